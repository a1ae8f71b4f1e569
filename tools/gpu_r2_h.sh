#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/link/run_link_test.py 2>&1 | tail -3
MB_FETCH_BULK=1 timeout 300 python tools/dbg_fetch.py > gpurun_out/r2h_dbg1.log 2>&1; echo "dbg bulk rc=$?"; tail -4 gpurun_out/r2h_dbg1.log | cut -c1-300
MB_FETCH_BULK=0 timeout 300 python tools/dbg_fetch.py > gpurun_out/r2h_dbg0.log 2>&1; echo "dbg nobulk rc=$?"; tail -2 gpurun_out/r2h_dbg0.log | cut -c1-300
MB_FETCH_BULK=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/dbg_fetch.py > gpurun_out/r2h_san.log 2>&1; echo "sanitizer rc=$?"; grep -v "^$" gpurun_out/r2h_san.log | grep -i "error\|invalid\|at 0x\|by thread\|=====" | head -20
