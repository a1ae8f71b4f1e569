#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/f3_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/f3_pytest.log
python -c "
import __graft_entry__ as g
g.smoke(); print('smoke ok')
" 2>&1 | tail -3
