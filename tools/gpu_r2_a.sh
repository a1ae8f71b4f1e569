#!/bin/bash
# round-2 GPU pass A: tests, bench (with parity block), ncu at the bench shape (application replay: no save/restore of the 128 GB table)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2a_bench.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum
MB_GRAPH=0 timeout 900 ncu --replay-mode application --metrics $M --clock-control none \
   -k regex:'rows_kernel|edge_backward|segment_reduce_kernel|loss_kernel|gemm_tc_group' -s 21 -c 7 --csv --log-file gpurun_out/r2a_ncu_bench_shape.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2a_ncu.out 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2a_ncu.out
