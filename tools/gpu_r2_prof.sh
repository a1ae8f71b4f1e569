#!/bin/bash
# profiling pass: every kernel of the bench-shape steps (4e7-row table), application replay (the process is re-run per pass: nothing is saved / restored)
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum
MB_GRAPH=0 timeout 1200 ncu --replay-mode application --metrics $M --clock-control none -c 600 --csv --log-file gpurun_out/r2_ncu_step.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-extra-shapes --no-buffered > gpurun_out/r2_ncu_step.out 2>&1; echo "ncu step rc=$?"
grep -c gpu__time gpurun_out/r2_ncu_step.csv; tail -2 gpurun_out/r2_ncu_step.out | cut -c1-200
