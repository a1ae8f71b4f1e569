#!/bin/bash
# compute-sanitizer memcheck over the backward contraction kernel tests (tensor-memory-A / shared-memory-A, identity conversion, bf16x3 and single bf16)
mkdir -p gpurun_out
timeout 150 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_decoder.py -m gpu -q -k "kernels_vs_fp64 or single_bf16" > gpurun_out/r2_sanitizer_backward.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2_sanitizer_backward.log
