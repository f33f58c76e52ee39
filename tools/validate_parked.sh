#!/bin/bash
# First GPU call of the next round: run everything that was written after round 1's GPU budget was spent.
#   gpurun --timeout 600 -- 'bash tools/validate_parked.sh'
export VIDI_RUN_UNVALIDATED=1
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_preprocess_gpu.py -m gpu -q > gpurun_out/parked_tests.log 2>&1; tail -15 gpurun_out/parked_tests.log
timeout 120 python tools/bench_kernels.py attn 2>&1 | grep attn_dense | tee gpurun_out/parked_attn_poly.log | cut -c1-400
