mkdir -p gpurun_out
VIDI_EVIDENCE_DIR=gpurun_out timeout 1500 python -m pytest tests -m gpu -q --timeout 400 --durations=6 > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -12 gpurun_out/r02_pytest_gpu_final.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
cap() {  # name kernel-regex case
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$2" -s 2 -c 1 -o gpurun_out/r02_prof_$1 -f \
      python tools/bench_kernels.py one $3 > gpurun_out/r02_prof_$1.stdout 2>&1
}
cap gemm2_gateup126k gemm2_bf16_kernel gate_up126k_2cta
cap gemm2_vit_qkv_ts gemm2_bf16_kernel vit_qkv
cap gemm2_vit_fc2_res gemm2_bf16_kernel vit_fc2_2cta_res
ls -la gpurun_out | grep "r02_prof_gemm2"
