#!/bin/bash
# ncu full captures of the round-1 final kernels (run under gpurun)
set -x
mkdir -p gpurun_out
cap() {  # name kernel-regex case
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s 2 -c 1 -o gpurun_out/prof2_$1 -f \
      python tools/bench_kernels.py one $3 > gpurun_out/prof2_$1.stdout 2>&1
}
cap attn_pp attn_fwd2_sm100 attn_vit
cap xattn xattn_splitkv_sm100 xattn
cap gemm2_vit_qkv gemm2_bf16_kernel vit_qkv
cap gemm2_vit_fc2 gemm2_bf16_kernel vit_fc2
cap gemm_gateup126k gemm_bf16_kernel gate_up126k
cap layernorm layernorm_kernel layernorm
ls -la gpurun_out | grep prof2
