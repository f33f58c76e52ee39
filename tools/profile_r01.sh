#!/bin/bash
# ncu evidence for round 1 (run under gpurun; outputs land in gpurun_out/ and are summarised into profiles/)
set -x
mkdir -p gpurun_out
KRE='regex:^(gemm_bf16|attn_|xattn_|layernorm|rmsnorm|residual_norm|mm_finish|pool_s2d|patch_im2col|whisper_im2col|embed_gather|sinusoid|split3|cast_f32|rope)'
# (1) per-launch device time of the library's kernels over load + 1 warm-up + 1 timed c2 step
#     (cold-cache, serialised -> compare SHARES); the last gpu_launches rows are the timed step
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 6000 --csv --log-file gpurun_out/launches_c2.csv \
    python bench.py --workload c2 --steps 1 --quick > gpurun_out/launches_c2.stdout 2>&1
# (2) full captures of the hot kernels (third launch of each)
cap() {  # name kernel-regex case
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s 2 -c 1 -o gpurun_out/prof_$1 -f \
      python tools/bench_kernels.py one $3 > gpurun_out/prof_$1.stdout 2>&1
}
cap gemm_gateup gemm_bf16_kernel gate_up
cap gemm_vit_fc2 gemm_bf16_kernel vit_fc2
cap gemm_vit_qkv gemm_bf16_kernel vit_qkv
cap attn_vit attn_fwd_sm100 attn_vit
cap xattn xattn_splitkv xattn
cap residual_norm residual_norm_kernel residual_norm
ls -la gpurun_out
