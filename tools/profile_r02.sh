#!/bin/bash
# Round-2 evidence run (1 GPU, under gpurun): full GPU test suite, smoke, the bench line, the ncu launch list of one C3 step and
# `ncu --set full` captures of the kernels whose numbers DESIGN.md quotes.  Outputs land in gpurun_out/ and are summarised into profiles/.
mkdir -p gpurun_out
VIDI_EVIDENCE_DIR=gpurun_out timeout 1500 python -m pytest tests -m gpu -q --timeout 400 --durations=8 > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -6 gpurun_out/r02_pytest_gpu_final.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_c3_n1_final.json 2> gpurun_out/r02_bench_c3_n1_final.err; tail -c 1500 gpurun_out/r02_bench_c3_n1_final.json
KRE='regex:^(gemm|attn_|xattn_|layernorm|rmsnorm|residual_norm|mm_finish|pool_s2d|patch_im2col|whisper_im2col|embed_gather|sinusoid|split3|cast_f32|rope|text_qk)'
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 40000 --csv --log-file gpurun_out/r02_launches_c3.csv \
    python bench.py --workload c3 --steps 1 --quick --no-cpu-baseline > gpurun_out/r02_launches_c3.stdout 2>&1
cap() {  # name kernel-regex case
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$2" -s 2 -c 1 -o gpurun_out/r02_prof_$1 -f \
      python tools/bench_kernels.py one $3 > gpurun_out/r02_prof_$1.stdout 2>&1
}
cap gemm_gateup126k gemm_bf16_kernel gate_up126k
cap xattn_seg xattn_splitkv_sm100 xattn
cap gemm_skinny_down gemm_skinny_kernel text_down
cap attn_vit attn_fwd3_sm100 attn_vit
ls -la gpurun_out | grep r02_prof
