"""Sustained micro-benchmark of one SigLIP / Whisper encoder block at the in-step shapes (128 frames = 93 312 rows; 16 chunks = 24 000 rows):
the block's ops run back to back for `--reps` x `--layers` iterations, exactly as engine._tower_layer issues them, so the chip is in the
same power-capped steady state as inside a bench step (single isolated launches overstate these K = 1152/1280 GEMMs by 10-30 %).
Per-site CUDA-event timings -> TF/s per GEMM site, GB/s for LayerNorm, plus the SM clock seen by nvidia-smi.  A/B knobs come from the
environment (read by the launchers): VIDI_GEMM2_RELAXED=1, VIDI_GEMM2_TMASTORE=1 ...   Not the headline bench."""
import argparse
import json
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
from vidi_b200 import ops  # noqa: E402

BF = torch.bfloat16


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tower", default="vit", choices=["vit", "aud"])
    ap.add_argument("--layers", type=int, default=26)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--cta2", type=int, default=1)
    ap.add_argument("--no-attn", action="store_true")
    ap.add_argument("--tag", default="")
    ap.add_argument("--bn", type=int, default=0, help="force the column tile of the four GEMMs (0 = ops.pick_block_n)")
    ap.add_argument("--frames", type=int, default=0, help="frames (vit) / chunks (aud) per block; default 128 / 16")
    a = ap.parse_args()
    if a.tower == "vit":
        B, S, H, dh, d, ff, act = 128, 729, 16, 72, 1152, 4304, ops.ACT_GELU_TANH
    else:
        B, S, H, dh, d, ff, act = 16, 1500, 20, 64, 1280, 5120, ops.ACT_GELU_ERF
    if a.frames:
        B = a.frames
    M = B * S
    bn = a.bn or None
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s, sc=1.0: (torch.randn(*s, device="cuda", generator=g) * sc)
    x = rn(M, d).to(BF)
    wqkv, wo, w1, w2 = rn(3 * d, d, sc=0.02).to(BF), rn(d, d, sc=0.02).to(BF), rn(ff, d, sc=0.02).to(BF), rn(d, ff, sc=0.02).to(BF)
    bqkv, bo, b1, b2 = rn(3 * d, sc=0.02), rn(d, sc=0.02), rn(ff, sc=0.02), rn(d, sc=0.02)
    lw, lb = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
    h = torch.empty_like(x)
    qkv = torch.empty(M, 3 * d, device="cuda", dtype=BF)
    att = rn(M, d).to(BF)
    m = torch.empty(M, ff, device="cuda", dtype=BF)
    cta2 = bool(a.cta2)

    def block():
        ops.layernorm(x, lw, lb, 1e-6, out=h)
        ops.gemm(h, wqkv, bias=bqkv, out=qkv, tag="qkv", cta2=cta2, block_n=bn)
        if not a.no_attn:
            ops.attn_dense(qkv, B, S, H, dh, dh ** -0.5, out=att)
        ops.gemm(att, wo, bias=bo, residual=x, out=x, tag="out", cta2=cta2, block_n=bn)
        ops.layernorm(x, lw, lb, 1e-6, out=h)
        ops.gemm(h, w1, bias=b1, act=act, out=m, tag="fc1", cta2=cta2, block_n=bn)
        ops.gemm(m, w2, bias=b2, residual=x, out=x, tag="fc2", cta2=cta2, block_n=bn)

    for _ in range(4):
        block()
    torch.cuda.synchronize()
    clocks = []
    stop = False

    def sample():
        import pynvml as nv
        nv.nvmlInit()
        hd = nv.nvmlDeviceGetHandleByIndex(0)
        while not stop:
            try:
                clocks.append((float(nv.nvmlDeviceGetClockInfo(hd, nv.NVML_CLOCK_SM)), nv.nvmlDeviceGetPowerUsage(hd) / 1000.0))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)
    th = threading.Thread(target=sample, daemon=True)
    th.start()
    ops.PROFILE, ops.PROFILE_OPS = [], {}
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.reps * a.layers):
        block()
    e.record()
    torch.cuda.synchronize()
    stop = True
    total = s.elapsed_time(e)
    by = {}
    for tag, fl, e0, e1 in ops.PROFILE:
        d_ = by.setdefault(tag, [0.0, 0.0, 0])
        d_[0] += fl; d_[1] += e0.elapsed_time(e1); d_[2] += 1
    other = {k: round(sum(p.elapsed_time(q) for p, q in v), 2) for k, v in ops.PROFILE_OPS.items()}
    ops.PROFILE, ops.PROFILE_OPS = None, None
    gemm_fl, gemm_ms = sum(v[0] for v in by.values()), sum(v[1] for v in by.values())
    clocks.sort()
    print(json.dumps(dict(tag=a.tag, tower=a.tower, cta2=cta2, rows=M, blocks=a.reps * a.layers, total_ms=round(total, 2),
                          gemm_tflops=round(gemm_fl / gemm_ms / 1e9, 1),
                          sites={k: dict(tflops=round(v[0] / v[1] / 1e9, 1), us=round(v[1] / v[2] * 1e3, 1)) for k, v in sorted(by.items())},
                          other_ms=other, ln_gbs=round(2 * a.reps * a.layers * M * d * 4 / other.get("layernorm", 1e9) / 1e6, 1),
                          sm_mhz_median=clocks[len(clocks) // 2][0] if clocks else None,
                          power_w_median=sorted(c[1] for c in clocks)[len(clocks) // 2] if clocks else None)), flush=True)


if __name__ == "__main__":
    main()
