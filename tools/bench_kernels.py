"""Micro-benchmarks of the hot kernels (CUDA events, L2-flushed between iterations) next to the library bar
(torch.matmul -> cuBLAS, flash-attn 2) the reference would run on the same B200.  Not the headline bench."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from vidi_b200 import ops  # noqa: E402
from vidi_b200.weights import pack_glu  # noqa: E402

BF = torch.bfloat16
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=8, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def gemm_case(name, M, N, K, glu=0, bn=None):
    a = torch.randn(M, K, device="cuda").to(BF)
    w = (torch.randn(N, K, device="cuda") * 0.02).to(BF)
    out = torch.empty(M, N // 2 if glu else N, device="cuda", dtype=BF)
    t = timeit(lambda: ops.gemm(a, w, out=out, glu=glu, block_n=bn, cta2=False))
    t2 = {}
    if M >= 1024:
        for b2 in (256, 192):
            try:
                t2[b2] = round(2.0 * M * N * K / timeit(lambda: ops.gemm(a, w, out=out, glu=glu, block_n=b2, cta2=True)) / 1e9, 1)
            except Exception as ex:  # noqa: BLE001
                t2[b2] = repr(ex)[:60]
    ref = torch.empty(M, N, device="cuda", dtype=BF)
    t_ref = timeit(lambda: torch.matmul(a, w.t(), out=ref))
    fl = 2.0 * M * N * K
    print(json.dumps(dict(kernel="gemm", name=name, M=M, N=N, K=K, glu=glu, ms=round(t, 4), tflops=round(fl / t / 1e9, 1),
                          cublas_ms=round(t_ref, 4), cublas_tflops=round(fl / t_ref / 1e9, 1), cta2_tflops=t2)), flush=True)


def xattn_case(T, N, splits=None):
    Hq, Hkv, dh = 16, 8, 256
    q = torch.randn(T, Hq * dh, device="cuda").to(BF)
    kv = torch.randn(N, 2 * Hkv * dh, device="cuda").to(BF)
    sp = splits or ops.xattn_splits(N, Hkv)
    op = torch.empty(sp, T, Hq, dh, device="cuda"); ls = torch.empty(sp, T, Hq, device="cuda")
    t = timeit(lambda: ops.xattn_splitkv(q, kv[:, :2048], kv[:, 2048:], None, Hq, Hkv, dh, 1 / 16, 50.0, sp, opart=op, lse=ls))
    by = N * 8192
    print(json.dumps(dict(kernel="xattn_splitkv", T=T, N=N, splits=sp, ms=round(t, 4), gbs=round(by / t / 1e6, 1))), flush=True)


def dense_case(B, S, H, dh):
    d = H * dh
    qkv = torch.randn(B * S, 3 * d, device="cuda").to(BF)
    out = torch.empty(B * S, d, device="cuda", dtype=BF)
    t = timeit(lambda: ops.attn_dense(qkv, B, S, H, dh, dh ** -0.5, out=out))
    fl = 4.0 * B * H * S * S * dh
    t_m = timeit(lambda: ops.attn_dense(qkv, B, S, H, dh, dh ** -0.5, out=out, impl="mma"))
    t_1 = timeit(lambda: ops.attn_dense(qkv, B, S, H, dh, dh ** -0.5, out=out, impl="v1"))
    t_2 = timeit(lambda: ops.attn_dense(qkv, B, S, H, dh, dh ** -0.5, out=out, impl="v2"))
    rec = dict(kernel="attn_dense", B=B, S=S, H=H, dh=dh, ms=round(t, 4), tflops=round(fl / t / 1e9, 1), v1_tflops=round(fl / t_1 / 1e9, 1), v2_tflops=round(fl / t_2 / 1e9, 1), mma_ms=round(t_m, 4),
               mma_tflops=round(fl / t_m / 1e9, 1))
    if True:                                                    # A/B variants: FMA-pipe exp2 on every N-th score pair
        for n in (2, 3, 4):
            t_p = timeit(lambda: ops.attn_dense(qkv, B, S, H, dh, dh ** -0.5, out=out, impl=f"poly{n}"))
            rec[f"poly{n}_tflops"] = round(fl / t_p / 1e9, 1)
    try:
        from flash_attn import flash_attn_func
        q, k, v = [x.reshape(B, S, H, dh) for x in qkv.split(d, 1)]
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        t2 = timeit(lambda: flash_attn_func(q, k, v))
        rec.update(fa2_ms=round(t2, 4), fa2_tflops=round(fl / t2 / 1e9, 1))
    except Exception as ex:  # noqa: BLE001
        rec["fa2"] = f"unavailable: {type(ex).__name__}"
    print(json.dumps(rec), flush=True)


def rownorm_case(rows, D):
    x = torch.randn(rows, D, device="cuda").to(BF); y = torch.randn(rows, D, device="cuda").to(BF)
    w = torch.randn(D, device="cuda").to(BF); h = torch.empty_like(x)
    t = timeit(lambda: ops.residual_norm(x, y, w, w, h, 1e-6, 1, True))
    print(json.dumps(dict(kernel="residual_norm", rows=rows, D=D, ms=round(t, 4), gbs=round(rows * D * 2 * 4 / t / 1e6, 1))), flush=True)


if __name__ == "__main__" and not (len(sys.argv) > 2 and sys.argv[1] == "one"):
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "gemm"):
        M = 15750
        gemm_case("kv_proj", M, 4096, 3584)
        gemm_case("v_o_fold", M, 3584, 2048)
        gemm_case("gate_up_geglu", M, 28672, 3584, glu=1)
        gemm_case("down", M, 3584, 14336)
        gemm_case("vit_qkv", 64 * 729, 3456, 1152)
        gemm_case("vit_fc1", 64 * 729, 4304, 1152)
        gemm_case("vit_fc2", 64 * 729, 1152, 4304)
        gemm_case("vit_out", 64 * 729, 1152, 1152)
        gemm_case("square8k", 8192, 8192, 8192)
        gemm_case("kv_proj_M126k", 126000, 4096, 3584)
        gemm_case("gate_up_M126k", 126000, 28672, 3584, glu=1)
        gemm_case("down_M126k", 126000, 3584, 14336)
        gemm_case("text_gateup", 32, 28672, 3584, glu=1)
        gemm_case("text_down", 32, 3584, 14336)
        gemm_case("text_down_bn64", 32, 3584, 14336, bn=64)
    if which in ("all", "text"):
        # text-stream GEMMs: weight-streaming (swap-AB) kernel vs the general kernel (block_n < 0 forces it), GB/s of W read
        for name, M, N, K, glu in (("qkv", 32, 8192, 3584, 0), ("o", 32, 3584, 4096, 0), ("gate_up", 32, 28672, 3584, 1), ("down", 32, 3584, 14336, 0),
                                   ("lm_head", 32, 256000, 3584, 0), ("qkv_dec", 1, 8192, 3584, 0), ("gate_up_dec", 1, 28672, 3584, 1),
                                   ("down_dec", 1, 3584, 14336, 0), ("lm_head_dec", 1, 256000, 3584, 0)):
            a = torch.randn(M, K, device="cuda").to(BF); w = (torch.randn(N, K, device="cuda") * 0.02).to(BF)
            out = torch.empty(M, N // 2 if glu else N, device="cuda", dtype=BF)
            t_s = timeit(lambda: ops.gemm(a, w, out=out, glu=glu, cta2=False))
            t_g = timeit(lambda: ops.gemm(a, w, out=out, glu=glu, cta2=False, block_n=-(256 if glu else 64)))
            print(json.dumps(dict(kernel="gemm_text", name=name, M=M, N=N, K=K, skinny_us=round(t_s * 1e3, 1), skinny_gbs=round(N * K * 2 / t_s / 1e6, 1),
                                  general_us=round(t_g * 1e3, 1), general_gbs=round(N * K * 2 / t_g / 1e6, 1))), flush=True)
    if which in ("all", "attn"):
        for sp in (18, 37, 55, 74):
            xattn_case(32, 126000, sp)
        xattn_case(32, 16000)
        xattn_case(32, 15750, 18)
        dense_case(64, 729, 16, 72)
        dense_case(16, 1500, 20, 64)
        rownorm_case(126000, 3584)


def one(name):
    """`python tools/bench_kernels.py one <case>`: a few plain launches of one kernel, for ncu captures."""
    cases = {
        "gate_up": lambda: gemm_launcher(15750, 28672, 3584, 1),
        "kv_proj": lambda: gemm_launcher(15750, 4096, 3584, 0),
        "vit_fc2": lambda: gemm_launcher(46656, 1152, 4304, 0),
        "vit_qkv": lambda: gemm_launcher(46656, 3456, 1152, 0),
        "gate_up126k": lambda: gemm_launcher(126000, 28672, 3584, 1, cta2=False),
        "gate_up126k_2cta": lambda: gemm_launcher(126000, 28672, 3584, 1, cta2=True),
        "vit_fc2_2cta_res": lambda: gemm_res_launcher(93312, 1152, 4304),
        "text_down": lambda: gemm_launcher(32, 3584, 14336, 0, cta2=False),
        "text_gate_up": lambda: gemm_launcher(32, 28672, 3584, 1, cta2=False),
    }
    if name in cases:
        fn = cases[name]()
    elif name == "attn_vit":
        qkv = torch.randn(64 * 729, 3 * 1152, device="cuda").to(BF); out = torch.empty(64 * 729, 1152, device="cuda", dtype=BF)
        fn = lambda: ops.attn_dense(qkv, 64, 729, 16, 72, 72 ** -0.5, out=out)
    elif name == "xattn":           # one launch over the image (90 000 keys) and audio (36 000 keys) segments, as the text pass issues it
        q = torch.randn(32, 4096, device="cuda").to(BF); kv = torch.randn(126000, 4096, device="cuda").to(BF)
        sp = ops.xattn_split_plan([90000, 36000], 8)
        op = torch.empty(sum(sp) * 32 * 16 * 256, device="cuda"); ls = torch.empty(sum(sp) * 32 * 16, device="cuda")
        fn = lambda: ops.xattn_splitkv_seg(q, kv[:, :2048], kv[:, 2048:], [(0, 90000, None), (90000, 36000, None)], sp, 16, 8, 256, 1 / 16, 50.0, op, ls)
    elif name == "layernorm":
        x = torch.randn(128 * 729, 1152, device="cuda").to(BF); w = torch.ones(1152, device="cuda"); b = torch.zeros(1152, device="cuda")
        y = torch.empty_like(x)
        fn = lambda: ops.layernorm(x, w, b, 1e-6, out=y)
    elif name == "residual_norm":
        x = torch.randn(126000, 3584, device="cuda").to(BF); y = torch.randn(126000, 3584, device="cuda").to(BF)
        w = torch.randn(3584, device="cuda").to(BF); h = torch.empty_like(x)
        fn = lambda: ops.residual_norm(x, y, w, w, h, 1e-6, 1, True)
    else:
        raise SystemExit(f"unknown case {name}")
    for _ in range(5):
        fn()
    torch.cuda.synchronize()


def gemm_res_launcher(M, N, K):
    """in-place residual GEMM of a tower block (x += a W^T + b): TMA-store epilogue with the residual rows fetched by TMA"""
    a = torch.randn(M, K, device="cuda").to(BF); w = (torch.randn(N, K, device="cuda") * 0.02).to(BF)
    x = torch.randn(M, N, device="cuda").to(BF); b = torch.randn(N, device="cuda") * 0.02
    return lambda: ops.gemm(a, w, bias=b, residual=x, out=x, cta2=True)


def gemm_launcher(M, N, K, glu, cta2=None):
    a = torch.randn(M, K, device="cuda").to(BF); w = (torch.randn(N, K, device="cuda") * 0.02).to(BF)
    out = torch.empty(M, N // 2 if glu else N, device="cuda", dtype=BF)
    return lambda: ops.gemm(a, w, out=out, glu=glu, cta2=cta2)


if len(sys.argv) > 2 and sys.argv[1] == "one":
    one(sys.argv[2])
