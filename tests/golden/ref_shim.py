"""Import shim that lets the *reference's own* Python modules run on this CPU-only container.

Used ONLY by tests/golden/make_golden.py to generate committed fixtures (the reference tree does not
exist on the GPU box).  The reference's source is imported from /root/reference unmodified; what is
stubbed is third-party plumbing it cannot get here:
  * deepspeed / decord / langid / orjson            -- absent packages, import-time only
  * liger_kernel monkey patch                        -- Triton RoPE kernel, GPU only
  * transformers.cache_utils.HybridCache             -- removed in transformers 5.x (only constructed in generate)
  * is_flash_attn_2_available()                      -- asserts at import; flash_attn_func itself is replaced by an
                                                        eager fp32 restatement of FA2's documented semantics
"""
import importlib.machinery
import sys
import types

import torch

REF_ROOT = "/root/reference/Vidi1.5_9B"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def eager_flash_attn_func(q, k, v, dropout_p=0.0, softmax_scale=None, causal=False, softcap=0.0, deterministic=False, **kw):
    """flash_attn_func semantics: q,k,v [B,S,H,D] (k/v heads may divide q heads), returns [B,Sq,H,D]."""
    B, Sq, H, D = q.shape
    Hk = k.shape[2]
    qf, kf, vf = q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2)
    if Hk != H:
        kf = kf.repeat_interleave(H // Hk, 1); vf = vf.repeat_interleave(H // Hk, 1)
    scale = softmax_scale if softmax_scale is not None else D ** -0.5
    s = qf @ kf.transpose(-1, -2) * scale
    if softcap and softcap > 0:
        s = softcap * torch.tanh(s / softcap)
    if causal:
        Sk = k.shape[1]
        i = torch.arange(Sq)[:, None] + (Sk - Sq); j = torch.arange(Sk)[None, :]
        s = s.masked_fill(j > i, float("-inf"))
    return (torch.softmax(s, -1) @ vf).transpose(1, 2).to(q.dtype)


def install():
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    class _Comm(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return lambda *a, **k: None
    comm = _Comm("deepspeed.comm")
    comm.__spec__ = importlib.machinery.ModuleSpec("deepspeed.comm", None)
    comm.ProcessGroup = object
    ds = _stub("deepspeed")
    ds.comm = comm
    sys.modules["deepspeed.comm"] = comm
    _stub("liger_kernel")
    _stub("liger_kernel.transformers")
    _stub("liger_kernel.transformers.monkey_patch", apply_liger_kernel_to_gemma2=lambda **kw: None,
          LigerRMSNorm=type("LigerRMSNorm", (torch.nn.Module,), {}),
          __all__=["apply_liger_kernel_to_gemma2", "LigerRMSNorm"])
    _stub("langid")
    _stub("orjson")
    _stub("decord", VideoReader=object, cpu=lambda *a, **k: None)

    import transformers  # noqa: F401
    import transformers.cache_utils as cu
    import transformers.utils as tu
    if not hasattr(cu, "HybridCache"):
        class HybridCache(cu.DynamicCache):
            pass
        cu.HybridCache = HybridCache
    tu.is_flash_attn_2_available = lambda: True
    tu.is_flash_attn_greater_or_equal = lambda v: True
    import vidi.model.lmm.dattn.gemma as G           # noqa: E402  (reference code, unmodified)
    import vidi.model.lmm.dattn.xattn as X
    X.flash_attn_func = eager_flash_attn_func
    return G, X


if __name__ == "__main__":
    G, X = install()
    print("reference modules imported:", G.__file__)


def eager_flash_attn_varlen_func(q, k, v, cu_seqlens_q=None, cu_seqlens_k=None, max_seqlen_q=None, max_seqlen_k=None,
                                 dropout_p=0.0, softmax_scale=None, causal=False, softcap=0.0, deterministic=False, **kw):
    """flash_attn_varlen_func semantics: packed q [Tq,H,D], k/v [Tk,Hk,D], per-sequence boundaries in cu_seqlens_*."""
    outs = []
    for b in range(len(cu_seqlens_q) - 1):
        qs, qe = int(cu_seqlens_q[b]), int(cu_seqlens_q[b + 1])
        ks, ke = int(cu_seqlens_k[b]), int(cu_seqlens_k[b + 1])
        outs.append(eager_flash_attn_func(q[None, qs:qe], k[None, ks:ke], v[None, ks:ke], softmax_scale=softmax_scale,
                                          causal=causal, softcap=softcap)[0])
    return torch.cat(outs, 0)
