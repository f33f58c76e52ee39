"""world_size=2 gloo test of the multi-rank host logic: shard plan + the per-rank pre-merge of the key splits (xchg.cu
xattn_premerge_push) + the exchange of the reduced [O | LSE] blocks + the rank-strided log-sum-exp merge indexing that
engine._TextRun hands to vidi_xattn_merge2 (spr = 1, rank stride = block).  The per-rank partials are produced by the oracle's
attention on CPU; the CUDA kernels' own parity, and the engine's multi-rank branches, are covered by the -m gpu tests
(test_kernels_gpu.py::test_premerge_then_rank_strided_merge2, test_engine_gpu.py::test_multirank_text_path_lockstep_equals_single_rank,
test_dist_nccl_gpu.py)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def merge_like_kernel(gathered, P, spr, rank_stride, o_off, l_off, rows, dh):
    """Python mirror of xattn_merge_kernel's addressing (attn.cu)."""
    out = torch.zeros(rows, dh)
    lse = torch.stack([gathered[(p // spr) * rank_stride + l_off + (p % spr) * rows:][:rows] for p in range(P)])
    L = lse.max(0).values
    w = torch.where(torch.isinf(lse), torch.zeros_like(lse), torch.exp(lse - L))
    for p in range(P):
        base = (p // spr) * rank_stride + o_off + (p % spr) * rows * dh
        out += w[p][:, None] * gathered[base:base + rows * dh].view(rows, dh)
    return out / w.sum(0)[:, None]


def premerge_like_kernel(O, Ls):
    """Python mirror of xattn_premerge_push_kernel: [P, rows, dh], [P, rows] -> flat [O rows*dh | LSE rows]"""
    Lm = Ls.max(0).values
    w = torch.where(torch.isinf(Ls), torch.zeros_like(Ls), torch.exp(Ls - torch.where(torch.isinf(Lm), torch.zeros_like(Lm), Lm)))
    den = w.sum(0)
    o = (w[:, :, None] * O).sum(0) * torch.where(den > 0, 1.0 / den, torch.zeros_like(den))[:, None]
    lse = torch.where(den > 0, Lm + torch.log(den), torch.full_like(den, float("-inf")))
    return torch.cat([o.reshape(-1), lse])


def worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vidi_b200.config import vidi15_mini
    from vidi_b200.engine import make_plan
    cfg = vidi15_mini()
    c = cfg.llm
    T, F, Cn, asz, spr = 5, 5, 3, 7000, 2
    plan = make_plan(cfg, F, Cn, asz, rank, world)
    g = torch.Generator().manual_seed(7)
    q = torch.randn(c.heads, T, c.head_dim, generator=g)
    K_img = torch.randn(c.kv_heads, plan.n_img_total, c.head_dim, generator=g)
    V_img = torch.randn(c.kv_heads, plan.n_img_total, c.head_dim, generator=g)
    K_aud = torch.randn(c.kv_heads, plan.n_aud_total, c.head_dim, generator=g)
    V_aud = torch.randn(c.kv_heads, plan.n_aud_total, c.head_dim, generator=g)
    rows, dh = T * c.heads, c.head_dim

    def partials(K, V, lo, hi):
        """spr key-splits of this rank's shard -> ([spr, rows, dh], [spr, rows]) like xattn_splitkv."""
        O, Ls = [], []
        n = hi - lo
        step = max(1, -(-n // spr))
        for s in range(spr):
            a, b = lo + s * step, min(hi, lo + (s + 1) * step)
            if b <= a:
                O.append(torch.zeros(rows, dh)); Ls.append(torch.full((rows,), float("-inf"))); continue
            k = K[:, a:b].repeat_interleave(c.groups, 0); v = V[:, a:b].repeat_interleave(c.groups, 0)
            s_ = 50 * torch.tanh((q @ k.transpose(-1, -2)) / 16 / 50)
            Ls.append(torch.logsumexp(s_, -1).transpose(0, 1).reshape(rows))       # row = t*Hq + h
            O.append((torch.softmax(s_, -1) @ v).transpose(0, 1).reshape(rows, dh))
        return torch.stack(O), torch.stack(Ls)

    Oi, Li = partials(K_img, V_img, plan.f0 * plan.tpf, plan.f1 * plan.tpf)
    Oa, La = partials(K_aud, V_aud, plan.a0, plan.a1)
    flat = torch.cat([premerge_like_kernel(Oi, Li), premerge_like_kernel(Oa, La)])     # one reduced block per rank
    gathered = torch.empty(world * flat.numel())
    dist.all_gather_into_tensor(gathered, flat)
    sz = rows * (dh + 1)
    out_i = merge_like_kernel(gathered, world, 1, flat.numel(), 0, rows * dh, rows, dh)
    out_a = merge_like_kernel(gathered, world, 1, flat.numel(), sz, sz + rows * dh, rows, dh)

    def full(K, V):
        k = K.repeat_interleave(c.groups, 0); v = V.repeat_interleave(c.groups, 0)
        s_ = 50 * torch.tanh((q @ k.transpose(-1, -2)) / 16 / 50)
        return (torch.softmax(s_, -1) @ v).transpose(0, 1).reshape(rows, dh)
    ok = torch.allclose(out_i, full(K_img, V_img), atol=1e-5) and torch.allclose(out_a, full(K_aud, V_aud), atol=1e-5)
    t = torch.tensor([1.0 if ok else 0.0]); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(bool(t.item() == 1.0))
    dist.destroy_process_group()


def test_two_rank_partial_exchange_gloo():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=worker, args=(r, 2, port, ret)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert ret.get(timeout=5) is True
