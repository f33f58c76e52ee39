"""Timing probes for the tower attention kernel (not a benchmark): cost per work item and per 128-key tile over a sweep of sequence
lengths.  During the v2 -> v3 work the kernel also had a VIDI_ATTN_DBG switch that disabled single pipeline stages (exp2, P V MMAs,
S loads, P store ...; results wrong by design) — those hooks are gone from the product kernel; the raw numbers they produced are in
profiles/r01_attn_probes_raw.txt (the `dbg` column is therefore always 0 now)."""
import json, os, sys, subprocess
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vidi_b200 import ops

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def run(B, S, H, dh, impl="auto"):
    qkv = (torch.randn(B * S, 3 * H * dh, device="cuda") * 1.0).to(torch.bfloat16)
    out = torch.empty(B * S, H * dh, device="cuda", dtype=torch.bfloat16)
    return timeit(lambda: ops.attn_dense(qkv, B, S, H, dh, dh ** -0.5, out=out, impl=impl))

if __name__ == "__main__":
    dbg = os.environ.get("VIDI_ATTN_DBG", "0")
    rows = []
    for dh, H in ((72, 16), (64, 20)):
        for S, B in ((256, 148 * 2), (768, 74), (1536, 37), (3072, 37), (729, 64), (1500, 16)):
            t = run(B, S, H, dh)
            nq2 = (S + 255) // 256; nk = (S + 127) // 128
            items = B * H * nq2
            per_sm_items = -(-items // 148)
            clk = t * 1e-3 * 1.75e9
            rows.append(dict(dbg=int(dbg), dh=dh, S=S, B=B, ms=round(t, 4), items=items, nk=nk,
                             clk_per_item=round(clk / per_sm_items), clk_per_tile=round(clk / per_sm_items / nk)))
    for r in rows: print(json.dumps(r))
