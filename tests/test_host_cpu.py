"""CPU-only checks: the C-ABI library loads and exports every declared symbol, host-side shard/token logic,
weight repacking algebra, and the facade's error behaviour.  No kernel is launched here."""
import ctypes
import math
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from vidi_b200 import lib
    hdr = open(os.path.join(ROOT, "include", "vidi_b200.h")).read()
    declared = set(re.findall(r"\b(vidi_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    L = lib.load()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/vidi_b200.h but not exported"
    assert declared == set(lib.SIGNATURES) | set(lib.EXTRA_SYMBOLS)
    assert L.vidi_abi_version() == 2
    assert isinstance(L.vidi_launch_count(), int)


def test_ops_refuse_cpu_tensors():
    from vidi_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_engine_fails_loudly_without_gpu():
    from vidi_b200.config import vidi15_mini
    from vidi_b200.engine import Vidi15Engine
    with pytest.raises(RuntimeError, match="no CPU path"):
        Vidi15Engine(vidi15_mini(), {}, device="cuda")


def test_token_math_follows_reference():
    """multimodal.py:175-180 + utils.py:152-171: 196 tokens/frame up to 306 frames, then shrink, floor 10x10."""
    from vidi_b200.config import vidi15_9b
    c = vidi15_9b()
    assert c.image_tokens(8) == 1568 and c.image_tokens(80) == 15680 and c.image_tokens(306) == 306 * 196
    assert c.image_hw(307) != (28, 28)
    assert c.image_tokens(600) == 60000 and c.image_tokens(3600) == 90000 and c.image_tokens(1800) == 45000
    assert c.audio_tokens(800) == 80 and c.audio_tokens(360000) == 36000 and c.audio_tokens(1234) == 123
    assert c.vis.run_layers == 26 and c.vis.patches == 729


@pytest.mark.parametrize("F,C,asz,world", [(3600, 120, 360000, 8), (80, 3, 8000, 2), (5, 3, 7000, 3), (7, 1, 500, 4), (600, 20, 60000, 8)])
def test_shard_plan_partitions_exactly(F, C, asz, world):
    from vidi_b200.config import vidi15_9b
    from vidi_b200.engine import make_plan
    cfg = vidi15_9b()
    plans = [make_plan(cfg, F, C, asz, r, world) for r in range(world)]
    assert plans[0].f0 == 0 and plans[-1].f1 == F and plans[0].c0 == 0 and plans[-1].c1 == C
    for a, b in zip(plans, plans[1:]):
        assert a.f1 == b.f0 and a.c1 == b.c0 and a.a1 == b.a0
    assert sum(p.n_img for p in plans) == cfg.image_tokens(F) == plans[0].n_img_total
    assert sum(p.n_aud for p in plans) == cfg.audio_tokens(asz) == plans[0].n_aud_total
    assert max(p.f1 - p.f0 for p in plans) - min(p.f1 - p.f0 for p in plans) <= 1
    for p in plans:                                   # every audio token of a rank comes from that rank's chunks
        assert p.c0 * p.tpc <= p.a0 or p.n_aud == 0
        assert p.a1 <= p.c1 * p.tpc


def test_fold_o_proj_equals_repeat_kv_then_project():
    """K13: o_proj(repeat_kv(V)) == V @ W_o'^T with W_o' = sum over the group axis (gemma.py:77-78,195-197)."""
    from vidi_b200.weights import fold_o_proj
    D, Hkv, G, dh, N = 48, 2, 3, 8, 11
    Wo = torch.randn(D, Hkv * G * dh, dtype=torch.float64)
    V = torch.randn(N, Hkv * dh, dtype=torch.float64)
    Vrep = V.view(N, Hkv, dh).repeat_interleave(G, 1).reshape(N, -1)
    ref = Vrep @ Wo.t()
    out = V @ fold_o_proj(Wo, Hkv, G, dh).double().t()
    assert torch.allclose(out, ref, atol=1e-5)


def test_pack_glu_layout():
    from vidi_b200.weights import pack_glu
    I, K, bn = 512, 16, 256
    wg = torch.arange(I * K, dtype=torch.float32).view(I, K); wu = -wg
    p = pack_glu(wg, wu, bn)
    h = bn // 2
    for t in range(I // h):
        assert torch.equal(p[t * bn:t * bn + h], wg[t * h:(t + 1) * h])
        assert torch.equal(p[t * bn + h:(t + 1) * bn], wu[t * h:(t + 1) * h])


def test_projector_column_permutation_matches_pool_kernel_order():
    """weights.py permutes W1 columns from the reference's c*m*m+q order (utils.py:143-150) to the pool kernel's q*d+c."""
    from oracle import vidi15_ref as R
    d, m = 6, 2
    x = torch.randn(1, d, 4, 4)
    ref_order = R.space_to_depth(x, m).permute(0, 2, 3, 1)                       # [...,(c, q)]
    ker_order = ref_order.reshape(1, 2, 2, d, m * m).permute(0, 1, 2, 4, 3).reshape(1, 2, 2, m * m * d)
    W = torch.randn(5, d * m * m)
    Wp = W.reshape(5, d, m * m).permute(0, 2, 1).reshape(5, m * m * d)
    assert torch.allclose(ref_order @ W.t(), ker_order @ Wp.t(), atol=1e-5)


def test_facade_errors_and_sentinel_handling():
    from vidi_b200.model import DattnGemma2ForCausalLM, IMAGE_TOKEN_INDEX, config_from_hf_json
    ids = torch.tensor([2, IMAGE_TOKEN_INDEX, 5, 6, 7])
    assert DattnGemma2ForCausalLM._strip(ids, None).tolist() == [2, 5, 6, 7]
    assert DattnGemma2ForCausalLM._strip(ids, torch.tensor([1, 1, 1, 0, 1])).tolist() == [2, 5, 7]
    with pytest.raises(AssertionError, match="at most one image"):
        DattnGemma2ForCausalLM._strip(torch.tensor([IMAGE_TOKEN_INDEX, IMAGE_TOKEN_INDEX, 3]), None)
    cfg = config_from_hf_json({"hidden_size": 3584, "mm_image_pool_size": 2, "mm_audio_pool_size": 5})
    assert cfg.llm.layers == 42 and cfg.llm.kv_dim == 2048 and cfg.mm_time_interval == 10000


def test_postprocess_matches_ask_arithmetic():
    """inference.py:52-66: regex (\\d\\.\\d+)-(\\d\\.\\d+), t = frac * length, HH:MM:SS with int() truncation."""
    from vidi_b200.postprocess import format_time_ranges, parse_ranges
    from vidi_b200.pipeline import DEFAULT_IMAGE_TOKEN, PROMPT_VIDI15
    txt = " 0.10-0.25, 0.5-0.75 and 0.9000-1.0 "
    assert parse_ranges(txt) == [(0.10, 0.25), (0.5, 0.75), (0.9, 1.0)]
    assert format_time_ranges(txt, 3600.0) == "00:06:00-00:15:00, 00:30:00-00:45:00, 00:54:00-01:00:00"
    assert format_time_ranges("0.01-0.02", 25.0) == "00:00:00-00:00:00"
    assert format_time_ranges("no ranges here", 100.0) == ""
    q = "a dog running."
    assert DEFAULT_IMAGE_TOKEN + "\n" + PROMPT_VIDI15.format(q[:-1]) == "<image>\nDuring which time segments in the video can we see a dog running?"


def test_fold_layernorm_algebra():
    """weights.fold_layernorm: rstd * (x W'^T - mean * colsum(W')) + b'  ==  Linear(LayerNorm(x))  (what gemm_ln's epilogue applies)."""
    import torch
    import torch.nn.functional as F
    from vidi_b200.weights import fold_layernorm
    g = torch.Generator().manual_seed(5)
    M, D, N = 37, 96, 40
    x = torch.randn(M, D, generator=g) * 2 + 1.5
    w = (torch.randn(N, D, generator=g) * 0.1).to(torch.bfloat16); b = torch.randn(N, generator=g)
    gamma = 1 + 0.2 * torch.randn(D, generator=g); beta = 0.3 * torch.randn(D, generator=g)
    wp, cs, bp = fold_layernorm(w, b, gamma, beta)
    assert wp.dtype == torch.bfloat16 and cs.dtype == torch.float32 and bp.dtype == torch.float32
    mean = x.mean(1, keepdim=True); rstd = torch.rsqrt(x.var(1, unbiased=False, keepdim=True) + 1e-6)
    y = rstd * (x @ wp.float().t() - mean * cs[None, :]) + bp
    ref = F.layer_norm(x, (D,), gamma, beta, 1e-6) @ w.float().t() + b
    assert float((y - ref).norm() / ref.norm()) < 4e-3          # only the bf16 rounding of W*gamma differs


def test_bench_reference_arm_prints_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys, rank 0 only."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "impl", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "tokens/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"] == dict(value=line["value"], unit=line["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert "workload" in line["config"]
    # every other rank exits without work
    env["RANK"] = "1"; env["WORLD_SIZE"] = "2"; env["LOCAL_RANK"] = "1"
    out1 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, timeout=120, cwd=root, env=env)
    assert out1.returncode == 0 and out1.stdout.strip() == ""


def test_header_is_plain_c99_and_links_from_c(tmp_path):
    """include/vidi_b200.h is a C header (no C++ / torch types): a strict-C99 translation unit includes it, links the in-tree shared
    library and calls an entry point that needs no GPU."""
    import os
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "vidi_b200")
    exe = str(tmp_path / "abi_smoke")
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
                         os.path.join(root, "tests", "abi_smoke.c"), "-o", exe, "-L", libdir, "-lvidi_b200", f"-Wl,-rpath,{libdir}",
                         "-Wl,--allow-shlib-undefined"], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0 and run.stdout.startswith("abi="), (run.stdout, run.stderr)


def test_text_pass_descriptor_layout_matches_the_c_compiler(tmp_path):
    """lib.VidiTextPass / VidiTextLayerW / VidiTextSeg (ctypes) must have exactly the layout gcc gives the structs of
    include/vidi_b200.h -- the descriptor crosses the C ABI by pointer."""
    import ctypes as C
    import subprocess
    from vidi_b200.lib import VidiTextLayerW, VidiTextPass, VidiTextSeg
    fields = [n for n, _ in VidiTextPass._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "vidi_b200.h"\nint main(void){\n'
                   'printf("%zu %zu %zu\\n", sizeof(VidiTextLayerW), sizeof(VidiTextSeg), sizeof(VidiTextPass));\n'
                   + "".join(f'printf("%zu\\n", offsetof(VidiTextPass, {n}));\n' for n in fields) + "return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert [int(x) for x in out[:3]] == [C.sizeof(VidiTextLayerW), C.sizeof(VidiTextSeg), C.sizeof(VidiTextPass)]
    assert [int(x) for x in out[3:]] == [getattr(VidiTextPass, n).offset for n in fields]


def test_xattn_split_plan_properties():
    """key-split plan of the one-launch cross attention: both segments together fill the SMs once (<= n_sms // kv_heads splits),
    every segment keeps at least one split and at most one per 512 keys, and the plan is the same on every rank (it is computed from
    the per-rank share of the GLOBAL key counts)."""
    from vidi_b200 import ops
    assert ops.xattn_split_plan([90000, 36000], 8) == [13, 5]                     # C3 on one GPU: 18 splits x 8 KV heads = 144 CTAs
    assert ops.xattn_split_plan([11250, 4500], 8) == [13, 5]                      # C3 on 8 GPUs (per-rank share)
    for keys in ([0, 300], [1, 1], [511, 513], [5000], [126000], [100, 70000], [70000, 100], [0], [0, 0]):
        for hkv in (2, 8):
            plan = ops.xattn_split_plan(keys, hkv, 148)
            assert len(plan) == len(keys) and all(p >= 1 for p in plan)
            assert sum(plan) <= max(len(keys), 148 // hkv)
            assert all(p <= max(1, (k + 511) // 512) for p, k in zip(plan, keys))


def test_exchange_arena_layout_arithmetic():
    """exchange.py: arena = data fp32 [2 slots][world][cap] | flags uint32 [2 slots][world] | counter | err; one rank block holds up to
    two streams of (O [rows, dh] | LSE [rows]); the text pass's slot / flag addressing in csrc/textpass.cu uses the same numbers."""
    from vidi_b200.exchange import PartialExchange, arena_bytes
    rows, dh = 256 * 16, 256
    cap = PartialExchange.capacity(rows, dh)
    assert cap == 2 * rows * (dh + 1)
    for world in (2, 4, 8):
        n = arena_bytes(world, cap)
        assert n == 2 * world * cap * 4 + 2 * world * 4 + 8 and n % 4 == 0
        flags_off = 2 * world * cap * 4
        for seq in (1, 2, 3, 42, 43):
            slot = seq & 1
            for r in range(world):
                data = slot * world * cap * 4 + r * cap * 4          # block of rank r in the slot
                flag = flags_off + (slot * world + r) * 4
                assert 0 <= data and data + cap * 4 <= flags_off and flags_off <= flag < flags_off + 2 * world * 4
    assert 8 * arena_bytes(8, cap) < 2 ** 31                          # C3 text rows at 8 ranks: well under 2 GB in total
