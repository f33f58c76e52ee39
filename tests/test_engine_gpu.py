"""End-to-end parity: CUDA engine (through the C ABI) vs the fp32 CPU oracle on the same seeded inputs.

Tolerances (bf16 engine vs fp32 oracle; stated per SURVEY.md 8c):
  * stream activations / K,V per layer: relative L2 <= 2e-2
  * final logits (after the 30*tanh soft-cap): max-abs <= 0.25 and relative L2 <= 3e-2
  * greedy token ids identical wherever the oracle's top-1 margin exceeds the logit tolerance
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def build(cfg, seed=1234):
    from oracle import synth
    from vidi_b200.engine import Vidi15Engine
    sd = synth.make_state_dict(cfg, seed=seed)
    # the engine sees bf16 weights; the oracle sees the same bf16-rounded values in fp32
    sd_r = {k: (v if "mm_rand_pos" in k else v.to(BF).float()) for k, v in sd.items()}
    eng = Vidi15Engine(cfg, {k: v.clone() for k, v in sd_r.items()}, device="cuda")
    return sd_r, eng


def run_case(cfg, n_frames, n_chunks, n_text, audio_size=None, check_stages=True):
    from oracle import synth, vidi15_ref as R
    sd, eng = build(cfg)
    ids, images, mels, asz = synth.make_inputs(cfg, n_frames, n_chunks, n_text=n_text, audio_size=audio_size)
    images = images.to(BF).float(); mels = mels.to(BF).float()          # both sides see bf16-representable inputs
    ref_logits, inter = R.prefill(sd, cfg, ids, images, mels, asz, normalizer_dtype=BF, return_intermediates=True)
    ids_dev = R.strip_image_token(ids).cuda()
    logits, st = eng.prefill(ids_dev, images.cuda().to(BF), mels.cuda().to(BF), asz, return_state=True)
    torch.cuda.synchronize()
    n_img, n_aud = inter["image_embeds"].shape[0], inter["audio_embeds"].shape[0]
    assert st["streams"].shape[0] == n_img + n_aud
    if check_stages:
        nrm = R.normalizer(cfg, BF)
        kv0 = st["kv"][0]
        kd = cfg.llm.kv_dim
        K0 = torch.cat([inter["kv"][0][0][0], inter["kv"][0][1][0]], 0)
        V0 = torch.cat([inter["kv"][0][0][1], inter["kv"][0][1][1]], 0)
        assert rel(kv0[:, :kd], K0) < 2e-2, ("K layer0", rel(kv0[:, :kd], K0))
        assert rel(kv0[:, kd:], V0) < 2e-2, ("V layer0", rel(kv0[:, kd:], V0))
        Ll = cfg.llm.layers - 1
        KL = torch.cat([inter["kv"][Ll][0][0], inter["kv"][Ll][1][0]], 0)
        assert rel(st["kv"][Ll][:, :kd], KL) < 3e-2, ("K last", rel(st["kv"][Ll][:, :kd], KL))
    err = float((logits.cpu() - ref_logits).abs().max())
    assert rel(logits, ref_logits) < 3e-2, rel(logits, ref_logits)
    assert err < 0.25, err
    top2 = ref_logits.topk(2, -1).values
    confident = (top2[:, 0] - top2[:, 1]) > 2 * err
    assert torch.equal(logits.cpu().argmax(-1)[confident], ref_logits.argmax(-1)[confident])
    return logits, ref_logits


def test_prefill_mini_config1_shape():
    """BASELINE config 1 shape (8 frames, 1 audio chunk) at mini dims."""
    from vidi_b200.config import vidi15_mini
    run_case(vidi15_mini(), n_frames=8, n_chunks=1, n_text=24)


def test_ln_fold_matches_default():
    """Optional tower path with both LayerNorms folded into the consuming GEMMs (engine.enable_ln_fold): tower outputs agree with
    the default LayerNorm-kernel path and with the fp32 oracle."""
    from oracle import synth, vidi15_ref as R
    from vidi_b200.config import vidi15_mini
    cfg = vidi15_mini()
    sd, eng = build(cfg)
    ids, images, mels, asz = synth.make_inputs(cfg, 5, 2, n_text=8)
    img = images.to(BF).cuda().flatten(0, 1) if images.dim() == 5 else images.to(BF).cuda()
    mel = mels.to(BF).cuda().flatten(0, 1) if mels.dim() == 4 else mels.to(BF).cuda()
    v0, a0 = eng.siglip(img).float(), eng.whisper(mel).float()
    eng.enable_ln_fold(True)
    v1, a1 = eng.siglip(img).float(), eng.whisper(mel).float()
    eng.enable_ln_fold(False)
    assert rel(v1, v0) < 1.5e-2 and rel(a1, a0) < 1.5e-2, (rel(v1, v0), rel(a1, a0))
    ref_v = R.siglip_tower(sd, cfg, img.float().cpu())
    assert rel(v1, ref_v.reshape(v1.shape)) < 2.5e-2, rel(v1, ref_v.reshape(v1.shape))


def test_prefill_mini_resized_featuremap():
    """Force the > max_image_tokens branch (bilinear shrink, utils.py:152-171) with a small cap."""
    import dataclasses
    from vidi_b200.config import vidi15_mini
    cfg = dataclasses.replace(vidi15_mini(), max_image_tokens=400)      # 6 frames * 784 > 1600 -> 16x16 maps
    assert cfg.image_hw(6) != (28, 28)
    run_case(cfg, n_frames=6, n_chunks=1, n_text=9, check_stages=False)


def test_prefill_ragged_audio_and_odd_text():
    from vidi_b200.config import vidi15_mini
    run_case(vidi15_mini(llm_layers=3), n_frames=3, n_chunks=2, n_text=33, audio_size=4321, check_stages=False)


def test_prefill_true_dims_depth_cut():
    """True 9B hidden sizes (3584 / 1152 / 1280, head dims 256 / 72 / 64), depth cut to keep the CPU oracle fast."""
    from vidi_b200.config import vidi15_true_dims
    run_case(vidi15_true_dims(llm_layers=2, vis_layers=2, aud_layers=1, vocab=4096), n_frames=2, n_chunks=1, n_text=16,
             audio_size=400)


def test_sharded_equals_single_rank_fake_multirank():
    """Run the shard function for ranks 0..P-1 serially on one device and merge: must match the 1-rank result
    (SURVEY.md section 4c).  Exercises ShardPlan offsets, global positional indices and the partial merge."""
    from oracle import synth, vidi15_ref as R
    from vidi_b200 import ops
    from vidi_b200.config import vidi15_mini
    from vidi_b200.engine import make_plan
    cfg = vidi15_mini()
    sd, eng = build(cfg)
    ids, images, mels, asz = synth.make_inputs(cfg, 5, 3, n_text=11, audio_size=7000)
    images = images.cuda().to(BF); mels = mels.cuda().to(BF)
    ids_dev = R.strip_image_token(ids).cuda()
    full, st = eng.prefill(ids_dev, images, mels, asz, return_state=True)
    world = 3
    S_parts, kv_parts = [], []
    for r in range(world):
        eng.rank, eng.world = r, world
        plan = make_plan(cfg, 5, 3, asz, r, world)
        S, seg = eng.encode_streams(images[plan.f0:plan.f1], mels[plan.c0:plan.c1], plan)
        S_parts.append((S.clone(), seg, plan))
        kv_parts.append(eng.stream_pass(S))
    eng.rank, eng.world = 0, 1
    # stream rows of all ranks, re-ordered to [all image rows | all audio rows], must equal the 1-rank stream
    img_rows = torch.cat([kv[:, seg[0][0]:seg[0][0] + seg[0][1]] for kv, (_, seg, _) in zip(kv_parts, S_parts)], 1)
    aud_rows = torch.cat([kv[:, seg[1][0]:seg[1][0] + seg[1][1]] for kv, (_, seg, _) in zip(kv_parts, S_parts)], 1)
    merged = torch.cat([img_rows, aud_rows], 1)
    assert merged.shape == st["kv"].shape
    assert rel(merged, st["kv"]) < 1e-2
    seg_full = [(0, img_rows.shape[1], None, 1.0, img_rows.shape[1]), (img_rows.shape[1], aud_rows.shape[1], None, 1.0, aud_rows.shape[1])]
    logits2 = eng.text_pass(ids_dev, merged.contiguous(), seg_full)
    assert rel(logits2, full) < 1e-2


def _lockstep_world(eng, cfg, world, ids_dev, images, mels, asz, F, Cn, family_mistral=False):
    """Run the REAL multi-rank text path for `world` fake ranks on one device: each fake rank is a shallow copy of the engine (shared
    weights) with its own rank / world / peer-exchange handle; encode + stream pass per rank, then all ranks' _TextRun objects advance
    layer by layer in lock step (begin: partials + push; end: flag wait + rank merge).  Exercises engine.py's `world > 1` branches,
    splits sized from ceil(n_total / world), premerge + peer stores + rank-strided merge -- everything but the IPC mapping."""
    import copy
    from vidi_b200.engine import _TextRun, make_plan
    from vidi_b200.exchange import PartialExchange
    c = cfg.llm
    xs = PartialExchange.local_group(world, PartialExchange.capacity(ids_dev.numel() * c.heads, c.head_dim))
    runs, plans = [], []
    for r in range(world):
        e = copy.copy(eng)
        e.rank, e.world, e.xchg = r, world, xs[r]
        plan = make_plan(cfg, F, Cn, asz, r, world)
        S, seg = e.encode_streams(images[plan.f0:plan.f1], mels[plan.c0:plan.c1], plan)
        kv = e.stream_pass(S)
        run = _TextRun(e, ids_dev, kv, seg, None, 0)
        assert run.mode == "p2p"
        runs.append(run); plans.append(plan)
    for l in range(c.layers):
        for run in runs:
            run.layer_begin(l)
        for run in runs:
            run.layer_end(l)
    outs = [run.finish() for run in runs]
    torch.cuda.synchronize()
    for x in xs:
        x.check()
    return outs, plans


@pytest.mark.parametrize("world,F,Cn,asz", [(2, 5, 3, 7000), (3, 5, 2, 4100), (8, 9, 3, 8200)])
def test_multirank_text_path_lockstep_equals_single_rank(world, F, Cn, asz):
    """N-rank vs 1-rank logits through the engine's own multi-rank code (replacement of Gather.forward, all_to_all.py:361):
    uneven frame split, ranks with zero audio chunks (world 3: 2 chunks; world 8: 3 chunks), a partial last audio chunk."""
    from oracle import synth, vidi15_ref as R
    from vidi_b200.config import vidi15_mini
    cfg = vidi15_mini()
    sd, eng = build(cfg)
    ids, images, mels, asz = synth.make_inputs(cfg, F, Cn, n_text=11, audio_size=asz)
    images = images.cuda().to(BF); mels = mels.cuda().to(BF)
    ids_dev = R.strip_image_token(ids).cuda()
    full = eng.prefill(ids_dev, images, mels, asz)
    outs, plans = _lockstep_world(eng, cfg, world, ids_dev, images, mels, asz, F, Cn)
    assert any(p.n_aud == 0 for p in plans) or world == 2
    assert sum(p.n_img for p in plans) == plans[0].n_img_total and sum(p.n_aud for p in plans) == plans[0].n_aud_total
    for o in outs[1:]:
        assert torch.equal(o, outs[0])               # every rank merges the same partials in the same order
    assert rel(outs[0], full) < 1e-2, rel(outs[0], full)
    ref = R.prefill(sd, cfg, ids, images.float().cpu(), mels.float().cpu(), asz, normalizer_dtype=BF)
    err = float((outs[0].cpu() - ref).abs().max())
    assert rel(outs[0], ref) < 3e-2 and err < 0.25
    top2 = ref.topk(2, -1).values
    confident = (top2[:, 0] - top2[:, 1]) > 2 * err
    assert torch.equal(outs[0].cpu().argmax(-1)[confident], ref.argmax(-1)[confident])


@pytest.mark.parametrize("family", ["gemma", "mistral"])
def test_native_text_pass_equals_per_layer_path(family):
    """vidi_text_pass (csrc/textpass.cu, one C call for all layers) issues the same kernels with the same arguments as the per-layer
    Python path: prefill logits, the text K||V cache it fills and three decode steps must be BIT-identical, for both model families,
    with a key-padding mask in play (zero frame) and logits_to_keep."""
    from oracle import synth
    from vidi_b200.config import vidi15_mini, vidi7b_mini
    from vidi_b200.engine import Vidi15Engine
    cfg = vidi15_mini(llm_layers=3) if family == "gemma" else vidi7b_mini(llm_layers=3)
    sd = synth.make_state_dict(cfg, seed=77)
    eng = Vidi15Engine(cfg, {k: (v if "mm_rand_pos" in k else v.to(BF)) for k, v in sd.items()}, device="cuda")
    ids, images, mels, asz = synth.make_inputs(cfg, 4, 2, n_text=21, audio_size=4000)
    images[1] = 0                                        # an all-zero frame -> masked keys (multimodal.py:202)
    ids_dev = ids[ids != -200].cuda()
    img, mel = images.cuda().to(BF), mels.cuda().to(BF)
    outs = {}
    for native in (False, True):
        eng.native_text = native
        tc = eng.new_text_cache(32)
        lg, st = eng.prefill(ids_dev, img, mel, asz, text_cache=tc, logits_to_keep=5, return_state=True)
        steps = [lg]
        for tok in (17, 400, 3):
            steps.append(eng.text_pass(torch.tensor([tok], device="cuda"), st["kv"], st["seg"], text_cache=tc, logits_to_keep=1))
        no_cache = eng.text_pass(ids_dev, st["kv"], st["seg"])
        outs[native] = (steps, tc["kv"][:, :tc["len"]].clone(), tc["len"], no_cache)
    assert outs[True][2] == outs[False][2] == 21 + 3
    assert torch.equal(outs[True][1], outs[False][1])
    for a, b in zip(outs[True][0], outs[False][0]):
        assert a.shape == b.shape and torch.equal(a, b)
    assert outs[True][0][0].shape == (5, cfg.llm.vocab) and torch.equal(outs[True][3], outs[False][3])


def test_prefill_c3_shape_true_dims_depth_cut():
    """BASELINE config 3's SHAPE features at true 9B hidden dims with the depth cut (SURVEY.md 8c allows layer-subsampled checks): the
    10 x 10 floor of the resize branch (multimodal.py:175-180, utils.py:152-171) -> 25 tokens per frame, several audio chunks with the
    last one partial (the floor(size/2) / floor(./5) trimming of multimodal.py:224-236), 32 text tokens; engine vs fp32 oracle."""
    import dataclasses
    from vidi_b200.config import vidi15_true_dims
    cfg = dataclasses.replace(vidi15_true_dims(llm_layers=2, vis_layers=2, aud_layers=1, vocab=4096), max_image_tokens=300)
    F, Cn = 14, 3
    assert cfg.image_hw(F) == (10, 10) and cfg.image_tokens(F) == F * 25
    run_case(cfg, n_frames=F, n_chunks=Cn, n_text=32, audio_size=Cn * 3000 - 1234, check_stages=True)


@pytest.mark.parametrize("name,l,N_", [("t", 3600, 10000), ("t", 36000, 10000), ("h", 5, 2)])
def test_pos_table_at_c3_lengths(name, l, N_):
    """LearnablePosEmbd at BASELINE config 3's lengths (pos.py:41-58): l = 3600 frames and l = 36 000 audio tokens against
    mm_time_interval = 10 000; the engine evaluates the fp32 MLP as split-bf16 GEMMs (3 terms) and rms-normalises -- vs the oracle's
    fp32 evaluation on the rows of a shard in the middle and at both ends."""
    from oracle import synth, vidi15_ref as R
    from vidi_b200.config import vidi15_true_dims
    cfg = vidi15_true_dims(llm_layers=1, vis_layers=2, aud_layers=1, vocab=256)
    sd, eng = build(cfg)
    D = cfg.llm.hidden
    ref = R.xhat(R.pos_embed(sd, f"model.mm_rand_pos_{name}", l, N_, D), cfg.mm_eps)
    for i0, rows in {(0, min(l, 40)), (max(0, l // 2 - 7), min(l - max(0, l // 2 - 7), 33)), (max(0, l - 29), min(l, 29))}:
        tab = eng.pos_table(name, rows, i0, l, N_)
        got, want = tab[:rows].float().cpu(), ref[i0:i0 + rows]
        # the table is rounded to bf16 after the rms-norm (it is added to bf16 activations): compare at bf16 resolution
        assert float((got - want).abs().max()) <= 2 ** -7 * float(want.abs().max()) + 1e-3, (name, l, i0, float((got - want).abs().max()))
        assert rel(got, want) < 4e-3


def test_vidi7b_prefill_mini():
    """Vidi-7B (Mistral Dattn): learned-conv pooling, SwiGLU, no post-norms / soft-caps, dh=128."""
    from oracle import synth, vidi7b_ref as R7
    from oracle.vidi15_ref import strip_image_token
    from vidi_b200.config import vidi7b_mini
    from vidi_b200.engine import Vidi15Engine
    cfg = vidi7b_mini()
    sd = synth.make_state_dict(cfg, seed=4242)
    sd = {k: (v if "mm_rand_pos" in k else v.to(BF).float()) for k, v in sd.items()}
    eng = Vidi15Engine(cfg, {k: v.clone() for k, v in sd.items()}, device="cuda")
    assert not eng.gemma and eng.normalizer == 1.0
    ids, images, mels, asz = synth.make_inputs(cfg, 5, 2, n_text=19, audio_size=4100)
    images = images.to(BF).float(); mels = mels.to(BF).float()
    ref, inter = R7.prefill(sd, cfg, ids, images, mels, asz, return_intermediates=True)
    logits, st = eng.prefill(strip_image_token(ids).cuda(), images.cuda().to(BF), mels.cuda().to(BF), asz, return_state=True)
    torch.cuda.synchronize()
    n_img = inter["image_embeds"].shape[0]
    assert n_img == 5 * cfg.mm_image_pool_size ** 2 and st["streams"].shape[0] == n_img + inter["audio_embeds"].shape[0]
    kd = cfg.llm.kv_dim
    K0 = torch.cat([inter["kv"][0][0][0], inter["kv"][0][1][0]], 0)
    assert rel(st["kv"][0][:, :kd], K0) < 2e-2
    err = float((logits.cpu() - ref).abs().max())
    assert rel(logits, ref) < 3e-2 and err < 0.25, (rel(logits, ref), err)
    top2 = ref.topk(2, -1).values
    confident = (top2[:, 0] - top2[:, 1]) > 2 * err
    assert torch.equal(logits.cpu().argmax(-1)[confident], ref.argmax(-1)[confident])


def test_facade_forward_generate_match_oracle():
    """Drop-in surface (model.py): forward(..., images=, audios=, audio_sizes=) logits and greedy generate() ids vs the oracle,
    including the decode steps that run q_len=1 against the cached image/audio/text K,V (gemma.py:64-65,603-655)."""
    from oracle import synth, vidi15_ref as R
    from vidi_b200.config import vidi15_mini
    from vidi_b200.model import DattnGemma2ForCausalLM, IMAGE_TOKEN_INDEX
    cfg = vidi15_mini()
    sd = {k: (v if "mm_rand_pos" in k else v.to(BF).float()) for k, v in synth.make_state_dict(cfg, seed=99).items()}
    model = DattnGemma2ForCausalLM(cfg, {k: v.clone() for k, v in sd.items()}, device="cuda")
    ids, images, mels, asz = synth.make_inputs(cfg, 4, 1, n_text=14, seed=5)
    images = images.to(BF).float(); mels = mels.to(BF).float()
    assert int((ids == IMAGE_TOKEN_INDEX).sum()) == 1
    out = model(ids[None], images=images[None], audios=mels[None], audio_sizes=[asz])       # host fp32 inputs, no mask
    ref = R.prefill(sd, cfg, ids, images, mels, asz, normalizer_dtype=BF)
    assert out.logits.shape == (1, 14, cfg.llm.vocab)
    assert rel(out.logits[0], ref) < 3e-2
    k_img, v_img = out.past_image_key_values[0]
    assert k_img.shape == (1, cfg.image_tokens(4), cfg.llm.kv_dim) and len(out.past_image_key_values) == cfg.llm.layers
    assert out.past_audio_key_values[0][0].shape[1] == cfg.audio_tokens(asz)
    # greedy decode: compare step by step while the oracle's margin is decisive
    n_new = 6
    gen = model.generate(ids[None], images=images[None], audios=mels[None], audio_sizes=[asz], do_sample=False,
                         max_new_tokens=n_new, use_cache=True, disable_compile=True, pad_token_id=0)
    ref_ids, margins = R.greedy_generate(sd, cfg, ids, images, mels, asz, max_new_tokens=n_new, normalizer_dtype=BF, return_margins=True)
    assert min(margins) > 1.0                          # decisive fixture (input seed 5): full equality is a fair demand
    assert gen[0].tolist() == ref_ids                  # see tests/test_generate_gpu.py for a non-degenerate decode
    with pytest.raises(NotImplementedError):
        model.generate(ids[None], inputs_embeds=torch.zeros(1))
    with pytest.raises(ValueError):
        model.forward(None)


def test_prefill_long_text_multiple_query_blocks():
    """T = 150 text tokens -> 300 virtual rows = 3 query blocks in the cross-attention kernel, text attention over 150 keys."""
    from vidi_b200.config import vidi15_mini
    run_case(vidi15_mini(), n_frames=2, n_chunks=1, n_text=150, audio_size=300, check_stages=False)


def test_prefill_single_modality():
    """images only / audios only (encode_videos handles either being None, multimodal.py:254-265)."""
    from oracle import synth, vidi15_ref as R
    from vidi_b200.config import vidi15_mini
    cfg = vidi15_mini()
    sd, eng = build(cfg)
    ids, images, mels, asz = synth.make_inputs(cfg, 3, 1, n_text=10, audio_size=700)
    images = images.to(BF).float(); mels = mels.to(BF).float()
    ids_dev = R.strip_image_token(ids).cuda()
    ref_i = R.prefill(sd, cfg, ids, images, None, None, normalizer_dtype=BF)
    out_i = eng.prefill(ids_dev, images.cuda().to(BF), None, 0)
    assert rel(out_i, ref_i) < 3e-2
    ref_a = R.prefill(sd, cfg, ids, None, mels, asz, normalizer_dtype=BF)
    out_a = eng.prefill(ids_dev, None, mels.cuda().to(BF), asz)
    assert rel(out_a, ref_a) < 3e-2


def test_facade_batch_with_padding_mask():
    """B=2 with right padding + attention_mask: each sample is stripped of padding and the sentinel (multimodal.py:363-397)."""
    from oracle import synth, vidi15_ref as R
    from vidi_b200.config import vidi15_mini
    from vidi_b200.model import DattnGemma2ForCausalLM
    cfg = vidi15_mini()
    sd = {k: (v if "mm_rand_pos" in k else v.to(BF).float()) for k, v in synth.make_state_dict(cfg, seed=11).items()}
    model = DattnGemma2ForCausalLM(cfg, {k: v.clone() for k, v in sd.items()}, device="cuda")
    ids0, img0, mel0, a0 = synth.make_inputs(cfg, 2, 1, n_text=9, seed=1, audio_size=500)
    ids1, img1, mel1, a1 = synth.make_inputs(cfg, 2, 1, n_text=5, seed=2, audio_size=900)
    img0, img1, mel0, mel1 = [t.to(BF).float() for t in (img0, img1, mel0, mel1)]
    L = max(ids0.numel(), ids1.numel())
    ids = torch.zeros(2, L, dtype=torch.long); mask = torch.zeros(2, L, dtype=torch.long)
    ids[0, :ids0.numel()] = ids0; mask[0, :ids0.numel()] = 1
    ids[1, :ids1.numel()] = ids1; mask[1, :ids1.numel()] = 1
    out = model(ids, attention_mask=mask, images=torch.stack([img0, img1]), audios=torch.stack([mel0, mel1]), audio_sizes=[a0, a1])
    r0 = R.prefill(sd, cfg, ids0, img0, mel0, a0, normalizer_dtype=BF)
    r1 = R.prefill(sd, cfg, ids1, img1, mel1, a1, normalizer_dtype=BF)
    assert out.logits.shape == (2, 9, cfg.llm.vocab)
    assert rel(out.logits[0], r0) < 3e-2 and rel(out.logits[1, :5], r1) < 3e-2
    assert float(out.logits[1, 5:].abs().max()) == 0.0


class _LazyStateDict:
    """HF-layout state_dict view that materialises one tensor at a time (generated on the GPU, handed to the CPU oracle as
    fp32): lets the fp32 oracle walk the FULL 9B model without a 37 GB host copy.  Values are identical to
    synth.make_state_dict(cfg, seed, device='cuda', dtype=bf16) because every tensor has its own seeded generator."""

    def __init__(self, cfg, seed):
        from vidi_b200 import synth
        self.cfg, self.seed, self.synth = cfg, seed, synth
        self.specs = synth.tensor_specs(cfg)

    def __contains__(self, k):
        return k in self.specs

    def get(self, k, default=None):
        return self[k] if k in self.specs else default

    def __getitem__(self, k):
        import time
        t0 = time.perf_counter()
        t = self.synth.make_tensor(self.cfg, k, self.specs[k], self.seed, "cuda", BF)
        out = t.float().cpu()
        self.seconds = getattr(self, "seconds", 0.0) + time.perf_counter() - t0      # weight materialisation, not oracle compute
        return out


def test_config1_full_depth_true_9b_vs_oracle():
    """BASELINE config 1 (8 frames, 8 s audio, 1 680 tokens) at the TRUE Vidi1.5-9B dims and FULL depth (42 + 26 + 32 layers):
    engine logits vs the fp32 CPU oracle on identical weights.  bf16 drift through the full stack is the stated tolerance."""
    from oracle import synth, vidi15_ref as R
    from vidi_b200.config import vidi15_9b
    from vidi_b200.engine import Vidi15Engine
    cfg = vidi15_9b()
    seed = 1234
    sd_gpu = synth.make_state_dict(cfg, seed=seed, device="cuda", dtype=BF)
    eng = Vidi15Engine(cfg, sd_gpu, device="cuda", pop_state_dict=True)
    del sd_gpu
    ids, images, mels, asz = synth.make_inputs(cfg, 8, 1, n_text=32, audio_size=800)
    images = images.to(BF).float(); mels = mels.to(BF).float()
    logits = eng.prefill(R.strip_image_token(ids).cuda(), images.cuda().to(BF), mels.cuda().to(BF), asz)
    torch.cuda.synchronize()
    import json, os, time
    torch.set_num_threads(min(64, torch.get_num_threads()))
    lazy = _LazyStateDict(cfg, seed)
    t0 = time.perf_counter()
    ref = R.prefill(lazy, cfg, ids, images, mels, asz, normalizer_dtype=BF)
    wall = time.perf_counter() - t0
    # config 1 timed IN FULL on the host cores (no extrapolation): oracle compute = wall - weight materialisation
    if os.environ.get("VIDI_EVIDENCE_DIR"):
        n_tok = cfg.image_tokens(8) + cfg.audio_tokens(asz) + 32
        cpu_s = wall - getattr(lazy, "seconds", 0.0)
        with open(os.path.join(os.environ["VIDI_EVIDENCE_DIR"], "r02_cpu_c1_full.json"), "w") as f:
            json.dump(dict(workload="BASELINE config 1: 8 frames, 8 s audio, 32 text tokens, true Vidi1.5-9B dims, full depth", tokens=n_tok,
                           oracle_seconds=round(cpu_s, 2), weight_materialisation_seconds=round(getattr(lazy, "seconds", 0.0), 2),
                           tokens_per_s=round(n_tok / cpu_s, 2), threads=torch.get_num_threads(), host_cores=os.cpu_count(),
                           kind="port (fp32 oracle, full run, not extrapolated)"), f)
    r, err = rel(logits, ref), float((logits.cpu() - ref).abs().max())
    top2 = ref.topk(2, -1).values
    confident = (top2[:, 0] - top2[:, 1]) > 2 * err
    agree = torch.equal(logits.cpu().argmax(-1)[confident], ref.argmax(-1)[confident])
    print(f"[config1 full depth] rel-L2 {r:.4f} max-abs {err:.4f} confident positions {int(confident.sum())}/32 top-1 agree {agree}")
    assert logits.shape == (32, cfg.llm.vocab)
    assert r < 3e-2 and err < 0.25, (r, err)
    assert agree


@pytest.mark.parametrize("N", [126000])
def test_full_size_properties_cross_attention(N):
    """Size-independent properties at the BASELINE 128k size: (i) the merged result does not depend on the split count,
    (ii) cross attention is permutation-invariant in the keys (no RoPE, non-causal), (iii) identical calls are bit-identical."""
    from vidi_b200 import ops
    T, Hq, Hkv, dh = 32, 16, 8, 256
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    q = torch.randn(T, Hq * dh, generator=g, device="cuda").to(BF)
    kv = torch.randn(N, 2 * Hkv * dh, generator=g, device="cuda").to(BF)

    def run(kvt, splits):
        op, lse = ops.xattn_splitkv(q, kvt[:, :Hkv * dh], kvt[:, Hkv * dh:], None, Hq, Hkv, dh, 1 / 16, 50.0, splits)
        out = torch.zeros(T * Hq, dh, device="cuda")
        return ops.xattn_merge(op, lse, out)
    a, b = run(kv, 18), run(kv, 37)
    assert rel(a, b) < 2e-3
    perm = torch.randperm(N, generator=g, device="cuda")
    c = run(kv[perm].contiguous(), 18)
    assert rel(c, a) < 2e-3
    assert torch.equal(run(kv, 18), a)


def test_load_pretrained_model_from_safetensors_dir(tmp_path):
    """builder.py:24-64 surface: a checkpoint directory with config.json + HF-layout *.safetensors shards loads into a model
    whose logits equal the directly-constructed one."""
    import dataclasses
    import json
    from safetensors.torch import save_file
    from oracle import synth
    from vidi_b200.config import vidi15_mini
    from vidi_b200.model import DattnGemma2ForCausalLM, load_pretrained_model
    cfg = vidi15_mini()
    sd = {k: (v if "mm_rand_pos" in k else v.to(BF)) for k, v in synth.make_state_dict(cfg, seed=21).items()}
    keys = sorted(sd)
    save_file({k: sd[k].contiguous() for k in keys[: len(keys) // 2]}, str(tmp_path / "model-00001-of-00002.safetensors"))
    save_file({k: sd[k].contiguous() for k in keys[len(keys) // 2:]}, str(tmp_path / "model-00002-of-00002.safetensors"))
    c = cfg.llm
    hf = dict(model_type="dattn_gemma2", hidden_size=c.hidden, num_attention_heads=c.heads, num_key_value_heads=c.kv_heads,
              head_dim=c.head_dim, intermediate_size=c.inter, num_hidden_layers=c.layers, vocab_size=c.vocab,
              rms_norm_eps=c.rms_eps, query_pre_attn_scalar=c.query_pre_attn_scalar, attn_logit_softcapping=c.attn_softcap,
              final_logit_softcapping=c.final_softcap, sliding_window=c.sliding_window, mm_image_pool_size=2,
              mm_audio_pool_size=5, mm_time_interval=10000, mm_std=cfg.mm_std, mm_input_type="video",
              vision_config=dataclasses.asdict(cfg.vis), audio_config=dataclasses.asdict(cfg.aud))
    (tmp_path / "config.json").write_text(json.dumps(hf))
    model, tok, img_proc, aud_proc = load_pretrained_model(str(tmp_path))
    assert tok is None and img_proc.size["height"] == cfg.vis.image and aud_proc.sampling_rate == 16000
    model.config.mm_splits = 32                                            # inference.py:86 keeps working
    with pytest.raises(NotImplementedError):
        load_pretrained_model(str(tmp_path), load_8bit=True)
    direct = DattnGemma2ForCausalLM(cfg, {k: v.clone() for k, v in sd.items()}, device="cuda")
    ids, images, mels, asz = synth.make_inputs(cfg, 2, 1, n_text=8, seed=3, audio_size=400)
    a = model(ids[None], images=images[None], audios=mels[None], audio_sizes=[asz]).logits
    b = direct(ids[None], images=images[None], audios=mels[None], audio_sizes=[asz]).logits
    assert torch.equal(a, b)
