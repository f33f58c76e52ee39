"""Vidi1.5-9B prefill engine: host-side orchestration of the sm_100a kernels.

Design: the image and audio streams are token-wise independent through all decoder
layers (gemma.py:183-202 reads only the stream itself), so
  1. the towers + projectors run per frame / per audio chunk               (encode_images / encode_audios)
  2. ONE concatenated [N_img+N_aud, D] stream runs the whole layer stack, writing the per-layer
     K||V cache [L, N, 2*kv_dim] and applying the diagonal V2V + GeGLU updates   (stream_pass)
  3. the short text stream then runs its layers against that cache with the split-KV
     cross-attention kernel; across GPUs only the (O, LSE) partials are exchanged  (text_pass)
Frames / chunks / tokens are sharded contiguously over ranks (ShardPlan); the text stream is replicated.
Nothing here computes on the CPU; torch supplies memory and streams; the multi-GPU exchange of the partials is peer-memory stores
(exchange.py / csrc/xchg.cu), with one NCCL all-gather per layer only as the fall-back when the arenas cannot be mapped.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch

from . import ops
from .weights import load_vidi15

BF16 = torch.bfloat16
BIG = 1 << 30


@dataclass
class ShardPlan:
    """Contiguous partition of frames / audio chunks / stream tokens over ranks."""
    rank: int
    world: int
    F: int
    C: int
    f0: int
    f1: int
    c0: int
    c1: int
    hw: tuple
    tpf: int            # image tokens per frame
    n_img_total: int
    n_aud_total: int    # s2
    s1: int
    tpc: int            # audio tokens per chunk
    a0: int             # first global audio token of this rank
    a1: int

    @property
    def n_img(self) -> int:
        return (self.f1 - self.f0) * self.tpf

    @property
    def n_aud(self) -> int:
        return max(0, self.a1 - self.a0)


def _split(n: int, world: int, rank: int):
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def make_plan(cfg, n_frames: int, n_chunks: int, audio_size: int, rank: int = 0, world: int = 1) -> ShardPlan:
    if hasattr(cfg, "image_hw"):                 # Vidi1.5: pad/resize + space-to-depth
        hw = cfg.image_hw(n_frames) if n_frames else (28, 28)
        m = cfg.mm_image_pool_size
        tpf = (hw[0] // m) * (hw[1] // m)
    else:                                        # Vidi-7B: learned conv pool to pool x pool tokens per frame
        hw = (cfg.mm_image_pool_size, cfg.mm_image_pool_size)
        tpf = cfg.mm_image_pool_size ** 2
    f0, f1 = _split(n_frames, world, rank)
    c0, c1 = _split(n_chunks, world, rank)
    ratio = cfg.aud.max_source_positions / cfg.aud.nb_max_frames
    s1 = int(math.floor(audio_size * ratio)) if n_chunks else 0
    s1 = min(s1, n_chunks * cfg.aud.max_source_positions)
    s2 = int(math.floor(s1 / cfg.mm_audio_pool_size))
    assert cfg.aud.max_source_positions % cfg.mm_audio_pool_size == 0
    tpc = cfg.aud.max_source_positions // cfg.mm_audio_pool_size
    a0, a1 = min(c0 * tpc, s2), min(c1 * tpc, s2)
    return ShardPlan(rank, world, n_frames, n_chunks, f0, f1, c0, c1, hw, tpf, n_frames * tpf, s2, s1, tpc, a0, a1)


class Vidi15Engine:
    def __init__(self, cfg, state_dict: dict, device="cuda", rank: int = 0, world: int = 1, group=None,
                 pop_state_dict: bool = False, vit_chunk: int = 128, aud_chunk: int = 16, exchange: str = "auto",
                 max_text_rows: int = 256):
        if not torch.cuda.is_available():
            raise RuntimeError("vidi_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.cfg = cfg
        self.device = torch.device(device)
        self.rank, self.world, self.group = rank, world, group
        # Multi-rank text pass: (O, LSE) partials cross ranks through peer-mapped arenas ("p2p", csrc/xchg.cu) -- or, when the arenas
        # cannot be mapped (or exchange="nccl"), through one NCCL all-gather per layer.  Both are GPU paths with identical results.
        self.xchg, self.exchange_note = None, "single rank"
        if world > 1:
            self.exchange_note = "nccl all-gather of the per-rank reduced (O, LSE) blocks"
            if exchange in ("auto", "p2p"):
                from .exchange import PartialExchange
                try:
                    self.xchg = PartialExchange.connect(rank, world, group, PartialExchange.capacity(max_text_rows * cfg.llm.heads, cfg.llm.head_dim),
                                                        self.device)
                    self.exchange_note = "peer-memory stores over NVLink + flag wait (no collective call)"
                except RuntimeError as ex:
                    if exchange == "p2p":
                        raise
                    self.exchange_note += f" (peer arenas unavailable: {ex})"
        self.W = load_vidi15(state_dict, cfg, self.device, ops, pop=pop_state_dict)
        # Gemma2 family (Vidi1.5) vs Mistral family (Vidi-7B): SURVEY.md 3.3
        self.gemma = hasattr(cfg.llm, "final_softcap")
        # torch.tensor(hidden**0.5, dtype=act): the normaliser is rounded to the activation dtype (gemma.py:353); Mistral has none
        self.normalizer = float(torch.tensor(cfg.llm.hidden ** 0.5, dtype=BF16).float()) if self.gemma else 1.0
        self.glu = ops.GLU_GELU_TANH if self.gemma else ops.GLU_SILU
        # GEMM variant per site, from same-box A/B runs of the full step (profiles/).  Round 1 and early round 2: the CTA-pair kernel won on
        # the tower / projector shapes only and lost 1-8 % on the stream-pass GEMMs at M ~ 1e5; with the TMA-store epilogue in both kernels
        # the pair kernel wins there too (4 159 / 4 176 vs 4 197 ms/step, gate||up 1 274 -> 1 322 TF/s: profiles/r02_ab_llm_cta2_tmastore.txt).
        self.llm_cta2 = True
        # Optional: text pass on a side stream, one layer behind the stream pass (see prefill).  Measured NEGATIVE on B200
        # (profiles/r01_ab_text_overlap_{1,2}gpu.txt: +0.5 % / +2 % step time): the ~700 small kernels that slip in between the
        # persistent GEMMs delay those kernels' CTAs more than the hidden text latency is worth.  Kept off by default.
        self.overlap_text = False
        self.fold_ln = False         # see enable_ln_fold()
        # diagnostic: added to the key-split count of the cross attention.  Results are mathematically identical for any split
        # count; bench.py / the tests use it to measure the fp32 re-association noise floor that N-rank vs 1-rank comparisons
        # of a 42-layer random-weight stack sit on.
        self.split_bias = 0
        # text stream through the native executor (csrc/textpass.cu: one C call per pass instead of ~500 ctypes launches); the
        # per-layer Python path (_TextRun) stays for the side-stream overlap mode, the NCCL exchange and the lock-step tests
        self.native_text = True
        self.vit_chunk, self.aud_chunk = vit_chunk, aud_chunk
        self.n_sms = torch.cuda.get_device_properties(self.device).multi_processor_count

    # ------------------------------------------------------------------------------------------
    # towers
    # ------------------------------------------------------------------------------------------
    def enable_ln_fold(self, on: bool = True):
        """Optional tower path without LayerNorm kernels: both LayerNorms of a pre-LN block are folded into the GEMM that consumes
        them (weights.fold_layernorm; row statistics travel from the producing GEMM's epilogue to the consuming one's).
        Measured NEGATIVE on B200 for these K=1152/1280 GEMMs (profiles/r01_ab_ln_fold.txt): their epilogue is already the pacing
        stage, the extra epilogue work costs more GEMM time (+335 ms/step) than the LayerNorm passes it removes (-171 ms).
        Kept off by default; parity-tested (tests/test_engine_gpu.py::test_ln_fold_matches_default)."""
        if on and not hasattr(self.W.vis.layers[0], "wqkv_f"):
            from .weights import fold_tower_layer
            for L in list(self.W.vis.layers) + list(self.W.aud.layers):
                fold_tower_layer(L)
        self.fold_ln = on

    def _tower_layer(self, x, st, st2, L, B, S, heads, dh, eps, act):
        """one pre-LN encoder block, in place on x (HF SiglipEncoderLayer / WhisperEncoderLayer)."""
        if self.fold_ln:     # st / st2: per-row (sum, sumsq) of x written by the GEMM that produced it
            qkv = ops.gemm_ln(x, L.wqkv_f, bias=L.bqkv_f, ln=(st, L.cqkv, eps), tag="tower")
            a = ops.attn_dense(qkv, B, S, heads, dh, dh ** -0.5)
            ops.gemm_ln(a, L.wo, bias=L.bo, residual=x, out=x, stats=st2, tag="tower")
            m = ops.gemm_ln(x, L.w1_f, bias=L.b1_f, act=act, ln=(st2, L.c1, eps), tag="tower")
            ops.gemm_ln(m, L.w2, bias=L.b2, residual=x, out=x, stats=st, tag="tower")
            return x
        h = ops.layernorm(x, L.ln1_w, L.ln1_b, eps)
        qkv = ops.gemm(h, L.wqkv, bias=L.bqkv, tag="tower")
        a = ops.attn_dense(qkv, B, S, heads, dh, dh ** -0.5)
        ops.gemm(a, L.wo, bias=L.bo, residual=x, out=x, tag="tower")
        h = ops.layernorm(x, L.ln2_w, L.ln2_b, eps, out=h)
        m = ops.gemm(h, L.w1, bias=L.b1, act=act, tag="tower")
        ops.gemm(m, L.w2, bias=L.b2, residual=x, out=x, tag="tower")
        return x

    def _ln_stats(self, rows: int, width: int):
        if not self.fold_ln:
            return None, None
        st = torch.empty(rows, ops.ln_stats_parts(width), 2, device=self.device, dtype=torch.float32)
        return st, torch.empty_like(st)

    def siglip(self, images: torch.Tensor) -> torch.Tensor:
        """images [f,3,S,S] bf16 -> hidden_states[-2] [f*P, dv]  (siglip.py:29-34)."""
        v, Wv = self.cfg.vis, self.W.vis
        f = images.shape[0]
        A = ops.patch_im2col(images, v.patch, Wv.kpad)
        st, st2 = self._ln_stats(A.shape[0], v.hidden)
        if self.fold_ln:
            x = ops.gemm_ln(A, Wv.patch_w, bias=Wv.patch_b, residual=Wv.pos, res_mod=v.patches, stats=st, tag="vit",
                            alg_k=3 * v.patch * v.patch)
        else:
            x = ops.gemm(A, Wv.patch_w, bias=Wv.patch_b, residual=Wv.pos, res_mod=v.patches, tag="vit", alg_k=3 * v.patch * v.patch)
        del A
        for L in Wv.layers:
            x = self._tower_layer(x, st, st2, L, f, v.patches, v.heads, v.head_dim, v.eps, ops.ACT_GELU_TANH)
        return x

    def whisper(self, mels: torch.Tensor) -> torch.Tensor:
        """mels [c,128,3000] bf16 -> [c*1500, da]  (whisper.py:26-27)."""
        a, Wa = self.cfg.aud, self.W.aud
        c, T = mels.shape[0], mels.shape[2]
        A1 = ops.whisper_im2col1(mels)
        x1 = ops.gemm(A1, Wa.conv1_w, bias=Wa.conv1_b, act=ops.ACT_GELU_ERF)
        del A1
        A2 = ops.whisper_im2col2(x1, c, T)
        del x1
        st, st2 = self._ln_stats(A2.shape[0], a.d_model)
        if self.fold_ln:
            x = ops.gemm_ln(A2, Wa.conv2_w, bias=Wa.conv2_b, act=ops.ACT_GELU_ERF, residual=Wa.pos, res_mod=T // 2, stats=st)
        else:
            x = ops.gemm(A2, Wa.conv2_w, bias=Wa.conv2_b, act=ops.ACT_GELU_ERF, residual=Wa.pos, res_mod=T // 2)
        del A2
        for L in Wa.layers:
            x = self._tower_layer(x, st, st2, L, c, T // 2, a.heads, a.head_dim, a.eps, ops.ACT_GELU_ERF)
        return ops.layernorm(x, Wa.ln_w, Wa.ln_b, a.eps)

    def pos_table(self, name: str, rows: int, i0: int, l: int, N: int) -> torch.Tensor:
        """rms_norm(LearnablePosEmbd(...)) rows i0..i0+rows of l -> fp32 [rows, D]  (pos.py:41-58, norm.py:9-16)."""
        assert l > 1, "LearnablePosEmbd asserts x.shape[dim] > 1 (pos.py:42)"
        D, P = self.cfg.llm.hidden, self.W.pos[name]
        if rows == 0:
            return torch.zeros(1, D, device=self.device)
        A = ops.sinusoid_split(self.W.div_term, rows, i0, l, N, D)
        h = ops.gemm(A, P.w0, bias=P.b0, act=ops.ACT_GELU_ERF, out_fp32=True, tag="pos", alg_k=D)
        y = ops.gemm(ops.split3(h, 0), P.w2, bias=P.b2, out_fp32=True, tag="pos", alg_k=D)
        return ops.rmsnorm_f32(y, self.cfg.mm_eps, round_bf16=True)

    def _project(self, x, proj):
        h = ops.gemm(x, proj.w1, bias=proj.b1, act=ops.ACT_GELU_ERF)
        return ops.gemm(h, proj.w2, bias=proj.b2)

    def encode_images(self, images: torch.Tensor, plan: ShardPlan, sample_valid: bool = True):
        """This rank's frames [f1-f0,3,S,S] bf16 -> (stream rows [n_img, D] already * sqrt(D), mask uint8)
        (multimodal.py:156-208, gemma.py:353-355)."""
        cfg, v = self.cfg, self.cfg.vis
        D, m = cfg.llm.hidden, cfg.mm_image_pool_size
        fl = plan.f1 - plan.f0
        assert images.shape[0] == fl and images.dtype == BF16
        h, w = plan.hw
        proj = torch.empty(plan.n_img, D, device=self.device, dtype=BF16)
        for s in range(0, fl, self.vit_chunk):
            e = min(fl, s + self.vit_chunk)
            P = self.siglip(images[s:e])
            if self.gemma:
                X = ops.pool_s2d(P, e - s, v.side, h, w, m)
            else:                                    # Vidi_7B/model/mm_vision/pool.py:20-26
                k = self.W.img_pool_k
                A = ops.conv_window_gather(P, e - s, v.side, k)
                Y = ops.gemm(A, self.W.img_pool_w, tag="tower")
                X = ops.bilinear_ac(Y, e - s, v.side - k + 1, m)
                del A, Y
            del P
            hid = ops.gemm(X, self.W.img_proj.w1, bias=self.W.img_proj.b1, act=ops.ACT_GELU_ERF)
            ops.gemm(hid, self.W.img_proj.w2, bias=self.W.img_proj.b2, out=proj[s * plan.tpf:e * plan.tpf])
            del X, hid
        hp, wp = (h // m, w // m) if self.gemma else (m, m)
        th = self.pos_table("h", hp, 0, hp, m)
        tw = self.pos_table("w", wp, 0, wp, m)
        tt = self.pos_table("t", fl, plan.f0, plan.F, cfg.mm_time_interval)
        out, mask = ops.mm_finish(proj, self.W.img_norm, self.W.llm_norm, [th, tw, tt], [wp, 1, hp * wp], [hp, wp, BIG],
                                  [0, 0, 0], 0, sample_valid, self.normalizer, cfg.mm_eps, out=proj)
        return out, mask

    def encode_audios(self, mels: torch.Tensor, plan: ShardPlan, sample_valid: bool = True):
        """This rank's chunks [c1-c0,128,3000] bf16 -> (stream rows [n_aud, D], mask)  (multimodal.py:210-252)."""
        cfg, a = self.cfg, self.cfg.aud
        D, k = cfg.llm.hidden, cfg.mm_audio_pool_size
        cl = plan.c1 - plan.c0
        assert mels.shape[0] == cl and mels.dtype == BF16
        n = plan.n_aud
        feats = torch.empty(cl * a.max_source_positions, a.d_model, device=self.device, dtype=BF16)
        for s in range(0, cl, self.aud_chunk):
            e = min(cl, s + self.aud_chunk)
            feats[s * a.max_source_positions:e * a.max_source_positions] = self.whisper(mels[s:e])
        if n == 0:
            return torch.empty(0, D, device=self.device, dtype=BF16), torch.empty(0, device=self.device, dtype=torch.uint8)
        A = feats.view(-1)[: n * k * a.d_model].view(n, k * a.d_model)         # Conv1d(k=s=5) == reshape + GEMM
        pooled = ops.gemm(A, self.W.aud_pool)
        proj = self._project(pooled, self.W.aud_proj)
        tt = self.pos_table("t", n, plan.a0, plan.n_aud_total, cfg.mm_time_interval)
        return ops.mm_finish(proj, self.W.aud_norm, self.W.llm_norm, [tt], [1], [BIG], [0], 0, sample_valid,
                             self.normalizer, cfg.mm_eps, out=proj)

    # ------------------------------------------------------------------------------------------
    # decoder: stream pass
    # ------------------------------------------------------------------------------------------
    def _side_stream(self):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def _llm_cta2(self):
        """GEMM variant of the stream-pass sites: None = automatic (CTA pair for M >= 1024 and 128..256-wide tiles), False = always 1-CTA"""
        return None if self.llm_cta2 else False

    def stream_pass(self, S: torch.Tensor, kv: Optional[torch.Tensor] = None, on_kv=None) -> torch.Tensor:
        """S [n, D] (modified in place) -> K||V cache [L, n, 2*kv_dim] bf16.
        Per layer (gemma.py:183-202 with Q5/Q6 of SURVEY 3.4 dropped): K||V = G(S,w_in) W_kv^T;
        S += G(V W_o'^T, w_post); S += G(MLP(G(S,w_pre)), w_postff).  The last layer only needs K||V."""
        c = self.cfg.llm
        n, D = S.shape
        Ls = self.W.layers
        if kv is None:
            kv = torch.empty(len(Ls), n, 2 * c.kv_dim, device=self.device, dtype=BF16)
        if n == 0:
            return kv
        gm = self.gemma
        h = ops.rmsnorm(S, Ls[0].n_in, c.rms_eps, gm)
        y = torch.empty_like(S)
        g = torch.empty(n, c.inter, device=self.device, dtype=BF16)
        for l, L in enumerate(Ls):
            ops.gemm(h, L.wkv, out=kv[l], tag="llm_kv", cta2=self._llm_cta2())
            if on_kv is not None:
                on_kv(l)                 # K||V of layer l is enqueued: the text stream may consume it
            if l == len(Ls) - 1:
                break
            ops.gemm(kv[l][:, c.kv_dim:], L.wo_fold, out=y, tag="llm_vo", cta2=self._llm_cta2())
            if gm:      # S += G(y, w_post); h = G(S, w_preff)          (gemma.py:198-202,116-118)
                ops.residual_norm(S, y, L.n_post, L.n_preff, h, c.rms_eps, 1, True)
            else:       # S += y; h = norm(S, w_post_attention)          (mistral.py:223-225,131-133)
                ops.residual_norm(S, y, None, L.n_post, h, c.rms_eps, 0, False)
            ops.gemm(h, L.wgu, glu=self.glu, out=g, tag="llm_gateup", cta2=self._llm_cta2())
            ops.gemm(g, L.wd, out=y, tag="llm_down", cta2=self._llm_cta2())
            if gm:
                ops.residual_norm(S, y, L.n_postff, Ls[l + 1].n_in, h, c.rms_eps, 1, True)
            else:
                ops.residual_norm(S, y, None, Ls[l + 1].n_in, h, c.rms_eps, 0, False)
        return kv

    # ------------------------------------------------------------------------------------------
    # decoder: text pass
    # ------------------------------------------------------------------------------------------
    def new_text_cache(self, max_len: int) -> dict:
        c = self.cfg.llm
        return dict(kv=torch.empty(c.layers, max_len, 2 * c.kv_dim, device=self.device, dtype=BF16), len=0)

    def text_pass(self, ids: torch.Tensor, kv: torch.Tensor, seg: list, text_cache: Optional[dict] = None,
                  logits_to_keep: int = 0) -> torch.Tensor:
        """ids [Tq] int64 (sentinel already stripped) -> logits fp32 [Tq or k, vocab].
        seg: list of (row0, n_rows, kmask or None, gate, n_total) describing the image / audio row ranges of the
        local K||V cache.  (gemma.py:160-175, 185-192, 206-221, 236-238, 564-569)"""
        if self.native_text and (self.world == 1 or not seg or (self.xchg is not None and
                                                                self.xchg.fits(len(seg), ids.numel() * self.cfg.llm.heads, self.cfg.llm.head_dim))):
            return self._text_pass_native(ids, kv, seg, text_cache, logits_to_keep)
        run = _TextRun(self, ids, kv, seg, text_cache, logits_to_keep)
        for l in range(len(self.W.layers)):
            run.layer(l)
        return run.finish()

    def split_plan(self, seg) -> list:
        """key splits per stream segment: together they fill the SMs once (one launch covers both segments), shared in proportion to
        the segments' key counts; sized from the per-rank share of the GLOBAL counts so every rank uses the same plan"""
        c = self.cfg.llm
        return [max(1, s + self.split_bias) for s in ops.xattn_split_plan([-(-g[4] // self.world) for g in seg], c.kv_heads, self.n_sms)]

    def _layer_table(self):
        """HOST array of per-layer weight pointers for vidi_text_pass (built once)"""
        if getattr(self, "_ltab", None) is None:
            from .lib import VidiTextLayerW
            tab = (VidiTextLayerW * len(self.W.layers))()
            for t, L in zip(tab, self.W.layers):
                t.wqkv, t.wo, t.wgu, t.wd = L.wqkv.data_ptr(), L.wo.data_ptr(), L.wgu.data_ptr(), L.wd.data_ptr()
                t.n_in, t.n_post = L.n_in.data_ptr(), L.n_post.data_ptr()
                t.n_preff = L.n_preff.data_ptr() if hasattr(L, "n_preff") else None
                t.n_postff = L.n_postff.data_ptr() if hasattr(L, "n_postff") else None
            self._ltab = tab
        return self._ltab

    def _text_pass_native(self, ids, kv, seg, text_cache, logits_to_keep):
        """engine.text_pass through the native executor (csrc/textpass.cu): one C call enqueues every layer's kernels -- the same
        kernels with the same arguments as _TextRun, so the results are bit-identical to the per-layer Python path."""
        import ctypes as C
        from . import lib as _lib
        from .lib import VidiTextPass
        c = self.cfg.llm
        gm = self.gemma
        Tq = ids.numel()
        d = VidiTextPass()
        d.Tq, d.pos0, d.layers, d.hidden, d.heads, d.kv_heads, d.head_dim = Tq, 0, c.layers, c.hidden, c.heads, c.kv_heads, c.head_dim
        d.inter, d.vocab, d.gemma, d.glu = c.inter, c.vocab, int(gm), self.glu
        d.sliding_window = int(getattr(c, "sliding_window", 0) or 0)
        d.logits_keep = int(logits_to_keep or 0)
        d.rms_eps, d.normalizer = c.rms_eps, self.normalizer
        d.scale = c.query_pre_attn_scalar ** -0.5 if gm else c.head_dim ** -0.5
        d.attn_softcap = (c.attn_softcap or 0.0) if gm else 0.0
        d.final_softcap = (c.final_softcap or 0.0) if gm else 0.0
        d.layer_w = self._layer_table()
        d.embed, d.final_norm, d.lm_head = self.W.embed.data_ptr(), self.W.final_norm.data_ptr(), self.W.lm_head.data_ptr()
        d.inv_freq, d.ids = self.W.inv_freq.data_ptr(), ids.data_ptr()
        assert ids.dtype == torch.int64 and ids.is_contiguous()
        if text_cache is not None:
            tkv = text_cache["kv"]
            d.pos0 = text_cache["len"]
            assert d.pos0 + Tq <= tkv.shape[1], "text KV cache too small"
            d.text_kv, d.text_kv_layer_stride, d.text_kv_ld = tkv.data_ptr(), tkv.stride(0), tkv.stride(1)
        d.stream_kv = kv.data_ptr() if kv.numel() else None
        d.stream_layer_stride, d.stream_ld = (kv.stride(0), kv.stride(1)) if kv.numel() else (0, 2 * c.kv_dim)
        d.nseg, d.stream_rows = len(seg), kv.shape[1]
        for i, ((r0, nr, kmask, gate, n_total), sp) in enumerate(zip(seg, self.split_plan(seg))):
            sg = d.seg[i]
            sg.row0, sg.rows, sg.gate, sg.splits = r0, nr, gate, sp
            sg.kmask = kmask.data_ptr() if (kmask is not None and kmask.numel()) else None
        multi = self.world > 1 and len(seg) > 0           # a text-only query has nothing to exchange: every rank runs it alone
        d.world, d.rank = (self.world, self.rank) if multi else (1, 0)
        x = self.xchg
        if multi:
            d.seq0, d.cap = x.seq, x.cap
            for r in range(self.world):
                d.peer_data[r] = x.arenas[r]
                d.peer_flags[r] = x.arenas[r] + x.flags_off
            d.counter, d.err = x.counter_ptr, x.err_ptr
        L = _lib.load()
        need = int(L.vidi_text_pass_workspace_bytes(C.byref(d)))
        ws = getattr(self, "_text_ws", None)
        if ws is None or ws.numel() < need:
            ws = self._text_ws = torch.empty(need, device=self.device, dtype=torch.uint8)
        keep = Tq if not logits_to_keep or logits_to_keep >= Tq else logits_to_keep
        logits = torch.empty(keep, c.vocab, device=self.device, dtype=torch.float32)
        d.workspace, d.workspace_bytes, d.logits = ws.data_ptr(), ws.numel(), logits.data_ptr()
        _lib.check(L.vidi_text_pass(C.byref(d), ops._stream()), "text_pass")
        if multi:
            x.seq += c.layers
        if text_cache is not None:
            text_cache["len"] = d.pos0 + Tq
        return logits

    # ------------------------------------------------------------------------------------------
    # whole prefill for one sample
    # ------------------------------------------------------------------------------------------
    def encode_streams(self, images, mels, plan: ShardPlan, image_valid=True, audio_valid=True):
        """-> (S [n_img+n_aud, D], seg list for text_pass)"""
        parts, seg = [], []
        r0 = 0
        if images is not None:
            X, mX = self.encode_images(images, plan, image_valid)
            parts.append(X)
            seg.append((r0, X.shape[0], mX if image_valid else None, 1.0 if image_valid else 0.0, plan.n_img_total))
            r0 += X.shape[0]
        if mels is not None:
            A, mA = self.encode_audios(mels, plan, audio_valid)
            parts.append(A)
            seg.append((r0, A.shape[0], mA if audio_valid else None, 1.0 if audio_valid else 0.0, plan.n_aud_total))
            r0 += A.shape[0]
        if not parts:
            return torch.empty(0, self.cfg.llm.hidden, device=self.device, dtype=BF16), seg
        S = parts[0] if len(parts) == 1 else torch.cat(parts, 0)
        return S, seg

    @torch.no_grad()
    def prefill(self, ids: torch.Tensor, images: Optional[torch.Tensor], mels: Optional[torch.Tensor], audio_size: int,
                n_frames_total: Optional[int] = None, n_chunks_total: Optional[int] = None, logits_to_keep: int = 0,
                text_cache: Optional[dict] = None, image_valid=True, audio_valid=True, return_state: bool = False):
        """ids: [T] int64 device tensor without the -200 sentinel.  images / mels are THIS RANK's shard
        (see make_plan) in bf16 on the device.  Returns logits fp32 [T or k, vocab]."""
        F = n_frames_total if n_frames_total is not None else (images.shape[0] if images is not None else 0)
        Cn = n_chunks_total if n_chunks_total is not None else (mels.shape[0] if mels is not None else 0)
        plan = make_plan(self.cfg, F, Cn, audio_size or 0, self.rank, self.world)
        S, seg = self.encode_streams(images, mels, plan, image_valid, audio_valid)
        if self.overlap_text and S.shape[0] > 0:
            # The text stream's layer l only needs K||V of layer l: run it on a side stream one step behind the stream
            # pass, so its ~500 small launches (and, multi-GPU, the per-layer exchange) hide under the big GEMMs.
            c = self.cfg.llm
            kv = torch.empty(c.layers, S.shape[0], 2 * c.kv_dim, device=self.device, dtype=BF16)
            main = torch.cuda.current_stream()
            side = self._side_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                run = _TextRun(self, ids, kv, seg, text_cache, logits_to_keep)

            def on_kv(l):
                ev = torch.cuda.Event()
                ev.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(ev)
                    run.layer(l)
            self.stream_pass(S, kv=kv, on_kv=on_kv)
            with torch.cuda.stream(side):
                logits = run.finish()
            main.wait_stream(side)
            logits.record_stream(main)
        else:
            kv = self.stream_pass(S)
            logits = self.text_pass(ids, kv, seg, text_cache=text_cache, logits_to_keep=logits_to_keep)
        if self.xchg is not None:
            self.xchg.check()
        if return_state:
            return logits, dict(kv=kv, seg=seg, plan=plan, streams=S)
        return logits


class _TextRun:
    """The text stream of one prefill / decode step, advanced layer by layer (so that it can trail the stream pass on a
    side stream).  Semantics: gemma.py:160-175 (T2T), :185-192 / :206-221 (T2V / T2A), :236-238 (combine), :564-569 (lm_head);
    Mistral family: mistral.py:190-264."""

    def __init__(self, eng: Vidi15Engine, ids: torch.Tensor, kv: torch.Tensor, seg: list, text_cache: Optional[dict],
                 logits_to_keep: int):
        self.e, self.kv, self.seg, self.text_cache, self.keep = eng, kv, seg, text_cache, logits_to_keep
        c = eng.cfg.llm
        self.c = c
        self.Tq = ids.numel()
        self.pos0 = 0
        if text_cache is not None:
            self.pos0 = text_cache["len"]
            assert self.pos0 + self.Tq <= text_cache["kv"].shape[1], "text KV cache too small"
        gm = eng.gemma
        self.scale = c.query_pre_attn_scalar ** -0.5 if gm else c.head_dim ** -0.5
        self.cap = (c.attn_softcap or 0.0) if gm else 0.0
        Ls = eng.W.layers
        self.H = ops.embed_gather(ids, eng.W.embed, eng.normalizer)
        self.h = ops.rmsnorm(self.H, Ls[0].n_in, c.rms_eps, gm)
        self.rows = self.Tq * c.heads
        dh = c.head_dim
        # one flat fp32 buffer per layer holds every stream's [O | LSE] split partials of this rank
        self.splits = eng.split_plan(seg)
        ptot = sum(self.splits)
        # flat = O [P0+P1][rows][dh] | LSE [P0+P1][rows]  (segment 1's partials follow segment 0's: ops.xattn_splitkv_seg)
        self.flat = torch.empty(max(1, ptot * self.rows * (dh + 1)), device=eng.device, dtype=torch.float32)
        self.O_all = self.flat[:ptot * self.rows * dh]
        self.L_all = self.flat[ptot * self.rows * dh:]
        # multi-rank: each rank first reduces its own splits to ONE (O, LSE) partial per stream (xchg.cu), and only those cross ranks:
        #   "p2p"  -- stored straight into every peer's arena over NVLink, merge kernel waits on flags (no NCCL, no host sync)
        #   "nccl" -- one all_gather_into_tensor of the reduced block per layer (used when the arenas cannot be mapped)
        self.mode = "local"
        if eng.world > 1 and seg:
            x = eng.xchg
            self.mode = "p2p" if (x is not None and x.fits(len(seg), self.rows, dh)) else "nccl"
            if self.mode == "nccl":
                assert eng.group is not None or torch.distributed.is_initialized(), "multi-rank text pass needs a process group or a peer arena"
                self.block = len(seg) * self.rows * (dh + 1)
                self.pre = torch.empty(self.block, device=eng.device, dtype=torch.float32)
                self.gathered = torch.empty(eng.world * self.block, device=eng.device, dtype=torch.float32)
        self.att = torch.empty(self.Tq, c.q_dim, device=eng.device, dtype=torch.float32)
        self.y = torch.empty(self.Tq, c.hidden, device=eng.device, dtype=BF16)
        self.h2 = torch.empty_like(self.H)
        self.a = torch.empty(self.Tq, c.q_dim, device=eng.device, dtype=BF16)
        self.qrope = torch.empty(self.Tq, c.q_dim, device=eng.device, dtype=BF16)
        self.krope = torch.empty(self.Tq, 2 * c.kv_dim, device=eng.device, dtype=BF16) if text_cache is None else None

    def layer(self, l: int):
        self.layer_begin(l)
        self.layer_end(l)

    def layer_begin(self, l: int):
        """text self-attention + this rank's cross-attention partials of layer l, sent on their way to the peers"""
        e, c = self.e, self.c
        gm = e.gemma
        L = e.W.layers[l]
        Tq, pos0, rows = self.Tq, self.pos0, self.rows
        qd, kd, dh = c.q_dim, c.kv_dim, c.head_dim
        qkv = ops.gemm(self.h, L.wqkv, tag="text")
        # one launch: q_rope = RoPE(q); text K||V rows = RoPE(k) | v (straight into the cache when there is one)
        if self.text_cache is not None:
            tkv = self.text_cache["kv"][l]
            ops.text_qk_prep(qkv, self.qrope, tkv[pos0:pos0 + Tq], c.heads, c.kv_heads, dh, e.W.inv_freq, pos0)
            kview, vview = tkv[:pos0 + Tq, :kd], tkv[:pos0 + Tq, kd:]
        else:
            ops.text_qk_prep(qkv, self.qrope, self.krope, c.heads, c.kv_heads, dh, e.W.inv_freq, pos0)
            kview, vview = self.krope[:, :kd], self.krope[:, kd:]
        window = (c.sliding_window if l % 2 == 0 else 0) if gm else (getattr(c, "sliding_window", 0) or 0)
        ops.attn_text(self.qrope, kview, vview, pos0, c.heads, c.kv_heads, dh, self.scale, self.cap, window, out=self.att)
        kvl = self.kv[l]
        srcs = []
        if self.seg:
            ops.xattn_splitkv_seg(qkv[:, :qd], kvl[:, :kd], kvl[:, kd:], [(s[0], s[1], s[2]) for s in self.seg], self.splits, c.heads,
                                  c.kv_heads, dh, self.scale, self.cap, self.O_all, self.L_all)
            p0 = 0
            for sp in self.splits:
                srcs.append((self.O_all[p0 * rows * dh:], self.L_all[p0 * rows:], sp))
                p0 += sp
        self.srcs = srcs
        if self.mode == "p2p":
            ops.xchg_push(e.xchg, srcs, rows, dh)
        elif self.mode == "nccl":
            ops.xattn_premerge(srcs, rows, dh, self.pre)
            torch.distributed.all_gather_into_tensor(self.gathered, self.pre, group=e.group)

    def layer_end(self, l: int):
        """merge the partials of all ranks, o_proj, residual + MLP of layer l"""
        e, c = self.e, self.c
        gm = e.gemma
        Ls = e.W.layers
        L = Ls[l]
        rows, dh = self.rows, c.head_dim
        gates = [s[3] for s in self.seg]
        # one launch: a = bf16(att_text + gate_img * merge(img partials) + gate_aud * merge(aud partials))
        if self.mode == "p2p":
            a = ops.xchg_merge(e.xchg, gates, self.att, self.a, rows, dh)
        elif self.mode == "nccl":
            srcs = []
            for si, gate in enumerate(gates):
                o = self.gathered[si * rows * (dh + 1):]
                srcs.append((o, o[rows * dh:], e.world, 1, self.block, self.block, gate))
            a = ops.xattn_merge2(srcs, self.att, self.a, rows, dh)
        else:
            srcs = [(o, l_, sp, sp, 0, 0, gate) for (o, l_, sp), gate in zip(self.srcs, gates)]
            a = ops.xattn_merge2(srcs, self.att, self.a, rows, dh)
        ops.gemm(a, L.wo, out=self.y, tag="text")
        w_next = Ls[l + 1].n_in if l + 1 < len(Ls) else e.W.final_norm
        if gm:
            ops.residual_norm(self.H, self.y, L.n_post, L.n_preff, self.h2, c.rms_eps, 1, True)
        else:       # H = residual + (a_t + a_i + a_a) W_o^T ; h2 = post_attention_layernorm(H)   (mistral.py:263,131-133)
            ops.residual_norm(self.H, self.y, None, L.n_post, self.h2, c.rms_eps, 0, False)
        g = ops.gemm(self.h2, L.wgu, glu=e.glu, tag="text")
        ops.gemm(g, L.wd, out=self.y, tag="text")
        if gm:
            ops.residual_norm(self.H, self.y, L.n_postff, w_next, self.h, c.rms_eps, 1, True)
        else:
            ops.residual_norm(self.H, self.y, None, w_next, self.h, c.rms_eps, 0, False)

    def finish(self) -> torch.Tensor:
        e, c = self.e, self.c
        if self.text_cache is not None:
            self.text_cache["len"] = self.pos0 + self.Tq
        hn = self.h if not self.keep else self.h[-self.keep:]
        if e.gemma:     # 30 * tanh(logits / 30)   (gemma.py:566-569)
            return ops.gemm(hn, e.W.lm_head, act=ops.ACT_SOFTCAP, act_param=c.final_softcap, out_fp32=True, tag="text")
        return ops.gemm(hn, e.W.lm_head, out_fp32=True, tag="text")      # logits.float() (mistral.py:615-616)
