// Native executor of the text stream (one call = all decoder layers of one prefill / decode step), see include/vidi_b200.h.
// Host code only: it sequences the library's own kernels (GEMM, fused RoPE prep, text attention, split-KV cross attention,
// pre-merge + peer push, flag-wait merge, residual+norm) on one stream.  Reference: DattnGemma2DecoderLayer.forward for the text
// rows (Vidi1.5_9B/vidi/model/lmm/dattn/gemma.py:125-244), DattnGemma2Model.forward (:362-411), lm_head + soft-cap (:564-569);
// Mistral family Vidi_7B/model/lmm/dattn/mistral.py:190-264, 615-616.
#include "../../include/vidi_b200.h"
#include "common.cuh"

namespace vb {
int gemm_bf16(const void*, int64_t, const void*, int64_t, void*, int64_t, int, int, int, const float*, const void*, int64_t,
              int, int, float, int, int, int, cudaStream_t);
int rmsnorm(const void*, int64_t, const void*, void*, int64_t, int, int, float, int, float, cudaStream_t);
int residual_norm(void*, int64_t, const void*, int64_t, const void*, const void*, void*, int64_t, int, int, float, int, int,
                  cudaStream_t);
int embed_gather(const int64_t*, const void*, void*, int, int, int, float, cudaStream_t);
int xattn_splitkv_seg(const void*, int64_t, const void*, const void*, int64_t, int, int, const int*, const int*, const int*,
                      const uint8_t* const*, int, int, int, int, float, float, float*, float*, int, int*, cudaStream_t);
int text_qk_prep(const void*, int64_t, void*, int64_t, void*, int64_t, int, int, int, int, const float*, int, cudaStream_t);
int xattn_merge2(const float*, const float*, int, int, int64_t, int64_t, float, const float*, const float*, int, int, int64_t, int64_t,
                 float, int, const float*, int, int, void*, const unsigned int*, int, unsigned int, int*, cudaStream_t);
int xattn_premerge_push(const float*, const float*, int, const float*, const float*, int, int, int, int, float* const*,
                        unsigned int* const*, int, int64_t, unsigned int, unsigned int*, cudaStream_t);
int attn_text(const void*, int64_t, const void*, const void*, int64_t, int, int, int, int, int, int, float, float, int, float*,
              cudaStream_t);

namespace {

struct Carve {
    uint8_t* base;
    int64_t off = 0;
    template <typename T>
    T* take(int64_t n) {
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += ((n * (int64_t)sizeof(T) + 255) / 256) * 256;
        return p;
    }
};

struct Scratch {
    __nv_bfloat16 *H, *h, *h2, *y, *qkv, *qrope, *krope, *a, *g;
    float *att, *flat;
    int64_t bytes;
};

Scratch carve(const VidiTextPass& d, void* ws) {
    Carve c{reinterpret_cast<uint8_t*>(ws)};
    const int64_t T = d.Tq, D = d.hidden, qd = (int64_t)d.heads * d.head_dim, kd = (int64_t)d.kv_heads * d.head_dim;
    const int64_t rows = T * d.heads;
    Scratch s;
    s.H = c.take<__nv_bfloat16>(T * D);
    s.h = c.take<__nv_bfloat16>(T * D);
    s.h2 = c.take<__nv_bfloat16>(T * D);
    s.y = c.take<__nv_bfloat16>(T * D);
    s.qkv = c.take<__nv_bfloat16>(T * (qd + 2 * kd));
    s.qrope = c.take<__nv_bfloat16>(T * qd);
    s.krope = c.take<__nv_bfloat16>(T * 2 * kd);
    s.a = c.take<__nv_bfloat16>(T * qd);
    s.g = c.take<__nv_bfloat16>(T * (int64_t)d.inter);
    s.att = c.take<float>(T * qd);
    int64_t flat = 0;
    for (int i = 0; i < d.nseg; ++i) flat += (int64_t)d.seg[i].splits * rows * (d.head_dim + 1);
    s.flat = c.take<float>(flat > 0 ? flat : 1);
    s.bytes = c.off;
    return s;
}

// column tile of the 1-CTA GEMM, as ops.pick_block_n chooses it for the Python path (same tiles -> same bits)
int pick_block_n(int M, int N, bool glu) {
    if (glu) return 256;
    if (M <= 256 && N <= 16384) return 64;
    if (N % 256 != 0 && N % 192 == 0 && N < 2048) return 192;
    if (N >= 1024 || N % 256 == 0) return 256;
    return N > 64 ? 128 : 64;
}

int gemm(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K, int act, float act_param,
         int out_fp32, int glu, cudaStream_t st) {
    return gemm_bf16(A, lda, W, ldw, C, ldc, M, N, K, nullptr, nullptr, 0, 0, act, act_param, out_fp32, glu, pick_block_n(M, N, glu != 0), st);
}

}  // namespace

int64_t text_pass_workspace_bytes(const VidiTextPass* d) { return carve(*d, nullptr).bytes; }

#define TP(expr)                       \
    do {                               \
        int rc_ = (expr);              \
        if (rc_ != 0) return rc_;      \
    } while (0)

int text_pass(const VidiTextPass* dp, int64_t* launches, cudaStream_t st) {
    const VidiTextPass& d = *dp;
    VB_REQUIRE(d.Tq > 0 && d.layers > 0 && d.nseg >= 0 && d.nseg <= 2 && d.layer_w && d.workspace && d.logits, "text_pass: bad descriptor");
    VB_REQUIRE((reinterpret_cast<uintptr_t>(d.workspace) & 255) == 0, "text_pass: workspace must be 256-byte aligned");
    const Scratch s = carve(d, d.workspace);
    VB_REQUIRE(s.bytes <= d.workspace_bytes, "text_pass: workspace too small (%lld < %lld)", (long long)d.workspace_bytes, (long long)s.bytes);
    VB_REQUIRE(d.text_kv != nullptr || d.pos0 == 0, "text_pass: pos0 > 0 needs a text K||V cache");
    VB_REQUIRE(d.world == 1 || (d.world > 1 && d.world <= 16 && d.counter && d.err && d.nseg > 0), "text_pass: bad exchange arguments");
    const int T = d.Tq, D = d.hidden, dh = d.head_dim, Hq = d.heads, Hkv = d.kv_heads;
    const int qd = Hq * dh, kd = Hkv * dh, rows = T * Hq;
    const bool gm = d.gemma != 0;
    if (d.world > 1) VB_REQUIRE((int64_t)d.nseg * rows * (dh + 1) <= d.cap, "text_pass: exchange arena too small for %d text rows", T);
    int64_t n = 0;

    TP(embed_gather(d.ids, d.embed, s.H, T, D, d.vocab, d.normalizer, st)); ++n;
    TP(rmsnorm(s.H, D, d.layer_w[0].n_in, s.h, D, T, D, d.rms_eps, gm ? 1 : 0, 1.0f, st)); ++n;
    for (int l = 0; l < d.layers; ++l) {
        const VidiTextLayerW& W = d.layer_w[l];
        TP(gemm(s.h, D, W.wqkv, D, s.qkv, qd + 2 * kd, T, qd + 2 * kd, D, 0, 0.f, 0, 0, st)); ++n;
        // q_rope = RoPE(q); text K||V rows = RoPE(k) | v, straight into the cache when there is one
        __nv_bfloat16* tkv = d.text_kv ? reinterpret_cast<__nv_bfloat16*>(d.text_kv) + (int64_t)l * d.text_kv_layer_stride : s.krope;
        const int64_t tld = d.text_kv ? d.text_kv_ld : 2 * kd;
        TP(text_qk_prep(s.qkv, qd + 2 * kd, s.qrope, qd, tkv + (int64_t)d.pos0 * tld, tld, T, Hq, Hkv, dh, d.inv_freq, d.pos0, st)); ++n;
        const int window = gm ? ((l % 2 == 0) ? d.sliding_window : 0) : d.sliding_window;
        TP(attn_text(s.qrope, qd, tkv, tkv + kd, tld, T, d.pos0 + T, d.pos0, Hq, Hkv, dh, d.scale, d.attn_softcap, window, s.att, st)); ++n;
        // cross attention partials of this rank: image rows, then audio rows
        const __nv_bfloat16* kvl = reinterpret_cast<const __nv_bfloat16*>(d.stream_kv) + (int64_t)l * d.stream_layer_stride;
        // both segments in one call (one launch on the tcgen05 path): flat = O [P0+P1][rows][dh] | LSE [P0+P1][rows]
        const float *O[2] = {nullptr, nullptr}, *L[2] = {nullptr, nullptr};
        if (d.nseg > 0) {
            int row0[2] = {0, 0}, nrows[2] = {0, 0}, sp[2] = {1, 1};
            const uint8_t* masks[2] = {nullptr, nullptr};
            int ptot = 0;
            for (int i = 0; i < d.nseg; ++i) {
                row0[i] = (int)d.seg[i].row0; nrows[i] = d.seg[i].rows; sp[i] = d.seg[i].splits; masks[i] = d.seg[i].kmask;
                ptot += sp[i];
            }
            float* Oall = s.flat;
            float* Lall = s.flat + (int64_t)ptot * rows * dh;
            int nl = 1;
            TP(xattn_splitkv_seg(s.qkv, qd + 2 * kd, kvl, kvl + kd, d.stream_ld, d.stream_rows, d.nseg, row0, nrows, sp, masks, T, Hq, Hkv, dh,
                                 d.scale, d.attn_softcap, Oall, Lall, 0, &nl, st));
            n += nl;
            O[0] = Oall; L[0] = Lall;
            if (d.nseg > 1) { O[1] = Oall + (int64_t)sp[0] * rows * dh; L[1] = Lall + (int64_t)sp[0] * rows; }
        }
        // a = bf16(att_text + sum_s gate_s * merge_s): one launch; multi-rank: pre-merge + push to the peers, then flag-wait merge
        const float g0 = d.nseg > 0 ? d.seg[0].gate : 0.f, g1 = d.nseg > 1 ? d.seg[1].gate : 0.f;
        const int p0 = d.nseg > 0 ? d.seg[0].splits : 0, p1 = d.nseg > 1 ? d.seg[1].splits : 0;
        if (d.world > 1) {
            const unsigned int seq = d.seq0 + (unsigned int)l + 1u;
            const int slot = (int)(seq & 1u);
            float* base[16]; unsigned int* flag[16];
            for (int r = 0; r < d.world; ++r) {
                base[r] = d.peer_data[r] + (int64_t)slot * d.world * d.cap;
                flag[r] = d.peer_flags[r] + slot * d.world + d.rank;
            }
            TP(xattn_premerge_push(O[0], L[0], p0, O[1], L[1], p1, d.nseg, rows, dh, base, flag, d.world, (int64_t)d.rank * d.cap, seq,
                                   d.counter, st)); ++n;
            const float* mine = d.peer_data[d.rank] + (int64_t)slot * d.world * d.cap;
            const float* o1 = mine + (int64_t)rows * (dh + 1);
            TP(xattn_merge2(mine, mine + (int64_t)rows * dh, d.world, 1, d.cap, d.cap, g0, d.nseg > 1 ? o1 : nullptr,
                            d.nseg > 1 ? o1 + (int64_t)rows * dh : nullptr, d.world, 1, d.cap, d.cap, g1, d.nseg, s.att, rows, dh, s.a,
                            d.peer_flags[d.rank] + slot * d.world, d.world, seq, d.err, st)); ++n;
        } else {
            TP(xattn_merge2(O[0], L[0], p0, p0 > 0 ? p0 : 1, 0, 0, g0, O[1], L[1], p1, p1 > 0 ? p1 : 1, 0, 0, g1, d.nseg, s.att, rows, dh, s.a,
                            nullptr, 0, 0u, nullptr, st)); ++n;
        }
        TP(gemm(s.a, qd, W.wo, qd, s.y, D, T, D, qd, 0, 0.f, 0, 0, st)); ++n;
        const void* w_next = l + 1 < d.layers ? d.layer_w[l + 1].n_in : d.final_norm;
        if (gm) { TP(residual_norm(s.H, D, s.y, D, W.n_post, W.n_preff, s.h2, D, T, D, d.rms_eps, 1, 1, st)); }
        else    { TP(residual_norm(s.H, D, s.y, D, nullptr, W.n_post, s.h2, D, T, D, d.rms_eps, 0, 0, st)); }
        ++n;
        TP(gemm(s.h2, D, W.wgu, D, s.g, d.inter, T, 2 * d.inter, D, 0, 0.f, 0, d.glu, st)); ++n;
        TP(gemm(s.g, d.inter, W.wd, d.inter, s.y, D, T, D, d.inter, 0, 0.f, 0, 0, st)); ++n;
        if (gm) { TP(residual_norm(s.H, D, s.y, D, W.n_postff, w_next, s.h, D, T, D, d.rms_eps, 1, 1, st)); }
        else    { TP(residual_norm(s.H, D, s.y, D, nullptr, w_next, s.h, D, T, D, d.rms_eps, 0, 0, st)); }
        ++n;
    }
    const int keep = d.logits_keep > 0 && d.logits_keep < T ? d.logits_keep : T;
    const __nv_bfloat16* hn = s.h + (int64_t)(T - keep) * D;
    if (gm) { TP(gemm(hn, D, d.lm_head, D, d.logits, d.vocab, keep, d.vocab, D, VIDI_ACT_SOFTCAP, d.final_softcap, 1, 0, st)); }
    else    { TP(gemm(hn, D, d.lm_head, D, d.logits, d.vocab, keep, d.vocab, D, 0, 0.f, 1, 0, st)); }
    ++n;
    if (launches) *launches = n;
    return 0;
}

}  // namespace vb
