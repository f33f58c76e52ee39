// extern "C" surface of libvidi_b200.so (declared in include/vidi_b200.h).  Thin forwarding layer: validates nothing
// beyond what the launchers check, counts launches, never synchronises.
#include <atomic>

#include "../../include/vidi_b200.h"
#include "common.cuh"

namespace vb {
const char* last_error();
int gemm_bf16(const void*, int64_t, const void*, int64_t, void*, int64_t, int, int, int, const float*, const void*, int64_t,
              int, int, float, int, int, int, cudaStream_t);
int gemm2_bf16_ln(const void*, int64_t, const void*, int64_t, void*, int64_t, int, int, int, const float*, const void*, int64_t,
                  int, int, float, int, int, int, const float*, int, const float*, float, float*, cudaStream_t);
int gemm2_bf16(const void*, int64_t, const void*, int64_t, void*, int64_t, int, int, int, const float*, const void*, int64_t,
               int, int, float, int, int, int, cudaStream_t);
int rmsnorm(const void*, int64_t, const void*, void*, int64_t, int, int, float, int, float, cudaStream_t);
int resample_u8(const uint8_t*, uint8_t*, int64_t, int, int, int, const int*, const int*, int, cudaStream_t);
int resample_u8_to_chw_bf16(const uint8_t*, void*, int, int, int, int, const int*, const int*, int, float, float, float, cudaStream_t);
int logmel_frames(const float*, const float*, void*, int, int, cudaStream_t);
int logmel_power(const float*, int64_t, void*, int64_t, cudaStream_t);
int logmel_finish(const float*, int, int, float*, void*, cudaStream_t);
int residual_norm(void*, int64_t, const void*, int64_t, const void*, const void*, void*, int64_t, int, int, float, int, int,
                  cudaStream_t);
int layernorm(const void*, int64_t, const float*, const float*, void*, int64_t, int, int, float, cudaStream_t);
int mm_finish(const void*, int64_t, const void*, const void*, const float* const*, const int*, const int*, const int*, int,
              int, int, float, void*, int64_t, uint8_t*, int, int, float, cudaStream_t);
int rmsnorm_f32(const float*, float*, int, int, float, int, cudaStream_t);
int patch_im2col(const void*, void*, int, int, int, int, cudaStream_t);
int whisper_im2col1(const void*, void*, int, int, int, cudaStream_t);
int whisper_im2col2(const void*, void*, int, int, int, cudaStream_t);
int pool_s2d(const void*, void*, int, int, int, int, int, int, cudaStream_t);
int conv_window_gather(const void*, void*, int, int, int, int, cudaStream_t);
int bilinear_ac(const void*, void*, int, int, int, int, cudaStream_t);
int embed_gather(const int64_t*, const void*, void*, int, int, int, float, cudaStream_t);
int sinusoid_split(const float*, void*, int, int, int, int, int, cudaStream_t);
int split3(const float*, void*, int64_t, int, int, cudaStream_t);
int cast_f32_bf16(const float*, void*, int64_t, cudaStream_t);
int attn_dense(const void*, int64_t, int, int, int, void*, int64_t, int, int, int, int, float, cudaStream_t);
int attn_dense_sm100(const void*, int64_t, void*, int64_t, int, int, int, int, float, cudaStream_t);
int attn_dense_sm100_pp(const void*, int64_t, void*, int64_t, int, int, int, int, float, cudaStream_t);
int attn_dense_sm100_v3(const void*, int64_t, void*, int64_t, int, int, int, int, float, cudaStream_t);
int attn_dense_sm100_v3_poly(const void*, int64_t, void*, int64_t, int, int, int, int, float, int, cudaStream_t);
int xattn_splitkv(const void*, int64_t, const void*, const void*, int64_t, const uint8_t*, int, int, int, int, int, int, float,
                  float, float*, float*, int, cudaStream_t);
int xattn_splitkv_seg(const void*, int64_t, const void*, const void*, int64_t, int, int, const int*, const int*, const int*,
                      const uint8_t* const*, int, int, int, int, float, float, float*, float*, int, int*, cudaStream_t);
int xattn_merge(const float*, const float*, int, int, int64_t, int64_t, int, int, float, int, float*, cudaStream_t);
int rope_inplace(void*, int64_t, int, int, int, int, const float*, int, cudaStream_t);
int text_qk_prep(const void*, int64_t, void*, int64_t, void*, int64_t, int, int, int, int, const float*, int, cudaStream_t);
int xattn_merge2(const float*, const float*, int, int, int64_t, int64_t, float, const float*, const float*, int, int, int64_t, int64_t,
                 float, int, const float*, int, int, void*, const unsigned int*, int, unsigned int, int*, cudaStream_t);
int xattn_premerge_push(const float*, const float*, int, const float*, const float*, int, int, int, int, float* const*,
                        unsigned int* const*, int, int64_t, unsigned int, unsigned int*, cudaStream_t);
int64_t text_pass_workspace_bytes(const VidiTextPass*);
int text_pass(const VidiTextPass*, int64_t*, cudaStream_t);
int p2p_alloc(int64_t, void**, void*);
int p2p_open(const void*, void**);
int p2p_close(void*);
int p2p_free(void*);
int attn_text(const void*, int64_t, const void*, const void*, int64_t, int, int, int, int, int, int, float, float, int, float*,
              cudaStream_t);
}  // namespace vb

static std::atomic<int64_t> g_launches{0};
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define COUNT(expr) (g_launches.fetch_add(1, std::memory_order_relaxed), (expr))

extern "C" {

const char* vidi_last_error(void) { return vb::last_error(); }
int vidi_abi_version(void) { return 2; }
int64_t vidi_launch_count(void) { return g_launches.load(); }
void vidi_reset_launch_count(void) { g_launches.store(0); }

int vidi_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K,
                   const float* bias, const void* residual, int64_t ldr, int res_mod, int act, float act_param,
                   int out_fp32, int glu, int block_n, void* stream) {
    return COUNT(vb::gemm_bf16(A, lda, W, ldw, C, ldc, M, N, K, bias, residual, ldr, res_mod, act, act_param, out_fp32, glu,
                               block_n, ST(stream)));
}
int vidi_gemm_bf16_2cta(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K,
                        const float* bias, const void* residual, int64_t ldr, int res_mod, int act, float act_param,
                        int out_fp32, int glu, int block_n, void* stream) {
    return COUNT(vb::gemm2_bf16(A, lda, W, ldw, C, ldc, M, N, K, bias, residual, ldr, res_mod, act, act_param, out_fp32, glu,
                                block_n, ST(stream)));
}
int vidi_gemm_bf16_2cta_ln(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K,
                           const float* bias, const void* residual, int64_t ldr, int res_mod, int act, float act_param,
                           int block_n, const float* ln_stats, int ln_parts, const float* ln_colsum, float ln_eps,
                           float* stats_out, void* stream) {
    return COUNT(vb::gemm2_bf16_ln(A, lda, W, ldw, C, ldc, M, N, K, bias, residual, ldr, res_mod, act, act_param, 0, 0, block_n,
                                   ln_stats, ln_parts, ln_colsum, ln_eps, stats_out, ST(stream)));
}
int vidi_resample_u8(const uint8_t* src, uint8_t* dst, int64_t outer, int in_size, int out_size, int inner, const int32_t* xmin,
                     const int32_t* kk, int ksize, void* stream) {
    return COUNT(vb::resample_u8(src, dst, outer, in_size, out_size, inner, xmin, kk, ksize, ST(stream)));
}
int vidi_resample_u8_to_chw_bf16(const uint8_t* src, void* dst, int F, int in_h, int out_h, int W, const int32_t* ymin,
                                 const int32_t* kk, int ksize, float rescale, float mean, float stdv, void* stream) {
    return COUNT(vb::resample_u8_to_chw_bf16(src, dst, F, in_h, out_h, W, ymin, kk, ksize, rescale, mean, stdv, ST(stream)));
}
int vidi_logmel_frames(const float* audio, const float* window, void* out, int C, int n_samples, void* stream) {
    return COUNT(vb::logmel_frames(audio, window, out, C, n_samples, ST(stream)));
}
int vidi_logmel_power(const float* Y, int64_t ldy, void* out, int64_t rows, void* stream) {
    return COUNT(vb::logmel_power(Y, ldy, out, rows, ST(stream)));
}
int vidi_logmel_finish(const float* M, int C, int mels, float* chunk_max, void* out, void* stream) {
    return COUNT(vb::logmel_finish(M, C, mels, chunk_max, out, ST(stream)));        /* two launches: max, then finish */
}
int vidi_rmsnorm(const void* x, int64_t ldx, const void* w, void* y, int64_t ldy, int rows, int D, float eps, int add_one,
                 float out_scale, void* stream) {
    return COUNT(vb::rmsnorm(x, ldx, w, y, ldy, rows, D, eps, add_one, out_scale, ST(stream)));
}
int vidi_residual_norm(void* x, int64_t ldx, const void* y, int64_t ldy, const void* w_post, const void* w_next, void* h,
                       int64_t ldh, int rows, int D, float eps, int post_mode, int next_add_one, void* stream) {
    return COUNT(vb::residual_norm(x, ldx, y, ldy, w_post, w_next, h, ldh, rows, D, eps, post_mode, next_add_one, ST(stream)));
}
int vidi_layernorm(const void* x, int64_t ldx, const float* w, const float* b, void* y, int64_t ldy, int rows, int D,
                   float eps, void* stream) {
    return COUNT(vb::layernorm(x, ldx, w, b, y, ldy, rows, D, eps, ST(stream)));
}
int vidi_mm_finish(const void* proj, int64_t ldp, const void* w_mod, const void* w_llm, const float* const* tabs,
                   const int* divs, const int* mods, const int* offs, int ntab, int n_offset, int sample_valid,
                   float normalizer, void* out, int64_t ldo, uint8_t* mask, int rows, int D, float eps, void* stream) {
    return COUNT(vb::mm_finish(proj, ldp, w_mod, w_llm, tabs, divs, mods, offs, ntab, n_offset, sample_valid, normalizer, out,
                               ldo, mask, rows, D, eps, ST(stream)));
}
int vidi_rmsnorm_f32(const float* x, float* y, int rows, int D, float eps, int round_bf16, void* stream) {
    return COUNT(vb::rmsnorm_f32(x, y, rows, D, eps, round_bf16, ST(stream)));
}
int vidi_patch_im2col(const void* images, void* out, int F, int S, int patch, int Kpad, void* stream) {
    return COUNT(vb::patch_im2col(images, out, F, S, patch, Kpad, ST(stream)));
}
int vidi_whisper_im2col1(const void* mel, void* out, int C, int mels, int T, void* stream) {
    return COUNT(vb::whisper_im2col1(mel, out, C, mels, T, ST(stream)));
}
int vidi_whisper_im2col2(const void* x, void* out, int C, int T, int d, void* stream) {
    return COUNT(vb::whisper_im2col2(x, out, C, T, d, ST(stream)));
}
int vidi_pool_s2d(const void* P, void* X, int F, int side, int d, int h, int w, int m, void* stream) {
    return COUNT(vb::pool_s2d(P, X, F, side, d, h, w, m, ST(stream)));
}
int vidi_conv_window_gather(const void* P, void* A, int F, int side, int d, int k, void* stream) {
    return COUNT(vb::conv_window_gather(P, A, F, side, d, k, ST(stream)));
}
int vidi_bilinear_ac(const void* X, void* Y, int F, int si, int so, int d, void* stream) {
    return COUNT(vb::bilinear_ac(X, Y, F, si, so, d, ST(stream)));
}
int vidi_embed_gather(const int64_t* ids, const void* E, void* out, int T, int D, int vocab, float normalizer, void* stream) {
    return COUNT(vb::embed_gather(ids, E, out, T, D, vocab, normalizer, ST(stream)));
}
int vidi_sinusoid_split(const float* div_term, void* out, int rows, int i0, int l, int N, int D, void* stream) {
    return COUNT(vb::sinusoid_split(div_term, out, rows, i0, l, N, D, ST(stream)));
}
int vidi_split3(const float* x, void* out, int64_t rows, int D, int mode, void* stream) {
    return COUNT(vb::split3(x, out, rows, D, mode, ST(stream)));
}
int vidi_cast_f32_bf16(const float* x, void* y, int64_t n, void* stream) {
    return COUNT(vb::cast_f32_bf16(x, y, n, ST(stream)));
}
int vidi_attn_dense(const void* qkv, int64_t ld, int q_off, int k_off, int v_off, void* out, int64_t ldo, int B, int S, int H,
                    int dh, float scale, void* stream) {
    // packed Q|K|V projection output with a tower head dim -> tcgen05/TMEM/TMA kernel; anything else -> mma.sync kernel
    if ((dh == 72 || dh == 64) && q_off == 0 && k_off == H * dh && v_off == 2 * H * dh && ld == (int64_t)3 * H * dh)
        return COUNT(S > 128 ? vb::attn_dense_sm100_v3(qkv, ld, out, ldo, B, S, H, dh, scale, ST(stream))
                             : vb::attn_dense_sm100(qkv, ld, out, ldo, B, S, H, dh, scale, ST(stream)));
    return COUNT(vb::attn_dense(qkv, ld, q_off, k_off, v_off, out, ldo, B, S, H, dh, scale, ST(stream)));
}
int vidi_attn_dense_v1(const void* qkv, int64_t ld, int q_off, int k_off, int v_off, void* out, int64_t ldo, int B, int S,
                       int H, int dh, float scale, void* stream) {
    if ((dh == 72 || dh == 64) && q_off == 0 && k_off == H * dh && v_off == 2 * H * dh && ld == (int64_t)3 * H * dh)
        return COUNT(vb::attn_dense_sm100(qkv, ld, out, ldo, B, S, H, dh, scale, ST(stream)));
    return COUNT(vb::attn_dense(qkv, ld, q_off, k_off, v_off, out, ldo, B, S, H, dh, scale, ST(stream)));
}
int vidi_attn_dense_v2(const void* qkv, int64_t ld, int q_off, int k_off, int v_off, void* out, int64_t ldo, int B, int S,
                       int H, int dh, float scale, void* stream) {
    if ((dh == 72 || dh == 64) && q_off == 0 && k_off == H * dh && v_off == 2 * H * dh && ld == (int64_t)3 * H * dh)
        return COUNT(vb::attn_dense_sm100_pp(qkv, ld, out, ldo, B, S, H, dh, scale, ST(stream)));
    return COUNT(vb::attn_dense(qkv, ld, q_off, k_off, v_off, out, ldo, B, S, H, dh, scale, ST(stream)));
}
int vidi_attn_dense_poly(const void* qkv, int64_t ld, void* out, int64_t ldo, int B, int S, int H, int dh, float scale, int poly_mod,
                         void* stream) {
    return COUNT(vb::attn_dense_sm100_v3_poly(qkv, ld, out, ldo, B, S, H, dh, scale, poly_mod, ST(stream)));
}
int vidi_attn_dense_mma(const void* qkv, int64_t ld, int q_off, int k_off, int v_off, void* out, int64_t ldo, int B, int S,
                        int H, int dh, float scale, void* stream) {
    return COUNT(vb::attn_dense(qkv, ld, q_off, k_off, v_off, out, ldo, B, S, H, dh, scale, ST(stream)));
}
int vidi_xattn_splitkv(const void* Q, int64_t ldq, const void* K, const void* V, int64_t ldkv, const uint8_t* kmask, int T,
                       int N, int Hq, int Hkv, int dh, int splits, float scale, float softcap, float* Opart, float* LSE,
                       void* stream) {
    return COUNT(vb::xattn_splitkv(Q, ldq, K, V, ldkv, kmask, T, N, Hq, Hkv, dh, splits, scale, softcap, Opart, LSE, 0, ST(stream)));
}
int vidi_xattn_splitkv_mma(const void* Q, int64_t ldq, const void* K, const void* V, int64_t ldkv, const uint8_t* kmask, int T,
                           int N, int Hq, int Hkv, int dh, int splits, float scale, float softcap, float* Opart, float* LSE,
                           void* stream) {
    return COUNT(vb::xattn_splitkv(Q, ldq, K, V, ldkv, kmask, T, N, Hq, Hkv, dh, splits, scale, softcap, Opart, LSE, 1, ST(stream)));
}
int vidi_xattn_splitkv_seg(const void* Q, int64_t ldq, const void* K, const void* V, int64_t ldkv, int n_rows_total, int nseg,
                           const int32_t* row0, const int32_t* rows, const int32_t* splits, const uint8_t* const* masks, int T, int Hq,
                           int Hkv, int dh, float scale, float softcap, float* Opart, float* LSE, void* stream) {
    int n = 1;
    const int rc = vb::xattn_splitkv_seg(Q, ldq, K, V, ldkv, n_rows_total, nseg, row0, rows, splits, masks, T, Hq, Hkv, dh, scale, softcap,
                                         Opart, LSE, 0, &n, ST(stream));
    g_launches.fetch_add(n, std::memory_order_relaxed);
    return rc;
}
int vidi_xattn_merge(const float* Opart, const float* LSE, int P, int splits_per_rank, int64_t rank_stride_o,
                     int64_t rank_stride_l, int rows, int dh, float gate, int accumulate, float* out, void* stream) {
    return COUNT(vb::xattn_merge(Opart, LSE, P, splits_per_rank, rank_stride_o, rank_stride_l, rows, dh, gate, accumulate, out,
                                 ST(stream)));
}
int vidi_text_qk_prep(const void* qkv, int64_t ld, void* q_rope, int64_t ldq, void* kv_out, int64_t ldkv, int Tq, int Hq, int Hkv,
                      int dh, const float* inv_freq, int pos0, void* stream) {
    return COUNT(vb::text_qk_prep(qkv, ld, q_rope, ldq, kv_out, ldkv, Tq, Hq, Hkv, dh, inv_freq, pos0, ST(stream)));
}
int vidi_xattn_merge2(const float* O0, const float* L0, int P0, int spr0, int64_t rso0, int64_t rsl0, float gate0, const float* O1,
                      const float* L1, int P1, int spr1, int64_t rso1, int64_t rsl1, float gate1, int nsrc, const float* att,
                      int rows, int dh, void* out_bf16, void* stream) {
    return COUNT(vb::xattn_merge2(O0, L0, P0, spr0, rso0, rsl0, gate0, O1, L1, P1, spr1, rso1, rsl1, gate1, nsrc, att, rows, dh,
                                  out_bf16, nullptr, 0, 0u, nullptr, ST(stream)));
}
int vidi_xattn_merge2_sync(const float* O0, const float* L0, int P0, int spr0, int64_t rso0, int64_t rsl0, float gate0,
                           const float* O1, const float* L1, int P1, int spr1, int64_t rso1, int64_t rsl1, float gate1, int nsrc,
                           const float* att, int rows, int dh, void* out_bf16, const uint32_t* flags, int nflags, uint32_t seq,
                           int* err, void* stream) {
    return COUNT(vb::xattn_merge2(O0, L0, P0, spr0, rso0, rsl0, gate0, O1, L1, P1, spr1, rso1, rsl1, gate1, nsrc, att, rows, dh,
                                  out_bf16, flags, nflags, seq, err, ST(stream)));
}
int vidi_xattn_premerge_push(const float* O0, const float* L0, int P0, const float* O1, const float* L1, int P1, int nsrc, int rows,
                             int dh, float* const* peer_base, uint32_t* const* peer_flag, int world, int64_t my_block_off,
                             uint32_t seq, uint32_t* counter, void* stream) {
    return COUNT(vb::xattn_premerge_push(O0, L0, P0, O1, L1, P1, nsrc, rows, dh, peer_base, peer_flag, world, my_block_off, seq,
                                         counter, ST(stream)));
}
int64_t vidi_text_pass_workspace_bytes(const VidiTextPass* d) { return vb::text_pass_workspace_bytes(d); }
int vidi_text_pass(const VidiTextPass* d, void* stream) {
    int64_t n = 0;
    const int rc = vb::text_pass(d, &n, ST(stream));
    g_launches.fetch_add(n, std::memory_order_relaxed);
    return rc;
}
int vidi_p2p_alloc(int64_t bytes, void** ptr, void* handle) { return vb::p2p_alloc(bytes, ptr, handle); }
int vidi_p2p_open(const void* handle, void** ptr) { return vb::p2p_open(handle, ptr); }
int vidi_p2p_close(void* ptr) { return vb::p2p_close(ptr); }
int vidi_p2p_free(void* ptr) { return vb::p2p_free(ptr); }
int vidi_rope_inplace(void* x, int64_t ld, int col_off, int T, int heads, int dh, const float* inv_freq, int pos0, void* stream) {
    return COUNT(vb::rope_inplace(x, ld, col_off, T, heads, dh, inv_freq, pos0, ST(stream)));
}
int vidi_attn_text(const void* Q, int64_t ldq, const void* K, const void* V, int64_t ldkv, int Tq, int Tk, int pos0, int Hq,
                   int Hkv, int dh, float scale, float softcap, int window, float* out, void* stream) {
    return COUNT(vb::attn_text(Q, ldq, K, V, ldkv, Tq, Tk, pos0, Hq, Hkv, dh, scale, softcap, window, out, ST(stream)));
}

}  // extern "C"
