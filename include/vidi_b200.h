/* vidi_b200 -- C ABI of the B200 (sm_100a) kernel library behind the Vidi prefill path.
 *
 * The reference (bytedance/vidi) has no native code and no FFI: every GPU instruction on its hot path
 * comes from PyTorch/cuBLAS/cuDNN, flash-attn 2 and liger-kernel calls made from Python.  Each entry
 * point below therefore cites the *reference call site(s)* whose third-party kernels it replaces
 * (paths relative to the reference repo root; K-numbers refer to SURVEY.md section 2.2).
 *
 * Conventions: plain pointers and sizes only; every pointer is a DEVICE pointer unless stated; bf16
 * tensors are row-major with an explicit leading dimension in ELEMENTS; `stream` is a cudaStream_t
 * passed as void*; functions return 0 on success, non-zero on error (vidi_last_error() has the text);
 * no function synchronises the stream.  No CPU fallback exists anywhere in this library.
 */
#ifndef VIDI_B200_H
#define VIDI_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* vidi_last_error(void);
int vidi_abi_version(void);
/* number of kernel launches issued through this library since the last reset (bench "gpu_launches") */
int64_t vidi_launch_count(void);
void vidi_reset_launch_count(void);

/* activation / GLU codes for vidi_gemm_bf16 */
#define VIDI_ACT_NONE 0
#define VIDI_ACT_GELU_ERF 1   /* nn.GELU(): projector, pos-MLP, Whisper      (mm_layer/mlp.py:20, mm_vision/pos.py:38) */
#define VIDI_ACT_GELU_TANH 2  /* gelu_pytorch_tanh: SigLIP MLP                (HF modeling_siglip.py SiglipMLP)       */
#define VIDI_ACT_SOFTCAP 3    /* cap * tanh(x / cap): final logit soft-cap    (lmm/dattn/gemma.py:566-569)            */
#define VIDI_ACT_SILU 4
#define VIDI_GLU_NONE 0
#define VIDI_GLU_GELU_TANH 1  /* Gemma2MLP  down(gelu_tanh(gate) * up)        (gemma.py:116-123 -> HF Gemma2MLP)      */
#define VIDI_GLU_SILU 2       /* MistralMLP down(silu(gate) * up)             (Vidi_7B mistral.py:131-137)            */

/* C[M,N] = epi(A[M,K] . W[N,K]^T): tcgen05/TMEM/TMA persistent GEMM.  Replaces every nn.Linear / conv-as-GEMM on the
 * path: k_proj/v_proj (gemma.py:61-62), o_proj (gemma.py:94,197), Gemma2MLP (gemma.py:116-123), lm_head (gemma.py:565),
 * SigLIP/Whisper linears (mm_vision/siglip.py:30, mm_audio/whisper.py:27), projector MLP (mm_layer/mlp.py:18-22),
 * patch conv and Conv1d pools (multimodal.py:85-88,232).
 *  bias: fp32 [N] or NULL.  residual: bf16 [M or res_mod, ldr] or NULL, added after the activation; if res_mod > 0 the
 *  residual row is (row % res_mod) (position-embedding add).  glu != 0: W rows are packed per block_n tile as
 *  [block_n/2 gate rows | block_n/2 up rows] and C has N/2 columns.  out_fp32: C is float.  block_n in {64,128,256}. */
int vidi_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K,
                   const float* bias, const void* residual, int64_t ldr, int res_mod, int act, float act_param,
                   int out_fp32, int glu, int block_n, void* stream);

/* same contract on a CTA pair (tcgen05 cta_group::2, 256 x block_n tile per pair, block_n in {128,256}) */
int vidi_gemm_bf16_2cta(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K,
                        const float* bias, const void* residual, int64_t ldr, int res_mod, int act, float act_param,
                        int out_fp32, int glu, int block_n, void* stream);

/* CTA-pair GEMM with a LayerNorm folded in and/or row statistics out — the pre-LN blocks of the towers (HF SiglipEncoderLayer:
 * layer_norm1 -> self_attn, layer_norm2 -> mlp; WhisperEncoderLayer likewise) without a LayerNorm pass over HBM:
 *   ln_stats != NULL: A is the RAW residual stream x [M,K]; W must be pre-multiplied by gamma (W' = W diag(gamma)), ln_colsum[n] =
 *     sum_k W'[n,k], bias must already include W beta; the epilogue computes rstd*(acc - mean*ln_colsum) + bias per row, with the
 *     row's (sum, sumsq) given as ln_parts partial pairs: ln_stats is float [M, ln_parts, 2];
 *   stats_out != NULL: float [M, 2*ceil(N/block_n), 2] receives the per-row (sum, sumsq) of the bf16 values stored to C, one pair
 *     per (column tile, column half) — the ln_stats input of the GEMM that consumes C as its LayerNorm-ed operand.
 * bf16 output, no GLU; other arguments as vidi_gemm_bf16. */
int vidi_gemm_bf16_2cta_ln(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K,
                           const float* bias, const void* residual, int64_t ldr, int res_mod, int act, float act_param,
                           int block_n, const float* ln_stats, int ln_parts, const float* ln_colsum, float ln_eps,
                           float* stats_out, void* stream);

/* y = x_hat(x,eps) * (add_one ? 1+w : w) * out_scale.  Gemma2RMSNorm (gemma.py:107-111,162,184) / vidi RMSNorm (mm_layer/norm.py:17-25) */
int vidi_rmsnorm(const void* x, int64_t ldx, const void* w, void* y, int64_t ldy, int rows, int D, float eps, int add_one,
                 float out_scale, void* stream);
/* x += post_mode ? G(y,w_post) : y ; h = norm(x, w_next) (h may be NULL).  gemma.py:198-202 and 116-123 fused */
int vidi_residual_norm(void* x, int64_t ldx, const void* y, int64_t ldy, const void* w_post, const void* w_next, void* h,
                       int64_t ldh, int rows, int D, float eps, int post_mode, int next_add_one, void* stream);
/* nn.LayerNorm with fp32 affine (SigLIP / Whisper encoder layers) */
int vidi_layernorm(const void* x, int64_t ldx, const float* w, const float* b, void* y, int64_t ldy, int rows, int D,
                   float eps, void* stream);
/* RMSNorm + up to 3 positional-table adds + token mask + llm norm + sqrt(D) scale in one pass
 * (multimodal.py:193-206,241-250; gemma.py:353-356).  tabs: HOST array of ntab device pointers (fp32 [*,D]). */
int vidi_mm_finish(const void* proj, int64_t ldp, const void* w_mod, const void* w_llm, const float* const* tabs,
                   const int* divs, const int* mods, const int* offs, int ntab, int n_offset, int sample_valid,
                   float normalizer, void* out, int64_t ldo, uint8_t* mask, int rows, int D, float eps, void* stream);
/* weight-less rms_norm of fp32 rows (rms_norm(pos_mlp(.)), mm_layer/norm.py:9-16) */
int vidi_rmsnorm_f32(const float* x, float* y, int rows, int D, float eps, int round_bf16, void* stream);

/* layout kernels */
int vidi_patch_im2col(const void* images, void* out, int F, int S, int patch, int Kpad, void* stream);      /* K1 */
int vidi_whisper_im2col1(const void* mel, void* out, int C, int mels, int T, void* stream);                /* K8 conv1 */
int vidi_whisper_im2col2(const void* x, void* out, int C, int T, int d, void* stream);                     /* K8 conv2 */
/* Conv2DPool: pad + bilinear + space_to_depth gather (mm_vision/pool.py:23-32, utils.py:134-150), K4 */
int vidi_pool_s2d(const void* P, void* X, int F, int side, int d, int h, int w, int m, void* stream);
/* Vidi-7B learned pool (Vidi_7B/model/mm_vision/pool.py:6-26): k x k stride-1 window gather feeding the conv GEMM, and the
 * bilinear(align_corners=True) resize of the token-major map [F,si,si,d] -> [F,so,so,d] */
int vidi_conv_window_gather(const void* P, void* A, int F, int side, int d, int k, void* stream);
int vidi_bilinear_ac(const void* X, void* Y, int F, int si, int so, int d, void* stream);
int vidi_embed_gather(const int64_t* ids, const void* E, void* out, int T, int D, int vocab, float normalizer, void* stream);
int vidi_sinusoid_split(const float* div_term, void* out, int rows, int i0, int l, int N, int D, void* stream); /* pos.py:18-26,48-56 */
int vidi_split3(const float* x, void* out, int64_t rows, int D, int mode, void* stream);
int vidi_cast_f32_bf16(const float* x, void* y, int64_t n, void* stream);

/* frame pre-processing: the two passes of PIL.Image.resize(..., BICUBIC) as process_images' 'resize' branch applies it
 * (vidi/dataset/img_utils.py:181-187; Pillow libImaging/Resample.c 8-bit path), bit-exact integer arithmetic.  Tap tables from the host
 * (precompute_coeffs + normalize_coeffs_8bpc): xmin int32 [out_size], kk int32 [out_size, ksize] (22-bit fixed point).
 *   vidi_resample_u8:             src uint8 [outer, in_size, inner] -> dst uint8 [outer, out_size, inner]   (horizontal pass: inner = 3)
 *   vidi_resample_u8_to_chw_bf16: src uint8 [F, in_h, W, 3] -> dst bf16 [F, 3, out_h, W] = ((v * rescale) - mean) / std   (vertical pass
 *                                 fused with the SiglipImageProcessor affine and the HWC -> CHW change) */
int vidi_resample_u8(const uint8_t* src, uint8_t* dst, int64_t outer, int in_size, int out_size, int inner, const int32_t* xmin,
                     const int32_t* kk, int ksize, void* stream);
int vidi_resample_u8_to_chw_bf16(const uint8_t* src, void* dst, int F, int in_h, int out_h, int W, const int32_t* ymin,
                                 const int32_t* kk, int ksize, float rescale, float mean, float stdv, void* stream);

/* audio pre-processing: operand builders and finisher of the Whisper log-mel (process_audio, vidi/dataset/vid_utils.py:52-63 -> HF
 * WhisperFeatureExtractor: hann 400 / hop 160 centred reflect-padded STFT, power, 128 Slaney mel filters, log10 clamp 1e-10, per-chunk
 * max - 8 floor, (x + 4) / 4, last frame dropped).  The DFT and the mel projection run as vidi_gemm_bf16 in 3-term split-bf16 form:
 *   vidi_logmel_frames: audio fp32 [C, 480000], window fp32 [400]  -> A1 bf16 [C*3001, 1200] = [hi | hi | lo] of frame * window
 *   vidi_logmel_power:  Y fp32 [rows, ldy] = [Re 0..200 | Im 0..200 | ...] -> A2 bf16 [rows, 624] = [hi | hi | lo] of |X_k|^2 (padded to 208)
 *   vidi_logmel_finish: M fp32 [C*3001, mels], chunk_max fp32 [C] (scratch) -> out bf16 [C, mels, 3000] */
int vidi_logmel_frames(const float* audio, const float* window, void* out, int C, int n_samples, void* stream);
int vidi_logmel_power(const float* Y, int64_t ldy, void* out, int64_t rows, void* stream);
int vidi_logmel_finish(const float* M, int C, int mels, float* chunk_max, void* out, void* stream);

/* attention */
/* bidirectional flash attention for the towers (flash_attn_func via HF SiglipAttention / WhisperAttention), K3/K8 */
int vidi_attn_dense(const void* qkv, int64_t ld, int q_off, int k_off, int v_off, void* out, int64_t ldo, int B, int S, int H,
                    int dh, float scale, void* stream);
/* same contract, first-generation tcgen05 kernel (one query block per item, split-key softmax warpgroups); A/B bar */
int vidi_attn_dense_v1(const void* qkv, int64_t ld, int q_off, int k_off, int v_off, void* out, int64_t ldo, int B, int S,
                       int H, int dh, float scale, void* stream);
/* same contract, second-generation kernel (ping-pong query blocks, P through shared memory); A/B bar for the current one
 * (P and O resident in tensor memory) */
int vidi_attn_dense_v2(const void* qkv, int64_t ld, int q_off, int k_off, int v_off, void* out, int64_t ldo, int B, int S,
                       int H, int dh, float scale, void* stream);
/* A/B entry, not on the product path and not yet validated on a GPU: the current kernel with every poly_mod-th (2, 3 or 4) pair of scores
 * exponentiated by an FMA-pipe polynomial instead of MUFU.EX2 (packed qkv [B*S, 3*H*dh], dh 64 / 72, S > 128) */
int vidi_attn_dense_poly(const void* qkv, int64_t ld, void* out, int64_t ldo, int B, int S, int H, int dh, float scale, int poly_mod,
                         void* stream);
/* same contract, always the warp-level mma.sync kernel (generic strides / head dims; kept as the A/B bar for the tcgen05 path) */
int vidi_attn_dense_mma(const void* qkv, int64_t ld, int q_off, int k_off, int v_off, void* out, int64_t ldo, int B, int S,
                        int H, int dh, float scale, void* stream);
/* split-KV cross attention, replaces flash_cross_attention_forward (lmm/dattn/xattn.py:141-263) as called from
 * DattnGemma2Attention.forward_xattn (gemma.py:81-91).  Opart fp32 [splits,T,Hq,dh], LSE fp32 [splits,T,Hq]. */
int vidi_xattn_splitkv(const void* Q, int64_t ldq, const void* K, const void* V, int64_t ldkv, const uint8_t* kmask, int T,
                       int N, int Hq, int Hkv, int dh, int splits, float scale, float softcap, float* Opart, float* LSE,
                       void* stream);
/* Both key segments of a layer (image rows and audio rows of the same K||V cache: the T2V and T2A calls of gemma.py:185-192 and
 * 206-221) in one call; on the tcgen05 path ONE launch whose grid covers the key splits of both segments.  K / V point at cache row 0
 * of the layer (n_rows_total rows); segment i = rows [row0[i], row0[i]+rows[i]) in splits[i] key ranges with mask masks[i] (or NULL).
 * Opart fp32 [splits[0]+splits[1]][T][Hq][dh], LSE fp32 [splits[0]+splits[1]][T][Hq] -- segment 1's partials follow segment 0's. */
int vidi_xattn_splitkv_seg(const void* Q, int64_t ldq, const void* K, const void* V, int64_t ldkv, int n_rows_total, int nseg,
                           const int32_t* row0, const int32_t* rows, const int32_t* splits, const uint8_t* const* masks, int T, int Hq,
                           int Hkv, int dh, float scale, float softcap, float* Opart, float* LSE, void* stream);
/* same contract, always the warp-level mma.sync kernel (any head dim in {128,256}, soft-cap optional: the Vidi-7B path) */
int vidi_xattn_splitkv_mma(const void* Q, int64_t ldq, const void* K, const void* V, int64_t ldkv, const uint8_t* kmask, int T,
                           int N, int Hq, int Hkv, int dh, int splits, float scale, float softcap, float* Opart, float* LSE,
                           void* stream);
/* LSE merge of P partials into fp32 out [rows, dh] (replaces the reference's all-gather of full K/V: all_to_all.py:361).
 * partial p = (rank p / splits_per_rank, split p % splits_per_rank) is at Opart + rank*rank_stride_o + split*rows*dh and
 * LSE + rank*rank_stride_l + split*rows (strides in floats): the all-gathered per-rank [O | LSE] blocks merge in place. */
int vidi_xattn_merge(const float* Opart, const float* LSE, int P, int splits_per_rank, int64_t rank_stride_o,
                     int64_t rank_stride_l, int rows, int dh, float gate, int accumulate, float* out, void* stream);
/* fused text-stream helpers (fewer launches per layer / per decoded token):
 *  vidi_text_qk_prep: q_rope = RoPE(q); kv_out rows = RoPE(k) | v, read from the fused qkv projection (HF apply_rotary_pos_emb,
 *                     gemma.py:165-175);
 *  vidi_xattn_merge2: out_bf16 = bf16(att_text + sum_s gate_s * LSE-merge(partials_s)) for up to two streams (gemma.py:236). */
int vidi_text_qk_prep(const void* qkv, int64_t ld, void* q_rope, int64_t ldq, void* kv_out, int64_t ldkv, int Tq, int Hq, int Hkv,
                      int dh, const float* inv_freq, int pos0, void* stream);
int vidi_xattn_merge2(const float* O0, const float* L0, int P0, int spr0, int64_t rso0, int64_t rsl0, float gate0, const float* O1,
                      const float* L1, int P1, int spr1, int64_t rso1, int64_t rsl1, float gate1, int nsrc, const float* att,
                      int rows, int dh, void* out_bf16, void* stream);
/* Multi-GPU exchange of the text stream's cross-attention partials over NVLink peer memory; replaces the reference's
 * sequence-parallel Gather.forward / all-gather (lmm/dattn/sequence_parallel/all_to_all.py:361, split.py:72-93).
 *  vidi_xattn_premerge_push: LSE-merge this rank's P0 (P1) key splits of stream 0 (1) into one (O [rows,dh], LSE [rows]) partial per
 *      stream and store it at float offset my_block_off of every peer_base[r] (r < world; peer-mapped device pointers, layout per
 *      stream: O | LSE); when peer_flag is non-NULL the last block publishes `seq` in *peer_flag[r] of every rank (release, system
 *      scope).  counter: one zero-initialised uint32 of scratch on this device.
 *  vidi_xattn_merge2_sync: vidi_xattn_merge2 that first waits until flags[0..nflags) (this rank's own flag words, one per source
 *      rank) have all reached `seq`; *err is set to 1 if a peer never arrives (bounded spin, ~3 s).
 *  vidi_p2p_*: the exchange arena itself -- cudaMalloc + CUDA IPC export / import (one process per GPU, same node). */
int vidi_xattn_premerge_push(const float* O0, const float* L0, int P0, const float* O1, const float* L1, int P1, int nsrc, int rows,
                             int dh, float* const* peer_base, uint32_t* const* peer_flag, int world, int64_t my_block_off,
                             uint32_t seq, uint32_t* counter, void* stream);
int vidi_xattn_merge2_sync(const float* O0, const float* L0, int P0, int spr0, int64_t rso0, int64_t rsl0, float gate0,
                           const float* O1, const float* L1, int P1, int spr1, int64_t rso1, int64_t rsl1, float gate1, int nsrc,
                           const float* att, int rows, int dh, void* out_bf16, const uint32_t* flags, int nflags, uint32_t seq,
                           int* err, void* stream);
int vidi_p2p_alloc(int64_t bytes, void** ptr, void* ipc_handle_64_bytes);
int vidi_p2p_open(const void* ipc_handle_64_bytes, void** ptr);
int vidi_p2p_close(void* ptr);
int vidi_p2p_free(void* ptr);
int vidi_rope_inplace(void* x, int64_t ld, int col_off, int T, int heads, int dh, const float* inv_freq, int pos0, void* stream);
/* causal text self attention with soft-cap and sliding window (HF Gemma2Attention via gemma.py:165-175), K16 */
int vidi_attn_text(const void* Q, int64_t ldq, const void* K, const void* V, int64_t ldkv, int Tq, int Tk, int pos0, int Hq,
                   int Hkv, int dh, float scale, float softcap, int window, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------------------
 * The whole text stream of one prefill / decode step in ONE call: every decoder layer's T2T self attention, T2V / T2A split-KV
 * cross attention against the image / audio K||V cache, the (multi-GPU) exchange of the partials, o_proj, the norms and the GLU
 * MLP, then the final norm and lm_head -- DattnGemma2Model.forward's layer loop for the text rows (gemma.py:362-409 with
 * :160-175, :185-192, :206-221, :236-238, :411, :564-569; Mistral family: Vidi_7B mistral.py:190-264, 615-616).  The reference
 * issues these as ~40 eager PyTorch ops per layer; here a native loop enqueues 12 kernels per layer on `stream` with no Python in
 * between, which is what makes decode (Tq = 1) and the replicated text pass of the multi-GPU prefill launch-latency free.
 * All pointers are device pointers except layer_w (HOST array).  Nothing is synchronised. */
typedef struct VidiTextLayerW {
    const void* wqkv;     /* bf16 [heads*head_dim + 2*kv_heads*head_dim, hidden]   q | k | v rows                           */
    const void* wo;       /* bf16 [hidden, heads*head_dim]                                                                  */
    const void* wgu;      /* bf16 [2*inter, hidden], gate/up rows interleaved per 256-row tile (weights.pack_glu)           */
    const void* wd;       /* bf16 [hidden, inter]                                                                           */
    const void* n_in;     /* bf16 [hidden] input_layernorm                                                                  */
    const void* n_post;   /* bf16 [hidden] post_attention_layernorm                                                         */
    const void* n_preff;  /* bf16 [hidden] pre_feedforward_layernorm  (Gemma2 family; NULL for Mistral)                     */
    const void* n_postff; /* bf16 [hidden] post_feedforward_layernorm (Gemma2 family; NULL for Mistral)                     */
} VidiTextLayerW;
typedef struct VidiTextSeg {
    int64_t row0;          /* first row of this stream (image or audio) in the K||V cache of every layer                    */
    int32_t rows;          /* keys of this stream held by this rank (may be 0)                                              */
    int32_t splits;        /* key splits (>= 1)                                                                             */
    const uint8_t* kmask;  /* uint8 [rows] key-padding mask or NULL                                                         */
    float gate;            /* `any(mask)` gate of gemma.py:192 as 1.0 / 0.0                                                 */
    int32_t reserved;
} VidiTextSeg;
typedef struct VidiTextPass {
    int32_t Tq, pos0, layers, hidden, heads, kv_heads, head_dim, inter, vocab;
    int32_t gemma;           /* 1: Gemma2 family (post norms, soft-caps, (1+w) norms, alternating window); 0: Mistral family */
    int32_t glu;             /* VIDI_GLU_GELU_TANH / VIDI_GLU_SILU                                                          */
    int32_t sliding_window;  /* Gemma2: applied on even layers; Mistral: on every layer; 0 = none                           */
    int32_t logits_keep;     /* 0: logits for all Tq rows, k: for the last k rows                                           */
    float rms_eps, scale, attn_softcap, final_softcap, normalizer;
    const VidiTextLayerW* layer_w;   /* HOST array [layers] */
    const void* embed;       /* bf16 [vocab, hidden] */
    const void* final_norm;  /* bf16 [hidden] */
    const void* lm_head;     /* bf16 [vocab, hidden] */
    const float* inv_freq;   /* fp32 [head_dim/2] */
    const int64_t* ids;      /* int64 [Tq], sentinel already stripped */
    void* text_kv;           /* bf16 [layers][max_len][2*kv_dim] text K||V cache (rows pos0..pos0+Tq are written), or NULL when
                                pos0 == 0 and nothing is kept (scratch from the workspace is used)                          */
    int64_t text_kv_layer_stride, text_kv_ld;    /* in elements */
    const void* stream_kv;   /* bf16 [layers][N][2*kv_dim] image+audio K||V cache of this rank */
    int64_t stream_layer_stride, stream_ld;      /* in elements */
    int32_t nseg;
    int32_t stream_rows;     /* rows N of the stream K||V cache (per layer) */
    int32_t world, rank;     /* world > 1: partials cross ranks through the peer arenas below (see vidi_xattn_premerge_push) */
    uint32_t seq0;           /* sequence number of the last exchange issued before this call; layer l uses seq0 + l + 1     */
    int32_t reserved0;
    VidiTextSeg seg[2];
    float* peer_data[16];    /* arena data base of every rank (peer-mapped): fp32 [2 slots][world][cap]                      */
    uint32_t* peer_flags[16];/* arena flag base of every rank: uint32 [2 slots][world]                                      */
    int64_t cap;             /* floats per rank block */
    uint32_t* counter;       /* one zero-initialised uint32 on this device */
    int32_t* err;            /* set to 1 if a peer never delivered */
    void* workspace;         /* >= vidi_text_pass_workspace_bytes() bytes, 256-byte aligned */
    int64_t workspace_bytes;
    float* logits;           /* fp32 [logits_keep ? logits_keep : Tq][vocab] */
} VidiTextPass;
int64_t vidi_text_pass_workspace_bytes(const VidiTextPass* d);
int vidi_text_pass(const VidiTextPass* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif
