// Weight-streaming GEMM for the text stream (M <= 64 rows: the ~32 prompt tokens of a prefill, 1 token of a decode step):
//     C[M, N] = epilogue(A[M, K] . W[N, K]^T)
// Every text GEMM is bound by reading W once from HBM (18.5 GB per pass over the 42 layers of Vidi1.5-9B + lm_head).  The
// general kernel (gemm_sm100.cu) puts the M rows on the 128-wide MMA M axis: each 64-wide k-block then moves 16 KB of (mostly
// zero) A rows next to 8 KB of weights and the L2 -> SM path, not HBM, is the limit (measured 2.5 TB/s).  Here the operands are
// SWAPPED:
//   * W rows are the MMA M axis: a CTA streams [128 rows x K] of W through a deep TMA ring (16 KB of weights per stage),
//     the text rows are the MMA N axis (TN = 16 / 32 / 64 columns, 2-8 KB per stage): D[128 w-rows, TN tokens] in TMEM;
//   * GLU sites (gate || up packed per 256 rows, weights.pack_glu): two MMAs per k-step into two accumulators, gelu_tanh / silu
//     product formed by the thread that holds both (lane = W row);
//   * N = 3584 sites (o_proj, down_proj) have only 28 row tiles: the K range is SPLIT over `ksplit` CTAs so that every SM streams;
//     partials go to an fp32 workspace and the LAST CTA to arrive for a tile (one atomic counter per tile) adds them in split
//     order -- a fixed order, so results are bit-reproducible run to run and rank to rank -- applies the epilogue and stores;
//   * persistent CTAs, double-buffered accumulators (epilogue of item i overlaps the stream of item i+1).
// Replaces, for the text rows, q/k/v/o_proj, Gemma2MLP / MistralMLP and lm_head (+ soft-cap): gemma.py:61-62,94,116-123,564-569.
#include "common.cuh"

namespace vb {
namespace sk {

enum Act : int { ACT_NONE = 0, ACT_SOFTCAP = 3 };
enum Glu : int { GLU_NONE = 0, GLU_GELU_TANH = 1, GLU_SILU = 2 };

constexpr int BW = 128;             // W rows per MMA (UMMA M)
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;       // warp0 TMA, warp1 MMA + TMEM alloc, warps 2-5 epilogue (128 lanes = 128 W rows)

struct Params {
    int M, N, K;                    // N = W rows (2 * out columns for GLU)
    void* C;
    int64_t ldc;
    int act;
    float act_param;
    int out_fp32;
    int glu;
    int ksplit, kb_per_split, num_k;
    int n_tiles;
    float* ws;                      // fp32 [ksplit][n_out][TN] partials (ksplit > 1)
    unsigned int* counters;         // [n_tiles], zero between launches
};

template <int TN, bool GLU>
struct Cfg {
    static constexpr int kWBytes = (GLU ? 2 : 1) * BW * BLOCK_K * 2;     // 16 / 32 KB of weights per stage
    static constexpr int kABytes = TN * BLOCK_K * 2;
    static constexpr int kStageBytes = kWBytes + kABytes;
    static constexpr int kStages = (200 * 1024) / kStageBytes > 12 ? 12 : (200 * 1024) / kStageBytes;
    static constexpr int kAccCols = (GLU ? 2 : 1) * TN;                  // per accumulator buffer
    static constexpr int kTmemCols = 2 * kAccCols <= 32 ? 32 : 2 * kAccCols <= 64 ? 64 : 2 * kAccCols <= 128 ? 128 : 256;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 512;
};

template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[N]);
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld_32x32b_x16(taddr, r); }
template <>
__device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_32x32b_x32(taddr, r); }
template <>
__device__ __forceinline__ void tmem_ld_cols<64>(uint32_t taddr, uint32_t (&r)[64]) {
    tmem_ld_32x32b_x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
    tmem_ld_32x32b_x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
}

template <int TN, bool GLU>
__global__ void __launch_bounds__(kThreads, 1)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a, const Params p) {
    using C = Cfg<TN, GLU>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_w = smem;
    uint8_t* smem_a = smem + C::kStages * C::kWBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + C::kStages;
    uint64_t* tmem_full = bars + 2 * C::kStages;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    __shared__ int s_last;

    const int warp = threadIdx.x >> 5;
    const int items = p.n_tiles * p.ksplit;        // item = (row tile, k split); splits of a tile are adjacent items
    constexpr int ROWS = GLU ? 2 * BW : BW;        // W rows per tile

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_a);
    }
    if (warp == 1) {
        if (elect_one()) {
            for (int i = 0; i < C::kStages; ++i) {
                mbar_init(&full_bar[i], 1);
                mbar_init(&empty_bar[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&tmem_full[i], 1);
                mbar_init(&tmem_empty[i], 128);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<C::kTmemCols>(tmem_ptr);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int it = blockIdx.x; it < items; it += gridDim.x) {
                const int tile = it / p.ksplit, ks = it - tile * p.ksplit;
                const int kb0 = ks * p.kb_per_split, kb1 = min(p.num_k, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], C::kStageBytes);
                    uint8_t* sw = smem_w + stage * C::kWBytes;
                    tma_load_2d(sw, &tmap_w, &full_bar[stage], kb * BLOCK_K, tile * ROWS, kEvictFirst);
                    if (GLU) tma_load_2d(sw + BW * BLOCK_K * 2, &tmap_w, &full_bar[stage], kb * BLOCK_K, tile * ROWS + BW, kEvictFirst);
                    tma_load_2d(smem_a + stage * C::kABytes, &tmap_a, &full_bar[stage], kb * BLOCK_K, 0, kEvictLast);
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(BW, TN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int it = blockIdx.x; it < items; it += gridDim.x) {
                const int tile = it / p.ksplit, ks = it - tile * p.ksplit;
                const int kb0 = ks * p.kb_per_split, kb1 = min(p.num_k, kb0 + p.kb_per_split);
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * C::kAccCols;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t w_desc = umma_desc_k_sw128(smem_u32(smem_w + stage * C::kWBytes));
                    const uint64_t a_desc = umma_desc_k_sw128(smem_u32(smem_a + stage * C::kABytes));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        umma_f16(d, w_desc + 2 * k, a_desc + 2 * k, idesc, (kb > kb0) || k != 0);
                        if (GLU) {
                            const uint64_t u_desc = umma_desc_k_sw128(smem_u32(smem_w + stage * C::kWBytes + BW * BLOCK_K * 2));
                            umma_f16(d + TN, u_desc + 2 * k, a_desc + 2 * k, idesc, (kb > kb0) || k != 0);
                        }
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // epilogue: warp w owns TMEM lanes (w % 4) * 32 .. +32; thread = one W row of the tile, TN token columns
        const int q = warp & 3;
        const int lrow = q * 32 + lane_id();
        const int n_out_total = GLU ? p.N / 2 : p.N;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int it = blockIdx.x; it < items; it += gridDim.x) {
            const int tile = it / p.ksplit, ks = it - tile * p.ksplit;
            mbar_wait_warp(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * C::kAccCols + ((uint32_t)(q * 32) << 16);
            float v[TN];
            {
                uint32_t r[TN];
                tmem_ld_cols<TN>(taddr, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < TN; ++j) v[j] = __uint_as_float(r[j]);
            }
            float u[GLU ? TN : 1];
            if (GLU) {
                uint32_t r[TN];
                tmem_ld_cols<TN>(taddr + TN, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < TN; ++j) u[j] = __uint_as_float(r[j]);
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }

            const int n = tile * BW + lrow;                       // output column (GLU: tile covers 128 output columns, too)
            const bool n_ok = n < n_out_total;
            bool finish = true;
            if (p.ksplit > 1) {
                // partial of this k range -> workspace [ks][n][TN] (gate then up for GLU); the last CTA of the tile reduces
                constexpr int W = GLU ? 2 * TN : TN;
                float* wp = p.ws + ((int64_t)ks * n_out_total + n) * W;
                if (n_ok) {
#pragma unroll
                    for (int j = 0; j < TN; j += 4) *reinterpret_cast<float4*>(wp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    if (GLU) {
#pragma unroll
                        for (int j = 0; j < TN; j += 4) *reinterpret_cast<float4*>(wp + TN + j) = make_float4(u[j], u[j + 1], u[j + 2], u[j + 3]);
                    }
                }
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (threadIdx.x == 64) s_last = (atomicAdd(&p.counters[tile], 1u) == (unsigned int)p.ksplit - 1) ? 1 : 0;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                finish = s_last != 0;
                if (finish) {
                    __threadfence();
                    if (threadIdx.x == 64) p.counters[tile] = 0;          // ready for the next launch
                    if (n_ok) {
#pragma unroll
                        for (int j = 0; j < TN; ++j) v[j] = 0.f;
                        if (GLU) {
#pragma unroll
                            for (int j = 0; j < TN; ++j) u[j] = 0.f;
                        }
                        for (int s = 0; s < p.ksplit; ++s) {              // fixed order: bit-reproducible
                            const float* rp = p.ws + ((int64_t)s * n_out_total + n) * W;
#pragma unroll
                            for (int j = 0; j < TN; j += 4) {
                                const float4 t = __ldcg(reinterpret_cast<const float4*>(rp + j));
                                v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
                            }
                            if (GLU) {
#pragma unroll
                                for (int j = 0; j < TN; j += 4) {
                                    const float4 t = __ldcg(reinterpret_cast<const float4*>(rp + TN + j));
                                    u[j] += t.x; u[j + 1] += t.y; u[j + 2] += t.z; u[j + 3] += t.w;
                                }
                            }
                        }
                    }
                }
            }
            if (finish && n_ok) {
                if (GLU) {
#pragma unroll
                    for (int j = 0; j < TN; ++j) {
                        const float g = p.glu == GLU_GELU_TANH ? gelu_tanh_fast(v[j]) : v[j] / (1.0f + __expf(-v[j]));
                        v[j] = g * u[j];
                    }
                } else if (p.act == ACT_SOFTCAP) {
                    const float inv = 1.0f / p.act_param;
#pragma unroll
                    for (int j = 0; j < TN; ++j) v[j] = p.act_param * tanhf(v[j] * inv);
                }
                // out[m][n]: for a fixed token m the 32 lanes of a warp write 32 consecutive columns
                if (p.out_fp32) {
                    float* cp = reinterpret_cast<float*>(p.C) + n;
#pragma unroll
                    for (int j = 0; j < TN; ++j)
                        if (j < p.M) cp[(int64_t)j * p.ldc] = v[j];
                } else {
                    __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.C) + n;
#pragma unroll
                    for (int j = 0; j < TN; ++j)
                        if (j < p.M) cp[(int64_t)j * p.ldc] = __float2bfloat16(v[j]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<C::kTmemCols>(tmem_base);
    }
}

struct Scratch {
    float* ws = nullptr;
    unsigned int* counters = nullptr;
    size_t ws_bytes = 0, n_counters = 0;
};

// per-device scratch (workspace of the split-K partials + tile counters); grown on demand, stream-ordered use only
static int scratch(size_t ws_bytes, size_t n_counters, Scratch** out) {
    static Scratch tab[64];
    int dev = 0;
    cudaGetDevice(&dev);
    Scratch& s = tab[dev & 63];
    if (s.ws_bytes < ws_bytes) {
        if (s.ws) { VB_CUDA_CHECK(cudaDeviceSynchronize()); VB_CUDA_CHECK(cudaFree(s.ws)); }
        VB_CUDA_CHECK(cudaMalloc(&s.ws, ws_bytes));
        s.ws_bytes = ws_bytes;
    }
    if (s.n_counters < n_counters) {
        if (s.counters) { VB_CUDA_CHECK(cudaDeviceSynchronize()); VB_CUDA_CHECK(cudaFree(s.counters)); }
        VB_CUDA_CHECK(cudaMalloc(&s.counters, n_counters * sizeof(unsigned int)));
        VB_CUDA_CHECK(cudaMemset(s.counters, 0, n_counters * sizeof(unsigned int)));
        VB_CUDA_CHECK(cudaDeviceSynchronize());
        s.n_counters = n_counters;
    }
    *out = &s;
    return 0;
}

template <int TN, bool GLU>
static int launch(const void* A, int64_t lda, const void* W, int64_t ldw, Params p, cudaStream_t st) {
    using C = Cfg<TN, GLU>;
    constexpr int ROWS = GLU ? 2 * BW : BW;
    p.n_tiles = (p.N + ROWS - 1) / ROWS;
    p.num_k = (p.K + BLOCK_K - 1) / BLOCK_K;
    const int sms = num_sms();
    // split K until the item count fills the SMs (>= 8 k-blocks per split so the pipeline still streams)
    int ks = 1;
    while (p.n_tiles * (ks + 1) <= sms && p.num_k / (ks + 1) >= 8) ++ks;
    p.ksplit = ks;
    p.kb_per_split = (p.num_k + ks - 1) / ks;
    if (ks > 1) {
        Scratch* s;
        const int n_out = GLU ? p.N / 2 : p.N;
        int rc = scratch((size_t)ks * n_out * (GLU ? 2 : 1) * TN * sizeof(float), (size_t)p.n_tiles, &s);
        if (rc) return rc;
        p.ws = s->ws; p.counters = s->counters;
    }
    CUtensorMap tw, ta;
    int rc;
    if ((rc = make_tmap_2d_bf16(&tw, W, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)ldw * 2, BLOCK_K, BW))) return rc;
    if ((rc = make_tmap_2d_bf16(&ta, A, (uint64_t)p.K, (uint64_t)p.M, (uint64_t)lda * 2, BLOCK_K, TN))) return rc;
    VB_SET_SMEM_ONCE(C::kSmemBytes, gemm_skinny_kernel<TN, GLU>);
    const int items = p.n_tiles * ks;
    gemm_skinny_kernel<TN, GLU><<<items < sms ? items : sms, kThreads, C::kSmemBytes, st>>>(tw, ta, p);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace sk

bool gemm_skinny_supports(int M, int N, int K, const float* bias, const void* residual, int act, int glu) {
    (void)K;
    return M >= 1 && M <= 64 && bias == nullptr && residual == nullptr && (act == sk::ACT_NONE || act == sk::ACT_SOFTCAP) &&
           (glu == 0 || N % 256 == 0) && N >= 128;
}

int gemm_skinny(const void* A, int64_t lda, const void* W, int64_t ldw, void* Cout, int64_t ldc, int M, int N, int K, int act,
                float act_param, int out_fp32, int glu, cudaStream_t st) {
    VB_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm_skinny: K/lda/ldw must be multiples of 8");
    VB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, "gemm_skinny: alignment");
    sk::Params p{};
    p.M = M; p.N = N; p.K = K; p.C = Cout; p.ldc = ldc; p.act = act; p.act_param = act_param; p.out_fp32 = out_fp32; p.glu = glu;
    if (M <= 16) return glu ? sk::launch<16, true>(A, lda, W, ldw, p, st) : sk::launch<16, false>(A, lda, W, ldw, p, st);
    if (M <= 32) return glu ? sk::launch<32, true>(A, lda, W, ldw, p, st) : sk::launch<32, false>(A, lda, W, ldw, p, st);
    return glu ? sk::launch<64, true>(A, lda, W, ldw, p, st) : sk::launch<64, false>(A, lda, W, ldw, p, st);
}

}  // namespace vb
