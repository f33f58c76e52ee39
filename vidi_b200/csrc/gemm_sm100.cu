// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] . W[N,K]^T)
//
//   * A and W are both K-major (nn.Linear layout: y = x W^T), moved global->shared by TMA
//     (cp.async.bulk.tensor, 128-byte swizzle), 4-stage mbarrier ring.
//   * tcgen05.mma (cta_group::1, kind::f16, M=128, N=BLOCK_N, K=16) issued by one elected thread,
//     fp32 accumulators live in TMEM, double-buffered (2 x BLOCK_N columns) so the epilogue of tile i
//     overlaps the main loop of tile i+1.
//   * Epilogue warps read TMEM with tcgen05.ld (32 lanes x 32 columns), apply bias / activation /
//     GeGLU / residual / soft-cap in fp32, and store bf16 (or fp32) rows with 128-bit stores.
//   * One CTA per SM (grid = min(#tiles, #SMs)), static round-robin over a grouped tile order that
//     keeps G m-blocks of A resident in L2 while W streams.
//
// This one kernel family serves every dense contraction on the Vidi prefill path (SURVEY.md 2.2
// K1,K2,K5,K8,K9,K12,K13,K14,K17,K18).
#include <cstdlib>

#include "common.cuh"

namespace vb {

enum Act : int { ACT_NONE = 0, ACT_GELU_ERF = 1, ACT_GELU_TANH = 2, ACT_SOFTCAP = 3, ACT_SILU = 4 };
enum Glu : int { GLU_NONE = 0, GLU_GELU_TANH = 1, GLU_SILU = 2 };

struct GemmParams {
    int M, N, K;
    void* C;            // bf16 or fp32, row stride ldc (elements)
    int64_t ldc;
    const float* bias;  // [N] fp32 or null
    const __nv_bfloat16* residual;  // [M, ldr] or null; added after activation
    int64_t ldr;
    int res_mod;        // > 0: residual row = row % res_mod (position-embedding add)
    int act;
    float act_param;    // soft-cap value
    int out_fp32;
    int glu;            // GLU_*: W rows are packed per tile as [BLOCK_N/2 gate | BLOCK_N/2 up]; C has N/2 columns
    int group_m;
};

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 384;  // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 spare, warps4-11 epilogue (2 column halves x 4 lane quarters)

template <int BLOCK_N, bool TS = false>
struct GemmCfg {
    static constexpr int kStages = (BLOCK_N == 256) ? 4 : (BLOCK_N == 192) ? 5 : 6;
    static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
    static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kOutBytes = TS ? 8 * 32 * 128 : 0;          // TMA-store epilogue: one [32 rows x 64 cols] bf16 box per epilogue warp
    static constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                     : (2 * BLOCK_N <= 256) ? 256 : 512;
    static constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int group_m, int& m_blk, int& n_blk) {
    const int per_group = group_m * num_n;
    const int g = tile / per_group;
    const int first_m = g * group_m;
    const int gm = min(group_m, num_m - first_m);
    const int r = tile - g * per_group;
    m_blk = first_m + r % gm;
    n_blk = r / gm;
}

// TS: bf16 outputs (plain and GLU) leave through shared memory and TMA tensor stores.  Each epilogue warp stages a [32 rows x 64 columns]
// box (128-byte swizzle, conflict-free st.shared.v4) and one lane issues cp.async.bulk.tensor (UTMASTG): full 128-byte lines reach L2
// instead of 32 half-sector writes per warp instruction, and rows >= M / columns >= N are clipped by the tensor map.
template <int BLOCK_N, bool TS = false>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_c, const GemmParams p) {
    using Cfg = GemmCfg<BLOCK_N, TS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
    uint8_t* smem_out = smem + Cfg::kStages * Cfg::kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kOutBytes);
    uint64_t* full_bar = bars;                       // [kStages]   TMA -> MMA
    uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]   MMA -> TMA
    uint64_t* tmem_full = bars + 2 * Cfg::kStages;   // [2]         MMA -> epilogue
    uint64_t* tmem_empty = tmem_full + 2;            // [2]         epilogue -> MMA
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int num_m = (p.M + BLOCK_M - 1) / BLOCK_M;
    const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
    const int num_tiles = num_m * num_n;
    const int num_k = (p.K + BLOCK_K - 1) / BLOCK_K;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < Cfg::kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 256);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int m_blk, n_blk;
                tile_coords(tile, num_m, num_n, p.group_m, m_blk, n_blk);
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                    tma_load_2d(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kb * BLOCK_K,
                                m_blk * BLOCK_M, kEvictNormal);
                    tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * BLOCK_K,
                                n_blk * BLOCK_N, kEvictNormal);
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t a_desc = umma_desc_k_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
                    const uint64_t b_desc = umma_desc_k_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // advance 16 elements = 32 bytes inside the 128-byte swizzle atom: +2 in the (addr>>4) field
                        umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = (warp - 4) & 3;                // == warp % 4: TMEM lane quarter this warp may access
        const int wg = (warp - 4) >> 2;               // which half of the tile columns this warp drains
        const int row_in_tile = ew * 32 + lane_id();
        int acc = 0;
        uint32_t acc_phase = 0;
        const int out_cols_total = p.glu ? p.N / 2 : p.N;
        uint8_t* stage_box = smem_out + (warp - 4) * 4096;              // TS: this warp's staging box
        uint8_t* srow = stage_box + lane_id() * 128;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int m_blk, n_blk;
            tile_coords(tile, num_m, num_n, p.group_m, m_blk, n_blk);
            mbar_wait_warp(&tmem_full[acc], acc_phase);          // one polling lane per warp (8 pollers instead of 256)
            tc_fence_after();
            const int row = m_blk * BLOCK_M + row_in_tile;
            const int row_warp0 = m_blk * BLOCK_M + ew * 32;
            const bool row_ok = row < p.M;
            const uint32_t taddr = tmem_base + acc * BLOCK_N + ((uint32_t)(ew * 32) << 16);
            if (p.glu) {
                constexpr int HALF = BLOCK_N / 2;
                const int col0 = n_blk * HALF;
                if (TS) {                                            // previous box of this warp must have left shared memory
                    if (lane_id() == 0) tma_store_wait_read0();
                    __syncwarp();
                }
#pragma unroll 1
                for (int c = wg * (HALF / 2); c < (wg + 1) * (HALF / 2); c += 16) {
                    uint32_t g[16], u[16];
                    tmem_ld_32x32b_x16(taddr + c, g);
                    tmem_ld_32x32b_x16(taddr + HALF + c, u);
                    tmem_ld_wait();
                    if (row_ok && col0 + c < out_cols_total) {
                        uint32_t o[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float g0 = __uint_as_float(g[2 * j]), g1 = __uint_as_float(g[2 * j + 1]);
                            float u0 = __uint_as_float(u[2 * j]), u1 = __uint_as_float(u[2 * j + 1]);
                            if (p.glu == GLU_GELU_TANH) {
                                g0 = gelu_tanh_fast(g0); g1 = gelu_tanh_fast(g1);
                            } else {
                                g0 = g0 / (1.0f + __expf(-g0)); g1 = g1 / (1.0f + __expf(-g1));
                            }
                            o[j] = pack_bf16(g0 * u0, g1 * u1);
                        }
                        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + (int64_t)row * p.ldc + col0 + c;
                        if (TS && HALF / 2 == 64) {                 // this warp's 64 output columns = one box; 16 columns = 2 x 16 B
                            const int ci = (c - wg * (HALF / 2)) >> 3;
                            *reinterpret_cast<uint4*>(srow + ((ci ^ (lane_id() & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                            *reinterpret_cast<uint4*>(srow + (((ci + 1) ^ (lane_id() & 7)) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
                        } else if (col0 + c + 16 <= out_cols_total) {
                            *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
                            *reinterpret_cast<uint4*>(dst + 8) = make_uint4(o[4], o[5], o[6], o[7]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                __nv_bfloat162 pr = *reinterpret_cast<__nv_bfloat162*>(&o[j]);
                                if (col0 + c + 2 * j < out_cols_total) dst[2 * j] = pr.x;
                                if (col0 + c + 2 * j + 1 < out_cols_total) dst[2 * j + 1] = pr.y;
                            }
                        }
                    }
                }
                if (TS && HALF / 2 == 64) {
                    fence_proxy_async();
                    __syncwarp();
                    if (lane_id() == 0 && row_warp0 < p.M) {
                        tma_store_2d(&tmap_c, stage_box, col0 + wg * (HALF / 2), row_warp0);
                        tma_store_commit();
                    }
                }
            } else {
                const int col0 = n_blk * BLOCK_N;
                const int c_begin = wg * (BLOCK_N / 2);
#pragma unroll 1
                for (int c = wg * (BLOCK_N / 2); c < (wg + 1) * (BLOCK_N / 2); c += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(taddr + c, r);
                    tmem_ld_wait();
                    const int cbase = col0 + c;
                    if (TS && (((c - c_begin) >> 5) & 1) == 0) {      // first chunk of a box: the previous box must have left shared memory
                        if (lane_id() == 0) tma_store_wait_read0();
                        __syncwarp();
                    }
                    if (row_ok && cbase < p.N) {
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                        const bool full = cbase + 32 <= p.N;
                        if (p.bias) {
                            if (full) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4) {
                                    const float4 b = *reinterpret_cast<const float4*>(p.bias + cbase + j);
                                    v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                                }
                            } else {
                                #pragma unroll
                                for (int j = 0; j < 32; ++j) if (cbase + j < p.N) v[j] += p.bias[cbase + j];
                            }
                        }
                        if (p.act == ACT_GELU_ERF) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
                        } else if (p.act == ACT_GELU_TANH) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_fast(v[j]);
                        } else if (p.act == ACT_SOFTCAP) {
                            const float inv = 1.0f / p.act_param;
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = p.act_param * tanhf(v[j] * inv);
                        } else if (p.act == ACT_SILU) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = v[j] / (1.0f + __expf(-v[j]));
                        }
                        if (p.residual) {
                            const __nv_bfloat16* rp = p.residual + (int64_t)(p.res_mod > 0 ? row % p.res_mod : row) * p.ldr + cbase;
                            if (full) {
#pragma unroll
                                for (int j = 0; j < 32; j += 8) {
                                    const uint4 q = *reinterpret_cast<const uint4*>(rp + j);
                                    float2 f;
                                    f = unpack_bf16(q.x); v[j] += f.x; v[j + 1] += f.y;
                                    f = unpack_bf16(q.y); v[j + 2] += f.x; v[j + 3] += f.y;
                                    f = unpack_bf16(q.z); v[j + 4] += f.x; v[j + 5] += f.y;
                                    f = unpack_bf16(q.w); v[j + 6] += f.x; v[j + 7] += f.y;
                                }
                            } else {
                                #pragma unroll
                                for (int j = 0; j < 32; ++j) if (cbase + j < p.N) v[j] += __bfloat162float(rp[j]);
                            }
                        }
                        if (p.out_fp32) {
                            float* dst = reinterpret_cast<float*>(p.C) + (int64_t)row * p.ldc + cbase;
                            if (full) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4)
                                    *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                            } else {
                                #pragma unroll
                                for (int j = 0; j < 32; ++j) if (cbase + j < p.N) dst[j] = v[j];
                            }
                        } else if (TS) {
                            const int half = ((c - c_begin) >> 5) & 1;
#pragma unroll
                            for (int j = 0; j < 32; j += 8)
                                *reinterpret_cast<uint4*>(srow + ((((half << 2) | (j >> 3)) ^ (lane_id() & 7)) << 4)) =
                                    make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]),
                                               pack_bf16(v[j + 4], v[j + 5]), pack_bf16(v[j + 6], v[j + 7]));
                        } else {
                            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + (int64_t)row * p.ldc + cbase;
                            if (full) {
#pragma unroll
                                for (int j = 0; j < 32; j += 8)
                                    *reinterpret_cast<uint4*>(dst + j) =
                                        make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]),
                                                   pack_bf16(v[j + 4], v[j + 5]), pack_bf16(v[j + 6], v[j + 7]));
                            } else {
                                #pragma unroll
                                for (int j = 0; j < 32; ++j) if (cbase + j < p.N) dst[j] = __float2bfloat16(v[j]);
                            }
                        }
                    }
                    if (TS && !p.out_fp32 && ((((c - c_begin) >> 5) & 1) == 1)) {     // box complete
                        fence_proxy_async();
                        __syncwarp();
                        if (lane_id() == 0 && row_warp0 < p.M) {
                            tma_store_2d(&tmap_c, stage_box, col0 + c_begin + (((c - c_begin) >> 6) << 6), row_warp0);
                            tma_store_commit();
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (TS && lane_id() == 0) tma_store_wait_read0();             // the last boxes must have left shared memory before the CTA exits
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

template <int BLOCK_N, bool TS>
static int launch_gemm_ts(const void* A, int64_t lda, const void* W, int64_t ldw, const GemmParams& p, cudaStream_t st) {
    using Cfg = GemmCfg<BLOCK_N, TS>;
    CUtensorMap ta, tb, tc;
    int rc;
    if ((rc = make_tmap_2d_bf16(&ta, A, (uint64_t)p.K, (uint64_t)p.M, (uint64_t)lda * 2, BLOCK_K, BLOCK_M))) return rc;
    if ((rc = make_tmap_2d_bf16(&tb, W, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)ldw * 2, BLOCK_K, BLOCK_N))) return rc;
    tc = ta;
    if (TS && (rc = make_tmap_2d_bf16(&tc, p.C, (uint64_t)(p.glu ? p.N / 2 : p.N), (uint64_t)p.M, (uint64_t)p.ldc * 2, 64, 32))) return rc;
    VB_SET_SMEM_ONCE(Cfg::kSmemBytes, gemm_bf16_kernel<BLOCK_N, TS>);
    const int num_m = (p.M + BLOCK_M - 1) / BLOCK_M, num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
    const int tiles = num_m * num_n;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    gemm_bf16_kernel<BLOCK_N, TS><<<grid, kNumThreads, Cfg::kSmemBytes, st>>>(ta, tb, tc, p);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// TMA-store epilogue for bf16 outputs of the 256- and 128-wide tiles (a warp's column half is then a whole number of 64-column boxes);
// VIDI_GEMM_TMASTORE=0 restores the per-thread 16-byte stores for A/B.
template <int BLOCK_N>
static int launch_gemm(const void* A, int64_t lda, const void* W, int64_t ldw, const GemmParams& p, cudaStream_t st) {
    static const int ts = getenv("VIDI_GEMM_TMASTORE") ? atoi(getenv("VIDI_GEMM_TMASTORE")) : 1;
    constexpr bool kCanTS = BLOCK_N == 256 || BLOCK_N == 128;
    if (kCanTS && ts && !p.out_fp32 && (!p.glu || BLOCK_N == 256) && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && p.ldc % 8 == 0)
        return launch_gemm_ts<BLOCK_N, kCanTS>(A, lda, W, ldw, p, st);
    return launch_gemm_ts<BLOCK_N, false>(A, lda, W, ldw, p, st);
}

bool gemm_skinny_supports(int M, int N, int K, const float* bias, const void* residual, int act, int glu);
int gemm_skinny(const void*, int64_t, const void*, int64_t, void*, int64_t, int, int, int, int, float, int, int, cudaStream_t);

int gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K,
              const float* bias, const void* residual, int64_t ldr, int res_mod, int act, float act_param, int out_fp32,
              int glu, int block_n, cudaStream_t st) {
    VB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
    // text-stream shapes (a few rows against a whole weight matrix): weight-streaming kernel with swapped operands
    // (gemm_skinny_sm100.cu); block_n < 0 forces the general kernel (A/B and tests)
    static const int skinny = getenv("VIDI_GEMM_SKINNY") ? atoi(getenv("VIDI_GEMM_SKINNY")) : 1;
    if (skinny && block_n >= 0 && gemm_skinny_supports(M, N, K, bias, residual, act, glu) && ldc % 8 == 0 &&
        (reinterpret_cast<uintptr_t>(C) & 15) == 0)
        return gemm_skinny(A, lda, W, ldw, C, ldc, M, N, K, act, act_param, out_fp32, glu, st);
    if (block_n < 0) block_n = -block_n;
    VB_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm_bf16: K/lda/ldw must be multiples of 8 (TMA 16B rows)");
    VB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(C) & 15) == 0, "gemm_bf16: pointers must be 16B aligned");
    VB_REQUIRE(ldc % 8 == 0, "gemm_bf16: ldc must be a multiple of 8");
    VB_REQUIRE(!residual || ldr % 8 == 0, "gemm_bf16: ldr must be a multiple of 8");
    GemmParams p;
    p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc; p.bias = bias;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(residual); p.ldr = ldr; p.res_mod = res_mod;
    p.act = act; p.act_param = act_param; p.out_fp32 = out_fp32; p.glu = glu;
    // m-blocks per L2-resident group: keep the group's A rows (group_m * 128 * K * 2 B) near 48 MB of the 126 MB L2 so the
    // weight matrix is re-streamed from HBM as few times as possible (ncu: 62 -> ~16 passes of W at M=126k, K=3584)
    {
        const int64_t per_block = (int64_t)BLOCK_M * K * 2;
        int64_t gm = (48ll << 20) / per_block;
        p.group_m = (int)(gm < 8 ? 8 : gm > 64 ? 64 : gm);
    }
    if (glu) VB_REQUIRE(N % block_n == 0, "gemm_bf16: GLU needs N %% block_n == 0 (packed gate|up tiles)");
    if (block_n == 256) return launch_gemm<256>(A, lda, W, ldw, p, st);
    if (block_n == 192) return launch_gemm<192>(A, lda, W, ldw, p, st);
    if (block_n == 128) return launch_gemm<128>(A, lda, W, ldw, p, st);
    if (block_n == 64) return launch_gemm<64>(A, lda, W, ldw, p, st);
    VB_REQUIRE(false, "gemm_bf16: unsupported block_n %d", block_n);
}

}  // namespace vb
