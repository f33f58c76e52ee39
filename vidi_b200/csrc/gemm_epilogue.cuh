// Epilogue of the bf16 GEMM family (shared by the 2-CTA kernel): one call drains this warp's share of one
// accumulator tile from TMEM and applies bias / activation / GLU / residual / soft-cap, storing bf16 or fp32 rows.
#pragma once
#include "common.cuh"

namespace vb {
namespace g2 {

enum Act : int { ACT_NONE = 0, ACT_GELU_ERF = 1, ACT_GELU_TANH = 2, ACT_SOFTCAP = 3, ACT_SILU = 4 };
enum Glu : int { GLU_NONE = 0, GLU_GELU_TANH = 1, GLU_SILU = 2 };

struct GemmParams {
    int M, N, K;
    void* C;
    int64_t ldc;
    const float* bias;
    const __nv_bfloat16* residual;
    int64_t ldr;
    int res_mod;
    int act;
    float act_param;
    int out_fp32;
    int glu;
    int group_m;
    // LayerNorm folded into this GEMM (pre-LN transformer blocks of the towers): A holds the raw residual stream x, W was
    // pre-multiplied by gamma, and the epilogue applies  out = rstd*(acc - mean*colsum) + bias'  per row, with (sum, sumsq) of
    // each row of x given as ln_parts partial pairs written by the GEMM that produced x (stats_out below).
    const float2* ln_stats = nullptr;
    int ln_parts = 0;
    const float* ln_colsum = nullptr;    // [N]: sum_k W'[n,k]
    float ln_inv_k = 0.f, ln_eps = 0.f;
    // per-row (sum, sumsq) of the bf16 values this GEMM stores, one pair per (column tile, column half): [M, 2*num_n_tiles]
    float2* stats_out = nullptr;
    int stats_parts = 0;
    int relaxed_arrive = 0;              // A/B: release the accumulator with a relaxed (not release.cluster) arrive
    int ragged_tail = 1;                 // A/B: narrow MMA on the ragged last column tile (0: full-width MMA over zero-filled W rows)
};

// taddr: TMEM address of this warp's lane quarter at the accumulator's first column; row: global output row of this thread;
// wg: which half of the tile columns this warp drains; n_blk: tile column index.
// TS: bf16 rows leave through shared memory and TMA tensor stores (UTMASTG): every pair of 32-column chunks is staged as one
// [32 rows x 64 columns] box in this warp's 4 KB buffer (128-byte swizzle: 16-byte chunk index XOR (row & 7), conflict-free
// st.shared.v4) and written by ONE cp.async.bulk.tensor store, which also clips rows >= M and columns >= N.
// RT (implies TS): the residual rows also travel by TMA.  Before the accumulator is even ready, one lane fetches the residual box(es) of
// this warp's column half into the staging buffers (one 4 KB buffer per box, mbarrier rbar[b]); each thread then reads ITS 16-byte
// slots, adds, and writes the result back in place, and the same buffer leaves through the TMA store -- no per-thread row-strided
// global loads (32 sector requests per warp instruction) are left in the epilogue.
template <int BLOCK_N, bool LN = false, bool TS = false, bool RT = false>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t taddr, int row, int n_blk, int wg,
                                              const CUtensorMap* tmap_c = nullptr, uint8_t* stage = nullptr, int row_warp0 = 0,
                                              uint64_t* rbar = nullptr, uint32_t rphase = 0) {
    const bool row_ok = row < p.M;
    const int out_cols_total = p.glu ? p.N / 2 : p.N;
    if (p.glu) {
        constexpr int HALF = BLOCK_N / 2;
        constexpr bool GTS = TS && !RT && HALF / 2 == 64;       // this warp's 64 output columns = one TMA-store box (256-wide tiles)
        const int col0 = n_blk * HALF;
        if (GTS) {                                              // previous box of this warp must have left shared memory
            if ((threadIdx.x & 31) == 0) tma_store_wait_read0();
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
                if (GTS) {
                    uint8_t* srow = stage + (threadIdx.x & 31) * 128;
                    const int ci = (c - wg * (HALF / 2)) >> 3;
                    *reinterpret_cast<uint4*>(srow + ((ci ^ (threadIdx.x & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<uint4*>(srow + (((ci + 1) ^ (threadIdx.x & 7)) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
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
        if (GTS) {
            fence_proxy_async();
            __syncwarp();
            if ((threadIdx.x & 31) == 0 && row_warp0 < p.M) {
                tma_store_2d(tmap_c, stage, col0 + wg * (HALF / 2), row_warp0);
                tma_store_commit();
            }
        }
    } else {
        const int col0 = n_blk * BLOCK_N;
        float ln_a = 1.f, ln_b = 0.f;                       // folded LayerNorm: v = ln_a * acc - ln_b * colsum
        if (LN && p.ln_stats) {
            float s1 = 0.f, s2 = 0.f;
            if (row_ok) {
                const float2* sp = p.ln_stats + (int64_t)row * p.ln_parts;
#pragma unroll 4
                for (int i = 0; i < p.ln_parts; ++i) {
                    const float2 t = sp[i];
                    s1 += t.x; s2 += t.y;
                }
            }
            const float mean = s1 * p.ln_inv_k;
            ln_a = rsqrtf(fmaxf(s2 * p.ln_inv_k - mean * mean, 0.f) + p.ln_eps);
            ln_b = ln_a * mean;
        }
        float so1 = 0.f, so2 = 0.f;                         // row statistics of what this thread stores
        // The residual row segment of chunk c+1 is requested before chunk c is processed, so its HBM latency overlaps the
        // TMEM load + math + stores of the current chunk instead of serialising four ~1 us round trips per tile.
        const __nv_bfloat16* res_row =
            (!RT && p.residual && row_ok) ? p.residual + (int64_t)(p.res_mod > 0 ? row % p.res_mod : row) * p.ldr : nullptr;
        uint4 rq_next[4];
        // columns at and beyond N were never computed when the last tile ran a narrower MMA: stop at the tile's valid width
        const int c_begin = wg * (BLOCK_N / 2), c_end = min((wg + 1) * (BLOCK_N / 2), p.ragged_tail ? (((p.N - col0 + 31) >> 5) << 5) : BLOCK_N);
        if (res_row && col0 + c_begin + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 4; ++j) rq_next[j] = *reinterpret_cast<const uint4*>(res_row + col0 + c_begin + j * 8);
        }
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += 32) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(taddr + c, r);
            uint4 rq[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) rq[j] = rq_next[j];
            if (res_row && c + 32 < c_end && col0 + c + 64 <= p.N) {
#pragma unroll
                for (int j = 0; j < 4; ++j) rq_next[j] = *reinterpret_cast<const uint4*>(res_row + col0 + c + 32 + j * 8);
            }
            tmem_ld_wait();
            const int cbase = col0 + c;
            if (TS && !RT && (((c - c_begin) >> 5) & 1) == 0) {   // first chunk of a box: the previous box must have left shared memory
                if ((threadIdx.x & 31) == 0) tma_store_wait_read0();
                __syncwarp();
            }
            uint8_t* sbox = RT ? stage + ((c - c_begin) >> 6) * 4096 : stage;      // RT: one buffer per box of this column half
            if (RT && (((c - c_begin) >> 5) & 1) == 0) mbar_wait(&rbar[(c - c_begin) >> 6], rphase);   // residual box has landed
            if (row_ok && cbase < p.N) {
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                const bool full = cbase + 32 <= p.N;
                if (LN && p.ln_stats) {
                    if (full) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 cs = *reinterpret_cast<const float4*>(p.ln_colsum + cbase + j);
                            v[j] = ln_a * v[j] - ln_b * cs.x; v[j + 1] = ln_a * v[j + 1] - ln_b * cs.y;
                            v[j + 2] = ln_a * v[j + 2] - ln_b * cs.z; v[j + 3] = ln_a * v[j + 3] - ln_b * cs.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (cbase + j < p.N) v[j] = ln_a * v[j] - ln_b * p.ln_colsum[cbase + j];
                    }
                }
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
                if (RT) {
                    const int half = ((c - c_begin) >> 5) & 1;
                    const uint8_t* rrow = sbox + (threadIdx.x & 31) * 128;
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        const uint4 q = *reinterpret_cast<const uint4*>(rrow + ((((half << 2) | (j >> 3)) ^ (threadIdx.x & 7)) << 4));
                        float2 f;
                        f = unpack_bf16(q.x); v[j] += f.x; v[j + 1] += f.y;
                        f = unpack_bf16(q.y); v[j + 2] += f.x; v[j + 3] += f.y;
                        f = unpack_bf16(q.z); v[j + 4] += f.x; v[j + 5] += f.y;
                        f = unpack_bf16(q.w); v[j + 6] += f.x; v[j + 7] += f.y;
                    }
                } else if (p.residual) {
                    const __nv_bfloat16* rp = p.residual + (int64_t)(p.res_mod > 0 ? row % p.res_mod : row) * p.ldr + cbase;
                    if (full) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            const uint4 q = rq[j >> 3];
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
                    // staged: 4 x 16 B of this row into the box of chunk pair (c / 64); statistics as in the direct path
                    const int half = ((c - c_begin) >> 5) & 1;
                    uint8_t* srow = sbox + (threadIdx.x & 31) * 128;
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        const uint4 q = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]),
                                                   pack_bf16(v[j + 4], v[j + 5]), pack_bf16(v[j + 6], v[j + 7]));
                        *reinterpret_cast<uint4*>(srow + ((((half << 2) | (j >> 3)) ^ (threadIdx.x & 7)) << 4)) = q;
                        if (LN && p.stats_out) {
                            float2 f;
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                f = unpack_bf16(t == 0 ? q.x : t == 1 ? q.y : t == 2 ? q.z : q.w);
                                if (cbase + j + 2 * t < p.N) { so1 += f.x; so2 = fmaf(f.x, f.x, so2); }
                                if (cbase + j + 2 * t + 1 < p.N) { so1 += f.y; so2 = fmaf(f.y, f.y, so2); }
                            }
                        }
                    }
                } else {
                    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + (int64_t)row * p.ldc + cbase;
                    if (full) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            const uint4 q = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]),
                                                       pack_bf16(v[j + 4], v[j + 5]), pack_bf16(v[j + 6], v[j + 7]));
                            *reinterpret_cast<uint4*>(dst + j) = q;
                            if (LN && p.stats_out) {                     // statistics of the ROUNDED values (what the next LN sees)
                                float2 f;
                                f = unpack_bf16(q.x); so1 += f.x + f.y; so2 = fmaf(f.x, f.x, fmaf(f.y, f.y, so2));
                                f = unpack_bf16(q.y); so1 += f.x + f.y; so2 = fmaf(f.x, f.x, fmaf(f.y, f.y, so2));
                                f = unpack_bf16(q.z); so1 += f.x + f.y; so2 = fmaf(f.x, f.x, fmaf(f.y, f.y, so2));
                                f = unpack_bf16(q.w); so1 += f.x + f.y; so2 = fmaf(f.x, f.x, fmaf(f.y, f.y, so2));
                            }
                        }
                    } else {
                        #pragma unroll
                        for (int j = 0; j < 32; ++j) if (cbase + j < p.N) {
                            const __nv_bfloat16 q = __float2bfloat16(v[j]);
                            dst[j] = q;
                            if (LN) {
                                const float f = __bfloat162float(q);
                                so1 += f; so2 = fmaf(f, f, so2);
                            }
                        }
                    }
                }
            }
            if (TS && ((((c - c_begin) >> 5) & 1) == 1 || c + 32 >= c_end)) {     // box complete (or last, half-filled box of a ragged tile)
                fence_proxy_async();
                __syncwarp();
                if ((threadIdx.x & 31) == 0 && row_warp0 < p.M) {
                    tma_store_2d(tmap_c, sbox, col0 + c_begin + (((c - c_begin) >> 6) << 6), row_warp0);
                    tma_store_commit();
                }
            }
        }
        if (RT) {                // a box the ragged tile never touched still has a residual load in flight: consume its barrier phase
            constexpr int NBOX = BLOCK_N / 128;
#pragma unroll
            for (int b = 0; b < NBOX; ++b)
                if (c_begin + b * 64 >= c_end) mbar_wait(&rbar[b], rphase);
        }
        if (LN && p.stats_out && row_ok) p.stats_out[(int64_t)row * p.stats_parts + n_blk * 2 + wg] = make_float2(so1, so2);
    }
}

}  // namespace g2
}  // namespace vb
