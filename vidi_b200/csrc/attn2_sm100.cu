// Ping-pong variant of the tcgen05 tower attention (attn_sm100.cu): one work item = (batch, head, PAIR of 128-query blocks).
// Each softmax warpgroup owns one query block end-to-end (thread = one query row x all 128 keys of a tile), so there is no
// cross-warpgroup max exchange and no block barrier in the tile loop; the MMA issuer alternates between the two blocks
//     QK_A(j+1) | softmax_B(j)      PV_A(j) | softmax_A(j+1) ...
// and every K/V tile brought in by TMA is used by both blocks (half the shared-memory fill traffic per query).
//   TMEM: S_A [0,128) S_B [128,256) O_A [256,352) O_B [384,480); smem: Q 2 items x 2 blocks, K/V 2 stages, P_A, P_B, ones.
// Softmax inner loop (the bound: MUFU ex2 + issue slots), per score: FFMA, MUFU.EX2, 1/2 CVT.bf16x2, 1/4 LOP3 —
//   * the row sum l is not added up by the threads but by the tensor core, as extra all-ones columns of V, so l arrives in
//     TMEM next to O and is rescaled with it.  dh=72: the V tail tile [128 keys][cols 64..79] has 8 zero-filled (TMA OOB)
//     columns 72..79 which a helper warp overwrites with 1.0 after each V tile lands; dh=64: one more N=16 MMA per k-step
//     against a constant all-ones operand;
//   * the reference m is exact for the first key tile and afterwards only moved when some p would exceed 2^8 ("lazy
//     re-referencing"); p is computed pre-scaled by 2^-7 so that test is one bit of the packed bf16 words (exponent MSB),
//     OR-ed together as they are stored; the exact-max pass + redo only runs on a tile where that bit fired;
//   * TMEM loads are software-pipelined (next 32 columns in flight while the current ones are exponentiated), and
//     setmaxnreg moves registers from the TMA/MMA warps to the softmax warpgroups (no spills).
#include "common.cuh"

namespace vb {

constexpr float kLog2eB = 1.4426950408889634f;

template <int DH>
struct Fa2Cfg {
    static constexpr int BM = 128, BN = 128;
    static constexpr int TAIL = (DH > 64) ? 16 : 0;
    static constexpr int OCOLS = 64 + TAIL;
    static constexpr int kMainBytes = 128 * 64 * 2;            // [128][64] SW128
    static constexpr int kTailBytes = 128 * 16 * 2;            // [128][16] SW32
    static constexpr int kTileBytes = kMainBytes + (TAIL ? kTailBytes : 0);
    static constexpr int kSlot = 20 * 1024;
    static constexpr int kPBytes = 2 * kMainBytes;             // [128][128] bf16
    // Q: [item parity 2][block 2] slots | K [2] | V [2] | P [2 blocks]
    static constexpr int kOffQ = 0, kOffK = 4 * kSlot, kOffV = 6 * kSlot, kOffP = 8 * kSlot;
    static constexpr int kOffOnes = kOffP + 2 * kPBytes;       // 512 B of bf16 1.0: B operand of the row-sum MMA
    static constexpr int kOffBar = kOffOnes + 512;
    static constexpr int kSmem = kOffBar + 256 + 1024;        // 160K + 64K + ... = 225.75 KB
    static constexpr int LCOL = DH;                            // TMEM column (within an O block) holding the row sum
};

struct Fa2Params {
    int S, H, B;
    int npair;               // query-block pairs per (b,h)
    int nkt;                 // key tiles
    int items;
    float scale_log2;
    __nv_bfloat16* out;
    int64_t ldo;
};

template <int DH>
__global__ void __launch_bounds__(384, 1)
attn_fwd2_sm100_kernel(const __grid_constant__ CUtensorMap tm_main, const __grid_constant__ CUtensorMap tm_tail,
                       const Fa2Params p) {
    using C = Fa2Cfg<DH>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
    uint64_t* q_full = bars;          // [2] per item parity
    uint64_t* q_empty = bars + 2;     // [2]
    uint64_t* kv_full = bars + 4;     // [2]
    uint64_t* kv_empty = bars + 6;    // [2]
    uint64_t* s_full = bars + 8;      // [2] per block
    uint64_t* s_empty = bars + 10;    // [2] (128 arrivals)
    uint64_t* p_full = bars + 12;     // [2] (128 arrivals)
    uint64_t* p_empty = bars + 14;    // [2]
    uint64_t* o_full = bars + 16;     // [2]
    uint64_t* o_empty = bars + 18;    // [2] (128 arrivals)
    uint64_t* v_ones = bars + 20;     // [2] V tail ones written (dh=72)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 22);

    const int warp = threadIdx.x >> 5;
    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tm_main);
        if (C::TAIL) tma_prefetch_desc(&tm_tail);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
            mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1);
            mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 128);
            mbar_init(&p_full[i], 128); mbar_init(&p_empty[i], 1);
            mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 128);
            mbar_init(&v_ones[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_ptr);
    if (warp == 3) {                                                   // 256 x bf16 1.0
        reinterpret_cast<uint4*>(smem + C::kOffOnes)[lane_id()] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int nkt = p.nkt;

    if (warp < 4) {
        setmaxnreg_dec<104>();
    if (warp == 0) {
        // ============================ TMA producer ============================
        if (elect_one()) {
            uint32_t g = 0, it = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
                const int pr = item % p.npair;
                const int h = (item / p.npair) % p.H;
                const int b = item / (p.npair * p.H);
                const int qi = it & 1;
                mbar_wait(&q_empty[qi], ((it >> 1) & 1) ^ 1);
                mbar_expect_tx(&q_full[qi], 2 * C::kTileBytes);
                for (int blk = 0; blk < 2; ++blk) {
                    uint8_t* sq = smem + C::kOffQ + (qi * 2 + blk) * C::kSlot;
                    const int q0 = (pr * 2 + blk) * C::BM;
                    tma_load_4d(sq, &tm_main, &q_full[qi], 0, h, q0, b, kEvictNormal);
                    if (C::TAIL) tma_load_4d(sq + C::kMainBytes, &tm_tail, &q_full[qi], 64, h, q0, b, kEvictNormal);
                }
                for (int j = 0; j < nkt; ++j, ++g) {
                    const int st = g & 1;
                    mbar_wait(&kv_empty[st], ((g >> 1) & 1) ^ 1);
                    mbar_expect_tx(&kv_full[st], 2 * C::kTileBytes);
                    uint8_t* sk = smem + C::kOffK + st * C::kSlot;
                    uint8_t* sv = smem + C::kOffV + st * C::kSlot;
                    tma_load_4d(sk, &tm_main, &kv_full[st], 0, p.H + h, j * C::BN, b, kEvictLast);
                    tma_load_4d(sv, &tm_main, &kv_full[st], 0, 2 * p.H + h, j * C::BN, b, kEvictLast);
                    if (C::TAIL) {
                        tma_load_4d(sk + C::kMainBytes, &tm_tail, &kv_full[st], 64, p.H + h, j * C::BN, b, kEvictLast);
                        tma_load_4d(sv + C::kMainBytes, &tm_tail, &kv_full[st], 64, 2 * p.H + h, j * C::BN, b, kEvictLast);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        if (elect_one()) {
            constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128);
            constexpr uint32_t idesc_pv_main = umma_idesc_bf16(128, 64, 0, 1);
            constexpr uint32_t idesc_pv_tail = umma_idesc_bf16(128, 16, 0, 1);
            uint32_t g = 0, it = 0;
            // t = per-block tile counter (same for both blocks): barrier parities derive from it
            auto issue_qk = [&](int blk, uint32_t gg, const uint8_t* sq, bool wait_kv) {
                const int st = gg & 1;
                if (wait_kv) mbar_wait(&kv_full[st], (gg >> 1) & 1);
                mbar_wait(&s_empty[blk], (gg & 1) ^ 1);
                tc_fence_after();
                const uint8_t* sk = smem + C::kOffK + st * C::kSlot;
                const uint64_t a = umma_desc_k_sw128(smem_u32(sq));
                const uint64_t bdesc = umma_desc_k_sw128(smem_u32(sk));
                const uint32_t d = tmem_base + blk * 128;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(d, a + 2 * k, bdesc + 2 * k, idesc_qk, k != 0);
                if (C::TAIL)
                    umma_f16(d, umma_desc_k_sw32(smem_u32(sq + C::kMainBytes)), umma_desc_k_sw32(smem_u32(sk + C::kMainBytes)),
                             idesc_qk, 1);
                umma_commit(&s_full[blk]);
            };
            auto issue_pv = [&](int blk, uint32_t gg) {
                const int st = gg & 1;
                if (C::TAIL && blk == 0) mbar_wait(&v_ones[st], (gg >> 1) & 1);
                mbar_wait(&p_full[blk], gg & 1);
                mbar_wait(&o_empty[blk], (gg & 1) ^ 1);
                tc_fence_after();
                const uint8_t* sp = smem + C::kOffP + blk * C::kPBytes;
                const uint8_t* sv = smem + C::kOffV + st * C::kSlot;
                const uint32_t d = tmem_base + 256 + blk * 128;
                const uint64_t b1 = umma_desc_mn_sw32(smem_u32(smem + C::kOffOnes), 2048, 256);   // every element is 1.0
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint64_t a = umma_desc_k_sw128(smem_u32(sp + (kk >> 2) * C::kMainBytes)) + 2 * (kk & 3);
                    const uint64_t bm = umma_desc_mn_sw128(smem_u32(sv + kk * 16 * 128), 8192, 1024);
                    umma_f16(d, a, bm, idesc_pv_main, kk != 0);
                    if (C::TAIL) {
                        const uint64_t bt = umma_desc_mn_sw32(smem_u32(sv + C::kMainBytes + kk * 16 * 32), 2048, 256);
                        umma_f16(d + 64, a, bt, idesc_pv_tail, kk != 0);
                    }
                    if (!C::TAIL) umma_f16(d + C::LCOL, a, b1, idesc_pv_tail, kk != 0);  // row sums of P
                }
                umma_commit(&o_full[blk]);
                umma_commit(&p_empty[blk]);
            };
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
                const int qi = it & 1;
                mbar_wait(&q_full[qi], (it >> 1) & 1);
                const uint8_t* sqA = smem + C::kOffQ + (qi * 2 + 0) * C::kSlot;
                const uint8_t* sqB = smem + C::kOffQ + (qi * 2 + 1) * C::kSlot;
                issue_qk(0, g, sqA, true);
                issue_qk(1, g, sqB, false);
                for (int j = 0; j < nkt; ++j) {
                    const bool more = j + 1 < nkt;
                    issue_pv(0, g + j);
                    if (more) issue_qk(0, g + j + 1, sqA, true);
                    issue_pv(1, g + j);
                    umma_commit(&kv_empty[(g + j) & 1]);               // both blocks' QK^T and PV of tile j are issued
                    if (more) issue_qk(1, g + j + 1, sqB, false);
                    else umma_commit(&q_empty[qi]);
                }
                g += nkt;
            }
        }
    } else if (warp == 3 && C::TAIL) {
        // ============================ V tail: columns 72..79 := 1.0 (row-sum columns) ============================
        uint32_t g = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            for (int j = 0; j < nkt; ++j, ++g) {
                const int st = g & 1;
                mbar_wait(&kv_full[st], (g >> 1) & 1);
                uint8_t* svt = smem + C::kOffV + st * C::kSlot + C::kMainBytes;
#pragma unroll
                for (int r = lane_id(); r < 128; r += 32)                          // SW32: 16-byte chunk index ^= bit 2 of the row
                    *reinterpret_cast<uint4*>(svt + r * 32 + ((1 ^ ((r >> 2) & 1)) << 4)) =
                        make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
                fence_proxy_async();
                __syncwarp();
                if (lane_id() == 0) mbar_arrive(&v_ones[st]);
            }
        }
    }
    } else {
        // ============================ softmax / output: warpgroup `blk` owns query block `blk` ============================
        setmaxnreg_inc<200>();
        const int ew = (warp - 4) & 3;
        const int blk = (warp - 4) >> 2;
        const int row = ew * 32 + lane_id();
        const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
        const uint32_t s_addr = tmem_base + blk * 128 + lane_addr;
        const uint32_t o_addr = tmem_base + 256 + blk * 128 + lane_addr;
        uint8_t* sp_row = smem + C::kOffP + blk * C::kPBytes + row * 128;
        uint32_t g = 0;
        auto take_o = [&](float (&o)[C::OCOLS], float& l, float corr_prev, uint32_t gg) {
            mbar_wait(&o_full[blk], gg & 1);
            tc_fence_after();
            uint32_t t0[32], t1[32];
            tmem_ld_32x32b_x32(o_addr, t0);
            tmem_ld_32x32b_x32(o_addr + 32, t1);
            uint32_t tl = 0;
            if (!C::TAIL) tl = tmem_ld_32x32b_x1(o_addr + C::LCOL);
            tmem_ld_wait();
            if (C::TAIL) {
                uint32_t t2[16];
                tmem_ld_32x32b_x16(o_addr + 64, t2);
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = o[i] * corr_prev + __uint_as_float(t0[i]);
#pragma unroll
                for (int i = 0; i < 32; ++i) o[32 + i] = o[32 + i] * corr_prev + __uint_as_float(t1[i]);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) o[64 + i] = o[64 + i] * corr_prev + __uint_as_float(t2[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = o[i] * corr_prev + __uint_as_float(t0[i]);
#pragma unroll
                for (int i = 0; i < 32; ++i) o[32 + i] = o[32 + i] * corr_prev + __uint_as_float(t1[i]);
            }
            if (!C::TAIL) l = l * corr_prev + __uint_as_float(tl);
            tc_fence_before();
            mbar_arrive(&o_empty[blk]);
        };
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            const int pr = item % p.npair;
            const int h = (item / p.npair) % p.H;
            const int b = item / (p.npair * p.H);
            // a warp whose 32 query rows all lie beyond S keeps the barrier protocol but does no arithmetic
            const bool dead = (pr * 2 + blk) * C::BM + ew * 32 >= p.S;
            float m = 0.f, l = 0.f, corr_prev = 1.f;
            float o[C::OCOLS];
#pragma unroll
            for (int i = 0; i < C::OCOLS; ++i) o[i] = 0.f;
            for (int j = 0; j < nkt; ++j, ++g) {
                mbar_wait(&s_full[blk], g & 1);
                tc_fence_after();
                const int nvalid = p.S - j * C::BN;                      // keys of this tile inside the sequence (may exceed 128)
                // exact row max of the tile (first tile of an item, and the rare re-reference)
                auto row_max = [&]() {
                    float mx = -INFINITY;
                    uint32_t ra[32], rb[32];
                    tmem_ld_32x32b_x32(s_addr, ra);
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        uint32_t (&r)[32] = (cc & 1) ? rb : ra;
                        uint32_t (&rn)[32] = (cc & 1) ? ra : rb;
                        tmem_ld_wait();
                        if (cc < 3) tmem_ld_32x32b_x32(s_addr + (cc + 1) * 32, rn);
                        if (nvalid >= (cc + 1) * 32) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) if (cc * 32 + i < nvalid) mx = fmaxf(mx, __uint_as_float(r[i]));
                        }
                    }
                    return mx;
                };
                // P = exp2(S*scale - ref - 7) -> bf16 -> swizzled smem; returns the OR of all packed words
                auto exp_pass = [&](float neg_ref) {
                    uint32_t ored = 0;
                    uint32_t ra[32], rb[32];
                    tmem_ld_32x32b_x32(s_addr, ra);
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        uint32_t (&r)[32] = (cc & 1) ? rb : ra;
                        uint32_t (&rn)[32] = (cc & 1) ? ra : rb;
                        tmem_ld_wait();
                        if (cc < 3) tmem_ld_32x32b_x32(s_addr + (cc + 1) * 32, rn);
                        uint8_t* spa = sp_row + (cc >> 1) * C::kMainBytes;
                        if (nvalid < (cc + 1) * 32) {                   // ragged tail of the sequence: mask (warp-uniform branch)
#pragma unroll
                            for (int i = 0; i < 32; ++i) if (cc * 32 + i >= nvalid) r[i] = 0xff800000u;
                        }
#pragma unroll
                        for (int c8 = 0; c8 < 4; ++c8) {
                            float pv[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pv[e])
                                    : "f"(fmaf(__uint_as_float(r[c8 * 8 + e]), p.scale_log2, neg_ref)));
                            const uint4 w = make_uint4(pack_bf16(pv[0], pv[1]), pack_bf16(pv[2], pv[3]), pack_bf16(pv[4], pv[5]),
                                                       pack_bf16(pv[6], pv[7]));
                            ored |= (w.x | w.y) | (w.z | w.w);
                            const int chunk = (cc & 1) * 4 + c8;               // 16-byte chunk index within the 64-key atom
                            *reinterpret_cast<uint4*>(spa + ((chunk ^ (row & 7)) << 4)) = w;
                        }
                    }
                    return ored;
                };
                float corr = 1.f;
                if (!dead) {
                    if (j == 0) m = row_max();
                    mbar_wait(&p_empty[blk], (g & 1) ^ 1);
                    const uint32_t ored = exp_pass(fmaf(-m, p.scale_log2, -7.f));
                    // exponent MSB of a bf16 half set  <=>  p * 2^-7 >= 2 (or inf / NaN): the reference is stale for that row
                    if (j > 0 && __any_sync(0xffffffffu, (ored & 0x40004000u) != 0)) {
                        const float m_new = fmaxf(m, row_max());
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(corr) : "f"((m - m_new) * p.scale_log2));
                        m = m_new;
                        exp_pass(fmaf(-m, p.scale_log2, -7.f));
                    }
                } else {
                    mbar_wait(&p_empty[blk], (g & 1) ^ 1);
                }
                tc_fence_before();
                mbar_arrive(&s_empty[blk]);
                fence_proxy_async();
                mbar_arrive(&p_full[blk]);
                if (j > 0) take_o(o, l, corr_prev, g - 1);
                corr_prev = corr;
            }
            take_o(o, l, corr_prev, g - 1);
            const float inv = 1.f / (C::TAIL ? o[DH < C::OCOLS ? DH : 0] : l);
            const int q = (pr * 2 + blk) * C::BM + row;
            if (q < p.S) {
                __nv_bfloat16* dst = p.out + ((int64_t)b * p.S + q) * p.ldo + h * DH;
#pragma unroll
                for (int i = 0; i < DH; i += 8)
                    *reinterpret_cast<uint4*>(dst + i) =
                        make_uint4(pack_bf16(o[i] * inv, o[i + 1] * inv), pack_bf16(o[i + 2] * inv, o[i + 3] * inv),
                                   pack_bf16(o[i + 4] * inv, o[i + 5] * inv), pack_bf16(o[i + 6] * inv, o[i + 7] * inv));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

template <int DH>
static int launch_fa2(const void* qkv, int64_t ld, void* out, int64_t ldo, int B, int S, int H, float scale, cudaStream_t st) {
    using C = Fa2Cfg<DH>;
    CUtensorMap tm_main, tm_tail;
    uint64_t dims[4] = {(uint64_t)DH, (uint64_t)(3 * H), (uint64_t)S, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)DH * 2, (uint64_t)ld * 2, (uint64_t)ld * 2 * (uint64_t)S};
    uint32_t box_main[4] = {64, 1, 128, 1};
    uint32_t box_tail[4] = {16, 1, 128, 1};
    int rc;
    if ((rc = make_tmap_nd_bf16(&tm_main, qkv, 4, dims, strides, box_main, 128))) return rc;
    if (C::TAIL) {
        if ((rc = make_tmap_nd_bf16(&tm_tail, qkv, 4, dims, strides, box_tail, 32))) return rc;
    } else {
        tm_tail = tm_main;
    }
    VB_SET_SMEM_ONCE(C::kSmem, attn_fwd2_sm100_kernel<DH>);
    Fa2Params p;
    p.S = S; p.H = H; p.B = B;
    p.npair = (S + 2 * C::BM - 1) / (2 * C::BM);
    p.nkt = (S + C::BN - 1) / C::BN;
    p.items = B * H * p.npair;
    p.scale_log2 = scale * kLog2eB;
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.ldo = ldo;
    const int grid = p.items < num_sms() ? p.items : num_sms();
    attn_fwd2_sm100_kernel<DH><<<grid, 384, C::kSmem, st>>>(tm_main, tm_tail, p);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int attn_dense_sm100_pp(const void* qkv, int64_t ld, void* out, int64_t ldo, int B, int S, int H, int dh, float scale,
                        cudaStream_t st) {
    VB_REQUIRE(ld == (int64_t)3 * H * dh, "attn_dense_sm100_pp: qkv must be packed [B*S, 3*H*dh]");
    VB_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && ldo % 8 == 0, "attn_dense_sm100_pp: alignment");
    if (B == 0 || S == 0) return 0;
    if (dh == 72) return launch_fa2<72>(qkv, ld, out, ldo, B, S, H, scale, st);
    if (dh == 64) return launch_fa2<64>(qkv, ld, out, ldo, B, S, H, scale, st);
    VB_REQUIRE(false, "attn_dense_sm100_pp: unsupported head_dim %d", dh);
}

}  // namespace vb
