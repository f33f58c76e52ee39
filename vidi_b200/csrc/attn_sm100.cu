// tcgen05 / TMEM / TMA bidirectional flash attention for the SigLIP (dh=72) and Whisper (dh=64) towers (K3 / K8).
//
//   persistent CTAs, one work item = (batch b, head h, 128-query block); 384 threads:
//     warp 0      TMA producer   Q tile (double-buffered across items) and 128-key K/V tiles (2 stages), 4-D tensor maps
//                                over qkv viewed as [dh, 3*H, S, B] so rows >= S and head columns >= dh are zero-filled
//     warp 1      MMA issuer     S_j = Q K_j^T  (kind::f16, M=128, N=128) into TMEM (2 buffers); O_j = P_j V_j (fresh
//                                accumulator per tile, 2 buffers) with P read K-major from shared memory and V read
//                                MN-major straight from its row-major TMA tile (no transpose anywhere)
//     warp 2      TMEM alloc
//     warps 4-11  softmax        two warpgroups x 4 lane quarters: thread = one query row x half of the 128 keys.
//                                Online softmax in base 2, P -> bf16 -> 128B-swizzled smem; O_j is pulled from TMEM and
//                                accumulated in registers with the running-max correction, so TMEM is never rescaled.
//   Head dim 72 is handled as a 64-wide 128B-swizzle tile plus a 16-wide 32B-swizzle tail tile whose columns 72..79 are
//   TMA zero fill: 4 + 1 UMMA k-steps for QK^T, N = 64 + 16 for PV.
#include "common.cuh"

namespace vb {

constexpr float kLog2eA = 1.4426950408889634f;

template <int DH>
struct FaCfg {
    static constexpr int BM = 128, BN = 128;
    static constexpr int TAIL = (DH > 64) ? 16 : 0;            // padded tail columns (DH - 64 rounded up to 16)
    static constexpr int OCOLS = 64 + TAIL;                    // accumulator columns of O
    static constexpr int kMainBytes = 128 * 64 * 2;            // 16 KB  [128 rows][64] SW128
    static constexpr int kTailBytes = 128 * 16 * 2;            // 4 KB   [128 rows][16] SW32
    static constexpr int kTileBytes = kMainBytes + (TAIL ? kTailBytes : 0);
    static constexpr int kPBytes = 2 * kMainBytes;             // P [128][128] as two 64-wide SW128 atoms
    // smem map (1024-aligned pieces): Q[2] | K[2] | V[2] | P[2] | barriers
    static constexpr int kSlot = 20 * 1024;                    // tile slot (main 16K + tail 4K)
    static constexpr int kOffQ = 0, kOffK = 2 * kSlot, kOffV = 4 * kSlot, kOffP = 6 * kSlot;
    static constexpr int kOffBar = kOffP + 2 * kPBytes;
    static constexpr int kSmem = kOffBar + 256 + 1024;
};

struct FaParams {
    int S, H, B;
    int nqb;                 // query blocks per (b,h)
    int nkt;                 // key tiles
    int items;               // B * H * nqb
    float scale_log2;
    __nv_bfloat16* out;
    int64_t ldo;
};

template <int DH>
__global__ void __launch_bounds__(384, 1)
attn_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tm_main, const __grid_constant__ CUtensorMap tm_tail,
                      const FaParams p) {
    using C = FaCfg<DH>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
    uint64_t* q_full = bars;          // [2]
    uint64_t* q_empty = bars + 2;     // [2]
    uint64_t* kv_full = bars + 4;     // [2]
    uint64_t* kv_empty = bars + 6;    // [2]
    uint64_t* s_full = bars + 8;      // [2]  MMA -> softmax
    uint64_t* s_empty = bars + 10;    // [2]  softmax -> MMA (256 arrivals)
    uint64_t* p_full = bars + 12;     // [2]  softmax -> MMA (256 arrivals)
    uint64_t* p_empty = bars + 14;    // [2]  MMA -> softmax
    uint64_t* o_full = bars + 16;     // [2]  MMA -> softmax
    uint64_t* o_empty = bars + 18;    // [2]  softmax -> MMA (256 arrivals)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 20);

    const int warp = threadIdx.x >> 5;
    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tm_main);
        if (C::TAIL) tma_prefetch_desc(&tm_tail);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
            mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1);
            mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 256);
            mbar_init(&p_full[i], 256); mbar_init(&p_empty[i], 1);
            mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 256);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // TMEM columns: S buffers at 0 / 128, O buffers at 256 / 384
    const int nkt = p.nkt;

    if (warp == 0) {
        // ============================ TMA producer ============================
        if (elect_one()) {
            uint32_t g = 0;                                            // global KV tile counter
            uint32_t it = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
                const int qb = item % p.nqb;
                const int h = (item / p.nqb) % p.H;
                const int b = item / (p.nqb * p.H);
                const int qi = it & 1;
                mbar_wait(&q_empty[qi], ((it >> 1) & 1) ^ 1);
                mbar_expect_tx(&q_full[qi], C::kTileBytes);
                uint8_t* sq = smem + C::kOffQ + qi * C::kSlot;
                tma_load_4d(sq, &tm_main, &q_full[qi], 0, h, qb * C::BM, b, kEvictNormal);
                if (C::TAIL) tma_load_4d(sq + C::kMainBytes, &tm_tail, &q_full[qi], 64, h, qb * C::BM, b, kEvictNormal);
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
            constexpr uint32_t idesc_pv_main = umma_idesc_bf16(128, 64, 0, 1);       // B (= V) is MN-major
            constexpr uint32_t idesc_pv_tail = umma_idesc_bf16(128, 16, 0, 1);
            uint32_t g = 0, it = 0;
            auto issue_qk = [&](uint32_t gg, const uint8_t* sq) {
                const int st = gg & 1;
                mbar_wait(&kv_full[st], (gg >> 1) & 1);
                mbar_wait(&s_empty[st], ((gg >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint8_t* sk = smem + C::kOffK + st * C::kSlot;
                const uint64_t a = umma_desc_k_sw128(smem_u32(sq));
                const uint64_t bdesc = umma_desc_k_sw128(smem_u32(sk));
                const uint32_t d = tmem_base + st * 128;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(d, a + 2 * k, bdesc + 2 * k, idesc_qk, k != 0);
                if (C::TAIL)
                    umma_f16(d, umma_desc_k_sw32(smem_u32(sq + C::kMainBytes)), umma_desc_k_sw32(smem_u32(sk + C::kMainBytes)),
                             idesc_qk, 1);
                umma_commit(&s_full[st]);
            };
            auto issue_pv = [&](uint32_t gg) {
                const int st = gg & 1;
                mbar_wait(&p_full[st], (gg >> 1) & 1);
                mbar_wait(&o_empty[st], ((gg >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint8_t* sp = smem + C::kOffP + st * C::kPBytes;
                const uint8_t* sv = smem + C::kOffV + st * C::kSlot;
                const uint32_t d = tmem_base + 256 + st * 128;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {                      // 16 keys per step
                    const uint64_t a = umma_desc_k_sw128(smem_u32(sp + (kk >> 2) * C::kMainBytes)) + 2 * (kk & 3);
                    const uint64_t bm = umma_desc_mn_sw128(smem_u32(sv + kk * 16 * 128), 8192, 1024);
                    umma_f16(d, a, bm, idesc_pv_main, kk != 0);
                    if (C::TAIL) {
                        const uint64_t bt = umma_desc_mn_sw32(smem_u32(sv + C::kMainBytes + kk * 16 * 32), 2048, 256);
                        umma_f16(d + 64, a, bt, idesc_pv_tail, kk != 0);
                    }
                }
                umma_commit(&o_full[st]);
                umma_commit(&kv_empty[st]);
                umma_commit(&p_empty[st]);
            };
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
                const int qi = it & 1;
                mbar_wait(&q_full[qi], (it >> 1) & 1);
                const uint8_t* sq = smem + C::kOffQ + qi * C::kSlot;
                issue_qk(g, sq);
                for (int j = 0; j < nkt; ++j) {
                    if (j + 1 < nkt) issue_qk(g + j + 1, sq);
                    else umma_commit(&q_empty[qi]);                   // all QK^T of this item issued: Q slot reusable when done
                    issue_pv(g + j);
                }
                g += nkt;
            }
        }
    } else if (warp >= 4) {
        // ============================ softmax / output ============================
        const int ew = (warp - 4) & 3;
        const int wg = (warp - 4) >> 2;                               // 0: keys [0,64) + O cols [0,40) ; 1: keys [64,128) + rest
        const int row = ew * 32 + lane_id();
        const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
        uint32_t g = 0;
        // Each thread owns one query row and half of every 128-key tile; the two halves keep separate running (m, l)
        // and separate O partial sums over *their* keys, merged through shared memory at the end of the item.
        // To keep a single running max per row (needed because O_j is produced from BOTH halves' P), the halves
        // exchange their tile maxima through smem before computing P.
        __shared__ float xmax[2][2][128];                              // [buffer][wg][row]
        __shared__ float xsum[2][128];                                 // [wg][row] final l exchange
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            const int qb = item % p.nqb;
            const int h = (item / p.nqb) % p.H;
            const int b = item / (p.nqb * p.H);
            float m = -INFINITY, l = 0.f, corr_prev = 0.f;
            float o[C::OCOLS / 2];
#pragma unroll
            for (int i = 0; i < C::OCOLS / 2; ++i) o[i] = 0.f;
            for (int j = 0; j < nkt; ++j, ++g) {
                const int st = g & 1;
                const uint32_t ph = (g >> 1) & 1;
                mbar_wait(&s_full[st], ph);
                tc_fence_after();
                const uint32_t s_addr = tmem_base + st * 128 + wg * 64 + lane_addr;
                uint32_t r0[32], r1[32];
                tmem_ld_32x32b_x32(s_addr, r0);
                tmem_ld_32x32b_x32(s_addr + 32, r1);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&s_empty[st]);                             // S is in registers now
                const int kbase = j * C::BN + wg * 64;
                // running max on the RAW scores (scale > 0 is monotone); the scale is folded into the exp2 FMA below
                if (kbase + 64 > p.S) {                                // only the last key tile can be ragged
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (kbase + i >= p.S) r0[i] = 0xff800000u;     // -inf
                        if (kbase + 32 + i >= p.S) r1[i] = 0xff800000u;
                    }
                }
                float mx = -INFINITY;
#pragma unroll
                for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(r0[i]), __uint_as_float(r1[i])));
                xmax[st][wg][row] = mx;
                asm volatile("bar.sync 1, 256;" ::: "memory");         // softmax warps only
                const float m_new = fmaxf(m, fmaxf(mx, xmax[st][wg ^ 1][row]));
                float corr;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(corr) : "f"((m - m_new) * p.scale_log2));   // first tile: 2^-inf = 0
                m = m_new;
                const float neg_m = -m_new * p.scale_log2;
                mbar_wait(&p_empty[st], ph ^ 1);
                uint8_t* sp = smem + C::kOffP + st * C::kPBytes + wg * C::kMainBytes + row * 128;
                float rs = 0.f;
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {                       // 8 chunks of 8 keys (16 B)
                    float pv[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int i = c8 * 8 + e;
                        const float sv = __uint_as_float(i < 32 ? r0[i] : r1[i - 32]);
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pv[e]) : "f"(fmaf(sv, p.scale_log2, neg_m)));
                        rs += pv[e];
                    }
                    const uint4 q = make_uint4(pack_bf16(pv[0], pv[1]), pack_bf16(pv[2], pv[3]), pack_bf16(pv[4], pv[5]),
                                               pack_bf16(pv[6], pv[7]));
                    *reinterpret_cast<uint4*>(sp + ((c8 ^ (row & 7)) << 4)) = q;
                }
                l = l * corr + rs;
                fence_proxy_async();
                mbar_arrive(&p_full[st]);
                // consume the previous tile's O while the tensor core works on this one
                if (j > 0) {
                    const int sp_ = (g - 1) & 1;
                    mbar_wait(&o_full[sp_], ((g - 1) >> 1) & 1);
                    tc_fence_after();
                    const uint32_t o_addr = tmem_base + 256 + sp_ * 128 + wg * (C::OCOLS / 2) + lane_addr;
                    if (C::OCOLS == 80) {
                        uint32_t t0[32], t1[16];
                        tmem_ld_32x32b_x32(o_addr, t0);
                        // remaining 8 columns
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                     : "=r"(t1[0]), "=r"(t1[1]), "=r"(t1[2]), "=r"(t1[3]), "=r"(t1[4]), "=r"(t1[5]), "=r"(t1[6]), "=r"(t1[7])
                                     : "r"(o_addr + 32));
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = o[i] * corr_prev + __uint_as_float(t0[i]);
#pragma unroll
                        for (int i = 0; i < 8; ++i) o[32 + i] = o[32 + i] * corr_prev + __uint_as_float(t1[i]);
                    } else {
                        uint32_t t0[32];
                        tmem_ld_32x32b_x32(o_addr, t0);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = o[i] * corr_prev + __uint_as_float(t0[i]);
                    }
                    tc_fence_before();
                    mbar_arrive(&o_empty[sp_]);
                }
                corr_prev = corr;
            }
            // last tile's O
            {
                const uint32_t gl = g - 1;
                const int sp_ = gl & 1;
                mbar_wait(&o_full[sp_], (gl >> 1) & 1);
                tc_fence_after();
                const uint32_t o_addr = tmem_base + 256 + sp_ * 128 + wg * (C::OCOLS / 2) + lane_addr;
                if (C::OCOLS == 80) {
                    uint32_t t0[32], t1[8];
                    tmem_ld_32x32b_x32(o_addr, t0);
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(t1[0]), "=r"(t1[1]), "=r"(t1[2]), "=r"(t1[3]), "=r"(t1[4]), "=r"(t1[5]), "=r"(t1[6]), "=r"(t1[7])
                                 : "r"(o_addr + 32));
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = o[i] * corr_prev + __uint_as_float(t0[i]);
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[32 + i] = o[32 + i] * corr_prev + __uint_as_float(t1[i]);
                } else {
                    uint32_t t0[32];
                    tmem_ld_32x32b_x32(o_addr, t0);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = o[i] * corr_prev + __uint_as_float(t0[i]);
                }
                tc_fence_before();
                mbar_arrive(&o_empty[sp_]);
            }
            // the two key-halves kept separate row sums; the row max is shared, so l_total = l_0 + l_1
            xsum[wg][row] = l;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const float inv = 1.f / (l + xsum[wg ^ 1][row]);
            asm volatile("bar.sync 1, 256;" ::: "memory");             // xsum may be overwritten by the next item
            const int q = qb * C::BM + row;
            if (q < p.S) {
                __nv_bfloat16* dst = p.out + ((int64_t)b * p.S + q) * p.ldo + h * DH + wg * (C::OCOLS / 2);
                constexpr int NOUT = (DH - 0) / 2 > C::OCOLS / 2 ? C::OCOLS / 2 : C::OCOLS / 2;
                // wg 0 writes O columns [0, OCOLS/2), wg 1 writes [OCOLS/2, DH)
                const int ncols = wg == 0 ? C::OCOLS / 2 : DH - C::OCOLS / 2;
#pragma unroll
                for (int i = 0; i < NOUT; i += 8) {
                    if (i < ncols) {
                        const uint4 v4 = make_uint4(pack_bf16(o[i] * inv, o[i + 1] * inv), pack_bf16(o[i + 2] * inv, o[i + 3] * inv),
                                                    pack_bf16(o[i + 4] * inv, o[i + 5] * inv), pack_bf16(o[i + 6] * inv, o[i + 7] * inv));
                        *reinterpret_cast<uint4*>(dst + i) = v4;
                    }
                }
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
static int launch_fa(const void* qkv, int64_t ld, void* out, int64_t ldo, int B, int S, int H, float scale, cudaStream_t st) {
    using C = FaCfg<DH>;
    CUtensorMap tm_main, tm_tail;
    // qkv viewed as [dh (contig), 3*H heads, S tokens, B frames]; sections Q | K | V are head indices [0,H), [H,2H), [2H,3H)
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
    VB_SET_SMEM_ONCE(C::kSmem, attn_fwd_sm100_kernel<DH>);
    FaParams p;
    p.S = S; p.H = H; p.B = B;
    p.nqb = (S + C::BM - 1) / C::BM;
    p.nkt = (S + C::BN - 1) / C::BN;
    p.items = B * H * p.nqb;
    p.scale_log2 = scale * kLog2eA;
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.ldo = ldo;
    const int grid = p.items < num_sms() ? p.items : num_sms();
    attn_fwd_sm100_kernel<DH><<<grid, 384, C::kSmem, st>>>(tm_main, tm_tail, p);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// qkv must be the packed [B*S, 3*H*dh] projection output (Q | K | V sections), ld == 3*H*dh
int attn_dense_sm100(const void* qkv, int64_t ld, void* out, int64_t ldo, int B, int S, int H, int dh, float scale,
                     cudaStream_t st) {
    VB_REQUIRE(ld == (int64_t)3 * H * dh, "attn_dense_sm100: qkv must be packed [B*S, 3*H*dh]");
    VB_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && ldo % 8 == 0, "attn_dense_sm100: alignment");
    if (B == 0 || S == 0) return 0;
    if (dh == 72) return launch_fa<72>(qkv, ld, out, ldo, B, S, H, scale, st);
    if (dh == 64) return launch_fa<64>(qkv, ld, out, ldo, B, S, H, scale, st);
    VB_REQUIRE(false, "attn_dense_sm100: unsupported head_dim %d", dh);
}

}  // namespace vb
