// tcgen05 / TMEM / TMA split-KV cross attention for the soft-capped (Gemma2) text->image / text->audio path (K15).
//
// Semantics (gemma.py:50-96, xattn.py:141-263): non-causal, no RoPE, s = cap*tanh(q.k*scale/cap), key-padding mask,
// GQA with un-repeated K/V.  One CTA = (key split, KV head, block of up to 128 "virtual rows" r = t*G + g) and streams
// its key range in 64-key tiles:
//     warp 0   TMA: Q once (3-D map [dh, Hq, T] -> rows ordered t*G+g), K/V tiles (2-D maps over the K||V cache),
//              2 stages of 64 KB -> 128 KB in flight per SM
//     warp 1   S_j = Q K_j^T (M=128, N=64, 16 k-steps) into TMEM (2 buffers); O += P_j V_j (N=256, V read MN-major from
//              its row-major tile, 4 k-steps) accumulated IN TMEM across the whole split
//     warps 4-11  softmax: thread = (row, 32 keys).  Because the logits are soft-capped to |s| <= cap, softmax is
//              evaluated against a FIXED reference  p = exp2(s*log2e - M_REF)  with M_REF chosen so that p can neither
//              overflow nor underflow for any admissible logit (|s*log2e| <= 72.2 for cap = 50): no running max, no
//              rescaling of O, no cross-thread max exchange -- the kernel is a pure K/V stream.
// Output: normalised partial O [split][T][Hq][DH] fp32 and natural-log LSE [split][T][Hq] (-inf for an empty split).
//
// CAPPED = false (Vidi-7B: dh = 128, no soft-cap, Vidi_7B/model/lmm/dattn/xattn.py:99-175): the same stream, but the reference is PER
// ROW and fixed after the first key tile (row max of that tile; the two threads of a row agree through shared memory once).  Because
// P is rounded to bf16 and accumulated in fp32 -- both with 8 exponent bits -- a later logit above the reference costs no precision,
// it only moves p above 1; the row keeps its running max of (s - ref) and, in the never-observed case that it exceeds 2^100, the CTA
// repeats its split once with the exact row max as the reference (flag + second pass), so no admissible input can overflow.
#include "common.cuh"

namespace vb {

constexpr float kLog2eX = 1.4426950408889634f;
constexpr float kLn2X = 0.6931471805599453f;

template <int DH>
struct XsCfg {
    static constexpr int BN = 64;                              // keys per tile
    static constexpr int ATOMS = DH / 64;                      // 64-column swizzle atoms per row
    static constexpr int kAtomQ = 128 * 128;                   // [128 rows][64] bf16 = 16 KB
    static constexpr int kAtomKV = BN * 128;                   // [64 keys][64] bf16 = 8 KB
    static constexpr int kQBytes = ATOMS * kAtomQ;             // 64 KB (DH=256)
    static constexpr int kKBytes = ATOMS * kAtomKV;            // 32 KB
    static constexpr int kStage = 2 * kKBytes;                 // K + V
    static constexpr int kPBytes = 128 * 128;                  // [128 rows][64 keys] bf16 = 16 KB
    static constexpr int kOffQ = 0, kOffKV = kQBytes, kOffP = kOffKV + 2 * kStage, kOffBar = kOffP + 2 * kPBytes;
    static constexpr int kSmem = kOffBar + 192 + 1024;
    static constexpr int kTmemO = 128;                         // O columns start (S buffers at 0 and 64)
};

struct XsParams {
    int T, N, Hq, G;
    // up to two key segments (image rows, audio rows of the K||V cache) share one grid: global split index -> (segment, local split)
    int nseg;
    int seg_split0[2];           // first global split of the segment
    int seg_row0[2];             // first cache row of the segment
    int seg_rows[2];             // keys in the segment
    int seg_kps[2];              // keys per split (multiple of 64)
    const uint8_t* seg_mask[2];  // key-padding mask of the segment (indexed from its first row) or nullptr
    int rows_per_block;          // 128 / G tokens * G
    float scale_log2;            // un-capped: scale * log2(e)
    float scale_over_cap;        // scale / cap
    float cap_log2;              // cap * log2(e)
    float m_ref;                 // fixed softmax reference (log2 units)
    float* Opart;
    float* LSE;
};

template <int DH, bool CAPPED>
__global__ void __launch_bounds__(384, 1)
xattn_splitkv_sm100_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                           const __grid_constant__ CUtensorMap tm_v, const XsParams p) {
    using C = XsCfg<DH>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
    uint64_t* q_full = bars;          // [1]
    uint64_t* k_full = bars + 1;      // [2]   K and V tiles have separate rings: a K slot is released as soon as its
    uint64_t* k_empty = bars + 3;     // [2]   QK^T has been issued, without waiting for softmax + PV of that tile
    uint64_t* s_full = bars + 5;      // [2]
    uint64_t* s_empty = bars + 7;     // [2] (256 arrivals)
    uint64_t* p_full = bars + 9;      // [2] (256 arrivals)
    uint64_t* p_empty = bars + 11;    // [2]
    uint64_t* o_full = bars + 13;     // [1]
    uint64_t* v_full = bars + 14;     // [2]
    uint64_t* v_empty = bars + 16;    // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 18);

    const int split = blockIdx.x, hk = blockIdx.y, qb = blockIdx.z;
    const int warp = threadIdx.x >> 5;
    const int sg = (p.nseg > 1 && split >= p.seg_split0[1]) ? 1 : 0;
    const int seg_lo = p.seg_row0[sg], seg_hi = seg_lo + p.seg_rows[sg];
    const int k_begin = seg_lo + (split - p.seg_split0[sg]) * p.seg_kps[sg];
    const int k_end = min(seg_hi, k_begin + p.seg_kps[sg]);
    const uint8_t* kmask = p.seg_mask[sg];
    const int ntiles = (max(0, k_end - k_begin) + C::BN - 1) / C::BN;
    const int t0 = qb * (128 / p.G);

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
    }
    __shared__ float xsum[2][128];
    __shared__ int retry_flag;
    if (warp == 1 && elect_one()) {
        retry_flag = 0;
        mbar_init(q_full, 1); mbar_init(o_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
            mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 256);
            mbar_init(&p_full[i], 256); mbar_init(&p_empty[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // softmax-thread state (only meaningful for warps >= 4), kept across the optional second pass of the un-capped variant
    const int ew = (warp - 4) & 3;                     // TMEM lane quarter == row group
    const int ch = (warp - 4) >> 2;                    // which 32 keys of the 64-key tile
    const int row = ew * 32 + lane_id();
    const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
    const int nrows = min(p.rows_per_block, (p.T - t0) * p.G);
    const bool active = row < nrows;                   // warp-uniform except in the boundary warp
    const bool warp_active = ew * 32 < nrows;
    float l4[4] = {0.f, 0.f, 0.f, 0.f};               // independent partial row sums
    float mref = CAPPED ? p.m_ref : 0.f;               // softmax reference in log2 units (per row when un-capped)
    float smax = -INFINITY;                            // un-capped: running max of the row's scaled logits

    int attempt = 0;
    while (true) {
    const int base = attempt * ntiles;                 // global tile counter: ring slots and barrier parities continue across passes
    if (warp == 0) {
        // Q + K producer
        if (elect_one() && ntiles > 0) {
            if (attempt == 0) {
                mbar_expect_tx(q_full, C::kQBytes);
                for (int a = 0; a < C::ATOMS; ++a)
                    tma_load_3d(smem + C::kOffQ + a * C::kAtomQ, &tm_q, q_full, a * 64, hk * p.G, t0, kEvictLast);
            }
            for (int j = 0; j < ntiles; ++j) {
                const int jj = base + j, st = jj & 1;
                mbar_wait(&k_empty[st], ((jj >> 1) & 1) ^ 1);
                mbar_expect_tx(&k_full[st], C::kKBytes);
                uint8_t* sk = smem + C::kOffKV + st * C::kStage;
                const int key = k_begin + j * C::BN;
                for (int a = 0; a < C::ATOMS; ++a)
                    tma_load_2d(sk + a * C::kAtomKV, &tm_k, &k_full[st], hk * DH + a * 64, key, kEvictFirst);
            }
        }
    } else if (warp == 3) {
        // V producer
        if (elect_one() && ntiles > 0) {
            for (int j = 0; j < ntiles; ++j) {
                const int jj = base + j, st = jj & 1;
                mbar_wait(&v_empty[st], ((jj >> 1) & 1) ^ 1);
                mbar_expect_tx(&v_full[st], C::kKBytes);
                uint8_t* sv = smem + C::kOffKV + st * C::kStage + C::kKBytes;
                const int key = k_begin + j * C::BN;
                for (int a = 0; a < C::ATOMS; ++a)
                    tma_load_2d(sv + a * C::kAtomKV, &tm_v, &v_full[st], hk * DH + a * 64, key, kEvictFirst);
            }
        }
    } else if (warp == 1) {
        if (elect_one() && ntiles > 0) {
            constexpr uint32_t idesc_qk = umma_idesc_bf16(128, C::BN);
            constexpr uint32_t idesc_pv = umma_idesc_bf16(128, DH, 0, 1);            // V is MN-major
            mbar_wait(q_full, 0);
            auto issue_qk = [&](int j) {
                const int jj = base + j, st = jj & 1;
                mbar_wait(&k_full[st], (jj >> 1) & 1);
                mbar_wait(&s_empty[st], ((jj >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint8_t* sk = smem + C::kOffKV + st * C::kStage;
                const uint32_t d = tmem_base + st * C::BN;
#pragma unroll
                for (int kk = 0; kk < DH / 16; ++kk) {
                    const uint64_t a = umma_desc_k_sw128(smem_u32(smem + C::kOffQ + (kk >> 2) * C::kAtomQ)) + 2 * (kk & 3);
                    const uint64_t b = umma_desc_k_sw128(smem_u32(sk + (kk >> 2) * C::kAtomKV)) + 2 * (kk & 3);
                    umma_f16(d, a, b, idesc_qk, kk != 0);
                }
                umma_commit(&k_empty[st]);
                umma_commit(&s_full[st]);
            };
            auto issue_pv = [&](int j) {
                const int jj = base + j, st = jj & 1;
                mbar_wait(&v_full[st], (jj >> 1) & 1);
                mbar_wait(&p_full[st], (jj >> 1) & 1);
                tc_fence_after();
                const uint8_t* sv = smem + C::kOffKV + st * C::kStage + C::kKBytes;
                const uint8_t* sp = smem + C::kOffP + st * C::kPBytes;
                const uint32_t d = tmem_base + C::kTmemO;
#pragma unroll
                for (int kk = 0; kk < C::BN / 16; ++kk) {
                    const uint64_t a = umma_desc_k_sw128(smem_u32(sp)) + 2 * kk;
                    // V tile: ATOMS chunks of [64 keys][64 dh]; chunk stride (LBO) = kAtomKV, 8-key groups (SBO) = 1024 B
                    const uint64_t b = umma_desc_mn_sw128(smem_u32(sv + kk * 16 * 128), C::kAtomKV, 1024);
                    umma_f16(d, a, b, idesc_pv, (j | kk) != 0);
                }
                umma_commit(&v_empty[st]);
                umma_commit(&p_empty[st]);
            };
            issue_qk(0);
            for (int j = 0; j < ntiles; ++j) {
                if (j + 1 < ntiles) issue_qk(j + 1);
                issue_pv(j);
            }
            umma_commit(o_full);
        }
    } else if (warp >= 4) {
        for (int j = 0; j < ntiles; ++j) {
            const int jj = base + j, st = jj & 1;
            const uint32_t ph = (jj >> 1) & 1;
            mbar_wait(&s_full[st], ph);
            tc_fence_after();
            uint32_t r[32];
            if (warp_active) {
                tmem_ld_32x32b_x32(tmem_base + st * C::BN + ch * 32 + lane_addr, r);
                tmem_ld_wait();
            }
            tc_fence_before();
            mbar_arrive(&s_empty[st]);
            const int kb = k_begin + j * C::BN + ch * 32;
            uint32_t mbits = 0xffffffffu;
            if (warp_active && kmask) {
                mbits = 0;
                const uint8_t* km = kmask + (kb - seg_lo);                            // kb - seg_lo is a multiple of 32
                if (kb + 32 <= seg_hi) {
                    const uint4 m0 = *reinterpret_cast<const uint4*>(km);
                    const uint4 m1 = *reinterpret_cast<const uint4*>(km + 16);
                    const uint32_t w[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
                    for (int i = 0; i < 32; ++i) mbits |= (((w[i >> 2] >> ((i & 3) * 8)) & 0xffu) ? 1u : 0u) << i;
                } else {
                    for (int i = 0; i < 32 && kb + i < seg_hi; ++i) mbits |= (km[i] ? 1u : 0u) << i;
                }
            }
            if (!CAPPED && j == 0 && attempt == 0) {
                // per-row reference = the row's max over the first key tile; the two threads of a row agree through shared memory
                float m = -INFINITY;
                if (warp_active) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if ((kb + i < k_end) && ((mbits >> i) & 1u)) m = fmaxf(m, __uint_as_float(r[i]) * p.scale_log2);
                }
                xsum[ch][row] = m;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                m = fmaxf(xsum[0][row], xsum[1][row]);
                mref = (m == -INFINITY) ? 0.f : m;
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            mbar_wait(&p_empty[st], ph ^ 1);
            if (warp_active) {
                uint8_t* sp = smem + C::kOffP + st * C::kPBytes + row * 128;
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    float pv[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int i = c8 * 8 + e;
                        float sl;
                        if (CAPPED) sl = p.cap_log2 * tanh_fast(__uint_as_float(r[i]) * p.scale_over_cap);
                        else sl = __uint_as_float(r[i]) * p.scale_log2;
                        float pe;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pe) : "f"(sl - mref));
                        const bool ok = (kb + i < k_end) && ((mbits >> i) & 1u);
                        pv[e] = ok ? pe : 0.f;
                        if (!CAPPED && ok) smax = fmaxf(smax, sl);
                        l4[e & 3] += pv[e];
                    }
                    const uint4 q = make_uint4(pack_bf16(pv[0], pv[1]), pack_bf16(pv[2], pv[3]), pack_bf16(pv[4], pv[5]),
                                               pack_bf16(pv[6], pv[7]));
                    *reinterpret_cast<uint4*>(sp + (((ch * 4 + c8) ^ (row & 7)) << 4)) = q;
                }
                fence_proxy_async();
            }
            mbar_arrive(&p_full[st]);
        }
        // the fixed reference is only valid while p = 2^(s - ref) stays inside fp32 / bf16 range for the row's largest logit
        if (!CAPPED && attempt == 0 && active && smax != -INFINITY && fabsf(smax - mref) > 100.f) retry_flag = 1;
    }
    if (CAPPED) break;
    __syncthreads();
    const bool again = attempt == 0 && retry_flag != 0;
    if (!again) break;
    // second pass over the same keys with the exact row max as the reference (cannot trigger again)
    if (warp >= 4) {
        if (ntiles > 0) { mbar_wait(o_full, 0); tc_fence_after(); }      // pass 1's MMAs have drained before TMEM is reused
        xsum[ch][row] = smax;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float m = fmaxf(xsum[0][row], xsum[1][row]);
        mref = (m == -INFINITY) ? 0.f : m;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        l4[0] = l4[1] = l4[2] = l4[3] = 0.f;
        tc_fence_before();
    }
    __syncthreads();
    attempt = 1;
    }

    if (warp >= 4) {
        // ---- epilogue: normalise O (TMEM) by the row sum and write the partial ----
        const float l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        xsum[ch][row] = l;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float lt = xsum[0][row] + xsum[1][row];
        if (ntiles > 0) {
            mbar_wait(o_full, attempt & 1);
            tc_fence_after();
        }
        if (warp_active) {
            const int t = t0 + row / p.G, head = hk * p.G + row % p.G;
            const int64_t rowid = ((int64_t)split * p.T + (active ? t : 0)) * p.Hq + head;
            const float inv = lt > 0.f ? 1.f / lt : 0.f;
            float* op = p.Opart + rowid * DH + ch * (DH / 2);
#pragma unroll 1
            for (int c = 0; c < DH / 2; c += 32) {
                uint32_t o[32];
                if (ntiles > 0) {
                    tmem_ld_32x32b_x32(tmem_base + C::kTmemO + ch * (DH / 2) + c + lane_addr, o);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = 0;
                }
                if (active) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4*>(op + c + i) =
                            make_float4(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv,
                                        __uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
                }
            }
            if (active && ch == 0) p.LSE[rowid] = lt > 0.f ? (mref + log2f(lt)) * kLn2X : -INFINITY;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// Host launcher.  CAPPED: softcap > 0 with softcap*log2(e) <= 80 (Gemma2: 50), DH == 256; un-capped: DH == 128 (Vidi-7B).  G in {1,2,4,8}.
// K / V point at cache row 0 of the layer; segment i covers rows [row0[i], row0[i] + rows[i]) split into splits[i] key ranges.
// Opart fp32 [splits[0] + splits[1]][T][Hq][DH], LSE fp32 [splits[0] + splits[1]][T][Hq]: segment 1's partials follow segment 0's.
template <int DH, bool CAPPED>
static int launch_xs(const void* Q, int64_t ldq, const void* K, const void* V, int64_t ldkv, int n_rows_total, int nseg,
                     const int* row0, const int* rows, const int* splits, const uint8_t* const* masks, int T, int Hq, int Hkv,
                     float scale, float softcap, float* Opart, float* LSE, cudaStream_t st) {
    using C = XsCfg<DH>;
    const int G = Hq / Hkv;
    if (CAPPED) VB_REQUIRE(softcap > 0.f && softcap * kLog2eX <= 80.f, "xattn_splitkv_sm100 needs a soft-cap with cap*log2e <= 80");
    VB_REQUIRE(128 % G == 0 && nseg >= 1 && nseg <= 2, "xattn_splitkv_sm100: G=%d nseg=%d", G, nseg);
    CUtensorMap tq, tk, tv;
    int rc;
    {   // Q viewed as [dh, Hq, T]: a box {64, G, 128/G} lands as rows ordered t*G + g
        uint64_t dims[3] = {(uint64_t)DH, (uint64_t)Hq, (uint64_t)T};
        uint64_t strides[2] = {(uint64_t)DH * 2, (uint64_t)ldq * 2};
        uint32_t box[3] = {64, (uint32_t)G, (uint32_t)(128 / G)};
        if ((rc = make_tmap_nd_bf16(&tq, Q, 3, dims, strides, box, 128))) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)Hkv * DH, (uint64_t)(n_rows_total > 0 ? n_rows_total : 1)};
        uint64_t strides[1] = {(uint64_t)ldkv * 2};
        uint32_t box[2] = {64, (uint32_t)C::BN};
        if ((rc = make_tmap_nd_bf16(&tk, K, 2, dims, strides, box, 128))) return rc;
        if ((rc = make_tmap_nd_bf16(&tv, V, 2, dims, strides, box, 128))) return rc;
    }
    VB_SET_SMEM_ONCE(C::kSmem, xattn_splitkv_sm100_kernel<DH, CAPPED>);
    XsParams p;
    p.T = T; p.N = n_rows_total; p.Hq = Hq; p.G = G; p.nseg = nseg;
    int total = 0;
    for (int i = 0; i < 2; ++i) {
        const bool on = i < nseg;
        VB_REQUIRE(!on || (splits[i] >= 1 && rows[i] >= 0 && row0[i] >= 0 && (rows[i] == 0 || row0[i] + rows[i] <= n_rows_total)),
                   "xattn_splitkv_sm100: bad segment %d (row0=%d rows=%d splits=%d of %d cache rows)", i, on ? row0[i] : 0,
                   on ? rows[i] : 0, on ? splits[i] : 0, n_rows_total);
        int kps = on ? (rows[i] + splits[i] - 1) / splits[i] : 64;
        kps = ((kps + 63) / 64) * 64;
        if (kps == 0) kps = 64;
        p.seg_split0[i] = total;
        p.seg_row0[i] = on ? row0[i] : 0;
        p.seg_rows[i] = on ? rows[i] : 0;
        p.seg_kps[i] = kps;
        p.seg_mask[i] = on && masks ? masks[i] : nullptr;
        VB_REQUIRE(p.seg_mask[i] == nullptr || (reinterpret_cast<uintptr_t>(p.seg_mask[i]) & 15) == 0, "xattn_splitkv_sm100: mask alignment");
        if (on) total += splits[i];
    }
    p.rows_per_block = 128;
    p.scale_log2 = scale * kLog2eX;
    p.scale_over_cap = CAPPED ? scale / softcap : 0.f;
    p.cap_log2 = CAPPED ? softcap * kLog2eX : 0.f;
    // admissible logits: |s*log2e| <= cap_log2.  With M_REF = cap_log2 - 96:  s - M_REF in [96 - 2*cap_log2, 96]
    // -> p in [2^-48.3, 2^96] for cap = 50: no underflow to zero, no overflow of p, of the row sum or of P.V in fp32.
    p.m_ref = CAPPED ? p.cap_log2 - 96.f : 0.f;
    p.Opart = Opart; p.LSE = LSE;
    dim3 grid(total, Hkv, (T * G + 127) / 128);
    xattn_splitkv_sm100_kernel<DH, CAPPED><<<grid, 384, C::kSmem, st>>>(tq, tk, tv, p);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// which (dh, soft-cap) combinations the tcgen05 kernel covers: Gemma2 (256, capped) and Mistral (128, un-capped)
bool xattn_sm100_supports(int dh, float softcap) {
    return (dh == 256 && softcap > 0.f && softcap * kLog2eX <= 80.f) || (dh == 128 && softcap == 0.f);
}

int xattn_splitkv_sm100_seg(const void* Q, int64_t ldq, const void* K, const void* V, int64_t ldkv, int n_rows_total, int nseg,
                            const int* row0, const int* rows, const int* splits, const uint8_t* const* masks, int T, int Hq, int Hkv,
                            int dh, float scale, float softcap, float* Opart, float* LSE, cudaStream_t st) {
    if (dh == 256 && softcap > 0.f)
        return launch_xs<256, true>(Q, ldq, K, V, ldkv, n_rows_total, nseg, row0, rows, splits, masks, T, Hq, Hkv, scale, softcap, Opart, LSE, st);
    if (dh == 128 && softcap == 0.f)
        return launch_xs<128, false>(Q, ldq, K, V, ldkv, n_rows_total, nseg, row0, rows, splits, masks, T, Hq, Hkv, scale, softcap, Opart, LSE, st);
    VB_REQUIRE(false, "xattn_splitkv_sm100: unsupported head_dim %d / soft-cap %g", dh, (double)softcap);
}

int xattn_splitkv_sm100(const void* Q, int64_t ldq, const void* K, const void* V, int64_t ldkv, const uint8_t* kmask, int T,
                        int N, int Hq, int Hkv, int dh, int splits, float scale, float softcap, float* Opart, float* LSE, cudaStream_t st) {
    const int row0 = 0;
    return xattn_splitkv_sm100_seg(Q, ldq, K, V, ldkv, N, 1, &row0, &N, &splits, &kmask, T, Hq, Hkv, dh, scale, softcap, Opart, LSE, st);
}

}  // namespace vb
