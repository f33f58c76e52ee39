// Attention kernels of the Vidi prefill path.
//   attn_dense      : bidirectional flash attention for the SigLIP (dh=72) and Whisper (dh=64) towers (K3/K8)
//   xattn_splitkv   : text->image / text->audio cross attention (K15): tiny Q (groups*T rows per KV head) against
//                     a very long un-repeated K/V stream; flash-decoding style split over keys, emits (O, LSE)
//   xattn_merge     : log-sum-exp merge of the partials (across splits and, after the all-gather, across ranks)
//   attn_text       : causal / sliding-window / soft-capped self attention of the short text stream (K16)
//   rope_inplace    : rotate-half RoPE on the q|k sections of the text qkv buffer
// Round-1 implementation of the two tile kernels uses warp-level mma.sync (bf16 m16n8k16) with cp.async staging;
// the tcgen05/TMA version of xattn_splitkv is the next step.
#include "common.cuh"

namespace vb {

// ---- warp-level MMA helpers ----------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t addr, uint32_t& r0, uint32_t& r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr float kLog2e = 1.4426950408889634f;

// =================================================================================================
// Dense bidirectional flash attention.  qkv [B*S, ld] with Q at column q_off + h*DH, K at k_off + h*DH,
// V at v_off + h*DH; out [B*S, ldo] at column h*DH.  One CTA = 64 queries of one (batch, head); 4 warps x 16 rows.
// =================================================================================================
template <int DH, int DHP>
struct DenseCfg {
    static constexpr int BM = 64, BN = 64;
    static constexpr int LDS = DHP + 8;                       // elements; (LDS*2/16) odd -> conflict-free ldmatrix
    static constexpr int kSmem = (BM + 4 * BN) * LDS * 2;     // Q + 2x(K,V)
};

template <int DH, int DHP>
__global__ void __launch_bounds__(128)
attn_dense_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t ld, int q_off, int k_off, int v_off,
                  __nv_bfloat16* __restrict__ out, int64_t ldo, int S, int H, float scale_log2) {
    using Cfg = DenseCfg<DH, DHP>;
    constexpr int BM = Cfg::BM, BN = Cfg::BN, LDS = Cfg::LDS;
    constexpr int CH = DH / 8;                                // 16-byte chunks per row
    constexpr int KS = DHP / 16;                              // k-steps of QK^T
    constexpr int NT = DH / 8;                                // n-tiles of PV
    extern __shared__ __align__(16) uint8_t smem_dense[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_dense);
    __nv_bfloat16* sK = sQ + BM * LDS;                        // [2][BN][LDS]
    __nv_bfloat16* sV = sK + 2 * BN * LDS;                    // [2][BN][LDS]

    const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t row0 = (int64_t)b * S;
    const int q0 = qb * BM;
    const int ntiles = (S + BN - 1) / BN;

    // zero the pad columns [DH, DHP) of Q and both K buffers once (V pad columns are never read)
    if (DHP > DH) {
        for (int i = tid; i < (BM + 2 * BN) * (DHP - DH); i += 128) {
            const int r = i / (DHP - DH), c = DH + i % (DHP - DH);
            sQ[r * LDS + c] = __float2bfloat16(0.f);          // sQ and sK are contiguous: r spans both
        }
    }
    // Q tile
    for (int i = tid; i < BM * CH; i += 128) {
        const int r = i / CH, c = i % CH;
        const bool ok = q0 + r < S;
        cp_async16(smem_u32(sQ + r * LDS + c * 8), qkv + (row0 + (ok ? q0 + r : 0)) * ld + q_off + h * DH + c * 8, ok);
    }
    auto load_kv = [&](int tile, int buf) {
        const int k0 = tile * BN;
        for (int i = tid; i < BN * CH; i += 128) {
            const int r = i / CH, c = i % CH;
            const bool ok = k0 + r < S;
            const __nv_bfloat16* src = qkv + (row0 + (ok ? k0 + r : 0)) * ld + h * DH + c * 8;
            cp_async16(smem_u32(sK + (buf * BN + r) * LDS + c * 8), src + k_off, ok);
            cp_async16(smem_u32(sV + (buf * BN + r) * LDS + c * 8), src + v_off, ok);
        }
    };
    load_kv(0, 0);
    cp_async_commit();

    float o[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // rows g and g+8 of this warp's 16
    uint32_t qf[KS][4];

    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < ntiles) load_kv(t + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (t == 0) {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
                ldsm_x4(smem_u32(sQ + (warp * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8), qf[ks][0], qf[ks][1],
                        qf[ks][2], qf[ks][3]);
        }
        // S = Q K^T  (16 x 64 per warp)
        float s[BN / 8][4];
#pragma unroll
        for (int i = 0; i < BN / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
        const __nv_bfloat16* kb = sK + buf * BN * LDS;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
            for (int np = 0; np < BN / 16; ++np) {
                uint32_t b0, b1, b2, b3;
                const int r = np * 16 + (lane & 7) + ((lane >> 4) << 3);
                const int c = ks * 16 + ((lane >> 3) & 1) * 8;
                ldsm_x4(smem_u32(kb + r * LDS + c), b0, b1, b2, b3);
                mma_bf16(s[2 * np], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b0, b1);
                mma_bf16(s[2 * np + 1], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b2, b3);
            }
        }
        // mask keys beyond S, online softmax in base 2
        const int kbase = t * BN + (lane & 3) * 2;
        float mx0 = m0, mx1 = m1;
#pragma unroll
        for (int i = 0; i < BN / 8; ++i) {
            const int kc = kbase + i * 8;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool ok = (kc + (e & 1)) < S;
                s[i][e] = ok ? s[i][e] * scale_log2 : -INFINITY;
            }
            mx0 = fmaxf(mx0, fmaxf(s[i][0], s[i][1]));
            mx1 = fmaxf(mx1, fmaxf(s[i][2], s[i][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float c0 = exp2f(m0 - mx0), c1 = exp2f(m1 - mx1);   // m=-inf on first tile -> exp2(-inf)=0
        m0 = mx0; m1 = mx1;
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int i = 0; i < BN / 8; ++i) {
            s[i][0] = exp2f(s[i][0] - m0); s[i][1] = exp2f(s[i][1] - m0);
            s[i][2] = exp2f(s[i][2] - m1); s[i][3] = exp2f(s[i][3] - m1);
            rs0 += s[i][0] + s[i][1]; rs1 += s[i][2] + s[i][3];
        }
        l0 = l0 * c0 + rs0; l1 = l1 * c1 + rs1;
#pragma unroll
        for (int i = 0; i < NT; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
        // O += P V
        const __nv_bfloat16* vb = sV + buf * BN * LDS;
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
            const uint32_t a0 = pack_bf16(s[2 * kk][0], s[2 * kk][1]), a1 = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
            const uint32_t a2 = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]), a3 = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
            for (int np = 0; np < NT / 2; ++np) {
                uint32_t b0, b1, b2, b3;
                const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = np * 16 + (lane >> 4) * 8;
                ldsm_x4_t(smem_u32(vb + r * LDS + c), b0, b1, b2, b3);
                mma_bf16(o[2 * np], a0, a1, a2, a3, b0, b1);
                mma_bf16(o[2 * np + 1], a0, a1, a2, a3, b2, b3);
            }
            if (NT & 1) {
                uint32_t b0, b1;
                const int r = kk * 16 + (lane & 15);
                ldsm_x2_t(smem_u32(vb + r * LDS + (NT - 1) * 8), b0, b1);
                mma_bf16(o[NT - 1], a0, a1, a2, a3, b0, b1);
            }
        }
        __syncthreads();
    }
    // finalize: reduce l over the quad, normalise, stage through this warp's sQ rows, coalesced store
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    __nv_bfloat16* st = sQ + warp * 16 * LDS;
    const int g = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        *reinterpret_cast<uint32_t*>(st + g * LDS + i * 8 + tq * 2) = pack_bf16(o[i][0] * i0, o[i][1] * i0);
        *reinterpret_cast<uint32_t*>(st + (g + 8) * LDS + i * 8 + tq * 2) = pack_bf16(o[i][2] * i1, o[i][3] * i1);
    }
    __syncwarp();
    for (int i = lane; i < 16 * CH; i += 32) {
        const int r = i / CH, c = i % CH;
        const int q = q0 + warp * 16 + r;
        if (q < S)
            *reinterpret_cast<uint4*>(out + (row0 + q) * ldo + h * DH + c * 8) = *reinterpret_cast<const uint4*>(st + r * LDS + c * 8);
    }
}

// =================================================================================================
// Split-KV cross attention (gemma.py:50-96, xattn.py:141-263 semantics: non-causal, no RoPE, scale,
// tanh soft-cap, key-padding mask).  Q [T, Hq*DH]; K [N, ldkv] / V [N, ldkv] un-repeated (Hkv heads).
// CTA (split, kv head, q block): 64 virtual rows r -> (t = r / G, g = r % G), q head = hk*G + g.
// Writes normalised partial O [split][T][Hq][DH] fp32 and LSE [split][T][Hq] (natural log, -inf if empty).
// =================================================================================================
template <int DH>
struct XCfg {
    static constexpr int BM = 64, BN = 32, STAGES = 3;
    static constexpr int LDS = DH + 8;
    static constexpr int kSmem = (BM + STAGES * 2 * BN) * LDS * 2;
};

template <int DH>
__global__ void __launch_bounds__(128)
xattn_splitkv_kernel(const __nv_bfloat16* __restrict__ Q, int64_t ldq, const __nv_bfloat16* __restrict__ K,
                     const __nv_bfloat16* __restrict__ V, int64_t ldkv, const uint8_t* __restrict__ kmask, int T, int N,
                     int Hq, int G, int keys_per_split, float scale, float softcap, float* __restrict__ Opart,
                     float* __restrict__ LSE) {
    using Cfg = XCfg<DH>;
    constexpr int BM = Cfg::BM, BN = Cfg::BN, LDS = Cfg::LDS, STAGES = Cfg::STAGES;
    constexpr int CH = DH / 8, KS = DH / 16, NT = DH / 8;
    extern __shared__ __align__(16) uint8_t smem_x[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_x);
    __nv_bfloat16* sK = sQ + BM * LDS;                        // [STAGES][BN][LDS]
    __nv_bfloat16* sV = sK + STAGES * BN * LDS;

    const int split = blockIdx.x, hk = blockIdx.y, qb = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nrows = T * G;
    const int k_begin = split * keys_per_split;
    const int k_end = min(N, k_begin + keys_per_split);
    const int ntiles = (max(0, k_end - k_begin) + BN - 1) / BN;

    for (int i = tid; i < BM * CH; i += 128) {
        const int r = i / CH, c = i % CH;
        const int vr = qb * BM + r;
        const bool ok = vr < nrows;
        const int t = ok ? vr / G : 0, g = ok ? vr % G : 0;
        cp_async16(smem_u32(sQ + r * LDS + c * 8), Q + (int64_t)t * ldq + (hk * G + g) * DH + c * 8, ok);
    }
    auto load_kv = [&](int tile, int buf) {
        const int k0 = k_begin + tile * BN;
        for (int i = tid; i < BN * CH; i += 128) {
            const int r = i / CH, c = i % CH;
            const bool ok = k0 + r < k_end;
            const int64_t off = (int64_t)(ok ? k0 + r : 0) * ldkv + hk * DH + c * 8;
            cp_async16(smem_u32(sK + (buf * BN + r) * LDS + c * 8), K + off, ok);
            cp_async16(smem_u32(sV + (buf * BN + r) * LDS + c * 8), V + off, ok);
        }
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < ntiles) load_kv(s, s);
        cp_async_commit();
    }

    float o[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const float inv_cap = softcap > 0.f ? 1.f / softcap : 0.f;
    const float cap_log2 = softcap * kLog2e;

    for (int t = 0; t < ntiles; ++t) {
        const int buf = t % STAGES;
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (t + STAGES - 1 < ntiles) load_kv(t + STAGES - 1, (t + STAGES - 1) % STAGES);
        cp_async_commit();

        float s[BN / 8][4];
#pragma unroll
        for (int i = 0; i < BN / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
        const __nv_bfloat16* kb = sK + buf * BN * LDS;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            uint32_t a0, a1, a2, a3;
            ldsm_x4(smem_u32(sQ + (warp * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8), a0, a1, a2, a3);
#pragma unroll
            for (int np = 0; np < BN / 16; ++np) {
                uint32_t b0, b1, b2, b3;
                const int r = np * 16 + (lane & 7) + ((lane >> 4) << 3);
                const int c = ks * 16 + ((lane >> 3) & 1) * 8;
                ldsm_x4(smem_u32(kb + r * LDS + c), b0, b1, b2, b3);
                mma_bf16(s[2 * np], a0, a1, a2, a3, b0, b1);
                mma_bf16(s[2 * np + 1], a0, a1, a2, a3, b2, b3);
            }
        }
        const int kbase = k_begin + t * BN + (lane & 3) * 2;
        float mx0 = m0, mx1 = m1;
#pragma unroll
        for (int i = 0; i < BN / 8; ++i) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int key = kbase + i * 8 + (e & 1);
                bool ok = key < k_end;
                if (ok && kmask) ok = kmask[key] != 0;
                float x = s[i][e] * scale;
                if (softcap > 0.f) x = cap_log2 * tanh_fast(x * inv_cap); else x *= kLog2e;
                s[i][e] = ok ? x : -INFINITY;
            }
            mx0 = fmaxf(mx0, fmaxf(s[i][0], s[i][1]));
            mx1 = fmaxf(mx1, fmaxf(s[i][2], s[i][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        // guard fully-masked rows: keep the reference max finite so exp2(-inf - (-inf)) never appears
        const float r0 = (mx0 == -INFINITY) ? 0.f : mx0, r1 = (mx1 == -INFINITY) ? 0.f : mx1;
        const float c0 = exp2f(m0 - r0), c1 = exp2f(m1 - r1);
        m0 = mx0; m1 = mx1;
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int i = 0; i < BN / 8; ++i) {
            s[i][0] = exp2f(s[i][0] - r0); s[i][1] = exp2f(s[i][1] - r0);
            s[i][2] = exp2f(s[i][2] - r1); s[i][3] = exp2f(s[i][3] - r1);
            rs0 += s[i][0] + s[i][1]; rs1 += s[i][2] + s[i][3];
        }
        l0 = l0 * c0 + rs0; l1 = l1 * c1 + rs1;
#pragma unroll
        for (int i = 0; i < NT; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
        const __nv_bfloat16* vb = sV + buf * BN * LDS;
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
            const uint32_t a0 = pack_bf16(s[2 * kk][0], s[2 * kk][1]), a1 = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
            const uint32_t a2 = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]), a3 = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
            for (int np = 0; np < NT / 2; ++np) {
                uint32_t b0, b1, b2, b3;
                const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = np * 16 + (lane >> 4) * 8;
                ldsm_x4_t(smem_u32(vb + r * LDS + c), b0, b1, b2, b3);
                mma_bf16(o[2 * np], a0, a1, a2, a3, b0, b1);
                mma_bf16(o[2 * np + 1], a0, a1, a2, a3, b2, b3);
            }
        }
    }
    cp_async_wait<0>();
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const int g8 = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int vr = qb * BM + warp * 16 + g8 + half * 8;
        if (vr >= nrows) continue;
        const int t = vr / G, g = vr % G;
        const float l = half ? l1 : l0, m = half ? m1 : m0;
        const float inv = l > 0.f ? 1.f / l : 0.f;
        const int64_t rowid = ((int64_t)split * T + t) * Hq + hk * G + g;
        float* op = Opart + rowid * DH;
#pragma unroll
        for (int i = 0; i < NT; ++i)
            *reinterpret_cast<float2*>(op + i * 8 + tq * 2) = make_float2(o[i][half * 2] * inv, o[i][half * 2 + 1] * inv);
        if (tq == 0) LSE[rowid] = l > 0.f ? (m + log2f(l)) * 0.6931471805599453f : -INFINITY;
    }
}

// merge P partials: out[row, :] (+)= gate * sum_p exp(lse_p - L) O_p / sum_p exp(lse_p - L)   (gate: gemma.py:192)
// Partial p = (rank r = p / spr, split s = p % spr) lives at Opart + r*rank_stride_o + s*rows*DH and
// LSE + r*rank_stride_l + s*rows, so the buffer produced by one all-gather of every rank's flat
// [O | LSE] block can be merged in place (spr = splits per rank).
__global__ void __launch_bounds__(128)
xattn_merge_kernel(const float* __restrict__ Opart, const float* __restrict__ LSE, int P, int spr,
                   int64_t rank_stride_o, int64_t rank_stride_l, int rows, int DH, float gate,
                   int accumulate, float* __restrict__ out) {
    const int row = blockIdx.x;
    extern __shared__ float wts[];                 // [P] weights, then 4 floats of reduction scratch
    __shared__ float red[4];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // phase 1: every partial's LSE is read once (thread-parallel over p), block max
    float lmax = -INFINITY;
    for (int p = tid; p < P; p += 128) {
        const float l = LSE[(p / spr) * rank_stride_l + (int64_t)(p % spr) * rows + row];
        wts[p] = l;
        lmax = fmaxf(lmax, l);
    }
    lmax = warp_max(lmax);
    if (lane == 0) red[warp] = lmax;
    __syncthreads();
    const float L = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    // phase 2: weights + denominator
    float dsum = 0.f;
    for (int p = tid; p < P; p += 128) {
        const float l = wts[p];
        const float w = (l == -INFINITY) ? 0.f : __expf(l - L);
        wts[p] = w;
        dsum += w;
    }
    dsum = warp_sum(dsum);
    if (lane == 0) red[warp] = dsum;
    __syncthreads();
    const float denom = red[0] + red[1] + red[2] + red[3];
    const float invd = denom > 0.f ? gate / denom : 0.f;
    // phase 3: weighted sum of the partial rows, float2 per thread (DH = 256 -> one pass, DH = 128 -> half the threads)
    for (int c = tid * 2; c < DH; c += 256) {
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll 4
        for (int p = 0; p < P; ++p) {
            const float w = wts[p];
            const float2 v = *reinterpret_cast<const float2*>(Opart + (p / spr) * rank_stride_o + ((int64_t)(p % spr) * rows + row) * DH + c);
            acc.x += w * v.x; acc.y += w * v.y;
        }
        float2* o = reinterpret_cast<float2*>(out + (int64_t)row * DH + c);
        float2 prev = accumulate ? *o : make_float2(0.f, 0.f);
        *o = make_float2(prev.x + acc.x * invd, prev.y + acc.y * invd);
    }
}

// =================================================================================================
// text stream: RoPE + causal attention (Tq small).  q position of row i is pos0 + i.
// =================================================================================================
__global__ void rope_kernel(__nv_bfloat16* __restrict__ x, int64_t ld, int col_off, int heads, int DH,
                            const float* __restrict__ inv_freq, int pos0) {
    const int t = blockIdx.x, h = blockIdx.y;
    __nv_bfloat16* p = x + (int64_t)t * ld + col_off + h * DH;
    const int half = DH / 2;
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
        const float ang = (float)(pos0 + t) * inv_freq[i];
        float sn, cs;
        sincosf(ang, &sn, &cs);
        // HF computes cos/sin in fp32 and casts them to the activation dtype before the rotation
        cs = __bfloat162float(__float2bfloat16(cs)); sn = __bfloat162float(__float2bfloat16(sn));
        const float a = __bfloat162float(p[i]), b = __bfloat162float(p[i + half]);
        p[i] = __float2bfloat16(a * cs - b * sn);
        p[i + half] = __float2bfloat16(b * cs + a * sn);
    }
}

// one CTA per (query row, head); 4 warps stride over keys; fp32 out [Tq, Hq*DH] (overwrites)
template <int DH>
__global__ void __launch_bounds__(128)
attn_text_kernel(const __nv_bfloat16* __restrict__ Q, int64_t ldq, const __nv_bfloat16* __restrict__ K,
                 const __nv_bfloat16* __restrict__ V, int64_t ldkv, int Tk, int pos0, int G, float scale, float softcap,
                 int window, float* __restrict__ out, int Hq) {
    constexpr int PER = DH / 32;
    const int i = blockIdx.x, h = blockIdx.y, hk = h / G;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qpos = pos0 + i;
    float q[PER];
#pragma unroll
    for (int e = 0; e < PER; ++e) q[e] = __bfloat162float(Q[(int64_t)i * ldq + h * DH + lane + 32 * e]);
    float m = -INFINITY, l = 0.f, acc[PER];
#pragma unroll
    for (int e = 0; e < PER; ++e) acc[e] = 0.f;
    int jlo = 0;
    if (window > 0) jlo = max(0, qpos - window + 1);
    const int jhi = min(Tk - 1, qpos);
    for (int j = jlo + warp; j <= jhi; j += 4) {
        const __nv_bfloat16* kr = K + (int64_t)j * ldkv + hk * DH;
        float d = 0.f;
#pragma unroll
        for (int e = 0; e < PER; ++e) d += q[e] * __bfloat162float(kr[lane + 32 * e]);
        d = warp_sum(d) * scale;
        if (softcap > 0.f) d = softcap * tanhf(d / softcap);
        const float mn = fmaxf(m, d);
        const float c = __expf(m - mn), pj = __expf(d - mn);
        const __nv_bfloat16* vr = V + (int64_t)j * ldkv + hk * DH;
#pragma unroll
        for (int e = 0; e < PER; ++e) acc[e] = acc[e] * c + pj * __bfloat162float(vr[lane + 32 * e]);
        l = l * c + pj;
        m = mn;
    }
    __shared__ float sm[4], sl[4], sacc[4][DH];
    if (lane == 0) { sm[warp] = m; sl[warp] = l; }
#pragma unroll
    for (int e = 0; e < PER; ++e) sacc[warp][lane + 32 * e] = acc[e];
    __syncthreads();
    float M = fmaxf(fmaxf(sm[0], sm[1]), fmaxf(sm[2], sm[3]));
    float den = 0.f;
    float wgt[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) { wgt[w] = (sm[w] == -INFINITY) ? 0.f : __expf(sm[w] - M); den += wgt[w] * sl[w]; }
    for (int c = threadIdx.x; c < DH; c += 128) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) v += wgt[w] * sacc[w][c];
        out[(int64_t)i * Hq * DH + h * DH + c] = den > 0.f ? v / den : 0.f;
    }
}

// fused text q/k preparation (one launch instead of clone + rope + copy + rope): from qkv [Tq, qd + 2*kd]
//   q_rope [Tq, qd]          = RoPE(q)                      (cross attention keeps using the un-roped q inside qkv)
//   kv_out [Tq, 2*kd] rows   = RoPE(k) | v                  (written straight into the text K||V cache rows pos0..)
__global__ void text_qk_prep_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t ld, __nv_bfloat16* __restrict__ q_rope,
                                    int64_t ldq, __nv_bfloat16* __restrict__ kv_out, int64_t ldkv, int Hq, int Hkv, int DH,
                                    const float* __restrict__ inv_freq, int pos0) {
    const int t = blockIdx.x, hh = blockIdx.y;
    const int half = DH / 2;
    const int qd = Hq * DH, kd = Hkv * DH;
    const __nv_bfloat16* src;
    __nv_bfloat16* dst;
    bool rope = true;
    if (hh < Hq) { src = qkv + (int64_t)t * ld + hh * DH; dst = q_rope + (int64_t)t * ldq + hh * DH; }
    else if (hh < Hq + Hkv) { const int h = hh - Hq; src = qkv + (int64_t)t * ld + qd + h * DH; dst = kv_out + (int64_t)t * ldkv + h * DH; }
    else { const int h = hh - Hq - Hkv; src = qkv + (int64_t)t * ld + qd + kd + h * DH; dst = kv_out + (int64_t)t * ldkv + kd + h * DH; rope = false; }
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
        const float a = __bfloat162float(src[i]), b = __bfloat162float(src[i + half]);
        if (rope) {
            float sn, cs;
            sincosf((float)(pos0 + t) * inv_freq[i], &sn, &cs);
            cs = __bfloat162float(__float2bfloat16(cs)); sn = __bfloat162float(__float2bfloat16(sn));
            dst[i] = __float2bfloat16(a * cs - b * sn);
            dst[i + half] = __float2bfloat16(b * cs + a * sn);
        } else {
            dst[i] = src[i];
            dst[i + half] = src[i + half];
        }
    }
}

// fused merge of up to two streams' partials + text attention + bf16 cast (one launch instead of merge x2 + cast):
//   out_bf16[row, :] = bf16( att_text[row, :] + sum_s gate_s * merge(partials_s)[row, :] )
struct MergeSrc {
    const float* O; const float* L; int P; int spr; int64_t rso; int64_t rsl; float gate;
};
__global__ void __launch_bounds__(128)
xattn_merge2_kernel(MergeSrc s0, MergeSrc s1, int nsrc, const float* __restrict__ att, int rows, int DH,
                    __nv_bfloat16* __restrict__ out, const unsigned int* flags, int nflags, unsigned int seq, int* err) {
    const int row = blockIdx.x;
    extern __shared__ float wts2[];               // [P0 + P1]
    __shared__ float red[4];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (flags != nullptr) {
        // receive side of the peer-memory exchange (xchg.cu): wait until every rank's partial of this exchange has landed in
        // this GPU's buffer.  Flags carry a monotonically increasing sequence number; a bounded spin turns a dead peer into an
        // error flag instead of a hung GPU.
        if (tid < nflags) {
            const long long t0 = clock64();
            while (true) {
                unsigned int v;
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + tid) : "memory");
                if ((int)(v - seq) >= 0) break;
                if (clock64() - t0 > 6000000000ll) { atomicExch(err, 1); break; }
            }
        }
        __syncthreads();
    }
    float inv[2] = {0.f, 0.f};
    int base = 0;
    for (int si = 0; si < nsrc; ++si) {
        const MergeSrc& s = si == 0 ? s0 : s1;
        float* w = wts2 + base;
        float lmax = -INFINITY;
        for (int p = tid; p < s.P; p += 128) {
            const float l = __ldcg(s.L + (p / s.spr) * s.rsl + (int64_t)(p % s.spr) * rows + row);
            w[p] = l; lmax = fmaxf(lmax, l);
        }
        lmax = warp_max(lmax);
        if (lane == 0) red[warp] = lmax;
        __syncthreads();
        const float Lm = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
        __syncthreads();
        float dsum = 0.f;
        for (int p = tid; p < s.P; p += 128) {
            const float l = w[p];
            const float e = (l == -INFINITY) ? 0.f : __expf(l - Lm);
            w[p] = e; dsum += e;
        }
        dsum = warp_sum(dsum);
        if (lane == 0) red[warp] = dsum;
        __syncthreads();
        const float den = red[0] + red[1] + red[2] + red[3];
        inv[si] = den > 0.f ? s.gate / den : 0.f;
        __syncthreads();
        base += s.P;
    }
    for (int c = tid * 2; c < DH; c += 256) {
        float2 acc = *reinterpret_cast<const float2*>(att + (int64_t)row * DH + c);
        base = 0;
        for (int si = 0; si < nsrc; ++si) {
            const MergeSrc& s = si == 0 ? s0 : s1;
            const float* w = wts2 + base;
            float2 a = make_float2(0.f, 0.f);
#pragma unroll 4
            for (int p = 0; p < s.P; ++p) {
                const float2 v = __ldcg(reinterpret_cast<const float2*>(s.O + (p / s.spr) * s.rso + ((int64_t)(p % s.spr) * rows + row) * DH + c));
                a.x += w[p] * v.x; a.y += w[p] * v.y;
            }
            acc.x += a.x * inv[si]; acc.y += a.y * inv[si];
            base += s.P;
        }
        *reinterpret_cast<uint32_t*>(out + (int64_t)row * DH + c) = pack_bf16(acc.x, acc.y);
    }
}

// =================================================================================================
// host launchers
// =================================================================================================
template <int DH, int DHP>
static int launch_dense(const void* qkv, int64_t ld, int q_off, int k_off, int v_off, void* out, int64_t ldo, int B, int S,
                        int H, float scale, cudaStream_t st) {
    using Cfg = DenseCfg<DH, DHP>;
    VB_SET_SMEM_ONCE(Cfg::kSmem, attn_dense_kernel<DH, DHP>);
    dim3 grid((S + Cfg::BM - 1) / Cfg::BM, H, B);
    attn_dense_kernel<DH, DHP><<<grid, 128, Cfg::kSmem, st>>>((const __nv_bfloat16*)qkv, ld, q_off, k_off, v_off,
                                                              (__nv_bfloat16*)out, ldo, S, H, scale * kLog2e);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int attn_dense(const void* qkv, int64_t ld, int q_off, int k_off, int v_off, void* out, int64_t ldo, int B, int S, int H,
               int dh, float scale, cudaStream_t st) {
    VB_REQUIRE(ld % 8 == 0 && ldo % 8 == 0 && q_off % 8 == 0 && k_off % 8 == 0 && v_off % 8 == 0, "attn_dense: alignment");
    if (B == 0 || S == 0) return 0;
    if (dh == 72) return launch_dense<72, 80>(qkv, ld, q_off, k_off, v_off, out, ldo, B, S, H, scale, st);
    if (dh == 64) return launch_dense<64, 64>(qkv, ld, q_off, k_off, v_off, out, ldo, B, S, H, scale, st);
    if (dh == 128) return launch_dense<128, 128>(qkv, ld, q_off, k_off, v_off, out, ldo, B, S, H, scale, st);
    VB_REQUIRE(false, "attn_dense: unsupported head_dim %d (have 64, 72, 128)", dh);
}

template <int DH>
static int launch_xattn(const void* Q, int64_t ldq, const void* K, const void* V, int64_t ldkv, const uint8_t* kmask, int T,
                        int N, int Hq, int Hkv, int splits, int keys_per_split, float scale, float softcap, float* Opart,
                        float* LSE, cudaStream_t st) {
    using Cfg = XCfg<DH>;
    VB_SET_SMEM_ONCE(Cfg::kSmem, xattn_splitkv_kernel<DH>);
    const int G = Hq / Hkv;
    dim3 grid(splits, Hkv, (T * G + Cfg::BM - 1) / Cfg::BM);
    xattn_splitkv_kernel<DH><<<grid, 128, Cfg::kSmem, st>>>((const __nv_bfloat16*)Q, ldq, (const __nv_bfloat16*)K,
                                                            (const __nv_bfloat16*)V, ldkv, kmask, T, N, Hq, G,
                                                            keys_per_split, scale, softcap, Opart, LSE);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int xattn_splitkv_sm100(const void*, int64_t, const void*, const void*, int64_t, const uint8_t*, int, int, int, int, int, int,
                        float, float, float*, float*, cudaStream_t);
bool xattn_sm100_supports(int dh, float softcap);

int xattn_splitkv(const void* Q, int64_t ldq, const void* K, const void* V, int64_t ldkv, const uint8_t* kmask, int T, int N,
                  int Hq, int Hkv, int dh, int splits, float scale, float softcap, float* Opart, float* LSE, int force_mma,
                  cudaStream_t st) {
    VB_REQUIRE(T > 0 && N >= 0 && splits > 0 && Hq % Hkv == 0, "xattn_splitkv: bad shape T=%d N=%d splits=%d", T, N, splits);
    VB_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0, "xattn_splitkv: alignment");
    int kps = (N + splits - 1) / splits;
    kps = ((kps + 63) / 64) * 64;
    if (kps == 0) kps = 64;
    // soft-capped dh=256 (Gemma2) and un-capped dh=128 (Mistral) -> tcgen05/TMEM/TMA streaming kernel; otherwise the warp-level kernel
    const int G = Hq / Hkv;
    if (!force_mma && xattn_sm100_supports(dh, softcap) && 128 % G == 0 &&
        (reinterpret_cast<uintptr_t>(Q) & 15) == 0 && (reinterpret_cast<uintptr_t>(K) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(V) & 15) == 0 && (kmask == nullptr || (reinterpret_cast<uintptr_t>(kmask) & 15) == 0))
        return xattn_splitkv_sm100(Q, ldq, K, V, ldkv, kmask, T, N, Hq, Hkv, dh, splits, scale, softcap, Opart, LSE, st);
    if (dh == 256) return launch_xattn<256>(Q, ldq, K, V, ldkv, kmask, T, N, Hq, Hkv, splits, kps, scale, softcap, Opart, LSE, st);
    if (dh == 128) return launch_xattn<128>(Q, ldq, K, V, ldkv, kmask, T, N, Hq, Hkv, splits, kps, scale, softcap, Opart, LSE, st);
    VB_REQUIRE(false, "xattn_splitkv: unsupported head_dim %d (have 128, 256)", dh);
}

int xattn_splitkv_sm100_seg(const void*, int64_t, const void*, const void*, int64_t, int, int, const int*, const int*, const int*,
                            const uint8_t* const*, int, int, int, int, float, float, float*, float*, cudaStream_t);

// Both streams of a layer (image rows, audio rows of the same K||V cache) in one call -- and, on the tcgen05 path, ONE launch whose
// grid covers the splits of both segments (no second prologue / tail per layer).  K / V point at cache row 0 of the layer.
// Opart [splits[0] + splits[1]][T][Hq][dh], LSE [splits[0] + splits[1]][T][Hq]; segment 1's partials follow segment 0's.
int xattn_splitkv_seg(const void* Q, int64_t ldq, const void* K, const void* V, int64_t ldkv, int n_rows_total, int nseg, const int* row0,
                      const int* rows, const int* splits, const uint8_t* const* masks, int T, int Hq, int Hkv, int dh, float scale,
                      float softcap, float* Opart, float* LSE, int force_mma, int* launches, cudaStream_t st) {
    VB_REQUIRE(T > 0 && nseg >= 1 && nseg <= 2 && Hq % Hkv == 0, "xattn_splitkv_seg: bad shape T=%d nseg=%d", T, nseg);
    VB_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0, "xattn_splitkv_seg: alignment");
    const int G = Hq / Hkv;
    bool fast = !force_mma && xattn_sm100_supports(dh, softcap) && 128 % G == 0 &&
                (reinterpret_cast<uintptr_t>(Q) & 15) == 0 && (reinterpret_cast<uintptr_t>(K) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(V) & 15) == 0;
    for (int i = 0; i < nseg; ++i)
        if (masks && masks[i] && (reinterpret_cast<uintptr_t>(masks[i]) & 15) != 0) fast = false;
    if (fast) {
        if (launches) *launches = 1;
        return xattn_splitkv_sm100_seg(Q, ldq, K, V, ldkv, n_rows_total, nseg, row0, rows, splits, masks, T, Hq, Hkv, dh, scale, softcap,
                                       Opart, LSE, st);
    }
    int64_t po = 0;
    for (int i = 0; i < nseg; ++i) {        // warp-level kernel: one launch per segment, same output layout
        const __nv_bfloat16* kk = reinterpret_cast<const __nv_bfloat16*>(K) + (int64_t)row0[i] * ldkv;
        const __nv_bfloat16* vv = reinterpret_cast<const __nv_bfloat16*>(V) + (int64_t)row0[i] * ldkv;
        int rc = xattn_splitkv(Q, ldq, kk, vv, ldkv, masks ? masks[i] : nullptr, T, rows[i], Hq, Hkv, dh, splits[i], scale, softcap,
                               Opart + po * T * Hq * dh, LSE + po * T * Hq, 1, st);
        if (rc) return rc;
        po += splits[i];
    }
    if (launches) *launches = nseg;
    return 0;
}

int xattn_merge(const float* Opart, const float* LSE, int P, int spr, int64_t rank_stride_o, int64_t rank_stride_l, int rows,
                int dh, float gate, int accumulate, float* out, cudaStream_t st) {
    if (rows == 0) return 0;
    VB_REQUIRE(P > 0 && spr > 0 && P % spr == 0, "xattn_merge: P=%d must be a multiple of splits-per-rank %d", P, spr);
    xattn_merge_kernel<<<rows, 128, P * sizeof(float), st>>>(Opart, LSE, P, spr, rank_stride_o, rank_stride_l, rows, dh, gate,
                                                          accumulate, out);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int rope_inplace(void* x, int64_t ld, int col_off, int T, int heads, int dh, const float* inv_freq, int pos0, cudaStream_t st) {
    if (T == 0) return 0;
    rope_kernel<<<dim3(T, heads), 64, 0, st>>>((__nv_bfloat16*)x, ld, col_off, heads, dh, inv_freq, pos0);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int attn_text(const void* Q, int64_t ldq, const void* K, const void* V, int64_t ldkv, int Tq, int Tk, int pos0, int Hq,
              int Hkv, int dh, float scale, float softcap, int window, float* out, cudaStream_t st) {
    if (Tq == 0) return 0;
    const int G = Hq / Hkv;
    dim3 grid(Tq, Hq);
    if (dh == 256)
        attn_text_kernel<256><<<grid, 128, 0, st>>>((const __nv_bfloat16*)Q, ldq, (const __nv_bfloat16*)K,
                                                    (const __nv_bfloat16*)V, ldkv, Tk, pos0, G, scale, softcap, window, out, Hq);
    else if (dh == 128)
        attn_text_kernel<128><<<grid, 128, 0, st>>>((const __nv_bfloat16*)Q, ldq, (const __nv_bfloat16*)K,
                                                    (const __nv_bfloat16*)V, ldkv, Tk, pos0, G, scale, softcap, window, out, Hq);
    else
        VB_REQUIRE(false, "attn_text: unsupported head_dim %d", dh);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}


int text_qk_prep(const void* qkv, int64_t ld, void* q_rope, int64_t ldq, void* kv_out, int64_t ldkv, int Tq, int Hq, int Hkv,
                 int dh, const float* inv_freq, int pos0, cudaStream_t st) {
    if (Tq == 0) return 0;
    text_qk_prep_kernel<<<dim3(Tq, Hq + 2 * Hkv), 64, 0, st>>>((const __nv_bfloat16*)qkv, ld, (__nv_bfloat16*)q_rope, ldq,
                                                              (__nv_bfloat16*)kv_out, ldkv, Hq, Hkv, dh, inv_freq, pos0);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int xattn_merge2(const float* O0, const float* L0, int P0, int spr0, int64_t rso0, int64_t rsl0, float gate0, const float* O1,
                 const float* L1, int P1, int spr1, int64_t rso1, int64_t rsl1, float gate1, int nsrc, const float* att,
                 int rows, int dh, void* out_bf16, const unsigned int* flags, int nflags, unsigned int seq, int* err,
                 cudaStream_t st) {
    if (rows == 0) return 0;
    VB_REQUIRE(nsrc >= 0 && nsrc <= 2 && dh % 2 == 0, "xattn_merge2: nsrc=%d dh=%d", nsrc, dh);
    VB_REQUIRE(flags == nullptr || (nflags >= 1 && nflags <= 128 && err != nullptr), "xattn_merge2: bad flag arguments (nflags=%d)", nflags);
    MergeSrc s0{O0, L0, nsrc > 0 ? P0 : 0, spr0 > 0 ? spr0 : 1, rso0, rsl0, gate0};
    MergeSrc s1{O1, L1, nsrc > 1 ? P1 : 0, spr1 > 0 ? spr1 : 1, rso1, rsl1, gate1};
    xattn_merge2_kernel<<<rows, 128, (s0.P + s1.P + 1) * sizeof(float), st>>>(s0, s1, nsrc, att, rows, dh, (__nv_bfloat16*)out_bf16,
                                                                              flags, nflags, seq, err);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace vb
