// 2-CTA (cta_group::2) variant of the persistent bf16 GEMM: a CTA pair on one TPC computes a 256 x BLOCK_N tile.
//
//   * each CTA TMA-loads its own 128 rows of A and HALF of the W tile (BLOCK_N/2 rows); the pair's tcgen05.mma
//     (M=256, issued by the leader CTA only) reads A from both CTAs' shared memory and the two W halves across the
//     pair, so per-CTA L2->SMEM traffic per flop drops by a third versus the 1-CTA 128 x 256 tile and the
//     stage gets small enough for a 6-deep ring;
//   * all TMA transactions of both CTAs complete on the LEADER's full barrier (peer bit cleared in the mbarrier
//     address); tcgen05.commit multicasts the "stage free" / "accumulator ready" arrivals to both CTAs;
//   * each CTA drains its own 128 TMEM lanes with the shared epilogue; the follower's epilogue threads release the
//     accumulator on the leader's barrier through a shared::cluster arrive.
#include <cstdlib>

#include "gemm_epilogue.cuh"

namespace vb {
namespace g2 {

constexpr int BLOCK_M = 128;        // per CTA; the pair covers 256 rows
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 384;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> leader CTA

template <int BLOCK_N, bool TS = false, bool RT = false>
struct Cfg {
    static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;            // 16 KB
    static constexpr int kBBytes = (BLOCK_N / 2) * BLOCK_K * 2;      // this CTA's half of W
    static constexpr int kStageBytes = kABytes + kBBytes;
    // TMA-store epilogue: one [32 rows x 64 cols] bf16 box per epilogue warp; RT (residual by TMA): one box per 64 columns of the warp's half
    static constexpr int kOutBytes = RT ? 8 * (BLOCK_N / 128) * 4096 : TS ? 8 * 4096 : 0;
    static constexpr int kRing = RT ? (BLOCK_N == 256 ? 160 : 176) * 1024 : TS ? 176 * 1024 : 200 * 1024;
    static constexpr int kStages = kRing / kStageBytes > 8 ? 8 : kRing / kStageBytes;
    static constexpr int kTmemCols = 512;
    static constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + 1024 + 512 + (RT ? 128 : 0);
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(leader_bar) & kPeerMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :
        : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once) on the same barrier offset in both CTAs of the pair when all prior MMAs of this thread are done
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}
// the accumulator hand-back only has to order the tcgen05.ld's (already complete after tcgen05.wait::ld + the before_thread_sync fence);
// a release at cluster scope would also wait for this thread's outstanding global stores (shows up as MEMBAR stalls in ncu)
__device__ __forceinline__ void mbar_arrive_leader_relaxed(uint64_t* bar) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int group_m, int& m_blk, int& n_blk) {
    const int per_group = group_m * num_n;
    const int g = tile / per_group;
    const int first_m = g * group_m;
    const int gm = min(group_m, num_m - first_m);
    const int r = tile - g * per_group;
    m_blk = first_m + r % gm;
    n_blk = r / gm;
}

template <int BLOCK_N, bool LN, bool TS = false, bool RT = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_r, const GemmParams p) {
    using C = Cfg<BLOCK_N, TS, RT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + C::kStages * C::kABytes;
    uint8_t* smem_out = smem + C::kStages * C::kStageBytes;          // TS: 8 x 4 KB staging boxes (1024-aligned)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes + C::kOutBytes);
    uint64_t* full_bar = bars;                     // [kStages]  used on the leader only (TMA of both CTAs -> leader MMA)
    uint64_t* empty_bar = bars + C::kStages;       // [kStages]  per CTA (MMA commit multicast -> each producer)
    uint64_t* tmem_full = bars + 2 * C::kStages;   // [2]        per CTA (commit multicast -> each epilogue)
    uint64_t* tmem_empty = tmem_full + 2;          // [2]        used on the leader only (both epilogues -> MMA)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    uint64_t* res_bar = tmem_empty + 4;            // [8 warps][2 boxes]  RT: residual boxes landed (per CTA)

    const int warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
    const int num_m = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
    const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
    const int num_tiles = num_m * num_n;
    const int num_k = (p.K + BLOCK_K - 1) / BLOCK_K;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < C::kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 2 * 256);
        }
        if (RT)
            for (int i = 0; i < 16; ++i) mbar_init(&res_bar[i], 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_2sm<C::kTmemCols>(tmem_ptr);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                int m_blk, n_blk;
                tile_coords(tile, num_m, num_n, p.group_m, m_blk, n_blk);
                const int row_a = m_blk * 2 * BLOCK_M + (int)rank * BLOCK_M;
                // ragged last column tile: the MMA is issued only n_eff = roundup(N - n0, 16) wide, each CTA supplies n_eff / 2 rows of W
                const int n_eff = p.ragged_tail ? min(BLOCK_N, ((p.N - n_blk * BLOCK_N + 15) >> 4) << 4) : BLOCK_N;
                const int row_b = n_blk * BLOCK_N + (int)rank * (n_eff / 2);
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (leader) mbar_expect_tx(&full_bar[stage], 2 * C::kStageBytes);
                    tma_load_2d_2sm(smem_a + stage * C::kABytes, &tmap_a, &full_bar[stage], kb * BLOCK_K, row_a);
                    tma_load_2d_2sm(smem_b + stage * C::kBBytes, &tmap_b, &full_bar[stage], kb * BLOCK_K, row_b);
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                int m_blk_, n_blk_;
                tile_coords(tile, num_m, num_n, p.group_m, m_blk_, n_blk_);
                // N = 1152 / 3456 / 4304 are not multiples of 256: the last column tile runs a narrower MMA instead of multiplying zeros
                const int n_eff = p.ragged_tail ? min(BLOCK_N, ((p.N - n_blk_ * BLOCK_N + 15) >> 4) << 4) : BLOCK_N;
                const uint32_t idesc = umma_idesc_bf16(2 * BLOCK_M, (uint32_t)n_eff);
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t a_desc = umma_desc_k_sw128(smem_u32(smem_a + stage * C::kABytes));
                    const uint64_t b_desc = umma_desc_k_sw128(smem_u32(smem_b + stage * C::kBBytes));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                        umma_f16_2sm(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    umma_commit_2sm(&empty_bar[stage]);
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit_2sm(&tmem_full[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (both CTAs, own 128 rows) =====================
        const int ew = (warp - 4) & 3;
        const int wg = (warp - 4) >> 2;
        const int row_in_tile = (int)rank * BLOCK_M + ew * 32 + lane_id();
        int acc = 0;
        uint32_t acc_phase = 0, rphase = 0;
        constexpr int NBOX = BLOCK_N / 128;            // 64-column boxes per warp (its column half)
        uint8_t* my_stage = smem_out + (warp - 4) * (RT ? NBOX * 4096 : 4096);
        uint64_t* my_rbar = res_bar + (warp - 4) * 2;
        for (int tile = pair; tile < num_tiles; tile += num_pairs) {
            int m_blk, n_blk;
            tile_coords(tile, num_m, num_n, p.group_m, m_blk, n_blk);
            const int row_warp0 = m_blk * 2 * BLOCK_M + (int)rank * BLOCK_M + ew * 32;
            if (RT) {
                // the residual boxes of this tile do not depend on the accumulator: fetch them while the MMAs still run
                if (lane_id() == 0) {
                    tma_store_wait_read0();                       // this warp's earlier stores have left its buffers
#pragma unroll
                    for (int b = 0; b < NBOX; ++b) {
                        mbar_expect_tx(&my_rbar[b], 4096);
                        tma_load_2d(my_stage + b * 4096, &tmap_r, &my_rbar[b], n_blk * BLOCK_N + wg * (BLOCK_N / 2) + b * 64, row_warp0,
                                    kEvictFirst);
                    }
                }
                __syncwarp();
            }
            mbar_wait_warp(&tmem_full[acc], acc_phase);          // one polling lane per warp
            tc_fence_after();
            const int row = m_blk * 2 * BLOCK_M + row_in_tile;
            const uint32_t taddr = tmem_base + acc * BLOCK_N + ((uint32_t)(ew * 32) << 16);
            epilogue_tile<BLOCK_N, LN, TS, RT>(p, taddr, row, n_blk, wg, &tmap_c, my_stage, row_warp0, my_rbar, rphase);
            rphase ^= 1;
            tc_fence_before();
            if (p.relaxed_arrive) mbar_arrive_leader_relaxed(&tmem_empty[acc]);
            else mbar_arrive_leader(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (TS && lane_id() == 0) tma_store_wait_read0();             // the last boxes must have left shared memory before the CTA exits
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2sm<C::kTmemCols>(tmem_base);
    }
}

template <int BLOCK_N, bool LN, bool TS, bool RT = false>
static int launch_ts(const void* A, int64_t lda, const void* W, int64_t ldw, const GemmParams& p, cudaStream_t st) {
    using C = Cfg<BLOCK_N, TS, RT>;
    CUtensorMap ta, tb, tc, tr;
    int rc;
    if ((rc = make_tmap_2d_bf16(&ta, A, (uint64_t)p.K, (uint64_t)p.M, (uint64_t)lda * 2, BLOCK_K, BLOCK_M))) return rc;
    if ((rc = make_tmap_2d_bf16(&tb, W, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)ldw * 2, BLOCK_K, BLOCK_N / 2))) return rc;
    tc = ta;
    if (TS && (rc = make_tmap_2d_bf16(&tc, p.C, (uint64_t)(p.glu ? p.N / 2 : p.N), (uint64_t)p.M, (uint64_t)p.ldc * 2, 64, 32))) return rc;
    tr = tc;
    if (RT && (rc = make_tmap_2d_bf16(&tr, p.residual, (uint64_t)p.N, (uint64_t)p.M, (uint64_t)p.ldr * 2, 64, 32))) return rc;
    VB_SET_SMEM_ONCE(C::kSmemBytes, gemm2_bf16_kernel<BLOCK_N, LN, TS, RT>);
    const int num_m = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M), num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
    const int tiles = num_m * num_n;
    const int max_pairs = num_sms() / 2;
    const int pairs = tiles < max_pairs ? tiles : max_pairs;
    gemm2_bf16_kernel<BLOCK_N, LN, TS, RT><<<2 * pairs, kNumThreads, C::kSmemBytes, st>>>(ta, tb, tc, tr, p);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// TMA-store epilogue: plain bf16 outputs (no GLU, no fp32) of the 256- and 128-wide tiles (a warp's column half is then a whole number
// of 64-column boxes), 16-byte aligned rows.  Sustained tower block 1 084 -> 1 224 TF/s, C3 step 4 596 -> 4 365 ms
// (profiles/r02_ab_tma_store.txt); VIDI_GEMM2_TMASTORE=0 restores the per-thread 16-byte stores for A/B.
template <int BLOCK_N, bool LN>
static int launch(const void* A, int64_t lda, const void* W, int64_t ldw, const GemmParams& p, cudaStream_t st) {
    static const int ts = getenv("VIDI_GEMM2_TMASTORE") ? atoi(getenv("VIDI_GEMM2_TMASTORE")) : 1;
    constexpr bool kCanTS = !LN && (BLOCK_N == 256 || BLOCK_N == 128);
    if (kCanTS && ts && (!p.glu || BLOCK_N == 256) && !p.out_fp32 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && p.ldc % 8 == 0) {
        // residual rows by TMA too (row-aligned residual only; the position-embedding add with res_mod keeps the register path)
        static const int rt = getenv("VIDI_GEMM2_RESTMA") ? atoi(getenv("VIDI_GEMM2_RESTMA")) : 1;
        if (rt && !p.glu && p.residual && p.res_mod == 0 && p.ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0)
            return launch_ts<BLOCK_N, LN, kCanTS, kCanTS>(A, lda, W, ldw, p, st);
        return launch_ts<BLOCK_N, LN, kCanTS>(A, lda, W, ldw, p, st);
    }
    return launch_ts<BLOCK_N, LN, false>(A, lda, W, ldw, p, st);
}

}  // namespace g2

// same contract as gemm_bf16 (gemm_sm100.cu); block_n in {128, 256}; GLU tiles: each CTA's W half must hold whole
// [gate | up] groups, so the packed group width is block_n/2 (see weights.pack_glu) -- handled by the caller.
int gemm2_bf16_ln(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K,
                  const float* bias, const void* residual, int64_t ldr, int res_mod, int act, float act_param, int out_fp32,
                  int glu, int block_n, const float* ln_stats, int ln_parts, const float* ln_colsum, float ln_eps,
                  float* stats_out, cudaStream_t st) {
    VB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm2_bf16: empty problem");
    VB_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0, "gemm2_bf16: K/lda/ldw/ldc must be multiples of 8");
    if (glu) VB_REQUIRE(N % block_n == 0, "gemm2_bf16: GLU needs N %% block_n == 0 (packed gate|up tiles)");
    g2::GemmParams p;
    p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc; p.bias = bias;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(residual); p.ldr = ldr; p.res_mod = res_mod;
    p.act = act; p.act_param = act_param; p.out_fp32 = out_fp32; p.glu = glu;
    // default: relaxed hand-back (+4.7 % on the tower block, profiles/r02_tower_ab.txt); VIDI_GEMM2_RELAXED=0 restores the release for A/B
    static const int relaxed = getenv("VIDI_GEMM2_RELAXED") ? atoi(getenv("VIDI_GEMM2_RELAXED")) : 1;
    p.relaxed_arrive = relaxed;
    static const int ragged = getenv("VIDI_GEMM2_RAGGED") ? atoi(getenv("VIDI_GEMM2_RAGGED")) : 1;
    p.ragged_tail = ragged;
    if (ln_stats || stats_out) VB_REQUIRE(!glu && !out_fp32, "gemm2_bf16_ln: LayerNorm fold / row statistics need a plain bf16 output");
    if (ln_stats) {
        VB_REQUIRE(ln_parts > 0 && ln_colsum != nullptr, "gemm2_bf16_ln: ln_stats needs ln_parts > 0 and ln_colsum");
        p.ln_stats = reinterpret_cast<const float2*>(ln_stats); p.ln_parts = ln_parts; p.ln_colsum = ln_colsum;
        p.ln_inv_k = 1.0f / (float)K; p.ln_eps = ln_eps;
    }
    if (stats_out) {
        p.stats_out = reinterpret_cast<float2*>(stats_out);
        p.stats_parts = 2 * ((N + block_n - 1) / block_n);
    }
    {   // pair-blocks (256 rows) per L2-resident group.  Measured in-step on C3 (profiles/r02_ab_l2_group.txt): 48 / 32 / 24 / 16 MB of A
        // rows per group -> 4 112 / 4 056 / 4 070 / 4 100 ms per step; the stream-pass sites (K >= 2048) are best at 32 MB (gate||up
        // 1 349 -> 1 388 TF/s, DRAM reads of that launch 26.7 -> ~20 GB), the K = 1152 / 1280 tower sites at 16 MB (1 286 -> 1 322 TF/s).
        // The A panel competes with W tiles, the output stream and -- the L2 being two die-local halves -- a second copy of itself.
        const int64_t per_block = (int64_t)2 * g2::BLOCK_M * K * 2;
        static const int l2_mb_env = getenv("VIDI_GEMM2_L2MB") ? atoi(getenv("VIDI_GEMM2_L2MB")) : 0;
        const int l2_mb = l2_mb_env > 0 ? l2_mb_env : (K < 2048 ? 16 : 32);
        int64_t gm = ((int64_t)l2_mb << 20) / per_block;
        p.group_m = (int)(gm < 4 ? 4 : gm > 32 ? 32 : gm);
    }
    if (ln_stats || stats_out) {          // separate instantiation: the plain kernel carries none of the LayerNorm-fold code
        if (block_n == 256) return g2::launch<256, true>(A, lda, W, ldw, p, st);
        if (block_n == 192) return g2::launch<192, true>(A, lda, W, ldw, p, st);
        if (block_n == 128) return g2::launch<128, true>(A, lda, W, ldw, p, st);
    } else {
        if (block_n == 256) return g2::launch<256, false>(A, lda, W, ldw, p, st);
        if (block_n == 192) return g2::launch<192, false>(A, lda, W, ldw, p, st);
        if (block_n == 128) return g2::launch<128, false>(A, lda, W, ldw, p, st);
    }
    VB_REQUIRE(false, "gemm2_bf16: unsupported block_n %d", block_n);
}

int gemm2_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K,
               const float* bias, const void* residual, int64_t ldr, int res_mod, int act, float act_param, int out_fp32,
               int glu, int block_n, cudaStream_t st) {
    return gemm2_bf16_ln(A, lda, W, ldw, C, ldc, M, N, K, bias, residual, ldr, res_mod, act, act_param, out_fp32, glu, block_n,
                         nullptr, 0, nullptr, 0.f, nullptr, st);
}

}  // namespace vb
