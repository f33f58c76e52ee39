// Tower attention v3 (SigLIP dh=72 / Whisper dh=64 encoder self-attention, reference: siglip / whisper encoder layers called
// from Vidi1.5_9B/vidi/model/multimodal_encoder/*).  One work item = (batch, head, PAIR of 128-query blocks); each softmax
// warpgroup owns one query block (thread = one query row).  What v2 (attn2_sm100.cu) measured and this kernel changes:
//   * v2 was bound by neither MUFU, issue slots nor the tensor pipe but by the hand-off chain
//     softmax(j) -> P V(j) -> Q K(j+1)^T -> softmax(j+1) (timing probes: removing the exp2, the P V MMAs or the S loads did not
//     change the time per tile).  Here S is DOUBLE-BUFFERED per query block with 64-key tiles: Q K(j+1)^T is issued two tiles
//     ahead, so a softmax warp finds its next S tile already complete and only the MMA issuer sees hand-off latency;
//   * the probabilities never touch shared memory: P = exp2(S*scale - ref) is written as packed bf16 back INTO THE TMEM
//     COLUMNS OF ITS S TILE (tcgen05.st) and the P V MMA takes its A operand from TMEM (tcgen05.mma [d], [a_tmem], b_desc);
//   * O accumulates in TMEM over the whole key loop (enable_input_d), together with the row sum l (all-ones V columns:
//     dh=72 uses the 8 zero-filled columns 72..79 of the V tail tile, rewritten to 1.0 by a helper warp; dh=64 issues one
//     extra N=16 MMA against a constant ones operand).  The softmax reference is exact after the first key tile and is then
//     only moved when some p would exceed 2^8 (p is computed pre-scaled by 2^-7, so the test is the exponent MSB of the
//     packed bf16 words, OR-ed together); only then is O rescaled in TMEM (ld, multiply, st) — no per-tile correction step.
//   TMEM: block A S/P buffers [0,64) [64,128), block B [128,192) [192,256), O_A [256,336), O_B [384,464).
// Q K(j+2)^T overwrites the buffer P(j) V(j) reads; it is issued after it by the same thread (the tensor pipe executes one
// thread's MMAs in issue order).
#include "common.cuh"

namespace vb {

constexpr float kLog2eC = 1.4426950408889634f;

template <int DH>
struct Fa3Cfg {
    static constexpr int BM = 128, BN = 64, KV = 6;
    static constexpr int TAIL = (DH > 64) ? 16 : 0;
    static constexpr int LCOL = DH;                            // TMEM column (within an O block) holding the row sum
    static constexpr int OSPAN = 80;                           // columns of an O block that carry data (O | l)
    static constexpr int kQMain = 128 * 64 * 2, kQTail = 128 * 16 * 2;     // Q block: [128][64] SW128 | [128][16] SW32
    static constexpr int kQBytes = kQMain + (TAIL ? kQTail : 0);
    static constexpr int kQSlot = 20 * 1024;
    static constexpr int kKMain = BN * 64 * 2, kKTail = BN * 16 * 2;       // K / V tile: [64][64] SW128 | [64][16] SW32
    static constexpr int kKBytes = kKMain + (TAIL ? kKTail : 0);
    static constexpr int kKSlot = 10 * 1024;
    // Q: [item parity 2][block 2] | K [KV] | V [KV] | ones | barriers
    static constexpr int kOffQ = 0, kOffK = 4 * kQSlot, kOffV = kOffK + KV * kKSlot;
    static constexpr int kOffOnes = kOffV + KV * kKSlot;
    static constexpr int kOffBar = kOffOnes + 512;
    static constexpr int kSmem = kOffBar + 512 + 1024;        // 200 KB + ...
};

struct Fa3Params {
    int S, H, B;
    int npair;               // query-block pairs per (b,h)
    int nkt;                 // key tiles
    int items;
    float scale_log2;
    __nv_bfloat16* out;
    int64_t ldo;
};

// D[tmem] (+)= A[tmem] * B[smem]: A is M x 16 bf16, row i in TMEM lane i, elements (2c, 2c+1) packed in 32-bit column c
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        :
        : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :
        : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        :
        : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

// 2^x on the FMA pipe (no MUFU): x = n + f with n = rint(x) by the 1.5 * 2^23 trick, a degree-3 polynomial for 2^f on [-0.5, 0.5]
// (max relative error 7.5e-5, 26x below the bf16 half-ulp of P — tools/exp2_poly.py) and n added to the exponent field.  Used for a
// share of the scores (kPolyMod below) so that the XU and the issue slots are loaded evenly; NOT ENABLED in the shipped
// instantiation (kPolyMod = 0) until it has been validated and timed on a B200.
__device__ __forceinline__ float exp2_poly3(float x) {
    const float xf = x + 12582912.f;                                   // 1.5 * 2^23: low mantissa bits now hold rint(x)
    const float f = x - (xf - 12582912.f);
    float q = fmaf(0.0551716387f, f, 0.242611125f);
    q = fmaf(q, f, 0.693260968f);
    q = fmaf(q, f, 0.999928057f);
    const int bits = __float_as_int(q) + (__float_as_int(xf) << 23);  // (n + 0x4B400000) << 23 == n << 23 mod 2^32
    return x < -126.f ? 0.f : __int_as_float(bits);                    // ragged / far-below-reference keys -> exactly 0
}

template <int DH, int kPolyMod>
__global__ void __launch_bounds__(384, 1)
attn_fwd3_sm100_kernel(const __grid_constant__ CUtensorMap tm_q_main, const __grid_constant__ CUtensorMap tm_q_tail,
                       const __grid_constant__ CUtensorMap tm_k_main, const __grid_constant__ CUtensorMap tm_k_tail,
                       const Fa3Params p) {
    using C = Fa3Cfg<DH>;
    constexpr int KV = C::KV;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
    uint64_t* q_full = bars;          // [2] per item parity
    uint64_t* q_empty = bars + 2;     // [2] (2 arrivals: one per MMA issuer)
    uint64_t* kv_full = bars + 4;     // [KV]
    uint64_t* kv_empty = bars + 12;   // [KV] (2 arrivals: one per MMA issuer)
    uint64_t* v_ones = bars + 20;     // [KV] V tail ones written (dh=72)
    uint64_t* s_full = bars + 28;     // [block 2][buffer 2]: S tile complete in TMEM
    uint64_t* p_full = bars + 32;     // [block 2][buffer 2] (128 arrivals): P written to TMEM, O rescaled if needed
    uint64_t* o_full = bars + 36;     // [block 2][tile parity 2]: P V of that tile accumulated (a waiter may skip phases, so one
                                      // barrier per parity keeps it at most one phase behind: S(t) complete => P V(t-2) complete)
    uint64_t* o_empty = bars + 40;    // [2] (128 arrivals) one phase per item: O read out
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 42);

    const int warp = threadIdx.x >> 5;
    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tm_q_main);
        tma_prefetch_desc(&tm_k_main);
        if (C::TAIL) { tma_prefetch_desc(&tm_q_tail); tma_prefetch_desc(&tm_k_tail); }
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 2);
            mbar_init(&o_empty[i], 128);
        }
        for (int i = 0; i < 4; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); mbar_init(&o_full[i], 1); }
        for (int i = 0; i < KV; ++i) {
            mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 2); mbar_init(&v_ones[i], 1);
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
                    mbar_expect_tx(&q_full[qi], 2 * C::kQBytes);
                    for (int blk = 0; blk < 2; ++blk) {
                        uint8_t* sq = smem + C::kOffQ + (qi * 2 + blk) * C::kQSlot;
                        const int q0 = (pr * 2 + blk) * C::BM;
                        tma_load_4d(sq, &tm_q_main, &q_full[qi], 0, h, q0, b, kEvictNormal);
                        if (C::TAIL) tma_load_4d(sq + C::kQMain, &tm_q_tail, &q_full[qi], 64, h, q0, b, kEvictNormal);
                    }
                    for (int j = 0; j < nkt; ++j, ++g) {
                        const int st = g % KV;
                        mbar_wait(&kv_empty[st], ((g / KV) & 1) ^ 1);
                        mbar_expect_tx(&kv_full[st], 2 * C::kKBytes);
                        uint8_t* sk = smem + C::kOffK + st * C::kKSlot;
                        uint8_t* sv = smem + C::kOffV + st * C::kKSlot;
                        tma_load_4d(sk, &tm_k_main, &kv_full[st], 0, p.H + h, j * C::BN, b, kEvictLast);
                        tma_load_4d(sv, &tm_k_main, &kv_full[st], 0, 2 * p.H + h, j * C::BN, b, kEvictLast);
                        if (C::TAIL) {
                            tma_load_4d(sk + C::kKMain, &tm_k_tail, &kv_full[st], 64, p.H + h, j * C::BN, b, kEvictLast);
                            tma_load_4d(sv + C::kKMain, &tm_k_tail, &kv_full[st], 64, 2 * p.H + h, j * C::BN, b, kEvictLast);
                        }
                    }
                }
            }
        } else if (warp == 1 || warp == 2) {
            // ============================ MMA issuers: warp 1 drives query block A, warp 2 block B ============================
            // (one issuer for both blocks couples them: a late softmax of A would hold back B's Q K^T prefetch behind A's P V)
            // The CTA's key tiles form one stream t = 0,1,2,... across its items; tile t of a block lives in S buffer t&1.
            // Q K^T runs two tiles ahead of P V:   QK(0) QK(1) | PV(0) QK(2) | PV(1) QK(3) | ...
            if (elect_one()) {
                const int blk = warp - 1;
                constexpr uint32_t idesc_qk = umma_idesc_bf16(128, C::BN);
                constexpr uint32_t idesc_pv_main = umma_idesc_bf16(128, 64, 0, 1);
                constexpr uint32_t idesc_pv_tail = umma_idesc_bf16(128, 16, 0, 1);
                const int n_items = (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
                const uint32_t T = (uint32_t)n_items * (uint32_t)nkt;
                struct Cursor { uint32_t t, it; int j; };
                auto advance = [&](Cursor& c) { ++c.t; if (++c.j == nkt) { c.j = 0; ++c.it; } };
                auto issue_qk = [&](const Cursor& c) {
                    const int st = c.t % KV;
                    const int qi = c.it & 1;
                    if (c.j == 0) mbar_wait(&q_full[qi], (c.it >> 1) & 1);
                    mbar_wait(&kv_full[st], (c.t / KV) & 1);
                    tc_fence_after();
                    const uint8_t* sq = smem + C::kOffQ + (qi * 2 + blk) * C::kQSlot;
                    const uint8_t* sk = smem + C::kOffK + st * C::kKSlot;
                    const uint64_t a = umma_desc_k_sw128(smem_u32(sq));
                    const uint64_t bdesc = umma_desc_k_sw128(smem_u32(sk));
                    const uint32_t d = tmem_base + blk * 128 + (c.t & 1) * 64;
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(d, a + 2 * k, bdesc + 2 * k, idesc_qk, k != 0);
                    if (C::TAIL)
                        umma_f16(d, umma_desc_k_sw32(smem_u32(sq + C::kQMain)), umma_desc_k_sw32(smem_u32(sk + C::kKMain)), idesc_qk, 1);
                    umma_commit(&s_full[blk * 2 + (c.t & 1)]);
                    if (c.j == nkt - 1) umma_commit(&q_empty[qi]);             // this block's last Q K^T of the item is issued
                };
                auto issue_pv = [&](const Cursor& c) {
                    const int st = c.t % KV;
                    const bool first = c.j == 0;
                    if (C::TAIL) mbar_wait(&v_ones[st], (c.t / KV) & 1);
                    if (first) mbar_wait(&o_empty[blk], (c.it & 1) ^ 1);        // previous item's O has been read out
                    mbar_wait(&p_full[blk * 2 + (c.t & 1)], (c.t >> 1) & 1);
                    tc_fence_after();
                    const uint8_t* sv = smem + C::kOffV + st * C::kKSlot;
                    const uint32_t a = tmem_base + blk * 128 + (c.t & 1) * 64;   // P: 128 lanes x 32 columns of packed bf16
                    const uint32_t d = tmem_base + 256 + blk * 128;
                    const uint64_t b1 = umma_desc_mn_sw32(smem_u32(smem + C::kOffOnes), 2048, 256);   // every element is 1.0
#pragma unroll
                    for (int kk = 0; kk < C::BN / 16; ++kk) {
                        const uint32_t acc = (!first || kk != 0) ? 1u : 0u;
                        const uint64_t bm = umma_desc_mn_sw128(smem_u32(sv + kk * 16 * 128), 8192, 1024);
                        umma_f16_ts(d, a + kk * 8, bm, idesc_pv_main, acc);
                        if (C::TAIL) {
                            const uint64_t bt = umma_desc_mn_sw32(smem_u32(sv + C::kKMain + kk * 16 * 32), 2048, 256);
                            umma_f16_ts(d + 64, a + kk * 8, bt, idesc_pv_tail, acc);
                        } else {
                            umma_f16_ts(d + C::LCOL, a + kk * 8, b1, idesc_pv_tail, acc);              // row sums of P
                        }
                    }
                    umma_commit(&o_full[blk * 2 + (c.t & 1)]);
                    umma_commit(&kv_empty[st]);                                 // this block's Q K^T and P V of the tile are issued
                };
                Cursor cq{0, 0, 0}, cp{0, 0, 0};
                for (int pre = 0; pre < 2 && cq.t < T; ++pre) { issue_qk(cq); advance(cq); }
                while (cp.t < T) {
                    issue_pv(cp);
                    advance(cp);
                    if (cq.t < T) { issue_qk(cq); advance(cq); }
                }
            }
        } else if (warp == 3 && C::TAIL) {
            // ============================ V tail: columns 72..79 := 1.0 (row-sum columns) ============================
            const int n_items = (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
            const uint32_t T = (uint32_t)n_items * (uint32_t)nkt;
            for (uint32_t g = 0; g < T; ++g) {
                const int st = g % KV;
                mbar_wait(&kv_full[st], (g / KV) & 1);
                uint8_t* svt = smem + C::kOffV + st * C::kKSlot + C::kKMain;
#pragma unroll
                for (int r = lane_id(); r < C::BN; r += 32)                        // SW32: 16-byte chunk index ^= bit 2 of the row
                    *reinterpret_cast<uint4*>(svt + r * 32 + ((1 ^ ((r >> 2) & 1)) << 4)) =
                        make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
                fence_proxy_async();
                __syncwarp();
                if (lane_id() == 0) mbar_arrive(&v_ones[st]);
            }
        }
    } else {
        // ============================ softmax / output: warpgroup `blk` owns query block `blk` ============================
        setmaxnreg_inc<200>();
        const int ew = (warp - 4) & 3;
        const int blk = (warp - 4) >> 2;
        const int row = ew * 32 + lane_id();
        const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
        const uint32_t o_addr = tmem_base + 256 + blk * 128 + lane_addr;
        uint32_t g = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            const int pr = item % p.npair;
            const int h = (item / p.npair) % p.H;
            const int b = item / (p.npair * p.H);
            // a warp whose 32 query rows all lie beyond S keeps the barrier protocol but does no arithmetic
            const bool dead = (pr * 2 + blk) * C::BM + ew * 32 >= p.S;
            float m = 0.f;
            for (int j = 0; j < nkt; ++j, ++g) {
                const int buf = g & 1;
                const uint32_t s_addr = tmem_base + blk * 128 + buf * 64 + lane_addr;
                mbar_wait(&s_full[blk * 2 + buf], (g >> 1) & 1);
                tc_fence_after();
                const int nvalid = p.S - j * C::BN;                      // keys of this tile inside the sequence (may exceed 64)
                // exact row max of the tile (first tile of an item, and the rare re-reference)
                auto row_max = [&]() {
                    float mx = -INFINITY;
                    uint32_t ra[32], rb[32];
                    tmem_ld_32x32b_x32(s_addr, ra);
                    tmem_ld_32x32b_x32(s_addr + 32, rb);
                    tmem_ld_wait();
                    if (nvalid >= 64) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(ra[i]), __uint_as_float(rb[i])));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (i < nvalid) mx = fmaxf(mx, __uint_as_float(ra[i]));
                            if (32 + i < nvalid) mx = fmaxf(mx, __uint_as_float(rb[i]));
                        }
                    }
                    return mx;
                };
                // pk = bf16x2(exp2(S*scale - ref - 7)), key pairs (2c, 2c+1) -> word c; returns the OR of all words
                auto exp_regs = [&](float neg_ref, uint32_t (&pk)[32]) {
                    uint32_t ored = 0;
                    uint32_t ra[32], rb[32];
                    tmem_ld_32x32b_x32(s_addr, ra);
                    tmem_ld_wait();
                    tmem_ld_32x32b_x32(s_addr + 32, rb);
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        uint32_t (&r)[32] = cc ? rb : ra;
                        if (cc) tmem_ld_wait();
                        if (nvalid < (cc + 1) * 32) {                   // ragged tail of the sequence: mask (warp-uniform branch)
#pragma unroll
                            for (int i = 0; i < 32; ++i) if (cc * 32 + i >= nvalid) r[i] = 0xff800000u;
                        }
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float e0, e1;
                            if (kPolyMod > 0 && (i % (kPolyMod > 0 ? kPolyMod : 1)) == 0) {     // this pair of scores goes to the FMA pipe
                                e0 = exp2_poly3(fmaf(__uint_as_float(r[2 * i]), p.scale_log2, neg_ref));
                                e1 = exp2_poly3(fmaf(__uint_as_float(r[2 * i + 1]), p.scale_log2, neg_ref));
                            } else {
                                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(__uint_as_float(r[2 * i]), p.scale_log2, neg_ref)));
                                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(__uint_as_float(r[2 * i + 1]), p.scale_log2, neg_ref)));
                            }
                            pk[cc * 16 + i] = pack_bf16(e0, e1);
                            ored |= pk[cc * 16 + i];
                        }
                    }
                    return ored;
                };
                if (!dead) {
                    if (j == 0) m = row_max();
                    uint32_t pk[32];
                    const uint32_t ored = exp_regs(fmaf(-m, p.scale_log2, -7.f), pk);
                    // exponent MSB of a bf16 half set  <=>  p * 2^-7 >= 2 (or inf / NaN): the reference is stale for that row
                    if (j > 0 && __any_sync(0xffffffffu, (ored & 0x40004000u) != 0)) {
                        const float m_new = fmaxf(m, row_max());
                        float corr;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(corr) : "f"((m - m_new) * p.scale_log2));
                        m = m_new;
                        exp_regs(fmaf(-m, p.scale_log2, -7.f), pk);
                        // rescale what P V has accumulated so far (O and the row sum) to the new reference
                        mbar_wait(&o_full[blk * 2 + ((g - 1) & 1)], ((g - 1) >> 1) & 1);
                        tc_fence_after();
#pragma unroll
                        for (int c = 0; c < C::OSPAN; c += 16) {
                            uint32_t t[16];
                            tmem_ld_32x32b_x16(o_addr + c, t);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * corr);
                            tmem_st_x16(o_addr + c, t);
                        }
                    }
                    tmem_st_x32(s_addr, pk);
                    tmem_st_wait();
                }
                tc_fence_before();
                mbar_arrive(&p_full[blk * 2 + buf]);
            }
            // ---- item epilogue: O / l -> bf16 -> global ----
            mbar_wait(&o_full[blk * 2 + ((g - 1) & 1)], ((g - 1) >> 1) & 1);
            tc_fence_after();
            if (!dead) {
                uint32_t t0[32], t1[32], t2[16];
                tmem_ld_32x32b_x32(o_addr, t0);
                tmem_ld_32x32b_x32(o_addr + 32, t1);
                tmem_ld_32x32b_x16(o_addr + 64, t2);                   // dh=72: O[64..71] | l x 8;   dh=64: l x 16
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&o_empty[blk]);
                const float inv = 1.f / __uint_as_float(t2[C::LCOL - 64]);
                const int q = (pr * 2 + blk) * C::BM + row;
                if (q < p.S) {
                    __nv_bfloat16* dst = p.out + ((int64_t)b * p.S + q) * p.ldo + h * DH;
#pragma unroll
                    for (int i = 0; i < 32; i += 8)
                        *reinterpret_cast<uint4*>(dst + i) = make_uint4(
                            pack_bf16(__uint_as_float(t0[i]) * inv, __uint_as_float(t0[i + 1]) * inv),
                            pack_bf16(__uint_as_float(t0[i + 2]) * inv, __uint_as_float(t0[i + 3]) * inv),
                            pack_bf16(__uint_as_float(t0[i + 4]) * inv, __uint_as_float(t0[i + 5]) * inv),
                            pack_bf16(__uint_as_float(t0[i + 6]) * inv, __uint_as_float(t0[i + 7]) * inv));
#pragma unroll
                    for (int i = 0; i < 32; i += 8)
                        *reinterpret_cast<uint4*>(dst + 32 + i) = make_uint4(
                            pack_bf16(__uint_as_float(t1[i]) * inv, __uint_as_float(t1[i + 1]) * inv),
                            pack_bf16(__uint_as_float(t1[i + 2]) * inv, __uint_as_float(t1[i + 3]) * inv),
                            pack_bf16(__uint_as_float(t1[i + 4]) * inv, __uint_as_float(t1[i + 5]) * inv),
                            pack_bf16(__uint_as_float(t1[i + 6]) * inv, __uint_as_float(t1[i + 7]) * inv));
                    if (DH > 64)
                        *reinterpret_cast<uint4*>(dst + 64) = make_uint4(
                            pack_bf16(__uint_as_float(t2[0]) * inv, __uint_as_float(t2[1]) * inv),
                            pack_bf16(__uint_as_float(t2[2]) * inv, __uint_as_float(t2[3]) * inv),
                            pack_bf16(__uint_as_float(t2[4]) * inv, __uint_as_float(t2[5]) * inv),
                            pack_bf16(__uint_as_float(t2[6]) * inv, __uint_as_float(t2[7]) * inv));
                }
            } else {
                tc_fence_before();
                mbar_arrive(&o_empty[blk]);
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

template <int DH, int kPolyMod>
static int launch_fa3(const void* qkv, int64_t ld, void* out, int64_t ldo, int B, int S, int H, float scale, cudaStream_t st) {
    using C = Fa3Cfg<DH>;
    CUtensorMap tm_q_main, tm_q_tail, tm_k_main, tm_k_tail;
    uint64_t dims[4] = {(uint64_t)DH, (uint64_t)(3 * H), (uint64_t)S, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)DH * 2, (uint64_t)ld * 2, (uint64_t)ld * 2 * (uint64_t)S};
    uint32_t box_q_main[4] = {64, 1, (uint32_t)C::BM, 1}, box_q_tail[4] = {16, 1, (uint32_t)C::BM, 1};
    uint32_t box_k_main[4] = {64, 1, (uint32_t)C::BN, 1}, box_k_tail[4] = {16, 1, (uint32_t)C::BN, 1};
    int rc;
    if ((rc = make_tmap_nd_bf16(&tm_q_main, qkv, 4, dims, strides, box_q_main, 128))) return rc;
    if ((rc = make_tmap_nd_bf16(&tm_k_main, qkv, 4, dims, strides, box_k_main, 128))) return rc;
    if (C::TAIL) {
        if ((rc = make_tmap_nd_bf16(&tm_q_tail, qkv, 4, dims, strides, box_q_tail, 32))) return rc;
        if ((rc = make_tmap_nd_bf16(&tm_k_tail, qkv, 4, dims, strides, box_k_tail, 32))) return rc;
    } else {
        tm_q_tail = tm_q_main;
        tm_k_tail = tm_k_main;
    }
    VB_SET_SMEM_ONCE(C::kSmem, attn_fwd3_sm100_kernel<DH, kPolyMod>);
    Fa3Params p;
    p.S = S; p.H = H; p.B = B;
    p.npair = (S + 2 * C::BM - 1) / (2 * C::BM);
    p.nkt = (S + C::BN - 1) / C::BN;
    p.items = B * H * p.npair;
    p.scale_log2 = scale * kLog2eC;
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.ldo = ldo;
    const int grid = p.items < num_sms() ? p.items : num_sms();
    attn_fwd3_sm100_kernel<DH, kPolyMod><<<grid, 384, C::kSmem, st>>>(tm_q_main, tm_q_tail, tm_k_main, tm_k_tail, p);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int attn_dense_sm100_v3(const void* qkv, int64_t ld, void* out, int64_t ldo, int B, int S, int H, int dh, float scale,
                        cudaStream_t st) {
    VB_REQUIRE(ld == (int64_t)3 * H * dh, "attn_dense_sm100_v3: qkv must be packed [B*S, 3*H*dh]");
    VB_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && ldo % 8 == 0, "attn_dense_sm100_v3: alignment");
    if (B == 0 || S == 0) return 0;
    if (dh == 72) return launch_fa3<72, 0>(qkv, ld, out, ldo, B, S, H, scale, st);
    if (dh == 64) return launch_fa3<64, 0>(qkv, ld, out, ldo, B, S, H, scale, st);
    VB_REQUIRE(false, "attn_dense_sm100_v3: unsupported head_dim %d", dh);
}

// A/B entry for the FMA-pipe exp2 share (kPolyMod = 2, 3 or 4: every kPolyMod-th pair of scores bypasses MUFU).  Not used by
// vidi_attn_dense; exists so that the option can be validated and timed (tests/test_preprocess_gpu.py is not the place: see
// tests/test_kernels_gpu.py::test_attn_dense_poly, parked with the other not-yet-run GPU tests).
int attn_dense_sm100_v3_poly(const void* qkv, int64_t ld, void* out, int64_t ldo, int B, int S, int H, int dh, float scale,
                             int poly_mod, cudaStream_t st) {
    VB_REQUIRE(ld == (int64_t)3 * H * dh, "attn_dense_sm100_v3_poly: qkv must be packed [B*S, 3*H*dh]");
    VB_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && ldo % 8 == 0, "attn_dense_sm100_v3_poly: alignment");
    VB_REQUIRE(S > 128, "attn_dense_sm100_v3_poly: S must exceed 128 (needs >= 3 key tiles)");
    if (B == 0) return 0;
#define VB_POLY_CASE(D, P) if (dh == D && poly_mod == P) return launch_fa3<D, P>(qkv, ld, out, ldo, B, S, H, scale, st)
    VB_POLY_CASE(72, 2); VB_POLY_CASE(72, 3); VB_POLY_CASE(72, 4);
    VB_POLY_CASE(64, 2); VB_POLY_CASE(64, 3); VB_POLY_CASE(64, 4);
#undef VB_POLY_CASE
    VB_REQUIRE(false, "attn_dense_sm100_v3_poly: unsupported head_dim %d / poly_mod %d", dh, poly_mod);
}

}  // namespace vb
