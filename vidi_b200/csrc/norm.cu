// HBM-bound row kernels: RMSNorm (Gemma "(1+w)" and vidi "w" flavours), fused residual+post-norm+next-norm,
// LayerNorm with bias, the fused multimodal "embed finish" pass, and the (O,LSE) partial merge is in xattn.cu.
// One warp per row, 128-bit loads/stores, fp32 statistics with warp-shuffle reductions; rows are kept packed in
// registers between the statistic pass and the write pass so every byte crosses HBM exactly once.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace vb {

constexpr int kMaxVec = 16;          // uint4 (8 bf16) per lane -> rows up to 32*16*8 = 4096 columns
constexpr int kRowWarps = 4;         // warps (rows) per CTA

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
    float2 t;
    t = unpack_bf16(q.x); f[0] = t.x; f[1] = t.y;
    t = unpack_bf16(q.y); f[2] = t.x; f[3] = t.y;
    t = unpack_bf16(q.z); f[4] = t.x; f[5] = t.y;
    t = unpack_bf16(q.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// ------------------------------------------------------------------------------------------------
// y = xhat(x, eps) * (add_one ? 1 + w : w) * out_scale
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowWarps * 32)
rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ w,
               __nv_bfloat16* __restrict__ y, int64_t ldy, int rows, int D, float eps, int add_one, float out_scale) {
    const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = lane_id();
    const int nvec = D >> 3;
    const uint4* xr = reinterpret_cast<const uint4*>(x + (int64_t)row * ldx);
    uint4 cache[kMaxVec];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            cache[i] = ld_stream(xr + v);
            float f[8]; unpack8(cache[i], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
        }
    }
    ss = warp_sum(ss);
    const float inv = rsqrtf(ss / (float)D + eps) * out_scale;
    const uint4* wr = reinterpret_cast<const uint4*>(w);
    uint4* yr = reinterpret_cast<uint4*>(y + (int64_t)row * ldy);
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            float f[8], g[8];
            unpack8(cache[i], f);
            unpack8(__ldg(wr + v), g);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = f[j] * inv * (add_one ? 1.0f + g[j] : g[j]);
            yr[v] = pack8(f);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// x <- x + (post_mode ? G(y, w_post) : y);  optionally h = norm(x, w_next)  [Gemma: (1+w), Mistral: w]
// This is the pass between two GEMMs of a decoder layer (gemma.py:196-202 / 116-123): the o-proj (or
// down-proj) output y is post-normed, added into the residual stream, and the next GEMM's normalised
// input is produced in the same sweep: 2 reads + 2 writes of one row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowWarps * 32)
residual_norm_kernel(__nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ y, int64_t ldy,
                     const __nv_bfloat16* __restrict__ w_post, const __nv_bfloat16* __restrict__ w_next,
                     __nv_bfloat16* __restrict__ h, int64_t ldh, int rows, int D, float eps, int post_mode,
                     int next_add_one) {
    const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = lane_id();
    const int nvec = D >> 3;
    const uint4* yr = reinterpret_cast<const uint4*>(y + (int64_t)row * ldy);
    uint4* xr = reinterpret_cast<uint4*>(x + (int64_t)row * ldx);
    uint4 cache[kMaxVec];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            cache[i] = ld_stream(yr + v);
            float f[8]; unpack8(cache[i], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
        }
    }
    float inv = 1.0f;
    if (post_mode) { ss = warp_sum(ss); inv = rsqrtf(ss / (float)D + eps); }
    const uint4* wp = reinterpret_cast<const uint4*>(w_post);
    float ss2 = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            float f[8], r[8];
            unpack8(cache[i], f);
            unpack8(xr[v], r);
            if (post_mode) {
                float g[8]; unpack8(__ldg(wp + v), g);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    // post-norm output is rounded to bf16 before the residual add, as the reference's module boundary does
                    const float t = __bfloat162float(__float2bfloat16(f[j] * inv * (1.0f + g[j])));
                    r[j] += t;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) r[j] += f[j];
            }
            const uint4 packed = pack8(r);
            xr[v] = packed;
            cache[i] = packed;
            float q[8]; unpack8(packed, q);
#pragma unroll
            for (int j = 0; j < 8; ++j) ss2 += q[j] * q[j];
        }
    }
    if (h == nullptr) return;
    ss2 = warp_sum(ss2);
    const float inv2 = rsqrtf(ss2 / (float)D + eps);
    const uint4* wn = reinterpret_cast<const uint4*>(w_next);
    uint4* hr = reinterpret_cast<uint4*>(h + (int64_t)row * ldh);
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            float f[8], g[8];
            unpack8(cache[i], f);
            unpack8(__ldg(wn + v), g);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = f[j] * inv2 * (next_add_one ? 1.0f + g[j] : g[j]);
            hr[v] = pack8(f);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm with affine (towers): y = (x - mean) * rsqrt(var + eps) * w + b ; w,b fp32
// ------------------------------------------------------------------------------------------------
template <int kVec>
__global__ void __launch_bounds__(kRowWarps * 32)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                 const float* __restrict__ b, __nv_bfloat16* __restrict__ y, int64_t ldy, int rows, int D, float eps) {
    const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = lane_id();
    const int nvec = D >> 3;
    const uint4* xr = reinterpret_cast<const uint4*>(x + (int64_t)row * ldx);
    uint4 cache[kVec];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            cache[i] = ld_stream(xr + v);
            float f[8]; unpack8(cache[i], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += f[j];
        }
    }
    const float mean = warp_sum(s) / (float)D;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            float f[8]; unpack8(cache[i], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = f[j] - mean; ss += d * d; }
        }
    }
    const float inv = rsqrtf(warp_sum(ss) / (float)D + eps);
    uint4* yr = reinterpret_cast<uint4*>(y + (int64_t)row * ldy);
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            float f[8]; unpack8(cache[i], f);
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(w) + 2 * v);
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(w) + 2 * v + 1);
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(b) + 2 * v);
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(b) + 2 * v + 1);
            f[0] = (f[0] - mean) * inv * w0.x + b0.x; f[1] = (f[1] - mean) * inv * w0.y + b0.y;
            f[2] = (f[2] - mean) * inv * w0.z + b0.z; f[3] = (f[3] - mean) * inv * w0.w + b0.w;
            f[4] = (f[4] - mean) * inv * w1.x + b1.x; f[5] = (f[5] - mean) * inv * w1.y + b1.y;
            f[6] = (f[6] - mean) * inv * w1.z + b1.z; f[7] = (f[7] - mean) * inv * w1.w + b1.w;
            yr[v] = pack8(f);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Fused multimodal finish (multimodal.py:193-206 / 241-250; gemma.py:353-356):
//   x = w_mod * xhat(proj)  + sum_i pos_i[idx_i(n)]        (pos tables are already weight-less rms-normed)
//   m = (sum|x| != 0) & sample_valid
//   out = w_llm * xhat(x) * m * normalizer
// idx_i(n) = ((n + n_offset) / div_i) % mod_i + off_i.  Tables are fp32 [*, D].
// ------------------------------------------------------------------------------------------------
struct FinishTables {
    const float* tab[3];
    int div[3];
    int mod[3];
    int off[3];
    int ntab;
};

__global__ void __launch_bounds__(kRowWarps * 32)
mm_finish_kernel(const __nv_bfloat16* __restrict__ proj, int64_t ldp, const __nv_bfloat16* __restrict__ w_mod,
                 const __nv_bfloat16* __restrict__ w_llm, FinishTables t, int n_offset, int sample_valid,
                 float normalizer, __nv_bfloat16* __restrict__ out, int64_t ldo, uint8_t* __restrict__ mask, int rows,
                 int D, float eps) {
    const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = lane_id();
    const int nvec = D >> 3;
    const uint4* pr = reinterpret_cast<const uint4*>(proj + (int64_t)row * ldp);
    uint4 cache[kMaxVec];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            cache[i] = ld_stream(pr + v);
            float f[8]; unpack8(cache[i], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
        }
    }
    const float inv = rsqrtf(warp_sum(ss) / (float)D + eps);
    const float* trow[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        trow[k] = nullptr;
        if (k < t.ntab) {
            const int idx = ((row + n_offset) / t.div[k]) % t.mod[k] + t.off[k];
            trow[k] = t.tab[k] + (int64_t)idx * D;
        }
    }
    // second sweep: x in fp32 cannot be kept in 16 uint4, so it is rebuilt from the packed cache twice
    float ss2 = 0.f, sabs = 0.f;
    const uint4* wm = reinterpret_cast<const uint4*>(w_mod);
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            float f[8], g[8];
            unpack8(cache[i], f);
            unpack8(__ldg(wm + v), g);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = f[j] * inv * g[j];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (trow[k]) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(trow[k]) + 2 * v);
                    const float4 b = __ldg(reinterpret_cast<const float4*>(trow[k]) + 2 * v + 1);
                    f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w;
                    f[4] += b.x; f[5] += b.y; f[6] += b.z; f[7] += b.w;
                }
            }
            // the reference holds x in the activation dtype here; round so the statistics match
            const uint4 packed = pack8(f);
            cache[i] = packed;
            unpack8(packed, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) { ss2 += f[j] * f[j]; sabs += fabsf(f[j]); }
        }
    }
    ss2 = warp_sum(ss2);
    sabs = warp_sum(sabs);
    const bool valid = (sabs != 0.f) && sample_valid;
    if (lane == 0 && mask) mask[row] = valid ? 1 : 0;
    const float inv2 = valid ? rsqrtf(ss2 / (float)D + eps) : 0.f;
    const uint4* wl = reinterpret_cast<const uint4*>(w_llm);
    uint4* orow = reinterpret_cast<uint4*>(out + (int64_t)row * ldo);
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            float f[8], g[8];
            unpack8(cache[i], f);
            unpack8(__ldg(wl + v), g);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                // w_llm * xhat(x) * mask is held in the activation dtype before the * normalizer (gemma.py:353-356)
                const float t0 = __bfloat162float(__float2bfloat16(f[j] * inv2 * g[j]));
                f[j] = t0 * normalizer;
            }
            orow[v] = pack8(f);
        }
    }
}

// fp32 rows -> weight-less rms-normed fp32 rows (pos tables: rms_norm(pos_mlp(...)), norm.py:9-16)
__global__ void __launch_bounds__(kRowWarps * 32)
rmsnorm_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int rows, int D, float eps, int round_bf16) {
    const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = lane_id();
    const float* xr = x + (int64_t)row * D;
    float ss = 0.f;
    for (int c = lane; c < D; c += 32) {
        float v = xr[c];
        if (round_bf16) v = __bfloat162float(__float2bfloat16(v));   // pe.to(x.dtype) (pos.py:58)
        ss += v * v;
    }
    const float inv = rsqrtf(warp_sum(ss) / (float)D + eps);
    float* yr = y + (int64_t)row * D;
    for (int c = lane; c < D; c += 32) {
        float v = xr[c];
        if (round_bf16) v = __bfloat162float(__float2bfloat16(v));
        v *= inv;
        if (round_bf16) v = __bfloat162float(__float2bfloat16(v));
        yr[c] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
static inline int row_grid(int rows) { return (rows + kRowWarps - 1) / kRowWarps; }

int rmsnorm(const void* x, int64_t ldx, const void* w, void* y, int64_t ldy, int rows, int D, float eps, int add_one,
            float out_scale, cudaStream_t st) {
    VB_REQUIRE(D % 8 == 0 && D <= kMaxVec * 256 && ldx % 8 == 0 && ldy % 8 == 0, "rmsnorm: D=%d unsupported", D);
    if (rows == 0) return 0;
    rmsnorm_kernel<<<row_grid(rows), kRowWarps * 32, 0, st>>>((const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)w,
                                                               (__nv_bfloat16*)y, ldy, rows, D, eps, add_one, out_scale);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int residual_norm(void* x, int64_t ldx, const void* y, int64_t ldy, const void* w_post, const void* w_next, void* h,
                  int64_t ldh, int rows, int D, float eps, int post_mode, int next_add_one, cudaStream_t st) {
    VB_REQUIRE(D % 8 == 0 && D <= kMaxVec * 256 && ldx % 8 == 0 && ldy % 8 == 0 && ldh % 8 == 0,
               "residual_norm: D=%d unsupported", D);
    if (rows == 0) return 0;
    residual_norm_kernel<<<row_grid(rows), kRowWarps * 32, 0, st>>>(
        (__nv_bfloat16*)x, ldx, (const __nv_bfloat16*)y, ldy, (const __nv_bfloat16*)w_post,
        (const __nv_bfloat16*)w_next, (__nv_bfloat16*)h, ldh, rows, D, eps, post_mode, next_add_one);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int layernorm(const void* x, int64_t ldx, const float* w, const float* b, void* y, int64_t ldy, int rows, int D,
              float eps, cudaStream_t st) {
    VB_REQUIRE(D % 8 == 0 && D <= kMaxVec * 256 && ldx % 8 == 0 && ldy % 8 == 0, "layernorm: D=%d unsupported", D);
    if (rows == 0) return 0;
    // per-lane cache sized to the row width: fewer registers -> more rows in flight per SM for the 1152 / 1280-wide towers
    const int vec = (D / 8 + 31) / 32;
#define VB_LN(V) layernorm_kernel<V><<<row_grid(rows), kRowWarps * 32, 0, st>>>((const __nv_bfloat16*)x, ldx, w, b, (__nv_bfloat16*)y, ldy, rows, D, eps)
    if (vec <= 2) VB_LN(2); else if (vec <= 5) VB_LN(5); else if (vec <= 8) VB_LN(8); else VB_LN(16);
#undef VB_LN
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int mm_finish(const void* proj, int64_t ldp, const void* w_mod, const void* w_llm, const float* const* tabs,
              const int* divs, const int* mods, const int* offs, int ntab, int n_offset, int sample_valid,
              float normalizer, void* out, int64_t ldo, uint8_t* mask, int rows, int D, float eps, cudaStream_t st) {
    VB_REQUIRE(D % 8 == 0 && D <= kMaxVec * 256 && ntab >= 0 && ntab <= 3, "mm_finish: D=%d ntab=%d unsupported", D, ntab);
    if (rows == 0) return 0;
    FinishTables t;
    t.ntab = ntab;
    for (int i = 0; i < 3; ++i) {
        t.tab[i] = i < ntab ? tabs[i] : nullptr;
        t.div[i] = i < ntab ? divs[i] : 1;
        t.mod[i] = i < ntab ? mods[i] : 1;
        t.off[i] = i < ntab ? offs[i] : 0;
    }
    mm_finish_kernel<<<row_grid(rows), kRowWarps * 32, 0, st>>>(
        (const __nv_bfloat16*)proj, ldp, (const __nv_bfloat16*)w_mod, (const __nv_bfloat16*)w_llm, t, n_offset,
        sample_valid, normalizer, (__nv_bfloat16*)out, ldo, mask, rows, D, eps);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int rmsnorm_f32(const float* x, float* y, int rows, int D, float eps, int round_bf16, cudaStream_t st) {
    if (rows == 0) return 0;
    rmsnorm_f32_kernel<<<row_grid(rows), kRowWarps * 32, 0, st>>>(x, y, rows, D, eps, round_bf16);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace vb
