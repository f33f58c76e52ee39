// Layout / gather kernels around the GEMMs: patch im2col, Whisper conv im2col, pad+bilinear+space-to-depth
// pooling gather, token embedding gather, sinusoid + split-precision packing for the fp32 positional MLPs.
// All are single-pass, 128-bit vectorised where the layout allows it.
#include "common.cuh"

namespace vb {

// ------------------------------------------------------------------------------------------------
// SigLIP patch im2col: images [F,3,S,S] bf16 -> A [F*P, Kpad] with k = c*p*p + ky*p + kx (conv weight order),
// zero padded to Kpad (multiple of 64).  (HF SiglipVisionEmbeddings.patch_embedding, K1)
// ------------------------------------------------------------------------------------------------
__global__ void patch_im2col_kernel(const __nv_bfloat16* __restrict__ img, __nv_bfloat16* __restrict__ out, int F,
                                    int S, int patch, int side, int Kpad) {
    const int row = blockIdx.x;                       // f*side*side + py*side + px
    const int f = row / (side * side);
    const int pr = row % (side * side);
    const int py = pr / side, px = pr % side;
    const int K = 3 * patch * patch;
    __nv_bfloat16* o = out + (int64_t)row * Kpad;
    for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
        __nv_bfloat16 v = __float2bfloat16(0.f);
        if (k < K) {
            const int c = k / (patch * patch);
            const int r = k % (patch * patch);
            const int ky = r / patch, kx = r % patch;
            v = img[(((int64_t)f * 3 + c) * S + (py * patch + ky)) * S + (px * patch + kx)];
        }
        o[k] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// Whisper conv1 im2col: mel [C,mels,T] (channel-major) -> A [C*T, 3*mels], A[(ch,t), j*mels + c] = mel[ch,c,t+j-1]
// ------------------------------------------------------------------------------------------------
__global__ void whisper_im2col1_kernel(const __nv_bfloat16* __restrict__ mel, __nv_bfloat16* __restrict__ out, int C,
                                       int mels, int T) {
    // block = 32 time steps x all (j,c); transpose through smem so both sides are coalesced
    extern __shared__ __nv_bfloat16 tile[];           // [mels][34]
    const int ch = blockIdx.y;
    const int t0 = blockIdx.x * 32;
    const __nv_bfloat16* m = mel + (int64_t)ch * mels * T;
    for (int i = threadIdx.x; i < mels * 34; i += blockDim.x) {
        const int c = i / 34, tt = i % 34;
        const int t = t0 + tt - 1;
        tile[i] = (t >= 0 && t < T) ? m[(int64_t)c * T + t] : __float2bfloat16(0.f);
    }
    __syncthreads();
    const int K = 3 * mels;
    for (int i = threadIdx.x; i < 32 * K; i += blockDim.x) {
        const int tt = i / K, k = i % K;
        const int j = k / mels, c = k % mels;
        if (t0 + tt < T) out[((int64_t)ch * T + t0 + tt) * K + k] = tile[c * 34 + tt + j];
    }
}

// Whisper conv2 im2col (k=3, stride 2, pad 1): x [C,T,d] token-major -> A [C*T/2, 3*d], row (ch,t') = rows 2t'-1..2t'+1
__global__ void whisper_im2col2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int C, int T,
                                       int d) {
    const int To = T / 2;
    const int row = blockIdx.x;
    const int ch = row / To, tp = row % To;
    const int nvec = d >> 3;
    for (int i = threadIdx.x; i < 3 * nvec; i += blockDim.x) {
        const int j = i / nvec, v = i % nvec;
        const int t = 2 * tp + j - 1;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (t >= 0 && t < T) q = *reinterpret_cast<const uint4*>(x + ((int64_t)ch * T + t) * d + v * 8);
        *reinterpret_cast<uint4*>(out + (int64_t)row * 3 * d + j * d + v * 8) = q;
    }
}

// ------------------------------------------------------------------------------------------------
// Conv2DPool gather (pool.py:23-32 + utils.py:143-150): tower output P [F, side*side, d] (token-major) ->
// X [F*(h/m)*(w/m), m*m*d] with X[(f,i,j), q*d + c] = R(f, c, m*i+dy, m*j+dx), q = dy*m+dx, where R is the
// zero-padded (side -> side+1) map, bilinearly resized (align_corners=False) to (h,w) when h != side+1.
// The channel order q*d+c (instead of the reference's c*m*m+q) is absorbed into the projector weight at load.
// ------------------------------------------------------------------------------------------------
__global__ void pool_s2d_kernel(const __nv_bfloat16* __restrict__ P, __nv_bfloat16* __restrict__ X, int F, int side,
                                int d, int h, int w, int m) {
    const int ho = h / m, wo = w / m;
    const int row = blockIdx.x;                       // (f, i, j)
    const int f = row / (ho * wo);
    const int ij = row % (ho * wo);
    const int i = ij / wo, j = ij % wo;
    const int pad = side + 1;
    const bool resize = (h != pad) || (w != pad);
    const float sh = (float)pad / (float)h, sw = (float)pad / (float)w;
    const int nvec = d >> 3;
    const __nv_bfloat16* Pf = P + (int64_t)f * side * side * d;
    for (int q = 0; q < m * m; ++q) {
        const int y = m * i + q / m, x = m * j + q % m;
        __nv_bfloat16* o = X + (int64_t)row * (m * m * d) + q * d;
        if (!resize) {
            const bool ok = (y < side) && (x < side);
            const uint4* src = reinterpret_cast<const uint4*>(Pf + (int64_t)(y * side + x) * d);
            for (int v = threadIdx.x; v < nvec; v += blockDim.x)
                reinterpret_cast<uint4*>(o)[v] = ok ? src[v] : make_uint4(0, 0, 0, 0);
        } else {
            // PyTorch upsample_bilinear2d, align_corners=False: src = scale*(dst+0.5)-0.5 clamped at 0
            float sy = sh * ((float)y + 0.5f) - 0.5f; if (sy < 0.f) sy = 0.f;
            float sx = sw * ((float)x + 0.5f) - 0.5f; if (sx < 0.f) sx = 0.f;
            const int y0 = (int)sy, x0 = (int)sx;
            const int y1 = y0 + (y0 < pad - 1 ? 1 : 0), x1 = x0 + (x0 < pad - 1 ? 1 : 0);
            const float ly = sy - (float)y0, lx = sx - (float)x0;
            const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
            const bool v00 = y0 < side && x0 < side, v01 = y0 < side && x1 < side;
            const bool v10 = y1 < side && x0 < side, v11 = y1 < side && x1 < side;
            const uint4* s00 = reinterpret_cast<const uint4*>(Pf + (int64_t)(y0 * side + x0) * d);
            const uint4* s01 = reinterpret_cast<const uint4*>(Pf + (int64_t)(y0 * side + x1) * d);
            const uint4* s10 = reinterpret_cast<const uint4*>(Pf + (int64_t)(y1 * side + x0) * d);
            const uint4* s11 = reinterpret_cast<const uint4*>(Pf + (int64_t)(y1 * side + x1) * d);
            for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
                float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                auto add = [&](const uint4* s, float wgt, bool ok) {
                    if (!ok) return;
                    const uint4 qv = s[v];
                    float2 t;
                    t = unpack_bf16(qv.x); acc[0] += wgt * t.x; acc[1] += wgt * t.y;
                    t = unpack_bf16(qv.y); acc[2] += wgt * t.x; acc[3] += wgt * t.y;
                    t = unpack_bf16(qv.z); acc[4] += wgt * t.x; acc[5] += wgt * t.y;
                    t = unpack_bf16(qv.w); acc[6] += wgt * t.x; acc[7] += wgt * t.y;
                };
                add(s00, w00, v00); add(s01, w01, v01); add(s10, w10, v10); add(s11, w11, v11);
                reinterpret_cast<uint4*>(o)[v] = make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]),
                                                            pack_bf16(acc[4], acc[5]), pack_bf16(acc[6], acc[7]));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// token embedding gather * normalizer (gemma.py:305,353-354): out[t] = bf16(E[ids[t]] * normalizer)
// ------------------------------------------------------------------------------------------------
__global__ void embed_gather_kernel(const int64_t* __restrict__ ids, const __nv_bfloat16* __restrict__ E,
                                    __nv_bfloat16* __restrict__ out, int T, int D, int vocab, float normalizer) {
    const int t = blockIdx.x;
    int64_t id = ids[t];
    if (id < 0 || id >= vocab) id = 0;
    const uint4* src = reinterpret_cast<const uint4*>(E + id * D);
    uint4* dst = reinterpret_cast<uint4*>(out + (int64_t)t * D);
    for (int v = threadIdx.x; v < (D >> 3); v += blockDim.x) {
        const uint4 q = src[v];
        float2 a = unpack_bf16(q.x), b = unpack_bf16(q.y), c = unpack_bf16(q.z), d2 = unpack_bf16(q.w);
        dst[v] = make_uint4(pack_bf16(a.x * normalizer, a.y * normalizer), pack_bf16(b.x * normalizer, b.y * normalizer),
                            pack_bf16(c.x * normalizer, c.y * normalizer),
                            pack_bf16(d2.x * normalizer, d2.y * normalizer));
    }
}

// ------------------------------------------------------------------------------------------------
// Positional MLP support (pos.py:11-65), evaluated at (near) fp32 accuracy on the bf16 tensor cores by the
// 3-term split  x*w ~= xh*wh + xh*wl + xl*wh  (xh = bf16(x), xl = bf16(x - xh)), laid out along K so one GEMM
// does all three terms:  A' = [xh | xh | xl]  (K' = 3K),  W' = [wh | wl | wh].
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16(v);
    lo = __float2bfloat16(v - __bfloat162float(hi));
}

// sinusoid rows for positions i0..i0+rows-1 of l total: p = i/(l-1)*(N-1); pe[2k]=sin(p*div[k]), pe[2k+1]=cos(..)
__global__ void sinusoid_split_kernel(const float* __restrict__ div_term, __nv_bfloat16* __restrict__ out, int rows,
                                      int i0, int l, int N, int D) {
    const int r = blockIdx.x;
    const float p = (float)(i0 + r) / (float)(l - 1) * (float)(N - 1);
    __nv_bfloat16* o = out + (int64_t)r * 3 * D;
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        const float a = p * div_term[c >> 1];
        const float v = (c & 1) ? cosf(a) : sinf(a);
        __nv_bfloat16 hi, lo;
        split_bf16(v, hi, lo);
        o[c] = hi; o[D + c] = hi; o[2 * D + c] = lo;
    }
}

// fp32 [rows, D] -> split A' [rows, 3D]  (mode 0: activations [hi|hi|lo]; mode 1: weights [hi|lo|hi])
__global__ void split3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t rows, int D, int mode) {
    const int64_t r = blockIdx.x;
    const float* xr = x + r * D;
    __nv_bfloat16* o = out + r * 3 * D;
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        __nv_bfloat16 hi, lo;
        split_bf16(xr[c], hi, lo);
        if (mode == 0) { o[c] = hi; o[D + c] = hi; o[2 * D + c] = lo; }
        else           { o[c] = hi; o[D + c] = lo; o[2 * D + c] = hi; }
    }
}

// fp32 -> bf16 cast with optional row gather (used for pos tables when a bf16 copy is wanted)
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = __float2bfloat16(x[i]);
}

// ------------------------------------------------------------------------------------------------
// Vidi-7B learned pooling (Vidi_7B/model/mm_vision/pool.py:6-26): Conv2d(d->d, k x k, stride 1, no bias) as a GEMM over
// the window gather below, then bilinear(align_corners=True) to s_out x s_out.
//   window gather: P [F, side*side, d] -> A [F*so*so, k*k*d], so = side-k+1, A[(f,y,x), (ky*k+kx)*d + c] = P[f,(y+ky)*side+(x+kx),c]
// ------------------------------------------------------------------------------------------------
__global__ void conv_window_gather_kernel(const __nv_bfloat16* __restrict__ P, __nv_bfloat16* __restrict__ A, int side, int d,
                                          int k) {
    const int so = side - k + 1;
    const int row = blockIdx.x;
    const int f = row / (so * so), yx = row % (so * so);
    const int y = yx / so, x = yx % so;
    const int nvec = d >> 3;
    for (int i = threadIdx.x; i < k * k * nvec; i += blockDim.x) {
        const int q = i / nvec, v = i % nvec;
        const int ky = q / k, kx = q % k;
        reinterpret_cast<uint4*>(A + (int64_t)row * k * k * d + q * d)[v] =
            reinterpret_cast<const uint4*>(P + ((int64_t)f * side * side + (y + ky) * side + (x + kx)) * d)[v];
    }
}
// X [F, si, si, d] token-major -> Y [F, so, so, d], PyTorch upsample_bilinear2d with align_corners=True
__global__ void bilinear_ac_kernel(const __nv_bfloat16* __restrict__ X, __nv_bfloat16* __restrict__ Y, int si, int so, int d) {
    const int row = blockIdx.x;
    const int f = row / (so * so), yx = row % (so * so);
    const int y = yx / so, x = yx % so;
    const float sc = so > 1 ? (float)(si - 1) / (float)(so - 1) : 0.f;
    const float sy = sc * (float)y, sx = sc * (float)x;
    const int y0 = min((int)sy, si - 1), x0 = min((int)sx, si - 1);
    const int y1 = y0 + (y0 < si - 1 ? 1 : 0), x1 = x0 + (x0 < si - 1 ? 1 : 0);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    const __nv_bfloat16* base = X + (int64_t)f * si * si * d;
    const uint4* s00 = reinterpret_cast<const uint4*>(base + (int64_t)(y0 * si + x0) * d);
    const uint4* s01 = reinterpret_cast<const uint4*>(base + (int64_t)(y0 * si + x1) * d);
    const uint4* s10 = reinterpret_cast<const uint4*>(base + (int64_t)(y1 * si + x0) * d);
    const uint4* s11 = reinterpret_cast<const uint4*>(base + (int64_t)(y1 * si + x1) * d);
    uint4* o = reinterpret_cast<uint4*>(Y + (int64_t)row * d);
    for (int v = threadIdx.x; v < (d >> 3); v += blockDim.x) {
        const uint4 a = s00[v], b = s01[v], c = s10[v], e = s11[v];
        const uint32_t aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w}, cc[4] = {c.x, c.y, c.z, c.w},
                       ee[4] = {e.x, e.y, e.z, e.w};
        uint32_t r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 fa = unpack_bf16(aa[j]), fb = unpack_bf16(bb[j]), fc = unpack_bf16(cc[j]), fe = unpack_bf16(ee[j]);
            r[j] = pack_bf16(w00 * fa.x + w01 * fb.x + w10 * fc.x + w11 * fe.x, w00 * fa.y + w01 * fb.y + w10 * fc.y + w11 * fe.y);
        }
        o[v] = make_uint4(r[0], r[1], r[2], r[3]);
    }
}

// ------------------------------------------------------------------------------------------------
int conv_window_gather(const void* P, void* A, int F, int side, int d, int k, cudaStream_t st) {
    VB_REQUIRE(d % 8 == 0 && k >= 1 && k <= side, "conv_window_gather: d=%d k=%d side=%d", d, k, side);
    if (F == 0) return 0;
    const int so = side - k + 1;
    conv_window_gather_kernel<<<F * so * so, 128, 0, st>>>((const __nv_bfloat16*)P, (__nv_bfloat16*)A, side, d, k);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int bilinear_ac(const void* X, void* Y, int F, int si, int so, int d, cudaStream_t st) {
    VB_REQUIRE(d % 8 == 0 && si >= 1 && so >= 1, "bilinear_ac: d=%d si=%d so=%d", d, si, so);
    if (F == 0) return 0;
    bilinear_ac_kernel<<<F * so * so, 128, 0, st>>>((const __nv_bfloat16*)X, (__nv_bfloat16*)Y, si, so, d);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int patch_im2col(const void* img, void* out, int F, int S, int patch, int Kpad, cudaStream_t st) {
    const int side = S / patch;
    if (F == 0) return 0;
    patch_im2col_kernel<<<F * side * side, 128, 0, st>>>((const __nv_bfloat16*)img, (__nv_bfloat16*)out, F, S, patch,
                                                          side, Kpad);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int whisper_im2col1(const void* mel, void* out, int C, int mels, int T, cudaStream_t st) {
    if (C == 0) return 0;
    dim3 grid((T + 31) / 32, C);
    whisper_im2col1_kernel<<<grid, 256, mels * 34 * 2, st>>>((const __nv_bfloat16*)mel, (__nv_bfloat16*)out, C, mels, T);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int whisper_im2col2(const void* x, void* out, int C, int T, int d, cudaStream_t st) {
    VB_REQUIRE(d % 8 == 0 && T % 2 == 0, "whisper_im2col2: d=%d T=%d unsupported", d, T);
    if (C == 0) return 0;
    whisper_im2col2_kernel<<<C * (T / 2), 128, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, C, T, d);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int pool_s2d(const void* P, void* X, int F, int side, int d, int h, int w, int m, cudaStream_t st) {
    VB_REQUIRE(d % 8 == 0 && h % m == 0 && w % m == 0, "pool_s2d: d=%d h=%d w=%d m=%d unsupported", d, h, w, m);
    if (F == 0) return 0;
    pool_s2d_kernel<<<F * (h / m) * (w / m), 128, 0, st>>>((const __nv_bfloat16*)P, (__nv_bfloat16*)X, F, side, d, h, w, m);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int embed_gather(const int64_t* ids, const void* E, void* out, int T, int D, int vocab, float normalizer, cudaStream_t st) {
    VB_REQUIRE(D % 8 == 0, "embed_gather: D=%d", D);
    if (T == 0) return 0;
    embed_gather_kernel<<<T, 128, 0, st>>>(ids, (const __nv_bfloat16*)E, (__nv_bfloat16*)out, T, D, vocab, normalizer);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int sinusoid_split(const float* div_term, void* out, int rows, int i0, int l, int N, int D, cudaStream_t st) {
    VB_REQUIRE(l > 1, "LearnablePosEmbd needs l > 1 (pos.py:42)");
    if (rows == 0) return 0;
    sinusoid_split_kernel<<<rows, 256, 0, st>>>(div_term, (__nv_bfloat16*)out, rows, i0, l, N, D);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int split3(const float* x, void* out, int64_t rows, int D, int mode, cudaStream_t st) {
    if (rows == 0) return 0;
    split3_kernel<<<(unsigned)rows, 256, 0, st>>>(x, (__nv_bfloat16*)out, rows, D, mode);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int cast_f32_bf16(const float* x, void* y, int64_t n, cudaStream_t st) {
    if (n == 0) return 0;
    cast_f32_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, (__nv_bfloat16*)y, n);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace vb
