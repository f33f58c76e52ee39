// Frame pre-processing on the device: the reference's `process_images` 'resize' branch
// (Vidi1.5_9B/vidi/dataset/img_utils.py:181-187) is PIL.Image.resize((S,S), BICUBIC) on uint8 RGB followed by the SigLIP
// processor's affine.  These kernels restate Pillow's 8-bit resampler (libImaging/Resample.c: ImagingResampleHorizontal_8bpc /
// ImagingResampleVertical_8bpc) in the same integer arithmetic — 22-bit fixed-point taps, accumulator seeded with 1 << 21,
// arithmetic shift, clamp to [0,255], a uint8 image between the two passes — so results are bit-exact; the tap tables
// (xmin[out], kk[out][ksize]) come from the host (vidi_b200/preprocess.py::pil_bicubic_coeffs == precompute_coeffs +
// normalize_coeffs_8bpc).  HBM-bound byte work: one thread per output byte, consecutive threads on consecutive bytes.
//   pass 1 (horizontal):  [F*H, W, 3] -> [F*H, S, 3] uint8
//   pass 2 (vertical) fused with the affine and the layout change:  [F, H, S*3] -> [F, 3, S, S] bf16 = ((v/255) - mean) / std
#include "common.cuh"

namespace vb {

constexpr int kPilPrecisionBits = 32 - 8 - 2;

__device__ __forceinline__ int pil_clip8(int acc) {
    const int v = acc >> kPilPrecisionBits;                   // arithmetic shift: floor, like the C source's lookup index
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// src [outer, in_size, inner] -> dst [outer, out_size, inner]; one thread per output byte
__global__ void resample_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int64_t outer, int in_size,
                                   int out_size, int inner, const int* __restrict__ xmin, const int* __restrict__ kk, int ksize) {
    const int64_t total = outer * out_size * inner;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % inner);
        const int o = (int)((i / inner) % out_size);
        const int64_t r = i / ((int64_t)inner * out_size);
        const uint8_t* s = src + (r * in_size + xmin[o]) * inner + c;
        const int* k = kk + (int64_t)o * ksize;
        const int nmax = in_size - xmin[o];                    // taps past the edge have zero weight; do not read them
        int acc = 1 << (kPilPrecisionBits - 1);
        for (int t = 0; t < ksize && t < nmax; ++t) acc += (int)s[(int64_t)t * inner] * k[t];
        dst[i] = (uint8_t)pil_clip8(acc);
    }
}

// src [F, in_h, W, 3] uint8 (after the horizontal pass) -> dst [F, 3, out_h, W] bf16, value = ((v * rescale) - mean) / std
__global__ void resample_u8_to_chw_bf16_kernel(const uint8_t* __restrict__ src, __nv_bfloat16* __restrict__ dst, int F, int in_h,
                                               int out_h, int W, const int* __restrict__ ymin, const int* __restrict__ kk,
                                               int ksize, float rescale, float mean, float stdv) {
    const int64_t total = (int64_t)F * out_h * W * 3;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        // read-coalesced order: (f, y, x, c) with c fastest — the planar store is a 3-way strided write of 2-byte elements
        const int c = (int)(i % 3);
        const int x = (int)((i / 3) % W);
        const int y = (int)((i / (3 * (int64_t)W)) % out_h);
        const int f = (int)(i / (3 * (int64_t)W * out_h));
        const int64_t row_stride = (int64_t)W * 3;
        const uint8_t* s = src + ((int64_t)f * in_h + ymin[y]) * row_stride + (int64_t)x * 3 + c;
        const int* k = kk + (int64_t)y * ksize;
        const int nmax = in_h - ymin[y];
        int acc = 1 << (kPilPrecisionBits - 1);
        for (int t = 0; t < ksize && t < nmax; ++t) acc += (int)s[(int64_t)t * row_stride] * k[t];
        // three separately rounded fp32 operations, as the HF processor does them (no FMA contraction: it changes the bf16 result)
        const float v = __fdiv_rn(__fsub_rn(__fmul_rn((float)pil_clip8(acc), rescale), mean), stdv);
        dst[(((int64_t)f * 3 + c) * out_h + y) * W + x] = __float2bfloat16(v);
    }
}

// ------------------------------------------------------------------------------------------------------------------------------
// Whisper log-mel on the device (process_audio, vid_utils.py:52-63 -> HF WhisperFeatureExtractor): the 400-point DFT of the 3001
// hann-windowed, reflect-padded frames of a 30-s chunk and the 201 -> 128 mel projection are two tensor-core GEMMs in the
// 3-term split-bf16 form already used for the fp32 positional MLPs (x*w ~ xh*wh + xh*wl + xl*wh: A rows hold [hi|hi|lo], W rows
// [hi|lo|hi]); these kernels only build the operands and finish the features:
//   logmel_frames:  audio fp32 [C, n]            -> A1 bf16 [C*3001, 3*400]      frame t = samples t*160-200 .. +400 (reflected), * hann
//   (GEMM 1: A1 x Wdft^T -> Y fp32 [rows, 408] = [Re(0..200) | Im(0..200) | 0])
//   logmel_power:   Y                             -> A2 bf16 [rows, 3*208]        |X_k|^2, k = 0..200, zero padded
//   (GEMM 2: A2 x Wmel^T -> M fp32 [rows, 128])
//   logmel_max:     M -> per-chunk max of log10(max(M, 1e-10)) over the first 3000 frames
//   logmel_finish:  M -> out bf16 [C, 128, 3000] = (max(log10(...), chunk_max - 8) + 4) / 4     (last frame dropped)
// ------------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_hi_lo(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16(v);
    lo = __float2bfloat16(v - __bfloat162float(hi));
}

constexpr int kFft = 400, kHop = 160, kBins = 201, kBinsPad = 208, kFramesPerChunk = 3001;

__global__ void logmel_frames_kernel(const float* __restrict__ audio, const float* __restrict__ window, __nv_bfloat16* __restrict__ out,
                                     int n_samples) {
    const int64_t row = blockIdx.x;                                // chunk * 3001 + t
    const int ch = (int)(row / kFramesPerChunk), t = (int)(row % kFramesPerChunk);
    const float* a = audio + (int64_t)ch * n_samples;
    __nv_bfloat16* o = out + row * (3 * kFft);
    for (int c = threadIdx.x; c < kFft; c += blockDim.x) {
        int i = t * kHop - kFft / 2 + c;                            // centred frame; torch.stft(center=True, pad_mode="reflect")
        if (i < 0) i = -i;
        if (i >= n_samples) i = 2 * (n_samples - 1) - i;
        __nv_bfloat16 hi, lo;
        split_hi_lo(a[i] * window[c], hi, lo);
        o[c] = hi; o[kFft + c] = hi; o[2 * kFft + c] = lo;
    }
}

__global__ void logmel_power_kernel(const float* __restrict__ Y, int64_t ldy, __nv_bfloat16* __restrict__ out) {
    const int64_t row = blockIdx.x;
    const float* y = Y + row * ldy;
    __nv_bfloat16* o = out + row * (3 * kBinsPad);
    for (int k = threadIdx.x; k < kBinsPad; k += blockDim.x) {
        float p = 0.f;
        if (k < kBins) { const float re = y[k], im = y[kBins + k]; p = re * re + im * im; }
        __nv_bfloat16 hi, lo;
        split_hi_lo(p, hi, lo);
        o[k] = hi; o[kBinsPad + k] = hi; o[2 * kBinsPad + k] = lo;
    }
}

// one block per chunk: max over frames 0..2999 and all mels of log10(max(M, 1e-10))
__global__ void logmel_max_kernel(const float* __restrict__ M, int mels, float* __restrict__ chunk_max) {
    const int ch = blockIdx.x;
    const float* m = M + (int64_t)ch * kFramesPerChunk * mels;
    float mx = -INFINITY;
    const int n = (kFramesPerChunk - 1) * mels;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, m[i]);   // log10 is monotone: take it once at the end
    __shared__ float red[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
        mx = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (threadIdx.x == 0) chunk_max[ch] = log10f(fmaxf(mx, 1e-10f));
    }
}

// M [C*3001, mels] fp32 -> out [C, mels, 3000] bf16; 32 x 32 tiles transposed through shared memory
__global__ void logmel_finish_kernel(const float* __restrict__ M, int mels, const float* __restrict__ chunk_max,
                                     __nv_bfloat16* __restrict__ out) {
    __shared__ float tile[32][33];
    const int ch = blockIdx.z, t0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
    const int T = kFramesPerChunk - 1;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 256 threads: 8 rows per sweep
    for (int r = ty; r < 32; r += 8) {
        const int t = t0 + r, m = m0 + tx;
        tile[r][tx] = (t < T && m < mels) ? M[((int64_t)ch * kFramesPerChunk + t) * mels + m] : 1.f;
    }
    __syncthreads();
    const float floor_v = chunk_max[ch] - 8.0f;
    for (int r = ty; r < 32; r += 8) {
        const int m = m0 + r, t = t0 + tx;
        if (m < mels && t < T) {
            const float v = fmaxf(log10f(fmaxf(tile[tx][r], 1e-10f)), floor_v);
            out[((int64_t)ch * mels + m) * T + t] = __float2bfloat16((v + 4.0f) / 4.0f);
        }
    }
}

int logmel_frames(const float* audio, const float* window, void* out, int C, int n_samples, cudaStream_t st) {
    VB_REQUIRE(n_samples == (kFramesPerChunk - 1) * kHop, "logmel_frames: chunks must be 30 s at 16 kHz (480000 samples), got %d", n_samples);
    if (C == 0) return 0;
    logmel_frames_kernel<<<(unsigned)(C * kFramesPerChunk), 128, 0, st>>>(audio, window, (__nv_bfloat16*)out, n_samples);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int logmel_power(const float* Y, int64_t ldy, void* out, int64_t rows, cudaStream_t st) {
    VB_REQUIRE(ldy >= 2 * kBins, "logmel_power: ldy %lld < 402", (long long)ldy);
    if (rows == 0) return 0;
    logmel_power_kernel<<<(unsigned)rows, 64, 0, st>>>(Y, ldy, (__nv_bfloat16*)out);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int logmel_finish(const float* M, int C, int mels, float* chunk_max, void* out, cudaStream_t st) {
    if (C == 0) return 0;
    logmel_max_kernel<<<C, 1024, 0, st>>>(M, mels, chunk_max);
    VB_CUDA_CHECK(cudaGetLastError());
    dim3 grid((kFramesPerChunk - 1 + 31) / 32, (mels + 31) / 32, C);
    logmel_finish_kernel<<<grid, 256, 0, st>>>(M, mels, chunk_max, (__nv_bfloat16*)out);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int resample_u8(const uint8_t* src, uint8_t* dst, int64_t outer, int in_size, int out_size, int inner, const int* xmin,
                const int* kk, int ksize, cudaStream_t st) {
    VB_REQUIRE(in_size > 0 && out_size > 0 && inner > 0 && ksize > 0, "resample_u8: bad sizes");
    if (outer == 0) return 0;
    const int64_t total = outer * out_size * inner;
    const int64_t want = (total + 255) / 256;
    const int grid = (int)(want < (int64_t)num_sms() * 16 ? want : (int64_t)num_sms() * 16);
    resample_u8_kernel<<<grid, 256, 0, st>>>(src, dst, outer, in_size, out_size, inner, xmin, kk, ksize);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int resample_u8_to_chw_bf16(const uint8_t* src, void* dst, int F, int in_h, int out_h, int W, const int* ymin, const int* kk,
                            int ksize, float rescale, float mean, float stdv, cudaStream_t st) {
    VB_REQUIRE(in_h > 0 && out_h > 0 && W > 0 && ksize > 0 && stdv != 0.f, "resample_u8_to_chw_bf16: bad sizes");
    if (F == 0) return 0;
    const int64_t total = (int64_t)F * out_h * W * 3;
    const int64_t want = (total + 255) / 256;
    const int grid = (int)(want < (int64_t)num_sms() * 16 ? want : (int64_t)num_sms() * 16);
    resample_u8_to_chw_bf16_kernel<<<grid, 256, 0, st>>>(src, (__nv_bfloat16*)dst, F, in_h, out_h, W, ymin, kk, ksize, rescale,
                                                          mean, stdv);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace vb
