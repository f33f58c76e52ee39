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
        const float v = ((float)pil_clip8(acc) * rescale - mean) / stdv;       // same fp32 operation order as the HF processor
        dst[(((int64_t)f * 3 + c) * out_h + y) * W + x] = __float2bfloat16(v);
    }
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
