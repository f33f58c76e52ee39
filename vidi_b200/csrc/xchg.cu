// Cross-rank exchange of the text stream's cross-attention partials over NVLink peer memory.
//
// Replaces the reference's sequence-parallel Gather (Vidi1.5_9B/vidi/model/lmm/dattn/sequence_parallel/all_to_all.py:361,
// split.py:72-93): there every rank all-gathers full activations per op; here every rank holds a shard of the image / audio
// K||V cache, computes split-KV partials (O, LSE) of the ~32 text rows against its shard, and only those partials cross
// the NVSwitch.  One kernel per layer does the whole send side:
//
//   xattn_premerge_push:  per text row, LSE-merge this rank's key splits of each stream into ONE (O [DH], LSE) partial and
//                         store it straight into slot[rank] of every peer's exchange buffer (peer-mapped pointers, plain
//                         st.global over NVLink); the last block to finish publishes a sequence number in every peer's flag
//                         word (fence.sys + st.release.sys).
//   xattn_merge2 (attn.cu) is the receive side: it spins on its own flag words until all `world` sequence numbers have
//                         arrived, then merges the world partials per stream and adds the text self-attention.
//
// No NCCL call, no host synchronisation, 2 launches per layer.  Slots are double-buffered by sequence parity; a rank cannot
// overwrite a slot a peer is still reading because its push of exchange n+2 is stream-ordered after its merge of n+1, which
// waited for that peer's push of n+1, which is stream-ordered after the peer's merge of n.
//
// Exchange block of one rank (floats):  stream 0: O [rows, DH] | LSE [rows]   stream 1: O [rows, DH] | LSE [rows]
#include <cstring>

#include "common.cuh"

namespace vb {

constexpr int kMaxPeers = 16;

struct PushSrc {
    const float* O;     // [P, rows, DH]
    const float* L;     // [P, rows]
    int P;              // key splits of this rank
};
struct PushDst {
    float* base[kMaxPeers];          // exchange buffer of every peer (peer-mapped), already offset to the slot
    unsigned int* flag[kMaxPeers];   // flag word [slot][this rank] in every peer's flag array (nullptr: no signalling)
};

__global__ void __launch_bounds__(128)
xattn_premerge_push_kernel(PushSrc s0, PushSrc s1, int nsrc, int rows, int DH, PushDst dst, int world, int64_t my_block_off,
                           unsigned int seq, unsigned int* __restrict__ counter) {
    const int row = blockIdx.x;
    extern __shared__ float wts[];                 // [max P]
    __shared__ float red[4];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int64_t soff = my_block_off;
    for (int si = 0; si < nsrc; ++si) {
        const PushSrc& s = si == 0 ? s0 : s1;
        float lmax = -INFINITY;
        for (int p = tid; p < s.P; p += 128) {
            const float l = s.L[(int64_t)p * rows + row];
            wts[p] = l;
            lmax = fmaxf(lmax, l);
        }
        lmax = warp_max(lmax);
        if (lane == 0) red[warp] = lmax;
        __syncthreads();
        const float Lm = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
        __syncthreads();
        float dsum = 0.f;
        for (int p = tid; p < s.P; p += 128) {
            const float l = wts[p];
            const float e = (l == -INFINITY) ? 0.f : __expf(l - Lm);
            wts[p] = e;
            dsum += e;
        }
        dsum = warp_sum(dsum);
        if (lane == 0) red[warp] = dsum;
        __syncthreads();
        const float den = red[0] + red[1] + red[2] + red[3];
        const float inv = den > 0.f ? 1.0f / den : 0.f;
        const float lse = den > 0.f ? Lm + __logf(den) : -INFINITY;     // a shard with no valid key contributes weight 0
        for (int c = tid * 2; c < DH; c += 256) {
            float2 a = make_float2(0.f, 0.f);
#pragma unroll 4
            for (int p = 0; p < s.P; ++p) {
                const float2 v = *reinterpret_cast<const float2*>(s.O + ((int64_t)p * rows + row) * DH + c);
                a.x += wts[p] * v.x; a.y += wts[p] * v.y;
            }
            a.x *= inv; a.y *= inv;
            for (int r = 0; r < world; ++r)
                *reinterpret_cast<float2*>(dst.base[r] + soff + (int64_t)row * DH + c) = a;
        }
        if (tid == 0)
            for (int r = 0; r < world; ++r) dst.base[r][soff + (int64_t)rows * DH + row] = lse;
        __syncthreads();
        soff += (int64_t)rows * (DH + 1);
    }
    if (counter == nullptr) return;
    // publish: every block fences its peer stores system-wide, the last one to arrive writes the sequence number
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        const unsigned int prev = atomicAdd(counter, 1u);
        if (prev == (unsigned int)rows - 1) {
            *counter = 0;                           // all blocks have arrived; the next launch on this stream starts from 0
            __threadfence_system();
            for (int r = 0; r < world; ++r)
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst.flag[r]), "r"(seq) : "memory");
        }
    }
}

int xattn_premerge_push(const float* O0, const float* L0, int P0, const float* O1, const float* L1, int P1, int nsrc, int rows,
                        int dh, float* const* peer_base, unsigned int* const* peer_flag, int world, int64_t my_block_off,
                        unsigned int seq, unsigned int* counter, cudaStream_t st) {
    if (rows == 0 || nsrc == 0) return 0;
    VB_REQUIRE(nsrc >= 1 && nsrc <= 2 && dh % 2 == 0 && world >= 1 && world <= kMaxPeers, "xattn_premerge_push: nsrc=%d dh=%d world=%d",
               nsrc, dh, world);
    VB_REQUIRE(P0 > 0 && (nsrc < 2 || P1 > 0), "xattn_premerge_push: empty split list");
    PushSrc s0{O0, L0, P0}, s1{O1, L1, nsrc > 1 ? P1 : 0};
    PushDst d;
    for (int r = 0; r < kMaxPeers; ++r) {
        d.base[r] = r < world ? peer_base[r] : nullptr;
        d.flag[r] = (r < world && peer_flag) ? peer_flag[r] : nullptr;
    }
    const int pmax = P0 > s1.P ? P0 : s1.P;
    xattn_premerge_push_kernel<<<rows, 128, pmax * sizeof(float), st>>>(s0, s1, nsrc, rows, dh, d, world, my_block_off, seq,
                                                                        peer_flag ? counter : nullptr);
    VB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------------------
// peer memory: one cudaMalloc'ed exchange arena per rank, exported / imported through CUDA IPC handles (one process per GPU).
// ------------------------------------------------------------------------------------------------------------------------------
int p2p_alloc(int64_t bytes, void** ptr, void* handle64) {
    VB_REQUIRE(bytes > 0 && ptr && handle64, "p2p_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    VB_CUDA_CHECK(cudaMalloc(ptr, (size_t)bytes));
    VB_CUDA_CHECK(cudaMemset(*ptr, 0, (size_t)bytes));
    VB_CUDA_CHECK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    VB_CUDA_CHECK(cudaIpcGetMemHandle(&h, *ptr));
    memcpy(handle64, &h, 64);
    return 0;
}
int p2p_open(const void* handle64, void** ptr) {
    VB_REQUIRE(ptr && handle64, "p2p_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    VB_CUDA_CHECK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int p2p_close(void* ptr) {
    if (ptr) VB_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
    return 0;
}
int p2p_free(void* ptr) {
    if (ptr) VB_CUDA_CHECK(cudaFree(ptr));
    return 0;
}

}  // namespace vb
