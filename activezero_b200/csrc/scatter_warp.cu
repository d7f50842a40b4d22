// Integer scatter warp.
// Reference: /root/reference/utils/warp_ops.py:20-47 (NVRTC kernels) and :55-95 (host).
//
// The reference runs ONE THREAD PER IMAGE ROW with a serial column loop whose write order makes
// the last writer win; that order is equivalent to "among the sources landing on a destination
// column, the one with the largest |disp| wins" (pos kernel: j descending => smallest j = largest
// disp last; neg kernel: j ascending => largest j = most negative disp last; ties cannot occur
// within one sign).  Here one CTA owns one (n, y) disparity row: every column proposes
// key = (|disp|+1)*W + j to a shared-memory atomicMax on its destination (integer max is
// order-independent => deterministic, bit-exact), then all C channel rows that share this
// disparity row (warp_ops.py:27, dbase) are gathered with coalesced stores; holes get 0.
#include "common.cuh"

namespace az {

constexpr int kSWThreads = 256;

// grid = (H, N); smem: W uint32 keys.  SHIFT (W <= 32768): key = (|disp|+1) << 16 | j -- the same order as
// (|disp|+1) * W + j, unpacked with a mask instead of a 32-bit modulo by a run-time W (ncu at B=8, 544x960: 97
// instructions per pixel with the multiply / modulo form, issue 65 %).
template <bool SHIFT>
__global__ void __launch_bounds__(kSWThreads) scatter_warp_kernel(const float* __restrict__ src,
                                                                 const int32_t* __restrict__ disp,
                                                                 float* __restrict__ dst,
                                                                 int32_t* __restrict__ sign_flags, int C, int H,
                                                                 int W) {
    extern __shared__ uint32_t key[];
    const int y = blockIdx.x, n = blockIdx.y;
    const size_t HW = (size_t)H * W;
    const int32_t* drow = disp + (size_t)n * HW + (size_t)y * W;
    for (int j = threadIdx.x; j < W; j += kSWThreads) key[j] = 0u;
    __syncthreads();
    int flags = 0;
    for (int j = threadIdx.x; j < W; j += kSWThreads) {
        const int d = drow[j];
        flags |= (d > 0 ? 1 : 0) | (d < 0 ? 2 : 0);
        const long long idx = (long long)j + d;
        if (idx >= 0 && idx < W) {
            const uint32_t ad = (uint32_t)(d < 0 ? -d : d);  // < W here
            atomicMax(&key[(int)idx], SHIFT ? (((ad + 1u) << 16) | (uint32_t)j) : ((ad + 1u) * (uint32_t)W + (uint32_t)j));
        }
    }
    if (sign_flags != nullptr) {
        flags = __reduce_or_sync(0xffffffffu, flags);
        if ((threadIdx.x & 31) == 0 && flags != 0) atomicOr(sign_flags, flags);
    }
    __syncthreads();
    for (int c = 0; c < C; ++c) {
        const float* srow = src + ((size_t)n * C + c) * HW + (size_t)y * W;
        float* orow = dst + ((size_t)n * C + c) * HW + (size_t)y * W;
        for (int x = threadIdx.x; x < W; x += kSWThreads) {
            const uint32_t k = key[x];
            orow[x] = k != 0u ? __ldg(srow + (SHIFT ? (k & 0xffffu) : (k % (uint32_t)W))) : 0.f;
        }
    }
}

}  // namespace az

using namespace az;

extern "C" int az_scatter_warp(const float* src, const int32_t* disp, float* dst, int32_t* sign_flags, int64_t N,
                               int64_t C, int64_t H, int64_t W, void* stream) {
    if (!src || !disp || !dst || N <= 0 || C <= 0 || H <= 0 || W <= 0) return AZ_ERR_BAD_ARG;
    // key packing needs (W+1)*W < 2^32; smem row of W keys
    if (W > 46340 || N > 65535 || W * 4 > 200 * 1024) return AZ_ERR_BAD_ARG;
    const size_t smem = (size_t)W * sizeof(uint32_t);
    const bool shift = W <= 32768 && tuning("AZ_SCATTER_SHIFT", 1) != 0;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(shift ? scatter_warp_kernel<true> : scatter_warp_kernel<false>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    dim3 grid((unsigned)H, (unsigned)N);
    if (shift)
        scatter_warp_kernel<true><<<grid, kSWThreads, smem, (cudaStream_t)stream>>>(src, disp, dst, sign_flags, (int)C, (int)H, (int)W);
    else
        scatter_warp_kernel<false><<<grid, kSWThreads, smem, (cudaStream_t)stream>>>(src, disp, dst, sign_flags, (int)C, (int)H, (int)W);
    AZ_LAUNCH_CHECK();
    return 0;
}
