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

// ------------------------------------------------------------------------------------------
// The trainer's ground-truth chain in ONE launch (SURVEY.md §8a row a10's caller, /root/reference/train.py:255-272,
// test.py:109-110): the right-view disparity arrives at twice the working resolution, is halved with
// F.interpolate(scale_factor=0.5, mode="nearest") (= source pixel (2y, 2j)), truncated to int32 (`.type(torch.int)`),
// scatter-warped onto the left view with ITSELF as the payload (apply_disparity_cu(img_disp_r, img_disp_r.int())) and
// thresholded into the training mask (0 < disp_gt_l < max_disp).  The reference spends an interpolate, a cast, the
// scatter warp, two comparisons and a product (six launches, five HxW temporaries); here a CTA reads every other pixel
// of one source row, races the keys in shared memory and writes the warped disparity and the mask.
// grid = (H, N), H = H2 / 2, W = W2 / 2; smem: W keys + W values.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSWThreads) scatter_warp_gt_kernel(const float* __restrict__ disp2x,
                                                                    float* __restrict__ disp_out,
                                                                    uint8_t* __restrict__ mask_out,
                                                                    int32_t* __restrict__ sign_flags, float max_disp,
                                                                    int H2, int W2, int H, int W) {
    extern __shared__ uint32_t key[];
    float* val = reinterpret_cast<float*>(key + W);
    const int y = blockIdx.x, n = blockIdx.y;
    const float* srow = disp2x + ((size_t)n * H2 + 2 * (size_t)y) * W2;
    for (int j = threadIdx.x; j < W; j += kSWThreads) {
        key[j] = 0u;
        val[j] = __ldg(srow + 2 * j);
    }
    __syncthreads();
    int flags = 0;
    for (int j = threadIdx.x; j < W; j += kSWThreads) {
        const int d = __float2int_rz(val[j]);  // .type(torch.int): truncation toward zero
        flags |= (d > 0 ? 1 : 0) | (d < 0 ? 2 : 0);
        const long long idx = (long long)j + d;
        if (idx >= 0 && idx < W) {
            const uint32_t ad = (uint32_t)(d < 0 ? -d : d);  // < W here
            atomicMax(&key[(int)idx], ((ad + 1u) << 16) | (uint32_t)j);
        }
    }
    if (sign_flags != nullptr) {
        flags = __reduce_or_sync(0xffffffffu, flags);
        if ((threadIdx.x & 31) == 0 && flags != 0) atomicOr(sign_flags, flags);
    }
    __syncthreads();
    const size_t obase = ((size_t)n * H + y) * W;
    for (int x = threadIdx.x; x < W; x += kSWThreads) {
        const uint32_t k = key[x];
        const float v = k != 0u ? val[k & 0xffffu] : 0.f;
        disp_out[obase + x] = v;
        if (mask_out != nullptr) mask_out[obase + x] = (v < max_disp && v > 0.f) ? 1 : 0;
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

extern "C" int az_scatter_warp_gt(const float* disp2x, float* disp_out, uint8_t* mask_out, int32_t* sign_flags,
                                  float max_disp, int64_t N, int64_t H2, int64_t W2, void* stream) {
    if (!disp2x || !disp_out || N <= 0 || H2 < 2 || W2 < 2) return AZ_ERR_BAD_ARG;
    const int64_t H = H2 / 2, W = W2 / 2;  // floor(size * 0.5), F.interpolate with recompute_scale_factor=False
    if (W > 32768 || N > 65535 || H * W >= (1ll << 31) || H2 * W2 >= (1ll << 31)) return AZ_ERR_BAD_ARG;
    const size_t smem = (size_t)W * 2 * sizeof(uint32_t);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(scatter_warp_gt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    dim3 grid((unsigned)H, (unsigned)N);
    scatter_warp_gt_kernel<<<grid, kSWThreads, smem, (cudaStream_t)stream>>>(disp2x, disp_out, mask_out, sign_flags, max_disp,
                                                                          (int)H2, (int)W2, (int)H, (int)W);
    AZ_LAUNCH_CHECK();
    return 0;
}
