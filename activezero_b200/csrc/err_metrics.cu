// Fused error-metric reductions (a "next" row).
// Reference: /root/reference/utils/cascade_metrics.py:16-57 (compute_err_metric): EPE, bad1/bad2,
// clipped depth error and the 2/4/8 mm depth outlier rates -- seven boolean-mask gathers each followed
// by a .item() host sync (train.py:348-351, every training step).  Here: one pass over the four maps,
// eight accumulators, per-CTA partials reduced in a fixed order by the last kernel, ONE 64-byte read.
//   out[0] = n (mask count)            out[1] = sum |disp_pred - disp_gt|
//   out[2] = #(|ddisp| > 1)            out[3] = #(|ddisp| > 2)
//   out[4] = sum clip(|depth_gt*1000 - depth_pred*1000|, 0, 100)
//   out[5..7] = #(|depth_gt - depth_pred| > 2e-3 / 4e-3 / 8e-3)
// fp32 element arithmetic in the reference's op order (the counts are then exact); double sums.
#include "common.cuh"

namespace az {

constexpr int kEMThreads = 256;
constexpr int kEMVals = 8;

// grid = (nblk, B); partial: [B*nblk][8] doubles
__global__ void __launch_bounds__(kEMThreads) err_metrics_kernel(const float* __restrict__ disp_gt,
                                                                const float* __restrict__ depth_gt,
                                                                const float* __restrict__ disp_pred,
                                                                const float* __restrict__ depth_pred,
                                                                const float* __restrict__ focal,
                                                                const float* __restrict__ baseline,
                                                                const uint8_t* __restrict__ mask,
                                                                double* __restrict__ partial, int64_t HW) {
    __shared__ double red[32];
    const int b = blockIdx.y;
    const float fb = depth_pred == nullptr ? __fmul_rn(focal[b], baseline[b]) : 0.f;  // focal_length * baseline (:39)
    double acc[kEMVals];
#pragma unroll
    for (int k = 0; k < kEMVals; ++k) acc[k] = 0.0;
    for (int64_t p = (int64_t)blockIdx.x * kEMThreads + threadIdx.x; p < HW; p += (int64_t)gridDim.x * kEMThreads) {
        const size_t o = (size_t)b * HW + p;
        if (mask[o] == 0) continue;
        const float dg = disp_gt[o], dp = disp_pred[o], zg = depth_gt[o];
        const float zp = depth_pred == nullptr ? __fdiv_rn(fb, dp) : depth_pred[o];
        const float dd = fabsf(__fsub_rn(dg, dp));
        const float e1000 = fabsf(__fsub_rn(__fmul_rn(zg, 1000.0f), __fmul_rn(zp, 1000.0f)));  // :41-43
        const float zd = fabsf(__fsub_rn(zg, zp));                                              // :47
        acc[0] += 1.0;
        acc[1] += (double)dd;
        acc[2] += dd > 1.0f ? 1.0 : 0.0;
        acc[3] += dd > 2.0f ? 1.0 : 0.0;
        // torch.clip propagates NaN (a diverged model must not report a finite depth error); fminf/fmaxf drop it
        acc[4] += (e1000 != e1000) ? (double)e1000 : (double)fminf(fmaxf(e1000, 0.0f), 100.0f);
        acc[5] += zd > 2e-3f ? 1.0 : 0.0;
        acc[6] += zd > 4e-3f ? 1.0 : 0.0;
        acc[7] += zd > 8e-3f ? 1.0 : 0.0;
    }
#pragma unroll
    for (int k = 0; k < kEMVals; ++k) {
        const double s = block_sum(acc[k], red);
        if (threadIdx.x == 0) partial[((size_t)b * gridDim.x + blockIdx.x) * kEMVals + k] = s;
    }
}

__global__ void __launch_bounds__(256) err_metrics_finalize_kernel(const double* __restrict__ partial, int64_t n,
                                                                  double* __restrict__ out) {
    __shared__ double red[32];
    for (int k = 0; k < kEMVals; ++k) {
        double s = 0.0;
        for (int64_t r = threadIdx.x; r < n; r += 256) s += partial[r * kEMVals + k];
        s = block_sum(s, red);
        if (threadIdx.x == 0) out[k] = s;
    }
}

static int64_t em_blocks(int64_t HW) {
    int64_t n = ceil_div(HW, kEMThreads * 4);
    return n < 1 ? 1 : (n > 4 * kNumSMs ? 4 * kNumSMs : n);
}

}  // namespace az

using namespace az;

extern "C" int64_t az_error_metrics_workspace_bytes(int64_t B, int64_t H, int64_t W) {
    return B * em_blocks(H * W) * kEMVals * (int64_t)sizeof(double);
}

extern "C" int az_error_metrics(const float* disp_gt, const float* depth_gt, const float* disp_pred,
                                const float* depth_pred, const float* focal_length, const float* baseline,
                                const uint8_t* mask, double* out, void* workspace, int64_t B, int64_t H, int64_t W,
                                void* stream) {
    if (!disp_gt || !depth_gt || !disp_pred || !mask || !out || !workspace || B <= 0 || H <= 0 || W <= 0)
        return AZ_ERR_BAD_ARG;
    if (!depth_pred && (!focal_length || !baseline)) return AZ_ERR_BAD_ARG;
    if (B > 65535) return AZ_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t HW = H * W, nblk = em_blocks(HW);
    dim3 grid((unsigned)nblk, (unsigned)B);
    err_metrics_kernel<<<grid, kEMThreads, 0, st>>>(disp_gt, depth_gt, disp_pred, depth_pred, focal_length, baseline,
                                                   mask, (double*)workspace, HW);
    AZ_LAUNCH_CHECK();
    err_metrics_finalize_kernel<<<1, 256, 0, st>>>((const double*)workspace, B * nblk, out);
    AZ_LAUNCH_CHECK();
    return 0;
}
