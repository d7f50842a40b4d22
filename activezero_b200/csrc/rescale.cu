// Bilinear rescaling of the multi-scale reprojection loss.
// Reference: /root/reference/utils/reprojection.py:153-158
//   input_L_rs = F.interpolate(input_L, scale_factor=r, mode="bilinear")          (same for input_R)
//   pred_disp_l_rs = F.interpolate(pred_disp_l, scale_factor=r, mode="bilinear") * r
//   mask_rs = F.interpolate(mask, scale_factor=r, mode="bilinear").type(torch.bool)
// i.e. twelve eager launches per call in round 1.  One launch here produces all four rescaled tensors of a
// scale; a second one is the (deterministic, gather-style) gradient w.r.t. the disparity.
//
// torch's upsample_bilinear2d, align_corners=False, scale_factor given (so the kernel uses 1/r, not in/out):
//   src = max((dst + 0.5) * (1/r) - 0.5, 0);  i0 = int(src);  i1 = i0 + (i0 < in-1);  l1 = src - i0;  l0 = 1 - l1
//   out = h0*(w0*a + w1*b) + h1*(w0*c + w1*d)      (horizontal first, then vertical; see blend4)
#include "common.cuh"

namespace az {

struct Lin1 {
    int i0, i1;
    float l0, l1;
};

__device__ __forceinline__ Lin1 lin_index(float scale, int dst, int in_size) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    Lin1 r;
    r.i0 = min((int)src, in_size - 1);
    r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
    r.l1 = src - (float)r.i0;
    r.l0 = 1.0f - r.l1;
    return r;
}

// torch's CUDA kernel (upsample_bilinear2d_out_frame: what the reference runs) and its CPU kernel at realistic
// image sizes both evaluate h0*(w0*a + w1*b) + h1*(w0*c + w1*d): horizontal first, then vertical.  (For small
// outputs the CPU kernel switches to four pre-multiplied corner weights, which differs in the last bit; checked in
// the authoring container with torch 2.11.)  The same order, un-fused, reproduces F.interpolate bit for bit.
__device__ __forceinline__ float blend4(float a, float b, float c, float d, const Lin1& ly, const Lin1& lx) {
    const float top = __fadd_rn(__fmul_rn(lx.l0, a), __fmul_rn(lx.l1, b));
    const float bot = __fadd_rn(__fmul_rn(lx.l0, c), __fmul_rn(lx.l1, d));
    return __fadd_rn(__fmul_rn(ly.l0, top), __fmul_rn(ly.l1, bot));
}

__device__ __forceinline__ float bil_sample(const float* __restrict__ p, int W, const Lin1& ly, const Lin1& lx) {
    const float a = __ldg(p + (size_t)ly.i0 * W + lx.i0), b = __ldg(p + (size_t)ly.i0 * W + lx.i1);
    const float c = __ldg(p + (size_t)ly.i1 * W + lx.i0), d = __ldg(p + (size_t)ly.i1 * W + lx.i1);
    return blend4(a, b, c, d, ly, lx);
}

// planes: [0, BC) tgt, [BC, 2BC) src, [2BC, 2BC+B) disp (x dmul), [2BC+B, 2BC+2B) mask
// grid = (ceil(Wo/128), Ho, 2*B*C + 2*B)
__global__ void __launch_bounds__(128) rescale_fwd_kernel(const float* __restrict__ tgt, const float* __restrict__ src,
                                                          const float* __restrict__ disp,
                                                          const uint8_t* __restrict__ mask, float* __restrict__ tgt_o,
                                                          float* __restrict__ src_o, float* __restrict__ disp_o,
                                                          uint8_t* __restrict__ mask_o, int BC, int B, int H, int W,
                                                          int Ho, int Wo, float sh, float sw, float dmul) {
    const int xo = blockIdx.x * 128 + threadIdx.x;
    if (xo >= Wo) return;
    const int yo = blockIdx.y;
    int pl = blockIdx.z;
    const Lin1 ly = lin_index(sh, yo, H), lx = lin_index(sw, xo, W);
    const size_t HW = (size_t)H * W, HWo = (size_t)Ho * Wo;
    const size_t o = (size_t)yo * Wo + xo;
    if (pl < BC) {
        tgt_o[(size_t)pl * HWo + o] = bil_sample(tgt + (size_t)pl * HW, W, ly, lx);
        return;
    }
    pl -= BC;
    if (pl < BC) {
        src_o[(size_t)pl * HWo + o] = bil_sample(src + (size_t)pl * HW, W, ly, lx);
        return;
    }
    pl -= BC;
    if (pl < B) {
        if (disp != nullptr) disp_o[(size_t)pl * HWo + o] = __fmul_rn(bil_sample(disp + (size_t)pl * HW, W, ly, lx), dmul);
        return;
    }
    pl -= B;
    if (mask_o == nullptr) return;
    // the float mask (0/1) is interpolated and cast to bool: true iff a tap with non-zero weight is set
    uint8_t m = 1;
    if (mask != nullptr) {
        const uint8_t* mp = mask + (size_t)pl * HW;
        const float a = mp[(size_t)ly.i0 * W + lx.i0] ? 1.f : 0.f, b = mp[(size_t)ly.i0 * W + lx.i1] ? 1.f : 0.f;
        const float c = mp[(size_t)ly.i1 * W + lx.i0] ? 1.f : 0.f, d = mp[(size_t)ly.i1 * W + lx.i1] ? 1.f : 0.f;
        m = blend4(a, b, c, d, ly, lx) != 0.f;
    }
    mask_o[(size_t)pl * HWo + o] = m;
}

// gin[b,y,x] = dmul * sum over low-res pixels (yo,xo) tapping (y,x) of wy*wx*gout[b,yo,xo]; each full-res pixel
// gathers from the <= 3x3 low-res candidates around its own position: deterministic, no atomics.
// grid = (ceil(W/128), H, B)
__global__ void __launch_bounds__(128) rescale_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin, int H,
                                                          int W, int Ho, int Wo, float sh, float sw, float dmul) {
    const int x = blockIdx.x * 128 + threadIdx.x;
    if (x >= W) return;
    const int y = blockIdx.y, b = blockIdx.z;
    const float* g = gout + (size_t)b * Ho * Wo;
    const int yc = (int)floorf(((float)y + 0.5f) / sh - 0.5f), xc = (int)floorf(((float)x + 0.5f) / sw - 0.5f);
    float acc = 0.f;
    for (int yo = max(yc - 1, 0); yo <= min(yc + 2, Ho - 1); ++yo) {
        const Lin1 ly = lin_index(sh, yo, H);
        float wy = 0.f;
        if (ly.i0 == y) wy += ly.l0;
        if (ly.i1 == y) wy += ly.l1;
        if (wy == 0.f) continue;
        for (int xo = max(xc - 1, 0); xo <= min(xc + 2, Wo - 1); ++xo) {
            const Lin1 lx = lin_index(sw, xo, W);
            float wx = 0.f;
            if (lx.i0 == x) wx += lx.l0;
            if (lx.i1 == x) wx += lx.l1;
            if (wx != 0.f) acc = fmaf(wy * wx, __ldg(g + (size_t)yo * Wo + xo), acc);
        }
    }
    gin[((size_t)b * H + y) * W + x] = acc * dmul;
}

}  // namespace az

using namespace az;

extern "C" int az_bilinear_rescale_fwd(const float* tgt, const float* src, const float* disp, const uint8_t* mask,
                                       float* tgt_o, float* src_o, float* disp_o, uint8_t* mask_o, int64_t B, int64_t C,
                                       int64_t H, int64_t W, int64_t Ho, int64_t Wo, float scale_h, float scale_w,
                                       float disp_mul, void* stream) {
    if (!tgt || !src || !tgt_o || !src_o || B <= 0 || C <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0)
        return AZ_ERR_BAD_ARG;
    if (disp && !disp_o) return AZ_ERR_BAD_ARG;
    if (Ho > 65535 || 2 * B * C + 2 * B > 65535 || H * W >= (1ll << 31)) return AZ_ERR_BAD_ARG;
    dim3 grid((unsigned)ceil_div(Wo, 128), (unsigned)Ho, (unsigned)(2 * B * C + 2 * B));
    rescale_fwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(tgt, src, disp, mask, tgt_o, src_o, disp_o, mask_o,
                                                               (int)(B * C), (int)B, (int)H, (int)W, (int)Ho, (int)Wo,
                                                               scale_h, scale_w, disp_mul);
    AZ_LAUNCH_CHECK();
    return 0;
}

extern "C" int az_bilinear_rescale_bwd(const float* gout, float* gin, int64_t B, int64_t H, int64_t W, int64_t Ho,
                                       int64_t Wo, float scale_h, float scale_w, float disp_mul, void* stream) {
    if (!gout || !gin || B <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0) return AZ_ERR_BAD_ARG;
    if (H > 65535 || B > 65535 || scale_h < 1.0f || scale_w < 1.0f) return AZ_ERR_BAD_ARG;  // downscaling (r <= 1) only
    dim3 grid((unsigned)ceil_div(W, 128), (unsigned)H, (unsigned)B);
    rescale_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(gout, gin, (int)H, (int)W, (int)Ho, (int)Wo, scale_h,
                                                               scale_w, disp_mul);
    AZ_LAUNCH_CHECK();
    return 0;
}
