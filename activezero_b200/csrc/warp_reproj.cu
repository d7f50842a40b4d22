// Bilinear disparity warp and the fused warp + masked-MSE reprojection losses
// (SURVEY.md §8a rows a6, a7, a8).
// Reference: /root/reference/utils/reprojection.py:13-35 (apply_disparity), :81-96
// (get_reprojection_error_old), :99-127 (get_reproj_error_patch).
//
// The sample position follows the reference's fp32 op order (common.cuh: sample_pos), the
// stand-alone warp follows torch's corner order/FMA chain so it is bit-identical to the CPU
// oracle; the fused losses never materialise the unfolded / warped [B,C*ps*ps,H,W] tensors:
// one CTA owns one output row, builds the 11 (ps) vertically-interpolated source rows and the
// ps target rows in shared memory (zero borders folded into a padded layout, so the tap loop
// has no bounds checks), and each thread walks the ps x ps taps of its pixels accumulating the
// squared residual and d(residual)/d(xs).  The kernels are shared-memory/issue bound, not
// HBM bound (compulsory traffic is ~4 image planes) -- see DESIGN.md.
#include "common.cuh"

namespace az {

// ------------------------------------------------------------------------------------------
// a6 forward: out[b,c,i,j].  grid = (ceil(W/256), H, B)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) warp_fwd_kernel(const float* __restrict__ img, const float* __restrict__ disp,
                                                       const float* __restrict__ lin_x,
                                                       const float* __restrict__ lin_y, float* __restrict__ out,
                                                       int C, int H, int W) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= W) return;
    const int i = blockIdx.y, b = blockIdx.z;
    const size_t HW = (size_t)H * W;
    const float d = disp[(size_t)b * HW + (size_t)i * W + j];
    const Axis ax = make_axis(sample_pos(__ldg(lin_x + j), __fdiv_rn(d, (float)W), (float)W), W);
    const Axis ay = make_axis(sample_pos(__ldg(lin_y + i), 0.0f, (float)H), H);
    const float nw = __fmul_rn(ax.e, ay.e), ne = __fmul_rn(ax.w, ay.e);
    const float sw = __fmul_rn(ax.e, ay.w), se = __fmul_rn(ax.w, ay.w);
    const bool m00 = ax.v0 && ay.v0, m01 = ax.v1 && ay.v0, m10 = ax.v0 && ay.v1, m11 = ax.v1 && ay.v1;
    const size_t o00 = (size_t)ay.i0 * W + ax.i0;
    for (int c = 0; c < C; ++c) {
        const float* p = img + ((size_t)b * C + c) * HW;
        const float v00 = m00 ? __ldg(p + o00) : 0.f, v01 = m01 ? __ldg(p + o00 + 1) : 0.f;
        const float v10 = m10 ? __ldg(p + o00 + W) : 0.f, v11 = m11 ? __ldg(p + o00 + W + 1) : 0.f;
        float acc = __fmul_rn(v00, nw);
        acc = __fmaf_rn(v01, ne, acc);
        acc = __fmaf_rn(v10, sw, acc);
        acc = __fmaf_rn(v11, se, acc);
        out[((size_t)b * C + c) * HW + (size_t)i * W + j] = acc;
    }
}

// a6 forward, four adjacent pixels per thread (W % 4 == 0, 16-byte aligned rows): one 128-bit disparity load,
// sixteen independent gathers in flight and one 128-bit store per channel -- the one-pixel form above keeps only
// ~16 bytes per thread in flight and is bound by the two dependent memory round trips per wave.
// Same op order as warp_fwd_kernel (bit-identical results).  grid = (ceil(H*W/4/256), B)
struct Bil4 {
    float nw[4], ne[4], sw[4], se[4];
    int o00[4];
    unsigned m;  // bit 4*t + {0,1,2,3}: corner {00,01,10,11} of pixel t is inside the image
};

__device__ __forceinline__ Bil4 bil4_setup(const float4 d4, const float4 lx4, const Axis ay, int W) {
    Bil4 q;
    const float dd[4] = {d4.x, d4.y, d4.z, d4.w}, lx[4] = {lx4.x, lx4.y, lx4.z, lx4.w};
    q.m = 0u;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const Axis ax = make_axis(sample_pos(lx[t], __fdiv_rn(dd[t], (float)W), (float)W), W);
        q.nw[t] = __fmul_rn(ax.e, ay.e);
        q.ne[t] = __fmul_rn(ax.w, ay.e);
        q.sw[t] = __fmul_rn(ax.e, ay.w);
        q.se[t] = __fmul_rn(ax.w, ay.w);
        q.o00[t] = ay.i0 * W + ax.i0;
        const unsigned mm = (ax.v0 && ay.v0 ? 1u : 0u) | (ax.v1 && ay.v0 ? 2u : 0u) | (ax.v0 && ay.v1 ? 4u : 0u) |
                            (ax.v1 && ay.v1 ? 8u : 0u);
        q.m |= mm << (4 * t);
    }
    return q;
}

__device__ __forceinline__ void bil4_gather(const float* __restrict__ p, const Bil4& q, int W, float (&v00)[4],
                                            float (&v01)[4], float (&v10)[4], float (&v11)[4]) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const unsigned mm = q.m >> (4 * t);
        v00[t] = (mm & 1u) ? __ldg(p + q.o00[t]) : 0.f;
        v01[t] = (mm & 2u) ? __ldg(p + q.o00[t] + 1) : 0.f;
        v10[t] = (mm & 4u) ? __ldg(p + q.o00[t] + W) : 0.f;
        v11[t] = (mm & 8u) ? __ldg(p + q.o00[t] + W + 1) : 0.f;
    }
}

__global__ void __launch_bounds__(256) warp_fwd4_kernel(const float* __restrict__ img, const float* __restrict__ disp,
                                                        const float* __restrict__ lin_x,
                                                        const float* __restrict__ lin_y, float* __restrict__ out,
                                                        int C, int H, int W) {
    const int HW = H * W;
    const int p = 4 * (blockIdx.x * 256 + threadIdx.x);
    if (p >= HW) return;
    const int b = blockIdx.y;
    const int i = p / W, j = p - i * W;
    const float4 d4 = *reinterpret_cast<const float4*>(disp + (size_t)b * HW + p);
    const float4 lx4 = __ldg(reinterpret_cast<const float4*>(lin_x + j));
    const Axis ay = make_axis(sample_pos(__ldg(lin_y + i), 0.0f, (float)H), H);
    const Bil4 q = bil4_setup(d4, lx4, ay, W);
    for (int c = 0; c < C; ++c) {
        const float* pc = img + ((size_t)b * C + c) * HW;
        float v00[4], v01[4], v10[4], v11[4], o[4];
        bil4_gather(pc, q, W, v00, v01, v10, v11);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float acc = __fmul_rn(v00[t], q.nw[t]);
            acc = __fmaf_rn(v01[t], q.ne[t], acc);
            acc = __fmaf_rn(v10[t], q.sw[t], acc);
            o[t] = __fmaf_rn(v11[t], q.se[t], acc);
        }
        *reinterpret_cast<float4*>(out + ((size_t)b * C + c) * HW + p) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// a8 (ps = 1) fused warp + masked squared residual, four adjacent pixels per thread, direct gathers (no
// shared-memory staging: with a single tap per pixel there is nothing to reuse).  One CTA per image row;
// writes the warped image (same op order as warp_fwd_kernel: bit-identical to the oracle), the pre-gradient
// and the per-row partial sums consumed by reproj_finalize_kernel.  grid = (H, B), 256 threads.
template <int MINB>
__global__ void __launch_bounds__(256, MINB) reproj_ps1_kernel(const float* __restrict__ tgt, const float* __restrict__ src,
                                                         const float* __restrict__ disp, float sign,
                                                         const uint8_t* __restrict__ mask,
                                                         const float* __restrict__ lin_x,
                                                         const float* __restrict__ lin_y, float* __restrict__ warped,
                                                         float* __restrict__ gpre, double* __restrict__ partial, int C,
                                                         int H, int W) {
    __shared__ double red_t[8];
    __shared__ int red_c[8];
    const int i = blockIdx.x, b = blockIdx.y;
    const int HW = H * W;
    const size_t rowbase = (size_t)b * HW + (size_t)i * W;
    const Axis ay = make_axis(sample_pos(__ldg(lin_y + i), 0.0f, (float)H), H);
    double tot = 0.0;
    int cnt = 0;
    for (int j = 4 * threadIdx.x; j < W; j += 4 * 256) {
        float4 d4 = *reinterpret_cast<const float4*>(disp + rowbase + j);
        d4.x *= sign; d4.y *= sign; d4.z *= sign; d4.w *= sign;
        const float4 lx4 = __ldg(reinterpret_cast<const float4*>(lin_x + j));
        const Bil4 q = bil4_setup(d4, lx4, ay, W);
        const unsigned m4 = mask == nullptr ? 0x01010101u : *reinterpret_cast<const unsigned*>(mask + rowbase + j);
        float g[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < C; ++c) {
            const size_t plane = ((size_t)b * C + c) * HW;
            float v00[4], v01[4], v10[4], v11[4], o[4];
            bil4_gather(src + plane, q, W, v00, v01, v10, v11);
            const float4 t4 = *reinterpret_cast<const float4*>(tgt + plane + (size_t)i * W + j);
            const float tt[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float acc = __fmul_rn(v00[t], q.nw[t]);
                acc = __fmaf_rn(v01[t], q.ne[t], acc);
                acc = __fmaf_rn(v10[t], q.sw[t], acc);
                o[t] = __fmaf_rn(v11[t], q.se[t], acc);
                if ((m4 >> (8 * t)) & 0xffu) {
                    const float r = o[t] - tt[t];
                    tot += (double)(r * r);
                    g[t] = fmaf(r, ay.e * (v01[t] - v00[t]) + ay.w * (v11[t] - v10[t]), g[t]);
                }
            }
            if (warped != nullptr)
                *reinterpret_cast<float4*>(warped + plane + (size_t)i * W + j) = make_float4(o[0], o[1], o[2], o[3]);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t)
            cnt += ((m4 >> (8 * t)) & 0xffu) ? 1 : 0;
        if (gpre != nullptr) *reinterpret_cast<float4*>(gpre + rowbase + j) = make_float4(g[0], g[1], g[2], g[3]);
    }
    // one reduction for both: the count is an integer (redux.sync), the sum of squares stays in double; a fixed
    // order (lanes by xor-shuffle, then warps 0..7) keeps the result run-to-run identical
    tot = warp_sum(tot);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) {
        red_t[threadIdx.x >> 5] = tot;
        red_c[threadIdx.x >> 5] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double bs = 0.0;
        int bc = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            bs += red_t[w];
            bc += red_c[w];
        }
        const size_t r = (size_t)b * H + i;
        partial[2 * r] = bs;
        partial[2 * r + 1] = (double)bc;
    }
}

// a6 backward w.r.t. the disparity (gimg != nullptr: round 1's atomic image gradient, kept for AZ_WARP_BWD_IMG=0).
__global__ void __launch_bounds__(256) warp_bwd_kernel(const float* __restrict__ img, const float* __restrict__ disp,
                                                       const float* __restrict__ lin_x,
                                                       const float* __restrict__ lin_y,
                                                       const float* __restrict__ gout, float* __restrict__ gimg,
                                                       float* __restrict__ gdisp, int C, int H, int W) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= W) return;
    const int i = blockIdx.y, b = blockIdx.z;
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)i * W + j;
    const float d = disp[(size_t)b * HW + pix];
    const Axis ax = make_axis(sample_pos(__ldg(lin_x + j), __fdiv_rn(d, (float)W), (float)W), W);
    const Axis ay = make_axis(sample_pos(__ldg(lin_y + i), 0.0f, (float)H), H);
    const float nw = ax.e * ay.e, ne = ax.w * ay.e, sw = ax.e * ay.w, se = ax.w * ay.w;
    const bool m00 = ax.v0 && ay.v0, m01 = ax.v1 && ay.v0, m10 = ax.v0 && ay.v1, m11 = ax.v1 && ay.v1;
    const size_t o00 = (size_t)ay.i0 * W + ax.i0;
    float gix = 0.f;
    for (int c = 0; c < C; ++c) {
        const size_t base = ((size_t)b * C + c) * HW;
        const float g = gout[base + pix];
        if (gdisp != nullptr) {
            const float* p = img + base;
            const float v00 = m00 ? __ldg(p + o00) : 0.f, v01 = m01 ? __ldg(p + o00 + 1) : 0.f;
            const float v10 = m10 ? __ldg(p + o00 + W) : 0.f, v11 = m11 ? __ldg(p + o00 + W + 1) : 0.f;
            gix += g * (ay.e * (v01 - v00) + ay.w * (v11 - v10));
        }
        if (gimg != nullptr) {
            float* q = gimg + base;
            if (m00) atomicAdd(q + o00, nw * g);
            if (m01) atomicAdd(q + o00 + 1, ne * g);
            if (m10) atomicAdd(q + o00 + W, sw * g);
            if (m11) atomicAdd(q + o00 + W + 1, se * g);
        }
    }
    if (gdisp != nullptr) {
        // chain of grid_sample's unnormalise (x W/2), `2*flow-1` (x 2) and `disp/width` (/ W)
        gdisp[(size_t)b * HW + pix] = ((gix * (0.5f * (float)W)) * 2.0f) / (float)W;
    }
}


// a6 backward w.r.t. the IMAGE, deterministic (round 2; torch's grid_sampler_2d_backward and round 1's kernel use
// float atomics, whose summation order changes from run to run).  One CTA owns one image row y' of one (b, c):
// ys depends on the output row only and grows by H/(H-1) per row, so y' is a bilinear corner of at most three
// output rows i (y0(i) in {y'-1, y'}); the threads walk those rows and add every output pixel's two horizontal
// corner contributions into a shared row of 64-bit FIXED-POINT accumulators (integer addition is associative:
// the result does not depend on the order of the atomics), scaled per row so that the worst case -- every pixel of
// the three rows landing on one column -- cannot overflow.  The row is then converted and STORED: every element
// of gimg is written exactly once, no zero-fill by the caller, no global atomics.
// grid = (H, C, B), 256 threads, dynamic smem = W * 8 bytes.
__global__ void __launch_bounds__(256) warp_bwd_img_kernel(const float* __restrict__ disp,
                                                           const float* __restrict__ lin_x,
                                                           const float* __restrict__ lin_y,
                                                           const float* __restrict__ gout, float* __restrict__ gimg,
                                                           int C, int H, int W) {
    extern __shared__ unsigned long long facc[];
    __shared__ float red[8];
    __shared__ int rows_s[4];
    __shared__ float wts_s[4];
    const int yp = blockIdx.x, c = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x;
    const size_t HW = (size_t)H * W;
    const float* gplane = gout + ((size_t)b * C + c) * HW;
    const float* dplane = disp + (size_t)b * HW;
    for (int x = tid; x < W; x += 256) facc[x] = 0ull;
    if (tid == 0) {
        int n = 0;
        for (int i = max(yp - 2, 0); i <= min(yp + 2, H - 1); ++i) {
            const Axis ay = make_axis(sample_pos(__ldg(lin_y + i), 0.0f, (float)H), H);
            float wy = 0.f;
            if (ay.v0 && ay.i0 == yp) wy += ay.e;
            if (ay.v1 && ay.i0 + 1 == yp) wy += ay.w;
            if (wy != 0.f && n < 4) {
                rows_s[n] = i;
                wts_s[n] = wy;
                ++n;
            }
        }
        for (; n < 4; ++n) rows_s[n] = -1;
    }
    __syncthreads();
    // scale: 2^61 / (sum over the candidate rows of W * max |wy * g|)
    float m = 0.f;
    for (int r = 0; r < 4; ++r) {
        const int i = rows_s[r];
        if (i < 0) continue;
        const float wy = wts_s[r];
        for (int j = tid; j < W; j += 256) m = fmaxf(m, fabsf(wy * __ldg(gplane + (size_t)i * W + j)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) red[tid >> 5] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) m = fmaxf(m, red[k]);
    float* orow = gimg + ((size_t)b * C + c) * HW + (size_t)yp * W;
    if (m == 0.f || !(m < 3.0e38f)) {  // all-zero upstream gradient, or a non-finite one (propagated as NaN)
        const float fill = (m == 0.f) ? 0.f : __int_as_float(0x7fc00000);
        for (int x = tid; x < W; x += 256) orow[x] = fill;
        return;
    }
    const double scale = ldexp(1.0, 61) / ((double)m * 4.0 * (double)W);
    for (int r = 0; r < 4; ++r) {
        const int i = rows_s[r];
        if (i < 0) continue;
        const float wy = wts_s[r];
        for (int j = tid; j < W; j += 256) {
            const float d = __ldg(dplane + (size_t)i * W + j);
            const Axis ax = make_axis(sample_pos(__ldg(lin_x + j), __fdiv_rn(d, (float)W), (float)W), W);
            const float g = wy * __ldg(gplane + (size_t)i * W + j);
            if (ax.v0) atomicAdd(&facc[ax.i0], (unsigned long long)__double2ll_rn((double)(ax.e * g) * scale));
            if (ax.v1) atomicAdd(&facc[ax.i0 + 1], (unsigned long long)__double2ll_rn((double)(ax.w * g) * scale));
        }
    }
    __syncthreads();
    const double inv = 1.0 / scale;
    for (int x = tid; x < W; x += 256) orow[x] = (float)((double)(long long)facc[x] * inv);
}

// ------------------------------------------------------------------------------------------
// a7/a8 fused loss.  grid = (H, B); one CTA per output row; dynamic smem:
//   Rs[ps][Wp] | Ls[ps][Wp]   (Wp = W + 2*(p+1), data starts at column OFF = p+1)
// ------------------------------------------------------------------------------------------
constexpr int kLossThreads = 512;

template <int PS>
__device__ __forceinline__ void patch_taps(const float* __restrict__ rs, const float* __restrict__ ls, int Wp, int ps,
                                           float bx0, float bx1, float gx0, float gx1, float& acc, float& gacc) {
    const int n = PS > 0 ? PS : ps;
    float a0 = 0.f, a1 = 0.f, g0 = 0.f, g1 = 0.f;
#pragma unroll
    for (int ky = 0; ky < n; ++ky) {
        const float* rr = rs + ky * Wp;
        const float* ll = ls + ky * Wp;
        float prev = rr[0];
#pragma unroll
        for (int kx = 0; kx < n; ++kx) {
            const float nxt = rr[kx + 1];
            const float wu = fmaf(bx1, nxt, bx0 * prev);
            const float dwu = fmaf(gx1, nxt, -gx0 * prev);
            const float r = wu - ll[kx];
            if (kx & 1) { a1 = fmaf(r, r, a1); g1 = fmaf(r, dwu, g1); }
            else        { a0 = fmaf(r, r, a0); g0 = fmaf(r, dwu, g0); }
            prev = nxt;
        }
    }
    acc = a0 + a1;
    gacc = g0 + g1;
}

template <int PS>
__global__ void __launch_bounds__(kLossThreads) reproj_loss_kernel(
    const float* __restrict__ tgt, const float* __restrict__ src, const float* __restrict__ disp, float sign,
    const uint8_t* __restrict__ mask, const float* __restrict__ lin_x, const float* __restrict__ lin_y, int ps,
    float* __restrict__ warped, float* __restrict__ gpre, double* __restrict__ partial, int C, int H, int W) {
    extern __shared__ __align__(16) float sm[];
    __shared__ double red[32];
    const int i = blockIdx.x, b = blockIdx.y;
    const int p = (ps - 1) >> 1, OFF = p + 1, Wp = W + 2 * OFF;
    float* Rs = sm;
    float* Ls = sm + (size_t)ps * Wp;
    const size_t HW = (size_t)H * W;
    const size_t rowbase = (size_t)b * HW + (size_t)i * W;

    const Axis ay = make_axis(sample_pos(__ldg(lin_y + i), 0.0f, (float)H), H);
    const float ay0 = ay.v0 ? ay.e : 0.f, ay1 = ay.v1 ? ay.w : 0.f;
    const bool want_warped = (warped != nullptr) && (ps == 1);

    double tot = 0.0, cnt = 0.0;
    for (int c = 0; c < C; ++c) {
        const float* sp = src + ((size_t)b * C + c) * HW;
        const float* tp = tgt + ((size_t)b * C + c) * HW;
        __syncthreads();  // the previous channel's taps are done with the shared rows
        for (int xx = threadIdx.x; xx < Wp; xx += kLossThreads) {
            const int x = xx - OFF;
            const bool inx = (x >= 0) && (x < W);
            // source rows y0-p .. y0+1+p, blended pairwise with the (validity-folded) row weights
            int yy = ay.i0 - p;
            float prev = (inx && yy >= 0 && yy < H) ? __ldg(sp + (size_t)yy * W + x) : 0.f;
            for (int ky = 0; ky < ps; ++ky) {
                ++yy;
                const float nxt = (inx && yy >= 0 && yy < H) ? __ldg(sp + (size_t)yy * W + x) : 0.f;
                Rs[ky * Wp + xx] = fmaf(ay1, nxt, ay0 * prev);
                prev = nxt;
                const int ty = i + ky - p;
                Ls[ky * Wp + xx] = (inx && ty >= 0 && ty < H) ? __ldg(tp + (size_t)ty * W + x) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll 1
        for (int j = threadIdx.x; j < W; j += kLossThreads) {
            const bool m = mask == nullptr ? true : (mask[rowbase + j] != 0);
            float a = 0.f, g = 0.f;
            if (m || want_warped) {
                const float d = sign * disp[rowbase + j];
                const Axis ax = make_axis(sample_pos(__ldg(lin_x + j), __fdiv_rn(d, (float)W), (float)W), W);
                const int x0 = min(max(ax.i0, -1), W - 1);  // outside this range both corners are invalid
                const float bx0 = ax.v0 ? ax.e : 0.f, bx1 = ax.v1 ? ax.w : 0.f;
                const float* rs = Rs + OFF + x0 - p;
                if (m)
                    patch_taps<PS>(rs, Ls + (OFF + j - p), Wp, ps, bx0, bx1, ax.v0 ? 1.f : 0.f, ax.v1 ? 1.f : 0.f, a, g);
                if (want_warped)  // ps == 1: the single tap *is* the warped pixel (the mask does not apply to it)
                    warped[((size_t)b * C + c) * HW + (size_t)i * W + j] = fmaf(bx1, rs[1], bx0 * rs[0]);
            }
            tot += (double)a;
            if (c == 0 && m) cnt += 1.0;
            if (gpre != nullptr) gpre[rowbase + j] = (c == 0 ? 0.f : gpre[rowbase + j]) + g;
        }
    }
    const double bs = block_sum(tot, red);
    const double bc = block_sum(cnt, red);
    if (threadIdx.x == 0) {
        const size_t r = (size_t)b * H + i;
        partial[2 * r] = bs;
        partial[2 * r + 1] = bc;
    }
}

// final reduction of the per-row partials, fixed order => deterministic.  1 CTA.
__global__ void __launch_bounds__(1024) reproj_finalize_kernel(const double* __restrict__ partial, int64_t n, double K,
                                                               float* __restrict__ loss_out,
                                                               double* __restrict__ stats) {
    __shared__ double red[32];
    double s = 0.0, c = 0.0;
    for (int64_t r = threadIdx.x; r < n; r += 1024) { s += partial[2 * r]; c += partial[2 * r + 1]; }
    s = block_sum(s, red);
    c = block_sum(c, red);
    if (threadIdx.x == 0) {
        stats[0] = s;
        stats[1] = c;
        loss_out[0] = (float)(s / (K * c));  // 0/0 = NaN for an empty mask, like F.mse_loss of nothing
    }
}

__global__ void __launch_bounds__(256) reproj_bwd_kernel(const float* __restrict__ gpre, const double* __restrict__ stats,
                                                         const float* __restrict__ gloss, float sign, double K,
                                                         float* __restrict__ gdisp, int64_t n) {
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= n) return;
    const float scale = (float)((double)gloss[0] * (double)sign * 2.0 / (K * stats[1]));
    gdisp[t] = scale * gpre[t];
}

// ------------------------------------------------------------------------------------------
// Fold (overlap-sum) of the warped unfolded planes, cropped: reprojection.py:120-125.
// vis[b,c,y,x] = sum_{ky,kx} Wu[(c,ky,kx)][y+p-ky, x+p-kx].   grid = (H, C, B)
//
// Source-centric + systolic: for tap row ky the contributing source row is i = y+p-ky.  A lane
// owns two adjacent SOURCE pixels j of that row, computes their sampling parameters once and
// walks the ps horizontal taps with a sliding window over the blended source row
// (Wu[kx] = bx0*A[kx] + bx1*A[kx+1]); tap kx of source j lands on output x = j+kx-p, so the
// sum over kx is a diagonal across lanes, accumulated with a shuffle chain
//     acc_s <- Wu_s[k] + acc_{s+1}         (s = source index, one __shfl_down per two taps).
// After ps steps the lane of source s holds the output x = s+p; a pass of one warp covers 64
// sources and yields 64-(ps-1) outputs.  Per tap: 1 LDS + 2 FMA + 1 FADD + 1/2 SHFL.
// smem: Bk[ps][Wp] blended source rows (Wp = W + 2*(p+1), data starts at column OFF = p+1).
// ------------------------------------------------------------------------------------------
constexpr int kFoldThreads = 256;

struct FoldSrc {
    int base;        // index of A[0] in the blended row
    float bx0, bx1;  // corner weights with validity folded in
};

__device__ __forceinline__ FoldSrc fold_src(const float* __restrict__ drow, const float* __restrict__ lin_x, float sign,
                                            int j, int W, int OFF, int p) {
    FoldSrc f;
    f.base = OFF;  // any in-range index; weights are 0
    f.bx0 = 0.f;
    f.bx1 = 0.f;
    if (j >= 0 && j < W) {
        const Axis ax = make_axis(sample_pos(__ldg(lin_x + j), __fdiv_rn(sign * __ldg(drow + j), (float)W), (float)W), W);
        f.base = OFF + min(max(ax.i0, -1), W - 1) - p;
        f.bx0 = ax.v0 ? ax.e : 0.f;
        f.bx1 = ax.v1 ? ax.w : 0.f;
    }
    return f;
}

template <int PS>
__global__ void __launch_bounds__(kFoldThreads) patch_fold_systolic_kernel(const float* __restrict__ src,
                                                                          const float* __restrict__ disp, float sign,
                                                                          const float* __restrict__ lin_x,
                                                                          const float* __restrict__ lin_y,
                                                                          float* __restrict__ vis, int C, int H, int W) {
    extern __shared__ __align__(16) float sm[];
    constexpr int p = (PS - 1) / 2, OFF = p + 1;
    constexpr int STEP = 64 - (PS - 1);  // outputs per warp pass
    const int y = blockIdx.x, c = blockIdx.y, b = blockIdx.z;
    const int Wp = W + 2 * OFF;
    const size_t HW = (size_t)H * W;
    const float* sp = src + ((size_t)b * C + c) * HW;
    const float* dp = disp + (size_t)b * HW;

    // blended source rows, one per tap row
    for (int ky = 0; ky < PS; ++ky) {
        const int i = y + p - ky;
        const bool rowok = (i >= 0) && (i < H);
        float ay0 = 0.f, ay1 = 0.f;
        int r0 = 0;
        if (rowok) {
            const Axis ay = make_axis(sample_pos(__ldg(lin_y + i), 0.0f, (float)H), H);
            ay0 = ay.v0 ? ay.e : 0.f;
            ay1 = ay.v1 ? ay.w : 0.f;
            r0 = ay.i0 + ky - p;  // image row read by corner y0 of tap row ky
        }
        for (int xx = threadIdx.x; xx < Wp; xx += kFoldThreads) {
            const int x = xx - OFF;
            float v = 0.f;
            if (rowok && x >= 0 && x < W) {
                const float a = (r0 >= 0 && r0 < H) ? __ldg(sp + (size_t)r0 * W + x) : 0.f;
                const float bb = (r0 + 1 >= 0 && r0 + 1 < H) ? __ldg(sp + (size_t)(r0 + 1) * W + x) : 0.f;
                v = fmaf(ay1, bb, ay0 * a);
            }
            sm[ky * Wp + xx] = v;
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = kFoldThreads >> 5;
    const int npass = (W + STEP - 1) / STEP;
    float* orow = vis + ((size_t)b * C + c) * HW + (size_t)y * W;
    for (int q = warp; q < npass; q += nwarps) {
        const int x_out = q * STEP + 2 * lane;  // output column of this lane's first source
        const int j0 = x_out - p;               // its two sources
        float tot0 = 0.f, tot1 = 0.f;
#pragma unroll 1
        for (int ky = 0; ky < PS; ++ky) {
            const int i = y + p - ky;
            if (i < 0 || i >= H) continue;  // warp-uniform
            const float* drow = dp + (size_t)i * W;
            const FoldSrc s0 = fold_src(drow, lin_x, sign, j0, W, OFF, p);
            const FoldSrc s1 = fold_src(drow, lin_x, sign, j0 + 1, W, OFF, p);
            const float* r0 = sm + ky * Wp + s0.base;
            const float* r1 = sm + ky * Wp + s1.base;
            float a0 = r0[0], a1 = r1[0];
            float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
            for (int k = 0; k < PS; ++k) {
                const float n0 = r0[k + 1], n1 = r1[k + 1];
                const float w0 = fmaf(s0.bx1, n0, s0.bx0 * a0);
                const float w1 = fmaf(s1.bx1, n1, s1.bx0 * a1);
                const float from_next = __shfl_down_sync(0xffffffffu, acc0, 1);
                acc0 = w0 + acc1;
                acc1 = w1 + from_next;
                a0 = n0;
                a1 = n1;
            }
            tot0 += acc0;
            tot1 += acc1;
        }
        if (2 * lane < STEP) {
            if (x_out < W) orow[x_out] = tot0;
            if (x_out + 1 < W) orow[x_out + 1] = tot1;
        }
    }
}

// generic fallback (any odd ps): output-centric gather.  grid = (H, C, B)
// smem: Bk[ps][Wp] blended source rows | XS[ps][W] sample x of source row i = y+p-ky
__global__ void __launch_bounds__(512) patch_fold_kernel(const float* __restrict__ src,
                                                                 const float* __restrict__ disp, float sign,
                                                                 const float* __restrict__ lin_x,
                                                                 const float* __restrict__ lin_y, int ps,
                                                                 float* __restrict__ vis, int C, int H, int W) {
    constexpr int kT = 512;
    extern __shared__ __align__(16) float sm[];
    const int y = blockIdx.x, c = blockIdx.y, b = blockIdx.z;
    const int p = (ps - 1) >> 1, OFF = p + 1, Wp = W + 2 * OFF;
    float* Bk = sm;
    float* XS = sm + (size_t)ps * Wp;
    const size_t HW = (size_t)H * W;
    const float* sp = src + ((size_t)b * C + c) * HW;
    const float* dp = disp + (size_t)b * HW;

    for (int ky = 0; ky < ps; ++ky) {
        const int i = y + p - ky;  // source (unfolded-plane) row contributing through tap row ky
        const bool rowok = (i >= 0) && (i < H);
        float ay0 = 0.f, ay1 = 0.f;
        int r0 = 0;
        if (rowok) {
            const Axis ay = make_axis(sample_pos(__ldg(lin_y + i), 0.0f, (float)H), H);
            ay0 = ay.v0 ? ay.e : 0.f;
            ay1 = ay.v1 ? ay.w : 0.f;
            r0 = ay.i0 + ky - p;  // image row read by corner y0 of tap row ky
        }
        for (int xx = threadIdx.x; xx < Wp; xx += kT) {
            const int x = xx - OFF;
            float v = 0.f;
            if (rowok && x >= 0 && x < W) {
                const float a = (r0 >= 0 && r0 < H) ? __ldg(sp + (size_t)r0 * W + x) : 0.f;
                const float bb = (r0 + 1 >= 0 && r0 + 1 < H) ? __ldg(sp + (size_t)(r0 + 1) * W + x) : 0.f;
                v = fmaf(ay1, bb, ay0 * a);
            }
            Bk[ky * Wp + xx] = v;
        }
        for (int j = threadIdx.x; j < W; j += kT) {
            float xs = -8.0f;  // floor = -8: both corners invalid
            if (rowok) xs = sample_pos(__ldg(lin_x + j), __fdiv_rn(sign * dp[(size_t)i * W + j], (float)W), (float)W);
            XS[ky * W + j] = xs;
        }
    }
    __syncthreads();
    for (int x = threadIdx.x; x < W; x += kT) {
        float a0 = 0.f, a1 = 0.f;
        for (int ky = 0; ky < ps; ++ky) {
            const float* bk = Bk + ky * Wp + OFF;
            const float* xs = XS + ky * W;
            for (int kx = 0; kx < ps; ++kx) {
                const int j = x + p - kx;
                if (j < 0 || j >= W) continue;
                const Axis ax = make_axis(xs[j], W);
                const int x0 = min(max(ax.i0, -1), W - 1) + kx - p;
                const float t = fmaf(ax.v1 ? ax.w : 0.f, bk[x0 + 1], (ax.v0 ? ax.e : 0.f) * bk[x0]);
                if (kx & 1) a1 += t; else a0 += t;
            }
        }
        vis[((size_t)b * C + c) * HW + (size_t)y * W + x] = a0 + a1;
    }
}

}  // namespace az

using namespace az;

namespace az {
// patch_loss_fold.cu: loss + Fold image in one pass; AZ_ERR_BAD_ARG when the shape does not fit
int plf_dispatch(const float* tgt, const float* src, const float* disp, float sign, const uint8_t* mask,
                 const float* lin_x, const float* lin_y, int ps, float* vis, float* gpre, double* partial, int B, int C,
                 int H, int W, int* nbands_out, cudaStream_t st);
// patch_loss_fold_v3.cu: round-2 form of the same pass (lanes over tap rows: conflict-free window loads)
int plf3_dispatch(const float* tgt, const float* src, const float* disp, float sign, const uint8_t* mask,
                  const float* lin_x, const float* lin_y, int ps, float* vis, float* gpre, double* partial, int B, int C,
                  int H, int W, int* nbands_out, cudaStream_t st);
}

extern "C" int az_warp_fwd(const float* img, const float* disp, const float* lin_x, const float* lin_y, float* out,
                           int64_t B, int64_t C, int64_t H, int64_t W, void* stream) {
    if (!img || !disp || !lin_x || !lin_y || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0) return AZ_ERR_BAD_ARG;
    if (H > 65535 || B > 65535 || H * W >= (1ll << 31)) return AZ_ERR_BAD_ARG;
    if (W % 4 == 0 && aligned16(img) && aligned16(disp) && aligned16(out) && aligned16(lin_x)) {
        dim3 grid4((unsigned)ceil_div(H * W / 4, 256), (unsigned)B);
        warp_fwd4_kernel<<<grid4, 256, 0, (cudaStream_t)stream>>>(img, disp, lin_x, lin_y, out, (int)C, (int)H, (int)W);
        AZ_LAUNCH_CHECK();
        return 0;
    }
    dim3 grid((unsigned)ceil_div(W, 256), (unsigned)H, (unsigned)B);
    warp_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, disp, lin_x, lin_y, out, (int)C, (int)H, (int)W);
    AZ_LAUNCH_CHECK();
    return 0;
}

extern "C" int az_warp_bwd(const float* img, const float* disp, const float* lin_x, const float* lin_y,
                           const float* gout, float* gimg, float* gdisp, int64_t B, int64_t C, int64_t H, int64_t W,
                           void* stream) {
    if (!img || !disp || !lin_x || !lin_y || !gout || B <= 0 || C <= 0 || H <= 0 || W <= 0) return AZ_ERR_BAD_ARG;
    if (H > 65535 || B > 65535 || H * W >= (1ll << 31)) return AZ_ERR_BAD_ARG;
    if (!gimg && !gdisp) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool det_img = gimg != nullptr && tuning("AZ_WARP_BWD_IMG", 1) == 1 && C <= 65535 && W * 8 <= 200 * 1024;
    if (gimg != nullptr && !det_img) {  // round 1's float-atomic path accumulates: it needs zeros to start from
        cudaError_t e = cudaMemsetAsync(gimg, 0, (size_t)B * C * H * W * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
    }
    if (gdisp != nullptr || (gimg != nullptr && !det_img)) {
        dim3 grid((unsigned)ceil_div(W, 256), (unsigned)H, (unsigned)B);
        warp_bwd_kernel<<<grid, 256, 0, st>>>(img, disp, lin_x, lin_y, gout, det_img ? nullptr : gimg, gdisp, (int)C, (int)H,
                                              (int)W);
        AZ_LAUNCH_CHECK();
    }
    if (det_img) {
        const size_t smem = (size_t)W * 8;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(warp_bwd_img_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        dim3 grid((unsigned)H, (unsigned)C, (unsigned)B);
        warp_bwd_img_kernel<<<grid, 256, smem, st>>>(disp, lin_x, lin_y, gout, gimg, (int)C, (int)H, (int)W);
        AZ_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int64_t az_reproj_workspace_bytes(int64_t B, int64_t H) { return B * H * 2 * (int64_t)sizeof(double); }

template <int PS>
static int launch_loss(const float* tgt, const float* src, const float* disp, float sign, const uint8_t* mask,
                       const float* lin_x, const float* lin_y, int ps, float* warped, float* gpre, double* partial,
                       int B, int C, int H, int W, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(reproj_loss_kernel<PS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(220 * 1024));
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)H, (unsigned)B);
    reproj_loss_kernel<PS><<<grid, kLossThreads, smem, st>>>(tgt, src, disp, sign, mask, lin_x, lin_y, ps, warped, gpre,
                                                            partial, C, H, W);
    return (int)cudaGetLastError();
}

extern "C" int az_reproj_loss_fwd(const float* tgt, const float* src, const float* disp, float sign,
                                  const uint8_t* mask, const float* lin_x, const float* lin_y, int64_t ps,
                                  float* warped, float* gpre, float* loss_out, double* stats, void* workspace,
                                  int64_t B, int64_t C, int64_t H, int64_t W, void* stream) {
    if (!tgt || !src || !disp || !lin_x || !lin_y || !loss_out || !stats || !workspace) return AZ_ERR_BAD_ARG;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || ps < 1 || (ps % 2) == 0) return AZ_ERR_BAD_ARG;
    if (B > 65535 || H * W >= (1ll << 31)) return AZ_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int p = (int)(ps - 1) / 2;
    const size_t Wp = (size_t)W + 2 * (p + 1);
    const size_t smem = 2 * (size_t)ps * Wp * sizeof(float);
    double* partial = (double*)workspace;
    int rc;
    // round 1's kernels stage full rows of width W + ps + 1: beyond ~2600 columns at ps = 11 only the x-tiled round-2
    // kernel (loss + Fold image) can run
    const bool r1_fits = smem <= 220 * 1024;
    if (!r1_fits && !(warped != nullptr && ps != 1)) return AZ_ERR_BAD_ARG;
    if (warped != nullptr && ps != 1) {
        // get_reproj_error_patch: loss + Fold image in one pass (patch_loss_fold.cu) when the shape fits its
        // shared-memory plan, else the stand-alone loss kernel followed by the stand-alone Fold
        cudaError_t e = cudaMemsetAsync(warped, 0, (size_t)B * C * H * W * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
        int nbands = 0;
        rc = AZ_ERR_BAD_ARG;
        if (H <= (1 << 24) && tuning("AZ_PATCH_IMPL", 1) == 1)
            rc = plf3_dispatch(tgt, src, disp, sign, mask, lin_x, lin_y, (int)ps, warped, gpre, partial, (int)B, (int)C,
                               (int)H, (int)W, &nbands, st);
        if (rc == AZ_ERR_BAD_ARG && H <= (1 << 24) && r1_fits)
            rc = plf_dispatch(tgt, src, disp, sign, mask, lin_x, lin_y, (int)ps, warped, gpre, partial, (int)B, (int)C,
                              (int)H, (int)W, &nbands, st);
        if (rc == 0) {
            reproj_finalize_kernel<<<1, 1024, 0, st>>>(partial, B * nbands, (double)(C * ps * ps), loss_out, stats);
            AZ_LAUNCH_CHECK();
            return 0;
        }
        if (rc != AZ_ERR_BAD_ARG || !r1_fits) return rc;
        rc = az_reproj_loss_fwd(tgt, src, disp, sign, mask, lin_x, lin_y, ps, nullptr, gpre, loss_out, stats, workspace,
                                B, C, H, W, stream);
        if (rc != 0) return rc;
        return az_patch_fold(src, disp, sign, lin_x, lin_y, ps, warped, B, C, H, W, stream);
    }
    if (ps == 1 && W % 4 == 0 && H <= 65535 && aligned16(tgt) && aligned16(src) && aligned16(disp) && aligned16(lin_x) &&
        (warped == nullptr || aligned16(warped)) && (gpre == nullptr || aligned16(gpre)) &&
        (mask == nullptr || (reinterpret_cast<uintptr_t>(mask) & 3u) == 0)) {
        dim3 grid((unsigned)H, (unsigned)B);
        if (tuning("AZ_PS1_MINB", 3) == 4)
            reproj_ps1_kernel<4><<<grid, 256, 0, st>>>(tgt, src, disp, sign, mask, lin_x, lin_y, warped, gpre, partial, (int)C,
                                                      (int)H, (int)W);
        else
            reproj_ps1_kernel<3><<<grid, 256, 0, st>>>(tgt, src, disp, sign, mask, lin_x, lin_y, warped, gpre, partial, (int)C,
                                                      (int)H, (int)W);
        AZ_LAUNCH_CHECK();
        reproj_finalize_kernel<<<1, 1024, 0, st>>>(partial, B * H, (double)C, loss_out, stats);
        AZ_LAUNCH_CHECK();
        return 0;
    }
#define AZ_LOSS_CASE(N)                                                                                              \
    case N:                                                                                                          \
        rc = launch_loss<N>(tgt, src, disp, sign, mask, lin_x, lin_y, (int)ps, warped, gpre, partial, (int)B, (int)C, \
                            (int)H, (int)W, smem, st);                                                               \
        break;
    switch (ps) {
        AZ_LOSS_CASE(1)
        AZ_LOSS_CASE(3)
        AZ_LOSS_CASE(5)
        AZ_LOSS_CASE(7)
        AZ_LOSS_CASE(9)
        AZ_LOSS_CASE(11)
        default:
            rc = launch_loss<0>(tgt, src, disp, sign, mask, lin_x, lin_y, (int)ps, warped, gpre, partial, (int)B, (int)C,
                                (int)H, (int)W, smem, st);
    }
#undef AZ_LOSS_CASE
    if (rc != 0) return rc;
    reproj_finalize_kernel<<<1, 1024, 0, st>>>(partial, B * H, (double)(C * ps * ps), loss_out, stats);
    AZ_LAUNCH_CHECK();
    return 0;
}

extern "C" int az_reproj_loss_bwd(const float* gpre, const double* stats, const float* gloss, float sign,
                                  float* gdisp, int64_t B, int64_t C, int64_t H, int64_t W, int64_t ps,
                                  void* stream) {
    if (!gpre || !stats || !gloss || !gdisp || B <= 0 || C <= 0 || H <= 0 || W <= 0 || ps < 1) return AZ_ERR_BAD_ARG;
    const int64_t n = B * H * W;
    reproj_bwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(gpre, stats, gloss, sign,
                                                                                    (double)(C * ps * ps), gdisp, n);
    AZ_LAUNCH_CHECK();
    return 0;
}

template <int PS>
static int launch_fold(const float* src, const float* disp, float sign, const float* lin_x, const float* lin_y, float* vis,
                       int B, int C, int H, int W, cudaStream_t st) {
    const size_t Wp = (size_t)W + 2 * ((PS - 1) / 2 + 1);
    const size_t smem = (size_t)PS * Wp * sizeof(float);
    if (smem > 220 * 1024) return AZ_ERR_BAD_ARG;
    cudaError_t e = cudaFuncSetAttribute(patch_fold_systolic_kernel<PS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(220 * 1024));
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)H, (unsigned)C, (unsigned)B);
    patch_fold_systolic_kernel<PS><<<grid, kFoldThreads, smem, st>>>(src, disp, sign, lin_x, lin_y, vis, C, H, W);
    return (int)cudaGetLastError();
}

extern "C" int az_patch_fold(const float* src, const float* disp, float sign, const float* lin_x, const float* lin_y,
                             int64_t ps, float* vis, int64_t B, int64_t C, int64_t H, int64_t W, void* stream) {
    if (!src || !disp || !lin_x || !lin_y || !vis || B <= 0 || C <= 0 || H <= 0 || W <= 0 || ps < 1 || (ps % 2) == 0)
        return AZ_ERR_BAD_ARG;
    if (B > 65535 || C > 65535 || H * W >= (1ll << 31)) return AZ_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    switch (ps) {
        case 1: return launch_fold<1>(src, disp, sign, lin_x, lin_y, vis, (int)B, (int)C, (int)H, (int)W, st);
        case 3: return launch_fold<3>(src, disp, sign, lin_x, lin_y, vis, (int)B, (int)C, (int)H, (int)W, st);
        case 5: return launch_fold<5>(src, disp, sign, lin_x, lin_y, vis, (int)B, (int)C, (int)H, (int)W, st);
        case 7: return launch_fold<7>(src, disp, sign, lin_x, lin_y, vis, (int)B, (int)C, (int)H, (int)W, st);
        case 9: return launch_fold<9>(src, disp, sign, lin_x, lin_y, vis, (int)B, (int)C, (int)H, (int)W, st);
        case 11: return launch_fold<11>(src, disp, sign, lin_x, lin_y, vis, (int)B, (int)C, (int)H, (int)W, st);
        case 13: return launch_fold<13>(src, disp, sign, lin_x, lin_y, vis, (int)B, (int)C, (int)H, (int)W, st);
        default: break;
    }
    const int p = (int)(ps - 1) / 2;
    const size_t Wp = (size_t)W + 2 * (p + 1);
    const size_t smem = (size_t)ps * (Wp + W) * sizeof(float);
    if (smem > 220 * 1024) return AZ_ERR_BAD_ARG;
    cudaError_t e = cudaFuncSetAttribute(patch_fold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(220 * 1024));
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)H, (unsigned)C, (unsigned)B);
    patch_fold_kernel<<<grid, 512, smem, st>>>(src, disp, sign, lin_x, lin_y, (int)ps, vis, (int)C, (int)H, (int)W);
    AZ_LAUNCH_CHECK();
    return 0;
}
