// Fused soft-argmin (softmax over D + disparity regression), forward and backward
// (SURVEY.md §8a rows a4/a5).
// Reference: /root/reference/nets/psmnet/psmnet.py:200-201, 204-205, 212-217 and
//            /root/reference/nets/psmnet/psmnet_submodule.py:80-89.
//
// cost is [B,D,H,W] with D the slowest in-image axis (stride H*W), so the reduction over D is
// a per-thread serial loop over planes while a warp reads 512 contiguous bytes of one plane per
// request (128-bit loads, 8 planes in flight per thread).  One pass over the logits: online
// softmax in chunks of 8 planes -- chunk max, one rescale per chunk (skipped when the running
// max does not move), tree-summed chunk, running sums kept in double so the result stays well
// inside the 1e-4 px gate for flat distributions.  The forward also emits the log-sum-exp so
// the backward is a single read-modify-write pass with no second reduction.
#include "common.cuh"

namespace az {

constexpr int kSAThreads = 256;
constexpr int kChunk = 8;
constexpr int kBwdDGroup = 16;  // planes per backward CTA (multiple of kChunk)
constexpr float kLog2e = 1.4426950408889634f;

template <int V> struct Vec;
template <> struct Vec<4> {
    using T = float4;
    static __device__ __forceinline__ void unpack(const float4& v, float* f) { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
    static __device__ __forceinline__ float4 pack(const float* f) { return make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct Vec<1> {
    using T = float;
    static __device__ __forceinline__ void unpack(const float& v, float* f) { f[0] = v; }
    static __device__ __forceinline__ float pack(const float* f) { return f[0]; }
};

// One thread owns V adjacent pixels.  n_vec = B*H*W/V work items; plane = H*W (floats).
template <int V>
__global__ void __launch_bounds__(kSAThreads) soft_argmin_fwd_kernel(const float* __restrict__ cost,
                                                                    float* __restrict__ disp,
                                                                    float* __restrict__ lse, int D,
                                                                    int64_t plane, int64_t n_vec) {
    using VT = typename Vec<V>::T;
    const int64_t t = (int64_t)blockIdx.x * kSAThreads + threadIdx.x;
    if (t >= n_vec) return;
    const int64_t pix = t * V;
    const int64_t b = pix / plane, hw = pix - b * plane;
    const float* src = cost + b * D * plane + hw;

    // m = running max.  Exponents are formed as (x - m) * log2e: the subtraction is exact for the
    // entries that carry weight (x close to m), so accuracy does not depend on the magnitude of the
    // logits, and every rounding (term exponent, rescale exponent) is relative to a SMALL number, which
    // keeps old and new partial sums consistent across rescales.
    float m[V];
    double s[V], ws[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { m[j] = -INFINITY; s[j] = 0.0; ws[j] = 0.0; }

    for (int d0 = 0; d0 < D; d0 += kChunk) {
        float x[kChunk][V];
#pragma unroll
        for (int k = 0; k < kChunk; ++k) {
            if (d0 + k < D) {
                const VT v = ld_stream(reinterpret_cast<const VT*>(src + (int64_t)(d0 + k) * plane));
                Vec<V>::unpack(v, x[k]);
            } else {
#pragma unroll
                for (int j = 0; j < V; ++j) x[k][j] = -INFINITY;
            }
        }
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float cm = x[0][j];
#pragma unroll
            for (int k = 1; k < kChunk; ++k) cm = fmaxf(cm, x[k][j]);
            if (cm > m[j]) {  // rescale the running sums (exp2(-inf) = 0 on the first chunk)
                const double sc = (double)fast_ex2((m[j] - cm) * kLog2e);
                s[j] *= sc;
                ws[j] *= sc;
                m[j] = cm;
            }
            const float ref = (m[j] == -INFINITY) ? 0.f : m[j];  // leading -inf planes contribute 0
            float e[kChunk];
#pragma unroll
            for (int k = 0; k < kChunk; ++k) e[k] = fast_ex2((x[k][j] - ref) * kLog2e);
            // tree sums of the chunk in fp32, running sums in fp64
            const float s01 = e[0] + e[1], s23 = e[2] + e[3], s45 = e[4] + e[5], s67 = e[6] + e[7];
            const float w01 = e[1], w23 = fmaf(e[3], 3.f, e[2] * 2.f);
            const float w45 = fmaf(e[5], 5.f, e[4] * 4.f), w67 = fmaf(e[7], 7.f, e[6] * 6.f);
            const float cs = (s01 + s23) + (s45 + s67);
            const float cw = (w01 + w23) + (w45 + w67);  // sum_k k*e[k], k local
            s[j] += (double)cs;
            ws[j] += (double)cw + (double)d0 * (double)cs;
        }
    }
    float o[V], l[V];
#pragma unroll
    for (int j = 0; j < V; ++j) o[j] = (float)(ws[j] / s[j]);
    *reinterpret_cast<VT*>(disp + pix) = Vec<V>::pack(o);
    if (lse != nullptr) {
        // saved for the backward: plane 0 = max logit m, plane 1 = log2(sum_d 2^((x_d - m)*log2e));
        // the backward forms p_d = 2^((x_d - m)*log2e - plane1) with the same exact subtraction.
        const int64_t total = n_vec * V;
        *reinterpret_cast<VT*>(lse + pix) = Vec<V>::pack(m);
#pragma unroll
        for (int j = 0; j < V; ++j) l[j] = log2f((float)s[j]);
        *reinterpret_cast<VT*>(lse + total + pix) = Vec<V>::pack(l);
    }
}

template <int V>
__global__ void __launch_bounds__(kSAThreads) soft_argmin_bwd_kernel(const float* __restrict__ cost,
                                                                    const float* __restrict__ disp,
                                                                    const float* __restrict__ lse,
                                                                    const float* __restrict__ gdisp,
                                                                    float* __restrict__ gcost, int D,
                                                                    int64_t plane, int64_t n_vec, int dgroup) {
    // grid.y splits the D planes into groups of `dgroup`: short-lived CTAs keep the store half of this
    // read-modify-write stream nearer the write-only ceiling (benchmarks/micro/store_patterns.cu)
    using VT = typename Vec<V>::T;
    const int64_t t = (int64_t)blockIdx.x * kSAThreads + threadIdx.x;
    if (t >= n_vec) return;
    const int64_t pix = t * V;
    const int64_t b = pix / plane, hw = pix - b * plane;
    const float* src = cost + b * D * plane + hw;
    float* dst = gcost + b * D * plane + hw;
    const int dbeg = blockIdx.y * dgroup, dend = min(D, dbeg + dgroup);
    float o[V], mx[V], lb[V], g[V];
    Vec<V>::unpack(*reinterpret_cast<const VT*>(disp + pix), o);
    Vec<V>::unpack(*reinterpret_cast<const VT*>(lse + pix), mx);
    Vec<V>::unpack(*reinterpret_cast<const VT*>(lse + n_vec * V + pix), lb);
    Vec<V>::unpack(*reinterpret_cast<const VT*>(gdisp + pix), g);
    for (int d0 = dbeg; d0 < dend; d0 += kChunk) {
        VT v[kChunk];
#pragma unroll
        for (int k = 0; k < kChunk; ++k)
            if (d0 + k < dend) v[k] = ld_stream(reinterpret_cast<const VT*>(src + (int64_t)(d0 + k) * plane));
#pragma unroll
        for (int k = 0; k < kChunk; ++k) {
            if (d0 + k < dend) {
                float x[V], r[V];
                Vec<V>::unpack(v[k], x);
                const float dd = (float)(d0 + k);
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const float p = fast_ex2(fmaf(x[j] - mx[j], kLog2e, -lb[j]));
                    r[j] = p * (dd - o[j]) * g[j];
                }
                st_stream(reinterpret_cast<VT*>(dst + (int64_t)(d0 + k) * plane), Vec<V>::pack(r));
            }
        }
    }
}

}  // namespace az

using namespace az;

extern "C" int az_soft_argmin_fwd(const float* cost, float* disp, float* lse, int64_t B, int64_t D, int64_t H,
                                  int64_t W, void* stream) {
    if (!cost || !disp || B <= 0 || D <= 0 || H <= 0 || W <= 0) return AZ_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t plane = H * W, total = B * plane;
    const bool vec = (plane % 4 == 0) && aligned16(cost) && aligned16(disp) && (lse == nullptr || aligned16(lse));
    if (vec) {
        const int64_t n = total / 4;
        soft_argmin_fwd_kernel<4><<<(unsigned)ceil_div(n, kSAThreads), kSAThreads, 0, st>>>(cost, disp, lse, (int)D,
                                                                                           plane, n);
    } else {
        soft_argmin_fwd_kernel<1><<<(unsigned)ceil_div(total, kSAThreads), kSAThreads, 0, st>>>(cost, disp, lse, (int)D,
                                                                                               plane, total);
    }
    AZ_LAUNCH_CHECK();
    return 0;
}

extern "C" int az_soft_argmin_bwd(const float* cost, const float* disp, const float* lse, const float* gdisp,
                                  float* gcost, int64_t B, int64_t D, int64_t H, int64_t W, void* stream) {
    if (!cost || !disp || !lse || !gdisp || !gcost || B <= 0 || D <= 0 || H <= 0 || W <= 0) return AZ_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t plane = H * W, total = B * plane;
    const bool vec = (plane % 4 == 0) && aligned16(cost) && aligned16(disp) && aligned16(lse) && aligned16(gdisp) &&
                     aligned16(gcost);
    const int dgroup = kBwdDGroup;
    const unsigned ngroups = (unsigned)ceil_div(D, dgroup);
    if (ngroups > 65535) return AZ_ERR_BAD_ARG;
    if (vec) {
        const int64_t n = total / 4;
        dim3 grid((unsigned)ceil_div(n, kSAThreads), ngroups);
        soft_argmin_bwd_kernel<4><<<grid, kSAThreads, 0, st>>>(cost, disp, lse, gdisp, gcost, (int)D, plane, n, dgroup);
    } else {
        dim3 grid((unsigned)ceil_div(total, kSAThreads), ngroups);
        soft_argmin_bwd_kernel<1><<<grid, kSAThreads, 0, st>>>(cost, disp, lse, gdisp, gcost, (int)D, plane, total,
                                                               dgroup);
    }
    AZ_LAUNCH_CHECK();
    return 0;
}
