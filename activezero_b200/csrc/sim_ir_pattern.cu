// Sim-domain IR pattern extraction on the GPU (a "next" row).
// Reference: /root/reference/datasets/dataset_utils.py:12-17 (get_ir_pattern) and :33-46
// (get_smoothed_ir_pattern2) -- per-sample CPU work inside the DataLoader workers
// (datasets/messytable.py:221-232, 408-428): |ir - no_ir|, min-max normalise, subtract the
// cv2.resize(INTER_AREA) down(//ks)-then-up resampling, threshold.
//
// float64 like numpy.  The two cv2.resize calls are restated from OpenCV's imgproc/resize.cpp in its own
// summation order (oracle/stereo_oracle.py: resize_area_down / resize_area_up are bit-identical to cv2 4.13):
//   shrink : integer ratio -> resizeAreaFast_ (block taps row-major, four at a time, times float 1/area);
//            otherwise computeResizeAreaTab fractional-area weights (stored as float), x then y accumulation;
//   enlarge: the two-tap linear resize with area-mode coefficients fx = (dx+1) - (sx+1)*inv_scale (float).
// No FMA contraction anywhere (_rn intrinsics) so the threshold decision matches bit for bit.
#include "common.cuh"

namespace az {

__global__ void __launch_bounds__(256) sip_init_kernel(unsigned long long* __restrict__ minmax, int B) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t < B) {
        minmax[2 * t] = 0x7FF0000000000000ull;  // +inf
        minmax[2 * t + 1] = 0ull;               // +0.0
    }
}

__device__ __forceinline__ double sip_load(const void* p, int is_u8, size_t o) {
    return is_u8 ? (double)reinterpret_cast<const uint8_t*>(p)[o] / 255.0 : reinterpret_cast<const double*>(p)[o];
}

// diff = |ir - img| and per-image min / max.  grid = (ceil(HW/256), B)
__global__ void __launch_bounds__(256) sip_diff_kernel(const void* __restrict__ ir, const void* __restrict__ img,
                                                       int is_u8, double* __restrict__ diff,
                                                       unsigned long long* __restrict__ minmax, int64_t HW) {
    const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int b = blockIdx.y;
    double mn = __longlong_as_double(0x7FF0000000000000ll), mx = 0.0;
    if (p < HW) {
        const size_t o = (size_t)b * HW + p;
        const double v = fabs(__dsub_rn(sip_load(ir, is_u8, o), sip_load(img, is_u8, o)));
        diff[o] = v;
        mn = v;
        mx = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&minmax[2 * b], (unsigned long long)__double_as_longlong(mn));
        atomicMax(&minmax[2 * b + 1], (unsigned long long)__double_as_longlong(mx));
    }
}

struct Norm {
    double mn, range;
    __device__ __forceinline__ double operator()(double v) const { return __dsub_rn(v, mn) / range; }
};

// One axis of computeResizeAreaTab for destination index d: up to `n` taps (index, float weight).
struct AreaTaps {
    int first;      // source index of the optional leading partial tap, or -1
    float a_first;
    int s1, s2;     // full taps [s1, s2)
    float a_full;
    int last;       // source index of the optional trailing partial tap, or -1
    float a_last;
};

__device__ __forceinline__ AreaTaps area_taps(int d, int ssize, double scale) {
    AreaTaps t;
    const double f1 = __dmul_rn((double)d, scale);
    const double f2 = __dadd_rn(f1, scale);
    const double cell = fmin(scale, __dsub_rn((double)ssize, f1));
    int s1 = (int)ceil(f1), s2 = (int)floor(f2);
    s2 = min(s2, ssize - 1);
    s1 = min(s1, s2);
    t.first = -1;
    t.last = -1;
    t.a_first = t.a_last = 0.f;
    if (__dsub_rn((double)s1, f1) > 1e-3) {
        t.first = s1 - 1;
        t.a_first = (float)(__dsub_rn((double)s1, f1) / cell);
    }
    t.s1 = s1;
    t.s2 = s2;
    t.a_full = (float)(1.0 / cell);
    if (__dsub_rn(f2, (double)s2) > 1e-3) {
        t.last = s2;
        t.a_last = (float)(fmin(fmin(__dsub_rn(f2, (double)s2), 1.0), cell) / cell);
    }
    return t;
}

// shrink [H,W] -> [hs,ws].  grid = (ceil(ws*hs/128), B)
__global__ void __launch_bounds__(128) sip_shrink_kernel(const double* __restrict__ diff,
                                                         const unsigned long long* __restrict__ minmax,
                                                         double* __restrict__ small, int H, int W, int hs, int ws) {
    const int t = blockIdx.x * 128 + threadIdx.x;
    if (t >= hs * ws) return;
    const int b = blockIdx.y, dy = t / ws, dx = t - dy * ws;
    Norm nz;
    nz.mn = __longlong_as_double((long long)minmax[2 * b]);
    nz.range = __dsub_rn(__longlong_as_double((long long)minmax[2 * b + 1]), nz.mn);
    const double* src = diff + (size_t)b * H * W;
    const double sx_f = 1.0 / ((double)ws / (double)W), sy_f = 1.0 / ((double)hs / (double)H);
    const int ix = (int)rint(sx_f), iy = (int)rint(sy_f);
    double result;
    if (fabs(sx_f - ix) < 2.220446049250313e-16 && fabs(sy_f - iy) < 2.220446049250313e-16) {
        // resizeAreaFast_: row-major taps, four at a time, times float(1/area)
        const double scale = (double)(1.0f / (float)(ix * iy));
        const int area = ix * iy;
        double sum = 0.0;
        int k = 0;
        auto tap = [&](int kk) { return nz(src[(size_t)(dy * iy + kk / ix) * W + dx * ix + kk % ix]); };
        for (; k + 4 <= area; k += 4)
            sum = __dadd_rn(sum, __dadd_rn(__dadd_rn(__dadd_rn(tap(k), tap(k + 1)), tap(k + 2)), tap(k + 3)));
        for (; k < area; ++k) sum = __dadd_rn(sum, tap(k));
        result = __dmul_rn(sum, scale);
    } else {
        const AreaTaps tx = area_taps(dx, W, sx_f), ty = area_taps(dy, H, sy_f);
        auto hrow = [&](int sy) {
            const double* r = src + (size_t)sy * W;
            double buf = 0.0;
            if (tx.first >= 0) buf = __dadd_rn(buf, __dmul_rn(nz(r[tx.first]), (double)tx.a_first));
            for (int sx = tx.s1; sx < tx.s2; ++sx) buf = __dadd_rn(buf, __dmul_rn(nz(r[sx]), (double)tx.a_full));
            if (tx.last >= 0) buf = __dadd_rn(buf, __dmul_rn(nz(r[tx.last]), (double)tx.a_last));
            return buf;
        };
        bool have = false;
        double sum = 0.0;
        auto acc = [&](int sy, float beta) {
            const double v = __dmul_rn((double)beta, hrow(sy));
            sum = have ? __dadd_rn(sum, v) : v;
            have = true;
        };
        if (ty.first >= 0) acc(ty.first, ty.a_first);
        for (int sy = ty.s1; sy < ty.s2; ++sy) acc(sy, ty.a_full);
        if (ty.last >= 0) acc(ty.last, ty.a_last);
        result = sum;
    }
    small[(size_t)b * hs * ws + t] = result;
}

// area-mode linear coefficient of destination index d when enlarging ssize -> dsize
struct UpCoef {
    int s;        // source index (clamped)
    float a0, a1; // weights of s and s+1
    bool edge;    // d >= xmax: the horizontal pass copies S[s] * 1
};

__device__ __forceinline__ UpCoef up_coef(int d, int ssize, int dsize) {
    const double inv = (double)dsize / (double)ssize;
    const double scale = 1.0 / inv;
    int s = (int)floor(__dmul_rn((double)d, scale));
    float f = (float)__dsub_rn((double)(d + 1), __dmul_rn((double)(s + 1), inv));
    f = f <= 0.f ? 0.f : __fsub_rn(f, floorf(f));
    UpCoef c;
    c.edge = false;
    if (s < 0) { f = 0.f; s = 0; }
    if (s + 1 >= ssize) {
        c.edge = true;
        if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
    }
    c.s = s;
    c.a0 = __fsub_rn(1.0f, f);
    c.a1 = f;
    return c;
}

// pattern = (norm(diff) - enlarge(small) > thr), or (norm(diff) > thr) when small == nullptr.
// grid = (ceil(W/256), H, B)
__global__ void __launch_bounds__(256) sip_pattern_kernel(const double* __restrict__ diff,
                                                          const unsigned long long* __restrict__ minmax,
                                                          const double* __restrict__ small, float* __restrict__ pattern,
                                                          int H, int W, int hs, int ws, double threshold) {
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x >= W) return;
    const int y = blockIdx.y, b = blockIdx.z;
    Norm nz;
    nz.mn = __longlong_as_double((long long)minmax[2 * b]);
    nz.range = __dsub_rn(__longlong_as_double((long long)minmax[2 * b + 1]), nz.mn);
    const size_t o = ((size_t)b * H + y) * W + x;
    const double v = nz(diff[o]);
    double ref = 0.0;
    if (small != nullptr) {
        const double* sm = small + (size_t)b * hs * ws;
        const UpCoef cx = up_coef(x, ws, W), cy = up_coef(y, hs, H);
        const int r0 = min(max(cy.s, 0), hs - 1), r1 = min(max(cy.s + 1, 0), hs - 1);
        auto hpass = [&](int r) {
            const double* row = sm + (size_t)r * ws;
            if (cx.edge) return __dmul_rn(row[cx.s], 1.0);
            return __dadd_rn(__dmul_rn(row[cx.s], (double)cx.a0), __dmul_rn(row[cx.s + 1], (double)cx.a1));
        };
        ref = __dadd_rn(__dmul_rn(hpass(r0), (double)cy.a0), __dmul_rn(hpass(r1), (double)cy.a1));
    }
    pattern[o] = (__dsub_rn(v, ref) > threshold) ? 1.0f : 0.0f;
}

}  // namespace az

using namespace az;

extern "C" int64_t az_sim_ir_pattern_workspace_bytes(int64_t B, int64_t H, int64_t W, int64_t ks) {
    const int64_t small = ks > 0 ? B * (H / ks) * (W / ks) : 0;
    return (B * H * W + small + 2 * B) * (int64_t)sizeof(double);
}

extern "C" int az_sim_ir_pattern(const void* img_ir, const void* img_no_ir, int is_u8, float* pattern, void* workspace,
                                 int64_t B, int64_t H, int64_t W, int64_t ks, double threshold, void* stream) {
    if (!img_ir || !img_no_ir || !pattern || !workspace || B <= 0 || H <= 0 || W <= 0 || ks < 0) return AZ_ERR_BAD_ARG;
    if (B > 65535 || H > 65535 || H * W >= (1ll << 31)) return AZ_ERR_BAD_ARG;
    const int64_t hs = ks > 0 ? H / ks : 0, ws = ks > 0 ? W / ks : 0;
    if (ks > 0 && (hs < 1 || ws < 1)) return AZ_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    double* diff = (double*)workspace;
    double* small = diff + B * H * W;
    unsigned long long* minmax = (unsigned long long*)(small + B * hs * ws);
    const int64_t HW = H * W;
    sip_init_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, st>>>(minmax, (int)B);
    AZ_LAUNCH_CHECK();
    sip_diff_kernel<<<dim3((unsigned)ceil_div(HW, 256), (unsigned)B), 256, 0, st>>>(img_ir, img_no_ir, is_u8, diff, minmax,
                                                                                    HW);
    AZ_LAUNCH_CHECK();
    if (ks > 0) {
        sip_shrink_kernel<<<dim3((unsigned)ceil_div(hs * ws, 128), (unsigned)B), 128, 0, st>>>(diff, minmax, small, (int)H,
                                                                                               (int)W, (int)hs, (int)ws);
        AZ_LAUNCH_CHECK();
    }
    sip_pattern_kernel<<<dim3((unsigned)ceil_div(W, 256), (unsigned)H, (unsigned)B), 256, 0, st>>>(
        diff, minmax, ks > 0 ? small : nullptr, pattern, (int)H, (int)W, (int)hs, (int)ws, threshold);
    AZ_LAUNCH_CHECK();
    return 0;
}
