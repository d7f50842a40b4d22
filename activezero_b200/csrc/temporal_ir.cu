// Temporal IR pattern extraction and local contrast normalisation
// (SURVEY.md §8a rows a11, a12).
// Reference: /root/reference/tools/temporal_ir.py:35-40 (get_smoothed_ir_pattern), :93-114
// (per-pixel regression over the frame stack, normalise, threshold) and
// /root/reference/utils/reprojection.py:175-200 (local_contrast_norm).
//
// a11 runs in float64 like numpy (uint8 frames are promoted).  Three launches:
//   1. slope/diff per pixel + per-image min/max (positive doubles order like uint64 => integer
//      atomicMin/atomicMax, deterministic);
//   2. tiled ks x ks box blur of the min-max-normalised diff with BORDER_REFLECT_101 (what
//      cv2.blur uses by default), threshold.
// cv2 builds the box sum with running row/column sums whose rounding differs from a direct sum at
// the 1e-16 level; pixels whose margin to the threshold is below 1e-6 are outside the parity gate.
#include "common.cuh"

namespace az {

// workspace layout per image b: diff[H*W] doubles, then (after all images) minmax[2*B] as uint64 bit patterns
__global__ void __launch_bounds__(256) tir_init_kernel(unsigned long long* __restrict__ minmax, int B) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t < B) {
        minmax[2 * t] = 0x7FF0000000000000ull;  // +inf
        minmax[2 * t + 1] = 0ull;               // +0.0
    }
}

template <int V>
__device__ __forceinline__ void tir_load(const uint8_t* __restrict__ f, double* y) {
    if (V == 4) {
        const uchar4 q = *reinterpret_cast<const uchar4*>(f);
        y[0] = (double)q.x; y[1 % V] = (double)q.y; y[2 % V] = (double)q.z; y[3 % V] = (double)q.w;
    } else {
        y[0] = (double)f[0];
    }
}

// grid = (ceil(H*W/V/256), B); a thread owns V adjacent pixels (V = 4: one 32-bit load per frame).
// temporal_ir.py:94-107, numpy float64 op order (no FMA contraction); the frames are read twice (mean,
// then centred products) -- the second pass hits L1/L2.
template <int V>
__global__ void __launch_bounds__(256) tir_slope_kernel(const uint8_t* __restrict__ frames, double* __restrict__ diff,
                                                        unsigned long long* __restrict__ minmax, int T, int64_t HW) {
    const int64_t p = ((int64_t)blockIdx.x * 256 + threadIdx.x) * V;
    const int b = blockIdx.y;
    const bool on = p < HW;
    const double t_avg = (double)((T - 1) * T / 2) / (double)T;  // np.average of the int ramp
    double mn = __longlong_as_double(0x7FF0000000000000ll), mx = 0.0;
    if (on) {
        const uint8_t* f = frames + (size_t)b * T * HW + p;
        double ysum[V], y[V], num[V], y_avg[V];
#pragma unroll
        for (int j = 0; j < V; ++j) { ysum[j] = 0.0; num[j] = 0.0; }
        for (int t = 0; t < T; ++t) {
            tir_load<V>(f + (size_t)t * HW, y);
#pragma unroll
            for (int j = 0; j < V; ++j) ysum[j] += y[j];  // exact (integers)
        }
#pragma unroll
        for (int j = 0; j < V; ++j) y_avg[j] = ysum[j] / (double)T;
        double den = 0.0;
        for (int t = 0; t < T; ++t) {
            const double dt = (double)t - t_avg;
            den = __dadd_rn(den, __dmul_rn(dt, dt));
            tir_load<V>(f + (size_t)t * HW, y);
#pragma unroll
            for (int j = 0; j < V; ++j) num[j] = __dadd_rn(num[j], __dmul_rn(y[j] - y_avg[j], dt));
        }
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const double slope = num[j] / den;
            const double icpt = __dsub_rn(y_avg[j], __dmul_rn(slope, t_avg));
            const double first = __dadd_rn(__dmul_rn(slope, 0.0), icpt);
            const double last = __dadd_rn(__dmul_rn(slope, (double)(T - 1)), icpt);
            const double v = fabs(__dsub_rn(last, first) / 255.0);  // :110-111
            diff[(size_t)b * HW + p + j] = v;
            mn = fmin(mn, v);
            mx = fmax(mx, v);
        }
    }
    // block min / max, then one integer atomic each
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

__device__ __forceinline__ int reflect101(int i, int n) {
    // BORDER_REFLECT_101: gfedcb|abcdefgh|gfedcba ; n == 1 degenerates to 0
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * n - 2 - i;
    }
    return i;
}

constexpr int kTirTX = 32, kTirTY = 8;

// grid = (ceil(W/32), ceil(H/8), B); smem tile (8+ks-1) x (32+ks-1) doubles + row sums
__global__ void __launch_bounds__(kTirTX * kTirTY) tir_pattern_kernel(const double* __restrict__ diff,
                                                                     const unsigned long long* __restrict__ minmax,
                                                                     float* __restrict__ pattern, int H, int W, int ks,
                                                                     double threshold) {
    extern __shared__ double tile[];
    const int h = ks >> 1;
    const int TW = kTirTX + ks - 1, TH = kTirTY + ks - 1;
    double* rows = tile + TW * TH;  // [TH][kTirTX] horizontal window sums
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * kTirTX, y0 = blockIdx.y * kTirTY;
    const double mn = __longlong_as_double((long long)minmax[2 * b]);
    const double mx = __longlong_as_double((long long)minmax[2 * b + 1]);
    const double range = __dsub_rn(mx, mn);
    const double* d = diff + (size_t)b * H * W;
    const int tid = threadIdx.y * kTirTX + threadIdx.x;
    for (int t = tid; t < TW * TH; t += kTirTX * kTirTY) {
        const int ty = t / TW, tx = t - ty * TW;
        const int yy = reflect101(y0 + ty - h, H), xx = reflect101(x0 + tx - h, W);
        tile[t] = fabs(__dsub_rn(d[(size_t)yy * W + xx], mn) / range);  // :113 normalise, :36 abs
    }
    __syncthreads();
    for (int t = tid; t < TH * kTirTX; t += kTirTX * kTirTY) {
        const int ty = t / kTirTX, tx = t - ty * kTirTX;
        double s = 0.0;
        for (int k = 0; k < ks; ++k) s += tile[ty * TW + tx + k];
        rows[t] = s;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x < W && y < H) {
        double s = 0.0;
        for (int k = 0; k < ks; ++k) s += rows[(threadIdx.y + k) * kTirTX + threadIdx.x];
        const double blur = __dmul_rn(s, 1.0 / (double)(ks * ks));
        const double v = tile[(threadIdx.y + h) * TW + threadIdx.x + h];
        pattern[(size_t)b * H * W + (size_t)y * W + x] = (__dsub_rn(v, blur) > threshold) ? 1.0f : 0.0f;
    }
}

// ------------------------------------------------------------------------------------------
// a12 LCN (utils/reprojection.py:175-200): zero-padded ks x ks window mean and POPULATION std, two-pass
// (mean first, then squared deviations -- the one-pass E[x^2]-mean^2 form cancels catastrophically in
// near-constant windows).
// Fast path (ks in {3..13}): a thread owns FOUR adjacent outputs; per window row it reads its 4+ks-1 values
// as aligned 128-bit shared loads and reuses them for the four windows (12x fewer shared loads than one
// output per thread).  grid = (ceil(W/128), ceil(H/8), B), block (32, 8); tile (8+ks-1) x (128+16) floats.
// ------------------------------------------------------------------------------------------
constexpr int kLcnTW = 128, kLcnTH = 8, kLcnPad = 16;  // pad >= ks-1+3, multiple of 4

template <int KS>
__global__ void __launch_bounds__(256) lcn_strip_kernel(const float* __restrict__ image, float* __restrict__ normed,
                                                        float* __restrict__ stdo, int Cin, int H, int W, float eps) {
    extern __shared__ __align__(16) float ftile[];
    constexpr int h = KS / 2;
    constexpr int TW = kLcnTW + kLcnPad, TH = kLcnTH + KS - 1;
    constexpr int NV = (KS + 3 + 3) / 4;  // float4 vectors covering the 4 + KS - 1 values of a window row
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * kLcnTW, y0 = blockIdx.y * kLcnTH;
    const float* im = image + (size_t)b * Cin * H * W;  // channel 0 (reprojection.py:184-185)
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int t = tid; t < TW * TH; t += 256) {
        const int ty = t / TW, tx = t - ty * TW;
        const int yy = y0 + ty - h, xx = x0 + tx - h;
        ftile[t] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(im + (size_t)yy * W + xx) : 0.f;
    }
    __syncthreads();
    const int xl = threadIdx.x * 4;  // first of this thread's four output columns inside the tile
    const int y = y0 + threadIdx.y;
    if (y >= H || x0 + xl >= W) return;
    const float inv = 1.0f / (float)(KS * KS);
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ky = 0; ky < KS; ++ky) {
        float v[NV * 4];
        const float4* row = reinterpret_cast<const float4*>(ftile + (threadIdx.y + ky) * TW + xl);
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const float4 t = row[q];
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < KS; ++k) s[j] += v[j + k];
    }
    float mean[4], q2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) mean[j] = s[j] * inv;
    float centre[4];
#pragma unroll
    for (int ky = 0; ky < KS; ++ky) {
        float v[NV * 4];
        const float4* row = reinterpret_cast<const float4*>(ftile + (threadIdx.y + ky) * TW + xl);
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const float4 t = row[q];
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
        if (ky == h) {
#pragma unroll
            for (int j = 0; j < 4; ++j) centre[j] = v[j + h];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                const float d = v[j + k] - mean[j];
                q2[j] = fmaf(d, d, q2[j]);
            }
    }
    const size_t o = (size_t)b * H * W + (size_t)y * W + x0 + xl;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (x0 + xl + j < W) {
            const float sd = sqrtf(q2[j] * inv);  // population std (unbiased=False, :193-197)
            normed[o + j] = (centre[j] - mean[j]) / (sd + eps);
            stdo[o + j] = sd;
        }
    }
}

// generic fallback: one output per thread.  grid = (ceil(W/32), ceil(H/8), B); tile (8+ks-1) x (32+ks-1)
__global__ void __launch_bounds__(kTirTX * kTirTY) lcn_kernel(const float* __restrict__ image,
                                                             float* __restrict__ normed, float* __restrict__ stdo,
                                                             int Cin, int H, int W, int ks, float eps) {
    extern __shared__ float ftile[];
    const int h = ks >> 1;
    const int TW = kTirTX + ks - 1, TH = kTirTY + ks - 1;
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * kTirTX, y0 = blockIdx.y * kTirTY;
    const float* im = image + (size_t)b * Cin * H * W;
    const int tid = threadIdx.y * kTirTX + threadIdx.x;
    for (int t = tid; t < TW * TH; t += kTirTX * kTirTY) {
        const int ty = t / TW, tx = t - ty * TW;
        const int yy = y0 + ty - h, xx = x0 + tx - h;
        ftile[t] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(im + (size_t)yy * W + xx) : 0.f;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x < W && y < H) {
        const float inv = 1.0f / (float)(ks * ks);
        float s = 0.f;
        for (int ky = 0; ky < ks; ++ky)
            for (int kx = 0; kx < ks; ++kx) s += ftile[(threadIdx.y + ky) * TW + threadIdx.x + kx];
        const float mean = s * inv;
        float q = 0.f;
        for (int ky = 0; ky < ks; ++ky)
            for (int kx = 0; kx < ks; ++kx) {
                const float t = ftile[(threadIdx.y + ky) * TW + threadIdx.x + kx] - mean;
                q = fmaf(t, t, q);
            }
        const float sd = sqrtf(q * inv);
        const float v = ftile[(threadIdx.y + h) * TW + threadIdx.x + h];
        const size_t o = (size_t)b * H * W + (size_t)y * W + x;
        normed[o] = (v - mean) / (sd + eps);
        stdo[o] = sd;
    }
}

template <int KS>
static int launch_lcn_strip(const float* image, float* normed, float* stdo, int B, int Cin, int H, int W, float eps,
                            cudaStream_t st) {
    const size_t smem = (size_t)(kLcnTW + kLcnPad) * (kLcnTH + KS - 1) * sizeof(float);
    dim3 grid((unsigned)ceil_div(W, kLcnTW), (unsigned)ceil_div(H, kLcnTH), (unsigned)B);
    lcn_strip_kernel<KS><<<grid, dim3(32, 8), smem, st>>>(image, normed, stdo, Cin, H, W, eps);
    return (int)cudaGetLastError();
}

}  // namespace az

using namespace az;

extern "C" int64_t az_temporal_ir_workspace_bytes(int64_t B, int64_t H, int64_t W) {
    return (B * H * W + 2 * B) * (int64_t)sizeof(double);
}

extern "C" int az_temporal_ir(const uint8_t* frames, float* pattern, void* workspace, int64_t B, int64_t T, int64_t H,
                              int64_t W, int64_t ks, double threshold, void* stream) {
    if (!frames || !pattern || !workspace || B <= 0 || T < 2 || H <= 0 || W <= 0 || ks < 1 || (ks % 2) == 0)
        return AZ_ERR_BAD_ARG;
    if (B > 65535 || H * W >= (1ll << 31) || ks > 63) return AZ_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    double* diff = (double*)workspace;
    unsigned long long* minmax = (unsigned long long*)(diff + B * H * W);
    const int64_t HW = H * W;
    tir_init_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, st>>>(minmax, (int)B);
    AZ_LAUNCH_CHECK();
    if (HW % 4 == 0 && (reinterpret_cast<uintptr_t>(frames) & 3u) == 0) {
        dim3 g1((unsigned)ceil_div(HW / 4, 256), (unsigned)B);
        tir_slope_kernel<4><<<g1, 256, 0, st>>>(frames, diff, minmax, (int)T, HW);
    } else {
        dim3 g1((unsigned)ceil_div(HW, 256), (unsigned)B);
        tir_slope_kernel<1><<<g1, 256, 0, st>>>(frames, diff, minmax, (int)T, HW);
    }
    AZ_LAUNCH_CHECK();
    const int TW = kTirTX + (int)ks - 1, TH = kTirTY + (int)ks - 1;
    const size_t smem = ((size_t)TW * TH + (size_t)TH * kTirTX) * sizeof(double);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(tir_pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    dim3 g2((unsigned)ceil_div(W, kTirTX), (unsigned)ceil_div(H, kTirTY), (unsigned)B);
    tir_pattern_kernel<<<g2, dim3(kTirTX, kTirTY), smem, st>>>(diff, minmax, pattern, (int)H, (int)W, (int)ks, threshold);
    AZ_LAUNCH_CHECK();
    return 0;
}

extern "C" int az_local_contrast_norm(const float* image, float* normed, float* stdo, int64_t B, int64_t Cin, int64_t H,
                                      int64_t W, int64_t ks, float eps, void* stream) {
    if (!image || !normed || !stdo || B <= 0 || Cin <= 0 || H <= 0 || W <= 0 || ks < 1 || (ks % 2) == 0 || ks > 63)
        return AZ_ERR_BAD_ARG;
    if (B > 65535 || H * W >= (1ll << 31)) return AZ_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    switch (ks) {
        case 3: return launch_lcn_strip<3>(image, normed, stdo, (int)B, (int)Cin, (int)H, (int)W, eps, st);
        case 5: return launch_lcn_strip<5>(image, normed, stdo, (int)B, (int)Cin, (int)H, (int)W, eps, st);
        case 7: return launch_lcn_strip<7>(image, normed, stdo, (int)B, (int)Cin, (int)H, (int)W, eps, st);
        case 9: return launch_lcn_strip<9>(image, normed, stdo, (int)B, (int)Cin, (int)H, (int)W, eps, st);
        case 11: return launch_lcn_strip<11>(image, normed, stdo, (int)B, (int)Cin, (int)H, (int)W, eps, st);
        case 13: return launch_lcn_strip<13>(image, normed, stdo, (int)B, (int)Cin, (int)H, (int)W, eps, st);
        default: break;
    }
    const int TW = kTirTX + (int)ks - 1, TH = kTirTY + (int)ks - 1;
    const size_t smem = (size_t)TW * TH * sizeof(float);
    dim3 grid((unsigned)ceil_div(W, kTirTX), (unsigned)ceil_div(H, kTirTY), (unsigned)B);
    lcn_kernel<<<grid, dim3(kTirTX, kTirTY), smem, st>>>(image, normed, stdo, (int)Cin, (int)H, (int)W,
                                                                          (int)ks, eps);
    AZ_LAUNCH_CHECK();
    return 0;
}
