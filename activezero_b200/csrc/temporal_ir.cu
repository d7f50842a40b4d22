// Temporal IR pattern extraction and local contrast normalisation
// (SURVEY.md §8a rows a11, a12).
// Reference: /root/reference/tools/temporal_ir.py:35-40 (get_smoothed_ir_pattern), :93-114
// (per-pixel regression over the frame stack, normalise, threshold) and
// /root/reference/utils/reprojection.py:175-200 (local_contrast_norm).
//
// a11 (numpy promotes the uint8 frames to float64) is evaluated in EXACT integer arithmetic, see below.  Launches:
//   1. a = |sum_t y_t (2t - (T-1))| per pixel + per-image min/max (integer atomics, deterministic);
//   2. tiled ks x ks box sum of a with BORDER_REFLECT_101 (what cv2.blur uses by default) and the threshold test.
// cv2 / numpy round at the 1e-16 level; pixels whose margin to the threshold is below 1e-6 are outside the parity gate.
#include "common.cuh"

namespace az {

// ---- exact integer form (round 2) ----------------------------------------------------------------
// With t = 0..T-1 the reference's float64 chain (temporal_ir.py:93-114) is, in exact arithmetic,
//   numerator = sum_t (y_t - ybar)(t - tbar) = N2 / 2,   N2 = sum_t y_t (2t - (T-1))          (an INTEGER)
//   denominator = T (T^2 - 1) / 12,   diff = |slope (T-1)| / 255 = |N2| * 6 / (255 T (T+1))
//   diff_n = (diff - min) / (max - min) = (a - a_min) / R,   a = |N2|,  R = a_max - a_min
//   pattern = diff_n - boxmean_ks(diff_n) > thr   <=>   ks^2 a - boxsum_ks(a) > thr ks^2 R     (a_min cancels)
// so the whole operator is integer sums and ONE floating-point product per image: no float64 pipe (round 1's
// kernels spent three IEEE divisions per pixel there), 4-byte instead of 8-byte intermediates, and a result that
// does not depend on float64 rounding at all (the reference's own 1e-16 noise only matters for pixels within
// ~1e-13 of the threshold, far inside the 1e-6 band the parity gate excludes).
// workspace: a[B*H*W] int32, then minmax[2*B] int32
__global__ void __launch_bounds__(256) tir_init_kernel(int* __restrict__ minmax, int B) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t < B) {
        minmax[2 * t] = 0x7fffffff;
        minmax[2 * t + 1] = 0;
    }
}

// grid = (ceil(H*W/V/256), B); a thread owns V adjacent pixels (V = 4: one 32-bit load per frame, one 128-bit store)
// TT > 0: the frame count at compile time (7 in tools/temporal_ir.py:64-70): the T loads are issued back to back and the
// weights 2t - (T-1) are immediates (ncu, run-time T: 62 instructions per pixel, issue 67 %); TT = 0: run-time T.
template <int V, int TT>
__global__ void __launch_bounds__(256) tir_slope_kernel(const uint8_t* __restrict__ frames, int* __restrict__ aout,
                                                        int* __restrict__ minmax, int T_rt, int64_t HW) {
    const int T = TT > 0 ? TT : T_rt;
    const int64_t p = ((int64_t)blockIdx.x * 256 + threadIdx.x) * V;
    const int b = blockIdx.y;
    int mn = 0x7fffffff, mx = 0;
    if (p < HW) {
        const uint8_t* f = frames + (size_t)b * T * HW + p;
        int acc[V];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = 0;
        if (TT > 0 && V == 4) {
            uchar4 q[TT > 0 ? TT : 1];
#pragma unroll
            for (int t = 0; t < TT; ++t) q[t] = __ldg(reinterpret_cast<const uchar4*>(f + (size_t)t * HW));
#pragma unroll
            for (int t = 0; t < TT; ++t) {
                const int wgt = 2 * t - (TT - 1);
                acc[0] += wgt * (int)q[t].x; acc[1 % V] += wgt * (int)q[t].y; acc[2 % V] += wgt * (int)q[t].z; acc[3 % V] += wgt * (int)q[t].w;
            }
        } else {
            for (int t = 0; t < T; ++t) {
                const int wgt = 2 * t - (T - 1);
                if (V == 4) {
                    const uchar4 q = *reinterpret_cast<const uchar4*>(f + (size_t)t * HW);
                    acc[0] += wgt * (int)q.x; acc[1 % V] += wgt * (int)q.y; acc[2 % V] += wgt * (int)q.z; acc[3 % V] += wgt * (int)q.w;
                } else {
                    acc[0] += wgt * (int)f[(size_t)t * HW];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < V; ++j) {
            acc[j] = abs(acc[j]);
            mn = min(mn, acc[j]);
            mx = max(mx, acc[j]);
        }
        int* o = aout + (size_t)b * HW + p;
        if (V == 4) *reinterpret_cast<int4*>(o) = make_int4(acc[0], acc[1 % V], acc[2 % V], acc[3 % V]);
        else o[0] = acc[0];
    }
    mn = __reduce_min_sync(0xffffffffu, mn);  // redux.sync: one instruction each
    mx = __reduce_max_sync(0xffffffffu, mx);
    // one atomic pair per CTA (per-warp atomics on the 2 words of an image serialise in L2: 130 k of them at 544x960)
    __shared__ int smn[8], smx[8];
    if ((threadIdx.x & 31) == 0) {
        smn[threadIdx.x >> 5] = mn;
        smx[threadIdx.x >> 5] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            mn = min(mn, smn[k]);
            mx = max(mx, smx[k]);
        }
        atomicMin(&minmax[2 * b], mn);
        atomicMax(&minmax[2 * b + 1], mx);
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

constexpr int kTirTX = 32, kTirTY = 8;  // tile of the generic LCN fallback below

// Box sum + threshold.  grid = (ceil(W/64), ceil(H/32), B), 256 threads; separable sliding-window INTEGER sums:
//   smem tile (32+ks-1) x (64+ks-1) of a (BORDER_REFLECT_101, what cv2.blur uses by default), horizontal window
//   sums for runs of 8 outputs per thread, vertical window sums for runs of 8 rows per thread, then the test
//   ks^2 a - boxsum > thr ks^2 R  (left side exact in int64, right side one double product per image).
constexpr int kTpTW = 64, kTpTH = 32;

// KS > 0: window size known at compile time (tile extents and the divisions by them become constants); KS = 0: run-time ks.
// Round 2 (ncu at B=8, 544x960: 183 instructions per pixel, 63 % of them in the tile fill, issue 73 %): the fill divides
// by a compile-time tile width, reflects with one step per side (valid whenever the image is larger than the halo; tiny
// images take the loop) and keeps four independent gathers in flight per thread, instead of two run-time divisions and
// two reflection loops per halo element with one load in flight.
__device__ __forceinline__ int reflect101_fast(int i, int n, bool simple) {
    if (simple) {  // n > halo: one reflection on either side is exact for every index a written pixel uses
        i = i < 0 ? -i : i;
        i = i >= n ? 2 * n - 2 - i : i;
        return min(max(i, 0), n - 1);  // tile cells far beyond the image (unused) still read inside it
    }
    return reflect101(i, n);
}

// VEC (W % 4 == 0, 16-byte aligned `ain`): the tile's 64 own columns are filled with 128-bit loads and stores (the tile
// is laid out so that they start at a multiple of four: OFFX leading pad columns), only the 2h halo columns and quads
// that touch the right image border go through the per-element reflection.
template <int KS, bool VEC>
__global__ void __launch_bounds__(256) tir_pattern_kernel(const int* __restrict__ ain, const int* __restrict__ minmax,
                                                          float* __restrict__ pattern, int H, int W, int ks_rt,
                                                          double threshold) {
    extern __shared__ __align__(16) int itile[];
    const int ks = KS > 0 ? KS : ks_rt;
    const int h = ks >> 1;
    const int OFFX = ((h + 3) & ~3) - h;                        // pad columns in front of the tile
    const int IW = kTpTW + ks - 1, IH = kTpTH + ks - 1;
    const int IWP = (OFFX + IW + 3) & ~3;                       // row pitch (multiple of 4)
    int* rows = itile + IWP * IH;  // [IH][TW] horizontal window sums
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * kTpTW, y0 = blockIdx.y * kTpTH;
    // lhs > rhs with an integer lhs  <=>  lhs >= floor(rhs) + 1: one integer compare per pixel instead of an
    // int64 -> double conversion (NaN / +inf: nothing passes, as with the floating-point compare)
    const double rhs = threshold * (double)(ks * ks) * (double)(minmax[2 * b + 1] - minmax[2 * b]);
    long long thr;
    if (!(rhs < 9.0e18)) thr = 0x7fffffffffffffffll;
    else if (rhs < -9.0e18) thr = -0x7fffffffffffffffll - 1;
    else thr = (long long)floor(rhs) + 1;
    const bool never = !(rhs < 9.0e18);
    const int* d = ain + (size_t)b * H * W;
    const int tid = threadIdx.x;
    const bool simple = H > h && W > h;
    if (VEC) {
        // items of a tile row: 16 quads (own columns) then 2h halo columns
        const int per_row = kTpTW / 4 + 2 * h;
        for (int t0 = tid; t0 < per_row * IH; t0 += 2 * 256) {
            int4 v[2];
            int sc[2];
            int dst[2];
            bool isq[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int t = t0 + u * 256;
                const int ty = t / per_row, it = t - ty * per_row;
                const int yy = reflect101_fast(y0 + ty - h, H, simple);
                const int* srow = d + (size_t)yy * W;
                dst[u] = -1;
                isq[u] = false;
                if (t < per_row * IH) {
                    if (it < kTpTW / 4) {
                        const int xx = x0 + 4 * it;
                        dst[u] = ty * IWP + OFFX + h + 4 * it;
                        isq[u] = true;
                        if (xx + 3 < W) {
                            v[u] = __ldg(reinterpret_cast<const int4*>(srow + xx));
                        } else {  // beyond the right border: reflected, element by element
                            v[u].x = __ldg(srow + reflect101_fast(xx, W, simple));
                            v[u].y = __ldg(srow + reflect101_fast(xx + 1, W, simple));
                            v[u].z = __ldg(srow + reflect101_fast(xx + 2, W, simple));
                            v[u].w = __ldg(srow + reflect101_fast(xx + 3, W, simple));
                        }
                    } else {
                        const int k = it - kTpTW / 4;                       // 0 .. 2h-1
                        const int tx = k < h ? k : kTpTW + k;               // left halo, right halo
                        dst[u] = ty * IWP + OFFX + tx;
                        sc[u] = __ldg(srow + reflect101_fast(x0 + tx - h, W, simple));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (dst[u] >= 0) {
                    if (isq[u]) *reinterpret_cast<int4*>(itile + dst[u]) = v[u];
                    else itile[dst[u]] = sc[u];
                }
            }
        }
    } else {
        // four independent gathers in flight per thread (the divisions are by a compile-time IW when KS > 0)
        for (int t0 = tid; t0 < IW * IH; t0 += 4 * 256) {
            int v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int t = t0 + u * 256;
                const int ty = t / IW, tx = t - ty * IW;
                const int yy = reflect101_fast(y0 + ty - h, H, simple), xx = reflect101_fast(x0 + tx - h, W, simple);
                v[u] = t < IW * IH ? __ldg(d + (size_t)yy * W + xx) : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int t = t0 + u * 256;
                if (t < IW * IH) itile[(t / IW) * IWP + OFFX + (t - (t / IW) * IW)] = v[u];
            }
        }
    }
    __syncthreads();
    for (int it = tid; it < IH * (kTpTW / 8); it += 256) {
        const int ty = it / (kTpTW / 8), xl = 8 * (it - ty * (kTpTW / 8));
        const int* r = itile + ty * IWP + OFFX + xl;
        int s = 0;
#pragma unroll
        for (int k = 0; k < (KS > 0 ? KS : 1); ++k) s += r[k];
        if (KS == 0)
            for (int k = 1; k < ks; ++k) s += r[k];
        int* o = rows + ty * kTpTW + xl;
        o[0] = s;
#pragma unroll
        for (int j = 1; j < 8; ++j) {
            s += r[j + ks - 1] - r[j - 1];
            o[j] = s;
        }
    }
    __syncthreads();
    const int c = tid & (kTpTW - 1), r0 = (tid / kTpTW) * 8;
    const int x = x0 + c;
    if (x >= W) return;
    long long s = 0;
#pragma unroll
    for (int k = 0; k < (KS > 0 ? KS : 1); ++k) s += rows[(r0 + k) * kTpTW + c];
    if (KS == 0)
        for (int k = 1; k < ks; ++k) s += rows[(r0 + k) * kTpTW + c];
    const long long k2 = (long long)ks * ks;
    float* prow = pattern + (size_t)b * H * W + (size_t)(y0 + r0) * W + x;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int y = y0 + r0 + r;
        if (y < H) {
            const long long lhs = k2 * (long long)itile[(r0 + r + h) * IWP + OFFX + c + h] - s;
            prow[(size_t)r * W] = (!never && lhs >= thr) ? 1.0f : 0.0f;
        }
        if (r < 7) s += rows[(r0 + r + ks) * kTpTW + c] - rows[(r0 + r) * kTpTW + c];
    }
}

// ------------------------------------------------------------------------------------------
// a12 LCN (utils/reprojection.py:175-200): zero-padded ks x ks window mean and POPULATION std.
// Fast path (ks in {3..13}): separable sliding-window sums of x and x^2 in FLOAT64.  The one-pass form
// E[x^2] - mean^2 cancels catastrophically in float32 for near-constant windows (which is why the first
// version of this kernel made two passes over the ks*ks window: 2 x 81 x 3 float ops per pixel); with the
// products (exact in double: 24 x 24 bits) and both sums carried in double the relative error of the
// variance is ~1e-16 * mean^2 / var, below what the float32 two-pass form (and torch's own float32 std)
// achieves, at ~30 double operations per pixel.  grid = (ceil(W/64), ceil(H/32), B), 256 threads:
//   phase 1: per tile row, horizontal window sums (s1, s2) for 4 adjacent outputs per thread (slide by one);
//   phase 2: per column, vertical window sums over runs of 8 rows (slide by one), then the epilogue.
// ------------------------------------------------------------------------------------------
constexpr int kLsTW = 64, kLsTH = 32;
// Phase 1 maps the lanes of a warp to consecutive tile ROWS (each lane slides along its own row), so the row
// pitches are padded to make the per-lane 16-byte accesses land in distinct bank groups: input tile pitch 84
// floats (336 B: 336*l mod 128 takes 8 distinct values), sums pitch 65 double2 (1040 B).
constexpr int kLsIW = kLsTW + 20, kLsHP = kLsTW + 1;

template <int KS>
__global__ void __launch_bounds__(256) lcn_sep_kernel(const float* __restrict__ image, float* __restrict__ normed,
                                                      float* __restrict__ stdo, int Cin, int H, int W, float eps) {
    extern __shared__ __align__(16) unsigned char lraw[];
    constexpr int h = KS / 2, IH = kLsTH + KS - 1, IWU = kLsTW + KS - 1;
    constexpr int NV = (KS + 3 + 3) / 4;  // float4 vectors covering the 4 + KS - 1 inputs of four adjacent windows
    double2* hs = reinterpret_cast<double2*>(lraw);                    // [IH][HP] horizontal (sum x, sum x^2)
    float* tile = reinterpret_cast<float*>(hs + IH * kLsHP);           // [IH][IW]
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * kLsTW, y0 = blockIdx.y * kLsTH;
    const float* im = image + (size_t)b * Cin * H * W;  // channel 0 (reprojection.py:184-185)
    const int tid = threadIdx.x;
    if ((h & 3) == 0 && (W & 3) == 0 && (reinterpret_cast<uintptr_t>(im) & 15u) == 0) {
        // kernel_size 9 (the reference's default): the tile's first column x0 - 4 is 16-byte aligned, a quad is wholly
        // inside or wholly outside the row -> 128-bit loads, one quarter of the load instructions
        constexpr int Q = (IWU + 3) / 4;
        for (int t = tid; t < IH * Q; t += 256) {
            const int ty = t / Q, q = t - ty * Q;
            const int yy = y0 + ty - h, xx = x0 - h + 4 * q;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(reinterpret_cast<const float4*>(im + (size_t)yy * W + xx));
            *reinterpret_cast<float4*>(tile + ty * kLsIW + 4 * q) = v;
        }
    } else {
        for (int t = tid; t < IH * kLsIW; t += 256) {
            const int ty = t / kLsIW, tx = t - ty * kLsIW;
            const int yy = y0 + ty - h, xx = x0 + tx - h;
            tile[t] = (tx < IWU && yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(im + (size_t)yy * W + xx) : 0.f;
        }
    }
    __syncthreads();
    for (int it = tid; it < IH * (kLsTW / 4); it += 256) {
        const int j4 = it / IH, ty = it - j4 * IH, xl = 4 * j4;
        double d[NV * 4];
        const float4* row = reinterpret_cast<const float4*>(tile + ty * kLsIW + xl);
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const float4 t = row[q];
            d[4 * q] = (double)t.x; d[4 * q + 1] = (double)t.y; d[4 * q + 2] = (double)t.z; d[4 * q + 3] = (double)t.w;
        }
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int k = 0; k < KS; ++k) { s1 += d[k]; s2 = fma(d[k], d[k], s2); }
        double2* o = hs + ty * kLsHP + xl;
        o[0] = make_double2(s1, s2);
#pragma unroll
        for (int j = 1; j < 4; ++j) {
            s1 += d[j + KS - 1] - d[j - 1];
            s2 += d[j + KS - 1] * d[j + KS - 1] - d[j - 1] * d[j - 1];
            o[j] = make_double2(s1, s2);
        }
    }
    __syncthreads();
    const int c = tid & (kLsTW - 1), r0 = (tid / kLsTW) * 8;  // 4 row groups of 8 rows
    const int x = x0 + c;
    if (x >= W) return;
    const double inv = 1.0 / (double)(KS * KS);
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        const double2 v = hs[(r0 + k) * kLsHP + c];
        s1 += v.x;
        s2 += v.y;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int y = y0 + r0 + r;
        if (y < H) {
            const double mean = s1 * inv;
            const double var = fmax(s2 * inv - mean * mean, 0.0);  // population variance (unbiased=False, :193-197)
            const float sd = sqrtf((float)var);
            const float centre = tile[(r0 + r + h) * kLsIW + c + h];
            const size_t o = (size_t)b * H * W + (size_t)y * W + x;
            normed[o] = (centre - (float)mean) / (sd + eps);
            stdo[o] = sd;
        }
        if (r < 7) {
            const double2 add = hs[(r0 + r + KS) * kLsHP + c], sub = hs[(r0 + r) * kLsHP + c];
            s1 += add.x - sub.x;
            s2 += add.y - sub.y;
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
static int launch_lcn_sep(const float* image, float* normed, float* stdo, int B, int Cin, int H, int W, float eps,
                          cudaStream_t st) {
    const size_t smem = (size_t)(kLsTH + KS - 1) * (kLsHP * sizeof(double2) + kLsIW * sizeof(float));
    cudaError_t e = cudaFuncSetAttribute(lcn_sep_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)ceil_div(W, kLsTW), (unsigned)ceil_div(H, kLsTH), (unsigned)B);
    lcn_sep_kernel<KS><<<grid, 256, smem, st>>>(image, normed, stdo, Cin, H, W, eps);
    return (int)cudaGetLastError();
}

}  // namespace az

using namespace az;

extern "C" int64_t az_temporal_ir_workspace_bytes(int64_t B, int64_t H, int64_t W) {
    return (B * H * W + 2 * B) * (int64_t)sizeof(int);
}

extern "C" int az_temporal_ir(const uint8_t* frames, float* pattern, void* workspace, int64_t B, int64_t T, int64_t H,
                              int64_t W, int64_t ks, double threshold, void* stream) {
    if (!frames || !pattern || !workspace || B <= 0 || T < 2 || H <= 0 || W <= 0 || ks < 1 || (ks % 2) == 0)
        return AZ_ERR_BAD_ARG;
    if (B > 65535 || H * W >= (1ll << 31) || ks > 63 || T > 4096) return AZ_ERR_BAD_ARG;  // |N2| <= 255 T^2 / 2 fits int32
    cudaStream_t st = (cudaStream_t)stream;
    int* diff = (int*)workspace;
    int* minmax = diff + B * H * W;
    const int64_t HW = H * W;
    tir_init_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, st>>>(minmax, (int)B);
    AZ_LAUNCH_CHECK();
    if (HW % 4 == 0 && (reinterpret_cast<uintptr_t>(frames) & 3u) == 0 && aligned16(workspace)) {
        dim3 g1((unsigned)ceil_div(HW / 4, 256), (unsigned)B);
        if (T == 7) tir_slope_kernel<4, 7><<<g1, 256, 0, st>>>(frames, diff, minmax, (int)T, HW);
        else if (T == 4) tir_slope_kernel<4, 4><<<g1, 256, 0, st>>>(frames, diff, minmax, (int)T, HW);
        else tir_slope_kernel<4, 0><<<g1, 256, 0, st>>>(frames, diff, minmax, (int)T, HW);
    } else {
        dim3 g1((unsigned)ceil_div(HW, 256), (unsigned)B);
        tir_slope_kernel<1, 0><<<g1, 256, 0, st>>>(frames, diff, minmax, (int)T, HW);
    }
    AZ_LAUNCH_CHECK();
    const int hh = (int)(ks >> 1), OFFX = ((hh + 3) & ~3) - hh;
    const int IW = kTpTW + (int)ks - 1, IH = kTpTH + (int)ks - 1, IWP = (OFFX + IW + 3) & ~3;
    const size_t smem = ((size_t)IWP * IH + (size_t)IH * kTpTW) * sizeof(int);
    dim3 g2((unsigned)ceil_div(W, kTpTW), (unsigned)ceil_div(H, kTpTH), (unsigned)B);
    const bool vec = W % 4 == 0 && aligned16(workspace) && tuning("AZ_TIR_VEC", 1) != 0;
    if (ks == 11 && tuning("AZ_TIR_KS_TEMPLATE", 1) != 0) {  // tools/temporal_ir.py's window
        if (vec) tir_pattern_kernel<11, true><<<g2, 256, smem, st>>>(diff, minmax, pattern, (int)H, (int)W, (int)ks, threshold);
        else tir_pattern_kernel<11, false><<<g2, 256, smem, st>>>(diff, minmax, pattern, (int)H, (int)W, (int)ks, threshold);
    } else {
        auto kern = vec ? tir_pattern_kernel<0, true> : tir_pattern_kernel<0, false>;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        kern<<<g2, 256, smem, st>>>(diff, minmax, pattern, (int)H, (int)W, (int)ks, threshold);
    }
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
        case 3: return launch_lcn_sep<3>(image, normed, stdo, (int)B, (int)Cin, (int)H, (int)W, eps, st);
        case 5: return launch_lcn_sep<5>(image, normed, stdo, (int)B, (int)Cin, (int)H, (int)W, eps, st);
        case 7: return launch_lcn_sep<7>(image, normed, stdo, (int)B, (int)Cin, (int)H, (int)W, eps, st);
        case 9: return launch_lcn_sep<9>(image, normed, stdo, (int)B, (int)Cin, (int)H, (int)W, eps, st);
        case 11: return launch_lcn_sep<11>(image, normed, stdo, (int)B, (int)Cin, (int)H, (int)W, eps, st);
        case 13: return launch_lcn_sep<13>(image, normed, stdo, (int)B, (int)Cin, (int)H, (int)W, eps, st);
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
