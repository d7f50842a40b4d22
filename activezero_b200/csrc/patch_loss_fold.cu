// Patch reprojection loss AND its Fold image in one pass -- what get_reproj_error_patch returns
// (/root/reference/utils/reprojection.py:99-127).
//
//   loss = mean_{mask} (Wu - Lu)^2,   Wu = apply_disparity(Unfold(R), -disp),  Lu = Unfold(L)   (:102-118)
//   vis  = Fold(Wu) cropped                                                                     (:120-125)
//
// The reference materialises three [B, ps*ps, H, W] tensors.  Here every tap Wu[(ky,kx)][i,j] is
// formed once in registers and feeds (a) the squared residual (and d/dxs) of its own pixel and
// (b) a systolic shuffle chain that sums the taps landing on one Fold pixel.
//
// Work decomposition.  One CTA owns a band of consecutive SOURCE rows i of one image.  For each row
// it stages, in shared memory,
//   Rs  the vertically interpolated source rows (ys depends on i only): one row per tap row ky,
//   Ls  the target rows i-p..i+p (a ring: one new row per source row),
//   Ps  per source pixel j: the window start x0(j) and the horizontal weight (sample_pos once per pixel),
// double-buffered, so staging of row i+1 overlaps the taps of row i (one barrier per row).
// A warp owns a strip of 128 consecutive sources (4 per lane) of which S = 128-(ps-1) are "owned"
// (the rest is the halo the chain needs).  Tap rows are processed two at a time as PACKED fp32x2
// values (FFMA2/FADD2, one issue slot for two taps): every shared-memory array interleaves rows
// (y, y+1) for even y, so one 64-bit load yields the operand pair.  ps tap rows = (ps+1)/2 row
// pairs, one half of one pair is a dummy whose results are discarded.  The row pairs of a strip are
// split over G warp groups.  The Fold partial sums live in a shared ring (Vacc) of the ps output
// rows still open; a row is written to HBM the moment its last contribution (ky = 0) arrives.
// Rows shared with the neighbouring band get one atomic add per element onto zero-filled memory --
// two addends per element, hence still deterministic.
//
// Shared-memory bandwidth: lane l owns sources 4l..4l+3, so for a smooth disparity map a plain row
// layout would make the Rs loads 4-way bank conflicted (stride 4 words).  Rs is therefore stored
// de-interleaved by column mod 4 ("planes"): element xx lives at plane xx&3, slot xx>>2, and the
// window walk xx = base+k uses four base pointers per source with compile-time offsets.  Ls and
// Vacc use the same plane layout, so a lane's 4 adjacent columns are four 64-bit accesses that are
// each stride-1 across the warp (a 128-bit access at a 32-byte lane stride is 2-way conflicted).
#include "common.cuh"

namespace az {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ u64 lds64(uint32_t addr) {
    u64 v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64u(uint32_t addr, u64 v) {
    asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, float lo, float hi) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(lo), "f"(hi) : "memory");
}

struct PlfArgs {
    const float* tgt;
    const float* src;
    const float* disp;
    const uint8_t* mask;
    const float* lin_x;
    const float* lin_y;
    float* vis;
    float* gpre;
    double* partial;
    float sign;
    int C, H, W;
    int band_rows;  // source rows per CTA
    int npass;      // strips per row = ceil(W / S)
    int G;          // warp groups sharing the row pairs of a strip
    int vec;        // W % 4 == 0 and every image pointer 16-byte aligned
    int pitchR, pitchL, pitchV, pitchP;
};

template <int PS>
struct PlfGeom {
    static constexpr int P = (PS - 1) / 2;
    static constexpr int NP = (PS + 1) / 2;              // row pairs per source row
    static constexpr int S = (128 - (PS - 1)) & ~3;      // owned sources (= Fold outputs) per strip
    static constexpr int OFF_R = 16;                     // column of x = 0 in a staged source row (>= 12 zeros before it)
    static constexpr int OFF_L = 2 * P;                  // column of x = 0 in a staged target row
    static constexpr int LRING = NP + 1;                 // target row pairs resident (one being staged)
    static constexpr int VRING = NP;                     // open Fold row pairs
    static constexpr int NLW = PS + 3;                   // target columns a lane needs (4 sources, PS taps)
    static constexpr int YB = 16;                        // bias that keeps (row + YB) non-negative
};

__device__ __forceinline__ float4 plf_load4(const float* __restrict__ img, int y, int x, int H, int W, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y >= 0 && y < H) {
        const float* p = img + (size_t)y * W + x;
        if (vec) {
            v = __ldg(reinterpret_cast<const float4*>(p));
        } else {
            if (x < W) v.x = __ldg(p);
            if (x + 1 < W) v.y = __ldg(p + 1);
            if (x + 2 < W) v.z = __ldg(p + 2);
            if (x + 3 < W) v.w = __ldg(p + 3);
        }
    }
    return v;
}

// ---- staging of one source row i into buffer `buf` -------------------------------------------
template <int PS>
__device__ __forceinline__ void plf_stage_row(const PlfArgs& a, const float* __restrict__ sp,
                                              const float* __restrict__ tp, const float* __restrict__ drow_img,
                                              const uint8_t* __restrict__ mimg, int i, int buf, bool first,
                                              uint32_t sRs, uint32_t sLs, float* PsW, int* PsC) {
    using T = PlfGeom<PS>;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int H = a.H, W = a.W;
    const int nq = (W + 3) >> 2;
    const bool vec = a.vec != 0;
    const int e0 = (i - T::P + T::YB) & 1;
    const int pid0 = (i - T::P + T::YB) >> 1;

    // (1) blended source rows, pair-interleaved and de-interleaved by column mod 4
    if (tid < nq) {
        const Axis ay = make_axis(sample_pos(__ldg(a.lin_y + i), 0.0f, (float)H), H);
        const float ay0 = ay.v0 ? ay.e : 0.f, ay1 = ay.v1 ? ay.w : 0.f;
        const int ytop = ay.i0 - T::P - e0;  // image row feeding slot 0 through corner y0
        const int x = 4 * tid;
        const uint32_t q2 = (uint32_t)a.pitchR * 2u;  // bytes between planes: (pitchR/4) slots * 8 B
        uint32_t dst = sRs + (uint32_t)buf * (uint32_t)(T::NP * 2 * a.pitchR * 4) + (uint32_t)((T::OFF_R + x) >> 2) * 8u;
        float4 prev = plf_load4(sp, ytop, x, H, W, vec);
#pragma unroll
        for (int r = 0; r < T::NP; ++r) {
            const float4 mid = plf_load4(sp, ytop + 2 * r + 1, x, H, W, vec);
            const float4 nxt = plf_load4(sp, ytop + 2 * r + 2, x, H, W, vec);
            sts64(dst, fmaf(ay1, mid.x, ay0 * prev.x), fmaf(ay1, nxt.x, ay0 * mid.x));
            sts64(dst + q2, fmaf(ay1, mid.y, ay0 * prev.y), fmaf(ay1, nxt.y, ay0 * mid.y));
            sts64(dst + 2 * q2, fmaf(ay1, mid.z, ay0 * prev.z), fmaf(ay1, nxt.z, ay0 * mid.z));
            sts64(dst + 3 * q2, fmaf(ay1, mid.w, ay0 * prev.w), fmaf(ay1, nxt.w, ay0 * mid.w));
            dst += (uint32_t)a.pitchR * 8u;
            prev = nxt;
        }
    }

    // (2) per-source sampling parameters, and (3) the target row pair(s) entering the ring
    const int tid2 = nthreads - 1 - tid;
    if (tid2 < nq) {
        const int x = 4 * tid2;
        const float4 d4 = plf_load4(drow_img, i, x, H, W, vec);
        const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
        uint32_t m4 = 0x01010101u;
        if (mimg != nullptr) {
            const uint8_t* mp = mimg + (size_t)i * W + x;
            if (vec) {
                m4 = __ldg(reinterpret_cast<const uint32_t*>(mp));
            } else {
                m4 = 0;
                for (int t = 0; t < 4; ++t)
                    if (x + t < W) m4 |= (uint32_t)(__ldg(mp + t) != 0) << (8 * t);
            }
        }
        float* pw = PsW + buf * a.pitchP + T::P + x;
        int* pc = PsC + buf * a.pitchP + T::P + x;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (x + t < W) {
                const Axis ax = make_axis(
                    sample_pos(__ldg(a.lin_x + x + t), __fdiv_rn(a.sign * dd[t], (float)W), (float)W), W);
                int base = 0, flags = 0;
                float wx = 0.f;
                if (ax.v0 || ax.v1) {
                    base = T::OFF_R + min(max(ax.i0, -1), W - 1) - T::P;
                    wx = ax.w;
                    flags = (ax.v0 ? 0 : 1) | (ax.v1 ? 0 : 2);
                }
                pw[t] = wx;
                pc[t] = (base << 3) | (((m4 >> (8 * t)) & 0xffu) ? 4 : 0) | flags;
            }
        }
    }
    const int np_new = first ? T::NP : (((i - T::P + T::YB) & 1) == 0 ? 1 : 0);  // a new pair enters when e0 flips to 0
    for (int n = 0; n < np_new; ++n) {
        const int pid = first ? pid0 + n : pid0 + T::NP - 1;
        const int y = 2 * pid - T::YB;
        const uint32_t rowb = sLs + (uint32_t)((pid % T::LRING) * a.pitchL) * 8u;
        const uint32_t ql8 = (uint32_t)a.pitchL * 2u;  // plane stride in bytes
        for (int m = tid2; m < nq; m += nthreads) {
            const int x = 4 * m;
            const float4 u = plf_load4(tp, y, x, H, W, vec);
            const float4 v = plf_load4(tp, y + 1, x, H, W, vec);
            const uint32_t sb = rowb + (uint32_t)((T::OFF_L + x) >> 2) * 8u;  // slot of column OFF_L + x
            // column OFF_L + x + vv lives in plane (OFF_L + vv) & 3, slot (OFF_L + x + vv) >> 2
            sts64(sb + (uint32_t)((T::OFF_L + 0) & 3) * ql8 + (uint32_t)(((T::OFF_L & 3) + 0) >> 2) * 8u, u.x, v.x);
            sts64(sb + (uint32_t)((T::OFF_L + 1) & 3) * ql8 + (uint32_t)(((T::OFF_L & 3) + 1) >> 2) * 8u, u.y, v.y);
            sts64(sb + (uint32_t)((T::OFF_L + 2) & 3) * ql8 + (uint32_t)(((T::OFF_L & 3) + 2) >> 2) * 8u, u.z, v.z);
            sts64(sb + (uint32_t)((T::OFF_L + 3) & 3) * ql8 + (uint32_t)(((T::OFF_L & 3) + 3) >> 2) * 8u, u.w, v.w);
        }
    }
}

// ---- taps of the row pairs [r_lo, r_hi) of one strip -----------------------------------------
template <int PS, bool GRAD, bool EDGE>
__device__ __forceinline__ void plf_taps(const PlfArgs& a, int i, int i0, int rows_here, int r_lo, int r_hi,
                                         const int (&code)[4], const float (&wxs)[4], uint32_t sRsBuf, uint32_t sLs,
                                         uint32_t sV, int lstart, int vcol, bool vlane, float* __restrict__ vimg,
                                         float (&sq_out)[4], float (&g_out)[4]) {
    using T = PlfGeom<PS>;
    const int e0 = (i - T::P + T::YB) & 1;
    const int pid0 = (i - T::P + T::YB) >> 1;
    const bool asc = e0 == 1;  // the pair holding the dummy half is processed first in its group
    const uint32_t pairBytes = (uint32_t)a.pitchR * 8u;
    const uint32_t Q8 = (uint32_t)a.pitchR * 2u;  // plane stride in bytes

    uint32_t addr[4][4];
    u64 wx2[4], m0[4], m1[4], sq2[4], g2[4];
    const int r_first = asc ? r_lo : r_hi - 1;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int base = code[t] >> 3;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int xx = base + j;
            addr[t][j] = sRsBuf + (uint32_t)r_first * pairBytes + (uint32_t)(xx & 3) * Q8 + (uint32_t)(xx >> 2) * 8u;
        }
        wx2[t] = pk2(wxs[t], wxs[t]);
        if (EDGE) {
            const float f0 = (code[t] & 1) ? 0.f : 1.f, f1 = (code[t] & 2) ? 0.f : 1.f;
            m0[t] = pk2(f0, f0);
            m1[t] = pk2(f1, f1);
        }
        sq2[t] = 0ull;
        g2[t] = 0ull;
    }
    const uint32_t step = asc ? pairBytes : (0u - pairBytes);

#pragma unroll 1
    for (int n = 0; n < r_hi - r_lo; ++n) {
        const int r = asc ? r_lo + n : r_hi - 1 - n;
        const int pid = pid0 + r;
        // target window: NLW columns x 2 rows
        u64 Lw[T::NLW];
        const uint32_t la = sLs + (uint32_t)((pid % T::LRING) * a.pitchL) * 8u + (uint32_t)(lstart >> 2) * 8u;
#pragma unroll
        for (int m = 0; m < T::NLW; ++m) Lw[m] = lds64(la + (uint32_t)(m & 3) * (uint32_t)(a.pitchL * 2) + 8u * (m >> 2));
        u64 a2[4], acc2[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            a2[t] = lds64(addr[t][0]);
            acc2[t] = 0ull;
        }
#pragma unroll
        for (int k = 0; k < PS; ++k) {
            u64 wv[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const u64 n2 = lds64(addr[t][(k + 1) & 3] + 8u * ((k + 1) >> 2));
                u64 am = a2[t], nm = n2;
                if (EDGE) {
                    am = mul2(am, m0[t]);
                    nm = mul2(nm, m1[t]);
                }
                const u64 dk = sub2(nm, am);
                const u64 w = fma2(wx2[t], dk, am);
                const u64 e = sub2(w, Lw[t + k]);
                sq2[t] = fma2(e, e, sq2[t]);
                if (GRAD) g2[t] = fma2(e, dk, g2[t]);
                wv[t] = w;
                a2[t] = n2;
            }
            const u64 fn = __shfl_down_sync(0xffffffffu, acc2[0], 1);
            acc2[0] = add2(wv[0], acc2[1]);
            acc2[1] = add2(wv[1], acc2[2]);
            acc2[2] = add2(wv[2], acc2[3]);
            acc2[3] = add2(wv[3], fn);
        }
        // the half outside the ps tap rows (ky = -1 or ky = ps) contributes nothing
        const bool dlo = (e0 == 1) && (r == 0), dhi = (e0 == 0) && (r == T::NP - 1);
        if (dlo || dhi) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float lo, hi;
                upk2(acc2[t], lo, hi);
                acc2[t] = dlo ? pk2(0.f, hi) : pk2(lo, 0.f);
                upk2(sq2[t], lo, hi);
                sq2[t] = dlo ? pk2(0.f, hi) : pk2(lo, 0.f);
                if (GRAD) {
                    upk2(g2[t], lo, hi);
                    g2[t] = dlo ? pk2(0.f, hi) : pk2(lo, 0.f);
                }
            }
        }
        // Fold ring: rows (2*pid - YB, +1), columns vcol..vcol+3
        if (vlane) {
            const uint32_t va = sV + (uint32_t)((pid % T::VRING) * a.pitchV) * 8u + (uint32_t)(vcol >> 2) * 8u;
            const uint32_t qv8 = (uint32_t)a.pitchV * 2u;  // plane stride in bytes
            u64 v0 = lds64(va), v1 = lds64(va + qv8), v2 = lds64(va + 2 * qv8), v3 = lds64(va + 3 * qv8);
            v0 = add2(v0, acc2[0]);
            v1 = add2(v1, acc2[1]);
            v2 = add2(v2, acc2[2]);
            v3 = add2(v3, acc2[3]);
            if (r == 0) {
                // ky = 0 is the last contribution to output row y = i - p: emit it and recycle its half
                float l0, h0, l1, h1, l2, h2, l3, h3;
                upk2(v0, l0, h0);
                upk2(v1, l1, h1);
                upk2(v2, l2, h2);
                upk2(v3, l3, h3);
                const int y = i - T::P;
                const float o0 = e0 ? h0 : l0, o1 = e0 ? h1 : l1, o2 = e0 ? h2 : l2, o3 = e0 ? h3 : l3;
                if (y >= 0 && vcol < a.W) {
                    float* op = vimg + (size_t)y * a.W + vcol;
                    const bool shared_row = (i0 > 0 && y < i0 + T::P) ||
                                            (i0 + rows_here < a.H && y > i0 + rows_here - 1 - T::P);
                    if (shared_row) {
                        atomicAdd(op, o0);
                        if (vcol + 1 < a.W) atomicAdd(op + 1, o1);
                        if (vcol + 2 < a.W) atomicAdd(op + 2, o2);
                        if (vcol + 3 < a.W) atomicAdd(op + 3, o3);
                    } else if (a.vec) {
                        *reinterpret_cast<float4*>(op) = make_float4(o0, o1, o2, o3);
                    } else {
                        op[0] = o0;
                        if (vcol + 1 < a.W) op[1] = o1;
                        if (vcol + 2 < a.W) op[2] = o2;
                        if (vcol + 3 < a.W) op[3] = o3;
                    }
                }
                if (e0) { v0 = pk2(l0, 0.f); v1 = pk2(l1, 0.f); v2 = pk2(l2, 0.f); v3 = pk2(l3, 0.f); }
                else    { v0 = pk2(0.f, h0); v1 = pk2(0.f, h1); v2 = pk2(0.f, h2); v3 = pk2(0.f, h3); }
            }
            sts64u(va, v0);
            sts64u(va + qv8, v1);
            sts64u(va + 2 * qv8, v2);
            sts64u(va + 3 * qv8, v3);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int j = 0; j < 4; ++j) addr[t][j] += step;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        float lo, hi;
        upk2(sq2[t], lo, hi);
        sq_out[t] = lo + hi;
        if (GRAD) {
            upk2(g2[t], lo, hi);
            g_out[t] = lo + hi;
        } else {
            g_out[t] = 0.f;
        }
    }
}

constexpr int kPlfMaxThreads = 576;

template <int PS, bool GRAD>
__global__ void __launch_bounds__(kPlfMaxThreads, 1) patch_loss_fold_v2_kernel(const PlfArgs a) {
    using T = PlfGeom<PS>;
    extern __shared__ __align__(16) float sm[];
    __shared__ double red[32];
    const int tid = threadIdx.x, nthreads = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int q = warp % a.npass, grp = warp / a.npass;
    const int band = blockIdx.x, b = blockIdx.y;
    const int H = a.H, W = a.W;
    const int i0 = band * a.band_rows;
    const int rows_here = min(a.band_rows, H - i0);
    const size_t HW = (size_t)H * W;

    // shared-memory carve-up (floats)
    const int rsBuf = T::NP * 2 * a.pitchR;
    float* Rs = sm;                                    // [2][NP][4 planes][pitchR/4][2]
    float* Ls = Rs + 2 * rsBuf;                        // [LRING][4 planes][pitchL/4][2]
    float* Vacc = Ls + T::LRING * 2 * a.pitchL;        // [VRING][4 planes][pitchV/4][2]
    float* PsW = Vacc + T::VRING * 2 * a.pitchV;       // [2][pitchP] horizontal weight of each source
    int* PsC = reinterpret_cast<int*>(PsW + 2 * a.pitchP);  // [2][pitchP] window start << 3 | mask << 2 | edge flags
    float* Gs = reinterpret_cast<float*>(PsC + 2 * a.pitchP);  // [2][G-1][npass*128] gradient partials of groups 1.. (GRAD)
    const int pitchG = a.npass * 128, gidx = q * 128 + 4 * lane;  // per-strip slots: strips overlap in Ps indices
    const int total_floats = 2 * rsBuf + T::LRING * 2 * a.pitchL + T::VRING * 2 * a.pitchV + 4 * a.pitchP +
                             (GRAD ? 2 * (a.G - 1) * pitchG : 0);
    const uint32_t sRs = (uint32_t)__cvta_generic_to_shared(Rs);
    const uint32_t sLs = (uint32_t)__cvta_generic_to_shared(Ls);
    const uint32_t sV = (uint32_t)__cvta_generic_to_shared(Vacc);

    // lane-static geometry
    const int li0 = 4 * lane;
    const int s0 = q * T::S - T::P + li0;  // image column of this lane's first source
    const int pidx = q * T::S + li0;       // its index in Ps (= s0 + P)
    bool own[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) own[t] = (li0 + t >= T::P) && (li0 + t < T::P + T::S) && (s0 + t < W);
    const int lstart = min(q * T::S + li0, (a.pitchL - T::NLW) & ~3);
    const int vcol = q * T::S + li0;
    const bool vlane = (li0 < T::S) && (vcol < a.pitchV);
    const int r_lo = grp * T::NP / a.G, r_hi = (grp + 1) * T::NP / a.G;

    const float* dimg = a.disp + (size_t)b * HW;
    const uint8_t* mimg = a.mask == nullptr ? nullptr : a.mask + (size_t)b * HW;

    double tot = 0.0, cnt = 0.0;
    for (int c = 0; c < a.C; ++c) {
        const float* sp = a.src + ((size_t)b * a.C + c) * HW;
        const float* tp = a.tgt + ((size_t)b * a.C + c) * HW;
        float* vimg = a.vis + ((size_t)b * a.C + c) * HW;
        __syncthreads();
        for (int t = tid; t < total_floats / 4; t += nthreads)
            reinterpret_cast<float4*>(sm)[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        plf_stage_row<PS>(a, sp, tp, dimg, mimg, i0, 0, true, sRs, sLs, PsW, PsC);
        __syncthreads();

        float gkeep[4] = {0.f, 0.f, 0.f, 0.f};
        for (int ii = 0; ii < rows_here; ++ii) {
            const int i = i0 + ii, buf = ii & 1;
            if (ii + 1 < rows_here)
                plf_stage_row<PS>(a, sp, tp, dimg, mimg, i + 1, buf ^ 1, false, sRs, sLs, PsW, PsC);
            if (GRAD && grp == 0 && ii > 0) {
                // gradient of row i-1: own partial + the other groups' partials (written before the barrier)
                const float* gs = Gs + ((ii - 1) & 1) * (a.G - 1) * pitchG + gidx;
                float* gp = a.gpre + (size_t)b * HW + (size_t)(i - 1) * W + s0;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    if (own[t]) {
                        float g = gkeep[t];
                        for (int gg = 0; gg < a.G - 1; ++gg) g += gs[gg * pitchG + t];
                        gp[t] = (c == 0 ? 0.f : gp[t]) + g;
                    }
                }
            }
            // this row's parameters
            const int4 c4 = *reinterpret_cast<const int4*>(PsC + buf * a.pitchP + pidx);
            const float4 w4 = *reinterpret_cast<const float4*>(PsW + buf * a.pitchP + pidx);
            const int code[4] = {c4.x, c4.y, c4.z, c4.w};
            const float wxs[4] = {w4.x, w4.y, w4.z, w4.w};
            const bool edge = __any_sync(0xffffffffu, ((c4.x | c4.y | c4.z | c4.w) & 3) != 0);
            const uint32_t sRsBuf = sRs + (uint32_t)buf * (uint32_t)(rsBuf * 4);
            float sq[4], g[4];
            if (!edge)
                plf_taps<PS, GRAD, false>(a, i, i0, rows_here, r_lo, r_hi, code, wxs, sRsBuf, sLs, sV, lstart, vcol,
                                          vlane, vimg, sq, g);
            else
                plf_taps<PS, GRAD, true>(a, i, i0, rows_here, r_lo, r_hi, code, wxs, sRsBuf, sLs, sV, lstart, vcol,
                                         vlane, vimg, sq, g);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const bool m = own[t] && (code[t] & 4);
                if (m) {
                    tot += (double)sq[t];
                    if (grp == 0 && c == 0) cnt += 1.0;
                }
                g[t] = m ? g[t] : 0.f;
            }
            if (GRAD) {
                if (grp == 0) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) gkeep[t] = g[t];
                } else {
                    *reinterpret_cast<float4*>(Gs + (ii & 1) * (a.G - 1) * pitchG + (grp - 1) * pitchG + gidx) =
                        make_float4(g[0], g[1], g[2], g[3]);
                }
            }
            __syncthreads();
        }
        if (GRAD && grp == 0) {
            const int ii = rows_here;
            const float* gs = Gs + ((ii - 1) & 1) * (a.G - 1) * pitchG + gidx;
            float* gp = a.gpre + (size_t)b * HW + (size_t)(i0 + ii - 1) * W + s0;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (own[t]) {
                    float g = gkeep[t];
                    for (int gg = 0; gg < a.G - 1; ++gg) g += gs[gg * pitchG + t];
                    gp[t] = (c == 0 ? 0.f : gp[t]) + g;
                }
            }
        }
        // Fold rows still open at the end of the band: y in (i_last - p, i_last + p]
        const int i_last = i0 + rows_here - 1;
        for (int t = tid; t < 2 * T::P * W; t += nthreads) {
            const int yy = t / W, x = t - yy * W;
            const int y = i_last - T::P + 1 + yy;
            if (y < 0 || y >= H) continue;
            const int pid = (y + T::YB) >> 1, half = (y + T::YB) & 1;
            const float v = Vacc[((pid % T::VRING) * a.pitchV + (x & 3) * (a.pitchV >> 2) + (x >> 2)) * 2 + half];
            const bool shared_row = (i0 > 0 && y < i0 + T::P) || (i_last + 1 < H && y > i_last - T::P);
            if (shared_row) atomicAdd(vimg + (size_t)y * W + x, v);
            else vimg[(size_t)y * W + x] = v;
        }
    }
    const double bs = block_sum(tot, red);
    const double bc = block_sum(cnt, red);
    if (tid == 0) {
        const size_t r = (size_t)b * gridDim.x + band;
        a.partial[2 * r] = bs;
        a.partial[2 * r + 1] = bc;
    }
}

// host: geometry, band size and launch.  Returns AZ_ERR_BAD_ARG when the shape does not fit (the caller
// then runs the stand-alone loss and fold kernels).
template <int PS>
static int plf_launch(PlfArgs a, int B, int* nbands_out, cudaStream_t st) {
    using T = PlfGeom<PS>;
    const int W = a.W, H = a.H;
    a.npass = (W + T::S - 1) / T::S;
    const int max_warps = kPlfMaxThreads / 32;
    if (a.npass > max_warps) return AZ_ERR_BAD_ARG;
    int G = max_warps / a.npass;  // warp groups sharing the row pairs of a strip
    if (G > 3) G = 3;
    if (G > T::NP) G = T::NP;
    a.G = G;
    // plane stride (pitchR / 4 slots of 8 bytes) a multiple of 16 slots = 128 bytes: the bank of element xx is then
    // (xx >> 2) mod 16 whatever its plane, so two lanes of a half-warp only collide when their window starts fall
    // into the same aligned group of four columns -- one disparity step in four instead of (almost) every step
    a.pitchR = (T::OFF_R + W + T::P + 1 + 63) & ~63;
    a.pitchL = (W + 3 * T::P + 4 + 3) & ~3;
    a.pitchV = (W + 3) & ~3;
    a.pitchP = (a.npass - 1) * T::S + 128;
    const bool grad = a.gpre != nullptr;
    const size_t floats = (size_t)2 * T::NP * 2 * a.pitchR + (size_t)T::LRING * 2 * a.pitchL +
                          (size_t)T::VRING * 2 * a.pitchV + (size_t)4 * a.pitchP +
                          (grad ? (size_t)2 * (G - 1) * a.npass * 128 : 0);
    const size_t smem = floats * sizeof(float);
    if (smem > 232448 - 1024) return AZ_ERR_BAD_ARG;
    // band size: minimise waves x (rows per band + fixed per-band cost); a wave is every SM filled with as many
    // CTAs as its shared memory, registers (96 per thread) and thread slots allow (one at 544x960)
    const int threads = 32 * G * a.npass;
    int per_sm = (int)(232448 / (smem + 1024));
    if (per_sm > 65536 / (96 * threads)) per_sm = 65536 / (96 * threads);
    if (per_sm > 2048 / threads) per_sm = 2048 / threads;
    if (per_sm < 1) per_sm = 1;
    const int64_t slots = (int64_t)kNumSMs * per_sm;
    int best_nb = 1;
    double best = 1e30;
    for (int nb = 1; nb <= H; ++nb) {
        const int rows = (H + nb - 1) / nb;
        if (nb > 1 && rows < 2 * T::P) break;  // every full band >= 2p rows: a Fold row is shared by at most two bands
        const int nb_eff = (H + rows - 1) / rows;
        const double waves = (double)(((int64_t)B * nb_eff + slots - 1) / slots);
        const double cost = waves * (rows + 3.0);
        if (cost < best - 1e-9) { best = cost; best_nb = nb_eff; }
        if (rows <= 4) break;
    }
    a.band_rows = (H + best_nb - 1) / best_nb;
    const int nbands = (H + a.band_rows - 1) / a.band_rows;
    if (nbands > 65535) return AZ_ERR_BAD_ARG;
    *nbands_out = nbands;
    dim3 grid((unsigned)nbands, (unsigned)B);
    cudaError_t e;
    if (grad) {
        e = cudaFuncSetAttribute(patch_loss_fold_v2_kernel<PS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        patch_loss_fold_v2_kernel<PS, true><<<grid, threads, smem, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(patch_loss_fold_v2_kernel<PS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        patch_loss_fold_v2_kernel<PS, false><<<grid, threads, smem, st>>>(a);
    }
    return (int)cudaGetLastError();
}

// Entry used by az_reproj_loss_fwd (warp_reproj.cu).  vis must be zero-filled by the caller.
int plf_dispatch(const float* tgt, const float* src, const float* disp, float sign, const uint8_t* mask,
                 const float* lin_x, const float* lin_y, int ps, float* vis, float* gpre, double* partial, int B, int C,
                 int H, int W, int* nbands_out, cudaStream_t st) {
    PlfArgs a;
    a.tgt = tgt; a.src = src; a.disp = disp; a.mask = mask; a.lin_x = lin_x; a.lin_y = lin_y;
    a.vis = vis; a.gpre = gpre; a.partial = partial; a.sign = sign;
    a.C = C; a.H = H; a.W = W;
    a.vec = (W % 4 == 0) && aligned16(tgt) && aligned16(src) && aligned16(disp) && aligned16(vis) &&
            (mask == nullptr || (reinterpret_cast<uintptr_t>(mask) & 3u) == 0);
    switch (ps) {
        case 3: return plf_launch<3>(a, B, nbands_out, st);
        case 5: return plf_launch<5>(a, B, nbands_out, st);
        case 7: return plf_launch<7>(a, B, nbands_out, st);
        case 9: return plf_launch<9>(a, B, nbands_out, st);
        case 11: return plf_launch<11>(a, B, nbands_out, st);
        case 13: return plf_launch<13>(a, B, nbands_out, st);
        default: return AZ_ERR_BAD_ARG;
    }
}

}  // namespace az
