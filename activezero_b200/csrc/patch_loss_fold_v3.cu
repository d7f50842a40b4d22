// Patch reprojection loss AND its Fold image in one pass, round-2 form ("v3").
// (/root/reference/utils/reprojection.py:99-127.)
//
//   loss = mean_{mask} (Wu - Lu)^2,   Wu = apply_disparity(Unfold(R), -disp),  Lu = Unfold(L)   (:102-118)
//   vis  = Fold(Wu) cropped                                                                     (:120-125)
//
// Why a second form.  Round 1's kernel (patch_loss_fold.cu) maps a lane to four adjacent SOURCE
// pixels; every lane then walks its own 12-column window of the blended source rows, starting at
// its own x0(disparity).  On a field whose integer disparity changes from pixel to pixel (the
// bench's, and SURVEY §8d's U(0,64)) those are 16 random 8-byte gathers per shared-memory
// wavefront: ncu counted 1.95x the ideal wavefronts (2.4x on the window loads) and the kernel sat
// at the LSU limit.  Random per-lane gathers cannot be made conflict-free by a layout.
//
// Here the lanes of a QUARTER-WARP share one group of four adjacent sources and differ in the tap
// ROW PAIR they process (lane = row pair, 6 of 8 lanes busy at ps = 11).  Every window load is a
// 128-bit load of two columns x two rows; the eight lanes of a quarter-warp read the same columns of
// different rows, and the row pitch is chosen so that those land in different 16-byte banks: the
// loads are conflict-free BY CONSTRUCTION, whatever the disparities are.  The rest follows from that
// mapping:
//   * a quarter-warp is a "walker" over 4n consecutive groups.  Its target window Lw and its Fold
//     accumulator are 16-entry circular register files indexed statically (the walk is unrolled
//     four groups at a time): per group the walker loads only the 4 new target columns and flushes
//     only the 4 Fold columns that no later group of the walk touches -- no shuffles, no halo;
//   * the Fold partial sums of a warp live in a PRIVATE strip of the shared ring (its own columns
//     plus ps-1 of overhang), so warps never write the same word; the strips are added in a fixed
//     order when a finished row is written to HBM (deterministic);
//   * the window start parity (128-bit loads are 16-byte aligned) is resolved with selects on the
//     loaded registers.
// Per pixel: 7 (window) + ~0.6 (target) + ~1.6 (Fold) 128-byte shared-memory wavefronts instead of
// 16.6 measured for round 1's kernel.
//
// Everything else is as in round 1: one CTA owns a band of source rows of one image; per row it
// stages the vertically blended source rows (ys depends on the row only), the per-pixel sampling
// parameters and -- as a ring -- the target rows; rows are processed two at a time as packed
// fp32x2 values (FFMA2/FADD2), absolute even/odd row pairs, one dummy half-row per source row;
// finished Fold rows leave through a ring; rows shared with the neighbouring band get one atomic add
// per element onto zero-filled memory (two addends: deterministic).
#include "common.cuh"

namespace az {

typedef unsigned long long u64;

namespace v3 {

__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ void lds128(uint32_t addr, u64& a, u64& b) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ void sts128(uint32_t addr, u64 a, u64 b) {
    asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, float lo, float hi) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(lo), "f"(hi) : "memory");
}
__device__ __forceinline__ u64 sel2(bool p, u64 a, u64 b) {  // p ? a : b
    float al, ah, bl, bh;
    upk2(a, al, ah);
    upk2(b, bl, bh);
    return pk2(p ? al : bl, p ? ah : bh);
}

struct Args {
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
    int NW;         // warps per CTA
    int n;          // a walker (quarter-warp) covers 4n groups; a warp 16n groups = 64n sources
    int nbuf;       // 1 or 2 buffers for the blended rows / parameters
    int vec;        // W % 4 == 0 and every image pointer 16-byte aligned
    int PR, PL, PV; // row pitches (in 8-byte elements) of Rs, Ls, Vacc; PR/2, PL/2, PV/2 odd
    int GT;         // groups per x-tile incl. padding = 16 n NW
    int ntile;      // x-tiles per row (1 unless the row does not fit: W > ~1400 at ps = 11); a tile is 4 GT sources
    int sb_rows;    // rows of the tile-boundary side buffer = band_rows + 2p
};

template <int PS>
struct Geom {
    static constexpr int P = (PS - 1) / 2;
    static constexpr int NP = (PS + 1) / 2;               // row pairs per source row (<= 8: one lane each)
    static constexpr int NV = PS + 3;                     // columns a lane holds per source (even): window + parity slack
    static constexpr int OFF_R = (PS + 3 + 3) & ~3;       // column of x = 0 in a staged source row (>= NV zeros before it)
    static constexpr int RING = NP + 1;                   // target / Fold row pairs resident
    static constexpr int YB = 16;                         // bias that keeps (row + YB) non-negative
    static_assert(NP <= 8 && NV <= 16, "ps <= 13");
};

__device__ __forceinline__ float4 load4(const float* __restrict__ img, int y, int x, int H, int W, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y >= 0 && y < H && x >= 0 && x < W) {
        const float* p = img + (size_t)y * W + x;
        if (vec) {
            v = __ldg(reinterpret_cast<const float4*>(p));
        } else {
            v.x = __ldg(p);
            if (x + 1 < W) v.y = __ldg(p + 1);
            if (x + 2 < W) v.z = __ldg(p + 2);
            if (x + 3 < W) v.w = __ldg(p + 3);
        }
    }
    return v;
}

// ---- staging of one source row i into buffer `buf` -------------------------------------------
template <int PS>
__device__ __forceinline__ void stage_row(const Args& a, const float* __restrict__ sp, const float* __restrict__ tp,
                                          const float* __restrict__ dimg, const uint8_t* __restrict__ mimg, int i,
                                          int X0, int buf, bool first, uint32_t sRs, uint32_t sLs, float* PsW, int* PsC) {
    using T = Geom<PS>;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int H = a.H, W = a.W;
    const int nq = (W + 3) >> 2;
    const bool vec = a.vec != 0;
    const int e0 = (i - T::P + T::YB) & 1;
    const int pid0 = (i - T::P + T::YB) >> 1;

    // (1) blended source rows: Rs[buf][r][OFF_R + x] = (row 2r - e0, row 2r + 1 - e0) of the tap window; the half
    //     outside the ps tap rows (ky = -1 or ky = ps) is stored as zero so that it adds nothing to the Fold sums
    for (int m = tid; m < nq; m += nthreads) {
        const Axis ay = make_axis(sample_pos(__ldg(a.lin_y + i), 0.0f, (float)H), H);
        const float ay0 = ay.v0 ? ay.e : 0.f, ay1 = ay.v1 ? ay.w : 0.f;
        const int ytop = ay.i0 - T::P - e0;  // image row feeding (r = 0, low half) through corner y0
        const int x = 4 * m;
        uint32_t dst = sRs + ((uint32_t)(buf * T::NP) * (uint32_t)a.PR + (uint32_t)(T::OFF_R + x)) * 8u;
        float4 prev = load4(sp, ytop, x, H, W, vec);
#pragma unroll
        for (int r = 0; r < T::NP; ++r) {
            const float4 mid = load4(sp, ytop + 2 * r + 1, x, H, W, vec);
            const float4 nxt = load4(sp, ytop + 2 * r + 2, x, H, W, vec);
            const float zl = (e0 == 1 && r == 0) ? 0.f : 1.f, zh = (e0 == 0 && r == T::NP - 1) ? 0.f : 1.f;
            const float l0 = zl * fmaf(ay1, mid.x, ay0 * prev.x), h0 = zh * fmaf(ay1, nxt.x, ay0 * mid.x);
            const float l1 = zl * fmaf(ay1, mid.y, ay0 * prev.y), h1 = zh * fmaf(ay1, nxt.y, ay0 * mid.y);
            const float l2 = zl * fmaf(ay1, mid.z, ay0 * prev.z), h2 = zh * fmaf(ay1, nxt.z, ay0 * mid.z);
            const float l3 = zl * fmaf(ay1, mid.w, ay0 * prev.w), h3 = zh * fmaf(ay1, nxt.w, ay0 * mid.w);
            sts128(dst, pk2(l0, h0), pk2(l1, h1));
            sts128(dst + 16u, pk2(l2, h2), pk2(l3, h3));
            dst += (uint32_t)a.PR * 8u;
            prev = nxt;
        }
    }

    // (2) per-source sampling parameters (threads taken from the other end of the CTA)
    const int tid2 = nthreads - 1 - tid;
    const int gq = a.GT;  // quads incl. padding groups: their sources are invalid
    for (int m = tid2; m < gq; m += nthreads) {
        const int xl = 4 * m, x = X0 + xl;  // tile-local / image column of the quad
        const float4 d4 = load4(dimg, i, x, H, W, vec);
        const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
        uint32_t m4 = 0x01010101u;
        if (mimg != nullptr && x < W) {
            const uint8_t* mp = mimg + (size_t)i * W + x;
            if (vec) {
                m4 = __ldg(reinterpret_cast<const uint32_t*>(mp));
            } else {
                m4 = 0;
                for (int t = 0; t < 4; ++t)
                    if (x + t < W) m4 |= (uint32_t)(__ldg(mp + t) != 0) << (8 * t);
            }
        }
        int cc[4];
        float ww[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            int base = 0, flags = 0, msk = 0;
            float wx = 0.f;
            if (x + t < W) {
                const Axis ax =
                    make_axis(sample_pos(__ldg(a.lin_x + x + t), __fdiv_rn(a.sign * dd[t], (float)W), (float)W), W);
                if (ax.v0 || ax.v1) {
                    base = T::OFF_R + min(max(ax.i0, -1), W - 1) - T::P;
                    wx = ax.w;
                    flags = (ax.v0 ? 0 : 1) | (ax.v1 ? 0 : 2);
                }
                msk = ((m4 >> (8 * t)) & 0xffu) ? 4 : 0;
            }
            cc[t] = (base << 3) | msk | flags;
            ww[t] = wx;
        }
        *reinterpret_cast<int4*>(PsC + buf * 4 * a.GT + xl) = make_int4(cc[0], cc[1], cc[2], cc[3]);
        *reinterpret_cast<float4*>(PsW + buf * 4 * a.GT + xl) = make_float4(ww[0], ww[1], ww[2], ww[3]);
    }

    // (3) the target row pair(s) entering the ring: Ls[slot][P + xl] = (row y, row y + 1), y = 2 pid - YB, for the
    //     tile-local columns xl in [-p, 4 GT + p) (image columns X0 + xl; zero outside the image)
    const int np_new = first ? T::NP : (e0 == 0 ? 1 : 0);  // a new pair enters when e0 flips to 0
    for (int nn = 0; nn < np_new; ++nn) {
        const int pid = first ? pid0 + nn : pid0 + T::NP - 1;
        const int y = 2 * pid - T::YB;
        const uint32_t rowb = sLs + ((uint32_t)(pid % T::RING) * (uint32_t)a.PL) * 8u;
        for (int m = tid2; m < a.GT + 4; m += nthreads) {
            const int xl = 4 * (m - 2);
            const float4 u = load4(tp, y, X0 + xl, H, W, vec);
            const float4 v = load4(tp, y + 1, X0 + xl, H, W, vec);
            const float uu[4] = {u.x, u.y, u.z, u.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int idx = xl + t + T::P;
                if (idx >= 0 && idx < a.PL) sts64(rowb + (uint32_t)idx * 8u, uu[t], vv[t]);
            }
        }
    }
}

// ---- one group of a walk: U = position of the group in the unrolled-by-4 walk ---------------------
// accV / Lw: 16-entry circular register files; this group's window occupies entries (4U + m) & 15, m = 0..NV-1.
template <int PS, bool GRAD, bool EDGE, int U>
__device__ __forceinline__ void walk_group(const int (&code)[4], const float (&wxs)[4], uint32_t rsrow, u64 hm2,
                                           u64 (&accV)[16], const u64 (&Lw)[16], float (&sq)[4], float (&gr)[4]) {
    using T = Geom<PS>;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int xb = code[t] >> 3;
        const bool odd = (xb & 1) != 0;
        const uint32_t addr = rsrow + (uint32_t)(xb & ~1) * 8u;
        u64 V[T::NV];
#pragma unroll
        for (int j = 0; j < T::NV / 2; ++j) lds128(addr + 16u * j, V[2 * j], V[2 * j + 1]);
        u64 A[PS + 1];
#pragma unroll
        for (int k = 0; k <= PS; ++k) A[k] = sel2(odd, V[k + 1], V[k]);
        const u64 wx2 = pk2(wxs[t], wxs[t]);
        u64 m0 = 0ull, m1 = 0ull;
        if (EDGE) {
            const float f0 = (code[t] & 1) ? 0.f : 1.f, f1 = (code[t] & 2) ? 0.f : 1.f;
            m0 = pk2(f0, f0);
            m1 = pk2(f1, f1);
        }
        u64 sq2 = 0ull, g2 = 0ull;
#pragma unroll
        for (int k = 0; k < PS; ++k) {
            u64 am = A[k], nm = A[k + 1];
            if (EDGE) {
                am = mul2(am, m0);
                nm = mul2(nm, m1);
            }
            const u64 dk = sub2(nm, am);
            const u64 w = fma2(wx2, dk, am);
            const u64 e = sub2(w, Lw[(4 * U + t + k) & 15]);
            sq2 = fma2(e, e, sq2);
            if (GRAD) g2 = fma2(e, dk, g2);
            accV[(4 * U + t + k) & 15] = add2(accV[(4 * U + t + k) & 15], w);
        }
        float lo, hi;
        upk2(mul2(sq2, hm2), lo, hi);
        sq[t] = lo + hi;
        if (GRAD) {
            upk2(mul2(g2, hm2), lo, hi);
            gr[t] = lo + hi;
        } else {
            gr[t] = 0.f;
        }
    }
}


// Writes Fold row y of x-tile `xt` to HBM from ring slot `vrow_f` (floats; + half selects the row of the pair).
// Column x (tile-local xl = x - X0) sums, in a fixed order, the private strips that cover it: the strip to the left
// (its right overhang), its own, the strip to the right (its left overhang).  Between x-tiles the 2p columns
// around the seam are handed over through the side buffer SB: tile t keeps them (their sums over its own sources)
// instead of emitting its last p columns, tile t+1 adds them to its own share and emits columns [X0 - p, ...).
// Every element of the image is therefore written once per CTA (rows shared with the neighbouring band: one atomic
// add per CTA onto zero-filled memory, two addends).
template <int PS>
__device__ __forceinline__ void emit_row(const Args& a, float* __restrict__ vslot, int half, bool recycle,
                                         const float* __restrict__ sb_in, float* __restrict__ sb_out,
                                         float* __restrict__ vimg_row, bool shared_row, bool store, int xt, int X0,
                                         int RWd) {
    using T = Geom<PS>;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int SW = 64 * a.n, TW = 4 * a.GT, W = a.W;
    const bool last = xt + 1 == a.ntile;
    const float* vr = vslot + half;
    const float2 z2 = make_float2(0.f, 0.f);
    // tile-local columns [-p, TW + p): every entry of the slot is read (and, when the slot is recycled, cleared) by
    // exactly one thread -- no barrier between the reads and the clears
    for (int xl = tid - T::P; xl < TW + T::P; xl += nthreads) {
        const int w0 = min(max(xl, 0) / SW, a.NW - 1), xw = xl - w0 * SW;  // xw in [-p, SW + p)
        const int li = w0 * RWd + xw + T::P;
        const bool from_left = xw < T::P && w0 > 0;                       // right overhang of the strip to the left
        const bool from_right = xw >= SW - T::P && xw < SW && w0 + 1 < a.NW;  // left overhang of the strip to the right
        float v = 0.f;
        if (xt > 0 && xl < T::P) v += sb_in[xl + T::P];  // the previous tile's share of the seam columns
        if (from_left) v += vr[2 * (li - RWd + SW)];
        v += vr[2 * li];
        if (from_right) v += vr[2 * (li + RWd - SW)];
        if (recycle) {
            if (from_left) *reinterpret_cast<float2*>(vslot + 2 * (li - RWd + SW)) = z2;
            *reinterpret_cast<float2*>(vslot + 2 * li) = z2;
            if (from_right) *reinterpret_cast<float2*>(vslot + 2 * (li + RWd - SW)) = z2;
        }
        const int x = X0 + xl;
        if (!last && xl >= TW - T::P) {
            sb_out[xl - (TW - T::P)] = v;  // seam: handed to the next tile (the other half of the double buffer)
        } else if (store && x >= 0 && x < W) {
            float* op = vimg_row + x;
            if (shared_row) atomicAdd(op, v);
            else *op = v;
        }
    }
}

constexpr int kMaxThreads = 480;  // 15 warps: 136 registers per thread (the walk keeps ~90 live packed values)

template <int PS, bool GRAD>
__global__ void __launch_bounds__(kMaxThreads, 1) patch_loss_fold_v3_kernel(const Args a) {
    using T = Geom<PS>;
    extern __shared__ __align__(16) float sm[];
    __shared__ double red[32];
    const int tid = threadIdx.x, nthreads = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int q = lane >> 3, r8 = lane & 7;
    const bool active = r8 < T::NP;
    const int rr = active ? r8 : T::NP - 1;  // idle lanes shadow the last row pair (same addresses: broadcast) and never store
    const int band = blockIdx.x, b = blockIdx.y;
    const int H = a.H, W = a.W;
    const int i0 = band * a.band_rows;
    const int rows_here = min(a.band_rows, H - i0);
    const size_t HW = (size_t)H * W;

    // shared-memory carve-up (8-byte elements unless noted)
    const int rsBuf = T::NP * a.PR;
    float* Rs = sm;                                                  // [nbuf][NP][PR][2]
    float* Ls = Rs + (size_t)2 * a.nbuf * rsBuf;                     // [RING][PL][2]
    float* Vacc = Ls + (size_t)2 * T::RING * a.PL;                   // [RING][PV][2]
    float* PsW = Vacc + (size_t)2 * T::RING * a.PV;                  // [nbuf][4 GT] horizontal weight of each source
    int* PsC = reinterpret_cast<int*>(PsW + (size_t)a.nbuf * 4 * a.GT);  // [nbuf][4 GT] window start << 3 | mask << 2 | edge flags
    const int total_floats = 2 * a.nbuf * rsBuf + 2 * T::RING * a.PL + 2 * T::RING * a.PV + 2 * a.nbuf * 4 * a.GT;
    float* SB = sm + total_floats;                                   // [2][sb_rows][2p] sums handed from x-tile t to t+1 (double buffer)
    const uint32_t sRs = (uint32_t)__cvta_generic_to_shared(Rs);
    const uint32_t sLs = (uint32_t)__cvta_generic_to_shared(Ls);
    const uint32_t sV = (uint32_t)__cvta_generic_to_shared(Vacc);

    // walker geometry: this quarter-warp walks the groups gw0 .. gw0 + 4n - 1 of its warp's 16n
    const int RWd = 64 * a.n + 2 * T::P;                  // width of a warp's private Fold strip
    const int gw0 = q * 4 * a.n;                          // first local group of the walk
    const int gglob0 = warp * 16 * a.n + gw0;             // ... and its global index
    const uint32_t vstrip = (uint32_t)(warp * RWd);       // this warp's strip in a Vacc row

    const float* dimg = a.disp + (size_t)b * HW;
    const uint8_t* mimg = a.mask == nullptr ? nullptr : a.mask + (size_t)b * HW;

    double tot = 0.0, cnt = 0.0;
    for (int c = 0; c < a.C; ++c) {
        const float* sp = a.src + ((size_t)b * a.C + c) * HW;
        const float* tp = a.tgt + ((size_t)b * a.C + c) * HW;
        float* vimg = a.vis + ((size_t)b * a.C + c) * HW;
        for (int xt = 0; xt < a.ntile; ++xt) {
        const int X0 = xt * 4 * a.GT;  // first source (= image column) of this x-tile
        __syncthreads();
        for (int t = tid; t < total_floats / 4; t += nthreads)
            reinterpret_cast<float4*>(sm)[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        stage_row<PS>(a, sp, tp, dimg, mimg, i0, X0, 0, true, sRs, sLs, PsW, PsC);
        __syncthreads();

        for (int ii = 0; ii < rows_here; ++ii) {
            const int i = i0 + ii, buf = a.nbuf == 2 ? (ii & 1) : 0;
            if (a.nbuf == 2 && ii + 1 < rows_here)
                stage_row<PS>(a, sp, tp, dimg, mimg, i + 1, X0, buf ^ 1, false, sRs, sLs, PsW, PsC);

            // ---------------- taps of row i ----------------
            const int e0 = (i - T::P + T::YB) & 1;
            const int pid0 = (i - T::P + T::YB) >> 1;
            const int slot = (pid0 + rr) % T::RING;
            // half of this lane's pair that lies outside the ps tap rows (or the whole pair for an idle lane)
            const float hl = (!active || (e0 == 1 && rr == 0)) ? 0.f : 1.f;
            const float hh = (!active || (e0 == 0 && rr == T::NP - 1)) ? 0.f : 1.f;
            const u64 hm2 = pk2(hl, hh);
            const uint32_t rsrow = sRs + ((uint32_t)(buf * T::NP + rr) * (uint32_t)a.PR) * 8u;
            const uint32_t lsrow = sLs + ((uint32_t)slot * (uint32_t)a.PL) * 8u;
            const uint32_t vrow = sV + ((uint32_t)slot * (uint32_t)a.PV + vstrip) * 8u;
            const int* pc = PsC + buf * 4 * a.GT;
            const float* pw = PsW + buf * 4 * a.GT;
            float* gprow = GRAD ? a.gpre + (size_t)b * HW + (size_t)i * W : nullptr;

            u64 accV[16], Lw[16];
            float rowsq = 0.f, rowcnt = 0.f;  // this lane's share of the row: <= 16 n terms, one fp64 add per row
#pragma unroll
            for (int m = 0; m < 16; ++m) accV[m] = 0ull;
            // target window of the first group: entries 0 .. NV-1
#pragma unroll
            for (int j = 0; j < T::NV / 2; ++j) lds128(lsrow + (uint32_t)(4 * gglob0 + 2 * j) * 8u, Lw[2 * j], Lw[2 * j + 1]);
#pragma unroll
            for (int m = T::NV; m < 16; ++m) Lw[m] = 0ull;

            for (int blk = 0; blk < a.n; ++blk) {
#define AZ_V3_GROUP(U)                                                                                                  \
    {                                                                                                                   \
        const int gl = gw0 + 4 * blk + U, gg = gglob0 + 4 * blk + U;                                                    \
        const int4 c4 = *reinterpret_cast<const int4*>(pc + 4 * gg);                                                    \
        const float4 w4 = *reinterpret_cast<const float4*>(pw + 4 * gg);                                                \
        const int code[4] = {c4.x, c4.y, c4.z, c4.w};                                                                   \
        const float wxs[4] = {w4.x, w4.y, w4.z, w4.w};                                                                  \
        if (4 * blk + U > 0) { /* the four new target columns of this group's window */                                 \
            lds128(lsrow + (uint32_t)(4 * gg + T::NV - 4) * 8u, Lw[(4 * U + T::NV - 4) & 15], Lw[(4 * U + T::NV - 3) & 15]); \
            lds128(lsrow + (uint32_t)(4 * gg + T::NV - 2) * 8u, Lw[(4 * U + T::NV - 2) & 15], Lw[(4 * U + T::NV - 1) & 15]); \
        }                                                                                                               \
        const bool edge = __any_sync(0xffffffffu, ((c4.x | c4.y | c4.z | c4.w) & 3) != 0);                              \
        float sq[4], gr[4];                                                                                             \
        if (!edge) walk_group<PS, GRAD, false, U>(code, wxs, rsrow, hm2, accV, Lw, sq, gr);                             \
        else walk_group<PS, GRAD, true, U>(code, wxs, rsrow, hm2, accV, Lw, sq, gr);                                    \
        /* the four lowest Fold columns of the window are final for this walk: add them to the strip */                 \
        {                                                                                                               \
            const uint32_t va = vrow + (uint32_t)(4 * gl) * 8u;                                                         \
            u64 v0, v1, v2, v3;                                                                                         \
            lds128(va, v0, v1);                                                                                         \
            lds128(va + 16u, v2, v3);                                                                                   \
            v0 = add2(v0, accV[(4 * U + 0) & 15]);                                                                      \
            v1 = add2(v1, accV[(4 * U + 1) & 15]);                                                                      \
            v2 = add2(v2, accV[(4 * U + 2) & 15]);                                                                      \
            v3 = add2(v3, accV[(4 * U + 3) & 15]);                                                                      \
            if (active) {                                                                                               \
                sts128(va, v0, v1);                                                                                     \
                sts128(va + 16u, v2, v3);                                                                               \
            }                                                                                                           \
            accV[(4 * U + 0) & 15] = 0ull;                                                                              \
            accV[(4 * U + 1) & 15] = 0ull;                                                                              \
            accV[(4 * U + 2) & 15] = 0ull;                                                                              \
            accV[(4 * U + 3) & 15] = 0ull;                                                                              \
        }                                                                                                               \
        /* loss: every lane keeps its own row pair's share; the count is taken once per pixel */                        \
        _Pragma("unroll") for (int t = 0; t < 4; ++t) {                                                                 \
            if (code[t] & 4) {                                                                                          \
                rowsq += sq[t];                                                                                         \
                rowcnt += 1.0f;                                                                                         \
            }                                                                                                           \
        }                                                                                                               \
        if (GRAD) { /* sum the row pairs of each source over the 8 lanes: transposing butterfly, 4 shuffles */          \
            const bool b2 = (lane & 4) != 0, b1 = (lane & 2) != 0;                                                      \
            float k0 = b2 ? gr[2] : gr[0], k1 = b2 ? gr[3] : gr[1];                                                     \
            const float s0 = b2 ? gr[0] : gr[2], s1 = b2 ? gr[1] : gr[3];                                               \
            k0 += __shfl_xor_sync(0xffffffffu, s0, 4);                                                                  \
            k1 += __shfl_xor_sync(0xffffffffu, s1, 4);                                                                  \
            float k = b1 ? k1 : k0;                                                                                     \
            k += __shfl_xor_sync(0xffffffffu, b1 ? k0 : k1, 2);                                                         \
            k += __shfl_xor_sync(0xffffffffu, k, 1);                                                                    \
            const int t = (b2 ? 2 : 0) + (b1 ? 1 : 0);                                                                  \
            const int j = X0 + 4 * gg + t;                                                                              \
            if ((lane & 1) == 0 && j < W) {                                                                             \
                const int cd = t == 0 ? c4.x : (t == 1 ? c4.y : (t == 2 ? c4.z : c4.w));                                \
                const float g = (cd & 4) ? k : 0.f;                                                                     \
                gprow[j] = (c == 0 ? 0.f : gprow[j]) + g;                                                               \
            }                                                                                                           \
        }                                                                                                               \
    }
                AZ_V3_GROUP(0)
                AZ_V3_GROUP(1)
                AZ_V3_GROUP(2)
                AZ_V3_GROUP(3)
#undef AZ_V3_GROUP
            }
            tot += (double)rowsq;
            if (r8 == 0 && c == 0) cnt += (double)rowcnt;
            // tail of the walk: the NV - 4 columns beyond the last group's own four (entries 0 .. NV-5 after the
            // last U = 3 group: its window started at entry 12)
            {
                const uint32_t va = vrow + (uint32_t)(4 * (gw0 + 4 * a.n)) * 8u;
#pragma unroll
                for (int j = 0; j < (T::NV - 4) / 2; ++j) {
                    u64 v0, v1;
                    lds128(va + 16u * j, v0, v1);
                    v0 = add2(v0, accV[(2 * j) & 15]);
                    v1 = add2(v1, accV[(2 * j + 1) & 15]);
                    if (active) sts128(va + 16u * j, v0, v1);
                }
            }
            __syncthreads();

            // ---------------- Fold row y = i - p is complete: write it out ----------------
            {
                const int y = i - T::P;
                const int half = e0;  // ky = 0 is (r = 0, half e0) of pair pid0
                const bool shared_row = (i0 > 0 && y < i0 + T::P) || (i0 + rows_here < H && y > i0 + rows_here - 1 - T::P);
                float* vz = Vacc + ((size_t)(pid0 % T::RING) * a.PV) * 2;
                // half == 1: both rows of the pair are out -> the slot is cleared for the pair that enters >= 3 rows later
                const size_t sbo = (size_t)(y - (i0 - T::P)) * 2 * T::P, sbh = (size_t)a.sb_rows * 2 * T::P;
                emit_row<PS>(a, vz, half, half == 1, SB + ((xt + 1) & 1) * sbh + sbo, SB + (xt & 1) * sbh + sbo,
                             vimg + (size_t)max(y, 0) * W, shared_row, y >= 0, xt, X0, RWd);
            }
            if (a.nbuf == 1 && ii + 1 < rows_here) {
                stage_row<PS>(a, sp, tp, dimg, mimg, i + 1, X0, 0, false, sRs, sLs, PsW, PsC);
                __syncthreads();
            }
        }
        __syncthreads();
        // Fold rows still open at the end of the band: y in (i_last - p, i_last + p]
        const int i_last = i0 + rows_here - 1;
        for (int yy = 0; yy < 2 * T::P; ++yy) {
            const int y = i_last - T::P + 1 + yy;
            if (y < 0 || y >= H) continue;
            const int pid = (y + T::YB) >> 1, half = (y + T::YB) & 1;
            const bool shared_row = (i0 > 0 && y < i0 + T::P) || (i_last + 1 < H && y > i_last - T::P);
            const size_t sbo = (size_t)(y - (i0 - T::P)) * 2 * T::P, sbh = (size_t)a.sb_rows * 2 * T::P;
            emit_row<PS>(a, Vacc + ((size_t)(pid % T::RING) * a.PV) * 2, half, false, SB + ((xt + 1) & 1) * sbh + sbo,
                         SB + (xt & 1) * sbh + sbo, vimg + (size_t)y * W, shared_row, true, xt, X0, RWd);
        }
        }  // x-tiles
    }
    const double bs = block_sum(tot, red);
    const double bc = block_sum(cnt, red);
    if (tid == 0) {
        const size_t r = (size_t)b * gridDim.x + band;
        a.partial[2 * r] = bs;
        a.partial[2 * r + 1] = bc;
    }
}

// host: geometry, band size and launch.  Returns AZ_ERR_BAD_ARG when the shape does not fit (the caller then
// runs round 1's kernels).
template <int PS>
static int launch(Args a, int B, int* nbands_out, cudaStream_t st) {
    using T = Geom<PS>;
    const int W = a.W, H = a.H;
    const int G = (W + 3) / 4;
    const bool grad = a.gpre != nullptr;
    const int max_warps = kMaxThreads / 32;
    auto round_pitch = [](int v) { v = (v + 1) & ~1; return (v & 3) == 2 ? v : v + 2; };  // even, half of it odd
    int best_n = 0;
    size_t smem = 0;
    // geometry: as few x-tiles as possible, then as few groups per walker as possible, double-buffered if it fits.
    // The band size (hence the side buffer) is fixed afterwards; 2p * (H + 2p) floats bound it here.
    for (int ntile = 1; ntile <= 4 && best_n == 0; ++ntile) {
        const int Gt = (G + ntile - 1) / ntile;
        for (int n = 1; n <= 8 && best_n == 0; ++n) {
            const int NW = (Gt + 16 * n - 1) / (16 * n);
            if (NW > max_warps) continue;
            a.n = n;
            a.NW = NW;
            a.GT = 16 * n * NW;
            a.ntile = (G + a.GT - 1) / a.GT;
            a.PR = round_pitch(T::OFF_R + 4 * G + T::NV);
            a.PL = round_pitch(4 * a.GT + T::NV);
            a.PV = round_pitch(NW * (64 * n + 2 * T::P));
            const size_t sb = a.ntile > 1 ? (size_t)4 * T::P * (H + 2 * T::P) : 0;
            for (int nbuf = 2; nbuf >= 1; --nbuf) {
                const size_t fl = (size_t)2 * nbuf * T::NP * a.PR + (size_t)2 * T::RING * a.PL +
                                  (size_t)2 * T::RING * a.PV + (size_t)2 * nbuf * 4 * a.GT;
                const size_t sb_cap = sb < 16384 ? sb : 16384;  // the side buffer only needs band_rows + 2p rows
                if ((fl + sb_cap) * sizeof(float) <= 232448 - 1024) {
                    a.nbuf = nbuf;
                    smem = ((fl + 3) & ~(size_t)3) * sizeof(float);
                    best_n = n;
                    break;
                }
            }
        }
    }
    if (best_n == 0) return AZ_ERR_BAD_ARG;
    const int threads = 32 * a.NW;
    // band size: minimise waves x (rows per band + fixed per-band cost)
    int per_sm = (int)(232448 / (smem + 1024));
    if (per_sm > 65536 / (128 * threads)) per_sm = 65536 / (128 * threads);
    if (per_sm > 2048 / threads) per_sm = 2048 / threads;
    if (per_sm < 1) per_sm = 1;
    // AZ_PATCH_SMS (developer knob): size the bands for a partition of the GPU (a green context with fewer SMs,
    // benchmarks/sm_partition_pipeline.py) instead of all 148 SMs
    int sms = tuning("AZ_PATCH_SMS", kNumSMs);
    if (sms < 1 || sms > kNumSMs) sms = kNumSMs;
    const int64_t slots = (int64_t)sms * per_sm;
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
    a.sb_rows = a.band_rows + 2 * T::P;
    if (a.ntile > 1) {
        smem += (size_t)2 * a.sb_rows * 2 * T::P * sizeof(float);
        if (smem > 232448 - 1024) return AZ_ERR_BAD_ARG;
    }
    const int nbands = (H + a.band_rows - 1) / a.band_rows;
    if (nbands > 65535) return AZ_ERR_BAD_ARG;
    *nbands_out = nbands;
    dim3 grid((unsigned)nbands, (unsigned)B);
    cudaError_t e;
    if (grad) {
        e = cudaFuncSetAttribute(patch_loss_fold_v3_kernel<PS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        patch_loss_fold_v3_kernel<PS, true><<<grid, threads, smem, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(patch_loss_fold_v3_kernel<PS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        patch_loss_fold_v3_kernel<PS, false><<<grid, threads, smem, st>>>(a);
    }
    return (int)cudaGetLastError();
}

}  // namespace v3

// Entry used by az_reproj_loss_fwd (warp_reproj.cu).  vis must be zero-filled by the caller.
int plf3_dispatch(const float* tgt, const float* src, const float* disp, float sign, const uint8_t* mask,
                  const float* lin_x, const float* lin_y, int ps, float* vis, float* gpre, double* partial, int B, int C,
                  int H, int W, int* nbands_out, cudaStream_t st) {
    v3::Args a;
    a.tgt = tgt; a.src = src; a.disp = disp; a.mask = mask; a.lin_x = lin_x; a.lin_y = lin_y;
    a.vis = vis; a.gpre = gpre; a.partial = partial; a.sign = sign;
    a.C = C; a.H = H; a.W = W;
    a.vec = (W % 4 == 0) && aligned16(tgt) && aligned16(src) && aligned16(disp) && aligned16(vis) &&
            (mask == nullptr || (reinterpret_cast<uintptr_t>(mask) & 3u) == 0);
    switch (ps) {
        case 3: return v3::launch<3>(a, B, nbands_out, st);
        case 5: return v3::launch<5>(a, B, nbands_out, st);
        case 7: return v3::launch<7>(a, B, nbands_out, st);
        case 9: return v3::launch<9>(a, B, nbands_out, st);
        case 11: return v3::launch<11>(a, B, nbands_out, st);
        case 13: return v3::launch<13>(a, B, nbands_out, st);
        default: return AZ_ERR_BAD_ARG;
    }
}

}  // namespace az
