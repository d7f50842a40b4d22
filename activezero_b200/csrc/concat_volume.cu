// Concat cost volume, forward and backward.
// Reference: /root/reference/nets/psmnet/psmnet.py:151-165 (= psmnet_3.py:149-163).
//
// Forward is a pure store stream (2C*Dq*H*W floats out for 2*C*H*W floats in): a CTA owns a
// contiguous 4 KB run of one (b, channel) feature plane and writes it to eight consecutive
// disparity planes with coalesced, streaming 128-bit stores (left half: the float4 stays in
// registers; right half: the shifted row is the static window of two aligned quads read from the
// L2-resident feature plane).  Short-lived CTAs keep the store stream closer to the write-only
// ceiling of the part (7.5 TB/s = memset; the copy-measured "peak" of 6.5 TB/s is a read+write mix).
//
// Backward is a pure load stream: one thread per float4 of gL / gR walks the Dq disparity planes
// with 8 independent 128-bit streaming loads in flight (left half: straight down; right half:
// along the x+i diagonal, read as two aligned float4 per plane and a static register window),
// accumulating in a fixed order: gather-style and atomic-free.  The zero triangle x < i of the
// gradient volume is never read.  (A first version staged the slab in shared memory with 1-D
// bulk async copies and reached only 47 % of the measured HBM peak against 97 % for this
// direct-load form; see profiles/README.md.)
#include "common.cuh"

namespace az {

constexpr int kFwdThreads = 256;
constexpr int kFwdPPT = 1;  // float4 positions per thread
constexpr int kFwdDG = 8;   // disparity planes per CTA (multiple of 4)

// window [k, k+4) of the 8 floats (a, b), k in 1..4 static after unrolling
__device__ __forceinline__ float4 cwin8(const float4& a, const float4& b, int k) {
    switch (k) {
        case 1: return make_float4(a.y, a.z, a.w, b.x);
        case 2: return make_float4(a.z, a.w, b.x, b.y);
        case 3: return make_float4(a.w, b.x, b.y, b.z);
        default: return b;
    }
}

// ------------------------------------------------------------------------------------------
// forward, vectorised (W % 4 == 0, 16-byte aligned bases)
// grid = (ceil(H*W/4 / (256*PPT)), 2C * ceil(Dq/kFwdDG), B): a CTA owns 256 consecutive float4 (4 KB) of one
// (b, channel) feature plane and writes them to kFwdDG consecutive disparity planes.  Short-lived CTAs
// matter: a pure store stream with this pattern reaches 7.0 TB/s when a CTA sweeps all 48 planes and
// 7.5 TB/s (= memset) with 8 planes per CTA (benchmarks/micro/store_patterns.cu); the first version of this
// kernel swept all planes in 8 KB runs from four pre-shifted shared-memory copies of the rows (0.515 ms at
// B=8 against 0.488 ms for this form; 8 KB runs with 8 planes per CTA: 0.495 ms, 16 planes: 0.498 ms).
// Left half: the float4 stays in registers, only the x >= i mask changes.
// Right half: out[i][x..x+3] = R[x-i .. x-i+3] (0 left of the row).  The shift by i = 4m + r is the window
// [4-r, 8-r) of two ALIGNED quads, A at x-4m-4 and B at x-4m, read straight from the feature plane (4 MB per
// pair, L2 resident; a quad is wholly inside or wholly left of its row); B of step m+1 is A of step m, so a
// thread issues three 128-bit loads for its eight planes and selects the windows with static register indices.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFwdThreads) concat_fwd_vec4_kernel(const float* __restrict__ L,
                                                                     const float* __restrict__ R,
                                                                     float* __restrict__ vol, int C, int H,
                                                                     int W, int Dq, int ngroups) {
    const int W4 = W >> 2;
    const int P = kFwdThreads * kFwdPPT;
    const int oc = blockIdx.y / ngroups, dg = blockIdx.y - oc * ngroups, b = blockIdx.z;
    const int i0 = dg * kFwdDG, i1 = min(Dq, i0 + kFwdDG);
    const bool right = oc >= C;
    const int c = right ? oc - C : oc;
    const int p0 = blockIdx.x * P;
    const int pend = min(p0 + P, H * W4);
    const size_t HW = (size_t)H * W;
    const float* src = (right ? R : L) + ((size_t)b * C + c) * HW;
    float* out = vol + ((size_t)b * 2 * C + oc) * Dq * HW;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);

#pragma unroll
    for (int k = 0; k < kFwdPPT; ++k) {
        const int p = p0 + threadIdx.x + k * kFwdThreads;
        if (p >= pend) continue;
        const int x = (p % W4) * 4;
        const float* sp = src + (size_t)p * 4;  // feature[y][x]
        float* op = out + (size_t)p * 4;
        if (!right) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(sp));
            for (int i = i0; i < i1; ++i) {
                float4 o = v;
                if (x + 3 < i) { o = zero; }
                else if (x < i) {
                    if (x + 0 < i) o.x = 0.f;
                    if (x + 1 < i) o.y = 0.f;
                    if (x + 2 < i) o.z = 0.f;
                }
                st_stream(reinterpret_cast<float4*>(op + (size_t)i * HW), o);
            }
        } else {
            float4 Bv = (x - i0 >= 0) ? __ldg(reinterpret_cast<const float4*>(sp - i0)) : zero;
            for (int m4 = i0; m4 < i1; m4 += 4) {
                const float4 Av = (x - m4 - 4 >= 0) ? __ldg(reinterpret_cast<const float4*>(sp - m4 - 4)) : zero;
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (m4 + r < i1)
                        st_stream(reinterpret_cast<float4*>(op + (size_t)(m4 + r) * HW), cwin8(Av, Bv, 4 - r));
                Bv = Av;
            }
        }
    }
}


// ------------------------------------------------------------------------------------------
// forward with 256-bit accesses (W % 8 == 0, 32-byte aligned bases): sm_100's LDG.E.256 / STG.E.256 halve the number
// of store instructions and L1 requests of the store stream.  A thread owns 8 consecutive floats of one (b, channel)
// plane and writes them to the kFwdDG = 8 consecutive disparity planes i0 .. i0+7, i0 a multiple of 8, so the right
// half's shift by i = i0 + r is the window [8-r, 16-r) of two ALIGNED octets, A at x-i0-8 and B at x-i0 (an octet is
// wholly inside or wholly left of its row): two loads and eight stores per thread, static register indices.
// grid = (ceil(H*W/8 / THREADS), 2C * ceil(Dq/8), B); a CTA writes runs of THREADS*32 bytes.
// ------------------------------------------------------------------------------------------
template <int THREADS>
__global__ void __launch_bounds__(THREADS) concat_fwd_vec8_kernel(const float* __restrict__ L,
                                                                  const float* __restrict__ R,
                                                                  float* __restrict__ vol, int C, int H, int W,
                                                                  int Dq, int ngroups) {
    static_assert(kFwdDG == 8, "the octet window needs plane groups of 8");
    const int W8 = W >> 3;
    const int oc = blockIdx.y / ngroups, dg = blockIdx.y - oc * ngroups, b = blockIdx.z;
    const int i0 = dg * kFwdDG, i1 = min(Dq, i0 + kFwdDG);
    const bool right = oc >= C;
    const int c = right ? oc - C : oc;
    const int p = blockIdx.x * THREADS + threadIdx.x;
    if (p >= H * W8) return;
    const size_t HW = (size_t)H * W;
    const int x = (p % W8) * 8;
    const float* sp = (right ? R : L) + ((size_t)b * C + c) * HW + (size_t)p * 8;  // feature[y][x]
    float* op = vol + ((size_t)b * 2 * C + oc) * Dq * HW + (size_t)p * 8;
    if (!right) {
        const float8 v = ldg8(sp);
#pragma unroll
        for (int r = 0; r < kFwdDG; ++r) {
            const int i = i0 + r;
            if (i >= i1) break;
            float8 o = v;
            if (x < i) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (x + k < i) o.v[k] = 0.f;
            }
            st_stream8(op + (size_t)i * HW, o);
        }
    } else {
        float ab[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) ab[k] = 0.f;
        if (x - i0 >= 0) {
            const float8 Bv = ldg8(sp - i0);
#pragma unroll
            for (int k = 0; k < 8; ++k) ab[8 + k] = Bv.v[k];
        }
        if (x - i0 - 8 >= 0) {
            const float8 Av = ldg8(sp - i0 - 8);
#pragma unroll
            for (int k = 0; k < 8; ++k) ab[k] = Av.v[k];
        }
#pragma unroll
        for (int r = 0; r < kFwdDG; ++r) {
            if (i0 + r >= i1) break;
            float8 o;
#pragma unroll
            for (int k = 0; k < 8; ++k) o.v[k] = ab[8 - r + k];
            st_stream8(op + (size_t)(i0 + r) * HW, o);
        }
    }
}

// ------------------------------------------------------------------------------------------
// forward through the bulk-store engine (round 2, north-star "TMA bulk copies"): the output never
// crosses the register file.  The LSU only stages the input rows in shared memory (1/48 + 4/48 of the
// output bytes); every output byte is written by cp.async.bulk shared -> global (SASS UBLKCP).
//   right half: out[C+c][i][y][:] = R[y][: - i].  Four copies of each row, pre-shifted by r = 0..3 floats
//     behind a zero prefix of PAD >= Dq floats, make the source of plane i = 4m + r the 16-byte aligned
//     address copy_r[y] + PAD - 4m: one 4*W-byte bulk copy per (plane, row), no waits (the staged rows are
//     never modified).
//   left half:  out[c][i][y][:] = L[y][:] with the first i columns zeroed.  The rows are staged densely
//     (pitch W), so an 8-row chunk of a plane is ONE contiguous copy on both sides; the thread that owns a
//     chunk walks the planes in order and zeroes column i-1 in place before plane i
//     (cp.async.bulk.wait_group.read, 8 stores, fence.proxy.async, next copy).
// grid = (nL + nR, C, B): blockIdx.x < nL -> left CTA of kBulkRowsL rows, else right CTA of kBulkRowsR rows.
// ------------------------------------------------------------------------------------------
constexpr int kBulkThreads = 128, kBulkRowsL = 32, kBulkChunk = 8, kBulkRowsR = 8;

__global__ void __launch_bounds__(kBulkThreads) concat_fwd_bulk_kernel(const float* __restrict__ L,
                                                                       const float* __restrict__ R,
                                                                       float* __restrict__ vol, int C, int H, int W,
                                                                       int Dq, int nL, int PAD, int halves) {
    extern __shared__ __align__(128) float sm[];
    const int tid = threadIdx.x;
    const int c = blockIdx.y, b = blockIdx.z;
    const size_t HW = (size_t)H * W;
    const int W4 = W >> 2;
    if ((int)blockIdx.x < nL) {
        if (!(halves & 1)) return;
        const int y0 = blockIdx.x * kBulkRowsL, nrows = min(kBulkRowsL, H - y0);
        const float* src = L + ((size_t)b * C + c) * HW + (size_t)y0 * W;
        for (int t = tid; t < nrows * W4; t += kBulkThreads)
            reinterpret_cast<float4*>(sm)[t] = __ldg(reinterpret_cast<const float4*>(src) + t);
        fence_async_smem();
        __syncthreads();
        const int nchunks = (nrows + kBulkChunk - 1) / kBulkChunk;
        if (tid < nchunks) {
            const int r0 = tid * kBulkChunk, nr = min(kBulkChunk, nrows - r0);
            float* chunk = sm + (size_t)r0 * W;
            float* out = vol + ((size_t)b * 2 * C + c) * Dq * HW + (size_t)(y0 + r0) * W;
            const uint32_t bytes = (uint32_t)(nr * W) * 4u;
            for (int i = 0; i < Dq; ++i) {
                if (i > 0 && i <= W) {
                    bulk_wait_read<0>();  // plane i-1 has been read out of shared memory
                    for (int r = 0; r < nr; ++r) chunk[r * W + i - 1] = 0.f;
                    fence_async_smem();
                }
                bulk_s2g(out + (size_t)i * HW, chunk, bytes);
                bulk_commit();
            }
            bulk_wait_read<0>();
        }
    } else {
        if (!(halves & 2)) return;
        const int y0 = ((int)blockIdx.x - nL) * kBulkRowsR, nrows = min(kBulkRowsR, H - y0);
        const int PW = PAD + W;                       // row pitch of a staged copy (multiple of 4)
        const float* src = R + ((size_t)b * C + c) * HW + (size_t)y0 * W;
        // copy r, row y: [zeros(PAD + r) | R[y][0 .. W - r)]
        const int PW4 = PW >> 2, PAD4 = PAD >> 2;
        for (int t = tid; t < nrows * PW4; t += kBulkThreads) {
            const int y = t / PW4, q = t - y * PW4;   // quad q of the padded row
            float4 cur = make_float4(0.f, 0.f, 0.f, 0.f), prev = cur;
            if (q >= PAD4) cur = __ldg(reinterpret_cast<const float4*>(src + (size_t)y * W) + (q - PAD4));
            if (q > PAD4) prev = __ldg(reinterpret_cast<const float4*>(src + (size_t)y * W) + (q - PAD4 - 1));
            float* d = sm + (size_t)y * PW + 4 * q;
            const size_t cs = (size_t)kBulkRowsR * PW;  // floats per copy
            *reinterpret_cast<float4*>(d) = cur;
            *reinterpret_cast<float4*>(d + cs) = make_float4(prev.w, cur.x, cur.y, cur.z);
            *reinterpret_cast<float4*>(d + 2 * cs) = make_float4(prev.z, prev.w, cur.x, cur.y);
            *reinterpret_cast<float4*>(d + 3 * cs) = make_float4(prev.y, prev.z, prev.w, cur.x);
        }
        fence_async_smem();
        __syncthreads();
        float* out = vol + ((size_t)b * 2 * C + C + c) * Dq * HW + (size_t)y0 * W;
        const uint32_t bytes = (uint32_t)W * 4u;
        for (int idx = tid; idx < Dq * nrows; idx += kBulkThreads) {
            const int i = idx / nrows, y = idx - i * nrows;
            const float* s = sm + (size_t)(i & 3) * kBulkRowsR * PW + (size_t)y * PW + PAD - (i & ~3);
            bulk_s2g(out + (size_t)i * HW + (size_t)y * W, s, bytes);
        }
        bulk_commit();
        bulk_wait_read<0>();
    }
}

// ------------------------------------------------------------------------------------------
// forward, scalar fallback (any W, any alignment): one thread per output element of a
// (b, oc, i) plane.  grid = (ceil(H*W/256), Dq, B*2C)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) concat_fwd_scalar_kernel(const float* __restrict__ L,
                                                                const float* __restrict__ R,
                                                                float* __restrict__ vol, int C, int H, int W,
                                                                int Dq) {
    const int hw = blockIdx.x * 256 + threadIdx.x;
    if (hw >= H * W) return;
    const int i = blockIdx.y;
    const int boc = blockIdx.z, b = boc / (2 * C), oc = boc - b * 2 * C;
    const int x = hw % W;
    float v = 0.f;
    if (x >= i) {
        if (oc < C) v = __ldg(L + ((size_t)b * C + oc) * H * W + hw);
        else        v = __ldg(R + ((size_t)b * C + (oc - C)) * H * W + hw - i);
    }
    st_stream(vol + (((size_t)boc * Dq + i) * H) * W + hw, v);
}

// ------------------------------------------------------------------------------------------
// backward, direct path (W % 4 == 0, aligned): one thread per float4 of gL / gR, serial loop over
// the disparity planes with 8 independent 128-bit streaming loads in flight.  Coalesced along W;
// the zero triangle x < i of the volume gradient is never read.  grid = (ceil(H*W/4/256), 2C, B)
// Right half: output columns x..x+3 need g[i][x+i .. x+i+3]; with i = 4m+r that is the window
// [r, r+4) of the two ALIGNED float4 at columns x+4m and x+4m+4, selected with static register
// indices inside the unrolled r loop.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float sel7(const float4& a, const float4& b, int k) {
    switch (k) {
        case 0: return a.x; case 1: return a.y; case 2: return a.z; case 3: return a.w;
        case 4: return b.x; case 5: return b.y; default: return b.z;
    }
}

__global__ void __launch_bounds__(256) concat_bwd_direct_kernel(const float* __restrict__ gvol,
                                                                float* __restrict__ gL, float* __restrict__ gR,
                                                                int C, int H, int W, int Dq) {
    const int W4 = W >> 2;
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= H * W4) return;
    const int oc = blockIdx.y, b = blockIdx.z;
    const bool right = oc >= C;
    float* gout = right ? gR : gL;
    if (gout == nullptr) return;
    const int c = right ? oc - C : oc;
    const int y = p / W4, x = (p - y * W4) * 4;
    const size_t HW = (size_t)H * W;
    const float* g = gvol + ((size_t)b * 2 * C + oc) * Dq * HW + (size_t)p * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!right) {
        const int lim = min(Dq, x + 4);  // planes i > x+3 only hold the zero triangle for these columns
        for (int i0 = 0; i0 < lim; i0 += 8) {
            float4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                v[k] = (i0 + k < lim) ? ld_stream(reinterpret_cast<const float4*>(g + (size_t)(i0 + k) * HW)) : zero;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = i0 + k;
                acc.x += (x + 0 >= i) ? v[k].x : 0.f;
                acc.y += (x + 1 >= i) ? v[k].y : 0.f;
                acc.z += (x + 2 >= i) ? v[k].z : 0.f;
                acc.w += v[k].w;  // x+3 >= i holds for every loaded plane
            }
        }
    } else {
        const int lim = min(Dq, W - x);  // x+i < W
        for (int m4 = 0; m4 < lim; m4 += 4) {
            float4 A[4], Bv[4];
            const bool inB = (x + m4 + 4) < W;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const bool on = (m4 + r) < lim;
                const float* q = g + (size_t)(m4 + r) * HW + m4;
                A[r] = on ? __ldg(reinterpret_cast<const float4*>(q)) : zero;
                Bv[r] = (on && inB) ? __ldg(reinterpret_cast<const float4*>(q + 4)) : zero;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                acc.x += sel7(A[r], Bv[r], r + 0);
                acc.y += sel7(A[r], Bv[r], r + 1);
                acc.z += sel7(A[r], Bv[r], r + 2);
                acc.w += sel7(A[r], Bv[r], r + 3);
            }
        }
    }
    *reinterpret_cast<float4*>(gout + ((size_t)b * C + c) * HW + (size_t)p * 4) = acc;
}

// backward, scalar fallback: one thread per output element.  grid = (ceil(H*W/256), 2C, B)
__global__ void __launch_bounds__(256) concat_bwd_scalar_kernel(const float* __restrict__ gvol,
                                                                float* __restrict__ gL, float* __restrict__ gR,
                                                                int C, int H, int W, int Dq) {
    const int hw = blockIdx.x * 256 + threadIdx.x;
    if (hw >= H * W) return;
    const int oc = blockIdx.y, b = blockIdx.z;
    const bool right = oc >= C;
    float* gout = right ? gR : gL;
    if (gout == nullptr) return;
    const int c = right ? oc - C : oc;
    const int x = hw % W;
    const size_t HW = (size_t)H * W;
    const float* g = gvol + ((size_t)b * 2 * C + oc) * Dq * HW + hw;
    float acc = 0.f;
    if (right) {
        const int lim = min(Dq, W - x);
        for (int i = 0; i < lim; ++i) acc += ld_stream(g + (size_t)i * HW + i);
    } else {
        const int lim = min(Dq, x + 1);
        for (int i = 0; i < lim; ++i) acc += ld_stream(g + (size_t)i * HW);
    }
    gout[((size_t)b * C + c) * HW + hw] = acc;
}


// ------------------------------------------------------------------------------------------
// channels_last_3d variants ("emit the volume in the layout cuDNN wants").
// Logical tensor [B,2C,Dq,H,W] with torch.channels_last_3d strides, i.e. memory order
// [B][Dq][H][W][2C]: the 3-D aggregation that consumes the volume (plain cuDNN, not this library)
// runs 2-2.6x faster on B200 in this layout (benchmarks/agg_layout_probe.py), and converting an
// NCDHW volume afterwards costs a second pass over the 401 MB.
// Forward: a CTA owns 64 consecutive x of one (b, y): it stages the 64 left and 64 + Dq - 1 right
// feature vectors TRANSPOSED in shared memory (position-major, C floats per position), then for
// every disparity writes 64 positions x 2C floats = one contiguous 16 KB run with 128-bit streaming
// stores (LDS.128 -> STG.128, no arithmetic).  grid = (ceil(W/64), H, B).
// Backward: a 32-position tile; a thread owns (position, four channels), walks the Dq planes with 8
// independent 128-bit loads in flight (left half: straight down; right half: along the x + i
// diagonal), fixed order, atomic-free; the sums are transposed through shared memory so that gL / gR
// are written as coalesced rows.
// ------------------------------------------------------------------------------------------
constexpr int kClTX = 64, kClTXB = 32, kClDG = 8, kClThreads = 256;  // positions per CTA (forward / backward), planes per forward CTA

__global__ void __launch_bounds__(kClThreads) concat_fwd_ndhwc_kernel(const float* __restrict__ L,
                                                                    const float* __restrict__ R,
                                                                    float* __restrict__ vol, int C, int H, int W,
                                                                    int Dq, int ntiles) {
    extern __shared__ __align__(16) float smem[];
    const int P = C + 4;  // pitch of one position (floats): multiple of 4, not of 32
    float* LsT = smem;                // [kClTX][P]
    float* RsT = smem + kClTX * P;    // [kClTX + kClDG - 1][P]: x0 - (i1-1) .. x0 + kClTX - 1
    const int tile = blockIdx.x % ntiles, dg = blockIdx.x / ntiles;  // short-lived CTAs: kClDG planes each
    const int i0 = dg * kClDG, i1 = min(Dq, i0 + kClDG);
    const int x0 = tile * kClTX, y = blockIdx.y, b = blockIdx.z;
    const int nr = kClTX + (i1 - i0) - 1;
    const size_t HW = (size_t)H * W;
    const float* Lb = L + (size_t)b * C * HW + (size_t)y * W;
    const float* Rb = R + (size_t)b * C * HW + (size_t)y * W;
    for (int t = threadIdx.x; t < C * kClTX; t += kClThreads) {
        const int c = t / kClTX, xl = t - c * kClTX, x = x0 + xl;
        LsT[xl * P + c] = x < W ? __ldg(Lb + (size_t)c * HW + x) : 0.f;
    }
    for (int t = threadIdx.x; t < C * nr; t += kClThreads) {
        const int c = t / nr, r = t - c * nr, x = x0 - (i1 - 1) + r;
        RsT[r * P + c] = (x >= 0 && x < W) ? __ldg(Rb + (size_t)c * HW + x) : 0.f;
    }
    __syncthreads();
    const int C4 = C >> 2, Q = 2 * C4;  // float4 per position
    const int npos = min(kClTX, W - x0);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    float* vrow = vol + (((size_t)b * Dq * H + y) * W + x0) * (size_t)(2 * C);  // d = 0
    const size_t dstep = (size_t)H * W * (2 * C);
    if (Q <= kClThreads && (kClThreads % Q) == 0 && kClTX * Q <= 4 * kClThreads &&
        ((kClThreads / Q) >= kClTX || (kClTX % (kClThreads / Q)) == 0)) {
        // fast path (C = 4, 8, 16, 32, 64): a thread keeps its channel quad q and walks <= 4 positions,
        // no integer division in the sweep; the left-half value does not depend on d and stays in registers
        const int q = threadIdx.x % Q, xstep = kClThreads / Q, nk = xstep >= kClTX ? 1 : kClTX / xstep;
        const bool left = q < C4;
        float4 vl[4];
        const float* rp[4];
        int xs[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int xl = threadIdx.x / Q + k * xstep;
            xs[k] = (k < nk && xl < npos) ? x0 + xl : -1;  // -1: nothing to write
            vl[k] = zero;
            rp[k] = RsT;
            if (xs[k] >= 0) {
                if (left) vl[k] = *reinterpret_cast<const float4*>(LsT + xl * P + 4 * q);
                else rp[k] = RsT + (xl + i1 - 1) * P + 4 * (q - C4);
            }
        }
        for (int d = i0; d < i1; ++d) {
            float4* od = reinterpret_cast<float4*>(vrow + (size_t)d * dstep) + threadIdx.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (xs[k] >= 0) {
                    float4 v = zero;
                    if (xs[k] >= d) v = left ? vl[k] : *reinterpret_cast<const float4*>(rp[k] - d * P);
                    st_stream(od + k * kClThreads, v);
                }
            }
        }
        return;
    }
    for (int d = i0; d < i1; ++d) {
        float4* od = reinterpret_cast<float4*>(vrow + (size_t)d * dstep);
        for (int t = threadIdx.x; t < npos * Q; t += kClThreads) {
            const int xl = t / Q, q = t - xl * Q;
            float4 v = zero;
            if (x0 + xl >= d)
                v = q < C4 ? *reinterpret_cast<const float4*>(LsT + xl * P + 4 * q)
                           : *reinterpret_cast<const float4*>(RsT + (xl - d + i1 - 1) * P + 4 * (q - C4));
            st_stream(od + t, v);
        }
    }
}

__global__ void __launch_bounds__(kClThreads) concat_bwd_ndhwc_kernel(const float* __restrict__ gvol,
                                                                    float* __restrict__ gL, float* __restrict__ gR,
                                                                    int C, int H, int W, int Dq) {
    extern __shared__ __align__(16) float smem[];  // Ts[2C][kClTXB + 1]
    const int x0 = blockIdx.x * kClTXB, y = blockIdx.y, b = blockIdx.z;
    const int C4 = C >> 2, Q = 2 * C4, C2 = 2 * C;
    const size_t plane = (size_t)H * W * C2;  // floats between consecutive disparities
    const float* gb = gvol + ((size_t)b * Dq * H + y) * (size_t)W * C2;
    for (int t = threadIdx.x; t < kClTXB * Q; t += kClThreads) {
        const int xl = t / Q, q = t - xl * Q, x = x0 + xl;
        const bool right = q >= C4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x < W && (right ? gR != nullptr : gL != nullptr)) {
            // left: g[d][x] for d <= x;  right: g[d][x + d] while x + d < W
            const int nd = right ? min(Dq, W - x) : min(Dq, x + 1);
            const float* p = gb + (size_t)x * C2 + 4 * q;
            const size_t step = plane + (right ? (size_t)C2 : 0);
            int d = 0;
            for (; d + 8 <= nd; d += 8) {
                float4 v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = ld_stream(reinterpret_cast<const float4*>(p + (size_t)(d + k) * step));
#pragma unroll
                for (int k = 0; k < 8; ++k) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
            }
            for (; d < nd; ++d) {
                const float4 v = ld_stream(reinterpret_cast<const float4*>(p + (size_t)d * step));
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        float* ts = smem + (4 * q) * (kClTXB + 1) + xl;
        ts[0] = acc.x;
        ts[kClTXB + 1] = acc.y;
        ts[2 * (kClTXB + 1)] = acc.z;
        ts[3 * (kClTXB + 1)] = acc.w;
    }
    __syncthreads();
    const size_t HW = (size_t)H * W;
    for (int t = threadIdx.x; t < C2 * kClTXB; t += kClThreads) {
        const int ch = t / kClTXB, xl = t - ch * kClTXB, x = x0 + xl;
        if (x >= W) continue;
        float* dst = ch < C ? gL : gR;
        if (dst == nullptr) continue;
        const int c = ch < C ? ch : ch - C;
        dst[((size_t)b * C + c) * HW + (size_t)y * W + x] = smem[ch * (kClTXB + 1) + xl];
    }
}

}  // namespace az

using namespace az;

extern "C" int az_concat_volume_fwd(const float* L, const float* R, float* vol, int64_t B, int64_t C, int64_t H,
                                    int64_t W, int64_t Dq, void* stream) {
    if (!L || !R || !vol || B <= 0 || C <= 0 || H <= 0 || W <= 0 || Dq <= 0) return AZ_ERR_BAD_ARG;
    if (H * W >= (1ll << 31) / 4 || B * 2 * C > 65535 * 32) return AZ_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (W % 4 == 0) && aligned16(L) && aligned16(R) && aligned16(vol) && B <= 65535 && 2 * C <= 65535;
    if (vec) {
        const int W4 = (int)(W / 4);
        const int P = kFwdThreads * kFwdPPT;
        const int64_t ngroups = ceil_div(Dq, kFwdDG);
        // 0 = register/LSU store path; 1 = both halves through the bulk-store engine; 2 = right half bulk, left half LSU
        const int impl = tuning("AZ_CONCAT_FWD", 0);
        const int PAD = (int)((Dq + 3) / 4 * 4);
        const size_t smemL = (size_t)kBulkRowsL * W * 4, smemR = (size_t)4 * kBulkRowsR * (PAD + W) * 4;
        const size_t smemB = smemL > smemR ? smemL : smemR;
        if ((impl == 1 || impl == 2) && smemB <= 200 * 1024 && C <= 65535 && Dq <= W) {
            const int nL = (int)ceil_div(H, kBulkRowsL), nR = (int)ceil_div(H, kBulkRowsR);
            cudaError_t e = cudaFuncSetAttribute(concat_fwd_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) return (int)e;
            dim3 gb((unsigned)(nL + nR), (unsigned)C, (unsigned)B);
            concat_fwd_bulk_kernel<<<gb, kBulkThreads, smemB, st>>>(L, R, vol, (int)C, (int)H, (int)W, (int)Dq, nL, PAD,
                                                                    impl == 1 ? 3 : 2);
            AZ_LAUNCH_CHECK();
            if (impl == 2) {  // left half: channels [0, C) of the LSU kernel
                dim3 grid((unsigned)ceil_div(H * W4, P), (unsigned)(C * ngroups), (unsigned)B);
                concat_fwd_vec4_kernel<<<grid, kFwdThreads, 0, st>>>(L, R, vol, (int)C, (int)H, (int)W, (int)Dq, (int)ngroups);
                AZ_LAUNCH_CHECK();
            }
            return 0;
        }
        // 3 / 4 / 5 = 256-bit loads and stores, CTAs of 128 / 256 / 64 threads (4 / 8 / 2 KB runs)
        if (impl >= 3 && impl <= 5 && W % 8 == 0 && aligned32(L) && aligned32(R) && aligned32(vol) && 2 * C * ngroups <= 65535) {
            const int64_t n8 = H * (W / 8);
            const int nt = impl == 3 ? 128 : impl == 4 ? 256 : 64;
            dim3 grid((unsigned)ceil_div(n8, nt), (unsigned)(2 * C * ngroups), (unsigned)B);
            if (nt == 128) concat_fwd_vec8_kernel<128><<<grid, 128, 0, st>>>(L, R, vol, (int)C, (int)H, (int)W, (int)Dq, (int)ngroups);
            else if (nt == 256) concat_fwd_vec8_kernel<256><<<grid, 256, 0, st>>>(L, R, vol, (int)C, (int)H, (int)W, (int)Dq, (int)ngroups);
            else concat_fwd_vec8_kernel<64><<<grid, 64, 0, st>>>(L, R, vol, (int)C, (int)H, (int)W, (int)Dq, (int)ngroups);
            AZ_LAUNCH_CHECK();
            return 0;
        }
        if (2 * C * ngroups <= 65535) {
            dim3 grid((unsigned)ceil_div(H * W4, P), (unsigned)(2 * C * ngroups), (unsigned)B);
            concat_fwd_vec4_kernel<<<grid, kFwdThreads, 0, st>>>(L, R, vol, (int)C, (int)H, (int)W, (int)Dq, (int)ngroups);
            AZ_LAUNCH_CHECK();
            return 0;
        }
    }
    dim3 grid((unsigned)ceil_div(H * W, 256), (unsigned)Dq, (unsigned)(B * 2 * C));
    concat_fwd_scalar_kernel<<<grid, 256, 0, st>>>(L, R, vol, (int)C, (int)H, (int)W, (int)Dq);
    AZ_LAUNCH_CHECK();
    return 0;
}

extern "C" int az_concat_volume_bwd(const float* gvol, float* gL, float* gR, int64_t B, int64_t C, int64_t H,
                                    int64_t W, int64_t Dq, void* stream) {
    if (!gvol || B <= 0 || C <= 0 || H <= 0 || W <= 0 || Dq <= 0) return AZ_ERR_BAD_ARG;
    if (H * W >= (1ll << 31) / 4 || B > 65535 || 2 * C > 65535) return AZ_ERR_BAD_ARG;
    if (!gL && !gR) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec4 = (W % 4 == 0) && aligned16(gvol) && (!gL || aligned16(gL)) && (!gR || aligned16(gR));
    if (vec4) {
        dim3 grid((unsigned)ceil_div(H * (W / 4), 256), (unsigned)(2 * C), (unsigned)B);
        concat_bwd_direct_kernel<<<grid, 256, 0, st>>>(gvol, gL, gR, (int)C, (int)H, (int)W, (int)Dq);
        AZ_LAUNCH_CHECK();
        return 0;
    }
    dim3 grid((unsigned)ceil_div(H * W, 256), (unsigned)(2 * C), (unsigned)B);
    concat_bwd_scalar_kernel<<<grid, 256, 0, st>>>(gvol, gL, gR, (int)C, (int)H, (int)W, (int)Dq);
    AZ_LAUNCH_CHECK();
    return 0;
}

extern "C" int az_concat_volume_fwd_ndhwc(const float* L, const float* R, float* vol, int64_t B, int64_t C, int64_t H,
                                          int64_t W, int64_t Dq, void* stream) {
    if (!L || !R || !vol || B <= 0 || C <= 0 || H <= 0 || W <= 0 || Dq <= 0) return AZ_ERR_BAD_ARG;
    if ((C % 4) != 0 || !aligned16(vol) || B > 65535 || H > 65535 || H * W >= (1ll << 31) / 4) return AZ_ERR_BAD_ARG;
    const size_t smem = (size_t)(2 * kClTX + kClDG - 1) * (size_t)(C + 4) * sizeof(float);
    if (smem > 200 * 1024) return AZ_ERR_BAD_ARG;
    const int64_t ntiles = ceil_div(W, kClTX), ngroups = ceil_div(Dq, kClDG);
    if (ntiles * ngroups >= (1ll << 31)) return AZ_ERR_BAD_ARG;
    cudaError_t e = cudaFuncSetAttribute(concat_fwd_ndhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)(ntiles * ngroups), (unsigned)H, (unsigned)B);
    concat_fwd_ndhwc_kernel<<<grid, kClThreads, smem, (cudaStream_t)stream>>>(L, R, vol, (int)C, (int)H, (int)W, (int)Dq,
                                                                              (int)ntiles);
    AZ_LAUNCH_CHECK();
    return 0;
}

extern "C" int az_concat_volume_bwd_ndhwc(const float* gvol, float* gL, float* gR, int64_t B, int64_t C, int64_t H,
                                          int64_t W, int64_t Dq, void* stream) {
    if (!gvol || B <= 0 || C <= 0 || H <= 0 || W <= 0 || Dq <= 0) return AZ_ERR_BAD_ARG;
    if ((C % 4) != 0 || !aligned16(gvol) || B > 65535 || H > 65535 || H * W >= (1ll << 31) / 4) return AZ_ERR_BAD_ARG;
    if (!gL && !gR) return 0;
    const size_t smem = (size_t)(2 * C) * (kClTXB + 1) * sizeof(float);
    if (smem > 200 * 1024) return AZ_ERR_BAD_ARG;
    cudaError_t e = cudaFuncSetAttribute(concat_bwd_ndhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)ceil_div(W, kClTXB), (unsigned)H, (unsigned)B);
    concat_bwd_ndhwc_kernel<<<grid, kClThreads, smem, (cudaStream_t)stream>>>(gvol, gL, gR, (int)C, (int)H, (int)W, (int)Dq);
    AZ_LAUNCH_CHECK();
    return 0;
}
