// Group-wise correlation volume, forward and backward.
// NOT IN THE REFERENCE: the definition is GwcNet's,
//   vol[b,g,i,y,x] = (1/cpg) * sum_{c in group g} L[b,c,y,x] * R[b,c,y,x-i]   (x >= i, else 0)
// with cpg = C/G.  Parity is against this repository's own torch restatement (parity unpinned).
//
// Same data movement as the concat volume: the forward stages the cpg left/right rows of one
// (b, g, row-run) once (right rows as four pre-shifted copies behind a zero prefix => aligned
// 128-bit shared loads for every disparity), then streams Dq planes of 128-bit stores; the
// backward pulls the gradient slab into shared memory with bulk async copies and reduces it
// gather-style (no atomics) into gL and gR.
#include "common.cuh"

namespace az {

constexpr int kGwcThreads = 256;

// grid = (ceil(H*W/4 / 256), G, B); one float4 position per thread.
template <int CPG>
__global__ void __launch_bounds__(kGwcThreads) gwc_fwd_vec4_kernel(const float* __restrict__ L,
                                                                  const float* __restrict__ R,
                                                                  float* __restrict__ vol, int C, int H, int W,
                                                                  int Dq, int pad, int rows_cap) {
    extern __shared__ __align__(16) float smem[];
    const int W4 = W >> 2;
    const int g = blockIdx.y, b = blockIdx.z, G = gridDim.y;
    const int p0 = blockIdx.x * kGwcThreads;
    const int pend = min(p0 + kGwcThreads, H * W4);
    const size_t HW = (size_t)H * W;
    const int y_first = p0 / W4, y_last = (pend - 1) / W4, nrows = y_last - y_first + 1;
    const int S = pad + W;
    const int copy_stride = rows_cap * S;      // floats per shifted copy of one channel
    const int ch_stride = 4 * copy_stride;     // floats per channel
    const float* Lg = L + ((size_t)b * C + (size_t)g * CPG) * HW;
    const float* Rg = R + ((size_t)b * C + (size_t)g * CPG) * HW;

    for (int t = threadIdx.x; t < CPG * copy_stride; t += kGwcThreads)
        reinterpret_cast<float4*>(smem)[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    for (int t = threadIdx.x; t < CPG * nrows * W; t += kGwcThreads) {
        const int c = t / (nrows * W), rem = t - c * nrows * W;
        const int y = rem / W, xx = rem - y * W;
        const float val = __ldg(Rg + (size_t)c * HW + (size_t)(y_first + y) * W + xx);
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (xx + r < W) smem[c * ch_stride + r * copy_stride + y * S + pad + xx + r] = val;
    }
    __syncthreads();

    const int pp = p0 + threadIdx.x;
    if (pp >= pend) return;
    const int y = pp / W4, x = (pp - y * W4) * 4;
    float4 l[CPG];
#pragma unroll
    for (int c = 0; c < CPG; ++c) l[c] = __ldg(reinterpret_cast<const float4*>(Lg + (size_t)c * HW + (size_t)pp * 4));
    const float* sp = smem + (y - y_first) * S + pad + x;
    float* out = vol + ((size_t)b * G + g) * Dq * HW + (size_t)pp * 4;
    const float inv = 1.0f / (float)CPG;
    for (int i = 0; i < Dq; ++i) {
        const float* q = sp + (i & 3) * copy_stride - (i & ~3);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int c = 0; c < CPG; ++c) {
            const float4 r = *reinterpret_cast<const float4*>(q + c * ch_stride);
            acc.x = fmaf(l[c].x, r.x, acc.x);
            acc.y = fmaf(l[c].y, r.y, acc.y);
            acc.z = fmaf(l[c].z, r.z, acc.z);
            acc.w = fmaf(l[c].w, r.w, acc.w);
        }
        acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
        st_stream(reinterpret_cast<float4*>(out + (size_t)i * HW), acc);
    }
}

// scalar fallback, any shape.  grid = (ceil(H*W/256), Dq, B*G)
__global__ void __launch_bounds__(256) gwc_fwd_scalar_kernel(const float* __restrict__ L, const float* __restrict__ R,
                                                             float* __restrict__ vol, int C, int G, int H, int W,
                                                             int Dq) {
    const int hw = blockIdx.x * 256 + threadIdx.x;
    if (hw >= H * W) return;
    const int i = blockIdx.y, bg = blockIdx.z, b = bg / G, g = bg - b * G;
    const int cpg = C / G, x = hw % W;
    const size_t HW = (size_t)H * W;
    float acc = 0.f;
    if (x >= i) {
        const float* l = L + ((size_t)b * C + (size_t)g * cpg) * HW + hw;
        const float* r = R + ((size_t)b * C + (size_t)g * cpg) * HW + hw - i;
        for (int c = 0; c < cpg; ++c) acc = fmaf(__ldg(l + (size_t)c * HW), __ldg(r + (size_t)c * HW), acc);
        acc *= 1.0f / (float)cpg;
    }
    st_stream(vol + (((size_t)bg * Dq + i) * H) * W + hw, acc);
}

// ------------------------------------------------------------------------------------------
// backward, bulk-async path: grid = (H, G, B), one image row per CTA.
// smem: gs[Dq][W] gradient rows | Ls[cpg][W] | Rs[cpg][pad+W] (zero prefix) | mbarriers
// gL[c,x] = inv * sum_{i<=x} g[i,x] * R[c,x-i];   gR[c,x] = inv * sum_{x+i<W} g[i,x+i] * L[c,x+i]
// ------------------------------------------------------------------------------------------
constexpr int kGwcBwdGroup = 8;

template <int CPG>
__global__ void __launch_bounds__(kGwcThreads) gwc_bwd_bulk_kernel(const float* __restrict__ gvol,
                                                                  const float* __restrict__ L,
                                                                  const float* __restrict__ R, float* __restrict__ gL,
                                                                  float* __restrict__ gR, int C, int H, int W, int Dq,
                                                                  int pad) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int y = blockIdx.x, g = blockIdx.y, b = blockIdx.z, G = gridDim.y;
    const size_t HW = (size_t)H * W;
    float* gs = reinterpret_cast<float*>(smem_raw);
    float* Ls = gs + (size_t)Dq * W;
    float* Rs = Ls + (size_t)CPG * W;
    const int S = pad + W;
    const int ngroups = (Dq + kGwcBwdGroup - 1) / kGwcBwdGroup;
    uint64_t* bars = reinterpret_cast<uint64_t*>(Rs + (size_t)CPG * S);
    const float* gsrc = gvol + ((size_t)b * G + g) * Dq * HW + (size_t)y * W;
    const size_t fbase = ((size_t)b * C + (size_t)g * CPG) * HW + (size_t)y * W;

    if (threadIdx.x == 0) {
        for (int q = 0; q < ngroups; ++q) mbar_init(&bars[q], 1);
        fence_mbar_init();
        for (int q = 0; q < ngroups; ++q) {
            const int i0 = q * kGwcBwdGroup, cnt = min(kGwcBwdGroup, Dq - i0);
            mbar_arrive_expect_tx(&bars[q], (uint32_t)(cnt * W * sizeof(float)));
            for (int i = i0; i < i0 + cnt; ++i)
                bulk_g2s(gs + (size_t)i * W, gsrc + (size_t)i * HW, (uint32_t)(W * sizeof(float)), &bars[q]);
        }
    }
    for (int t = threadIdx.x; t < CPG * S; t += kGwcThreads) {
        const int c = t / S, xx = t - c * S - pad;
        Rs[t] = xx >= 0 ? __ldg(R + fbase + (size_t)c * HW + xx) : 0.f;
    }
    for (int t = threadIdx.x; t < CPG * W; t += kGwcThreads) {
        const int c = t / W, xx = t - c * W;
        Ls[t] = __ldg(L + fbase + (size_t)c * HW + xx);
    }
    __syncthreads();  // Ls/Rs visible; mbarrier inits visible to the waiters

    const float inv = 1.0f / (float)CPG;
    for (int x0 = 0; x0 < W; x0 += kGwcThreads) {
        const int x = x0 + threadIdx.x;
        float aL[CPG], aR[CPG];
#pragma unroll
        for (int c = 0; c < CPG; ++c) { aL[c] = 0.f; aR[c] = 0.f; }
        for (int q = 0; q < ngroups; ++q) {
            mbar_wait(&bars[q], 0);
            if (x >= W) continue;
            const int i0 = q * kGwcBwdGroup, cnt = min(kGwcBwdGroup, Dq - i0);
            for (int i = i0; i < i0 + cnt; ++i) {
                if (i <= x) {
                    const float gv = gs[(size_t)i * W + x];
#pragma unroll
                    for (int c = 0; c < CPG; ++c) aL[c] = fmaf(gv, Rs[c * S + pad + x - i], aL[c]);
                }
                if (x + i < W) {
                    const float gv = gs[(size_t)i * W + x + i];
#pragma unroll
                    for (int c = 0; c < CPG; ++c) aR[c] = fmaf(gv, Ls[c * W + x + i], aR[c]);
                }
            }
        }
        if (x < W) {
#pragma unroll
            for (int c = 0; c < CPG; ++c) {
                if (gL != nullptr) gL[fbase + (size_t)c * HW + x] = aL[c] * inv;
                if (gR != nullptr) gR[fbase + (size_t)c * HW + x] = aR[c] * inv;
            }
        }
    }
}

// scalar fallback: one thread per (b, c, y, x).  grid = (ceil(H*W/256), C, B)
__global__ void __launch_bounds__(256) gwc_bwd_scalar_kernel(const float* __restrict__ gvol,
                                                             const float* __restrict__ L, const float* __restrict__ R,
                                                             float* __restrict__ gL, float* __restrict__ gR, int C,
                                                             int G, int H, int W, int Dq) {
    const int hw = blockIdx.x * 256 + threadIdx.x;
    if (hw >= H * W) return;
    const int c = blockIdx.y, b = blockIdx.z;
    const int cpg = C / G, g = c / cpg, x = hw % W;
    const size_t HW = (size_t)H * W;
    const float* gv = gvol + ((size_t)b * G + g) * Dq * HW + hw;
    const float* l = L + ((size_t)b * C + c) * HW + hw;
    const float* r = R + ((size_t)b * C + c) * HW + hw;
    float aL = 0.f, aR = 0.f;
    for (int i = 0; i < Dq; ++i) {
        if (i <= x) aL = fmaf(ld_stream(gv + (size_t)i * HW), __ldg(r - i), aL);
        if (x + i < W) aR = fmaf(ld_stream(gv + (size_t)i * HW + i), __ldg(l + i), aR);
    }
    const float inv = 1.0f / (float)cpg;
    if (gL != nullptr) gL[((size_t)b * C + c) * HW + hw] = aL * inv;
    if (gR != nullptr) gR[((size_t)b * C + c) * HW + hw] = aR * inv;
}

template <int CPG>
static int launch_gwc_fwd(const float* L, const float* R, float* vol, int B, int C, int G, int H, int W, int Dq,
                          cudaStream_t st, bool* done) {
    const int W4 = W / 4, pad = (Dq + 3) / 4 * 4;
    const int rows_cap = (kGwcThreads + W4 - 1) / W4 + 1;
    const size_t smem = (size_t)CPG * 4 * rows_cap * (pad + W) * sizeof(float);
    *done = false;
    if (smem > 200 * 1024) return 0;
    cudaError_t e = cudaFuncSetAttribute(gwc_fwd_vec4_kernel<CPG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)ceil_div((int64_t)H * W4, kGwcThreads), (unsigned)G, (unsigned)B);
    gwc_fwd_vec4_kernel<CPG><<<grid, kGwcThreads, smem, st>>>(L, R, vol, C, H, W, Dq, pad, rows_cap);
    *done = true;
    return (int)cudaGetLastError();
}

template <int CPG>
static int launch_gwc_bwd(const float* gvol, const float* L, const float* R, float* gL, float* gR, int B, int C, int G,
                          int H, int W, int Dq, cudaStream_t st, bool* done) {
    const int pad = (Dq + 3) / 4 * 4;
    const int ngroups = (Dq + kGwcBwdGroup - 1) / kGwcBwdGroup;
    const size_t smem = ((size_t)Dq * W + (size_t)CPG * W + (size_t)CPG * (pad + W)) * sizeof(float) +
                        (size_t)ngroups * sizeof(uint64_t);
    *done = false;
    if (smem > 200 * 1024) return 0;
    cudaError_t e = cudaFuncSetAttribute(gwc_bwd_bulk_kernel<CPG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)H, (unsigned)G, (unsigned)B);
    gwc_bwd_bulk_kernel<CPG><<<grid, kGwcThreads, smem, st>>>(gvol, L, R, gL, gR, C, H, W, Dq, pad);
    *done = true;
    return (int)cudaGetLastError();
}

}  // namespace az

using namespace az;

extern "C" int az_gwc_volume_fwd(const float* L, const float* R, float* vol, int64_t B, int64_t C, int64_t H,
                                 int64_t W, int64_t Dq, int64_t G, void* stream) {
    if (!L || !R || !vol || B <= 0 || C <= 0 || H <= 0 || W <= 0 || Dq <= 0 || G <= 0 || C % G != 0)
        return AZ_ERR_BAD_ARG;
    if (H * W >= (1ll << 31) / 4 || B > 65535 || G > 65535) return AZ_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t cpg = C / G;
    if ((W % 4 == 0) && aligned16(L) && aligned16(R) && aligned16(vol)) {
        bool done = false;
        int rc = 0;
        switch (cpg) {
            case 1: rc = launch_gwc_fwd<1>(L, R, vol, (int)B, (int)C, (int)G, (int)H, (int)W, (int)Dq, st, &done); break;
            case 2: rc = launch_gwc_fwd<2>(L, R, vol, (int)B, (int)C, (int)G, (int)H, (int)W, (int)Dq, st, &done); break;
            case 4: rc = launch_gwc_fwd<4>(L, R, vol, (int)B, (int)C, (int)G, (int)H, (int)W, (int)Dq, st, &done); break;
            case 8: rc = launch_gwc_fwd<8>(L, R, vol, (int)B, (int)C, (int)G, (int)H, (int)W, (int)Dq, st, &done); break;
            default: break;
        }
        if (rc != 0) return rc;
        if (done) return 0;
    }
    if (Dq > 65535 || B * G > 65535) return AZ_ERR_BAD_ARG;
    dim3 grid((unsigned)ceil_div(H * W, 256), (unsigned)Dq, (unsigned)(B * G));
    gwc_fwd_scalar_kernel<<<grid, 256, 0, st>>>(L, R, vol, (int)C, (int)G, (int)H, (int)W, (int)Dq);
    AZ_LAUNCH_CHECK();
    return 0;
}

extern "C" int az_gwc_volume_bwd(const float* gvol, const float* L, const float* R, float* gL, float* gR, int64_t B,
                                 int64_t C, int64_t H, int64_t W, int64_t Dq, int64_t G, void* stream) {
    if (!gvol || !L || !R || B <= 0 || C <= 0 || H <= 0 || W <= 0 || Dq <= 0 || G <= 0 || C % G != 0)
        return AZ_ERR_BAD_ARG;
    if (H * W >= (1ll << 31) / 4 || B > 65535 || C > 65535 || H > 65535) return AZ_ERR_BAD_ARG;
    if (!gL && !gR) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t cpg = C / G;
    if ((W % 4 == 0) && aligned16(gvol)) {
        bool done = false;
        int rc = 0;
        switch (cpg) {
            case 1: rc = launch_gwc_bwd<1>(gvol, L, R, gL, gR, (int)B, (int)C, (int)G, (int)H, (int)W, (int)Dq, st, &done); break;
            case 2: rc = launch_gwc_bwd<2>(gvol, L, R, gL, gR, (int)B, (int)C, (int)G, (int)H, (int)W, (int)Dq, st, &done); break;
            case 4: rc = launch_gwc_bwd<4>(gvol, L, R, gL, gR, (int)B, (int)C, (int)G, (int)H, (int)W, (int)Dq, st, &done); break;
            case 8: rc = launch_gwc_bwd<8>(gvol, L, R, gL, gR, (int)B, (int)C, (int)G, (int)H, (int)W, (int)Dq, st, &done); break;
            default: break;
        }
        if (rc != 0) return rc;
        if (done) return 0;
    }
    dim3 grid((unsigned)ceil_div(H * W, 256), (unsigned)C, (unsigned)B);
    gwc_bwd_scalar_kernel<<<grid, 256, 0, st>>>(gvol, L, R, gL, gR, (int)C, (int)G, (int)H, (int)W, (int)Dq);
    AZ_LAUNCH_CHECK();
    return 0;
}
