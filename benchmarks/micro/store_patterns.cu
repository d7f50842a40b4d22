// Store-bandwidth microbenchmark (not part of the library): how close can a pure store stream get to
// torch's fill_ (7.5 TB/s measured on this B200) with the access patterns the concat-volume kernel could use?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_patterns store_patterns.cu && ./store_patterns
#include <cstdio>
#include <cuda_runtime.h>

constexpr long long PLANE4 = 136LL * 240 / 4;      // float4 per (c,d) plane  (8160)
constexpr int DQ = 48, NOC = 64, NB = 8;
constexpr long long TOTAL4 = PLANE4 * DQ * NOC * NB;  // 3.2 GB / 16

template <bool CS> __device__ __forceinline__ void st(float4* p, float4 v) { if (CS) __stcs(p, v); else *p = v; }

// (1) linear: thread t writes float4 t, t + stride, ...
template <bool CS> __global__ void k_linear(float4* out, long long n, float4 v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) st<CS>(out + i, v);
}
// (2) the concat pattern: CTA (chunk, oc*b) writes 512 float4 (8 KB) at each of the 48 planes (stride 130 KB)
template <bool CS, int PPT> __global__ void __launch_bounds__(256) k_plane_sweep(float4* out, float4 v) {
    const long long base = (long long)blockIdx.y * DQ * PLANE4;
    const int p0 = blockIdx.x * 256 * PPT;
    for (int d = 0; d < DQ; ++d)
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const int p = p0 + k * 256 + threadIdx.x;
            if (p < PLANE4) st<CS>(out + base + (long long)d * PLANE4 + p, v);
        }
}
// (3) same amount of work per CTA (48 x 8 KB) but one contiguous 393 KB region per CTA
template <bool CS> __global__ void __launch_bounds__(256) k_cta_contig(float4* out, long long n, float4 v) {
    const long long base = (long long)blockIdx.x * 48 * 512;
    for (int j = 0; j < 48 * 2; ++j) {
        const long long i = base + (long long)j * 256 + threadIdx.x;
        if (i < n) st<CS>(out + i, v);
    }
}
// (4) plane sweep, d-major grid: blockIdx.x = chunk + 16 * d-group ... CTA writes only DG planes
template <bool CS, int DG> __global__ void __launch_bounds__(256) k_plane_dgroup(float4* out, float4 v) {
    const long long base = (long long)blockIdx.z * DQ * PLANE4 + (long long)blockIdx.y * DG * PLANE4;
    const int p0 = blockIdx.x * 512;
    for (int d = 0; d < DG; ++d)
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int p = p0 + k * 256 + threadIdx.x;
            if (p < PLANE4) st<CS>(out + base + (long long)d * PLANE4 + p, v);
        }
}

// (5) as (4) with the value LOADED first from a small L2-resident plane (what the concat kernel does): how much
// does the load -> store dependency of a short-lived CTA cost?
template <int DG> __global__ void __launch_bounds__(256) k_plane_dgroup_ld(const float4* __restrict__ feat, float4* out) {
    const long long base = (long long)blockIdx.z * DQ * PLANE4 + (long long)blockIdx.y * DG * PLANE4;
    const int p0 = blockIdx.x * 512;
    float4 v[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int p = p0 + k * 256 + threadIdx.x;
        v[k] = p < PLANE4 ? __ldg(feat + (long long)(blockIdx.z % 64) * PLANE4 + p) : make_float4(0, 0, 0, 0);
    }
    for (int d = 0; d < DG; ++d)
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int p = p0 + k * 256 + threadIdx.x;
            if (p < PLANE4) __stcs(out + base + (long long)d * PLANE4 + p, v[k]);
        }
}

template <class F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 10;
}

int main() {
    float4* out; cudaMalloc(&out, TOTAL4 * 16);
    const float4 v = make_float4(1, 2, 3, 4);
    const double gb = TOTAL4 * 16 / 1e9;
    auto rep = [&](const char* n, float ms) { printf("%-44s %.4f ms  %.0f GB/s\n", n, ms, gb / ms * 1e3); };
    rep("memset", timeit([&] { cudaMemsetAsync(out, 0, TOTAL4 * 16); }));
    rep("linear grid-stride 148*8 CTAs, default", timeit([&] { k_linear<false><<<148 * 8, 256>>>(out, TOTAL4, v); }));
    rep("linear grid-stride 148*8 CTAs, .cs", timeit([&] { k_linear<true><<<148 * 8, 256>>>(out, TOTAL4, v); }));
    rep("linear one-shot (1 float4/thread), default", timeit([&] { k_linear<false><<<(unsigned)(TOTAL4 / 256), 256>>>(out, TOTAL4, v); }));
    rep("linear one-shot (1 float4/thread), .cs", timeit([&] { k_linear<true><<<(unsigned)(TOTAL4 / 256), 256>>>(out, TOTAL4, v); }));
    rep("plane sweep 8KB x 48 planes, .cs (concat)", timeit([&] { k_plane_sweep<true, 2><<<dim3(16, NOC * NB), 256>>>(out, v); }));
    rep("plane sweep 8KB x 48 planes, default", timeit([&] { k_plane_sweep<false, 2><<<dim3(16, NOC * NB), 256>>>(out, v); }));
    rep("plane sweep 4KB x 48 planes, .cs", timeit([&] { k_plane_sweep<true, 1><<<dim3(32, NOC * NB), 256>>>(out, v); }));
    rep("plane sweep 32KB x 48 planes, .cs", timeit([&] { k_plane_sweep<true, 8><<<dim3(4, NOC * NB), 256>>>(out, v); }));
    rep("CTA-contiguous 393KB, .cs", timeit([&] { k_cta_contig<true><<<(unsigned)((TOTAL4 + 48 * 512 - 1) / (48 * 512)), 256>>>(out, TOTAL4, v); }));
    rep("CTA-contiguous 393KB, default", timeit([&] { k_cta_contig<false><<<(unsigned)((TOTAL4 + 48 * 512 - 1) / (48 * 512)), 256>>>(out, TOTAL4, v); }));
    rep("plane sweep, 8 planes per CTA, .cs", timeit([&] { k_plane_dgroup<true, 8><<<dim3(16, 6, NOC * NB), 256>>>(out, v); }));
    rep("plane sweep, 1 plane per CTA, .cs", timeit([&] { k_plane_dgroup<true, 1><<<dim3(16, 48, NOC * NB), 256>>>(out, v); }));
    rep("plane sweep, 1 plane per CTA, default", timeit([&] { k_plane_dgroup<false, 1><<<dim3(16, 48, NOC * NB), 256>>>(out, v); }));
    float4* feat; cudaMalloc(&feat, 64 * PLANE4 * 16); cudaMemset(feat, 0, 64 * PLANE4 * 16);
    rep("plane sweep, 8 planes per CTA, loaded value", timeit([&] { k_plane_dgroup_ld<8><<<dim3(16, 6, NOC * NB), 256>>>(feat, out); }));
    rep("plane sweep, 16 planes per CTA, loaded value", timeit([&] { k_plane_dgroup_ld<16><<<dim3(16, 3, NOC * NB), 256>>>(feat, out); }));
    rep("plane sweep, 48 planes per CTA, loaded value", timeit([&] { k_plane_dgroup_ld<48><<<dim3(16, 1, NOC * NB), 256>>>(feat, out); }));
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
