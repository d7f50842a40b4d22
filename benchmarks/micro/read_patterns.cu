// Load-bandwidth microbenchmark (not part of the library): what a pure load stream reaches on this B200 with the
// access patterns of the soft-argmin / concat-backward kernels (a thread owns one float4 of pixels and walks the
// D planes, 8 independent 128-bit loads in flight), against a linear read.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o read_patterns read_patterns.cu && ./read_patterns
#include <cstdio>
#include <cuda_runtime.h>

constexpr long long HW4 = 544LL * 960 / 4;  // float4 per plane (130560)
constexpr int D = 192, NB = 8;
constexpr long long TOTAL4 = HW4 * D * NB;

template <bool CS> __device__ __forceinline__ float4 ld(const float4* p) { return CS ? __ldcs(p) : __ldg(p); }
__device__ __forceinline__ void acc4(float4& a, const float4& v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }

template <bool CS> __global__ void k_linear(const float4* in, long long n, float* sink) {
    float4 a = make_float4(0, 0, 0, 0);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        float4 v0 = ld<CS>(in + i), v1 = ld<CS>(in + i + stride), v2 = ld<CS>(in + i + 2 * stride), v3 = ld<CS>(in + i + 3 * stride);
        acc4(a, v0); acc4(a, v1); acc4(a, v2); acc4(a, v3);
    }
    for (; i < n; i += stride) acc4(a, ld<CS>(in + i));
    if (a.x + a.y + a.z + a.w == 123.456f) *sink = 1.f;
}
// plane walk: grid = (ceil(HW4/256), NB, D/DG): a thread reads its float4 from DG consecutive planes, U in flight
template <bool CS, int DG, int U> __global__ void __launch_bounds__(256) k_plane_walk(const float4* in, float* sink) {
    const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
    if (p >= HW4) return;
    const float4* q = in + ((long long)blockIdx.y * D + (long long)blockIdx.z * DG) * HW4 + p;
    float4 a = make_float4(0, 0, 0, 0);
    for (int d = 0; d < DG; d += U) {
        float4 v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) v[k] = ld<CS>(q + (long long)(d + k) * HW4);
#pragma unroll
        for (int k = 0; k < U; ++k) acc4(a, v[k]);
    }
    if (a.x + a.y + a.z + a.w == 123.456f) *sink = 1.f;
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
    float4* in; float* sink; cudaMalloc(&in, TOTAL4 * 16); cudaMalloc(&sink, 4); cudaMemset(in, 0, TOTAL4 * 16);
    const double gb = TOTAL4 * 16 / 1e9;
    auto rep = [&](const char* n, float ms) { printf("%-52s %.4f ms  %.0f GB/s\n", n, ms, gb / ms * 1e3); };
    const unsigned gx = (unsigned)((HW4 + 255) / 256);
    rep("linear grid-stride 148*8 CTAs, ldg", timeit([&] { k_linear<false><<<148 * 8, 256>>>(in, TOTAL4, sink); }));
    rep("linear grid-stride 148*8 CTAs, ld.cs", timeit([&] { k_linear<true><<<148 * 8, 256>>>(in, TOTAL4, sink); }));
    rep("linear grid-stride 148*16 CTAs x 512 thr, ld.cs", timeit([&] { k_linear<true><<<148 * 16, 512>>>(in, TOTAL4, sink); }));
    rep("plane walk 192 planes/CTA, 8 in flight, ld.cs", timeit([&] { k_plane_walk<true, 192, 8><<<dim3(gx, NB, 1), 256>>>(in, sink); }));
    rep("plane walk 192 planes/CTA, 8 in flight, ldg", timeit([&] { k_plane_walk<false, 192, 8><<<dim3(gx, NB, 1), 256>>>(in, sink); }));
    rep("plane walk 192 planes/CTA, 16 in flight, ld.cs", timeit([&] { k_plane_walk<true, 192, 16><<<dim3(gx, NB, 1), 256>>>(in, sink); }));
    rep("plane walk 48 planes/CTA, 8 in flight, ld.cs", timeit([&] { k_plane_walk<true, 48, 8><<<dim3(gx, NB, 4), 256>>>(in, sink); }));
    rep("plane walk 16 planes/CTA, 8 in flight, ld.cs", timeit([&] { k_plane_walk<true, 16, 8><<<dim3(gx, NB, 12), 256>>>(in, sink); }));
    rep("plane walk 8 planes/CTA, 8 in flight, ld.cs", timeit([&] { k_plane_walk<true, 8, 8><<<dim3(gx, NB, 24), 256>>>(in, sink); }));
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
