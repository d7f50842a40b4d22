// Which CTAs become resident on an SM beside a big resident CTA of another stream?
// Kernel A ("host"): 148 CTAs x TA threads, RA registers (template), SA bytes of dynamic shared memory; spins ~200 us.
// Kernel B ("guest"): many CTAs of TB threads with <= 32 or <= 64 registers; each records the global timer at start.
// A is launched first on a high-priority stream, B on a second stream; the report is the fraction of B's CTAs that
// STARTED before A ended (0 = the kernels serialise).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o coresidency coresidency.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

template <int REGS>
__global__ void __launch_bounds__(512, 1) host_kernel(unsigned long long* span, float* sink, long long spin_ns, int has_smem) {
    extern __shared__ float sm[];
    // keep ~REGS live values so that ptxas allocates them
    float v[REGS > 100 ? 118 : 8];
#pragma unroll
    for (int i = 0; i < (REGS > 100 ? 118 : 8); ++i) v[i] = threadIdx.x * 0.001f + i;
    const unsigned long long t0 = gtimer();
    if (threadIdx.x == 0 && has_smem) sm[0] = 1.f;
    while ((long long)(gtimer() - t0) < spin_ns) {
#pragma unroll
        for (int i = 0; i < (REGS > 100 ? 118 : 8); ++i) v[i] = fmaf(v[i], 1.0001f, 0.5f);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < (REGS > 100 ? 118 : 8); ++i) s += v[i];
    if (s == 12345.f) sink[0] = s + (has_smem ? sm[0] : 0.f);
    if (threadIdx.x == 0) {
        atomicMin(&span[0], t0);
        atomicMax(&span[1], gtimer());
    }
}

__global__ void guest_kernel(unsigned long long* starts, float* out) {
    if (threadIdx.x == 0) starts[blockIdx.x] = gtimer();
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    float4* o = reinterpret_cast<float4*>(out) + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) __stcs(o + i, v);
}

template <int REGS>
void run(int ta, int sa, int tb, int nb, int carve_guest) {
    unsigned long long *span, *starts;
    float *sink, *out;
    cudaMalloc(&span, 16);
    cudaMalloc(&starts, nb * 8);
    cudaMalloc(&sink, 4);
    cudaMalloc(&out, (size_t)nb * tb * 16 * 16);
    unsigned long long init[2] = {~0ull, 0ull};
    cudaMemcpy(span, init, 16, cudaMemcpyHostToDevice);
    cudaMemset(starts, 0, nb * 8);
    cudaFuncSetAttribute(host_kernel<REGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, sa);
    if (carve_guest >= 0) cudaFuncSetAttribute(guest_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve_guest);
    int lo, hi;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    cudaStream_t s1, s2;
    cudaStreamCreateWithPriority(&s1, cudaStreamNonBlocking, hi);
    cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, lo);
    cudaDeviceSynchronize();
    host_kernel<REGS><<<148, ta, sa, s1>>>(span, sink, 200000, sa > 0);
    guest_kernel<<<nb, tb, 0, s2>>>(starts, out);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<unsigned long long> st(nb);
    unsigned long long sp[2];
    cudaMemcpy(st.data(), starts, nb * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(sp, span, 16, cudaMemcpyDeviceToHost);
    int during = 0;
    for (int i = 0; i < nb; ++i) during += (st[i] >= sp[0] && st[i] < sp[1]);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, host_kernel<REGS>);
    printf("{\"host_threads\": %d, \"host_regs\": %d, \"host_smem\": %d, \"guest_threads\": %d, \"guest_carveout\": %d, "
           "\"guest_ctas\": %d, \"started_during_host\": %d, \"frac\": %.3f, \"host_us\": %.1f, \"err\": \"%s\"}\n",
           ta, fa.numRegs, sa, tb, carve_guest, nb, during, (double)during / nb, (sp[1] - sp[0]) / 1000.0, cudaGetErrorString(e));
    cudaFree(span); cudaFree(starts); cudaFree(sink); cudaFree(out);
    cudaStreamDestroy(s1); cudaStreamDestroy(s2);
}

int main() {
    const int nb = 148 * 64;
    run<8>(480, 0, 32, nb, -1);  // warm-up (lazy module load delays the first guest launch): discard
    printf("# light host (18 registers): shared memory of the host decides\n");
    for (int kb : {0, 100, 200, 216, 224, 227}) run<8>(480, kb * 1024, 32, nb, kb >= 200 ? 100 : -1);
    printf("# register-heavy host (126 registers), no shared memory: 15, 14, 13, 12 warps; guests of 1 and 4 warps\n");
    for (int ta : {480, 448, 416, 384}) {
        run<128>(ta, 0, 32, nb, -1);
        run<128>(ta, 0, 128, nb / 4, -1);
    }
    return 0;
}
