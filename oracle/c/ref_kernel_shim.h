/* Shim that lets g++ compile the reference's CUDA-C kernel string
 * (/root/reference/utils/warp_ops.py:20-47) for the CPU: the kernels use only
 * __global__ and the x components of blockIdx/blockDim/threadIdx.
 * TEST INFRASTRUCTURE ONLY (see oracle/build_ref.py). */
#pragma once
#define __global__
struct az_dim3_shim { int x, y, z; };
static thread_local az_dim3_shim blockIdx, blockDim, threadIdx;
