// cuda_emu.h -- TEST-ONLY single-threaded emulation of the CUDA subset the SIMT kernels use.
//
// Purpose: this container has no GPU and every gpurun call costs minutes, so the `-m "not gpu"`
// tests compile csrc/e2t.cu a second time with g++ (-DE2T_EMU -include this file) and run the very
// same kernel source and host orchestration on the CPU to catch indexing / sequencing bugs before
// any GPU time is spent.  It is NOT a product path: the ecog2txt_b200 package only ever loads
// libe2t.so (nvcc, sm_100a) and raises if it is missing; nothing outside tests/ references this.
//
// Model: blocks run one after another; the threads of a block are fibers (hand-rolled x86-64
// context switch) scheduled round-robin, yielding at __syncthreads / warp shuffles.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
struct int4 { int x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return {x, y}; }

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };

namespace emu {
extern dim3 g_gridDim, g_blockDim, g_blockIdx, g_threadIdx;
extern char* g_dyn_smem;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void sync_block();
void sync_warp();
int lane();
float shfl_f(float v, int src_lane);
}  // namespace emu

#define gridDim emu::g_gridDim
#define blockDim emu::g_blockDim
#define blockIdx emu::g_blockIdx
#define threadIdx emu::g_threadIdx

static inline void __syncthreads() { emu::sync_block(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::sync_warp(); }
// blocks run one after another and fibers only switch at sync points, so a plain static counter is race-free
static inline int __syncthreads_count(int pred) {
  static int count = 0, result = 0;
  if (threadIdx.x == 0) count = 0;
  emu::sync_block();
  count += pred ? 1 : 0;
  emu::sync_block();
  if (threadIdx.x == 0) result = count;
  emu::sync_block();
  return result;
}
static inline float __shfl_xor_sync(unsigned, float v, int m) { return emu::shfl_f(v, emu::lane() ^ m); }
static inline float __shfl_down_sync(unsigned, float v, int d) {
  int s = emu::lane() + d;
  return emu::shfl_f(v, s < 32 ? s : emu::lane());
}
static inline float __shfl_sync(unsigned, float v, int s) { return emu::shfl_f(v, s); }
static inline int __shfl_xor_sync(unsigned, int v, int m) {
  float f; memcpy(&f, &v, 4); f = emu::shfl_f(f, emu::lane() ^ m); memcpy(&v, &f, 4); return v;
}
static inline int __shfl_down_sync(unsigned, int v, int d) {
  float f; memcpy(&f, &v, 4); int s = emu::lane() + d; f = emu::shfl_f(f, s < 32 ? s : emu::lane());
  memcpy(&v, &f, 4); return v;
}
static inline int __shfl_sync(unsigned, int v, int s) {
  float f; memcpy(&f, &v, 4); f = emu::shfl_f(f, s); memcpy(&v, &f, 4); return v;
}
static inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline void __threadfence() {}
static inline int atomicMax(int* p, int v) { int o = *p; *p = std::max(o, v); return o; }
// fast-math intrinsics: plain libm here
#define __expf(x) expf(x)
#define __logf(x) logf(x)
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __ldg(const float* p) { return *p; }
static inline int __ldg(const int* p) { return *p; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
using std::max;
using std::min;

// runtime API subset
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }

#define E2T_LAUNCH(kern, grid, block, smem, stream, ...) \
  emu::launch((grid), (block), (smem), [&]() { kern(__VA_ARGS__); })
#define E2T_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::g_dyn_smem)
