// Micro-benchmark of the mbarrier producer/consumer handshake cost (no TMA, no MMA).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void wait_bounded(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void wait_loop(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void wait_test(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void wait_hint(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity), "r"(10000000u) : "memory");
}
__device__ __forceinline__ void wait_test_sleep(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(20);
  }
}
template <int W> __device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
  if (W == 0) wait_bounded(bar, parity); else if (W == 1) wait_loop(bar, parity); else if (W == 2) wait_test(bar, parity);
  else if (W == 3) wait_hint(bar, parity); else wait_test_sleep(bar, parity);
}
// variant: 0 = two threads in different warps ping-pong over `stages` barriers; 1 = + tcgen05 fence in consumer;
// 2 = single thread arrive+wait on its own barrier; 3 = producer/consumer in same warp? (n/a)
template <int W>
__global__ void k(int n, int stages, int variant, long long* out, int stride) {
  __shared__ __align__(128) uint64_t bars[16 * 16 * 2];
  uint64_t* full_bar = bars; uint64_t* empty_bar = bars + 8 * stride;
#define full_bar(i) full_bar[(i) * stride]
#define empty_bar(i) empty_bar[(i) * stride]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) { mbar_init(smem_u32(&full_bar(s)), 1); mbar_init(smem_u32(&empty_bar(s)), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  if (variant == 2) {
    if (threadIdx.x == 0) {
      for (int i = 0; i < n; ++i) { mbar_arrive(smem_u32(&full_bar(0))); wait<W>(smem_u32(&full_bar(0)), i & 1); }
      out[blockIdx.x] = clock64() - t0;
    }
    return;
  }
  if (warp == 1 && lane == 0) {
    for (int kb = 0; kb < n; ++kb) {
      const int s = kb % stages; const uint32_t ph = (kb / stages) & 1;
      wait<W>(smem_u32(&empty_bar(s)), ph ^ 1);
      mbar_arrive(smem_u32(&full_bar(s)));
    }
  } else if (warp == 2 && lane == 0) {
    for (int kb = 0; kb < n; ++kb) {
      const int s = kb % stages; const uint32_t ph = (kb / stages) & 1;
      wait<W>(smem_u32(&full_bar(s)), ph);
      if (variant == 1) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      mbar_arrive(smem_u32(&empty_bar(s)));
    }
    out[blockIdx.x] = clock64() - t0;
  }
}
template <int W> void run(const char* name, long long* d) {
  for (int stride : {1})
  for (int variant = 0; variant < 3; variant += 2)
    for (int stages : {1, 6}) {
      if (variant == 2 && stages != 1) continue;
      const int n = 4000;
      k<W><<<148, 128>>>(n, stages, variant, d, stride); k<W><<<148, 128>>>(n, stages, variant, d, stride);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return; }
      long long h[148]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
      double a = 0; for (auto c : h) a += c;
      printf("%-8s stride=%2d variant=%d stages=%d: %.0f cyc/iter\n", name, stride, variant, stages, a / 148 / n);
    }
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  run<1>("loop", d); run<3>("hint", d); run<4>("testsleep", d);
  return 0;
}
