// Micro-benchmark: per-SM TMA ingest rate of a 128 x 400 fp32 tile (13 boxes of 128 x 32) from L2 when many CTAs pull
// the same tile (the persistent LSTM's per-step h exchange), distinct tiles, or a cluster shares it by multicast.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "kernels_simt.cuh"
#include "gemm_tc.cuh"
using namespace tc;
__device__ __forceinline__ uint32_t cluster_ctarank_() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t m) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(m) : "memory");
}
// mode 0: every CTA loads tile (blockIdx % ntiles); mode 1: cluster multicast (each CTA issues chunks kk % csz == rank)
__global__ void __launch_bounds__(128, 1) k_ingest(const __grid_constant__ CUtensorMap map, int iters, int ntiles, int nchunk, int csz, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + nchunk * 16384);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (csz > 1) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  const int rank = csz > 1 ? (int)cluster_ctarank_() : 0;
  const int tile = (blockIdx.x / csz) % ntiles;
  long long t0 = clock64();
  if (warp == 0) {
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
        mbar_expect_tx(smem_u32(bar), nchunk * 16384);
        for (int kc = 0; kc < nchunk; ++kc) {
          if (csz == 1) tma_load_2d(smem_u32(smem + kc * 16384), &map, smem_u32(bar), kc * 32, tile * 128);
          else if (kc % csz == rank) tma_load_2d_mc(smem_u32(smem + kc * 16384), &map, smem_u32(bar), kc * 32, tile * 128, (uint16_t)((1u << csz) - 1));
        }
      }
      __syncwarp();
      mbar_wait(smem_u32(bar), it & 1);
      if (csz > 1) { /* peers may still be receiving into our smem next round only after everyone passed: cluster barrier */
      }
    }
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
  }
  if (csz > 1) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
}
int main() {
  const int rows = 128 * 128, cols = 416;
  float* d; long long* dcy;
  cudaMalloc(&d, (size_t)rows * cols * 4); cudaMemset(d, 0, (size_t)rows * cols * 4); cudaMalloc(&dcy, 1024 * 8);
  CUtensorMap m = make_map(d, rows, 400, cols, 128);
  const int nchunk = 13, iters = 200;
  size_t smem = nchunk * 16384 + 64 + 1024;
  cudaFuncSetAttribute(k_ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  struct Cfg { int grid, ntiles, csz; } cfgs[] = {{100, 4, 1}, {100, 100, 1}, {148, 4, 1}, {148, 148, 1}, {1, 1, 1}, {100, 4, 5}, {100, 4, 2}, {96, 4, 8}, {100, 100, 5}};
  for (auto c : cfgs) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(c.grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = c.csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
      cudaError_t e = cudaLaunchKernelEx(&cfg, k_ingest, m, iters, c.ntiles, nchunk, c.csz, dcy);
      if (e != cudaSuccess) { printf("launch err %s\n", cudaGetErrorString(e)); return 1; }
      e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    }
    std::vector<long long> cy(c.grid);
    cudaMemcpy(cy.data(), dcy, c.grid * 8, cudaMemcpyDeviceToHost);
    double a = 0; for (auto v : cy) a += v; a /= c.grid;
    printf("grid=%3d distinct tiles=%3d cluster=%d: %.0f cyc per 208 KB tile -> %.1f B/clk/SM\n", c.grid, c.ntiles, c.csz, a / iters, nchunk * 16384.0 / (a / iters));
  }
  return 0;
}
