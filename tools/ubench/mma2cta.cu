// Micro-benchmark: K loop of a tf32 GEMM with cta_group::2 (CTA pair, M = 256, each SM holds its 128 rows of A and HALF of B).
// Question: does halving the B ingest per SM lift the single-CTA loop (~790 cycles per 128x256x32 stage, pipe floor 512)?
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "kernels_simt.cuh"
#include "gemm_tc.cuh"
using namespace tc;
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion is signalled on the LEADER CTA's mbarrier (peer bit of the shared::cluster address cleared)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
template <int BN>
__global__ void __launch_bounds__(192, 1) k_ub2(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                int num_kb, int kb_wrap, int stages, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr uint32_t A_BYTES = 128 * BK * 4, B_BYTES = (BN / 2) * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stages * STAGE_BYTES);
  uint64_t* full_bar = bars; uint64_t* empty_bar = bars + stages; uint64_t* done_bar = bars + 2 * stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ctarank();
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(done_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_before_sync(); __syncthreads(); cluster_sync_(); fence_after_sync();
  const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, *tmem_slot);
  const int pair = blockIdx.x / 2;
  const int m0 = (pair % 32) * 256 + (int)rank * 128;     // this CTA's 128 rows of the 256-row tile
  const int n0 = (int)rank * (BN / 2);                    // its half of the B rows
  long long t0 = clock64();
  if (warp == 0) {
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % stages; const uint32_t ph = (kb / stages) & 1;
      mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
      if (elect_one()) {
        const uint32_t fb = smem_u32(&full_bar[s]);
        if (rank == 0) mbar_expect_tx(fb, 2 * STAGE_BYTES);      // both CTAs' loads complete on the leader's barrier
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        tma_load_2d_2sm(sa, &map_a, fb, (kb % kb_wrap) * BK, m0);
        tma_load_2d_2sm(sa + A_BYTES, &map_b, fb, (kb % kb_wrap) * BK, n0);
      }
      __syncwarp();
    }
  } else if (warp == 1 && rank == 0) {
    const uint32_t idesc = make_idesc_tf32(256, BN, 0);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % stages; const uint32_t ph = (kb / stages) & 1;
      mbar_wait(smem_u32(&full_bar[s]), ph);
      fence_after_sync();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_tf32_2sm(tmem_base, make_smem_desc(sa + k * UMMA_K * 4), make_smem_desc(sa + A_BYTES + k * UMMA_K * 4), idesc, 1u);
        umma_commit_2sm(smem_u32(&empty_bar[s]));
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit_2sm(smem_u32(done_bar));
    __syncwarp();
  }
  if (warp == 1) {
    mbar_wait(smem_u32(done_bar), 0);
    if (lane == 0) out[blockIdx.x] = clock64() - t0;
  }
  fence_before_sync(); __syncthreads(); cluster_sync_();
  if (warp == 1) { fence_after_sync(); asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory"); }
}
template <int BN> void run(int stages, float* dA, float* dB, int M, int K, int num_kb, long long* dcy) {
  CUtensorMap ma = make_map(dA, M, K, K, 128);
  CUtensorMap mb = make_map(dB, BN, K, K, BN / 2);
  size_t smem = (size_t)stages * (128 * BK * 4 + (BN / 2) * BK * 4) + (2 * stages + 2) * 8 + 1024;
  auto kfn = k_ub2<BN>;
  cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, kfn, ma, mb, num_kb, K / BK, stages, dcy);
    if (e != cudaSuccess) { printf("launch err %s\n", cudaGetErrorString(e)); exit(1); }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
  }
  std::vector<long long> cy(148);
  cudaMemcpy(cy.data(), dcy, 148 * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (auto c : cy) avg += c; avg /= 148;
  printf("2-CTA M=256 BN=%3d stages=%d: %.0f cyc per k-block (32 of K), MMA pipe floor %d, per-SM stage bytes %d\n", BN, stages, avg / num_kb,
         BN / 2 * 4, 128 * BK * 4 + (BN / 2) * BK * 4);
}
int main() {
  const int M = 8704, K = 800;
  float *dA, *dB; long long* dcy;
  cudaMalloc(&dA, (size_t)M * K * 4); cudaMalloc(&dB, (size_t)256 * K * 4); cudaMalloc(&dcy, 1024 * 8);
  cudaMemset(dA, 0, (size_t)M * K * 4); cudaMemset(dB, 0, (size_t)256 * K * 4);
  run<256>(6, dA, dB, M, K, 2000, dcy);
  run<128>(6, dA, dB, M, K, 2000, dcy);
  return 0;
}
