// Micro-benchmark: what bounds a tf32 tcgen05 K-loop on B200 -- TMA ingest, MMA issue or the smem port?
//   mode 0: TMA + MMA pipeline; mode 1: TMA only; mode 2: MMA only (no loads)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ecog2txt_b200/csrc -o tma_mma tools/ubench/tma_mma.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "kernels_simt.cuh"
#include "gemm_tc.cuh"
using namespace tc;

template <int BN>
__global__ void __launch_bounds__(192, 1)
k_ub(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int num_kb, int kb_wrap, int mode,
     int stages, int m_tiles, long long* out_cycles, int R) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr uint32_t A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stages * STAGE_BYTES);
  uint64_t* full_bar = bars; uint64_t* empty_bar = bars + stages; uint64_t* done_bar = bars + 2 * stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 4 && lane == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(done_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 5) tmem_alloc(smem_u32(tmem_slot), BN < 32 ? 32 : BN);
  fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int m0 = (blockIdx.x % m_tiles) * BM;
  const int n0 = 0;
  long long t0 = clock64();
  if (mode == 10 || mode == 11) {
    // like 8/9 but only lane 0 polls the mbarrier, then the warp reconverges
    if (warp == 4) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % stages; const uint32_t ph = (kb / stages) & 1;
        if (lane == 0) mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
        __syncwarp();
        const uint32_t fb = smem_u32(&full_bar[s]);
        if (elect_one()) {
          if (mode == 10) {
            mbar_expect_tx(fb, STAGE_BYTES);
            const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
            tma_load_2d(sa, &map_a, fb, (kb % kb_wrap) * BK, m0);
            tma_load_2d(sa + A_BYTES, &map_b, fb, (kb % kb_wrap) * BK, n0);
          } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fb) : "memory");
          }
        }
        __syncwarp();
      }
    } else if (warp == 5) {
      constexpr uint32_t idesc = make_idesc_tf32(BM, BN, 0);
      const uint32_t tb = __reduce_max_sync(0xffffffffu, tmem_base);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % stages; const uint32_t ph = (kb / stages) & 1;
        if (lane == 0) mbar_wait(smem_u32(&full_bar[s]), ph);
        __syncwarp();
        fence_after_sync();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_tf32(tb, make_smem_desc(sa + k * UMMA_K * 4), make_smem_desc(sa + A_BYTES + k * UMMA_K * 4), idesc, 1u);
          umma_commit(smem_u32(&empty_bar[s]));
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(smem_u32(done_bar));
      __syncwarp();
      mbar_wait(smem_u32(done_bar), 0);
      if (lane == 0) out_cycles[blockIdx.x] = clock64() - t0;
    }
  } else if (mode == 8 || mode == 9) {
    // warp-convergent roles: every lane runs the loop and the waits, one elected lane issues (mode 9: MMA only)
    if (warp == 4) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % stages; const uint32_t ph = (kb / stages) & 1;
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
        const uint32_t fb = smem_u32(&full_bar[s]);
        if (elect_one()) {
          if (mode == 8) {
            mbar_expect_tx(fb, STAGE_BYTES);
            const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
            tma_load_2d(sa, &map_a, fb, (kb % kb_wrap) * BK, m0);
            tma_load_2d(sa + A_BYTES, &map_b, fb, (kb % kb_wrap) * BK, n0);
          } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fb) : "memory");
          }
        }
        __syncwarp();
      }
    } else if (warp == 5) {
      constexpr uint32_t idesc = make_idesc_tf32(BM, BN, 0);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % stages; const uint32_t ph = (kb / stages) & 1;
        mbar_wait(smem_u32(&full_bar[s]), ph);
        fence_after_sync();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_tf32(tmem_base, make_smem_desc(sa + k * UMMA_K * 4), make_smem_desc(sa + A_BYTES + k * UMMA_K * 4), idesc, 1u);
          umma_commit(smem_u32(&empty_bar[s]));
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(smem_u32(done_bar));
      __syncwarp();
      mbar_wait(smem_u32(done_bar), 0);
      if (lane == 0) out_cycles[blockIdx.x] = clock64() - t0;
    }
  } else if (mode == 7) {
    // issue-cost probe: single thread, clock64 around each operation class
    if (warp == 5 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BM, BN, 0);
      long long d_tma = 0, d_wfull = 0, d_fence = 0, d_mma = 0, d_commit = 0, d_wcommit = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % stages; const uint32_t ph = kb & 1;
        const uint32_t fb = smem_u32(&full_bar[0]);
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        long long c0 = clock64();
        mbar_expect_tx(fb, STAGE_BYTES);
        tma_load_2d(sa, &map_a, fb, (kb % kb_wrap) * BK, m0);
        tma_load_2d(sa + A_BYTES, &map_b, fb, (kb % kb_wrap) * BK, n0);
        long long c1 = clock64();
        mbar_wait(fb, ph);
        long long c2 = clock64();
        fence_after_sync();
        long long c3 = clock64();
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_tf32(tmem_base, make_smem_desc(sa + k * UMMA_K * 4), make_smem_desc(sa + A_BYTES + k * UMMA_K * 4), idesc, 1u);
        long long c4 = clock64();
        umma_commit(smem_u32(&empty_bar[0]));
        long long c5 = clock64();
        mbar_wait(smem_u32(&empty_bar[0]), ph);
        long long c6 = clock64();
        d_tma += c1 - c0; d_wfull += c2 - c1; d_fence += c3 - c2; d_mma += c4 - c3; d_commit += c5 - c4; d_wcommit += c6 - c5;
      }
      out_cycles[blockIdx.x] = clock64() - t0;
      if (blockIdx.x == 0)
        printf("  probe BN=%d: tma_issue %lld  wait_full %lld  fence %lld  mma_issue(4) %lld  commit %lld  wait_commit %lld\n", BN,
               d_tma / num_kb, d_wfull / num_kb, d_fence / num_kb, d_mma / num_kb, d_commit / num_kb, d_wcommit / num_kb);
    }
  } else if (mode == 5 || mode == 6) {
    // single thread issues TMA and MMA: every wait is on a hardware-async completion (complete_tx / tcgen05.commit)
    if (warp == 5 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BM, BN, 0);
      auto load = [&](int kb) {
        const int s = kb % stages;
        const uint32_t fb = smem_u32(&full_bar[s]);
        mbar_expect_tx(fb, STAGE_BYTES);
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        tma_load_2d(sa, &map_a, fb, (kb % kb_wrap) * BK, m0);
        tma_load_2d(sa + A_BYTES, &map_b, fb, (kb % kb_wrap) * BK, n0);
      };
      for (int kb = 0; kb < stages - R && kb < num_kb; ++kb) load(kb);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % stages; const uint32_t ph = (kb / stages) & 1;
        mbar_wait(smem_u32(&full_bar[s]), ph);
        fence_after_sync();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_tf32(tmem_base, make_smem_desc(sa + k * UMMA_K * 4), make_smem_desc(sa + A_BYTES + k * UMMA_K * 4), idesc, 1u);
        umma_commit(smem_u32(&empty_bar[s]));
        // refill the stage freed by the PREVIOUS k-block (its MMAs are ahead of ours in the tensor pipe)
        const int nk = kb + stages - R;
        if (nk < num_kb) {
          if (kb >= R) { const int ps = (kb - R) % stages; mbar_wait(smem_u32(&empty_bar[ps]), ((kb - R) / stages) & 1); }
          if (mode == 5) load(nk);
          else { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[nk % stages])) : "memory"); }
        }
      }
      umma_commit(smem_u32(done_bar)); mbar_wait(smem_u32(done_bar), 0);
      out_cycles[blockIdx.x] = clock64() - t0;
    }
  } else if (mode == 3) {
    if (warp == 5 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BM, BN, 0);
      for (int kb = 0; kb < num_kb; ++kb) {
        const uint32_t sa = smem_u32(smem + (kb % stages) * STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_tf32(tmem_base, make_smem_desc(sa + k * UMMA_K * 4), make_smem_desc(sa + A_BYTES + k * UMMA_K * 4), idesc, 1u);
      }
      umma_commit(smem_u32(done_bar)); mbar_wait(smem_u32(done_bar), 0);
      out_cycles[blockIdx.x] = clock64() - t0;
    }
  } else if (warp == 4 && lane == 0) {
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % stages; const uint32_t ph = (kb / stages) & 1;
      mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
      const uint32_t fb = smem_u32(&full_bar[s]);
      if (mode != 2 && mode != 4) {
        mbar_expect_tx(fb, STAGE_BYTES);
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        tma_load_2d(sa, &map_a, fb, (kb % kb_wrap) * BK, m0);
        tma_load_2d(sa + A_BYTES, &map_b, fb, (kb % kb_wrap) * BK, n0);
      } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fb) : "memory");
      }
    }
  } else if (warp == 5 && lane == 0) {
    constexpr uint32_t idesc = make_idesc_tf32(BM, BN, 0);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % stages; const uint32_t ph = (kb / stages) & 1;
      mbar_wait(smem_u32(&full_bar[s]), ph);
      fence_after_sync();
      const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
      if (mode != 1 && mode != 4) {
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_tf32(tmem_base, make_smem_desc(sa + k * UMMA_K * 4), make_smem_desc(sa + A_BYTES + k * UMMA_K * 4), idesc, 1u);
        umma_commit(smem_u32(&empty_bar[s]));
      } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty_bar[s])) : "memory");
      }
    }
    if (mode != 1 && mode != 4) { umma_commit(smem_u32(done_bar)); mbar_wait(smem_u32(done_bar), 0); }
    long long t1 = clock64();
    out_cycles[blockIdx.x] = t1 - t0;
  }
  fence_before_sync(); __syncthreads();
  if (warp == 5) { fence_after_sync(); tmem_dealloc(tmem_base, BN < 32 ? 32 : BN); }
}

template <int BN>
void run(int mode, int stages, int grid, float* dA, float* dB, int M, int K, int num_kb, long long* dcy, int R = 1) {
  CUtensorMap ma = make_map(dA, M, K, K, BM);
  CUtensorMap mb = make_map(dB, BN, K, K, BN);
  size_t smem = (size_t)stages * (BM * BK * 4 + BN * BK * 4) + (2 * stages + 2) * 8 + 1024;
  auto kfn = k_ub<BN>;
  cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int m_tiles = M / BM;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    kfn<<<grid, 192, smem>>>(ma, mb, num_kb, K / BK, mode, stages, m_tiles, dcy, R);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
  }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> cy(grid);
  cudaMemcpy(cy.data(), dcy, grid * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (auto c : cy) avg += c; avg /= grid;
  const double bytes = (double)(BM + BN) * BK * 4;
  printf("BN=%3d mode=%d R=%d stages=%d grid=%3d: %.1f us, %.0f cyc/kblock, TMA %.1f B/clk/SM, MMA floor %d cyc/kblock, agg %.2f TB/s\n", BN,
         mode, R, stages, grid, ms * 1e3, avg / num_kb, mode == 2 ? 0.0 : bytes / (avg / num_kb), BM * BN / 256 * 4,
         mode == 2 ? 0.0 : bytes * num_kb * grid / (ms * 1e-3) / 1e12);
}

int main() {
  const int M = 8704, K = 800;
  float *dA, *dB; long long* dcy;
  cudaMalloc(&dA, (size_t)M * K * 4); cudaMalloc(&dB, (size_t)256 * K * 4); cudaMalloc(&dcy, 1024 * 8);
  cudaMemset(dA, 0, (size_t)M * K * 4); cudaMemset(dB, 0, (size_t)256 * K * 4);
  const int nkb = 2000;
  for (int grid : {148}) {
    for (int mode : {8, 10, 11}) for (int R : {1}) {
      run<64>(mode, 6, grid, dA, dB, M, K, nkb, dcy, R);
      run<128>(mode, 6, grid, dA, dB, M, K, nkb, dcy, R);
      run<256>(mode, 4, grid, dA, dB, M, K, nkb, dcy, R);
    }
  }
  return 0;
}
