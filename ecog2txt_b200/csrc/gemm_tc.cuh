// gemm_tc.cuh -- tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = A[M,K] * B[N,K]^T (+bias) (+beta*C)
//
// Both operands are K-major fp32 matrices in HBM, consumed by the tensor cores as TF32
// (kind::tf32: 10-bit mantissa operands, fp32 accumulate in TMEM).  One CTA computes one
// 128 x BN tile:
//   warp 4      TMA producer   cp.async.bulk.tensor.2d, 128B-swizzled [rows x 32 fp32] boxes, kStages ring
//   warp 5      MMA issuer     one elected lane issues tcgen05.mma.cta_group::1.kind::tf32 (K=8 per instr),
//                              tcgen05.commit releases smem stages / signals the epilogue; owns the TMEM allocation
//   warps 0-3   epilogue       tcgen05.ld 32x32b.x32 (thread = one accumulator row), bias / beta, float4 stores
// Out-of-bounds rows / K columns are zero-filled by TMA, so M, N, K need no padding; leading
// dimensions must be multiples of 4 floats (16-byte TMA strides).
#pragma once
#include <cuda.h>

#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>

#include "common.cuh"

namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;          // fp32 elements per 128-byte swizzle row
constexpr int UMMA_K = 8;       // tf32: 32 bytes per instruction
constexpr int kStages = 6;
constexpr int kThreads = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a wrong descriptor must not hang the GPU -- trap instead (surfaces as a CUDA error).
// The suspend-time hint lets the hardware park the thread until the phase completes instead of re-polling: measured
// (tools/ubench/mbar.cu) a producer/consumer hand-off costs ~160 cycles with the hint vs ~300 without, because the
// re-polls of the waiting warp delay the other warp's barrier operations.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 14); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(1000000u)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled smem matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) (8 rows * 128 B = 1024)
// | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major tf32 operands: the only legal smem layout is SWIZZLE_128B_BASE32B (cutlass sm100_common.inl:92,
// UMMA::Layout_MN_SW128_32B_Atom = Swizzle<2,5,2> o (128 B of MN elements) x (4 k rows)): every k row holds
// 32 consecutive MN elements, 32-byte chunks are XOR-permuted by (row & 3), 4 k rows = one 512 B atom.
// LBO = byte stride between 32-element MN chunks, SBO = byte stride between 4-row k groups.
// The matching TMA mode is CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;   // SWIZZLE_128B_BASE32B
  return d;
}
// cute::UMMA::InstrDescriptor: c_format F32=1 [4,6) | a_format TF32=2 [7,10) | b_format TF32=2 [10,13)
// | a_major [15] b_major [16] (0 = K, 1 = MN) | N>>3 [17,23) | M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int mn_major = 0) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)mn_major << 15) | ((uint32_t)mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ================================================================================================
// Persistent warp-specialised GEMM (one CTA per SM, loops over work items = (m tile, n tile, k split)).
//   warp 0      TMA producer   every lane runs the loop and the mbarrier waits (warp-convergent control flow
//                              keeps descriptors in uniform registers), one elected lane issues
//   warp 1      MMA issuer     same; owns the 512-column TMEM allocation = two BN-wide accumulators, so the
//                              epilogue of work item i overlaps the K loop of item i+1
//   warps 2-5   epilogue       tcgen05.ld (thread = accumulator row) -> padded smem transpose -> row-contiguous
//                              128-byte global accesses (bias / beta*C / split-K partials)
// BN is a run-time multiple of 16 (TN: of 32) up to 256: measured on B200 (tools/ubench/tma_mma.cu) one
// tcgen05.mma kind::tf32 costs the issuing thread ~70 cycles and one smem-stage hand-off ~300 cycles whatever
// its size, so the K loop only approaches the tensor-pipe floor (BN/2 cycles per instruction) with wide tiles.
// ================================================================================================
struct TcGemmP {
  float* C; i64 ldc;
  int M, N, K;
  const float* bias; float beta;
  int BN, stages;              // tile width, smem ring depth
  int tiles_m, tiles_n, ksplit, kb_per_split;
  float* ws;                   // split-K partials [ksplit][M][N] (ksplit > 1), reduced by k_splitk_reduce
  int nkb0;                    // K-concatenated product C = A0 B0^T + A1 B1^T: k-blocks [0, nkb0) come from the first
  int nkb_total;               // operand pair, [nkb0, nkb_total) from the second (nkb0 == nkb_total: single pair)
};

constexpr int kGemmThreads = 192;
constexpr int kEpiPad = 33;    // floats per staged row: conflict-free column writes and row reads

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive_cta(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// ---- cta_group::2 (CTA pair) primitives: the pair computes a 256 x BN tile, each SM holds its 128 rows of A and HALF of B,
// so the operand ingest per SM and k-block drops from (128 + BN) to (128 + BN/2) rows -- measured 596 vs 790 cycles per
// 256-wide k-block (tools/ubench/mma2cta.cu).  Only the leader CTA (cluster rank 0) issues MMAs; both load and drain.
__device__ __forceinline__ uint32_t cluster_rank_() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all_() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load into this CTA's smem whose completion is signalled on the LEADER's mbarrier (peer bit of the address cleared)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// commit arriving on the mbarrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
// arrive on the barrier at this smem offset in the leader CTA (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t local_bar) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(0));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

template <bool TN, bool TWO>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
          const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b2, TcGemmP p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // TWO: this CTA stages its own 128 rows of A and BN/2 rows of B
  const int bn_local = TWO ? p.BN / 2 : p.BN;
  const uint32_t A_BYTES = BM * BK * 4, B_BYTES = (uint32_t)bn_local * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
  const uint32_t rank = TWO ? cluster_rank_() : 0u;
  const int tile_rows = TWO ? 2 * BM : BM;
  const int cta0 = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;          // work-item index / stride in CTAs or CTA pairs
  const int ncta = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  float* stage_c = reinterpret_cast<float*>(smem + (size_t)p.stages * STAGE_BYTES);      // [4 warps][32][kEpiPad]
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_c + 4 * 32 * kEpiPad);
  uint64_t* full_bar = bars;                      // [stages]
  uint64_t* empty_bar = bars + p.stages;          // [stages]
  uint64_t* acc_full = bars + 2 * p.stages;       // [2]
  uint64_t* acc_empty = acc_full + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(&acc_full[a]), 1); mbar_init(smem_u32(&acc_empty[a]), TWO ? 8 : 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if (!TWO) tmem_alloc(smem_u32(tmem_slot), 512);
    else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  fence_before_sync();
  __syncthreads();
  if (TWO) cluster_sync_all_();      // the peer's barriers are initialised before anything is signalled across the pair
  fence_after_sync();
  // warp-uniform by construction (REDUX writes a uniform register): keeps the tcgen05 operands out of vector registers
  const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, *tmem_slot);

  const int n_items = p.tiles_m * p.tiles_n * p.ksplit;
  const int num_kb_total = p.nkb_total;

  if (warp == 0) {
    // ================= TMA producer =================
    int it = 0;
    for (int item = cta0; item < n_items; item += ncta) {
      const int ks = item / (p.tiles_m * p.tiles_n);
      const int tile = item - ks * (p.tiles_m * p.tiles_n);
      const int m0 = (tile / p.tiles_n) * tile_rows + (int)rank * BM;              // this CTA's 128 rows of A
      const int n0 = (tile % p.tiles_n) * p.BN + (int)rank * bn_local;             // its rows of B
      const int kb0 = ks * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, num_kb_total);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
        if (elect_one()) {
          const uint32_t fb = smem_u32(&full_bar[s]);
          // TWO: both CTAs' loads complete on the leader's barrier, which therefore expects twice the bytes
          if (!TWO) mbar_expect_tx(fb, STAGE_BYTES);
          else if (rank == 0) mbar_expect_tx(fb, 2 * STAGE_BYTES);
          const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
          const bool second = kb >= p.nkb0;
          const CUtensorMap* ma = second ? &map_a2 : &map_a;
          const CUtensorMap* mb = second ? &map_b2 : &map_b;
          const int kc = (second ? kb - p.nkb0 : kb) * BK;
          if (!TN) {
            if (!TWO) { tma_load_2d(sa, ma, fb, kc, m0); tma_load_2d(sa + A_BYTES, mb, fb, kc, n0); }
            else { tma_load_2d_2sm(sa, ma, fb, kc, m0); tma_load_2d_2sm(sa + A_BYTES, mb, fb, kc, n0); }
          } else {
#pragma unroll
            for (int i = 0; i < BM / 32; ++i) {
              if (!TWO) tma_load_2d(sa + i * (BK * 128), ma, fb, m0 + i * 32, kc);
              else tma_load_2d_2sm(sa + i * (BK * 128), ma, fb, m0 + i * 32, kc);
            }
            for (int i = 0; i < bn_local / 32; ++i) {
              if (!TWO) tma_load_2d(sa + A_BYTES + i * (BK * 128), mb, fb, n0 + i * 32, kc);
              else tma_load_2d_2sm(sa + A_BYTES + i * (BK * 128), mb, fb, n0 + i * 32, kc);
            }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (TWO: the leader CTA only) =================
    const uint32_t idesc = make_idesc_tf32(tile_rows, p.BN, TN ? 1 : 0);
    int it = 0, n_done = 0;
    if (!TWO || rank == 0)
    for (int item = cta0; item < n_items; item += ncta, ++n_done) {
      const int ks = item / (p.tiles_m * p.tiles_n);
      const int kb0 = ks * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, num_kb_total);
      const int acc = n_done & 1;
      const uint32_t acc_ph = (n_done >> 1) & 1;
      mbar_wait(smem_u32(&acc_empty[acc]), acc_ph ^ 1);     // the epilogue drained this accumulator
      fence_after_sync();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 256);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        mbar_wait(smem_u32(&full_bar[s]), ph);
        fence_after_sync();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = TN ? make_smem_desc_mn(sa + k * 1024, BK * 128, 512) : make_smem_desc(sa + k * UMMA_K * 4);
            const uint64_t db = TN ? make_smem_desc_mn(sa + A_BYTES + k * 1024, BK * 128, 512)
                                   : make_smem_desc(sa + A_BYTES + k * UMMA_K * 4);
            if (!TWO) umma_tf32(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_tf32_2sm(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (!TWO) {
            umma_commit(smem_u32(&empty_bar[s]));             // frees this smem stage when the MMAs retire
            if (kb == kb1 - 1) umma_commit(smem_u32(&acc_full[acc]));
          } else {
            umma_commit_2sm(smem_u32(&empty_bar[s]));         // ... in both CTAs of the pair
            if (kb == kb1 - 1) umma_commit_2sm(smem_u32(&acc_full[acc]));
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue: warp w reads TMEM lanes [32*(w%4), +32) =================
    const int quad = warp & 3;
    float* st = stage_c + (size_t)(warp - 2) * 32 * kEpiPad;
    int n_done = 0;
    for (int item = cta0; item < n_items; item += ncta, ++n_done) {
      const int ks = item / (p.tiles_m * p.tiles_n);
      const int tile = item - ks * (p.tiles_m * p.tiles_n);
      const int m0 = (tile / p.tiles_n) * tile_rows + (int)rank * BM, n0 = (tile % p.tiles_n) * p.BN;   // own rows, all BN columns
      const int acc = n_done & 1;
      const uint32_t acc_ph = (n_done >> 1) & 1;
      mbar_wait(smem_u32(&acc_full[acc]), acc_ph);
      fence_after_sync();
      const int row0 = m0 + quad * 32;
      const int n_end = min(n0 + p.BN, p.N);
      float* outp; i64 ldo; const float* bias; float beta;
      if (p.ksplit > 1) { outp = p.ws + (size_t)ks * p.M * p.N; ldo = p.N; bias = nullptr; beta = 0.f; }
      else { outp = p.C; ldo = p.ldc; bias = p.bias; beta = p.beta; }
      const int rows_ok = min(32, p.M - row0);
      // TMEM (thread = row, 32 columns) -> swizzled float4 staging -> 16-byte global accesses, 4 rows x 128 B per warp
      // instruction.  Slot (j ^ (row & 7)) keeps both the row-wise writes and the 8-lanes-per-row reads conflict-free.
      float4* st4 = reinterpret_cast<float4*>(st);
      const bool vec_ok = ((ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(outp) & 15) == 0);
      const int rsub = lane >> 3, slot = lane & 7;
      for (int c0 = n0; c0 < n_end; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 256 + (c0 - n0)), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) st4[lane * 8 + (j ^ (lane & 7))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        const int col = c0 + 4 * slot;
        if (col < n_end) {
          float bv[4] = {0.f, 0.f, 0.f, 0.f};
          if (bias) {
#pragma unroll
            for (int e = 0; e < 4; ++e) if (col + e < n_end) bv[e] = bias[col + e];
          }
          const bool full4 = vec_ok && col + 4 <= n_end;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + rsub;
            if (r < rows_ok) {
              float4 o = st4[r * 8 + (slot ^ (r & 7))];
              o.x += bv[0]; o.y += bv[1]; o.z += bv[2]; o.w += bv[3];
              float* cp = outp + (i64)(row0 + r) * ldo + col;
              if (full4) {
                if (beta != 0.f) {
                  const float4 old = *reinterpret_cast<const float4*>(cp);
                  o.x += beta * old.x; o.y += beta * old.y; o.z += beta * old.z; o.w += beta * old.w;
                }
                *reinterpret_cast<float4*>(cp) = o;
              } else {
                const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (col + e < n_end) cp[e] = beta != 0.f ? ov[e] + beta * cp[e] : ov[e];
              }
            }
          }
        }
        __syncwarp();
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (!TWO) mbar_arrive_cta(smem_u32(&acc_empty[acc]));
        else mbar_arrive_leader(smem_u32(&acc_empty[acc]));     // the leader's MMA warp waits for both CTAs' epilogues
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (TWO) cluster_sync_all_();      // nobody frees TMEM / leaves while the peer may still signal or read across the pair
  if (warp == 1) {
    fence_after_sync();
    if (!TWO) tmem_dealloc(tmem_base, 512);
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// C[m,n] = sum_s ws[s][m][n] (+ bias[n]) (+ beta*C): fixed summation order -> deterministic split-K
__global__ void k_splitk_reduce(const float* __restrict__ ws, int ksplit, int M, int N, float* C, i64 ldc,
                                const float* __restrict__ bias, float beta) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (i64)M * N) return;
  const int m = (int)(i / N), n = (int)(i - (i64)m * N);
  float a = 0.f;
  for (int s = 0; s < ksplit; ++s) a += ws[(size_t)s * M * N + i];
  if (bias) a += bias[n];
  float* c = C + (i64)m * ldc + n;
  if (beta != 0.f) a += beta * (*c);
  *c = a;
}
// The same sum (same order, so bit-identical) with 16-byte accesses and four partial loads in flight; N % 4 == 0, 16-byte
// aligned rows of C.  unpermH > 0: the GEMM's columns are in the gate order of the persistent recurrent kernels
// (e2t_gate_perm) and C is the canonical tensor -- column n' = 64 (u/16) + 32 ((u%16)/8) + 8 g + u%8 lands at g H + u, so a
// weight gradient goes straight into the flat gradient buffer (runs of 8 columns stay contiguous: float4 stores).
__global__ void __launch_bounds__(256) k_splitk_reduce4(const float4* __restrict__ ws, int ksplit, int M, int N4, float* C,
                                                        i64 ldc, const float* __restrict__ bias, float beta, int unpermH) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const i64 total = (i64)M * N4;
  if (i >= total) return;
  const int m = (int)(i / N4), n = (int)(i - (i64)m * N4) * 4;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* w = ws + i;
  int s = 0;
  for (; s + 4 <= ksplit; s += 4) {
    const float4 v0 = w[(size_t)s * total], v1 = w[(size_t)(s + 1) * total];
    const float4 v2 = w[(size_t)(s + 2) * total], v3 = w[(size_t)(s + 3) * total];
    a.x += v0.x; a.y += v0.y; a.z += v0.z; a.w += v0.w;
    a.x += v1.x; a.y += v1.y; a.z += v1.z; a.w += v1.w;
    a.x += v2.x; a.y += v2.y; a.z += v2.z; a.w += v2.w;
    a.x += v3.x; a.y += v3.y; a.z += v3.z; a.w += v3.w;
  }
  for (; s < ksplit; ++s) {
    const float4 v = w[(size_t)s * total];
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  if (bias) { a.x += bias[n]; a.y += bias[n + 1]; a.z += bias[n + 2]; a.w += bias[n + 3]; }
  int nn = n;
  if (unpermH > 0) {
    const int r = n & 63;
    nn = ((r & 31) >> 3) * unpermH + (n >> 6) * 16 + (r >> 5) * 8 + (r & 7);
  }
  float4* c = reinterpret_cast<float4*>(C + (i64)m * ldc + nn);
  if (beta != 0.f) { const float4 o = *c; a.x += beta * o.x; a.y += beta * o.y; a.z += beta * o.z; a.w += beta * o.w; }
  *c = a;
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D fp32 tensor map over a row-major [rows, cols] matrix with leading dimension ld (elements);
// box = [box_rows x 32 cols], 128-byte swizzle, zero OOB fill.
inline CUtensorMap make_map(const float* ptr, i64 rows, i64 cols, i64 ld, int box_rows, bool atom32 = false) {
  if (rows <= 0 || cols <= 0) throw std::runtime_error("e2t: empty tensor map");
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  EncodeTiledFn fn = encode_fn();
  if (!fn) throw std::runtime_error("e2t: cuTensorMapEncodeTiled entry point not found");
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("e2t: cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return m;
}

inline int sm_count_() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

// Launches made while this is set are NOT persistent: one CTA per (tile, k-split) instead of min(work, SMs) CTAs looping over
// their share.  Used for the GEMMs that run on a low-priority side stream beside a persistent recurrent kernel (e2t.cu:
// SideScope): short-lived CTAs give the SMs back every few microseconds, so the high-priority kernel's CTAs are never held up
// for the duration of a whole GEMM.
inline bool& short_lived_ctas() { static thread_local bool v = false; return v; }

constexpr size_t kSmemCap = 227 * 1024;
inline size_t gemm_fixed_smem() { return (size_t)4 * 32 * kEpiPad * 4 + (2 * 8 + 4 + 1) * 8 + 1024; }
inline size_t gemm_smem_bytes(int BN, int stages) { return (size_t)stages * (BM * BK * 4 + (size_t)BN * BK * 4) + gemm_fixed_smem(); }

// split-K workspace (grown on demand; one per process is enough: launches are stream-ordered per handle, and
// handles on different streams would race -- so it is keyed by stream)
struct SplitWs { float* p = nullptr; size_t n = 0; };
inline SplitWs& split_ws(cudaStream_t st) {
  static std::mutex mu;
  static std::map<cudaStream_t, SplitWs> tab;
  std::lock_guard<std::mutex> g(mu);
  return tab[st];
}

// Tile width and split-K from a small cost model fitted to B200 measurements (tools/sweep_gemm.py): a k-block costs
// ~650 cycles up to BN = 128 (stage hand-off + MMA issue, not the tensor pipe) and ~1.1 more per extra column; a CTA pays
// ~6000 cycles of prologue / drain, a split-K reduction ~8000 cycles plus its traffic.
inline void pick_tiling(int M, int N, int num_kb, bool tn, int nsm, int* bn_out, int* ks_out, int* two_out) {
  double best = 1e30;
  int best_bn = 0, best_ks = 1, best_two = 0;
  static const bool allow_two = getenv("E2T_GEMM_NO_2SM") == nullptr;
  for (int two = 0; two <= ((allow_two && M > BM) ? 1 : 0); ++two) {
    // single CTA: BN multiple of 16 (TN: 32); CTA pair: each SM stages BN/2 rows of B -> multiples of 32 (TN: 64)
    const int step = (tn ? 32 : 16) * (two ? 2 : 1);
    const int tm = two ? (M + 2 * BM - 1) / (2 * BM) : (M + BM - 1) / BM;
    const int units = two ? nsm / 2 : nsm;
    const int n_cap = (N + step - 1) / step * step;
    for (int bn = std::min(256, n_cap); bn >= std::min(2 * step, n_cap); bn -= step) {
      const int tiles = tm * ((N + bn - 1) / bn);
      // cycles per 32-deep k-block (tools/sweep_gemm.py, tools/ubench/mma2cta.cu)
      const double per_kb = two ? 520.0 + std::max(0, bn - 128) * 0.6 : 650.0 + std::max(0, bn - 128) * 1.1;
      const int ks_max = tiles >= units ? 1 : std::max(1, std::min(8, num_kb / 8));
      for (int ks = 1; ks <= ks_max; ++ks) {
        const int kb = (num_kb + ks - 1) / ks;
        const double waves = std::ceil((double)tiles * ks / units);
        double cost = (two ? 7000.0 : 6000.0) + waves * (kb * per_kb + bn / 32.0 * 500.0);
        if (ks > 1) cost += 8000.0 + (double)ks * M * N * 4.0 / 3.0e12 * 1.965e9;
        if (cost < best - 1.0) { best = cost; best_bn = bn; best_ks = ks; best_two = two; }
      }
    }
  }
  *bn_out = best_bn; *ks_out = best_ks; *two_out = best_two;
}

// Optional second operand pair (A2, B2, K2): C = A B^T + A2 B2^T (+bias, +beta C) in ONE pass over C -- the two LSTM
// directions' contributions to d(input), or [x_t, h_{t-1}] [Wx; Wh] of a decoder step, without a read-modify-write of C.
// C_unperm / unpermH: the product's columns are in the recurrent kernels' gate order and `C_unperm` (leading dimension N) is
// the canonical tensor; when the launch splits K, its reduction writes there directly (returns true) and C is not touched;
// otherwise C holds the permuted product as usual (returns false) and the caller un-permutes.
template <bool TN>
inline bool launch_gemm(cudaStream_t st, const float* A, i64 lda, const float* B, i64 ldb, float* C, i64 ldc, int M, int N,
                        int K, const float* bias, float beta, const float* A2 = nullptr, i64 lda2 = 0,
                        const float* B2 = nullptr, i64 ldb2 = 0, int K2 = 0, float* C_unperm = nullptr, int unpermH = 0) {
  const int nsm = sm_count_();
  TcGemmP p{};
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.bias = bias; p.beta = beta;
  p.nkb0 = (K + BK - 1) / BK;
  const int num_kb = p.nkb0 + (A2 ? (K2 + BK - 1) / BK : 0);
  p.nkb_total = num_kb;
  int two = 0;
  pick_tiling(M, N, num_kb, TN, nsm, &p.BN, &p.ksplit, &two);
  p.tiles_m = two ? (M + 2 * BM - 1) / (2 * BM) : (M + BM - 1) / BM;
  p.tiles_n = (N + p.BN - 1) / p.BN;
  int tiles = p.tiles_m * p.tiles_n;
  {  // diagnostic overrides for shape sweeps (tools/sweep_gemm.py)
    const char* e_bn = getenv("E2T_GEMM_BN");
    const char* e_ks = getenv("E2T_GEMM_KSPLIT");
    if (e_bn && atoi(e_bn) > 0) {
      p.BN = atoi(e_bn);
      p.tiles_n = (N + p.BN - 1) / p.BN;
    }
    if (e_ks && atoi(e_ks) > 0) p.ksplit = std::min(atoi(e_ks), num_kb);
  }
  p.kb_per_split = (num_kb + p.ksplit - 1) / p.ksplit;
  p.ksplit = (num_kb + p.kb_per_split - 1) / p.kb_per_split;     // no empty splits
  const int bn_local = two ? p.BN / 2 : p.BN;      // B rows staged per SM
  int stages = 8;
  while (stages > 2 && gemm_smem_bytes(bn_local, stages) > kSmemCap) --stages;
  p.stages = stages;
  if (p.ksplit > 1) {
    SplitWs& w = split_ws(st);
    const size_t need = (size_t)p.ksplit * M * N;
    if (w.n < need) {
      if (w.p) { E2T_CHECK(cudaStreamSynchronize(st)); E2T_CHECK(cudaFree(w.p)); }
      E2T_CHECK(cudaMalloc(&w.p, need * sizeof(float)));
      w.n = need;
    }
    p.ws = w.p;
  }
  CUtensorMap ma, mb, ma2, mb2;
  if (!TN) { ma = make_map(A, M, K, lda, BM); mb = make_map(B, N, K, ldb, bn_local); }
  else { ma = make_map(A, K, M, lda, BK, true); mb = make_map(B, K, N, ldb, BK, true); }
  ma2 = ma; mb2 = mb;
  if (A2) {
    if (!TN) { ma2 = make_map(A2, M, K2, lda2, BM); mb2 = make_map(B2, N, K2, ldb2, bn_local); }
    else { ma2 = make_map(A2, K2, M, lda2, BK, true); mb2 = make_map(B2, K2, N, ldb2, BK, true); }
  }
  static bool attr_set = false;
  if (!attr_set) {
    E2T_CHECK(cudaFuncSetAttribute(k_gemm_tc<TN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCap));
    E2T_CHECK(cudaFuncSetAttribute(k_gemm_tc<TN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCap));
    attr_set = true;
  }
  tiles = p.tiles_m * p.tiles_n;
  const size_t smem_bytes = gemm_smem_bytes(bn_local, p.stages);
  if (!two) {
    const int grid = short_lived_ctas() ? tiles * p.ksplit : std::min(tiles * p.ksplit, nsm);
    k_gemm_tc<TN, false><<<grid, kGemmThreads, smem_bytes, st>>>(ma, mb, ma2, mb2, p);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * (short_lived_ctas() ? tiles * p.ksplit : std::min(tiles * p.ksplit, nsm / 2))));
    cfg.blockDim = dim3(kGemmThreads); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    E2T_CHECK(cudaLaunchKernelEx(&cfg, k_gemm_tc<TN, true>, ma, mb, ma2, mb2, p));
  }
  if (p.ksplit > 1) {
    const i64 n = (i64)M * N;
    static const bool scalar_reduce = getenv("E2T_SPLITK_SCALAR") != nullptr;
    const bool unperm = C_unperm != nullptr && unpermH > 0 && (unpermH & 3) == 0 && (N & 63) == 0 && N == 4 * unpermH &&
                        (reinterpret_cast<uintptr_t>(C_unperm) & 15) == 0 && !scalar_reduce;
    float* Co = unperm ? C_unperm : C;
    const i64 ldo = unperm ? (i64)N : ldc;
    if (!scalar_reduce && (N & 3) == 0 && (ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(Co) & 15) == 0)
      k_splitk_reduce4<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(p.ws), p.ksplit, M, N / 4,
                                                                         Co, ldo, bias, beta, unperm ? unpermH : 0);
    else
      k_splitk_reduce<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p.ws, p.ksplit, M, N, C, ldc, bias, beta);
    return unperm;
  }
  return false;
}

}  // namespace tc

// ---- interface used by e2t.cu -----------------------------------------------------------------------
static inline bool tc_gemm_nt_supported(const float* A, i64 lda, const float* B, i64 ldb, const float* C, i64 ldc, int M,
                                        int N, int K) {
  (void)C; (void)ldc;
  if (M < 1 || N < 8 || K < 8) return false;
  if ((lda & 3) || (ldb & 3)) return false;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return false;
  // tiny problems are launch-bound either way; keep them on the fp32 path (also keeps tiny parity tests exact)
  if ((i64)M * N * K < (i64)64 * 64 * 64) return false;
  return true;
}

static inline void tc_gemm_nt(cudaStream_t st, const float* A, i64 lda, const float* B, i64 ldb, float* C, i64 ldc, int M,
                              int N, int K, const float* bias, float beta) {
  tc::launch_gemm<false>(st, A, lda, B, ldb, C, ldc, M, N, K, bias, beta);
}
// C = A B^T + A2 B2^T (+bias) (+beta C): all four operands K-major
static inline void tc_gemm_nt2(cudaStream_t st, const float* A, i64 lda, const float* B, i64 ldb, int K, const float* A2,
                               i64 lda2, const float* B2, i64 ldb2, int K2, float* C, i64 ldc, int M, int N,
                               const float* bias, float beta) {
  tc::launch_gemm<false>(st, A, lda, B, ldb, C, ldc, M, N, K, bias, beta, A2, lda2, B2, ldb2, K2);
}

// C[M,N] = A[K,M]^T B[K,N]  (A, B row-major with leading dims lda, ldb): weight gradients.
static inline bool tc_gemm_tn_supported(const float* A, i64 lda, const float* B, i64 ldb, int M, int N, int K) {
  if (M < 8 || N < 8 || K < 32) return false;
  if ((lda & 3) || (ldb & 3)) return false;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return false;
  if ((i64)M * N * K < (i64)64 * 64 * 64) return false;
  return true;
}
static inline bool tc_gemm_tn(cudaStream_t st, const float* A, i64 lda, const float* B, i64 ldb, float* C, i64 ldc, int M,
                              int N, int K, const float* bias, float beta, float* C_unperm = nullptr, int unpermH = 0) {
  return tc::launch_gemm<true>(st, A, lda, B, ldb, C, ldc, M, N, K, bias, beta, nullptr, 0, nullptr, 0, 0, C_unperm, unpermH);
}

// times `iters` back-to-back launches of one GEMM shape on zero-filled operands (diagnostic; e2t_bench_gemm)
// uniform(-0.5, 0.5) operands for the GEMM timing (all-zero operands do not toggle the datapath: the tensor pipe then runs
// at a power / clock point no real GEMM sees, which is how an 8192^3 "1033 TFLOP/s" was once measured)
__global__ void k_fill_uniform(float* p, size_t n, uint32_t seed) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x = (uint32_t)i * 2654435761u + seed;
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  p[i] = (float)(x >> 8) * (1.0f / 16777216.0f) - 0.5f;
}

static inline float tc_gemm_bench(cudaStream_t st, int M, int N, int K, bool tn, float beta, int iters) {
  const i64 lda = tn ? (M + 3) / 4 * 4 : (K + 3) / 4 * 4, ldb = tn ? (N + 3) / 4 * 4 : (K + 3) / 4 * 4, ldc = (N + 3) / 4 * 4;
  float *dA, *dB, *dC;
  const size_t na = (size_t)(tn ? K : M) * lda, nb = (size_t)(tn ? K : N) * ldb, nc = (size_t)M * ldc;
  E2T_CHECK(cudaMalloc(&dA, na * 4)); E2T_CHECK(cudaMalloc(&dB, nb * 4)); E2T_CHECK(cudaMalloc(&dC, nc * 4));
  if (getenv("E2T_BENCH_ZERO_OPERANDS")) {
    E2T_CHECK(cudaMemsetAsync(dA, 0, na * 4, st)); E2T_CHECK(cudaMemsetAsync(dB, 0, nb * 4, st));
  } else {
    k_fill_uniform<<<(unsigned)((na + 255) / 256), 256, 0, st>>>(dA, na, 1u);
    k_fill_uniform<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(dB, nb, 2u);
  }
  E2T_CHECK(cudaMemsetAsync(dC, 0, nc * 4, st));
  cudaEvent_t e0, e1;
  E2T_CHECK(cudaEventCreate(&e0)); E2T_CHECK(cudaEventCreate(&e1));
  for (int i = 0; i < iters + 2; ++i) {
    if (i == 2) E2T_CHECK(cudaEventRecord(e0, st));
    if (tn) tc_gemm_tn(st, dA, lda, dB, ldb, dC, ldc, M, N, K, nullptr, beta);
    else tc_gemm_nt(st, dA, lda, dB, ldb, dC, ldc, M, N, K, nullptr, beta);
  }
  E2T_CHECK(cudaEventRecord(e1, st));
  E2T_CHECK(cudaStreamSynchronize(st));
  E2T_CHECK(cudaGetLastError());
  float ms = 0.f;
  E2T_CHECK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  return ms / iters;
}

// C[M,N] = A[K,M]^T B[K,N] + A2[K2,M]^T B2[K2,N]: two row ranges of a weight gradient in one pass
static inline void tc_gemm_tn2(cudaStream_t st, const float* A, i64 lda, const float* B, i64 ldb, int K, const float* A2,
                               i64 lda2, const float* B2, i64 ldb2, int K2, float* C, i64 ldc, int M, int N) {
  tc::launch_gemm<true>(st, A, lda, B, ldb, C, ldc, M, N, K, nullptr, 0.f, A2, lda2, B2, ldb2, K2);
}

// A/B random, C_tc vs fp32 SIMT reference; exercises bias, beta and ragged M/N/K edges. Returns max |diff|.
static inline float tc_gemm_selftest(cudaStream_t st, int M, int N, int K, bool tn = false) {
  const i64 lda = tn ? (M + 3) / 4 * 4 : (K + 3) / 4 * 4, ldb = tn ? (N + 3) / 4 * 4 : (K + 3) / 4 * 4, ldc = (N + 3) / 4 * 4;
  std::vector<float> hA((size_t)(tn ? K : M) * lda), hB((size_t)(tn ? K : N) * ldb), hC((size_t)M * ldc), hbias(N);
  uint32_t s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
  for (auto& v : hA) v = rnd();
  for (auto& v : hB) v = rnd();
  for (auto& v : hC) v = rnd();
  for (auto& v : hbias) v = rnd();
  float *dA, *dB, *dC1, *dC2, *dbias;
  E2T_CHECK(cudaMalloc(&dA, hA.size() * 4)); E2T_CHECK(cudaMalloc(&dB, hB.size() * 4));
  E2T_CHECK(cudaMalloc(&dC1, hC.size() * 4)); E2T_CHECK(cudaMalloc(&dC2, hC.size() * 4));
  E2T_CHECK(cudaMalloc(&dbias, hbias.size() * 4));
  E2T_CHECK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  E2T_CHECK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  E2T_CHECK(cudaMemcpy(dC1, hC.data(), hC.size() * 4, cudaMemcpyHostToDevice));
  E2T_CHECK(cudaMemcpy(dC2, hC.data(), hC.size() * 4, cudaMemcpyHostToDevice));
  E2T_CHECK(cudaMemcpy(dbias, hbias.data(), hbias.size() * 4, cudaMemcpyHostToDevice));
  if (tn) tc_gemm_tn(st, dA, lda, dB, ldb, dC1, ldc, M, N, K, dbias, 1.0f);
  else tc_gemm_nt(st, dA, lda, dB, ldb, dC1, ldc, M, N, K, dbias, 1.0f);
  GemmP g{};
  g.A = dA; g.sam = tn ? 1 : lda; g.sak = tn ? lda : 1; g.B = dB; g.sbk = tn ? ldb : 1; g.sbn = tn ? 1 : ldb;
  g.C = dC2; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K; g.bias = dbias; g.beta = 1.0f;
  dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64));
  k_gemm<0><<<grid, 256, 0, st>>>(g);
  E2T_CHECK(cudaStreamSynchronize(st));
  E2T_CHECK(cudaGetLastError());
  std::vector<float> c1(hC.size()), c2(hC.size());
  E2T_CHECK(cudaMemcpy(c1.data(), dC1, hC.size() * 4, cudaMemcpyDeviceToHost));
  E2T_CHECK(cudaMemcpy(c2.data(), dC2, hC.size() * 4, cudaMemcpyDeviceToHost));
  float md = 0.f;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) md = fmaxf(md, fabsf(c1[(size_t)m * ldc + n] - c2[(size_t)m * ldc + n]));
  cudaFree(dA); cudaFree(dB); cudaFree(dC1); cudaFree(dC2); cudaFree(dbias);
  return md;
}
