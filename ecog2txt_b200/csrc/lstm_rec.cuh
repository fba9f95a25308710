// lstm_rec.cuh -- persistent recurrent LSTM kernels for sm_100a (A5 / A10 of SURVEY.md section 8a).
//
// One cooperative launch runs ALL time steps of BOTH directions of one BiLSTM layer (forward pass) or
// of its BPTT (backward pass).  The recurrent weight slice of every CTA is staged ONCE by TMA into
// shared memory and stays there for the whole sequence; per step only the previous hidden state
// (forward) / the previous dz (backward) is streamed through a TMA ring, multiplied on the tcgen05
// tensor cores (kind::tf32, fp32 accumulate in TMEM) and the gate non-linearities + cell update
// (forward) or the gate derivatives (backward) are applied straight out of tcgen05.ld registers.
//
// Decomposition (grid = 2 directions x batch tiles of 128 rows x H/16 unit slices):
//   forward : CTA owns hidden units [16j, 16j+16) -> its 64 gate columns (i|j|f|o x 16 units), K = H.
//             D[128 x 64] = h_prev[128 x H] * WhT_slice[64 x H]^T ; epilogue thread = one batch row,
//             keeps its 16 cell values in registers across all steps.
//   backward: k_lstm_bptt (below), reduce-scatter: every CTA multiplies only its own 64 dz columns for all H units and the
//             partial dh goes through an L2 workspace.  The older all-gather form k_lstm_rec<true> (N = 16 columns of
//             dh_rec, K = 4H: 819 KB of dz per CTA per step) is kept as the fallback for H > 512 only.
// CTAs of one (direction, batch tile) chain exchange h / partial dh through L2 and a per-step arrival counter:
// epilogue stores -> fence -> red.release ; waiter ld.relaxed spin -> fence.acq_rel (-> fence.proxy.async -> TMA).
// All CTAs must be co-resident: cooperative launch, grid <= #SMs.
//
// warp roles (k_lstm_rec): 0 = TMA producer (+ counter wait), 1 = MMA issuer (+ TMEM owner), 2..9 = epilogue.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "gemm_tc.cuh"

namespace rec {

using namespace tc;

constexpr int kU = 16;            // hidden units per CTA
constexpr int kBM = 128;          // batch rows per CTA
constexpr int kUT = E2T_REC_UT;   // hidden units per epilogue thread (8: every global access is a full 32 B sector)
constexpr int kEpiWarps = 4 * (kU / kUT);          // kU/kUT warps per TMEM lane quadrant
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kThreadsRec = 64 + kEpiThreads;      // + TMA producer warp + MMA warp
constexpr uint32_t A_STAGE_BYTES = kBM * BK * 4;   // 16 KB
constexpr int kGroup = 3;                          // k-chunks per mbarrier phase of the A-tile ring

__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void red_release_gpu(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
// publish the step: make the epilogue's generic-proxy global stores visible (gpu scope, async proxy) and bump the counter
__device__ __forceinline__ void publish_step(int* ctr, int mode) {
  if (mode == 0) { fence_proxy_async_all(); red_release_gpu(ctr, 1); }
  else if (mode == 1) { red_release_gpu(ctr, 1); }
  else if (mode == 2) { __threadfence(); atomicAdd(ctr, 1); }
  else { fence_proxy_async_global(); red_release_gpu(ctr, 1); }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// Bounded wait on a global arrival counter (a lost arrival must trap, not hang the GPU).
__device__ __forceinline__ void wait_counter(const int* p, int target) {
  // relaxed polling (an acquire load in the loop costs one L1 invalidation per iteration), one acquire fence after
  const long long t0 = clock64();
  while (ld_relaxed_gpu(p) < target) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
  fence_acq_rel_gpu();
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// n contiguous floats (n % 4 == 0, 16-byte aligned) <-> registers
template <int N>
__device__ __forceinline__ void ldv(float* dst, const float* src) {
#pragma unroll
  for (int i = 0; i < N / 4; ++i) {
    const float4 v = *reinterpret_cast<const float4*>(src + 4 * i);
    dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
  }
}
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): a thread's 32 contiguous bytes are one full sector per request,
// half the L1 wavefronts of two 128-bit accesses.  n % 8 == 0, 32-byte aligned.
template <int N>
__device__ __forceinline__ void ldv8(float* d, const float* src) {
#pragma unroll
  for (int i = 0; i < N / 8; ++i)
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(d[8 * i]), "=f"(d[8 * i + 1]), "=f"(d[8 * i + 2]), "=f"(d[8 * i + 3]), "=f"(d[8 * i + 4]),
                   "=f"(d[8 * i + 5]), "=f"(d[8 * i + 6]), "=f"(d[8 * i + 7])
                 : "l"(src + 8 * i));
}
template <int N>
__device__ __forceinline__ void stv8(float* dst, const float* v) {
#pragma unroll
  for (int i = 0; i < N / 8; ++i)
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + 8 * i), "f"(v[8 * i]), "f"(v[8 * i + 1]),
                 "f"(v[8 * i + 2]), "f"(v[8 * i + 3]), "f"(v[8 * i + 4]), "f"(v[8 * i + 5]), "f"(v[8 * i + 6]), "f"(v[8 * i + 7])
                 : "memory");
}
template <int N>
__device__ __forceinline__ void ldv_cg(float* dst, const float* src) {
#pragma unroll
  for (int i = 0; i < N / 4; ++i) {
    const float4 v = __ldcg(reinterpret_cast<const float4*>(src + 4 * i));
    dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
  }
}
template <int N>
__device__ __forceinline__ void stv(float* dst, const float* src) {
#pragma unroll
  for (int i = 0; i < N / 4; ++i)
    *reinterpret_cast<float4*>(dst + 4 * i) = make_float4(src[4 * i], src[4 * i + 1], src[4 * i + 2], src[4 * i + 3]);
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
template <int NCOL>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* v);
template <>
__device__ __forceinline__ void tmem_ld_cols<4>(uint32_t taddr, float* v) { tmem_ld4(taddr, v); tmem_ld_wait(); }
template <>
__device__ __forceinline__ void tmem_ld_cols<8>(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
  tmem_ld_wait();
}
template <>
__device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, float* v) { tmem_ld32(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <>
__device__ __forceinline__ void tmem_ld_cols<64>(uint32_t taddr, float* v) {
  tmem_ld32(taddr, v);
  tmem_ld32(taddr + 32, v + 32);
}

// fp32 gate non-linearities on the SFU: ex2.approx + rcp (abs error ~1e-7, far below the tf32 operand rounding)
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigm(float x) { return rcp_approx(1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) {
  // tanh(x) = 1 - 2 / (1 + e^{2x}); saturates cleanly (e^{2x} -> inf gives 1, -> 0 gives -1)
  return 1.0f - 2.0f * rcp_approx(1.0f + __expf(2.0f * x));
}

struct RecFwdP {
  float* gates[2];        // [T', B, 4H] x-projection (+bias) in, gate activations out
  float* cs[2];           // [T', B, H]
  float* hs;              // [T', B, 2H]   fwd half | bwd half
  float* hd;              // dropped copy of hs (nullable)
  const int* lens2;       // [B] (nullable: all steps valid)
  int* counters;          // [2][n_bt][steps], zeroed before launch
  int steps, B, H, n_bt, n_slices, nkc, stages;
  DropP dp; int drop_F;
  long long* dbg;         // E2T_REC_DEBUG=1: per-step clock64 stamps of CTA 0 ([steps][8]), else NULL
  int pub_mode;
};

struct RecBwdP {
  float* gates[2];        // gate activations in, dz out (in place)
  const float* cs[2];
  const float* dhs;       // [T', B, 2H] grad wrt layer outputs (nullable)
  const int* lens2;
  const float* dc_inject; int ldi;       // [B, ldi] final-state cell grad (nullable), column offset d*H
  const int* inject_t;                   // fwd direction: time index per row at which to inject (nullable -> 0)
  int* counters;
  int steps, B, H, n_bt, n_slices, nkc, stages;
  long long* dbg;
  int pub_mode;
};

// shared-memory carve (dynamic, 1024-aligned): [W resident: nkc * WCHUNK] [A ring: stages * 16 KB] [barriers]
template <bool BWD>
struct Geo {
  static constexpr int N = BWD ? kU : 4 * kU;                   // MMA N = accumulator columns
  static constexpr uint32_t WCHUNK = (uint32_t)N * BK * 4;      // bytes of one resident weight k-chunk
  static constexpr uint32_t TMEM_COLS = BWD ? 32 : 64;
};

template <bool BWD, typename P>
__global__ void __launch_bounds__(kThreadsRec, 1)
k_lstm_rec(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
           const __grid_constant__ CUtensorMap map_w0, const __grid_constant__ CUtensorMap map_w1, P p) {
  using G = Geo<BWD>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* smem_w = smem;
  unsigned char* smem_a = smem + (size_t)p.nkc * G::WCHUNK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + (size_t)p.stages * kGroup * A_STAGE_BYTES);
  uint64_t* full_bar = bars;                       // [stages]
  uint64_t* empty_bar = bars + p.stages;           // [stages]
  uint64_t* w_bar = bars + 2 * p.stages;
  uint64_t* acc_full = w_bar + 1;
  uint64_t* acc_empty = w_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blockIdx.x = ((d * n_bt) + bt) * n_slices + j
  const int j = blockIdx.x % p.n_slices;
  const int bt = (blockIdx.x / p.n_slices) % p.n_bt;
  const int d = blockIdx.x / (p.n_slices * p.n_bt);
  const bool reverse = d == 1;
  const CUtensorMap* map_a = d ? &map_a1 : &map_a0;
  const CUtensorMap* map_w = d ? &map_w1 : &map_w0;
  const int steps = p.steps, B = p.B, H = p.H;
  int* counters = p.counters + (size_t)(d * p.n_bt + bt) * steps;
  // per-CTA starting K chunk (the K order is free; spreads the chain's CTAs over the tile)
  const int kc_rot = (int)(((long long)j * p.nkc) / p.n_slices);
  long long* dbg = (blockIdx.x == 0) ? p.dbg : nullptr;
  // every CTA also stamps one probe step with the device-wide timer ([steps*8 + blockIdx*8 + slot]): who is late?
  long long* dbgx = p.dbg ? p.dbg + (size_t)p.steps * 8 + (size_t)blockIdx.x * 8 : nullptr;
#define REC_STAMP(step, slot) do { if (dbg) dbg[(step) * 8 + (slot)] = clock64(); \
    if (dbgx && (step) == 10) { unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); dbgx[slot] = (long long)gt_; } } while (0)

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(w_bar), 1);
    mbar_init(smem_u32(acc_full), 1);
    mbar_init(smem_u32(acc_empty), kEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), G::TMEM_COLS);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, *tmem_slot);   // warp-uniform: tcgen05 operands stay in uniform registers

  if (warp == 0) {
    // ================= TMA producer (warp-convergent: every lane polls / waits, one elected lane issues) =========
    if (elect_one()) {
      // resident weights, once
      const uint32_t wb = smem_u32(w_bar);
      mbar_expect_tx(wb, (uint32_t)p.nkc * G::WCHUNK);
      for (int kc = 0; kc < p.nkc; ++kc) {
        // forward: 64 permuted gate rows of Wh^T ; backward: 16 unit rows of Wh (permuted gate columns = K)
        tma_load_2d(smem_u32(smem_w + (size_t)kc * G::WCHUNK), map_w, wb, kc * BK, j * G::N);
      }
    }
    __syncwarp();
    int it = 0;
    for (int s = 1; s < steps; ++s) {
      // forward pass : step s (forward order) consumes h of step s-1
      // backward pass: q-th processed step (q = s) consumes dz of the (q-1)-th processed step
      int t_src;
      if (!BWD) { const int t = reverse ? steps - 1 - s : s; t_src = reverse ? t + 1 : t - 1; }
      else      { const int sf = steps - 1 - s; const int t = reverse ? steps - 1 - sf : sf; t_src = reverse ? t - 1 : t + 1; }
      wait_counter(counters + (s - 1), p.n_slices);     // all lanes poll the same word: one L2 request per round
      fence_proxy_async_all();
      if (lane == 0) REC_STAMP(s, 0);
      // the A tile arrives in groups of kGroup k-chunks per mbarrier phase: one producer/consumer hand-off costs about as
      // much as the four MMAs of a chunk (tools/ubench/tma_mma.cu), so fewer, larger phases shorten the step
      for (int kk0 = 0; kk0 < p.nkc; kk0 += kGroup, ++it) {
        const int st = it % p.stages;                 // ring slot = group of kGroup chunk buffers
        const uint32_t ph = (it / p.stages) & 1;
        const int n = min(kGroup, p.nkc - kk0);
        mbar_wait(smem_u32(&empty_bar[st]), ph ^ 1);
        if (elect_one()) {
          const uint32_t fb = smem_u32(&full_bar[st]);
          mbar_expect_tx(fb, (uint32_t)n * A_STAGE_BYTES);
          for (int i = 0; i < n; ++i) {
            int kc = kk0 + i + kc_rot;                // the K order is free: spread the chain's CTAs over the tile
            if (kc >= p.nkc) kc -= p.nkc;
            tma_load_2d(smem_u32(smem_a + (size_t)(st * kGroup + i) * A_STAGE_BYTES), map_a, fb, kc * BK, t_src * B + bt * kBM);
          }
        }
        __syncwarp();
      }
      if (lane == 0) REC_STAMP(s, 1);
    }
  } else if (warp == 1) {
    // ================= MMA issuer (warp-convergent) =================
    constexpr uint32_t idesc = make_idesc_tf32(kBM, G::N, 0);
    mbar_wait(smem_u32(w_bar), 0);
    fence_after_sync();
    // K-major SWIZZLE_128B descriptors advance linearly with the byte offset (>> 4) as long as bits [0,14) do not wrap
    const uint64_t desc_a0 = make_smem_desc(smem_u32(smem_a));
    const uint64_t desc_w0 = make_smem_desc(smem_u32(smem_w));
    int it = 0;
    for (int s = 1; s < steps; ++s) {
      mbar_wait(smem_u32(acc_empty), ((s - 1) & 1) ^ 1);   // epilogue drained the previous accumulator
      fence_after_sync();
      for (int kk0 = 0; kk0 < p.nkc; kk0 += kGroup, ++it) {
        const int st = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        const int n = min(kGroup, p.nkc - kk0);
        mbar_wait(smem_u32(&full_bar[st]), ph);
        fence_after_sync();
        if (kk0 == 0 && lane == 0) REC_STAMP(s, 2);
        if (elect_one()) {
          for (int i = 0; i < n; ++i) {
            int kc = kk0 + i + kc_rot;
            if (kc >= p.nkc) kc -= p.nkc;
            const uint64_t da = desc_a0 + (uint64_t)(((uint32_t)(st * kGroup + i) * A_STAGE_BYTES) >> 4);
            const uint64_t dw = desc_w0 + (uint64_t)(((uint32_t)kc * G::WCHUNK) >> 4);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_tf32(tmem_base, da + (uint64_t)(k * UMMA_K * 4 >> 4), dw + (uint64_t)(k * UMMA_K * 4 >> 4), idesc,
                        (kk0 + i > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(smem_u32(&empty_bar[st]));
          if (kk0 + kGroup >= p.nkc) umma_commit(smem_u32(acc_full));
        }
        __syncwarp();
      }
      if (lane == 0) REC_STAMP(s, 3);
    }
  } else {
    // ================= epilogue: thread = (batch row, kUT hidden units) =================
    // warp w may only read TMEM lanes [32*(w%4), +32): kU/kUT warps per lane quadrant, one per group of kUT units,
    // so that every scheduler holds several epilogue warps and the gate math (latency-bound in one warp) overlaps.
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int ug = ew >> 2;                          // unit group
    const int r = quad * 32 + lane;
    const int b = bt * kBM + r;
    const bool row_ok = b < B;
    const int u0 = j * kU + ug * kUT;                // first hidden unit of this thread
    const int z0 = j * 4 * kU + ug * 4 * kUT;        // its 4*kUT contiguous gate columns (permuted layout [ug][gate][kUT])
    const int len2 = row_ok ? (p.lens2 ? p.lens2[b] : steps) : 0;
    // TMEM columns follow the B-operand rows: forward [ug][gate][kUT] (4*kUT per thread), backward units (kUT per thread)
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(BWD ? ug * kUT : ug * 4 * kUT);
    float carry[kUT];                                // forward: cell state ; backward: dc through time
#pragma unroll
    for (int i = 0; i < kUT; ++i) carry[i] = 0.f;
    const bool publisher = threadIdx.x == 64;

    if constexpr (!BWD) {
      float* gates = d ? p.gates[1] : p.gates[0];
      float* cs = d ? p.cs[1] : p.cs[0];
      const int col0 = d * H;
      for (int s = 0; s < steps; ++s) {
        const int t = reverse ? steps - 1 - s : s;
        const bool valid = row_ok && t < len2;
        float* zrow = gates + ((i64)t * B + b) * 4 * H + z0;   // [gate][kUT] contiguous
        float z[4 * kUT];
        if (valid) ldv8<4 * kUT>(z, zrow);
        float acc[4 * kUT];
        if (s > 0) {
          mbar_wait(smem_u32(acc_full), (s - 1) & 1);
          fence_after_sync();
          if (publisher) REC_STAMP(s, 4);
          tmem_ld_cols<4 * kUT>(taddr, acc);
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(acc_empty));
        } else {
#pragma unroll
          for (int i = 0; i < 4 * kUT; ++i) acc[i] = 0.f;
        }
        float hv[kUT];
        if (valid) {
#pragma unroll
          for (int e = 0; e < kUT; ++e) {
            const float gi = sigm(z[e] + acc[e]);
            const float gj = tanh_fast(z[kUT + e] + acc[kUT + e]);
            const float gf = sigm(z[2 * kUT + e] + acc[2 * kUT + e] + 1.0f);
            const float go = sigm(z[3 * kUT + e] + acc[3 * kUT + e]);
            const float c = gf * carry[e] + gi * gj;
            carry[e] = c;
            hv[e] = go * tanh_fast(c);
            z[e] = gi; z[kUT + e] = gj; z[2 * kUT + e] = gf; z[3 * kUT + e] = go;
          }
        } else {
#pragma unroll
          for (int e = 0; e < kUT; ++e) { carry[e] = 0.f; hv[e] = 0.f; }
        }
        // 1) the hidden state is all the other CTAs of the chain wait for: store it, then publish
        if (row_ok) stv8<kUT>(p.hs + ((i64)t * B + b) * 2 * H + col0 + u0, hv);
        if (publisher) REC_STAMP(s, 5);
        named_bar_sync(1, kEpiThreads);
        if (publisher) {
          REC_STAMP(s, 6);
          // generic-proxy stores of the epilogue warps (ordered by the barrier) -> gpu scope -> async proxy (TMA)
          publish_step(counters + s, p.pub_mode & 15);
          REC_STAMP(s, 7);
        }
        // 2) everything only the backward pass needs (gate activations, cell state, dropout copy) goes out after
        //    the release, off the inter-CTA critical path -- and only once the release has been issued: a gpu-scope
        //    release waits for the SM's outstanding stores, so 48 KB of them issued meanwhile would sit in front of it
        named_bar_sync(3, kEpiThreads);
        if (row_ok) {
          if (valid) stv8<4 * kUT>(zrow, z);
          stv8<kUT>(cs + ((i64)t * B + b) * H + u0, carry);
          if (p.hd) {
            const uint32_t idx0 = (uint32_t)(((i64)t * B + b) * p.drop_F + col0 + u0);
            const uint32_t key = p.dp.key, thresh = p.dp.thresh;
            const float inv = p.dp.inv;
            float o[kUT];
#pragma unroll
            for (int e = 0; e < kUT; ++e) o[e] = (valid && e2t_keep(key, idx0 + e, thresh)) ? hv[e] * inv : 0.f;
            stv8<kUT>(p.hd + ((i64)t * B + b) * 2 * H + col0 + u0, o);
          }
        }
      }
    } else {
      float* gates = d ? p.gates[1] : p.gates[0];
      const float* cs = d ? p.cs[1] : p.cs[0];
      const int col0 = d * H;
      for (int q = 0; q < steps; ++q) {
        const int sf = steps - 1 - q;                       // forward-order step index being differentiated
        const int t = reverse ? steps - 1 - sf : sf;
        const int tp = reverse ? t + 1 : t - 1;             // step processed before t in the forward pass
        const bool valid = row_ok && t < len2;
        float* zrow = gates + ((i64)t * B + b) * 4 * H + z0;
        float gz[4 * kUT], cv[kUT], cpv[kUT], dhv[kUT];
        if (valid) {
          ldv<4 * kUT>(gz, zrow);
          ldv<kUT>(cv, cs + ((i64)t * B + b) * H + u0);
          if (sf > 0) ldv<kUT>(cpv, cs + ((i64)tp * B + b) * H + u0);
          else {
#pragma unroll
            for (int e = 0; e < kUT; ++e) cpv[e] = 0.f;
          }
          if (p.dhs) ldv<kUT>(dhv, p.dhs + ((i64)t * B + b) * 2 * H + col0 + u0);
          else {
#pragma unroll
            for (int e = 0; e < kUT; ++e) dhv[e] = 0.f;
          }
        }
        float acc[kUT];
        if (q > 0) {
          mbar_wait(smem_u32(acc_full), (q - 1) & 1);
          fence_after_sync();
          if (publisher) REC_STAMP(q, 4);
          tmem_ld_cols<kUT>(taddr, acc);
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(acc_empty));
        } else {
#pragma unroll
          for (int i = 0; i < kUT; ++i) acc[i] = 0.f;
        }
        if (row_ok) {
          if (valid) {
            bool inject = false;
            if (p.dc_inject) {
              const int ti = (d == 0 && p.inject_t) ? p.inject_t[b] : 0;
              inject = ti == t;
            }
#pragma unroll
            for (int e = 0; e < kUT; ++e) {
              const float gi = gz[e], gj = gz[kUT + e], gf = gz[2 * kUT + e], go = gz[3 * kUT + e];
              const float dh = dhv[e] + acc[e];
              float dc = carry[e];
              if (inject) dc += p.dc_inject[(i64)b * p.ldi + col0 + u0 + e];
              const float tc_ = tanh_fast(cv[e]);
              gz[3 * kUT + e] = dh * tc_ * go * (1.f - go);
              dc += dh * go * (1.f - tc_ * tc_);
              gz[e] = dc * gj * gi * (1.f - gi);
              gz[kUT + e] = dc * gi * (1.f - gj * gj);
              gz[2 * kUT + e] = dc * cpv[e] * gf * (1.f - gf);
              carry[e] = dc * gf;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 4 * kUT; ++i) gz[i] = 0.f;
#pragma unroll
            for (int e = 0; e < kUT; ++e) carry[e] = 0.f;
          }
          stv<4 * kUT>(zrow, gz);
        }
        if (publisher) REC_STAMP(q, 5);
        named_bar_sync(1, kEpiThreads);
        if (publisher) {
          REC_STAMP(q, 6);
          publish_step(counters + q, p.pub_mode & 15);
          REC_STAMP(q, 7);
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, G::TMEM_COLS);
  }
}

// ================================================================================================
// BPTT, reduce-scatter formulation (k_lstm_bptt).  The all-gather kernel above must stream dz_{t+1}[128 x 4H] into every
// CTA each step (819 KB, 200 tcgen05.mma of K = 8: ~23 us per step at H = 400).  Here every CTA multiplies ONLY ITS OWN
// 64 gate columns of dz (already in its registers) with the matching 64 columns of Wh for ALL H units:
//     partial_j[128 x H] = dz[:, G_j] (128 x 64, written by the epilogue threads into a swizzled smem tile)
//                          x Wh[:, G_j]^T (H x 64, resident in smem for the whole sequence)        -- 16 MMAs per step
// and stores the partial to an L2-resident workspace; the owner of units U_i sums the n_slices partials of its 16
// columns in a fixed order (deterministic) when it differentiates the next step.  Per step and CTA: 205 KB written,
// 205 KB read, no TMA on the critical path (writer and readers all use the generic proxy, so the hand-off is a plain
// release / acquire on the per-step counter), the recurrent weights are read from HBM once per layer.
// warps: 0 = weight loader (TMA, once), 1 = MMA issuer (+TMEM owner), 2..9 = compute (thread = batch row x 8 units).
// ================================================================================================
struct RecBptt {
  float* gates[2];        // gate activations in, dz out (in place), permuted gate columns
  const float* cs[2];
  const float* dhs;       // [T', B, 2H] grad wrt layer outputs (nullable)
  const int* lens2;
  const float* dc_inject; int ldi;
  const int* inject_t;
  int* counters;          // [2][n_bt][steps], zeroed before launch
  float* pws;             // [2 parity][2 dir][n_bt][writer slice][H/4 unit quads][128 rows][4 units] partial dh:
                          // every warp-wide 16-byte access (32 rows of one unit quad) is 512 contiguous bytes
  int steps, B, H, n_bt, n_slices;
  int n_parts, part;      // the H accumulator columns are produced by n_parts MMAs of N <= 256 (multiples of 16)
  int wbox_rows;          // rows per weight TMA box (H or H/2)
  long long* dbg;
};
static_assert(kUT == 8 && kU == 16, "k_lstm_bptt maps one 32-column k-chunk of the dz tile to one unit group");

constexpr int kBUT = 8;                               // hidden units per compute thread
constexpr int kBGroups = kU / kBUT;                   // 2 unit groups per slice -> 2 warps per TMEM lane quadrant
constexpr int kBComputeThreads = 32 * 4 * kBGroups;   // 256
constexpr int kBpttThreads = kBComputeThreads;          // no dedicated issue warps: 256 threads -> 255 registers each

__global__ void __launch_bounds__(kBpttThreads, 1)
k_lstm_bptt(const __grid_constant__ CUtensorMap map_w0, const __grid_constant__ CUtensorMap map_w1, RecBptt p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int H = p.H, steps = p.steps, B = p.B;
  const uint32_t WCH = (uint32_t)H * 128;                 // one 32-column chunk of the resident weights [H rows x 128 B]
  unsigned char* smem_w = smem;                           // [2 chunks][H][32] K-major, 128B swizzle
  unsigned char* smem_a = smem + 2 * (size_t)WCH;         // [2 chunks][128][32]  (2 * WCH is a multiple of 1024: H % 16 == 0)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + 2 * A_STAGE_BYTES);
  uint64_t* w_bar = bars;
  uint64_t* a_full = bars + 1;     // compute threads -> MMA warp: the dz tile is in smem
  uint64_t* acc_full = bars + 2;   // MMA warp -> compute threads: partial is in TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x % p.n_slices;
  const int bt = (blockIdx.x / p.n_slices) % p.n_bt;
  const int d = blockIdx.x / (p.n_slices * p.n_bt);
  const bool reverse = d == 1;
  const CUtensorMap* map_w = d ? &map_w1 : &map_w0;
  int* counters = p.counters + (size_t)(d * p.n_bt + bt) * steps;
  long long* dbg = (blockIdx.x == 0) ? p.dbg : nullptr;

  if (warp == 0 && lane == 0) {
    mbar_init(smem_u32(w_bar), 1);
    mbar_init(smem_u32(a_full), 1);
    mbar_init(smem_u32(acc_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, *tmem_slot);

  if (warp == 0) {
    if (elect_one()) {
      // Wh[u, G_j]: rows = all H units, columns = this CTA's 64 permuted gate columns (two 32-column chunks)
      const uint32_t wb = smem_u32(w_bar);
      mbar_expect_tx(wb, 2 * WCH);
      for (int c = 0; c < 2; ++c)
        for (int r0 = 0; r0 < H; r0 += p.wbox_rows)    // box rows <= 256, a multiple of 8 (every piece 1024-aligned)
          tma_load_2d(smem_u32(smem_w + (size_t)c * WCH + (size_t)r0 * 128), map_w, wb, j * 64 + c * 32, r0);
    }
    __syncwarp();
    mbar_wait(smem_u32(w_bar), 0);
  }
  const uint64_t desc_a0 = make_smem_desc(smem_u32(smem_a));
  const uint64_t desc_w0 = make_smem_desc(smem_u32(smem_w));
  {
    // ================= compute threads: (batch row, kBUT hidden units) =================
    const int quad = warp & 3;
    const int sg = warp >> 2;                        // unit sub-group (kBUT = 8 units = one group of the permuted gate layout)
    const int r = quad * 32 + lane;
    const int b = bt * kBM + r;
    const bool row_ok = b < B;
    const int u0 = j * kU + sg * kBUT;
    const int z0 = j * 4 * kU + sg * 4 * kBUT;       // this thread's 32 contiguous gate columns [gate][8]
    const int len2 = row_ok ? (p.lens2 ? p.lens2[b] : steps) : 0;
    float* gates = d ? p.gates[1] : p.gates[0];
    const float* cs = d ? p.cs[1] : p.cs[0];
    const int col0 = d * H;
    const bool publisher = threadIdx.x == 0;
    const size_t chain_sz = (size_t)p.n_slices * kBM * H;                 // one (parity, dir, bt) block of the workspace
    float* pws_chain = p.pws + (size_t)(d * p.n_bt + bt) * chain_sz;      // + parity * 2 * n_bt * chain_sz
    const size_t par_stride = (size_t)2 * p.n_bt * chain_sz;
    // this thread's dz row inside the swizzled A tile: chunk sg (its 32 contiguous k), row r, 16-byte pieces XOR (r & 7)
    const uint32_t a_row = smem_u32(smem_a) + (uint32_t)sg * A_STAGE_BYTES + (uint32_t)r * 128;
    // TMEM -> workspace: this thread copies columns [hcol0, hcol0 + H / kBGroups) of its row
    const int ncol = H / kBGroups;
    const int hcol0 = sg * ncol;
    float carry[kBUT];
#pragma unroll
    for (int i = 0; i < kBUT; ++i) carry[i] = 0.f;

    for (int q = 0; q < steps; ++q) {
      const int sf = steps - 1 - q;
      const int t = reverse ? steps - 1 - sf : sf;
      const int tp = reverse ? t + 1 : t - 1;
      const bool valid = row_ok && t < len2;
      float* zrow = gates + ((i64)t * B + b) * 4 * H + z0;
      float gz[4 * kBUT], cv[kBUT], cpv[kBUT], dhv[kBUT];
#pragma unroll
      for (int e = 0; e < kBUT; ++e) { cv[e] = 0.f; cpv[e] = 0.f; dhv[e] = 0.f; }
      if (valid) {
        ldv8<4 * kBUT>(gz, zrow);                    // 256-bit accesses: half the L1 wavefronts of 128-bit ones
        ldv8<kBUT>(cv, cs + ((i64)t * B + b) * H + u0);
        if (sf > 0) ldv8<kBUT>(cpv, cs + ((i64)tp * B + b) * H + u0);
        if (p.dhs) ldv8<kBUT>(dhv, p.dhs + ((i64)t * B + b) * 2 * H + col0 + u0);
      }
      float acc[kBUT];
#pragma unroll
      for (int i = 0; i < kBUT; ++i) acc[i] = 0.f;
      if (q > 0) {
        // every CTA of the chain has stored its partial of step q-1
        if (publisher) wait_counter(counters + (q - 1), p.n_slices);
        named_bar_sync(1, kBComputeThreads);
        if (publisher && dbg) dbg[q * 8 + 0] = clock64();
        if (valid) {
          // L2-only loads (the buffer is rewritten by other SMs every two steps); kBatch slices = 2 kBatch independent
          // 16-byte loads in flight per thread, summed in slice order (deterministic)
          const float* src = pws_chain + (size_t)((q - 1) & 1) * par_stride + ((size_t)(u0 / 4) * kBM + r) * 4;
          constexpr int kBatch = 13;
          for (int i0 = 0; i0 < p.n_slices; i0 += kBatch) {
            float4 v[kBatch][kBUT / 4];
#pragma unroll
            for (int i = 0; i < kBatch; ++i)
              if (i0 + i < p.n_slices) {
#pragma unroll
                for (int h4 = 0; h4 < kBUT / 4; ++h4)
                  v[i][h4] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(i0 + i) * kBM * H + (size_t)h4 * kBM * 4));
              }
#pragma unroll
            for (int i = 0; i < kBatch; ++i)
              if (i0 + i < p.n_slices) {
#pragma unroll
                for (int h4 = 0; h4 < kBUT / 4; ++h4) {
                  acc[4 * h4] += v[i][h4].x; acc[4 * h4 + 1] += v[i][h4].y; acc[4 * h4 + 2] += v[i][h4].z; acc[4 * h4 + 3] += v[i][h4].w;
                }
              }
          }
        }
        if (publisher && dbg) dbg[q * 8 + 1] = clock64();
      }
      if (valid) {
        bool inject = false;
        if (p.dc_inject) {
          const int ti = (d == 0 && p.inject_t) ? p.inject_t[b] : 0;
          inject = ti == t;
        }
#pragma unroll
        for (int e = 0; e < kBUT; ++e) {
          const float gi = gz[e], gj = gz[kBUT + e], gf = gz[2 * kBUT + e], go = gz[3 * kBUT + e];
          const float dh = dhv[e] + acc[e];
          float dc = carry[e];
          if (inject) dc += p.dc_inject[(i64)b * p.ldi + col0 + u0 + e];
          const float tc_ = tanh_fast(cv[e]);
          gz[3 * kBUT + e] = dh * tc_ * go * (1.f - go);
          dc += dh * go * (1.f - tc_ * tc_);
          gz[e] = dc * gj * gi * (1.f - gi);
          gz[kBUT + e] = dc * gi * (1.f - gj * gj);
          gz[2 * kBUT + e] = dc * cpv[e] * gf * (1.f - gf);
          carry[e] = dc * gf;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4 * kBUT; ++i) gz[i] = 0.f;
#pragma unroll
        for (int e = 0; e < kBUT; ++e) carry[e] = 0.f;
      }
      if (q + 1 < steps) {
        // dz -> swizzled A tile (rows past B are zero), visible to the tensor core, then hand over to the MMA warp
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t addr = a_row + (uint32_t)((c ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(gz[4 * c]), "f"(gz[4 * c + 1]),
                       "f"(gz[4 * c + 2]), "f"(gz[4 * c + 3]) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        fence_before_sync();                 // this thread's tcgen05.ld of the previous step precede the next MMAs
        named_bar_sync(2, kBComputeThreads);
        if (warp == 0) {
          fence_after_sync();
          if (publisher && dbg) dbg[q * 8 + 2] = clock64();
          if (elect_one()) {
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {
                const uint64_t da = desc_a0 + (uint64_t)((c * A_STAGE_BYTES + k * UMMA_K * 4) >> 4);
                for (int pt = 0; pt < p.n_parts; ++pt) {
                  const int n = min(p.part, H - pt * p.part);
                  const uint64_t dw = desc_w0 + (uint64_t)((c * WCH + (uint32_t)(pt * p.part) * 128 + k * UMMA_K * 4) >> 4);
                  umma_tf32(tmem_base + (uint32_t)(pt * p.part), da, dw, make_idesc_tf32(kBM, n, 0), (c > 0 || k > 0) ? 1u : 0u);
                }
              }
            umma_commit(smem_u32(acc_full));
          }
          __syncwarp();
        }
      }
      // dz to HBM for the weight-gradient GEMMs (off the inter-CTA critical path: overlaps the MMA)
      if (row_ok) stv8<4 * kBUT>(zrow, gz);
      if (q + 1 < steps) {
        mbar_wait(smem_u32(acc_full), q & 1);
        fence_after_sync();
        if (publisher && dbg) dbg[q * 8 + 3] = clock64();
        // unit quad (column / 4) of writer slice j lives at ((j * H/4 + quad) * 128 + row) * 4
        float* dst = pws_chain + (size_t)(q & 1) * par_stride + (size_t)j * kBM * H + ((size_t)(hcol0 / 4) * kBM + r) * 4;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)hcol0;
        // TMEM reads (64 B/clk) and the global stores overlap: the next 16 columns are in flight while these are stored
        int c = 0;
        float va[16], vb[16];
        if (ncol >= 16) tmem_ld16_nowait(taddr, va);
        for (; c + 16 <= ncol; c += 32) {
          tmem_ld_wait();
          if (c + 32 <= ncol) tmem_ld16_nowait(taddr + (uint32_t)(c + 16), vb);
#pragma unroll
          for (int g = 0; g < 4; ++g) stv<4>(dst + (size_t)(c / 4 + g) * kBM * 4, va + g * 4);
          if (c + 32 <= ncol) {
            tmem_ld_wait();
            if (c + 48 <= ncol) tmem_ld16_nowait(taddr + (uint32_t)(c + 32), va);
#pragma unroll
            for (int g = 0; g < 4; ++g) stv<4>(dst + (size_t)((c + 16) / 4 + g) * kBM * 4, vb + g * 4);
          }
        }
        c = ncol / 16 * 16;
        for (; c < ncol; c += 4) {
          float v[4];
          tmem_ld_cols<4>(taddr + (uint32_t)c, v);
          stv<4>(dst + (size_t)(c / 4) * kBM * 4, v);
        }
        fence_before_sync();
        if (publisher && dbg) dbg[q * 8 + 4] = clock64();
        named_bar_sync(1, kBComputeThreads);     // all partial stores of this CTA are ordered before the release below
        if (publisher) {
          red_release_gpu(counters + q, 1);
          if (dbg) dbg[q * 8 + 5] = clock64();
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- host side ------------------------------------------------------------------------------------
// generic fp32 tensor map, up to 3 dims; dims/strides innermost first (strides in elements, dim0 stride = 1)
inline CUtensorMap make_map_nd(const float* ptr, int nd, const i64* dims, const i64* strides_elems, const int* box) {
  CUtensorMap m;
  cuuint64_t gdim[3]; cuuint64_t gstr[2]; cuuint32_t bx[3]; cuuint32_t estr[3] = {1, 1, 1};
  for (int i = 0; i < nd; ++i) { gdim[i] = (cuuint64_t)dims[i]; bx[i] = (cuuint32_t)box[i]; }
  for (int i = 1; i < nd; ++i) gstr[i - 1] = (cuuint64_t)strides_elems[i] * 4;
  EncodeTiledFn fn = encode_fn();
  if (!fn) throw std::runtime_error("e2t: cuTensorMapEncodeTiled entry point not found");
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)nd, const_cast<float*>(ptr), gdim, gstr, bx, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("e2t: cuTensorMapEncodeTiled (nd) failed with code " + std::to_string((int)r));
  return m;
}

template <bool BWD>
inline size_t rec_smem_bytes(int nkc, int stages) {
  return (size_t)nkc * Geo<BWD>::WCHUNK + (size_t)stages * kGroup * A_STAGE_BYTES + (2 * stages + 4) * 8 + 1024;
}

// picks the A-ring depth that fits; returns 0 if the resident weights do not fit at all
template <bool BWD>
inline int rec_pick_stages(int nkc) {
  const size_t cap = 227 * 1024;
  for (int s = 4; s >= 2; --s)     // ring depth in groups of kGroup chunks
    if (rec_smem_bytes<BWD>(nkc, s) <= cap) return s;
  return 0;
}

inline int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

// Can the persistent kernels run a layer of this shape?  (H multiple of 16; weights fit; grid co-resident)
inline bool rec_supported(int B, int H, int steps) {
  if (H % kU != 0 || H < kU || steps < 1 || B < 1) return false;
  const int n_bt = (B + kBM - 1) / kBM, n_slices = H / kU;
  if (2 * n_bt * n_slices > sm_count()) return false;
  if (!rec_pick_stages<false>((H + BK - 1) / BK)) return false;
  if (!rec_pick_stages<true>((4 * H + BK - 1) / BK)) return false;
  return true;
}

template <bool BWD, typename P>
inline void rec_launch(cudaStream_t st, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w0,
                       const CUtensorMap& w1, P& p) {
  auto kfn = k_lstm_rec<BWD, P>;
  const size_t smem = rec_smem_bytes<BWD>(p.nkc, p.stages);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    E2T_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  E2T_CHECK(cudaMemsetAsync(p.counters, 0, (size_t)2 * p.n_bt * p.steps * sizeof(int), st));
  // diagnostic: E2T_REC_DEBUG=<n> prints the per-step clock64 timeline of CTA 0 for the first n launches
  static int dbg_left = getenv("E2T_REC_DEBUG") ? atoi(getenv("E2T_REC_DEBUG")) : 0;
  p.dbg = nullptr;
  static int pub_mode = getenv("E2T_REC_PUB") ? atoi(getenv("E2T_REC_PUB")) : 3;
  p.pub_mode = pub_mode;
  dim3 grid((unsigned)(2 * p.n_bt * p.n_slices));
  const size_t dbg_n = ((size_t)p.steps + grid.x) * 8;
  if (dbg_left > 0) {
    E2T_CHECK(cudaMalloc(&p.dbg, dbg_n * sizeof(long long)));
    E2T_CHECK(cudaMemsetAsync(p.dbg, 0, dbg_n * sizeof(long long), st));
  }
  // (a cluster sharing every A tile by TMA multicast was measured slower than unicast -- tools/ubench/ingest.cu: the
  // per-SM ingest of a 208 KB tile is ~2300 cycles alone or hot-shared by 100 CTAs, ~3200 with multicast -- and removed)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(kThreadsRec); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  attrs[na].id = cudaLaunchAttributeCooperative; attrs[na].val.cooperative = 1; ++na;   // all CTAs co-resident
  cfg.attrs = attrs; cfg.numAttrs = na;
  E2T_CHECK(cudaLaunchKernelEx(&cfg, kfn, a0, a1, w0, w1, p));
  if (p.dbg) {
    --dbg_left;
    std::vector<long long> hst(dbg_n);
    E2T_CHECK(cudaStreamSynchronize(st));
    E2T_CHECK(cudaMemcpy(hst.data(), p.dbg, hst.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(p.dbg);
    if (p.steps > 11) {
      // probe step 10, all CTAs, ns relative to the earliest flag-seen
      long long t0 = -1;
      for (unsigned c = 0; c < grid.x; ++c) { const long long v = hst[((size_t)p.steps + c) * 8]; if (v && (t0 < 0 || v < t0)) t0 = v; }
      fprintf(stderr, "[rec %s probe step 10] cta: flag_seen tma_issued first_full mma_committed acc_seen h_stored bar released (ns)\n", BWD ? "bwd" : "fwd");
      for (unsigned c = 0; c < grid.x; ++c) {
        const long long* e = &hst[((size_t)p.steps + c) * 8];
        fprintf(stderr, "  cta %3u:", c);
        for (int k = 0; k < 8; ++k) fprintf(stderr, " %6lld", e[k] ? e[k] - t0 : -1);
        fprintf(stderr, "\n");
      }
    }
    fprintf(stderr, "[rec %s] steps=%d B=%d H=%d nkc=%d ring=%dx%d grid=%u  (cycles rel. to flag-seen of each step)\n"
                    "  step  flag->tma_issued  ->first_full  ->mma_committed  ->acc_seen  ->h_stored  ->bar_passed  ->released | step_total\n",
            BWD ? "bwd" : "fwd", p.steps, p.B, p.H, p.nkc, p.stages, kGroup, grid.x);
    for (int s = 1; s < p.steps; ++s) {
      const long long* e = &hst[(size_t)s * 8];
      const long long prev = s > 1 ? hst[(size_t)(s - 1) * 8] : 0;
      fprintf(stderr, "  %4d  %8lld %8lld %8lld %8lld %8lld %8lld %8lld | %8lld\n", s, e[1] - e[0], e[2] - e[0], e[3] - e[0],
              e[4] - e[0], e[5] - e[0], e[6] - e[0], e[7] ? e[7] - e[0] : 0, prev ? e[0] - prev : 0);
    }
  }
}

// Forward of one BiLSTM layer.  KT[d]: packed transposed kernels [4H, ldkt] (Wh^T at column In).
inline void rec_forward(cudaStream_t st, float* const gates[2], float* const cs[2], float* hs, float* hd,
                        const float* const KT[2], int ldkt, int In, const int* lens2, int* counters, int steps, int B,
                        int H, DropP dp, int drop_F) {
  RecFwdP p{};
  for (int d = 0; d < 2; ++d) { p.gates[d] = gates[d]; p.cs[d] = cs[d]; }
  p.hs = hs; p.hd = hd; p.lens2 = lens2; p.counters = counters;
  p.steps = steps; p.B = B; p.H = H;
  p.n_bt = (B + kBM - 1) / kBM; p.n_slices = H / kU; p.nkc = (H + BK - 1) / BK;
  p.stages = rec_pick_stages<false>(p.nkc);
  p.dp = dp; p.drop_F = drop_F;
  CUtensorMap ma[2], mw[2];
  for (int d = 0; d < 2; ++d) {
    // A: h of this direction, [steps*B rows, H cols] with row pitch 2H; K tail / OOB rows zero-filled
    const i64 adims[2] = {H, (i64)steps * B}, astr[2] = {1, 2 * (i64)H};
    const int abox[2] = {BK, kBM};
    ma[d] = make_map_nd(hs + (i64)d * H, 2, adims, astr, abox);
    // W: Wh^T rows in the permuted gate order (64 consecutive rows = the 4 gates of one CTA's 16 units)
    const i64 wdims[2] = {H, 4 * (i64)H}, wstr[2] = {1, (i64)ldkt};
    const int wbox[2] = {BK, 4 * kU};
    mw[d] = make_map_nd(KT[d] + In, 2, wdims, wstr, wbox);
  }
  rec_launch<false>(st, ma[0], ma[1], mw[0], mw[1], p);
}

// BPTT of one BiLSTM layer.  K[d]: canonical kernels [In+H, 4H]; gates hold activations in, dz out.
inline void rec_backward(cudaStream_t st, float* const gates[2], const float* const cs[2], const float* dhs,
                         const float* const K[2], int In, const int* lens2, const float* dc_inject, int ldi,
                         const int* inject_t, int* counters, int steps, int B, int H) {
  RecBwdP p{};
  for (int d = 0; d < 2; ++d) { p.gates[d] = gates[d]; p.cs[d] = cs[d]; }
  p.dhs = dhs; p.lens2 = lens2; p.dc_inject = dc_inject; p.ldi = ldi; p.inject_t = inject_t; p.counters = counters;
  p.steps = steps; p.B = B; p.H = H;
  p.n_bt = (B + kBM - 1) / kBM; p.n_slices = H / kU; p.nkc = (4 * H + BK - 1) / BK;
  p.stages = rec_pick_stages<true>(p.nkc);
  CUtensorMap ma[2], mw[2];
  for (int d = 0; d < 2; ++d) {
    const i64 adims[2] = {4 * (i64)H, (i64)steps * B}, astr[2] = {1, 4 * (i64)H};
    const int abox[2] = {BK, kBM};
    ma[d] = make_map_nd(gates[d], 2, adims, astr, abox);
    // W: rows u of Wh (canonical [H, 4H] block below the In input rows) = K-major B operand of dz Wh^T
    const i64 wdims[2] = {4 * (i64)H, H}, wstr[2] = {1, 4 * (i64)H};
    const int wbox[2] = {BK, kU};
    mw[d] = make_map_nd(K[d] + (i64)In * 4 * H, 2, wdims, wstr, wbox);
  }
  rec_launch<true>(st, ma[0], ma[1], mw[0], mw[1], p);
}

inline size_t bptt_smem_bytes(int H) { return (size_t)2 * H * 128 + 2 * A_STAGE_BYTES + 8 * 8 + 1024; }
inline size_t bptt_ws_floats(int B, int H) {
  const int n_bt = (B + kBM - 1) / kBM, n_slices = H / kU;
  return (size_t)2 * 2 * n_bt * n_slices * kBM * H;
}
inline bool bptt_supported(int B, int H) {
  if (H % kU != 0 || H < kU || H > 512) return false;
  const int n_bt = (B + kBM - 1) / kBM, n_slices = H / kU;
  if (2 * n_bt * n_slices > sm_count()) return false;
  return bptt_smem_bytes(H) <= 227 * 1024;
}

// BPTT of one BiLSTM layer, reduce-scatter formulation.  K[d]: canonical kernels [In+H, 4H] with permuted gate columns.
inline void rec_backward_rs(cudaStream_t st, float* const gates[2], const float* const cs[2], const float* dhs,
                            const float* const K[2], int In, const int* lens2, const float* dc_inject, int ldi,
                            const int* inject_t, int* counters, float* pws, int steps, int B, int H) {
  RecBptt p{};
  for (int d = 0; d < 2; ++d) { p.gates[d] = gates[d]; p.cs[d] = cs[d]; }
  p.dhs = dhs; p.lens2 = lens2; p.dc_inject = dc_inject; p.ldi = ldi; p.inject_t = inject_t; p.counters = counters;
  p.pws = pws; p.steps = steps; p.B = B; p.H = H;
  p.n_bt = (B + kBM - 1) / kBM; p.n_slices = H / kU;
  p.n_parts = (H + 255) / 256;
  p.part = (((H + p.n_parts - 1) / p.n_parts) + 15) / 16 * 16;
  CUtensorMap mw[2];
  for (int d = 0; d < 2; ++d) {
    // Wh = rows [In, In+H) of the canonical kernel: [H rows (units), 4H permuted gate columns]; box = 32 columns x 200 rows
    const i64 wdims[2] = {4 * (i64)H, H}, wstr[2] = {1, 4 * (i64)H};
    p.wbox_rows = H <= 256 ? H : H / 2;
    const int wbox[2] = {BK, p.wbox_rows};
    mw[d] = make_map_nd(K[d] + (i64)In * 4 * H, 2, wdims, wstr, wbox);
  }
  auto kfn = k_lstm_bptt;
  const size_t smem = bptt_smem_bytes(H);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    E2T_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  E2T_CHECK(cudaMemsetAsync(p.counters, 0, (size_t)2 * p.n_bt * p.steps * sizeof(int), st));
  static int dbg_left = getenv("E2T_REC_DEBUG") ? atoi(getenv("E2T_REC_DEBUG")) : 0;
  p.dbg = nullptr;
  if (dbg_left > 0) {
    E2T_CHECK(cudaMalloc(&p.dbg, (size_t)p.steps * 8 * sizeof(long long)));
    E2T_CHECK(cudaMemsetAsync(p.dbg, 0, (size_t)p.steps * 8 * sizeof(long long), st));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * p.n_bt * p.n_slices)); cfg.blockDim = dim3(kBpttThreads);
  cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeCooperative; attrs[0].val.cooperative = 1;   // all CTAs co-resident
  cfg.attrs = attrs; cfg.numAttrs = 1;
  E2T_CHECK(cudaLaunchKernelEx(&cfg, kfn, mw[0], mw[1], p));
  if (p.dbg) {
    --dbg_left;
    std::vector<long long> hst((size_t)p.steps * 8);
    E2T_CHECK(cudaStreamSynchronize(st));
    E2T_CHECK(cudaMemcpy(hst.data(), p.dbg, hst.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(p.dbg);
    fprintf(stderr, "[rec bptt] steps=%d B=%d H=%d grid=%u (cycles rel. to flag-seen)\n"
                    "  step  ->partials_summed  ->dz_in_smem  ->acc_seen  ->partial_stored  ->released | step_total\n",
            p.steps, p.B, p.H, cfg.gridDim.x);
    for (int q = 1; q + 1 < p.steps; ++q) {
      const long long* e = &hst[(size_t)q * 8];
      const long long prev = q > 1 ? hst[(size_t)(q - 1) * 8] : 0;
      fprintf(stderr, "  %4d  %8lld %8lld %8lld %8lld %8lld | %8lld\n", q, e[1] - e[0], e[2] - e[0], e[3] - e[0], e[4] - e[0],
              e[5] - e[0], prev ? e[0] - prev : 0);
    }
  }
}

}  // namespace rec
