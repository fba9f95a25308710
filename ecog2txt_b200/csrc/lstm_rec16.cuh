// lstm_rec16.cuh -- persistent BiLSTM forward recurrence, second generation (A5 of SURVEY.md section 8a).
//
// Same decomposition as k_lstm_rec<false> (lstm_rec.cuh): grid = 2 directions x batch tiles of 128 rows x H/16 unit slices,
// one cooperative launch runs ALL time steps, the CTA's recurrent weight slice stays in shared memory for the whole
// sequence, accumulators live in TMEM, gate non-linearities + cell update come straight out of tcgen05.ld registers.
// What changed, and why (B200 timelines of the first generation, profiles/r1p_rec_timeline.txt: 8.5 us per step of which
// 2.65 us were 52 MMA issues, 0.7 us the release and 3.4 us the counter hand-off between the 25 CTAs of a chain):
//
//  * 16-bit operands.  h is a product of a sigmoid and a tanh, |h| < 1, and the recurrent weights are O(0.1): both are
//    exactly representable in fp16 with the SAME 11-bit significand the tensor core keeps of a tf32 operand (range
//    6e-5 .. 65504 normal, absolute error < 3e-8 below).  kind::f16 consumes K = 16 per instruction instead of 8: 25 MMAs
//    per step instead of 52, half the shared memory for the weights (56 KB) and half the bytes of h through L2.
//  * No flags, no counters, no fences on the critical path: the DATA is its own flag.  The exchange buffer
//    hx [dir][t][b][H] (fp16) is filled with the bit pattern 0xFFFF before the launch (a NaN that the epilogue can never
//    produce: cvt.rn.f16.f32 canonicalises NaN to 0x7FFF); every step writes its own time slot exactly once, so a consumer
//    simply re-loads a 16-byte piece until none of its words is the fill pattern.  One L2 round trip replaces
//    store -> release -> counter -> poll -> acquire -> TMA.
//  * The 256 epilogue threads are also the loaders: they would otherwise idle while the step's h arrives.  Each pulls its
//    pieces with 16-byte ld.relaxed.gpu, writes them into the 128B-swizzled K-major operand tile, fences the async proxy and
//    arrives on the k-chunk's mbarrier; the MMA warp issues chunk by chunk as they complete.
//
// 8 warps, all loader + epilogue (thread = batch row x 8 hidden units); warp 0 also loads the weights (TMA, once), owns the
// TMEM allocation and issues the MMAs once its own pieces are in place (a ninth warp would cap every thread at 168 registers --
// ptxas budgets for 384 threads -- and spill the 28 outstanding 16-byte loads).
#pragma once
#include <cuda_fp16.h>

#include "lstm_rec.cuh"

namespace rec16 {

using namespace tc;
using rec::kU;
using rec::kBM;
using rec::kUT;

constexpr int kWorkThreads = 256;
constexpr int kThreads16 = kWorkThreads + 64;     // + MMA warp (8) + load / check warp (9)
constexpr int kKC = 64;                         // fp16 elements per 128-byte swizzle row = one k-chunk
constexpr uint32_t kAChunk = kBM * 128;         // 16 KB: 128 rows x 128 B
constexpr uint32_t kWChunk = 4 * kU * 128;      // 8 KB: 64 gate rows x 128 B
constexpr uint32_t kFill32 = 0xFFFFFFFFu;       // two fp16 fill patterns
constexpr int kMaxRedo = 3;                     // bounded re-pulls of a tile whose accumulator came out NaN
constexpr int kDbg = 32;                        // timeline slots per step (E2T_REC_DEBUG)

// cute::UMMA::InstrDescriptor for kind::f16: c_format F32=1 [4,6) | a_format F16=0 [7,10) | b_format F16=0 [10,13)
// | a_major / b_major = K (0) | N>>3 [17,23) | M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ uint4 ld_relaxed_v4(const void* p) {
  uint4 v;
  // E2T_STRONG_POLL: ld.relaxed.gpu (LDG.STRONG.GPU) -- measured far slower than L1-bypassing weak loads (ld.cg), which
  // read the same L2 copy; asm volatile keeps the compiler from caching the value across poll rounds
#ifdef E2T_STRONG_POLL
  asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
#else
  asm volatile("ld.global.cg.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
#endif
  return v;
}
__device__ __forceinline__ void st_relaxed_v4(void* p, uint4 v) {
#ifdef E2T_STRONG_POLL
  asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#else
  asm volatile("st.global.cg.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#endif
}
__device__ __forceinline__ bool has_fill(const uint4& v) {
  return v.x == kFill32 || v.y == kFill32 || v.z == kFill32 || v.w == kFill32;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);      // NaN -> 0x7FFF, never the 0xFFFF fill pattern
  return *reinterpret_cast<const uint32_t*>(&h);
}

struct Fwd16P {
  float* gates[2];        // [T', B, 4H] x-projection (+bias) in, gate activations out (permuted gate order)
  __half* hx;             // fp16 exchange buffer [2][T'][n_bt][NKC][128 rows][64] = the operand tile's shared-memory image
                          // (128B swizzle applied by the writers), pre-filled with 0xFFFF
  const int* lens2;       // [B] (nullable: all steps valid)
  int steps, B, H, n_bt, n_slices;
  int has_hd;
  int dbg_skip;           // E2T_REC_DBGSKIP, timing experiments only (results are WRONG): bit 0 no result stores, bit 1 no x-projection loads; bit 2 (results stay right) forces re-pulls
  DropP dp; int drop_F;
  long long* dbg;         // E2T_REC_DEBUG: per-step clock64 stamps of CTA 0 ([steps][8]), else NULL
  int* trap_rec;          // mapped host memory (nullable): who timed out where, written right before the trap
};
// A wait that ran out of patience records (site, block, thread, step, chunk, extra) for the host, then traps.
__device__ __noinline__ void timeout_trap(int* rec, int site, int s, int kc, int extra) {
  if (rec && atomicCAS(rec, 0, site) == 0) {
    rec[1] = (int)blockIdx.x; rec[2] = (int)threadIdx.x; rec[3] = s; rec[4] = kc; rec[5] = extra;
    __threadfence_system();
  }
  __trap();
}
__device__ __forceinline__ void mbar_wait_rec(uint32_t bar, uint32_t parity, int* rec, int site, int s, int kc) {
  // bounded by TIME (2^31 cycles ~ 1 s), not by a try count: how long one try_wait parks the thread is up to the hardware,
  // and 2^14 tries (gemm_tc.cuh: mbar_wait) were measured to run out on healthy launches of this kernel
  uint32_t done = 0;
  const long long t0 = clock64();
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(1000000u)
        : "memory");
    if (done) return;
    if (clock64() - t0 > (1LL << 31)) break;
  }
  timeout_trap(rec, site, s, kc, (int)parity);
}
// tensor maps of one launch (one kernel parameter: TMA descriptors must live in param / const space)
struct Fwd16Maps {
  CUtensorMap w[2];       // Wh^T fp16 [4H, Hp], box 64 x 64, 128B swizzle (load, once)
  CUtensorMap gates[2];   // [T', B, 4H] fp32, box 32 x 128 x 1, 128B swizzle (store)
  CUtensorMap cs[2];      // [T', B, H] fp32, box 16 x 128 x 1, 64B swizzle (store)
  CUtensorMap hs, hd;     // [T', B, 2H] fp32, box 16 x 128 x 1, 64B swizzle (store)
};

constexpr uint32_t kStageGates = 2 * kBM * 128;   // two swizzled [128 x 32 fp32] sub-tiles (unit group 0 / 1)
constexpr uint32_t kStageSmall = kBM * 64;        // dense [128 x 16 fp32]
constexpr uint32_t kStageBytes = kStageGates + 3 * kStageSmall;

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// contiguous global -> shared bulk copy (no tensor map: one request, not one per box row)
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint32_t ld_cg_u32(const void* p) {
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ bool bar_red_or(int id, int nthreads, bool pred) {
  uint32_t out;
  asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, %2, %3, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(out) : "r"((uint32_t)pred), "r"(id), "r"(nthreads) : "memory");
  return out != 0;
}

// Per step and CTA: 112 KB of h in (7 contiguous bulk copies of the tile's shared-memory image), 32 KB of x-projection in
// (TMA), 24 KB of c / h / dropped h out (TMA bulk stores from swizzled stages), 32 KB of gate activations and 4 KB of fp16
// h out through the LSU.  History (profiles/r2*_timeline.txt, cycles per step at 1.965 GHz):
//  * h pulled with 16-byte LSU loads: 15-25 B/clk per SM, 6500 cycles; TMA / bulk copies ingest the tile in ~1800;
//  * x-projection prefetched into registers with 32-byte LSU loads at a 6400-byte row stride: 2400 cycles of LSU queueing
//    in front of the staging stores of the same threads;
//  * probe by the epilogue threads after their staging: the whole staging phase (5-6 k cycles) sat on the critical path;
//  * explicit fill-pattern check of the landed tile by one warp: 32 ld.shared per lane and chunk = ~630 cycles per chunk,
//    4400 per step, although all seven chunks had landed after ~1800 (r2q/r2r).
// Now: warps 0-7 epilogue (thread = batch row x 8 units), warp 8 MMA issue, warp 9 probe + copies.
//  * The load warp polls one word of every producer warp's h store of the previous step (the probe); when all are there the
//    tile is almost surely complete in L2 and is pulled.  No check of what landed: the fill pattern 0xFFFF is an fp16 NaN,
//    so a piece that was not there yet turns the whole accumulator row into NaN.  The epilogue threads test one accumulator
//    word of their row, agree with one bar.red.or, and in the (never yet observed) bad case the step's copies and MMAs are
//    simply repeated (at most kMaxRedo times, so that a genuinely diverged model cannot hang the kernel).
//  * The x-projection tile of step s+1 is requested as soon as every thread has read the tile of step s into registers.
template <int NKC>
__global__ void __launch_bounds__(kThreads16, 1)
k_lstm_fwd16(const __grid_constant__ Fwd16Maps maps, Fwd16P p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* smem_w = smem;                                   // [NKC][64 rows][128 B]
  unsigned char* smem_a = smem + (size_t)NKC * kWChunk;           // [NKC][128 rows][128 B]
  unsigned char* smem_o = smem_a + (size_t)NKC * kAChunk;         // x-projection tile (2 x 16 KB) | stages cs | hs | hd
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_o + kStageBytes);
  uint64_t* w_bar = bars;
  uint64_t* acc_full = bars + 1;
  uint64_t* a_full = bars + 2;                                    // [NKC]
  uint64_t* z_bar = bars + 2 + NKC;                               // x-projection tile of the step has landed
  uint64_t* verdict_bar = bars + 3 + NKC;                         // epilogue -> MMA / load warp: accumulator accepted or redo
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 + 2 * NKC);
  volatile uint32_t* verdict = tmem_slot + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x % p.n_slices;
  const int bt = (blockIdx.x / p.n_slices) % p.n_bt;
  const int d = blockIdx.x / (p.n_slices * p.n_bt);
  const bool reverse = d == 1;
  const int steps = p.steps, B = p.B, H = p.H;
  long long* dbg = (blockIdx.x == 0) ? p.dbg : nullptr;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(w_bar), 1);
    mbar_init(smem_u32(acc_full), 1);
    mbar_init(smem_u32(z_bar), 1);
    mbar_init(smem_u32(verdict_bar), 1);
    for (int k = 0; k < NKC; ++k) mbar_init(smem_u32(&a_full[k]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 8) tmem_alloc(smem_u32(tmem_slot), 128);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, *tmem_slot);
  // this chain's tiles: tile of step t at hx_chain + t * tile_stride (bytes); piece (row, k) of a tile at
  // (k / 64) * 16 KB + row * 128 + (((k % 64) / 8) ^ (row & 7)) * 16
  const size_t tile_bytes = (size_t)NKC * kAChunk;
  const size_t tile_stride = (size_t)p.n_bt * tile_bytes;
  unsigned char* hx_chain = reinterpret_cast<unsigned char*>(p.hx) + ((size_t)d * steps * p.n_bt + bt) * tile_bytes;

  if (warp == 8) {
    // ================= MMA warp: weight TMA (once); per step the MMAs of every chunk as it lands =================
    if (elect_one()) {
      const uint32_t wb = smem_u32(w_bar);
      mbar_expect_tx(wb, (uint32_t)NKC * kWChunk);
      for (int kc = 0; kc < NKC; ++kc)     // 64 permuted gate rows of Wh^T (fp16), 64 k columns; the K tail is zero-filled
        tma_load_2d(smem_u32(smem_w + (size_t)kc * kWChunk), &maps.w[d], wb, kc * kKC, j * 4 * kU);
    }
    __syncwarp();
    mbar_wait_rec(smem_u32(w_bar), 0, p.trap_rec, 5, 0, 0);
    fence_after_sync();
    constexpr uint32_t idesc = make_idesc_f16(kBM, 4 * kU);
    const uint64_t desc_a0 = make_smem_desc(smem_u32(smem_a));
    const uint64_t desc_w0 = make_smem_desc(smem_u32(smem_w));
    uint32_t round = 0;                                // tiles pulled so far = phase of a_full / verdict_bar
    for (int s = 1; s < steps; ++s) {
      for (;;) {
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc) {
          mbar_wait_rec(smem_u32(&a_full[kc]), round & 1u, p.trap_rec, 7, s, kc);
          fence_after_sync();
          if (elect_one()) {
            const int nk = min(4, (H - kc * kKC) / 16);       // K = 16 per instruction; H % 16 == 0
            for (int k = 0; k < nk; ++k)
              umma_f16(tmem_base, desc_a0 + (uint64_t)((kc * kAChunk + k * 32) >> 4), desc_w0 + (uint64_t)((kc * kWChunk + k * 32) >> 4),
                       idesc, (kc > 0 || k > 0) ? 1u : 0u);
            if (kc == NKC - 1) umma_commit(smem_u32(acc_full));
          }
          __syncwarp();
          if (dbg && lane == 0) dbg[s * kDbg + 16 + kc] = clock64();
        }
        if (dbg && lane == 0) dbg[s * kDbg + 2] = clock64();
        mbar_wait_rec(smem_u32(verdict_bar), round & 1u, p.trap_rec, 9, s, 0);
        ++round;
        if (*verdict == 0u) break;
      }
    }
  } else if (warp == 9) {
    // ================= load warp: probe, then the h tile as NKC contiguous bulk copies =================
    // probe word idx = 32 k + lane: producer slice idx / 8, its epilogue warp idx % 8 (lane 0 of that warp: row 32 (w & 3),
    // unit group w >> 2)
    constexpr int kProbes = (kWorkThreads + 31) / 32;      // n_slices * 8 <= kWorkThreads words
    const int n_probe = p.n_slices * 8;
    uint32_t probe_off[kProbes];
#pragma unroll
    for (int k = 0; k < kProbes; ++k) {
      const int idx = 32 * k + lane;
      const int prow = 32 * (idx & 3), pk = (idx >> 3) * kU + ((idx >> 2) & 1) * kUT;
      probe_off[k] = (uint32_t)(pk / kKC) * kAChunk + (uint32_t)prow * 128 + (uint32_t)((((pk % kKC) / 8) ^ (prow & 7)) << 4);
    }
    uint32_t round = 0;
    for (int s = 1; s < steps; ++s) {
      const int t = reverse ? steps - 1 - s : s;
      const int t_src = reverse ? t + 1 : t - 1;
      const unsigned char* base = hx_chain + (size_t)t_src * tile_stride;
      // words 32 kc .. 32 kc + 31 belong to the four producer slices of chunk kc: a chunk is pulled as soon as ITS producers
      // are in, so that when the slowest CTA of the chain arrives only one chunk (and its MMAs) is still outstanding
      for (bool first = true;; first = false) {
        uint32_t pending = 0, pulled = 0;
        if (first) {
#pragma unroll
          for (int k = 0; k < kProbes; ++k)
            if (32 * k + lane < n_probe) pending |= 1u << k;
        }
        const long long t0 = clock64();
        while (pulled != (1u << NKC) - 1u) {
          uint32_t v[kProbes];
#pragma unroll
          for (int k = 0; k < kProbes; ++k)
            if (pending & (1u << k)) v[k] = ld_cg_u32(base + probe_off[k]);
#pragma unroll
          for (int k = 0; k < kProbes; ++k)
            if ((pending & (1u << k)) && v[k] != kFill32) pending &= ~(1u << k);
#pragma unroll
          for (int kc = 0; kc < NKC; ++kc) {
            if (pulled & (1u << kc)) continue;
            if (!__all_sync(0xffffffffu, kc >= kProbes || !(pending & (1u << kc)))) continue;
            if (dbg && lane == 0 && pulled == 0) dbg[s * kDbg + 1] = clock64();
            pulled |= 1u << kc;
            if (elect_one()) {
              const uint32_t fb = smem_u32(&a_full[kc]);
              mbar_expect_tx(fb, kAChunk);
              bulk_load(smem_u32(smem_a + (size_t)kc * kAChunk), base + (size_t)kc * kAChunk, kAChunk, fb);
            }
            __syncwarp();
          }
          if (clock64() - t0 > (1LL << 31)) timeout_trap(p.trap_rec, 1, s, 0, (int)pending);
        }
        if (dbg && lane == 0 && first) dbg[s * kDbg + 0] = clock64();
        mbar_wait_rec(smem_u32(verdict_bar), round & 1u, p.trap_rec, 6, s, 0);
        ++round;
        if (*verdict == 0u) break;
        if (dbg && lane == 0) dbg[s * kDbg + 7] += 1;           // tiles pulled again
      }
    }
  } else {
    // ================= epilogue: thread = (batch row, 8 hidden units) =================
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may read
    const int ug = warp >> 2;                        // unit group (8 units)
    const int r = quad * 32 + lane;
    const int b = bt * kBM + r;
    const bool row_ok = b < B;
    const int u0 = j * kU + ug * kUT;
    const int z0 = j * 4 * kU + ug * 4 * kUT;        // 32 contiguous gate columns [gate][8] in the permuted layout
    const int len2 = row_ok ? (p.lens2 ? p.lens2[b] : steps) : 0;
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ug * 4 * kUT);
    float* gates = d ? p.gates[1] : p.gates[0];
    const int col0 = d * H;
    float carry[kUT];
#pragma unroll
    for (int i = 0; i < kUT; ++i) carry[i] = 0.f;
    // this thread's 32 gate columns [gate][8] of its row in the swizzled x-projection tile
    const uint32_t st_g = smem_u32(smem_o) + (uint32_t)ug * (kBM * 128) + (uint32_t)r * 128;
    // small stages [128 rows][16 fp32], 64B swizzle: 16-byte chunk c of row r sits at chunk c ^ ((r >> 1) & 3)
    const uint32_t st_s0 = smem_u32(smem_o) + kStageGates + (uint32_t)r * 64 + (uint32_t)(((2 * ug) ^ ((r >> 1) & 3)) << 4);
    const uint32_t st_s1 = smem_u32(smem_o) + kStageGates + (uint32_t)r * 64 + (uint32_t)(((2 * ug + 1) ^ ((r >> 1) & 3)) << 4);
    // this thread's 16-byte piece of the exchange tile (8 units of h, fp16)
    const uint32_t hx_off = (uint32_t)(u0 / kKC) * kAChunk + (uint32_t)r * 128 + (uint32_t)((((u0 % kKC) / 8) ^ (r & 7)) << 4);
    const uint32_t so = smem_u32(smem_o);
    const uint32_t zb = smem_u32(z_bar);
    uint32_t acc_round = 0;

    if (threadIdx.x == 0) {                            // x-projection tile of the first step
      const int t0s = reverse ? steps - 1 : 0;
      mbar_expect_tx(zb, kStageGates);
      rec::tma_load_3d(so, &maps.gates[d], zb, j * 4 * kU, bt * kBM, t0s);
      rec::tma_load_3d(so + kBM * 128, &maps.gates[d], zb, j * 4 * kU + 4 * kUT, bt * kBM, t0s);
    }
    for (int s = 0; s < steps; ++s) {
      const int t = reverse ? steps - 1 - s : s;
      const bool valid = row_ok && t < len2;
      float z[4 * kUT];
      float acc[4 * kUT];
      mbar_wait_rec(zb, (uint32_t)s & 1u, p.trap_rec, 8, s, 0);
#pragma unroll
      for (int c = 0; c < 8; ++c)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(z[4 * c]), "=f"(z[4 * c + 1]), "=f"(z[4 * c + 2]), "=f"(z[4 * c + 3])
                     : "r"(st_g + (uint32_t)((c ^ (r & 7)) << 4)) : "memory");
      // the previous step's bulk stores must have finished reading the small stages before anyone writes them again
      if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      rec::named_bar_sync(2, kWorkThreads);           // every thread holds its x-projection: the tile can be refilled
      if (threadIdx.x == 0 && s + 1 < steps) {
        const int tn = reverse ? t - 1 : t + 1;
        if (p.dbg_skip & 2) rec::mbar_arrive(zb);
        else {
          mbar_expect_tx(zb, kStageGates);
          rec::tma_load_3d(so, &maps.gates[d], zb, j * 4 * kU, bt * kBM, tn);
          rec::tma_load_3d(so + kBM * 128, &maps.gates[d], zb, j * 4 * kU + 4 * kUT, bt * kBM, tn);
          if (s + 2 < steps) {                         // and the one after that is asked into L2
            const int tnn = reverse ? t - 2 : t + 2;
            tma_prefetch_3d(&maps.gates[d], j * 4 * kU, bt * kBM, tnn);
            tma_prefetch_3d(&maps.gates[d], j * 4 * kU + 4 * kUT, bt * kBM, tnn);
          }
        }
      }
      __syncwarp();
      if (s > 0) {
        for (int tries = 0;; ++tries) {
          mbar_wait_rec(smem_u32(acc_full), acc_round & 1u, p.trap_rec, 2, s, 0);
          ++acc_round;
          fence_after_sync();
          if (dbg && threadIdx.x == 0) dbg[s * kDbg + 3] = clock64();
          rec::tmem_ld_cols<4 * kUT>(taddr, acc);
          fence_before_sync();
          // a piece of h that had not arrived (fill pattern = NaN) makes every accumulator word of its row NaN
          const bool nan_row = (__float_as_uint(acc[0]) & 0x7fffffffu) > 0x7f800000u;
          // (E2T_REC_DBGSKIP bit 2: tests force a redo on every fifth step to exercise the path)
          const bool redo = (bar_red_or(3, kWorkThreads, nan_row) || ((p.dbg_skip & 4) && s % 5 == 0 && tries == 0)) && tries < kMaxRedo;
          if (threadIdx.x == 0) {
            *verdict = redo ? 1u : 0u;
            rec::mbar_arrive(smem_u32(verdict_bar));
          }
          __syncwarp();
          if (!redo) break;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4 * kUT; ++i) acc[i] = 0.f;
      }
      float hv[kUT];
      // (one reciprocal for the product of several denominators -- 7 instead of 10 SFU results per unit -- was measured:
      //  the epilogue of CTA 0 shrank by 270 cycles, the launch got 3 % SLOWER, gpurun_out/r2y; the plain form stays)
      if (valid) {
#pragma unroll
        for (int e = 0; e < kUT; ++e) {
          const float gi = rec::sigm(z[e] + acc[e]);
          const float gj = rec::tanh_fast(z[kUT + e] + acc[kUT + e]);
          const float gf = rec::sigm(z[2 * kUT + e] + acc[2 * kUT + e] + 1.0f);
          const float go = rec::sigm(z[3 * kUT + e] + acc[3 * kUT + e]);
          const float c = gf * carry[e] + gi * gj;
          carry[e] = c;
          hv[e] = go * rec::tanh_fast(c);
          z[e] = gi; z[kUT + e] = gj; z[2 * kUT + e] = gf; z[3 * kUT + e] = go;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4 * kUT; ++i) z[i] = 0.f;
#pragma unroll
        for (int e = 0; e < kUT; ++e) { carry[e] = 0.f; hv[e] = 0.f; }
      }
      // 1) what the other CTAs of the chain wait for: this thread's 8 units of h as ONE 16-byte store (zeros past the
      //    length, and for the padding rows of the last batch tile: every row of the tile must leave the fill pattern)
      {
        uint4 hp;
        hp.x = pack_h2(hv[0], hv[1]); hp.y = pack_h2(hv[2], hv[3]); hp.z = pack_h2(hv[4], hv[5]); hp.w = pack_h2(hv[6], hv[7]);
        st_relaxed_v4(hx_chain + (size_t)t * tile_stride + hx_off, hp);
      }
      if (dbg && threadIdx.x == 0) dbg[s * kDbg + 4] = clock64();
      // 2) what the backward pass and the next layer read.  Gate activations: 128 contiguous bytes per thread, straight from
      //    registers (the shared-memory tile they came from is already being refilled).  c / h / dropped h: small swizzled
      //    stages, written out by TMA bulk stores (rows past B are clipped by the tensor maps).
      if (row_ok && !(p.dbg_skip & 1)) rec::stv8<4 * kUT>(gates + ((i64)t * B + b) * 4 * H + z0, z);
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st_s0), "f"(carry[0]), "f"(carry[1]), "f"(carry[2]), "f"(carry[3]) : "memory");
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st_s1), "f"(carry[4]), "f"(carry[5]), "f"(carry[6]), "f"(carry[7]) : "memory");
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st_s0 + kStageSmall), "f"(hv[0]), "f"(hv[1]), "f"(hv[2]), "f"(hv[3]) : "memory");
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st_s1 + kStageSmall), "f"(hv[4]), "f"(hv[5]), "f"(hv[6]), "f"(hv[7]) : "memory");
      if (p.has_hd) {
        const uint32_t idx0 = (uint32_t)(((i64)t * B + b) * p.drop_F + col0 + u0);
        const uint32_t key = p.dp.key, thresh = p.dp.thresh;
        const float inv = p.dp.inv;
        float o[kUT];
#pragma unroll
        for (int e = 0; e < kUT; ++e) o[e] = (valid && e2t_keep(key, idx0 + e, thresh)) ? hv[e] * inv : 0.f;
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st_s0 + 2 * kStageSmall), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st_s1 + 2 * kStageSmall), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7]) : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      rec::named_bar_sync(4, kWorkThreads);
      if (threadIdx.x == 0) {
        if (dbg) dbg[s * kDbg + 9] = clock64();
        if (!(p.dbg_skip & 1)) {
          tma_store_3d(&maps.cs[d], so + kStageGates, j * kU, bt * kBM, t);
          tma_store_3d(&maps.hs, so + kStageGates + kStageSmall, col0 + j * kU, bt * kBM, t);
          if (p.has_hd) tma_store_3d(&maps.hd, so + kStageGates + 2 * kStageSmall, col0 + j * kU, bt * kBM, t);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (dbg) dbg[s * kDbg + 5] = clock64();
      }
      __syncwarp();
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 8) {
    fence_after_sync();
    tmem_dealloc(tmem_base, 128);
  }
}

// fp32 [rows, ld_src] -> fp16 [rows, ld_dst] (columns [0, cols); the padding columns up to ld_dst are zeroed), several
// matrices per launch: the 16-bit copies of Wh^T the recurrence multiplies with
struct PackJob { const float* src; __half* dst; int rows, cols, ld_src, ld_dst; };
struct PackJobs { PackJob j[16]; int n; };
__global__ void k_pack_f16(PackJobs jobs) {
  const PackJob& jb = jobs.j[blockIdx.y];
  if (((jb.cols | jb.ld_src | jb.ld_dst) & 3) == 0 && (reinterpret_cast<uintptr_t>(jb.src) & 15) == 0) {
    // four elements per thread: one 16-byte load, one 8-byte store
    const int q = jb.ld_dst / 4;
    const i64 n4 = (i64)jb.rows * q;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (i64)gridDim.x * blockDim.x) {
      const int r = (int)(i / q), c = (int)(i % q) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < jb.cols) v = *reinterpret_cast<const float4*>(jb.src + (i64)r * jb.ld_src + c);
      const __half2 lo = __floats2half2_rn(fminf(fmaxf(v.x, -65504.f), 65504.f), fminf(fmaxf(v.y, -65504.f), 65504.f));
      const __half2 hi = __floats2half2_rn(fminf(fmaxf(v.z, -65504.f), 65504.f), fminf(fmaxf(v.w, -65504.f), 65504.f));
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&lo); o.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(jb.dst + (i64)r * jb.ld_dst + c) = o;
    }
    return;
  }
  const i64 n = (i64)jb.rows * jb.ld_dst;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
    const int r = (int)(i / jb.ld_dst), c = (int)(i % jb.ld_dst);
    float v = c < jb.cols ? jb.src[(i64)r * jb.ld_src + c] : 0.f;
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    jb.dst[i] = __float2half_rn(v);
  }
}

// ---- host side ------------------------------------------------------------------------------------
inline CUtensorMap make_map_f16(const __half* ptr, i64 rows, i64 cols, i64 ld, int box_rows, int box_cols) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t bx[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  EncodeTiledFn fn = encode_fn();
  if (!fn) throw std::runtime_error("e2t: cuTensorMapEncodeTiled entry point not found");
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(ptr), gdim, gstr, bx, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("e2t: cuTensorMapEncodeTiled (f16) failed with code " + std::to_string((int)r));
  return m;
}

inline CUtensorMap make_map_f32_3d(const float* ptr, const i64* dims, const i64* strides_elems, const int* box, int swizzle_bytes) {
  CUtensorMap m;
  cuuint64_t gdim[3]; cuuint64_t gstr[2]; cuuint32_t bx[3]; cuuint32_t estr[3] = {1, 1, 1};
  for (int i = 0; i < 3; ++i) { gdim[i] = (cuuint64_t)dims[i]; bx[i] = (cuuint32_t)box[i]; }
  for (int i = 1; i < 3; ++i) gstr[i - 1] = (cuuint64_t)strides_elems[i] * 4;
  EncodeTiledFn fn = encode_fn();
  if (!fn) throw std::runtime_error("e2t: cuTensorMapEncodeTiled entry point not found");
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("e2t: cuTensorMapEncodeTiled (3d) failed with code " + std::to_string((int)r));
  return m;
}

inline int nkc16(int H) { return (H + kKC - 1) / kKC; }
inline size_t fwd16_smem_bytes(int H) {
  return (size_t)nkc16(H) * (kWChunk + kAChunk) + kStageBytes + (3 + 16) * 8 + 16 + 1024;
}
inline int hp16(int H) { return (H + 7) / 8 * 8; }
// exchange buffer: one shared-memory image of the operand tile (NKC chunks of 16 KB) per direction, step and batch tile
inline int bp16(int B) { return (B + kBM - 1) / kBM * kBM; }
inline size_t hx16_halves(int B, int H, int steps) { return (size_t)2 * steps * (bp16(B) / kBM) * nkc16(H) * (kAChunk / 2); }

inline bool fwd16_supported(int B, int H) {
  if (H % kU != 0 || H < kU || H > 512 || B < 1) return false;
  const int n_bt = (B + kBM - 1) / kBM, n_slices = H / kU;
  if (2 * n_bt * n_slices > rec::sm_count()) return false;
  if (n_slices * 8 > kWorkThreads) return false;            // one probe thread per (producer slice, warp)
  return fwd16_smem_bytes(H) <= 227 * 1024;
}

template <int NKC>
inline void fwd16_launch_t(cudaStream_t st, const Fwd16Maps& maps, Fwd16P& p) {
  auto kfn = k_lstm_fwd16<NKC>;
  const size_t smem = fwd16_smem_bytes(p.H);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    E2T_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * p.n_bt * p.n_slices)); cfg.blockDim = dim3(kThreads16);
  cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeCooperative; attrs[0].val.cooperative = 1;   // all CTAs co-resident
  static const bool coop = getenv("E2T_REC_NOCOOP") == nullptr;      // A/B: plain launch (the grid fits one wave by construction)
  cfg.attrs = attrs; cfg.numAttrs = coop ? 1 : 0;
  E2T_CHECK(cudaLaunchKernelEx(&cfg, kfn, maps, p));
}

// Forward of one BiLSTM layer.  WhT16[d]: fp16 copies of Wh^T [4H (permuted gate rows), Hp]; hx: exchange buffer of at
// least hx16_halves(B, H, steps) halves.
inline void rec_forward16(cudaStream_t st, float* const gates[2], float* const cs[2], float* hs, float* hd,
                          const __half* const WhT16[2], __half* hx, const int* lens2, int steps, int B, int H, DropP dp,
                          int drop_F, bool fill_hx = true) {
  Fwd16P p{};
  for (int d = 0; d < 2; ++d) p.gates[d] = gates[d];
  p.hx = hx; p.lens2 = lens2; p.has_hd = hd != nullptr;
  p.steps = steps; p.B = B; p.H = H;
  p.n_bt = (B + kBM - 1) / kBM; p.n_slices = H / kU;
  p.dp = dp; p.drop_F = drop_F;
  static const int skip = getenv("E2T_REC_DBGSKIP") ? atoi(getenv("E2T_REC_DBGSKIP")) : 0;
  p.dbg_skip = skip;
  Fwd16Maps maps;
  const i64 dg[3] = {4 * (i64)H, B, steps}, sg[3] = {1, 4 * (i64)H, (i64)B * 4 * H};
  const i64 dc[3] = {H, B, steps}, sc[3] = {1, H, (i64)B * H};
  const i64 dh[3] = {2 * (i64)H, B, steps}, sh[3] = {1, 2 * (i64)H, (i64)B * 2 * H};
  const int bg[3] = {4 * kUT, kBM, 1}, bs[3] = {kU, kBM, 1};
  for (int d = 0; d < 2; ++d) {
    maps.w[d] = make_map_f16(WhT16[d], 4 * (i64)H, H, hp16(H), 4 * kU, kKC);
    maps.gates[d] = make_map_f32_3d(gates[d], dg, sg, bg, 128);
    maps.cs[d] = make_map_f32_3d(cs[d], dc, sc, bs, 64);
  }
  maps.hs = make_map_f32_3d(hs, dh, sh, bs, 64);
  maps.hd = make_map_f32_3d(hd ? hd : hs, dh, sh, bs, 64);
  // (fill_hx = false: the caller keeps the buffer all-0xFF between launches, e2t.cu: XBuf)
  if (fill_hx) E2T_CHECK(cudaMemsetAsync(hx, 0xFF, hx16_halves(B, H, steps) * sizeof(__half), st));
  static int dbg_left = getenv("E2T_REC_DEBUG") ? atoi(getenv("E2T_REC_DEBUG")) : 0;
  p.dbg = nullptr;
  if (dbg_left > 0) {
    E2T_CHECK(cudaMalloc(&p.dbg, (size_t)(steps + 1) * kDbg * sizeof(long long)));
    E2T_CHECK(cudaMemsetAsync(p.dbg, 0, (size_t)(steps + 1) * kDbg * sizeof(long long), st));
  }
  // E2T_REC_TRAPINFO=1: a wait that times out leaves a record in mapped host memory before it traps (the launch is then
  // synchronised here so that the record can be printed: diagnostics only)
  static int* trap_host = nullptr;
  static int* trap_dev = nullptr;
  static const bool trapinfo = getenv("E2T_REC_TRAPINFO") != nullptr;
  if (trapinfo && !trap_host) {
    E2T_CHECK(cudaHostAlloc(&trap_host, 64, cudaHostAllocMapped));
    memset(trap_host, 0, 64);
    E2T_CHECK(cudaHostGetDevicePointer(&trap_dev, trap_host, 0));
  }
  p.trap_rec = trapinfo ? trap_dev : nullptr;
  switch (nkc16(H)) {
    case 1: fwd16_launch_t<1>(st, maps, p); break;
    case 2: fwd16_launch_t<2>(st, maps, p); break;
    case 3: fwd16_launch_t<3>(st, maps, p); break;
    case 4: fwd16_launch_t<4>(st, maps, p); break;
    case 5: fwd16_launch_t<5>(st, maps, p); break;
    case 6: fwd16_launch_t<6>(st, maps, p); break;
    case 7: fwd16_launch_t<7>(st, maps, p); break;
    default: throw std::runtime_error("e2t: rec_forward16 needs H <= 448");
  }
  if (trapinfo) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess)
      fprintf(stderr, "[rec fwd16 TRAP] %s: site=%d (1 probe, 2 acc_full, 3 a_full, 4 check, 5 weights, 6 probe_bar, 7 chk_bar) block=%d thread=%d step=%d chunk=%d extra=0x%x "
                      "(steps=%d B=%d H=%d has_hd=%d)\n", cudaGetErrorString(e), trap_host[0], trap_host[1], trap_host[2], trap_host[3],
              trap_host[4], (unsigned)trap_host[5], steps, B, H, p.has_hd);
    E2T_CHECK(e);
  }
  if (p.dbg) {
    --dbg_left;
    std::vector<long long> hst((size_t)(steps + 1) * kDbg);
    E2T_CHECK(cudaStreamSynchronize(st));
    E2T_CHECK(cudaMemcpy(hst.data(), p.dbg, hst.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(p.dbg);
    const int nkc = nkc16(H);
    fprintf(stderr, "[rec fwd16] steps=%d B=%d H=%d grid=%d (cycles of CTA 0, rel. to the request of the step's LAST chunk)\n"
                    "  step  repulls first_chunk_asked [chunk's MMAs issued ...] ->acc_seen ->h_stored ->staged ->stores_issued | step_total\n",
            steps, B, H, 2 * p.n_bt * p.n_slices);
    for (int s = 1; s + 1 < steps; ++s) {
      const long long* e = &hst[(size_t)s * kDbg];
      const long long prev = s > 1 ? hst[(size_t)(s - 1) * kDbg] : 0;
      fprintf(stderr, "  %4d  %4lld %6lld [", s, e[7], e[1] - e[0]);
      for (int kc = 0; kc < nkc; ++kc) fprintf(stderr, " %lld", e[16 + kc] - e[0]);
      fprintf(stderr, "] %6lld %6lld %6lld %6lld | %6lld\n", e[3] - e[0], e[4] - e[0], e[9] - e[0], e[5] - e[0], prev ? e[0] - prev : 0);
    }
  }
}

// ================================================================================================
// BPTT, second generation (k_lstm_bptt2): the reduce-scatter formulation of k_lstm_bptt (lstm_rec.cuh) -- every CTA multiplies
// only its own 64 dz columns with the matching columns of Wh for all H units and the [128, H] partial dh goes through an
// L2-resident workspace -- with the counter / release / acquire hand-off replaced by TAGGED DATA:
//   * every fp32 word of a partial carries, in its least-significant mantissa bit, the parity of the number of times its
//     workspace slot has been written (the slot of (step parity, direction, batch tile, writer) is rewritten every second
//     step).  A reader re-loads a 16-byte piece until all four tag bits show the value it expects; stale data from two steps
//     earlier carries the other value.  No flags, no fences, no resets; one bit (6e-8 relative) of the partial is given up.
//   * a writer drains its accumulator owner by owner in the order the owners sum (owner j-1 first), and an owner sums its
//     n partials starting with writer j+1: the hand-off pipelines, and the summation order is a fixed function of the
//     slice index (bit-identical gradients run to run).
// The workspace must hold consistent tags: the host memsets it whenever (batch tiles, H) change and tracks the write counts
// per step parity (BpttTags).
// ================================================================================================
struct BpttTags { uint32_t epoch[2]; int n_bt, H; };

struct Bptt2P {
  rec::RecBptt r;
  uint32_t epoch0, epoch1;       // writes so far to the parity-0 / parity-1 slots
};

__device__ __forceinline__ uint4 ld_relaxed_v4f(const float* p) { return ld_relaxed_v4(p); }
__device__ __forceinline__ bool tags_ok(const uint4& v, uint32_t tw) {
  return ((((v.x ^ tw) | (v.y ^ tw) | (v.z ^ tw) | (v.w ^ tw)) & 1u) == 0u);
}

__global__ void __launch_bounds__(rec::kBpttThreads, 1)
k_lstm_bptt2(const __grid_constant__ CUtensorMap map_w0, const __grid_constant__ CUtensorMap map_w1, Bptt2P pp) {
  using namespace rec;
  const RecBptt& p = pp.r;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int H = p.H, steps = p.steps, B = p.B;
  const uint32_t WCH = (uint32_t)H * 128;                 // one 32-column chunk of the resident weights [H rows x 128 B]
  unsigned char* smem_w = smem;                           // [2 chunks][H][32] K-major, 128B swizzle
  unsigned char* smem_a = smem + 2 * (size_t)WCH;         // [2 chunks][128][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + 2 * A_STAGE_BYTES);
  uint64_t* w_bar = bars;
  uint64_t* acc_full = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = p.n_slices;
  const int j = blockIdx.x % n;
  const int bt = (blockIdx.x / n) % p.n_bt;
  const int d = blockIdx.x / (n * p.n_bt);
  const bool reverse = d == 1;
  const CUtensorMap* map_w = d ? &map_w1 : &map_w0;
  long long* dbg = (blockIdx.x == 0) ? p.dbg : nullptr;

  if (warp == 0 && lane == 0) {
    mbar_init(smem_u32(w_bar), 1);
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
      const uint32_t wb = smem_u32(w_bar);
      mbar_expect_tx(wb, 2 * WCH);
      for (int c = 0; c < 2; ++c)
        for (int r0 = 0; r0 < H; r0 += p.wbox_rows)
          tma_load_2d(smem_u32(smem_w + (size_t)c * WCH + (size_t)r0 * 128), map_w, wb, j * 64 + c * 32, r0);
    }
    __syncwarp();
    mbar_wait(smem_u32(w_bar), 0);
  }
  const uint64_t desc_a0 = make_smem_desc(smem_u32(smem_a));
  const uint64_t desc_w0 = make_smem_desc(smem_u32(smem_w));

  const int quad = warp & 3;
  const int sg = warp >> 2;                        // unit sub-group (8 units)
  const int r = quad * 32 + lane;
  const int b = bt * kBM + r;
  const bool row_ok = b < B;
  const int u0 = j * kU + sg * kBUT;
  const int z0 = j * 4 * kU + sg * 4 * kBUT;
  const int len2 = row_ok ? (p.lens2 ? p.lens2[b] : steps) : 0;
  float* gates = d ? p.gates[1] : p.gates[0];
  const float* cs = d ? p.cs[1] : p.cs[0];
  const int col0 = d * H;
  const bool t0thread = threadIdx.x == 0;
  const size_t chain_sz = (size_t)n * kBM * H;
  float* pws_chain = p.pws + (size_t)(d * p.n_bt + bt) * chain_sz;
  const size_t par_stride = (size_t)2 * p.n_bt * chain_sz;
  const uint32_t a_row = smem_u32(smem_a) + (uint32_t)sg * A_STAGE_BYTES + (uint32_t)r * 128;
  const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
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
      ldv8<4 * kBUT>(gz, zrow);
      ldv8<kBUT>(cv, cs + ((i64)t * B + b) * H + u0);
      if (sf > 0) ldv8<kBUT>(cpv, cs + ((i64)tp * B + b) * H + u0);
      if (p.dhs) ldv8<kBUT>(dhv, p.dhs + ((i64)t * B + b) * 2 * H + col0 + u0);
    }
    float acc[kBUT];
#pragma unroll
    for (int i = 0; i < kBUT; ++i) acc[i] = 0.f;
    if (q > 0) {
      if (t0thread && dbg) dbg[q * 8 + 0] = clock64();
      if (row_ok) {
        // partial dh of step q-1 from every writer, summed in the order j+1, j+2, ..., j (mod n); two halves of <= 13
        // writers, all pieces of a half in flight together, re-polled until their tags say "written at step q-1".
        // Rows past their length poll too (and discard): seeing step q-1 of EVERY writer is what keeps this CTA from
        // running two steps ahead and overwriting a slot another owner has not read yet.
        const int pq = (q - 1) & 1;
        const uint32_t tw = ((pq ? pp.epoch1 : pp.epoch0) + (uint32_t)((q - 1) >> 1)) & 1u;
        const float* src = pws_chain + (size_t)pq * par_stride + ((size_t)(u0 / 4) * kBM + r) * 4;
        constexpr int KB = 13;
        const long long t0 = clock64();
        for (int k0 = 1; k0 <= n; k0 += KB) {
          uint4 v[KB][2];
          uint32_t pending = 0;
#pragma unroll
          for (int kk = 0; kk < KB; ++kk)
            if (k0 + kk <= n) pending |= 3u << (2 * kk);
          while (pending) {
#pragma unroll
            for (int kk = 0; kk < KB; ++kk) {
              int i = j + k0 + kk;
              i -= (i >= n) ? n : 0;
              i -= (i >= n) ? n : 0;
#pragma unroll
              for (int h4 = 0; h4 < 2; ++h4)
                if (pending & (1u << (2 * kk + h4)))
                  v[kk][h4] = ld_relaxed_v4f(src + (size_t)i * kBM * H + (size_t)h4 * kBM * 4);
            }
#pragma unroll
            for (int kk = 0; kk < KB; ++kk)
#pragma unroll
              for (int h4 = 0; h4 < 2; ++h4)
                if ((pending & (1u << (2 * kk + h4))) && tags_ok(v[kk][h4], tw)) pending &= ~(1u << (2 * kk + h4));
            if (pending && clock64() - t0 > 4000000000LL) __trap();
          }
#pragma unroll
          for (int kk = 0; kk < KB; ++kk)
            if (k0 + kk <= n) {
#pragma unroll
              for (int h4 = 0; h4 < 2; ++h4) {
                acc[4 * h4] += __uint_as_float(v[kk][h4].x & ~1u); acc[4 * h4 + 1] += __uint_as_float(v[kk][h4].y & ~1u);
                acc[4 * h4 + 2] += __uint_as_float(v[kk][h4].z & ~1u); acc[4 * h4 + 3] += __uint_as_float(v[kk][h4].w & ~1u);
              }
            }
        }
      }
      if (t0thread && dbg) dbg[q * 8 + 1] = clock64();
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
      // dz -> swizzled A tile (rows past B are zero), visible to the tensor core, then warp 0 issues the MMAs
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
        if (t0thread && dbg) dbg[q * 8 + 2] = clock64();
        if (elect_one()) {
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t da = desc_a0 + (uint64_t)((c * A_STAGE_BYTES + k * UMMA_K * 4) >> 4);
              for (int pt = 0; pt < p.n_parts; ++pt) {
                const int nn = min(p.part, H - pt * p.part);
                const uint64_t dw = desc_w0 + (uint64_t)((c * WCH + (uint32_t)(pt * p.part) * 128 + k * UMMA_K * 4) >> 4);
                umma_tf32(tmem_base + (uint32_t)(pt * p.part), da, dw, make_idesc_tf32(kBM, nn, 0), (c > 0 || k > 0) ? 1u : 0u);
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
      if (t0thread && dbg) dbg[q * 8 + 3] = clock64();
      // drain the accumulator owner by owner: owner o = j - k sums this writer at position k, so k = 1 goes out first;
      // this thread takes every second owner (k = 1 + sg, 3 + sg, ...); each owner = 16 columns = 4 unit quads
      const int pq = q & 1;
      const uint32_t tw = ((pq ? pp.epoch1 : pp.epoch0) + (uint32_t)(q >> 1)) & 1u;
      float* dstw = pws_chain + (size_t)pq * par_stride + (size_t)j * kBM * H + (size_t)r * 4;
      float va[16], vb[16];
      int k = 1 + sg;
      int o = j - k; o += (o < 0) ? n : 0;
      if (k <= n) tmem_ld16_nowait(tlane + (uint32_t)(16 * o), va);
      while (k <= n) {
        tmem_ld_wait();
        const int k2 = k + 2;
        int o2 = j - k2; o2 += (o2 < 0) ? n : 0; o2 += (o2 < 0) ? n : 0;
        if (k2 <= n) tmem_ld16_nowait(tlane + (uint32_t)(16 * o2), vb);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 w;
          w.x = (__float_as_uint(va[4 * g]) & ~1u) | tw; w.y = (__float_as_uint(va[4 * g + 1]) & ~1u) | tw;
          w.z = (__float_as_uint(va[4 * g + 2]) & ~1u) | tw; w.w = (__float_as_uint(va[4 * g + 3]) & ~1u) | tw;
          st_relaxed_v4(dstw + (size_t)(4 * o + g) * kBM * 4, w);
        }
        if (k2 > n) break;
        tmem_ld_wait();
        const int k3 = k2 + 2;
        int o3 = j - k3; o3 += (o3 < 0) ? n : 0; o3 += (o3 < 0) ? n : 0;
        if (k3 <= n) tmem_ld16_nowait(tlane + (uint32_t)(16 * o3), va);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 w;
          w.x = (__float_as_uint(vb[4 * g]) & ~1u) | tw; w.y = (__float_as_uint(vb[4 * g + 1]) & ~1u) | tw;
          w.z = (__float_as_uint(vb[4 * g + 2]) & ~1u) | tw; w.w = (__float_as_uint(vb[4 * g + 3]) & ~1u) | tw;
          st_relaxed_v4(dstw + (size_t)(4 * o2 + g) * kBM * 4, w);
        }
        k = k3; o = o3;
      }
      if (t0thread && dbg) dbg[q * 8 + 4] = clock64();
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// BPTT of one BiLSTM layer (second generation).  Arguments as rec::rec_backward_rs; `tags` is the persistent tag state of `pws`.
inline void rec_backward_rs2(cudaStream_t st, float* const gates[2], const float* const cs[2], const float* dhs,
                             const float* const K[2], int In, const int* lens2, const float* dc_inject, int ldi,
                             const int* inject_t, float* pws, size_t pws_floats, BpttTags& tags, int steps, int B, int H) {
  using namespace rec;
  Bptt2P pp{};
  RecBptt& p = pp.r;
  for (int d = 0; d < 2; ++d) { p.gates[d] = gates[d]; p.cs[d] = cs[d]; }
  p.dhs = dhs; p.lens2 = lens2; p.dc_inject = dc_inject; p.ldi = ldi; p.inject_t = inject_t; p.counters = nullptr;
  p.pws = pws; p.steps = steps; p.B = B; p.H = H;
  p.n_bt = (B + kBM - 1) / kBM; p.n_slices = H / kU;
  p.n_parts = (H + 255) / 256;
  p.part = (((H + p.n_parts - 1) / p.n_parts) + 15) / 16 * 16;
  CUtensorMap mw[2];
  for (int d = 0; d < 2; ++d) {
    const i64 wdims[2] = {4 * (i64)H, H}, wstr[2] = {1, 4 * (i64)H};
    p.wbox_rows = H <= 256 ? H : H / 2;
    const int wbox[2] = {BK, p.wbox_rows};
    mw[d] = make_map_nd(K[d] + (i64)In * 4 * H, 2, wdims, wstr, wbox);
  }
  if (tags.n_bt != p.n_bt || tags.H != H) {
    // new geometry: the slots hold tags of another layout -- start from all-zero tags, next expected tag = 1
    E2T_CHECK(cudaMemsetAsync(pws, 0, pws_floats * sizeof(float), st));
    tags.epoch[0] = tags.epoch[1] = 1;
    tags.n_bt = p.n_bt; tags.H = H;
  }
  pp.epoch0 = tags.epoch[0]; pp.epoch1 = tags.epoch[1];
  // steps - 1 partials are written (q = 0 .. steps-2): parity 0 gets ceil, parity 1 floor of (steps - 1) / 2
  tags.epoch[0] += (uint32_t)(steps / 2);
  tags.epoch[1] += (uint32_t)((steps - 1) / 2);
  auto kfn = k_lstm_bptt2;
  const size_t smem = bptt_smem_bytes(H);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    E2T_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  static int dbg_left = getenv("E2T_REC_DEBUG") ? atoi(getenv("E2T_REC_DEBUG")) : 0;
  p.dbg = nullptr;
  if (dbg_left > 0) {
    E2T_CHECK(cudaMalloc(&p.dbg, (size_t)steps * 8 * sizeof(long long)));
    E2T_CHECK(cudaMemsetAsync(p.dbg, 0, (size_t)steps * 8 * sizeof(long long), st));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * p.n_bt * p.n_slices)); cfg.blockDim = dim3(kBpttThreads);
  cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeCooperative; attrs[0].val.cooperative = 1;   // all CTAs co-resident
  cfg.attrs = attrs; cfg.numAttrs = 1;
  E2T_CHECK(cudaLaunchKernelEx(&cfg, kfn, mw[0], mw[1], pp));
  if (p.dbg) {
    --dbg_left;
    std::vector<long long> hst((size_t)steps * 8);
    E2T_CHECK(cudaStreamSynchronize(st));
    E2T_CHECK(cudaMemcpy(hst.data(), p.dbg, hst.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(p.dbg);
    fprintf(stderr, "[rec bptt2] steps=%d B=%d H=%d grid=%u (cycles of CTA 0 thread 0, rel. to the start of the step's poll)\n"
                    "  step  ->partials_summed  ->dz_in_smem  ->acc_seen  ->partial_stored | step_total\n",
            steps, B, H, cfg.gridDim.x);
    for (int q = 1; q + 1 < steps; ++q) {
      const long long* e = &hst[(size_t)q * 8];
      const long long prev = q > 1 ? hst[(size_t)(q - 1) * 8] : 0;
      fprintf(stderr, "  %4d  %8lld %8lld %8lld %8lld | %8lld\n", q, e[1] - e[0], e[2] - e[0], e[3] - e[0], e[4] - e[0],
              prev ? e[0] - prev : 0);
    }
  }
}

}  // namespace rec16
