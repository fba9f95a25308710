// lstm_dec16.cuh -- persistent teacher-forced decoder recurrence (A8 of SURVEY.md section 8a): ONE launch runs the L steps of
// the decoder LSTM (Hd = 800 in config 2) that used to be L x (K-concatenated GEMM + gate kernel) = 22 launches at their
// ~20 us latency floor.  Same exchange as k_lstm_fwd16 (lstm_rec16.cuh): grid = batch tiles of 128 rows x Hd/16 unit slices,
// the CTA's 64 rows of Wh^T stay in shared memory (fp16), every step's h travels as 16-byte fp16 pieces through an L2-resident
// tile image pre-filled with the NaN pattern 0xFFFF, a load warp polls one word per producer warp and pulls the chunks whose
// producers are in, a missing piece poisons the accumulator row with NaN and the step is pulled again.  Differences:
//   * Hd = 800 makes the operand tile 13 chunks x 16 KB = 208 KB: it does not fit next to 104 KB of weights, so the chunks
//     stream through a ring of 4 slots (slot freed by a tcgen05.commit behind the chunk's MMAs);
//   * canonical gate order [i | j | f | o] x Hd (no permuted copies of the decoder weights / biases / gradients): the CTA's
//     64 weight rows are four TMA boxes of 16 rows, the accumulator columns are [gate][16 units], the x-projection tile is
//     four [128 x 16] boxes, gate activations go out as four 32-byte pieces per thread;
//   * one direction, no utterance lengths, the first step multiplies the bridge state h0 (written as tile 0 by the prologue)
//     and starts from the cell state c0.
// The x-projection (embedding Wx + b) of all L steps is ONE GEMM in front of the launch (lstm_xproj), as in the encoder.
#pragma once
#include "lstm_rec16.cuh"
#include "lstm_bptt3.cuh"

namespace rec16 {

constexpr int kDecRing = 4;

struct Dec16Maps {
  CUtensorMap w;          // Wh^T fp16 [4Hd (canonical gate rows), Hp], box 64 x 16, 128B swizzle (load, once)
  CUtensorMap z;          // [L, B, 4Hd] fp32, box 16 x 128 x 1, 64B swizzle (load)
  CUtensorMap cs, hs;     // [L, B, Hd] fp32, box 16 x 128 x 1, 64B swizzle (store)
};

struct Dec16P {
  float* z;               // [L, B, 4Hd] x-projection (+bias) in, gate activations out
  const float* h0;        // [B, Hd]
  const float* c0;        // [B, Hd]
  unsigned char* hx;      // exchange: [L + 1][n_bt][NKC] x 16 KB chunk images, pre-filled with 0xFF
  int L, B, H, n_bt, n_slices, nkc;
  int dbg_force;
  long long* dbg;
  int* trap_rec;
};

template <int NKC>
__global__ void __launch_bounds__(kThreads16, 1)
k_dec_fwd16(const __grid_constant__ Dec16Maps maps, Dec16P p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* smem_w = smem;                                   // [NKC][64 rows = 4 gates x 16 units][128 B]
  unsigned char* smem_a = smem + (size_t)NKC * kWChunk;           // [kDecRing][128 rows][128 B]
  unsigned char* smem_z = smem_a + (size_t)kDecRing * kAChunk;    // x-projection: 4 gate tiles [128][16 fp32], 64B swizzle
  unsigned char* smem_o = smem_z + 4 * kPiece;                    // stages cs | hs
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_o + 2 * kPiece);
  uint64_t* w_bar = bars;
  uint64_t* acc_full = bars + 1;
  uint64_t* z_bar = bars + 2;
  uint64_t* verdict_bar = bars + 3;
  uint64_t* a_full = bars + 4;                                    // [kDecRing]
  uint64_t* a_free = bars + 4 + kDecRing;                         // [kDecRing]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 + 2 * kDecRing);
  volatile uint32_t* verdict = tmem_slot + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x % p.n_slices;
  const int bt = blockIdx.x / p.n_slices;
  const int L = p.L, B = p.B, H = p.H;
  long long* dbg = (blockIdx.x == 0) ? p.dbg : nullptr;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(w_bar), 1);
    mbar_init(smem_u32(acc_full), 1);
    mbar_init(smem_u32(z_bar), 1);
    mbar_init(smem_u32(verdict_bar), 1);
    for (int k = 0; k < kDecRing; ++k) { mbar_init(smem_u32(&a_full[k]), 1); mbar_init(smem_u32(&a_free[k]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 8) tmem_alloc(smem_u32(tmem_slot), 64);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, *tmem_slot);
  // tile of step s (= h of step s - 1; tile 0 = h0) at hx_chain + s * tile_stride
  const size_t tile_bytes = (size_t)NKC * kAChunk;
  const size_t tile_stride = (size_t)p.n_bt * tile_bytes;
  unsigned char* hx_chain = p.hx + (size_t)bt * tile_bytes;

  if (warp == 8) {
    // ================= MMA warp =================
    if (elect_one()) {
      const uint32_t wb = smem_u32(w_bar);
      mbar_expect_tx(wb, (uint32_t)NKC * kWChunk);
      for (int kc = 0; kc < NKC; ++kc)
        for (int g = 0; g < 4; ++g)      // 16 rows of gate g: units 16 j .. 16 j + 15; the K tail is zero-filled
          tma_load_2d(smem_u32(smem_w + (size_t)kc * kWChunk + (size_t)g * 2048), &maps.w, wb, kc * kKC, g * H + j * kU);
    }
    __syncwarp();
    mbar_wait_rec(smem_u32(w_bar), 0, p.trap_rec, 5, 0, 0);
    fence_after_sync();
    constexpr uint32_t idesc = make_idesc_f16(kBM, 4 * kU);
    const uint64_t desc_a0 = make_smem_desc(smem_u32(smem_a));
    const uint64_t desc_w0 = make_smem_desc(smem_u32(smem_w));
    uint32_t q = 0, round = 0;                         // chunks consumed so far (ring position), tiles so far (verdict phase)
    for (int s = 0; s < L; ++s) {
      for (;;) {
        for (int kc = 0; kc < NKC; ++kc, ++q) {
          const uint32_t slot = q % kDecRing;
          mbar_wait_rec(smem_u32(&a_full[slot]), (q / kDecRing) & 1u, p.trap_rec, 7, s, kc);
          fence_after_sync();
          if (elect_one()) {
            const int nk = min(4, (H - kc * kKC) / 16);
            for (int k = 0; k < nk; ++k)
              umma_f16(tmem_base, desc_a0 + (uint64_t)((slot * kAChunk + k * 32) >> 4), desc_w0 + (uint64_t)((kc * kWChunk + k * 32) >> 4),
                       idesc, (kc > 0 || k > 0) ? 1u : 0u);
            umma_commit(smem_u32(&a_free[slot]));      // the slot may be refilled once these MMAs have read it
            if (kc == NKC - 1) umma_commit(smem_u32(acc_full));
          }
          __syncwarp();
        }
        if (dbg && lane == 0) dbg[s * kDbg + 2] = clock64();
        mbar_wait_rec(smem_u32(verdict_bar), round & 1u, p.trap_rec, 9, s, 0);
        ++round;
        if (*verdict == 0u) break;
      }
    }
  } else if (warp == 9) {
    // ================= load warp: probe, chunks through the ring in order =================
    // probe word 32 k + lane: producer slice (32 k + lane) / 8, its epilogue warp % 8 -- the four slices of chunk k
    constexpr int kProbes = NKC;
    const int n_probe = p.n_slices * 8;
    uint32_t probe_off[kProbes];
#pragma unroll
    for (int k = 0; k < kProbes; ++k) {
      const int idx = 32 * k + lane;
      const int prow = 32 * (idx & 3), pk = (idx >> 3) * kU + ((idx >> 2) & 1) * kUT;
      probe_off[k] = (uint32_t)(pk / kKC) * kAChunk + (uint32_t)prow * 128 + (uint32_t)((((pk % kKC) / 8) ^ (prow & 7)) << 4);
    }
    uint32_t q = 0, round = 0;
    for (int s = 0; s < L; ++s) {
      const unsigned char* base = hx_chain + (size_t)s * tile_stride;
      for (bool first = true;; first = false) {
        uint32_t pending = 0;
        if (first) {
#pragma unroll
          for (int k = 0; k < kProbes; ++k)
            if (32 * k + lane < n_probe) pending |= 1u << k;
        }
        int next = 0;                                   // next chunk to pull (in order: the ring is consumed in order)
        const long long t0 = clock64();
        while (next < NKC) {
          uint32_t v[kProbes];
#pragma unroll
          for (int k = 0; k < kProbes; ++k)
            if (pending & (1u << k)) v[k] = ld_cg_u32(base + probe_off[k]);
#pragma unroll
          for (int k = 0; k < kProbes; ++k)
            if ((pending & (1u << k)) && v[k] != kFill32) pending &= ~(1u << k);
          const uint32_t ready = ~__reduce_or_sync(0xffffffffu, pending);       // bit k: every lane has seen its word of chunk k
          while (next < NKC && (ready & (1u << next))) {
            const uint32_t slot = q % kDecRing;
            if (q >= (uint32_t)kDecRing) mbar_wait_rec(smem_u32(&a_free[slot]), ((q / kDecRing) - 1u) & 1u, p.trap_rec, 3, s, next);
            if (dbg && lane == 0 && next == 0) dbg[s * kDbg + 1] = clock64();
            if (elect_one()) {
              const uint32_t fb = smem_u32(&a_full[slot]);
              mbar_expect_tx(fb, kAChunk);
              bulk_load(smem_u32(smem_a + (size_t)slot * kAChunk), base + (size_t)next * kAChunk, kAChunk, fb);
            }
            __syncwarp();
            ++next; ++q;
          }
          if (clock64() - t0 > (1LL << 31)) timeout_trap(p.trap_rec, 1, s, next, (int)pending);
        }
        if (dbg && lane == 0 && first) dbg[s * kDbg + 0] = clock64();
        mbar_wait_rec(smem_u32(verdict_bar), round & 1u, p.trap_rec, 6, s, 0);
        ++round;
        if (*verdict == 0u) break;
        if (dbg && lane == 0) dbg[s * kDbg + 7] += 1;
      }
    }
  } else {
    // ================= epilogue: thread = (batch row, 8 hidden units) =================
    const int quad = warp & 3;
    const int ug = warp >> 2;
    const int r = quad * 32 + lane;
    const int b = bt * kBM + r;
    const bool row_ok = b < B;
    const int u0 = j * kU + ug * kUT;
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ug * kUT);        // + 16 g
    const uint32_t sw0 = (uint32_t)r * 64 + (uint32_t)(((2 * ug) ^ ((r >> 1) & 3)) << 4);
    const uint32_t sw1 = (uint32_t)r * 64 + (uint32_t)(((2 * ug + 1) ^ ((r >> 1) & 3)) << 4);
    const uint32_t sz = smem_u32(smem_z), so = smem_u32(smem_o);
    const uint32_t hx_off = (uint32_t)(u0 / kKC) * kAChunk + (uint32_t)r * 128 + (uint32_t)((((u0 % kKC) / 8) ^ (r & 7)) << 4);
    const uint32_t zb = smem_u32(z_bar);
    uint32_t acc_round = 0;
    float carry[kUT];

    if (threadIdx.x == 0) {                            // x-projection tiles of the first step
      mbar_expect_tx(zb, 4 * kPiece);
      for (int g = 0; g < 4; ++g) rec::tma_load_3d(sz + (uint32_t)g * kPiece, &maps.z, zb, g * H + j * kU, bt * kBM, 0);
    }
    {
      // tile 0 = the bridge state h0 (zeros for the padding rows: the fill must go), cell state from c0
      float h0v[kUT];
#pragma unroll
      for (int e = 0; e < kUT; ++e) { h0v[e] = 0.f; carry[e] = 0.f; }
      if (row_ok) {
        rec::ldv8<kUT>(h0v, p.h0 + (i64)b * H + u0);
        rec::ldv8<kUT>(carry, p.c0 + (i64)b * H + u0);
      }
      uint4 hp;
      hp.x = pack_h2(h0v[0], h0v[1]); hp.y = pack_h2(h0v[2], h0v[3]); hp.z = pack_h2(h0v[4], h0v[5]); hp.w = pack_h2(h0v[6], h0v[7]);
      st_relaxed_v4(hx_chain + hx_off, hp);
    }
    for (int s = 0; s < L; ++s) {
      float z[4 * kUT], acc[4 * kUT];
      mbar_wait_rec(zb, (uint32_t)s & 1u, p.trap_rec, 8, s, 0);
#pragma unroll
      for (int g = 0; g < 4; ++g) { lds_v4(z + 8 * g, sz + (uint32_t)g * kPiece + sw0); lds_v4(z + 8 * g + 4, sz + (uint32_t)g * kPiece + sw1); }
      if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      rec::named_bar_sync(2, kWorkThreads);
      if (threadIdx.x == 0 && s + 1 < L) {
        mbar_expect_tx(zb, 4 * kPiece);
        for (int g = 0; g < 4; ++g) rec::tma_load_3d(sz + (uint32_t)g * kPiece, &maps.z, zb, g * H + j * kU, bt * kBM, s + 1);
      }
      __syncwarp();
      for (int tries = 0;; ++tries) {
        mbar_wait_rec(smem_u32(acc_full), acc_round & 1u, p.trap_rec, 2, s, 0);
        ++acc_round;
        fence_after_sync();
        if (dbg && threadIdx.x == 0) dbg[s * kDbg + 3] = clock64();
#pragma unroll
        for (int g = 0; g < 4; ++g) rec::tmem_ld_cols<8>(tlane + (uint32_t)(16 * g), acc + 8 * g);
        fence_before_sync();
        const bool nan_row = (__float_as_uint(acc[0]) & 0x7fffffffu) > 0x7f800000u;
        const bool force = p.dbg_force && s % 5 == 1 && tries == 0;
        const bool redo = (bar_red_or(3, kWorkThreads, nan_row) || force) && tries < kMaxRedo;
        if (threadIdx.x == 0) {
          *verdict = redo ? 1u : 0u;
          rec::mbar_arrive(smem_u32(verdict_bar));
        }
        __syncwarp();
        if (!redo) break;
      }
      float hv[kUT];
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
      if (s + 1 < L) {     // the next step's operand: this thread's 8 units as one 16-byte piece of tile s + 1
        uint4 hp;
        hp.x = pack_h2(hv[0], hv[1]); hp.y = pack_h2(hv[2], hv[3]); hp.z = pack_h2(hv[4], hv[5]); hp.w = pack_h2(hv[6], hv[7]);
        st_relaxed_v4(hx_chain + (size_t)(s + 1) * tile_stride + hx_off, hp);
      }
      if (dbg && threadIdx.x == 0) dbg[s * kDbg + 4] = clock64();
      // gate activations for the backward pass (canonical layout: four 32-byte pieces), c / h through the swizzled stages
      if (row_ok) {
        float* zrow = p.z + ((i64)s * B + b) * 4 * H + u0;
#pragma unroll
        for (int g = 0; g < 4; ++g) rec::stv8<kUT>(zrow + (i64)g * H, z + 8 * g);
      }
      sts_v4(so + sw0, carry); sts_v4(so + sw1, carry + 4);
      sts_v4(so + kPiece + sw0, hv); sts_v4(so + kPiece + sw1, hv + 4);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      rec::named_bar_sync(4, kWorkThreads);
      if (threadIdx.x == 0) {
        tma_store_3d(&maps.cs, so, j * kU, bt * kBM, s);
        tma_store_3d(&maps.hs, so + kPiece, j * kU, bt * kBM, s);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      __syncwarp();
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 8) {
    fence_after_sync();
    tmem_dealloc(tmem_base, 64);
  }
}

// ---- host side ------------------------------------------------------------------------------------
inline size_t dec16_smem_bytes(int H) {
  return (size_t)nkc16(H) * kWChunk + (size_t)kDecRing * kAChunk + 6 * kPiece + (4 + 2 * kDecRing) * 8 + 16 + 1024;
}
inline size_t dec16_hx_bytes(int B, int H, int L) { return (size_t)(L + 1) * (bp16(B) / kBM) * nkc16(H) * kAChunk; }
inline bool dec16_supported(int B, int H) {
  if (H % kU != 0 || H < 4 * kU || B < 1) return false;
  const int nkc = nkc16(H);
  if (nkc < 8 || nkc > 13) return false;                 // smaller decoders: the per-step kernels (instantiations 8..13 below)
  const int n_bt = (B + kBM - 1) / kBM, n_slices = H / kU;
  if (n_bt * n_slices > rec::sm_count()) return false;
  return dec16_smem_bytes(H) <= 227 * 1024;
}

template <int NKC>
inline void dec16_launch_t(cudaStream_t st, const Dec16Maps& maps, Dec16P& p) {
  auto kfn = k_dec_fwd16<NKC>;
  const size_t smem = dec16_smem_bytes(p.H);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    E2T_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(p.n_bt * p.n_slices)); cfg.blockDim = dim3(kThreads16);
  cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeCooperative; attrs[0].val.cooperative = 1;   // all CTAs co-resident
  cfg.attrs = attrs; cfg.numAttrs = 1;
  E2T_CHECK(cudaLaunchKernelEx(&cfg, kfn, maps, p));
}

// Teacher-forced decoder forward.  z: [L, B, 4Hd] x-projection + bias in, gate activations out (canonical gate order);
// WhT16: fp16 Wh^T [4Hd, hp16(Hd)]; hx: >= dec16_hx_bytes, ALL 0xFF on entry (the caller wipes it behind the launch).
inline void dec_forward16(cudaStream_t st, float* z, float* cs, float* hs, const __half* WhT16, unsigned char* hx, const float* h0,
                          const float* c0, int L, int B, int H) {
  Dec16P p{};
  p.z = z; p.h0 = h0; p.c0 = c0; p.hx = hx;
  p.L = L; p.B = B; p.H = H; p.n_bt = (B + kBM - 1) / kBM; p.n_slices = H / kU; p.nkc = nkc16(H);
  static const int force = getenv("E2T_REC_DBGSKIP") ? (atoi(getenv("E2T_REC_DBGSKIP")) & 4) : 0;
  p.dbg_force = force;
  Dec16Maps maps;
  maps.w = make_map_f16(WhT16, 4 * (i64)H, H, hp16(H), kU, kKC);
  const i64 dz[3] = {4 * (i64)H, B, L}, sz[3] = {1, 4 * (i64)H, (i64)B * 4 * H};
  const i64 dc[3] = {H, B, L}, sc[3] = {1, H, (i64)B * H};
  const int bs[3] = {kU, kBM, 1};
  maps.z = make_map_f32_3d(z, dz, sz, bs, 64);
  maps.cs = make_map_f32_3d(cs, dc, sc, bs, 64);
  maps.hs = make_map_f32_3d(hs, dc, sc, bs, 64);
  static int dbg_left = getenv("E2T_REC_DEBUG") ? atoi(getenv("E2T_REC_DEBUG")) : 0;
  p.dbg = nullptr;
  if (dbg_left > 0) {
    E2T_CHECK(cudaMalloc(&p.dbg, (size_t)(L + 1) * kDbg * sizeof(long long)));
    E2T_CHECK(cudaMemsetAsync(p.dbg, 0, (size_t)(L + 1) * kDbg * sizeof(long long), st));
  }
  static int* trap_host = nullptr;
  static int* trap_dev = nullptr;
  static const bool trapinfo = getenv("E2T_REC_TRAPINFO") != nullptr;
  if (trapinfo && !trap_host) {
    E2T_CHECK(cudaHostAlloc(&trap_host, 64, cudaHostAllocMapped));
    memset(trap_host, 0, 64);
    E2T_CHECK(cudaHostGetDevicePointer(&trap_dev, trap_host, 0));
  }
  p.trap_rec = trapinfo ? trap_dev : nullptr;
  switch (p.nkc) {
    case 8: dec16_launch_t<8>(st, maps, p); break;
    case 9: dec16_launch_t<9>(st, maps, p); break;
    case 10: dec16_launch_t<10>(st, maps, p); break;
    case 11: dec16_launch_t<11>(st, maps, p); break;
    case 12: dec16_launch_t<12>(st, maps, p); break;
    case 13: dec16_launch_t<13>(st, maps, p); break;
    default: throw std::runtime_error("e2t: dec_forward16 needs 8 <= ceil(Hd / 64) <= 13");
  }
  if (trapinfo) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess)
      fprintf(stderr, "[dec fwd16 TRAP] %s: site=%d (1 probe, 2 acc_full, 3 a_free, 5 weights, 6/9 verdict, 7 a_full, 8 z_bar) block=%d thread=%d "
                      "step=%d chunk=%d extra=0x%x (L=%d B=%d H=%d)\n", cudaGetErrorString(e), trap_host[0], trap_host[1], trap_host[2],
              trap_host[3], trap_host[4], (unsigned)trap_host[5], L, B, H);
    E2T_CHECK(e);
  }
  if (p.dbg) {
    --dbg_left;
    std::vector<long long> hst((size_t)(L + 1) * kDbg);
    E2T_CHECK(cudaStreamSynchronize(st));
    E2T_CHECK(cudaMemcpy(hst.data(), p.dbg, hst.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(p.dbg);
    fprintf(stderr, "[dec fwd16] L=%d B=%d H=%d grid=%d (cycles of CTA 0, rel. to the request of the step's LAST chunk)\n"
                    "  step  repulls first_chunk_asked ->mma_issued ->acc_seen ->h_stored | step_total\n", L, B, H, p.n_bt * p.n_slices);
    for (int s = 1; s < L; ++s) {
      const long long* e = &hst[(size_t)s * kDbg];
      const long long prev = hst[(size_t)(s - 1) * kDbg];
      fprintf(stderr, "  %4d  %4lld %8lld %8lld %8lld %8lld | %8lld\n", s, e[7], e[1] - e[0], e[2] - e[0], e[3] - e[0], e[4] - e[0], e[0] - prev);
    }
  }
}


// ================================================================================================
// k_dec_bwd16 -- BPTT of the teacher-forced decoder LSTM in ONE launch (was L x (dz Wh^T GEMM at its ~20 us floor: 400
// dependent MMAs of K = 8 + gate-derivative kernel) + one more GEMM for the bridge-state gradient).
// All-gather formulation on the k_dec_fwd16 skeleton: every CTA keeps Wh[its 16 units, all 4Hd gate columns] (fp16, 100 KB)
// in shared memory and per step streams the [128, 4Hd] dz tile of the previous step (fp16 x 2^e, 50 chunks of 16 KB at
// Hd = 800) through the 4-slot ring: 200 MMAs of N = 16, K = 16 into a [128 x 16] accumulator = the recurrent part of dh for
// its own units.  (The two-hop exchange of k_lstm_bptt3 needs 102 KB of weights + 80 KB of chunks + 80 KB of pieces at
// n = 50 slices: it does not fit; with only L = 11 steps the simpler exchange is what the time buys.)
// Hand-off exactly as in the forward kernels: 16-byte pieces into per-step slots pre-filled with the fp16 NaN pattern, the
// load warp polls one word per (piece column, row quadrant) of the next ring-depth chunks, a missing piece poisons the
// accumulator row and the tile is pulled again.  Round L (no gate math) multiplies dz of the first step: the gradient of
// the bridge state, dh0; dc0 is the cell-gradient carry.
// ================================================================================================
struct DecBwdMaps {
  CUtensorMap w;          // Wh fp16 [Hd units, 4Hd canonical gate columns], box 64 x 16, 128B swizzle (load, once)
  CUtensorMap g;          // [L, B, 4Hd] fp32 gate activations, box 16 x 128 x 1, 64B swizzle (load)
  CUtensorMap cs;         // [L, B, Hd] fp32, box 16 x 128 x 1, 64B swizzle (load)
  CUtensorMap c0;         // [1, B, Hd]
  CUtensorMap dh;         // [L, B, Hd] gradient from the layers above
};

struct DecBwdP {
  float* gates;           // [L, B, 4Hd] gate activations in, dz out (in place)
  float* dh0;             // [B, Hd] out
  float* dc0;             // [B, Hd] out
  unsigned char* dzx;     // exchange: [L][n_bt][4Hd / 64] x 16 KB chunk images, pre-filled with 0xFF
  const int* scale_in;    // bits of the largest |incoming gradient| (k_absmax)
  int L, B, H, n_bt, n_slices, nkc;
  int dbg_force;
  long long* dbg;
  int* trap_rec;
};

constexpr uint32_t kDecBwdW = 16 * 128;      // one 64-column chunk of the resident weights: 16 unit rows x 128 B
constexpr int kBwdPair = 2;                  // chunks per ring slot of k_dec_bwd16
constexpr int kBwdSlots = kDecRing / kBwdPair;

__global__ void __launch_bounds__(kThreads16, 1)
k_dec_bwd16(const __grid_constant__ DecBwdMaps maps, DecBwdP p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int NKC = p.nkc;
  unsigned char* smem_w = smem;                                   // [NKC][16 rows][128 B]
  unsigned char* smem_a = smem + (size_t)NKC * kDecBwdW;          // [kDecRing][128 rows][128 B]
  unsigned char* smem_in = smem_a + (size_t)kDecRing * kAChunk;   // 4 gate tiles | c(t-1) | dh from above | c(t) of the first step
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_in + 7 * kPiece);
  uint64_t* w_bar = bars;
  uint64_t* acc_full = bars + 1;
  uint64_t* in_bar = bars + 2;
  uint64_t* verdict_bar = bars + 3;
  uint64_t* a_full = bars + 4;
  uint64_t* a_free = bars + 4 + kDecRing;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 + 2 * kDecRing);
  volatile uint32_t* verdict = tmem_slot + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x % p.n_slices;
  const int bt = blockIdx.x / p.n_slices;
  const int L = p.L, B = p.B, H = p.H;
  long long* dbg = (blockIdx.x == 0) ? p.dbg : nullptr;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(w_bar), 1);
    mbar_init(smem_u32(acc_full), 1);
    mbar_init(smem_u32(in_bar), 1);
    mbar_init(smem_u32(verdict_bar), 1);
    for (int k = 0; k < kDecRing; ++k) { mbar_init(smem_u32(&a_full[k]), 1); mbar_init(smem_u32(&a_free[k]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 8) tmem_alloc(smem_u32(tmem_slot), 32);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, *tmem_slot);
  // dz tile of BPTT step q at dz_chain + q * tile_stride; piece (row, k) at (k / 64) * 16 KB + row * 128 + (((k % 64) / 8) ^ (row & 7)) * 16
  const size_t tile_bytes = (size_t)NKC * kAChunk;
  const size_t tile_stride = (size_t)p.n_bt * tile_bytes;
  unsigned char* dz_chain = p.dzx + (size_t)bt * tile_bytes;

  if (warp == 8) {
    // ================= MMA warp =================
    if (elect_one()) {
      const uint32_t wb = smem_u32(w_bar);
      mbar_expect_tx(wb, (uint32_t)NKC * kDecBwdW);
      for (int kc = 0; kc < NKC; ++kc)
        tma_load_2d(smem_u32(smem_w + (size_t)kc * kDecBwdW), &maps.w, wb, kc * kKC, j * kU);
    }
    __syncwarp();
    mbar_wait_rec(smem_u32(w_bar), 0, p.trap_rec, 5, 0, 0);
    fence_after_sync();
    constexpr uint32_t idesc = make_idesc_f16(kBM, kU);
    const uint64_t desc_a0 = make_smem_desc(smem_u32(smem_a));
    const uint64_t desc_w0 = make_smem_desc(smem_u32(smem_w));
    // the ring is handed over in slots of kBwdPair chunks (one barrier round trip, fence and commit per 8 MMAs instead of
    // per 4: the per-chunk overhead was ~150 of ~570 cycles)
    uint32_t q = 0, round = 0;
    for (int s = 0; s < L; ++s) {
      for (;;) {
        for (int kc = 0; kc < NKC; kc += kBwdPair, ++q) {
          const uint32_t slot = q % kBwdSlots;
          mbar_wait_rec(smem_u32(&a_full[slot]), (q / kBwdSlots) & 1u, p.trap_rec, 7, s, kc);
          fence_after_sync();
          if (elect_one()) {
#pragma unroll
            for (int c2 = 0; c2 < kBwdPair; ++c2)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16(tmem_base, desc_a0 + (uint64_t)(((slot * kBwdPair + c2) * kAChunk + k * 32) >> 4),
                         desc_w0 + (uint64_t)(((kc + c2) * kDecBwdW + k * 32) >> 4), idesc, (kc > 0 || c2 > 0 || k > 0) ? 1u : 0u);
            umma_commit(smem_u32(&a_free[slot]));
            if (kc + kBwdPair >= NKC) umma_commit(smem_u32(acc_full));
          }
          __syncwarp();
        }
        if (dbg && lane == 0) dbg[s * kDbg + 2] = clock64();
        mbar_wait_rec(smem_u32(verdict_bar), round & 1u, p.trap_rec, 9, s, 0);
        ++round;
        if (*verdict == 0u) break;
      }
    }
  } else if (warp == 9) {
    // ================= load warp: chunks of the tile of round s through the ring, in order =================
    // lane l watches piece l & 7 of row 32 (l >> 3) of a chunk: the first word of a producer warp's lane-0 store
    const uint32_t lane_off = (uint32_t)(lane >> 3) * 4096 + (uint32_t)(lane & 7) * 16;
    uint32_t q = 0, round = 0;
    for (int s = 0; s < L; ++s) {
      const unsigned char* base = dz_chain + (size_t)s * tile_stride;
      for (bool first = true;; first = false) {
        int next = 0;              // next chunk to pull
        int seen_upto = first ? 0 : NKC;      // chunks [0, seen_upto) are known complete (after a verdict: all)
        const long long t0 = clock64();
        while (next < NKC) {
          // poll the window [seen_upto, min(NKC, next + ring depth in chunks))
          const int hi = min(NKC, next + kDecRing);
          uint32_t v[kDecRing];
#pragma unroll
          for (int k = 0; k < kDecRing; ++k)
            if (seen_upto + k < hi) v[k] = ld_cg_u32(base + (size_t)(seen_upto + k) * kAChunk + lane_off);
          int adv = 0;
#pragma unroll
          for (int k = 0; k < kDecRing; ++k)
            if (seen_upto + k < hi && adv == k && __all_sync(0xffffffffu, v[k] != kFill32)) adv = k + 1;
          seen_upto += adv;
          while (next + kBwdPair <= seen_upto) {          // a slot = kBwdPair consecutive chunks = one contiguous 32 KB copy
            const uint32_t slot = q % kBwdSlots;
            if (q >= (uint32_t)kBwdSlots) mbar_wait_rec(smem_u32(&a_free[slot]), ((q / kBwdSlots) - 1u) & 1u, p.trap_rec, 3, s, next);
            if (dbg && lane == 0 && next == 0) dbg[s * kDbg + 1] = clock64();
            if (elect_one()) {
              const uint32_t fb = smem_u32(&a_full[slot]);
              mbar_expect_tx(fb, kBwdPair * kAChunk);
              bulk_load(smem_u32(smem_a + (size_t)slot * kBwdPair * kAChunk), base + (size_t)next * kAChunk, kBwdPair * kAChunk, fb);
            }
            __syncwarp();
            next += kBwdPair; ++q;
          }
          if (clock64() - t0 > (1LL << 31)) timeout_trap(p.trap_rec, 1, s, next, seen_upto);
        }
        if (dbg && lane == 0 && first) dbg[s * kDbg + 0] = clock64();
        mbar_wait_rec(smem_u32(verdict_bar), round & 1u, p.trap_rec, 6, s, 0);
        ++round;
        if (*verdict == 0u) break;
        if (dbg && lane == 0) dbg[s * kDbg + 7] += 1;
      }
    }
  } else {
    // ================= compute threads: (batch row, 8 hidden units) =================
    const int quad = warp & 3;
    const int ug = warp >> 2;
    const int r = quad * 32 + lane;
    const int b = bt * kBM + r;
    const bool row_ok = b < B;
    const int u0 = j * kU + ug * kUT;
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ug * kUT);
    const uint32_t sw0 = (uint32_t)r * 64 + (uint32_t)(((2 * ug) ^ ((r >> 1) & 3)) << 4);
    const uint32_t sw1 = (uint32_t)r * 64 + (uint32_t)(((2 * ug + 1) ^ ((r >> 1) & 3)) << 4);
    const uint32_t sin = smem_u32(smem_in);
    const uint32_t ib = smem_u32(in_bar);
    const int e_scale = bptt3_scale_exp(p.scale_in);
    const float S = exp2f((float)e_scale), invS = exp2f((float)-e_scale);
    uint32_t acc_round = 0;
    float carry[kUT], cv[kUT];
#pragma unroll
    for (int i = 0; i < kUT; ++i) carry[i] = 0.f;

    auto request_inputs = [&](int t, bool with_ct) {      // thread 0: the tiles of time step t
      mbar_expect_tx(ib, (uint32_t)(with_ct ? 7 : 6) * kPiece);
      for (int g = 0; g < 4; ++g) rec::tma_load_3d(sin + (uint32_t)g * kPiece, &maps.g, ib, g * H + j * kU, bt * kBM, t);
      if (t > 0) rec::tma_load_3d(sin + 4 * kPiece, &maps.cs, ib, j * kU, bt * kBM, t - 1);
      else rec::tma_load_3d(sin + 4 * kPiece, &maps.c0, ib, j * kU, bt * kBM, 0);
      rec::tma_load_3d(sin + 5 * kPiece, &maps.dh, ib, j * kU, bt * kBM, t);
      if (with_ct) rec::tma_load_3d(sin + 6 * kPiece, &maps.cs, ib, j * kU, bt * kBM, t);
    };
    if (threadIdx.x == 0) request_inputs(L - 1, true);

    for (int q = 0; q <= L; ++q) {
      const int t = L - 1 - q;
      float acc[kUT];
#pragma unroll
      for (int e = 0; e < kUT; ++e) acc[e] = 0.f;
      float gz[4 * kUT], cpv[kUT], dhv[kUT];
      if (q < L) {
        mbar_wait_rec(ib, (uint32_t)q & 1u, p.trap_rec, 8, q, 0);
#pragma unroll
        for (int g = 0; g < 4; ++g) { lds_v4(gz + 8 * g, sin + (uint32_t)g * kPiece + sw0); lds_v4(gz + 8 * g + 4, sin + (uint32_t)g * kPiece + sw1); }
        lds_v4(cpv, sin + 4 * kPiece + sw0); lds_v4(cpv + 4, sin + 4 * kPiece + sw1);
        lds_v4(dhv, sin + 5 * kPiece + sw0); lds_v4(dhv + 4, sin + 5 * kPiece + sw1);
        if (q == 0) { lds_v4(cv, sin + 6 * kPiece + sw0); lds_v4(cv + 4, sin + 6 * kPiece + sw1); }
        __syncwarp();
        rec::named_bar_sync(2, kWorkThreads);
        if (threadIdx.x == 0 && q + 1 < L) request_inputs(t - 1, false);
        __syncwarp();
      }
      if (q > 0) {       // recurrent part of dh: dz of round q - 1 times Wh^T, this thread's 8 units
        for (int tries = 0;; ++tries) {
          mbar_wait_rec(smem_u32(acc_full), acc_round & 1u, p.trap_rec, 2, q, 0);
          ++acc_round;
          fence_after_sync();
          if (dbg && threadIdx.x == 0) dbg[(q - 1) * kDbg + 3] = clock64();
          rec::tmem_ld_cols<8>(tlane, acc);
          fence_before_sync();
          const bool nan_row = (__float_as_uint(acc[0]) & 0x7fffffffu) > 0x7f800000u;
          const bool force = p.dbg_force && q % 5 == 3 && tries == 0;
          const bool redo = (bar_red_or(3, kWorkThreads, nan_row) || force) && tries < kMaxRedo;
          if (threadIdx.x == 0) {
            *verdict = redo ? 1u : 0u;
            rec::mbar_arrive(smem_u32(verdict_bar));
          }
          __syncwarp();
          if (!redo) break;
        }
#pragma unroll
        for (int e = 0; e < kUT; ++e) acc[e] *= invS;
      }
      if (q == L) {      // gradient of the bridge state
        if (row_ok) {
          rec::stv8<kUT>(p.dh0 + (i64)b * H + u0, acc);
          rec::stv8<kUT>(p.dc0 + (i64)b * H + u0, carry);
        }
        break;
      }
#pragma unroll
      for (int e = 0; e < kUT; ++e) {
        const float gi = gz[e], gj = gz[kUT + e], gf = gz[2 * kUT + e], go = gz[3 * kUT + e];
        const float dh = dhv[e] + acc[e];
        float dc = carry[e];
        const float tc_ = rec::tanh_fast(cv[e]);
        gz[3 * kUT + e] = dh * tc_ * go * (1.f - go);
        dc += dh * go * (1.f - tc_ * tc_);
        gz[e] = dc * gj * gi * (1.f - gi);
        gz[kUT + e] = dc * gi * (1.f - gj * gj);
        gz[2 * kUT + e] = dc * cpv[e] * gf * (1.f - gf);
        carry[e] = dc * gf;
      }
      if (!row_ok) {
#pragma unroll
        for (int i = 0; i < 4 * kUT; ++i) gz[i] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < kUT; ++e) cv[e] = cpv[e];
      // this thread's four 16-byte pieces of the dz tile of round q (gate g: K index g Hd + u0)
      {
        unsigned char* dst = dz_chain + (size_t)q * tile_stride + (uint32_t)r * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int kidx = g * H + u0;
          uint4 w;
          w.x = pack_h2_sat(gz[8 * g] * S, gz[8 * g + 1] * S); w.y = pack_h2_sat(gz[8 * g + 2] * S, gz[8 * g + 3] * S);
          w.z = pack_h2_sat(gz[8 * g + 4] * S, gz[8 * g + 5] * S); w.w = pack_h2_sat(gz[8 * g + 6] * S, gz[8 * g + 7] * S);
          st_relaxed_v4(dst + (size_t)(kidx / kKC) * kAChunk + (uint32_t)((((kidx % kKC) / 8) ^ (r & 7)) << 4), w);
        }
        if (dbg && threadIdx.x == 0) dbg[q * kDbg + 4] = clock64();
      }
      if (row_ok) {      // dz for the weight-gradient GEMMs, canonical layout
        float* zrow = p.gates + ((i64)t * B + b) * 4 * H + u0;
#pragma unroll
        for (int g = 0; g < 4; ++g) rec::stv8<kUT>(zrow + (i64)g * H, gz + 8 * g);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 8) {
    fence_after_sync();
    tmem_dealloc(tmem_base, 32);
  }
}

inline size_t decbwd_smem_bytes(int H) {
  return (size_t)(4 * H / kKC) * kDecBwdW + (size_t)kDecRing * kAChunk + 7 * kPiece + (4 + 2 * kDecRing) * 8 + 16 + 1024;
}
inline size_t decbwd_dzx_bytes(int B, int H, int L) { return (size_t)L * (bp16(B) / kBM) * (4 * H / kKC) * kAChunk; }
inline bool decbwd_supported(int B, int H) {
  if (H % kU != 0 || H < 8 * kU || B < 1 || (4 * H / kKC) % kBwdPair != 0) return false;
  const int n_bt = (B + kBM - 1) / kBM, n_slices = H / kU;
  if (n_bt * n_slices > rec::sm_count()) return false;
  return decbwd_smem_bytes(H) <= 227 * 1024;
}

// BPTT of the teacher-forced decoder.  gates: activations in, dz out; Wh16: fp16 Wh [Hd, 4Hd]; dhs: gradient from above
// [L, B, Hd]; dzx: >= decbwd_dzx_bytes, ALL 0xFF on entry; scale_ws: 4 ints of device scratch.
inline void dec_backward16(cudaStream_t st, float* gates, const float* cs, const float* c0, const float* dhs, const __half* Wh16,
                           unsigned char* dzx, int* scale_ws, float* dh0, float* dc0, int L, int B, int H) {
  DecBwdP p{};
  p.gates = gates; p.dh0 = dh0; p.dc0 = dc0; p.dzx = dzx; p.scale_in = scale_ws;
  p.L = L; p.B = B; p.H = H; p.n_bt = (B + kBM - 1) / kBM; p.n_slices = H / kU; p.nkc = 4 * H / kKC;
  static const int force = getenv("E2T_REC_DBGSKIP") ? (atoi(getenv("E2T_REC_DBGSKIP")) & 4) : 0;
  p.dbg_force = force;
  DecBwdMaps maps;
  maps.w = make_map_f16(Wh16, H, 4 * (i64)H, 4 * (i64)H, kU, kKC);
  const i64 dg[3] = {4 * (i64)H, B, L}, sg[3] = {1, 4 * (i64)H, (i64)B * 4 * H};
  const i64 dc[3] = {H, B, L}, sc[3] = {1, H, (i64)B * H};
  const i64 d0[3] = {H, B, 1};
  const int bs[3] = {kU, kBM, 1};
  maps.g = make_map_f32_3d(gates, dg, sg, bs, 64);
  maps.cs = make_map_f32_3d(cs, dc, sc, bs, 64);
  maps.c0 = make_map_f32_3d(c0, d0, sc, bs, 64);
  maps.dh = make_map_f32_3d(dhs, dc, sc, bs, 64);
  E2T_CHECK(cudaMemsetAsync(scale_ws, 0, 4 * sizeof(int), st));
  {
    const long long na = (long long)L * B * H;
    if (na & 3) throw std::runtime_error("e2t: dec_backward16 needs 4-float aligned gradient buffers");
    k_absmax<<<148, 256, 0, st>>>(reinterpret_cast<const float4*>(dhs), na / 4, nullptr, 0, scale_ws);
    E2T_CHECK(cudaGetLastError());
  }
  static int dbg_left = getenv("E2T_REC_DEBUG") ? atoi(getenv("E2T_REC_DEBUG")) : 0;
  p.dbg = nullptr;
  if (dbg_left > 0) {
    E2T_CHECK(cudaMalloc(&p.dbg, (size_t)(L + 2) * kDbg * sizeof(long long)));
    E2T_CHECK(cudaMemsetAsync(p.dbg, 0, (size_t)(L + 2) * kDbg * sizeof(long long), st));
  }
  static int* trap_host = nullptr;
  static int* trap_dev = nullptr;
  static const bool trapinfo = getenv("E2T_REC_TRAPINFO") != nullptr;
  if (trapinfo && !trap_host) {
    E2T_CHECK(cudaHostAlloc(&trap_host, 64, cudaHostAllocMapped));
    memset(trap_host, 0, 64);
    E2T_CHECK(cudaHostGetDevicePointer(&trap_dev, trap_host, 0));
  }
  p.trap_rec = trapinfo ? trap_dev : nullptr;
  auto kfn = k_dec_bwd16;
  const size_t smem = decbwd_smem_bytes(H);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    E2T_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(p.n_bt * p.n_slices)); cfg.blockDim = dim3(kThreads16);
  cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeCooperative; attrs[0].val.cooperative = 1;   // all CTAs co-resident
  cfg.attrs = attrs; cfg.numAttrs = 1;
  E2T_CHECK(cudaLaunchKernelEx(&cfg, kfn, maps, p));
  if (trapinfo) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess)
      fprintf(stderr, "[dec bwd16 TRAP] %s: site=%d (1 probe, 2 acc_full, 3 a_free, 5 weights, 6/9 verdict, 7 a_full, 8 in_bar) block=%d thread=%d "
                      "step=%d chunk=%d extra=0x%x (L=%d B=%d H=%d)\n", cudaGetErrorString(e), trap_host[0], trap_host[1], trap_host[2],
              trap_host[3], trap_host[4], (unsigned)trap_host[5], L, B, H);
    E2T_CHECK(e);
  }
  if (p.dbg) {
    --dbg_left;
    std::vector<long long> hst((size_t)(L + 2) * kDbg);
    E2T_CHECK(cudaStreamSynchronize(st));
    E2T_CHECK(cudaMemcpy(hst.data(), p.dbg, hst.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(p.dbg);
    fprintf(stderr, "[dec bwd16] L=%d B=%d H=%d grid=%d chunks/tile=%d (cycles of CTA 0, rel. to the request of the round's LAST chunk)\n"
                    "  round  repulls first_chunk_asked ->mma_issued ->acc_seen ->dz_published(next round) | round_total\n", L, B, H,
            p.n_bt * p.n_slices, p.nkc);
    for (int s = 1; s < L; ++s) {
      const long long* e = &hst[(size_t)s * kDbg];
      const long long prev = hst[(size_t)(s - 1) * kDbg];
      fprintf(stderr, "  %4d  %4lld %8lld %8lld %8lld %8lld | %8lld\n", s, e[7], e[1] - e[0], e[2] - e[0], e[3] - e[0], e[kDbg + 4] - e[0], e[0] - prev);
    }
  }
}

}  // namespace rec16
