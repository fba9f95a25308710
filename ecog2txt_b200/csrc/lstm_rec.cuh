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
//   backward: CTA owns the same 16 units -> N = 16 columns of dh_rec, K = 4H.
//             D[128 x 16] = dz_next[128 x 4H] * Wh_slice[16 x 4H]^T ; dc carried in registers.
// CTAs of one (direction, batch tile) chain exchange h / dz through L2 (their natural HBM buffers) and
// a per-step arrival counter: epilogue stores -> fence -> red.release ; producer ld.acquire spin ->
// fence.proxy.async -> TMA.  All CTAs must be co-resident: cooperative launch, grid <= #SMs.
//
// warp roles: 0 = TMA producer (+ counter wait), 1 = MMA issuer (+ TMEM owner), 2..5 = epilogue.
#pragma once
#include "gemm_tc.cuh"

namespace rec {

using namespace tc;

constexpr int kU = 16;            // hidden units per CTA
constexpr int kBM = 128;          // batch rows per CTA
constexpr int kThreadsRec = 192;
constexpr uint32_t A_STAGE_BYTES = kBM * BK * 4;   // 16 KB

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
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
  const long long t0 = clock64();
  while (ld_acquire_gpu(p) < target) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
template <int NCOL>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* v);
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

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }

struct RecFwdP {
  float* gates[2];        // [T', B, 4H] x-projection (+bias) in, gate activations out
  float* cs[2];           // [T', B, H]
  float* hs;              // [T', B, 2H]   fwd half | bwd half
  float* hd;              // dropped copy of hs (nullable)
  const int* lens2;       // [B] (nullable: all steps valid)
  int* counters;          // [2][n_bt][steps], zeroed before launch
  int steps, B, H, n_bt, n_slices, nkc, stages;
  DropP dp; int drop_F;
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
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + (size_t)p.stages * A_STAGE_BYTES);
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

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(w_bar), 1);
    mbar_init(smem_u32(acc_full), 1);
    mbar_init(smem_u32(acc_empty), 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), G::TMEM_COLS);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      // resident weights, once
      const uint32_t wb = smem_u32(w_bar);
      mbar_expect_tx(wb, (uint32_t)p.nkc * G::WCHUNK);
      for (int kc = 0; kc < p.nkc; ++kc) {
        if (!BWD) tma_load_3d(smem_u32(smem_w + (size_t)kc * G::WCHUNK), map_w, wb, kc * BK, j * kU, 0);
        else      tma_load_2d(smem_u32(smem_w + (size_t)kc * G::WCHUNK), map_w, wb, kc * BK, j * kU);
      }
      int it = 0;
      for (int s = 1; s < steps; ++s) {
        // forward pass : step s (forward order) consumes h of step s-1
        // backward pass: q-th processed step (q = s) consumes dz of the (q-1)-th processed step
        int t_src;
        if (!BWD) { const int t = reverse ? steps - 1 - s : s; t_src = reverse ? t + 1 : t - 1; }
        else      { const int sf = steps - 1 - s; const int t = reverse ? steps - 1 - sf : sf; t_src = reverse ? t - 1 : t + 1; }
        wait_counter(counters + (s - 1), p.n_slices);
        fence_proxy_async_all();
        for (int kc = 0; kc < p.nkc; ++kc, ++it) {
          const int st = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(smem_u32(&empty_bar[st]), ph ^ 1);
          const uint32_t fb = smem_u32(&full_bar[st]);
          mbar_expect_tx(fb, A_STAGE_BYTES);
          tma_load_2d(smem_u32(smem_a + (size_t)st * A_STAGE_BYTES), map_a, fb, kc * BK, t_src * B + bt * kBM);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(kBM, G::N, 0);
      mbar_wait(smem_u32(w_bar), 0);
      fence_after_sync();
      int it = 0;
      for (int s = 1; s < steps; ++s) {
        mbar_wait(smem_u32(acc_empty), ((s - 1) & 1) ^ 1);   // epilogue drained the previous accumulator
        fence_after_sync();
        for (int kc = 0; kc < p.nkc; ++kc, ++it) {
          const int st = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(smem_u32(&full_bar[st]), ph);
          fence_after_sync();
          const uint32_t sa = smem_u32(smem_a + (size_t)st * A_STAGE_BYTES);
          const uint32_t sw = smem_u32(smem_w + (size_t)kc * G::WCHUNK);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_tf32(tmem_base, make_smem_desc(sa + k * UMMA_K * 4), make_smem_desc(sw + k * UMMA_K * 4), idesc,
                      (kc > 0 || k > 0) ? 1u : 0u);
          umma_commit(smem_u32(&empty_bar[st]));
        }
        umma_commit(smem_u32(acc_full));
      }
    }
  } else {
    // ================= epilogue: thread = batch row, 16 hidden units =================
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may read
    const int r = quad * 32 + lane;
    const int b = bt * kBM + r;
    const bool row_ok = b < B;
    const int u0 = j * kU;
    const int len2 = row_ok ? (p.lens2 ? p.lens2[b] : steps) : 0;
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    float carry[kU];                                 // forward: cell state ; backward: dc through time
#pragma unroll
    for (int i = 0; i < kU; ++i) carry[i] = 0.f;

    if constexpr (!BWD) {
      float* gates = p.gates[d];
      float* cs = p.cs[d];
      const int col0 = d * H;
      for (int s = 0; s < steps; ++s) {
        const int t = reverse ? steps - 1 - s : s;
        const bool valid = row_ok && t < len2;
        float* zrow = gates + ((i64)t * B + b) * 4 * H + u0;
        float4 zx[4][kU / 4];
        if (valid) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int q = 0; q < kU / 4; ++q) zx[g][q] = *reinterpret_cast<const float4*>(zrow + (i64)g * H + q * 4);
        }
        float acc[4 * kU];
        if (s > 0) {
          mbar_wait(smem_u32(acc_full), (s - 1) & 1);
          fence_after_sync();
          tmem_ld_cols<4 * kU>(taddr, acc);
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(acc_empty));
        } else {
#pragma unroll
          for (int i = 0; i < 4 * kU; ++i) acc[i] = 0.f;
        }
        if (row_ok) {
          float hv[kU];
          if (valid) {
#pragma unroll
            for (int q = 0; q < kU / 4; ++q) {
              float zi[4] = {zx[0][q].x, zx[0][q].y, zx[0][q].z, zx[0][q].w};
              float zj[4] = {zx[1][q].x, zx[1][q].y, zx[1][q].z, zx[1][q].w};
              float zf[4] = {zx[2][q].x, zx[2][q].y, zx[2][q].z, zx[2][q].w};
              float zo[4] = {zx[3][q].x, zx[3][q].y, zx[3][q].z, zx[3][q].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int u = q * 4 + e;
                const float gi = sigm(zi[e] + acc[u]);
                const float gj = tanhf(zj[e] + acc[kU + u]);
                const float gf = sigm(zf[e] + acc[2 * kU + u] + 1.0f);
                const float go = sigm(zo[e] + acc[3 * kU + u]);
                const float c = gf * carry[u] + gi * gj;
                carry[u] = c;
                hv[u] = go * tanhf(c);
                zi[e] = gi; zj[e] = gj; zf[e] = gf; zo[e] = go;
              }
              *reinterpret_cast<float4*>(zrow + q * 4) = make_float4(zi[0], zi[1], zi[2], zi[3]);
              *reinterpret_cast<float4*>(zrow + (i64)H + q * 4) = make_float4(zj[0], zj[1], zj[2], zj[3]);
              *reinterpret_cast<float4*>(zrow + (i64)2 * H + q * 4) = make_float4(zf[0], zf[1], zf[2], zf[3]);
              *reinterpret_cast<float4*>(zrow + (i64)3 * H + q * 4) = make_float4(zo[0], zo[1], zo[2], zo[3]);
            }
          } else {
#pragma unroll
            for (int u = 0; u < kU; ++u) { carry[u] = 0.f; hv[u] = 0.f; }
          }
          float* crow = cs + ((i64)t * B + b) * H + u0;
          float* hrow = p.hs + ((i64)t * B + b) * 2 * H + col0 + u0;
#pragma unroll
          for (int q = 0; q < kU / 4; ++q) {
            *reinterpret_cast<float4*>(crow + q * 4) = make_float4(carry[q * 4], carry[q * 4 + 1], carry[q * 4 + 2], carry[q * 4 + 3]);
            *reinterpret_cast<float4*>(hrow + q * 4) = make_float4(hv[q * 4], hv[q * 4 + 1], hv[q * 4 + 2], hv[q * 4 + 3]);
          }
          if (p.hd) {
            float* drow = p.hd + ((i64)t * B + b) * 2 * H + col0 + u0;
            const uint32_t idx0 = (uint32_t)(((i64)t * B + b) * p.drop_F + col0 + u0);
#pragma unroll
            for (int u = 0; u < kU; ++u)
              drow[u] = (valid && e2t_keep(p.dp.key, idx0 + u, p.dp.thresh)) ? hv[u] * p.dp.inv : 0.f;
          }
        }
        // publish: generic-proxy stores -> visible to the async proxy (TMA) of every CTA in the chain
        fence_proxy_async_all();
        named_bar_sync(1, 128);
        if (threadIdx.x == 64) {
          __threadfence();
          red_release_gpu(counters + s, 1);
        }
      }
    } else {
      float* gates = p.gates[d];
      const float* cs = p.cs[d];
      const int col0 = d * H;
      for (int q = 0; q < steps; ++q) {
        const int sf = steps - 1 - q;                       // forward-order step index being differentiated
        const int t = reverse ? steps - 1 - sf : sf;
        const int tp = reverse ? t + 1 : t - 1;             // step processed before t in the forward pass
        const bool valid = row_ok && t < len2;
        float* zrow = gates + ((i64)t * B + b) * 4 * H + u0;
        float4 gz[4][kU / 4], cc[kU / 4], cp[kU / 4], dho[kU / 4];
        if (valid) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int w = 0; w < kU / 4; ++w) gz[g][w] = *reinterpret_cast<const float4*>(zrow + (i64)g * H + w * 4);
          const float* crow = cs + ((i64)t * B + b) * H + u0;
#pragma unroll
          for (int w = 0; w < kU / 4; ++w) {
            cc[w] = *reinterpret_cast<const float4*>(crow + w * 4);
            cp[w] = sf > 0 ? *reinterpret_cast<const float4*>(cs + ((i64)tp * B + b) * H + u0 + w * 4)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
            dho[w] = p.dhs ? *reinterpret_cast<const float4*>(p.dhs + ((i64)t * B + b) * 2 * H + col0 + u0 + w * 4)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        float acc[kU];
        if (q > 0) {
          mbar_wait(smem_u32(acc_full), (q - 1) & 1);
          fence_after_sync();
          tmem_ld_cols<kU>(taddr, acc);
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(acc_empty));
        } else {
#pragma unroll
          for (int i = 0; i < kU; ++i) acc[i] = 0.f;
        }
        if (row_ok) {
          if (valid) {
            bool inject = false;
            if (p.dc_inject) {
              const int ti = (d == 0 && p.inject_t) ? p.inject_t[b] : 0;
              inject = ti == t;
            }
#pragma unroll
            for (int w = 0; w < kU / 4; ++w) {
              const float gi[4] = {gz[0][w].x, gz[0][w].y, gz[0][w].z, gz[0][w].w};
              const float gj[4] = {gz[1][w].x, gz[1][w].y, gz[1][w].z, gz[1][w].w};
              const float gf[4] = {gz[2][w].x, gz[2][w].y, gz[2][w].z, gz[2][w].w};
              const float go[4] = {gz[3][w].x, gz[3][w].y, gz[3][w].z, gz[3][w].w};
              const float cv[4] = {cc[w].x, cc[w].y, cc[w].z, cc[w].w};
              const float cpv[4] = {cp[w].x, cp[w].y, cp[w].z, cp[w].w};
              const float dhv[4] = {dho[w].x, dho[w].y, dho[w].z, dho[w].w};
              float di[4], dj[4], df[4], dO[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int u = w * 4 + e;
                const float dh = dhv[e] + acc[u];
                float dc = carry[u];
                if (inject) dc += p.dc_inject[(i64)b * p.ldi + col0 + u0 + u];
                const float tc_ = tanhf(cv[e]);
                dO[e] = dh * tc_ * go[e] * (1.f - go[e]);
                dc += dh * go[e] * (1.f - tc_ * tc_);
                di[e] = dc * gj[e] * gi[e] * (1.f - gi[e]);
                dj[e] = dc * gi[e] * (1.f - gj[e] * gj[e]);
                df[e] = dc * cpv[e] * gf[e] * (1.f - gf[e]);
                carry[u] = dc * gf[e];
              }
              *reinterpret_cast<float4*>(zrow + w * 4) = make_float4(di[0], di[1], di[2], di[3]);
              *reinterpret_cast<float4*>(zrow + (i64)H + w * 4) = make_float4(dj[0], dj[1], dj[2], dj[3]);
              *reinterpret_cast<float4*>(zrow + (i64)2 * H + w * 4) = make_float4(df[0], df[1], df[2], df[3]);
              *reinterpret_cast<float4*>(zrow + (i64)3 * H + w * 4) = make_float4(dO[0], dO[1], dO[2], dO[3]);
            }
          } else {
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
              for (int w = 0; w < kU / 4; ++w) *reinterpret_cast<float4*>(zrow + (i64)g * H + w * 4) = z4;
#pragma unroll
            for (int u = 0; u < kU; ++u) carry[u] = 0.f;
          }
        }
        fence_proxy_async_all();
        named_bar_sync(1, 128);
        if (threadIdx.x == 64) {
          __threadfence();
          red_release_gpu(counters + q, 1);
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
  return (size_t)nkc * Geo<BWD>::WCHUNK + (size_t)stages * A_STAGE_BYTES + (2 * stages + 4) * 8 + 1024;
}

// picks the A-ring depth that fits; returns 0 if the resident weights do not fit at all
template <bool BWD>
inline int rec_pick_stages(int nkc) {
  const size_t cap = 227 * 1024;
  for (int s = 8; s >= 2; --s)
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
  void* args[] = {(void*)&a0, (void*)&a1, (void*)&w0, (void*)&w1, (void*)&p};
  dim3 grid((unsigned)(2 * p.n_bt * p.n_slices));
  E2T_CHECK(cudaLaunchCooperativeKernel((const void*)kfn, grid, dim3(kThreadsRec), args, smem, st));
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
    // W: Wh^T viewed as (k, u, gate): element (k,u,g) at KT[(g*H+u)*ldkt + In + k]
    const i64 wdims[3] = {H, H, 4}, wstr[3] = {1, (i64)ldkt, (i64)H * ldkt};
    const int wbox[3] = {BK, kU, 4};
    mw[d] = make_map_nd(KT[d] + In, 3, wdims, wstr, wbox);
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

}  // namespace rec
