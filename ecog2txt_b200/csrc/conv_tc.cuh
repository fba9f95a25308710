// conv_tc.cuh -- the temporal convolution (A3 + A4 of SURVEY.md section 8a) on tcgen05, HBM-bound by design.
//
// The reference reverses every utterance within its length (tf.reverse_sequence, trainers.py:808-810) and applies a
// stride-W, width-W conv2d (trainers.py:813-818; kernel [1, W, C, E], plotters.py:511-514).  Both are ONE gather-GEMM here:
//   forward   Y[(t2,b), e]   = sum_{w,c} x[b, len_b-1-(t2 W + w), c] * Wc[w, c, e]          (M = T2*B, K = W*C, N = E)
//   backward  dWc[(w,c), e]  = sum_{t2,b} x[b, len_b-1-(t2 W + w), c] * dY[(t2,b), e]      (M = W*C,  K = T2*B, N = E)
// The ECoG tensor x [B, T, C] is read exactly once per kernel with 16-byte loads (512 B contiguous per row in the
// backward pass, 128 B in the forward pass); the reversal and the zero padding of the last window are index
// arithmetic in the gather warps, which write the operand tile straight into the 128B-swizzled shared-memory layout
// the tensor core expects (K-major SWIZZLE_128B forward, MN-major SWIZZLE_128B_BASE32B backward), then
// fence.proxy.async + mbarrier.  The small operand (weights / dY) arrives by TMA.  Split-K over the otherwise idle SMs
// with a fixed-order reduction (k_splitk_reduce) keeps the result deterministic.
//
// warp roles: 0 = TMA producer (B operand), 1 = MMA issuer (+TMEM owner), 2-5 = epilogue, 6-21 = gather (4 groups of 4 warps).
#pragma once
#include <type_traits>

#include "gemm_tc.cuh"

namespace conv {

using namespace tc;

constexpr int kGatherWarps = 16;
constexpr int kGatherThreads = 32 * kGatherWarps;          // 512: two 16-byte pieces of every 16 KB A tile each
constexpr int kConvThreads = 64 + 128 + kGatherThreads;    // 704
constexpr int kGatherGroups = 4;                            // independent groups of 4 gather warps, one k-chunk each in flight

struct ConvP {
  const float* x; const int* lens;
  int Bsz, T, C, W, T2;
  int M, N, K;                  // GEMM view (see header)
  int BN, stages, ksplit, chunks_per_split, n_chunks;
  int groups;                   // independent gather groups (<= stages: a group runs at most one ring phase ahead)
  float* out; i64 ldo;          // ksplit == 1: result (+bias) ; else unused
  const float* bias;
  float* ws;                    // [ksplit][M][N] partials
};

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void sts_f4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// PRECISE (forward only): fp32-accurate products from three tf32 MMAs per k-step (a = a_hi + a_lo, b = b_hi + b_lo,
// a b ~ a_hi b_hi + a_hi b_lo + a_lo b_hi; the hi parts are tf32-exact so the tensor core's operand conversion cannot
// disturb them).  The conv output feeds a ReLU: with plain tf32 products pre-activations within ~1e-3 of zero come out on
// the other side of the kink than in the fp32 reference, and one flipped row changes a whole column of dWc by ~1/sqrt(K).
template <bool BWD, bool PRECISE>
__global__ void __launch_bounds__(kConvThreads, 1)
k_conv_tc(const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_b_lo, ConvP p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // stage layout: [A (hi)] [A lo (PRECISE)] [B (hi)] [B lo (PRECISE)]
  constexpr uint32_t NP = PRECISE ? 2 : 1;
  const uint32_t A_BYTES = BM * BK * 4, B_BYTES = (uint32_t)p.BN * BK * 4, STAGE_BYTES = NP * (A_BYTES + B_BYTES);
  const uint32_t B_OFF = NP * A_BYTES;
  float* stage_c = reinterpret_cast<float*>(smem + (size_t)p.stages * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_c + 4 * 32 * kEpiPad);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + p.stages;
  uint64_t* acc_full = bars + 2 * p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&full_bar[s]), kGatherWarps / p.groups + 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(acc_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 256);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, *tmem_slot);

  // work item: blockIdx.x = ks * tiles_m + tm
  const int tiles_m = (p.M + BM - 1) / BM;
  const int tm = blockIdx.x % tiles_m, ks = blockIdx.x / tiles_m;
  const int m0 = tm * BM;
  const int j0 = ks * p.chunks_per_split, j1 = min(j0 + p.chunks_per_split, p.n_chunks);
  const int cpf = p.C / 32;       // 32-channel chunks per frame

  if (warp == 0) {
    // ================= TMA producer: the small operand =================
    for (int j = j0, it = 0; j < j1; ++j, ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (it / p.stages) & 1;
      mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
      if (elect_one()) {
        const uint32_t fb = smem_u32(&full_bar[s]);
        mbar_expect_tx(fb, NP * B_BYTES);
        const uint32_t sb = smem_u32(smem + (size_t)s * STAGE_BYTES) + B_OFF;
        if (!BWD) {
          tma_load_2d(sb, &map_b, fb, j * BK, 0);                               // Wc^T [E, W*C]: k columns of chunk j
          if (PRECISE) tma_load_2d(sb + B_BYTES, &map_b_lo, fb, j * BK, 0);
        } else for (int i = 0; i < p.BN / 32; ++i) tma_load_2d(sb + i * (BK * 128), &map_b, fb, i * 32, j * BK);   // dY [K, E]
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = make_idesc_tf32(BM, p.BN, BWD ? 1 : 0);
    for (int j = j0, it = 0; j < j1; ++j, ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (it / p.stages) & 1;
      mbar_wait(smem_u32(&full_bar[s]), ph);
      fence_after_sync();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t da = BWD ? make_smem_desc_mn(sa + k * 1024, BK * 128, 512) : make_smem_desc(sa + k * UMMA_K * 4);
          const uint64_t db = BWD ? make_smem_desc_mn(sa + B_OFF + k * 1024, BK * 128, 512)
                                  : make_smem_desc(sa + B_OFF + k * UMMA_K * 4);
          umma_tf32(tmem_base, da, db, idesc, (j > j0 || k > 0) ? 1u : 0u);
          if (PRECISE) {
            umma_tf32(tmem_base, da, make_smem_desc(sa + B_OFF + B_BYTES + k * UMMA_K * 4), idesc, 1u);     // a_hi b_lo
            umma_tf32(tmem_base, make_smem_desc(sa + A_BYTES + k * UMMA_K * 4), db, idesc, 1u);             // a_lo b_hi
          }
        }
        umma_commit(smem_u32(&empty_bar[s]));
        if (j == j1 - 1) umma_commit(smem_u32(acc_full));
      }
      __syncwarp();
    }
  } else if (warp < 6) {
    // ================= epilogue =================
    const int quad = warp & 3;
    float* st = stage_c + (size_t)(warp - 2) * 32 * kEpiPad;
    mbar_wait(smem_u32(acc_full), 0);
    fence_after_sync();
    const int row0 = m0 + quad * 32;
    const int rows_ok = min(32, p.M - row0);
    float* outp; i64 ldo; const float* bias;
    if (p.ksplit > 1) { outp = p.ws + (size_t)ks * p.M * p.N; ldo = p.N; bias = nullptr; }
    else { outp = p.out; ldo = p.ldo; bias = p.bias; }
    for (int c0 = 0; c0 < p.N; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) st[lane * kEpiPad + jj] = v[jj];
      __syncwarp();
      const int col = c0 + lane;
      if (col < p.N) {
        const float bv = bias ? bias[col] : 0.f;
        float* cp = outp + (i64)row0 * ldo + col;
#pragma unroll 8
        for (int r = 0; r < rows_ok; ++r) cp[(i64)r * ldo] = st[r * kEpiPad + lane] + bv;
      }
      __syncwarp();
    }
  } else {
    // ================= gather warps: x -> swizzled A tile =================
    // Four groups of four warps; k-chunk number `it` belongs to group it % 4, whose 128 threads move its 1024 16-byte
    // pieces (8 each).  A thread's fence.proxy.async waits for ITS outstanding global loads, so with every warp working on
    // every chunk the loads prefetched for later chunks serialised the pipeline at one memory latency per chunk (measured:
    // 2450 cycles per 16 KB chunk, 20 % of HBM peak); with four independent groups four chunks are in flight per latency.
    // A group may run at most one mbarrier phase ahead of the consumer, i.e. groups <= ring stages (p.groups: 4 or 1).
    const int tg = threadIdx.x - 192;                 // 0..511
    const int tpg = kGatherThreads / p.groups;        // threads per group
    const int grp = tg / tpg, tgi = tg - grp * tpg;   // group, thread in group
    //   forward : piece p -> row r = p / 8 (tile row = output row m0 + r), 16-byte piece q = p % 8 of its 128-byte k-chunk
    //   backward: piece p -> k-row r = p / 32 (chunk row), sub-tile i = (p % 32) / 8 (32 channels each), piece q = p % 8
    constexpr int NPC = 8;                            // max pieces per thread and chunk: p = tgi + tpg h, h < 2 * groups
    const int npc = 2 * p.groups;
    const int q = tgi & 7;
    const int sub = BWD ? (tgi >> 3) & 3 : 0;
    int rr[NPC];
    uint32_t soff[NPC];
#pragma unroll
    for (int h = 0; h < NPC; ++h) {
      if (!BWD) { rr[h] = ((tgi + tpg * h) >> 3) & 127; soff[h] = (uint32_t)(rr[h] * 128 + ((q ^ (rr[h] & 7)) << 4)); }
      else {
        rr[h] = ((tgi + tpg * h) >> 5) & 31;
        soff[h] = (uint32_t)(sub * (BK * 128) + rr[h] * 128 + ((((q >> 1) ^ (rr[h] & 3)) << 5) | ((q & 1) << 4)));
      }
    }
    // forward: per-row constants (the row does not change with the k chunk)
    const float* fbase[NPC]; int fs0[NPC], flen[NPC];
    // backward: channel / window position of this thread's pieces (constant), rows change with the chunk
    int bw = 0, bc = 0; bool bok = false;
    if (!BWD) {
#pragma unroll
      for (int h = 0; h < NPC; ++h) {
        const int m = m0 + rr[h];
        fbase[h] = nullptr; fs0[h] = 0; flen[h] = 0;
        if (m < p.M) {
          const int t2 = m / p.Bsz, b = m - t2 * p.Bsz;
          flen[h] = p.lens[b]; fs0[h] = t2 * p.W;
          fbase[h] = p.x + (i64)b * p.T * p.C + q * 4;
        }
      }
    } else {
      const int mb = m0 + sub * 32;                   // first (w,c) index of this 32-channel sub-tile
      bok = mb < p.M;
      bw = mb / p.C; bc = mb - bw * p.C + q * 4;
    }
    auto load_piece = [&](int j, int h) -> float4 {
      if (!BWD) {
        const int w = j / cpf, c0 = (j - w * cpf) * 32;
        const int sidx = fs0[h] + w;
        if (fbase[h] == nullptr || sidx >= flen[h]) return make_float4(0.f, 0.f, 0.f, 0.f);
        return ldg_f4(fbase[h] + (i64)(flen[h] - 1 - sidx) * p.C + c0);
      } else {
        const int k = j * BK + rr[h];
        if (!bok || k >= p.K) return make_float4(0.f, 0.f, 0.f, 0.f);
        const int t2 = k / p.Bsz, b = k - t2 * p.Bsz;
        const int len = __ldg(p.lens + b);
        const int sidx = t2 * p.W + bw;
        if (sidx >= len) return make_float4(0.f, 0.f, 0.f, 0.f);
        return ldg_f4(p.x + ((i64)b * p.T + (len - 1 - sidx)) * p.C + bc);
      }
    };
    const int n_it = j1 - j0;
    // G groups x D chunks of register prefetch per group (8 float4 per thread either way): G = 4, D = 1 when the ring has
    // >= 4 stages, else every warp works on every chunk (G = 1) with 4 chunks of loads in flight per thread
    auto run = [&](auto G_, auto D_) {
      constexpr int G = decltype(G_)::value, D = decltype(D_)::value, NP_ = 2 * G;
      float4 v[D][NP_];
#pragma unroll
      for (int d = 0; d < D; ++d)
        if (grp + d * G < n_it) {
#pragma unroll
          for (int h = 0; h < NP_; ++h) v[d][h] = load_piece(j0 + grp + d * G, h);
        }
      for (int itb = grp; itb < n_it; itb += G * D) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const int it = itb + d * G;
          if (it < n_it) {
            const int s = it % p.stages;
            const uint32_t ph = (it / p.stages) & 1;
            mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
            const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
#pragma unroll
            for (int h = 0; h < NP_; ++h) {
              if (!PRECISE) sts_f4(sa + soff[h], v[d][h]);
              else {
                const float4 a = v[d][h];
                float4 hi, lo;
                hi.x = __uint_as_float(__float_as_uint(a.x) & 0xffffe000u); lo.x = a.x - hi.x;
                hi.y = __uint_as_float(__float_as_uint(a.y) & 0xffffe000u); lo.y = a.y - hi.y;
                hi.z = __uint_as_float(__float_as_uint(a.z) & 0xffffe000u); lo.z = a.z - hi.z;
                hi.w = __uint_as_float(__float_as_uint(a.w) & 0xffffe000u); lo.w = a.w - hi.w;
                sts_f4(sa + soff[h], hi);
                sts_f4(sa + A_BYTES + soff[h], lo);
              }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(smem_u32(&full_bar[s]));
            if (it + G * D < n_it) {
#pragma unroll
              for (int h = 0; h < NP_; ++h) v[d][h] = load_piece(j0 + it + G * D, h);
            }
          }
        }
      }
    };
    if (p.groups == 4) run(std::integral_constant<int, 4>{}, std::integral_constant<int, 1>{});
    else run(std::integral_constant<int, 1>{}, std::integral_constant<int, 4>{});
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, 256);
  }
}

inline size_t conv_smem_bytes(int BN, int stages, int np) {
  return (size_t)stages * np * (BM * BK * 4 + (size_t)BN * BK * 4) + gemm_fixed_smem();
}

// w -> tf32-exact high part and the remainder (both fp32 arrays): the B operands of the PRECISE forward
__global__ void k_split_tf32(const float* w, float* hi, float* lo, i64 n) {   // hi may alias w
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float a = w[i];
  const float h = __uint_as_float(__float_as_uint(a) & 0xffffe000u);
  hi[i] = h;
  lo[i] = a - h;
}

// can the tensor-core gather kernels run this geometry?  (32-channel k-chunks, 16-byte rows, N fits one tile)
inline bool conv_tc_supported(int C, int W, int E, int T2B) {
  (void)T2B;   // decided per geometry, not per call: the packed weights (hi/lo split) depend on it
  return C % 32 == 0 && E % 4 == 0 && E >= 8 && E <= 256 && W >= 1;
}

// mode 1 (forward): out[T2*B, E] = gather(x) Wc + bias, wT = Wc^T [E, ldw] (K-major).
// mode 2 (backward): out[W*C, E] = gather(x)^T dY, dY [T2*B, E] with leading dimension ldy.
// Bop_lo != NULL (forward only): PRECISE mode, Bop / Bop_lo = hi / lo parts of Wc^T (k_split_tf32).
template <bool BWD>
inline void launch_conv(cudaStream_t st, const float* x, const int* lens, int Bsz, int T, int C, int W, int T2,
                        const float* Bop, const float* Bop_lo, i64 ldb, float* out, i64 ldo, int E, const float* bias) {
  const bool precise = !BWD && Bop_lo != nullptr;
  const int np = precise ? 2 : 1;
  const int nsm = sm_count_();
  ConvP p{};
  p.x = x; p.lens = lens; p.Bsz = Bsz; p.T = T; p.C = C; p.W = W; p.T2 = T2;
  p.N = E;
  if (!BWD) { p.M = T2 * Bsz; p.K = W * C; p.BN = (E + 15) / 16 * 16; }
  else { p.M = W * C; p.K = T2 * Bsz; p.BN = (E + 31) / 32 * 32; }
  p.n_chunks = (p.K + BK - 1) / BK;
  const int tiles_m = (p.M + BM - 1) / BM;
  p.ksplit = std::max(1, std::min(nsm / tiles_m, p.n_chunks / 8));
  p.chunks_per_split = (p.n_chunks + p.ksplit - 1) / p.ksplit;
  p.ksplit = (p.n_chunks + p.chunks_per_split - 1) / p.chunks_per_split;
  int stages = 8;
  while (stages > 2 && conv_smem_bytes(p.BN, stages, np) > kSmemCap) --stages;
  p.stages = stages;
  p.groups = stages >= kGatherGroups ? kGatherGroups : 1;
  p.out = out; p.ldo = ldo; p.bias = bias;
  if (p.ksplit > 1) {
    SplitWs& w = split_ws(st);
    const size_t need = (size_t)p.ksplit * p.M * p.N;
    if (w.n < need) {
      if (w.p) { E2T_CHECK(cudaStreamSynchronize(st)); E2T_CHECK(cudaFree(w.p)); }
      E2T_CHECK(cudaMalloc(&w.p, need * sizeof(float)));
      w.n = need;
    }
    p.ws = w.p;
  }
  CUtensorMap mb = BWD ? make_map(Bop, p.K, E, ldb, BK, true) : make_map(Bop, E, p.K, ldb, p.BN);
  CUtensorMap mlo = precise ? make_map(Bop_lo, E, p.K, ldb, p.BN) : mb;
  static bool attr_set = false;
  if (!attr_set) {
    E2T_CHECK(cudaFuncSetAttribute(k_conv_tc<BWD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCap));
    if (!BWD) E2T_CHECK(cudaFuncSetAttribute(k_conv_tc<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCap));
    attr_set = true;
  }
  const size_t smem_bytes = conv_smem_bytes(p.BN, p.stages, np);
  if (precise) k_conv_tc<false, true><<<tiles_m * p.ksplit, kConvThreads, smem_bytes, st>>>(mb, mlo, p);
  else k_conv_tc<BWD, false><<<tiles_m * p.ksplit, kConvThreads, smem_bytes, st>>>(mb, mlo, p);
  if (p.ksplit > 1) {
    const i64 n = (i64)p.M * p.N;
    k_splitk_reduce<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p.ws, p.ksplit, p.M, p.N, out, ldo, bias, 0.f);
  }
}

}  // namespace conv
