// kernels_simt.cuh -- fp32 CUDA-core kernels of the seq2seq hot path (every op of SURVEY.md §8a).
// These are the validation / small-shape path; the tcgen05 kernels in gemm_tc.cuh replace the GEMMs
// and the recurrence at production shapes.  Written against the CUDA subset that tests/emu can run.
#pragma once
#include "common.cuh"

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide reductions for blockDim.x == 128 or 256 (multiple of 32, <= 1024); `red` = 32 floats smem
__device__ __forceinline__ float block_sum(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < nw; ++i) r += red[i];
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < nw; ++i) r = fmaxf(r, red[i]);
  return r;
}

// ------------------------------------------------------------------------------------------------
// A2: lengths.  One warp per utterance; scans frames from the tail for the last non-zero frame
// (nn.sequences_tools, trainers.py:806-807) unless lens_in is given; lens2 = ceil(len/W).
// ------------------------------------------------------------------------------------------------
__global__ void k_lengths(const float* __restrict__ x, const int* __restrict__ lens_in, int* lens,
                          int* lens2, int* tlast, int B, int T, int C, int W) {
  int b = blockIdx.x;
  int lane = threadIdx.x;
  int len = 0;
  if (lens_in) {
    len = lens_in[b];
    len = len < 0 ? 0 : (len > T ? T : len);
  } else {
    for (int t = T - 1; t >= 0; --t) {
      const float* row = x + ((i64)b * T + t) * C;
      int any = 0;
      for (int c = lane; c < C; c += 32) any |= (row[c] != 0.0f);
      for (int o = 16; o > 0; o >>= 1) any |= __shfl_xor_sync(0xffffffffu, any, o);
      if (any) { len = t + 1; break; }
    }
  }
  if (lane == 0) {
    lens[b] = len;
    int l2 = (len + W - 1) / W;
    lens2[b] = l2;
    tlast[b] = l2 - 1;
  }
}

// ------------------------------------------------------------------------------------------------
// Generic strided fp32 GEMM  C[m,n] = sum_k A(m,k) B(k,n) (+bias[n]) (+beta*C)
//   A(m,k) = A[m*sam + k*sak]   B(k,n) = B[k*sbk + n*sbn]
// CONV=1: A(m,k) gathers the reversed, zero-padded ECoG window (A3+A4 fused): row m = t2*Bsz+b,
//         k = w*C+c  ->  x[b, len_b-1-(t2*W+w), c]   (tf.reverse_sequence + stride-W conv)
// CONV=2: the transpose of that gather (m = w*C+c, k = t2*Bsz+b), for dW_conv = A^T dY.
// 64x64x16 tiles, 256 threads, 4x4 register micro-tile.
// ------------------------------------------------------------------------------------------------
struct GemmP {
  const float* A; i64 sam, sak;
  const float* B; i64 sbk, sbn;
  float* C; i64 ldc;
  int M, N, K;
  const float* bias; float beta;
  const float* x; const int* lens; int Bsz, T, Cch, Wd;
};

template <int CONV>
__device__ __forceinline__ float gemm_load_a(const GemmP& p, int m, int k) {
  if (CONV == 0) return p.A[(i64)m * p.sam + (i64)k * p.sak];
  int row = CONV == 1 ? m : k, col = CONV == 1 ? k : m;
  int t2 = row / p.Bsz, b = row - t2 * p.Bsz;
  int w = col / p.Cch, c = col - w * p.Cch;
  int s = t2 * p.Wd + w, len = p.lens[b];
  if (s >= len) return 0.0f;
  return p.x[((i64)b * p.T + (len - 1 - s)) * p.Cch + c];
}

template <int CONV>
__global__ void __launch_bounds__(256) k_gemm(GemmP p) {
  __shared__ float As[16][65];
  __shared__ float Bs[16][65];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_kfast = (CONV == 1) || (CONV == 0 && p.sak == 1);
  const bool b_nfast = p.sbn == 1;
  for (int k0 = 0; k0 < p.K; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * 256;
      int kk, mm;
      if (a_kfast) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
      int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < p.M && k < p.K) ? gemm_load_a<CONV>(p, m, k) : 0.f;
      int nn;
      if (b_nfast) { nn = e & 63; kk = e >> 6; } else { kk = e & 15; nn = e >> 4; }
      int n = n0 + nn;
      k = k0 + kk;
      Bs[kk][nn] = (n < p.N && k < p.K) ? p.B[(i64)k * p.sbk + (i64)n * p.sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      float* c = p.C + (i64)m * p.ldc + n;
      if (p.beta != 0.f) v += p.beta * (*c);
      *c = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// activation + dropout over X[rows, F] (leading dim ld), in place.  Hash index = row*F + f.
// ------------------------------------------------------------------------------------------------
__global__ void k_act_dropout(float* X, i64 rows, int F, int ld, int act, DropP dp) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * F) return;
  i64 r = i / F;
  int f = (int)(i - r * F);
  float v = X[r * ld + f];
  if (act == 1) v = fmaxf(v, 0.f);
  if (dp.thresh) v = e2t_keep(dp.key, (uint32_t)i, dp.thresh) ? v * dp.inv : 0.f;
  X[r * ld + f] = v;
}
// backward of the above: dX <- dX * d(out)/d(pre); `out` is the stored forward output.
__global__ void k_act_dropout_bwd(float* dX, const float* out, i64 rows, int F, int ld, int act, DropP dp) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * F) return;
  i64 r = i / F;
  int f = (int)(i - r * F);
  float g = dX[r * ld + f];
  if (dp.thresh) g = e2t_keep(dp.key, (uint32_t)i, dp.thresh) ? g * dp.inv : 0.f;
  if (act == 1 && !(out[r * ld + f] > 0.f)) g = 0.f;
  dX[r * ld + f] = g;
}
// dropout-only backward for RNN layer outputs: dX[r, f] *= keep/(1-p)
__global__ void k_dropout_bwd(float* dX, i64 n, DropP dp) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dX[i] = e2t_keep(dp.key, (uint32_t)i, dp.thresh) ? dX[i] * dp.inv : 0.f;
}
// same, four consecutive elements per thread (n % 4 == 0, 16-byte aligned buffer): 16-byte accesses
__global__ void k_dropout_bwd4(float4* dX, i64 n4, DropP dp) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 g = dX[i];
  const uint32_t e = (uint32_t)(4 * i);
  g.x = e2t_keep(dp.key, e, dp.thresh) ? g.x * dp.inv : 0.f;
  g.y = e2t_keep(dp.key, e + 1, dp.thresh) ? g.y * dp.inv : 0.f;
  g.z = e2t_keep(dp.key, e + 2, dp.thresh) ? g.z * dp.inv : 0.f;
  g.w = e2t_keep(dp.key, e + 3, dp.thresh) ? g.w * dp.inv : 0.f;
  dX[i] = g;
}

// ------------------------------------------------------------------------------------------------
// A5/A8: LSTM cell, one time step (TF1 LSTMCell: gates i,j,f,o; forget_bias 1; App. D item 4).
// z [B,4H] holds pre-activations on entry and (sig i, tanh j, sig(f+1), sig o) on exit.
// Rows with t >= lens2[b] (dynamic_rnn past the length): output 0, cell 0, gates untouched.
// ------------------------------------------------------------------------------------------------
struct LstmFwdP {
  float* z; const float* c_prev; float* c_out;
  float* h_out; float* h_drop; int ldh;
  const int* lens2; int t, B, H;
  DropP dp; int drop_F, drop_col0;   // hash index = (t*B+b)*drop_F + drop_col0 + u
};
__global__ void k_lstm_fwd(LstmFwdP p) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.B * p.H) return;
  int b = i / p.H, u = i - b * p.H;
  bool valid = p.lens2 ? (p.t < p.lens2[b]) : true;
  float* hrow = p.h_out + (i64)b * p.ldh + u;
  if (!valid) {
    *hrow = 0.f;
    if (p.h_drop) p.h_drop[(i64)b * p.ldh + u] = 0.f;
    p.c_out[i] = 0.f;
    return;
  }
  float* z = p.z + (i64)b * 4 * p.H + u;
  float gi = sigmoidf_(z[0]);
  float gj = tanhf(z[p.H]);
  float gf = sigmoidf_(z[2 * p.H] + 1.0f);
  float go = sigmoidf_(z[3 * p.H]);
  float cp = p.c_prev ? p.c_prev[i] : 0.f;
  float c = gf * cp + gi * gj;
  float h = go * tanhf(c);
  z[0] = gi; z[p.H] = gj; z[2 * p.H] = gf; z[3 * p.H] = go;
  p.c_out[i] = c;
  *hrow = h;
  if (p.h_drop) {
    uint32_t idx = (uint32_t)(((i64)p.t * p.B + b) * p.drop_F + p.drop_col0 + u);
    p.h_drop[(i64)b * p.ldh + u] = e2t_keep(p.dp.key, idx, p.dp.thresh) ? h * p.dp.inv : 0.f;
  }
}

// Backward of one step.  gz: gate activations in, dz (d loss / d pre-activations) out.
struct LstmBwdP {
  float* gz; const float* c_t; const float* c_prev;
  const float* dh_out; int ldh;        // grad wrt this step's output (nullable)
  const float* dh_rec;                 // [B,H] grad through the recurrence (nullable)
  float* dc_rec;                       // [B,H] in/out
  const float* dc_inject; int ldi;     // final-state cell grad, added when t == inject_t[b] (nullable)
  const int* inject_t; int inject_const;  // inject_t nullable -> use inject_const
  const int* lens2; int t, B, H;
};
__global__ void k_lstm_bwd(LstmBwdP p) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.B * p.H) return;
  int b = i / p.H, u = i - b * p.H;
  float* z = p.gz + (i64)b * 4 * p.H + u;
  bool valid = p.lens2 ? (p.t < p.lens2[b]) : true;
  if (!valid) {
    z[0] = 0.f; z[p.H] = 0.f; z[2 * p.H] = 0.f; z[3 * p.H] = 0.f;
    p.dc_rec[i] = 0.f;
    return;
  }
  float dh = 0.f;
  if (p.dh_out) dh += p.dh_out[(i64)b * p.ldh + u];
  if (p.dh_rec) dh += p.dh_rec[i];
  float dc = p.dc_rec[i];
  if (p.dc_inject) {
    int ti = p.inject_t ? p.inject_t[b] : p.inject_const;
    if (ti == p.t) dc += p.dc_inject[(i64)b * p.ldi + u];
  }
  float gi = z[0], gj = z[p.H], gf = z[2 * p.H], go = z[3 * p.H];
  float c = p.c_t[i];
  float cp = p.c_prev ? p.c_prev[i] : 0.f;
  float tc = tanhf(c);
  float d_o = dh * tc * go * (1.f - go);
  dc += dh * go * (1.f - tc * tc);
  float d_i = dc * gj * gi * (1.f - gi);
  float d_j = dc * gi * (1.f - gj * gj);
  float d_f = dc * cp * gf * (1.f - gf);
  p.dc_rec[i] = dc * gf;
  z[0] = d_i; z[p.H] = d_j; z[2 * p.H] = d_f; z[3 * p.H] = d_o;
}

// ------------------------------------------------------------------------------------------------
// bridge (App. D item 5): decoder (c,h)_0 = concat(fwd final, bwd final) of the last encoder layer.
// hs [T',B,2H]; c_fw, c_bw [T',B,H]; fwd final sits at t = lens2-1, bwd final at t = 0.
// ------------------------------------------------------------------------------------------------
__global__ void k_gather_final(const float* hs, const float* c_fw, const float* c_bw, const int* lens2,
                               float* h0, float* c0, int B, int H) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 2 * H) return;
  int b = i / (2 * H), j = i - b * 2 * H;
  int d = j >= H, u = j - d * H;
  int l2 = lens2[b];
  if (l2 <= 0) { h0[i] = 0.f; c0[i] = 0.f; return; }
  int t = d ? 0 : l2 - 1;
  h0[i] = hs[((i64)t * B + b) * 2 * H + j];
  c0[i] = (d ? c_bw : c_fw)[((i64)t * B + b) * H + u];
}
// transpose of the gather for h: dhs[t_final, b, j] += dh0[b, j]  (dhs zeroed beforehand)
__global__ void k_scatter_final(float* dhs, const float* dh0, const int* lens2, int B, int H) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 2 * H) return;
  int b = i / (2 * H), j = i - b * 2 * H;
  int d = j >= H;
  int l2 = lens2[b];
  if (l2 <= 0) return;
  int t = d ? 0 : l2 - 1;
  dhs[((i64)t * B + b) * 2 * H + j] += dh0[i];
}

// ------------------------------------------------------------------------------------------------
// A7: Luong ("general") attention over the last encoder layer's outputs, fused score / masked softmax / context.
// The reference model has no attention; north_star asks for it, so it is an optional module
// (e2t_config.attention) whose normative spec is oracle/seq2seq_oracle.py: decoder_step (parity unpinned).
//   q = h Wq^T (GEMM, outside) ; score[s] = q . enc[s] for s < lens2[b] ; alpha = softmax(score) ; ctx = sum_s alpha[s] enc[s]
// One block per decoder row r (time-major rows r = k*R + j of q / ctx / alpha; encoder batch index b = j / bdiv, so that the
// beam rows of one utterance share its encoder outputs).  enc [T2, Benc, F] time-major.  Scores live in shared memory
// (T2 floats); each thread keeps its F/blockDim query elements in registers; nothing but alpha and ctx goes to HBM.
// Bahdanau ("additive") variant, kp != NULL: score[s] = sum_a v[a] tanh(q[a] + kp[s,b,a]) with kp = enc Wk^T [T2, Benc, F]
// computed once per batch (GEMM, outside) and q = h Wq^T; softmax and context are the same.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_attn_fwd(const float* __restrict__ q, const float* __restrict__ enc, const int* __restrict__ lens2, float* alpha,
           float* ctx, int R, int Benc, int bdiv, int T2, int F, int ld_alpha, const float* __restrict__ kp,
           const float* __restrict__ v) {
  E2T_DYN_SMEM(float, sc);                 // [T2] scores -> probabilities
  __shared__ float red[32];
  const int r = blockIdx.x;
  const int j = r % R;
  const int b = j / bdiv;
  const int len = min(lens2[b], T2);
  const float* qr = q + (i64)r * F;
  for (int s = 0; s < len; ++s) {
    const float* er = (kp ? kp : enc) + ((i64)s * Benc + b) * F;
    float a = 0.f;
    if (kp) { for (int u = threadIdx.x; u < F; u += blockDim.x) a = fmaf(v[u], tanhf(qr[u] + er[u]), a); }
    else { for (int u = threadIdx.x; u < F; u += blockDim.x) a = fmaf(qr[u], er[u], a); }
    a = block_sum(a, red);
    if (threadIdx.x == 0) sc[s] = a;
  }
  __syncthreads();
  float m = -3.0e38f;
  for (int s = threadIdx.x; s < len; s += blockDim.x) m = fmaxf(m, sc[s]);
  m = block_max(m, red);
  float z = 0.f;
  for (int s = threadIdx.x; s < len; s += blockDim.x) { const float e = expf(sc[s] - m); sc[s] = e; z += e; }
  z = block_sum(z, red);
  const float inv = len > 0 ? 1.0f / z : 0.f;
  __syncthreads();
  for (int s = threadIdx.x; s < T2; s += blockDim.x) {
    const float pv = s < len ? sc[s] * inv : 0.f;
    if (s < len) sc[s] = pv;
    alpha[(i64)r * ld_alpha + s] = pv;
  }
  __syncthreads();
  for (int u = threadIdx.x; u < F; u += blockDim.x) {
    float a = 0.f;
    for (int s = 0; s < len; ++s) a = fmaf(sc[s], enc[((i64)s * Benc + b) * F + u], a);
    ctx[(i64)r * F + u] = a;
  }
}
// backward, phase 1 (one block per decoder row): dalpha[s] = dctx . enc[s] ; dscore = alpha * (dalpha - sum alpha dalpha)
// (written over alpha's companion buffer dscore) ; dq[u] = sum_s dscore[s] enc[s][u]
// Bahdanau (kp != NULL; q holds the forward queries): th = tanh(q[a] + kp[s,b,a]) is recomputed;
// dq[a] = sum_s dscore[s] v[a] (1 - th^2) and dvrow[r, a] = sum_s dscore[s] th (column-summed over r afterwards, fixed order).
__global__ void __launch_bounds__(128)
k_attn_bwd_q(const float* __restrict__ dctx, const float* __restrict__ enc, const int* __restrict__ lens2,
             const float* __restrict__ alpha, float* dscore, float* dq, int R, int Benc, int T2, int F, int ld_alpha,
             const float* __restrict__ kp, const float* __restrict__ v, const float* __restrict__ q, float* dvrow) {
  E2T_DYN_SMEM(float, sc);                 // [T2] dalpha -> dscore
  __shared__ float red[32];
  const int r = blockIdx.x;
  const int b = r % R;
  const int len = min(lens2[b], T2);
  const float* dr = dctx + (i64)r * F;
  const float* ar = alpha + (i64)r * ld_alpha;
  for (int s = 0; s < len; ++s) {
    const float* er = enc + ((i64)s * Benc + b) * F;
    float a = 0.f;
    for (int u = threadIdx.x; u < F; u += blockDim.x) a = fmaf(dr[u], er[u], a);
    a = block_sum(a, red);
    if (threadIdx.x == 0) sc[s] = a;
  }
  __syncthreads();
  float t = 0.f;
  for (int s = threadIdx.x; s < len; s += blockDim.x) t = fmaf(ar[s], sc[s], t);
  t = block_sum(t, red);
  __syncthreads();
  for (int s = threadIdx.x; s < T2; s += blockDim.x) {
    const float v = s < len ? ar[s] * (sc[s] - t) : 0.f;
    if (s < len) sc[s] = v;
    dscore[(i64)r * ld_alpha + s] = v;
  }
  __syncthreads();
  for (int u = threadIdx.x; u < F; u += blockDim.x) {
    float a = 0.f;
    if (kp) {
      const float qu = q[(i64)r * F + u], vu = v[u];
      float dv = 0.f;
      for (int s = 0; s < len; ++s) {
        const float th = tanhf(qu + kp[((i64)s * Benc + b) * F + u]);
        a = fmaf(sc[s] * vu, 1.f - th * th, a);
        dv = fmaf(sc[s], th, dv);
      }
      dvrow[(i64)r * F + u] = dv;
    } else {
      for (int s = 0; s < len; ++s) a = fmaf(sc[s], enc[((i64)s * Benc + b) * F + u], a);
    }
    dq[(i64)r * F + u] = a;
  }
}
// ---- block-per-utterance versions (F <= 32 * kAttnNF) ----------------------------------------------------------------
// One block per utterance b, one warp per decoder row of that utterance (its L teacher-forced steps, or its beams).  The
// encoder rows (or the projected keys) of the utterance are staged ONCE per sweep in shared memory, SC rows at a time, and
// every warp consumes them from there: L2 traffic drops from rows x T2 x F to B x T2 x F per sweep (measured at config 2:
// the row-per-block kernels move 614 MB per launch through L2, 215 us).  A lane keeps its F/32 query / accumulator elements in
// registers, dot products are warp-shuffle reductions, the T2 scores of a row live in that warp's slice of shared memory.
// smem: [SC * F] tile + [warps * T2] scores.
constexpr int kAttnNF = 32;          // register elements per lane -> F <= 1024
__device__ __forceinline__ float attn_tanh(float x) {   // 1 - 2 / (1 + e^{2x}): saturates cleanly, ~1e-7 absolute error
  return 1.0f - 2.0f * __fdividef(1.0f, 1.0f + __expf(2.0f * x));
}
// cooperative copy of rows [c0, c0 + n) of utterance b (time-major [T2, Benc, F]) into the tile
// (four 16-byte loads in flight per thread before the first store: with one block of a dozen warps per SM a load -> store -> load
// chain would expose one L2 latency per element -- measured 100 us per block)
__device__ __forceinline__ void attn_stage(float* tile, const float* __restrict__ src, int c0, int n, int Benc, int b, int F) {
  constexpr int UN = 4;
  const int nt = blockDim.x;
  if ((F & 3) == 0) {          // 16-byte rows
    const int F4 = F >> 2, total = n * F4;
    float4* t4 = reinterpret_cast<float4*>(tile);
    for (int base = threadIdx.x; base < total; base += nt * UN) {
      float4 val[UN];
#pragma unroll
      for (int j = 0; j < UN; ++j) {
        const int i4 = base + j * nt;
        if (i4 < total) {
          const int rr = i4 / F4, cc = i4 - rr * F4;
          val[j] = *reinterpret_cast<const float4*>(src + ((i64)(c0 + rr) * Benc + b) * F + 4 * cc);
        }
      }
#pragma unroll
      for (int j = 0; j < UN; ++j) {
        const int i4 = base + j * nt;
        if (i4 < total) t4[i4] = val[j];
      }
    }
  } else {
    const int total = n * F;
    for (int base = threadIdx.x; base < total; base += nt * UN) {
      float val[UN];
#pragma unroll
      for (int j = 0; j < UN; ++j) {
        const int i = base + j * nt;
        if (i < total) { const int rr = i / F, cc = i - rr * F; val[j] = src[((i64)(c0 + rr) * Benc + b) * F + cc]; }
      }
#pragma unroll
      for (int j = 0; j < UN; ++j) {
        const int i = base + j * nt;
        if (i < total) tile[i] = val[j];
      }
    }
  }
}
// NF = register elements per lane; EXACT: F == 32 * NF (no bounds checks, constant offsets -- the guarded generic version
// executed 6x the instructions: 70.7 M warp instructions per launch at config 2, 220 us); BAH: additive (Bahdanau) score.
#define E2T_ATTN_OK(i, u) (EXACT || ((i) < nf && (u) < F))
template <int NF, bool EXACT, bool BAH>
__global__ void __launch_bounds__(512)
k_attn_fwd_w(const float* __restrict__ q, const float* __restrict__ enc, const int* __restrict__ lens2, float* alpha,
             float* ctx, int R, int Benc, int bdiv, int L, int T2, int F_, int ld_alpha, const float* __restrict__ kp,
             const float* __restrict__ v, int SC) {
  E2T_DYN_SMEM(float, tile);               // [SC][F], then [warps][T2]
  const int F = EXACT ? 32 * NF : F_;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int len = min(lens2[b], T2);
  const int nf = (F + 31) >> 5;
  (void)nf;
  float* sc = tile + (size_t)SC * F + (size_t)warp * T2;
  const float* src = BAH ? kp : enc;
  const int n_rows = L * bdiv;
  for (int g0 = 0; g0 < n_rows; g0 += nw) {
    const int idx = g0 + warp;
    const bool active = idx < n_rows;
    const int k = idx / bdiv, jb = idx - k * bdiv;
    const i64 r = (i64)k * R + (i64)b * bdiv + jb;
    float reg[NF], vv[BAH ? NF : 1];
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      const int u = lane + 32 * i;
      const bool ok = active && E2T_ATTN_OK(i, u);
      reg[i] = ok ? q[r * F + u] : 0.f;
      if (BAH) vv[i] = ok ? v[u] : 0.f;
    }
    // sweep 1: scores
    for (int c0 = 0; c0 < len; c0 += SC) {
      const int n = min(SC, len - c0);
      __syncthreads();
      attn_stage(tile, src, c0, n, Benc, b, F);
      __syncthreads();
      if (active) {
        for (int s = 0; s < n; ++s) {
          const float* er = tile + s * F + lane;
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int i = 0; i < NF; ++i) {
            if (E2T_ATTN_OK(i, lane + 32 * i)) {
              const float t = BAH ? vv[i] * attn_tanh(reg[i] + er[32 * i]) : reg[i] * er[32 * i];
              if (i & 1) a1 += t; else a0 += t;
            }
          }
          const float a = warp_sum(a0 + a1);
          if (lane == 0) sc[c0 + s] = a;
        }
      }
    }
    __syncwarp();
    if (active) {
      float m = -3.0e38f;
      for (int s = lane; s < len; s += 32) m = fmaxf(m, sc[s]);
      m = warp_max(m);
      float z = 0.f;
      for (int s = lane; s < len; s += 32) { const float e = expf(sc[s] - m); sc[s] = e; z += e; }
      z = warp_sum(z);
      const float inv = len > 0 ? 1.0f / z : 0.f;
      for (int s = lane; s < T2; s += 32) {
        const float pv = s < len ? sc[s] * inv : 0.f;
        if (s < len) sc[s] = pv;
        alpha[r * ld_alpha + s] = pv;
      }
    }
    __syncwarp();
    // sweep 2: context (the tile still holds the encoder rows when one chunk covers the utterance and no keys were staged)
#pragma unroll
    for (int i = 0; i < NF; ++i) reg[i] = 0.f;
    const bool reuse = !BAH && len <= SC;
    for (int c0 = 0; c0 < len; c0 += SC) {
      const int n = min(SC, len - c0);
      if (!reuse) {
        __syncthreads();
        attn_stage(tile, enc, c0, n, Benc, b, F);
        __syncthreads();
      }
      if (active) {
        for (int s = 0; s < n; ++s) {
          const float* er = tile + s * F + lane;
          const float pv = sc[c0 + s];
#pragma unroll
          for (int i = 0; i < NF; ++i)
            if (E2T_ATTN_OK(i, lane + 32 * i)) reg[i] = fmaf(pv, er[32 * i], reg[i]);
        }
      }
    }
    if (active) {
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const int u = lane + 32 * i;
        if (E2T_ATTN_OK(i, u)) ctx[r * F + u] = reg[i];
      }
    }
  }
}
// backward phase 1, same mapping (training only: bdiv = 1, rows r = k*B + b)
template <int NF, bool EXACT, bool BAH>
__global__ void __launch_bounds__(512)
k_attn_bwd_q_w(const float* __restrict__ dctx, const float* __restrict__ enc, const int* __restrict__ lens2,
               const float* __restrict__ alpha, float* dscore, float* dq, int R, int Benc, int L, int T2, int F_, int ld_alpha,
               const float* __restrict__ kp, const float* __restrict__ v, const float* __restrict__ q, float* dvrow, int SC) {
  E2T_DYN_SMEM(float, tile);               // [SC][F], then [warps][T2] dalpha -> dscore
  const int F = EXACT ? 32 * NF : F_;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int len = min(lens2[b], T2);
  const int nf = (F + 31) >> 5;
  (void)nf;
  float* sc = tile + (size_t)SC * F + (size_t)warp * T2;
  for (int g0 = 0; g0 < L; g0 += nw) {
    const int k = g0 + warp;
    const bool active = k < L;
    const i64 r = (i64)k * R + b;
    const float* ar = alpha + r * ld_alpha;
    float reg[NF];
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      const int u = lane + 32 * i;
      reg[i] = (active && E2T_ATTN_OK(i, u)) ? dctx[r * F + u] : 0.f;
    }
    // sweep 1: dalpha[s] = dctx . enc[s]
    for (int c0 = 0; c0 < len; c0 += SC) {
      const int n = min(SC, len - c0);
      __syncthreads();
      attn_stage(tile, enc, c0, n, Benc, b, F);
      __syncthreads();
      if (active) {
        for (int s = 0; s < n; ++s) {
          const float* er = tile + s * F + lane;
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int i = 0; i < NF; ++i) {
            if (E2T_ATTN_OK(i, lane + 32 * i)) {
              if (i & 1) a1 = fmaf(reg[i], er[32 * i], a1); else a0 = fmaf(reg[i], er[32 * i], a0);
            }
          }
          const float a = warp_sum(a0 + a1);
          if (lane == 0) sc[c0 + s] = a;
        }
      }
    }
    __syncwarp();
    if (active) {
      float t = 0.f;
      for (int s = lane; s < len; s += 32) t = fmaf(ar[s], sc[s], t);
      t = warp_sum(t);
      for (int s = lane; s < T2; s += 32) {
        const float dv = s < len ? ar[s] * (sc[s] - t) : 0.f;
        if (s < len) sc[s] = dv;
        dscore[r * ld_alpha + s] = dv;
      }
    }
    __syncwarp();
    // sweep 2: dq (Luong: over the encoder rows again; Bahdanau: over the keys, tanh recomputed)
    float dvv[BAH ? NF : 1];
    if (BAH) {
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const int u = lane + 32 * i;
        reg[i] = (active && E2T_ATTN_OK(i, u)) ? q[r * F + u] : 0.f;      // forward query
        dvv[i] = 0.f;
      }
    }
    float acc[NF];
#pragma unroll
    for (int i = 0; i < NF; ++i) acc[i] = 0.f;
    const bool reuse = !BAH && len <= SC;
    for (int c0 = 0; c0 < len; c0 += SC) {
      const int n = min(SC, len - c0);
      if (!reuse) {
        __syncthreads();
        attn_stage(tile, BAH ? kp : enc, c0, n, Benc, b, F);
        __syncthreads();
      }
      if (active) {
        for (int s = 0; s < n; ++s) {
          const float* er = tile + s * F + lane;
          const float ds = sc[c0 + s];
#pragma unroll
          for (int i = 0; i < NF; ++i) {
            if (E2T_ATTN_OK(i, lane + 32 * i)) {
              if (BAH) {
                const float th = attn_tanh(reg[i] + er[32 * i]);
                acc[i] = fmaf(ds, 1.f - th * th, acc[i]);
                dvv[i] = fmaf(ds, th, dvv[i]);
              } else {
                acc[i] = fmaf(ds, er[32 * i], acc[i]);
              }
            }
          }
        }
      }
    }
    if (active) {
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const int u = lane + 32 * i;
        if (E2T_ATTN_OK(i, u)) {
          if (BAH) { dq[r * F + u] = acc[i] * v[u]; dvrow[r * F + u] = dvv[i]; }
          else dq[r * F + u] = acc[i];
        }
      }
    }
  }
}
// backward, phase 2 (one block per utterance b; thread = feature u): denc[s,b,u] += sum_k alpha[k,b,s] dctx[k,b,u]
// + dscore[k,b,s] q[k,b,u], the L decoder steps summed in order (deterministic, no atomics).
// Bahdanau (kp != NULL): the score reaches the encoder through the keys instead:
// dkp[s,b,a] = sum_k dscore[k,b,s] v[a] (1 - tanh^2(q[k,b,a] + kp[s,b,a])) (rows s >= len' get 0); dWk and the encoder
// gradient through Wk follow as two GEMMs outside.
__global__ void __launch_bounds__(256)
k_attn_bwd_enc(const float* __restrict__ dctx, const float* __restrict__ q, const float* __restrict__ alpha,
               const float* __restrict__ dscore, const int* __restrict__ lens2, float* denc, int L, int B, int T2, int F,
               int ld_alpha, const float* __restrict__ kp, const float* __restrict__ v, float* dkp) {
  E2T_DYN_SMEM(float, sm);                 // [2][L][T2]: alpha, dscore of this utterance
  const int b = blockIdx.x;
  const int len = min(lens2[b], T2);
  for (int i = threadIdx.x; i < L * T2; i += blockDim.x) {
    const int k = i / T2, s = i - k * T2;
    sm[i] = alpha[((i64)k * B + b) * ld_alpha + s];
    sm[L * T2 + i] = dscore[((i64)k * B + b) * ld_alpha + s];
  }
  __syncthreads();
  // the L (dctx, q) values of feature u stay in registers across the T2 encoder positions, KC decoder steps at a time
  constexpr int KC = 12;
  for (int u = threadIdx.x; u < F; u += blockDim.x) {
    const float vu = kp ? v[u] : 0.f;
    for (int k0 = 0; k0 < L; k0 += KC) {
      float dc[KC], qq[KC];
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) {
        const bool ok = k0 + kk < L;
        const i64 row = ((i64)(k0 + kk) * B + b) * F + u;
        dc[kk] = ok ? dctx[row] : 0.f;
        qq[kk] = ok ? q[row] : 0.f;
      }
      // eight encoder positions at a time: their denc / kp / dkp values are loaded together (independent loads in flight)
      constexpr int SB = 8;
      for (int s0 = 0; s0 < T2; s0 += SB) {
        float old[SB], kpv[SB], okp[SB];
#pragma unroll
        for (int j = 0; j < SB; ++j) {
          const int s = s0 + j;
          const i64 e = ((i64)s * B + b) * F + u;
          const bool in = s < len;
          old[j] = in ? denc[e] : 0.f;
          kpv[j] = (in && kp) ? kp[e] : 0.f;
          okp[j] = (in && kp && k0 > 0) ? dkp[e] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < SB; ++j) {
          const int s = s0 + j;
          if (s >= T2) continue;
          const i64 e = ((i64)s * B + b) * F + u;
          if (s >= len) { if (kp && k0 == 0) dkp[e] = 0.f; continue; }
          float a = 0.f, dk = 0.f;
#pragma unroll
          for (int kk = 0; kk < KC; ++kk) {
            if (k0 + kk < L) {
              a = fmaf(sm[(k0 + kk) * T2 + s], dc[kk], a);
              const float ds = sm[L * T2 + (k0 + kk) * T2 + s];
              if (kp) {
                const float th = attn_tanh(qq[kk] + kpv[j]);
                dk = fmaf(ds * vu, 1.f - th * th, dk);
              } else {
                a = fmaf(ds, qq[kk], a);
              }
            }
          }
          denc[e] = old[j] + a;
          if (kp) dkp[e] = okp[j] + dk;
        }
      }
    }
  }
}
// X <- tanh(X) in place ; and its backward dX <- dX * (1 - out^2)
__global__ void k_tanh_fwd(float* X, i64 n) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) X[i] = tanhf(X[i]);
}
__global__ void k_tanh_bwd(float* dX, const float* out, i64 n) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const float o = out[i]; dX[i] *= 1.f - o * o; }
}

// ------------------------------------------------------------------------------------------------
// A8: decoder inputs.  y [B,L] -> time-major prev[k,b] (teacher forcing, start token first) and tgt[k,b]
// ------------------------------------------------------------------------------------------------
__global__ void k_shift_targets(const int* y, int* prev, int* tgt, int B, int L, int start_id, int V) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * L) return;
  int k = i / B, b = i - k * B;
  int t = y[b * L + k];
  t = t < 0 ? 0 : (t >= V ? V - 1 : t);
  tgt[i] = t;
  int pv = k == 0 ? start_id : y[b * L + k - 1];
  prev[i] = pv < 0 ? 0 : (pv >= V ? V - 1 : pv);
}
// e[r, :] = dropout(act(Emb[tok[r], :] + bias)); rows r = k*B+b; out leading dim ld
__global__ void k_embed_fwd(const int* tok, const float* emb, const float* bias, float* out, i64 rows,
                            int D, int ld, int act, DropP dp) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * D) return;
  i64 r = i / D;
  int j = (int)(i - r * D);
  float v = emb[(i64)tok[r] * D + j] + bias[j];
  if (act == 1) v = fmaxf(v, 0.f);
  if (dp.thresh) v = e2t_keep(dp.key, (uint32_t)i, dp.thresh) ? v * dp.inv : 0.f;
  out[r * ld + j] = v;
}
// dEmb[tok[r], j] += dpre[r, j]  (dpre already passed through k_act_dropout_bwd)
__global__ void k_embed_bwd(const int* tok, const float* dpre, float* demb, i64 rows, int D, int ld) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * D) return;
  i64 r = i / D;
  int j = (int)(i - r * D);
  atomicAdd(demb + (i64)tok[r] * D + j, dpre[r * ld + j]);
}

// ------------------------------------------------------------------------------------------------
// A9: masked softmax cross-entropy, one block (128 threads) per row r = k*B+b of logits [rows, V].
// loss_row[r] = scale * (lse - logit[tgt]) (0 where tgt == pad); with_grad: logits <- dlogits.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_softmax_ce(float* logits, int ld, int V, const int* tgt,
                                                    int pad_id, float scale, float* loss_row, int with_grad) {
  __shared__ float red[32];
  i64 r = blockIdx.x;
  float* row = logits + r * ld;
  int t = tgt[r];
  if (t == pad_id) {
    if (with_grad)
      for (int j = threadIdx.x; j < V; j += blockDim.x) row[j] = 0.f;
    if (threadIdx.x == 0) loss_row[r] = 0.f;
    return;
  }
  float mx = -3.0e38f;
  for (int j = threadIdx.x; j < V; j += blockDim.x) mx = fmaxf(mx, row[j]);
  mx = block_max(mx, red);
  float s = 0.f;
  for (int j = threadIdx.x; j < V; j += blockDim.x) s += expf(row[j] - mx);
  s = block_sum(s, red);
  float lse = mx + logf(s);
  if (threadIdx.x == 0) loss_row[r] = scale * (lse - row[t]);
  if (with_grad) {
    __syncthreads();
    for (int j = threadIdx.x; j < V; j += blockDim.x) {
      float g = expf(row[j] - lse);
      if (j == t) g -= 1.f;
      row[j] = scale * g;
    }
  }
}
// deterministic single-block reduction of the per-row losses and the unmasked-token count
// count_f (nullable): the token count as fp32 in the slot behind the gradient buffer (rides on the gradient all-reduce);
// acc (nullable): running [loss sum, token count] over the training steps since the last read (no per-step host sync)
__global__ void __launch_bounds__(256) k_reduce_loss(const float* loss_row, const int* tgt, int pad_id, int rows,
                                                     float* loss_out, int* ntok_out, float* count_f, double* acc) {
  __shared__ float red[32];
  float s = 0.f, n = 0.f;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) {
    s += loss_row[i];
    n += (tgt[i] != pad_id) ? 1.f : 0.f;
  }
  s = block_sum(s, red);
  n = block_sum(n, red);
  if (threadIdx.x == 0) {
    *loss_out = s; *ntok_out = (int)(n + 0.5f);
    if (count_f) *count_f = n;
    if (acc) { acc[0] += (double)s; acc[1] += (double)n; }
  }
}

// hi = x with the 13 low mantissa bits cleared (exactly representable in TF32), lo = x - hi: operands of a 3xTF32 product
// (A_hi B_hi + A_lo B_hi + A_hi B_lo), which is fp32-accurate on the tf32 tensor cores.
__global__ void k_split_tf32(const float* __restrict__ in, float* hi, float* lo, i64 n) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = in[i];
  const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  hi[i] = h;
  lo[i] = x - h;
}

// ------------------------------------------------------------------------------------------------
// A6: loss of the encoder-targets head, one warp per row r = t2*B + b of out [rows, F] (leading dim ld).
// The target of row (t2, b) is frame len_b - 1 - t2*W of the caller's [B,T,...] array: reversed within the length,
// then every W-th frame (trainers.py:791-795).  Rows with t2 >= lens2[b] (and, categorical, pad-class rows) are masked.
//   kind 0 (gaussian):    loss = 0.5 * sum_f (out - tgt)^2 ;  d out = out - tgt
//   kind 1 (categorical): loss = lse(out) - out[cls]       ;  d out = softmax - onehot
// loss_row[r] = scale * loss, cnt_row[r] = 1 if unmasked; with_grad: out <- scale * d out (0 on masked rows).
// ------------------------------------------------------------------------------------------------
struct AuxP {
  float* out; int ld, F;
  const void* tgt; int kind;
  const int* lens; const int* lens2;
  int B, T, W, rows;
  float scale;
  float* loss_row; int* cnt_row;
  int with_grad;
};
__global__ void __launch_bounds__(32) k_aux_loss(AuxP p) {
  const int r = blockIdx.x, lane = threadIdx.x;
  const int t2 = r / p.B, b = r - t2 * p.B;
  float* row = p.out + (i64)r * p.ld;
  bool valid = t2 < p.lens2[b];
  const int frame = p.lens[b] - 1 - t2 * p.W;
  int cls = 0;
  if (valid && p.kind == 1) {
    cls = static_cast<const int*>(p.tgt)[(i64)b * p.T + frame];
    valid = cls > 0 && cls < p.F;
  }
  if (!valid) {
    if (p.with_grad)
      for (int f = lane; f < p.F; f += 32) row[f] = 0.f;
    if (lane == 0) { p.loss_row[r] = 0.f; p.cnt_row[r] = 0; }
    return;
  }
  float loss;
  if (p.kind == 0) {
    const float* tg = static_cast<const float*>(p.tgt) + ((i64)b * p.T + frame) * p.F;
    float s = 0.f;
    for (int f = lane; f < p.F; f += 32) {
      const float e = row[f] - tg[f];
      s += e * e;
      if (p.with_grad) row[f] = p.scale * e;
    }
    loss = 0.5f * warp_sum(s);
  } else {
    float mx = -3.0e38f;
    for (int f = lane; f < p.F; f += 32) mx = fmaxf(mx, row[f]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int f = lane; f < p.F; f += 32) s += expf(row[f] - mx);
    s = warp_sum(s);
    const float lse = mx + logf(s);
    loss = lse - row[cls];
    if (p.with_grad) {
      __syncwarp();
      for (int f = lane; f < p.F; f += 32) {
        float g = expf(row[f] - lse);
        if (f == cls) g -= 1.f;
        row[f] = p.scale * g;
      }
    }
  }
  if (lane == 0) { p.loss_row[r] = p.scale * loss; p.cnt_row[r] = 1; }
}
// deterministic single-block reduction of the per-row head losses and the unmasked-frame count
__global__ void __launch_bounds__(256) k_reduce_aux(const float* loss_row, const int* cnt_row, int rows, float* loss_out,
                                                    int* cnt_out, double* acc) {
  __shared__ float red[32];
  float s = 0.f, n = 0.f;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) { s += loss_row[i]; n += (float)cnt_row[i]; }
  s = block_sum(s, red);
  n = block_sum(n, red);
  if (threadIdx.x == 0) { *loss_out = s; *cnt_out = (int)(n + 0.5f); if (acc) { acc[2] += (double)s; acc[3] += (double)n; } }
}

// ------------------------------------------------------------------------------------------------
// A13: input saliency.  tmp [T2*B, W*C] = d(conv pre-activation) Wc^T holds the gradient of every conv window; frame t of
// utterance b is window position s = len_b - 1 - t (the reversal of A3), i.e. row (s / W)*B + b, column (s % W)*C + c.
// tf.reverse_sequence leaves the padding where it is (s = t for t >= len_b), so the zero frames that complete the last
// window still receive a gradient; windows past len' have none (their outputs are masked by the recurrence).
// ------------------------------------------------------------------------------------------------
__global__ void k_saliency_scatter(const float* __restrict__ tmp, i64 ldt, const int* __restrict__ lens, float* dx, int B, int T,
                                   int C, int W) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (i64)B * T * C) return;
  const int c = (int)(i % C);
  const i64 bt = i / C;
  const int t = (int)(bt % T), b = (int)(bt / T);
  const int len = lens[b];
  const int s = t < len ? len - 1 - t : t, t2 = s / W, w = s - t2 * W;
  dx[i] = t2 * W < len ? tmp[((i64)t2 * B + b) * ldt + (i64)w * C + c] : 0.f;
}
// sq[b, c] = sum_t dx[b, t, c]^2 (fixed order)
__global__ void k_saliency_norms(const float* __restrict__ dx, float* sq, int B, int T, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (c >= C) return;
  const float* p = dx + (i64)b * T * C + c;
  float s = 0.f;
  for (int t = 0; t < T; ++t) { const float v = p[(i64)t * C]; s += v * v; }
  sq[(i64)b * C + c] = s;
}

// ------------------------------------------------------------------------------------------------
// column sums (bias gradients): out[n] (+)= sum_m X[m*ld + n].  block = 32 columns x 8 row lanes.
// ------------------------------------------------------------------------------------------------
// blockIdx.y splits the rows into gridDim.y contiguous chunks; chunk y writes out[y*out_stride + n]
// (two deterministic passes for tall matrices: partial sums into a scratch, then a colsum of the scratch).
__global__ void __launch_bounds__(256) k_colsum(const float* X, i64 rows, int N, int ld, float* out, int accumulate,
                                                i64 out_stride) {
  __shared__ float part[8][33];
  int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  int n = blockIdx.x * 32 + cx;
  const i64 chunk = (rows + gridDim.y - 1) / gridDim.y;
  const i64 m0 = (i64)blockIdx.y * chunk;
  i64 m1 = m0 + chunk;
  if (m1 > rows) m1 = rows;
  float s = 0.f;
  if (n < N)
    for (i64 m = m0 + ry; m < m1; m += 8) s += X[m * ld + n];
  part[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && n < N) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += part[i][cx];
    float* o = out + (i64)blockIdx.y * out_stride + n;
    *o = accumulate ? *o + t : t;
  }
}

// ------------------------------------------------------------------------------------------------
// A10: TF1 Adam (lr_t folds the bias corrections) + ExponentialMovingAverage, fused, elementwise.
// ------------------------------------------------------------------------------------------------
// count_dev != NULL: the gradient scale is 1 / max(*count_dev, 1) (the global token count left on the device by the
// data-parallel all-reduce), so that the host never has to read it back.
__global__ void k_adam_ema(float* p, const float* g, float* m, float* v, float* s, i64 n, float grad_scale,
                           float lr_t, float b1, float b2, float eps, float ema_decay, const float* count_dev) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (count_dev) grad_scale = 1.0f / fmaxf(*count_dev, 1.0f);
  float gi = g[i] * grad_scale;
  float mi = b1 * m[i] + (1.f - b1) * gi;
  float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  float pi = p[i] - lr_t * mi / (sqrtf(vi) + eps);
  m[i] = mi; v[i] = vi; p[i] = pi;
  if (ema_decay > 0.f) s[i] = ema_decay * s[i] + (1.f - ema_decay) * pi;
}

// ------------------------------------------------------------------------------------------------
// A11 greedy: one block (128 threads) per row: argmax (lowest index on ties) + log-prob of it under
// softmax(logits / temperature); appends to tokens[b, k]; finished rows emit pad and keep `prev`.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_greedy_pick(const float* logits, int ld, int V, float inv_temp, int k,
                                                     int max_len, int pad_id, int eos_id, int* prev, int* done,
                                                     int* tokens, float* logp) {
  __shared__ float red[32];
  __shared__ int best_idx_s;
  int b = blockIdx.x;
  const float* row = logits + (i64)b * ld;
  float mx = -3.0e38f;
  for (int j = threadIdx.x; j < V; j += blockDim.x) mx = fmaxf(mx, row[j]);
  mx = block_max(mx, red);
  // lowest index attaining the max: encode as -index and take the max
  float bi = -3.0e38f;
  for (int j = threadIdx.x; j < V; j += blockDim.x)
    if (row[j] == mx) bi = fmaxf(bi, -(float)j);
  bi = block_max(bi, red);
  float s = 0.f;
  for (int j = threadIdx.x; j < V; j += blockDim.x) s += expf((row[j] - mx) * inv_temp);
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    int best = (int)(-bi + 0.5f);
    best_idx_s = best;
    int was_done = done[b];
    tokens[(i64)b * max_len + k] = was_done ? pad_id : best;
    if (logp) logp[(i64)b * max_len + k] = was_done ? 0.f : -logf(s);  // (mx-mx)*inv_temp - log(sum)
    if (!was_done) prev[b] = best;
    done[b] = was_done | (best == eos_id);
  }
}

// ------------------------------------------------------------------------------------------------
// A11 / N3, the online-predictor regime (construct_online_predictor, trainers.py:925-949: one utterance at a time).  With
// R <= 8 state rows a decoder step is two matrix-VECTOR products over 18 MB of weights that sit in L2 after the first
// step; the batched path runs it as 5 launches (embedding, a 128-row tensor-core GEMM that is 127/128 padding, cell,
// projection GEMM, pick).  Here it is two:
//   k_dec_small_cell  z = [act(Emb[prev] + b_e), h] [Wx; Wh] + b (rows of the transposed kernel are contiguous: one warp
//                     per gate row, 16-byte loads, the R input vectors staged once per block in shared memory in the
//                     row's own layout), then the cell update of the block's units -- fp32 CUDA cores
//   k_dec_small_pick  logits = h Wp^T + b_p (one warp per vocabulary row), per-block (max, lowest arg-max, sum of exp)
//                     partials, and the LAST block to arrive (ticket counter) merges them and does what k_greedy_pick
//                     does: token, log-probability under softmax(logits / temperature), done flags
// ------------------------------------------------------------------------------------------------
constexpr int kDecSmallRows = 8;      // state rows handled by the small-batch decode kernels
constexpr int kDecSmallUnits = 2;     // hidden units per block of k_dec_small_cell (4 gates x 2 units = 8 warps)
struct DecSmallP {
  const int* prev; const float* emb; const float* emb_b; int act;
  const float* KT; int ldk; const float* bias;          // transposed decoder kernel [4Hd][ldk]: [0, D) embedding part, [Dp, Dp + Hd) state part
  const float* h_in; const float* c_in; float* h_out; float* c_out;
  int R, D, Dp, Hd;
};
__global__ void __launch_bounds__(256) k_dec_small_cell(DecSmallP p) {
  E2T_DYN_SMEM(float, v);                   // [R][Kp] input vectors in the layout of a kernel row, then zs [8][R]
  const int Kp = p.Dp + p.Hd, R = p.R;
  for (int i = threadIdx.x; i < R * Kp; i += blockDim.x) {
    const int r = i / Kp, k = i - r * Kp;
    float x = 0.f;
    if (k < p.D) {
      x = p.emb[(i64)p.prev[r] * p.D + k] + p.emb_b[k];
      if (p.act == 1) x = fmaxf(x, 0.f);
    } else if (k >= p.Dp) x = p.h_in[(i64)r * p.Hd + (k - p.Dp)];
    v[i] = x;
  }
  __syncthreads();
  float* zs = v + R * Kp;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int g = w / kDecSmallUnits, uu = w - g * kDecSmallUnits;
  const int u = blockIdx.x * kDecSmallUnits + uu;
  if (u < p.Hd) {
    const int n = g * p.Hd + u;
    const float4* wrow = reinterpret_cast<const float4*>(p.KT + (i64)n * p.ldk);
    const float4* v4 = reinterpret_cast<const float4*>(v);
    const int K4 = Kp >> 2;
    float acc[kDecSmallRows];
#pragma unroll
    for (int r = 0; r < kDecSmallRows; ++r) acc[r] = 0.f;
    for (int k4 = lane; k4 < K4; k4 += 32) {
      const float4 a = wrow[k4];
#pragma unroll
      for (int r = 0; r < kDecSmallRows; ++r)
        if (r < R) {
          const float4 x = v4[r * K4 + k4];
          acc[r] += a.x * x.x + a.y * x.y + a.z * x.z + a.w * x.w;
        }
    }
    const float bn = p.bias[n];
#pragma unroll
    for (int r = 0; r < kDecSmallRows; ++r)
      if (r < R) {
        const float t = warp_sum(acc[r]);
        if (lane == 0) zs[w * R + r] = t + bn;
      }
  }
  __syncthreads();
  if ((int)threadIdx.x < kDecSmallUnits * R) {
    const int uu2 = threadIdx.x / R, r = threadIdx.x - uu2 * R;
    const int u2 = blockIdx.x * kDecSmallUnits + uu2;
    if (u2 < p.Hd) {      // TF1 LSTMCell: gates i, j, f, o; forget bias 1 (k_lstm_fwd)
      const float gi = sigmoidf_(zs[(0 * kDecSmallUnits + uu2) * R + r]);
      const float gj = tanhf(zs[(1 * kDecSmallUnits + uu2) * R + r]);
      const float gf = sigmoidf_(zs[(2 * kDecSmallUnits + uu2) * R + r] + 1.0f);
      const float go = sigmoidf_(zs[(3 * kDecSmallUnits + uu2) * R + r]);
      const float c = gf * p.c_in[(i64)r * p.Hd + u2] + gi * gj;
      p.c_out[(i64)r * p.Hd + u2] = c;
      p.h_out[(i64)r * p.Hd + u2] = go * tanhf(c);
    }
  }
}
struct DecPickP {
  const float* h; const float* Wp; const float* bp;     // state rows [R][Hd], canonical projection [V][Hd], bias [V]
  int V, Hd, R; float inv_temp; int k, max_len, pad_id, eos_id;
  int* prev; int* done; int* tokens; float* logp;
  float* ws; int* counter;                              // [gridDim.x][R][3] partials; ticket counter (0 between launches)
};
__global__ void __launch_bounds__(256) k_dec_small_pick(DecPickP p) {
  E2T_DYN_SMEM(float, hs);                  // [R][Hd] state rows, then lg [8][R]
  __shared__ int is_last;
  const int R = p.R, Hd = p.Hd;
  for (int i = threadIdx.x; i < R * Hd; i += blockDim.x) hs[i] = p.h[i];
  __syncthreads();
  float* lg = hs + R * Hd;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int vrow = blockIdx.x * 8 + w;
  if (vrow < p.V) {
    const float4* wrow = reinterpret_cast<const float4*>(p.Wp + (i64)vrow * Hd);
    const float4* h4 = reinterpret_cast<const float4*>(hs);
    const int K4 = Hd >> 2;
    float acc[kDecSmallRows];
#pragma unroll
    for (int r = 0; r < kDecSmallRows; ++r) acc[r] = 0.f;
    for (int k4 = lane; k4 < K4; k4 += 32) {
      const float4 a = wrow[k4];
#pragma unroll
      for (int r = 0; r < kDecSmallRows; ++r)
        if (r < R) {
          const float4 x = h4[r * K4 + k4];
          acc[r] += a.x * x.x + a.y * x.y + a.z * x.z + a.w * x.w;
        }
    }
    const float bv = p.bp[vrow];
#pragma unroll
    for (int r = 0; r < kDecSmallRows; ++r)
      if (r < R) {
        const float t = warp_sum(acc[r]);
        if (lane == 0) lg[w * R + r] = t + bv;
      }
  } else if (lane < R) lg[w * R + lane] = -3.0e38f;
  __syncthreads();
  if ((int)threadIdx.x < R) {
    const int r = threadIdx.x;
    float mx = -3.0e38f; int wi = 0;
    for (int i = 0; i < 8; ++i)
      if (lg[i * R + r] > mx) { mx = lg[i * R + r]; wi = i; }       // strict >: the lowest index wins ties
    float sum = 0.f;
    for (int i = 0; i < 8; ++i)
      if (blockIdx.x * 8 + i < (unsigned)p.V) sum += expf((lg[i * R + r] - mx) * p.inv_temp);
    float* o = p.ws + ((i64)blockIdx.x * R + r) * 3;
    o[0] = mx; o[1] = (float)(blockIdx.x * 8 + wi); o[2] = sum;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(p.counter, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (w < R) {        // one warp per state row merges the blocks' partials (blocks cover ascending vocabulary ranges)
    const int r = w;
    const volatile float* ws = p.ws;
    float mx = -3.0e38f;
    for (int b = lane; b < (int)gridDim.x; b += 32) mx = fmaxf(mx, ws[((i64)b * R + r) * 3]);
    mx = warp_max(mx);
    float bi = -3.0e38f, sum = 0.f;
    for (int b = lane; b < (int)gridDim.x; b += 32) {
      const float bm = ws[((i64)b * R + r) * 3];
      if (bm == mx) bi = fmaxf(bi, -ws[((i64)b * R + r) * 3 + 1]);       // lowest index attaining the max
      sum += ws[((i64)b * R + r) * 3 + 2] * expf((bm - mx) * p.inv_temp);
    }
    bi = warp_max(bi);
    sum = warp_sum(sum);
    if (lane == 0) {
      const int best = (int)(-bi + 0.5f);
      const int was_done = p.done[r];
      p.tokens[(i64)r * p.max_len + p.k] = was_done ? p.pad_id : best;
      if (p.logp) p.logp[(i64)r * p.max_len + p.k] = was_done ? 0.f : -logf(sum);
      if (!was_done) p.prev[r] = best;
      p.done[r] = was_done | (best == p.eos_id);
    }
  }
  if (threadIdx.x == 0) *p.counter = 0;
}

// ------------------------------------------------------------------------------------------------
// tiled transpose: out[n*ldo + k] = in[k*ldi + n] for k < K, n < N  (weight re-packing)
// ------------------------------------------------------------------------------------------------
// Gate-column permutation of the layers run by the persistent recurrent kernels (lstm_rec.cuh): canonical
// column n = g*H + u  ->  n' = 64*(u/16) + 32*((u%16)/8) + 8*g + (u%8), so that the 4 gates x 8 units one
// epilogue thread owns are 32 contiguous floats (one 128 B line) of a gates row and 32 contiguous TMEM columns.
#define E2T_REC_UT 8   /* hidden units per epilogue thread of the persistent kernels */
__host__ __device__ __forceinline__ int e2t_gate_perm(int n, int H) {
  int g = n / H, u = n - g * H;
  return 64 * (u >> 4) + (4 * E2T_REC_UT) * ((u & 15) / E2T_REC_UT) + E2T_REC_UT * g + (u % E2T_REC_UT);
}
// permH > 0: output row n is replaced by e2t_gate_perm(n, permH)
__global__ void __launch_bounds__(256) k_transpose(const float* in, i64 ldi, float* out, i64 ldo, int K, int N, int permH) {
  __shared__ float tile[32][33];
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int i = ty; i < 32; i += 8) {
    int k = k0 + i, n = n0 + tx;
    tile[i][tx] = (k < K && n < N) ? in[(i64)k * ldi + n] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    int n = n0 + i, k = k0 + tx;
    if (n < N && k < K) out[(i64)(permH > 0 ? e2t_gate_perm(n, permH) : n) * ldo + k] = tile[tx][i];
  }
}
// out[r, perm(n)] = in[r, n] (forward = 1: canonical -> permuted copy) or out[r, n] = in[r, perm(n)] (forward = 0)
__global__ void k_permute_cols(const float* in, float* out, i64 rows, int N, int permH, int forward) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * N) return;
  i64 r = i / N;
  int n = (int)(i - r * N);
  int np = e2t_gate_perm(n, permH);
  if (forward) out[r * N + np] = in[i];
  else out[i] = in[r * N + np];
}

// ------------------------------------------------------------------------------------------------
// Batched small jobs.  A training step needs ~60 tiny weight re-packs (transposes, gate-column permutations, tf32
// splits) and bias-gradient column sums; launched one by one each costs a launch + a tail of a few microseconds on an
// otherwise idle GPU.  k_batch runs a whole list of them in ONE launch (block -> job by a scan of the job table).
//   E2T_JOB_TRANSPOSE   out[perm(n)][k] = in[k][n]   (permH > 0: row n -> e2t_gate_perm(n); out2 != NULL: tf32 hi -> out, lo -> out2)
//   E2T_JOB_PERMUTE     out[r][perm(n)] = in[r][n] (flag = 1) or out[r][n] = in[r][perm(n)] (flag = 0)
//   E2T_JOB_COLSUM      out[y][n] = sum over the y-th of `flag` row chunks of in[m][n]   (deterministic partial sums;
//                       flag = 1: the final sum; a second job over the partials finishes the two-pass reduction)
// ------------------------------------------------------------------------------------------------
enum { E2T_JOB_TRANSPOSE = 0, E2T_JOB_PERMUTE = 1, E2T_JOB_COLSUM = 2 };
struct BatchJob {
  int type, blk0, nblk;
  int K, N, permH, flag;
  long long ldi, ldo, rows;
  const float* in;
  float* out;
  float* out2;
};
__global__ void __launch_bounds__(256) k_batch(const BatchJob* __restrict__ jobs, int n_jobs) {
  __shared__ float tile[32][33];
  // job of this block = number of jobs whose first block is <= blockIdx.x, minus one: one parallel probe instead of a
  // serial scan of the table (n_jobs <= blockDim.x)
  const int pred = (int)threadIdx.x < n_jobs && jobs[threadIdx.x].blk0 <= (int)blockIdx.x;
  const int ji = __syncthreads_count(pred) - 1;
  const BatchJob J = jobs[ji];
  const int lb = blockIdx.x - J.blk0;
  if (J.type == E2T_JOB_TRANSPOSE) {
    const int tiles_n = (J.N + 31) / 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int k0 = (lb / tiles_n) * 32, n0 = (lb % tiles_n) * 32;
    for (int i = ty; i < 32; i += 8) {
      const int k = k0 + i, n = n0 + tx;
      tile[i][tx] = (k < J.K && n < J.N) ? J.in[(i64)k * J.ldi + n] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
      const int n = n0 + i, k = k0 + tx;
      if (n < J.N && k < J.K) {
        const i64 o = (i64)(J.permH > 0 ? e2t_gate_perm(n, J.permH) : n) * J.ldo + k;
        const float a = tile[tx][i];
        if (J.out2) {
          const float hi = __uint_as_float(__float_as_uint(a) & 0xffffe000u);
          J.out[o] = hi; J.out2[o] = a - hi;
        } else J.out[o] = a;
      }
    }
  } else if (J.type == E2T_JOB_PERMUTE) {
    // block = 1024 consecutive columns of one row (no 64-bit division per element); nblk = rows * ceil(N / 1024)
    const int bpr = (J.N + 1023) / 1024;
    const int r = lb / bpr;
    const i64 base = (i64)r * J.N;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int n = (lb - r * bpr) * 1024 + q * 256 + threadIdx.x;
      if (n < J.N) {
        const int np = e2t_gate_perm(n, J.permH);
        if (J.flag) J.out[base + np] = J.in[base + n];
        else J.out[base + n] = J.in[base + np];
      }
    }
  } else if (J.K == 4) {
    // column sums, 128 columns per block: a thread owns 4 consecutive columns (16-byte loads, 512 B per warp access, four
    // row loads in flight) -- rows 16-byte aligned (ldi % 4 == 0, 16-byte aligned base); K = 4 marks the job as such
    __shared__ float4 part4[8][32];
    const int nx = (J.N + 127) / 128;
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int n = (lb % nx) * 128 + 4 * cx, y = lb / nx;
    const i64 chunk = (J.rows + J.flag - 1) / J.flag;
    const i64 m0 = (i64)y * chunk;
    i64 m1 = m0 + chunk;
    if (m1 > J.rows) m1 = J.rows;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n + 3 < J.N) {
      i64 m = m0 + ry;
      for (; m + 24 < m1; m += 32) {
        const float4 a = *reinterpret_cast<const float4*>(J.in + m * J.ldi + n);
        const float4 b4 = *reinterpret_cast<const float4*>(J.in + (m + 8) * J.ldi + n);
        const float4 c4 = *reinterpret_cast<const float4*>(J.in + (m + 16) * J.ldi + n);
        const float4 d4 = *reinterpret_cast<const float4*>(J.in + (m + 24) * J.ldi + n);
        acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
        acc.x += b4.x; acc.y += b4.y; acc.z += b4.z; acc.w += b4.w;
        acc.x += c4.x; acc.y += c4.y; acc.z += c4.z; acc.w += c4.w;
        acc.x += d4.x; acc.y += d4.y; acc.z += d4.z; acc.w += d4.w;
      }
      for (; m < m1; m += 8) {
        const float4 a = *reinterpret_cast<const float4*>(J.in + m * J.ldi + n);
        acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
      }
    } else if (n < J.N) {     // ragged right edge: scalar
      for (i64 m = m0 + ry; m < m1; m += 8) {
        const float* rp = J.in + m * J.ldi + n;
        acc.x += rp[0];
        if (n + 1 < J.N) acc.y += rp[1];
        if (n + 2 < J.N) acc.z += rp[2];
      }
    }
    part4[ry][cx] = acc;
    __syncthreads();
    if (ry == 0 && n < J.N) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = 0; i < 8; ++i) { const float4 q4 = part4[i][cx]; t.x += q4.x; t.y += q4.y; t.z += q4.z; t.w += q4.w; }
      float* o = J.out + (i64)y * J.ldo + n;
      o[0] = t.x;
      if (n + 1 < J.N) o[1] = t.y;
      if (n + 2 < J.N) o[2] = t.z;
      if (n + 3 < J.N) o[3] = t.w;
    }
  } else {
    const int nx = (J.N + 31) / 32;
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int n = (lb % nx) * 32 + cx, y = lb / nx;
    const i64 chunk = (J.rows + J.flag - 1) / J.flag;
    const i64 m0 = (i64)y * chunk;
    i64 m1 = m0 + chunk;
    if (m1 > J.rows) m1 = J.rows;
    float sacc = 0.f;
    if (n < J.N)
      for (i64 m = m0 + ry; m < m1; m += 8) sacc += J.in[m * J.ldi + n];
    tile[ry][cx] = sacc;
    __syncthreads();
    if (ry == 0 && n < J.N) {
      float t = 0.f;
      for (int i = 0; i < 8; ++i) t += tile[i][cx];
      J.out[(i64)y * J.ldo + n] = t;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// A11 beam search helpers (width <= 32).  State rows are r = b*beam + j.
// k_beam_topk: one block per utterance.  For each live beam the candidates are
//   score[j] + log_softmax(logits[r]/T)[v]; finished beams contribute only (score[j], pad).
// Selects the `beam` best (score desc, then flat index j*V+v asc) by repeated block arg-max.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_beam_topk(const float* logits, int ld, int V, float inv_temp, int beam,
                                                   const float* score_in, const int* done_in, int pad_id,
                                                   float* lse_ws, float* score_out, int* src_out, int* tok_out) {
  __shared__ float red[32];
  __shared__ float sel_score[32];
  __shared__ int sel_flat[32];
  int b = blockIdx.x;
  // log-sum-exp per beam row at temperature
  for (int j = 0; j < beam; ++j) {
    const float* row = logits + ((i64)b * beam + j) * ld;
    float mx = -3.0e38f;
    for (int v = threadIdx.x; v < V; v += blockDim.x) mx = fmaxf(mx, row[v] * inv_temp);
    mx = block_max(mx, red);
    float s = 0.f;
    for (int v = threadIdx.x; v < V; v += blockDim.x) s += expf(row[v] * inv_temp - mx);
    s = block_sum(s, red);
    if (threadIdx.x == 0) lse_ws[b * beam + j] = mx + logf(s);
  }
  __syncthreads();
  for (int n = 0; n < beam; ++n) {
    float best = -3.0e38f;
    int best_flat = 0x7fffffff;
    for (int j = 0; j < beam; ++j) {
      int r = b * beam + j;
      float sc = score_in[r];
      if (done_in[r]) {
        if (threadIdx.x == 0) {
          int flat = j * V + pad_id;
          bool taken = false;
          for (int q = 0; q < n; ++q) taken |= (sel_flat[q] == flat);
          if (!taken && (sc > best || (sc == best && flat < best_flat))) { best = sc; best_flat = flat; }
        }
        continue;
      }
      const float* row = logits + (i64)r * ld;
      float lse = lse_ws[r];
      for (int v = threadIdx.x; v < V; v += blockDim.x) {
        float cand = sc + (row[v] * inv_temp - lse);
        int flat = j * V + v;
        if (cand > best || (cand == best && flat < best_flat)) {
          bool taken = false;
          for (int q = 0; q < n; ++q) taken |= (sel_flat[q] == flat);
          if (!taken) { best = cand; best_flat = flat; }
        }
      }
    }
    // block arg-max with lowest-flat tie-break: first the max score, then the min flat among holders
    float bm = block_max(best, red);
    float cand_flat = (best == bm) ? -(float)best_flat : -3.0e38f;
    float bf = block_max(cand_flat, red);
    if (threadIdx.x == 0) {
      sel_score[n] = bm;
      sel_flat[n] = (int)(-bf + 0.5f);
    }
    __syncthreads();
  }
  if (threadIdx.x < beam) {
    int n = threadIdx.x;
    score_out[b * beam + n] = sel_score[n];
    src_out[b * beam + n] = sel_flat[n] / V;
    tok_out[b * beam + n] = sel_flat[n] % V;
  }
}
// Same selection, one pass over the logits (V <= 32 * NV, beam <= 32): one WARP per beam row keeps the row's candidates
// score[j] + log_softmax(logits/T)[v] in registers (NV per lane), finds the row's `beam` best by warp arg-max rounds
// (score desc, flat index asc -- the same total order as above), and warp 0 merges the beam x beam survivors.  The
// block-wide version above re-scans all beam x V candidates with two block reductions per selected beam: 114 us per decode
// step at beam 8, V = 1806, 43 % of a beam-8 decode (profiles/r1u_decode_breakdown.txt).
template <int NV>
__global__ void __launch_bounds__(256) k_beam_topk_w(const float* __restrict__ logits, int ld, int V, float inv_temp, int beam,
                                                     const float* __restrict__ score_in, const int* __restrict__ done_in,
                                                     int pad_id, float* score_out, int* src_out, int* tok_out) {
  __shared__ float cand_score[32 * 32];
  __shared__ int cand_flat[32 * 32];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float NEG = -3.0e38f;
  for (int j = warp; j < beam; j += nw) {
    const int r = b * beam + j;
    const float sc = score_in[r];
    if (done_in[r]) {     // a finished beam contributes only (score, pad)
      if (lane < beam) { cand_score[j * beam + lane] = lane == 0 ? sc : NEG; cand_flat[j * beam + lane] = j * V + pad_id + lane; }
      continue;
    }
    const float* row = logits + (i64)r * ld;
    float c[NV];
    float mx = NEG;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      c[i] = v < V ? row[v] * inv_temp : NEG;
      mx = fmaxf(mx, c[i]);
    }
    mx = warp_max(mx);
    float se = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + 32 * i < V) se += expf(c[i] - mx);
    se = warp_sum(se);
    const float off = sc - (mx + logf(se));
#pragma unroll
    for (int i = 0; i < NV; ++i) c[i] = (lane + 32 * i < V) ? c[i] + off : NEG;
    for (int n = 0; n < beam; ++n) {
      float best = NEG;
      int bi = 0;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (c[i] > best) { best = c[i]; bi = i; }      // ascending i, strict >: lowest index on ties
      int flat = j * V + lane + 32 * bi;
      if (best == NEG) flat = 0x7fffffff;              // this lane has nothing left
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int of = __shfl_xor_sync(0xffffffffu, flat, o);
        if (ob > best || (ob == best && of < flat)) { best = ob; flat = of; }
      }
      if (lane == 0) { cand_score[j * beam + n] = best; cand_flat[j * beam + n] = flat; }
      const int loc = flat - j * V;                    // the owner removes the winner from its registers
      if (flat != 0x7fffffff && (loc & 31) == lane) {
        const int wi = loc >> 5;
#pragma unroll
        for (int i = 0; i < NV; ++i)
          if (i == wi) c[i] = NEG;
      }
    }
  }
  __syncthreads();
  if (warp == 0) {
    const int nc = beam * beam;
    for (int n = 0; n < beam; ++n) {
      float best = NEG;
      int flat = 0x7fffffff, pos = -1;
      for (int p = lane; p < nc; p += 32) {
        const float s = cand_score[p];
        const int f = cand_flat[p];
        if (s > best || (s == best && f < flat)) { best = s; flat = f; pos = p; }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int of = __shfl_xor_sync(0xffffffffu, flat, o);
        const int op = __shfl_xor_sync(0xffffffffu, pos, o);
        if (ob > best || (ob == best && of < flat)) { best = ob; flat = of; pos = op; }
      }
      if (lane == 0) {
        score_out[b * beam + n] = best;
        src_out[b * beam + n] = flat / V;
        tok_out[b * beam + n] = flat % V;
        if (pos >= 0) { cand_score[pos] = NEG; cand_flat[pos] = 0x7fffffff; }
      }
      __syncwarp();
    }
  }
}
// reorder beam state after top-k: rows of (h,c) gathered from src; finished sources keep their old state
struct BeamStepP {
  const float* h_new; const float* c_new; const float* h_old; const float* c_old;
  float* h_out; float* c_out; int Hd;
  const int* src; const int* tok; const int* done_in; const int* prev_in;
  int* done_out; int* prev_out; const int* toks_in; int* toks_out;
  int beam, k, max_len, pad_id, eos_id, rows;
};
__global__ void k_beam_reorder(BeamStepP p) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (i64)p.rows * p.Hd) return;
  int r = (int)(i / p.Hd), u = (int)(i - (i64)r * p.Hd);
  int b = r / p.beam;
  int sr = b * p.beam + p.src[r];
  bool was_done = p.done_in[sr] != 0;
  p.h_out[i] = was_done ? p.h_old[(i64)sr * p.Hd + u] : p.h_new[(i64)sr * p.Hd + u];
  p.c_out[i] = was_done ? p.c_old[(i64)sr * p.Hd + u] : p.c_new[(i64)sr * p.Hd + u];
  if (u == 0) {
    int tk = p.tok[r];
    for (int q = 0; q < p.max_len; ++q) {
      int v = q < p.k ? p.toks_in[(i64)sr * p.max_len + q] : p.pad_id;
      if (q == p.k) v = was_done ? p.pad_id : tk;
      p.toks_out[(i64)r * p.max_len + q] = v;
    }
    p.prev_out[r] = was_done ? p.prev_in[sr] : tk;
    p.done_out[r] = was_done | (tk == p.eos_id);
  }
}
__global__ void k_fill_int(int* p, int v, i64 n) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void k_fill_float(float* p, float v, i64 n) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
// beam init: scores[b,0]=0, others -1e30; h/c rows replicated from [B,Hd] to [B*beam,Hd]
__global__ void k_beam_init(const float* h0, const float* c0, float* h, float* c, float* score, int B, int beam, int Hd) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (i64)B * beam * Hd) return;
  int r = (int)(i / Hd), u = (int)(i - (i64)r * Hd);
  int b = r / beam, j = r - b * beam;
  h[i] = h0[(i64)b * Hd + u];
  c[i] = c0[(i64)b * Hd + u];
  if (u == 0) score[r] = j == 0 ? 0.f : -1.0e30f;
}
