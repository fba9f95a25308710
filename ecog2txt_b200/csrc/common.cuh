// common.cuh -- launch macro, error plumbing and the dropout hash shared by all kernels.
#pragma once
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#ifndef E2T_EMU
#include <cuda_runtime.h>
// Real CUDA build.  (The g++ emulation build used by the CPU-only tests force-includes
// tests/emu/cuda_emu.h, which defines E2T_LAUNCH / E2T_DYN_SMEM for fibers instead.)
#define E2T_LAUNCH(kern, grid, block, smem, stream, ...) \
  kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define E2T_DYN_SMEM(type, name) extern __shared__ __align__(1024) unsigned char name##_raw_[]; \
  type* name = reinterpret_cast<type*>(name##_raw_)
#endif

typedef long long i64;

#define E2T_CHECK(expr)                                                                          \
  do {                                                                                           \
    cudaError_t e_ = (expr);                                                                     \
    if (e_ != cudaSuccess)                                                                       \
      throw std::runtime_error(std::string(#expr) + " failed: " + cudaGetErrorString(e_) + " @" + \
                               __FILE__ + ":" + std::to_string(__LINE__));                       \
  } while (0)

#define E2T_REQUIRE(cond, msg)                                                        \
  do {                                                                                \
    if (!(cond)) throw std::runtime_error(std::string("e2t: ") + (msg) + " [" #cond "]"); \
  } while (0)

// ---- counter-based dropout hash: identical integer recipe in oracle/seq2seq_oracle.py ----------
__host__ __device__ __forceinline__ uint32_t e2t_mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t e2t_stream_key(uint32_t seed, uint32_t stream) {
  return e2t_mix32(seed * 0x9E3779B1u + stream * 0x85EBCA77u + 0x165667B1u);
}
__host__ __device__ __forceinline__ bool e2t_keep(uint32_t key, uint32_t idx, uint32_t thresh) {
  uint32_t h = e2t_mix32(idx ^ key);
  h = e2t_mix32(h + 0x27D4EB2Fu);
  return h >= thresh;
}
static inline uint32_t e2t_thresh(float p) {
  double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}
// dropout stream ids (oracle: STREAM_CONV, STREAM_ENC0 + l, STREAM_DEMB)
enum { E2T_STREAM_CONV = 0, E2T_STREAM_ENC0 = 1, E2T_STREAM_DEMB = 64, E2T_STREAM_AUX = 96, E2T_STREAM_PROJ = 112 };

struct DropP {  // p == 0 <=> thresh == 0 && inv == 1
  uint32_t key, thresh;
  float inv;
};
