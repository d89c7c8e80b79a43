// Shared helpers for the r3d_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/r3d_b200.h"

namespace r3d {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// cudaFuncSetAttribute and streams/events belong to ONE device; a process that drives several GPUs from several
// threads (nn.DataParallel) needs them once per device.  Returns true the first time it is called for the current
// device with this flag array (benign race: the guarded calls are idempotent).
constexpr int kMaxDevices = 64;
inline int current_device_index() {
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < kMaxDevices) ? d : 0;
}
inline bool per_device_once(bool (&done)[kMaxDevices]) {
  const int d = current_device_index();
  if (done[d]) return false;
  done[d] = true;
  return true;
}

// The host-buffer entry points take their device arena from the stream-ordered allocator on every call.  By default
// the pool hands unused memory back to the driver at every synchronisation, so each call would map several GB again
// (tens of ms up to a second).  Keep it: raise the release threshold of the device's default pool once.
inline void keep_async_pool() {
  static bool done[kMaxDevices] = {};
  if (!per_device_once(done)) return;
  cudaMemPool_t pool = nullptr;
  if (cudaDeviceGetDefaultMemPool(&pool, current_device_index()) == cudaSuccess && pool) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
}

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// ---- optional per-stage timing (CUDA events on the launching stream) -------------
enum Stage {
  ST_SCORE_PARTIAL = 0, ST_SCORE_FINALIZE, ST_BOTTOMK, ST_EXCHANGE_FWD, ST_EXCHANGE_BWD, ST_COLSUM_FINALIZE,
  ST_BN_STATS, ST_BN_BWD, ST_GRAM, ST_JACOBI_INIT, ST_JACOBI_INNER, ST_JACOBI_UPDATE, ST_JACOBI_EXTRACT,
  ST_REFINE_Y, ST_SIGMA, ST_ENTROPY, ST_COEF, ST_BWD_GEMM, ST_TOKEN_INFO, ST_BLOCK, ST_JACOBI_VUPDATE,
  ST_JACOBI_LOCAL, ST_NUM
};
bool profiling_enabled();
void stage_begin(int stage, cudaStream_t st);
void stage_end(int stage, cudaStream_t st);
struct StageScope {
  int stage; cudaStream_t st; bool on;
  StageScope(int s, cudaStream_t t) : stage(s), st(t), on(profiling_enabled()) { if (on) stage_begin(stage, st); }
  ~StageScope() { if (on) stage_end(stage, st); }
};
#define R3D_STAGE(id, st) r3d::StageScope _stage_scope_##id(r3d::id, st)

#define R3D_CHECK(cond, ...)            \
  do {                                  \
    if (!(cond)) {                      \
      r3d::set_error(__VA_ARGS__);      \
      return 1;                         \
    }                                   \
  } while (0)

#define R3D_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      r3d::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 2;                                                                          \
    }                                                                                    \
  } while (0)

#define R3D_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      r3d::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 3;                                                                          \
    }                                                                                    \
    r3d::count_launch();                                                                 \
  } while (0)

// ---- element <-> float vector access (128-bit when V == VecOf<T>::N) -----------
template <typename T>
struct VecOf;
template <>
struct VecOf<float> {
  static constexpr int N = 4;
};
template <>
struct VecOf<__nv_bfloat16> {
  static constexpr int N = 8;
};

// load_vec / store_vec: streaming (evict-first) accesses for tensors whose LAST use on the path this is (exchange
// inputs, gradients).  load_vec_keep: default cache policy for a tensor the next kernel reads again -- the score and
// BatchNorm-statistics passes read X (67 MB at the headline shape, inside the 126 MB L2) right before the exchange does.
template <typename T, int V>
__device__ __forceinline__ void load_vec(const T* __restrict__ p, float (&f)[V]);
template <typename T, int V>
__device__ __forceinline__ void load_vec_keep(const T* __restrict__ p, float (&f)[V]);
template <>
__device__ __forceinline__ void load_vec_keep<float, 4>(const float* __restrict__ p, float (&f)[4]) {
  float4 v = *reinterpret_cast<const float4*>(p);
  f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
template <>
__device__ __forceinline__ void load_vec_keep<float, 1>(const float* __restrict__ p, float (&f)[1]) { f[0] = *p; }
template <>
__device__ __forceinline__ void load_vec_keep<__nv_bfloat16, 8>(const __nv_bfloat16* __restrict__ p, float (&f)[8]) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <>
__device__ __forceinline__ void load_vec_keep<__nv_bfloat16, 1>(const __nv_bfloat16* __restrict__ p, float (&f)[1]) {
  f[0] = __bfloat162float(*p);
}

template <>
__device__ __forceinline__ void load_vec<float, 4>(const float* __restrict__ p, float (&f)[4]) {
  float4 v = __ldcs(reinterpret_cast<const float4*>(p));
  f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
template <>
__device__ __forceinline__ void load_vec<float, 1>(const float* __restrict__ p, float (&f)[1]) {
  f[0] = __ldcs(p);
}
template <>
__device__ __forceinline__ void load_vec<__nv_bfloat16, 8>(const __nv_bfloat16* __restrict__ p, float (&f)[8]) {
  uint4 v = __ldcs(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <>
__device__ __forceinline__ void load_vec<__nv_bfloat16, 1>(const __nv_bfloat16* __restrict__ p, float (&f)[1]) {
  f[0] = __bfloat162float(*p);
}

template <typename T, int V>
__device__ __forceinline__ void store_vec(T* __restrict__ p, const float (&f)[V]);

template <>
__device__ __forceinline__ void store_vec<float, 4>(float* __restrict__ p, const float (&f)[4]) {
  __stcs(reinterpret_cast<float4*>(p), make_float4(f[0], f[1], f[2], f[3]));
}
template <>
__device__ __forceinline__ void store_vec<float, 1>(float* __restrict__ p, const float (&f)[1]) {
  __stcs(p, f[0]);
}
template <>
__device__ __forceinline__ void store_vec<__nv_bfloat16, 8>(__nv_bfloat16* __restrict__ p, const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  __stcs(reinterpret_cast<uint4*>(p), make_uint4(w[0], w[1], w[2], w[3]));
}
template <>
__device__ __forceinline__ void store_vec<__nv_bfloat16, 1>(__nv_bfloat16* __restrict__ p, const float (&f)[1]) {
  *p = __float2bfloat16_rn(f[0]);
}

// ---- the (rows, C) streaming tile every elementwise/column-reduction kernel uses ----
// A CTA of 256 threads is TX column-vector lanes x TY row lanes; gridDim.y covers
// column chunks of TX*V channels, gridDim.x covers row chunks.  Each thread owns ONE
// column vector for all of its rows, so per-channel state (masks, alpha, accumulators)
// lives in registers.
struct Tile {
  int tx_log2;           // TX = 1 << tx_log2
  int col_chunks;        // gridDim.y
  int row_chunks;        // gridDim.x
  int64_t rows_per_cta;  // multiple of TY
};

inline int ilog2_ceil(int64_t v) {
  int l = 0;
  while ((int64_t(1) << l) < v) ++l;
  return l;
}

// ctas_per_sm: how many CTAs of this kernel we want resident per SM.
inline Tile make_tile(int64_t rows, int64_t C, int V, int ctas_per_sm) {
  Tile t;
  int64_t cv = C / V;
  t.tx_log2 = ilog2_ceil(cv);
  if (t.tx_log2 > 8) t.tx_log2 = 8;
  int TX = 1 << t.tx_log2, TY = 256 / TX;
  t.col_chunks = int((cv + TX - 1) / TX);
  int64_t want = int64_t(kNumSMs) * ctas_per_sm / t.col_chunks;
  if (want < 1) want = 1;
  int64_t rpc = (rows + want - 1) / want;
  rpc = ((rpc + TY - 1) / TY) * TY;
  if (rpc < TY) rpc = TY;
  t.rows_per_cta = rpc;
  t.row_chunks = int((rows + rpc - 1) / rpc);
  if (t.row_chunks < 1) t.row_chunks = 1;
  return t;
}

template <typename T>
inline bool vec_ok(const void* p, int64_t C) {
  return (reinterpret_cast<uintptr_t>(p) % 16 == 0) && (C % VecOf<T>::N == 0);
}

}  // namespace r3d
