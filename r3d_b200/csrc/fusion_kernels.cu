// Token-fusion kernels for sm_100a: channel score, bottom-k, exchange/blend
// forward and backward, BatchNorm front end.  All HBM-bound streaming kernels:
// 128-bit accesses, one column vector per thread for all of its rows, grids sized
// in multiples of the SM count, deterministic two-stage column reductions.
//
// Reference behaviour replaced (paths relative to the reference repo):
//   score     model/futr_safuser_tokenfusion.py:49-50
//   bottom-k  model/futr_safuser_tokenfusion.py:52-54
//   exchange  model/futr_safuser_tokenfusion.py:56-62, ..._vary.py:48-57,
//             ..._batchnormalization.py:45-46,62-75
#include <stdarg.h>

#include <vector>

#include "common.cuh"

namespace r3d {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }

// ---- stage timing -----------------------------------------------------------------
struct StageRec { int stage; cudaEvent_t a, b; };
static thread_local bool g_prof = false;
static thread_local std::vector<StageRec>* g_recs = nullptr;
static thread_local std::vector<cudaEvent_t>* g_pool = nullptr;
static thread_local cudaEvent_t g_open[ST_NUM];
static thread_local int64_t g_stage_launch0[ST_NUM];
static thread_local double g_ms[ST_NUM];
static thread_local int64_t g_cnt[ST_NUM];
static thread_local int64_t g_kl[ST_NUM];

bool profiling_enabled() { return g_prof; }
static cudaEvent_t get_event() {
  if (!g_pool) g_pool = new std::vector<cudaEvent_t>();
  if (!g_pool->empty()) { cudaEvent_t e = g_pool->back(); g_pool->pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
void stage_begin(int stage, cudaStream_t st) {
  cudaEvent_t e = get_event();
  cudaEventRecord(e, st);
  g_open[stage] = e;
  g_stage_launch0[stage] = g_launches;
}
void stage_end(int stage, cudaStream_t st) {
  cudaEvent_t e = get_event();
  cudaEventRecord(e, st);
  if (!g_recs) g_recs = new std::vector<StageRec>();
  g_recs->push_back({stage, g_open[stage], e});
  g_cnt[stage] += 1;
  g_kl[stage] += g_launches - g_stage_launch0[stage];
}
static void drain_records() {
  if (!g_recs) return;
  for (auto& r : *g_recs) {
    cudaEventSynchronize(r.b);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) g_ms[r.stage] += ms;
    g_pool->push_back(r.a);
    g_pool->push_back(r.b);
  }
  g_recs->clear();
}

// half-width bf16 vectors (8-byte accesses) for the register-heavy BN backward
template <>
__device__ __forceinline__ void load_vec<__nv_bfloat16, 4>(const __nv_bfloat16* __restrict__ p, float (&f)[4]) {
  uint2 v = __ldcs(reinterpret_cast<const uint2*>(p));
  f[0] = __uint_as_float(v.x << 16);
  f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16);
  f[3] = __uint_as_float(v.y & 0xffff0000u);
}
template <>
__device__ __forceinline__ void store_vec<__nv_bfloat16, 4>(__nv_bfloat16* __restrict__ p, const float (&f)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(f[2], f[3]);
  __stcs(reinterpret_cast<uint2*>(p), make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b)));
}

// Row chunking is a function of `rows` alone so that the finalize kernels can
// recompute it without knowing dtype or vector width.
int g_row_chunk_mult = 4;   // row chunks are capped at this many per SM (tuning knob "row_chunk_mult")
static inline int fixed_row_chunks(int64_t rows) {
  int64_t rc = rows / 16;
  if (rc < 1) rc = 1;
  if (rc > kNumSMs * g_row_chunk_mult) rc = kNumSMs * g_row_chunk_mult;
  return int(rc);
}

struct Grid2 {
  int tx_log2, col_chunks, row_chunks;
  int64_t rows_per_cta;
};
static inline Grid2 make_grid(int64_t rows, int64_t C, int V) {
  Grid2 g;
  int64_t cv = C / V;
  g.tx_log2 = ilog2_ceil(cv);
  if (g.tx_log2 > 8) g.tx_log2 = 8;
  int TX = 1 << g.tx_log2, TY = 256 / TX;
  g.col_chunks = int((cv + TX - 1) / TX);
  g.row_chunks = fixed_row_chunks(rows);
  int64_t rpc = (rows + g.row_chunks - 1) / g.row_chunks;
  g.rows_per_cta = ((rpc + TY - 1) / TY) * TY;
  return g;
}

// ------------------------------------------------------------------------------
// a1: per-CTA column sums of |x|  (stage 1)
// ------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(256) score_partial_kernel(const T* __restrict__ rgb, const T* __restrict__ depth,
                                                            int64_t rows, int64_t C, int tx_log2,
                                                            int64_t rows_per_cta, float* __restrict__ partial) {
  __shared__ float red[256 * V];
  const int TX = 1 << tx_log2, TY = 256 >> tx_log2;
  const int tx = threadIdx.x & (TX - 1), ty = threadIdx.x >> tx_log2;
  const int64_t cv = int64_t(blockIdx.y) * TX + tx;
  const bool active = cv * V < C;
  const T* __restrict__ x = blockIdx.z ? depth : rgb;
  const int64_t r0 = int64_t(blockIdx.x) * rows_per_cta;
  const int64_t r1 = min(rows, r0 + rows_per_cta);
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  if (active) {
    const T* p = x + cv * V;
    // batches of 8 independent 128-bit loads per thread; rows past the end contribute +0 (fixed summation order)
    for (int64_t r = r0 + ty; r < r1; r += 8 * TY) {
      float a[8][V];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (r + u * TY < r1) load_vec_keep<T, V>(p + (r + u * TY) * C, a[u]);
        else {
#pragma unroll
          for (int i = 0; i < V; ++i) a[u][i] = 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < V; ++i)
        acc[i] += ((fabsf(a[0][i]) + fabsf(a[1][i])) + (fabsf(a[2][i]) + fabsf(a[3][i]))) +
                  ((fabsf(a[4][i]) + fabsf(a[5][i])) + (fabsf(a[6][i]) + fabsf(a[7][i])));
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) red[(ty * TX + tx) * V + i] = acc[i];
  __syncthreads();
  if (ty == 0 && active) {
    float* out = partial + (int64_t(blockIdx.z) * gridDim.x + blockIdx.x) * C + cv * V;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float s = 0.f;
      for (int y = 0; y < TY; ++y) s += red[(y * TX + tx) * V + i];   // fixed order
      out[i] = s;
    }
  }
}

// Stage 2: 32 channels x 8 partial-sum lanes per CTA; every lane adds its slice of the row chunks in index
// order, then lane 0 adds the 8 lane sums in order -- a fixed summation tree, so results are bit-reproducible.
// Optional tail (the packed statistic of the multi-GPU path): tail[0] = sum of er[0..n_er) in a fixed
// order, tail[1] = rows -- written on the device so that the step needs no ATen reduction / fill kernels.
__global__ void __launch_bounds__(256) score_finalize_kernel(const float* __restrict__ partial, int row_chunks,
                                                             int64_t rows, int64_t C, float* __restrict__ sums_out,
                                                             float* __restrict__ score_out,
                                                             const float* __restrict__ er, int n_er,
                                                             float* __restrict__ tail) {
  __shared__ float red[8][33];
  const int cl = threadIdx.x & 31, g = threadIdx.x >> 5;
  if (tail != nullptr && blockIdx.x == 0 && blockIdx.y == 0) {
    __shared__ float ered[256];
    float a = 0.f;
    for (int i = threadIdx.x; i < n_er; i += 256) a += er[i];
    ered[threadIdx.x] = a;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
      if (threadIdx.x < st) ered[threadIdx.x] += ered[threadIdx.x + st];
      __syncthreads();
    }
    if (threadIdx.x == 0) { tail[0] = er ? ered[0] : 0.f; tail[1] = float(rows); }
  }
  const int64_t c = int64_t(blockIdx.x) * 32 + cl;
  float s = 0.f;
  if (c < C) {
    const float* p = partial + int64_t(blockIdx.y) * row_chunks * C + c;
    const int per = (row_chunks + 7) / 8;
    const int i0 = g * per, i1 = min(row_chunks, i0 + per);
    for (int i = i0; i < i1; ++i) s += p[int64_t(i) * C];
  }
  red[g][cl] = s;
  __syncthreads();
  if (g == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][cl];
    if (sums_out) sums_out[blockIdx.y * C + c] = t;
    if (score_out) score_out[blockIdx.y * C + c] = t / float(rows);
  }
}

// ------------------------------------------------------------------------------
// a4: bottom-k.  64-bit keys (order-preserving score bits << 32 | channel), so
// the order is total: ascending score, ties -> lower channel, NaN last.
// Bitonic network; exchange distances < 32 run on warp shuffles.
// ------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long make_key(float s, uint32_t idx) {
  uint32_t b = __float_as_uint(s);
  if (s != s) b = 0xffffffffu;                       // NaN last
  else if (s == 0.f) b = 0x80000000u;                // -0 == +0
  else b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return (static_cast<unsigned long long>(b) << 32) | idx;
}

// denom != nullptr: the keys are score[i] / *denom (IEEE division, as torch.div would produce) -- the score sums of
// the packed statistic divided by the global row count, without a separate elementwise kernel.
__global__ void __launch_bounds__(1024) bottomk_kernel(const float* __restrict__ score, int64_t C, int64_t k, int n2,
                                                       int64_t* __restrict__ idx_out, const float* __restrict__ denom,
                                                       float* __restrict__ score_out) {
  extern __shared__ unsigned long long keys[];
  const float* sraw = score + int64_t(blockIdx.x) * C;
  const float dn = denom ? *denom : 1.f;
  struct Scaled {
    const float* p; float d; bool on;
    __device__ float operator[](int i) const { return on ? __fdiv_rn(p[i], d) : p[i]; }
  } s{sraw, dn, denom != nullptr};
  if (score_out) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) score_out[int64_t(blockIdx.x) * C + i] = s[i];
  }
  int64_t* out = idx_out + int64_t(blockIdx.x) * k;
  const int tid = threadIdx.x;
  if (n2 <= 1024) {
    // one key per thread, register resident
    unsigned long long key = (tid < C) ? make_key(s[tid], uint32_t(tid)) : ~0ull;
    for (int kk = 2; kk <= n2; kk <<= 1) {
      for (int j = kk >> 1; j > 0; j >>= 1) {
        unsigned long long other;
        if (j >= 32) {
          keys[tid] = key;
          __syncthreads();
          other = keys[tid ^ j];
          __syncthreads();
        } else {
          other = __shfl_xor_sync(0xffffffffu, key, j);
        }
        const bool asc = (tid & kk) == 0, lower = (tid & j) == 0;
        const unsigned long long lo = key < other ? key : other, hi = key < other ? other : key;
        key = (lower == asc) ? lo : hi;
      }
    }
    if (tid < k) out[tid] = int64_t(key & 0xffffffffull);
  } else {
    for (int i = tid; i < n2; i += blockDim.x) keys[i] = (i < C) ? make_key(s[i], uint32_t(i)) : ~0ull;
    __syncthreads();
    for (int kk = 2; kk <= n2; kk <<= 1) {
      for (int j = kk >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < n2; i += blockDim.x) {
          int p = i ^ j;
          if (p > i) {
            const bool asc = (i & kk) == 0;
            unsigned long long a = keys[i], b = keys[p];
            if ((a > b) == asc) { keys[i] = b; keys[p] = a; }
          }
        }
        __syncthreads();
      }
    }
    for (int i = tid; i < k; i += blockDim.x) out[i] = int64_t(keys[i] & 0xffffffffull);
  }
}

// ------------------------------------------------------------------------------
// Channel-membership bits for this CTA's column chunk, built in shared memory
// from the two index lists (bit 0: c in S_rgb, bit 1: c in S_depth).
// ------------------------------------------------------------------------------
template <int V>
__device__ __forceinline__ void build_sel(int* sel, int64_t c0, int TXV, const int64_t* __restrict__ idx_r,
                                          const int64_t* __restrict__ idx_d, int64_t k, uint32_t& mr, uint32_t& md,
                                          int tx) {
  for (int i = threadIdx.x; i < TXV; i += 256) sel[i] = 0;
  __syncthreads();
  for (int64_t i = threadIdx.x; i < k; i += 256) {
    int64_t a = idx_r[i] - c0, b = idx_d[i] - c0;
    if (a >= 0 && a < TXV) atomicOr(&sel[a], 1);
    if (b >= 0 && b < TXV) atomicOr(&sel[b], 2);
  }
  __syncthreads();
  mr = 0; md = 0;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    int s = sel[tx * V + i];
    mr |= uint32_t(s & 1) << i;
    md |= uint32_t((s >> 1) & 1) << i;
  }
}

template <int BLEND>
__device__ __forceinline__ float blend_val(float own, float other, float a) {
  // explicit _rn intrinsics: no FMA contraction, so fp32 results equal the
  // reference's op-by-op arithmetic bit for bit
  if (BLEND == R3D_BLEND_SWAP) return other;
  if (BLEND == R3D_BLEND_SCALE) return __fmul_rn(a, other);
  return __fadd_rn(__fmul_rn(a, own), __fmul_rn(__fsub_rn(1.f, a), other));
}

// ------------------------------------------------------------------------------
// a5/a6/a7: exchange + stack, writes (rows, 2, C) directly
// ------------------------------------------------------------------------------
template <typename T, int V, int BLEND, bool AFFINE>
__global__ void __launch_bounds__(256) exchange_fwd_kernel(const T* __restrict__ rgb, const T* __restrict__ depth,
                                                           const int64_t* __restrict__ idx_r,
                                                           const int64_t* __restrict__ idx_d, int64_t k,
                                                           const float* __restrict__ alpha,
                                                           const float* __restrict__ affine, T* __restrict__ out,
                                                           int64_t rows, int64_t C, int tx_log2,
                                                           int64_t rows_per_cta) {
  __shared__ int sel[256 * V];
  const int TX = 1 << tx_log2, TY = 256 >> tx_log2;
  const int tx = threadIdx.x & (TX - 1), ty = threadIdx.x >> tx_log2;
  const int64_t c0 = int64_t(blockIdx.y) * TX * V;
  uint32_t mr, md;
  build_sel<V>(sel, c0, TX * V, idx_r, idx_d, k, mr, md, tx);
  const int64_t c = c0 + int64_t(tx) * V;
  if (c >= C) return;
  float a[V];
  float sc_r[V], sh_r[V], sc_d[V], sh_d[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    a[i] = (BLEND != R3D_BLEND_SWAP) ? alpha[c + i] : 0.f;
    if (AFFINE) {
      sc_r[i] = affine[c + i];
      sh_r[i] = affine[C + c + i];
      sc_d[i] = affine[2 * C + c + i];
      sh_d[i] = affine[3 * C + c + i];
    }
  }
  const int64_t r0 = int64_t(blockIdx.x) * rows_per_cta;
  const int64_t r1 = min(rows, r0 + rows_per_cta);
  const T* pr = rgb + c;
  const T* pd = depth + c;
  T* po = out + c;

  auto body = [&](const float (&r)[V], const float (&d)[V], int64_t row) {
    float o_r[V], o_d[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float rv = r[i], dv = d[i];
      if (AFFINE) {
        rv = fmaf(rv, sc_r[i], sh_r[i]);
        dv = fmaf(dv, sc_d[i], sh_d[i]);
      }
      o_r[i] = ((mr >> i) & 1u) ? blend_val<BLEND>(rv, dv, a[i]) : rv;
      o_d[i] = ((md >> i) & 1u) ? blend_val<BLEND>(dv, rv, a[i]) : dv;
    }
    store_vec<T, V>(po + row * 2 * C, o_r);
    store_vec<T, V>(po + row * 2 * C + C, o_d);
  };

  int64_t r = r0 + ty;
  for (; r + 3 * TY < r1; r += 4 * TY) {   // four rows = eight independent loads in flight
    float ra[4][V], da[4][V];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      load_vec<T, V>(pr + (r + u * TY) * C, ra[u]);
      load_vec<T, V>(pd + (r + u * TY) * C, da[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) body(ra[u], da[u], r + u * TY);
  }
  for (; r < r1; r += TY) {
    float ra[V], da[V];
    load_vec<T, V>(pr + r * C, ra);
    load_vec<T, V>(pd + r * C, da);
    body(ra, da, r);
  }
}

// ------------------------------------------------------------------------------
// a8: exchange backward.  NSUM = number of per-channel sums accumulated
// (0: swap, 1: d_alpha, 5: d_alpha + the four BatchNorm backward sums).
// ------------------------------------------------------------------------------
template <typename T, int V, int BLEND, bool AFFINE, bool BN>
__global__ void __launch_bounds__(256) exchange_bwd_kernel(
    const T* __restrict__ g, const T* __restrict__ rgb, const T* __restrict__ depth,
    const int64_t* __restrict__ idx_r, const int64_t* __restrict__ idx_d, int64_t k, const float* __restrict__ alpha,
    const float* __restrict__ affine, const float* __restrict__ bn_norm, T* __restrict__ d_rgb,
    T* __restrict__ d_depth, float* __restrict__ colsum_partial, int64_t rows, int64_t C, int tx_log2,
    int64_t rows_per_cta) {
  constexpr int NSUM = (BLEND == R3D_BLEND_SWAP) ? 0 : (BN ? 5 : 1);
  __shared__ int sel[256 * V];
  __shared__ float red[(NSUM > 0 ? 256 * V : 1)];
  const int TX = 1 << tx_log2, TY = 256 >> tx_log2;
  const int tx = threadIdx.x & (TX - 1), ty = threadIdx.x >> tx_log2;
  const int64_t c0 = int64_t(blockIdx.y) * TX * V;
  uint32_t mr, md;
  build_sel<V>(sel, c0, TX * V, idx_r, idx_d, k, mr, md, tx);
  const int64_t c = c0 + int64_t(tx) * V;
  const bool active = c < C;
  float acc[NSUM > 0 ? NSUM : 1][V];
#pragma unroll
  for (int s = 0; s < (NSUM > 0 ? NSUM : 1); ++s)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[s][i] = 0.f;

  if (active) {
    float a[V];
#pragma unroll
    for (int i = 0; i < V; ++i) a[i] = (BLEND != R3D_BLEND_SWAP) ? alpha[c + i] : 0.f;
    const int64_t r0 = int64_t(blockIdx.x) * rows_per_cta;
    const int64_t r1 = min(rows, r0 + rows_per_cta);
    int64_t r = r0 + ty;
    if (BLEND == R3D_BLEND_SWAP) {
      // pure mask-select: keep two rows (four 128-bit loads) in flight
      for (; r + TY < r1; r += 2 * TY) {
        float gr4[2][V], gd4[2][V];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          load_vec<T, V>(g + (r + u * TY) * 2 * C + c, gr4[u]);
          load_vec<T, V>(g + (r + u * TY) * 2 * C + C + c, gd4[u]);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          float dr[V], dd[V];
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const bool ir = (mr >> i) & 1u, id = (md >> i) & 1u;
            dr[i] = (ir ? 0.f : gr4[u][i]) + (id ? gd4[u][i] : 0.f);
            dd[i] = (id ? 0.f : gd4[u][i]) + (ir ? gr4[u][i] : 0.f);
          }
          store_vec<T, V>(d_rgb + (r + u * TY) * C + c, dr);
          store_vec<T, V>(d_depth + (r + u * TY) * C + c, dd);
        }
      }
    }
    for (; r < r1; r += TY) {
      float gr[V], gd[V], dr[V], dd[V];
      load_vec<T, V>(g + r * 2 * C + c, gr);
      load_vec<T, V>(g + r * 2 * C + C + c, gd);
      if (BLEND == R3D_BLEND_SWAP) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const bool ir = (mr >> i) & 1u, id = (md >> i) & 1u;
          dr[i] = (ir ? 0.f : gr[i]) + (id ? gd[i] : 0.f);
          dd[i] = (id ? 0.f : gd[i]) + (ir ? gr[i] : 0.f);
        }
      } else {
        float xr[V], xd[V];
        load_vec<T, V>(rgb + r * C + c, xr);
        load_vec<T, V>(depth + r * C + c, xd);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const bool ir = (mr >> i) & 1u, id = (md >> i) & 1u;
          float rv = xr[i], dv = xd[i];
          if (AFFINE) {
            rv = fmaf(rv, affine[c + i], affine[C + c + i]);
            dv = fmaf(dv, affine[2 * C + c + i], affine[3 * C + c + i]);
          }
          if (BLEND == R3D_BLEND_SCALE) {
            dr[i] = (ir ? 0.f : gr[i]) + (id ? a[i] * gd[i] : 0.f);
            dd[i] = (id ? 0.f : gd[i]) + (ir ? a[i] * gr[i] : 0.f);
            acc[0][i] += (ir ? gr[i] * dv : 0.f) + (id ? gd[i] * rv : 0.f);
          } else {
            dr[i] = gr[i] * (ir ? a[i] : 1.f) + (id ? gd[i] * (1.f - a[i]) : 0.f);
            dd[i] = gd[i] * (id ? a[i] : 1.f) + (ir ? gr[i] * (1.f - a[i]) : 0.f);
            acc[0][i] += (ir ? gr[i] * (rv - dv) : 0.f) + (id ? gd[i] * (dv - rv) : 0.f);
          }
          if (BN) {
            const float nr = fmaf(xr[i], bn_norm[c + i], bn_norm[C + c + i]);
            const float nd = fmaf(xd[i], bn_norm[2 * C + c + i], bn_norm[3 * C + c + i]);
            acc[1][i] += dr[i];
            acc[2][i] += dr[i] * nr;
            acc[3][i] += dd[i];
            acc[4][i] += dd[i] * nd;
          }
        }
      }
      store_vec<T, V>(d_rgb + r * C + c, dr);
      store_vec<T, V>(d_depth + r * C + c, dd);
    }
  }
  if (NSUM > 0) {
    const int row_chunks = gridDim.x;
    for (int s = 0; s < NSUM; ++s) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < V; ++i) red[(ty * TX + tx) * V + i] = acc[s][i];
      __syncthreads();
      if (ty == 0 && active) {
        float* out = colsum_partial + (int64_t(s) * row_chunks + blockIdx.x) * C + c;
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float t = 0.f;
          for (int y = 0; y < TY; ++y) t += red[(y * TX + tx) * V + i];
          out[i] = t;
        }
      }
    }
  }
}

__global__ void colsum_finalize_kernel(const float* __restrict__ partial, int row_chunks, int64_t C,
                                       float* __restrict__ out) {
  int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float* p = partial + int64_t(blockIdx.y) * row_chunks * C + c;
  float s = 0.f;
  for (int i = 0; i < row_chunks; ++i) s += p[int64_t(i) * C];
  out[blockIdx.y * C + c] = s;
}

// ------------------------------------------------------------------------------
// a3: BatchNorm statistics.  Per-thread Welford over its rows, Chan merge across
// the CTA's row lanes, per-CTA (n, mean, M2) partials, fixed-order final merge.
// ------------------------------------------------------------------------------
__device__ __forceinline__ void chan_merge(float& n, float& mean, float& m2, float nb, float meanb, float m2b) {
  if (nb == 0.f) return;
  const float nt = n + nb, delta = meanb - mean;
  mean += delta * (nb / nt);
  m2 += m2b + delta * delta * (n * nb / nt);
  n = nt;
}

template <typename T, int V>
__global__ void __launch_bounds__(256) bn_partial_kernel(const T* __restrict__ rgb, const T* __restrict__ depth,
                                                         int64_t rows, int64_t C, int tx_log2, int64_t rows_per_cta,
                                                         float* __restrict__ partial) {
  __shared__ float s_mean[256 * V], s_m2[256 * V], s_n[256];
  const int TX = 1 << tx_log2, TY = 256 >> tx_log2;
  const int tx = threadIdx.x & (TX - 1), ty = threadIdx.x >> tx_log2;
  const int64_t cv = int64_t(blockIdx.y) * TX + tx;
  const bool active = cv * V < C;
  const T* __restrict__ x = blockIdx.z ? depth : rgb;
  const int64_t r0 = int64_t(blockIdx.x) * rows_per_cta;
  const int64_t r1 = min(rows, r0 + rows_per_cta);
  float mean[V], m2[V], n = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) { mean[i] = 0.f; m2[i] = 0.f; }
  if (active) {
    const T* p = x + cv * V;
    for (int64_t r = r0 + ty; r < r1; r += TY) {
      float a[V];
      load_vec_keep<T, V>(p + r * C, a);
      n += 1.f;
      const float inv = 1.f / n;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float d = a[i] - mean[i];
        mean[i] += d * inv;
        m2[i] += d * (a[i] - mean[i]);
      }
    }
  }
  s_n[ty * TX + tx] = n;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    s_mean[(ty * TX + tx) * V + i] = mean[i];
    s_m2[(ty * TX + tx) * V + i] = m2[i];
  }
  __syncthreads();
  if (ty == 0 && active) {
    // layout: [mod][stat(3: n, mean, m2)][row_chunk][C]
    const int64_t rc = gridDim.x;
    float* base = partial + int64_t(blockIdx.z) * 3 * rc * C;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float nn = 0.f, mm = 0.f, qq = 0.f;
      for (int y = 0; y < TY; ++y)
        chan_merge(nn, mm, qq, s_n[y * TX + tx], s_mean[(y * TX + tx) * V + i], s_m2[(y * TX + tx) * V + i]);
      const int64_t c = cv * V + i;
      base[(0 * rc + blockIdx.x) * C + c] = nn;
      base[(1 * rc + blockIdx.x) * C + c] = mm;
      base[(2 * rc + blockIdx.x) * C + c] = qq;
    }
  }
}

__global__ void bn_finalize_kernel(const float* __restrict__ partial, int row_chunks, int64_t C,
                                   float* __restrict__ stats_out) {
  int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int64_t rc = row_chunks;
  const float* base = partial + int64_t(blockIdx.y) * 3 * rc * C;
  float n = 0.f, mean = 0.f, m2 = 0.f;
  for (int i = 0; i < row_chunks; ++i)
    chan_merge(n, mean, m2, base[(0 * rc + i) * C + c], base[(1 * rc + i) * C + c], base[(2 * rc + i) * C + c]);
  float* o = stats_out + int64_t(blockIdx.y) * 3 * C;
  o[c] = mean;
  o[C + c] = m2 / n;
  o[2 * C + c] = m2 / fmaxf(n - 1.f, 1.f);
}

// dx = gamma*rstd*(dy - sum(dy)/N - xn*sum(dy*xn)/N), in place on dy.
template <typename T, int V>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const T* __restrict__ rgb, const T* __restrict__ depth,
                                                           const float* __restrict__ bn_norm,
                                                           const float* __restrict__ gamma_r,
                                                           const float* __restrict__ gamma_d,
                                                           const float* __restrict__ colsums, T* __restrict__ d_rgb,
                                                           T* __restrict__ d_depth, int64_t rows, int64_t C,
                                                           int tx_log2, int64_t rows_per_cta) {
  const int TX = 1 << tx_log2, TY = 256 >> tx_log2;
  const int tx = threadIdx.x & (TX - 1), ty = threadIdx.x >> tx_log2;
  const int64_t c = (int64_t(blockIdx.y) * TX + tx) * V;
  if (c >= C) return;
  const int mod = blockIdx.z;
  const T* __restrict__ x = mod ? depth : rgb;
  T* __restrict__ dy = mod ? d_depth : d_rgb;
  const float* gamma = mod ? gamma_d : gamma_r;
  const float invN = 1.f / float(rows);
  float rstd[V], nsh[V], k0[V], k1[V], k2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    rstd[i] = bn_norm[(2 * mod) * C + c + i];
    nsh[i] = bn_norm[(2 * mod + 1) * C + c + i];
    k0[i] = gamma[c + i] * rstd[i];
    k1[i] = colsums[(1 + 2 * mod) * C + c + i] * invN;
    k2[i] = colsums[(2 + 2 * mod) * C + c + i] * invN;
  }
  const int64_t r0 = int64_t(blockIdx.x) * rows_per_cta;
  const int64_t r1 = min(rows, r0 + rows_per_cta);
  for (int64_t r = r0 + ty; r < r1; r += TY) {
    float xv[V], gv[V], o[V];
    load_vec<T, V>(x + r * C + c, xv);
    load_vec<T, V>(dy + r * C + c, gv);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float xn = fmaf(xv[i], rstd[i], nsh[i]);
      o[i] = k0[i] * (gv[i] - k1[i] - xn * k2[i]);
    }
    store_vec<T, V>(dy + r * C + c, o);
  }
}

}  // namespace r3d

// ================================================================================
// C ABI
// ================================================================================
using namespace r3d;

extern "C" const char* r3d_last_error(void) { return g_err; }
extern "C" int r3d_abi_version(void) { return 2; }
extern "C" int r3d_profile_enable(int on) {
  const int prev = g_prof ? 1 : 0;
  g_prof = on != 0;
  return prev;
}
extern "C" int r3d_profile_num_stages(void) { return ST_NUM; }
extern "C" const char* r3d_profile_stage_name(int stage) {
  static const char* names[ST_NUM] = {"score_partial", "score_finalize", "bottomk", "exchange_fwd", "exchange_bwd",
                                      "colsum_finalize", "bn_stats", "bn_bwd", "gram", "jacobi_init", "jacobi_inner",
                                      "jacobi_update", "jacobi_extract", "refine_y", "sigma", "entropy", "coef",
                                      "bwd_gemm", "token_info", "block", "jacobi_vupdate", "jacobi_local"};
  return (stage >= 0 && stage < ST_NUM) ? names[stage] : "?";
}
extern "C" int r3d_profile_read(double* ms_out, int64_t* calls_out, int64_t* launches_out, int reset) {
  drain_records();
  for (int i = 0; i < ST_NUM; ++i) {
    if (ms_out) ms_out[i] = g_ms[i];
    if (calls_out) calls_out[i] = g_cnt[i];
    if (launches_out) launches_out[i] = g_kl[i];
    if (reset) { g_ms[i] = 0; g_cnt[i] = 0; g_kl[i] = 0; }
  }
  return 0;
}
extern "C" int64_t r3d_launch_count(int reset) {
  int64_t v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

static int check_common(const void* a, const void* b, int64_t rows, int64_t C, int dtype) {
  R3D_CHECK(a != nullptr && b != nullptr, "null tensor pointer");
  R3D_CHECK(rows >= 0 && C >= 1, "bad shape rows=%lld C=%lld", (long long)rows, (long long)C);
  R3D_CHECK(dtype == R3D_F32 || dtype == R3D_BF16, "dtype must be R3D_F32 or R3D_BF16, got %d", dtype);
  return 0;
}

extern "C" size_t r3d_score_workspace_floats(int64_t rows, int64_t C) {
  return size_t(2) * fixed_row_chunks(rows) * size_t(C);
}

template <typename T>
static int score_partial_t(const void* rgb, const void* depth, int64_t rows, int64_t C, float* partial,
                           cudaStream_t st) {
  const bool vec = vec_ok<T>(rgb, C) && vec_ok<T>(depth, C);
  constexpr int VN = VecOf<T>::N;
  Grid2 g = make_grid(rows, C, vec ? VN : 1);
  dim3 grid(g.row_chunks, g.col_chunks, 2);
  R3D_STAGE(ST_SCORE_PARTIAL, st);
  if (vec)
    score_partial_kernel<T, VN><<<grid, 256, 0, st>>>((const T*)rgb, (const T*)depth, rows, C, g.tx_log2,
                                                      g.rows_per_cta, partial);
  else
    score_partial_kernel<T, 1><<<grid, 256, 0, st>>>((const T*)rgb, (const T*)depth, rows, C, g.tx_log2,
                                                     g.rows_per_cta, partial);
  R3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int r3d_channel_score_partial(const void* rgb, const void* depth, int64_t rows, int64_t C, int dtype,
                                         float* partial, void* stream) {
  if (int e = check_common(rgb, depth, rows, C, dtype)) return e;
  R3D_CHECK(partial != nullptr, "null workspace");
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == R3D_F32 ? score_partial_t<float>(rgb, depth, rows, C, partial, st)
                          : score_partial_t<__nv_bfloat16>(rgb, depth, rows, C, partial, st);
}

extern "C" int r3d_score_finalize(const float* partial, int64_t rows, int64_t C, float* sums_out, float* score_out,
                                  void* stream) {
  R3D_CHECK(partial != nullptr, "null workspace");
  R3D_CHECK(rows >= 1 && C >= 1, "bad shape");
  dim3 grid((unsigned)((C + 31) / 32), 2);
  R3D_STAGE(ST_SCORE_FINALIZE, (cudaStream_t)stream);
  score_finalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(partial, fixed_row_chunks(rows), rows, C, sums_out,
                                                               score_out, nullptr, 0, nullptr);
  R3D_LAUNCH_CHECK();
  return 0;
}

// Packed statistic from column-sum partials that other kernels produced as a by-product (the GEMM epilogue of the RGB
// embedding, the LayerNorm+ReLU kernel of the depth projection): blockIdx.y = modality, its own number of partial rows.
__global__ void __launch_bounds__(256) score_pack_kernel(const float* __restrict__ part_r, int parts_r,
                                                         const float* __restrict__ part_d, int parts_d, int64_t rows,
                                                         int64_t C, const float* __restrict__ er, int n_er,
                                                         float* __restrict__ packed) {
  __shared__ float red[8][33];
  const int cl = threadIdx.x & 31, g = threadIdx.x >> 5;
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    __shared__ float ered[256];
    float a = 0.f;
    for (int i = threadIdx.x; i < n_er; i += 256) a += er[i];
    ered[threadIdx.x] = a;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
      if (threadIdx.x < st) ered[threadIdx.x] += ered[threadIdx.x + st];
      __syncthreads();
    }
    if (threadIdx.x == 0) { packed[2 * C] = er ? ered[0] : 0.f; packed[2 * C + 1] = float(rows); }
  }
  const float* p = blockIdx.y ? part_d : part_r;
  const int parts = blockIdx.y ? parts_d : parts_r;
  const int64_t c = int64_t(blockIdx.x) * 32 + cl;
  float s = 0.f;
  if (c < C) {
    const int per = (parts + 7) / 8;
    const int i0 = g * per, i1 = min(parts, i0 + per);
    for (int i = i0; i < i1; ++i) s += p[int64_t(i) * C + c];
  }
  red[g][cl] = s;
  __syncthreads();
  if (g == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][cl];
    packed[blockIdx.y * C + c] = t;
  }
}

extern "C" int r3d_score_pack(const float* part_r, int64_t parts_r, const float* part_d, int64_t parts_d, int64_t rows,
                              int64_t C, const float* er, int64_t n_er, float* packed_out, void* stream) {
  R3D_CHECK(part_r && part_d && packed_out, "null pointer");
  R3D_CHECK(parts_r >= 1 && parts_d >= 1 && rows >= 1 && C >= 1 && n_er >= 0, "bad shape");
  dim3 grid((unsigned)((C + 31) / 32), 2);
  R3D_STAGE(ST_SCORE_FINALIZE, (cudaStream_t)stream);
  score_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(part_r, (int)parts_r, part_d, (int)parts_d, rows, C, er,
                                                           (int)n_er, packed_out);
  R3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int r3d_score_finalize_packed(const float* partial, int64_t rows, int64_t C, const float* er, int64_t n_er,
                                         float* packed_out, void* stream) {
  R3D_CHECK(partial != nullptr && packed_out != nullptr, "null pointer");
  R3D_CHECK(rows >= 1 && C >= 1 && n_er >= 0, "bad shape");
  dim3 grid((unsigned)((C + 31) / 32), 2);
  R3D_STAGE(ST_SCORE_FINALIZE, (cudaStream_t)stream);
  score_finalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(partial, fixed_row_chunks(rows), rows, C, packed_out,
                                                               nullptr, er, (int)n_er, packed_out + 2 * C);
  R3D_LAUNCH_CHECK();
  return 0;
}

static int bottomk_impl(const float* score, int nvec, int64_t C, int64_t k, int64_t* idx_out, const float* denom,
                        float* score_out, void* stream);
extern "C" int r3d_bottomk(const float* score, int nvec, int64_t C, int64_t k, int64_t* idx_out, void* stream) {
  return bottomk_impl(score, nvec, C, k, idx_out, nullptr, nullptr, stream);
}
extern "C" int r3d_bottomk_scaled(const float* sums, int nvec, int64_t C, int64_t k, const float* denom,
                                  int64_t* idx_out, float* score_out, void* stream) {
  R3D_CHECK(denom != nullptr, "null denominator");
  return bottomk_impl(sums, nvec, C, k, idx_out, denom, score_out, stream);
}
static int bottomk_impl(const float* score, int nvec, int64_t C, int64_t k, int64_t* idx_out, const float* denom,
                        float* score_out, void* stream) {
  R3D_CHECK(score != nullptr, "null score");
  R3D_CHECK(C >= 1 && C <= 8192, "bottomk supports 1 <= C <= 8192, got %lld", (long long)C);
  R3D_CHECK(k >= 0 && k <= C, "k=%lld out of range for C=%lld", (long long)k, (long long)C);
  if (k == 0 || nvec == 0) return 0;
  R3D_CHECK(idx_out != nullptr, "null idx_out");
  int n2 = 32;
  while (n2 < C) n2 <<= 1;
  const int threads = n2 <= 1024 ? n2 : 1024;
  const size_t smem = size_t(n2) * sizeof(unsigned long long);
  if (smem > 48 * 1024)
    R3D_CUDA(cudaFuncSetAttribute(bottomk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  R3D_STAGE(ST_BOTTOMK, (cudaStream_t)stream);
  bottomk_kernel<<<nvec, threads, smem, (cudaStream_t)stream>>>(score, C, k, n2, idx_out, denom, score_out);
  R3D_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t r3d_bn_workspace_floats(int64_t rows, int64_t C) {
  return size_t(2) * 3 * fixed_row_chunks(rows) * size_t(C);
}

template <typename T>
static int bn_stats_t(const void* rgb, const void* depth, int64_t rows, int64_t C, float* ws, float* stats,
                      cudaStream_t st) {
  const bool vec = vec_ok<T>(rgb, C) && vec_ok<T>(depth, C);
  constexpr int VN = VecOf<T>::N;
  Grid2 g = make_grid(rows, C, vec ? VN : 1);
  dim3 grid(g.row_chunks, g.col_chunks, 2);
  R3D_STAGE(ST_BN_STATS, st);
  if (vec)
    bn_partial_kernel<T, VN><<<grid, 256, 0, st>>>((const T*)rgb, (const T*)depth, rows, C, g.tx_log2, g.rows_per_cta, ws);
  else
    bn_partial_kernel<T, 1><<<grid, 256, 0, st>>>((const T*)rgb, (const T*)depth, rows, C, g.tx_log2, g.rows_per_cta, ws);
  R3D_LAUNCH_CHECK();
  dim3 g2((unsigned)((C + 255) / 256), 2);
  bn_finalize_kernel<<<g2, 256, 0, st>>>(ws, g.row_chunks, C, stats);
  R3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int r3d_bn_stats(const void* rgb, const void* depth, int64_t rows, int64_t C, int dtype, float* workspace,
                            float* stats_out, void* stream) {
  if (int e = check_common(rgb, depth, rows, C, dtype)) return e;
  R3D_CHECK(rows >= 1, "BatchNorm statistics need at least one row");
  R3D_CHECK(workspace != nullptr && stats_out != nullptr, "null workspace/stats");
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == R3D_F32 ? bn_stats_t<float>(rgb, depth, rows, C, workspace, stats_out, st)
                          : bn_stats_t<__nv_bfloat16>(rgb, depth, rows, C, workspace, stats_out, st);
}

template <typename T, int V>
static int exchange_fwd_v(const void* rgb, const void* depth, const int64_t* idx_r, const int64_t* idx_d, int64_t k,
                          const float* alpha, const float* affine, int blend, void* out, int64_t rows, int64_t C,
                          cudaStream_t st) {
  Grid2 g = make_grid(rows, C, V);
  dim3 grid(g.row_chunks, g.col_chunks, 1);
  R3D_STAGE(ST_EXCHANGE_FWD, st);
#define R3D_FWD(BL, AF)                                                                                           \
  exchange_fwd_kernel<T, V, BL, AF><<<grid, 256, 0, st>>>((const T*)rgb, (const T*)depth, idx_r, idx_d, k, alpha, \
                                                          affine, (T*)out, rows, C, g.tx_log2, g.rows_per_cta)
  if (blend == R3D_BLEND_SWAP) {
    if (affine) R3D_FWD(R3D_BLEND_SWAP, true); else R3D_FWD(R3D_BLEND_SWAP, false);
  } else if (blend == R3D_BLEND_SCALE) {
    if (affine) R3D_FWD(R3D_BLEND_SCALE, true); else R3D_FWD(R3D_BLEND_SCALE, false);
  } else {
    if (affine) R3D_FWD(R3D_BLEND_CONVEX, true); else R3D_FWD(R3D_BLEND_CONVEX, false);
  }
#undef R3D_FWD
  R3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int r3d_exchange_fwd(const void* rgb, const void* depth, const int64_t* idx_r, const int64_t* idx_d,
                                int64_t k, const float* alpha, const float* affine, int blend, void* out,
                                int64_t rows, int64_t C, int dtype, void* stream) {
  if (int e = check_common(rgb, depth, rows, C, dtype)) return e;
  R3D_CHECK(out != nullptr, "null output");
  R3D_CHECK(blend >= 0 && blend <= 2, "bad blend mode %d", blend);
  R3D_CHECK(k >= 0 && k <= C, "k=%lld out of range", (long long)k);
  R3D_CHECK(k == 0 || (idx_r && idx_d), "null index pointer with k > 0");
  R3D_CHECK(blend == R3D_BLEND_SWAP || alpha != nullptr, "alpha required for blend mode %d", blend);
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == R3D_F32) {
    const bool vec = vec_ok<float>(rgb, C) && vec_ok<float>(depth, C) && vec_ok<float>(out, C);
    return vec ? exchange_fwd_v<float, 4>(rgb, depth, idx_r, idx_d, k, alpha, affine, blend, out, rows, C, st)
               : exchange_fwd_v<float, 1>(rgb, depth, idx_r, idx_d, k, alpha, affine, blend, out, rows, C, st);
  }
  using B = __nv_bfloat16;
  const bool vec = vec_ok<B>(rgb, C) && vec_ok<B>(depth, C) && vec_ok<B>(out, C);
  if (vec && !affine) return exchange_fwd_v<B, 8>(rgb, depth, idx_r, idx_d, k, alpha, affine, blend, out, rows, C, st);
  if (vec) return exchange_fwd_v<B, 4>(rgb, depth, idx_r, idx_d, k, alpha, affine, blend, out, rows, C, st);
  return exchange_fwd_v<B, 1>(rgb, depth, idx_r, idx_d, k, alpha, affine, blend, out, rows, C, st);
}

extern "C" size_t r3d_exchange_bwd_workspace_floats(int64_t rows, int64_t C) {
  return size_t(5) * fixed_row_chunks(rows) * size_t(C);
}

template <typename T, int V>
static int exchange_bwd_v(const void* g, const void* rgb, const void* depth, const int64_t* idx_r,
                          const int64_t* idx_d, int64_t k, const float* alpha, const float* affine,
                          const float* bn_norm, int blend, void* d_rgb, void* d_depth, float* part, int64_t rows,
                          int64_t C, cudaStream_t st) {
  Grid2 gg = make_grid(rows, C, V);
  dim3 grid(gg.row_chunks, gg.col_chunks, 1);
  R3D_STAGE(ST_EXCHANGE_BWD, st);
#define R3D_BWD(BL, AF, BN)                                                                                        \
  exchange_bwd_kernel<T, V, BL, AF, BN><<<grid, 256, 0, st>>>((const T*)g, (const T*)rgb, (const T*)depth, idx_r,  \
                                                              idx_d, k, alpha, affine, bn_norm, (T*)d_rgb,         \
                                                              (T*)d_depth, part, rows, C, gg.tx_log2, gg.rows_per_cta)
  if (blend == R3D_BLEND_SWAP) R3D_BWD(R3D_BLEND_SWAP, false, false);
  else if (blend == R3D_BLEND_SCALE) {
    if (affine) R3D_BWD(R3D_BLEND_SCALE, true, false); else R3D_BWD(R3D_BLEND_SCALE, false, false);
  } else {
    if (bn_norm) R3D_BWD(R3D_BLEND_CONVEX, true, true);
    else if (affine) R3D_BWD(R3D_BLEND_CONVEX, true, false);
    else R3D_BWD(R3D_BLEND_CONVEX, false, false);
  }
#undef R3D_BWD
  R3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int r3d_exchange_bwd(const void* g, const void* rgb, const void* depth, const int64_t* idx_r,
                                const int64_t* idx_d, int64_t k, const float* alpha, const float* affine,
                                const float* bn_norm, int blend, void* d_rgb, void* d_depth, float* colsum_partial,
                                int64_t rows, int64_t C, int dtype, void* stream) {
  R3D_CHECK(g && d_rgb && d_depth, "null gradient pointer");
  R3D_CHECK(rows >= 0 && C >= 1, "bad shape");
  R3D_CHECK(dtype == R3D_F32 || dtype == R3D_BF16, "bad dtype %d", dtype);
  R3D_CHECK(blend >= 0 && blend <= 2, "bad blend mode %d", blend);
  R3D_CHECK(k >= 0 && k <= C, "k out of range");
  R3D_CHECK(k == 0 || (idx_r && idx_d), "null index pointer with k > 0");
  if (blend != R3D_BLEND_SWAP)
    R3D_CHECK(rgb && depth && alpha && colsum_partial, "blend mode %d needs rgb, depth, alpha and a workspace", blend);
  R3D_CHECK(!bn_norm || (affine && blend == R3D_BLEND_CONVEX), "bn_norm requires affine and the convex blend");
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == R3D_F32) {
    bool vec = vec_ok<float>(g, C) && vec_ok<float>(d_rgb, C) && vec_ok<float>(d_depth, C);
    if (blend != R3D_BLEND_SWAP) vec = vec && vec_ok<float>(rgb, C) && vec_ok<float>(depth, C);
    return vec ? exchange_bwd_v<float, 4>(g, rgb, depth, idx_r, idx_d, k, alpha, affine, bn_norm, blend, d_rgb,
                                          d_depth, colsum_partial, rows, C, st)
               : exchange_bwd_v<float, 1>(g, rgb, depth, idx_r, idx_d, k, alpha, affine, bn_norm, blend, d_rgb,
                                          d_depth, colsum_partial, rows, C, st);
  }
  using B = __nv_bfloat16;
  bool vec = vec_ok<B>(g, C) && vec_ok<B>(d_rgb, C) && vec_ok<B>(d_depth, C);
  if (blend != R3D_BLEND_SWAP) vec = vec && vec_ok<B>(rgb, C) && vec_ok<B>(depth, C);
  if (vec && blend == R3D_BLEND_SWAP)
    return exchange_bwd_v<B, 8>(g, rgb, depth, idx_r, idx_d, k, alpha, affine, bn_norm, blend, d_rgb, d_depth,
                                colsum_partial, rows, C, st);
  if (vec)
    return exchange_bwd_v<B, 4>(g, rgb, depth, idx_r, idx_d, k, alpha, affine, bn_norm, blend, d_rgb, d_depth,
                                colsum_partial, rows, C, st);
  return exchange_bwd_v<B, 1>(g, rgb, depth, idx_r, idx_d, k, alpha, affine, bn_norm, blend, d_rgb, d_depth,
                              colsum_partial, rows, C, st);
}

extern "C" int r3d_exchange_bwd_finalize(const float* colsum_partial, int64_t rows, int64_t C, float* colsums_out,
                                         void* stream) {
  R3D_CHECK(colsum_partial && colsums_out, "null pointer");
  dim3 grid((unsigned)((C + 255) / 256), 5);
  R3D_STAGE(ST_COLSUM_FINALIZE, (cudaStream_t)stream);
  colsum_finalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(colsum_partial, fixed_row_chunks(rows), C,
                                                                colsums_out);
  R3D_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int bn_bwd_apply_t(const void* rgb, const void* depth, const float* bn_norm, const float* gr, const float* gd,
                          const float* colsums, void* d_rgb, void* d_depth, int64_t rows, int64_t C,
                          cudaStream_t st) {
  const bool vec = vec_ok<T>(rgb, C) && vec_ok<T>(depth, C) && vec_ok<T>(d_rgb, C) && vec_ok<T>(d_depth, C);
  R3D_STAGE(ST_BN_BWD, st);
  if (vec) {
    Grid2 g = make_grid(rows, C, 4);
    dim3 grid(g.row_chunks, g.col_chunks, 2);
    bn_bwd_apply_kernel<T, 4><<<grid, 256, 0, st>>>((const T*)rgb, (const T*)depth, bn_norm, gr, gd, colsums,
                                                    (T*)d_rgb, (T*)d_depth, rows, C, g.tx_log2, g.rows_per_cta);
  } else {
    Grid2 g = make_grid(rows, C, 1);
    dim3 grid(g.row_chunks, g.col_chunks, 2);
    bn_bwd_apply_kernel<T, 1><<<grid, 256, 0, st>>>((const T*)rgb, (const T*)depth, bn_norm, gr, gd, colsums,
                                                    (T*)d_rgb, (T*)d_depth, rows, C, g.tx_log2, g.rows_per_cta);
  }
  R3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int r3d_bn_bwd_apply(const void* rgb, const void* depth, const float* bn_norm, const float* gamma_r,
                                const float* gamma_d, const float* colsums, void* d_rgb, void* d_depth, int64_t rows,
                                int64_t C, int dtype, void* stream) {
  if (int e = check_common(rgb, depth, rows, C, dtype)) return e;
  R3D_CHECK(bn_norm && gamma_r && gamma_d && colsums && d_rgb && d_depth, "null pointer");
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == R3D_F32
             ? bn_bwd_apply_t<float>(rgb, depth, bn_norm, gamma_r, gamma_d, colsums, d_rgb, d_depth, rows, C, st)
             : bn_bwd_apply_t<__nv_bfloat16>(rgb, depth, bn_norm, gamma_r, gamma_d, colsums, d_rgb, d_depth, rows, C, st);
}

// Host-buffer entry point: H2D -> score -> finalize -> bottom-k -> exchange -> D2H.
extern "C" int r3d_token_fusion_host(const void* rgb_host, const void* depth_host, int64_t B, int64_t T, int64_t C,
                                     int dtype, int64_t k, void* out_host, int64_t* idx_r_host, int64_t* idx_d_host,
                                     void* stream) {
  R3D_CHECK(rgb_host && depth_host && out_host, "null host pointer");
  R3D_CHECK(dtype == R3D_F32 || dtype == R3D_BF16, "bad dtype %d", dtype);
  R3D_CHECK(B >= 1 && T >= 1 && C >= 1 && C <= 8192, "bad shape");
  R3D_CHECK(k >= 0 && k <= C, "k out of range");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = B * T;
  const size_t es = dtype == R3D_F32 ? 4 : 2;
  const size_t nb = size_t(rows) * C * es;
  const size_t wsf = r3d_score_workspace_floats(rows, C);
  char* buf = nullptr;
  const size_t off_d = nb, off_o = 2 * nb, off_ws = 4 * nb, off_sc = off_ws + wsf * 4, off_idx = off_sc + 2 * C * 4;
  const size_t total = off_idx + 2 * size_t(k > 0 ? k : 1) * 8;
  keep_async_pool();
  R3D_CUDA(cudaMallocAsync((void**)&buf, total, st));
  int rc = 0;
  do {
    if (cudaMemcpyAsync(buf, rgb_host, nb, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(buf + off_d, depth_host, nb, cudaMemcpyHostToDevice, st) != cudaSuccess) {
      set_error("H2D copy failed"); rc = 2; break;
    }
    float* ws = (float*)(buf + off_ws);
    float* sc = (float*)(buf + off_sc);
    int64_t* idx = (int64_t*)(buf + off_idx);
    if ((rc = r3d_channel_score_partial(buf, buf + off_d, rows, C, dtype, ws, st))) break;
    if ((rc = r3d_score_finalize(ws, rows, C, nullptr, sc, st))) break;
    if ((rc = r3d_bottomk(sc, 2, C, k, idx, st))) break;
    if ((rc = r3d_exchange_fwd(buf, buf + off_d, idx, idx + k, k, nullptr, nullptr, R3D_BLEND_SWAP, buf + off_o, rows,
                               C, dtype, st))) break;
    if (cudaMemcpyAsync(out_host, buf + off_o, 2 * nb, cudaMemcpyDeviceToHost, st) != cudaSuccess) {
      set_error("D2H copy failed"); rc = 2; break;
    }
    if (k > 0 && idx_r_host) cudaMemcpyAsync(idx_r_host, idx, k * 8, cudaMemcpyDeviceToHost, st);
    if (k > 0 && idx_d_host) cudaMemcpyAsync(idx_d_host, idx + k, k * 8, cudaMemcpyDeviceToHost, st);
  } while (0);
  cudaFreeAsync(buf, st);
  cudaError_t e = cudaStreamSynchronize(st);
  if (rc == 0 && e != cudaSuccess) {
    set_error("stream sync failed: %s", cudaGetErrorString(e));
    rc = 2;
  }
  return rc;
}
