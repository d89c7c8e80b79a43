// M-modality fuser (BASELINE.json configs[4]: RGB + depth + gaze, T = 2048, D = 1024).
//
// The reference fuser is written for a dict of modal features (`M = len(modal_feats)`,
// model/futr_safuser_tokenfusion.py:76) but hard-codes two of them (keys 'rgb' / 'depth' at :79, a 2 x 2 mask at :77);
// its only gaze consumer (model/futr_unsupervised_multimodal.py:16-32) is not a fuser.  The M-modality form fixed here
// (PARITY UNPINNED beyond M = 2, where it reproduces the reference; oracle: oracle/torch_port.py:PortCMFuserM):
//   * exchange: modality m replaces its k lowest-score channels by the same channels of modality (m + 1) mod M --
//     for M = 2 exactly tokenfusion.py:56-62 -- and the M streams are stacked to (B, T, M, C);
//   * Block: self-attention over the M modality tokens of every (b, t) with the reference's -inf diagonal mask
//     (tokenfusion.py:68-72 generalised to M x M): token m attends to the other M - 1 tokens.  For M = 2 that is the
//     closed form of SURVEY.md F4; for M >= 3 it is a real (M - 1)-way softmax per head, computed by the kernels below
//     straight from the qkv GEMM output -- M(M-1) dot products of head_dim per (row, head), one warp each;
//   * the final LayerNorm is followed by the mean over the M tokens (tokenfusion.py:93-95).
//
// Kernels (all HBM-bound, 128-bit accesses):
//   exchange_one_fwd / _bwd   one output stream per launch (reads own + partner, writes a strided slice of the stack)
//   mtoken_attn_fwd / _bwd    qkv (R, M, 3C) -> out (R, M, C); backward recomputes the softmax weights
//   token_mean_fwd / _bwd     (R, M, C) <-> (R, C)
#include "common.cuh"

namespace r3d {

namespace {

// out[row, m_slot, c] = c in S ? other[row, c] : own[row, c]   (out row pitch = out_pitch elements)
template <typename T, int V>
__global__ void __launch_bounds__(256) exchange_one_fwd_kernel(const T* __restrict__ own, const T* __restrict__ other,
                                                               const int64_t* __restrict__ idx, int64_t k,
                                                               T* __restrict__ out, int64_t out_pitch, int64_t rows,
                                                               int64_t C) {
  extern __shared__ uint8_t sel[];          // C bytes: 1 = exchanged channel
  for (int64_t c = threadIdx.x; c < C; c += 256) sel[c] = 0;
  __syncthreads();
  for (int64_t i = threadIdx.x; i < k; i += 256) {
    const int64_t c = idx[i];
    if (c >= 0 && c < C) sel[c] = 1;
  }
  __syncthreads();
  const int64_t cv = C / V, total = rows * cv;
  for (int64_t e = int64_t(blockIdx.x) * 256 + threadIdx.x; e < total; e += int64_t(gridDim.x) * 256) {
    const int64_t row = e / cv, col = (e - row * cv) * V;
    float a[V], b[V], o[V];
    load_vec<T, V>(own + row * C + col, a);
    load_vec<T, V>(other + row * C + col, b);
#pragma unroll
    for (int i = 0; i < V; ++i) o[i] = sel[col + i] ? b[i] : a[i];
    store_vec<T, V>(out + row * out_pitch + col, o);
  }
}

// d_x_j[row, c] = (c in S_j ? 0 : g_j[row, c]) + (c in S_prev ? g_prev[row, c] : 0)
template <typename T, int V>
__global__ void __launch_bounds__(256) exchange_one_bwd_kernel(const T* __restrict__ g_own, const T* __restrict__ g_prev,
                                                               int64_t g_pitch, const int64_t* __restrict__ idx_own,
                                                               const int64_t* __restrict__ idx_prev, int64_t k,
                                                               T* __restrict__ dx, int64_t rows, int64_t C) {
  extern __shared__ uint8_t sel[];          // bit 0: c in S_own, bit 1: c in S_prev
  for (int64_t c = threadIdx.x; c < C; c += 256) sel[c] = 0;
  __syncthreads();
  // the two lists are written by disjoint passes, so plain byte stores suffice (no two threads touch the same bit set
  // of one byte concurrently within a pass; a barrier separates the passes)
  for (int64_t i = threadIdx.x; i < k; i += 256) { const int64_t c = idx_own[i]; if (c >= 0 && c < C) sel[c] = 1; }
  __syncthreads();
  for (int64_t i = threadIdx.x; i < k; i += 256) { const int64_t c = idx_prev[i]; if (c >= 0 && c < C) sel[c] |= 2; }
  __syncthreads();
  const int64_t cv = C / V, total = rows * cv;
  for (int64_t e = int64_t(blockIdx.x) * 256 + threadIdx.x; e < total; e += int64_t(gridDim.x) * 256) {
    const int64_t row = e / cv, col = (e - row * cv) * V;
    float a[V], b[V], o[V];
    load_vec<T, V>(g_own + row * g_pitch + col, a);
    load_vec<T, V>(g_prev + row * g_pitch + col, b);
#pragma unroll
    for (int i = 0; i < V; ++i) o[i] = ((sel[col + i] & 1) ? 0.f : a[i]) + ((sel[col + i] & 2) ? b[i] : 0.f);
    store_vec<T, V>(dx + row * C + col, o);
  }
}

// ---- M-token masked self-attention (diagonal masked out), one warp per (row, head) -------------------------
constexpr int MT_MAX = 4;          // modalities supported by the register layout
constexpr int HD_MAX = 8;          // head_dim / 32 elements per lane (head_dim <= 256)

template <typename T>
__device__ __forceinline__ float ldf(const T* p) { return float(*p); }

template <typename T>
__global__ void __launch_bounds__(256) mtoken_attn_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ out, int64_t rows,
                                                              int M, int C, int heads, float scale) {
  const int hd = C / heads, per = hd / 32;
  const int lane = threadIdx.x & 31;
  const int64_t unit0 = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
  for (int64_t unit = unit0; unit < rows * heads; unit += int64_t(gridDim.x) * 8) {
    const int64_t row = unit / heads;
    const int h = int(unit - row * heads);
    const T* base = qkv + row * M * 3 * C + h * hd + lane * per;        // token m: + m * 3C; q | k | v: + 0, C, 2C
    float q[MT_MAX][HD_MAX], kk[MT_MAX][HD_MAX], vv[MT_MAX][HD_MAX];
#pragma unroll
    for (int m = 0; m < MT_MAX; ++m) {
      if (m >= M) break;
#pragma unroll
      for (int e = 0; e < HD_MAX; ++e) {
        if (e >= per) break;
        q[m][e] = ldf(base + m * 3 * C + e);
        kk[m][e] = ldf(base + m * 3 * C + C + e);
        vv[m][e] = ldf(base + m * 3 * C + 2 * C + e);
      }
    }
#pragma unroll
    for (int m = 0; m < MT_MAX; ++m) {
      if (m >= M) break;
      float s[MT_MAX], mx = -3.0e38f;
#pragma unroll
      for (int j = 0; j < MT_MAX; ++j) {
        if (j >= M) break;
        float d = 0.f;
#pragma unroll
        for (int e = 0; e < HD_MAX; ++e) { if (e >= per) break; d = fmaf(q[m][e], kk[j][e], d); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        s[j] = d * scale;
        if (j != m) mx = fmaxf(mx, s[j]);
      }
      float den = 0.f;
#pragma unroll
      for (int j = 0; j < MT_MAX; ++j) { if (j >= M) break; s[j] = (j == m) ? 0.f : __expf(s[j] - mx); den += s[j]; }
      const float inv = 1.f / den;
      T* o = out + (row * M + m) * C + h * hd + lane * per;
#pragma unroll
      for (int e = 0; e < HD_MAX; ++e) {
        if (e >= per) break;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < MT_MAX; ++j) { if (j >= M) break; acc = fmaf(s[j] * inv, vv[j][e], acc); }
        o[e] = T(acc);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) mtoken_attn_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ dout,
                                                              T* __restrict__ dqkv, int64_t rows, int M, int C, int heads,
                                                              float scale) {
  const int hd = C / heads, per = hd / 32;
  const int lane = threadIdx.x & 31;
  const int64_t unit0 = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
  for (int64_t unit = unit0; unit < rows * heads; unit += int64_t(gridDim.x) * 8) {
    const int64_t row = unit / heads;
    const int h = int(unit - row * heads);
    const T* base = qkv + row * M * 3 * C + h * hd + lane * per;
    float q[MT_MAX][HD_MAX], kk[MT_MAX][HD_MAX], vv[MT_MAX][HD_MAX], go[MT_MAX][HD_MAX];
    float dq[MT_MAX][HD_MAX], dk[MT_MAX][HD_MAX], dv[MT_MAX][HD_MAX];
#pragma unroll
    for (int m = 0; m < MT_MAX; ++m) {
      if (m >= M) break;
#pragma unroll
      for (int e = 0; e < HD_MAX; ++e) {
        if (e >= per) break;
        q[m][e] = ldf(base + m * 3 * C + e);
        kk[m][e] = ldf(base + m * 3 * C + C + e);
        vv[m][e] = ldf(base + m * 3 * C + 2 * C + e);
        go[m][e] = ldf(dout + (row * M + m) * C + h * hd + lane * per + e);
        dq[m][e] = 0.f; dk[m][e] = 0.f; dv[m][e] = 0.f;
      }
    }
#pragma unroll
    for (int m = 0; m < MT_MAX; ++m) {
      if (m >= M) break;
      float w[MT_MAX], dw[MT_MAX], mx = -3.0e38f;
#pragma unroll
      for (int j = 0; j < MT_MAX; ++j) {
        if (j >= M) break;
        float d = 0.f, g = 0.f;
#pragma unroll
        for (int e = 0; e < HD_MAX; ++e) {
          if (e >= per) break;
          d = fmaf(q[m][e], kk[j][e], d);
          g = fmaf(go[m][e], vv[j][e], g);                       // d out_m / d w[m][j] = dout_m . v_j
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          d += __shfl_xor_sync(0xffffffffu, d, o);
          g += __shfl_xor_sync(0xffffffffu, g, o);
        }
        w[j] = d * scale; dw[j] = g;
        if (j != m) mx = fmaxf(mx, w[j]);
      }
      float den = 0.f;
#pragma unroll
      for (int j = 0; j < MT_MAX; ++j) { if (j >= M) break; w[j] = (j == m) ? 0.f : __expf(w[j] - mx); den += w[j]; }
      const float inv = 1.f / den;
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < MT_MAX; ++j) { if (j >= M) break; w[j] *= inv; dot = fmaf(w[j], dw[j], dot); }
#pragma unroll
      for (int j = 0; j < MT_MAX; ++j) {
        if (j >= M) break;
        const float ds = w[j] * (dw[j] - dot) * scale;           // softmax backward, then the 1/sqrt(hd) scale
#pragma unroll
        for (int e = 0; e < HD_MAX; ++e) {
          if (e >= per) break;
          dv[j][e] = fmaf(w[j], go[m][e], dv[j][e]);
          dq[m][e] = fmaf(ds, kk[j][e], dq[m][e]);
          dk[j][e] = fmaf(ds, q[m][e], dk[j][e]);
        }
      }
    }
    T* ob = dqkv + row * M * 3 * C + h * hd + lane * per;
#pragma unroll
    for (int m = 0; m < MT_MAX; ++m) {
      if (m >= M) break;
#pragma unroll
      for (int e = 0; e < HD_MAX; ++e) {
        if (e >= per) break;
        ob[m * 3 * C + e] = T(dq[m][e]);
        ob[m * 3 * C + C + e] = T(dk[m][e]);
        ob[m * 3 * C + 2 * C + e] = T(dv[m][e]);
      }
    }
  }
}

// out[r, c] = mean over m of x[r, m, c]  /  dx[r, m, c] = dy[r, c] / M
template <typename T, int V>
__global__ void __launch_bounds__(256) token_mean_fwd_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t rows,
                                                             int M, int64_t C) {
  const int64_t cv = C / V, total = rows * cv;
  const float invM = 1.f / float(M);
  for (int64_t e = int64_t(blockIdx.x) * 256 + threadIdx.x; e < total; e += int64_t(gridDim.x) * 256) {
    const int64_t row = e / cv, col = (e - row * cv) * V;
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
    for (int m = 0; m < M; ++m) {
      float a[V];
      load_vec<T, V>(x + (row * M + m) * C + col, a);
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] += a[i];
    }
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] *= invM;
    store_vec<T, V>(out + row * C + col, acc);
  }
}
template <typename T, int V>
__global__ void __launch_bounds__(256) token_mean_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int64_t rows,
                                                             int M, int64_t C) {
  const int64_t cv = C / V, total = rows * cv;
  const float invM = 1.f / float(M);
  for (int64_t e = int64_t(blockIdx.x) * 256 + threadIdx.x; e < total; e += int64_t(gridDim.x) * 256) {
    const int64_t row = e / cv, col = (e - row * cv) * V;
    float a[V];
    load_vec<T, V>(dy + row * C + col, a);
#pragma unroll
    for (int i = 0; i < V; ++i) a[i] *= invM;
    for (int m = 0; m < M; ++m) store_vec<T, V>(dx + (row * M + m) * C + col, a);
  }
}

inline int ew_grid(int64_t total) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, int64_t(kNumSMs) * 16));
}

}  // namespace

}  // namespace r3d

using namespace r3d;

static int multi_check(int64_t rows, int64_t C, int dtype) {
  R3D_CHECK(rows >= 0 && C >= 1, "bad shape");
  R3D_CHECK(dtype == R3D_F32 || dtype == R3D_BF16, "bad dtype %d", dtype);
  R3D_CHECK(C % (dtype == R3D_F32 ? 4 : 8) == 0, "C must be a multiple of the 128-bit vector width");
  R3D_CHECK(C <= 48 * 1024, "C too large");
  return 0;
}

// One stream of the M-modality exchange: out[row, c] (row pitch out_pitch elements, e.g. M * C for slot m of the
// stacked (rows, M, C) tensor, out already offset to the slot) = c in idx ? other[row, c] : own[row, c].
extern "C" int r3d_exchange_one_fwd(const void* own, const void* other, const int64_t* idx, int64_t k, void* out,
                                    int64_t out_pitch, int64_t rows, int64_t C, int dtype, void* stream) {
  if (int e = multi_check(rows, C, dtype)) return e;
  if (rows == 0) return 0;
  R3D_CHECK(own && other && out && (k == 0 || idx), "null pointer");
  R3D_CHECK(((uintptr_t(own) | uintptr_t(other) | uintptr_t(out)) & 15) == 0 && out_pitch % 8 == 0, "unaligned tensors");
  cudaStream_t st = (cudaStream_t)stream;
  R3D_STAGE(ST_EXCHANGE_FWD, st);
  if (dtype == R3D_F32)
    exchange_one_fwd_kernel<float, 4><<<ew_grid(rows * C / 4), 256, size_t(C), st>>>((const float*)own, (const float*)other,
                                                                                   idx, k, (float*)out, out_pitch, rows, C);
  else
    exchange_one_fwd_kernel<__nv_bfloat16, 8><<<ew_grid(rows * C / 8), 256, size_t(C), st>>>(
        (const __nv_bfloat16*)own, (const __nv_bfloat16*)other, idx, k, (__nv_bfloat16*)out, out_pitch, rows, C);
  R3D_LAUNCH_CHECK();
  return 0;
}

// Gradient of modality j: dx = g_own masked outside idx_own + g_prev inside idx_prev (prev = the modality whose
// partner j is); g_own / g_prev point at their slots of the stacked gradient, row pitch g_pitch.
extern "C" int r3d_exchange_one_bwd(const void* g_own, const void* g_prev, int64_t g_pitch, const int64_t* idx_own,
                                    const int64_t* idx_prev, int64_t k, void* dx, int64_t rows, int64_t C, int dtype,
                                    void* stream) {
  if (int e = multi_check(rows, C, dtype)) return e;
  if (rows == 0) return 0;
  R3D_CHECK(g_own && g_prev && dx && (k == 0 || (idx_own && idx_prev)), "null pointer");
  R3D_CHECK(((uintptr_t(g_own) | uintptr_t(g_prev) | uintptr_t(dx)) & 15) == 0 && g_pitch % 8 == 0, "unaligned tensors");
  cudaStream_t st = (cudaStream_t)stream;
  R3D_STAGE(ST_EXCHANGE_BWD, st);
  if (dtype == R3D_F32)
    exchange_one_bwd_kernel<float, 4><<<ew_grid(rows * C / 4), 256, size_t(C), st>>>(
        (const float*)g_own, (const float*)g_prev, g_pitch, idx_own, idx_prev, k, (float*)dx, rows, C);
  else
    exchange_one_bwd_kernel<__nv_bfloat16, 8><<<ew_grid(rows * C / 8), 256, size_t(C), st>>>(
        (const __nv_bfloat16*)g_own, (const __nv_bfloat16*)g_prev, g_pitch, idx_own, idx_prev, k, (__nv_bfloat16*)dx, rows, C);
  R3D_LAUNCH_CHECK();
  return 0;
}

static int attn_check(int64_t rows, int M, int64_t C, int heads, int dtype) {
  R3D_CHECK(rows >= 0 && M >= 2 && M <= MT_MAX, "M-token attention supports 2 <= M <= %d modalities", MT_MAX);
  R3D_CHECK(heads >= 1 && C % heads == 0 && (C / heads) % 32 == 0 && C / heads <= 32 * HD_MAX,
            "head_dim = C / heads must be a multiple of 32 and <= %d", 32 * HD_MAX);
  R3D_CHECK(dtype == R3D_F32 || dtype == R3D_BF16, "bad dtype %d", dtype);
  return 0;
}

// qkv (rows, M, 3C) as nn.Linear(C, 3C) lays it out (model/extras/transformerblock.py:22-23) -> out (rows, M, C):
// softmax over the OTHER tokens (diagonal masked with -inf, tokenfusion.py:68-72), scale = head_dim^-0.5.
extern "C" int r3d_mtoken_attn_fwd(const void* qkv, void* out, int64_t rows, int M, int64_t C, int heads, int dtype,
                                   void* stream) {
  if (int e = attn_check(rows, M, C, heads, dtype)) return e;
  if (rows == 0) return 0;
  R3D_CHECK(qkv && out, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const float scale = 1.f / sqrtf(float(C / heads));
  const int grid = (int)std::min<int64_t>((rows * heads + 7) / 8, int64_t(kNumSMs) * 16);
  R3D_STAGE(ST_BLOCK, st);
  if (dtype == R3D_F32)
    mtoken_attn_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)qkv, (float*)out, rows, M, (int)C, heads, scale);
  else
    mtoken_attn_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, rows, M,
                                                               (int)C, heads, scale);
  R3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int r3d_mtoken_attn_bwd(const void* qkv, const void* dout, void* dqkv, int64_t rows, int M, int64_t C,
                                   int heads, int dtype, void* stream) {
  if (int e = attn_check(rows, M, C, heads, dtype)) return e;
  if (rows == 0) return 0;
  R3D_CHECK(qkv && dout && dqkv, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const float scale = 1.f / sqrtf(float(C / heads));
  const int grid = (int)std::min<int64_t>((rows * heads + 7) / 8, int64_t(kNumSMs) * 16);
  R3D_STAGE(ST_BLOCK, st);
  if (dtype == R3D_F32)
    mtoken_attn_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)qkv, (const float*)dout, (float*)dqkv, rows, M, (int)C,
                                                       heads, scale);
  else
    mtoken_attn_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dout,
                                                               (__nv_bfloat16*)dqkv, rows, M, (int)C, heads, scale);
  R3D_LAUNCH_CHECK();
  return 0;
}

// (rows, M, C) -> (rows, C): mean over the M modality tokens (tokenfusion.py:95), and its backward.
extern "C" int r3d_token_mean(const void* x, void* out, int64_t rows, int M, int64_t C, int dtype, int backward,
                              void* stream) {
  if (int e = multi_check(rows, C, dtype)) return e;
  R3D_CHECK(M >= 1, "bad M");
  if (rows == 0) return 0;
  R3D_CHECK(x && out, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  R3D_STAGE(ST_BLOCK, st);
  if (dtype == R3D_F32) {
    if (backward) token_mean_bwd_kernel<float, 4><<<ew_grid(rows * C / 4), 256, 0, st>>>((const float*)x, (float*)out, rows, M, C);
    else token_mean_fwd_kernel<float, 4><<<ew_grid(rows * C / 4), 256, 0, st>>>((const float*)x, (float*)out, rows, M, C);
  } else {
    using Bf = __nv_bfloat16;
    if (backward) token_mean_bwd_kernel<Bf, 8><<<ew_grid(rows * C / 8), 256, 0, st>>>((const Bf*)x, (Bf*)out, rows, M, C);
    else token_mean_fwd_kernel<Bf, 8><<<ew_grid(rows * C / 8), 256, 0, st>>>((const Bf*)x, (Bf*)out, rows, M, C);
  }
  R3D_LAUNCH_CHECK();
  return 0;
}
