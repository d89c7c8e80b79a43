// Row LayerNorm forward / backward for the fuser Block (first step): the reference runs three
// nn.LayerNorm over (B*T*2, C) rows per fuser call -- norm1 / norm2 of the Block
// (model/extras/transformerblock.py:122,127,132,134) and the final norm (model/futr_safuser_tokenfusion.py:25,93).
// At the headline shape (131072 rows x 512, bf16) ATen's LayerNorm backward spends 1.66 ms per fuser step in the
// gamma/beta reduction alone (39 % of the whole fuser forward+backward).  Here:
//   forward : one warp per row, the row lives in registers (128-bit loads), two-pass mean / variance in fp32, one
//             read + one write of the tensor; saves mean and rstd (fp32 per row);
//   backward: same mapping; dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma, and every lane keeps
//             fp32 accumulators of dy*xhat and dy for ITS columns over all rows of the persistent CTA, reduced across
//             the CTA's warps through shared memory into one partial row per CTA, then a fixed-order finalize --
//             deterministic, no atomics, dy and x are read exactly once.
// HBM bound: forward 2 * rows * C * s bytes, backward 3 * rows * C * s.
#include "common.cuh"

namespace r3d {

namespace {

constexpr int LN_WARPS = 8;

// PAIR: the rows come in pairs (the fuser's two modality tokens per (b, t)); y has rows/2 rows and receives the mean
// of the two normalised tokens (model/futr_safuser_tokenfusion.py:93-95: norm, then mean over the token dimension).
// RELU (depth projection, model/futr_safuser_tokenfusion.py:195-197: LayerNorm then F.relu): y = max(LN(x), 0) and the
// per-CTA column sums of |y| (rounded values) go to colsum_partial[blockIdx.x][C] -- the channel-score partials of
// tokenfusion.py:49-50 without another pass over the tensor.
template <typename T, int V, int NCH, bool PAIR, bool RELU = false>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ gamma,
                                                               const T* __restrict__ beta, int64_t rows, int C,
                                                               float eps, T* __restrict__ y,
                                                               float* __restrict__ mean_out,
                                                               float* __restrict__ rstd_out, int swap,
                                                               float* __restrict__ colsum_partial = nullptr) {
  __shared__ float red_f[RELU ? LN_WARPS : 1][RELU ? 32 * V : 1];
  float cs[RELU ? NCH : 1][RELU ? V : 1];
  if (RELU) {
#pragma unroll
    for (int k = 0; k < NCH; ++k)
#pragma unroll
      for (int i = 0; i < V; ++i) cs[k][i] = 0.f;
  }
  // swap (not with PAIR): the normalised row r is WRITTEN to row r ^ 1 -- the closed-form 2-token attention hands
  // token m the V of token 1 - m, and doing the swap here keeps every GEMM of the Block plain
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float gm[NCH][V], bt[NCH][V];
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int col = (lane + 32 * k) * V;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      gm[k][i] = (col + i < C) ? float(gamma[col + i]) : 0.f;
      bt[k][i] = (col + i < C) ? float(beta[col + i]) : 0.f;
    }
  }
  const float invC = 1.f / float(C);
  const int64_t units = PAIR ? rows / 2 : rows;          // a warp handles one row, or one pair of rows
  for (int64_t unit = int64_t(blockIdx.x) * LN_WARPS + warp; unit < units; unit += int64_t(gridDim.x) * LN_WARPS) {
    float acc[NCH][V];
#pragma unroll
    for (int h = 0; h < (PAIR ? 2 : 1); ++h) {
      const int64_t row = PAIR ? 2 * unit + h : unit;
      float xv[NCH][V];
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        const int col = (lane + 32 * k) * V;
        if (col < C) load_vec<T, V>(x + row * C + col, xv[k]);
        else {
#pragma unroll
          for (int i = 0; i < V; ++i) xv[k][i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < V; ++i) s += xv[k][i];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * invC;
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        const int col = (lane + 32 * k) * V;
        if (col < C) {
#pragma unroll
          for (int i = 0; i < V; ++i) { const float d = xv[k][i] - mean; q = fmaf(d, d, q); }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = rsqrtf(q * invC + eps);
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float o = fmaf((xv[k][i] - mean) * rstd, gm[k][i], bt[k][i]);
          if (RELU) {
            o = fmaxf(o, 0.f);
            cs[k][i] += float(T(o));                     // the stored (rounded) value, |o| = o after the ReLU
          }
          acc[k][i] = (PAIR && h == 1) ? 0.5f * (acc[k][i] + o) : o;
        }
      }
      if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
    }
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int col = (lane + 32 * k) * V;
      if (col < C) store_vec<T, V>(y + (unit ^ int64_t(swap)) * C + col, acc[k]);
    }
  }
  if (RELU && colsum_partial != nullptr) {
    // one partial row per CTA, warps added in index order (deterministic)
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
#pragma unroll
      for (int i = 0; i < V; ++i) red_f[warp][lane * V + i] = cs[k][i];
      __syncthreads();
      for (int e = threadIdx.x; e < 32 * V; e += LN_WARPS * 32) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < LN_WARPS; ++w) t += red_f[w][e];
        const int col = 32 * k * V + e;
        if (col < C) colsum_partial[int64_t(blockIdx.x) * C + col] = t;
      }
      __syncthreads();
    }
  }
}

template <typename T, int V, int NCH, bool PAIR>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                               const float* __restrict__ mean_in,
                                                               const float* __restrict__ rstd_in,
                                                               const T* __restrict__ gamma, int64_t rows, int C,
                                                               T* __restrict__ dx, float* __restrict__ partial,
                                                               int swap, const T* __restrict__ addend,
                                                               const T* __restrict__ relu_beta = nullptr) {
  // relu_beta != nullptr: the forward applied ReLU after the affine; dy is masked where gamma * xhat + beta <= 0
  // swap: dy is READ from row r ^ 1 (backward of the swapped write above); addend: dx[r] += addend[r] (the gradient
  // arriving over the residual connection, fused instead of a separate elementwise add)
  __shared__ float red[LN_WARPS][32 * V];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float gm[NCH][V], ag[NCH][V], ab[NCH][V];
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int col = (lane + 32 * k) * V;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      gm[k][i] = (col + i < C) ? float(gamma[col + i]) : 0.f;
      ag[k][i] = 0.f; ab[k][i] = 0.f;
    }
  }
  const float invC = 1.f / float(C);
  for (int64_t row = int64_t(blockIdx.x) * LN_WARPS + warp; row < rows; row += int64_t(gridDim.x) * LN_WARPS) {
    const float mean = mean_in[row], rstd = rstd_in[row];
    float xh[NCH][V], g[NCH][V];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int col = (lane + 32 * k) * V;
      if (col < C) {
        float xv[V], dv[V];
        load_vec<T, V>(x + row * C + col, xv);
        load_vec<T, V>(dy + (PAIR ? (row >> 1) : (row ^ int64_t(swap))) * C + col, dv);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          if (PAIR) dv[i] *= 0.5f;
          xh[k][i] = (xv[i] - mean) * rstd;
          if (relu_beta != nullptr && fmaf(xh[k][i], gm[k][i], float(relu_beta[col + i])) <= 0.f) dv[i] = 0.f;
          g[k][i] = dv[i] * gm[k][i];
          s1 += g[k][i];
          s2 = fmaf(g[k][i], xh[k][i], s2);
          ag[k][i] = fmaf(dv[i], xh[k][i], ag[k][i]);
          ab[k][i] += dv[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) { xh[k][i] = 0.f; g[k][i] = 0.f; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float m1 = s1 * invC, m2 = s2 * invC;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int col = (lane + 32 * k) * V;
      if (col < C) {
        float o[V];
#pragma unroll
        for (int i = 0; i < V; ++i) o[i] = rstd * (g[k][i] - m1 - xh[k][i] * m2);
        if (addend != nullptr) {
          float av[V];
          load_vec<T, V>(addend + row * C + col, av);
#pragma unroll
          for (int i = 0; i < V; ++i) o[i] += av[i];
        }
        store_vec<T, V>(dx + row * C + col, o);
      }
    }
  }
  // one partial row per CTA: [2][gridDim.x][C] (dgamma partials, then dbeta partials), warps added in index order
  float* pg = partial + int64_t(blockIdx.x) * C;
  float* pb = partial + (int64_t(gridDim.x) + blockIdx.x) * C;
#pragma unroll
  for (int which = 0; which < 2; ++which) {
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
#pragma unroll
      for (int i = 0; i < V; ++i) red[warp][lane * V + i] = which ? ab[k][i] : ag[k][i];
      __syncthreads();
      for (int e = threadIdx.x; e < 32 * V; e += LN_WARPS * 32) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < LN_WARPS; ++w) t += red[w][e];
        const int col = 32 * k * V + e;       // chunk k covers columns [32*k*V, 32*(k+1)*V)
        if (col < C) (which ? pb : pg)[col] = t;
      }
      __syncthreads();
    }
  }
}

// out[which][c] = sum over parts of partial[which][part][c], fixed order (32 channels x 8 lanes per CTA)
__global__ void __launch_bounds__(256) ln_finalize_kernel(const float* __restrict__ partial, int parts, int C,
                                                          float* __restrict__ out) {
  __shared__ float red[8][33];
  const int cl = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float s = 0.f;
  if (c < C) {
    const float* p = partial + int64_t(blockIdx.y) * parts * C + c;
    const int per = (parts + 7) / 8;
    const int i0 = g * per, i1 = min(parts, i0 + per);
    for (int i = i0; i < i1; ++i) s += p[int64_t(i) * C];
  }
  red[g][cl] = s;
  __syncthreads();
  if (g == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][cl];
    out[int64_t(blockIdx.y) * C + c] = t;
  }
}

// out[row] = (a ? a[row] : 0) + b[row ^ 1]: the closed-form 2-token attention hands token m the projected V of token
// 1-m, so the residual add reads its second operand with the two rows of a pair swapped; with a == null
// it is the backward of that operand (a pair-swapped copy).
template <typename T, int V>
__global__ void __launch_bounds__(256) swap_add_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                       T* __restrict__ out, int64_t rows, int C) {
  const int vec_per_row = C / V;
  const int64_t total = rows * vec_per_row;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
    const int64_t row = e / vec_per_row;
    const int col = int(e % vec_per_row) * V;
    float bv[V], o[V];
    load_vec<T, V>(b + (row ^ 1) * C + col, bv);
    if (a != nullptr) {
      float av[V];
      load_vec<T, V>(a + row * C + col, av);
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = av[i] + bv[i];
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = bv[i];
    }
    store_vec<T, V>(out + row * C + col, o);
  }
}

inline int ln_grid(int64_t rows) {
  const int64_t want = (rows + LN_WARPS - 1) / LN_WARPS;
  return int(std::max<int64_t>(1, std::min<int64_t>(want, int64_t(kNumSMs) * 2)));
}

template <typename T, int V, bool PAIR>
int ln_fwd_t(const void* x, const void* gamma, const void* beta, int64_t rows, int C, float eps, void* y, float* mean,
             float* rstd, cudaStream_t st, int swap = 0) {
  const int nch = (C + 32 * V - 1) / (32 * V);
  const int grid = ln_grid(PAIR ? rows / 2 : rows);
#define R3D_LN_FWD(N)                                                                                         \
  ln_fwd_kernel<T, V, N, PAIR><<<grid, LN_WARPS * 32, 0, st>>>((const T*)x, (const T*)gamma, (const T*)beta, rows, C, \
                                                               eps, (T*)y, mean, rstd, swap)
  if (nch <= 1) R3D_LN_FWD(1);
  else if (nch <= 2) R3D_LN_FWD(2);
  else if (nch <= 4) R3D_LN_FWD(4);
  else R3D_LN_FWD(8);
#undef R3D_LN_FWD
  R3D_LAUNCH_CHECK();
  return 0;
}

template <typename T, int V>
int ln_fwd_relu_t(const void* x, const void* gamma, const void* beta, int64_t rows, int C, float eps, void* y, float* mean,
                  float* rstd, float* colsum_partial, cudaStream_t st) {
  const int nch = (C + 32 * V - 1) / (32 * V);
  const int grid = ln_grid(rows);
#define R3D_LN_FWDR(N)                                                                                              \
  ln_fwd_kernel<T, V, N, false, true><<<grid, LN_WARPS * 32, 0, st>>>((const T*)x, (const T*)gamma, (const T*)beta, rows, \
                                                                      C, eps, (T*)y, mean, rstd, 0, colsum_partial)
  if (nch <= 1) R3D_LN_FWDR(1);
  else if (nch <= 2) R3D_LN_FWDR(2);
  else if (nch <= 4) R3D_LN_FWDR(4);
  else R3D_LN_FWDR(8);
#undef R3D_LN_FWDR
  R3D_LAUNCH_CHECK();
  return 0;
}

template <typename T, int V, bool PAIR>
int ln_bwd_t(const void* dy, const void* x, const float* mean, const float* rstd, const void* gamma, int64_t rows, int C,
             void* dx, float* partial, cudaStream_t st, int swap = 0, const void* addend = nullptr,
             const void* relu_beta = nullptr) {
  const int nch = (C + 32 * V - 1) / (32 * V);
  const int grid = ln_grid(rows);
#define R3D_LN_BWD(N)                                                                                             \
  ln_bwd_kernel<T, V, N, PAIR><<<grid, LN_WARPS * 32, 0, st>>>((const T*)dy, (const T*)x, mean, rstd, (const T*)gamma, \
                                                               rows, C, (T*)dx, partial, swap, (const T*)addend, \
                                                               (const T*)relu_beta)
  if (nch <= 1) R3D_LN_BWD(1);
  else if (nch <= 2) R3D_LN_BWD(2);
  else if (nch <= 4) R3D_LN_BWD(4);
  else R3D_LN_BWD(8);
#undef R3D_LN_BWD
  R3D_LAUNCH_CHECK();
  return 0;
}

int ln_check(const void* a, const void* b, int64_t rows, int64_t C, int dtype) {
  R3D_CHECK(a && b, "null pointer");
  R3D_CHECK(rows >= 0 && C >= 1, "bad shape");
  R3D_CHECK(dtype == R3D_F32 || dtype == R3D_BF16, "bad dtype %d", dtype);
  const int V = dtype == R3D_F32 ? 4 : 8;
  R3D_CHECK(C % V == 0, "LayerNorm width C=%lld must be a multiple of %d for this dtype", (long long)C, V);
  R3D_CHECK(C <= 32 * V * 8, "LayerNorm width C=%lld exceeds %d", (long long)C, 32 * V * 8);
  R3D_CHECK((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0,
            "LayerNorm tensors must be 16-byte aligned");
  return 0;
}

}  // namespace

}  // namespace r3d

using namespace r3d;

extern "C" size_t r3d_ln_bwd_workspace_floats(int64_t rows, int64_t C) {
  return size_t(2) * size_t(ln_grid(rows)) * size_t(C);
}

extern "C" int r3d_ln_fwd(const void* x, const void* gamma, const void* beta, int64_t rows, int64_t C, int dtype,
                          float eps, int pair_mean, void* y, float* mean, float* rstd, void* stream) {
  if (int e = ln_check(x, y, rows, C, dtype)) return e;
  R3D_CHECK(gamma && beta && mean && rstd, "null pointer");
  R3D_CHECK(!pair_mean || rows % 2 == 0, "pair_mean needs an even number of rows");
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  R3D_STAGE(ST_BLOCK, st);
  if (pair_mean)
    return dtype == R3D_F32 ? ln_fwd_t<float, 4, true>(x, gamma, beta, rows, int(C), eps, y, mean, rstd, st)
                            : ln_fwd_t<__nv_bfloat16, 8, true>(x, gamma, beta, rows, int(C), eps, y, mean, rstd, st);
  return dtype == R3D_F32 ? ln_fwd_t<float, 4, false>(x, gamma, beta, rows, int(C), eps, y, mean, rstd, st)
                          : ln_fwd_t<__nv_bfloat16, 8, false>(x, gamma, beta, rows, int(C), eps, y, mean, rstd, st);
}

extern "C" int r3d_ln_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const void* gamma,
                          int64_t rows, int64_t C, int dtype, int pair_mean, void* dx, float* workspace,
                          float* dgamma_dbeta, void* stream) {
  if (int e = ln_check(dy, x, rows, C, dtype)) return e;
  R3D_CHECK(mean && rstd && gamma && dx && workspace && dgamma_dbeta, "null pointer");
  R3D_CHECK((reinterpret_cast<uintptr_t>(dx) & 15) == 0, "dx must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) {
    R3D_CUDA(cudaMemsetAsync(dgamma_dbeta, 0, size_t(2) * C * sizeof(float), st));
    return 0;
  }
  R3D_STAGE(ST_BLOCK, st);
  int e;
  if (pair_mean)
    e = dtype == R3D_F32 ? ln_bwd_t<float, 4, true>(dy, x, mean, rstd, gamma, rows, int(C), dx, workspace, st)
                         : ln_bwd_t<__nv_bfloat16, 8, true>(dy, x, mean, rstd, gamma, rows, int(C), dx, workspace, st);
  else
    e = dtype == R3D_F32 ? ln_bwd_t<float, 4, false>(dy, x, mean, rstd, gamma, rows, int(C), dx, workspace, st)
                         : ln_bwd_t<__nv_bfloat16, 8, false>(dy, x, mean, rstd, gamma, rows, int(C), dx, workspace, st);
  if (e) return e;
  ln_finalize_kernel<<<dim3((unsigned)((C + 31) / 32), 2), 256, 0, st>>>(workspace, ln_grid(rows), int(C), dgamma_dbeta);
  R3D_LAUNCH_CHECK();
  return 0;
}

// flags: bit 0 = pair_mean, bit 1 = swap the two rows of every pair (forward: on the write; backward: on the read of
// dy).  addend (backward only, may be NULL): dx += addend.
extern "C" int r3d_ln_fwd2(const void* x, const void* gamma, const void* beta, int64_t rows, int64_t C, int dtype,
                           float eps, int flags, void* y, float* mean, float* rstd, void* stream) {
  if ((flags & 2) == 0) return r3d_ln_fwd(x, gamma, beta, rows, C, dtype, eps, flags & 1, y, mean, rstd, stream);
  R3D_CHECK((flags & 1) == 0, "pair_mean and swap cannot be combined");
  if (int e = ln_check(x, y, rows, C, dtype)) return e;
  R3D_CHECK(gamma && beta && mean && rstd, "null pointer");
  R3D_CHECK(rows % 2 == 0, "swap needs an even number of rows");
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  R3D_STAGE(ST_BLOCK, st);
  return dtype == R3D_F32 ? ln_fwd_t<float, 4, false>(x, gamma, beta, rows, int(C), eps, y, mean, rstd, st, 1)
                          : ln_fwd_t<__nv_bfloat16, 8, false>(x, gamma, beta, rows, int(C), eps, y, mean, rstd, st, 1);
}

extern "C" int r3d_ln_bwd2(const void* dy, const void* x, const float* mean, const float* rstd, const void* gamma,
                           int64_t rows, int64_t C, int dtype, int flags, const void* addend, void* dx,
                           float* workspace, float* dgamma_dbeta, void* stream) {
  if ((flags & 2) == 0 && addend == nullptr)
    return r3d_ln_bwd(dy, x, mean, rstd, gamma, rows, C, dtype, flags & 1, dx, workspace, dgamma_dbeta, stream);
  R3D_CHECK((flags & 1) == 0, "pair_mean cannot be combined with swap / addend");
  if (int e = ln_check(dy, x, rows, C, dtype)) return e;
  R3D_CHECK(mean && rstd && gamma && dx && workspace && dgamma_dbeta, "null pointer");
  R3D_CHECK((reinterpret_cast<uintptr_t>(dx) & 15) == 0 && (reinterpret_cast<uintptr_t>(addend) & 15) == 0,
            "dx / addend must be 16-byte aligned");
  R3D_CHECK((flags & 2) == 0 || rows % 2 == 0, "swap needs an even number of rows");
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) {
    R3D_CUDA(cudaMemsetAsync(dgamma_dbeta, 0, size_t(2) * C * sizeof(float), st));
    return 0;
  }
  R3D_STAGE(ST_BLOCK, st);
  const int sw = (flags & 2) ? 1 : 0;
  int e = dtype == R3D_F32
              ? ln_bwd_t<float, 4, false>(dy, x, mean, rstd, gamma, rows, int(C), dx, workspace, st, sw, addend)
              : ln_bwd_t<__nv_bfloat16, 8, false>(dy, x, mean, rstd, gamma, rows, int(C), dx, workspace, st, sw, addend);
  if (e) return e;
  ln_finalize_kernel<<<dim3((unsigned)((C + 31) / 32), 2), 256, 0, st>>>(workspace, ln_grid(rows), int(C), dgamma_dbeta);
  R3D_LAUNCH_CHECK();
  return 0;
}

// Depth projection tail (model/futr_safuser_tokenfusion.py:195-197): y = relu(LayerNorm(x)), plus the channel-score
// partial sums of |y|: colsum_partial holds r3d_ln_bwd_workspace_floats(rows, C) / 2 floats = one row per CTA, finalised
// by r3d_colsum_finalize(partial, r3d_ln_relu_parts(rows), C, ...).
extern "C" int64_t r3d_ln_relu_parts(int64_t rows) { return ln_grid(rows); }
extern "C" int r3d_ln_relu_fwd(const void* x, const void* gamma, const void* beta, int64_t rows, int64_t C, int dtype,
                               float eps, void* y, float* mean, float* rstd, float* colsum_partial, void* stream) {
  if (int e = ln_check(x, y, rows, C, dtype)) return e;
  R3D_CHECK(gamma && beta && mean && rstd && colsum_partial, "null pointer");
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  R3D_STAGE(ST_BLOCK, st);
  return dtype == R3D_F32 ? ln_fwd_relu_t<float, 4>(x, gamma, beta, rows, int(C), eps, y, mean, rstd, colsum_partial, st)
                          : ln_fwd_relu_t<__nv_bfloat16, 8>(x, gamma, beta, rows, int(C), eps, y, mean, rstd,
                                                            colsum_partial, st);
}
// backward of the above: dy is masked by the ReLU (recomputed from x, mean, rstd, gamma, beta), then LayerNorm backward
extern "C" int r3d_ln_relu_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const void* gamma,
                               const void* beta, int64_t rows, int64_t C, int dtype, void* dx, float* workspace,
                               float* dgamma_dbeta, void* stream) {
  if (int e = ln_check(dy, x, rows, C, dtype)) return e;
  R3D_CHECK(mean && rstd && gamma && beta && dx && workspace && dgamma_dbeta, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) {
    R3D_CUDA(cudaMemsetAsync(dgamma_dbeta, 0, size_t(2) * C * sizeof(float), st));
    return 0;
  }
  R3D_STAGE(ST_BLOCK, st);
  int e = dtype == R3D_F32
              ? ln_bwd_t<float, 4, false>(dy, x, mean, rstd, gamma, rows, int(C), dx, workspace, st, 0, nullptr, beta)
              : ln_bwd_t<__nv_bfloat16, 8, false>(dy, x, mean, rstd, gamma, rows, int(C), dx, workspace, st, 0, nullptr, beta);
  if (e) return e;
  ln_finalize_kernel<<<dim3((unsigned)((C + 31) / 32), 2), 256, 0, st>>>(workspace, ln_grid(rows), int(C), dgamma_dbeta);
  R3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int r3d_swap_add(const void* a, const void* b, int64_t rows, int64_t C, int dtype, void* out, void* stream) {
  R3D_CHECK(b && out, "null pointer");
  R3D_CHECK(rows >= 0 && rows % 2 == 0 && C >= 1, "bad shape (rows must be even)");
  R3D_CHECK(dtype == R3D_F32 || dtype == R3D_BF16, "bad dtype %d", dtype);
  const int V = dtype == R3D_F32 ? 4 : 8;
  R3D_CHECK(C % V == 0, "C=%lld must be a multiple of %d for this dtype", (long long)C, V);
  R3D_CHECK((reinterpret_cast<uintptr_t>(b) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(a) & 15) == 0, "tensors must be 16-byte aligned");
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = rows * (C / V);
  const int grid = int(std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, int64_t(kNumSMs) * 16)));
  R3D_STAGE(ST_BLOCK, st);
  if (dtype == R3D_F32)
    swap_add_kernel<float, 4><<<grid, 256, 0, st>>>((const float*)a, (const float*)b, (float*)out, rows, int(C));
  else
    swap_add_kernel<__nv_bfloat16, 8><<<grid, 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b,
                                                            (__nv_bfloat16*)out, rows, int(C));
  R3D_LAUNCH_CHECK();
  return 0;
}
