// Token-axis selection path (north_star kernels 3-6).
//
// The reference ships only the CHANNEL exchange (model/futr_safuser_tokenfusion.py:33-66, "Token Fusion by Channel
// Exchanging"); the paper prose it implements (README.md:13) describes the token form: per sample, score
// every token by how much of the sample's spectrum it carries, take the k least informative tokens of a modality and
// swap in the other modality's tokens at those positions.  No reference code exists for it -- PARITY UNPINNED; the
// oracle is oracle/fuser_oracle.py:token_fusion_tokens.  Kernels:
//
//   r3d_token_scores        s[b, t] = sum_j p_j u_{t j}^2  (u_j: left singular vectors of the (T, C) sample; p = sigma / sum sigma)
//                           from what r3d_erank_fwd saved: U when T < C (Gram on the token side), Y = U^T X rows when T >= C
//   r3d_bottomk             (fusion_kernels.cu) per-sample bottom-k of the scores, ties -> lower token index
//   r3d_token_mask          index lists (B, k) x 2 -> one byte per token: bit 0 = t in S_rgb(b), bit 1 = t in S_depth(b)
//   r3d_token_exchange_fwd  out[b, t, 0, :] = t in S_rgb(b) ? depth[b, t, :] : rgb[b, t, :],  out[b, t, 1, :] mirrored;
//                           128-bit copies, writes the stacked (B, T, 2, C) tensor directly: 2 N s read + 2 N s written
//   r3d_token_exchange_bwd  masked select per token (the index sets are duplicate-free per sample, so the "scatter-add"
//                           of north_star degenerates to a select: no atomics): 2 N s read + 2 N s written
#include "common.cuh"

namespace r3d {

// one CTA per sample; all 256 threads
__device__ __forceinline__ float tk_block_reduce(float v, bool is_max, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, w) : (v + w);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
  for (int i = 1; i < 8; ++i) r = is_max ? fmaxf(r, sh[i]) : (r + sh[i]);
  return r;
}

// w_j = p_j / (sigma_j^2 ||u_j||^2) (tside == 0: scores from Y rows, Y[j, t] = sigma_j u_{t j} ||u_j||) or
// w_j = p_j / ||u_j||^2 (tside == 1: scores from U rows directly); kept singular values only (sigma > rtol sigma_max)
__global__ void __launch_bounds__(256) token_scores_kernel(const float* __restrict__ sigma, const float* __restrict__ U,
                                                           const float* __restrict__ Y, int n, int m, int T, int tside,
                                                           float rtol, float* __restrict__ out) {
  __shared__ float sh[8];
  extern __shared__ float wj[];
  const int b = blockIdx.x;
  const float* sg = sigma + int64_t(b) * n;
  const float* u = U + int64_t(b) * n * n;
  float mx = 0.f;
  for (int j = threadIdx.x; j < n; j += 256) mx = fmaxf(mx, sg[j]);
  const float smax = tk_block_reduce(mx, true, sh);
  const float cut = rtol * smax;
  float s = 0.f;
  for (int j = threadIdx.x; j < n; j += 256) { const float v = sg[j]; if (v > cut) s += v; }
  const float S = tk_block_reduce(s, false, sh);
  // row norms of U (unit up to fp32 drift), one warp per row
  for (int j = threadIdx.x >> 5; j < n; j += 8) {
    float su = 0.f;
    for (int c = threadIdx.x & 31; c < n; c += 32) { const float v = u[int64_t(j) * n + c]; su = fmaf(v, v, su); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) su += __shfl_xor_sync(0xffffffffu, su, o);
    if ((threadIdx.x & 31) == 0) {
      const float v = sg[j];
      float w = 0.f;
      if (v > cut && S > 0.f && su > 0.f) w = tside ? (v / S) / su : (v / S) / (v * v * su);
      wj[j] = w;
    }
  }
  __syncthreads();
  const float* src = tside ? u : (Y + int64_t(b) * n * m);     // rows j, columns t (pitch n or m)
  const int pitch = tside ? n : m;
  for (int t = threadIdx.x; t < T; t += 256) {
    float acc = 0.f;
    for (int j = 0; j < n; ++j) { const float v = src[int64_t(j) * pitch + t]; acc = fmaf(wj[j] * v, v, acc); }
    out[int64_t(b) * T + t] = acc;
  }
}

__global__ void __launch_bounds__(256) token_mask_kernel(const int64_t* __restrict__ idx_r,
                                                         const int64_t* __restrict__ idx_d, int64_t k, int64_t T,
                                                         uint8_t* __restrict__ mask) {
  extern __shared__ int bits[];
  const int b = blockIdx.x;
  for (int t = threadIdx.x; t < T; t += 256) bits[t] = 0;
  __syncthreads();
  for (int64_t i = threadIdx.x; i < k; i += 256) {
    const int64_t a = idx_r[int64_t(b) * k + i], c = idx_d[int64_t(b) * k + i];
    if (a >= 0 && a < T) atomicOr(&bits[a], 1);
    if (c >= 0 && c < T) atomicOr(&bits[c], 2);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += 256) mask[int64_t(b) * T + t] = (uint8_t)bits[t];
}

// one 128-bit vector per thread per iteration; a row's vectors share one mask byte
template <typename T, int V>
__global__ void __launch_bounds__(256) token_exchange_fwd_kernel(const T* __restrict__ rgb, const T* __restrict__ depth,
                                                                 const uint8_t* __restrict__ mask, T* __restrict__ out,
                                                                 int64_t rows, int64_t C) {
  const int64_t cv = C / V, total = rows * cv;
  const int64_t stride = int64_t(gridDim.x) * 256;
  for (int64_t e0 = int64_t(blockIdx.x) * 256 + threadIdx.x; e0 < total; e0 += 4 * stride) {
    float r[4][V], d[4][V];
    int64_t row[4], col[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t e = e0 + u * stride;
      ok[u] = e < total;
      row[u] = ok[u] ? e / cv : 0;
      col[u] = ok[u] ? (e - row[u] * cv) * V : 0;
      if (ok[u]) {
        load_vec<T, V>(rgb + row[u] * C + col[u], r[u]);
        load_vec<T, V>(depth + row[u] * C + col[u], d[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
      const int mk = mask[row[u]];
      T* o = out + row[u] * 2 * C + col[u];
      if (mk & 1) store_vec<T, V>(o, d[u]); else store_vec<T, V>(o, r[u]);
      if (mk & 2) store_vec<T, V>(o + C, r[u]); else store_vec<T, V>(o + C, d[u]);
    }
  }
}

// d_rgb[row] = (t in S_r ? 0 : g_r) + (t in S_d ? g_d : 0);  d_depth[row] = (t in S_d ? 0 : g_d) + (t in S_r ? g_r : 0)
template <typename T, int V>
__global__ void __launch_bounds__(256) token_exchange_bwd_kernel(const T* __restrict__ g, const uint8_t* __restrict__ mask,
                                                                 T* __restrict__ d_rgb, T* __restrict__ d_depth,
                                                                 int64_t rows, int64_t C) {
  const int64_t cv = C / V, total = rows * cv;
  const int64_t stride = int64_t(gridDim.x) * 256;
  for (int64_t e0 = int64_t(blockIdx.x) * 256 + threadIdx.x; e0 < total; e0 += 4 * stride) {
    float gr[4][V], gd[4][V];
    int64_t row[4], col[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t e = e0 + u * stride;
      ok[u] = e < total;
      row[u] = ok[u] ? e / cv : 0;
      col[u] = ok[u] ? (e - row[u] * cv) * V : 0;
      if (ok[u]) {
        load_vec<T, V>(g + row[u] * 2 * C + col[u], gr[u]);
        load_vec<T, V>(g + row[u] * 2 * C + C + col[u], gd[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
      const int mk = mask[row[u]];
      float o_r[V], o_d[V];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        // same association as autograd: the copy-through term first, then the swapped-in term
        o_r[i] = ((mk & 1) ? 0.f : gr[u][i]) + ((mk & 2) ? gd[u][i] : 0.f);
        o_d[i] = ((mk & 2) ? 0.f : gd[u][i]) + ((mk & 1) ? gr[u][i] : 0.f);
      }
      store_vec<T, V>(d_rgb + row[u] * C + col[u], o_r);
      store_vec<T, V>(d_depth + row[u] * C + col[u], o_d);
    }
  }
}

template <typename T, int V>
static int token_fwd_launch(const void* rgb, const void* depth, const uint8_t* mask, void* out, int64_t rows, int64_t C,
                            cudaStream_t st) {
  const int64_t total = rows * (C / V);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((total + 1023) / 1024, int64_t(kNumSMs) * 8));
  R3D_STAGE(ST_EXCHANGE_FWD, st);
  token_exchange_fwd_kernel<T, V><<<grid, 256, 0, st>>>((const T*)rgb, (const T*)depth, mask, (T*)out, rows, C);
  R3D_LAUNCH_CHECK();
  return 0;
}
template <typename T, int V>
static int token_bwd_launch(const void* g, const uint8_t* mask, void* d_rgb, void* d_depth, int64_t rows, int64_t C,
                            cudaStream_t st) {
  const int64_t total = rows * (C / V);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((total + 1023) / 1024, int64_t(kNumSMs) * 8));
  R3D_STAGE(ST_EXCHANGE_BWD, st);
  token_exchange_bwd_kernel<T, V><<<grid, 256, 0, st>>>((const T*)g, mask, (T*)d_rgb, (T*)d_depth, rows, C);
  R3D_LAUNCH_CHECK();
  return 0;
}

}  // namespace r3d

using namespace r3d;

extern "C" int r3d_token_scores(const float* sigma, const float* U, const float* Y, int64_t B, int64_t T, int64_t C,
                                float rtol, float* score_out, void* stream) {
  R3D_CHECK(sigma && U && Y && score_out, "null pointer");
  R3D_CHECK(B >= 1 && T >= 1 && C >= 1, "bad shape");
  const bool tside = T < C;                       // same side rule as r3d_erank_fwd
  const int64_t n = tside ? T : C, m = tside ? C : T;
  R3D_CHECK(n <= 8192, "min(T, C) exceeds 8192");
  R3D_STAGE(ST_TOKEN_INFO, (cudaStream_t)stream);
  token_scores_kernel<<<(unsigned)B, 256, size_t(n) * 4, (cudaStream_t)stream>>>(sigma, U, Y, int(n), int(m), int(T),
                                                                                tside ? 1 : 0, rtol, score_out);
  R3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int r3d_token_mask(const int64_t* idx_r, const int64_t* idx_d, int64_t B, int64_t T, int64_t k,
                              uint8_t* mask_out, void* stream) {
  R3D_CHECK(mask_out != nullptr, "null pointer");
  R3D_CHECK(B >= 0 && T >= 1 && T <= 8192 && k >= 0 && k <= T, "bad shape B=%lld T=%lld k=%lld", (long long)B,
            (long long)T, (long long)k);
  R3D_CHECK(k == 0 || (idx_r && idx_d), "null index pointer with k > 0");
  if (B == 0) return 0;
  token_mask_kernel<<<(unsigned)B, 256, size_t(T) * 4, (cudaStream_t)stream>>>(idx_r, idx_d, k, T, mask_out);
  R3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int r3d_token_exchange_fwd(const void* rgb, const void* depth, const uint8_t* mask, void* out, int64_t rows,
                                      int64_t C, int dtype, void* stream) {
  R3D_CHECK(rows >= 0 && C >= 1, "bad shape");
  R3D_CHECK(dtype == R3D_F32 || dtype == R3D_BF16, "bad dtype %d", dtype);
  if (rows == 0) return 0;
  R3D_CHECK(rgb && depth && mask && out, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == R3D_F32) {
    const bool vec = vec_ok<float>(rgb, C) && vec_ok<float>(depth, C) && vec_ok<float>(out, C);
    return vec ? token_fwd_launch<float, 4>(rgb, depth, mask, out, rows, C, st)
               : token_fwd_launch<float, 1>(rgb, depth, mask, out, rows, C, st);
  }
  using Bf = __nv_bfloat16;
  const bool vec = vec_ok<Bf>(rgb, C) && vec_ok<Bf>(depth, C) && vec_ok<Bf>(out, C);
  return vec ? token_fwd_launch<Bf, 8>(rgb, depth, mask, out, rows, C, st)
             : token_fwd_launch<Bf, 1>(rgb, depth, mask, out, rows, C, st);
}

extern "C" int r3d_token_exchange_bwd(const void* g, const uint8_t* mask, void* d_rgb, void* d_depth, int64_t rows,
                                      int64_t C, int dtype, void* stream) {
  R3D_CHECK(rows >= 0 && C >= 1, "bad shape");
  R3D_CHECK(dtype == R3D_F32 || dtype == R3D_BF16, "bad dtype %d", dtype);
  if (rows == 0) return 0;
  R3D_CHECK(g && mask && d_rgb && d_depth, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == R3D_F32) {
    const bool vec = vec_ok<float>(g, C) && vec_ok<float>(d_rgb, C) && vec_ok<float>(d_depth, C);
    return vec ? token_bwd_launch<float, 4>(g, mask, d_rgb, d_depth, rows, C, st)
               : token_bwd_launch<float, 1>(g, mask, d_rgb, d_depth, rows, C, st);
  }
  using Bf = __nv_bfloat16;
  const bool vec = vec_ok<Bf>(g, C) && vec_ok<Bf>(d_rgb, C) && vec_ok<Bf>(d_depth, C);
  return vec ? token_bwd_launch<Bf, 8>(g, mask, d_rgb, d_depth, rows, C, st)
             : token_bwd_launch<Bf, 1>(g, mask, d_rgb, d_depth, rows, C, st);
}
