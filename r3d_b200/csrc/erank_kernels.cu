// Effective-rank chain for sm_100a (no reference symbol exists,
// the definition is SURVEY.md appendix B):
//
//   Gram on the smaller side  ->  block two-sided Jacobi eigensolver (fp32, shared
//   memory inner solves + GEMM-shaped tile updates)  ->  Rayleigh refinement
//   Y = U^T A, sigma_j = ||y_j|| / ||u_j||  ->  cut-off / normalise / entropy / exp.
//   Backward: dX = U diag(coef) Y  (one GEMM, eigenvector outer products).
//
// Side choice: T < C -> G = X X^T (n = T);  T >= C -> G = X^T X (n = C).  The
// channel side is preferred on ties because per-channel scale differences make
// X^T X a graded matrix, which two-sided Jacobi resolves to relative accuracy.
// No transposes are materialised: the GEMMs read X through (row, col) strides.
#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "jacobi_tc.cuh"
#include "pgemm.cuh"

namespace r3d {

extern int g_panel_debug;
extern int g_row_chunk_mult;
extern int g_panel_grid_cap;
static thread_local Options g_opts;
Options& options() { return g_opts; }

constexpr int JB = 32;        // Jacobi block width
constexpr int JM = 2 * JB;    // inner problem size (one block pair)
constexpr int JMAX_SWEEPS = 32;

// ------------------------------------------------------------------------------
// Generic batched SIMT GEMM, fp32 accumulate:  C[b] = op(A[b]) * op(B[b])
//   a(i,k) = TRANS_A ? A[k*lda + i] : A[i*lda + k]
//   b(k,j) = TRANS_B ? B[j*ldb + k] : B[k*ldb + j]
// optional kscale[b][k] multiplies a(i,k).  64x64x16 tiles, 4x4 per thread.
// Used for Y = U^T A and the backward GEMM, and as the SIMT Gram (gram_impl = 1).
// ------------------------------------------------------------------------------
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void from_f(float* p, float v, bool acc) { *p = acc ? (*p + v) : v; }
__device__ __forceinline__ void from_f(__nv_bfloat16* p, float v, bool acc) {
  *p = __float2bfloat16_rn(acc ? (__bfloat162float(*p) + v) : v);
}

template <typename TA, typename TB, typename TC, bool TRANS_A, bool TRANS_B>
__global__ void __launch_bounds__(256) sgemm_batched_kernel(const TA* __restrict__ A, const TB* __restrict__ Bm,
                                                            TC* __restrict__ Cm, int M, int N, int K, int64_t lda,
                                                            int64_t ldb, int64_t ldc, int64_t strideA,
                                                            int64_t strideB, int64_t strideC,
                                                            const float* __restrict__ kscale, int64_t strideS,
                                                            int accumulate) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int bz = blockIdx.z;
  A += int64_t(bz) * strideA;
  Bm += int64_t(bz) * strideB;
  Cm += int64_t(bz) * strideC;
  if (kscale) kscale += int64_t(bz) * strideS;
  const int i_base = blockIdx.y * 64, j_base = blockIdx.x * 64;
  const int tid = threadIdx.x;
  const int ti = tid / 16, tj = tid % 16;
  float acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;

  for (int k0 = 0; k0 < K; k0 += 16) {
    // ---- stage A tile (64 x 16) into As[k][i] ----
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = tid + e * 256;
      int i, k;
      if (TRANS_A) { i = idx % 64; k = idx / 64; } else { k = idx % 16; i = idx / 16; }
      const int gi = i_base + i, gk = k0 + k;
      float v = 0.f;
      if (gi < M && gk < K) {
        v = to_f(TRANS_A ? A[int64_t(gk) * lda + gi] : A[int64_t(gi) * lda + gk]);
        if (kscale) v *= kscale[gk];
      }
      As[k][i] = v;
    }
    // ---- stage B tile (16 x 64) into Bs[k][j] ----
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = tid + e * 256;
      int j, k;
      if (TRANS_B) { k = idx % 16; j = idx / 16; } else { j = idx % 64; k = idx / 64; }
      const int gj = j_base + j, gk = k0 + k;
      float v = 0.f;
      if (gj < N && gk < K) v = to_f(TRANS_B ? Bm[int64_t(gj) * ldb + gk] : Bm[int64_t(gk) * ldb + gj]);
      Bs[k][j] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ti * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tj * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(av[u], bv[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int gi = i_base + ti * 4 + u;
    if (gi >= M) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int gj = j_base + tj * 4 + v;
      if (gj < N) from_f(Cm + int64_t(gi) * ldc + gj, acc[u][v], accumulate != 0);
    }
  }
}

template <typename TA, typename TB, typename TC>
static int sgemm_launch(bool ta, bool tb, const TA* A, const TB* Bm, TC* Cm, int M, int N, int K, int64_t lda,
                        int64_t ldb, int64_t ldc, int64_t sA, int64_t sB, int64_t sC, const float* kscale,
                        int64_t sS, int accumulate, int batch, cudaStream_t st) {
  dim3 grid((N + 63) / 64, (M + 63) / 64, batch);
#define R3D_GEMM(TA_, TB_) \
  sgemm_batched_kernel<TA, TB, TC, TA_, TB_><<<grid, 256, 0, st>>>(A, Bm, Cm, M, N, K, lda, ldb, ldc, sA, sB, sC, kscale, sS, accumulate)
  if (ta && tb) R3D_GEMM(true, true);
  else if (ta) R3D_GEMM(true, false);
  else if (tb) R3D_GEMM(false, true);
  else R3D_GEMM(false, false);
#undef R3D_GEMM
  R3D_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------
// Jacobi workspace layout (all per batch):
//   Gp  (np, np)  padded symmetric working matrix        Vt (np, np) rows -> eigenvectors
//   Qb  (nt, 2, 64, 64) TRANSPOSED rotation products Q^T of the current round, nt = nb/2 tasks, PRE-SPLIT for the 3xTF32
//       tensor-core updates: plane 0 = hi (round-to-nearest TF32), plane 1 = lo = Q^T - hi (exact), so hi + lo == Q^T
//   cnt (JMAX_SWEEPS) significant rotations per sweep     qflag (nt) task rotated anything
//   nu  (1) absolute significance floor
// ------------------------------------------------------------------------------
struct JacobiWs {
  float* Gp; float* Vt; float* H; float* Qb[2]; int* cnt; int* qflag[2]; float* nu;   // nact = cnt + B*JMAX_SWEEPS
  int* psync;   // 2 * kPanelSyncGroups + 1 counters of the merged panel schedule (cleared by the inner solver)
  int np, nb, nt;
  // spread schedule (np % 128 == 0): group-local problems of a super-round, B * np/128 matrices of 128 x 128
  float* Sg; float* Sh; float* Pv; float* Pt[2];   // local G, its ping-pong scratch, local V (= P), P^T (double-buffered)
  float* Ql[2];                                    // Q^T of the local rounds (2 tasks of 64 x 64 per group)
  int* lflag[3]; int* gflag[2];                    // per local task flags of the three rounds; per group flags
  // chained schedule (jacobi_schedule = 2): Q^T and task flags of two slots x three rounds (alias the buffers above)
  float* Qc[6]; int* qflagc[6];
};

constexpr int QSTR = 2 * JM * JM;        // floats per task in a Q^T buffer: hi plane, lo plane
__device__ __forceinline__ float q_hi_of(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

static inline int jacobi_np(int64_t n) { return int(((n + JM - 1) / JM) * JM); }

// The Jacobi working matrices (Gp, H, V) are stored column-block-major: [np/32 column blocks][np rows][32],
// so a column-panel tile (128 rows x one 32-column block) is one contiguous 16 KB run in HBM and a 32x32
// sub-block is a contiguous 4 KB run.
__host__ __device__ __forceinline__ int64_t boff(int np, int r, int c) {
  return (int64_t(c >> 5) * np + r) * JB + (c & (JB - 1));
}

static size_t jacobi_ws_bytes_one(int64_t B, int64_t n) {   // what jacobi_carve(ws, B, n) uses, plus alignment slack
  const size_t np = jacobi_np(n), nt = np / JM;
  size_t f = size_t(B) * (3 * np * np + 2 * nt * QSTR + 1);
  size_t i = size_t(B) * (JMAX_SWEEPS + 2 * nt) + JMAX_SWEEPS + (2 * kPanelSyncGroups + 8);
  if (np % 128 == 0) {                                  // spread / chained schedules: Sg, Sh, Pv, Pt[2], Ql[2], Qc[5] + flags
    f += size_t(B) * 8 * np * 128;
    i += size_t(B) * 8 * nt;
  }
  return f * 4 + i * 4 + 4096;
}
constexpr int kMaxChunksWs = 4;                         // == kMaxChunks (declared with the stream sets below)
static size_t jacobi_ws_bytes(int64_t B, int64_t n) {   // room for up to kMaxChunks chunk carvings of the batch
  return jacobi_ws_bytes_one(B, n) + size_t(kMaxChunksWs) * (jacobi_ws_bytes_one(0, n) + 512);
}

static JacobiWs jacobi_carve(void* ws, int64_t B, int64_t n) {
  JacobiWs w;
  w.np = jacobi_np(n); w.nb = w.np / JB; w.nt = w.np / JM;
  char* p = (char*)ws;
  p = (char*)((uintptr_t(p) + 255) & ~uintptr_t(255));
  const size_t np2 = size_t(w.np) * w.np;
  w.Gp = (float*)p; p += size_t(B) * np2 * 4;
  w.Vt = (float*)p; p += size_t(B) * np2 * 4;
  w.H = (float*)p; p += size_t(B) * np2 * 4;
  w.Qb[0] = (float*)p; p += size_t(B) * w.nt * QSTR * 4;
  w.Qb[1] = (float*)p; p += size_t(B) * w.nt * QSTR * 4;
  w.nu = (float*)p; p += size_t(B) * 4;
  w.cnt = (int*)p; p += (size_t(B) * JMAX_SWEEPS + JMAX_SWEEPS) * 4;   // per-matrix counts, then nact[JMAX_SWEEPS]
  w.qflag[0] = (int*)p; p += size_t(B) * w.nt * 4;
  w.qflag[1] = (int*)p; p += size_t(B) * w.nt * 4;
  w.psync = (int*)p; p += size_t(2 * kPanelSyncGroups + 8) * 4;
  w.Sg = w.Sh = w.Pv = w.Pt[0] = w.Pt[1] = nullptr;
  if (w.np % 128 == 0) {
    p = (char*)((uintptr_t(p) + 255) & ~uintptr_t(255));
    const size_t gsz = size_t(B) * w.np * 128 * 4;     // B * (np/128) groups x 128 x 128 floats
    w.Sg = (float*)p; p += gsz;
    w.Sh = (float*)p; p += gsz;
    w.Pv = (float*)p; p += gsz;
    w.Pt[0] = (float*)p; p += gsz;
    w.Pt[1] = (float*)p; p += gsz;
    w.Ql[0] = (float*)p; p += gsz;                    // B * ng * 2 local tasks x 2 planes x 64 x 64
    w.Ql[1] = (float*)p; p += gsz;
    for (int k = 0; k < 3; ++k) { w.lflag[k] = (int*)p; p += size_t(B) * w.nt * 4; }   // B * ng * 2 local tasks = B * nt
    for (int k = 0; k < 2; ++k) { w.gflag[k] = (int*)p; p += size_t(B) * (w.nt / 2) * 4; }
    // the chained schedule never runs the group-local problems: five of its six Q^T buffers (B * nt * 2 * 64 * 64 floats
    // = one group buffer each) live in Sg, Sh, Pv, Pt[0], Pt[1]
    float* const big[5] = {w.Sg, w.Sh, w.Pv, w.Pt[0], w.Pt[1]};
    for (int k = 0; k < 5; ++k) w.Qc[k] = big[k];
    w.Qc[5] = (float*)p; p += gsz;
    for (int k = 0; k < 3; ++k) w.qflagc[k] = w.lflag[k];
    for (int k = 3; k < 6; ++k) { w.qflagc[k] = (int*)p; p += size_t(B) * w.nt * 4; }
  }
  return w;
}

// Gp <- zero-padded copy of G; Vt <- I; cnt <- 0; nu <- 2^-21 * max |diag|.
__global__ void jacobi_init_kernel(const float* __restrict__ G, int n, int np, float* __restrict__ Gp,
                                   float* __restrict__ Vt, int* __restrict__ cnt, float* __restrict__ nu,
                                   float nu_ulps) {
  const int b = blockIdx.y;
  const float* g = G + int64_t(b) * n * n;
  float* gp = Gp + int64_t(b) * np * np;
  float* vt = Vt + int64_t(b) * np * np;
  const int64_t total = int64_t(np) * np;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
    // e enumerates the blocked layout directly: e = (cb * np + i) * 32 + cj
    const int cj = int(e & (JB - 1));
    const int64_t t = e >> 5;
    const int i = int(t % np), j = int(t / np) * JB + cj;
    gp[e] = (i < n && j < n) ? g[int64_t(i) * n + j] : 0.f;
    vt[e] = (i == j) ? 1.f : 0.f;
  }
  if (blockIdx.x == 0) {
    __shared__ float smax[256];
    float m = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(g[int64_t(i) * n + i]));
    smax[threadIdx.x] = m;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s) smax[threadIdx.x] = fmaxf(smax[threadIdx.x], smax[threadIdx.x + s]);
      __syncthreads();
    }
    // nu_ulps > 0: absolute floor nu_ulps * 2^-23 * max|diag| (default 4);  nu_ulps < 0: scale-free mode, stored negative:
    // a rotation counts as significant when it passes the relative test and at least one of its two diagonal entries
    // exceeds |nu_ulps| * max|diag| (see jacobi_rotation)
    if (threadIdx.x == 0) nu[b] = nu_ulps > 0.f ? smax[0] * nu_ulps * 1.1920929e-7f : smax[0] * nu_ulps;
    for (int i = threadIdx.x; i < JMAX_SWEEPS; i += blockDim.x) cnt[b * JMAX_SWEEPS + i] = 0;
  }
}

// nact[sweep] = number of matrices that still rotated something significant in sweep-1.  Every Jacobi kernel
// returns at once when it is zero, so the launches after global convergence cost launch latency only.
__global__ void jacobi_active_kernel(int* __restrict__ cnt, int B, int sweep) {
  __shared__ int total;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  int local = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) local += (cnt[b * JMAX_SWEEPS + sweep - 1] != 0);
  if (local) atomicAdd(&total, local);
  __syncthreads();
  if (threadIdx.x == 0) cnt[B * JMAX_SWEEPS + sweep] = total;
}

// pair t of round r over m (even) blocks: r >= 0 circle-method round robin; r < 0 the XOR matching with mask -r
// (block i is paired with i ^ mask; i = t with a zero bit inserted at the mask's top bit)
__device__ __forceinline__ void rr_pair(int m, int r, int t, int& a, int& b) {
  if (r < 0) {
    const int mask = -r, hb = 31 - __clz(mask);
    a = ((t >> hb) << (hb + 1)) | (t & ((1 << hb) - 1));
    b = a ^ mask;
    return;
  }
  if (m == 2) { a = 0; b = 1; return; }
  int x, y;
  if (t == 0) { x = r; y = m - 1; }
  else { x = (r + t) % (m - 1); y = (r - t + (m - 1)) % (m - 1); }
  a = min(x, y); b = max(x, y);
}

__device__ __forceinline__ int blk_row(int I, int J, int i) { return (i < JB) ? I * JB + i : J * JB + (i - JB); }

// ------------------------------------------------------------------------------
// Inner solver: one CTA runs parallel-ordered two-sided Jacobi on the 64x64 sub-block
// G[IJ, IJ] of one block pair in shared memory and emits the accumulated rotation
// product Q^T (64x64).  Per step one warp derives the 32 disjoint rotations; then every
// thread applies the row rotation a AND the column rotation b to 2x2 "pair blocks"
// {p_a,q_a} x {p_b,q_b} of S in registers (S <- J^T S J needs no barrier between its two
// halves) and rotates rows of Q^T.  Two barriers per step.  One inner sweep per visit is
// enough: the outer sweep count is set by the block round-robin (measured).
//
// Round 0 of a sweep visits every pair of the 64 columns (63 steps, circle method), so the
// pairs inside each block are rotated once per sweep.  Later rounds rotate only the 32x32
// cross pairs with the XOR schedule (t, 32 + (t ^ s)), s = 0..31: half the steps, and an
// aligned group of four columns maps to an aligned group of four, so every shared-memory
// access of the step is a conflict-free 128-bit access.
// ------------------------------------------------------------------------------
constexpr int SP = JM + 4;   // row pitch of S and Qt: 16-byte aligned rows

// Scaled ("fast Givens") rotations.  The kernel stores S~ and Q~^T with S = D S~ D and Q^T = D Q~^T (D = diag(d),
// d starts at 1).  The Jacobi rotation J = c [[1, t], [-t, 1]] on the true matrices becomes, on the stored ones,
//     x~_p <- x~_p - tau_pq x~_q ,   x~_q <- x~_q + tau_qp x~_p      (two FFMAs per element pair instead of four ops)
// with tau_pq = t d_q / d_p, tau_qp = t d_p / d_q, followed by d_p <- c d_p, d_q <- c d_q.  c in [1/sqrt(2), 1], so
// over one visit d stays above 2^-32 and nothing over- or underflows; D is applied once when Q^T is written out.
// The relative convergence test |S_pq| > tol sqrt(|S_pp S_qq|) is scale invariant, so it reads S~ directly.
__device__ __forceinline__ bool jacobi_rotation(float spp, float sqq, float spq, float dp, float dq, float tol,
                                                float nu_abs, float& tau_pq, float& tau_qp, float& c, bool& sig) {
  const bool rt = (spq != 0.f) && (fabsf(spq) > tol * sqrtf(fabsf(spp * sqq)));
  // significance (what keeps the sweeps going): an absolute floor on the true off-diagonal entry (nu_abs > 0), or --
  // for a graded matrix, whose entries are accurate relative to their own rows (the second pass) -- the relative test
  // itself unless both directions lie below the numerical-rank cut-off (-nu_abs), where only noise is left
  sig = rt && (nu_abs >= 0.f ? fabsf(spq) * dp * dq > nu_abs : fmaxf(fabsf(spp) * dp * dp, fabsf(sqq) * dq * dq) > -nu_abs);
  tau_pq = 0.f; tau_qp = 0.f; c = 1.f;
  if (rt) {
    const float r = __fdividef(dq, dp);                                  // d_q / d_p
    const float zeta = __fdividef(r * sqq - __fdividef(spp, r), 2.f * spq);   // (S_qq - S_pp) / (2 S_pq) on true values
    const float tt = __fdividef(zeta >= 0.f ? 1.f : -1.f, fabsf(zeta) + sqrtf(fmaf(zeta, zeta, 1.f)));
    c = rsqrtf(fmaf(tt, tt, 1.f));
    tau_pq = tt * r;
    tau_qp = __fdividef(tt, r);
  }
  return rt;
}

__device__ __forceinline__ void permute4(float4& v, int sel) {   // v[i] <- v[i ^ sel], sel warp-uniform
  if (sel & 1) { float t = v.x; v.x = v.y; v.y = t; t = v.z; v.z = v.w; v.w = t; }
  if (sel & 2) { float t = v.x; v.x = v.z; v.z = t; t = v.y; v.y = v.w; v.w = t; }
}

__global__ void __launch_bounds__(256, 5) jacobi_inner_kernel(float* __restrict__ Gp, int np, int nb, int nt, int round,
                                                              int sweep, int* __restrict__ cnt,
                                                              int* __restrict__ qflag, float* __restrict__ Qb,
                                                              float tol, const float* __restrict__ nu,
                                                              int max_inner, int* __restrict__ psync, int bdiv,
                                                              int generic) {
  // bdiv > 1: the batch is the set of group-local problems of the spread schedule, bdiv per matrix; the convergence
  // bookkeeping (cnt, nact, nu) stays per matrix
  const int b = blockIdx.y, t = blockIdx.x, bm = b / bdiv;
  if (sweep > 0 && cnt[(gridDim.y / bdiv) * JMAX_SWEEPS + sweep] == 0) return;   // every matrix converged
  if (psync != nullptr && b == 0 && t == 0)                             // counters of this round's merged panel launch
    for (int i = threadIdx.x; i < 2 * kPanelSyncGroups + 1; i += 256) psync[i] = 0;
  if (sweep > 0 && cnt[bm * JMAX_SWEEPS + sweep - 1] == 0) return;   // this matrix converged
  __shared__ __align__(16) float S[JM][SP];          // S~
  __shared__ __align__(16) float Qt[JM][SP];         // Q~^T: Qt[i][k] = Q[k][i] / d_i
  __shared__ __align__(16) float4 rot[JB];           // generic rounds: {tau_pq, tau_qp, bits(p), bits(q)}
  __shared__ __align__(16) float ra_[JB], rb_[JB];   // cross rounds: tau_pq and tau_qp of pair t
  __shared__ float dsc[JM];                          // d
  __shared__ int s_any, s_sig, s_tot;
  const int tid = threadIdx.x;
  int I, J;
  rr_pair(nb, round, t, I, J);
  float* g = Gp + int64_t(b) * np * np;
  for (int e = tid; e < JM * JM; e += 256) {
    const int i = e / JM, j = e % JM;
    S[i][j] = g[boff(np, blk_row(I, J, i), blk_row(I, J, j))];
    Qt[i][j] = (i == j) ? 1.f : 0.f;
  }
  if (tid < JM) dsc[tid] = 1.f;
  if (tid == 0) { s_tot = 0; s_sig = 0; }
  __syncthreads();
  // symmetrise (the tile updates leave eps-level asymmetry)
  for (int e = tid; e < JM * JM; e += 256) {
    const int i = e / JM, j = e % JM;
    if (i < j) {
      const float v = 0.5f * (S[i][j] + S[j][i]);
      S[i][j] = v; S[j][i] = v;
    }
  }
  const float nu_abs = nu[bm];
  __syncthreads();
  {
    // Early out: when no pair of this visit passes the rotation test on the initial block, the first step rotates
    // nothing, the block stays as it is, and so does every later step -- the visit is the identity.  Typical for the
    // last sweep of a pass and for the null space of rank-deficient samples.
    int any = 0;
    for (int e = tid; e < JM * JM; e += 256) {
      const int i = e / JM, j = e % JM;
      if (i < j && (generic || (i < JB && j >= JB))) {
        const float v = S[i][j];
        any |= (v != 0.f) && (fabsf(v) > tol * sqrtf(fabsf(S[i][i] * S[j][j])));
      }
    }
    if (!__syncthreads_or(any)) {
      float* qo = Qb + (int64_t(b) * nt + t) * QSTR;
      for (int e = tid; e < JM * JM; e += 256) { qo[e] = (e / JM == e % JM) ? 1.f : 0.f; qo[JM * JM + e] = 0.f; }
      if (tid == 0) qflag[b * nt + t] = 0;
      return;
    }
  }

  int sig_total = 0;
  for (int it = 0; it < max_inner; ++it) {
    if (generic) {
      // ---------------- generic schedule: all pairs of the 64 columns ----------------
      for (int s = 0; s < JM - 1; ++s) {
        if (tid < JB) {
          int p, q;
          rr_pair(JM, s, tid, p, q);
          float tpq, tqp, c; bool sg;
          const float dp = dsc[p], dq = dsc[q];
          const bool rt = jacobi_rotation(S[p][p], S[q][q], S[p][q], dp, dq, tol, nu_abs, tpq, tqp, c, sg);
          rot[tid] = make_float4(tpq, tqp, __int_as_float(p), __int_as_float(q));
          if (rt) { dsc[p] = dp * c; dsc[q] = dq * c; }
          const unsigned any = __ballot_sync(0xffffffffu, rt), sgm = __ballot_sync(0xffffffffu, sg);
          if (tid == 0) { s_any = (any != 0); s_sig += __popc(sgm); s_tot += __popc(any); }
        }
        __syncthreads();
        if (s_any) {
#pragma unroll
          for (int u = 0; u < (JB * JB) / 256; ++u) {
            const int item = tid + u * 256;
            const float4 ra = rot[item / JB], rb = rot[item % JB];
            const int pa = __float_as_int(ra.z), qa = __float_as_int(ra.w);
            const int pb = __float_as_int(rb.z), qb = __float_as_int(rb.w);
            const float x00 = S[pa][pb], x01 = S[pa][qb], x10 = S[qa][pb], x11 = S[qa][qb];
            const float y00 = fmaf(-ra.x, x10, x00), y10 = fmaf(ra.y, x00, x10);
            const float y01 = fmaf(-ra.x, x11, x01), y11 = fmaf(ra.y, x01, x11);
            S[pa][pb] = fmaf(-rb.x, y01, y00);
            S[pa][qb] = fmaf(rb.y, y00, y01);
            S[qa][pb] = fmaf(-rb.x, y11, y10);
            S[qa][qb] = fmaf(rb.y, y10, y11);
          }
#pragma unroll
          for (int u = 0; u < (JB * (JM / 4)) / 256; ++u) {
            const int item = tid + u * 256;
            const float4 ra = rot[item / (JM / 4)];
            const int col = (item % (JM / 4)) * 4;
            const int pa = __float_as_int(ra.z), qa = __float_as_int(ra.w);
            const float4 vp = *reinterpret_cast<const float4*>(&Qt[pa][col]);
            const float4 vq = *reinterpret_cast<const float4*>(&Qt[qa][col]);
            *reinterpret_cast<float4*>(&Qt[pa][col]) =
                make_float4(fmaf(-ra.x, vq.x, vp.x), fmaf(-ra.x, vq.y, vp.y), fmaf(-ra.x, vq.z, vp.z),
                            fmaf(-ra.x, vq.w, vp.w));
            *reinterpret_cast<float4*>(&Qt[qa][col]) =
                make_float4(fmaf(ra.y, vp.x, vq.x), fmaf(ra.y, vp.y, vq.y), fmaf(ra.y, vp.z, vq.z),
                            fmaf(ra.y, vp.w, vq.w));
          }
        }
        __syncthreads();
      }
    } else {
      // ---------------- cross schedule: pairs (t, 32 + (t ^ s)) ----------------
      const int a = tid >> 3;            // row pair owned by this thread
      const int g4 = (tid & 7) * 4;      // four column pairs b = g4 .. g4+3
      for (int s = 0; s < JB; ++s) {
        if (tid < JB) {
          const int p = tid, q = JB + (tid ^ s);
          float tpq, tqp, c; bool sg;
          const float dp = dsc[p], dq = dsc[q];
          const bool rt = jacobi_rotation(S[p][p], S[q][q], S[p][q], dp, dq, tol, nu_abs, tpq, tqp, c, sg);
          ra_[tid] = tpq; rb_[tid] = tqp;
          if (rt) { dsc[p] = dp * c; dsc[q] = dq * c; }
          const unsigned any = __ballot_sync(0xffffffffu, rt), sgm = __ballot_sync(0xffffffffu, sg);
          if (tid == 0) { s_any = (any != 0); s_sig += __popc(sgm); s_tot += __popc(any); }
        }
        __syncthreads();
        if (s_any) {
          const int qa = JB + (a ^ s);
          const int gq = JB + (g4 ^ (s & ~3));      // aligned group holding the partners of columns g4..g4+3
          const int sel = s & 3;
          const float ta = ra_[a], ua = rb_[a];     // tau_pq, tau_qp of the row pair
          const float4 tb = *reinterpret_cast<const float4*>(&ra_[g4]);
          const float4 ub = *reinterpret_cast<const float4*>(&rb_[g4]);
          float4 x00 = *reinterpret_cast<const float4*>(&S[a][g4]);
          float4 x10 = *reinterpret_cast<const float4*>(&S[qa][g4]);
          float4 x01 = *reinterpret_cast<const float4*>(&S[a][gq]);
          float4 x11 = *reinterpret_cast<const float4*>(&S[qa][gq]);
          permute4(x01, sel);                       // now component i is the partner column of g4 + i
          permute4(x11, sel);
          float4 o00, o01, o10, o11;
#define R3D_BLOCK(F)                                                                   \
          {                                                                              \
            const float y00 = fmaf(-ta, x10.F, x00.F), y10 = fmaf(ua, x00.F, x10.F);     \
            const float y01 = fmaf(-ta, x11.F, x01.F), y11 = fmaf(ua, x01.F, x11.F);     \
            o00.F = fmaf(-tb.F, y01, y00); o01.F = fmaf(ub.F, y00, y01);                 \
            o10.F = fmaf(-tb.F, y11, y10); o11.F = fmaf(ub.F, y10, y11);                 \
          }
          R3D_BLOCK(x) R3D_BLOCK(y) R3D_BLOCK(z) R3D_BLOCK(w)
#undef R3D_BLOCK
          permute4(o01, sel);                       // the permutation is an involution
          permute4(o11, sel);
          *reinterpret_cast<float4*>(&S[a][g4]) = o00;
          *reinterpret_cast<float4*>(&S[qa][g4]) = o10;
          *reinterpret_cast<float4*>(&S[a][gq]) = o01;
          *reinterpret_cast<float4*>(&S[qa][gq]) = o11;
          // Q~t <- T^T Q~t: rows a and qa; each quarter-warp touches 128 contiguous bytes per access
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int col = g4 + 32 * h;
            const float4 vp = *reinterpret_cast<const float4*>(&Qt[a][col]);
            const float4 vq = *reinterpret_cast<const float4*>(&Qt[qa][col]);
            *reinterpret_cast<float4*>(&Qt[a][col]) =
                make_float4(fmaf(-ta, vq.x, vp.x), fmaf(-ta, vq.y, vp.y), fmaf(-ta, vq.z, vp.z), fmaf(-ta, vq.w, vp.w));
            *reinterpret_cast<float4*>(&Qt[qa][col]) =
                make_float4(fmaf(ua, vp.x, vq.x), fmaf(ua, vp.y, vq.y), fmaf(ua, vp.z, vq.z), fmaf(ua, vp.w, vq.w));
          }
        }
        __syncthreads();
      }
    }
    const int sig_now = s_sig;        // stable: the step loop ended on a barrier
    __syncthreads();                  // nobody may start the next sweep before all have read it
    if (sig_now == sig_total) break;  // nothing significant in this inner sweep
    sig_total = sig_now;
  }
  if (tid == 0 && sig_total > 0) atomicAdd(&cnt[bm * JMAX_SWEEPS + sweep], sig_total);
  float* qo = Qb + (int64_t(b) * nt + t) * QSTR;
  for (int e = tid; e < JM * JM; e += 256) {                                            // Q^T = D Q~^T, row-major, hi | lo
    const float x = dsc[e / JM] * Qt[e / JM][e % JM], h = q_hi_of(x);
    qo[e] = h; qo[JM * JM + e] = x - h;
  }
  if (tid == 0) qflag[b * nt + t] = (s_tot > 0);
}

// ------------------------------------------------------------------------------
// Register-resident inner solver for the cross rounds (round > 0).
//
// The shared-memory solver above moves all of S and Q^T through shared memory every step (64 KB per CTA-step).
// Here the whole state lives in registers.  With the pairing (a, 32 + (a ^ sigma)), a "pair block" (a, b) of S is
//     TL = S[a][b]   TR = S[a][32+(b^sigma)]   BL = S[32+(a^sigma)][b]   BR = S[32+(a^sigma)][32+(b^sigma)]
// and both the row rotation of pair a and the column rotation of pair b act inside it.  A thread owns the 2x2
// pair blocks a in {2A, 2A+1}, b in {2B, 2B+1} (A, B: 4 bits each) and, of Q~^T, rows a and 32+(a^sigma) over
// columns 4B..4B+3: 32 state registers.  sigma runs through the 5-bit Gray code, so between steps exactly one
// bit beta of sigma flips and TR moves to the thread whose b differs in bit beta, BL (and the bottom rows of
// Q~^T) to the one whose a differs, BR both.  Threads are numbered by (A, D = A ^ B):
//     lane = A0 | A1<<1 | D0<<2 | D1<<3 | A2<<4,   warp = D2 | D3<<1 | A3<<2
// so that
//     beta = 0 (16 of 31 transitions): inside the thread -- a register renaming;
//     beta = 1, 2 (12 transitions):    lane bits only -- 20 __shfl_xor per thread;
//     beta = 3, 4 (3 transitions):     warp bits -- the moving state goes through shared memory once;
// and the 16 threads holding the diagonal pair blocks (D = 0) sit in two warps (0 and 4), which derive the 32
// rotations of a step one per lane (the second rotation of a diagonal thread is handed to lane ^ 4).  Per step:
// rotations in 2 warps, one barrier, 48 FFMAs per thread.  Same thresholds and outputs (Q^T, qflag, cnt) as the
// kernel above; the rotation uses 5 MUFU ops (no IEEE sqrt/div: c and tau only have to be consistent to ~2 ulp,
// the final row normalisation of the eigenvectors absorbs the drift, as it does for rsqrtf above).
// ------------------------------------------------------------------------------
__device__ __forceinline__ bool jacobi_rotation_fast(float spp, float sqq, float spq, float dp, float dq, float tol2,
                                                     float nu_abs, float& tau_pq, float& tau_qp, float& c, bool& sig) {
  const float dpq = dp * dq;
  const bool rt = spq * spq > tol2 * fabsf(spp * sqq);
  sig = rt && (nu_abs >= 0.f ? fabsf(spq) * dpq > nu_abs : fmaxf(fabsf(spp) * dp * dp, fabsf(sqq) * dq * dq) > -nu_abs);
  tau_pq = 0.f; tau_qp = 0.f; c = 1.f;
  if (rt) {
    const float rp = __frcp_rn(dpq);                 // dp, dq in [2^-32, 1]: dpq >= 2^-64, no overflow
    const float num = dq * dq * sqq - dp * dp * spp; // (S_qq - S_pp) on true values, times dp dq
    float zeta = num * rp * __fdividef(0.5f, spq);   // (S_qq - S_pp) / (2 S_pq)
    const float az = fminf(fabsf(zeta), 1e18f);
    const float zz = fmaf(az, az, 1.f);
    const float tt = copysignf(__fdividef(1.f, az + zz * rsqrtf(zz)), zeta);
    c = rsqrtf(fmaf(tt, tt, 1.f));
    tau_pq = tt * dq * dq * rp;                      // t d_q / d_p
    tau_qp = tt * dp * dp * rp;                      // t d_p / d_q
  }
  return rt;
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB) jacobi_inner_cross_kernel(float* __restrict__ Gp, int np, int nb, int nt,
                                                                    int round, int sweep, int* __restrict__ cnt,
                                                                    int* __restrict__ qflag, float* __restrict__ Qb,
                                                                    float tol, const float* __restrict__ nu,
                                                                    int* __restrict__ psync, int bdiv) {
  const int b = blockIdx.y, t = blockIdx.x, bm = b / bdiv;              // bdiv: see jacobi_inner_kernel
  if (sweep > 0 && cnt[(gridDim.y / bdiv) * JMAX_SWEEPS + sweep] == 0) return;   // every matrix converged
  if (psync != nullptr && b == 0 && t == 0)                             // counters of this round's merged panel launch
    for (int i = threadIdx.x; i < 2 * kPanelSyncGroups + 1; i += 256) psync[i] = 0;
  if (sweep > 0 && cnt[bm * JMAX_SWEEPS + sweep - 1] == 0) return;      // this matrix converged
  // ONE 17 KB staging buffer (initial tile, the three cross-warp transitions, final Q^T): small enough that a CTA
  // of this kernel fits next to two resident CTAs of the panel update (V on the side stream)
  __shared__ __align__(16) float S[JM][SP];
  __shared__ __align__(16) float2 par[2][JB];      // {tau_pq, tau_qp} of pair a, double buffered by step parity
  __shared__ float dsc[JM];
  __shared__ int s_sig, s_tot;
  const int tid = threadIdx.x, lane = tid & 31;
  // Logical warp id.  The two rotation warps (logical 0 and 4) issue about twice the instructions of the others;
  // they are placed on physical warps {0,1} or {2,3} (alternating with the CTA's position in the launch order), so
  // that the four schedulers of an SM, which serve physical warps w % 4 of several resident CTAs, stay balanced.
  const int pw = (tid >> 5) ^ ((((blockIdx.y * gridDim.x + blockIdx.x) / kNumSMs) & 1) << 1);
  const int warp = ((pw & 1) << 2) | (pw >> 1);
  const int A = (lane & 3) | (((lane >> 4) & 1) << 2) | (((warp >> 2) & 1) << 3);
  const int D = ((lane >> 2) & 3) | ((warp & 3) << 2);
  const int Bc = A ^ D;
  int I, J;
  rr_pair(nb, round, t, I, J);
  const float* g = Gp + int64_t(b) * np * np;
  {
    float v[JM * JM / 256];
#pragma unroll
    for (int u = 0; u < JM * JM / 256; ++u) {
      const int e = tid + u * 256, i = e / JM, j = e % JM;
      v[u] = g[boff(np, blk_row(I, J, i), blk_row(I, J, j))];
    }
#pragma unroll
    for (int u = 0; u < JM * JM / 256; ++u) {
      const int e = tid + u * 256;
      S[e / JM][e % JM] = v[u];
    }
  }
  if (tid < JM) dsc[tid] = 1.f;
  if (tid == 0) { s_tot = 0; s_sig = 0; }
  __syncthreads();

  float TL[2][2], TR[2][2], BL[2][2], BR[2][2], QT[2][4], QB[2][4];
  int sg = 0;                                       // current sigma
  // initial state: symmetrised S (the tile updates leave eps-level asymmetry), Q~^T = I
#pragma unroll
  for (int ra = 0; ra < 2; ++ra) {
    const int a = 2 * A + ra, u = JB + a;
#pragma unroll
    for (int rb = 0; rb < 2; ++rb) {
      const int c = 2 * Bc + rb, v = JB + c;
      TL[ra][rb] = 0.5f * (S[a][c] + S[c][a]);
      TR[ra][rb] = 0.5f * (S[a][v] + S[v][a]);
      BL[ra][rb] = 0.5f * (S[u][c] + S[c][u]);
      BR[ra][rb] = 0.5f * (S[u][v] + S[v][u]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      QT[ra][c] = (a == 4 * Bc + c) ? 1.f : 0.f;
      QB[ra][c] = (u == 4 * Bc + c) ? 1.f : 0.f;
    }
  }
  const float nu_abs = nu[bm], tol2 = tol * tol;
  {
    // Early out (see jacobi_inner_kernel): the thread's four TR entries are cross pairs (a, 32 + c), and the 256
    // threads together hold all 1024 of them; if none passes the rotation test, the whole visit is the identity.
    int any = 0;
#pragma unroll
    for (int ra = 0; ra < 2; ++ra)
#pragma unroll
      for (int rb = 0; rb < 2; ++rb) {
        const int a = 2 * A + ra, v = JB + 2 * Bc + rb;
        const float x = TR[ra][rb];
        any |= x * x > tol2 * fabsf(S[a][a] * S[v][v]);
      }
    if (!__syncthreads_or(any)) {
      float4* qo = reinterpret_cast<float4*>(Qb + (int64_t(b) * nt + t) * QSTR);
#pragma unroll
      for (int u = 0; u < JM * JM / 4 / 256; ++u) {
        const int e = tid + u * 256, i = e / (JM / 4), j4 = (e % (JM / 4)) * 4;
        qo[e] = make_float4(i == j4 ? 1.f : 0.f, i == j4 + 1 ? 1.f : 0.f, i == j4 + 2 ? 1.f : 0.f, i == j4 + 3 ? 1.f : 0.f);
        qo[JM * JM / 4 + e] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (tid == 0) qflag[b * nt + t] = 0;
      return;
    }
  }
  const bool rot_warp = (warp & 3) == 0;            // warps 0 and 4 hold the diagonal pair blocks (D == 0) ...
  const bool rot_lane = rot_warp && (lane & 8) == 0; // ... in lanes with D1 == 0; lane bit 2 (D0) picks the rotation
  const int rsel = (lane >> 2) & 1;
  int n_tot = 0, n_sig = 0;

  for (int s2 = 0; s2 < JB; s2 += 2) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = s2 + h;
      if (rot_warp) {
        // the diagonal thread (lane & ~4) holds both rotations of its two pairs; lane | 4 takes the second one
        const float pp1 = __shfl_sync(0xffffffffu, TL[1][1], lane & ~4);
        const float qq1 = __shfl_sync(0xffffffffu, BR[1][1], lane & ~4);
        const float pq1 = __shfl_sync(0xffffffffu, TR[1][1], lane & ~4);
        if (rot_lane) {
          const int p = 2 * A + rsel, q = JB + (p ^ sg);
          const float spp = rsel ? pp1 : TL[0][0], sqq = rsel ? qq1 : BR[0][0], spq = rsel ? pq1 : TR[0][0];
          float tpq, tqp, c; bool sgn;
          const float dp = dsc[p], dq = dsc[q];
          const bool rt = jacobi_rotation_fast(spp, sqq, spq, dp, dq, tol2, nu_abs, tpq, tqp, c, sgn);
          par[h][p] = make_float2(tpq, tqp);
          if (rt) { dsc[p] = dp * c; dsc[q] = dq * c; ++n_tot; }
          if (sgn) ++n_sig;
        }
      }
      __syncthreads();
      const float4 pa = *reinterpret_cast<const float4*>(&par[h][2 * A]);    // {tau_pq, tau_qp} of a = 2A, 2A+1
      const float4 pb = *reinterpret_cast<const float4*>(&par[h][2 * Bc]);
#pragma unroll
      for (int ra = 0; ra < 2; ++ra) {
        const float ta = ra ? pa.z : pa.x, ua = ra ? pa.w : pa.y;
#pragma unroll
        for (int rb = 0; rb < 2; ++rb) {
          const float tb = rb ? pb.z : pb.x, ub = rb ? pb.w : pb.y;
          const float x00 = TL[ra][rb], x01 = TR[ra][rb], x10 = BL[ra][rb], x11 = BR[ra][rb];
          const float y00 = fmaf(-ta, x10, x00), y10 = fmaf(ua, x00, x10);
          const float y01 = fmaf(-ta, x11, x01), y11 = fmaf(ua, x01, x11);
          TL[ra][rb] = fmaf(-tb, y01, y00); TR[ra][rb] = fmaf(ub, y00, y01);
          BL[ra][rb] = fmaf(-tb, y11, y10); BR[ra][rb] = fmaf(ub, y10, y11);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float vp = QT[ra][c], vq = QB[ra][c];
          QT[ra][c] = fmaf(-ta, vq, vp);
          QB[ra][c] = fmaf(ua, vp, vq);
        }
      }
      // ---- move to the next sigma ----
      if (h == 0) {                                  // bit 0 flips: renaming inside the thread
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          float x = BL[0][i]; BL[0][i] = BL[1][i]; BL[1][i] = x;
          x = TR[i][0]; TR[i][0] = TR[i][1]; TR[i][1] = x;
        }
        float x = BR[0][0]; BR[0][0] = BR[1][1]; BR[1][1] = x;
        x = BR[0][1]; BR[0][1] = BR[1][0]; BR[1][0] = x;
#pragma unroll
        for (int c = 0; c < 4; ++c) { const float y = QB[0][c]; QB[0][c] = QB[1][c]; QB[1][c] = y; }
        sg ^= 1;
      } else if (k + 1 < JB) {
        const int beta = __ffs(k + 1) - 1;           // Gray code: sigma_{k+1} = sigma_k ^ (1 << ctz(k+1)), beta >= 1
        if (beta <= 2) {
          // bit beta-1 of A and of B: TR moves along D (lane mask 4 << (beta-1)), BR along A (1 << (beta-1)),
          // BL and the bottom rows of Q~^T along both
          const int ma = 1 << (beta - 1), md = 4 << (beta - 1);
#pragma unroll
          for (int ra = 0; ra < 2; ++ra) {
#pragma unroll
            for (int rb = 0; rb < 2; ++rb) {
              BL[ra][rb] = __shfl_xor_sync(0xffffffffu, BL[ra][rb], ma | md);
              TR[ra][rb] = __shfl_xor_sync(0xffffffffu, TR[ra][rb], md);
              BR[ra][rb] = __shfl_xor_sync(0xffffffffu, BR[ra][rb], ma);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) QB[ra][c] = __shfl_xor_sync(0xffffffffu, QB[ra][c], ma | md);
          }
          sg ^= 1 << beta;
        } else {                                     // warp-id bits: through shared memory, S part then Q part
#pragma unroll
          for (int ra = 0; ra < 2; ++ra) {
            const int a = 2 * A + ra, u = JB + (a ^ sg);
#pragma unroll
            for (int rb = 0; rb < 2; ++rb) {
              const int c = 2 * Bc + rb, v = JB + (c ^ sg);
              S[a][v] = TR[ra][rb]; S[u][c] = BL[ra][rb]; S[u][v] = BR[ra][rb];
            }
          }
          __syncthreads();
          const int sg2 = sg ^ (1 << beta);
#pragma unroll
          for (int ra = 0; ra < 2; ++ra) {
            const int a = 2 * A + ra, u = JB + (a ^ sg2);
#pragma unroll
            for (int rb = 0; rb < 2; ++rb) {
              const int c = 2 * Bc + rb, v = JB + (c ^ sg2);
              TR[ra][rb] = S[a][v]; BL[ra][rb] = S[u][c]; BR[ra][rb] = S[u][v];
            }
          }
          __syncthreads();
#pragma unroll
          for (int ra = 0; ra < 2; ++ra) {
            const int u = (2 * A + ra) ^ sg;         // bottom row of Q~^T, stored at row u of the buffer
            *reinterpret_cast<float4*>(&S[u][4 * Bc]) = make_float4(QB[ra][0], QB[ra][1], QB[ra][2], QB[ra][3]);
          }
          __syncthreads();
#pragma unroll
          for (int ra = 0; ra < 2; ++ra) {
            const int u = (2 * A + ra) ^ sg2;
            const float4 q4 = *reinterpret_cast<const float4*>(&S[u][4 * Bc]);
            QB[ra][0] = q4.x; QB[ra][1] = q4.y; QB[ra][2] = q4.z; QB[ra][3] = q4.w;
          }
          sg = sg2;
        }
      }
    }
  }
  if (n_tot) atomicAdd(&s_tot, n_tot);
  if (n_sig) atomicAdd(&s_sig, n_sig);
  // Q~^T to shared memory in canonical order, then Q^T = D Q~^T out, row-major and coalesced
#pragma unroll
  for (int ra = 0; ra < 2; ++ra) {
    const int a = 2 * A + ra, u = JB + (a ^ sg);
    *reinterpret_cast<float4*>(&S[a][4 * Bc]) = make_float4(QT[ra][0], QT[ra][1], QT[ra][2], QT[ra][3]);
    *reinterpret_cast<float4*>(&S[u][4 * Bc]) = make_float4(QB[ra][0], QB[ra][1], QB[ra][2], QB[ra][3]);
  }
  __syncthreads();
  if (tid == 0 && s_sig > 0) atomicAdd(&cnt[bm * JMAX_SWEEPS + sweep], s_sig);
  float4* qo = reinterpret_cast<float4*>(Qb + (int64_t(b) * nt + t) * QSTR);
#pragma unroll
  for (int u = 0; u < JM * JM / 4 / 256; ++u) {
    const int e = tid + u * 256, i = e / (JM / 4), j4 = (e % (JM / 4)) * 4;
    const float d = dsc[i];
    const float4 q4 = *reinterpret_cast<const float4*>(&S[i][j4]);
    const float4 x = make_float4(d * q4.x, d * q4.y, d * q4.z, d * q4.w);
    const float4 h = make_float4(q_hi_of(x.x), q_hi_of(x.y), q_hi_of(x.z), q_hi_of(x.w));
    qo[e] = h;
    qo[JM * JM / 4 + e] = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
  }
  if (tid == 0) qflag[b * nt + t] = (s_tot > 0);
}

// ------------------------------------------------------------------------------
// Tile update of one round:  G[IJ_a, IJ_c] <- Q_a^T G[IJ_a, IJ_c] Q_c  (nt*nt tiles)
// and Vt[IJ_a, slab] <- Q_a^T Vt[IJ_a, slab]  (nt * np/64 tiles).  64x64x64 products
// from shared memory, 4x4 outputs per thread.
// ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) jacobi_update_kernel(float* __restrict__ Gp, float* __restrict__ Vt, int np,
                                                            int nb, int nt, int round, int sweep,
                                                            const int* __restrict__ cnt,
                                                            const int* __restrict__ qflag,
                                                            const float* __restrict__ Qb) {
  const int b = blockIdx.y;
  if (sweep > 0 && cnt[gridDim.y * JMAX_SWEEPS + sweep] == 0) return;
  if (sweep > 0 && cnt[b * JMAX_SWEEPS + sweep - 1] == 0) return;
  extern __shared__ float sm[];
  float (*Tm)[JM + 4] = reinterpret_cast<float (*)[JM + 4]>(sm);
  float (*Qa)[JM + 4] = reinterpret_cast<float (*)[JM + 4]>(sm + JM * (JM + 4));
  float (*Qc)[JM + 4] = reinterpret_cast<float (*)[JM + 4]>(sm + 2 * JM * (JM + 4));
  const int tid = threadIdx.x;
  const int x = blockIdx.x;
  const bool two_sided = x < nt * nt;
  int a, c, slab = 0;
  if (two_sided) { a = x / nt; c = x % nt; }
  else { const int xx = x - nt * nt; const int nslab = np / JM; a = xx / nslab; slab = xx % nslab; c = a; }
  const bool fa = qflag[b * nt + a] != 0, fc = two_sided && (qflag[b * nt + c] != 0);
  if (!fa && !fc) return;
  int Ia, Ja, Ic, Jc;
  rr_pair(nb, round, a, Ia, Ja);
  rr_pair(nb, round, c, Ic, Jc);
  float* base = (two_sided ? Gp : Vt) + int64_t(b) * np * np;
  const float* qa = Qb + (int64_t(b) * nt + a) * QSTR;          // hi plane, then lo plane: hi + lo is exact
  const float* qc = Qb + (int64_t(b) * nt + c) * QSTR;
  for (int e = tid; e < JM * JM; e += 256) {
    const int i = e / JM, j = e % JM;
    const int gj = two_sided ? blk_row(Ic, Jc, j) : slab * JM + j;
    Tm[i][j] = base[boff(np, blk_row(Ia, Ja, i), gj)];
    Qa[j][i] = qa[e] + qa[JM * JM + e];   // Qb holds Q^T: Qa[k][i] = Q[k][i] = qa[i*64+k]
    if (two_sided) Qc[j][i] = qc[e] + qc[JM * JM + e];
  }
  __syncthreads();
  const int ti = tid / 16, tj = tid % 16;
  float acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;
  // T1[i][j] = sum_k Qa[k][i] * Tm[k][j]
#pragma unroll 8
  for (int k = 0; k < JM; ++k) {
    const float4 av = *reinterpret_cast<const float4*>(&Qa[k][ti * 4]);
    const float4 bv = *reinterpret_cast<const float4*>(&Tm[k][tj * 4]);
    const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(aa[u], bb[v], acc[u][v]);
  }
  if (two_sided) {
    __syncthreads();   // everyone is done reading Tm
#pragma unroll
    for (int u = 0; u < 4; ++u)
      *reinterpret_cast<float4*>(&Tm[ti * 4 + u][tj * 4]) = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;
    // out[i][j] = sum_k T1[i][k] * Qc[k][j]
#pragma unroll 8
    for (int k = 0; k < JM; ++k) {
      const float4 bv = *reinterpret_cast<const float4*>(&Qc[k][tj * 4]);
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float aa = Tm[ti * 4 + u][k];
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(aa, bb[v], acc[u][v]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = ti * 4 + u;
    const int j0 = tj * 4;
    const int gj = two_sided ? blk_row(Ic, Jc, j0) : slab * JM + j0;   // 4 consecutive j stay inside one block
    *reinterpret_cast<float4*>(&base[boff(np, blk_row(Ia, Ja, i), gj)]) =
        make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
  }
}

// ------------------------------------------------------------------------------
// Spread schedule.  With nb = 2^k blocks, the rounds of a sweep are the XOR matchings i <-> i ^ mask, mask = 1..nb-1.
// Three masks {a, b, a^b} (a two-dimensional subspace of GF(2)^k minus zero) only ever pair blocks inside the cosets
// {i0, i0^a, i0^b, i0^a^b}: 4-block groups of 128 columns.  A super-round gathers the 128 x 128 diagonal group blocks
// of G into a batch of small matrices (Sg), runs the three rounds on them with the same kernels (inner solve, panel
// update with V accumulation: the local "V" is the product P of the three rounds' rotations), and then streams G and
// V ONCE with the 128 x 128 products P (K = N = 128) instead of three times with 64 x 64 rotations.  The subspaces
// of a sweep partition the masks (a spread of PG(k-1, 2), built from GF(2^k) over GF(4) when k is even; for odd k a
// greedy partial spread and the left-over masks as ordinary single rounds), so every block pair still meets exactly
// once per sweep.  The group-local working set (B * np/128 matrices of 64 KB) stays in L2.
// ------------------------------------------------------------------------------
struct SuperRound { int a, b; };             // b == 0: single round with mask a; else the triple {a, b, a^b}

static std::vector<SuperRound> spread_plan(int nb) {
  std::vector<SuperRound> plan;
  int k = 0;
  while ((1 << k) < nb) ++k;
  if ((1 << k) != nb || k < 3) return plan;                 // needs a power of two >= 8 blocks
  std::vector<char> used(nb, 0);
  if (k % 2 == 0) {
    static const int prim[9] = {0, 0, 0x7, 0, 0x13, 0, 0x43, 0, 0x11d};   // primitive polynomials of degree 2, 4, 6, 8
    if (k <= 8) {
      std::vector<int> e(nb - 1);
      e[0] = 1;
      for (int i = 1; i < nb - 1; ++i) { int v = e[i - 1] << 1; if (v & nb) v ^= prim[k]; e[i] = v; }
      const int third = (nb - 1) / 3;
      for (int i = 0; i < third; ++i) {                       // g^i GF(4)^* = {g^i, g^(i + N/3), g^(i + 2N/3)}
        plan.push_back({e[i], e[i + third]});
        used[e[i]] = used[e[i + third]] = used[e[i] ^ e[i + third]] = 1;
      }
    }
  }
  // whatever is left (odd k, or k > 8): greedy triples, then single rounds
  for (int a = 1; a < nb; ++a) {
    if (used[a]) continue;
    int found = 0;
    for (int b = a + 1; b < nb && !found; ++b)
      if (!used[b] && !used[a ^ b] && (a ^ b) > a) { found = b; }
    used[a] = 1;
    if (found) { used[found] = used[a ^ found] = 1; plan.push_back({a, found}); }
    else plan.push_back({a, 0});
  }
  // the first round of a sweep rotates the pairs inside each block too (generic schedule): it must be a round that
  // exists in every plan -- the first mask of the first super-round, whatever that is
  return plan;
}

static PanelGroups groups_of(const SuperRound& sr) {
  auto top = [](int v) { int h = 0; while ((v >> (h + 1)) != 0) ++h; return h; };
  const int p1 = top(sr.a);
  const int b2 = ((sr.b >> p1) & 1) ? (sr.b ^ sr.a) : sr.b;
  const int p2 = top(b2);
  PanelGroups g;
  g.ga = sr.a; g.gb = sr.b; g.plo = std::min(p1, p2); g.phi = std::max(p1, p2);
  return g;
}

// Sg[b*ng + g] <- the 128 x 128 diagonal block of group g (16 sub-blocks of 32 x 32 = contiguous 4 KB runs in the
// blocked layout of both matrices); Pv <- I.  One CTA per (group, matrix).
__global__ void __launch_bounds__(256) group_gather_kernel(const float* __restrict__ Gp, int np, int nb, int ng,
                                                           PanelGroups grp, int sweep, const int* __restrict__ cnt,
                                                           float* __restrict__ Sg, float* __restrict__ Pv) {
  const int g = blockIdx.x, b = blockIdx.y;
  if (sweep > 0 && cnt[gridDim.y * JMAX_SWEEPS + sweep] == 0) return;
  if (sweep > 0 && cnt[b * JMAX_SWEEPS + sweep - 1] == 0) return;
  int x = g;
  x = ((x >> grp.plo) << (grp.plo + 1)) | (x & ((1 << grp.plo) - 1));
  x = ((x >> grp.phi) << (grp.phi + 1)) | (x & ((1 << grp.phi) - 1));
  int blk[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) blk[j] = x ^ ((j & 1) ? grp.ga : 0) ^ ((j & 2) ? grp.gb : 0);
  const float4* src = reinterpret_cast<const float4*>(Gp + int64_t(b) * np * np);
  float4* dst = reinterpret_cast<float4*>(Sg + (int64_t(b) * ng + g) * 128 * 128);
  float4* pv = reinterpret_cast<float4*>(Pv + (int64_t(b) * ng + g) * 128 * 128);
  // local element (r, c) sits at ((c >> 5) * 128 + r) * 32 + (c & 31); as float4 index e = (cb * 128 + r) * 8 + j4
  for (int e = threadIdx.x; e < 128 * 128 / 4; e += 256) {
    const int j4 = e & 7, r = (e >> 3) & 127, cb = e >> 10;
    const int64_t s = (int64_t(blk[cb]) * np + blk[r >> 5] * JB + (r & 31)) * 8 + j4;
    dst[e] = src[s];
    const int d = r - cb * 32 - j4 * 4;                 // diagonal position inside this float4, if 0..3
    pv[e] = make_float4(d == 0 ? 1.f : 0.f, d == 1 ? 1.f : 0.f, d == 2 ? 1.f : 0.f, d == 3 ? 1.f : 0.f);
  }
}

// Pt[b*ng + g][n][k] = P[k][n] (P = the local V, blocked layout) -- the K-major B operand of the K = 128 panel update;
// gflag = OR of the six local task flags of the super-round.  One CTA per (group, matrix).
__global__ void __launch_bounds__(256) group_transpose_kernel(const float* __restrict__ Pv, int ng, int sweep,
                                                              const int* __restrict__ cnt,
                                                              const int* __restrict__ f0, const int* __restrict__ f1,
                                                              const int* __restrict__ f2, float* __restrict__ Pt,
                                                              int* __restrict__ gflag) {
  const int g = blockIdx.x, b = blockIdx.y;
  if (sweep > 0 && cnt[gridDim.y * JMAX_SWEEPS + sweep] == 0) return;
  if (sweep > 0 && cnt[b * JMAX_SWEEPS + sweep - 1] == 0) return;
  __shared__ float tile[32][33];
  const int64_t gi = int64_t(b) * ng + g;
  const float* pv = Pv + gi * 128 * 128;
  float* pt = Pt + gi * 128 * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = 0; t < 16; ++t) {                       // 32 x 32 tile (k block kb, column block cb)
    const int cb = t >> 2, kb = t & 3;
#pragma unroll
    for (int r = warp; r < 32; r += 8) tile[r][lane] = pv[(cb * 128 + kb * 32 + r) * 32 + lane];   // [k][n]
    __syncthreads();
#pragma unroll
    for (int r = warp; r < 32; r += 8) pt[(cb * 32 + r) * 128 + kb * 32 + lane] = tile[lane][r];     // [n][k]
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int i = int(gi) * 2;
    gflag[gi] = f0[i] | f0[i + 1] | f1[i] | f1[i + 1] | f2[i] | f2[i + 1];
  }
}

// lambda = diag(Gp); Ut (n x n compact) = rows of Vt[:n, :n] renormalised to unit length (fp32 rotation
// products drift from orthonormal by ~1e-4 over a few thousand rotations); one warp per row.
// sweeps = number of sweeps executed (the last one found nothing significant to rotate).
__global__ void __launch_bounds__(256) jacobi_extract_kernel(const float* __restrict__ Gp,
                                                             const float* __restrict__ Vt, int n, int np,
                                                             const int* __restrict__ cnt, int max_sweeps,
                                                             float* __restrict__ lambda, float* __restrict__ Ut,
                                                             int* __restrict__ sweeps, int v_is_columns) {
  const int b = blockIdx.y;
  const float* gp = Gp + int64_t(b) * np * np;
  const float* vt = Vt + int64_t(b) * np * np;
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += gridDim.x * 8) {
    // eigenvector i is row i of Vt (SIMT update) or column i of V (tensor-core update); blocked layout
    float ss = 0.f;
    for (int j = lane; j < n; j += 32) {
      const float v = v_is_columns ? vt[boff(np, j, i)] : vt[boff(np, i, j)];
      ss = fmaf(v, v, ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = ss > 0.f ? rsqrtf(ss) : 0.f;
    if (Ut) {
      float* out = Ut + int64_t(b) * n * n + int64_t(i) * n;
      for (int j = lane; j < n; j += 32) out[j] = (v_is_columns ? vt[boff(np, j, i)] : vt[boff(np, i, j)]) * inv;
    }
    if (lambda && lane == 0) lambda[int64_t(b) * n + i] = gp[boff(np, i, i)];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && sweeps) {
    int s = 0;
    while (s < max_sweeps && cnt[b * JMAX_SWEEPS + s] != 0) ++s;
    sweeps[b] = s < max_sweeps ? s + 1 : -max_sweeps;    // negative: every sweep up to the cap still rotated (not converged)
  }
}

// Same outputs for the tensor-core path, where eigenvector i is COLUMN i of V (blocked layout: a 32-column block
// is np contiguous 128-byte rows).  One CTA per (column block, matrix): column norms from coalesced row reads,
// then 32x32 tiles transposed through shared memory so that both the reads and the writes of U^T are full lines
// (the row-per-warp kernel above reads one float per 128-byte line in this layout: 258 -> ~50 us per launch).
__global__ void __launch_bounds__(256) jacobi_extract_cols_kernel(const float* __restrict__ Gp,
                                                                  const float* __restrict__ V, int n, int np,
                                                                  const int* __restrict__ cnt, int max_sweeps,
                                                                  float* __restrict__ lambda, float* __restrict__ Ut,
                                                                  int* __restrict__ sweeps) {
  __shared__ float red[8][33];
  __shared__ float inv[32];
  __shared__ float tile[32][33];
  const int b = blockIdx.y, ib = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* vb = V + int64_t(b) * np * np + int64_t(ib) * np * JB;   // rows j = 0..np-1 of this column block, 32 floats each
  float ss = 0.f;
  for (int j = warp; j < n; j += 8) {
    const float v = vb[int64_t(j) * JB + lane];
    ss = fmaf(v, v, ss);
  }
  red[warp][lane] = ss;
  __syncthreads();
  if (warp == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][lane];
    inv[lane] = t > 0.f ? rsqrtf(t) : 0.f;
    const int i = ib * JB + lane;
    if (lambda && i < n) lambda[int64_t(b) * n + i] = Gp[int64_t(b) * np * np + boff(np, i, i)];
  }
  __syncthreads();
  if (Ut) {
    float* ub = Ut + int64_t(b) * n * n;
    for (int j0 = 0; j0 < n; j0 += 32) {
#pragma unroll
      for (int r = warp; r < 32; r += 8) tile[r][lane] = (j0 + r < n) ? vb[int64_t(j0 + r) * JB + lane] : 0.f;
      __syncthreads();
#pragma unroll
      for (int c = warp; c < 32; c += 8) {
        const int i = ib * JB + c, j = j0 + lane;
        if (i < n && j < n) ub[int64_t(i) * n + j] = tile[lane][c] * inv[c];
      }
      __syncthreads();
    }
  }
  if (ib == 0 && threadIdx.x == 0 && sweeps) {
    int s = 0;
    while (s < max_sweeps && cnt[b * JMAX_SWEEPS + s] != 0) ++s;
    sweeps[b] = s < max_sweeps ? s + 1 : -max_sweeps;    // negative: every sweep up to the cap still rotated (not converged)
  }
}

// ------------------------------------------------------------------------------
// sigma_j = ||Y[j,:]|| / ||Ut[j,:]||   (one warp per row)
// ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) row_sigma_kernel(const float* __restrict__ Y, const float* __restrict__ Ut,
                                                        int64_t rows_total, int n, int m,
                                                        float* __restrict__ sigma) {
  const int64_t row = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows_total) return;
  const int lane = threadIdx.x & 31;
  const float* y = Y + row * m;
  const float* u = Ut + row * n;
  float sy = 0.f, su = 0.f;
  for (int c = lane; c < m; c += 32) { const float v = y[c]; sy = fmaf(v, v, sy); }
  for (int c = lane; c < n; c += 32) { const float v = u[c]; su = fmaf(v, v, su); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
    su += __shfl_xor_sync(0xffffffffu, su, o);
  }
  if (lane == 0) sigma[row] = (su > 0.f) ? sqrtf(sy / su) : 0.f;
}

// two-pass solver: total sweeps of both passes; negative when the FINAL pass hit its cap (the first pass may hit its
// cap by design -- whatever it leaves, the second pass removes)
__global__ void sweeps_merge_kernel(int32_t* __restrict__ sweeps, const int32_t* __restrict__ sweeps2, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int s1 = abs(sweeps[b]), s2 = sweeps2[b];
  sweeps[b] = s2 < 0 ? -(s1 - s2) : (s1 + s2);
}

// block-wide reductions for the per-sample spectrum kernels (blockDim = 256)
__device__ __forceinline__ float block_reduce(float v, bool is_max, float* sh) {
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

struct Spectrum { float smax, S, H, erank; };

// cut-off / normalise / entropy / exp for one sample (all 256 threads participate)
__device__ __forceinline__ Spectrum spectrum_stats(const float* __restrict__ sg, int n, float rtol, float* sh) {
  float m = 0.f;
  for (int j = threadIdx.x; j < n; j += 256) m = fmaxf(m, sg[j]);
  const float smax = block_reduce(m, true, sh);
  const float cut = rtol * smax;
  float s = 0.f;
  for (int j = threadIdx.x; j < n; j += 256) { const float v = sg[j]; if (v > cut) s += v; }
  const float S = block_reduce(s, false, sh);
  float h = 0.f;
  if (S > 0.f) {
    for (int j = threadIdx.x; j < n; j += 256) {
      const float v = sg[j];
      if (v > cut) { const float p = v / S; h -= p * logf(p); }
    }
  }
  const float H = block_reduce(h, false, sh);
  Spectrum r; r.smax = smax; r.S = S; r.H = H; r.erank = (S > 0.f) ? expf(H) : 0.f;
  return r;
}

__global__ void __launch_bounds__(256) erank_entropy_kernel(const float* __restrict__ sigma, int n, float rtol,
                                                            float* __restrict__ erank) {
  __shared__ float sh[8];
  const Spectrum sp = spectrum_stats(sigma + int64_t(blockIdx.x) * n, n, rtol, sh);
  if (threadIdx.x == 0) erank[blockIdx.x] = sp.erank;
}

// coef_j = g * erank * (-(ln p_j + H) / S) / sigma_j  for kept j, else 0
__global__ void __launch_bounds__(256) erank_coef_kernel(const float* __restrict__ sigma, const float* __restrict__ g,
                                                         int n, float rtol, float* __restrict__ coef) {
  __shared__ float sh[8];
  const float* sg = sigma + int64_t(blockIdx.x) * n;
  const Spectrum sp = spectrum_stats(sg, n, rtol, sh);
  const float cut = rtol * sp.smax, gb = g[blockIdx.x];
  for (int j = threadIdx.x; j < n; j += 256) {
    const float v = sg[j];
    float cf = 0.f;
    if (v > cut && sp.S > 0.f) {
      const float p = v / sp.S;
      cf = gb * sp.erank * (-(logf(p) + sp.H) / sp.S) / v;
    }
    coef[int64_t(blockIdx.x) * n + j] = cf;
  }
}

// a13: s_t = sum_j p_j Ut[j][t]^2 / ||Ut[j]||^2
__global__ void __launch_bounds__(256) token_info_kernel(const float* __restrict__ sigma, const float* __restrict__ Ut,
                                                         int n, float rtol, float* __restrict__ out) {
  __shared__ float sh[8];
  extern __shared__ float pj[];
  const float* sg = sigma + int64_t(blockIdx.x) * n;
  const float* u = Ut + int64_t(blockIdx.x) * n * n;
  const Spectrum sp = spectrum_stats(sg, n, rtol, sh);
  const float cut = rtol * sp.smax;
  for (int j = threadIdx.x; j < n; j += 256) pj[j] = (sg[j] > cut && sp.S > 0.f) ? sg[j] / sp.S : 0.f;
  __syncthreads();
  // normalise p_j by the squared row norm of Ut (rows are unit up to fp32 drift)
  for (int j = threadIdx.x >> 5; j < n; j += 8) {
    float su = 0.f;
    for (int c = threadIdx.x & 31; c < n; c += 32) { const float v = u[int64_t(j) * n + c]; su = fmaf(v, v, su); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) su += __shfl_xor_sync(0xffffffffu, su, o);
    if ((threadIdx.x & 31) == 0 && su > 0.f) pj[j] /= su;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < n; t += 256) {
    float s = 0.f;
    for (int j = 0; j < n; ++j) { const float v = u[int64_t(j) * n + t]; s = fmaf(pj[j] * v, v, s); }
    out[int64_t(blockIdx.x) * n + t] = s;
  }
}

int gram_tcgen05_launch(const void* x, int64_t B, int64_t T, int64_t C, int dtype, void* workspace, float* G,
                        cudaStream_t st);   // gram_tcgen05.cu
bool gram_tcgen05_supported(int64_t B, int64_t T, int64_t C, int dtype);

}  // namespace r3d

using namespace r3d;

// ================================================================================
// C ABI
// ================================================================================
static inline void side(int64_t T, int64_t C, int64_t& n, int64_t& m, bool& tside) {
  tside = T < C;
  n = tside ? T : C;
  m = tside ? C : T;
}

// fp32 Q^T (tasks, 64, 64) -> the pre-split layout (tasks, 2, 64, 64) the inner solver emits (test hooks only)
__global__ void split_q_kernel(const float* __restrict__ q, float* __restrict__ out, int64_t total) {
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
    const int64_t task = e / (JM * JM), r = e % (JM * JM);
    const float x = q[e], h = q_hi_of(x);
    out[task * QSTR + r] = h;
    out[task * QSTR + JM * JM + r] = x - h;
  }
}
static int split_q(const float* q, int64_t tasks, float** out, cudaStream_t st) {
  R3D_CUDA(cudaMallocAsync((void**)out, size_t(tasks) * QSTR * sizeof(float), st));
  split_q_kernel<<<(unsigned)std::min<int64_t>((tasks * JM * JM + 255) / 256, 4096), 256, 0, st>>>(q, *out, tasks * JM * JM);
  R3D_LAUNCH_CHECK();
  return 0;
}

// Debug/test hook: one tensor-core panel-update round on caller-provided buffers (all (B, np, np) fp32,
// Qb (B, np/64, 64, 64)); every task is applied (qflag = 1, nothing converged).
extern "C" int r3d_debug_panel_round(float* G, float* H, float* V, const float* Qb, int64_t B, int np, int round,
                                     int* scratch /* B*32 + B*np/64 ints */, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  R3D_CHECK(panel_tc_supported(np), "np must be a multiple of 128");
  const int nt = np / JM;
  int* cnt = scratch;
  int* qflag = scratch + B * JMAX_SWEEPS;
  R3D_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * B * JMAX_SWEEPS, st));
  std::vector<int> ones(B * nt, 1);
  R3D_CUDA(cudaMemcpyAsync(qflag, ones.data(), sizeof(int) * B * nt, cudaMemcpyHostToDevice, st));
  R3D_CUDA(cudaStreamSynchronize(st));
  float* qs = nullptr;
  if (int e = split_q(Qb, B * nt, &qs, st)) return e;
  PanelTc ptc;
  int rc = panel_tc_prepare(&ptc, G, H, V, qs, qs, B, np);
  if (!rc) rc = panel_tc_update_v(&ptc, 0, round, 0, cnt, qflag, st);
  if (!rc) {
    if (options().panel_sym != 0 && panel_sym_supported(np)) {     // one in-place pass; H is not touched
      rc = panel_sym_prepare(&ptc);
      if (!rc) rc = panel_sym_update_g(&ptc, 0, round, 0, cnt, qflag, st);
    } else {
      rc = panel_tc_update_g(&ptc, 0, round, 0, cnt, qflag, st);
    }
  }
  cudaFreeAsync(qs, st);
  return rc;
}

// Debug/test hook of the chained V update: V <- V Q1 Q2 Q3 for the XOR rounds with masks ga, gb, ga ^ gb.  Q3 holds the
// three rounds' Q^T buffers back to back (3 x (B, np/64, 64, 64)); scratch: B * 32 + 3 * B * np/64 ints.
extern "C" int r3d_debug_vchain(float* V, float* Q3, int64_t B, int np, int ga, int gb, int* scratch, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  R3D_CHECK(panel_chain_supported(np), "np / 32 must be a power of two >= 8");
  R3D_CHECK(ga > 0 && gb > 0 && ga != gb && ga < np / JB && gb < np / JB, "bad masks");
  const int nt = np / JM;
  int* cnt = scratch;
  int* qflag = scratch + B * JMAX_SWEEPS;
  R3D_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * B * JMAX_SWEEPS, st));
  std::vector<int> ones(3 * B * nt, 1);
  R3D_CUDA(cudaMemcpyAsync(qflag, ones.data(), sizeof(int) * 3 * B * nt, cudaMemcpyHostToDevice, st));
  R3D_CUDA(cudaStreamSynchronize(st));
  float* qs = nullptr;
  if (int e = split_q(Q3, 3 * B * nt, &qs, st)) return e;
  PanelTc ptc;
  int rc = panel_tc_prepare(&ptc, V, nullptr, V, qs, qs, B, np);
  const size_t qsz = size_t(B) * nt * QSTR;
  float* qc[6] = {qs, qs + qsz, qs + 2 * qsz, qs, qs + qsz, qs + 2 * qsz};
  if (!rc) rc = panel_tc_prepare_chain(&ptc, qc);
  const int* fl[3] = {qflag, qflag + B * nt, qflag + 2 * B * nt};
  if (!rc) rc = panel_tc_update_v_chain(&ptc, 0, groups_of(SuperRound{ga, gb}), 0, cnt, fl, st);
  cudaFreeAsync(qs, st);
  return rc;
}

// Host-only test hooks (no GPU work): the round plan of a sweep for nb blocks -- entries (a, b): b == 0 a single XOR round
// with mask a, else the super-round {a, b, a ^ b}; returns the number of entries (0: the block count gets the circle
// method) -- and the chained V update's tile bookkeeping for group g of the super-round (ga, gb).
extern "C" int r3d_debug_round_plan(int nb, int32_t* out_pairs, int cap) {
  const std::vector<SuperRound> plan = spread_plan(nb);
  for (size_t i = 0; i < plan.size() && (int)i < cap; ++i) { out_pairs[2 * i] = plan[i].a; out_pairs[2 * i + 1] = plan[i].b; }
  return (int)plan.size();
}
extern "C" int r3d_debug_chain_plan(int ga, int gb, int g, int32_t* out22) {
  R3D_CHECK(ga > 0 && gb > 0 && ga != gb && out22 != nullptr, "bad arguments");
  int tmp[22];
  panel_chain_plan_host(groups_of(SuperRound{ga, gb}), g, tmp);
  for (int i = 0; i < 22; ++i) out22[i] = tmp[i];
  return 0;
}

extern "C" int r3d_panel_tiles(uint64_t* out3, int reset) {
  R3D_CHECK(out3 != nullptr, "null pointer");
  unsigned long long v[3];
  if (int e = panel_tiles_read(v, reset)) return e;
  unsigned long long units = 0;                     // the one-pass symmetric G update counts in the same 64 KB units
  if (int e = panel_sym_units_read(&units, reset)) return e;
  out3[0] = v[0] + units; out3[1] = v[1]; out3[2] = v[2];
  return 0;
}


extern "C" int r3d_set_option(const char* key, double value) {
  R3D_CHECK(key != nullptr, "null option key");
  const std::string k(key);
  if (k == "jacobi_update_tc") options().jacobi_update_tc = value != 0.0;
  else if (k == "jacobi_tol") options().jacobi_tol = (float)value;
  else if (k == "jacobi_max_sweeps") options().jacobi_max_sweeps = (int)value;
  else if (k == "jacobi_overlap_v") options().jacobi_overlap_v = value != 0.0;
  else if (k == "jacobi_chunks") options().jacobi_chunks = (int)value;
  else if (k == "gemm_tc") options().gemm_tc = value != 0.0;
  else if (k == "erank_passes") options().erank_passes = (int)value;
  else if (k == "erank_pass2_sweeps") options().erank_pass2_sweeps = (int)value;
  else if (k == "erank_pass1_sweeps") options().erank_pass1_sweeps = (int)value;
  else if (k == "jacobi_tol_pass1") options().jacobi_tol_pass1 = (float)value;
  else if (k == "jacobi_nu_pass1") options().jacobi_nu_pass1 = (float)value;
  else if (k == "jacobi_nu_pass2") options().jacobi_nu_pass2 = (float)value;
  else if (k == "jacobi_inner_regs") options().jacobi_inner_regs = (int)value;
  else if (k == "panel_merged") options().panel_merged = (int)value;
  else if (k == "row_chunk_mult") g_row_chunk_mult = std::max(1, std::min(16, (int)value));
  else if (k == "panel_group_mb") options().panel_group_mb = std::max(1, (int)value);
  else if (k == "panel_ring") options().panel_ring = (int)value;
  else if (k == "jacobi_v_after_g") options().jacobi_v_after_g = value != 0.0;
  else if (k == "jacobi_schedule") options().jacobi_schedule = (int)value;
  else if (k == "jacobi_own_streams") options().jacobi_own_streams = (int)value;
  else if (k == "panel_sym") options().panel_sym = (int)value;
  else if (k == "lin_fast") options().lin_fast = (int)value;
  else if (k == "panel_debug") g_panel_debug = (int)value;
  else if (k == "panel_grid_cap") g_panel_grid_cap = (int)value;
  else R3D_CHECK(false, "unknown option '%s'", key);
  return 0;
}

extern "C" size_t r3d_jacobi_workspace_bytes(int64_t B, int64_t n) { return jacobi_ws_bytes(B, n); }

// Workspace layout: [G (B,n,n) f32][coef (B,n) f32][Jacobi workspace][bf16 planes: U (3,B,n,n) | Y (3,B,n,m) |
// X (3,B,T,C), the last only for fp32 inputs]
struct ErankWs {
  float* G; float* coef; void* jws; __nv_bfloat16* Upl; __nv_bfloat16* Ypl; __nv_bfloat16* Xpl;
  float* U2; __nv_bfloat16* U2pl;     // second refinement pass only
};
static size_t align256(size_t v) { return (v + 255) & ~size_t(255); }
static ErankWs erank_carve(void* workspace, int64_t B, int64_t T, int64_t C, int dtype) {
  int64_t n, m; bool ts; side(T, C, n, m, ts);
  ErankWs w;
  char* p = (char*)((uintptr_t(workspace) + 255) & ~uintptr_t(255));
  w.G = (float*)p; p += align256(size_t(B) * n * n * 4);
  w.coef = (float*)p; p += align256(size_t(B) * n * 4);
  w.jws = p; p += align256(jacobi_ws_bytes(B, n));
  w.Upl = (__nv_bfloat16*)p; p += align256(size_t(3) * B * n * n * 2);
  w.Ypl = (__nv_bfloat16*)p; p += align256(size_t(3) * B * n * m * 2);
  w.Xpl = (dtype == R3D_F32) ? (__nv_bfloat16*)p : nullptr;
  if (dtype == R3D_F32) p += align256(size_t(3) * B * T * C * 2);
  w.U2 = (float*)p; p += align256(size_t(B) * n * n * 4);
  w.U2pl = (__nv_bfloat16*)p;
  return w;
}

extern "C" size_t r3d_erank_workspace_bytes(int64_t B, int64_t T, int64_t C, int dtype) {
  int64_t n, m; bool ts; side(T, C, n, m, ts);
  size_t bytes = align256(size_t(B) * n * n * 4) + align256(size_t(B) * n * 4) + align256(jacobi_ws_bytes(B, n)) +
                 align256(size_t(3) * B * n * n * 2) + align256(size_t(3) * B * n * m * 2) + 1024;
  if (dtype == R3D_F32) bytes += align256(size_t(3) * B * T * C * 2);
  bytes += align256(size_t(B) * n * n * 4) + align256(size_t(3) * B * n * n * 2);   // second refinement pass
  return bytes;
}

// ---- tensor-core (bf16-plane) versions of the three GEMMs; return -1 when the shape needs the SIMT fallback ----
static bool tc_gemm_ok(int64_t T, int64_t C) {
  return options().gemm_tc != 0 && T % 8 == 0 && C % 8 == 0;
}

// Y (B,n,m) f32 = Ut * A   (A = the (n, m) short-side-major view of x)
static int refine_Y_tc(const void* x, int dtype, const float* Ut, const ErankWs& w, int64_t B, int64_t T, int64_t C,
                       float* Y, cudaStream_t st) {
  int64_t n, m; bool ts; side(T, C, n, m, ts);
  R3D_STAGE(ST_REFINE_Y, st);
  if (int e = split_planes(Ut, R3D_F32, w.Upl, B * n * n, 3, n, nullptr, st)) return e;
  const void* xb = x;
  int pb = 1;
  if (dtype == R3D_F32) {
    if (int e = split_planes(x, R3D_F32, w.Xpl, B * T * C, 3, C, nullptr, st)) return e;
    xb = w.Xpl; pb = 3;
  }
  PGemm g{};
  g.A = w.Upl; g.B = xb; g.batch = (int)B; g.pa = 3; g.pb = pb;
  g.a_kmajor = 1;                   // Ut[j][k]
  g.b_kmajor = ts ? 0 : 1;          // token side: X[r (K)][c (N)] is MN-major; channel side: X[t (N)][c (K)] is K-major
  g.M = (int)n; g.N = (int)m; g.K = (int)n;
  pgemm_products(g, 3, pb, true);
  g.out_mode = 0; g.C = Y; g.ldc = m; g.strideC = n * m;
  return pgemm_launch(g, st);
}

// dX (B,T,C) (+)= U diag(coef) Y
static int bwd_gemm_tc(const float* Ut, const float* Y, const float* coef, const ErankWs& w, int64_t B, int64_t T,
                       int64_t C, int dtype, void* dx, int accumulate, cudaStream_t st) {
  int64_t n, m; bool ts; side(T, C, n, m, ts);
  R3D_STAGE(ST_BWD_GEMM, st);
  const int P = dtype == R3D_F32 ? 3 : 2;
  if (int e = split_planes(Ut, R3D_F32, w.Upl, B * n * n, P, n, nullptr, st)) return e;
  if (int e = split_planes(Y, R3D_F32, w.Ypl, B * n * m, P, m, coef, st)) return e;     // rows scaled by coef_j
  PGemm g{};
  g.batch = (int)B; g.pa = P; g.pb = P; g.a_kmajor = 0; g.b_kmajor = 0;   // both operands are [rows j = K][cols]
  if (ts) { g.A = w.Upl; g.B = w.Ypl; }      // dX[t][c] = sum_j Ut[j][t] * (coef Y)[j][c]
  else    { g.A = w.Ypl; g.B = w.Upl; }      // dX[t][c] = sum_j (coef Y)[j][t] * Ut[j][c]
  g.M = (int)T; g.N = (int)C; g.K = (int)n;
  pgemm_products(g, P, P, dtype == R3D_F32);
  g.out_mode = (dtype == R3D_F32 ? 0 : 2) + (accumulate ? 1 : 0);
  g.C = dx; g.ldc = C; g.strideC = T * C;
  return pgemm_launch(g, st);
}

static int jacobi_run_chunked(const float* G, int64_t B, int64_t n, void* workspace, float* lambda_out, float* U_out,
                              int32_t* sweeps_out, int max_sweeps, cudaStream_t st, float tol_override = 0.f);

// Optional second refinement pass.  After the first pass the rows of Y = U^T A are orthogonal only up to the
// absolute error of an fp32 Gram (eps * lambda_max), which is large relative to the smallest singular directions.
// G2 = Y Y^T is nearly diagonal and GRADED (its small entries are represented to fp32 relative accuracy), which
// two-sided Jacobi resolves to relative accuracy (Demmel-Veselic): G2 = V2^T diag V2, then U <- V2 U, Y <- V2 Y.
// sweep cap of the second pass: the option, or (negative = automatic, the default) 6 up to n = 512 and 8 beyond --
// with the first-pass floor at 8192 ulps the second pass takes 3-5 sweeps at n <= 512 and 7 at n = 2048 (decay spectrum)
static int pass2_cap(int64_t n) {
  const int o = options().erank_pass2_sweeps;
  return o >= 0 ? o : (n <= 512 ? 6 : 8);
}

static int second_pass_tc(const ErankWs& w, int64_t B, int64_t n, int64_t m, float* U, float* Y, int32_t* sweeps2,
                          cudaStream_t st) {
  if (int e = split_planes(Y, R3D_F32, w.Ypl, B * n * m, 3, m, nullptr, st)) return e;
  {
    R3D_STAGE(ST_GRAM, st);
    PGemm g{};
    g.A = w.Ypl; g.B = w.Ypl; g.batch = (int)B; g.pa = 3; g.pb = 3; g.a_kmajor = 1; g.b_kmajor = 1;
    g.M = (int)n; g.N = (int)n; g.K = (int)m;
    pgemm_products(g, 3, 3, true);
    g.out_mode = 0; g.C = w.G; g.ldc = n; g.strideC = n * n;
    if (int e = pgemm_launch(g, st)) return e;
  }
  if (int e = jacobi_run_chunked(w.G, B, n, w.jws, nullptr, w.U2, sweeps2, pass2_cap(n), st, -1.f)) return e;
  R3D_STAGE(ST_REFINE_Y, st);
  if (int e = split_planes(w.U2, R3D_F32, w.U2pl, B * n * n, 3, n, nullptr, st)) return e;
  if (int e = split_planes(U, R3D_F32, w.Upl, B * n * n, 3, n, nullptr, st)) return e;
  PGemm g{};
  g.A = w.U2pl; g.batch = (int)B; g.pa = 3; g.pb = 3; g.a_kmajor = 1; g.b_kmajor = 0;   // B operand: [rows j = K][cols]
  g.M = (int)n; g.K = (int)n;
  pgemm_products(g, 3, 3, true);
  g.out_mode = 0;
  g.B = w.Upl; g.N = (int)n; g.C = U; g.ldc = n; g.strideC = n * n;          // U <- V2 U
  if (int e = pgemm_launch(g, st)) return e;
  g.B = w.Ypl; g.N = (int)m; g.C = Y; g.ldc = m; g.strideC = n * m;          // Y <- V2 Y
  return pgemm_launch(g, st);
}

// The same pass with the SIMT GEMMs (shapes the tensor-core path does not take: T or C not a multiple of 8,
// unaligned x).  The plane buffers are free at this point and serve as the fp32 temporaries.
static int second_pass_simt(const ErankWs& w, int64_t B, int64_t n, int64_t m, float* U, float* Y, int32_t* sweeps2,
                            cudaStream_t st) {
  {
    R3D_STAGE(ST_GRAM, st);
    if (int e = sgemm_launch<float, float, float>(false, true, Y, Y, w.G, int(n), int(n), int(m), m, m, n, n * m, n * m,
                                                  n * n, nullptr, 0, 0, int(B), st)) return e;
  }
  if (int e = jacobi_run_chunked(w.G, B, n, w.jws, nullptr, w.U2, sweeps2, pass2_cap(n), st, -1.f)) return e;
  R3D_STAGE(ST_REFINE_Y, st);
  float* Un = reinterpret_cast<float*>(w.Upl);
  float* Yn = reinterpret_cast<float*>(w.Ypl);
  if (int e = sgemm_launch<float, float, float>(false, false, w.U2, U, Un, int(n), int(n), int(n), n, n, n, n * n, n * n,
                                                n * n, nullptr, 0, 0, int(B), st)) return e;
  if (int e = sgemm_launch<float, float, float>(false, false, w.U2, Y, Yn, int(n), int(m), int(n), n, m, m, n * n, n * m,
                                                n * m, nullptr, 0, 0, int(B), st)) return e;
  R3D_CUDA(cudaMemcpyAsync(U, Un, size_t(B) * n * n * 4, cudaMemcpyDeviceToDevice, st));
  R3D_CUDA(cudaMemcpyAsync(Y, Yn, size_t(B) * n * m * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// Gram of an fp32 input on tensor cores: 6 products of the 3 bf16 planes
static int gram_f32_tc(const void* x, const ErankWs& w, int64_t B, int64_t T, int64_t C, float* G, cudaStream_t st) {
  int64_t n, m; bool ts; side(T, C, n, m, ts);
  R3D_STAGE(ST_GRAM, st);
  if (int e = split_planes(x, R3D_F32, w.Xpl, B * T * C, 3, C, nullptr, st)) return e;
  PGemm g{};
  g.A = w.Xpl; g.B = w.Xpl; g.batch = (int)B; g.pa = 3; g.pb = 3;
  g.a_kmajor = ts ? 1 : 0; g.b_kmajor = ts ? 1 : 0;      // token side: X[t][c = K]; channel side: X[t = K][c]
  g.M = (int)n; g.N = (int)n; g.K = (int)m;
  pgemm_products(g, 3, 3, true);
  g.out_mode = 0; g.C = G; g.ldc = n; g.strideC = n * n;
  return pgemm_launch(g, st);
}

template <typename T>
static int gram_simt(const void* x, int64_t B, int64_t Tt, int64_t C, float* G, cudaStream_t st) {
  int64_t n, m; bool ts; side(Tt, C, n, m, ts);
  const T* X = (const T*)x;
  // T-side: G = X X^T  (a(i,k)=X[i*C+k], b(k,j)=X[j*C+k]);  C-side: G = X^T X (a(i,k)=X[k*C+i], b(k,j)=X[k*C+j])
  R3D_STAGE(ST_GRAM, st);
  return sgemm_launch<T, T, float>(!ts, ts, X, X, G, int(n), int(n), int(m), C, C, n, Tt * C, Tt * C, n * n, nullptr,
                                   0, 0, int(B), st);
}

extern "C" int r3d_gram(const void* x, int64_t B, int64_t T, int64_t C, int dtype, int gram_impl, void* workspace,
                        float* G_out, void* stream) {
  R3D_CHECK(x && G_out, "null pointer");
  R3D_CHECK(B >= 1 && T >= 1 && C >= 1, "bad shape");
  R3D_CHECK(dtype == R3D_F32 || dtype == R3D_BF16, "bad dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  if (gram_impl == 0 && gram_tcgen05_supported(B, T, C, dtype))
    return gram_tcgen05_launch(x, B, T, C, dtype, workspace, G_out, st);
  if (gram_impl == 0 && dtype == R3D_F32 && workspace && tc_gemm_ok(T, C) &&
      (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const ErankWs w = erank_carve(workspace, B, T, C, dtype);
    return gram_f32_tc(x, w, B, T, C, G_out, st);
  }
  return dtype == R3D_F32 ? gram_simt<float>(x, B, T, C, G_out, st) : gram_simt<__nv_bfloat16>(x, B, T, C, G_out, st);
}

// Library-owned streams.  (1) Per chunk, the V <- V Q update of round r runs on a side stream and overlaps the
// inner solve of round r+1 (which needs only G and leaves HBM idle).  (2) The batch is split into two chunks
// whose Jacobi iterations run on two streams, so one chunk's issue-bound inner solve overlaps the other's
// HBM-bound panel passes.  Fork/join with events only, so the pattern is also legal under stream capture.
constexpr int kMaxChunks = 4;
static_assert(kMaxChunks == kMaxChunksWs, "workspace slack is sized for kMaxChunks carvings");
struct StreamSet {
  bool ready = false;
  cudaStream_t chunk[kMaxChunks] = {};     // chunk 0 uses the caller's stream
  cudaStream_t vst[kMaxChunks] = {};
  cudaEvent_t ev_inner[kMaxChunks][2], ev_v[kMaxChunks][2], ev_fork, ev_join[kMaxChunks];
  cudaEvent_t ev_inner_c[kMaxChunks][2], ev_v_c[kMaxChunks][2];   // chained schedule: per slot
  int create() {
    for (int c = 0; c < kMaxChunks; ++c) {
      // The chunk streams carry the critical path (inner solve -> G update -> inner solve ...); the V streams only have
      // to be done two super-rounds later.  Higher priority for the former lets their CTAs take free SM slots first.
      int lo = 0, hi = 0;
      R3D_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      R3D_CUDA(cudaStreamCreateWithPriority(&chunk[c], cudaStreamNonBlocking, hi));
      R3D_CUDA(cudaStreamCreateWithPriority(&vst[c], cudaStreamNonBlocking, lo));
      R3D_CUDA(cudaEventCreateWithFlags(&ev_join[c], cudaEventDisableTiming));
      for (int i = 0; i < 2; ++i) {
        R3D_CUDA(cudaEventCreateWithFlags(&ev_inner[c][i], cudaEventDisableTiming));
        R3D_CUDA(cudaEventCreateWithFlags(&ev_v[c][i], cudaEventDisableTiming));
        R3D_CUDA(cudaEventCreateWithFlags(&ev_inner_c[c][i], cudaEventDisableTiming));
        R3D_CUDA(cudaEventCreateWithFlags(&ev_v_c[c][i], cudaEventDisableTiming));
      }
    }
    R3D_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    ready = true;
    return 0;
  }
};
// Stream sets are pooled per device for the whole process: a host thread borrows one on first use and hands it back
// when it exits (nn.DataParallel starts fresh threads for every forward, so thread-owned sets would be re-created
// -- or, without a destructor, leaked -- on every step).  A set is never shared by two live threads: events recorded
// by one thread must not be overwritten by another.
static std::mutex g_pool_mu;
static std::vector<StreamSet*> g_pool_free[kMaxDevices];
static std::atomic<int> g_sets_created{0};
struct StreamLease {
  StreamSet* set[kMaxDevices] = {};
  ~StreamLease() {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (int d = 0; d < kMaxDevices; ++d)
      if (set[d]) g_pool_free[d].push_back(set[d]);
  }
  int get(StreamSet** out) {
    const int d = current_device_index();
    if (!set[d]) {
      {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (!g_pool_free[d].empty()) { set[d] = g_pool_free[d].back(); g_pool_free[d].pop_back(); }
      }
      if (!set[d]) {
        StreamSet* s = new StreamSet();
        if (int e = s->create()) { delete s; return e; }
        g_sets_created.fetch_add(1);
        set[d] = s;
      }
    }
    *out = set[d];
    return 0;
  }
};
static thread_local StreamLease g_lease;
struct StreamRef {                         // keeps the old `g_streams.x` / `g_streams.ensure()` spelling
  StreamSet* s = nullptr;
  int ensure() { return s ? 0 : g_lease.get(&s); }
  StreamSet* operator->() { return s; }
};

static int jacobi_run(const float* G, int64_t B, int64_t n, void* workspace, float* lambda_out, float* U_out,
                      int32_t* sweeps_out, int max_sweeps, cudaStream_t st, int chunk = 0, float tol_override = 0.f) {
  if (max_sweeps <= 0 || max_sweeps > JMAX_SWEEPS) max_sweeps = options().jacobi_max_sweeps;
  if (max_sweeps <= 0 || max_sweeps > JMAX_SWEEPS) max_sweeps = 16;
  JacobiWs w = jacobi_carve(workspace, B, n);
  // tol_override > 0: first pass of the two-pass solver (its own threshold and raised significance floor);
  // tol_override < 0: second pass (default threshold, floor jacobi_nu_pass2); 0: single-pass solver
  const float tol = tol_override > 0.f ? tol_override : options().jacobi_tol;
  const bool tc = options().jacobi_update_tc != 0 && panel_tc_supported(w.np);
  PanelTc ptc;
  const bool overlap = tc && options().jacobi_overlap_v != 0;
  StreamRef g_streams;
  std::vector<SuperRound> plan;
  const bool sym = tc && options().panel_sym != 0 && panel_sym_supported(w.np);
  const bool want_chain = sym && options().jacobi_schedule == 2 && panel_chain_supported(w.np);
  if (tc && (options().jacobi_schedule == 1 || want_chain) && w.Sg != nullptr) plan = spread_plan(w.nb);
  const bool chain = want_chain && !plan.empty();
  const bool spread = !plan.empty() && !chain;
  if (tc) {
    if (int e = panel_tc_prepare(&ptc, w.Gp, w.H, w.Vt, w.Qb[0], w.Qb[1], B, w.np)) return e;
    if (sym) { if (int e = panel_sym_prepare(&ptc)) return e; }
    if (overlap) { if (int e = g_streams.ensure()) return e; }
  }
  {
    dim3 grid(std::min<int64_t>((int64_t(w.np) * w.np + 255) / 256, 64), (unsigned)B);
    R3D_STAGE(ST_JACOBI_INIT, st);
    jacobi_init_kernel<<<grid, 256, 0, st>>>(G, int(n), w.np, w.Gp, w.Vt, w.cnt, w.nu,
                                             tol_override > 0.f ? options().jacobi_nu_pass1
                                                                : (tol_override < 0.f ? options().jacobi_nu_pass2 : 4.f));
    R3D_LAUNCH_CHECK();
  }
  const size_t upd_smem = size_t(3) * JM * (JM + 4) * sizeof(float);
  R3D_CUDA(cudaFuncSetAttribute(jacobi_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd_smem));
  const int rounds = w.nb - 1;
  const int upd_tiles = w.nt * w.nt + w.nt * (w.np / JM);
  int iter = 0;
  bool v_pending[2] = {false, false};
  // V(big) of an iteration on the side stream: `launch_v(vs)` issues it
  auto fork_v = [&](int qb, auto&& launch_v) -> int {
    cudaStream_t vs = g_streams->vst[chunk];
    R3D_CUDA(cudaEventRecord(g_streams->ev_inner[chunk][qb], st));
    R3D_CUDA(cudaStreamWaitEvent(vs, g_streams->ev_inner[chunk][qb], 0));
    if (int e = launch_v(vs)) return e;
    R3D_CUDA(cudaEventRecord(g_streams->ev_v[chunk][qb], vs));
    v_pending[qb] = true;
    return 0;
  };
  // one ordinary round (pairing code r: >= 0 circle method, < 0 XOR mask) on the full matrices
  auto plain_round = [&](int r, bool generic, int sweep, int qb) -> int {
    {
      R3D_STAGE(ST_JACOBI_INNER, st);
      if (!generic && options().jacobi_inner_regs != 0)
        (options().jacobi_inner_regs == 3 ? jacobi_inner_cross_kernel<3> : jacobi_inner_cross_kernel<4>)<<<dim3(w.nt, (unsigned)B), 256, 0, st>>>(w.Gp, w.np, w.nb, w.nt, r, sweep, w.cnt,
                                                                          w.qflag[qb], w.Qb[qb], tol, w.nu, w.psync, 1);
      else
        jacobi_inner_kernel<<<dim3(w.nt, (unsigned)B), 256, 0, st>>>(w.Gp, w.np, w.nb, w.nt, r, sweep, w.cnt,
                                                                    w.qflag[qb], w.Qb[qb], tol, w.nu, 1, w.psync, 1,
                                                                    generic ? 1 : 0);
      R3D_LAUNCH_CHECK();
    }
    if (tc) {
      auto vlaunch = [&](cudaStream_t vs) { return panel_tc_update_v(&ptc, qb, r, sweep, w.cnt, w.qflag[qb], vs); };
      if (overlap && options().jacobi_v_after_g == 0) { if (int e = fork_v(qb, vlaunch)) return e; }
      else if (!overlap) { if (int e = vlaunch(st)) return e; }
      if (sym) { if (int e = panel_sym_update_g(&ptc, qb, r, sweep, w.cnt, w.qflag[qb], st)) return e; }
      else if (int e = panel_tc_update_g(&ptc, qb, r, sweep, w.cnt, w.qflag[qb], st, spread ? nullptr : w.psync)) return e;
      // jacobi_v_after_g: start V(r) only when the G passes of round r are done, so that it overlaps inner(r+1)
      // instead of competing with the G passes for HBM bandwidth
      if (overlap && options().jacobi_v_after_g != 0) { if (int e = fork_v(qb, vlaunch)) return e; }
    } else {
      R3D_STAGE(ST_JACOBI_UPDATE, st);
      jacobi_update_kernel<<<dim3(upd_tiles, (unsigned)B), 256, upd_smem, st>>>(w.Gp, w.Vt, w.np, w.nb, w.nt, r,
                                                                               sweep, w.cnt, w.qflag[qb], w.Qb[qb]);
      R3D_LAUNCH_CHECK();
    }
    return 0;
  };
  // chained schedule: the three XOR rounds of a super-round run as ordinary rounds on G (inner solve + one-pass
  // symmetric update each), their V updates as ONE chained pass over V on the side stream
  bool vc_pending[2] = {false, false};
  int chain_iter = 0;
  if (chain) { if (int e = panel_tc_prepare_chain(&ptc, w.Qc)) return e; }
  auto chain_round = [&](const SuperRound& sr, bool first_of_sweep, int sweep) -> int {
    const int slot = (chain_iter++) & 1;
    if (overlap && vc_pending[slot]) {      // the chained pass that last read this slot's Q buffers must be done
      R3D_CUDA(cudaStreamWaitEvent(st, g_streams->ev_v_c[chunk][slot], 0));
      vc_pending[slot] = false;
    }
    const int masks[3] = {sr.a, sr.b, sr.a ^ sr.b};
    for (int k = 0; k < 3; ++k) {
      const int qi = 3 * slot + k, code = -masks[k];
      const bool generic = first_of_sweep && k == 0;
      {
        R3D_STAGE(ST_JACOBI_INNER, st);
        if (!generic && options().jacobi_inner_regs != 0)
          (options().jacobi_inner_regs == 3 ? jacobi_inner_cross_kernel<3> : jacobi_inner_cross_kernel<4>)<<<dim3(w.nt, (unsigned)B), 256, 0, st>>>(w.Gp, w.np, w.nb, w.nt, code, sweep, w.cnt,
                                                                            w.qflagc[qi], w.Qc[qi], tol, w.nu, nullptr, 1);
        else
          jacobi_inner_kernel<<<dim3(w.nt, (unsigned)B), 256, 0, st>>>(w.Gp, w.np, w.nb, w.nt, code, sweep, w.cnt,
                                                                      w.qflagc[qi], w.Qc[qi], tol, w.nu, 1, nullptr, 1,
                                                                      generic ? 1 : 0);
        R3D_LAUNCH_CHECK();
      }
      // the chain needs the three inner solves only: fork right behind the third one, beside its G update
      if (k == 2 && overlap) R3D_CUDA(cudaEventRecord(g_streams->ev_inner_c[chunk][slot], st));
      if (int e = panel_sym_update_g(&ptc, 2 + qi, code, sweep, w.cnt, w.qflagc[qi], st)) return e;
    }
    const PanelGroups grp = groups_of(sr);
    const int* fl[3] = {w.qflagc[3 * slot], w.qflagc[3 * slot + 1], w.qflagc[3 * slot + 2]};
    if (!overlap) return panel_tc_update_v_chain(&ptc, slot, grp, sweep, w.cnt, fl, st);
    cudaStream_t vs = g_streams->vst[chunk];
    R3D_CUDA(cudaStreamWaitEvent(vs, g_streams->ev_inner_c[chunk][slot], 0));
    if (int e = panel_tc_update_v_chain(&ptc, slot, grp, sweep, w.cnt, fl, vs)) return e;
    R3D_CUDA(cudaEventRecord(g_streams->ev_v_c[chunk][slot], vs));
    vc_pending[slot] = true;
    return 0;
  };
  PanelTc loc;
  const int ng = w.np / 128;
  if (spread) {
    if (int e = panel_tc_prepare(&loc, w.Sg, w.Sh, w.Pv, w.Ql[0], w.Ql[1], B * ng, 128)) return e;
    loc.bdiv = ng; loc.local = 1;
    if (sym) { if (int e = panel_sym_prepare(&loc)) return e; }
    if (int e = panel_tc_prepare_groups(&ptc, w.Pt[0], w.Pt[1])) return e;
  }
  // one super-round of the spread schedule: three XOR rounds inside the 4-block groups, then ONE pass over G and V
  auto super_round = [&](const SuperRound& sr, bool first_of_sweep, int sweep, int pb) -> int {
    const PanelGroups grp = groups_of(sr);
    const dim3 ggrid((unsigned)ng, (unsigned)B);
    {
      R3D_STAGE(ST_JACOBI_LOCAL, st);
      group_gather_kernel<<<ggrid, 256, 0, st>>>(w.Gp, w.np, w.nb, ng, grp, sweep, w.cnt, w.Sg, w.Pv);
      R3D_LAUNCH_CHECK();
    }
    for (int k = 0; k < 3; ++k) {
      const int ql = k & 1, code = -(k + 1);
      const bool generic = first_of_sweep && k == 0;      // once per sweep the pairs inside each block rotate too
      {
        R3D_STAGE(ST_JACOBI_INNER, st);
        if (!generic && options().jacobi_inner_regs != 0)
          jacobi_inner_cross_kernel<4><<<dim3(2, (unsigned)(B * ng)), 256, 0, st>>>(w.Sg, 128, 4, 2, code, sweep, w.cnt, w.lflag[k],
                                                                                w.Ql[ql], tol, w.nu, nullptr, ng);
        else
          jacobi_inner_kernel<<<dim3(2, (unsigned)(B * ng)), 256, 0, st>>>(w.Sg, 128, 4, 2, code, sweep, w.cnt, w.lflag[k],
                                                                          w.Ql[ql], tol, w.nu, 1, nullptr, ng, generic ? 1 : 0);
        R3D_LAUNCH_CHECK();
      }
      if (int e = panel_tc_update_v(&loc, ql, code, sweep, w.cnt, w.lflag[k], st)) return e;          // P <- P Q_k
      if (k < 2) {
        if (sym) { if (int e = panel_sym_update_g(&loc, ql, code, sweep, w.cnt, w.lflag[k], st)) return e; }
        else if (int e = panel_tc_update_g(&loc, ql, code, sweep, w.cnt, w.lflag[k], st, nullptr)) return e;
      }
    }
    {
      R3D_STAGE(ST_JACOBI_LOCAL, st);
      group_transpose_kernel<<<ggrid, 256, 0, st>>>(w.Pv, ng, sweep, w.cnt, w.lflag[0], w.lflag[1], w.lflag[2], w.Pt[pb],
                                                    w.gflag[pb]);
      R3D_LAUNCH_CHECK();
    }
    auto vlaunch = [&](cudaStream_t vs) { return panel_tc_update_v_groups(&ptc, pb, grp, sweep, w.cnt, w.gflag[pb], vs); };
    if (overlap) { if (int e = fork_v(pb, vlaunch)) return e; }
    else { if (int e = vlaunch(st)) return e; }
    return panel_tc_update_g_groups(&ptc, pb, grp, sweep, w.cnt, w.gflag[pb], st);
  };
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    if (sweep > 0) {
      jacobi_active_kernel<<<1, 256, 0, st>>>(w.cnt, (int)B, sweep);
      R3D_LAUNCH_CHECK();
    }
    const int steps = (spread || chain) ? (int)plan.size() : rounds;
    for (int r = 0; r < steps; ++r, ++iter) {
      const int qb = tc ? (iter & 1) : 0;
      if (overlap && v_pending[qb]) {        // the V update that last read this Q / P^T buffer must be done
        R3D_CUDA(cudaStreamWaitEvent(st, g_streams->ev_v[chunk][qb], 0));
        v_pending[qb] = false;
      }
      if (chain && plan[r].b != 0) { if (int e = chain_round(plan[r], r == 0, sweep)) return e; }
      else if (!spread && !chain) { if (int e = plain_round(r, r == 0, sweep, qb)) return e; }
      else if (plan[r].b == 0) { if (int e = plain_round(-plan[r].a, false, sweep, qb)) return e; }
      else { if (int e = super_round(plan[r], r == 0, sweep, qb)) return e; }
    }
  }
  for (int i = 0; i < 2; ++i) {
    if (overlap && v_pending[i]) R3D_CUDA(cudaStreamWaitEvent(st, g_streams->ev_v[chunk][i], 0));
    if (overlap && vc_pending[i]) R3D_CUDA(cudaStreamWaitEvent(st, g_streams->ev_v_c[chunk][i], 0));
  }
  {
    dim3 grid((unsigned)std::min<int64_t>((n + 7) / 8, 64), (unsigned)B);
    R3D_STAGE(ST_JACOBI_EXTRACT, st);
    if (tc)
      jacobi_extract_cols_kernel<<<dim3((unsigned)((n + JB - 1) / JB), (unsigned)B), 256, 0, st>>>(
          w.Gp, w.Vt, int(n), w.np, w.cnt, max_sweeps, lambda_out, U_out, sweeps_out);
    else
      jacobi_extract_kernel<<<grid, 256, 0, st>>>(w.Gp, w.Vt, int(n), w.np, w.cnt, max_sweeps, lambda_out, U_out,
                                                  sweeps_out, 0);
    R3D_LAUNCH_CHECK();
  }
  return 0;
}

// number of side-stream sets created so far in this process (they are pooled: stays flat across short-lived threads)
extern "C" int r3d_stream_sets_created(void) { return g_sets_created.load(); }

// Split the batch into two chunks on two streams when it is large enough to fill the GPU twice over.
static int jacobi_run_chunked(const float* G, int64_t B, int64_t n, void* workspace, float* lambda_out, float* U_out,
                              int32_t* sweeps_out, int max_sweeps, cudaStream_t st, float tol_override) {
  const int np = jacobi_np(n);
  // chunks: as many as asked for (<= kMaxChunks) while every chunk still fills the GPU once with inner-solver CTAs
  int nch = std::min(options().jacobi_chunks, kMaxChunks);
  if (!(panel_tc_supported(np) && options().jacobi_update_tc != 0)) nch = 1;
  while (nch > 1 && (B / nch) * (np / JM) < kNumSMs) --nch;
  if (nch <= 1) return jacobi_run(G, B, n, workspace, lambda_out, U_out, sweeps_out, max_sweeps, st, 0, tol_override);
  StreamRef g_streams;
  if (int e = g_streams.ensure()) return e;
  R3D_CUDA(cudaEventRecord(g_streams->ev_fork, st));
  char* ws = (char*)workspace;
  int64_t b0 = 0;
  for (int c = 0; c < nch; ++c) {
    const int64_t Bc = B / nch + (c < B % nch ? 1 : 0);
    const bool own = c > 0 || options().jacobi_own_streams != 0;   // chunk 0 may stay on the caller's stream
    cudaStream_t sc = own ? g_streams->chunk[c] : st;
    if (own) R3D_CUDA(cudaStreamWaitEvent(sc, g_streams->ev_fork, 0));
    if (int e = jacobi_run(G + b0 * n * n, Bc, n, ws, lambda_out ? lambda_out + b0 * n : nullptr,
                           U_out ? U_out + b0 * n * n : nullptr, sweeps_out ? sweeps_out + b0 : nullptr, max_sweeps, sc, c,
                           tol_override)) return e;
    if (own) R3D_CUDA(cudaEventRecord(g_streams->ev_join[c], sc));
    ws += (jacobi_ws_bytes_one(Bc, n) + 255) & ~size_t(255);
    b0 += Bc;
  }
  for (int c = options().jacobi_own_streams != 0 ? 0 : 1; c < nch; ++c)
    R3D_CUDA(cudaStreamWaitEvent(st, g_streams->ev_join[c], 0));
  return 0;
}

extern "C" int r3d_jacobi_eigh(const float* G, int64_t B, int64_t n, void* workspace, float* lambda_out, float* U_out,
                               int32_t* sweeps_out, int max_sweeps, void* stream) {
  R3D_CHECK(G && workspace, "null pointer");
  R3D_CHECK(B >= 1 && n >= 1 && n <= 8192, "bad shape B=%lld n=%lld", (long long)B, (long long)n);
  return jacobi_run_chunked(G, B, n, workspace, lambda_out, U_out, sweeps_out, max_sweeps, (cudaStream_t)stream);
}

template <typename T>
static int refine_Y(const void* x, const float* Ut, int64_t B, int64_t Tt, int64_t C, float* Y, cudaStream_t st) {
  int64_t n, m; bool ts; side(Tt, C, n, m, ts);
  const T* X = (const T*)x;
  // T-side: Y[j,c] = sum_r Ut[j,r] X[r,c]          (NN)
  // C-side: Y[j,t] = sum_c Ut[j,c] X[t,c]          (NT)
  R3D_STAGE(ST_REFINE_Y, st);
  return sgemm_launch<float, T, float>(false, !ts, Ut, X, Y, int(n), int(m), int(n), n, C, m, n * n, Tt * C, n * m,
                                       nullptr, 0, 0, int(B), st);
}

extern "C" int r3d_erank_fwd(const void* x, int64_t B, int64_t T, int64_t C, int dtype, float rtol, int gram_impl,
                             void* workspace, float* erank_out, float* sigma_out, float* U_out, float* Y_out,
                             int32_t* sweeps_out, void* stream) {
  R3D_CHECK(x && workspace && erank_out && sigma_out && U_out && Y_out, "null pointer");
  R3D_CHECK(B >= 1 && T >= 1 && C >= 1, "bad shape");
  R3D_CHECK(dtype == R3D_F32 || dtype == R3D_BF16, "bad dtype %d", dtype);
  R3D_CHECK(rtol >= 0.f && rtol < 1.f, "rtol must be in [0, 1)");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t n, m; bool ts; side(T, C, n, m, ts);
  R3D_CHECK(n <= 8192, "min(T, C) = %lld exceeds 8192", (long long)n);
  const ErankWs w = erank_carve(workspace, B, T, C, dtype);
  if (int e = r3d_gram(x, B, T, C, dtype, gram_impl, workspace, w.G, st)) return e;
  const bool two_pass = options().erank_passes >= 2;
  int32_t* sweeps2 = sweeps_out ? reinterpret_cast<int32_t*>(w.coef) : nullptr;   // coef is free until the backward
  if (int e = jacobi_run_chunked(w.G, B, n, w.jws, nullptr, U_out, sweeps_out,
                                 two_pass ? options().erank_pass1_sweeps : 0, st,
                                 two_pass ? options().jacobi_tol_pass1 : 0.f)) return e;
  if (tc_gemm_ok(T, C) && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    if (int e = refine_Y_tc(x, dtype, U_out, w, B, T, C, Y_out, st)) return e;
    if (two_pass) {
      if (int e = second_pass_tc(w, B, n, m, U_out, Y_out, sweeps2, st)) return e;
    }
  } else {
    if (int e = (dtype == R3D_F32 ? refine_Y<float>(x, U_out, B, T, C, Y_out, st)
                                  : refine_Y<__nv_bfloat16>(x, U_out, B, T, C, Y_out, st))) return e;
    if (two_pass) {
      if (int e = second_pass_simt(w, B, n, m, U_out, Y_out, sweeps2, st)) return e;
    }
  }
  if (two_pass && sweeps_out) {
    sweeps_merge_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(sweeps_out, sweeps2, (int)B);
    R3D_LAUNCH_CHECK();
  }
  {
    R3D_STAGE(ST_SIGMA, st);
    row_sigma_kernel<<<(unsigned)((B * n + 7) / 8), 256, 0, st>>>(Y_out, U_out, B * n, int(n), int(m), sigma_out);
    R3D_LAUNCH_CHECK();
  }
  {
    R3D_STAGE(ST_ENTROPY, st);
    erank_entropy_kernel<<<(unsigned)B, 256, 0, st>>>(sigma_out, int(n), rtol, erank_out);
    R3D_LAUNCH_CHECK();
  }
  return 0;
}

template <typename T>
static int bwd_gemm(const float* Ut, const float* Y, const float* coef, int64_t B, int64_t Tt, int64_t C, void* dx,
                    int accumulate, cudaStream_t st) {
  int64_t n, m; bool ts; side(Tt, C, n, m, ts);
  T* D = (T*)dx;
  R3D_STAGE(ST_BWD_GEMM, st);
  if (ts)   // dX[t,c] = sum_j Ut[j,t] coef_j Y[j,c]   : a(t,j)=Ut[j*n+t] (TRANS_A), b(j,c)=Y[j*m+c]
    return sgemm_launch<float, float, T>(true, false, Ut, Y, D, int(Tt), int(C), int(n), n, m, C, n * n, n * m, Tt * C,
                                         coef, n, accumulate, int(B), st);
  // C-side: dX[t,c] = sum_j Y[j,t] coef_j Ut[j,c]     : a(t,j)=Y[j*m+t] (TRANS_A), b(j,c)=Ut[j*n+c]
  return sgemm_launch<float, float, T>(true, false, Y, Ut, D, int(Tt), int(C), int(n), m, n, C, n * m, n * n, Tt * C,
                                       coef, n, accumulate, int(B), st);
}

extern "C" int r3d_erank_bwd(const float* g, const float* erank, const float* sigma, const float* U, const float* Y,
                             int64_t B, int64_t T, int64_t C, int dtype, float rtol, void* workspace, void* dx,
                             int accumulate, void* stream) {
  R3D_CHECK(g && sigma && U && Y && workspace && dx, "null pointer");
  R3D_CHECK(dtype == R3D_F32 || dtype == R3D_BF16, "bad dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t n, m; bool ts; side(T, C, n, m, ts);
  const ErankWs w = erank_carve(workspace, B, T, C, dtype);
  float* coef = w.coef;
  {
    R3D_STAGE(ST_COEF, st);
    erank_coef_kernel<<<(unsigned)B, 256, 0, st>>>(sigma, g, int(n), rtol, coef);
    R3D_LAUNCH_CHECK();
  }
  if (tc_gemm_ok(T, C) && (reinterpret_cast<uintptr_t>(dx) & 15) == 0)
    return bwd_gemm_tc(U, Y, coef, w, B, T, C, dtype, dx, accumulate, st);
  return dtype == R3D_F32 ? bwd_gemm<float>(U, Y, coef, B, T, C, dx, accumulate, st)
                          : bwd_gemm<__nv_bfloat16>(U, Y, coef, B, T, C, dx, accumulate, st);
}

extern "C" int r3d_token_informativeness(const float* sigma, const float* U, int64_t B, int64_t n, float rtol,
                                         float* score_out, void* stream) {
  R3D_CHECK(sigma && U && score_out, "null pointer");
  R3D_CHECK(n >= 1 && n <= 8192, "bad n");
  R3D_STAGE(ST_TOKEN_INFO, (cudaStream_t)stream);
  token_info_kernel<<<(unsigned)B, 256, size_t(n) * 4, (cudaStream_t)stream>>>(sigma, U, int(n), rtol, score_out);
  R3D_LAUNCH_CHECK();
  return 0;
}


// ================================================================================
// The whole hot path from HOST buffers through the C ABI (what a non-Python caller binds; bench.py's `e2e_c_abi`):
//   H2D inputs (+ upstream gradient) -> erank(rgb), erank(depth) -> channel score -> bottom-k -> exchange/stack
//   -> exchange backward of the upstream gradient + d(mean erank)/dX accumulated -> D2H of every result.
// One device arena per call (cudaMallocAsync, stream-ordered); synchronises the stream before returning.
// ================================================================================
extern "C" int r3d_fuser_step_host(const void* rgb_host, const void* depth_host, const void* gst_host, int64_t B,
                                   int64_t T, int64_t C, int dtype, int64_t k, float rtol, float erank_weight,
                                   void* stacked_out_host, float* erank_out_host, void* d_rgb_host, void* d_depth_host,
                                   int64_t* idx_host, void* stream) {
  R3D_CHECK(rgb_host && depth_host && stacked_out_host && erank_out_host, "null host pointer");
  R3D_CHECK(dtype == R3D_F32 || dtype == R3D_BF16, "bad dtype %d", dtype);
  R3D_CHECK(B >= 1 && T >= 1 && C >= 1 && C <= 8192 && k >= 0 && k <= C, "bad shape");
  R3D_CHECK((gst_host == nullptr) == (d_rgb_host == nullptr) && (gst_host == nullptr) == (d_depth_host == nullptr),
            "the backward needs gst_host, d_rgb_host and d_depth_host together");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = B * T;
  int64_t n, m; bool ts; side(T, C, n, m, ts);
  const size_t es = dtype == R3D_F32 ? 4 : 2;
  const size_t nb = size_t(rows) * C * es;                       // one modality
  auto al = [](size_t v) { return (v + 255) & ~size_t(255); };
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += al(bytes); return o; };
  const size_t o_in = take(2 * nb), o_out = take(2 * nb), o_g = take(gst_host ? 2 * nb : 0), o_dg = take(gst_host ? 2 * nb : 0);
  const size_t o_ws = take(r3d_erank_workspace_bytes(2 * B, T, C, dtype));
  const size_t o_er = take(size_t(2 * B) * 4), o_sig = take(size_t(2 * B) * n * 4), o_U = take(size_t(2 * B) * n * n * 4);
  const size_t o_Y = take(size_t(2 * B) * n * m * 4), o_sw = take(size_t(2 * B) * 4), o_gv = take(size_t(2 * B) * 4);
  const size_t o_sws = take(r3d_score_workspace_floats(rows, C) * 4), o_pk = take(size_t(2 * C + 2) * 4);
  const size_t o_idx = take(size_t(2) * (k > 0 ? k : 1) * 8);
  char* buf = nullptr;
  keep_async_pool();
  R3D_CUDA(cudaMallocAsync((void**)&buf, off + 256, st));
  int rc = 0;
  do {
    if (cudaMemcpyAsync(buf + o_in, rgb_host, nb, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(buf + o_in + nb, depth_host, nb, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        (gst_host && cudaMemcpyAsync(buf + o_g, gst_host, 2 * nb, cudaMemcpyHostToDevice, st) != cudaSuccess)) {
      set_error("H2D copy failed"); rc = 2; break;
    }
    float* er = (float*)(buf + o_er);
    int64_t* idx = (int64_t*)(buf + o_idx);
    float* packed = (float*)(buf + o_pk);
    if ((rc = r3d_erank_fwd(buf + o_in, 2 * B, T, C, dtype, rtol, 0, buf + o_ws, er, (float*)(buf + o_sig),
                            (float*)(buf + o_U), (float*)(buf + o_Y), (int32_t*)(buf + o_sw), st))) break;
    if ((rc = r3d_channel_score_partial(buf + o_in, buf + o_in + nb, rows, C, dtype, (float*)(buf + o_sws), st))) break;
    if ((rc = r3d_score_finalize_packed((float*)(buf + o_sws), rows, C, er, 2 * B, packed, st))) break;
    if ((rc = r3d_bottomk_scaled(packed, 2, C, k, packed + 2 * C + 1, idx, nullptr, st))) break;
    if ((rc = r3d_exchange_fwd(buf + o_in, buf + o_in + nb, idx, idx + k, k, nullptr, nullptr, R3D_BLEND_SWAP, buf + o_out,
                               rows, C, dtype, st))) break;
    if (gst_host) {
      std::vector<float> gv(size_t(2 * B), erank_weight / float(2 * B));
      if (cudaMemcpyAsync(buf + o_gv, gv.data(), gv.size() * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) {
        set_error("H2D copy failed"); rc = 2; break;
      }
      if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("stream sync failed"); rc = 2; break; }   // gv is a stack vector
      if ((rc = r3d_exchange_bwd(buf + o_g, nullptr, nullptr, idx, idx + k, k, nullptr, nullptr, nullptr, R3D_BLEND_SWAP,
                                 buf + o_dg, buf + o_dg + nb, nullptr, rows, C, dtype, st))) break;
      if ((rc = r3d_erank_bwd((float*)(buf + o_gv), er, (float*)(buf + o_sig), (float*)(buf + o_U), (float*)(buf + o_Y),
                              2 * B, T, C, dtype, rtol, buf + o_ws, buf + o_dg, 1, st))) break;
      if (cudaMemcpyAsync(d_rgb_host, buf + o_dg, nb, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
          cudaMemcpyAsync(d_depth_host, buf + o_dg + nb, nb, cudaMemcpyDeviceToHost, st) != cudaSuccess) {
        set_error("D2H copy failed"); rc = 2; break;
      }
    }
    if (cudaMemcpyAsync(stacked_out_host, buf + o_out, 2 * nb, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaMemcpyAsync(erank_out_host, er, size_t(2 * B) * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) {
      set_error("D2H copy failed"); rc = 2; break;
    }
    if (k > 0 && idx_host) cudaMemcpyAsync(idx_host, idx, size_t(2) * k * 8, cudaMemcpyDeviceToHost, st);
  } while (0);
  cudaFreeAsync(buf, st);
  cudaError_t e = cudaStreamSynchronize(st);
  if (rc == 0 && e != cudaSuccess) {
    set_error("stream sync failed: %s", cudaGetErrorString(e));
    rc = 2;
  }
  return rc;
}
