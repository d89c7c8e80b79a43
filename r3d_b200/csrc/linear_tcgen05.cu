// Linear-layer GEMM on tcgen05 with fused epilogues (sm_100a) -- SURVEY.md rows f1 / f2 / f3:
// the GEMMs of the fuser Block (model/extras/transformerblock.py:79-93,118-135 as called from
// model/futr_safuser_tokenfusion.py:86-95) and of the RGB / depth input projections (tokenfusion.py:111,143,179-197),
// forward and backward, without a library GEMM and without separate elementwise kernels.
//
//     D (M x N)  =  epilogue( sum_k A[m, k] * B[n, k] )
//
// Operands are bf16 (fp32 tensors go through bf16 planes: x = b1 + b2 + b3, six plane products, fp32-level accuracy,
// as in pgemm_tcgen05.cu) and may each be K-major ([rows][K]) or MN-major ([K][rows]) -- TMA feeds both straight from
// the row-major tensors, so forward (X W^T), input gradient (dY W) and weight gradient (dY^T X) need no transposes.
//
// Persistent, warp-specialised: warp 0 TMA producer, warp 1 MMA issuer (tcgen05.mma.cta_group::1.kind::f16, M = 128,
// N = 128 or 256, K = 16, fp32 accumulate in TMEM), warp 2 TMEM allocator, warps 4-7 epilogue.  The accumulator is
// double-buffered in TMEM (2 x N columns), so the epilogue of tile i overlaps the main loop of tile i + 1; the
// shared-memory ring holds up to four 64-wide K slabs.  Weight gradients (K = rows of the batch, few output tiles) are
// split along K; the fp32 partial tiles are summed in a fixed order by lin_splitk_reduce_kernel.
//
// Epilogue (all optional, applied in this order to the fp32 accumulator x of element (m, n)):
//     x += bias[n];  aux_out[m, n] = x;  x = act(x) [GELU(erf) | ReLU];  x *= gelu'(aux_in[m, n]);  x += residual[m, n];
//     D[m, n] = round(x);  colsum[m_tile, n] = sum over the tile's rows of D[m, n] (or |D[m, n]|), rounded values,
// the last one being the bias-gradient / channel-score partial sums (fixed order, finalised by r3d_score_finalize or
// lin_colsum_finalize) that would otherwise cost one more pass over the tensor.
#include <cuda.h>

#include <algorithm>

#include "common.cuh"
#include "jacobi_tc.cuh"
#include "pgemm.cuh"
#include "tc_store.cuh"

namespace r3d {

namespace {

constexpr int LM = 128;
constexpr int LBK = 64;
constexpr int kLinThreads = 384;   // warps 0-2: producer / MMA / TMEM allocator, warp 3 idle, warps 4-11: epilogue
constexpr int kEpiWarps = 8;
// FAST variant (bf16 output, full tiles, no split-K): 16 epilogue warps on 16-column chunks.  The epilogue of the wide
// GEMMs is bound by latency (ncu: issue slots 32 % busy with 2 epilogue warps per scheduler, top stalls long_scoreboard /
// wait), so the cure is more warps in flight; 640 threads leave 102 registers each, hence the narrower chunks.
constexpr int kFastThreads = 640;
constexpr int kFastEpiWarps = 16;
constexpr int kFastStg = 32 * 48;  // staging bytes per warp: 32 rows x 32 B of payload, 48-byte pitch

__device__ __forceinline__ uint32_t ln_s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ln_bar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ln_s32(bar)), "r"(count));
}
__device__ __forceinline__ void ln_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ln_s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ln_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ln_s32(bar)) : "memory");
}
__device__ __forceinline__ void ln_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(ln_s32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void ln_tma_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(ln_s32(dst)), "l"(map), "r"(ln_s32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ bool ln_elect() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t ln_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((saddr >> 4) & 0x3fff);
  d |= uint64_t((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;       // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void ln_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void ln_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ln_s32(bar))
               : "memory");
}

struct LinDev {
  int M, N, K;
  int pa, pb, nprod, prod_a[6], prod_b[6];
  int a_kmajor, b_kmajor;
  int stages, stage_bytes;
  int splits, kb_per_split;      // split-K: `splits` K ranges of kb_per_split 64-wide slabs; output = fp32 partials
  float* partial;                // [splits][M][N] when splits > 1
  int out_f32;                   // dtype of D / residual / aux (1: fp32, 0: bf16); bias has the same dtype
  void* D; int64_t ldd;
  const void* bias;
  const void* residual; int64_t ldr;
  void* aux_out; const void* aux_in; int64_t ldx;
  float* colsum;                 // [m_tiles][N]
  int act, dgelu, colsum_abs;
};

// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, below fp32 GELU's own rounding): one exp, one reciprocal and a
// degree-5 polynomial instead of libm's branchy erff -- the epilogue warps are the bottleneck of the wide (N = 4C) GEMMs.
// e = exp(-z^2) is handed back because gelu' needs exp(-x^2 / 2) = e for z = x / sqrt(2).
__device__ __forceinline__ float erf_as(float z, float& e) {
  const float az = fabsf(z);
  const float t = __fdividef(1.f, fmaf(0.3275911f, az, 1.f));
  e = __expf(-az * az);
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  return copysignf(1.f - p * t * e, z);
}
__device__ __forceinline__ float gelu_f(float x) {
  float e;
  return 0.5f * x * (1.f + erf_as(x * 0.70710678118654752f, e));
}
__device__ __forceinline__ float dgelu_f(float x) {
  float e;
  const float er = erf_as(x * 0.70710678118654752f, e);
  return fmaf(x * 0.3989422804014327f, e, 0.5f * (1.f + er));
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
// bf16 tensors: GELU in its tanh form on the hardware tanh (one MUFU op, ~6 instructions per element instead of ~20).
// |gelu_tanh - gelu_erf| <= 3e-4 absolute, i.e. below half a bf16 ulp wherever |gelu| > 0.08 and far inside the 1e-2
// bf16 bar everywhere; fp32 tensors keep the erf form above.  The epilogue warps, not the tensor pipe, bound the wide
// (N = 4C) GEMMs, so instructions per element are what counts.
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float t = tanh_fast(x * fmaf(0.0356774081f, x * x, 0.7978845608f));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
__device__ __forceinline__ float dgelu_tanh_f(float x) {
  const float x2 = x * x;
  const float t = tanh_fast(x * fmaf(0.0356774081f, x2, 0.7978845608f));
  const float du = fmaf(0.1070322243f, x2, 0.7978845608f);
  return fmaf(0.5f * x * du, fmaf(-t, t, 1.f), fmaf(0.5f, t, 0.5f));
}

// element `i` (0..31) of a row segment starting at column col0 of a (rows, ld) tensor of dtype f32 / bf16
__device__ __forceinline__ void load_row32(const void* base, int f32, int64_t off, float (&o)[32]) {
  if (f32) {
    const float4* p = reinterpret_cast<const float4*>((const float*)base + off);
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float4 v = p[j]; o[4 * j] = v.x; o[4 * j + 1] = v.y; o[4 * j + 2] = v.z; o[4 * j + 3] = v.w; }
  } else {
    const uint4* p = reinterpret_cast<const uint4*>((const __nv_bfloat16*)base + off);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 v = p[j];
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        o[8 * j + 2 * u] = __uint_as_float(w[u] << 16);
        o[8 * j + 2 * u + 1] = __uint_as_float(w[u] & 0xffff0000u);
      }
    }
  }
}

// Coalesced store of a warp's 32 x 32 block as bf16 (lane = row, f[0..31] = 32 consecutive columns): staged through
// shared memory (80-byte pitch: 64 B of payload per row) and written as 16-byte pieces, 4 lanes per 64-byte row segment.
__device__ __forceinline__ void staged_store_bf16_32x32(uint8_t* stg, int lane, const float (&f)[32], __nv_bfloat16* dst,
                                                        int64_t pitch) {
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[8 * j + 2 * u], f[8 * j + 2 * u + 1]);
      w[u] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(stg + lane * kStgPitch + j * 16) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int id = j * 32 + lane, r = id >> 2, c = id & 3;
    const uint4 val = *reinterpret_cast<const uint4*>(stg + r * kStgPitch + c * 16);
    *reinterpret_cast<uint4*>(dst + r * pitch + c * 8) = val;
  }
}

// Coalesced load of a warp's 32 x 32 block (the mirror image of the staged stores): 16-byte pieces, 4 (bf16) or 8
// (fp32) lanes per row segment, staged through shared memory; lane = row on return.
__device__ __forceinline__ void staged_load_32x32(uint8_t* stg, int lane, const void* base, int f32, int64_t off0,
                                                  int64_t pitch, float (&o)[32]) {
  if (f32) {
    const float* src = (const float*)base + off0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {                     // 16 columns at a time (the staging row holds 64 B of payload)
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int id = j * 32 + lane, r = id >> 2, c = id & 3;
        *reinterpret_cast<float4*>(stg + r * kStgPitch + c * 16) =
            *reinterpret_cast<const float4*>(src + r * pitch + h * 16 + c * 4);
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(stg + lane * kStgPitch + j * 16);
        o[16 * h + 4 * j] = v.x; o[16 * h + 4 * j + 1] = v.y; o[16 * h + 4 * j + 2] = v.z; o[16 * h + 4 * j + 3] = v.w;
      }
    }
  } else {
    const __nv_bfloat16* src = (const __nv_bfloat16*)base + off0;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int id = j * 32 + lane, r = id >> 2, c = id & 3;
      *reinterpret_cast<uint4*>(stg + r * kStgPitch + c * 16) = *reinterpret_cast<const uint4*>(src + r * pitch + c * 8);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 v = *reinterpret_cast<const uint4*>(stg + lane * kStgPitch + j * 16);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        o[8 * j + 2 * u] = __uint_as_float(w[u] << 16);
        o[8 * j + 2 * u + 1] = __uint_as_float(w[u] & 0xffff0000u);
      }
    }
  }
}

// The bf16 row-block load split in two, so that the global-load latency hides behind the TMEM read and the
// arithmetic of the chunk: issue (four 16-byte pieces per lane into registers) early, finish (through the staging
// buffer, lane = row afterwards) where the values are needed.
__device__ __forceinline__ void row_block_issue(const void* base, int64_t off0, int64_t pitch, int lane, uint4 (&raw)[4]) {
  const __nv_bfloat16* src = (const __nv_bfloat16*)base + off0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int id = j * 32 + lane, r = id >> 2, c = id & 3;
    raw[j] = *reinterpret_cast<const uint4*>(src + r * pitch + c * 8);
  }
}
__device__ __forceinline__ void row_block_finish(uint8_t* stg, int lane, const uint4 (&raw)[4], float (&o)[32]) {
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int id = j * 32 + lane, r = id >> 2, c = id & 3;
    *reinterpret_cast<uint4*>(stg + r * kStgPitch + c * 16) = raw[j];
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 v = *reinterpret_cast<const uint4*>(stg + lane * kStgPitch + j * 16);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      o[8 * j + 2 * u] = __uint_as_float(w[u] << 16);
      o[8 * j + 2 * u + 1] = __uint_as_float(w[u] & 0xffff0000u);
    }
  }
}

// ---- 32 x 16 bf16 blocks (FAST epilogue): 2 lanes per 32-byte row segment ----
__device__ __forceinline__ void blk16_issue(const void* base, int64_t off0, int64_t pitch, int lane, uint4 (&raw)[2]) {
  const __nv_bfloat16* src = (const __nv_bfloat16*)base + off0;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int id = j * 32 + lane, r = id >> 1, c = id & 1;
    raw[j] = *reinterpret_cast<const uint4*>(src + r * pitch + c * 8);
  }
}
__device__ __forceinline__ void unpack16(const uint4& a, const uint4& b, float (&o)[16]) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    o[2 * u] = __uint_as_float(w[u] << 16);
    o[2 * u + 1] = __uint_as_float(w[u] & 0xffff0000u);
  }
}
__device__ __forceinline__ void blk16_finish(uint8_t* stg, int lane, const uint4 (&raw)[2], float (&o)[16]) {
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int id = j * 32 + lane, r = id >> 1, c = id & 1;
    *reinterpret_cast<uint4*>(stg + r * 48 + c * 16) = raw[j];
  }
  __syncwarp();
  unpack16(*reinterpret_cast<const uint4*>(stg + lane * 48), *reinterpret_cast<const uint4*>(stg + lane * 48 + 16), o);
}
__device__ __forceinline__ void blk16_store(uint8_t* stg, int lane, const float (&f)[16], __nv_bfloat16* dst, int64_t pitch) {
  uint32_t w[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * u], f[2 * u + 1]);
    w[u] = *reinterpret_cast<uint32_t*>(&h);
  }
  __syncwarp();
  *reinterpret_cast<uint4*>(stg + lane * 48) = make_uint4(w[0], w[1], w[2], w[3]);
  *reinterpret_cast<uint4*>(stg + lane * 48 + 16) = make_uint4(w[4], w[5], w[6], w[7]);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int id = j * 32 + lane, r = id >> 1, c = id & 1;
    *reinterpret_cast<uint4*>(dst + r * pitch + c * 8) = *reinterpret_cast<const uint4*>(stg + r * 48 + c * 16);
  }
}
// sum over the warp's 32 rows of 16 per-lane values: lanes 0-15 (and, duplicated, 16-31) end up with column (lane & 15)
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], 16);
#pragma unroll
  for (int half = 8; half >= 1; half >>= 1) {
    const bool up = (lane & half) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float send = up ? v[j] : v[j + half];
      const float recv = __shfl_xor_sync(0xffffffffu, send, half);
      v[j] = (up ? v[j + half] : v[j]) + recv;
    }
  }
  return v[0];
}

// sum over the warp's 32 rows of 32 per-lane values: lane l ends up holding the column-l total (31 shuffles)
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool up = (lane & half) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      // keep columns [0, half) if the lane's bit is clear, [half, 2 half) otherwise; hand the other half over
      const float send = up ? v[j] : v[j + half];
      const float recv = __shfl_xor_sync(0xffffffffu, send, half);
      v[j] = (up ? v[j + half] : v[j]) + recv;
    }
  }
  return v[0];
}

template <int TN, bool FAST = false>
__global__ void __launch_bounds__(FAST ? kFastThreads : kLinThreads, 1) lin_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                                   const __grid_constant__ CUtensorMap map_b,
                                                                                   LinDev g) {
  constexpr int EW = FAST ? kFastEpiWarps : kEpiWarps;
  constexpr int STGB = FAST ? kFastStg : kStgWarpBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)((uintptr_t(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_plane_bytes = LM * 128;              // 128 (m) x 64 (k) bf16
  const int b_plane_bytes = TN * 128;
  uint8_t* stg_base = smem + g.stages * g.stage_bytes;                 // 8 x 2560 B load / store staging
  float* cs_smem = (float*)(stg_base + EW * STGB);                     // [4 lane quarters][TN] column sums
  uint64_t* full_bar = (uint64_t*)(cs_smem + (FAST ? 8 : 4) * TN);
  uint64_t* empty_bar = full_bar + 4;
  uint64_t* tmem_full = empty_bar + 4;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (g.N + TN - 1) / TN, m_tiles = (g.M + LM - 1) / LM;
  const int total = g.splits * m_tiles * n_tiles;
  const int num_kb_all = (g.K + LBK - 1) / LBK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < g.stages; ++s) { ln_bar_init(&full_bar[s], 1); ln_bar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { ln_bar_init(&tmem_full[s], 1); ln_bar_init(&tmem_empty[s], EW * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ln_s32(tmem_slot)),
                 "n"(2 * TN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (split, m tile, n tile): n fastest, so CTAs running side by side share the A rows through L2
  auto decode = [&](int tile, int& sp, int& m0, int& n0, int& kb0, int& kb1) {
    const int per = m_tiles * n_tiles;
    sp = tile / per;
    const int r = tile - sp * per;
    m0 = (r / n_tiles) * LM;
    n0 = (r % n_tiles) * TN;
    kb0 = sp * g.kb_per_split;
    kb1 = min(num_kb_all, kb0 + g.kb_per_split);
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (ln_elect()) {
      int it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        int sp, m0, n0, kb0, kb1;
        decode(tile, sp, m0, n0, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % g.stages;
          const uint32_t ph = (it / g.stages) & 1;
          ln_wait(&empty_bar[s], ph ^ 1);
          uint8_t* stage = smem + s * g.stage_bytes;
          ln_expect_tx(&full_bar[s], g.stage_bytes);
          const int k0 = kb * LBK;
          for (int p = 0; p < g.pa; ++p) {
            uint8_t* dst = stage + p * a_plane_bytes;
            if (g.a_kmajor) {
              ln_tma_3d(dst, &map_a, &full_bar[s], k0, m0, p);                 // 64 k x 128 m rows
            } else {
              ln_tma_3d(dst, &map_a, &full_bar[s], m0, k0, p);                 // two 64-m slabs x 64 k rows
              ln_tma_3d(dst + 8192, &map_a, &full_bar[s], m0 + 64, k0, p);
            }
          }
          for (int p = 0; p < g.pb; ++p) {
            uint8_t* dst = stage + g.pa * a_plane_bytes + p * b_plane_bytes;
            if (g.b_kmajor) {
#pragma unroll
              for (int q = 0; q < TN / 128; ++q) ln_tma_3d(dst + q * 16384, &map_b, &full_bar[s], k0, n0 + q * 128, p);
            } else {
#pragma unroll
              for (int q = 0; q < TN / 64; ++q) ln_tma_3d(dst + q * 8192, &map_b, &full_bar[s], n0 + q * 64, k0, p);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (ln_elect()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(g.a_kmajor ? 0 : 1) << 15) |
                             (uint32_t(g.b_kmajor ? 0 : 1) << 16) | (uint32_t(TN >> 3) << 17) |
                             (uint32_t(LM >> 4) << 24);
      int it = 0, tt = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tt) {
        int sp, m0, n0, kb0, kb1;
        decode(tile, sp, m0, n0, kb0, kb1);
        const int ab = tt & 1;
        ln_wait(&tmem_empty[ab], ((tt >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem_base + ab * TN;
        uint32_t acc = 0;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % g.stages;
          const uint32_t ph = (it / g.stages) & 1;
          ln_wait(&full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_base = ln_s32(smem + s * g.stage_bytes);
          const uint32_t b_base = a_base + g.pa * a_plane_bytes;
          for (int pr = 0; pr < g.nprod; ++pr) {
            const uint32_t aa = a_base + g.prod_a[pr] * a_plane_bytes;
            const uint32_t bb = b_base + g.prod_b[pr] * b_plane_bytes;
#pragma unroll
            for (int k = 0; k < LBK / 16; ++k) {
              const uint64_t ad = g.a_kmajor ? ln_desc(aa + k * 32, 16, 1024) : ln_desc(aa + k * 2048, 8192, 1024);
              const uint64_t bd = g.b_kmajor ? ln_desc(bb + k * 32, 16, 1024) : ln_desc(bb + k * 2048, 8192, 1024);
              ln_umma(d, ad, bd, idesc, acc);
              acc = 1;
            }
          }
          ln_commit(&empty_bar[s]);
        }
        ln_commit(&tmem_full[ab]);
      }
    }
  } else if (FAST && warp >= 4) {
    // ===================== epilogue, FAST variant: bf16, full tiles, 16 warps x 16-column chunks =====================
    const int q = warp & 3, cset = (warp - 4) >> 2;        // lane quarter, column set 0..3
    uint8_t* stg = stg_base + (warp - 4) * STGB;
    __nv_bfloat16* Dp = (__nv_bfloat16*)g.D;
    int tt = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tt) {
      int sp, m0, n0, kb0, kb1;
      decode(tile, sp, m0, n0, kb0, kb1);
      const int ab = tt & 1;
      ln_wait(&tmem_full[ab], (tt >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row_w = m0 + q * 32;
#pragma unroll 1
      for (int c0 = cset * 16; c0 < TN; c0 += 64) {
        const int col0 = n0 + c0;
        uint4 raw_x[2], raw_r[2];
        if (g.dgelu) blk16_issue(g.aux_in, int64_t(row_w) * g.ldx + col0, g.ldx, lane, raw_x);
        if (g.residual) blk16_issue(g.residual, int64_t(row_w) * g.ldr + col0, g.ldr, lane, raw_r);
        uint32_t v[16];
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(ab * TN + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(v[j]);
        if (g.bias) {
          float bv[16];                                    // broadcast loads: every lane reads the same 32 bytes
          const uint4* bp = reinterpret_cast<const uint4*>((const __nv_bfloat16*)g.bias + col0);
          unpack16(bp[0], bp[1], bv);
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] += bv[j];
        }
        if (g.aux_out) blk16_store(stg, lane, x, (__nv_bfloat16*)g.aux_out + int64_t(row_w) * g.ldx + col0, g.ldx);
        if (g.act == 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] = gelu_tanh_f(bf16_round(x[j]));
        } else if (g.act == 2) {
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] = fmaxf(x[j], 0.f);
        }
        if (g.dgelu) {
          float h[16];
          blk16_finish(stg, lane, raw_x, h);
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] *= dgelu_tanh_f(h[j]);
        }
        if (g.residual) {
          float r[16];
          blk16_finish(stg, lane, raw_r, r);
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] += r[j];
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = bf16_round(x[j]);
        blk16_store(stg, lane, x, Dp + int64_t(row_w) * g.ldd + col0, g.ldd);
        if (g.colsum) {
          if (g.colsum_abs) {
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = fabsf(x[j]);
          }
          const float tot = warp_colsum16(x, lane);
          if (lane < 16) cs_smem[(ab * 4 + q) * TN + c0 + lane] = tot;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      ln_arrive(&tmem_empty[ab]);
      if (g.colsum) {
        // the per-quarter sums are double-buffered by tile parity: ONE barrier per tile (the barrier of the next tile
        // also orders this tile's reads before the buffer's reuse two tiles later), so warps can run ahead
        asm volatile("bar.sync 1, 512;" ::: "memory");
        const float* cs = cs_smem + ab * 4 * TN;
        const int t = threadIdx.x - 128;
        for (int c = t; c < TN; c += 512)
          g.colsum[int64_t(m0 / LM) * g.N + n0 + c] = ((cs[c] + cs[TN + c]) + cs[2 * TN + c]) + cs[3 * TN + c];
      }
    }
  } else if (!FAST && warp >= 4) {
    // ===================== epilogue =====================
    // eight warps: warp w reads TMEM lanes 32 (w % 4) .. + 31 (a hardware rule); warps 4-7 take the even 32-column
    // chunks of the tile, warps 8-11 the odd ones
    const int q = warp & 3, cset = (warp - 4) >> 2;
    uint8_t* stg = stg_base + (warp - 4) * kStgWarpBytes;
    int tt = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tt) {
      int sp, m0, n0, kb0, kb1;
      decode(tile, sp, m0, n0, kb0, kb1);
      const int ab = tt & 1;
      ln_wait(&tmem_full[ab], (tt >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row_w = m0 + q * 32, row = row_w + lane;
      const bool row_ok = row < g.M;
#pragma unroll 1
      for (int c0 = cset * 32; c0 < TN; c0 += 64) {
        const int col0 = n0 + c0;
        if (col0 >= g.N) break;                              // warp uniform
        // bf16 row tensors of the epilogue (gelu' input, residual): issue the loads now, use them after the TMEM read
        const bool pre_ok = !g.out_f32 && g.splits == 1 && (row_w + 32 <= g.M) && (col0 + 32 <= g.N);
        uint4 raw_x[4], raw_r[4];
        if (pre_ok && g.dgelu) row_block_issue(g.aux_in, int64_t(row_w) * g.ldx + col0, g.ldx, lane, raw_x);
        if (pre_ok && g.residual) row_block_issue(g.residual, int64_t(row_w) * g.ldr + col0, g.ldr, lane, raw_r);
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(ab * TN + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const bool full_blk = (row_w + 32 <= g.M) && (col0 + 32 <= g.N);     // warp uniform
        if (g.splits > 1) {
          // split-K partial tile: plain fp32, no epilogue; the reduce kernel applies nothing but the sum
          float* pbase = g.partial + (int64_t(sp) * g.M) * g.N;
          if (full_blk && (g.N & 3) == 0) {
            staged_store_32x32(stg, lane, v, pbase + int64_t(row_w) * g.N + col0, g.N, 0);
          } else if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < g.N) pbase[int64_t(row) * g.N + col0 + j] = __uint_as_float(v[j]);
          }
          continue;
        }
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
        const bool vec_in = full_blk;                       // 32 in-range columns: 128-bit row accesses
        if (g.bias) {
          if (col0 + 32 <= g.N && (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0) {
            float bv[32];                                   // broadcast 128-bit loads: every lane reads the same 32 values
            load_row32(g.bias, g.out_f32, col0, bv);
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] += bv[j];
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (col0 + j < g.N)
                x[j] += g.out_f32 ? ((const float*)g.bias)[col0 + j] : __bfloat162float(((const __nv_bfloat16*)g.bias)[col0 + j]);
            }
          }
        }
        if (g.aux_out) {
          if (full_blk) {
            if (g.out_f32) {
              uint32_t w[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) w[j] = __float_as_uint(x[j]);
              staged_store_32x32(stg, lane, w, (float*)g.aux_out + int64_t(row_w) * g.ldx + col0, g.ldx, 0);
            } else {
              staged_store_bf16_32x32(stg, lane, x, (__nv_bfloat16*)g.aux_out + int64_t(row_w) * g.ldx + col0, g.ldx);
            }
          } else if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < g.N) {
                if (g.out_f32) ((float*)g.aux_out)[int64_t(row) * g.ldx + col0 + j] = x[j];
                else ((__nv_bfloat16*)g.aux_out)[int64_t(row) * g.ldx + col0 + j] = __float2bfloat16_rn(x[j]);
              }
          }
        }
        if (g.act == 1) {                                   // GELU of the STORED (rounded) pre-activation
          if (g.out_f32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = gelu_f(x[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = gelu_tanh_f(bf16_round(x[j]));
          }
        } else if (g.act == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
        }
        if (g.dgelu && (row_ok || vec_in)) {
          float h[32];
          if (pre_ok) row_block_finish(stg, lane, raw_x, h);
          else if (vec_in) staged_load_32x32(stg, lane, g.aux_in, g.out_f32, int64_t(row_w) * g.ldx + col0, g.ldx, h);
          else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              h[j] = (col0 + j < g.N) ? (g.out_f32 ? ((const float*)g.aux_in)[int64_t(row) * g.ldx + col0 + j]
                                                   : __bfloat162float(((const __nv_bfloat16*)g.aux_in)[int64_t(row) * g.ldx + col0 + j]))
                                      : 0.f;
          }
          if (g.out_f32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] *= dgelu_f(h[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] *= dgelu_tanh_f(h[j]);
          }
        }
        if (g.residual && (row_ok || vec_in)) {
          float r[32];
          if (pre_ok) row_block_finish(stg, lane, raw_r, r);
          else if (vec_in) staged_load_32x32(stg, lane, g.residual, g.out_f32, int64_t(row_w) * g.ldr + col0, g.ldr, r);
          else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              r[j] = (col0 + j < g.N) ? (g.out_f32 ? ((const float*)g.residual)[int64_t(row) * g.ldr + col0 + j]
                                                   : __bfloat162float(((const __nv_bfloat16*)g.residual)[int64_t(row) * g.ldr + col0 + j]))
                                      : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] += r[j];
        }
        if (!g.out_f32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = bf16_round(x[j]);
        }
        if (full_blk) {
          if (g.out_f32) {
            uint32_t w[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) w[j] = __float_as_uint(x[j]);
            staged_store_32x32(stg, lane, w, (float*)g.D + int64_t(row_w) * g.ldd + col0, g.ldd, 0);
          } else {
            staged_store_bf16_32x32(stg, lane, x, (__nv_bfloat16*)g.D + int64_t(row_w) * g.ldd + col0, g.ldd);
          }
        } else if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < g.N) {
              if (g.out_f32) ((float*)g.D)[int64_t(row) * g.ldd + col0 + j] = x[j];
              else ((__nv_bfloat16*)g.D)[int64_t(row) * g.ldd + col0 + j] = __float2bfloat16_rn(x[j]);
            }
        }
        if (g.colsum) {
          float c[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) c[j] = row_ok ? (g.colsum_abs ? fabsf(x[j]) : x[j]) : 0.f;
          const float tot = warp_colsum32(c, lane);
          cs_smem[q * TN + c0 + lane] = tot;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      ln_arrive(&tmem_empty[ab]);
      if (g.colsum && g.splits == 1) {
        // the four epilogue warps (rows 0-31, 32-63, ...) add their column sums in warp order
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int t = threadIdx.x - 128;
        for (int c = t; c < TN; c += kEpiWarps * 32) {
          if (n0 + c < g.N)
            g.colsum[int64_t(m0 / LM) * g.N + n0 + c] =
                ((cs_smem[c] + cs_smem[TN + c]) + cs_smem[2 * TN + c]) + cs_smem[3 * TN + c];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * TN) : "memory");
  }
}

// D[m, n] (+)= sum over splits (in order) of partial[s][m, n]; output fp32 or bf16
__global__ void __launch_bounds__(256) lin_splitk_reduce_kernel(const float* __restrict__ partial, int splits, int64_t MN,
                                                                int64_t N, void* __restrict__ D, int64_t ldd, int out_f32,
                                                                int accumulate, const void* __restrict__ bias) {
  for (int64_t e = int64_t(blockIdx.x) * 256 + threadIdx.x; e < MN; e += int64_t(gridDim.x) * 256) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += partial[int64_t(k) * MN + e];
    const int64_t m = e / N, n = e - m * N;
    if (bias) s += out_f32 ? ((const float*)bias)[n] : __bfloat162float(((const __nv_bfloat16*)bias)[n]);
    if (out_f32) {
      float* p = (float*)D + m * ldd + n;
      *p = accumulate ? *p + s : s;
    } else {
      __nv_bfloat16* p = (__nv_bfloat16*)D + m * ldd + n;
      *p = __float2bfloat16_rn(accumulate ? __bfloat162float(*p) + s : s);
    }
  }
}

// out[n] = sum over parts (in order) of partial[part][n]: finalises the epilogue column sums (bias gradients)
__global__ void __launch_bounds__(256) lin_colsum_finalize_kernel(const float* __restrict__ partial, int parts, int64_t N,
                                                                  void* __restrict__ out, int out_f32) {
  __shared__ float red[8][33];
  const int cl = threadIdx.x & 31, gq = threadIdx.x >> 5;
  const int64_t c = int64_t(blockIdx.x) * 32 + cl;
  float s = 0.f;
  if (c < N) {
    const int per = (parts + 7) / 8;
    const int i0 = gq * per, i1 = min(parts, i0 + per);
    for (int i = i0; i < i1; ++i) s += partial[int64_t(i) * N + c];
  }
  red[gq][cl] = s;
  __syncthreads();
  if (gq == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][cl];
    if (out_f32) ((float*)out)[c] = t; else ((__nv_bfloat16*)out)[c] = __float2bfloat16_rn(t);
  }
}

// column sums of a (rows, C) tensor, stage 1: per-CTA partial rows (fixed order); abs optional
template <typename T>
__global__ void __launch_bounds__(256) lin_colsum_partial_kernel(const T* __restrict__ x, int64_t rows, int64_t C,
                                                                 int64_t rows_per_cta, float* __restrict__ partial) {
  // thread = one column (stride 256 over C in blockIdx.y chunks); rows in order
  const int64_t c = int64_t(blockIdx.y) * 256 + threadIdx.x;
  if (c >= C) return;
  const int64_t r0 = int64_t(blockIdx.x) * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int64_t r = r0;
  for (; r + 3 < r1; r += 4) {
    a0 += float(x[r * C + c]); a1 += float(x[(r + 1) * C + c]); a2 += float(x[(r + 2) * C + c]); a3 += float(x[(r + 3) * C + c]);
  }
  for (; r < r1; ++r) a0 += float(x[r * C + c]);
  partial[int64_t(blockIdx.x) * C + c] = (a0 + a1) + (a2 + a3);
}

// ReLU backward for the input projections (tokenfusion.py:183,197): dpre = y > 0 ? dy : 0, plus per-CTA column sums of
// dpre (= the bias gradient, finalised by lin_colsum_finalize_kernel).  Thread = two adjacent columns, rows in order.
template <typename T>
__global__ void __launch_bounds__(256) lin_relu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y,
                                                           int64_t rows, int64_t C, int64_t rows_per_cta,
                                                           T* __restrict__ dpre, float* __restrict__ partial) {
  const int64_t c = (int64_t(blockIdx.y) * 256 + threadIdx.x) * 2;
  if (c >= C) return;
  const bool two = c + 1 < C;
  const int64_t r0 = int64_t(blockIdx.x) * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  float a0 = 0.f, a1 = 0.f;
  for (int64_t r = r0; r < r1; ++r) {
    const float g0 = float(y[r * C + c]) > 0.f ? float(dy[r * C + c]) : 0.f;
    const float g1 = two && float(y[r * C + c + 1]) > 0.f ? float(dy[r * C + c + 1]) : 0.f;
    dpre[r * C + c] = T(g0);
    if (two) dpre[r * C + c + 1] = T(g1);
    a0 += float(T(g0)); a1 += float(T(g1));
  }
  partial[int64_t(blockIdx.x) * C + c] = a0;
  if (two) partial[int64_t(blockIdx.x) * C + c + 1] = a1;
}

// vectorised stage 1 of the column sums: a thread owns one 128-bit column vector (V columns) for every fourth row of
// its CTA's row chunk, four independent loads in flight; the four row lanes are added in order through shared memory
template <typename T, int V>
__global__ void __launch_bounds__(256) lin_colsum_partial_vec_kernel(const T* __restrict__ x, int64_t rows, int64_t C,
                                                                     int64_t rows_per_cta, float* __restrict__ partial) {
  __shared__ float red[4][64 * V];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int64_t c = (int64_t(blockIdx.y) * 64 + tx) * V;
  const bool active = c < C;
  const int64_t r0 = int64_t(blockIdx.x) * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  if (active) {
    for (int64_t r = r0 + ty; r < r1; r += 16) {
      float a[4][V];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (r + 4 * u < r1) load_vec<T, V>(x + (r + 4 * u) * C + c, a[u]);
        else {
#pragma unroll
          for (int i = 0; i < V; ++i) a[u][i] = 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] += (a[0][i] + a[1][i]) + (a[2][i] + a[3][i]);
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) red[ty][tx * V + i] = acc[i];
  __syncthreads();
  if (ty == 0 && active) {
#pragma unroll
    for (int i = 0; i < V; ++i)
      partial[int64_t(blockIdx.x) * C + c + i] = (red[0][tx * V + i] + red[1][tx * V + i]) + (red[2][tx * V + i] + red[3][tx * V + i]);
  }
}

typedef CUresult (*LinEncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
LinEncFn lin_encode() {
  static LinEncFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (LinEncFn)p;
  }
  return fn;
}

// physical tensor: [planes][rows][cols] bf16 with row pitch `ld` elements
int lin_make_map(CUtensorMap* m, const void* base, int64_t planes, int64_t rows, int64_t cols, int64_t ld,
                 int64_t plane_stride, bool kmajor) {
  LinEncFn enc = lin_encode();
  R3D_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  R3D_CHECK(ld % 8 == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0,
            "GEMM operands need a 16-byte aligned base and a row pitch that is a multiple of 8 elements");
  const cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)planes};
  const cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(planes > 1 ? plane_stride * 2 : rows * ld * 2)};
  const cuuint32_t box[3] = {64, (cuuint32_t)(kmajor ? 128 : 64), 1};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  R3D_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(linear) failed with %d", (int)r);
  return 0;
}

}  // namespace

}  // namespace r3d

using namespace r3d;

static size_t lin_align(size_t v) { return (v + 255) & ~size_t(255); }

// split-K plan: few output tiles and a long K (weight gradients) -> K ranges so that about one wave of CTAs is busy
static void lin_plan(int64_t M, int64_t N, int64_t K, int planes, int& TN, int& splits, int& kb_per) {
  TN = (N > 128 && planes == 1) ? 256 : 128;
  const int64_t tiles = ((M + LM - 1) / LM) * ((N + TN - 1) / TN);
  const int64_t num_kb = (K + LBK - 1) / LBK;
  splits = 1;
  if (tiles * 2 <= kNumSMs && num_kb >= 16) {
    // one wave: tiles * splits <= number of SMs (a second, nearly empty wave would double the time)
    int64_t want = std::min<int64_t>(kNumSMs / tiles, num_kb / 4);
    splits = (int)std::max<int64_t>(1, want);
  }
  kb_per = (int)((num_kb + splits - 1) / splits);
  splits = (int)((num_kb + kb_per - 1) / kb_per);
}

extern "C" size_t r3d_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int dtype) {
  size_t b = 1024;
  if (dtype == R3D_F32) b += lin_align(size_t(3) * M * K * 2) + lin_align(size_t(3) * N * K * 2);
  int TN, splits, kb_per;
  lin_plan(M, N, K, dtype == R3D_F32 ? 3 : 1, TN, splits, kb_per);
  if (splits > 1) b += lin_align(size_t(splits) * M * N * 4);
  return b;
}

extern "C" int r3d_gemm(const void* A, const void* Bm, void* D, int64_t M, int64_t N, int64_t K, int a_kmajor,
                        int b_kmajor, int dtype, const r3d_epilogue* epi, void* workspace, void* stream) {
  R3D_CHECK(A && Bm && D, "null pointer");
  R3D_CHECK(M >= 1 && N >= 1 && K >= 1 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "bad shape");
  R3D_CHECK(dtype == R3D_F32 || dtype == R3D_BF16, "bad dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  const bool f32 = dtype == R3D_F32;
  // physical shapes: K-major operand = [rows = M|N][cols = K]; MN-major operand = [rows = K][cols = M|N]
  const int64_t a_rows = a_kmajor ? M : K, a_cols = a_kmajor ? K : M;
  const int64_t b_rows = b_kmajor ? N : K, b_cols = b_kmajor ? K : N;
  R3D_CHECK(a_cols % 8 == 0 && b_cols % 8 == 0, "the contiguous dimension of both operands must be a multiple of 8");
  const void* Ap = A; const void* Bp = Bm;
  int planes = 1;
  char* ws = (char*)((uintptr_t(workspace) + 255) & ~uintptr_t(255));
  if (f32) {
    R3D_CHECK(workspace != nullptr, "fp32 GEMMs need the workspace (bf16 planes)");
    planes = 3;
    __nv_bfloat16* apl = (__nv_bfloat16*)ws; ws += lin_align(size_t(3) * M * K * 2);
    __nv_bfloat16* bpl = (__nv_bfloat16*)ws; ws += lin_align(size_t(3) * N * K * 2);
    if (int e = split_planes(A, R3D_F32, apl, M * K, 3, a_cols, nullptr, st)) return e;
    if (int e = split_planes(Bm, R3D_F32, bpl, N * K, 3, b_cols, nullptr, st)) return e;
    Ap = apl; Bp = bpl;
  }
  LinDev g{};
  g.M = (int)M; g.N = (int)N; g.K = (int)K;
  g.pa = planes; g.pb = planes; g.a_kmajor = a_kmajor != 0; g.b_kmajor = b_kmajor != 0;
  if (planes == 1) { g.nprod = 1; g.prod_a[0] = 0; g.prod_b[0] = 0; }
  else {
    static const int ia[6] = {0, 0, 1, 0, 2, 1}, ib[6] = {0, 1, 0, 2, 0, 1};
    g.nprod = 6;
    for (int i = 0; i < 6; ++i) { g.prod_a[i] = ia[i]; g.prod_b[i] = ib[i]; }
  }
  int TN, splits, kb_per;
  lin_plan(M, N, K, planes, TN, splits, kb_per);
  // split-K outputs go through the fp32 partial buffer and a reduce kernel that knows only the bias; any other
  // epilogue member keeps the GEMM in one piece
  const bool epi_beyond_bias = epi != nullptr && (epi->residual || epi->aux_out || epi->aux_in || epi->colsum_partial ||
                                                  epi->act != 0);
  if (epi_beyond_bias && splits > 1) { splits = 1; kb_per = (int)((K + LBK - 1) / LBK); }
  g.splits = splits; g.kb_per_split = kb_per;
  const void* reduce_bias = nullptr;
  if (splits > 1) {
    R3D_CHECK(workspace != nullptr, "split-K GEMMs need the workspace");
    g.partial = (float*)ws;
    if (epi != nullptr) reduce_bias = epi->bias;
  }
  g.stage_bytes = planes * LM * 128 + planes * TN * 128;
  g.stages = std::max(2, std::min(4, (197 * 1024) / g.stage_bytes));
  g.out_f32 = f32 ? 1 : 0;
  g.D = D; g.ldd = N;
  if (epi) {
    g.bias = epi->bias; g.residual = epi->residual; g.ldr = N;
    g.aux_out = epi->aux_out; g.aux_in = epi->aux_in; g.ldx = N;
    g.colsum = epi->colsum_partial; g.act = epi->act; g.dgelu = epi->aux_in != nullptr; g.colsum_abs = epi->colsum_abs;
    R3D_CHECK(g.act >= 0 && g.act <= 2, "bad activation %d", g.act);
    R3D_CHECK(N % 8 == 0 || (!g.residual && !g.aux_in), "row tensors of the epilogue need N %% 8 == 0");
  }
  CUtensorMap ma, mb;
  if (int e = lin_make_map(&ma, Ap, planes, a_rows, a_cols, a_cols, M * K, g.a_kmajor != 0)) return e;
  if (int e = lin_make_map(&mb, Bp, planes, b_rows, b_cols, b_cols, N * K, g.b_kmajor != 0)) return e;
  const int64_t tiles = int64_t(splits) * ((M + LM - 1) / LM) * ((N + TN - 1) / TN);
  const int grid = (int)std::min<int64_t>(tiles, kNumSMs);
  // FAST epilogue variant: bf16 output, no split-K, whole tiles only, 16-byte aligned row tensors
  const bool fast = !f32 && splits == 1 && M % LM == 0 && N % TN == 0 && N % 8 == 0 &&
                    ((uintptr_t(D) | uintptr_t(g.bias) | uintptr_t(g.residual) | uintptr_t(g.aux_out) | uintptr_t(g.aux_in)) & 15) == 0 &&
                    options().lin_fast != 0;
  const int smem = fast ? g.stages * g.stage_bytes + kFastEpiWarps * kFastStg + 8 * TN * 4 + 1024 + 256
                        : g.stages * g.stage_bytes + kEpiWarps * kStgWarpBytes + 4 * TN * 4 + 1024 + 256;
  R3D_CHECK(smem <= 227 * 1024, "linear GEMM: shared memory budget exceeded (%d)", smem);
  {
    R3D_STAGE(ST_BLOCK, st);
    if (fast) {
      static bool done[kMaxDevices] = {};
      if (per_device_once(done)) {
        R3D_CUDA(cudaFuncSetAttribute(lin_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        R3D_CUDA(cudaFuncSetAttribute(lin_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      }
      if (TN == 256) lin_kernel<256, true><<<grid, kFastThreads, smem, st>>>(ma, mb, g);
      else lin_kernel<128, true><<<grid, kFastThreads, smem, st>>>(ma, mb, g);
    } else if (TN == 256) {
      static bool done[kMaxDevices] = {};
      if (per_device_once(done))
        R3D_CUDA(cudaFuncSetAttribute(lin_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      lin_kernel<256><<<grid, kLinThreads, smem, st>>>(ma, mb, g);
    } else {
      static bool done[kMaxDevices] = {};
      if (per_device_once(done))
        R3D_CUDA(cudaFuncSetAttribute(lin_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      lin_kernel<128><<<grid, kLinThreads, smem, st>>>(ma, mb, g);
    }
    R3D_LAUNCH_CHECK();
    if (splits > 1) {
      const int64_t MN = M * N;
      const int rg = (int)std::min<int64_t>((MN + 255) / 256, int64_t(kNumSMs) * 8);
      lin_splitk_reduce_kernel<<<rg, 256, 0, st>>>(g.partial, splits, MN, N, D, N, g.out_f32, 0, reduce_bias);
      R3D_LAUNCH_CHECK();
    }
  }
  return 0;
}

extern "C" int r3d_colsum_finalize(const float* partial, int64_t parts, int64_t N, int dtype, void* out, void* stream) {
  R3D_CHECK(partial && out, "null pointer");
  R3D_CHECK(parts >= 1 && N >= 1, "bad shape");
  R3D_STAGE(ST_BLOCK, (cudaStream_t)stream);
  lin_colsum_finalize_kernel<<<(unsigned)((N + 31) / 32), 256, 0, (cudaStream_t)stream>>>(partial, (int)parts, N, out,
                                                                                         dtype == R3D_F32 ? 1 : 0);
  R3D_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t r3d_colsum_workspace_floats(int64_t rows, int64_t C) {
  const int64_t chunks = std::max<int64_t>(1, std::min<int64_t>((rows + 255) / 256, 2 * kNumSMs));
  return size_t(chunks) * size_t(C);
}

// out (C) = column sums of x (rows, C): the bias gradient of a Linear whose output gradient is x
extern "C" int r3d_colsum(const void* x, int64_t rows, int64_t C, int dtype, float* workspace, void* out, void* stream) {
  R3D_CHECK(x && workspace && out, "null pointer");
  R3D_CHECK(rows >= 1 && C >= 1, "bad shape");
  R3D_CHECK(dtype == R3D_F32 || dtype == R3D_BF16, "bad dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t chunks = std::max<int64_t>(1, std::min<int64_t>((rows + 255) / 256, 2 * kNumSMs));
  const int64_t rpc = (rows + chunks - 1) / chunks;
  dim3 grid((unsigned)chunks, (unsigned)((C + 255) / 256));
  R3D_STAGE(ST_BLOCK, st);
  const int V = dtype == R3D_F32 ? 4 : 8;
  if (C % V == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    dim3 vgrid((unsigned)chunks, (unsigned)((C / V + 63) / 64));
    if (dtype == R3D_F32) lin_colsum_partial_vec_kernel<float, 4><<<vgrid, 256, 0, st>>>((const float*)x, rows, C, rpc, workspace);
    else lin_colsum_partial_vec_kernel<__nv_bfloat16, 8><<<vgrid, 256, 0, st>>>((const __nv_bfloat16*)x, rows, C, rpc, workspace);
  } else if (dtype == R3D_F32) lin_colsum_partial_kernel<float><<<grid, 256, 0, st>>>((const float*)x, rows, C, rpc, workspace);
  else lin_colsum_partial_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, rows, C, rpc, workspace);
  R3D_LAUNCH_CHECK();
  lin_colsum_finalize_kernel<<<(unsigned)((C + 31) / 32), 256, 0, st>>>(workspace, (int)chunks, C, out,
                                                                        dtype == R3D_F32 ? 1 : 0);
  R3D_LAUNCH_CHECK();
  return 0;
}

// dpre (rows, C) = y > 0 ? dy : 0;  dbias (C, dtype) = column sums of dpre.  workspace: r3d_colsum_workspace_floats.
extern "C" int r3d_relu_bwd(const void* dy, const void* y, int64_t rows, int64_t C, int dtype, float* workspace,
                            void* dpre, void* dbias, void* stream) {
  R3D_CHECK(dy && y && workspace && dpre && dbias, "null pointer");
  R3D_CHECK(rows >= 1 && C >= 1, "bad shape");
  R3D_CHECK(dtype == R3D_F32 || dtype == R3D_BF16, "bad dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t chunks = std::max<int64_t>(1, std::min<int64_t>((rows + 255) / 256, 2 * kNumSMs));
  const int64_t rpc = (rows + chunks - 1) / chunks;
  dim3 grid((unsigned)chunks, (unsigned)((C + 511) / 512));
  R3D_STAGE(ST_BLOCK, st);
  if (dtype == R3D_F32)
    lin_relu_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)dy, (const float*)y, rows, C, rpc, (float*)dpre, workspace);
  else
    lin_relu_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, rows, C, rpc,
                                                             (__nv_bfloat16*)dpre, workspace);
  R3D_LAUNCH_CHECK();
  lin_colsum_finalize_kernel<<<(unsigned)((C + 31) / 32), 256, 0, st>>>(workspace, (int)chunks, C, dbias,
                                                                        dtype == R3D_F32 ? 1 : 0);
  R3D_LAUNCH_CHECK();
  return 0;
}
