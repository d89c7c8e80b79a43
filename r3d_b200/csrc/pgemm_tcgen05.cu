// Batched GEMM on tcgen05 with fp32-level accuracy from bf16 planes (sm_100a).
//
//   C[b] (M x N, fp32 accumulate in TMEM)  =  sum over (pa, pb) in `prods` of  A_pa[b] * B_pb[b]
//
// An fp32 matrix is split once into bf16 planes x = b1 + b2 + b3 (exact to 2^-24; split_planes_kernel), and
// the products b_i * b_j that matter are accumulated by kind::f16 MMAs.  16-bit operands may be K-major or
// MN-major under the plain SWIZZLE_128B layout, so -- unlike kind::tf32 -- every transpose combination the
// effective-rank chain needs is fed by TMA directly from the row-major tensors:
//
//   Y  = U^T A  (refinement)     A-op = U planes (K-major),  B-op = X (K-major, channel side) / (MN-major, token side)
//   dX = U diag(c) Y (backward)  both operands MN-major
//   G  = X^T X / X X^T for fp32 X (6 products of the 3 planes)
//
// Same warp-specialised structure as gram_tcgen05.cu: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM
// allocator, warps 4-7 epilogue; 2-stage mbarrier ring; M = 128, N = 128 or 256, K = 64 per stage.
#include <cuda.h>

#include <algorithm>

#include "common.cuh"
#include "pgemm.cuh"
#include "tc_store.cuh"

namespace r3d {

namespace {

constexpr int PG_M = 128;
constexpr int PG_BK = 64;
constexpr int PG_STAGES = 2;

__device__ __forceinline__ uint32_t pg_s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pg_bar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pg_s32(bar)), "r"(count));
}
__device__ __forceinline__ void pg_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pg_s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pg_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(pg_s32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void pg_tma_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(pg_s32(dst)), "l"(map), "r"(pg_s32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ bool pg_elect() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t pg_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((saddr >> 4) & 0x3fff);
  d |= uint64_t((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;       // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void pg_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void pg_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(pg_s32(bar))
               : "memory");
}

struct PGemmDev {
  int M, N, K;                 // logical sizes
  int batch;                   // number of matrices
  int pa, pb;                  // planes staged per operand
  int a_kmajor, b_kmajor;
  int nprod;
  int prod_a[6], prod_b[6];
  int out_mode;                // 0: fp32 store, 1: fp32 accumulate, 2: bf16 store, 3: bf16 accumulate
  void* C;
  int64_t ldc, strideC;
};

template <int TN>
__global__ void __launch_bounds__(256, 1) pgemm_kernel(const __grid_constant__ CUtensorMap map_a,
                                                       const __grid_constant__ CUtensorMap map_b, PGemmDev g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)((uintptr_t(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_plane_bytes = PG_M * 128;            // 16 KB: 128 (m) x 64 (k) bf16
  const int b_plane_bytes = TN * 128;
  const int stage_bytes = g.pa * a_plane_bytes + g.pb * b_plane_bytes;
  uint8_t* stg_base = smem + PG_STAGES * stage_bytes;              // epilogue staging, 2560 B per warp
  uint64_t* full_bar = (uint64_t*)(smem + PG_STAGES * stage_bytes + 4 * kStgWarpBytes);
  uint64_t* empty_bar = full_bar + PG_STAGES;
  uint64_t* tmem_full = empty_bar + PG_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (g.N + TN - 1) / TN, m_tiles = (g.M + PG_M - 1) / PG_M;
  const int b = blockIdx.x / (n_tiles * m_tiles);
  const int rem = blockIdx.x % (n_tiles * m_tiles);
  const int m0 = (rem / n_tiles) * PG_M, n0 = (rem % n_tiles) * TN;
  const int num_kb = (g.K + PG_BK - 1) / PG_BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < PG_STAGES; ++s) { pg_bar_init(&full_bar[s], 1); pg_bar_init(&empty_bar[s], 1); }
    pg_bar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(pg_s32(tmem_slot)), "n"(TN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (pg_elect()) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % PG_STAGES;
        const uint32_t ph = (kb / PG_STAGES) & 1;
        pg_wait(&empty_bar[s], ph ^ 1);
        uint8_t* stage = smem + s * stage_bytes;
        pg_expect_tx(&full_bar[s], stage_bytes);
        const int k0 = kb * PG_BK;
        for (int p = 0; p < g.pa; ++p) {
          uint8_t* dst = stage + p * a_plane_bytes;
          const int z = p * g.batch + b;
          if (g.a_kmajor) {
            pg_tma_3d(dst, &map_a, &full_bar[s], k0, m0, z);                 // 64 k x 128 m rows: 16 KB
          } else {
            pg_tma_3d(dst, &map_a, &full_bar[s], m0, k0, z);                 // two 64-m slabs x 64 k rows
            pg_tma_3d(dst + 8192, &map_a, &full_bar[s], m0 + 64, k0, z);
          }
        }
        for (int p = 0; p < g.pb; ++p) {
          uint8_t* dst = stage + g.pa * a_plane_bytes + p * b_plane_bytes;
          const int z = p * g.batch + b;
          if (g.b_kmajor) {
#pragma unroll
            for (int q = 0; q < TN / 128; ++q) pg_tma_3d(dst + q * 16384, &map_b, &full_bar[s], k0, n0 + q * 128, z);
          } else {
#pragma unroll
            for (int q = 0; q < TN / 64; ++q) pg_tma_3d(dst + q * 8192, &map_b, &full_bar[s], n0 + q * 64, k0, z);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (pg_elect()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(g.a_kmajor ? 0 : 1) << 15) |
                             (uint32_t(g.b_kmajor ? 0 : 1) << 16) | (uint32_t(TN >> 3) << 17) |
                             (uint32_t(PG_M >> 4) << 24);
      uint32_t acc = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % PG_STAGES;
        const uint32_t ph = (kb / PG_STAGES) & 1;
        pg_wait(&full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_base = pg_s32(smem + s * stage_bytes);
        const uint32_t b_base = a_base + g.pa * a_plane_bytes;
        for (int pr = 0; pr < g.nprod; ++pr) {
          const uint32_t aa = a_base + g.prod_a[pr] * a_plane_bytes;
          const uint32_t bb = b_base + g.prod_b[pr] * b_plane_bytes;
#pragma unroll
          for (int k = 0; k < PG_BK / 16; ++k) {
            const uint64_t ad = g.a_kmajor ? pg_desc(aa + k * 32, 16, 1024) : pg_desc(aa + k * 2048, 8192, 1024);
            const uint64_t bd = g.b_kmajor ? pg_desc(bb + k * 32, 16, 1024) : pg_desc(bb + k * 2048, 8192, 1024);
            pg_umma(tmem_base, ad, bd, idesc, acc);
            acc = 1;
          }
        }
        pg_commit(&empty_bar[s]);
      }
      pg_commit(tmem_full);
    }
  } else if (warp >= 4) {
    pg_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < TN; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(c0);
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
      const int row_w = m0 + q * 32, col0w = n0 + c0;
      if (g.out_mode <= 1 && row_w + 32 <= g.M && col0w + 32 <= g.N && (g.ldc & 3) == 0 &&
          ((reinterpret_cast<uintptr_t>(g.C) + (int64_t(b) * g.strideC) * 4) & 15) == 0) {
        // whole 32 x 32 fp32 block in range: coalesced store through shared memory (warp-uniform branch)
        staged_store_32x32(stg_base + q * kStgWarpBytes, lane, v,
                           (float*)g.C + int64_t(b) * g.strideC + int64_t(row_w) * g.ldc + col0w, g.ldc,
                           g.out_mode == 1);
      } else if (row < g.M) {
        const int col0 = n0 + c0;
        if (g.out_mode <= 1) {
          float* crow = (float*)g.C + int64_t(b) * g.strideC + int64_t(row) * g.ldc + col0;
          const bool vec = (col0 + 32 <= g.N) && ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0);
          if (vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                     __uint_as_float(v[j + 3]));
              if (g.out_mode == 1) {
                const float4 old = *reinterpret_cast<const float4*>(crow + j);
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
              }
              *reinterpret_cast<float4*>(crow + j) = o;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < g.N) crow[j] = __uint_as_float(v[j]) + (g.out_mode == 1 ? crow[j] : 0.f);
          }
        } else {
          __nv_bfloat16* crow = (__nv_bfloat16*)g.C + int64_t(b) * g.strideC + int64_t(row) * g.ldc + col0;
          const bool vec = (col0 + 32 <= g.N) && ((g.ldc & 7) == 0) && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0);
          if (vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float f[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) f[u] = __uint_as_float(v[j + u]);
              if (g.out_mode == 3) {
                const uint4 old = *reinterpret_cast<const uint4*>(crow + j);
                const uint32_t w[4] = {old.x, old.y, old.z, old.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  f[2 * u] += __uint_as_float(w[u] << 16);
                  f[2 * u + 1] += __uint_as_float(w[u] & 0xffff0000u);
                }
              }
              uint32_t o[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * u], f[2 * u + 1]);
                o[u] = *reinterpret_cast<uint32_t*>(&h);
              }
              *reinterpret_cast<uint4*>(crow + j) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < g.N) {
                const float old = g.out_mode == 3 ? __bfloat162float(crow[j]) : 0.f;
                crow[j] = __float2bfloat16_rn(__uint_as_float(v[j]) + old);
              }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TN) : "memory");
  }
}

// x (fp32 or bf16, `count` elements viewed as rows of `cols`) -> P bf16 planes [P][count]; optional per-row scale.
template <typename TIN>
__global__ void __launch_bounds__(256) split_planes_kernel(const TIN* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                           int64_t count, int P, int64_t cols,
                                                           const float* __restrict__ rowscale) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < count; e += stride) {
    float v = sizeof(TIN) == 4 ? (float)(*(const float*)(x + e)) : __bfloat162float(*(const __nv_bfloat16*)(x + e));
    if (rowscale) v *= rowscale[e / cols];
    float r = v;
    for (int p = 0; p < P; ++p) {
      const __nv_bfloat16 b = __float2bfloat16_rn(r);
      out[int64_t(p) * count + e] = b;
      r -= __bfloat162float(b);
    }
  }
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncFn pg_encode() {
  static EncFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncFn)p;
  }
  return fn;
}

// physical tensor: [planes * batch][rows][cols] bf16, row pitch = cols
int pg_make_map(CUtensorMap* m, const void* base, int64_t zcount, int64_t rows, int64_t cols, bool kmajor) {
  EncFn enc = pg_encode();
  R3D_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  R3D_CHECK(cols % 8 == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0, "pgemm operand needs cols %% 8 == 0");
  const cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)zcount};
  const cuuint64_t gstr[2] = {(cuuint64_t)cols * 2, (cuuint64_t)rows * cols * 2};
  const cuuint32_t box[3] = {64, (cuuint32_t)(kmajor ? 128 : 64), 1};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  R3D_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(pgemm) failed with %d", (int)r);
  return 0;
}

}  // namespace

int split_planes(const void* x, int in_dtype, __nv_bfloat16* out, int64_t count, int P, int64_t cols,
                 const float* rowscale, cudaStream_t st) {
  const int grid = (int)std::min<int64_t>((count + 255) / 256, int64_t(kNumSMs) * 16);
  if (in_dtype == R3D_F32)
    split_planes_kernel<float><<<grid, 256, 0, st>>>((const float*)x, out, count, P, cols, rowscale);
  else
    split_planes_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, out, count, P, cols, rowscale);
  R3D_LAUNCH_CHECK();
  return 0;
}

bool pgemm_operand_ok(const void* base, int64_t cols) {
  return cols % 8 == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0;
}

int pgemm_launch(const PGemm& a, cudaStream_t st) {
  R3D_CHECK(a.nprod >= 1 && a.nprod <= 6 && a.pa >= 1 && a.pa <= 3 && a.pb >= 1 && a.pb <= 3, "bad plane setup");
  CUtensorMap ma, mb;
  // physical shapes: K-major operand = [rows = M|N][cols = K]; MN-major operand = [rows = K][cols = M|N]
  if (int e = pg_make_map(&ma, a.A, int64_t(a.pa) * a.batch, a.a_kmajor ? a.M : a.K, a.a_kmajor ? a.K : a.M,
                          a.a_kmajor != 0)) return e;
  if (int e = pg_make_map(&mb, a.B, int64_t(a.pb) * a.batch, a.b_kmajor ? a.N : a.K, a.b_kmajor ? a.K : a.N,
                          a.b_kmajor != 0)) return e;
  // N tile: 256 when the staged planes fit two stages in shared memory, else 128
  int TN = 256;
  auto smem_for = [&](int tn) { return PG_STAGES * (a.pa * PG_M * 128 + a.pb * tn * 128) + 4 * kStgWarpBytes + 1024 + 256; };
  if (smem_for(256) > 220 * 1024 || a.N <= 128) TN = 128;
  R3D_CHECK(smem_for(TN) <= 227 * 1024, "pgemm: too many planes for shared memory");
  PGemmDev g;
  g.M = a.M; g.N = a.N; g.K = a.K; g.batch = a.batch; g.pa = a.pa; g.pb = a.pb;
  g.a_kmajor = a.a_kmajor; g.b_kmajor = a.b_kmajor; g.nprod = a.nprod;
  for (int i = 0; i < 6; ++i) { g.prod_a[i] = a.prod_a[i]; g.prod_b[i] = a.prod_b[i]; }
  g.out_mode = a.out_mode; g.C = a.C; g.ldc = a.ldc; g.strideC = a.strideC;
  const int64_t tiles = int64_t(a.batch) * ((a.M + PG_M - 1) / PG_M) * ((a.N + TN - 1) / TN);
  R3D_CHECK(tiles < (1ll << 31), "pgemm grid too large");
  const int smem = smem_for(TN);
  if (TN == 256) {
    R3D_CUDA(cudaFuncSetAttribute(pgemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    pgemm_kernel<256><<<(unsigned)tiles, 256, smem, st>>>(ma, mb, g);
  } else {
    R3D_CUDA(cudaFuncSetAttribute(pgemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    pgemm_kernel<128><<<(unsigned)tiles, 256, smem, st>>>(ma, mb, g);
  }
  R3D_LAUNCH_CHECK();
  return 0;
}

}  // namespace r3d
