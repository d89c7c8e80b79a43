// Batched Gram matrix on 5th-generation tensor cores (sm_100a) for bf16 X of shape (B, T, C), fp32
// accumulation in TMEM:  channel side  G[b] = X[b]^T X[b]  (n = C, T >= C; operands MN-major) and token side
// G[b] = X[b] X[b]^T  (n = T, T < C; operands K-major).  The two variants differ only in the TMA box, the
// shared-memory descriptor and the major bits of the instruction descriptor.
//
//   * operands are staged by TMA (cp.async.bulk.tensor.3d, 128-byte swizzle) straight
//     from the row-major (T, C) sample: a box of 64 channels x BK tokens is exactly one
//     MN-major SWIZZLE_128B UMMA slab, so no transpose is ever materialised;
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=256, K=16)
//     with both operands MN-major; the accumulator (128 lanes x 256 fp32 columns) lives
//     in TMEM; tcgen05.commit releases smem stages and signals the epilogue;
//   * four epilogue warps read the accumulator with tcgen05.ld (32x32b.x32) and store
//     128-byte row segments of G;
//   * warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..7 = epilogue;
//     a 2-stage full/empty mbarrier ring connects producer and issuer.  Two CTAs fit per
//     SM (97 KB smem, 256 TMEM columns each) so one CTA's epilogue overlaps the other's
//     main loop.
//
// SASS evidence to look for: UTCHMMA (tcgen05.mma), UTMALDG (TMA), LDTM (tcgen05.ld).
#include <cuda.h>

#include "common.cuh"
#include "tc_store.cuh"

namespace r3d {

namespace {

constexpr int TILE_M = 128;
constexpr int TILE_N = 256;
constexpr int BK = 64;            // tokens per pipeline stage
constexpr int STAGES = 2;            // x 2 resident CTAs per SM = 4 stages in flight per SM
constexpr int BOX_C = 64;         // channels per TMA box = 128 bytes of bf16 = one swizzle row
constexpr int SLAB_BYTES = BK * 128;                       // one 64-channel slab of a stage: 8 KB
constexpr int A_BYTES = (TILE_M / BOX_C) * SLAB_BYTES;     // 16 KB
constexpr int B_BYTES = (TILE_N / BOX_C) * SLAB_BYTES;     // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;             // 48 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 4 * kStgWarpBytes + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TMEM_COLS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}

// MN-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 (next 64-element MN slab),
//   [32,46) stride byte offset >> 4 (next 8-row K group), [46,48) version = 1, [61,64) layout = 2.
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((saddr >> 4) & 0x3fff);
  d |= uint64_t((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

// cute::UMMA::InstrDescriptor: c_format F32 (bit 4), a/b format BF16 (bits 7, 10),
// a/b major (bits 15, 16: 1 = MN-major, 0 = K-major), N >> 3 at [17,23), M >> 4 at [24,29).
constexpr uint32_t kInstrDescMN = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                  (uint32_t(TILE_N >> 3) << 17) | (uint32_t(TILE_M >> 4) << 24);
constexpr uint32_t kInstrDescK = (1u << 4) | (1u << 7) | (1u << 10) |
                                 (uint32_t(TILE_N >> 3) << 17) | (uint32_t(TILE_M >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TSIDE = false: n = C, K runs over tokens (BK = 64 tokens per stage).
// TSIDE = true : n = T, K runs over channels (64 channels = one 128-byte swizzle row per stage).
template <bool TSIDE>
__global__ void __launch_bounds__(256, 2) gram_bf16_kernel(const __grid_constant__ CUtensorMap tmap,
                                                           float* __restrict__ G, int T, int C) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)((uintptr_t(smem_raw) + 1023) & ~uintptr_t(1023));   // SWIZZLE_128B needs 1024 B
  uint8_t* stg_base = smem + STAGES * STAGE_BYTES;                 // epilogue staging, 2560 B per warp
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * STAGE_BYTES + 4 * kStgWarpBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = TSIDE ? T : C, kdim = TSIDE ? C : T;
  const int n_tiles = (n + TILE_N - 1) / TILE_N, m_tiles = (n + TILE_M - 1) / TILE_M;
  const int b = blockIdx.x / (n_tiles * m_tiles);
  const int rem = blockIdx.x % (n_tiles * m_tiles);
  const int m0 = (rem / n_tiles) * TILE_M, n0 = (rem % n_tiles) * TILE_N;
  const int num_kb = (kdim + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* stage = smem + s * STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        if (!TSIDE) {
          // boxes of 64 channels x 64 tokens: one MN-major slab each (8 KB)
#pragma unroll
          for (int h = 0; h < TILE_M / BOX_C; ++h)
            tma_load_3d(stage + h * SLAB_BYTES, &tmap, &full_bar[s], m0 + h * BOX_C, kb * BK, b);
#pragma unroll
          for (int q = 0; q < TILE_N / BOX_C; ++q)
            tma_load_3d(stage + A_BYTES + q * SLAB_BYTES, &tmap, &full_bar[s], n0 + q * BOX_C, kb * BK, b);
        } else {
          // boxes of 64 channels x 128 tokens: K-major rows of 128 bytes (16 KB each)
          tma_load_3d(stage, &tmap, &full_bar[s], kb * BK, m0, b);
#pragma unroll
          for (int q = 0; q < TILE_N / 128; ++q)
            tma_load_3d(stage + A_BYTES + q * 16384, &tmap, &full_bar[s], kb * BK, n0 + q * 128, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (elect_one()) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          if (!TSIDE) {
            // 16 tokens = two 8-row swizzle groups of 1024 B; slabs (64 channels) are SLAB_BYTES apart
            const uint64_t ad = make_desc_mn_sw128(a_addr + k * 2048, SLAB_BYTES, 1024);
            const uint64_t bd = make_desc_mn_sw128(b_addr + k * 2048, SLAB_BYTES, 1024);
            umma_bf16(tmem_base, ad, bd, kInstrDescMN, (kb | k) != 0);
          } else {
            // K-major: 16 channels = 32 bytes inside the 128-byte swizzle row; 8-row groups are 1024 B apart
            const uint64_t ad = make_desc_mn_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t bd = make_desc_mn_sw128(b_addr + k * 32, 16, 1024);
            umma_bf16(tmem_base, ad, bd, kInstrDescK, (kb | k) != 0);
          }
        }
        umma_commit(&empty_bar[s]);          // frees the smem stage once these MMAs have read it
      }
      umma_commit(tmem_full);                // accumulator complete
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> global =====
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;                          // this warp may touch TMEM lanes [32q, 32q+32)
    const int row = m0 + q * 32 + lane;
    float* grow = G + (int64_t(b) * n + row) * n + n0;
#pragma unroll 1
    for (int c0 = 0; c0 < TILE_N; c0 += 32) {
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
      const int row_w = m0 + q * 32;                                     // first row of this warp's block
      if (row_w + 32 <= n && n0 + c0 + 32 <= n && (n & 3) == 0) {
        // whole 32 x 32 block in range: coalesced store through shared memory (warp-uniform branch)
        staged_store_32x32(stg_base + q * kStgWarpBytes, lane, v, G + (int64_t(b) * n + row_w) * n + n0 + c0, n, 0);
      } else if (row < n) {
        if (n0 + c0 + 32 <= n && (n & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            __stcs(reinterpret_cast<float4*>(grow + c0 + j),
                   make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                               __uint_as_float(v[j + 3])));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + c0 + j < n) grow[c0 + j] = __uint_as_float(v[j]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

}  // namespace

bool gram_tcgen05_supported(int64_t B, int64_t T, int64_t C, int dtype) {
  // bf16 with a 16-byte-aligned row pitch; fp32 inputs take the SIMT Gram (32-bit MN-major operands need the
  // SWIZZLE_128B_BASE32B layout -- not implemented yet)
  const int64_t n = T < C ? T : C;
  const int64_t tiles = B * ((n + TILE_M - 1) / TILE_M) * ((n + TILE_N - 1) / TILE_N);
  return dtype == R3D_BF16 && C % 8 == 0 && T >= 1 && B >= 1 && tiles < (1ll << 31);
}

int gram_tcgen05_launch(const void* x, int64_t B, int64_t T, int64_t C, int dtype, void* workspace, float* G,
                        cudaStream_t st) {
  (void)workspace;
  R3D_CHECK(gram_tcgen05_supported(B, T, C, dtype), "shape/dtype not supported by the tcgen05 Gram");
  R3D_CHECK((reinterpret_cast<uintptr_t>(x) & 15) == 0, "x must be 16-byte aligned for TMA");
  EncodeTiledFn enc = get_encode();
  R3D_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  const bool tside = T < C;
  const int64_t n = tside ? T : C;
  CUtensorMap tmap;
  const cuuint64_t gdim[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B};
  const cuuint64_t gstride[2] = {(cuuint64_t)C * 2, (cuuint64_t)T * C * 2};   // bytes, dims 1 and 2
  const cuuint32_t box[3] = {BOX_C, (cuuint32_t)(tside ? 128 : BK), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(x), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  R3D_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d", (int)r);
  static bool attr_done[kMaxDevices] = {};
  if (per_device_once(attr_done)) {
    R3D_CUDA(cudaFuncSetAttribute(gram_bf16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    R3D_CUDA(cudaFuncSetAttribute(gram_bf16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  }
  const int grid = int(B * ((n + TILE_M - 1) / TILE_M) * ((n + TILE_N - 1) / TILE_N));
  R3D_STAGE(ST_GRAM, st);
  if (tside) gram_bf16_kernel<true><<<grid, 256, SMEM_BYTES, st>>>(tmap, G, int(T), int(C));
  else gram_bf16_kernel<false><<<grid, 256, SMEM_BYTES, st>>>(tmap, G, int(T), int(C));
  R3D_LAUNCH_CHECK();
  return 0;
}

}  // namespace r3d
