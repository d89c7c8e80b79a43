// One-pass symmetric two-sided Jacobi panel update on 5th-generation tensor cores (sm_100a), fp32-accurate.
//
// A round of the block two-sided Jacobi has nt disjoint block pairs ("tasks") c with rotation products Q_c (64x64)
// and must apply  G <- Q^T G Q.  In pair coordinates G is an nt x nt grid of 64x64 tiles and
//
//        G'[a, c] = Q_a^T  G[a, c]  Q_c ,
//
// so every tile can be updated on its own, IN PLACE, from two small rotation products -- no scratch matrix and no
// second pass.  G is symmetric, so only the tiles a <= c are computed; the mirror tile is the transposed store of the
// same accumulator.  Per round this reads 0.56 np^2 floats and writes np^2 (jacobi_tc.cu's two-pass scheme through the
// scratch H: 2 np^2 read + 2 np^2 written).
//
// Work unit: a 2x2 "super tile" (pair groups PA <= PC, two tasks each): rows (a1, a2) x columns (c1, c2), 128 x 128.
//   step 1   W[:, c] = [G[a1, c]; G[a2, c]] Q_c        for c = c1, c2     M = 128, N = 64, K = 64
//   step 2   D_a     = [W[a, c1]^T; W[a, c2]^T] Q_a    for a = a1, a2     M = 128, N = 64, K = 64
//            D_a[(c, j)][i] = G'[a, c][i][j]
// Both steps are the SAME UMMA shape with both operands K-major (the inner solver emits Q^T for that reason); between
// them W travels TMEM -> registers -> (hi, lo) split -> shared memory, transposed on the way: TMEM lane r = (a, k) of
// W becomes K index k of operand row (c, j), which is one 128-byte swizzled row segment per warp store.
// fp32 accuracy: 3xTF32 (x = hi + lo, hi = round-to-nearest TF32; hi hi + hi lo + lo hi, fp32 accumulate in TMEM).
//
// Pipeline per CTA (persistent over a strided list of super tiles, 2 CTAs per SM); a super tile makes 8 uses of a
// 2-stage ring of 48 KB stages (one 32-wide K slab of both operands, hi + lo):
//   warp 0      TMA producer: uses 0-3 the four 32x32 row blocks of G[(a1,a2), column block] + the hi and lo planes of the
//               K slab of Q_c^T (the inner solver emits Q^T pre-split); uses 4-7 only the planes of Q_a^T (the A part comes
//               from the transposers; these uses complete on a TMA barrier of their own, q_full)
//   warps 2-5   splitters (uses 0-3 only): the raw slab is the hi operand (kind::tf32 reads the upper 19 bits of the
//               container), they write lo = rn_tf32(x - trunc_tf32(x)) next to it
//   warp 1      MMA issuer: 12 x tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=64, K=8) per use; in uses 0-3 the eight that
//               need no lo plane of G go out when the TMA barrier fires, the other four after the splitters
//   warps 6-9   transposers + epilogue: W -> operand slabs of step 2 (warp q owns TMEM lanes 32q.. = K slab q of step 2),
//               then D_a -> global: direct tile by transposed stores (one 128-byte line per instruction), mirror tile
//               through the shared-memory staging of tc_store.cuh
// TMEM: 256 columns per CTA (W: 128, D_a1 | D_a2: 128).
#include <cuda.h>

#include <algorithm>

#include "common.cuh"
#include "jacobi_tc.cuh"
#include "tc_store.cuh"

namespace r3d {

extern int g_panel_debug;
extern int g_panel_grid_cap;

namespace {

constexpr int PB = 32;                 // Jacobi block width (JB in erank_kernels.cu)
constexpr int PM = 64;                 // pair width
constexpr int TM = 128;                // rows per MMA
constexpr int NSTAGE = 2;
constexpr int A_RAW = TM * PB * 4;     // 16 KB
constexpr int Q_RAW = PM * PB * 4;     // 8 KB
constexpr int STAGE = 2 * A_RAW + 2 * Q_RAW;
constexpr int STG_WARP = kStgWarpBytes;
constexpr int SMEM_TOTAL = NSTAGE * STAGE + 4 * STG_WARP + 1024 + 256;
constexpr int TMEM_COLS_S = 256;
constexpr int JMAXS = 32;              // JMAX_SWEEPS
constexpr int kSymThreads = 320;
constexpr int USES = 8;                // stage uses per super tile

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void bar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(s_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(s_u32(dst)), "l"(map), "r"(s_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(s_u32(dst)), "l"(map), "r"(s_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ bool elect() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}
// cute::UMMA::SmemDescriptor, SWIZZLE_128B (layout type 2), version 1
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((saddr >> 4) & 0x3fff);
  d |= uint64_t((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}
// InstrDescriptor: D fp32, A/B TF32, both K-major, N = 64, M = 128
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) |
                                (uint32_t(PM >> 3) << 17) | (uint32_t(TM >> 4) << 24);
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdescTf32), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_to(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar))
               : "memory");
}
__device__ __forceinline__ float tf32_rn(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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
}

// pairing of round r (same function as erank_kernels.cu / jacobi_tc.cu)
__device__ __forceinline__ void rr_pair_s(int m, int r, int t, int& a, int& b) {
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

// super tiles processed since the last reset, in units of 64 KB of HBM traffic (32 KB read + 32 KB written): a
// diagonal super tile (64 KB in, 64 KB out) counts 2, an off-diagonal one (64 KB in, 128 KB out) counts 3
__device__ unsigned long long g_sym_units;

struct SymTile { int b, pa, pc; bool run; };

// tile -> (matrix b, pair groups pa <= pc); run = false when the matrix converged or none of the four tasks rotated
__device__ __forceinline__ SymTile decode_sym(int tile, int nst, int npg, int nt, int sweep, int bdiv,
                                              const int* __restrict__ cnt, const int* __restrict__ qflag) {
  SymTile t;
  t.b = tile / nst;
  int r = tile - t.b * nst;
  int pa = 0;
  while (r >= npg - pa) { r -= npg - pa; ++pa; }
  t.pa = pa; t.pc = pa + r;
  t.run = true;
  if (sweep > 0 && cnt[(t.b / bdiv) * JMAXS + sweep - 1] == 0) t.run = false;
  else {
    const int* f = qflag + t.b * nt;
    if ((f[2 * t.pa] | f[2 * t.pa + 1] | f[2 * t.pc] | f[2 * t.pc + 1]) == 0) t.run = false;   // in place: nothing changes
  }
  return t;
}

__global__ void __maxnreg__(72) panel_sym_kernel(const __grid_constant__ CUtensorMap map_g32,
                                                 const __grid_constant__ CUtensorMap map_q, float* __restrict__ G,
                                                 int B, int np, int nb, int nt, int round, int sweep, int bdiv,
                                                 const int* __restrict__ cnt, const int* __restrict__ qflag) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)((uintptr_t(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg_base = smem + NSTAGE * STAGE;
  uint64_t* raw_full = (uint64_t*)(smem + NSTAGE * STAGE + 4 * STG_WARP);   // TMA -> splitters
  uint64_t* split_done = raw_full + NSTAGE;      // step 1: 128 splitters (A slab split) -> MMA
  uint64_t* tr_done = split_done + NSTAGE;       // step 2: the use's transposer warp (A slab written) -> MMA
  uint64_t* q_full = tr_done + NSTAGE;           // step 2: TMA (pre-split Q^T planes) -> MMA.  Its own barrier: every
                                                 // waiter of a barrier must see each of its phases (a waiter that skips
                                                 // phases can match a stale phase of the same parity)
  uint64_t* smem_empty = q_full + NSTAGE;        // MMA commit -> producer / transposer
  uint64_t* w_full = smem_empty + NSTAGE;        // MMA commit (step 1 done) -> transposers
  uint64_t* w_empty = w_full + 1;                // transposers (4 warps) -> MMA
  uint64_t* d_full = w_empty + 1;                // [2] MMA commit (step 2 of half h done) -> epilogue
  uint64_t* d_empty = d_full + 2;                // [2] epilogue (4 warps) -> MMA
  uint32_t* tmem_slot = (uint32_t*)(d_empty + 2);

  if (sweep > 0 && cnt[(B / bdiv) * JMAXS + sweep] == 0) return;   // every matrix converged
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npg = nt >> 1, nst = npg * (npg + 1) / 2;
  const int total = B * nst;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_g32) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      bar_init(&raw_full[s], 1); bar_init(&split_done[s], 128); bar_init(&tr_done[s], 1); bar_init(&smem_empty[s], 1);
      bar_init(&q_full[s], 1);
    }
    bar_init(w_full, 1); bar_init(w_empty, 4);
    for (int h = 0; h < 2; ++h) { bar_init(&d_full[h], 1); bar_init(&d_empty[h], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)),
                 "n"(TMEM_COLS_S) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect()) {
      int it = 0;
      unsigned long long units = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const SymTile ti = decode_sym(tile, nst, npg, nt, sweep, bdiv, cnt, qflag);
        if (!ti.run) continue;
        int rb[4], cb[4];
        rr_pair_s(nb, round, 2 * ti.pa, rb[0], rb[1]);
        rr_pair_s(nb, round, 2 * ti.pa + 1, rb[2], rb[3]);
        rr_pair_s(nb, round, 2 * ti.pc, cb[0], cb[1]);
        rr_pair_s(nb, round, 2 * ti.pc + 1, cb[2], cb[3]);
#pragma unroll 1
        for (int u = 0; u < USES; ++u, ++it) {
          const int s = it % NSTAGE;
          const uint32_t ph = (it / NSTAGE) & 1;
          bar_wait(&smem_empty[s], ph ^ 1);
          uint8_t* st = smem + s * STAGE;
          // Q^T arrives pre-split from the inner solver: hi plane at row 2 * task * 64, lo plane 64 rows below
          int qrow, qcol;
          uint64_t* fb = u < 4 ? &raw_full[s] : &q_full[s];
          if (u < 4) {
            bar_expect_tx(fb, A_RAW + 2 * Q_RAW);
            // K slab u of step 1 = column block cb[u]; rows = the four 32-row blocks of (a1, a2): contiguous 4 KB each
#pragma unroll
            for (int rg = 0; rg < 4; ++rg)
              tma_3d(st + rg * (PB * PB * 4), &map_g32, fb, 0, rb[rg] * PB, ti.b * nb + cb[u]);
            qcol = (u & 1) * PB; qrow = (ti.b * nt + 2 * ti.pc + (u >> 1)) * 2 * PM;
          } else {
            bar_expect_tx(fb, 2 * Q_RAW);
            const int h = (u - 4) >> 1, ks = (u - 4) & 1;
            qcol = ks * PB; qrow = (ti.b * nt + 2 * ti.pa + h) * 2 * PM;
          }
          tma_2d(st + 2 * A_RAW, &map_q, fb, qcol, qrow);
          tma_2d(st + 2 * A_RAW + Q_RAW, &map_q, fb, qcol, qrow + PM);
        }
        units += (ti.pa == ti.pc) ? 2 : 3;
      }
      if (units) atomicAdd(&g_sym_units, units);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect()) {
      int it = 0, tt = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const SymTile ti = decode_sym(tile, nst, npg, nt, sweep, bdiv, cnt, qflag);
        if (!ti.run) continue;
        const uint32_t tph = tt & 1;
        bar_wait(w_empty, tph ^ 1);                    // the transposers have read the previous W
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int u = 0; u < USES; ++u, ++it) {
          const int s = it % NSTAGE;
          const uint32_t ph = (it / NSTAGE) & 1;
          uint32_t d;
          if (u < 4) d = tmem_base + (u >> 1) * PM;    // W[:, c]
          else {
            const int h = (u - 4) >> 1;
            d = tmem_base + 2 * PM + h * PM;           // D_a
            if (((u - 4) & 1) == 0) {
              bar_wait(&d_empty[h], tph ^ 1);          // the epilogue has drained the previous D_a
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
          }
          // per super tile every stage sees 2 uses of step 1 and 2 of step 2 (8 uses, 2 stages): the phase parities of
          // the two hand-over barriers follow from the use index alone
          if (u < 4) bar_wait(&raw_full[s], (u >> 1) & 1);              // raw A slab (= hi operand) and Q planes landed
          else {
            bar_wait(&q_full[s], ((u - 4) >> 1) & 1);                   // pre-split Q^T planes landed (TMA)
            bar_wait(&tr_done[s], ((u - 4) >> 1) & 1);                  // A slab written by the transposer warp
          }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = s_u32(smem + s * STAGE), a_lo = a_hi + A_RAW;
          const uint32_t q_hi = a_hi + 2 * A_RAW, q_lo = q_hi + Q_RAW;
          uint32_t first = (u & 1);                    // each accumulator takes two K slabs
#pragma unroll
          for (int prod = 0; prod < 3; ++prod) {
            if (prod == 2 && u < 4) {                  // A_lo comes from the splitters
              bar_wait(&split_done[s], (u >> 1) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const uint32_t ab = (prod == 2) ? a_lo : a_hi;
            const uint32_t qb = (prod == 1) ? q_lo : q_hi;
#pragma unroll
            for (int kk = 0; kk < PB / 8; ++kk) {
              umma_tf32(d, desc_sw128(ab + kk * 32, 16, 1024), desc_sw128(qb + kk * 32, 16, 1024), first);
              first = 1;
            }
          }
          umma_commit_to(&smem_empty[s]);
          if (u == 3) umma_commit_to(w_full);
          if (u == 5) umma_commit_to(&d_full[0]);
          if (u == 7) umma_commit_to(&d_full[1]);
        }
        ++tt;
      }
    }
  } else if (warp >= 2 && warp < 6) {
    // ===================== splitters =====================
    const int t = threadIdx.x - 64;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const SymTile ti = decode_sym(tile, nst, npg, nt, sweep, bdiv, cnt, qflag);
      if (!ti.run) continue;
      // only the four A slabs of step 1 need splitting (Q^T arrives pre-split; step 2 gets A from the transposers)
#pragma unroll 1
      for (int u = 0; u < 4; ++u) {
        const int s = u % NSTAGE;                  // a super tile starts on stage 0 (8 uses per tile)
        bar_wait(&raw_full[s], (u >> 1) & 1);      // raw_full completes twice per super tile and stage
        uint8_t* st = smem + s * STAGE;
        float4* a_hi = reinterpret_cast<float4*>(st);
        float4* a_lo = reinterpret_cast<float4*>(st + A_RAW);
        // kind::tf32 reads the upper 19 bits of each fp32 container, so the raw slab IS the hi operand (hi = x with the
        // low 13 mantissa bits dropped); only lo = rn_tf32(x - hi) has to be written -- and the two products that do not
        // involve lo can be issued as soon as the TMA has landed
#pragma unroll 4
        for (int e = t; e < A_RAW / 16; e += 128) {
          const float4 x = a_hi[e];
          float4 l;
          l.x = tf32_rn(x.x - tf32_trunc(x.x));
          l.y = tf32_rn(x.y - tf32_trunc(x.y));
          l.z = tf32_rn(x.z - tf32_trunc(x.z));
          l.w = tf32_rn(x.w - tf32_trunc(x.w));
          a_lo[e] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        bar_arrive(&split_done[s]);
      }
    }
  } else if (warp >= 6) {
    // ===================== transposers + epilogue =====================
    const int q = warp & 3;                        // TMEM lanes [32q, 32q + 32)
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    int tt = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const SymTile ti = decode_sym(tile, nst, npg, nt, sweep, bdiv, cnt, qflag);
      if (!ti.run) continue;
      const uint32_t tph = tt & 1;
      int rb[4], cb[4];
      rr_pair_s(nb, round, 2 * ti.pa, rb[0], rb[1]);
      rr_pair_s(nb, round, 2 * ti.pa + 1, rb[2], rb[3]);
      rr_pair_s(nb, round, 2 * ti.pc, cb[0], cb[1]);
      rr_pair_s(nb, round, 2 * ti.pc + 1, cb[2], cb[3]);
      // ---- W (TMEM lane r = 32 q + lane, i.e. half h = q >> 1, k = 32 (q & 1) + lane) -> stage use 4 + q:
      //      operand row (c, j) = W column, K index = lane
      bar_wait(w_full, tph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      {
        const int it = tt * USES + 4 + q;
        const int s = it % NSTAGE;
        const uint32_t ph = (it / NSTAGE) & 1;
        bar_wait(&smem_empty[s], ph ^ 1);
        uint8_t* a_hi = smem + s * STAGE;
        uint8_t* a_lo = a_hi + A_RAW;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t v[32];
          tmem_ld32(tmem_base + lane_addr + uint32_t(cc * 32), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int row = cc * 32 + j;
            const int off = row * 128 + (((lane >> 2) ^ (row & 7)) << 4) + ((lane & 3) << 2);
            const float x = __uint_as_float(v[j]);
            const float hi = tf32_rn(x);
            *reinterpret_cast<float*>(a_hi + off) = hi;
            *reinterpret_cast<float*>(a_lo + off) = x - hi;
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) { bar_arrive(w_empty); bar_arrive(&tr_done[s]); }
      }
      // ---- D_a -> global
      float* out = G + int64_t(ti.b) * np * np;
      const int cblk = cb[q];                      // the warp's rows (c, j): column block of the direct tile
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        bar_wait(&d_full[h], tph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          uint32_t v[32];
          tmem_ld32(tmem_base + lane_addr + uint32_t(2 * PM + h * PM + half * 32), v);
          const int rblk = rb[2 * h + half];       // v[i] = G'[row rblk * 32 + i][column cblk * 32 + lane]
          // direct tile: for a fixed i the warp writes 32 consecutive floats of one row = one 128-byte line
          float* o = out + (int64_t(cblk) * np + rblk * PB) * PB + lane;
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i * PB] = __uint_as_float(v[i]);
          // mirror tile G'[cblk * 32 + lane][rblk * 32 + i]: lane = row, v = 32 consecutive columns
          if (ti.pa != ti.pc)
            staged_store_32x32(stg_base + q * STG_WARP, lane, v, out + (int64_t(rblk) * np + cblk * PB) * PB, PB, 0);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) bar_arrive(&d_empty[h]);
      }
      ++tt;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS_S) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn3 get_encode3() {
  static EncodeTiledFn3 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn3)p;
  }
  return fn;
}

}  // namespace

bool panel_sym_supported(int np) { return np % TM == 0; }

// 32-row boxes of the column-block-major G: [B * np/32 column blocks][np rows][32 floats]
int panel_sym_prepare(PanelTc* h) {
  EncodeTiledFn3 enc = get_encode3();
  R3D_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t gdim[3] = {(cuuint64_t)PB, (cuuint64_t)h->np, (cuuint64_t)(h->B * (h->np / PB))};
  const cuuint64_t gstr[2] = {(cuuint64_t)PB * 4, (cuuint64_t)h->np * PB * 4};
  const cuuint32_t box[3] = {PB, PB, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&h->map_g32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, h->G, gdim, gstr, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  R3D_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(panel 32) failed with %d", (int)r);
  static bool attr_done[kMaxDevices] = {};
  if (per_device_once(attr_done))
    R3D_CUDA(cudaFuncSetAttribute(panel_sym_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
  return 0;
}

// G <- Q^T G Q for one round, in place, one launch.
int panel_sym_update_g(PanelTc* h, int qbuf, int round, int sweep, const int* cnt, const int* qflag, cudaStream_t st) {
  const int npg = h->nt / 2;
  const int64_t total = h->B * (npg * (npg + 1) / 2);
  int grid = (int)std::min<int64_t>(total, 2 * kNumSMs);
  if (g_panel_grid_cap > 0) grid = std::min(grid, g_panel_grid_cap);
  StageScope scope(h->local ? ST_JACOBI_LOCAL : ST_JACOBI_UPDATE, st);
  // qbuf 0/1: the round-parity buffers; 2..7: the buffers of the chained schedule
  const CUtensorMap& mq = qbuf < 2 ? h->map_q[qbuf] : h->map_qc[qbuf - 2];
  panel_sym_kernel<<<grid, kSymThreads, SMEM_TOTAL, st>>>(h->map_g32, mq, h->G, (int)h->B, h->np, h->nb, h->nt,
                                                          round, sweep, h->bdiv, cnt, qflag);
  R3D_LAUNCH_CHECK();
  return 0;
}

int panel_sym_units_read(unsigned long long* out, int reset) {
  R3D_CUDA(cudaMemcpyFromSymbol(out, g_sym_units, sizeof(unsigned long long)));
  if (reset) {
    const unsigned long long z = 0;
    R3D_CUDA(cudaMemcpyToSymbol(g_sym_units, &z, sizeof(z)));
  }
  return 0;
}

}  // namespace r3d
