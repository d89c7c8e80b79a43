// Block-Jacobi panel update on 5th-generation tensor cores (sm_100a), fp32-accurate.
//
// One round of the block two-sided Jacobi applies, for every block pair (task) c of the
// round with rotation product Q_c (64x64):   G <- Q^T G Q   and   V <- V Q.
// Both are expressed with ONE GEMM shape, the column-panel update
//
//        Z[:, IJ_c]  <-  Z[:, IJ_c] * Q_c          (M = 128 rows per tile, K = 64, N = 64)
//
// because G' = (G Q)^T Q for symmetric G: pass 1 writes H^T = (G Q)^T (transposed
// store), pass 2 applies the same panel update to H^T.  V is updated in place.
//
// fp32 accuracy on TF32 tensor cores comes from the 3xTF32 split: x = hi + lo with hi exact in TF32 and
// D = A_hi B_hi + A_hi B_lo + A_lo B_hi accumulated in fp32 in TMEM (error ~2^-21 per product).  The rotation products
// Q^T arrive PRE-SPLIT from the inner solver (hi = round-to-nearest TF32, lo = Q^T - hi, two planes per block pair).
// A slab that comes from HBM is its own hi operand: kind::tf32 reads the upper 19 bits of each fp32 container, so only
// lo = rn_tf32(x - trunc_tf32(x)) is computed and written next to it.
//
// Pipeline per CTA (persistent over a strided tile list); the unit that travels through the 2-stage shared-memory
// ring is one 32-column K slab of a tile, not the tile:
//   warp 0      TMA producer: per slab one 32-column box of 128 panel rows + the hi and lo boxes of the 64 rows of Q^T
//               (all K-major SWIZZLE_128B)
//   warps 2-5   splitters: write the lo plane of the panel slab (group mode: also split P^T, which is plain fp32)
//   warp 1      MMA issuer: per slab 12 x tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=64, K=8); the eight that do not
//               read a lo plane written by the splitters are issued as soon as the TMA barrier fires
//   warps 6-9   epilogue: tcgen05.ld 32x32b.x32 -> (transposed) global stores
//   (warp 0 also allocates TMEM: 2 accumulator stages x 64 columns)
// panel_vchain_kernel (further down) chains the V updates of three XOR rounds on chip: one pass over V per three rounds.
// SASS evidence: UTCHMMA/UTCMMA (tcgen05.mma), UTMALDG (TMA), LDTM (tcgen05.ld).
#include <cuda.h>

#include <algorithm>

#include "common.cuh"
#include "jacobi_tc.cuh"
#include "tc_store.cuh"

namespace r3d {

namespace {

constexpr int PB = 32;                 // Jacobi block width (must equal JB in erank_kernels.cu)
constexpr int PM = 64;                 // panel width = 2 blocks
constexpr int TM = 128;                // rows per tile
// A pipeline stage holds ONE 32-column K slab of a tile (block I or block J of the pair): the tile's two slabs go
// through a 2-stage ring, so the TMA of the next slab runs while this one is split and multiplied, and a stage is
// held for half as long as when a stage was a whole tile (same 96 KB per CTA, two CTAs per SM).
constexpr int NSTAGE = 2;
constexpr int A_RAW = TM * PB * 4;     // 16 KB: one slab of the panel tile, hi part after the split (in place)
constexpr int Q_RAW = PM * PB * 4;     // 8 KB: the matching 32 k-columns of Q^T
constexpr int STAGE = 2 * A_RAW + 2 * Q_RAW;        // hi + lo of both operands: 48 KB
constexpr int STG_WARP = kStgWarpBytes;              // 2560 B of store staging per epilogue warp (tc_store.cuh)
constexpr int SMEM_TOTAL = NSTAGE * STAGE + 4 * STG_WARP + 1024 + 256;
constexpr int TMEM_COLS_P = 128;       // 2 accumulator stages x 64 fp32 columns
constexpr int JMAXS = 32;              // must equal JMAX_SWEEPS
// 10 warps: 0 = TMA producer + TMEM allocator, 1 = MMA issuer, 2-5 = splitters, 6-9 = epilogue.  320 threads x 72
// registers x 2 CTAs leave room (registers and shared memory) for one CTA of the inner solver on the same SM, so
// that the side-stream V update and the next inner solve can actually run together.
constexpr int kPanelThreads = 320;

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
// (32-bit MN-major operands need the special SWIZZLE_128B_BASE32B layout; the inner solver emits Q^T
//  instead, so B[n = j][k] = Qt[j][k] is K-major like the panel tile)
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

// round-to-nearest TF32 (10 explicit mantissa bits): unbiased split x = hi + lo, |lo| <= 2^-11 |x|
__device__ __forceinline__ float tf32_rn(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
// what the tensor core sees of a raw fp32 container in kind::tf32: the low 13 mantissa bits dropped
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// pairing of round r (same function as erank_kernels.cu): r >= 0 circle-method round robin over m blocks,
// r < 0 the XOR matching with mask -r (task t pairs block i with i ^ mask, i = t with a zero inserted at the
// mask's top bit)
__device__ __forceinline__ void rr_pair_tc(int m, int r, int t, int& a, int& b) {
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

// tiles actually processed (not skipped) since the last reset: [0] G passes, [1] in-place V passes -- lets bench.py
// divide the ALGORITHMIC bytes of the timed region (tiles x 64 KB) by its time instead of assuming that every launch
// of the fixed launch sequence did a full pass (launches after convergence exit at once)
// [2] tiles of the group-local 128x128 problems of the spread schedule (L2-resident working set, not HBM traffic)
__device__ unsigned long long g_panel_tiles[3];

struct PanelJob {
  // job 0 and job 1 may run in the same launch
  float* out0; float* out1;
  int transposed0, transposed1;
  int skip_on_qflag0, skip_on_qflag1;   // in-place jobs may skip tasks whose Q is the identity
  // merged mode (both passes of G <- Q^T G Q in ONE launch, see decode_merged): job 0 = pass 1 (G -> H ring),
  // job 1 = pass 2 (H ring -> G)
  int merged;
  int gs, ng, ring, lag;                // matrices per group, groups, ring slots (in groups), pass-2 lag (in groups)
  int* done1; int* done2; int* err;     // per-group completion counters (4 per tile), error flag
  // batch entries per matrix of the convergence bookkeeping (cnt / nact are indexed by b / bdiv): 1, or the number
  // of 128-column groups when the batch is the set of group-local problems of the spread schedule
  int bdiv;
  int counter;                          // which g_panel_tiles slot this launch adds to
  // group mode (SL == 4): a task is (group g, output half h) of the 4-block groups {i0, i0^ga, i0^gb, i0^ga^gb},
  // i0 = g with zero bits inserted at the pivot positions plo < phi
  int ga, gb, plo, phi;
};

// blocks of task c: SL == 2 -> the pair of the round; SL == 4 -> the four blocks of group c >> 1.  o0/o1: the two
// column blocks the task writes.
template <int SL>
__device__ __forceinline__ void task_blocks(const PanelJob& pj, int nb, int round, int c, int (&blk)[SL], int& o0,
                                            int& o1) {
  if constexpr (SL == 2) {
    rr_pair_tc(nb, round, c, blk[0], blk[1]);
    o0 = blk[0]; o1 = blk[1];
  } else {
    int x = c >> 1;
    x = ((x >> pj.plo) << (pj.plo + 1)) | (x & ((1 << pj.plo) - 1));
    x = ((x >> pj.phi) << (pj.phi + 1)) | (x & ((1 << pj.phi) - 1));
#pragma unroll
    for (int j = 0; j < 4; ++j) blk[j] = x ^ ((j & 1) ? pj.ga : 0) ^ ((j & 2) ? pj.gb : 0);
    const int h = c & 1;
    o0 = h ? blk[2] : blk[0];
    o1 = h ? blk[3] : blk[1];
  }
}

// tile index -> (job, matrix b, task c, row tile mt); returns false when the tile must be skipped
struct TileInfo { int job, b, c, mt, hb, group; bool run, valid; };

template <int SL>
__device__ __forceinline__ TileInfo decode_tile(int tile, int njobs, int B, int nt, int mtiles, int sweep,
                                                const int* __restrict__ cnt, const int* __restrict__ qflag,
                                                const PanelJob& pj) {
  TileInfo ti;
  ti.valid = true; ti.group = 0;
  if (SL == 2 && pj.merged) {
    // Merged schedule.  The batch is cut into groups of gs matrices; block s of the tile list interleaves the
    // pass-1 tiles of group s with the pass-2 tiles of group s - lag.  A pass-2 tile waits (done1) until every
    // pass-1 tile of its group has stored H; H lives in a ring of `ring` groups that stays in L2, so it never
    // travels to HBM; a pass-1 tile of group g >= ring waits (done2) for pass 2 of group g - ring before it
    // overwrites that ring slot.  Every wait targets tiles that precede the waiter in every CTA's list, and all
    // CTAs of the persistent grid are resident or will become resident, so the schedule cannot deadlock.
    const int per_mat = nt * mtiles, gsz = pj.gs * per_mat;
    const int s = tile / (2 * gsz), j = tile % (2 * gsz);
    ti.job = j & 1;
    const int idx = j >> 1;
    ti.group = ti.job == 0 ? s : s - pj.lag;
    const int bl = idx / per_mat, r = idx % per_mat;
    ti.b = ti.group * pj.gs + bl;
    ti.c = r / mtiles;
    ti.mt = r % mtiles;
    ti.valid = ti.group >= 0 && ti.group < pj.ng && ti.b < B;
    ti.hb = (ti.group % pj.ring) * pj.gs + bl;
    ti.run = ti.valid;
    if (ti.valid) {
      if (sweep > 0 && cnt[ti.b * JMAXS + sweep - 1] == 0) ti.run = false;          // matrix converged
      else {
        int any = 0;
        for (int c = 0; c < nt; ++c) any |= qflag[ti.b * nt + c];
        if (!any) ti.run = false;                                                   // nothing rotated this round
      }
    }
    return ti;
  }
  const int per_job = B * nt * mtiles;
  ti.job = tile / per_job;
  int r = tile % per_job;
  ti.b = r / (nt * mtiles);
  ti.hb = ti.b;
  r %= nt * mtiles;
  ti.c = r / mtiles;
  ti.mt = r % mtiles;
  ti.run = true;
  // flags: one per task (SL == 2) or one per group = two tasks (SL == 4)
  const int nfl = SL == 4 ? (nt >> 1) : nt, fc = SL == 4 ? (ti.c >> 1) : ti.c;
  if (sweep > 0 && cnt[(ti.b / pj.bdiv) * JMAXS + sweep - 1] == 0) ti.run = false;   // matrix converged
  else if ((ti.job == 0 ? pj.skip_on_qflag0 : pj.skip_on_qflag1)) {
    if (qflag[ti.b * nfl + fc] == 0) ti.run = false;                                 // in-place job: identity task
  } else {
    // ping-pong G passes: an identity task still has to be copied through, but if NO task of this matrix rotated
    // in this round, G is unchanged and both passes can be skipped for the whole matrix
    int any = 0;
    for (int c = 0; c < nfl; ++c) any |= qflag[ti.b * nfl + c];
    if (!any) ti.run = false;
  }
  (void)njobs;
  return ti;
}

// tiles of group g (the counters advance by 4 per tile: one per epilogue warp)
__device__ __forceinline__ int group_target(const PanelJob& pj, int g, int B, int nt, int mtiles) {
  return 4 * min(pj.gs, B - g * pj.gs) * nt * mtiles;
}
// bounded spin (about one second) on a device-scope counter; on timeout the error flag is raised and the caller
// proceeds, so that a scheduling bug shows up as a wrong result and an error code, never as a hung GPU
__device__ __forceinline__ void spin_until(const int* p, int target, int* err) {
  const long long t0 = clock64();
  for (;;) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (v >= target) break;
    if (clock64() - t0 > (1ll << 31)) { atomicExch(err, 1); break; }
    __nanosleep(100);
  }
}

template <int SL>
__global__ void __maxnreg__(72) panel_update_tc_kernel(const __grid_constant__ CUtensorMap map_in0,
                                                                 const __grid_constant__ CUtensorMap map_in1,
                                                                 const __grid_constant__ CUtensorMap map_q,
                                                                 PanelJob pj, int njobs, int B, int np, int nb, int nt,
                                                                 int round, int sweep, const int* __restrict__ cnt,
                                                                 const int* __restrict__ qflag, int debug) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)((uintptr_t(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg_base = smem + NSTAGE * STAGE;                       // epilogue staging (normal-store path)
  uint64_t* raw_full = (uint64_t*)(smem + NSTAGE * STAGE + 4 * STG_WARP);       // TMA -> splitters
  uint64_t* split_done = raw_full + NSTAGE;                      // splitters -> MMA (count 128)
  uint64_t* smem_empty = split_done + NSTAGE;                    // MMA commit -> producer
  uint64_t* tmem_full = smem_empty + NSTAGE;                     // MMA commit -> epilogue
  uint64_t* tmem_empty = tmem_full + 2;                          // epilogue -> MMA (count 128)
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  constexpr int SLABS = SL;                                      // K slabs per tile: the pair (2) or the group (4)
  if (sweep > 0 && cnt[(B / pj.bdiv) * JMAXS + sweep] == 0) return;   // every matrix converged: nothing to do
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mtiles = np / TM;
  const int total_tiles = (SL == 2 && pj.merged) ? (pj.ng + pj.lag) * 2 * pj.gs * nt * mtiles : njobs * B * nt * mtiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_in0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) { bar_init(&raw_full[s], 1); bar_init(&split_done[s], 128); bar_init(&smem_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { bar_init(&tmem_full[s], 1); bar_init(&tmem_empty[s], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)),
                 "n"(TMEM_COLS_P) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect()) {
      int it = 0, ntiles = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileInfo ti = decode_tile<SL>(tile, njobs, B, nt, mtiles, sweep, cnt, qflag, pj);
        if (!ti.run) continue;
        int blk[SL], o0, o1;
        task_blocks<SL>(pj, nb, round, ti.c, blk, o0, o1);
        if (SL == 2 && pj.merged && ti.job == 1) {
          // pass 2 reads what pass 1 of this group stored (generic proxy, other CTAs) through the async proxy
          spin_until(&pj.done1[ti.group], group_target(pj, ti.group, B, nt, mtiles), pj.err);
          asm volatile("fence.proxy.async;" ::: "memory");
        }
        const CUtensorMap* mp = ti.job == 0 ? &map_in0 : &map_in1;
        const int mb = (SL == 2 && pj.merged && ti.job == 1) ? ti.hb : ti.b;
        const int qrow = (ti.b * nt + ti.c) * PM;
#pragma unroll
        for (int slab = 0; slab < SLABS; ++slab, ++it) {
          const int s = it % NSTAGE;
          const uint32_t ph = (it / NSTAGE) & 1;
          bar_wait(&smem_empty[s], ph ^ 1);
          uint8_t* st = smem + s * STAGE;
          bar_expect_tx(&raw_full[s], A_RAW + (SL == 2 ? 2 : 1) * Q_RAW);
          // panel slab: rows [mt*128, +128) of column block I (slab 0) or J (slab 1): 16 KB contiguous in HBM
          tma_3d(st, mp, &raw_full[s], 0, ti.mt * TM, mb * nb + blk[slab]);
          // Q_c^T: 64 rows (j) x the 32 k-columns of this slab.  Pair mode: the inner solver wrote it pre-split (hi plane,
          // lo plane), both land where the MMA reads them.  Group mode: rows [64 h, 64 h + 64) of the group's 128 x 128
          // P^T (fp32, split in shared memory), qrow = (b * nt + c) * 64.
          if (SL == 2) {
            tma_2d(st + 2 * A_RAW, &map_q, &raw_full[s], slab * PB, 2 * qrow);
            tma_2d(st + 2 * A_RAW + Q_RAW, &map_q, &raw_full[s], slab * PB, 2 * qrow + PM);
          } else {
            tma_2d(st + 2 * A_RAW, &map_q, &raw_full[s], slab * PB, qrow);
          }
        }
        ++ntiles;
      }
      if (ntiles > 0) atomicAdd(&g_panel_tiles[pj.counter], (unsigned long long)ntiles);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect()) {
      int it = 0, tt = 0;                              // stage counter (slabs), tile counter (accumulators)
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileInfo ti = decode_tile<SL>(tile, njobs, B, nt, mtiles, sweep, cnt, qflag, pj);
        if (!ti.run) continue;
        const int acc = tt & 1;
        const uint32_t aph = (tt >> 1) & 1;
        bar_wait(&tmem_empty[acc], aph ^ 1);           // epilogue has drained this accumulator
        const uint32_t d = tmem_base + acc * PM;
        uint32_t first = 0;
#pragma unroll
        for (int slab = 0; slab < SLABS; ++slab, ++it) {
          const int s = it % NSTAGE;
          const uint32_t ph = (it / NSTAGE) & 1;
          // pair mode: the raw A slab is the hi operand (kind::tf32 reads the upper 19 bits of the container) and Q^T
          // arrives pre-split, so two of the three products start when the TMA has landed; the lo product waits for the
          // splitters.  Group mode: P^T is split in shared memory (hi rewritten), everything waits for the splitters.
          bar_wait(SL == 2 ? &raw_full[s] : &split_done[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = s_u32(smem + s * STAGE), a_lo = a_hi + A_RAW;
          const uint32_t q_hi = a_hi + 2 * A_RAW, q_lo = q_hi + Q_RAW;
#pragma unroll
          for (int prod = 0; prod < 3; ++prod) {
            if ((debug & 1) && prod > 0) break;
            if (debug & 16) break;                    // timing experiment: no MMAs at all
            if (SL == 2 && prod == 2) {
              bar_wait(&split_done[s], ph);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const uint32_t ab = (prod == 2) ? a_lo : a_hi;
            const uint32_t qb = (prod == 1) ? q_lo : q_hi;
#pragma unroll
            for (int kk = 0; kk < PB / 8; ++kk) {
              // both operands K-major: 32 B (8 tf32) per step inside the 128 B swizzle row of the slab
              const uint64_t ad = desc_sw128(ab + kk * 32, 16, 1024);
              const uint64_t bd = desc_sw128(qb + kk * 32, 16, 1024);
              umma_tf32(d, ad, bd, first);
              first = 1;
            }
          }
          umma_commit_to(&smem_empty[s]);
        }
        umma_commit_to(&tmem_full[acc]);
        ++tt;
      }
    }
  } else if (warp >= 2 && warp < 6) {
    // ===================== splitters: x -> (hi, lo) in shared memory =====================
    const int t = threadIdx.x - 64;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileInfo ti = decode_tile<SL>(tile, njobs, B, nt, mtiles, sweep, cnt, qflag, pj);
      if (!ti.run) continue;
#pragma unroll 1
      for (int slab = 0; slab < SLABS; ++slab, ++it) {
        const int s = it % NSTAGE;
        const uint32_t ph = (it / NSTAGE) & 1;
        bar_wait(&raw_full[s], ph);
        uint8_t* st = smem + s * STAGE;
        float4* a_hi = reinterpret_cast<float4*>(st);
        float4* a_lo = reinterpret_cast<float4*>(st + A_RAW);
        float4* q_hi = reinterpret_cast<float4*>(st + 2 * A_RAW);
        float4* q_lo = reinterpret_cast<float4*>(st + 2 * A_RAW + Q_RAW);
        if (!(debug & 4)) {
#pragma unroll 4
          for (int e = t; e < A_RAW / 16; e += 128) {
            const float4 x = a_hi[e];
            if (SL == 2) {                             // raw slab = hi operand: write lo only
              float4 l;
              l.x = tf32_rn(x.x - tf32_trunc(x.x));
              l.y = tf32_rn(x.y - tf32_trunc(x.y));
              l.z = tf32_rn(x.z - tf32_trunc(x.z));
              l.w = tf32_rn(x.w - tf32_trunc(x.w));
              a_lo[e] = l;
            } else {
              float4 h, l;
              h.x = tf32_rn(x.x); l.x = x.x - h.x;
              h.y = tf32_rn(x.y); l.y = x.y - h.y;
              h.z = tf32_rn(x.z); l.z = x.z - h.z;
              h.w = tf32_rn(x.w); l.w = x.w - h.w;
              a_hi[e] = h; a_lo[e] = l;
            }
          }
          if (SL != 2) {
#pragma unroll 4
            for (int e = t; e < Q_RAW / 16; e += 128) {
              const float4 x = q_hi[e];
              float4 h, l;
              h.x = tf32_rn(x.x); l.x = x.x - h.x;
              h.y = tf32_rn(x.y); l.y = x.y - h.y;
              h.z = tf32_rn(x.z); l.z = x.z - h.z;
              h.w = tf32_rn(x.w); l.w = x.w - h.w;
              q_hi[e] = h; q_lo[e] = l;
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to UMMA
        bar_arrive(&split_done[s]);
      }
    }
  } else if (warp >= 6) {
    // ===================== epilogue: TMEM -> global =====================
    const int q = warp & 3;                      // TMEM lanes [32q, 32q+32)
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileInfo ti = decode_tile<SL>(tile, njobs, B, nt, mtiles, sweep, cnt, qflag, pj);
      if (!ti.run) {
        // merged mode: a skipped tile still counts as done for the waiters of its group
        if (SL == 2 && pj.merged && ti.valid && lane == 0) atomicAdd(ti.job == 0 ? &pj.done1[ti.group] : &pj.done2[ti.group], 1);
        continue;
      }
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      bar_wait(&tmem_full[acc], aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      int blk[SL], I, J;
      task_blocks<SL>(pj, nb, round, ti.c, blk, I, J);
      if (SL == 2 && pj.merged && ti.job == 0 && ti.group >= pj.ring) {
        // the ring slot is reused: pass 2 of group - ring must have read it
        if (lane == 0) spin_until(&pj.done2[ti.group - pj.ring], group_target(pj, ti.group - pj.ring, B, nt, mtiles), pj.err);
        __syncwarp();
      }
      float* out = (ti.job == 0 ? pj.out0 : pj.out1) +
                   int64_t((SL == 2 && pj.merged && ti.job == 0) ? ti.hb : ti.b) * np * np;
      const bool tr = (ti.job == 0 ? pj.transposed0 : pj.transposed1) != 0;
      const int row = ti.mt * TM + q * 32 + lane;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * PM + half * 32);
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
        const int cblk = (half == 0 ? I : J);
        if (SL == 2 && (debug & 2)) {
          // dump the staged panel tile as the async proxy left it (after the split): element (row r, k = half*32 + j)
          const uint8_t* abase = smem + half * STAGE;      // slab `half` of this tile sits in stage `half` (NSTAGE == SLABS)
          const int r = q * 32 + lane;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int chunk = (j >> 2) ^ (r & 7);
            v[j] = *reinterpret_cast<const uint32_t*>(abase + r * 128 + chunk * 16 + (j & 3) * 4);
          }
        }
        if (debug & 8) continue;                      // timing experiment: no global stores
        if (tr) {
          // out[row' = cblk*32 + j][col' = row]: column block row/32 (warp uniform), lane = col' % 32, so the
          // warp writes 32 consecutive 128-byte rows = one contiguous 4 KB run
          float* o = out + (int64_t(row >> 5) * np + cblk * PB) * PB + (row & 31);
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j * PB] = __uint_as_float(v[j]);
        } else {
          // out[row][cblk*32 .. +32) for the warp's 32 rows = one contiguous 4 KB run.  A lane owns a row, so a
          // direct store would touch 32 different lines per instruction; stage 16 columns at a time through
          // shared memory and write 64-byte row segments with 4 lanes each (full sectors, 8 rows per instruction).
          staged_store_32x32(stg_base + q * STG_WARP, lane, v, out + (int64_t(cblk) * np + (row - lane)) * PB, PB, 0);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      bar_arrive(&tmem_empty[acc]);
      if (SL == 2 && pj.merged) {
        // publish this warp's stores (pass 1) / its tile's completed reads (pass 2) at device scope
        asm volatile("fence.proxy.async;" ::: "memory");
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(ti.job == 0 ? &pj.done1[ti.group] : &pj.done2[ti.group], 1);
      }
      ++it;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS_P) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Chained V update (jacobi_schedule = 2).  With the XOR ordering, three consecutive rounds {a, b, a^b} only pair blocks
// inside the 4-block cosets of the subspace <a, b> (jacobi_tc.cuh: PanelGroups), so  V <- V Q1 Q2 Q3  restricted to a
// coset is a product of three 128 x 128 block-diagonal-in-pairs matrices.  One tile = 128 rows x the 4 blocks of a
// coset: loaded ONCE (64 KB), multiplied by the three rounds' rotations with the intermediate results staying on chip
// (TMEM -> registers -> hi/lo split -> shared-memory operand slabs), stored ONCE: a third of the HBM traffic of three
// separate passes.  G keeps its per-round update (the inner solver needs the updated diagonal blocks every round).
//
// A tile makes 12 uses of the 2-stage ring, use u = 4 k + 2 p + s: round k, pair p of the coset, K slab s of the pair
// (M = 128, N = 64, K = 32 per use, 3xTF32).  Round 0 gets its A slabs by TMA, rounds 1-2 from the drainers (warps 6-9),
// which read the previous round's accumulator; the accumulators ping-pong between two 128-column TMEM buffers.
// ------------------------------------------------------------------------------------------------------------------
struct ChainJob {
  int ga, gb, plo, phi;                 // the coset structure of the three rounds (masks ga, gb, ga ^ gb)
  const int* qflag[3];                  // task flags of the three rounds
};
// per tile: sig = for each round the local block index (2 bits each) at accumulator column positions 0..3
// ([pair 0: I, J | pair 1: I, J]); task = the rounds' task indices of the two pairs (8 bits each)
struct ChainPlan { int x; uint32_t sig; uint64_t task; };

__host__ __device__ __forceinline__ int top_bit(int v) {
#ifdef __CUDA_ARCH__
  return 31 - __clz(v);
#else
  return 31 - __builtin_clz((unsigned)v);
#endif
}
__host__ __device__ __forceinline__ int chain_blk(const ChainJob& cj, int x, int j) {
  return x ^ ((j & 1) ? cj.ga : 0) ^ ((j & 2) ? cj.gb : 0);
}
__host__ __device__ __forceinline__ ChainPlan chain_plan(const ChainJob& cj, int g) {
  ChainPlan cp;
  int x = g;
  x = ((x >> cj.plo) << (cj.plo + 1)) | (x & ((1 << cj.plo) - 1));
  x = ((x >> cj.phi) << (cj.phi + 1)) | (x & ((1 << cj.phi) - 1));
  cp.x = x; cp.sig = 0; cp.task = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int lm = k + 1;
    const int mk = ((lm & 1) ? cj.ga : 0) ^ ((lm & 2) ? cj.gb : 0);
    const int hb = top_bit(mk);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int j = p == 0 ? 0 : (lm == 1 ? 2 : 1), jj = j ^ lm;
      const int sel = (chain_blk(cj, x, j) >> hb) & 1;            // the block with the mask's top bit clear comes first
      const int il = sel ? jj : j, jl = sel ? j : jj;
      cp.sig |= uint32_t(il | (jl << 2)) << (8 * k + 4 * p);
      const int I = chain_blk(cj, x, il);
      const int t = ((I >> (hb + 1)) << hb) | (I & ((1 << hb) - 1));
      cp.task |= uint64_t(t & 0xff) << (16 * k + 8 * p);
    }
  }
  return cp;
}
__host__ __device__ __forceinline__ int chain_sig(const ChainPlan& cp, int k, int pos) { return (cp.sig >> (8 * k + 2 * pos)) & 3; }
__host__ __device__ __forceinline__ int chain_task(const ChainPlan& cp, int k, int p) { return int((cp.task >> (16 * k + 8 * p)) & 0xff); }
// accumulator column position of local block l after round k
__host__ __device__ __forceinline__ int chain_pos(const ChainPlan& cp, int k, int l) {
  int pos = 0;
#pragma unroll
  for (int q = 1; q < 4; ++q) if (chain_sig(cp, k, q) == l) pos = q;
  return pos;
}

struct ChainTile { int b, g, mt; bool run; ChainPlan cp; };
__device__ __forceinline__ ChainTile decode_chain(int tile, int ng, int nt, int mtiles, int sweep,
                                                  const int* __restrict__ cnt, const ChainJob& cj) {
  ChainTile t;
  t.b = tile / (ng * mtiles);
  const int r = tile - t.b * ng * mtiles;
  t.g = r / mtiles;
  t.mt = r - t.g * mtiles;
  t.run = !(sweep > 0 && cnt[t.b * JMAXS + sweep - 1] == 0);
  if (t.run) {
    t.cp = chain_plan(cj, t.g);
    int any = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int p = 0; p < 2; ++p) any |= cj.qflag[k][t.b * nt + chain_task(t.cp, k, p)];
    t.run = any != 0;
  }
  return t;
}

constexpr int TMEM_COLS_C = 256;       // two 128-column accumulators
constexpr int CUSES = 12;

__device__ __forceinline__ void tmem_ld32c(uint32_t taddr, uint32_t (&v)[32]) {
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

__global__ void __maxnreg__(72) panel_vchain_kernel(const __grid_constant__ CUtensorMap map_v,
                                                    const __grid_constant__ CUtensorMap map_q0,
                                                    const __grid_constant__ CUtensorMap map_q1,
                                                    const __grid_constant__ CUtensorMap map_q2, float* __restrict__ V,
                                                    ChainJob cj, int B, int np, int nb, int nt, int sweep,
                                                    const int* __restrict__ cnt) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)((uintptr_t(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg_base = smem + NSTAGE * STAGE;
  uint64_t* raw_full = (uint64_t*)(smem + NSTAGE * STAGE + 4 * STG_WARP);   // TMA -> splitters
  uint64_t* split_done = raw_full + NSTAGE;      // round 0: 128 splitters (A slab split) -> MMA
  uint64_t* drain_done = split_done + NSTAGE;    // rounds 1-2: 4 drainer warps (A slab written) -> MMA
  uint64_t* q_full = drain_done + NSTAGE;        // rounds 1-2: TMA (pre-split Q^T planes) -> MMA.  A barrier of its own:
                                                 // every waiter of a barrier must see each of its phases (a waiter that
                                                 // skips phases can match a stale phase of the same parity)
  uint64_t* smem_empty = q_full + NSTAGE;        // MMA commit -> producer / drainers
  uint64_t* acc_full = smem_empty + NSTAGE;      // [2] MMA commit (a round is complete) -> drainers / epilogue
  uint64_t* acc_empty = acc_full + 2;            // [2] drainers / epilogue (4 warps) -> MMA
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);

  if (sweep > 0 && cnt[B * JMAXS + sweep] == 0) return;          // every matrix converged
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mtiles = np / TM, ng = nb >> 2;
  const int total = B * ng * mtiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q2) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      bar_init(&raw_full[s], 1); bar_init(&split_done[s], 128); bar_init(&drain_done[s], 4); bar_init(&smem_empty[s], 1);
      bar_init(&q_full[s], 1);
    }
    for (int s = 0; s < 2; ++s) { bar_init(&acc_full[s], 1); bar_init(&acc_empty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)),
                 "n"(TMEM_COLS_C) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect()) {
      int it = 0, ntiles = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const ChainTile ti = decode_chain(tile, ng, nt, mtiles, sweep, cnt, cj);
        if (!ti.run) continue;
#pragma unroll 1
        for (int u = 0; u < CUSES; ++u, ++it) {
          const int k = u >> 2, p = (u >> 1) & 1, sl = u & 1;
          const int s = it % NSTAGE;
          const uint32_t ph = (it / NSTAGE) & 1;
          bar_wait(&smem_empty[s], ph ^ 1);
          uint8_t* st = smem + s * STAGE;
          const CUtensorMap* mq = k == 0 ? &map_q0 : (k == 1 ? &map_q1 : &map_q2);
          const int qrow = (ti.b * nt + chain_task(ti.cp, k, p)) * 2 * PM;      // hi plane; lo plane 64 rows below
          uint64_t* fb = k == 0 ? &raw_full[s] : &q_full[s];
          if (k == 0) {
            bar_expect_tx(fb, A_RAW + 2 * Q_RAW);
            const int cb = chain_blk(cj, ti.cp.x, chain_sig(ti.cp, 0, 2 * p + sl));
            tma_3d(st, &map_v, fb, 0, ti.mt * TM, ti.b * nb + cb);
          } else {
            bar_expect_tx(fb, 2 * Q_RAW);
          }
          tma_2d(st + 2 * A_RAW, mq, fb, sl * PB, qrow);
          tma_2d(st + 2 * A_RAW + Q_RAW, mq, fb, sl * PB, qrow + PM);
        }
        ++ntiles;
      }
      // one chained tile moves 64 KB in + 64 KB out: two units of the V counter
      if (ntiles > 0) atomicAdd(&g_panel_tiles[1], (unsigned long long)(2 * ntiles));
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect()) {
      int it = 0, tt = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const ChainTile ti = decode_chain(tile, ng, nt, mtiles, sweep, cnt, cj);
        if (!ti.run) continue;
#pragma unroll 1
        for (int k = 0; k < 3; ++k) {
          const int buf = k & 1;
          const int n = buf == 0 ? 2 * tt + (k >> 1) : tt;       // how often this accumulator was produced before
          bar_wait(&acc_empty[buf], (n & 1) ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
          for (int ps = 0; ps < 4; ++ps, ++it) {
            const int s = it % NSTAGE;
            const uint32_t ph = (it / NSTAGE) & 1;
            // per tile every stage sees 2 uses of round 0 and 4 of rounds 1-2 (12 uses, 2 stages): the phase parities of
            // the two hand-over barriers follow from the use index alone
            if (k == 0) bar_wait(&raw_full[s], (ps >> 1) & 1);                   // raw A slab (= hi operand) + Q planes landed
            else {
              bar_wait(&q_full[s], ((4 * (k - 1) + ps) >> 1) & 1);               // pre-split Q^T planes landed (TMA)
              bar_wait(&drain_done[s], ((4 * (k - 1) + ps) >> 1) & 1);           // A slab written by the drainers
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d = tmem_base + buf * 128 + (ps >> 1) * PM;
            const uint32_t a_hi = s_u32(smem + s * STAGE), a_lo = a_hi + A_RAW;
            const uint32_t q_hi = a_hi + 2 * A_RAW, q_lo = q_hi + Q_RAW;
            uint32_t first = ps & 1;                             // the pair's accumulator takes two K slabs
#pragma unroll
            for (int prod = 0; prod < 3; ++prod) {
              if (prod == 2 && k == 0) {                         // A_lo of a TMA-fed slab comes from the splitters
                bar_wait(&split_done[s], (ps >> 1) & 1);
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
          }
          umma_commit_to(&acc_full[buf]);
        }
        ++tt;
      }
    }
  } else if (warp >= 2 && warp < 6) {
    // ===================== splitters =====================
    const int t = threadIdx.x - 64;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const ChainTile ti = decode_chain(tile, ng, nt, mtiles, sweep, cnt, cj);
      if (!ti.run) continue;
      // only the four A slabs of round 0 need splitting (Q^T arrives pre-split; rounds 1-2 get A from the drainers)
#pragma unroll 1
      for (int u = 0; u < 4; ++u) {
        const int s = u % NSTAGE;                  // a tile starts on stage 0 (12 uses per tile)
        bar_wait(&raw_full[s], (u >> 1) & 1);      // raw_full completes twice per tile and stage
        uint8_t* st = smem + s * STAGE;
        float4* a_hi = reinterpret_cast<float4*>(st);
        float4* a_lo = reinterpret_cast<float4*>(st + A_RAW);
        // kind::tf32 reads the upper 19 bits of each fp32 container, so the raw slab IS the hi operand; only
        // lo = rn_tf32(x - trunc_tf32(x)) is written, and the two products without lo start when the TMA has landed
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
    // ===================== drainers (rounds 1, 2) + epilogue =====================
    const int q = warp & 3;                        // TMEM lanes [32q, 32q + 32) = tile rows
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const int row = q * 32 + lane;                 // operand row of this lane
    int tt = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const ChainTile ti = decode_chain(tile, ng, nt, mtiles, sweep, cnt, cj);
      if (!ti.run) continue;
#pragma unroll 1
      for (int k = 1; k < 3; ++k) {
        const int src = (k - 1) & 1;               // accumulator of the previous round
        const int n = src == 0 ? 2 * tt : tt;      // its production index (round 0 -> buffer 0, round 1 -> buffer 1)
        bar_wait(&acc_full[src], n & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int ps = 0; ps < 4; ++ps) {
          const int it = tt * CUSES + k * 4 + ps;
          const int s = it % NSTAGE;
          const uint32_t ph = (it / NSTAGE) & 1;
          // K slab ps of round k = local block sig[k][ps] = column position pos of the previous accumulator
          const int pos = chain_pos(ti.cp, k - 1, chain_sig(ti.cp, k, ps));
          uint32_t v[32];
          tmem_ld32c(tmem_base + lane_addr + uint32_t(src * 128 + pos * 32), v);
          bar_wait(&smem_empty[s], ph ^ 1);
          uint8_t* a_hi = smem + s * STAGE + row * 128;
          uint8_t* a_lo = a_hi + A_RAW;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float4 h, l;
            const float x0 = __uint_as_float(v[4 * c]), x1 = __uint_as_float(v[4 * c + 1]);
            const float x2 = __uint_as_float(v[4 * c + 2]), x3 = __uint_as_float(v[4 * c + 3]);
            h.x = tf32_rn(x0); l.x = x0 - h.x;
            h.y = tf32_rn(x1); l.y = x1 - h.y;
            h.z = tf32_rn(x2); l.z = x2 - h.z;
            h.w = tf32_rn(x3); l.w = x3 - h.w;
            const int off = (c ^ (row & 7)) << 4;
            *reinterpret_cast<float4*>(a_hi + off) = h;
            *reinterpret_cast<float4*>(a_lo + off) = l;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) bar_arrive(&drain_done[s]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) bar_arrive(&acc_empty[src]);
      }
      // ---- round 2's accumulator (buffer 0, second production of this tile) -> V, in place
      bar_wait(&acc_full[0], (2 * tt + 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float* out = V + int64_t(ti.b) * np * np;
#pragma unroll 1
      for (int pos = 0; pos < 4; ++pos) {
        uint32_t v[32];
        tmem_ld32c(tmem_base + lane_addr + uint32_t(pos * 32), v);
        const int cb = chain_blk(cj, ti.cp.x, chain_sig(ti.cp, 2, pos));
        staged_store_32x32(stg_base + q * STG_WARP, lane, v, out + (int64_t(cb) * np + ti.mt * TM + q * 32) * PB, PB, 0);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) bar_arrive(&acc_empty[0]);
      ++tt;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS_C) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn2 get_encode2() {
  static EncodeTiledFn2 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn2)p;
  }
  return fn;
}

int make_map_panel(CUtensorMap* m, const float* base, int64_t B, int np) {
  EncodeTiledFn2 enc = get_encode2();
  R3D_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  // column-block-major storage: [B * np/32 column blocks][np rows][32 floats]
  const cuuint64_t gdim[3] = {(cuuint64_t)PB, (cuuint64_t)np, (cuuint64_t)(B * (np / PB))};
  const cuuint64_t gstr[2] = {(cuuint64_t)PB * 4, (cuuint64_t)np * PB * 4};
  const cuuint32_t box[3] = {PB, TM, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  R3D_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(panel) failed with %d", (int)r);
  return 0;
}
// rows of `pitch` floats (64: the Q^T of a block pair; 128: the P^T of a 4-block group); a box is 64 rows x 32 k-columns
int make_map_q(CUtensorMap* m, const float* base, int64_t rows, int pitch = PM) {
  EncodeTiledFn2 enc = get_encode2();
  R3D_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)pitch * 4};
  const cuuint32_t box[2] = {32, PM};
  const cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  R3D_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(Q) failed with %d", (int)r);
  return 0;
}

}  // namespace

int g_panel_debug = 0;
int g_panel_grid_cap = 0;

bool panel_tc_supported(int np) { return np % TM == 0; }

int panel_tc_prepare(PanelTc* h, float* G, float* H, float* V, const float* Qb0, const float* Qb1, int64_t B,
                     int np) {
  h->G = G; h->H = H; h->V = V; h->B = B; h->np = np; h->nb = np / PB; h->nt = np / PM;
  h->bdiv = 1; h->local = 0;
  if (int e = make_map_panel(&h->map_g, G, B, np)) return e;
  if (H != nullptr) { if (int e = make_map_panel(&h->map_h, H, B, np)) return e; }
  if (int e = make_map_panel(&h->map_v, V, B, np)) return e;
  // Q^T buffers are pre-split by the inner solver: per task a hi plane and a lo plane of 64 rows each
  if (int e = make_map_q(&h->map_q[0], Qb0, B * h->nt * 2 * PM)) return e;
  if (int e = make_map_q(&h->map_q[1], Qb1, B * h->nt * 2 * PM)) return e;
  static bool attr_done[kMaxDevices] = {};
  if (per_device_once(attr_done)) {
    R3D_CUDA(cudaFuncSetAttribute(panel_update_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    R3D_CUDA(cudaFuncSetAttribute(panel_update_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    R3D_CUDA(cudaFuncSetAttribute(panel_vchain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
  }
  return 0;
}

// P^T of the 4-block groups of the spread schedule: (B * np/128 groups, 128, 128) fp32 row-major, double-buffered
int panel_tc_prepare_groups(PanelTc* h, const float* Pt0, const float* Pt1) {
  const int64_t rows = h->B * (h->np / 128) * 128;
  if (int e = make_map_q(&h->map_p[0], Pt0, rows, 128)) return e;
  return make_map_q(&h->map_p[1], Pt1, rows, 128);
}

// Q^T buffers of the chained schedule: two slots of three rounds each
int panel_tc_prepare_chain(PanelTc* h, float* const Qc[6]) {
  for (int i = 0; i < 6; ++i)
    if (int e = make_map_q(&h->map_qc[i], Qc[i], h->B * h->nt * 2 * PM)) return e;
  return 0;
}

// Host mirror of the tile bookkeeping of panel_vchain_kernel for the CPU tests: group g of the coset structure `grp` ->
// out[0..3] the four global blocks, out[4 + 4 k + pos] the local block at accumulator position pos in round k,
// out[16 + 2 k + p] the task index (in the round's Q^T buffer) of pair p in round k.
void panel_chain_plan_host(const PanelGroups& grp, int g, int out[22]) {
  ChainJob cj{};
  cj.ga = grp.ga; cj.gb = grp.gb; cj.plo = grp.plo; cj.phi = grp.phi;
  const ChainPlan cp = chain_plan(cj, g);
  for (int j = 0; j < 4; ++j) out[j] = chain_blk(cj, cp.x, j);
  for (int k = 0; k < 3; ++k) {
    for (int pos = 0; pos < 4; ++pos) out[4 + 4 * k + pos] = chain_sig(cp, k, pos);
    for (int p = 0; p < 2; ++p) out[16 + 2 * k + p] = chain_task(cp, k, p);
  }
}

bool panel_chain_supported(int np) {
  const int nb = np / PB;
  return np % TM == 0 && nb >= 8 && (nb & (nb - 1)) == 0 && nb / 2 <= 256;
}

// V <- V Q1 Q2 Q3 for the three rounds of a super-round (masks ga, gb, ga ^ gb), one launch, one pass over V.
int panel_tc_update_v_chain(PanelTc* h, int slot, const PanelGroups& grp, int sweep, const int* cnt,
                            const int* const qflag[3], cudaStream_t st) {
  const int mtiles = h->np / TM;
  const int64_t tiles = h->B * (h->nb / 4) * mtiles;
  ChainJob cj;
  cj.ga = grp.ga; cj.gb = grp.gb; cj.plo = grp.plo; cj.phi = grp.phi;
  for (int k = 0; k < 3; ++k) cj.qflag[k] = qflag[k];
  int grid = (int)std::min<int64_t>(tiles, 2 * kNumSMs);
  if (g_panel_grid_cap > 0) grid = std::min(grid, g_panel_grid_cap);
  StageScope scope(ST_JACOBI_VUPDATE, st);
  panel_vchain_kernel<<<grid, kPanelThreads, SMEM_TOTAL, st>>>(h->map_v, h->map_qc[3 * slot], h->map_qc[3 * slot + 1],
                                                               h->map_qc[3 * slot + 2], h->V, cj, (int)h->B, h->np, h->nb,
                                                               h->nt, sweep, cnt);
  R3D_LAUNCH_CHECK();
  return 0;
}

int panel_tiles_read(unsigned long long out[3], int reset) {
  R3D_CUDA(cudaMemcpyFromSymbol(out, g_panel_tiles, sizeof(unsigned long long) * 3));
  if (reset) {
    const unsigned long long z[3] = {0, 0, 0};
    R3D_CUDA(cudaMemcpyToSymbol(g_panel_tiles, z, sizeof(z)));
  }
  return 0;
}

static int panel_launch(PanelTc* h, const CUtensorMap& in, const CUtensorMap& q, float* out, int transposed,
                        int skip_on_qflag, int round, int sweep, const int* cnt, const int* qflag, int stage_id,
                        cudaStream_t st, const PanelGroups* grp = nullptr) {
  const int mtiles = h->np / TM;
  const int64_t tiles = h->B * h->nt * mtiles;
  PanelJob pj{};
  pj.out0 = out; pj.transposed0 = transposed; pj.skip_on_qflag0 = skip_on_qflag;
  pj.out1 = nullptr; pj.transposed1 = 0; pj.skip_on_qflag1 = 0;
  pj.bdiv = h->bdiv;
  pj.counter = h->local ? 2 : (skip_on_qflag ? 1 : 0);
  int grid = (int)std::min<int64_t>(tiles, 2 * kNumSMs);
  if (g_panel_grid_cap > 0) grid = std::min(grid, g_panel_grid_cap);
  StageScope scope(stage_id, st);
  if (grp != nullptr) {
    pj.ga = grp->ga; pj.gb = grp->gb; pj.plo = grp->plo; pj.phi = grp->phi;
    panel_update_tc_kernel<4><<<grid, kPanelThreads, SMEM_TOTAL, st>>>(in, in, q, pj, 1, (int)h->B, h->np, h->nb, h->nt, 0,
                                                                       sweep, cnt, qflag, g_panel_debug);
  } else {
    panel_update_tc_kernel<2><<<grid, kPanelThreads, SMEM_TOTAL, st>>>(in, in, q, pj, 1, (int)h->B, h->np, h->nb, h->nt,
                                                                       round, sweep, cnt, qflag, g_panel_debug);
  }
  R3D_LAUNCH_CHECK();
  return 0;
}

// Both passes in one launch (merged schedule, H ring resident in L2).  `sync` holds 2 * kPanelSyncGroups counters
// + 1 error flag, all zero when the launch starts (the inner solver of the round clears them).
static int panel_launch_merged(PanelTc* h, const CUtensorMap& q, int round, int sweep, const int* cnt,
                               const int* qflag, int* sync, cudaStream_t st) {
  const int mtiles = h->np / TM;
  const int per_mat = h->nt * mtiles;
  PanelJob pj{};
  pj.out0 = h->H; pj.transposed0 = 1;
  pj.out1 = h->G; pj.transposed1 = 1;
  pj.merged = 1;
  const int64_t mat_bytes = int64_t(h->np) * h->np * 4;
  pj.gs = (int)std::max<int64_t>(1, std::min<int64_t>(h->B, (int64_t(options().panel_group_mb) << 20) / mat_bytes));
  pj.ng = (int)((h->B + pj.gs - 1) / pj.gs);
  pj.ring = std::min(pj.ng, std::max(3, options().panel_ring));
  pj.lag = 2;
  if (pj.ring <= pj.lag) pj.ring = pj.ng;             // tiny batches: no slot reuse at all
  pj.done1 = sync; pj.done2 = sync + kPanelSyncGroups; pj.err = sync + 2 * kPanelSyncGroups;
  pj.bdiv = 1; pj.counter = 0;
  const int64_t tiles = int64_t(pj.ng + pj.lag) * 2 * pj.gs * per_mat;
  int grid = (int)std::min<int64_t>(tiles, 2 * kNumSMs);
  if (g_panel_grid_cap > 0) grid = std::min(grid, g_panel_grid_cap);
  StageScope scope(ST_JACOBI_UPDATE, st);
  panel_update_tc_kernel<2><<<grid, kPanelThreads, SMEM_TOTAL, st>>>(h->map_g, h->map_h, q, pj, 2, (int)h->B, h->np, h->nb, h->nt,
                                                        round, sweep, cnt, qflag, g_panel_debug);
  R3D_LAUNCH_CHECK();
  return 0;
}

bool panel_tc_merged_ok(const PanelTc* h) {
  if (options().panel_merged == 0 || h->bdiv != 1) return false;
  const int64_t mat_bytes = int64_t(h->np) * h->np * 4;
  const int64_t gs = std::max<int64_t>(1, std::min<int64_t>(h->B, (int64_t(options().panel_group_mb) << 20) / mat_bytes));
  return (h->B + gs - 1) / gs <= kPanelSyncGroups;
}

int panel_tc_update_g(PanelTc* h, int qbuf, int round, int sweep, const int* cnt, const int* qflag, cudaStream_t st,
                      int* sync) {
  if (sync != nullptr && panel_tc_merged_ok(h))
    return panel_launch_merged(h, h->map_q[qbuf], round, sweep, cnt, qflag, sync, st);
  // pass 1: H^T = (G Q)^T (transposed store);  pass 2: G = H^T Q.  Identity tasks cannot be skipped (ping-pong).
  const int sid = h->local ? ST_JACOBI_LOCAL : ST_JACOBI_UPDATE;
  if (int e = panel_launch(h, h->map_g, h->map_q[qbuf], h->H, 1, 0, round, sweep, cnt, qflag, sid, st))
    return e;
  // G' is symmetric, so pass 2 may also use the transposed store (one full 128-byte line per store instruction)
  return panel_launch(h, h->map_h, h->map_q[qbuf], h->G, 1, 0, round, sweep, cnt, qflag, sid, st);
}

int panel_tc_update_v(PanelTc* h, int qbuf, int round, int sweep, const int* cnt, const int* qflag, cudaStream_t st) {
  return panel_launch(h, h->map_v, h->map_q[qbuf], h->V, 0, 1, round, sweep, cnt, qflag,
                      h->local ? ST_JACOBI_LOCAL : ST_JACOBI_VUPDATE, st);
}

// The same two updates with the 128 x 128 group products P of a super-round (K = N = 128): `gflag` holds one flag per
// (matrix, group).
int panel_tc_update_g_groups(PanelTc* h, int pbuf, const PanelGroups& grp, int sweep, const int* cnt, const int* gflag,
                             cudaStream_t st) {
  if (int e = panel_launch(h, h->map_g, h->map_p[pbuf], h->H, 1, 0, 0, sweep, cnt, gflag, ST_JACOBI_UPDATE, st, &grp))
    return e;
  return panel_launch(h, h->map_h, h->map_p[pbuf], h->G, 1, 0, 0, sweep, cnt, gflag, ST_JACOBI_UPDATE, st, &grp);
}
int panel_tc_update_v_groups(PanelTc* h, int pbuf, const PanelGroups& grp, int sweep, const int* cnt, const int* gflag,
                             cudaStream_t st) {
  return panel_launch(h, h->map_v, h->map_p[pbuf], h->V, 0, 1, 0, sweep, cnt, gflag, ST_JACOBI_VUPDATE, st, &grp);
}

}  // namespace r3d
