// Host-side handle of the tensor-core panel update (jacobi_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace r3d {
struct PanelTc {
  CUtensorMap map_g, map_h, map_v, map_q[2];   // Q^T is double-buffered (round parity)
  CUtensorMap map_p[2];                        // spread schedule: P^T of the 4-block groups, double-buffered
  CUtensorMap map_g32;                         // one-pass symmetric update: 32-row boxes of G (jacobi_sym.cu)
  CUtensorMap map_qc[6];                       // chained schedule: Q^T of two slots x three rounds (qbuf = 2 + 3 slot + k)
  float *G, *H, *V;
  int64_t B;
  int np, nb, nt;
  int bdiv;     // batch entries per matrix of the convergence bookkeeping (> 1: the batch is the group-local problems)
  int local;    // 1: group-local problem (stage / tile counters are booked separately)
};
// One super-round of the spread schedule: the block index set is partitioned into the cosets {i0, i0^ga, i0^gb,
// i0^ga^gb} of a two-dimensional subspace of GF(2)^k; plo < phi are the pivot bit positions of the subspace.
struct PanelGroups { int ga, gb, plo, phi; };
bool panel_tc_supported(int np);
int panel_tc_prepare(PanelTc* h, float* G, float* H, float* V, const float* Qb0, const float* Qb1, int64_t B, int np);
// G <- Q^T G Q for one round (two launches on `st`): pass 1 writes (G Q)^T into H, pass 2 H Q back into G.
// With `sync` (2 * kPanelSyncGroups + 1 zeroed ints) both passes run in ONE launch on a merged schedule that keeps H in L2.
constexpr int kPanelSyncGroups = 512;
int panel_tc_update_g(PanelTc* h, int qbuf, int round, int sweep, const int* cnt, const int* qflag, cudaStream_t st,
                      int* sync = nullptr);
// V <- V Q for one round (one launch on `st`); independent of the G update and of the next inner solve.
int panel_tc_update_v(PanelTc* h, int qbuf, int round, int sweep, const int* cnt, const int* qflag, cudaStream_t st);

// One-pass symmetric update (jacobi_sym.cu): G <- Q^T G Q in place, upper block triangle + mirrored store, one launch.
bool panel_sym_supported(int np);
int panel_sym_prepare(PanelTc* h);
int panel_sym_update_g(PanelTc* h, int qbuf, int round, int sweep, const int* cnt, const int* qflag, cudaStream_t st);
int panel_sym_units_read(unsigned long long* out, int reset);   // 64 KB units of HBM traffic processed

int panel_tc_prepare_groups(PanelTc* h, const float* Pt0, const float* Pt1);
int panel_tc_update_g_groups(PanelTc* h, int pbuf, const PanelGroups& grp, int sweep, const int* cnt, const int* gflag,
                             cudaStream_t st);
int panel_tc_update_v_groups(PanelTc* h, int pbuf, const PanelGroups& grp, int sweep, const int* cnt, const int* gflag,
                             cudaStream_t st);

// Chained V update (jacobi_schedule = 2): V <- V Q1 Q2 Q3 for three XOR rounds inside 4-block cosets, one pass over V.
bool panel_chain_supported(int np);
void panel_chain_plan_host(const PanelGroups& grp, int g, int out[22]);   // host mirror of the kernel's tile bookkeeping (tests)
int panel_tc_prepare_chain(PanelTc* h, float* const Qc[6]);
int panel_tc_update_v_chain(PanelTc* h, int slot, const PanelGroups& grp, int sweep, const int* cnt,
                            const int* const qflag[3], cudaStream_t st);

// Panel tiles (128 rows x 64 output columns, 32 KB in + 32 KB out of HBM traffic) processed since the last reset:
// [0] G passes, [1] V passes, [2] tiles of the group-local (L2-resident) problems of the spread schedule.
int panel_tiles_read(unsigned long long out[3], int reset);

struct Options {
  int jacobi_update_tc = 1;     // 1: tcgen05 3xTF32 panel update, 0: SIMT fp32 tile update
  float jacobi_tol = 1e-6f;     // relative rotation threshold |s_pq| > tol sqrt(s_pp s_qq) (1e-5 costs 0.3 ms less per step and
                                // ~10x in gradient accuracy: 8e-5 vs 8e-6 on square samples)
  int jacobi_max_sweeps = 16;
  int jacobi_chunks = 2;        // split the batch into this many chunks (<= 4) whose Jacobi iterations run on separate streams, so
                                // that one chunk's issue-bound inner solve shares the SMs with another chunk's latency-bound
                                // panel passes (a chunk must still fill the GPU once with inner-solver CTAs).  Headline shape:
                                // 1 -> 34.8 ms, 2 -> 32.7 ms, 3 -> 33.4 ms, 4 -> 34.3 ms per step (round 1's kernels were
                                // slower with 2: 69.8 vs 65.9 ms)
  float jacobi_tol_pass1 = 1e-6f; // first-pass relative threshold when a second pass follows (looser values are slower:
                                  // 1e-4 -> 56.7 ms, 1e-3 -> 59.5 ms vs 55.3 ms at 1e-5, measured before the raised floor)
  float jacobi_nu_pass1 = 8192.f; // first pass of the two-pass solver: absolute significance floor in units of 2^-23 max|diag|
                                  // (single-pass solver: 4).  The second pass removes what the first leaves, so the first may stop
                                  // early.  With the scale-free second pass (below), gradient error / step time on the hard inputs
                                  // of scripts/dbg_floor_sweep.py: 2048 -> <= 1.6e-5 / 32.4 ms, 8192 -> <= 1.4e-5 / 31.5 ms,
                                  // 16384 -> <= 1.9e-5 but two inputs reach the second pass's sweep cap, 32768 -> 7e-5,
                                  // 65536 -> 3e-3 (the smallest directions are lost)
  float jacobi_nu_pass2 = -1e-10f; // second pass.  > 0: absolute floor in the same units.  < 0 (default): scale-free -- a rotation
                                  // is significant when it passes the relative test and at least one of its two directions has
                                  // a diagonal entry above |value| * max|diag| (1e-10 = (1e-4)^2 * 1e-2: below the numerical-rank
                                  // cut-off only noise is left).  G2 = Y Y^T is graded (entries accurate relative to their own
                                  // rows), so a floor relative to lambda_max is wrong for it: with a dominant mean component
                                  // (ReLU features, lambda_max / lambda_bulk ~ 10^3) the old floor of 4 ulps stopped the pass
                                  // while the bulk was resolved to 1e-3 only (gradient errors 1e-3 .. 8e-3 on such inputs,
                                  // scripts/dbg_floor_sweep.py); scale-free: <= 1.6e-5 on every input tried, +1.5 ms per step
  int erank_pass1_sweeps = 12;  // sweep cap of the first pass of the two-pass solver (it converges in 8-11 with the raised floor;
                                // whatever a capped matrix still needs, the second pass does); 0 = jacobi_max_sweeps
  int erank_pass2_sweeps = -1;  // sweep cap of the second pass: < 0 (default) = automatic, 6 up to n = 512 and 8 beyond (the second
                                // pass takes 3-5 sweeps at n <= 512, 7 at n = 2048 with a decaying spectrum; each spare sweep costs
                                // ~35 launches per chunk that return at once); 0 = jacobi_max_sweeps; > 0 = that many
  int erank_passes = 2;         // 2: second refinement pass (G2 = Y Y^T -> Jacobi -> U, Y updated): relative accuracy for the
                                //    smallest singular directions (gradients <= 1e-4 on square samples), +15 % time;
                                //    1: single pass (erank itself is already <= 3e-6; gradients 2e-4 .. 1e-2)
  int panel_merged = 0;         // 1: both G passes in one launch, H in an L2-resident ring (DRAM traffic 570 -> 402 MB per round,
                                //    but measured slower, 31.8 vs 29.2 ms/step: the per-CTA tile pipeline, not HBM, bounds that
                                //    schedule); 0: two launches through HBM (default)
  int panel_group_mb = 8;       // merged schedule: MB of G per group
  int panel_ring = 6;           // merged schedule: ring slots (groups) of H
  int jacobi_inner_regs = 1;    // 1: register-resident inner solver for the cross rounds, 0: shared-memory solver everywhere
  int gemm_tc = 1;              // 1: refinement / backward / fp32 Gram GEMMs on tcgen05 via bf16 planes, 0: SIMT
  int jacobi_v_after_g = 0;     // 0: V(r) starts right after inner(r) and runs beside the G passes (58.9 ms); 1: V(r) starts after the
                                //    G passes of round r and runs beside inner(r+1) (60.1 ms with the register-resident inner solver)
  int lin_fast = 1;             // linear GEMM: 16-warp epilogue variant for bf16 full-tile problems (0: the generic 8-warp one)
  int jacobi_overlap_v = 1;     // run V <- V Q on a side stream, overlapped with the next inner solve
  int panel_sym = 1;            // 1: G <- Q^T G Q as ONE in-place pass over the upper block triangle with mirrored stores
                                //    (jacobi_sym.cu: 0.56 n^2 read + n^2 written per round); 0: two passes through the scratch
                                //    matrix H (jacobi_tc.cu: 2 n^2 read + 2 n^2 written)
  int jacobi_own_streams = 1;   // 1: with several chunks, chunk 0 also runs on a library-owned stream (forked from / joined to the
                                // caller's), so that all chunk streams have the same (high) priority and the V side streams the
                                // low one: 32.73 -> 32.46 ms; 0: chunk 0 on the caller's stream (33.4 ms: the chunks get unequal
                                // priorities)
  int jacobi_schedule = 2;      // 2 (default): where the block count is a power of two >= 8, the rounds of a sweep are the XOR
                                //    matchings grouped three at a time into super-rounds {a, b, a^b} that stay inside 4-block
                                //    (128-column) cosets; G is updated every round, the three V updates of a super-round run as
                                //    ONE chained pass over V (panel_vchain_kernel): 41.96 -> 37.26 ms per step;
                                // 1: spread schedule on the same grouping: group-local problems + one K = 128 pass over G and V
                                //    per super-round (measured slower, 44.5 ms);
                                // 0: circle-method round robin, one panel update of G and V per round (also the fallback for
                                //    other block counts)
};
Options& options();
}  // namespace r3d
