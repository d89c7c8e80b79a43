// Host-side handle of the tensor-core panel update (jacobi_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace r3d {
struct PanelTc {
  CUtensorMap map_g, map_h, map_v, map_q;
  float *G, *H, *V;
  int64_t B;
  int np, nb, nt;
};
bool panel_tc_supported(int np);
int panel_tc_prepare(PanelTc* h, float* G, float* H, float* V, const float* Qb, int64_t B, int np);
int panel_tc_round(PanelTc* h, int round, int sweep, const int* cnt, const int* qflag, cudaStream_t st);

struct Options {
  int jacobi_update_tc = 1;     // 1: tcgen05 3xTF32 panel update, 0: SIMT fp32 tile update
  float jacobi_tol = 1e-5f;     // relative off-diagonal threshold
  int jacobi_max_sweeps = 16;
};
Options& options();
}  // namespace r3d
