// Host-side interface of the bf16-plane tcgen05 GEMM (pgemm_tcgen05.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace r3d {

struct PGemm {
  // operands are bf16 plane stacks [planes][batch][rows][cols]:
  //   K-major  operand: rows = M (or N), cols = K      MN-major operand: rows = K, cols = M (or N)
  const void* A; const void* B;
  int M, N, K, batch;
  int pa, pb;                   // planes per operand (1..3)
  int a_kmajor, b_kmajor;
  int nprod;                    // plane products accumulated, in this order
  int prod_a[6], prod_b[6];
  int out_mode;                 // 0 fp32 store, 1 fp32 accumulate, 2 bf16 store, 3 bf16 accumulate
  void* C; int64_t ldc, strideC;
};

int pgemm_launch(const PGemm& a, cudaStream_t st);
bool pgemm_operand_ok(const void* base, int64_t cols);
// x (R3D_F32 / R3D_BF16, `count` elements in rows of `cols`) -> P bf16 planes [P][count]; optional per-row scale
int split_planes(const void* x, int in_dtype, __nv_bfloat16* out, int64_t count, int P, int64_t cols,
                 const float* rowscale, cudaStream_t st);

// products giving ~2^-24 (all six) or ~2^-16 (first three) relative accuracy
inline void pgemm_products(PGemm& g, int pa, int pb, bool full) {
  static const int ia[6] = {0, 0, 1, 0, 2, 1}, ib[6] = {0, 1, 0, 2, 0, 1};
  int n = 0;
  for (int i = 0; i < 6; ++i) {
    if (ia[i] >= pa || ib[i] >= pb) continue;
    if (!full && ia[i] + ib[i] > 1) continue;
    g.prod_a[n] = ia[i]; g.prod_b[n] = ib[i]; ++n;
  }
  g.nprod = n;
}

}  // namespace r3d
