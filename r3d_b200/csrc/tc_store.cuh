// Coalesced store of a warp's 32 x 32 fp32 accumulator block (lane = row, v[0..31] = 32 consecutive columns, as
// tcgen05.ld 32x32b.x32 delivers it).  A direct store would touch 32 different 128-byte lines per instruction;
// here 16 columns at a time are staged through shared memory and written as 64-byte row segments, 4 lanes per
// segment (full 32-byte sectors, 8 rows per store instruction).  `stg`: 2560 bytes of shared memory per warp.
#pragma once
#include <stdint.h>

namespace r3d {

constexpr int kStgPitch = 80;                 // bytes per staged row: 16 floats + 16 B pad -> conflict-free 128-bit access
constexpr int kStgWarpBytes = 32 * kStgPitch;

// dst points at element (first row of the warp's block, first of the 32 columns); pitch in floats.
// accumulate != 0 adds the previous contents (read with the same coalesced pattern).
__device__ __forceinline__ void staged_store_32x32(uint8_t* stg, int lane, const uint32_t (&v)[32], float* dst,
                                                   int64_t pitch, int accumulate) {
#pragma unroll
  for (int qt = 0; qt < 2; ++qt) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(stg + lane * kStgPitch + j * 16) =
          make_float4(__uint_as_float(v[16 * qt + 4 * j]), __uint_as_float(v[16 * qt + 4 * j + 1]),
                      __uint_as_float(v[16 * qt + 4 * j + 2]), __uint_as_float(v[16 * qt + 4 * j + 3]));
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int id = j * 32 + lane, r = id >> 2, c = id & 3;
      float4 val = *reinterpret_cast<const float4*>(stg + r * kStgPitch + c * 16);
      float4* p = reinterpret_cast<float4*>(dst + r * pitch + qt * 16 + c * 4);
      if (accumulate) {
        const float4 old = *p;
        val.x += old.x; val.y += old.y; val.z += old.z; val.w += old.w;
      }
      *p = val;
    }
  }
}

}  // namespace r3d
