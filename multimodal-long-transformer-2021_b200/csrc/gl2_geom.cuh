// Shared geometry of the gl2 kernels (persistent kernels specialised for compact global-local attention
// with local_radius == 64): tile / chunk sizes and the slot order of the per-row relative table.
#pragma once

namespace mlt {
namespace gl2 {

constexpr int TM = 128;          // rows per tile
constexpr int TK = 128;          // keys (queries) per chunk: two blocks of 64
constexpr int NST = 3;           // ring stages
constexpr int RAD = 64;          // local_radius the chunk schedule is built for
constexpr float LOG2E = 1.4426950408889634f;

// ids 0..2D in offset order -D..D (slot = offset + D), other ids unchanged
__host__ __device__ __forceinline__ int slot_of_id(int id, int D) {
  if (id <= D) return D + id;
  if (id <= 2 * D) return 2 * D - id;
  return id;
}

}  // namespace gl2
}  // namespace mlt
