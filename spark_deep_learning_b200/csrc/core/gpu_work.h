// Work item the engine hands to the CUDA executor (implemented in libsparkdl_cuda.so).
// Kept in its own header so the CUDA library does not pull in the rest of the engine.
#pragma once
#include <cstdint>

namespace sdl {

struct GpuWork {
  uint32_t type, dtype, op;
  int32_t root, device;
  uint32_t n;
  double prescale, postscale;
  void** in;
  void** out;
  int64_t* numel;
  uint64_t* ready_events;     // cudaEvent_t recorded by the producer stream (0 = none)
  const int64_t* dim0;        // allgather: rows per rank
  int64_t row_elems;
  const int32_t* splits;      // alltoall matrix [src * kMaxRanks... see note] in rows
  int32_t splits_stride;      // row stride of `splits` (world)
};
typedef int (*GpuExecFn)(void* ctx, const GpuWork* work, uint64_t* done_event, char* err, int errlen);

}  // namespace sdl
