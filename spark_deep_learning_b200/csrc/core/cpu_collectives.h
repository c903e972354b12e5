// Host-memory collectives over the shared staging area of ShmSegment.  Serves CPU tensors
// (BASELINE.json config 1: "HorovodRunner(np=1).run(main) on CPU ... allreduce on a 1k tensor")
// and the T1/T3 plumbing tests of SURVEY.md §4b.
#pragma once
#include <cstdint>

#include "shm.h"

namespace sdl {

// element-wise: dst[i] = dst[i] (op) src[i]
void reduce_into(void* dst, const void* src, int64_t n, uint32_t dtype, uint32_t op);
// element-wise scale of floating types (no-op for integers unless factor is integral)
void scale_buffer(void* buf, int64_t n, uint32_t dtype, double factor);

class CpuCollectives {
 public:
  explicit CpuCollectives(ShmSegment* shm) : shm_(shm) {}
  bool allreduce(void* buf, int64_t numel, uint32_t dtype, uint32_t op);
  bool allgather(const void* in, void* out, const int64_t* bytes_per_rank);
  bool broadcast(void* buf, int64_t bytes, int root);
  bool alltoall(const void* in, void* out, const int64_t* send_bytes, const int64_t* recv_bytes);
  bool reducescatter(const void* in, void* out, const int64_t* numel_per_rank, uint32_t dtype,
                     uint32_t op);
  bool barrier() { return shm_->barrier(0); }

 private:
  ShmSegment* shm_;
};

}  // namespace sdl
