// Shared definitions of the sparkdl-b200 native control plane.
//
// The reference has no native layer at all ("Native components: none"); this is
// the net-new runtime that backs the hvd.* calls a pickled main() makes (SURVEY.md §1c,
// "Scheduler: per-rank background engine: tensor queue -> negotiation -> fusion -> launch").
#pragma once
#include <atomic>
#include <cstdint>
#include <cstring>
#include <string>

namespace sdl {

constexpr int kMaxRanks = 64;        // CPU task slots on one box (GPU gangs are <= 8)
constexpr int kMaxName = 128;        // tensor-name bytes incl. NUL
constexpr int kMaxDims = 6;
constexpr int kMaxFused = 128;       // tensors fused into one response
constexpr int kReqRing = 256;        // outstanding requests per rank
constexpr int kRespRing = 64;        // outstanding responses
constexpr int kCacheWords = 16;      // response cache: 64 * kCacheWords entries, one announce bit each
constexpr int kCacheBits = 64 * kCacheWords;

enum DType : uint32_t {
  kU8 = 0, kI8 = 1, kI32 = 2, kI64 = 3, kF16 = 4, kBF16 = 5, kF32 = 6, kF64 = 7, kBool = 8,
  kF8E4M3 = 9,
};

inline size_t dtype_size(uint32_t dt) {
  switch (dt) {
    case kU8: case kI8: case kBool: case kF8E4M3: return 1;
    case kF16: case kBF16: return 2;
    case kI32: case kF32: return 4;
    default: return 8;
  }
}

enum ReduceOp : uint32_t { kSum = 0, kAverage = 1, kMin = 2, kMax = 3, kProduct = 4 };

enum ReqType : uint32_t {
  kAllreduce = 0, kAllgather = 1, kBroadcast = 2, kAlltoall = 3, kReducescatter = 4,
  kBarrier = 5, kJoin = 6, kShutdown = 7, kError = 8,
};

struct Request {
  uint32_t type;
  uint32_t dtype;
  int32_t device;        // -1 = host memory, >= 0 = CUDA device ordinal
  uint32_t op;
  int32_t root;
  uint32_t ndim;
  int64_t shape[kMaxDims];
  double prescale;
  double postscale;
  int32_t has_splits;
  int32_t splits[kMaxRanks];
  char name[kMaxName];
};

struct Response {
  uint64_t seq;
  uint32_t type;
  uint32_t dtype;
  int32_t device;
  uint32_t op;
  int32_t root;
  uint32_t n;                            // number of fused tensors
  double prescale;
  double postscale;
  int32_t last_joined;                   // kJoin: rank that joined last
  uint32_t has_error;
  uint32_t cacheable;                    // every rank may insert these names into its response cache
  uint32_t from_cache;                   // agreed through the announce bitmask, not through request records
  char error[512];
  int64_t numel[kMaxFused];              // per tensor element count (joined ranks need it)
  int64_t dim0[kMaxRanks];               // allgather: first-dim size per rank
  int64_t row_elems;                     // allgather/alltoall: elements per dim-0 row
  int32_t splits[kMaxRanks * kMaxRanks]; // alltoall: splits[src * world + dst]
  char names[kMaxFused][kMaxName];
};

inline uint64_t now_us();

}  // namespace sdl
