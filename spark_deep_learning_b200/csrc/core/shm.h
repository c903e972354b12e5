// POSIX shared-memory segment shared by all ranks of one job on one box, plus the futex
// primitives built on it.  This is the "CPU fallback: POSIX-shm + futex backend" of
// SURVEY.md §5.8 and at the same time the negotiation fabric of the engine.
#pragma once
#include <linux/futex.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <atomic>
#include <cstdint>
#include <string>

#include "common.h"

namespace sdl {

struct ShmBarrier {
  std::atomic<uint32_t> count;
  std::atomic<uint32_t> gen;
  char pad[56];
};

struct RankRing {
  std::atomic<uint64_t> head;   // written by the owning rank
  std::atomic<uint64_t> tail;   // written by the coordinator (rank 0)
  char pad[48];
  Request slots[kReqRing];
};

struct ShmHeader {
  std::atomic<uint64_t> magic;
  uint32_t world;
  uint32_t version;
  uint64_t stage_bytes;
  uint64_t data_offset;
  std::atomic<uint32_t> doorbell;      // futex word: bumped on every ring / response write
  std::atomic<uint32_t> abort_flag;    // any rank sets it to tear the job down
  std::atomic<uint32_t> attached;      // ranks that mapped the segment
  std::atomic<uint32_t> detached;
  ShmBarrier barriers[4];              // [0] data plane, [1] init/shutdown, [2..3] spare
  std::atomic<uint64_t> resp_head;     // responses published so far
  std::atomic<uint64_t> resp_tail[kMaxRanks];
  // response cache fast path: rank r announces "my cached tensor #i is ready" by setting bit i of its row; the
  // coordinator swaps the words out and ANDs the rows - no request record, no string handling
  std::atomic<uint64_t> cache_bits[kMaxRanks][kCacheWords];
  RankRing rings[kMaxRanks];
  Response resp[kRespRing];
};

constexpr uint64_t kShmMagic = 0x53444c4232303042ull;  // "SDLB200B"

inline long futex_wait(std::atomic<uint32_t>* addr, uint32_t expected, int timeout_us) {
  struct timespec ts;
  ts.tv_sec = timeout_us / 1000000;
  ts.tv_nsec = (timeout_us % 1000000) * 1000L;
  return syscall(SYS_futex, reinterpret_cast<uint32_t*>(addr), FUTEX_WAIT, expected, &ts,
                 nullptr, 0);
}

inline long futex_wake_all(std::atomic<uint32_t>* addr) {
  return syscall(SYS_futex, reinterpret_cast<uint32_t*>(addr), FUTEX_WAKE, 0x7fffffff, nullptr,
                 nullptr, 0);
}

class ShmSegment {
 public:
  ShmSegment() = default;
  ~ShmSegment();
  // rank 0 creates (O_EXCL) and initialises; everyone else opens and waits for the magic.
  // Returns empty string on success, else an error message.
  std::string open(const std::string& name, int rank, int world, uint64_t stage_bytes,
                   double timeout_s);
  void close(bool unlink);

  ShmHeader* hdr() const { return hdr_; }
  uint8_t* stage(int rank) const { return data_ + (uint64_t)(rank + 1) * stage_bytes_; }
  uint8_t* result() const { return data_; }
  uint64_t stage_bytes() const { return stage_bytes_; }
  int world() const { return world_; }
  int rank() const { return rank_; }

  // Sense-reversing barrier with spin-then-futex waiting.  Returns false if the job was
  // aborted (or the wait exceeded timeout_s) so callers can unwind instead of hanging.
  bool barrier(int idx = 0, double timeout_s = 600.0);
  void ring_doorbell();
  bool aborted() const { return hdr_ && hdr_->abort_flag.load(std::memory_order_acquire) != 0; }
  void abort_job();

 private:
  std::string name_;
  ShmHeader* hdr_ = nullptr;
  uint8_t* data_ = nullptr;
  void* base_ = nullptr;
  uint64_t map_bytes_ = 0;
  uint64_t stage_bytes_ = 0;
  int rank_ = 0, world_ = 1;
};

}  // namespace sdl
