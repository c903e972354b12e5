#include "shm.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>

#include <cerrno>
#include <cstdio>
#include <cstring>
#include <new>

namespace sdl {

static double mono_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + ts.tv_nsec * 1e-9;
}

ShmSegment::~ShmSegment() { close(false); }

std::string ShmSegment::open(const std::string& name, int rank, int world, uint64_t stage_bytes,
                             double timeout_s) {
  if (world < 1 || world > kMaxRanks) return "world size out of range for the shm backend";
  name_ = name;
  rank_ = rank;
  world_ = world;
  stage_bytes_ = (stage_bytes + 4095) & ~uint64_t(4095);
  uint64_t hdr_bytes = (sizeof(ShmHeader) + 4095) & ~uint64_t(4095);
  map_bytes_ = hdr_bytes + stage_bytes_ * (uint64_t)(world + 1);

  int fd = -1;
  if (rank == 0) {
    shm_unlink(name.c_str());
    fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) return std::string("shm_open(create) failed: ") + strerror(errno);
    if (ftruncate(fd, (off_t)map_bytes_) != 0) {
      ::close(fd);
      return std::string("ftruncate failed: ") + strerror(errno);
    }
  } else {
    double deadline = mono_s() + timeout_s;
    while (true) {
      fd = shm_open(name.c_str(), O_RDWR, 0600);
      if (fd >= 0) {
        struct stat st;
        if (fstat(fd, &st) == 0 && (uint64_t)st.st_size >= map_bytes_) break;
        ::close(fd);
        fd = -1;
      }
      if (mono_s() > deadline) return "timed out waiting for rank 0 to create the shm segment";
      usleep(1000);
    }
  }
  base_ = mmap(nullptr, map_bytes_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  ::close(fd);
  if (base_ == MAP_FAILED) {
    base_ = nullptr;
    return std::string("mmap failed: ") + strerror(errno);
  }
  hdr_ = reinterpret_cast<ShmHeader*>(base_);
  data_ = reinterpret_cast<uint8_t*>(base_) + hdr_bytes;
  if (rank == 0) {
    // ftruncate zero-fills, so every atomic starts at 0; only the scalars need values.
    hdr_->world = (uint32_t)world;
    hdr_->version = 1;
    hdr_->stage_bytes = stage_bytes_;
    hdr_->data_offset = hdr_bytes;
    hdr_->magic.store(kShmMagic, std::memory_order_release);
  } else {
    double deadline = mono_s() + timeout_s;
    while (hdr_->magic.load(std::memory_order_acquire) != kShmMagic) {
      if (mono_s() > deadline) return "timed out waiting for the shm header";
      usleep(200);
    }
    if ((int)hdr_->world != world) return "shm segment belongs to a job of a different size";
  }
  hdr_->attached.fetch_add(1, std::memory_order_acq_rel);
  return "";
}

void ShmSegment::close(bool unlink) {
  if (base_) {
    if (hdr_) hdr_->detached.fetch_add(1, std::memory_order_acq_rel);
    munmap(base_, map_bytes_);
    base_ = nullptr;
    hdr_ = nullptr;
    data_ = nullptr;
  }
  if (unlink && !name_.empty()) shm_unlink(name_.c_str());
}

void ShmSegment::ring_doorbell() {
  hdr_->doorbell.fetch_add(1, std::memory_order_release);
  futex_wake_all(&hdr_->doorbell);
}

void ShmSegment::abort_job() {
  if (!hdr_) return;
  hdr_->abort_flag.store(1, std::memory_order_release);
  for (auto& b : hdr_->barriers) {
    b.gen.fetch_add(1 << 16, std::memory_order_release);  // release anyone parked in a barrier
    futex_wake_all(&b.gen);
  }
  ring_doorbell();
}

bool ShmSegment::barrier(int idx, double timeout_s) {
  if (world_ == 1) return true;
  ShmBarrier& b = hdr_->barriers[idx];
  uint32_t g = b.gen.load(std::memory_order_acquire);
  if (b.count.fetch_add(1, std::memory_order_acq_rel) + 1 == (uint32_t)world_) {
    b.count.store(0, std::memory_order_relaxed);
    b.gen.fetch_add(1, std::memory_order_release);
    futex_wake_all(&b.gen);
    return !aborted();
  }
  // Sleeping costs a futex wake-up (milliseconds inside a VM): spin for ~200us first.
  const double spin_until = mono_s() + 200e-6;
  for (int i = 0;; ++i) {
    if (b.gen.load(std::memory_order_acquire) != g) return !aborted();
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
    if ((i & 255) == 255 && mono_s() > spin_until) break;
  }
  double deadline = mono_s() + timeout_s;
  while (b.gen.load(std::memory_order_acquire) == g) {
    if (aborted()) return false;
    futex_wait(&b.gen, g, 50000);
    if (mono_s() > deadline) return false;
  }
  return !aborted();
}

}  // namespace sdl
