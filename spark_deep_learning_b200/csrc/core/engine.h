// Per-rank background engine: tensor queue -> name-based negotiation over shared memory ->
// tensor fusion -> execution (host collectives here, CUDA collectives through a registered
// executor).  SURVEY.md §7 step 3 and "hard part" 5: ranks may submit named tensors in
// different orders; rank 0's engine thread is the coordinator that decides a single global
// order, so fusion and execution order are identical on every rank.
#pragma once
#include <atomic>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "common.h"
#include "cpu_collectives.h"
#include "gpu_work.h"
#include "shm.h"

namespace sdl {

struct EngineOptions {
  std::string shm_name;
  int rank = 0, world = 1;
  uint64_t stage_bytes = 8 << 20;
  int64_t fusion_threshold = 64 << 20;
  int cycle_us = 500;
  double stall_warn_s = 60.0, stall_shutdown_s = 0.0;
  double init_timeout_s = 120.0;
  std::string timeline_path;
};

struct HandleState {
  std::atomic<int> status{0};  // 0 pending, 1 done, -1 error
  std::string error;
  std::vector<uint8_t> out_buf;   // engine-allocated output (allgather / alltoall / reducescatter)
  int64_t out_dim0 = 0;
  std::vector<int32_t> recv_splits;
  uint64_t done_event = 0;
  int last_joined = -1;
};

struct Entry {
  int64_t handle;
  Request req;
  void* in;
  void* out;
  int64_t numel;
  uint64_t ready_event;
  uint64_t enqueue_us;
};

class Timeline {
 public:
  bool open(const std::string& path, int rank);   // false: empty path, already open, or fopen failed
  void event(const std::string& name, const char* phase, char ph, uint64_t ts_us);
  void close();
  bool enabled() const { return f_.load(std::memory_order_acquire) != nullptr; }

 private:
  std::atomic<FILE*> f_{nullptr};                 // opened / closed at run time (hvd.start_timeline / stop_timeline)
  int rank_ = 0;
  bool first_ = true;
  std::mutex mu_;
};

class Engine {
 public:
  Engine() = default;
  ~Engine();
  std::string init(const EngineOptions& opt);
  void shutdown();
  bool timeline_start(const std::string& path) { return timeline_.open(path, opt_.rank); }
  void timeline_stop() { timeline_.close(); }
  int64_t enqueue(const Request& req, void* in, void* out, int64_t numel, uint64_t ready_event);
  int poll(int64_t handle);
  int wait(int64_t handle, double timeout_s);
  std::shared_ptr<HandleState> state(int64_t handle);
  void release(int64_t handle);
  void set_gpu_executor(GpuExecFn fn, void* ctx) { gpu_fn_ = fn; gpu_ctx_ = ctx; }
  ShmSegment* shm() { return &shm_; }
  int rank() const { return opt_.rank; }
  int world() const { return opt_.world; }
  uint64_t responses_executed() const { return responses_executed_; }
  uint64_t tensors_fused() const { return tensors_fused_; }
  uint64_t cache_hits() const { return cache_hits_; }
  uint64_t cache_misses() const { return cache_misses_; }
  uint64_t cache_size() const { return cache_.size(); }
  double negotiation_us() const { return neg_count_ ? (double)neg_us_sum_ / (double)neg_count_ : 0.0; }
  void set_cache_enabled(bool on) { cache_enabled_ = on; }

 private:
  struct Record {
    Request first;
    uint64_t seen = 0;
    int64_t dim0[kMaxRanks];
    std::vector<int32_t> splits;  // world*world when alltoall
    std::string error;
    uint64_t order = 0;
    uint64_t first_us = 0;
    uint64_t last_warn_us = 0;
  };
  void loop();
  void publish_local();
  void coordinate();
  void absorb(int src, const Request& r);
  void emit_ready();
  void emit_cached();
  void cache_insert(const Response& r, const std::vector<Entry*>& es);
  int cache_lookup(const Request& r);
  void publish(Response* r);
  void check_stalls();
  bool consume_responses();
  void execute(const Response& r);
  void exec_allreduce(const Response& r, std::vector<Entry*>& es);
  void complete(Entry* e, const std::string& err);
  void complete_handle(int64_t h, const std::string& err);
  void fail_all(const std::string& err);

  EngineOptions opt_;
  ShmSegment shm_;
  std::unique_ptr<CpuCollectives> cpu_;
  std::thread thread_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<Entry> queue_;                               // submitted, not yet announced
  std::unordered_map<std::string, Entry> pending_;        // announced, waiting for a response
  std::unordered_map<int64_t, std::shared_ptr<HandleState>> handles_;
  int64_t next_handle_ = 1;
  bool running_ = false, stop_ = false, joined_local_ = false, shutdown_seen_ = false;
  std::string fatal_;
  GpuExecFn gpu_fn_ = nullptr;
  void* gpu_ctx_ = nullptr;
  // coordinator (rank 0)
  std::map<std::string, Record> table_;
  uint64_t order_ = 0, joined_mask_ = 0, shutdown_mask_ = 0;
  int last_joined_ = -1;
  std::vector<uint8_t> fuse_buf_;
  Timeline timeline_;
  uint64_t responses_executed_ = 0, tensors_fused_ = 0;
  // response cache (identical on every rank: entries are inserted while executing responses, which all ranks see in
  // the same order)
  std::vector<Request> cache_;
  std::unordered_map<std::string, int> cache_index_;
  bool cache_enabled_ = true;
  std::atomic<uint64_t> cache_hits_{0}, cache_misses_{0}, neg_us_sum_{0}, neg_count_{0};
  // coordinator side of the fast path
  uint64_t acc_bits_[kMaxRanks][kCacheWords] = {};
  std::vector<uint8_t> demoted_;          // cache ids that must go through the record path until their next response
  std::vector<uint64_t> cache_first_us_;  // stall reporting
};

}  // namespace sdl
