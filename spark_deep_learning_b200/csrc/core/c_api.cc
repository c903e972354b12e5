// C ABI of libsparkdl_core.so, consumed from Python through ctypes
// (spark_deep_learning_b200/parallel/core.py).
#include <unistd.h>

#include <chrono>
#include <cstring>
#include <string>
#include <thread>

#include "engine.h"

using namespace sdl;

static thread_local std::string g_err;

extern "C" {

void* sdl_engine_create() { return new Engine(); }

const char* sdl_engine_init(void* e, const char* shm_name, int rank, int world, uint64_t stage_bytes,
                            int64_t fusion_threshold, int cycle_us, double stall_warn_s,
                            double stall_shutdown_s, double init_timeout_s, const char* timeline) {
  EngineOptions o;
  o.shm_name = shm_name;
  o.rank = rank;
  o.world = world;
  o.stage_bytes = stage_bytes;
  o.fusion_threshold = fusion_threshold;
  o.cycle_us = cycle_us;
  o.stall_warn_s = stall_warn_s;
  o.stall_shutdown_s = stall_shutdown_s;
  o.init_timeout_s = init_timeout_s;
  o.timeline_path = timeline ? timeline : "";
  g_err = static_cast<Engine*>(e)->init(o);
  return g_err.c_str();
}

void sdl_engine_shutdown(void* e) { static_cast<Engine*>(e)->shutdown(); }
void sdl_engine_destroy(void* e) { delete static_cast<Engine*>(e); }
int sdl_timeline_start(void* e, const char* path) { return static_cast<Engine*>(e)->timeline_start(path ? path : "") ? 0 : -1; }
void sdl_timeline_stop(void* e) { static_cast<Engine*>(e)->timeline_stop(); }
void sdl_engine_abort(void* e) { static_cast<Engine*>(e)->shm()->abort_job(); }

void sdl_set_gpu_executor(void* e, void* fn, void* ctx) {
  static_cast<Engine*>(e)->set_gpu_executor(reinterpret_cast<GpuExecFn>(fn), ctx);
}

int64_t sdl_enqueue(void* e, uint32_t type, const char* name, uint32_t dtype, int device, uint32_t op,
                    int root, int ndim, const int64_t* shape, double prescale, double postscale,
                    void* in, void* out, const int32_t* splits, uint64_t ready_event) {
  Request r;
  memset(&r, 0, sizeof(r));
  r.type = type;
  r.dtype = dtype;
  r.device = device;
  r.op = op;
  r.root = root;
  r.ndim = (uint32_t)(ndim > kMaxDims ? kMaxDims : ndim);
  int64_t numel = 1;
  for (int i = 0; i < ndim; ++i) numel *= shape[i];
  // tensors with more than kMaxDims dims: fold the trailing dims (only dim 0 has semantics)
  for (int i = 0; i < (int)r.ndim; ++i) r.shape[i] = shape[i];
  if (ndim > kMaxDims) {
    int64_t tail = 1;
    for (int i = kMaxDims - 1; i < ndim; ++i) tail *= shape[i];
    r.shape[kMaxDims - 1] = tail;
  }
  r.prescale = prescale;
  r.postscale = postscale;
  if (splits) {
    r.has_splits = 1;
    int w = static_cast<Engine*>(e)->world();
    for (int i = 0; i < w && i < kMaxRanks; ++i) r.splits[i] = splits[i];
  }
  snprintf(r.name, kMaxName, "%s", name);
  return static_cast<Engine*>(e)->enqueue(r, in, out, numel, ready_event);
}

int sdl_poll(void* e, int64_t h) { return static_cast<Engine*>(e)->poll(h); }
int sdl_wait(void* e, int64_t h, double timeout_s) { return static_cast<Engine*>(e)->wait(h, timeout_s); }

int sdl_handle_error(void* e, int64_t h, char* buf, int len) {
  auto st = static_cast<Engine*>(e)->state(h);
  if (!st) return -1;
  snprintf(buf, len, "%s", st->error.c_str());
  return (int)st->error.size();
}
int64_t sdl_handle_out_bytes(void* e, int64_t h) {
  auto st = static_cast<Engine*>(e)->state(h);
  return st ? (int64_t)st->out_buf.size() : -1;
}
int64_t sdl_handle_out_dim0(void* e, int64_t h) {
  auto st = static_cast<Engine*>(e)->state(h);
  return st ? st->out_dim0 : -1;
}
int sdl_handle_copy_out(void* e, int64_t h, void* dst) {
  auto st = static_cast<Engine*>(e)->state(h);
  if (!st) return -1;
  if (!st->out_buf.empty()) memcpy(dst, st->out_buf.data(), st->out_buf.size());
  return 0;
}
int sdl_handle_recv_splits(void* e, int64_t h, int32_t* dst) {
  auto st = static_cast<Engine*>(e)->state(h);
  if (!st) return -1;
  for (size_t i = 0; i < st->recv_splits.size(); ++i) dst[i] = st->recv_splits[i];
  return (int)st->recv_splits.size();
}
int sdl_handle_last_joined(void* e, int64_t h) {
  auto st = static_cast<Engine*>(e)->state(h);
  return st ? st->last_joined : -1;
}
uint64_t sdl_handle_done_event(void* e, int64_t h) {
  auto st = static_cast<Engine*>(e)->state(h);
  return st ? st->done_event : 0;
}
void sdl_release(void* e, int64_t h) { static_cast<Engine*>(e)->release(h); }

uint64_t sdl_stat_responses(void* e) { return static_cast<Engine*>(e)->responses_executed(); }
uint64_t sdl_stat_tensors(void* e) { return static_cast<Engine*>(e)->tensors_fused(); }
uint64_t sdl_stat_cache_hits(void* e) { return static_cast<Engine*>(e)->cache_hits(); }
uint64_t sdl_stat_cache_misses(void* e) { return static_cast<Engine*>(e)->cache_misses(); }
uint64_t sdl_stat_cache_size(void* e) { return static_cast<Engine*>(e)->cache_size(); }
double sdl_stat_negotiation_us(void* e) { return static_cast<Engine*>(e)->negotiation_us(); }
void sdl_set_cache_enabled(void* e, int on) { static_cast<Engine*>(e)->set_cache_enabled(on != 0); }

// stand-alone helpers (unit tests of the host reduction kernels)
void sdl_reduce_into(void* dst, const void* src, int64_t n, uint32_t dtype, uint32_t op) {
  reduce_into(dst, src, n, dtype, op);
}
void sdl_scale_buffer(void* buf, int64_t n, uint32_t dtype, double f) { scale_buffer(buf, n, dtype, f); }

// Last-resort exit watchdog for worker processes: the worker leaves through the normal interpreter shutdown
// (so at-exit hooks run), and this detached native thread ends the process with `code` if the shutdown has not
// completed after `seconds` (a wedged CUDA context teardown must not hold the whole gang).
void sdl_exit_watchdog(double seconds, int code) {
  std::thread([seconds, code]() {
    std::this_thread::sleep_for(std::chrono::duration<double>(seconds));
    _exit(code);
  }).detach();
}

}  // extern "C"
