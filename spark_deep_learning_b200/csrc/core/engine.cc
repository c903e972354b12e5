#include "engine.h"

#include <time.h>

#include <algorithm>
#include <chrono>
#include <cstddef>
#include <cstdio>
#include <cstring>

namespace sdl {

static uint64_t mono_us() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (uint64_t)ts.tv_sec * 1000000ull + (uint64_t)ts.tv_nsec / 1000ull;
}

static const char* type_name(uint32_t t) {
  static const char* n[] = {"ALLREDUCE", "ALLGATHER", "BROADCAST", "ALLTOALL", "REDUCESCATTER",
                            "BARRIER", "JOIN", "SHUTDOWN", "ERROR"};
  return t <= kError ? n[t] : "?";
}

// ------------------------------------------------------------------------------------------
// Timeline: Chrome-trace JSON, same phase vocabulary as Horovod's HOROVOD_TIMELINE
// (SURVEY.md §5.1).
// ------------------------------------------------------------------------------------------
bool Timeline::open(const std::string& path, int rank) {
  if (path.empty()) return false;
  std::lock_guard<std::mutex> g(mu_);
  if (f_.load(std::memory_order_relaxed)) return false;
  std::string p = rank == 0 ? path : path + "." + std::to_string(rank);
  FILE* f = fopen(p.c_str(), "w");
  if (!f) return false;
  fputs("[\n", f);
  rank_ = rank;
  first_ = true;
  f_.store(f, std::memory_order_release);
  return true;
}

static std::string json_escape(const std::string& s) {
  std::string o;
  o.reserve(s.size());
  for (char c : s) {
    if (c == '"' || c == '\\') { o += '\\'; o += c; }
    else if ((unsigned char)c < 0x20) o += ' ';
    else o += c;
  }
  return o;
}

void Timeline::event(const std::string& raw_name, const char* phase, char ph, uint64_t ts_us) {
  if (!enabled()) return;
  const std::string name = json_escape(raw_name);
  std::lock_guard<std::mutex> g(mu_);
  FILE* f = f_.load(std::memory_order_relaxed);
  if (!f) return;                                  // stopped between the check and the lock
  size_t tid = std::hash<std::string>()(name) % 100000;
  fprintf(f, "%s{\"name\": \"%s\", \"cat\": \"%s\", \"ph\": \"%c\", \"ts\": %llu, \"pid\": %d, \"tid\": %zu}",
          first_ ? "" : ",\n", phase, name.c_str(), ph, (unsigned long long)ts_us, rank_, tid);
  first_ = false;
}

void Timeline::close() {
  std::lock_guard<std::mutex> g(mu_);
  FILE* f = f_.exchange(nullptr, std::memory_order_acq_rel);
  if (!f) return;
  fputs("\n]\n", f);
  fclose(f);
}

// ------------------------------------------------------------------------------------------
Engine::~Engine() {
  if (running_) shutdown();
}

std::string Engine::init(const EngineOptions& opt) {
  opt_ = opt;
  std::string err = shm_.open(opt.shm_name, opt.rank, opt.world, opt.stage_bytes, opt.init_timeout_s);
  if (!err.empty()) return err;
  if (!shm_.barrier(1, opt.init_timeout_s)) return "init barrier failed (a peer did not attach)";
  cpu_.reset(new CpuCollectives(&shm_));
  timeline_.open(opt.timeline_path, opt.rank);
  running_ = true;
  stop_ = false;
  thread_ = std::thread([this] { loop(); });
  return "";
}

void Engine::shutdown() {
  if (!running_) return;
  Request r;
  memset(&r, 0, sizeof(r));
  r.type = kShutdown;
  r.device = -1;
  snprintf(r.name, kMaxName, "__shutdown__");
  int64_t h = enqueue(r, nullptr, nullptr, 0, 0);
  if (wait(h, 15.0) == 0) {  // peers are gone: stop unilaterally
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
    }
    shm_.abort_job();
  }
  if (thread_.joinable()) thread_.join();
  running_ = false;
  timeline_.close();
  cpu_.reset();
  shm_.close(opt_.rank == 0);
}

int64_t Engine::enqueue(const Request& req, void* in, void* out, int64_t numel, uint64_t ready_event) {
  int64_t h;
  {
    std::lock_guard<std::mutex> g(mu_);
    h = next_handle_++;
    auto st = std::make_shared<HandleState>();
    handles_[h] = st;
    if (!fatal_.empty() || !running_) {
      st->status = -1;
      st->error = fatal_.empty() ? "engine is not running" : fatal_;
      cv_.notify_all();
      return h;
    }
    Entry e;
    e.handle = h;
    e.req = req;
    e.in = in;
    e.out = out;
    e.numel = numel;
    e.ready_event = ready_event;
    e.enqueue_us = mono_us();
    queue_.push_back(e);
  }
  if (timeline_.enabled())
    timeline_.event(req.name, (std::string("NEGOTIATE_") + type_name(req.type)).c_str(), 'B', mono_us());
  shm_.ring_doorbell();
  return h;
}

int Engine::poll(int64_t handle) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = handles_.find(handle);
  if (it == handles_.end()) return -2;
  return it->second->status;
}

int Engine::wait(int64_t handle, double timeout_s) {
  std::unique_lock<std::mutex> g(mu_);
  auto it = handles_.find(handle);
  if (it == handles_.end()) return -2;
  auto st = it->second;
  if (st->status == 0) {  // short optimistic spin before sleeping on the condition variable
    g.unlock();
    const uint64_t spin_until = mono_us() + 200;
    for (int i = 0; st->status.load(std::memory_order_acquire) == 0; ++i) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
      if ((i & 127) == 127 && mono_us() > spin_until) break;
    }
    g.lock();
  }
  auto deadline = std::chrono::steady_clock::now() + std::chrono::duration<double>(timeout_s);
  while (st->status == 0) {
    if (cv_.wait_until(g, deadline) == std::cv_status::timeout && st->status == 0) return 0;
  }
  return st->status;
}

std::shared_ptr<HandleState> Engine::state(int64_t handle) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = handles_.find(handle);
  return it == handles_.end() ? nullptr : it->second;
}

void Engine::release(int64_t handle) {
  std::lock_guard<std::mutex> g(mu_);
  handles_.erase(handle);
}

void Engine::complete_handle(int64_t h, const std::string& err) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = handles_.find(h);
  if (it != handles_.end()) {
    if (err.empty()) {
      it->second->status = 1;
    } else {
      it->second->status = -1;
      it->second->error = err;
    }
  }
  cv_.notify_all();
}

void Engine::complete(Entry* e, const std::string& err) { complete_handle(e->handle, err); }

void Engine::fail_all(const std::string& err) {
  std::lock_guard<std::mutex> g(mu_);
  fatal_ = err;
  for (auto& kv : pending_) {
    auto it = handles_.find(kv.second.handle);
    if (it != handles_.end() && it->second->status == 0) {
      it->second->status = -1;
      it->second->error = err;
    }
  }
  pending_.clear();
  for (auto& e : queue_) {
    auto it = handles_.find(e.handle);
    if (it != handles_.end() && it->second->status == 0) {
      it->second->status = -1;
      it->second->error = err;
    }
  }
  queue_.clear();
  cv_.notify_all();
}

// ------------------------------------------------------------------------------------------
// Engine thread
// ------------------------------------------------------------------------------------------
static constexpr uint64_t kLingerUs = 250;

void Engine::loop() {
  ShmHeader* hdr = shm_.hdr();
  uint64_t last_stall_check = mono_us();
  uint64_t last_active = 0;       // last time something was in flight: a synchronous caller's next op follows within us
  while (true) {
    uint32_t db = hdr->doorbell.load(std::memory_order_acquire);
    publish_local();
    if (opt_.rank == 0) {
      coordinate();
      uint64_t now = mono_us();
      if (now - last_stall_check > 1000000) {
        check_stalls();
        last_stall_check = now;
      }
    }
    consume_responses();
    if (shutdown_seen_) break;
    if (shm_.aborted()) {
      fail_all("sparkdl job aborted (a peer failed or the stall limit was reached)");
      break;
    }
    bool active;
    {
      std::lock_guard<std::mutex> g(mu_);
      if (stop_) break;
      active = !pending_.empty() || !queue_.empty();
    }
    active = active || (opt_.rank == 0 && !table_.empty());
    const uint64_t now_us = mono_us();
    if (active) last_active = now_us;
    // Collectives are in flight - or were a moment ago (a loop of blocking calls leaves the engine idle for a few
    // microseconds between operations): poll the doorbell for a while instead of paying a futex sleep / wake round
    // trip, the latency-critical part of negotiation (83 -> ~35 us per blocking allreduce at np=2 on the CPU box).
    if (active || now_us - last_active < kLingerUs) {
      const uint64_t spin_until = now_us + 300;
      bool rung = false;
      for (int i = 0;; ++i) {
        if (hdr->doorbell.load(std::memory_order_acquire) != db) { rung = true; break; }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        if ((i & 127) == 127 && mono_us() > spin_until) break;
      }
      if (rung) continue;
    }
    futex_wait(&hdr->doorbell, db, opt_.cycle_us);
  }
  fail_all("engine stopped");
}

void Engine::publish_local() {
  ShmHeader* hdr = shm_.hdr();
  RankRing& ring = hdr->rings[opt_.rank];
  bool wrote = false;
  std::vector<std::pair<int64_t, std::string>> failed;
  {
    std::lock_guard<std::mutex> g(mu_);
    while (!queue_.empty()) {
      Entry& e = queue_.front();
      std::string name(e.req.name);
      if (pending_.count(name)) {
        failed.emplace_back(e.handle, "duplicate tensor name '" + name +
                                          "' submitted while a previous operation with that name is in flight");
        queue_.pop_front();
        continue;
      }
      const int cid = cache_enabled_ ? cache_lookup(e.req) : -1;
      if (cid >= 0) {
        // steady state: this exact (name, type, dtype, shape, op, scales) was negotiated before - announce it with
        // one bit; the coordinator ANDs the bit rows of all ranks
        hdr->cache_bits[opt_.rank][cid >> 6].fetch_or(1ull << (cid & 63), std::memory_order_release);
        cache_hits_.fetch_add(1, std::memory_order_relaxed);
        pending_.emplace(name, e);
        queue_.pop_front();
        wrote = true;
        continue;
      }
      if (e.req.type == kAllreduce || e.req.type == kBroadcast) cache_misses_.fetch_add(1, std::memory_order_relaxed);
      uint64_t head = ring.head.load(std::memory_order_relaxed);
      uint64_t tail = ring.tail.load(std::memory_order_acquire);
      if (head - tail >= (uint64_t)kReqRing) break;  // coordinator is behind: retry next cycle
      ring.slots[head % kReqRing] = e.req;
      ring.head.store(head + 1, std::memory_order_release);
      if (e.req.type == kJoin) joined_local_ = true;
      pending_.emplace(name, e);
      queue_.pop_front();
      wrote = true;
    }
  }
  for (auto& f : failed) complete_handle(f.first, f.second);
  if (wrote && opt_.rank != 0) shm_.ring_doorbell();
}

static bool same_signature(const Request& a, const Request& b) {
  if (a.type != b.type || a.dtype != b.dtype || (a.device < 0) != (b.device < 0) || a.op != b.op || a.root != b.root ||
      a.ndim != b.ndim || a.prescale != b.prescale || a.postscale != b.postscale)
    return false;
  for (uint32_t i = 0; i < a.ndim; ++i)
    if (a.shape[i] != b.shape[i]) return false;
  return true;
}

// id of the cache entry that `r` may be announced through, or -1
int Engine::cache_lookup(const Request& r) {
  if (r.type != kAllreduce && r.type != kBroadcast) return -1;
  auto it = cache_index_.find(r.name);
  if (it == cache_index_.end()) return -1;
  return same_signature(cache_[it->second], r) ? it->second : -1;
}

// Called by every rank while executing a cacheable response (same order everywhere => same ids everywhere).
void Engine::cache_insert(const Response& r, const std::vector<Entry*>& es) {
  for (uint32_t k = 0; k < r.n; ++k) {
    if (!es[k]) continue;   // cannot happen for cacheable responses (no rank had joined)
    auto it = cache_index_.find(r.names[k]);
    if (it != cache_index_.end()) {
      cache_[it->second] = es[k]->req;          // signature changed: same id, new signature, on every rank
      if (opt_.rank == 0 && (size_t)it->second < demoted_.size()) demoted_[it->second] = 0;
      continue;
    }
    if ((int)cache_.size() >= kCacheBits) continue;
    cache_index_.emplace(r.names[k], (int)cache_.size());
    cache_.push_back(es[k]->req);
  }
}

static std::string shape_str(const Request& r) {
  std::string s = "[";
  for (uint32_t i = 0; i < r.ndim; ++i) s += (i ? ", " : "") + std::to_string(r.shape[i]);
  return s + "]";
}

void Engine::absorb(int src, const Request& r) {
  const uint64_t bit = 1ull << src;
  if (r.type == kJoin) {
    joined_mask_ |= bit;
    last_joined_ = src;
    return;
  }
  if (r.type == kShutdown) {
    shutdown_mask_ |= bit;
    return;
  }
  std::string name(r.name);
  auto it = table_.find(name);
  const int world = opt_.world;
  {
    // a rank announced a cached name with a full record (its signature changed): the other ranks' bits for that id
    // are folded into the record path until the next response re-establishes the entry
    auto ci = cache_index_.find(name);
    if (ci != cache_index_.end()) {
      if (demoted_.size() < cache_.size()) demoted_.resize(cache_.size(), 0);
      demoted_[ci->second] = 1;
    }
  }
  if (it == table_.end()) {
    Record rec;
    rec.first = r;
    rec.order = order_++;
    rec.first_us = mono_us();
    memset(rec.dim0, 0, sizeof(rec.dim0));
    if (r.type == kAlltoall) rec.splits.assign((size_t)world * world, 0);
    it = table_.emplace(name, std::move(rec)).first;
  }
  Record& rec = it->second;
  const Request& f = rec.first;
  auto fail = [&](const std::string& m) {
    if (rec.error.empty()) rec.error = m;
  };
  if (rec.seen & bit) fail("rank " + std::to_string(src) + " submitted tensor '" + name + "' twice");
  if (r.type != f.type)
    fail("mismatched collective for tensor '" + name + "': " + type_name(f.type) + " vs " + type_name(r.type));
  if (r.dtype != f.dtype) fail("mismatched data types for tensor '" + name + "'");
  if ((r.device < 0) != (f.device < 0)) fail("tensor '" + name + "' is on the CPU on some ranks and on a GPU on others");
  if (r.type == kAllreduce || r.type == kReducescatter) {
    if (r.op != f.op) fail("mismatched reduce ops for tensor '" + name + "'");
    if (r.prescale != f.prescale || r.postscale != f.postscale)
      fail("mismatched prescale/postscale factors for tensor '" + name + "'");
  }
  if (r.type == kBroadcast && r.root != f.root) fail("mismatched root ranks for broadcast '" + name + "'");
  if (r.ndim != f.ndim) {
    fail("mismatched tensor ranks for '" + name + "': " + shape_str(f) + " vs " + shape_str(r));
  } else {
    uint32_t from = (r.type == kAllgather || r.type == kAlltoall) ? 1 : 0;
    for (uint32_t i = from; i < r.ndim; ++i)
      if (r.shape[i] != f.shape[i]) {
        fail("mismatched shapes for tensor '" + name + "': " + shape_str(f) + " vs " + shape_str(r));
        break;
      }
  }
  rec.dim0[src] = r.ndim > 0 ? r.shape[0] : 1;
  if (r.type == kAlltoall && (int)rec.splits.size() == world * world) {
    if (r.has_splits) {
      int64_t sum = 0;
      for (int d = 0; d < world; ++d) {
        rec.splits[(size_t)src * world + d] = r.splits[d];
        sum += r.splits[d];
      }
      if (sum != rec.dim0[src]) fail("alltoall splits of '" + name + "' do not sum to the first dimension");
    } else {
      if (rec.dim0[src] % world != 0) fail("alltoall tensor '" + name + "' first dimension is not divisible by size");
      for (int d = 0; d < world; ++d) rec.splits[(size_t)src * world + d] = (int32_t)(rec.dim0[src] / world);
    }
  }
  rec.seen |= bit;
}

void Engine::coordinate() {
  ShmHeader* hdr = shm_.hdr();
  for (int src = 0; src < opt_.world; ++src) {
    RankRing& ring = hdr->rings[src];
    uint64_t tail = ring.tail.load(std::memory_order_relaxed);
    uint64_t head = ring.head.load(std::memory_order_acquire);
    if (tail == head) continue;
    while (tail < head) {
      absorb(src, ring.slots[tail % kReqRing]);
      ++tail;
    }
    ring.tail.store(tail, std::memory_order_release);
  }
  emit_cached();
  emit_ready();
}

// Response-cache fast path (coordinator): swap the announce words out of shared memory, OR them into the per-rank
// accumulators, AND the rows; every id whose bit survives is ready on all ranks.  Ready ids are emitted in id order,
// consecutive compatible allreduces fused exactly like the record path does.
void Engine::emit_cached() {
  if (cache_.empty()) return;
  ShmHeader* hdr = shm_.hdr();
  const int world = opt_.world;
  const int words = ((int)cache_.size() + 63) / 64;
  bool any = false;
  for (int r = 0; r < world; ++r)
    for (int w = 0; w < words; ++w) {
      if (hdr->cache_bits[r][w].load(std::memory_order_relaxed) == 0) continue;
      const uint64_t b = hdr->cache_bits[r][w].exchange(0, std::memory_order_acq_rel);
      acc_bits_[r][w] |= b;
      any = any || b != 0;
    }
  if (demoted_.size() < cache_.size()) demoted_.resize(cache_.size(), 0);
  if (cache_first_us_.size() < cache_.size()) cache_first_us_.resize(cache_.size(), 0);
  // ranks that joined, or ids demoted to the record path: hand the accumulated announcements to absorb()
  for (int w = 0; w < words; ++w)
    for (int r = 0; r < world; ++r) {
      uint64_t b = acc_bits_[r][w];
      while (b) {
        const int bit = __builtin_ctzll(b);
        b &= b - 1;
        const int id = w * 64 + bit;
        if (joined_mask_ != 0 || demoted_[id]) {
          acc_bits_[r][w] &= ~(1ull << bit);
          absorb(r, cache_[id]);
        } else if (cache_first_us_[id] == 0) {
          cache_first_us_[id] = mono_us();
        }
      }
    }
  if (!any && joined_mask_ == 0) {
    bool pending = false;
    for (int w = 0; w < words && !pending; ++w) pending = acc_bits_[0][w] != 0;
    if (!pending) return;
  }
  std::unique_ptr<Response> resp(new Response);
  std::vector<int> ready;
  for (int w = 0; w < words; ++w) {
    uint64_t m = ~0ull;
    for (int r = 0; r < world; ++r) m &= acc_bits_[r][w];
    while (m) {
      const int bit = __builtin_ctzll(m);
      m &= m - 1;
      ready.push_back(w * 64 + bit);
    }
  }
  size_t i = 0;
  while (i < ready.size()) {
    const Request& f = cache_[ready[i]];
    Response& r = *resp;
    memset(&r, 0, offsetof(Response, numel));
    r.type = f.type; r.dtype = f.dtype; r.device = f.device; r.op = f.op; r.root = f.root;
    r.prescale = f.prescale; r.postscale = f.postscale; r.last_joined = -1;
    r.cacheable = 1; r.from_cache = 1;
    r.n = 1;
    snprintf(r.names[0], kMaxName, "%s", f.name);
    auto numel_of_req = [](const Request& q) { int64_t n = 1; for (uint32_t d = 0; d < q.ndim; ++d) n *= q.shape[d]; return n; };
    r.numel[0] = numel_of_req(f);
    size_t consumed = 1;
    if (f.type == kAllreduce) {
      int64_t bytes = r.numel[0] * (int64_t)dtype_size(f.dtype);
      while (i + consumed < ready.size() && r.n < (uint32_t)kMaxFused) {
        const Request& g = cache_[ready[i + consumed]];
        if (g.type != kAllreduce || g.dtype != f.dtype || g.device != f.device || g.op != f.op ||
            g.prescale != f.prescale || g.postscale != f.postscale)
          break;
        const int64_t nb = numel_of_req(g) * (int64_t)dtype_size(g.dtype);
        if (bytes + nb > opt_.fusion_threshold) break;
        snprintf(r.names[r.n], kMaxName, "%s", g.name);
        r.numel[r.n] = numel_of_req(g);
        bytes += nb;
        ++r.n;
        ++consumed;
      }
    }
    for (size_t k = 0; k < consumed; ++k) {
      const int id = ready[i + k];
      for (int q = 0; q < world; ++q) acc_bits_[q][id >> 6] &= ~(1ull << (id & 63));
      cache_first_us_[id] = 0;
    }
    publish(&r);
    i += consumed;
  }
}

static int64_t numel_of(const Request& r) {
  int64_t n = 1;
  for (uint32_t i = 0; i < r.ndim; ++i) n *= r.shape[i];
  return n;
}

void Engine::emit_ready() {
  const int world = opt_.world;
  const uint64_t full = world == 64 ? ~0ull : ((1ull << world) - 1);
  std::vector<std::map<std::string, Record>::iterator> ready;
  for (auto it = table_.begin(); it != table_.end(); ++it)
    if (((it->second.seen | joined_mask_) & full) == full) ready.push_back(it);
  std::sort(ready.begin(), ready.end(),
            [](const auto& a, const auto& b) { return a->second.order < b->second.order; });
  std::unique_ptr<Response> resp(new Response);
  size_t i = 0;
  while (i < ready.size()) {
    Record& rec = ready[i]->second;
    Response& r = *resp;
    memset(&r, 0, offsetof(Response, numel));
    const Request& f = rec.first;
    r.type = f.type;
    r.dtype = f.dtype;
    r.device = f.device;
    r.op = f.op;
    r.root = f.root;
    r.prescale = f.prescale;
    r.postscale = f.postscale;
    r.last_joined = -1;
    r.n = 1;
    r.cacheable = (joined_mask_ == 0 && (f.type == kAllreduce || f.type == kBroadcast)) ? 1 : 0;
    snprintf(r.names[0], kMaxName, "%s", f.name);
    r.numel[0] = numel_of(f);
    if (rec.error.empty() && joined_mask_ != 0 && f.type != kAllreduce && f.type != kBarrier)
      rec.error = std::string(type_name(f.type)) + " of '" + f.name + "' is not supported while some ranks have joined";
    if (!rec.error.empty()) {
      r.has_error = 1;
      snprintf(r.error, sizeof(r.error), "%s", rec.error.c_str());
      table_.erase(ready[i]);
      publish(&r);
      ++i;
      continue;
    }
    size_t consumed = 1;
    if (f.type == kAllreduce) {
      int64_t bytes = r.numel[0] * (int64_t)dtype_size(f.dtype);
      while (i + consumed < ready.size() && r.n < (uint32_t)kMaxFused) {
        Record& nx = ready[i + consumed]->second;
        const Request& g = nx.first;
        if (!nx.error.empty() || g.type != kAllreduce || g.dtype != f.dtype || g.device != f.device ||
            g.op != f.op || g.prescale != f.prescale || g.postscale != f.postscale)
          break;
        int64_t nb = numel_of(g) * (int64_t)dtype_size(g.dtype);
        if (bytes + nb > opt_.fusion_threshold) break;
        snprintf(r.names[r.n], kMaxName, "%s", g.name);
        r.numel[r.n] = numel_of(g);
        bytes += nb;
        ++r.n;
        ++consumed;
      }
    } else if (f.type == kAllgather || f.type == kAlltoall) {
      int64_t row = 1;
      for (uint32_t d = 1; d < f.ndim; ++d) row *= f.shape[d];
      r.row_elems = row;
      for (int p = 0; p < world; ++p) r.dim0[p] = rec.dim0[p];
      if (f.type == kAlltoall)
        for (int k = 0; k < world * world; ++k) r.splits[k] = rec.splits[k];
    }
    for (size_t k = 0; k < consumed; ++k) {
      auto ci = cache_index_.find(ready[i + k]->first);
      if (ci != cache_index_.end() && (size_t)ci->second < demoted_.size()) demoted_[ci->second] = 0;
      table_.erase(ready[i + k]);
    }
    publish(&r);
    i += consumed;
  }
  if (joined_mask_ != 0 && (joined_mask_ & full) == full) {
    Response& r = *resp;
    memset(&r, 0, offsetof(Response, numel));
    r.type = kJoin;
    r.last_joined = last_joined_;
    r.n = 1;
    snprintf(r.names[0], kMaxName, "__join__");
    joined_mask_ = 0;
    publish(&r);
  }
  if (shutdown_mask_ != 0 && (shutdown_mask_ & full) == full) {
    Response& r = *resp;
    memset(&r, 0, offsetof(Response, numel));
    r.type = kShutdown;
    r.n = 1;
    snprintf(r.names[0], kMaxName, "__shutdown__");
    shutdown_mask_ = 0;
    publish(&r);
  }
}

void Engine::publish(Response* r) {
  ShmHeader* hdr = shm_.hdr();
  uint64_t head = hdr->resp_head.load(std::memory_order_relaxed);
  while (true) {
    uint64_t min_tail = head;
    for (int p = 0; p < opt_.world; ++p)
      min_tail = std::min(min_tail, hdr->resp_tail[p].load(std::memory_order_acquire));
    if (head - min_tail < (uint64_t)kRespRing) break;
    consume_responses();  // make room by draining our own backlog, then wait for the peers
    if (shm_.aborted()) return;
    usleep(50);
  }
  Response& slot = hdr->resp[head % kRespRing];
  // copy only what is used: header, numel/dim0/splits and the first n names
  memcpy(&slot, r, offsetof(Response, names));
  for (uint32_t k = 0; k < r->n; ++k) memcpy(slot.names[k], r->names[k], kMaxName);
  slot.seq = head;
  hdr->resp_head.store(head + 1, std::memory_order_release);
  shm_.ring_doorbell();
}

void Engine::check_stalls() {
  if (opt_.stall_warn_s <= 0 && opt_.stall_shutdown_s <= 0) return;
  const uint64_t now = mono_us();
  const int world = opt_.world;
  for (size_t id = 0; id < cache_first_us_.size(); ++id) {
    if (cache_first_us_[id] == 0) continue;
    const double age = (now - cache_first_us_[id]) * 1e-6;
    const double limit = opt_.stall_shutdown_s > 0 ? opt_.stall_shutdown_s : opt_.stall_warn_s;
    if (limit > 0 && age > limit) {
      // hand a stuck cached announcement to the record path: it carries the warnings / shutdown logic
      if (demoted_.size() < cache_.size()) demoted_.resize(cache_.size(), 0);
      demoted_[id] = 1;
      cache_first_us_[id] = 0;
    }
  }
  const uint64_t full = world == 64 ? ~0ull : ((1ull << world) - 1);
  for (auto& kv : table_) {
    Record& rec = kv.second;
    double age = (now - rec.first_us) * 1e-6;
    if (opt_.stall_warn_s > 0 && age > opt_.stall_warn_s &&
        (now - std::max(rec.last_warn_us, rec.first_us)) * 1e-6 > opt_.stall_warn_s) {
      std::string missing;
      for (int p = 0; p < world; ++p)
        if (!((rec.seen | joined_mask_) & (1ull << p))) missing += (missing.empty() ? "" : ", ") + std::to_string(p);
      fprintf(stderr,
              "[sparkdl engine] WARNING: tensor '%s' (%s) has been waiting %.0fs for ranks [%s]. "
              "One or more ranks submitted different tensors or are much slower.\n",
              kv.first.c_str(), type_name(rec.first.type), age, missing.c_str());
      fflush(stderr);
      rec.last_warn_us = now;
    }
    if (opt_.stall_shutdown_s > 0 && age > opt_.stall_shutdown_s && rec.error.empty()) {
      rec.error = "stalled: tensor '" + kv.first + "' was not submitted by every rank within " +
                  std::to_string((int)opt_.stall_shutdown_s) + "s";
      rec.seen = full;  // force emission as an error on every rank
    }
  }
}

bool Engine::consume_responses() {
  ShmHeader* hdr = shm_.hdr();
  uint64_t tail = hdr->resp_tail[opt_.rank].load(std::memory_order_relaxed);
  while (tail < hdr->resp_head.load(std::memory_order_acquire)) {
    const Response& r = hdr->resp[tail % kRespRing];
    if (r.type == kShutdown) shutdown_seen_ = true;
    execute(r);
    ++tail;
    hdr->resp_tail[opt_.rank].store(tail, std::memory_order_release);
    ++responses_executed_;
  }
  return shutdown_seen_;
}

void Engine::execute(const Response& r) {
  std::vector<Entry> owned;
  std::vector<Entry*> es(r.n, nullptr);
  owned.reserve(r.n);
  {
    std::lock_guard<std::mutex> g(mu_);
    for (uint32_t k = 0; k < r.n; ++k) {
      auto it = pending_.find(r.names[k]);
      if (it != pending_.end()) {
        owned.push_back(it->second);
        pending_.erase(it);
      } else {
        Entry e;
        memset(&e, 0, sizeof(e));
        e.handle = -1;
        owned.push_back(e);
      }
    }
  }
  for (uint32_t k = 0; k < r.n; ++k) es[k] = owned[k].handle >= 0 ? &owned[k] : nullptr;
  const uint64_t t0 = mono_us();
  if (timeline_.enabled())
    for (uint32_t k = 0; k < r.n; ++k)
      timeline_.event(r.names[k], (std::string("NEGOTIATE_") + type_name(r.type)).c_str(), 'E', t0);

  for (auto* e : es)
    if (e && e->enqueue_us) {
      neg_us_sum_.fetch_add(t0 - e->enqueue_us, std::memory_order_relaxed);
      neg_count_.fetch_add(1, std::memory_order_relaxed);
    }
  if (r.has_error) {
    for (auto* e : es)
      if (e) complete(e, r.error);
    return;
  }
  if (r.cacheable && cache_enabled_) cache_insert(r, es);
  const int world = opt_.world, rank = opt_.rank;
  const int64_t esz = (int64_t)dtype_size(r.dtype);
  std::string err;
  const char* phase = "";
  switch (r.type) {
    case kAllreduce:
      phase = r.device >= 0 ? "CUDA_ALLREDUCE" : "SHM_ALLREDUCE";
      if (timeline_.enabled())
        for (uint32_t k = 0; k < r.n; ++k) timeline_.event(r.names[k], phase, 'B', t0);
      exec_allreduce(r, es);
      if (timeline_.enabled()) {
        uint64_t t1 = mono_us();
        for (uint32_t k = 0; k < r.n; ++k) timeline_.event(r.names[k], phase, 'E', t1);
      }
      tensors_fused_ += r.n;
      return;
    case kAllgather: {
      Entry* e = es[0];
      if (!e) { err = "internal: allgather response without a local entry"; break; }
      if (r.device >= 0) {
        if (!gpu_fn_) { complete(e, "CUDA tensor submitted but no GPU executor is registered"); return; }
        // root == 1: the callers sized their output buffers for equal shards (world x own rows) without asking first.
        // The response carries every rank's rows, so all ranks take the same decision here: run when the shards are
        // equal, otherwise complete without touching the buffers - the caller sees out_dim0 != its guess and
        // re-submits with exact sizes (root == 0).  One negotiation instead of two in the common case.
        int64_t rows = 0;
        bool equal = true;
        for (int p = 0; p < world; ++p) {
          rows += r.dim0[p];
          equal = equal && r.dim0[p] == r.dim0[0];
        }
        if (auto st0 = state(e->handle)) st0->out_dim0 = rows;
        if (r.root == 1 && !equal) { complete(e, ""); return; }
        GpuWork w;
        memset(&w, 0, sizeof(w));
        w.type = r.type; w.dtype = r.dtype; w.device = r.device; w.n = 1;
        w.in = &e->in; w.out = &e->out; w.numel = &e->numel; w.ready_events = &e->ready_event;
        w.dim0 = r.dim0; w.row_elems = r.row_elems;
        char eb[512] = {0};
        uint64_t ev = 0;
        int rc = gpu_fn_(gpu_ctx_, &w, &ev, eb, sizeof(eb));
        if (auto st = state(e->handle)) st->done_event = ev;
        complete(e, rc == 0 ? "" : eb);
        return;
      }
      auto st = state(e->handle);
      std::vector<int64_t> bytes(world);
      int64_t total = 0, rows = 0;
      for (int p = 0; p < world; ++p) {
        bytes[p] = r.dim0[p] * r.row_elems * esz;
        total += bytes[p];
        rows += r.dim0[p];
      }
      st->out_buf.resize((size_t)total);
      st->out_dim0 = rows;
      if (!cpu_->allgather(e->in, st->out_buf.data(), bytes.data())) err = "allgather aborted";
      complete(e, err);
      if (!err.empty()) fail_all(err);
      return;
    }
    case kBroadcast: {
      Entry* e = es[0];
      if (!e) { err = "internal: broadcast response without a local entry"; break; }
      if (r.device >= 0) {
        if (!gpu_fn_) { complete(e, "CUDA tensor submitted but no GPU executor is registered"); return; }
        GpuWork w;
        memset(&w, 0, sizeof(w));
        w.type = r.type; w.dtype = r.dtype; w.device = r.device; w.root = r.root; w.n = 1;
        w.in = &e->in; w.out = &e->out; w.numel = &e->numel; w.ready_events = &e->ready_event;
        char eb[512] = {0};
        uint64_t ev = 0;
        int rc = gpu_fn_(gpu_ctx_, &w, &ev, eb, sizeof(eb));
        if (auto st = state(e->handle)) st->done_event = ev;
        complete(e, rc == 0 ? "" : eb);
        return;
      }
      if (e->out != e->in && rank == r.root) memcpy(e->out, e->in, e->numel * esz);
      if (!cpu_->broadcast(e->out, e->numel * esz, r.root)) err = "broadcast aborted";
      complete(e, err);
      if (!err.empty()) fail_all(err);
      return;
    }
    case kAlltoall: {
      Entry* e = es[0];
      if (!e) { err = "internal: alltoall response without a local entry"; break; }
      if (r.device >= 0) {
        if (!gpu_fn_) { complete(e, "CUDA tensor submitted but no GPU executor is registered"); return; }
        GpuWork w;
        memset(&w, 0, sizeof(w));
        w.type = r.type; w.dtype = r.dtype; w.device = r.device; w.n = 1;
        w.in = &e->in; w.out = &e->out; w.numel = &e->numel; w.ready_events = &e->ready_event;
        w.row_elems = r.row_elems; w.splits = r.splits; w.splits_stride = world;
        char eb[512] = {0};
        uint64_t ev = 0;
        int rc = gpu_fn_(gpu_ctx_, &w, &ev, eb, sizeof(eb));
        if (auto st = state(e->handle)) st->done_event = ev;
        complete(e, rc == 0 ? "" : eb);
        return;
      }
      auto st = state(e->handle);
      std::vector<int64_t> m((size_t)world * world);
      int64_t recv_rows = 0;
      st->recv_splits.resize(world);
      for (int s = 0; s < world; ++s)
        for (int d = 0; d < world; ++d) m[(size_t)s * world + d] = (int64_t)r.splits[s * world + d] * r.row_elems * esz;
      for (int s = 0; s < world; ++s) {
        st->recv_splits[s] = r.splits[s * world + rank];
        recv_rows += r.splits[s * world + rank];
      }
      st->out_buf.resize((size_t)(recv_rows * r.row_elems * esz));
      st->out_dim0 = recv_rows;
      if (!cpu_->alltoall(e->in, st->out_buf.data(), m.data(), nullptr)) err = "alltoall aborted";
      complete(e, err);
      if (!err.empty()) fail_all(err);
      return;
    }
    case kReducescatter: {
      Entry* e = es[0];
      if (!e) { err = "internal: reducescatter response without a local entry"; break; }
      if (r.device >= 0) {
        if (!gpu_fn_) { complete(e, "CUDA tensor submitted but no GPU executor is registered"); return; }
        GpuWork w;
        memset(&w, 0, sizeof(w));
        w.type = r.type; w.dtype = r.dtype; w.device = r.device; w.n = 1;
        w.op = r.op == kAverage ? (uint32_t)kSum : r.op;
        w.prescale = r.prescale; w.postscale = r.postscale;
        w.in = &e->in; w.out = &e->out; w.numel = &e->numel; w.ready_events = &e->ready_event;
        int64_t d0v = e->req.ndim > 0 ? e->req.shape[0] : 1;
        w.dim0 = &d0v;
        w.row_elems = d0v > 0 ? e->numel / d0v : 0;
        char eb[512] = {0};
        uint64_t ev = 0;
        int rc = gpu_fn_(gpu_ctx_, &w, &ev, eb, sizeof(eb));
        if (auto st = state(e->handle)) st->done_event = ev;
        complete(e, rc == 0 ? "" : eb);
        return;
      }
      auto st = state(e->handle);
      const int64_t d0 = e->req.ndim > 0 ? e->req.shape[0] : 1;
      const int64_t row = d0 > 0 ? e->numel / d0 : 0;
      std::vector<int64_t> per(world);
      for (int p = 0; p < world; ++p) per[p] = (d0 / world + (p < d0 % world ? 1 : 0)) * row;
      st->out_buf.resize((size_t)(per[rank] * esz));
      st->out_dim0 = d0 / world + (rank < d0 % world ? 1 : 0);
      std::vector<uint8_t> tmp;
      const void* src = e->in;
      if (r.prescale != 1.0) {
        tmp.assign((const uint8_t*)e->in, (const uint8_t*)e->in + e->numel * esz);
        scale_buffer(tmp.data(), e->numel, r.dtype, r.prescale);
        src = tmp.data();
      }
      if (!cpu_->reducescatter(src, st->out_buf.data(), per.data(), r.dtype, r.op == kAverage ? kSum : r.op))
        err = "reducescatter aborted";
      scale_buffer(st->out_buf.data(), per[rank], r.dtype, r.postscale);
      complete(e, err);
      if (!err.empty()) fail_all(err);
      return;
    }
    case kBarrier: {
      if (!cpu_->barrier()) err = "barrier aborted";
      if (es[0]) complete(es[0], err);
      if (!err.empty()) fail_all(err);
      return;
    }
    case kJoin: {
      joined_local_ = false;
      if (es[0]) {
        if (auto st = state(es[0]->handle)) st->last_joined = r.last_joined;
        complete(es[0], "");
      }
      return;
    }
    case kShutdown:
      if (es[0]) complete(es[0], "");
      return;
    default:
      err = "unknown response type";
  }
  for (auto* e : es)
    if (e) complete(e, err);
}

void Engine::exec_allreduce(const Response& r, std::vector<Entry*>& es) {
  const int64_t esz = (int64_t)dtype_size(r.dtype);
  const uint32_t op = r.op == kAverage ? (uint32_t)kSum : r.op;
  if (r.device >= 0) {
    if (!gpu_fn_) {
      for (auto* e : es)
        if (e) complete(e, "CUDA tensor submitted but no GPU executor is registered");
      return;
    }
    std::vector<void*> in(r.n), out(r.n);
    std::vector<int64_t> numel(r.n);
    std::vector<uint64_t> evs(r.n);
    for (uint32_t k = 0; k < r.n; ++k) {
      in[k] = es[k] ? es[k]->in : nullptr;  // nullptr = this rank joined: contribute zeros
      out[k] = es[k] ? es[k]->out : nullptr;
      numel[k] = r.numel[k];
      evs[k] = es[k] ? es[k]->ready_event : 0;
    }
    GpuWork w;
    memset(&w, 0, sizeof(w));
    w.type = kAllreduce; w.dtype = r.dtype; w.op = op; w.device = r.device; w.n = r.n;
    w.prescale = r.prescale; w.postscale = r.postscale;
    w.in = in.data(); w.out = out.data(); w.numel = numel.data(); w.ready_events = evs.data();
    char eb[512] = {0};
    uint64_t ev = 0;
    int rc = gpu_fn_(gpu_ctx_, &w, &ev, eb, sizeof(eb));
    for (auto* e : es) {
      if (!e) continue;
      if (auto st = state(e->handle)) st->done_event = ev;
      complete(e, rc == 0 ? "" : eb);
    }
    return;
  }
  std::string err;
  if (r.n == 1 && es[0]) {
    Entry* e = es[0];
    if (e->out != e->in) memcpy(e->out, e->in, e->numel * esz);
    scale_buffer(e->out, e->numel, r.dtype, r.prescale);
    if (!cpu_->allreduce(e->out, e->numel, r.dtype, op)) err = "allreduce aborted";
    scale_buffer(e->out, e->numel, r.dtype, r.postscale);
  } else {
    int64_t total = 0;
    for (uint32_t k = 0; k < r.n; ++k) total += r.numel[k];
    fuse_buf_.resize((size_t)(total * esz));
    int64_t off = 0;
    for (uint32_t k = 0; k < r.n; ++k) {
      if (es[k])
        memcpy(fuse_buf_.data() + off * esz, es[k]->in, r.numel[k] * esz);
      else
        memset(fuse_buf_.data() + off * esz, 0, r.numel[k] * esz);  // joined rank: zeros
      off += r.numel[k];
    }
    scale_buffer(fuse_buf_.data(), total, r.dtype, r.prescale);
    if (!cpu_->allreduce(fuse_buf_.data(), total, r.dtype, op)) err = "allreduce aborted";
    scale_buffer(fuse_buf_.data(), total, r.dtype, r.postscale);
    off = 0;
    for (uint32_t k = 0; k < r.n; ++k) {
      if (es[k]) memcpy(es[k]->out, fuse_buf_.data() + off * esz, r.numel[k] * esz);
      off += r.numel[k];
    }
  }
  for (auto* e : es)
    if (e) complete(e, err);
  if (!err.empty()) fail_all(err);
}

}  // namespace sdl
