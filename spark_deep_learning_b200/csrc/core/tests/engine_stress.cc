// Native stress test of the engine (SURVEY.md §5.2 "TSAN build of the C++ engine for T3").
// W processes (fork) x 2 submitter threads each: every thread submits its share of named allreduces in a
// rank-dependent order plus allgathers/broadcasts, waits, checks the sums.  Build with -fsanitize=thread to have
// ThreadSanitizer watch the engine thread, the submitters and the handle table.
//   usage: engine_stress <world> <tensors>      exit code 0 = all results correct on every rank
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../engine.h"

using namespace sdl;

static int run_rank(const std::string& shm, int rank, int world, int tensors) {
  Engine eng;
  EngineOptions o;
  o.shm_name = shm;
  o.rank = rank;
  o.world = world;
  o.stage_bytes = 1 << 16;
  o.fusion_threshold = 1 << 14;
  o.stall_warn_s = 0;
  std::string err = eng.init(o);
  if (!err.empty()) {
    fprintf(stderr, "rank %d init failed: %s\n", rank, err.c_str());
    return 2;
  }
  int bad = 0;
  // Rounds 0-2 repeat the same names and shapes (rounds 1-2 go through the response cache: one bit per tensor instead of
  // a request record); round 3 changes the length of every third tensor (cached signature no longer matches: the entry
  // is demoted to the record path and re-established); round 4 repeats round 3's shapes (cache hits again).
  auto worker = [&](int tid, int round) {
    std::vector<std::vector<float>> bufs;
    std::vector<int64_t> handles;
    std::vector<int> ids;
    for (int k = 0; k < tensors; ++k) {
      int i = (rank % 2 == 0) ? k : tensors - 1 - k;   // odd ranks submit in reverse order
      if (i % 2 != tid) continue;
      ids.push_back(i);
      const int len = 17 + i + ((round >= 3 && i % 3 == 0) ? 5 : 0);
      bufs.emplace_back(len, (float)(rank + i));
    }
    for (size_t j = 0; j < ids.size(); ++j) {
      Request r;
      memset(&r, 0, sizeof(r));
      r.type = kAllreduce;
      r.dtype = kF32;
      r.device = -1;
      r.op = kSum;
      r.ndim = 1;
      r.shape[0] = (int64_t)bufs[j].size();
      r.prescale = r.postscale = 1.0;
      snprintf(r.name, kMaxName, "t%d", ids[j]);
      handles.push_back(eng.enqueue(r, bufs[j].data(), bufs[j].data(), (int64_t)bufs[j].size(), 0));
    }
    for (size_t j = 0; j < ids.size(); ++j) {
      if (eng.wait(handles[j], 60.0) != 1) {
        __sync_fetch_and_add(&bad, 1);
        continue;
      }
      const float expect = (float)(world * ids[j]) + (float)(world * (world - 1) / 2);
      for (float v : bufs[j])
        if (v != expect) {
          __sync_fetch_and_add(&bad, 1);
          break;
        }
      eng.release(handles[j]);
    }
  };
  for (int round = 0; round < 5; ++round) {
    std::thread t0(worker, 0, round), t1(worker, 1, round);
    t0.join();
    t1.join();
  }
  if (eng.cache_hits() == 0) {
    fprintf(stderr, "rank %d: the response cache was never hit\n", rank);
    ++bad;
  }
  // one broadcast and one ragged allgather from the main thread
  {
    std::vector<int32_t> b(100, rank == 1 % world ? 7 : -1);
    Request r;
    memset(&r, 0, sizeof(r));
    r.type = kBroadcast; r.dtype = kI32; r.device = -1; r.root = 1 % world; r.ndim = 1; r.shape[0] = 100;
    r.prescale = r.postscale = 1.0;
    snprintf(r.name, kMaxName, "bcast");
    int64_t h = eng.enqueue(r, b.data(), b.data(), 100, 0);
    if (eng.wait(h, 60.0) != 1 || b[99] != 7) ++bad;
    eng.release(h);
    std::vector<int64_t> g(rank + 1, rank);
    memset(&r, 0, sizeof(r));
    r.type = kAllgather; r.dtype = kI64; r.device = -1; r.ndim = 1; r.shape[0] = rank + 1;
    r.prescale = r.postscale = 1.0;
    snprintf(r.name, kMaxName, "gather");
    h = eng.enqueue(r, g.data(), nullptr, rank + 1, 0);
    if (eng.wait(h, 60.0) != 1) ++bad;
    auto st = eng.state(h);
    if (!st || (int64_t)st->out_buf.size() != (int64_t)(world * (world + 1) / 2) * 8) ++bad;
    eng.release(h);
  }
  eng.shutdown();
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  const int world = argc > 1 ? atoi(argv[1]) : 3;
  const int tensors = argc > 2 ? atoi(argv[2]) : 64;
  char name[64];
  snprintf(name, sizeof(name), "/sparkdl-stress-%d", (int)getpid());
  std::vector<pid_t> kids;
  for (int r = 1; r < world; ++r) {
    pid_t p = fork();
    if (p == 0) _exit(run_rank(name, r, world, tensors));
    kids.push_back(p);
  }
  int rc = run_rank(name, 0, world, tensors);
  for (pid_t p : kids) {
    int st = 0;
    waitpid(p, &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = rc ? rc : 3;
  }
  printf("engine_stress world=%d tensors=%d -> %s\n", world, tensors, rc == 0 ? "OK" : "FAILED");
  return rc;
}
