// Host-side communicator: symmetric buffers (CUDA VMM posix-fd / cudaIpc), NVSwitch multicast
// binding, signal pads, communication streams and an event pool.  SURVEY.md §5.8 / §7 step 4.
//
// The cross-process exchange of descriptors (file descriptors over SCM_RIGHTS, IPC handles
// through the store) is orchestrated from Python (ops/comm.py); this layer does the CUDA part
// and owns everything the kernels need at launch time.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "sdl_cuda.cuh"

namespace sdlc {

struct SymmBuffer {
  size_t bytes = 0;          // usable bytes
  size_t alloc_bytes = 0;    // rounded to allocation granularity
  void* ptrs[kMaxPeers] = {nullptr};
  void* mc = nullptr;
  int kind = 0;              // 0 cudaMalloc (ipc / local), 1 VMM
  CUmemGenericAllocationHandle handle = 0;
  CUmemGenericAllocationHandle mc_handle = 0;
  CUmemGenericAllocationHandle peer_handles[kMaxPeers] = {0};
  bool mc_bound = false;
  bool live = false;
};

struct Comm {
  int rank = 0, world = 1, device = 0;
  std::vector<SymmBuffer> bufs;
  int pad_buf = -1;                       // symmetric signal-pad buffer
  int ll_buf = -1;                        // symmetric receive area of the flag-in-data small-message path (or -1)
  int staging_buf[kChannels] = {-1, -1, -1, -1};
  uint32_t* epoch = nullptr;              // device: kChannels * kMaxBlocks counters
  uint32_t* err_host = nullptr;           // pinned + mapped
  uint32_t* err_dev = nullptr;
  cudaStream_t streams[kChannels] = {nullptr};
  unsigned long long timeout_ns = 20000000000ull;
  int max_blocks = 32;
  size_t oneshot_max_bytes = 512 << 10;
  bool use_mc = true;
  // In-switch reduction pays when there are >= 3 copies to fold: at 2 ranks NVLS moves 1.5 S per direction through
  // the links, the P2P kernels 1.0 S (NVML link counters, benchmarks/nvlink_counters.py).  SPARKDL_NVLS_MIN_WORLD.
  int nvls_min_world = 3;
  std::mutex mu;
  // Owned events: an event handed out by acquire_event() is never re-recorded until every holder has
  // released it (a round-robin ring could wrap onto an event that a pending cudaStreamWaitEvent still needs).
  std::vector<cudaEvent_t> all_events;
  std::vector<cudaEvent_t> free_events;
  std::unordered_map<cudaEvent_t, int> event_refs;
  // pinned pointer-table ring for batched pack/unpack
  void* table_host = nullptr;
  void* table_dev = nullptr;
  cudaEvent_t table_events[16] = {nullptr};
  int table_next = 0;
  std::string last_error;
  unsigned long long launches = 0;        // kernels launched by this communicator

  CollCtx make_ctx(int channel, int buf, size_t offset) const;
  cudaEvent_t acquire_event(int refs = 1);
  void release_event(cudaEvent_t e);
};

const char* cu_error_name(CUresult r);
bool load_driver_api(std::string* err);

}  // namespace sdlc
