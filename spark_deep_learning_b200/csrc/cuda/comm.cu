#include <cstdlib>

#include "comm.h"

#include <unistd.h>

#include <cstdio>
#include <cstring>

namespace sdlc {

// ------------------------------------------------------------------------------------------
// Driver API entry points, resolved through the runtime so the library loads on a box
// without libcuda (the CPU build host) and needs no link-time dependency on it.
// ------------------------------------------------------------------------------------------
#define SDL_DRV(X)                                                                              \
  X(cuMemCreate) X(cuMemRelease) X(cuMemAddressReserve) X(cuMemAddressFree) X(cuMemMap)          \
  X(cuMemUnmap) X(cuMemSetAccess) X(cuMemExportToShareableHandle)                                \
  X(cuMemImportFromShareableHandle) X(cuMemGetAllocationGranularity) X(cuMulticastCreate)        \
  X(cuMulticastAddDevice) X(cuMulticastBindMem) X(cuMulticastGetGranularity) X(cuMulticastUnbind)\
  X(cuDeviceGet) X(cuDeviceGetAttribute) X(cuGetErrorName)

struct Drv {
#define X(n) decltype(&::n) n = nullptr;
  SDL_DRV(X)
#undef X
  bool ok = false;
} g_drv;

bool load_driver_api(std::string* err) {
  if (g_drv.ok) return true;
  bool all = true;
#define X(n)                                                                                     \
  {                                                                                              \
    void* fp = nullptr;                                                                          \
    cudaDriverEntryPointQueryResult qr;                                                          \
    cudaError_t e = cudaGetDriverEntryPoint(#n, &fp, cudaEnableDefault, &qr);                    \
    if (e != cudaSuccess || fp == nullptr) {                                                     \
      all = false;                                                                               \
      if (err) *err += std::string(#n) + " unavailable; ";                                       \
      cudaGetLastError();                                                                        \
    } else {                                                                                     \
      g_drv.n = reinterpret_cast<decltype(g_drv.n)>(fp);                                         \
    }                                                                                            \
  }
  SDL_DRV(X)
#undef X
  g_drv.ok = all;
  return all;
}

const char* cu_error_name(CUresult r) {
  const char* s = nullptr;
  if (g_drv.cuGetErrorName && g_drv.cuGetErrorName(r, &s) == CUDA_SUCCESS && s) return s;
  return "CUDA_ERROR_?";
}

#define CU_TRY(comm, expr)                                                                       \
  do {                                                                                           \
    CUresult _r = (expr);                                                                        \
    if (_r != CUDA_SUCCESS) {                                                                    \
      (comm)->last_error = std::string(#expr) + " -> " + cu_error_name(_r);                      \
      return -1;                                                                                 \
    }                                                                                            \
  } while (0)

#define RT_TRY(comm, expr)                                                                       \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      (comm)->last_error = std::string(#expr) + " -> " + cudaGetErrorString(_e);                 \
      cudaGetLastError();                                                                        \
      return -1;                                                                                 \
    }                                                                                            \
  } while (0)

CollCtx Comm::make_ctx(int channel, int buf, size_t offset) const {
  CollCtx c;
  memset(&c, 0, sizeof(c));
  const SymmBuffer& b = bufs[buf];
  const SymmBuffer& pad = bufs[pad_buf];
  for (int r = 0; r < world; ++r) {
    c.data.p[r] = static_cast<char*>(b.ptrs[r]) + offset;
    c.pads[r] = static_cast<uint32_t*>(pad.ptrs[r]) + (size_t)channel * kPadWordsPerChannel;
  }
  c.mc = (b.mc && use_mc) ? static_cast<char*>(b.mc) + offset : nullptr;
  c.epoch = epoch + (size_t)channel * kMaxBlocks;
  c.err = err_dev;
  c.rank = rank;
  c.world = world;
  c.timeout_ns = timeout_ns;
  return c;
}

cudaEvent_t Comm::acquire_event(int refs) {
  std::lock_guard<std::mutex> g(mu);
  cudaEvent_t e = nullptr;
  if (!free_events.empty()) {
    e = free_events.back();
    free_events.pop_back();
  } else {
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    all_events.push_back(e);
  }
  event_refs[e] = refs < 1 ? 1 : refs;
  return e;
}

void Comm::release_event(cudaEvent_t e) {
  if (!e) return;
  std::lock_guard<std::mutex> g(mu);
  auto it = event_refs.find(e);
  if (it == event_refs.end()) return;   // not ours, or already back on the free list
  if (--it->second > 0) return;
  event_refs.erase(it);
  free_events.push_back(e);
}

static size_t round_up(size_t x, size_t g) { return (x + g - 1) / g * g; }

}  // namespace sdlc

using namespace sdlc;

extern "C" {

void* sdlc_comm_create(int rank, int world, int device, double timeout_ms, int max_blocks,
                       long long oneshot_max_bytes, int use_mc) {
  if (world < 1 || world > kMaxPeers) return nullptr;
  Comm* c = new Comm();
  c->rank = rank;
  c->world = world;
  c->device = device;
  c->timeout_ns = (unsigned long long)(timeout_ms * 1e6);
  c->max_blocks = max_blocks < 1 ? 1 : (max_blocks > kMaxBlocks ? kMaxBlocks : max_blocks);
  c->oneshot_max_bytes = (size_t)oneshot_max_bytes;
  c->use_mc = use_mc != 0;
  if (const char* e = getenv("SPARKDL_NVLS_MIN_WORLD")) c->nvls_min_world = atoi(e) < 2 ? 2 : atoi(e);
  if (cudaSetDevice(device) != cudaSuccess) {
    delete c;
    return nullptr;
  }
  cudaMalloc(&c->epoch, sizeof(uint32_t) * kChannels * kMaxBlocks);
  cudaMemset(c->epoch, 0, sizeof(uint32_t) * kChannels * kMaxBlocks);
  cudaHostAlloc(&c->err_host, sizeof(uint32_t) * 16, cudaHostAllocMapped);
  memset(c->err_host, 0, sizeof(uint32_t) * 16);
  cudaHostGetDevicePointer(&c->err_dev, c->err_host, 0);
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  for (int i = 0; i < kChannels; ++i) cudaStreamCreateWithPriority(&c->streams[i], cudaStreamNonBlocking, hi);
  // pointer-table ring: 16 slots x 3 arrays x 1024 entries x 8 bytes
  cudaHostAlloc(&c->table_host, 16 * 3 * 1024 * 8, cudaHostAllocDefault);
  cudaMalloc(&c->table_dev, 16 * 3 * 1024 * 8);
  for (auto& e : c->table_events) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  cudaDeviceSynchronize();
  return c;
}

void sdlc_comm_destroy(void* h) {
  Comm* c = static_cast<Comm*>(h);
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (auto& s : c->streams)
    if (s) cudaStreamDestroy(s);
  for (auto& e : c->all_events) cudaEventDestroy(e);
  for (auto& e : c->table_events)
    if (e) cudaEventDestroy(e);
  if (c->epoch) cudaFree(c->epoch);
  if (c->err_host) cudaFreeHost(c->err_host);
  if (c->table_host) cudaFreeHost(c->table_host);
  if (c->table_dev) cudaFree(c->table_dev);
  // symmetric buffers are left to process teardown: peers may still have them mapped
  delete c;
}

const char* sdlc_last_error(void* h) { return static_cast<Comm*>(h)->last_error.c_str(); }

// out[0] = VMM posix-fd export supported, out[1] = multicast supported, out[2] = #devices,
// out[3] = bitmask of devices this device can access as a peer, out[4] = SM count
int sdlc_probe(void* h, int* out) {
  Comm* c = static_cast<Comm*>(h);
  memset(out, 0, sizeof(int) * 8);
  std::string err;
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  out[2] = ndev;
  int mask = 0;
  for (int d = 0; d < ndev && d < 31; ++d) {
    int can = (d == c->device);
    if (d != c->device) cudaDeviceCanAccessPeer(&can, c->device, d);
    if (can) mask |= 1 << d;
  }
  out[3] = mask;
  cudaDeviceGetAttribute(&out[4], cudaDevAttrMultiProcessorCount, c->device);
  if (!load_driver_api(&err)) {
    c->last_error = err;
    return 0;
  }
  CUdevice dev;
  if (g_drv.cuDeviceGet(&dev, c->device) != CUDA_SUCCESS) return 0;
  int v = 0;
  if (g_drv.cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev) == CUDA_SUCCESS) out[0] = v;
  v = 0;
  if (g_drv.cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) == CUDA_SUCCESS) out[1] = v;
  return 0;
}

// ---- symmetric buffers -----------------------------------------------------------------------
int sdlc_buf_alloc(void* h, size_t bytes, int kind) {
  Comm* c = static_cast<Comm*>(h);
  cudaSetDevice(c->device);
  SymmBuffer b;
  b.bytes = bytes;
  b.kind = kind;
  if (kind == 0) {
    b.alloc_bytes = round_up(bytes, 256);
    void* p = nullptr;
    RT_TRY(c, cudaMalloc(&p, b.alloc_bytes));
    RT_TRY(c, cudaMemset(p, 0, b.alloc_bytes));
    b.ptrs[c->rank] = p;
  } else {
    std::string err;
    if (!load_driver_api(&err)) {
      c->last_error = err;
      return -1;
    }
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = c->device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    CU_TRY(c, g_drv.cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    size_t mc_gran = 0;
    if (c->world > 1) {
      CUmulticastObjectProp mp;
      memset(&mp, 0, sizeof(mp));
      mp.numDevices = (unsigned)c->world;
      mp.size = round_up(bytes, gran);
      mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
      if (g_drv.cuMulticastGetGranularity(&mc_gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS) mc_gran = 0;
    }
    if (mc_gran > gran && mc_gran <= (size_t)(64 << 20)) gran = mc_gran;
    b.alloc_bytes = round_up(bytes, gran);
    CU_TRY(c, g_drv.cuMemCreate(&b.handle, b.alloc_bytes, &prop, 0));
    CUdeviceptr va = 0;
    CU_TRY(c, g_drv.cuMemAddressReserve(&va, b.alloc_bytes, gran, 0, 0));
    CU_TRY(c, g_drv.cuMemMap(va, b.alloc_bytes, 0, b.handle, 0));
    CUmemAccessDesc ad;
    memset(&ad, 0, sizeof(ad));
    ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ad.location.id = c->device;
    ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CU_TRY(c, g_drv.cuMemSetAccess(va, b.alloc_bytes, &ad, 1));
    b.ptrs[c->rank] = reinterpret_cast<void*>(va);
    RT_TRY(c, cudaMemset(b.ptrs[c->rank], 0, b.alloc_bytes));
  }
  RT_TRY(c, cudaDeviceSynchronize());
  b.live = true;
  std::lock_guard<std::mutex> g(c->mu);
  c->bufs.push_back(b);
  return (int)c->bufs.size() - 1;
}

int sdlc_buf_export_fd(void* h, int id) {
  Comm* c = static_cast<Comm*>(h);
  SymmBuffer& b = c->bufs[id];
  if (b.kind != 1) {
    c->last_error = "export_fd on a non-VMM buffer";
    return -1;
  }
  int fd = -1;
  CU_TRY(c, g_drv.cuMemExportToShareableHandle(&fd, b.handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  return fd;
}

int sdlc_buf_import_fd(void* h, int id, int peer, int fd) {
  Comm* c = static_cast<Comm*>(h);
  cudaSetDevice(c->device);
  SymmBuffer& b = c->bufs[id];
  CUmemGenericAllocationHandle ph;
  CU_TRY(c, g_drv.cuMemImportFromShareableHandle(&ph, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)),
                                                 CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  CUdeviceptr va = 0;
  CU_TRY(c, g_drv.cuMemAddressReserve(&va, b.alloc_bytes, 0, 0, 0));
  CU_TRY(c, g_drv.cuMemMap(va, b.alloc_bytes, 0, ph, 0));
  CUmemAccessDesc ad;
  memset(&ad, 0, sizeof(ad));
  ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ad.location.id = c->device;
  ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CU_TRY(c, g_drv.cuMemSetAccess(va, b.alloc_bytes, &ad, 1));
  b.peer_handles[peer] = ph;
  b.ptrs[peer] = reinterpret_cast<void*>(va);
  close(fd);
  return 0;
}

int sdlc_buf_export_ipc(void* h, int id, char* out64) {
  Comm* c = static_cast<Comm*>(h);
  SymmBuffer& b = c->bufs[id];
  cudaIpcMemHandle_t mh;
  RT_TRY(c, cudaIpcGetMemHandle(&mh, b.ptrs[c->rank]));
  static_assert(sizeof(mh) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(out64, &mh, 64);
  return 0;
}

int sdlc_buf_import_ipc(void* h, int id, int peer, const char* in64) {
  Comm* c = static_cast<Comm*>(h);
  cudaSetDevice(c->device);
  SymmBuffer& b = c->bufs[id];
  cudaIpcMemHandle_t mh;
  memcpy(&mh, in64, 64);
  void* p = nullptr;
  RT_TRY(c, cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
  b.ptrs[peer] = p;
  return 0;
}

int sdlc_buf_set_peer_ptr(void* h, int id, int peer, void* ptr) {
  static_cast<Comm*>(h)->bufs[id].ptrs[peer] = ptr;
  return 0;
}

void* sdlc_buf_ptr(void* h, int id, int peer) { return static_cast<Comm*>(h)->bufs[id].ptrs[peer]; }
void* sdlc_buf_mc_ptr(void* h, int id) { return static_cast<Comm*>(h)->bufs[id].mc; }
size_t sdlc_buf_bytes(void* h, int id) { return static_cast<Comm*>(h)->bufs[id].bytes; }
size_t sdlc_buf_alloc_bytes(void* h, int id) { return static_cast<Comm*>(h)->bufs[id].alloc_bytes; }

// ---- NVSwitch multicast (NVLS) -------------------------------------------------------------------
int sdlc_mc_create(void* h, int id) {
  Comm* c = static_cast<Comm*>(h);
  SymmBuffer& b = c->bufs[id];
  CUmulticastObjectProp mp;
  memset(&mp, 0, sizeof(mp));
  mp.numDevices = (unsigned)c->world;
  mp.size = b.alloc_bytes;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CU_TRY(c, g_drv.cuMulticastCreate(&b.mc_handle, &mp));
  int fd = -1;
  CU_TRY(c, g_drv.cuMemExportToShareableHandle(&fd, b.mc_handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  return fd;
}

int sdlc_mc_import(void* h, int id, int fd) {
  Comm* c = static_cast<Comm*>(h);
  SymmBuffer& b = c->bufs[id];
  CU_TRY(c, g_drv.cuMemImportFromShareableHandle(&b.mc_handle, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)),
                                                 CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  close(fd);
  return 0;
}

int sdlc_mc_add_device(void* h, int id) {
  Comm* c = static_cast<Comm*>(h);
  SymmBuffer& b = c->bufs[id];
  CUdevice dev;
  CU_TRY(c, g_drv.cuDeviceGet(&dev, c->device));
  CU_TRY(c, g_drv.cuMulticastAddDevice(b.mc_handle, dev));
  return 0;
}

int sdlc_mc_bind(void* h, int id) {
  Comm* c = static_cast<Comm*>(h);
  cudaSetDevice(c->device);
  SymmBuffer& b = c->bufs[id];
  CU_TRY(c, g_drv.cuMulticastBindMem(b.mc_handle, 0, b.handle, 0, b.alloc_bytes, 0));
  b.mc_bound = true;
  CUdeviceptr va = 0;
  CU_TRY(c, g_drv.cuMemAddressReserve(&va, b.alloc_bytes, 0, 0, 0));
  CU_TRY(c, g_drv.cuMemMap(va, b.alloc_bytes, 0, b.mc_handle, 0));
  CUmemAccessDesc ad;
  memset(&ad, 0, sizeof(ad));
  ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ad.location.id = c->device;
  ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CU_TRY(c, g_drv.cuMemSetAccess(va, b.alloc_bytes, &ad, 1));
  b.mc = reinterpret_cast<void*>(va);
  return 0;
}

int sdlc_comm_set_pads(void* h, int id) {
  Comm* c = static_cast<Comm*>(h);
  if (c->bufs[id].bytes < sizeof(uint32_t) * kChannels * kPadWordsPerChannel) {
    c->last_error = "signal pad buffer too small";
    return -1;
  }
  c->pad_buf = id;
  return 0;
}
size_t sdlc_pad_bytes() { return sizeof(uint32_t) * kChannels * kPadWordsPerChannel; }

// receive area of the LL (flag-in-data) allreduce: zero-filled symmetric buffer of sdlc_ll_bytes(); id < 0 disables
int sdlc_comm_set_ll(void* h, int id) {
  Comm* c = static_cast<Comm*>(h);
  if (id >= 0 && c->bufs[id].bytes < kLLChannelBytes * kChannels) {
    c->last_error = "LL buffer too small";
    return -1;
  }
  c->ll_buf = id;
  return 0;
}
size_t sdlc_ll_bytes() { return kLLChannelBytes * kChannels; }

int sdlc_comm_set_staging(void* h, int channel, int id) {
  static_cast<Comm*>(h)->staging_buf[channel] = id;
  return 0;
}
void sdlc_comm_set_use_mc(void* h, int on) { static_cast<Comm*>(h)->use_mc = on != 0; }
void sdlc_comm_set_max_blocks(void* h, int n) {
  static_cast<Comm*>(h)->max_blocks = n < 1 ? 1 : (n > kMaxBlocks ? kMaxBlocks : n);
}
void sdlc_comm_set_oneshot_max(void* h, long long n) { static_cast<Comm*>(h)->oneshot_max_bytes = (size_t)n; }

// ---- streams / events -------------------------------------------------------------------------------
unsigned long long sdlc_channel_stream(void* h, int ch) {
  return reinterpret_cast<unsigned long long>(static_cast<Comm*>(h)->streams[ch]);
}

unsigned long long sdlc_record_event(void* h, unsigned long long stream) {
  Comm* c = static_cast<Comm*>(h);
  cudaEvent_t e = c->acquire_event();
  if (!e || cudaEventRecord(e, reinterpret_cast<cudaStream_t>(stream)) != cudaSuccess) {
    c->last_error = "cudaEventRecord failed";
    cudaGetLastError();
    c->release_event(e);
    return 0;
  }
  return reinterpret_cast<unsigned long long>(e);
}

// Give an event from sdlc_record_event / the engine executor back to the pool.  Legal as soon as every
// cudaStreamWaitEvent / cudaEventSynchronize that needs it has been *issued* (a wait binds to the record that
// preceded it, so a later re-record cannot disturb it).
void sdlc_release_event(void* h, unsigned long long event) {
  static_cast<Comm*>(h)->release_event(reinterpret_cast<cudaEvent_t>(event));
}
unsigned long long sdlc_event_pool_size(void* h) {
  Comm* c = static_cast<Comm*>(h);
  std::lock_guard<std::mutex> g(c->mu);
  return c->all_events.size();
}

int sdlc_stream_wait_event(unsigned long long stream, unsigned long long event) {
  return cudaStreamWaitEvent(reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<cudaEvent_t>(event), 0) == cudaSuccess ? 0 : -1;
}

int sdlc_event_synchronize(unsigned long long event) {
  return cudaEventSynchronize(reinterpret_cast<cudaEvent_t>(event)) == cudaSuccess ? 0 : -1;
}

unsigned int sdlc_check_error(void* h) {
  return *reinterpret_cast<volatile uint32_t*>(static_cast<Comm*>(h)->err_host);
}
void sdlc_clear_error(void* h) { *reinterpret_cast<volatile uint32_t*>(static_cast<Comm*>(h)->err_host) = 0; }
unsigned long long sdlc_launch_count(void* h) { return static_cast<Comm*>(h)->launches; }
int sdlc_max_peers() { return kMaxPeers; }

}  // extern "C"
