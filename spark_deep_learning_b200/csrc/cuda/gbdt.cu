// Histogram GBDT kernels for sparkdl.xgboost (SURVEY.md §2f K8, §3.5 "HOT LOOP: per tree, per level:
// build histograms -> ALLREDUCE -> split").
//
// Gradients are quantised to int32 fixed point per boosting round, so histogram bins are exact
// int64 sums: atomics commute, results are bit-reproducible run to run and independent of the
// number of workers.  `gbdt_hist_allreduce_kernel` is the fused op: every rank accumulates its
// row shard into its symmetric histogram, the CTAs of a rank meet at a device-local barrier, and
// the SAME launch then runs the two-shot peer-memory reduction (rank r sums slice r from all peers
// and pushes the result to everyone) - one launch per tree level instead of hist kernel + NCCL.
#include <cstring>

#include "comm.h"

namespace sdlc {

// bins: [n, F] uint8 row-major; gh: [n] int2 (quantised grad, hess); pos: [n] node slot in this
// level or -1; hist: [nodes, F, 256, 2] int64.
__device__ __forceinline__ void hist_accumulate(const uint8_t* __restrict__ bins, const int2* __restrict__ gh,
                                                const int* __restrict__ pos, unsigned long long* hist, long long n,
                                                int F, long long tid, long long nth) {
  const long long total = n * F;
  for (long long i = tid; i < total; i += nth) {
    const long long row = i / F;
    const int f = (int)(i - row * F);
    const int node = pos[row];
    if (node < 0) continue;
    const int2 v = gh[row];
    unsigned long long* h = hist + (((long long)node * F + f) * 256 + bins[i]) * 2;
    atomicAdd(h, (unsigned long long)(long long)v.x);
    atomicAdd(h + 1, (unsigned long long)(long long)v.y);
  }
}

__global__ void __launch_bounds__(kThreads)
gbdt_hist_kernel(const uint8_t* __restrict__ bins, const int2* __restrict__ gh, const int* __restrict__ pos,
                 unsigned long long* hist, long long n, int F) {
  hist_accumulate(bins, gh, pos, hist, n, F, (long long)blockIdx.x * blockDim.x + threadIdx.x,
                  (long long)gridDim.x * blockDim.x);
}

// K8 fused: local build -> device-local grid barrier -> cross-rank two-shot int64 reduction.
// The build phase wants the whole GPU (one 1024-thread CTA per SM, all co-resident for the local barrier);
// the cross-rank phase is run by the first `cross_blocks` CTAs only (flag slots exist for kMaxBlocks CTAs).
constexpr int kHistThreads = 1024;
__global__ void __launch_bounds__(kHistThreads)
gbdt_hist_allreduce_kernel(CollCtx c, const uint8_t* __restrict__ bins, const int2* __restrict__ gh,
                           const int* __restrict__ pos, long long n, int F, long long hist_elems,
                           unsigned int* grid_ctr, unsigned int grid_epoch, int cross_blocks) {
  unsigned long long* hist = static_cast<unsigned long long*>(c.data.p[c.rank]);
  hist_accumulate(bins, gh, pos, hist, n, F, (long long)blockIdx.x * blockDim.x + threadIdx.x,
                  (long long)gridDim.x * blockDim.x);
  // all CTAs of this rank must have finished their atomics before any peer reads the histogram
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(grid_ctr, 1u);
    const unsigned int target = grid_epoch * gridDim.x;
    const unsigned long long t0 = globaltimer_ns();
    while ((int)(*reinterpret_cast<volatile unsigned int*>(grid_ctr) - target) < 0) {
      if (globaltimer_ns() - t0 > c.timeout_ns) {
        *reinterpret_cast<volatile uint32_t*>(c.err) = 0x40000000u | (uint32_t)c.rank;
        break;
      }
    }
    __threadfence();
  }
  __syncthreads();
  if ((int)blockIdx.x >= cross_blocks) return;
  EpochScope ep(c);
  coll_barrier(c, ep.base + 1);
  const long long nvec = (hist_elems + 1) / 2;  // 2 int64 per 16-byte vector
  const int world = c.world;
  const long long per = (nvec + world - 1) / world;
  const long long lo = min(nvec, (long long)c.rank * per), hi = min(nvec, lo + per);
  const long long stride = (long long)cross_blocks * blockDim.x;
  for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
    long long a0 = 0, a1 = 0;
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p) {
      if (p < world) {
        const uint4 v = ld_peer_v4(static_cast<const uint4*>(c.data.p[p]) + i);
        a0 += (long long)(((unsigned long long)v.y << 32) | v.x);
        a1 += (long long)(((unsigned long long)v.w << 32) | v.z);
      }
    }
    const uint4 o = make_uint4((uint32_t)a0, (uint32_t)((unsigned long long)a0 >> 32), (uint32_t)a1,
                               (uint32_t)((unsigned long long)a1 >> 32));
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p)
      if (p < world) st_peer_v4(static_cast<uint4*>(c.data.p[p]) + i, o);
  }
  coll_barrier(c, ep.base + 2);
  ep.commit(c, 2);
}

// Tree ensemble inference on raw float features.  Node arrays are concatenated over trees;
// tree t starts at node offset tree_off[t]; out[row, group[t]] += leaf value.
__global__ void __launch_bounds__(256)
gbdt_predict_kernel(const float* __restrict__ x, long long n, int F, float missing, int missing_is_nan,
                    const int* __restrict__ feat, const float* __restrict__ thr, const int* __restrict__ left,
                    const int* __restrict__ right, const uint8_t* __restrict__ dleft, const float* __restrict__ value,
                    const int* __restrict__ tree_off, const int* __restrict__ group, int n_trees, int n_groups,
                    float* __restrict__ out) {
  for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += (long long)gridDim.x * blockDim.x) {
    const float* xr = x + row * F;
    for (int t = 0; t < n_trees; ++t) {
      const int base = tree_off[t];
      int node = 0;
      while (left[base + node] >= 0) {
        const float v = xr[feat[base + node]];
        const bool miss = missing_is_nan ? isnan(v) : (v == missing || isnan(v));
        const bool go_left = miss ? (dleft[base + node] != 0) : (v < thr[base + node]);
        node = go_left ? left[base + node] : right[base + node];
      }
      out[row * n_groups + group[t]] += value[base + node];
    }
  }
}

}  // namespace sdlc

using namespace sdlc;

extern "C" {

int sdlc_gbdt_hist(const void* bins, const void* gh, const void* pos, void* hist, long long n, int F,
                   unsigned long long stream) {
  long long work = n * F;
  long long blocks = (work + kThreads * 8 - 1) / (kThreads * 8);
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 4) blocks = 148 * 4;
  gbdt_hist_kernel<<<(int)blocks, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(bins), static_cast<const int2*>(gh), static_cast<const int*>(pos),
      static_cast<unsigned long long*>(hist), n, F);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// hist lives in symmetric buffer `buf` (zeroed by the caller on the same stream before the launch)
int sdlc_gbdt_hist_allreduce(void* h, int channel, int buf, const void* bins, const void* gh, const void* pos, long long n,
                             int F, long long hist_elems, void* grid_ctr, unsigned int grid_epoch, unsigned long long stream) {
  Comm* c = static_cast<Comm*>(h);
  // the caller's stream as given (0 = the legacy default stream): the launch must stay ordered with the
  // memset before it and the reads after it
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const CollCtx ctx = c->make_ctx(channel, buf, 0);
  // every CTA must be co-resident (device-local barrier): exactly one 1024-thread CTA per SM
  int blocks = 0;
  cudaDeviceGetAttribute(&blocks, cudaDevAttrMultiProcessorCount, c->device);
  if (blocks < 1) blocks = 1;
  const int cross = blocks < kMaxBlocks ? blocks : kMaxBlocks;
  gbdt_hist_allreduce_kernel<<<blocks, kHistThreads, 0, s>>>(ctx, static_cast<const uint8_t*>(bins), static_cast<const int2*>(gh),
                                                             static_cast<const int*>(pos), n, F, hist_elems,
                                                             static_cast<unsigned int*>(grid_ctr), grid_epoch, cross);
  if (cudaGetLastError() != cudaSuccess) {
    c->last_error = "gbdt_hist_allreduce launch failed";
    return -1;
  }
  c->launches++;
  return blocks;
}

int sdlc_gbdt_predict(const float* x, long long n, int F, float missing, int missing_is_nan, const int* feat,
                      const float* thr, const int* left, const int* right, const unsigned char* dleft, const float* value,
                      const int* tree_off, const int* group, int n_trees, int n_groups, float* out, unsigned long long stream) {
  long long blocks = (n + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  gbdt_predict_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, n, F, missing, missing_is_nan, feat, thr, left, right, dleft, value, tree_off, group, n_trees, n_groups, out);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // extern "C"
