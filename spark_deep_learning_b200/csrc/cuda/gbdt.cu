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

// ---- shared-memory privatised build (round 2) --------------------------------------------------------------------------
// The global-atomic kernel above issues 2 int64 atomics per (row, feature): 128 M per level at 1 M rows x 64 features,
// which is the L2 atomic rate limit (0.79 ms).  Here a CTA owns (row chunk, feature tile): the histogram of EVERY node of
// the level for FT features lives in shared memory (nodes x FT x 256 bins x 16 B <= 192 KB), rows stream through with
// shared-memory atomics, and only the non-zero bins are flushed with global atomics at the end (a few percent of the
// original count).  binsT is the FEATURE-major copy of the bin matrix ([F, pitch], pitch % 16 == 0) so that a thread reads
// the bins of 4 consecutive rows of one feature with one 4-byte load; the row -> (node, g, h) loads are shared by all FT
// features of the tile.
//
// sm_100 has no native 64-bit shared-memory add (atom.shared.add.u64 compiles to an ATOMS.CAST.SPIN compare-and-swap
// loop), so every 64-bit counter is two 32-bit words in separate planes: the low word takes a native ATOMS.ADD that
// returns the old value, the wrap-around of that add is the carry, and the high word takes `hi + carry` (skipped when
// zero).  Every wrap produces exactly one carry, so the pair equals the 64-bit sum whatever the interleaving - the result
// stays bit-exact and order independent.  Planes (instead of interleaved words) keep the bank of a counter = bin % 32.
constexpr int kHistThreads = 1024;
constexpr int kSmemHistBytes = 192 * 1024;
constexpr int kMaxTileSlots = kSmemHistBytes / (256 * 16);   // (node, feature) pairs per CTA: 48

__device__ __forceinline__ void smem_add64(unsigned* lo_plane, unsigned* hi_plane, int cell, int v) {
  const unsigned lo = (unsigned)v, hi = v < 0 ? 0xffffffffu : 0u;
  const unsigned old = atomicAdd(lo_plane + cell, lo);
  const unsigned add = hi + ((old + lo) < old ? 1u : 0u);
  if (add) atomicAdd(hi_plane + cell, add);
}

__device__ __forceinline__ void hist_build_smem(const uint8_t* __restrict__ binsT, long long pitch, const int2* __restrict__ gh,
                                                const int* __restrict__ pos, unsigned long long* __restrict__ hist, long long n,
                                                int F, int nodes, int FT, int tiles, int row_chunks, unsigned* sh) {
  const int tile = (int)(blockIdx.x % tiles), chunk = (int)(blockIdx.x / tiles);
  if (chunk >= row_chunks) return;
  const int f0 = tile * FT, nf = min(FT, F - f0);
  const int cells = nodes * FT * 256;
  unsigned *g_lo = sh, *g_hi = sh + cells, *h_lo = sh + 2 * cells, *h_hi = sh + 3 * cells;
  for (int i = threadIdx.x; i < 4 * cells; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  const long long rows_per = (((n + row_chunks - 1) / row_chunks) + 3) & ~3ll;
  const long long r0 = (long long)chunk * rows_per, r1 = min(n, r0 + rows_per);
  for (long long base = r0 + 4ll * threadIdx.x; base < r1; base += 4ll * blockDim.x) {
    if (base + 4 <= r1) {
      const int4 p = *reinterpret_cast<const int4*>(pos + base);
      if ((p.x & p.y & p.z & p.w) < 0) continue;            // all four rows inactive
      const int4 a = *reinterpret_cast<const int4*>(gh + base), b = *reinterpret_cast<const int4*>(gh + base + 2);
      for (int j = 0; j < nf; ++j) {
        const uchar4 q = *reinterpret_cast<const uchar4*>(binsT + (size_t)(f0 + j) * pitch + base);
        if (p.x >= 0) { const int c = (p.x * FT + j) * 256 + q.x; smem_add64(g_lo, g_hi, c, a.x); smem_add64(h_lo, h_hi, c, a.y); }
        if (p.y >= 0) { const int c = (p.y * FT + j) * 256 + q.y; smem_add64(g_lo, g_hi, c, a.z); smem_add64(h_lo, h_hi, c, a.w); }
        if (p.z >= 0) { const int c = (p.z * FT + j) * 256 + q.z; smem_add64(g_lo, g_hi, c, b.x); smem_add64(h_lo, h_hi, c, b.y); }
        if (p.w >= 0) { const int c = (p.w * FT + j) * 256 + q.w; smem_add64(g_lo, g_hi, c, b.z); smem_add64(h_lo, h_hi, c, b.w); }
      }
    } else {
      for (long long row = base; row < r1; ++row) {
        const int node = pos[row];
        if (node < 0) continue;
        const int2 v = gh[row];
        for (int j = 0; j < nf; ++j) {
          const int c = (node * FT + j) * 256 + binsT[(size_t)(f0 + j) * pitch + row];
          smem_add64(g_lo, g_hi, c, v.x);
          smem_add64(h_lo, h_hi, c, v.y);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cells; i += blockDim.x) {
    const int slot = i >> 8, node = slot / FT, j = slot - node * FT;
    if (j >= nf) continue;
    const unsigned long long g = ((unsigned long long)g_hi[i] << 32) | g_lo[i], h = ((unsigned long long)h_hi[i] << 32) | h_lo[i];
    unsigned long long* o = hist + (((size_t)node * F + f0 + j) * 256 + (i & 255)) * 2;
    if (g) atomicAdd(o, g);
    if (h) atomicAdd(o + 1, h);
  }
}

__global__ void __launch_bounds__(kHistThreads)
gbdt_hist_smem_kernel(const uint8_t* __restrict__ binsT, long long pitch, const int2* __restrict__ gh,
                      const int* __restrict__ pos, unsigned long long* hist, long long n, int F, int nodes, int FT, int tiles,
                      int row_chunks) {
  extern __shared__ unsigned hist_sh[];
  hist_build_smem(binsT, pitch, gh, pos, hist, n, F, nodes, FT, tiles, row_chunks, hist_sh);
}

// K8 fused: local build -> device-local grid barrier -> cross-rank two-shot int64 reduction.
// The build phase wants the whole GPU (one 1024-thread CTA per SM, all co-resident for the local barrier);
// the cross-rank phase is run by the first `cross_blocks` CTAs only (flag slots exist for kMaxBlocks CTAs).
__global__ void __launch_bounds__(kHistThreads)
gbdt_hist_allreduce_kernel(CollCtx c, const uint8_t* __restrict__ bins, const uint8_t* __restrict__ binsT, long long pitch,
                           const int2* __restrict__ gh, const int* __restrict__ pos, long long n, int F, long long hist_elems,
                           unsigned int* grid_ctr, unsigned int grid_epoch, int cross_blocks, int nodes, int FT, int tiles,
                           int row_chunks) {
  extern __shared__ unsigned hist_sh[];
  unsigned long long* hist = static_cast<unsigned long long*>(c.data.p[c.rank]);
  if (binsT != nullptr && FT > 0)
    hist_build_smem(binsT, pitch, gh, pos, hist, n, F, nodes, FT, tiles, row_chunks, hist_sh);
  else
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

// tiling of the shared-memory build for `nodes` node slots: FT features per CTA (0: does not fit, use global atomics)
static void smem_tiling(int F, int nodes, int ctas, int* FT, int* tiles, int* row_chunks) {
  *FT = nodes > 0 ? kMaxTileSlots / nodes : 0;
  if (*FT > F) *FT = F;
  if (*FT < 1) { *FT = 0; *tiles = 0; *row_chunks = 0; return; }
  *tiles = (F + *FT - 1) / *FT;
  *row_chunks = ctas / *tiles;
  if (*row_chunks < 1) *row_chunks = 1;
}

// hist [nodes, F, 256, 2] int64, zeroed by the caller; binsT: feature-major bins [F, pitch] (pitch % 16 == 0).  Returns 0, or 1 when the
// level has too many nodes for the shared-memory tiling (the caller falls back to sdlc_gbdt_hist).
int sdlc_gbdt_hist_smem(const void* binsT, long long pitch, const void* gh, const void* pos, void* hist, long long n, int F, int nodes,
                        unsigned long long stream) {
  int FT, tiles, chunks;
  smem_tiling(F, nodes, 148, &FT, &tiles, &chunks);
  if (FT == 0) return 1;
  if (cudaFuncSetAttribute(gbdt_hist_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemHistBytes) != cudaSuccess) return -1;
  const size_t smem = (size_t)nodes * FT * 256 * 16;
  gbdt_hist_smem_kernel<<<tiles * chunks, kHistThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(binsT), pitch, static_cast<const int2*>(gh), static_cast<const int*>(pos),
      static_cast<unsigned long long*>(hist), n, F, nodes, FT, tiles, chunks);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// hist lives in symmetric buffer `buf` (zeroed by the caller on the same stream before the launch); binsT (optional):
// feature-major bins for the shared-memory build, nodes = node slots of this level
int sdlc_gbdt_hist_allreduce(void* h, int channel, int buf, const void* bins, const void* binsT, long long pitch, const void* gh,
                             const void* pos,
                             long long n, int F, int nodes, long long hist_elems, void* grid_ctr, unsigned int grid_epoch,
                             unsigned long long stream) {
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
  int FT = 0, tiles = 0, chunks = 0;
  if (binsT != nullptr) smem_tiling(F, nodes, blocks, &FT, &tiles, &chunks);
  // per-device attribute (several devices may be driven from one process): set on every launch, it is a host-side table write
  if (FT > 0 && cudaFuncSetAttribute(gbdt_hist_allreduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemHistBytes) != cudaSuccess)
    FT = 0;
  const size_t smem = FT > 0 ? (size_t)nodes * FT * 256 * 16 : 0;
  gbdt_hist_allreduce_kernel<<<blocks, kHistThreads, smem, s>>>(ctx, static_cast<const uint8_t*>(bins),
                                                                static_cast<const uint8_t*>(binsT), pitch, static_cast<const int2*>(gh),
                                                                static_cast<const int*>(pos), n, F, hist_elems,
                                                                static_cast<unsigned int*>(grid_ctr), grid_epoch, cross, nodes,
                                                                FT, tiles, chunks);
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
