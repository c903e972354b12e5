// sm_100a peer-memory collectives: every op is ONE kernel that synchronises with its peers
// through flag words in peer memory and moves data with P2P loads/stores or NVLS multimem
// instructions over NVSwitch.  No NCCL anywhere.  Inventory: SURVEY.md §2f K1 (one-shot),
// K2 (two-shot P2P + NVLS), K3 (pack/unpack), K4 (allgather), K5 (broadcast), K6/K7 (16-bit and
// block-scaled fp8 wire formats), K9 (barrier), K10 (reducescatter / alltoall).
#include <cstdio>
#include <cstring>

#include "comm.h"

namespace sdlc {

__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// =====================================================================================================
// K9  barrier only
// =====================================================================================================
__global__ void __launch_bounds__(kThreads) barrier_kernel(CollCtx c) {
  EpochScope ep(c);
  coll_barrier(c, ep.base + 1);
  ep.commit(c, 1);
}

// shared reduction step: v[p] holds the 16-byte vector of rank p, fixed summation order 0..W-1 so
// every rank computes bit-identical results
template <typename T>
__device__ __forceinline__ uint4 reduce_vectors(const uint4* v, int world, int op, float pre, float post) {
  using Tr = VecTraits<T>;
  typename Tr::Acc acc[Tr::N], t[Tr::N];
  vec_unpack<T>(v[0], acc);
#pragma unroll
  for (int i = 0; i < Tr::N; ++i) acc[i] = scale_val(acc[i], pre);
#pragma unroll
  for (int p = 1; p < kMaxPeers; ++p) {
    if (p < world) {
      vec_unpack<T>(v[p], t);
#pragma unroll
      for (int i = 0; i < Tr::N; ++i) acc[i] = reduce_op(acc[i], scale_val(t[i], pre), op);
    }
  }
#pragma unroll
  for (int i = 0; i < Tr::N; ++i) acc[i] = scale_val(acc[i], post);
  return vec_pack<T>(acc);
}

// =====================================================================================================
// K1  one-shot allreduce, in place, latency regime.  Every rank reads all peers and keeps the
// result in registers across the second barrier, then overwrites its own buffer.
// =====================================================================================================
constexpr int kOneShotR = 4;

template <typename T>
__global__ void __launch_bounds__(kThreads)
allreduce_oneshot_kernel(CollCtx c, long long nvec, int op, float pre, float post) {
  EpochScope ep(c);
  coll_barrier(c, ep.base + 1);
  uint4 res[kOneShotR];
  const int world = c.world;
#pragma unroll
  for (int r = 0; r < kOneShotR; ++r) {
    const long long idx = ((long long)r * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx < nvec) {
      uint4 v[kMaxPeers];
#pragma unroll
      for (int p = 0; p < kMaxPeers; ++p)
        if (p < world) v[p] = ld_peer_v4(static_cast<const uint4*>(c.data.p[p]) + idx);
      res[r] = reduce_vectors<T>(v, world, op, pre, post);
    }
  }
  coll_barrier(c, ep.base + 2);  // all peers have finished reading this block's region
  uint4* mine = static_cast<uint4*>(c.data.p[c.rank]);
#pragma unroll
  for (int r = 0; r < kOneShotR; ++r) {
    const long long idx = ((long long)r * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx < nvec) mine[idx] = res[r];
  }
  ep.commit(c, 2);
}

// =====================================================================================================
// K1b  LL allreduce (<= 64 KB): no barrier at all.  Every rank pushes its vector into a receive slot of every peer as
// {data32, flag32} pairs (64-bit stores are single-copy atomic, so a pair is never seen half-written) and spins on its
// own slots until the flag of this launch shows up in all of them: one NVLink store latency instead of the two
// signal-pad round trips of the one-shot kernel (13.4 -> ~6 us at np=8).  The flag is the per-block epoch counter, so
// stale slots never match; slots alternate between two halves by flag parity, and a rank can only reach the launch
// that reuses a half after every peer has sent (hence finished reading) the launch in between.  Fixed summation order.
// =====================================================================================================
template <typename T>
__global__ void __launch_bounds__(kThreads)
allreduce_ll_kernel(CollCtx c, PeerPtrs ll, const char* ll_mine, uint4* data_mine, long long nvec, int op, float pre, float post) {
  EpochScope ep(c);
  const uint32_t flag = ep.base + 1;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < nvec) {
    const int world = c.world;
    const size_t half = (size_t)(flag & 1u) * kMaxPeers * kLLVecs;
    uint4* data = data_mine + idx;
    const uint4 mine = *data;
    const unsigned long long f = (unsigned long long)flag << 32;
    const unsigned long long w0 = f | mine.x, w1 = f | mine.y, w2 = f | mine.z, w3 = f | mine.w;
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p)
      if (p < world && p != c.rank) {
        char* dst = static_cast<char*>(ll.p[p]) + (half + (size_t)c.rank * kLLVecs + idx) * kLLSlotBytes;
        st_peer_v2_u64(dst, w0, w1);
        st_peer_v2_u64(dst + 16, w2, w3);
      }
    uint4 v[kMaxPeers];
    const unsigned long long t0 = globaltimer_ns();
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p) {
      if (p >= world) continue;
      if (p == c.rank) { v[p] = mine; continue; }
      const char* src = ll_mine + (half + (size_t)p * kLLVecs + idx) * kLLSlotBytes;
      unsigned long long a, b, d, e;
      unsigned spins = 0;
      for (;;) {
        ld_peer_v2_u64(src, a, b);
        ld_peer_v2_u64(src + 16, d, e);
        if ((uint32_t)(a >> 32) == flag && (uint32_t)(b >> 32) == flag && (uint32_t)(d >> 32) == flag && (uint32_t)(e >> 32) == flag) break;
        if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > c.timeout_ns) {
          *reinterpret_cast<volatile uint32_t*>(c.err) = 0x80000000u | ((uint32_t)p << 8) | (uint32_t)c.rank;
          break;
        }
      }
      v[p] = make_uint4((uint32_t)a, (uint32_t)b, (uint32_t)d, (uint32_t)e);
    }
    *data = reduce_vectors<T>(v, world, op, pre, post);
  }
  ep.commit(c, 1);
}

// =====================================================================================================
// K2  two-shot allreduce over P2P, in place: rank r reduces slice r from all peers and pushes the
// result straight into every peer's buffer (reduce-scatter and all-gather in one pass).
// =====================================================================================================
template <typename T>
__global__ void __launch_bounds__(kThreads)
allreduce_twoshot_kernel(CollCtx c, long long nvec, int op, float pre, float post) {
  EpochScope ep(c);
  coll_barrier(c, ep.base + 1);
  const int world = c.world;
  const long long per = ceil_div_ll(nvec, world);
  const long long lo = min(nvec, (long long)c.rank * per), hi = min(nvec, lo + per);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += 2 * stride) {
    const long long i1 = i0 + stride;
    const bool has1 = i1 < hi;
    uint4 a[kMaxPeers], b[kMaxPeers];
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p)
      if (p < world) a[p] = ld_peer_v4(static_cast<const uint4*>(c.data.p[p]) + i0);
    if (has1) {
#pragma unroll
      for (int p = 0; p < kMaxPeers; ++p)
        if (p < world) b[p] = ld_peer_v4(static_cast<const uint4*>(c.data.p[p]) + i1);
    }
    const uint4 ra = reduce_vectors<T>(a, world, op, pre, post);
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p)
      if (p < world) st_peer_v4(static_cast<uint4*>(c.data.p[p]) + i0, ra);
    if (has1) {
      const uint4 rb = reduce_vectors<T>(b, world, op, pre, post);
#pragma unroll
      for (int p = 0; p < kMaxPeers; ++p)
        if (p < world) st_peer_v4(static_cast<uint4*>(c.data.p[p]) + i1, rb);
    }
  }
  coll_barrier(c, ep.base + 2);
  ep.commit(c, 2);
}

// =====================================================================================================
// K2  two-shot allreduce over NVLS: multimem.ld_reduce lets the switch sum the N copies, the fused
// scale runs in registers, multimem.st replicates the result to every GPU.  Sum only.
// =====================================================================================================
template <typename T> __device__ __forceinline__ uint4 mc_ld_reduce(const void* p);
template <> __device__ __forceinline__ uint4 mc_ld_reduce<float>(const void* p) { return mc_ld_reduce_f32(p); }
template <> __device__ __forceinline__ uint4 mc_ld_reduce<__nv_bfloat16>(const void* p) { return mc_ld_reduce_bf16(p); }
template <> __device__ __forceinline__ uint4 mc_ld_reduce<__half>(const void* p) { return mc_ld_reduce_f16(p); }

template <typename T>
__device__ __forceinline__ uint4 scale_vector(const uint4& v, float s) {
  using Tr = VecTraits<T>;
  typename Tr::Acc a[Tr::N];
  vec_unpack<T>(v, a);
#pragma unroll
  for (int i = 0; i < Tr::N; ++i) a[i] = scale_val(a[i], s);
  return vec_pack<T>(a);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
allreduce_multimem_kernel(CollCtx c, long long nvec, float scale) {
  EpochScope ep(c);
  coll_barrier(c, ep.base + 1);
  const long long per = ceil_div_ll(nvec, c.world);
  const long long lo = min(nvec, (long long)c.rank * per), hi = min(nvec, lo + per);
  const long long stride = (long long)gridDim.x * blockDim.x;
  uint4* mc = static_cast<uint4*>(c.mc);
  const bool do_scale = scale != 1.0f;
  constexpr int U = 8;  // independent ld_reduce in flight per thread (NVLink latency ~2 us)
  for (long long i0 = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + u * stride < hi) v[u] = mc_ld_reduce<T>(mc + i0 + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + u * stride < hi) mc_st_v4(mc + i0 + u * stride, do_scale ? scale_vector<T>(v[u], scale) : v[u]);
  }
  coll_barrier(c, ep.base + 2);
  ep.commit(c, 2);
}

// =====================================================================================================
// K6/K7  wire-format allreduce: storage dtype (fp32 / bf16) stays local, what crosses NVLink is a
// 16-bit cast or block-scaled fp8 (e4m3 payload + one fp32 scale per 128 elements), accumulated in
// fp32.  Quantise -> barrier -> reduce slice + requantise + push -> barrier -> dequantise: one kernel.
// =====================================================================================================
constexpr int kUnit = 128;  // elements per scale block; one HALF-warp handles one unit, 8 elements per lane

template <typename S> __device__ __forceinline__ void load8(const S* p, float* f);
template <> __device__ __forceinline__ void load8<float>(const float* p, float* f) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float* f) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  bf16x2_to_f(v.x, f[0], f[1]); bf16x2_to_f(v.y, f[2], f[3]); bf16x2_to_f(v.z, f[4], f[5]); bf16x2_to_f(v.w, f[6], f[7]);
}
template <typename S> __device__ __forceinline__ void store8(S* p, const float* f);
template <> __device__ __forceinline__ void store8<float>(float* p, const float* f) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float* f) {
  *reinterpret_cast<uint4*>(p) = make_uint4(f_to_bf16x2(f[0], f[1]), f_to_bf16x2(f[2], f[3]), f_to_bf16x2(f[4], f[5]), f_to_bf16x2(f[6], f[7]));
}

// max over the 16 lanes of a half-warp.  The two halves of a warp walk different units and may leave their
// loops at different times, so the shuffle names ONLY the lanes of the caller's half (a full-warp mask
// would deadlock when the other half has already exited).
__device__ __forceinline__ float halfwarp_max(float v) {
  const unsigned mask = (threadIdx.x & 16) ? 0xffff0000u : 0x0000ffffu;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(mask, v, o));
  return v;
}

__device__ __forceinline__ uint32_t quant_fp8x4(const float* f, float inv) {
  const __nv_fp8x2_storage_t lo = __nv_cvt_float2_to_fp8x2(make_float2(f[0] * inv, f[1] * inv), __NV_SATFINITE, __NV_E4M3);
  const __nv_fp8x2_storage_t hi = __nv_cvt_float2_to_fp8x2(make_float2(f[2] * inv, f[3] * inv), __NV_SATFINITE, __NV_E4M3);
  return (uint32_t)lo | ((uint32_t)hi << 16);
}
__device__ __forceinline__ void dequant_fp8x4(uint32_t q, float s, float* f) {
  const __half2_raw a = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(q & 0xffff), __NV_E4M3);
  const __half2_raw b = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(q >> 16), __NV_E4M3);
  const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&a));
  const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&b));
  f[0] = fa.x * s; f[1] = fa.y * s; f[2] = fb.x * s; f[3] = fb.y * s;
}

// wire layouts (hl = lane within the half-warp, 8 elements each):
//   16-bit: unit u at byte u*256, lane at +16*hl.   fp8: payload unit u at byte u*128 (+8*hl), scales at
//   byte n_units*128 + 4u (one fp32 per 128-element block).
struct WireRaw { uint4 q; float s; };

template <int WIRE>
__device__ __forceinline__ void wire_store(char* base, long long n_units, long long u, int hl, const float* f, bool remote) {
  if (WIRE == kWireFP8) {
    float amax = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) amax = fmaxf(amax, fabsf(f[i]));
    amax = halfwarp_max(amax);
    const float s = amax > 0.f ? amax * (1.0f / 448.0f) : 1.0f;
    const float inv = 1.0f / s;
    const uint2 q = make_uint2(quant_fp8x4(f, inv), quant_fp8x4(f + 4, inv));
    if (remote) {
      st_peer_v2(base + u * kUnit + hl * 8, q);
      if (hl == 0) st_peer_u32(base + n_units * kUnit + u * 4, __float_as_uint(s));
    } else {
      *reinterpret_cast<uint2*>(base + u * kUnit + hl * 8) = q;
      if (hl == 0) *reinterpret_cast<float*>(base + n_units * kUnit + u * 4) = s;
    }
  } else {
    uint4 q;
    if (WIRE == kWireBF16) q = make_uint4(f_to_bf16x2(f[0], f[1]), f_to_bf16x2(f[2], f[3]), f_to_bf16x2(f[4], f[5]), f_to_bf16x2(f[6], f[7]));
    else q = make_uint4(f_to_f16x2(f[0], f[1]), f_to_f16x2(f[2], f[3]), f_to_f16x2(f[4], f[5]), f_to_f16x2(f[6], f[7]));
    if (remote) st_peer_v4(base + u * (kUnit * 2) + hl * 16, q);
    else *reinterpret_cast<uint4*>(base + u * (kUnit * 2) + hl * 16) = q;
  }
}

// issue the loads of one unit (raw bits); decode later so that the loads of all peers are in flight together
template <int WIRE>
__device__ __forceinline__ WireRaw wire_load_raw(const char* base, long long n_units, long long u, int hl) {
  WireRaw r;
  if (WIRE == kWireFP8) {
    const uint2 q = ld_peer_v2(base + u * kUnit + hl * 8);
    r.q = make_uint4(q.x, q.y, 0u, 0u);
    r.s = __uint_as_float(ld_peer_u32(base + n_units * kUnit + u * 4));
  } else {
    r.q = ld_peer_v4(base + u * (kUnit * 2) + hl * 16);
    r.s = 1.f;
  }
  return r;
}
template <int WIRE>
__device__ __forceinline__ void wire_decode(const WireRaw& r, float* f) {
  if (WIRE == kWireFP8) {
    dequant_fp8x4(r.q.x, r.s, f);
    dequant_fp8x4(r.q.y, r.s, f + 4);
  } else if (WIRE == kWireBF16) {
    bf16x2_to_f(r.q.x, f[0], f[1]); bf16x2_to_f(r.q.y, f[2], f[3]); bf16x2_to_f(r.q.z, f[4], f[5]); bf16x2_to_f(r.q.w, f[6], f[7]);
  } else {
    f16x2_to_f(r.q.x, f[0], f[1]); f16x2_to_f(r.q.y, f[2], f[3]); f16x2_to_f(r.q.z, f[4], f[5]); f16x2_to_f(r.q.w, f[6], f[7]);
  }
}

template <typename S, int WIRE>
__global__ void __launch_bounds__(kThreads)
allreduce_wire_kernel(CollCtx c, S* storage, long long n, float pre, float post) {
  EpochScope ep(c);
  const int world = c.world, hl = threadIdx.x & 15, hw = threadIdx.x >> 4, hpb = blockDim.x >> 4;
  const long long n_units = ceil_div_ll(n, kUnit);
  const long long upr = ceil_div_ll(n_units, world);
  char* my_wire = static_cast<char*>(c.data.p[c.rank]);
  const long long step = (long long)gridDim.x * hpb;
  // phase 0: quantise my gradients into my wire buffer.  Block b owns the units j (within every
  // slice) with j % gridDim.x == b - exactly the units block b of the slice owner will read.
  for (int s = 0; s < world; ++s) {
    const long long s_lo = min(n_units, (long long)s * upr), s_len = min(n_units, s_lo + upr) - s_lo;
    for (long long j = blockIdx.x + (long long)gridDim.x * hw; j < s_len; j += step) {
      const long long u = s_lo + j, e = u * kUnit + hl * 8;
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (e < n) {                                  // n % 8 == 0 (checked on the host)
        load8<S>(storage + e, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] *= pre;
      }
      wire_store<WIRE>(my_wire, n_units, u, hl, f, false);
    }
  }
  coll_barrier(c, ep.base + 1);
  // phase 1: reduce my slice over all peers (fp32), requantise, push to every peer's wire buffer
  {
    const long long s_lo = min(n_units, (long long)c.rank * upr), s_len = min(n_units, s_lo + upr) - s_lo;
    for (long long j = blockIdx.x + (long long)gridDim.x * hw; j < s_len; j += step) {
      const long long u = s_lo + j;
      WireRaw raw[kMaxPeers];
#pragma unroll
      for (int p = 0; p < kMaxPeers; ++p)
        if (p < world) raw[p] = wire_load_raw<WIRE>(static_cast<const char*>(c.data.p[p]), n_units, u, hl);
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int p = 0; p < kMaxPeers; ++p) {
        if (p < world) {
          float f[8];
          wire_decode<WIRE>(raw[p], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += f[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] *= post;
#pragma unroll
      for (int p = 0; p < kMaxPeers; ++p)
        if (p < world) wire_store<WIRE>(static_cast<char*>(c.data.p[p]), n_units, u, hl, acc, true);
    }
  }
  coll_barrier(c, ep.base + 2);
  // phase 2: dequantise the reduced wire buffer back into my storage
  for (int s = 0; s < world; ++s) {
    const long long s_lo = min(n_units, (long long)s * upr), s_len = min(n_units, s_lo + upr) - s_lo;
    for (long long j = blockIdx.x + (long long)gridDim.x * hw; j < s_len; j += step) {
      const long long u = s_lo + j, e = u * kUnit + hl * 8;
      float f[8];
      wire_decode<WIRE>(wire_load_raw<WIRE>(my_wire, n_units, u, hl), f);
      if (e < n) store8<S>(storage + e, f);
    }
  }
  ep.commit(c, 2);
}

// =====================================================================================================
// generic byte copy helper (width chosen from the common alignment of dst, src and n)
// =====================================================================================================
__device__ __forceinline__ void copy_bytes(char* dst, const char* src, long long n, long long tid, long long nthreads,
                                           bool src_remote) {
  const unsigned long long al = (unsigned long long)(uintptr_t)dst | (unsigned long long)(uintptr_t)src;
  if ((al & 15) == 0) {
    const long long nv = n >> 4;
    for (long long i = tid; i < nv; i += nthreads) {
      const uint4 v = src_remote ? ld_peer_v4(src + (i << 4)) : *reinterpret_cast<const uint4*>(src + (i << 4));
      *reinterpret_cast<uint4*>(dst + (i << 4)) = v;
    }
    for (long long i = (nv << 4) + tid; i < n; i += nthreads) dst[i] = src[i];
  } else if ((al & 3) == 0) {
    const long long nv = n >> 2;
    for (long long i = tid; i < nv; i += nthreads)
      *reinterpret_cast<uint32_t*>(dst + (i << 2)) = src_remote ? ld_peer_u32(src + (i << 2)) : *reinterpret_cast<const uint32_t*>(src + (i << 2));
    for (long long i = (nv << 2) + tid; i < n; i += nthreads) dst[i] = src[i];
  } else {
    for (long long i = tid; i < n; i += nthreads) dst[i] = src[i];
  }
}

// =====================================================================================================
// K4/K5/K10  pull kernel: allgather, broadcast and alltoall are all "copy a byte range out of each
// peer's symmetric staging buffer into my local output".
// =====================================================================================================
struct PullTable {
  long long src_off[kMaxPeers];
  long long nbytes[kMaxPeers];
  long long dst_off[kMaxPeers];
};

__global__ void __launch_bounds__(kThreads) pull_kernel(CollCtx c, char* out, PullTable t) {
  EpochScope ep(c);
  coll_barrier(c, ep.base + 1);
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  for (int k = 0; k < c.world; ++k) {
    const int p = (c.rank + k) % c.world;  // start with my own chunk, then walk the ring
    if (t.nbytes[p] > 0)
      copy_bytes(out + t.dst_off[p], static_cast<const char*>(c.data.p[p]) + t.src_off[p], t.nbytes[p], tid, nth, p != c.rank);
  }
  coll_barrier(c, ep.base + 2);
  ep.commit(c, 2);
}

// K5  broadcast through the switch: the root writes its payload once with multimem.st, NVSwitch
// replicates it into every rank's staging buffer, then each rank copies staging -> output locally.
// 16-byte read / write of possibly unaligned user memory (alignment is a per-rank property, so it must
// never influence which kernel or how many barriers a rank runs)
__device__ __forceinline__ uint4 load16_any(const char* p) {
  const uintptr_t a = (uintptr_t)p;
  if ((a & 15) == 0) return *reinterpret_cast<const uint4*>(p);
  uint4 v;
  if ((a & 3) == 0) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
    v = make_uint4(q[0], q[1], q[2], q[3]);
  } else {
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i >> 2] |= (uint32_t)(uint8_t)p[i] << ((i & 3) * 8);
    v = make_uint4(w[0], w[1], w[2], w[3]);
  }
  return v;
}
__device__ __forceinline__ void store16_any(char* p, const uint4& v) {
  const uintptr_t a = (uintptr_t)p;
  if ((a & 15) == 0) { *reinterpret_cast<uint4*>(p) = v; return; }
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  if ((a & 3) == 0) {
    uint32_t* q = reinterpret_cast<uint32_t*>(p);
    q[0] = w[0]; q[1] = w[1]; q[2] = w[2]; q[3] = w[3];
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) p[i] = (char)((w[i >> 2] >> ((i & 3) * 8)) & 0xff);
  }
}

__global__ void __launch_bounds__(kThreads)
bcast_multimem_kernel(CollCtx c, const char* src, char* out, long long nbytes, int root) {
  EpochScope ep(c);
  coll_barrier(c, ep.base + 1);
  const long long nvec = nbytes >> 4;  // caller guarantees a 16-byte multiple (rank-independent condition)
  const long long per = ceil_div_ll(nvec, gridDim.x);
  const long long lo = min(nvec, (long long)blockIdx.x * per), hi = min(nvec, lo + per);
  if (c.rank == root) {
    uint4* mc = static_cast<uint4*>(c.mc);
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) mc_st_v4(mc + i, load16_any(src + (i << 4)));
  }
  coll_barrier(c, ep.base + 2);
  const uint4* mine = static_cast<const uint4*>(c.data.p[c.rank]);
  if (out != nullptr && (c.rank != root || out != src))
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) store16_any(out + (i << 4), ld_peer_v4(mine + i));
  coll_barrier(c, ep.base + 3);  // nobody reuses staging until every rank has drained it
  ep.commit(c, 3);
}

// =====================================================================================================
// K10  reduce-scatter: rank r reduces elements [lo, lo+cnt) of every peer's staging buffer into its
// local output (phase 1 of the two-shot algorithm on its own).
// =====================================================================================================
template <typename T> __device__ __forceinline__ T ld_scalar_peer(const T* p) { return __ldcv(p); }
template <> __device__ __forceinline__ __nv_bfloat16 ld_scalar_peer(const __nv_bfloat16* p) {
  unsigned short u = __ldcv(reinterpret_cast<const unsigned short*>(p));
  return *reinterpret_cast<__nv_bfloat16*>(&u);
}
template <> __device__ __forceinline__ __half ld_scalar_peer(const __half* p) {
  unsigned short u = __ldcv(reinterpret_cast<const unsigned short*>(p));
  return *reinterpret_cast<__half*>(&u);
}
template <typename T> __device__ __forceinline__ typename VecTraits<T>::Acc to_acc(T v) { return (typename VecTraits<T>::Acc)v; }
template <> __device__ __forceinline__ float to_acc(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_acc(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_acc(typename VecTraits<T>::Acc v) { return (T)v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_acc(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_acc(float v) { return __float2half_rn(v); }

template <typename T>
__global__ void __launch_bounds__(kThreads)
reducescatter_kernel(CollCtx c, T* out, long long lo, long long cnt, int op, float pre, float post) {
  using Tr = VecTraits<T>;
  EpochScope ep(c);
  coll_barrier(c, ep.base + 1);
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  const int world = c.world;
  const bool vec_ok = ((lo * (long long)sizeof(T)) & 15) == 0 && (cnt % Tr::N) == 0 && (((uintptr_t)out) & 15) == 0;
  if (vec_ok) {
    const long long v0 = lo / Tr::N, nv = cnt / Tr::N;
    for (long long i = tid; i < nv; i += nth) {
      uint4 v[kMaxPeers];
#pragma unroll
      for (int p = 0; p < kMaxPeers; ++p)
        if (p < world) v[p] = ld_peer_v4(static_cast<const uint4*>(c.data.p[p]) + v0 + i);
      reinterpret_cast<uint4*>(out)[i] = reduce_vectors<T>(v, world, op, pre, post);
    }
  } else {
    for (long long i = tid; i < cnt; i += nth) {
      typename Tr::Acc acc = scale_val(to_acc<T>(ld_scalar_peer<T>(static_cast<const T*>(c.data.p[0]) + lo + i)), pre);
      for (int p = 1; p < world; ++p)
        acc = reduce_op(acc, scale_val(to_acc<T>(ld_scalar_peer<T>(static_cast<const T*>(c.data.p[p]) + lo + i)), pre), op);
      out[i] = from_acc<T>(scale_val(acc, post));
    }
  }
  coll_barrier(c, ep.base + 2);
  ep.commit(c, 2);
}

// =====================================================================================================
// K3  batched pack / unpack between user tensors and the symmetric staging buffer (pointer table in
// device memory; blockIdx.y = table entry).  A null user pointer packs zeros (joined rank).
// =====================================================================================================
__global__ void __launch_bounds__(256)
pack_kernel(char* staging, const unsigned long long* ptrs, const long long* offs, const long long* sizes, int to_staging) {
  const int e = blockIdx.y;
  char* user = reinterpret_cast<char*>(ptrs[e]);
  char* st = staging + offs[e];
  const long long n = sizes[e];
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  if (to_staging) {
    if (user == nullptr) {
      for (long long i = tid; i < n; i += nth) st[i] = 0;
    } else {
      copy_bytes(st, user, n, tid, nth, false);
    }
  } else if (user != nullptr) {
    copy_bytes(user, st, n, tid, nth, false);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) scale_kernel(T* p, long long n, float s) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = from_acc<T>(scale_val(to_acc<T>(p[i]), s));
}

// ---- host-side dispatch ----------------------------------------------------------------------------------
// CTAs of a bandwidth kernel.  `max_blocks` (SPARKDL_COMM_CTAS, default 32) is tuned for 8 ranks, where every thread
// keeps loads to 7 peers in flight; with fewer peers a CTA has proportionally fewer bytes in flight over NVLink
// (round 1: 357-400 GB/s at np=2 with 32 CTAs vs 508 GB/s for NCCL), so the cap scales with 8 / world.  The result
// depends on rank-independent values only (the in-kernel barrier pairs CTA b of every rank).
static int pick_blocks(const Comm* c, long long work_items, int per_block, int base_cap = 0) {
  long long b = ceil_div_ll(work_items, per_block);
  if (b < 1) b = 1;
  long long cap = base_cap > 0 ? base_cap : c->max_blocks;
  if (c->world > 1 && c->world < 8) cap = cap * (8 / c->world);
  if (cap > kMaxBlocks) cap = kMaxBlocks;
  if (b > cap) b = cap;
  return (int)b;
}

#define LAUNCH_CHECK(c)                                                                          \
  do {                                                                                           \
    cudaError_t _e = cudaGetLastError();                                                         \
    if (_e != cudaSuccess) {                                                                     \
      (c)->last_error = std::string("kernel launch failed: ") + cudaGetErrorString(_e);          \
      return -1;                                                                                 \
    }                                                                                            \
    (c)->launches++;                                                                             \
  } while (0)

template <typename T>
static int launch_allreduce_t(Comm* c, const CollCtx& ctx, long long numel, int op, float pre, float post, int algo,
                              cudaStream_t s) {
  const long long nvec = ceil_div_ll(numel * (long long)sizeof(T), 16);
  const long long bytes = nvec * 16;
  constexpr bool mc_type = std::is_same<T, float>::value || std::is_same<T, __nv_bfloat16>::value || std::is_same<T, __half>::value;
  // LL pays while (world - 1) x message stays small: 64 KB at 2 ranks, ~18 KB at 8 (13.6 us NVLS vs 14.3 us LL at 64 KB, np=8)
  const bool ll_ok = c->world > 1 && c->ll_buf >= 0 && nvec > 0 && nvec <= kLLVecs;      // fits the receive slots
  const bool ll_pays = ll_ok && nvec * (c->world - 1) <= 2 * kLLVecs;
  if (algo == 4 && !ll_ok) algo = 0;
  if (algo == 0) {
    if (c->world == 1) algo = 1;
    else if (ll_pays) algo = 4;
    else if (bytes <= (long long)c->oneshot_max_bytes && nvec <= (long long)c->max_blocks * kThreads * kOneShotR) algo = 1;
    // in-switch reduction halves the bytes a rank receives only when there are >= 3 peers to fold: at 2 ranks the
    // P2P two-shot moves the same bytes and measures 1.6x faster (417 vs 696 us at 256 MB, sweep of round 2)
    else if (ctx.mc != nullptr && mc_type && op == kSum && c->world >= c->nvls_min_world) algo = 3;
    else algo = 2;
  }
  if (algo == 1) {
    if (nvec > (long long)kMaxBlocks * kThreads * kOneShotR) algo = 2;
  }
  if (algo == 3 && (ctx.mc == nullptr || !mc_type || op != kSum)) {
    c->last_error = "multimem allreduce needs a multicast-bound buffer, a float type and op=sum";
    return -1;
  }
  if (algo == 4) {
    PeerPtrs ll;
    const int channel = (int)((ctx.epoch - c->epoch) / kMaxBlocks);
    for (int p = 0; p < kMaxPeers; ++p)
      ll.p[p] = p < c->world ? static_cast<char*>(c->bufs[c->ll_buf].ptrs[p]) + (size_t)channel * kLLChannelBytes : nullptr;
    const int threads = nvec < kThreads ? (int)((nvec + 31) / 32 * 32) : kThreads;
    allreduce_ll_kernel<T><<<(int)ceil_div_ll(nvec, threads), threads, 0, s>>>(ctx, ll, static_cast<const char*>(ll.p[c->rank]),
                                                                              static_cast<uint4*>(ctx.data.p[c->rank]), nvec, op, pre, post);
  } else if (algo == 1) {
    int blocks = (int)ceil_div_ll(nvec, (long long)kThreads * kOneShotR);
    if (blocks < 1) blocks = 1;
    allreduce_oneshot_kernel<T><<<blocks, kThreads, 0, s>>>(ctx, nvec, op, pre, post);
  } else if (algo == 2) {
    const int blocks = pick_blocks(c, ceil_div_ll(nvec, c->world), kThreads * 2);
    allreduce_twoshot_kernel<T><<<blocks, kThreads, 0, s>>>(ctx, nvec, op, pre, post);
  } else {
    if constexpr (mc_type) {
      // up to 64 MB fewer CTAs are faster through the switch (np=8 sweep of round 2: 649 / 587 / 533 GB/s bus bandwidth
      // at 32 MB with 16 / 32 / 64 CTAs, 732 / 694 / 646 at 64 MB); from 128 MB on the configured count wins
      const int cap = (bytes <= (64ll << 20) && c->max_blocks > 16) ? 16 : 0;
      const int blocks = pick_blocks(c, ceil_div_ll(nvec, c->world), kThreads * 8, cap);
      allreduce_multimem_kernel<T><<<blocks, kThreads, 0, s>>>(ctx, nvec, pre * post);
    }
  }
  LAUNCH_CHECK(c);
  return algo;
}

}  // namespace sdlc

using namespace sdlc;

extern "C" {

// In-place allreduce of `numel` elements at byte `offset` of symmetric buffer `buf`.
// algo: 0 auto, 1 one-shot, 2 two-shot P2P, 3 two-shot NVLS, 4 LL (flag-in-data, <= 64 KB).  Returns the algorithm used or -1.
int sdlc_allreduce(void* h, int channel, int buf, long long offset, long long numel, int dtype, int op,
                   float pre, float post, int algo, unsigned long long stream) {
  Comm* c = static_cast<Comm*>(h);
  cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : c->streams[channel];
  const CollCtx ctx = c->make_ctx(channel, buf, (size_t)offset);
  if (op == kAvg) {
    op = kSum;
    post /= (float)c->world;
  }
  switch (dtype) {
    case kF32: return launch_allreduce_t<float>(c, ctx, numel, op, pre, post, algo, s);
    case kBF16: return launch_allreduce_t<__nv_bfloat16>(c, ctx, numel, op, pre, post, algo, s);
    case kF16: return launch_allreduce_t<__half>(c, ctx, numel, op, pre, post, algo, s);
    case kF64: return launch_allreduce_t<double>(c, ctx, numel, op, pre, post, algo, s);
    case kI32: return launch_allreduce_t<int32_t>(c, ctx, numel, op, pre, post, algo, s);
    case kI64: return launch_allreduce_t<int64_t>(c, ctx, numel, op, pre, post, algo, s);
    case kU8: case kBool: return launch_allreduce_t<uint8_t>(c, ctx, numel, op, pre, post, algo, s);
    case kI8: return launch_allreduce_t<int8_t>(c, ctx, numel, op, pre, post, algo, s);
    default: c->last_error = "unsupported dtype for CUDA allreduce"; return -1;
  }
}

// bytes of symmetric wire buffer needed for `numel` elements in wire format `wire`
long long sdlc_wire_bytes(long long numel, int wire) {
  const long long units = ceil_div_ll(numel, kUnit);
  return wire == kWireFP8 ? units * kUnit + units * 4 : units * kUnit * 2;
}

// Allreduce of a local storage tensor through a (smaller) symmetric wire buffer.
int sdlc_allreduce_wire(void* h, int channel, void* storage, long long numel, int dtype, int wire, int wire_buf,
                        long long wire_off, float pre, float post, unsigned long long stream) {
  Comm* c = static_cast<Comm*>(h);
  cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : c->streams[channel];
  const CollCtx ctx = c->make_ctx(channel, wire_buf, (size_t)wire_off);
  if ((numel & 7) != 0) {
    c->last_error = "wire-format allreduce needs a multiple of 8 elements";
    return -1;
  }
  const long long units = ceil_div_ll(numel, kUnit);
  // phases 0 and 2 are local HBM passes over the whole tensor: use up to kMaxBlocks CTAs, not just comm_ctas
  long long want = ceil_div_ll(units, (kThreads / 16) * 4);
  if (want < 1) want = 1;
  const int blocks = (int)(want > kMaxBlocks ? kMaxBlocks : want);
  if (dtype == kF32) {
    float* p = static_cast<float*>(storage);
    if (wire == kWireBF16) allreduce_wire_kernel<float, kWireBF16><<<blocks, kThreads, 0, s>>>(ctx, p, numel, pre, post);
    else if (wire == kWireF16) allreduce_wire_kernel<float, kWireF16><<<blocks, kThreads, 0, s>>>(ctx, p, numel, pre, post);
    else if (wire == kWireFP8) allreduce_wire_kernel<float, kWireFP8><<<blocks, kThreads, 0, s>>>(ctx, p, numel, pre, post);
    else { c->last_error = "bad wire format"; return -1; }
  } else if (dtype == kBF16) {
    __nv_bfloat16* p = static_cast<__nv_bfloat16*>(storage);
    if (wire == kWireFP8) allreduce_wire_kernel<__nv_bfloat16, kWireFP8><<<blocks, kThreads, 0, s>>>(ctx, p, numel, pre, post);
    else { c->last_error = "bf16 storage supports the fp8 wire format only"; return -1; }
  } else {
    c->last_error = "wire-format allreduce supports fp32 and bf16 storage";
    return -1;
  }
  LAUNCH_CHECK(c);
  return 0;
}

int sdlc_barrier(void* h, int channel, unsigned long long stream) {
  Comm* c = static_cast<Comm*>(h);
  cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : c->streams[channel];
  const CollCtx ctx = c->make_ctx(channel, c->pad_buf, 0);
  barrier_kernel<<<1, kThreads, 0, s>>>(ctx);
  LAUNCH_CHECK(c);
  return 0;
}

// allgather / broadcast / alltoall: pull byte ranges out of the peers' staging buffers
// `work_bytes` sizes the grid and MUST be identical on every rank (the in-kernel barrier pairs CTA b of
// every rank): callers pass a rank-independent bound, never their own byte count.
int sdlc_pull(void* h, int channel, int buf, void* out, const long long* src_off, const long long* nbytes,
              const long long* dst_off, long long work_bytes, unsigned long long stream) {
  Comm* c = static_cast<Comm*>(h);
  cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : c->streams[channel];
  const CollCtx ctx = c->make_ctx(channel, buf, 0);
  PullTable t;
  for (int p = 0; p < kMaxPeers; ++p) {
    t.src_off[p] = p < c->world ? src_off[p] : 0;
    t.nbytes[p] = p < c->world ? nbytes[p] : 0;
    t.dst_off[p] = p < c->world ? dst_off[p] : 0;
  }
  const int blocks = pick_blocks(c, ceil_div_ll(work_bytes, 16), kThreads * 4);
  pull_kernel<<<blocks, kThreads, 0, s>>>(ctx, static_cast<char*>(out), t);
  LAUNCH_CHECK(c);
  return 0;
}

int sdlc_bcast_multimem(void* h, int channel, int buf, const void* src, void* out, long long nbytes, int root,
                        unsigned long long stream) {
  Comm* c = static_cast<Comm*>(h);
  cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : c->streams[channel];
  const CollCtx ctx = c->make_ctx(channel, buf, 0);
  if (ctx.mc == nullptr || (nbytes & 15) != 0) {
    c->last_error = "multimem broadcast needs a multicast-bound staging buffer and 16-byte multiples";
    return -1;
  }
  const int blocks = pick_blocks(c, nbytes >> 4, kThreads * 4);
  bcast_multimem_kernel<<<blocks, kThreads, 0, s>>>(ctx, static_cast<const char*>(src), static_cast<char*>(out), nbytes, root);
  LAUNCH_CHECK(c);
  return 0;
}

// `work_elems`: rank-independent grid-sizing bound (see sdlc_pull)
int sdlc_reducescatter(void* h, int channel, int buf, void* out, long long lo, long long cnt, int dtype, int op,
                       float pre, float post, long long work_elems, unsigned long long stream) {
  Comm* c = static_cast<Comm*>(h);
  cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : c->streams[channel];
  const CollCtx ctx = c->make_ctx(channel, buf, 0);
  if (op == kAvg) {
    op = kSum;
    post /= (float)c->world;
  }
  const int blocks = pick_blocks(c, work_elems, kThreads * 4);
#define RS(T) reducescatter_kernel<T><<<blocks, kThreads, 0, s>>>(ctx, static_cast<T*>(out), lo, cnt, op, pre, post)
  switch (dtype) {
    case kF32: RS(float); break;
    case kBF16: RS(__nv_bfloat16); break;
    case kF16: RS(__half); break;
    case kF64: RS(double); break;
    case kI32: RS(int32_t); break;
    case kI64: RS(int64_t); break;
    case kU8: case kBool: RS(uint8_t); break;
    case kI8: RS(int8_t); break;
    default: c->last_error = "unsupported dtype for CUDA reducescatter"; return -1;
  }
#undef RS
  LAUNCH_CHECK(c);
  return 0;
}

// Batched copy between n user tensors and the staging buffer of `buf` (to_staging=1: pack).
int sdlc_pack(void* h, int buf, const unsigned long long* ptrs, const long long* offs, const long long* sizes, int n,
              int to_staging, unsigned long long stream) {
  Comm* c = static_cast<Comm*>(h);
  if (n <= 0) return 0;
  if (n > 1024) {   // the pointer table of one launch holds 1024 entries: larger batches go out as several launches
    for (int i = 0; i < n; i += 1024) {
      const int rc = sdlc_pack(h, buf, ptrs + i, offs + i, sizes + i, n - i < 1024 ? n - i : 1024, to_staging, stream);
      if (rc != 0) return rc;
    }
    return 0;
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int slot = c->table_next;
  c->table_next = (c->table_next + 1) % 16;
  cudaEventSynchronize(c->table_events[slot]);  // slot reuse: previous user has consumed the table
  char* hb = static_cast<char*>(c->table_host) + (size_t)slot * 3 * 1024 * 8;
  char* db = static_cast<char*>(c->table_dev) + (size_t)slot * 3 * 1024 * 8;
  memcpy(hb, ptrs, (size_t)n * 8);
  memcpy(hb + 1024 * 8, offs, (size_t)n * 8);
  memcpy(hb + 2 * 1024 * 8, sizes, (size_t)n * 8);
  cudaMemcpyAsync(db, hb, 3 * 1024 * 8, cudaMemcpyHostToDevice, s);
  long long maxb = 0;
  for (int i = 0; i < n; ++i) maxb = sizes[i] > maxb ? sizes[i] : maxb;
  int bx = (int)ceil_div_ll(maxb, 256 * 16 * 4);
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  dim3 grid(bx, n);
  pack_kernel<<<grid, 256, 0, s>>>(static_cast<char*>(c->bufs[buf].ptrs[c->rank]),
                                   reinterpret_cast<const unsigned long long*>(db),
                                   reinterpret_cast<const long long*>(db + 1024 * 8),
                                   reinterpret_cast<const long long*>(db + 2 * 1024 * 8), to_staging);
  cudaEventRecord(c->table_events[slot], s);
  LAUNCH_CHECK(c);
  return 0;
}

int sdlc_scale(void* h, void* p, long long n, int dtype, float s, unsigned long long stream) {
  Comm* c = static_cast<Comm*>(h);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int blocks = (int)ceil_div_ll(n, 256 * 8);
  if (blocks < 1) blocks = 1;
  if (blocks > 1184) blocks = 1184;
  switch (dtype) {
    case kF32: scale_kernel<float><<<blocks, 256, 0, st>>>(static_cast<float*>(p), n, s); break;
    case kBF16: scale_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<__nv_bfloat16*>(p), n, s); break;
    case kF16: scale_kernel<__half><<<blocks, 256, 0, st>>>(static_cast<__half*>(p), n, s); break;
    case kF64: scale_kernel<double><<<blocks, 256, 0, st>>>(static_cast<double*>(p), n, s); break;
    default: return 0;
  }
  LAUNCH_CHECK(c);
  return 0;
}

}  // extern "C"
