// Device-side building blocks of the sm_100a peer-memory collectives.
//
// Net-new (the reference has no kernels: SURVEY.md §2f "0 .cu files, 0 NCCL call sites").
// Everything here addresses *peer* memory directly: `CollCtx::data.p[r]` is rank r's symmetric
// buffer mapped into this process (CUDA VMM / cudaIpc), `mc` is the NVSwitch multicast alias of
// the same buffer (multimem.* PTX), and `pads` are flag words in peer memory used for in-kernel
// barriers (st.release.sys / ld.acquire.sys), so a whole collective is ONE kernel launch.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sdlc {

constexpr int kMaxPeers = 8;
constexpr int kMaxBlocks = 128;     // CTAs of one collective launch
constexpr int kChannels = 4;        // independent flag sets: 0 engine, 1 optimizer, 2 gbdt, 3 spare
constexpr int kPadWordsPerChannel = kMaxBlocks * kMaxPeers;
constexpr int kThreads = 512;
constexpr int kLLVecs = 4096;       // 16-byte vectors per message on the flag-in-data (LL) path: 64 KB
// LL receive area per channel: [parity 2][source rank kMaxPeers][kLLVecs] slots of 32 bytes (4 x {data32, flag32})
constexpr size_t kLLSlotBytes = 32;
constexpr size_t kLLChannelBytes = 2ull * kMaxPeers * kLLVecs * kLLSlotBytes;

// dtype / op codes: keep in sync with csrc/core/common.h
enum DType : uint32_t { kU8 = 0, kI8 = 1, kI32 = 2, kI64 = 3, kF16 = 4, kBF16 = 5, kF32 = 6, kF64 = 7, kBool = 8, kF8 = 9 };
enum Op : uint32_t { kSum = 0, kAvg = 1, kMin = 2, kMax = 3, kProd = 4 };
enum Wire : int { kWireNone = 0, kWireBF16 = 1, kWireF16 = 2, kWireFP8 = 3 };

struct PeerPtrs { void* p[kMaxPeers]; };

struct CollCtx {
  PeerPtrs data;            // symmetric payload buffer, one mapping per rank
  void* mc;                 // multicast (NVLS) alias of `data`, or nullptr
  uint32_t* pads[kMaxPeers];// signal pad of every rank for this channel
  uint32_t* epoch;          // local, per-block barrier epoch counters for this channel
  uint32_t* err;            // host-mapped error word (set on spin timeout)
  int rank, world;
  unsigned long long timeout_ns;
};

// ---- memory-model primitives --------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// 16-byte peer load that never allocates in L1 (peer lines are only coherent at the home L2)
__device__ __forceinline__ uint4 ld_peer_v4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer_v4(void* p, const uint4& v) {
  asm volatile("st.relaxed.sys.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint2 ld_peer_v2(const void* p) {
  uint2 v;
  asm volatile("ld.relaxed.sys.global.L1::no_allocate.v2.u32 {%0,%1}, [%2];"
               : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer_v2(void* p, const uint2& v) {
  asm volatile("st.relaxed.sys.global.L1::no_allocate.v2.u32 [%0], {%1,%2};"
               ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
// two 64-bit words; every 64-bit element is single-copy atomic, so a {data, flag} pair can never be seen half-written
__device__ __forceinline__ void st_peer_v2_u64(void* p, unsigned long long a, unsigned long long b) {
  asm volatile("st.relaxed.sys.global.L1::no_allocate.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_peer_v2_u64(const void* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.relaxed.sys.global.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ uint32_t ld_peer_u32(const void* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer_u32(void* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// NVLS: the switch reduces the N copies behind a multicast address / replicates a store
__device__ __forceinline__ uint4 mc_ld_reduce_f32(const void* p) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 mc_ld_reduce_bf16(const void* p) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 mc_ld_reduce_f16(const void* p) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st_v4(void* p, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- in-kernel barrier over peer flag words -------------------------------
// Flag slot for (block b, source rank s) inside a rank's pad: b * kMaxPeers + s.  Epochs only
// grow, so flags never need clearing.  Thread t < world signals peer t and waits for peer t.
__device__ __forceinline__ void coll_barrier(const CollCtx& c, uint32_t epoch) {
  __syncthreads();
  if (threadIdx.x < (unsigned)c.world) {
    const int peer = threadIdx.x;
    st_release_sys(c.pads[peer] + blockIdx.x * kMaxPeers + c.rank, epoch);
    const uint32_t* mine = c.pads[c.rank] + blockIdx.x * kMaxPeers + peer;
    const unsigned long long t0 = globaltimer_ns();
    unsigned spins = 0;
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
      if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > c.timeout_ns) {
        *reinterpret_cast<volatile uint32_t*>(c.err) = 0x80000000u | ((uint32_t)peer << 8) | (uint32_t)c.rank;
        break;
      }
    }
  }
  __syncthreads();
}

// Per-kernel epoch bookkeeping: every block reads its counter once, uses base+1, base+2, ...
// for its barriers and stores base+used at the end.
struct EpochScope {
  uint32_t base;
  __device__ __forceinline__ explicit EpochScope(const CollCtx& c) { base = c.epoch[blockIdx.x]; }
  __device__ __forceinline__ void commit(const CollCtx& c, uint32_t used) {
    __syncthreads();
    if (threadIdx.x == 0) c.epoch[blockIdx.x] = base + used;
  }
};

// ---- 16-byte vector <-> accumulator conversion ---------------------------------------------------
template <typename T> struct VecTraits;
template <> struct VecTraits<float> { static constexpr int N = 4; using Acc = float; };
template <> struct VecTraits<__nv_bfloat16> { static constexpr int N = 8; using Acc = float; };
template <> struct VecTraits<__half> { static constexpr int N = 8; using Acc = float; };
template <> struct VecTraits<double> { static constexpr int N = 2; using Acc = double; };
template <> struct VecTraits<int32_t> { static constexpr int N = 4; using Acc = int32_t; };
template <> struct VecTraits<int64_t> { static constexpr int N = 2; using Acc = int64_t; };
template <> struct VecTraits<uint8_t> { static constexpr int N = 16; using Acc = int32_t; };
template <> struct VecTraits<int8_t> { static constexpr int N = 16; using Acc = int32_t; };

template <typename T>
__device__ __forceinline__ void vec_unpack(const uint4& v, typename VecTraits<T>::Acc* a);
template <typename T>
__device__ __forceinline__ uint4 vec_pack(const typename VecTraits<T>::Acc* a);

template <> __device__ __forceinline__ void vec_unpack<float>(const uint4& v, float* a) {
  a[0] = __uint_as_float(v.x); a[1] = __uint_as_float(v.y); a[2] = __uint_as_float(v.z); a[3] = __uint_as_float(v.w);
}
template <> __device__ __forceinline__ uint4 vec_pack<float>(const float* a) {
  return make_uint4(__float_as_uint(a[0]), __float_as_uint(a[1]), __float_as_uint(a[2]), __float_as_uint(a[3]));
}
__device__ __forceinline__ void bf16x2_to_f(uint32_t u, float& lo, float& hi) {
  lo = __uint_as_float(u << 16);
  hi = __uint_as_float(u & 0xffff0000u);
}
__device__ __forceinline__ uint32_t f_to_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ void vec_unpack<__nv_bfloat16>(const uint4& v, float* a) {
  bf16x2_to_f(v.x, a[0], a[1]); bf16x2_to_f(v.y, a[2], a[3]); bf16x2_to_f(v.z, a[4], a[5]); bf16x2_to_f(v.w, a[6], a[7]);
}
template <> __device__ __forceinline__ uint4 vec_pack<__nv_bfloat16>(const float* a) {
  return make_uint4(f_to_bf16x2(a[0], a[1]), f_to_bf16x2(a[2], a[3]), f_to_bf16x2(a[4], a[5]), f_to_bf16x2(a[6], a[7]));
}
__device__ __forceinline__ void f16x2_to_f(uint32_t u, float& lo, float& hi) {
  __half2 h = *reinterpret_cast<__half2*>(&u);
  float2 f = __half22float2(h);
  lo = f.x; hi = f.y;
}
__device__ __forceinline__ uint32_t f_to_f16x2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ void vec_unpack<__half>(const uint4& v, float* a) {
  f16x2_to_f(v.x, a[0], a[1]); f16x2_to_f(v.y, a[2], a[3]); f16x2_to_f(v.z, a[4], a[5]); f16x2_to_f(v.w, a[6], a[7]);
}
template <> __device__ __forceinline__ uint4 vec_pack<__half>(const float* a) {
  return make_uint4(f_to_f16x2(a[0], a[1]), f_to_f16x2(a[2], a[3]), f_to_f16x2(a[4], a[5]), f_to_f16x2(a[6], a[7]));
}
template <> __device__ __forceinline__ void vec_unpack<double>(const uint4& v, double* a) {
  a[0] = __hiloint2double((int)v.y, (int)v.x); a[1] = __hiloint2double((int)v.w, (int)v.z);
}
template <> __device__ __forceinline__ uint4 vec_pack<double>(const double* a) {
  return make_uint4((uint32_t)__double2loint(a[0]), (uint32_t)__double2hiint(a[0]),
                    (uint32_t)__double2loint(a[1]), (uint32_t)__double2hiint(a[1]));
}
template <> __device__ __forceinline__ void vec_unpack<int32_t>(const uint4& v, int32_t* a) {
  a[0] = (int32_t)v.x; a[1] = (int32_t)v.y; a[2] = (int32_t)v.z; a[3] = (int32_t)v.w;
}
template <> __device__ __forceinline__ uint4 vec_pack<int32_t>(const int32_t* a) {
  return make_uint4((uint32_t)a[0], (uint32_t)a[1], (uint32_t)a[2], (uint32_t)a[3]);
}
template <> __device__ __forceinline__ void vec_unpack<int64_t>(const uint4& v, int64_t* a) {
  a[0] = (int64_t)(((uint64_t)v.y << 32) | v.x); a[1] = (int64_t)(((uint64_t)v.w << 32) | v.z);
}
template <> __device__ __forceinline__ uint4 vec_pack<int64_t>(const int64_t* a) {
  return make_uint4((uint32_t)a[0], (uint32_t)((uint64_t)a[0] >> 32), (uint32_t)a[1], (uint32_t)((uint64_t)a[1] >> 32));
}
template <> __device__ __forceinline__ void vec_unpack<uint8_t>(const uint4& v, int32_t* a) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (int32_t)((w[i >> 2] >> ((i & 3) * 8)) & 0xff);
}
template <> __device__ __forceinline__ uint4 vec_pack<uint8_t>(const int32_t* a) {
  uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i >> 2] |= ((uint32_t)a[i] & 0xff) << ((i & 3) * 8);
  return make_uint4(w[0], w[1], w[2], w[3]);
}
template <> __device__ __forceinline__ void vec_unpack<int8_t>(const uint4& v, int32_t* a) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (int32_t)(int8_t)((w[i >> 2] >> ((i & 3) * 8)) & 0xff);
}
template <> __device__ __forceinline__ uint4 vec_pack<int8_t>(const int32_t* a) {
  uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i >> 2] |= ((uint32_t)a[i] & 0xff) << ((i & 3) * 8);
  return make_uint4(w[0], w[1], w[2], w[3]);
}

template <typename A>
__device__ __forceinline__ A reduce_op(A x, A y, int op) {
  switch (op) {
    case kMin: return x < y ? x : y;
    case kMax: return x > y ? x : y;
    case kProd: return x * y;
    default: return x + y;
  }
}

template <typename A> __device__ __forceinline__ A scale_val(A x, float s) { return x; }
template <> __device__ __forceinline__ float scale_val<float>(float x, float s) { return x * s; }
template <> __device__ __forceinline__ double scale_val<double>(double x, float s) { return x * (double)s; }

}  // namespace sdlc
