// Fused "reduce-scatter + optimizer + all-gather" step (BASELINE.json north
// star: "FUSES the reduction with the gradient cast/scale ... so no separate NCCL call or
// elementwise kernel runs on that path").
//
// Gradients AND parameters of a bucket live in symmetric buffers.  Rank r owns slice r:
//   g      = (sum over peers of grad[slice r]) * scale        (P2P loads or multimem.ld_reduce)
//   p_new  = optimizer(p_old, g, state[slice r])               (SGD-momentum / Adam / AdamW, fp32 math,
//                                                               optional fp32 master weights)
//   param[slice r] on EVERY rank = p_new                       (P2P stores or multimem.st)
// One launch replaces allreduce + foreach-optimizer kernels, and the optimizer state (and master
// weights) are sharded 1/N per GPU.
#include <cstring>

#include "comm.h"

namespace sdlc {

struct OptimArgs {
  int kind;        // 0 SGD, 1 Adam (L2 in grad), 2 AdamW (decoupled)
  float lr, weight_decay;
  float momentum, dampening;
  int nesterov, first_step;
  float beta1, beta2, eps, bias_c1, bias_c2;  // bias_c = 1 - beta^t
};

// Hyper-parameters live in DEVICE memory (16 floats per param group) so that a captured CUDA graph of the
// training step keeps working when the learning rate or the step count changes:
//   [0] kind  [1] lr  [2] weight_decay  [3] momentum  [4] dampening  [5] nesterov  [6] beta1  [7] beta2
//   [8] eps   [9] number of optimizer steps already taken (advanced on the device by step_advance_kernel)
__device__ __forceinline__ OptimArgs load_hyper(const float* __restrict__ h) {
  OptimArgs a;
  a.kind = (int)h[0]; a.lr = h[1]; a.weight_decay = h[2]; a.momentum = h[3]; a.dampening = h[4];
  a.nesterov = (int)h[5]; a.beta1 = h[6]; a.beta2 = h[7]; a.eps = h[8];
  const float t = h[9] + 1.f;
  a.first_step = h[9] == 0.f;
  a.bias_c1 = 1.f - powf(a.beta1, t);
  a.bias_c2 = 1.f - powf(a.beta2, t);
  return a;
}

__global__ void step_advance_kernel(float* hyper) { hyper[9] += 1.f; }

template <typename T> __device__ __forceinline__ uint4 mc_ld_reduce_t(const void* p);
template <> __device__ __forceinline__ uint4 mc_ld_reduce_t<float>(const void* p) { return mc_ld_reduce_f32(p); }
template <> __device__ __forceinline__ uint4 mc_ld_reduce_t<__nv_bfloat16>(const void* p) { return mc_ld_reduce_bf16(p); }
template <> __device__ __forceinline__ uint4 mc_ld_reduce_t<__half>(const void* p) { return mc_ld_reduce_f16(p); }

// one vector of parameters: gr = summed gradient, pw = fp32 parameter (master), m1 / m2 = optimizer moments (all in
// registers, loaded by the caller so that every load of an iteration is in flight before the first use)
template <typename T>
__device__ __forceinline__ void optimizer_update(const float* gr, float* pw, float* m1, float* m2, float scale,
                                                 const OptimArgs& a, float inv_sqrt_bc2, float step_size) {
  using Tr = VecTraits<T>;
#pragma unroll
  for (int i = 0; i < Tr::N; ++i) {
    float gi = gr[i] * scale, p = pw[i];
    if (a.kind == 0) {
      gi = fmaf(a.weight_decay, p, gi);
      if (a.momentum != 0.f) {
        const float b = a.first_step ? gi : fmaf(a.momentum, m1[i], (1.f - a.dampening) * gi);
        m1[i] = b;
        gi = a.nesterov ? fmaf(a.momentum, b, gi) : b;
      }
      p = fmaf(-a.lr, gi, p);
    } else {
      if (a.kind == 1) gi = fmaf(a.weight_decay, p, gi);
      else p *= (1.f - a.lr * a.weight_decay);
      const float m = fmaf(a.beta1, m1[i], (1.f - a.beta1) * gi);
      const float v = fmaf(a.beta2, m2[i], (1.f - a.beta2) * gi * gi);
      m1[i] = m;
      m2[i] = v;
      const float denom = fmaf(sqrtf(v), inv_sqrt_bc2, a.eps);
      p = fmaf(-step_size, __fdividef(m, denom), p);
    }
    pw[i] = p;
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
fused_step_kernel(CollCtx g, PeerPtrs params, void* params_mc, float* __restrict__ s1, float* __restrict__ s2,
                  float* __restrict__ master, long long nvec, float scale, const float* __restrict__ hyper) {
  using Tr = VecTraits<T>;
  const OptimArgs a = load_hyper(hyper);
  EpochScope ep(g);
  coll_barrier(g, ep.base + 1);
  const int world = g.world;
  const long long per = (nvec + world - 1) / world;
  const long long lo = min(nvec, (long long)g.rank * per), hi = min(nvec, lo + per);
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool mc_in = g.mc != nullptr, mc_out = params_mc != nullptr;
  const bool use_m1 = a.kind != 0 || a.momentum != 0.f, use_m2 = a.kind != 0;
  const float inv_sqrt_bc2 = rsqrtf(a.bias_c2), step_size = a.lr / a.bias_c1;
  constexpr int U = 2;   // vectors per thread per iteration: all gradient loads are issued before any math
  for (long long i0 = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += U * stride) {
    float gr[U][Tr::N];
    bool valid[U];
    if (mc_in) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        valid[u] = i0 + u * stride < hi;
        if (valid[u]) v[u] = mc_ld_reduce_t<T>(static_cast<const uint4*>(g.mc) + i0 + u * stride);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (valid[u]) vec_unpack<T>(v[u], gr[u]);
    } else {
      // P2P fallback (no multicast): one peer at a time, U vectors in flight per peer - the loads of all peers at
      // once would need U x 8 x 4 registers and spill
#pragma unroll
      for (int u = 0; u < U; ++u) {
        valid[u] = i0 + u * stride < hi;
#pragma unroll
        for (int i = 0; i < Tr::N; ++i) gr[u][i] = 0.f;
      }
      for (int p = 0; p < world; ++p) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (valid[u]) v[u] = ld_peer_v4(static_cast<const uint4*>(g.data.p[p]) + i0 + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (!valid[u]) continue;
          float t[Tr::N];
          vec_unpack<T>(v[u], t);
#pragma unroll
          for (int i = 0; i < Tr::N; ++i) gr[u][i] += t[i];
        }
      }
    }
    // optimizer state of all U vectors: issued back to back (with the gradient loads above still in flight)
    float pw[U][Tr::N], m1[U][Tr::N], m2[U][Tr::N];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!valid[u]) continue;
      const long long idx = i0 + u * stride;
      const long long k0 = (idx - lo) * Tr::N;
      if (master != nullptr) {
#pragma unroll
        for (int i = 0; i < Tr::N; i += 4) *reinterpret_cast<float4*>(pw[u] + i) = __ldcs(reinterpret_cast<const float4*>(master + k0 + i));
      } else {
        vec_unpack<T>(static_cast<const uint4*>(params.p[g.rank])[idx], pw[u]);
      }
      if (use_m1) {
#pragma unroll
        for (int i = 0; i < Tr::N; i += 4) *reinterpret_cast<float4*>(m1[u] + i) = __ldcs(reinterpret_cast<const float4*>(s1 + k0 + i));
      }
      if (use_m2) {
#pragma unroll
        for (int i = 0; i < Tr::N; i += 4) *reinterpret_cast<float4*>(m2[u] + i) = __ldcs(reinterpret_cast<const float4*>(s2 + k0 + i));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!valid[u]) continue;
      const long long idx = i0 + u * stride;
      const long long k0 = (idx - lo) * Tr::N;
      optimizer_update<T>(gr[u], pw[u], m1[u], m2[u], scale, a, inv_sqrt_bc2, step_size);
      if (use_m1) {
#pragma unroll
        for (int i = 0; i < Tr::N; i += 4) __stcs(reinterpret_cast<float4*>(s1 + k0 + i), *reinterpret_cast<const float4*>(m1[u] + i));
      }
      if (use_m2) {
#pragma unroll
        for (int i = 0; i < Tr::N; i += 4) __stcs(reinterpret_cast<float4*>(s2 + k0 + i), *reinterpret_cast<const float4*>(m2[u] + i));
      }
      if (master != nullptr) {
#pragma unroll
        for (int i = 0; i < Tr::N; i += 4) __stcs(reinterpret_cast<float4*>(master + k0 + i), *reinterpret_cast<const float4*>(pw[u] + i));
      }
      const uint4 out = vec_pack<T>(pw[u]);
      if (mc_out) {
        mc_st_v4(static_cast<uint4*>(params_mc) + idx, out);
      } else {
#pragma unroll
        for (int p = 0; p < kMaxPeers; ++p)
          if (p < world) st_peer_v4(static_cast<uint4*>(params.p[p]) + idx, out);
      }
    }
  }
  coll_barrier(g, ep.base + 2);
  ep.commit(g, 2);
}

}  // namespace sdlc

using namespace sdlc;

extern "C" {

// hyper: DEVICE pointer to the 16-float hyper-parameter block of the bucket's param group (see load_hyper)
int sdlc_fused_step(void* h, int channel, int grad_buf, int param_buf, long long numel, int dtype, float* s1, float* s2,
                    float* master, float scale, const float* hyper, unsigned long long stream) {
  Comm* c = static_cast<Comm*>(h);
  cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : c->streams[channel];
  CollCtx g = c->make_ctx(channel, grad_buf, 0);
  const SymmBuffer& pb = c->bufs[param_buf];
  PeerPtrs pp;
  memset(&pp, 0, sizeof(pp));
  for (int r = 0; r < c->world; ++r) pp.p[r] = pb.ptrs[r];
  // NVLS (multimem.ld_reduce of the gradient shard, multimem.st of the updated parameters) from 3 ranks up; at 2 ranks
  // the peer-to-peer branch moves 1.0 S instead of 1.5 S per direction over the links
  const bool nvls = c->world >= c->nvls_min_world;
  void* pmc = (pb.mc && c->use_mc && nvls) ? pb.mc : nullptr;
  if (!nvls) g.mc = nullptr;
  const size_t esz = dtype == kF32 ? 4 : 2;
  const long long nvec = (numel * (long long)esz + 15) / 16;
  long long blocks = ((nvec + c->world - 1) / c->world + kThreads * 2 - 1) / (kThreads * 2);
  if (blocks < 1) blocks = 1;
  // a single rank owns the whole bucket and only streams local HBM: give it more CTAs
  long long cap = c->world == 1 ? 148 : (long long)c->max_blocks * (c->world < 8 ? 8 / c->world : 1);
  if (cap > kMaxBlocks && c->world > 1) cap = kMaxBlocks;
  if (blocks > cap) blocks = cap;
  if (dtype == kF32) {
    fused_step_kernel<float><<<(int)blocks, kThreads, 0, s>>>(g, pp, pmc, s1, s2, master, nvec, scale, hyper);
  } else if (dtype == kBF16) {
    fused_step_kernel<__nv_bfloat16><<<(int)blocks, kThreads, 0, s>>>(g, pp, pmc, s1, s2, master, nvec, scale, hyper);
  } else if (dtype == kF16) {
    fused_step_kernel<__half><<<(int)blocks, kThreads, 0, s>>>(g, pp, pmc, s1, s2, master, nvec, scale, hyper);
  } else {
    c->last_error = "fused optimizer step supports fp32 / bf16 / fp16 parameters";
    return -1;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    c->last_error = std::string("fused_step launch failed: ") + cudaGetErrorString(e);
    return -1;
  }
  c->launches++;
  return 0;
}

// advance the device-side step counter of one param group (launched once per optimizer step, after its buckets)
int sdlc_step_advance(void* h, int channel, float* hyper, unsigned long long stream) {
  Comm* c = static_cast<Comm*>(h);
  cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : c->streams[channel];
  step_advance_kernel<<<1, 1, 0, s>>>(hyper);
  c->launches++;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// elements of optimizer state this rank must allocate for a bucket of `numel` elements
long long sdlc_fused_state_elems(void* h, long long numel, int dtype) {
  Comm* c = static_cast<Comm*>(h);
  const size_t esz = dtype == kF32 ? 4 : 2;
  const long long n_per_vec = 16 / (long long)esz;
  const long long nvec = (numel * (long long)esz + 15) / 16;
  const long long per = (nvec + c->world - 1) / c->world;
  return per * n_per_vec;
}

// first element (in units of elements) of this rank's slice - for checkpointing the sharded state
long long sdlc_fused_slice_start(void* h, long long numel, int dtype) {
  Comm* c = static_cast<Comm*>(h);
  const size_t esz = dtype == kF32 ? 4 : 2;
  const long long n_per_vec = 16 / (long long)esz;
  const long long nvec = (numel * (long long)esz + 15) / 16;
  const long long per = (nvec + c->world - 1) / c->world;
  long long lo = (long long)c->rank * per;
  if (lo > nvec) lo = nvec;
  return lo * n_per_vec;
}

}  // extern "C"
