// Fused softmax cross-entropy over bf16 logits (the 30k-wide MLM head of BERT-large, BASELINE.json config 3).
//
// torch runs  logits.float() -> log_softmax -> nll_loss  forward and the mirror image backward: the [tokens, vocab]
// matrix (250 MB in bf16 at 4096 x 30528) is read or written nine times, four of them in fp32.  Here the forward is ONE
// pass (online max / sum-of-exponentials per row, fp32 math on bf16 data) that keeps only the per-row log-sum-exp, and
// the backward is ONE pass that overwrites nothing and writes d(logits) in bf16:  (softmax - onehot) * scale.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace xent {

constexpr int kThreads = 256;

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// (m, s): running maximum and sum of exp(x - m); merge is associative
__device__ __forceinline__ void merge(float& m, float& s, float m2, float s2) {
  const float mm = fmaxf(m, m2);
  if (mm == -INFINITY) return;   // both sides empty (threads beyond a short row)
  s = s * __expf(m - mm) + s2 * __expf(m2 - mm);
  m = mm;
}

// one block per row; V % 8 == 0; labels < 0 (ignore_index) contribute nothing
__global__ void __launch_bounds__(kThreads)
xent_fwd_kernel(const __nv_bfloat16* __restrict__ logits, const long long* __restrict__ labels, float* __restrict__ lse,
                float* __restrict__ loss, int V, long long ld) {
  const long long row = blockIdx.x;
  const uint4* p = reinterpret_cast<const uint4*>(logits + row * ld);
  const int nvec = V / 8;
  float m = -INFINITY, s = 0.f;
  for (int v = threadIdx.x; v < nvec; v += kThreads) {
    float f[8];
    unpack8(__ldg(p + v), f);
    float vm = f[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) vm = fmaxf(vm, f[i]);
    const float mm = fmaxf(m, vm);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += __expf(f[i] - mm);
    s = s * __expf(m - mm) + acc;
    m = mm;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    merge(m, s, m2, s2);
  }
  __shared__ float sm[kThreads / 32], ss[kThreads / 32];
  if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = m; ss[threadIdx.x >> 5] = s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    m = sm[0]; s = ss[0];
#pragma unroll
    for (int w = 1; w < kThreads / 32; ++w) merge(m, s, sm[w], ss[w]);
    const float l = m + __logf(s);
    lse[row] = l;
    const long long y = labels[row];
    loss[row] = (y >= 0 && y < V) ? l - __bfloat162float(logits[row * ld + y]) : 0.f;
  }
}

// dlogits[row, v] = (exp(logit - lse) - [v == label]) * scale[0] * gscale   (0 for ignored rows)
__global__ void __launch_bounds__(kThreads)
xent_bwd_kernel(const __nv_bfloat16* __restrict__ logits, const long long* __restrict__ labels, const float* __restrict__ lse,
                const float* __restrict__ grad_out, float gscale, __nv_bfloat16* __restrict__ dlogits, int V, long long ld) {
  const long long row = blockIdx.x;
  const uint4* p = reinterpret_cast<const uint4*>(logits + row * ld);
  uint4* q = reinterpret_cast<uint4*>(dlogits + row * ld);
  const long long y = labels[row];
  const bool live = y >= 0 && y < V;
  const float l = lse[row], g = live ? grad_out[0] * gscale : 0.f;
  const int nvec = V / 8;
  for (int v = threadIdx.x; v < nvec; v += kThreads) {
    float f[8];
    unpack8(__ldg(p + v), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float d = __expf(f[i] - l);
      if (v * 8 + i == y) d -= 1.f;
      f[i] = d * g;
    }
    __stcs(q + v, pack8(f));
  }
}

}  // namespace xent

extern "C" {

// logits: bf16 [T, ld] (V valid columns, V % 8 == 0, ld % 8 == 0); labels: int64 [T]; lse, loss: fp32 [T]
int sdlc_xent_fwd(const void* logits, const long long* labels, float* lse, float* loss, int T, int V, long long ld,
                  unsigned long long stream) {
  if (T <= 0) return 0;
  if ((V % 8) != 0 || (ld % 8) != 0) return -2;
  xent::xent_fwd_kernel<<<T, xent::kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(logits), labels, lse, loss, V, ld);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// grad_out: fp32 scalar on the device (upstream gradient of the mean loss); gscale: 1 / number of live rows
int sdlc_xent_bwd(const void* logits, const long long* labels, const float* lse, const float* grad_out, float gscale,
                  void* dlogits, int T, int V, long long ld, unsigned long long stream) {
  if (T <= 0) return 0;
  if ((V % 8) != 0 || (ld % 8) != 0) return -2;
  xent::xent_bwd_kernel<<<T, xent::kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(logits), labels, lse, grad_out, gscale, static_cast<__nv_bfloat16*>(dlogits), V, ld);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // extern "C"
