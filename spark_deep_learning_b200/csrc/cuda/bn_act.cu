// Fused training-mode BatchNorm (+ residual add) (+ ReLU) for NHWC activations, forward and backward.
//
// ResNet-50 (BASELINE.json config 2) spends as much time streaming activations through
// batch-norm / add / relu as in its convolutions.  Between two (library) convolutions we touch
// the activation exactly twice in forward (statistics, then normalise+add+relu -> bf16 + 1-bit
// relu mask) and twice in backward (reductions, then dx [+ dz for the residual branch]).
// All loops are 16-byte vectorised over channels; each thread owns a fixed group of 8 (bf16) or
// 4 (fp32) channels so per-channel parameters live in registers.  Reductions are deterministic:
// per-block partials + a finalize kernel, no float atomics.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

namespace bn {

constexpr int kThreads = 256;

template <typename T> struct V;
template <> struct V<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void load(const __nv_bfloat16* p, float* f) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static void store(__nv_bfloat16* p, const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <> struct V<float> {
  static constexpr int N = 4;
  __device__ static void load(const float* p, float* f) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  __device__ static void store(float* p, const float* f) { *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]); }
};

// Block-level reduction of K per-thread values across the threads that share a channel group.
// Thread t owns channel group t % CV; groups repeat every CV threads inside the block.
template <int K>
__device__ __forceinline__ void block_reduce_groups(float* vals, int CV, float* smem /* kThreads*K */) {
  const int t = threadIdx.x;
#pragma unroll
  for (int k = 0; k < K; ++k) smem[k * kThreads + t] = vals[k];
  __syncthreads();
  for (int s = kThreads / 2; s >= CV; s >>= 1) {
    if (t < s) {
#pragma unroll
      for (int k = 0; k < K; ++k) smem[k * kThreads + t] += smem[k * kThreads + t + s];
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < K; ++k) vals[k] = smem[k * kThreads + t];  // valid for t < CV
}

// ---------------- forward pass 1: per-block partial sum / sum of squares -------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
bn_stats_kernel(const T* __restrict__ x, float* __restrict__ partial, long long nvec, int C) {
  constexpr int N = V<T>::N;
  const int CV = C / N;
  extern __shared__ float smem[];
  const long long T_all = (long long)gridDim.x * kThreads;
  const long long g = (long long)blockIdx.x * kThreads + threadIdx.x;
  float acc[2 * N];
#pragma unroll
  for (int i = 0; i < 2 * N; ++i) acc[i] = 0.f;
  long long v = g;
  for (; v + 3 * T_all < nvec; v += 4 * T_all) {
    float f[4][N];
#pragma unroll
    for (int u = 0; u < 4; ++u) V<T>::load(x + (v + u * T_all) * N, f[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < N; ++i) {
        acc[i] += f[u][i];
        acc[N + i] = fmaf(f[u][i], f[u][i], acc[N + i]);
      }
  }
  for (; v < nvec; v += T_all) {
    float f[N];
    V<T>::load(x + v * N, f);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      acc[i] += f[i];
      acc[N + i] = fmaf(f[i], f[i], acc[N + i]);
    }
  }
  if (CV < kThreads) block_reduce_groups<2 * N>(acc, CV, smem);
  // channel group of this thread: (g % CV); T_all % CV == 0 and kThreads % CV == 0 (or CV % kThreads == 0)
  if (CV >= kThreads) {
    const int cv = (int)(g % CV);
    // several blocks may share a channel group set only when CV > kThreads: block b covers groups
    // [(b*kThreads) % CV, ...+kThreads); partial row = blockIdx.x, other groups stay zero (pre-zeroed)
#pragma unroll
    for (int i = 0; i < N; ++i) {
      partial[(long long)blockIdx.x * 2 * C + cv * N + i] = acc[i];
      partial[(long long)blockIdx.x * 2 * C + C + cv * N + i] = acc[N + i];
    }
  } else if (threadIdx.x < CV) {
    const int cv = threadIdx.x;  // blockIdx.x*kThreads % CV == 0
#pragma unroll
    for (int i = 0; i < N; ++i) {
      partial[(long long)blockIdx.x * 2 * C + cv * N + i] = acc[i];
      partial[(long long)blockIdx.x * 2 * C + C + cv * N + i] = acc[N + i];
    }
  }
}

// ---------------- finalize: mean / rstd / fused scale+shift, running statistics ----------------------
// Block = 8 channels x 128 row groups (1024 threads), grid = C / 8.  The loop over the per-block partial rows is pure
// L2 latency, so it is made as short as possible: with 592 partial rows a thread issues at most 5 independent 32-byte
// row reads in one go.  The 128 row groups are then folded in a fixed order (warp shuffles over the 4 groups of a warp,
// shared memory across the 32 warps, shuffles again in warp 0): deterministic, and threads (x, y = 0) own the result.
constexpr int kFinX = 8, kFinY = 128;

__device__ __forceinline__ void sum_partials(const float* __restrict__ partial, int nblocks, int C, int c, float& s, float& q) {
  __shared__ float sh[2][32][kFinX];
  float a = 0.f, b = 0.f;
  if (c < C) {
    int r = threadIdx.y;
    for (; r + 3 * kFinY < nblocks; r += 4 * kFinY) {
      float t0[4], t1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        t0[u] = partial[(long long)(r + u * kFinY) * 2 * C + c];
        t1[u] = partial[(long long)(r + u * kFinY) * 2 * C + C + c];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { a += t0[u]; b += t1[u]; }
    }
    float t0[3] = {0.f, 0.f, 0.f}, t1[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 3; ++u)
      if (r + u * kFinY < nblocks) {
        t0[u] = partial[(long long)(r + u * kFinY) * 2 * C + c];
        t1[u] = partial[(long long)(r + u * kFinY) * 2 * C + C + c];
      }
#pragma unroll
    for (int u = 0; u < 3; ++u) { a += t0[u]; b += t1[u]; }
  }
  // lane = (y & 3) * 8 + x: fold the 4 row groups of this warp
  a += __shfl_xor_sync(0xffffffffu, a, 8);  b += __shfl_xor_sync(0xffffffffu, b, 8);
  a += __shfl_xor_sync(0xffffffffu, a, 16); b += __shfl_xor_sync(0xffffffffu, b, 16);
  const int tid = threadIdx.y * kFinX + threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (lane < kFinX) { sh[0][warp][lane] = a; sh[1][warp][lane] = b; }
  __syncthreads();
  s = 0.f; q = 0.f;
  if (warp == 0) {
    // lane = part * 8 + x: part p adds warps 8p .. 8p+7 in order, then the 4 parts are folded
    const int x = lane & 7, part = lane >> 3;
#pragma unroll
    for (int w = 0; w < 8; ++w) { s += sh[0][part * 8 + w][x]; q += sh[1][part * 8 + w][x]; }
    s += __shfl_xor_sync(0xffffffffu, s, 8);  q += __shfl_xor_sync(0xffffffffu, q, 8);
    s += __shfl_xor_sync(0xffffffffu, s, 16); q += __shfl_xor_sync(0xffffffffu, q, 16);
  }
}

__global__ void __launch_bounds__(kFinX * kFinY)
bn_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, float inv_m, float eps,
                   const float* __restrict__ weight, const float* __restrict__ bias,
                   float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ scale,
                   float* __restrict__ shift, float* running_mean, float* running_var, float momentum,
                   float unbias) {
  const int c = blockIdx.x * kFinX + threadIdx.x;
  float s, q;
  sum_partials(partial, nblocks, C, c, s, q);
  if (c >= C || threadIdx.y != 0) return;
  const float m = s * inv_m;
  const float var = fmaxf(q * inv_m - m * m, 0.f);
  const float r = rsqrtf(var + eps);
  mean[c] = m;
  rstd[c] = r;
  const float a = (weight ? weight[c] : 1.f) * r;
  scale[c] = a;
  shift[c] = (bias ? bias[c] : 0.f) - m * a;
  if (running_mean) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * unbias;
  }
}

// ---------------- forward pass 2: y = relu(scale*x + shift (+ residual)), 1-bit relu mask ------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
bn_apply_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y, uint8_t* __restrict__ mask,
                const float* __restrict__ scale, const float* __restrict__ shift, long long nvec, int C, int relu) {
  constexpr int N = V<T>::N;
  const int CV = C / N;
  const long long T_all = (long long)gridDim.x * kThreads;
  const long long g = (long long)blockIdx.x * kThreads + threadIdx.x;
  const int cv = (int)(g % CV);
  float a[N], b[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    a[i] = scale[cv * N + i];
    b[i] = shift[cv * N + i];
  }
  for (long long v0 = g; v0 < nvec; v0 += 2 * T_all) {
    const long long v1 = v0 + T_all;
    const bool h1 = v1 < nvec;
    float f0[N], f1[N], r0[N], r1[N];
    V<T>::load(x + v0 * N, f0);
    if (h1) V<T>::load(x + v1 * N, f1);
    if (res) {
      V<T>::load(res + v0 * N, r0);
      if (h1) V<T>::load(res + v1 * N, r1);
    }
    uint32_t m0 = 0, m1 = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float o = fmaf(a[i], f0[i], b[i]);
      if (res) o += r0[i];
      if (relu) {
        if (o > 0.f) m0 |= 1u << i; else o = 0.f;
      }
      f0[i] = o;
    }
    V<T>::store(y + v0 * N, f0);
    if (h1) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        float o = fmaf(a[i], f1[i], b[i]);
        if (res) o += r1[i];
        if (relu) {
          if (o > 0.f) m1 |= 1u << i; else o = 0.f;
        }
        f1[i] = o;
      }
      V<T>::store(y + v1 * N, f1);
    }
    if (relu) {
      mask[v0] = (uint8_t)m0;
      if (h1) mask[v1] = (uint8_t)m1;
    }
  }
}

// ---------------- backward pass 1: partial sum(dz), sum(dz * x) -------------------------------------------
// (sum(dz * xhat) = rstd * (sum(dz*x) - mean * sum(dz)) is formed in the finalize kernel: the streaming
// loop keeps no per-channel constants in registers, which buys occupancy.)
// HAS2: the incoming gradient is the sum of two tensors (the block output feeds both the next block's first convolution
// and its identity / downsample branch); adding them here saves the separate add pass autograd would run.
template <typename T, bool HAS2>
__global__ void __launch_bounds__(kThreads)
bn_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ dy2, const T* __restrict__ x,
                     const uint8_t* __restrict__ mask, float* __restrict__ partial, long long nvec, int C, int relu) {
  constexpr int N = V<T>::N;
  const int CV = C / N;
  extern __shared__ float smem[];
  const long long T_all = (long long)gridDim.x * kThreads;
  const long long g = (long long)blockIdx.x * kThreads + threadIdx.x;
  const int cvv = (int)(g % CV);
  float acc[2 * N];
#pragma unroll
  for (int i = 0; i < 2 * N; ++i) acc[i] = 0.f;
  for (long long v0 = g; v0 < nvec; v0 += 2 * T_all) {
    const long long v1 = v0 + T_all;
    const bool h1 = v1 < nvec;
    float d0[N], d1[N], f0[N], f1[N];
    V<T>::load(dy + v0 * N, d0);
    V<T>::load(x + v0 * N, f0);
    uint32_t m0 = relu ? mask[v0] : 0xffu, m1 = 0xffu;
    if (h1) {
      V<T>::load(dy + v1 * N, d1);
      V<T>::load(x + v1 * N, f1);
      if (relu) m1 = mask[v1];
    }
    if (HAS2) {
      float e0[N], e1[N];
      V<T>::load(dy2 + v0 * N, e0);
      if (h1) V<T>::load(dy2 + v1 * N, e1);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        d0[i] += e0[i];
        if (h1) d1[i] += e1[i];
      }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const float dz = (m0 >> i) & 1u ? d0[i] : 0.f;
      acc[i] += dz;
      acc[N + i] = fmaf(dz, f0[i], acc[N + i]);
    }
    if (h1) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const float dz = (m1 >> i) & 1u ? d1[i] : 0.f;
        acc[i] += dz;
        acc[N + i] = fmaf(dz, f1[i], acc[N + i]);
      }
    }
  }
  if (CV < kThreads) block_reduce_groups<2 * N>(acc, CV, smem);
  if (CV >= kThreads || threadIdx.x < CV) {
    const int cv = CV >= kThreads ? cvv : threadIdx.x;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      partial[(long long)blockIdx.x * 2 * C + cv * N + i] = acc[i];
      partial[(long long)blockIdx.x * 2 * C + C + cv * N + i] = acc[N + i];
    }
  }
}

__global__ void __launch_bounds__(kFinX * kFinY)
bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, float inv_m,
                       const float* __restrict__ weight, const float* __restrict__ mean, const float* __restrict__ rstd,
                       float* __restrict__ dweight, float* __restrict__ dbias, float* __restrict__ c0,
                       float* __restrict__ c1, float* __restrict__ c2, int accumulate) {
  const int c = blockIdx.x * kFinX + threadIdx.x;
  float s, q;
  sum_partials(partial, nblocks, C, c, s, q);
  if (c >= C || threadIdx.y != 0) return;
  const float mu = mean[c], rs = rstd[c];
  const float dgamma = rs * (q - mu * s);      // sum(dz * xhat)
  // accumulate: dweight / dbias are views of a gradient bucket that other backward passes also add to
  dbias[c] = accumulate ? dbias[c] + s : s;
  dweight[c] = accumulate ? dweight[c] + dgamma : dgamma;
  // dx = a*(dz - s/M - xhat*dgamma/M)  with a = gamma*rstd, xhat = (x-mu)*rs   ==>   dx = A*dz + B*x + D
  const float a = (weight ? weight[c] : 1.f) * rs;
  const float k2 = dgamma * inv_m * rs;
  c0[c] = a;
  c1[c] = -a * k2;
  c2[c] = a * (k2 * mu - s * inv_m);
}

// ---------------- backward pass 2: dx = A*dz + B*x + D (and dz for the residual branch) -------------------
template <typename T> __device__ __forceinline__ void store_streaming(T* p, const float* f);
template <> __device__ __forceinline__ void store_streaming<__nv_bfloat16>(__nv_bfloat16* p, const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  __stcs(reinterpret_cast<uint4*>(p), make_uint4(w[0], w[1], w[2], w[3]));
}
template <> __device__ __forceinline__ void store_streaming<float>(float* p, const float* f) {
  __stcs(reinterpret_cast<float4*>(p), make_float4(f[0], f[1], f[2], f[3]));
}

// Five streams (dy, x, mask in; dx, dz out): U vectors per thread are loaded before any math or store so that enough
// bytes are in flight to cover HBM latency, and the outputs (each read exactly once, much later, by the next backward
// kernel) are written with streaming stores so they do not evict the inputs of this pass from L2.
template <typename T, int U, bool STREAM, int MINB, bool HAS2 = false>
__global__ void __launch_bounds__(kThreads, MINB)
bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ dy2, const T* __restrict__ x, const uint8_t* __restrict__ mask,
                    const float* __restrict__ c0, const float* __restrict__ c1, const float* __restrict__ c2,
                    T* __restrict__ dx, T* __restrict__ dz_out, long long nvec, int C, int relu) {
  constexpr int N = V<T>::N;
  const int CV = C / N;
  const long long T_all = (long long)gridDim.x * kThreads;
  const long long g = (long long)blockIdx.x * kThreads + threadIdx.x;
  const int cv = (int)(g % CV);
  float A[N], B[N], D[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    A[i] = c0[cv * N + i];
    B[i] = c1[cv * N + i];
    D[i] = c2[cv * N + i];
  }
  for (long long v0 = g; v0 < nvec; v0 += U * T_all) {
    float d[U][N], f[U][N];
    uint32_t m[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long v = v0 + u * T_all;
      m[u] = 0xffu;
      if (v < nvec) {
        V<T>::load(dy + v * N, d[u]);
        V<T>::load(x + v * N, f[u]);
        if (relu) m[u] = mask[v];
      }
    }
    if (HAS2) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long v = v0 + u * T_all;
        if (v < nvec) {
          float e[N];
          V<T>::load(dy2 + v * N, e);
#pragma unroll
          for (int i = 0; i < N; ++i) d[u][i] += e[i];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long v = v0 + u * T_all;
      if (v < nvec) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const float dz = (m[u] >> i) & 1u ? d[u][i] : 0.f;
          d[u][i] = dz;
          f[u][i] = fmaf(A[i], dz, fmaf(B[i], f[u][i], D[i]));
        }
        if (STREAM) {
          store_streaming<T>(dx + v * N, f[u]);
          if (dz_out) store_streaming<T>(dz_out + v * N, d[u]);
        } else {
          V<T>::store(dx + v * N, f[u]);
          if (dz_out) V<T>::store(dz_out + v * N, d[u]);
        }
      }
    }
  }
}

static int grid_for(long long nvec, int CV) {
  long long blocks = (nvec + kThreads * 4 - 1) / (kThreads * 4);
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 4) blocks = 148 * 4;
  // total threads must be a multiple of CV so that every thread keeps one channel group
  if (CV > kThreads) {
    const int mult = CV / kThreads;
    blocks = (blocks + mult - 1) / mult * mult;
  }
  return (int)blocks;
}

}  // namespace bn

using namespace bn;

extern "C" {

// workspace floats needed: grid * 2C (partials) ; returns the grid the launches will use
int sdlc_bn_grid(long long m, int c, int dtype) {
  const int n = dtype == 6 ? 4 : 8;
  return grid_for(m * (c / n), c / n);
}

// dtype: 5 = bf16, 6 = fp32.  Returns 0 on success.
int sdlc_bn_fwd(const void* x, const void* res, void* y, void* mask, const float* weight, const float* bias,
                float* running_mean, float* running_var, float* mean, float* rstd, float* scale, float* shift,
                float* partial, long long m, int c, float eps, float momentum, int relu, int dtype,
                unsigned long long stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int n = dtype == 6 ? 4 : 8;
  if (c % n != 0 || (dtype != 5 && dtype != 6)) return -2;
  const int cv = c / n;
  if ((cv < kThreads && kThreads % cv != 0) || (cv >= kThreads && cv % kThreads != 0)) return -3;
  const long long nvec = m * cv;
  const int grid = grid_for(nvec, cv);
  const size_t smem = cv < kThreads ? sizeof(float) * kThreads * 2 * n : 0;
  if (cv > kThreads) cudaMemsetAsync(partial, 0, sizeof(float) * (size_t)grid * 2 * c, s);
  if (dtype == 5) bn_stats_kernel<__nv_bfloat16><<<grid, kThreads, smem, s>>>(static_cast<const __nv_bfloat16*>(x), partial, nvec, c);
  else bn_stats_kernel<float><<<grid, kThreads, smem, s>>>(static_cast<const float*>(x), partial, nvec, c);
  const float unbias = m > 1 ? (float)m / (float)(m - 1) : 1.f;
  bn_finalize_kernel<<<(c + kFinX - 1) / kFinX, dim3(kFinX, kFinY), 0, s>>>(partial, grid, c, 1.f / (float)m, eps, weight, bias, mean, rstd, scale,
                                                     shift, running_mean, running_var, momentum, unbias);
  if (dtype == 5)
    bn_apply_kernel<__nv_bfloat16><<<grid, kThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(res),
                                                            static_cast<__nv_bfloat16*>(y), static_cast<uint8_t*>(mask), scale, shift, nvec, c, relu);
  else
    bn_apply_kernel<float><<<grid, kThreads, 0, s>>>(static_cast<const float*>(x), static_cast<const float*>(res), static_cast<float*>(y),
                                                    static_cast<uint8_t*>(mask), scale, shift, nvec, c, relu);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int sdlc_bn_bwd(const void* dy, const void* dy2, const void* x, const void* mask, const float* weight, const float* mean, const float* rstd,
                void* dx, void* dz, float* dweight, float* dbias, float* c0, float* c1, float* c2, float* partial,
                long long m, int c, int relu, int dtype, int accumulate, unsigned long long stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int n = dtype == 6 ? 4 : 8;
  if (c % n != 0 || (dtype != 5 && dtype != 6)) return -2;
  const int cv = c / n;
  const long long nvec = m * cv;
  const int grid = grid_for(nvec, cv);
  const size_t smem = cv < kThreads ? sizeof(float) * kThreads * 2 * n : 0;
  if (cv > kThreads) cudaMemsetAsync(partial, 0, sizeof(float) * (size_t)grid * 2 * c, s);
#define BWD_REDUCE(TT, H2)                                                                                                 \
  bn_bwd_reduce_kernel<TT, H2><<<grid, kThreads, smem, s>>>(static_cast<const TT*>(dy), static_cast<const TT*>(dy2),       \
                                                            static_cast<const TT*>(x), static_cast<const uint8_t*>(mask), \
                                                            partial, nvec, c, relu)
  if (dtype == 5) { if (dy2) BWD_REDUCE(__nv_bfloat16, true); else BWD_REDUCE(__nv_bfloat16, false); }
  else { if (dy2) BWD_REDUCE(float, true); else BWD_REDUCE(float, false); }
  bn_bwd_finalize_kernel<<<(c + kFinX - 1) / kFinX, dim3(kFinX, kFinY), 0, s>>>(partial, grid, c, 1.f / (float)m, weight, mean, rstd, dweight, dbias, c0, c1, c2, accumulate);
  static int variant = -1;
  if (variant < 0) { const char* e = getenv("SPARKDL_BN_BWD_VARIANT"); variant = e ? atoi(e) : 0; }
#define BWD_APPLY2(TT, UU, ST, MB, H2)                                                                                     \
  bn_bwd_apply_kernel<TT, UU, ST, MB, H2><<<grid, kThreads, 0, s>>>(static_cast<const TT*>(dy), static_cast<const TT*>(dy2), \
                                                                   static_cast<const TT*>(x), static_cast<const uint8_t*>(mask), \
                                                                   c0, c1, c2, static_cast<TT*>(dx), static_cast<TT*>(dz), nvec, c, relu)
#define BWD_APPLY(TT, UU, ST, MB) BWD_APPLY2(TT, UU, ST, MB, false)
  if (dy2) {   // fork: gradient = dy + dy2 (6 streams)
    static int fv = -1;
    if (fv < 0) { const char* e = getenv("SPARKDL_BN_FORK_VARIANT"); fv = e ? atoi(e) : 0; }
    if (dtype == 5) {
      switch (fv) {
        case 1: BWD_APPLY2(__nv_bfloat16, 2, false, 4, true); break;
        case 2: BWD_APPLY2(__nv_bfloat16, 1, false, 4, true); break;
        case 3: BWD_APPLY2(__nv_bfloat16, 2, true, 3, true); break;
        default: BWD_APPLY2(__nv_bfloat16, 2, false, 3, true); break;
      }
    } else {
      BWD_APPLY2(float, 2, true, 2, true);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }
  if (dtype == 5) {
    switch (variant) {
      case 1: BWD_APPLY(__nv_bfloat16, 2, true, 3); break;
      case 2: BWD_APPLY(__nv_bfloat16, 2, false, 3); break;
      case 3: BWD_APPLY(__nv_bfloat16, 4, false, 2); break;
      case 4: BWD_APPLY(__nv_bfloat16, 3, true, 3); break;
      case 5: BWD_APPLY(__nv_bfloat16, 2, false, 4); break;
      case 6: BWD_APPLY(__nv_bfloat16, 4, true, 2); break;
      default: BWD_APPLY(__nv_bfloat16, 2, false, 4); break;   // 64 registers -> 4 CTAs / SM: 85-89 % of the measured HBM peak (bn_variants, round 2)
    }
  } else {
    BWD_APPLY(float, 4, true, 2);
  }
#undef BWD_APPLY
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // extern "C"
