// Fused  y = LayerNorm(x + dropout(a))  for transformer blocks (BERT-large, BASELINE.json config 3), forward and
// backward, bf16 activations, fp32 statistics.
//
// Eager PyTorch runs dropout, add and layer-norm as three kernels forward (7 activation passes) and four backward;
// here each direction is one streaming pass: one warp per row, 16-byte vectors, row statistics by warp shuffles,
// a 1-bit dropout mask, Philox-4x32-10 counters taken from a DEVICE counter (so a captured CUDA graph draws fresh
// masks on every replay).  gamma/beta gradients: per-block partials + deterministic finalize.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lnf {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

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
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Philox-4x32-10: 128 random bits per (counter, key)
__device__ __forceinline__ uint4 philox(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// x, a, y, z: [T, H] bf16 (a may be null: plain LayerNorm(x)); gamma/beta: [H] bf16; mask: [T*H/8] bytes;
// mean/rstd: [T] fp32; rng: device uint64 {seed, offset}; NV = H / 256 vectors per lane
template <int NV>
__global__ void __launch_bounds__(kThreads)
ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ gamma,
              const __nv_bfloat16* __restrict__ beta, __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ z,
              uint8_t* __restrict__ mask, float* __restrict__ mean, float* __restrict__ rstd,
              const unsigned long long* __restrict__ rng, unsigned int call_id, int T, float p, float eps) {
  constexpr int H = NV * 256;
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * kWarps;
  float g[NV][8], b[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    unpack8(*reinterpret_cast<const uint4*>(gamma + (lane + 32 * v) * 8), g[v]);
    unpack8(*reinterpret_cast<const uint4*>(beta + (lane + 32 * v) * 8), b[v]);
  }
  const bool drop = a != nullptr && p > 0.f;
  const float keep_scale = drop ? 1.f / (1.f - p) : 1.f;
  const uint32_t thresh = (uint32_t)(p * 65536.f);
  uint2 key = make_uint2(0u, 0u);
  unsigned long long off = 0;
  if (drop) {
    key = make_uint2((uint32_t)rng[0], (uint32_t)(rng[0] >> 32));
    off = rng[1];
  }
  for (long long row = warp; row < T; row += nwarps) {
    float zf[NV][8];
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const long long vi = row * (H / 8) + lane + 32 * v;
      unpack8(reinterpret_cast<const uint4*>(x)[vi], zf[v]);
      if (a != nullptr) {
        float af[8];
        unpack8(reinterpret_cast<const uint4*>(a)[vi], af);
        uint32_t m = 0xffu;
        if (drop) {
          const uint4 r = philox(make_uint4((uint32_t)vi, (uint32_t)(vi >> 32), (uint32_t)off, call_id ^ (uint32_t)(off >> 32)), key);
          const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
          m = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (((rr[i >> 1] >> ((i & 1) * 16)) & 0xffffu) >= thresh) m |= 1u << i;
          mask[vi] = (uint8_t)m;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) zf[v][i] += ((m >> i) & 1u) ? af[i] * keep_scale : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) s += zf[v][i];
    }
    const float mu = warp_sum(s) * (1.f / H);
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = zf[v][i] - mu;
        q = fmaf(d, d, q);
      }
    const float rs = rsqrtf(warp_sum(q) * (1.f / H) + eps);
    if (lane == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const long long vi = row * (H / 8) + lane + 32 * v;
      if (z != nullptr) reinterpret_cast<uint4*>(z)[vi] = pack8(zf[v]);
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf((zf[v][i] - mu) * rs, g[v][i], b[v][i]);
      reinterpret_cast<uint4*>(y)[vi] = pack8(o);
    }
  }
}

__global__ void rng_advance_kernel(unsigned long long* rng) { rng[1] += 1; }

// backward: dz = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = gamma * dy ;  dx = dz ;  da = dz * mask / (1-p)
// partial: [gridDim.x, 2H] fp32 (dgamma, dbeta partial sums of this block's rows)
template <int NV>
__global__ void __launch_bounds__(kThreads)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ gamma,
              const uint8_t* __restrict__ mask, const float* __restrict__ mean, const float* __restrict__ rstd,
              __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ da, float* __restrict__ partial, int T, float p) {
  constexpr int H = NV * 256;
  __shared__ float acc_s[2 * H];
  for (int i = threadIdx.x; i < 2 * H; i += kThreads) acc_s[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * kWarps;
  float g[NV][8], dg[NV][8], db[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    unpack8(*reinterpret_cast<const uint4*>(gamma + (lane + 32 * v) * 8), g[v]);
#pragma unroll
    for (int i = 0; i < 8; ++i) { dg[v][i] = 0.f; db[v][i] = 0.f; }
  }
  const float keep_scale = (da != nullptr && p > 0.f) ? 1.f / (1.f - p) : 1.f;
  for (long long row = warp; row < T; row += nwarps) {
    const float mu = mean[row], rs = rstd[row];
    float gy[NV][8], xh[NV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const long long vi = row * (H / 8) + lane + 32 * v;
      float d[8], zf[8];
      unpack8(reinterpret_cast<const uint4*>(dy)[vi], d);
      unpack8(reinterpret_cast<const uint4*>(z)[vi], zf);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xh[v][i] = (zf[i] - mu) * rs;
        db[v][i] += d[i];
        dg[v][i] = fmaf(d[i], xh[v][i], dg[v][i]);
        gy[v][i] = d[i] * g[v][i];
        s1 += gy[v][i];
        s2 = fmaf(gy[v][i], xh[v][i], s2);
      }
    }
    const float m1 = warp_sum(s1) * (1.f / H), m2 = warp_sum(s2) * (1.f / H);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const long long vi = row * (H / 8) + lane + 32 * v;
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = rs * (gy[v][i] - m1 - xh[v][i] * m2);
      reinterpret_cast<uint4*>(dx)[vi] = pack8(o);
      if (da != nullptr) {
        const uint32_t m = (p > 0.f) ? mask[vi] : 0xffu;
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = ((m >> i) & 1u) ? o[i] * keep_scale : 0.f;
        reinterpret_cast<uint4*>(da)[vi] = pack8(o);
      }
    }
  }
  // combine the 8 warps of the block (shared-memory atomics, 2H addresses), then one partial row per block
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = (lane + 32 * v) * 8 + i;
      atomicAdd(&acc_s[c], dg[v][i]);
      atomicAdd(&acc_s[H + c], db[v][i]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * H; i += kThreads) partial[(long long)blockIdx.x * 2 * H + i] = acc_s[i];
}

// dgamma/dbeta = sum over blocks of the partial rows: 32 columns x 32 row-groups per block (the loop over partial
// rows is latency bound, so it is spread over 32 threads per column), fixed combination order.
__global__ void __launch_bounds__(1024)
ln_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, int H, float* __restrict__ dgamma,
                       float* __restrict__ dbeta) {
  __shared__ float sh[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < 2 * H)
    for (int b = threadIdx.y; b < nblocks; b += 32) s += partial[(long long)b * 2 * H + c];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < 2 * H) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 32; ++y) t += sh[y][threadIdx.x];
    if (c < H) dgamma[c] = t; else dbeta[c - H] = t;
  }
}

static int grid_rows(int T) {
  long long blocks = (T + kWarps - 1) / kWarps;
  if (blocks > 148 * 2) blocks = 148 * 2;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace lnf

using namespace lnf;

extern "C" {

int sdlc_ln_supported(int H) { return H == 256 || H == 512 || H == 768 || H == 1024; }
int sdlc_ln_grid(int T) { return grid_rows(T); }

#define LN_DISPATCH(NVV, CALL)                                                                                      \
  switch (NVV) {                                                                                                    \
    case 1: { constexpr int NV = 1; CALL; break; }                                                                  \
    case 2: { constexpr int NV = 2; CALL; break; }                                                                  \
    case 3: { constexpr int NV = 3; CALL; break; }                                                                  \
    case 4: { constexpr int NV = 4; CALL; break; }                                                                  \
    default: return -2;                                                                                             \
  }

int sdlc_ln_fwd(const void* x, const void* a, const void* gamma, const void* beta, void* y, void* z, void* mask, float* mean,
                float* rstd, unsigned long long* rng, unsigned int call_id, int T, int H, float p, float eps,
                unsigned long long stream) {
  if (!sdlc_ln_supported(H)) return -2;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = grid_rows(T);
  typedef const __nv_bfloat16* B;
  LN_DISPATCH(H / 256, (ln_fwd_kernel<NV><<<grid, kThreads, 0, s>>>((B)x, (B)a, (B)gamma, (B)beta, (__nv_bfloat16*)y,
                                                                    (__nv_bfloat16*)z, (uint8_t*)mask, mean, rstd, rng,
                                                                    call_id, T, p, eps)));
  if (a != nullptr && p > 0.f) rng_advance_kernel<<<1, 1, 0, s>>>(rng);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int sdlc_ln_bwd(const void* dy, const void* z, const void* gamma, const void* mask, const float* mean, const float* rstd,
                void* dx, void* da, float* partial, float* dgamma, float* dbeta, int T, int H, float p,
                unsigned long long stream) {
  if (!sdlc_ln_supported(H)) return -2;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = grid_rows(T);
  typedef const __nv_bfloat16* B;
  LN_DISPATCH(H / 256, (ln_bwd_kernel<NV><<<grid, kThreads, 0, s>>>((B)dy, (B)z, (B)gamma, (const uint8_t*)mask, mean, rstd,
                                                                    (__nv_bfloat16*)dx, (__nv_bfloat16*)da, partial, T, p)));
  ln_bwd_finalize_kernel<<<(2 * H + 31) / 32, dim3(32, 32), 0, s>>>(partial, grid, H, dgamma, dbeta);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // extern "C"
