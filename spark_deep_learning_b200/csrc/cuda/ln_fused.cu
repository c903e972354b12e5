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
// mean/rstd: [T] fp32; rng: device uint64 {seed, offset}; NV = H / 256 vectors per lane.
// One warp per row, ONE row per warp where the grid allows it: the only per-thread state is the row itself
// (gamma / beta are re-read through L1 in the output phase), so 3 CTAs fit per SM and a 4096-row call is a single wave
// (round 1 kept gamma/beta in 64 registers: 164 registers, 1 CTA / SM, 18.5 us for 32 MB of traffic).
template <int NV>
__global__ void __launch_bounds__(kThreads, NV >= 4 ? 2 : 3)
ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ gamma,
              const __nv_bfloat16* __restrict__ beta, __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ z,
              uint8_t* __restrict__ mask, float* __restrict__ mean, float* __restrict__ rstd,
              const unsigned long long* __restrict__ rng, unsigned int call_id, int T, float p, float eps) {
  constexpr int H = NV * 256;
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * kWarps;
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
    uint4 xv[NV], av[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {   // all loads of the row first
      const long long vi = row * (H / 8) + lane + 32 * v;
      xv[v] = reinterpret_cast<const uint4*>(x)[vi];
      if (a != nullptr) av[v] = reinterpret_cast<const uint4*>(a)[vi];
    }
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const long long vi = row * (H / 8) + lane + 32 * v;
      unpack8(xv[v], zf[v]);
      if (a != nullptr) {
        float af[8];
        unpack8(av[v], af);
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
      float g[8], b[8], o[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(gamma) + lane + 32 * v), g);
      unpack8(__ldg(reinterpret_cast<const uint4*>(beta) + lane + 32 * v), b);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf((zf[v][i] - mu) * rs, g[i], b[i]);
      reinterpret_cast<uint4*>(y)[vi] = pack8(o);
    }
  }
}

__global__ void rng_advance_kernel(unsigned long long* rng) { rng[1] += 1; }

// backward, row pass: dz = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = gamma * dy ;  dx = dz ;
// da = dz * mask / (1-p).  The row stays packed (bf16) in registers between the two phases.
template <int NV>
__global__ void __launch_bounds__(kThreads, NV >= 3 ? 2 : 3)
ln_bwd_row_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ gamma,
                  const uint8_t* __restrict__ mask, const float* __restrict__ mean, const float* __restrict__ rstd,
                  __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ da, int T, float p) {
  constexpr int H = NV * 256;
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * kWarps;
  const float keep_scale = (da != nullptr && p > 0.f) ? 1.f / (1.f - p) : 1.f;
  for (long long row = warp; row < T; row += nwarps) {
    const float mu = mean[row], rs = rstd[row];
    uint4 dv[NV], zv[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const long long vi = row * (H / 8) + lane + 32 * v;
      dv[v] = reinterpret_cast<const uint4*>(dy)[vi];
      zv[v] = reinterpret_cast<const uint4*>(z)[vi];
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float d[8], zf[8], g[8];
      unpack8(dv[v], d);
      unpack8(zv[v], zf);
      unpack8(__ldg(reinterpret_cast<const uint4*>(gamma) + lane + 32 * v), g);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float gy = d[i] * g[i];
        s1 += gy;
        s2 = fmaf(gy, (zf[i] - mu) * rs, s2);
      }
    }
    const float m1 = warp_sum(s1) * (1.f / H), m2 = warp_sum(s2) * (1.f / H);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const long long vi = row * (H / 8) + lane + 32 * v;
      float d[8], zf[8], g[8], o[8];
      unpack8(dv[v], d);
      unpack8(zv[v], zf);
      unpack8(__ldg(reinterpret_cast<const uint4*>(gamma) + lane + 32 * v), g);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = rs * (d[i] * g[i] - m1 - (zf[i] - mu) * rs * m2);
      reinterpret_cast<uint4*>(dx)[vi] = pack8(o);
      if (da != nullptr) {
        const uint32_t m = (p > 0.f) ? mask[vi] : 0xffu;
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = ((m >> i) & 1u) ? o[i] * keep_scale : 0.f;
        reinterpret_cast<uint4*>(da)[vi] = pack8(o);
      }
    }
  }
}

// backward, column pass: partial[b][0:H] = sum_rows dy * xhat, partial[b][H:2H] = sum_rows dy over the rows of block b.
// One thread per 8-column vector, rows of the slice streamed 4 at a time (the second read of dy / z mostly hits L2).
__global__ void __launch_bounds__(128)
ln_bwd_col_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z, const float* __restrict__ mean,
                  const float* __restrict__ rstd, float* __restrict__ partial, int T, int H) {
  const int cv = threadIdx.x;
  if (cv >= H / 8) return;
  const int R = gridDim.x;
  const int rows_per = (T + R - 1) / R;
  const int r0 = blockIdx.x * rows_per, r1 = min(T, r0 + rows_per);
  float dg[8], db[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { dg[i] = 0.f; db[i] = 0.f; }
  int r = r0;
  for (; r + 3 < r1; r += 4) {
    uint4 dv[4], zv[4];
    float mu[4], rs[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      dv[u] = reinterpret_cast<const uint4*>(dy)[(long long)(r + u) * (H / 8) + cv];
      zv[u] = reinterpret_cast<const uint4*>(z)[(long long)(r + u) * (H / 8) + cv];
      mu[u] = mean[r + u];
      rs[u] = rstd[r + u];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float d[8], zf[8];
      unpack8(dv[u], d);
      unpack8(zv[u], zf);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        db[i] += d[i];
        dg[i] = fmaf(d[i], (zf[i] - mu[u]) * rs[u], dg[i]);
      }
    }
  }
  for (; r < r1; ++r) {
    float d[8], zf[8];
    unpack8(reinterpret_cast<const uint4*>(dy)[(long long)r * (H / 8) + cv], d);
    unpack8(reinterpret_cast<const uint4*>(z)[(long long)r * (H / 8) + cv], zf);
    const float mu = mean[r], rs = rstd[r];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      db[i] += d[i];
      dg[i] = fmaf(d[i], (zf[i] - mu) * rs, dg[i]);
    }
  }
  float* out = partial + (long long)blockIdx.x * 2 * H + cv * 8;
#pragma unroll
  for (int i = 0; i < 8; i += 4) {
    *reinterpret_cast<float4*>(out + i) = make_float4(dg[i], dg[i + 1], dg[i + 2], dg[i + 3]);
    *reinterpret_cast<float4*>(out + H + i) = make_float4(db[i], db[i + 1], db[i + 2], db[i + 3]);
  }
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

static int grid_rows(int T) {   // row kernels: one row per warp where possible, at most 3 CTAs per SM
  long long blocks = (T + kWarps - 1) / kWarps;
  if (blocks > 148 * 3) blocks = 148 * 3;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}
static int grid_cols(int T) {   // column kernel: number of row slices (= partial rows)
  long long blocks = (T + 15) / 16;
  if (blocks > 148 * 2) blocks = 148 * 2;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace lnf

using namespace lnf;

extern "C" {

int sdlc_ln_supported(int H) { return H == 256 || H == 512 || H == 768 || H == 1024; }
int sdlc_ln_grid(int T) { return grid_cols(T); }   // partial rows of the backward column pass

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
  const int grid = grid_rows(T), cgrid = grid_cols(T);
  typedef const __nv_bfloat16* B;
  LN_DISPATCH(H / 256, (ln_bwd_row_kernel<NV><<<grid, kThreads, 0, s>>>((B)dy, (B)z, (B)gamma, (const uint8_t*)mask, mean, rstd,
                                                                        (__nv_bfloat16*)dx, (__nv_bfloat16*)da, T, p)));
  ln_bwd_col_kernel<<<cgrid, 128, 0, s>>>((B)dy, (B)z, mean, rstd, partial, T, H);
  ln_bwd_finalize_kernel<<<(2 * H + 31) / 32, dim3(32, 32), 0, s>>>(partial, cgrid, H, dgamma, dbeta);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // extern "C"
