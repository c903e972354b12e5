// 3x3 / stride 2 / pad 1 max-pooling for NHWC bf16 (the ResNet stem), forward + backward.
//
// torch's channels-last max-pool backward costs 1.7 ms per step at batch 256 (int64 indices, scatter).  Here the
// forward stores a 1-byte window-local argmax per output element and the backward is a pure gather: every input
// pixel looks at the (at most four) windows that contain it - no atomics, one streaming pass each way.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pool {

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

// x: [N,H,W,C] bf16, y: [N,OH,OW,C], idx: [N,OH,OW,C] uint8 (kh*3+kw of the arg-max), C % 8 == 0
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ idx,
                   int N, int H, int W, int C, int OH, int OW) {
  const int CV = C / 8;
  const long long total = (long long)N * OH * OW * CV;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(t % CV);
    long long r = t / CV;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH);
    const int n = (int)(r / OH);
    float best[8];
    uint32_t arg[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { best[i] = -INFINITY; arg[i] = 0; }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int h = oh * 2 - 1 + kh;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int w = ow * 2 - 1 + kw;
        if (w < 0 || w >= W) continue;
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(x + (((long long)n * H + h) * W + w) * C + cv * 8), f);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (f[i] > best[i] || f[i] != f[i]) { best[i] = f[i]; arg[i] = kh * 3 + kw; }
      }
    }
    const long long o = (((long long)n * OH + oh) * OW + ow) * C + cv * 8;
    *reinterpret_cast<uint4*>(y + o) = pack8(best);
    uint2 a;
    a.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
    a.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
    *reinterpret_cast<uint2*>(idx + o) = a;
  }
}

__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ idx, __nv_bfloat16* __restrict__ dx,
                   int N, int H, int W, int C, int OH, int OW) {
  const int CV = C / 8;
  const long long total = (long long)N * H * W * CV;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(t % CV);
    long long r = t / CV;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    // windows containing row h: oh*2-1 <= h <= oh*2+1
    const int oh0 = h >> 1, oh1 = (h + 1) >> 1;
    const int ow0 = w >> 1, ow1 = (w + 1) >> 1;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int oh = a == 0 ? oh0 : oh1;
      if ((a == 1 && oh1 == oh0) || oh >= OH) continue;
      const int kh = h - (oh * 2 - 1);
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int ow = b == 0 ? ow0 : ow1;
        if ((b == 1 && ow1 == ow0) || ow >= OW) continue;
        const int kw = w - (ow * 2 - 1);
        const uint32_t me = (uint32_t)(kh * 3 + kw);
        const long long o = (((long long)n * OH + oh) * OW + ow) * C + cv * 8;
        const uint2 ix = *reinterpret_cast<const uint2*>(idx + o);
        float g[8];
        unpack8(*reinterpret_cast<const uint4*>(dy + o), g);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t k = ((i < 4 ? ix.x : ix.y) >> ((i & 3) * 8)) & 0xffu;
          if (k == me) acc[i] += g[i];
        }
      }
    }
    *reinterpret_cast<uint4*>(dx + (((long long)n * H + h) * W + w) * C + cv * 8) = pack8(acc);
  }
}

// ---- 2x2 variants (even H, W): one thread owns a 2x2 block of outputs (forward) / inputs (backward), so the overlapping
// windows are loaded once: 25 instead of 36 input vectors per 4 outputs, 4 instead of 9 (dy, idx) pairs per 4 inputs.
__global__ void __launch_bounds__(256)
maxpool_fwd2x2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ idx,
                      int N, int H, int W, int C, int OH, int OW) {
  const int CV = C / 8, OH2 = OH / 2, OW2 = OW / 2;
  const long long total = (long long)N * OH2 * OW2 * CV;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(t % CV);
    long long r = t / CV;
    const int ow0 = (int)(r % OW2) * 2; r /= OW2;
    const int oh0 = (int)(r % OH2) * 2;
    const int n = (int)(r / OH2);
    float best[4][8];
    uint32_t arg[4][8];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int i = 0; i < 8; ++i) { best[o][i] = -INFINITY; arg[o][i] = 0; }
    // input rows 2*oh0-1 .. 2*oh0+3, cols 2*ow0-1 .. 2*ow0+3
#pragma unroll
    for (int dr = 0; dr < 5; ++dr) {
      const int h = oh0 * 2 - 1 + dr;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int dc = 0; dc < 5; ++dc) {
        const int w = ow0 * 2 - 1 + dc;
        if (w < 0 || w >= W) continue;
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + (((long long)n * H + h) * W + w) * C + cv * 8)), f);
        // output (a, b) of the block sees this pixel when dr - 2a and dc - 2b are in [0, 3)
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const int kh = dr - 2 * a;
          if (kh < 0 || kh > 2) continue;
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const int kw = dc - 2 * b;
            if (kw < 0 || kw > 2) continue;
            const int o = a * 2 + b;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (f[i] > best[o][i] || f[i] != f[i]) { best[o][i] = f[i]; arg[o][i] = kh * 3 + kw; }
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int o = a * 2 + b;
        const long long off = (((long long)n * OH + oh0 + a) * OW + ow0 + b) * C + cv * 8;
        __stcs(reinterpret_cast<uint4*>(y + off), pack8(best[o]));
        uint2 p;
        p.x = arg[o][0] | (arg[o][1] << 8) | (arg[o][2] << 16) | (arg[o][3] << 24);
        p.y = arg[o][4] | (arg[o][5] << 8) | (arg[o][6] << 16) | (arg[o][7] << 24);
        *reinterpret_cast<uint2*>(idx + off) = p;
      }
  }
}

__global__ void __launch_bounds__(256)
maxpool_bwd2x2_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ idx, __nv_bfloat16* __restrict__ dx,
                      int N, int H, int W, int C, int OH, int OW) {
  const int CV = C / 8, H2 = H / 2, W2 = W / 2;
  const long long total = (long long)N * H2 * W2 * CV;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(t % CV);
    long long r = t / CV;
    const int j = (int)(r % W2); r /= W2;
    const int i2 = (int)(r % H2);
    const int n = (int)(r / H2);
    // input pixels (2*i2 + p, 2*j + q); windows oh in {i2, i2+1}, ow in {j, j+1}:
    //   window row i2   covers input rows 2*i2-1 .. 2*i2+1  -> p = 0 (kh 1), p = 1 (kh 2)
    //   window row i2+1 covers input rows 2*i2+1 .. 2*i2+3  -> p = 1 (kh 0)
    float acc[4][8];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[o][i] = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int oh = i2 + a;
      if (oh >= OH) continue;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int ow = j + b;
        if (ow >= OW) continue;
        const long long off = (((long long)n * OH + oh) * OW + ow) * C + cv * 8;
        const uint2 ix = __ldg(reinterpret_cast<const uint2*>(idx + off));
        float g[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(dy + off)), g);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const int kh = (2 * i2 + p) - (oh * 2 - 1);
          if (kh < 0 || kh > 2) continue;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int kw = (2 * j + q) - (ow * 2 - 1);
            if (kw < 0 || kw > 2) continue;
            const uint32_t me = (uint32_t)(kh * 3 + kw);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t k = ((i < 4 ? ix.x : ix.y) >> ((i & 3) * 8)) & 0xffu;
              if (k == me) acc[p * 2 + q][i] += g[i];
            }
          }
        }
      }
    }
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = 0; q < 2; ++q)
        __stcs(reinterpret_cast<uint4*>(dx + (((long long)n * H + 2 * i2 + p) * W + 2 * j + q) * C + cv * 8), pack8(acc[p * 2 + q]));
  }
}

// ---- ResNet stem input: [N,3,H,W] (fp32 or bf16, any strides) -> space-to-depth NHWC bf16 [N, H/2+3, W/2+3, 16] ------------
// channel = (rh*2 + rw)*4 + c (c = 3 is zero), spatially zero-padded by 2 (top/left) and 1 (bottom/right), so that the
// 7x7 / stride-2 / pad-3 stem convolution becomes a dense 4x4 / stride-1 convolution with 16 input channels
// (models/resnet.py): 3.1x faster in cuDNN than the 3-channel original.
template <typename T>
__global__ void __launch_bounds__(256)
stem_s2d_kernel(const T* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int H, int W, long long sn, long long sc,
                long long sh, long long sw) {
  const int H2 = H / 2 + 3, W2 = W / 2 + 3;
  const long long total = (long long)N * H2 * W2;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int w2 = (int)(t % W2);
    long long r = t / W2;
    const int h2 = (int)(r % H2);
    const int n = (int)(r / H2);
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.f;
    const int hb = 2 * (h2 - 2), wb = 2 * (w2 - 2);
    if (hb >= 0 && hb < H && wb >= 0 && wb < W) {
#pragma unroll
      for (int rh = 0; rh < 2; ++rh)
#pragma unroll
        for (int rw = 0; rw < 2; ++rw)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            v[(rh * 2 + rw) * 4 + c] = (float)x[n * sn + c * sc + (long long)(hb + rh) * sh + (long long)(wb + rw) * sw];
    }
    uint4* o = reinterpret_cast<uint4*>(out + t * 16);
    o[0] = pack8(v);
    o[1] = pack8(v + 8);
  }
}

}  // namespace pool

extern "C" {

int sdlc_maxpool3s2_fwd(const void* x, void* y, void* idx, int N, int H, int W, int C, unsigned long long stream) {
  if (C % 8) return -2;
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  long long total = (long long)N * OH * OW * (C / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  // measured on B200 (bn_bench.py, 256x64x112x112): the 2x2 forward is SLOWER (321 us vs 219 us: 92 registers and a
  // branchy update), the 2x2 backward is 2.2x faster (157 us vs 344 us) - so only the backward uses its 2x2 variant
  if (false && (H % 4) == 0 && (W % 4) == 0) {
    total /= 4;
    blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    pool::maxpool_fwd2x2_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), static_cast<uint8_t*>(idx), N, H, W, C, OH, OW);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }
  pool::maxpool_fwd_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), static_cast<uint8_t*>(idx), N, H, W, C, OH, OW);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int sdlc_maxpool3s2_bwd(const void* dy, const void* idx, void* dx, int N, int H, int W, int C, unsigned long long stream) {
  if (C % 8) return -2;
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  long long total = (long long)N * H * W * (C / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  if ((H % 2) == 0 && (W % 2) == 0) {   // one thread per 2x2 block of inputs
    total /= 4;
    blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    pool::maxpool_bwd2x2_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(dy), static_cast<const uint8_t*>(idx), static_cast<__nv_bfloat16*>(dx), N, H, W, C, OH, OW);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }
  pool::maxpool_bwd_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const uint8_t*>(idx), static_cast<__nv_bfloat16*>(dx), N, H, W, C, OH, OW);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// x: [N,3,H,W] with element strides (sn, sc, sh, sw); dtype 6 = fp32, 5 = bf16; out: bf16 [N, H/2+3, W/2+3, 16]
int sdlc_stem_s2d(const void* x, void* out, int N, int H, int W, long long sn, long long sc, long long sh, long long sw,
                  int dtype, unsigned long long stream) {
  if ((H % 2) != 0 || (W % 2) != 0) return -2;
  const long long total = (long long)N * (H / 2 + 3) * (W / 2 + 3);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == 6)
    pool::stem_s2d_kernel<float><<<(int)blocks, 256, 0, s>>>(static_cast<const float*>(x), static_cast<__nv_bfloat16*>(out), N, H, W, sn, sc, sh, sw);
  else if (dtype == 5)
    pool::stem_s2d_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), N, H, W, sn, sc, sh, sw);
  else
    return -2;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // extern "C"
