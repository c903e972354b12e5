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

}  // namespace pool

extern "C" {

int sdlc_maxpool3s2_fwd(const void* x, void* y, void* idx, int N, int H, int W, int C, unsigned long long stream) {
  if (C % 8) return -2;
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  long long total = (long long)N * OH * OW * (C / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
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
  pool::maxpool_bwd_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const uint8_t*>(idx), static_cast<__nv_bfloat16*>(dx), N, H, W, C, OH, OW);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // extern "C"
