// EXPERIMENTAL variant of gemm_tcgen05.cu (opt-in: SPARKDL_GEMM_EPI16=1; never used by default).
//
// Same TMA -> smem ring -> tcgen05.mma -> double-buffered TMEM pipeline, but SIXTEEN epilogue warps (four per TMEM
// lane quarter, each draining a quarter of the tile's columns in 32-column chunks) instead of eight with 64-column
// chunks.  Motivation (NOTES_NEXT_ROUND.md): on the K=1024 BERT shapes the kernel is epilogue-bound at ~3.5x the
// epilogue's issue-bound time - with two epilogue warps per scheduler, each running a serial
// tcgen05.ld -> math -> st.shared -> ld.shared -> st.global chain, the schedulers idle.  Four warps per scheduler at
// half the chain length should hide that latency.  ptxas: 96 registers/thread at 640 threads (fits the 64K file).
// Written after the last GPU call of round 1: compiles for sm_100a, NOT yet run on hardware.  Validate with
//   SPARKDL_GEMM_EPI16=1 python -m pytest tests/test_gpu_gemm.py
// before relying on it.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>

namespace gemm16 {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;      // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 640; // 20 warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-19 epilogue
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 16;    // four warps per TMEM lane quarter, each owning a quarter of the columns
template <int BLOCK_N> struct Stages { static constexpr int value = BLOCK_N == 256 ? 4 : 6; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] . B[smem]^T ; both operands K-major, swizzle-128B
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//  [0,14) start address >> 4, [16,30) LBO >> 4 (=1, unused for swizzled K-major), [32,46) SBO >> 4
//  (8 rows x 128 B = 1024 B between 8-row groups), [46,48) version = 1 (sm_100), [61,64) layout = 2.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) at [4,6), a/b format BF16 (1) at
// [7,10)/[10,13), a/b major K (0) at 15/16, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ float gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float u = k0 * fmaf(k1 * x * x, x, x);
  float t;                                  // one MUFU op (max rel. error 2^-11, below bf16 resolution); the IEEE
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));   // division of the baseline costs an FCHK + slow-path CALL
  return 0.5f * x * (1.f + t);
}
// erf via Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7, far below bf16 resolution): one MUFU.EX2 + one MUFU.RCP
__device__ __forceinline__ float fast_erf(float z) {
  const float a = fabsf(z);
  const float t = __fdividef(1.f, fmaf(0.3275911f, a, 1.f));
  const float p = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
  const float r = 1.f - p * __expf(-a * a);
  return copysignf(r, z);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + fast_erf(x * 0.7071067811865476f)); }

template <int BLOCK_N>
struct SmemLayout {
  static constexpr int kStages = Stages<BLOCK_N>::value;
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = kEpiWarps * 32 * 64;    // per epilogue warp: 32 rows x 32 bf16 columns
  static constexpr int kTotal = kStages * kStageBytes + kStagingBytes + 1024 /* alignment slack */ + 256 /* barriers */;
};

// epilogue: 0 none, 1 GELU(tanh), 2 GELU(erf).  pre_act (optional) receives X.W^T + bias before the
// activation (needed by the backward of GELU).
template <int BLOCK_N>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ pre_act, const float* __restrict__ bias,
                 int M, int N, int K, int epilogue) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using L = SmemLayout<BLOCK_N>;
  constexpr int kStages = L::kStages;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * L::kABytes;
  uint8_t* staging = smem + kStages * L::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + L::kStagingBytes);
  uint64_t* full = bars;                    // [kStages]
  uint64_t* empty = bars + kStages;         // [kStages]
  uint64_t* tmem_full = bars + 2 * kStages; // [2]
  uint64_t* tmem_empty = tmem_full + 2;     // [2]
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M, n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  constexpr uint32_t kTmemCols = 2 * BLOCK_N;  // power of two: 256 or 512

  if (warp == 0 && elect_one()) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kEpiWarps);   // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_smem)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], L::kStageBytes);
          tma_load_2d(smem_a + stage * L::kABytes, &map_a, &full[stage], kb * BLOCK_K, m_blk * BLOCK_M);
          tma_load_2d(smem_b + stage * L::kBBytes, &map_b, &full[stage], kb * BLOCK_K, n_blk * BLOCK_N);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);   // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&full[stage], phase);              // TMA bytes have landed
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem_a + stage * L::kABytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * L::kBBytes);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc(a_addr + k * UMMA_K * 2);
            const uint64_t db = make_smem_desc(b_addr + k * UMMA_K * 2);
            umma_f16(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);                 // frees the smem slot when these MMAs retire
          if (kb == k_blocks - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= kEpiWarp0) {
    // ===== epilogue: TMEM -> registers -> bias / GELU -> bf16 -> swizzled smem -> coalesced global stores =====
    const int q = (warp - kEpiWarp0) & 3;            // == warp % 4: TMEM lanes [32q, 32q+32)
    const int part = (warp - kEpiWarp0) >> 2;        // which quarter of the tile's columns this warp drains
    constexpr int kColsPerWarp = BLOCK_N / 4;
    uint8_t* stage = staging + (warp - kEpiWarp0) * (32 * 64);
    // thread = row while reading TMEM; 8 rows x 64 contiguous bytes per store instruction when writing out.
    // 16-byte chunk j of row r lives at chunk (j ^ (r >> 1)) & 3: conflict-free both ways (64-byte rows).
    auto stage_out = [&](const float* f, __nv_bfloat16* dst, int row0, int col) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 o;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(f[8 * j + 0], f[8 * j + 1]), h1 = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]), h3 = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
        o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
        o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(stage + lane * 64 + (((j ^ (lane >> 1)) & 3) << 4)) = o;
      }
      __syncwarp();
      const int chunk = lane & 3;
      const int c = col + chunk * 8;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int r = it * 8 + (lane >> 2);
        const uint4 o = *reinterpret_cast<const uint4*>(stage + r * 64 + (((chunk ^ (r >> 1)) & 3) << 4));
        if (row0 + r < M && c < N) *reinterpret_cast<uint4*>(dst + (size_t)(row0 + r) * N + c) = o;
      }
      __syncwarp();
    };
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row0 = m_blk * BLOCK_M + q * 32;
#pragma unroll 1
      for (int c0 = part * kColsPerWarp; c0 < (part + 1) * kColsPerWarp; c0 += 32) {
        uint32_t w[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
              "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]),
              "=r"(w[16]), "=r"(w[17]), "=r"(w[18]), "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]),
              "=r"(w[24]), "=r"(w[25]), "=r"(w[26]), "=r"(w[27]), "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int col0 = n_blk * BLOCK_N + c0;
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(w[i]);
        if (bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (col0 + j < N) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + col0 + j));
              f[j] += bv.x; f[j + 1] += bv.y; f[j + 2] += bv.z; f[j + 3] += bv.w;
            }
          }
        }
        if (pre_act) stage_out(f, pre_act, row0, col0);
        if (epilogue == 1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = gelu_tanh(f[i]);
        } else if (epilogue == 2) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = gelu_erf(f[i]);
        }
        stage_out(f, out, row0, col0);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---------------- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static bool load_encode() {
  if (g_encode) return true;
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr) != cudaSuccess || !fp) {
    cudaGetLastError();
    return false;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fp);
  return true;
}

// row-major [rows, cols] bf16 matrix, box = [box_rows, 64 cols], 128-byte swizzle
static bool make_map(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace gemm16

using namespace gemm16;

extern "C" {

// Y[M,N] = act(X[M,K] . W[N,K]^T + bias).  Requirements: K % 8 == 0, N % 8 == 0, 16-byte aligned pointers.
// Returns 0 on success, negative on error.
int sdlc_gemm_bf16_epi16(const void* x, const void* w, const float* bias, void* out, void* pre_act, int M, int N, int K,
                   int epilogue, int block_n, int num_sms, unsigned long long stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((K % 8) != 0 || (N % 8) != 0) return -2;
  if (!load_encode()) return -3;
  CUtensorMap ma, mb;
  if (block_n != 256) block_n = 128;
  if (!make_map(&ma, x, (uint64_t)M, (uint64_t)K, BLOCK_M)) return -4;
  if (!make_map(&mb, w, (uint64_t)N, (uint64_t)K, (uint32_t)block_n)) return -4;
  const int tiles = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + block_n - 1) / block_n);
  if (num_sms <= 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = tiles < num_sms ? tiles : num_sms;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (block_n == 256) {
    const int smem = SmemLayout<256>::kTotal;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(gemm_bf16_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
    gemm_bf16_kernel<256><<<grid, kNumThreads, smem, s>>>(ma, mb, static_cast<__nv_bfloat16*>(out), static_cast<__nv_bfloat16*>(pre_act), bias, M, N, K, epilogue);
  } else {
    const int smem = SmemLayout<128>::kTotal;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(gemm_bf16_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
    gemm_bf16_kernel<128><<<grid, kNumThreads, smem, s>>>(ma, mb, static_cast<__nv_bfloat16*>(out), static_cast<__nv_bfloat16*>(pre_act), bias, M, N, K, epilogue);
  }
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -1;
}

}  // extern "C"
