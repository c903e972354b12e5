// bf16 GEMM on the 5th-generation tensor cores for sm_100a, hand-written: CTA-pair tiles (tcgen05.mma.cta_group::2,
// UMMA 256 x BLOCK_N x 16), operands staged by TMA into a 128B-swizzled shared-memory ring, accumulators in TMEM,
// fused epilogues drained with tcgen05.ld and written back with TMA stores.
//
//     D[M,N] = epilogue( A[M,K] . B[N,K]^T )
//
// Each operand is either K-major (stored [rows, K], K contiguous) or MN-major (stored [K, rows], rows contiguous),
// so the three GEMMs of a linear layer all run here without a transpose:
//     forward  Y  = X . W^T           A = X  [M,K]  K-major      B = W  [N,K]  K-major
//     dgrad    dX = dY . W            A = dY [M,K'] K-major      B = W  [K',N] MN-major
//     wgrad    dW = dY^T . X          A = dY [K'',M] MN-major    B = X  [K'',N] MN-major   (split-K over tokens)
// and a 1x1 convolution in NHWC is the forward case with M = N*H*W.
//
// Roles inside a CTA (384 threads; the two CTAs of a cluster run the same code, blackwell_cuda_programming.md §3):
//   warp 0      TMA producer: each CTA loads its own 128 rows of A and its half of the B tile
//               (cp.async.bulk.tensor.2d.cta_group::2, completion counted on the LEADER CTA's mbarrier)
//   warp 1      MMA issuer (leader CTA only): one elected lane issues tcgen05.mma.cta_group::2.kind::f16; the
//               hardware reads A/B from both CTAs' shared memory and writes 128 accumulator rows into each CTA's TMEM;
//               tcgen05.commit...multicast::cluster releases smem slots / publishes accumulators in BOTH CTAs
//   warp 2      TMEM allocator (tcgen05.alloc.cta_group::2, 2 x BLOCK_N fp32 columns: double-buffered accumulator)
//   warps 4-11  epilogue: tcgen05.ld 32x32b.x32 -> registers -> epilogue math -> swizzled smem -> TMA store
//               (double-buffered staging per warp), overlapped with the MMAs of the next tile
//
// Epilogues (template parameter EPI):
//   EPI_LINEAR    + bias[N], optional GELU (tanh / erf), optional second output with the pre-activation    -> bf16
//   EPI_DGELU     acc * gelu'(aux[M,N])   (backward of the fused bias+GELU: no separate elementwise pass)   -> bf16
//   EPI_ACC_BF16  dst (+)= acc            (wgrad written / accumulated straight into a gradient bucket)     -> bf16
//   EPI_ACC_F32   same with fp32 output; also the partial-sum output of split-K                             -> fp32
//   EPI_STATS     bf16 output + per-column sum / sum-of-squares of the rounded outputs (BatchNorm statistics of a
//                 1x1 convolution come out of the GEMM instead of a second pass over its output)
//
// Persistent: grid = 2 * min(#SMs / 2, work units); a work unit is (tile, k-slice); all roles walk the same static
// schedule (grouped raster: 8 M-tiles share their B tiles in L2).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace gemm {

constexpr int CTA_M = 128;       // accumulator rows per CTA (= TMEM lanes); a CTA pair owns a 256-row tile
constexpr int PAIR_M = 256;
constexpr int BLOCK_K = 64;      // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 384; // 12 warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 epilogue
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;     // two warps per TMEM lane quarter, each owning half of the tile's columns
constexpr int kStageBufBytes = 4096;                       // 32 rows x 128 bytes
constexpr int kMaxStages = 8;
constexpr int kMaxSmem = 232448;                           // 227 KB: the opt-in maximum per CTA
constexpr int kGroupM = 8;       // raster group: this many M-tiles walk the N-tiles together

enum { EPI_LINEAR = 0, EPI_DGELU = 1, EPI_ACC_BF16 = 2, EPI_ACC_F32 = 3, EPI_STATS = 4 };

template <int BLOCK_N>
struct Cfg {
  static constexpr int kBRows = BLOCK_N / 2;                 // rows of the B tile held by each CTA
  static constexpr int kABytes = CTA_M * BLOCK_K * 2;        // 16 KB
  static constexpr int kBBytes = kBRows * BLOCK_K * 2;       // 16 KB (BLOCK_N 256) / 8 KB (128)
  static constexpr int kStageBytes = kABytes + kBBytes;
  // shared memory: stages x (A + B) | 8 warps x staging_bufs x 4 KB | barriers; the ring takes what staging leaves
  static constexpr int smem_bytes(int stages, int staging_bufs) {
    return stages * kStageBytes + kEpiWarps * staging_bufs * kStageBufBytes + 1024 /* alignment slack */ + 256 /* barriers */;
  }
  static constexpr int max_stages(int staging_bufs) {
    int s = (kMaxSmem - smem_bytes(0, staging_bufs)) / kStageBytes;
    return s > kMaxStages ? kMaxStages : s;
  }
};

struct GemmParams {
  int M, N, K;
  int a_mn, b_mn;              // 1: operand is MN-major (stored [K, rows])
  int a_parts, a_part_len;     // A may be split into up to 3 buffers: along K (K-major A) or along M (MN-major A)
  int splits, kb_per_split;    // split-K: number of slices / k-blocks per slice
  int m_pad;                   // split-K workspace: rows per slice (m_tiles * 256)
  int act;                     // LINEAR: 0 none, 1 GELU(tanh), 2 GELU(erf); DGELU: 1 / 2
  int has_pre;                 // LINEAR: also store through map_out2 the pre-activation (1) or act'(pre-activation) (2)
  int accumulate;              // ACC: out = dst + acc
  const void* bias;            // LINEAR: fp32 (or bf16, bias_bf16) [N] or null
  int bias_bf16;
  const void* aux;             // DGELU: bf16 pre-activation [M, ld_aux]; ACC: destination read for accumulate
  long long ld_aux;
  float* stats;                // STATS: partial sums [2][stat_rows][N] (sum, then sum of squares)
  int stat_rows;               // STATS: number of 32-row groups = m_tiles * 8
  int stages, staging_bufs;    // smem ring depth; staging buffers per epilogue warp (1 or 2)
  int dbg;                     // lab switches (SPARKDL_GEMM_DBG): 1 = no TMA loads, 2 = no MMAs, 4 = no epilogue stores
};

// ---------------- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of `local` (a shared::cta address) inside CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on an mbarrier that lives in another CTA of the cluster (address from mapa).  Default semantics
// (.release.cta): a .release.cluster here compiles to MEMBAR.ALL.GPU + ERRBAR per arrival, which serialised the
// producer of the non-leader CTA behind its own in-flight TMA loads (ncu source page, round 2) - the data hand-off
// itself is ordered by complete_tx (TMA) and by tcgen05.fence (TMEM), not by this arrive.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
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
// TMA load of one box into THIS CTA's shared memory; the transaction bytes are counted on `bar_cluster_addr`,
// which may be the peer CTA's mbarrier (cta_group::2 form)
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] . B[smem]^T over a CTA pair (issued by the leader CTA only)
__device__ __forceinline__ void umma_f16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// once all previously issued tcgen05.mma have completed, arrive on the mbarrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* w) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
        "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]),
        "=r"(w[16]), "=r"(w[17]), "=r"(w[18]), "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]),
        "=r"(w[24]), "=r"(w[25]), "=r"(w[26]), "=r"(w[27]), "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31])
      : "r"(taddr));
}

// Shared-memory matrix descriptor, SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout):
//  [0,14) start address >> 4, [16,30) LBO >> 4, [32,46) SBO >> 4, [46,48) version = 1 (sm_100), [61,64) layout = 2.
//  K-major : 8-row groups are 1024 B apart (SBO); LBO unused.
//  MN-major: a [64 (MN) x BLOCK_K] TMA box is 8-row (K) groups 1024 B apart (SBO); the next 64 MN elements are the
//            next box, BLOCK_K * 128 B further (LBO).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) at [4,6), a/b format BF16 (1) at
// [7,10)/[10,13), a/b major (0 K, 1 MN) at 15/16, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float u = k0 * fmaf(k1 * x * x, x, x);
  return 0.5f * x * (1.f + tanh_approx(u));
}
// erf(x / sqrt2) as tanh of an odd polynomial: tanh(x (c1 + c3 x^2 + c5 x^4)), coefficients fitted (minimax on [0, 6]) to
// |err| < 3.7e-5 - below MUFU.TANH's own 2^-11 and two orders below bf16 resolution.  One MUFU and 5 FP32 instructions
// instead of the 2 MUFU + 14 of a rational erf: the epilogue of a K=1024 tile is issue-bound (8 warps do the math of
// 32768 outputs while the tensor pipe needs only 4.4 us per tile), so the instruction count is what matters.
// x^2 is clamped at 36: beyond it the result is +-1 anyway and the negative x^4 term must not take over.
__device__ __forceinline__ float erf_rsqrt2(float x) {
  const float t = fminf(x * x, 36.f);
  return tanh_approx(x * fmaf(t, fmaf(t, -3.15807060e-4f, 3.67982560e-2f), 7.97717834e-1f));
}
__device__ __forceinline__ float gelu_erf(float x) {
  const float hx = 0.5f * x;
  return fmaf(hx, erf_rsqrt2(x), hx);
}
// GELU and its derivative in one go (they share the tanh; the erf form adds one MUFU.EX2 for the density)
__device__ __forceinline__ void gelu_both(float x, int kind, float& y, float& d) {
  if (kind == 1) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    const float t = tanh_approx(k0 * fmaf(k1 * x * x, x, x));
    const float c = fmaf(0.5f, t, 0.5f);
    y = x * c;
    d = fmaf(0.5f * x * (1.f - t * t), k0 * fmaf(3.f * k1 * x, x, 1.f), c);
  } else {
    const float c = fmaf(0.5f, erf_rsqrt2(x), 0.5f);
    y = x * c;
    d = fmaf(x * __expf(-0.5f * x * x), 0.3989422804014327f, c);
  }
}
// d/dz GELU(z)
__device__ __forceinline__ float dgelu(float z, int kind) {
  if (kind == 3) return z;       // the forward pass stored the derivative itself (has_pre == 2)
  if (kind == 1) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    const float t = tanh_approx(k0 * fmaf(k1 * z * z, z, z));
    return 0.5f * (1.f + t) + 0.5f * z * (1.f - t * t) * k0 * fmaf(3.f * k1 * z, z, 1.f);
  }
  // Phi(z) + z phi(z): the cdf from the tanh form above, the density from one MUFU.EX2
  const float cdf = fmaf(0.5f, erf_rsqrt2(z), 0.5f);
  const float e = __expf(-0.5f * z * z);
  return fmaf(z * e, 0.3989422804014327f, cdf);
}
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

struct Unit { int m_blk, n_blk, kb0, kb1, slice; };

// static schedule shared by all roles: unit u -> (k slice, tile), tiles in a grouped raster
__device__ __forceinline__ Unit decode_unit(int u, int m_tiles, int n_tiles, int k_blocks, const GemmParams& p) {
  Unit w;
  const int tiles = m_tiles * n_tiles;
  w.slice = u / tiles;
  const int t = u - w.slice * tiles;
  const int per_group = kGroupM * n_tiles;
  const int g = t / per_group, r = t - g * per_group;
  const int first_m = g * kGroupM;
  const int gsz = min(kGroupM, m_tiles - first_m);
  w.n_blk = r / gsz;
  w.m_blk = first_m + (r - w.n_blk * gsz);
  w.kb0 = w.slice * p.kb_per_split;
  w.kb1 = min(k_blocks, w.kb0 + p.kb_per_split);
  return w;
}

template <int BLOCK_N, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a1,
            const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b,
            const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_out2, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // identical layout in both CTAs of the pair (the MMA addresses both through one descriptor)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using C = Cfg<BLOCK_N>;
  const int kStages = p.stages;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * C::kABytes;
  uint8_t* staging = smem + kStages * C::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kEpiWarps * p.staging_bufs * kStageBufBytes);
  uint64_t* full = bars;                       // [kMaxStages] used in the leader CTA: 2 arrivals + both CTAs' TMA bytes
  uint64_t* empty = bars + kMaxStages;         // [kMaxStages] one per CTA, released by the multicast commit
  uint64_t* tmem_full = bars + 2 * kMaxStages; // [2]          one per CTA, multicast commit
  uint64_t* tmem_empty = tmem_full + 2;        // [2]          used in the leader CTA: 16 epilogue warps of the pair
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int m_tiles = (p.M + PAIR_M - 1) / PAIR_M, n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int num_units = m_tiles * n_tiles * p.splits;
  constexpr uint32_t kTmemCols = 2 * BLOCK_N;  // power of two: 256 or 512

  cluster_sync();   // both CTAs are running before the pair-wide TMEM allocation
  if (warp == 0 && elect_one()) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 2);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_smem)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = cluster_id; u < num_units; u += num_clusters) {
        const Unit w = decode_unit(u, m_tiles, n_tiles, k_blocks, p);
        int a_row = w.m_blk * PAIR_M + (int)cta_rank * CTA_M;
        const int b_row = w.n_blk * BLOCK_N + (int)cta_rank * C::kBRows;
        // split A operand: [dQ | dK | dV] along K for dgrad, [dQ; dK; dV] along M for wgrad - three buffers, no concat
        const CUtensorMap* ma = &map_a;
        if (p.a_parts > 1 && p.a_mn) {
          const int part = a_row / p.a_part_len;
          a_row -= part * p.a_part_len;
          ma = part == 0 ? &map_a : (part == 1 ? &map_a1 : &map_a2);
        }
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          const uint32_t bar = mapa(smem_u32(&full[stage]), 0);
          const uint32_t sa = smem_u32(smem_a + stage * C::kABytes), sb = smem_u32(smem_b + stage * C::kBBytes);
          if (p.dbg & 1) {
            if (leader) mbar_arrive(&full[stage]);
            else mbar_arrive_cluster(bar);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
            continue;
          }
          if (!p.a_mn) {
            int k0 = kb * BLOCK_K;
            if (p.a_parts > 1) {
              const int part = k0 / p.a_part_len;
              k0 -= part * p.a_part_len;
              ma = part == 0 ? &map_a : (part == 1 ? &map_a1 : &map_a2);
            }
            tma_load_2d(sa, ma, bar, k0, a_row);
          } else {
#pragma unroll
            for (int i = 0; i < CTA_M / 64; ++i) tma_load_2d(sa + i * (BLOCK_K * 128), ma, bar, a_row + i * 64, kb * BLOCK_K);
          }
          if (!p.b_mn) {
            tma_load_2d(sb, &map_b, bar, kb * BLOCK_K, b_row);
          } else {
#pragma unroll
            for (int i = 0; i < C::kBRows / 64; ++i) tma_load_2d(sb + i * (BLOCK_K * 128), &map_b, bar, b_row + i * 64, kb * BLOCK_K);
          }
          if (leader) mbar_arrive_expect_tx(&full[stage], 2 * C::kStageBytes);
          else mbar_arrive_cluster(bar);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA) =====
    if (leader) {
      const uint32_t idesc = make_idesc(PAIR_M, BLOCK_N, p.a_mn, p.b_mn);
      // per UMMA_K step the start address advances by 32 B inside the swizzle row (K-major) or by 16 K-rows (MN-major)
      const uint32_t a_kstep = p.a_mn ? UMMA_K * 128 : UMMA_K * 2, b_kstep = p.b_mn ? UMMA_K * 128 : UMMA_K * 2;
      const uint32_t a_lbo = p.a_mn ? BLOCK_K * 128 : 16, b_lbo = p.b_mn ? BLOCK_K * 128 : 16;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0, n_done = 0;
      uint32_t acc_phase = 0;
      for (int u = cluster_id; u < num_units; u += num_clusters) {
        const Unit w = decode_unit(u, m_tiles, n_tiles, k_blocks, p);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);   // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&full[stage], phase);              // both CTAs' TMA bytes have landed
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_u32(smem_a + stage * C::kABytes);
            const uint32_t b_addr = smem_u32(smem_b + stage * C::kBBytes);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint64_t da = make_smem_desc(a_addr + k * a_kstep, a_lbo);
              const uint64_t db = make_smem_desc(b_addr + k * b_kstep, b_lbo);
              if (!(p.dbg & 2)) umma_f16_2cta(tmem_d, da, db, idesc, (kb > w.kb0 || k > 0) ? 1u : 0u);
            }
            umma_commit_2cta(&empty[stage]);           // frees this smem slot in both CTAs when the MMAs retire
            if (kb == w.kb1 - 1) umma_commit_2cta(&tmem_full[acc]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        ++n_done;
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      // the peer CTA arrives remotely on tmem_empty: wait for the last tile's arrivals before tearing down
      if (n_done > 0) {
        const int last = n_done - 1;
        mbar_wait(&tmem_empty[last & 1], (uint32_t)((last >> 1) & 1));
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===== epilogue (both CTAs): TMEM -> registers -> math -> swizzled smem -> TMA store =====
    const int e = warp - kEpiWarp0;
    const int q = e & 3;                             // == warp % 4: TMEM lanes [32q, 32q+32)
    const int half = e >> 2;                         // which half of the tile's columns this warp drains
    constexpr int kColsPerWarp = BLOCK_N / 2;
    constexpr int kChunk = (EPI == EPI_ACC_F32) ? 32 : 64;   // columns per staged box (128-byte rows)
    const int nbufs = p.staging_bufs;
    uint8_t* stage_buf = staging + e * (nbufs * kStageBufBytes);
    int sbuf = 0;
    const uint32_t tmem_empty_leader = mapa(smem_u32(&tmem_empty[0]), 0);

    // registers -> staging buffer (thread = row; 16-byte chunk j of row r lives at chunk j ^ (r & 7), the layout a
    // SWIZZLE_128B tensor map expects) -> one TMA store of the 32-row box
    auto store_box = [&](const uint4* o, const CUtensorMap* map, int col, int row) {
      if (p.dbg & 4) return;
      uint8_t* buf = stage_buf + sbuf * kStageBufBytes;
      if (lane == 0) {   // the store that used `buf` last has finished reading it
        if (nbufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) = o[j];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(map, smem_u32(buf), col, row);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      sbuf = (nbufs == 2) ? (sbuf ^ 1) : 0;
    };
    auto store_bf16 = [&](const float* f, const CUtensorMap* map, int col, int row) {
      uint4 o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j].x = pack_bf16(f[8 * j + 0], f[8 * j + 1]); o[j].y = pack_bf16(f[8 * j + 2], f[8 * j + 3]);
        o[j].z = pack_bf16(f[8 * j + 4], f[8 * j + 5]); o[j].w = pack_bf16(f[8 * j + 6], f[8 * j + 7]);
      }
      store_box(o, map, col, row);
    };

    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = cluster_id; u < num_units; u += num_clusters) {
      const Unit w = decode_unit(u, m_tiles, n_tiles, k_blocks, p);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row_in_out = w.m_blk * PAIR_M + (int)cta_rank * CTA_M + q * 32;       // first row of this warp's box
      const int row0 = (p.splits > 1 ? w.slice * p.m_pad : 0) + row_in_out;           // row in the output tensor map
      const int my_row = row_in_out + lane;
#pragma unroll 1
      for (int c0 = half * kColsPerWarp; c0 < (half + 1) * kColsPerWarp; c0 += kChunk) {
        const int col0 = w.n_blk * BLOCK_N + c0;
        const bool last_chunk = c0 + kChunk >= (half + 1) * kColsPerWarp;
        uint32_t v[kChunk];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c0);
        // operand of the epilogue read from global memory (issued before the TMEM load so the latencies overlap)
        uint4 g[8];
        if constexpr (EPI == EPI_DGELU) {
          const __nv_bfloat16* src = static_cast<const __nv_bfloat16*>(p.aux) + (size_t)my_row * p.ld_aux + col0;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            g[j] = (my_row < p.M && col0 + 8 * j < p.N) ? __ldg(reinterpret_cast<const uint4*>(src) + j) : make_uint4(0, 0, 0, 0);
        } else if constexpr (EPI == EPI_ACC_BF16) {
          if (p.accumulate) {
            const __nv_bfloat16* src = static_cast<const __nv_bfloat16*>(p.aux) + (size_t)my_row * p.ld_aux + col0;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              g[j] = (my_row < p.M && col0 + 8 * j < p.N) ? *(reinterpret_cast<const uint4*>(src) + j) : make_uint4(0, 0, 0, 0);
          }
        } else if constexpr (EPI == EPI_ACC_F32) {
          if (p.accumulate) {
            const float* src = static_cast<const float*>(p.aux) + (size_t)my_row * p.ld_aux + col0;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              g[j] = (my_row < p.M && col0 + 4 * j < p.N) ? *(reinterpret_cast<const uint4*>(src) + j) : make_uint4(0, 0, 0, 0);
          }
        }
        tmem_ld_x32(taddr, v);
        if constexpr (kChunk == 64) tmem_ld_x32(taddr + 32, v + 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (last_chunk) {
          // this warp's part of the accumulator is in registers: hand the TMEM buffer back to the MMA issuer early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(&tmem_empty[acc]);
            else mbar_arrive_cluster(tmem_empty_leader + (uint32_t)(acc * 8));
          }
        }
        float f[kChunk];
#pragma unroll
        for (int i = 0; i < kChunk; ++i) f[i] = __uint_as_float(v[i]);

        if constexpr (EPI == EPI_LINEAR || EPI == EPI_STATS) {
          if (p.bias && p.bias_bf16) {
            const __nv_bfloat16* bp = static_cast<const __nv_bfloat16*>(p.bias) + col0;
#pragma unroll
            for (int j = 0; j < kChunk; j += 8) {
              if (col0 + j < p.N) {
                const uint4 bv = __ldg(reinterpret_cast<const uint4*>(bp + j));
                f[j] += bf16lo(bv.x); f[j + 1] += bf16hi(bv.x); f[j + 2] += bf16lo(bv.y); f[j + 3] += bf16hi(bv.y);
                f[j + 4] += bf16lo(bv.z); f[j + 5] += bf16hi(bv.z); f[j + 6] += bf16lo(bv.w); f[j + 7] += bf16hi(bv.w);
              }
            }
          } else if (p.bias) {
            const float* bp = static_cast<const float*>(p.bias) + col0;
#pragma unroll
            for (int j = 0; j < kChunk; j += 4) {
              if (col0 + j < p.N) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(bp + j));
                f[j] += bv.x; f[j + 1] += bv.y; f[j + 2] += bv.z; f[j + 3] += bv.w;
              }
            }
          }
        }
        if constexpr (EPI == EPI_LINEAR) {
          if (p.has_pre == 2 && p.act != 0) {
            // second output = act'(z): the backward dgrad epilogue then only multiplies (EPI_DGELU with act 3)
            uint4 oy[8], od[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float y[8], d[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) gelu_both(f[8 * j + e], p.act, y[e], d[e]);
              oy[j] = make_uint4(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
              od[j] = make_uint4(pack_bf16(d[0], d[1]), pack_bf16(d[2], d[3]), pack_bf16(d[4], d[5]), pack_bf16(d[6], d[7]));
            }
            store_box(od, &map_out2, col0, row0);
            store_box(oy, &map_out, col0, row0);
            continue;
          }
          if (p.has_pre) store_bf16(f, &map_out2, col0, row0);
          if (p.act == 1) {
#pragma unroll
            for (int i = 0; i < kChunk; ++i) f[i] = gelu_tanh(f[i]);
          } else if (p.act == 2) {
#pragma unroll
            for (int i = 0; i < kChunk; ++i) f[i] = gelu_erf(f[i]);
          }
          store_bf16(f, &map_out, col0, row0);
        } else if constexpr (EPI == EPI_DGELU) {
          const uint32_t* gw = reinterpret_cast<const uint32_t*>(g);
#pragma unroll
          for (int i = 0; i < kChunk / 2; ++i) {
            f[2 * i] *= dgelu(bf16lo(gw[i]), p.act);
            f[2 * i + 1] *= dgelu(bf16hi(gw[i]), p.act);
          }
          store_bf16(f, &map_out, col0, row0);
        } else if constexpr (EPI == EPI_ACC_BF16) {
          if (p.accumulate) {
            const uint32_t* gw = reinterpret_cast<const uint32_t*>(g);
#pragma unroll
            for (int i = 0; i < kChunk / 2; ++i) {
              f[2 * i] += bf16lo(gw[i]);
              f[2 * i + 1] += bf16hi(gw[i]);
            }
          }
          store_bf16(f, &map_out, col0, row0);
        } else if constexpr (EPI == EPI_ACC_F32) {
          if (p.accumulate) {
            const float* gf = reinterpret_cast<const float*>(g);
#pragma unroll
            for (int i = 0; i < kChunk; ++i) f[i] += gf[i];
          }
          uint4 o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            o[j] = make_uint4(__float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]), __float_as_uint(f[4 * j + 2]), __float_as_uint(f[4 * j + 3]));
          store_box(o, &map_out, col0, row0);
        } else if constexpr (EPI == EPI_STATS) {
          // round to bf16 first: the statistics describe exactly the tensor that BatchNorm will read back
          uint4 o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            o[j].x = pack_bf16(f[8 * j + 0], f[8 * j + 1]); o[j].y = pack_bf16(f[8 * j + 2], f[8 * j + 3]);
            o[j].z = pack_bf16(f[8 * j + 4], f[8 * j + 5]); o[j].w = pack_bf16(f[8 * j + 6], f[8 * j + 7]);
          }
          uint8_t* buf = stage_buf + sbuf * kStageBufBytes;
          store_box(o, &map_out, col0, row0);
          // column sums over this warp's 32 rows, read back from the staging buffer (a TMA store only reads it):
          // lane l owns columns 2l, 2l+1 = word (l & 3) of 16-byte chunk (l >> 2); conflict-free under the swizzle
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
          const int rows_valid = min(32, p.M - row_in_out);
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            const uint32_t wv = *reinterpret_cast<const uint32_t*>(buf + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
            if (r < rows_valid) {
              const float a = bf16lo(wv), b = bf16hi(wv);
              s0 += a; s1 += b;
              q0 = fmaf(a, a, q0); q1 = fmaf(b, b, q1);
            }
          }
          const int c = col0 + 2 * lane;
          if (c < p.N) {
            const int srow = (w.m_blk * 2 + (int)cta_rank) * 4 + q;      // index of this 32-row group
            float* dst = p.stats + (size_t)srow * p.N + c;
            *reinterpret_cast<float2*>(dst) = make_float2(s0, s1);
            *reinterpret_cast<float2*>(dst + (size_t)p.stat_rows * p.N) = make_float2(q0, q1);
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores of this warp are complete
  }

  tc_fence_before();
  cluster_sync();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---------------- split-K second pass: dst (+)= sum over slices of the fp32 partial tiles ---------------------------
template <typename OutT>
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, int splits, long long slice_stride, int M, int N, OutT* __restrict__ dst,
                     long long ld_dst, int accumulate) {
  const long long nvec = (long long)M * (N / 4);
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(v / (N / 4)), c = (int)(v % (N / 4)) * 4;
    float4 a = *reinterpret_cast<const float4*>(ws + (size_t)r * N + c);
    for (int s = 1; s < splits; ++s) {   // fixed order: deterministic
      const float4 b = *reinterpret_cast<const float4*>(ws + (size_t)s * slice_stride + (size_t)r * N + c);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    OutT* d = dst + (size_t)r * ld_dst + c;
    if constexpr (sizeof(OutT) == 4) {
      if (accumulate) {
        const float4 o = *reinterpret_cast<const float4*>(d);
        a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
      }
      *reinterpret_cast<float4*>(d) = a;
    } else {
      if (accumulate) {
        const uint2 o = *reinterpret_cast<const uint2*>(d);
        a.x += bf16lo(o.x); a.y += bf16hi(o.x); a.z += bf16lo(o.y); a.w += bf16hi(o.y);
      }
      *reinterpret_cast<uint2*>(d) = make_uint2(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w));
    }
  }
}

// ---------------- BatchNorm statistics: deterministic tree over the per-32-row partials of EPI_STATS ---------------
// out[0][c] = sum_r part[0][r][c], out[1][c] = sum_r part[1][r][c]   (one block per 32 channels, fixed order)
__global__ void __launch_bounds__(256)
stats_reduce_kernel(const float* __restrict__ part, int rows, int N, float* __restrict__ out) {
  __shared__ float sh[8][33];
  const int which = blockIdx.y;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
  const float* src = part + (size_t)which * rows * N;
  float acc = 0.f;
  if (c < N)
    for (int r = ty; r < rows; r += 8) acc += src[(size_t)r * N + c];
  sh[ty][threadIdx.x & 31] = acc;
  __syncthreads();
  if (ty == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
    out[(size_t)which * N + c] = t;
  }
}

// dX_pre = dY * gelu'(pre): stand-alone version of the EPI_DGELU math (used when dY does not come out of a GEMM)
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ pre, __nv_bfloat16* __restrict__ out,
                long long nvec, int kind) {
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    const uint4 a = reinterpret_cast<const uint4*>(dy)[v], b = reinterpret_cast<const uint4*>(pre)[v];
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    uint32_t ow[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      ow[i] = pack_bf16(bf16lo(aw[i]) * dgelu(bf16lo(bw[i]), kind), bf16hi(aw[i]) * dgelu(bf16hi(bw[i]), kind));
    reinterpret_cast<uint4*>(out)[v] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

// ---------------- column sums (bias gradients): dst[n] (+)= sum_m x[m, n] ------------------------------------------
// Block = 8 column vectors (64 columns) x 32 rows; grid = (N / 64, R row slices).  Every block writes one partial
// row; the block that finishes last for its column group adds the R partials in a fixed order (deterministic),
// writes / accumulates the result and re-arms the ticket counter.
template <typename OutT>
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, int M, int N, long long ld, OutT* __restrict__ dst, int accumulate,
              float* __restrict__ partial, unsigned int* __restrict__ tickets) {
  __shared__ float sh[32][65];
  __shared__ unsigned int s_ticket;
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const int col = blockIdx.x * 64 + tx * 8;
  const int R = gridDim.y;
  const int rows_per = (M + R - 1) / R;
  const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (col < N) {
    int r = r0 + ty;
    for (; r + 96 < r1; r += 128) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(x + (size_t)(r + 32 * u) * ld + col));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc[0] += bf16lo(v[u].x); acc[1] += bf16hi(v[u].x); acc[2] += bf16lo(v[u].y); acc[3] += bf16hi(v[u].y);
        acc[4] += bf16lo(v[u].z); acc[5] += bf16hi(v[u].z); acc[6] += bf16lo(v[u].w); acc[7] += bf16hi(v[u].w);
      }
    }
    for (; r < r1; r += 32) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (size_t)r * ld + col));
      acc[0] += bf16lo(v.x); acc[1] += bf16hi(v.x); acc[2] += bf16lo(v.y); acc[3] += bf16hi(v.y);
      acc[4] += bf16lo(v.z); acc[5] += bf16hi(v.z); acc[6] += bf16lo(v.w); acc[7] += bf16hi(v.w);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sh[ty][tx * 8 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 32; ++y) t += sh[y][threadIdx.x];
    partial[(size_t)blockIdx.y * (gridDim.x * 64) + blockIdx.x * 64 + threadIdx.x] = t;   // pitch: N rounded up to 64
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&tickets[blockIdx.x], 1u);
  __syncthreads();
  if (s_ticket != (unsigned int)(R - 1)) return;
  __threadfence();
  if (threadIdx.x < 64) {
    const int c = blockIdx.x * 64 + threadIdx.x;
    float t = 0.f;
    for (int y = 0; y < R; ++y) t += partial[(size_t)y * (gridDim.x * 64) + c];
    if (c < N) {
      if constexpr (sizeof(OutT) == 4) {
        dst[c] = accumulate ? dst[c] + t : t;
      } else {
        dst[c] = __float2bfloat16(accumulate ? __bfloat162float(dst[c]) + t : t);
      }
    }
  }
  if (threadIdx.x == 0) tickets[blockIdx.x] = 0u;
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

// row-major [rows, cols] matrix with row pitch ld (elements), box = [box_rows, box_cols], 128-byte swizzle
// (box_cols * elem_bytes must be 128)
static bool make_map(CUtensorMap* map, const void* ptr, CUtensorMapDataType dt, int elem_bytes, uint64_t rows, uint64_t cols,
                     uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    fprintf(stderr, "sparkdl gemm: cuTensorMapEncodeTiled failed (%d): ptr %p rows %llu cols %llu ld %llu box %ux%u elem %d\n", (int)r, ptr,
            (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols, elem_bytes);
  return r == CUDA_SUCCESS;
}

template <int BLOCK_N, int EPI>
static cudaError_t launch(int grid, cudaStream_t s, const CUtensorMap* ma, const CUtensorMap& mb, const CUtensorMap& mo,
                          const CUtensorMap& mo2, const GemmParams& p) {
  const int smem = Cfg<BLOCK_N>::smem_bytes(p.stages, p.staging_bufs);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BLOCK_N, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  gemm_kernel<BLOCK_N, EPI><<<grid, kNumThreads, smem, s>>>(ma[0], ma[1], ma[2], mb, mo, mo2, p);
  return cudaGetLastError();
}

template <int EPI>
static cudaError_t launch_bn(int block_n, int grid, cudaStream_t s, const CUtensorMap* ma, const CUtensorMap& mb,
                             const CUtensorMap& mo, const CUtensorMap& mo2, const GemmParams& p) {
  return block_n == 256 ? launch<256, EPI>(grid, s, ma, mb, mo, mo2, p) : launch<128, EPI>(grid, s, ma, mb, mo, mo2, p);
}

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms <= 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return g_num_sms;
}

}  // namespace gemm

using namespace gemm;

extern "C" {

// General entry point.  D[M,N] = epi(A . B^T), bf16 operands.
//   a / lda, a_mn : A is [M, lda] (K-major, a_mn = 0) or [K, lda] (MN-major, a_mn = 1); likewise b for [N,K] / [K,N]
//   epi           : EPI_* above;  out / ld_out: destination (bf16, or fp32 for EPI_ACC_F32)
//   out2          : EPI_LINEAR: optional pre-activation output (bf16, same pitch as out)
//   bias, act     : EPI_LINEAR (act | 16: bias is bf16 instead of fp32);   aux / ld_aux, act: EPI_DGELU (bf16 pre-activation)
//   accumulate    : EPI_ACC_*: out += result (out is read through aux := out)
//   splits        : > 1 = split-K (EPI_ACC_* only); `workspace` must hold splits * ceil(M/256)*256 * N floats
//   stats         : EPI_STATS: fp32 [2][ceil(M/256)*8][N] partials followed by (caller-sized) [2][N] totals; the totals
//                   are produced by a second tiny kernel
//   block_n       : 256 / 128 / 0 = choose
// Requirements: 16-byte aligned pointers and pitches (lda, ldb, ld_out multiples of 8 elements; 4 for fp32 out).
// Returns 0 on success, negative on error.
int sdlc_gemm_ex3(const void* a, const void* a1, const void* a2, int a_part_len, long long lda, int a_mn, const void* b,
                  long long ldb, int b_mn, void* out, long long ld_out,
                 void* out2, int M, int N, int K, int epi, const void* bias, int act, const void* aux, long long ld_aux,
                 int accumulate, int splits, float* workspace, float* stats, int block_n, int sms, unsigned long long stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((lda % 8) != 0 || (ldb % 8) != 0 || (N % 8) != 0) return -2;
  if (epi < 0 || epi > EPI_STATS) return -2;
  if (!load_encode()) return -3;
  {
    // cuTensorMapEncodeTiled is a driver-API call: it needs the primary context bound to THIS thread.  autograd runs
    // backward on its own worker thread, and if our dgrad is the first CUDA work there nothing has bound it yet
    // (CUDA_ERROR_INVALID_CONTEXT, seen in round 2).  cudaSetDevice binds it; once per thread.
    static thread_local bool bound = false;
    if (!bound) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaSetDevice(dev);
      bound = true;
    }
  }
  if (sms <= 0) sms = num_sms();
  const int max_clusters = sms / 2 > 0 ? sms / 2 : 1;
  const int m_tiles = (M + PAIR_M - 1) / PAIR_M;
  const int k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  if (block_n != 256 && block_n != 128) {
    // 256-wide tiles need 2/3 of the L2->smem bytes per FLOP of 128-wide ones and measure ~1.3x faster per tile-FLOP
    // (gemm_lab, round 2); the narrow tile only pays when the wide one leaves most clusters idle in its last wave
    const int t256 = m_tiles * ((N + 255) / 256), t128 = m_tiles * ((N + 127) / 128);
    auto eff = [&](int t) { const int waves = (t + max_clusters - 1) / max_clusters; return (double)t / (waves * max_clusters); };
    block_n = (N > 128 && eff(t256) >= 0.72 * eff(t128)) ? 256 : 128;
  }
  const int n_tiles = (N + block_n - 1) / block_n;
  const bool acc_epi = epi == EPI_ACC_BF16 || epi == EPI_ACC_F32;
  if (splits < 1 || !acc_epi || !workspace) splits = 1;
  if (splits > k_blocks) splits = k_blocks;
  int kbps = (k_blocks + splits - 1) / splits;
  splits = (k_blocks + kbps - 1) / kbps;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K; p.a_mn = a_mn ? 1 : 0; p.b_mn = b_mn ? 1 : 0;
  p.a_parts = a1 ? (a2 ? 3 : 2) : 1;
  p.a_part_len = a_part_len;
  if (p.a_parts > 1) {
    const int split_dim = a_mn ? M : K;
    if (a_part_len <= 0 || a_part_len * p.a_parts != split_dim || (a_part_len % (a_mn ? PAIR_M : BLOCK_K)) != 0) return -2;
  }
  p.splits = splits; p.kb_per_split = kbps; p.m_pad = m_tiles * PAIR_M;
  p.act = act; p.has_pre = (epi == EPI_LINEAR && out2) ? ((act & 32) ? 2 : 1) : 0;
  p.bias = bias; p.bias_bf16 = (act & 16) ? 1 : 0; p.act = act & 15; p.aux = aux; p.ld_aux = ld_aux;
  p.accumulate = (accumulate && splits == 1) ? 1 : 0;
  if (p.accumulate && acc_epi) { p.aux = out; p.ld_aux = ld_out; }
  p.stats = stats; p.stat_rows = m_tiles * 8;
  static int dbg = -1, staging_env = -1;
  if (dbg < 0) { const char* e = getenv("SPARKDL_GEMM_DBG"); dbg = e ? atoi(e) : 0; }
  if (staging_env < 0) { const char* e = getenv("SPARKDL_GEMM_STAGING"); staging_env = e ? atoi(e) : 0; }
  p.dbg = dbg;
  // a deeper operand ring beats double-buffered epilogue staging when the K loop is long; EPI_STATS reads its
  // staging buffer back after the store was issued, so it keeps two
  p.staging_bufs = (staging_env == 1 || staging_env == 2) ? staging_env : 1;
  if (epi == EPI_STATS) p.staging_bufs = 2;
  p.stages = block_n == 256 ? Cfg<256>::max_stages(p.staging_bufs) : Cfg<128>::max_stages(p.staging_bufs);
  if (epi == EPI_STATS && !stats) return -2;
  if (epi == EPI_DGELU && (!aux || (ld_aux % 8) != 0)) return -2;

  CUtensorMap ma[3], mb, mo, mo2;
  const CUtensorMapDataType BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, F32 = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  bool ok = true;
  const void* aptr[3] = {a, a1, a2};
  for (int i = 0; i < p.a_parts; ++i) {
    // every part is a matrix of its own: K-major [M, K / parts], MN-major [K, M / parts]
    if (!a_mn) ok &= make_map(&ma[i], aptr[i], BF, 2, (uint64_t)M, (uint64_t)(K / p.a_parts), (uint64_t)lda, CTA_M, BLOCK_K);
    else ok &= make_map(&ma[i], aptr[i], BF, 2, (uint64_t)K, (uint64_t)(M / p.a_parts), (uint64_t)lda, BLOCK_K, 64);
  }
  for (int i = p.a_parts; i < 3; ++i) ma[i] = ma[0];
  if (!b_mn) ok &= make_map(&mb, b, BF, 2, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, (uint32_t)block_n / 2, BLOCK_K);
  else ok &= make_map(&mb, b, BF, 2, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BLOCK_K, 64);
  int run_epi = epi;
  if (splits > 1) {
    run_epi = EPI_ACC_F32;   // partial tiles go to the fp32 workspace, the second pass converts / accumulates
    ok &= make_map(&mo, workspace, F32, 4, (uint64_t)splits * p.m_pad, (uint64_t)N, (uint64_t)N, 32, 32);
  } else if (epi == EPI_ACC_F32) {
    if ((ld_out % 4) != 0) return -2;
    ok &= make_map(&mo, out, F32, 4, (uint64_t)M, (uint64_t)N, (uint64_t)ld_out, 32, 32);
  } else {
    if ((ld_out % 8) != 0) return -2;
    ok &= make_map(&mo, out, BF, 2, (uint64_t)M, (uint64_t)N, (uint64_t)ld_out, 32, 64);
  }
  if (p.has_pre) ok &= make_map(&mo2, out2, BF, 2, (uint64_t)M, (uint64_t)N, (uint64_t)ld_out, 32, 64);
  else mo2 = mo;
  if (!ok) return -4;

  const int units = m_tiles * n_tiles * splits;
  const int clusters = units < max_clusters ? units : max_clusters;
  const int grid = 2 * clusters;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e;
  switch (run_epi) {
    case EPI_LINEAR: e = launch_bn<EPI_LINEAR>(block_n, grid, s, ma, mb, mo, mo2, p); break;
    case EPI_DGELU: e = launch_bn<EPI_DGELU>(block_n, grid, s, ma, mb, mo, mo2, p); break;
    case EPI_ACC_BF16: e = launch_bn<EPI_ACC_BF16>(block_n, grid, s, ma, mb, mo, mo2, p); break;
    case EPI_ACC_F32: e = launch_bn<EPI_ACC_F32>(block_n, grid, s, ma, mb, mo, mo2, p); break;
    default: e = launch_bn<EPI_STATS>(block_n, grid, s, ma, mb, mo, mo2, p); break;
  }
  if (e != cudaSuccess) return -1;
  if (splits > 1) {
    const long long nvec = (long long)M * (N / 4);
    long long blocks = (nvec + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    const long long slice_stride = (long long)p.m_pad * N;
    if (epi == EPI_ACC_F32)
      splitk_reduce_kernel<float><<<(int)blocks, 256, 0, s>>>(workspace, splits, slice_stride, M, N, static_cast<float*>(out), ld_out, accumulate);
    else
      splitk_reduce_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, s>>>(workspace, splits, slice_stride, M, N, static_cast<__nv_bfloat16*>(out), ld_out, accumulate);
    if (cudaGetLastError() != cudaSuccess) return -1;
  }
  if (epi == EPI_STATS) {
    float* totals = stats + (size_t)2 * p.stat_rows * N;
    stats_reduce_kernel<<<dim3((N + 31) / 32, 2), 256, 0, s>>>(stats, p.stat_rows, N, totals);
    if (cudaGetLastError() != cudaSuccess) return -1;
  }
  return 0;
}

int sdlc_gemm_ex(const void* a, long long lda, int a_mn, const void* b, long long ldb, int b_mn, void* out, long long ld_out,
                 void* out2, int M, int N, int K, int epi, const void* bias, int act, const void* aux, long long ld_aux,
                 int accumulate, int splits, float* workspace, float* stats, int block_n, int sms, unsigned long long stream) {
  return sdlc_gemm_ex3(a, nullptr, nullptr, 0, lda, a_mn, b, ldb, b_mn, out, ld_out, out2, M, N, K, epi, bias, act, aux, ld_aux,
                       accumulate, splits, workspace, stats, block_n, sms, stream);
}

// number of kernels sdlc_gemm_ex launches for these arguments (the Python launch counter asks)
int sdlc_gemm_stat_rows(int M) { return ((M + PAIR_M - 1) / PAIR_M) * 8; }

// Y[M,N] = act(X[M,K] . W[N,K]^T + bias)  (kept for callers of the round-1 interface)
int sdlc_gemm_bf16(const void* x, const void* w, const float* bias, void* out, void* pre_act, int M, int N, int K,
                   int epilogue, int block_n, int sms, unsigned long long stream) {
  return sdlc_gemm_ex(x, K, 0, w, K, 0, out, N, pre_act, M, N, K, EPI_LINEAR, bias, epilogue, nullptr, 0, 0, 1, nullptr, nullptr,
                      block_n, sms, stream);
}

// dst[N] (+)= column sums of the bf16 matrix x[M, ld] (bias gradient of a linear layer).  N % 8 == 0, ld % 8 == 0.
// workspace: 64 * round_up(N, 64) floats; tickets: round_up(N, 64) / 64 zero-initialised uint32 (left zero on return)
int sdlc_colsum_bf16(const void* x, int M, int N, long long ld, void* dst, int dst_f32, int accumulate, float* workspace,
                     unsigned int* tickets, unsigned long long stream) {
  if (M <= 0 || N <= 0) return 0;
  if ((N % 8) != 0 || (ld % 8) != 0) return -2;
  const int cb = (N + 63) / 64;
  int R = 592 / cb;
  if (R > 64) R = 64;
  if (R > (M + 127) / 128) R = (M + 127) / 128;
  if (R < 1) R = 1;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dst_f32)
    colsum_kernel<float><<<dim3(cb, R), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), M, N, ld, static_cast<float*>(dst), accumulate, workspace, tickets);
  else
    colsum_kernel<__nv_bfloat16><<<dim3(cb, R), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), M, N, ld, static_cast<__nv_bfloat16*>(dst), accumulate, workspace, tickets);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// out = dy * gelu'(pre); numel % 8 == 0.  kind: 1 tanh, 2 erf
int sdlc_gelu_bwd(const void* dy, const void* pre, void* out, long long numel, int kind, unsigned long long stream) {
  const long long nvec = numel / 8;
  long long blocks = (nvec + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  gelu_bwd_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(pre), static_cast<__nv_bfloat16*>(out), nvec, kind);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // extern "C"
