// CPU histogram construction for the GBDT engine (the CUDA twin is csrc/cuda/gbdt.cu).
//
// hist[node][feature][bin][{g,h}] += gh[row] for every row with pos[row] = node >= 0.  Integer
// (int64) accumulation makes the result independent of the summation order, so the work is split
// over threads by FEATURE: every thread owns a disjoint slice of the output, no atomics, no
// private copies to merge, bit-identical to the single-threaded result.
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

namespace {

void hist_features(const uint8_t* bins, const int32_t* gh, const int32_t* pos, int64_t* hist, int64_t n, int F,
                   int f0, int f1) {
  constexpr int kBins = 256;
  for (int64_t r = 0; r < n; ++r) {
    const int32_t node = pos[r];
    if (node < 0) continue;
    const int64_t g = gh[2 * r], h = gh[2 * r + 1];
    const uint8_t* row = bins + r * F;
    int64_t* base = hist + (int64_t)node * F * kBins * 2;
    for (int f = f0; f < f1; ++f) {
      int64_t* cell = base + ((int64_t)f * kBins + row[f]) * 2;
      cell[0] += g;
      cell[1] += h;
    }
  }
}

}  // namespace

extern "C" {

// Returns 0 on success, -1 on bad arguments.  `hist` must be zero-initialised by the caller and hold
// n_nodes * F * 256 * 2 int64; pos values must lie in [-1, n_nodes).
int sdl_gbdt_hist_cpu(const uint8_t* bins, const int32_t* gh, const int32_t* pos, int64_t* hist, int64_t n, int F,
                      int n_nodes, int max_threads) {
  if (!hist || n < 0 || F <= 0 || n_nodes <= 0) return -1;
  if (n == 0) return 0;                            // an empty shard (its tensors have null data pointers)
  if (!bins || !gh || !pos) return -1;
  for (int64_t r = 0; r < n; ++r)
    if (pos[r] >= n_nodes) return -1;
  int hw = (int)std::thread::hardware_concurrency();
  if (hw <= 0) hw = 1;
  int T = std::min(std::min(hw, F), max_threads > 0 ? max_threads : 16);
  if (n * (int64_t)F < (1 << 16)) T = 1;          // not worth a thread start
  if (T <= 1) {
    hist_features(bins, gh, pos, hist, n, F, 0, F);
    return 0;
  }
  std::vector<std::thread> th;
  th.reserve(T);
  for (int t = 0; t < T; ++t) {
    const int f0 = (int)((int64_t)F * t / T), f1 = (int)((int64_t)F * (t + 1) / T);
    if (f0 < f1) th.emplace_back(hist_features, bins, gh, pos, hist, n, F, f0, f1);
  }
  for (auto& x : th) x.join();
  return 0;
}

}  // extern "C"
