// Native check of the feature-parallel CPU histogram (gbdt_cpu.cc) against a serial reference: empty input, inactive rows
// (pos = -1), 1..100 features, 1..64 nodes, 1 / 3 / default threads.  Built plain and with sanitizers by tests/test_core_native.py.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
extern "C" int sdl_gbdt_hist_cpu(const uint8_t* bins, const int32_t* gh, const int32_t* pos, int64_t* hist, int64_t n, int F, int n_nodes, int max_threads);
int main() {
  int bad = 0;
  const int cases[][3] = {{0, 1, 1}, {1, 1, 1}, {1000, 1, 3}, {5000, 7, 8}, {4097, 64, 32}, {333, 100, 1}, {2000, 3, 64}};
  for (auto& c : cases) {
    const int64_t n = c[0]; const int F = c[1], nodes = c[2];
    std::vector<uint8_t> bins((size_t)n * F + 1);
    std::vector<int32_t> gh((size_t)n * 2 + 2), pos((size_t)n + 1);
    for (size_t i = 0; i < bins.size(); ++i) bins[i] = (uint8_t)(rand() & 255);
    for (size_t i = 0; i < gh.size(); ++i) gh[i] = (rand() % 2001) - 1000;
    for (int64_t i = 0; i < n; ++i) pos[i] = (rand() % (nodes + 1)) - 1;     // -1 = inactive row
    std::vector<int64_t> hist((size_t)nodes * F * 256 * 2, 0), ref(hist.size(), 0);
    for (int threads : {1, 3, 0}) {
      std::fill(hist.begin(), hist.end(), 0);
      int rc = sdl_gbdt_hist_cpu(bins.data(), gh.data(), pos.data(), hist.data(), n, F, nodes, threads);
      if (rc != 0) { ++bad; continue; }
      std::fill(ref.begin(), ref.end(), 0);
      for (int64_t i = 0; i < n; ++i) if (pos[i] >= 0)
        for (int f = 0; f < F; ++f) {
          size_t k = (((size_t)pos[i] * F + f) * 256 + bins[(size_t)i * F + f]) * 2;
          ref[k] += gh[2 * i]; ref[k + 1] += gh[2 * i + 1];
        }
      if (hist != ref) { ++bad; fprintf(stderr, "mismatch n=%lld F=%d nodes=%d threads=%d\n", (long long)n, F, nodes, threads); }
    }
  }
  printf("gbdt_cpu asan harness: %s\n", bad ? "FAILED" : "OK");
  return bad ? 1 : 0;
}
