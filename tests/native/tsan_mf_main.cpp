#include <cstdint>
#include <cstdio>
#include <vector>
#include <random>
extern "C" int fps_host_mf_train(const int32_t*, const int32_t*, const float*, int64_t, int32_t, int32_t, int32_t, float,
                                 int32_t, float, float, uint64_t, int32_t, int32_t, float*, int64_t, float*, int64_t,
                                 uint8_t*, uint8_t*, double*);
int main() {
  std::mt19937 g(1);
  for (int trial = 0; trial < 12; ++trial) {
    int W = 1 + g() % 6, S = 1 + g() % 6, L = 1 + g() % 17, k = 1 + g() % 20;
    int nu = 50 + g() % 200, ni = 20 + g() % 100; int64_t n = 2000 + g() % 20000;
    std::vector<int32_t> u(n), it(n); std::vector<float> r(n);
    for (int64_t i = 0; i < n; ++i) { u[i] = g() % nu; it[i] = g() % ni; r[i] = (g() % 1000) / 1000.f; }
    std::vector<float> ut((size_t)nu * k), vt((size_t)ni * k); std::vector<uint8_t> a(nu), b(ni); double sse = 0;
    int rc = fps_host_mf_train(u.data(), it.data(), r.data(), n, W, S, k, 0.05f, trial & 1, -0.1f, 0.1f, 7, 1 + trial % 3, L,
                               ut.data(), nu, vt.data(), ni, a.data(), b.data(), &sse);
    std::printf("trial %d W=%d S=%d L=%d k=%d n=%lld rc=%d sse=%.3f\n", trial, W, S, L, k, (long long)n, rc, sse);
  }
  return 0;
}
