#include <cstdint>
#include <cstdio>
#include <vector>
#include <random>
extern "C" int fps_host_mf_train(const int32_t*, const int32_t*, const float*, int64_t, int32_t, int32_t, int32_t, float,
                                 int32_t, float, float, uint64_t, int32_t, int32_t, float*, int64_t, float*, int64_t,
                                 uint8_t*, uint8_t*, double*, int32_t, int32_t);
extern "C" int fps_host_pa_binary(const int64_t*, const int32_t*, const float*, const int32_t*, int64_t, int64_t,
                                  int32_t, float, int32_t, int32_t, int32_t, int32_t, float*, int32_t*, uint8_t*);
extern "C" int fps_host_pa_multiclass(const int64_t*, const int32_t*, const float*, const int32_t*, int64_t, int64_t,
                                      int32_t, int32_t, float, const float*, int32_t, int32_t, int32_t, int32_t,
                                      float*, int32_t*, uint8_t*);
static void pa_multi_trials(std::mt19937& g) {
  for (int trial = 0; trial < 6; ++trial) {
    int W = 1 + g() % 5, S = 1 + g() % 5, L = 2 + g() % 9, lim = 1 + g() % 40;
    int64_t feats = 200 + g() % 2000, n = 200 + g() % 1200;
    std::vector<int64_t> rp(1, 0); std::vector<int32_t> cols; std::vector<float> vals; std::vector<int32_t> lab(n);
    for (int64_t e = 0; e < n; ++e) {
      int nnz = 1 + g() % 40;
      for (int j = 0; j < nnz; ++j) { cols.push_back(g() % feats); vals.push_back((int)(g() % 2001 - 1000) / 1000.f); }
      rp.push_back((int64_t)cols.size());
      lab[e] = (int)(g() % (L + 1)) - 1;
    }
    std::vector<float> w((size_t)feats * L, 0.f); std::vector<int32_t> pred(n); std::vector<uint8_t> t(feats);
    int rc = fps_host_pa_multiclass(rp.data(), cols.data(), vals.data(), lab.data(), n, feats, L, trial % 5, 0.1f,
                                    nullptr, W, S, lim, trial & 1, w.data(), pred.data(), t.data());
    std::printf("pa-multi trial %d W=%d S=%d L=%d n=%lld rc=%d\n", trial, W, S, L, (long long)n, rc);
  }
}
static void pa_trials(std::mt19937& g) {
  for (int trial = 0; trial < 8; ++trial) {
    int W = 1 + g() % 5, S = 1 + g() % 5, L = 1 + g() % 40; int64_t feats = 200 + g() % 3000, n = 200 + g() % 1500;
    std::vector<int64_t> rp(1, 0); std::vector<int32_t> cols; std::vector<float> vals; std::vector<int32_t> lab(n);
    for (int64_t e = 0; e < n; ++e) {
      int nnz = 1 + g() % 60;
      for (int j = 0; j < nnz; ++j) { cols.push_back(g() % feats); vals.push_back((int)(g() % 2001 - 1000) / 1000.f); }
      rp.push_back((int64_t)cols.size());
      lab[e] = (int)(g() % 3) - 1;
    }
    std::vector<float> w(feats, 0.f); std::vector<int32_t> pred(n); std::vector<uint8_t> t(feats);
    int rc = fps_host_pa_binary(rp.data(), cols.data(), vals.data(), lab.data(), n, feats, trial % 3, 0.1f, W, S, L,
                                trial & 1, w.data(), pred.data(), t.data());
    std::printf("pa trial %d W=%d S=%d L=%d n=%lld rc=%d\n", trial, W, S, L, (long long)n, rc);
  }
}
int main() {
  std::mt19937 g(1);
  for (int trial = 0; trial < 12; ++trial) {
    int W = 1 + g() % 6, S = 1 + g() % 6, L = 1 + g() % 17, k = 1 + g() % 20;
    int nu = 50 + g() % 200, ni = 20 + g() % 100; int64_t n = 2000 + g() % 20000;
    std::vector<int32_t> u(n), it(n); std::vector<float> r(n);
    for (int64_t i = 0; i < n; ++i) { u[i] = g() % nu; it[i] = g() % ni; r[i] = (g() % 1000) / 1000.f; }
    std::vector<float> ut((size_t)nu * k), vt((size_t)ni * k); std::vector<uint8_t> a(nu), b(ni); double sse = 0;
    int rc = fps_host_mf_train(u.data(), it.data(), r.data(), n, W, S, k, 0.05f, trial & 1, -0.1f, 0.1f, 7, 1 + trial % 3, L,
                               ut.data(), nu, vt.data(), ni, a.data(), b.data(), &sse, trial % 4, 1 + trial % 7);
    std::printf("trial %d W=%d S=%d L=%d k=%d n=%lld rc=%d sse=%.3f\n", trial, W, S, L, k, (long long)n, rc, sse);
  }
  pa_trials(g);
  pa_multi_trials(g);
  return 0;
}
