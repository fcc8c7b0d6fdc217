// Native host runtime (C ABI, loaded with ctypes): the pieces of the data path that sit between the
// DataStream-shaped iterator and the pinned staging buffers of the device workers.
//
//  * rating partitioner / packer: the reference partitions the input stream with
//    `partitionCustom(user % workerParallelism)` (PSOnlineMatrixFactorization.scala:62-64); here one
//    multi-threaded pass buckets a rating block by owner worker and writes packed64 records
//    (user:26 | item:22 | fp16 rating:16) straight into per-worker (pinned) buffers.
//  * key interner: opaque 64-bit keys (e.g. String.hashCode of a word, or a hashed string id) to dense
//    slot ids for the device tables.
#include <atomic>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

static inline uint16_t f32_to_f16(float f) {
  uint32_t x;
  std::memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  int32_t exp = (int32_t)((x >> 23) & 0xFF) - 127 + 15;
  uint32_t man = x & 0x7FFFFFu;
  if (exp >= 31) return (uint16_t)(sign | 0x7C00u | (((x >> 23) & 0xFF) == 0xFF && man ? 0x200u : 0));
  if (exp <= 0) {
    if (exp < -10) return (uint16_t)sign;
    man |= 0x800000u;
    const int shift = 14 - exp;
    uint32_t h = man >> shift;
    const uint32_t rem = man & ((1u << shift) - 1), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1))) ++h;
    return (uint16_t)(sign | h);
  }
  uint32_t h = ((uint32_t)exp << 10) | (man >> 13);
  const uint32_t rem = man & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;
  return (uint16_t)(sign | h);
}

extern "C" {

// counts[w] <- number of ratings owned by worker w
int fps_host_count_by_worker(const int32_t* users, int64_t n, int32_t workers, int64_t* counts) {
  for (int w = 0; w < workers; ++w) counts[w] = 0;
  for (int64_t i = 0; i < n; ++i) counts[(uint32_t)users[i] % (uint32_t)workers]++;
  return 0;
}

// Scatter ratings into per-worker packed64 buffers; out[w] must hold counts[w] records.
// Order inside a worker's buffer preserves stream order (stable), like a Flink partitioner.
int fps_host_partition_pack(const int32_t* users, const int32_t* items, const float* ratings,
                            int64_t n, int32_t workers, uint64_t** out, int32_t n_threads) {
  if (n_threads < 1) n_threads = 1;
  const int64_t chunk = (n + n_threads - 1) / n_threads;
  std::vector<std::vector<int64_t>> cnt(n_threads, std::vector<int64_t>(workers, 0));
  auto count = [&](int t) {
    const int64_t b = t * chunk, e = std::min(n, b + chunk);
    for (int64_t i = b; i < e; ++i) cnt[t][(uint32_t)users[i] % (uint32_t)workers]++;
  };
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t) th.emplace_back(count, t);
  for (auto& x : th) x.join();
  th.clear();
  std::vector<std::vector<int64_t>> off(n_threads, std::vector<int64_t>(workers, 0));
  for (int w = 0; w < workers; ++w) {
    int64_t acc = 0;
    for (int t = 0; t < n_threads; ++t) { off[t][w] = acc; acc += cnt[t][w]; }
  }
  std::atomic<int> bad{0};
  auto scatter = [&](int t) {
    const int64_t b = t * chunk, e = std::min(n, b + chunk);
    std::vector<int64_t> pos = off[t];
    for (int64_t i = b; i < e; ++i) {
      const uint32_t u = (uint32_t)users[i], it = (uint32_t)items[i];
      if (u >= (1u << 26) || it >= (1u << 22)) { bad = 1; continue; }
      const uint32_t w = u % (uint32_t)workers;
      out[w][pos[w]++] = ((uint64_t)u << 38) | ((uint64_t)it << 16) | f32_to_f16(ratings[i]);
    }
  };
  for (int t = 0; t < n_threads; ++t) th.emplace_back(scatter, t);
  for (auto& x : th) x.join();
  return bad.load() ? -1 : 0;
}

// ---- key interner ---------------------------------------------------------------------------
struct Interner {
  std::unordered_map<int64_t, int32_t> map;
  std::vector<int64_t> keys;
  std::mutex mu;
};
void* fps_interner_new() { return new Interner(); }
void fps_interner_free(void* p) { delete static_cast<Interner*>(p); }
int64_t fps_interner_size(void* p) { return (int64_t)static_cast<Interner*>(p)->keys.size(); }
// slots[i] <- dense slot of keys[i]; unseen keys get the next free slot (insert) or -1 (lookup only)
int fps_interner_map(void* p, const int64_t* keys, int64_t n, int32_t* slots, int insert) {
  Interner* in = static_cast<Interner*>(p);
  std::lock_guard<std::mutex> g(in->mu);
  for (int64_t i = 0; i < n; ++i) {
    auto it = in->map.find(keys[i]);
    if (it != in->map.end()) { slots[i] = it->second; continue; }
    if (!insert) { slots[i] = -1; continue; }
    const int32_t s = (int32_t)in->keys.size();
    in->map.emplace(keys[i], s);
    in->keys.push_back(keys[i]);
    slots[i] = s;
  }
  return 0;
}
int fps_interner_keys(void* p, int64_t* out) {
  Interner* in = static_cast<Interner*>(p);
  std::lock_guard<std::mutex> g(in->mu);
  std::memcpy(out, in->keys.data(), in->keys.size() * sizeof(int64_t));
  return 0;
}

}  // extern "C"
