// Native host runtime (C ABI, loaded with ctypes): the pieces of the data path that sit between the
// DataStream-shaped iterator and the pinned staging buffers of the device workers.
//
//  * rating partitioner / packer: the reference partitions the input stream with
//    `partitionCustom(user % workerParallelism)` (PSOnlineMatrixFactorization.scala:62-64); here one
//    multi-threaded pass buckets a rating block by owner worker and writes packed64 records
//    (user:26 | item:22 | fp16 rating:16) straight into per-worker (pinned) buffers.
//  * key interner: opaque 64-bit keys (e.g. String.hashCode of a word, or a hashed string id) to dense
//    slot ids for the device tables.
//  * native asynchronous MF engine: worker and server threads exchanging pull / answer / push messages over
//    lock-free SPSC rings (the CPU backend of psOnlineMF / psOfflineMF at native speed, see the end).
#include <algorithm>
#include <atomic>
#include <cmath>
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

// ---------------------------------------------------------------------------------------------
// Native host engine for asynchronous SGD matrix factorisation (the CPU backend of psOnlineMF /
// psOfflineMF at native speed).  It runs the reference's protocol, not a shortcut:
//   * W worker threads own the users (user % W), S server threads own the item vectors (item % S)
//     (PSOnlineMatrixFactorization.scala:58-64);
//   * per rating: Pull(item) -> PullAnswer(item, vector) -> SGD delta -> local user update ->
//     Push(item, delta) -> paramUpdate = vector sum (PSOnlineMatrixFactorizationWorker.scala:42-89);
//   * at most `pull_limit` unanswered pulls per worker (WorkerLogic.addPullLimiter, WL:196-250);
//   * messages travel through single-producer / single-consumer rings, one per (worker, server) pair and
//     direction, so delivery is FIFO per pair like Flink's channels and answers are matched to ratings by
//     order (the reference's per-item rating queues rely on the same property).
// Ring capacities make sends non-blocking by construction: a worker has at most `pull_limit` pulls and
// `pull_limit` pushes in flight towards one server, a server at most `pull_limit` answers towards one worker.
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int MF_MAX_K = 128;

struct MfMsg {
  int32_t kind;  // 0 pull, 1 push, 2 answer
  int32_t id;
  float v[MF_MAX_K];
};

struct MfRing {  // SPSC
  std::vector<MfMsg> buf;
  size_t cap = 0;
  alignas(64) std::atomic<size_t> head{0};  // next to pop
  alignas(64) std::atomic<size_t> tail{0};  // next to push
  void init(size_t c) { cap = c; buf.resize(c); }
  MfMsg* begin_push() {
    const size_t t = tail.load(std::memory_order_relaxed);
    if (t - head.load(std::memory_order_acquire) >= cap) return nullptr;
    return &buf[t % cap];
  }
  void end_push() { tail.store(tail.load(std::memory_order_relaxed) + 1, std::memory_order_release); }
  MfMsg* front() {
    const size_t h = head.load(std::memory_order_relaxed);
    if (h == tail.load(std::memory_order_acquire)) return nullptr;
    return &buf[h % cap];
  }
  void pop() { head.store(head.load(std::memory_order_relaxed) + 1, std::memory_order_release); }
};

static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// deterministic "init on first pull": a pure function of (seed, id, column)
static inline float init_value(uint64_t seed, int64_t id, int j, float lo, float hi) {
  const uint64_t h = splitmix64(seed ^ splitmix64((uint64_t)id * 0x100000001B3ull + (uint64_t)j));
  return lo + (hi - lo) * (float)((h >> 40) * (1.0 / 16777216.0));
}

}  // namespace

extern "C" {

// users/items/ratings: the whole stream (every worker filters its own users, order preserved).
// user_table [num_users, k], item_table [num_items, k] are initialised here (init_value) and trained in
// place; *_touched are set for ids that occurred.  err_mode 0: e = sigmoid(r - u.v) (SGDUpdater.scala:8),
// 1: e = r - u.v.  Returns 0, or -1 for bad arguments, -2 if a non-finite update appeared.
int fps_host_mf_train(const int32_t* users, const int32_t* items, const float* ratings, int64_t n,
                      int32_t workers, int32_t servers, int32_t k, float lr, int32_t err_mode, float lo,
                      float hi, uint64_t seed, int32_t epochs, int32_t pull_limit, float* user_table,
                      int64_t num_users, float* item_table, int64_t num_items, uint8_t* user_touched,
                      uint8_t* item_touched, double* sum_sq_err, int32_t neg_rate, int32_t user_memory) {
  if (workers < 1 || servers < 1 || k < 1 || k > MF_MAX_K || pull_limit < 1 || epochs < 1 || neg_rate < 0 ||
      user_memory < 0)
    return -1;
  for (int64_t i = 0; i < n; ++i)
    if (users[i] < 0 || users[i] >= num_users || items[i] < 0 || items[i] >= num_items) return -1;
  for (int64_t u = 0; u < num_users; ++u)
    for (int j = 0; j < k; ++j) user_table[u * k + j] = init_value(seed * 2 + 2, u, j, lo, hi);
  for (int64_t it = 0; it < num_items; ++it)
    for (int j = 0; j < k; ++j) item_table[it * k + j] = init_value(seed * 2 + 1, it, j, lo, hi);

  const size_t cap_w2s = 2 * (size_t)pull_limit + 8, cap_s2w = (size_t)pull_limit + 8;
  std::vector<MfRing> w2s((size_t)workers * servers), s2w((size_t)workers * servers);
  for (auto& r : w2s) r.init(cap_w2s);
  for (auto& r : s2w) r.init(cap_s2w);
  std::atomic<int> workers_done{0};
  std::atomic<int> bad{0};
  std::vector<double> sq(workers, 0.0);

  struct Rec { int32_t user, item; float rating; };
  auto worker = [&](int w) {
    std::vector<int64_t> mine;
    for (int64_t i = 0; i < n; ++i)
      if (users[i] % workers == w) mine.push_back(i);
    std::vector<std::vector<Rec>> pend(servers);  // FIFO of records awaiting an answer, per server
    std::vector<size_t> pend_head(servers, 0);
    // negative sampling state (PSOnlineMatrixFactorizationWorker.scala:61-78): the items this worker has
    // seen so far, and per user the last `user_memory` rated items
    std::vector<int32_t> item_ids;
    std::vector<uint8_t> item_known(neg_rate > 0 ? (size_t)num_items : 0, 0);
    std::unordered_map<int32_t, std::vector<int32_t>> seen_q;  // ring, oldest first
    uint64_t rng = splitmix64(seed ^ (0xA5A5A5A5ull + (uint64_t)w));
    std::vector<Rec> todo;  // records to issue for the current positive rating (negatives first)
    size_t todo_pos = 0;
    double acc = 0.0;
    for (int ep = 0; ep < epochs; ++ep) {
      size_t next = 0;
      int outstanding = 0;
      while (next < mine.size() || todo_pos < todo.size() || outstanding > 0) {
        bool progressed = false;
        while (outstanding < pull_limit) {  // the pull limiter
          if (todo_pos == todo.size()) {
            if (next == mine.size()) break;
            const int64_t idx = mine[next++];
            todo.clear(); todo_pos = 0;
            if (neg_rate > 0) {
              std::vector<int32_t>& q = seen_q[users[idx]];
              if ((int)q.size() >= user_memory && !q.empty()) q.erase(q.begin());
              if (user_memory > 0) q.push_back(items[idx]);
              const int64_t room = (int64_t)item_ids.size() - (int64_t)q.size();
              const int64_t k_neg = std::max<int64_t>(0, std::min<int64_t>(room, neg_rate));
              for (int64_t t = 0; t < k_neg; ++t) {
                int32_t cand;
                do {
                  rng = splitmix64(rng);
                  cand = item_ids[(size_t)(rng % item_ids.size())];
                } while (std::find(q.begin(), q.end(), cand) != q.end());
                todo.push_back(Rec{users[idx], cand, 0.f});
              }
              if (!item_known[items[idx]]) { item_known[items[idx]] = 1; item_ids.push_back(items[idx]); }
            }
            todo.push_back(Rec{users[idx], items[idx], ratings[idx]});
          }
          const Rec& r = todo[todo_pos];
          const int s = r.item % servers;
          MfMsg* m = w2s[(size_t)w * servers + s].begin_push();
          if (m == nullptr) break;
          m->kind = 0; m->id = r.item;
          w2s[(size_t)w * servers + s].end_push();
          pend[s].push_back(r);
          ++todo_pos; ++outstanding; progressed = true;
        }
        for (int s = 0; s < servers; ++s) {
          MfRing& in = s2w[(size_t)w * servers + s];
          while (MfMsg* a = in.front()) {
            MfRing& out = w2s[(size_t)w * servers + s];
            MfMsg* p = out.begin_push();
            if (p == nullptr) break;  // cannot happen by the capacity argument; stay safe
            const Rec r = pend[s][pend_head[s]++];
            float* u = user_table + (int64_t)r.user * k;
            float dot = 0.f;
            for (int j = 0; j < k; ++j) dot += u[j] * a->v[j];
            const float resid = r.rating - dot;
            const float e = err_mode == 0 ? 1.f / (1.f + std::exp(-resid)) : resid;
            const float g = lr * e;
            if (!(std::fabs(g) <= 3.0e38f)) bad = 1;
            acc += (double)resid * resid;
            p->kind = 1; p->id = a->id;
            for (int j = 0; j < k; ++j) {
              p->v[j] = g * u[j];        // item delta (uses the user vector BEFORE its update, like delta())
              u[j] += g * a->v[j];       // worker-local user update
            }
            out.end_push();
            in.pop();
            user_touched[r.user] = 1;
            --outstanding; progressed = true;
          }
        }
        if (!progressed) std::this_thread::yield();
      }
      for (int s = 0; s < servers; ++s) { pend[s].clear(); pend_head[s] = 0; }
    }
    sq[w] = acc;
    workers_done.fetch_add(1, std::memory_order_release);
  };

  auto server = [&](int s) {
    while (true) {
      bool progressed = false;
      const bool all_done = workers_done.load(std::memory_order_acquire) == workers;
      for (int w = 0; w < workers; ++w) {
        MfRing& in = w2s[(size_t)w * servers + s];
        while (MfMsg* m = in.front()) {
          float* row = item_table + (int64_t)m->id * k;
          if (m->kind == 0) {
            MfRing& out = s2w[(size_t)w * servers + s];
            MfMsg* a = out.begin_push();
            if (a == nullptr) break;  // answer ring full: the worker will drain it
            a->kind = 2; a->id = m->id;
            std::memcpy(a->v, row, sizeof(float) * k);
            out.end_push();
            item_touched[m->id] = 1;
          } else {
            for (int j = 0; j < k; ++j) row[j] += m->v[j];  // paramUpdate = vectorSum
          }
          in.pop();
          progressed = true;
        }
      }
      if (!progressed) {
        if (all_done) {
          bool empty = true;
          for (int w = 0; w < workers; ++w) empty = empty && w2s[(size_t)w * servers + s].front() == nullptr;
          if (empty) break;
        }
        std::this_thread::yield();
      }
    }
  };

  std::vector<std::thread> th;
  for (int s = 0; s < servers; ++s) th.emplace_back(server, s);
  for (int w = 0; w < workers; ++w) th.emplace_back(worker, w);
  for (auto& t : th) t.join();
  double total = 0.0;
  for (double x : sq) total += x;
  if (sum_sq_err != nullptr) *sum_sq_err = total;
  return bad.load() ? -2 : 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Native host engine for the binary passive-aggressive classifiers (PA / PA-I / PA-II;
// PassiveAggressiveParameterServer.scala:283-340, PassiveAggressiveBinaryAlgorithm.scala:44-112):
// one scalar weight per feature on the server threads (hash or range partitioned), worker threads take
// the examples round-robin, pull every active feature of an example (bounded by the pull limiter),
// and once all answers are in either predict (unlabelled) or push tau*y*x_i per feature.
// Same SPSC-ring transport as the MF engine above.
// ---------------------------------------------------------------------------------------------
namespace {

struct PaMsg {
  int32_t kind;  // 0 pull, 1 push, 2 answer
  int32_t id;
  float v;
  int32_t pad;
};

struct PaRing {
  std::vector<PaMsg> buf;
  size_t cap = 0;
  alignas(64) std::atomic<size_t> head{0};
  alignas(64) std::atomic<size_t> tail{0};
  void init(size_t c) { cap = c; buf.resize(c); }
  bool push(const PaMsg& m) {
    const size_t t = tail.load(std::memory_order_relaxed);
    if (t - head.load(std::memory_order_acquire) >= cap) return false;
    buf[t % cap] = m;
    tail.store(t + 1, std::memory_order_release);
    return true;
  }
  bool pop(PaMsg& m) {
    const size_t h = head.load(std::memory_order_relaxed);
    if (h == tail.load(std::memory_order_acquire)) return false;
    m = buf[h % cap];
    head.store(h + 1, std::memory_order_release);
    return true;
  }
  bool empty() { return head.load(std::memory_order_relaxed) == tail.load(std::memory_order_acquire); }
};

}  // namespace

extern "C" {

// CSR examples (row_ptr[n+1], cols, vals); labels: +1 / -1, or 0 = unlabelled (predict only).
// algo: 0 PA, 1 PA-I, 2 PA-II; weights[feature_count] in/out (initial model); pred[n] out (1 / 0);
// touched[feature_count] out.  range_partition: contiguous feature ranges per server
// (RangePSLogicWithClose.scala:51-62) instead of feature % servers.
int fps_host_pa_binary(const int64_t* row_ptr, const int32_t* cols, const float* vals, const int32_t* labels,
                       int64_t n, int64_t feature_count, int32_t algo, float C, int32_t workers,
                       int32_t servers, int32_t pull_limit, int32_t range_partition, float* weights,
                       int32_t* pred, uint8_t* touched) {
  if (workers < 1 || servers < 1 || pull_limit < 1 || feature_count < 1 || algo < 0 || algo > 2) return -1;
  for (int64_t j = 0; j < row_ptr[n]; ++j)
    if (cols[j] < 0 || cols[j] >= feature_count) return -1;
  const int64_t div = (feature_count + servers - 1) / servers;
  auto owner = [&](int32_t f) -> int {
    if (!range_partition) return f % servers;
    const int64_t s = f / div;
    return (int)(s < servers ? s : servers - 1);
  };
  std::vector<PaRing> w2s((size_t)workers * servers), s2w((size_t)workers * servers);
  for (auto& r : w2s) r.init(2 * (size_t)pull_limit + 64);
  for (auto& r : s2w) r.init((size_t)pull_limit + 8);
  std::atomic<int> workers_done{0};
  std::atomic<int> bad{0};

  auto drain_answers = [&](int w, std::vector<std::vector<int64_t>>& pend, std::vector<size_t>& pend_head,
                           std::vector<float>& wbuf, int64_t& received) {
    bool any = false;
    PaMsg a;
    for (int s = 0; s < servers; ++s)
      while (s2w[(size_t)w * servers + s].pop(a)) {
        wbuf[pend[s][pend_head[s]++]] = a.v;
        ++received;
        any = true;
      }
    return any;
  };

  auto worker = [&](int w) {
    std::vector<std::vector<int64_t>> pend(servers);
    std::vector<size_t> pend_head(servers, 0);
    std::vector<float> wbuf;
    for (int64_t ex = w; ex < n; ex += workers) {
      const int64_t b = row_ptr[ex], m = row_ptr[ex + 1] - b;
      wbuf.assign((size_t)m, 0.f);
      for (int s = 0; s < servers; ++s) { pend[s].clear(); pend_head[s] = 0; }
      int64_t issued = 0, received = 0;
      while (received < m) {
        bool progressed = false;
        while (issued < m && issued - received < pull_limit) {  // the pull limiter
          const int s = owner(cols[b + issued]);
          if (!w2s[(size_t)w * servers + s].push(PaMsg{0, cols[b + issued], 0.f, 0})) break;
          pend[s].push_back(issued);
          ++issued;
          progressed = true;
        }
        progressed |= drain_answers(w, pend, pend_head, wbuf, received);
        if (!progressed) std::this_thread::yield();
      }
      float dot = 0.f, nsq = 0.f;
      for (int64_t j = 0; j < m; ++j) { dot += wbuf[j] * vals[b + j]; nsq += vals[b + j] * vals[b + j]; }
      pred[ex] = dot > 0.f ? 1 : 0;
      const int y = labels[ex];
      if (y == 0 || m == 0 || !(nsq > 0.f)) continue;
      const float loss = std::max(0.f, 1.f - (float)y * dot);
      float tau;
      if (algo == 0) tau = loss / nsq;
      else if (algo == 1) tau = std::min(C, loss / nsq);
      else tau = loss / (nsq + 1.f / (2.f * C));
      if (tau == 0.f) continue;
      if (!(std::fabs(tau) <= 3.0e38f)) bad = 1;
      for (int64_t j = 0; j < m; ++j) {
        const int s = owner(cols[b + j]);
        const PaMsg p{1, cols[b + j], tau * (float)y * vals[b + j], 0};
        while (!w2s[(size_t)w * servers + s].push(p)) std::this_thread::yield();  // server always drains
      }
    }
    workers_done.fetch_add(1, std::memory_order_release);
  };

  auto server = [&](int s) {
    PaMsg m;
    while (true) {
      bool progressed = false;
      const bool all_done = workers_done.load(std::memory_order_acquire) == workers;
      for (int w = 0; w < workers; ++w) {
        PaRing& in = w2s[(size_t)w * servers + s];
        PaRing& out = s2w[(size_t)w * servers + s];
        while (true) {
          const size_t h = in.head.load(std::memory_order_relaxed);
          if (h == in.tail.load(std::memory_order_acquire)) break;
          const PaMsg& peek = in.buf[h % in.cap];
          if (peek.kind == 0) {
            if (!out.push(PaMsg{2, peek.id, weights[peek.id], 0})) break;  // answer ring full: retry later
            touched[peek.id] = 1;
          } else {
            weights[peek.id] += peek.v;  // paramUpdate = +
            touched[peek.id] = 1;
          }
          in.head.store(h + 1, std::memory_order_release);
          progressed = true;
        }
      }
      if (!progressed) {
        if (all_done) {
          bool empty = true;
          for (int w = 0; w < workers; ++w) empty = empty && w2s[(size_t)w * servers + s].empty();
          if (empty) break;
        }
        std::this_thread::yield();
      }
    }
    (void)m;
  };

  std::vector<std::thread> th;
  for (int s = 0; s < servers; ++s) th.emplace_back(server, s);
  for (int w = 0; w < workers; ++w) th.emplace_back(worker, w);
  for (auto& t : th) t.join();
  return bad.load() ? -2 : 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Multiclass passive-aggressive on the native engine: one weight vector of `L` labels per feature;
// one-versus-all PA / PA-I / PA-II (PassiveAggressiveOneVersusAll.scala:38-123) and the cost-based PB / ML
// rules (PassiveAggressiveCostBased.scala:30-140).  Same protocol and rings as fps_host_pa_binary, with
// L-vector payloads (L <= MF_MAX_K).
// ---------------------------------------------------------------------------------------------
extern "C" {

// labels[ex] in [0, L) or -1 (= predict only).  algo: 0 PA, 1 PA-I, 2 PA-II (one versus all), 3 PB, 4 ML
// (cost based; cost[L*L] row = true label, may be null = 0/1 cost).  weights [feature_count, L] in/out.
int fps_host_pa_multiclass(const int64_t* row_ptr, const int32_t* cols, const float* vals, const int32_t* labels,
                           int64_t n, int64_t feature_count, int32_t L, int32_t algo, float C, const float* cost,
                           int32_t workers, int32_t servers, int32_t pull_limit, int32_t range_partition,
                           float* weights, int32_t* pred, uint8_t* touched) {
  if (workers < 1 || servers < 1 || pull_limit < 1 || feature_count < 1 || algo < 0 || algo > 4 || L < 1 ||
      L > MF_MAX_K)
    return -1;
  for (int64_t j = 0; j < row_ptr[n]; ++j)
    if (cols[j] < 0 || cols[j] >= feature_count) return -1;
  for (int64_t e = 0; e < n; ++e)
    if (labels[e] < -1 || labels[e] >= L) return -1;
  const int64_t div = (feature_count + servers - 1) / servers;
  auto owner = [&](int32_t f) -> int {
    if (!range_partition) return f % servers;
    const int64_t s = f / div;
    return (int)(s < servers ? s : servers - 1);
  };
  std::vector<MfRing> w2s((size_t)workers * servers), s2w((size_t)workers * servers);
  for (auto& r : w2s) r.init(2 * (size_t)pull_limit + 64);
  for (auto& r : s2w) r.init((size_t)pull_limit + 8);
  std::atomic<int> workers_done{0};
  std::atomic<int> bad{0};

  auto worker = [&](int w) {
    std::vector<std::vector<int64_t>> pend(servers);
    std::vector<size_t> pend_head(servers, 0);
    std::vector<float> dec(L), mult(L);
    for (int64_t ex = w; ex < n; ex += workers) {
      const int64_t b = row_ptr[ex], m = row_ptr[ex + 1] - b;
      std::fill(dec.begin(), dec.end(), 0.f);
      for (int s = 0; s < servers; ++s) { pend[s].clear(); pend_head[s] = 0; }
      int64_t issued = 0, received = 0;
      while (received < m) {
        bool progressed = false;
        while (issued < m && issued - received < pull_limit) {
          const int s = owner(cols[b + issued]);
          MfMsg* p = w2s[(size_t)w * servers + s].begin_push();
          if (p == nullptr) break;
          p->kind = 0; p->id = cols[b + issued];
          w2s[(size_t)w * servers + s].end_push();
          pend[s].push_back(issued);
          ++issued; progressed = true;
        }
        for (int s = 0; s < servers; ++s) {
          MfRing& in = s2w[(size_t)w * servers + s];
          while (MfMsg* a = in.front()) {
            const float x = vals[b + pend[s][pend_head[s]++]];
            for (int l = 0; l < L; ++l) dec[l] += x * a->v[l];  // d = W^T x
            in.pop();
            ++received; progressed = true;
          }
        }
        if (!progressed) std::this_thread::yield();
      }
      float nsq = 0.f;
      for (int64_t j = 0; j < m; ++j) nsq += vals[b + j] * vals[b + j];
      int arg = 0;
      for (int l = 1; l < L; ++l)
        if (dec[l] > dec[arg]) arg = l;
      pred[ex] = arg;
      const int y = labels[ex];
      if (y < 0 || m == 0 || !(nsq > 0.f)) continue;
      bool any = false;
      if (algo <= 2) {
        for (int l = 0; l < L; ++l) {
          const float yl = l == y ? 1.f : -1.f;
          const float loss = std::max(0.f, 1.f - yl * dec[l]);
          float tau;
          if (algo == 0) tau = loss / nsq;
          else if (algo == 1) tau = std::min(C, loss / nsq);
          else tau = loss / (nsq + 1.f / (2.f * C));
          mult[l] = tau * yl;
          any = any || mult[l] != 0.f;
        }
      } else {
        int q = arg;  // PB: the predicted label; ML: the label maximising score difference + sqrt(cost)
        if (algo == 4) {
          float bestv = -3.0e38f;
          for (int l = 0; l < L; ++l) {
            const float c = cost ? cost[(size_t)y * L + l] : (l == y ? 0.f : 1.f);
            const float v = dec[l] - dec[y] + std::sqrt(c);
            if (v > bestv) { bestv = v; q = l; }
          }
        }
        std::fill(mult.begin(), mult.end(), 0.f);
        if (q != y) {
          const float c = cost ? cost[(size_t)y * L + q] : 1.f;
          const float tau = (dec[q] - dec[y] + std::sqrt(c)) / (2.f * nsq);
          mult[y] = tau; mult[q] = -tau;
          any = tau != 0.f;
        }
      }
      if (!any) continue;
      for (int64_t j = 0; j < m; ++j) {
        const int s = owner(cols[b + j]);
        MfRing& out = w2s[(size_t)w * servers + s];
        MfMsg* p;
        while ((p = out.begin_push()) == nullptr) std::this_thread::yield();
        p->kind = 1; p->id = cols[b + j];
        for (int l = 0; l < L; ++l) {
          p->v[l] = vals[b + j] * mult[l];
          if (!(std::fabs(p->v[l]) <= 3.0e38f)) bad = 1;
        }
        out.end_push();
      }
    }
    workers_done.fetch_add(1, std::memory_order_release);
  };

  auto server = [&](int s) {
    while (true) {
      bool progressed = false;
      const bool all_done = workers_done.load(std::memory_order_acquire) == workers;
      for (int w = 0; w < workers; ++w) {
        MfRing& in = w2s[(size_t)w * servers + s];
        while (MfMsg* m = in.front()) {
          float* row = weights + (int64_t)m->id * L;
          if (m->kind == 0) {
            MfRing& out = s2w[(size_t)w * servers + s];
            MfMsg* a = out.begin_push();
            if (a == nullptr) break;
            a->kind = 2; a->id = m->id;
            std::memcpy(a->v, row, sizeof(float) * L);
            out.end_push();
          } else {
            for (int l = 0; l < L; ++l) row[l] += m->v[l];
          }
          touched[m->id] = 1;
          in.pop();
          progressed = true;
        }
      }
      if (!progressed) {
        if (all_done) {
          bool empty = true;
          for (int w = 0; w < workers; ++w) empty = empty && w2s[(size_t)w * servers + s].front() == nullptr;
          if (empty) break;
        }
        std::this_thread::yield();
      }
    }
  };

  std::vector<std::thread> th;
  for (int s = 0; s < servers; ++s) th.emplace_back(server, s);
  for (int w = 0; w < workers; ++w) th.emplace_back(worker, w);
  for (auto& t : th) t.join();
  return bad.load() ? -2 : 0;
}

}  // extern "C"
