// Argument block of the fused matrix-factorisation step (fps_core.cu register-staged kernel and
// fps_mf_tma.cu TMA-pipelined kernel).  Mirrored by ops/native.py::MfArgsC.
#pragma once
#include "fps_common.cuh"

struct MfArgs {
  const void* users;
  const void* items;
  const float* ratings;
  long long n_pos;
  int neg_rate;
  long long num_items;        // negative-sample id range [0, num_items)
  unsigned long long seed;    // negative-sample stream key
  unsigned long long step;    // negative-sample stream counter (micro-batch number)
  float* user_table;          // worker-local [n_local_users, stride]
  int user_div;               // workerParallelism: local slot = user / user_div
  int user_shift;             // log2(user_div) if power of two, else -1
  float lr;
  int err_mode;               // 0: reference parity sigmoid(r - u.v); 1: plain residual r - u.v;
                              // 2: logistic r - sigmoid(u.v) (skip-gram negative sampling)
  int format;                 // 0: users/items/ratings arrays; 1: packed64 records in `users`
                              //    (user:26 | item:22 | rating fp16:16) -- 8 B/update over PCIe
  float* stats;               // [0] += sum (r-u.v)^2, [1] += #updates
  int* nan_flag;              // set to 1 if a non-finite update was produced
  ShardTable item_tab;
  ShardTable user_tab;        // used when user_sharded != 0: the "user" rows also live on the PS
  int user_sharded;           //   (word2vec: input vectors and output vectors are both PS tables)
  int use_push_tab;           // != 0: item deltas are pushed into push_tab instead of item_tab
  ShardTable push_tab;        //   (worker-side delta staging of the item-cache mode, see fps_cache_sync)
  int l2_hints;               // != 0: item rows evict_last, user rows evict_first (L2-blocked batches)
  int pad2_;
  unsigned int* progress;    // optional: CTA 0 publishes the record index it has reached (replica
                             //   exchange kernels follow the sweep over the L2-blocked batch)
  // E5 worker output stream (ps.output((user, userVector)) after every update,
  // PSOnlineMatrixFactorizationWorker.scala:52): record idx is emitted iff idx % out_every == 0, into the
  // device staging area of an OutputRing at slot *out_staged + idx / out_every (dropped beyond out_cap)
  long long* out_ids;
  float* out_vecs;
  const unsigned long long* out_staged;
  long long out_cap;
  int out_every;             // 0 = no output stream
  int pad3_;
  int* credits;              // device-side pull limiter (WL:196-250): credits[0] = pulls that may still be
                             //   issued, credits[1] = stall counter; nullptr = unlimited
};
