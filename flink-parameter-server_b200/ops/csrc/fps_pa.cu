// K7: passive-aggressive step on a CSR micro-batch, fused with the pulls and pushes.
//
// One CTA per example.  Phase 1 pulls the rows w[idx] of every active feature (16-byte loads from
// the owning shard -- local HBM or NVLink peer) and accumulates the decision vector d = W^T x.
// Phase 2 computes the update multipliers (binary PA / PA-I / PA-II, one-versus-all PA variants,
// cost-based PB / ML).  Phase 3 pushes x_i * mult to every active feature's row with
// red.global.add.v4.f32, i.e. the additive paramUpdate runs in the owner's memory system.
// Unlabelled examples only predict.  This is the reference's worker
// (PassiveAggressiveParameterServer.scala:283-340: nnz pulls + nnz pushes per example, ~10^4 each in
// its test) as ONE kernel per micro-batch; examples inside a micro-batch race like the reference's
// asynchronous workers do.
//
// Update rules reproduced: PassiveAggressiveBinaryAlgorithm.scala:44-112,
// PassiveAggressiveOneVersusAll.scala:38-123, PassiveAggressiveCostBased.scala:30-140.
#include "fps_common.cuh"

enum FpsPaAlgo : int { PA_PA = 0, PA_PAI = 1, PA_PAII = 2, PA_PB = 3, PA_ML = 4 };
#define PA_UNLABELLED (-2147483647 - 1)
#define PA_MAX_LABELS 1024
#define PA_THREADS 128

struct PaArgs {
  const long long* row_ptr;  // [n + 1]
  const void* col_idx;       // [nnz] feature ids
  const float* values;       // [nnz]
  const int* labels;         // [n]: binary: +1 / -1, multiclass: class index, PA_UNLABELLED: predict
  int* pred;                 // [n] out: binary 1/0, multiclass argmax
  float* loss;               // [n] out (optional): suffered loss (binary) / max label loss
  const float* cost;         // [L, L] cost matrix for PB / ML (may be null)
  long long n;
  int binary;                // 1: scalar weights, labels are +-1
  int num_labels;            // L (1 for binary)
  int algo;                  // FpsPaAlgo
  float aggressiveness;      // C
  int* nan_flag;
  ShardTable tab;            // one row per feature: [stride >= L] floats
};

template <typename IdT, int LPR>
__global__ void __launch_bounds__(PA_THREADS)
    fps_pa_step_kernel(const __grid_constant__ PaArgs a) {
  __shared__ float s_dec[PA_MAX_LABELS];   // decision vector, then the multipliers
  __shared__ float s_norm;
  __shared__ int s_aux[2];
  const IdT* __restrict__ cols = reinterpret_cast<const IdT*>(a.col_idx);
  const int L = a.num_labels;
  const int nvec = a.tab.stride >> 2;
  const int lane = threadIdx.x & (LPR - 1);
  const int grp = threadIdx.x / LPR;
  constexpr int NGRP = PA_THREADS / LPR;
  constexpr int VPL = (PA_MAX_LABELS / 4 + LPR - 1) / LPR > 8 ? 8 : (PA_MAX_LABELS / 4 + LPR - 1) / LPR;

  for (long long ex = blockIdx.x; ex < a.n; ex += gridDim.x) {
    const long long b = a.row_ptr[ex], e = a.row_ptr[ex + 1];
    for (int i = threadIdx.x; i < L; i += PA_THREADS) s_dec[i] = 0.f;
    if (threadIdx.x == 0) s_norm = 0.f;
    __syncthreads();
    // ---- phase 1: d = W^T x, ||x||^2 ------------------------------------------------------
    float4 acc[VPL];
#pragma unroll
    for (int c = 0; c < VPL; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    float n2 = 0.f;
    for (long long j = b + grp; j < e; j += NGRP) {
      const float x = a.values[j];
      const float* row = fps_row_t<IdT>(a.tab, cols[j]);
      if (lane == 0) n2 += x * x;
#pragma unroll
      for (int c = 0; c < VPL; ++c) {
        const int q = lane + c * LPR;
        if (q < nvec) {
          const float4 w = fps_ld_row4(row + 4 * q);  // the PULL
          acc[c].x += x * w.x; acc[c].y += x * w.y; acc[c].z += x * w.z; acc[c].w += x * w.w;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < VPL; ++c) {
      const int q = lane + c * LPR;
      if (q < nvec) {
        if (4 * q + 0 < L) atomicAdd(&s_dec[4 * q + 0], acc[c].x);
        if (4 * q + 1 < L) atomicAdd(&s_dec[4 * q + 1], acc[c].y);
        if (4 * q + 2 < L) atomicAdd(&s_dec[4 * q + 2], acc[c].z);
        if (4 * q + 3 < L) atomicAdd(&s_dec[4 * q + 3], acc[c].w);
      }
    }
    if (lane == 0 && n2 != 0.f) atomicAdd(&s_norm, n2);
    __syncthreads();
    // ---- phase 2: prediction + multipliers (thread 0 for scalars, all threads for OVA) -----
    const int label = a.labels[ex];
    const float nsq = s_norm;
    if (threadIdx.x == 0) {
      int arg = 0;
      float best = s_dec[0];
      for (int i = 1; i < L; ++i)
        if (s_dec[i] > best) { best = s_dec[i]; arg = i; }
      a.pred[ex] = a.binary ? (s_dec[0] > 0.f ? 1 : 0) : arg;
      s_aux[0] = arg;
      s_aux[1] = -1;
    }
    __syncthreads();
    bool push = (label != PA_UNLABELLED) && (e > b) && nsq > 0.f;
    if (push) {
      if (a.binary || a.algo <= PA_PAII) {
        // binary: y = label (+-1); OVA: y_i = +1 for the true class, -1 otherwise
        float max_loss = 0.f;
        for (int i = threadIdx.x; i < L; i += PA_THREADS) {
          const float y = a.binary ? (float)label : (i == label ? 1.f : -1.f);
          const float l = fmaxf(0.f, 1.f - y * s_dec[i]);
          float tau;
          if (a.algo == PA_PA) tau = l / nsq;
          else if (a.algo == PA_PAI) tau = fminf(a.aggressiveness, l / nsq);
          else tau = l / (nsq + 1.f / (2.f * a.aggressiveness));
          s_dec[i] = tau * y;  // multiplier for label column i
          max_loss = fmaxf(max_loss, l);
        }
        if (a.loss != nullptr && threadIdx.x == 0 && L == 1) a.loss[ex] = max_loss;
      } else {
        // cost based PB / ML: two columns get +-tau, everything else 0
        if (threadIdx.x == 0) {
          int q = s_aux[0];
          if (a.algo == PA_ML) {
            float bestv = -3.0e38f;
            for (int i = 0; i < L; ++i) {
              const float c = a.cost ? a.cost[label * L + i] : (i == label ? 0.f : 1.f);
              const float v = s_dec[i] - s_dec[label] + sqrtf(c);
              if (v > bestv) { bestv = v; q = i; }
            }
          }
          float tau = 0.f;
          if (q != label) {
            const float c = a.cost ? a.cost[label * L + q] : 1.f;
            const float l = s_dec[q] - s_dec[label] + sqrtf(c);
            tau = l / (2.f * nsq);
            if (a.loss != nullptr) a.loss[ex] = l;
          }
          s_aux[1] = q;
          s_norm = tau;
        }
        __syncthreads();
        const int q = s_aux[1];
        const float tau = s_norm;
        for (int i = threadIdx.x; i < L; i += PA_THREADS)
          s_dec[i] = (q == label) ? 0.f : (i == label ? tau : (i == q ? -tau : 0.f));
      }
    }
    __syncthreads();
    // ---- phase 3: push x_i * mult to every active feature --------------------------------
    if (push) {
      bool bad = false;
      for (long long j = b + grp; j < e; j += NGRP) {
        const float x = a.values[j];
        float* row = fps_row_t<IdT>(a.tab, cols[j]);
#pragma unroll
        for (int c = 0; c < VPL; ++c) {
          const int q = lane + c * LPR;
          if (q < nvec && 4 * q < L) {
            float4 d;
            d.x = x * s_dec[4 * q + 0];
            d.y = (4 * q + 1 < L) ? x * s_dec[4 * q + 1] : 0.f;
            d.z = (4 * q + 2 < L) ? x * s_dec[4 * q + 2] : 0.f;
            d.w = (4 * q + 3 < L) ? x * s_dec[4 * q + 3] : 0.f;
            if (d.x != 0.f || d.y != 0.f || d.z != 0.f || d.w != 0.f) {
              if (!(fabsf(d.x) <= 3.0e38f)) bad = true;
              fps_red_add4(row + 4 * q, d);  // the PUSH fused with paramUpdate (+)
            }
          }
        }
      }
      if (bad && a.nan_flag != nullptr) *a.nan_flag = 1;
    }
    __syncthreads();
  }
}

template <typename IdT>
static int dispatch_pa(const PaArgs& a, int num_sms, cudaStream_t s) {
  const int nvec = a.tab.stride >> 2;
  long long blocks = a.n;
  const long long cap = (long long)num_sms * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const int g = (int)blocks;
  if (nvec <= 1) fps_pa_step_kernel<IdT, 1><<<g, PA_THREADS, 0, s>>>(a);
  else if (nvec <= 2) fps_pa_step_kernel<IdT, 2><<<g, PA_THREADS, 0, s>>>(a);
  else if (nvec <= 4) fps_pa_step_kernel<IdT, 4><<<g, PA_THREADS, 0, s>>>(a);
  else if (nvec <= 8) fps_pa_step_kernel<IdT, 8><<<g, PA_THREADS, 0, s>>>(a);
  else if (nvec <= 16) fps_pa_step_kernel<IdT, 16><<<g, PA_THREADS, 0, s>>>(a);
  else fps_pa_step_kernel<IdT, 32><<<g, PA_THREADS, 0, s>>>(a);
  return (int)cudaGetLastError();
}

extern "C" int fps_pa_step(const PaArgs* args, int id_bytes, int num_sms, cudaStream_t stream) {
  if (args->n <= 0) return 0;
  if (args->num_labels > PA_MAX_LABELS || args->num_labels < 1) return -1006;
  if (args->num_labels > args->tab.stride) return -1006;
  if (id_bytes == 4) return dispatch_pa<int>(*args, num_sms, stream);
  if (id_bytes == 8) return dispatch_pa<long long>(*args, num_sms, stream);
  return -1001;
}
