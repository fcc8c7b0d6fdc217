// E5 + K11: the worker output stream of the device tier.
//
// The reference's workers call ps.output((user, userVector)) after every update
// (PSOnlineMatrixFactorizationWorker.scala:52) and Flink carries the records to a sink.  Here the fused
// kernel drops the records into a DEVICE staging area; after every micro-batch
//   * fps_output_policy (1 thread) -- the device-side CountLogic / TimerLogic / AND / OR
//     (CountLogic.scala:5-29, TimerLogic.scala:6-51) on the staged records: "count >= max" and / or
//     "globaltimer deadline passed" -- decides whether to flush;
//   * fps_output_flush copies the staged records into a ring in PINNED HOST memory (mapped into the device)
//     and the last block publishes the new tail with a system-scope release store.  The host iterator
//     polls the tail word and reads the records: no stream synchronisation, no cudaMemcpy on that path.
// A full ring drops the newest records and counts them (no device-side waiting on the host).
#include "fps_common.cuh"
#include "fps_tma.cuh"

struct OutState {                    // device memory
  unsigned long long staged;         // records in the staging area
  unsigned long long tail;           // records ever published to the host ring
  unsigned long long last_ns;        // globaltimer of the last flush
  unsigned long long n_flush;        // OUT of the policy kernel: records to flush now (0 = none)
  unsigned long long flush_tail;     //   ... to ring position flush_tail
  unsigned long long flushes;
  unsigned long long dropped;
  unsigned int blocks_done;          // completion counter of the flush grid
  unsigned int pad_;
};

struct OutPolicy {
  unsigned long long count_max;      // 0 = off
  unsigned long long interval_ns;    // 0 = off
  unsigned long long n_new;          // records the training kernel just staged
  unsigned long long staging_cap;
  unsigned long long ring_cap;
  int require_all;
  int force;
};

__global__ void fps_output_policy_kernel(const OutPolicy p, OutState* st,
                                         const volatile unsigned long long* host_head) {
  if (threadIdx.x != 0) return;
  const unsigned long long now = fps_globaltimer_ns();
  unsigned long long staged = st->staged + p.n_new;
  if (staged > p.staging_cap) {      // the kernel dropped what did not fit
    st->dropped += staged - p.staging_cap;
    staged = p.staging_cap;
  }
  if (st->last_ns == 0) st->last_ns = now;
  const bool has_c = p.count_max != 0, has_t = p.interval_ns != 0;
  const bool cnt = has_c && staged >= p.count_max;
  const bool tim = has_t && staged != 0 && now - st->last_ns >= p.interval_ns;
  bool fire = p.force ? staged != 0
                      : (p.require_all ? ((has_c || has_t) && (!has_c || cnt) && (!has_t || tim)) : (cnt || tim));
  unsigned long long n = 0;
  if (fire) {
    const unsigned long long used = st->tail - *host_head;
    const unsigned long long room = p.ring_cap > used ? p.ring_cap - used : 0;
    n = staged < room ? staged : room;
    st->dropped += staged - n;
    st->flush_tail = st->tail;
    st->last_ns = now;
    st->flushes += 1;
    staged = 0;
  }
  st->n_flush = n;
  st->staged = staged;
  st->blocks_done = 0;
}

__global__ void __launch_bounds__(256)
    fps_output_flush_kernel(OutState* st, long long* __restrict__ s_ids, const float* __restrict__ s_vecs,
                            int stride, long long* ring_ids, float* ring_vecs, unsigned long long ring_cap,
                            volatile unsigned long long* host_tail) {
  const unsigned long long n = st->n_flush;
  if (n == 0) return;
  const unsigned long long t0 = st->flush_tail;
  const int nvec = stride >> 2;
  const unsigned long long total = n * (unsigned long long)nvec;
  for (unsigned long long x = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; x < total;
       x += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long rec = x / nvec;
    const int q = (int)(x - rec * nvec);
    const unsigned long long pos = (t0 + rec) % ring_cap;
    reinterpret_cast<float4*>(ring_vecs + pos * (unsigned long long)stride)[q] =
        reinterpret_cast<const float4*>(s_vecs + rec * (unsigned long long)stride)[q];
    if (q == 0) {
      ring_ids[pos] = s_ids[rec];
      s_ids[rec] = -1;               // a slot the next kernel does not fill (voided record) reads as "no record"
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(&st->blocks_done, 1u) + 1u;
    if (done == gridDim.x) {         // last block: everything is in host memory -> publish
      __threadfence_system();
      st->tail = t0 + n;
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(host_tail), "l"(t0 + n) : "memory");
    }
  }
}

extern "C" int fps_output_step(const OutPolicy* p, OutState* st, long long* s_ids, const float* s_vecs,
                               int stride, long long* ring_ids, float* ring_vecs,
                               unsigned long long* host_tail, const unsigned long long* host_head,
                               int num_sms, cudaStream_t stream) {
  fps_output_policy_kernel<<<1, 32, 0, stream>>>(*p, st, host_head);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  fps_output_flush_kernel<<<num_sms, 256, 0, stream>>>(st, s_ids, s_vecs, stride, ring_ids, ring_vecs,
                                                      p->ring_cap, host_tail);
  return (int)cudaGetLastError();
}
