// K8-K10: streaming-sketch updates fused with the push (one-sided reductions on the owner's HBM).
//
//   Bloom      : bitmap[key][pos >> 5] |= 1 << (pos & 31)        red.global.or.b32
//                (BloomPSLogic.scala:16-22: set bits)
//   tug-of-war : counter[key][j] += bit_j ? +1 : -1               red.global.add.s32
//                (BitSetBasedPSLogic.scala:14-25)
//   MinHash    : sig[key][j] = min(sig[key][j], hash_j(tweet) << 32 | tweet)   red.global.min.u64
//                (SendHashPSLogic.scala:17-40: keep the tweet with the smaller hash per slot)
//   Bloom query: est[key] = n(A) + n(B) - n(A u B), n(X) = -m/k ln(1 - |X|/m)   (popcount scan)
//                (BloomPredictPSLogic.scala:47-58, Utils.scala:29-43)
//
// Keys are dense slot ids produced by the host-side interning dictionary; rows live in a ShardTable
// whose 4-byte cells are reinterpreted as u32 / s32 / u64.  hash64 is the same seeded mixer as
// models/sketch/hashing.py::hash64, so host and device sketches agree bit for bit.
#include "fps_common.cuh"

__host__ __device__ __forceinline__ unsigned long long fps_hash64(unsigned long long x,
                                                                  unsigned long long seed) {
  unsigned long long z = x + 0x9E3779B97F4A7C15ull * (seed + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

enum FpsSketchKind : int { SK_BLOOM = 0, SK_TOW = 1, SK_MINHASH = 2 };

template <int KIND>
__global__ void __launch_bounds__(256)
    fps_sketch_update_kernel(const __grid_constant__ ShardTable t, const int* __restrict__ keys,
                             const long long* __restrict__ tweets, long long n, int num_hashes,
                             int array_size) {
  const long long total = n * (long long)num_hashes;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const long long rec = g / num_hashes;
    const int j = (int)(g - rec * num_hashes);
    const unsigned long long tw = (unsigned long long)tweets[rec];
    unsigned int* row = reinterpret_cast<unsigned int*>(fps_row32(t, keys[rec]));
    if (KIND == SK_BLOOM) {
      const unsigned int pos = (unsigned int)(fps_hash64(tw, (unsigned long long)j) %
                                              (unsigned long long)array_size);
      asm volatile("red.relaxed.sys.global.or.b32 [%0], %1;" ::"l"(row + (pos >> 5)),
                   "r"(1u << (pos & 31))
                   : "memory");
    } else if (KIND == SK_TOW) {
      const unsigned long long h = fps_hash64(tw, (unsigned long long)(j >> 6));
      const int d = ((h >> (j & 63)) & 1ull) ? 1 : -1;
      asm volatile("red.relaxed.sys.global.add.s32 [%0], %1;" ::"l"(row + j), "r"(d) : "memory");
    } else {
      const unsigned long long h = fps_hash64(tw, (unsigned long long)j) >> 32;
      const unsigned long long packed = (h << 32) | (tw & 0xFFFFFFFFull);
      asm volatile("red.relaxed.sys.global.min.u64 [%0], %1;" ::"l"(
                       reinterpret_cast<unsigned long long*>(row) + j),
                   "l"(packed)
                   : "memory");
    }
  }
}

extern "C" int fps_sketch_update(const ShardTable* t, int kind, const int* keys,
                                 const long long* tweets, long long n, int num_hashes,
                                 int array_size, int num_sms, cudaStream_t stream) {
  if (n <= 0) return 0;
  long long total = n * (long long)num_hashes;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)num_sms * 16) blocks = (long long)num_sms * 16;
  const int g = (int)blocks;
  if (kind == SK_BLOOM)
    fps_sketch_update_kernel<SK_BLOOM><<<g, 256, 0, stream>>>(*t, keys, tweets, n, num_hashes, array_size);
  else if (kind == SK_TOW)
    fps_sketch_update_kernel<SK_TOW><<<g, 256, 0, stream>>>(*t, keys, tweets, n, num_hashes, array_size);
  else if (kind == SK_MINHASH)
    fps_sketch_update_kernel<SK_MINHASH><<<g, 256, 0, stream>>>(*t, keys, tweets, n, num_hashes, array_size);
  else
    return -1007;
  return (int)cudaGetLastError();
}

// one warp per local key: popcount(B) and popcount(B | Q)
__global__ void __launch_bounds__(256)
    fps_bloom_query_kernel(const unsigned int* __restrict__ local_rows, long long n_rows, int stride_words,
                           int n_words, const unsigned int* __restrict__ query, float m, float k,
                           float* __restrict__ est) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < n_rows; r += n_warps) {
    const unsigned int* row = local_rows + r * (long long)stride_words;
    int cb = 0, cu = 0, cq = 0;
    for (int w = lane; w < n_words; w += 32) {
      const unsigned int b = row[w], q = query[w];
      cb += __popc(b); cu += __popc(b | q); cq += __popc(q);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cb += __shfl_xor_sync(0xffffffffu, cb, o);
      cu += __shfl_xor_sync(0xffffffffu, cu, o);
      cq += __shfl_xor_sync(0xffffffffu, cq, o);
    }
    if (lane == 0) {
      const float nA = -m / k * logf(1.f - (float)cq / m);
      const float nB = -m / k * logf(1.f - (float)cb / m);
      const float nU = -m / k * logf(1.f - (float)cu / m);
      est[r] = (cb == 0) ? -3.0e38f : nA + nB - nU;  // empty rows are not keys
    }
  }
}

extern "C" int fps_bloom_query(const unsigned int* local_rows, long long n_rows, int stride_words,
                               int n_words, const unsigned int* query, float m, float k, float* est,
                               int num_sms, cudaStream_t stream) {
  if (n_rows <= 0) return 0;
  long long blocks = (n_rows * 32 + 255) / 256;
  if (blocks > (long long)num_sms * 8) blocks = (long long)num_sms * 8;
  fps_bloom_query_kernel<<<(int)blocks, 256, 0, stream>>>(local_rows, n_rows, stride_words, n_words,
                                                          query, m, k, est);
  return (int)cudaGetLastError();
}
