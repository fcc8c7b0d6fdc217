// K6: pull (query vectors from the PS shards over NVLink) fused with a tcgen05 GEMM against the
// worker-local item table and a top-K candidate-filter epilogue.  sm_100a only.
//
//   scores[q, i] = <query[q, :], item[i, :]>        (TF32 tensor-core MMA, FP32 accumulate in TMEM)
//
//   A operand (queries, 128 rows/CTA): gathered ONCE per CTA from the owning PS shards with 16-byte
//       peer loads and written into shared memory in the canonical K-major SWIZZLE_128B layout
//       (manual XOR swizzle), then published to the async proxy with fence.proxy.async.
//   B operand (items): streamed by TMA (cp.async.bulk.tensor.2d, SASS UTMALDG) through a 4-stage
//       mbarrier ring, hardware-swizzled by the tensor map.
//   MMA: one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (SASS UTCHMMA-family), M=128,
//       N=128, K=8 per instruction; two 128-column accumulators in TMEM are double buffered so the
//       epilogue of tile t overlaps the MMAs of tile t+1 (tcgen05.commit -> mbarrier).
//   Epilogue (4 warps, one TMEM lane = one query row per thread, tcgen05.ld 32x32b):
//       mode 0: write raw scores                    (small problems / validation)
//       mode 1: per-(row, tile) maximum             (pass 1: gives an exact top-K lower bound theta)
//       mode 2: append (score, item) >= theta[row]  (pass 2: exact candidate set, usually << N)
//
// Exactness: theta[row] = K-th largest tile-maximum of pass 1 is attained by K distinct items, so the
// true K-th best score >= theta and pass 2 (bitwise identical scores) keeps every top-K item.
// This replaces the pointer-chasing LEMP scan (PSTopKGeneratorWorker.scala:49-113) with tile-level
// pruning that is validated against brute force (tests/test_gpu_topk.py).
#include <cuda.h>
#include "fps_common.cuh"

#define TK_M 128          // query rows per CTA (UMMA M)
#define TK_N 128          // items per tile (UMMA N)
#define TK_KB_FLOATS 32   // floats per 128-byte swizzle atom row
#define TK_MAX_STAGES 4
#define TK_THREADS 256
#define TK_MAX_KB 4       // dim <= 128 (A block + >= 2 B stages must fit in 227 KB of smem)

struct TopkArgs {
  const void* q_ids;        // [n_queries] ids of the query vectors in q_tab (null -> q_local)
  const float* q_local;     // [n_queries, stride] already-local queries (used when q_ids == null)
  ShardTable q_tab;         // PS table of the query vectors (users)
  int n_queries;
  int n_items;              // rows of the local item table
  int stride;               // floats per row (multiple of 4)
  int n_tiles;              // ceil(n_items / TK_N)
  int tiles_per_split;
  int n_splits;
  int mode;
  int n_stages;             // B-operand ring depth (2..4, from the smem budget)
  float* out_scores;        // mode 0: [n_queries, out_ld]
  long long out_ld;
  float* tile_max;          // mode 1: [n_queries, n_tiles]
  const float* theta;       // mode 2: [n_queries]
  int* cand_count;          // mode 2: [n_queries, n_splits] candidates found by each split (may exceed seg_cap)
  float* cand_score;        // mode 2: [n_queries, cand_cap], split s owns columns [s*seg_cap, (s+1)*seg_cap)
  int* cand_item;           // mode 2: [n_queries, cand_cap]
  int cand_cap;
  int seg_cap;              // cand_cap / n_splits (set by the launcher)
  int tile_lo;              // first tile to score
  int pad_;
  const int* tile_limit;    // optional device scalar: score only tiles < min(n_tiles, *tile_limit)
                            // (length-pruning bound computed on the device, no host round trip)
};

__device__ __forceinline__ uint32_t tk_smem(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void tk_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tk_smem(bar)), "r"(count));
}
__device__ __forceinline__ void tk_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tk_smem(bar)) : "memory");
}
__device__ __forceinline__ void tk_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tk_smem(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tk_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TK_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TK_DONE;\n"
      "bra TK_WAIT;\n"
      "TK_DONE:\n"
      "}\n" ::"r"(tk_smem(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tk_tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1,
                                               uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];" ::"r"(tk_smem(dst)),
      "l"(map), "r"(tk_smem(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (SmemDescriptor, version 1 = sm_100)
__device__ __forceinline__ uint64_t tk_desc(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;  // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;             // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;             // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tk_mma_tf32(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_c),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tk_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   tk_smem(bar))
               : "memory");
}
__device__ __forceinline__ void tk_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// MB = number of 128-row query blocks per CTA.  With MB = 2 every item tile fetched by TMA feeds two
// UMMA_M=128 accumulators (256 query rows), which halves the L2 -> SM item traffic per FLOP; the CTA
// then has 8 epilogue warps (warps 4..11) and uses all 512 TMEM columns (2 blocks x 2 buffers x 128).
template <typename IdT, int MODE, int MB>
__global__ void __launch_bounds__(128 + 128 * MB, 1)
    fps_topk_mma_kernel(const __grid_constant__ CUtensorMap item_map,
                        const __grid_constant__ TopkArgs a) {
  extern __shared__ __align__(1024) unsigned char tk_smem_raw[];
  const int KB = (a.stride + TK_KB_FLOATS - 1) / TK_KB_FLOATS;  // 128-byte K blocks
  const uint32_t kb_bytes = TK_M * 128;                         // one K block of a 128-row tile
  unsigned char* sA = tk_smem_raw;                              // [MB][KB][128 rows][128 B]
  unsigned char* sB = sA + (size_t)MB * KB * kb_bytes;          // [STAGES][KB][128 rows][128 B]
  const int NS = a.n_stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)NS * KB * kb_bytes);
  uint64_t* full = bars;                  // [STAGES] TMA -> MMA
  uint64_t* empty = bars + TK_MAX_STAGES;     // [STAGES] MMA -> TMA
  uint64_t* tfull = bars + 2 * TK_MAX_STAGES;   // [2] MMA -> epilogue
  uint64_t* tempty = tfull + 2;             // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qb = blockIdx.x / a.n_splits;
  const int split = blockIdx.x - qb * a.n_splits;
  // tiles are dealt round-robin to the splits (tile = tile_lo + split + i * n_splits): a device-side
  // tile limit or a hot region of the table (length-sorted items) then stays balanced over the CTAs
  int tiles_hi = a.n_tiles;
  if (a.tile_limit != nullptr) tiles_hi = min(tiles_hi, *a.tile_limit);
  const int first_tile = a.tile_lo + split;
  const int my_tiles = first_tile < tiles_hi ? (tiles_hi - first_tile + a.n_splits - 1) / a.n_splits : 0;
  const int row0 = qb * TK_M * MB;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      tk_mbar_init(&full[s], 1);
      tk_mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tk_mbar_init(&tfull[s], 1);
      tk_mbar_init(&tempty[s], 4 * MB);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {  // TMEM: 2 accumulators x 128 fp32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     tk_smem(tmem_slot)),
                 "r"(256 * MB));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }

  // ---- A operand: pull the 128 query rows (peer loads) into swizzled smem ------------------
  {
    const IdT* qids = reinterpret_cast<const IdT*>(a.q_ids);
    const int nvec = a.stride >> 2;
    const int chunks_per_row = KB * 8;
    for (int t = threadIdx.x; t < MB * TK_M * chunks_per_row; t += blockDim.x) {
      const int r = t / chunks_per_row;
      const int cc = t - r * chunks_per_row;  // 16-byte chunk index along K
      const int kb = cc >> 3, c = cc & 7;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int row = row0 + r;
      if (row < a.n_queries && cc < nvec) {
        const float* src = (qids != nullptr) ? fps_row_t<IdT>(a.q_tab, qids[row])
                                             : a.q_local + (size_t)row * a.stride;
        v = fps_ld_row4(src + 4 * cc);  // the PULL (local HBM or NVLink peer)
      }
      const int mb = r >> 7, rr = r & 127;
      unsigned char* dst = sA + ((size_t)mb * KB + kb) * kb_bytes + (size_t)rr * 128 + ((c ^ (rr & 7)) << 4);
      *reinterpret_cast<float4*>(dst) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> async proxy
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer (items) ===============================
    if (lane == 0) {
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i % NS;
        const uint32_t ph = (uint32_t)((i / NS) & 1);
        tk_mbar_wait(&empty[s], ph ^ 1u);
        tk_mbar_expect_tx(&full[s], (uint32_t)KB * kb_bytes);
        const int item0 = (first_tile + i * a.n_splits) * TK_N;
        for (int kb = 0; kb < KB; ++kb)
          tk_tma_load_2d(sB + ((size_t)s * KB + kb) * kb_bytes, &item_map, kb * TK_KB_FLOATS, item0,
                         &full[s]);
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // instruction descriptor: D=F32, A=B=TF32, K-major x K-major, N=128, M=128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TK_N >> 3) << 17) |
                           ((uint32_t)(TK_M >> 4) << 24);
    if (lane == 0) {
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i % NS;
        const uint32_t ph = (uint32_t)((i / NS) & 1);
        const int acc = i & 1;
        const uint32_t aph = (uint32_t)((i >> 1) & 1);
        tk_mbar_wait(&tempty[acc], aph ^ 1u);  // epilogue drained this accumulator
        tk_mbar_wait(&full[s], ph);            // TMA landed the item tile
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * MB + mb) * TK_N;
          for (int kb = 0; kb < KB; ++kb) {
            const uint32_t a_addr = tk_smem(sA + ((size_t)mb * KB + kb) * kb_bytes);
            const uint32_t b_addr = tk_smem(sB + ((size_t)s * KB + kb) * kb_bytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // UMMA_K = 8 tf32 = 32 bytes inside the 128-byte atom
              tk_mma_tf32(d_tmem, tk_desc(a_addr + k * 32), tk_desc(b_addr + k * 32), idesc,
                          (kb | k) != 0 ? 1u : 0u);
            }
          }
        }
        tk_commit(&empty[s]);     // smem stage reusable once these MMAs retire
        tk_commit(&tfull[acc]);   // accumulator ready for the epilogue
      }
    }
  } else if (warp >= 4) {
    // =============================== epilogue ===============================
    const int ew = warp & 3;                    // TMEM lane quadrant this warp may access
    const int emb = (warp - 4) >> 2;            // which 128-row query block this warp drains
    const int r = emb * TK_M + ew * 32 + lane;  // row inside the CTA's query rows
    const int row = row0 + r;
    const bool row_ok = row < a.n_queries;
    const float th = (MODE == 2 && row_ok) ? a.theta[row] : 0.f;
    int n_cand = 0;
    float* seg_s = nullptr;
    int* seg_i = nullptr;
    if (MODE == 2 && row_ok) {
      seg_s = a.cand_score + (size_t)row * a.cand_cap + (size_t)split * a.seg_cap;
      seg_i = a.cand_item + (size_t)row * a.cand_cap + (size_t)split * a.seg_cap;
    }
    for (int i = 0; i < my_tiles; ++i) {
      const int acc = i & 1;
      const uint32_t aph = (uint32_t)((i >> 1) & 1);
      const int tile = first_tile + i * a.n_splits;
      const int item0 = tile * TK_N;
      const bool full_tile = item0 + TK_N <= a.n_items;   // no per-element bound check needed
      tk_mbar_wait(&tfull[acc], aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float tmax = -3.0e38f;
#pragma unroll 1
      for (int c0 = 0; c0 < TK_N; c0 += 32) {
        uint32_t v[32];
        __syncwarp();  // tcgen05.ld is .sync.aligned: the whole warp must be converged here
        tk_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)((acc * MB + emb) * TK_N + c0), v);
        if (row_ok) {
        if (MODE == 1) {
          if (full_tile) {
#pragma unroll
            for (int j = 0; j < 32; ++j) tmax = fmaxf(tmax, __uint_as_float(v[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (item0 + c0 + j < a.n_items) tmax = fmaxf(tmax, __uint_as_float(v[j]));
          }
        } else if (MODE == 2) {
          // candidates are rare: first a branch-free "any >= theta" test over the 32 columns.  The
          // thread owns (row, split) for the whole kernel, so candidates go to a private segment of the
          // row's buffer with a register cursor: no atomics, no memory round trip in the epilogue
          // (a returning atomic per candidate / per chunk made pass 2 latency bound).
          // The test is hierarchical (4 groups of 8 columns): when one lane of the warp has a candidate
          // the whole warp walks the predicated append loop, so the loop is kept to the 8 columns of
          // the group that actually reached theta instead of all 32.
          float gmax[4];
#pragma unroll
          for (int gi = 0; gi < 4; ++gi) {
            float m = __uint_as_float(v[8 * gi]);
#pragma unroll
            for (int j = 1; j < 8; ++j) m = fmaxf(m, __uint_as_float(v[8 * gi + j]));
            gmax[gi] = m;
          }
          if (fmaxf(fmaxf(gmax[0], gmax[1]), fmaxf(gmax[2], gmax[3])) >= th) {
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) {
              if (gmax[gi] >= th) {
#pragma unroll
                for (int j = 8 * gi; j < 8 * gi + 8; ++j) {
                  const int item = item0 + c0 + j;
                  const float sc = __uint_as_float(v[j]);
                  if (sc >= th && (full_tile || item < a.n_items)) {
                    if (n_cand < a.seg_cap) {
                      seg_s[n_cand] = sc;
                      seg_i[n_cand] = item;
                    }
                    ++n_cand;
                  }
                }
              }
            }
          }
        } else {
          float* out = a.out_scores + (size_t)row * a.out_ld + item0 + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (full_tile || item0 + c0 + j < a.n_items) out[j] = __uint_as_float(v[j]);
        }
        }  // row_ok
      }
      if (MODE == 1 && row_ok) a.tile_max[(size_t)row * a.n_tiles + tile] = tmax;
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) tk_mbar_arrive(&tempty[acc]);
    }
    if (MODE == 2 && row_ok) a.cand_count[(size_t)row * a.n_splits + split] = n_cand;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256 * MB));
  }
}

// ---- host side ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

extern "C" int fps_topk_mma(TopkArgs* args_in, const float* item_table, int id_bytes,
                            int num_sms, cudaStream_t stream) {
  TopkArgs a = *args_in;
  if (a.n_queries <= 0 || a.n_items <= 0) return 0;
  const int KB = (a.stride + TK_KB_FLOATS - 1) / TK_KB_FLOATS;
  if (KB > TK_MAX_KB) return -1003;
  PFN_encodeTiled enc = get_encode();
  if (enc == nullptr) return -1004;
  CUtensorMap map;
  cuuint64_t gdim[2] = {(cuuint64_t)a.stride, (cuuint64_t)a.n_items};
  cuuint64_t gstr[1] = {(cuuint64_t)a.stride * 4};
  cuuint32_t box[2] = {TK_KB_FLOATS, TK_N};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)item_table, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return -1005;
  a.n_tiles = (a.n_items + TK_N - 1) / TK_N;
  const int MBv = (a.n_queries > TK_M && KB <= 2) ? 2 : 1;   // 2 query blocks per CTA when smem allows
  const int qblocks = (a.n_queries + TK_M * MBv - 1) / (TK_M * MBv);
  int splits = num_sms / qblocks;  // one wave of CTAs (1 CTA/SM: smem bound), no tail wave
  int span = a.n_tiles - a.tile_lo;  // tiles that may be scored (a device tile_limit can only lower it)
  if (span < 1) span = 1;
  if (splits > span) splits = span;
  if (splits < 1) splits = 1;
  a.tiles_per_split = (span + splits - 1) / splits;
  a.n_splits = (span + a.tiles_per_split - 1) / a.tiles_per_split;
  a.seg_cap = a.cand_cap / a.n_splits;
  args_in->n_splits = a.n_splits;   // the caller sizes cand_count [n_queries, n_splits] from these
  args_in->seg_cap = a.seg_cap;
  args_in->n_tiles = a.n_tiles;
  if (a.mode < 0) return 0;         // geometry query only
  if (a.mode == 2 && a.seg_cap < 1) return -1006;
  const size_t blk = (size_t)KB * TK_M * 128;
  int stages = (int)((220 * 1024 - blk * MBv - 2048) / blk);
  if (stages > TK_MAX_STAGES) stages = TK_MAX_STAGES;
  if (stages < 2) return -1003;
  a.n_stages = stages;
  const size_t smem = blk * (MBv + stages) + 16 * 8 + 16 + 1024;
  const int grid = qblocks * a.n_splits;
#define TK_LAUNCH2(IDT, MODE, MBT)                                                                 \
  do {                                                                                             \
    cudaError_t e = cudaFuncSetAttribute(fps_topk_mma_kernel<IDT, MODE, MBT>,                      \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
    if (e != cudaSuccess) return (int)e;                                                           \
    fps_topk_mma_kernel<IDT, MODE, MBT><<<grid, 128 + 128 * MBT, smem, stream>>>(map, a);          \
  } while (0)
#define TK_LAUNCH(IDT, MODE)                                                                       \
  do {                                                                                             \
    if (MBv == 2) TK_LAUNCH2(IDT, MODE, 2); else TK_LAUNCH2(IDT, MODE, 1);                         \
  } while (0)
  if (id_bytes == 8) {
    if (a.mode == 0) TK_LAUNCH(long long, 0); else if (a.mode == 1) TK_LAUNCH(long long, 1); else TK_LAUNCH(long long, 2);
  } else {
    if (a.mode == 0) TK_LAUNCH(int, 0); else if (a.mode == 1) TK_LAUNCH(int, 1); else TK_LAUNCH(int, 2);
  }
#undef TK_LAUNCH2
#undef TK_LAUNCH
  return (int)cudaGetLastError();
}
