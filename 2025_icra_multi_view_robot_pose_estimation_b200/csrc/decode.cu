// decode.cu — belief-map key-point decoder (kernel 1 of the hot path), sm_100a.
//
// Replaces extract_keypoints_from_heatmaps (model/Fr5_model_train.ipynb:4674-4705 and twins)
// and the inline arg-max loops (DIP_REAL.py:116-124, model/MvRoPose_FR3.py:299-304),
// batched over n_maps = B*V*K maps, plus the sub-pixel soft-arg-max the reference lacks.
//
// Roofline: HBM. Every map byte is read from DRAM exactly once (algorithmic bytes per map =
// H*W*sizeof(dtype); outputs are 28 B per map). Design:
//   * one CTA (or one thread-block cluster of S CTAs for maps > 192 KB) per map; each thread
//     streams 16-byte vectors with ld.global.nc.L1::no_allocate, 2 x 4 loads in flight
//     (software-pipelined batches), 4 CTAs / SM -> ~128 KB in flight per SM;
//   * pass 1 does the minimum ALU work per byte: a packed max.NaN tree per 16-byte chunk
//     (bf16x2 / f16x2 SIMD for 16-bit maps), one scalar (max, first-chunk) update per chunk,
//     and (global soft mode only) one 2-byte st.shared of the chunk maximum;
//   * the arg-max index is resolved afterwards by re-reading ONE chunk per thread, then a
//     warp-shuffle / shared-memory / DSMEM (value, index) reduction with torch.argmax's
//     first-maximum, NaN-is-maximal ordering;
//   * pass 2 (soft-arg-max) is exact with respect to the TRUE map maximum: threads scan
//     their chunk maxima in shared memory and re-read from L2 only chunks that can carry a
//     weight >= exp(-32) (a handful per peaked map), so no online-softmax rescaling and no
//     exp per element in the streaming loop (MUFU would cap a bf16 stream at ~75% of HBM).
#include <cooperative_groups.h>
#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace mvgeo {

constexpr int kDecThreads = 256;
constexpr int kDecWarps = kDecThreads / 32;
constexpr int kDecUnroll = 4;
constexpr int kMaxSplits = 8;  // portable cluster size

struct DecodeParams {
  const void* maps;
  int64_t n_maps;
  int H, W;
  int chunks_per_map;  // vector kernel: H*W*sizeof / 16
  int seg_chunks;      // chunks handled by one CTA of the cluster
  int splits;          // CTAs per map (cluster size)
  double scale_x, scale_y;
  float beta_log2e;  // beta * log2(e)
  float skip_delta;  // kSoftSkip / beta
  int radius;
  int apply_sigmoid;
  int64_t k_inner, out_stride, out_offset;
  int32_t* idx;
  float* peak;
  float* score;
  float* kp_hard;
  float* kp_soft;
};

constexpr int kMaxWarps = 9;  // 8 consumer warps + 1 producer warp in the TMA kernel

struct BlockScratch {
  float val[kMaxWarps];
  int idx[kMaxWarps];
  float sum[3][kMaxWarps];
  // per-CTA results, read by cluster peers through DSMEM
  float best_val;
  int best_idx;
  float part[3];
};

// Barrier over the first NW warps of the CTA. BAR == 0 is __syncthreads() (NW must then be every
// warp of the CTA); BAR > 0 is a named barrier, used by the persistent kernel whose producer
// warp never takes part in the per-map reductions.
template <int NW, int BAR>
__device__ __forceinline__ void group_sync() {
  if (BAR == 0) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NW * 32) : "memory");
}

// Group-wide (value, index) arg-max. Result valid in every participating thread.
template <int NW, int BAR>
__device__ __forceinline__ void block_argmax(float& v, int& i, BlockScratch& s) {
  warp_argmax(v, i);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s.val[warp] = v;
    s.idx[warp] = i;
  }
  group_sync<NW, BAR>();
  v = s.val[0];
  i = s.idx[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) {
    if (argmax_better(v, i, s.val[w], s.idx[w])) {
      v = s.val[w];
      i = s.idx[w];
    }
  }
  group_sync<NW, BAR>();
}

// Group-wide fixed-order sums of three accumulators. Result valid in every participating thread.
template <int NW, int BAR>
__device__ __forceinline__ void block_sum3(float& a, float& b, float& c, BlockScratch& s) {
  a = warp_sum(a);
  b = warp_sum(b);
  c = warp_sum(c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s.sum[0][warp] = a;
    s.sum[1][warp] = b;
    s.sum[2][warp] = c;
  }
  group_sync<NW, BAR>();
  a = b = c = 0.f;
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    a += s.sum[0][w];
    b += s.sum[1][w];
    c += s.sum[2][w];
  }
  group_sync<NW, BAR>();
}

__device__ __forceinline__ void write_outputs(const DecodeParams& p, int64_t map, float M, int best, float s,
                                              float sx, float sy, bool have_soft) {
  const int64_t o = (map / p.k_inner) * p.out_stride + p.out_offset + (map % p.k_inner);
  const int py = best / p.W, px = best - py * p.W;
  if (p.idx) p.idx[o] = best;
  if (p.peak) p.peak[o] = M;
  if (p.score) p.score[o] = p.apply_sigmoid ? 1.0f / (1.0f + expf(-M)) : M;
  const float hx = (float)((double)px * p.scale_x), hy = (float)((double)py * p.scale_y);
  if (p.kp_hard) {
    p.kp_hard[2 * o] = hx;
    p.kp_hard[2 * o + 1] = hy;
  }
  if (p.kp_soft) {
    float qx = hx, qy = hy;
    if (have_soft) {
      if (M != M) {
        qx = qy = __int_as_float(0x7fc00000);
      } else if (s > 0.f) {
        qx = (float)(((double)px + (double)sx / (double)s) * p.scale_x);
        qy = (float)(((double)py + (double)sy / (double)s) * p.scale_y);
      }
    }
    p.kp_soft[2 * o] = qx;
    p.kp_soft[2 * o + 1] = qy;
  }
}

// Window soft-arg-max around (px,py): every thread takes window cells tid, tid+256, ...
template <int DT, int NT>
__device__ __forceinline__ void window_accumulate(const DecodeParams& p, const void* map_base, float M, int px,
                                                  int py, float& s, float& sx, float& sy) {
  using E = Elem<DT>;
  const int r = p.radius, side = 2 * r + 1;
  for (int t = threadIdx.x; t < side * side; t += NT) {
    const int dy = t / side - r, dx = t % side - r;
    const int y = py + dy, x = px + dx;
    if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
      const float e = E::load(map_base, (int64_t)y * p.W + x);
      const float w = ex2_approx((e - M) * p.beta_log2e);
      s += w;
      sx += w * (float)dx;
      sy += w * (float)dy;
    }
  }
}

// ----------------------------------------------------------------------------------------
// Fast path: 16-byte aligned maps whose size is a multiple of 16 bytes.
// ----------------------------------------------------------------------------------------
// Per-chunk work of pass 1, shared by both streaming kernels: packed max.NaN tree, 2-byte
// st.shared of the chunk maximum (global soft mode), and the running (max, first chunk) update.
// The maximum is canonicalised (+0.0f turns -0 into +0) so that the update test can be a plain
// bit comparison of max.NaN results: equal values, -0/+0 and NaN/NaN all leave the FIRST chunk.
template <int DT, int MODE>
__device__ __forceinline__ void consume_chunk(const uint4& ch, int c, typename Elem<DT>::carrier* cmax, float& run_max,
                                              int& run_chunk) {
  using E = Elem<DT>;
  float cm;
  typename E::carrier packed;
  E::chunk_max2(ch, cm, packed);
  if (MODE == MVGEO_SOFT_GLOBAL) cmax[c] = packed;
  const float nm = max_nan_f32(run_max, cm + 0.0f);
  run_chunk = (__float_as_uint(nm) != __float_as_uint(run_max)) ? c : run_chunk;
  run_max = nm;
}

// Everything after the streaming pass: resolve the first maximal element, reduce (value, index)
// over the CTA and the cluster, run the soft-arg-max pass, write the outputs. NW warps take part
// (barrier BAR). cmax holds the chunk maxima of this CTA's segment, padded with -inf to a multiple
// of PER entries so that pass 2 can scan PER of them per 16-byte ld.shared.
template <int DT, int MODE, int NW, int BAR>
__device__ __forceinline__ void decode_epilogue(const DecodeParams& p, BlockScratch& sc,
                                                const typename Elem<DT>::carrier* cmax, const uint4* mp, int64_t map,
                                                int rank, int c_begin, int n, float run_max, int run_chunk) {
  using E = Elem<DT>;
  constexpr int PER = E::kPerChunk;
  constexpr int NT = NW * 32;
  const int S = p.splits;
  const int tid = threadIdx.x;
  const uint4* seg = mp + c_begin;

  float my_val = run_max;
  int my_idx = 0x7fffffff;
  if (run_chunk >= 0) {
    const uint4 ch = ld_stream(seg + run_chunk);
    const bool isn = (run_max != run_max);
#pragma unroll
    for (int j = PER - 1; j >= 0; --j) {
      const float e = E::get(ch, j);
      const bool hit = isn ? (e != e) : (e == run_max);
      if (hit) my_idx = (c_begin + run_chunk) * PER + j;
    }
  }
  block_argmax<NW, BAR>(my_val, my_idx, sc);

  cg::cluster_group cluster = cg::this_cluster();
  if (S > 1) {
    if (tid == 0) {
      sc.best_val = my_val;
      sc.best_idx = my_idx;
    }
    cluster.sync();
    float v = my_val;
    int i = my_idx;
    for (int r = 0; r < S; ++r) {
      if (r == rank) continue;
      const BlockScratch* peer = cluster.map_shared_rank(&sc, r);
      const float ov = peer->best_val;
      const int oi = peer->best_idx;
      if (argmax_better(v, i, ov, oi)) {
        v = ov;
        i = oi;
      }
    }
    my_val = v;
    my_idx = i;
  }
  const float M = my_val;
  const int best = my_idx;
  const int py = best / p.W, px = best - py * p.W;

  float s = 0.f, sx = 0.f, sy = 0.f;
  if (MODE == MVGEO_SOFT_GLOBAL) {
    const float thr = M - p.skip_delta;  // NaN peak: every comparison is false, nothing accumulates
    const uint4* cm4 = reinterpret_cast<const uint4*>(cmax);
    const int groups = (n + PER - 1) / PER;
    for (int g = tid; g < groups; g += NT) {
      const uint4 cv = cm4[g];  // PER chunk maxima (carriers have the element type of the map)
      if (E::chunk_max(cv) >= thr) {
#pragma unroll
        for (int j = 0; j < PER; ++j) {
          const int c = g * PER + j;
          if (c < n && E::get(cv, j) >= thr) {
            const uint4 ch = ld_stream(seg + c);
            const int flat0 = (c_begin + c) * PER;
            int y = flat0 / p.W, x = flat0 - y * p.W;
#pragma unroll
            for (int e_ = 0; e_ < PER; ++e_) {
              const float e = E::get(ch, e_);
              const float w = ex2_approx((e - M) * p.beta_log2e);
              s += w;
              sx += w * (float)(x - px);
              sy += w * (float)(y - py);
              if (++x == p.W) {
                x = 0;
                ++y;
              }
            }
          }
        }
      }
    }
  } else if (MODE == MVGEO_SOFT_WINDOW) {
    if (rank == 0) window_accumulate<DT, NT>(p, mp, M, px, py, s, sx, sy);
  }
  if (MODE != MVGEO_SOFT_NONE) {
    block_sum3<NW, BAR>(s, sx, sy, sc);
    if (S > 1 && MODE == MVGEO_SOFT_GLOBAL) {
      if (tid == 0) {
        sc.part[0] = s;
        sc.part[1] = sx;
        sc.part[2] = sy;
      }
      cluster.sync();
      if (rank == 0 && tid == 0) {
        s = sx = sy = 0.f;
        for (int r = 0; r < S; ++r) {  // fixed rank order: deterministic
          const BlockScratch* peer = cluster.map_shared_rank(&sc, r);
          s += peer->part[0];
          sx += peer->part[1];
          sy += peer->part[2];
        }
      }
    }
  }
  if (rank == 0 && tid == 0) write_outputs(p, map, M, best, s, sx, sy, MODE != MVGEO_SOFT_NONE);
  if (S > 1) cluster.sync();  // peers' shared memory must outlive the remote reads above
}

// -inf padding of the chunk-maxima array up to a multiple of PER entries (see decode_epilogue)
template <int DT>
__device__ __forceinline__ void pad_cmax(typename Elem<DT>::carrier* cmax, int n) {
  using E = Elem<DT>;
  const int c = n + (int)threadIdx.x;
  if ((int)threadIdx.x < E::kPerChunk && c < ((n + E::kPerChunk - 1) / E::kPerChunk) * E::kPerChunk)
    cmax[c] = E::pack(__int_as_float(0xff800000));
}

// ---- streaming pass, variant A: register-staged ld.global.nc (kept for A/B measurements) ----
template <int DT, int MODE>
__global__ void __launch_bounds__(kDecThreads, 4) decode_vec_kernel(const DecodeParams p) {
  using E = Elem<DT>;
  using carrier = typename E::carrier;
  __shared__ BlockScratch sc;
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  carrier* cmax = reinterpret_cast<carrier*>(dyn_smem);

  const int S = p.splits;
  const int64_t map = blockIdx.x / S;
  const int rank = (int)(blockIdx.x - map * S);
  const int tid = threadIdx.x;
  const uint4* mp = reinterpret_cast<const uint4*>(p.maps) + map * (int64_t)p.chunks_per_map;
  const int c_begin = rank * p.seg_chunks;
  const int n = max(min(c_begin + p.seg_chunks, p.chunks_per_map) - c_begin, 0);
  const uint4* seg = mp + c_begin;

  float run_max = __int_as_float(0xff800000);  // -inf
  int run_chunk = tid < n ? tid : -1;
  constexpr int kBatch = kDecThreads * kDecUnroll;
  const int full = n / kBatch;
  for (int it = 0; it < full; ++it) {
    uint4 v[kDecUnroll];
#pragma unroll
    for (int u = 0; u < kDecUnroll; ++u) v[u] = ld_stream(seg + it * kBatch + u * kDecThreads + tid);
#pragma unroll
    for (int u = 0; u < kDecUnroll; ++u)
      consume_chunk<DT, MODE>(v[u], it * kBatch + u * kDecThreads + tid, cmax, run_max, run_chunk);
  }
  for (int c = full * kBatch + tid; c < n; c += kDecThreads) {
    const uint4 v = ld_stream(seg + c);
    consume_chunk<DT, MODE>(v, c, cmax, run_max, run_chunk);
  }
  if (MODE == MVGEO_SOFT_GLOBAL) pad_cmax<DT>(cmax, n);
  decode_epilogue<DT, MODE, kDecWarps, 0>(p, sc, cmax, mp, map, rank, c_begin, n, run_max, run_chunk);
}

// ---- streaming pass, variant B: TMA bulk copies into a shared-memory ring -------------------
// One producer lane issues cp.async.bulk (SASS UBLKCP) tiles of 256*U chunks into a STAGES-deep
// ring with full/empty mbarriers per stage; the 8 consumer warps read each tile with
// conflict-free ld.shared.v4. Bytes in flight are set by the ring, not by registers, and the
// consumers carry no global address arithmetic.
// PERSIST (maps that fit one CTA): the grid is (SMs x resident CTAs) and every CTA walks maps
// blockIdx.x, +gridDim.x, ...; the producer keeps filling the ring with the NEXT map's tiles
// while the consumers are in the latency-bound epilogue of the current one (they synchronise on
// a consumer-only named barrier), so the epilogue no longer drains the memory pipeline.
// !PERSIST (maps split over a thread-block cluster): one segment per CTA, whole CTA in the epilogue.
template <int DT, int MODE, int U, int STAGES, bool PERSIST>
__global__ void __launch_bounds__(kDecThreads + 32) decode_tma_kernel(const DecodeParams p) {
  using E = Elem<DT>;
  using carrier = typename E::carrier;
  constexpr int kTile = kDecThreads * U;  // chunks per tile
  __shared__ BlockScratch sc;
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  uint4* ring = reinterpret_cast<uint4*>(dyn_smem);
  carrier* cmax = reinterpret_cast<carrier*>(dyn_smem + (size_t)STAGES * kTile * 16);

  const int S = PERSIST ? 1 : p.splits;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t first = PERSIST ? (int64_t)blockIdx.x : (int64_t)(blockIdx.x / S);
  const int64_t step = PERSIST ? (int64_t)gridDim.x : p.n_maps;  // !PERSIST: exactly one map
  const int rank = PERSIST ? 0 : (int)(blockIdx.x - first * S);
  const int c_begin = rank * p.seg_chunks;
  const int n = max(min(c_begin + p.seg_chunks, p.chunks_per_map) - c_begin, 0);
  const int n_tiles = (n + kTile - 1) / kTile;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kDecWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == kDecWarps) {
    // ------------------------------- producer ---------------------------------------------
    if (lane == 0) {
      int s = 0, k = 0;  // slot, and how many times the ring has wrapped
      for (int64_t map = first; map < p.n_maps; map += step) {
        const uint4* seg = reinterpret_cast<const uint4*>(p.maps) + map * (int64_t)p.chunks_per_map + c_begin;
        for (int t = 0; t < n_tiles; ++t) {
          // before re-using a slot for the k-th time, wait for the consumers' (k-1)-th release of it
          if (k > 0) mbar_wait(&empty_bar[s], (uint32_t)((k - 1) & 1));
          const uint32_t bytes = (uint32_t)min(kTile, n - t * kTile) * 16u;
          mbar_arrive_expect_tx(&full_bar[s], bytes);
          bulk_copy_g2s(ring + s * kTile, seg + (size_t)t * kTile, bytes, &full_bar[s]);
          if (++s == STAGES) {
            s = 0;
            ++k;
          }
        }
      }
    }
    if (PERSIST) return;  // the producer warp takes no part in the per-map reductions
  }

  // --------------------------------- consumers ---------------------------------------------
  int s = 0;
  uint32_t ph = 0;
  for (int64_t map = first; map < p.n_maps; map += step) {
    const uint4* mp = reinterpret_cast<const uint4*>(p.maps) + map * (int64_t)p.chunks_per_map;
    float run_max = __int_as_float(0xff800000);  // -inf
    int run_chunk = -1;
    if (warp < kDecWarps) {
      run_chunk = tid < n ? tid : -1;
      if (MODE == MVGEO_SOFT_GLOBAL) pad_cmax<DT>(cmax, n);
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait(&full_bar[s], ph);
        const uint4* tile = ring + s * kTile;
        const int base = t * kTile;
        if (base + kTile <= n) {
          uint4 v[U];
#pragma unroll
          for (int u = 0; u < U; ++u) v[u] = tile[u * kDecThreads + tid];
#pragma unroll
          for (int u = 0; u < U; ++u)
            consume_chunk<DT, MODE>(v[u], base + u * kDecThreads + tid, cmax, run_max, run_chunk);
        } else {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int c = base + u * kDecThreads + tid;
            if (c < n) consume_chunk<DT, MODE>(tile[u * kDecThreads + tid], c, cmax, run_max, run_chunk);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
        if (++s == STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
    }
    if (PERSIST)
      decode_epilogue<DT, MODE, kDecWarps, 1>(p, sc, cmax, mp, map, rank, c_begin, n, run_max, run_chunk);
    else
      decode_epilogue<DT, MODE, kDecWarps + 1, 0>(p, sc, cmax, mp, map, rank, c_begin, n, run_max, run_chunk);
  }
}

// ----------------------------------------------------------------------------------------
// Generic path: any H, W, alignment. One CTA per map, element-wise loads.
// ----------------------------------------------------------------------------------------
template <int DT, int MODE>
__global__ void __launch_bounds__(kDecThreads) decode_scalar_kernel(const DecodeParams p) {
  using E = Elem<DT>;
  __shared__ BlockScratch sc;
  const int64_t map = blockIdx.x;
  const int tid = threadIdx.x;
  const int n = p.H * p.W;
  const char* base = reinterpret_cast<const char*>(p.maps) + map * (int64_t)n * E::kBytes;

  float my_val = __int_as_float(0xff800000);
  int my_idx = tid < n ? tid : 0x7fffffff;
  for (int i = tid; i < n; i += kDecThreads) {
    const float e = E::load(base, i);
    const bool gt = (e > my_val) || ((e != e) && (my_val == my_val));
    if (gt) {
      my_val = e;
      my_idx = i;
    }
  }
  block_argmax<kDecWarps, 0>(my_val, my_idx, sc);
  const float M = my_val;
  const int best = my_idx;
  const int py = best / p.W, px = best - py * p.W;
  float s = 0.f, sx = 0.f, sy = 0.f;
  if (MODE == MVGEO_SOFT_GLOBAL) {
    const float thr = M - p.skip_delta;
    for (int i = tid; i < n; i += kDecThreads) {
      const float e = E::load(base, i);
      if (e >= thr) {
        const int y = i / p.W, x = i - y * p.W;
        const float w = ex2_approx((e - M) * p.beta_log2e);
        s += w;
        sx += w * (float)(x - px);
        sy += w * (float)(y - py);
      }
    }
  } else if (MODE == MVGEO_SOFT_WINDOW) {
    window_accumulate<DT, kDecThreads>(p, base, M, px, py, s, sx, sy);
  }
  if (MODE != MVGEO_SOFT_NONE) block_sum3<kDecWarps, 0>(s, sx, sy, sc);
  if (tid == 0) write_outputs(p, map, M, best, s, sx, sy, MODE != MVGEO_SOFT_NONE);
}

// Streaming-kernel configuration: tile = 256*kTmaU chunks, kTmaStages-deep ring.
#ifndef MVGEO_TMA_U
#define MVGEO_TMA_U 2
#endif
#ifndef MVGEO_TMA_STAGES
#define MVGEO_TMA_STAGES 4
#endif
constexpr int kTmaU = MVGEO_TMA_U;
constexpr int kTmaStages = MVGEO_TMA_STAGES;

template <typename K>
static int launch_clustered(K kern, const DecodeParams& p, unsigned grid, int threads, size_t smem, cudaStream_t st) {
  if (smem > 40 * 1024) MVGEO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)p.splits;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MVGEO_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  return MVGEO_OK;
}

template <int DT, int MODE>
static int launch_decode(const DecodeParams& p, bool vec, size_t cmax_bytes, int variant, cudaStream_t st) {
  if (!vec) {
    decode_scalar_kernel<DT, MODE><<<(unsigned)p.n_maps, kDecThreads, 0, st>>>(p);
    MVGEO_CHECK_LAUNCH();
    return MVGEO_OK;
  }
  const unsigned one_per_segment = (unsigned)(p.n_maps * p.splits);
  if (variant == 0)
    return launch_clustered(decode_vec_kernel<DT, MODE>, p, one_per_segment, kDecThreads, cmax_bytes, st);
  const size_t smem = (size_t)kTmaStages * kTmaU * kDecThreads * 16 + cmax_bytes;
  if (p.splits > 1)
    return launch_clustered(decode_tma_kernel<DT, MODE, kTmaU, kTmaStages, false>, p, one_per_segment,
                            kDecThreads + 32, smem, st);
  // persistent: one resident wave of CTAs, each walking maps blockIdx.x, +gridDim.x, ...
  auto kern = decode_tma_kernel<DT, MODE, kTmaU, kTmaStages, true>;
  if (smem > 40 * 1024) MVGEO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0, per_sm = 0;
  MVGEO_CUDA(cudaGetDevice(&dev));
  MVGEO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  MVGEO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kDecThreads + 32, smem));
  if (per_sm < 1) return MVGEO_EUNSUPPORTED;
  const int64_t resident = (int64_t)sms * per_sm;
  const unsigned grid = (unsigned)(p.n_maps < resident ? p.n_maps : resident);
  return launch_clustered(kern, p, grid, kDecThreads + 32, smem, st);
}

template <int DT>
static int dispatch_mode(const DecodeParams& p, int mode, bool vec, size_t smem, int variant, cudaStream_t st) {
  switch (mode) {
    case MVGEO_SOFT_NONE: return launch_decode<DT, MVGEO_SOFT_NONE>(p, vec, 0, variant, st);
    case MVGEO_SOFT_GLOBAL: return launch_decode<DT, MVGEO_SOFT_GLOBAL>(p, vec, smem, variant, st);
    case MVGEO_SOFT_WINDOW: return launch_decode<DT, MVGEO_SOFT_WINDOW>(p, vec, 0, variant, st);
  }
  return MVGEO_EINVAL;
}

}  // namespace mvgeo

using namespace mvgeo;

extern "C" int mvgeo_decode(const void* maps, int dtype, int64_t n_maps, int H, int W, double scale_x, double scale_y,
                            int soft_mode, float beta, int window_radius, int apply_sigmoid, int64_t k_inner,
                            int64_t out_stride, int64_t out_offset, int32_t* idx, float* peak, float* score,
                            float* kp_hard, float* kp_soft, void* stream) {
  if (n_maps < 0 || H <= 0 || W <= 0 || k_inner <= 0 || out_stride < 0 || out_offset < 0) return MVGEO_EINVAL;
  if ((int64_t)H * W > (int64_t)1 << 30) return MVGEO_EINVAL;
  if (dtype != MVGEO_F32 && dtype != MVGEO_BF16 && dtype != MVGEO_F16) return MVGEO_EINVAL;
  if (soft_mode < MVGEO_SOFT_NONE || soft_mode > MVGEO_SOFT_WINDOW) return MVGEO_EINVAL;
  if (soft_mode != MVGEO_SOFT_NONE && !(beta > 0.f)) return MVGEO_EINVAL;
  if (soft_mode == MVGEO_SOFT_WINDOW && (window_radius < 0 || window_radius > MVGEO_MAX_WINDOW_RADIUS)) return MVGEO_EINVAL;
  if (n_maps == 0) return MVGEO_OK;
  if (!maps) return MVGEO_ENULL;
  if (n_maps > (int64_t)0x7fffffff / kMaxSplits) return MVGEO_EINVAL;

  const int esize = dtype == MVGEO_F32 ? 4 : 2;
  const int64_t map_bytes = (int64_t)H * W * esize;
  const bool vec = (map_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(maps) & 15) == 0);

  DecodeParams p;
  p.maps = maps;
  p.n_maps = n_maps;
  p.H = H;
  p.W = W;
  p.scale_x = scale_x;
  p.scale_y = scale_y;
  p.beta_log2e = beta * kLog2e;
  p.skip_delta = soft_mode == MVGEO_SOFT_GLOBAL ? kSoftSkip / beta : 0.f;
  p.radius = window_radius;
  p.apply_sigmoid = apply_sigmoid;
  p.k_inner = k_inner;
  p.out_stride = out_stride;
  p.out_offset = out_offset;
  p.idx = idx;
  p.peak = peak;
  p.score = score;
  p.kp_hard = kp_hard;
  p.kp_soft = kp_soft;
  p.chunks_per_map = 0;
  p.seg_chunks = 0;
  p.splits = 1;
  size_t smem = 0;
  if (vec) {
    // The split count depends on the map size only (never on n_maps), so results are
    // bit-identical however the frames are sharded across GPUs.
    const int64_t chunks = map_bytes / 16;
    int splits = 1;
    if (map_bytes > 192 * 1024) splits = (int)min((int64_t)kMaxSplits, (map_bytes + 160 * 1024 - 1) / (160 * 1024));
    p.chunks_per_map = (int)chunks;
    p.seg_chunks = (int)((chunks + splits - 1) / splits);
    p.splits = splits;
    if (soft_mode == MVGEO_SOFT_GLOBAL) {
      smem = (size_t)((p.seg_chunks + 7) / 8 * 8) * (dtype == MVGEO_F32 ? 4 : 2);  // padded for the 16-byte scan
      if (smem > 150 * 1024) return MVGEO_EUNSUPPORTED;  // maps beyond ~6.4 MB (f32): use the window mode
    }
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // A/B switch for kernel development only (MVGEO_DECODE_VARIANT=0 selects the register-staged
  // ld.global variant); both variants produce bit-identical results.
  const char* ev = getenv("MVGEO_DECODE_VARIANT");
  const int variant = (ev && ev[0] == '0') ? 0 : 1;
  switch (dtype) {
    case MVGEO_F32: return dispatch_mode<MVGEO_F32>(p, soft_mode, vec, smem, variant, st);
    case MVGEO_BF16: return dispatch_mode<MVGEO_BF16>(p, soft_mode, vec, smem, variant, st);
    case MVGEO_F16: return dispatch_mode<MVGEO_F16>(p, soft_mode, vec, smem, variant, st);
  }
  return MVGEO_EINVAL;
}
