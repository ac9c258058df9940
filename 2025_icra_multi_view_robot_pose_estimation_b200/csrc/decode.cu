// decode.cu — belief-map key-point decoder (kernel 1 of the hot path), sm_100a.
//
// Replaces extract_keypoints_from_heatmaps (model/Fr5_model_train.ipynb:4674-4705 and twins)
// and the inline arg-max loops (DIP_REAL.py:116-124, model/MvRoPose_FR3.py:299-304),
// batched over n_maps = B*V*K maps, plus the sub-pixel soft-arg-max the reference lacks.
//
// Roofline: HBM. Every map byte is read from DRAM exactly once (algorithmic bytes per map =
// H*W*sizeof(dtype); outputs are 28 B per map). Design:
//   * PERSISTENT kernel: (SMs x resident CTAs) CTAs, each walking maps blockIdx.x, +gridDim.x, ...
//     One elected producer lane per map stream issues 1-D TMA bulk copies (cp.async.bulk, SASS UBLKCP) of 8 KB
//     tiles into a 4-stage shared-memory ring with full/empty mbarriers; 8 consumer warps read
//     the tiles with conflict-free ld.shared.v4. Bytes in flight are set by the ring, not by
//     registers, and the producer keeps prefetching the NEXT map while the consumers are in the
//     latency-bound per-map epilogue (consumer-only named barrier), so the DRAM pipe never drains;
//   * pass 1 does the minimum ALU work per byte: a packed max.NaN tree per 16-byte chunk
//     (bf16x2 / f16x2 SIMD for 16-bit maps), ONE running-maximum update per thread per tile
//     (a "slice" = the 2 chunks a thread owns in a tile), and (global soft mode only) one 2-byte
//     st.shared of the slice maximum;
//   * the raw chunks of the best slice stay in registers, so the first maximal element is
//     resolved without touching memory again; then a warp-shuffle / shared-memory (/ DSMEM)
//     (value, index) reduction with torch.argmax's first-maximum, NaN-is-maximal ordering;
//   * maps up to 112 KB run 4 independent map streams per CTA (consumer groups of 2 warps, each
//     with its own ring, mbarriers, named barrier and producer warp) so that epilogues overlap;
//   * pass 2 (soft-arg-max) is exact with respect to the TRUE map maximum: threads scan the
//     slice maxima in shared memory, 8 per ld.shared.v4, and re-read from L2 only slices that can
//     carry a weight >= exp(-32) (a handful per peaked map), so there is no online-softmax
//     rescaling and no exp per element in the streaming loop (MUFU would cap a bf16 stream at
//     ~75% of HBM). A flat map (nothing can be skipped) re-reads itself: bounded 2x, data-dependent;
//   * maps whose slice-maxima table cannot fit one CTA (> ~2.4 MB) are split over a thread-block
//     cluster of <= 8 CTAs and combined through distributed shared memory.
#include <cooperative_groups.h>

#include <atomic>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace mvgeo {

constexpr int kDecThreads = 256;
constexpr int kDecWarps = kDecThreads / 32;
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

// Barrier over a group of NW warps. bar == 0 is __syncthreads() (the group must then be the whole
// CTA); bar > 0 is a named barrier, used by the persistent kernel whose consumer warp groups
// reduce independently of one another and of the producer warp.
template <int NW>
__device__ __forceinline__ void group_sync(int bar) {
  // literal barrier ids so that ptxas reserves 5 hardware barriers per CTA, not all 16
  switch (bar) {
    case 0: __syncthreads(); break;
    case 1: asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory"); break;
    case 2: asm volatile("bar.sync 2, %0;" ::"n"(NW * 32) : "memory"); break;
    case 3: asm volatile("bar.sync 3, %0;" ::"n"(NW * 32) : "memory"); break;
    default: asm volatile("bar.sync 4, %0;" ::"n"(NW * 32) : "memory"); break;
  }
}

// Group-wide (value, index) arg-max; `lw` is the warp's index inside the group. Result valid in
// every participating thread.
template <int NW>
__device__ __forceinline__ void block_argmax(float& v, int& i, BlockScratch& s, int bar, int lw) {
  warp_argmax(v, i);
  if ((threadIdx.x & 31) == 0) {
    s.val[lw] = v;
    s.idx[lw] = i;
  }
  group_sync<NW>(bar);
  v = s.val[0];
  i = s.idx[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) {
    if (argmax_better(v, i, s.val[w], s.idx[w])) {
      v = s.val[w];
      i = s.idx[w];
    }
  }
  group_sync<NW>(bar);
}

// Group-wide fixed-order sums of three accumulators. Result valid in every participating thread.
template <int NW>
__device__ __forceinline__ void block_sum3(float& a, float& b, float& c, BlockScratch& s, int bar, int lw) {
  a = warp_sum(a);
  b = warp_sum(b);
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) {
    s.sum[0][lw] = a;
    s.sum[1][lw] = b;
    s.sum[2][lw] = c;
  }
  group_sync<NW>(bar);
  a = b = c = 0.f;
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    a += s.sum[0][w];
    b += s.sum[1][w];
    c += s.sum[2][w];
  }
  group_sync<NW>(bar);
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
                                                  int py, int t0, float& s, float& sx, float& sy) {
  using E = Elem<DT>;
  const int r = p.radius, side = 2 * r + 1;
  for (int t = t0; t < side * side; t += NT) {
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
// A "slice" is the U chunks {(t*U + u)*NT + gtid, u < U} that one consumer thread owns in tile t.
// Slice maxima ("carriers") are 16-bit: the map's own type for bf16 / f16 maps (exact), bf16
// rounded UP for f32 maps (a conservative filter: a slice is re-read whenever it might matter,
// and every re-read element is weighted exactly).
template <int DT> struct Carrier { using E = Elem<DT>; };
template <> struct Carrier<MVGEO_F32> { using E = Elem<MVGEO_BF16>; };

template <int DT>
__device__ __forceinline__ uint16_t to_carrier(float m) {
  if (DT == MVGEO_F32) return __bfloat16_as_ushort(__float2bfloat16_ru(m));
  return Elem<DT>::pack(m);
}

// Everything after the streaming pass, for one group: NW warps take part in the reductions
// (NT = 32 NW threads, barrier `bar`, `gt` = thread index in the group) of which the first NC
// threads are the consumers that streamed the map (NC == NT in the persistent kernel; the
// cluster kernel's producer warp joins the reductions with neutral values, NT = NC + 32).
// Resolve the first maximal element inside the winning slice, reduce (value, index) over the group
// and the cluster, run the soft-arg-max pass, write the outputs.
// smax[t*NC + c] = maximum of consumer c's slice of tile t, i.e. of chunks {(t*U + u)*NC + c}.
template <int DT, int MODE, int NW, int NC, int U>
__device__ __forceinline__ void decode_epilogue(const DecodeParams& p, BlockScratch& sc, const uint16_t* smax,
                                                const uint4* mp, int64_t map, int rank, int c_begin, int n,
                                                int n_tiles, float run_max, int run_tile, const uint4 (&run_v)[U],
                                                int bar, int gt) {
  using E = Elem<DT>;
  using CE = typename Carrier<DT>::E;
  constexpr int PER = E::kPerChunk;
  constexpr int NT = NW * 32;
  const int S = p.splits;
  const int lw = gt >> 5;
  const uint4* seg = mp + c_begin;

  // first maximal element inside the winning slice, straight from the registers that kept it
  // (no global re-read on the latency-critical path). Descending loops: the lowest index sticks.
  float my_val = run_max;
  int my_idx = 0x7fffffff;
  if (run_tile >= 0) {
    const bool isn = (run_max != run_max);
#pragma unroll
    for (int u = U - 1; u >= 0; --u) {
      const int c = (run_tile * U + u) * NC + gt;
#pragma unroll
      for (int j = PER - 1; j >= 0; --j) {
        const float e = E::get(run_v[u], j);
        const bool hit = (c < n) && (isn ? (e != e) : (e == run_max));
        if (hit) my_idx = (c_begin + c) * PER + j;
      }
    }
  }
  block_argmax<NW>(my_val, my_idx, sc, bar, lw);

  cg::cluster_group cluster = cg::this_cluster();
  if (S > 1) {
    if (gt == 0) {
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
    // Scan: every thread tests 8 slice maxima per 16-byte shared-memory load. Gather: a group
    // with passing slices is broadcast through the warp (ballot + shuffle) and its up-to 8*U
    // chunks are taken by 8*U different lanes, so all candidate re-reads are in flight together
    // instead of one thread walking them load-by-load. Lane-to-chunk assignment and the reduction
    // order are fixed: deterministic.
    static_assert(8 * U <= 32, "one warp pass per group of 8 slices");
    const float thr = M - p.skip_delta;  // NaN peak: every comparison is false, nothing accumulates
    const uint4* sm4 = reinterpret_cast<const uint4*>(smax);
    const int groups = n_tiles * (NC / 8);
    const int lane = gt & 31;
    for (int g0 = 0; g0 < groups; g0 += NT) {  // warp-uniform trip count
      const int g = g0 + gt;
      unsigned m8 = 0;
      if (g < groups) {
        const uint4 cv = sm4[g];  // 8 slice maxima
        if (CE::chunk_max(cv) >= thr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) m8 |= (CE::get(cv, j) >= thr) ? (1u << j) : 0u;
        }
      }
      unsigned ball = __ballot_sync(0xffffffffu, m8 != 0);
      while (ball) {
        const int src = __ffs(ball) - 1;
        ball &= ball - 1;
        const unsigned sm8 = __shfl_sync(0xffffffffu, m8, src);
        const int sg = __shfl_sync(0xffffffffu, g, src);
        const int j = lane / U, u = lane - j * U;
        if (lane < 8 * U && ((sm8 >> j) & 1u)) {
          const int e = sg * 8 + j;
          const int t = e / NC, owner = e - t * NC;
          const int c = (t * U + u) * NC + owner;
          if (c < n) {
            const uint4 ch = ld_stream(seg + c);
            const int flat0 = (c_begin + c) * PER;
            int y = flat0 / p.W, x = flat0 - y * p.W;
#pragma unroll
            for (int e_ = 0; e_ < PER; ++e_) {
              const float el = E::get(ch, e_);
              const float w = ex2_approx((el - M) * p.beta_log2e);
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
    if (rank == 0) window_accumulate<DT, NT>(p, mp, M, px, py, gt, s, sx, sy);
  }
  if (MODE != MVGEO_SOFT_NONE) {
    block_sum3<NW>(s, sx, sy, sc, bar, lw);
    if (S > 1 && MODE == MVGEO_SOFT_GLOBAL) {
      if (gt == 0) {
        sc.part[0] = s;
        sc.part[1] = sx;
        sc.part[2] = sy;
      }
      cluster.sync();
      if (rank == 0 && gt == 0) {
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
  if (rank == 0 && gt == 0) write_outputs(p, map, M, best, s, sx, sy, MODE != MVGEO_SOFT_NONE);
  if (S > 1) cluster.sync();  // peers' shared memory must outlive the remote reads above
}

// G independent consumer groups per CTA (8/G warps each, own ring, own mbarriers, own named
// barrier, own map sequence): small maps use G > 1 so that one group's latency-bound epilogue
// overlaps the other groups' streaming. Producer warp 8+g (one elected lane) feeds group g.
// PERSIST: grid = one resident wave, group (blockIdx.x, g) walks maps blockIdx.x*G + g, +gridDim.x*G, ...
// !PERSIST (G == 1): one segment of a cluster-split map per CTA, producer warp joins the epilogue.
template <int DT, int MODE, int U, int STAGES, int G, bool PERSIST>
__global__ void __launch_bounds__(kDecThreads + 32 * G) decode_tma_kernel(const DecodeParams p) {
  static_assert(PERSIST || G == 1, "cluster-split maps use one consumer group");
  constexpr int NW = kDecWarps / G;  // consumer warps per group
  constexpr int NT = NW * 32;
  constexpr int kTile = NT * U;  // chunks per tile of one group
  __shared__ BlockScratch sc[G];
  __shared__ __align__(8) uint64_t full_bar[G][STAGES];
  __shared__ __align__(8) uint64_t empty_bar[G][STAGES];
  extern __shared__ __align__(128) unsigned char dyn_smem[];

  const int S = PERSIST ? 1 : p.splits;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = PERSIST ? 0 : (int)(blockIdx.x % S);
  const int c_begin = rank * p.seg_chunks;
  const int n = max(min(c_begin + p.seg_chunks, p.chunks_per_map) - c_begin, 0);
  const int n_tiles = (n + kTile - 1) / kTile;
  const int64_t step = PERSIST ? (int64_t)gridDim.x * G : p.n_maps;  // !PERSIST: exactly one map

  if (tid == 0) {
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[g][s], 1);
        mbar_init(&empty_bar[g][s], NW);
      }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp >= kDecWarps) {
    // ------------------------------- producers: one warp per group (a lane suspended in try_wait
    // must not stall another group's producer), lane 0 issues ------------------------------------
    if (lane == 0) {
      const int g = warp - kDecWarps;
      uint4* ring = reinterpret_cast<uint4*>(dyn_smem) + (size_t)g * STAGES * kTile;
      const int64_t first = PERSIST ? (int64_t)blockIdx.x * G + g : (int64_t)(blockIdx.x / S);
      int s = 0, k = 0;  // slot, and how many times the ring has wrapped
      for (int64_t map = first; map < p.n_maps; map += step) {
        const uint4* seg = reinterpret_cast<const uint4*>(p.maps) + map * (int64_t)p.chunks_per_map + c_begin;
        for (int t = 0; t < n_tiles; ++t) {
          // before re-using a slot for the k-th time, wait for the consumers' (k-1)-th release of it
          if (k > 0) mbar_wait(&empty_bar[g][s], (uint32_t)((k - 1) & 1));
          const uint32_t bytes = (uint32_t)min(kTile, n - t * kTile) * 16u;
          mbar_arrive_expect_tx(&full_bar[g][s], bytes);
          bulk_copy_g2s(ring + s * kTile, seg + (size_t)t * kTile, bytes, &full_bar[g][s]);
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
  const bool consumer = warp < kDecWarps;
  const int g = consumer ? warp / NW : 0;
  const int gt = PERSIST ? tid - g * NT : tid;  // thread index inside the group
  const uint4* ring = reinterpret_cast<const uint4*>(dyn_smem) + (size_t)g * STAGES * kTile;
  uint16_t* smax = reinterpret_cast<uint16_t*>(dyn_smem + (size_t)G * STAGES * kTile * 16) + (size_t)g * n_tiles * NT;
  const int64_t first = PERSIST ? (int64_t)blockIdx.x * G + g : (int64_t)(blockIdx.x / S);
  int s = 0;
  uint32_t ph = 0;
  for (int64_t map = first; map < p.n_maps; map += step) {
    const uint4* mp = reinterpret_cast<const uint4*>(p.maps) + map * (int64_t)p.chunks_per_map;
    float run_max = __int_as_float(0xff800000);  // -inf
    int run_tile = -1;
    uint4 run_v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) run_v[u] = Elem<DT>::neg_inf_chunk();  // an all -inf map resolves to index 0
    if (consumer) {
      run_tile = gt < n ? 0 : -1;
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait(&full_bar[g][s], ph);
        const uint4* tile = ring + s * kTile;
        const int base = t * kTile;
        // vertical (packed) maximum over the thread's slice, then ONE horizontal step and ONE
        // running-maximum update per tile. The slice maximum is canonicalised (+0.0f turns -0 into
        // +0) so that the update test is a bit comparison of max.NaN results: equal values, -0/+0
        // and NaN/NaN all keep the FIRST slice.
        uint4 v[U];
        if (base + kTile <= n) {
#pragma unroll
          for (int u = 0; u < U; ++u) v[u] = tile[u * NT + gt];
        } else {
#pragma unroll
          for (int u = 0; u < U; ++u)
            v[u] = (base + u * NT + gt < n) ? tile[u * NT + gt] : Elem<DT>::neg_inf_chunk();
        }
        uint32_t vm = Elem<DT>::vmax(v[0]);
#pragma unroll
        for (int u = 1; u < U; ++u) vm = Elem<DT>::vmerge(vm, Elem<DT>::vmax(v[u]));
        const float sm = Elem<DT>::vfinish(vm) + 0.0f;
        if (MODE == MVGEO_SOFT_GLOBAL) smax[t * NT + gt] = to_carrier<DT>(sm);
        const float nm = max_nan_f32(run_max, sm);
        if (__float_as_uint(nm) != __float_as_uint(run_max)) {  // strictly better: remember the slice itself
          run_tile = t;
#pragma unroll
          for (int u = 0; u < U; ++u) run_v[u] = v[u];
        }
        run_max = nm;
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[g][s]);
        if (++s == STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
    }
    if (PERSIST)
      decode_epilogue<DT, MODE, NW, NT, U>(p, sc[g], smax, mp, map, rank, c_begin, n, n_tiles, run_max, run_tile, run_v,
                                       1 + g, gt);
    else
      decode_epilogue<DT, MODE, kDecWarps + 1, NT, U>(p, sc[0], smax, mp, map, rank, c_begin, n, n_tiles, run_max,
                                                  run_tile, run_v, 0, gt);
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
  block_argmax<kDecWarps>(my_val, my_idx, sc, 0, tid >> 5);
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
    window_accumulate<DT, kDecThreads>(p, base, M, px, py, tid, s, sx, sy);
  }
  if (MODE != MVGEO_SOFT_NONE) block_sum3<kDecWarps>(s, sx, sy, sc, 0, tid >> 5);
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
constexpr int kMaxTableBytes = 150 * 1024;  // slice-maxima table budget per CTA

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

template <int DT, int MODE, int G>
static int launch_persistent(const DecodeParams& p, size_t smem, cudaStream_t st) {
  auto kern = decode_tma_kernel<DT, MODE, kTmaU, kTmaStages, G, true>;
  // Resident-wave size for this (instantiation, shared-memory size, device): a pure function of
  // its key, cached because the occupancy query costs microseconds on a latency-bound call.
  // (Benign cache, not state: a racing thread recomputes the same value.)
  static std::atomic<uint64_t> cache{0};  // [63:48] device+1, [47:16] smem, [15:0] resident CTAs / SM count packed below
  int dev = 0;
  MVGEO_CUDA(cudaGetDevice(&dev));
  const uint64_t key = ((uint64_t)(dev + 1) << 48) | ((uint64_t)(smem & 0xffffffffu) << 16);
  uint64_t c = cache.load(std::memory_order_relaxed);
  int resident_ctas;
  if ((c & ~0xffffull) == key) {
    resident_ctas = (int)(c & 0xffff);
  } else {
    if (smem > 40 * 1024)
      MVGEO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 0, per_sm = 0;
    MVGEO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MVGEO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kDecThreads + 32 * G, smem));
    if (per_sm < 1) return MVGEO_EUNSUPPORTED;
    resident_ctas = sms * per_sm;
    if (resident_ctas > 0xffff) resident_ctas = 0xffff;
    cache.store(key | (uint64_t)resident_ctas, std::memory_order_relaxed);
  }
  const int64_t wanted = (p.n_maps + G - 1) / G;
  const unsigned grid = (unsigned)(wanted < resident_ctas ? wanted : resident_ctas);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(kDecThreads + 32 * G, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  MVGEO_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  return MVGEO_OK;
}

template <int DT, int MODE>
static int launch_decode(const DecodeParams& p, bool vec, int groups, size_t table_bytes, cudaStream_t st) {
  if (!vec) {
    decode_scalar_kernel<DT, MODE><<<(unsigned)p.n_maps, kDecThreads, 0, st>>>(p);
    MVGEO_CHECK_LAUNCH();
    return MVGEO_OK;
  }
  const size_t smem = (size_t)kTmaStages * kTmaU * kDecThreads * 16 + table_bytes;
  if (p.splits > 1)
    return launch_clustered(decode_tma_kernel<DT, MODE, kTmaU, kTmaStages, 1, false>, p,
                            (unsigned)(p.n_maps * p.splits), kDecThreads + 32, smem, st);
  switch (groups) {
    case 4: return launch_persistent<DT, MODE, 4>(p, smem, st);
    case 2: return launch_persistent<DT, MODE, 2>(p, smem, st);
    default: return launch_persistent<DT, MODE, 1>(p, smem, st);
  }
}

template <int DT>
static int dispatch_mode(const DecodeParams& p, int mode, bool vec, int groups, size_t table_bytes, cudaStream_t st) {
  switch (mode) {
    case MVGEO_SOFT_NONE: return launch_decode<DT, MVGEO_SOFT_NONE>(p, vec, groups, 0, st);
    case MVGEO_SOFT_GLOBAL: return launch_decode<DT, MVGEO_SOFT_GLOBAL>(p, vec, groups, table_bytes, st);
    case MVGEO_SOFT_WINDOW: return launch_decode<DT, MVGEO_SOFT_WINDOW>(p, vec, groups, 0, st);
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
  int groups = 1;
  if (vec) {
    // Work decomposition is a function of the map size only (never of n_maps), so results are
    // bit-identical however the frames are sharded across GPUs:
    //   small maps  -> several consumer groups per CTA, one map stream each (epilogues overlap):
    //                  4 up to 112 KB (native 128x128 fp32, C1), 2 up to 160 KB (C2 / C3), measured;
    //   large maps  -> one group; split over a cluster only when the slice-maxima table
    //                  (2 bytes per kTmaU chunks) cannot fit one CTA.
    const int64_t chunks = map_bytes / 16;
    groups = map_bytes <= 112 * 1024 ? 4 : (map_bytes <= 160 * 1024 ? 2 : 1);
    const int64_t tile = (int64_t)(kDecThreads / groups) * kTmaU;
    int splits = 1;
    while (splits < kMaxSplits && ((chunks + splits - 1) / splits + tile - 1) / tile * kDecThreads * 2 > kMaxTableBytes)
      ++splits;
    if (splits > 1) groups = 1;
    p.chunks_per_map = (int)chunks;
    p.seg_chunks = (int)((chunks + splits - 1) / splits);
    p.splits = splits;
    if (soft_mode == MVGEO_SOFT_GLOBAL) {
      const int64_t tile1 = (int64_t)(kDecThreads / groups) * kTmaU;
      smem = (size_t)((p.seg_chunks + tile1 - 1) / tile1) * kDecThreads * 2;  // all groups' tables
      if (smem > (size_t)kMaxTableBytes) return MVGEO_EUNSUPPORTED;  // maps beyond ~19 MB: use the window mode
    }
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (dtype) {
    case MVGEO_F32: return dispatch_mode<MVGEO_F32>(p, soft_mode, vec, groups, smem, st);
    case MVGEO_BF16: return dispatch_mode<MVGEO_BF16>(p, soft_mode, vec, groups, smem, st);
    case MVGEO_F16: return dispatch_mode<MVGEO_F16>(p, soft_mode, vec, groups, smem, st);
  }
  return MVGEO_EINVAL;
}
