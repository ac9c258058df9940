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

struct BlockScratch {
  float val[kDecWarps];
  int idx[kDecWarps];
  float sum[3][kDecWarps];
  // per-CTA results, read by cluster peers through DSMEM
  float best_val;
  int best_idx;
  float part[3];
};

// Block-wide (value, index) arg-max. Result valid in every thread.
__device__ __forceinline__ void block_argmax(float& v, int& i, BlockScratch& s) {
  warp_argmax(v, i);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s.val[warp] = v;
    s.idx[warp] = i;
  }
  __syncthreads();
  v = s.val[0];
  i = s.idx[0];
#pragma unroll
  for (int w = 1; w < kDecWarps; ++w) {
    if (argmax_better(v, i, s.val[w], s.idx[w])) {
      v = s.val[w];
      i = s.idx[w];
    }
  }
  __syncthreads();
}

// Block-wide fixed-order sums of three accumulators. Result valid in every thread.
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
  __syncthreads();
  a = b = c = 0.f;
#pragma unroll
  for (int w = 0; w < kDecWarps; ++w) {
    a += s.sum[0][w];
    b += s.sum[1][w];
    c += s.sum[2][w];
  }
  __syncthreads();
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
template <int DT>
__device__ __forceinline__ void window_accumulate(const DecodeParams& p, const void* map_base, float M, int px,
                                                  int py, float& s, float& sx, float& sy) {
  using E = Elem<DT>;
  const int r = p.radius, side = 2 * r + 1;
  for (int t = threadIdx.x; t < side * side; t += kDecThreads) {
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
template <int DT, int MODE>
__global__ void __launch_bounds__(kDecThreads, 4) decode_vec_kernel(const DecodeParams p) {
  using E = Elem<DT>;
  using carrier = typename E::carrier;
  constexpr int PER = E::kPerChunk;
  __shared__ BlockScratch sc;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  carrier* cmax = reinterpret_cast<carrier*>(dyn_smem);

  const int S = p.splits;
  const int64_t map = blockIdx.x / S;
  const int rank = (int)(blockIdx.x - map * S);
  const int tid = threadIdx.x;

  const uint4* mp = reinterpret_cast<const uint4*>(p.maps) + map * (int64_t)p.chunks_per_map;
  const int c_begin = rank * p.seg_chunks;
  const int c_end = min(c_begin + p.seg_chunks, p.chunks_per_map);
  const int n = max(c_end - c_begin, 0);
  const uint4* seg = mp + c_begin;

  // ---------------- pass 1: stream the segment once --------------------------------------
  float run_max = __int_as_float(0xff800000);  // -inf
  int run_chunk = tid < n ? tid : -1;
  constexpr int kBatch = kDecThreads * kDecUnroll;
  const int iters = (n + kBatch - 1) / kBatch;

  uint4 cur[kDecUnroll], nxt[kDecUnroll];
#pragma unroll
  for (int u = 0; u < kDecUnroll; ++u) {
    const int c = u * kDecThreads + tid;
    cur[u] = c < n ? ld_stream(seg + c) : E::neg_inf_chunk();
  }
  for (int it = 0; it < iters; ++it) {
    const int base = it * kBatch;
    if (it + 1 < iters) {
#pragma unroll
      for (int u = 0; u < kDecUnroll; ++u) {
        const int c = base + kBatch + u * kDecThreads + tid;
        nxt[u] = c < n ? ld_stream(seg + c) : E::neg_inf_chunk();
      }
    }
#pragma unroll
    for (int u = 0; u < kDecUnroll; ++u) {
      const int c = base + u * kDecThreads + tid;
      const float cm = E::chunk_max(cur[u]);
      if (MODE == MVGEO_SOFT_GLOBAL) {
        if (c < n) cmax[c] = E::pack(cm);
      }
      const bool gt = (cm > run_max) || ((cm != cm) && (run_max == run_max));
      if (gt) {
        run_max = cm;
        run_chunk = c;
      }
    }
#pragma unroll
    for (int u = 0; u < kDecUnroll; ++u) cur[u] = nxt[u];
  }

  // ---------------- resolve the first maximal element inside the winning chunk -----------
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
  block_argmax(my_val, my_idx, sc);

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

  // ---------------- pass 2: soft-arg-max sums ---------------------------------------------
  float s = 0.f, sx = 0.f, sy = 0.f;
  if (MODE == MVGEO_SOFT_GLOBAL) {
    const float thr = M - p.skip_delta;  // NaN peak: every comparison is false, nothing accumulates
    for (int c = tid; c < n; c += kDecThreads) {
      if (E::unpack(cmax[c]) >= thr) {
        const uint4 ch = ld_stream(seg + c);
        const int flat0 = (c_begin + c) * PER;
        int y = flat0 / p.W, x = flat0 - y * p.W;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
          const float e = E::get(ch, j);
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
  } else if (MODE == MVGEO_SOFT_WINDOW) {
    if (rank == 0) window_accumulate<DT>(p, mp, M, px, py, s, sx, sy);
  }
  if (MODE != MVGEO_SOFT_NONE) {
    block_sum3(s, sx, sy, sc);
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
  block_argmax(my_val, my_idx, sc);
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
    window_accumulate<DT>(p, base, M, px, py, s, sx, sy);
  }
  if (MODE != MVGEO_SOFT_NONE) block_sum3(s, sx, sy, sc);
  if (tid == 0) write_outputs(p, map, M, best, s, sx, sy, MODE != MVGEO_SOFT_NONE);
}

template <int DT, int MODE>
static int launch_decode(const DecodeParams& p, bool vec, size_t smem, cudaStream_t st) {
  if (!vec) {
    decode_scalar_kernel<DT, MODE><<<(unsigned)p.n_maps, kDecThreads, 0, st>>>(p);
    MVGEO_CHECK_LAUNCH();
    return MVGEO_OK;
  }
  auto kern = decode_vec_kernel<DT, MODE>;
  if (smem > 48 * 1024) MVGEO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(p.n_maps * p.splits), 1, 1);
  cfg.blockDim = dim3(kDecThreads, 1, 1);
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

template <int DT>
static int dispatch_mode(const DecodeParams& p, int mode, bool vec, size_t smem, cudaStream_t st) {
  switch (mode) {
    case MVGEO_SOFT_NONE: return launch_decode<DT, MVGEO_SOFT_NONE>(p, vec, 0, st);
    case MVGEO_SOFT_GLOBAL: return launch_decode<DT, MVGEO_SOFT_GLOBAL>(p, vec, smem, st);
    case MVGEO_SOFT_WINDOW: return launch_decode<DT, MVGEO_SOFT_WINDOW>(p, vec, 0, st);
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
      smem = (size_t)p.seg_chunks * (dtype == MVGEO_F32 ? 4 : 2);
      if (smem > 200 * 1024) return MVGEO_EUNSUPPORTED;  // maps beyond ~6.4 MB (f32): use the window mode
    }
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (dtype) {
    case MVGEO_F32: return dispatch_mode<MVGEO_F32>(p, soft_mode, vec, smem, st);
    case MVGEO_BF16: return dispatch_mode<MVGEO_BF16>(p, soft_mode, vec, smem, st);
    case MVGEO_F16: return dispatch_mode<MVGEO_F16>(p, soft_mode, vec, smem, st);
  }
  return MVGEO_EINVAL;
}
