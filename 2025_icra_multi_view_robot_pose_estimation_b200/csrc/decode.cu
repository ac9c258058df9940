// decode.cu — belief-map key-point decoder (kernel 1 of the hot path), sm_100a.
//
// Replaces extract_keypoints_from_heatmaps (model/Fr5_model_train.ipynb:4674-4705 and twins)
// and the inline arg-max loops (DIP_REAL.py:116-124, model/MvRoPose_FR3.py:299-304),
// batched over n_maps = B*V*K maps, plus the sub-pixel soft-arg-max the reference lacks.
//
// Roofline: HBM. Every map byte is read from DRAM exactly once in EVERY mode and for EVERY data
// distribution (algorithmic bytes per map = H*W*sizeof(dtype); outputs are 28 B per map). Design:
//   * PERSISTENT kernel: (SMs x resident CTAs) CTAs, each walking maps blockIdx.x, +gridDim.x, ...
//     One elected producer lane per map stream issues 2-D tensor-map TMA loads (cp.async.bulk.tensor.2d,
//     SASS UTMALDG.2D; the maps are described as a tensor of 128-byte rows) of 32 KB / G tiles into a
//     2-stage shared-memory ring with full/empty mbarriers, written with the 128-byte hardware swizzle so
//     that every consumer thread owns a RUN of 128 contiguous map bytes — one TMA row — and still reads it
//     with conflict-free ld.shared.v4 (64-byte runs, 16 KB / G tiles and 4 stages where the image rows do
//     not hold whole 128-byte runs). Bytes in flight are set by the ring, not by registers, and the
//     producer keeps prefetching the NEXT map while the consumers are in the per-map epilogue
//     (consumer-only named barrier), so the DRAM pipe never drains;
//   * arg-max: a packed max.NaN tree per 16-byte chunk (bf16x2 / f16x2 SIMD for 16-bit maps) and ONE
//     running-maximum update per thread per tile; the raw chunks of the best run are parked in shared
//     memory, so the first maximal element is resolved without touching global memory again: warp
//     maximum = one redux.sync, first run holding it = one redux.sync, first maximal element of that run =
//     one ballot (lane j looks at element j), with torch.argmax's first-maximum, NaN-is-maximal ordering;
//   * global soft-arg-max: ONLINE softmax in the same streaming loop. Every thread keeps
//     (sum w, sum w x, sum w y) with w = 2^(h*beta' - ref*beta'): one FFMA2 per two elements, one MUFU.EX2
//     per element, moments from suffix sums of the run (FADD2 only), position applied once per tile. The
//     reference moves only when a run climbs 64 log2 units above it, and the f32 sums are folded in double
//     into per-thread totals at that moment and every 16 tiles ("epochs"): no per-tile rescale, and the f32
//     rounding of a long background tail after a peak stays bounded. Cost and DRAM traffic are independent of
//     the data (round 1 re-read every slice within 32/beta of the maximum: free for a sharp peak at large
//     beta, but 1.06 TB/s for flat / low-amplitude maps or small beta);
//   * maps up to 2 MB run 4 independent map streams per CTA (consumer groups of 2 warps, each
//     with its own ring, mbarriers, named barrier and producer warp) so that epilogues overlap; maps up to
//     32 KB (the reference-native 128x128 under bf16 autocast) run 8 ONE-WARP streams that feed themselves:
//     lane 0 re-issues the TMA copy into the slot its own warp has just emptied — no producer warps, no
//     empty barriers, every per-map reduction warp-wide (16 tiles per thread and map instead of 8: +13 %);
//   * V per-view base pointers (the reference's dict view -> (B,K,H,W), model/MvRoPose_FR3.py:625)
//     are walked by ONE launch through V tensor maps: map m = (b, v, k) is read from view v's tensor,
//     results land in [B,V,K] order. No stack copy, no per-view launch;
//   * MSE variant: the heat-map loss against on-the-fly Gaussian targets from the same registers
//     (one-read training step).
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include <atomic>
#include <cstring>
#include <type_traits>

#include "common.cuh"

namespace mvgeo {

constexpr int kDecThreads = 256;
constexpr int kDecWarps = kDecThreads / 32;

struct DecodeParams {
  const void* view_base[MVGEO_MAX_VIEWS];  // n_views == 1: one dense [n_maps, H, W] array
  int n_views;
  int k_per_view;  // K (maps per frame and view) when n_views > 1
  int64_t n_maps;
  int64_t map_bytes;
  int H, W;
  int rows_per_map;    // streaming kernel: runs (128 or 64 bytes, one per thread and tile) per map
  double scale_x, scale_y;
  float beta_log2e;  // beta * log2(e)
  float skip_delta;  // scalar kernel only: kSoftSkip / beta
  int radius;
  int apply_sigmoid;
  int64_t k_inner, out_stride, out_offset;
  int32_t* idx;
  float* peak;
  float* score;
  float* kp_hard;
  float* kp_soft;
  // fused heat-map MSE against on-the-fly Gaussian targets (decode_tma_kernel<..., MSE = true>)
  const float* mse_kp;   // [n_maps, 2] target centres in map pixels (non-finite: all-zero target)
  float mse_k;           // log2(e) / (2 sigma^2)
  float* mse_partial;    // [n_maps] sum over the map of (pred - target)^2
};

// Map m in result order [B, V, K] -> its first byte. With one base pointer the maps are dense.
__device__ __forceinline__ const char* map_ptr(const DecodeParams& p, int64_t m) {
  if (p.n_views == 1) return reinterpret_cast<const char*>(p.view_base[0]) + m * p.map_bytes;
  const int64_t f = m / p.k_per_view;  // (b, v)
  const int k = (int)(m - f * p.k_per_view);
  const int64_t b = f / p.n_views;
  const int v = (int)(f - b * p.n_views);
  return reinterpret_cast<const char*>(p.view_base[v]) + (b * p.k_per_view + k) * p.map_bytes;
}

constexpr int kMaxWarps = 8;

struct BlockScratch {
  float val[kMaxWarps];
  int idx[kMaxWarps];
  float sum[3][kMaxWarps];
  uint32_t key[kMaxWarps];
  double dsum[3][kMaxWarps];
};

// Barrier over a group of NW warps. bar == 0 is __syncthreads() (the group must then be the whole
// CTA); bar > 0 is a named barrier, used by the persistent kernel whose consumer warp groups
// reduce independently of one another and of the producer warps.
template <int NW>
__device__ __forceinline__ void group_sync(int bar) {
  if (NW == 1) {  // a one-warp group
    __syncwarp();
    return;
  }
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

// (s, sx, sy) are the soft-arg-max sums RELATIVE TO THE HARD PEAK: sum w, sum w (x - px), sum w (y - py).
__device__ __forceinline__ void write_outputs(const DecodeParams& p, int64_t map, float M, int best, float s,
                                              float sx, float sy, bool have_soft) {
  int64_t o = map;  // dense [n_maps] results: the common case, no 64-bit division
  if (p.k_inner != 1 || p.out_stride != 1 || p.out_offset != 0)
    o = (map / p.k_inner) * p.out_stride + p.out_offset + (map % p.k_inner);
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
        // offset from the hard peak in float (relative error 6e-8 of an offset of a few pixels), position
        // and scaling in double like the hard key-point
        const float inv = 1.0f / s;
        qx = (float)(((double)px + (double)(sx * inv)) * p.scale_x);
        qy = (float)(((double)py + (double)(sy * inv)) * p.scale_y);
      }
    }
    p.kp_soft[2 * o] = qx;
    p.kp_soft[2 * o + 1] = qy;
  }
}

// Window soft-arg-max around (px,py): every thread takes window cells tid, tid+NT, ...
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
// Streaming path: 16-byte aligned maps made of whole 128-byte TMA rows (global soft mode: image rows that hold
// whole 128- or 64-byte runs, so a thread's run never straddles two image rows).
// ----------------------------------------------------------------------------------------

// Online soft-arg-max state of one thread: sums of w = 2^(h*beta' + nb), nb = -ref*beta', over the
// elements of the current EPOCH, with coordinates relative to the map centre (exact small floats).
// Every sum is an f32x2 pair (even / odd element of each 2-element group) so that the loop runs on
// packed FFMA2 / FADD2: one issue slot per two elements.
//
// Epochs. The reference `ref` is NOT the running maximum (with a per-thread running maximum SOME lane
// of a warp breaks its record in ~90% of the tiles and the whole warp pays a rescale: 15% of all issued
// instructions). Floating point keeps its relative precision for weights far from 1, so the reference only
// moves when a run maximum climbs more than kEpochWindow (log2 units) above it: the first finite run, and
// the few runs in which a thread climbs the peak. At that moment — and every kFoldPeriod 64-byte tiles (half as
// many 128-byte tiles) regardless —
// the epoch's f32 sums are folded, in DOUBLE, into the thread's per-map totals in shared memory and start
// again from zero. The periodic fold bounds what f32 accumulation can lose: once a thread has met the peak
// its first moments carry the lever arm to the map centre (hundreds of pixels) times the peak's mass, and
// every later background tile added on top would be rounded at THAT scale — measured up to 3e-4 px when the
// background holds ~1e-3 of the weight; with at most kFoldPeriod tiles per fold the bound is ~5e-5 px.
// Cost: nothing per element; ~70 instructions per fold, i.e. ~2% of the loop.
constexpr float kEpochWindow = 64.0f;  // range only: weights stay below 2^64, sums below 2^100
constexpr int kFoldPeriod = 16;  // tiles of 64-byte runs (power of two)
// Experiment kept for reproduction, OFF: every n-th element pair of a 16-bit run takes its exponential on the FMA
// pipe (exp2_poly2: 12 instructions per pair instead of 2 MUFU). With the XU pipe 72 % busy and the FMA pipe 25 %
// this looked like headroom; measured (interleaved A/B, all 20 regimes): n = 4: -6 %, n = 3: -5 %, n = 2: -21 %
// with 64-byte runs, n = 4: -4.4 % with 128-byte runs (XU 77 %). Every added instruction costs its issue slot.
#ifndef MVGEO_POLY_EVERY
#define MVGEO_POLY_EVERY 0
#endif
constexpr int kPolyEvery = MVGEO_POLY_EVERY;
struct SoftAcc {
  float nb;      // -ref * beta_log2e as rounded (the fold corrects with the SAME value)
  float ref_hi;  // ref + kEpochWindow / beta_log2e
  f32x2 s, sx, sy;       // sum w, sum w (x0 - ox), sum w (y - oy), x0 = the run's first column
  f32x2 sj;              // sum w i, i = index of the element's 2-element group inside its run
};

// Per-thread, per-map totals over the finished epochs: S, SX, SY relative to reference nb (32 bytes per thread in
// shared memory, touched only by the fold). Stored as arrays over the NT threads of a group — S[NT], SX[NT],
// SY[NT] (double), nb[NT] (float) — so that a warp's accesses are consecutive words: no bank conflicts.
struct SoftTotals {
  double S, SX, SY;
  float nb;  // +inf: empty
};
// the group's block is aligned to its size (NT * 32): thread index = (addr mod NT*8) / 8, nb[] starts at +NT*24
template <int NT>
__device__ __forceinline__ uint32_t totals_nb_addr(uint32_t addr) { return addr - ((addr & (NT * 8 - 1)) >> 1); }
template <int NT>
__device__ __forceinline__ void totals_store(uint32_t addr, const SoftTotals& t) {  // addr = group base + 8 * thread
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(t.S) : "memory");
  asm volatile("st.shared.f64 [%0+%2], %1;" ::"r"(addr), "d"(t.SX), "n"(NT * 8) : "memory");
  asm volatile("st.shared.f64 [%0+%2], %1;" ::"r"(addr), "d"(t.SY), "n"(NT * 16) : "memory");
  asm volatile("st.shared.f32 [%0+%2], %1;" ::"r"(totals_nb_addr<NT>(addr)), "f"(t.nb), "n"(NT * 24) : "memory");
}
template <int NT>
__device__ __forceinline__ SoftTotals totals_load(uint32_t addr) {
  SoftTotals t;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(t.S) : "r"(addr));
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(t.SX) : "r"(addr), "n"(NT * 8));
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(t.SY) : "r"(addr), "n"(NT * 16));
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(t.nb) : "r"(totals_nb_addr<NT>(addr)), "n"(NT * 24));
  return t;
}
// Fold the epoch held in `a` into the totals (common reference = the higher of the two, so the
// scale factor is <= 1 and can only underflow to a truly negligible 0), and clear the epoch.
template <int NT>
__device__ __forceinline__ void epoch_fold(SoftAcc& a, uint32_t totals_addr) {
  const float ts = sum2(a.s);
  if (ts > 0.f) {  // an empty epoch (or one poisoned by +inf - +inf) adds nothing
    float s_even, s_odd, h0, h1;
    unpack2(a.s, s_even, s_odd);
    const double S = (double)s_even + (double)s_odd;
    unpack2(a.sx, h0, h1);
    float j0, j1;
    unpack2(a.sj, j0, j1);
    // sum_j j*w_j over the runs = 2 * sum_i i*(w_i.lo + w_i.hi) + sum_i w_i.hi
    const double SX = ((double)h0 + (double)h1) + 2.0 * ((double)j0 + (double)j1) + (double)s_odd;
    unpack2(a.sy, h0, h1);
    const double SY = (double)h0 + (double)h1;
    SoftTotals t = totals_load<NT>(totals_addr);
    if (a.nb <= t.nb) {  // the epoch's reference is the higher one (nb = -ref*beta'): bring the totals to it
      const double f = (double)ex2_approx(a.nb - t.nb);  // empty totals: nb = +inf -> f = 0
      t.S = t.S * f + S;
      t.SX = t.SX * f + SX;
      t.SY = t.SY * f + SY;
      t.nb = a.nb;
    } else {
      const double f = (double)ex2_approx(t.nb - a.nb);
      t.S += S * f;
      t.SX += SX * f;
      t.SY += SY * f;
    }
    totals_store<NT>(totals_addr, t);
  }
  a.s = a.sx = a.sy = a.sj = 0ull;
}

// Order-preserving map float -> uint32 with NaN (any sign, any payload) on top, so that a warp-wide
// maximum with torch's NaN-is-maximal rule is ONE redux.sync instruction.
__device__ __forceinline__ uint32_t ord_key(float v) {
  const uint32_t u = __float_as_uint(v);
  if (v != v) return 0xffffffffu;
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_val(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// One map element from shared memory (32-bit shared address) as float.
template <int DT>
__device__ __forceinline__ float lds_elem(uint32_t addr) {
  if constexpr (DT == MVGEO_F32) {
    return lds32f(addr);
  } else {
    uint16_t b;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(b) : "r"(addr));
    return Elem<DT>::unpack(b);
  }
}

// Up to MVGEO_MAX_VIEWS TMA tensor maps, one per base pointer (kernel parameter, __grid_constant__): every
// view's maps seen as a 2-D tensor [rows of 128 bytes][128 bytes].
struct TensorMaps {
  CUtensorMap m[MVGEO_MAX_VIEWS];
};

// 2-D tiled TMA load (SASS UTMALDG): box = [rows x 128 B] at row `row` into `dst`, bytes counted on `bar`.
__device__ __forceinline__ void tma_load_rows(uint32_t dst_smem, const CUtensorMap* tm, int row, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   dst_smem),
               "l"(tm), "r"(0), "r"(row), "r"(bar)
               : "memory");
}

// The same, issued by a CONSUMER lane for the slot its own warp has just emptied: `dep` is a value computed from
// the slot's last ld.shared, so the copy cannot be issued before the warp's reads of the slot have returned.
__device__ __forceinline__ void tma_load_rows_after(uint32_t dst_smem, const CUtensorMap* tm, int row, uint32_t bar,
                                                    uint32_t dep) {
  asm volatile(
      "{\n\t.reg .b32 d;\n\tmov.b32 d, %5;\n\t"
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n\t}" ::"r"(
          dst_smem),
      "l"(tm), "r"(0), "r"(row), "r"(bar), "r"(dep)
      : "memory");
}

// G independent consumer groups per CTA (8/G warps each, own ring, own mbarriers, own named
// barrier, own map sequence): small maps use G > 1 so that one group's latency-bound epilogue
// overlaps the other groups' streaming. Producer warp 8+g (one elected lane) feeds group g.
// Grid = one resident wave; group (blockIdx.x, g) walks maps blockIdx.x*G + g, +gridDim.x*G, ...
//
// Tile = NT*U/8 rows of 128 bytes, loaded by ONE 2-D TMA tensor copy with the 128-byte swizzle. U = 8: thread gt owns
// row gt, a RUN of 128 contiguous map bytes (64 bf16 / 32 f32 elements inside one image row), chunk u at u ^ (gt & 7).
// U = 4: thread gt owns half (gt & 1) of row gt >> 1, a run of 64 bytes, chunk u at (4 (gt & 1) + u) ^ ((gt >> 1) & 7).
// Either way conflict-free ld.shared.v4 although the lane stride is 128 / 64 bytes. (64-byte TMA rows with the
// 64-byte swizzle work too but stream 4 % slower.) Owning a contiguous run makes the soft-arg-max bookkeeping per
// TILE instead of per 16-byte chunk: one (column, row) position, one pair of first-moment FFMA2, one suffix-sum pass
// over the run — and the longer the run, the less the per-tile fixed costs weigh (launch_persistent).
#ifndef MVGEO_DEC_MINB
#define MVGEO_DEC_MINB 2  // measured: 2 CTAs/SM without a register cap beat 3 CTAs/SM at 64 registers (spills)
#endif
constexpr int kRowBytes = 128;  // TMA row; a thread's run is a whole row (U = 8 chunks) or half of one (U = 4)
// MSE: the same pass also accumulates sum (pred - g)^2 per map against the separable Gaussian target
// g(x, y) = ex[x] * ey[y] built in shared memory per map (W + H exponentials, as csrc/encode.cu does) — the
// training step's heat-map loss (nn.MSELoss, model/MvRoPose_FR3.py:846-847) and the decode of the same
// prediction read the maps ONCE.
// G == 8: one-warp groups that feed THEMSELVES — lane 0 re-issues the TMA copy into the slot its warp has just
// emptied (no producer warps, no empty barriers, no named barriers: every per-map reduction is warp-wide).
// (Self-feeding groups of 2 / 4 warps — an atomic per slot, the last warp to empty it re-issues the copy — were
// measured too: within +-3 % of the producer warps, sign depending on the box; not kept.)
constexpr int producer_threads(int G) { return G >= 8 ? 0 : 32 * G; }
template <int DT, int MODE, int U, int STAGES, int G, bool MSE = false>
__global__ void __launch_bounds__(kDecThreads + producer_threads(G), MVGEO_DEC_MINB)
    decode_tma_kernel(const DecodeParams p, const __grid_constant__ TensorMaps tms) {
  static_assert(!MSE || MODE == MVGEO_SOFT_NONE, "the fused loss pass decodes the hard peak only");
  static_assert(U == 4 || U == 8, "a run is half a 128-byte TMA row or a whole one");
  constexpr int RPR = kRowBytes / (U * 16);  // runs per TMA row
  using E = Elem<DT>;
  constexpr int PER = E::kPerChunk;
  constexpr int EPR = U * PER;       // elements per run
  constexpr int NW = kDecWarps / G;  // consumer warps per group
  constexpr bool SELF = producer_threads(G) == 0;
  constexpr int NT = NW * 32;
  constexpr int kTile = NT * U;  // chunks per tile of one group
  constexpr uint32_t kTileBytes = kTile * 16;
  __shared__ BlockScratch sc[G];
  __shared__ __align__(8) uint64_t full_bar[G][STAGES];
  __shared__ __align__(8) uint64_t empty_bar[G][STAGES];
  extern __shared__ __align__(1024) unsigned char dyn_smem[];  // the 128-byte swizzle pattern repeats every 1024 bytes

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rows = p.rows_per_map;                   // runs per map
  const int n_full = rows / NT;                      // tiles without padding
  const int n_tiles = (rows + NT - 1) / NT;          // n_full or n_full + 1
  // (A balanced assignment — only ceil(n_maps / rounds) streams active so that no stream walks a ragged last round —
  // was measured and rejected: -1.6 % at C2, -11 % at C5. The streams left over in the last round run faster, and
  // parallelism per SM is worth more than an even finish.)
  const int64_t step = (int64_t)gridDim.x * G;

  if (tid == 0) {
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(smem_u32(&full_bar[g][s]), 1);
        if (!SELF) mbar_init(smem_u32(&empty_bar[g][s]), NW);
      }
    mbar_fence_init();
  }
  __syncthreads();

  const int g = warp >= kDecWarps ? warp - kDecWarps : warp / NW;  // the group this warp feeds / belongs to
  const int64_t map_first = (int64_t)blockIdx.x * G + g;
  const uint32_t ring_s = smem_u32(dyn_smem) + (uint32_t)g * STAGES * kTileBytes;
  const uint32_t full_s = smem_u32(&full_bar[g][0]);
  const uint32_t empty_s = smem_u32(&empty_bar[g][0]);

  // map (b, v, k) -> view v's tensor map and the map's first 128-byte row inside that view (maps are whole rows)
  auto locate = [&](int64_t map, const CUtensorMap*& tm, int& row0) {
    int v = 0;
    int64_t local = map;
    if (p.n_views > 1) {
      const int64_t f = map / p.k_per_view;
      const int64_t b = f / p.n_views;
      v = (int)(f - b * p.n_views);
      local = b * p.k_per_view + (map - f * p.k_per_view);
    }
    tm = &tms.m[v];
    row0 = (int)(local * (rows / RPR));
  };

  if (!SELF && warp >= kDecWarps) {
    // ------------------------------- producers: one warp per group (a lane suspended in try_wait
    // must not stall another group's producer), lane 0 issues ------------------------------------
    if (lane == 0) {
      int s = 0, k = 0;  // slot, and how many times the ring has wrapped
      for (int64_t map = map_first; map < p.n_maps; map += step) {
        const CUtensorMap* tm;
        int row0;
        locate(map, tm, row0);
        for (int t = 0; t < n_tiles; ++t) {
          // before re-using a slot for the k-th time, wait for the consumers' (k-1)-th release of it
          if (k > 0) mbar_wait(empty_s + 8 * s, (uint32_t)((k - 1) & 1));
          // the box is always delivered whole: rows past the end of the tensor arrive as zeros, rows of the
          // NEXT map as that map's data — the consumers mask both (they know how many runs the map has)
          mbar_arrive_expect_tx(full_s + 8 * s, kTileBytes);
          tma_load_rows(ring_s + s * kTileBytes, tm, row0 + t * (NT / RPR), full_s + 8 * s);
          if (++s == STAGES) {
            s = 0;
            ++k;
          }
        }
      }
    }
    return;  // the producer warps take no part in the per-map reductions
  }

  // --------------------------------- consumers ---------------------------------------------
  const int gt = tid - g * NT;  // thread index inside the group = the row of the tile this thread owns
  const int lw = gt >> 5;
  const int bar = 1 + g;
  BlockScratch& scr = sc[g];
  const uint32_t my_s = ring_s + (uint32_t)(gt / RPR) * kRowBytes;  // this thread's 128-byte row in slot 0
  // 128-byte swizzle: 16-byte chunk c of row r lives at c ^ (r & 7); this thread's chunk u is c = U (gt % RPR) + u
  const uint32_t sw = (uint32_t)((((gt % RPR) * U) ^ ((gt / RPR) & 7)) << 4);
  // behind the rings: the raw chunks of every thread's best run so far ([g][u][gt], 16 KB per CTA: the
  // epilogue finds the first maximal element there, not in global memory) ...
  const uint32_t cand_s = smem_u32(dyn_smem) + G * STAGES * kTileBytes + (uint32_t)(g * kTile + gt) * 16;
  // ... and every thread's per-map soft-arg-max totals (32 bytes each)
  const uint32_t tot_s = smem_u32(dyn_smem) + G * STAGES * kTileBytes + G * kTileBytes + (uint32_t)(g * NT * 32 + gt * 8);

  // Soft-arg-max geometry: coordinates are relative to the map centre; from one tile to the next a
  // run advances by (step_y rows, step_x columns) with at most one row wrap.
  const float ox = 0.5f * (float)p.W, oy = 0.5f * (float)p.H, Wf = (float)p.W;
  const float x_hi = Wf - ox;  // first column value that belongs to the next row
  const int step_yi = (NT * EPR) / p.W, step_xi = (NT * EPR) % p.W;
  const float step_y = (float)step_yi, step_x = (float)step_xi;
  const int e00 = gt * EPR, y00 = e00 / p.W, x00 = e00 - y00 * p.W;  // this thread's run of tile 0
  const f32x2 beta2 = pack2(p.beta_log2e, p.beta_log2e);
  const float window = kEpochWindow / p.beta_log2e;
  const float kNegInf = __int_as_float(0xff800000);

  // MSE: behind the totals, one table of ex[0..Wp) and ey[0..H) per consumer group
  const int Wp = (p.W + 3) & ~3;
  const uint32_t tab_s = smem_u32(dyn_smem) + G * STAGES * kTileBytes + G * kTileBytes + kDecThreads * 32 +
                         (uint32_t)g * (uint32_t)(Wp + ((p.H + 3) & ~3)) * 4;

  // SELF: lane 0's copy cursor runs STAGES tiles ahead of the warp's own consumption
  int64_t c_map = map_first;
  int c_tile = 0, c_row = 0;
  const CUtensorMap* c_tm = nullptr;
  static_assert(!SELF || NW == 1, "only one-warp groups feed themselves");
  auto issue = [&](int slot, uint32_t dep) {  // lane 0 only
    if (c_map < p.n_maps) {
      mbar_arrive_expect_tx(full_s + 8 * slot, kTileBytes);
      tma_load_rows_after(ring_s + slot * kTileBytes, c_tm, c_row, full_s + 8 * slot, dep);
      c_row += NT / RPR;
      if (++c_tile == n_tiles) {
        c_tile = 0;
        c_map += step;
        if (c_map < p.n_maps) locate(c_map, c_tm, c_row);
      }
    }
  };
  if (SELF && lane == 0) {
    if (c_map < p.n_maps) locate(c_map, c_tm, c_row);
    for (int i = 0; i < STAGES; ++i) issue(i, 0u);
  }

  int s = 0;
  uint32_t ph = 0;
  for (int64_t map = map_first; map < p.n_maps; map += step) {
    float run_max = kNegInf;
    int run_tile = gt < rows ? 0 : -1;
    SoftAcc a = {0.f, kNegInf, 0ull, 0ull, 0ull, 0ull};
    if (MODE == MVGEO_SOFT_GLOBAL) totals_store<NT>(tot_s, SoftTotals{0.0, 0.0, 0.0, __int_as_float(0x7f800000)});
    int ix = x00, iy = y00;
    f32x2 mse_acc = 0ull;
    if (MSE) {
      // the map's separable Gaussian tables (consumers of this group only; the previous map's last table
      // reads are behind its epilogue barriers)
      const float cx = p.mse_kp[2 * map], cy = p.mse_kp[2 * map + 1];
      const bool okc = isfinite(cx) && isfinite(cy);
      for (int i = gt; i < p.W + p.H; i += NT) {
        const float d = i < p.W ? (float)i - cx : (float)(i - p.W) - cy;
        sts32f(tab_s + (uint32_t)(i < p.W ? i : Wp + i - p.W) * 4, okc ? ex2_approx(-d * d * p.mse_k) : 0.f);
      }
      group_sync<NW>(bar);
    }
    float fx = (float)x00 - ox, fy = (float)y00 - oy;

    // One tile: wait for the slot, take the thread's run into registers, arg-max bookkeeping,
    // hand the slot back, then (global soft mode) the exponentials. FULL = no padding in the tile.
    auto tile_body = [&](auto full_tag, int t) {
      constexpr bool FULL = decltype(full_tag)::value;
      mbar_wait(full_s + 8 * s, ph);
      const uint32_t src = my_s + s * kTileBytes;
      const bool live = FULL || (t * NT + gt < rows);  // rows past the map's end belong to the next map / nobody
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        v[u] = lds128(src + (((uint32_t)u << 4) ^ sw));
        if (!FULL && !live) v[u] = E::neg_inf_chunk();
      }
      // vertical (packed) maximum over the thread's run, then ONE horizontal step and ONE
      // running-maximum update per tile. The run maximum is canonicalised (+0.0f turns -0 into
      // +0) so that the update test is a bit comparison of max.NaN results: equal values, -0/+0
      // and NaN/NaN all keep the FIRST tile.
      uint32_t vm = E::vmax(v[0]);
#pragma unroll
      for (int u = 1; u < U; ++u) vm = E::vmerge(vm, E::vmax(v[u]));
      const float sm = E::vfinish(vm) + 0.0f;
      // every register of the run has been consumed by the maximum: hand the slot back to the
      // producer BEFORE the exponentials, so the refill overlaps the arithmetic
      __syncwarp();
      if (lane == 0) {
        if (SELF)
          issue(s, __float_as_uint(sm));  // refill the slot this warp has just emptied
        else
          mbar_arrive(empty_s + 8 * s);
      }
      if (++s == STAGES) {
        s = 0;
        ph ^= 1;
      }
      const float nm = max_nan_f32(run_max, sm);
      if (__float_as_uint(nm) != __float_as_uint(run_max)) {  // strictly better: park the run in shared memory
        run_tile = t;
#pragma unroll
        for (int u = 0; u < U; ++u) sts128(cand_s + u * NT * 16, v[u]);
      }
      run_max = nm;
      if (MSE) {
        if (live) {  // padding must not reach the loss
          const float gy = -lds32f(tab_s + (uint32_t)(Wp + iy) * 4);
          const f32x2 ngy = pack2(gy, gy);
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int h = 0; h < PER / 4; ++h) {
              const uint4 gx = lds128(tab_s + (uint32_t)(ix + u * PER + 4 * h) * 4);
              float lo, hi;
              E::pair(v[u], 2 * h, lo, hi);
              f32x2 d = fma2(pack2(__uint_as_float(gx.x), __uint_as_float(gx.y)), ngy, pack2(lo, hi));
              mse_acc = fma2(d, d, mse_acc);
              E::pair(v[u], 2 * h + 1, lo, hi);
              d = fma2(pack2(__uint_as_float(gx.z), __uint_as_float(gx.w)), ngy, pack2(lo, hi));
              mse_acc = fma2(d, d, mse_acc);
            }
        }
        ix += step_xi;
        iy += step_yi;
        if (ix >= p.W) {
          ix -= p.W;
          ++iy;
        }
      }
      // (A data-dependent shortcut was measured and rejected: publishing the group's running maximum in shared
      // memory and skipping tiles whose runs are all 32/beta below it lifts sharp-peak / large-beta maps to 6.9 TB/s
      // but costs every other regime 7 %, and — because the skip depends on when another warp's atomicMax lands —
      // it makes the last bits of kp_soft timing-dependent. Results here are bit-identical run to run and however
      // the frames are sharded.)
      if (MODE == MVGEO_SOFT_GLOBAL) {
        // Fold the epoch (rare, warp-uniform or nearly so): when a run climbs out of the window — first finite
        // run, climbing the peak — the reference moves; and every kFoldPeriod tiles unconditionally, so that an
        // accumulator that holds the peak's mass never takes more than kFoldPeriod background tiles on top
        // (bounds the f32 rounding of the long tail after a peak, wherever in the map the peak sits).
        const bool climb = sm > a.ref_hi;
        constexpr int kFoldTiles = kFoldPeriod * 4 / U;  // the same number of elements per fold for either run length
        if (climb || (t & (kFoldTiles - 1)) == kFoldTiles - 1) {
          epoch_fold<NT>(a, tot_s);
          if (climb) {
            a.nb = -sm * p.beta_log2e;
            a.ref_hi = sm + window;
          }
        }
        const f32x2 nb2 = pack2(a.nb, a.nb);
        // The run's 2-element groups i = 0 .. NP-1 (element 2i, 2i+1), processed as two halves so that the
        // dependent add chains stay short. Suffix sums t_i = w_i + ... + w_last of a half give its sum cs = t_0
        // and sum_i i*w_i = t_1 + t_2 + ...: adds only, no per-element constants. For the run:
        //   sum_j j*w_j = 2 * sum_i i*(w_i.lo + w_i.hi) + sum_i w_i.hi      (the last term is the odd half of a.s,
        //   added at the fold), and the upper half's groups carry an extra offset NP/2.
        constexpr int NP = EPR / 2, HP = NP / 2, PPC = PER / 2;  // groups per run / per half / per chunk
        f32x2 cs[2], sj[2];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
          for (int i = HP - 1; i >= 0; --i) {
            const int gi = hf * HP + i;  // group index in the run
            float lo, hi;
            E::pair(v[gi / PPC], gi % PPC, lo, hi);
            const f32x2 t2 = fma2(pack2(lo, hi), beta2, nb2);
            f32x2 w;
            if (kPolyEvery > 0 && E::kBytes == 2 && gi % (kPolyEvery > 0 ? kPolyEvery : 1) == kPolyEvery - 1) {
              w = exp2_poly2(t2);  // every kPolyEvery-th pair of a 16-bit run: off the MUFU pipe
            } else {
              unpack2(t2, lo, hi);
              w = pack2(ex2_approx(lo), ex2_approx(hi));  // -inf padding: weight 0
            }
            if (i == HP - 1) {
              cs[hf] = w;
            } else {
              sj[hf] = (i == HP - 2) ? cs[hf] : add2(sj[hf], cs[hf]);
              cs[hf] = add2(cs[hf], w);
            }
          }
        }
        const f32x2 ct = add2(cs[0], cs[1]);
        a.s = add2(a.s, ct);
        a.sj = add2(a.sj, add2(sj[0], sj[1]));
        a.sj = fma2(cs[1], pack2((float)HP, (float)HP), a.sj);
        a.sx = fma2(ct, pack2(fx, fx), a.sx);
        a.sy = fma2(ct, pack2(fy, fy), a.sy);
        fx += step_x;
        fy += step_y;
        if (fx >= x_hi) {
          fx -= Wf;
          fy += 1.0f;
        }
      }
    };
    for (int t = 0; t < n_full; ++t) tile_body(std::true_type{}, t);
    if (n_full < n_tiles) tile_body(std::false_type{}, n_full);

    // ------------------------------ per-map epilogue (consumers of this group only) ------------
    // 1. the maximum: one redux.sync per warp on order-preserving keys, then NW values through smem
    const uint32_t wk = __reduce_max_sync(0xffffffffu, ord_key(run_max));
    uint32_t mk = wk;
    if (NW > 1) {
      if (lane == 0) scr.key[lw] = wk;
      group_sync<NW>(bar);
      mk = scr.key[0];
#pragma unroll
      for (int w = 1; w < NW; ++w) mk = max(mk, scr.key[w]);
    }
    const float M = ord_val(mk);

    // 2. the first maximal element. Every thread parked its best run in shared memory; the first RUN of the
    //    warp that holds the maximum is one redux.sync on (tile, thread), and the first maximal ELEMENT of that
    //    run one ballot: lane j looks at element j (instead of one lane scanning its 16 / 32 elements while
    //    the warp waits). An all -inf map parked nothing: every element is maximal and index 0 wins, below.
    const bool isn = (M != M);
    const bool holds = run_tile >= 0 && M != kNegInf &&
                       (isn ? (run_max != run_max) : (__float_as_uint(run_max) == __float_as_uint(M)));
    const int first_run = __reduce_min_sync(0xffffffffu, holds ? run_tile * NT + gt : 0x7fffffff);
    int my_idx = 0x7fffffff;
    if (first_run != 0x7fffffff) {  // warp-uniform
      __syncwarp();                 // the parked run was written by another lane
      // lane j looks at elements j, j + 32, ... of that run (a run holds 16 / 32 / 64 elements)
      unsigned hits = 0;
      int base = 0;
#pragma unroll
      for (int k = 0; k < EPR; k += 32) {
        bool hit = false;
        const int j = k + lane;
        if (j < EPR) {
          const uint32_t addr = smem_u32(dyn_smem) + G * STAGES * kTileBytes +
                                (uint32_t)(g * kTile + (j / PER) * NT + (first_run & (NT - 1))) * 16 +
                                (uint32_t)(j % PER) * E::kBytes;
          const float e = lds_elem<DT>(addr);
          hit = isn ? (e != e) : (e == M);
        }
        const unsigned b = __ballot_sync(0xffffffffu, hit);
        if (hits == 0 && b != 0) {
          hits = b;
          base = k;
        }
      }
      my_idx = first_run * EPR + base + __ffs(hits) - 1;
    }
    const int wi = my_idx;  // warp-uniform
    if (NW > 1 && lane == 0) scr.idx[lw] = wi;
    float mse_w = 0.f;
    if (MSE) {
      mse_w = warp_sum(sum2(mse_acc));
      if (NW > 1 && lane == 0) scr.sum[0][lw] = mse_w;
    }
    // 3. soft-arg-max sums, rescaled from the thread's reference to the true maximum. The moments are
    //    still relative to the map centre (the peak position is not reduced yet) and the thread that holds
    //    the peak carries a lever arm of hundreds of pixels that cancels in the end: the per-map
    //    reduction runs in double (a few instructions per thread and MAP, nothing per element).
    float ss = 0.f, sx = 0.f, sy = 0.f;
    double ds = 0.0, dx = 0.0, dy = 0.0;
    if (MODE == MVGEO_SOFT_GLOBAL) {
      epoch_fold<NT>(a, tot_s);  // the last epoch
      const SoftTotals t = totals_load<NT>(tot_s);
      if (t.S > 0.0) {
        // 2^((ref - M) beta') formed with the rounded nb the weights were formed with: its rounding cancels
        const double r = (double)ex2_approx(fmaf(-M, p.beta_log2e, -t.nb));
        ds = t.S * r;
        dx = t.SX * r;
        dy = t.SY * r;
      }
      ds = warp_sum_f64(ds);
      dx = warp_sum_f64(dx);
      dy = warp_sum_f64(dy);
      if (NW > 1 && lane == 0) {
        scr.dsum[0][lw] = ds;
        scr.dsum[1][lw] = dx;
        scr.dsum[2][lw] = dy;
      }
    }
    int best = wi;
    if (NW > 1) {
      group_sync<NW>(bar);
      best = scr.idx[0];
#pragma unroll
      for (int w = 1; w < NW; ++w) best = min(best, scr.idx[w]);
    }
    if (best == 0x7fffffff) best = 0;  // all -inf map: nothing was ever parked; every element is maximal, the first wins
    const int py = best / p.W, px = best - py * p.W;
    if (MODE == MVGEO_SOFT_WINDOW) {
      window_accumulate<DT, NT>(p, map_ptr(p, map), M, px, py, gt, ss, sx, sy);
      block_sum3<NW>(ss, sx, sy, scr, bar, lw);
    } else if (MODE == MVGEO_SOFT_GLOBAL) {
      if (gt == 0) {
        if (NW > 1) {
          ds = dx = dy = 0.0;
#pragma unroll
          for (int w = 0; w < NW; ++w) {  // fixed order: deterministic
            ds += scr.dsum[0][w];
            dx += scr.dsum[1][w];
            dy += scr.dsum[2][w];
          }
        }
        // shift the first moments from the map centre to the hard peak, then leave double
        ss = (float)ds;
        sx = (float)(dx + (double)(ox - (float)px) * ds);
        sy = (float)(dy + (double)(oy - (float)py) * ds);
      }
    }
    if (MSE && gt == 0) {
      float m = mse_w;
      if (NW > 1) {
        m = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) m += scr.sum[0][w];  // fixed order: deterministic
      }
      p.mse_partial[map] = m;
    }
    if (gt == 0) write_outputs(p, map, M, best, ss, sx, sy, MODE != MVGEO_SOFT_NONE);
    // scr.key is rewritten only after the next map's streaming loop; scr.idx / scr.sum only after
    // the next map's first barrier: no trailing barrier needed.
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
  const char* base = map_ptr(p, map);

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

// Streaming-kernel configuration. A thread owns a run of U 16-byte chunks per tile: U = 8 (a whole 128-byte TMA row,
// 2-stage ring of 8 KB / 4 KB tiles) wherever the image rows hold whole 128-byte runs — the per-tile fixed costs
// (mbarrier wait, maximum finish, parking, position, fold test, slot hand-back: ~85 of 218 instructions) are paid
// per 128 bytes instead of 64: +6 % at C2, +8 % at C5, +10 % on 32 KB maps; U = 4 (half a row, 4-stage ring)
// for global soft mode / the fused loss pass on maps whose rows hold whole 64-byte runs only.
#ifndef MVGEO_G8_MAX
#define MVGEO_G8_MAX (32 * 1024)
#endif
#ifndef MVGEO_G4_MAX
#define MVGEO_G4_MAX (2 * 1024 * 1024)
#endif
#ifndef MVGEO_G2_MAX
#define MVGEO_G2_MAX (8 * 1024 * 1024)
#endif
constexpr int ring_stages(int U) { return U == 8 ? 2 : 4; }
// per CTA: the rings, one tile of parked candidate runs and the 32-byte soft-arg-max totals per thread
constexpr size_t ring_bytes(int U) { return (size_t)(ring_stages(U) + 1) * U * kDecThreads * 16 + (size_t)kDecThreads * 32; }
constexpr int kMaxDevices = 64;


// cuTensorMapEncodeTiled through the runtime (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
  static std::atomic<void*> cached{nullptr};  // a pure function of the driver: racing threads store the same pointer
  void* f = cached.load(std::memory_order_acquire);
  if (!f) {
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || !f)
      return nullptr;
    cached.store(f, std::memory_order_release);
  }
  return reinterpret_cast<EncodeTiledFn>(f);
}

// One tensor map per base pointer: [rows of 128 bytes] x [128 bytes], box = box_rows x 128 bytes, 128-byte swizzle.
static int build_tensor_maps(const DecodeParams& p, int box_rows, TensorMaps& tms) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc) return MVGEO_EUNSUPPORTED;
  memset(&tms, 0, sizeof(tms));
  const int64_t maps_per_view = p.n_views == 1 ? p.n_maps : p.n_maps / p.n_views;
  const uint64_t rows_total = (uint64_t)maps_per_view * (uint64_t)(p.map_bytes / kRowBytes);
  if (rows_total > 0x7fffffffull) return MVGEO_EUNSUPPORTED;  // TMA coordinates are 32-bit signed (137 GB per view)
  for (int v = 0; v < p.n_views; ++v) {
    const cuuint64_t gdim[2] = {(cuuint64_t)kRowBytes, (cuuint64_t)rows_total};
    const cuuint64_t gstride[1] = {(cuuint64_t)kRowBytes};
    const cuuint32_t box[2] = {(cuuint32_t)kRowBytes, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&tms.m[v], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(p.view_base[v]), gdim, gstride, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return MVGEO_EINVAL;
  }
  return MVGEO_OK;
}

constexpr size_t kMseTableMax = 16 * 1024;  // shared memory of the fused loss pass's Gaussian tables (2 CTAs/SM still fit)

template <int DT, int MODE, int G, int U, bool MSE = false>
static int launch_persistent(const DecodeParams& p, cudaStream_t st) {
  auto kern = decode_tma_kernel<DT, MODE, U, ring_stages(U), G, MSE>;
  // Resident-wave size per (instantiation, device): a pure function of its key, cached because the
  // occupancy query costs microseconds on a latency-bound call. The dynamic shared-memory opt-in is
  // set to the SAME constant by every caller (the ring never changes size), so concurrent callers
  // cannot disturb one another; a racing thread recomputes and stores the same value.
  constexpr size_t kSmemMax = ring_bytes(U) + (MSE ? kMseTableMax : 0);
  const size_t smem = kSmemMax;  // constant per instantiation: the cached occupancy is exact
  if (MSE && (size_t)G * (((p.W + 3) & ~3) + ((p.H + 3) & ~3)) * 4 > kMseTableMax)
    return MVGEO_EUNSUPPORTED;  // very wide / tall maps: use the two separate passes
  TensorMaps tms;
  const int rc = build_tensor_maps(p, kDecThreads / G * (U * 16) / kRowBytes, tms);
  if (rc) return rc;
  static std::atomic<int> cache[kMaxDevices];
  int dev = 0;
  MVGEO_CUDA(cudaGetDevice(&dev));
  int resident_ctas = (dev >= 0 && dev < kMaxDevices) ? cache[dev].load(std::memory_order_acquire) : 0;
  if (resident_ctas == 0) {
    MVGEO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    int sms = 0, per_sm = 0;
    MVGEO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MVGEO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kDecThreads + producer_threads(G), kSmemMax));
    if (per_sm < 1) return MVGEO_EUNSUPPORTED;
    resident_ctas = sms * per_sm;
    if (dev >= 0 && dev < kMaxDevices) cache[dev].store(resident_ctas, std::memory_order_release);
  }
  const int64_t wanted = (p.n_maps + G - 1) / G;
  const unsigned grid = (unsigned)(wanted < resident_ctas ? wanted : resident_ctas);
  kern<<<grid, kDecThreads + producer_threads(G), smem, st>>>(p, tms);
  MVGEO_CHECK_LAUNCH();
  return MVGEO_OK;
}

template <int DT, int MODE, int U>
static int launch_groups(const DecodeParams& p, int groups, cudaStream_t st) {
  switch (groups) {
    case 8: return launch_persistent<DT, MODE, 8, U>(p, st);
    case 4: return launch_persistent<DT, MODE, 4, U>(p, st);
    case 2: return launch_persistent<DT, MODE, 2, U>(p, st);
    default: return launch_persistent<DT, MODE, 1, U>(p, st);
  }
}

// run_chunks: 16-byte chunks per thread and tile of the streaming kernel (8 or 4), 0 = the element-wise kernel
template <int DT, int MODE>
static int launch_decode(const DecodeParams& p, int run_chunks, int groups, cudaStream_t st) {
  if (run_chunks == 0) {
    decode_scalar_kernel<DT, MODE><<<(unsigned)p.n_maps, kDecThreads, 0, st>>>(p);
    MVGEO_CHECK_LAUNCH();
    return MVGEO_OK;
  }
  if constexpr (MODE == MVGEO_SOFT_GLOBAL) {  // only the online soft-arg-max needs runs that stay inside an image row
    if (run_chunks == 4) return launch_groups<DT, MODE, 4>(p, groups, st);
  }
  return launch_groups<DT, MODE, 8>(p, groups, st);
}

template <int DT>
static int dispatch_mode(const DecodeParams& p, int mode, int run_chunks, int groups, cudaStream_t st) {
  switch (mode) {
    case MVGEO_SOFT_NONE: return launch_decode<DT, MVGEO_SOFT_NONE>(p, run_chunks, groups, st);
    case MVGEO_SOFT_GLOBAL: return launch_decode<DT, MVGEO_SOFT_GLOBAL>(p, run_chunks, groups, st);
    case MVGEO_SOFT_WINDOW: return launch_decode<DT, MVGEO_SOFT_WINDOW>(p, run_chunks, groups, st);
  }
  return MVGEO_EINVAL;
}

static int decode_impl(const void* const* view_maps, int n_views, int k_per_view, int dtype, int64_t n_maps, int H,
                       int W, double scale_x, double scale_y, int soft_mode, float beta, int window_radius,
                       int apply_sigmoid, int64_t k_inner, int64_t out_stride, int64_t out_offset, int32_t* idx,
                       float* peak, float* score, float* kp_hard, float* kp_soft, void* stream) {
  if (n_maps < 0 || H <= 0 || W <= 0 || k_inner <= 0 || out_stride < 0 || out_offset < 0) return MVGEO_EINVAL;
  if ((int64_t)H * W > (int64_t)1 << 30) return MVGEO_EINVAL;
  if (dtype != MVGEO_F32 && dtype != MVGEO_BF16 && dtype != MVGEO_F16) return MVGEO_EINVAL;
  if (soft_mode < MVGEO_SOFT_NONE || soft_mode > MVGEO_SOFT_WINDOW) return MVGEO_EINVAL;
  if (soft_mode != MVGEO_SOFT_NONE && !(beta > 0.f)) return MVGEO_EINVAL;
  if (soft_mode == MVGEO_SOFT_WINDOW && (window_radius < 0 || window_radius > MVGEO_MAX_WINDOW_RADIUS)) return MVGEO_EINVAL;
  if (n_views < 1 || n_views > MVGEO_MAX_VIEWS || k_per_view < 1) return MVGEO_EINVAL;
  if (n_maps == 0) return MVGEO_OK;
  if (!view_maps) return MVGEO_ENULL;
  if (n_maps > (int64_t)0x7fffffff) return MVGEO_EINVAL;

  const int esize = dtype == MVGEO_F32 ? 4 : 2;
  DecodeParams p = {};
  p.map_bytes = (int64_t)H * W * esize;
  // Work decomposition is a function of the map shape only (never of n_maps or the data), so results
  // are bit-identical however the frames are sharded across GPUs.
  // streaming kernel: maps made of whole 128-byte TMA rows, every thread a run of 128 bytes per tile; global soft
  // mode: runs that never straddle two image rows — 128-byte runs if the rows allow, else 64-byte runs
  int run_chunks = (p.map_bytes % kRowBytes == 0) ? 8 : 0;
  if (run_chunks && soft_mode == MVGEO_SOFT_GLOBAL)
    run_chunks = (W * esize) % 128 == 0 ? 8 : (W * esize) % 64 == 0 ? 4 : 0;
  for (int v = 0; v < n_views; ++v) {
    if (!view_maps[v]) return MVGEO_ENULL;
    p.view_base[v] = view_maps[v];
    if (reinterpret_cast<uintptr_t>(view_maps[v]) & 15) run_chunks = 0;
  }
  p.n_views = n_views;
  p.k_per_view = k_per_view;
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
  p.rows_per_map = run_chunks ? (int)(p.map_bytes / (run_chunks * 16)) : 0;
  // several consumer groups per CTA, one map stream each (one group's per-map epilogue overlaps the others'
  // streaming): 4 groups up to 2 MB maps (every BASELINE config; measured +3 % at C2 and +5 % at C5 over 2 / 1
  // groups), 2 up to 8 MB, one group (all 8 consumer warps on one map) beyond — a function of the map size only.
  const int groups = p.map_bytes <= MVGEO_G8_MAX ? 8 : p.map_bytes <= MVGEO_G4_MAX ? 4 : (p.map_bytes <= MVGEO_G2_MAX ? 2 : 1);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (dtype) {
    case MVGEO_F32: return dispatch_mode<MVGEO_F32>(p, soft_mode, run_chunks, groups, st);
    case MVGEO_BF16: return dispatch_mode<MVGEO_BF16>(p, soft_mode, run_chunks, groups, st);
    case MVGEO_F16: return dispatch_mode<MVGEO_F16>(p, soft_mode, run_chunks, groups, st);
  }
  return MVGEO_EINVAL;
}

}  // namespace mvgeo

using namespace mvgeo;

extern "C" int mvgeo_decode(const void* maps, int dtype, int64_t n_maps, int H, int W, double scale_x, double scale_y,
                            int soft_mode, float beta, int window_radius, int apply_sigmoid, int64_t k_inner,
                            int64_t out_stride, int64_t out_offset, int32_t* idx, float* peak, float* score,
                            float* kp_hard, float* kp_soft, void* stream) {
  const void* one[1] = {maps};  // NULL is reported by decode_impl after the size / enum checks
  return decode_impl(one, 1, 1, dtype, n_maps, H, W, scale_x, scale_y, soft_mode, beta, window_radius, apply_sigmoid,
                     k_inner, out_stride, out_offset, idx, peak, score, kp_hard, kp_soft, stream);
}

namespace mvgeo {
int finish_mse(const float* partial, int64_t n_maps, double N, float weight, float* loss, cudaStream_t st);  // encode.cu

// the fused loss pass keeps 64-byte runs: its Gaussian tables (16 KB) and the 128-byte-run ring do not fit twice per SM
constexpr int kMseU = 4;
template <int DT>
static int launch_decode_mse(const DecodeParams& p, int groups, cudaStream_t st) {
  switch (groups) {
    case 8: return launch_persistent<DT, MVGEO_SOFT_NONE, 8, kMseU, true>(p, st);
    case 4: return launch_persistent<DT, MVGEO_SOFT_NONE, 4, kMseU, true>(p, st);
    case 2: return launch_persistent<DT, MVGEO_SOFT_NONE, 2, kMseU, true>(p, st);
    default: return launch_persistent<DT, MVGEO_SOFT_NONE, 1, kMseU, true>(p, st);
  }
}
}  // namespace mvgeo

extern "C" int mvgeo_decode_mse(const void* maps, int dtype, int64_t n_maps, int H, int W, double scale_x, double scale_y,
                                int apply_sigmoid, const float* kp_target, float sigma, float weight, int32_t* idx,
                                float* peak, float* score, float* kp_hard, float* partial, float* loss, void* stream) {
  if (n_maps < 0 || H <= 0 || W <= 0 || !(sigma > 0.f)) return MVGEO_EINVAL;
  if ((int64_t)H * W > (int64_t)1 << 30 || n_maps > (int64_t)0x7fffffff) return MVGEO_EINVAL;
  if (dtype != MVGEO_F32 && dtype != MVGEO_BF16 && dtype != MVGEO_F16) return MVGEO_EINVAL;
  if (n_maps == 0) return MVGEO_OK;
  if (!maps || !kp_target || !partial || !loss) return MVGEO_ENULL;
  const int esize = dtype == MVGEO_F32 ? 4 : 2;
  // the fused pass is the streaming kernel only: 16-byte aligned maps of whole 128-byte TMA rows whose image rows
  // hold whole 64-byte runs
  if ((reinterpret_cast<uintptr_t>(maps) & 15) || (W * esize) % (kMseU * 16) != 0 || ((int64_t)H * W * esize) % kRowBytes != 0)
    return MVGEO_EUNSUPPORTED;
  DecodeParams p = {};
  p.view_base[0] = maps;
  p.n_views = 1;
  p.k_per_view = 1;
  p.n_maps = n_maps;
  p.map_bytes = (int64_t)H * W * esize;
  p.H = H;
  p.W = W;
  p.rows_per_map = (int)(p.map_bytes / (kMseU * 16));
  p.scale_x = scale_x;
  p.scale_y = scale_y;
  p.apply_sigmoid = apply_sigmoid;
  p.k_inner = 1;
  p.out_stride = 1;
  p.idx = idx;
  p.peak = peak;
  p.score = score;
  p.kp_hard = kp_hard;
  p.mse_kp = kp_target;
  p.mse_k = kLog2e / (2.0f * sigma * sigma);
  p.mse_partial = partial;
  int groups = p.map_bytes <= MVGEO_G8_MAX ? 8 : p.map_bytes <= MVGEO_G4_MAX ? 4 : (p.map_bytes <= MVGEO_G2_MAX ? 2 : 1);
  // every consumer group holds its own pair of Gaussian tables: fewer groups for wide / tall maps (shape only)
  while (groups > 1 && (size_t)groups * (((W + 3) & ~3) + ((H + 3) & ~3)) * 4 > kMseTableMax) groups >>= 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc;
  switch (dtype) {
    case MVGEO_F32: rc = launch_decode_mse<MVGEO_F32>(p, groups, st); break;
    case MVGEO_BF16: rc = launch_decode_mse<MVGEO_BF16>(p, groups, st); break;
    default: rc = launch_decode_mse<MVGEO_F16>(p, groups, st); break;
  }
  if (rc) return rc;
  return finish_mse(partial, n_maps, (double)n_maps * H * W, weight, loss, st);
}

extern "C" int mvgeo_decode_views(const void* const* view_maps, int n_views, int dtype, int64_t B, int K, int H, int W,
                                  double scale_x, double scale_y, int soft_mode, float beta, int window_radius,
                                  int apply_sigmoid, int32_t* idx, float* peak, float* score, float* kp_hard,
                                  float* kp_soft, void* stream) {
  if (B < 0 || K < 1 || n_views < 1 || n_views > MVGEO_MAX_VIEWS) return MVGEO_EINVAL;
  return decode_impl(view_maps, n_views, K, dtype, B * n_views * K, H, W, scale_x, scale_y, soft_mode, beta,
                     window_radius, apply_sigmoid, 1, 1, 0, idx, peak, score, kp_hard, kp_soft, stream);
}
