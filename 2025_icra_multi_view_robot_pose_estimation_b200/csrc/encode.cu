// encode.cu — Gaussian belief-map encoder and fused heat-map MSE loss (fwd + bwd), sm_100a.
//
// Replaces create_gt_heatmap (model/MvRoPose_FR3.py:65-73, model/DREAM_Train.py:60-69) as it
// is used by RobotPoseDataset.__getitem__ (model/MvRoPose_FR3.py:214-222), and
// nn.MSELoss()(pred, gt) * loss_weight_kpt (model/MvRoPose_FR3.py:846-847,975) with the
// target rasterised on the fly instead of being materialised and copied host->device.
//
// Roofline: HBM. Encoder: H*W*sizeof bytes written per map, nothing read. MSE: H*W*sizeof
// read (+ the same written when a gradient is requested). The Gaussian is separable, so a CTA
// evaluates W + H exponentials into shared memory once per map and every pixel costs one
// shared-memory read and one multiply — an exp per pixel would make MUFU, not HBM, the limit
// for 16-bit maps (13 elements/clk/SM needed, 16/clk/SM available).
#include "common.cuh"

namespace mvgeo {

constexpr int kEncThreads = 256;
constexpr int kEncWarps = kEncThreads / 32;
constexpr int kEncMaxDim = 4096;  // H + W <= kEncMaxDim (shared-memory tables)

int launch_sum(const float* x, int64_t n, float* out, cudaStream_t st);  // fk.cu

template <int DT> __device__ __forceinline__ uint4 pack_chunk(const float* v);
template <> __device__ __forceinline__ uint4 pack_chunk<MVGEO_F32>(const float* v) {
  return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
}
template <> __device__ __forceinline__ uint4 pack_chunk<MVGEO_BF16>(const float* v) {
  uint4 r;
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  r.z = *reinterpret_cast<uint32_t*>(&c);
  r.w = *reinterpret_cast<uint32_t*>(&d);
  return r;
}
template <> __device__ __forceinline__ uint4 pack_chunk<MVGEO_F16>(const float* v) {
  uint4 r;
  __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
  __half2 c = __floats2half2_rn(v[4], v[5]), d = __floats2half2_rn(v[6], v[7]);
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  r.z = *reinterpret_cast<uint32_t*>(&c);
  r.w = *reinterpret_cast<uint32_t*>(&d);
  return r;
}

// Separable Gaussian tables for one map: ex[x] = exp(-(x-cx)^2 / 2 sigma^2), ey likewise, and
// the reference's cut-off `heatmap < eps(double) * heatmap.max()` (MvRoPose_FR3.py:72).
// Returns false (all-zero target) when the centre is not finite.
__device__ __forceinline__ bool gaussian_tables(const float* __restrict__ kp, int64_t map, int H, int W,
                                                float inv_2s2_log2e, float* ex, float* ey, float& thr) {
  const float cx = kp[2 * map], cy = kp[2 * map + 1];
  const bool ok = isfinite(cx) && isfinite(cy);
  for (int i = threadIdx.x; i < W + H; i += kEncThreads) {
    const float d = i < W ? (float)i - cx : (float)(i - W) - cy;
    const float e = ok ? ex2_approx(-d * d * inv_2s2_log2e) : 0.f;
    if (i < W) ex[i] = e; else ey[i - W] = e;
  }
  // max of the map = value at the pixel nearest to the centre (clamped into the map)
  const float nx = fminf(fmaxf(rintf(cx), 0.f), (float)(W - 1)) - cx;
  const float ny = fminf(fmaxf(rintf(cy), 0.f), (float)(H - 1)) - cy;
  thr = ok ? 2.220446049250313e-16f * ex2_approx(-(nx * nx + ny * ny) * inv_2s2_log2e) : 0.f;
  __syncthreads();
  return ok;
}

template <int DT, bool VEC>
__global__ void __launch_bounds__(kEncThreads) encode_kernel(const float* __restrict__ kp, int64_t n_maps, int H,
                                                             int W, float inv_2s2_log2e, void* __restrict__ maps) {
  using E = Elem<DT>;
  extern __shared__ float tab[];
  float* ex = tab;
  float* ey = tab + W;
  const int64_t map = blockIdx.x;
  float thr;
  gaussian_tables(kp, map, H, W, inv_2s2_log2e, ex, ey, thr);
  const int n = H * W;
  if (VEC) {
    constexpr int PER = E::kPerChunk;
    uint4* out = reinterpret_cast<uint4*>(maps) + map * (int64_t)(n / PER);
    for (int c = threadIdx.x; c < n / PER; c += kEncThreads) {
      const int flat0 = c * PER;
      int y = flat0 / W, x = flat0 - y * W;
      float vy = ey[y];
      float v[PER];
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const float g = ex[x] * vy;
        v[j] = g < thr ? 0.f : g;
        if (++x == W) {
          x = 0;
          ++y;
          vy = y < H ? ey[y] : 0.f;
        }
      }
      out[c] = pack_chunk<DT>(v);
    }
  } else {
    char* out = reinterpret_cast<char*>(maps) + map * (int64_t)n * E::kBytes;
    for (int i = threadIdx.x; i < n; i += kEncThreads) {
      const int y = i / W, x = i - y * W;
      const float g = ex[x] * ey[y];
      E::store(out, i, g < thr ? 0.f : g);
    }
  }
}

// pred - gaussian(kp): per-map sum of squares (fixed-order reduction) and optional gradient.
template <int DT, bool VEC, bool GRAD>
__global__ void __launch_bounds__(kEncThreads)
    mse_kernel(const void* __restrict__ pred, const float* __restrict__ kp, int64_t n_maps, int H, int W,
               float inv_2s2_log2e, float grad_scale, float* __restrict__ partial, void* __restrict__ grad) {
  using E = Elem<DT>;
  extern __shared__ float tab[];
  __shared__ float red[kEncWarps];
  float* ex = tab;
  float* ey = tab + W;
  const int64_t map = blockIdx.x;
  float thr;
  gaussian_tables(kp, map, H, W, inv_2s2_log2e, ex, ey, thr);
  const int n = H * W;
  float acc = 0.f;
  if (VEC) {
    constexpr int PER = E::kPerChunk;
    const uint4* in = reinterpret_cast<const uint4*>(pred) + map * (int64_t)(n / PER);
    uint4* gout = GRAD ? reinterpret_cast<uint4*>(grad) + map * (int64_t)(n / PER) : nullptr;
    for (int c = threadIdx.x; c < n / PER; c += kEncThreads) {
      const uint4 ch = ld_stream(in + c);
      const int flat0 = c * PER;
      int y = flat0 / W, x = flat0 - y * W;
      float vy = ey[y];
      float gv[PER];
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        float g = ex[x] * vy;
        g = g < thr ? 0.f : g;
        const float d = E::get(ch, j) - g;
        acc += d * d;
        gv[j] = d * grad_scale;
        if (++x == W) {
          x = 0;
          ++y;
          vy = y < H ? ey[y] : 0.f;
        }
      }
      if (GRAD) gout[c] = pack_chunk<DT>(gv);
    }
  } else {
    const char* in = reinterpret_cast<const char*>(pred) + map * (int64_t)n * E::kBytes;
    char* gout = GRAD ? reinterpret_cast<char*>(grad) + map * (int64_t)n * E::kBytes : nullptr;
    for (int i = threadIdx.x; i < n; i += kEncThreads) {
      const int y = i / W, x = i - y * W;
      float g = ex[x] * ey[y];
      g = g < thr ? 0.f : g;
      const float d = E::load(in, i) - g;
      acc += d * d;
      if (GRAD) E::store(gout, i, d * grad_scale);
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kEncWarps; ++w) t += red[w];
    partial[map] = t;
  }
}

__global__ void scale_kernel(float* x, float s) { x[0] *= s; }

template <int DT>
static int launch_encode(const float* kp, int64_t n_maps, int H, int W, float k, void* maps, bool vec, cudaStream_t st) {
  const size_t smem = (size_t)(H + W) * sizeof(float);
  if (vec) encode_kernel<DT, true><<<(unsigned)n_maps, kEncThreads, smem, st>>>(kp, n_maps, H, W, k, maps);
  else encode_kernel<DT, false><<<(unsigned)n_maps, kEncThreads, smem, st>>>(kp, n_maps, H, W, k, maps);
  MVGEO_CHECK_LAUNCH();
  return MVGEO_OK;
}

template <int DT>
static int launch_mse(const void* pred, const float* kp, int64_t n_maps, int H, int W, float k, float gs,
                      float* partial, void* grad, bool vec, cudaStream_t st) {
  const size_t smem = (size_t)(H + W) * sizeof(float);
  const unsigned g = (unsigned)n_maps;
  if (vec) {
    if (grad) mse_kernel<DT, true, true><<<g, kEncThreads, smem, st>>>(pred, kp, n_maps, H, W, k, gs, partial, grad);
    else mse_kernel<DT, true, false><<<g, kEncThreads, smem, st>>>(pred, kp, n_maps, H, W, k, gs, partial, grad);
  } else {
    if (grad) mse_kernel<DT, false, true><<<g, kEncThreads, smem, st>>>(pred, kp, n_maps, H, W, k, gs, partial, grad);
    else mse_kernel<DT, false, false><<<g, kEncThreads, smem, st>>>(pred, kp, n_maps, H, W, k, gs, partial, grad);
  }
  MVGEO_CHECK_LAUNCH();
  return MVGEO_OK;
}

static int check_maps_args(int64_t n_maps, int H, int W, int dtype, float sigma) {
  if (n_maps < 0 || H <= 0 || W <= 0 || H + W > kEncMaxDim || !(sigma > 0.f)) return MVGEO_EINVAL;
  if (dtype != MVGEO_F32 && dtype != MVGEO_BF16 && dtype != MVGEO_F16) return MVGEO_EINVAL;
  if (n_maps > 0x7fffffff) return MVGEO_EINVAL;
  return MVGEO_OK;
}

}  // namespace mvgeo

using namespace mvgeo;

extern "C" int mvgeo_encode_gaussian(const float* kp, int64_t n_maps, int H, int W, float sigma, int dtype, void* maps,
                                     void* stream) {
  int rc = check_maps_args(n_maps, H, W, dtype, sigma);
  if (rc) return rc;
  if (n_maps == 0) return MVGEO_OK;
  if (!kp || !maps) return MVGEO_ENULL;
  const int esize = dtype == MVGEO_F32 ? 4 : 2;
  const bool vec = ((int64_t)H * W * esize) % 16 == 0 && (reinterpret_cast<uintptr_t>(maps) & 15) == 0;
  const float k = kLog2e / (2.0f * sigma * sigma);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (dtype) {
    case MVGEO_F32: return launch_encode<MVGEO_F32>(kp, n_maps, H, W, k, maps, vec, st);
    case MVGEO_BF16: return launch_encode<MVGEO_BF16>(kp, n_maps, H, W, k, maps, vec, st);
    default: return launch_encode<MVGEO_F16>(kp, n_maps, H, W, k, maps, vec, st);
  }
}

extern "C" int mvgeo_heatmap_mse(const void* pred, int dtype, const float* kp, int64_t n_maps, int H, int W,
                                 float sigma, float weight, float* partial, float* loss, void* grad, void* stream) {
  int rc = check_maps_args(n_maps, H, W, dtype, sigma);
  if (rc) return rc;
  if (n_maps == 0) return MVGEO_OK;
  if (!pred || !kp || !partial || !loss) return MVGEO_ENULL;
  const int esize = dtype == MVGEO_F32 ? 4 : 2;
  const bool vec = ((int64_t)H * W * esize) % 16 == 0 && (reinterpret_cast<uintptr_t>(pred) & 15) == 0 &&
                   (!grad || (reinterpret_cast<uintptr_t>(grad) & 15) == 0);
  const float k = kLog2e / (2.0f * sigma * sigma);
  const double N = (double)n_maps * H * W;
  const float gs = (float)(2.0 * (double)weight / N);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (dtype) {
    case MVGEO_F32: rc = launch_mse<MVGEO_F32>(pred, kp, n_maps, H, W, k, gs, partial, grad, vec, st); break;
    case MVGEO_BF16: rc = launch_mse<MVGEO_BF16>(pred, kp, n_maps, H, W, k, gs, partial, grad, vec, st); break;
    default: rc = launch_mse<MVGEO_F16>(pred, kp, n_maps, H, W, k, gs, partial, grad, vec, st); break;
  }
  if (rc) return rc;
  rc = launch_sum(partial, n_maps, loss, st);
  if (rc) return rc;
  scale_kernel<<<1, 1, 0, st>>>(loss, (float)((double)weight / N));
  MVGEO_CHECK_LAUNCH();
  return MVGEO_OK;
}
