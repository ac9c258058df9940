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

// Gaussian value of PER consecutive pixels of one row starting at x (x % 4 == 0): two/one 16-byte
// table reads instead of PER scalar ones.
template <int PER>
__device__ __forceinline__ void row_values(const float* ex, int x, float vy, float thr, float* v) {
#pragma unroll
  for (int q = 0; q < PER / 4; ++q) {
    const float4 e = *reinterpret_cast<const float4*>(ex + x + 4 * q);
    const float g0 = e.x * vy, g1 = e.y * vy, g2 = e.z * vy, g3 = e.w * vy;
    v[4 * q + 0] = g0 < thr ? 0.f : g0;
    v[4 * q + 1] = g1 < thr ? 0.f : g1;
    v[4 * q + 2] = g2 < thr ? 0.f : g2;
    v[4 * q + 3] = g3 < thr ? 0.f : g3;
  }
}

// Same without the eps*max cut-off: in the MSE the cut-off only replaces values < 2.2e-16 by 0,
// which cannot change a float32 difference pred - g (the encoder keeps the cut-off: exact zeros
// are part of create_gt_heatmap's contract).
template <int PER>
__device__ __forceinline__ void row_values_nocut(const float* ex, int x, float vy, float* v) {
#pragma unroll
  for (int q = 0; q < PER / 4; ++q) {
    const float4 e = *reinterpret_cast<const float4*>(ex + x + 4 * q);
    v[4 * q + 0] = e.x * vy;
    v[4 * q + 1] = e.y * vy;
    v[4 * q + 2] = e.z * vy;
    v[4 * q + 3] = e.w * vy;
  }
}

__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// VEC: 16-byte stores; ROW additionally W % PER == 0, so a chunk never straddles two rows.
template <int DT, bool VEC, bool ROW>
__global__ void __launch_bounds__(kEncThreads) encode_kernel(const float* __restrict__ kp, int64_t n_maps, int H,
                                                             int W, float inv_2s2_log2e, void* __restrict__ maps) {
  using E = Elem<DT>;
  extern __shared__ __align__(16) float tab[];
  float* ex = tab;
  float* ey = tab + ((W + 3) & ~3);
  const int n = H * W;
  for (int64_t map = blockIdx.x; map < n_maps; map += gridDim.x) {
    float thr;
    gaussian_tables(kp, map, H, W, inv_2s2_log2e, ex, ey, thr);
    if (VEC) {
      constexpr int PER = E::kPerChunk;
      uint4* out = reinterpret_cast<uint4*>(maps) + map * (int64_t)(n / PER);
      const int cpr = W / PER;  // chunks per row (ROW only)
      for (int c = threadIdx.x; c < n / PER; c += kEncThreads) {
        float v[PER];
        if (ROW) {
          const int y = c / cpr, x = (c - y * cpr) * PER;
          row_values<PER>(ex, x, ey[y], thr, v);
        } else {
          const int flat0 = c * PER;
          int y = flat0 / W, x = flat0 - y * W;
          float vy = ey[y];
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
        }
        st_stream(out + c, pack_chunk<DT>(v));
      }
    } else {
      char* out = reinterpret_cast<char*>(maps) + map * (int64_t)n * E::kBytes;
      for (int i = threadIdx.x; i < n; i += kEncThreads) {
        const int y = i / W, x = i - y * W;
        const float g = ex[x] * ey[y];
        E::store(out, i, g < thr ? 0.f : g);
      }
    }
    __syncthreads();  // tables are rebuilt for the next map
  }
}

// pred - gaussian(kp): per-map sum of squares (fixed-order reduction) and optional gradient.
// Four 16-byte loads in flight per thread (the loads of a batch are issued before any is used).
template <int DT, bool VEC, bool ROW, bool GRAD>
__global__ void __launch_bounds__(kEncThreads, 4)
    mse_kernel(const void* __restrict__ pred, const float* __restrict__ kp, int64_t n_maps, int H, int W,
               float inv_2s2_log2e, float grad_scale, const float* __restrict__ dloss, float* __restrict__ partial,
               void* __restrict__ grad) {
  using E = Elem<DT>;
  extern __shared__ __align__(16) float tab[];
  __shared__ float red[kEncWarps];
  if (GRAD && dloss) grad_scale *= dloss[0];  // upstream gradient of the scalar loss, read on the device
  float* ex = tab;
  float* ey = tab + ((W + 3) & ~3);
  const int n = H * W;
  for (int64_t map = blockIdx.x; map < n_maps; map += gridDim.x) {
    float thr;
    gaussian_tables(kp, map, H, W, inv_2s2_log2e, ex, ey, thr);
    float acc = 0.f;
    if (VEC) {
      constexpr int PER = E::kPerChunk;
      constexpr int UN = 4;
      const int nc = n / PER, cpr = ROW ? W / PER : 1;
      const uint4* in = reinterpret_cast<const uint4*>(pred) + map * (int64_t)nc;
      uint4* gout = GRAD ? reinterpret_cast<uint4*>(grad) + map * (int64_t)nc : nullptr;
      for (int c0 = threadIdx.x; c0 < nc; c0 += kEncThreads * UN) {
        uint4 ch[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const int c = c0 + u * kEncThreads;
          ch[u] = c < nc ? ld_stream(in + c) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const int c = c0 + u * kEncThreads;
          if (c < nc) {
            float g[PER], gv[PER];
            if (ROW) {
              const int y = c / cpr, x = (c - y * cpr) * PER;
              row_values_nocut<PER>(ex, x, ey[y], g);
            } else {
              const int flat0 = c * PER;
              int y = flat0 / W, x = flat0 - y * W;
              float vy = ey[y];
#pragma unroll
              for (int j = 0; j < PER; ++j) {
                g[j] = ex[x] * vy;
                if (++x == W) {
                  x = 0;
                  ++y;
                  vy = y < H ? ey[y] : 0.f;
                }
              }
            }
#pragma unroll
            for (int j = 0; j < PER; ++j) {
              const float d = E::get(ch[u], j) - g[j];
              acc += d * d;
              gv[j] = d * grad_scale;
            }
            if (GRAD) st_stream(gout + c, pack_chunk<DT>(gv));
          }
        }
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
    __syncthreads();  // red[] and the tables are reused for the next map
  }
}

__global__ void scale_kernel(float* x, float s) { x[0] *= s; }

// Grid: enough CTAs for ~8 per SM; CTAs walk maps blockIdx.x, +gridDim.x (tables rebuilt per map).
static unsigned maps_grid(int64_t n_maps) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t cap = (int64_t)sms * 8;
  return (unsigned)(n_maps < cap ? n_maps : cap);
}

template <int DT>
static int launch_encode(const float* kp, int64_t n_maps, int H, int W, float k, void* maps, bool vec, cudaStream_t st) {
  const size_t smem = (size_t)(((W + 3) & ~3) + H) * sizeof(float);
  const unsigned g = maps_grid(n_maps);
  const bool row = vec && (W % Elem<DT>::kPerChunk == 0);
  if (row) encode_kernel<DT, true, true><<<g, kEncThreads, smem, st>>>(kp, n_maps, H, W, k, maps);
  else if (vec) encode_kernel<DT, true, false><<<g, kEncThreads, smem, st>>>(kp, n_maps, H, W, k, maps);
  else encode_kernel<DT, false, false><<<g, kEncThreads, smem, st>>>(kp, n_maps, H, W, k, maps);
  MVGEO_CHECK_LAUNCH();
  return MVGEO_OK;
}

template <int DT, bool VEC, bool ROW>
static void launch_mse2(const void* pred, const float* kp, int64_t n_maps, int H, int W, float k, float gs,
                        const float* dloss, float* partial, void* grad, unsigned g, size_t smem, cudaStream_t st) {
  if (grad)
    mse_kernel<DT, VEC, ROW, true><<<g, kEncThreads, smem, st>>>(pred, kp, n_maps, H, W, k, gs, dloss, partial, grad);
  else
    mse_kernel<DT, VEC, ROW, false><<<g, kEncThreads, smem, st>>>(pred, kp, n_maps, H, W, k, gs, dloss, partial, grad);
}

template <int DT>
static int launch_mse(const void* pred, const float* kp, int64_t n_maps, int H, int W, float k, float gs,
                      const float* dloss, float* partial, void* grad, bool vec, cudaStream_t st) {
  const size_t smem = (size_t)(((W + 3) & ~3) + H) * sizeof(float);
  const unsigned g = maps_grid(n_maps);
  const bool row = vec && (W % Elem<DT>::kPerChunk == 0);
  if (row) launch_mse2<DT, true, true>(pred, kp, n_maps, H, W, k, gs, dloss, partial, grad, g, smem, st);
  else if (vec) launch_mse2<DT, true, false>(pred, kp, n_maps, H, W, k, gs, dloss, partial, grad, g, smem, st);
  else launch_mse2<DT, false, false>(pred, kp, n_maps, H, W, k, gs, dloss, partial, grad, g, smem, st);
  MVGEO_CHECK_LAUNCH();
  return MVGEO_OK;
}

// loss[0] = weight / N * sum(partial): shared by mvgeo_heatmap_mse and the fused decode + MSE pass (decode.cu)
int finish_mse(const float* partial, int64_t n_maps, double N, float weight, float* loss, cudaStream_t st) {
  int rc = launch_sum(partial, n_maps, loss, st);
  if (rc) return rc;
  scale_kernel<<<1, 1, 0, st>>>(loss, (float)((double)weight / N));
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
                                 float sigma, float weight, const float* dloss, float* partial, float* loss, void* grad,
                                 void* stream) {
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
    case MVGEO_F32: rc = launch_mse<MVGEO_F32>(pred, kp, n_maps, H, W, k, gs, dloss, partial, grad, vec, st); break;
    case MVGEO_BF16: rc = launch_mse<MVGEO_BF16>(pred, kp, n_maps, H, W, k, gs, dloss, partial, grad, vec, st); break;
    default: rc = launch_mse<MVGEO_F16>(pred, kp, n_maps, H, W, k, gs, dloss, partial, grad, vec, st); break;
  }
  if (rc) return rc;
  return finish_mse(partial, n_maps, N, weight, loss, st);
}
