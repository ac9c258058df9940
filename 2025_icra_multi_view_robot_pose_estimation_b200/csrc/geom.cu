// geom.cu — the geometry tail of the pipeline in ONE launch: DLT triangulation of every key-point
// and FK + reprojection consistency of every (frame, view), plus the deterministic loss sum.
//
// The two stages are independent of each other (both consume the decoded key-points), so the
// grid simply holds both: CTAs [0, n_dlt) run dlt_body, CTAs [n_dlt, n_dlt + n_fk) run
// fk_reproj_fwd_body. The FK CTA that finishes last (atomic ticket after a __threadfence) sums
// frame_loss[0..B) in a fixed order and resets the ticket, so the scalar loss needs no third
// launch and no float atomics. Replaces three latency-bound launches (triangulate, fk_reproj_fwd,
// sum) of ~8-15 us each by one; the stand-alone entry points remain for callers that need only
// one stage.
#include "dlt_device.cuh"
#include "fk_device.cuh"

namespace mvgeo {

static_assert(kDltThreads == kFkThreads, "both roles share one block size");

template <bool BASE>
__global__ void __launch_bounds__(kFkThreads)
    geometry_kernel(const float* __restrict__ kp, const float* __restrict__ wgt, const float* __restrict__ Pm,
                    int64_t B, int V, int K, float min_weight, int weighted, float* __restrict__ X_tri,
                    float* __restrict__ resid, int32_t* __restrict__ n_views, const mvgeo_chain ch,
                    const float* __restrict__ q, const float* __restrict__ R_view,
                    const mvgeo_camera* __restrict__ cams, float scale, float* __restrict__ X_fk,
                    float* __restrict__ uv_fk, float* frame_loss, float* __restrict__ loss, unsigned int* ticket,
                    int n_dlt, int n_fk) {
  __shared__ float sP[MVGEO_MAX_VIEWS * 12];
  __shared__ float part[kFkThreads];
  __shared__ bool is_last;
  if ((int)blockIdx.x < n_dlt) {
    for (int i = threadIdx.x; i < V * 12; i += kDltThreads) sP[i] = Pm[i];
    __syncthreads();
    dlt_body(kp, wgt, sP, B, V, K, min_weight, weighted, X_tri, resid, n_views, blockIdx.x);
    return;
  }
  // consistency of FK against the same key-points the triangulation used (unweighted, like mvgeo_pipeline)
  fk_reproj_fwd_body<BASE>(ch, q, B, R_view, cams, V, kp, nullptr, scale, X_fk, uv_fk, frame_loss, part,
                           (int64_t)blockIdx.x - n_dlt);
  if (!loss) return;
  // Several warps of this CTA wrote frame_loss entries: each writer fences its own stores, the
  // barrier makes sure EVERY writer of the CTA has done so, and only then is the ticket taken.
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == (unsigned)(n_fk - 1));
  __syncthreads();
  if (!is_last) return;
  const volatile float* fl = frame_loss;  // written by other CTAs: bypass L1
  float a = 0.f;
  for (int64_t i = threadIdx.x; i < B; i += kFkThreads) a += fl[i];
  __syncthreads();
  part[threadIdx.x] = a;
  __syncthreads();
  for (int s = kFkThreads / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    loss[0] = part[0];
    *ticket = 0u;  // ready for the next launch
  }
}

}  // namespace mvgeo

using namespace mvgeo;

extern "C" int mvgeo_geometry(const float* kp, const float* w, const float* P, const mvgeo_chain* chain,
                              const float* q, int64_t B, const float* R_view, const mvgeo_camera* cams, int V, int K,
                              float min_weight, int weighted, float lambda, float* X_tri, float* resid,
                              int32_t* n_views, float* X_fk, float* uv_fk, float* frame_loss, float* loss,
                              int32_t* ticket, void* stream) {
  if (!chain) return MVGEO_ENULL;
  if (chain->n_joints < 1 || chain->n_joints > MVGEO_MAX_JOINTS) return MVGEO_EINVAL;
  if (chain->convention != MVGEO_DH_STANDARD && chain->convention != MVGEO_DH_MODIFIED) return MVGEO_EINVAL;
  if (B < 0 || V < 1 || V > MVGEO_MAX_VIEWS || K != chain->n_joints + (chain->emit_base ? 1 : 0)) return MVGEO_EINVAL;
  if (B == 0) return MVGEO_OK;
  if (!kp || !P || !q || !cams || !X_tri || !frame_loss) return MVGEO_ENULL;
  if (loss && !ticket) return MVGEO_ENULL;
  const int64_t n_pts = B * K;
  const int64_t n_dlt = (n_pts + kDltThreads / 4 - 1) / (kDltThreads / 4);
  const int fpc = kFkThreads / V;
  const int64_t n_fk = (B + fpc - 1) / fpc;
  if (n_dlt + n_fk > 0x7fffffff) return MVGEO_EINVAL;
  const float scale = (float)((double)lambda / ((double)B * V * K * 2.0));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)(n_dlt + n_fk);
  unsigned int* tk = reinterpret_cast<unsigned int*>(ticket);
  if (chain->emit_base)
    geometry_kernel<true><<<grid, kFkThreads, 0, st>>>(kp, w, P, B, V, K, min_weight, weighted, X_tri, resid, n_views,
                                                       *chain, q, R_view, cams, scale, X_fk, uv_fk, frame_loss, loss,
                                                       tk, (int)n_dlt, (int)n_fk);
  else
    geometry_kernel<false><<<grid, kFkThreads, 0, st>>>(kp, w, P, B, V, K, min_weight, weighted, X_tri, resid,
                                                        n_views, *chain, q, R_view, cams, scale, X_fk, uv_fk,
                                                        frame_loss, loss, tk, (int)n_dlt, (int)n_fk);
  MVGEO_CHECK_LAUNCH();
  return MVGEO_OK;
}
