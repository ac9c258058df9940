// api.cu — library-level entry points: version / error strings, the fused device pipeline
// (decode -> triangulate -> FK + reprojection consistency on one stream, no host sync) and the
// host-buffer pipeline (pinned host memory in, chunked H2D overlapped with compute, D2H out).
#include <new>

#include "common.cuh"

using namespace mvgeo;

extern "C" int mvgeo_version(void) { return MVGEO_VERSION; }

extern "C" const char* mvgeo_error_string(int code) {
  switch (code) {
    case MVGEO_OK: return "ok";
    case MVGEO_EINVAL: return "invalid argument (size or enum out of range)";
    case MVGEO_ENULL: return "required pointer is NULL";
    case MVGEO_EALIGN: return "pointer is not aligned as documented";
    case MVGEO_EUNSUPPORTED: return "request not supported by this build";
    case MVGEO_ENOMEM: return "out of memory";
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown mvgeo error";
}

static int check_cfg(const mvgeo_pipeline_cfg* c, const mvgeo_chain* chain) {
  if (!c || !chain) return MVGEO_ENULL;
  if (c->V < 1 || c->V > MVGEO_MAX_VIEWS || c->K < 1 || c->H < 1 || c->W < 1) return MVGEO_EINVAL;
  if (chain->n_joints < 1 || chain->n_joints > MVGEO_MAX_JOINTS) return MVGEO_EINVAL;
  if (c->K != chain->n_joints + (chain->emit_base ? 1 : 0)) return MVGEO_EINVAL;
  return MVGEO_OK;
}

extern "C" int mvgeo_pipeline(const mvgeo_pipeline_cfg* cfg, const void* maps, int64_t B, const float* P,
                              const mvgeo_chain* chain, const float* q, const float* R_view,
                              const mvgeo_camera* cams, const mvgeo_pipeline_out* out, void* stream) {
  int rc = check_cfg(cfg, chain);
  if (rc) return rc;
  if (B < 0) return MVGEO_EINVAL;
  if (B == 0) return MVGEO_OK;
  if (!maps || !P || !q || !cams || !out) return MVGEO_ENULL;
  const bool soft = cfg->soft_mode != MVGEO_SOFT_NONE;
  const float* kp_tri = (soft && cfg->tri_use_soft) ? out->kp_soft : out->kp_hard;
  if (!kp_tri || !out->score || !out->X_tri) return MVGEO_ENULL;
  const int64_t n_maps = B * cfg->V * cfg->K;
  rc = mvgeo_decode(maps, cfg->dtype, n_maps, cfg->H, cfg->W, cfg->scale_x, cfg->scale_y, cfg->soft_mode, cfg->beta,
                    cfg->window_radius, cfg->apply_sigmoid, 1, 1, 0, out->idx, out->peak, out->score, out->kp_hard,
                    out->kp_soft, stream);
  if (rc) return rc;
  if (out->ticket && out->frame_loss)  // geometry tail in one launch
    return mvgeo_geometry(kp_tri, out->score, P, chain, q, B, R_view, cams, cfg->V, cfg->K, cfg->min_score,
                          cfg->tri_weighted, cfg->lambda, out->X_tri, out->tri_resid, out->tri_views, out->X_fk,
                          out->uv_fk, out->frame_loss, out->loss, out->ticket, stream);
  rc = mvgeo_triangulate(kp_tri, out->score, P, B, cfg->V, cfg->K, cfg->min_score, cfg->tri_weighted, out->X_tri,
                         out->tri_resid, out->tri_views, stream);
  if (rc) return rc;
  return mvgeo_fk_reproj_fwd(chain, q, B, R_view, cams, cfg->V, out->frame_loss ? kp_tri : nullptr, nullptr,
                             cfg->lambda, out->X_fk, out->uv_fk, out->frame_loss, out->loss, stream);
}

// ------------------------------------------------------------------------ host pipeline
namespace {
constexpr int kSlots = 2;

struct Slot {
  cudaStream_t stream = nullptr;
  void* maps = nullptr;
  float* q = nullptr;
  mvgeo_pipeline_out out = {};
};
}  // namespace

struct mvgeo_ctx {
  int device = 0;
  mvgeo_pipeline_cfg cfg = {};
  mvgeo_chain chain = {};
  int64_t chunk = 0;
  size_t frame_bytes = 0;
  float* P = nullptr;
  float* R_view = nullptr;
  mvgeo_camera* cams = nullptr;
  Slot slot[kSlots];
};

template <typename T> static cudaError_t dev_alloc(T** p, size_t n) { return cudaMalloc((void**)p, n * sizeof(T)); }

extern "C" int mvgeo_ctx_destroy(mvgeo_ctx* c) {
  if (!c) return MVGEO_OK;
  cudaSetDevice(c->device);
  for (Slot& s : c->slot) {
    if (s.stream) cudaStreamSynchronize(s.stream);
    cudaFree(s.maps); cudaFree(s.q);
    cudaFree(s.out.idx); cudaFree(s.out.peak); cudaFree(s.out.score); cudaFree(s.out.kp_hard);
    cudaFree(s.out.kp_soft); cudaFree(s.out.X_tri); cudaFree(s.out.tri_resid); cudaFree(s.out.tri_views);
    cudaFree(s.out.X_fk); cudaFree(s.out.uv_fk); cudaFree(s.out.frame_loss);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  cudaFree(c->P); cudaFree(c->R_view); cudaFree(c->cams);
  delete c;
  return MVGEO_OK;
}

extern "C" int mvgeo_ctx_create(mvgeo_ctx** ctx, int device, const mvgeo_pipeline_cfg* cfg, const mvgeo_chain* chain,
                                int64_t chunk_frames) {
  if (!ctx) return MVGEO_ENULL;
  int rc = check_cfg(cfg, chain);
  if (rc) return rc;
  if (chunk_frames < 1) return MVGEO_EINVAL;
  MVGEO_CUDA(cudaSetDevice(device));
  mvgeo_ctx* c = new (std::nothrow) mvgeo_ctx;
  if (!c) return MVGEO_ENOMEM;
  c->device = device;
  c->cfg = *cfg;
  c->chain = *chain;
  c->chunk = chunk_frames;
  const int V = cfg->V, K = cfg->K;
  c->frame_bytes = (size_t)V * K * cfg->H * cfg->W * (cfg->dtype == MVGEO_F32 ? 4 : 2);
  cudaError_t e = cudaSuccess;
  auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  ok(dev_alloc(&c->P, (size_t)V * 12));
  ok(dev_alloc(&c->R_view, (size_t)V * 9));
  ok(dev_alloc(&c->cams, (size_t)V));
  for (Slot& s : c->slot) {
    const size_t n = (size_t)chunk_frames, vk = (size_t)V * K;
    ok(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    ok(cudaMalloc(&s.maps, n * c->frame_bytes));
    ok(dev_alloc(&s.q, n * chain->n_joints));
    ok(dev_alloc(&s.out.idx, n * vk));
    ok(dev_alloc(&s.out.peak, n * vk));
    ok(dev_alloc(&s.out.score, n * vk));
    ok(dev_alloc(&s.out.kp_hard, n * vk * 2));
    ok(dev_alloc(&s.out.kp_soft, n * vk * 2));
    ok(dev_alloc(&s.out.X_tri, n * K * 3));
    ok(dev_alloc(&s.out.tri_resid, n * K));
    ok(dev_alloc(&s.out.tri_views, n * K));
    ok(dev_alloc(&s.out.X_fk, n * vk * 3));
    ok(dev_alloc(&s.out.uv_fk, n * vk * 2));
    ok(dev_alloc(&s.out.frame_loss, n));
  }
  if (e != cudaSuccess) {
    mvgeo_ctx_destroy(c);
    return e == cudaErrorMemoryAllocation ? MVGEO_ENOMEM : (int)e;
  }
  *ctx = c;
  return MVGEO_OK;
}

extern "C" int mvgeo_pipeline_host(mvgeo_ctx* c, const void* maps_host, int64_t B, const float* P_host,
                                   const float* q_host, const float* R_view_host, const mvgeo_camera* cams_host,
                                   const mvgeo_pipeline_out* oh) {
  if (!c || !oh) return MVGEO_ENULL;
  if (B < 0) return MVGEO_EINVAL;
  if (B == 0) return MVGEO_OK;
  if (!maps_host || !P_host || !q_host || !cams_host) return MVGEO_ENULL;
  MVGEO_CUDA(cudaSetDevice(c->device));
  const int V = c->cfg.V, K = c->cfg.K, J = c->chain.n_joints;
  const size_t vk = (size_t)V * K;
  cudaStream_t s0 = c->slot[0].stream;
  MVGEO_CUDA(cudaMemcpyAsync(c->P, P_host, sizeof(float) * V * 12, cudaMemcpyHostToDevice, s0));
  MVGEO_CUDA(cudaMemcpyAsync(c->cams, cams_host, sizeof(mvgeo_camera) * V, cudaMemcpyHostToDevice, s0));
  if (R_view_host)
    MVGEO_CUDA(cudaMemcpyAsync(c->R_view, R_view_host, sizeof(float) * V * 9, cudaMemcpyHostToDevice, s0));
  MVGEO_CUDA(cudaStreamSynchronize(s0));  // constants visible to both slot streams
  // the per-frame consistency term is normalised by the frames of the WHOLE call, not the chunk
  mvgeo_pipeline_cfg cfg = c->cfg;
  int n_chunks = 0;
  for (int64_t f0 = 0; f0 < B; f0 += c->chunk, ++n_chunks) {
    Slot& s = c->slot[n_chunks % kSlots];
    const int64_t n = (B - f0 < c->chunk) ? (B - f0) : c->chunk;
    const char* src = reinterpret_cast<const char*>(maps_host) + (size_t)f0 * c->frame_bytes;
    MVGEO_CUDA(cudaMemcpyAsync(s.maps, src, (size_t)n * c->frame_bytes, cudaMemcpyHostToDevice, s.stream));
    MVGEO_CUDA(cudaMemcpyAsync(s.q, q_host + f0 * J, sizeof(float) * n * J, cudaMemcpyHostToDevice, s.stream));
    mvgeo_pipeline_out o = s.out;
    o.loss = nullptr;
    o.ticket = nullptr;
    if (!oh->kp_soft && !(cfg.soft_mode != MVGEO_SOFT_NONE && cfg.tri_use_soft)) o.kp_soft = nullptr;
    cfg.lambda = c->cfg.lambda * (float)((double)n / (double)B);
    int rc = mvgeo_pipeline(&cfg, s.maps, n, c->P, &c->chain, s.q, R_view_host ? c->R_view : nullptr, c->cams, &o,
                            s.stream);
    if (rc) return rc;
#define MVGEO_D2H(field, count)                                                                        \
  if (oh->field)                                                                                       \
    MVGEO_CUDA(cudaMemcpyAsync(oh->field + (size_t)f0 * (count), s.out.field, sizeof(*oh->field) * n * (count), \
                               cudaMemcpyDeviceToHost, s.stream));
    MVGEO_D2H(idx, vk) MVGEO_D2H(peak, vk) MVGEO_D2H(score, vk) MVGEO_D2H(kp_hard, vk * 2)
    if (o.kp_soft) { MVGEO_D2H(kp_soft, vk * 2) }
    MVGEO_D2H(X_tri, (size_t)K * 3) MVGEO_D2H(tri_resid, (size_t)K) MVGEO_D2H(tri_views, (size_t)K)
    MVGEO_D2H(X_fk, vk * 3) MVGEO_D2H(uv_fk, vk * 2) MVGEO_D2H(frame_loss, (size_t)1)
#undef MVGEO_D2H
  }
  for (Slot& s : c->slot) MVGEO_CUDA(cudaStreamSynchronize(s.stream));
  if (oh->loss) {  // fixed-order host sum of the per-frame terms (B floats)
    if (!oh->frame_loss) return MVGEO_ENULL;
    double t = 0.0;
    for (int64_t i = 0; i < B; ++i) t += (double)oh->frame_loss[i];
    oh->loss[0] = (float)t;
  }
  return MVGEO_OK;
}
