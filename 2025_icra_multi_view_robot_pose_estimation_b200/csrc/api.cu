// api.cu — library-level entry points: version / error strings, the fused device pipeline
// (decode -> triangulate -> FK + reprojection consistency on one stream, no host sync) and the
// host-buffer pipeline (pinned host memory in, chunked H2D overlapped with compute, D2H out).
#include <cstring>
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

// maps: one dense [B,V,K,H,W] array (view_maps == NULL) or V per-view [B,K,H,W] arrays.
static int pipeline_impl(const mvgeo_pipeline_cfg* cfg, const void* maps, const void* const* view_maps, int64_t B,
                         const float* P, const mvgeo_chain* chain, const float* q, const float* R_view,
                         const mvgeo_camera* cams, const mvgeo_pipeline_out* out, void* stream) {
  int rc = check_cfg(cfg, chain);
  if (rc) return rc;
  if (B < 0) return MVGEO_EINVAL;
  if (B == 0) return MVGEO_OK;
  if ((!maps && !view_maps) || !P || !q || !cams || !out) return MVGEO_ENULL;
  const bool soft = cfg->soft_mode != MVGEO_SOFT_NONE;
  const float* kp_tri = (soft && cfg->tri_use_soft) ? out->kp_soft : out->kp_hard;
  if (!kp_tri || !out->score || !out->X_tri) return MVGEO_ENULL;
  const int64_t n_maps = B * cfg->V * cfg->K;
  if (view_maps)
    rc = mvgeo_decode_views(view_maps, cfg->V, cfg->dtype, B, cfg->K, cfg->H, cfg->W, cfg->scale_x, cfg->scale_y,
                            cfg->soft_mode, cfg->beta, cfg->window_radius, cfg->apply_sigmoid, out->idx, out->peak,
                            out->score, out->kp_hard, out->kp_soft, stream);
  else
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

extern "C" int mvgeo_pipeline(const mvgeo_pipeline_cfg* cfg, const void* maps, int64_t B, const float* P,
                              const mvgeo_chain* chain, const float* q, const float* R_view,
                              const mvgeo_camera* cams, const mvgeo_pipeline_out* out, void* stream) {
  if (B > 0 && !maps) return MVGEO_ENULL;
  return pipeline_impl(cfg, maps, nullptr, B, P, chain, q, R_view, cams, out, stream);
}

extern "C" int mvgeo_pipeline_views(const mvgeo_pipeline_cfg* cfg, const void* const* view_maps, int64_t B,
                                    const float* P, const mvgeo_chain* chain, const float* q, const float* R_view,
                                    const mvgeo_camera* cams, const mvgeo_pipeline_out* out, void* stream) {
  if (B > 0 && !view_maps) return MVGEO_ENULL;
  return pipeline_impl(cfg, nullptr, view_maps, B, P, chain, q, R_view, cams, out, stream);
}

// ------------------------------------------------------------------------ host pipeline
namespace {
constexpr int kSlots = 2;

// Restores the caller's current device on every exit path (a multi-GPU PyTorch process must not
// find its thread switched to another GPU after a library call).
struct DeviceGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Per-frame result fields, in record order. A chunk's results live back to back in ONE device
// buffer (field-major: all idx of the chunk, then all peak, ...), so a chunk costs ONE D2H copy.
enum Field { F_IDX, F_PEAK, F_SCORE, F_KP_HARD, F_KP_SOFT, F_X_TRI, F_RESID, F_VIEWS, F_X_FK, F_UV_FK, F_FRAME_LOSS, F_COUNT };

struct Slot {
  cudaStream_t stream = nullptr;
  void* maps = nullptr;
  float* q = nullptr;
  char* rec = nullptr;  // chunk * rec_words * 4 bytes
};
}  // namespace

struct mvgeo_ctx {
  int device = 0;
  mvgeo_pipeline_cfg cfg = {};
  mvgeo_chain chain = {};
  int64_t chunk = 0;
  size_t frame_bytes = 0;
  size_t words[F_COUNT] = {};  // 32-bit words per frame of every field
  size_t rec_words = 0;        // their sum
  float* P = nullptr;
  float* R_view = nullptr;
  mvgeo_camera* cams = nullptr;
  char* host_rec = nullptr;  // pinned staging for the records of one call
  size_t host_cap = 0;       // frames
  Slot slot[kSlots];
};

template <typename T> static cudaError_t dev_alloc(T** p, size_t n) { return cudaMalloc((void**)p, n * sizeof(T)); }

extern "C" int mvgeo_ctx_destroy(mvgeo_ctx* c) {
  if (!c) return MVGEO_OK;
  DeviceGuard guard(c->device);
  for (Slot& s : c->slot) {
    if (s.stream) cudaStreamSynchronize(s.stream);
    cudaFree(s.maps); cudaFree(s.q); cudaFree(s.rec);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  cudaFree(c->P); cudaFree(c->R_view); cudaFree(c->cams);
  if (c->host_rec) cudaFreeHost(c->host_rec);
  delete c;
  return MVGEO_OK;
}

extern "C" int mvgeo_ctx_create(mvgeo_ctx** ctx, int device, const mvgeo_pipeline_cfg* cfg, const mvgeo_chain* chain,
                                int64_t chunk_frames) {
  if (!ctx) return MVGEO_ENULL;
  int rc = check_cfg(cfg, chain);
  if (rc) return rc;
  if (chunk_frames < 1) return MVGEO_EINVAL;
  DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return (int)guard.err;
  mvgeo_ctx* c = new (std::nothrow) mvgeo_ctx;
  if (!c) return MVGEO_ENOMEM;
  c->device = device;
  c->cfg = *cfg;
  c->chain = *chain;
  c->chunk = chunk_frames;
  const size_t V = cfg->V, K = cfg->K, vk = V * K;
  c->frame_bytes = vk * cfg->H * cfg->W * (cfg->dtype == MVGEO_F32 ? 4 : 2);
  const size_t words[F_COUNT] = {vk, vk, vk, vk * 2, vk * 2, K * 3, K, K, vk * 3, vk * 2, 1};
  for (int f = 0; f < F_COUNT; ++f) {
    c->words[f] = words[f];
    c->rec_words += words[f];
  }
  cudaError_t e = cudaSuccess;
  auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  ok(dev_alloc(&c->P, V * 12));
  ok(dev_alloc(&c->R_view, V * 9));
  ok(dev_alloc(&c->cams, V));
  for (Slot& s : c->slot) {
    const size_t n = (size_t)chunk_frames;
    ok(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    ok(cudaMalloc(&s.maps, n * c->frame_bytes));
    ok(dev_alloc(&s.q, n * chain->n_joints));
    ok(dev_alloc(&s.rec, n * c->rec_words * 4));
  }
  if (e != cudaSuccess) {
    mvgeo_ctx_destroy(c);
    return e == cudaErrorMemoryAllocation ? MVGEO_ENOMEM : (int)e;
  }
  *ctx = c;
  return MVGEO_OK;
}

// Field f of a chunk of n frames inside a record buffer (device or host staging).
static inline char* field_ptr(const mvgeo_ctx* c, char* rec, int64_t n, int f) {
  size_t off = 0;
  for (int i = 0; i < f; ++i) off += c->words[i];
  return rec + off * (size_t)n * 4;
}

extern "C" int mvgeo_pipeline_host(mvgeo_ctx* c, const void* maps_host, int64_t B, const float* P_host,
                                   const float* q_host, const float* R_view_host, const mvgeo_camera* cams_host,
                                   const mvgeo_pipeline_out* oh) {
  if (!c || !oh) return MVGEO_ENULL;
  if (B < 0) return MVGEO_EINVAL;
  if (B == 0) return MVGEO_OK;
  if (!maps_host || !P_host || !q_host || !cams_host) return MVGEO_ENULL;
  if (oh->loss && !oh->frame_loss) return MVGEO_ENULL;
  DeviceGuard guard(c->device);
  if (guard.err != cudaSuccess) return (int)guard.err;
  const int V = c->cfg.V, J = c->chain.n_joints;
  const size_t rec_bytes = c->rec_words * 4;
  if ((size_t)B > c->host_cap) {  // staging for the whole call's records (1-2 KB per frame), grown on demand
    if (c->host_rec) cudaFreeHost(c->host_rec);
    c->host_rec = nullptr;
    c->host_cap = 0;
    MVGEO_CUDA(cudaHostAlloc((void**)&c->host_rec, (size_t)B * rec_bytes, cudaHostAllocDefault));
    c->host_cap = (size_t)B;
  }
  cudaStream_t s0 = c->slot[0].stream;
  // From here on copies into caller memory may be in flight: every exit drains both streams first.
  auto drain = [&](int rc) {
    for (Slot& s : c->slot) cudaStreamSynchronize(s.stream);
    return rc;
  };
#define MVGEO_TRY(call)                                  \
  do {                                                   \
    cudaError_t e__ = (call);                            \
    if (e__ != cudaSuccess) return drain((int)e__);      \
  } while (0)
  MVGEO_TRY(cudaMemcpyAsync(c->P, P_host, sizeof(float) * V * 12, cudaMemcpyHostToDevice, s0));
  MVGEO_TRY(cudaMemcpyAsync(c->cams, cams_host, sizeof(mvgeo_camera) * V, cudaMemcpyHostToDevice, s0));
  if (R_view_host)
    MVGEO_TRY(cudaMemcpyAsync(c->R_view, R_view_host, sizeof(float) * V * 9, cudaMemcpyHostToDevice, s0));
  MVGEO_TRY(cudaStreamSynchronize(s0));  // constants visible to both slot streams
  // the per-frame consistency term is normalised by the frames of the WHOLE call, not the chunk
  mvgeo_pipeline_cfg cfg = c->cfg;
  const bool want_soft = oh->kp_soft || (cfg.soft_mode != MVGEO_SOFT_NONE && cfg.tri_use_soft);
  int n_chunks = 0;
  for (int64_t f0 = 0; f0 < B; f0 += c->chunk, ++n_chunks) {
    Slot& s = c->slot[n_chunks % kSlots];
    const int64_t n = (B - f0 < c->chunk) ? (B - f0) : c->chunk;
    const char* src = reinterpret_cast<const char*>(maps_host) + (size_t)f0 * c->frame_bytes;
    MVGEO_TRY(cudaMemcpyAsync(s.maps, src, (size_t)n * c->frame_bytes, cudaMemcpyHostToDevice, s.stream));
    MVGEO_TRY(cudaMemcpyAsync(s.q, q_host + f0 * J, sizeof(float) * n * J, cudaMemcpyHostToDevice, s.stream));
    mvgeo_pipeline_out o = {};
    o.idx = (int32_t*)field_ptr(c, s.rec, n, F_IDX);
    o.peak = (float*)field_ptr(c, s.rec, n, F_PEAK);
    o.score = (float*)field_ptr(c, s.rec, n, F_SCORE);
    o.kp_hard = (float*)field_ptr(c, s.rec, n, F_KP_HARD);
    o.kp_soft = want_soft ? (float*)field_ptr(c, s.rec, n, F_KP_SOFT) : nullptr;
    o.X_tri = (float*)field_ptr(c, s.rec, n, F_X_TRI);
    o.tri_resid = (float*)field_ptr(c, s.rec, n, F_RESID);
    o.tri_views = (int32_t*)field_ptr(c, s.rec, n, F_VIEWS);
    o.X_fk = (float*)field_ptr(c, s.rec, n, F_X_FK);
    o.uv_fk = (float*)field_ptr(c, s.rec, n, F_UV_FK);
    o.frame_loss = (float*)field_ptr(c, s.rec, n, F_FRAME_LOSS);
    cfg.lambda = c->cfg.lambda * (float)((double)n / (double)B);
    int rc = mvgeo_pipeline(&cfg, s.maps, n, c->P, &c->chain, s.q, R_view_host ? c->R_view : nullptr, c->cams, &o,
                            s.stream);
    if (rc) return drain(rc);
    // ONE device-to-host copy per chunk: the whole record block
    MVGEO_TRY(cudaMemcpyAsync(c->host_rec + (size_t)f0 * rec_bytes, s.rec, (size_t)n * rec_bytes, cudaMemcpyDeviceToHost,
                              s.stream));
  }
  for (Slot& s : c->slot) MVGEO_TRY(cudaStreamSynchronize(s.stream));
#undef MVGEO_TRY
  // scatter the records into the caller's per-field arrays (host memcpy, ~1-2 KB per frame)
  void* dst[F_COUNT] = {oh->idx, oh->peak, oh->score, oh->kp_hard, want_soft ? oh->kp_soft : nullptr, oh->X_tri,
                        oh->tri_resid, oh->tri_views, oh->X_fk, oh->uv_fk, oh->frame_loss};
  for (int64_t f0 = 0; f0 < B; f0 += c->chunk) {
    const int64_t n = (B - f0 < c->chunk) ? (B - f0) : c->chunk;
    char* rec = c->host_rec + (size_t)f0 * rec_bytes;
    for (int f = 0; f < F_COUNT; ++f)
      if (dst[f])
        memcpy(reinterpret_cast<char*>(dst[f]) + (size_t)f0 * c->words[f] * 4, field_ptr(c, rec, n, f),
               (size_t)n * c->words[f] * 4);
  }
  if (oh->loss) {  // fixed-order host sum of the per-frame terms (B floats)
    double t = 0.0;
    for (int64_t i = 0; i < B; ++i) t += (double)oh->frame_loss[i];
    oh->loss[0] = (float)t;
  }
  return MVGEO_OK;
}
