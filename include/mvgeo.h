/*
 * mvgeo.h — C ABI of libmvgeo.so, the B200 (sm_100a) geometry hot path for
 * multi-view robot pose estimation: belief-map decode -> DLT triangulation ->
 * DH forward kinematics + reprojection loss (forward and backward).
 *
 * This is the drop-in boundary. The reference
 * (Najongs/2025_ICRA_Multi_View_Robot_Pose_Estimation) has no FFI layer: its
 * boundary is a set of free Python functions re-declared per script. Every entry
 * point below names the reference function(s) it replaces (file:line, raw .ipynb
 * line numbers for notebooks). A Python caller binds these with ctypes
 * (see INTEGRATION.md); no torch / pybind types appear in any signature.
 *
 * Conventions
 *   - All array pointers are DEVICE pointers unless the name ends in `_host`.
 *   - All arrays are dense, row-major, innermost dimension last.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Every call is asynchronous and stream-ordered; nothing synchronises the
 *     device. The library keeps no global mutable state and is re-entrant
 *     (the reference runs one decode loop per camera thread, DIP_REAL.py:178-185).
 *   - Return value: 0 on success, <0 for argument errors (mvgeo_status),
 *     >0 for a cudaError_t raised by the launch. Numerical failure (fewer than two
 *     valid views, all -inf maps ...) is signalled in-band with NaN / counts,
 *     never by a return code (reference behaviour: errors are swallowed to None,
 *     model/MvRoPose_FR3.py:229-231).
 */
#ifndef MVGEO_H_
#define MVGEO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVGEO_VERSION 100 /* 0.1.0 */
#define MVGEO_MAX_JOINTS 8
#define MVGEO_MAX_VIEWS 16
#define MVGEO_MAX_WINDOW_RADIUS 15

typedef enum mvgeo_status {
  MVGEO_OK = 0,
  MVGEO_EINVAL = -1,       /* bad size / enum value */
  MVGEO_ENULL = -2,        /* required pointer is NULL */
  MVGEO_EALIGN = -3,       /* pointer not aligned as documented */
  MVGEO_EUNSUPPORTED = -4, /* valid request this build cannot serve (e.g. wrong GPU arch) */
  MVGEO_ENOMEM = -5
} mvgeo_status;

typedef enum mvgeo_dtype { MVGEO_F32 = 0, MVGEO_BF16 = 1, MVGEO_F16 = 2 } mvgeo_dtype;

typedef enum mvgeo_soft_mode {
  MVGEO_SOFT_NONE = 0,   /* hard arg-max only; kp_soft (if given) = kp_hard        */
  MVGEO_SOFT_GLOBAL = 1, /* soft-arg-max over the whole map                        */
  MVGEO_SOFT_WINDOW = 2  /* soft-arg-max over the (2r+1)^2 window around the peak  */
} mvgeo_soft_mode;

typedef enum mvgeo_robot { MVGEO_ROBOT_FR3 = 0, MVGEO_ROBOT_FR5 = 1, MVGEO_ROBOT_MECA500 = 2 } mvgeo_robot;

typedef enum mvgeo_dh_convention {
  MVGEO_DH_STANDARD = 0, /* T_i = Rz(theta) Tz(d) Tx(a) Rx(alpha)  (Fr5, Meca500, MV-model) */
  MVGEO_DH_MODIFIED = 1  /* T_i = Rx(alpha) Tx(a) Rz(theta) Tz(d)  (Craig; FR3)             */
} mvgeo_dh_convention;

/* One serial chain. theta_i = (q_i + theta_offset_i) * angle_scale, so degree-valued
 * robots (Fr5, Meca500) carry angle_scale = pi/180 and offsets in degrees. cos/sin of
 * alpha are stored pre-evaluated in float64-then-rounded form, like the reference
 * evaluates them (np.cos(math.radians(alpha))). emit_base = 1 prepends the base origin
 * (K = n_joints + 1 points), 0 returns joints only (MV-model.ipynb ForwardKinematics). */
typedef struct mvgeo_chain {
  int32_t n_joints;
  int32_t convention; /* mvgeo_dh_convention */
  int32_t emit_base;
  float angle_scale;
  float a[MVGEO_MAX_JOINTS];
  float d[MVGEO_MAX_JOINTS];
  float cos_alpha[MVGEO_MAX_JOINTS];
  float sin_alpha[MVGEO_MAX_JOINTS];
  float theta_offset[MVGEO_MAX_JOINTS];
} mvgeo_chain;

/* One calibrated pinhole camera, world -> pixel:  x = K * distort(R X + t).
 * R is the row-major Rodrigues matrix of the ArUco rvec; dist = [k1 k2 p1 p2 k3]
 * (OpenCV order, dataset/4_Calib_cam_save.py:45-51). 24 floats, 96 bytes. */
typedef struct mvgeo_camera {
  float R[9];
  float t[3];
  float fx, fy, cx, cy;
  float dist[5];
  float _pad[3];
} mvgeo_camera;

/* ------------------------------------------------------------------ misc */
int mvgeo_version(void);
const char* mvgeo_error_string(int code);
/* Built-in DH tables: FR3 model/MvRoPose_FR3.py:94-101 (first 7 rows; the flange row is
 * never applied, :121), Fr5 model/Fr5_model_train.ipynb:258-265, Meca500
 * visualization/Meca500_vis.ipynb:65-70. Host call. */
int mvgeo_chain_builtin(int robot, mvgeo_chain* out_host);

/* ---------------------------------------------------------------- decode
 * Replaces extract_keypoints_from_heatmaps (model/Fr5_model_train.ipynb:4674-4705,
 * model/Franka_research3_model_train.ipynb:3634-3665, model/DREAM_model_train.ipynb:1517-1548)
 * and the inline arg-max loops (DIP_REAL.py:116-124, model/MvRoPose_FR3.py:299-304,
 * model/DREAM_Train.py:371-385,448-460), batched over n_maps = B*V*K maps.
 *
 *   maps      [n_maps, H, W] of `dtype`. Streaming (TMA) kernel: base 16-byte aligned and H*W*sizeof a multiple
 *             of 128 bytes; global soft mode also needs image rows (W*sizeof) of whole 64-byte runs (whole 128-byte
 *             runs: the faster instantiation). Any other shape / alignment takes a plain element-wise kernel —
 *             same results, never an error.
 *   idx       [n_maps] int32   flat arg-max y*W+x; first maximum wins, NaN is maximal
 *                              (torch.argmax semantics)                       (nullable)
 *   peak      [n_maps] f32     the raw maximum                                (nullable)
 *   score     [n_maps] f32     sigmoid(peak) if apply_sigmoid else peak       (nullable)
 *   kp_hard   [n_maps, 2] f32  (x*scale_x, y*scale_y), product formed in double and
 *                              rounded once, like `x * (original_w / w)`      (nullable)
 *   kp_soft   [n_maps, 2] f32  sub-pixel soft-arg-max (same scaling)          (nullable)
 *
 * Output placement: out_index(m) = (m / k_inner) * out_stride + out_offset + (m % k_inner);
 * a contiguous [n_maps] result is k_inner=1, out_stride=1, out_offset=0. The strided form
 * lets a list of V per-view tensors [B,K,H,W] (the reference's dict of views,
 * model/MvRoPose_FR3.py:625) land in one [B,V,K] result without a stack copy:
 * call once per view with k_inner=K, out_stride=V*K, out_offset=v*K.
 */
int mvgeo_decode(const void* maps, int dtype, int64_t n_maps, int H, int W,
                 double scale_x, double scale_y,
                 int soft_mode, float beta, int window_radius, int apply_sigmoid,
                 int64_t k_inner, int64_t out_stride, int64_t out_offset,
                 int32_t* idx, float* peak, float* score, float* kp_hard, float* kp_soft,
                 void* stream);

/* The same decoder over V separate per-view tensors — the reference network returns
 * dict view -> (B,K,H,W) (model/MvRoPose_FR3.py:584-627) — in ONE launch and without a stack copy:
 *   view_maps  HOST array of n_views DEVICE pointers, each [B, K, H, W] of `dtype` (read during the call only)
 * Results are dense [B, n_views, K] (kp_*: [B, n_views, K, 2]); every byte of every view is read once.
 */
int mvgeo_decode_views(const void* const* view_maps, int n_views, int dtype, int64_t B, int K, int H, int W,
                       double scale_x, double scale_y,
                       int soft_mode, float beta, int window_radius, int apply_sigmoid,
                       int32_t* idx, float* peak, float* score, float* kp_hard, float* kp_soft,
                       void* stream);

/* ---------------------------------------------------------- triangulation
 * New functionality (the reference has no triangulation).
 * Homogeneous DLT: per key-point, rows u*P[2]-P[0], v*P[2]-P[1] for every valid view,
 * X = smallest eigenvector of A^T A, de-homogenised. A view is valid when its weight
 * >= min_weight (mirrors the score filter model/Fr5_model_train.ipynb:4721-4725) and its
 * key-point is finite; `weighted` != 0 additionally scales the view's rows by its weight.
 *
 *   kp      [B, V, K, 2] f32  pixels (same pixel frame as P)
 *   w       [B, V, K]    f32  (nullable: all views valid, weight 1)
 *   P       [V, 3, 4]    f32  projection matrices K [R|t]
 *   X       [B, K, 3]    f32  NaN when fewer than 2 valid views or the point is at infinity
 *   resid   [B, K]       f32  RMS reprojection error in pixels over valid views (nullable)
 *   n_views [B, K]       i32  number of valid views                             (nullable)
 */
int mvgeo_triangulate(const float* kp, const float* w, const float* P,
                      int64_t B, int V, int K, float min_weight, int weighted,
                      float* X, float* resid, int32_t* n_views, void* stream);

/* Quaternion averaging with the same register-resident 4x4 symmetric eigen-solver: the unit
 * eigenvector of the largest eigenvalue of sum_i w_i q_i q_i^T — average_quaternion,
 * dataset/Fr5_preprocessing.py:57-65 (= dataset/Franka_research3_preprocessing.py:59-67), which the
 * ArUco extrinsics pipeline applies per marker. Sign: hemisphere of the group's first quaternion.
 *   q [G, N, 4] f32 (any fixed component order), w [G, N] f32 (nullable), out [G, 4] f32 */
int mvgeo_quat_mean(const float* q, const float* w, int64_t G, int N, float* out, void* stream);

/* --------------------------------------------------- forward kinematics
 * Replaces angle_to_joint_coordinate (FR3 model/MvRoPose_FR3.py:90-131; Fr5
 * model/Fr5_model_train.ipynb:256-288), forward_kinematics (Meca500
 * visualization/Meca500_vis.ipynb:62-82) and ForwardKinematics.forward
 * (model/MV-model.ipynb:858-874), batched.
 *
 *   chain   host pointer (copied into the launch)
 *   q       [B, n_joints] f32  joint values in the chain's native unit
 *   R_view  [V, 3, 3] f32      per-view base rotation applied on the left (nullable = identity, V=1)
 *   X       [B, V, K, 3] f32   K = n_joints + emit_base
 */
int mvgeo_fk(const mvgeo_chain* chain, const float* q, int64_t B,
             const float* R_view, int V, float* X, void* stream);

/* ------------------------------------------------------------ projection
 * Replaces joint_coordinate_to_pixel_plane (model/MvRoPose_FR3.py:133-141 and twins),
 * project_to_pixel (visualization/Fr5_vis.ipynb:111-115, Meca500_vis.ipynb:84-87) and
 * project_3d_to_2d (model/MV-model.ipynb:879-899): cv2.projectPoints with the
 * 5-coefficient Brown-Conrady model.
 *
 *   X     [B, Vx, K, 3] f32, Vx = V when x_per_view != 0 else 1 (same points for every camera)
 *   cams  [V] mvgeo_camera (device)
 *   uv    [B, V, K, 2] f32
 */
int mvgeo_project(const float* X, int x_per_view, const mvgeo_camera* cams,
                  int64_t B, int V, int K, float* uv, void* stream);

/* Inverse of the distortion: cv2.undistortPoints(kp, K, dist, P=K) (OpenCV's fixed-point iteration,
 * `iters` = 5 is its default), so key-points decoded from RAW (not cv2.undistort-ed) images can be
 * triangulated with pinhole projection matrices. The reference undistorts whole images instead
 * (cv2.undistort, model/MvRoPose_FR3.py:212, DIP_REAL.py:105); this is the per-key-point equivalent.
 *   kp, out  [B, V, K, 2] f32 pixels;  cams [V] (intrinsics and dist are used) */
int mvgeo_undistort_points(const float* kp, const mvgeo_camera* cams, int64_t B, int V, int K, int iters,
                           float* out, void* stream);

/* ------------------------------------- FK + reprojection loss, fwd / bwd
 * Replaces RobotPoseNet.forward's FK -> project chain plus the FK-consistency term of
 * robot_pose_loss (model/MV-model.ipynb:915-950), and makes it differentiable:
 *   loss = lambda * sum_{b,v,k,c} w[b,v,k] * (uv[b,v,k,c] - gt_uv[b,v,k,c])^2 / (B*V*K*2)
 * (F.mse_loss 'mean' reduction, MV-model.ipynb:949; w == NULL means all ones; points with
 * non-finite gt are skipped).
 *
 *   frame_loss [B] f32     per-frame contribution (already divided by B*V*K*2 and scaled)
 *   loss       [1] f32     deterministic fixed-order sum of frame_loss        (nullable)
 *   X_out      [B,V,K,3]   (nullable)      uv_out [B,V,K,2]   (nullable)
 * Backward:
 *   dloss      [1] f32     upstream gradient (nullable = 1)
 *   dq         [B, n_joints] f32  d loss / d q in the chain's native unit
 */
int mvgeo_fk_reproj_fwd(const mvgeo_chain* chain, const float* q, int64_t B,
                        const float* R_view, const mvgeo_camera* cams, int V,
                        const float* gt_uv, const float* w, float lambda,
                        float* X_out, float* uv_out, float* frame_loss, float* loss,
                        void* stream);
int mvgeo_fk_reproj_bwd(const mvgeo_chain* chain, const float* q, int64_t B,
                        const float* R_view, const mvgeo_camera* cams, int V,
                        const float* gt_uv, const float* w, float lambda,
                        const float* dloss, float* dq, void* stream);

/* ------------------------------------------------------------- geometry tail, one launch
 * mvgeo_triangulate + mvgeo_fk_reproj_fwd (+ the loss sum) in ONE kernel: the two stages are
 * independent (both consume the decoded key-points `kp`, which is also the consistency target, unweighted),
 * so their CTAs share a grid; the last FK CTA (atomic ticket) sums frame_loss in a fixed order.
 * Arguments as in the two stand-alone calls. `ticket` [1] i32 must be 0 on entry and is 0 again on
 * exit (allocate once, zero once); required when `loss` is given.
 */
int mvgeo_geometry(const float* kp, const float* w, const float* P, const mvgeo_chain* chain,
                   const float* q, int64_t B, const float* R_view, const mvgeo_camera* cams, int V, int K,
                   float min_weight, int weighted, float lambda,
                   float* X_tri, float* resid, int32_t* n_views, float* X_fk, float* uv_fk,
                   float* frame_loss, float* loss, int32_t* ticket, void* stream);

/* -------------------------------------------------- camera-pose refinement (PnP)
 * The step after the hot path in the reference: estimate_camera_pose
 * (model/Fr5_model_train.ipynb:4707-4753): FK points + decoded key-points with score >= threshold
 * (at least 4) -> cv2.solvePnPRansac -> plausibility gate 0.5 m < |t| < 5 m
 * (model/Franka_research3_model_train.ipynb:3696-3701) -> ArUco prior on failure (:4978-4993).
 * Levenberg-Marquardt on the reprojection error per (frame, view), started from the prior pose
 * held in cams[v].R / cams[v].t (so a refused or failed solve returns the prior).
 *   X       [B, Vx, K, 3] f32 object points, Vx = V when x_per_view != 0 else 1
 *   kp      [B, V, K, 2]  f32 image points;  w [B, V, K] f32 (nullable)
 *   rvec    [B, V, 3] f32 (cv2 Rodrigues vector), tvec [B, V, 3] f32
 *   rms     [B, V] f32 RMS reprojection error over the points used (NaN if < 4)     (nullable)
 *   status  [B, V] i32 bit 0: >= 4 valid points (solved), bit 1: converged, bit 2: plausible |t|  (nullable)
 */
int mvgeo_pnp_refine(const float* X, int x_per_view, const float* kp, const float* w,
                     const mvgeo_camera* cams, int64_t B, int V, int K, float min_weight, int max_iters,
                     float* rvec, float* tvec, float* rms, int32_t* status, void* stream);

/* Pose WITHOUT a prior: replaces cv2.solvePnPRansac(object_points, image_points, camera_matrix, dist_coeffs,
 * flags=cv2.SOLVEPNP_EPNP) inside estimate_camera_pose (model/Fr5_model_train.ipynb:4735-4741,
 * model/Franka_research3_model_train.ipynb:3690-3692). K <= 16 key-points make random sampling pointless:
 * every point triplet is a P3P hypothesis, scored on all valid points (inliers within reproj_thresh pixels —
 * OpenCV's default reprojectionError is 8 — then the inliers' squared error); the best one is refined by the
 * Levenberg-Marquardt of mvgeo_pnp_refine on its inliers. cams[v] supplies intrinsics and distortion only.
 * Refusals of the reference are kept: fewer than 4 valid points (:4728) or fewer than 4 inliers (:4743)
 * give status 0 and a NaN pose (the reference returns None).
 *   arguments as mvgeo_pnp_refine, plus
 *   inliers [B, V] i32  bit k set: key-point k is an inlier of the returned pose      (nullable)
 */
int mvgeo_pnp_solve(const float* X, int x_per_view, const float* kp, const float* w,
                    const mvgeo_camera* cams, int64_t B, int V, int K, float min_weight, float reproj_thresh,
                    int max_iters, float* rvec, float* tvec, float* rms, int32_t* status, int32_t* inliers,
                    void* stream);

/* -------------------------------------------------- GT belief-map encoder
 * Replaces create_gt_heatmap (model/MvRoPose_FR3.py:65-73, model/DREAM_Train.py:60-69):
 * exp(-((x-cx)^2+(y-cy)^2)/(2 sigma^2)), values below eps(double)*max set to 0.
 *   kp    [n_maps, 2] f32 centre in map pixels; non-finite centre -> all-zero map
 *   maps  [n_maps, H, W] of `dtype`
 */
int mvgeo_encode_gaussian(const float* kp, int64_t n_maps, int H, int W, float sigma,
                          int dtype, void* maps, void* stream);

/* ------------------------------------------ heat-map MSE loss, fwd / bwd
 * nn.MSELoss()(pred, gt) * weight (model/MvRoPose_FR3.py:846-847,975) with the target
 * generated on the fly from key-point centres (no materialised GT maps):
 *   partial [n_maps] f32 scratch, loss [1] f32,
 *   grad [n_maps,H,W] of `dtype` (nullable) = dloss * d loss / d pred, dloss [1] f32 device scalar
 *   (the upstream gradient; nullable = 1), so autograd needs no extra pass to scale the gradient.
 */
int mvgeo_heatmap_mse(const void* pred, int dtype, const float* kp, int64_t n_maps, int H, int W,
                      float sigma, float weight, const float* dloss, float* partial, float* loss, void* grad,
                      void* stream);

/* ---------------------------------- one-read training step: decode + heat-map MSE in ONE pass
 * mvgeo_decode (hard arg-max: idx, peak, score, kp_hard) and the forward of mvgeo_heatmap_mse over the SAME
 * predicted maps, which are read from HBM once (the reference reads them for nn.MSELoss,
 * model/MvRoPose_FR3.py:846-847, and again when decoding for evaluation, :299-304).
 *   kp_target [n_maps, 2] f32 target centres in MAP pixels; partial [n_maps] f32 scratch; loss [1] f32
 * Needs 16-byte aligned maps of whole 128-byte rows (H*W*sizeof % 128 == 0) whose image rows hold whole 64-byte
 * runs (W*sizeof % 64 == 0) and (W + H) * 4 bytes of shared-memory tables <= 16 KB; otherwise MVGEO_EUNSUPPORTED
 * (call the two stand-alone entry points). The gradient pass is
 * mvgeo_heatmap_mse with `grad` (it reads the prediction and writes the gradient).
 */
int mvgeo_decode_mse(const void* maps, int dtype, int64_t n_maps, int H, int W, double scale_x, double scale_y,
                     int apply_sigmoid, const float* kp_target, float sigma, float weight,
                     int32_t* idx, float* peak, float* score, float* kp_hard,
                     float* partial, float* loss, void* stream);

/* -------------------------------------------------------- fused pipeline
 * decode -> triangulate -> FK -> reprojection consistency, one stream, no host sync: two launches
 * (decode, geometry) when out->ticket is given, four otherwise.
 * Device-resident inputs. Any output pointer may be NULL except those a later stage
 * needs (kp_hard, kp_soft when soft_mode != NONE, score, X_tri).
 */
typedef struct mvgeo_pipeline_cfg {
  int32_t dtype, H, W, V, K;
  int32_t soft_mode, window_radius, apply_sigmoid;
  int32_t tri_use_soft; /* triangulate the soft key-points (else the hard ones) */
  int32_t tri_weighted;
  float beta, min_score, lambda;
  double scale_x, scale_y;
} mvgeo_pipeline_cfg;

typedef struct mvgeo_pipeline_out {
  int32_t* idx;      /* [B,V,K]   */
  float* peak;       /* [B,V,K]   */
  float* score;      /* [B,V,K]   */
  float* kp_hard;    /* [B,V,K,2] */
  float* kp_soft;    /* [B,V,K,2] */
  float* X_tri;      /* [B,K,3]   */
  float* tri_resid;  /* [B,K]     */
  int32_t* tri_views;/* [B,K]     */
  float* X_fk;       /* [B,V,K,3] */
  float* uv_fk;      /* [B,V,K,2] */
  float* frame_loss; /* [B]       */
  float* loss;       /* [1]       */
  int32_t* ticket;   /* [1] zero on entry (and again on exit): lets the geometry tail run as ONE launch;
                        NULL = separate triangulate / FK / sum launches */
} mvgeo_pipeline_out;

int mvgeo_pipeline(const mvgeo_pipeline_cfg* cfg, const void* maps, int64_t B,
                   const float* P, const mvgeo_chain* chain, const float* q,
                   const float* R_view, const mvgeo_camera* cams,
                   const mvgeo_pipeline_out* out, void* stream);
/* mvgeo_pipeline over cfg->V separate per-view tensors [B, K, H, W] (HOST array of DEVICE pointers,
 * the values of the reference's dict of views): same two launches, same results bit for bit, no
 * torch.stack copy of the maps. */
int mvgeo_pipeline_views(const mvgeo_pipeline_cfg* cfg, const void* const* view_maps, int64_t B,
                         const float* P, const mvgeo_chain* chain, const float* q,
                         const float* R_view, const mvgeo_camera* cams,
                         const mvgeo_pipeline_out* out, void* stream);

/* Host-buffer pipeline: the call a CPU-tensor caller makes (the reference moves the maps
 * to the CPU before decoding, DIP_REAL.py:113). Inputs and outputs are HOST pointers
 * (pinned for full PCIe rate); the context owns streams, events and device staging and
 * overlaps H2D of frame chunk i+1 with the kernels of chunk i. Synchronous on return. */
typedef struct mvgeo_ctx mvgeo_ctx;
int mvgeo_ctx_create(mvgeo_ctx** ctx, int device, const mvgeo_pipeline_cfg* cfg,
                     const mvgeo_chain* chain, int64_t chunk_frames);
int mvgeo_ctx_destroy(mvgeo_ctx* ctx);
int mvgeo_pipeline_host(mvgeo_ctx* ctx, const void* maps_host, int64_t B,
                        const float* P_host, const float* q_host, const float* R_view_host,
                        const mvgeo_camera* cams_host, const mvgeo_pipeline_out* out_host);

#ifdef __cplusplus
}
#endif
#endif /* MVGEO_H_ */
