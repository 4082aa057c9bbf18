/*
 * bdpose.h — C ABI of libbdpose.so: the B200 (sm_100a) bin-and-delta pose hot path.
 *
 * The reference (JHUVisionLab/multi-modal-regression) has no FFI of its own: its boundary is a set
 * of Python call sites.  Each entry point below names the reference call site (file:line under the
 * reference checkout) whose arithmetic it replaces.  The Python mirror of the reference API
 * (multi-modal-regression_b200/{axisAngle,quaternion,binDeltaLosses,binDeltaGenerators,
 * binDeltaModels,poseModels}.py) binds these symbols with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in _host;
 *   - the caller owns every buffer; no entry point allocates device memory, synchronises the device
 *     or the stream, or touches a stream other than the one passed in (`stream` is a cudaStream_t
 *     passed as void*; NULL = legacy default stream);
 *   - row-major, densely packed unless a leading dimension is passed;
 *   - return value: BDP_OK (0) or a negative BDP_ERR_* code; bdp_last_error() returns a
 *     thread-local human-readable message for the last failure on the calling thread;
 *   - thread-safe for concurrent calls on distinct streams and distinct buffers.
 */
#ifndef BDPOSE_H_
#define BDPOSE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BDP_OK 0
#define BDP_ERR_ARG (-1)         /* bad argument (shape, alignment, enum) */
#define BDP_ERR_CUDA (-2)        /* a CUDA runtime call / launch failed */
#define BDP_ERR_UNSUPPORTED (-3) /* valid request this build cannot serve */

#define BDP_ABI_VERSION 1

int bdp_abi_version(void);
const char* bdp_last_error(void);
/* SM count of the current device (grid sizing on the Python side). */
int bdp_sm_count(void);

/* ------------------------------------------------------------------------------------------------
 * (b) fused bin-delta loss:  cross-entropy over the K pose bins  +  argmax key gather  +  pose
 *     composition  +  pose loss, forward AND the hand-derived backward in one pass over the logits.
 *
 * Replaces, per pose_mode:
 *   BDP_POSE_NONE        nn.CrossEntropyLoss only                    learnGeodesicBDModel.py:69,178
 *   BDP_POSE_MSE         SimpleLoss / loss_m0  (use_keys=0)           binDeltaLosses.py:16-28, 243-256
 *                        GeodesicLoss with my_loss=None (use_keys=1)  binDeltaLosses.py:31-50
 *   BDP_POSE_GEODESIC_AA GeodesicLoss(my_loss=axisAngle.geodesic_loss) binDeltaLosses.py:44-50 +
 *                        axisAngle.py:110-120; script form learnGeodesicBDModel.py:175-180
 *   BDP_POSE_GEODESIC_Q  GeodesicLossQ(my_loss=quaternion.geodesic_loss) binDeltaLosses.py:53-72 +
 *                        quaternion.py:156-163
 *   BDP_POSE_RIEMANNIAN  RiemannianLoss (Rodrigues delta composed on the key rotation, trace-acos)
 *                        binDeltaLosses.py:211-239; learnRiemannianBDModel.py:69-95
 * With logits == NULL the cross-entropy part is skipped and the call is the stand-alone pose loss
 * (axisAngle.geodesic_loss axisAngle.py:103-120, quaternion.geodesic_loss quaternion.py:149-163,
 * RiemannianLoss.my_loss binDeltaLosses.py:221-225 via BDP_POSE_ROTMAT).
 * ---------------------------------------------------------------------------------------------- */
enum {
  BDP_POSE_NONE = 0,
  BDP_POSE_MSE = 1,
  BDP_POSE_GEODESIC_AA = 2,
  BDP_POSE_GEODESIC_Q = 3,
  BDP_POSE_RIEMANNIAN = 4,
  BDP_POSE_ROTMAT = 5 /* pred is a [B,9] rotation matrix: acos(clamp((tr(PᵀT)-1)/2)); no keys */
};

/* bytes of zero-initialised scratch the loss call needs for B rows (block partials + ticket).
 * The call leaves the ticket zeroed again, so one allocation can be reused forever. */
int64_t bdp_bd_loss_workspace_bytes(int64_t B);

/*
 * logits     [B, ld_logits] fp32, K valid columns (NULL: no cross-entropy)
 * bin_true   [B] int64 target bin (required iff logits != NULL)
 * pred       [B, ndim] fp32 predicted delta / pose   (ndim 3 | 4 | 9 by pose_mode)
 * keys       use_keys: [K, ndim] fp32 key poses (axis-angle / quaternion), or [K, 9] rotation
 *            matrices for BDP_POSE_RIEMANNIAN; the key is chosen by argmax_k logits (lowest index
 *            on ties) and carries no gradient to the logits (learnGeodesicBDModel.py:175-176)
 * target     [B, tdim] fp32 ground truth: ndim columns, 9 for RIEMANNIAN / ROTMAT
 * out_loss   [2] fp32: {mean CE, mean pose loss}   (always written; unused slot = 0)
 * row_ce, row_pose  [B] fp32 per-sample values (either may be NULL)
 * grad_logits [B, ld_logits] fp32 = grad_scale * (softmax - onehot)              (NULL: skip)
 * grad_pred   [B, ndim] fp32 = grad_scale * d(row pose loss)/d pred              (NULL: skip)
 * grad_scale  1/B for the mean-reduced losses (the reference default), 1 for reduce=False
 *             (learnProbabilisticBDModel.py:70); <= 0 selects 1/B
 * argmax_out  [B] int64 chosen bin (NULL: skip)
 */
int bdp_bd_loss_fwd_bwd(const float* logits, int64_t B, int64_t K, int64_t ld_logits,
                        const int64_t* bin_true, const float* pred, int ndim, const float* keys,
                        int use_keys, const float* target, int pose_mode, float* out_loss,
                        float* row_ce, float* row_pose, float* grad_logits, float* grad_pred,
                        float grad_scale, int64_t* argmax_out, void* workspace,
                        int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (d) evaluation: batched geodesic error in degrees + Acc@30deg + (per-class) median.
 *   axisAngle.get_error / get_error2   axisAngle.py:45-66, 70-95
 *   quaternion.get_error / get_error2  quaternion.py:33-51, 55-76
 * ---------------------------------------------------------------------------------------------- */
enum { BDP_REPR_AXIS_ANGLE = 0, BDP_REPR_QUATERNION = 1 };
enum { BDP_F32 = 0, BDP_F64 = 1 };

/* y_gt, y_hat [N, 3|4] of `dtype`; err_deg [N] fp64 (degrees, as the reference returns) */
int bdp_geodesic_error_deg(const void* y_gt, const void* y_hat, int dtype, int repr, int64_t N,
                           double* err_deg, void* stream);

int64_t bdp_error_stats_workspace_bytes(int64_t N, int num_classes);
/*
 * err_deg [N] fp64 (>= 0).  labels [N] int64 class ids in [0,num_classes) or NULL (one class).
 * Outputs (device): median [num_classes] fp64 with np.median semantics (mean of the two middle
 * order statistics for even counts, NaN for an empty class — axisAngle.py:92);
 * count [num_classes] int64; below30 [1] int64 = #(err < 30); max_err [1] fp64.
 */
int bdp_error_stats(const double* err_deg, const int64_t* labels, int64_t N, int num_classes,
                    double* median, int64_t* count, int64_t* below30, double* max_err,
                    void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Expected pose loss of the soft-bin losses (binDeltaLosses.py:119-123, 141-146, 162-167, 176-180,
 * 311-316, 327-333; the reference loops over the K bins in python):
 *   rows[b] = sum_k softmax(logits[b])_k * L(target[b], keys[k] + delta[b(,k)])
 * L = axis-angle geodesic (pose_mode BDP_POSE_GEODESIC_AA, ndim 3) or the quaternion geodesic with
 * the prediction in the un-normalised `ytrue` slot (BDP_POSE_GEODESIC_Q, ndim 4), as the reference
 * calls my_loss(ydata, pose).  delta [B, ndim] (per_bin = 0) or [B, K, ndim] (per_bin = 1).
 * grad_logits [B, K] = d rows[b] / d logits, grad_delta (shape of delta) = d rows[b] / d delta.
 */
int bdp_expected_pose_loss(const float* logits, int64_t B, int64_t K, int64_t ld_logits,
                           const float* delta, int per_bin, int ndim, const float* keys,
                           const float* target, int pose_mode, float* rows, float* grad_logits,
                           float* grad_delta, void* stream);

/*
 * Test-time pose composition of the scripts' testing() loops, per prediction:
 * bin = argmax_k score[k] (first maximum, np.argmax), then
 *   BDP_COMPOSE_ADD            out = dict[bin] + residual        learnGeodesicBDModel.py:217-219
 *   BDP_COMPOSE_ADD_NORMALIZE  y = dict[bin] + residual, out = y / max(||y||, 1e-10)
 *                                                                learnGeodesicBDModel_quaternion.py:217-218
 *   BDP_COMPOSE_RIEMANNIAN     out = get_y(R_key[bin] . get_R(residual))   (dict = [K,9] key rotations)
 *                                                                learnRiemannianBDModel.py:247
 * score [N, ld_score] fp32 (K valid columns), residual [N, ndim] fp32, dict [K, ndim] (or [K,9]) fp64
 * -> out [N, ndim] fp64 (numpy's float64 + float32 result type), bin_out [N] int64 (NULL: skip).
 */
enum { BDP_COMPOSE_ADD = 0, BDP_COMPOSE_ADD_NORMALIZE = 1, BDP_COMPOSE_RIEMANNIAN = 2 };
int bdp_compose_prediction(const float* score, int64_t N, int K, int64_t ld_score,
                           const float* residual, int ndim, const double* dict, int mode, double* out,
                           int64_t* bin_out, void* stream);

/* out[0] = min_{i != j} ||keys_i - keys_j||^2, keys [K, d] fp64 — helperFunctions.get_gamma
 * (helperFunctions.py:51-58) is 1 / (2 * that). */
int bdp_min_key_gap(const double* keys, int K, int d, double* out, void* stream);

/*
 * One optimizer step of helperFunctions.mySGD (helperFunctions.py:62-120; the snapshot-ensemble SGD
 * of the evaluate* scripts) over MANY tensors in one launch.  table_dev: device array of n_tensors
 * rows; per tensor  d = g (+ weight_decay * p);  with momentum: buf = d on the first step, else
 * momentum * buf + (1 - dampening) * d;  d = d + momentum * buf (nesterov) or buf;  p -= step_size * d.
 * step_size is the cyclical learning rate of the row's own step counter (computed by the caller).
 * The gradient tensors are read only (the reference adds the weight decay into .grad in place).
 */
typedef struct bdp_sgd_tensor {
  float* p;        /* parameter, updated in place */
  const float* g;  /* gradient */
  float* buf;      /* momentum buffer (may be NULL when momentum == 0) */
  int64_t n;
  float step_size;
  int32_t first;   /* 1: the momentum buffer is created by this step */
} bdp_sgd_tensor;
int bdp_sgd_step(const bdp_sgd_tensor* table_dev, int n_tensors, int64_t max_numel, float weight_decay,
                 float momentum, float dampening, int nesterov, void* stream);

/*
 * Column statistics of the k-means preprocessing (scikit-learn KMeans.fit: X -= X.mean(0);
 * tol = mean(var(X)) * tol — sklearn/cluster/_kmeans.py), in fixed point so that they do not depend on
 * the order of summation or on how the rows are split over GPUs.  x [N, d] fp64; `out` is ADDED to /
 * max-ed into (zero it first; all-reduce it across ranks afterwards):
 *   mode 0  out[0] = max |x|                         (uint64 bit pattern of the double, atomic max)
 *   mode 1  out[0..d) += sum floor(x * scale), out[d..2d) += sum trunc(frac(x * scale) * 2^32)  (int64)
 *   mode 2  y = x - mean[col]; out[0] = max (x - mean)^2, out[1] = max |x - mean|   (bit patterns)
 *   mode 3  out[0..d) += sum llrint(x^2 * scale)     (x already centred; int64)
 */
int bdp_fit_stats(const double* x, int64_t N, int d, int mode, const double* mean, double scale,
                  double* y, void* out, void* stream);

/* buf[i] *= *scale_dev, i < n; returns at once (no memory traffic) when *scale_dev == 1 — the upstream
 * scalar of a loss whose gradient the forward launch already wrote (buf 16-byte aligned). */
int bdp_scale_inplace(float* buf, int64_t n, const float* scale_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (c) nearest-dictionary-key assignment + residual, and the k-means Lloyd step.
 *   kmeans.predict + residual        binDeltaGenerators.py:27-30, 78-82 (quaternion keys: 67)
 *   Riemannian residual + ydata_rot  binDeltaGenerators.py:131-137
 *   argmax |K·q| assignment          learnObjectnetModel.py:108-109
 *   KMeans(K).fit                    learnKmeansDictionary.py:41-42  (scikit-learn Lloyd E+M step:
 *                                    sklearn/cluster/_k_means_lloyd.pyx, _k_means_common.pyx)
 * ---------------------------------------------------------------------------------------------- */

/*
 * x        [N, d] rotations, fp32 or fp64 (x_dtype), d = 3 or 4
 * centers  [K, d] fp64 dictionary (cluster_centers_)
 * labels32 [N] int32 and/or labels64 [N] int64 nearest key, argmin_k ||x-c_k||^2 evaluated
 *          faithfully to fp64 (fp32 screening + fp64 re-check of near ties), lowest index on ties
 * residual [N, d] fp32 = float(x - c[label])                  (NULL: skip)
 * min_sqdist [N] fp64 = ||x - c[label]||^2                    (NULL: skip)
 */
int bdp_assign_nearest(const void* x, int x_dtype, int64_t N, int d, const double* centers, int K,
                       int32_t* labels32, int64_t* labels64, float* residual, double* min_sqdist,
                       void* stream);

/* q [N,4] unit quaternions (fp32 or fp64 by q_dtype), keys [K,4] fp64: bin = argmax_k |<key_k, q>|
 * evaluated in fp64 (lowest index on ties), residual = float(q - keys[bin]).
 * learnObjectnetModel.py:108-109 (fixed 16-key dictionary 60-66) */
int bdp_assign_quatdot(const void* q, int q_dtype, int64_t N, const double* keys, int K,
                       int64_t* bin, float* residual, void* stream);

/* Soft assignment: p[n,k] = exp(-gamma ||x_n - c_k||^2) / sum_j exp(-gamma ||x_n - c_j||^2)
 * (exp / normalise without max subtraction, as numpy does it in the reference), residual[n] =
 * x_n - sum_k p[n,k] c_k; evaluated in fp64, stored as fp32.
 * x [N,d] fp32|fp64 (x_dtype), d = 3|4, centers [K,d] fp64 -> p [N,K] fp32 (NULL: skip),
 * residual [N,d] fp32 (NULL: skip).
 * binDeltaGenerators.py:104-108 (XPBDGeneratorQ, gamma = 10); dataGenerators.py:155-157, 166;
 * ablationFunctions.py:146-150 */
int bdp_assign_soft(const void* x, int x_dtype, int64_t N, int d, const double* centers, int K,
                    double gamma, float* p, float* residual, void* stream);

/* Riemannian residual of RBDGenerator: rot[n] = exp([x_n]x) (axisAngle.get_R), residual[n] =
 * log(key_rot[bin_n]^T rot[n]) (axisAngle.get_y, zero vector when the axis norm <= 1e-6), all
 * evaluated in fp64 and stored as fp32 (the reference's `.float()`).
 * x [N,3] fp32|fp64 (x_dtype), key_rot [K,9] fp64, bin [N] int64 -> rot [N,9] fp32 (NULL: skip),
 * residual [N,3] fp32 (NULL: skip; otherwise key_rot and bin are required).
 * binDeltaGenerators.py:120, 131, 137; dataGenerators.py:173-178 */
int bdp_riemannian_residual(const void* x, int x_dtype, int64_t N, const double* key_rot, int K,
                            const int64_t* bin, float* rot, float* residual, void* stream);

/* axis-angle [N,3] fp64 -> rotation matrices [N,9] fp64 (axisAngle.get_R, axisAngle.py:33-41) and
 * unit quaternions [N,4] fp64 (quaternion.convert_dictionary, quaternion.py:79-92). Either output
 * may be NULL. */
int bdp_convert_axis_angle(const double* aa, int64_t N, double* rotmat, double* quat, void* stream);

/* Pose targets from Euler angles: euler_deg [N,3] fp64 = (azimuth, elevation, camera tilt) in degrees
 * -> R = Rz(ct) Rx(el) Rz(az) (helperFunctions.rotation_matrix, helperFunctions.py:37-48) -> axis-angle
 * [N,3] fp64 (axisAngle.get_y, axisAngle.py:19-29) and / or unit quaternion [N,4] fp64
 * (quaternion.get_y, quaternion.py:18-29).  Either output may be NULL.  The reference runs this per
 * image in python (learnKmeansDictionary.py:31-37, dataGenerators.py:55-69); rendered images pass -ct. */
int bdp_euler_to_pose(const double* euler_deg, int64_t N, double* aa, double* quat, void* stream);

/*
 * One Lloyd iteration, E-step + M-step accumulation (sklearn lloyd_iter_chunked_dense):
 *   labels[n] = argmin_k ||x_n - c_k||^2 (fp64-faithful, lowest index on ties); the per-cluster
 *   coordinate sums are accumulated EXACTLY in two-limb int64 fixed point (order-independent, so
 *   the result is bit-identical for any grid, any launch and any sharding over GPUs), together
 *   with member counts, the number of labels that changed, and the inertia of this assignment.
 *
 * x        [N, d] fp64 (d = 3 or 4), N < 2^30
 * centers  [K, d] fp64
 * labels   [N] int32 in/out (previous labels in, new labels out; -1 initially)
 * acc      [K, 2*d + 1] int64, ADDED to (zero it before the first shard; all-reduce(SUM) it
 *          across ranks afterwards): per cluster {hi_0, lo_0, ..., hi_{d-1}, lo_{d-1}, count}
 *          where sum_x = hi * 2^-fix_hi_bits + lo * 2^-(fix_hi_bits+32)
 * fix_hi_bits  fixed-point scale: requires |x| < 2^(31 - fix_hi_bits)
 * stats    [2] int64 ADDED to: {changed labels, 0}; inertia [1] fp64 ADDED to (NULL: skip)
 * update   0 = E-step only (labels, changed, inertia; acc untouched)
 */
int bdp_kmeans_lloyd_step(const double* x, int64_t N, int d, const double* centers, int K,
                          int32_t* labels, int64_t* acc, int fix_hi_bits, int64_t* stats,
                          double* inertia, int update, void* stream);

/*
 * Key grid: candidate pruning for the nearest-key query (same labels as the brute-force entry points,
 * ~5 candidate keys per rotation instead of K).  A uniform grid over the dictionary's bounding box
 * stores, per cell, the ascending list of keys that can be nearest to some point of the cell; points
 * outside the grid and cells with overflowing lists take the brute-force path inside the kernel.
 *   bdp_keygrid_bytes(K, d)       device bytes the caller allocates for the grid (-1: unsupported;
 *                                 supported: d = 3|4, 1 <= K <= 4096)
 *   bdp_keygrid_build(...)        (re)builds the grid for `centers` [K, d] fp64 — two small
 *                                 launches, no host synchronisation; rebuild whenever centers change
 *   bdp_assign_nearest_grid       bdp_assign_nearest with a prebuilt grid (16-byte aligned)
 *   bdp_kmeans_lloyd_step_grid    bdp_kmeans_lloyd_step with a prebuilt grid
 * Replaces the same reference call sites as the brute-force forms (binDeltaGenerators.py:27-30,
 * learnKmeansDictionary.py:41-42).
 */
/* Diagnostics of a query batch against a built grid: stats [5] int64 (zeroed by the caller) receive
 * {points, points outside the grid, points in overflowed cells, sum of candidate-list lengths over
 * the other points, longest list seen}. */
int bdp_keygrid_stats(const void* x, int x_dtype, int64_t N, int d, int K, const void* grid,
                      int64_t grid_bytes, int64_t* stats, void* stream);
int64_t bdp_keygrid_bytes(int K, int d);
int bdp_keygrid_build(const double* centers, int K, int d, void* grid, int64_t grid_bytes,
                      void* stream);
/* Fixed-geometry grid for the k-means loop (learnKmeansDictionary.py:41-42: one data set, many
 * dictionaries).  bdp_keygrid_prepare lays the grid over the caller's box [box_lo, box_hi] ([d] fp64,
 * device; the bounding box of the rows) instead of the dictionary's and marks every cell "scan the
 * dictionary"; bdp_keygrid_occupancy sets occ[c] = 1 (occ [bdp_keygrid_coarse_cells(K, d)] int32,
 * zeroed by the caller) for every coarse cell c that holds a row of x [N, d] fp64 — with the query
 * kernel's own point -> cell arithmetic.  bdp_kmeans_run then rebuilds only the listed cells. */
int64_t bdp_keygrid_coarse_cells(int K, int d);
int bdp_keygrid_prepare(const double* box_lo, const double* box_hi, int K, int d, void* grid,
                        int64_t grid_bytes, void* stream);
int bdp_keygrid_occupancy(const double* x, int64_t N, int d, int K, const void* grid,
                          int64_t grid_bytes, int32_t* occ, void* stream);
/* Rows of a fit in cell order of a PREPARED grid (bdp_keygrid_prepare): perm [N] int32 receives the
 * permutation (row i of x_sorted [N, d] is row perm[i] of x), occ (NULL: skip) the occupied coarse
 * cells as bdp_keygrid_occupancy marks them.  A warp of the E+M kernel then works on one or two
 * cells (one cell-record line, the same candidate keys in every lane).  Labels and cluster sums do
 * not depend on the order of the rows; bdp_scatter_i32 (dst[perm[i]] = src[i]) takes the labels back
 * to the caller's order.  workspace: bdp_cellsort_workspace_bytes(N, K, d) device bytes. */
int64_t bdp_cellsort_workspace_bytes(int64_t N, int K, int d);
int bdp_cellsort(const double* x, int64_t N, int d, int K, void* grid, int64_t grid_bytes,
                 int32_t* occ, void* workspace, int64_t workspace_bytes, int32_t* perm,
                 double* x_sorted, void* stream);
int bdp_scatter_i32(const int32_t* src, const int32_t* perm, int64_t N, int32_t* dst, void* stream);
int bdp_assign_nearest_grid(const void* x, int x_dtype, int64_t N, int d, const double* centers,
                            int K, const void* grid, int64_t grid_bytes, int32_t* labels32,
                            int64_t* labels64, float* residual, double* min_sqdist, void* stream);
int bdp_kmeans_lloyd_step_grid(const double* x, int64_t N, int d, const double* centers, int K,
                               const void* grid, int64_t grid_bytes, int32_t* labels, int64_t* acc,
                               int fix_hi_bits, int64_t* stats, double* inertia, int update,
                               void* stream);

/*
 * One whole Lloyd iteration in one call: acc_stats [K*(2d+1) + 2] (accumulators followed by the
 * {changed, unused} counters — the layout that is all-reduced across ranks) is zeroed, the key grid
 * is rebuilt for `centers` (grid == NULL: brute-force scan), the E+M step runs, and when centers_new
 * is given the M-step is finalised (single rank; with several ranks pass NULL, all-reduce acc_stats,
 * then call bdp_kmeans_finalize).
 */
int bdp_kmeans_iteration(const double* x, int64_t N, int d, const double* centers, int K, void* grid,
                         int64_t grid_bytes, int32_t* labels, int64_t* acc_stats, int fix_hi_bits,
                         double* inertia, int update, double* centers_new, double* shift2,
                         int64_t* n_empty, void* stream);

/*
 * M-step finalisation on the device (no host round trip): centers_new = sum / count from the
 * fixed-point accumulators (exactly rounded sum, then one division), empty clusters take the
 * centre of the heaviest cluster (sklearn _average_centers), shift2[0] = sum ||new - old||^2,
 * n_empty[0] = number of empty clusters (the caller runs relocation when non-zero).
 */
int bdp_kmeans_finalize(const int64_t* acc, int K, int d, int fix_hi_bits,
                        const double* centers_old, double* centers_new, double* shift2,
                        int64_t* n_empty, void* stream);

/*
 * Multi-GPU Lloyd iterations with the exchange fused into the M-step finalisation — no collective
 * library call, no memset and no host round trip inside the iteration loop.
 * Replaces learnKmeansDictionary.py:41-42 (KMeans.fit) at 1/2/4/8 GPUs.
 *
 * Exchange buffer (one per rank, bdp_kmeans_xchg_bytes(K, d) bytes, ZERO-initialised, 16-byte aligned,
 * allocated in memory every rank of the node can address — CUDA VMM / IPC / torch symmetric memory):
 *     int64 acc[2][K*(2d+1) + 2]     two accumulators {cluster sums, counts, changed, unused},
 *                                    used alternately by even and odd iterations
 *     uint64 flags[BDP_KMEANS_MAX_RANKS]   flags[r] = last iteration rank r has published here
 *     uint64 gflags[BDP_KMEANS_MAX_RANKS]  gflags[r] = last iteration whose key-grid slab rank r has
 *                                    stored into this rank's grid (sharded build)
 * xchg[r] is THIS process's address of rank r's buffer (xchg[rank] is the local one); xchg_multicast
 * is an NVLS multicast mapping of the same buffers (sums are then taken by one in-switch
 * multimem.ld_reduce per word) or NULL (one load per rank and word over NVLink).
 *
 * Control block: bdp_kmeans_ctl_bytes() zero-initialised device bytes; its head is a
 * struct bdp_kmeans_status the caller may read back (after synchronising the stream).
 *
 * bdp_kmeans_exchange_finalize   the exchange step of ONE iteration: publish this rank's accumulator
 *     acc[parity], wait for the peers', sum them (integer sums: bit-identical on every rank and for
 *     every world size), centers_new = sum / count, shift^2, empty-cluster census, stopping decision
 *     (check != 0: scikit-learn's rules — labels unchanged, or shift^2 <= tol_abs); zeroes
 *     acc[parity ^ 1].  flag_value = 1-based global iteration index (strictly increasing per fit).
 * bdp_kmeans_run   n_iters iterations back to back on one stream: iteration i = iter0 + k reads
 *     centers2[i & 1] ([2][K,d] fp64 ping-pong), rebuilds the key grid (grid == NULL: brute-force
 *     scan), runs the E+M step on this rank's rows into acc[i & 1] and calls the exchange step, which
 *     writes centers2[(i & 1) ^ 1].  Once the status leaves BDP_KMEANS_RUNNING the remaining launches
 *     do nothing; status.iter_done tells which centres buffer is current.
 *     incremental != 0 (needs the key grid): the accumulators PERSIST across iterations — the
 *     exchange step carries acc[parity] over to acc[parity ^ 1] instead of zeroing it and the E+M
 *     kernel only moves the rotations whose label changed (out of the old cluster, into the new one).
 *     The sums are integers, so this equals a recomputation bit for bit; after the first iterations
 *     few labels change and the M-step costs next to nothing.  `labels` must then be the labels the
 *     accumulators were built from (-1 and zeroed accumulators at the start of a fit).
 *     em_events (NULL: none): 4 * n_iters caller-owned cudaEvent_t handles (benchmark
 *     instrumentation; needs the key grid), recorded per iteration before the grid build, before the E+M
 *     kernel, after it, and after the exchange+finalise kernel.
 *     grid_peers (NULL: every rank builds the whole grid itself) = this process's addresses of
 *     every rank's key-grid buffer, allocated like the exchange buffers: the build is then SHARDED —
 *     a rank builds the cells of one slab of the grid and stores them into every rank's buffer over
 *     NVLink, and the E+M kernel waits until all slabs have arrived (flags in the exchange buffer).
 *     cells (NULL: the grid geometry follows the dictionary's bounding box and every cell is rebuilt
 *     every iteration) = ascending list of the n_cells coarse cells of a FIXED-geometry grid
 *     (bdp_keygrid_prepare + bdp_keygrid_occupancy) that hold rows of the fit, identical on every
 *     rank: only those cells are rebuilt (the rest of the box is never queried), a rank's slab is its
 *     share of the list, and the exchange kernel renews the build's counters and fp32 keys, so an
 *     iteration is three launches.  Labels do not depend on it: a query that reaches a cell no build
 *     has written scans the dictionary.
 *     BDP_KMEANS_NEEDS_HOST: an empty cluster appeared; the caller relocates (scikit-learn
 *     _relocate_empty_clusters_dense) on the summed accumulator, finalises with bdp_kmeans_finalize,
 *     clears status.state and resumes with iter0 = status.iter_done.
 */
#define BDP_KMEANS_MAX_RANKS 8
enum { BDP_KMEANS_RUNNING = 0, BDP_KMEANS_STRICT = 1, BDP_KMEANS_TOL = 2, BDP_KMEANS_NEEDS_HOST = 3 };
typedef struct bdp_kmeans_status {
  int32_t state, reserved;
  int64_t iter_done; /* iterations completed */
  int64_t changed;   /* labels changed in the last completed iteration, all ranks */
  int64_t n_empty;   /* empty clusters in the last completed iteration */
  double shift2;     /* sum ||c_new - c_old||^2 of the last completed iteration */
} bdp_kmeans_status;
int64_t bdp_kmeans_ctl_bytes(void);
int64_t bdp_kmeans_xchg_bytes(int K, int d);
int bdp_kmeans_exchange_finalize(void* const* xchg, const void* xchg_multicast, int world, int rank,
                                 int K, int d, int fix_hi_bits, int parity, int64_t flag_value,
                                 int check, int incremental, double tol_abs,
                                 const double* centers_old, double* centers_new, void* ctl,
                                 void* stream);
int bdp_kmeans_run(const double* x, int64_t N, int d, double* centers2, int K, void* grid,
                   int64_t grid_bytes, void* const* grid_peers, int32_t* labels, void* const* xchg,
                   const void* xchg_multicast, int world, int rank, int fix_hi_bits, int64_t iter0,
                   int n_iters, int check, int incremental, double tol_abs, void* ctl,
                   void* const* em_events, const int32_t* cells, int n_cells, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (a) category-conditioned bin-delta heads.
 *   bin_3layer / res_3layer           binDeltaModels.py:62-91 (fc1, bn1, relu, fc2, bn2, relu, fc3)
 *   OneBinDeltaModel.forward          binDeltaModels.py:112-121 (all C heads on all B, label select)
 *   soft mixing (joint model)         learnJointCatPoseModel_weighted.py:107-115
 *   ObjectNet one-hot-concat heads    objectnetHelperFunctions.py:110-172
 * Activations are batch-major [B, ld] (ld >= features, a multiple of 4): the GEMM puts the batch
 * on the 128 TMEM lanes and streams the weights on the N side, 176-256 rows of them per MMA, which is
 * what lets a small batch still pull the weights at HBM speed.
 * ---------------------------------------------------------------------------------------------- */

/*
 * Grouped TF32 tensor-core GEMM (tcgen05.mma, TMEM accumulators, TMA operand streaming):
 *     D_g[m, n] = sum_k A_g(m, k) * B_g(n, k),   g = 0..G-1,   fp32 in / fp32 accumulate / fp32 out
 * operand element (r, k), r = m or n:   major 0 (K-major):  P[g*gstride + r*ld + k]
 *                                       major 1 (MN-major): P[g*gstride + k*ld + r]
 * gstride == 0 with G > 1 shares the operand between all groups.  ld and gstride are in floats and
 * must be multiples of 4; base pointers 16-byte aligned (TMA).
 * output  c_layout 0: C[s*c_sstride + g*c_gstride + m*ldc + n]     1: C[... + n*ldc + m]
 * splits > 1 cuts K into bdp_gemm_tf32_splits(K, splits) ranges whose partial products land in
 * consecutive c_sstride slabs (the caller sums them: deterministic, no atomics).
 * precise 0: plain TF32 (operands truncated to 10 mantissa bits, ~1e-3 relative error);
 * precise 1: 3xTF32 — operand tiles are split into tf32 hi/lo halves in shared memory and three MMAs
 *            are accumulated per k-step: fp32-class results (~1e-6) at the same HBM traffic.
 * Serves nn.Linear forward (W x^T), input gradient (W^T dy) and weight gradient (dy a^T).
 */
int bdp_gemm_tf32(const float* A, int a_major, int64_t a_ld, int64_t a_gstride, const float* B,
                  int b_major, int64_t b_ld, int64_t b_gstride, float* C, int c_layout, int64_t ldc,
                  int64_t c_gstride, int64_t M, int64_t N, int64_t K, int G, int splits,
                  int64_t c_sstride, int precise, void* stream);
int bdp_gemm_tf32_splits(int64_t K, int splits);

/*
 * BatchNorm1d + ReLU (nn.BatchNorm1d defaults: eps 1e-5, momentum 0.1, biased variance for
 * normalisation, unbiased for the running estimate; binDeltaModels.py:66-73).
 * h [B, ld] pre-activation (GEMM output), F valid columns.
 * training != 0: batch statistics; save_mean / save_invstd [F] are written; running_mean / running_var
 *                [F] are updated in place (NULL: skip) — num_batches_tracked is the caller's.
 * training == 0: normalise with running_mean / running_var (save_mean / save_invstd, when given,
 *                receive running_mean and rsqrt(running_var + eps)).
 * a [B, ld] = relu(gamma * xhat + beta)   (columns F..ld-1 are not touched)
 */
int bdp_bn_relu_fwd(const float* h, int64_t F, int64_t B, int64_t ld, const float* gamma,
                    const float* beta, float* running_mean, float* running_var, float* save_mean,
                    float* save_invstd, float eps, float momentum, int training, float* a,
                    void* stream);
/*
 * Backward of the above: da [B, ld] gradient w.r.t. the post-ReLU output, a the saved output (ReLU
 * mask), h the saved pre-activation.  dh [B, ld], dgamma [F], dbeta [F].  training == 0 uses the
 * running statistics (mean = running_mean, invstd = rsqrt(running_var + eps)) passed in
 * save_mean / save_invstd and drops the batch-coupling terms.
 */
int bdp_bn_relu_bwd(const float* da, const float* a, const float* h, const float* gamma,
                    const float* save_mean, const float* save_invstd, int64_t F, int64_t B,
                    int64_t ld, int training, float* dh, float* dgamma, float* dbeta, void* stream);

/*
 * fc3 + category mixing.  a2 [B, ld] post-ReLU activations, head h in columns [h*N2, (h+1)*N2); w3
 * [H, O, N2], b3 [H, O]; mix [B, H] mixing weights (one-hot of the class label —
 * binDeltaModels.py:116-119 — or softmax probabilities — learnJointCatPoseModel_weighted.py:110-115).
 *   y[b, o] = sum_h mix[b,h] * (b3[h,o] + sum_j w3[h,o,j] * a2[b, h*N2 + j])          y [B, O]
 * Heads with mix[b,h] == 0 are skipped (exactly what the bmm with a one-hot produces).
 */
int bdp_head_fc3_fwd(const float* a2, int64_t ld, const float* w3, const float* b3,
                     const float* mix, int64_t B, int H, int O, int N2, float* y, void* stream);
/*
 * Backward of fc3 + mixing, dy [B, O]:
 *   da2 [B, ld]      = mix[b,h] * sum_o dy[b,o] w3[h,o,j]             (dense, zeros where mix == 0)
 *   dw3 [H, O, N2]   = sum_b mix[b,h] dy[b,o] a2[b, h*N2+j]           db3 [H, O] = sum_b mix[b,h] dy[b,o]
 *   dmix [B, H]      = sum_o dy[b,o] * (b3[h,o] + w3[h,o,:] . a2[b,h,:])   (NULL: skip)
 */
int bdp_head_fc3_bwd(const float* dy, const float* a2, int64_t ld, const float* w3,
                     const float* b3, const float* mix, int64_t B, int H, int O, int N2,
                     float* da2, float* dw3, float* db3, float* dmix, void* stream);

/* out[i] = sum_s parts[s*stride + i], i < n  (split-K reduction, fixed order) */
int bdp_sum_slabs(const float* parts, int64_t n, int S, int64_t stride, float* out, void* stream);
/*
 * Whole-head launch sequences: one call runs every kernel of the forward / backward pass of a stack
 * of H three-layer heads (OneBinDeltaModel.forward, binDeltaModels.py:112-121, and its autograd
 * backward).  The heads are split into fc3 groups (OneBinDeltaModel: bin heads with K outputs, res
 * heads with ndim outputs) that share the mixing weights mix [B, heads per group].
 */
#define BDP_HEAD_MAX_GROUPS 4
typedef struct bdp_head_desc {
  int32_t H, N0, N1, N2;              /* heads, fc1 in, fc1 out, fc2 out (multiples of 4) */
  int32_t n_groups, training, precise, reserved;
  int32_t group_heads[BDP_HEAD_MAX_GROUPS], group_out[BDP_HEAD_MAX_GROUPS];
  const float *w1, *g1, *be1;         /* [H,N1,N0], [H,N1], [H,N1] */
  const float *w2, *g2, *be2;         /* [H,N2,N1], [H,N2], [H,N2] */
  float *rm1, *rv1, *rm2, *rv2;       /* BatchNorm running statistics (updated when training) */
  const float* w3[BDP_HEAD_MAX_GROUPS];   /* [Hg, O_g, N2] */
  const float* b3[BDP_HEAD_MAX_GROUPS];   /* [Hg, O_g] */
  float eps, momentum;
} bdp_head_desc;

/* floats of the per-call saved-activation buffer: h1|a1 [B,H*N1], h2|a2 [B,H*N2], mean/invstd */
int64_t bdp_head_saved_floats(const bdp_head_desc* d, int64_t B);
/* floats of the backward scratch buffer (reusable across calls on one stream) */
int64_t bdp_head_bwd_workspace_floats(const bdp_head_desc* d, int64_t B);
/* x [B,N0], mix [B,Hg] -> y[g] [B,O_g]; `saved` is kept by the caller for the backward pass */
int bdp_head_forward(const bdp_head_desc* d, const float* x, const float* mix, int64_t B,
                     float* saved, float* const* y, void* stream);
/* dy[g] [B,O_g] -> stacked parameter gradients (dense, zeros included), dmix [B,Hg] (NULL: skip),
 * dx [B,N0] (NULL: skip) */
int bdp_head_backward(const bdp_head_desc* d, const float* x, const float* mix, int64_t B,
                      const float* saved, const float* const* dy, float* ws, float* dw1, float* dg1,
                      float* dbe1, float* dw2, float* dg2, float* dbe2, float* const* dw3,
                      float* const* db3, float* dmix, float* dx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BDPOSE_H_ */
