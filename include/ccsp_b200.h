/*
 * ccsp_b200.h — C ABI of libccsp_b200.so: the B200-native reverse-diffusion CCSP sampling path.
 *
 * The reference (zt-yang/diffusion-ccsp) is pure Python/PyTorch and has NO FFI for this path, so
 * there is no existing binding to mirror.  Each entry point below names the reference interface it
 * replaces (paths relative to the reference checkout); INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: opaque handles, raw pointers and sizes, no C++/torch types;
 *   - every function returns CCSP_OK (0) or a negative CcspStatus; ccsp_last_error() returns the
 *     message of the last failure on the calling thread; no exceptions cross the boundary;
 *   - "host" pointers are ordinary CPU memory; "dev" pointers are CUDA device pointers on the
 *     model's device, caller-owned, and must stay valid until the stream work completes;
 *   - kernels are enqueued on the `stream` argument (a cudaStream_t passed as void*; NULL = legacy
 *     default stream) and the calls do not synchronise unless documented;
 *   - a model/plan is bound to the CUDA device current at creation and is not thread-safe
 *     (use one per host thread / rank);
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     CCSP_ERR_CUDA.
 */
#ifndef CCSP_B200_H_
#define CCSP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CCSP_ABI_VERSION 1
#define CCSP_HIDDEN_DIM 256   /* hidden_dim of every published configuration (train_utils.py:104) */
#define CCSP_MAX_POSE_DIM 8
#define CCSP_MAX_TYPES 16

typedef enum {
  CCSP_OK = 0,
  CCSP_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  CCSP_ERR_CUDA = -2,      /* CUDA runtime error (message in ccsp_last_error) */
  CCSP_ERR_STATE = -3      /* call order violated (e.g. sample before plan) */
} CcspStatus;

/* arithmetic used for the two dense layers of the per-edge MLP (first layer pose part, decoder) */
typedef enum {
  CCSP_MATH_FP32 = 0,      /* FP32 FMA on CUDA cores (validation path, bit-for-bit deterministic) */
  CCSP_MATH_TF32X3 = 1,    /* tcgen05 kind::tf32, 3-term split (hi*hi + hi*lo + lo*hi), FP32 accumulate in TMEM */
  CCSP_MATH_BF16X3 = 2,    /* tcgen05 kind::f16 (bf16), 3-term split, FP32 accumulate in TMEM */
  CCSP_MATH_TF32 = 3,      /* single-pass TF32 (fast, ~1e-3) */
  CCSP_MATH_BF16 = 4       /* single-pass BF16 (fastest, ~1e-2) */
} CcspMath;

typedef struct CcspModel CcspModel;
typedef struct CcspPlan CcspPlan;

/*
 * Weights of networks/denoise_fn.py::ConstraintDiffuser (model='Diffusion-CCSP'), in the reference's
 * own layouts: every matrix is an nn.Linear.weight, row-major [out_features, in_features], FP32.
 * Pointers may be host or device memory (copied with cudaMemcpyDefault during ccsp_model_create).
 *   geom_encoder   denoise_fn.py:227-232   w0 [128,G]  b0 [128]  w2 [256,128] b2 [256]
 *   grasp_encoder  denoise_fn.py:236-241   w0 [128,Gr] ...        (NULL unless robot mode)
 *   pose_encoder   denoise_fn.py:245-250   w0 [128,P]  ...
 *   pose_decoder   denoise_fn.py:253-257   w0 [128,256] b0 [128] w2 [P,128] b2 [P]
 *   time_mlp       denoise_fn.py:259-264   w1 [1024,256] b1 [1024] w3 [256,1024] b3 [256]
 *   mlps[c]        denoise_fn.py:293-308   w [512, 256*(5 or 6)] b [512]; column blocks
 *                  [ (grasp_i) | geom_i | geom_j | pose_i | pose_j | time ]  (denoise_fn.py:346-354)
 */
typedef struct {
  int32_t hidden_dim;      /* must equal CCSP_HIDDEN_DIM */
  int32_t geom_dim;        /* G = dims[0][0] */
  int32_t pose_dim;        /* P = dims[-1][0], 1..CCSP_MAX_POSE_DIM */
  int32_t grasp_dim;       /* dims[1][0] in robot mode, else 0 */
  int32_t num_types;       /* len(constraint_sets) */
  int32_t normalize;       /* ConstraintDiffuser(normalize=...) */
  const float *geom_w0, *geom_b0, *geom_w2, *geom_b2;
  const float *grasp_w0, *grasp_b0, *grasp_w2, *grasp_b2;
  const float *pose_w0, *pose_b0, *pose_w2, *pose_b2;
  const float *dec_w0, *dec_b0, *dec_w2, *dec_b2;
  const float *time_w1, *time_b1, *time_w3, *time_b3;
  const float *const *mlp_w;   /* [num_types] */
  const float *const *mlp_b;   /* [num_types] */
} CcspModelDesc;

/* Schedule tables of networks/ddpm.py::GaussianDiffusion (ddpm.py:184-226), HOST float arrays [T]. */
typedef struct {
  int32_t T;                                   /* num_timesteps */
  const float *sqrt_recip_alphas_cumprod;      /* ddpm.py:213 */
  const float *sqrt_recipm1_alphas_cumprod;    /* ddpm.py:214 */
  const float *posterior_mean_coef1;           /* ddpm.py:223 */
  const float *posterior_mean_coef2;           /* ddpm.py:225 */
  const float *posterior_log_variance_clipped; /* ddpm.py:222 */
  const float *ula_grad_scale;                 /* _sqrt_recipm1_alphas_cumprod_custom, ddpm.py:215; NULL if no ULA */
  const float *step_sizes;                     /* eval(step_sizes), ddpm.py:207; NULL if no ULA */
  const int32_t *samples_per_step;             /* [T] ULA steps per timestep (ddpm.py:294-302); NULL = none */
  int32_t ebm_per_steps;                       /* denoise_fn.ebm_per_steps (ddpm.py:330); <=0 means 1 */
} CcspSchedule;

/* Where the Gaussian draws come from (reference: torch.randn at ddpm.py:121-122, 273, 292). */
typedef struct {
  const float *x_init;   /* dev [n,P] or NULL: start state replacing 0.5*randn (mask still pinned)        */
  const float *noise;    /* dev [1+sum_t(1+K_t), n, P] or NULL: injected draws in reference draw order     */
  uint64_t seed;         /* used when noise == NULL: in-kernel Philox4x32-10 keyed on                      */
  uint64_t node_offset;  /*   (seed, draw index, node_offset + node) so shards reproduce the global stream */
} CcspNoise;

const char *ccsp_last_error(void);
/* developer aid: where a device-side trap came from (source line or code << 40 | block << 24 | thread), readable even after the
 * CUDA context died; 0 = none */
unsigned long long ccsp_debug_trap_info(void);
/* developer aid (CCSP_PERSIST_TRACE=1): global-timer timeline of the persistent kernels, [event 0..7][iteration 0..31], ns */
unsigned long long ccsp_debug_persist_trace(int event, int iter);
/* Host-only (no device): the node-range boundaries {0, .., n} of the `want` (1..4) independent scene groups ("chains") that
 * ccsp_plan_create cuts a batch into when CCSP_CHAINS is set (a PyG batch is a disjoint union of scene graphs: the property
 * `Batch.batch` encodes in the reference, networks/denoise_fn.py:466-508 never crosses scenes).  bounds_out holds want + 1
 * entries; returns the number of groups (1 when the graph cannot be cut that often), -1 on bad arguments. */
int ccsp_debug_chain_cuts(const int64_t *edge_index, const float *edge_attr, int64_t n, int64_t E, int32_t num_types, int32_t want,
                          int64_t *bounds_out);
int ccsp_abi_version(void);
/* Number of kernels launched by this library on the calling thread since the last reset (bench.py's
 * `gpu_launches`). */
uint64_t ccsp_launch_count(void);
void ccsp_reset_launch_count(void);

/* Replaces: ConstraintDiffuser.__init__ + load_state_dict (denoise_fn.py:185-291, ddpm.py:503-514).
 * Packs the weights into kernel layouts on the current device. Synchronous. */
int ccsp_model_create(const CcspModelDesc *desc, CcspModel **out);
void ccsp_model_destroy(CcspModel *m);
int ccsp_model_set_math(CcspModel *m, int math /* CcspMath */);
int ccsp_model_get_math(const CcspModel *m);

/* Replaces: the per-call graph handling of ConstraintDiffuser.forward (denoise_fn.py:466-478, 508,
 * 313-339: re-upload of batch.x / edge_index, per-type `where`, gathers of the run-constant geometry
 * embeddings).  Inputs are HOST arrays in the reference's batch layout (data_transforms.py:181-200):
 *   x [n,F] f32, edge_index [2,E] i64, edge_attr [E] f32 (type ids), mask [n] i8.
 * pose_begin = dims[-1][1]; grasp_begin = dims[1][1] (ignored unless the model has a grasp encoder).
 * Sorts edges by type, builds the destination-CSR for the deterministic scatter, 1/sqrt(deg), the
 * pinned rows, and the per-edge static pre-activation (geometry/grasp part of mlps[c] + bias).
 * Returns after the plan is resident in HBM (synchronises `stream`). */
int ccsp_plan_create(CcspModel *m, const float *x, int64_t n, int32_t F,
                     const int64_t *edge_index, const float *edge_attr, const int8_t *mask, int64_t E,
                     int32_t pose_begin, int32_t grasp_begin, void *stream, CcspPlan **out);
void ccsp_plan_destroy(CcspPlan *p);
int64_t ccsp_plan_num_nodes(const CcspPlan *p);
int64_t ccsp_plan_num_edges(const CcspPlan *p);
/* padded edge rows actually processed by the edge kernels (multiple of the 128-row tile) */
int64_t ccsp_plan_num_edge_rows(const CcspPlan *p);

/* Replaces: ConstraintDiffuser.forward(poses_in, batch, t) (denoise_fn.py:453-537), non-energy branch.
 * poses dev [n,P] -> out dev [n,P]. Asynchronous on `stream`. */
int ccsp_denoise(CcspPlan *p, const float *poses, int32_t t, float *out, void *stream);

/* Replaces: GaussianDiffusion.p_sample_loop (ddpm.py:260-340) with p_sample (ddpm.py:245-258) and
 * AnnealedULASampler.sample_step (ddpm.py:955-966).  out dev [n,P]; history dev [T+1,n,P] or NULL
 * (ddpm.py:323-324, 335-336).  Asynchronous on `stream`. */
int ccsp_sample(CcspPlan *p, const CcspSchedule *sched, const CcspNoise *noise,
                float *out, float *history, void *stream);

/* Sampled device timing of the kernels ccsp_sample launches (feeds bench.py's `roofline`): when
 * stride > 0 every stride-th denoiser evaluation is bracketed by CUDA events recorded on the launch
 * stream.  ccsp_plan_get_timing synchronises on the recorded events, returns the accumulated
 * per-kernel times and resets the accumulators. No reference counterpart (the reference only has
 * wall-clock time.time() around p_sample_loop, ddpm.py:344-349). */
typedef struct {
  int64_t samples;      /* evaluations sampled */
  double ms_edge_l1;    /* first-layer kernel (or the fused edge kernel) */
  double ms_edge_dec;   /* decoder kernel (0 when fused into the first) */
  double ms_node;       /* scatter-reduce + update + pose-encoder kernel */
} CcspTiming;
int ccsp_plan_set_timing(CcspPlan *p, int32_t stride);
int ccsp_plan_get_timing(CcspPlan *p, CcspTiming *out);
/* bytes copied host->device by ccsp_plan_create for this plan (bench.py's e2e.h2d_bytes_per_step) */
int64_t ccsp_plan_h2d_bytes(const CcspPlan *p);

/* ---------------------------------------------------------------------------------------------------------
 * N1: the success check behind "solved scenes / s", whole batch in one launch.
 * Replaces: the per-graph CPU loop of Trainer.evaluate (networks/ddpm.py:620-713): clamp to [-1,1] (:620),
 * get_all_features (:807-821), NaN skip (:644-645), render_world_from_graph (envs/data_utils.py:221-357) ->
 * world.check_constraints_satisfied (envs/worlds.py:734-764 qualitative, :377-388 boxes) with python-fcl box-box
 * collisions among tiles and tray walls (envs/collisions.py:58-130; exclusions worlds.py:380-388, 398) and
 * compute_qualitative_constraints (envs/data_utils.py:427-621) + set inclusion (data_utils.py:418-424).
 * 2-D box worlds only (4-feature rows: collisions; 6-feature rows: collisions + the 13 qualitative relations);
 * triangle / 3-D / robot worlds need trimesh, FCL Convex or PyBullet and stay on the reference's CPU path. */
typedef enum { CCSP_WORLD_BOXES = 0, CCSP_WORLD_QUALITATIVE = 1 } CcspWorldKind;

typedef struct {
  int32_t kind;                   /* CcspWorldKind */
  int32_t num_scenes;             /* S */
  int32_t F, P, pose_begin;       /* row width of x, pose width, first pose column (dims[-1][1]) */
  int32_t clamp;                  /* 1: clamp the poses to [-1,1] first (ddpm.py:620) */
  const float *x;                 /* dev [n,F]  batch.x (geometry columns are read from here) */
  const int32_t *scene_node_ptr;  /* dev [S+1]  nodes of scene j = [ptr[j], ptr[j+1]), first one is the container; <= 28 tiles */
  const int32_t *scene_edge_ptr;  /* dev [S+1]  edges grouped by scene (batch.edge_extract); ignored for CCSP_WORLD_BOXES */
  const int32_t *edge_a, *edge_b; /* dev [E]    endpoints minus the scene's smallest edge index (ddpm.py:690-691) */
  const int32_t *edge_type;       /* dev [E]    int(edge_attr), >= 0; ids >= 13 are skipped (data_utils.py:180-181) */
  const float *world_dims;        /* dev [S,2]  (w_tray, l_tray) = batch.world_dims[j] */
} CcspCheckDesc;

/* poses dev [n,P] (the sampler's output, unclamped) -> solved dev u8 [S] (1 = no collision and no missing constraint);
 * counts dev i32 [S,2] or NULL: (#collisions, #missing constraints), (-1,-1) for NaN rows.  Asynchronous on `stream`. */
int ccsp_check_solved(const CcspCheckDesc *desc, const float *poses, uint8_t *solved, int32_t *counts, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * N2: the training step.  The reference computes loss = GaussianDiffusion(batch, debug=False, tag='EBM')
 * (networks/ddpm.py:387-389 -> p_losses :363-385 -> q_sample :353-361 with masked noise :114-117 -> ConstraintDiffuser.forward
 * denoise_fn.py:453-537 -> F.mse_loss :383) and gets the gradients from autograd (loss.backward(), ddpm.py:136-142, 533-534),
 * then torch.optim.Adam (ddpm.py:466, 542).  Here one call computes the loss and d(grad_scale * loss)/d(parameter) for every
 * parameter with hand-written FP32 kernels (fixed summation order: bit-reproducible). */
typedef struct CcspTrainGraph CcspTrainGraph;

typedef struct {
  int32_t hidden_dim;      /* must equal CCSP_HIDDEN_DIM */
  int32_t geom_dim, pose_dim, grasp_dim, num_types, normalize;   /* as in CcspModelDesc */
  int32_t row_width;       /* F: columns of batch.x */
  int32_t pose_begin;      /* dims[-1][1] */
  int32_t grasp_begin;     /* dims[1][1] in robot mode */
} CcspTrainDims;

/* Parameters (or their gradients) of ConstraintDiffuser as DEVICE pointers in the reference's nn.Linear layouts
 * ([out_features, in_features] row-major, see CcspModelDesc); mlp_w / mlp_b are HOST arrays of num_types device pointers. */
typedef struct {
  float *geom_w0, *geom_b0, *geom_w2, *geom_b2;
  float *grasp_w0, *grasp_b0, *grasp_w2, *grasp_b2;    /* NULL unless robot mode */
  float *pose_w0, *pose_b0, *pose_w2, *pose_b2;
  float *dec_w0, *dec_b0, *dec_w2, *dec_b2;
  float *time_w1, *time_b1, *time_w3, *time_b3;
  float *const *mlp_w;
  float *const *mlp_b;
} CcspParams;

/* Replaces: the per-call graph handling of ConstraintDiffuser.forward for a training batch (same inputs as ccsp_plan_create:
 * HOST x [n,F] f32, edge_index [2,E] i64, edge_attr [E] f32, mask [n] i8); allocates the activation buffers. Synchronous. */
int ccsp_train_graph_create(const CcspTrainDims *dims, const float *x, int64_t n, const int64_t *edge_index,
                            const float *edge_attr, const int8_t *mask, int64_t E, void *stream, CcspTrainGraph **out);
void ccsp_train_graph_destroy(CcspTrainGraph *g);
/* edges of constraint type c in the batch; a type without edges is not part of the reference's autograd graph, so its
 * mlps[c] gradients are None there (denoise_fn.py:514-515) — here they come back as zeros */
int64_t ccsp_train_graph_num_edges_of_type(const CcspTrainGraph *g, int32_t c);

/* Replaces: p_losses + loss.backward().  t = the batch's timestep (ddpm.py:388 draws ONE per batch), the two schedule scalars
 * are sqrt_alphas_cumprod[t] / sqrt_one_minus_alphas_cumprod[t] (ddpm.py:356-357), noise dev [n,P] = `all_noise` of p_losses, used as
 * given (conditional_noise zeroes the rows of pinned nodes, ddpm.py:114-117: the caller does that when it draws the noise).  loss_l1: 0 = 'l2' (mse), 1 = 'l1' (ddpm.py:380-383).
 * Writes *loss_out (dev scalar, the unscaled loss) and OVERWRITES every buffer of `grads` with d(grad_scale * loss)/d(param);
 * out_recon (dev [n,P], nullable) receives the denoiser output.  Asynchronous on `stream`. */
int ccsp_train_step(CcspTrainGraph *g, const CcspParams *weights, const CcspParams *grads, int32_t t,
                    float sqrt_alphas_cumprod_t, float sqrt_one_minus_alphas_cumprod_t, const float *noise,
                    int32_t loss_l1, float grad_scale, float *loss_out, float *out_recon, void *stream);

/* N4: the energy form of the denoiser.  Replaces: ConstraintDiffuser.forward(tag='EBM') with
 * energy_wrapper=True (denoise_fn.py:518-521, 539-548) and its wrapper ComposedEBMDenoiseFn (:57-83): the energy
 * E = sum over edges and both endpoints of |pose_decoder(...) - x[arg]|^2 (_compute_energy, :373-375) and its gradient
 * dE/dx, which the reference takes with torch.autograd.grad (:550-555) and here is derived analytically (back through the
 * decoder, the first layer's pose columns and the pose encoder, plus the direct -x term).  No normalisation and no pinning
 * in this branch (the sampler pins).  x dev [n,P] -> energy_out dev scalar, grad_out dev [n,P].  FP32; asynchronous. */
int ccsp_energy_grad(CcspTrainGraph *g, const CcspParams *weights, int32_t t, const float *x, float *energy_out,
                     float *grad_out, void *stream);

/* Replaces: torch.optim.Adam.step for one flat tensor (defaults: no weight decay, no amsgrad; ddpm.py:466): step >= 1 is the
 * 1-based update count.  All pointers dev [count].  Asynchronous on `stream`. */
int ccsp_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t count, int32_t step,
                   float lr, float beta1, float beta2, float eps, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CCSP_B200_H_ */
