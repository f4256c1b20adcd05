/*
 * b200i.h -- C ABI of the B200-native INSITE hot path (libb200insite.so).
 *
 * The reference (samholt/ODE-Discovery-for-Longitudinal-Heterogeneous-Treatment-Effects-Inference)
 * is pure Python and has no FFI for this path; its boundary is the Python call signatures listed in
 * SURVEY.md §8(b).  Each entry point below names the reference function it replaces (file:line under
 * the reference root).  The Python host layer (package dir, `cancer_simulation.py`, `sindy.py`) keeps
 * the reference signatures and calls these through ctypes; INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in `_host`; all arrays are row-major,
 *    C-contiguous, float64 unless stated, exactly the numpy layout of the reference dicts
 *    (SURVEY.md App. D), so `torch.from_numpy(a).cuda().data_ptr()` can be passed as is;
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 *    stream-ordered, nothing synchronises unless documented;
 *  - caller owns every buffer; nothing is allocated inside except where a function takes a
 *    workspace pointer + size;
 *  - return value: 0 ok; <0 argument error (B200I_E_*); >0 a cudaError_t.  b200i_last_error()
 *    returns a thread-local message for the last non-zero return.
 */
#ifndef B200I_H
#define B200I_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200I_API __attribute__((visibility("default")))
#else
#define B200I_API
#endif

#define B200I_OK 0
#define B200I_E_ARG (-1)        /* NULL pointer / negative size */
#define B200I_E_UNSUPPORTED (-2)/* configuration outside the kernels' range (e.g. lag != 0) */
#define B200I_E_ALIGN (-3)      /* pointer or row pitch not 16-byte aligned for the TMA path */
#define B200I_E_WORKSPACE (-4)  /* workspace too small */
#define B200I_E_DRIVER (-5)     /* cuTensorMapEncodeTiled unavailable / failed */

#define B200I_LS_ROBUST 0       /* line_search argument of the b200i_insite_bfgs* entry points */
#define B200I_LS_JAX 1

#define B200I_NUM_PARAMS 10     /* rows of the parameter block, order below */
/* parameter block `params` is (10, N) row-major, rows in this order (keys of the dict returned by
 * generate_params, cancer_simulation.py:195-203, 84-88):
 *   0 initial_volumes 1 alpha 2 rho 3 beta 4 beta_c 5 K
 *   6 chemo_sigmoid_intercepts 7 radio_sigmoid_intercepts 8 chemo_sigmoid_betas 9 radio_sigmoid_betas */

/* scalar constants of the simulator, computed by the host in float64 exactly as the reference
 * does (cancer_simulation.py:34-44, 231-241) and passed by value. */
typedef struct {
    double death_threshold;   /* TUMOUR_DEATH_THRESHOLD = calc_volume(13)            :44  */
    double cell_density;      /* TUMOUR_CELL_DENSITY = 5.8e8                          :43  */
    double sphere_coef;       /* 4 / 3 * pi  (calc_diameter denominator)              :39  */
    double chemo_amt;         /* 5.0                                                  :233 */
    double radio_amt;         /* 2.0                                                  :231 */
    double drug_decay;        /* exp(-log(2) / drug_half_life) = 0.5                  :338 */
    int32_t window_size;      /* 1..15                                                :252 */
    int32_t lag;              /* only 0 is supported by the CUDA path                 :253 */
} b200i_sim_consts;

B200I_API const char *b200i_last_error(void);
B200I_API int b200i_version(void);
/* number of SMs of the current device (grid sizing is done inside; exposed for reporting) */
B200I_API int b200i_device_sms(int *sms_out);

/* ------------------------------------------------------------------------------------------------
 * K1  simulate_factual -- cancer_simulation.py:218-375 (loop :282-354).
 * in : params (10,N); noise/recovery_rvs/chemo_rvs/radio_rvs (N,T) pre-drawn in the reference's
 *      order (:275-279; noise already multiplied by 0.01); assigned_actions (N,T,2) or NULL (:317).
 * out: nine (N,T) arrays + sequence_lengths (N,), keys of the dict at :356-367.  Every element of
 *      every output is written (zeros included) -- buffers need not be pre-zeroed.
 * variant: 0 = auto: the lean tiled kernel (csrc/sim_factual_ws.cuh) when T is even and all (N,T) pointers are
 *      16 B aligned -- one 16-column box per chunk when rows start on 128-byte lines (pitched entry point, = 12), two
 *      otherwise (= 10) -- else the generic kernel; 1 = generic thread-per-patient kernel (odd T, assigned_actions);
 *      2 = the first-generation TMA kernel (cross-check); 20, 21 = data-movement-only builds of 10 / 12 (profiling aid:
 *      outputs are copies of the draws).  See csrc/sim_factual.cu::dispatch_tma.
 * gram_workspace: NULL, or a workspace of b200i_gram_workspace_bytes() bytes: the kernel then also
 *      accumulates the population statistics of K4 on the fly (fused theta_gram; variants 1, 2, 10, 12) and
 *      leaves the reduced result in the first B200I_STATS_DOUBLES doubles of the workspace.  Measured slower than
 *      the two separate launches on B200 (2.0 vs 1.2 + 0.5 ms at 1M patients), so the pipeline does not use it.
 * ---------------------------------------------------------------------------------------------- */
B200I_API int b200i_sim_factual(int64_t n, int32_t T, const b200i_sim_consts *consts,
                      const double *params,
                      const double *noise, const double *recovery_rvs,
                      const double *chemo_rvs, const double *radio_rvs,
                      const double *assigned_actions,
                      double *cancer_volume, double *chemo_dosage, double *radio_dosage,
                      double *chemo_application, double *radio_application,
                      double *chemo_probabilities, double *radio_probabilities,
                      double *death_flags, double *recovery_flags, double *sequence_lengths,
                      const double *static_feature, double fd_dt, void *gram_workspace,
                      int32_t variant, void *stream);
/* The same with a row pitch (in elements, even, >= T) for the (N,T) arrays, cudaMallocPitch style: element (i,t)
 * lives at base[i * row_pitch + t].  A pitch that makes rows a multiple of 128 bytes (64 for T = 60) keeps every
 * 16-column box of the tiled kernel on 128-byte lines -- 1.23 instead of 1.46 ms per 1M patients on B200.  Only the
 * tiled kernels (variant 0, 10-13) take pitched rows. */
B200I_API int b200i_sim_factual_pitched(int64_t n, int32_t T, int64_t row_pitch, const b200i_sim_consts *consts,
                      const double *params,
                      const double *noise, const double *recovery_rvs,
                      const double *chemo_rvs, const double *radio_rvs,
                      const double *assigned_actions,
                      double *cancer_volume, double *chemo_dosage, double *radio_dosage,
                      double *chemo_application, double *radio_application,
                      double *chemo_probabilities, double *radio_probabilities,
                      double *death_flags, double *recovery_flags, double *sequence_lengths,
                      const double *static_feature, double fd_dt, void *gram_workspace,
                      int32_t variant, void *stream);

/* ------------------------------------------------------------------------------------------------
 * K4  theta_gram -- the data reduction behind SINDY.fit: constant-treatment snippeting
 * (pkpd/utils.py:433-462), order-1 finite differences and the [1,x0,u0,x0*u0] library of the
 * pysindy call at sindy.py:203-213, reduced to per-treatment normal equations, plus the moments
 * get_scaling_params needs (cancer_simulation.py:776-796).
 * in : cancer_volume/chemo_application/radio_application (N,T), sequence_lengths (N,) float64,
 *      static_feature (N,) float64 (the un-scaled patient type), chemo_dosage/radio_dosage (N,T) or
 *      NULL (moments of those two are then left 0).
 * out: stats[B200I_STATS_DOUBLES] (layout below), reduced over all N patients in a fixed order
 *      (deterministic for a fixed N and launch shape).
 * workspace: b200i_gram_workspace_bytes() bytes, 16 B aligned; stats live at its start.
 * ---------------------------------------------------------------------------------------------- */
#define B200I_GRAM_PER_TREATMENT 15 /* 10 upper-triangular Gram entries (row-major: 00 01 02 03 11 12 13 22 23 33),
                                       4 right-hand sides (Theta^T xdot), 1 sample count */
#define B200I_MOMENTS 8             /* sum v, sum v^2, sum C, sum C^2, sum d, sum d^2 over active entries; active count; N */
#define B200I_STATS_DOUBLES (4 * B200I_GRAM_PER_TREATMENT + B200I_MOMENTS)
B200I_API int64_t b200i_gram_workspace_bytes(void);
B200I_API int b200i_theta_gram(int64_t n, int32_t T, double fd_dt,
                     const double *cancer_volume, const double *chemo_application,
                     const double *radio_application, const double *sequence_lengths,
                     const double *static_feature,
                     const double *chemo_dosage, const double *radio_dosage,
                     void *gram_workspace, void *stream);
/* mode 1 = the joint model's reduction (sindy.py:160-171 with joint_model=True; pkpd/utils.py:493-497, 656-672): one
 * trajectory per patient, x_k = cancer_volume[k+1] for k < sequence_length, inputs of sample k = the applications of
 * column k, order-1 finite differences over the whole trajectory; same 15 statistics per treatment code, from which
 * b200i_stlsq_joint assembles the 11x11 normal equations.  mode 0 = b200i_theta_gram_pitched. */
B200I_API int b200i_theta_gram_mode(int64_t n, int32_t T, int64_t row_pitch, int32_t mode, double fd_dt,
                     const double *cancer_volume, const double *chemo_application,
                     const double *radio_application, const double *sequence_lengths,
                     const double *static_feature,
                     const double *chemo_dosage, const double *radio_dosage,
                     void *gram_workspace, void *stream);
/* Joint-model STLSQ + unbias (pysindy semantics as b200i_stlsq_population) on the 11-term library
 * [1,x0,u0,u1,u2,x0u0,x0u1,x0u2,u0u1,u0u2,u1u2] (u0 chemo, u1 radio, u2 static feature).
 * out: coefs11 (11,), support11 (11,) int32, coefs44 (4,4): the expression with |c| > drop_below restricted to each
 *      treatment code chemo + 2*radio, i.e. what b200i_ode_rollout integrates (pass drop_below < 0 there). */
B200I_API int b200i_stlsq_joint(const double *stats, double threshold, double alpha, int32_t max_iter, double drop_below,
                     double *coefs11, int32_t *support11, double *coefs44, void *stream);
/* Lean fit of the device-resident pipeline.  b200i_sim_factual_side = b200i_sim_factual_pitched (tiled kernel, no
 * assigned_actions, no fused statistics) plus two side outputs of the simulator kernel: codes_out (N, code_pitch)
 * uint8 = chemo + 2*radio application per step, and patient_moments_out (6, N) = per-patient sums over the active
 * entries of volume, volume^2, chemo dosage, its square, radio dosage, its square.  b200i_theta_gram_codes finishes
 * the population statistics from cancer_volume + those two (0.6 instead of 2.4 GB per million patients); same
 * statistics layout and reduction as b200i_theta_gram_mode.  code_pitch: multiple of 16, >= T rounded up to 16;
 * moments_stride: elements between the six rows of patient_moments (0 = n; larger when the launch covers a row range
 * of a bigger cohort). */
B200I_API int b200i_sim_factual_side(int64_t n, int32_t T, int64_t row_pitch, const b200i_sim_consts *consts,
                      const double *params,
                      const double *noise, const double *recovery_rvs,
                      const double *chemo_rvs, const double *radio_rvs,
                      double *cancer_volume, double *chemo_dosage, double *radio_dosage,
                      double *chemo_application, double *radio_application,
                      double *chemo_probabilities, double *radio_probabilities,
                      double *death_flags, double *recovery_flags, double *sequence_lengths,
                      uint8_t *codes_out, int64_t code_pitch, double *patient_moments_out,
                      int32_t variant, void *stream);
B200I_API int b200i_theta_gram_codes(int64_t n, int32_t T, int64_t row_pitch, int32_t mode, double fd_dt,
                     const double *cancer_volume, const uint8_t *codes, int64_t code_pitch,
                     const double *sequence_lengths, const double *static_feature,
                     const double *patient_moments, int64_t moments_stride, void *gram_workspace, void *stream);
/* K1L  simulate_factual with device-generated draws (throughput mode, SURVEY.md 8d "lean variant").
 * The reference draws its four (N,T) arrays from numpy's sequential global stream (cancer_simulation.py:275-279);
 * here they come from Philox4x32-10 counted by (patient_base + i, column / 2, stream) and keyed by `seed`
 * (csrc/philox.cuh), so a patient's draws are independent of the launch shape, the shard and the number of GPUs.
 * b200i_philox_draws writes exactly those draws as the four arrays of the K1 contract (noise already x 0.01);
 * b200i_sim_factual on them reproduces b200i_sim_factual_rng bit for bit.
 * b200i_sim_factual_rng: in : params (10, params_stride >= N): a launch may cover a column range of a larger block
 *        (chunked host->device pipelines), seed, patient_base (global index of row 0).
 *   out: cancer_volume (N,T) with row_pitch (even, >= T); codes_out (N, code_pitch) uint8 = chemo + 2*radio
 *        application per step, or NULL (code_pitch: multiple of 16, >= T rounded up to 16; bytes >= T of a row are
 *        written as 0 up to T rounded up to 16); sequence_lengths (N,) float64; patient_moments_out
 *        (6, moments_stride >= N) as b200i_sim_factual_side, or NULL.
 *   gram_workspace != NULL (with static_feature (N,), fd_dt > 0): the population statistics of K4 (Gram + moments,
 *        same layout and meaning as b200i_theta_gram) are accumulated in the same kernel; patient_moments_out is
 *        then ignored.  T even, 4..1024.
 *   variant: 0 = default (= 2: phased kernel, generator loop and one-column simulator loop that each fit the L0
 *        instruction cache, 16 warps per SM); 1 = first generation (four unrolled columns, generator inlined);
 *        both give identical bits. */
B200I_API int b200i_philox_draws(int64_t n, int32_t T, int64_t row_pitch, uint64_t seed, int64_t patient_base,
                     double *noise, double *recovery_rvs, double *chemo_rvs, double *radio_rvs, void *stream);
B200I_API int b200i_sim_factual_rng(int64_t n, int32_t T, int64_t row_pitch, const b200i_sim_consts *consts,
                     const double *params, int64_t params_stride, uint64_t seed, int64_t patient_base,
                     double *cancer_volume, uint8_t *codes_out, int64_t code_pitch, double *sequence_lengths,
                     double *patient_moments_out, int64_t moments_stride, const double *static_feature, double fd_dt,
                     void *gram_workspace, int32_t variant, void *stream);
/* Host-resident parameters: params_host (10,N) and static_host (N,) [or NULL] are PINNED HOST arrays; they are copied
 * into params / static_feature in `chunks` column ranges on copy_stream, and every range is simulated on `stream`
 * (b200i_sim_factual_rng with the per-patient moments) as soon as it has arrived, so the PCIe transfer overlaps the
 * simulation.  Same outputs as one b200i_sim_factual_rng launch over all N patients.
 * uniform_mask: bit r set = parameter row r is one scalar for the whole cohort (the reference's generate_params builds
 * K and its four sigmoid rows that way, cancer_simulation.py:83-88, :202): the row is not read from params_host but filled on the
 * device with uniform_values_host[r] (host array of 10 doubles; NULL when the mask is 0).
 * chunk_gram_workspaces != NULL (chunks x b200i_gram_workspace_bytes() bytes, with fd_dt > 0 and stats_out[68]): each
 * range's share of the population statistics (b200i_theta_gram_codes) is launched right behind its simulation and the
 * shares are summed in range order into stats_out, so only the last range's work is left when the copy ends. */
B200I_API int b200i_upload_simulate_rng(int64_t n, int32_t T, int64_t row_pitch, const b200i_sim_consts *consts,
                     const double *params_host, uint32_t uniform_mask, const double *uniform_values_host,
                     const double *static_host, double *params, double *static_feature,
                     uint64_t seed, int64_t patient_base, double *cancer_volume, uint8_t *codes_out, int64_t code_pitch,
                     double *sequence_lengths, double *patient_moments_out, int32_t chunks,
                     double fd_dt, void *chunk_gram_workspaces, double *stats_out,
                     void *copy_stream, void *stream);
/* The same from the REDUCED parameter set: what get_standard_params actually draws per patient (cancer_simulation.py:
 * 96-215: initial volume, alpha, rho, beta_c and the patient type) crosses PCIe, everything generate_params derives from
 * it is rebuilt on the device exactly as the reference builds it -- beta = alpha / 10 (:185, derive_beta != 0; the same
 * IEEE division), the static feature = the patient type (one byte per patient: patient_types_host pinned uint8 (N,),
 * patient_types_dev a device staging buffer of N bytes), and the cohort-wide scalar rows through uniform_mask.  With
 * K and the four sigmoid rows uniform that is 33 instead of 88 bytes per patient. */
B200I_API int b200i_upload_simulate_rng_reduced(int64_t n, int32_t T, int64_t row_pitch, const b200i_sim_consts *consts,
                     const double *params_host, uint32_t uniform_mask, const double *uniform_values_host,
                     int32_t derive_beta, const uint8_t *patient_types_host, uint8_t *patient_types_dev,
                     double *params, double *static_feature, uint64_t seed, int64_t patient_base,
                     double *cancer_volume, uint8_t *codes_out, int64_t code_pitch, double *sequence_lengths,
                     double *patient_moments_out, int32_t chunks, double fd_dt, void *chunk_gram_workspaces,
                     double *stats_out, void *copy_stream, void *stream);
/* b200i_upload_simulate_rng_reduced for a STREAM of cohorts: consecutive calls overlap.  The copies of a call do not wait
 * for everything queued on `stream` before it (the previous step's all-reduce, STLSQ and result copies) but, chunk by
 * chunk, only for the previous call's kernels that read that chunk's parameter rows; the kernels stay ordered behind
 * `stream`.  The caller keeps the pinned inputs of a call unchanged until copy_stream has passed the call (record an
 * event on copy_stream after it) and reads a call's results after an event recorded on `stream` behind it
 * (cohort.GeneratedFitPipeline.submit / HostStep).  Same outputs, bit for bit, as the serial entry point. */
B200I_API int b200i_upload_simulate_rng_pipelined(int64_t n, int32_t T, int64_t row_pitch, const b200i_sim_consts *consts,
                     const double *params_host, uint32_t uniform_mask, const double *uniform_values_host,
                     int32_t derive_beta, const uint8_t *patient_types_host, uint8_t *patient_types_dev,
                     double *params, double *static_feature, uint64_t seed, int64_t patient_base,
                     double *cancer_volume, uint8_t *codes_out, int64_t code_pitch, double *sequence_lengths,
                     double *patient_moments_out, int32_t chunks, double fd_dt, void *chunk_gram_workspaces,
                     double *stats_out, void *copy_stream, void *stream);
/* The same for (N,T) arrays with a row pitch (elements, even, >= T); see b200i_sim_factual_pitched. */
B200I_API int b200i_theta_gram_pitched(int64_t n, int32_t T, int64_t row_pitch, double fd_dt,
                     const double *cancer_volume, const double *chemo_application,
                     const double *radio_application, const double *sequence_lengths,
                     const double *static_feature,
                     const double *chemo_dosage, const double *radio_dosage,
                     void *gram_workspace, void *stream);

/* ------------------------------------------------------------------------------------------------
 * K5  population STLSQ -- pysindy STLSQ(threshold, alpha, max_iter) + unbias refit, semantics of the
 * vendored copy pkpd/utils.py:244-327 and of sindy.py:194,203-213: per treatment, on the 4x4 normal
 * equations: repeat { solve (G[ind,ind] + alpha I) c = b[ind]; ind &= |c| >= threshold } until the
 * support stops changing (<= max_iter), then OLS on the final support.
 * in : stats (device, B200I_STATS_DOUBLES, e.g. after the cross-GPU all-reduce).
 * out: coefs (4,4) device float64 = SINDY.joint_coefs (sindy.py:334); support (4,4) int32 device.
 * ---------------------------------------------------------------------------------------------- */
B200I_API int b200i_stlsq_population(const double *stats, double threshold, double alpha, int32_t max_iter,
                           double *coefs, int32_t *support, void *stream);

/* ------------------------------------------------------------------------------------------------
 * K6  ode_rollout -- SINDY._get_non_fine_tuned_predictions (sindy.py:371-431) and
 * predict_with_reduced_coefs (sindy.py:767-778) with the Euler integrator of pkpd/utils.py:68-90:
 * open loop from x0, `substeps` explicit-Euler sub-steps of dt/substeps per interval, the ODE of the
 * interval chosen by the treatment code (argmax of the one-hot, sindy.py:310,499).
 * in : x0 (R,), static_feature (R,), codes (R,W) uint8 in 0..3, coefs (4,4) shared by all rows
 *      (coefs_per_row == 0) or (R,4,4) (coefs_per_row == 1).  Terms with |c| <= drop_below are
 *      dropped (1e-3 for the population path, pkpd/utils.py:388; pass a negative value to keep all).
 * out: pred (R,W) un-scaled volumes.
 * ---------------------------------------------------------------------------------------------- */
B200I_API int b200i_ode_rollout(int64_t rows, int32_t W, double dt, int32_t substeps,
                      const double *x0, const double *static_feature, const uint8_t *codes,
                      const double *coefs, int32_t coefs_per_row, double drop_below,
                      double *pred, void *stream);
/* FP32 variant (BASELINE config C4, "FP32 vs FP64"): same interface, state / coefficients / Euler steps in float32. */
B200I_API int b200i_ode_rollout_f32(int64_t rows, int32_t W, double dt, int32_t substeps,
                      const double *x0, const double *static_feature, const uint8_t *codes,
                      const double *coefs, int32_t coefs_per_row, double drop_below,
                      double *pred, void *stream);

/* ------------------------------------------------------------------------------------------------
 * treatment codes: code = chemo + 2*radio of the (R,Wfull) application arrays, columns [0,W)
 * (dataset.py:127-141: one-hot index of [none, chemo, radio, both]).
 * ---------------------------------------------------------------------------------------------- */
B200I_API int b200i_treatment_codes(int64_t rows, int32_t W, int32_t row_pitch,
                          const double *chemo_application, const double *radio_application,
                          uint8_t *codes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * masked squared-error sums for TimeVaryingCausalModel.get_normalised_masked_rmse /
 * get_normalised_n_step_rmses (time_varying_model.py:236-313): for predictions and targets (R,W)
 * in un-scaled units and integer active lengths (active_entries[i,:len]=1, dataset.py:162-164):
 * out: sums (3*W + 2) doubles: se_col[W] (sum over rows of active squared error per column),
 *      cnt_col[W] (active rows per column), se_last_col[W] (squared error at the last active
 *      entry, binned by column), then total se_last and count_last.  W <= 128.
 * workspace: b200i_masked_se_workspace_bytes() bytes (block partials; ordered, atomics-free sum).
 * ---------------------------------------------------------------------------------------------- */
B200I_API int64_t b200i_masked_se_workspace_bytes(void);
B200I_API int b200i_masked_se(int64_t rows, int32_t W, const double *pred, const double *target,
                    const int32_t *active_len, double *sums, void *workspace, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Counterfactual generators -- compact, device-resident representation.
 *
 * The reference materialises one dense row per (patient, t, option) (R ~ 227 N rows of width T for
 * the one-step set, R ~ 562 N rows of width T+H for treatment sequences, cancer_simulation.py:423-430,
 * :621-630): 0.9 TB at N = 1M.  Here a cohort is stored per patient:
 *   factual       (N,T)    float64  clipped factual trajectory F (F[t+1] written at step t)
 *   codes         (N,T)    uint8    factual option index at step t: 2*chemo + radio (reference option
 *                                   order (0,0),(0,1),(1,0),(1,1), :513); 0 after the last step
 *   n_steps       (N,)     int32    executed steps (t_last + 1)
 *   one-step:  cf (N,T-1,4)   float64  un-clipped next volume under each of the 4 options (:536-538)
 *   tr.-seq.:  cf (N,T-1,2H,H) float64 the H projected volumes of each of the 2H sliding options
 *                                   (:721-743); valid (N,T-1) uint16 bit o set iff option o produced
 *                                   a row (no NaN, :745-746)
 *   n_rows        (N,)     int32    rows the reference would have emitted for this patient
 *   row_offsets   (N+1,)   int64    exclusive prefix sum of n_rows = reference row index of the
 *                                   patient's first row (test_idx at :432/:632)
 * b200i_expand_* turn a row range of this into the reference's dense arrays.
 *
 * Cross-row window: the reference computes patient i's treatment probabilities from OUTPUT ROW i
 * (cancer_simulation.py:471, :671), i.e. from a row emitted by an earlier patient j(i).  A cohort is
 * therefore simulated in dependency levels (level k+1 = the patients whose row index falls into rows
 * emitted by level k); the level loop runs inside these calls and synchronises the stream once per
 * level (depth ~ log_227 N resp. log_562 N).  For a shard of a larger cohort pass `source`: the
 * compact arrays of the global prefix of patients that own rows [0, global_base + n), simulated
 * redundantly on every rank; then no level loop is needed.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int64_t n;                    /* patients in the source cohort (global indices 0..n-1) */
    const double *factual;        /* (n,T) */
    const uint8_t *codes;         /* (n,T) */
    const double *cf;             /* (n,T-1,4) or (n,T-1,2H,H) */
    const uint16_t *valid;        /* (n,T-1) treatment-seq only, else NULL */
    const int64_t *row_offsets;   /* (n+1,) */
} b200i_cf_source;

/* K2  simulate_counterfactual_1_step -- cancer_simulation.py:378-563 (loop :435-552).
 * draws: per-patient rows in the reference's order (:440-453): noise (N,T) [x0.01], recovery, chemo,
 * radio (N,T).  total_rows_host (may be NULL) receives row_offsets[N] (synchronises). */
B200I_API int b200i_sim_cf_one_step(int64_t n, int32_t T, const b200i_sim_consts *consts, const double *params,
                          const double *noise, const double *recovery_rvs, const double *chemo_rvs,
                          const double *radio_rvs, int64_t global_base, const b200i_cf_source *source,
                          double *factual, uint8_t *codes, double *cf, int32_t *n_steps, int32_t *n_rows,
                          int64_t *row_offsets, int64_t *total_rows_host, int32_t *levels_host, void *stream);

/* K3  simulate_counterfactuals_treatment_seq, cf_seq_mode='sliding_treatment' --
 * cancer_simulation.py:566-773 (loop :635-760).  noise is (N,T+H) [x0.01] (:640). */
B200I_API int b200i_sim_cf_treatment_seq(int64_t n, int32_t T, int32_t H, const b200i_sim_consts *consts,
                               const double *params, const double *noise, const double *recovery_rvs,
                               const double *chemo_rvs, const double *radio_rvs, int64_t global_base,
                               const b200i_cf_source *source, double *factual, uint8_t *codes, double *cf,
                               uint16_t *valid, int32_t *n_steps, int32_t *n_rows, int64_t *row_offsets,
                               int64_t *total_rows_host, int32_t *levels_host, void *stream);

/* dense reference rows [row_begin, row_end) of a compact cohort (keys of the dicts at :554-559 and
 * :762-769).  Output arrays hold (row_end-row_begin) rows of width T (one-step) / T+H (sequences). */
B200I_API int b200i_expand_cf_one_step(int64_t n, int32_t T, const double *factual, const uint8_t *codes,
                             const double *cf, const int64_t *row_offsets, const double *patient_types,
                             int64_t row_begin, int64_t row_end, double *cancer_volume,
                             double *chemo_application, double *radio_application, double *sequence_lengths,
                             double *patient_types_rows, void *stream);
B200I_API int b200i_expand_cf_treatment_seq(int64_t n, int32_t T, int32_t H, const double *factual,
                                  const uint8_t *codes, const double *cf, const uint16_t *valid,
                                  const int64_t *row_offsets, const double *patient_types, int64_t row_begin,
                                  int64_t row_end, double *cancer_volume, double *chemo_application,
                                  double *radio_application, double *sequence_lengths,
                                  double *patient_types_rows, double *patient_ids, double *patient_current_t,
                                  void *stream);

/* ------------------------------------------------------------------------------------------------
 * Individualisation of the population ODE.  Row interface shared by both estimators:
 *   x (R,W) float64       un-scaled prev_outputs of the processed dataset (sindy.py:555-556)
 *   codes (R,W) uint8     treatment index per step = argmax of the one-hot current_treatments
 *                         (chemo + 2*radio; sindy.py:499)
 *   the fit window of a row is its first n_fit = sequence_length - projection_horizon transitions
 *   x[k] -> x[k+1] (mask of f_to_min_func, sindy.py:786)
 *   output coefs (R,4,4): per-row coefficient matrix for b200i_ode_rollout(coefs_per_row = 1)
 *
 * K5b b200i_stlsq_batched: per row and treatment, sequentially thresholded ridge regression shrunk to
 *   the population coefficients on the population support:
 *     repeat { c = argmin (1/n)|Theta c - xdot|^2 + lam |c - prior|^2 on the support; drop |c| < threshold }
 *   (normal equations from the same snippet / finite-difference / library rules as b200i_theta_gram; 4x4
 *   Cholesky in registers).  Support = |prior| > support_tol.  Treatments that do not occur in the window
 *   keep the prior.  fit_len (R,) int32 = n_fit, already clamped by the caller to [0, W-1].
 *   Lineage: determine_individualized_equation_coefs, pkpd_simulation.py:791-836 (dormant in the reference).
 *
 * K7 b200i_insite_bfgs: the reference's live estimator, _fine_tuning_inner (sindy.py:587-631) with
 *   f_to_min_func (:781-794): jax.scipy.optimize.minimize(method='BFGS') over all 16 coefficients from theta0.
 *   Rows with sequence_length <= projection_horizon keep theta0 (:571-585).
 *   line_search:
 *     B200I_LS_JAX    the semantics of jax's minimize_bfgs / line_search / _zoom (un-vendored dependency of the reference,
 *                     restated): start step min(1, 1.01 * 2 (f_k - f_{k-1}) / dphi_0), bracketing by doubling (10 trials),
 *                     zoom with cubic / quadratic / bisection trial points, and jax's FAILURE rules -- zoom fails when the
 *                     signed bracket width a_hi - a_lo is <= 1e-10 (so at once for a reversed bracket) or after 30
 *                     trials; a failed line search ends BFGS with status 3 and leaves x + a p behind, a = the last trial
 *                     point if it satisfied both Wolfe conditions, else 1.  Call it with gtol = 1e-5 and max_iter =
 *                     200 * (number of coefficients): jax ignores the reference's tol=1e-12 argument (minimize() does
 *                     not forward `tol` to minimize_bfgs).  With these settings the two INSITE lines of the reference's
 *                     committed logs are reproduced to 5e-15 / 2e-6 relative (tests/test_gpu_insite.py).
 *     B200I_LS_ROBUST Nocedal-Wright strong-Wolfe search that accepts the best sufficient-decrease point when the
 *                     curvature test cannot be met (FP64 noise floor) and never returns a point worse than theta0.
 *   status_out (R,) int32: low byte 0 converged / 1 max_iter / 3 zoom failed (the rows the reference replaces by
 *   theta0, sindy.py:628-631) / 5 bracketing exhausted / 4 perfect start / 6 no improvement (robust mode: theta0
 *   kept) / -2 skipped; bits 8.. = BFGS iterations.  fval_out (R,2) = objective at theta0 and at the returned
 *   coefficients.
 * ---------------------------------------------------------------------------------------------- */
B200I_API int b200i_stlsq_batched(int64_t rows, int32_t W, double fd_dt, const double *x, const uint8_t *codes,
                        const int32_t *fit_len, const double *static_feature, const double *prior,
                        double support_tol, double lam, double threshold, int32_t max_iter,
                        double *coefs_out, void *stream);
B200I_API int b200i_insite_bfgs(int64_t rows, int32_t W, double dt, int32_t substeps, const double *x,
                      const uint8_t *codes, const int32_t *sequence_lengths, int32_t projection_horizon,
                      const double *static_feature, const double *theta0, double lam, double gtol,
                      int32_t max_iter, int32_t line_search, double *coefs_out, int32_t *status_out, double *fval_out,
                      void *stream);
/* The same for the joint ("one ODE") model, sindy.py:503-517 with joint_model=True: theta0 / coefs_out hold the 11
 * coefficients of [1, x0, u0, u1, u2, x0 u0, x0 u1, x0 u2, u0 u1, u0 u2, u1 u2] (x0 volume, u0 chemo, u1 radio
 * application, u2 static feature); codes = chemo + 2*radio application per step; coefs_out is (rows, 11); the
 * penalty is lam * mean over the 11 coefficients. */
B200I_API int b200i_insite_bfgs_joint(int64_t rows, int32_t W, double dt, int32_t substeps, const double *x,
                      const uint8_t *codes, const int32_t *sequence_lengths, int32_t projection_horizon,
                      const double *static_feature, const double *theta0, double lam, double gtol,
                      int32_t max_iter, int32_t line_search, double *coefs_out, int32_t *status_out, double *fval_out,
                      void *stream);


/* ------------------------------------------------------------------------------------------------
 * Evaluation on COMPACT counterfactual cohorts (BASELINE config C3: tau = 1..5 rollouts for 1M patients).
 *
 * The reference evaluates one dense row per (patient, t, option): SINDY.get_predictions /
 * get_autoregressive_predictions (sindy.py:371-431, 433-715, 717-760) followed by
 * get_normalised_masked_rmse(one_step_counterfactual=True) / get_normalised_n_step_rmses
 * (time_varying_model.py:236-313).  All rows of one (patient, t) share the factual prefix and -- in INSITE mode -- the
 * fit problem (fit window = F[0..t] for the one-step rows, sindy.py:786 with projection_horizon 1, F[0..t+1] for the
 * sequence rows with projection_horizon H), so these entry points read the per-patient arrays of K2 / K3 directly.
 *
 * coefs: (4,4) for the whole cohort (coefs_per_step == 0; rows = treatment code chemo + 2*radio, as
 *        b200i_stlsq_population writes them) or (n, T-1, 4, 4) per (patient, t) (coefs_per_step == 1; what the two
 *        *_prefix entry points below write).  Terms with |c| <= drop_below are dropped (pass < 0 to keep all).
 * codes are the generators' factual option indices 2*chemo + radio; n_steps = executed steps per patient.
 * b200i_cf_eval_one_step:     sums (3*(T-1) + 2) doubles, the layout of b200i_masked_se over the (R, T-1) rows.
 * b200i_cf_eval_treatment_seq: sums (2*H) doubles: squared error per projection step, then the number of scored rows
 *                              (valid options) per step.  T <= 128, H <= 8.
 * Deterministic (ordered, atomics-free reduction) for a fixed n on a given device.
 * ---------------------------------------------------------------------------------------------- */
B200I_API int b200i_cf_eval_one_step(int64_t n, int32_t T, double dt, int32_t substeps, const double *factual,
                           const uint8_t *codes, const double *cf, const int32_t *n_steps,
                           const double *static_feature, const double *coefs, int32_t coefs_per_step,
                           double drop_below, double *sums, void *stream);
B200I_API int b200i_cf_eval_treatment_seq(int64_t n, int32_t T, int32_t H, double dt, int32_t substeps,
                           const double *factual, const uint8_t *codes, const double *cf, const uint16_t *valid,
                           const int32_t *n_steps, const double *static_feature, const double *coefs,
                           int32_t coefs_per_step, double drop_below, double *sums, void *stream);
/* Individualisation per (patient, t) of a compact cohort: row (i,t) fits on the first t + fit_offset transitions of
 * the patient's factual trajectory (fit_offset 0: one-step rows, 1: sequence rows); t >= n_steps[i] or an empty
 * window keeps theta0 / the prior.  Same estimators, arguments and outputs as b200i_insite_bfgs /
 * b200i_stlsq_batched, with coefs_out (n, T-1, 4, 4), status_out (n, T-1), fval_out (n, T-1, 2): one fit instead of
 * the 4 (one-step) or <= 2H (sequence) identical ones the dense rows repeat. */
B200I_API int b200i_insite_bfgs_prefix(int64_t n, int32_t T, int32_t fit_offset, double dt, int32_t substeps,
                           const double *factual, const uint8_t *codes, const int32_t *n_steps,
                           const double *static_feature, const double *theta0, double lam, double gtol,
                           int32_t max_iter, int32_t line_search, double *coefs_out, int32_t *status_out, double *fval_out,
                           void *stream);
B200I_API int b200i_stlsq_prefix(int64_t n, int32_t T, int32_t fit_offset, double fd_dt, const double *factual,
                           const uint8_t *codes, const int32_t *n_steps, const double *static_feature,
                           const double *prior, double support_tol, double lam, double threshold, int32_t max_iter,
                           double *coefs_out, void *stream);


/* ------------------------------------------------------------------------------------------------
 * Irregular sampling (BASELINE config C4).  The reference's integrator takes any time grid -- odeint(func, y0, t),
 * pkpd/utils.py:68-90: dts = diff(t), every interval split into STEPS_FOR_DT Euler sub-steps of dts[k] / STEPS_FOR_DT
 * when hmax < dts[0] -- but its live path always feeds the uniform STANDARD_DT (sindy.py:90, :395, :565;
 * FiniteDifference(is_uniform=True), :195).  These entry points take the interval lengths explicitly:
 *   dts[k] = t[k+1] - t[k], the time between column k and column k+1 of a row; (W,) shared by all rows
 *   (dts_per_row == 0) or (R, W) per row (dts_per_row == 1), W as in the uniform entry point of the same name
 *   (theta_gram: T-1 intervals between the T volumes).
 * Rollouts step with h = dts[k] / substeps, fits differentiate with (x[k+1] - x[k]) / dts[k] (pysindy
 * FiniteDifference(order=1) on a time array).  A dts array filled with the uniform dt reproduces the uniform entry
 * points bit for bit.  Known answers: the reference's in-file odeint tests (dy/dt = 1 => y = t on a dense and on a
 * two-point grid, pkpd/utils.py:759-828).
 * b200i_stlsq_batched_dts also offers FP32 STORAGE of the volumes (x_f32 instead of x; arithmetic stays FP64,
 * SURVEY.md App. E.5): pass exactly one of x / x_f32; dts may be NULL (uniform fd_dt).
 * estimator: 0 = ridge-to-prior (b200i_stlsq_batched); 1 = the ridge / threshold loop of the reference's dormant
 * per-patient optimiser LSQIntialMask (pkpd/utils.py:244-327, used at pkpd_simulation.py:778-836): `prior` only
 * supplies the initial support |prior| > support_tol (the reference: 1e-14, :251-253), every ridge step solves
 * (Theta^T Theta + lam I) c = Theta^T xdot on the current support (sklearn ridge_regression, :228; lam = alpha), then
 * thresholds.  That is the reference's estimator WITHOUT pysindy's unbias refit -- the branch pkpd_simulation.py:795-797
 * keeps when the unbiased coefficients overflow; the refit itself is an OLS on a per-patient design that is rank
 * deficient by construction (the static feature is constant within a patient) and is not offered.
 * ---------------------------------------------------------------------------------------------- */
B200I_API int b200i_ode_rollout_dts(int64_t rows, int32_t W, int32_t substeps, const double *x0,
                           const double *static_feature, const uint8_t *codes, const double *coefs,
                           int32_t coefs_per_row, double drop_below, const double *dts, int32_t dts_per_row,
                           int32_t fp32, double *pred, void *stream);
B200I_API int b200i_stlsq_batched_dts(int64_t rows, int32_t W, const double *x, const float *x_f32, const uint8_t *codes,
                           const int32_t *fit_len, const double *static_feature, const double *prior,
                           double support_tol, double lam, double threshold, int32_t max_iter, double fd_dt,
                           const double *dts, int32_t dts_per_row, int32_t estimator, double *coefs_out, void *stream);
B200I_API int b200i_insite_bfgs_dts(int64_t rows, int32_t W, int32_t substeps, const double *x, const uint8_t *codes,
                           const int32_t *sequence_lengths, int32_t projection_horizon, const double *static_feature,
                           const double *theta0, double lam, double gtol, int32_t max_iter, int32_t line_search,
                           const double *dts, int32_t dts_per_row, double *coefs_out, int32_t *status_out, double *fval_out,
                           void *stream);
B200I_API int b200i_theta_gram_dts(int64_t n, int32_t T, const double *cancer_volume, const double *chemo_application,
                           const double *radio_application, const double *sequence_lengths,
                           const double *static_feature, const double *dts, int32_t dts_per_row,
                           void *gram_workspace, void *stream);

/* ------------------------------------------------------------------------------------------------
 * model.use_smoothed_finite_difference (sindy.py:196-198): pysindy SmoothedFiniteDifference with scipy's
 * savgol_filter(window_length=2, polyorder=1) run over every fitting trajectory before the order-1 finite difference.
 * With an even window the filter is the half-sample two-point mean: interior samples become (x[i] + x[i+1]) / 2, the
 * first and last sample of a trajectory keep their value (mode='interp' refits them through two points).  Trajectories
 * as b200i_theta_gram cuts them (joint = 0: constant-treatment snippets, which share their end sample with the next
 * snippet's first sample -- both edges; joint = 1: columns 1..L of a patient).  smoothed_out (N,T) then replaces
 * cancer_volume in b200i_theta_gram / _mode / _dts.  Not in-place.
 * ---------------------------------------------------------------------------------------------- */
B200I_API int b200i_smooth_snippets(int64_t n, int32_t T, const double *cancer_volume, const double *chemo_application,
                           const double *radio_application, const double *sequence_lengths, int32_t joint,
                           double *smoothed_out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * model.ablation_more_complex_basis_functions (sindy.py:185-186): PolynomialLibrary(degree=4, interaction_only=False)
 * over [x0 = volume, u0 = patient type] for the four per-treatment models -- 15 monomials x0^a u0^b in sklearn's
 * order: 1, x0, u0, x0^2, x0 u0, u0^2, x0^3, x0^2 u0, x0 u0^2, u0^3, x0^4, x0^3 u0, x0^2 u0^2, x0 u0^3, u0^4.
 * The library is rank deficient on every cancer_sim cohort (three patient types) and spans 13 orders of magnitude, so
 * the statistics are the R factors of [Theta | xdot] (tall-skinny QR by Givens rotations), not normal equations, and
 * the STLSQ passes solve through a one-sided Jacobi SVD of R[:, support]: ridge (Theta^T Theta + alpha I)^-1 Theta^T xdot
 * as sklearn's ridge_regression, the final un-biasing as scipy.linalg.lstsq (minimum norm, singular values below
 * rcond * s_max dropped; rcond <= 0 selects eps * max(samples, 15), numpy.linalg.lstsq's default.  On cancer_sim data
 * the genuine singular values of a support stop at ~3e-4 * s_max and the three null directions sit at ~1e-17 * s_max,
 * so every cut-off between scipy's own default of the reference's time -- machine epsilon, a factor 4 above that
 * noise -- and today's scikit-learn (cond = 1e-6) gives the same solution).
 *
 * b200i_poly_tsqr: trajectories, finite differences and sample rows exactly as b200i_theta_gram (dense (N,T) rows).
 *   r_out: 4 * 256 + 4 doubles = the four row-major 16x16 upper-triangular factors (columns 0..14 the monomials, 15 the
 *   derivative) followed by the four sample counts.  workspace: b200i_poly_workspace_bytes().
 * b200i_poly_stlsq: r_out -> coefs (4,15) float64, support (4,15) int32 (pysindy STLSQ + unbias, pkpd/utils.py:274-310).
 * b200i_poly_rollout: b200i_ode_rollout for dx/dt = sum_j c[code][j] x^a_j u^b_j (terms with |c| <= drop_below dropped,
 *   pkpd/utils.py:388), explicit Euler with `substeps` per interval (pkpd/utils.py:40,68-90); W <= 128.
 * ---------------------------------------------------------------------------------------------- */
#define B200I_POLY_TERMS 15
#define B200I_POLY_R_DOUBLES (4 * 256 + 4)
B200I_API int64_t b200i_poly_workspace_bytes(void);
B200I_API int b200i_poly_tsqr(int64_t n, int32_t T, double fd_dt, const double *cancer_volume,
                     const double *chemo_application, const double *radio_application, const double *sequence_lengths,
                     const double *static_feature, void *workspace, double *r_out, void *stream);
B200I_API int b200i_poly_stlsq(const double *r_factors, double threshold, double alpha, int32_t max_iter, double rcond,
                      double *coefs_out, int32_t *support_out, void *stream);
B200I_API int b200i_poly_rollout(int64_t rows, int32_t W, double dt, int32_t substeps, const double *x0,
                        const double *static_feature, const uint8_t *codes, const double *coefs, double drop_below,
                        double *pred, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200I_H */
