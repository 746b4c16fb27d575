/*
 * larnd_b200 — C ABI of the B200-native (sm_100a) larnd-sim hot path.
 *
 * The reference (pgranger23/larnd-sim-jax) has no FFI boundary: its boundary is the set of
 * Python/JAX functions of src/larndsim/sim_jax.py.  Each entry point below replaces the XLA
 * lowering of one of them and is what a jax.ffi custom-call handler (or the ctypes binding
 * of larndsim_b200) binds to:
 *
 *   larnd_lut_create        <- load_lut's device-side products + the per-call
 *                              jnp.cumsum(response_template)          consts_jax.py:387-449, sim_jax.py:228
 *   larnd_lut_forward       <- simulate_wfs = simulate_drift_new + unique/renumber
 *                              + simulate_signals                      sim_jax.py:689-736 (375-453, 717-725, 142-286)
 *   larnd_lut_backward      <- jax.grad through simulate_wfs           (VJP w.r.t. the fitted Params leaves)
 *   larnd_fee_forward       <- simulate_stochastic = get_adc_values + digitize + id2pixel
 *                              + get_pixel_coordinates + get_hit_z + parse_output
 *                                                                      sim_jax.py:738-769, fee_jax.py:57-71,170-279
 *   larnd_fee_backward      <- jax.grad through get_adc_values/digitize (VJP w.r.t. the waveforms)
 *   larnd_mc_forward        <- simulate_drift(mc_diff) + current_mc + accumulate_signals_parametrized
 *                                                                      sim_jax.py:122-139,289-335; detsim_jax.py:207-228,618-639
 *   larnd_mc_backward       <- jax.grad through the same
 *   larnd_tracks_stage      <- shift_tracks / quench / drift as separate stages
 *                                                                      sim_jax.py:109-119, quenching_jax.py:38-75, drifting_jax.py:19-58
 *   larnd_signals_stream_*  <- simulate_signals with its own (materialised-stream) argument list and its VJP
 *                                                                      sim_jax.py:142-286
 *   larnd_current_mc[_backward], larnd_accumulate_parametrized[_backward]
 *                           <- current_mc, accumulate_signals_parametrized        detsim_jax.py:619-639, 209-228
 *
 * Conventions: plain pointers and sizes, no allocation and no host synchronisation inside
 * (workspace is caller-provided, its size queried with larnd_workspace_bytes); every pointer
 * named *_d is DEVICE memory; `stream` is a cudaStream_t passed as void*; functions return 0 on
 * success or a negative LARND_E_* code, with a message available from larnd_last_error().
 * All arithmetic is float32; ids are int32 (the reference never enables x64).
 */
#ifndef LARND_B200_H
#define LARND_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LARND_ABI_VERSION 2
#define LARND_MAX_TPC 8
#define LARND_MAX_TEMPLATES 128
#define LARND_NB_TRAN_BINS 5   /* params.nb_tran_diff_bins, consts_jax.py:158,294 */
#define LARND_MAX_ADC 10       /* params.MAX_ADC_VALUES,    consts_jax.py:279   */

enum {
  LARND_OK = 0,
  LARND_E_ARG = -1,       /* invalid argument / unsupported configuration */
  LARND_E_CUDA = -2,      /* CUDA runtime error (message has the detail) */
  LARND_E_CAPACITY = -3   /* caller-provided capacity too small (host-side check) */
};

/* `flags` of larnd_lut_forward / _accumulate / _backward (every behaviour switch is an explicit argument: the library
 * reads no environment variable). */
#define LARND_FLAG_SKIP_GARBAGE 0x1  /* drop what the reference routes to the garbage row 0 (NOT reference-identical for wfs[0]) */
#define LARND_FLAG_IMPL_CHUNK   0x2  /* force the chunk kernels (default: by batch size)            */
#define LARND_FLAG_IMPL_SORTED  0x4  /* force the class-sorted tile kernels where they are supported  */
#define LARND_FLAG_NO_SPLIT     0x8  /* class-sorted kernels: serve every tile with the 6-position variant */
#define LARND_FLAG_REUSE_RUNS   0x10 /* larnd_lut_backward / _backward_steps: the workspace still holds the sorted run / tile tables the
                                      * forward call (larnd_lut_accumulate / _forward) built from the same records, on the same
                                      * stream or synchronised with it; the library checks (workspace, n, lut, n_ticks) against its
                                      * record of the last build and rebuilds when they do not match */
#define LARND_FLAG_WFS_ZERO     0x20 /* larnd_lut_accumulate / _forward: wfs_d is already all-zero (fresh cudaMemset, or cleaned by
                                      * larnd_fee_forward_ex(LARND_FEE_CLEAR_WFS)): skip the memset of the waveform buffer */
#define LARND_FLAG_PUBLIC_MASK  0x3f

/* Differentiable parameters (optimize/ranges.py:7-21).  Gradients are returned in this order. */
enum {
  LARND_P_AB = 0, LARND_P_KB, LARND_P_EFIELD, LARND_P_LIFETIME, LARND_P_LONG_DIFF, LARND_P_TRAN_DIFF,
  LARND_P_SHIFT_X, LARND_P_SHIFT_Y, LARND_P_SHIFT_Z, LARND_P_ALPHA, LARND_P_BETA, LARND_P_R_PARAM,
  LARND_P_LAR_DENSITY, LARND_P_MEV_TO_ELECTRONS, LARND_P_VDRIFT /* never read by the reference: grad == 0 */,
  LARND_NPARAMS
};

/* Column indices of the (N, ncols) float32 `tracks` array (row-major, reference `fields` tuple). */
typedef struct larnd_columns {
  int32_t ncols;
  int32_t eventID, x, y, z, z_start, z_end, dx, dEdx, dE, t0;
} larnd_columns_t;

/* Simulation constants, already rounded the way the reference's weak-typed Python floats are
 * (one f32 rounding of the double-precision constant expression).  Mirrors the fields of
 * Params_template that the hot path reads (consts_jax.py:80-160). */
typedef struct larnd_params {
  /* recombination (quenching_jax.py:18-75) */
  int32_t recombination_mode;   /* 1 BOX, 2 BIRKS, 3 ELLIPSOID */
  float Ab, kb, alpha, beta, inv_R2 /* 1/R_param^2 */, efield_rho /* eField*lArDensity */, MeVToElectrons;
  /* drift (drifting_jax.py:19-58) */
  float vdrift, lifetime, long_diff, tran_diff, size_margin;
  float shift_x, shift_y, shift_z;
  int32_t n_tpc;
  float tpc_borders[LARND_MAX_TPC][3][2];
  /* pixelisation (detsim_jax.py:232-244,494-512; sim_jax.py:398-450) */
  float pixel_pitch, bin_width /* pitch / nb_sampling_bins_per_pixel */, half_pitch;
  int32_t nb_sampling_bins_per_pixel, n_pixels_x, n_pixels_y, number_pix_neighbors;
  float tran_bin_edges[LARND_NB_TRAN_BINS + 1];  /* jnp.linspace(-2w, 3w, 6) */
  /* time axis (sim_jax.py:147-178) */
  float t_sampling;
  int32_t n_ticks;        /* int(time_interval[1]/t_sampling)+1, includes the garbage tick 0 */
  int32_t signal_length;  /* L */
  int32_t n_templates;    /* len(long_diff_template) */
  float long_diff_template[LARND_MAX_TEMPLATES];
  /* front end (fee_jax.py:57-71,170-279) */
  float discrimination_threshold, reset_noise_charge, uncorrelated_noise_charge;
  float gain, v_cm, v_ref_minus_cm, v_pedestal, adc_counts, hit_prob_threshold;
  int32_t hold_interval;  /* round((3+ADC_HOLD_DELAY)*CLOCK_CYCLE/t_sampling) */
  int32_t max_adc_values;
  /* MC-current mode (detsim_jax.py:618-639) */
  int32_t diffusion_in_current_sim;
  /* derivative helpers computed on the host in double precision */
  float dvdrift_dEfield;  /* d get_vdrift / d eField */
  float eField, lArDensity, R_param;
  float ts_vdrift;        /* float32(t_sampling) * float32(vdrift): XLA folds the two scalars of get_hit_z (detsim_jax.py:318) */
} larnd_params_t;

typedef struct larnd_lut larnd_lut_t; /* opaque: compacted response rows + cumulative tables on the device */

const char* larnd_last_error(void);
int larnd_abi_version(void);

/* Template bank of load_lut (consts_jax.py:427-447): bank_d (n_templates, nx, ny, nt) float32 with
 * bank[t] = response (*) gauss[t] ('same' convolution along time, zero padded) for t >= 1 and bank[0] = response.
 * gauss_d: (n_templates, taps) normalised Gaussians, taps odd (2*long_diff_extent + 1). */
int larnd_build_bank(const float* response_d, int nx, int ny, int nt, const float* gauss_d, int n_templates, int taps,
                     float* bank_d, void* stream);

/* Builds the device tables for one (response_template, signal_length) pair:
 *   neighbour rows  R0[ci][cj][Nt-L..Nt)          (template 0, all Nx*Ny bins)
 *   main rows       Rm[tpl][ci<5][cj<5][Nt-L..Nt) (all templates, the 5x5 collecting bins)
 *   cumulative sums C0[ci][cj][0..Nt), Cm[tpl][ci<5][cj<5][0..Nt)  == jnp.cumsum(response_template,-1)
 * `bank_d` is the (n_templates, nx, ny, nt) float32 template bank on the device (load_lut's output);
 * it is only read during this call. */
int larnd_lut_create(const float* bank_d, int n_templates, int nx, int ny, int nt, int signal_length,
                     void* stream, larnd_lut_t** out);
/* Second half of the table build, needed before any accumulate / backward call: the neighbourhood-sum rows for
 * (nb_sampling_bins_per_pixel, number_pix_neighbors) — everything a segment deposits on neighbour pixels that are not main
 * pixels lands in waveform row 0 (sim_jax.py:724-725), by linearity (sum over the whole neighbourhood) - (owned ones).
 * This is the only call that allocates / mutates a LUT after creation; the per-batch entry points take a const handle,
 * never allocate, and fail with LARND_E_ARG when the tables do not match their parameters.  Not concurrent with
 * per-batch calls on the same handle. */
int larnd_lut_prepare_neighbours(larnd_lut_t* lut, int nb_sampling_bins_per_pixel, int number_pix_neighbors, void* stream);
void larnd_lut_destroy(larnd_lut_t* lut);

/* Workspace (device) size for a batch of n_segments whose local event ids are < n_events
 * (padding rows carry eventID -1) — both larnd_lut_* and larnd_mc_* use it. */
size_t larnd_workspace_bytes(int64_t n_segments, int32_t n_events, int32_t n_tpc, int32_t n_pixels_x, int32_t n_pixels_y);

/* simulate_wfs.  Outputs:
 *   unique_pixels_d (npix_capacity) int32 : sorted unique main-pixel ids, front-padded with -1 exactly
 *                                           like sim_jax.py:717-721 for a padded size of npix_capacity
 *   wfs_d (npix_capacity, n_ticks) float32: FULL waveform rows including garbage column 0
 *                                           (simulate_signals' return; simulate_wfs is wfs[:,1:]), row stride
 *                                           wfs_row_stride floats (>= n_ticks).  With a 16-byte aligned base and a
 *                                           stride that is a multiple of 4 and >= n_ticks + 3 (2004 for 2001 ticks) the
 *                                           tile kernel flushes whole frames with red.global.add.v4.f32; any other
 *                                           stride (e.g. n_ticks itself) is served with scalar reductions.  The
 *                                           padding columns are zeroed.
 *   counts_d[4] int32                     : {n_unique_main_pixels, n_negative_ids, overflow_flag, n_chunks}
 * overflow_flag != 0 means npix_capacity < n_unique+1: outputs are then invalid.
 * flags: LARND_FLAG_* above. */
int larnd_lut_forward(const float* tracks_d, int64_t n_segments, const larnd_columns_t* cols,
                      const larnd_params_t* params, const larnd_lut_t* lut, int32_t n_events,
                      int32_t npix_capacity, int32_t flags, void* workspace_d, size_t workspace_bytes,
                      int32_t* unique_pixels_d, float* wfs_d, int64_t wfs_row_stride, int32_t* counts_d, void* stream);

/* Deterministic simulate_wfs (the counterpart of running the reference with --xla_gpu_deterministic_ops,
 * optimize/example_run.py:44-47): same outputs as larnd_lut_forward, but bitwise reproducible from run to run.  The default
 * kernels add float32 window sums with red.global.add in arrival order (low bits vary, ~1e-7 relative); here every window
 * sum — a fixed function of its chunk of segments — is added as a 64-bit fixed-point integer (2^-20 resolution), which is
 * order-independent, and converted to float32 once.  Uses the chunk kernels for every batch size (slower at spill size)
 * and a caller-provided scratch of larnd_deterministic_scratch_bytes(npix_capacity, n_ticks) bytes. */
size_t larnd_deterministic_scratch_bytes(int32_t npix_capacity, int32_t n_ticks);
int larnd_lut_accumulate_deterministic(int64_t n_segments, const larnd_params_t* params, const larnd_lut_t* lut,
                                       int32_t n_events, int32_t npix_capacity, int32_t flags, void* workspace_d,
                                       size_t workspace_bytes, int32_t* unique_pixels_d, float* wfs_d, int64_t wfs_row_stride,
                                       int32_t* counts_d, void* det_scratch_d, size_t det_scratch_bytes, void* stream);
int larnd_lut_forward_deterministic(const float* tracks_d, int64_t n_segments, const larnd_columns_t* cols,
                                    const larnd_params_t* params, const larnd_lut_t* lut, int32_t n_events,
                                    int32_t npix_capacity, int32_t flags, void* workspace_d, size_t workspace_bytes,
                                    int32_t* unique_pixels_d, float* wfs_d, int64_t wfs_row_stride, int32_t* counts_d,
                                    void* det_scratch_d, size_t det_scratch_bytes, void* stream);

/* Only the drift/pixelisation stage + unique/renumber (simulate_drift_new + sim_jax.py:717-725):
 * fills the workspace segment records and unique_pixels_d/counts_d.  Lets a caller size wfs exactly
 * (the reference's pad_size(n_unique+1)) before calling larnd_lut_accumulate. */
int larnd_lut_prepare(const float* tracks_d, int64_t n_segments, const larnd_columns_t* cols,
                      const larnd_params_t* params, const larnd_lut_t* lut, int32_t n_events,
                      void* workspace_d, size_t workspace_bytes, int32_t* counts_d, void* stream);
int larnd_lut_accumulate(int64_t n_segments, const larnd_params_t* params, const larnd_lut_t* lut,
                         int32_t n_events, int32_t npix_capacity, int32_t flags, void* workspace_d,
                         size_t workspace_bytes, int32_t* unique_pixels_d, float* wfs_d, int64_t wfs_row_stride,
                         int32_t* counts_d, void* stream);

/* VJP of simulate_wfs w.r.t. the LARND_NPARAMS fitted parameters.  Must follow a forward/prepare call
 * on the same workspace.  g_wfs_d is (npix_capacity, n_ticks) with row stride g_row_stride floats
 * (column 0 = garbage tick, never read: a gradient of simulate_wfs' (npix, n_ticks-1) output is passed as
 * g - 1 with stride n_ticks-1).  counts_d = the forward call's counts.  flags bit0: gradients of rows the
 * forward treats as garbage (id < 0 or not a main pixel) are taken as zero and skipped.
 * grad_params_d[LARND_NPARAMS] is ACCUMULATED into (caller zeroes). */
int larnd_lut_backward(int64_t n_segments, const larnd_params_t* params, const larnd_lut_t* lut,
                       int32_t n_events, int32_t npix_capacity, int32_t flags, void* workspace_d,
                       size_t workspace_bytes, const int32_t* counts_d, const float* g_wfs_d, int64_t g_row_stride,
                       float* grad_params_d, void* stream);

/* The two VJPs above chained WITHOUT the dense waveform gradient.  The VJP of get_adc_values is a step function of the tick
 * (at most 2 * MAX_ADC_VALUES steps per pixel row: a hit's integral reaches back to the previous subtraction), so
 * larnd_fee_backward_steps emits the (last column, coefficient) list of every row — 168 bytes instead of n_ticks floats —
 * and larnd_lut_backward_steps turns every correlation sum_k g[t + k] R[k] into differences of the response's running sum at
 * the step positions.  Same gradients as larnd_fee_backward + larnd_lut_backward (rows of pixel ids < 0 carry no gradient, as
 * for every loss built on parse_output's hits); the 2 GB gradient array of a spill-sized batch is neither written nor read.
 * steps_d: larnd_fee_steps_bytes(npix) bytes, npix = the npix_capacity of the forward call. */
size_t larnd_fee_steps_bytes(int32_t npix);
int larnd_fee_backward_steps(const float* g_adc_d, const float* saved_d, const int32_t* unique_pixels_d, int32_t npix,
                             const larnd_params_t* params, void* steps_d, size_t steps_bytes, int32_t raw_charge, void* stream);
int larnd_lut_backward_steps(int64_t n_segments, const larnd_params_t* params, const larnd_lut_t* lut, int32_t n_events,
                             int32_t npix_capacity, int32_t flags, void* workspace_d, size_t workspace_bytes,
                             const int32_t* counts_d, const void* steps_d, size_t steps_bytes, float* grad_params_d, void* stream);

/* simulate_stochastic on (npix, n_ticks-1) waveforms (row stride wfs_row_stride floats, first column =
 * reference tick index 0 of wfs[:,1:]).  noise_d: NULL (noise-free) or standard normals laid out as
 * [base(npix) | extra(10,npix) | pass(10,npix) | fail(10,npix)].
 * Dense outputs (npix,10): adc_d (digitized ADC), ticks_d (float, integer valued), pixel_z_d.
 * Per-pixel outputs (npix): pixel_x_d, pixel_y_d, event_d (int32).
 * Compacted outputs (capacity npix*10, first n_valid entries meaningful, parse_output order):
 *   hit_adc_d, hit_x_d, hit_y_d, hit_z_d, hit_ticks_d, hit_prob_d (float), hit_event_d, hit_pixel_d (int32),
 *   n_valid_d[1].  saved_d (npix, 32) float: per-row state needed by larnd_fee_backward ([0..9] crossing ticks, [10..12]
 *   hit / subtraction / digitiser-slope masks, [13..22] the integrated charge of every hit BEFORE the digitiser = the
 *   `adc` array get_adc_values itself returns, fee_jax.py:170-279).
 * Row loads: when the row stride is a multiple of four floats, the rows are fetched with one cp.async.bulk (TMA) each from
 * the 16-byte aligned address at or below the row start; if wfs_d itself is k = (wfs_d mod 16) / 4 floats past a 16-byte
 * boundary (k = 1 for simulate_wfs' view [:, 1:] of the padded buffer) the k floats in front of every row and the round-up
 * to a whole vector behind it are READ as well (never used) and must belong to the caller's buffer — true for any row
 * stride >= round_up(k + n_ticks - 1, 4).  Other layouts take the vector-load path. */
int larnd_fee_forward(const float* wfs_d, int64_t wfs_row_stride, const int32_t* unique_pixels_d, int32_t npix,
                      const larnd_params_t* params, const float* noise_d,
                      float* adc_d, float* ticks_d, float* pixel_z_d, float* pixel_x_d, float* pixel_y_d,
                      int32_t* event_d, float* saved_d,
                      float* hit_adc_d, float* hit_x_d, float* hit_y_d, float* hit_z_d, float* hit_ticks_d,
                      float* hit_prob_d, int32_t* hit_event_d, int32_t* hit_pixel_d, int32_t* n_valid_d,
                      void* scratch_d, size_t scratch_bytes, void* stream);
/* larnd_fee_forward with flags.  LARND_FEE_CLEAR_WFS: the waveform buffer is scratch of a hits-only pipeline — every row's
 * non-zero samples (and its garbage column) are zeroed as soon as the row has been read, so the buffer is all-zero again when
 * the call completes and the next larnd_lut_accumulate on it may pass LARND_FLAG_WFS_ZERO instead of paying its 2 GB memset
 * (8 KB per row written vs ~0.6 KB).  Needs the TMA row layout described above. */
#define LARND_FEE_CLEAR_WFS 0x1
int larnd_fee_forward_ex(const float* wfs_d, int64_t wfs_row_stride, const int32_t* unique_pixels_d, int32_t npix,
                      const larnd_params_t* params, const float* noise_d,
                      float* adc_d, float* ticks_d, float* pixel_z_d, float* pixel_x_d, float* pixel_y_d,
                      int32_t* event_d, float* saved_d,
                      float* hit_adc_d, float* hit_x_d, float* hit_y_d, float* hit_z_d, float* hit_ticks_d,
                      float* hit_prob_d, int32_t* hit_event_d, int32_t* hit_pixel_d, int32_t* n_valid_d,
                      void* scratch_d, size_t scratch_bytes, int32_t fee_flags, void* stream);
size_t larnd_fee_scratch_bytes(int32_t npix);

/* VJP of get_adc_values+digitize: g_adc_d (npix,10) -> g_wfs_d (npix, n_ticks-1) with row stride.
 * raw_charge != 0: g_adc_d is the gradient w.r.t. the INTEGRATED CHARGE get_adc_values returns (saved_d[row*32 + 13 + k],
 * fee_jax.py:229-265) instead of the digitised ADC, i.e. the digitiser slope and its clipping are left out. */
int larnd_fee_backward(const float* g_adc_d, const float* ticks_d, const float* saved_d, int32_t npix,
                       const larnd_params_t* params, float* g_wfs_d, int64_t g_row_stride, int32_t raw_charge, void* stream);

/* Noise-averaged ("probabilistic") front end, forward: get_adc_values_average_noise_vmap (fee_jax.py:390-461).
 * wfs_d: (npix, n_ticks) waveforms (row stride in floats).  Outputs (npix, MAX_ADC_VALUES, n_ticks-1) each:
 *   log_prob_d : log-probability that hit number k of the pixel triggers at tick t (log_total_hit_dist_tick)
 *   charge_d   : expected integrated charge of such a hit (esperance_value; digitize() maps it to ADC)
 *   top_ticks_d: optional (npix, MAX_ADC_VALUES, n_paths) int32, the ticks kept by the beam search (may be NULL)
 * n_paths = params.fee_paths_scaling (20); sigma = params->reset_noise_charge must be > 0. */
size_t larnd_prob_fee_scratch_bytes(int32_t npix, int32_t n_ticks, int32_t n_paths, int32_t n_steps);
/* state_d: optional (npix, MAX_ADC_VALUES, 2*n_paths) float, the path charges and log-probabilities BEFORE every step;
 * flags_d: optional (MAX_ADC_VALUES+1) int32, whether step k ran (the reference stops all pixels globally).  Both, and
 * top_ticks_d, are what larnd_prob_fee_backward needs. */
int larnd_prob_fee_forward(const float* wfs_d, int64_t wfs_row_stride, int32_t npix, int32_t n_ticks,
                           const larnd_params_t* params, int32_t n_paths, float stop_threshold, float* log_prob_d,
                           float* charge_d, int32_t* top_ticks_d, float* state_d, int32_t* flags_d, void* scratch_d,
                           size_t scratch_bytes, void* stream);
/* VJP of the above w.r.t. the waveforms (beam ticks and stop flags are the forward's, like lax.stop_gradient in
 * fee_jax.py:380): g_log_prob_d, g_charge_d (npix, MAX_ADC_VALUES, n_ticks-1) -> g_wfs_d (npix, n_ticks). */
size_t larnd_prob_fee_bwd_scratch_bytes(int32_t npix, int32_t n_ticks, int32_t n_paths);
int larnd_prob_fee_backward(const float* wfs_d, int64_t wfs_row_stride, int32_t npix, int32_t n_ticks,
                            const larnd_params_t* params, int32_t n_paths, const float* g_log_prob_d, const float* g_charge_d,
                            const float* log_prob_d, const float* state_d, const int32_t* top_ticks_d, const int32_t* flags_d,
                            float* g_wfs_d, int64_t g_row_stride, void* scratch_d, size_t scratch_bytes, void* stream);

/* MC-current mode with number_pix_neighbors = 0 and mc_diff = True.  rnd_d: (N,3) standard normals
 * (the reference draws random.normal(key,(N,3)), detsim_jax.py:393).  Same output convention as
 * larnd_lut_forward. */
int larnd_mc_forward(const float* tracks_d, int64_t n_segments, const larnd_columns_t* cols,
                     const larnd_params_t* params, const float* rnd_d, int32_t n_events, int32_t npix_capacity,
                     void* workspace_d, size_t workspace_bytes, int32_t* unique_pixels_d, float* wfs_d,
                     int32_t* counts_d, void* stream);
int larnd_mc_backward(const float* tracks_d, int64_t n_segments, const larnd_columns_t* cols,
                      const larnd_params_t* params, const float* rnd_d, int32_t n_events, int32_t npix_capacity,
                      void* workspace_d, size_t workspace_bytes, const int32_t* counts_d, const float* g_wfs_d,
                      int64_t g_row_stride, float* grad_params_d, void* stream);

/* Device-side chop_tracks (replaces optimize/dataio.py:63-106, a Python loop per raw row on the host).
 * raw_d: (m, ncols) float32 raw segments in the file's column order; every raw row becomes
 * max(ceil(length / precision), 1) rows of the same width with start/end/mid points, dx and dE subdivided exactly as
 * numpy evaluates the reference expressions (float64 for steps*precision*direction, float32 elsewhere).
 *   larnd_chop_count : offsets_d[0..m) = exclusive prefix of the piece counts, offsets_d[m] = total (int64, device)
 *   larnd_chop_tracks: writes the total x ncols output; does nothing if the total exceeds `capacity` rows
 *                      (the caller reads offsets_d[m] to size / validate the output). */
typedef struct larnd_chop_columns {
  int32_t ncols;
  int32_t x, y, z, x_start, y_start, z_start, x_end, y_end, z_end, dx, dE;
} larnd_chop_columns_t;
int larnd_chop_count(const float* raw_d, int64_t m, const larnd_chop_columns_t* cols, double precision,
                     int64_t* offsets_d, void* stream);
int larnd_chop_tracks(const float* raw_d, int64_t m, const larnd_chop_columns_t* cols, double precision,
                      const int64_t* offsets_d, float* out_d, int64_t capacity, void* stream);

/* larnd_lut_prepare for RAW (un-chopped) rows: chop_tracks fused into the drift / pixelisation stage.  Segment s of the batch
 * (0 <= s < n_segments, the caller's capacity; workspace sized for n_segments) is piece s - offsets_d[i] of raw row i, its
 * ten simulation columns formed in registers with chop_tracks' arithmetic (bit-identical to larnd_chop_tracks), so the
 * chopped (n, ncols) batch of optimize/dataio.py:63-106 — 104 B per segment written and read back — never exists.
 * Segments beyond offsets_d[m] are the invalid rows pad_batch appends (:340-373).  offsets_d: larnd_chop_count's output for
 * the same rows and precision.  offsets_d[m] > n_segments sets bit 3 of counts_d[2] (outputs invalid, kernels bail out).
 * Follow with larnd_lut_accumulate(n_segments, ...) / the backward entry points as after larnd_lut_prepare. */
int larnd_lut_prepare_raw(const float* raw_d, int64_t m, const larnd_chop_columns_t* chop_cols, const larnd_columns_t* cols,
                          double precision, const int64_t* offsets_d, int64_t n_segments, const larnd_params_t* params,
                          const larnd_lut_t* lut, int32_t n_events, void* workspace_d, size_t workspace_bytes,
                          int32_t* counts_d, void* stream);

/* Batch assembly around the chop (TracksDataset.__getitem__ / pad_batch, optimize/dataio.py:340-406, :47-61): the file's
 * rows stay on the device; a batch is a list of row indices with the batch-local event id of every row.
 *   larnd_batch_gather: out_d (m, ncols) = raw_d[rows_d[i]] with column event_col replaced by local_event_d[i]
 *   larnd_pad_rows    : rows [*n_valid_d, capacity) of batch_d become invalid rows — eventID / trackID / pixel_plane -1,
 *                       every other column 0 (np.pad + _invalidate_rows); n_valid_d is a DEVICE scalar (e.g. the total
 *                       larnd_chop_count left in offsets_d[m]), so the batch is produced without a host synchronisation */
typedef struct larnd_pad_columns {
  int32_t eventID, trackID, pixel_plane;   /* column indices; -1 = column absent */
} larnd_pad_columns_t;
int larnd_batch_gather(const float* raw_d, int32_t ncols, const int64_t* rows_d, const int32_t* local_event_d, int64_t m,
                       int32_t event_col, float* out_d, void* stream);
int larnd_pad_rows(float* batch_d, int32_t ncols, const int64_t* n_valid_d, int64_t capacity, const larnd_pad_columns_t* cols,
                   void* stream);

/* jax.random-compatible random numbers (Threefry-2x32; replaces jax.random.key/split/normal at fee_jax.py:186,237-255,271,
 * detsim_jax.py:393, sim_jax.py:359-360,757).  key = the two uint32 words of jax.random.key(seed) = {seed >> 32, seed};
 * partitionable = 1 follows jax_threefry_partitionable (default since JAX 0.5.0), 0 the original counter layout.
 *   larnd_rng_split     : random.split(key, num) -> keys_out[num][2] (host memory, computed on the host)
 *   larnd_rng_normal    : random.normal(key, shape) for prod(shape) = n float32 values, row-major, into device memory
 *   larnd_rng_fee_noise : all standard normals get_adc_values draws, in larnd_fee_forward's noise layout
 *                         [base(npix) | extra(n_adc,npix) | pass(n_adc,npix) | fail(n_adc,npix)] */
int larnd_rng_split(const uint32_t key[2], int num, int partitionable, uint32_t* keys_out);
int larnd_rng_normal(const uint32_t key[2], int64_t n, int partitionable, float* out_d, void* stream);
int larnd_rng_fee_noise(const uint32_t key[2], int32_t npix, int32_t n_adc, int partitionable, float* noise_d, void* stream);

/* Weighted RBF field of a point set: for every target t (n_targets, 3)
 *   field[t] = { sum_j w_j K(t, z_j),  sum_j w_j K(t, z_j) (z_j - t) }   with K = exp(-|t - z|^2 / (2 sigma^2)),
 * the building block of the MMD loss (losses_jax.py:14-39) and of its gradient w.r.t. positions and weights.  Pairs
 * further apart than 15 sigma (exactly 0 in float32; e.g. hits of different events, offset by 1e5) are skipped per tile. */
size_t larnd_rbf_field_scratch_bytes(int32_t n_targets, int32_t n_sources);
int larnd_rbf_field(const float* targets_d, int32_t n_targets, const float* sources_d, const float* weights_d,
                    int32_t n_sources, float sigma, float* field_d /* (n_targets, 4) */, void* scratch_d, size_t scratch_bytes,
                    void* stream);

/* Dense mse_adc (losses_jax.py:58-82 on the hits of simulate_stochastic, sim_jax.py:738-769) WITHOUT hit compaction: the
 * inputs are larnd_fee_forward's dense (npix, MAX_ADC_VALUES) outputs, every slot parse_output would drop enters with weight
 * 0.  Lets a fit step (optimize/fit_params.py:731) run forward + loss + backward without a host synchronisation.
 *   larnd_mse_adc_sums     -> sums_d[0..2] = {Kxx, Kxy, Sx} of THIS rank's events (sums_d[3..4] = {Kyy, Sy} of the target are
 *                             written by the caller; all five are all-reduced by the caller when events are sharded)
 *   larnd_mse_adc_backward -> loss_d[4] = {loss, mmd term, charge term, Sx}, g_adc_d (npix, MAX_ADC_VALUES) = dL/d adc and
 *                             grad_params_d[LARND_P_EFIELD] += dL/d(hit z) * d(hit z)/d eField (detsim_jax.py:318)
 * ref_points_d (n_ref, 3) = (x + event * 1e5, y, z) and ref_weights_d = Q * hit_prob of the target hits.  The same
 * scratch_d (larnd_mse_adc_scratch_bytes) must be passed to both calls: it carries the kernel fields between them. */
size_t larnd_mse_adc_scratch_bytes(int32_t npix, int32_t max_adc_values, int32_t n_ref);
int larnd_mse_adc_sums(const float* adc_d, const float* ticks_d, const float* pixel_z_d, const float* pixel_x_d,
                       const float* pixel_y_d, const int32_t* event_d, const int32_t* unique_pixels_d, int32_t npix,
                       const larnd_params_t* params, const float* ref_points_d, const float* ref_weights_d, int32_t n_ref,
                       float sigma, float* sums_d, void* scratch_d, size_t scratch_bytes, void* stream);
int larnd_mse_adc_backward(const float* sums_d, const float* adc_d, const float* ticks_d, const float* pixel_z_d,
                           const float* pixel_x_d, const float* pixel_y_d, const int32_t* event_d,
                           const int32_t* unique_pixels_d, int32_t npix, const larnd_params_t* params, int32_t n_ref,
                           float sigma, float lambda_q, float* loss_d, float* g_adc_d, float* grad_params_d,
                           void* scratch_d, size_t scratch_bytes, void* stream);

/* ---- Stream-form operators: the reference functions whose arguments are the materialised per-segment arrays.  The
 * fused entry points above never build those arrays; these exist so that code calling the reference's stages one by one
 * (quench -> drift -> simulate_drift_new -> simulate_signals, or simulate_drift -> current_mc ->
 * accumulate_signals_parametrized) keeps working on the device. ---- */

/* shift_tracks (sim_jax.py:109-119), quench (quenching_jax.py:38-75), drift (drifting_jax.py:19-58): tracks (n, ncols)
 * in -> updated copy out (may alias).  stages: bit0 shift, bit1 quench, bit2 drift, applied in that order.  `cols` as for
 * larnd_lut_forward; `oc` names the extra columns the stages read/write. */
typedef struct larnd_track_columns {
  int32_t x_start, x_end, y_start, y_end;                                        /* shifted with x, y (z_start/z_end: `cols`) */
  int32_t n_electrons, long_diff, tran_diff, pixel_plane, t, t_start, t_end;    /* written by quench / drift */
} larnd_track_columns_t;
int larnd_tracks_stage(const float* tracks_d, int64_t n, const larnd_columns_t* cols, const larnd_track_columns_t* oc,
                       const larnd_params_t* params, int32_t stages, float* out_d, void* stream);

/* simulate_signals (sim_jax.py:142-286) with the reference's argument list:
 *   unique_pixels (npix) int32 sorted; main entries (n_main = N*T^2): pixels int32, t0_after_diff, nelectrons, long_diff
 *   float32, currents_idx (n_main,2) int32 in [0,5); neighbour entries: nelectrons_neigh (N), t0_neigh (N),
 *   pix_renumbering_neigh (N*P^2) int32 rows, currents_idx_neigh (N*P^2,2) int32.
 * wfs_d (npix, n_ticks) is zeroed and filled (garbage column 0 included).  The response values and running sums come from
 * the larnd_lut_t tables (== response_template / jnp.cumsum(response_template)).  status_d[0]: bit0 a main response index
 * was outside the collecting 5x5 bins, bit1 a neighbour index outside the LUT (those entries are skipped).
 * The backward call returns the VJP w.r.t. the five float streams (g_wfs_d: (npix, n_ticks) with row stride). */
int larnd_signals_stream_forward(const int32_t* unique_pixels_d, int32_t npix, const int32_t* pixels_d,
                                 const float* t0_after_diff_d, const float* nelectrons_d, const float* long_diff_d,
                                 const int32_t* currents_idx_d, int64_t n_main, const float* nelectrons_neigh_d,
                                 const int32_t* pix_renumbering_neigh_d, const float* t0_neigh_d,
                                 const int32_t* currents_idx_neigh_d, int64_t n_segments, const larnd_params_t* params,
                                 const larnd_lut_t* lut, float* wfs_d, int32_t* status_d, void* stream);
int larnd_signals_stream_backward(const int32_t* unique_pixels_d, int32_t npix, const int32_t* pixels_d,
                                  const float* t0_after_diff_d, const float* nelectrons_d, const float* long_diff_d,
                                  const int32_t* currents_idx_d, int64_t n_main, const float* nelectrons_neigh_d,
                                  const int32_t* pix_renumbering_neigh_d, const float* t0_neigh_d,
                                  const int32_t* currents_idx_neigh_d, int64_t n_segments, const larnd_params_t* params,
                                  const larnd_lut_t* lut, const float* g_wfs_d, int64_t g_row_stride,
                                  float* g_nelectrons_d, float* g_t0_after_diff_d, float* g_long_diff_d,
                                  float* g_nelectrons_neigh_d, float* g_t0_neigh_d, int32_t* status_d, void* stream);

/* Legacy entry points the reference keeps beside simulate_signals (no caller in the reference itself; signatures kept):
 *   simulate_signals_new   sim_jax.py:456-617      main entries (n_main = Nseg * 25) + neighbour ENTRIES
 *   accumulate_signals     detsim_jax.py:157-205   n_main = 0, entries = (currents_idx, charge, pixID, cathode_ticks)
 * Truncating tick (t0 / t_sampling).astype(int), no sub-tick split, bare searchsorted for the main pixels, boundary
 * correction of the main entries from the running sum of template 0 (the reference's response_cum.take without template
 * offset), .at[] index semantics (negative flat index wraps once, out of range dropped).  The per-entry arrays are the
 * reference's own: charge / pixID / cathode_ticks per (segment, neighbour pixel).  wfs_d (npix, n_ticks) is ACCUMULATED
 * INTO.  status_d[0]: bits 0-2 as larnd_signals_stream_forward, bit 3 a tick outside [0, Nt) (clamped for the correction). */
int larnd_signals_legacy_forward(const int32_t* unique_pixels_d, int32_t npix, const int32_t* pixels_d,
                                 const float* t0_after_diff_d, const float* nelectrons_d, const float* long_diff_d,
                                 const int32_t* currents_idx_d, int64_t n_main, const float* charge_entries_d,
                                 const int32_t* pix_id_entries_d, const int32_t* cathode_ticks_entries_d,
                                 const int32_t* currents_idx_entries_d, int64_t n_entries, const larnd_params_t* params,
                                 const larnd_lut_t* lut, float* wfs_d, int32_t* status_d, void* stream);

/* current_lut (detsim_jax.py:642-660): t0 = response_full_drift_t - electrons[:, t]; currents_idx = clip(int(|x - px| /
 * response_bin_size), 0, nx-1), likewise y.  electrons (n, ncols), pixels_coord (n, 2) -> t0 (n), currents_idx (n, 2). */
int larnd_current_lut(const float* electrons_d, int64_t n, int32_t ncols, int32_t col_x, int32_t col_y, int32_t col_t,
                      const float* pixels_coord_d, float response_full_drift_t, float response_bin_size, int32_t nx,
                      int32_t ny, float* t0_d, int32_t* currents_idx_d, void* stream);

/* current_mc (detsim_jax.py:619-639): electrons (n, ncols) + pixel centres (n, 2) -> t0_tick (n) int32, signals (n, 51).
 * The backward call returns the VJP w.r.t. the electrons' x, y, z, long_diff, n_electrons columns (other columns of
 * g_electrons_d are zeroed) and w.r.t. the pixel centres. */
typedef struct larnd_current_columns {
  int32_t ncols, x, y, z, long_diff, n_electrons, pixel_plane;
} larnd_current_columns_t;
int larnd_current_mc(const float* electrons_d, int64_t n, const larnd_current_columns_t* cols, const float* pixels_coord_d,
                     const larnd_params_t* params, int32_t* t0_tick_d, float* signals_d, void* stream);
int larnd_current_mc_backward(const float* electrons_d, int64_t n, const larnd_current_columns_t* cols,
                              const float* pixels_coord_d, const larnd_params_t* params, const float* g_signals_d,
                              float* g_electrons_d, float* g_pixels_coord_d, void* stream);

/* accumulate_signals_parametrized (detsim_jax.py:209-228): wfs_d (npix, n_ticks) += signals (n, n_signal_ticks) at row
 * pix_id, ticks start+k (< 0 or >= n_ticks-1 -> column 0, else +1).  Backward: g_signals = gather of g_wfs. */
int larnd_accumulate_parametrized(float* wfs_d, int32_t npix, int32_t n_ticks, const float* signals_d, int32_t n_signal_ticks,
                                  const int32_t* pix_id_d, const int32_t* start_ticks_d, int64_t n, void* stream);
int larnd_accumulate_parametrized_backward(const float* g_wfs_d, int32_t npix, int32_t n_ticks, float* g_signals_d,
                                           int32_t n_signal_ticks, const int32_t* pix_id_d, const int32_t* start_ticks_d,
                                           int64_t n, void* stream);

/* Optional device-side timing of the dominant kernels (used by bench.py for the roofline numbers): when
 * enabled, CUDA events are recorded on the launching stream immediately around
 *   slot 0: k_prepare   slot 1: k_lut_accumulate   slot 2: k_lut_backward   slot 3: k_fee_forward
 * larnd_profile_read synchronises on those events and returns the elapsed milliseconds of the LAST launch of
 * each kernel (-1 if it has not run since larnd_profile_enable(1)). */
/* Number of kernels this library has launched in this process so far (every launch site counts itself): lets a caller
 * report how many of OUR kernels ran inside a timed region instead of asserting a constant. */
uint64_t larnd_launch_count(void);

#define LARND_PROF_SLOTS 4
int larnd_profile_enable(int on);
int larnd_profile_read(float* ms_out /* [LARND_PROF_SLOTS] */);

/* Layout of the per-segment records inside the workspace (for tests / debugging): field f of segment s
 * is ((float*)workspace_d)[f * n_segments + s]; integer fields are bit-cast int32. */
enum {
  LARND_F_Q = 0, LARND_F_FRAC, LARND_F_SL, LARND_F_A, LARND_F_B, LARND_F_C,
  LARND_F_WX0, LARND_F_WX1, LARND_F_WX2, LARND_F_WX3, LARND_F_WX4,
  LARND_F_WY0, LARND_F_WY1, LARND_F_WY2, LARND_F_WY3, LARND_F_WY4,
  LARND_F_TD, LARND_F_X0, LARND_F_Y0, LARND_F_ST, LARND_F_REC, LARND_F_FT,
  LARND_F_XI /* recombination csi */, LARND_F_COS2 /* cos^2(phi), ellipsoid model */,
  LARND_I_T0 /* Nt-L-ct */, LARND_I_IDX, LARND_I_BX, LARND_I_BY, LARND_I_EP /* event*n_tpc+plane */,
  LARND_I_FLAGS /* bit0 TPC mask, bit1 z>z_anode, bit2 z>z_cathode */, LARND_I_MAINPIX,
  LARND_NFIELDS
};

#ifdef __cplusplus
}
#endif
#endif /* LARND_B200_H */
