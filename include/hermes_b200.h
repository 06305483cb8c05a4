/* hermes_b200.h -- C-ABI of the B200-native HermesPy channel hot path.
 *
 * One shared library (libhermes_b200.so, sm_100a only) exposes the batched replacements of the
 * reference's per-sample numpy routines.  Plain pointers and sizes only; no torch / numpy types.
 * Every entry point returns an hb_status (0 == HB_OK); hb_last_error() gives the message of the
 * last failure on the calling thread.  No exceptions cross this boundary and there is no CPU
 * fallback: without a CUDA device every compute entry point returns HB_ERR_NO_DEVICE.
 *
 * Reference interfaces replaced (paths relative to the HermesPy tree, v1.6.0):
 *   hb_fading_propagate*   <- MultipathFadingSample._propagate          hermespy/channel/fading/fading.py:371-406
 *                             + MultipathFadingSample.__path_impulse_generator        fading.py:293-343
 *   hb_fading_state        <- MultipathFadingSample.state (SISO tap gains)            fading.py:345-358
 *   hb_kron_mix            <- antenna-correlation mixing  R_rx @ S @ R_tx             fading.py:480-489
 *   hb_fading_sample       <- MultipathFadingRealization._sample (normals -> parameters)  fading.py:468-515
 *   hb_cdl_*               <- ClusterDelayLineSample.__ray_impulse_generator / _propagate
 *                             hermespy/channel/cdl/cluster_delay_lines.py:409-558
 *   hb_stats_accumulate    <- ScalarEvaluationResult.add_artifact  hermespy/core/pymonte/scalar.py:101-125
 *   hb_bit_errors          <- BitErrorEvaluator.evaluate/.artifact hermespy/modem/evaluators.py:231-259
 *   hb_receive_combine     <- receive superposition + AWGNRealization.add_to
 *                             hermespy/simulation/simulated_device.py:1899-1915, simulation/rf/noise/model.py:140-160
 *
 * Layouts (row-major, batch first):
 *   x      [B, Ntx, T]        complex64 (float2) or complex128 (double2), read-only
 *   y      [B, Nrx, T + D]    same element type as x, fully overwritten
 *   omega  [B, L, N+1] f64    per-sample angular increment of every sinusoid (column 0 = LOS term)
 *   phi    [B, L, N+1] f64    start phases
 *   amp    [B, L, 2]   f64    (LOS amplitude, per-sinusoid NLOS amplitude) incl. sqrt(gain*power_l)
 *   spatial[B, Nrx, Ntx] complex128
 *   so that  h_l[n] = amp_los e^{j(omega_l0 n + phi_l0)} + amp_nlos sum_k e^{j(omega_lk n + phi_lk)}
 *   and      y = spatial @ sum_l shift_{d_l}(x * h_l)          (d_l = tap_delay[l], ascending).
 */
#ifndef HERMES_B200_H
#define HERMES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define HB_API __attribute__((visibility("default")))
#else
#define HB_API
#endif

#define HB_VERSION 110 /* 0.1.1: antenna element models, per-link delay tables, fused receive, SINC mode */
#define HB_MAX_TAPS 256
#define HB_MAX_POLY_ORDER 8
#define HB_NUM_KERNEL_KINDS 9
#define HB_CDL_MAX_TERMS 1024
#define HB_CDL_MAX_GROUPS 64

typedef enum hb_status {
  HB_OK = 0,
  HB_ERR_INVALID = 1,     /* bad argument (mirrors the reference's ValueError) */
  HB_ERR_CUDA = 2,        /* CUDA runtime failure, see hb_last_error() */
  HB_ERR_UNSUPPORTED = 3, /* shape outside the compiled kernel set */
  HB_ERR_NO_DEVICE = 4    /* no sm_100 device visible; there is no CPU fallback */
} hb_status;

typedef enum hb_precision {
  HB_F32 = 0, /* complex64 arithmetic, rel. L2 <= 1e-5 against the float64 reference */
  HB_F64 = 1  /* float64 parity mode (<= 1e-12, bit-exact BER counts): FP64 Taylor path (truncation bound 1e-14) while the
                 frame's largest phase stays below 1000 rad, per-sample FP64 sincos beyond (or with HB_SOS_DIRECT) */
} hb_precision;

typedef enum hb_sos_mode {
  HB_SOS_AUTO = 0,  /* pick POLY when the error bound allows it, else DIRECT */
  HB_SOS_POLY = 1,  /* per-tile Taylor moments of the sum of sinusoids (FMA pipe)  */
  HB_SOS_DIRECT = 2, /* one sincos per sinusoid per sample (MUFU pipe)             */
  HB_SOS_POLY_GATHER = 3, /* POLY, forcing the per-group gather kernel (sparse / very long delay spreads) */
  HB_SOS_POLY_WINDOW = 4, /* POLY, forcing the cp.async-staged sliding-window kernel where it is eligible  */
  HB_SOS_POLY_TMA = 5,    /* POLY, preferring the persistent TMA-pipelined window kernel (what AUTO/POLY pick
                             for complex64 frames with T % 16 == 0, T + D >= 2048 and delays below 128 samples) */
  HB_SOS_POLY_FUSED = 6,  /* POLY, large arrays (16..64 antennas per side): the single-kernel GEMM + delay-line variant
                             (fading_fused.cuh).  Measured slower than the two-kernel path on B200 (profiles/r02_c4.md),
                             therefore not what AUTO picks; kept selectable */
  HB_SOS_POLY_SISO = 7    /* POLY, forcing the single-antenna time-packed kernel (1 x 1 links; measured level with the window
                             kernel that AUTO picks there, profiles/r02_siso.md) */
} hb_sos_mode;

typedef enum hb_poly_variant {
  HB_VARIANT_GATHER = 0, /* tdl_poly_kernel: one shared-memory read per (delay group, antenna, output)  */
  HB_VARIANT_WINDOW = 1, /* tdl_window_kernel: register sliding window along the delay axis            */
  HB_VARIANT_TMA = 2,    /* tdl_tma_kernel: the same walk, persistent CTAs fed by TMA (swizzled time-pair loads) */
  HB_VARIANT_SISO = 4,   /* tdl_siso_kernel (1 x 1 links): planar staging, FFMA2 packed over pairs of consecutive outputs */
  HB_VARIANT_FUSED = 3   /* fused_gemm_tdl_kernel (16..64 antennas per side): spatial GEMM on tcgen05, delay lines on its
                            accumulator through a shared-memory history ring -- the intermediate never reaches HBM */
} hb_poly_variant;

/* Launch-uniform description of one batched fading propagation. */
typedef struct hb_fading_problem {
  int32_t batch;          /* B: links (drop x sweep point x direction) in this launch       */
  int32_t num_tx;         /* Ntx streams of x                                                */
  int32_t num_rx;         /* Nrx streams of y                                                */
  int32_t num_samples;    /* T                                                               */
  int32_t max_delay;      /* D = round(max(delay) * fs)  (fading.py:372); y has T + D samples */
  int32_t num_taps;       /* L <= HB_MAX_TAPS                                                */
  int32_t num_sinusoids;  /* N (NLOS sinusoids per tap; the LOS term is extra)               */
  int32_t precision;      /* hb_precision                                                    */
  int32_t io_complex128;  /* 0: x, y are complex64; 1: complex128                            */
  int32_t sos_mode;       /* hb_sos_mode                                                     */
  double omega_max;       /* upper bound of |omega| over the batch (rad / sample); 0 = static */
  const int32_t* tap_delay; /* HOST int32[L]: rint(delay_l * fs) (fading.py:297), ascending   */
  const double* omega;    /* DEVICE f64 [B, L, N+1]                                          */
  const double* phi;      /* DEVICE f64 [B, L, N+1]                                          */
  const double* amp;      /* DEVICE f64 [B, L, 2]                                            */
  const void* spatial;    /* DEVICE complex128 [B, Nrx, Ntx]                                 */
} hb_fading_problem;

/* What hb_fading_plan() decided for a problem (also filled by propagate when info != NULL). */
typedef struct hb_fading_plan_info {
  int32_t mode;        /* HB_SOS_POLY or HB_SOS_DIRECT */
  int32_t tile;        /* output samples per CTA */
  int32_t poly_order;  /* number of Taylor terms P (POLY only) */
  int32_t num_groups;  /* distinct integer delays */
  int32_t num_tiles;   /* CTAs per link */
  int32_t launches;    /* kernels launched per call */
  double error_bound;  /* bound on the relative truncation error of the POLY expansion */
  int32_t variant;     /* hb_poly_variant (POLY only) */
  int32_t poly_tile;   /* samples per Taylor expansion window (POLY only; a multiple of tile) */
} hb_fading_plan_info;

HB_API int hb_version(void);
HB_API const char* hb_last_error(void);
/* Number of CUDA devices visible (0 when none / driver missing). */
HB_API int hb_device_count(void);
/* Make `device` the CUDA device of the calling thread (cudaSetDevice).  The host-buffer entries and every
 * device-pointer entry run on the calling thread's current device; a plugin that serves one GPU per process calls
 * this once per thread (hermespy_b200.dropin does it on every call: worker threads default to device 0). */
HB_API int hb_set_device(int device);

HB_API int hb_fading_plan(const hb_fading_problem* p, hb_fading_plan_info* info);

/* Fractional delays (extension; the reference rounds, fading.py:297): polyphase expansion of L taps with real-valued
 * delays into windowed-sinc integer-delay taps.  Tap l at delay_samples[l] = floor + eps becomes the 2 W taps
 *   j = floor - W + 1 .. floor + W   with weight  g(j - delay) = sinc(u) I0(beta sqrt(1 - (u / W)^2)) / I0(beta),
 * taps at negative delays (non-causal precursor) and zero-weight taps are dropped, integer delays stay single taps.
 * HOST arrays: out_delay / out_weight / out_source hold up to `capacity` entries, sorted by delay (stable in l):
 * expanded tap t has integer delay out_delay[t], the sinusoid parameters of tap out_source[t] and its amplitudes
 * scaled by out_weight[t].  Feeding the expanded set to hb_fading_propagate* (whose kernels merge equal delays) IS the
 * InterpolationMode.SINC path (hermespy/core/definitions.py:82-93).  Returns the number of expanded taps through
 * num_out, HB_ERR_UNSUPPORTED when it exceeds capacity. */
HB_API int hb_fading_sinc_taps(const double* delay_samples, int32_t num_taps, int32_t half_width, double kaiser_beta,
                               int32_t capacity, int32_t* out_delay, double* out_weight, int32_t* out_source,
                               int32_t* num_out);

/* Device-resident propagation.  x, y and all DEVICE members of p live on the current device;
 * the work is enqueued on `stream` (a cudaStream_t; NULL = legacy default stream) and the call
 * returns without synchronizing. */
HB_API int hb_fading_propagate(const hb_fading_problem* p, const void* x, void* y, void* stream,
                        hb_fading_plan_info* info);

/* Host-buffer propagation (the call a drop-in plugin makes): every pointer of `p`, x and y are HOST
 * buffers (pinned or pageable).  The library stages chunks of `chunk_links` links through its own
 * device workspace on three streams (H2D / kernels / D2H overlapped) and returns after y is complete.
 * chunk_links <= 0 lets the library choose. */
HB_API int hb_fading_propagate_host(const hb_fading_problem* p, const void* x, void* y, int32_t chunk_links,
                             hb_fading_plan_info* info);

/* SISO tap gains of the channel state, h[B, G, T] complex (element type per io_complex128), one row per
 * distinct integer delay, plus the delays themselves in group_delay_out (HOST int32[G], may be NULL).
 * CSI[b, i, j, n, group_delay[g]] = spatial[b, i, j] * h[b, g, n]   (fading.py:351-364). */
HB_API int hb_fading_state(const hb_fading_problem* p, void* h, int32_t* group_delay_out, void* stream);

/* ---- 3GPP cluster delay line ------------------------------------------------------------------------------
 * Batched replacement of ClusterDelayLineSample._propagate / .state and the per-ray array responses
 * (hermespy/channel/cdl/cluster_delay_lines.py:409-592, hermespy/core/antennas.py:138-210, 883-1000).
 * Antenna elements: the reference's four element models -- IdealAntenna, LinearAntenna(slant), PatchAntenna, Dipole
 * (core/antennas.py:392-622) -- each with its own orientation inside the array (element_mode below).
 *
 * A "ray term" is one iteration of the reference's ray generator: (cluster sub-partition, ray) in generator order.
 * Per term the host supplies the four ray angles, the 2x2 Jones matrix and the real amplitude
 * sqrt(P_c / num_rays) * nlos_scale; the delay index int((tau + offset) * fs) is launch-uniform.  The optional
 * line-of-sight term (cluster_delay_lines.py:498-523) is synthesized by the library from the device poses.
 */
#define HB_ELEMENT_STRIDE 12
typedef enum hb_element_mode { HB_ELEMENTS_IDEAL = 0, HB_ELEMENTS_UNIFORM = 1, HB_ELEMENTS_PER_ELEMENT = 2 } hb_element_mode;
typedef enum hb_element_kind {
  HB_ELEMENT_IDEAL = 0,  /* F = [2^-1/2, 2^-1/2]                                    core/antennas.py:435-436 */
  HB_ELEMENT_LINEAR = 1, /* F = [cos(slant), sin(slant)]                            core/antennas.py:509-510 */
  HB_ELEMENT_PATCH = 2,  /* F = [max(0.1, (0.1 + 0.9 e^{-1.315 az^2}) cos^2 ze), 0] core/antennas.py:556-560 */
  HB_ELEMENT_DIPOLE = 3  /* F = [cos(pi/2 cos ze) / sin ze, 0] (0 at ze = 0)        core/antennas.py:610-614 */
} hb_element_kind;

typedef enum hb_cdl_variant {
  HB_CDL_VARIANT_AUTO = 0,   /* tensor-core kernel from 8 transmit antennas when it takes the problem, else the gather kernel */
  HB_CDL_VARIANT_GATHER = 1, /* cdl_poly_kernel: FP32 pipe, moments and x tile from shared memory                           */
  HB_CDL_VARIANT_UMMA = 2,   /* cdl_umma_kernel: tcgen05 3xTF32, delay groups folded into K (HB_ERR_UNSUPPORTED if not eligible) */
  HB_CDL_VARIANT_UMMA_BF16 = 3 /* cdl_umma_bf16_kernel: the same GEMM as BF16x3 (kind::f16), 18 % less operand traffic; up to 32
                                  accumulator columns per term (2 P Nrx <= 32) */
} hb_cdl_variant;

typedef struct hb_cdl_problem {
  int32_t batch;            /* B links                                                              */
  int32_t num_tx, num_rx;   /* antenna counts                                                       */
  int32_t num_samples;      /* T                                                                    */
  int32_t max_delay;        /* D = ceil(max_delay * fs) (cluster_delay_lines.py:528)                */
  int32_t num_terms;        /* Rn <= HB_CDL_MAX_TERMS - 1: non-line-of-sight ray terms per link     */
  int32_t line_of_sight;    /* 1: add the LOS term                                                  */
  int32_t los_delay;        /* delay index of the LOS term: int((cluster_delays[0] + offset) * fs)  */
  int32_t precision;        /* hb_precision                                                         */
  int32_t io_complex128;    /* element type of x / y                                                */
  double carrier_frequency; /* Hz                                                                   */
  double sampling_rate;     /* Hz (LinkState.bandwidth)                                             */
  double los_amplitude;     /* sqrt(K / (1 + K)), K linear (cluster_delay_lines.py:516)             */
  double max_speed;         /* bound on |v_rx - v_tx| over the batch in m/s (0 = static links)      */
  const int32_t* term_delay;   /* HOST int32[Rn] delay index per term                               */
  const double* angles;        /* DEVICE f64 [B, Rn, 4]: aoa, zoa, aod, zod (radians)               */
  const void* jones;           /* DEVICE complex128 [B, Rn, 2, 2]                                   */
  const double* amplitude;     /* DEVICE f64 [B, Rn]                                                */
  const double* tx_pose;       /* DEVICE f64 [B, 12]: rotation (row-major 3x3, array -> global), translation */
  const double* rx_pose;       /* DEVICE f64 [B, 12]                                                */
  const double* rel_velocity;  /* DEVICE f64 [B, 3]: v_rx - v_tx (global frame)                     */
  const double* tx_topology;   /* DEVICE f64 [Ntx, 3] element positions in the array frame          */
  const double* rx_topology;   /* DEVICE f64 [Nrx, 3]                                               */
  /* Antenna element models (core/antennas.py:138-210 global_characteristics, :392-622 local patterns).
   * element_mode HB_ELEMENTS_IDEAL: unrotated IdealAntenna elements, the two tables are ignored (may be NULL);
   * HB_ELEMENTS_UNIFORM: every element of an array equals row 0 of its table (rank-one ray matrices);
   * HB_ELEMENTS_PER_ELEMENT: one row per element (rank-two ray matrices a_rx J a_tx^T).
   * Row layout (HB_ELEMENT_STRIDE doubles): [0..8] rotation element frame -> array frame (row-major 3x3),
   * [9] hb_element_kind, [10] kind parameter (LinearAntenna: slant in radians), [11] reserved. */
  int32_t element_mode;        /* hb_element_mode                                                   */
  int32_t variant;             /* hb_cdl_variant: which K6 kernel the POLY mode runs (f32 only)     */
  const double* tx_elements;   /* DEVICE f64 [Ntx or 1, HB_ELEMENT_STRIDE]                          */
  const double* rx_elements;   /* DEVICE f64 [Nrx or 1, HB_ELEMENT_STRIDE]                          */
  /* Heterogeneous batches: links whose realizations drew their own cluster delays, cluster counts and line-of-sight
   * state (the stochastic 3GPP scenarios, cluster_delay_lines.py:1824-2013) in ONE launch set.  link_term_delay != NULL
   * switches the delay structure from launch-uniform to per link: every link gets its own term -> delay-group table in
   * device memory, padded to the batch maximum of groups.  num_terms and line_of_sight stay batch-wide: links with fewer
   * rays are padded by the caller with zero-amplitude rays, links without a line of sight carry link_los_amplitude 0.
   * The three arrays are HOST memory in both the device-pointer and the host-buffer entry (like term_delay). */
  const int32_t* link_term_delay;   /* HOST int32 [B, Rn], NULL = term_delay for every link               */
  const int32_t* link_los_delay;    /* HOST int32 [B] (only read when line_of_sight), NULL = los_delay    */
  const double* link_los_amplitude; /* HOST f64 [B] (only read when line_of_sight), NULL = los_amplitude  */
} hb_cdl_problem;

typedef struct hb_cdl_plan_info {
  int32_t mode;        /* HB_SOS_POLY (grouped moment path) or HB_SOS_DIRECT (per-ray FP64 path) */
  int32_t tile;
  int32_t poly_order;
  int32_t num_groups;
  int32_t num_tiles;
  int32_t launches;
  double error_bound;
  int32_t variant;     /* hb_cdl_variant that runs (POLY only): GATHER or UMMA */
  int32_t poly_tile;   /* samples per Taylor window of the moments: == tile (GATHER), 1/2/4/8 tiles (UMMA) */
} hb_cdl_plan_info;

HB_API int hb_cdl_plan(const hb_cdl_problem* p, hb_cdl_plan_info* info);
/* Device-resident: x [B, Ntx, T] -> y [B, Nrx, T + D]; enqueued on `stream`, no synchronization. */
HB_API int hb_cdl_propagate(const hb_cdl_problem* p, const void* x, void* y, void* stream, hb_cdl_plan_info* info);
/* Host buffers for every pointer (chunked H2D / kernels / D2H pipeline). */
HB_API int hb_cdl_propagate_host(const hb_cdl_problem* p, const void* x, void* y, int32_t chunk_links,
                                 hb_cdl_plan_info* info);
/* Channel state per delay group: h [B, G, Nrx, Ntx, T] complex (io_complex128), FP64 evaluation;
 * group_delay_out: HOST int32[G] (may be NULL); returns G through num_groups_out. */
HB_API int hb_cdl_state(const hb_cdl_problem* p, void* h, int32_t* group_delay_out, int32_t* num_groups_out,
                        void* stream);

/* ---- evaluator statistics (the only data that ever crosses GPUs) and antenna-correlation mixing ---------------
 * Bit errors of num_drops drops (BitErrorEvaluator, evaluators.py:239-259): tx_bits / rx_bits are DEVICE uint8
 * [num_drops, num_bits] in 0/1 format, tx_len / rx_len (DEVICE int32[num_drops], may be NULL = num_bits) the valid
 * lengths, the shorter sequence is zero-padded.  Outputs (DEVICE): errors[i] = sum |tx - rx|, bits[i] = max(len),
 * artifact[i] = errors / bits (np.mean of the indicators; may be NULL). */
HB_API int hb_bit_errors(const uint8_t* tx_bits, const uint8_t* rx_bits, const int32_t* tx_len, const int32_t* rx_len,
                         int32_t num_drops, int32_t num_bits, int64_t* errors, int64_t* bits, double* artifact,
                         void* stream);
/* Running statistics per grid cell (ScalarEvaluationResult.add_artifact, scalar.py:109-115): for every drop i with
 * cell[i] == c:  stats[c] += (artifact, artifact^2, 1),  counts[c] += (errors, bits).  stats: DEVICE f64
 * [num_cells, 3], counts: DEVICE int64 [num_cells, 2] (errors / bits / counts may be NULL).  Fixed summation order.
 * These buffers are what the evaluator all-reduce (ncclAllReduce SUM) reads. */
HB_API int hb_stats_accumulate(const double* artifact, const int32_t* cell, const int64_t* errors, const int64_t* bits,
                               int32_t num_drops, int32_t num_cells, double* stats, int64_t* counts, void* stream);
/* out[b] = R_rx @ spatial[b] @ R_tx, complex128 DEVICE, FP64 (fading.py:480-489: the covariance matrices themselves,
 * not their square roots).  r_rx [Nrx, Nrx] / r_tx [Ntx, Ntx] are shared by the batch; NULL = identity.
 * out may alias spatial. */
HB_API int hb_kron_mix(const void* r_rx, const void* spatial, const void* r_tx, void* out, int32_t batch,
                       int32_t num_rx, int32_t num_tx, void* stream);

/* ---- receive side: superposition of the impinging signals + additive white Gaussian noise, one fused pass ------------
 * out[b, i, m] = sum_k in_k[b, i, m - offset_k]  +  noise_scale[b] (noise_re[b, i, m] + j noise_im[b, i, m])
 * Replaces, for impinging signals of one sampling rate / carrier frequency with whole-sample delays, the superposition
 * loop of SimulatedDevice.process_input (hermespy/simulation/simulated_device.py:1899-1915) and AWGNRealization.add_to
 * (hermespy/simulation/rf/noise/model.py:140-160).  The standard normals come from the caller's numpy generator (noise is
 * part of the drop's random stream): two DEVICE float64 planes [B, Nrx, T]; noise_scale: DEVICE f64 [B] = sqrt(P_b / 2).
 * noise_re == NULL: superposition only.  inputs: HOST array of descriptors with DEVICE sample pointers [B, Nrx, T_k]
 * (element type per io_complex128); out: DEVICE [B, Nrx, num_out_samples].  complex128 results are bit-identical to the
 * reference's numpy arithmetic (same summation order, no fused multiply-add). */
typedef struct hb_receive_input {
  const void* samples;  /* DEVICE [B, Nrx, num_samples] */
  int32_t num_samples;  /* T_k */
  int32_t offset;       /* first output sample this signal lands on (delay in samples) */
} hb_receive_input;
HB_API int hb_receive_combine(const hb_receive_input* inputs, int32_t num_inputs, const double* noise_re,
                              const double* noise_im, const double* noise_scale, void* out, int32_t batch, int32_t num_rx,
                              int32_t num_out_samples, int32_t io_complex128, void* stream);

/* a3 on the device: the standard normals of B static realizations (drawn by the caller's numpy generator -- the draw
 * order is part of parity) -> the kernel parameter blocks of B links.  Replaces MultipathFadingRealization._sample
 * (hermespy/channel/fading/fading.py:468-515) with ConsistentUniform.sample = norm.cdf (hermespy/channel/consistent.py:475-485)
 * for decorrelation distance = inf.  normals: DEVICE f64 [B, num_scalars], num_scalars = dim^2 + 2 L + 2 L N in declaration
 * order (fading.py:742-754); amp_table: DEVICE f64 [L, 2] = (los_gain_l, nlos_gain_l) sqrt(gain power_l); los_rate / nlos_rate:
 * Doppler "frequencies" divided by the sampling rate (used as angular rates); reciprocal: transpose the antenna
 * phases (fading.py:517-538).  Outputs (DEVICE): omega, phi [B, L, N + 1], amp [B, L, 2], spatial complex128 [B, Nrx, Ntx]
 * (before antenna correlation: follow with hb_kron_mix). */
HB_API int hb_fading_sample(const double* normals, int32_t batch, int32_t num_scalars, int32_t antenna_dim, int32_t num_rx,
                            int32_t num_tx, int32_t num_taps, int32_t num_sinusoids, const double* amp_table,
                            double los_rate, double nlos_rate, int32_t reciprocal, double* omega, double* phi, double* amp,
                            void* spatial, void* stream);

/* K4 for large arrays (SURVEY 8(b) `hb_spatial_gemm_3xtf32`): y[b] = spatial[b] @ z[b], the `spatial_response @
 * propagated` product of fading.py:395, on the tcgen05 tensor cores in 3xTF32 (FP32-equivalent accuracy, FP32
 * accumulation in tensor memory).  spatial: DEVICE complex128 [B, Nrx, Ntx]; z: DEVICE complex64 [B, Ntx, T];
 * y: DEVICE complex64 [B, Nrx, T].  Any Nrx, Ntx (blocks of 64 x 64 antennas per launch). */
HB_API int hb_spatial_gemm_3xtf32(const void* spatial, const void* z, void* y, int32_t batch, int32_t num_rx,
                                  int32_t num_tx, int32_t num_samples, void* stream);

/* Per-kernel accounting.  Kinds: 0 sos_poly_coef, 1 tdl_poly, 2 tdl_direct, 3 sos_state, 4 cdl_rays,
 * 5 cdl_propagate, 6 spatial_gemm, 7 stats, 8 misc.
 * hb_launch_counts: launches of each kind since load (always counted).
 * hb_profile_begin/end: between the two calls every kernel launch is bracketed by CUDA events recorded on
 * its own launch stream; hb_profile_end synchronizes those events and returns summed device time (ms) and
 * launch counts per kind.  Meant for bench.py's live roofline measurement. */
typedef struct hb_profile_report {
  double ms[HB_NUM_KERNEL_KINDS];
  int64_t launches[HB_NUM_KERNEL_KINDS];
} hb_profile_report;
HB_API void hb_launch_counts(int64_t* counts /* [HB_NUM_KERNEL_KINDS] */);
HB_API int hb_profile_begin(void);
HB_API int hb_profile_end(hb_profile_report* report);

/* Release cached device workspaces / streams of the calling process. */
HB_API void hb_release(void);

/* Leave `num_sms` streaming multiprocessors out of the grids of the persistent kernels (tdl_tma, spatial GEMM) so that a
 * concurrent collective (the NCCL all-reduce of the evaluator statistics, montecarlo.py) finds room to run instead of
 * queueing behind a kernel that owns every SM.  0 (default) uses all SMs.  Returns the previous value. */
HB_API int hb_reserve_sms(int num_sms);

#ifdef __cplusplus
}
#endif
#endif /* HERMES_B200_H */
