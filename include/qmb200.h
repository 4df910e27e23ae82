/*
 * qmb200.h -- C ABI of libqmb200.so: the B200-native (sm_100a) implementation of
 * QuakeMigrate's coalescence hot path.
 *
 * Two tiers of entry points:
 *
 *  T1  the five symbols of the reference C extension, with byte-identical signatures, so that
 *      libqmb200.so can literally be loaded by an unmodified quakemigrate.core
 *      (quakemigrate/core/libnames.py:35-38 opens  core/src/qmlib<EXT_SUFFIX>  with ctypes.CDLL;
 *      prototypes: quakemigrate/core/src/qmlib.h:28-44, export list qmlib.def:3-7;
 *      ctypes argtypes: quakemigrate/core/lib.py:38-49,128,173,211,249).
 *      Host pointers in, host pointers out, GPU inside.  They honour the reference contract
 *      (caller allocates everything; map4d arrives zeroed and is accumulated into; STA/LTA
 *      outputs arrive pre-filled) and therefore move N*T doubles over PCIe: plumbing, not speed.
 *
 *  T2  the fused path QuakeScan._compute needs (quakemigrate/signal/scan.py:594-647):
 *      device-resident travel-time tables, onset functions computed/kept on the device, the
 *      4-D volume never materialised for detect(), a marginal window for locate(), node-sharded
 *      multi-GPU with NCCL.  All functions return an int status (0 = QMB_OK) and record a
 *      message retrievable with qmb_last_error(); host buffers are always caller-owned, exactly
 *      like the reference (C never allocates, frees or retains host pointers).
 *
 * There is NO CPU fallback anywhere behind this header: without a CUDA device every T2 call
 * returns QMB_ERR_CUDA and every T1 call fails (see qmb_legacy_status).
 */
#ifndef QMB200_H
#define QMB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * T1 -- reference-compatible symbols (qmlib.h:28-44)
 * ---------------------------------------------------------------------------------------- */

/* qmlib.h:34-38; numpy mirror quakemigrate/core/lib.py:33-36 */
typedef struct {
  int n;
  int nsta;
  int nlta;
} stalta_header;

/* replaces migrate(), quakemigrate/core/src/migratelib.c:40-65 */
void migrate(double *onsets, int32_t *lookup_tables, double *map4d, int32_t fsmp, int32_t lsmp,
             int32_t n_samples, int32_t n_stations, int32_t available, int64_t n_nodes,
             int64_t threads);

/* replaces find_max_coa(), quakemigrate/core/src/migratelib.c:85-111 */
void find_max_coa(double *map4d, double *max_coa, double *max_norm_coa, int64_t *max_coa_idx,
                  int32_t n_samples, int64_t n_nodes, int64_t threads);

/* replace quakemigrate/core/src/onsetlib.c:35-59, :79-108, :126-148 */
void overlapping_sta_lta(const double *signal, const stalta_header *head, double *onset);
void centred_sta_lta(const double *signal, const stalta_header *head, double *onset);
void recursive_sta_lta(const double *signal, const stalta_header *head, double *onset);

/* The five symbols above return void like the reference's, so a failure (no device, out of memory,
 * a travel time beyond the post-pad -- where the reference would read past its onset rows) cannot
 * be returned: it is recorded per thread, reported on stderr, and the call returns with the output
 * buffers untouched.  qmb_legacy_status() returns the status of this thread's last T1 call (QMB_OK
 * = 0, else a QMB_ERR_* code; message in qmb_last_error()) and clears it; quakemigrate_b200.core
 * checks it after every call and raises.  Set QMB_LEGACY_ABORT=1 to abort() instead -- the only way
 * an unmodified quakemigrate.core, which never asks, can be stopped from using a zero map. */
int qmb_legacy_status(void);

/* ------------------------------------------------------------------------------------------
 * T2 -- fused device-resident path
 * ---------------------------------------------------------------------------------------- */

enum {
  QMB_OK = 0,
  QMB_ERR_CUDA = 1,        /* CUDA runtime / driver failure (includes "no device") */
  QMB_ERR_ARG = 2,         /* invalid argument (the reference's ValueError cases, lib.py:105-110) */
  QMB_ERR_UNSUPPORTED = 3, /* shape outside what the kernels were built for */
  QMB_ERR_NCCL = 4,        /* NCCL not loadable / collective failure */
  QMB_ERR_STATE = 5        /* call sequence error (e.g. fetch before scan) */
};

enum { QMB_STALTA_CLASSIC = 0, QMB_STALTA_CENTRED = 1, QMB_STALTA_RECURSIVE = 2 };
/* signal_transform of STALTAOnset._onset (stalta.py:516-523): x^2, |x|, |hilbert(x)|, |hilbert(x)|^2 */
enum {
  QMB_TRANSFORM_NONE = 0,
  QMB_TRANSFORM_ENERGY = 1,
  QMB_TRANSFORM_ABS = 2,
  QMB_TRANSFORM_ENV = 3,
  QMB_TRANSFORM_ENV_SQUARED = 4
};

typedef struct qmb_ctx *qmb_ctx_t;       /* one per (process, GPU) */
typedef struct qmb_lut *qmb_lut_t;       /* device-resident travel-time table (node slab) */
typedef struct qmb_onsets *qmb_onsets_t; /* device-resident log-onset rows */

const char *qmb_last_error(void);
const char *qmb_version(void);
int qmb_device_count(int *count);

int qmb_ctx_create(int device, qmb_ctx_t *out);
int qmb_ctx_destroy(qmb_ctx_t ctx);
int qmb_ctx_sync(qmb_ctx_t ctx);
/* kernel variant: nodes per warp (4 or 8); 0 = default.  Must be set before qmb_lut_upload. */
int qmb_ctx_set_variant(qmb_ctx_t ctx, int nodes_per_warp);
/* Operand format of the DETECT scan's gathers: mode 0 = automatic (default), 1 = float32 rows,
 * 2 = 16-bit fixed-point rows (two samples per 32-bit shared-memory word, exact integer stacks of
 * the quantised rows).  Whatever the format, the per-sample winner is decided among the candidates
 * inside a rigorous error band by re-evaluating them in float64 in the reference's summation
 * order, so arg-max indices and max_coa do not depend on it; max_norm_coa divides by a node sum
 * built from the scan's stacks, and the automatic mode keeps float32 rows whenever the estimated
 * error of that sum is not clearly inside 1e-5 (tiny grids, `available` << rows).  Also settable
 * with QMB_SCAN_PRECISION=fp32|q16.  qmb_ctx_last_scan_bytes_per_gather reports what the last
 * detect scan moved per gathered operand (2 or 4 bytes). */
int qmb_ctx_set_scan_precision(qmb_ctx_t ctx, int mode);
/* Fixed-point rows of the detect scan are quantised to 16 - log2(group) bits so that `group` rows
 * (1, 2, 4 or 8) can be added word-wise before the two 16-bit halves are separated: 1 = full 16-bit
 * rows, two instructions per shared-memory word; 4 = 14-bit rows, one.  The candidate band is
 * S * q_step (q_step doubles per bit dropped); results do not depend on it (float64 decision among
 * the candidates, verified node sums).  Applies to onsets and tables prepared afterwards; also
 * QMB_SCAN_QGROUP. */
int qmb_ctx_set_scan_qgroup(qmb_ctx_t ctx, int group);
int qmb_ctx_last_scan_bytes_per_gather(qmb_ctx_t ctx, int *bytes);
/* What the last detect scan did.  The 16-bit scan is VERIFIED per call: after the scan, 1-2 windows
 * of 248 samples are re-scanned over all local nodes with float32 rows and the per-sample node sums
 * (the denominator of max_norm_coa, migratelib.c:108) are compared; check_max_rel_err is the worst
 * relative difference seen on any rank.  In the automatic mode a difference above 2e-6 discards the
 * 16-bit result and the whole scan is redone with float32 rows (fell_back_fp32 = 1), so that
 * max_norm_coa stays inside the 1e-5 bar also on inputs whose quantisation residuals are correlated.
 * two_candidate_samples: samples whose two leading nodes were re-evaluated in float64;
 * slow_path_samples: samples re-scanned over all nodes in float64 (>= 3 candidates inside the band). */
typedef struct {
  int32_t operand_bytes;  /* 2: 16-bit fixed-point rows, 4: float32 rows (what produced the result) */
  int32_t fell_back_fp32;
  int32_t check_windows, check_samples;
  double check_max_rel_err;
  int32_t slow_path_samples, two_candidate_samples;
  float scan_ms, tail_ms, total_ms, check_ms; /* CUDA-event times: scan kernel, everything after it
                                                 (merge, exchange, candidates, slow path, check), both,
                                                 and the check windows alone (part of tail_ms) */
  int32_t host_syncs_in_timed_region; /* cudaStreamSynchronize calls between the first and last event */
  int32_t operand_bits;               /* 32 (float32 rows) or the bits of the fixed-point rows (13..16) */
} qmb_scan_info_t;
int qmb_ctx_last_scan_info(qmb_ctx_t ctx, qmb_scan_info_t *info);
/* number of launches of this library's own kernels since context creation */
int qmb_ctx_launch_count(qmb_ctx_t ctx, int64_t *count);

/* benchmarking helpers: evict L2 by writing a 256 MiB scratch buffer; page-lock caller memory
 * so that H2D/D2H copies of host buffers run at full PCIe rate (cudaHostRegister). */
int qmb_ctx_flush_l2(qmb_ctx_t ctx);
/* CUDA-event stopwatch on the context's stream (where all kernels of this library are launched):
 * start records an event, stop records a second one, waits for it and returns the device time
 * between them (host gaps between the bracketed calls included). */
int qmb_ctx_timer_start(qmb_ctx_t ctx);
int qmb_ctx_timer_stop(qmb_ctx_t ctx, float *ms);
int qmb_host_register(void *ptr, int64_t bytes);
int qmb_host_unregister(void *ptr);

/* --- multi-GPU (one process per GPU; ids are exchanged by the host application) ---------- */
#define QMB_NCCL_ID_BYTES 128
int qmb_comm_unique_id(void *id_out /* QMB_NCCL_ID_BYTES */);
int qmb_comm_init(qmb_ctx_t ctx, const void *id, int world, int rank);
int qmb_comm_destroy(qmb_ctx_t ctx);

/* --- travel-time table ---------------------------------------------------------------------
 * `lut` is what LUT.serve_traveltimes returns (quakemigrate/lut/lut.py:502-538): int32,
 * C-order (nx, ny, nz, n_rows), station axis fastest == the `lookup_tables` argument of
 * migrate().  For node-sharded runs each rank uploads its x-slab only:
 * (nx_local, ny, nz, n_rows) plus the flat index of its first node and the global node count
 * (needed for max_norm_coa = max * N / sum, migratelib.c:108).  Negative entries clamp to 0
 * (migratelib.c:55).  Upload once per availability pattern; reuse across calls. */
int qmb_lut_upload(qmb_ctx_t ctx, const int32_t *lut, int32_t nx_local, int32_t ny, int32_t nz,
                   int32_t n_rows, int64_t node_offset, int64_t n_nodes_global, qmb_lut_t *out);
/* The same for a slab given as a range of the C-order node list, [node_first, node_first +
 * n_nodes_local): whole z-columns (multiples of nz), but free to begin and end inside an x-plane, so
 * that G ranks can hold N / G nodes each whatever nx is.  `lut` holds the rows of exactly that range,
 * (n_nodes_local, n_rows). */
int qmb_lut_upload_range(qmb_ctx_t ctx, const int32_t *lut, int32_t ny, int32_t nz, int32_t n_rows,
                         int64_t node_first, int64_t n_nodes_local, int64_t n_nodes_global, qmb_lut_t *out);
int qmb_lut_free(qmb_lut_t lut);

/* LUT.serve_traveltimes on the device (quakemigrate/lut/lut.py:502-538).  Upload every
 * (station, phase) float64 travel-time grid [seconds, (n_grids, nx, ny, nz) C-order] once; then
 * each availability pattern is served without touching the host: row r of the table is
 * rint(grids[rows[r]] * sampling_rate) as int32 (negative -> 0), for the x-slab [x0, x1)
 * (x1 <= 0 means nx).  The result is identical to qmb_lut_upload of the host-served array. */
typedef struct qmb_ttgrids *qmb_ttgrids_t;
int qmb_ttgrids_upload(qmb_ctx_t ctx, const double *grids, int32_t n_grids, int32_t nx, int32_t ny,
                       int32_t nz, qmb_ttgrids_t *out);
int qmb_ttgrids_free(qmb_ttgrids_t grids);
/* LUT.decimate (quakemigrate/lut/lut.py:102-140) on resident grids: every df-th node per axis,
 * centred; returns new grids (the caller rescales node_spacing, lut.py:128).  qmb_ttgrids_shape
 * writes (n_grids, nx, ny, nz). */
int qmb_ttgrids_decimate(qmb_ctx_t ctx, qmb_ttgrids_t grids, const int32_t *df /* 3 */, qmb_ttgrids_t *out);
int qmb_ttgrids_shape(qmb_ttgrids_t grids, int32_t *shape4);
int qmb_lut_serve(qmb_ctx_t ctx, qmb_ttgrids_t grids, const int32_t *rows, int32_t n_rows,
                  double sampling_rate, int32_t x0, int32_t x1, qmb_lut_t *out);
typedef struct {
  int64_t n_nodes_local, n_nodes_global, node_offset;
  int32_t n_rows, max_tau, max_tile_span, n_tiles, nodes_per_tile;
  int64_t device_bytes;
  double duplicate_fraction; /* neighbouring nodes (inside a warp group) sharing a row's travel time */
  int32_t variant, reserved; /* kernel tiling variant id actually used (after span fall-back) */
} qmb_lut_info_t;
int qmb_lut_info(qmb_lut_t lut, qmb_lut_info_t *info);

/* --- onset functions -------------------------------------------------------------------------
 * qmb_onsets_upload: `onsets` are the LINEAR onset functions returned by
 * Onset.calculate_onsets (float64 (n_rows, n_total), base.py:103-106); the clip(0.01)+log of
 * quakemigrate/core/lib.py:93-94 happens on the device.
 * qmb_onsets_from_waveforms: STALTAOnset._onset on the device (stalta.py:516-548): transform,
 * per-channel STA/LTA (onsetlib.c), RMS over the channels of each row, clip(min_onset_value);
 * optionally returns the linear onsets to the host (the picker needs them, gaussian.py:142).
 *   waveforms   float64 (n_channels, n_total), host
 *   row_ptr     int32 (n_rows+1): channels of row r are row_chan[row_ptr[r] .. row_ptr[r+1])
 *   nsta, nlta  int32 (n_rows): window lengths in samples per row (stalta.py:394-396)
 */
int qmb_onsets_upload(qmb_ctx_t ctx, const double *onsets, int32_t n_rows, int64_t n_total,
                      qmb_onsets_t *out);
int qmb_onsets_from_waveforms(qmb_ctx_t ctx, const double *waveforms, int32_t n_channels,
                              int64_t n_total, int32_t n_rows, const int32_t *row_ptr,
                              const int32_t *row_chan, const int32_t *nsta, const int32_t *nlta,
                              int32_t position, int32_t transform, double min_onset_value,
                              double *onsets_out /* nullable, (n_rows, n_total) */,
                              qmb_onsets_t *out);
int qmb_onsets_free(qmb_onsets_t onsets);

/* --- waveform pre-processing ---------------------------------------------------
 * pre_process (quakemigrate/signal/onsets/stalta.py:203-209) and util.decimate
 * (quakemigrate/util.py:499-518) for blocks of equally long float64 traces kept on the device:
 *   detrend("linear") + detrend("constant")  (detrend != 0)
 *   -> cosine taper over int(taper_fraction * n) samples at either end  (obspy Trace.taper)
 *   -> IIR filter given as scipy second-order sections, (n_sections, 6) rows b0 b1 b2 a0 a1 a2,
 *      n_sections in 0..4 (0: no filter); zerophase != 0: forward, then over the reversed trace
 *      (obspy.signal.filter.bandpass/lowpass(zerophase=True))
 *   -> keep every `decimate`-th sample (Trace.decimate(factor, no_filter=True)).
 * qmb_waveforms_filter reads channels `channels[0..n_sel)` of `in` (null: all) and writes channels
 * [out_chan0, out_chan0 + n_sel) of `out`, whose traces must hold ceil(n / decimate) samples.
 * kernel_ms (nullable): device time of the kernels of this call (CUDA events).
 * qmb_onsets_from_device_waveforms is qmb_onsets_from_waveforms on such a block: raw traces ->
 * filtered traces -> onset functions without a host round trip.
 * qmb_preprocess is the one-call form with host buffers. */
typedef struct qmb_waveforms *qmb_waveforms_t;
int qmb_waveforms_alloc(qmb_ctx_t ctx, int32_t n_channels, int64_t n, qmb_waveforms_t *out);
int qmb_waveforms_upload(qmb_ctx_t ctx, const double *x, int32_t n_channels, int64_t n,
                         qmb_waveforms_t *out);
int qmb_waveforms_fetch(qmb_waveforms_t w, int32_t chan0, int32_t n_chan, double *host);
int qmb_waveforms_free(qmb_waveforms_t w);
int qmb_waveforms_filter(qmb_ctx_t ctx, qmb_waveforms_t in, const int32_t *channels, int32_t n_sel,
                         const double *sos, int32_t n_sections, int32_t detrend, double taper_fraction,
                         int32_t zerophase, int32_t decimate, qmb_waveforms_t out, int32_t out_chan0,
                         float *kernel_ms);
/* util.upsample's fence-post linear interpolation (quakemigrate/util.py:544-551) of every trace of a
 * block: (n - 1) * upfactor + 1 samples per trace, bit-identical to the numpy expression. */
int qmb_waveforms_upsample(qmb_ctx_t ctx, qmb_waveforms_t in, int32_t upfactor, qmb_waveforms_t *out);
int qmb_onsets_from_device_waveforms(qmb_ctx_t ctx, qmb_waveforms_t w, int32_t n_rows,
                                     const int32_t *row_ptr, const int32_t *row_chan,
                                     const int32_t *nsta, const int32_t *nlta, int32_t position,
                                     int32_t transform, double min_onset_value,
                                     double *onsets_out /* nullable */, qmb_onsets_t *out);
int qmb_preprocess(qmb_ctx_t ctx, const double *x, int32_t n_channels, int64_t n, const double *sos,
                   int32_t n_sections, int32_t detrend, double taper_fraction, int32_t zerophase,
                   int32_t decimate, double *out /* (n_channels, ceil(n / decimate)) */);

/* np.abs(scipy.signal.hilbert(x, axis=-1)) of a host (n_channels, n) block: the envelope behind the
 * "env" / "env_squared" transforms (stalta.py:520-523).  A DFT of exactly the trace length like
 * scipy's, by this library's own transforms: Bluestein's chirp-z convolution on power-of-two
 * Stockham FFT passes (no FFT library). */
int qmb_envelope(qmb_ctx_t ctx, const double *x, int32_t n_channels, int64_t n, double *out);

/* batched STA/LTA on (n_channels, n) host arrays; `onset` must arrive pre-filled like the
 * reference wrappers do (lib.py:204,242,281) -- only the valid range is written. */
int qmb_sta_lta(qmb_ctx_t ctx, const double *signal, int32_t n_channels, int64_t n, int32_t nsta,
                int32_t nlta, int32_t position, double *onset);

/* --- detect: fused migrate + stack + per-sample max / argmax / normalised max -----------------
 * Equivalent to  find_max_coa(migrate(onsets, lut, fsmp, lsmp, available, threads), threads)
 * (scan.py:635-638) without ever materialising map4d.  n_samples = n_total - fsmp - lsmp.
 * qmb_detect_device runs on device-resident inputs and leaves results on the device
 * (qmb_detect_fetch copies them out); qmb_detect is the one-call host-buffer form.
 * With a communicator attached, every rank calls with its own LUT slab and identical onsets and
 * all ranks receive the global result.  kernel_ms (nullable, 4 floats): device time of
 * [scan kernel, merge+finalise, total, slow-path samples (count)] measured with CUDA events. */
int qmb_detect_device(qmb_ctx_t ctx, qmb_lut_t lut, qmb_onsets_t onsets, int32_t fsmp,
                      int32_t lsmp, int32_t available, float *kernel_ms);
int qmb_detect_fetch(qmb_ctx_t ctx, double *max_coa, double *max_norm_coa, int64_t *max_coa_idx,
                     int32_t n_samples);
int qmb_detect(qmb_ctx_t ctx, qmb_lut_t lut, const double *onsets, int32_t n_rows,
               int64_t n_total, int32_t fsmp, int32_t lsmp, int32_t available, double *max_coa,
               double *max_norm_coa, int64_t *max_coa_idx);

/* --- detect output stage ----------------------------------------------------
 * qmb_detect_fetch_quantised: the last detect result in the integer form ScanmSEED.append stores
 * (quakemigrate/io/scanmseed.py:102-130, _data2int :226-240), produced on the device:
 *   coa_q, coa_n_q  int32 (n): round(min(x, 21474) * 1e5), numpy rounding (half to even)
 *   ijk             int32 (n, 3): C-order unravel of max_coa_idx by node_count (lut.py:165-166)
 *   xyz             float64 (n, 3): ll_corner + ijk * node_spacing (lut.py:177), grid projection
 *   z_q             int32 (n): round(xyz[:, 2] * z_factor), z_factor = 1e3 * ucf (scanmseed.py:127)
 * Any output pointer may be null.  Longitude/latitude need the inverse map projection
 * (lut.py:236, pyproj): host work on `xyz`, outside this library. */
int qmb_detect_fetch_quantised(qmb_ctx_t ctx, const int32_t *node_count /* 3 */,
                               const double *ll_corner /* 3 */, const double *node_spacing /* 3 */,
                               double z_factor, int32_t *coa_q, int32_t *coa_n_q, int32_t *ijk,
                               double *xyz, int32_t *z_q, int32_t n_samples);

/* STEIM-1 / STEIM-2 data sections of miniSEED records (HOST code, no GPU involved): what
 * st.write(format="MSEED", encoding="STEIM2") produces through obspy/libmseed in the reference
 * (scanmseed.py:216-224; STEIM-1 is its fall-back when a difference exceeds 30 bits).
 * qmb_steim_encode packs as many of samples[0..n) as fit into n_frames 64-byte frames (`out`,
 * n_frames*64 bytes, zero-filled here), d0 = first difference (samples[0] minus the last sample of
 * the previous record, 0 for the first record of a trace); *n_packed = samples consumed.
 * Returns QMB_ERR_UNSUPPORTED when a difference does not fit the encoding.
 * qmb_steim_decode is the inverse for one record (first_diff nullable) and verifies the reverse
 * integration constant. */
int qmb_steim_encode(const int32_t *samples, int64_t n, int32_t d0, int32_t version, int32_t n_frames,
                     uint8_t *out, int64_t *n_packed);
int qmb_steim_decode(const uint8_t *frames, int32_t n_frames, int32_t version, int64_t n_samples,
                     int32_t *out, int32_t *first_diff);

/* --- locate: detect outputs + the marginal window of the 4-D map ------------------------------
 * After a detect pass over the whole span, re-migrates only samples [win0, win1) (end-exclusive,
 * quakemigrate/io/event.py:433-435) and returns
 *   coa_map  float64 (n_nodes_local): sum over the window of map4d (scan.py:720; the caller
 *            divides by its max, scan.py:721)
 *   map4d_win nullable float64 (n_nodes_local, win1-win0): the window slab itself. */
int qmb_locate_window(qmb_ctx_t ctx, qmb_lut_t lut, qmb_onsets_t onsets, int32_t fsmp,
                      int32_t lsmp, int32_t available, int32_t win0, int32_t win1,
                      double *coa_map, double *map4d_win);

/* --- locate, 3-D map stage ---------------------------------------------------
 * Streaming float64 kernels over the (nx, ny, nz) marginalised coalescence map; host pointers in
 * and out (the map is 8*N bytes).
 * qmb_gaufilt3d : QuakeScan._gaufilt3d (quakemigrate/signal/scan.py:1008-1043): Gaussian
 *                 smoothing in scipy fftconvolve(mode="same") semantics, mirrored and repeated,
 *                 each pass normalised by its maximum.
 * qmb_map3d_peak: nanargmax (first, C order), nanmax and nanmean of a map (scan.py:869,878,972).
 * qmb_covfit3d  : reductions of QuakeScan._covfit3d (scan.py:939-1005) over nodes with
 *                 map > thresh (win <= 0: whole grid; else +-(win-1)//2 nodes around the maximum):
 *                 out12[0..2] = weighted expectation (index * spacing), out12[3..11] = covariance. */
int qmb_gaufilt3d(qmb_ctx_t ctx, const double *map3d, int32_t nx, int32_t ny, int32_t nz, double sgm,
                  double *smoothed);
/* The same stage on maps that never leave the device (QuakeScan._calculate_location,
 * scan.py:696-733): qmb_locate_window_device leaves the marginal sum of the window in a map handle;
 * normalise = map / nanmax(map) (scan.py:721); gaufilt returns a new handle; peak_device returns
 * nanargmax / nanmax / nanmean (32 bytes); window copies a small box (the 5^3 spline and 7^3
 * Gaussian-fit windows, scan.py:1045-1070) to the host; covfit returns the 12 moments.  A located
 * event moves a few KB over PCIe instead of six copies of the 8*N-byte map. */
typedef struct qmb_map3d *qmb_map3d_t;
int qmb_locate_window_device(qmb_ctx_t ctx, qmb_lut_t lut, qmb_onsets_t onsets, int32_t fsmp,
                             int32_t lsmp, int32_t available, int32_t win0, int32_t win1,
                             qmb_map3d_t *coa_map);
int qmb_map3d_upload(qmb_ctx_t ctx, const double *map3d, int32_t nx, int32_t ny, int32_t nz,
                     qmb_map3d_t *out);
int qmb_map3d_fetch(qmb_map3d_t map, double *host);
int qmb_map3d_free(qmb_map3d_t map);
int qmb_map3d_normalise(qmb_ctx_t ctx, qmb_map3d_t map);
int qmb_map3d_gaufilt(qmb_ctx_t ctx, qmb_map3d_t map, double sgm, qmb_map3d_t *out);
int qmb_map3d_peak_device(qmb_ctx_t ctx, qmb_map3d_t map, int64_t *argmax, double *maxval, double *mean);
/* QuakeScan._splineloc's dense evaluation (scan.py:789-809): the cubic radial-basis interpolant of
 * an n^3 window (weights = solution of scipy's Rbf system, C order) on the upscale-times finer
 * grid, reduced to its first arg-max (flat C-order index into the (M, M, M) fine grid,
 * M = (n - 1) * upscale + 1) and value.  16 bytes back instead of 0.1-0.2 s of host time. */
int qmb_rbf_cubic_peak(qmb_ctx_t ctx, const double *weights, int32_t n, int32_t upscale,
                       int64_t *argmax, double *maxval);
int qmb_map3d_window(qmb_ctx_t ctx, qmb_map3d_t map, int32_t x0, int32_t x1, int32_t y0, int32_t y1,
                     int32_t z0, int32_t z1, double *host);
int qmb_map3d_covfit(qmb_ctx_t ctx, qmb_map3d_t map, const double *spacing, double thresh, int32_t win,
                     double *out12);
int qmb_map3d_peak(qmb_ctx_t ctx, const double *map3d, int64_t n, int64_t *argmax, double *maxval,
                   double *mean);
int qmb_covfit3d(qmb_ctx_t ctx, const double *map3d, int32_t nx, int32_t ny, int32_t nz,
                 const double *spacing, double thresh, int32_t win, double *out12);

/* Micro-benchmark of the pipe that bounds the scan kernel: sustained fp32 gather-adds per second
 * of conflict-free LDS.32 + FADD streams (13 warps per SM, like k_scan).  Used by bench.py to
 * report the scan against a MEASURED shared-memory ceiling next to the nominal 32/clk/SM. */
int qmb_microbench_smem_pipe(qmb_ctx_t ctx, int32_t iters, double *gather_adds_per_s);
/* exploratory form: vec = floats per lane per load (1, 2, 4), warps per SM in [1, 32] */
int qmb_microbench_smem_pipe_ex(qmb_ctx_t ctx, int32_t iters, int32_t vec, int32_t warps,
                                double *gather_adds_per_s);

#ifdef __cplusplus
}
#endif
#endif /* QMB200_H */
