// qmb_internal.cuh -- shared declarations of libqmb200 (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/qmb200.h"

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
void qmb_set_error(const char *fmt, ...);

#define QMB_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      qmb_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));     \
      return QMB_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define QMB_TRY(call)                                                                          \
  do {                                                                                         \
    int s__ = (call);                                                                          \
    if (s__ != QMB_OK) return s__;                                                             \
  } while (0)

#define QMB_FAIL(code, ...)                                                                    \
  do {                                                                                         \
    qmb_set_error(__VA_ARGS__);                                                                \
    return (code);                                                                             \
  } while (0)

// ---------------------------------------------------------------------------------------------
// geometry constants of the scan kernel
// ---------------------------------------------------------------------------------------------
// Scan-kernel variants: W consumer warps x NPW nodes per warp form a node tile of tx*ty*tz nodes;
// every lane carries K time samples, so a time tile is 32*K samples.
struct QmbVariant {
  int id, W, NPW, K, tx, ty, tz, CR;  // CR = onset rows staged per pipeline stage
};
constexpr int QMB_N_VARIANTS = 10;
constexpr QmbVariant QMB_VARIANTS[QMB_N_VARIANTS] = {
    {0, 16, 4, 8, 4, 4, 4, 16},  // 17 warps -> 96 regs/thread
    {1, 16, 8, 4, 4, 4, 8, 16},
    {2, 8, 8, 8, 4, 4, 4, 16},   // 9 warps -> 168 regs/thread
    {3, 8, 8, 4, 4, 4, 4, 16},
    {4, 12, 4, 8, 4, 4, 3, 16},  // 13 warps -> 128 regs/thread
    {5, 12, 8, 4, 4, 4, 6, 16},
    {6, 8, 4, 8, 4, 4, 2, 16},
    {7, 12, 4, 8, 4, 4, 3, 32},  // as 4, 32 rows per stage: half the stage hand-offs
    {8, 15, 4, 8, 5, 4, 3, 32},  // 16 warps in all -> still 128 regs/thread, 60-node tiles
    {9, 15, 4, 8, 5, 4, 3, 16},
};
constexpr int QMB_DEFAULT_VARIANT = 7;  // measured best on B200 (scripts/sweep_variants.sh)
constexpr int QMB_DEFAULT_CHUNK_ROWS = 16;  // onset rows (station x phase) staged per pipeline stage
constexpr int QMB_SMEM_BUDGET = 227 * 1024;
constexpr int QMB_DEFAULT_QGROUP = 2;  // 15-bit rows, two rows per carry-free group (measured fastest)

// ---------------------------------------------------------------------------------------------
// device buffer helper (RAII-free; owned by the structs below)
// ---------------------------------------------------------------------------------------------
// Recycling of large device buffers (qmb_api.cu): the handles a detect / locate loop creates and
// frees every timestep or event (trace blocks, onset rows, maps) have the same sizes every time, and
// cudaMalloc / cudaFree of hundreds of MB cost milliseconds to tens of milliseconds each and
// serialise against every other driver call.  A freed buffer of >= 1 MiB is kept (after a device
// synchronise, which cudaFree implies as well, so it carries no pending work) and handed to the
// next request of exactly that size on that device.  Bounded by QMB_POOL_MB (default 8192; 0: off)
// and 32 entries; qmb_ctx_destroy returns the device's entries.
void *qmb_pool_take(size_t bytes);
bool qmb_pool_give(void *p, size_t bytes);
void qmb_pool_clear(int device);

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  bool pooled = false;  // recycle through qmb_pool_* (handle-owned buffers) instead of cudaFree
  static DevBuf Pooled() {
    DevBuf b;
    b.pooled = true;
    return b;
  }
  int ensure(size_t count) {
    if (count <= n && p) return QMB_OK;
    release();
    if (count == 0) return QMB_OK;
    if (pooled) p = (T *)qmb_pool_take(count * sizeof(T));
    if (!p) QMB_CUDA(cudaMalloc((void **)&p, count * sizeof(T)));
    n = count;
    return QMB_OK;
  }
  void release() {
    if (p && !(pooled && qmb_pool_give(p, n * sizeof(T)))) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  size_t bytes() const { return n * sizeof(T); }
};

// ---------------------------------------------------------------------------------------------
// handles
// ---------------------------------------------------------------------------------------------
struct qmb_ctx {
  int device = 0;
  int sm_count = 148;
  int variant = QMB_DEFAULT_VARIANT;  // index into QMB_VARIANTS used for tables uploaded next
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_user[2] = {nullptr, nullptr};  // qmb_ctx_timer_start / stop
  int64_t launches = 0;
  int scan_qgroup = QMB_DEFAULT_QGROUP;  // rows per carry-free group of the fixed-point scan (1, 2, 4, 8)
  int scan_fp32 = 0;      // detect scan operands: 0 automatic (16-bit fixed point when accurate enough), 1 float32, 2 16-bit
  int last_scan_q16 = 0;  // what the last detect scan actually used

  // NCCL
  void *nccl_comm = nullptr;
  int world = 1, rank = 0;

  // scan scratch (grown on demand)
  DevBuf<float> pm1, pm2, pm3;
  DevBuf<uint32_t> pi1, pi2;
  DevBuf<double> psum;
  // merged per-sample state, packed [sum f64][m1 m2 m3 f32][i1 i2 u32] x Tpad (28 bytes per sample):
  // st_loc = this rank's fold of its partials (also the all-gather send buffer), st_glob = the
  // fold over ranks (world > 1), st_grp = the 32 group folds of the two-stage merge, st_chk = the
  // fold of a float32 check window (node-sum verification of the 16-bit scan)
  DevBuf<unsigned char> st_loc, st_glob, st_grp, st_chk;
  DevBuf<double> ex;             // [2][T] float64 stacks of the two leading candidates
  DevBuf<unsigned char> gather;  // all-gather receive staging
  // results
  DevBuf<double> r_max, r_norm;
  DevBuf<long long> r_idx;
  DevBuf<unsigned char> slow_flag;  // [T padded to 16] 1 = sample needs the all-node float64 re-scan
  DevBuf<int> slow_list;            // [0] = listed count, [1..] = samples in ascending order
  DevBuf<double> slow_val;
  DevBuf<unsigned long long> slow_idx;
  DevBuf<unsigned int> status;      // device-side counters of the last scan (QMB_ST_*)
  unsigned int *h_status = nullptr; // pinned mirror
  qmb_scan_info_t info = {};        // what the last detect scan did (qmb_ctx_last_scan_info)
  DevBuf<unsigned char> flush;  // L2 eviction scratch (benchmarks)
  DevBuf<double> stage_lin;     // H2D staging of linear onsets (reused across calls)
  DevBuf<int32_t> q_i32;        // output stage: quantised channels + unravelled indices
  DevBuf<double> q_f64;         // output stage: grid coordinates
  DevBuf<double> map_scratch;   // locate 3-D map stage: convolution ping-pong, reductions, taps
  DevBuf<double> pp_scratch;    // pre-processing: stats, matrix powers, chunk/tile states, pass-1 output
  DevBuf<int> pp_sel;           // pre-processing: selected channel indices
  // Hilbert envelopes (own Bluestein + Stockham transforms, qmb_envelope.cu)
  int64_t fft_n = 0;          // trace length the cached chirp / chirp filter belong to
  DevBuf<double> fft_buf;     // two complex work buffers of the envelope transform
  DevBuf<double> fft_chirp;   // [w (n)][transformed chirp filter (M)][scratch (M)] complex
  struct qmb_onsets *tmp_onsets = nullptr;  // onset rows of the one-call qmb_detect (reused)
  int32_t last_n_samples = 0;
};

// slots of qmb_ctx::status
enum { QMB_ST_SLOW_TOTAL = 0, QMB_ST_TWO_CAND = 1, QMB_ST_CHECK_ERR = 2, QMB_ST_CHECK_BAD = 3,
       QMB_ST_SLOW_LISTED = 4, QMB_ST_AGREE = 5, QMB_ST_COUNT = 8 };

struct qmb_lut {
  qmb_ctx *ctx = nullptr;
  int device = 0;  // copied from ctx: handles may outlive their context
  int32_t nxl = 0, ny = 0, nz = 0, S = 0, Spad = 0, chunk_rows = QMB_DEFAULT_CHUNK_ROWS;
  // a slab is a contiguous range of the C-order node list in whole z-columns: it covers nxl x-planes,
  // skips the first y_first rows of the first one and keeps only y_last rows of the last (0: all)
  int32_t y_first = 0, y_last = 0;
  int64_t node_offset = 0, n_global = 0, n_local = 0;
  QmbVariant v = QMB_VARIANTS[QMB_DEFAULT_VARIANT];  // variant this table was tiled for
  int tx = 4, ty = 4, tz = 4, NB = 64;
  int ntx = 0, nty = 0, ntz = 0, n_tiles = 0;
  int32_t max_tau = 0, max_span = 0;
  double dup_frac = 0.0;  // fraction of neighbouring nodes (inside a warp's group) sharing the row's offset
  DevBuf<uint16_t> tau16;   // [n_tiles][Spad][NB]   tau - tmin(tile,row)
  DevBuf<int32_t> tmin;     // [n_tiles][Spad]       multiple of 4
  DevBuf<uint16_t> span;    // [n_tiles][Spad]       max(tau16) within the tile
  DevBuf<uint32_t> nodeidx; // [n_tiles][NB]         global flat node index or 0xFFFFFFFF (pad)
  DevBuf<int> boxes;        // [n_tiles][9]          tile origin, shape, region end
  DevBuf<uint32_t> slotmap; // [n_local]             tile * NB + slot of each local node
  // 16-bit scan: word offset of (tile, row, slot) inside the staged row = (tau16 >> 1), plus
  // RSq / 2 when tau16 is odd (the pair then comes from the "odd" stream staged behind the
  // "even" one); RSq = elements per staged stream
  DevBuf<uint16_t> tauq16;  // [n_tiles][Spad][NB]
  int32_t RSq = 0;
  // packed scan (groups of qG rows): BYTE offset of the pair inside the stage, row position included,
  // laid out [n_tiles][Spad / qG][NB][qG] so that one vector load fetches a node's offsets of a group
  DevBuf<uint32_t> tauq32;
  DevBuf<uint16_t> tauq16p;  // qG == 2 and the stage's rows fit 64 KiB: the same offsets in 16 bits
  int32_t qG = 1;
};

struct qmb_ttgrids {  // every (station, phase) travel-time grid, float64 seconds, on the device
  qmb_ctx *ctx = nullptr;
  int device = 0;
  int32_t n_grids = 0, nx = 0, ny = 0, nz = 0;
  DevBuf<double> tt;  // [n_grids][nx*ny*nz]
};

struct qmb_onsets {
  qmb_ctx *ctx = nullptr;
  int device = 0;
  int32_t S = 0;
  int64_t P = 0;          // samples per row as seen by the caller (fsmp + T + lsmp)
  int64_t stride = 0;     // device row stride in elements (multiple of 4)
  int64_t front = 0;      // leading pad elements (so that front + fsmp can be 16B-aligned)
  DevBuf<float> log32 = DevBuf<float>::Pooled();     // [S][stride]
  DevBuf<double> log64 = DevBuf<double>::Pooled();   // [S][stride]
  DevBuf<double> rowmax;  // [S] max |log onset| per row (for the fp32 error bound)
  double rowmax_sum = 0.0;  // host copy of sum(rowmax), taken when the rows are prepared
  std::vector<double> h_rowmax;
  // 16-bit fixed-point copy for the detect scan: u = rint((x - qmin) / qstep) in [0, 65535].
  // Two streams per row over a common position axis pos = qfront + sample:
  //   stream 0 ("even"): q16[(2 s + 0) * qstride + pos] = u(sample)
  //   stream 1 ("odd") : q16[(2 s + 1) * qstride + pos] = u(sample + 1)
  // so that the pair (u(i), u(i + 1)) is one aligned 32-bit word for every i.
  DevBuf<uint16_t> q16 = DevBuf<uint16_t>::Pooled();
  DevBuf<double> qrange;  // [S][2] min / max of each row
  int64_t qstride = 0, qfront = 0;
  int32_t qG = 1, qbits = 16, qlevels = 65535;  // rows per carry-free group <-> bits per sample
  double qmin = 0.0, qstep = 0.0;  // qstep == 0: no quantised copy (e.g. rows of the T1 migrate())
};

struct qmb_map3d {  // float64 (nx, ny, nz) map resident on the device (locate 3-D stage)
  qmb_ctx *ctx = nullptr;
  int device = 0;
  int32_t nx = 0, ny = 0, nz = 0;
  DevBuf<double> v = DevBuf<double>::Pooled();
};

struct qmb_waveforms {  // device-resident float64 traces, (C, n) C-order
  qmb_ctx *ctx = nullptr;
  int device = 0;
  int32_t C = 0;
  int64_t n = 0;
  DevBuf<double> data = DevBuf<double>::Pooled();
};

constexpr int64_t QMB_ONSET_FRONT = 4;     // leading pad (elements); alignment slack
constexpr int64_t QMB_ONSET_TAIL = 1024;   // trailing zero pad: >= time tile + max tile span + 8
constexpr int64_t QMB_ONSET_QFRONT = 8;    // leading pad of the 16-bit streams (elements; 16 bytes)

// ---------------------------------------------------------------------------------------------
// cross-file entry points
// ---------------------------------------------------------------------------------------------
struct QmbServeSource {  // float64 travel-time grids resident on the device (qmb_ttgrids)
  const double *tt;
  const int *rowsel;
  double sr;
  long long stride, node0;
};
int qmb_lut_rsq(const qmb_lut *L);
int qmb_lut_build(qmb_ctx *ctx, qmb_lut *L, const int32_t *d_raw, const QmbServeSource *src = nullptr);
int qmb_onsets_prepare(qmb_ctx *ctx, qmb_onsets *O, const double *d_linear /* (S,P) */, int already_log);
int qmb_onsets_quantise(qmb_ctx *ctx, qmb_onsets *O);
int qmb_scan_run(qmb_ctx *ctx, qmb_lut *L, qmb_onsets *O, int32_t fsmp, int32_t lsmp,
                 int32_t available, float *kernel_ms);
int qmb_window_run(qmb_ctx *ctx, qmb_lut *L, qmb_onsets *O, int32_t fsmp, int32_t lsmp,
                   int32_t available, int32_t win0, int32_t win1, double *d_coa_map,
                   double *d_map4d_win, int accumulate);
int qmb_stalta_device(qmb_ctx *ctx, const double *d_signal, int32_t n_channels, int64_t n,
                      int64_t stride, int32_t nsta, int32_t nlta, int32_t position,
                      int32_t transform, double *d_onset);
int qmb_stalta_device_ex(qmb_ctx *ctx, const double *d_signal, int32_t n_channels, int64_t n,
                         int64_t stride_in, int64_t stride_out, int32_t nsta, int32_t nlta,
                         int32_t position, int32_t transform, double *d_onset,
                         const int *d_src_chan, const int *d_nsta, const int *d_nlta);
int qmb_fill_f64(qmb_ctx *ctx, double *d, double v, size_t n);
int qmb_onsets_combine(qmb_ctx *ctx, qmb_onsets *O, const double *d_chan, int64_t chan_stride,
                       const int *d_row_ptr, double min_onset, double *d_lin_out);
int qmb_nccl_allgather(qmb_ctx *ctx, const void *send, void *recv, size_t bytes_per_rank);
int qmb_nccl_allreduce_sum_f64(qmb_ctx *ctx, double *buf, size_t count);
int qmb_nccl_allreduce_u32(qmb_ctx *ctx, unsigned int *buf, size_t count, int take_max);
int qmb_preprocess_device(qmb_ctx *ctx, const double *d_x, int64_t x_stride, const int *d_chan_in,
                          int32_t n_sel, int64_t n, const double *sos, int32_t nsec, int32_t detrend,
                          double taper_fraction, int32_t zerophase, int32_t decimate, double *d_out,
                          int64_t out_stride, int32_t out_chan0, float *kernel_ms);
int qmb_envelope_device(qmb_ctx *ctx, const double *d_x, int64_t x_stride, int32_t n_channels, int64_t n,
                        double *d_out, int64_t out_stride);
void qmb_fft_release(qmb_ctx *ctx);
qmb_ctx *qmb_default_ctx();
int qmb_map3d_new(qmb_ctx *ctx, int32_t nx, int32_t ny, int32_t nz, qmb_map3d **out);
