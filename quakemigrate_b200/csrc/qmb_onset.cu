// qmb_onset.cu -- onset functions on the device.
//
//  * STA/LTA characteristic functions (classic/overlapping, centred, recursive), restating the
//    window definitions of quakemigrate/core/src/onsetlib.c:35-59, :79-108, :126-148 as
//    shuffle-based running sums over time samples (float64 throughout);
//  * the per-row recipe of STALTAOnset._onset (quakemigrate/signal/onsets/stalta.py:516-548):
//    transform -> per-channel STA/LTA -> RMS over channels -> clip(min_onset_value);
//  * the clip(0.01) + log of quakemigrate/core/lib.py:93-94, producing the float32 rows the scan
//    kernel gathers from and a float64 copy used for exact re-evaluation of arg-max candidates.
//
// The reference keeps two sliding sums per trace and updates them sample by sample; here every
// CTA builds a LOCAL inclusive prefix sum over (its output tile + the window halo) and forms the
// window sums as prefix differences.  Rounding therefore differs from the reference at the
// 1e-13 level on ordinary data (and is better conditioned after large-amplitude transients,
// where the reference's sliding sums keep the rounding residue of the transient forever).
#include <math.h>

#include <vector>

#include "qmb_internal.cuh"

namespace {

constexpr int SL_THREADS = 256;
constexpr int SL_TILE = 2048;  // outputs per CTA

__device__ __forceinline__ double apply_transform(double x, int transform) {
  if (transform == QMB_TRANSFORM_ENERGY) return x * x;
  if (transform == QMB_TRANSFORM_ABS) return fabs(x);
  return x;
}

// Inclusive prefix sum of `len` doubles held in shared memory (in place).
// Chunks of SL_THREADS elements: warp-shuffle scan, cross-warp offsets, running carry.
__device__ void block_prefix_inplace(double *buf, int len) {
  __shared__ double warp_tot[SL_THREADS / 32];
  __shared__ double carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0.0;
  __syncthreads();
  for (int base = 0; base < len; base += SL_THREADS) {
    const int i = base + tid;
    double v = (i < len) ? buf[i] : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += up;
    }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    double off = carry_s;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (i < len) buf[i] = v + off;
    __syncthreads();
    if (tid == SL_THREADS - 1) carry_s = v + off;
    __syncthreads();
  }
}

struct StaltaParams {
  const double *x;   // (C, stride_in)
  double *onset;     // (C, stride_out)
  long long n, stride_in, stride_out;
  int nsta, nlta, position, transform;
  // optional per-output-row overrides (row = blockIdx.y): source channel and window lengths
  const int *src_chan, *nsta_arr, *nlta_arr;
};

// classic : onset[i] = (sum x[i-nsta+1..i]) / (sum x[i-nlta+1..i]) * nlta/nsta,   i in [nlta-1, n)
// centred : onset[i] = (sum x[i+1..i+nsta]) / (sum x[i-nlta+1..i]) * nlta/nsta,   i in [nlta-1, n-nsta)
//           with lta <= 0 -> 1.0 for i >= nlta (onsetlib.c:102-106)
__global__ void __launch_bounds__(SL_THREADS) k_stalta_window(StaltaParams p) {
  extern __shared__ double sbuf[];  // prefix over x[lo .. hi)
  const int c = blockIdx.y;
  if (p.nsta_arr) {
    p.nsta = p.nsta_arr[c];
    p.nlta = p.nlta_arr[c];
    if (p.nsta < 1 || p.nsta > p.nlta || p.nlta > p.n) return;
  }
  const long long t0 = (long long)blockIdx.x * SL_TILE;
  const bool centred = p.position == QMB_STALTA_CENTRED;
  const long long first = p.nlta - 1;
  const long long last = centred ? p.n - p.nsta : p.n;  // exclusive
  long long o0 = max(t0, first), o1 = min(t0 + (long long)SL_TILE, last);
  if (o0 >= o1) return;
  // samples needed: [o0 - nlta + 1, o1 - 1 + (centred ? nsta : 0)]
  const long long lo = o0 - p.nlta + 1;
  const long long hi = o1 + (centred ? p.nsta : 0);  // exclusive
  const int len = (int)(hi - lo);
  const double *x = p.x + (size_t)(p.src_chan ? p.src_chan[c] : c) * p.stride_in;
  for (int i = threadIdx.x; i < len; i += SL_THREADS) sbuf[i + 1] = apply_transform(x[lo + i], p.transform);
  if (threadIdx.x == 0) sbuf[0] = 0.0;
  __syncthreads();
  block_prefix_inplace(sbuf + 1, len);
  __syncthreads();
  // sbuf[k] = sum of x[lo .. lo+k)
  const double frac = (double)p.nlta / (double)p.nsta;
  double *out = p.onset + (size_t)c * p.stride_out;
  for (long long i = o0 + threadIdx.x; i < o1; i += SL_THREADS) {
    const int k = (int)(i - lo);  // index of sample i within the buffer
    const double lta = sbuf[k + 1] - sbuf[k + 1 - p.nlta];
    double sta;
    if (centred) sta = sbuf[k + 1 + p.nsta] - sbuf[k + 1];
    else sta = sbuf[k + 1] - sbuf[k + 1 - p.nsta];
    double v = sta / lta * frac;
    if (centred && i >= p.nlta && !(lta > 0.0)) v = 1.0;
    out[i] = v;
  }
}

// recursive EMA form (onsetlib.c:126-148): one warp per channel, lanes own contiguous segments.
__global__ void __launch_bounds__(32) k_stalta_recursive(StaltaParams p) {
  const int c = blockIdx.x, lane = threadIdx.x;
  if (p.nsta_arr) {
    p.nsta = p.nsta_arr[c];
    p.nlta = p.nlta_arr[c];
  }
  const double *x = p.x + (size_t)(p.src_chan ? p.src_chan[c] : c) * p.stride_in;
  double *out = p.onset + (size_t)c * p.stride_out;
  const double a = 1.0 / (double)p.nsta, b = 1.0 / (double)p.nlta;
  // samples 1 .. n-1 are produced; split evenly
  const long long m = p.n - 1;
  if (m <= 0) return;
  const long long seg = (m + 31) / 32;
  const long long b0 = 1 + lane * seg, b1 = min(b0 + seg, p.n);
  // pass 1: segment response from zero state and decay factors
  double s = 0.0, l = 0.0, ds = 1.0, dl = 1.0;
  for (long long i = b0; i < b1; ++i) {
    const double v = apply_transform(x[i], p.transform);
    s = a * v + (1 - a) * s;
    l = b * v + (1 - b) * l;
    ds *= (1 - a);
    dl *= (1 - b);
  }
  // incoming state of each lane = fold of the previous lanes
  double in_s = 0.0, in_l = 0.0;
  for (int src = 0; src < 31; ++src) {
    const double ss = __shfl_sync(0xffffffffu, s, src), ll = __shfl_sync(0xffffffffu, l, src);
    const double dss = __shfl_sync(0xffffffffu, ds, src), dll = __shfl_sync(0xffffffffu, dl, src);
    // state after lane `src` given the state before it
    const double cur_s = __shfl_sync(0xffffffffu, in_s, src), cur_l = __shfl_sync(0xffffffffu, in_l, src);
    const double after_s = dss * cur_s + ss, after_l = dll * cur_l + ll;
    if (lane == src + 1) {
      in_s = after_s;
      in_l = after_l;
    }
  }
  // pass 2: replay with the right incoming state
  s = in_s;
  l = in_l;
  for (long long i = b0; i < b1; ++i) {
    const double v = apply_transform(x[i], p.transform);
    s = a * v + (1 - a) * s;
    l = b * v + (1 - b) * l;
    out[i] = s / l;
  }
  __syncwarp();
  if (p.nlta < p.n)
    for (long long i = lane; i < p.nlta; i += 32) out[i] = 1.0;
}

__global__ void k_fill(double *p, double v, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t step = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += step) p[i] = v;
}

// ---- linear onsets -> device log rows --------------------------------------------------------
struct PrepParams {
  const double *lin;  // (S, P) linear onsets, or nullptr when combining channels
  long long P, stride, front;
  float *log32;
  double *log64;
  unsigned long long *rowmax_bits;
  // channel-combine mode
  const double *chan;  // (n_pairs, chan_stride) STA/LTA per (row, channel) pair
  long long chan_stride;
  const int *row_ptr;
  double min_onset;
  double *lin_out;  // nullable (S, P)
  int already_log;  // input rows are log-onsets already (reference C ABI of migrate())
};

__global__ void __launch_bounds__(256) k_onset_prepare(PrepParams p) {
  const int s = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // position in padded row
  const long long src = i - p.front;
  double lv = 0.0;
  bool inside = src >= 0 && src < p.P;
  if (inside) {
    double lin;
    if (p.lin) {
      lin = p.lin[(size_t)s * p.P + src];
    } else {
      // RMS over the row's channels then clip (stalta.py:544-546)
      const int c0 = p.row_ptr[s], c1 = p.row_ptr[s + 1];
      double acc = 0.0;
      for (int k = c0; k < c1; ++k) {
        const double o = p.chan[(size_t)k * p.chan_stride + src];
        acc += o * o;
      }
      lin = sqrt(acc / (double)(c1 - c0));
      lin = fmax(lin, p.min_onset);
      if (p.lin_out) p.lin_out[(size_t)s * p.P + src] = lin;
    }
    lv = p.already_log ? lin : log(fmax(lin, 0.01));  // lib.py:93-94
  }
  if (i < p.stride) {
    p.log64[(size_t)s * p.stride + i] = lv;
    p.log32[(size_t)s * p.stride + i] = (float)lv;
  }
  // per-row max |log onset| (non-negative doubles order like their bit patterns)
  double m = inside ? fabs(lv) : 0.0;
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ double wm[8];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmax(m, wm[w]);
    atomicMax(&p.rowmax_bits[s], (unsigned long long)__double_as_longlong(m));
  }
}

// ---- 16-bit fixed-point rows for the detect scan -------------------------------------------------
// range[2 s], range[2 s + 1] = min, max of the log-onset row s (valid samples only)
__global__ void __launch_bounds__(256) k_onset_range(const double *log64, long long stride, long long front,
                                                     long long P, double *range) {
  const int s = blockIdx.x;
  const double *row = log64 + (size_t)s * stride + front;
  double mn = INFINITY, mx = -INFINITY;
  for (long long i = threadIdx.x; i < P; i += blockDim.x) {
    const double v = row[i];
    mn = fmin(mn, v);
    mx = fmax(mx, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __shared__ double smn[8], smx[8];
  if ((threadIdx.x & 31) == 0) {
    smn[threadIdx.x >> 5] = mn;
    smx[threadIdx.x >> 5] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      mn = fmin(mn, smn[w]);
      mx = fmax(mx, smx[w]);
    }
    range[2 * s] = mn;
    range[2 * s + 1] = mx;
  }
}

// u(i) = rint((x_i - qmin) / qstep) clamped to [0, 65535]; 0 outside the row.  Position pos holds
// u(pos - qfront) in the even stream and u(pos - qfront + 1) in the odd stream.
__global__ void __launch_bounds__(256) k_onset_quantise(const double *log64, long long stride, long long front,
                                                        long long P, uint16_t *q16, long long qstride,
                                                        long long qfront, double qmin, double inv_step,
                                                        double top) {
  const int s = blockIdx.y;
  const long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= qstride) return;
  const double *row = log64 + (size_t)s * stride + front;
  auto quant = [&](long long i) -> uint16_t {
    if (i < 0 || i >= P) return (uint16_t)0;
    const double u = rint((row[i] - qmin) * inv_step);
    return (uint16_t)fmin(fmax(u, 0.0), top);
  };
  const long long i = pos - qfront;
  q16[((size_t)2 * s + 0) * qstride + pos] = quant(i);
  q16[((size_t)2 * s + 1) * qstride + pos] = quant(i + 1);
}

}  // namespace

// Build the 16-bit streams of O (after its float64 rows exist).  One common grid for all rows:
// qmin = the smallest value present (so that the clip floor, the most frequent value of real
// onsets, is represented exactly), qstep = (max - min) / 65535.
// host copy of sum_rows max|log onset| (the float32 error bound of the scan needs it per call;
// taking it here keeps the scan itself free of host round trips)
static int onsets_rowmax_to_host(qmb_ctx *ctx, qmb_onsets *O, bool sync) {
  O->h_rowmax.resize((size_t)O->S);
  QMB_CUDA(cudaMemcpyAsync(O->h_rowmax.data(), O->rowmax.p, O->S * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (sync) QMB_CUDA(cudaStreamSynchronize(ctx->stream));
  return QMB_OK;
}
static void onsets_rowmax_sum(qmb_onsets *O) {
  double a = 0.0;
  for (double v : O->h_rowmax) a += v;
  O->rowmax_sum = a;
}

int qmb_onsets_quantise(qmb_ctx *ctx, qmb_onsets *O) {
  O->qstep = 0.0;
  if (O->S <= 0 || O->P <= 0) return QMB_OK;
  QMB_TRY(O->qrange.ensure((size_t)2 * O->S));
  k_onset_range<<<O->S, 256, 0, ctx->stream>>>(O->log64.p, O->stride, O->front, O->P, O->qrange.p);
  ctx->launches++;
  QMB_CUDA(cudaGetLastError());
  std::vector<double> range((size_t)2 * O->S);
  QMB_TRY(onsets_rowmax_to_host(ctx, O, false));
  QMB_CUDA(cudaMemcpyAsync(range.data(), O->qrange.p, range.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  QMB_CUDA(cudaStreamSynchronize(ctx->stream));
  onsets_rowmax_sum(O);
  double mn = range[0], mx = range[1];
  for (int s = 0; s < O->S; ++s) {
    if (range[2 * s] < mn) mn = range[2 * s];
    if (range[2 * s + 1] > mx) mx = range[2 * s + 1];
  }
  if (!(mn <= mx) || !isfinite(mn) || !isfinite(mx)) return QMB_OK;  // NaN/inf rows: float32 scan only
  // rows per carry-free group of the packed scan <-> bits per sample (G rows of (16 - log2 G)-bit
  // values add up inside 16 bits); rows beyond S up to a multiple of 8 are zero (they pad a group)
  O->qG = ctx->scan_qgroup;
  O->qbits = O->qG == 8 ? 13 : (O->qG == 4 ? 14 : (O->qG == 2 ? 15 : 16));
  O->qlevels = (1 << O->qbits) - 1;
  O->qmin = mn;
  O->qstep = mx > mn ? (mx - mn) / (double)O->qlevels : 1.0;
  O->qfront = QMB_ONSET_QFRONT;
  O->qstride = ((O->qfront + O->P + QMB_ONSET_TAIL + 7) / 8) * 8;
  const size_t S8 = ((size_t)O->S + 7) & ~(size_t)7;
  QMB_TRY(O->q16.ensure((size_t)2 * S8 * O->qstride));
  if (S8 > (size_t)O->S)
    QMB_CUDA(cudaMemsetAsync(O->q16.p + (size_t)2 * O->S * O->qstride, 0,
                             (size_t)2 * (S8 - O->S) * O->qstride * sizeof(uint16_t), ctx->stream));
  dim3 grid((unsigned)((O->qstride + 255) / 256), (unsigned)O->S);
  k_onset_quantise<<<grid, 256, 0, ctx->stream>>>(O->log64.p, O->stride, O->front, O->P, O->q16.p, O->qstride,
                                                  O->qfront, O->qmin, 1.0 / O->qstep, (double)O->qlevels);
  ctx->launches++;
  QMB_CUDA(cudaGetLastError());
  return QMB_OK;
}

int qmb_stalta_device(qmb_ctx *ctx, const double *d_signal, int32_t n_channels, int64_t n,
                      int64_t stride, int32_t nsta, int32_t nlta, int32_t position,
                      int32_t transform, double *d_onset) {
  return qmb_stalta_device_ex(ctx, d_signal, n_channels, n, stride, stride, nsta, nlta, position,
                              transform, d_onset, nullptr, nullptr, nullptr);
}

// General form: output row r reads channel src_chan[r] (identity when null) with windows
// nsta_arr[r] / nlta_arr[r] (scalars nsta/nlta = the maxima, used to size shared memory).
int qmb_stalta_device_ex(qmb_ctx *ctx, const double *d_signal, int32_t n_channels, int64_t n,
                         int64_t stride_in, int64_t stride_out, int32_t nsta, int32_t nlta,
                         int32_t position, int32_t transform, double *d_onset,
                         const int *d_src_chan, const int *d_nsta, const int *d_nlta) {
  if (nsta < 1 || nlta < 1) QMB_FAIL(QMB_ERR_ARG, "STA/LTA windows must be >= 1 sample");
  if (n_channels <= 0 || n <= 0) return QMB_OK;
  StaltaParams p;
  p.x = d_signal; p.onset = d_onset; p.n = n; p.stride_in = stride_in; p.stride_out = stride_out;
  p.nsta = nsta; p.nlta = nlta; p.position = position; p.transform = transform;
  p.src_chan = d_src_chan; p.nsta_arr = d_nsta; p.nlta_arr = d_nlta;
  if (position == QMB_STALTA_RECURSIVE) {
    k_stalta_recursive<<<n_channels, 32, 0, ctx->stream>>>(p);
    ctx->launches++;
    QMB_CUDA(cudaGetLastError());
    return QMB_OK;
  }
  if (!d_nsta && nsta > nlta)
    QMB_FAIL(QMB_ERR_ARG, "nsta (%d) > nlta (%d): the reference reads out of bounds here", nsta, nlta);
  if (!d_nsta && nlta > n) return QMB_OK;  // no valid output sample
  const size_t smem = (size_t)(SL_TILE + nlta + nsta + 2) * sizeof(double);
  if (smem > (size_t)QMB_SMEM_BUDGET)
    QMB_FAIL(QMB_ERR_UNSUPPORTED, "STA/LTA windows of %d+%d samples exceed the on-chip tile", nsta, nlta);
  QMB_CUDA(cudaFuncSetAttribute(k_stalta_window, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((n + SL_TILE - 1) / SL_TILE), (unsigned)n_channels);
  k_stalta_window<<<grid, SL_THREADS, smem, ctx->stream>>>(p);
  ctx->launches++;
  QMB_CUDA(cudaGetLastError());
  return QMB_OK;
}

int qmb_fill_f64(qmb_ctx *ctx, double *d, double v, size_t n) {
  if (!n) return QMB_OK;
  const int blocks = (int)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256);
  k_fill<<<blocks, 256, 0, ctx->stream>>>(d, v, n);
  ctx->launches++;
  QMB_CUDA(cudaGetLastError());
  return QMB_OK;
}

static int onsets_alloc(qmb_ctx *ctx, qmb_onsets *O) {
  O->front = QMB_ONSET_FRONT;
  O->stride = ((O->front + O->P + QMB_ONSET_TAIL + 3) / 4) * 4;
  QMB_TRY(O->log32.ensure((size_t)O->S * O->stride));
  QMB_TRY(O->log64.ensure((size_t)O->S * O->stride));
  QMB_TRY(O->rowmax.ensure((size_t)O->S));
  QMB_CUDA(cudaMemsetAsync(O->rowmax.p, 0, O->S * sizeof(double), ctx->stream));
  return QMB_OK;
}

int qmb_onsets_prepare(qmb_ctx *ctx, qmb_onsets *O, const double *d_linear, int already_log) {
  QMB_TRY(onsets_alloc(ctx, O));
  PrepParams p = {};
  p.already_log = already_log;
  p.lin = d_linear; p.P = O->P; p.stride = O->stride; p.front = O->front;
  p.log32 = O->log32.p; p.log64 = O->log64.p;
  p.rowmax_bits = (unsigned long long *)O->rowmax.p;
  dim3 grid((unsigned)((O->stride + 255) / 256), (unsigned)O->S);
  k_onset_prepare<<<grid, 256, 0, ctx->stream>>>(p);
  ctx->launches++;
  QMB_CUDA(cudaGetLastError());
  if (already_log) {  // rows of the reference-compatible migrate(): float32 window pass only
    O->qstep = 0.0;
    QMB_TRY(onsets_rowmax_to_host(ctx, O, true));
    onsets_rowmax_sum(O);
    return QMB_OK;
  }
  return qmb_onsets_quantise(ctx, O);
}

int qmb_onsets_combine(qmb_ctx *ctx, qmb_onsets *O, const double *d_chan, int64_t chan_stride,
                       const int *d_row_ptr, double min_onset, double *d_lin_out) {
  QMB_TRY(onsets_alloc(ctx, O));
  PrepParams p = {};
  p.lin = nullptr; p.P = O->P; p.stride = O->stride; p.front = O->front;
  p.log32 = O->log32.p; p.log64 = O->log64.p;
  p.rowmax_bits = (unsigned long long *)O->rowmax.p;
  p.chan = d_chan; p.chan_stride = chan_stride; p.row_ptr = d_row_ptr;
  p.min_onset = min_onset; p.lin_out = d_lin_out;
  dim3 grid((unsigned)((O->stride + 255) / 256), (unsigned)O->S);
  k_onset_prepare<<<grid, 256, 0, ctx->stream>>>(p);
  ctx->launches++;
  QMB_CUDA(cudaGetLastError());
  return qmb_onsets_quantise(ctx, O);
}
