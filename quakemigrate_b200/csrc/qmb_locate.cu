// qmb_locate.cu -- the 3-D map stage that follows the marginal window in locate()
// (SURVEY.md section 8f rank 1): streaming float64 kernels over the (nx, ny, nz) marginalised
// coalescence map.  They restate, on the device,
//   * QuakeScan._gaufilt3d  (quakemigrate/signal/scan.py:1008-1043): convolve with a full-size
//     Gaussian (util.gaussian_3d, util.py:76-116) in scipy's mode="same" sense, mirror, normalise
//     by the maximum, convolve again, mirror, normalise;
//   * the reductions of QuakeScan._covfit3d (scan.py:939-1005): weighted first moments and
//     central second moments of the nodes above a threshold (optionally inside a window);
//   * nanargmax / nanmean of a map (scan.py:869, :878, :972).
// The small dense algebra that follows (pinv of a 343x10 system, 3x3 eigen-problem, the cubic RBF
// on a 5^3 window) stays on the host (quakemigrate_b200/locate.py).
//
// scipy.signal.fftconvolve(a, g, mode="same") with g the same shape as a returns, per axis of
// length n,  out[i] = sum_j a[j] * g[i - j + (n-1)//2];  with g[m] = exp(-(m-(n-1)/2)^2 / (2 s^2))
// this is a Gaussian centred half a sample off when n is even -- the reference mirrors and
// repeats the convolution to cancel that shift, and so do we.  The Gaussian is separable, so
// three 1-D passes give the same result as the 3-D FFT product (to rounding, ~1e-16 of the peak).
#include <math.h>

#include <vector>

#include "qmb_internal.cuh"

namespace {

struct Dims {
  int nx, ny, nz;
};

// one 1-D pass along `axis` (0:x, 1:y, 2:z): out[i] = sum_m w[m] * in[i + c - m]
__global__ void __launch_bounds__(256) k_conv_axis(const double *in, double *out, Dims d, int axis,
                                                   const double *w, int m_lo, int m_hi, int c) {
  const long long n_total = (long long)d.nx * d.ny * d.nz;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_total) return;
  const int iz = (int)(idx % d.nz), iy = (int)((idx / d.nz) % d.ny), ix = (int)(idx / ((long long)d.nz * d.ny));
  const int n = axis == 0 ? d.nx : (axis == 1 ? d.ny : d.nz);
  const int i = axis == 0 ? ix : (axis == 1 ? iy : iz);
  const long long stride = axis == 0 ? (long long)d.ny * d.nz : (axis == 1 ? d.nz : 1);
  const long long base = idx - (long long)i * stride;
  double acc = 0.0;
  for (int m = m_lo; m <= m_hi; ++m) {
    const int j = i + c - m;
    if (j >= 0 && j < n) acc += w[m] * in[base + (long long)j * stride];
  }
  out[idx] = acc;
}

// per-block (max ignoring NaN, first index of it, sum of non-NaN, count of non-NaN)
struct PeakPartial {
  double mx, sum;
  long long idx, cnt;
};

__global__ void __launch_bounds__(256) k_peak_partial(const double *map, long long n, PeakPartial *part) {
  double mx = -INFINITY, sum = 0.0;
  long long idx = -1, cnt = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = map[i];
    if (v != v) continue;
    sum += v;
    ++cnt;
    if (v > mx || idx < 0) { mx = v; idx = i; }  // ascending i within a thread: first maximum
  }
  __shared__ double s_mx[256], s_sum[256];
  __shared__ long long s_idx[256], s_cnt[256];
  const int t = threadIdx.x;
  s_mx[t] = mx; s_sum[t] = sum; s_idx[t] = idx; s_cnt[t] = cnt;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) {
      const double v = s_mx[t + o];
      const long long j = s_idx[t + o];
      if (j >= 0 && (s_idx[t] < 0 || v > s_mx[t] || (v == s_mx[t] && j < s_idx[t]))) { s_mx[t] = v; s_idx[t] = j; }
      s_sum[t] += s_sum[t + o];
      s_cnt[t] += s_cnt[t + o];
    }
    __syncthreads();
  }
  if (t == 0) part[blockIdx.x] = {s_mx[0], s_sum[0], s_idx[0], s_cnt[0]};
}

__global__ void k_peak_final(const PeakPartial *part, int n_parts, PeakPartial *out) {
  PeakPartial r = {-INFINITY, 0.0, -1, 0};
  for (int q = 0; q < n_parts; ++q) {  // single thread: deterministic order
    const PeakPartial p = part[q];
    r.sum += p.sum;
    r.cnt += p.cnt;
    if (p.idx >= 0 && (r.idx < 0 || p.mx > r.mx || (p.mx == r.mx && p.idx < r.idx))) { r.mx = p.mx; r.idx = p.idx; }
  }
  *out = r;
}

// out[mirror(i)] = in[i] * scale   (reverse all three axes == reverse the flat C-order index)
__global__ void __launch_bounds__(256) k_mirror_scale(const double *in, double *out, long long n, const PeakPartial *peak) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[n - 1 - i] = in[i] / peak->mx;
}

// ---- covariance fit reductions -------------------------------------------------------------------
struct CovParams {
  const double *map;
  Dims d;
  double sx, sy, sz;  // node spacing
  double thresh;
  int use_win, x1, x2, y1, y2, z1, z2;  // window [x1,x2) ... when use_win
  double xe, ye, ze;  // pass 2 only
};

template <int PASS>
__global__ void __launch_bounds__(256) k_cov_partial(CovParams p, double *part /* [blocks][8] */) {
  double a[7] = {0, 0, 0, 0, 0, 0, 0};
  const long long n = (long long)p.d.nx * p.d.ny * p.d.nz;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double w = p.map[i];
    if (!(w > p.thresh)) continue;  // also drops NaN
    const int iz = (int)(i % p.d.nz), iy = (int)((i / p.d.nz) % p.d.ny), ix = (int)(i / ((long long)p.d.nz * p.d.ny));
    if (p.use_win && (ix < p.x1 || ix >= p.x2 || iy < p.y1 || iy >= p.y2 || iz < p.z1 || iz >= p.z2)) continue;
    const double x = ix * p.sx, y = iy * p.sy, z = iz * p.sz;
    if (PASS == 1) {
      a[0] += w; a[1] += w * x; a[2] += w * y; a[3] += w * z;
    } else {
      const double dx = x - p.xe, dy = y - p.ye, dz = z - p.ze;
      a[0] += w * dx * dx; a[1] += w * dy * dy; a[2] += w * dz * dz;
      a[3] += w * dx * dy; a[4] += w * dx * dz; a[5] += w * dy * dz;
    }
  }
  __shared__ double s[7][256];
  for (int k = 0; k < 7; ++k) s[k][threadIdx.x] = a[k];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o)
      for (int k = 0; k < 7; ++k) s[k][threadIdx.x] += s[k][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0)
    for (int k = 0; k < 7; ++k) part[(size_t)blockIdx.x * 8 + k] = s[k][0];
}

__global__ void k_cov_final(const double *part, int n_parts, double *out /* [8] */) {
  double a[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int q = 0; q < n_parts; ++q)
    for (int k = 0; k < 7; ++k) a[k] += part[(size_t)q * 8 + k];
  for (int k = 0; k < 7; ++k) out[k] = a[k];
}

int gaussian_taps(int n, double sgm, std::vector<double> &w, int &m_lo, int &m_hi) {
  w.assign(n, 0.0);
  const double c = (n - 1) / 2.0;
  m_lo = n;
  m_hi = -1;
  for (int m = 0; m < n; ++m) {
    const double r = m - c;
    const double v = exp(-(r * r) / (2.0 * sgm * sgm));  // util.py:110-114
    w[m] = v;
    // taps below 1e-22 of the central one cannot change a float64 result (the reference's FFT
    // product is itself only good to ~1e-16 of the peak): skip them
    if (v > 1e-22) {
      if (m < m_lo) m_lo = m;
      m_hi = m;
    }
  }
  return QMB_OK;
}

constexpr int RED_BLOCKS = 592;

// scratch of the map stage, carved out of one context-owned buffer: [a n][b n][part][peak][cov 8*(RED+1)][taps]
struct MapScratch {
  double *a, *b, *cov_part, *cov_out, *taps;
  PeakPartial *part, *peak;
};
int map_scratch(qmb_ctx *ctx, long long n, int n_taps, MapScratch &s) {
  const size_t pp = (sizeof(PeakPartial) * (RED_BLOCKS + 1) + 7) / 8;
  const size_t total = 2 * (size_t)n + pp + 8 * (RED_BLOCKS + 1) + (size_t)n_taps + 8;
  QMB_TRY(ctx->map_scratch.ensure(total));
  double *q = ctx->map_scratch.p;
  s.a = q; q += n;
  s.b = q; q += n;
  s.part = (PeakPartial *)q; s.peak = s.part + RED_BLOCKS; q += pp;
  s.cov_part = q; s.cov_out = q + 8 * RED_BLOCKS; q += 8 * (RED_BLOCKS + 1);
  s.taps = q;
  return QMB_OK;
}

int peak_of(qmb_ctx *ctx, const double *d_map, long long n, PeakPartial *d_part, PeakPartial *d_peak) {
  k_peak_partial<<<RED_BLOCKS, 256, 0, ctx->stream>>>(d_map, n, d_part);
  k_peak_final<<<1, 1, 0, ctx->stream>>>(d_part, RED_BLOCKS, d_peak);
  ctx->launches += 2;
  QMB_CUDA(cudaGetLastError());
  return QMB_OK;
}

// three separable passes: src -> dst (tmp is clobbered)
int conv3(qmb_ctx *ctx, const double *src, double *dst, double *tmp, Dims d, const double *d_w, const int *lo,
          const int *hi, const int *off) {
  const long long n = (long long)d.nx * d.ny * d.nz;
  const int blocks = (int)((n + 255) / 256);
  const int ns[3] = {d.nx, d.ny, d.nz};
  const double *in = src;
  double *bufs[3] = {dst, tmp, dst};
  for (int axis = 0; axis < 3; ++axis) {
    k_conv_axis<<<blocks, 256, 0, ctx->stream>>>(in, bufs[axis], d, axis, d_w + off[axis], lo[axis], hi[axis],
                                                 (ns[axis] - 1) / 2);
    ctx->launches++;
    in = bufs[axis];
  }
  QMB_CUDA(cudaGetLastError());
  return QMB_OK;
}

__global__ void __launch_bounds__(256) k_scale_by_peak(double *map, long long n, const PeakPartial *peak) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) map[i] = map[i] / peak->mx;
}

// copy the box [lo, hi) of a map into a dense (hi-lo) C-order block
__global__ void k_map_window(const double *map, Dims d, int x0, int y0, int z0, int wx, int wy, int wz, double *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= wx * wy * wz) return;
  const int iz = i % wz, iy = (i / wz) % wy, ix = i / (wz * wy);
  out[i] = map[((long long)(x0 + ix) * d.ny + (y0 + iy)) * d.nz + (z0 + iz)];
}

// Cubic radial-basis interpolant of an n^3 window evaluated on the `up`-times finer grid
// (QuakeScan._splineloc, scan.py:776-804: scipy's Rbf(function="cubic") called on
// meshgrid(linspace(0, n-1, (n-1)*up+1))).  Output point (a, b, c) and centre (a', b', c'), both in
// C order, sit at the coordinates (b, a, c) -- numpy.meshgrid's default 'xy' indexing -- which only
// fixes the order of the three squares inside the distance.  out[p] = sum_j w[j] * r_pj^3.
__global__ void __launch_bounds__(128) k_rbf_cubic(const double *w, int n, int M, double step, double *out) {
  extern __shared__ double sw[];
  const int n3 = n * n * n;
  for (int j = threadIdx.x; j < n3; j += blockDim.x) sw[j] = w[j];
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= M * M * M) return;
  const int c = p % M, b = (p / M) % M, a = p / (M * M);
  // numpy.linspace: arange(num) * step, the last sample replaced by the stop value
  const double x = b == M - 1 ? (double)(n - 1) : b * step;
  const double y = a == M - 1 ? (double)(n - 1) : a * step;
  const double z = c == M - 1 ? (double)(n - 1) : c * step;
  double acc = 0.0;
  int j = 0;
  for (int ca = 0; ca < n; ++ca)
    for (int cb = 0; cb < n; ++cb)
      for (int cc = 0; cc < n; ++cc, ++j) {
        const double dx = x - cb, dy = y - ca, dz = z - cc;
        const double r = sqrt(dx * dx + dy * dy + dz * dz);
        acc += r * r * r * sw[j];
      }
  out[p] = acc;
}

int check_map(qmb_ctx *ctx, qmb_map3d *m) {
  if (!ctx || !m) QMB_FAIL(QMB_ERR_ARG, "null argument");
  if (m->ctx != ctx) QMB_FAIL(QMB_ERR_ARG, "map belongs to a different context");
  QMB_CUDA(cudaSetDevice(ctx->device));
  return QMB_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// device-resident maps
// ---------------------------------------------------------------------------------------------
int qmb_map3d_new(qmb_ctx *ctx, int32_t nx, int32_t ny, int32_t nz, qmb_map3d **out) {
  *out = nullptr;
  if (nx <= 0 || ny <= 0 || nz <= 0) QMB_FAIL(QMB_ERR_ARG, "bad map shape");
  qmb_map3d *m = new qmb_map3d();
  m->ctx = ctx; m->device = ctx->device; m->nx = nx; m->ny = ny; m->nz = nz;
  const int rc = m->v.ensure((size_t)nx * ny * nz);
  if (rc != QMB_OK) {
    delete m;
    return rc;
  }
  *out = m;
  return QMB_OK;
}

extern "C" int qmb_map3d_upload(qmb_ctx_t ctx, const double *map3d, int32_t nx, int32_t ny, int32_t nz,
                                qmb_map3d_t *out) {
  if (!ctx || !map3d || !out) QMB_FAIL(QMB_ERR_ARG, "null argument");
  QMB_CUDA(cudaSetDevice(ctx->device));
  qmb_map3d *m = nullptr;
  QMB_TRY(qmb_map3d_new(ctx, nx, ny, nz, &m));
  cudaError_t e = cudaMemcpyAsync(m->v.p, map3d, m->v.bytes(), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    qmb_map3d_free(m);
    QMB_FAIL(QMB_ERR_CUDA, "qmb_map3d_upload: %s", cudaGetErrorString(e));
  }
  *out = m;
  return QMB_OK;
}

extern "C" int qmb_map3d_fetch(qmb_map3d_t m, double *host) {
  if (!m || !host) QMB_FAIL(QMB_ERR_ARG, "null argument");
  QMB_CUDA(cudaSetDevice(m->device));
  QMB_CUDA(cudaMemcpyAsync(host, m->v.p, m->v.bytes(), cudaMemcpyDeviceToHost, m->ctx->stream));
  QMB_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return QMB_OK;
}

extern "C" int qmb_map3d_free(qmb_map3d_t m) {
  if (!m) return QMB_OK;
  cudaSetDevice(m->device);
  m->v.release();
  delete m;
  return QMB_OK;
}

// nanargmax (first, C order), nanmax, nanmean (scan.py:869,878,972); 32 bytes come back
extern "C" int qmb_map3d_peak_device(qmb_ctx_t ctx, qmb_map3d_t m, int64_t *argmax, double *maxval, double *mean) {
  QMB_TRY(check_map(ctx, m));
  const long long n = (long long)m->nx * m->ny * m->nz;
  MapScratch s;
  QMB_TRY(map_scratch(ctx, 0, 0, s));
  QMB_TRY(peak_of(ctx, m->v.p, n, s.part, s.peak));
  PeakPartial h;
  QMB_CUDA(cudaMemcpyAsync(&h, s.peak, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  QMB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (argmax) *argmax = h.idx;
  if (maxval) *maxval = h.mx;
  if (mean) *mean = h.cnt > 0 ? h.sum / (double)h.cnt : NAN;
  return QMB_OK;
}

// Peak of the cubic RBF interpolant on the upscaled window grid (scan.py:801-809): `weights` are the
// n^3 interpolation weights (the solution of scipy's Rbf system, C order of the window); returns the
// first arg-max (C order of the (M, M, M) fine grid, M = (n - 1) * upscale + 1) and its value.
extern "C" int qmb_rbf_cubic_peak(qmb_ctx_t ctx, const double *weights, int32_t n, int32_t upscale, int64_t *argmax,
                                  double *maxval) {
  if (!ctx || !weights || !argmax) QMB_FAIL(QMB_ERR_ARG, "null argument");
  if (n < 2 || n > 16 || upscale < 1 || upscale > 64) QMB_FAIL(QMB_ERR_ARG, "bad window (%d) or upscale (%d)", n, upscale);
  QMB_CUDA(cudaSetDevice(ctx->device));
  const int M = (n - 1) * upscale + 1, n3 = n * n * n;
  const long long total = (long long)M * M * M;
  MapScratch s;
  QMB_TRY(map_scratch(ctx, total, n3, s));
  QMB_CUDA(cudaMemcpyAsync(s.taps, weights, sizeof(double) * n3, cudaMemcpyHostToDevice, ctx->stream));
  const double step = (double)(n - 1) / (double)(M - 1);
  k_rbf_cubic<<<(int)((total + 127) / 128), 128, sizeof(double) * n3, ctx->stream>>>(s.taps, n, M, step, s.a);
  ctx->launches++;
  QMB_CUDA(cudaGetLastError());
  QMB_TRY(peak_of(ctx, s.a, total, s.part, s.peak));
  PeakPartial h;
  QMB_CUDA(cudaMemcpyAsync(&h, s.peak, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  QMB_CUDA(cudaStreamSynchronize(ctx->stream));
  *argmax = h.idx;
  if (maxval) *maxval = h.mx;
  return QMB_OK;
}

// map / nanmax(map) in place (scan.py:721)
extern "C" int qmb_map3d_normalise(qmb_ctx_t ctx, qmb_map3d_t m) {
  QMB_TRY(check_map(ctx, m));
  const long long n = (long long)m->nx * m->ny * m->nz;
  MapScratch s;
  QMB_TRY(map_scratch(ctx, 0, 0, s));
  QMB_TRY(peak_of(ctx, m->v.p, n, s.part, s.peak));
  k_scale_by_peak<<<(int)((n + 255) / 256), 256, 0, ctx->stream>>>(m->v.p, n, s.peak);
  ctx->launches++;
  QMB_CUDA(cudaGetLastError());
  return QMB_OK;
}

// QuakeScan._gaufilt3d (scan.py:1008-1043): new device map
extern "C" int qmb_map3d_gaufilt(qmb_ctx_t ctx, qmb_map3d_t m, double sgm, qmb_map3d_t *out) {
  QMB_TRY(check_map(ctx, m));
  if (!out) QMB_FAIL(QMB_ERR_ARG, "null argument");
  *out = nullptr;
  if (!(sgm > 0)) QMB_FAIL(QMB_ERR_ARG, "bad sigma");
  const Dims d = {m->nx, m->ny, m->nz};
  const long long n = (long long)d.nx * d.ny * d.nz;
  std::vector<double> w, wa;
  int lo[3], hi[3], off[3];
  const int ns[3] = {d.nx, d.ny, d.nz};
  for (int a = 0; a < 3; ++a) {
    gaussian_taps(ns[a], sgm, w, lo[a], hi[a]);
    off[a] = (int)wa.size();
    wa.insert(wa.end(), w.begin(), w.end());
  }
  MapScratch s;
  QMB_TRY(map_scratch(ctx, n, (int)wa.size(), s));
  qmb_map3d *r = nullptr;
  QMB_TRY(qmb_map3d_new(ctx, d.nx, d.ny, d.nz, &r));
  // the taps are consumed by kernels of this stream only; a pageable H2D copy returns after the
  // source has been staged, so `wa` may go out of scope
  cudaError_t e = cudaMemcpyAsync(s.taps, wa.data(), wa.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  int rc = e == cudaSuccess ? QMB_OK : QMB_ERR_CUDA;
  if (rc != QMB_OK) qmb_set_error("qmb_map3d_gaufilt: %s", cudaGetErrorString(e));
  const int blocks = (int)((n + 255) / 256);
  const double *src = m->v.p;
  for (int rep = 0; rep < 2 && rc == QMB_OK; ++rep) {
    // src -> (conv) -> a ; a -> (mirror, / max) -> r
    rc = conv3(ctx, src, s.a, s.b, d, s.taps, lo, hi, off);
    if (rc == QMB_OK) rc = peak_of(ctx, s.a, n, s.part, s.peak);
    if (rc == QMB_OK) {
      k_mirror_scale<<<blocks, 256, 0, ctx->stream>>>(s.a, r->v.p, n, s.peak);
      ctx->launches++;
    }
    src = r->v.p;
  }
  if (rc == QMB_OK && cudaGetLastError() != cudaSuccess) {
    qmb_set_error("qmb_map3d_gaufilt: kernel launch failed");
    rc = QMB_ERR_CUDA;
  }
  if (rc != QMB_OK) {
    qmb_map3d_free(r);
    return rc;
  }
  *out = r;
  return QMB_OK;
}

// box [x0,x1) x [y0,y1) x [z0,z1) of a map to the host, C order (the 5^3 / 7^3 fit windows)
extern "C" int qmb_map3d_window(qmb_ctx_t ctx, qmb_map3d_t m, int32_t x0, int32_t x1, int32_t y0, int32_t y1,
                                int32_t z0, int32_t z1, double *host) {
  QMB_TRY(check_map(ctx, m));
  if (!host) QMB_FAIL(QMB_ERR_ARG, "null argument");
  if (x0 < 0 || y0 < 0 || z0 < 0 || x1 > m->nx || y1 > m->ny || z1 > m->nz || x0 >= x1 || y0 >= y1 || z0 >= z1)
    QMB_FAIL(QMB_ERR_ARG, "window outside the map");
  const long long cnt = (long long)(x1 - x0) * (y1 - y0) * (z1 - z0);
  if (cnt > (1 << 20)) QMB_FAIL(QMB_ERR_ARG, "window too large (use qmb_map3d_fetch)");
  MapScratch s;
  QMB_TRY(map_scratch(ctx, cnt, 0, s));
  const Dims d = {m->nx, m->ny, m->nz};
  k_map_window<<<(int)((cnt + 127) / 128), 128, 0, ctx->stream>>>(m->v.p, d, x0, y0, z0, x1 - x0, y1 - y0, z1 - z0, s.a);
  ctx->launches++;
  QMB_CUDA(cudaGetLastError());
  QMB_CUDA(cudaMemcpyAsync(host, s.a, cnt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  QMB_CUDA(cudaStreamSynchronize(ctx->stream));
  return QMB_OK;
}

// reductions of QuakeScan._covfit3d (scan.py:939-1005).  out[0..2] = expectation (xe, ye, ze) in
// grid units * spacing (relative to the grid origin), out[3..11] = 3x3 covariance matrix (row major).
// win <= 0: whole grid; else nodes within +-(win-1)//2 of the maximum (scan.py:1045-1070).
extern "C" int qmb_map3d_covfit(qmb_ctx_t ctx, qmb_map3d_t m, const double *spacing, double thresh, int32_t win,
                                double *out12) {
  QMB_TRY(check_map(ctx, m));
  if (!spacing || !out12) QMB_FAIL(QMB_ERR_ARG, "null argument");
  const int nx = m->nx, ny = m->ny, nz = m->nz;
  const long long n = (long long)nx * ny * nz;
  MapScratch s;
  QMB_TRY(map_scratch(ctx, 0, 0, s));
  CovParams p = {};
  p.map = m->v.p; p.d = {nx, ny, nz}; p.sx = spacing[0]; p.sy = spacing[1]; p.sz = spacing[2]; p.thresh = thresh;
  if (win > 0) {
    QMB_TRY(peak_of(ctx, m->v.p, n, s.part, s.peak));
    PeakPartial h;
    QMB_CUDA(cudaMemcpyAsync(&h, s.peak, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    QMB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h.idx < 0) QMB_FAIL(QMB_ERR_ARG, "map has no finite value");
    const int iz = (int)(h.idx % nz), iy = (int)((h.idx / nz) % ny), ix = (int)(h.idx / ((long long)nz * ny));
    const int w2 = (win - 1) / 2;
    auto clip = [](int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); };
    p.use_win = 1;
    p.x1 = clip(ix - w2, nx); p.x2 = clip(ix + w2 + 1, nx);
    p.y1 = clip(iy - w2, ny); p.y2 = clip(iy + w2 + 1, ny);
    p.z1 = clip(iz - w2, nz); p.z2 = clip(iz + w2 + 1, nz);
  }
  double h1[8], h2[8];
  k_cov_partial<1><<<RED_BLOCKS, 256, 0, ctx->stream>>>(p, s.cov_part);
  k_cov_final<<<1, 1, 0, ctx->stream>>>(s.cov_part, RED_BLOCKS, s.cov_out);
  ctx->launches += 2;
  QMB_CUDA(cudaGetLastError());
  QMB_CUDA(cudaMemcpyAsync(h1, s.cov_out, 7 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  QMB_CUDA(cudaStreamSynchronize(ctx->stream));
  const double ssw = h1[0];
  p.xe = h1[1] / ssw; p.ye = h1[2] / ssw; p.ze = h1[3] / ssw;
  k_cov_partial<2><<<RED_BLOCKS, 256, 0, ctx->stream>>>(p, s.cov_part);
  k_cov_final<<<1, 1, 0, ctx->stream>>>(s.cov_part, RED_BLOCKS, s.cov_out);
  ctx->launches += 2;
  QMB_CUDA(cudaGetLastError());
  QMB_CUDA(cudaMemcpyAsync(h2, s.cov_out, 7 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  QMB_CUDA(cudaStreamSynchronize(ctx->stream));
  out12[0] = p.xe; out12[1] = p.ye; out12[2] = p.ze;
  const double cxx = h2[0] / ssw, cyy = h2[1] / ssw, czz = h2[2] / ssw, cxy = h2[3] / ssw, cxz = h2[4] / ssw,
               cyz = h2[5] / ssw;
  const double cov[9] = {cxx, cxy, cxz, cxy, cyy, cyz, cxz, cyz, czz};
  for (int k = 0; k < 9; ++k) out12[3 + k] = cov[k];
  return QMB_OK;
}

// ---------------------------------------------------------------------------------------------
// host-buffer forms (upload -> device form -> fetch)
// ---------------------------------------------------------------------------------------------
extern "C" int qmb_gaufilt3d(qmb_ctx_t ctx, const double *map3d, int32_t nx, int32_t ny, int32_t nz, double sgm,
                             double *smoothed) {
  if (!ctx || !map3d || !smoothed) QMB_FAIL(QMB_ERR_ARG, "null argument");
  if (nx <= 0 || ny <= 0 || nz <= 0 || !(sgm > 0)) QMB_FAIL(QMB_ERR_ARG, "bad map shape or sigma");
  qmb_map3d_t m = nullptr, r = nullptr;
  QMB_TRY(qmb_map3d_upload(ctx, map3d, nx, ny, nz, &m));
  int rc = qmb_map3d_gaufilt(ctx, m, sgm, &r);
  if (rc == QMB_OK) rc = qmb_map3d_fetch(r, smoothed);
  qmb_map3d_free(m);
  qmb_map3d_free(r);
  return rc;
}

extern "C" int qmb_map3d_peak(qmb_ctx_t ctx, const double *map3d, int64_t n, int64_t *argmax, double *maxval,
                              double *mean) {
  if (!ctx || !map3d || n <= 0 || n > 0x7fffffffLL) QMB_FAIL(QMB_ERR_ARG, "bad argument");
  qmb_map3d_t m = nullptr;
  QMB_TRY(qmb_map3d_upload(ctx, map3d, 1, 1, (int32_t)n, &m));
  const int rc = qmb_map3d_peak_device(ctx, m, argmax, maxval, mean);
  qmb_map3d_free(m);
  return rc;
}

extern "C" int qmb_covfit3d(qmb_ctx_t ctx, const double *map3d, int32_t nx, int32_t ny, int32_t nz,
                            const double *spacing, double thresh, int32_t win, double *out12) {
  if (!ctx || !map3d || !spacing || !out12) QMB_FAIL(QMB_ERR_ARG, "null argument");
  if (nx <= 0 || ny <= 0 || nz <= 0) QMB_FAIL(QMB_ERR_ARG, "bad map shape");
  qmb_map3d_t m = nullptr;
  QMB_TRY(qmb_map3d_upload(ctx, map3d, nx, ny, nz, &m));
  const int rc = qmb_map3d_covfit(ctx, m, spacing, thresh, win, out12);
  qmb_map3d_free(m);
  return rc;
}

// ---------------------------------------------------------------------------------------------
// micro-benchmark: the ceiling of the pipe that bounds k_scan.  Each warp streams conflict-free
// 128-byte wavefronts from shared memory into independent fp32 accumulators -- the scan kernel's
// inner loop with everything else removed (no staging, no offsets, no epilogue).  Measured on
// B200 (scripts/smem_pipe_sweep.py): 99.5 % of the nominal 32 gather-adds/clk/SM with 12, 16, 24
// or 32 warps per SM for LDS.32/64/128 alike; 91 % with 13 warps (one scheduler carries 4).
// ---------------------------------------------------------------------------------------------
namespace {
// VEC floats per lane per load (1: LDS.32, 2: LDS.64, 4: LDS.128), any warp count
template <int VEC, int NV>
__global__ void k_lds_peak_vec(float *out, int iters) {
  extern __shared__ float sm[];
  const int n_words = 32 * 1024;
  for (int i = threadIdx.x; i < n_words; i += blockDim.x) sm[i] = (float)(i & 7) * 0.125f;
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) acc[j] = 0.f;
  uint32_t off = (warp * 1024u) & (n_words / 2 - 1);
#pragma unroll 2
  for (int it = 0; it < iters; ++it) {
    const uint32_t a = base + ((off + lane * VEC) << 2);
    float v[NV];
#pragma unroll
    for (int j = 0; j < NV / VEC; ++j) {
      const uint32_t aj = a + 128u * VEC * j;
      if constexpr (VEC == 1) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[j]) : "r"(aj));
      if constexpr (VEC == 2)
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[2 * j]), "=f"(v[2 * j + 1]) : "r"(aj));
      if constexpr (VEC == 4)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v[4 * j]), "=f"(v[4 * j + 1]), "=f"(v[4 * j + 2]), "=f"(v[4 * j + 3])
                     : "r"(aj));
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] += v[j];
    off += 1028u;  // multiple of 4: keeps vector loads aligned
    if (off >= (uint32_t)(n_words - 32 * NV - 128)) off -= (uint32_t)(n_words - 32 * NV - 128) & ~3u;
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) s += acc[j];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int VEC, int NV>
int run_vec(qmb_ctx *ctx, int warps, int iters, float *d_out, float *ms) {
  const size_t smem = 128 * 1024;
  QMB_CUDA(cudaFuncSetAttribute(k_lds_peak_vec<VEC, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_lds_peak_vec<VEC, NV><<<ctx->sm_count, warps * 32, smem, ctx->stream>>>(d_out, iters / 8 + 1);
  QMB_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  k_lds_peak_vec<VEC, NV><<<ctx->sm_count, warps * 32, smem, ctx->stream>>>(d_out, iters);
  QMB_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  ctx->launches += 2;
  QMB_CUDA(cudaStreamSynchronize(ctx->stream));
  QMB_CUDA(cudaEventElapsedTime(ms, ctx->ev[0], ctx->ev[1]));
  return QMB_OK;
}
}  // namespace

// vec in {1,2,4}, warps in [1,32]; 32 (warps <= 16) or 16 (more warps) values per lane per iteration
extern "C" int qmb_microbench_smem_pipe_ex(qmb_ctx_t ctx, int32_t iters, int32_t vec, int32_t warps,
                                           double *gather_adds_per_s) {
  if (!ctx || !gather_adds_per_s || iters <= 0 || warps < 1 || warps > 32) QMB_FAIL(QMB_ERR_ARG, "bad argument");
  QMB_CUDA(cudaSetDevice(ctx->device));
  float *d_out = nullptr;
  QMB_CUDA(cudaMalloc((void **)&d_out, (size_t)ctx->sm_count * 1024 * sizeof(float)));
  float ms = 0.f;
  int rc = QMB_ERR_ARG;
  const bool big = warps <= 16;
  if (vec == 1) rc = big ? run_vec<1, 32>(ctx, warps, iters, d_out, &ms) : run_vec<1, 16>(ctx, warps, iters, d_out, &ms);
  else if (vec == 2) rc = big ? run_vec<2, 32>(ctx, warps, iters, d_out, &ms) : run_vec<2, 16>(ctx, warps, iters, d_out, &ms);
  else if (vec == 4) rc = big ? run_vec<4, 32>(ctx, warps, iters, d_out, &ms) : run_vec<4, 16>(ctx, warps, iters, d_out, &ms);
  else qmb_set_error("vec must be 1, 2 or 4");
  cudaFree(d_out);
  if (rc != QMB_OK) return rc;
  *gather_adds_per_s = (double)ctx->sm_count * warps * 32.0 * (big ? 32.0 : 16.0) * (double)iters / ((double)ms * 1e-3);
  return QMB_OK;
}

// The ceiling k_scan is judged against: LDS.32 + FADD streams on 12 warps per SM (its consumer
// warp count), 32 independent gathers per lane per iteration.
extern "C" int qmb_microbench_smem_pipe(qmb_ctx_t ctx, int32_t iters, double *gather_adds_per_s) {
  return qmb_microbench_smem_pipe_ex(ctx, iters, 1, 12, gather_adds_per_s);
}
