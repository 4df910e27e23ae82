// qmb_envelope.cu -- Hilbert envelopes for the "env" / "env_squared" signal transforms of
// STALTAOnset._onset (quakemigrate/signal/onsets/stalta.py:520-523):
//     np.abs(scipy.signal.hilbert(tr.data))          and its square.
// scipy.signal.hilbert is the FFT construction of the analytic signal: X = fft(x, N = len(x)),
// keep bin 0 (and bin N/2 for even N), double bins 1 .. ceil(N/2)-1, zero the negative
// frequencies, inverse transform -- a DFT of exactly the trace length (409 260 = 2^2*3*5*19*359 at
// cfg3), not of a padded one.
//
// Own transforms, no FFT library:
//   * power-of-two complex FFT (float64): Stockham auto-sort passes of radix 16 (8 / 4 / 2 for the
//     last one), one kernel per pass, ping-pong between two buffers.  Thread i of a pass with
//     p = product of the earlier radices:  k = i mod p,  u[r] = x[i + r*T] * e^(-+2 pi i r k / (p R)),
//     v = DFT_R(u) in registers (radix-2 butterflies, bit-reversed order), y[(i - k) R + k + s p] = v[s].
//     Reads are coalesced; a thread's R stores are p elements apart, coalesced across the warp for
//     p >= 32; the first pass (p = 1, 16 consecutive results per thread) goes through shared memory.
//   * arbitrary N by Bluestein's chirp-z identity  n k = (n^2 + k^2 - (k - n)^2) / 2:
//         DFT_N(a)[k] = conj(w[k]) * sum_n a[n] conj(w[n]) w[k - n],   w[m] = e^(i pi m^2 / N),
//     i.e. a circular convolution of length M = 2^ceil(log2(2N-1)) done with the FFT above; the
//     transform of the chirp filter is computed once per trace length and cached in the context
//     (m^2 mod 2N in 64-bit integers, sincospi: the chirp is accurate for any N).
//   * the inverse DFT is conj(DFT(conj(.))) / N, and only its modulus is needed, so the envelope is
//     |second convolution| / (M N): four FFT_M per trace plus three pointwise kernels, the chirp
//     filter multiply fused into the last pass of each forward FFT.
#include <stdlib.h>

#include "qmb_internal.cuh"

namespace {

typedef double2 cplx;

__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }

// cos / sin of 2 pi t / 16, t = 0 .. 7
__device__ __constant__ double TW16_C[8] = {1.0, 0.92387953251128674, 0.70710678118654752, 0.38268343236508977,
                                            0.0, -0.38268343236508977, -0.70710678118654752, -0.92387953251128674};
__device__ __constant__ double TW16_S[8] = {0.0, 0.38268343236508977, 0.70710678118654752, 0.92387953251128674,
                                            1.0, 0.92387953251128674, 0.70710678118654752, 0.38268343236508977};

// One Stockham pass of radix R = 2^LR over `batch` transforms of length M (blockIdx.y = transform).
// sign = -1: forward (e^-), +1: inverse (unnormalised).  `mul` (nullable, length M): the result is
// multiplied element-wise by it on the way out (the chirp filter after the last forward pass).
constexpr int FFT_TPB = 128;
template <int LR>
__global__ void __launch_bounds__(FFT_TPB) k_fft_pass(const cplx *__restrict__ x, cplx *__restrict__ y, int log2M, int p,
                                                  double sign, const cplx *__restrict__ mul) {
  constexpr int R = 1 << LR;
  const long long M = 1LL << log2M, T = M >> LR;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  const cplx *xb = x + (size_t)blockIdx.y * M;
  cplx *yb = y + (size_t)blockIdx.y * M;
  const long long k = i & (long long)(p - 1), j = ((i - k) << LR) + k;
  cplx u[R];
#pragma unroll
  for (int r = 0; r < R; ++r) u[r] = xb[i + r * T];
  if (p > 1) {
    // e^(sign 2 pi i r k / (p R)), r = 1 .. R-1: k / (p R) is exact (power-of-two denominator)
    cplx w[R];
    double s, c;
    sincospi(2.0 * (double)k / ((double)p * (double)R), &s, &c);
    w[1] = make_double2(c, sign * s);
#pragma unroll
    for (int r = 2; r < R; ++r) w[r] = (r & 1) ? cmul(w[r - 1], w[1]) : cmul(w[r / 2], w[r / 2]);
#pragma unroll
    for (int r = 1; r < R; ++r) u[r] = cmul(u[r], w[r]);
  }
  // DFT_R in registers: decimation in frequency, results in bit-reversed positions
#pragma unroll
  for (int span = R / 2; span >= 1; span >>= 1) {
#pragma unroll
    for (int base = 0; base < R; base += 2 * span) {
#pragma unroll
      for (int t = 0; t < span; ++t) {
        const cplx a = u[base + t], b = u[base + t + span];
        u[base + t] = make_double2(a.x + b.x, a.y + b.y);
        const cplx d = make_double2(a.x - b.x, a.y - b.y);
        const int q = t * (8 / span);  // angle 2 pi t / (2 span) = 2 pi q / 16
        u[base + t + span] = q == 0 ? d : cmul(d, make_double2(TW16_C[q], sign * TW16_S[q]));
      }
    }
  }
  if constexpr (LR == 4) {
    // First pass of a transform (p = 1): a thread's 16 results are 16 consecutive elements, so the
    // warp's stores would stride 256 bytes.  Transpose through shared memory (rows padded to 17
    // elements: conflict-free 16-byte accesses) and write the block's 2048 elements coalesced.
    if (p == 1 && (T % FFT_TPB) == 0) {
      extern __shared__ double2 stage[];
#pragma unroll
      for (int pos = 0; pos < R; ++pos) {
        int s = 0;
#pragma unroll
        for (int b = 0; b < LR; ++b) s |= ((pos >> b) & 1) << (LR - 1 - b);
        stage[threadIdx.x * (R + 1) + s] = u[pos];
      }
      __syncthreads();
      const long long base = (long long)blockIdx.x * FFT_TPB * R;
#pragma unroll
      for (int m = 0; m < R; ++m) {
        const int e = threadIdx.x + FFT_TPB * m;
        const cplx v = stage[(e >> LR) * (R + 1) + (e & (R - 1))];
        const long long o = base + e;
        yb[o] = mul ? cmul(v, mul[o]) : v;
      }
      return;
    }
  }
#pragma unroll
  for (int pos = 0; pos < R; ++pos) {
    int s = 0;
#pragma unroll
    for (int b = 0; b < LR; ++b) s |= ((pos >> b) & 1) << (LR - 1 - b);
    const long long o = j + (long long)s * p;
    yb[o] = mul ? cmul(u[pos], mul[o]) : u[pos];
  }
}

// w[n] = e^(i pi n^2 / N), n < N  (n^2 mod 2N exactly, then sincospi)
__global__ void __launch_bounds__(256) k_chirp(cplx *w, long long N) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const unsigned long long r = ((unsigned long long)n * (unsigned long long)n) % (2ull * (unsigned long long)N);
  double s, c;
  sincospi((double)r / (double)N, &s, &c);
  w[n] = make_double2(c, s);
}

// circular chirp filter of length M: b[m] = b[M - m] = w[m], zero in between
__global__ void __launch_bounds__(256) k_chirp_filter(const cplx *w, long long N, long long M, cplx *b) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  cplx v = make_double2(0.0, 0.0);
  if (m < N) v = w[m];
  else if (M - m < N) v = w[M - m];
  b[m] = v;
}

// c[n] = x[n] conj(w[n]), zero-padded to M
__global__ void __launch_bounds__(256) k_chirp_in(const double *x, long long x_stride, const cplx *w, long long N,
                                                  long long M, cplx *c) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= M) return;
  cplx v = make_double2(0.0, 0.0);
  if (n < N) {
    const double a = x[(size_t)blockIdx.y * x_stride + n];
    const cplx wn = w[n];
    v = make_double2(a * wn.x, -a * wn.y);
  }
  c[(size_t)blockIdx.y * M + n] = v;
}

// conv = M * (first convolution).  A[k] = conj(w[k]) conv[k] / M is the spectrum; weight it with the
// analytic-signal window h and set up the inverse transform as a forward one of the conjugate:
// c2[k] = conj(h[k] A[k]) conj(w[k]), zero-padded to M.  In place.
__global__ void __launch_bounds__(256) k_analytic_mid(cplx *z, const cplx *w, long long N, long long M, double inv_m) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= M) return;
  cplx *zz = z + (size_t)blockIdx.y * M;
  double h = 0.0;
  if (k < N) {
    if (k == 0) h = 1.0;
    else if ((N & 1) == 0 && k == N / 2) h = 1.0;
    else if (k < (N + 1) / 2) h = 2.0;
  }
  cplx v = make_double2(0.0, 0.0);
  if (h != 0.0) {
    const cplx wk = w[k];
    cplx a = cmul(cconj(wk), zz[k]);  // M A[k]
    a.x *= h * inv_m;
    a.y *= h * inv_m;
    v = cmul(cconj(a), cconj(wk));
  }
  zz[k] = v;
}

// |analytic signal|[n] = |second convolution|[n] / (M N)
__global__ void __launch_bounds__(256) k_envelope(const cplx *z, long long N, long long M, double *out,
                                                  long long out_stride, double scale) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const cplx v = z[(size_t)blockIdx.y * M + n];
  out[(size_t)blockIdx.y * out_stride + n] = hypot(v.x * scale, v.y * scale);
}

// batch transforms of length 2^log2M: a -> result pointer (a or b); `mul` applied after the last pass
int fft_pow2(qmb_ctx *ctx, cplx *a, cplx *b, int log2M, int batch, double sign, const cplx *mul, cplx **result) {
  cplx *src = a, *dst = b;
  int p = 1, rem = log2M;
  if (rem == 0) {  // length 1: the transform is the identity
    *result = a;
    return QMB_OK;
  }
  while (rem > 0) {
    const int LR = rem >= 4 ? 4 : rem;
    const long long T = (1LL << log2M) >> LR;
    const dim3 grid((unsigned)((T + FFT_TPB - 1) / FFT_TPB), (unsigned)batch);
    const cplx *m = (rem == LR) ? mul : nullptr;
    switch (LR) {
      case 4: k_fft_pass<4><<<grid, FFT_TPB, sizeof(cplx) * FFT_TPB * 17, ctx->stream>>>(src, dst, log2M, p, sign, m); break;
      case 3: k_fft_pass<3><<<grid, FFT_TPB, 0, ctx->stream>>>(src, dst, log2M, p, sign, m); break;
      case 2: k_fft_pass<2><<<grid, FFT_TPB, 0, ctx->stream>>>(src, dst, log2M, p, sign, m); break;
      default: k_fft_pass<1><<<grid, FFT_TPB, 0, ctx->stream>>>(src, dst, log2M, p, sign, m); break;
    }
    ctx->launches++;
    QMB_CUDA(cudaGetLastError());
    cplx *t = src; src = dst; dst = t;
    p <<= LR;
    rem -= LR;
  }
  *result = src;
  return QMB_OK;
}

}  // namespace

void qmb_fft_release(qmb_ctx *ctx) {
  ctx->fft_n = 0;
  ctx->fft_buf.release();
  ctx->fft_chirp.release();
}

// d_out[c][i] = |analytic signal of d_x[c]|[i]; channels are rows of stride x_stride / out_stride
int qmb_envelope_device(qmb_ctx *ctx, const double *d_x, int64_t x_stride, int32_t n_channels, int64_t n,
                        double *d_out, int64_t out_stride) {
  if (n_channels <= 0 || n <= 0) return QMB_OK;
  if (n > (1LL << 28)) QMB_FAIL(QMB_ERR_UNSUPPORTED, "trace too long for the envelope transform (%lld samples)", (long long)n);
  int log2M = 0;
  while ((1LL << log2M) < 2 * n - 1) ++log2M;
  const long long M = 1LL << log2M;
  // channels per batch: two complex work buffers of at most 1 GiB each
  long long batch = (1ll << 30) / (16 * M);
  if (batch < 1) batch = 1;
  if (batch > n_channels) batch = n_channels;
  if (batch > 65535) batch = 65535;
  QMB_TRY(ctx->fft_buf.ensure((size_t)batch * M * 4));
  cplx *bufA = (cplx *)ctx->fft_buf.p, *bufB = bufA + (size_t)batch * M;
  const unsigned gm = (unsigned)((M + 255) / 256);
  // chirp w (N) and the transformed chirp filter (M), cached per trace length
  if (ctx->fft_n != n) {
    ctx->fft_n = 0;
    QMB_TRY(ctx->fft_chirp.ensure((size_t)(n + 2 * M) * 2));
    cplx *w = (cplx *)ctx->fft_chirp.p, *f0 = w + n, *f1 = f0 + M;
    k_chirp<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(w, n);
    k_chirp_filter<<<gm, 256, 0, ctx->stream>>>(w, n, M, f0);
    ctx->launches += 2;
    QMB_CUDA(cudaGetLastError());
    cplx *res = nullptr;
    QMB_TRY(fft_pow2(ctx, f0, f1, log2M, 1, -1.0, nullptr, &res));
    if (res != f0) QMB_CUDA(cudaMemcpyAsync(f0, res, sizeof(cplx) * M, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->fft_n = n;
  }
  const cplx *w = (const cplx *)ctx->fft_chirp.p, *filt = w + n;
  for (int c0 = 0; c0 < n_channels; c0 += (int)batch) {
    const int b = (int)((n_channels - c0) < batch ? (n_channels - c0) : batch);
    const dim3 grid_m(gm, (unsigned)b), grid_n((unsigned)((n + 255) / 256), (unsigned)b);
    k_chirp_in<<<grid_m, 256, 0, ctx->stream>>>(d_x + (size_t)c0 * x_stride, x_stride, w, n, M, bufA);
    ctx->launches++;
    QMB_CUDA(cudaGetLastError());
    cplx *f = nullptr, *g = nullptr;
    QMB_TRY(fft_pow2(ctx, bufA, bufB, log2M, b, -1.0, filt, &f));
    QMB_TRY(fft_pow2(ctx, f, f == bufA ? bufB : bufA, log2M, b, +1.0, nullptr, &g));
    k_analytic_mid<<<grid_m, 256, 0, ctx->stream>>>(g, w, n, M, 1.0 / (double)M);
    ctx->launches++;
    QMB_CUDA(cudaGetLastError());
    QMB_TRY(fft_pow2(ctx, g, g == bufA ? bufB : bufA, log2M, b, -1.0, filt, &f));
    QMB_TRY(fft_pow2(ctx, f, f == bufA ? bufB : bufA, log2M, b, +1.0, nullptr, &g));
    k_envelope<<<grid_n, 256, 0, ctx->stream>>>(g, n, M, d_out + (size_t)c0 * out_stride, out_stride,
                                                1.0 / ((double)M * (double)n));
    ctx->launches++;
    QMB_CUDA(cudaGetLastError());
  }
  return QMB_OK;
}

// host in / host out: np.abs(scipy.signal.hilbert(x, axis=-1)) for a (n_channels, n) block
extern "C" int qmb_envelope(qmb_ctx_t ctx, const double *x, int32_t n_channels, int64_t n, double *out) {
  if (!ctx || !x || !out) QMB_FAIL(QMB_ERR_ARG, "null argument");
  if (n_channels <= 0 || n <= 0) return QMB_OK;
  QMB_CUDA(cudaSetDevice(ctx->device));
  const size_t count = (size_t)n_channels * n;
  double *d_in = nullptr, *d_res = nullptr;
  QMB_CUDA(cudaMalloc((void **)&d_in, count * sizeof(double)));
  cudaError_t e = cudaMalloc((void **)&d_res, count * sizeof(double));
  if (e != cudaSuccess) {
    cudaFree(d_in);
    QMB_FAIL(QMB_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
  }
  int rc = [&]() -> int {
    QMB_CUDA(cudaMemcpyAsync(d_in, x, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    QMB_TRY(qmb_envelope_device(ctx, d_in, n, n_channels, n, d_res, n));
    QMB_CUDA(cudaMemcpyAsync(out, d_res, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    QMB_CUDA(cudaStreamSynchronize(ctx->stream));
    return QMB_OK;
  }();
  cudaStreamSynchronize(ctx->stream);
  cudaFree(d_in);
  cudaFree(d_res);
  return rc;
}
