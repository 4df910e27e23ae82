// qmb_envelope.cu -- Hilbert envelopes for the "env" / "env_squared" signal transforms of
// STALTAOnset._onset (quakemigrate/signal/onsets/stalta.py:520-523):
//     np.abs(scipy.signal.hilbert(tr.data))          and its square.
// scipy.signal.hilbert is the FFT construction of the analytic signal: X = fft(x, N = len(x)),
// keep bin 0 (and bin N/2 for even N), double bins 1 .. ceil(N/2)-1, zero the negative
// frequencies, inverse transform.  The same three steps here: a batched real-to-complex cuFFT
// (library FFT, opened with dlopen so that nothing else depends on it), our weighting kernel, a
// batched complex inverse cuFFT, and our |.| kernel with the 1/N of the inverse folded in.
#include <cufft.h>
#include <dlfcn.h>
#include <stdlib.h>

#include <mutex>

#include "qmb_internal.cuh"

namespace {

struct CufftApi {
  void *handle = nullptr;
  cufftResult (*PlanMany)(cufftHandle *, int, int *, int *, int, int, int *, int, int, cufftType, int) = nullptr;
  cufftResult (*ExecD2Z)(cufftHandle, cufftDoubleReal *, cufftDoubleComplex *) = nullptr;
  cufftResult (*ExecZ2Z)(cufftHandle, cufftDoubleComplex *, cufftDoubleComplex *, int) = nullptr;
  cufftResult (*SetStream)(cufftHandle, cudaStream_t) = nullptr;
  cufftResult (*Destroy)(cufftHandle) = nullptr;
};
CufftApi g_fft;
std::mutex g_fft_mu;

int cufft_load() {
  std::lock_guard<std::mutex> lock(g_fft_mu);
  if (g_fft.handle) return QMB_OK;
  const char *env = getenv("QMB_CUFFT_LIB");
  const char *names[] = {env, "libcufft.so.11", "libcufft.so", "/usr/local/cuda/lib64/libcufft.so.11",
                         "/usr/local/cuda/lib64/libcufft.so"};
  void *h = nullptr;
  for (const char *nm : names) {
    if (!nm) continue;
    h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
    if (h) break;
  }
  if (!h) QMB_FAIL(QMB_ERR_UNSUPPORTED, "cannot dlopen libcufft (%s); set QMB_CUFFT_LIB", dlerror());
#define QMB_SYM(field, name)                                                   \
  *(void **)(&g_fft.field) = dlsym(h, name);                                   \
  if (!g_fft.field) QMB_FAIL(QMB_ERR_UNSUPPORTED, "libcufft lacks symbol %s", name)
  QMB_SYM(PlanMany, "cufftPlanMany");
  QMB_SYM(ExecD2Z, "cufftExecD2Z");
  QMB_SYM(ExecZ2Z, "cufftExecZ2Z");
  QMB_SYM(SetStream, "cufftSetStream");
  QMB_SYM(Destroy, "cufftDestroy");
#undef QMB_SYM
  g_fft.handle = h;
  return QMB_OK;
}

#define QMB_FFT(call)                                                                  \
  do {                                                                                 \
    cufftResult r__ = (call);                                                          \
    if (r__ != CUFFT_SUCCESS) {                                                        \
      qmb_set_error("%s:%d %s -> cufftResult %d", __FILE__, __LINE__, #call, (int)r__); \
      return QMB_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)

// bins 0 .. n/2 come from the real-to-complex transform; weight them and clear the rest
__global__ void __launch_bounds__(256) k_analytic_weights(cufftDoubleComplex *z, long long n) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  cufftDoubleComplex *zz = z + (size_t)blockIdx.y * n;
  double h;
  if (k == 0) h = 1.0;
  else if ((n & 1) == 0 && k == n / 2) h = 1.0;
  else if (k < (n + 1) / 2) h = 2.0;
  else h = 0.0;
  cufftDoubleComplex v = {0.0, 0.0};
  if (h != 0.0) {
    v = zz[k];
    v.x *= h;
    v.y *= h;
  }
  zz[k] = v;
}

__global__ void __launch_bounds__(256) k_envelope(const cufftDoubleComplex *z, long long n, double *out,
                                                  long long out_stride, double inv_n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const cufftDoubleComplex v = z[(size_t)blockIdx.y * n + i];
  out[(size_t)blockIdx.y * out_stride + i] = hypot(v.x * inv_n, v.y * inv_n);
}

}  // namespace

void qmb_fft_release(qmb_ctx *ctx) {
  if (ctx->fft_n && g_fft.handle) {
    g_fft.Destroy(ctx->fft_fwd);
    g_fft.Destroy(ctx->fft_inv);
  }
  ctx->fft_n = 0;
  ctx->fft_buf.release();
}

// d_out[c][i] = |analytic signal of d_x[c]|[i]; channels are rows of stride x_stride / out_stride
int qmb_envelope_device(qmb_ctx *ctx, const double *d_x, int64_t x_stride, int32_t n_channels, int64_t n,
                        double *d_out, int64_t out_stride) {
  if (n_channels <= 0 || n <= 0) return QMB_OK;
  if (n > 0x7fffffffLL || x_stride > 0x7fffffffLL) QMB_FAIL(QMB_ERR_UNSUPPORTED, "trace too long for a single FFT");
  QMB_TRY(cufft_load());
  // channels per batch: complex scratch of at most 1 GiB
  long long batch = (1ll << 30) / (16 * n);
  if (batch < 1) batch = 1;
  if (batch > n_channels) batch = n_channels;
  QMB_TRY(ctx->fft_buf.ensure((size_t)batch * n * 2));
  cufftDoubleComplex *z = (cufftDoubleComplex *)ctx->fft_buf.p;
  for (int c0 = 0; c0 < n_channels; c0 += (int)batch) {
    const int b = (int)((n_channels - c0) < batch ? (n_channels - c0) : batch);
    if (ctx->fft_n != n || ctx->fft_batch != b || ctx->fft_idist != x_stride) {
      if (ctx->fft_n) {
        g_fft.Destroy(ctx->fft_fwd);
        g_fft.Destroy(ctx->fft_inv);
        ctx->fft_n = 0;
      }
      int dims[1] = {(int)n}, embed[1] = {(int)n};
      QMB_FFT(g_fft.PlanMany(&ctx->fft_fwd, 1, dims, embed, 1, (int)x_stride, embed, 1, (int)n, CUFFT_D2Z, b));
      cufftResult r = g_fft.PlanMany(&ctx->fft_inv, 1, dims, embed, 1, (int)n, embed, 1, (int)n, CUFFT_Z2Z, b);
      if (r != CUFFT_SUCCESS) {
        g_fft.Destroy(ctx->fft_fwd);
        QMB_FAIL(QMB_ERR_CUDA, "cufftPlanMany(Z2Z, n=%lld, batch=%d) -> cufftResult %d", (long long)n, b, (int)r);
      }
      ctx->fft_n = n; ctx->fft_batch = b; ctx->fft_idist = x_stride;
      QMB_FFT(g_fft.SetStream(ctx->fft_fwd, ctx->stream));
      QMB_FFT(g_fft.SetStream(ctx->fft_inv, ctx->stream));
    }
    QMB_FFT(g_fft.ExecD2Z(ctx->fft_fwd, (cufftDoubleReal *)(d_x + (size_t)c0 * x_stride), z));
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)b);
    k_analytic_weights<<<grid, 256, 0, ctx->stream>>>(z, n);
    QMB_CUDA(cudaGetLastError());
    QMB_FFT(g_fft.ExecZ2Z(ctx->fft_inv, z, z, CUFFT_INVERSE));
    k_envelope<<<grid, 256, 0, ctx->stream>>>(z, n, d_out + (size_t)c0 * out_stride, out_stride, 1.0 / (double)n);
    QMB_CUDA(cudaGetLastError());
    ctx->launches += 2;
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
