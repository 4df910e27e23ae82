// qmb_legacy.cu -- the five symbols of the reference C extension (quakemigrate/core/src/qmlib.h:28-44)
// with byte-identical signatures, implemented on the GPU.  An unmodified quakemigrate.core can
// load libqmb200.so in place of core/src/qmlib<EXT_SUFFIX> (core/libnames.py:35-38).
//
// These honour the reference's ownership contract (the caller allocates every buffer, map4d is
// accumulated into, STA/LTA outputs arrive pre-filled), which forces the full (n_nodes, n_samples)
// float64 volume over PCIe: they are the plumbing tier.  The fast path is the T2 API.
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "qmb_internal.cuh"

namespace {

std::mutex g_default_mu;
qmb_ctx *g_default = nullptr;

// The five reference symbols return void (qmlib.h:28-44), so a failure cannot be returned.  It is
// recorded (qmb_legacy_status / qmb_last_error -- quakemigrate_b200.core checks them after every call
// and raises), reported on stderr, and the call returns with the caller's output buffers untouched.
// QMB_LEGACY_ABORT=1 restores a hard abort() for an unmodified quakemigrate.core, which has no way
// of noticing the status and would otherwise go on with a zero map.
thread_local int g_legacy_status = QMB_OK;

void legacy_fail(const char *what, int status) {
  g_legacy_status = status ? status : QMB_ERR_CUDA;
  fprintf(stderr, "libqmb200: %s failed: %s\n(no CPU fallback exists in this library; output left untouched)\n", what,
          qmb_last_error());
  const char *ab = getenv("QMB_LEGACY_ABORT");
  if (ab && atoi(ab) != 0) abort();
}

struct DevMem {  // device allocation released on scope exit
  void *p = nullptr;
  ~DevMem() {
    if (p) cudaFree(p);
  }
  int alloc(size_t bytes) {
    QMB_CUDA(cudaMalloc(&p, bytes));
    return QMB_OK;
  }
  template <typename T>
  T *as() const { return (T *)p; }
};

// per-sample (max, first arg-max, sum) over a slab of nodes; one thread per sample, nodes split
// across blockIdx.y groups.  Node order inside a group is ascending, ">" is strict, so ties keep
// the lowest index exactly like migratelib.c:96-105.
__global__ void __launch_bounds__(128) k_colreduce(const double *map, long long n_nodes, int n_samples,
                                                   long long node_base, int groups, double *g_max,
                                                   long long *g_idx, double *g_sum) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_samples) return;
  const int g = blockIdx.y;
  const long long per = (n_nodes + groups - 1) / groups;
  const long long n0 = g * per, n1 = min(n_nodes, n0 + per);
  double mx = -1.0, sm = 0.0;
  long long ix = -1;
  for (long long n = n0; n < n1; ++n) {
    const double c = map[(size_t)n * n_samples + t];
    sm += c;
    if (ix < 0 || c > mx) { mx = c; ix = node_base + n; }
  }
  g_max[(size_t)g * n_samples + t] = mx;
  g_idx[(size_t)g * n_samples + t] = ix;
  g_sum[(size_t)g * n_samples + t] = sm;
}

// fold group partials into the running per-sample state (slabs arrive in ascending node order)
__global__ void __launch_bounds__(128) k_colfold(const double *g_max, const long long *g_idx, const double *g_sum,
                                                 int groups, int n_samples, double *r_max, long long *r_idx,
                                                 double *r_sum, int first) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_samples) return;
  double mx = first ? -1.0 : r_max[t], sm = first ? 0.0 : r_sum[t];
  long long ix = first ? -1 : r_idx[t];
  for (int g = 0; g < groups; ++g) {
    const long long gi = g_idx[(size_t)g * n_samples + t];
    if (gi < 0) continue;
    const double c = g_max[(size_t)g * n_samples + t];
    sm += g_sum[(size_t)g * n_samples + t];
    if (ix < 0 || c > mx) { mx = c; ix = gi; }
  }
  r_max[t] = mx; r_idx[t] = ix; r_sum[t] = sm;
}

__global__ void k_colfinish(const double *r_max, const double *r_sum, double n_nodes, int n_samples,
                            double *norm) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_samples) norm[t] = r_max[t] * n_nodes / r_sum[t];
}

}  // namespace

qmb_ctx *qmb_default_ctx() {
  std::lock_guard<std::mutex> lock(g_default_mu);
  if (!g_default) {
    const char *d = getenv("QMB_DEVICE");
    qmb_ctx_t c = nullptr;
    if (qmb_ctx_create(d ? atoi(d) : 0, &c) != QMB_OK) return nullptr;  // message in qmb_last_error
    g_default = c;
  }
  return g_default;
}

extern "C" int qmb_legacy_status(void) {
  const int s = g_legacy_status;
  g_legacy_status = QMB_OK;
  return s;
}

namespace {

int migrate_impl(double *onsets, int32_t *lookup_tables, double *map4d, int32_t fsmp, int32_t lsmp,
                 int32_t n_samples, int32_t n_stations, int32_t available, int64_t n_nodes) {
  qmb_ctx *ctx = qmb_default_ctx();
  if (!ctx) return QMB_ERR_CUDA;
  QMB_CUDA(cudaSetDevice(ctx->device));
  const int64_t P = (int64_t)fsmp + lsmp + n_samples;

  // onsets: upload once, skip the clip/log
  struct OnsetsGuard {
    qmb_onsets *O = new qmb_onsets();
    ~OnsetsGuard() { qmb_onsets_free(O); }
  } og;
  qmb_onsets *O = og.O;
  O->ctx = ctx; O->device = ctx->device; O->S = n_stations; O->P = P;
  {
    DevMem lin;
    const size_t on_bytes = (size_t)n_stations * P * sizeof(double);
    QMB_TRY(lin.alloc(on_bytes));
    QMB_CUDA(cudaMemcpyAsync(lin.p, onsets, on_bytes, cudaMemcpyHostToDevice, ctx->stream));
    QMB_TRY(qmb_onsets_prepare(ctx, O, lin.as<double>(), 1));
    QMB_CUDA(cudaStreamSynchronize(ctx->stream));
  }

  // node slabs sized so that the float64 slab of the volume stays under ~2 GiB on the device
  int64_t slab = (int64_t)((2ull << 30) / ((size_t)n_samples * sizeof(double)));
  slab = (slab / 64) * 64;
  if (slab < 64) slab = 64;
  DevMem map;
  const int64_t slab_alloc = slab < n_nodes ? slab : n_nodes;
  QMB_TRY(map.alloc((size_t)slab_alloc * n_samples * sizeof(double)));

  for (int64_t n0 = 0; n0 < n_nodes; n0 += slab) {
    const int64_t nn = (n_nodes - n0 < slab) ? n_nodes - n0 : slab;
    // pseudo-grid (ceil(nn/16), 4, 4): 64 consecutive flat nodes form one tile
    struct LutGuard {
      qmb_lut *L = new qmb_lut();
      ~LutGuard() { qmb_lut_free(L); }
    } lg;
    qmb_lut *L = lg.L;
    L->ctx = ctx;
    L->device = ctx->device;
    L->nxl = (int32_t)((nn + 15) / 16); L->ny = 4; L->nz = 4; L->S = n_stations;
    L->node_offset = 0; L->n_global = nn; L->n_local = nn;
    {
      DevMem raw;
      const size_t raw_bytes = (size_t)nn * n_stations * sizeof(int32_t);
      QMB_TRY(raw.alloc(raw_bytes));
      QMB_CUDA(cudaMemcpyAsync(raw.p, lookup_tables + (size_t)n0 * n_stations, raw_bytes, cudaMemcpyHostToDevice,
                               ctx->stream));
      QMB_TRY(qmb_lut_build(ctx, L, raw.as<int32_t>()));
    }
    if (L->max_tau > lsmp)
      QMB_FAIL(QMB_ERR_ARG,
               "travel time %d samples > lsmp %d: the reference reads past the onset rows here (post_pad must be "
               ">= the maximum travel time, scan.py:213)",
               L->max_tau, lsmp);
    const size_t map_bytes = (size_t)nn * n_samples * sizeof(double);
    // the reference accumulates into the caller's (zeroed) array: bring its contents in
    QMB_CUDA(cudaMemcpyAsync(map.p, map4d + (size_t)n0 * n_samples, map_bytes, cudaMemcpyHostToDevice, ctx->stream));
    QMB_TRY(qmb_window_run(ctx, L, O, fsmp, lsmp, available, 0, n_samples, nullptr, map.as<double>(), 1));
    QMB_CUDA(cudaMemcpyAsync(map4d + (size_t)n0 * n_samples, map.p, map_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    QMB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return QMB_OK;
}

int find_max_coa_impl(double *map4d, double *max_coa, double *max_norm_coa, int64_t *max_coa_idx, int32_t n_samples,
                      int64_t n_nodes) {
  qmb_ctx *ctx = qmb_default_ctx();
  if (!ctx) return QMB_ERR_CUDA;
  QMB_CUDA(cudaSetDevice(ctx->device));
  int64_t slab = (int64_t)((2ull << 30) / ((size_t)n_samples * sizeof(double)));
  if (slab < 1) slab = 1;
  if (slab > n_nodes) slab = n_nodes;
  const int groups = 64;
  const size_t T = (size_t)n_samples;
  DevMem map, gmax, gsum, gidx, rmax, rsum, ridx, norm;
  QMB_TRY(map.alloc((size_t)slab * T * 8));
  QMB_TRY(gmax.alloc(groups * T * 8)); QMB_TRY(gsum.alloc(groups * T * 8)); QMB_TRY(gidx.alloc(groups * T * 8));
  QMB_TRY(rmax.alloc(T * 8)); QMB_TRY(rsum.alloc(T * 8)); QMB_TRY(ridx.alloc(T * 8)); QMB_TRY(norm.alloc(T * 8));
  const int tb = (n_samples + 127) / 128;
  int first = 1;
  for (int64_t n0 = 0; n0 < n_nodes; n0 += slab) {
    const int64_t nn = (n_nodes - n0 < slab) ? n_nodes - n0 : slab;
    QMB_CUDA(cudaMemcpyAsync(map.p, map4d + (size_t)n0 * T, (size_t)nn * T * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_colreduce<<<dim3(tb, groups), 128, 0, ctx->stream>>>(map.as<double>(), nn, n_samples, n0, groups, gmax.as<double>(),
                                                           gidx.as<long long>(), gsum.as<double>());
    k_colfold<<<tb, 128, 0, ctx->stream>>>(gmax.as<double>(), gidx.as<long long>(), gsum.as<double>(), groups, n_samples,
                                           rmax.as<double>(), ridx.as<long long>(), rsum.as<double>(), first);
    ctx->launches += 2;
    QMB_CUDA(cudaGetLastError());
    first = 0;
  }
  k_colfinish<<<tb, 128, 0, ctx->stream>>>(rmax.as<double>(), rsum.as<double>(), (double)n_nodes, n_samples, norm.as<double>());
  ctx->launches++;
  QMB_CUDA(cudaGetLastError());
  // results land in scratch first: the caller's arrays are written only when everything worked
  std::vector<double> h_max(T), h_norm(T);
  std::vector<int64_t> h_idx(T);
  QMB_CUDA(cudaMemcpyAsync(h_max.data(), rmax.p, T * 8, cudaMemcpyDeviceToHost, ctx->stream));
  QMB_CUDA(cudaMemcpyAsync(h_norm.data(), norm.p, T * 8, cudaMemcpyDeviceToHost, ctx->stream));
  QMB_CUDA(cudaMemcpyAsync(h_idx.data(), ridx.p, T * 8, cudaMemcpyDeviceToHost, ctx->stream));
  QMB_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(max_coa, h_max.data(), T * 8);
  memcpy(max_norm_coa, h_norm.data(), T * 8);
  memcpy(max_coa_idx, h_idx.data(), T * 8);
  return QMB_OK;
}

}  // namespace

// migrate(), quakemigrate/core/src/migratelib.c:40-65.  `onsets` are LOG onsets here (the Python
// wrapper clips and logs before the call, lib.py:93-94).  `threads` is accepted and ignored.
extern "C" void migrate(double *onsets, int32_t *lookup_tables, double *map4d, int32_t fsmp, int32_t lsmp,
                        int32_t n_samples, int32_t n_stations, int32_t available, int64_t n_nodes,
                        int64_t threads) {
  (void)threads;
  g_legacy_status = QMB_OK;
  if (n_nodes <= 0 || n_samples <= 0 || n_stations <= 0) return;
  const int rc = migrate_impl(onsets, lookup_tables, map4d, fsmp, lsmp, n_samples, n_stations, available, n_nodes);
  if (rc != QMB_OK) legacy_fail("migrate", rc);
}

// find_max_coa(), quakemigrate/core/src/migratelib.c:85-111
extern "C" void find_max_coa(double *map4d, double *max_coa, double *max_norm_coa, int64_t *max_coa_idx,
                             int32_t n_samples, int64_t n_nodes, int64_t threads) {
  (void)threads;
  g_legacy_status = QMB_OK;
  if (n_nodes <= 0 || n_samples <= 0) return;
  const int rc = find_max_coa_impl(map4d, max_coa, max_norm_coa, max_coa_idx, n_samples, n_nodes);
  if (rc != QMB_OK) legacy_fail("find_max_coa", rc);
}

static void legacy_stalta(const double *signal, const stalta_header *head, double *onset, int position,
                          const char *name) {
  g_legacy_status = QMB_OK;
  if (!head || head->n <= 0) return;
  qmb_ctx *ctx = qmb_default_ctx();
  const int rc = ctx ? qmb_sta_lta(ctx, signal, 1, head->n, head->nsta, head->nlta, position, onset) : QMB_ERR_CUDA;
  if (rc != QMB_OK) legacy_fail(name, rc);
}

// quakemigrate/core/src/onsetlib.c:35-59
extern "C" void overlapping_sta_lta(const double *signal, const stalta_header *head, double *onset) {
  legacy_stalta(signal, head, onset, QMB_STALTA_CLASSIC, "overlapping_sta_lta");
}
// quakemigrate/core/src/onsetlib.c:79-108
extern "C" void centred_sta_lta(const double *signal, const stalta_header *head, double *onset) {
  legacy_stalta(signal, head, onset, QMB_STALTA_CENTRED, "centred_sta_lta");
}
// quakemigrate/core/src/onsetlib.c:126-148
extern "C" void recursive_sta_lta(const double *signal, const stalta_header *head, double *onset) {
  legacy_stalta(signal, head, onset, QMB_STALTA_RECURSIVE, "recursive_sta_lta");
}
