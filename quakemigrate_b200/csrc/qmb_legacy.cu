// qmb_legacy.cu -- the five symbols of the reference C extension (quakemigrate/core/src/qmlib.h:28-44)
// with byte-identical signatures, implemented on the GPU.  An unmodified quakemigrate.core can
// load libqmb200.so in place of core/src/qmlib<EXT_SUFFIX> (core/libnames.py:35-38).
//
// These honour the reference's ownership contract (the caller allocates every buffer, map4d is
// accumulated into, STA/LTA outputs arrive pre-filled), which forces the full (n_nodes, n_samples)
// float64 volume over PCIe: they are the plumbing tier.  The fast path is the T2 API.
#include <stdlib.h>

#include <mutex>

#include "qmb_internal.cuh"

namespace {

std::mutex g_default_mu;
qmb_ctx *g_default = nullptr;

[[noreturn]] void die(const char *what) {
  fprintf(stderr, "libqmb200: %s: %s\n(no CPU fallback exists in this library)\n", what, qmb_last_error());
  abort();
}

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
    if (qmb_ctx_create(d ? atoi(d) : 0, &c) != QMB_OK) die("cannot create the default CUDA context");
    g_default = c;
  }
  return g_default;
}

// migrate(), quakemigrate/core/src/migratelib.c:40-65.  `onsets` are LOG onsets here (the Python
// wrapper clips and logs before the call, lib.py:93-94).  `threads` is accepted and ignored.
extern "C" void migrate(double *onsets, int32_t *lookup_tables, double *map4d, int32_t fsmp, int32_t lsmp,
                        int32_t n_samples, int32_t n_stations, int32_t available, int64_t n_nodes,
                        int64_t threads) {
  (void)threads;
  if (n_nodes <= 0 || n_samples <= 0 || n_stations <= 0) return;
  qmb_ctx *ctx = qmb_default_ctx();
  if (cudaSetDevice(ctx->device) != cudaSuccess) die("cudaSetDevice");
  const int64_t P = (int64_t)fsmp + lsmp + n_samples;

  // onsets: upload once, skip the clip/log
  qmb_onsets *O = new qmb_onsets();
  O->ctx = ctx; O->device = ctx->device; O->S = n_stations; O->P = P;
  double *d_lin = nullptr;
  const size_t on_bytes = (size_t)n_stations * P * sizeof(double);
  if (cudaMalloc((void **)&d_lin, on_bytes) != cudaSuccess) die("cudaMalloc(onsets)");
  if (cudaMemcpyAsync(d_lin, onsets, on_bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) die("H2D onsets");
  if (qmb_onsets_prepare(ctx, O, d_lin, 1) != QMB_OK) die("onset preparation");
  cudaStreamSynchronize(ctx->stream);
  cudaFree(d_lin);

  // node slabs sized so that the float64 slab of the volume stays under ~2 GiB on the device
  int64_t slab = (int64_t)((2ull << 30) / ((size_t)n_samples * sizeof(double)));
  slab = (slab / 64) * 64;
  if (slab < 64) slab = 64;
  double *d_map = nullptr;
  const int64_t slab_alloc = slab < n_nodes ? slab : n_nodes;
  if (cudaMalloc((void **)&d_map, (size_t)slab_alloc * n_samples * sizeof(double)) != cudaSuccess) die("cudaMalloc(map slab)");

  for (int64_t n0 = 0; n0 < n_nodes; n0 += slab) {
    const int64_t nn = (n_nodes - n0 < slab) ? n_nodes - n0 : slab;
    // pseudo-grid (ceil(nn/16), 4, 4): 64 consecutive flat nodes form one tile
    qmb_lut *L = new qmb_lut();
    L->ctx = ctx;
    L->device = ctx->device;
    L->nxl = (int32_t)((nn + 15) / 16); L->ny = 4; L->nz = 4; L->S = n_stations;
    L->node_offset = 0; L->n_global = nn; L->n_local = nn;
    int32_t *d_raw = nullptr;
    const size_t raw_bytes = (size_t)nn * n_stations * sizeof(int32_t);
    if (cudaMalloc((void **)&d_raw, raw_bytes) != cudaSuccess) die("cudaMalloc(lut slab)");
    if (cudaMemcpyAsync(d_raw, lookup_tables + (size_t)n0 * n_stations, raw_bytes, cudaMemcpyHostToDevice,
                        ctx->stream) != cudaSuccess)
      die("H2D lut");
    if (qmb_lut_build(ctx, L, d_raw) != QMB_OK) die("travel-time table tiling");
    cudaFree(d_raw);
    if (L->max_tau > lsmp) {
      qmb_set_error("travel time %d samples > lsmp %d: the reference reads out of bounds here", L->max_tau, lsmp);
      die("migrate");
    }
    const size_t map_bytes = (size_t)nn * n_samples * sizeof(double);
    // the reference accumulates into the caller's (zeroed) array: bring its contents in
    if (cudaMemcpyAsync(d_map, map4d + (size_t)n0 * n_samples, map_bytes, cudaMemcpyHostToDevice, ctx->stream) !=
        cudaSuccess)
      die("H2D map4d");
    if (qmb_window_run(ctx, L, O, fsmp, lsmp, available, 0, n_samples, nullptr, d_map, 1) != QMB_OK) die("migrate kernel");
    if (cudaMemcpyAsync(map4d + (size_t)n0 * n_samples, d_map, map_bytes, cudaMemcpyDeviceToHost, ctx->stream) !=
        cudaSuccess)
      die("D2H map4d");
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      qmb_set_error("%s", cudaGetErrorString(cudaGetLastError()));
      die("migrate");
    }
    qmb_lut_free(L);
  }
  cudaFree(d_map);
  qmb_onsets_free(O);
}

// find_max_coa(), quakemigrate/core/src/migratelib.c:85-111
extern "C" void find_max_coa(double *map4d, double *max_coa, double *max_norm_coa, int64_t *max_coa_idx,
                             int32_t n_samples, int64_t n_nodes, int64_t threads) {
  (void)threads;
  if (n_nodes <= 0 || n_samples <= 0) return;
  qmb_ctx *ctx = qmb_default_ctx();
  if (cudaSetDevice(ctx->device) != cudaSuccess) die("cudaSetDevice");
  int64_t slab = (int64_t)((2ull << 30) / ((size_t)n_samples * sizeof(double)));
  if (slab < 1) slab = 1;
  if (slab > n_nodes) slab = n_nodes;
  const int groups = 64;
  double *d_map, *d_gmax, *d_gsum, *d_rmax, *d_rsum, *d_norm;
  long long *d_gidx, *d_ridx;
  const size_t T = (size_t)n_samples;
  if (cudaMalloc((void **)&d_map, (size_t)slab * T * 8) != cudaSuccess) die("cudaMalloc(map slab)");
  if (cudaMalloc((void **)&d_gmax, groups * T * 8) != cudaSuccess || cudaMalloc((void **)&d_gsum, groups * T * 8) != cudaSuccess ||
      cudaMalloc((void **)&d_gidx, groups * T * 8) != cudaSuccess || cudaMalloc((void **)&d_rmax, T * 8) != cudaSuccess ||
      cudaMalloc((void **)&d_rsum, T * 8) != cudaSuccess || cudaMalloc((void **)&d_ridx, T * 8) != cudaSuccess ||
      cudaMalloc((void **)&d_norm, T * 8) != cudaSuccess)
    die("cudaMalloc(find_max_coa scratch)");
  const int tb = (n_samples + 127) / 128;
  int first = 1;
  for (int64_t n0 = 0; n0 < n_nodes; n0 += slab) {
    const int64_t nn = (n_nodes - n0 < slab) ? n_nodes - n0 : slab;
    if (cudaMemcpyAsync(d_map, map4d + (size_t)n0 * T, (size_t)nn * T * 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
      die("H2D map4d");
    k_colreduce<<<dim3(tb, groups), 128, 0, ctx->stream>>>(d_map, nn, n_samples, n0, groups, d_gmax, d_gidx, d_gsum);
    k_colfold<<<tb, 128, 0, ctx->stream>>>(d_gmax, d_gidx, d_gsum, groups, n_samples, d_rmax, d_ridx, d_rsum, first);
    ctx->launches += 2;
    first = 0;
  }
  k_colfinish<<<tb, 128, 0, ctx->stream>>>(d_rmax, d_rsum, (double)n_nodes, n_samples, d_norm);
  ctx->launches++;
  cudaMemcpyAsync(max_coa, d_rmax, T * 8, cudaMemcpyDeviceToHost, ctx->stream);
  cudaMemcpyAsync(max_norm_coa, d_norm, T * 8, cudaMemcpyDeviceToHost, ctx->stream);
  cudaMemcpyAsync(max_coa_idx, d_ridx, T * 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    qmb_set_error("%s", cudaGetErrorString(cudaGetLastError()));
    die("find_max_coa");
  }
  cudaFree(d_map); cudaFree(d_gmax); cudaFree(d_gsum); cudaFree(d_gidx); cudaFree(d_rmax); cudaFree(d_rsum);
  cudaFree(d_ridx); cudaFree(d_norm);
}

static void legacy_stalta(const double *signal, const stalta_header *head, double *onset, int position,
                          const char *name) {
  if (!head || head->n <= 0) return;
  qmb_ctx *ctx = qmb_default_ctx();
  if (qmb_sta_lta(ctx, signal, 1, head->n, head->nsta, head->nlta, position, onset) != QMB_OK) die(name);
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
