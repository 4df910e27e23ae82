// qmb_api.cu -- the C ABI of libqmb200.so (see include/qmb200.h): contexts, NCCL plumbing,
// table / onset upload, detect, locate window, and the five reference-compatible symbols.
#include <dlfcn.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "qmb_internal.cuh"

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void qmb_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char *qmb_last_error(void) { return g_err; }
extern "C" const char *qmb_version(void) { return "qmb200 0.1 (sm_100a)"; }

extern "C" int qmb_device_count(int *count) {
  if (!count) QMB_FAIL(QMB_ERR_ARG, "null argument");
  *count = 0;
  QMB_CUDA(cudaGetDeviceCount(count));
  return QMB_OK;
}

// ---------------------------------------------------------------------------------------------
// recycled device buffers (see DevBuf in qmb_internal.cuh)
// ---------------------------------------------------------------------------------------------
namespace {
struct PoolEntry {
  int device;
  size_t bytes;
  void *p;
};
std::mutex g_pool_mu;
std::vector<PoolEntry> g_pool;
size_t g_pool_bytes = 0;
constexpr size_t POOL_MIN_BYTES = 1u << 20;
constexpr size_t POOL_MAX_ENTRIES = 32;
size_t pool_cap_bytes() {
  static const size_t cap = [] {
    const char *env = getenv("QMB_POOL_MB");
    const long long mb = env ? atoll(env) : 8192;
    return (size_t)(mb < 0 ? 0 : mb) << 20;
  }();
  return cap;
}
}  // namespace

void *qmb_pool_take(size_t bytes) {
  if (bytes < POOL_MIN_BYTES) return nullptr;
  int device = -1;
  if (cudaGetDevice(&device) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(g_pool_mu);
  for (size_t i = 0; i < g_pool.size(); ++i)
    if (g_pool[i].device == device && g_pool[i].bytes == bytes) {
      void *p = g_pool[i].p;
      g_pool_bytes -= bytes;
      g_pool.erase(g_pool.begin() + i);
      return p;
    }
  return nullptr;
}

bool qmb_pool_give(void *p, size_t bytes) {
  if (!p || bytes < POOL_MIN_BYTES || bytes > pool_cap_bytes()) return false;
  int device = -1;
  if (cudaGetDevice(&device) != cudaSuccess) return false;
  // like cudaFree: nothing queued on the buffer survives the hand-back
  if (cudaDeviceSynchronize() != cudaSuccess) return false;
  std::lock_guard<std::mutex> lock(g_pool_mu);
  // make room: drop the oldest entries (they are the least likely to be asked for again)
  while (!g_pool.empty() && (g_pool.size() >= POOL_MAX_ENTRIES || g_pool_bytes + bytes > pool_cap_bytes())) {
    if (g_pool[0].device == device) cudaFree(g_pool[0].p);
    else {
      cudaSetDevice(g_pool[0].device);
      cudaFree(g_pool[0].p);
      cudaSetDevice(device);
    }
    g_pool_bytes -= g_pool[0].bytes;
    g_pool.erase(g_pool.begin());
  }
  g_pool.push_back({device, bytes, p});
  g_pool_bytes += bytes;
  return true;
}

void qmb_pool_clear(int device) {
  int cur = -1;
  cudaGetDevice(&cur);
  std::lock_guard<std::mutex> lock(g_pool_mu);
  for (size_t i = 0; i < g_pool.size();) {
    if (g_pool[i].device == device) {
      cudaSetDevice(device);
      cudaFree(g_pool[i].p);
      g_pool_bytes -= g_pool[i].bytes;
      g_pool.erase(g_pool.begin() + i);
    } else {
      ++i;
    }
  }
  if (cur >= 0) cudaSetDevice(cur);
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
extern "C" int qmb_ctx_create(int device, qmb_ctx_t *out) {
  if (!out) QMB_FAIL(QMB_ERR_ARG, "null argument");
  *out = nullptr;
  int n = 0;
  QMB_CUDA(cudaGetDeviceCount(&n));
  if (n <= 0) QMB_FAIL(QMB_ERR_CUDA, "no CUDA device visible (libqmb200 has no CPU fallback)");
  if (device < 0 || device >= n) QMB_FAIL(QMB_ERR_ARG, "device %d out of range [0, %d)", device, n);
  QMB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  QMB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    QMB_FAIL(QMB_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device,
             prop.major, prop.minor);
  qmb_ctx *ctx = new qmb_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  QMB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  for (int i = 0; i < 4; ++i) QMB_CUDA(cudaEventCreate(&ctx->ev[i]));
  if (const char *pr = getenv("QMB_SCAN_PRECISION"))
    ctx->scan_fp32 = strcmp(pr, "fp32") == 0 ? 1 : (strcmp(pr, "q16") == 0 ? 2 : 0);
  if (const char *g = getenv("QMB_SCAN_QGROUP")) {
    const int gi = atoi(g);
    if (gi == 1 || gi == 2 || gi == 4 || gi == 8) ctx->scan_qgroup = gi;
  }
  const char *v = getenv("QMB_VARIANT");
  if (v && atoi(v) >= 0 && atoi(v) < QMB_N_VARIANTS) ctx->variant = atoi(v);
  *out = ctx;
  return QMB_OK;
}

extern "C" int qmb_ctx_set_variant(qmb_ctx_t ctx, int variant) {
  if (!ctx) QMB_FAIL(QMB_ERR_ARG, "null context");
  if (variant < 0) variant = QMB_DEFAULT_VARIANT;
  if (variant >= QMB_N_VARIANTS) QMB_FAIL(QMB_ERR_ARG, "variant must be in [0, %d)", QMB_N_VARIANTS);
  ctx->variant = variant;
  return QMB_OK;
}

extern "C" int qmb_ctx_set_scan_precision(qmb_ctx_t ctx, int mode) {
  if (!ctx) QMB_FAIL(QMB_ERR_ARG, "null context");
  if (mode < 0 || mode > 2) QMB_FAIL(QMB_ERR_ARG, "precision mode must be 0 (auto), 1 (float32) or 2 (16-bit)");
  ctx->scan_fp32 = mode;
  return QMB_OK;
}

extern "C" int qmb_ctx_set_scan_qgroup(qmb_ctx_t ctx, int group) {
  if (!ctx) QMB_FAIL(QMB_ERR_ARG, "null context");
  if (group <= 0) group = QMB_DEFAULT_QGROUP;
  if (group != 1 && group != 2 && group != 4 && group != 8) QMB_FAIL(QMB_ERR_ARG, "row group must be 1, 2, 4 or 8");
  ctx->scan_qgroup = group;
  return QMB_OK;
}

extern "C" int qmb_ctx_last_scan_bytes_per_gather(qmb_ctx_t ctx, int *bytes) {
  if (!ctx || !bytes) QMB_FAIL(QMB_ERR_ARG, "null argument");
  *bytes = ctx->last_scan_q16 ? 2 : 4;
  return QMB_OK;
}

extern "C" int qmb_ctx_last_scan_info(qmb_ctx_t ctx, qmb_scan_info_t *info) {
  if (!ctx || !info) QMB_FAIL(QMB_ERR_ARG, "null argument");
  if (ctx->last_n_samples <= 0) QMB_FAIL(QMB_ERR_STATE, "no detect scan has run on this context");
  *info = ctx->info;
  return QMB_OK;
}

extern "C" int qmb_ctx_launch_count(qmb_ctx_t ctx, int64_t *count) {
  if (!ctx || !count) QMB_FAIL(QMB_ERR_ARG, "null argument");
  *count = ctx->launches;
  return QMB_OK;
}

extern "C" int qmb_ctx_sync(qmb_ctx_t ctx) {
  if (!ctx) QMB_FAIL(QMB_ERR_ARG, "null context");
  QMB_CUDA(cudaSetDevice(ctx->device));
  QMB_CUDA(cudaStreamSynchronize(ctx->stream));
  return QMB_OK;
}

// CUDA-event stopwatch on the context's stream (the stream every kernel of this library runs on)
extern "C" int qmb_ctx_timer_start(qmb_ctx_t ctx) {
  if (!ctx) QMB_FAIL(QMB_ERR_ARG, "null context");
  QMB_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->ev_user[0]) {
    QMB_CUDA(cudaEventCreate(&ctx->ev_user[0]));
    QMB_CUDA(cudaEventCreate(&ctx->ev_user[1]));
  }
  QMB_CUDA(cudaEventRecord(ctx->ev_user[0], ctx->stream));
  return QMB_OK;
}

extern "C" int qmb_ctx_timer_stop(qmb_ctx_t ctx, float *ms) {
  if (!ctx || !ms) QMB_FAIL(QMB_ERR_ARG, "null argument");
  if (!ctx->ev_user[0]) QMB_FAIL(QMB_ERR_STATE, "timer was not started");
  QMB_CUDA(cudaSetDevice(ctx->device));
  QMB_CUDA(cudaEventRecord(ctx->ev_user[1], ctx->stream));
  QMB_CUDA(cudaEventSynchronize(ctx->ev_user[1]));
  QMB_CUDA(cudaEventElapsedTime(ms, ctx->ev_user[0], ctx->ev_user[1]));
  return QMB_OK;
}

extern "C" int qmb_ctx_flush_l2(qmb_ctx_t ctx) {
  if (!ctx) QMB_FAIL(QMB_ERR_ARG, "null context");
  QMB_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = 256ull << 20;
  QMB_TRY(ctx->flush.ensure(bytes));
  QMB_CUDA(cudaMemsetAsync(ctx->flush.p, (int)(ctx->launches & 0xff), bytes, ctx->stream));
  QMB_CUDA(cudaStreamSynchronize(ctx->stream));
  return QMB_OK;
}

extern "C" int qmb_host_register(void *ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) QMB_FAIL(QMB_ERR_ARG, "bad host range");
  QMB_CUDA(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault));
  return QMB_OK;
}

extern "C" int qmb_host_unregister(void *ptr) {
  if (!ptr) return QMB_OK;
  QMB_CUDA(cudaHostUnregister(ptr));
  return QMB_OK;
}

extern "C" int qmb_comm_destroy(qmb_ctx_t ctx);

extern "C" int qmb_ctx_destroy(qmb_ctx_t ctx) {
  if (!ctx) return QMB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  qmb_comm_destroy(ctx);
  ctx->pm1.release(); ctx->pm2.release(); ctx->pm3.release(); ctx->pi1.release(); ctx->pi2.release();
  ctx->psum.release(); ctx->st_loc.release(); ctx->st_glob.release(); ctx->st_grp.release(); ctx->st_chk.release();
  ctx->ex.release(); ctx->gather.release(); ctx->slow_flag.release(); ctx->status.release();
  if (ctx->h_status) cudaFreeHost(ctx->h_status);
  ctx->r_max.release(); ctx->r_norm.release(); ctx->r_idx.release(); ctx->slow_list.release();
  ctx->slow_val.release(); ctx->slow_idx.release(); ctx->flush.release(); ctx->stage_lin.release();
  ctx->q_i32.release(); ctx->q_f64.release(); ctx->map_scratch.release(); ctx->pp_scratch.release(); ctx->pp_sel.release();
  qmb_fft_release(ctx);
  if (ctx->tmp_onsets) {
    ctx->tmp_onsets->log32.release(); ctx->tmp_onsets->log64.release(); ctx->tmp_onsets->rowmax.release();
    ctx->tmp_onsets->q16.release(); ctx->tmp_onsets->qrange.release();
    delete ctx->tmp_onsets;
    ctx->tmp_onsets = nullptr;
  }
  qmb_pool_clear(ctx->device);
  for (int i = 0; i < 4; ++i)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (int i = 0; i < 2; ++i)
    if (ctx->ev_user[i]) cudaEventDestroy(ctx->ev_user[i]);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return QMB_OK;
}

// ---------------------------------------------------------------------------------------------
// NCCL, loaded lazily so that single-GPU use has no dependency on it
// ---------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

int nccl_load() {
  std::lock_guard<std::mutex> lock(g_nccl_mu);
  if (g_nccl.handle) return QMB_OK;
  const char *env = getenv("QMB_NCCL_LIB");
  const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *nm : names) {
    if (!nm) continue;
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) QMB_FAIL(QMB_ERR_NCCL, "cannot dlopen libnccl.so.2 (%s); set QMB_NCCL_LIB", dlerror());
#define QMB_SYM(field, name)                                                   \
  *(void **)(&g_nccl.field) = dlsym(h, name);                                  \
  if (!g_nccl.field) QMB_FAIL(QMB_ERR_NCCL, "libnccl lacks symbol %s", name)
  QMB_SYM(GetUniqueId, "ncclGetUniqueId");
  QMB_SYM(CommInitRank, "ncclCommInitRank");
  QMB_SYM(CommDestroy, "ncclCommDestroy");
  QMB_SYM(AllGather, "ncclAllGather");
  QMB_SYM(AllReduce, "ncclAllReduce");
  QMB_SYM(GetErrorString, "ncclGetErrorString");
#undef QMB_SYM
  g_nccl.handle = h;
  return QMB_OK;
}
}  // namespace

#define QMB_NCCL(call)                                                                          \
  do {                                                                                          \
    ncclResult_t r__ = (call);                                                                  \
    if (r__ != ncclSuccess) {                                                                   \
      qmb_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__));  \
      return QMB_ERR_NCCL;                                                                      \
    }                                                                                           \
  } while (0)

extern "C" int qmb_comm_unique_id(void *id_out) {
  if (!id_out) QMB_FAIL(QMB_ERR_ARG, "null argument");
  static_assert(sizeof(ncclUniqueId) <= QMB_NCCL_ID_BYTES, "id size");
  QMB_TRY(nccl_load());
  ncclUniqueId id;
  QMB_NCCL(g_nccl.GetUniqueId(&id));
  memset(id_out, 0, QMB_NCCL_ID_BYTES);
  memcpy(id_out, &id, sizeof(id));
  return QMB_OK;
}

extern "C" int qmb_comm_init(qmb_ctx_t ctx, const void *id, int world, int rank) {
  if (!ctx || !id) QMB_FAIL(QMB_ERR_ARG, "null argument");
  if (world < 1 || rank < 0 || rank >= world) QMB_FAIL(QMB_ERR_ARG, "bad world/rank %d/%d", rank, world);
  QMB_TRY(nccl_load());
  QMB_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  ncclComm_t comm;
  QMB_NCCL(g_nccl.CommInitRank(&comm, world, uid, rank));
  ctx->nccl_comm = comm;
  ctx->world = world;
  ctx->rank = rank;
  return QMB_OK;
}

extern "C" int qmb_comm_destroy(qmb_ctx_t ctx) {
  if (!ctx || !ctx->nccl_comm) return QMB_OK;
  g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
  ctx->world = 1;
  ctx->rank = 0;
  return QMB_OK;
}

int qmb_nccl_allgather(qmb_ctx *ctx, const void *send, void *recv, size_t bytes_per_rank) {
  if (!ctx->nccl_comm) QMB_FAIL(QMB_ERR_STATE, "no communicator attached");
  QMB_NCCL(g_nccl.AllGather(send, recv, bytes_per_rank, ncclChar, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  return QMB_OK;
}

int qmb_nccl_allreduce_u32(qmb_ctx *ctx, unsigned int *buf, size_t count, int take_max) {
  if (!ctx->nccl_comm) QMB_FAIL(QMB_ERR_STATE, "no communicator attached");
  QMB_NCCL(g_nccl.AllReduce(buf, buf, count, ncclUint32, take_max ? ncclMax : ncclMin, (ncclComm_t)ctx->nccl_comm,
                            ctx->stream));
  return QMB_OK;
}

int qmb_nccl_allreduce_sum_f64(qmb_ctx *ctx, double *buf, size_t count) {
  if (!ctx->nccl_comm) QMB_FAIL(QMB_ERR_STATE, "no communicator attached");
  QMB_NCCL(g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  return QMB_OK;
}

// ---------------------------------------------------------------------------------------------
// travel-time table
// ---------------------------------------------------------------------------------------------
// shared by the two upload forms: raw rows of the slab -> device, tile, free the raw copy
static int lut_upload_common(qmb_ctx_t ctx, qmb_lut *L, const int32_t *lut, qmb_lut_t *out) {
  int32_t *d_raw = nullptr;
  const size_t raw_bytes = (size_t)L->n_local * L->S * sizeof(int32_t);
  cudaError_t e = cudaMalloc((void **)&d_raw, raw_bytes);
  if (e != cudaSuccess) {
    delete L;
    QMB_FAIL(QMB_ERR_CUDA, "cudaMalloc(%zu bytes) for the raw table: %s", raw_bytes, cudaGetErrorString(e));
  }
  int rc = QMB_OK;
  e = cudaMemcpyAsync(d_raw, lut, raw_bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) {
    qmb_set_error("H2D of the travel-time table: %s", cudaGetErrorString(e));
    rc = QMB_ERR_CUDA;
  }
  if (rc == QMB_OK) rc = qmb_lut_build(ctx, L, d_raw);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(d_raw);
  if (rc != QMB_OK) {
    qmb_lut_free(L);
    return rc;
  }
  *out = L;
  return QMB_OK;
}

extern "C" int qmb_lut_upload(qmb_ctx_t ctx, const int32_t *lut, int32_t nx_local, int32_t ny, int32_t nz,
                              int32_t n_rows, int64_t node_offset, int64_t n_nodes_global, qmb_lut_t *out) {
  if (!ctx || !lut || !out) QMB_FAIL(QMB_ERR_ARG, "null argument");
  *out = nullptr;
  if (nx_local <= 0 || ny <= 0 || nz <= 0 || n_rows <= 0) QMB_FAIL(QMB_ERR_ARG, "empty travel-time table");
  const int64_t n_local = (int64_t)nx_local * ny * nz;
  if (n_nodes_global <= 0) n_nodes_global = n_local;
  if (node_offset < 0 || node_offset + n_local > n_nodes_global)
    QMB_FAIL(QMB_ERR_ARG, "node slab [%lld, %lld) outside the global grid of %lld nodes", (long long)node_offset,
             (long long)(node_offset + n_local), (long long)n_nodes_global);
  if (n_nodes_global >= 0xFFFFFFFFLL) QMB_FAIL(QMB_ERR_UNSUPPORTED, "more than 2^32-2 nodes");
  QMB_CUDA(cudaSetDevice(ctx->device));
  qmb_lut *L = new qmb_lut();
  L->ctx = ctx;
  L->device = ctx->device;
  L->nxl = nx_local; L->ny = ny; L->nz = nz; L->S = n_rows;
  L->node_offset = node_offset; L->n_global = n_nodes_global; L->n_local = n_local;
  return lut_upload_common(ctx, L, lut, out);
}

// A slab given as a range of the C-order node list: [node_first, node_first + n_nodes_local), whole
// z-columns (both multiples of nz).  It may begin and end inside an x-plane, so that G ranks can
// hold N / G nodes each whatever nx is (x-plane slabs of a 201-plane grid over 8 ranks are 26 or 25
// planes: 3.5 % imbalance).
extern "C" int qmb_lut_upload_range(qmb_ctx_t ctx, const int32_t *lut, int32_t ny, int32_t nz, int32_t n_rows,
                                    int64_t node_first, int64_t n_nodes_local, int64_t n_nodes_global,
                                    qmb_lut_t *out) {
  if (!ctx || !lut || !out) QMB_FAIL(QMB_ERR_ARG, "null argument");
  *out = nullptr;
  if (ny <= 0 || nz <= 0 || n_rows <= 0 || n_nodes_local <= 0) QMB_FAIL(QMB_ERR_ARG, "empty travel-time table");
  const int64_t plane = (int64_t)ny * nz;
  if (node_first < 0 || node_first % nz != 0 || n_nodes_local % nz != 0)
    QMB_FAIL(QMB_ERR_ARG, "a node range must consist of whole z-columns (multiples of nz = %d)", nz);
  if (n_nodes_global <= 0 || n_nodes_global % plane != 0 || node_first + n_nodes_local > n_nodes_global)
    QMB_FAIL(QMB_ERR_ARG, "node range [%lld, %lld) outside the global grid of %lld nodes", (long long)node_first,
             (long long)(node_first + n_nodes_local), (long long)n_nodes_global);
  if (n_nodes_global >= 0xFFFFFFFFLL) QMB_FAIL(QMB_ERR_UNSUPPORTED, "more than 2^32-2 nodes");
  QMB_CUDA(cudaSetDevice(ctx->device));
  const int64_t last = node_first + n_nodes_local;
  const int64_t xa = node_first / plane, xb = (last + plane - 1) / plane;
  qmb_lut *L = new qmb_lut();
  L->ctx = ctx;
  L->device = ctx->device;
  L->nxl = (int32_t)(xb - xa); L->ny = ny; L->nz = nz; L->S = n_rows;
  L->y_first = (int32_t)((node_first % plane) / nz);
  L->y_last = (int32_t)((last % plane) / nz);
  L->node_offset = node_first; L->n_global = n_nodes_global; L->n_local = n_nodes_local;
  return lut_upload_common(ctx, L, lut, out);
}

// --- serving from device-resident float64 grids (LUT.serve_traveltimes on the device) ------------
extern "C" int qmb_ttgrids_upload(qmb_ctx_t ctx, const double *grids, int32_t n_grids, int32_t nx, int32_t ny,
                                  int32_t nz, qmb_ttgrids_t *out) {
  if (!ctx || !grids || !out) QMB_FAIL(QMB_ERR_ARG, "null argument");
  *out = nullptr;
  if (n_grids <= 0 || nx <= 0 || ny <= 0 || nz <= 0) QMB_FAIL(QMB_ERR_ARG, "empty travel-time grids");
  QMB_CUDA(cudaSetDevice(ctx->device));
  qmb_ttgrids *G = new qmb_ttgrids();
  G->ctx = ctx; G->device = ctx->device; G->n_grids = n_grids; G->nx = nx; G->ny = ny; G->nz = nz;
  const size_t count = (size_t)n_grids * nx * ny * nz;
  int rc = G->tt.ensure(count);
  if (rc == QMB_OK) {
    cudaError_t e = cudaMemcpyAsync(G->tt.p, grids, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      qmb_set_error("H2D of the travel-time grids: %s", cudaGetErrorString(e));
      rc = QMB_ERR_CUDA;
    }
  }
  if (rc != QMB_OK) {
    G->tt.release();
    delete G;
    return rc;
  }
  *out = G;
  return QMB_OK;
}

extern "C" int qmb_ttgrids_free(qmb_ttgrids_t G) {
  if (!G) return QMB_OK;
  cudaSetDevice(G->device);
  G->tt.release();
  delete G;
  return QMB_OK;
}

extern "C" int qmb_lut_serve(qmb_ctx_t ctx, qmb_ttgrids_t G, const int32_t *rows, int32_t n_rows,
                             double sampling_rate, int32_t x0, int32_t x1, qmb_lut_t *out) {
  if (!ctx || !G || !rows || !out) QMB_FAIL(QMB_ERR_ARG, "null argument");
  *out = nullptr;
  if (G->ctx != ctx) QMB_FAIL(QMB_ERR_ARG, "grids belong to a different context");
  if (n_rows <= 0) QMB_FAIL(QMB_ERR_ARG, "no rows selected");
  if (x1 <= 0) x1 = G->nx;
  if (x0 < 0 || x0 >= x1 || x1 > G->nx) QMB_FAIL(QMB_ERR_ARG, "x-slab [%d, %d) outside [0, %d)", x0, x1, G->nx);
  for (int r = 0; r < n_rows; ++r)
    if (rows[r] < 0 || rows[r] >= G->n_grids) QMB_FAIL(QMB_ERR_ARG, "row %d selects grid %d of %d", r, rows[r], G->n_grids);
  if (!(sampling_rate > 0)) QMB_FAIL(QMB_ERR_ARG, "sampling rate must be positive");
  QMB_CUDA(cudaSetDevice(ctx->device));
  const int64_t plane = (int64_t)G->ny * G->nz;
  qmb_lut *L = new qmb_lut();
  L->ctx = ctx; L->device = ctx->device;
  L->nxl = x1 - x0; L->ny = G->ny; L->nz = G->nz; L->S = n_rows;
  L->node_offset = x0 * plane; L->n_global = (int64_t)G->nx * plane; L->n_local = (int64_t)(x1 - x0) * plane;
  if (L->n_global >= 0xFFFFFFFFLL) {
    delete L;
    QMB_FAIL(QMB_ERR_UNSUPPORTED, "more than 2^32-2 nodes");
  }
  int *d_rows = nullptr;
  cudaError_t e = cudaMalloc((void **)&d_rows, n_rows * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_rows, rows, n_rows * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
  int rc = QMB_OK;
  if (e != cudaSuccess) {
    qmb_set_error("row selection upload: %s", cudaGetErrorString(e));
    rc = QMB_ERR_CUDA;
  }
  QmbServeSource src = {G->tt.p, d_rows, sampling_rate, (long long)L->n_global, (long long)L->node_offset};
  if (rc == QMB_OK) rc = qmb_lut_build(ctx, L, nullptr, &src);
  cudaStreamSynchronize(ctx->stream);
  if (d_rows) cudaFree(d_rows);
  if (rc != QMB_OK) {
    qmb_lut_free(L);
    return rc;
  }
  *out = L;
  return QMB_OK;
}

extern "C" int qmb_lut_free(qmb_lut_t L) {
  if (!L) return QMB_OK;
  cudaSetDevice(L->device);
  L->tau16.release(); L->tmin.release(); L->span.release(); L->nodeidx.release();
  L->boxes.release(); L->slotmap.release(); L->tauq16.release(); L->tauq32.release(); L->tauq16p.release();
  delete L;
  return QMB_OK;
}

extern "C" int qmb_lut_info(qmb_lut_t L, qmb_lut_info_t *info) {
  if (!L || !info) QMB_FAIL(QMB_ERR_ARG, "null argument");
  info->n_nodes_local = L->n_local; info->n_nodes_global = L->n_global; info->node_offset = L->node_offset;
  info->n_rows = L->S; info->max_tau = L->max_tau; info->max_tile_span = L->max_span;
  info->n_tiles = L->n_tiles; info->nodes_per_tile = L->NB;
  info->duplicate_fraction = L->dup_frac; info->variant = L->v.id; info->reserved = 0;
  info->device_bytes = (int64_t)(L->tau16.bytes() + L->tmin.bytes() + L->span.bytes() + L->nodeidx.bytes() +
                                  L->boxes.bytes() + L->slotmap.bytes());
  return QMB_OK;
}

// ---------------------------------------------------------------------------------------------
// onsets
// ---------------------------------------------------------------------------------------------
extern "C" int qmb_onsets_free(qmb_onsets_t O) {
  if (!O) return QMB_OK;
  cudaSetDevice(O->device);
  O->log32.release(); O->log64.release(); O->rowmax.release(); O->q16.release(); O->qrange.release();
  delete O;
  return QMB_OK;
}

// H2D through the context's reusable staging buffer, then clip/log into O's (reusable) rows.
static int onsets_fill(qmb_ctx *ctx, qmb_onsets *O, const double *onsets, int32_t n_rows, int64_t n_total) {
  const size_t count = (size_t)n_rows * n_total;
  QMB_TRY(ctx->stage_lin.ensure(count));
  O->S = n_rows;
  O->P = n_total;
  QMB_CUDA(cudaMemcpyAsync(ctx->stage_lin.p, onsets, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  QMB_TRY(qmb_onsets_prepare(ctx, O, ctx->stage_lin.p, 0));
  QMB_CUDA(cudaStreamSynchronize(ctx->stream));  // the staging buffer may be reused right away
  return QMB_OK;
}

extern "C" int qmb_onsets_upload(qmb_ctx_t ctx, const double *onsets, int32_t n_rows, int64_t n_total,
                                 qmb_onsets_t *out) {
  if (!ctx || !onsets || !out) QMB_FAIL(QMB_ERR_ARG, "null argument");
  *out = nullptr;
  if (n_rows <= 0 || n_total <= 0) QMB_FAIL(QMB_ERR_ARG, "empty onset array");
  QMB_CUDA(cudaSetDevice(ctx->device));
  qmb_onsets *O = new qmb_onsets();
  O->ctx = ctx;
  O->device = ctx->device;
  int rc = onsets_fill(ctx, O, onsets, n_rows, n_total);
  if (rc != QMB_OK) {
    qmb_onsets_free(O);
    return rc;
  }
  *out = O;
  return QMB_OK;
}

// `waveforms` is a host array (dev_wf == nullptr) or traces already resident on the device.
static int onsets_from_traces(qmb_ctx_t ctx, const double *waveforms, const double *dev_wf, int32_t n_channels,
                              int64_t n_total, int32_t n_rows, const int32_t *row_ptr,
                              const int32_t *row_chan, const int32_t *nsta, const int32_t *nlta,
                              int32_t position, int32_t transform, double min_onset_value,
                              double *onsets_out, qmb_onsets_t *out) {
  if (!ctx || (!waveforms && !dev_wf) || !row_ptr || !row_chan || !nsta || !nlta || !out)
    QMB_FAIL(QMB_ERR_ARG, "null argument");
  *out = nullptr;
  if (n_channels <= 0 || n_total <= 0 || n_rows <= 0) QMB_FAIL(QMB_ERR_ARG, "empty input");
  if (position != QMB_STALTA_CLASSIC && position != QMB_STALTA_CENTRED)
    QMB_FAIL(QMB_ERR_ARG, "position must be classic (0) or centred (1)");
  const int n_pairs = row_ptr[n_rows];
  if (row_ptr[0] != 0 || n_pairs <= 0) QMB_FAIL(QMB_ERR_ARG, "malformed row_ptr");
  std::vector<int> pair_nsta(n_pairs), pair_nlta(n_pairs);
  int max_nsta = 1, max_nlta = 1;
  for (int r = 0; r < n_rows; ++r) {
    if (row_ptr[r + 1] <= row_ptr[r]) QMB_FAIL(QMB_ERR_ARG, "row %d has no channels", r);
    if (nsta[r] < 1 || nlta[r] < nsta[r]) QMB_FAIL(QMB_ERR_ARG, "row %d: need 1 <= nsta <= nlta", r);
    for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k) {
      if (row_chan[k] < 0 || row_chan[k] >= n_channels) QMB_FAIL(QMB_ERR_ARG, "channel index out of range");
      pair_nsta[k] = nsta[r];
      pair_nlta[k] = nlta[r];
    }
    if (nsta[r] > max_nsta) max_nsta = nsta[r];
    if (nlta[r] > max_nlta) max_nlta = nlta[r];
  }
  QMB_CUDA(cudaSetDevice(ctx->device));
  double *d_wf = nullptr, *d_pair = nullptr, *d_lin = nullptr, *d_env = nullptr;
  int *d_meta = nullptr;  // [row_ptr (n_rows+1)][row_chan (n_pairs)][nsta (n_pairs)][nlta (n_pairs)]
  const size_t wf_bytes = (size_t)n_channels * n_total * sizeof(double);
  const size_t pair_elems = (size_t)n_pairs * n_total;
  int rc = QMB_OK;
  qmb_onsets *O = new qmb_onsets();
  O->ctx = ctx; O->device = ctx->device; O->S = n_rows; O->P = n_total;
  auto cleanup = [&]() {
    cudaStreamSynchronize(ctx->stream);
    if (d_wf) cudaFree(d_wf);
    if (d_env) cudaFree(d_env);
    if (d_pair) cudaFree(d_pair);
    if (d_lin) cudaFree(d_lin);
    if (d_meta) cudaFree(d_meta);
  };
#define QMB_STEP(expr)                          \
  if (rc == QMB_OK) {                           \
    rc = [&]() -> int { expr; return QMB_OK; }(); \
  }
  if (!dev_wf) QMB_STEP(QMB_CUDA(cudaMalloc((void **)&d_wf, wf_bytes)));
  QMB_STEP(QMB_CUDA(cudaMalloc((void **)&d_pair, pair_elems * sizeof(double))));
  QMB_STEP(QMB_CUDA(cudaMalloc((void **)&d_meta, (size_t)(n_rows + 1 + 3 * n_pairs) * sizeof(int))));
  if (onsets_out) QMB_STEP(QMB_CUDA(cudaMalloc((void **)&d_lin, (size_t)n_rows * n_total * sizeof(double))));
  if (!dev_wf) QMB_STEP(QMB_CUDA(cudaMemcpyAsync(d_wf, waveforms, wf_bytes, cudaMemcpyHostToDevice, ctx->stream)));
  const double *d_src = dev_wf ? dev_wf : d_wf;
  if (transform == QMB_TRANSFORM_ENV || transform == QMB_TRANSFORM_ENV_SQUARED) {
    // np.abs(hilbert(tr.data)) [** 2] (stalta.py:520-523): envelopes first, then |.| or (.)^2 is
    // the identity or the energy transform of the STA/LTA kernels
    QMB_STEP(QMB_CUDA(cudaMalloc((void **)&d_env, wf_bytes)));
    QMB_STEP(QMB_TRY(qmb_envelope_device(ctx, d_src, n_total, n_channels, n_total, d_env, n_total)));
    d_src = d_env;
    transform = transform == QMB_TRANSFORM_ENV ? QMB_TRANSFORM_NONE : QMB_TRANSFORM_ENERGY;
  }
  int *d_row_ptr = d_meta, *d_row_chan = d_meta ? d_meta + n_rows + 1 : nullptr;
  int *d_nsta = d_meta ? d_row_chan + n_pairs : nullptr, *d_nlta = d_meta ? d_nsta + n_pairs : nullptr;
  QMB_STEP(QMB_CUDA(cudaMemcpyAsync(d_row_ptr, row_ptr, (n_rows + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream)));
  QMB_STEP(QMB_CUDA(cudaMemcpyAsync(d_row_chan, row_chan, n_pairs * sizeof(int), cudaMemcpyHostToDevice, ctx->stream)));
  QMB_STEP(QMB_CUDA(cudaMemcpyAsync(d_nsta, pair_nsta.data(), n_pairs * sizeof(int), cudaMemcpyHostToDevice, ctx->stream)));
  QMB_STEP(QMB_CUDA(cudaMemcpyAsync(d_nlta, pair_nlta.data(), n_pairs * sizeof(int), cudaMemcpyHostToDevice, ctx->stream)));
  // pre-fill 1.0 like the reference wrappers (lib.py:204,242), then the valid range
  QMB_STEP(QMB_TRY(qmb_fill_f64(ctx, d_pair, 1.0, pair_elems)));
  QMB_STEP(QMB_TRY(qmb_stalta_device_ex(ctx, d_src, n_pairs, n_total, n_total, n_total, max_nsta, max_nlta,
                                        position, transform, d_pair, d_row_chan, d_nsta, d_nlta)));
  QMB_STEP(QMB_TRY(qmb_onsets_combine(ctx, O, d_pair, n_total, d_row_ptr, min_onset_value, d_lin)));
  if (onsets_out)
    QMB_STEP(QMB_CUDA(cudaMemcpyAsync(onsets_out, d_lin, (size_t)n_rows * n_total * sizeof(double),
                                      cudaMemcpyDeviceToHost, ctx->stream)));
#undef QMB_STEP
  cleanup();
  if (rc != QMB_OK) {
    qmb_onsets_free(O);
    return rc;
  }
  *out = O;
  return QMB_OK;
}

extern "C" int qmb_onsets_from_waveforms(qmb_ctx_t ctx, const double *waveforms, int32_t n_channels,
                                         int64_t n_total, int32_t n_rows, const int32_t *row_ptr,
                                         const int32_t *row_chan, const int32_t *nsta, const int32_t *nlta,
                                         int32_t position, int32_t transform, double min_onset_value,
                                         double *onsets_out, qmb_onsets_t *out) {
  if (!waveforms) QMB_FAIL(QMB_ERR_ARG, "null argument");
  return onsets_from_traces(ctx, waveforms, nullptr, n_channels, n_total, n_rows, row_ptr, row_chan, nsta, nlta,
                            position, transform, min_onset_value, onsets_out, out);
}

extern "C" int qmb_onsets_from_device_waveforms(qmb_ctx_t ctx, qmb_waveforms_t w, int32_t n_rows,
                                                const int32_t *row_ptr, const int32_t *row_chan,
                                                const int32_t *nsta, const int32_t *nlta, int32_t position,
                                                int32_t transform, double min_onset_value,
                                                double *onsets_out, qmb_onsets_t *out) {
  if (!ctx || !w) QMB_FAIL(QMB_ERR_ARG, "null argument");
  if (w->ctx != ctx) QMB_FAIL(QMB_ERR_ARG, "handles belong to a different context");
  return onsets_from_traces(ctx, nullptr, w->data.p, w->C, w->n, n_rows, row_ptr, row_chan, nsta, nlta,
                            position, transform, min_onset_value, onsets_out, out);
}

extern "C" int qmb_sta_lta(qmb_ctx_t ctx, const double *signal, int32_t n_channels, int64_t n, int32_t nsta,
                           int32_t nlta, int32_t position, double *onset) {
  if (!ctx || !signal || !onset) QMB_FAIL(QMB_ERR_ARG, "null argument");
  if (n_channels <= 0 || n <= 0) return QMB_OK;
  QMB_CUDA(cudaSetDevice(ctx->device));
  double *d_in = nullptr, *d_out = nullptr;
  const size_t bytes = (size_t)n_channels * n * sizeof(double);
  QMB_CUDA(cudaMalloc((void **)&d_in, bytes));
  cudaError_t e = cudaMalloc((void **)&d_out, bytes);
  if (e != cudaSuccess) {
    cudaFree(d_in);
    QMB_FAIL(QMB_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
  }
  int rc = QMB_OK;
  auto step = [&](cudaError_t err) {
    if (rc == QMB_OK && err != cudaSuccess) {
      qmb_set_error("qmb_sta_lta: %s", cudaGetErrorString(err));
      rc = QMB_ERR_CUDA;
    }
  };
  step(cudaMemcpyAsync(d_in, signal, bytes, cudaMemcpyHostToDevice, ctx->stream));
  step(cudaMemcpyAsync(d_out, onset, bytes, cudaMemcpyHostToDevice, ctx->stream));  // caller's pre-fill
  if (rc == QMB_OK) rc = qmb_stalta_device(ctx, d_in, n_channels, n, n, nsta, nlta, position, QMB_TRANSFORM_NONE, d_out);
  if (rc == QMB_OK) step(cudaMemcpyAsync(onset, d_out, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  step(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_in);
  cudaFree(d_out);
  return rc;
}

// ---------------------------------------------------------------------------------------------
// detect
// ---------------------------------------------------------------------------------------------
extern "C" int qmb_detect_device(qmb_ctx_t ctx, qmb_lut_t lut, qmb_onsets_t onsets, int32_t fsmp, int32_t lsmp,
                                 int32_t available, float *kernel_ms) {
  if (!ctx || !lut || !onsets) QMB_FAIL(QMB_ERR_ARG, "null argument");
  if (lut->ctx != ctx || onsets->ctx != ctx) QMB_FAIL(QMB_ERR_ARG, "handles belong to a different context");
  QMB_CUDA(cudaSetDevice(ctx->device));
  return qmb_scan_run(ctx, lut, onsets, fsmp, lsmp, available, kernel_ms);
}

extern "C" int qmb_detect_fetch(qmb_ctx_t ctx, double *max_coa, double *max_norm_coa, int64_t *max_coa_idx,
                                int32_t n_samples) {
  if (!ctx || !max_coa || !max_norm_coa || !max_coa_idx) QMB_FAIL(QMB_ERR_ARG, "null argument");
  if (ctx->last_n_samples <= 0) QMB_FAIL(QMB_ERR_STATE, "no detect result to fetch");
  if (n_samples != ctx->last_n_samples)
    QMB_FAIL(QMB_ERR_ARG, "n_samples %d does not match the last scan (%d)", n_samples, ctx->last_n_samples);
  QMB_CUDA(cudaSetDevice(ctx->device));
  const size_t n = (size_t)n_samples;
  QMB_CUDA(cudaMemcpyAsync(max_coa, ctx->r_max.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  QMB_CUDA(cudaMemcpyAsync(max_norm_coa, ctx->r_norm.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  QMB_CUDA(cudaMemcpyAsync(max_coa_idx, ctx->r_idx.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  QMB_CUDA(cudaStreamSynchronize(ctx->stream));
  return QMB_OK;
}

extern "C" int qmb_detect(qmb_ctx_t ctx, qmb_lut_t lut, const double *onsets, int32_t n_rows, int64_t n_total,
                          int32_t fsmp, int32_t lsmp, int32_t available, double *max_coa, double *max_norm_coa,
                          int64_t *max_coa_idx) {
  if (!ctx || !lut) QMB_FAIL(QMB_ERR_ARG, "null argument");
  if (n_rows != lut->S)
    QMB_FAIL(QMB_ERR_ARG, "Mismatch between number of stations for data and LUT, %d:%d", n_rows, lut->S);
  if (!onsets || n_total <= 0) QMB_FAIL(QMB_ERR_ARG, "empty onset array");
  QMB_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->tmp_onsets) {
    ctx->tmp_onsets = new qmb_onsets();
    ctx->tmp_onsets->ctx = ctx;
    ctx->tmp_onsets->device = ctx->device;
  }
  QMB_TRY(onsets_fill(ctx, ctx->tmp_onsets, onsets, n_rows, n_total));
  QMB_TRY(qmb_detect_device(ctx, lut, ctx->tmp_onsets, fsmp, lsmp, available, nullptr));
  return qmb_detect_fetch(ctx, max_coa, max_norm_coa, max_coa_idx, (int32_t)(n_total - fsmp - lsmp));
}

// ---------------------------------------------------------------------------------------------
// locate window
// ---------------------------------------------------------------------------------------------
extern "C" int qmb_locate_window(qmb_ctx_t ctx, qmb_lut_t lut, qmb_onsets_t onsets, int32_t fsmp, int32_t lsmp,
                                 int32_t available, int32_t win0, int32_t win1, double *coa_map,
                                 double *map4d_win) {
  if (!ctx || !lut || !onsets) QMB_FAIL(QMB_ERR_ARG, "null argument");
  if (!coa_map && !map4d_win) QMB_FAIL(QMB_ERR_ARG, "nothing requested");
  if (lut->ctx != ctx || onsets->ctx != ctx) QMB_FAIL(QMB_ERR_ARG, "handles belong to a different context");
  QMB_CUDA(cudaSetDevice(ctx->device));
  const size_t n_local = (size_t)lut->n_local;
  const size_t W = (size_t)(win1 > win0 ? win1 - win0 : 0);
  double *d_coa = nullptr, *d_map = nullptr;
  int rc = QMB_OK;
  // the window slab is produced in node-slabs of bounded size and streamed out
  QMB_CUDA(cudaMalloc((void **)&d_coa, n_local * sizeof(double)));
  if (map4d_win) {
    cudaError_t e = cudaMalloc((void **)&d_map, n_local * W * sizeof(double));
    if (e != cudaSuccess) {
      cudaFree(d_coa);
      QMB_FAIL(QMB_ERR_CUDA, "cudaMalloc(%zu bytes) for the window slab: %s", n_local * W * sizeof(double),
               cudaGetErrorString(e));
    }
  }
  rc = qmb_window_run(ctx, lut, onsets, fsmp, lsmp, available, win0, win1, d_coa, d_map, 0);
  auto step = [&](cudaError_t err) {
    if (rc == QMB_OK && err != cudaSuccess) {
      qmb_set_error("qmb_locate_window: %s", cudaGetErrorString(err));
      rc = QMB_ERR_CUDA;
    }
  };
  if (rc == QMB_OK && coa_map)
    step(cudaMemcpyAsync(coa_map, d_coa, n_local * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (rc == QMB_OK && map4d_win)
    step(cudaMemcpyAsync(map4d_win, d_map, n_local * W * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  step(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_coa);
  if (d_map) cudaFree(d_map);
  return rc;
}

// the marginal map stays on the device (locate 3-D stage works on the handle)
extern "C" int qmb_locate_window_device(qmb_ctx_t ctx, qmb_lut_t lut, qmb_onsets_t onsets, int32_t fsmp,
                                        int32_t lsmp, int32_t available, int32_t win0, int32_t win1,
                                        qmb_map3d_t *coa_map) {
  if (!ctx || !lut || !onsets || !coa_map) QMB_FAIL(QMB_ERR_ARG, "null argument");
  *coa_map = nullptr;
  if (lut->ctx != ctx || onsets->ctx != ctx) QMB_FAIL(QMB_ERR_ARG, "handles belong to a different context");
  QMB_CUDA(cudaSetDevice(ctx->device));
  if (lut->y_first || lut->y_last)
    QMB_FAIL(QMB_ERR_UNSUPPORTED, "a device map needs a slab of whole x-planes (this table is a ragged node range)");
  qmb_map3d *m = nullptr;
  QMB_TRY(qmb_map3d_new(ctx, lut->nxl, lut->ny, lut->nz, &m));
  const int rc = qmb_window_run(ctx, lut, onsets, fsmp, lsmp, available, win0, win1, m->v.p, nullptr, 0);
  if (rc != QMB_OK) {
    qmb_map3d_free(m);
    return rc;
  }
  *coa_map = m;
  return QMB_OK;
}
