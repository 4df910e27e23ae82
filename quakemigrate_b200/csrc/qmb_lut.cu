// qmb_lut.cu -- device-resident travel-time tables, re-tiled for the scan kernel.
//
// Input is exactly the array the reference hands to migrate(): int32 (n_nodes, n_rows), station
// axis fastest (quakemigrate/core/src/migratelib.c:53; produced by LUT.serve_traveltimes,
// quakemigrate/lut/lut.py:502-538).  The scan kernel wants, per spatially compact TILE of nodes
// (4x4x4 or 4x4x8) and per onset row:
//     tmin   = (min over the tile of max(0,tau)) rounded down to a multiple of 8 samples
//     tau16  = max(0,tau) - tmin    (uint16; the tile's tau-range is small because it is compact)
//     span   = max over the tile of tau16
// so that one 16-byte-aligned bulk copy of  [tmin, tmin + span + time_tile)  of the onset row
// (float32: 4 samples per 16 bytes; 16-bit fixed point: 8) serves every node of the tile.
//     tauq16 = word offset of the node's sample PAIR inside the staged 16-bit row (see
//              qmb_internal.cuh): tau16 >> 1, plus RSq / 2 for odd tau16 ("odd" stream).  The clamp of negative travel times (migratelib.c:55) is
// applied here, once.
#include <stdlib.h>
#include <string.h>

#include "qmb_internal.cuh"

namespace {

constexpr int LT_THREADS = 128;
constexpr int LT_ROWS = 32;  // rows handled per pass through shared memory

struct LutTileParams {
  const int32_t *raw;  // (nxl*ny*nz, S), or nullptr when serving from float64 grids:
  const double *tt;    //   seconds, [grid][node of the FULL grid]
  const int *rowsel;   //   grid index of onset row s
  double sr;           //   samples per second: tau = rint(tt * sr)  (lut.py:538)
  long long tt_stride, tt_node0;  // nodes per grid; flat index of this slab's first node
  int nxl, ny, nz, S, Spad;
  int NB;
  const int *boxes;    // [n_tiles][9]: origin x0,y0,z0; tile shape sx,sy,sz; region end x1,y1,z1
  uint32_t *slotmap;   // [n_local]: tile * NB + slot of every node
  long long node_offset, n_valid;
  long long first;     // flat index (inside the covering box) of the slab's first node: a slab may start
                       // and end in the middle of an x-plane (whole z-columns), see qmb_lut_upload_range
  uint16_t *tau16;
  int32_t *tmin;
  uint16_t *span;
  uint32_t *nodeidx;
  int *stats;  // [0] max tau, [1] max span, [2] overflow flag
  int NPW;     // nodes per consumer warp (for the duplicate statistic)
  unsigned long long *dup;  // [0] same-offset neighbours inside a warp's node group, [1] pairs examined
};

__global__ void __launch_bounds__(LT_THREADS) k_lut_tile(LutTileParams p) {
  __shared__ int s_tau[LT_ROWS][128 + 1];
  __shared__ long long s_node[128];  // clamped local flat index per slot
  const int j = blockIdx.x;
  const int *bx = p.boxes + (size_t)j * 9;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NWARPS = LT_THREADS / 32;

  if (tid < p.NB) {
    const int sx = bx[3], sy = bx[4], sz = bx[5];
    (void)sx;
    const int dz = tid % sz, dy = (tid / sz) % sy, dx = tid / (sz * sy);
    const int ix = bx[0] + dx, iy = bx[1] + dy, iz = bx[2] + dz;
    // slots outside the tile's region (or the grid) are padding; they replicate the nearest
    // in-region node so that they do not widen the tile's travel-time range
    const int cx = min(ix, bx[6] - 1), cy = min(iy, bx[7] - 1), cz = min(iz, bx[8] - 1);
    long long local = ((long long)cx * p.ny + cy) * p.nz + cz - p.first;
    const long long flat = ((long long)ix * p.ny + iy) * p.nz + iz - p.first;
    const bool valid = ix < bx[6] && iy < bx[7] && iz < bx[8] && flat >= 0 && flat < p.n_valid;
    if (local >= p.n_valid) local = p.n_valid - 1;  // ragged last x-plane of a pseudo-grid
    if (local < 0) local = 0;
    s_node[tid] = local;
    p.nodeidx[(size_t)j * p.NB + tid] = valid ? (uint32_t)(p.node_offset + flat) : 0xFFFFFFFFu;
    if (valid) p.slotmap[flat] = (uint32_t)j * (uint32_t)p.NB + (uint32_t)tid;
  }
  __syncthreads();

  int my_max_tau = 0, my_max_span = 0, overflow = 0;
  unsigned int my_dup = 0, my_pairs = 0;
  for (int s0 = 0; s0 < p.Spad; s0 += LT_ROWS) {
    // gather: one warp per node slot, lanes over rows (coalesced 128-byte reads of the raw table)
    for (int q = warp; q < p.NB; q += NWARPS) {
      const int s = s0 + lane;
      int v = 0;
      if (s < p.S) {
        if (p.raw) v = p.raw[s_node[q] * p.S + s];
        else v = (int)rint(p.tt[(size_t)p.rowsel[s] * p.tt_stride + p.tt_node0 + s_node[q]] * p.sr);
        v = max(0, v);
      }
      s_tau[lane][q] = v;
    }
    __syncthreads();
    // reduce + emit: one warp per row
    for (int r = warp; r < LT_ROWS; r += NWARPS) {
      const int s = s0 + r;
      if (s >= p.Spad) break;
      int mn = 0x7fffffff, mx = 0;
      for (int q = lane; q < p.NB; q += 32) {
        mn = min(mn, s_tau[r][q]);
        mx = max(mx, s_tau[r][q]);
      }
      for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      }
      const int base = mn & ~7;  // 16-byte aligned window starts for the float32 AND the 16-bit rows
      const int sp = mx - base;
      if (sp > 65535) overflow = 1;
      for (int q = lane; q < p.NB; q += 32) {
        p.tau16[((size_t)j * p.Spad + s) * p.NB + q] = (uint16_t)min(s_tau[r][q] - base, 65535);
        if (s < p.S && (q % p.NPW) != 0) {  // neighbour inside the same consumer warp's node group
          ++my_pairs;
          if (s_tau[r][q] == s_tau[r][q - 1]) ++my_dup;
        }
      }
      if (lane == 0) {
        p.tmin[(size_t)j * p.Spad + s] = base;
        p.span[(size_t)j * p.Spad + s] = (uint16_t)min(sp, 65535);
      }
      my_max_tau = max(my_max_tau, mx);
      my_max_span = max(my_max_span, sp);
    }
    __syncthreads();
  }
  for (int o = 16; o > 0; o >>= 1) {
    my_dup += __shfl_xor_sync(0xffffffffu, my_dup, o);
    my_pairs += __shfl_xor_sync(0xffffffffu, my_pairs, o);
  }
  if (lane == 0) {
    atomicMax(&p.stats[0], my_max_tau);
    atomicMax(&p.stats[1], my_max_span);
    if (overflow) atomicMax(&p.stats[2], 1);
    atomicAdd(&p.dup[0], (unsigned long long)my_dup);
    atomicAdd(&p.dup[1], (unsigned long long)my_pairs);
  }
}

__global__ void __launch_bounds__(256) k_lut_tauq(const uint16_t *tau16, uint16_t *tauq16, size_t n, int half_words) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned t = tau16[i];
  tauq16[i] = (uint16_t)((t >> 1) + (t & 1u) * (unsigned)half_words);
}

// byte offsets for the packed scan: [tile][row group][slot][G]; value = position of the row inside
// its pipeline stage + the pair's word inside the row (the odd stream sits RSq * 2 bytes further)
__global__ void __launch_bounds__(256) k_lut_tauq32(const uint16_t *tau16, uint32_t *tauq32, uint16_t *tauq16p, size_t n,
                                                    int Spad, int NB, int CR, int G, int RSq) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // index into tau16: (tile*Spad + s)*NB + slot
  if (i >= n) return;
  const unsigned t = tau16[i];
  const int slot = (int)(i % NB);
  const size_t js = i / NB;
  const int s = (int)(js % Spad);
  const size_t j = js / Spad;
  const uint32_t off = (uint32_t)(s % CR) * (uint32_t)RSq * 4u + 4u * (t >> 1) + (t & 1u) * (uint32_t)RSq * 2u;
  const size_t o = ((j * (size_t)(Spad / G) + (size_t)(s / G)) * NB + slot) * G + (s % G);
  tauq32[o] = off;
  if (tauq16p) tauq16p[o] = (uint16_t)off;
}

}  // namespace

// Elements per staged 16-bit stream for this table: the widest window [tmin, tmin + span + TT + 1]
// rounded up to 16 bytes, plus 8 elements so that consecutive rows do not start on the same bank.
int qmb_lut_rsq(const qmb_lut *L) {
  const int TT = 32 * L->v.K;
  return ((L->max_span + TT + 2 + 7) & ~7) + 8;
}

// ---------------------------------------------------------------------------------------------
// host: cover the slab with node tiles.  The bulk uses the variant's preferred shape; what does
// not divide evenly is covered by up to three remainder slabs, each tiled with the factorisation
// of NB that wastes the fewest slots (ties: most compact), so padding stays ~0.2 % instead of the
// ~6 % a single padded shape costs on e.g. a 101x101x51 grid.
// ---------------------------------------------------------------------------------------------
namespace {

struct Region {
  int x0, y0, z0, x1, y1, z1;
};

// The widest tile decides how much shared memory a pipeline stage takes for the whole table (its
// travel-time range sets the staged row length), so a shape is also kept compact: no edge longer
// than 8 nodes unless that costs more than a quarter more slots than the best elongated shape
// (a single z-column can only be cut into 1x1xNB pieces).
void best_shape(int NB, int bx, int by, int bz, int &sx, int &sy, int &sz) {
  constexpr int LONGEST = 8;
  long long best_vol[2] = {-1, -1};  // [0] any shape, [1] compact shapes
  int best_sum[2] = {0, 0}, shape[2][3] = {{1, 1, NB}, {1, 1, NB}};
  for (int a = 1; a <= NB; ++a) {
    if (NB % a) continue;
    for (int b = 1; b <= NB / a; ++b) {
      if ((NB / a) % b) continue;
      const int c = NB / a / b;
      const long long vol = (long long)((bx + a - 1) / a) * a * ((by + b - 1) / b) * b * ((bz + c - 1) / c) * c;
      const int sum = a + b + c;
      const bool compact = a <= LONGEST && b <= LONGEST && c <= LONGEST;
      for (int k = 0; k < (compact ? 2 : 1); ++k)
        if (best_vol[k] < 0 || vol < best_vol[k] || (vol == best_vol[k] && sum < best_sum[k])) {
          best_vol[k] = vol; best_sum[k] = sum;
          shape[k][0] = a; shape[k][1] = b; shape[k][2] = c;
        }
    }
  }
  const int k = (best_vol[1] >= 0 && 4 * best_vol[1] <= 5 * best_vol[0]) ? 1 : 0;
  sx = shape[k][0]; sy = shape[k][1]; sz = shape[k][2];
}

void tile_region(const Region &r, int sx, int sy, int sz, std::vector<int> &boxes) {
  for (int x = r.x0; x < r.x1; x += sx)
    for (int y = r.y0; y < r.y1; y += sy)
      for (int z = r.z0; z < r.z1; z += sz) {
        const int b[9] = {x, y, z, sx, sy, sz, r.x1, r.y1, r.z1};
        boxes.insert(boxes.end(), b, b + 9);
      }
}

}  // namespace

static int lut_build_variant(qmb_ctx *ctx, qmb_lut *L, const int32_t *d_raw, const QmbServeSource *src, int variant) {
  L->v = QMB_VARIANTS[variant];
  L->tx = L->v.tx;
  L->ty = L->v.ty;
  L->tz = L->v.tz;
  L->NB = L->tx * L->ty * L->tz;
  std::vector<int> boxes;
  // bulk tiles of the variant's shape + up to three remainder slabs per box; a slab that starts /
  // ends inside an x-plane is three boxes: the partial first plane, the whole planes, the partial
  // last plane
  auto tile_box = [&](const Region &b) {
    const int ex = b.x1 - b.x0, ey = b.y1 - b.y0, ez = b.z1 - b.z0;
    if (ex <= 0 || ey <= 0 || ez <= 0) return;
    const int Mx = (ex / L->tx) * L->tx, My = (ey / L->ty) * L->ty, Mz = (ez / L->tz) * L->tz;
    int sx, sy, sz;
    if (Mx > 0 && My > 0 && Mz > 0)
      tile_region({b.x0, b.y0, b.z0, b.x0 + Mx, b.y0 + My, b.z0 + Mz}, L->tx, L->ty, L->tz, boxes);
    const Region rem[3] = {
        {b.x0 + Mx, b.y0, b.z0, b.x1, b.y1, b.z1},                            // x remainder (everything if Mx == 0)
        {b.x0, b.y0 + My, b.z0, b.x0 + Mx, b.y1, b.z1},                       // y remainder
        {b.x0, b.y0, b.z0 + (My > 0 ? Mz : 0), b.x0 + Mx, b.y0 + My, b.z1},   // z remainder
    };
    for (const Region &r : rem) {
      if (r.x1 <= r.x0 || r.y1 <= r.y0 || r.z1 <= r.z0) continue;
      best_shape(L->NB, r.x1 - r.x0, r.y1 - r.y0, r.z1 - r.z0, sx, sy, sz);
      tile_region(r, sx, sy, sz, boxes);
    }
  };
  {
    int xs = 0, xe = L->nxl;
    if (L->nxl == 1 && (L->y_first > 0 || L->y_last > 0)) {
      tile_box({0, L->y_first, 0, 1, L->y_last > 0 ? L->y_last : L->ny, L->nz});
      xs = xe = 0;
    } else {
      if (L->y_first > 0) {
        tile_box({0, L->y_first, 0, 1, L->ny, L->nz});
        xs = 1;
      }
      if (L->y_last > 0) {
        tile_box({L->nxl - 1, 0, 0, L->nxl, L->y_last, L->nz});
        xe = L->nxl - 1;
      }
    }
    tile_box({xs, 0, 0, xe, L->ny, L->nz});
  }
  const long long nt = (long long)(boxes.size() / 9);
  if (nt > 0x7fffffffLL / 128) QMB_FAIL(QMB_ERR_UNSUPPORTED, "too many node tiles (%lld)", nt);
  L->n_tiles = (int)nt;
  QMB_TRY(L->boxes.ensure(boxes.size()));
  QMB_TRY(L->slotmap.ensure((size_t)L->n_local));
  QMB_CUDA(cudaMemcpyAsync(L->boxes.p, boxes.data(), boxes.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  L->chunk_rows = L->v.CR;
  L->Spad = ((L->S + L->chunk_rows - 1) / L->chunk_rows) * L->chunk_rows;

  QMB_TRY(L->tau16.ensure((size_t)L->n_tiles * L->Spad * L->NB));
  QMB_TRY(L->tmin.ensure((size_t)L->n_tiles * L->Spad));
  QMB_TRY(L->span.ensure((size_t)L->n_tiles * L->Spad));
  QMB_TRY(L->nodeidx.ensure((size_t)L->n_tiles * L->NB));

  int *d_stats = nullptr;  // 4 ints followed (8-byte aligned) by 2 64-bit counters
  QMB_CUDA(cudaMalloc((void **)&d_stats, 8 * sizeof(int)));
  QMB_CUDA(cudaMemsetAsync(d_stats, 0, 8 * sizeof(int), ctx->stream));

  LutTileParams p = {};
  p.raw = d_raw;
  if (src) {
    p.tt = src->tt; p.rowsel = src->rowsel; p.sr = src->sr; p.tt_stride = src->stride; p.tt_node0 = src->node0;
  }
  p.nxl = L->nxl; p.ny = L->ny; p.nz = L->nz; p.S = L->S; p.Spad = L->Spad;
  p.NB = L->NB;
  p.boxes = L->boxes.p;
  p.slotmap = L->slotmap.p;
  p.node_offset = L->node_offset;
  p.n_valid = L->n_local;
  p.first = (long long)L->y_first * L->nz;
  p.tau16 = L->tau16.p; p.tmin = L->tmin.p; p.span = L->span.p; p.nodeidx = L->nodeidx.p;
  p.stats = d_stats;
  p.NPW = L->v.NPW;
  p.dup = (unsigned long long *)(d_stats + 4);
  k_lut_tile<<<L->n_tiles, LT_THREADS, 0, ctx->stream>>>(p);
  ctx->launches++;
  QMB_CUDA(cudaGetLastError());
  int h_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  QMB_CUDA(cudaMemcpyAsync(h_stats, d_stats, 8 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  QMB_CUDA(cudaStreamSynchronize(ctx->stream));  // also: `boxes` (host vector) was read by the H2D copy
  cudaFree(d_stats);
  L->max_tau = h_stats[0];
  L->max_span = h_stats[1];
  {
    unsigned long long d[2];
    memcpy(d, h_stats + 4, sizeof(d));
    L->dup_frac = d[1] ? (double)d[0] / (double)d[1] : 0.0;
  }
  if (h_stats[2])
    QMB_FAIL(QMB_ERR_UNSUPPORTED,
             "travel-time range inside one %dx%dx%d node tile exceeds 65535 samples", L->tx, L->ty,
             L->tz);
  // offsets for the 16-bit scan (needs max_span); skipped when they would not fit 16 bits
  L->RSq = qmb_lut_rsq(L);
  if ((L->v.K & 1) == 0 && L->max_span / 2 + L->RSq / 2 < 65536) {
    const size_t n = (size_t)L->n_tiles * L->Spad * L->NB;
    QMB_TRY(L->tauq16.ensure(n));
    k_lut_tauq<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(L->tau16.p, L->tauq16.p, n, L->RSq / 2);
    ctx->launches++;
    QMB_CUDA(cudaGetLastError());
    L->qG = 1;
    L->tauq32.release();
    if (ctx->scan_qgroup > 1 && (L->chunk_rows % ctx->scan_qgroup) == 0 && L->v.NPW == 4) {
      L->qG = ctx->scan_qgroup;
      QMB_TRY(L->tauq32.ensure(n));
      // 16-bit form for groups of two rows when every offset of a stage fits
      L->tauq16p.release();
      const long long max_off = (long long)(L->chunk_rows - 1) * L->RSq * 4 + 4LL * (L->max_span / 2) + 2LL * L->RSq;
      if (L->qG == 2 && max_off < 65536) QMB_TRY(L->tauq16p.ensure(n));
      k_lut_tauq32<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(L->tau16.p, L->tauq32.p, L->tauq16p.p, n,
                                                                        L->Spad, L->NB, L->chunk_rows, L->qG, L->RSq);
      ctx->launches++;
      QMB_CUDA(cudaGetLastError());
    }
  } else {
    L->RSq = 0;
    L->tauq16.release();
    L->tauq32.release();
    L->tauq16p.release();
    L->qG = 1;
  }
  return QMB_OK;
}

// Pipeline stages the scan kernel would get for this table (mirrors scan_setup in qmb_scan.cu).
int qmb_lut_stage_count(const qmb_lut *L) {
  const int TT = 32 * L->v.K;
  const int RS = ((L->max_span + TT + 3) & ~3) + 4;
  int stage_bytes = L->chunk_rows * RS * 4 + L->chunk_rows * L->NB * 2 + L->NB * 4;
  stage_bytes = (stage_bytes + 127) & ~127;
  return (QMB_SMEM_BUDGET - 1024) / (stage_bytes + 16);
}

// Tile with the context's variant; if the travel-time range inside a tile is so wide that fewer
// than three pipeline stages fit in shared memory, fall back to variants with fewer rows per stage
// (4) and then smaller node tiles (6) before giving up.
int qmb_lut_build(qmb_ctx *ctx, qmb_lut *L, const int32_t *d_raw, const QmbServeSource *src) {
  const int order[3] = {ctx->variant, 4, 6};
  int rc = QMB_OK;
  for (int k = 0; k < 3; ++k) {
    if (k > 0 && order[k] == order[k - 1]) continue;
    rc = lut_build_variant(ctx, L, d_raw, src, order[k]);
    if (rc != QMB_OK) return rc;
    const int stages = qmb_lut_stage_count(L);
    if (stages >= 3 || (k == 2 && stages >= 2)) return QMB_OK;
  }
  if (qmb_lut_stage_count(L) >= 2) return QMB_OK;
  QMB_FAIL(QMB_ERR_UNSUPPORTED,
           "travel-time range inside one node tile (%d samples) is too wide for the shared-memory pipeline",
           L->max_span);
}

// ---------------------------------------------------------------------------------------------
// LUT.decimate on the device (quakemigrate/lut/lut.py:102-140): keep every df-th node per axis,
// starting c1 nodes in so that the kept nodes are centred; one strided gather per grid.
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) k_grid_decimate(const double *in, double *out, int nx, int ny, int nz, int mx,
                                                       int my, int mz, int dx, int dy, int dz, int cx, int cy, int cz) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (long long)mx * my * mz;
  if (m >= per) return;
  const int g = blockIdx.y;
  const int k = (int)(m % mz), j = (int)((m / mz) % my), i = (int)(m / ((long long)mz * my));
  const long long src = ((long long)(cx + i * dx) * ny + (cy + j * dy)) * nz + (cz + k * dz);
  out[(size_t)g * per + m] = in[(size_t)g * nx * ny * nz + src];
}
}  // namespace

extern "C" int qmb_ttgrids_decimate(qmb_ctx_t ctx, qmb_ttgrids_t G, const int32_t *df, qmb_ttgrids_t *out) {
  if (!ctx || !G || !df || !out) QMB_FAIL(QMB_ERR_ARG, "null argument");
  *out = nullptr;
  if (G->ctx != ctx) QMB_FAIL(QMB_ERR_ARG, "grids belong to a different context");
  if (df[0] < 1 || df[1] < 1 || df[2] < 1) QMB_FAIL(QMB_ERR_ARG, "decimation factors must be >= 1");
  QMB_CUDA(cudaSetDevice(ctx->device));
  const int n[3] = {G->nx, G->ny, G->nz};
  int m[3], c[3];
  for (int a = 0; a < 3; ++a) {
    m[a] = 1 + (n[a] - 1) / df[a];                         // lut.py:123
    c[a] = (n[a] - df[a] * (m[a] - 1) - 1) / 2;            // lut.py:124
  }
  qmb_ttgrids *D = new qmb_ttgrids();
  D->ctx = ctx; D->device = ctx->device; D->n_grids = G->n_grids; D->nx = m[0]; D->ny = m[1]; D->nz = m[2];
  const long long per = (long long)m[0] * m[1] * m[2];
  int rc = D->tt.ensure((size_t)G->n_grids * per);
  if (rc != QMB_OK) {
    delete D;
    return rc;
  }
  dim3 grid((unsigned)((per + 255) / 256), (unsigned)G->n_grids);
  k_grid_decimate<<<grid, 256, 0, ctx->stream>>>(G->tt.p, D->tt.p, n[0], n[1], n[2], m[0], m[1], m[2], df[0], df[1], df[2],
                                                 c[0], c[1], c[2]);
  ctx->launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    D->tt.release();
    delete D;
    QMB_FAIL(QMB_ERR_CUDA, "qmb_ttgrids_decimate: %s", cudaGetErrorString(e));
  }
  *out = D;
  return QMB_OK;
}

extern "C" int qmb_ttgrids_shape(qmb_ttgrids_t G, int32_t *shape4) {
  if (!G || !shape4) QMB_FAIL(QMB_ERR_ARG, "null argument");
  shape4[0] = G->n_grids; shape4[1] = G->nx; shape4[2] = G->ny; shape4[3] = G->nz;
  return QMB_OK;
}
