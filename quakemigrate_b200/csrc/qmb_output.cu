// qmb_output.cu -- the detect OUTPUT stage.
//
//  * k_output_stage: the per-sample results of the fused scan, converted on the device into the
//    integer channels ScanmSEED.append stores (quakemigrate/io/scanmseed.py:102-130):
//    clip at 21474 (:103-104), round(x * 1e5) -> int32 (_data2int :226-240, numpy round = half to
//    even), the C-order unravel of the arg-max node (quakemigrate/lut/lut.py:165-166), the grid
//    coordinate ll_corner + ijk * node_spacing (lut.py:177) and the Z channel
//    round(z * 1e3 * ucf) (:127).  Longitude / latitude need the inverse map projection (pyproj in
//    the reference, lut.py:236): host work, outside this library.
//  * STEIM-1 / STEIM-2 codec (host code): the data section of the miniSEED records a .scanmseed
//    file consists of.  The reference delegates to obspy -> libmseed
//    (st.write(format="MSEED", encoding="STEIM2"), scanmseed.py:216-224); libmseed is absent
//    from /root/reference, so its published packing rule is restated here: per 32-bit word take
//    the largest group of upcoming differences that fits (7x4, 6x5, 5x6, 4x8, 3x10, 2x15, 1x30 bits;
//    STEIM-1: 4x8, 2x16, 1x32), word 0 of each 64-byte frame holds the sixteen 2-bit codes, words
//    1 and 2 of a record's first frame hold the first and last sample.  Pinned byte-for-byte by the
//    reference's own golden file examples/benchmarks/Icequake_Iceland/2014_180.scanmseed
//    (tests/test_scanmseed.py).
#include <string.h>

#include "qmb_internal.cuh"

namespace {

struct OutParams {
  const double *max_coa, *max_norm;
  const long long *idx;
  int n;
  long long ny, nz;
  double ll[3], sp[3], zf;
  int *coa_q, *coan_q, *ijk, *z_q;
  double *xyz;
};

__device__ __forceinline__ int quantise(double v, double factor) {
  // np.round(data * factor).astype(np.int32): one multiply, round half to even
  return __double2int_rn(__dmul_rn(v, factor));
}

__global__ void __launch_bounds__(256) k_output_stage(OutParams p) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.n) return;
  if (p.coa_q) p.coa_q[t] = quantise(fmin(p.max_coa[t], 21474.0), 1e5);
  if (p.coan_q) p.coan_q[t] = quantise(fmin(p.max_norm[t], 21474.0), 1e5);
  const long long f = p.idx[t];
  const long long k = f % p.nz, j = (f / p.nz) % p.ny, i = f / (p.nz * p.ny);
  const long long c[3] = {i, j, k};
  double g[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) g[a] = __dadd_rn(p.ll[a], __dmul_rn((double)c[a], p.sp[a]));  // no FMA: lut.py:177
  if (p.ijk) {
    p.ijk[3 * t + 0] = (int)i;
    p.ijk[3 * t + 1] = (int)j;
    p.ijk[3 * t + 2] = (int)k;
  }
  if (p.xyz) {
    p.xyz[3 * (size_t)t + 0] = g[0];
    p.xyz[3 * (size_t)t + 1] = g[1];
    p.xyz[3 * (size_t)t + 2] = g[2];
  }
  if (p.z_q) p.z_q[t] = quantise(g[2], p.zf);
}

}  // namespace

extern "C" int qmb_detect_fetch_quantised(qmb_ctx_t ctx, const int32_t *node_count, const double *ll_corner,
                                          const double *node_spacing, double z_factor, int32_t *coa_q,
                                          int32_t *coa_n_q, int32_t *ijk, double *xyz, int32_t *z_q,
                                          int32_t n_samples) {
  if (!ctx || !node_count || !ll_corner || !node_spacing) QMB_FAIL(QMB_ERR_ARG, "null argument");
  if (ctx->last_n_samples <= 0) QMB_FAIL(QMB_ERR_STATE, "no detect result to fetch");
  if (n_samples != ctx->last_n_samples)
    QMB_FAIL(QMB_ERR_ARG, "n_samples %d does not match the last scan (%d)", n_samples, ctx->last_n_samples);
  if (node_count[0] <= 0 || node_count[1] <= 0 || node_count[2] <= 0)
    QMB_FAIL(QMB_ERR_ARG, "node_count must be positive");
  QMB_CUDA(cudaSetDevice(ctx->device));
  const size_t n = (size_t)n_samples;
  // one int32 block [coa | coa_n | z | ijk(3n)] and one float64 block [xyz(3n)]
  QMB_TRY(ctx->q_i32.ensure(6 * n));
  QMB_TRY(ctx->q_f64.ensure(3 * n));
  OutParams p = {};
  p.max_coa = ctx->r_max.p; p.max_norm = ctx->r_norm.p; p.idx = ctx->r_idx.p; p.n = n_samples;
  p.ny = node_count[1]; p.nz = node_count[2];
  for (int a = 0; a < 3; ++a) { p.ll[a] = ll_corner[a]; p.sp[a] = node_spacing[a]; }
  p.zf = z_factor;
  p.coa_q = coa_q ? ctx->q_i32.p : nullptr;
  p.coan_q = coa_n_q ? ctx->q_i32.p + n : nullptr;
  p.z_q = z_q ? ctx->q_i32.p + 2 * n : nullptr;
  p.ijk = ijk ? ctx->q_i32.p + 3 * n : nullptr;
  p.xyz = xyz ? ctx->q_f64.p : nullptr;
  k_output_stage<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(p);
  ctx->launches++;
  QMB_CUDA(cudaGetLastError());
  if (coa_q) QMB_CUDA(cudaMemcpyAsync(coa_q, p.coa_q, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (coa_n_q) QMB_CUDA(cudaMemcpyAsync(coa_n_q, p.coan_q, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (z_q) QMB_CUDA(cudaMemcpyAsync(z_q, p.z_q, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (ijk) QMB_CUDA(cudaMemcpyAsync(ijk, p.ijk, 3 * n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (xyz) QMB_CUDA(cudaMemcpyAsync(xyz, p.xyz, 3 * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  QMB_CUDA(cudaStreamSynchronize(ctx->stream));
  return QMB_OK;
}

// ---------------------------------------------------------------------------------------------
// STEIM-1 / STEIM-2 (host)
// ---------------------------------------------------------------------------------------------
namespace {

inline void put_be32(uint8_t *dst, uint32_t v) {
  dst[0] = (uint8_t)(v >> 24); dst[1] = (uint8_t)(v >> 16); dst[2] = (uint8_t)(v >> 8); dst[3] = (uint8_t)v;
}
inline uint32_t get_be32(const uint8_t *src) {
  return ((uint32_t)src[0] << 24) | ((uint32_t)src[1] << 16) | ((uint32_t)src[2] << 8) | (uint32_t)src[3];
}
inline int64_t sign_extend(uint32_t v, int bits) {
  const uint32_t m = 1u << (bits - 1);
  v &= (bits == 32) ? 0xFFFFFFFFu : ((1u << bits) - 1u);
  return (int64_t)(v ^ m) - (int64_t)m;
}

// smallest STEIM-2 field (bits) holding d; 0 = does not fit 30 bits
inline int width2(int64_t d) {
  if (d >= -8 && d <= 7) return 4;
  if (d >= -16 && d <= 15) return 5;
  if (d >= -32 && d <= 31) return 6;
  if (d >= -128 && d <= 127) return 8;
  if (d >= -512 && d <= 511) return 10;
  if (d >= -16384 && d <= 16383) return 15;
  if (d >= -536870912LL && d <= 536870911LL) return 30;
  return 0;
}
inline int width1(int64_t d) {
  if (d >= -128 && d <= 127) return 8;
  if (d >= -32768 && d <= 32767) return 16;
  if (d >= -2147483648LL && d <= 2147483647LL) return 32;
  return 0;
}

struct Group { int count, bits, code, dnib; };  // dnib < 0: no 2-bit sub-code
constexpr Group GROUPS2[7] = {{7, 4, 3, 2}, {6, 5, 3, 1}, {5, 6, 3, 0}, {4, 8, 1, -1},
                              {3, 10, 2, 3}, {2, 15, 2, 2}, {1, 30, 2, 1}};
constexpr Group GROUPS1[3] = {{4, 8, 1, -1}, {2, 16, 2, -1}, {1, 32, 3, -1}};

}  // namespace

extern "C" int qmb_steim_encode(const int32_t *samples, int64_t n, int32_t d0, int32_t version,
                                int32_t n_frames, uint8_t *out, int64_t *n_packed) {
  if (!samples || !out || !n_packed) QMB_FAIL(QMB_ERR_ARG, "null argument");
  if (version != 1 && version != 2) QMB_FAIL(QMB_ERR_ARG, "STEIM version must be 1 or 2");
  if (n <= 0 || n_frames <= 0) QMB_FAIL(QMB_ERR_ARG, "nothing to encode");
  const Group *groups = version == 2 ? GROUPS2 : GROUPS1;
  const int n_groups = version == 2 ? 7 : 3;
  const int look = groups[0].count;
  memset(out, 0, (size_t)n_frames * 64);
  int64_t i = 0;
  for (int f = 0; f < n_frames && i < n; ++f) {
    uint8_t *frame = out + (size_t)f * 64;
    uint32_t codes = 0;
    for (int w = (f == 0 ? 3 : 1); w < 16 && i < n; ++w) {
      int64_t d[7];
      int bits[7];
      const int avail = (int)((n - i) < look ? (n - i) : look);
      for (int k = 0; k < avail; ++k) {
        const int64_t j = i + k;
        d[k] = j == 0 ? (int64_t)d0 : (int64_t)samples[j] - (int64_t)samples[j - 1];
        bits[k] = version == 2 ? width2(d[k]) : width1(d[k]);
      }
      const Group *g = nullptr;
      for (int q = 0; q < n_groups && !g; ++q) {
        if (avail < groups[q].count) continue;
        bool ok = true;
        for (int k = 0; k < groups[q].count; ++k) ok = ok && bits[k] != 0 && bits[k] <= groups[q].bits;
        if (ok) g = &groups[q];
      }
      if (!g)
        QMB_FAIL(QMB_ERR_UNSUPPORTED, "difference %lld at sample %lld does not fit STEIM-%d", (long long)d[0],
                 (long long)i, version);
      uint32_t word = 0;
      for (int k = 0; k < g->count; ++k) {
        const uint32_t field = g->bits == 32 ? (uint32_t)d[k] : ((uint32_t)d[k] & ((1u << g->bits) - 1u));
        word = g->bits == 32 ? field : ((word << g->bits) | field);
      }
      if (g->dnib >= 0) word |= (uint32_t)g->dnib << 30;
      put_be32(frame + 4 * w, word);
      codes |= (uint32_t)g->code << (30 - 2 * w);
      i += g->count;
    }
    put_be32(frame, codes);
  }
  put_be32(out + 4, (uint32_t)samples[0]);      // forward integration constant
  put_be32(out + 8, (uint32_t)samples[i - 1]);  // reverse integration constant
  *n_packed = i;
  return QMB_OK;
}

extern "C" int qmb_steim_decode(const uint8_t *frames, int32_t n_frames, int32_t version, int64_t n_samples,
                                int32_t *out, int32_t *first_diff) {
  if (!frames || !out) QMB_FAIL(QMB_ERR_ARG, "null argument");
  if (version != 1 && version != 2) QMB_FAIL(QMB_ERR_ARG, "STEIM version must be 1 or 2");
  if (n_samples <= 0) return QMB_OK;
  if (n_frames <= 0) QMB_FAIL(QMB_ERR_ARG, "no frames");
  const int64_t x0 = sign_extend(get_be32(frames + 4), 32), xn = sign_extend(get_be32(frames + 8), 32);
  int64_t i = 0, cur = x0;
  for (int f = 0; f < n_frames && i < n_samples; ++f) {
    const uint8_t *frame = frames + (size_t)f * 64;
    const uint32_t codes = get_be32(frame);
    for (int w = (f == 0 ? 3 : 1); w < 16 && i < n_samples; ++w) {
      const int code = (codes >> (30 - 2 * w)) & 3;
      if (code == 0) continue;
      const uint32_t word = get_be32(frame + 4 * w);
      int count, bits;
      if (version == 1) {
        count = code == 1 ? 4 : code == 2 ? 2 : 1;
        bits = 32 / count;
      } else if (code == 1) {
        count = 4; bits = 8;
      } else {
        const int dnib = word >> 30;
        if (code == 2) {
          if (dnib == 0) QMB_FAIL(QMB_ERR_ARG, "invalid STEIM-2 sub-code (frame %d, word %d)", f, w);
          count = dnib; bits = dnib == 1 ? 30 : dnib == 2 ? 15 : 10;
        } else {
          if (dnib == 3) QMB_FAIL(QMB_ERR_ARG, "invalid STEIM-2 sub-code (frame %d, word %d)", f, w);
          count = 5 + dnib; bits = dnib == 0 ? 6 : dnib == 1 ? 5 : 4;
        }
      }
      for (int k = 0; k < count && i < n_samples; ++k, ++i) {
        const int64_t d = sign_extend(bits == 32 ? word : (word >> (bits * (count - 1 - k))), bits);
        if (i == 0) {
          if (first_diff) *first_diff = (int32_t)d;  // x0 is stored explicitly; d0 links to the previous record
        } else {
          cur += d;
        }
        out[i] = (int32_t)cur;
      }
    }
  }
  if (i != n_samples) QMB_FAIL(QMB_ERR_ARG, "record holds %lld of %lld samples", (long long)i, (long long)n_samples);
  if (cur != xn)
    QMB_FAIL(QMB_ERR_ARG, "reverse integration constant mismatch (%lld != %lld)", (long long)cur, (long long)xn);
  return QMB_OK;
}
