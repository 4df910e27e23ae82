// qmb_scan.cu -- the hot path: fused migrate + stack + per-sample max / arg-max / node-sum.
//
// Replaces, for detect(), the pair  migrate() + find_max_coa()
// (quakemigrate/core/src/migratelib.c:40-65 and :85-111, called from QuakeScan._compute,
// quakemigrate/signal/scan.py:635-638) WITHOUT materialising the (n_nodes, n_samples) volume.
//
// Work decomposition
//   item      = (time tile of 32*K samples) x (contiguous range of node tiles)
//   CTA       = persistent, one per SM; walks items  blockIdx.x, +gridDim.x, ...
//   node tile = W*NPW spatially compact nodes (default 12*4 = 48: 4x4x3 boxes), laid out by qmb_lut.cu
//   stage     = (node tile, chunk of CR onset rows): the producer warp bulk-copies, per row, the
//               window [tmin, tmin + span + 32K) of the log-onset row (float32, or the two 16-bit streams) into shared memory
//               (cp.async.bulk -> SASS UBLKCP, completion on an mbarrier) plus the tile's uint16
//               tau offsets; a ring of stages decouples it from the W consumer warps.
//   consumer  = warp w owns NPW nodes of the tile; LANES ARE TIME SAMPLES (lane l handles
//               samples 32k + l), so every shared-memory gather of a shifted onset row is a
//               contiguous, conflict-free 128-byte wavefront whatever the travel times are, and
//               the per-sample reduction over nodes lives in registers (no shuffles, no atomics):
//               each lane keeps, for its K samples, the running top-3 values (with the node
//               indices of the top two) and the float64 sum of exp(stack / available).
//
// MODE 4 (default for detect): the rows are staged as 16-bit fixed point, two samples per 32-bit
// word, so one LDS.32 feeds two gather-adds and the stacks are exact integer sums of the quantised
// rows; delta below becomes S * qstep / 2 (+ float32 conversion).  MODE 0 / 1 keep float32 rows.
//
// Exactness of the arg-max.  Stacks are accumulated in float32 in station order.  With
//   delta = n_rows * 2^-24 * sum_rows max|log onset|   (rigorous bound on |fp32 stack - exact|)
// a node whose fp32 stack is more than 2*delta below the fp32 maximum cannot be the float64
// arg-max.  Hence: if m1 - m2 > 2 delta the fp32 winner is THE winner; else if m1 - m3 > 2 delta
// the two leading candidates are re-evaluated in float64 with the reference's own summation
// order and compared as exp(stack/available) with lowest-index tie-break (migratelib.c:96-105);
// otherwise (3+ candidates, e.g. exact ties on clipped data) that sample is re-scanned over all
// nodes in float64 (slow path, rare).  The reported max_coa is always exp() of the float64
// re-evaluated stack of the winner.
#include <math.h>

#include <stdlib.h>

#include <algorithm>

#include "qmb_internal.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + bulk async copy (TMA engine, 1-D form)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "QMB_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra QMB_DONE_%=;\n"
      "bra QMB_WAIT_%=;\n"
      "QMB_DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------------------------------------
// scan kernel
// ---------------------------------------------------------------------------------------------
struct ScanParams {
  // onsets
  const float *on32;     // [S][row_stride]
  // MODE 4 (16-bit fixed point): [S][2][row_stride] uint16, see qmb_onsets in qmb_internal.cuh;
  // stack = qc0 + qstep * sum(u)
  const uint16_t *on16;
  float qstep, qc0;
  // 1 and 65536 as kernel parameters: multiplying by them keeps the word / upper-half accumulation
  // on the fma pipe (IMAD, IMAD.HI) -- the compiler cannot fold a multiplier it does not know --
  // while shifts and adds would all land on the alu pipe, which is the one that saturates
  uint32_t q_one, q_64k;
  int q_form;  // 0: IADD3 + LEA.HI (alu)  1: IMAD word sums  2: 1 + IMAD.HI for odd words  3: IMAD / IADD3 alternating
  long long row_stride;  // elements
  long long t_origin;    // index within a row of (tau = 0, sample t' = 0); multiple of 4
  int S, n_chunks;       // n_chunks = stages per node tile (rows / CR)
  // LUT tiles
  const uint16_t *tau16;
  const int32_t *tmin;
  const uint16_t *span;
  const uint32_t *nodeidx;
  int n_tiles, Spad;
  // tiling
  int RS;            // shared row stride (floats), multiple of 4
  int n_stages;
  int stage_bytes;
  int n_time_tiles, R;   // items = n_time_tiles x R node ranges (ranges balanced to +-1 tile)
  float scale;       // log2(e) / available
  // per-(item range, warp) partial results, indexed [(r*16 + w) * Tpad + t']
  float *pm1, *pm2, *pm3;
  uint32_t *pi1, *pi2;
  double *psum;
  long long Tpad;
  // MODE 1 (marginal window / materialise): outputs indexed by LOCAL node
  double *map;          // nullable [n_local][T]; accumulated into when map_accumulate
  double *coa_map;      // nullable [n_local]: sum over the window of exp(stack/available)
  int map_accumulate;
  int shift;            // t' = t + shift
  long long T;          // output samples
  long long node_offset;
  double avail;
};

template <int W, int NPW, int K, int MODE, int CR>
__global__ void __launch_bounds__((W + 1) * 32, 1) k_scan(const ScanParams p) {
  constexpr int NB = W * NPW;
  constexpr int TT = 32 * K;
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // stage layout: [16 rows x RS floats][16 x NB uint16 tau][NB uint32 node index]
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t rows_bytes = (uint32_t)CR * p.RS * 4u;
  constexpr uint32_t tau_bytes = (uint32_t)CR * NB * 2u;
  const uint32_t bar_base = smem_base + (uint32_t)p.n_stages * p.stage_bytes;  // full[], then empty[]

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(bar_base + 8u * s, 1);
      mbar_init(bar_base + 8u * (p.n_stages + s), W);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int n_items = p.n_time_tiles * p.R;
  int st = 0;
  uint32_t ph = 0;

  if (warp == W) {
    // ===================================== producer ==========================================
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int r = it / p.n_time_tiles, tt = it - r * p.n_time_tiles;
      const int tile0 = (int)((long long)r * p.n_tiles / p.R);
      const int tile1 = (int)((long long)(r + 1) * p.n_tiles / p.R);
      const long long tbase = p.t_origin + (long long)tt * TT;
      for (int j = tile0; j < tile1; ++j) {
        for (int c = 0; c < p.n_chunks; ++c) {
          const int s = c * CR + lane;
          const bool act = lane < CR && s < p.S;
          // metadata loads are issued before the wait so their latency overlaps it
          int tmin = 0, len = 0;
          if (act) {
            tmin = p.tmin[(size_t)j * p.Spad + s];
            if constexpr (MODE >= 4) len = ((int)p.span[(size_t)j * p.Spad + s] + TT + 2 + 7) & ~7;
            else len = ((int)p.span[(size_t)j * p.Spad + s] + TT + 3) & ~3;
          }
          // MODE 4: `bytes` per stream (2 bytes per sample), two streams per row
          uint32_t bytes = MODE >= 4 ? (uint32_t)len * 2u : (uint32_t)len * 4u;
          uint32_t total = MODE >= 4 ? 2u * bytes : bytes;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
          total += tau_bytes + (c == 0 ? (uint32_t)NB * 4u : 0u);
          const uint32_t full = bar_base + 8u * st, empty = bar_base + 8u * (p.n_stages + st);
          mbar_wait(empty, ph ^ 1u);
          const uint32_t stage = smem_base + (uint32_t)st * p.stage_bytes;
          if (lane == 0) mbar_expect_tx(full, total);
          __syncwarp();
          if constexpr (MODE >= 4) {
            if (act) {
              const uint16_t *src = p.on16 + (size_t)2 * s * p.row_stride + tbase + tmin;
              const uint32_t dst = stage + (uint32_t)lane * p.RS * 4u;
              bulk_g2s(dst, src, bytes, full);                                       // pairs (2i, 2i+1)
              bulk_g2s(dst + (uint32_t)p.RS * 2u, src + p.row_stride, bytes, full);  // pairs (2i+1, 2i+2)
            }
          } else if (act) {
            bulk_g2s(stage + (uint32_t)lane * p.RS * 4u, p.on32 + (size_t)s * p.row_stride + tbase + tmin,
                     bytes, full);
          }
          if (lane == 0)
            bulk_g2s(stage + rows_bytes, p.tau16 + ((size_t)j * p.Spad + (size_t)c * CR) * NB,
                     tau_bytes, full);
          if (lane == 1 && c == 0)
            bulk_g2s(stage + rows_bytes + tau_bytes, p.nodeidx + (size_t)j * NB, NB * 4u, full);
          if (++st == p.n_stages) { st = 0; ph ^= 1u; }
        }
      }
    }
    return;
  }

  // ======================================= consumers ===========================================
  const float NEG_INF = __int_as_float(0xff800000);
  for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
    const int r = it / p.n_time_tiles, tt = it - r * p.n_time_tiles;
    const int tile0 = (int)((long long)r * p.n_tiles / p.R);
    const int tile1 = (int)((long long)(r + 1) * p.n_tiles / p.R);

    float m1[K], m2[K], m3[K];
    uint32_t i1[K], i2[K];
    double sum[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      m1[k] = m2[k] = m3[k] = NEG_INF;
      i1[k] = i2[k] = 0xFFFFFFFFu;
      sum[k] = 0.0;
    }

    if constexpr (MODE >= 4) {
      constexpr int QF = MODE - 4;
      // ---- 16-bit fixed-point rows: lane l owns the sample PAIRS (64 kk + 2 l, + 1), kk < K/2.
      // One LDS.32 fetches both samples of a pair (half the shared-memory bytes per gather-add);
      // A accumulates the whole words, Hs their upper halves: sum(lo) = A - (Hs << 16) exactly
      // (mod 2^32, and sum(lo) < 2^32), sum(hi) = Hs.
      static_assert(NPW == 4 && (K % 2) == 0, "16-bit scan is built for 4 nodes per warp");
      constexpr int KW = K / 2, H = NPW / 2;
      for (int j = tile0; j < tile1; ++j) {
        uint32_t A[NPW][KW], Hs[NPW][KW];
#pragma unroll
        for (int q = 0; q < NPW; ++q)
#pragma unroll
          for (int k = 0; k < KW; ++k) A[q][k] = Hs[q][k] = 0u;
        uint32_t nidx[NPW];
        for (int c = 0; c < p.n_chunks; ++c) {
          const uint32_t full = bar_base + 8u * st, empty = bar_base + 8u * (p.n_stages + st);
          const uint32_t stage = smem_base + (uint32_t)st * p.stage_bytes;
          mbar_wait(full, ph);
          if (c == 0) {
            const uint32_t ia = stage + rows_bytes + tau_bytes + (uint32_t)(warp * NPW) * 4u;
#pragma unroll
            for (int q = 0; q < NPW; ++q)
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(nidx[q]) : "r"(ia + 4u * q));
          }
          const int nrows = min(CR, p.S - c * CR);
          const uint32_t tau_addr = stage + rows_bytes + (uint32_t)(warp * NPW) * 2u;
          uint32_t row_addr = stage + (uint32_t)lane * 4u;
          auto load_off = [&](int sl, uint32_t (&off)[NPW]) {
            uint32_t a, b;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(tau_addr + (uint32_t)sl * NB * 2u));
            off[0] = a & 0xffffu; off[1] = a >> 16; off[2] = b & 0xffffu; off[3] = b >> 16;
          };
          auto load_half = [&](int h, const uint32_t (&off)[NPW], uint32_t base, uint32_t (&v)[H][KW]) {
#pragma unroll
            for (int q = 0; q < H; ++q) {
              const uint32_t a = base + off[h * H + q] * 4u;
#pragma unroll
              for (int k = 0; k < KW; ++k)
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[q][k]) : "r"(a + 128u * k));
            }
          };
          auto add_half = [&](int h, const uint32_t (&v)[H][KW]) {
#pragma unroll
            for (int q = 0; q < H; ++q)
#pragma unroll
              for (int k = 0; k < KW; ++k) {
                if (QF == 0 || (QF == 3 && (k & 1))) A[h * H + q][k] += v[q][k];
                else A[h * H + q][k] = v[q][k] * p.q_one + A[h * H + q][k];
                if (QF == 2 && (k & 1)) Hs[h * H + q][k] = __umulhi(v[q][k], p.q_64k) + Hs[h * H + q][k];
                else Hs[h * H + q][k] += v[q][k] >> 16;
              }
          };
          uint32_t off[NPW], off_next[NPW];
          uint32_t v0[H][KW], v1[H][KW];
          load_off(0, off);
          load_half(0, off, row_addr, v0);
#pragma unroll 2
          for (int sl = 0; sl + 1 < nrows; ++sl) {
            load_half(1, off, row_addr, v1);
            load_off(sl + 1, off_next);
            add_half(0, v0);
            row_addr += (uint32_t)p.RS * 4u;
            load_half(0, off_next, row_addr, v0);
            add_half(1, v1);
#pragma unroll
            for (int q = 0; q < NPW; ++q) off[q] = off_next[q];
          }
          load_half(1, off, row_addr, v1);
          add_half(0, v0);
          add_half(1, v1);
          __syncwarp();
          if (lane == 0) mbar_arrive(empty);
          if (++st == p.n_stages) { st = 0; ph ^= 1u; }
        }
        // ---- per-tile epilogue: slot k = 2 kk + h <-> sample 64 kk + 2 lane + h ---------------
#pragma unroll
        for (int kk = 0; kk < KW; ++kk) {
          float e0 = 0.f, e1 = 0.f;
#pragma unroll
          for (int q = 0; q < NPW; ++q) {
            const bool pad = nidx[q] == 0xFFFFFFFFu;
            const uint32_t hi = Hs[q][kk], lo = A[q][kk] - (hi << 16);
            const float vv[2] = {pad ? NEG_INF : fmaf((float)lo, p.qstep, p.qc0),
                                 pad ? NEG_INF : fmaf((float)hi, p.qstep, p.qc0)};
            e0 += ex2_approx(vv[0] * p.scale);
            e1 += ex2_approx(vv[1] * p.scale);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int k = 2 * kk + h;
              const float v = vv[h];
              const bool g1 = v > m1[k];
              const bool g2 = v > m2[k];
              m3[k] = g2 ? m2[k] : fmaxf(m3[k], v);
              m2[k] = g1 ? m1[k] : (g2 ? v : m2[k]);
              i2[k] = g1 ? i1[k] : (g2 ? nidx[q] : i2[k]);
              m1[k] = g1 ? v : m1[k];
              i1[k] = g1 ? nidx[q] : i1[k];
            }
          }
          sum[2 * kk] += (double)e0;
          sum[2 * kk + 1] += (double)e1;
        }
      }
      // ---- item done: publish (pairs of neighbouring samples -> 8 / 16-byte stores) -----------
      const size_t base = ((size_t)r * W + warp) * p.Tpad + (size_t)tt * TT + 2u * lane;
#pragma unroll
      for (int kk = 0; kk < KW; ++kk) {
        const size_t o = base + 64u * kk;
        *reinterpret_cast<float2 *>(p.pm1 + o) = make_float2(m1[2 * kk], m1[2 * kk + 1]);
        *reinterpret_cast<float2 *>(p.pm2 + o) = make_float2(m2[2 * kk], m2[2 * kk + 1]);
        *reinterpret_cast<float2 *>(p.pm3 + o) = make_float2(m3[2 * kk], m3[2 * kk + 1]);
        *reinterpret_cast<uint2 *>(p.pi1 + o) = make_uint2(i1[2 * kk], i1[2 * kk + 1]);
        *reinterpret_cast<uint2 *>(p.pi2 + o) = make_uint2(i2[2 * kk], i2[2 * kk + 1]);
        *reinterpret_cast<double2 *>(p.psum + o) = make_double2(sum[2 * kk], sum[2 * kk + 1]);
      }
      continue;
    }

    for (int j = tile0; j < tile1; ++j) {
      float acc[NPW][K];
#pragma unroll
      for (int q = 0; q < NPW; ++q)
#pragma unroll
        for (int k = 0; k < K; ++k) acc[q][k] = 0.f;
      uint32_t nidx[NPW];

      for (int c = 0; c < p.n_chunks; ++c) {
        const uint32_t full = bar_base + 8u * st, empty = bar_base + 8u * (p.n_stages + st);
        const uint32_t stage = smem_base + (uint32_t)st * p.stage_bytes;
        mbar_wait(full, ph);
        if (c == 0) {
          const uint32_t ia = stage + rows_bytes + tau_bytes + (uint32_t)(warp * NPW) * 4u;
#pragma unroll
          for (int q = 0; q < NPW; ++q)
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(nidx[q]) : "r"(ia + 4u * q));
        }
        const int nrows = min(CR, p.S - c * CR);
        const uint32_t tau_addr = stage + rows_bytes + (uint32_t)(warp * NPW) * 2u;
        uint32_t row_addr = stage + (uint32_t)lane * 4u;
        // Software-pipelined over rows with a half-row skew: the gathers of one half of the
        // warp's nodes are always in flight while the other half is being accumulated, so at
        // least NPW*K/2 independent LDS are outstanding whatever order ptxas picks.
        constexpr int H = NPW / 2;
        auto load_off = [&](int sl, uint32_t (&off)[NPW]) {
          if constexpr (MODE == 3) {  // ablation: no offset fetch (timing only, results meaningless)
#pragma unroll
            for (int q = 0; q < NPW; ++q) off[q] = 4u * q + (uint32_t)(sl & 3);
          } else if constexpr (NPW == 4) {
            uint32_t a, b;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(tau_addr + (uint32_t)sl * NB * 2u));
            off[0] = a & 0xffffu; off[1] = a >> 16; off[2] = b & 0xffffu; off[3] = b >> 16;
          } else {
            uint32_t a, b, cc, d;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(a), "=r"(b), "=r"(cc), "=r"(d)
                         : "r"(tau_addr + (uint32_t)sl * NB * 2u));
            off[0] = a & 0xffffu; off[1] = a >> 16; off[2] = b & 0xffffu; off[3] = b >> 16;
            off[4] = cc & 0xffffu; off[5] = cc >> 16; off[6] = d & 0xffffu; off[7] = d >> 16;
          }
        };
        auto load_half = [&](int h, const uint32_t (&off)[NPW], uint32_t base, float (&v)[H][K]) {
#pragma unroll
          for (int q = 0; q < H; ++q) {
            const uint32_t a = base + off[h * H + q] * 4u;
#pragma unroll
            for (int k = 0; k < K; ++k) v[q][k] = lds_f32(a + 128u * k);
          }
        };
        auto add_half = [&](int h, const float (&v)[H][K]) {
#pragma unroll
          for (int q = 0; q < H; ++q)
#pragma unroll
            for (int k = 0; k < K; ++k) acc[h * H + q][k] += v[q][k];
        };
        uint32_t off[NPW], off_next[NPW];
        float v0[H][K], v1[H][K];
        load_off(0, off);
        load_half(0, off, row_addr, v0);
#pragma unroll 2
        for (int sl = 0; sl + 1 < nrows; ++sl) {
          load_half(1, off, row_addr, v1);
          load_off(sl + 1, off_next);
          add_half(0, v0);
          row_addr += (uint32_t)p.RS * 4u;
          load_half(0, off_next, row_addr, v0);
          add_half(1, v1);
#pragma unroll
          for (int q = 0; q < NPW; ++q) off[q] = off_next[q];
        }
        load_half(1, off, row_addr, v1);
        add_half(0, v0);
        add_half(1, v1);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty);
        if (++st == p.n_stages) { st = 0; ph ^= 1u; }
      }

      if constexpr (MODE == 1) {
        // ---- materialise / marginalise: exp(stack/available) per (node, sample) in float64 ----
#pragma unroll
        for (int q = 0; q < NPW; ++q) {
          if (nidx[q] == 0xFFFFFFFFu) continue;  // warp-uniform
          const long long local = (long long)nidx[q] - p.node_offset;
          double nodesum = 0.0;
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const long long t = (long long)tt * TT + 32 * k + lane - p.shift;
            if (t >= 0 && t < p.T) {
              double x = (double)acc[q][k];
              if (p.map && p.map_accumulate) x += p.map[(size_t)local * p.T + t];
              const double c = exp(x / p.avail);
              if (p.map) p.map[(size_t)local * p.T + t] = c;
              nodesum += c;
            }
          }
          if (p.coa_map) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) nodesum += __shfl_xor_sync(0xffffffffu, nodesum, o);
            if (lane == 0) atomicAdd(&p.coa_map[local], nodesum);
          }
        }
        continue;
      }
      if constexpr (MODE == 2) {  // ablation: no epilogue (timing only): keep the stacks alive cheaply
#pragma unroll
        for (int k = 0; k < K; ++k) {
          float t = acc[0][k];
#pragma unroll
          for (int q = 1; q < NPW; ++q) t += acc[q][k];
          m1[k] = fmaxf(m1[k], t);
        }
        continue;
      }
      // ---- per-tile epilogue: fold NPW nodes into the running per-sample reduction ----------
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float esum = 0.f;
#pragma unroll
        for (int q = 0; q < NPW; ++q) {
          const bool pad = nidx[q] == 0xFFFFFFFFu;
          const float v = pad ? NEG_INF : acc[q][k];
          esum += ex2_approx(v * p.scale);  // exp(stack / available); exp2(-inf) = 0
          const bool g1 = v > m1[k];
          const bool g2 = v > m2[k];
          m3[k] = g2 ? m2[k] : fmaxf(m3[k], v);
          m2[k] = g1 ? m1[k] : (g2 ? v : m2[k]);
          i2[k] = g1 ? i1[k] : (g2 ? nidx[q] : i2[k]);
          m1[k] = g1 ? v : m1[k];
          i1[k] = g1 ? nidx[q] : i1[k];
        }
        sum[k] += (double)esum;
      }
    }

    if constexpr (MODE == 1) continue;
    // ---- item done: publish this warp's partial reduction -------------------------------------
    const size_t base = ((size_t)r * W + warp) * p.Tpad + (size_t)tt * TT + lane;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const size_t o = base + 32u * k;
      p.pm1[o] = m1[k]; p.pm2[o] = m2[k]; p.pm3[o] = m3[k];
      p.pi1[o] = i1[k]; p.pi2[o] = i2[k];
      p.psum[o] = sum[k];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// top-3 bookkeeping shared by the merge kernels
// ---------------------------------------------------------------------------------------------
struct Top3 {
  float m1, m2, m3;
  uint32_t i1, i2;
};
__device__ __forceinline__ void top3_insert(Top3 &a, float v, uint32_t i) {
  const bool g1 = v > a.m1, g2 = v > a.m2;
  a.m3 = g2 ? a.m2 : fmaxf(a.m3, v);
  a.m2 = g1 ? a.m1 : (g2 ? v : a.m2);
  a.i2 = g1 ? a.i1 : (g2 ? i : a.i2);
  a.m1 = g1 ? v : a.m1;
  a.i1 = g1 ? i : a.i1;
}

// merge `n_parts` partial reductions per sample: part q lives at [q * stride + t]
__global__ void __launch_bounds__(256) k_merge(const float *pm1, const float *pm2, const float *pm3,
                                               const uint32_t *pi1, const uint32_t *pi2,
                                               const double *psum, int n_parts, long long stride,
                                               long long stride_sum, long long n, float *gm1, float *gm2, float *gm3,
                                               uint32_t *gi1, uint32_t *gi2, double *gsum) {
  // blockIdx.y selects a group of parts; group g writes its fold at offset g * n (two-stage merge)
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int per = (n_parts + gridDim.y - 1) / gridDim.y;
  const int q0 = blockIdx.y * per, q1 = min(n_parts, q0 + per);
  Top3 a;
  a.m1 = a.m2 = a.m3 = __int_as_float(0xff800000);
  a.i1 = a.i2 = 0xFFFFFFFFu;
  double s = 0.0;
  for (int q = q0; q < q1; ++q) {
    const size_t o = (size_t)q * stride + t;
    top3_insert(a, pm1[o], pi1[o]);
    top3_insert(a, pm2[o], pi2[o]);
    a.m3 = fmaxf(a.m3, pm3[o]);
    s += psum[(size_t)q * stride_sum + t];
  }
  const size_t w = (size_t)blockIdx.y * n + t;
  gm1[w] = a.m1; gm2[w] = a.m2; gm3[w] = a.m3;
  gi1[w] = a.i1; gi2[w] = a.i2;
  gsum[w] = s;
}

// ---------------------------------------------------------------------------------------------
// float64 re-evaluation of arg-max candidates (reference summation order: rows ascending)
// ---------------------------------------------------------------------------------------------
struct ExactParams {
  const double *on64;
  long long row_stride, origin;  // origin = front + fsmp (index of tau = 0, t = 0)
  int S, Spad;
  const uint16_t *tau16;
  const int32_t *tmin;
  const uint32_t *slotmap;  // [n_local] tile * NB + slot
  int NB;
  long long node_offset, n_local;
};

// stack(node, t) in float64; node is a GLOBAL flat index owned by this rank
__device__ double exact_stack(const ExactParams &e, long long gnode, long long t) {
  const uint32_t where = e.slotmap[gnode - e.node_offset];
  const int j = (int)(where / (uint32_t)e.NB), q = (int)(where % (uint32_t)e.NB);
  double acc = 0.0;
  for (int s = 0; s < e.S; ++s) {
    const size_t m = (size_t)j * e.Spad + s;
    const long long tau = (long long)e.tmin[m] + e.tau16[m * e.NB + q];
    acc += e.on64[(size_t)s * e.row_stride + e.origin + tau + t];
  }
  return acc;
}

// ex1/ex2[t'] = float64 stack of candidate 1 / 2 when owned by this rank, else 0 (summed across
// ranks afterwards).  Candidate 2 is only evaluated when it can matter.
__global__ void __launch_bounds__(128) k_exact(ExactParams e, const float *gm1, const float *gm2,
                                               const uint32_t *gi1, const uint32_t *gi2, float thr,
                                               int shift, long long n_samples, double *ex1, double *ex2) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // output sample
  if (t >= n_samples) return;
  const long long tp = t + shift;
  double a = 0.0, b = 0.0;
  const long long c1 = gi1[tp], c2 = gi2[tp];
  if (c1 != 0xFFFFFFFFLL && c1 >= e.node_offset && c1 < e.node_offset + e.n_local) a = exact_stack(e, c1, t);
  if (gm1[tp] - gm2[tp] <= thr && c2 != 0xFFFFFFFFLL && c2 >= e.node_offset &&
      c2 < e.node_offset + e.n_local)
    b = exact_stack(e, c2, t);
  ex1[t] = a;
  ex2[t] = b;
}

// decide winner per sample; queue samples that need the all-node float64 re-scan
__global__ void __launch_bounds__(256) k_finalize(const float *gm1, const float *gm2, const float *gm3,
                                                  const uint32_t *gi1, const uint32_t *gi2,
                                                  const double *gsum, const double *ex1, const double *ex2,
                                                  float thr, int shift, long long n_samples, double avail,
                                                  double n_nodes_global, double *r_max, double *r_norm,
                                                  long long *r_idx, int *slow_list, int slow_cap) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_samples) return;
  const long long tp = t + shift;
  const float a1 = gm1[tp], a2 = gm2[tp], a3 = gm3[tp];
  double best = exp(ex1[t] / avail);
  long long idx = gi1[tp];
  if (a1 - a2 <= thr) {
    const double other = exp(ex2[t] / avail);
    const long long oi = gi2[tp];
    if (oi != 0xFFFFFFFFLL && (other > best || (other == best && oi < idx))) {
      best = other;
      idx = oi;
    }
    if (a1 - a3 <= thr) {  // three or more candidates: full float64 re-scan of this sample
      const int slot = atomicAdd(&slow_list[0], 1);
      if (slot < slow_cap) slow_list[1 + slot] = (int)t;
    }
  }
  r_max[t] = best;
  r_idx[t] = idx;
  r_norm[t] = best * n_nodes_global / gsum[tp];
}

// slow path: for each queued sample, float64 coalescence of every local node; per-CTA best
// (value, lowest index) written to partial arrays [cta][slot].
constexpr int SLOW_BATCH = 8;
__global__ void __launch_bounds__(128) k_slow(ExactParams e, const int *slow_list, int n_slow, int batch_offset,
                                              double avail, const uint32_t *nodeidx, int n_tiles,
                                              double *part_val, unsigned long long *part_idx) {
  // grid.x = CTAs striding over tiles, grid.y = batch of SLOW_BATCH queued samples
  const int b0 = (blockIdx.y + batch_offset) * SLOW_BATCH;
  const int nb = min(SLOW_BATCH, n_slow - b0);
  long long ts[SLOW_BATCH];
  for (int b = 0; b < SLOW_BATCH; ++b) ts[b] = (b < nb) ? slow_list[1 + b0 + b] : 0;
  double best[SLOW_BATCH];
  unsigned long long bidx[SLOW_BATCH];
  for (int b = 0; b < SLOW_BATCH; ++b) { best[b] = -1.0; bidx[b] = ~0ull; }
  const int slots_per_cta = blockDim.x;  // one thread per node slot, several tiles per pass
  const int tiles_per_pass = slots_per_cta / e.NB > 0 ? slots_per_cta / e.NB : 1;
  for (int j0 = blockIdx.x * tiles_per_pass; j0 < n_tiles; j0 += gridDim.x * tiles_per_pass) {
    const int j = j0 + threadIdx.x / e.NB, q = threadIdx.x % e.NB;
    if (j >= n_tiles || (int)threadIdx.x >= tiles_per_pass * e.NB) continue;
    const uint32_t gi = nodeidx[(size_t)j * e.NB + q];
    if (gi == 0xFFFFFFFFu) continue;
    double acc[SLOW_BATCH];
    for (int b = 0; b < SLOW_BATCH; ++b) acc[b] = 0.0;
    for (int s = 0; s < e.S; ++s) {
      const size_t m = (size_t)j * e.Spad + s;
      const long long tau = (long long)e.tmin[m] + e.tau16[m * e.NB + q];
      const double *row = e.on64 + (size_t)s * e.row_stride + e.origin + tau;
#pragma unroll
      for (int b = 0; b < SLOW_BATCH; ++b) acc[b] += row[ts[b]];
    }
#pragma unroll
    for (int b = 0; b < SLOW_BATCH; ++b) {
      const double c = exp(acc[b] / avail);
      if (c > best[b] || (c == best[b] && (unsigned long long)gi < bidx[b])) { best[b] = c; bidx[b] = gi; }
    }
  }
  // block reduce (value desc, index asc)
  __shared__ double sv[128];
  __shared__ unsigned long long si[128];
  for (int b = 0; b < nb; ++b) {
    sv[threadIdx.x] = best[b];
    si[threadIdx.x] = bidx[b];
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) {
        const double v = sv[threadIdx.x + o];
        const unsigned long long i = si[threadIdx.x + o];
        if (v > sv[threadIdx.x] || (v == sv[threadIdx.x] && i < si[threadIdx.x])) {
          sv[threadIdx.x] = v;
          si[threadIdx.x] = i;
        }
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      part_val[(size_t)blockIdx.x * n_slow + b0 + b] = sv[0];
      part_idx[(size_t)blockIdx.x * n_slow + b0 + b] = si[0];
    }
    __syncthreads();
  }
}

// fold [n_parts][n_slow] partials; optionally scatter into the result arrays
__global__ void k_slow_merge(const double *part_val, const unsigned long long *part_idx, int n_parts,
                             int n_slow, double *out_val, unsigned long long *out_idx) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_slow) return;
  double best = -1.0;
  unsigned long long bi = ~0ull;
  for (int q = 0; q < n_parts; ++q) {
    const double v = part_val[(size_t)q * n_slow + b];
    const unsigned long long i = part_idx[(size_t)q * n_slow + b];
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
  out_val[b] = best;
  out_idx[b] = bi;
}

__global__ void k_slow_scatter(const int *slow_list, int n_slow, const double *val,
                               const unsigned long long *idx, const double *gsum, int shift,
                               double n_nodes_global, double *r_max, double *r_norm, long long *r_idx) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_slow) return;
  const int t = slow_list[1 + b];
  r_max[t] = val[b];
  r_idx[t] = (long long)idx[b];
  r_norm[t] = val[b] * n_nodes_global / gsum[t + shift];
}

template <int W, int NPW, int K, int MODE, int CR>
int launch_scan(qmb_ctx *ctx, const ScanParams &p, int grid, size_t smem) {
  QMB_CUDA(cudaFuncSetAttribute(k_scan<W, NPW, K, MODE, CR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_scan<W, NPW, K, MODE, CR><<<grid, (W + 1) * 32, smem, ctx->stream>>>(p);
  ctx->launches++;
  QMB_CUDA(cudaGetLastError());
  return QMB_OK;
}

template <int MODE>
int dispatch_scan(qmb_ctx *ctx, const QmbVariant &v, const ScanParams &p, int grid, size_t smem) {
  switch (v.id) {
    case 0: return launch_scan<16, 4, 8, MODE, 16>(ctx, p, grid, smem);
    case 1: return launch_scan<16, 8, 4, MODE, 16>(ctx, p, grid, smem);
    case 2: return launch_scan<8, 8, 8, MODE, 16>(ctx, p, grid, smem);
    case 3: return launch_scan<8, 8, 4, MODE, 16>(ctx, p, grid, smem);
    case 4: return launch_scan<12, 4, 8, MODE, 16>(ctx, p, grid, smem);
    case 5: return launch_scan<12, 8, 4, MODE, 16>(ctx, p, grid, smem);
    case 6: return launch_scan<8, 4, 8, MODE, 16>(ctx, p, grid, smem);
    case 7:
      if constexpr (MODE == 0) {  // QMB_ABLATE=2|3: timing-only ablations of the default kernel
        const char *ab = getenv("QMB_ABLATE");
        if (ab && atoi(ab) == 2) return launch_scan<12, 4, 8, 2, 32>(ctx, p, grid, smem);
        if (ab && atoi(ab) == 3) return launch_scan<12, 4, 8, 3, 32>(ctx, p, grid, smem);
      }
      return launch_scan<12, 4, 8, MODE, 32>(ctx, p, grid, smem);
    case 8: return launch_scan<15, 4, 8, MODE, 32>(ctx, p, grid, smem);
    case 9: return launch_scan<15, 4, 8, MODE, 16>(ctx, p, grid, smem);
  }
  QMB_FAIL(QMB_ERR_ARG, "unknown kernel variant %d", v.id);
}

constexpr int QMB_Q16_FORM = 1;  // default accumulation form (see ScanParams::q_form)
// 16-bit fixed-point detect scan: built for the 4-nodes-per-warp variants
bool q16_variant(const QmbVariant &v) { return v.NPW == 4 && (v.K % 2) == 0; }
int dispatch_scan_q16(qmb_ctx *ctx, const QmbVariant &v, const ScanParams &p, int grid, size_t smem) {
  switch (v.id) {
    case 0: return launch_scan<16, 4, 8, 4 + QMB_Q16_FORM, 16>(ctx, p, grid, smem);
    case 4: return launch_scan<12, 4, 8, 4 + QMB_Q16_FORM, 16>(ctx, p, grid, smem);
    case 6: return launch_scan<8, 4, 8, 4 + QMB_Q16_FORM, 16>(ctx, p, grid, smem);
    case 8: return launch_scan<15, 4, 8, 4 + QMB_Q16_FORM, 32>(ctx, p, grid, smem);
    case 9: return launch_scan<15, 4, 8, 4 + QMB_Q16_FORM, 16>(ctx, p, grid, smem);
    case 7:
      if (p.q_form == 0) return launch_scan<12, 4, 8, 4, 32>(ctx, p, grid, smem);
      if (p.q_form == 1) return launch_scan<12, 4, 8, 5, 32>(ctx, p, grid, smem);
      if (p.q_form == 3) return launch_scan<12, 4, 8, 7, 32>(ctx, p, grid, smem);
      return launch_scan<12, 4, 8, 6, 32>(ctx, p, grid, smem);
  }
  QMB_FAIL(QMB_ERR_ARG, "kernel variant %d has no 16-bit form", v.id);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------
// Tiling decisions shared by the detect scan and the window pass.
static int scan_setup(qmb_ctx *ctx, qmb_lut *L, qmb_onsets *O, long long first, long long T,
                      int32_t available, ScanParams &p, size_t &smem, int &shift_out, bool use_q = false) {
  const int K = L->v.K, TT = 32 * K, NB = L->NB;
  p.on32 = O->log32.p;
  p.row_stride = O->stride;
  long long front = O->front;
  int align = 3;  // time tiles start on 16-byte boundaries of the staged rows
  p.tau16 = L->tau16.p;
  p.RS = ((L->max_span + TT + 3) & ~3) + 4;
  if (use_q) {  // 16-bit fixed-point rows (detect scan only)
    p.on16 = O->q16.p;
    p.row_stride = O->qstride;
    front = O->qfront;
    align = 7;
    p.tau16 = L->tauq16.p;
    p.RS = L->RSq;  // elements per stream; a staged row is two streams = RS * 4 bytes, as for float32
    p.qstep = (float)O->qstep;
    p.qc0 = (float)((double)L->S * O->qmin);
    p.q_one = 1u;
    p.q_64k = 65536u;
    p.q_form = QMB_Q16_FORM;
    if (const char *f = getenv("QMB_Q16_FORM")) p.q_form = atoi(f);
  }
  const int shift = (int)((front + first) & align);  // time tiles start `shift` samples early
  shift_out = shift;
  p.shift = shift;
  p.T = T;
  p.t_origin = front + first - shift;
  p.S = L->S;
  p.n_chunks = L->Spad / L->chunk_rows;
  p.tmin = L->tmin.p; p.span = L->span.p; p.nodeidx = L->nodeidx.p;
  p.n_tiles = L->n_tiles; p.Spad = L->Spad;
  p.stage_bytes = L->chunk_rows * p.RS * 4 + L->chunk_rows * NB * 2 + NB * 4;
  p.stage_bytes = (p.stage_bytes + 127) & ~127;
  const int avail_smem = QMB_SMEM_BUDGET - 1024;
  p.n_stages = avail_smem / (p.stage_bytes + 16);
  if (p.n_stages > 12) p.n_stages = 12;
  if (p.n_stages < 2)
    QMB_FAIL(QMB_ERR_UNSUPPORTED, "travel-time range inside a node tile (%d samples) too large for shared memory", L->max_span);
  smem = (size_t)p.n_stages * p.stage_bytes + (size_t)p.n_stages * 16 + 128;
  p.n_time_tiles = (int)((T + shift + TT - 1) / TT);
  // Items = time tiles x R node ranges (ranges balanced to +-1 tile), dealt round-robin to the
  // persistent CTAs.  All items cost the same, so the tail loss is 1 - items / (rounds * grid):
  // pick the R in [R0, 4 R0] that minimises it (>= 8 tiles per item, partial buffers <= 3 GiB).
  const int grid = ctx->sm_count;
  const int W = L->v.W;
  int R0 = (32 * grid + p.n_time_tiles - 1) / p.n_time_tiles;
  if (R0 < 1) R0 = 1;
  int maxR = L->n_tiles / 8;
  if (maxR < 1) maxR = 1;
  const long long part_bytes_per_R = (long long)W * p.n_time_tiles * TT * 28;
  const int memR = (int)std::max<long long>(1, (3LL << 30) / std::max<long long>(1, part_bytes_per_R));
  if (maxR > memR) maxR = memR;
  if (R0 > maxR) R0 = maxR;
  int R = R0;
  double best_eff = -1.0;
  for (int cand = R0; cand <= std::min(maxR, 4 * R0); ++cand) {
    const long long items = (long long)cand * p.n_time_tiles;
    const double eff = (double)items / (double)(((items + grid - 1) / grid) * grid);
    if (eff > best_eff + 1e-9) { best_eff = eff; R = cand; }
  }
  if (const char *env = getenv("QMB_NODE_RANGES")) {  // tuning override
    const int forced = atoi(env);
    if (forced >= 1 && forced <= std::max(1, L->n_tiles)) R = forced;
  }
  p.R = R;
  p.scale = (float)(1.4426950408889634 / (double)available);
  p.Tpad = (long long)p.n_time_tiles * TT;
  p.node_offset = L->node_offset;
  p.avail = (double)available;
  return QMB_OK;
}

static int check_scan_args(qmb_lut *L, qmb_onsets *O, int32_t fsmp, int32_t lsmp, int32_t available,
                           long long &T) {
  if (L->S != O->S)
    QMB_FAIL(QMB_ERR_ARG, "Mismatch between number of stations for data and LUT, %d:%d", O->S, L->S);
  T = O->P - fsmp - lsmp;
  if (fsmp < 0 || lsmp < 0 || T <= 0) QMB_FAIL(QMB_ERR_ARG, "Data array smaller than coalescence array.");
  if (T > 0x7fffffffLL - 1024) QMB_FAIL(QMB_ERR_UNSUPPORTED, "n_samples too large");
  if (available <= 0) QMB_FAIL(QMB_ERR_ARG, "available must be positive");
  if (L->max_tau > lsmp)
    QMB_FAIL(QMB_ERR_ARG,
             "maximum travel time (%d samples) exceeds the post-pad (%d samples): the reference "
             "reads past the onset rows here (post_pad must be >= max traveltime, scan.py:213)",
             L->max_tau, lsmp);
  return QMB_OK;
}

// Marginal window / materialisation pass: samples [win0, win1) of the 4-D map.
int qmb_window_run(qmb_ctx *ctx, qmb_lut *L, qmb_onsets *O, int32_t fsmp, int32_t lsmp,
                   int32_t available, int32_t win0, int32_t win1, double *d_coa_map,
                   double *d_map4d_win, int accumulate) {
  long long T;
  QMB_TRY(check_scan_args(L, O, fsmp, lsmp, available, T));
  if (win0 < 0 || win1 > T || win0 >= win1) QMB_FAIL(QMB_ERR_ARG, "window [%d, %d) outside [0, %lld)", win0, win1, T);
  ScanParams p = {};
  size_t smem;
  int shift;
  QMB_TRY(scan_setup(ctx, L, O, (long long)fsmp + win0, (long long)win1 - win0, available, p, smem, shift));
  p.map = d_map4d_win;
  p.coa_map = d_coa_map;
  p.map_accumulate = accumulate;
  if (d_coa_map) QMB_CUDA(cudaMemsetAsync(d_coa_map, 0, (size_t)L->n_local * sizeof(double), ctx->stream));
  return dispatch_scan<1>(ctx, L->v, p, ctx->sm_count, smem);
}

int qmb_scan_run(qmb_ctx *ctx, qmb_lut *L, qmb_onsets *O, int32_t fsmp, int32_t lsmp,
                 int32_t available, float *kernel_ms) {
  long long T;
  QMB_TRY(check_scan_args(L, O, fsmp, lsmp, available, T));
  ScanParams p = {};
  size_t smem;
  int shift;
  // detect scans read 16-bit fixed-point rows unless float32 was asked for (or the table / rows
  // have no 16-bit form); candidates are re-evaluated in float64 either way
  bool use_q = ctx->scan_fp32 != 1 && O->qstep > 0.0 && O->q16.p && L->RSq > 0 && L->tauq16.p &&
               q16_variant(L->v) && O->S <= 256;  // sum(u) < 2^24: exact in float32
  if (use_q && ctx->scan_fp32 == 0) {
    // automatic choice: the 16-bit rows leave arg-max and max_coa exact, but max_norm_coa divides
    // by a node sum built from the quantised stacks.  One node's coalescence is off by about
    // sigma = qstep * sqrt(S / 12) / available (relative), and the sum over N nodes averages that
    // down by ~sqrt(N) (less where neighbouring nodes share samples: a factor 32 is held back).
    // Keep float32 rows when that estimate is not clearly inside the 1e-5 bar (tiny grids, or
    // `available` much smaller than the number of rows).
    const double sigma = O->qstep * sqrt((double)O->S / 12.0) / (double)available;
    if (32.0 * sigma / sqrt((double)L->n_global) > 5e-6) use_q = false;
  }
  if (use_q && scan_setup(ctx, L, O, fsmp, T, available, p, smem, shift, true) != QMB_OK) {
    use_q = false;  // e.g. too few pipeline stages for the (slightly wider) 16-bit rows
    p = ScanParams{};
  }
  if (!use_q) QMB_TRY(scan_setup(ctx, L, O, fsmp, T, available, p, smem, shift, false));
  const int R = p.R, grid = ctx->sm_count;

  const size_t n_part = (size_t)R * L->v.W * p.Tpad;
  QMB_TRY(ctx->pm1.ensure(n_part)); QMB_TRY(ctx->pm2.ensure(n_part)); QMB_TRY(ctx->pm3.ensure(n_part));
  QMB_TRY(ctx->pi1.ensure(n_part)); QMB_TRY(ctx->pi2.ensure(n_part)); QMB_TRY(ctx->psum.ensure(n_part));
  p.pm1 = ctx->pm1.p; p.pm2 = ctx->pm2.p; p.pm3 = ctx->pm3.p;
  p.pi1 = ctx->pi1.p; p.pi2 = ctx->pi2.p; p.psum = ctx->psum.p;
  const size_t Tp = (size_t)p.Tpad;
  QMB_TRY(ctx->gm1.ensure(Tp * 33)); QMB_TRY(ctx->gm2.ensure(Tp * 33)); QMB_TRY(ctx->gm3.ensure(Tp * 33));
  QMB_TRY(ctx->gi1.ensure(Tp * 33)); QMB_TRY(ctx->gi2.ensure(Tp * 33)); QMB_TRY(ctx->gsum.ensure(Tp * 33));
  QMB_TRY(ctx->ex1.ensure(Tp)); QMB_TRY(ctx->ex2.ensure(Tp));
  QMB_TRY(ctx->r_max.ensure((size_t)T)); QMB_TRY(ctx->r_norm.ensure((size_t)T)); QMB_TRY(ctx->r_idx.ensure((size_t)T));
  const int slow_cap = (int)T;
  QMB_TRY(ctx->slow_list.ensure((size_t)slow_cap + 1));

  // fp32 error bound (see file header): delta = S * 2^-24 * sum_rows max|v|, threshold 2*delta
  std::vector<double> rowmax(O->S);
  QMB_CUDA(cudaMemcpyAsync(rowmax.data(), O->rowmax.p, O->S * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  QMB_CUDA(cudaStreamSynchronize(ctx->stream));
  double A = 0.0;
  for (double v : rowmax) A += v;
  float thr = (float)(2.0 * 1.001 * (double)O->S * 5.9604644775390625e-08 * A) + 1e-30f;
  if (use_q) {
    // |quantised stack - exact| <= S * qstep / 2 (rounding to the grid) + the float32 evaluation of
    // qc0 + qstep * sum(u): rounding of qstep, of qc0 and of the fma, each <= 2^-24 of
    // S * (|qmin| + 65535 qstep)
    const double S = (double)O->S, mag = S * (fabs(O->qmin) + 65535.0 * O->qstep);
    const double delta = 0.5 * S * O->qstep + 4.0 * 5.9604644775390625e-08 * mag;
    thr = (float)(2.0 * 1.001 * delta) + 1e-30f;
  }
  ctx->last_scan_q16 = use_q ? 1 : 0;

  QMB_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  int rc;
  rc = use_q ? dispatch_scan_q16(ctx, L->v, p, grid, smem) : dispatch_scan<0>(ctx, L->v, p, grid, smem);
  QMB_TRY(rc);
  QMB_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));

  // ---- merge partials (local), then across ranks ----------------------------------------------
  const int mb = (int)((p.Tpad + 255) / 256);
  {
    // many parts (short scans over big grids): fold in G groups first, then fold the groups.
    // The group results live behind the final arrays (offset Tpad).
    const int n_parts = R * L->v.W;
    const int G = n_parts >= 256 ? 32 : 1;
    if (G > 1) {
      k_merge<<<dim3(mb, G), 256, 0, ctx->stream>>>(p.pm1, p.pm2, p.pm3, p.pi1, p.pi2, p.psum, n_parts, p.Tpad,
                                                    p.Tpad, p.Tpad, ctx->gm1.p + Tp, ctx->gm2.p + Tp,
                                                    ctx->gm3.p + Tp, ctx->gi1.p + Tp, ctx->gi2.p + Tp,
                                                    ctx->gsum.p + Tp);
      k_merge<<<mb, 256, 0, ctx->stream>>>(ctx->gm1.p + Tp, ctx->gm2.p + Tp, ctx->gm3.p + Tp, ctx->gi1.p + Tp,
                                           ctx->gi2.p + Tp, ctx->gsum.p + Tp, G, p.Tpad, p.Tpad, p.Tpad,
                                           ctx->gm1.p, ctx->gm2.p, ctx->gm3.p, ctx->gi1.p, ctx->gi2.p, ctx->gsum.p);
      ctx->launches++;
    } else {
      k_merge<<<mb, 256, 0, ctx->stream>>>(p.pm1, p.pm2, p.pm3, p.pi1, p.pi2, p.psum, n_parts, p.Tpad, p.Tpad,
                                           p.Tpad, ctx->gm1.p, ctx->gm2.p, ctx->gm3.p, ctx->gi1.p, ctx->gi2.p,
                                           ctx->gsum.p);
    }
  }
  ctx->launches++;
  QMB_CUDA(cudaGetLastError());
  if (ctx->world > 1) {
    // pack the six arrays contiguously, all-gather, merge again
    const size_t per = Tp * (3 * sizeof(float) + 2 * sizeof(uint32_t) + sizeof(double));
    QMB_TRY(ctx->gather.ensure(per * (size_t)(ctx->world + 1)));
    unsigned char *send = ctx->gather.p, *recv = ctx->gather.p + per;
    size_t o = 0;
    QMB_CUDA(cudaMemcpyAsync(send + o, ctx->gsum.p, Tp * 8, cudaMemcpyDeviceToDevice, ctx->stream)); o += Tp * 8;
    QMB_CUDA(cudaMemcpyAsync(send + o, ctx->gm1.p, Tp * 4, cudaMemcpyDeviceToDevice, ctx->stream)); o += Tp * 4;
    QMB_CUDA(cudaMemcpyAsync(send + o, ctx->gm2.p, Tp * 4, cudaMemcpyDeviceToDevice, ctx->stream)); o += Tp * 4;
    QMB_CUDA(cudaMemcpyAsync(send + o, ctx->gm3.p, Tp * 4, cudaMemcpyDeviceToDevice, ctx->stream)); o += Tp * 4;
    QMB_CUDA(cudaMemcpyAsync(send + o, ctx->gi1.p, Tp * 4, cudaMemcpyDeviceToDevice, ctx->stream)); o += Tp * 4;
    QMB_CUDA(cudaMemcpyAsync(send + o, ctx->gi2.p, Tp * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    QMB_TRY(qmb_nccl_allgather(ctx, send, recv, per));
    const long long stride_f = (long long)(per / 4), stride_d = (long long)(per / 8);
    k_merge<<<mb, 256, 0, ctx->stream>>>((const float *)(recv + Tp * 8), (const float *)(recv + Tp * 12),
                                         (const float *)(recv + Tp * 16), (const uint32_t *)(recv + Tp * 20),
                                         (const uint32_t *)(recv + Tp * 24), (const double *)recv, ctx->world,
                                         stride_f, stride_d, p.Tpad, ctx->gm1.p, ctx->gm2.p, ctx->gm3.p, ctx->gi1.p,
                                         ctx->gi2.p, ctx->gsum.p);
    ctx->launches++;
    QMB_CUDA(cudaGetLastError());
  }

  // ---- float64 re-evaluation of the candidates + decision -------------------------------------
  ExactParams e = {};
  e.on64 = O->log64.p; e.row_stride = O->stride; e.origin = O->front + fsmp;
  e.S = L->S; e.Spad = L->Spad; e.tau16 = L->tau16.p; e.tmin = L->tmin.p;
  e.slotmap = L->slotmap.p; e.NB = L->NB;
  e.node_offset = L->node_offset; e.n_local = L->n_local;
  k_exact<<<(int)((T + 127) / 128), 128, 0, ctx->stream>>>(e, ctx->gm1.p, ctx->gm2.p, ctx->gi1.p, ctx->gi2.p,
                                                            thr, shift, T, ctx->ex1.p, ctx->ex2.p);
  ctx->launches++;
  QMB_CUDA(cudaGetLastError());
  if (ctx->world > 1) {
    QMB_TRY(qmb_nccl_allreduce_sum_f64(ctx, ctx->ex1.p, (size_t)T));
    QMB_TRY(qmb_nccl_allreduce_sum_f64(ctx, ctx->ex2.p, (size_t)T));
  }
  QMB_CUDA(cudaMemsetAsync(ctx->slow_list.p, 0, sizeof(int), ctx->stream));
  k_finalize<<<(int)((T + 255) / 256), 256, 0, ctx->stream>>>(
      ctx->gm1.p, ctx->gm2.p, ctx->gm3.p, ctx->gi1.p, ctx->gi2.p, ctx->gsum.p, ctx->ex1.p, ctx->ex2.p, thr,
      shift, T, (double)available, (double)L->n_global, ctx->r_max.p, ctx->r_norm.p, ctx->r_idx.p,
      ctx->slow_list.p, slow_cap);
  ctx->launches++;
  QMB_CUDA(cudaGetLastError());

  // ---- slow path (needs the queue length on the host) ------------------------------------------
  int n_slow = 0;
  QMB_CUDA(cudaMemcpyAsync(&n_slow, ctx->slow_list.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  QMB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (n_slow > slow_cap) n_slow = slow_cap;
  if (n_slow > 0) {
    // the queue was filled with atomics: put it in ascending sample order so that every rank
    // (and every run) walks it identically -- the cross-rank fold pairs entries by position
    std::vector<int> q((size_t)n_slow);
    QMB_CUDA(cudaMemcpyAsync(q.data(), ctx->slow_list.p + 1, (size_t)n_slow * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    QMB_CUDA(cudaStreamSynchronize(ctx->stream));
    std::sort(q.begin(), q.end());
    QMB_CUDA(cudaMemcpyAsync(ctx->slow_list.p + 1, q.data(), (size_t)n_slow * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    QMB_CUDA(cudaStreamSynchronize(ctx->stream));
    const int ctas = 4 * ctx->sm_count;
    const int n_batches = (n_slow + SLOW_BATCH - 1) / SLOW_BATCH;
    const size_t per_rank = (size_t)n_slow;
    QMB_TRY(ctx->slow_val.ensure((size_t)(ctas + ctx->world + 1) * per_rank));
    QMB_TRY(ctx->slow_idx.ensure((size_t)(ctas + ctx->world + 1) * per_rank));
    double *pv = ctx->slow_val.p;
    unsigned long long *pi = ctx->slow_idx.p;
    double *mv = pv + (size_t)ctas * per_rank;             // merged (local)  [n_slow]
    unsigned long long *mi = pi + (size_t)ctas * per_rank;
    // chunk batches over grid.y (<= 65535)
    for (int b0 = 0; b0 < n_batches; b0 += 32768) {
      const int nb = (n_batches - b0 < 32768) ? n_batches - b0 : 32768;
      k_slow<<<dim3(ctas, nb), 128, 0, ctx->stream>>>(e, ctx->slow_list.p, n_slow, b0, (double)available,
                                                      L->nodeidx.p, L->n_tiles, pv, pi);
      ctx->launches++;
      QMB_CUDA(cudaGetLastError());
    }
    k_slow_merge<<<(n_slow + 127) / 128, 128, 0, ctx->stream>>>(pv, pi, ctas, n_slow, mv, mi);
    ctx->launches++;
    QMB_CUDA(cudaGetLastError());
    if (ctx->world > 1) {
      // gather every rank's (value, index) per queued sample and fold again
      QMB_TRY(ctx->gather.ensure((size_t)(ctx->world) * per_rank * 16 + 16));
      double *gv = (double *)ctx->gather.p;
      unsigned long long *gi = (unsigned long long *)(ctx->gather.p + (size_t)ctx->world * per_rank * 8);
      QMB_TRY(qmb_nccl_allgather(ctx, mv, gv, per_rank * 8));
      QMB_TRY(qmb_nccl_allgather(ctx, mi, gi, per_rank * 8));
      k_slow_merge<<<(n_slow + 127) / 128, 128, 0, ctx->stream>>>(gv, gi, ctx->world, n_slow, mv, mi);
      ctx->launches++;
      QMB_CUDA(cudaGetLastError());
    }
    k_slow_scatter<<<(n_slow + 127) / 128, 128, 0, ctx->stream>>>(ctx->slow_list.p, n_slow, mv, mi, ctx->gsum.p,
                                                                  shift, (double)L->n_global, ctx->r_max.p,
                                                                  ctx->r_norm.p, ctx->r_idx.p);
    ctx->launches++;
    QMB_CUDA(cudaGetLastError());
  }
  QMB_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
  QMB_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->last_n_samples = (int32_t)T;
  if (kernel_ms) {
    float a = 0.f, b = 0.f, c = 0.f;
    cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&c, ctx->ev[0], ctx->ev[2]);
    kernel_ms[0] = a; kernel_ms[1] = b; kernel_ms[2] = c; kernel_ms[3] = (float)n_slow;
  }
  return QMB_OK;
}
