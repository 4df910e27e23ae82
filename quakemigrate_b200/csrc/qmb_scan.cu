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
#include <string.h>

#include <algorithm>
#include <type_traits>

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
  uint32_t q_one, q_64k, q_m64k;  // q_m64k = -65536
  int q_form;  // 0: IADD3 + LEA.HI (alu)  1: IMAD word sums  2: 1 + IMAD.HI for odd words  3: IMAD / IADD3 alternating
  // MODE 8+ (carry-free packed lanes): rows quantised to 16 - log2(G) bits, G rows are added word-wise
  // before the halves are split; tauq32 = byte offsets [tile][row group][slot][G]; Sg = S rounded up to G
  const uint32_t *tauq32;
  const uint16_t *tauq16p;  // MODE 11: 16-bit byte offsets, [tile][row pair][slot][2]
  int Sg;
  float qa, qb;  // exp2 argument of a stack: qa * sum(u) + qb = (qc0 + qstep * sum(u)) * scale
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
  constexpr uint32_t tau_bytes = (uint32_t)CR * NB * ((MODE >= 8 && MODE != 11) ? 4u : 2u);
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
          const bool act = lane < CR && s < (MODE >= 8 ? p.Sg : p.S);  // MODE 8+: zero rows pad a row group
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
          if (lane == 0) {
            if constexpr (MODE == 11)
              bulk_g2s(stage + rows_bytes, p.tauq16p + ((size_t)j * p.Spad + (size_t)c * CR) * NB, tau_bytes, full);
            else if constexpr (MODE >= 8)
              bulk_g2s(stage + rows_bytes, p.tauq32 + ((size_t)j * p.Spad + (size_t)c * CR) * NB, tau_bytes, full);
            else
              bulk_g2s(stage + rows_bytes, p.tau16 + ((size_t)j * p.Spad + (size_t)c * CR) * NB, tau_bytes, full);
          }
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

    if constexpr (MODE >= 8) {
      // ---- carry-free packed lanes -------------------------------------------------------------
      // Rows are quantised to 16 - log2(G) bits, so the two 16-bit halves of G words can be added
      // as ONE 32-bit integer without a carry crossing from the lower into the upper half:
      //   t = w_0 + ... + w_{G-1}   (IADD / IADD3: half an instruction per word)
      //   Hs += t >> 16 ; A += t    (once per G rows)
      // against two instructions per word for full 16-bit rows.  Offsets are byte offsets inside
      // the stage (row position included), G of them per (row group, node) in one vector load.
      // MODE 11: G = 2 with 16-bit byte offsets (row position included), two per 32-bit word
      constexpr bool HALF_OFFSETS = MODE == 11;
      constexpr int G = HALF_OFFSETS ? 2 : 1 << (MODE - 7);
      constexpr int H = G / 2;  // row pairs per group
      static_assert(NPW == 4 && (K % 2) == 0 && (CR % G) == 0 && G >= 2 && G <= 8, "packed scan geometry");
      constexpr int KW = K / 2;
      constexpr uint32_t OG = (uint32_t)NB * G * (HALF_OFFSETS ? 2u : 4u);  // offset bytes per row group
      int M1[K], M2[K], M3[K];  // running top-3 of the integer stacks (sum of the quantised rows)
#pragma unroll
      for (int k = 0; k < K; ++k) M1[k] = M2[k] = M3[k] = -1;
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
          const int ngroups = min(CR, p.Sg - c * CR) / G;
          const uint32_t lane_base = stage + (uint32_t)lane * 4u;
          const uint32_t oa = stage + rows_bytes + (uint32_t)(warp * NPW) * (G * (HALF_OFFSETS ? 2u : 4u));
          auto lds_v4 = [&](uint32_t addr, uint32_t *o) {
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3])
                         : "r"(addr));
          };
          auto load_rows = [&](uint32_t o0, uint32_t o1, uint32_t (&vv)[2][KW]) {
            const uint32_t a0 = lane_base + o0, a1 = lane_base + o1;
#pragma unroll
            for (int k = 0; k < KW; ++k) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(vv[0][k]) : "r"(a0 + 128u * k));
#pragma unroll
            for (int k = 0; k < KW; ++k) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(vv[1][k]) : "r"(a1 + 128u * k));
          };
          uint32_t v[2][2][KW];
          uint32_t t[KW];
          // The prefetches that reach into the next group are issued unconditionally (no second code
          // path, no register shuffling at the loop edge): after the last group of the stage they
          // re-read this group, and their results are dropped.
          if constexpr (HALF_OFFSETS) {
            // one 64-bit load carries the offsets of TWO nodes x two rows, 16 bits each:
            // ow[pair][node] = offset(row 0) | offset(row 1) << 16 -- half the shared-memory wavefronts
            // of the 32-bit form (a broadcast vector load costs one wavefront per 8 bytes)
            uint32_t ow[2][2];
            auto lds_v2 = [&](uint32_t addr, uint32_t *o) {
              asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(o[0]), "=r"(o[1]) : "r"(addr));
            };
            auto load_node = [&](uint32_t w, uint32_t (&vv)[2][KW]) { load_rows(w & 0xffffu, w >> 16, vv); };
            lds_v2(oa, ow[0]);
            lds_v2(oa + 8u, ow[1]);
            load_node(ow[0][0], v[0]);
            auto group = [&](uint32_t ogn, auto last_tag) {
              constexpr bool LAST = decltype(last_tag)::value;  // last group of the stage: nothing to prefetch
#pragma unroll
              for (int q = 0; q < NPW; ++q) {
                const int qn = (q + 1) & 3;  // q == 3: node 0 of the next group (ow[0] was refilled after q == 1)
                if (!(LAST && q == 3)) load_node(ow[qn >> 1][qn & 1], v[(q + 1) & 1]);
                uint32_t (&vv)[2][KW] = v[q & 1];
#pragma unroll
                for (int k = 0; k < KW; ++k) {
                  t[k] = vv[0][k] + vv[1][k];
                  Hs[q][k] += t[k] >> 16;
                  A[q][k] = t[k] * p.q_one + A[q][k];
                }
                if (!LAST && q == 1) lds_v2(ogn, ow[0]);
                if (!LAST && q == 3) lds_v2(ogn + 8u, ow[1]);
              }
            };
            for (int g = 0; g + 1 < ngroups; ++g) group(oa + (uint32_t)(g + 1) * OG, std::false_type{});
            group(0u, std::true_type{});
          } else if constexpr (G == 2) {
            // one vector load carries the offsets of TWO nodes x two rows: ov[pair][2 * node + row]
            uint32_t ov[2][4];
            lds_v4(oa, ov[0]);
            lds_v4(oa + 16u, ov[1]);
            load_rows(ov[0][0], ov[0][1], v[0]);
            for (int g = 0; g < ngroups; ++g) {
              const uint32_t ogn = g + 1 < ngroups ? oa + (uint32_t)(g + 1) * OG : oa + (uint32_t)g * OG;
#pragma unroll
              for (int q = 0; q < NPW; ++q) {
                const int qn = (q + 1) & 3;  // q == 3: node 0 of the next group (ov[0] was refilled after q == 1)
                load_rows(ov[qn >> 1][2 * (qn & 1)], ov[qn >> 1][2 * (qn & 1) + 1], v[(q + 1) & 1]);
                uint32_t (&vv)[2][KW] = v[q & 1];
#pragma unroll
                for (int k = 0; k < KW; ++k) {
                  t[k] = vv[0][k] + vv[1][k];
                  Hs[q][k] += t[k] >> 16;
                  A[q][k] = t[k] * p.q_one + A[q][k];
                }
                if (q == 1) lds_v4(ogn, ov[0]);
                if (q == 3) lds_v4(ogn + 16u, ov[1]);
              }
            }
          } else {
            uint32_t off[2][G];
            auto lds_off = [&](uint32_t addr, uint32_t (&o)[G]) {
#pragma unroll
              for (int i = 0; i < G; i += 4) lds_v4(addr + 4u * i, &o[i]);
            };
            lds_off(oa, off[0]);
            lds_off(oa + G * 4u, off[1]);
            load_rows(off[0][0], off[0][1], v[0]);
            for (int g = 0; g < ngroups; ++g) {
              const uint32_t og = oa + (uint32_t)g * OG;
              const uint32_t ogn = g + 1 < ngroups ? og + OG : og;
#pragma unroll
              for (int q = 0; q < NPW; ++q) {
#pragma unroll
                for (int h = 0; h < H; ++h) {
                  const int n = q * H + h;
                  // loads of the next unit first, then the adds of this one
                  if (h + 1 < H) load_rows(off[q & 1][2 * h + 2], off[q & 1][2 * h + 3], v[(n + 1) & 1]);
                  else if (q + 1 < NPW) load_rows(off[(q + 1) & 1][0], off[(q + 1) & 1][1], v[(n + 1) & 1]);
                  else load_rows(off[0][0], off[0][1], v[(n + 1) & 1]);
                  uint32_t (&vv)[2][KW] = v[n & 1];
#pragma unroll
                  for (int k = 0; k < KW; ++k) {
                    if (h == 0) t[k] = vv[0][k] + vv[1][k];
                    else t[k] = t[k] + vv[0][k] + vv[1][k];
                  }
                  if (h == H - 1) {
#pragma unroll
                    for (int k = 0; k < KW; ++k) {
                      Hs[q][k] += t[k] >> 16;
                      A[q][k] = t[k] * p.q_one + A[q][k];
                    }
                    // this node's offsets are consumed: fetch those of the node two units ahead
                    if (q + 2 < NPW) lds_off(og + (uint32_t)(q + 2) * (G * 4u), off[q & 1]);
                    else lds_off(ogn + (uint32_t)(q + 2 - NPW) * (G * 4u), off[q & 1]);
                  }
                }
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(empty);
          if (++st == p.n_stages) { st = 0; ph ^= 1u; }
        }
        // ---- per-tile epilogue: slot k = 2 kk + h <-> sample 64 kk + 2 lane + h ---------------
        // The running top-3 lives on the exact integer stacks (one float conversion per survivor
        // at the end of the item).  A stack that does not beat the current third value cannot
        // change the state, and after a few hundred tiles that is the rule: the four nodes of the
        // warp are folded into their maximum and the insertion is skipped unless some lane needs it
        // (warp vote).  Padding slots (warp-uniform, rare) take part with stack -1 = never.
        auto epilogue = [&](auto haspad_tag) {
          constexpr bool HASPAD = decltype(haspad_tag)::value;
#pragma unroll
          for (int kk = 0; kk < KW; ++kk) {
            int u[2][NPW];
            float e0 = 0.f, e1 = 0.f;
#pragma unroll
            for (int q = 0; q < NPW; ++q) {
              const bool pad = HASPAD && nidx[q] == 0xFFFFFFFFu;
              const uint32_t hi = Hs[q][kk], lo = hi * p.q_m64k + A[q][kk];
              const float x0 = ex2_approx(fmaf((float)lo, p.qa, p.qb));  // exp(stack / available)
              const float x1 = ex2_approx(fmaf((float)hi, p.qa, p.qb));
              e0 += pad ? 0.f : x0;
              e1 += pad ? 0.f : x1;
              u[0][q] = pad ? -1 : (int)lo;
              u[1][q] = pad ? -1 : (int)hi;
            }
            sum[2 * kk] += (double)e0;
            sum[2 * kk + 1] += (double)e1;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int k = 2 * kk + h;
              const int top = max(max(u[h][0], u[h][1]), max(u[h][2], u[h][3]));
              if (__any_sync(0xffffffffu, top > M3[k])) {
#pragma unroll
                for (int q = 0; q < NPW; ++q) {
                  const int v1 = u[h][q];
                  const bool g1 = v1 > M1[k];
                  const bool g2 = v1 > M2[k];
                  M3[k] = g2 ? M2[k] : max(M3[k], v1);
                  M2[k] = g1 ? M1[k] : (g2 ? v1 : M2[k]);
                  i2[k] = g1 ? i1[k] : (g2 ? nidx[q] : i2[k]);
                  M1[k] = g1 ? v1 : M1[k];
                  i1[k] = g1 ? nidx[q] : i1[k];
                }
              }
            }
          }
        };
        // padding slots are warp-uniform and rare (ragged grid edges): a separate instance
        if (nidx[0] == 0xFFFFFFFFu || nidx[1] == 0xFFFFFFFFu || nidx[2] == 0xFFFFFFFFu || nidx[3] == 0xFFFFFFFFu)
          epilogue(std::true_type{});
        else
          epilogue(std::false_type{});
      }
      // ---- item done: publish (pairs of neighbouring samples -> 8 / 16-byte stores) -----------
      const size_t base = ((size_t)r * W + warp) * p.Tpad + (size_t)tt * TT + 2u * lane;
      auto tofloat = [&](int v) { return v < 0 ? NEG_INF : fmaf((float)v, p.qstep, p.qc0); };
#pragma unroll
      for (int kk = 0; kk < KW; ++kk) {
        const size_t o = base + 64u * kk;
        *reinterpret_cast<float2 *>(p.pm1 + o) = make_float2(tofloat(M1[2 * kk]), tofloat(M1[2 * kk + 1]));
        *reinterpret_cast<float2 *>(p.pm2 + o) = make_float2(tofloat(M2[2 * kk]), tofloat(M2[2 * kk + 1]));
        *reinterpret_cast<float2 *>(p.pm3 + o) = make_float2(tofloat(M3[2 * kk]), tofloat(M3[2 * kk + 1]));
        *reinterpret_cast<uint2 *>(p.pi1 + o) = make_uint2(i1[2 * kk], i1[2 * kk + 1]);
        *reinterpret_cast<uint2 *>(p.pi2 + o) = make_uint2(i2[2 * kk], i2[2 * kk + 1]);
        *reinterpret_cast<double2 *>(p.psum + o) = make_double2(sum[2 * kk], sum[2 * kk + 1]);
      }
      continue;
    } else if constexpr (MODE >= 4) {
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
// merged per-sample state: packed [sum f64][m1 m2 m3 f32][i1 i2 u32], each array Tp long
// ---------------------------------------------------------------------------------------------
struct MState {
  double *sum;
  float *m1, *m2, *m3;
  uint32_t *i1, *i2;
};
constexpr size_t MSTATE_BYTES = 28;  // per sample
__host__ __device__ inline MState mstate(unsigned char *b, size_t Tp) {
  MState s;
  s.sum = (double *)b;
  s.m1 = (float *)(b + Tp * 8);
  s.m2 = (float *)(b + Tp * 12);
  s.m3 = (float *)(b + Tp * 16);
  s.i1 = (uint32_t *)(b + Tp * 20);
  s.i2 = (uint32_t *)(b + Tp * 24);
  return s;
}

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

// merge `n_parts` partial reductions per sample: part q lives at [q * stride + t] (float / uint32
// arrays) and [q * stride_sum + t] (sums).  blockIdx.y selects a group of parts; group g writes
// its fold into the packed state at out + g * Tp * 28 bytes (two-stage merge).
__global__ void __launch_bounds__(256) k_merge(const float *pm1, const float *pm2, const float *pm3,
                                               const uint32_t *pi1, const uint32_t *pi2,
                                               const double *psum, int n_parts, long long stride,
                                               long long stride_sum, long long n, unsigned char *out) {
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
  const MState g = mstate(out + (size_t)blockIdx.y * (size_t)n * MSTATE_BYTES, (size_t)n);
  g.m1[t] = a.m1; g.m2[t] = a.m2; g.m3[t] = a.m3;
  g.i1[t] = a.i1; g.i2[t] = a.i2;
  g.sum[t] = s;
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

// ex1/ex2[t] = float64 stack of candidate 1 / 2 when owned by this rank, else 0 (summed across
// ranks afterwards).  Candidate 2 is only evaluated when it can matter.
__global__ void __launch_bounds__(128) k_exact(ExactParams e, MState g, float thr, int shift, long long n_samples,
                                               double *ex1, double *ex2) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // output sample
  if (t >= n_samples) return;
  const long long tp = t + shift;
  double a = 0.0, b = 0.0;
  const long long c1 = g.i1[tp], c2 = g.i2[tp];
  if (c1 != 0xFFFFFFFFLL && c1 >= e.node_offset && c1 < e.node_offset + e.n_local) a = exact_stack(e, c1, t);
  if (g.m1[tp] - g.m2[tp] <= thr && c2 != 0xFFFFFFFFLL && c2 >= e.node_offset &&
      c2 < e.node_offset + e.n_local)
    b = exact_stack(e, c2, t);
  ex1[t] = a;
  ex2[t] = b;
}

// decide the winner per sample; FLAG samples that need the all-node float64 re-scan (three or
// more candidates inside the band).  The flags are compacted in sample order by k_slow_compact,
// so the queue is identical on every rank and in every run without a host sort.
__global__ void __launch_bounds__(256) k_finalize(MState g, const double *ex1, const double *ex2, float thr, int shift,
                                                  long long n_samples, double avail, double n_nodes_global,
                                                  double *r_max, double *r_norm, long long *r_idx,
                                                  unsigned char *slow_flag, unsigned int *status) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool two = false;
  if (t < n_samples) {
    const long long tp = t + shift;
    const float a1 = g.m1[tp], a2 = g.m2[tp], a3 = g.m3[tp];
    double best = exp(ex1[t] / avail);
    long long idx = g.i1[tp];
    bool slow = false;
    if (a1 - a2 <= thr) {
      two = true;
      const double other = exp(ex2[t] / avail);
      const long long oi = g.i2[tp];
      double known = ex1[t];  // best EXACT stack among the two evaluated candidates
      if (oi != 0xFFFFFFFFLL) {
        known = fmax(known, ex2[t]);
        if (other > best || (other == best && oi < idx)) {
          best = other;
          idx = oi;
        }
      }
      // Every other node (the third included) has a scan stack <= m3, hence an exact stack
      // <= m3 + delta (thr = 2 delta).  Only if that could reach the best known exact stack does the
      // sample need the all-node float64 re-scan -- roughly m1 - m3 < delta instead of < 2 delta.
      // The margin keeps exp() of the two sides distinguishable in float64.
      slow = !(known > (double)a3 + 0.5 * (double)thr + 1e-10 * fmax(1.0, fabs(known)));
    }
    slow_flag[t] = slow ? 1 : 0;
    r_max[t] = best;
    r_idx[t] = idx;
    r_norm[t] = best * n_nodes_global / g.sum[tp];
  }
  const unsigned m = __ballot_sync(0xffffffffu, two);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(&status[QMB_ST_TWO_CAND], (unsigned)__popc(m));
}

// Ordered compaction of the slow-path flags (one CTA): the samples of rank [skip, skip + cap) in
// ascending order go to slow_list[1..], their number to slow_list[0] and status[SLOW_LISTED], the
// number of flagged samples to status[SLOW_TOTAL].  `flag` is zero-padded to a multiple of 16.
__global__ void __launch_bounds__(1024) k_slow_compact(const unsigned char *flag, long long n, int skip, int cap,
                                                       int *slow_list, unsigned int *status) {
  __shared__ int wsum[32];
  __shared__ int total_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long chunk = (((n + 1023) / 1024) + 15) & ~15LL;
  const long long b0 = (long long)tid * chunk, b1 = min(n, b0 + chunk);
  int cnt = 0;
  for (long long i = b0; i < b1; i += 16) {
    const uint4 w = *reinterpret_cast<const uint4 *>(flag + i);
    cnt += __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
  }
  int incl = cnt;
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane], wi = w;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += v;
    }
    wsum[lane] = wi - w;  // exclusive
    if (lane == 31) total_s = wi;
  }
  __syncthreads();
  int r = wsum[warp] + incl - cnt;
  if (cnt) {
    for (long long i = b0; i < b1; ++i)
      if (flag[i]) {
        if (r >= skip && r < skip + cap) slow_list[1 + r - skip] = (int)i;
        ++r;
      }
  }
  if (tid == 0) {
    const int total = total_s;
    int listed = total - skip;
    listed = listed < 0 ? 0 : (listed > cap ? cap : listed);
    slow_list[0] = listed;
    status[QMB_ST_SLOW_TOTAL] = (unsigned)total;
    status[QMB_ST_SLOW_LISTED] = (unsigned)listed;
  }
}

// slow path: for each queued sample, float64 coalescence of every local node; per-CTA best
// (value, lowest index) written to partial arrays [cta][slot].  Persistent: the queue length is
// read from the device, batches of SLOW_BATCH samples are walked inside the kernel.
constexpr int SLOW_BATCH = 8;
constexpr int SLOW_CAP = 4096;  // queued samples handled per round (more: further rounds)
// Node subsets that can hold the arg-max: the scan kept one running maximum per (node range r,
// warp w) and sample -- the partial pm1[(r * W + w) * Tpad + t] -- and the exact arg-max node has a
// scan stack >= m1 - thr (exact(n*) >= exact(i1) >= m1 - delta, scan(n*) >= exact(n*) - delta), so
// only subsets whose partial maximum reaches m1 - thr need the float64 walk for that sample.
struct SlowFilter {
  const float *pm1;  // null: no filter (partials already overwritten)
  const float *gm1;  // merged m1 (all ranks)
  long long Tpad;
  int R, W, NPW, shift;
  float thr;
};
__global__ void __launch_bounds__(128) k_slow(ExactParams e, SlowFilter f, const int *slow_list, int cap, double avail,
                                              const uint32_t *nodeidx, int n_tiles, double *part_val,
                                              unsigned long long *part_idx) {
  const int n_slow = slow_list[0];
  __shared__ double sv[128];
  __shared__ unsigned long long si[128];
  const int slots_per_cta = blockDim.x;  // one thread per node slot, several tiles per pass
  const int tiles_per_pass = slots_per_cta / e.NB > 0 ? slots_per_cta / e.NB : 1;
  for (int b0 = 0; b0 < n_slow; b0 += SLOW_BATCH) {
    const int nb = min(SLOW_BATCH, n_slow - b0);
    long long ts[SLOW_BATCH];
    for (int b = 0; b < SLOW_BATCH; ++b) ts[b] = (b < nb) ? slow_list[1 + b0 + b] : 0;
    double best[SLOW_BATCH];
    unsigned long long bidx[SLOW_BATCH];
    for (int b = 0; b < SLOW_BATCH; ++b) { best[b] = -1.0; bidx[b] = ~0ull; }
    for (int j0 = blockIdx.x * tiles_per_pass; j0 < n_tiles; j0 += gridDim.x * tiles_per_pass) {
      const int j = j0 + threadIdx.x / e.NB, q = threadIdx.x % e.NB;
      if (j >= n_tiles || (int)threadIdx.x >= tiles_per_pass * e.NB) continue;
      const uint32_t gi = nodeidx[(size_t)j * e.NB + q];
      if (gi == 0xFFFFFFFFu) continue;
      unsigned hot = (1u << nb) - 1u;
      if (f.pm1) {
        // range of tile j: the r with r * n_tiles / R <= j < (r + 1) * n_tiles / R (k_scan's split)
        int r = (int)((((long long)j + 1) * f.R - 1) / n_tiles);
        while (r > 0 && (long long)r * n_tiles / f.R > j) --r;
        while (r + 1 < f.R && (long long)(r + 1) * n_tiles / f.R <= j) ++r;
        const float *pm = f.pm1 + ((size_t)r * f.W + q / f.NPW) * (size_t)f.Tpad + f.shift;
        hot = 0u;
        for (int b = 0; b < nb; ++b)
          if (pm[ts[b]] >= f.gm1[ts[b] + f.shift] - f.thr) hot |= 1u << b;
        if (!hot) continue;
      }
      double acc[SLOW_BATCH];
      for (int b = 0; b < SLOW_BATCH; ++b) acc[b] = 0.0;
      // four rows' table look-ups and gathers in flight at a time (the walk is latency-bound); the
      // additions stay in row order like the reference's
      for (int s0 = 0; s0 < e.S; s0 += 4) {
        const double *row[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int s = min(s0 + u, e.S - 1);
          const size_t m = (size_t)j * e.Spad + s;
          const long long tau = (long long)e.tmin[m] + e.tau16[m * e.NB + q];
          row[u] = e.on64 + (size_t)s * e.row_stride + e.origin + tau;
        }
        double v[4][SLOW_BATCH];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int b = 0; b < SLOW_BATCH; ++b) v[u][b] = ((hot >> b) & 1u) ? row[u][ts[b]] : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (s0 + u < e.S) {
#pragma unroll
            for (int b = 0; b < SLOW_BATCH; ++b) acc[b] += v[u][b];
          }
      }
#pragma unroll
      for (int b = 0; b < SLOW_BATCH; ++b) {
        if (!((hot >> b) & 1u)) continue;
        const double c = exp(acc[b] / avail);
        if (c > best[b] || (c == best[b] && (unsigned long long)gi < bidx[b])) { best[b] = c; bidx[b] = gi; }
      }
    }
    // block reduce (value desc, index asc)
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
        part_val[(size_t)blockIdx.x * cap + b0 + b] = sv[0];
        part_idx[(size_t)blockIdx.x * cap + b0 + b] = si[0];
      }
      __syncthreads();
    }
  }
}

// fold [n_parts][cap] partials of the listed samples (count read from the device)
__global__ void k_slow_merge(const double *part_val, const unsigned long long *part_idx, int n_parts, int cap,
                             const int *slow_list, double *out_val, unsigned long long *out_idx) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= slow_list[0]) return;
  double best = -1.0;
  unsigned long long bi = ~0ull;
  for (int q = 0; q < n_parts; ++q) {
    const double v = part_val[(size_t)q * cap + b];
    const unsigned long long i = part_idx[(size_t)q * cap + b];
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
  out_val[b] = best;
  out_idx[b] = bi;
}

__global__ void k_slow_scatter(const int *slow_list, const double *val, const unsigned long long *idx,
                               const double *gsum, int shift, double n_nodes_global, double *r_max,
                               double *r_norm, long long *r_idx) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= slow_list[0]) return;
  const int t = slow_list[1 + b];
  r_max[t] = val[b];
  r_idx[t] = (long long)idx[b];
  r_norm[t] = val[b] * n_nodes_global / gsum[t + shift];
}

// node-sum verification of the 16-bit scan: worst relative difference between the sums of the
// 16-bit scan (q_sum, this rank's nodes) and of a float32 re-scan of the same samples (f_sum)
__global__ void __launch_bounds__(256) k_check_sum(const double *f_sum, const double *q_sum, int n,
                                                   unsigned int *status) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  float rel = 0.f;
  if (t < n) {
    const double f = f_sum[t], q = q_sum[t];
    rel = (float)(fabs(q - f) / f);
    if (!(rel >= 0.f)) rel = __int_as_float(0x7fc00000);  // NaN orders above everything as bits
  }
  unsigned bits = __float_as_uint(rel);
  for (int o = 16; o > 0; o >>= 1) bits = max(bits, __shfl_xor_sync(0xffffffffu, bits, o));
  if ((threadIdx.x & 31) == 0 && bits) atomicMax(&status[QMB_ST_CHECK_ERR], bits);
}
__global__ void k_check_decide(unsigned int *status, float tol) {
  const float e = __uint_as_float(status[QMB_ST_CHECK_ERR]);
  status[QMB_ST_CHECK_BAD] = (e <= tol) ? 0u : 1u;
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

// carry-free packed lanes: G rows per group (2, 4, 8 <-> 15-, 14-, 13-bit rows)
template <int MODE>
int dispatch_scan_packed_mode(qmb_ctx *ctx, const QmbVariant &v, const ScanParams &p, int grid, size_t smem) {
  switch (v.id) {
    case 4: return launch_scan<12, 4, 8, MODE, 16>(ctx, p, grid, smem);
    case 6: return launch_scan<8, 4, 8, MODE, 16>(ctx, p, grid, smem);
    case 7: return launch_scan<12, 4, 8, MODE, 32>(ctx, p, grid, smem);
  }
  QMB_FAIL(QMB_ERR_ARG, "kernel variant %d has no packed 16-bit form", v.id);
}
bool packed_variant(const QmbVariant &v) { return v.id == 4 || v.id == 6 || v.id == 7; }
int dispatch_scan_packed(qmb_ctx *ctx, const QmbVariant &v, const ScanParams &p, int grid, size_t smem, int G) {
  if (G == 2 && p.tauq16p) return dispatch_scan_packed_mode<11>(ctx, v, p, grid, smem);
  if (G == 2) return dispatch_scan_packed_mode<8>(ctx, v, p, grid, smem);
  if (G == 4) return dispatch_scan_packed_mode<9>(ctx, v, p, grid, smem);
  if (G == 8) return dispatch_scan_packed_mode<10>(ctx, v, p, grid, smem);
  QMB_FAIL(QMB_ERR_ARG, "row group %d is not 2, 4 or 8", G);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------
// Tiling decisions shared by the detect scan and the window pass.
static int scan_setup(qmb_ctx *ctx, qmb_lut *L, qmb_onsets *O, long long first, long long T,
                      int32_t available, ScanParams &p, size_t &smem, int &shift_out, bool use_q = false,
                      int items_per_sm = 32, int qG = 1) {
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
    p.q_m64k = 0xFFFF0000u;
    p.q_form = QMB_Q16_FORM;
    if (const char *f = getenv("QMB_Q16_FORM")) p.q_form = atoi(f);
    if (qG > 1) {
      p.tauq32 = L->tauq32.p;
      p.tauq16p = getenv("QMB_WIDE_OFFSETS") ? nullptr : L->tauq16p.p;  // tuning: force the 32-bit offsets
      p.Sg = ((L->S + qG - 1) / qG) * qG;
    }
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
  const bool half_offsets = use_q && qG == 2 && L->tauq16p.p != nullptr && !getenv("QMB_WIDE_OFFSETS");
  p.stage_bytes = L->chunk_rows * p.RS * 4 + L->chunk_rows * NB * ((use_q && qG > 1 && !half_offsets) ? 4 : 2) + NB * 4;
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
  int R0 = (items_per_sm * grid + p.n_time_tiles - 1) / p.n_time_tiles;
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
  if (use_q) {
    p.qa = (float)(O->qstep * 1.4426950408889634 / (double)available);
    p.qb = (float)((double)L->S * O->qmin * 1.4426950408889634 / (double)available);
  }
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


// ---------------------------------------------------------------------------------------------
// detect: scan -> merge -> (exchange) -> float64 candidates -> decision -> slow path -> check
// ---------------------------------------------------------------------------------------------
namespace {

struct CheckWindow {
  ScanParams p;
  size_t smem;
  int shift;
  long long w0, len;
};

constexpr long long CHECK_LEN = 248;   // one time tile of the float32 kernel whatever the alignment shift
constexpr float CHECK_TOL = 2e-6f;     // node-sum difference above which the 16-bit result is discarded

int scan_run_impl(qmb_ctx *ctx, qmb_lut *L, qmb_onsets *O, int32_t fsmp, int32_t lsmp, int32_t available,
                  float *kernel_ms, bool force_fp32, float earlier_ms) {
  long long T;
  QMB_TRY(check_scan_args(L, O, fsmp, lsmp, available, T));
  ScanParams p = {};
  size_t smem;
  int shift;
  // detect scans read 16-bit fixed-point rows unless float32 was asked for (or the table / rows
  // have no 16-bit form); candidates are re-evaluated in float64 either way.  The automatic mode
  // keeps float32 rows for short scans (the verification windows would cost more than they save)
  // and verifies the node sums of every 16-bit scan (see CheckWindow).
  bool use_q = !force_fp32 && ctx->scan_fp32 != 1 && O->qstep > 0.0 && O->q16.p && L->RSq > 0 && L->tauq16.p &&
               q16_variant(L->v) && O->S <= 256;  // sum(u) < 2^24: exact in float32
  if (use_q && ctx->scan_fp32 == 0 && T < 2048) use_q = false;
  // rows quantised for carry-free groups of qG rows need the matching offset table; the two-op
  // 16-bit kernel reads any quantisation
  int qG = (O->qG > 1 && L->qG == O->qG && L->tauq32.p && packed_variant(L->v)) ? O->qG : 1;
  if (use_q && qG > 1 && scan_setup(ctx, L, O, fsmp, T, available, p, smem, shift, true, 32, qG) != QMB_OK) {
    qG = 1;
    p = ScanParams{};
  }
  if (use_q && qG == 1 && scan_setup(ctx, L, O, fsmp, T, available, p, smem, shift, true) != QMB_OK) {
    use_q = false;  // e.g. too few pipeline stages for the (slightly wider) 16-bit rows
    p = ScanParams{};
  }
  QMB_TRY(ctx->status.ensure(QMB_ST_COUNT));
  if (!ctx->h_status) QMB_CUDA(cudaHostAlloc((void **)&ctx->h_status, QMB_ST_COUNT * sizeof(unsigned int), cudaHostAllocDefault));
  if (ctx->world > 1 && !force_fp32) {
    // the operand format fixes the candidate threshold, and every rank must fold the exchanged
    // state identically: 16-bit rows only when every rank can use them
    ctx->h_status[QMB_ST_AGREE] = use_q ? (unsigned)qG : 0u;
    QMB_CUDA(cudaMemcpyAsync(ctx->status.p + QMB_ST_AGREE, ctx->h_status + QMB_ST_AGREE, sizeof(unsigned int),
                             cudaMemcpyHostToDevice, ctx->stream));
    QMB_TRY(qmb_nccl_allreduce_u32(ctx, ctx->status.p + QMB_ST_AGREE, 1, 0));
    QMB_CUDA(cudaMemcpyAsync(ctx->h_status + QMB_ST_AGREE, ctx->status.p + QMB_ST_AGREE, sizeof(unsigned int),
                             cudaMemcpyDeviceToHost, ctx->stream));
    QMB_CUDA(cudaStreamSynchronize(ctx->stream));
    // min over ranks: 0 -> float32 everywhere; 1 -> two-op 16-bit kernel everywhere; else the group
    const unsigned agreed = ctx->h_status[QMB_ST_AGREE];
    if (use_q && agreed == 0u) {
      use_q = false;
      p = ScanParams{};
    } else if (use_q && (int)agreed != qG) {
      qG = 1;
      p = ScanParams{};
      if (scan_setup(ctx, L, O, fsmp, T, available, p, smem, shift, true) != QMB_OK)
        QMB_FAIL(QMB_ERR_STATE, "ranks disagree on the scan's operand format and this rank cannot follow");
    }
  }
  if (!use_q) QMB_TRY(scan_setup(ctx, L, O, fsmp, T, available, p, smem, shift, false));
  const int R = p.R, grid = ctx->sm_count, W = L->v.W;

  // ---- verification windows of the 16-bit scan (float32 re-scan of a few time tiles) -----------
  CheckWindow chk[2];
  int n_chk = 0;
  size_t n_part_chk = 0, tp_chk = 0;
  if (use_q) {
    n_chk = T >= 16384 ? 2 : 1;
    for (int c = 0; c < n_chk; ++c) {
      CheckWindow &cw = chk[c];
      cw.len = std::min<long long>(CHECK_LEN, T);
      cw.w0 = ((2LL * c + 1) * T) / (2LL * n_chk) - cw.len / 2;
      cw.w0 = std::max<long long>(0, std::min<long long>(cw.w0, T - cw.len));
      cw.p = ScanParams{};
      QMB_TRY(scan_setup(ctx, L, O, (long long)fsmp + cw.w0, cw.len, available, cw.p, cw.smem, cw.shift, false, 4));
      n_part_chk = std::max(n_part_chk, (size_t)cw.p.R * W * (size_t)cw.p.Tpad);
      tp_chk = std::max(tp_chk, (size_t)cw.p.Tpad);
    }
  }

  const size_t n_part = std::max<size_t>((size_t)R * W * (size_t)p.Tpad, n_part_chk);
  QMB_TRY(ctx->pm1.ensure(n_part)); QMB_TRY(ctx->pm2.ensure(n_part)); QMB_TRY(ctx->pm3.ensure(n_part));
  QMB_TRY(ctx->pi1.ensure(n_part)); QMB_TRY(ctx->pi2.ensure(n_part)); QMB_TRY(ctx->psum.ensure(n_part));
  p.pm1 = ctx->pm1.p; p.pm2 = ctx->pm2.p; p.pm3 = ctx->pm3.p;
  p.pi1 = ctx->pi1.p; p.pi2 = ctx->pi2.p; p.psum = ctx->psum.p;
  const size_t Tp = (size_t)p.Tpad;
  constexpr int GROUPS = 32;
  QMB_TRY(ctx->st_loc.ensure(Tp * MSTATE_BYTES));
  QMB_TRY(ctx->st_grp.ensure(std::max(Tp, tp_chk) * MSTATE_BYTES * GROUPS));
  if (ctx->world > 1) {
    QMB_TRY(ctx->st_glob.ensure(Tp * MSTATE_BYTES));
    QMB_TRY(ctx->gather.ensure(std::max(Tp * MSTATE_BYTES, (size_t)SLOW_CAP * 16) * (size_t)ctx->world));
  }
  if (n_chk) QMB_TRY(ctx->st_chk.ensure(tp_chk * MSTATE_BYTES));
  QMB_TRY(ctx->ex.ensure(2 * (size_t)T));
  QMB_TRY(ctx->r_max.ensure((size_t)T)); QMB_TRY(ctx->r_norm.ensure((size_t)T)); QMB_TRY(ctx->r_idx.ensure((size_t)T));
  const size_t flag_bytes = (((size_t)T + 15) & ~(size_t)15) + 16;
  QMB_TRY(ctx->slow_flag.ensure(flag_bytes));
  QMB_TRY(ctx->slow_list.ensure((size_t)SLOW_CAP + 1));
  const int slow_ctas = 4 * ctx->sm_count;
  QMB_TRY(ctx->slow_val.ensure((size_t)(slow_ctas + 1) * SLOW_CAP));
  QMB_TRY(ctx->slow_idx.ensure((size_t)(slow_ctas + 1) * SLOW_CAP));

  // fp32 error bound (see file header): delta = S * 2^-24 * sum_rows max|v|, threshold 2*delta
  float thr = (float)(2.0 * 1.001 * (double)O->S * 5.9604644775390625e-08 * O->rowmax_sum) + 1e-30f;
  if (use_q) {
    // |quantised stack - exact| <= S * qstep / 2 (rounding to the grid) + the float32 evaluation of
    // qc0 + qstep * sum(u): rounding of qstep, of qc0 and of the fma, each <= 2^-24 of
    // S * (|qmin| + 65535 qstep)
    const double S = (double)O->S, mag = S * (fabs(O->qmin) + (double)O->qlevels * O->qstep);
    const double delta = 0.5 * S * O->qstep + 4.0 * 5.9604644775390625e-08 * mag;
    thr = (float)(2.0 * 1.001 * delta) + 1e-30f;
  }
  ctx->last_scan_q16 = use_q ? 1 : 0;

  QMB_CUDA(cudaMemsetAsync(ctx->status.p, 0, QMB_ST_COUNT * sizeof(unsigned int), ctx->stream));
  QMB_CUDA(cudaMemsetAsync(ctx->slow_flag.p, 0, flag_bytes, ctx->stream));
  QMB_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  if (use_q && qG > 1) QMB_TRY(dispatch_scan_packed(ctx, L->v, p, grid, smem, qG));
  else QMB_TRY(use_q ? dispatch_scan_q16(ctx, L->v, p, grid, smem) : dispatch_scan<0>(ctx, L->v, p, grid, smem));
  QMB_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));

  // ---- merge partials (local), then across ranks ----------------------------------------------
  // many parts (short scans over big grids): fold in GROUPS groups first, then fold the groups
  auto merge_parts = [&](const ScanParams &sp, unsigned char *out) -> int {
    const int n_parts = sp.R * W;
    const long long tp = sp.Tpad;
    const int mb = (int)((tp + 255) / 256);
    if (n_parts >= 256) {
      k_merge<<<dim3(mb, GROUPS), 256, 0, ctx->stream>>>(sp.pm1, sp.pm2, sp.pm3, sp.pi1, sp.pi2, sp.psum, n_parts,
                                                         tp, tp, tp, ctx->st_grp.p);
      const MState g = mstate(ctx->st_grp.p, (size_t)tp);
      k_merge<<<mb, 256, 0, ctx->stream>>>(g.m1, g.m2, g.m3, g.i1, g.i2, g.sum, GROUPS, tp * 7, (tp * 7) / 2, tp, out);
      ctx->launches += 2;
    } else {
      k_merge<<<mb, 256, 0, ctx->stream>>>(sp.pm1, sp.pm2, sp.pm3, sp.pi1, sp.pi2, sp.psum, n_parts, tp, tp, tp, out);
      ctx->launches++;
    }
    QMB_CUDA(cudaGetLastError());
    return QMB_OK;
  };
  QMB_TRY(merge_parts(p, ctx->st_loc.p));
  unsigned char *glob = ctx->st_loc.p;
  const int mb = (int)((p.Tpad + 255) / 256);
  if (ctx->world > 1) {
    // the packed local state is the send buffer: all-gather, fold the ranks identically everywhere
    const size_t per = Tp * MSTATE_BYTES;
    QMB_TRY(qmb_nccl_allgather(ctx, ctx->st_loc.p, ctx->gather.p, per));
    const MState g = mstate(ctx->gather.p, Tp);
    k_merge<<<mb, 256, 0, ctx->stream>>>(g.m1, g.m2, g.m3, g.i1, g.i2, g.sum, ctx->world, (long long)(per / 4),
                                         (long long)(per / 8), p.Tpad, ctx->st_glob.p);
    ctx->launches++;
    QMB_CUDA(cudaGetLastError());
    glob = ctx->st_glob.p;
  }
  const MState G = mstate(glob, Tp);

  // ---- float64 re-evaluation of the candidates + decision -------------------------------------
  ExactParams e = {};
  e.on64 = O->log64.p; e.row_stride = O->stride; e.origin = O->front + fsmp;
  e.S = L->S; e.Spad = L->Spad; e.tau16 = L->tau16.p; e.tmin = L->tmin.p;
  e.slotmap = L->slotmap.p; e.NB = L->NB;
  e.node_offset = L->node_offset; e.n_local = L->n_local;
  double *ex1 = ctx->ex.p, *ex2 = ctx->ex.p + T;
  k_exact<<<(int)((T + 127) / 128), 128, 0, ctx->stream>>>(e, G, thr, shift, T, ex1, ex2);
  ctx->launches++;
  QMB_CUDA(cudaGetLastError());
  if (ctx->world > 1) QMB_TRY(qmb_nccl_allreduce_sum_f64(ctx, ctx->ex.p, 2 * (size_t)T));  // one owner per candidate
  k_finalize<<<(int)((T + 255) / 256), 256, 0, ctx->stream>>>(G, ex1, ex2, thr, shift, T, (double)available,
                                                              (double)L->n_global, ctx->r_max.p, ctx->r_norm.p,
                                                              ctx->r_idx.p, ctx->slow_flag.p, ctx->status.p);
  ctx->launches++;
  QMB_CUDA(cudaGetLastError());

  // ---- slow path, driven from the device: no host round trip ------------------------------------
  double *pv = ctx->slow_val.p, *mv = pv + (size_t)slow_ctas * SLOW_CAP;
  unsigned long long *pi = ctx->slow_idx.p, *mi = pi + (size_t)slow_ctas * SLOW_CAP;
  auto slow_round = [&](int skip) -> int {
    // the scan's partial maxima are intact only until the check windows reuse their buffers
    SlowFilter f = {};
    if (skip == 0 && !getenv("QMB_SLOW_NOFILTER")) {
      f.pm1 = p.pm1; f.gm1 = G.m1; f.Tpad = p.Tpad; f.R = R; f.W = W; f.NPW = L->v.NPW; f.shift = shift; f.thr = thr;
    }
    k_slow_compact<<<1, 1024, 0, ctx->stream>>>(ctx->slow_flag.p, T, skip, SLOW_CAP, ctx->slow_list.p, ctx->status.p);
    k_slow<<<slow_ctas, 128, 0, ctx->stream>>>(e, f, ctx->slow_list.p, SLOW_CAP, (double)available, L->nodeidx.p,
                                               L->n_tiles, pv, pi);
    k_slow_merge<<<SLOW_CAP / 128, 128, 0, ctx->stream>>>(pv, pi, slow_ctas, SLOW_CAP, ctx->slow_list.p, mv, mi);
    ctx->launches += 3;
    QMB_CUDA(cudaGetLastError());
    if (ctx->world > 1) {
      // every rank lists the same samples (identical merged state): exchange a fixed-size block
      // [values | indices] per rank and fold again
      unsigned char *send = ctx->st_grp.p;  // scratch, free at this point
      QMB_CUDA(cudaMemcpyAsync(send, mv, (size_t)SLOW_CAP * 8, cudaMemcpyDeviceToDevice, ctx->stream));
      QMB_CUDA(cudaMemcpyAsync(send + (size_t)SLOW_CAP * 8, mi, (size_t)SLOW_CAP * 8, cudaMemcpyDeviceToDevice, ctx->stream));
      QMB_TRY(qmb_nccl_allgather(ctx, send, ctx->gather.p, (size_t)SLOW_CAP * 16));
      // rank r: values at gather + r * CAP * 16, indices CAP * 8 bytes further: a "part" stride of 2 * CAP
      k_slow_merge<<<SLOW_CAP / 128, 128, 0, ctx->stream>>>((const double *)ctx->gather.p,
                                                            (const unsigned long long *)(ctx->gather.p + (size_t)SLOW_CAP * 8),
                                                            ctx->world, 2 * SLOW_CAP, ctx->slow_list.p, mv, mi);
      ctx->launches++;
      QMB_CUDA(cudaGetLastError());
    }
    k_slow_scatter<<<SLOW_CAP / 128, 128, 0, ctx->stream>>>(ctx->slow_list.p, mv, mi, G.sum, shift, (double)L->n_global,
                                                            ctx->r_max.p, ctx->r_norm.p, ctx->r_idx.p);
    ctx->launches++;
    QMB_CUDA(cudaGetLastError());
    return QMB_OK;
  };
  QMB_TRY(slow_round(0));

  // ---- node-sum verification of the 16-bit scan ---------------------------------------------------
  QMB_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
  int chk_samples = 0;
  for (int c = 0; c < n_chk; ++c) {
    CheckWindow &cw = chk[c];
    cw.p.pm1 = ctx->pm1.p; cw.p.pm2 = ctx->pm2.p; cw.p.pm3 = ctx->pm3.p;
    cw.p.pi1 = ctx->pi1.p; cw.p.pi2 = ctx->pi2.p; cw.p.psum = ctx->psum.p;
    QMB_TRY(dispatch_scan<0>(ctx, L->v, cw.p, grid, cw.smem));
    QMB_TRY(merge_parts(cw.p, ctx->st_chk.p));
    const MState F = mstate(ctx->st_chk.p, (size_t)cw.p.Tpad), Q = mstate(ctx->st_loc.p, Tp);
    k_check_sum<<<(int)((cw.len + 255) / 256), 256, 0, ctx->stream>>>(F.sum + cw.shift, Q.sum + shift + cw.w0,
                                                                      (int)cw.len, ctx->status.p);
    ctx->launches++;
    QMB_CUDA(cudaGetLastError());
    chk_samples += (int)cw.len;
  }
  if (n_chk) {
    if (ctx->world > 1) QMB_TRY(qmb_nccl_allreduce_u32(ctx, ctx->status.p + QMB_ST_CHECK_ERR, 1, 1));
    k_check_decide<<<1, 1, 0, ctx->stream>>>(ctx->status.p, CHECK_TOL);
    ctx->launches++;
    QMB_CUDA(cudaGetLastError());
  }
  QMB_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
  QMB_CUDA(cudaMemcpyAsync(ctx->h_status, ctx->status.p, QMB_ST_COUNT * sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
  QMB_CUDA(cudaStreamSynchronize(ctx->stream));

  float ms_scan = 0.f, ms_tail = 0.f, ms_total = 0.f, ms_check = 0.f;
  cudaEventElapsedTime(&ms_scan, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&ms_tail, ctx->ev[1], ctx->ev[2]);
  cudaEventElapsedTime(&ms_total, ctx->ev[0], ctx->ev[2]);
  cudaEventElapsedTime(&ms_check, ctx->ev[3], ctx->ev[2]);
  const int n_slow = (int)ctx->h_status[QMB_ST_SLOW_TOTAL];
  int syncs = 0;

  const bool bad = n_chk && ctx->h_status[QMB_ST_CHECK_BAD] != 0u;
  float check_err = 0.f;
  memcpy(&check_err, &ctx->h_status[QMB_ST_CHECK_ERR], sizeof(float));
  if (bad && ctx->scan_fp32 == 0) {
    // the quantised node sums are not trustworthy on this input: redo the scan with float32 rows
    QMB_TRY(scan_run_impl(ctx, L, O, fsmp, lsmp, available, kernel_ms, true, earlier_ms + ms_total));
    ctx->info.fell_back_fp32 = 1;
    ctx->info.check_windows = n_chk;
    ctx->info.check_samples = chk_samples;
    ctx->info.check_max_rel_err = (double)check_err;
    ctx->info.check_ms = ms_check;
    ctx->info.host_syncs_in_timed_region += 1;
    return QMB_OK;
  }

  if (n_slow > SLOW_CAP) {
    // more queued samples than one round holds (pathological, e.g. all-tied data): further rounds
    for (int skip = SLOW_CAP; skip < n_slow; skip += SLOW_CAP) QMB_TRY(slow_round(skip));
    QMB_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
    QMB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ms_tail, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ms_total, ctx->ev[0], ctx->ev[2]);
    syncs = 1;
  }
  ctx->last_n_samples = (int32_t)T;
  qmb_scan_info_t &inf = ctx->info;
  inf = qmb_scan_info_t{};
  inf.operand_bytes = use_q ? 2 : 4;
  inf.operand_bits = use_q ? O->qbits : 32;
  inf.fell_back_fp32 = 0;
  inf.check_windows = n_chk;
  inf.check_samples = chk_samples;
  inf.check_max_rel_err = n_chk ? (double)check_err : 0.0;
  inf.slow_path_samples = n_slow;
  inf.two_candidate_samples = (int)ctx->h_status[QMB_ST_TWO_CAND];
  inf.scan_ms = ms_scan; inf.tail_ms = ms_tail; inf.total_ms = ms_total + earlier_ms; inf.check_ms = ms_check;
  inf.host_syncs_in_timed_region = syncs;
  if (kernel_ms) {
    kernel_ms[0] = ms_scan; kernel_ms[1] = ms_tail; kernel_ms[2] = ms_total + earlier_ms; kernel_ms[3] = (float)n_slow;
  }
  return QMB_OK;
}

}  // namespace

int qmb_scan_run(qmb_ctx *ctx, qmb_lut *L, qmb_onsets *O, int32_t fsmp, int32_t lsmp,
                 int32_t available, float *kernel_ms) {
  return scan_run_impl(ctx, L, O, fsmp, lsmp, available, kernel_ms, false, 0.f);
}
