// qmb_preprocess.cu -- waveform pre-processing on the device.
//
// The reference runs, per trace and serially on the host (obspy -> scipy.signal):
//   pre_process (quakemigrate/signal/onsets/stalta.py:203-209)
//       detrend("linear") -> detrend("constant") -> taper("cosine", 5 %) -> zero-phase Butterworth
//       band-pass (sosfilt forward, sosfilt on the reversed trace)
//   decimate    (quakemigrate/util.py:499-518)
//       the same with a low-pass at the new Nyquist, then data[::factor].
//
// A recursive (IIR) filter is a serial recurrence in time, but it is LINEAR: the state at the end
// of a span is  M^len * state_in + (state reached from zero).  So time is cut into chunks of
// PP_L samples, one thread each:
//   k_pre_partial / k_pre_stats  mean and least-squares slope of every trace (tile partial sums,
//                folded in a fixed order);
//   k_sos_state  every chunk is filtered from a ZERO state (input staged through shared memory,
//                detrend + taper applied on the fly); a Hillis-Steele scan over the tile's
//                PP_TPB chunks with the pre-computed powers M^1, M^2, M^4 ... turns the per-chunk
//                end states into "state at chunk start, given zero state at tile start";
//   k_sos_carry  one warp per trace chains the tiles: in[T+1] = M^PP_TPB * in[T] + end[T];
//   k_sos_apply  every chunk is filtered again, now from its TRUE initial state
//                (chunk_state + M^t * tile_in), i.e. with the very operation sequence of scipy's
//                sosfilt (direct form II transposed, section by section); output through shared
//                memory, coalesced, optionally keeping every `decimate`-th sample.
// The backward pass of the zero-phase filter is the same three kernels walking the trace from its
// end.  All arithmetic is float64.  HBM traffic: 8 B read (moments) + 2 x (8 + 8 + 8) B per
// sample for a zero-phase filter; the sections' arithmetic (10 flops per section and sample,
// twice) hides under it.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "qmb_internal.cuh"

namespace {

#ifndef QMB_PP_L
#define QMB_PP_L 32
#endif
#ifndef QMB_PP_TPB
#define QMB_PP_TPB 256
#endif
constexpr int PP_L = QMB_PP_L;           // samples per chunk = per thread
constexpr int PP_TPB = QMB_PP_TPB;       // chunks (threads) per tile
constexpr int PP_TILE = PP_L * PP_TPB;   // samples per tile / CTA
constexpr int PP_MAX_SEC = 4;            // second-order sections per filter (template instances)
constexpr double PP_PI = 3.14159265358979323846;

struct PreParams {
  const double *x;        // input traces
  long long x_stride;
  const int *chan_in;     // nullable: input channel of selection s (blockIdx.y); identity if null
  double *y;              // output traces
  long long y_stride;
  int chan_out0;          // output channel of selection s is chan_out0 + s
  long long n;            // samples per trace
  int n_sel, ntiles;
  int apply_pre;          // detrend + taper while loading (first pass only)
  const double *stats;    // [n_sel][2]  mean, slope about the trace centre
  double tbar;            // (n - 1) / 2
  long long wlen;         // taper half-length in samples (0: none)
  const double *taper;    // [2][wlen]: weights of the first / last wlen samples (k_taper_table)
  int decimate;           // keep samples with index % decimate == 0 (final pass only)
  double coef[PP_MAX_SEC * 5];  // per section: b0 b1 b2 a1 a2 (a0 == 1)
  const double *mpow;     // [(PP_TPB + 1)][D][D]: M^k, M = zero-input state transition of one chunk
  double *chunk_state;    // [n_sel][ntiles * PP_TPB][D]
  double *tile_state;     // [n_sel][ntiles][D]
  double *tile_in;        // [n_sel][ntiles][D]
};

// scipy.signal.sosfilt's inner loop (_sosfilt): direct form II transposed, section after section
template <int NSEC>
__device__ __forceinline__ double cascade_step(const double *c, double (&s)[2 * NSEC], double x) {
#pragma unroll
  for (int k = 0; k < NSEC; ++k) {
    const double y = c[5 * k + 0] * x + s[2 * k];
    s[2 * k] = c[5 * k + 1] * x - c[5 * k + 3] * y + s[2 * k + 1];
    s[2 * k + 1] = c[5 * k + 2] * x - c[5 * k + 4] * y;
    x = y;
  }
  return x;
}

// cosine taper weight of sample i (obspy Trace.taper(type="cosine") with wlen samples per side)
__device__ __noinline__ double taper_weight(long long i, long long n, long long wlen) {
  const long long w1 = wlen - 1;
  if (i < wlen)  // 0.5 * (1 - cos(pi * k / (wlen - 1))), k = i
    return w1 > 0 ? 0.5 * (1.0 - cos(PP_PI * (double)i / (double)w1)) : 0.0;
  if (i >= n - wlen) {  // 0.5 * (1 + cos(pi * j / (wlen - 1))), j = i - (n - wlen)
    const long long j = i - (n - wlen);
    return w1 > 0 ? 0.5 * (1.0 + cos(PP_PI * (double)j / (double)w1)) : 0.0;
  }
  return 1.0;
}

// The taper weights of a call, once: the tiles at either end of a trace would otherwise spend longer
// in cos() than the others spend filtering (and, in the single-pass form, keep their successors waiting).
__global__ void __launch_bounds__(256) k_taper_table(double *tw, long long n, long long wlen) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= wlen) return;
  tw[k] = taper_weight(k, n, wlen);
  tw[wlen + k] = taper_weight(n - wlen + k, n, wlen);
}

// detrend("linear") + detrend("constant") + cosine taper of sample i (element-wise kernels)
__device__ __forceinline__ double pre_sample(const PreParams &p, double v, long long i, double mean,
                                             double slope) {
  v -= mean + slope * ((double)i - p.tbar);
  if (p.wlen > 0) {
    if (i < p.wlen) v *= p.taper[i];
    else if (i >= p.n - p.wlen) v *= p.taper[p.wlen + (i - (p.n - p.wlen))];
  }
  return v;
}

// Stage tile T of selection s in shared memory, in PROCESSING order q (q-th processed sample is
// x[q] forward, x[n-1-q] backward); chunk t occupies tile[t * (PP_L + 1) .. + PP_L).
// All PP_L loads of a thread are issued before the first use (one DRAM latency per tile, not 32);
// the line is removed in registers; the taper touches only the tiles at either end of the trace
// (CTA-uniform branch, rolled loop: the cosine code exists once).
template <bool REV>
__device__ __forceinline__ void load_tile(const PreParams &p, double *tile, int T, int s) {
  const int ch = p.chan_in ? p.chan_in[s] : s;
  const double *__restrict__ x = p.x + (size_t)ch * p.x_stride;
  const long long q0 = (long long)T * PP_TILE + threadIdx.x;
  double v[PP_L];
#pragma unroll
  for (int j = 0; j < PP_L; ++j) {
    const long long q = q0 + (long long)j * PP_TPB;
    v[j] = q < p.n ? __ldg(x + (REV ? p.n - 1 - q : q)) : 0.0;
  }
  // k = threadIdx.x + j * PP_TPB  ->  k / PP_L = threadIdx.x / PP_L + j * (PP_TPB / PP_L)
  double *dst = tile + threadIdx.x + threadIdx.x / PP_L;
  constexpr int STEP = PP_TPB + PP_TPB / PP_L;
  if (p.apply_pre) {
    const double slope = p.stats[2 * s + 1];
    const double step = REV ? -(double)PP_TPB : (double)PP_TPB;
    // value of the fitted line at this thread's first sample; later samples are +-PP_TPB apart
    const double line0 = p.stats[2 * s] + slope * ((double)(REV ? p.n - 1 - q0 : q0) - p.tbar);
#pragma unroll
    for (int j = 0; j < PP_L; ++j) {
      const long long q = q0 + (long long)j * PP_TPB;
      if (q < p.n) v[j] -= line0 + slope * (step * (double)j);
    }
  }
#pragma unroll
  for (int j = 0; j < PP_L; ++j) dst[j * STEP] = v[j];
  if (p.apply_pre && p.wlen > 0) {
    // samples of this tile: processing positions [qa, qb] <-> trace indices [lo, hi]
    const long long qa = (long long)T * PP_TILE, qb = min(qa + PP_TILE, p.n) - 1;
    const long long lo = REV ? p.n - 1 - qb : qa, hi = REV ? p.n - 1 - qa : qb;
    if (lo < p.wlen || hi >= p.n - p.wlen) {
#pragma unroll 1
      for (int j = 0; j < PP_L; ++j) {
        const long long q = q0 + (long long)j * PP_TPB;
        if (q >= p.n) break;
        const long long i = REV ? p.n - 1 - q : q;
        if (i < p.wlen) dst[j * STEP] *= p.taper[i];
        else if (i >= p.n - p.wlen) dst[j * STEP] *= p.taper[p.wlen + (i - (p.n - p.wlen))];
      }
    }
  }
}

// ---- per-trace mean and slope (least-squares line) ------------------------------------------------
// k_pre_partial: one CTA per tile, sum x and sum (i - tbar) x of its 8192 samples -> partial[s][T][2]
// k_pre_stats  : one warp per trace folds the partials in a fixed order (deterministic)
__global__ void __launch_bounds__(PP_TPB) k_pre_partial(PreParams p, double *partial) {
  const int T = blockIdx.x, s = blockIdx.y;
  const int ch = p.chan_in ? p.chan_in[s] : s;
  const double *__restrict__ x = p.x + (size_t)ch * p.x_stride;
  const long long q0 = (long long)T * PP_TILE + threadIdx.x;
  double v[PP_L];
#pragma unroll
  for (int j = 0; j < PP_L; ++j) {
    const long long q = q0 + (long long)j * PP_TPB;
    v[j] = q < p.n ? __ldg(x + q) : 0.0;
  }
  double s0 = 0.0, s1 = 0.0;
#pragma unroll
  for (int j = 0; j < PP_L; ++j) {
    const long long q = q0 + (long long)j * PP_TPB;
    s0 += v[j];
    s1 += ((double)q - p.tbar) * v[j];
  }
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  __shared__ double w0[PP_TPB / 32], w1[PP_TPB / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    w0[warp] = s0;
    w1[warp] = s1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    s0 = 0.0;
    s1 = 0.0;
    for (int w = 0; w < PP_TPB / 32; ++w) {
      s0 += w0[w];
      s1 += w1[w];
    }
    partial[2 * ((size_t)s * p.ntiles + T)] = s0;
    partial[2 * ((size_t)s * p.ntiles + T) + 1] = s1;
  }
}

__global__ void __launch_bounds__(32) k_pre_stats(PreParams p, const double *partial, double *stats, int detrend) {
  const int s = blockIdx.x, lane = threadIdx.x;
  double s0 = 0.0, s1 = 0.0;
  if (detrend)
    for (int T = lane; T < p.ntiles; T += 32) {
      s0 += partial[2 * ((size_t)s * p.ntiles + T)];
      s1 += partial[2 * ((size_t)s * p.ntiles + T) + 1];
    }
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  if (lane == 0) {
    const double nn = (double)p.n;
    const double sxx = nn * (nn * nn - 1.0) / 12.0;  // sum (i - tbar)^2
    stats[2 * s] = detrend ? s0 / nn : 0.0;
    stats[2 * s + 1] = (detrend && p.n > 1) ? s1 / sxx : 0.0;
  }
}

// ---- pass 1: zero-state chunk responses + in-tile scan -------------------------------------------
template <int NSEC, bool REV>
__global__ void __launch_bounds__(PP_TPB) k_sos_state(PreParams p) {
  constexpr int D = 2 * NSEC;
  extern __shared__ double sm[];
  double *tile = sm;                          // PP_TPB * (PP_L + 1)
  double *sv = sm + PP_TPB * (PP_L + 1);      // PP_TPB * D
  const int t = threadIdx.x, T = blockIdx.x, s = blockIdx.y;
  load_tile<REV>(p, tile, T, s);
  __syncthreads();
  double st[D];
#pragma unroll
  for (int d = 0; d < D; ++d) st[d] = 0.0;
  const double *my = tile + t * (PP_L + 1);
#pragma unroll 4
  for (int k = 0; k < PP_L; ++k) cascade_step<NSEC>(p.coef, st, my[k]);
#pragma unroll
  for (int d = 0; d < D; ++d) sv[t * D + d] = st[d];
  // inclusive scan: after step o, st = end state of chunk t given zero state o chunks earlier
  for (int o = 1; o < PP_TPB; o <<= 1) {
    __syncthreads();
    double add[D];
    if (t >= o) {
      const double *M = p.mpow + (size_t)o * D * D;
      const double *prev = sv + (t - o) * D;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) acc += M[i * D + j] * prev[j];
        add[i] = acc;
      }
    }
    __syncthreads();
    if (t >= o) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        st[i] += add[i];
        sv[t * D + i] = st[i];
      }
    }
  }
  __syncthreads();
  double *cst = p.chunk_state + (((size_t)s * p.ntiles + T) * PP_TPB + t) * D;
#pragma unroll
  for (int d = 0; d < D; ++d) cst[d] = t ? sv[(t - 1) * D + d] : 0.0;
  if (t == PP_TPB - 1) {
    double *ts = p.tile_state + ((size_t)s * p.ntiles + T) * D;
#pragma unroll
    for (int d = 0; d < D; ++d) ts[d] = st[d];
  }
}

// ---- pass 2: chain the tiles of each trace --------------------------------------------------------
// One warp per trace.  32 tile end states at a time are loaded side by side into shared memory
// (one global latency per 32 tiles instead of one per tile), lane 0 runs the short recurrence, and
// the 32 incoming states go out side by side again.
template <int NSEC>
__global__ void __launch_bounds__(32) k_sos_carry(PreParams p) {
  constexpr int D = 2 * NSEC;
  __shared__ double zs[32 * D], ins[32 * D], Ms[D * D];
  const int s = blockIdx.x, lane = threadIdx.x;
  for (int i = lane; i < D * D; i += 32) Ms[i] = p.mpow[(size_t)PP_TPB * D * D + i];
  double cur[D];
#pragma unroll
  for (int d = 0; d < D; ++d) cur[d] = 0.0;
  for (int T0 = 0; T0 < p.ntiles; T0 += 32) {
    const int cnt = min(32, p.ntiles - T0);
    const double *z = p.tile_state + ((size_t)s * p.ntiles + T0) * D;
    for (int i = lane; i < cnt * D; i += 32) zs[i] = z[i];
    __syncwarp();
    if (lane == 0) {
      for (int k = 0; k < cnt; ++k) {
        double nx[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
          ins[k * D + i] = cur[i];
          double acc = zs[k * D + i];
#pragma unroll
          for (int j = 0; j < D; ++j) acc += Ms[i * D + j] * cur[j];
          nx[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < D; ++i) cur[i] = nx[i];
      }
    }
    __syncwarp();
    double *tin = p.tile_in + ((size_t)s * p.ntiles + T0) * D;
    for (int i = lane; i < cnt * D; i += 32) tin[i] = ins[i];
    __syncwarp();
  }
}

// ---- pass 3: filter every chunk from its true initial state ----------------------------------------
template <int NSEC, bool REV>
__global__ void __launch_bounds__(PP_TPB) k_sos_apply(PreParams p) {
  constexpr int D = 2 * NSEC;
  extern __shared__ double sm[];
  double *tile = sm;
  const int t = threadIdx.x, T = blockIdx.x, s = blockIdx.y;
  load_tile<REV>(p, tile, T, s);
  const double *tin = p.tile_in + ((size_t)s * p.ntiles + T) * D;
  const double *cst = p.chunk_state + (((size_t)s * p.ntiles + T) * PP_TPB + t) * D;
  const double *M = p.mpow + (size_t)t * D * D;
  double st[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    double acc = cst[i];
#pragma unroll
    for (int j = 0; j < D; ++j) acc += M[i * D + j] * tin[j];
    st[i] = acc;
  }
  __syncthreads();
  double *my = tile + t * (PP_L + 1);
#pragma unroll 4
  for (int k = 0; k < PP_L; ++k) my[k] = cascade_step<NSEC>(p.coef, st, my[k]);
  __syncthreads();
  double *y = p.y + (size_t)(p.chan_out0 + s) * p.y_stride;
  for (int k = t; k < PP_TILE; k += PP_TPB) {
    const long long q = (long long)T * PP_TILE + k;
    if (q >= p.n) break;
    const long long i = REV ? p.n - 1 - q : q;
    if (p.decimate <= 1) y[i] = tile[k + k / PP_L];
    else if (i % p.decimate == 0) y[i / p.decimate] = tile[k + k / PP_L];
  }
}

// ---- single pass per direction ---------------------------------------------------------------------
// One kernel does what k_sos_state + k_sos_carry + k_sos_apply do, reading every sample once and
// writing it once (16 bytes per sample and direction instead of 24):
//   1. the tile is staged in shared memory and every chunk is filtered IN PLACE from a zero state
//      (zero-state response y0, chunk end states);
//   2. the in-tile scan turns the end states into chunk-start states relative to the tile start;
//   3. the state at the tile start comes from the predecessors of the same trace by a decoupled
//      look-back: every tile publishes its aggregate (end state from a zero start) as soon as it has
//      it, and its inclusive end state once it knows its own start; a tile walks back until it meets
//      an inclusive state and replays  in <- M^256 in + z  forward from there (tiles are handed out
//      in launch order by a ticket, so every predecessor is running or done);
//   4. the filter is linear: the true output is y0 plus the zero-input response of the cascade to
//      the chunk's true start state, out[k] = y0[k] + sum_d resp[d][k] * state[d] -- D multiply-adds
//      per sample, independent across samples, instead of a second serial pass of the cascade.
// Rounding differs from scipy's sosfilt at the 1e-15 level of the trace's peak (the same level as
// the three-kernel form, whose re-associated state propagation this shares).
struct alignas(16) DescEntry {
  double v;
  unsigned long long valid;
};
__device__ __forceinline__ void desc_store(DescEntry *e, double v) {
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(e), "l"(__double_as_longlong(v)), "l"(1ull) : "memory");
}
__device__ __forceinline__ bool desc_load(const DescEntry *e, double &v) {
  unsigned long long a, b;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(e) : "memory");
  v = __longlong_as_double((long long)a);
  return b != 0ull;
}

struct FusedScratch {
  int *ticket;      // [1] next tile to hand out (4 ints reserved)
  DescEntry *desc;  // [n_sel * ntiles][2][D]: aggregate, inclusive end state; zeroed before every pass
  const double *resp;  // [D][PP_L] zero-input response of the cascade output to unit start states
  const double *ppow;  // [ntiles][D][D] powers of the tile transition P = M^PP_TPB
};

template <int NSEC, bool REV>
__global__ void __launch_bounds__(PP_TPB) k_sos_fused(PreParams p, FusedScratch f) {
  constexpr int D = 2 * NSEC;
  extern __shared__ double sm[];
  double *tile = sm;                            // PP_TPB * (PP_L + 1)
  double *sv = tile + PP_TPB * (PP_L + 1);      // PP_TPB * D: scan buffer, then true chunk-start states
  double *tin_s = sv + PP_TPB * D;              // D
  double *z_s = tin_s + D;                      // D: the tile's aggregate
  double *R = z_s + D;                          // D * PP_L
  __shared__ int s_ticket;
  const int t = threadIdx.x;
  if (t == 0) s_ticket = atomicAdd(f.ticket, 1);
  for (int i = t; i < D * PP_L; i += PP_TPB) R[i] = f.resp[i];
  __syncthreads();
  const int T = s_ticket % p.ntiles, s = s_ticket / p.ntiles;
  load_tile<REV>(p, tile, T, s);
  __syncthreads();
  // 1. zero-state response in place
  double st[D];
#pragma unroll
  for (int d = 0; d < D; ++d) st[d] = 0.0;
  double *my = tile + t * (PP_L + 1);
#pragma unroll 4
  for (int k = 0; k < PP_L; ++k) my[k] = cascade_step<NSEC>(p.coef, st, my[k]);
#pragma unroll
  for (int d = 0; d < D; ++d) sv[t * D + d] = st[d];
  // 2. inclusive scan over the chunks of the tile
  for (int o = 1; o < PP_TPB; o <<= 1) {
    __syncthreads();
    double add[D];
    if (t >= o) {
      const double *M = p.mpow + (size_t)o * D * D;
      const double *prev = sv + (t - o) * D;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) acc += M[i * D + j] * prev[j];
        add[i] = acc;
      }
    }
    __syncthreads();
    if (t >= o) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        st[i] += add[i];
        sv[t * D + i] = st[i];
      }
    }
  }
  __syncthreads();
  double c0[D];  // state at the start of this chunk, given a zero state at the start of the tile
#pragma unroll
  for (int d = 0; d < D; ++d) c0[d] = t ? sv[(t - 1) * D + d] : 0.0;
  // 3. look-back, one warp: lane l inspects the predecessor at distance 32 w + l, all 32 at once.
  // A descriptor entry is 16 bytes, {value, valid}, written and read with ONE vector access, so a
  // reader sees either nothing or the value: no flags array, no memory fences.
  if (t == PP_TPB - 1) {
#pragma unroll
    for (int d = 0; d < D; ++d) z_s[d] = st[d];
  }
  __syncthreads();
  if (t < 32) {
    const int lane = t;
    const size_t row = (size_t)s * p.ntiles;
    DescEntry *mine = f.desc + (row + T) * 2 * D;  // [0, D): aggregate, [D, 2 D): inclusive end state
    if (T > 0 && lane < D) desc_store(mine + lane, z_s[lane]);
    double acc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] = 0.0;
    // everything that does not depend on the predecessors is fetched before the first poll, so the
    // critical path of the look-back is one round trip to the descriptors, not three
    double Prow[D];
#pragma unroll
    for (int q = 0; q < D; ++q) Prow[q] = lane < D ? p.mpow[(size_t)PP_TPB * D * D + lane * D + q] : 0.0;
    bool done = T == 0;
    for (int base = 0; !done; base += 32) {
      const int dist = base + lane, j = T - 1 - dist;
      constexpr bool PREFETCH_P = D <= 4;  // 16 values; 36 or 64 would spill
      double v[D], Pd[PREFETCH_P ? D * D : 1];
      bool incl = false;
      if (j >= 0) {
        if constexpr (PREFETCH_P) {
#pragma unroll
          for (int q = 0; q < D * D; ++q) Pd[q] = f.ppow[(size_t)dist * D * D + q];
        }
        const DescEntry *e = f.desc + (row + j) * 2 * D;
        for (;;) {  // the inclusive state if it is there, else the aggregate, else wait; one round trip per poll
          double vi[D];
          bool ok_i = true, ok_a = true;
#pragma unroll
          for (int q = 0; q < D; ++q) ok_i &= desc_load(e + D + q, vi[q]);
#pragma unroll
          for (int q = 0; q < D; ++q) ok_a &= desc_load(e + q, v[q]);
          if (ok_i) {
#pragma unroll
            for (int q = 0; q < D; ++q) v[q] = vi[q];
            incl = true;
            break;
          }
          if (ok_a) break;
        }
      }
      const unsigned has_incl = __ballot_sync(0xffffffffu, incl);
      const int k = has_incl ? __ffs(has_incl) - 1 : 32;  // nearest tile of this window with a known end state
      if (j >= 0 && lane <= k) {
        // in[T] = sum_{d < dist_k} P^d z_{T-1-d} + P^dist_k E_{T-1-dist_k}
        const double *Pg = f.ppow + (size_t)dist * D * D;
#pragma unroll
        for (int i = 0; i < D; ++i) {
          double a2 = 0.0;
#pragma unroll
          for (int q = 0; q < D; ++q) a2 += (PREFETCH_P ? Pd[i * D + q] : Pg[i * D + q]) * v[q];
          acc[i] += a2;
        }
      }
      done = has_incl != 0u;  // tile 0 only ever publishes an inclusive state, so the walk ends there at the latest
    }
#pragma unroll
    for (int d = 0; d < D; ++d)
      for (int o = 16; o > 0; o >>= 1) acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], o);
    if (lane < D) {  // lane i: component i of the inclusive end state P in + z
      double e = z_s[lane];
#pragma unroll
      for (int q = 0; q < D; ++q) e += Prow[q] * acc[q];
      desc_store(mine + D + lane, e);
    }
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < D; ++i) tin_s[i] = acc[i];
    }
  }
  __syncthreads();
  // 4. true start state of every chunk, then the zero-input correction, coalesced
  {
    const double *M = p.mpow + (size_t)t * D * D;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double acc = c0[i];
#pragma unroll
      for (int j = 0; j < D; ++j) acc += M[i * D + j] * tin_s[j];
      sv[t * D + i] = acc;
    }
  }
  __syncthreads();
  double *y = p.y + (size_t)(p.chan_out0 + s) * p.y_stride;
  for (int k = t; k < PP_TILE; k += PP_TPB) {
    const long long q = (long long)T * PP_TILE + k;
    if (q >= p.n) break;
    const int c = k / PP_L, r = k % PP_L;
    double v = tile[k + c];
#pragma unroll
    for (int d = 0; d < D; ++d) v += R[d * PP_L + r] * sv[c * D + d];
    const long long i = REV ? p.n - 1 - q : q;
    if (p.decimate <= 1) y[i] = v;
    else if (i % p.decimate == 0) y[i / p.decimate] = v;
  }
}

// detrend + taper only (no filter sections)
__global__ void __launch_bounds__(256) k_pre_only(PreParams p) {
  const int s = blockIdx.y;
  const int ch = p.chan_in ? p.chan_in[s] : s;
  const double *x = p.x + (size_t)ch * p.x_stride;
  double *y = p.y + (size_t)(p.chan_out0 + s) * p.y_stride;
  const double mean = p.stats[2 * s], slope = p.stats[2 * s + 1];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const double v = pre_sample(p, x[i], i, mean, slope);
  if (p.decimate <= 1) y[i] = v;
  else if (i % p.decimate == 0) y[i / p.decimate] = v;
}

// M[i][j] = state i after `steps` zero-input samples starting from unit state j
void chunk_transition(const double *coef, int nsec, int steps, std::vector<double> &M) {
  const int D = 2 * nsec;
  M.assign((size_t)D * D, 0.0);
  std::vector<double> s(D);
  for (int j = 0; j < D; ++j) {
    for (int d = 0; d < D; ++d) s[d] = d == j ? 1.0 : 0.0;
    for (int k = 0; k < steps; ++k) {
      double x = 0.0;
      for (int q = 0; q < nsec; ++q) {
        const double *c = coef + 5 * q;
        const double y = c[0] * x + s[2 * q];
        s[2 * q] = c[1] * x - c[3] * y + s[2 * q + 1];
        s[2 * q + 1] = c[2] * x - c[4] * y;
        x = y;
      }
    }
    for (int i = 0; i < D; ++i) M[(size_t)i * D + j] = s[i];
  }
}

// resp[d][k] = output of the cascade at step k of a zero-input run started from unit state d
void zero_input_response(const double *coef, int nsec, int steps, std::vector<double> &R) {
  const int D = 2 * nsec;
  R.assign((size_t)D * steps, 0.0);
  std::vector<double> s(D);
  for (int j = 0; j < D; ++j) {
    for (int d = 0; d < D; ++d) s[d] = d == j ? 1.0 : 0.0;
    for (int k = 0; k < steps; ++k) {
      double x = 0.0;
      for (int q = 0; q < nsec; ++q) {
        const double *c = coef + 5 * q;
        const double y = c[0] * x + s[2 * q];
        s[2 * q] = c[1] * x - c[3] * y + s[2 * q + 1];
        s[2 * q + 1] = c[2] * x - c[4] * y;
        x = y;
      }
      R[(size_t)j * steps + k] = x;
    }
  }
}

template <int NSEC>
int run_pass_fused(qmb_ctx *ctx, PreParams &p, const FusedScratch &f, bool rev) {
  constexpr int D = 2 * NSEC;
  const size_t smem = (size_t)(PP_TPB * (PP_L + 1) + PP_TPB * D + 2 * D + D * PP_L) * sizeof(double);
  const unsigned n_ctas = (unsigned)p.ntiles * (unsigned)p.n_sel;
  // ticket + descriptors are contiguous: one memset
  QMB_CUDA(cudaMemsetAsync(f.ticket, 0, 16 + (size_t)p.ntiles * p.n_sel * 2 * D * sizeof(DescEntry), ctx->stream));
  if (rev) {
    QMB_CUDA(cudaFuncSetAttribute(k_sos_fused<NSEC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_sos_fused<NSEC, true><<<n_ctas, PP_TPB, smem, ctx->stream>>>(p, f);
  } else {
    QMB_CUDA(cudaFuncSetAttribute(k_sos_fused<NSEC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_sos_fused<NSEC, false><<<n_ctas, PP_TPB, smem, ctx->stream>>>(p, f);
  }
  QMB_CUDA(cudaGetLastError());
  ctx->launches++;
  return QMB_OK;
}

int run_pass_fused_n(qmb_ctx *ctx, int nsec, PreParams &p, const FusedScratch &f, bool rev) {
  switch (nsec) {
    case 1: return run_pass_fused<1>(ctx, p, f, rev);
    case 2: return run_pass_fused<2>(ctx, p, f, rev);
    case 3: return run_pass_fused<3>(ctx, p, f, rev);
    case 4: return run_pass_fused<4>(ctx, p, f, rev);
  }
  QMB_FAIL(QMB_ERR_UNSUPPORTED, "filters with %d second-order sections are not built (1..%d)", nsec, PP_MAX_SEC);
}

template <int NSEC>
int run_pass(qmb_ctx *ctx, PreParams &p, bool rev) {
  constexpr int D = 2 * NSEC;
  const size_t smem_state = (size_t)(PP_TPB * (PP_L + 1) + PP_TPB * D) * sizeof(double);
  const size_t smem_apply = (size_t)(PP_TPB * (PP_L + 1)) * sizeof(double);
  dim3 grid((unsigned)p.ntiles, (unsigned)p.n_sel);
  if (rev) {
    QMB_CUDA(cudaFuncSetAttribute(k_sos_state<NSEC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_state));
    QMB_CUDA(cudaFuncSetAttribute(k_sos_apply<NSEC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_apply));
    k_sos_state<NSEC, true><<<grid, PP_TPB, smem_state, ctx->stream>>>(p);
  } else {
    QMB_CUDA(cudaFuncSetAttribute(k_sos_state<NSEC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_state));
    QMB_CUDA(cudaFuncSetAttribute(k_sos_apply<NSEC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_apply));
    k_sos_state<NSEC, false><<<grid, PP_TPB, smem_state, ctx->stream>>>(p);
  }
  QMB_CUDA(cudaGetLastError());
  k_sos_carry<NSEC><<<p.n_sel, 32, 0, ctx->stream>>>(p);
  QMB_CUDA(cudaGetLastError());
  if (rev) k_sos_apply<NSEC, true><<<grid, PP_TPB, smem_apply, ctx->stream>>>(p);
  else k_sos_apply<NSEC, false><<<grid, PP_TPB, smem_apply, ctx->stream>>>(p);
  QMB_CUDA(cudaGetLastError());
  ctx->launches += 3;
  return QMB_OK;
}

int run_pass_n(qmb_ctx *ctx, int nsec, PreParams &p, bool rev) {
  switch (nsec) {
    case 1: return run_pass<1>(ctx, p, rev);
    case 2: return run_pass<2>(ctx, p, rev);
    case 3: return run_pass<3>(ctx, p, rev);
    case 4: return run_pass<4>(ctx, p, rev);
  }
  QMB_FAIL(QMB_ERR_UNSUPPORTED, "filters with %d second-order sections are not built (1..%d)", nsec, PP_MAX_SEC);
}

}  // namespace

// Device-level entry: x (any channels, selected through d_chan_in) -> out channels
// [out_chan0, out_chan0 + n_sel).  sos: host (nsec, 6) rows b0 b1 b2 a0 a1 a2 like scipy.
int qmb_preprocess_device(qmb_ctx *ctx, const double *d_x, int64_t x_stride, const int *d_chan_in,
                          int32_t n_sel, int64_t n, const double *sos, int32_t nsec, int32_t detrend,
                          double taper_fraction, int32_t zerophase, int32_t decimate, double *d_out,
                          int64_t out_stride, int32_t out_chan0, float *kernel_ms) {
  if (n_sel <= 0 || n <= 0) return QMB_OK;
  if (nsec < 0 || nsec > PP_MAX_SEC)
    QMB_FAIL(QMB_ERR_UNSUPPORTED, "filters with %d second-order sections are not built (0..%d)", nsec, PP_MAX_SEC);
  if (nsec > 0 && !sos) QMB_FAIL(QMB_ERR_ARG, "null filter coefficients");
  if (decimate < 1) QMB_FAIL(QMB_ERR_ARG, "decimation factor must be >= 1");
  if (n_sel > 65535) QMB_FAIL(QMB_ERR_UNSUPPORTED, "more than 65535 traces per call (%d): split the block", n_sel);
  if (taper_fraction < 0.0 || taper_fraction > 0.5) QMB_FAIL(QMB_ERR_ARG, "taper fraction must be in [0, 0.5]");
  const int D = 2 * nsec;
  PreParams p = {};
  p.x = d_x; p.x_stride = x_stride; p.chan_in = d_chan_in;
  p.n = n; p.n_sel = n_sel; p.ntiles = (int)((n + PP_TILE - 1) / PP_TILE);
  p.tbar = 0.5 * (double)(n - 1);
  {  // Trace.taper: wlen = min(int(max_percentage * npts), int(npts / 2))
    long long w = (long long)(taper_fraction * (double)n);
    if (w > n / 2) w = n / 2;
    p.wlen = w;
  }
  for (int k = 0; k < nsec; ++k) {
    const double a0 = sos[6 * k + 3];
    if (a0 == 0.0) QMB_FAIL(QMB_ERR_ARG, "section %d has a0 == 0", k);
    p.coef[5 * k + 0] = sos[6 * k + 0] / a0;
    p.coef[5 * k + 1] = sos[6 * k + 1] / a0;
    p.coef[5 * k + 2] = sos[6 * k + 2] / a0;
    p.coef[5 * k + 3] = sos[6 * k + 4] / a0;
    p.coef[5 * k + 4] = sos[6 * k + 5] / a0;
  }
  // scratch: stats | partial sums | mpow | chunk_state | tile_state | tile_in | forward output
  const size_t n_stats = (size_t)2 * n_sel + (size_t)2 * n_sel * p.ntiles;
  const size_t n_mpow = (size_t)(PP_TPB + 1) * D * D;
  const size_t n_chunk = (size_t)n_sel * p.ntiles * PP_TPB * D;
  const size_t n_tile = (size_t)n_sel * p.ntiles * D;
  const size_t n_tmp = (nsec > 0 && zerophase) ? (size_t)n_sel * n : 0;
  // single-pass form (default): no per-chunk state arrays, but tile descriptors + the response table
  const bool fused = nsec > 0 && !getenv("QMB_PP_LEGACY");
  const size_t n_resp = (size_t)D * PP_L + (fused ? (size_t)p.ntiles * D * D : 0);  // response table, then P^d
  const size_t n_desc = fused ? (size_t)n_sel * p.ntiles * 2 * D * 2 : 0;  // 16-byte entries, counted in doubles
  const size_t n_flags = fused ? 2 : 0;                                     // the ticket (16 bytes)
  const size_t n_taper = (size_t)2 * p.wlen;
  QMB_TRY(ctx->pp_scratch.ensure(n_stats + n_taper + n_mpow + n_resp + (fused ? n_desc + n_flags + 2 : n_chunk + 2 * n_tile) + n_tmp + 8));
  double *base = ctx->pp_scratch.p;
  double *d_stats = base; base += n_stats;
  double *d_taper = base; base += n_taper;
  p.taper = d_taper;
  double *d_mpow = base; base += n_mpow;
  double *d_resp = base; base += n_resp;
  FusedScratch fs = {};
  if (fused) {
    if (((uintptr_t)base & 15u) != 0) ++base;  // 16-byte alignment of the descriptor entries
    fs.ticket = (int *)base; base += n_flags;
    fs.desc = (DescEntry *)base; base += n_desc;
    fs.resp = d_resp;
    fs.ppow = d_resp + (size_t)D * PP_L;
  } else {
    p.chunk_state = base; base += n_chunk;
    p.tile_state = base; base += n_tile;
    p.tile_in = base; base += n_tile;
  }
  double *d_tmp = base;
  p.stats = d_stats;
  p.mpow = d_mpow;
  if (nsec > 0) {
    std::vector<double> M, pw(n_mpow, 0.0);
    chunk_transition(p.coef, nsec, PP_L, M);
    for (int i = 0; i < D; ++i) pw[(size_t)i * D + i] = 1.0;
    for (int k = 1; k <= PP_TPB; ++k) {  // M^k = M^(k-1) * M
      const double *a = &pw[(size_t)(k - 1) * D * D];
      double *c = &pw[(size_t)k * D * D];
      for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {
          double acc = 0.0;
          for (int l = 0; l < D; ++l) acc += a[i * D + l] * M[(size_t)l * D + j];
          c[i * D + j] = acc;
        }
    }
    std::vector<double> R;
    zero_input_response(p.coef, nsec, PP_L, R);
    pw.insert(pw.end(), R.begin(), R.end());  // d_resp follows d_mpow
    if (fused) {  // P^0 .. P^(ntiles-1), P = M^PP_TPB, for the look-back across tiles
      std::vector<double> P(pw.begin() + (size_t)PP_TPB * D * D, pw.begin() + (size_t)(PP_TPB + 1) * D * D);
      std::vector<double> cur((size_t)D * D, 0.0), nxt((size_t)D * D);
      for (int i = 0; i < D; ++i) cur[(size_t)i * D + i] = 1.0;
      for (int d = 0; d < p.ntiles; ++d) {
        pw.insert(pw.end(), cur.begin(), cur.end());
        for (int i = 0; i < D; ++i)
          for (int j = 0; j < D; ++j) {
            double acc = 0.0;
            for (int l = 0; l < D; ++l) acc += cur[(size_t)i * D + l] * P[(size_t)l * D + j];
            nxt[(size_t)i * D + j] = acc;
          }
        cur.swap(nxt);
      }
    }
    // pageable source: the copy is staged before the call returns
    QMB_CUDA(cudaMemcpyAsync(d_mpow, pw.data(), (n_mpow + n_resp) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    QMB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  if (kernel_ms) QMB_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  double *d_partial = d_stats + (size_t)2 * n_sel;
  if (detrend) {
    k_pre_partial<<<dim3((unsigned)p.ntiles, (unsigned)n_sel), PP_TPB, 0, ctx->stream>>>(p, d_partial);
    ctx->launches++;
    QMB_CUDA(cudaGetLastError());
  }
  k_pre_stats<<<n_sel, 32, 0, ctx->stream>>>(p, d_partial, d_stats, detrend ? 1 : 0);
  ctx->launches++;
  QMB_CUDA(cudaGetLastError());
  if (p.wlen > 0) {
    k_taper_table<<<(unsigned)((p.wlen + 255) / 256), 256, 0, ctx->stream>>>(d_taper, n, p.wlen);
    ctx->launches++;
    QMB_CUDA(cudaGetLastError());
  }
  p.apply_pre = 1;
  if (nsec == 0) {
    p.y = d_out; p.y_stride = out_stride; p.chan_out0 = out_chan0; p.decimate = decimate;
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)n_sel);
    k_pre_only<<<grid, 256, 0, ctx->stream>>>(p);
    ctx->launches++;
    QMB_CUDA(cudaGetLastError());
  } else if (!zerophase) {
    p.y = d_out; p.y_stride = out_stride; p.chan_out0 = out_chan0; p.decimate = decimate;
    QMB_TRY(fused ? run_pass_fused_n(ctx, nsec, p, fs, false) : run_pass_n(ctx, nsec, p, false));
  } else {
    p.y = d_tmp; p.y_stride = n; p.chan_out0 = 0; p.decimate = 1;
    QMB_TRY(fused ? run_pass_fused_n(ctx, nsec, p, fs, false) : run_pass_n(ctx, nsec, p, false));
    p.x = d_tmp; p.x_stride = n; p.chan_in = nullptr; p.apply_pre = 0;
    p.y = d_out; p.y_stride = out_stride; p.chan_out0 = out_chan0; p.decimate = decimate;
    QMB_TRY(fused ? run_pass_fused_n(ctx, nsec, p, fs, true) : run_pass_n(ctx, nsec, p, true));
  }
  if (kernel_ms) {
    QMB_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    QMB_CUDA(cudaEventSynchronize(ctx->ev[1]));
    QMB_CUDA(cudaEventElapsedTime(kernel_ms, ctx->ev[0], ctx->ev[1]));
  }
  return QMB_OK;
}

// ---------------------------------------------------------------------------------------------
// device-resident traces
// ---------------------------------------------------------------------------------------------
extern "C" int qmb_waveforms_free(qmb_waveforms_t w) {
  if (!w) return QMB_OK;
  cudaSetDevice(w->device);
  w->data.release();
  delete w;
  return QMB_OK;
}

extern "C" int qmb_waveforms_alloc(qmb_ctx_t ctx, int32_t n_channels, int64_t n, qmb_waveforms_t *out) {
  if (!ctx || !out) QMB_FAIL(QMB_ERR_ARG, "null argument");
  *out = nullptr;
  if (n_channels <= 0 || n <= 0) QMB_FAIL(QMB_ERR_ARG, "empty trace block");
  QMB_CUDA(cudaSetDevice(ctx->device));
  qmb_waveforms *w = new qmb_waveforms();
  w->ctx = ctx; w->device = ctx->device; w->C = n_channels; w->n = n;
  int rc = w->data.ensure((size_t)n_channels * n);
  if (rc == QMB_OK) {
    cudaError_t e = cudaMemsetAsync(w->data.p, 0, w->data.bytes(), ctx->stream);
    if (e != cudaSuccess) {
      qmb_set_error("cudaMemsetAsync: %s", cudaGetErrorString(e));
      rc = QMB_ERR_CUDA;
    }
  }
  if (rc != QMB_OK) {
    qmb_waveforms_free(w);
    return rc;
  }
  *out = w;
  return QMB_OK;
}

extern "C" int qmb_waveforms_upload(qmb_ctx_t ctx, const double *x, int32_t n_channels, int64_t n,
                                    qmb_waveforms_t *out) {
  if (!x) QMB_FAIL(QMB_ERR_ARG, "null argument");
  QMB_TRY(qmb_waveforms_alloc(ctx, n_channels, n, out));
  cudaError_t e = cudaMemcpyAsync((*out)->data.p, x, (size_t)n_channels * n * sizeof(double),
                                  cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    qmb_waveforms_free(*out);
    *out = nullptr;
    QMB_FAIL(QMB_ERR_CUDA, "H2D copy of the traces: %s", cudaGetErrorString(e));
  }
  return QMB_OK;
}

extern "C" int qmb_waveforms_fetch(qmb_waveforms_t w, int32_t chan0, int32_t n_chan, double *host) {
  if (!w || !host) QMB_FAIL(QMB_ERR_ARG, "null argument");
  if (chan0 < 0 || n_chan < 0 || chan0 + n_chan > w->C) QMB_FAIL(QMB_ERR_ARG, "channel range out of bounds");
  if (n_chan == 0) return QMB_OK;
  QMB_CUDA(cudaSetDevice(w->device));
  QMB_CUDA(cudaMemcpyAsync(host, w->data.p + (size_t)chan0 * w->n, (size_t)n_chan * w->n * sizeof(double),
                           cudaMemcpyDeviceToHost, w->ctx->stream));
  QMB_CUDA(cudaStreamSynchronize(w->ctx->stream));
  return QMB_OK;
}

extern "C" int qmb_waveforms_filter(qmb_ctx_t ctx, qmb_waveforms_t in, const int32_t *channels, int32_t n_sel,
                                    const double *sos, int32_t n_sections, int32_t detrend,
                                    double taper_fraction, int32_t zerophase, int32_t decimate,
                                    qmb_waveforms_t out, int32_t out_chan0, float *kernel_ms) {
  if (!ctx || !in || !out) QMB_FAIL(QMB_ERR_ARG, "null argument");
  if (in->ctx != ctx || out->ctx != ctx) QMB_FAIL(QMB_ERR_ARG, "handles belong to a different context");
  if (in == out) QMB_FAIL(QMB_ERR_ARG, "in-place filtering is not supported");
  if (!channels) n_sel = in->C;
  if (n_sel <= 0) QMB_FAIL(QMB_ERR_ARG, "no channels selected");
  if (decimate < 1) QMB_FAIL(QMB_ERR_ARG, "decimation factor must be >= 1");
  const int64_t n_out = (in->n + decimate - 1) / decimate;
  if (out->n != n_out) QMB_FAIL(QMB_ERR_ARG, "output block has %lld samples per trace, need %lld", (long long)out->n, (long long)n_out);
  if (out_chan0 < 0 || out_chan0 + n_sel > out->C) QMB_FAIL(QMB_ERR_ARG, "output channel range out of bounds");
  QMB_CUDA(cudaSetDevice(ctx->device));
  int *d_sel = nullptr;
  if (channels) {
    for (int k = 0; k < n_sel; ++k)
      if (channels[k] < 0 || channels[k] >= in->C) QMB_FAIL(QMB_ERR_ARG, "channel index out of range");
    QMB_TRY(ctx->pp_sel.ensure((size_t)n_sel));
    d_sel = ctx->pp_sel.p;
    QMB_CUDA(cudaMemcpyAsync(d_sel, channels, (size_t)n_sel * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    QMB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return qmb_preprocess_device(ctx, in->data.p, in->n, d_sel, n_sel, in->n, sos, n_sections, detrend,
                               taper_fraction, zerophase, decimate, out->data.p, out->n, out_chan0, kernel_ms);
}

// util.upsample's interpolation (quakemigrate/util.py:544-551): fence-post linear interpolation,
// out[i * up + r] = r/up * x[i+1] + (up-r)/up * x[i]; products and sum rounded separately like numpy
namespace {
__global__ void __launch_bounds__(256) k_upsample(const double *x, long long n, int up, double *y, long long m) {
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= m) return;
  const double *xs = x + (size_t)blockIdx.y * n;
  const long long i = o / up;
  const int r = (int)(o % up);
  double v = xs[i];
  if (r) v = __dadd_rn(__dmul_rn((double)r / (double)up, xs[i + 1]), __dmul_rn((double)(up - r) / (double)up, v));
  y[(size_t)blockIdx.y * m + o] = v;
}
}  // namespace

extern "C" int qmb_waveforms_upsample(qmb_ctx_t ctx, qmb_waveforms_t in, int32_t upfactor, qmb_waveforms_t *out) {
  if (!ctx || !in || !out) QMB_FAIL(QMB_ERR_ARG, "null argument");
  *out = nullptr;
  if (in->ctx != ctx) QMB_FAIL(QMB_ERR_ARG, "handles belong to a different context");
  if (upfactor < 1) QMB_FAIL(QMB_ERR_ARG, "upfactor must be >= 1");
  QMB_CUDA(cudaSetDevice(ctx->device));
  const int64_t m = (in->n - 1) * upfactor + 1;
  QMB_TRY(qmb_waveforms_alloc(ctx, in->C, m, out));
  dim3 grid((unsigned)((m + 255) / 256), (unsigned)in->C);
  k_upsample<<<grid, 256, 0, ctx->stream>>>(in->data.p, in->n, upfactor, (*out)->data.p, m);
  ctx->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    qmb_waveforms_free(*out);
    *out = nullptr;
    QMB_FAIL(QMB_ERR_CUDA, "qmb_waveforms_upsample: %s", cudaGetErrorString(e));
  }
  return QMB_OK;
}

// one call, host in / host out
extern "C" int qmb_preprocess(qmb_ctx_t ctx, const double *x, int32_t n_channels, int64_t n, const double *sos,
                              int32_t n_sections, int32_t detrend, double taper_fraction, int32_t zerophase,
                              int32_t decimate, double *out) {
  if (!ctx || !x || !out) QMB_FAIL(QMB_ERR_ARG, "null argument");
  if (decimate < 1) QMB_FAIL(QMB_ERR_ARG, "decimation factor must be >= 1");
  qmb_waveforms_t in = nullptr, res = nullptr;
  QMB_TRY(qmb_waveforms_upload(ctx, x, n_channels, n, &in));
  int rc = qmb_waveforms_alloc(ctx, n_channels, (n + decimate - 1) / decimate, &res);
  if (rc == QMB_OK)
    rc = qmb_waveforms_filter(ctx, in, nullptr, n_channels, sos, n_sections, detrend, taper_fraction, zerophase,
                              decimate, res, 0, nullptr);
  if (rc == QMB_OK) rc = qmb_waveforms_fetch(res, 0, n_channels, out);
  qmb_waveforms_free(in);
  qmb_waveforms_free(res);
  return rc;
}
