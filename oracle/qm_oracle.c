/*
 * qm_oracle.c -- CPU ORACLE for the QuakeMigrate coalescence hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (quakemigrate_b200 + libqmb200.so) never links, imports or calls it.
 *
 * This is an independent restatement (plain C99, IEEE double, no fast-math)
 * of the arithmetic performed by the reference C extension:
 *
 *   oracle_stack        <- migrate          quakemigrate/core/src/migratelib.c:40-65
 *   oracle_peak_trace   <- find_max_coa     quakemigrate/core/src/migratelib.c:85-111
 *   oracle_stalta_*     <- *_sta_lta        quakemigrate/core/src/onsetlib.c:35-59,
 *                                           :79-108, :126-148
 *
 * The loop nests are organised differently from the reference (gather per
 * output element; node-outer running reduction) but the ORDER of floating
 * point operations applied to every individual output element is the same,
 * so on IEEE hardware without value-changing optimisations the results are
 * bit-identical to a strict (-O2, no -ffast-math) build of the reference and
 * agree with the reference's own -Ofast build to a few ulp (exp / reciprocal).
 *
 * Parity pin: tests/test_oracle.py checks this file against
 *   (1) the reference KAT vectors of tests/test_onsets.py:23-39,
 *   (2) oracle/_ref/qmlib.so = the reference sources compiled with the
 *       reference flags (setup.py:111-119) where /root/reference exists,
 *   (3) golden vectors generated from (2) and committed in tests/golden/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* Must match qmlib.h:34-38 and the numpy mirror in core/lib.py:33-36. */
typedef struct {
  int n;
  int nsta;
  int nlta;
} oracle_stalta_header;

/*
 * Migrate-and-stack.  For node n and output sample t:
 *
 *   v  = map4d[n, t]                         (caller passes zeros, lib.py:101)
 *   v += onsets[s, max(0, tau[n, s]) + fsmp + t]   for s = 0 .. S-1, in order
 *   map4d[n, t] = exp(v / available)
 *
 * onsets rows have length fsmp + n_samples + lsmp (migratelib.c:57); the
 * travel-time table is (n_nodes, n_stations) with the station axis fastest
 * (migratelib.c:53).  Negative travel times clamp to 0 (migratelib.c:55);
 * there is no upper clamp in the reference and none here.
 */
void oracle_stack(const double *onsets, const int32_t *tau, double *map4d,
                  int32_t fsmp, int32_t lsmp, int32_t n_samples,
                  int32_t n_stations, int32_t available, int64_t n_nodes,
                  int64_t threads) {
  const int64_t row = (int64_t)fsmp + (int64_t)lsmp + (int64_t)n_samples;
  if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads((int)threads)
  for (int64_t n = 0; n < n_nodes; ++n) {
    const int32_t *tt = tau + n * (int64_t)n_stations;
    double *out = map4d + n * (int64_t)n_samples;
    for (int32_t t = 0; t < n_samples; ++t) {
      double v = out[t];
      for (int32_t s = 0; s < n_stations; ++s) {
        const int32_t shift = tt[s] > 0 ? tt[s] : 0;
        v += onsets[(int64_t)s * row + shift + fsmp + t];
      }
      out[t] = exp(v / (double)available);
    }
  }
}

/*
 * Per-sample peak of the 4-D map: value, FIRST flat node index attaining it
 * (strict ">" update, seeded from node 0: migratelib.c:96-105) and the
 * normalised peak  max * n_nodes / sum_n map4d[n, t]  (migratelib.c:108).
 * The node sum is accumulated in ascending node order, like the reference.
 *
 * Organisation differs from the reference (which walks one strided column per
 * sample): time is cut in blocks that fit L1/L2 and nodes are streamed once
 * per block, so each (t) sees nodes in the same ascending order.
 */
void oracle_peak_trace(const double *map4d, double *max_coa,
                       double *max_norm_coa, int64_t *max_coa_idx,
                       int32_t n_samples, int64_t n_nodes, int64_t threads) {
  enum { TBLK = 512 };
  if (threads < 1) threads = 1;
  const int32_t n_blocks = (n_samples + TBLK - 1) / TBLK;
#pragma omp parallel for schedule(dynamic, 1) num_threads((int)threads)
  for (int32_t b = 0; b < n_blocks; ++b) {
    const int32_t t0 = b * TBLK;
    const int32_t len = (n_samples - t0 < TBLK) ? n_samples - t0 : TBLK;
    double best[TBLK], total[TBLK];
    int64_t where[TBLK];
    for (int32_t i = 0; i < len; ++i) {
      best[i] = map4d[t0 + i];
      total[i] = map4d[t0 + i];
      where[i] = 0;
    }
    for (int64_t n = 1; n < n_nodes; ++n) {
      const double *rowp = map4d + n * (int64_t)n_samples + t0;
      for (int32_t i = 0; i < len; ++i) {
        const double c = rowp[i];
        total[i] += c;
        if (c > best[i]) {
          best[i] = c;
          where[i] = n;
        }
      }
    }
    for (int32_t i = 0; i < len; ++i) {
      max_coa[t0 + i] = best[i];
      max_norm_coa[t0 + i] = best[i] * (double)n_nodes / total[i];
      max_coa_idx[t0 + i] = where[i];
    }
  }
}

/*
 * Classic ("overlapping") STA/LTA: both windows END at sample i.
 * Valid outputs: i in [nlta-1, n).  Samples before keep the caller's pre-fill
 * (1.0, lib.py:204).  Running sums are updated as  sum += (new - old)  which
 * is the association the reference uses (onsetlib.c:46-57); this matters for
 * bit-exactness of the drift.
 */
void oracle_stalta_overlapping(const double *x, const oracle_stalta_header *h,
                               double *onset) {
  const int n = h->n, ns = h->nsta, nl = h->nlta;
  const double scale = (double)nl / (double)ns;
  double s_short = 0.0, s_long;
  int i = 0;
  while (i < ns) s_short += x[i++];
  s_long = s_short;
  while (i < nl) {
    const double in = x[i];
    s_long += in;
    s_short += in - x[i - ns];
    ++i;
  }
  onset[nl - 1] = s_short / s_long * scale;
  for (i = nl; i < n; ++i) {
    const double in = x[i];
    s_short += in - x[i - ns];
    s_long += in - x[i - nl];
    onset[i] = s_short / s_long * scale;
  }
}

/*
 * Centred STA/LTA: the LTA window ends AT sample i, the STA window starts
 * just after it.  Valid outputs: i in [nlta-1, n-nsta).  For i >= nlta a
 * non-positive LTA yields exactly 1.0 (onsetlib.c:102-106); the first valid
 * sample has no such guard (onsetlib.c:97).
 */
void oracle_stalta_centred(const double *x, const oracle_stalta_header *h,
                           double *onset) {
  const int n = h->n, ns = h->nsta, nl = h->nlta;
  const double scale = (double)nl / (double)ns;
  double s_long = 0.0, s_short = 0.0;
  for (int i = 0; i < nl; ++i) s_long += x[i];
  for (int i = nl; i < nl + ns; ++i) s_short += x[i];
  onset[nl - 1] = s_short / s_long * scale;
  for (int i = nl; i < n - ns; ++i) {
    s_short += x[i + ns] - x[i];
    s_long += x[i] - x[i - nl];
    onset[i] = (s_long > 0.0) ? s_short / s_long * scale : 1.0;
  }
}

/*
 * Recursive (exponential moving average) STA/LTA, onsetlib.c:126-148.
 * Sample 0 keeps the caller's pre-fill unless the nulling below overwrites it;
 * the first nlta samples are forced to 1.0 only when nlta < n.
 */
void oracle_stalta_recursive(const double *x, const oracle_stalta_header *h,
                             double *onset) {
  const int n = h->n;
  const double a = 1.0 / (double)h->nsta, b = 1.0 / (double)h->nlta;
  double s_short = 0.0, s_long = 0.0;
  for (int i = 1; i < n; ++i) {
    const double in = x[i];
    s_short = a * in + (1 - a) * s_short;
    s_long = b * in + (1 - b) * s_long;
    onset[i] = s_short / s_long;
  }
  if (h->nlta < n)
    for (int i = 0; i < h->nlta; ++i) onset[i] = 1.0;
}

/*
 * Fused detect oracle: same arithmetic as oracle_stack followed by
 * oracle_peak_trace, but without materialising the (N, T) volume, so that the
 * full-size BASELINE configs can be checked on a host with ordinary RAM.
 * Per output sample the node order is ascending and the per-node stack order
 * is station-ascending, so results are bit-identical to the two-step path.
 * Time is processed in blocks; each thread owns whole blocks.
 */
void oracle_detect_fused(const double *onsets, const int32_t *tau, int32_t fsmp,
                         int32_t lsmp, int32_t n_samples, int32_t n_stations,
                         int32_t available, int64_t n_nodes, double *max_coa,
                         double *max_norm_coa, int64_t *max_coa_idx,
                         int64_t threads) {
  enum { TBLK = 256 };
  const int64_t row = (int64_t)fsmp + (int64_t)lsmp + (int64_t)n_samples;
  if (threads < 1) threads = 1;
  const int32_t n_blocks = (n_samples + TBLK - 1) / TBLK;
#pragma omp parallel for schedule(dynamic, 1) num_threads((int)threads)
  for (int32_t b = 0; b < n_blocks; ++b) {
    const int32_t t0 = b * TBLK;
    const int32_t len = (n_samples - t0 < TBLK) ? n_samples - t0 : TBLK;
    double best[TBLK], total[TBLK], v[TBLK];
    int64_t where[TBLK];
    for (int64_t n = 0; n < n_nodes; ++n) {
      const int32_t *tt = tau + n * (int64_t)n_stations;
      for (int32_t i = 0; i < len; ++i) v[i] = 0.0;
      for (int32_t s = 0; s < n_stations; ++s) {
        const int32_t shift = tt[s] > 0 ? tt[s] : 0;
        const double *src = onsets + (int64_t)s * row + shift + fsmp + t0;
        for (int32_t i = 0; i < len; ++i) v[i] += src[i];
      }
      for (int32_t i = 0; i < len; ++i) {
        const double c = exp(v[i] / (double)available);
        if (n == 0) {
          best[i] = c;
          total[i] = c;
          where[i] = 0;
        } else {
          total[i] += c;
          if (c > best[i]) {
            best[i] = c;
            where[i] = n;
          }
        }
      }
    }
    for (int32_t i = 0; i < len; ++i) {
      max_coa[t0 + i] = best[i];
      max_norm_coa[t0 + i] = best[i] * (double)n_nodes / total[i];
      max_coa_idx[t0 + i] = where[i];
    }
  }
}
