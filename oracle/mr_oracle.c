/*
 * oracle/mr_oracle.c — CPU restatement of the MusicRecommendation scoring hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (musicrecommendation_b200/, the
 * C-ABI library libmrscore.so) may include, link, call or execute this file.  It is used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the
 * checker and as the timed CPU baseline.
 *
 * PARITY STATUS: "parity unpinned".  The reference (Scala 2.12, /root/reference) ships no tests,
 * no golden vectors and no datasets and there is no JVM in this image, so
 * the restatement cannot be checked against outputs of the reference itself.  It is pinned by
 *   (1) the hand-derived known-answer fixture of SURVEY.md §4.3 (tests/golden/fixture_4_3.json),
 *   (2) a differential test between the two independent restatements in this file:
 *       mro_naive_*  — loop-for-loop transliteration of MusicRecommender.scala as written
 *       mro_canon_*  — CSR / inverted-index restatement in exact integer arithmetic.
 *
 * Reference citations are file:line into /root/reference/src/main/scala/ :
 *   MR   = music_recommandation/MusicRecommender.scala
 *   MAIN = main.scala, DIST = distributed.scala, UTIL = my_utils/MyUtils.scala
 *
 * Data model handed to every function (what MR:26-62 builds, with strings replaced by dense
 * int32 ids assigned in ascending String.compareTo order):
 *   train CSR  tr_ptr[T+1], tr_col[]   songs of train user v, ascending, unique   (MR:55)
 *   test  CSR  te_ptr[U+1], te_col[]   visible songs of test user u               (MR:56)
 *   deg_tr[v] = trainUsersToSongsMap(v).length,  deg_te[u] = testUsersToSongsMap(u).length (MR:147)
 *   deg_song[s] = songsToUsersMap(s).length — listeners in train AND test-visible (MR:41,53,237)
 *
 * Canonical exact arithmetic (what the GPU path must reproduce bit for bit):
 *   q_k(x)    = llrint(2^k / sqrt((double)x))             (0 when x == 0)
 *   UBM  Sint[u,s] = sum_{v in train, s in I_v} |I_u ∩ I_v| * q_24(deg_tr[v])          (int64, exact)
 *        score     = (double)Sint * (2^-24 / sqrt((double)deg_te[u]))
 *   IBM  Sint[u,s] = sum_{j in I_u, j != s} |U_s^train ∩ U_j^train| * q_26(deg_song[j]) (int64, exact)
 *        score     = (double)Sint * (2^-26 / sqrt((double)deg_song[s]))
 *   (worst-case relative quantisation error 0.5*sqrt(deg)/2^k: 2e-6 for UBM at |I_v| = 4400 (MSD maximum), 7.6e-6 at 65535;
 *    2.5e-6 for IBM at deg 110k)
 * Integer sums are associative, so any summation order / sharding gives the same bits; the
 * result is within ~1e-8 relative of the reference's fp64 expression c/(sqrt(a)*sqrt(b)) summed
 * left to right (tolerance in north_star: 1e-5).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MRO_API __attribute__((visibility("default")))

typedef struct {
  int32_t T, U, S;
  const int64_t *tr_ptr; const int32_t *tr_col;
  const int64_t *te_ptr; const int32_t *te_col;
  const int32_t *deg_tr, *deg_te, *deg_song;
} mro_data;

static const double QSCALE_UBM = 16777216.0;         /* 2^24: weighted co-occurrence rows fit u32 entries (plus a small exception list) */
static const double QSCALE_IBM = 67108864.0;         /* 2^26: sums of up to ~90 terms fit a uint32 panel entry */

/* ------------------------------------------------------------------------------------------ */
/* small helpers                                                                               */
/* ------------------------------------------------------------------------------------------ */

/* Array.contains — linear scan, as the reference does (MR:109, 144, 162, 234, 253). */
static int lin_contains(const int32_t *a, int64_t n, int32_t x) {
  for (int64_t i = 0; i < n; ++i) if (a[i] == x) return 1;
  return 0;
}

/* model: 0 = UBM scale, 1 = IBM scale */
MRO_API int64_t mro_q(int32_t deg, int model) {
  if (deg <= 0) return 0;
  return llrint((model == 0 ? QSCALE_UBM : QSCALE_IBM) / sqrt((double)deg));
}
MRO_API double mro_rs(int32_t deg, int model) {
  if (deg <= 0) return 0.0;
  return (1.0 / (model == 0 ? QSCALE_UBM : QSCALE_IBM)) / sqrt((double)deg);
}

MRO_API int mro_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1; the timed CPU legs of bench.py set the thread count explicitly. */
MRO_API void mro_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* song -> listeners map over train AND test-visible users (MR:41, 53, 60-62); train users are
 * numbered 0..T-1, test users T..T+U-1.  Returned arrays are malloc'ed. */
static void build_song_to_users(const mro_data *d, int64_t **ptr_out, int32_t **usr_out) {
  int32_t S = d->S;
  int64_t *ptr = (int64_t *)calloc((size_t)S + 1, sizeof(int64_t));
  for (int64_t i = 0; i < d->tr_ptr[d->T]; ++i) ptr[d->tr_col[i] + 1]++;
  for (int64_t i = 0; i < d->te_ptr[d->U]; ++i) ptr[d->te_col[i] + 1]++;
  for (int32_t s = 0; s < S; ++s) ptr[s + 1] += ptr[s];
  int32_t *usr = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ptr[S] > 0 ? ptr[S] : 1));
  int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (size_t)(S + 1));
  memcpy(fill, ptr, sizeof(int64_t) * (size_t)(S + 1));
  for (int32_t v = 0; v < d->T; ++v)
    for (int64_t i = d->tr_ptr[v]; i < d->tr_ptr[v + 1]; ++i) usr[fill[d->tr_col[i]]++] = v;
  for (int32_t u = 0; u < d->U; ++u)
    for (int64_t i = d->te_ptr[u]; i < d->te_ptr[u + 1]; ++i) usr[fill[d->te_col[i]]++] = d->T + u;
  free(fill);
  *ptr_out = ptr; *usr_out = usr;
}

/* train-only song -> listeners (CSC of the train matrix). */
static void build_train_csc(const mro_data *d, int64_t **ptr_out, int32_t **usr_out) {
  int32_t S = d->S;
  int64_t *ptr = (int64_t *)calloc((size_t)S + 1, sizeof(int64_t));
  for (int64_t i = 0; i < d->tr_ptr[d->T]; ++i) ptr[d->tr_col[i] + 1]++;
  for (int32_t s = 0; s < S; ++s) ptr[s + 1] += ptr[s];
  int32_t *usr = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ptr[S] > 0 ? ptr[S] : 1));
  int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (size_t)(S + 1));
  memcpy(fill, ptr, sizeof(int64_t) * (size_t)(S + 1));
  for (int32_t v = 0; v < d->T; ++v)
    for (int64_t i = d->tr_ptr[v]; i < d->tr_ptr[v + 1]; ++i) usr[fill[d->tr_col[i]]++] = v;
  free(fill);
  *ptr_out = ptr; *usr_out = usr;
}

/* ------------------------------------------------------------------------------------------ */
/* NAIVE restatement — MusicRecommender.scala as written (small inputs only)                   */
/* ------------------------------------------------------------------------------------------ */

/* UBM cosineSimilarity, MR:140-149: numerator = songs.map(both contain ? 1 : 0).sum (MR:142-145),
 * denominator = sqrt(|I_u|) * sqrt(|I_v|) (MR:147), numerator / denominator else 0.0 (MR:148). */
static double naive_ubm_cos(const mro_data *d, int32_t u, int32_t v) {
  const int32_t *Iu = d->te_col + d->te_ptr[u]; int64_t nu = d->te_ptr[u + 1] - d->te_ptr[u];
  const int32_t *Iv = d->tr_col + d->tr_ptr[v]; int64_t nv = d->tr_ptr[v + 1] - d->tr_ptr[v];
  int numerator = 0;
  for (int32_t song = 0; song < d->S; ++song)
    numerator += (lin_contains(Iu, nu, song) && lin_contains(Iv, nv, song)) ? 1 : 0;
  double denominator = sqrt((double)d->deg_te[u]) * sqrt((double)d->deg_tr[v]);
  return denominator != 0 ? numerator / denominator : 0.0;
}

/* UBM rank, MR:159-166: for u2 <- trainUsers if I_u2 contains song yield cos(user, u2); .sum */
static double naive_ubm_rank(const mro_data *d, int32_t u, int32_t s) {
  double sum = 0.0;
  for (int32_t v = 0; v < d->T; ++v) {
    const int32_t *Iv = d->tr_col + d->tr_ptr[v]; int64_t nv = d->tr_ptr[v + 1] - d->tr_ptr[v];
    if (lin_contains(Iv, nv, s)) sum += naive_ubm_cos(d, u, v);
  }
  return sum;
}

/* getModel, MR:105-111 (song-major loop, listened pairs skipped).  Output is dense row-major
 * [U][S] with NaN at the pairs the reference does not emit, i.e. already in the (user, song)
 * order the alignment sort MAIN:57-59 produces.  par != 0 mirrors getModelP MR:119-125. */
MRO_API void mro_naive_ubm(const mro_data *d, double *out, int par) {
  int64_t S = d->S, U = d->U;
#pragma omp parallel for collapse(2) schedule(dynamic, 8) if (par)
  for (int64_t s = 0; s < S; ++s)
    for (int64_t u = 0; u < U; ++u) {
      const int32_t *Iu = d->te_col + d->te_ptr[u]; int64_t nu = d->te_ptr[u + 1] - d->te_ptr[u];
      out[u * S + s] = lin_contains(Iu, nu, (int32_t)s) ? NAN : naive_ubm_rank(d, (int32_t)u, (int32_t)s);
    }
}

/* IBM cosineSimilarity, MR:230-239: numerator iterates trainUsers only (MR:232) but the listener
 * lists — and therefore the denominator lengths (MR:237) — include test-visible listeners. */
static double naive_ibm_cos(const mro_data *d, const int64_t *sp, const int32_t *su, int32_t s1, int32_t s2) {
  const int32_t *U1 = su + sp[s1]; int64_t n1 = sp[s1 + 1] - sp[s1];
  const int32_t *U2 = su + sp[s2]; int64_t n2 = sp[s2 + 1] - sp[s2];
  int numerator = 0;
  for (int32_t user = 0; user < d->T; ++user)
    numerator += (lin_contains(U1, n1, user) && lin_contains(U2, n2, user)) ? 1 : 0;
  double denominator = sqrt((double)d->deg_song[s1]) * sqrt((double)d->deg_song[s2]);
  return denominator != 0 ? numerator / denominator : 0;
}

/* IBM rank, MR:249-257: for s2 <- songs if s2 != song if I_u contains s2 yield cos(song, s2); .sum */
static double naive_ibm_rank(const mro_data *d, const int64_t *sp, const int32_t *su, int32_t u, int32_t s) {
  const int32_t *Iu = d->te_col + d->te_ptr[u]; int64_t nu = d->te_ptr[u + 1] - d->te_ptr[u];
  double sum = 0.0;
  for (int32_t s2 = 0; s2 < d->S; ++s2)
    if (s2 != s && lin_contains(Iu, nu, s2)) sum += naive_ibm_cos(d, sp, su, s, s2);
  return sum;
}

MRO_API void mro_naive_ibm(const mro_data *d, double *out, int par) {
  int64_t S = d->S, U = d->U;
  int64_t *sp; int32_t *su;
  build_song_to_users(d, &sp, &su);
#pragma omp parallel for collapse(2) schedule(dynamic, 8) if (par)
  for (int64_t s = 0; s < S; ++s)
    for (int64_t u = 0; u < U; ++u) {
      const int32_t *Iu = d->te_col + d->te_ptr[u]; int64_t nu = d->te_ptr[u + 1] - d->te_ptr[u];
      out[u * S + s] = lin_contains(Iu, nu, (int32_t)s) ? NAN : naive_ibm_rank(d, sp, su, (int32_t)u, (int32_t)s);
    }
  free(sp); free(su);
}

/* Bounded-sample variant for the timed CPU baseline: scores only songs s with s % stride == phase
 * for test users [u0,u1); returns the number of scored pairs.  Same loops, fewer of them. */
MRO_API int64_t mro_naive_sample(const mro_data *d, int model, int32_t u0, int32_t u1, int32_t stride,
                                 int32_t phase, int par, double *checksum) {
  int64_t S = d->S; int64_t *sp = NULL; int32_t *su = NULL;
  if (model == 1) build_song_to_users(d, &sp, &su);
  int64_t pairs = 0; double acc = 0.0;
#pragma omp parallel for collapse(2) schedule(dynamic, 4) reduction(+ : pairs, acc) if (par)
  for (int64_t s = phase; s < S; s += stride)
    for (int64_t u = u0; u < u1; ++u) {
      const int32_t *Iu = d->te_col + d->te_ptr[u]; int64_t nu = d->te_ptr[u + 1] - d->te_ptr[u];
      if (lin_contains(Iu, nu, (int32_t)s)) continue;
      acc += model == 0 ? naive_ubm_rank(d, (int32_t)u, (int32_t)s) : naive_ibm_rank(d, sp, su, (int32_t)u, (int32_t)s);
      pairs++;
    }
  if (sp) { free(sp); free(su); }
  if (checksum) *checksum = acc;
  return pairs;
}

/* ------------------------------------------------------------------------------------------ */
/* CANONICAL restatement — CSR / inverted index, exact integer accumulation                    */
/* ------------------------------------------------------------------------------------------ */

/* |I_u ∩ I_v| for all (u, v): out[U][T] int32 (MR:142-145). */
MRO_API void mro_counts_ubm(const mro_data *d, int32_t *out) {
  int64_t *cp; int32_t *cu; build_train_csc(d, &cp, &cu);
  memset(out, 0, sizeof(int32_t) * (size_t)d->U * (size_t)d->T);
#pragma omp parallel for schedule(dynamic, 4)
  for (int32_t u = 0; u < d->U; ++u) {
    int32_t *row = out + (int64_t)u * d->T;
    for (int64_t i = d->te_ptr[u]; i < d->te_ptr[u + 1]; ++i) {
      int32_t j = d->te_col[i];
      for (int64_t k = cp[j]; k < cp[j + 1]; ++k) row[cu[k]]++;
    }
  }
  free(cp); free(cu);
}

/* G[rows[r], :] = |U_i^train ∩ U_j^train| for the listed songs i: out[n][S] int32 (MR:232-235). */
MRO_API void mro_gram_rows(const mro_data *d, const int32_t *rows, int32_t n, int32_t *out) {
  int64_t *cp; int32_t *cu; build_train_csc(d, &cp, &cu);
  memset(out, 0, sizeof(int32_t) * (size_t)n * (size_t)d->S);
#pragma omp parallel for schedule(dynamic, 4)
  for (int32_t r = 0; r < n; ++r) {
    int32_t i = rows[r]; int32_t *row = out + (int64_t)r * d->S;
    for (int64_t k = cp[i]; k < cp[i + 1]; ++k) {
      int32_t v = cu[k];
      for (int64_t m = d->tr_ptr[v]; m < d->tr_ptr[v + 1]; ++m) row[d->tr_col[m]]++;
    }
  }
  free(cp); free(cu);
}

/* Exact fixed-point score numerators for test users [u0,u1): out[(u-u0)][S] int64.
 * model 0 = UBM (MR:140-166), 1 = IBM (MR:230-257); listened pairs are computed too (the caller
 * masks them) except that IBM always excludes j == s (MR:252). */
MRO_API void mro_canon_sint(const mro_data *d, int model, int32_t u0, int32_t u1, int64_t *out) {
  int64_t *cp; int32_t *cu; build_train_csc(d, &cp, &cu);
  int64_t S = d->S;
#pragma omp parallel
  {
    int64_t *w = (int64_t *)calloc((size_t)(d->T > 0 ? d->T : 1), sizeof(int64_t));
#pragma omp for schedule(dynamic, 1)
    for (int32_t u = u0; u < u1; ++u) {
      int64_t *row = out + (int64_t)(u - u0) * S;
      memset(row, 0, sizeof(int64_t) * (size_t)S);
      if (model == 0) {
        /* w[v] = |I_u ∩ I_v| ; Sint[s] = sum_{v: s in I_v} w[v] * q(deg_tr[v]) */
        for (int64_t i = d->te_ptr[u]; i < d->te_ptr[u + 1]; ++i) {
          int32_t j = d->te_col[i];
          for (int64_t k = cp[j]; k < cp[j + 1]; ++k) w[cu[k]]++;
        }
        for (int32_t v = 0; v < d->T; ++v) if (w[v]) {
          int64_t term = w[v] * mro_q(d->deg_tr[v], 0);
          for (int64_t m = d->tr_ptr[v]; m < d->tr_ptr[v + 1]; ++m) row[d->tr_col[m]] += term;
          w[v] = 0;
        }
      } else {
        /* Sint[s] = sum_{j in I_u, j != s} G[s,j] * q(deg_song[j]);  G[s,j] = #{v: s,j in I_v} */
        for (int64_t i = d->te_ptr[u]; i < d->te_ptr[u + 1]; ++i) {
          int32_t j = d->te_col[i]; int64_t qj = mro_q(d->deg_song[j], 1);
          for (int64_t k = cp[j]; k < cp[j + 1]; ++k) {
            int32_t v = cu[k];
            for (int64_t m = d->tr_ptr[v]; m < d->tr_ptr[v + 1]; ++m) {
              int32_t s = d->tr_col[m];
              if (s != j) row[s] += qj;
            }
          }
        }
      }
    }
    free(w);
  }
  free(cp); free(cu);
}

/* Dense fp64 scores for users [u0,u1): out[(u-u0)][S], NaN where the user listened (MR:109). */
MRO_API void mro_canon_scores(const mro_data *d, int model, int32_t u0, int32_t u1, double *out) {
  int64_t S = d->S; int64_t n = (int64_t)(u1 - u0) * S;
  int64_t *sint = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  mro_canon_sint(d, model, u0, u1, sint);
#pragma omp parallel for schedule(static)
  for (int32_t u = u0; u < u1; ++u) {
    const int64_t *si = sint + (int64_t)(u - u0) * S; double *row = out + (int64_t)(u - u0) * S;
    double ru = mro_rs(d->deg_te[u], 0);
    for (int64_t s = 0; s < S; ++s)
      row[s] = model == 0 ? (double)si[s] * ru : (double)si[s] * mro_rs(d->deg_song[s], 1);
    for (int64_t i = d->te_ptr[u]; i < d->te_ptr[u + 1]; ++i) row[d->te_col[i]] = NAN;
  }
  free(sint);
}

/* ------------------------------------------------------------------------------------------ */
/* Blends — MR:317-481 on the (user, song)-sorted compact arrays of MAIN:57-59                 */
/* ------------------------------------------------------------------------------------------ */

/* java.util.Random (scala.util.Random wraps it, MR:439): 48-bit LCG. */
typedef struct { uint64_t s; } jrandom;
static void jr_seed(jrandom *r, uint64_t seed) { r->s = (seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1); }
static float jr_next_float(jrandom *r) {
  r->s = (r->s * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
  return (float)(int32_t)(r->s >> 24) / (float)(1 << 24);
}

enum { MRO_LC = 0, MRO_AGG = 1, MRO_STOCH = 2 };

/* Returns 0, or -1 when the parameter is outside [0,1] for AGG / STOCH (System.exit(-1), MR:366-369,
 * 434-437; LinearCombination has no range check, MR:317-330).  first_index is the position of
 * ubm[0] in the full sorted model (0 for a whole model) so shards reproduce the global index that
 * Aggregation (MR:381) and the sequential Random stream (MR:447) depend on; n_total is ubm.length
 * of the whole model (MR:371). */
MRO_API int mro_blend(int kind, double param, uint64_t seed, const double *ubm, const double *ibm,
                      double *out, int64_t n, int64_t first_index, int64_t n_total) {
  if (kind == MRO_LC) {
    for (int64_t i = 0; i < n; ++i) out[i] = ubm[i] * param + ibm[i] * (1 - param); /* MR:328 */
    return 0;
  }
  if (param < 0 || param > 1) return -1;
  if (kind == MRO_AGG) {
    /* (itemBasedPercentage * length).toInt, MR:372 — truncation toward zero */
    int64_t thr = (int64_t)(param * (double)n_total);
    for (int64_t i = 0; i < n; ++i) out[i] = (first_index + i < thr) ? ibm[i] : ubm[i]; /* MR:381-382 */
    return 0;
  }
  if (kind == MRO_STOCH) {
    jrandom r; jr_seed(&r, seed);
    for (int64_t i = 0; i < first_index; ++i) (void)jr_next_float(&r);
    for (int64_t i = 0; i < n; ++i) out[i] = ((double)jr_next_float(&r) < param) ? ibm[i] : ubm[i]; /* MR:447 */
    return 0;
  }
  return -2;
}

/* ------------------------------------------------------------------------------------------ */
/* Top-k — new derived output: score desc, song id asc, NaN (listened) excluded */
/* ------------------------------------------------------------------------------------------ */
typedef struct { double score; int32_t song; } mro_pair;
static int pair_cmp(const void *a, const void *b) {
  const mro_pair *x = (const mro_pair *)a, *y = (const mro_pair *)b;
  if (x->score > y->score) return -1;
  if (x->score < y->score) return 1;
  return (x->song > y->song) - (x->song < y->song);
}

MRO_API void mro_topk(const double *scores, int32_t n_users, int32_t S, int32_t k, int32_t *out_song,
                      double *out_score, int32_t *out_len) {
#pragma omp parallel
  {
    mro_pair *buf = (mro_pair *)malloc(sizeof(mro_pair) * (size_t)(S > 0 ? S : 1));
#pragma omp for schedule(dynamic, 1)
    for (int32_t u = 0; u < n_users; ++u) {
      const double *row = scores + (int64_t)u * S; int32_t n = 0;
      for (int32_t s = 0; s < S; ++s) if (!isnan(row[s])) { buf[n].score = row[s]; buf[n].song = s; n++; }
      qsort(buf, (size_t)n, sizeof(mro_pair), pair_cmp);
      int32_t len = n < k ? n : k;
      for (int32_t i = 0; i < k; ++i) {
        out_song[(int64_t)u * k + i] = i < len ? buf[i].song : -1;
        out_score[(int64_t)u * k + i] = i < len ? buf[i].score : 0.0;
      }
      out_len[u] = len;
    }
    free(buf);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Evaluation — the reference's threshold-sweep "mAP", MR:521-639                 */
/* ------------------------------------------------------------------------------------------ */

/* scores: dense [U][S] with NaN at non-emitted pairs.  Labels: CSR lab_ptr[U+1], lab_col[] with
 * song ids; ids >= S denote label songs that occur nowhere in train/test (they can never be
 * predicted).  new_songs[n_new] = the distinct label songs (MR:72,79), canonical ascending order.
 * n_thresholds = 10 for MR:590, 11 for DIST:395. */
MRO_API double mro_evaluate(const double *scores, int32_t U, int32_t S, const int64_t *lab_ptr,
                            const int32_t *lab_col, const int32_t *new_songs, int32_t n_new,
                            int32_t n_thresholds) {
  static const double TH[11] = {0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0};
  int64_t n = (int64_t)U * S;
  double mn = INFINITY, mx = -INFINITY; /* MR:524-525 */
  for (int64_t i = 0; i < n; ++i) if (!isnan(scores[i])) { if (scores[i] < mn) mn = scores[i]; if (scores[i] > mx) mx = scores[i]; }
  double total = 0.0;
  int nt = n_thresholds;
  for (int32_t c = 0; c < n_new; ++c) {
    int32_t song = new_songs[c];
    double prec[11], rec[11];
    for (int t = 0; t < nt; ++t) {
      int tp = 0, fp = 0, fn = 0; /* confusionMatrix, MR:541-553 */
      for (int32_t u = 0; u < U; ++u) {
        int predicted = 0;
        if (song < S) {
          double x = scores[(int64_t)u * S + song];
          if (!isnan(x)) predicted = ((x - mn) / (mx - mn) > TH[t]); /* MR:529; NaN compares false */
        }
        int labelled = lin_contains(lab_col + lab_ptr[u], lab_ptr[u + 1] - lab_ptr[u], song);
        tp += predicted && labelled; fp += predicted && !labelled; fn += !predicted && labelled;
      }
      prec[t] = (tp + fp > 0) ? (double)tp / (tp + fp) : 0.0; /* MR:561-566 */
      rec[t] = (tp + fn > 0) ? (double)tp / (tp + fn) : 0.0;  /* MR:574-579 */
    }
    double ap = 0.0; /* singleAveragePrecision, MR:600-610: List.sum = left fold from 0 */
    for (int t = 0; t < nt; ++t) {
      double term;
      if (t == nt - 1) term = 0.0;
      else if (t == nt - 2) term = (rec[t] - 0.0) * prec[t];
      else term = (rec[t] - rec[t + 1]) * prec[t];
      ap += term;
    }
    total += ap; /* foldLeft(0.0)(_+_), MR:626 */
  }
  return total / n_new;
}

/* MyUtils.roundAt(p, n), UTIL:17: { val s = math pow (10, p); (math round n * s) / s } */
MRO_API double mro_round_at(int p, double x) {
  double s = pow(10.0, p);
  return (double)llround(floor(x * s + 0.5)) / s; /* Math.round(double) = floor(x + 0.5) */
}

/* ------------------------------------------------------------------------------------------ */
/* FP64 restatement on the CSR — the reference's expressions, summed in ascending-id order     */
/* ------------------------------------------------------------------------------------------ */

/* A third, independent restatement used only by the ranking differential test: the reference's own fp64
 * terms  c / (sqrt(a) * sqrt(b))  (MR:147-148, 237-238) summed left to right with trainUsers / songs in
 * ascending id order (MR:166, 257 sum in HashSet order), computed through the inverted index
 * instead of linear `contains` scans so that it is feasible on an MSD-shaped shard.  No fixed point
 * anywhere: it shows that ranking the canonical integers ranks the reference's doubles.
 * out[(u-u0)][S] fp64, NaN at listened pairs. */
MRO_API void mro_fp64_scores(const mro_data *d, int model, int32_t u0, int32_t u1, double *out) {
  int64_t *cp; int32_t *cu; build_train_csc(d, &cp, &cu);
  int64_t S = d->S;
#pragma omp parallel
  {
    int32_t *cnt = (int32_t *)calloc((size_t)((d->T > S ? d->T : S) + 1), sizeof(int32_t));
    int32_t *touched = (int32_t *)malloc(sizeof(int32_t) * (size_t)((d->T > S ? d->T : S) + 1));
#pragma omp for schedule(dynamic, 1)
    for (int32_t u = u0; u < u1; ++u) {
      double *row = out + (int64_t)(u - u0) * S;
      for (int64_t s = 0; s < S; ++s) row[s] = 0.0;
      if (model == 0) {
        /* cnt[v] = |I_u ∩ I_v| (MR:142-145); then for v ascending: row[s] += cnt / (sqrt|I_u| * sqrt|I_v|) for s in I_v (MR:161-166) */
        for (int64_t i = d->te_ptr[u]; i < d->te_ptr[u + 1]; ++i) {
          int32_t j = d->te_col[i];
          for (int64_t k = cp[j]; k < cp[j + 1]; ++k) cnt[cu[k]]++;
        }
        for (int32_t v = 0; v < d->T; ++v) if (cnt[v]) {
          double denominator = sqrt((double)d->deg_te[u]) * sqrt((double)d->deg_tr[v]);
          double w = denominator != 0 ? cnt[v] / denominator : 0.0;
          for (int64_t m = d->tr_ptr[v]; m < d->tr_ptr[v + 1]; ++m) row[d->tr_col[m]] += w;
          cnt[v] = 0;
        }
      } else {
        /* for s2 = j in I_u ascending (MR:251-253): cnt[s] = |U_s ∩ U_j| over train users (MR:232-235); row[s] += cnt / (sqrt d_s * sqrt d_j) */
        for (int64_t i = d->te_ptr[u]; i < d->te_ptr[u + 1]; ++i) {
          int32_t j = d->te_col[i]; int32_t nt = 0;
          for (int64_t k = cp[j]; k < cp[j + 1]; ++k) {
            int32_t v = cu[k];
            for (int64_t m = d->tr_ptr[v]; m < d->tr_ptr[v + 1]; ++m) { int32_t s = d->tr_col[m]; if (cnt[s]++ == 0) touched[nt++] = s; }
          }
          for (int32_t t = 0; t < nt; ++t) {
            int32_t s = touched[t];
            if (s != j) {
              double denominator = sqrt((double)d->deg_song[s]) * sqrt((double)d->deg_song[j]);
              row[s] += denominator != 0 ? cnt[s] / denominator : 0;
            }
            cnt[s] = 0;
          }
        }
      }
      for (int64_t i = d->te_ptr[u]; i < d->te_ptr[u + 1]; ++i) row[d->te_col[i]] = NAN;
    }
    free(cnt); free(touched);
  }
  free(cp); free(cu);
}

/* ------------------------------------------------------------------------------------------ */
/* mAP@k on the ranked lists — new derived metric (north_star "mAP@500")  */
/* ------------------------------------------------------------------------------------------ */

/* The reference has no ranking and therefore no mAP@k (its "mAP" is the threshold sweep above, MR:588-627).
 * Definition used here = the Million Song Dataset Challenge's truncated mAP: for test user u with hidden
 * (label) songs L_u and ranked recommendations r_1..r_n (n = top_len[u] <= k, listened songs never appear):
 *     AP@k(u) = ( sum_{i=1..n} [r_i in L_u] * hits_i / i ) / min(|L_u|, k),   hits_i = #{t <= i : r_t in L_u}
 * and mAP@k = mean of AP@k over the test users that have at least one label row (users without labels are
 * skipped, as a user without hidden songs has no defined precision).  Terms are added in rank order, users
 * in ascending id order (both fp64, left folds) — the GPU kernel does the same, so results are bit-identical.
 * Label rows are ascending/unique; ap_out (optional) receives the per-user AP (0 for users without labels). */
MRO_API double mro_map_at_k(const int32_t *top_song, const int32_t *top_len, int32_t U, int32_t k, const int64_t *lab_ptr,
                            const int32_t *lab_col, double *ap_out) {
  double total = 0.0; int64_t n_eval = 0;
  for (int32_t u = 0; u < U; ++u) {
    const int32_t *lab = lab_col + lab_ptr[u]; int64_t nl = lab_ptr[u + 1] - lab_ptr[u];
    double ap = 0.0;
    if (nl > 0) {
      int32_t hits = 0; int32_t n = top_len[u] < k ? top_len[u] : k;
      for (int32_t i = 0; i < n; ++i) {
        int32_t s = top_song[(int64_t)u * k + i];
        int64_t lo = 0, hi = nl;
        while (lo < hi) { int64_t m = (lo + hi) >> 1; if (lab[m] < s) lo = m + 1; else hi = m; }
        if (lo < nl && lab[lo] == s) { hits++; ap += (double)hits / (double)(i + 1); }
      }
      ap = ap / (double)(nl < k ? nl : k);
      total += ap; n_eval++;
    }
    if (ap_out) ap_out[u] = ap;
  }
  return n_eval > 0 ? total / (double)n_eval : 0.0;
}
