/*
 * mrs_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C fp64 restatement of the rating-prediction hot path of the reference
 * (src/main/scala/shared/predictions.scala, cited below as P:<line>).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library, and only as the checker or the timed CPU baseline -- never as part of
 * the product path (movie-recommender-system_b200/ never links or imports it).
 *
 * PINNING STATUS: "parity partially pinned".  The reference is Scala/Spark; no JVM exists
 * in this image, so it cannot be executed here, and the MovieLens files its committed
 * answer JSONs were computed on are not in the reference tree.  What IS pinned
 * (tests/test_oracle_golden.py): the formula-level known answers that the committed
 * JSONs make self-checkable (baseline-100k.json:9-12, distributed-25m-4.json:10-13,
 * knn-100k.json:8, the uniform==baseline identity of personalized-100k.json:8-9, ...).
 * Dataset-level values (MAEs etc.) are unpinned until real data is supplied; the test
 * suite auto-checks them if data/ml-100k appears.
 *
 * Floating-point contract: the JVM never contracts a*b+c; build this file
 * with -ffp-contract=off.  Canonical summation orders where the reference's order is a
 * hash-iteration order that cannot be reproduced without a JVM:
 *   user / item rating sums ........ train-file order (exact anyway for half-star data)
 *   item deviation sums ............ train-file order            (P:180 is HashMap order)
 *   per-user squared norm .......... ascending item id           (P:474 is HashMap order)
 *   cosine dot over intersection ... ascending item id           (P:424 is HashSet order)
 *   weighted sum over raters ....... train-file order within the item  (P:513, exact)
 *   MAE ............................ test-file order                    (P:81, exact)
 *   neighbour ties ................. (sim desc, user id asc)     (P:608 is HashSet order)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

enum { ORC_GLOBAL = 0, ORC_USER = 1, ORC_ITEM = 2, ORC_ITEMDEV = 3, ORC_BASELINE = 4, ORC_PERSONALIZED = 5 };
enum { ORC_SIM_UNIFORM = 0, ORC_SIM_COSINE = 1, ORC_SIM_JACCARD = 2 };

typedef struct {
  int64_t n;
  int32_t umax, imax; /* largest ids seen; tables are direct-indexed 0..max */
  int32_t *u, *i;
  double *r;
  double gavg;
  double *usum, *uavg;
  int32_t *ucnt; /* ucnt[u]==0 <=> user unknown ("not in the train set") */
  double *isum, *iavg;
  int32_t *icnt;
  double *dev;     /* normalised deviation of rating j, file order (P:155-169) */
  double *idevsum; /* per item */
  double *idevavg;
  /* user-major rows sorted by ascending item id (canonical order for norms / dots) */
  int64_t *urow; /* umax+2 */
  int32_t *ucol;
  double *udev;  /* deviation, same order as ucol */
  double *upre;  /* preprocessed rating r~ (P:470-481) */
  double *unorm; /* per user */
  /* item-major lists in train-file order (P:493 groupBy preserves order) */
  int64_t *icolp; /* imax+2 */
  int32_t *irow;  /* rater user id */
  double *idev;   /* that rater's deviation for the item */
  /* lazily built, cached full neighbour lists per similarity kind (P:601 allNeighbors) */
  int32_t n_known_users;
  int32_t *known_users; /* ascending id */
  int32_t **nb_ids[3];
  double **nb_sims[3];
  /* scratch for dense-row similarity */
  double *scratch_val;
  int32_t *scratch_stamp;
  int32_t stamp;
} orc_model;

/* P:57-61 */
ORC_API double orc_scale(double x, double y) {
  if (x > y) return 5 - y;
  else if (x < y) return y - 1;
  else return 1;
}

/* P:18 mean: left-to-right reduce / length, empty -> 0.0 */
ORC_API double orc_mean(const double *s, int64_t n) {
  if (n <= 0) return 0.0;
  double acc = s[0];
  for (int64_t j = 1; j < n; ++j) acc = acc + s[j];
  return acc / (double)n;
}

/* P:19-25 population standard deviation */
ORC_API double orc_std(const double *s, int64_t n) {
  if (n <= 0) return 0.0;
  double m = orc_mean(s, n), acc = 0.0;
  for (int64_t j = 0; j < n; ++j) acc += (m - s[j]) * (m - s[j]);
  return sqrt(acc / (double)n);
}

typedef struct { int32_t col; int64_t src; } colsrc;
static int cmp_colsrc(const void *a, const void *b) {
  const colsrc *x = a, *y = b;
  if (x->col != y->col) return (x->col > y->col) - (x->col < y->col);
  return (x->src > y->src) - (x->src < y->src);
}

ORC_API void orc_free(orc_model *m) {
  if (!m) return;
  for (int s = 0; s < 3; ++s) {
    if (m->nb_ids[s]) {
      for (int32_t u = 0; u <= m->umax + 1; ++u) { free(m->nb_ids[s][u]); free(m->nb_sims[s][u]); }
      free(m->nb_ids[s]); free(m->nb_sims[s]);
    }
  }
  free(m->u); free(m->i); free(m->r); free(m->usum); free(m->uavg); free(m->ucnt);
  free(m->isum); free(m->iavg); free(m->icnt); free(m->dev); free(m->idevsum); free(m->idevavg);
  free(m->urow); free(m->ucol); free(m->udev); free(m->upre); free(m->unorm);
  free(m->icolp); free(m->irow); free(m->idev); free(m->known_users);
  free(m->scratch_val); free(m->scratch_stamp);
  free(m);
}

/*
 * Fit everything the baseline family needs (P:94-198) plus the sorted layouts used by the
 * personalized / kNN functions.  Ids must be >= 0.  Duplicate (u,i) pairs are outside the
 * contract (P:168 keeps the last one; MovieLens has none) -- not handled here.
 */
ORC_API orc_model *orc_fit(const int32_t *u, const int32_t *i, const double *r, int64_t n) {
  orc_model *m = calloc(1, sizeof *m);
  m->n = n;
  m->u = malloc(sizeof(int32_t) * (n ? n : 1));
  m->i = malloc(sizeof(int32_t) * (n ? n : 1));
  m->r = malloc(sizeof(double) * (n ? n : 1));
  memcpy(m->u, u, sizeof(int32_t) * n);
  memcpy(m->i, i, sizeof(int32_t) * n);
  memcpy(m->r, r, sizeof(double) * n);
  int32_t umax = 0, imax = 0;
  for (int64_t j = 0; j < n; ++j) { if (u[j] > umax) umax = u[j]; if (i[j] > imax) imax = i[j]; }
  m->umax = umax; m->imax = imax;
  size_t U = (size_t)umax + 2, I = (size_t)imax + 2;
  m->usum = calloc(U, sizeof(double)); m->uavg = calloc(U, sizeof(double)); m->ucnt = calloc(U, sizeof(int32_t));
  m->isum = calloc(I, sizeof(double)); m->iavg = calloc(I, sizeof(double)); m->icnt = calloc(I, sizeof(int32_t));
  m->idevsum = calloc(I, sizeof(double)); m->idevavg = calloc(I, sizeof(double));
  m->dev = malloc(sizeof(double) * (n ? n : 1));

  /* P:94 average = mean(ratings.map(_.rating)) -- reduce(_+_) in file order */
  m->gavg = orc_mean(r, n);
  /* P:113 usersAvg, P:134 itemsAvg: groupBy keeps file order inside each group */
  for (int64_t j = 0; j < n; ++j) {
    m->usum[u[j]] = m->ucnt[u[j]] ? m->usum[u[j]] + r[j] : r[j]; m->ucnt[u[j]]++;
    m->isum[i[j]] = m->icnt[i[j]] ? m->isum[i[j]] + r[j] : r[j]; m->icnt[i[j]]++;
  }
  for (size_t k = 0; k < U; ++k) if (m->ucnt[k]) m->uavg[k] = m->usum[k] / (double)m->ucnt[k];
  for (size_t k = 0; k < I; ++k) if (m->icnt[k]) m->iavg[k] = m->isum[k] / (double)m->icnt[k];
  /* P:155-169 computeNormalizeDeviation; P:176-186 itemsAvgDev (sum, count) then sum/count */
  for (int64_t j = 0; j < n; ++j) {
    double ua = m->uavg[u[j]];
    m->dev[j] = (r[j] - ua) / orc_scale(r[j], ua);
    m->idevsum[i[j]] = m->dev[j] + m->idevsum[i[j]]; /* P:183: x._2 + cur._1 */
  }
  for (size_t k = 0; k < I; ++k) if (m->icnt[k]) m->idevavg[k] = m->idevsum[k] / (double)m->icnt[k];

  /* user-major, ascending item id */
  m->urow = calloc(U, sizeof(int64_t));
  for (size_t k = 0; k + 1 < U; ++k) m->urow[k + 1] = m->urow[k] + m->ucnt[k];
  m->ucol = malloc(sizeof(int32_t) * (n ? n : 1));
  m->udev = malloc(sizeof(double) * (n ? n : 1));
  m->upre = malloc(sizeof(double) * (n ? n : 1));
  m->unorm = calloc(U, sizeof(double));
  {
    int64_t *cur = malloc(sizeof(int64_t) * U);
    memcpy(cur, m->urow, sizeof(int64_t) * U);
    colsrc *tmp = malloc(sizeof(colsrc) * (n ? n : 1));
    for (int64_t j = 0; j < n; ++j) { int64_t p = cur[u[j]]++; tmp[p].col = i[j]; tmp[p].src = j; }
    for (int32_t k = 0; k <= umax; ++k)
      qsort(tmp + m->urow[k], (size_t)m->ucnt[k], sizeof(colsrc), cmp_colsrc);
    for (int64_t p = 0; p < n; ++p) { m->ucol[p] = tmp[p].col; m->udev[p] = m->dev[tmp[p].src]; }
    free(tmp); free(cur);
  }
  /* P:470-481 preprocessedRating: weight = sqrt(sum dev^2); weight != 0 ? dev/weight : 0 */
  m->n_known_users = 0;
  for (int32_t k = 0; k <= umax; ++k) {
    if (!m->ucnt[k]) continue;
    m->n_known_users++;
    double ss = 0.0;
    for (int64_t p = m->urow[k]; p < m->urow[k + 1]; ++p) ss = ss + m->udev[p] * m->udev[p];
    double w = sqrt(ss);
    m->unorm[k] = w;
    for (int64_t p = m->urow[k]; p < m->urow[k + 1]; ++p) m->upre[p] = (w != 0) ? m->udev[p] / w : 0.0;
  }
  m->known_users = malloc(sizeof(int32_t) * (m->n_known_users ? m->n_known_users : 1));
  { int32_t c = 0; for (int32_t k = 0; k <= umax; ++k) if (m->ucnt[k]) m->known_users[c++] = k; }

  /* item-major, train-file order */
  m->icolp = calloc(I, sizeof(int64_t));
  for (size_t k = 0; k + 1 < I; ++k) m->icolp[k + 1] = m->icolp[k] + m->icnt[k];
  m->irow = malloc(sizeof(int32_t) * (n ? n : 1));
  m->idev = malloc(sizeof(double) * (n ? n : 1));
  {
    int64_t *cur = malloc(sizeof(int64_t) * I);
    memcpy(cur, m->icolp, sizeof(int64_t) * I);
    for (int64_t j = 0; j < n; ++j) { int64_t p = cur[i[j]]++; m->irow[p] = u[j]; m->idev[p] = m->dev[j]; }
    free(cur);
  }
  m->scratch_val = calloc(I, sizeof(double));
  m->scratch_stamp = calloc(I, sizeof(int32_t));
  m->stamp = 0;
  return m;
}

static int known_u(const orc_model *m, int32_t u) { return u >= 0 && u <= m->umax && m->ucnt[u] > 0; }
static int known_i(const orc_model *m, int32_t i) { return i >= 0 && i <= m->imax && m->icnt[i] > 0; }

ORC_API double orc_global_avg(const orc_model *m) { return m->gavg; }
ORC_API int32_t orc_umax(const orc_model *m) { return m->umax; }
ORC_API int32_t orc_imax(const orc_model *m) { return m->imax; }
ORC_API int32_t orc_user_count(const orc_model *m, int32_t u) { return known_u(m, u) ? m->ucnt[u] : 0; }
ORC_API int32_t orc_item_count(const orc_model *m, int32_t i) { return known_i(m, i) ? m->icnt[i] : 0; }
/* P:126 / P:287 fallbacks */
ORC_API double orc_user_avg(const orc_model *m, int32_t u) { return known_u(m, u) ? m->uavg[u] : m->gavg; }
/* P:147 / P:308 */
ORC_API double orc_item_avg(const orc_model *m, int32_t i) { return known_i(m, i) ? m->iavg[i] : m->gavg; }
/* P:197 / P:354 */
ORC_API double orc_item_avg_dev(const orc_model *m, int32_t i) { return known_i(m, i) ? m->idevavg[i] : 0.0; }
ORC_API double orc_user_norm(const orc_model *m, int32_t u) { return known_u(m, u) ? m->unorm[u] : 0.0; }
/* deviation of train rating j (file order), P:167 */
ORC_API void orc_deviations(const orc_model *m, double *out) { memcpy(out, m->dev, sizeof(double) * m->n); }

/* P:155 / P:470 lookups keyed by (u,i): returns 1 if the pair is in train */
ORC_API int32_t orc_pair_values(const orc_model *m, int32_t u, int32_t i, double *dev, double *pre) {
  if (!known_u(m, u)) return 0;
  for (int64_t p = m->urow[u]; p < m->urow[u + 1]; ++p)
    if (m->ucol[p] == i) { if (dev) *dev = m->udev[p]; if (pre) *pre = m->upre[p]; return 1; }
  return 0;
}

/* P:229 / P:383 / P:578: avg + dev * scale(avg + dev, avg) -- the branch is taken on the rounded sum */
ORC_API double orc_combine(double avg, double dev) { return avg + dev * orc_scale(avg + dev, avg); }

/* ---------------- baseline-family predictors ---------------- */
static double predict_baseline_family(const orc_model *m, int kind, int32_t u, int32_t i) {
  switch (kind) {
    case ORC_GLOBAL: return m->gavg;                  /* P:105 */
    case ORC_USER: return orc_user_avg(m, u);         /* P:126 */
    case ORC_ITEM: return orc_item_avg(m, i);         /* P:147 */
    case ORC_ITEMDEV: return orc_item_avg_dev(m, i);  /* P:197 */
    default: {                                        /* P:217-236, P:373-390 */
      if (!known_u(m, u)) return m->gavg;             /* P:222-224 */
      double ua = m->uavg[u];
      double d = orc_item_avg_dev(m, i);
      return orc_combine(ua, d);                      /* P:229 */
    }
  }
}

/* ---------------- similarities ---------------- */
/* P:407-433 adjusted cosine: sum over the item intersection of r~_ui * r~_vi (ascending item id) */
ORC_API double orc_cosine(const orc_model *m, int32_t u, int32_t v) {
  if (!known_u(m, u) || !known_u(m, v)) return 0.0; /* empty intersection -> empty sum = 0.0 */
  int64_t p = m->urow[u], pe = m->urow[u + 1], q = m->urow[v], qe = m->urow[v + 1];
  double acc = 0.0;
  while (p < pe && q < qe) {
    int32_t a = m->ucol[p], b = m->ucol[q];
    if (a == b) { acc = acc + m->upre[p] * m->upre[q]; ++p; ++q; }
    else if (a < b) ++p; else ++q;
  }
  return acc;
}

/* P:440-464 Jaccard: |I(u) n I(v)| / (|I(u)| + |I(v)| - |I(u) n I(v)|); 0/0 -> NaN like the JVM */
ORC_API double orc_jaccard(const orc_model *m, int32_t u, int32_t v) {
  int64_t nu = known_u(m, u) ? m->ucnt[u] : 0, nv = known_u(m, v) ? m->ucnt[v] : 0, inter = 0;
  if (nu && nv) {
    int64_t p = m->urow[u], pe = m->urow[u + 1], q = m->urow[v], qe = m->urow[v + 1];
    while (p < pe && q < qe) {
      int32_t a = m->ucol[p], b = m->ucol[q];
      if (a == b) { ++inter; ++p; ++q; } else if (a < b) ++p; else ++q;
    }
  }
  return (double)inter / (double)(nu + nv - inter);
}

static double base_sim(const orc_model *m, int simkind, int32_t u, int32_t v) {
  switch (simkind) {
    case ORC_SIM_UNIFORM: return 1.0; /* P:400 */
    case ORC_SIM_COSINE: return orc_cosine(m, u, v);
    default: return orc_jaccard(m, u, v);
  }
}

typedef struct { double s; int32_t id; } nbr;
/* Candidate order of `(allUsers - u).toSeq` (P:608), which the stable sort of P:610 keeps among equal similarities:
 * mode 0 = ascending id (north_star: "index-ordered tie-breaking"); mode 1 = iteration order of a Scala 2.11
 * immutable.HashSet[Int]: hash-trie walk in ascending successive 5-bit groups (least significant first) of
 * improve(id) = { h = id + ~(id << 9); h ^= h >>> 14; h += h << 4; h ^ (h >>> 10) } -- recalled from the 2.11 library
 * source, not verifiable without a JVM. */
static int g_tie_mode = 0;
static uint64_t hashset_order_key(int32_t id) {
  uint32_t h = (uint32_t)id + ~((uint32_t)id << 9);
  h ^= h >> 14;
  h += h << 4;
  h ^= h >> 10;
  uint64_t key = 0;
  for (int k = 0; k < 7; ++k) key = (key << 5) | ((h >> (5 * k)) & 31u);
  return key;
}
static int cmp_nbr(const void *a, const void *b) {
  const nbr *x = a, *y = b;
  if (x->s > y->s) return -1; /* P:610 sortWith(_._2 > _._2), stable */
  if (y->s > x->s) return 1;
  if (g_tie_mode == 1) {
    uint64_t kx = hashset_order_key(x->id), ky = hashset_order_key(y->id);
    return (kx > ky) - (kx < ky);
  }
  return (x->id > y->id) - (x->id < y->id); /* stable w.r.t. ascending-id candidate order */
}

/* P:596-617 getNeighbors: full sorted candidate list of u (all known users except u), cached. */
static void neighbours_full(orc_model *m, int simkind, int32_t u, const int32_t **ids, const double **sims, int32_t *cnt) {
  int32_t nc = m->n_known_users - (known_u(m, u) ? 1 : 0);
  *cnt = nc;
  /* every out-of-range user behaves identically (no ratings), so they share slot umax+1 */
  int32_t slot = (u >= 0 && u <= m->umax) ? u : m->umax + 1;
  if (!m->nb_ids[simkind]) {
    m->nb_ids[simkind] = calloc((size_t)m->umax + 2, sizeof(int32_t *));
    m->nb_sims[simkind] = calloc((size_t)m->umax + 2, sizeof(double *));
  }
  if (m->nb_ids[simkind][slot]) { *ids = m->nb_ids[simkind][slot]; *sims = m->nb_sims[simkind][slot]; return; }
  nbr *tmp = malloc(sizeof(nbr) * (nc ? nc : 1));
  int32_t c = 0;
  if (simkind == ORC_SIM_COSINE && known_u(m, u)) {
    /* dense row of u, then one ascending pass over each candidate's items: identical
       sequence of operations to the two-pointer intersection in orc_cosine */
    int32_t st = ++m->stamp;
    for (int64_t p = m->urow[u]; p < m->urow[u + 1]; ++p) { m->scratch_val[m->ucol[p]] = m->upre[p]; m->scratch_stamp[m->ucol[p]] = st; }
    for (int32_t t = 0; t < m->n_known_users; ++t) {
      int32_t x = m->known_users[t];
      if (x == u) continue; /* P:608 allUsers - u */
      double acc = 0.0;
      for (int64_t q = m->urow[x]; q < m->urow[x + 1]; ++q)
        if (m->scratch_stamp[m->ucol[q]] == st) acc = acc + m->scratch_val[m->ucol[q]] * m->upre[q];
      tmp[c].s = acc; tmp[c].id = x; ++c;
    }
  } else {
    for (int32_t t = 0; t < m->n_known_users; ++t) {
      int32_t x = m->known_users[t];
      if (x == u) continue;
      tmp[c].s = base_sim(m, simkind, u, x); tmp[c].id = x; ++c;
    }
  }
  qsort(tmp, (size_t)c, sizeof(nbr), cmp_nbr);
  int32_t *oi = malloc(sizeof(int32_t) * (c ? c : 1));
  double *os = malloc(sizeof(double) * (c ? c : 1));
  for (int32_t t = 0; t < c; ++t) { oi[t] = tmp[t].id; os[t] = tmp[t].s; }
  free(tmp);
  *ids = oi; *sims = os;
  m->nb_ids[simkind][slot] = oi; m->nb_sims[simkind][slot] = os;
}

/* tie order of the neighbour lists built from now on (see cmp_nbr); drops the cached lists of this model */
ORC_API void orc_set_tie_order(orc_model *m, int32_t mode) {
  g_tie_mode = mode == 1 ? 1 : 0;
  for (int s = 0; s < 3; ++s) {
    if (m->nb_ids[s]) {
      for (int32_t u = 0; u <= m->umax + 1; ++u) { free(m->nb_ids[s][u]); free(m->nb_sims[s][u]); }
      free(m->nb_ids[s]); free(m->nb_sims[s]);
      m->nb_ids[s] = NULL; m->nb_sims[s] = NULL;
    }
  }
}

/* P:603-616: first k of the sorted list.  Returns the number written (min(k, candidates)). */
ORC_API int32_t orc_neighbors(orc_model *m, int simkind, int32_t k, int32_t u, int32_t *ids_out, double *sims_out, int32_t cap) {
  const int32_t *ids; const double *sims; int32_t cnt;
  neighbours_full(m, simkind, u, &ids, &sims, &cnt);
  int32_t w = cnt < k ? cnt : k;
  if (w < 0) w = 0;
  if (w > cap) w = cap;
  for (int32_t t = 0; t < w; ++t) { if (ids_out) ids_out[t] = ids[t]; if (sims_out) sims_out[t] = sims[t]; }
  return w;
}

/* P:626-649 getSimilarity: s(u,v) if v is among the first k neighbours of u, else 0.0; k<=0 means "no kNN" */
ORC_API double orc_similarity(orc_model *m, int simkind, int32_t k, int32_t u, int32_t v) {
  if (k <= 0) return base_sim(m, simkind, u, v);
  const int32_t *ids; const double *sims; int32_t cnt;
  neighbours_full(m, simkind, u, &ids, &sims, &cnt);
  int32_t w = cnt < k ? cnt : k;
  double acc = 0.0; /* P:638-641 map(...).sum */
  for (int32_t t = 0; t < w; ++t) acc = acc + (ids[t] == v ? sims[t] : 0.0);
  return acc;
}

/* P:489-549 weightedSumDeviation */
ORC_API double orc_wsd(orc_model *m, int simkind, int32_t k, int32_t u, int32_t i) {
  if (!known_i(m, i)) return 0.0; /* empty rater list: ssSum = (0,0) -> 0.0 (P:527-529) */
  double num = 0.0, den = 0.0;
  if (k > 0) {
    /* membership of each rater in N_k(u) via a rank table built on the scratch-free path */
    const int32_t *ids; const double *sims; int32_t cnt;
    neighbours_full(m, simkind, u, &ids, &sims, &cnt);
    int32_t w = cnt < k ? cnt : k;
    /* small dense map user -> sim for this u; rebuilt per call (oracle favours clarity over speed,
       but keep it O(k + raters)) */
    static __thread double *map = NULL; static __thread int32_t *mstamp = NULL; static __thread int32_t mcap = 0, mst = 0;
    if (mcap < m->umax + 2) {
      free(map); free(mstamp);
      mcap = m->umax + 2; map = calloc((size_t)mcap, sizeof(double)); mstamp = calloc((size_t)mcap, sizeof(int32_t)); mst = 0;
    }
    ++mst;
    for (int32_t t = 0; t < w; ++t) { map[ids[t]] = sims[t]; mstamp[ids[t]] = mst; }
    for (int64_t p = m->icolp[i]; p < m->icolp[i + 1]; ++p) {
      int32_t x = m->irow[p];
      double s = (mstamp[x] == mst) ? map[x] : 0.0;
      num = num + m->idev[p] * s;  /* P:522 acc._1 + a._1*a._2 */
      den = den + fabs(s);         /* P:522 acc._2 + a._2.abs */
    }
  } else {
    for (int64_t p = m->icolp[i]; p < m->icolp[i + 1]; ++p) {
      double s = base_sim(m, simkind, u, m->irow[p]);
      num = num + m->idev[p] * s;
      den = den + fabs(s);
    }
  }
  return (den > 0) ? num / den : 0.0; /* P:527-529 */
}

/* P:557-586 predictor */
ORC_API double orc_predict_personalized(orc_model *m, int simkind, int32_t k, int32_t u, int32_t i) {
  if (!known_u(m, u)) return m->gavg; /* P:572-573 */
  double ua = m->uavg[u];
  double w = orc_wsd(m, simkind, k, u, i);
  return orc_combine(ua, w); /* P:578 */
}

ORC_API double orc_predict(orc_model *m, int kind, int simkind, int32_t k, int32_t u, int32_t i) {
  if (kind == ORC_PERSONALIZED) return orc_predict_personalized(m, simkind, k, u, i);
  return predict_baseline_family(m, kind, u, i);
}

ORC_API void orc_predict_batch(orc_model *m, int kind, int simkind, int32_t k, const int32_t *us, const int32_t *is, int64_t n, double *out) {
  for (int64_t j = 0; j < n; ++j) out[j] = orc_predict(m, kind, simkind, k, us[j], is[j]);
}

/* P:69-86 MAE: foldLeft over the data set in order, then sum/count (0/0 -> NaN for an empty set) */
ORC_API double orc_mae(orc_model *m, int kind, int simkind, int32_t k, const int32_t *tu, const int32_t *ti, const double *tr, int64_t nt) {
  double acc = 0.0; int64_t cnt = 0;
  for (int64_t j = 0; j < nt; ++j) {
    acc = fabs(tr[j] - orc_predict(m, kind, simkind, k, tu[j], ti[j])) + acc; /* P:83 f(x) + acc._1 */
    cnt++;
  }
  return acc / (double)cnt;
}

/* P:651-674 recommendations: unrated items, order (score desc, item id asc), first n */
typedef struct { double s; int32_t id; } reco;
static int cmp_reco(const void *a, const void *b) {
  const reco *x = a, *y = b;
  if (x->s == y->s) return (x->id > y->id) - (x->id < y->id); /* P:655-656 */
  return (x->s > y->s) ? -1 : 1;                               /* P:658 */
}
ORC_API int32_t orc_recommend(orc_model *m, int kind, int simkind, int32_t k, int32_t user, int32_t n, int32_t *items_out, double *scores_out) {
  char *rated = calloc((size_t)m->imax + 2, 1);
  if (known_u(m, user)) for (int64_t p = m->urow[user]; p < m->urow[user + 1]; ++p) rated[m->ucol[p]] = 1;
  reco *c = malloc(sizeof(reco) * ((size_t)m->imax + 2));
  int32_t nc = 0;
  for (int32_t it = 0; it <= m->imax; ++it) {
    if (!m->icnt[it] || rated[it]) continue; /* P:667 */
    c[nc].id = it; c[nc].s = orc_predict(m, kind, simkind, k, user, it); ++nc;
  }
  qsort(c, (size_t)nc, sizeof(reco), cmp_reco);
  int32_t w = nc < n ? nc : n;
  for (int32_t t = 0; t < w; ++t) { items_out[t] = c[t].id; scores_out[t] = c[t].s; }
  free(c); free(rated);
  return w;
}

/*
 * Spark twin of the baseline MAE pass, P:246-391, as timed by distributed/DistributedBaseline.scala:45-47:
 *   MeanAbsoluteErrorSpark(baselinePredictorSpark(train), test)
 * Partitions = threads; reduceByKey (P:267-268) = per-partition (count,sum) combine, merged in
 * partition order; meanSpark (P:247) = sum / count.  Used as the multi-threaded CPU baseline
 * ("local[N]" stand-in) and as a second statement of the same maths for parity (A.9 item 7).
 */
ORC_API double orc_baseline_mae_spark(const int32_t *u, const int32_t *i, const double *r, int64_t n,
                                      const int32_t *tu, const int32_t *ti, const double *tr, int64_t nt,
                                      int32_t nthreads, double *out_global_avg) {
  int32_t umax = 0, imax = 0;
  for (int64_t j = 0; j < n; ++j) { if (u[j] > umax) umax = u[j]; if (i[j] > imax) imax = i[j]; }
  size_t U = (size_t)umax + 1, I = (size_t)imax + 1;
  if (nthreads < 1) nthreads = 1;
  int P = nthreads;
  double *usum = calloc(U, sizeof(double)), *isum = calloc(I, sizeof(double));
  int32_t *ucnt = calloc(U, sizeof(int32_t)), *icnt = calloc(I, sizeof(int32_t));
  double *pus = calloc(U * (size_t)P, sizeof(double)); int32_t *puc = calloc(U * (size_t)P, sizeof(int32_t));
  double *pis = calloc(I * (size_t)P, sizeof(double)); int32_t *pic = calloc(I * (size_t)P, sizeof(int32_t));
  double *psum = calloc((size_t)P, sizeof(double));
  /* getGlobalAvg P:265 + getUsersAvg P:274 (map-side combine per partition) */
#pragma omp parallel for num_threads(P) schedule(static)
  for (int p = 0; p < P; ++p) {
    int64_t lo = n * p / P, hi = n * (p + 1) / P;
    double s = 0.0;
    for (int64_t j = lo; j < hi; ++j) { s += r[j]; pus[(size_t)p * U + u[j]] += r[j]; puc[(size_t)p * U + u[j]]++; }
    psum[p] = s;
  }
  double gs = 0.0;
  for (int p = 0; p < P; ++p) gs += psum[p];
  double gavg = n ? gs / (double)n : 0.0;
#pragma omp parallel for num_threads(P) schedule(static)
  for (int64_t k = 0; k < (int64_t)U; ++k) {
    double s = 0.0; int32_t c = 0;
    for (int p = 0; p < P; ++p) { s += pus[(size_t)p * U + k]; c += puc[(size_t)p * U + k]; }
    ucnt[k] = c; usum[k] = c ? s / (double)c : 0.0; /* usum now holds the average (P:268 x._2/x._1) */
  }
  /* getNormalizedDev P:316-329 + reduceByKeySpark on (item,(1,dev)) P:342 */
#pragma omp parallel for num_threads(P) schedule(static)
  for (int p = 0; p < P; ++p) {
    int64_t lo = n * p / P, hi = n * (p + 1) / P;
    for (int64_t j = lo; j < hi; ++j) {
      double ua = ucnt[u[j]] ? usum[u[j]] : gavg; /* P:325 */
      pis[(size_t)p * I + i[j]] += (r[j] - ua) / orc_scale(r[j], ua); /* P:327 */
      pic[(size_t)p * I + i[j]]++;
    }
  }
#pragma omp parallel for num_threads(P) schedule(static)
  for (int64_t k = 0; k < (int64_t)I; ++k) {
    double s = 0.0; int32_t c = 0;
    for (int p = 0; p < P; ++p) { s += pis[(size_t)p * I + k]; c += pic[(size_t)p * I + k]; }
    icnt[k] = c; isum[k] = c ? s / (double)c : 0.0;
  }
  /* MeanAbsoluteErrorSpark P:256-258 */
  double *perr = calloc((size_t)P, sizeof(double));
#pragma omp parallel for num_threads(P) schedule(static)
  for (int p = 0; p < P; ++p) {
    int64_t lo = nt * p / P, hi = nt * (p + 1) / P;
    double s = 0.0;
    for (int64_t j = lo; j < hi; ++j) {
      double pred;
      int32_t uu = tu[j], ii = ti[j];
      if (uu < 0 || uu > umax || !ucnt[uu]) pred = gavg;      /* P:377-378 */
      else {
        double ua = usum[uu];
        double d = (ii >= 0 && ii <= imax && icnt[ii]) ? isum[ii] : 0.0; /* P:381 */
        pred = orc_combine(ua, d);                             /* P:383 */
      }
      s += fabs(pred - tr[j]);                                  /* P:257 */
    }
    perr[p] = s;
  }
  double es = 0.0;
  for (int p = 0; p < P; ++p) es += perr[p];
  if (out_global_avg) *out_global_avg = gavg;
  free(usum); free(isum); free(ucnt); free(icnt); free(pus); free(puc); free(pis); free(pic); free(psum); free(perr);
  return es / (double)nt;
}

ORC_API int32_t orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
