/*
 * mmidx_oracle.c -- TEST INFRASTRUCTURE ONLY (see mmidx_oracle.h).
 *
 * Every function restates one reference method in the reference's own operation
 * order: binary64, one rounding per sub / mul / add, no fused multiply-add
 * (build with -ffp-contract=off), early exits kept where the Java has them.
 * PARITY UNPINNED by the reference (it has no tests); see header.
 */
#include "mmidx_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------
 * com.aliasi.util.BoundedPriorityQueue<Result> with comparator J/utilities/Result.java:38-45
 * (smaller distance == "larger" element).  LingPipe 4.0.1 semantics:
 *   - TreeSet of entries; every entry gets an increasing sequence id when created;
 *   - set order: comparator-larger first (ascending distance); equal-by-comparator entries
 *     are ordered by sequence id with the EARLIER entry sorting LATER;
 *   - offer(e): size<max -> insert.  Otherwise w=last(); if compare(e,w)<=0 (e.dist >= w.dist)
 *     reject; else insert e and remove w.
 * Kept as an array sorted in iteration order: (dist asc, seq desc).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    double dist;
    int64_t seq;
    int32_t id;
} orc_entry;

struct orc_bpq {
    orc_entry *e;
    int size, max_size;
    int64_t next_seq;
};

orc_bpq *orc_bpq_new(int max_size) {
    if (max_size < 1) return NULL; /* LingPipe ctor throws IllegalArgumentException */
    orc_bpq *q = (orc_bpq *)malloc(sizeof(orc_bpq));
    q->e = (orc_entry *)malloc(sizeof(orc_entry) * (size_t)(max_size + 1));
    q->size = 0;
    q->max_size = max_size;
    q->next_seq = 0;
    return q;
}

void orc_bpq_free(orc_bpq *q) {
    if (!q) return;
    free(q->e);
    free(q);
}

int orc_bpq_size(const orc_bpq *q) { return q->size; }

double orc_bpq_last_distance(const orc_bpq *q) { return q->e[q->size - 1].dist; }

/* position of a NEW entry (largest seq so far): before every entry with dist >= its dist */
static int bpq_insert_pos(const orc_bpq *q, double dist) {
    int lo = 0, hi = q->size;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (q->e[mid].dist < dist)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

int orc_bpq_offer(orc_bpq *q, int id, double dist) {
    if (q->size >= q->max_size) {
        /* Result.compare(e, last) <= 0  <=>  !(e.dist < last.dist) */
        if (!(dist < q->e[q->size - 1].dist)) return 0;
    }
    int pos = bpq_insert_pos(q, dist);
    memmove(&q->e[pos + 1], &q->e[pos], sizeof(orc_entry) * (size_t)(q->size - pos));
    q->e[pos].dist = dist;
    q->e[pos].id = id;
    q->e[pos].seq = q->next_seq++;
    q->size++;
    if (q->size > q->max_size) q->size--; /* remove(last) */
    return 1;
}

void orc_bpq_to_arrays(const orc_bpq *q, int32_t *ids, double *dists) {
    for (int i = 0; i < q->size; i++) {
        if (ids) ids[i] = q->e[i].id;
        if (dists) dists[i] = q->e[i].dist;
    }
}

/* ------------------------------------------------------------------------------------------
 * Linear.computeNearestNeighborsInternal  J/datastructures/Linear.java:138-163
 * ------------------------------------------------------------------------------------------ */
int orc_linear_search(const double *X, int64_t n, int d, const double *q, int k, int32_t *out_ids,
                      double *out_dist) {
    orc_bpq *nn = orc_bpq_new(k);
    if (!nn) return -1;
    double lowest = DBL_MAX;
    for (int64_t i = 0; i < n; i++) {
        int skip = 0;
        const double *x = X + i * (int64_t)d;
        double l2 = 0;
        for (int j = 0; j < d; j++) {
            double a = q[j] - x[j];
            l2 += a * a;
            if (l2 > lowest) {
                skip = 1;
                break;
            }
        }
        if (!skip) {
            orc_bpq_offer(nn, (int)i, l2);
            if (i >= k) lowest = orc_bpq_last_distance(nn);
        }
    }
    int cnt = nn->size;
    orc_bpq_to_arrays(nn, out_ids, out_dist);
    orc_bpq_free(nn);
    return cnt;
}

/* ------------------------------------------------------------------------------------------
 * PQ
 * ------------------------------------------------------------------------------------------ */
/* PQ.computeLookupADC PQ.java:387-399 / IVFPQ.java:525-538 */
void orc_pq_lut(const double *P, int m, int ks, int S, const double *v, double *lut) {
    for (int i = 0; i < m; i++) {
        int start = i * S;
        for (int j = 0; j < ks; j++) {
            const double *c = P + ((int64_t)i * ks + j) * S;
            double acc = 0;
            for (int k = 0; k < S; k++) {
                double a = v[start + k] - c[k];
                acc += a * a;
            }
            lut[(int64_t)i * ks + j] = acc;
        }
    }
}

/* PQ.computeNearestProductIndex PQ.java:411-429 / IVFPQ.java:613-631 */
int orc_pq_nearest_product_index(const double *P, int ks, int S, int sub, const double *subvec) {
    int best = -1;
    double min_d = DBL_MAX;
    for (int i = 0; i < ks; i++) {
        const double *c = P + ((int64_t)sub * ks + i) * S;
        double dist = 0;
        for (int j = 0; j < S; j++) {
            double a = c[j] - subvec[j];
            dist += a * a;
            if (dist >= min_d) break;
        }
        if (dist < min_d) {
            min_d = dist;
            best = i;
        }
    }
    return best;
}

/* RandomRotation.rotate RandomRotation.java:44-49: CommonOps.mult(original[1 x d], randomMatrix[d x d], transformed):
 * transformed[j] = sum_i v[i] * R[i][j], i ascending, product rounded then added (Java has no fused multiply-add).
 * EJML is un-vendored third-party code: the accumulation order is the natural row-vector-times-matrix one, unverified. */
void orc_apply_rotation(const double *R, int d, const double *v, double *out) {
    for (int j = 0; j < d; j++) {
        double acc = 0;
        for (int i = 0; i < d; i++) {
            double p = v[i] * R[(size_t)i * d + j];
            acc += p;
        }
        out[j] = acc;
    }
}

/* The transformation in force for the calls below: TransformationType is ONE of None / RandomRotation / RandomPermutation
 * (PQ.java:30-32; `if RandomRotation ... else if RandomPermutation`, PQ.java:237-241).  A rotation is installed with
 * orc_set_rotation (test infrastructure: a process-wide setting, not thread-safe to change during a batch call). */
static const double *g_rot = NULL;
static int g_rot_d = 0;
void orc_set_rotation(const double *R, int d) {
    g_rot = R;
    g_rot_d = d;
}

/* BASELINE.md variant (A): pay the reference's per-candidate allocations in orc_ivfpq_search (timing only) */
static int g_faithful_costs = 0;
void orc_set_faithful_costs(int on) { g_faithful_costs = on; }

/* RandomPermutation.permute RandomPermutation.java:50-56: permuted[i] = vector[perm[i]] */
static void apply_perm(const int32_t *perm, const double *v, int d, double *out) {
    if (g_rot && g_rot_d == d) {
        orc_apply_rotation(g_rot, d, v, out);
    } else if (perm) {
        for (int i = 0; i < d; i++) out[i] = v[perm[i]];
    } else {
        memcpy(out, v, sizeof(double) * (size_t)d);
    }
}

/* PCA.sampleToEigenSpace PCA.java:188-208: CommonOps.sub(sample, means, sample); CommonOps.mult(V_t, sample, projected)
 * (matrix times vector: total += V_t[i][j] * sample[j], j ascending); with whitening the result is L2-normalised
 * (Normalization.normalizeL2).  V_t is the matrix PCA.loadPCAFromFile leaves in place (whitening folded in, :281-300). */
void orc_normalize_l2(double *v, int64_t n);
void orc_pca_project(const double *Vt, const double *means, int nc, int ss, const double *x, int l2, double *out) {
    double *s = (double *)malloc(sizeof(double) * (size_t)ss);
    for (int j = 0; j < ss; j++) s[j] = x[j] - means[j];
    for (int i = 0; i < nc; i++) {
        double total = 0;
        for (int j = 0; j < ss; j++) {
            double p = Vt[(size_t)i * ss + j] * s[j];
            total += p;
        }
        out[i] = total;
    }
    free(s);
    if (l2) orc_normalize_l2(out, nc);
}

/* PQ.indexVectorInternal PQ.java:232-268 (transform then per-sub-vector argmin) */
void orc_pq_encode(const double *P, int m, int ks, int S, const int32_t *perm, const double *x, int32_t *code) {
    int d = m * S;
    double *v = (double *)malloc(sizeof(double) * (size_t)d);
    apply_perm(perm, x, d, v);
    for (int i = 0; i < m; i++) code[i] = orc_pq_nearest_product_index(P, ks, S, i, v + i * S);
    free(v);
}

static inline int code_at(const void *codes, int ks, int64_t idx) {
    return ks <= 256 ? (int)((const uint8_t *)codes)[idx] : (int)((const uint16_t *)codes)[idx];
}

/* PQ.computeKnnADC PQ.java:290-322 */
int orc_pq_search(const double *P, int m, int ks, int S, const int32_t *perm, const void *codes, int64_t n,
                  const double *q, int k, int32_t *out_ids, double *out_dist) {
    orc_bpq *nn = orc_bpq_new(k);
    if (!nn) return -1;
    int d = m * S;
    double *v = (double *)malloc(sizeof(double) * (size_t)d);
    double *lut = (double *)malloc(sizeof(double) * (size_t)m * ks);
    apply_perm(perm, q, d, v);
    orc_pq_lut(P, m, ks, S, v, lut);
    for (int64_t i = 0; i < n; i++) {
        double l2 = 0;
        int64_t start = i * m;
        for (int j = 0; j < m; j++) l2 += lut[(int64_t)j * ks + code_at(codes, ks, start + j)];
        orc_bpq_offer(nn, (int)i, l2);
    }
    int cnt = nn->size;
    orc_bpq_to_arrays(nn, out_ids, out_dist);
    orc_bpq_free(nn);
    free(v);
    free(lut);
    return cnt;
}

/* ------------------------------------------------------------------------------------------
 * IVFPQ
 * ------------------------------------------------------------------------------------------ */
/* IVFPQ.computeNearestCoarseIndex IVFPQ.java:547-564 */
int orc_coarse_nearest(const double *C, int nlist, int d, const double *v) {
    int best = -1;
    double min_d = DBL_MAX;
    for (int i = 0; i < nlist; i++) {
        const double *c = C + (int64_t)i * d;
        double dist = 0;
        for (int j = 0; j < d; j++) {
            double a = c[j] - v[j];
            dist += a * a;
            if (dist >= min_d) break;
        }
        if (dist < min_d) {
            min_d = dist;
            best = i;
        }
    }
    return best;
}

/* IVFPQ.computeNearestCoarseIndices IVFPQ.java:575-601.  Requires 1 <= w <= nlist
 * (the Java throws for w=0 and NPEs for w>nlist). */
void orc_coarse_topw(const double *C, int nlist, int d, const double *v, int w, int32_t *out) {
    orc_bpq *bpq = orc_bpq_new(w);
    double lowest = DBL_MAX;
    for (int i = 0; i < nlist; i++) {
        int skip = 0;
        const double *c = C + (int64_t)i * d;
        double l2 = 0;
        for (int j = 0; j < d; j++) {
            double a = c[j] - v[j];
            l2 += a * a;
            if (l2 > lowest) {
                skip = 1;
                break;
            }
        }
        if (!skip) {
            orc_bpq_offer(bpq, i, l2);
            if (i >= w) lowest = orc_bpq_last_distance(bpq);
        }
    }
    orc_bpq_to_arrays(bpq, out, NULL); /* poll() x w == iteration order */
    orc_bpq_free(bpq);
}

/* IVFPQ.indexVectorInternal IVFPQ.java:309-355; returns the list id, writes raw code[m] */
int orc_ivfpq_encode(const double *C, int nlist, int d, const double *P, int m, int ks, const int32_t *perm,
                     const double *x, int32_t *code) {
    int S = d / m;
    int l = orc_coarse_nearest(C, nlist, d, x);
    double *r = (double *)malloc(sizeof(double) * (size_t)d);
    double *v = (double *)malloc(sizeof(double) * (size_t)d);
    const double *c = C + (int64_t)l * d;
    for (int i = 0; i < d; i++) r[i] = c[i] - x[i]; /* computeResidualVector :642-648 */
    apply_perm(perm, r, d, v);
    for (int i = 0; i < m; i++) code[i] = orc_pq_nearest_product_index(P, ks, S, i, v + i * S);
    free(r);
    free(v);
    return l;
}

/* IVFPQ.computeKnnIVFADC IVFPQ.java:408-450 */
int orc_ivfpq_search(const double *C, int nlist, int d, const double *P, int m, int ks, const int32_t *perm,
                     const int64_t *list_off, const void *codes, const int32_t *iids, const double *q, int k,
                     int w, int32_t *out_ids, double *out_dist) {
    orc_bpq *nn = orc_bpq_new(k);
    if (!nn) return -1;
    int S = d / m;
    int32_t *probes = (int32_t *)malloc(sizeof(int32_t) * (size_t)w);
    double *r = (double *)malloc(sizeof(double) * (size_t)d);
    double *v = (double *)malloc(sizeof(double) * (size_t)d);
    double *lut = (double *)malloc(sizeof(double) * (size_t)m * ks);
    orc_coarse_topw(C, nlist, d, q, w, probes);
    const int cbytes = ks <= 256 ? m : 2 * m;
    for (int i = 0; i < w; i++) {
        int l = probes[i];
        const double *c = C + (int64_t)l * d;
        for (int t = 0; t < d; t++) r[t] = c[t] - q[t];
        apply_perm(perm, r, d, v);
        orc_pq_lut(P, m, ks, S, v, lut);
        for (int64_t pos = list_off[l]; pos < list_off[l + 1]; pos++) {
            double l2 = 0;
            int64_t start = pos * m;
            if (g_faithful_costs) {
                /* the reference's cost structure per candidate: pqByteCodes[l].toArray(start, m) copies the code into a
                 * fresh array (IVFPQ.java:434), `new Result(iid, dist)` (:443-444), and an accepted offer allocates a
                 * TreeSet entry while the evicted one becomes garbage.  Same arithmetic, same result. */
                unsigned char *copy = (unsigned char *)malloc((size_t)cbytes);
                memcpy(copy, (const unsigned char *)codes + pos * cbytes, (size_t)cbytes);
                for (int j = 0; j < m; j++) l2 += lut[(int64_t)j * ks + code_at(copy, ks, j)];
                volatile double *res = (volatile double *)malloc(2 * sizeof(double));
                res[0] = (double)iids[pos];
                res[1] = l2;
                if (orc_bpq_offer(nn, iids[pos], res[1])) {
                    void *node = malloc(48);
                    *(volatile char *)node = 0;
                    free(node);
                }
                free((void *)res);
                free(copy);
                continue;
            }
            for (int j = 0; j < m; j++) l2 += lut[(int64_t)j * ks + code_at(codes, ks, start + j)];
            orc_bpq_offer(nn, iids[pos], l2);
        }
    }
    int cnt = nn->size;
    orc_bpq_to_arrays(nn, out_ids, out_dist);
    orc_bpq_free(nn);
    free(probes);
    free(r);
    free(v);
    free(lut);
    return cnt;
}

/* ------------------------------------------------------------------------------------------
 * VLAD
 * ------------------------------------------------------------------------------------------ */
/* AbstractFeatureAggregator.computeNearestCentroid AFA.java:136-155 */
int orc_nearest_centroid(const double *codebook, int K, int D, const double *desc) {
    int best = -1;
    double min_d = DBL_MAX;
    for (int i = 0; i < K; i++) {
        const double *c = codebook + (int64_t)i * D;
        double dist = 0;
        for (int j = 0; j < D; j++) {
            double a = c[j] - desc[j];
            dist += a * a;
            if (dist >= min_d) break;
        }
        if (dist < min_d) {
            min_d = dist;
            best = i;
        }
    }
    return best;
}

/* VladAggregator.aggregateInternal VladAggregator.java:56-70 */
void orc_vlad(const double *codebook, int K, int D, const double *desc, int64_t n, double *out,
              int32_t *out_assign) {
    memset(out, 0, sizeof(double) * (size_t)K * D);
    for (int64_t t = 0; t < n; t++) {
        const double *x = desc + t * D;
        int nn = orc_nearest_centroid(codebook, K, D, x);
        if (out_assign) out_assign[t] = nn;
        const double *c = codebook + (int64_t)nn * D;
        double *o = out + (int64_t)nn * D;
        for (int i = 0; i < D; i++) o[i] += x[i] - c[i];
    }
}

/* Normalization.normalizeL2 Normalization.java:21-37 */
void orc_normalize_l2(double *v, int64_t n) {
    double norm2 = 0;
    for (int64_t i = 0; i < n; i++) norm2 += v[i] * v[i];
    norm2 = sqrt(norm2);
    if (norm2 == 0) {
        for (int64_t i = 0; i < n; i++) v[i] = 1;
    } else {
        for (int64_t i = 0; i < n; i++) v[i] = v[i] / norm2;
    }
}

/* Normalization.normalizePower Normalization.java:74-79 (a = 0.5 uses sqrt: Math.pow(x,0.5) == sqrt(x)
 * for every finite non-negative x except that pow is allowed 1 ulp; StrictMath/fdlibm pow(x,0.5) is
 * correctly rounded for these inputs in practice -- "next" row f3, tolerance 1e-15 in tests). */
void orc_normalize_power(double *v, int64_t n, double a) {
    for (int64_t i = 0; i < n; i++) {
        double s = (v[i] > 0) - (v[i] < 0);
        v[i] = s * pow(fabs(v[i]), a);
    }
}

/* ------------------------------------------------------------------------------------------
 * RandomPermutation ctor RandomPermutation.java:29-40: java.util.Random(seed) (48-bit LCG) +
 * Collections.shuffle(list, rnd): for (i = size; i > 1; i--) swap(list, i-1, rnd.nextInt(i)).
 * ------------------------------------------------------------------------------------------ */
static int32_t jrand_next(uint64_t *s, int bits) {
    *s = (*s * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
    return (int32_t)((int64_t)(*s) >> (48 - bits));
}

static int32_t jrand_next_int(uint64_t *s, int32_t bound) {
    int32_t r = jrand_next(s, 31);
    int32_t m = bound - 1;
    if ((bound & m) == 0) {
        r = (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
    } else {
        for (int32_t u = r; (int32_t)((uint32_t)u - (uint32_t)(r = u % bound) + (uint32_t)m) < 0;
             u = jrand_next(s, 31))
            ;
    }
    return r;
}

void orc_random_permutation(int seed, int dim, int32_t *out) {
    uint64_t s = ((uint64_t)(int64_t)seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1);
    for (int i = 0; i < dim; i++) out[i] = i;
    for (int i = dim; i > 1; i--) {
        int j = jrand_next_int(&s, i);
        int32_t t = out[i - 1];
        out[i - 1] = out[j];
        out[j] = t;
    }
}

/* ------------------------------------------------------------------------------------------
 * Batch helpers: N caller threads on one shared read-only index, the way several Java threads
 * may call the unsynchronised computeNearestNeighbors (AbstractSearchStructure.java:281).
 * Plain pthreads, dynamic chunking over items.
 * ------------------------------------------------------------------------------------------ */
int orc_num_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef void (*item_fn)(int64_t i, void *ctx);
typedef struct {
    item_fn fn;
    void *ctx;
    int64_t n, chunk;
    atomic_llong next;
} pf_job;

static void *pf_worker(void *arg) {
    pf_job *job = (pf_job *)arg;
    for (;;) {
        int64_t b = atomic_fetch_add(&job->next, job->chunk);
        if (b >= job->n) break;
        int64_t e = b + job->chunk < job->n ? b + job->chunk : job->n;
        for (int64_t i = b; i < e; i++) job->fn(i, job->ctx);
    }
    return NULL;
}

static void parallel_for(int64_t n, int64_t chunk, int nthreads, item_fn fn, void *ctx) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    pf_job job;
    job.fn = fn;
    job.ctx = ctx;
    job.n = n;
    job.chunk = chunk < 1 ? 1 : chunk;
    atomic_init(&job.next, 0);
    if (nthreads == 1 || n <= 1) {
        pf_worker(&job);
        return;
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    int started = 0;
    for (int t = 0; t < nthreads - 1; t++)
        if (pthread_create(&th[started], NULL, pf_worker, &job) == 0) started++;
    pf_worker(&job);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    free(th);
}

static void fill_tail(int32_t *ids, double *dist, int cnt, int k) {
    for (int i = cnt < 0 ? 0 : cnt; i < k; i++) {
        ids[i] = -1;
        dist[i] = INFINITY;
    }
}

typedef struct {
    const double *C, *P, *Q, *X, *codebook, *desc;
    const int32_t *perm, *iids;
    const int64_t *list_off, *offsets;
    const void *codes;
    int nlist, d, m, ks, S, k, w, K, D;
    int64_t n;
    int32_t *out_ids, *out_count, *out_list, *out_codes, *out_assign;
    double *out_dist, *out;
} bctx;

static void it_ivfpq_search(int64_t i, void *p) {
    bctx *b = (bctx *)p;
    int c = orc_ivfpq_search(b->C, b->nlist, b->d, b->P, b->m, b->ks, b->perm, b->list_off, b->codes, b->iids,
                             b->Q + i * b->d, b->k, b->w, b->out_ids + i * b->k, b->out_dist + i * b->k);
    fill_tail(b->out_ids + i * b->k, b->out_dist + i * b->k, c, b->k);
    if (b->out_count) b->out_count[i] = c;
}

void orc_ivfpq_search_batch(const double *C, int nlist, int d, const double *P, int m, int ks,
                            const int32_t *perm, const int64_t *list_off, const void *codes,
                            const int32_t *iids, const double *Q, int64_t nq, int k, int w, int32_t *out_ids,
                            double *out_dist, int32_t *out_count, int nthreads) {
    bctx b = {0};
    b.C = C; b.nlist = nlist; b.d = d; b.P = P; b.m = m; b.ks = ks; b.perm = perm; b.list_off = list_off;
    b.codes = codes; b.iids = iids; b.Q = Q; b.k = k; b.w = w; b.out_ids = out_ids; b.out_dist = out_dist;
    b.out_count = out_count;
    parallel_for(nq, 4, nthreads, it_ivfpq_search, &b);
}

static void it_pq_search(int64_t i, void *p) {
    bctx *b = (bctx *)p;
    int c = orc_pq_search(b->P, b->m, b->ks, b->S, b->perm, b->codes, b->n, b->Q + i * b->d, b->k,
                          b->out_ids + i * b->k, b->out_dist + i * b->k);
    fill_tail(b->out_ids + i * b->k, b->out_dist + i * b->k, c, b->k);
    if (b->out_count) b->out_count[i] = c;
}

void orc_pq_search_batch(const double *P, int m, int ks, int S, const int32_t *perm, const void *codes,
                         int64_t n, const double *Q, int64_t nq, int k, int32_t *out_ids, double *out_dist,
                         int32_t *out_count, int nthreads) {
    bctx b = {0};
    b.P = P; b.m = m; b.ks = ks; b.S = S; b.d = m * S; b.perm = perm; b.codes = codes; b.n = n; b.Q = Q;
    b.k = k; b.out_ids = out_ids; b.out_dist = out_dist; b.out_count = out_count;
    parallel_for(nq, 1, nthreads, it_pq_search, &b);
}

static void it_linear_search(int64_t i, void *p) {
    bctx *b = (bctx *)p;
    int c = orc_linear_search(b->X, b->n, b->d, b->Q + i * b->d, b->k, b->out_ids + i * b->k,
                              b->out_dist + i * b->k);
    fill_tail(b->out_ids + i * b->k, b->out_dist + i * b->k, c, b->k);
    if (b->out_count) b->out_count[i] = c;
}

void orc_linear_search_batch(const double *X, int64_t n, int d, const double *Q, int64_t nq, int k,
                             int32_t *out_ids, double *out_dist, int32_t *out_count, int nthreads) {
    bctx b = {0};
    b.X = X; b.n = n; b.d = d; b.Q = Q; b.k = k; b.out_ids = out_ids; b.out_dist = out_dist;
    b.out_count = out_count;
    parallel_for(nq, 1, nthreads, it_linear_search, &b);
}

static void it_ivfpq_encode(int64_t i, void *p) {
    bctx *b = (bctx *)p;
    b->out_list[i] = orc_ivfpq_encode(b->C, b->nlist, b->d, b->P, b->m, b->ks, b->perm, b->X + i * b->d,
                                      b->out_codes + i * b->m);
}

void orc_ivfpq_encode_batch(const double *C, int nlist, int d, const double *P, int m, int ks,
                            const int32_t *perm, const double *X, int64_t n, int32_t *out_list,
                            int32_t *out_codes, int nthreads) {
    bctx b = {0};
    b.C = C; b.nlist = nlist; b.d = d; b.P = P; b.m = m; b.ks = ks; b.perm = perm; b.X = X;
    b.out_list = out_list; b.out_codes = out_codes;
    parallel_for(n, 64, nthreads, it_ivfpq_encode, &b);
}

static void it_pq_encode(int64_t i, void *p) {
    bctx *b = (bctx *)p;
    orc_pq_encode(b->P, b->m, b->ks, b->S, b->perm, b->X + i * b->d, b->out_codes + i * b->m);
}

void orc_pq_encode_batch(const double *P, int m, int ks, int S, const int32_t *perm, const double *X,
                         int64_t n, int32_t *out_codes, int nthreads) {
    bctx b = {0};
    b.P = P; b.m = m; b.ks = ks; b.S = S; b.d = m * S; b.perm = perm; b.X = X; b.out_codes = out_codes;
    parallel_for(n, 64, nthreads, it_pq_encode, &b);
}

static void it_vlad(int64_t i, void *p) {
    bctx *b = (bctx *)p;
    orc_vlad(b->codebook, b->K, b->D, b->desc + b->offsets[i] * b->D, b->offsets[i + 1] - b->offsets[i],
             b->out + i * (int64_t)b->K * b->D, b->out_assign ? b->out_assign + b->offsets[i] : NULL);
}

void orc_vlad_batch(const double *codebook, int K, int D, const double *desc, const int64_t *offsets,
                    int64_t n_img, double *out, int32_t *out_assign, int nthreads) {
    bctx b = {0};
    b.codebook = codebook; b.K = K; b.D = D; b.desc = desc; b.offsets = offsets; b.out = out;
    b.out_assign = out_assign;
    parallel_for(n_img, 4, nthreads, it_vlad, &b);
}
