/*
 * mmidx_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, IEEE binary64, no FMA contraction) of the search /
 * encode arithmetic of MKLab-ITI/multimedia-indexing, package
 * gr.iti.mklab.visual.datastructures (Linear, PQ, IVFPQ) and
 * gr.iti.mklab.visual.aggregation (VladAggregator).
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for
 * this path and cannot be built here (no JVM, no third-party jars).  The oracle
 * is pinned only by (a) hand-derived known-answer tests and (b) an independent
 * pure-Python restatement (tests/pyref.py) whose outputs are committed under
 * tests/golden/.  The LingPipe 4.0.1 BoundedPriorityQueue tie rules are
 * restated from the published library behaviour.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (libmmidx.so)
 * never links or calls it.
 *
 * Reference citations use J/ = src/main/java/gr/iti/mklab/visual/.
 */
#ifndef MMIDX_ORACLE_H
#define MMIDX_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- BoundedPriorityQueue<Result> (LingPipe 4.0.1 + J/utilities/Result.java:38-45) ---- */
typedef struct orc_bpq orc_bpq;
orc_bpq *orc_bpq_new(int max_size);
void orc_bpq_free(orc_bpq *q);
int orc_bpq_offer(orc_bpq *q, int id, double dist); /* 1 = accepted */
int orc_bpq_size(const orc_bpq *q);
double orc_bpq_last_distance(const orc_bpq *q);
/* iteration order (== toArray / successive poll()): ascending distance, later-offered first among ties */
void orc_bpq_to_arrays(const orc_bpq *q, int32_t *ids, double *dists);

/* ---- Linear: J/datastructures/Linear.java:138-163 ---- */
int orc_linear_search(const double *X, int64_t n, int d, const double *q, int k,
                      int32_t *out_ids, double *out_dist);

/* ---- PQ: J/datastructures/PQ.java ---- */
/* computeLookupADC PQ.java:387-399 (== IVFPQ.java:525-538). P is [m][ks][S] row-major, lut is [m][ks]. */
void orc_pq_lut(const double *P, int m, int ks, int S, const double *v, double *lut);
/* computeNearestProductIndex PQ.java:411-429 (== IVFPQ.java:613-631) */
int orc_pq_nearest_product_index(const double *P, int ks, int S, int sub, const double *subvec);
/* indexVectorInternal PQ.java:232-268: code[m] are raw centroid indices 0..ks-1 (Java stores code-128 as a byte) */
void orc_pq_encode(const double *P, int m, int ks, int S, const int32_t *perm, const double *x, int32_t *code);
/* computeKnnADC PQ.java:290-322; codes are raw indices, uint8 when ks<=256 else uint16 */
int orc_pq_search(const double *P, int m, int ks, int S, const int32_t *perm, const void *codes, int64_t n,
                  const double *q, int k, int32_t *out_ids, double *out_dist);

/* ---- IVFPQ: J/datastructures/IVFPQ.java ---- */
int orc_coarse_nearest(const double *C, int nlist, int d, const double *v);                 /* :547-564 */
void orc_coarse_topw(const double *C, int nlist, int d, const double *v, int w, int32_t *out); /* :575-601 */
/* indexVectorInternal :309-355 (coarse assign, residual = centroid - vector :642-648, transform, PQ encode) */
int orc_ivfpq_encode(const double *C, int nlist, int d, const double *P, int m, int ks,
                     const int32_t *perm, const double *x, int32_t *code);
/* computeKnnIVFADC :408-450. Lists in CSR form: list l holds positions [list_off[l], list_off[l+1]) of
 * codes (row-major [pos][m]) and iids, in insertion order. */
int orc_ivfpq_search(const double *C, int nlist, int d, const double *P, int m, int ks,
                     const int32_t *perm, const int64_t *list_off, const void *codes,
                     const int32_t *iids, const double *q, int k, int w,
                     int32_t *out_ids, double *out_dist);

/* ---- VLAD: J/aggregation/AbstractFeatureAggregator.java:136-155, VladAggregator.java:56-70 ---- */
int orc_nearest_centroid(const double *codebook, int K, int D, const double *desc);
void orc_vlad(const double *codebook, int K, int D, const double *desc, int64_t n, double *out /*[K*D]*/,
              int32_t *out_assign /* [n] or NULL */);
/* Normalization.java:21-37 (L2; zero vector -> all ones) and :74-79 (signed power) */
/* RandomRotation.rotate RandomRotation.java:44-49; orc_set_rotation installs / clears (NULL) the rotation that the
 * encode / search functions apply instead of `perm` */
void orc_apply_rotation(const double *R, int d, const double *v, double *out);
void orc_set_rotation(const double *R, int d);
/* BASELINE.md variant (A) of the CPU baseline: orc_ivfpq_search additionally pays the reference's per-candidate
 * allocations (code copy IVFPQ.java:434, `new Result` :443-444, queue entry per accepted offer).  Results unchanged. */
void orc_set_faithful_costs(int on);
/* PCA.sampleToEigenSpace PCA.java:188-208 */
void orc_pca_project(const double *Vt, const double *means, int nc, int ss, const double *x, int l2, double *out);
void orc_normalize_l2(double *v, int64_t n);
void orc_normalize_power(double *v, int64_t n, double a);

/* ---- RandomPermutation: J/utilities/RandomPermutation.java:29-40 (java.util.Random + Collections.shuffle) ---- */
void orc_random_permutation(int seed, int dim, int32_t *out);

/* ---- batch helpers (query-level threads over a shared read-only index; used as the CPU baseline) ---- */
void orc_ivfpq_search_batch(const double *C, int nlist, int d, const double *P, int m, int ks,
                            const int32_t *perm, const int64_t *list_off, const void *codes,
                            const int32_t *iids, const double *Q, int64_t nq, int k, int w,
                            int32_t *out_ids, double *out_dist, int32_t *out_count, int nthreads);
void orc_pq_search_batch(const double *P, int m, int ks, int S, const int32_t *perm, const void *codes,
                         int64_t n, const double *Q, int64_t nq, int k, int32_t *out_ids, double *out_dist,
                         int32_t *out_count, int nthreads);
void orc_linear_search_batch(const double *X, int64_t n, int d, const double *Q, int64_t nq, int k,
                             int32_t *out_ids, double *out_dist, int32_t *out_count, int nthreads);
void orc_ivfpq_encode_batch(const double *C, int nlist, int d, const double *P, int m, int ks,
                            const int32_t *perm, const double *X, int64_t n, int32_t *out_list,
                            int32_t *out_codes, int nthreads);
void orc_pq_encode_batch(const double *P, int m, int ks, int S, const int32_t *perm, const double *X,
                         int64_t n, int32_t *out_codes, int nthreads);
void orc_vlad_batch(const double *codebook, int K, int D, const double *desc, const int64_t *offsets,
                    int64_t n_img, double *out, int32_t *out_assign, int nthreads);
int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
