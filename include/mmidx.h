/*
 * mmidx.h -- C ABI of libmmidx.so, the B200 (sm_100a) drop-in for the search / encode hot path of
 * MKLab-ITI/multimedia-indexing (package gr.iti.mklab.visual.datastructures + VladAggregator).
 *
 * Plain C, flat pointers and sizes, int status codes, no callbacks: bindable from JNI, JDK-22 FFM,
 * ctypes or cgo.  Every entry point names the reference method it replaces
 * (J/ = src/main/java/gr/iti/mklab/visual/).  INTEGRATION.md shows the Java/JNI side.
 *
 * All arithmetic follows the reference's binary64 model (one rounding per sub/mul/add, index-ascending
 * accumulation, no FMA), so PQ codes, list ids, neighbour ids AND distances are bit-identical to the
 * Java path on the same inputs.  There is no CPU fallback: every compute entry point fails with
 * MMIDX_ERR_CUDA when no sm_100 device is usable.
 *
 * Host entry points take HOST pointers and do their own H2D/D2H copies.  The *_dev entry points take
 * DEVICE pointers plus a cudaStream_t (passed as void*) and never synchronise the host unless stated.
 *
 * Threading: concurrent mmidx_search* calls on one index are allowed once it is "sealed" (the first
 * search after an add seals it); mmidx_add* are exclusive (mirrors `synchronized indexVector`,
 * AbstractSearchStructure.java:229).
 */
#ifndef MMIDX_H
#define MMIDX_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mmidx_index mmidx_t;

/* index kinds: Linear.java / PQ.java / IVFPQ.java */
enum { MMIDX_LINEAR = 0, MMIDX_PQ = 1, MMIDX_IVFPQ = 2 };

/* status codes (0 = ok).  Each maps onto a reference `throw new Exception(...)` / `return false`. */
enum {
    MMIDX_OK = 0,
    MMIDX_ERR_INVALID = 1,   /* bad argument (null pointer, k<1 == BoundedPriorityQueue ctor throw, ...) */
    MMIDX_ERR_DIM = 2,       /* "The dimensionality of the vector is wrong!" IVFPQ.java:311 PQ.java:234 Linear.java:113;
                                "The given number of subvectors is not valid!" PQ.java:148-150 */
    MMIDX_ERR_FULL = 3,      /* "Maximum index capacity reached" -> indexVector returns false, ASS.java:232-235 */
    MMIDX_ERR_STATE = 4,     /* quantizer not loaded / index not in memory ASS.java:282-284 */
    MMIDX_ERR_CUDA = 5,      /* CUDA runtime failure or no sm_100 device: the product has no CPU path */
    MMIDX_ERR_UNSUPPORTED = 6, /* outside the built limits (k > MMIDX_MAX_K, ks > 65536, ...) */
    MMIDX_ERR_W = 7          /* w < 1 (BoundedPriorityQueue ctor throws) or w > nlist (NPE at IVFPQ.java:598) */
};

#define MMIDX_MAX_K 1024

typedef struct {
    int32_t type;        /* MMIDX_LINEAR | MMIDX_PQ | MMIDX_IVFPQ */
    int32_t d;           /* vectorLength */
    int64_t max_n;       /* maxNumVectors */
    int32_t m;           /* numSubVectors (PQ, IVFPQ); d % m == 0 */
    int32_t ks;          /* numProductCentroids; <=256 -> byte codes, else short codes (PQ.java:254-266) */
    int32_t nlist;       /* numCoarseCentroids (IVFPQ) */
    int32_t w;           /* lists probed; <=0 -> reference default (int)(nlist*0.1), IVFPQ.java:188 */
    int32_t device;      /* CUDA device ordinal; <0 -> current device */
    int32_t shard_rank;  /* multi-GPU: this index keeps only IVF lists l with l % shard_count == shard_rank */
    int32_t shard_count; /* (PQ/Linear: contiguous iid ranges are the caller's job); 0 or 1 -> unsharded */
} mmidx_params;

/* ---- lifecycle: the IVFPQ / PQ / Linear constructors (IVFPQ.java:174-177, PQ.java:142-144, Linear.java:67-68) ---- */
int mmidx_create(const mmidx_params *p, mmidx_t **out);
int mmidx_destroy(mmidx_t *ix); /* closeInternal(), ASS.java:755 */

/* loadProductQuantizer PQ.java:210-223 / IVFPQ.java:275-292: P is [m][ks][d/m] row-major binary64 */
int mmidx_set_product_quantizer(mmidx_t *ix, const double *P);
/* loadCoarseQuantizer IVFPQ.java:297-300: C is [nlist][d] row-major binary64 */
int mmidx_set_coarse_quantizer(mmidx_t *ix, const double *C);
/* TransformationType.RandomPermutation (PQ.java:237-241,294-298; IVFPQ.java:319-323,420-424):
 * perm[d] as produced by RandomPermutation.java:29-40; permuted[i] = v[perm[i]].  NULL clears it. */
int mmidx_set_permutation(mmidx_t *ix, const int32_t *perm);
/* TransformationType (PQ.java:30-32, same ordinals) applied before product quantization: to the vector in PQ
 * (PQ.java:237-241, 294-298), to the residual in IVFPQ (IVFPQ.java:319-323, 420-424).
 *   MMIDX_TRANSFORM_NONE        perm, R ignored
 *   MMIDX_TRANSFORM_ROTATION    R[d][d] row-major: transformed = v R (RandomRotation.rotate, RandomRotation.java:44-49: the 1 x d
 *                               row vector times the matrix, products added for i ascending).  The matrix has to be SUPPLIED:
 *                               the reference draws it with EJML's RandomMatrices.createOrthogonal (un-vendored third-party
 *                               code), e.g. dumped once from a JVM.  Searches then run the binary64 ADC-table kernels.
 *   MMIDX_TRANSFORM_PERMUTATION perm[d] as for mmidx_set_permutation */
enum { MMIDX_TRANSFORM_NONE = 0, MMIDX_TRANSFORM_ROTATION = 1, MMIDX_TRANSFORM_PERMUTATION = 2 };
int mmidx_set_transform(mmidx_t *ix, int32_t kind, const int32_t *perm, const double *R);
/* multi-GPU: owner[nlist] = shard that stores each inverted list (default l % shard_count); every rank must be
 * given the same map, before the first vector is indexed.  NULL restores the default. */
int mmidx_set_shard_map(mmidx_t *ix, const int32_t *owner);
/* IVFPQ.setW IVFPQ.java:95-97 */
int mmidx_set_w(mmidx_t *ix, int32_t w);

/* ---- indexing: indexVectorInternal (IVFPQ.java:309-355, PQ.java:232-268, Linear.java:111-121) ----
 * Appends n vectors X[n][d]; they get iids loadCounter .. loadCounter+n-1 in order.
 * out_list[n] (IVFPQ, may be NULL) receives the coarse list id; out_codes (PQ/IVFPQ, may be NULL) the RAW
 * centroid indices, uint8[n][m] when ks<=256 else uint16[n][m] -- the Java byte is (code-128), PQ.java:555 --
 * so the caller can persist (listId, code) exactly as IVFPQ.appendPersistentIndex does (IVFPQ.java:760-772).
 * For a sharded index every rank is given the same X; vectors of foreign lists are encoded but not stored. */
int mmidx_add(mmidx_t *ix, int64_t n, const double *X, int32_t *out_list, void *out_codes);
/* indexPQCode IVFPQ.java:357-386 / loadIndexInMemory IVFPQ.java:680-728: insert pre-computed codes.
 * list_ids is ignored for PQ.  codes are raw indices (uint8 / uint16 as above). */
int mmidx_add_codes(mmidx_t *ix, int64_t n, const int32_t *list_ids, const void *codes);
/* encode only, nothing stored (the arithmetic of indexVectorInternal without the append) */
int mmidx_encode(mmidx_t *ix, int64_t n, const double *X, int32_t *out_list, void *out_codes);
/* mmidx_add with the vectors already in HBM: dX[n][d], d_out_list / d_out_codes are DEVICE pointers (may be NULL).
 * Lets a caller index a database that is produced on the device (bench.py configs[3]: 10M vectors per shard).
 * Runs on the index's own stream and returns when the vectors are stored: dX must be COMPLETE when the call is made
 * (synchronise the stream that produces it first). */
int mmidx_add_dev(mmidx_t *ix, int64_t n, const double *dX, int32_t *d_out_list, void *d_out_codes);

/* ---- search: computeNearestNeighborsInternal(k, double[]) (IVFPQ.java:408-450 computeKnnIVFADC,
 * PQ.java:290-322 computeKnnADC, Linear.java:138-163) for nq queries Q[nq][d].
 * out_iids[nq][k], out_dist[nq][k] (squared L2, ADC for PQ/IVFPQ) in BoundedPriorityQueue iteration order
 * (ascending distance, later-offered first among equal distances); out_count[nq] = min(k, #candidates)
 * (ASS.java:346); unused slots hold iid -1 / +inf. */
int mmidx_search(mmidx_t *ix, int64_t nq, const double *Q, int32_t k, int32_t *out_iids, double *out_dist,
                 int32_t *out_count);
/* same with everything already in HBM; asynchronous on `stream` */
int mmidx_search_dev(mmidx_t *ix, int64_t nq, const double *dQ, int32_t k, int32_t *d_iids, double *d_dist,
                     int32_t *d_count, void *stream);
/* ---- multi-GPU (one process per GPU; the IVF lists are sharded, l % shard_count == shard_rank) ----
 * The reference keeps ONE queue for all probed lists (IVFPQ.java:409,445).  A sharded search therefore is:
 *   0. (optional) every rank runs mmidx_coarse_probe_dev on its nq/G slice of the queries and the probe lists
 *      are all-gathered, so the coarse stage is not repeated G times;
 *   1. mmidx_search_shard_dev on every rank: local top-k of the lists this rank owns, plus for every result the
 *      offer sequence number d_seq[nq][k] (probe rank * 2^32 + position in list) and per query d_tie[nq], the
 *      distance at which locally tied candidates were cut (-1 if none);
 *   2. all-gather the four arrays (NCCL) into [nparts][nq][k] / [nparts][nq] buffers;
 *   3. mmidx_merge_topk_dev: final top-k in BoundedPriorityQueue order; d_amb_list/d_amb_count receive the
 *      queries whose k-th boundary is an exact binary64 tie that was cut (normally none);
 *   4. only if *d_amb_count != 0: mmidx_tie_collect_shard_dev on every rank, all-gather, mmidx_tie_finish_dev
 *      (replays the queue's tie rule exactly, see csrc/tie_resolve.cuh).
 * All device side, asynchronous on `stream` (step 4's collect host-syncs: it is the rare path). nq <= 32768. */
int mmidx_search_shard_dev(mmidx_t *ix, int64_t nq, const double *dQ, int32_t k,
                           const int32_t *d_probes /* [nq][w] from mmidx_coarse_probe_dev, or NULL: computed here */,
                           int32_t *d_iids, double *d_dist, int64_t *d_seq, double *d_tie, int32_t *d_count,
                           void *stream);
int mmidx_merge_topk_dev(int64_t nq, int32_t k, int32_t nparts, const int32_t *d_iids, const double *d_dist,
                         const int64_t *d_seq, const double *d_tie, const int32_t *d_count,
                         int32_t *d_out_iids, double *d_out_dist, int64_t *d_out_seq, int32_t *d_out_count,
                         int32_t *d_amb_list /*[nq]*/, int32_t *d_amb_count /*[1]*/, void *stream);
int mmidx_tie_collect_shard_dev(mmidx_t *ix, int64_t nq, const double *dQ, int32_t k, const double *d_res_dist,
                                const int32_t *d_amb_list, const int32_t *d_amb_count, int64_t *d_l_seq /*[nq][k]*/,
                                int32_t *d_l_iid /*[nq][k]*/, int32_t *d_l_eq /*[nq][k]*/, int32_t *d_l_cnt /*[nq]*/,
                                void *stream);
int mmidx_tie_finish_dev(int64_t nq, int32_t k, int32_t nparts, const int64_t *d_l_seq, const int32_t *d_l_iid,
                         const int32_t *d_l_eq, const int32_t *d_l_cnt, const int32_t *d_amb_list,
                         const int32_t *d_amb_count, int32_t *d_res_iids, double *d_res_dist, void *stream);

/* ---- multi-GPU, whole step inside the library (one process per GPU of ONE NVSwitch node, world <= 8) ----
 * The job is G = S x R ranks, rank = group * S + shard: S list shards hold one copy of the index (this index must have
 * been created with shard_count = S, shard_rank = rank % S), R groups each serve their own query batch.  The
 * exchange does not go through a collective library: every rank owns an exchange window in HBM which its peers map
 * over CUDA IPC, and the kernels that produce a row (coarse verification, fused ADC scan, merge, tie pass) store it
 * straight into the window of the rank that consumes it over NVLink; the exchange points are epoch flags in the same
 * windows (csrc/comm.cuh).  The reference's single queue for all probed lists (IVFPQ.java:409,445) is rebuilt by the
 * owner of a query slice from the S per-shard queues, including the ordered tie rule.
 *   1. every rank: mmidx_comm_create -> 128-byte handle;  2. all-gather the handles with any host-side transport;
 *   3. every rank: mmidx_comm_attach(handles[world]);     4. mmidx_search_multi_dev / mmidx_search_multi per step. */
#define MMIDX_COMM_HANDLE_BYTES 128
/* max_gq / k_max size the window: the largest group batch and k that will be searched */
int mmidx_comm_create(mmidx_t *ix, int32_t rank, int32_t world, int32_t list_shards, int64_t max_gq, int32_t k_max,
                      void *handle_out /* MMIDX_COMM_HANDLE_BYTES */);
int mmidx_comm_attach(mmidx_t *ix, const void *handles /* [world][MMIDX_COMM_HANDLE_BYTES], by rank */);
int mmidx_comm_destroy(mmidx_t *ix);
/* One step: the gq queries dQ[gq][d] of this rank's GROUP (every shard of a group passes the same queries; every rank of
 * the job must call with the same gq, k, gather_all, the same number of times).  Asynchronous on `stream`.
 * Results live in this rank's window (valid until the second next step on this stream): job-wide arrays
 * iids[R * gqp][k], dist[R * gqp][k], count[R * gqp] with gqp = S * ceil(gq / S); group g's query q is row g * gqp + q.
 * This rank produced rows [*row0, *row0 + *nrows) (its slice of its group's batch; the whole batch when S == 1);
 * gather_all != 0 additionally delivers every rank's rows to every rank (stores into all windows + one more
 * exchange point), so the arrays are complete everywhere. */
int mmidx_search_multi_dev(mmidx_t *ix, int64_t gq, const double *dQ, int32_t k, int32_t gather_all,
                           const int32_t **d_iids, const double **d_dist, const int32_t **d_count, int64_t *row0,
                           int64_t *nrows, void *stream);
/* same with HOST buffers: copies the group's queries in, runs the step, copies THIS rank's rows out
 * (queries [*first_query, *first_query + *nrows) of the group batch) and synchronises */
int mmidx_search_multi(mmidx_t *ix, int64_t gq, const double *Q, int32_t k, int32_t *out_iids, double *out_dist,
                       int32_t *out_count, int64_t *first_query, int64_t *nrows);

/* IVFPQ.computeNearestCoarseIndices IVFPQ.java:575-601: out[nq][w], ascending coarse distance */
int mmidx_coarse_probe(mmidx_t *ix, int64_t nq, const double *Q, int32_t w, int32_t *out);
/* same, device side and asynchronous: lets G ranks each probe nq/G queries and all-gather the probe lists
 * (the coarse quantizer is replicated, so the result does not depend on which rank computes it) */
int mmidx_coarse_probe_dev(mmidx_t *ix, int64_t nq, const double *dQ, int32_t w, int32_t *d_out, void *stream);
/* computeLookupADC PQ.java:387-399: out[nq][m][ks] for already-transformed (residual) vectors V[nq][d] */
int mmidx_pq_lut(mmidx_t *ix, int64_t nq, const double *V, double *out);

/* ---- introspection ---- */
int mmidx_size(mmidx_t *ix, int64_t *out);                      /* getLoadCounter ASS.java:711 */
int mmidx_list_sizes(mmidx_t *ix, int32_t *out /*[nlist]*/);    /* outputItemsPerList IVFPQ.java:654-673 */
int mmidx_get_vector(mmidx_t *ix, int64_t iid, double *out);    /* Linear.getVector Linear.java:253-281 */
/* algorithmic bytes the reference's scan touches for these queries: sum over probed lists of
 * len*(m*code_bytes+4) (IVFPQ), n*m*code_bytes (PQ), n*d*8 (Linear); used by bench.py's roofline */
int mmidx_scan_bytes(mmidx_t *ix, int64_t nq, const double *Q, int64_t *out_total);
/* device time in ms of the stages of the most recent search call on this index (CUDA events on the
 * launching stream): [0]=coarse [1]=LUT build [2]=ADC scan+top-k [3]=merge/tie [4]=whole call. Host-syncs. */
int mmidx_last_timings(mmidx_t *ix, float *out5);
/* same for a multi-GPU step, finer: [0..4] as above ([3] = everything after the scan), [5]=exchange points (flag stores
 * + waiting for the peers, i.e. their skew), [6]=merge of the per-shard queues, [7]=cross-shard tie pass kernels */
int mmidx_last_timings_multi(mmidx_t *ix, float *out8);
/* record CUDA events around the stages of later search calls (off by default; not thread-safe) */
int mmidx_enable_timings(mmidx_t *ix, int32_t on);
/* number of kernels the most recent call launched */
int mmidx_last_launches(mmidx_t *ix, int32_t *out);
/* debug counters of the fused IVFPQ scan kernel, accumulated since mmidx_create when MMIDX_STATS=1 was set (else
 * MMIDX_ERR_STATE): [0]=candidates scanned [1]=lists re-scanned after a collector overflow of the barrier-free sweep
 * [2]=survivors evaluated in binary64 [3]=queries handed to the table-free exact kernel. Host-syncs. */
int mmidx_debug_stats(mmidx_t *ix, uint64_t *out4);

/* ---- VLAD: VladAggregator.aggregateInternal VladAggregator.java:56-70 over
 * AbstractFeatureAggregator.computeNearestCentroid AFA.java:136-155, batched over images.
 * codebook[K][D]; desc[offsets[n_img]][D]; image i owns descriptors offsets[i]..offsets[i+1]-1;
 * out[n_img][K*D]; out_assign[sum n] optional nearest-centroid indices. Stateless and re-entrant. */
int mmidx_vlad(const double *codebook, int32_t K, int32_t D, int64_t n_img, const int64_t *offsets,
               const double *desc, double *out, int32_t *out_assign, int32_t device);
int mmidx_vlad_dev(const double *d_codebook, int32_t K, int32_t D, int64_t n_img, const int64_t *d_offsets,
                   int64_t n_desc, const double *d_desc, double *d_out, int32_t *d_assign, void *stream);

/* ---- the steps around the VLAD path, batched on the device ----
 * Normalization.normalizePower (signum(x) * pow(|x|, a), Normalization.java:74-79; a = 0.5 uses the correctly rounded
 * square root) and / or normalizeL2 (Normalization.java:21-37: squares added for i ascending, zero norm -> all ones) applied
 * to every row of X[rows][len] in place. */
int mmidx_normalize_rows(double *X, int64_t rows, int64_t len, int32_t do_power, double a, int32_t do_l2, int32_t device);
int mmidx_normalize_rows_dev(double *dX, int64_t rows, int64_t ld, int32_t len, int32_t do_power, double a, int32_t do_l2,
                             void *stream);
/* VladAggregatorMultipleVocabularies.aggregate (VladAggregatorMultipleVocabularies.java:84-101): nvoc codebooks stacked in
 * codebooks[sum Ks][D]; every vocabulary aggregates the same descriptors; normalize != 0: power(0.5) + L2 per sub-VLAD and,
 * for nvoc > 1, L2 of the concatenation.  out[n_img][sum_v Ks[v] * D]. */
int mmidx_vlad_multi(const double *codebooks, int32_t nvoc, const int32_t *Ks, int32_t D, int64_t n_img, const int64_t *offsets,
                     const double *desc, int32_t normalize, double *out, int32_t device);
int mmidx_vlad_multi_dev(const double *d_codebooks, int32_t nvoc, const int32_t *Ks /* host */, int32_t D, int64_t n_img,
                         const int64_t *d_offsets, int64_t n_desc, const double *d_desc, int32_t normalize, double *d_out,
                         void *stream);
/* PCA.sampleToEigenSpace (PCA.java:188-208) for n vectors: out[i] = V_t (X[i] - means), V_t[nc][ss] (for whitening the
 * caller passes diag(eigenvalue^-0.5) V_t, as PCA.loadPCAFromFile builds it, and l2_normalize = 1).  Products are added for j
 * ascending; EJML's own order is un-vendored, so parity with the Java path is claimed at 1e-4 relative. */
int mmidx_pca_project(const double *Vt, const double *means, int32_t nc, int32_t ss, int64_t n, const double *X,
                      int32_t l2_normalize, double *out, int32_t device);
int mmidx_pca_project_dev(const double *d_Vt, const double *d_means, int32_t nc, int32_t ss, int64_t n, const double *dX,
                          int32_t l2_normalize, double *d_out, void *stream);

/* thread-local message of the last failing call on this thread */
const char *mmidx_last_error(void);
/* "libmmidx <version> sm_100a" */
const char *mmidx_version(void);

#ifdef __cplusplus
}
#endif
#endif
