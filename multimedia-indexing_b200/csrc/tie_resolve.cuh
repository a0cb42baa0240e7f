// tie_resolve.cuh -- ordered tie pass: exact BoundedPriorityQueue membership when binary64 distances tie
// AT the k-th boundary (LingPipe 4.0.1 BoundedPriorityQueue as used at IVFPQ.java:409,445 PQ.java:291,318
// Linear.java:140,156 IVFPQ.java:576,590; semantics in SURVEY.md A.2).
//
// Let T be the final k-th smallest distance of a query and call an entry "le" when dist <= T, "eq" when
// dist == T.  Replaying offer() in offer order (seq ascending) gives:
//   * while fewer than k le-entries have been offered, every eq-entry is accepted (worst > T or not full);
//   * from the offer t* that brings the number of le-entries to k, the queue is full with worst == T:
//     later eq-entries are rejected (compare(e, last) <= 0), and every later entry with dist < T evicts
//     last() == the EARLIEST-offered eq-entry still in the queue.
// Hence the final queue = all entries with dist < T (nless of them) + the (k - nless) LATEST-offered of the
// eq-entries with seq <= t*.  The first k le-entries in offer order contain exactly the le-entries with
// seq <= t*, so it is enough to collect those (k_tie_collect_*), merge them over shards by seq and keep the
// latest eq-entries (k_tie_finish).  Only queries the collectors flagged ambiguous take this path.
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace mmidx {

// first-k le-entries of each ambiguous query, offer order.  Indexed by query id.
struct TieLists {
    unsigned long long *seq;  // [nq][k]
    int32_t *pay;             // [nq][k]
    int32_t *eq;              // [nq][k]  1 when dist == T
    int32_t *cnt;             // [nq]
};

// block-wide exclusive scan of a 0/1 flag; returns this thread's offset, *total = block sum.  256 threads.
__device__ __forceinline__ int block_scan_flag(bool f, int *warp_sums, int *total) {
    const unsigned b = __ballot_sync(0xffffffffu, f);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();  // warp_sums reuse
    if (lane == 0) warp_sums[warp] = __popc(b);
    __syncthreads();
    int off = 0, tot = 0;
#pragma unroll
    for (int wv = 0; wv < MMIDX_NT / 32; ++wv) {
        int v = warp_sums[wv];
        if (wv < warp) off += v;
        tot += v;
    }
    *total = tot;
    return off + __popc(b & ((1u << lane) - 1u));
}

// One ordered sweep over `len` candidates of one segment.  dist_of(i) must reproduce the main scan's bits.
template <typename DistFn, typename PayFn>
__device__ __forceinline__ void tie_sweep(int64_t len, unsigned long long seq_base, double T, int k, int &found,
                                          int *warp_sums, const TieLists &o, int64_t q, DistFn dist_of, PayFn pay_of) {
    for (int64_t base = 0; base < len && found < k; base += MMIDX_NT) {
        const int64_t i = base + threadIdx.x;
        const bool valid = i < len;
        double dv = valid ? dist_of(i) : 0.0;
        const bool le = valid && dv <= T;
        int total;
        const int pos = found + block_scan_flag(le, warp_sums, &total);
        if (le && pos < k) {
            o.seq[q * k + pos] = seq_base + (unsigned long long)i;
            o.pay[q * k + pos] = pay_of(i);
            o.eq[q * k + pos] = (dv == T) ? 1 : 0;
        }
        found += total;
    }
}

// ---- collectors: grid = any; CTAs stride over the ambiguous-query list -----------------------------

// rows of a distance matrix (coarse top-w, IVFPQ.java:575-601): candidate i = column i
__global__ void __launch_bounds__(MMIDX_NT) k_tie_collect_rows(const double *__restrict__ D, int ncol, int k,
                                                               const double *__restrict__ res_dist,
                                                               const int32_t *__restrict__ amb_list,
                                                               const int32_t *__restrict__ amb_count, TieLists o) {
    __shared__ int warp_sums[MMIDX_NT / 32];
    const int na = *amb_count;
    for (int a = blockIdx.x; a < na; a += gridDim.x) {
        const int64_t q = amb_list[a];
        const double T = res_dist[q * k + k - 1];
        const double *row = D + q * (int64_t)ncol;
        int found = 0;
        tie_sweep(ncol, 0ull, T, k, found, warp_sums, o, q, [row](int64_t i) { return row[i]; },
                  [](int64_t i) { return (int)i; });
        if (threadIdx.x == 0) o.cnt[q] = min(found, k);
        __syncthreads();
    }
}

struct TieCodeArgs {
    const double *luts;       // IVFPQ: [nq][w][lut_stride]; PQ: [nq][lut_stride]
    int64_t lut_stride;
    const uint8_t *codes;
    const int32_t *iids;      // IVFPQ only
    const int32_t *probes;    // IVFPQ only: [nq][w]
    const int64_t *list_off;  // IVFPQ only
    const int32_t *list_len;  // IVFPQ only
    int64_t n;                // PQ only
    int w, m, ks, code_bytes, k;
};

__device__ __forceinline__ double adc_dist(const double *__restrict__ lut, const uint8_t *__restrict__ cp, int m, int ks) {
    double d0 = 0.0;
    for (int j = 0; j < m; ++j) {
        int code = (ks <= 256) ? (int)cp[j] : (int)reinterpret_cast<const uint16_t *>(cp)[j];
        d0 = __dadd_rn(d0, lut[(int64_t)j * ks + code]);
    }
    return d0;
}

// IVFPQ: offer order = probe rank, then list position (IVFPQ.java:414,429)
__global__ void __launch_bounds__(MMIDX_NT) k_tie_collect_ivfpq(TieCodeArgs a, const double *__restrict__ res_dist,
                                                                const int32_t *__restrict__ amb_list,
                                                                const int32_t *__restrict__ amb_count, TieLists o) {
    __shared__ int warp_sums[MMIDX_NT / 32];
    const int na = *amb_count;
    for (int ai = blockIdx.x; ai < na; ai += gridDim.x) {
        const int64_t q = amb_list[ai];
        const double T = res_dist[q * a.k + a.k - 1];
        int found = 0;
        for (int p = 0; p < a.w && found < a.k; ++p) {
            const int l = a.probes[q * a.w + p];
            const int64_t start = a.list_off[l];
            const double *lut = a.luts + (q * a.w + p) * a.lut_stride;
            const uint8_t *cp = a.codes + start * a.code_bytes;
            const int32_t *li = a.iids + start;
            const int m = a.m, ks = a.ks, cb = a.code_bytes;
            tie_sweep(a.list_len[l], ((unsigned long long)p) << 32, T, a.k, found, warp_sums, o, q,
                      [=](int64_t i) { return adc_dist(lut, cp + i * cb, m, ks); },
                      [li](int64_t i) { return li[i]; });
        }
        if (threadIdx.x == 0) o.cnt[q] = min(found, a.k);
        __syncthreads();
    }
}

// flat PQ: offer order = iid (PQ.java:303)
__global__ void __launch_bounds__(MMIDX_NT) k_tie_collect_pq(TieCodeArgs a, const double *__restrict__ res_dist,
                                                             const int32_t *__restrict__ amb_list,
                                                             const int32_t *__restrict__ amb_count, TieLists o) {
    __shared__ int warp_sums[MMIDX_NT / 32];
    const int na = *amb_count;
    for (int ai = blockIdx.x; ai < na; ai += gridDim.x) {
        const int64_t q = amb_list[ai];
        const double T = res_dist[q * a.k + a.k - 1];
        const double *lut = a.luts + q * a.lut_stride;
        const uint8_t *cp = a.codes;
        const int m = a.m, ks = a.ks, cb = a.code_bytes;
        int found = 0;
        tie_sweep(a.n, 0ull, T, a.k, found, warp_sums, o, q,
                  [=](int64_t i) { return adc_dist(lut, cp + i * cb, m, ks); }, [](int64_t i) { return (int)i; });
        if (threadIdx.x == 0) o.cnt[q] = min(found, a.k);
        __syncthreads();
    }
}

// Linear: offer order = iid (Linear.java:143)
__global__ void __launch_bounds__(MMIDX_NT) k_tie_collect_linear(const double *__restrict__ Q, const double *__restrict__ Xb,
                                                                 int64_t n, int d, int k,
                                                                 const double *__restrict__ res_dist,
                                                                 const int32_t *__restrict__ amb_list,
                                                                 const int32_t *__restrict__ amb_count, TieLists o) {
    __shared__ int warp_sums[MMIDX_NT / 32];
    const int na = *amb_count;
    for (int ai = blockIdx.x; ai < na; ai += gridDim.x) {
        const int64_t q = amb_list[ai];
        const double T = res_dist[q * k + k - 1];
        const double *qv = Q + q * (int64_t)d;
        int found = 0;
        tie_sweep(n, 0ull, T, k, found, warp_sums, o, q,
                  [=](int64_t i) {
                      const double *xp = Xb + (i >> 5) * (int64_t)d * 32 + (i & 31);
                      double acc = 0.0;
                      for (int j = 0; j < d; ++j) acc = sqacc(acc, qv[j], xp[(int64_t)j * 32]);
                      return acc;
                  },
                  [](int64_t i) { return (int)i; });
        if (threadIdx.x == 0) o.cnt[q] = min(found, k);
        __syncthreads();
    }
}

// ---- finish: merge the per-shard first-k lists by seq, keep the latest eq-entries, patch the result ----
// lists are laid out [nparts][nq][k] (an all-gather buffer; nparts == 1 on one GPU), each ascending in seq.
// res_* is the query's sorted top-k (n == k); its entries with dist < T are final, the tail is rewritten.
__global__ void __launch_bounds__(MMIDX_NT) k_tie_finish(int nparts, int64_t nq, int k, const unsigned long long *__restrict__ l_seq,
                                                         const int32_t *__restrict__ l_pay, const int32_t *__restrict__ l_eq,
                                                         const int32_t *__restrict__ l_cnt, const int32_t *__restrict__ amb_list,
                                                         const int32_t *__restrict__ amb_count, int32_t *__restrict__ res_iids,
                                                         double *__restrict__ res_dist, unsigned long long *__restrict__ res_seq,
                                                         PeerSink sink) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long *a_seq = reinterpret_cast<unsigned long long *>(smem_raw);  // [k] first-k le by seq
    int32_t *a_pay = reinterpret_cast<int32_t *>(a_seq + k);                       // [k]
    int32_t *a_eq = a_pay + k;                                                     // [k]
    __shared__ int warp_sums[MMIDX_NT / 32];
    __shared__ int s_nless;
    const int na = *amb_count;
    for (int ai = blockIdx.x; ai < na; ai += gridDim.x) {
        const int64_t q = amb_list[ai];
        const double T = res_dist[q * k + k - 1];
        __syncthreads();
        if (threadIdx.x == 0) s_nless = 0;
        __syncthreads();
        // nless = #result entries with dist < T (sorted ascending -> they are the prefix)
        int loc = 0;
        for (int i = threadIdx.x; i < k; i += MMIDX_NT) loc += (res_dist[q * k + i] < T) ? 1 : 0;
        if (loc) atomicAdd(&s_nless, loc);
        // global rank by seq of every list entry (seqs are unique); rank < k -> a_*[rank]
        for (int e = threadIdx.x; e < nparts * k; e += MMIDX_NT) {
            const int part = e / k, i = e - part * k;
            const int64_t row = (int64_t)part * nq + q;
            if (i >= l_cnt[row]) continue;
            const unsigned long long s = l_seq[row * k + i];
            int rank = i;
            for (int b = 0; b < nparts; ++b) {
                if (b == part) continue;
                const int64_t rb = (int64_t)b * nq + q;
                int lo = 0, hi = l_cnt[rb];
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (l_seq[rb * k + mid] < s)
                        lo = mid + 1;
                    else
                        hi = mid;
                }
                rank += lo;
            }
            if (rank < k) {
                a_seq[rank] = s;
                a_pay[rank] = l_pay[row * k + i];
                a_eq[rank] = l_eq[row * k + i];
            }
        }
        __syncthreads();
        const int nless = s_nless;
        const int keep = k - nless;  // eq-entries that survive
        // neq = number of eq entries among the first k le-entries
        int neq = 0;
        for (int base = 0; base < k; base += MMIDX_NT) {
            int i = base + threadIdx.x;
            int total;
            block_scan_flag(i < k && a_eq[i] != 0, warp_sums, &total);
            neq += total;
        }
        int run = 0;
        for (int base = 0; base < k; base += MMIDX_NT) {
            int i = base + threadIdx.x;
            bool f = i < k && a_eq[i] != 0;
            int total;
            int er = run + block_scan_flag(f, warp_sums, &total);  // rank among eq entries, offer order
            if (f && er >= neq - keep) {
                int dst = nless + (neq - 1 - er);  // later-offered first
                res_iids[q * k + dst] = a_pay[i];
                res_dist[q * k + dst] = T;
                if (res_seq) res_seq[q * k + dst] = a_seq[i];
            }
            run += total;
        }
        __syncthreads();
        if (sink.mode == 2) {  // multi-GPU: the patched row replaces the copy every peer already holds
            int p0, p1;
            const long long row = sink_dest(sink, q, p0, p1);
            PeerSink ps = sink;
            ps.fields &= (SINK_IIDS | SINK_DIST);
            const int32_t *ri = res_iids + q * k;
            const double *rd = res_dist + q * k;
            sink_row(ps, p0, p1, row, k, k, -1.0, [=](int i) { return ri[i]; }, [=](int i) { return rd[i]; },
                     [](int) { return 0ull; });
            __syncthreads();
        }
    }
}

}  // namespace mmidx
