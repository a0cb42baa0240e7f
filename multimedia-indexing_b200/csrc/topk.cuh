// topk.cuh -- CTA-wide bounded top-k collector with the reference's BoundedPriorityQueue semantics.
//
// Replaces com.aliasi.util.BoundedPriorityQueue<Result> (LingPipe 4.0.1) as used at IVFPQ.java:409,445,
// PQ.java:291,318, Linear.java:140,156 and IVFPQ.java:576,590.  Semantics restated in SURVEY.md A.2:
//   * result = the k smallest distances; output order = ascending distance, later-offered first among ties;
//   * only exact binary64 ties AT the k-th boundary depend on offer order.  The collector detects that case
//     (flag "ambiguous") and the caller re-runs the query through the ordered tie pass (tie_resolve.cuh).
//
// Design: candidates that pass a running admission threshold are appended (warp-aggregated shared-memory
// atomic) to an unordered buffer of CAP entries; when the buffer could overflow, an 8x8-bit radix select
// finds the k-th smallest key and compacts in place.  Keys are the raw bits of non-negative doubles, whose
// unsigned order equals the numeric order.  Every candidate carries its offer sequence number.
#pragma once
#include "common.cuh"

#ifndef MMIDX_RANK_SORT
#define MMIDX_RANK_SORT 1  // A/B switch (profiles/run_r2_ab2.sh)
#endif

namespace mmidx {

// histogram increment with warp aggregation: lanes that hit the same bin issue ONE shared-memory atomic
// (the high-order bytes of similar distances all fall into one or two bins, which would serialise 32-way).
// Must be called by all 32 lanes of a converged warp.
__device__ __forceinline__ void hist_add(unsigned int *hist, bool act, unsigned bin) {
    const unsigned key = act ? bin : 0xffffffffu;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (act && (threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&hist[bin], (unsigned)__popc(peers));
}

template <int CAP>
struct TopK {
    static constexpr int ROUND = CAP / 2;     // max pushes between two maybe_compact() calls
    static constexpr int KEEP_MAX = CAP / 2;  // entries kept by a compaction; k <= KEEP_MAX
    static constexpr int PER = CAP / MMIDX_NT;

    double dist[CAP];
    unsigned long long seq[CAP];
    int pay[CAP];
    unsigned int hist[256];
    double thr;       // admission threshold: an upper bound of the final k-th smallest distance
    double tie_drop;  // distance at which entries tied with the threshold were discarded (-1: never)
    int cnt;
    int strict;  // 1: candidates equal to thr are no longer admitted (that tie is already flagged)
    int s_bin, s_krem, s_bincnt, s_newcnt, s_tiecnt;

    __device__ __forceinline__ void init() {
        if (threadIdx.x == 0) {
            thr = __longlong_as_double(0x7ff0000000000000LL);  // +inf
            tie_drop = -1.0;
            cnt = 0;
            strict = 0;
        }
        // other threads must not evaluate maybe_compact()'s predicate on uninitialised shared memory
        __syncthreads();
    }

    // to be called by all 32 lanes of a converged warp
    __device__ __forceinline__ void push(bool pred, double d, unsigned long long s, int p) {
        const unsigned full = 0xffffffffu;
        unsigned mask = __ballot_sync(full, pred);
        if (mask == 0) return;
        int lane = threadIdx.x & 31;
        int leader = __ffs(mask) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&cnt, __popc(mask));
        base = __shfl_sync(full, base, leader);
        if (pred) {
            int slot = base + __popc(mask & ((1u << lane) - 1u));
            dist[slot] = d;
            seq[slot] = s;
            pay[slot] = p;
        }
    }

    // k-th smallest key (1-based k) among the first n entries; all threads; n >= k >= 1
    __device__ void select_kth(int n, int k, unsigned long long &kth, int &need, int &neq) {
        const int tid = threadIdx.x;
        unsigned long long prefix = 0;
        unsigned krem = (unsigned)k;
#pragma unroll 1
        for (int pass = 0; pass < 8; ++pass) {
            const int shift = 56 - 8 * pass;
            hist[tid] = 0;
            __syncthreads();
            for (int i0 = 0; i0 < n; i0 += MMIDX_NT) {
                const int i = i0 + tid;
                bool act = i < n;
                unsigned bin = 0;
                if (act) {
                    unsigned long long key = (unsigned long long)__double_as_longlong(dist[i]);
                    act = (pass == 0 || (key >> (shift + 8)) == prefix);
                    bin = (unsigned)(key >> shift) & 255u;
                }
                hist_add(hist, act, bin);
            }
            __syncthreads();
            if (tid < 32) {
                unsigned loc[8];
                unsigned s = 0;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    loc[b] = hist[tid * 8 + b];
                    s += loc[b];
                }
                unsigned incl = s;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (tid >= o) incl += t;
                }
                unsigned excl = incl - s;
                if (excl < krem && krem <= incl) {
                    unsigned r = krem - excl;
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        if (r != 0u) {
                            if (r <= loc[b]) {
                                s_bin = tid * 8 + b;
                                s_krem = (int)r;
                                s_bincnt = (int)loc[b];
                                r = 0u;
                            } else {
                                r -= loc[b];
                            }
                        }
                    }
                }
            }
            __syncthreads();
            prefix = (prefix << 8) | (unsigned long long)s_bin;
            krem = (unsigned)s_krem;
        }
        kth = prefix;
        need = (int)krem;
        neq = s_bincnt;
        __syncthreads();  // s_* are rewritten by the next call
    }

    // keep every entry below the k-th smallest key plus at most (keep_max - #below) entries equal to it
    __device__ void compact(int k, int keep_max) {
        const int tid = threadIdx.x;
        const int n = cnt;
        unsigned long long kth;
        int need, neq;
        select_kth(n, k, kth, need, neq);
        const int nless = k - need;
        const int tie_room = keep_max - nless;
        double d[PER];
        unsigned long long s[PER];
        int p[PER];
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            int i = tid + e * MMIDX_NT;
            if (i < n) {
                d[e] = dist[i];
                s[e] = seq[i];
                p[e] = pay[i];
            }
        }
        if (tid == 0) {
            s_newcnt = 0;
            s_tiecnt = 0;
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            int i = tid + e * MMIDX_NT;
            if (i < n) {
                unsigned long long key = (unsigned long long)__double_as_longlong(d[e]);
                bool keep = key < kth;
                if (key == kth) keep = atomicAdd(&s_tiecnt, 1) < tie_room;
                if (keep) {
                    int slot = atomicAdd(&s_newcnt, 1);
                    dist[slot] = d[e];
                    seq[slot] = s[e];
                    pay[slot] = p[e];
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            cnt = s_newcnt;
            thr = __longlong_as_double((long long)kth);
            if (neq > tie_room) {
                strict = 1;
                tie_drop = thr;
            } else {
                strict = 0;
            }
        }
        __syncthreads();
    }

    // block-uniform: compacts when fewer than ROUND free slots remain.  Contains the round barrier.
    __device__ __forceinline__ void maybe_compact(int k) {
        if (__syncthreads_or(*(volatile int *)&cnt > CAP - ROUND)) compact(k, KEEP_MAX);
    }

    // Replays the queue's tie rule when the first n entries are a COMPLETE candidate set, i.e. they contain every
    // candidate whose distance is <= T, the k-th smallest (n > k).  BoundedPriorityQueue keeps, among the entries with
    // dist <= T, the first k in offer order (set A); of the entries of A tied at T only the `need` latest-offered
    // survive, and tied entries offered after A were rejected on arrival.  Losers get dist = +inf and
    // drop out in finalize().  flag: scratch for n ints.  All threads; block-uniform.
    __device__ void kill_tie_losers(int n, int k, int *flag) {
        const int tid = threadIdx.x;
        __syncthreads();
        unsigned long long kth;
        int need, neq;
        select_kth(n, k, kth, need, neq);
        if (neq > need) {  // block-uniform
            const double T = __longlong_as_double((long long)kth);
            // rank among the le-entries by offer sequence; A = the first k of them
            for (int i = tid; i < n; i += MMIDX_NT) {
                int f = 0;
                if (dist[i] <= T) {
                    int rank = 0;
                    const unsigned long long si = seq[i];
                    for (int j = 0; j < n; ++j) rank += (dist[j] <= T && seq[j] < si) ? 1 : 0;
                    f = (rank < k) ? ((dist[i] == T) ? 2 : 1) : 3;  // 2: tied entry inside A, 3: offered after A
                }
                flag[i] = f;
            }
            __syncthreads();
            for (int i = tid; i < n; i += MMIDX_NT) {
                const int f = flag[i];
                bool kill = (f == 3 && dist[i] == T);
                if (f == 2) {
                    int later = 0;  // tied entries of A offered after this one
                    const unsigned long long si = seq[i];
                    for (int j = 0; j < n; ++j) later += (flag[j] == 2 && seq[j] > si) ? 1 : 0;
                    kill = later >= need;  // only the `need` latest-offered tied entries stay
                }
                if (kill) dist[i] = __longlong_as_double(0x7ff0000000000000LL);
            }
            __syncthreads();
        }
    }

    // sorts the first n entries (n <= CAP) in BoundedPriorityQueue iteration order: ascending distance, among equal
    // distances the later-offered (larger seq) first.  Entries [n, next power of two) are overwritten with padding.
    __device__ void sort_first(int n) {
        const int tid = threadIdx.x;
#if MMIDX_RANK_SORT
        if (n <= MMIDX_NT) {
            // One element per thread: its position is the number of entries that precede it in the queue's order
            // (distance ascending, then the later offer first; the offer sequence is unique, so positions are too).
            // n broadcast reads per thread and two barriers instead of log^2 n / 2 exchange stages with a barrier each.
            __syncthreads();
            double dv = 0.0;
            unsigned long long sv = 0ull;
            int pv = 0, rank = 0;
            if (tid < n) {
                dv = dist[tid];
                sv = seq[tid];
                pv = pay[tid];
#pragma unroll 4
                for (int i = 0; i < n; ++i) {
                    const double di = dist[i];
                    const unsigned long long si = seq[i];
                    rank += (di < dv || (di == dv && si > sv)) ? 1 : 0;
                }
            }
            __syncthreads();
            if (tid < n) {
                dist[rank] = dv;
                seq[rank] = sv;
                pay[rank] = pv;
            }
            __syncthreads();
            return;
        }
#endif
        int n2 = 1;
        while (n2 < n) n2 <<= 1;
        __syncthreads();
        for (int i = n + tid; i < n2; i += MMIDX_NT) {
            dist[i] = __longlong_as_double(0x7ff0000000000000LL);
            seq[i] = 0ull;
            pay[i] = -1;
        }
        for (int size = 2; size <= n2; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                __syncthreads();
                for (int t = tid; t < (n2 >> 1); t += MMIDX_NT) {
                    int lo = 2 * t - (t & (stride - 1));
                    int hi = lo + stride;
                    bool up = (lo & size) == 0;
                    double dl = dist[lo], dh = dist[hi];
                    unsigned long long sl = seq[lo], sh = seq[hi];
                    bool lo_after_hi = (dl > dh) || (dl == dh && sl < sh);
                    if (lo_after_hi == up) {
                        int pl = pay[lo], ph = pay[hi];
                        dist[lo] = dh;
                        dist[hi] = dl;
                        seq[lo] = sh;
                        seq[hi] = sl;
                        pay[lo] = ph;
                        pay[hi] = pl;
                    }
                }
            }
        }
        __syncthreads();
    }

    // reduce to the final <=k entries, sorted in BoundedPriorityQueue iteration order
    // (ascending distance; among equal distances the later-offered, i.e. larger seq, first).
    // Returns the number of results; *ambiguous is set when an exact tie at the k-th boundary was cut.
    __device__ int finalize(int k, bool *ambiguous) {
        __syncthreads();
        if (cnt > k) compact(k, k);
        const int n = cnt;
        sort_first(n);
        *ambiguous = (n == k) && (tie_drop == dist[n - 1]);
        return n;
    }
};

}  // namespace mmidx
