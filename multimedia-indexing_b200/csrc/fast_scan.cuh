// fast_scan.cuh -- fused IVFADC search kernel: fp32 ADC filter + exact binary64 verification.
//
// Replaces k_lut_build + k_ivfpq_scan for the common geometry (ks <= 256, m in {8,16}) and returns the SAME
// bits: a candidate is only ever accepted on its exact binary64 distance, computed in the reference's
// operation order (computeResidualVector IVFPQ.java:642-648, computeLookupADC :525-538, ADC sum :435-438).
// The fp32 table is used to REJECT candidates that provably cannot enter the queue.
//
// Per probe the reference builds LUT[j][c] = sum_t ((C_l - q)[jS+t] - P[j][c][t])^2 (m*ks*S triples).  Here
//   LUT[j][c] = T1[l][j][c] + T2[q][j][c] + s[q][l][j]
//     T1 = ||C_l,j - P_j,c||^2          per index  (k_build_t1, binary64 then rounded to fp32; nlist*m*ks floats)
//     T2 = 2 q_j . P_j,c                per query  (fp32 FMA chain in the kernel prologue)
//     s  = sum_t q_t (q_t - 2 C_l,t)    per probe  (binary64, m values)
// so a probe costs one 4*m*ks-byte TMA load of T1[l] and two fp32 adds per entry.
//
// Error bound (u = 2^-24; derivation in DESIGN.md 5.1): every fp32 entry differs from the real LUT value by
// at most u*1.002*(3*T1 + (2S+8)*||q_j||*||P_j,c|| + 2|s_j|), and an m-term fp32 sum adds m*u*d32.  With
//   Bq  = 1.02 * u * sum_j max_probes (3*T1max[l][j] + (4S+14)*||q_j||*Pmax[j] + 2|s_j|),   rel = m*u
// the exact distance of a candidate lies in [d32 - rel*d32 - Bq, d32 + rel*d32 + Bq].  Both ends are monotone
// in d32, so the whole selection runs on fp32 keys: if x is the k-th smallest d32 seen so far, U = x*(1+rel)+Bq
// bounds the final k-th exact distance from above and every candidate with d32 > (U+Bq)*(1+2rel) is provably
// outside the result.  The collector keeps everything at or below that slack line (k entries plus the few inside
// the error band); only those survivors (~k per query) are evaluated exactly at the end.  If the band ever holds
// more entries than the collector can keep (massive exact duplicates) the query is handed to the direct kernel.
//
// fp32 range.  The bound above is relative; it holds while no fp32 intermediate overflows and while underflow errors
// stay below the absolute slack FAST_ABS_SLACK that Bq also carries.  Both are guaranteed inside a magnitude window:
// every norm that enters the filter (max ||C_l||, max ||P_j,c||, ||q_j||) is 0 or lies in [FAST_MAG_MIN, FAST_MAG_MAX].
// Then no product of two components exceeds 1e24 (sums of d of them stay far below 3.4e38), and a component that
// underflows fp32 (spacing 1.4e-45) times a factor <= 1e12, over the m(S+3) terms of a distance, errs by < 1e-30.
// An index outside the window never takes the fast path (host check in prepare_fast); a QUERY outside the window is
// flagged by k_fast_prep (bq = NaN) and evaluated by the table-free exact kernel.  Results are the reference's bits either way.
#pragma once
#include "common.cuh"
#include "kernels.cuh"
#include "tie_resolve.cuh"
#include "topk.cuh"

// ---- build-time switches of the fused scan kernel (defaults = the measured best; profiles/run_r2_ab2.sh builds variants) ----
#ifndef MMIDX_SCAN_LB16
#define MMIDX_SCAN_LB16 3  // resident CTAs per SM the m = 16 kernel is compiled for: shared memory allows 3, so 80
                           // registers instead of 64 (config 4 scan 4.29 -> 4.01 ms, profiles/README.md)
#endif
#ifndef MMIDX_SCAN_HOIST16
#define MMIDX_SCAN_HOIST16 1  // m = 16 (80 registers): the list's first code word is requested before the table build (-2 %)
#endif
#ifndef MMIDX_SELECT_PASSES
#define MMIDX_SELECT_PASSES 3  // radix passes of the collector's threshold selection (4: the exact k-th key)
#endif
#ifndef MMIDX_SCAN_STAGE
#define MMIDX_SCAN_STAGE 1  // A/B switch (profiles/run_r2_ab2.sh): probe descriptors staged in shared memory
#endif

namespace mmidx {

// ---- index-time tables ------------------------------------------------------------------------------
// T1[l][j*ks + c] and t1max[l][j] (must be zero-filled before launch).  grid (nlist, m), thread <-> c.
__global__ void __launch_bounds__(MMIDX_NT) k_build_t1(const double *__restrict__ C, const double *__restrict__ P,
                                                       const int32_t *__restrict__ perm, int d, int m, int ks, int S,
                                                       float *__restrict__ T1, float *__restrict__ t1max) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *cv = reinterpret_cast<double *>(smem_raw);  // [S] sub-vector of the (permuted) coarse centroid
    const int l = blockIdx.x, j = blockIdx.y;
    for (int t = threadIdx.x; t < S; t += MMIDX_NT) {
        int src = j * S + t;
        if (perm) src = perm[src];
        cv[t] = C[(int64_t)l * d + src];
    }
    __syncthreads();
    float mx = 0.f;
    for (int c = threadIdx.x; c < ks; c += MMIDX_NT) {
        const double *pc = P + ((int64_t)j * ks + c) * S;
        double acc = 0.0;
        for (int t = 0; t < S; ++t) acc = sqacc(acc, cv[t], pc[t]);
        float f = __double2float_rn(acc);
        T1[(int64_t)l * m * ks + (int64_t)j * ks + c] = f;
        mx = fmaxf(mx, f);
    }
    // non-negative floats order like their bit patterns
    atomicMax(reinterpret_cast<int *>(&t1max[(int64_t)l * m + j]), __float_as_int(mx));
}

// P32t[j][t][c] = fp32(P[j][c][t]) (thread <-> c reads coalesce) and pmax[j] >= max_c ||P_j,c||. grid m.
__global__ void __launch_bounds__(MMIDX_NT) k_build_p32t(const double *__restrict__ P, int m, int ks, int S,
                                                         float *__restrict__ P32t, float *__restrict__ pmax) {
    const int j = blockIdx.x;
    float mx = 0.f;
    for (int c = threadIdx.x; c < ks; c += MMIDX_NT) {
        const double *pc = P + ((int64_t)j * ks + c) * S;
        double n2 = 0.0;
        for (int t = 0; t < S; ++t) {
            P32t[((int64_t)j * S + t) * ks + c] = __double2float_rn(pc[t]);
            n2 += pc[t] * pc[t];
        }
        mx = fmaxf(mx, __double2float_ru(sqrt(n2) * (1.0 + 1e-12)));
    }
    atomicMax(reinterpret_cast<int *>(&pmax[j]), __float_as_int(mx));
}

// ---- bank-conflict-aware order inside an inverted list ------------------------------------------------
// The scan's cost is shared-memory wavefronts: one warp-wide lookup into sub-table j takes as many cycles as
// the most loaded bank has DISTINCT addresses (bank = code % 32 for fp32 entries).  Order inside a list is free
// (the queue's offer order is carried separately as `orank`), so each list is re-ordered greedily such that the 32
// codes one warp instruction touches spread over the banks: pick by pick, the entry that raises the fewest per-
// sub-quantizer maxima (then the least sum_j (2*load_j[bank]+1)) over sub-quantizers whose address is not yet in the
// group.  The pool is the whole list up to RCH entries (longer lists: chunks of RCH).  Measured on the bench index
// (scratch/sim_reorder.c): 2.77 wavefronts per lookup in insertion order, 1.71 with a 512-entry pool and the plain
// sum cost, 1.11 with this pool and cost.
// M == 8: a thread scans two adjacent codes per 128-bit load, so a warp instruction covers the even (then the
// odd) positions of a 64-entry block; M == 16: 32 consecutive positions.  grid nlist; src[] is list-relative.
constexpr int RCH = 4096;

template <int M>
__global__ void __launch_bounds__(MMIDX_NT) k_reorder_lists(const uint8_t *__restrict__ codes, const int64_t *__restrict__ list_off,
                                                            const int32_t *__restrict__ list_len, int32_t *__restrict__ src) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint8_t(*cs)[M] = reinterpret_cast<uint8_t(*)[M]>(smem_raw);  // [RCH][M]
    __shared__ int load[M][32];
    __shared__ int mx[M];
    __shared__ unsigned seen[M][8];
    __shared__ unsigned red[MMIDX_NT / 32];
    __shared__ unsigned s_win;
    constexpr int BLK = (M == 8) ? 64 : 32;
    constexpr int PER = RCH / MMIDX_NT;
    constexpr int MAXPEN = 1000;  // raising a sub-quantizer's maximum outweighs any sum of load terms
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = blockIdx.x;
    const int64_t start = list_off[l];
    const int len = list_len[l];
    for (int cb = 0; cb < len; cb += RCH) {
        const int n = min(RCH, len - cb);
        __syncthreads();
        for (int e = tid; e < n * M; e += MMIDX_NT) (&cs[0][0])[e] = codes[(start + cb) * M + e];
        unsigned alive = 0u;
#pragma unroll
        for (int r = 0; r < PER; ++r)
            if (tid + r * MMIDX_NT < n) alive |= 1u << r;
        for (int bb = 0; bb < n; bb += BLK) {
            const int rblk = min(BLK, n - bb);
            for (int h = 0; h < BLK / 32; ++h) {
                const int gsize = (M == 8) ? ((h == 0) ? (rblk + 1) / 2 : rblk / 2) : rblk;
                __syncthreads();
                for (int e = tid; e < M * 32; e += MMIDX_NT) (&load[0][0])[e] = 0;
                for (int e = tid; e < M * 8; e += MMIDX_NT) (&seen[0][0])[e] = 0u;
                if (tid < M) mx[tid] = 0;
                for (int t = 0; t < gsize; ++t) {
                    __syncthreads();
                    unsigned best = 0xffffffffu;
#pragma unroll
                    for (int r = 0; r < PER; ++r) {
                        if ((alive >> r) & 1u) {
                            const int idx = tid + r * MMIDX_NT;
                            unsigned cw[M / 4];
                            if (M == 8) {
                                const uint2 v = *reinterpret_cast<const uint2 *>(&cs[idx][0]);
                                cw[0] = v.x;
                                cw[1] = v.y;
                            } else {
                                const uint4 v = *reinterpret_cast<const uint4 *>(&cs[idx][0]);
                                cw[0] = v.x;
                                cw[1] = v.y;
                                cw[2 % (M / 4)] = v.z;
                                cw[3 % (M / 4)] = v.w;
                            }
                            int cost = 0;
#pragma unroll
                            for (int j = 0; j < M; ++j) {
                                const unsigned c = (cw[j >> 2] >> (8 * (j & 3))) & 255u;
                                const bool dup = (seen[j][c >> 5] >> (c & 31u)) & 1u;
                                const int ld = load[j][c & 31u];
                                cost += dup ? 0 : ((ld + 1 > mx[j]) ? MAXPEN : 0) + 2 * ld + 1;
                            }
                            best = min(best, ((unsigned)cost << 12) | (unsigned)idx);
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
                    if (lane == 0) red[warp] = best;
                    __syncthreads();
                    if (tid == 0) {
                        unsigned b = red[0];
                        for (int wv = 1; wv < MMIDX_NT / 32; ++wv) b = min(b, red[wv]);
                        s_win = b;
                        const int idx = (int)(b & 0xfffu);
                        const int pos = (M == 8) ? (bb + 2 * t + h) : (bb + t);
                        src[start + cb + pos] = cb + idx;
                        for (int j = 0; j < M; ++j) {
                            const unsigned c = cs[idx][j];
                            if (!((seen[j][c >> 5] >> (c & 31u)) & 1u)) {
                                seen[j][c >> 5] |= 1u << (c & 31u);
                                const int nl = ++load[j][c & 31u];
                                if (nl > mx[j]) mx[j] = nl;
                            }
                        }
                    }
                    __syncthreads();
                    const int widx = (int)(s_win & 0xfffu);
                    if ((widx & (MMIDX_NT - 1)) == tid) alive &= ~(1u << (widx / MMIDX_NT));
                }
            }
        }
    }
}

// apply the order: ocodes/oiids/orank[start + i] = codes/iids/rank of insertion entry src[start + i]
__global__ void k_apply_order(const uint8_t *__restrict__ codes, const int32_t *__restrict__ iids,
                              const int64_t *__restrict__ list_off, const int32_t *__restrict__ list_len,
                              const int32_t *__restrict__ src, int m, uint8_t *__restrict__ ocodes,
                              int32_t *__restrict__ oiids, int32_t *__restrict__ orank) {
    const int l = blockIdx.x;
    const int64_t start = list_off[l];
    const int len = list_len[l];
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
        const int s = src[start + i];
        for (int b = 0; b < m; ++b) ocodes[(start + i) * m + b] = codes[(start + s) * m + b];
        oiids[start + i] = iids[start + s];
        orank[start + i] = s;
    }
}

// One descriptor per (query, owned probe slot): everything the scan kernel needs about a probe in one place, so a
// probe costs one independent 16-byte load instead of a chain of four dependent ones (oprobes -> probes -> list_off /
// list_len -> sall).  Followed in memory by the m floats s[q][probe][j]; stride = fast_desc_stride(m) bytes.
struct __align__(16) ProbeHdr {
    long long start;  // first position of the list in the CSR arrays
    int len;          // > 0
    int p;            // probe rank (offer order of the queue, IVFPQ.java:414)
    int l;            // coarse centroid / list id
    int pad[3];
};
static_assert(sizeof(ProbeHdr) == 32, "ProbeHdr layout");
__host__ __device__ constexpr int fast_desc_stride(int m) { return (int)sizeof(ProbeHdr) + 4 * m; }

// flat PQ index as pseudo lists: probes[q][p] = p for every query; iids[i] = i
__global__ void k_fill_flat_probes(int64_t nq, int w, int32_t *__restrict__ probes) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq * w) probes[i] = (int32_t)(i % w);
}
__global__ void k_iota_negate(int64_t n_iota, int32_t *__restrict__ iota, int64_t n_neg, const double *__restrict__ src,
                              double *__restrict__ neg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_iota) iota[i] = (int32_t)i;
    if (i < n_neg) neg[i] = -src[i];
}

// Launch order of a batch that fills only a few waves of CTAs: heaviest queries first (longest-processing-time rule),
// so the last wave is made of short CTAs.  One CTA, nq <= ORDER_MAX; bitonic sort of (work desc, query asc).
constexpr int ORDER_MAX = 4096;
constexpr int ORDER_NT = 1024;
__global__ void __launch_bounds__(ORDER_NT) k_order_queries(int nq, const int32_t *__restrict__ work, int32_t *__restrict__ qorder) {
    __shared__ unsigned long long key[ORDER_MAX];
    int n2 = 1;
    while (n2 < nq) n2 <<= 1;
    for (int i = threadIdx.x; i < n2; i += ORDER_NT)
        key[i] = i < nq ? ((unsigned long long)(0x7fffffffu - (unsigned)work[i]) << 32) | (unsigned)i : ~0ull;
    for (int size = 2; size <= n2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (n2 >> 1); t += ORDER_NT) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = key[lo], b = key[hi];
                if ((a > b) == up) {
                    key[lo] = b;
                    key[hi] = a;
                }
            }
        }
    __syncthreads();
    for (int i = threadIdx.x; i < nq; i += ORDER_NT) qorder[i] = (int32_t)(key[i] & 0xffffffffu);
}

struct FastArgs {
    const double *Q;          // [nq][d]
    const double *C;          // [nlist][d]
    const double *P;          // [m][ks][S]
    const float *T1;          // [nlist][m*ks]
    const float *P32t;        // [m][S][ks]
    const float *t1max;       // [nlist][m]
    const float *pmax;        // [m]
    const int32_t *perm;      // [d] or NULL
    const int32_t *probes;    // [nq][w]
    const uint8_t *codes;     // CSR, insertion order (direct kernel, tie pass)
    const int32_t *iids;
    const uint8_t *ocodes;    // CSR, bank-conflict-aware order inside each list (k_reorder_lists); fast kernel only
    const int32_t *oiids;
    const int32_t *orank;     // insertion rank (= offer order inside the list) of every reordered entry
    const int64_t *list_off;
    const int32_t *list_len;
    const double *bq;         // [nq]        error radius of the fp32 distances of query q (k_fast_prep)
    const float *T2;          // [nq][m*256] per-query term of the table decomposition (k_fast_t2)
    const int32_t *oprobes;   // [nq][w]     ranks of the probes with a non-empty list on this shard, ascending
    const int32_t *ocnt;      // [nq]
    const int32_t *qorder;    // [nq] query handled by CTA row y (k_order_queries), or NULL: y itself
    const unsigned char *desc;  // [nq][w] probe descriptors in oprobes order (ProbeHdr + s[m]), k_fast_prep
    int d, m, ks, S, w, k, nsplit;
    int flat;                   // flat PQ index: every (pseudo) list shares coarse row 0 (a zero centroid)
    int resolve_ties;           // 1: replay the queue's tie rule inside the kernel (unsharded, nsplit == 1)
    int desc_stage;             // 1: the launch carries w * fast_desc_stride(m) extra bytes of shared memory for the descriptors
    int32_t *fb_list;           // (q*nsplit + s) items whose error band overflowed the collector
    int32_t *fb_count;
    unsigned long long *stats;  // optional [4]: candidates, re-scanned lists, exact evaluations, direct fallbacks
};

// order-preserving map of fp32 bit patterns to unsigned keys (d32 may be slightly negative)
__device__ __forceinline__ unsigned f32_key(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_unkey(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// shared-memory gather with a compile-time offset: address = base + 4*byte, one LEA + one LDS per lookup
template <int OFF>
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.volatile.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
    return v;
}
// ADC table entry of sub-quantizer J for byte I of a packed code word (ks == 256: 1 KiB per sub-table)
// address = lut + 4 * byte_I(word) in ONE instruction: a 4-way byte dot product with the constant 4 << 8I (SASS IDP.4A;
// measured 3.4 % faster than PRMT + IMAD, profiles/README.md)
#define MMIDX_LK(J, word, I) lds_f32<(J) * 1024>(__dp4a((unsigned)(word), 4u << (8 * (I)), (unsigned)lut))

constexpr int FAST_POS_BITS = 22;  // lists longer than 2^22 entries disable the fast path (host check); w <= 1024

// CTA-wide collector on fp32 keys with an error-band slack (see the header comment).
template <int CAP>
struct TopK32 {
    static constexpr int ROUND = CAP / 2;
    static constexpr int KEEP_MAX = CAP / 2;
    static constexpr int PER = CAP / MMIDX_NT;
    float key[CAP];
    unsigned int pk[CAP];  // (probe rank << 22) | position inside the (re-ordered) list
    unsigned int hist[2][256];
    float thr32;  // admission threshold: candidates with d32 > thr32 are provably outside the result
    int cnt, overflow;
    int ovf;  // a barrier-free list scan ran out of slots: its entries are dropped and the list is scanned again
    int s_bin, s_krem, s_newcnt;

    __device__ __forceinline__ void init() {
        if (threadIdx.x == 0) {
            thr32 = __int_as_float(0x7f800000);
            cnt = 0;
            overflow = 0;
            ovf = 0;
        }
        __syncthreads();
    }

    // Barrier-free scan: both candidates of one 128-bit load (p1 == false for a single candidate); all 32 lanes of a
    // converged warp.  There is no capacity guarantee here: entries past CAP are not stored and `ovf` is raised; every
    // slot below min(cnt, CAP) still holds a valid entry, so drop_tag() can undo the list.
    __device__ __forceinline__ void push2(bool p0, float d0, bool p1, float d1, unsigned int packed0) {
        const unsigned m0 = __ballot_sync(0xffffffffu, p0), m1 = __ballot_sync(0xffffffffu, p1);
        const int lane = threadIdx.x & 31;
        const int n0 = __popc(m0), tot = n0 + __popc(m1);
        int base = 0;
        if (lane == 0) {
            base = atomicAdd(&cnt, tot);
            if (base + tot > CAP) ovf = 1;
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        const unsigned lt = (1u << lane) - 1u;
        if (p0) {
            const int slot = base + __popc(m0 & lt);
            if (slot < CAP) {
                key[slot] = d0;
                pk[slot] = packed0;
            }
        }
        if (p1) {
            const int slot = base + n0 + __popc(m1 & lt);
            if (slot < CAP) {
                key[slot] = d1;
                pk[slot] = packed0 + 1u;
            }
        }
    }

    // remove every entry of probe slot `tag` at list position >= pos0 (after an overflowed barrier-free scan of the list
    // segment that starts there); all threads, block-uniform
    __device__ __noinline__ void drop_tag(unsigned int tag, unsigned int pos0) {
        const int tid = threadIdx.x;
        const int n = min(cnt, CAP);
        float d[PER];
        unsigned int pp[PER];
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            const int i = tid + e * MMIDX_NT;
            if (i < n) {
                d[e] = key[i];
                pp[e] = pk[i];
            }
        }
        if (tid == 0) s_newcnt = 0;
        __syncthreads();
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            const int i = tid + e * MMIDX_NT;
            if (i < n && ((pp[e] >> FAST_POS_BITS) != tag || (pp[e] & ((1u << FAST_POS_BITS) - 1u)) < pos0)) {
                const int slot = atomicAdd(&s_newcnt, 1);
                key[slot] = d[e];
                pk[slot] = pp[e];
            }
        }
        __syncthreads();
        if (tid == 0) {
            cnt = s_newcnt;
            ovf = 0;
        }
        __syncthreads();
    }

    // all 32 lanes of a converged warp
    __device__ __forceinline__ void push(bool pred, float d, unsigned int packed) {
        const unsigned mask = __ballot_sync(0xffffffffu, pred);
        if (mask == 0) return;
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(mask) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&cnt, __popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (pred) {
            const int slot = base + __popc(mask & ((1u << lane) - 1u));
            key[slot] = d;
            pk[slot] = packed;
        }
    }

    // An upper bound of the k-th smallest key (1-based) among the first n entries, n >= k >= 1, that is at most 2^-15
    // (relative) above it: MMIDX_SELECT_PASSES = 3 radix passes of 8 bits locate the key's 24-bit prefix and the largest value
    // with that prefix is returned (any bound >= the k-th key keeps the survivor set complete; a looser one only keeps a
    // few more).  A non-finite bound (prefix of inf / nan patterns) takes the fourth pass: the exact key.
    // Two histograms alternate, so that zeroing the next one needs no barrier of its own.
    __device__ float select_kth(int n, int k) {
        const int tid = threadIdx.x;
        unsigned prefix = 0, krem = (unsigned)k;
        hist[0][tid] = 0;
        __syncthreads();
        int pass = 0;
#pragma unroll 1
        for (; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            unsigned int *h = hist[pass & 1];
            hist[(pass + 1) & 1][tid] = 0;  // last read two barriers ago
            for (int i0 = 0; i0 < n; i0 += MMIDX_NT) {
                const int i = i0 + tid;
                bool act = i < n;
                unsigned bin = 0;
                if (act) {
                    const unsigned kk = f32_key(key[i]);
                    act = (pass == 0 || (kk >> (shift + 8)) == prefix);
                    bin = (kk >> shift) & 255u;
                }
                hist_add(h, act, bin);
            }
            __syncthreads();
            if (tid < 32) {
                unsigned loc[8], sum = 0;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    loc[b] = h[tid * 8 + b];
                    sum += loc[b];
                }
                unsigned incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (tid >= o) incl += t;
                }
                const unsigned excl = incl - sum;
                if (excl < krem && krem <= incl) {
                    unsigned r = krem - excl;
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        if (r != 0u) {
                            if (r <= loc[b]) {
                                s_bin = tid * 8 + b;
                                s_krem = (int)r;
                                r = 0u;
                            } else {
                                r -= loc[b];
                            }
                        }
                    }
                }
            }
            __syncthreads();
            prefix = (prefix << 8) | (unsigned)s_bin;
            krem = (unsigned)s_krem;
            if (pass == MMIDX_SELECT_PASSES - 1 && pass < 3) {
                const float edge = f32_unkey((prefix << (shift)) | ((1u << shift) - 1u));
                if (edge <= 3.4028234663852886e38f) return edge;  // block-uniform
            }
        }
        return f32_unkey(prefix);
    }

    // keep every entry at or below the slack line of the k-th smallest key
    __device__ __noinline__ void compact(int k, double bq, double rel) {
        const int tid = threadIdx.x;
        const int n = cnt;
        const float kth = select_kth(n, k);
        const double U = (double)kth + rel * fabs((double)kth) + bq;
        const float keep32 = __double2float_ru((U + bq) * (1.0 + 2.0 * rel));
        float d[PER];
        unsigned int pp[PER];
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            const int i = tid + e * MMIDX_NT;
            if (i < n) {
                d[e] = key[i];
                pp[e] = pk[i];
            }
        }
        if (tid == 0) s_newcnt = 0;
        __syncthreads();
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            const int i = tid + e * MMIDX_NT;
            if (i < n && d[e] <= keep32) {
                const int slot = atomicAdd(&s_newcnt, 1);
                if (slot < KEEP_MAX) {
                    key[slot] = d[e];
                    pk[slot] = pp[e];
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            if (s_newcnt > KEEP_MAX) overflow = 1;  // the error band holds more than we can keep: direct kernel
            cnt = min(s_newcnt, KEEP_MAX);
            thr32 = keep32;
        }
        __syncthreads();
    }

    __device__ __forceinline__ void maybe_compact(int k, double bq, double rel) {
        if (__syncthreads_or(*(volatile int *)&cnt > CAP - ROUND)) compact(k, bq, rel);
    }
};

// exact binary64 ADC distance of one stored code, straight from the quantizers (reference operation order)
__device__ __forceinline__ double exact_adc(const double *__restrict__ Cl, const double *__restrict__ qv,
                                            const int32_t *__restrict__ perm, const double *__restrict__ P,
                                            const uint8_t *__restrict__ code, int m, int ks, int S) {
    double dist = 0.0;
    for (int j = 0; j < m; ++j) {
        const int c = (ks <= 256) ? (int)code[j] : (int)reinterpret_cast<const uint16_t *>(code)[j];
        const double *pc = P + ((int64_t)j * ks + c) * S;
        double acc = 0.0;
        for (int t = 0; t < S; ++t) {
            int src = j * S + t;
            if (perm) src = perm[src];
            // residual = centroid - query (one rounding), then (r - p)^2 accumulated in order
            double r = __dsub_rn(Cl[src], qv[src]);
            acc = sqacc(acc, r, pc[t]);
        }
        dist = __dadd_rn(dist, acc);
    }
    return dist;
}

// exact collector for the survivors (more survivors than this -> direct kernel); their codes are staged in the dead
// fp32 key array of the first collector, which bounds it at 4 * CAP32 / M
template <int CAP32, int M>
struct FastExactCap {
    static constexpr int value = (4 * CAP32 / M) < 512 ? (4 * CAP32 / M) : 512;
};


// T2[q][j*256 + c] = 2 * sum_t q32[perm(jS+t)] * P32t[j][t][c]: the per-query term of the table decomposition, for
// a whole batch.  grid (ceil(nq / T2_QB), m), thread <-> c; the thread keeps its S codebook values in registers and
// the CTA's query sub-vectors sit in shared memory (read as warp-wide broadcasts).  fp32 FMA chain, t ascending
// (the error bound in the header assumes exactly this chain).
constexpr int T2_QB = 32;

template <int ST>
__global__ void __launch_bounds__(MMIDX_NT) k_fast_t2(const double *__restrict__ Q, const int32_t *__restrict__ perm,
                                                      const float *__restrict__ P32t, int64_t nq, int d, int m, int S_rt,
                                                      float *__restrict__ T2) {
    const int S = ST > 0 ? ST : S_rt;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *qs = reinterpret_cast<float *>(smem_raw);  // [T2_QB][S]
    const int j = blockIdx.y, c = threadIdx.x;        // ks == 256 == blockDim.x
    const int64_t q0 = (int64_t)blockIdx.x * T2_QB;
    const int nb = (int)min((int64_t)T2_QB, nq - q0);
    for (int e = threadIdx.x; e < nb * S; e += MMIDX_NT) {
        const int qi = e / S, t = e - qi * S;
        int src = j * S + t;
        if (perm) src = perm[src];
        qs[e] = __double2float_rn(Q[(q0 + qi) * (int64_t)d + src]);
    }
    __syncthreads();
    const float *pp = P32t + (int64_t)j * S * 256 + c;
    if (ST > 0) {
        float pr[ST > 0 ? ST : 1];
#pragma unroll
        for (int t = 0; t < ST; ++t) pr[t] = pp[t * 256];
        for (int qi = 0; qi < nb; ++qi) {
            float acc = 0.f;
#pragma unroll
            for (int t = 0; t < ST; ++t) acc = fmaf(qs[qi * ST + t], pr[t], acc);
            T2[(q0 + qi) * (int64_t)m * 256 + j * 256 + c] = 2.f * acc;
        }
    } else {
        for (int qi = 0; qi < nb; ++qi) {
            float acc = 0.f;
            for (int t = 0; t < S; ++t) acc = fmaf(qs[qi * S + t], pp[(int64_t)t * 256], acc);
            T2[(q0 + qi) * (int64_t)m * 256 + j * 256 + c] = 2.f * acc;
        }
    }
}

// Per-query pre-pass.  For every probe whose list is non-empty on this shard, in probe-rank order: a ProbeHdr and
// s[j] = sum_t q_t (q_t - 2 C_l,t) over the (permuted) sub-vector j, in binary64 then rounded to fp32; and the
// error radius Bq of the query (header comment).  grid nq.
// MT/ST > 0: compile-time m and S (S a power of two <= 32: S consecutive lanes share one (probe, j) pair and read
// the centroid sub-vector as one contiguous segment); 0: generic.
template <int MT, int ST>
__global__ void __launch_bounds__(MMIDX_NT) k_fast_prep(const double *__restrict__ Q, const double *__restrict__ C,
                                                        const int32_t *__restrict__ perm, const int32_t *__restrict__ probes,
                                                        const float *__restrict__ t1max, const float *__restrict__ pmax,
                                                        const int64_t *__restrict__ list_off,
                                                        const int32_t *__restrict__ list_len, int d, int m_rt, int S_rt,
                                                        int w, int flat, unsigned char *__restrict__ desc, double *__restrict__ bq,
                                                        int32_t *__restrict__ oprobes, int32_t *__restrict__ ocnt,
                                                        int32_t *__restrict__ work) {
    const int m = MT > 0 ? MT : m_rt;
    const int S = ST > 0 ? ST : S_rt;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // the (permuted) query as m sub-vectors of stride S + 1: threads with different j then read different banks
    double *qp = reinterpret_cast<double *>(smem_raw);       // [m][S + 1]
    float *qn = reinterpret_cast<float *>(qp + m * (S + 1));  // [m]  upper bound of ||q_j||
    float *bterm = qn + m;                              // [m]
    int *slot_of = reinterpret_cast<int *>(bterm + m);  // [w]  descriptor slot of probe rank p, -1: not on this shard
    int *lst = slot_of + w;                             // [w]  list id of probe rank p
    const int tid = threadIdx.x;
    const int64_t q = blockIdx.x;
    const int dstride = fast_desc_stride(m);
    unsigned char *dq = desc + q * (int64_t)w * dstride;
    for (int i = tid; i < d; i += MMIDX_NT) {
        const int j = i / S, t = i - j * S;
        qp[j * (S + 1) + t] = Q[q * (int64_t)d + (perm ? perm[i] : i)];
    }
    if (tid < m) bterm[tid] = 0.f;
    __syncthreads();
    if (tid < m) {
        double n2 = 0.0;
        for (int t = 0; t < S; ++t) n2 += qp[tid * (S + 1) + t] * qp[tid * (S + 1) + t];
        qn[tid] = __fsqrt_ru(__double2float_ru(n2 * (1.0 + 1e-12)));
    }
    const int32_t *pr = probes + q * w;
    const double cS = 4.0 * S + 14.0;
    // probe ranks with a non-empty list on this shard, in rank order (lists other shards own have length 0 here)
    if (tid < 32) {
        int base = 0;
        long long cand = 0;  // candidates this query will scan on this shard
        for (int p0 = 0; p0 < w; p0 += 32) {
            const int p = p0 + tid;
            int l = 0, len = 0;
            if (p < w) {
                l = pr[p];
                len = list_len[l];
                lst[p] = flat ? 0 : l;
            }
            cand += len;
            const bool own = len > 0;
            const unsigned mk = __ballot_sync(0xffffffffu, own);
            if (p < w) slot_of[p] = -1;
            if (own) {
                const int slot = base + __popc(mk & ((1u << tid) - 1u));
                slot_of[p] = slot;
                oprobes[q * w + slot] = p;
                ProbeHdr h;
                h.start = list_off[l];
                h.len = len;
                h.p = p;
                h.l = flat ? 0 : l;  // row of C / T1 / t1max
                h.pad[0] = h.pad[1] = h.pad[2] = 0;
                *reinterpret_cast<ProbeHdr *>(dq + (int64_t)slot * dstride) = h;
            }
            base += __popc(mk);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cand += __shfl_xor_sync(0xffffffffu, cand, o);
        if (tid == 0) {
            ocnt[q] = base;
            if (work) work[q] = (int32_t)min(cand, (long long)0x7fffffff);
        }
    }
    __syncthreads();
    // One thread per (probe, sub-quantizer) pair: its S terms are added serially from 128-bit loads of the centroid's
    // sub-vector (a 16-lane shuffle tree per pair was 10x the instructions: 142 -> 45 us per 10 000 queries).
    for (int e = tid; e < w * m; e += MMIDX_NT) {
        const int p = e / m, j = e - p * m;
        const int slot = slot_of[p];
        if (slot < 0) continue;
        const int l = lst[p];
        const double *Cl = C + (int64_t)l * d;
        const double *qj = qp + j * (S + 1);
        double acc = 0.0;
        if (ST > 0 && (ST & 1) == 0 && perm == nullptr) {
            constexpr int SS = ST > 0 ? ST : 2;
            const double2 *c2 = reinterpret_cast<const double2 *>(Cl + j * SS);  // 16-byte aligned: d and S are even
#pragma unroll
            for (int t2 = 0; t2 < SS / 2; ++t2) {
                const double2 c = c2[t2];
                const double q0 = qj[2 * t2], q1 = qj[2 * t2 + 1];
                acc += q0 * (q0 - 2.0 * c.x);
                acc += q1 * (q1 - 2.0 * c.y);
            }
        } else {
            for (int t = 0; t < S; ++t) {
                int src = j * S + t;
                if (perm) src = perm[src];
                const double qt = qj[t];
                acc += qt * (qt - 2.0 * Cl[src]);
            }
        }
        reinterpret_cast<float *>(dq + (int64_t)slot * dstride + sizeof(ProbeHdr))[j] = __double2float_rn(acc);
        const double term = 3.0 * (double)t1max[(int64_t)l * m + j] + cS * (double)qn[j] * (double)pmax[j] + 2.0 * fabs(acc);
        atomicMax(reinterpret_cast<int *>(&bterm[j]), __float_as_int(__double2float_ru(term)));
    }
    __syncthreads();
    if (tid == 0) {
        double b = 0.0;
        bool ok = true;
        for (int j = 0; j < m; ++j) {
            b += (double)bterm[j];
            ok = ok && fast_mag_ok((double)qn[j]);  // NaN / inf / outside the fp32-safe window -> false
        }
        ok = ok && b < 1e37;  // every term finite
        // NaN marks a query the fp32 filter must not touch: the scan kernel hands it to the exact table-free kernel
        bq[q] = ok ? 1.02 * 5.9604644775390625e-08 * b + FAST_ABS_SLACK : __longlong_as_double(0x7ff8000000000000LL);
    }
}

// ---- packed fp32 pairs (Blackwell add.f32x2, SASS FADD2): one issue slot adds both candidates of a 128-bit load ----
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long x, unsigned long long y) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y));
    return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

// fp32 ADC sums of the code word(s) of one 128-bit load; `lut` is the 32-bit shared-window address of the table.  M == 8: two candidates (c.xy, c.zw), d0/d1 are their sums,
// terms added for j ascending.  M == 16: one candidate, d0 = the sum (even-j and odd-j chains added last), d1 unused.
template <int M>
__device__ __forceinline__ void adc_pair(const uint32_t lut, const uint4 c, float &d0, float &d1) {
    unsigned long long acc;
    if (M == 8) {
        acc = f2_pack(MMIDX_LK(0, c.x, 0), MMIDX_LK(0, c.z, 0));
        acc = f2_add(acc, f2_pack(MMIDX_LK(1, c.x, 1), MMIDX_LK(1, c.z, 1)));
        acc = f2_add(acc, f2_pack(MMIDX_LK(2, c.x, 2), MMIDX_LK(2, c.z, 2)));
        acc = f2_add(acc, f2_pack(MMIDX_LK(3, c.x, 3), MMIDX_LK(3, c.z, 3)));
        acc = f2_add(acc, f2_pack(MMIDX_LK(4, c.y, 0), MMIDX_LK(4, c.w, 0)));
        acc = f2_add(acc, f2_pack(MMIDX_LK(5, c.y, 1), MMIDX_LK(5, c.w, 1)));
        acc = f2_add(acc, f2_pack(MMIDX_LK(6, c.y, 2), MMIDX_LK(6, c.w, 2)));
        acc = f2_add(acc, f2_pack(MMIDX_LK(7, c.y, 3), MMIDX_LK(7, c.w, 3)));
        f2_unpack(acc, d0, d1);
    } else {
        acc = f2_pack(MMIDX_LK(0, c.x, 0), MMIDX_LK(1, c.x, 1));
        acc = f2_add(acc, f2_pack(MMIDX_LK(2, c.x, 2), MMIDX_LK(3, c.x, 3)));
        acc = f2_add(acc, f2_pack(MMIDX_LK(4, c.y, 0), MMIDX_LK(5, c.y, 1)));
        acc = f2_add(acc, f2_pack(MMIDX_LK(6, c.y, 2), MMIDX_LK(7, c.y, 3)));
        acc = f2_add(acc, f2_pack(MMIDX_LK(8, c.z, 0), MMIDX_LK(9, c.z, 1)));
        acc = f2_add(acc, f2_pack(MMIDX_LK(10, c.z, 2), MMIDX_LK(11, c.z, 3)));
        acc = f2_add(acc, f2_pack(MMIDX_LK(12, c.w, 0), MMIDX_LK(13, c.w, 1)));
        acc = f2_add(acc, f2_pack(MMIDX_LK(14, c.w, 2), MMIDX_LK(15, c.w, 3)));
        float lo, hi;
        f2_unpack(acc, lo, hi);
        d0 = lo + hi;
        d1 = 0.f;
    }
}

// One list with a block barrier per ROUND candidates: the collector can never run out of slots.  Used while the
// admission threshold is still +inf (everything is pushed) and to redo a list whose barrier-free scan overflowed.
// lc / len: the list segment to scan; ptag already carries the segment's first list position (tag | pos0).
template <int CAP32, int M>
__device__ __noinline__ void scan_list_rounds(TopK32<CAP32> &c32, const uint32_t lut, const uint8_t *__restrict__ lc, int len,
                                 unsigned int ptag, int k, double bq, double rel) {
    constexpr int ROUND = TopK32<CAP32>::ROUND;
    constexpr int CPT = (M == 8) ? 2 : 1;         // candidates per 128-bit load
    constexpr int NE = ROUND / (MMIDX_NT * CPT);  // loads per thread per round
    const int tid = threadIdx.x;
    for (int base = 0; base < len; base += ROUND) {
        c32.maybe_compact(k, bq, rel);  // round barrier
        const float thr32 = c32.thr32;
        uint4 cw[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const int i0 = base + (e * MMIDX_NT + tid) * CPT;
            cw[e] = make_uint4(0, 0, 0, 0);
            if (i0 < len) cw[e] = ld_nc_u4(lc + (int64_t)i0 * M);
        }
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            // warp-uniform skip of the list tail (push() only synchronises within the warp)
            if (base + (e * MMIDX_NT + (tid & ~31)) * CPT >= len) continue;
            const int i0 = base + (e * MMIDX_NT + tid) * CPT;
            float d0, d1;
            adc_pair<M>(lut, cw[e], d0, d1);
            c32.push((i0 < len) && d0 <= thr32, d0, ptag + (unsigned int)i0);
            if (M == 8) c32.push((i0 + 1 < len) && d1 <= thr32, d1, ptag + (unsigned int)(i0 + 1));
        }
    }
    __syncthreads();  // every push of this list is visible
}

// grid (nsplit, nq).  CTA (s, q) handles owned probe slots s, s+nsplit, ... of query q in rank order.
// Shared memory: TopK32 | t2 | lut[2] | qv | mbarrier | the query's probe descriptors (when the host grants the room).
// Per probe: one descriptor read, wait for the TMA of the probe's T1 row -- it lands straight in the table
// buffer the previous probe is not using -- then the fp32 table lut = (T1[l] + T2) + s is built IN PLACE with 128-bit shared
// accesses, ONE block barrier (which also settles the collector), the TMA of the next T1 row into the other buffer (its
// last reader, the previous probe's sweep, ended at that barrier), then a barrier-free sweep over the list with the next
// 128-bit code load in flight.  (No staging buffer: 36 KB per CTA at m = 8, 69 KB at m = 16 -- three CTAs per SM instead of two;
// the m = 16 instantiation is compiled for exactly those three, i.e. with 80 registers.)
// LONG: some list is longer than FAST_SEG entries (the pseudo lists of a large flat PQ index: 10^5 entries).  Such lists are
// swept in segments of FAST_SEG entries with a settle in between, so that the admission threshold is re-read while it
// tightens and a collector overflow re-scans one segment, not the list (without this a 1 GiB flat scan ran at 25 % of the
// HBM peak, with it at 45 % for one query and 69 % for eight: profiles/README.md).  IVF lists are short: the LONG = false
// instantiation is the kernel exactly as tuned for them.  (Also measured and rejected for the HBM-streaming regime: four
// 128-bit loads in flight per thread at 3 CTAs/SM -- slower at every batch size.)
constexpr int FAST_SEG = 8192;

template <int CAP32, int M, bool LONG>
__global__ void __launch_bounds__(MMIDX_NT, (M == 16 && MMIDX_SCAN_LB16 == 3) ? 3 : 4) k_ivfpq_scan_fast(FastArgs a, TopkOut o) {
    constexpr int ECAP = FastExactCap<CAP32, M>::value;
    constexpr int ks = 256;  // the fused kernel is specialised for full byte codes (host checks ks == 256)
    constexpr int nent = M * ks;
    constexpr int NV = nent / (4 * MMIDX_NT);  // float4 entries of one table per thread
    constexpr int DSTRIDE = fast_desc_stride(M);
    constexpr int KEEP = TopK32<CAP32>::KEEP_MAX;
    constexpr unsigned POS_MASK = (1u << FAST_POS_BITS) - 1u;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TopK32<CAP32> &c32 = *reinterpret_cast<TopK32<CAP32> *>(smem_raw);
    const size_t c32_bytes = (sizeof(TopK32<CAP32>) + 127) & ~(size_t)127;
    // region A (scan phase): t2 | lut0 | lut1.   Aliased in the final phase by: TopK<ECAP> | s_l | xs
    unsigned char *regA = smem_raw + c32_bytes;
    float *t2 = reinterpret_cast<float *>(regA);   // [nent]
    float *lut0 = t2 + nent;                       // [nent]  TMA target of the even probes, table built in place
    float *lut1 = lut0 + nent;                     // [nent]  ... of the odd probes
    const uint32_t lut0_s = smem_u32(lut0), lut1_s = smem_u32(lut1);
    const size_t tk_bytes = (sizeof(TopK<ECAP>) + 127) & ~(size_t)127;
    // final phase needs room for at least one survivor's terms (host: fast_smem_bytes)
    const size_t fin_bytes = tk_bytes + ECAP * sizeof(int) + (size_t)M * (a.S + 1) * sizeof(double);
    const size_t regA_bytes = max((size_t)3 * nent * sizeof(float), fin_bytes);
    double *qv = reinterpret_cast<double *>(regA + ((regA_bytes + 15) & ~(size_t)15));  // [d] raw query
    uint64_t *bars = reinterpret_cast<uint64_t *>(qv + a.d);                            // [1]

    const int tid = threadIdx.x;
    const int s = blockIdx.x;
    const int64_t q = a.qorder ? (int64_t)a.qorder[blockIdx.y] : (int64_t)blockIdx.y;
    const unsigned char *dqg = a.desc + q * (int64_t)a.w * DSTRIDE;
    const uint32_t t1_bytes = (uint32_t)(nent * sizeof(float));
    const int S = a.S;
    constexpr double rel = (double)M * 5.9604644775390625e-08;
    const float finf = __int_as_float(0x7f800000);

    const double bq_q = a.bq[q];
    const bool out_of_range = !(bq_q == bq_q);  // k_fast_prep: the query is outside the fp32-safe window
    const int nop = out_of_range ? 0 : a.ocnt[q];
    c32.init();
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_fence_init();
        if (s < nop) {
            const int l0 = reinterpret_cast<const ProbeHdr *>(dqg + (int64_t)s * DSTRIDE)->l;
            mbar_arrive_expect_tx(&bars[0], t1_bytes);
            tma_load_1d(lut0, a.T1 + (int64_t)l0 * nent, t1_bytes, &bars[0]);
        }
    }
    for (int i = tid; i < a.d; i += MMIDX_NT) qv[i] = a.Q[q * (int64_t)a.d + i];
    // The query's probe descriptors, staged once: a probe then starts with a shared-memory read instead of a dependent
    // L2 round trip (the host grants the room only where it does not cost a resident CTA).
    const unsigned char *dq = dqg;
    if (a.desc_stage) {
        uint4 *dsm = reinterpret_cast<uint4 *>(bars + 2);
        const uint4 *src = reinterpret_cast<const uint4 *>(dqg);
        for (int i = tid; i < nop * (DSTRIDE / 16); i += MMIDX_NT) dsm[i] = src[i];
        dq = reinterpret_cast<const unsigned char *>(dsm);
    }
    __syncthreads();
    // ---- per-query prologue: this query's T2 row (k_fast_t2) into shared memory.  Thread tid owns the float4
    //      entries tid + 256*r of every table (t2, stage, lut): it only ever re-reads what it wrote itself. ----
    float4 *t2v = reinterpret_cast<float4 *>(t2);
    {
        const float4 *t2g = reinterpret_cast<const float4 *>(a.T2 + q * (int64_t)nent);
#pragma unroll
        for (int r = 0; r < NV; ++r) t2v[tid + MMIDX_NT * r] = t2g[tid + MMIDX_NT * r];
    }

    // Settles the collector at a block barrier: if the barrier-free scan of slot `pslot` overflowed, its entries are
    // dropped and the list is scanned again with round barriers from its (still intact) table `plut`; then the buffer
    // is compacted when it is more than half full, or to get a first finite threshold.
    // Settles the collector at a block barrier: if the barrier-free scan of slot `pslot` (LONG: of its segment starting at
    // seg0) overflowed, its entries are dropped and it is scanned again with round barriers from its (still intact) table
    // `plut`; then the buffer is compacted when it is more than half full, or to get a first finite threshold.
    constexpr int SEG = FAST_SEG;
    auto settle = [&](int pslot, const uint32_t plut, int seg0) {
        const int need = (c32.ovf != 0) | (*(volatile int *)&c32.cnt > KEEP) |
                         ((c32.thr32 == finf) & (*(volatile int *)&c32.cnt >= a.k));
        if (__syncthreads_or(need)) {
            const int ov = c32.ovf;
            __syncthreads();
            if (ov) {  // block-uniform
                if (a.stats && tid == 0) atomicAdd(&a.stats[1], 1ull);
                c32.drop_tag((unsigned int)pslot, (unsigned int)seg0);
                const ProbeHdr *h = reinterpret_cast<const ProbeHdr *>(dq + (int64_t)pslot * DSTRIDE);
                scan_list_rounds<CAP32, M>(c32, plut, a.ocodes + (h->start + seg0) * M, LONG ? min(SEG, h->len - seg0) : h->len,
                                           (((unsigned int)pslot) << FAST_POS_BITS) + (unsigned int)seg0, a.k, a.bq[q], rel);
            }
            const int n = c32.cnt;
            const bool isinf32 = c32.thr32 == finf;
            __syncthreads();  // everybody has read the state before the next list starts pushing
            if (n >= a.k && (n > KEEP || isinf32)) c32.compact(a.k, a.bq[q], rel);
        }
    };

    int it = 0;
    int last_seg0 = 0;  // first position of the segment whose settle is still pending (the previous probe's last one)
    for (int ii = s; ii < nop; ii += a.nsplit, ++it) {
        float *lut = (it & 1) ? lut1 : lut0;
        const uint32_t lut_s = (it & 1) ? lut1_s : lut0_s;
        const unsigned char *dp = dq + (int64_t)ii * DSTRIDE;
        const int4 hd = *reinterpret_cast<const int4 *>(dp);  // start (lo, hi), len, probe rank
        float sj[NV];
#pragma unroll
        for (int r = 0; r < NV; ++r) sj[r] = reinterpret_cast<const float *>(dp + sizeof(ProbeHdr))[(tid + MMIDX_NT * r) >> 6];
        int lnext = 0;
        if (tid == 0 && ii + a.nsplit < nop) lnext = reinterpret_cast<const ProbeHdr *>(dp + (int64_t)a.nsplit * DSTRIDE)->l;
        const int64_t start = (int64_t)(((unsigned long long)(unsigned int)hd.y << 32) | (unsigned long long)(unsigned int)hd.x);
        const int len = hd.z;  // > 0 by construction of the descriptors
        const uint8_t *lc = a.ocodes + start * M;
        uint4 first = make_uint4(0, 0, 0, 0);
        if (MMIDX_SCAN_HOIST16 && M == 16 && tid < len) first = ld_nc_u4(lc + (int64_t)tid * M);
        mbar_wait(&bars[0], (uint32_t)(it & 1));
        // ADC table of this probe: lut = (T1[l] + T2) + s
        {
            float4 *lutv = reinterpret_cast<float4 *>(lut);
#pragma unroll
            for (int r = 0; r < NV; ++r) {
                const int e4 = tid + MMIDX_NT * r;
                const float4 sv = lutv[e4], tv = t2v[e4];  // sv: the T1 row the TMA delivered into this buffer
                float4 ov;
                ov.x = (sv.x + tv.x) + sj[r];
                ov.y = (sv.y + tv.y) + sj[r];
                ov.z = (sv.z + tv.z) + sj[r];
                ov.w = (sv.w + tv.w) + sj[r];
                lutv[e4] = ov;
            }
        }
        // the probe's barrier: publishes the table, ends every read of `stage` and every push of the previous list
        settle(ii - a.nsplit, (it & 1) ? lut0_s : lut1_s, last_seg0);
        if (tid == 0 && ii + a.nsplit < nop) {
            // the other buffer's last readers (the previous probe's sweep and its possible re-scan in settle) are done:
            // the next probe's T1 row lands there while this list is scanned
            fence_proxy_async();
            mbar_arrive_expect_tx(&bars[0], t1_bytes);
            tma_load_1d((it & 1) ? lut0 : lut1, a.T1 + (int64_t)lnext * nent, t1_bytes, &bars[0]);
        }
        const unsigned int ltag = ((unsigned int)ii) << FAST_POS_BITS;
        const uint8_t *list_codes = lc;
        const int list_len = len;
        for (int seg0 = 0; seg0 < (LONG ? list_len : 1); seg0 += SEG) {
            if (LONG) {
                if (seg0 > 0) settle(ii, lut_s, seg0 - SEG);  // block-uniform: the previous segment of this list
                last_seg0 = seg0;
            }
            const uint8_t *lc = LONG ? list_codes + (int64_t)seg0 * M : list_codes;
            const int len = LONG ? min(SEG, list_len - seg0) : list_len;
            const unsigned int ptag = LONG ? ltag + (unsigned int)seg0 : ltag;
            const float thr32 = c32.thr32;
            if (thr32 == finf) {  // block-uniform: nothing can be rejected yet
                scan_list_rounds<CAP32, M>(c32, lut_s, lc, len, ptag, a.k, a.bq[q], rel);
            } else {
                // One 128-bit load per thread and step; the load of the next step is issued before the current one is
                // consumed.  (Measured alternatives, profiles/README.md: L1 prefetch one or two steps ahead, two or four
                // loads per step, 3 CTAs/SM with 80 registers -- all slower than this form.)
                // Two steps per trip with the roles of the two code registers swapped, so that nothing is moved between
                // them (SASS: 72 -> 60 instructions per step); a lane past the end of the list keeps a stale word -- every
                // byte is a valid table index and the (i < len) predicate gates the offer.
                constexpr int CPT = (M == 8) ? 2 : 1;
                constexpr int STEP = MMIDX_NT * CPT;
                int i0 = tid * CPT;
                const uint8_t *pc = lc + (int64_t)i0 * M;
                // (requesting the list's first word before the table build, so that its latency overlaps the build and the
                //  barrier, was measured: 1.5 % slower -- four more live registers across the settle)
                uint4 cur = first, nxt = make_uint4(0, 0, 0, 0);
                if (!(MMIDX_SCAN_HOIST16 && M == 16 && seg0 == 0) && i0 < len) cur = ld_nc_u4(pc);
                auto offer = [&](const uint4 cw, const int i) {
                    float d0, d1;
                    adc_pair<M>(lut_s, cw, d0, d1);
                    const bool p0 = (i < len) && d0 <= thr32;
                    const bool p1 = (M == 8) && (i + 1 < len) && d1 <= thr32;
                    if (__any_sync(0xffffffffu, p0 | p1)) c32.push2(p0, d0, p1, d1, ptag | (unsigned int)i);
                };
                for (int wb = (tid & ~31) * CPT; wb < len; wb += 2 * STEP) {  // warp-uniform trip count
                    if (i0 + STEP < len) nxt = ld_vol_u4(pc + (int64_t)STEP * M);
                    offer(cur, i0);
                    if (wb + STEP >= len) break;  // warp-uniform
                    if (i0 + 2 * STEP < len) cur = ld_vol_u4(pc + (int64_t)2 * STEP * M);
                    offer(nxt, i0 + STEP);
                    i0 += 2 * STEP;
                    pc += (int64_t)2 * STEP * M;
                }
            }
        }
    }
    if (it > 0) settle(s + (it - 1) * a.nsplit, ((it - 1) & 1) ? lut1_s : lut0_s, last_seg0);

    // ---- final phase: shrink to the error band, evaluate the survivors exactly, exact top-k ----
    __syncthreads();
    const int n_before = c32.cnt;
    __syncthreads();
    if (n_before > a.k) c32.compact(a.k, a.bq[q], rel);
    const int nsurv = c32.cnt;
    const bool overflow = c32.overflow != 0 || nsurv > ECAP || out_of_range;
    TopK<ECAP> &tk = *reinterpret_cast<TopK<ECAP> *>(regA);  // aliases t2/stage/lut: no TMA is in flight any more
    int *s_l = reinterpret_cast<int *>(regA + tk_bytes);                      // [ECAP] list id of every survivor
    double *xs = reinterpret_cast<double *>(regA + tk_bytes + ECAP * sizeof(int));  // [NB][M][S + 1] squared terms
    uint8_t *scode = reinterpret_cast<uint8_t *>(c32.key);                    // [ECAP][M]; the fp32 keys are dead now
    static_assert(sizeof(c32.key) >= (size_t)ECAP * M, "survivor codes are staged in the fp32 key array");
    tk.init();
    if (!overflow) {
        // Exact binary64 distance of every survivor in the reference's operation order (computeResidualVector
        // IVFPQ.java:642-648, computeLookupADC :525-538, ADC sum :435-438).
        // A: one thread per survivor gathers its code, list id, offer sequence and iid (slot e of the exact collector).
        for (int e = tid; e < nsurv; e += MMIDX_NT) {
            const unsigned int packed = c32.pk[e];
            const ProbeHdr *h = reinterpret_cast<const ProbeHdr *>(dq + (int64_t)(packed >> FAST_POS_BITS) * DSTRIDE);
            const int64_t ps = h->start + (int64_t)(packed & POS_MASK);
            if (M == 8)
                reinterpret_cast<uint2 *>(scode)[e] = *reinterpret_cast<const uint2 *>(a.ocodes + ps * M);
            else
                reinterpret_cast<uint4 *>(scode)[e] = *reinterpret_cast<const uint4 *>(a.ocodes + ps * M);
            s_l[e] = h->l;
            tk.seq[e] = (((unsigned long long)(unsigned int)h->p) << 32) | (unsigned long long)(unsigned int)a.orank[ps];
            tk.pay[e] = a.oiids[ps];
        }
        __syncthreads();
        // B: batches of NB survivors.  thread <-> element (survivor, j, t) with coalesced loads of the centroid and
        //    codebook rows: x = ((C_l - q) - P_j,c)[t]^2.  C: thread <-> (survivor, j) adds its S terms for t ascending
        //    (= LUT[j][code_j]).  D: thread <-> survivor adds the m table entries for j ascending.
        const int d = a.d;
        const int row = S + 1;
        const int NB = (int)((regA_bytes - tk_bytes - ECAP * sizeof(int)) / ((size_t)M * row * sizeof(double)));
        const int shS = ((S & (S - 1)) == 0) ? (31 - __clz(S)) : -1;
        const int shD = ((d & (d - 1)) == 0) ? (31 - __clz(d)) : -1;
        for (int b0 = 0; b0 < nsurv; b0 += NB) {
            const int nb = min(NB, nsurv - b0);
            const int total = nb * d;
            if (d <= MMIDX_NT && MMIDX_NT % d == 0) {
                // every thread keeps ONE element index (j, t): only the survivor changes from pass to pass
                const int epb = MMIDX_NT / d;
                const int e0 = (shD >= 0) ? (tid >> shD) : (tid / d);
                const int i = tid - e0 * d;
                const int j = (shS >= 0) ? (i >> shS) : (i / S);
                const int t = i - j * S;
                int src = i;
                if (a.perm) src = a.perm[i];
                const double qs = qv[src];
                const double *Cs = a.C + src;
                const double *Pj = a.P + (int64_t)j * ks * S + t;
                // (unrolling this loop four times for more loads in flight was measured: 1 % slower)
                for (int e = e0; e < nb; e += epb) {
                    const unsigned int code = scode[(b0 + e) * M + j];
                    const double r = __dsub_rn(Cs[(int64_t)s_l[b0 + e] * d], qs);  // residual = centroid - query
                    const double df = __dsub_rn(r, Pj[code * S]);
                    xs[(e * M + j) * row + t] = __dmul_rn(df, df);
                }
            } else {
                for (int g = tid; g < total; g += MMIDX_NT) {
                    const int e = (shD >= 0) ? (g >> shD) : (g / d);
                    const int i = g - e * d;
                    const int j = (shS >= 0) ? (i >> shS) : (i / S);
                    const int t = i - j * S;
                    int src = i;
                    if (a.perm) src = a.perm[i];
                    const unsigned int code = scode[(b0 + e) * M + j];
                    const double r = __dsub_rn(a.C[(int64_t)s_l[b0 + e] * d + src], qv[src]);  // residual = centroid - query
                    const double df = __dsub_rn(r, a.P[((int64_t)j * ks + code) * S + t]);
                    xs[(e * M + j) * row + t] = __dmul_rn(df, df);
                }
            }
            __syncthreads();
            for (int item = tid; item < nb * M; item += MMIDX_NT) {
                double *xr = xs + item * row;
                double acc = 0.0;
                for (int t = 0; t < S; ++t) acc = __dadd_rn(acc, xr[t]);
                xr[S] = acc;
            }
            __syncthreads();
            for (int e = tid; e < nb; e += MMIDX_NT) {
                double dv = 0.0;
#pragma unroll
                for (int j = 0; j < M; ++j) dv = __dadd_rn(dv, xs[(e * M + j) * row + S]);
                tk.dist[b0 + e] = dv;
            }
            __syncthreads();
        }
        if (tid == 0) tk.cnt = nsurv;
        __syncthreads();
    }
    // The survivors (~k + the error band) are sorted as a whole; the result is their first k entries unless an exact
    // binary64 tie is cut at the k-th boundary (rare), which takes the queue-order path below.
    bool written = false;
    if (!overflow) {
        tk.sort_first(nsurv);
        const bool cut = nsurv > a.k && tk.dist[a.k - 1] == tk.dist[a.k];  // block-uniform
        if (!cut || !a.resolve_ties) {
            // sharded / split mode: a cut tie is only flagged; the merge replays the queue over all parts
            write_sorted(tk, o, q, s, a.k, min(nsurv, a.k), cut);
            written = true;
        }
    }
    // Exact ties cut at the k-th boundary: without overflow the survivors contain EVERY candidate with an exact
    // distance <= T (a candidate outside the band is strictly farther than T), so the queue's rule
    // (tie_resolve.cuh) can be replayed here on the survivors alone: among the first k entries with dist <= T in
    // offer order, the latest-offered tied entries survive.  Losers get dist = +inf and drop out in finalize().
    if (!written && !overflow) {  // a.resolve_ties && a tie cut at the boundary
        // Without overflow the survivors contain EVERY candidate with an exact distance <= T (a candidate outside
        // the band is strictly farther than T), so the queue's rule can be replayed on the survivors alone.
        tk.kill_tie_losers(nsurv, a.k, reinterpret_cast<int *>(c32.key));  // the fp32 keys are dead: scratch [nsurv]
    }
    if (a.stats && tid == 0) {
        unsigned long long n_cand = 0;
        for (int ii = s; ii < nop; ii += a.nsplit) n_cand += (unsigned long long)reinterpret_cast<const ProbeHdr *>(dq + (int64_t)ii * DSTRIDE)->len;
        atomicAdd(&a.stats[0], n_cand);
        atomicAdd(&a.stats[2], (unsigned long long)nsurv);
        if (overflow) atomicAdd(&a.stats[3], 1ull);
    }
    if (overflow && tid == 0) {
        const int slot = atomicAdd(a.fb_count, 1);
        a.fb_list[slot] = (int32_t)(q * a.nsplit + s);
    }
    // on overflow an empty result is written here and the direct kernel overwrites it
    if (!written) write_result(tk, o, q, s, a.k, -1.0);
}

// Direct (table-free) exact scan for the rare items the fast kernel could not finish: every candidate of the
// item's probes is evaluated in binary64 from the quantizers.  grid: any; CTAs stride over fb_list.
template <int CAP>
__global__ void __launch_bounds__(MMIDX_NT) k_ivfpq_scan_direct(FastArgs a, TopkOut o) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TopK<CAP> &tk = *reinterpret_cast<TopK<CAP> *>(smem_raw);
    constexpr int ROUND = TopK<CAP>::ROUND;
    const int nfb = *a.fb_count;
    for (int fi = blockIdx.x; fi < nfb; fi += gridDim.x) {
        const int item = a.fb_list[fi];
        const int64_t q = item / a.nsplit;
        const int s = item - (int)q * a.nsplit;
        const int32_t *pr = a.probes + q * a.w;
        const double *qv = a.Q + q * (int64_t)a.d;
        const int32_t *op = a.oprobes + q * a.w;  // same probe subset as the fast kernel's CTA (s, q)
        const int nop = a.ocnt[q];
        __syncthreads();
        tk.init();
        for (int ii = s; ii < nop; ii += a.nsplit) {
            const int p = op[ii];
            const int l = pr[p];
            const int64_t start = a.list_off[l];
            const int len = a.list_len[l];
            const double *Cl = a.C + (int64_t)(a.flat ? 0 : l) * a.d;
            for (int base = 0; base < len; base += ROUND) {
                tk.maybe_compact(a.k);
                const double thr = tk.thr;
                const bool strict = tk.strict != 0;
                for (int e = 0; e < ROUND / MMIDX_NT; ++e) {
                    const int i = base + e * MMIDX_NT + threadIdx.x;
                    const bool valid = i < len;
                    double dv = 0.0;
                    if (valid) dv = exact_adc(Cl, qv, a.perm, a.P, a.codes + (start + i) * a.m, a.m, a.ks, a.S);
                    const bool pred = valid && (dv < thr || (dv == thr && !strict));
                    tk.push(pred, dv, (((unsigned long long)p) << 32) | (unsigned long long)i, pred ? a.iids[start + i] : 0);
                }
            }
        }
        write_result(tk, o, q, s, a.k, -1.0);
    }
}

// tie pass collector without ADC tables in memory: exact distances straight from the quantizers.
struct TieDirectArgs {
    const double *Q, *C, *P;
    const int32_t *perm;
    const int32_t *probes;
    const uint8_t *codes;
    const int32_t *iids;
    const int64_t *list_off;
    const int32_t *list_len;
    int d, m, ks, S, w, k, code_bytes;
    int flat;
};

__global__ void __launch_bounds__(MMIDX_NT) k_tie_collect_ivfpq_direct(TieDirectArgs a, const double *__restrict__ res_dist,
                                                                       const int32_t *__restrict__ amb_list,
                                                                       const int32_t *__restrict__ amb_count, TieLists o) {
    __shared__ int warp_sums[MMIDX_NT / 32];
    const int na = *amb_count;
    for (int ai = blockIdx.x; ai < na; ai += gridDim.x) {
        const int64_t q = amb_list[ai];
        const double T = res_dist[q * a.k + a.k - 1];
        const double *qv = a.Q + q * (int64_t)a.d;
        int found = 0;
        for (int p = 0; p < a.w && found < a.k; ++p) {
            const int l = a.probes[q * a.w + p];
            const int64_t start = a.list_off[l];
            const double *Cl = a.C + (int64_t)(a.flat ? 0 : l) * a.d;
            const uint8_t *cp = a.codes + start * a.code_bytes;
            const int32_t *li = a.iids + start;
            const TieDirectArgs &r = a;
            tie_sweep(a.list_len[l], ((unsigned long long)p) << 32, T, a.k, found, warp_sums, o, q,
                      [=](int64_t i) { return exact_adc(Cl, qv, r.perm, r.P, cp + i * r.code_bytes, r.m, r.ks, r.S); },
                      [li](int64_t i) { return li[i]; });
        }
        if (threadIdx.x == 0) o.cnt[q] = min(found, a.k);
        __syncthreads();
    }
}

}  // namespace mmidx
