// comm.cuh -- multi-GPU exchange of libmmidx over peer memory (NVLink 5 / NVSwitch P2P), one process per GPU.
//
// The reference keeps ONE BoundedPriorityQueue for all probed lists of a query (IVFPQ.java:409,445).  With the
// inverted lists sharded over S GPUs (BASELINE north_star) a query's queue is the merge of S per-shard queues, so a
// search step has real exchange points.  They are NOT collective launches here: every rank owns an "exchange window"
// in its HBM that its peers map (CUDA IPC), and the kernels that produce a row (coarse verification, fused scan,
// merge, tie finish) store it straight into the window of the rank that consumes it (PeerSink, kernels.cuh).  What
// is left of a collective is its synchronisation: k_comm_sync publishes "my stores of epoch e have landed" into each
// peer's flag row and waits for the peers' flags -- one tiny kernel per exchange point, CUDA-graph friendly because the
// epoch lives in device memory.
//
// Layout of G = S x R ranks: rank = group * S + shard.  A group holds the whole index (S list shards) and serves its
// own batch of gq queries; sl = ceil(gq / S) consecutive queries form the slice whose final queue shard t merges.
//   stage 0  probes   each shard runs the coarse stage for its slice, rows -> every group peer        (BCAST)
//   stage 1  partials each shard scans its lists for all gq queries, row of query q -> owner q / sl   (ROUTE)
//   stage 2  ties     owner publishes (query, T) of the slice queries whose k-th boundary tie was cut (normally none)
//   stage 3  lists    every shard's first-k entries with dist <= T in offer order -> owner          (tie_resolve.cuh)
//   stage 4  final    optional: merged rows -> every rank of the job                                  (BCAST)
// Every data region exists twice (epoch parity): a rank can run at most one step ahead of a peer, because passing an
// exchange point of step e+1 needs the peer's flag of step e+1, which it raises after finishing step e.
#pragma once
#include "common.cuh"
#include "fast_scan.cuh"
#include "kernels.cuh"
#include "tie_resolve.cuh"

namespace mmidx {

constexpr int COMM_NSTAGE = 5;

struct WinLayout {
    // absolute offsets
    size_t flags = 0;          // unsigned [COMM_NSTAGE][MMIDX_MAX_PEERS]
    size_t epoch = 0;          // unsigned [1]   (only the owner touches it)
    size_t data0 = 0;          // parity-0 data; parity 1 at data0 + parity_stride
    size_t parity_stride = 0;
    size_t total = 0;
    // offsets relative to a parity base
    size_t probes = 0;                                        // int32 [gqp][w]
    size_t p_iids = 0, p_dist = 0, p_seq = 0, p_cnt = 0, p_tie = 0;  // partials [S][sl][k] / [S][sl]
    size_t a_cnt = 0, a_q = 0, a_T = 0;                       // published ties: int32 [S], int32 [S][sl], double [S][sl]
    size_t t_seq = 0, t_pay = 0, t_eq = 0, t_cnt = 0;         // tie lists [S][sl][k] / [S][sl]
    size_t o_iids = 0, o_dist = 0, o_cnt = 0;                 // final rows [R * gqp][k] / [R * gqp]
    int64_t sl_max = 0, gqp_max = 0;
};

static inline size_t win_align(size_t x) { return (x + 255) & ~(size_t)255; }

// identical on every rank: a pure function of the job geometry
static inline WinLayout make_layout(int S, int R, int64_t max_gq, int k_max, int w_max) {
    WinLayout L;
    L.sl_max = (max_gq + S - 1) / S;
    L.gqp_max = L.sl_max * S;
    size_t o = 0;
    L.flags = o;
    o = win_align(o + sizeof(unsigned) * COMM_NSTAGE * MMIDX_MAX_PEERS);
    L.epoch = o;
    o = win_align(o + sizeof(unsigned));
    L.data0 = o;
    size_t r = 0;
    auto take = [&](size_t bytes) {
        size_t at = r;
        r = win_align(r + bytes);
        return at;
    };
    const size_t rows = (size_t)L.gqp_max;  // S * sl_max
    L.probes = take(rows * (size_t)w_max * 4);
    if (S > 1) {
        L.p_iids = take(rows * k_max * 4);
        L.p_dist = take(rows * k_max * 8);
        L.p_seq = take(rows * k_max * 8);
        L.p_cnt = take(rows * 4);
        L.p_tie = take(rows * 8);
        L.a_cnt = take((size_t)S * 4);
        L.a_q = take(rows * 4);
        L.a_T = take(rows * 8);
        L.t_seq = take(rows * k_max * 8);
        L.t_pay = take(rows * k_max * 4);
        L.t_eq = take(rows * k_max * 4);
        L.t_cnt = take(rows * 4);
    }
    const size_t orows = rows * (size_t)R;
    L.o_iids = take(orows * k_max * 4);
    L.o_dist = take(orows * k_max * 8);
    L.o_cnt = take(orows * 4);
    L.parity_stride = r;
    L.total = L.data0 + 2 * r;
    return L;
}

// ---- synchronisation ------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// first kernel of a step: a new epoch
__global__ void k_comm_begin(unsigned *epoch) { *epoch = *epoch + 1u; }

struct CommSync {
    unsigned *remote[MMIDX_MAX_PEERS];       // my slot in the flag row of peer t
    const unsigned *local[MMIDX_MAX_PEERS];  // peer t's slot in my flag row
    const unsigned *epoch;
    int n;
    unsigned long long timeout_ns;
};

// One exchange point.  Every store of the preceding kernels of this stream has been performed (kernel boundary), so the
// release store of the epoch orders them before the flag; then wait until every peer has done the same.  A peer that
// never arrives (crashed rank) trips the timeout and traps instead of hanging the GPU.
__global__ void k_comm_sync(CommSync c) {
    const int t = threadIdx.x;
    if (t < c.n) {
        const unsigned e = *c.epoch;
        __threadfence_system();
        st_release_sys(c.remote[t], e);
        const unsigned long long t0 = global_ns();
        while ((int)(ld_acquire_sys(c.local[t]) - e) < 0) {
            if (global_ns() - t0 > c.timeout_ns) {
                printf("libmmidx: peer %d did not reach exchange epoch %u within %llu s\n", t, e, c.timeout_ns / 1000000000ull);
                __trap();
            }
        }
    }
}

// ---- stage 2: the slice owner publishes its cut ties to the group ----------------------------------------------
struct AmbPublish {
    const int32_t *amb_list;   // slice-local query ids flagged by the merge
    const int32_t *amb_count;
    const double *res_dist;    // merged rows of the slice [sl][k]
    unsigned char *base[MMIDX_MAX_PEERS];  // windows of all S group members (parity applied), self included
    long long off_cnt, off_q, off_T;
    int S, shard, sl, k;
};

__global__ void __launch_bounds__(MMIDX_NT) k_comm_publish_ties(AmbPublish a) {
    const int na = *a.amb_count;
    for (int t = 0; t < a.S; ++t) {
        unsigned char *b = a.base[t];
        if (threadIdx.x == 0) reinterpret_cast<int32_t *>(b + a.off_cnt)[a.shard] = na;
        int32_t *dq = reinterpret_cast<int32_t *>(b + a.off_q) + (long long)a.shard * a.sl;
        double *dT = reinterpret_cast<double *>(b + a.off_T) + (long long)a.shard * a.sl;
        for (int i = threadIdx.x; i < na; i += MMIDX_NT) {
            const int ql = a.amb_list[i];
            dq[i] = a.shard * a.sl + ql;  // group query id
            dT[i] = a.res_dist[(long long)ql * a.k + a.k - 1];
        }
    }
}

// ---- stage 3: first-k entries with dist <= T in offer order, over the lists THIS shard stores -------------------
struct TieMultiArgs {
    TieDirectArgs t;           // quantizers, group queries, probes of all group queries, this shard's CSR
    const int32_t *a_cnt;      // my window: [S] ties published by every owner
    const int32_t *a_q;        // [S][sl]
    const double *a_T;         // [S][sl]
    unsigned char *base[MMIDX_MAX_PEERS];  // windows of all S group members (parity applied)
    long long off_seq, off_pay, off_eq, off_cnt;
    int S, shard, sl;
};

// Per probed list the exact binary64 ADC table is rebuilt in shared memory (same arithmetic as k_lut_build:
// computeResidualVector IVFPQ.java:642-648 + computeLookupADC :525-538) and the list is swept in offer order with m lookups
// per candidate -- ~50 us per flagged query instead of re-deriving every candidate from the quantizers.  use_lut == 0
// (table larger than the shared memory the host granted): table-free evaluation.
__global__ void __launch_bounds__(MMIDX_NT) k_tie_collect_multi(TieMultiArgs a, int use_lut) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int warp_sums[MMIDX_NT / 32];
    const TieDirectArgs &r = a.t;
    double *rv = reinterpret_cast<double *>(smem_raw);  // [d]   transformed residual of the current probe
    double *lut = rv + r.d;                             // [m][ks]
    for (int owner = 0; owner < a.S; ++owner) {
        const int na = a.a_cnt[owner];
        unsigned char *b = a.base[owner];
        TieLists o;
        // rows of this shard inside the owner's lists: (shard * sl + slice-local query)
        o.seq = reinterpret_cast<unsigned long long *>(b + a.off_seq) + (long long)a.shard * a.sl * r.k;
        o.pay = reinterpret_cast<int32_t *>(b + a.off_pay) + (long long)a.shard * a.sl * r.k;
        o.eq = reinterpret_cast<int32_t *>(b + a.off_eq) + (long long)a.shard * a.sl * r.k;
        o.cnt = reinterpret_cast<int32_t *>(b + a.off_cnt) + (long long)a.shard * a.sl;
        for (int ai = blockIdx.x; ai < na; ai += gridDim.x) {
            const int64_t q = a.a_q[(long long)owner * a.sl + ai];
            const double T = a.a_T[(long long)owner * a.sl + ai];
            const int64_t ql = q - (int64_t)owner * a.sl;
            const double *qv = r.Q + q * (int64_t)r.d;
            int found = 0;
            for (int p = 0; p < r.w && found < r.k; ++p) {
                const int l = r.probes[q * r.w + p];
                const int len = r.list_len[l];  // 0 for lists another shard stores
                if (len == 0) continue;         // block-uniform
                const int64_t start = r.list_off[l];
                const double *Cl = r.C + (int64_t)(r.flat ? 0 : l) * r.d;
                const uint8_t *cp = r.codes + start * r.code_bytes;
                const int32_t *li = r.iids + start;
                if (use_lut) {
                    __syncthreads();
                    for (int i = threadIdx.x; i < r.d; i += MMIDX_NT) {
                        const int src = r.perm ? r.perm[i] : i;
                        rv[i] = __dsub_rn(Cl[src], qv[src]);  // residual = centroid - query, then the permutation
                    }
                    __syncthreads();
                    for (int e = threadIdx.x; e < r.m * r.ks; e += MMIDX_NT) {
                        const int j = e / r.ks, c = e - j * r.ks;
                        const double *pc = r.P + ((int64_t)j * r.ks + c) * r.S;
                        double acc = 0.0;
                        for (int t = 0; t < r.S; ++t) acc = sqacc(acc, rv[j * r.S + t], pc[t]);
                        lut[e] = acc;
                    }
                    __syncthreads();
                    const int m = r.m, ks = r.ks, cb = r.code_bytes;
                    tie_sweep(len, ((unsigned long long)p) << 32, T, r.k, found, warp_sums, o, ql,
                              [=](int64_t i) { return adc_dist(lut, cp + i * cb, m, ks); }, [li](int64_t i) { return li[i]; });
                } else {
                    tie_sweep(len, ((unsigned long long)p) << 32, T, r.k, found, warp_sums, o, ql,
                              [=](int64_t i) { return exact_adc(Cl, qv, r.perm, r.P, cp + i * r.code_bytes, r.m, r.ks, r.S); },
                              [li](int64_t i) { return li[i]; });
                }
            }
            if (threadIdx.x == 0) o.cnt[ql] = min(found, r.k);
            __syncthreads();
        }
    }
}

}  // namespace mmidx
