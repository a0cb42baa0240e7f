// kernels.cuh -- the sm_100a kernels of libmmidx (exact binary64 path).
//
// K1a k_sqdist_matrix   coarse distances          IVFPQ.computeNearestCoarseIndices   IVFPQ.java:575-601
// K1b k_select_rows     top-w of each row         (BoundedPriorityQueue + poll() x w) IVFPQ.java:590-599
// K1c k_assign_nearest  argmin over centroids     computeNearestCoarseIndex           IVFPQ.java:547-564
//                                                 computeNearestCentroid              AFA.java:136-155
// K2  k_lut_build       residual + ADC table      computeResidualVector + computeLookupADC IVFPQ.java:642-648,525-538
// K3  k_ivfpq_scan      ADC scan + top-k          computeKnnIVFADC                    IVFPQ.java:429-446
//     k_pq_scan         flat ADC scan + top-k     computeKnnADC                       PQ.java:303-319
// K4  k_merge_topk      merge partial top-k       (one queue shared by all probes)    IVFPQ.java:409,445
// K6  k_pq_encode       PQ encode                 computeNearestProductIndex          PQ.java:411-429 IVFPQ.java:613-631
// K7  k_vlad_order / k_vlad_accumulate            VladAggregator.aggregateInternal    VladAggregator.java:56-70
// K8  k_linear_scan     exact kNN                 Linear.computeNearestNeighborsInternal Linear.java:138-163
//
// All early exits of the Java loops are result-neutral and are not reproduced; every sum is
// the full index-ascending binary64 sum, so results are bit-identical.
#pragma once
#include "common.cuh"
#include "topk.cuh"

namespace mmidx {

constexpr int QT = 8;    // rows of A handled per CTA in the centroid-distance kernels
constexpr int JC = 128;  // dimension chunk staged in shared memory

// ------------------------------------------------------------------------------------------------------
// Shared inner block: acc[qi] += (Bt[j][b] - A[a0+qi][j])^2 for j in [0,d), exact order.
// Bt is the centroid matrix stored transposed ([d][nb]) so that thread<->centroid loads coalesce.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_rows(double2 (*qs)[QT / 2], const double *__restrict__ A, int64_t a0,
                                           int64_t na, int d, int j0, int jn) {
    // qs[jj][qi]; consecutive threads walk jj so the global reads are contiguous per row
    double *q1 = reinterpret_cast<double *>(qs);
    for (int e = threadIdx.x; e < jn * QT; e += MMIDX_NT) {
        int qi = e / jn, jj = e - qi * jn;
        int64_t a = a0 + qi;
        q1[jj * QT + qi] = (a < na) ? A[a * (int64_t)d + j0 + jj] : 0.0;
    }
}

__device__ __forceinline__ void accumulate_chunk(double (&acc)[QT], const double2 (*qs)[QT / 2],
                                                 const double *__restrict__ Bt, int64_t nb, int64_t b, int j0,
                                                 int jn) {
#pragma unroll 4
    for (int jj = 0; jj < jn; ++jj) {
        double c = Bt[(int64_t)(j0 + jj) * nb + b];
#pragma unroll
        for (int h = 0; h < QT / 2; ++h) {
            double2 q = qs[jj][h];
            acc[2 * h] = sqacc(acc[2 * h], c, q.x);
            acc[2 * h + 1] = sqacc(acc[2 * h + 1], c, q.y);
        }
    }
}

// K1a: D[a][b] = sum_j (B[b][j] - A[a][j])^2.  grid (ceil(nb/256), ceil(na/QT)).
__global__ void __launch_bounds__(MMIDX_NT) k_sqdist_matrix(const double *__restrict__ A, const double *__restrict__ Bt,
                                                            int64_t na, int nb, int d, double *__restrict__ D) {
    __shared__ double2 qs[JC][QT / 2];
    const int64_t b = (int64_t)blockIdx.x * MMIDX_NT + threadIdx.x;
    const int64_t a0 = (int64_t)blockIdx.y * QT;
    double acc[QT];
#pragma unroll
    for (int i = 0; i < QT; ++i) acc[i] = 0.0;
    for (int j0 = 0; j0 < d; j0 += JC) {
        int jn = min(JC, d - j0);
        __syncthreads();
        stage_rows(qs, A, a0, na, d, j0, jn);
        __syncthreads();
        if (b < nb) accumulate_chunk(acc, qs, Bt, nb, b, j0, jn);
    }
    if (b < nb) {
#pragma unroll
        for (int qi = 0; qi < QT; ++qi)
            if (a0 + qi < na) D[(a0 + qi) * (int64_t)nb + b] = acc[qi];
    }
}

// K1c: out[a] = argmin_b sum_j (B[b][j] - A[a][j])^2, lowest index among equal minima
// (`distance < minDistance`, IVFPQ.java:558, AFA.java:149).  grid ceil(na/QT).
__global__ void __launch_bounds__(MMIDX_NT) k_assign_nearest(const double *__restrict__ A, const double *__restrict__ Bt,
                                                             int64_t na, int nb, int d, int32_t *__restrict__ out) {
    __shared__ double2 qs[JC][QT / 2];
    __shared__ double red_d[MMIDX_NT / 32][QT];
    __shared__ int red_i[MMIDX_NT / 32][QT];
    const int64_t a0 = (int64_t)blockIdx.x * QT;
    double best[QT];
    int bidx[QT];
#pragma unroll
    for (int i = 0; i < QT; ++i) {
        best[i] = 1.7976931348623157e308;  // Double.MAX_VALUE
        bidx[i] = 0x7fffffff;
    }
    for (int b0 = 0; b0 < nb; b0 += MMIDX_NT) {
        const int b = b0 + threadIdx.x;
        double acc[QT];
#pragma unroll
        for (int i = 0; i < QT; ++i) acc[i] = 0.0;
        for (int j0 = 0; j0 < d; j0 += JC) {
            int jn = min(JC, d - j0);
            __syncthreads();
            stage_rows(qs, A, a0, na, d, j0, jn);
            __syncthreads();
            if (b < nb) accumulate_chunk(acc, qs, Bt, nb, b, j0, jn);
        }
        if (b < nb) {
#pragma unroll
            for (int qi = 0; qi < QT; ++qi)
                if (acc[qi] < best[qi]) {
                    best[qi] = acc[qi];
                    bidx[qi] = b;
                }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int qi = 0; qi < QT; ++qi) {
        double v = best[qi];
        int ix = bidx[qi];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_down_sync(0xffffffffu, v, o);
            int oi = __shfl_down_sync(0xffffffffu, ix, o);
            if (ov < v || (ov == v && oi < ix)) {
                v = ov;
                ix = oi;
            }
        }
        if (lane == 0) {
            red_d[warp][qi] = v;
            red_i[warp][qi] = ix;
        }
    }
    __syncthreads();
    if (threadIdx.x < QT) {
        int qi = threadIdx.x;
        double v = red_d[0][qi];
        int ix = red_i[0][qi];
        for (int wv = 1; wv < MMIDX_NT / 32; ++wv) {
            double ov = red_d[wv][qi];
            int oi = red_i[wv][qi];
            if (ov < v || (ov == v && oi < ix)) {
                v = ov;
                ix = oi;
            }
        }
        if (a0 + qi < na) out[a0 + qi] = (ix == 0x7fffffff) ? -1 : ix;
    }
}

// ------------------------------------------------------------------------------------------------------
// Result sink shared by every top-k kernel: part-major arrays [nq][nparts][k].
//
// Multi-GPU (comm.cuh): a sink can also deliver the row straight into the HBM of peer GPUs (NVLink P2P stores into
// the peers' exchange windows), so the per-shard top-k / the final rows travel from the kernel that produces them --
// there is no pack kernel, no staging buffer and no collective launch on the data path.
//   mode 1 (ROUTE): the row of group query gq goes ONLY to the window of its slice owner, peer gq / sl
//   mode 2 (BCAST): the row is written locally (iids/dist/... below) AND copied to every peer in base[]
// ------------------------------------------------------------------------------------------------------
constexpr int MMIDX_MAX_PEERS = 8;
enum { SINK_IIDS = 1, SINK_DIST = 2, SINK_SEQ = 4, SINK_CNT = 8, SINK_TIE = 16 };

struct PeerSink {
    int mode;       // 0: local only
    int npeer;      // ROUTE: number of slice owners (list shards); BCAST: number of remote peers
    int sl;         // ROUTE: queries per owner slice
    int fields;     // SINK_* mask of the arrays the peers receive
    long long q0;   // index of this launch's query 0 inside the group batch
    long long row0; // ROUTE: first row of this shard's block in an owner's arrays; BCAST: row of group query 0
    unsigned char *base[MMIDX_MAX_PEERS];  // peer windows (parity offset applied by the host)
    long long off_iids, off_dist, off_seq, off_cnt, off_tie;  // byte offsets of the destination arrays in a window
};

struct TopkOut {
    int32_t *iids;            // [nq][nparts][k]
    double *dist;             // [nq][nparts][k]
    unsigned long long *seq;  // [nq][nparts][k] or NULL
    int32_t *cnt;             // [nq][nparts]
    double *tie;              // [nq][nparts] distance at which tied candidates were discarded, -1 if none; or NULL
    int32_t *amb_list;        // final stage only: queries whose k-th boundary tie was cut (or NULL)
    int32_t *amb_count;
    int nparts;
    PeerSink sink;            // nparts == 1 whenever sink.mode != 0
};

// copies row `row` (k entries taken from the getters) into the windows of peers [p0, p1)
template <typename GI, typename GD, typename GS>
__device__ __forceinline__ void sink_row(const PeerSink &ps, int p0, int p1, long long row, int k, int n, double tiev,
                                         GI gi, GD gd, GS gs) {
    for (int i = threadIdx.x; i < k; i += MMIDX_NT) {
        const bool v = i < n;
        const int32_t iv = v ? gi(i) : -1;
        const double dv = v ? gd(i) : __longlong_as_double(0x7ff0000000000000LL);
        const unsigned long long sv = v ? gs(i) : 0ull;
        for (int p = p0; p < p1; ++p) {
            unsigned char *b = ps.base[p];
            if (ps.fields & SINK_IIDS) reinterpret_cast<int32_t *>(b + ps.off_iids)[row * k + i] = iv;
            if (ps.fields & SINK_DIST) reinterpret_cast<double *>(b + ps.off_dist)[row * k + i] = dv;
            if (ps.fields & SINK_SEQ) reinterpret_cast<unsigned long long *>(b + ps.off_seq)[row * k + i] = sv;
        }
    }
    if (threadIdx.x == 0) {
        for (int p = p0; p < p1; ++p) {
            unsigned char *b = ps.base[p];
            if (ps.fields & SINK_CNT) reinterpret_cast<int32_t *>(b + ps.off_cnt)[row] = n;
            if (ps.fields & SINK_TIE) reinterpret_cast<double *>(b + ps.off_tie)[row] = tiev;
        }
    }
}

// destination of group query q of this launch: peers [p0, p1) and the row inside their arrays
__device__ __forceinline__ long long sink_dest(const PeerSink &ps, int64_t q, int &p0, int &p1) {
    const long long gq = q + ps.q0;
    if (ps.mode == 1) {
        p0 = (int)(gq / ps.sl);
        p1 = p0 + 1;
        return ps.row0 + (gq - (long long)p0 * ps.sl);
    }
    p0 = 0;
    p1 = ps.npeer;
    return ps.row0 + gq;
}

// writes the first n (<= k) entries of a collector that is already in output order (finalize() / sort_first())
template <int CAP>
__device__ void write_sorted(TopK<CAP> &tk, const TopkOut &o, int64_t q, int part, int k, int n, bool amb) {
    const double tiev = (n == k && amb) ? tk.dist[n - 1] : -1.0;
    if (o.sink.mode != 1) {
        const int64_t base = (q * o.nparts + part) * (int64_t)k;
        for (int i = threadIdx.x; i < k; i += MMIDX_NT) {
            bool v = i < n;
            o.iids[base + i] = v ? tk.pay[i] : -1;
            o.dist[base + i] = v ? tk.dist[i] : __longlong_as_double(0x7ff0000000000000LL);
            if (o.seq) o.seq[base + i] = v ? tk.seq[i] : 0ull;
        }
        if (threadIdx.x == 0) {
            o.cnt[q * o.nparts + part] = n;
            if (o.tie) o.tie[q * o.nparts + part] = tiev;
        }
    }
    if (threadIdx.x == 0 && o.amb_list && amb) {
        int slot = atomicAdd(o.amb_count, 1);
        o.amb_list[slot] = (int32_t)q;
    }
    if (o.sink.mode != 0) {
        int p0, p1;
        const long long row = sink_dest(o.sink, q, p0, p1);
        sink_row(o.sink, p0, p1, row, k, n, tiev, [&](int i) { return tk.pay[i]; }, [&](int i) { return tk.dist[i]; },
                 [&](int i) { return tk.seq[i]; });
    }
}

template <int CAP>
__device__ void write_result(TopK<CAP> &tk, const TopkOut &o, int64_t q, int part, int k, double extra_tie) {
    bool amb;
    const int n = tk.finalize(k, &amb);
    // a partial result that discarded ties at distance t makes the final answer ambiguous iff t == final T
    if (n == k && extra_tie == tk.dist[n - 1]) amb = true;
    write_sorted(tk, o, q, part, k, n, amb);
}

// K1b: top-w entries of each row of D in queue order. grid nq.
template <int CAP>
__global__ void __launch_bounds__(MMIDX_NT) k_select_rows(const double *__restrict__ D, int ncol, int w, TopkOut o) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TopK<CAP> &tk = *reinterpret_cast<TopK<CAP> *>(smem_raw);
    const int64_t q = blockIdx.x;
    const double *row = D + q * (int64_t)ncol;
    tk.init();
    for (int base = 0; base < ncol; base += TopK<CAP>::ROUND) {
        tk.maybe_compact(w);
        const double thr = tk.thr;
        const bool strict = tk.strict != 0;
        for (int i = base + threadIdx.x; i < base + TopK<CAP>::ROUND; i += MMIDX_NT) {
            bool valid = i < ncol;
            double dv = valid ? row[i] : 0.0;
            bool pred = valid && (dv < thr || (dv == thr && !strict));
            tk.push(pred, dv, (unsigned long long)i, i);
        }
    }
    write_result(tk, o, q, 0, w, -1.0);
}

// ------------------------------------------------------------------------------------------------------
// K2: ADC lookup tables.  grid (ceil(npairs/PT), m), thread <-> product centroid.
//   pair g = q*w + p;  l = probes[g];  v[i] = C[l][perm[i]] - Q[q][perm[i]]   (residual = centroid - vector,
//   IVFPQ.java:645; then RandomPermutation.permute RandomPermutation.java:50-56)
//   lut[g][j][c] = sum_t (v[jS+t] - P[j][c][t])^2                              (IVFPQ.java:529-535)
// probes == NULL: flat PQ, v = permuted query (PQ.java:294-301), w == 1.
// ------------------------------------------------------------------------------------------------------
constexpr int LUT_PT = 32;

template <int S>
__global__ void __launch_bounds__(MMIDX_NT) k_lut_build(const double *__restrict__ Q, const double *__restrict__ C,
                                                        const int32_t *__restrict__ probes,
                                                        const int32_t *__restrict__ perm, const double *__restrict__ P,
                                                        int64_t npairs, int w, int d, int m, int ks, int s_rt,
                                                        int64_t lut_stride, double *__restrict__ lut) {
    const int SS = (S > 0) ? S : s_rt;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *vt = reinterpret_cast<double *>(smem_raw);  // [LUT_PT][SS]
    const int j = blockIdx.y;
    const int64_t g0 = (int64_t)blockIdx.x * LUT_PT;
    const int np = (int)min((int64_t)LUT_PT, npairs - g0);
    for (int e = threadIdx.x; e < np * SS; e += MMIDX_NT) {
        int pi = e / SS, t = e - pi * SS;
        int64_t g = g0 + pi;
        int64_t q = g / w;
        int src = j * SS + t;
        if (perm) src = perm[src];
        double qv = Q[q * (int64_t)d + src];
        double val = qv;
        if (probes) {
            int l = probes[g];
            val = __dsub_rn(C[(int64_t)l * d + src], qv);
        }
        vt[pi * SS + t] = val;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < ks; c += MMIDX_NT) {
        const double *pc = P + ((int64_t)j * ks + c) * SS;
        if (S > 0) {
            double pr[S > 0 ? S : 1];
#pragma unroll
            for (int t = 0; t < S; ++t) pr[t] = pc[t];
            for (int pi = 0; pi < np; ++pi) {
                const double *v = vt + pi * S;
                double acc = 0.0;
#pragma unroll
                for (int t = 0; t < S; ++t) acc = sqacc(acc, v[t], pr[t]);
                lut[(g0 + pi) * lut_stride + (int64_t)j * ks + c] = acc;
            }
        } else {
            for (int pi = 0; pi < np; ++pi) {
                const double *v = vt + pi * SS;
                double acc = 0.0;
                for (int t = 0; t < SS; ++t) acc = sqacc(acc, v[t], pc[t]);
                lut[(g0 + pi) * lut_stride + (int64_t)j * ks + c] = acc;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// K3: ADC scan.  dist = sum_{j ascending} lut[j][code_j] starting from 0.0 (IVFPQ.java:435-438).
// One round = ROUND candidates of one list, consumed by all threads, then a block-uniform compaction check.
// ------------------------------------------------------------------------------------------------------
struct CodeLayout {
    int m;           // sub-quantizers
    int ks;          // centroids per sub-quantizer
    int code_bytes;  // m (ks<=256) or 2m
};

template <int CAP, typename PayFn>
__device__ __forceinline__ void scan_segment(TopK<CAP> &tk, const double *__restrict__ lut,
                                             const uint8_t *__restrict__ codes, int64_t len, const CodeLayout L,
                                             unsigned long long seq_hi, int64_t pos0, int k, PayFn pay) {
    constexpr int ROUND = TopK<CAP>::ROUND;
    constexpr int PER = ROUND / MMIDX_NT;  // candidates per thread per round (2 or 4)
    const int tid = threadIdx.x;
    for (int64_t base = 0; base < len; base += ROUND) {
        tk.maybe_compact(k);
        const double thr = tk.thr;
        const bool strict = tk.strict != 0;
        if (L.ks <= 256 && L.m == 8) {
            // 128-bit loads: two 8-byte codes per load
#pragma unroll
            for (int e = 0; e < PER / 2; ++e) {
                int64_t p = base + (int64_t)(e * MMIDX_NT + tid) * 2;
                bool v0 = p < len, v1 = p + 1 < len;
                uint4 c = make_uint4(0, 0, 0, 0);
                if (v0) c = ld_nc_u4(codes + p * 8);
                double d0 = lut[c.x & 255u];
                d0 = __dadd_rn(d0, lut[256 + ((c.x >> 8) & 255u)]);
                d0 = __dadd_rn(d0, lut[512 + ((c.x >> 16) & 255u)]);
                d0 = __dadd_rn(d0, lut[768 + (c.x >> 24)]);
                d0 = __dadd_rn(d0, lut[1024 + (c.y & 255u)]);
                d0 = __dadd_rn(d0, lut[1280 + ((c.y >> 8) & 255u)]);
                d0 = __dadd_rn(d0, lut[1536 + ((c.y >> 16) & 255u)]);
                d0 = __dadd_rn(d0, lut[1792 + (c.y >> 24)]);
                double d1 = lut[c.z & 255u];
                d1 = __dadd_rn(d1, lut[256 + ((c.z >> 8) & 255u)]);
                d1 = __dadd_rn(d1, lut[512 + ((c.z >> 16) & 255u)]);
                d1 = __dadd_rn(d1, lut[768 + (c.z >> 24)]);
                d1 = __dadd_rn(d1, lut[1024 + (c.w & 255u)]);
                d1 = __dadd_rn(d1, lut[1280 + ((c.w >> 8) & 255u)]);
                d1 = __dadd_rn(d1, lut[1536 + ((c.w >> 16) & 255u)]);
                d1 = __dadd_rn(d1, lut[1792 + (c.w >> 24)]);
                bool a0 = v0 && (d0 < thr || (d0 == thr && !strict));
                bool a1 = v1 && (d1 < thr || (d1 == thr && !strict));
                tk.push(a0, d0, seq_hi | (unsigned long long)(pos0 + p), a0 ? pay(pos0 + p) : 0);
                tk.push(a1, d1, seq_hi | (unsigned long long)(pos0 + p + 1), a1 ? pay(pos0 + p + 1) : 0);
            }
        } else if (L.ks <= 256 && L.m == 16) {
#pragma unroll
            for (int e = 0; e < PER; ++e) {
                int64_t p = base + e * MMIDX_NT + tid;
                bool v0 = p < len;
                uint4 c = make_uint4(0, 0, 0, 0);
                if (v0) c = ld_nc_u4(codes + p * 16);
                uint32_t wds[4] = {c.x, c.y, c.z, c.w};
                double d0 = 0.0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    uint32_t code = (wds[j >> 2] >> ((j & 3) * 8)) & 255u;
                    double t = lut[j * 256 + code];
                    d0 = (j == 0) ? t : __dadd_rn(d0, t);
                }
                bool a0 = v0 && (d0 < thr || (d0 == thr && !strict));
                tk.push(a0, d0, seq_hi | (unsigned long long)(pos0 + p), a0 ? pay(pos0 + p) : 0);
            }
        } else {
            // generic layout: byte or short codes, any m
#pragma unroll 1
            for (int e = 0; e < PER; ++e) {
                int64_t p = base + e * MMIDX_NT + tid;
                bool v0 = p < len;
                double d0 = 0.0;
                if (v0) {
                    const uint8_t *cp = codes + p * L.code_bytes;
                    for (int j = 0; j < L.m; ++j) {
                        int code = (L.ks <= 256) ? (int)cp[j] : (int)reinterpret_cast<const uint16_t *>(cp)[j];
                        d0 = __dadd_rn(d0, lut[(int64_t)j * L.ks + code]);
                    }
                }
                bool a0 = v0 && (d0 < thr || (d0 == thr && !strict));
                tk.push(a0, d0, seq_hi | (unsigned long long)(pos0 + p), a0 ? pay(pos0 + p) : 0);
            }
        }
    }
}

struct IvfScanArgs {
    const int32_t *probes;     // [nq][w] in probe order
    const double *luts;        // [nq][w][lut_stride], lut_stride = m*ks rounded up to even (16-byte TMA granules)
    int64_t lut_stride;
    const uint8_t *codes;      // CSR, list starts 16-byte aligned
    const int32_t *iids;       // same positions as codes
    const int64_t *list_off;   // [nlist] start position (in codes) of each list
    const int32_t *list_len;   // [nlist]
    int w, k, nsplit;
    CodeLayout L;
};

// grid (nsplit, nq).  CTA (s,q) scans probes s, s+nsplit, ... of query q in rank order with one collector;
// the LUT of the next probe is fetched by TMA bulk copy while the current list is scanned.
template <int CAP>
__global__ void __launch_bounds__(MMIDX_NT) k_ivfpq_scan(IvfScanArgs a, TopkOut o) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TopK<CAP> &tk = *reinterpret_cast<TopK<CAP> *>(smem_raw);
    const size_t tk_bytes = (sizeof(TopK<CAP>) + 127) & ~(size_t)127;
    const uint32_t lut_bytes = (uint32_t)(a.lut_stride * sizeof(double));
    double *lutbuf0 = reinterpret_cast<double *>(smem_raw + tk_bytes);
    double *lutbuf1 = reinterpret_cast<double *>(smem_raw + tk_bytes + lut_bytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + tk_bytes + 2 * (size_t)lut_bytes);
    const int s = blockIdx.x;
    const int64_t q = blockIdx.y;
    const int32_t *pr = a.probes + q * a.w;
    const double *qlut = a.luts + q * (int64_t)a.w * a.lut_stride;

    tk.init();
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0 && s < a.w) {
        mbar_arrive_expect_tx(&bars[0], lut_bytes);
        tma_load_1d(lutbuf0, qlut + (int64_t)s * a.lut_stride, lut_bytes, &bars[0]);
    }
    int it = 0;
    for (int p = s; p < a.w; p += a.nsplit, ++it) {
        const int cur = it & 1;
        // buffer cur^1 was last read during iteration it-1; every thread has passed the barrier that ends it
        if (threadIdx.x == 0 && p + a.nsplit < a.w) {
            fence_proxy_async();
            mbar_arrive_expect_tx(&bars[cur ^ 1], lut_bytes);
            tma_load_1d(cur ? lutbuf0 : lutbuf1, qlut + (int64_t)(p + a.nsplit) * a.lut_stride, lut_bytes,
                        &bars[cur ^ 1]);
        }
        mbar_wait(&bars[cur], (uint32_t)((it >> 1) & 1));
        const double *lut = cur ? lutbuf1 : lutbuf0;
        const int l = pr[p];
        const int64_t start = a.list_off[l];
        const int64_t len = a.list_len[l];
        const int32_t *li = a.iids + start;
        scan_segment(tk, lut, a.codes + start * a.L.code_bytes, len, a.L, ((unsigned long long)p) << 32, 0, a.k,
                     [li](int64_t pos) { return li[pos]; });
        __syncthreads();  // all reads of `lut` done before it is refilled two iterations later
    }
    write_result(tk, o, q, s, a.k, -1.0);
}

struct PqScanArgs {
    const double *luts;    // [nq][lut_stride]
    int64_t lut_stride;
    const uint8_t *codes;  // [n][code_bytes], iid == position
    int64_t n;
    int64_t chunk;  // positions per split (multiple of the round size)
    int k;
    CodeLayout L;
};

// grid (nsplit, nq): CTA (s,q) scans positions [s*chunk, (s+1)*chunk).
template <int CAP>
__global__ void __launch_bounds__(MMIDX_NT) k_pq_scan(PqScanArgs a, TopkOut o) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TopK<CAP> &tk = *reinterpret_cast<TopK<CAP> *>(smem_raw);
    const size_t tk_bytes = (sizeof(TopK<CAP>) + 127) & ~(size_t)127;
    const uint32_t lut_bytes = (uint32_t)(a.lut_stride * sizeof(double));
    double *lut = reinterpret_cast<double *>(smem_raw + tk_bytes);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + tk_bytes + lut_bytes);
    const int s = blockIdx.x;
    const int64_t q = blockIdx.y;
    tk.init();
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, lut_bytes);
        tma_load_1d(lut, a.luts + q * a.lut_stride, lut_bytes, bar);
    }
    mbar_wait(bar, 0);
    const int64_t p0 = (int64_t)s * a.chunk;
    const int64_t len = max((int64_t)0, min(a.chunk, a.n - p0));
    scan_segment(tk, lut, a.codes + p0 * a.L.code_bytes, len, a.L, 0ull, p0, a.k,
                 [](int64_t pos) { return (int)pos; });
    write_result(tk, o, q, s, a.k, -1.0);
}

// ------------------------------------------------------------------------------------------------------
// K4: merge nparts partial results per query.  Input layout: element (part, q, i) at
// ((part*part_stride_q + q*q_stride) + i) so that both [nq][nparts][k] scratch and an all-gathered
// [nparts][nq][k] buffer can be merged.  grid nq.
// ------------------------------------------------------------------------------------------------------
struct MergeArgs {
    const int32_t *iids;
    const double *dist;
    const unsigned long long *seq;
    const int32_t *cnt;
    const double *tie;  // may be NULL
    int64_t part_stride, q_stride;          // in units of k-rows: row(part,q) = part*part_stride + q*q_stride
    int nparts, k;
};

template <int CAP>
__global__ void __launch_bounds__(MMIDX_NT) k_merge_topk(MergeArgs a, TopkOut o) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TopK<CAP> &tk = *reinterpret_cast<TopK<CAP> *>(smem_raw);
    constexpr int ROUND = TopK<CAP>::ROUND;
    const int64_t q = blockIdx.x;
    tk.init();
    const int total = a.nparts * a.k;
    for (int base = 0; base < total; base += ROUND) {
        tk.maybe_compact(a.k);
        const double thr = tk.thr;
        const bool strict = tk.strict != 0;
        for (int e = base + threadIdx.x; e < base + ROUND; e += MMIDX_NT) {
            bool valid = e < total;
            int part = valid ? e / a.k : 0;
            int i = e - part * a.k;
            int64_t row = (int64_t)part * a.part_stride + q * a.q_stride;
            valid = valid && i < a.cnt[row];
            double dv = 0.0;
            unsigned long long sv = 0ull;
            int pv = 0;
            if (valid) {
                dv = a.dist[row * a.k + i];
                sv = a.seq[row * a.k + i];
                pv = a.iids[row * a.k + i];
            }
            bool pred = valid && (dv < thr || (dv == thr && !strict));
            tk.push(pred, dv, sv, pv);
        }
    }
    // All partial queues fit the collector uncompacted (total <= CAP: no round above compacted).  If an exact tie is cut at
    // the k-th boundary and NO part discarded candidates at that distance, the union of the parts holds every candidate with
    // dist <= T, so the queue's tie rule can be replayed right here (tie_resolve.cuh) -- the ordered tie pass, which
    // re-evaluates whole lists, is left for the case where a part's own cut hides candidates.
    __shared__ int s_amb, s_part_cut;
    int *flag = reinterpret_cast<int *>(smem_raw + ((sizeof(TopK<CAP>) + 127) & ~(size_t)127));  // [CAP] (host sizes it)
    __syncthreads();
    const int nall = tk.cnt;
    if (total <= CAP && nall > a.k) {
        tk.sort_first(nall);
        const bool cut = tk.dist[a.k - 1] == tk.dist[a.k];  // block-uniform
        if (cut) {
            if (threadIdx.x == 0) {
                int pc = 0;
                if (a.tie)
                    for (int part = 0; part < a.nparts; ++part)
                        if (a.tie[(int64_t)part * a.part_stride + q * a.q_stride] == tk.dist[a.k - 1]) pc = 1;
                s_part_cut = pc;
            }
            __syncthreads();
            if (s_part_cut == 0) tk.kill_tie_losers(nall, a.k, flag);  // losers get dist = +inf and drop out below
        }
    }
    bool amb;
    const int n = tk.finalize(a.k, &amb);
    // a part that discarded candidates tied at distance t makes the answer ambiguous iff t == final T
    // (T <= every part's own k-th distance, so comparing each part's value with T is exact)
    if (threadIdx.x == 0) {
        if (a.tie && n == a.k) {
            for (int part = 0; part < a.nparts; ++part)
                if (a.tie[(int64_t)part * a.part_stride + q * a.q_stride] == tk.dist[n - 1]) amb = true;
        }
        s_amb = amb ? 1 : 0;
    }
    __syncthreads();
    write_sorted(tk, o, q, 0, a.k, n, s_amb != 0);  // o.nparts == 1
}

// ------------------------------------------------------------------------------------------------------
// K6: PQ encode.  grid (ceil(n/256), m); thread <-> vector, sub-quantizer j's centroids staged in shared
// memory in chunks and read as warp-wide broadcasts.  code = argmin_c sum_t (P[j][c][t] - v[jS+t])^2 with
// the lowest index among equal minima (PQ.java:423).  list != NULL: v = permuted (C[list] - x) (IVFPQ.java:316-323).
// ------------------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(MMIDX_NT) k_pq_encode(const double *__restrict__ X, const double *__restrict__ C,
                                                        const int32_t *__restrict__ list,
                                                        const int32_t *__restrict__ perm, const double *__restrict__ P,
                                                        int64_t n, int d, int m, int ks, int cchunk,
                                                        uint8_t *__restrict__ out, int code_bytes) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *pc = reinterpret_cast<double *>(smem_raw);  // [cchunk][S]
    const int j = blockIdx.y;
    const int64_t v = (int64_t)blockIdx.x * MMIDX_NT + threadIdx.x;
    const bool valid = v < n;
    double x[S];
#pragma unroll
    for (int t = 0; t < S; ++t) {
        double val = 0.0;
        if (valid) {
            int src = j * S + t;
            if (perm) src = perm[src];
            val = X[v * (int64_t)d + src];
            if (list) val = __dsub_rn(C[(int64_t)list[v] * d + src], val);
        }
        x[t] = val;
    }
    double best = 1.7976931348623157e308;  // Double.MAX_VALUE
    int bidx = -1;
    for (int c0 = 0; c0 < ks; c0 += cchunk) {
        int cn = min(cchunk, ks - c0);
        __syncthreads();
        const double *src = P + ((int64_t)j * ks + c0) * S;
        for (int e = threadIdx.x; e < cn * S; e += MMIDX_NT) pc[e] = src[e];
        __syncthreads();
        for (int c = 0; c < cn; ++c) {
            const double *cc = pc + c * S;
            double acc = 0.0;
#pragma unroll
            for (int t = 0; t < S; ++t) acc = sqacc(acc, cc[t], x[t]);
            if (acc < best) {
                best = acc;
                bidx = c0 + c;
            }
        }
    }
    if (valid) {
        if (ks <= 256)
            out[v * code_bytes + j] = (uint8_t)bidx;
        else
            reinterpret_cast<uint16_t *>(out + v * code_bytes)[j] = (uint16_t)bidx;
    }
}

// generic sub-vector length (S not instantiated): sub-vector re-read from global (L1-resident)
__global__ void __launch_bounds__(MMIDX_NT) k_pq_encode_generic(const double *__restrict__ X, const double *__restrict__ C,
                                                                const int32_t *__restrict__ list,
                                                                const int32_t *__restrict__ perm,
                                                                const double *__restrict__ P, int64_t n, int d, int m,
                                                                int ks, int S, uint8_t *__restrict__ out,
                                                                int code_bytes) {
    const int j = blockIdx.y;
    const int64_t v = (int64_t)blockIdx.x * MMIDX_NT + threadIdx.x;
    if (v >= n) return;
    double best = 1.7976931348623157e308;
    int bidx = -1;
    for (int c = 0; c < ks; ++c) {
        const double *cc = P + ((int64_t)j * ks + c) * S;
        double acc = 0.0;
        for (int t = 0; t < S; ++t) {
            int src = j * S + t;
            if (perm) src = perm[src];
            double val = X[v * (int64_t)d + src];
            if (list) val = __dsub_rn(C[(int64_t)list[v] * d + src], val);
            acc = sqacc(acc, cc[t], val);
        }
        if (acc < best) {
            best = acc;
            bidx = c;
        }
    }
    if (ks <= 256)
        out[v * code_bytes + j] = (uint8_t)bidx;
    else
        reinterpret_cast<uint16_t *>(out + v * code_bytes)[j] = (uint16_t)bidx;
}

// ------------------------------------------------------------------------------------------------------
// Index maintenance (not arithmetic): CSR scatter of the append log, Linear block packing.
// ------------------------------------------------------------------------------------------------------
__global__ void k_scatter_codes(const uint8_t *__restrict__ log_codes, const int32_t *__restrict__ log_iid,
                                const int64_t *__restrict__ dst, int64_t n, int code_bytes,
                                uint8_t *__restrict__ codes, int32_t *__restrict__ iids) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t t = dst[i];
    for (int b = 0; b < code_bytes; ++b) codes[t * code_bytes + b] = log_codes[i * code_bytes + b];
    iids[t] = log_iid[i];
}

__global__ void k_gather_rows(const uint8_t *__restrict__ src, const int64_t *__restrict__ idx, int64_t n,
                              int row_bytes, uint8_t *__restrict__ dstp) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t s = idx[i];
    for (int b = 0; b < row_bytes; ++b) dstp[i * row_bytes + b] = src[s * row_bytes + b];
}

// Linear database layout: blocks of 32 vectors, dimension-major inside a block:
//   Xb[(i/32)*d*32 + j*32 + (i%32)] = X[i][j]   -> a warp reads 256 contiguous bytes per dimension.
__global__ void k_linear_pack(const double *__restrict__ X, int64_t n, int d, int64_t i0, double *__restrict__ Xb) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * d) return;
    int64_t i = e / d;
    int j = (int)(e - i * d);
    int64_t gi = i0 + i;
    Xb[(gi >> 5) * (int64_t)d * 32 + (int64_t)j * 32 + (gi & 31)] = X[e];
}

__global__ void k_linear_unpack_row(const double *__restrict__ Xb, int d, int64_t gi, double *__restrict__ out) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < d) out[j] = Xb[(gi >> 5) * (int64_t)d * 32 + (int64_t)j * 32 + (gi & 31)];
}

// ------------------------------------------------------------------------------------------------------
// K8: Linear exact kNN. grid (nsplit, nq): CTA (s,q) scans vectors [s*chunk, (s+1)*chunk).
// dist = sum_j (q[j] - x[j])^2, j ascending (Linear.java:147-149).
// ------------------------------------------------------------------------------------------------------
struct LinearArgs {
    const double *Q;   // [nq][d]
    const double *Xb;  // packed
    int64_t n, chunk;
    int d, k;
};

template <int CAP>
__global__ void __launch_bounds__(MMIDX_NT) k_linear_scan(LinearArgs a, TopkOut o) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TopK<CAP> &tk = *reinterpret_cast<TopK<CAP> *>(smem_raw);
    const size_t tk_bytes = (sizeof(TopK<CAP>) + 127) & ~(size_t)127;
    double *qs = reinterpret_cast<double *>(smem_raw + tk_bytes);  // [d]
    constexpr int ROUND = TopK<CAP>::ROUND;
    const int s = blockIdx.x;
    const int64_t q = blockIdx.y;
    tk.init();
    for (int j = threadIdx.x; j < a.d; j += MMIDX_NT) qs[j] = a.Q[q * (int64_t)a.d + j];
    __syncthreads();
    const int64_t p0 = (int64_t)s * a.chunk;
    const int64_t pend = min(a.n, p0 + a.chunk);
    for (int64_t base = p0; base < pend; base += ROUND) {
        tk.maybe_compact(a.k);
        const double thr = tk.thr;
        const bool strict = tk.strict != 0;
#pragma unroll 1
        for (int e = 0; e < ROUND / MMIDX_NT; ++e) {
            int64_t i = base + e * MMIDX_NT + threadIdx.x;
            bool valid = i < pend;
            double acc = 0.0;
            if (valid) {
                const double *xp = a.Xb + (i >> 5) * (int64_t)a.d * 32 + (i & 31);
#pragma unroll 8
                for (int j = 0; j < a.d; ++j) acc = sqacc(acc, qs[j], xp[(int64_t)j * 32]);
            }
            bool pred = valid && (acc < thr || (acc == thr && !strict));
            tk.push(pred, acc, (unsigned long long)i, (int)i);
        }
    }
    write_result(tk, o, q, s, a.k, -1.0);
}

// ------------------------------------------------------------------------------------------------------
// K7: VLAD.  Assignment = k_assign_nearest over all descriptors.  Accumulation must follow descriptor order
// per centroid (binary64 addition is not associative): k_vlad_order builds, per image, the stable
// by-centroid ordering of its descriptors; k_vlad_accumulate sums `desc[i] - codebook[nn][i]` in that order.
// grid n_img for both.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MMIDX_NT) k_vlad_order(const int32_t *__restrict__ assign,
                                                         const int64_t *__restrict__ offsets, int K,
                                                         int32_t *__restrict__ order, int32_t *__restrict__ cstart) {
    // cstart: [n_img][K+1] start of each centroid's run inside order[offsets[img] ..]
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int *cnt = reinterpret_cast<int *>(smem_raw);  // [K+1]
    const int64_t img = blockIdx.x;
    const int64_t o0 = offsets[img];
    const int n = (int)(offsets[img + 1] - o0);
    for (int c = threadIdx.x; c <= K; c += MMIDX_NT) cnt[c] = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += MMIDX_NT) atomicAdd(&cnt[assign[o0 + t]], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int c = 0; c < K; ++c) {
            int v = cnt[c];
            cnt[c] = run;
            run += v;
        }
        cnt[K] = run;
    }
    __syncthreads();
    for (int c = threadIdx.x; c <= K; c += MMIDX_NT) cstart[img * (int64_t)(K + 1) + c] = cnt[c];
    // stable fill: thread c walks the image's assignments in descriptor order
    for (int c = threadIdx.x; c < K; c += MMIDX_NT) {
        int w = cnt[c];
        int end = cnt[c + 1];
        for (int t = 0; t < n && w < end; ++t)
            if (assign[o0 + t] == c) order[o0 + w++] = t;
    }
}

__global__ void __launch_bounds__(MMIDX_NT) k_vlad_accumulate(const double *__restrict__ codebook,
                                                              const double *__restrict__ desc,
                                                              const int64_t *__restrict__ offsets,
                                                              const int32_t *__restrict__ order,
                                                              const int32_t *__restrict__ cstart, int K, int D,
                                                              double *__restrict__ out, int64_t ld) {
    // out rows are `ld` apart: K*D for a single vocabulary, the multi-VLAD length when several vocabularies are concatenated
    const int64_t img = blockIdx.x;
    const int64_t o0 = offsets[img];
    const int32_t *cs = cstart + img * (int64_t)(K + 1);
    double *vo = out + img * ld;
    for (int e = threadIdx.x; e < K * D; e += MMIDX_NT) {
        int c = e / D, i = e - c * D;
        double cb = codebook[(int64_t)c * D + i];
        double acc = 0.0;  // `new double[...]` is zero-filled, VladAggregator.java:57
        for (int r = cs[c]; r < cs[c + 1]; ++r) {
            int t = order[o0 + r];
            acc = __dadd_rn(acc, __dsub_rn(desc[(o0 + t) * (int64_t)D + i], cb));
        }
        vo[e] = acc;
    }
}

}  // namespace mmidx
