// coarse_fast.cuh -- coarse stage (IVFPQ.computeNearestCoarseIndices, IVFPQ.java:575-601) as fp32 filter + exact
// binary64 verification.  Returns the SAME probe lists, in the same order, as the exact kernels (k_sqdist_matrix +
// k_select_rows): a centroid is only ever ranked on its exact binary64 distance, computed in the reference's order.
//
//   a[q][c] = fp32(||C_c||^2) - 2 * dot32(fp32(q), fp32(C_c))          (k_coarse_f32: tiled FFMA GEMM)
// differs from t = ||C_c||^2 - 2 q.C_c (the exact distance minus ||q||^2, which does not change the ranking) by at most
//   B(q) = 1.02 u (2 Cmax^2 + (2d + 8) ||q|| Cmax) + (d + 2) 2^-52 (||q|| + Cmax)^2,   u = 2^-24,  Cmax = max_c ||C_c||
// (rounding of the inputs and of ||C||^2: u each; fp32 dot product of d terms in any order: gamma_d; the final fma;
//  Cauchy-Schwarz for sum |q_j C_j|; the last term covers the rounding of the binary64 reference sum itself).
// If x is the w-th smallest a[q][.], the w-th smallest exact distance T satisfies T - ||q||^2 <= x + B, and every
// centroid with exact distance <= T has a <= x + 2B.  k_coarse_verify keeps exactly those (a few more than w),
// evaluates them in binary64 (sum_j (C[c][j] - q[j])^2, j ascending, IVFPQ.java:583), and selects the top w with the
// queue's order and tie rule -- the survivors contain every centroid at distance <= T, so the rule is replayed locally.
// A query whose band holds more survivors than the collector can keep is ranked by the exact sweep in the same kernel.
#pragma once
#include "common.cuh"
#include "kernels.cuh"
#include "tie_resolve.cuh"
#include "topk.cuh"

namespace mmidx {

// C32[c][j] = fp32(C[c][j]); c2[c] = fp32(||C_c||^2); cmax[0] >= max_c ||C_c|| (zero-filled before launch). grid nlist.
__global__ void __launch_bounds__(MMIDX_NT) k_coarse_tables(const double *__restrict__ C, int d, float *__restrict__ C32,
                                                            float *__restrict__ c2, float *__restrict__ cmax) {
    __shared__ double red[MMIDX_NT / 32];
    const int c = blockIdx.x;
    double n2 = 0.0;
    for (int j = threadIdx.x; j < d; j += MMIDX_NT) {
        const double v = C[(int64_t)c * d + j];
        C32[(int64_t)c * d + j] = __double2float_rn(v);
        n2 += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = n2;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < MMIDX_NT / 32; ++i) s += red[i];
        c2[c] = __double2float_rn(s);
        const float nrm = __double2float_ru(sqrt(s) * (1.0 + 1e-12));
        atomicMax(reinterpret_cast<int *>(cmax), __float_as_int(nrm));  // non-negative floats order like their bits
    }
}

// A32[q][c] = c2[c] - 2 * sum_j fp32(Q[q][j]) * C32[c][j].  64 x 128 output tile per CTA, 8 x 4 per thread, K step 16.
constexpr int CG_BM = 64, CG_BN = 128, CG_BK = 16;

__global__ void __launch_bounds__(MMIDX_NT) k_coarse_f32(const double *__restrict__ Q, const float *__restrict__ C32,
                                                         const float *__restrict__ c2, int64_t nq, int nlist, int d,
                                                         float *__restrict__ A32) {
    __shared__ float As[CG_BK][CG_BM + 4];
    __shared__ float Bs[CG_BK][CG_BN + 4];
    const int tid = threadIdx.x;
    const int64_t q0 = (int64_t)blockIdx.y * CG_BM;
    const int n0 = blockIdx.x * CG_BN;
    const int tr = tid >> 5, tc = tid & 31;  // thread <-> rows tr*8.., columns tc + 32*i
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < d; k0 += CG_BK) {
        __syncthreads();
        for (int e = tid; e < CG_BM * CG_BK; e += MMIDX_NT) {
            const int r = e / CG_BK, kk = e - r * CG_BK;
            const int64_t q = q0 + r;
            As[kk][r] = (q < nq && k0 + kk < d) ? __double2float_rn(Q[q * (int64_t)d + k0 + kk]) : 0.f;
        }
        for (int e = tid; e < CG_BN * CG_BK; e += MMIDX_NT) {
            const int n = e / CG_BK, kk = e - n * CG_BK;
            Bs[kk][n] = (n0 + n < nlist && k0 + kk < d) ? C32[(int64_t)(n0 + n) * d + k0 + kk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < CG_BK; ++kk) {
            float av[8], bv[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) av[i] = As[kk][tr * 8 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tc + 32 * j];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t q = q0 + tr * 8 + i;
        if (q >= nq) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tc + 32 * j;
            if (n < nlist) A32[q * (int64_t)nlist + n] = fmaf(-2.f, acc[i][j], c2[n]);
        }
    }
}

// grid nq.  Shared memory: TopK<CAP> | qv [d] | keys [nlist] fp32 | surv [CAP] | xs [vb][d + 1]

template <int CAP>
__global__ void __launch_bounds__(MMIDX_NT) k_coarse_verify(const double *__restrict__ Q, const double *__restrict__ C,
                                                            const float *__restrict__ A32, const float *__restrict__ cmax,
                                                            int nlist, int d, int w, int vb, TopkOut o) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TopK<CAP> &tk = *reinterpret_cast<TopK<CAP> *>(smem_raw);
    const size_t tk_bytes = (sizeof(TopK<CAP>) + 127) & ~(size_t)127;
    double *qv = reinterpret_cast<double *>(smem_raw + tk_bytes);  // [d]
    float *key = reinterpret_cast<float *>(qv + d);                 // [nlist]
    int *surv = reinterpret_cast<int *>(key + ((nlist + 1) & ~1));             // [CAP] centroid ids, later scratch of the tie rule
    double *xs = reinterpret_cast<double *>(surv + CAP);            // [vb][d + 1] squared terms of a batch of survivors
    __shared__ unsigned int hist[256];
    __shared__ int s_bin, s_krem, s_ns;
    __shared__ double s_q2;
    const int tid = threadIdx.x;
    const int64_t q = blockIdx.x;
    const float *arow = A32 + q * (int64_t)nlist;
    for (int i = tid; i < d; i += MMIDX_NT) qv[i] = Q[q * (int64_t)d + i];
    for (int i = tid; i < nlist; i += MMIDX_NT) key[i] = arow[i];
    if (tid == 0) s_ns = 0;
    tk.init();  // barrier
    if (tid < 32) {
        double n2 = 0.0;
        for (int i = tid; i < d; i += 32) n2 += qv[i] * qv[i];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, off);
        if (tid == 0) s_q2 = n2;
    }
    // ---- an upper bound x of the w-th smallest fp32 key: 2 radix passes of 8 bits over the order-preserving unsigned
    //      image locate its 16-bit prefix; x = the largest value with that prefix (any x >= the w-th key keeps the
    //      survivor set complete, a tighter x only keeps it smaller) ----
    unsigned prefix = 0, krem = (unsigned)w;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int shift = 24 - 8 * pass;
        hist[tid] = 0;
        __syncthreads();
        for (int i0 = 0; i0 < nlist; i0 += MMIDX_NT) {
            const int i = i0 + tid;
            bool act = i < nlist;
            unsigned bin = 0;
            if (act) {
                unsigned u = __float_as_uint(key[i]);
                u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
                act = (pass == 0 || (u >> (shift + 8)) == prefix);
                bin = (u >> shift) & 255u;
            }
            hist_add(hist, act, bin);
        }
        __syncthreads();
        if (tid < 32) {
            unsigned loc[8], sum = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                loc[b] = hist[tid * 8 + b];
                sum += loc[b];
            }
            unsigned incl = sum;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                unsigned t = __shfl_up_sync(0xffffffffu, incl, off);
                if (tid >= off) incl += t;
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
    }
    const unsigned edge = (prefix << 16) | 0xffffu;  // largest unsigned image with the selected 16-bit prefix
    float kth = __uint_as_float((edge & 0x80000000u) ? (edge & 0x7fffffffu) : ~edge);
    if (!(kth <= 3.4028234663852886e38f)) kth = 3.4028234663852886e38f;  // prefix of inf / nan patterns
    // ---- survivors: a <= kth + 2B ----
    const double qn = sqrt(s_q2) * (1.0 + 1e-12), cm = (double)cmax[0];
    const double B = 1.02 * 5.9604644775390625e-08 * (2.0 * cm * cm + (2.0 * d + 8.0) * qn * cm) +
                     (d + 2.0) * 2.220446049250313e-16 * (qn + cm) * (qn + cm);
    const float lim = __double2float_ru((double)kth + 2.0 * B + 1e-30);
    for (int i0 = 0; i0 < nlist; i0 += MMIDX_NT) {
        const int i = i0 + tid;
        const bool pred = i < nlist && key[i] <= lim;
        const unsigned mask = __ballot_sync(0xffffffffu, pred);
        if (mask) {
            const int lane = tid & 31, leader = __ffs(mask) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&s_ns, __popc(mask));
            base = __shfl_sync(0xffffffffu, base, leader);
            const int slot = base + __popc(mask & ((1u << lane) - 1u));
            if (pred && slot < CAP) surv[slot] = i;
        }
    }
    __syncthreads();
    // outside the fp32-safe magnitude window (fast_scan.cuh) the filter proves nothing: rank by the exact sweep below
    const bool in_range = fast_mag_ok(qn) && fast_mag_ok(cm) && B == B && B < 1e300;
    const int ns = in_range ? s_ns : CAP + 1;
    if (ns <= CAP) {
        // exact binary64 distance of every survivor, `vb` at a time: the block squares the terms with coalesced loads of
        // the centroid rows, then one thread per survivor adds them for j ascending -- the reference's loop
        // (IVFPQ.java:582-585), one rounding per operation
        for (int b0 = 0; b0 < ns; b0 += vb) {
            const int nb = min(vb, ns - b0);
            for (int e = tid >> 5; e < nb; e += MMIDX_NT / 32) {  // warp <-> centroid row
                const double *cr = C + (int64_t)surv[b0 + e] * d;
                double *xr = xs + e * (d + 1);
                for (int j = tid & 31; j < d; j += 32) {
                    const double df = __dsub_rn(cr[j], qv[j]);
                    xr[j] = __dmul_rn(df, df);
                }
            }
            __syncthreads();
            if (tid < nb) {
                const double *xr = xs + tid * (d + 1);
                double acc = 0.0;
#pragma unroll 16
                for (int j = 0; j < d; ++j) acc = __dadd_rn(acc, xr[j]);
                const int c = surv[b0 + tid];
                tk.dist[b0 + tid] = acc;
                tk.seq[b0 + tid] = (unsigned long long)c;  // offer order = centroid index (IVFPQ.java:579)
                tk.pay[b0 + tid] = c;
            }
            __syncthreads();
        }
        if (tid == 0) tk.cnt = ns;
        __syncthreads();
        tk.sort_first(ns);
        if (ns > w && tk.dist[w - 1] == tk.dist[w]) {  // block-uniform: an exact tie is cut at the w-th boundary
            tk.kill_tie_losers(ns, w, surv);
            write_result(tk, o, q, 0, w, -1.0);
        } else {
            write_sorted(tk, o, q, 0, w, min(ns, w), false);
        }
        return;
    }
    // ---- the band is wider than the collector (massive near-ties): exact sweep over all centroids ----
    for (int base = 0; base < nlist; base += TopK<CAP>::ROUND) {
        tk.maybe_compact(w);
        const double thr = tk.thr;
        const bool strict = tk.strict != 0;
        for (int i = base + tid; i < base + TopK<CAP>::ROUND; i += MMIDX_NT) {
            const bool valid = i < nlist;
            double dv = 0.0;
            if (valid) {
                const double *cr = C + (int64_t)i * d;
                for (int j = 0; j < d; ++j) dv = sqacc(dv, cr[j], qv[j]);
            }
            const bool pred = valid && (dv < thr || (dv == thr && !strict));
            tk.push(pred, dv, (unsigned long long)i, i);
        }
    }
    write_result(tk, o, q, 0, w, -1.0);
}

// ordered tie pass without a distance matrix (rows the exact sweep above flagged): distances straight from C
__global__ void __launch_bounds__(MMIDX_NT) k_tie_collect_rows_direct(const double *__restrict__ Q, const double *__restrict__ C,
                                                                      int ncol, int d, int k,
                                                                      const double *__restrict__ res_dist,
                                                                      const int32_t *__restrict__ amb_list,
                                                                      const int32_t *__restrict__ amb_count, TieLists o) {
    __shared__ int warp_sums[MMIDX_NT / 32];
    const int na = *amb_count;
    for (int a = blockIdx.x; a < na; a += gridDim.x) {
        const int64_t q = amb_list[a];
        const double T = res_dist[q * k + k - 1];
        const double *qv = Q + q * (int64_t)d;
        int found = 0;
        tie_sweep(ncol, 0ull, T, k, found, warp_sums, o, q,
                  [=](int64_t i) {
                      const double *cr = C + i * d;
                      double dv = 0.0;
                      for (int j = 0; j < d; ++j) dv = sqacc(dv, cr[j], qv[j]);
                      return dv;
                  },
                  [](int64_t i) { return (int)i; });
        if (threadIdx.x == 0) o.cnt[q] = min(found, k);
        __syncthreads();
    }
}

}  // namespace mmidx
