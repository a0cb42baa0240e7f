// argmin_filter.cuh -- exact nearest-centroid assignment with an fp32 filter, for codebooks small enough to tile:
//   AbstractFeatureAggregator.computeNearestCentroid   AFA.java:136-155     (VLAD: K x D codebook)
//   PQ / IVFPQ.computeNearestProductIndex              PQ.java:411-429, IVFPQ.java:613-631   (per sub-quantizer: ks x S)
// idx = argmin_c sum_t (B[c][t] - x[t])^2 in binary64, t ascending, lowest index among equal minima (`distance < minDistance`).
//
// Brute force in binary64 is 3 K D dependent-ish operations per vector on the (slow) fp64 pipe.  Here
//   a[c] = fp32(||B_c||^2) - 2 dot32(fp32(x), fp32(B_c))            (register-tiled FFMA, like an SGEMM)
// differs from t[c] = ||B_c||^2 - 2 x.B_c (the exact distance minus ||x||^2, which does not change the argmin) by at most the
// radius R(x) of coarse_fast.cuh (same expression with D, ||x||, Bmax).  With a1 = min_c a[c], the exact argmin c* satisfies
// a[c*] <= a1 + 2R, so if the SECOND smallest a is above a1 + 2R the smallest one is the exact argmin -- no binary64 work
// at all for that vector.  Otherwise (near ties, exact ties, or magnitudes outside the fp32-safe window) the vector is put
// on a list and k_argmin_exact_list evaluates it in binary64 in the reference's order.  Same indices as the exact kernels,
// bit for bit (tests: codes / assignments against the oracle, duplicated centroids, scaled data).
#pragma once
#include "common.cuh"

namespace mmidx {

// how row v's D-vector is formed: x[t] = X[v * ldx + src] or, with a residual, C[list[v] * ldc + src] - X[v * ldx + src]
// (IVFPQ.java:316 computeResidualVector: centroid - vector), src = perm ? perm[col0 + t] : col0 + t
struct ArgminRows {
    const double *X;
    int64_t ldx;
    const double *C;      // or NULL
    const int32_t *list;  // [n] row of C per vector (with C)
    int64_t ldc;
    const int32_t *perm;  // or NULL
    int col0;
};

__device__ __forceinline__ double argmin_row_value(const ArgminRows &r, int64_t v, int t) {
    int src = r.col0 + t;
    if (r.perm) src = r.perm[src];
    const double x = r.X[v * r.ldx + src];
    return r.C ? __dsub_rn(r.C[(int64_t)r.list[v] * r.ldc + src], x) : x;
}

// fp32 tables of a codebook B[nb][K][D] (nb sub-quantizers): B32 = fp32(B), b2 = fp32(||B_c||^2), bmax[b] >= max_c ||B_c||
// (zero-filled before launch).  grid (K, nb).
__global__ void __launch_bounds__(MMIDX_NT) k_argmin_tables(const double *__restrict__ B, int K, int D, float *__restrict__ B32,
                                                            float *__restrict__ b2, float *__restrict__ bmax) {
    __shared__ double red[MMIDX_NT / 32];
    const int c = blockIdx.x, b = blockIdx.y;
    const double *row = B + ((int64_t)b * K + c) * D;
    float *o = B32 + ((int64_t)b * K + c) * D;
    double n2 = 0.0;
    for (int t = threadIdx.x; t < D; t += MMIDX_NT) {
        const double v = row[t];
        o[t] = __double2float_rn(v);
        n2 += v * v;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = n2;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < MMIDX_NT / 32; ++i) s += red[i];
        b2[(int64_t)b * K + c] = __double2float_rn(s);
        atomicMax(reinterpret_cast<int *>(&bmax[b]), __float_as_int(__double2float_ru(sqrt(s) * (1.0 + 1e-12))));
    }
}

// CTA tile: AF_TV vectors x AF_TK centroids per pass, 16 x 16 threads, 4 vectors x 8 centroids per thread, D step 16.
// Thread (ty, tx) owns vectors ty*4 .. ty*4+3 and centroids tx*4 .. tx*4+3 and 64 + tx*4 .. 64 + tx*4+3, so every operand of a
// k-step comes from three 128-bit shared loads (one broadcast, two with 16 lanes x 16 B contiguous).
constexpr int AF_TV = 64, AF_TK = 128, AF_DK = 16;
__device__ __forceinline__ int af_centroid(int tx, int j) { return j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4); }

struct Min2 {
    float a1, a2;  // smallest and second smallest filter value seen
    int i1;        // index of the smallest
};
__device__ __forceinline__ void min2_add(Min2 &m, float a, int i) {
    if (a < m.a1) {
        m.a2 = m.a1;
        m.a1 = a;
        m.i1 = i;
    } else if (a < m.a2) {
        m.a2 = a;
    }
}
__device__ __forceinline__ void min2_merge(Min2 &m, float a1, float a2, int i1) {
    if (a1 < m.a1) {
        m.a2 = fminf(m.a1, a2);
        m.a1 = a1;
        m.i1 = i1;
    } else {
        m.a2 = fminf(m.a2, a1);  // a1 >= m.a1: it is a runner-up (a tie counts as ambiguous)
    }
}

// grid (ceil(n / AF_TV), nb).  Block (b = blockIdx.y) uses codebook b: B32[b][K][D], b2[b][K], bmax[b], rows.col0 + b * D.
// out[v * out_stride + b] receives the index when the filter decides; otherwise v * nb + b is appended to amb_list.
// out16: write uint16 codes (ks > 256), else out8 (PQ) or out32 (VLAD / generic, out_stride in elements).
__global__ void __launch_bounds__(MMIDX_NT) k_argmin_filter(ArgminRows rows, const float *__restrict__ B32, const float *__restrict__ b2,
                                                            const float *__restrict__ bmax, int64_t n, int K, int D, int nb,
                                                            uint8_t *__restrict__ out8, uint16_t *__restrict__ out16,
                                                            int32_t *__restrict__ out32, int64_t out_stride,
                                                            int64_t *__restrict__ amb_list, int32_t *__restrict__ amb_count) {
    __shared__ __align__(16) float Xs[AF_DK][AF_TV + 4];
    __shared__ __align__(16) float Bs[AF_DK][AF_TK + 4];
    __shared__ double xn2[AF_TV];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int b = blockIdx.y;
    const int64_t v0 = (int64_t)blockIdx.x * AF_TV;
    ArgminRows r = rows;
    r.col0 = rows.col0 + b * D;
    const float *Bb = B32 + (int64_t)b * K * D;
    const float *b2b = b2 + (int64_t)b * K;
    // ||x||^2 of the tile's vectors in binary64 (4 threads per vector)
    {
        const int vi = tid >> 2, part = tid & 3;
        const int64_t v = v0 + vi;
        double s = 0.0;
        if (v < n)
            for (int t = part; t < D; t += 4) {
                const double x = argmin_row_value(r, v, t);
                s += x * x;
            }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (part == 0) xn2[vi] = s;
    }
    Min2 best[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        best[i].a1 = best[i].a2 = __int_as_float(0x7f800000);
        best[i].i1 = 0;
    }
    for (int k0 = 0; k0 < K; k0 += AF_TK) {
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        for (int d0 = 0; d0 < D; d0 += AF_DK) {
            __syncthreads();
            for (int e = tid; e < AF_TV * AF_DK; e += MMIDX_NT) {
                const int vi = e / AF_DK, t = e - vi * AF_DK;
                const int64_t v = v0 + vi;
                Xs[t][vi] = (v < n && d0 + t < D) ? __double2float_rn(argmin_row_value(r, v, d0 + t)) : 0.f;
            }
            for (int e = tid; e < AF_TK * AF_DK; e += MMIDX_NT) {
                const int ci = e / AF_DK, t = e - ci * AF_DK;
                const int c = k0 + ci;
                Bs[t][ci] = (c < K && d0 + t < D) ? Bb[(int64_t)c * D + d0 + t] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int t = 0; t < AF_DK; ++t) {
                const float4 x4 = *reinterpret_cast<const float4 *>(&Xs[t][ty * 4]);
                const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[t][tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[t][64 + tx * 4]);
                const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
                const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xv[i], bv[j], acc[i][j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = k0 + af_centroid(tx, j);
            if (c < K) {
                const float cc = b2b[c];
#pragma unroll
                for (int i = 0; i < 4; ++i) min2_add(best[i], fmaf(-2.f, acc[i][j], cc), c);
            }
        }
    }
    // merge over the 16 threads (one half-warp) that share the same 4 vectors
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int s = 8; s > 0; s >>= 1) {
            const float o1 = __shfl_xor_sync(0xffffffffu, best[i].a1, s);
            const float o2 = __shfl_xor_sync(0xffffffffu, best[i].a2, s);
            const int oi = __shfl_xor_sync(0xffffffffu, best[i].i1, s);
            min2_merge(best[i], o1, o2, oi);
        }
    }
    if (tx == 0) {
        const double bm = (double)bmax[b];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t v = v0 + ty * 4 + i;
            if (v >= n) continue;
            const double xn = sqrt(xn2[ty * 4 + i]) * (1.0 + 1e-12);
            const double R = 1.02 * 5.9604644775390625e-08 * (2.0 * bm * bm + (2.0 * D + 8.0) * xn * bm) +
                             (D + 2.0) * 2.220446049250313e-16 * (xn + bm) * (xn + bm) + FAST_ABS_SLACK;
            const bool in_range = fast_mag_ok(xn) && fast_mag_ok(bm) && R == R && R < 1e300;
            // decided iff every other centroid is provably farther: a2 > a1 + 2R (K == 1: a2 = +inf)
            const bool decided = in_range && (double)best[i].a2 > (double)best[i].a1 + 2.0 * R;
            if (decided) {
                if (out8) out8[v * out_stride + b] = (uint8_t)best[i].i1;
                if (out16) out16[v * out_stride + b] = (uint16_t)best[i].i1;
                if (out32) out32[v * out_stride + b] = best[i].i1;
            } else {
                const int slot = atomicAdd(amb_count, 1);
                amb_list[slot] = v * nb + b;
            }
        }
    }
}

// Large codebooks (the coarse quantizer at index time, IVFPQ.computeNearestCoarseIndex IVFPQ.java:547-564): the filter matrix
// A32[n][K] comes from the GEMM kernels of coarse_fast.cuh (tensor cores or FFMA; `coef` is that kernel's radius coefficient),
// and this kernel takes each row's minimum.  One warp per row; decided / listed exactly as above.
__global__ void __launch_bounds__(MMIDX_NT) k_rowmin_filter(const float *__restrict__ A32, const double *__restrict__ X, int64_t n,
                                                            int K, int D, const float *__restrict__ cmax, double coef,
                                                            int32_t *__restrict__ out32, int64_t *__restrict__ amb_list,
                                                            float *__restrict__ amb_thr, int32_t *__restrict__ amb_count) {
    const int lane = threadIdx.x & 31;
    const int64_t v = ((int64_t)blockIdx.x * MMIDX_NT + threadIdx.x) >> 5;
    if (v >= n) return;
    const float *row = A32 + v * (int64_t)K;
    Min2 m;
    m.a1 = m.a2 = __int_as_float(0x7f800000);
    m.i1 = 0;
    for (int c = lane; c < K; c += 32) min2_add(m, row[c], c);
    double x2 = 0.0;
    for (int t = lane; t < D; t += 32) {
        const double x = X[v * (int64_t)D + t];
        x2 += x * x;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const float o1 = __shfl_xor_sync(0xffffffffu, m.a1, s);
        const float o2 = __shfl_xor_sync(0xffffffffu, m.a2, s);
        const int oi = __shfl_xor_sync(0xffffffffu, m.i1, s);
        min2_merge(m, o1, o2, oi);
        x2 += __shfl_xor_sync(0xffffffffu, x2, s);
    }
    if (lane == 0) {
        const double xn = sqrt(x2) * (1.0 + 1e-12), cm = (double)cmax[0];
        const double R = 1.02 * 5.9604644775390625e-08 * 2.0 * cm * cm + coef * xn * cm +
                         (D + 2.0) * 2.220446049250313e-16 * (xn + cm) * (xn + cm) + FAST_ABS_SLACK;
        const bool in_range = fast_mag_ok(xn) && fast_mag_ok(cm) && R == R && R < 1e300;
        if (in_range && (double)m.a2 > (double)m.a1 + 2.0 * R) {
            out32[v] = m.i1;
        } else {
            const int slot = atomicAdd(amb_count, 1);
            amb_list[slot] = v;
            // every centroid that can still be the exact argmin has a <= a1 + 2R; outside the fp32-safe window: all of them
            amb_thr[slot] = in_range ? __double2float_ru((double)m.a1 + 2.0 * R) : __int_as_float(0x7f800000);
        }
    }
}

// the listed rows in binary64, but only over the centroids the filter could not exclude (a <= thr): typically two or three of
// the K.  One warp per row; a lane that finds a candidate adds its D terms for t ascending (the reference's loop); strict `<`
// and the (distance, index) merge keep the lowest index among equal minima.
__global__ void __launch_bounds__(MMIDX_NT) k_rowmin_exact_list(const float *__restrict__ A32, const double *__restrict__ X,
                                                                const double *__restrict__ B, int K, int D,
                                                                const int64_t *__restrict__ amb_list, const float *__restrict__ amb_thr,
                                                                const int32_t *__restrict__ amb_count, int32_t *__restrict__ out32) {
    const int na = *amb_count;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * MMIDX_NT) >> 5;
    for (int a = (blockIdx.x * MMIDX_NT + threadIdx.x) >> 5; a < na; a += warps) {
        const int64_t v = amb_list[a];
        const float thr = amb_thr[a];
        const float *row = A32 + v * (int64_t)K;
        const double *x = X + v * (int64_t)D;
        double best = 1.7976931348623157e308;  // Double.MAX_VALUE
        int bidx = 0x7fffffff;
        const bool all = thr == __int_as_float(0x7f800000);  // outside the fp32-safe window the filter values mean nothing
        for (int c = lane; c < K; c += 32) {
            if (!all && !(row[c] <= thr)) continue;
            const double *bc = B + (int64_t)c * D;
            double acc = 0.0;
            for (int t = 0; t < D; ++t) acc = sqacc(acc, bc[t], x[t]);
            if (acc < best) {
                best = acc;
                bidx = c;
            }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, s);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, s);
            if (ob < best || (ob == best && oi < bidx)) {
                best = ob;
                bidx = oi;
            }
        }
        if (lane == 0) out32[v] = bidx == 0x7fffffff ? -1 : bidx;
    }
}

// the listed (vector, codebook) pairs in binary64, the reference's loop: one warp per pair, lane <-> centroids c = lane, lane+32..;
// each lane adds its centroid's terms for t ascending; strict `<` keeps the lowest index among equal minima.  grid: any.
__global__ void __launch_bounds__(MMIDX_NT) k_argmin_exact_list(ArgminRows rows, const double *__restrict__ B, int K, int D, int nb,
                                                                const int64_t *__restrict__ amb_list,
                                                                const int32_t *__restrict__ amb_count, uint8_t *__restrict__ out8,
                                                                uint16_t *__restrict__ out16, int32_t *__restrict__ out32,
                                                                int64_t out_stride) {
    const int na = *amb_count;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * MMIDX_NT) >> 5;
    for (int a = (blockIdx.x * MMIDX_NT + threadIdx.x) >> 5; a < na; a += warps) {
        const int64_t item = amb_list[a];
        const int64_t v = item / nb;
        const int b = (int)(item - v * nb);
        ArgminRows r = rows;
        r.col0 = rows.col0 + b * D;
        const double *Bb = B + (int64_t)b * K * D;
        double best = 1.7976931348623157e308;  // Double.MAX_VALUE
        int bidx = 0x7fffffff;
        for (int c = lane; c < K; c += 32) {
            const double *bc = Bb + (int64_t)c * D;
            double acc = 0.0;
            for (int t = 0; t < D; ++t) acc = sqacc(acc, bc[t], argmin_row_value(r, v, t));
            if (acc < best) {
                best = acc;
                bidx = c;
            }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, s);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, s);
            if (ob < best || (ob == best && oi < bidx)) {
                best = ob;
                bidx = oi;
            }
        }
        if (lane == 0) {
            const int idx = bidx == 0x7fffffff ? -1 : bidx;
            if (out8) out8[v * out_stride + b] = (uint8_t)idx;
            if (out16) out16[v * out_stride + b] = (uint16_t)idx;
            if (out32) out32[v * out_stride + b] = idx;
        }
    }
}

}  // namespace mmidx
