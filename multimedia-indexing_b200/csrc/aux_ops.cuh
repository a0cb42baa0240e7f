// aux_ops.cuh -- the steps right before / after the hot path, on the device:
//   k_rows_normalize   Normalization.normalizePower + normalizeL2            J/utilities/Normalization.java:21-37,74-79
//                      as used by VladAggregatorMultipleVocabularies          J/aggregation/VladAggregatorMultipleVocabularies.java:84-101
//   k_pca_project      PCA.sampleToEigenSpace (mean subtraction + V_t x)      J/dimreduction/PCA.java:188-208
//   k_rotate_vectors   RandomRotation.rotate applied before PQ                J/utilities/RandomRotation.java:44-49,
//                                                                             PQ.java:237-241,294-298 IVFPQ.java:319-323,420-424
// Numeric model as everywhere in this library: binary64, one rounding per operation, sums in the reference's index order.
#pragma once
#include "common.cuh"

namespace mmidx {

// One CTA per row v[len] (rows `ld` apart).  do_power: v[i] = signum(v[i]) * pow(|v[i]|, a)  (Normalization.java:74-79);
// a == 0.5 uses the correctly rounded square root (Java's Math.pow may differ from it in the last bit).
// do_l2: norm = sqrt(sum_i v[i]*v[i]) with the squares added for i ascending (one thread walks the staged squares, so the
// norm has the reference's bits), then v[i] /= norm, or every element = 1 when norm == 0 (Normalization.java:21-37).
constexpr int NORM_CHUNK = 4096;

__global__ void __launch_bounds__(MMIDX_NT) k_rows_normalize(double *__restrict__ X, int64_t ld, int len, int do_power, double a,
                                                             int do_l2) {
    __shared__ double sq[NORM_CHUNK];
    __shared__ double s_norm;
    double *v = X + (int64_t)blockIdx.x * ld;
    double acc = 0.0;
    for (int c0 = 0; c0 < len; c0 += NORM_CHUNK) {
        const int n = min(NORM_CHUNK, len - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += MMIDX_NT) {
            double x = v[c0 + i];
            if (do_power) {
                const double mag = (a == 0.5) ? sqrt(fabs(x)) : pow(fabs(x), a);
                const double sg = (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : x);  // Math.signum: +-0 and NaN pass through
                x = __dmul_rn(sg, mag);
                v[c0 + i] = x;
            }
            sq[i] = __dmul_rn(x, x);
        }
        __syncthreads();
        if (do_l2 && threadIdx.x == 0)
            for (int i = 0; i < n; ++i) acc = __dadd_rn(acc, sq[i]);
    }
    if (!do_l2) return;
    if (threadIdx.x == 0) s_norm = sqrt(acc);
    __syncthreads();
    const double norm = s_norm;
    for (int i = threadIdx.x; i < len; i += MMIDX_NT) v[i] = (norm == 0.0) ? 1.0 : __ddiv_rn(v[i], norm);
}

// Y[n][nc] = (X[n][ss] - mean) V_t^T, V_t[nc][ss]: y_i = sum_j V_t[i][j] * (x_j - mean_j), products rounded then added
// for j ascending (the row-times-vector loop; EJML's own order is un-vendored third-party code, so parity with the Java
// path is claimed at 1e-4 relative only).  64 x 64 output tile per CTA, 4 x 4 per thread, K step 16.
constexpr int PCA_T = 64, PCA_K = 16;

__global__ void __launch_bounds__(MMIDX_NT) k_pca_project(const double *__restrict__ X, const double *__restrict__ means,
                                                          const double *__restrict__ Vt, int64_t n, int nc, int ss,
                                                          double *__restrict__ Y) {
    __shared__ double As[PCA_K][PCA_T + 1];
    __shared__ double Bs[PCA_K][PCA_T + 1];
    const int tid = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * PCA_T;
    const int c0 = blockIdx.x * PCA_T;
    const int tr = tid >> 4, tc = tid & 15;  // rows tr*4.., columns tc + 16*j
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int k0 = 0; k0 < ss; k0 += PCA_K) {
        __syncthreads();
        for (int e = tid; e < PCA_T * PCA_K; e += MMIDX_NT) {
            const int r = e / PCA_K, kk = e - r * PCA_K;
            const int64_t row = r0 + r;
            const int col = k0 + kk;
            As[kk][r] = (row < n && col < ss) ? __dsub_rn(X[row * (int64_t)ss + col], means[col]) : 0.0;  // CommonOps.sub
            const int comp = c0 + r;
            Bs[kk][r] = (comp < nc && col < ss) ? Vt[(int64_t)comp * ss + col] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < PCA_K; ++kk) {
            double av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[kk][tr * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tc + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __dadd_rn(acc[i][j], __dmul_rn(bv[j], av[i]));
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t row = r0 + tr * 4 + i;
        if (row >= n) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int comp = c0 + tc + 16 * j;
            if (comp < nc) Y[row * (int64_t)nc + comp] = acc[i][j];
        }
    }
}

// out[g][j] = sum_i v[i] * R[i][j], i ascending, product rounded then added (CommonOps.mult of the 1 x d row vector with
// the d x d matrix, RandomRotation.java:44-49).  v = X[g] for a flat PQ, or the residual C[list] - X[g / w]
// (IVFPQ.java:316-323 at index time with w == 1 and list = the assigned lists, :417-424 at search time with
// list = probes[g]).  grid = number of vectors, thread <-> output component j (R rows are read coalesced).
__global__ void __launch_bounds__(MMIDX_NT) k_rotate_vectors(const double *__restrict__ X, const double *__restrict__ C,
                                                             const int32_t *__restrict__ list, int w,
                                                             const double *__restrict__ R, int d, double *__restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *v = reinterpret_cast<double *>(smem_raw);  // [d]
    const int64_t g = blockIdx.x;
    const double *x = X + (g / w) * (int64_t)d;
    const double *c = list ? C + (int64_t)list[g] * d : nullptr;
    for (int i = threadIdx.x; i < d; i += MMIDX_NT) v[i] = c ? __dsub_rn(c[i], x[i]) : x[i];
    __syncthreads();
    for (int j = threadIdx.x; j < d; j += MMIDX_NT) {
        double acc = 0.0;
        for (int i = 0; i < d; ++i) acc = __dadd_rn(acc, __dmul_rn(v[i], R[(int64_t)i * d + j]));
        out[g * (int64_t)d + j] = acc;
    }
}

}  // namespace mmidx
