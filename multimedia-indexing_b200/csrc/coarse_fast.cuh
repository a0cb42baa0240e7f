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
#include <cuda_bf16.h>

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

// ---- the same filter matrix on the tensor cores (north_star: "tensor cores only for the batched coarse-quantizer GEMM") ----
// bf16 x 3 split: q = qh + ql + eq, C = ch + cl + ec with bf16 qh, ql, ch, cl (|ql| <= 2^-9 |q|, |eq| <= (2^-18 + 2^-24) |q|),
//   q.C ~= sum_j qh ch + (qh cl + ql ch)                                (the ql cl term is dropped)
// evaluated with mma.sync.m16n8k16 (SASS HMMA.16816.F32.BF16): bf16 x bf16 products are exact in fp32; the main chain and the
// 2^-9-times-smaller correction chain have their own fp32 accumulators.  Because the filter is only ever used to REJECT
// (k_coarse_verify ranks the survivors on exact binary64 distances) it needs a valid error radius, not IEEE semantics:
//   representation   |q.C - split| <= 3.1 * 2^-18 ||q|| ||C||
//   accumulation     one HMMA adds 16 exact products to the accumulator; modelled as |err| <= 2^-18 (sum |products| + |acc|)
//                    (published measurements of NVIDIA tensor cores: truncation to the largest addend's ulp, <= 17 * 2^-23;
//                    2^-18 doubles that), d/16 instructions per chain: <= 2^-18 (1 + d/16) ||q|| ||C||
//   fp32 epilogue    c2 - 2 (main + corr): a few u
// => coefficient TC_COEF(d) on ||q|| Cmax below (k_coarse_verify's radius takes it instead of the FFMA one).  The
// accumulation model is an assumption about the hardware, so mmidx_set_coarse_quantizer MEASURES the kernel against binary64
// on the loaded centroids before enabling it (self_check_coarse_mma; margin 4x) and falls back to k_coarse_f32 otherwise.
__host__ __device__ inline double coarse_coef_ffma(int d) { return 1.02 * 5.9604644775390625e-08 * (2.0 * d + 8.0); }
__host__ __device__ inline double coarse_coef_mma(int d) {
    return 1.05 * 2.0 * 3.814697265625e-06 * (3.1 + 1.0 + d / 16.0) + 8.0 * 5.9604644775390625e-08;
}

// Ch / Cl: bf16 split of the coarse centroids, [nlist][dpad] (dpad = d rounded up to 16, zero filled)
__global__ void k_coarse_split_tables(const double *__restrict__ C, int nlist, int d, int dpad, unsigned short *__restrict__ Ch,
                                      unsigned short *__restrict__ Cl) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)nlist * dpad) return;
    const int c = (int)(e / dpad), j = (int)(e - (int64_t)c * dpad);
    float f = 0.f;
    if (j < d) f = __double2float_rn(C[(int64_t)c * d + j]);
    const __nv_bfloat16 h = __float2bfloat16_rn(f);
    const __nv_bfloat16 l = __float2bfloat16_rn(f - __bfloat162float(h));
    Ch[e] = __bfloat16_as_ushort(h);
    Cl[e] = __bfloat16_as_ushort(l);
}

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// CTA tile 128 queries x 64 centroids, 8 warps as 4 x 2, warp tile 32 x 32 (2 x 4 HMMA tiles), K chunks of 64 in two
// shared-memory stages filled with cp.async (the next chunk lands while the current one is multiplied); fragments come
// from ldmatrix.x4.  Both operands arrive pre-split (k_coarse_split_tables: the queries once per batch instead of once per
// column of CTAs -- at nlist = 8192 that was 128 conversions of every query from binary64).
// Shared rows are padded to 72 bf16 (144 bytes): the 8 rows of an ldmatrix phase fall into distinct bank groups.
constexpr int TM_BM = 128, TM_BN = 64, TM_BK = 64, TM_LD = TM_BK + 8, TM_STAGES = 2;
constexpr size_t TM_STAGE_ELEMS = (size_t)2 * (TM_BM + TM_BN) * TM_LD;
constexpr size_t TM_SMEM = TM_STAGES * TM_STAGE_ELEMS * sizeof(unsigned short);

__device__ __forceinline__ void cp_async_16(void *smem, const void *gmem, bool valid) {
    const int sz = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void *smem) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(smem)));
}

__global__ void __launch_bounds__(MMIDX_NT) k_coarse_mma(const unsigned short *__restrict__ Qh, const unsigned short *__restrict__ Ql,
                                                         const unsigned short *__restrict__ Ch, const unsigned short *__restrict__ Cl,
                                                         const float *__restrict__ c2, int64_t nq, int nlist, int dpad,
                                                         float *__restrict__ A32) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned short *sm = reinterpret_cast<unsigned short *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;  // warp tile origin: rows wm*32, columns wn*32
    const int g = lane >> 2, tig = lane & 3;
    const int64_t q0 = (int64_t)blockIdx.y * TM_BM;
    const int n0 = blockIdx.x * TM_BN;
    float main_[2][4][4], corr[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) main_[i][j][r] = corr[i][j][r] = 0.f;
    auto load_stage = [&](int stage, int k0) {
        unsigned short *Ah = sm + stage * TM_STAGE_ELEMS, *Al = Ah + TM_BM * TM_LD;
        unsigned short *Bh = Al + TM_BM * TM_LD, *Bl = Bh + TM_BN * TM_LD;
        for (int e = tid; e < TM_BM * (TM_BK / 8); e += MMIDX_NT) {
            const int r = e >> 3, k8 = (e & 7) * 8;
            const bool valid = q0 + r < nq && k0 + k8 < dpad;
            const int64_t off = valid ? (q0 + r) * (int64_t)dpad + k0 + k8 : 0;
            cp_async_16(Ah + r * TM_LD + k8, Qh + off, valid);
            cp_async_16(Al + r * TM_LD + k8, Ql + off, valid);
        }
        for (int e = tid; e < TM_BN * (TM_BK / 8); e += MMIDX_NT) {
            const int r = e >> 3, k8 = (e & 7) * 8;
            const bool valid = n0 + r < nlist && k0 + k8 < dpad;
            const int64_t off = valid ? (int64_t)(n0 + r) * dpad + k0 + k8 : 0;
            cp_async_16(Bh + r * TM_LD + k8, Ch + off, valid);
            cp_async_16(Bl + r * TM_LD + k8, Cl + off, valid);
        }
        asm volatile("cp.async.commit_group;");
    };
    const int nk = (dpad + TM_BK - 1) / TM_BK;
    load_stage(0, 0);
    for (int kc = 0; kc < nk; ++kc) {
        if (kc + 1 < nk) {
            load_stage((kc + 1) & 1, (kc + 1) * TM_BK);
            asm volatile("cp.async.wait_group 1;");  // everything but the chunk just requested has landed
        } else {
            asm volatile("cp.async.wait_group 0;");
        }
        __syncthreads();
        const unsigned short *Ah = sm + (kc & 1) * TM_STAGE_ELEMS, *Al = Ah + TM_BM * TM_LD;
        const unsigned short *Bh = Al + TM_BM * TM_LD, *Bl = Bh + TM_BN * TM_LD;
#pragma unroll
        for (int ks = 0; ks < TM_BK; ks += 16) {
            uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                // matrices: (rows 0-7, k 0-7), (rows 8-15, k 0-7), (rows 0-7, k 8-15), (rows 8-15, k 8-15) = a0..a3
                const int off = (wm * 32 + i * 16 + (lane & 15)) * TM_LD + ks + 8 * (lane >> 4);
                ldmatrix_x4(ah[i], Ah + off);
                ldmatrix_x4(al[i], Al + off);
            }
#pragma unroll
            for (int jp = 0; jp < 2; ++jp) {
                // matrices: (n 0-7, k 0-7), (n 0-7, k 8-15), (n 8-15, k 0-7), (n 8-15, k 8-15) = b[2jp][0..1], b[2jp+1][0..1]
                const int off = (wn * 32 + jp * 16 + (lane & 7) + ((lane >> 4) << 3)) * TM_LD + ks + 8 * ((lane >> 3) & 1);
                uint32_t r[4];
                ldmatrix_x4(r, Bh + off);
                bh[2 * jp][0] = r[0], bh[2 * jp][1] = r[1], bh[2 * jp + 1][0] = r[2], bh[2 * jp + 1][1] = r[3];
                ldmatrix_x4(r, Bl + off);
                bl[2 * jp][0] = r[0], bl[2 * jp][1] = r[1], bl[2 * jp + 1][0] = r[2], bl[2 * jp + 1][1] = r[3];
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    mma_bf16_16816(main_[i][j], ah[i], bh[j]);
                    mma_bf16_16816(corr[i][j], ah[i], bl[j]);
                    mma_bf16_16816(corr[i][j], al[i], bh[j]);
                }
        }
        __syncthreads();  // the stage is refilled by the request of the next trip
    }
    // a = c2[c] - 2 (main + corr)
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int64_t q = q0 + wm * 32 + i * 16 + g + half * 8;
            if (q >= nq) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + wn * 32 + j * 8 + 2 * tig;
                const float d0 = main_[i][j][half * 2] + corr[i][j][half * 2];
                const float d1 = main_[i][j][half * 2 + 1] + corr[i][j][half * 2 + 1];
                if (n < nlist) A32[q * (int64_t)nlist + n] = fmaf(-2.f, d0, c2[n]);
                if (n + 1 < nlist) A32[q * (int64_t)nlist + n + 1] = fmaf(-2.f, d1, c2[n + 1]);
            }
        }
}

// max over a sample of |a32 - (||C||^2 - 2 q.C)| / (||q|| Cmax) in binary64: what mmidx_set_coarse_quantizer compares with
// the coefficient above.  Queries = the first nsamp centroids (perturbed by the caller); one thread per (q, c) pair.
__global__ void k_coarse_filter_error(const double *__restrict__ Q, const double *__restrict__ C, const float *__restrict__ A32,
                                      int nsamp, int nlist, int d, double cmax, double *__restrict__ out_max) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double ratio = 0.0;
    if (e < (int64_t)nsamp * nlist) {
        const int q = (int)(e / nlist), c = (int)(e - (int64_t)q * nlist);
        double dot = 0.0, c2 = 0.0, q2 = 0.0;
        for (int j = 0; j < d; ++j) {
            const double qv = Q[(int64_t)q * d + j], cv = C[(int64_t)c * d + j];
            dot = fma(qv, cv, dot);
            c2 = fma(cv, cv, c2);
            q2 = fma(qv, qv, q2);
        }
        const double t = c2 - 2.0 * dot;
        const double den = sqrt(q2) * cmax;
        // the c2 rounding (u Cmax^2) is accounted for separately in the radius: take it out of the measurement
        const double err = fmax(0.0, fabs((double)A32[e] - t) - 1.2e-7 * cmax * cmax);
        ratio = den > 0.0 ? err / den : 0.0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ratio = fmax(ratio, __shfl_xor_sync(0xffffffffu, ratio, o));
    if ((threadIdx.x & 31) == 0 && ratio > 0.0)
        atomicMax(reinterpret_cast<unsigned long long *>(out_max), (unsigned long long)__double_as_longlong(ratio));
}

// grid nq.  Shared memory: TopK<CAP> | qv [d] | keys [nlist] fp32 | surv [CAP] | xs [vb][d + 1]

// KPT > 0: nlist <= KPT * 256 and every thread keeps its KPT filter keys in registers (no key array in shared memory: at
// nlist = 8192 that array was 32 KB per CTA and every pass re-read it); KPT == 0: keys staged in shared memory, any nlist.
template <int CAP, int KPT>
__global__ void __launch_bounds__(MMIDX_NT, KPT == 0 ? (CAP <= 256 ? 6 : 4) : (KPT <= 4 ? 6 : (KPT <= 16 ? 4 : 3))) k_coarse_verify(const double *__restrict__ Q, const double *__restrict__ C,
                                                            const float *__restrict__ A32, const float *__restrict__ cmax,
                                                            int nlist, int d, int w, int vb, double coef, TopkOut o) {
    constexpr bool REGK = KPT > 0;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TopK<CAP> &tk = *reinterpret_cast<TopK<CAP> *>(smem_raw);
    const size_t tk_bytes = (sizeof(TopK<CAP>) + 127) & ~(size_t)127;
    double *qv = reinterpret_cast<double *>(smem_raw + tk_bytes);  // [d]
    float *key = reinterpret_cast<float *>(qv + d);                 // [nlist] (KPT == 0 only)
    int *surv = reinterpret_cast<int *>(key + (REGK ? 0 : ((nlist + 1) & ~1)));  // [CAP] centroid ids, later scratch of the tie rule
    double *xs = reinterpret_cast<double *>(surv + CAP);            // [vb][d + 1] squared terms of a batch of survivors
    __shared__ unsigned int hist2[2][256];  // the passes alternate, so that zeroing the next one needs no barrier of its own
    __shared__ int s_bin, s_krem, s_ns;
    __shared__ double s_q2;
    const int tid = threadIdx.x;
    const int64_t q = blockIdx.x;
    const float *arow = A32 + q * (int64_t)nlist;
    for (int i = tid; i < d; i += MMIDX_NT) qv[i] = Q[q * (int64_t)d + i];
    float kreg[REGK ? KPT : 1];
    if (REGK) {
#pragma unroll
        for (int r = 0; r < (REGK ? KPT : 1); ++r) {
            const int i = tid + r * MMIDX_NT;
            kreg[r] = i < nlist ? arow[i] : 0.f;
        }
    } else {
        for (int i = tid; i < nlist; i += MMIDX_NT) key[i] = arow[i];
    }
    if (tid == 0) s_ns = 0;
    hist2[0][tid] = 0;
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
        unsigned int *hist = hist2[pass & 1];
        hist2[(pass + 1) & 1][tid] = 0;
        auto count_key = [&](int i, float kv) {
            bool act = i < nlist;
            unsigned bin = 0;
            if (act) {
                unsigned u = __float_as_uint(kv);
                u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
                act = (pass == 0 || (u >> (shift + 8)) == prefix);
                bin = (u >> shift) & 255u;
            }
            hist_add(hist, act, bin);
        };
        if (REGK) {
#pragma unroll
            for (int r = 0; r < (REGK ? KPT : 1); ++r)
                if (r * MMIDX_NT < nlist) count_key(tid + r * MMIDX_NT, kreg[r]);  // block-uniform guard
        } else {
            for (int i0 = 0; i0 < nlist; i0 += MMIDX_NT) count_key(i0 + tid, (i0 + tid < nlist) ? key[i0 + tid] : 0.f);
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
    // coef: coarse_coef_ffma(d) or coarse_coef_mma(d), whichever kernel produced A32
    const double B = 1.02 * 5.9604644775390625e-08 * 2.0 * cm * cm + coef * qn * cm +
                     (d + 2.0) * 2.220446049250313e-16 * (qn + cm) * (qn + cm);
    const float lim = __double2float_ru((double)kth + 2.0 * B + 1e-30);
    auto keep_key = [&](int i, float kv) {
        const bool pred = i < nlist && kv <= lim;
        const unsigned mask = __ballot_sync(0xffffffffu, pred);
        if (mask) {
            const int lane = tid & 31, leader = __ffs(mask) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&s_ns, __popc(mask));
            base = __shfl_sync(0xffffffffu, base, leader);
            const int slot = base + __popc(mask & ((1u << lane) - 1u));
            if (pred && slot < CAP) surv[slot] = i;
        }
    };
    if (REGK) {
#pragma unroll
        for (int r = 0; r < (REGK ? KPT : 1); ++r)
            if (r * MMIDX_NT < nlist) keep_key(tid + r * MMIDX_NT, kreg[r]);
    } else {
        for (int i0 = 0; i0 < nlist; i0 += MMIDX_NT) keep_key(i0 + tid, (i0 + tid < nlist) ? key[i0 + tid] : 0.f);
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
                if ((d & 1) == 0) {  // two dimensions per lane and 128-bit load (rows are 16-byte aligned when d is even)
                    for (int j = 2 * (tid & 31); j < d; j += 64) {
                        const double2 cv = *reinterpret_cast<const double2 *>(cr + j);
                        const double2 qq = *reinterpret_cast<const double2 *>(qv + j);
                        const double d0 = __dsub_rn(cv.x, qq.x), d1 = __dsub_rn(cv.y, qq.y);
                        xr[j] = __dmul_rn(d0, d0);
                        xr[j + 1] = __dmul_rn(d1, d1);
                    }
                } else {
                    for (int j = tid & 31; j < d; j += 32) {
                        const double df = __dsub_rn(cr[j], qv[j]);
                        xr[j] = __dmul_rn(df, df);
                    }
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
