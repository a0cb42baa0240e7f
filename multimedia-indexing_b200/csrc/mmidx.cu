// mmidx.cu -- host runtime + C ABI of libmmidx.so (see include/mmidx.h).
//
// Index objects own their HBM storage:
//   Linear : vectors packed in blocks of 32 (dimension-major inside a block)      Linear.java:34,80 vectorsList
//   PQ     : flat code array [n][code_bytes], iid == position                     PQ.java:65,71 pqByteCodes/pqShortCodes
//   IVFPQ  : append log (codes + list ids in iid order) and, once sealed, a CSR copy grouped by list in
//            insertion order with 16-entry aligned list starts                    IVFPQ.java:72-83 invertedLists/pqByteCodes[l]
// There is no CPU compute path: every entry point that does arithmetic launches the sm_100a kernels.
#include "../../include/mmidx.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <map>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <unistd.h>
#include <vector>

#include "kernels.cuh"
#include "tie_resolve.cuh"
#include "coarse_fast.cuh"
#include "fast_scan.cuh"
#include "comm.cuh"
#include "aux_ops.cuh"
#include "argmin_filter.cuh"

using namespace mmidx;

// ---------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return fail(MMIDX_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)
#define RET(call)                 \
    do {                          \
        int r_ = (call);          \
        if (r_ != MMIDX_OK) return r_; \
    } while (0)

extern "C" const char *mmidx_last_error(void) { return g_err.c_str(); }
extern "C" const char *mmidx_version(void) { return "libmmidx 0.1 sm_100a"; }

// ---------------------------------------------------------------------------------------------------------
// device buffers
// ---------------------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    // grow to at least `bytes`, keeping the first `keep` bytes
    int reserve(size_t bytes, size_t keep, cudaStream_t st) {
        if (bytes <= cap) return MMIDX_OK;
        size_t ncap = std::max(bytes, cap + cap / 2);
        ncap = (ncap + 255) & ~(size_t)255;
        void *np = nullptr;
        CK(cudaMalloc(&np, ncap));
        if (p && keep) CK(cudaMemcpyAsync(np, p, keep, cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
        if (p) cudaFree(p);
        p = np;
        cap = ncap;
        return MMIDX_OK;
    }
    template <typename T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

// stream-ordered scratch, released (stream-ordered) when the call returns
// Per-(index, stream) scratch arena of the search hot path: one grow-only device block with stack (bump / rollback)
// allocation, so a search call issues no cudaMallocAsync / cudaFreeAsync once the arena has seen its batch size (a
// search call makes ~25 scratch allocations; at small per-GPU batches the call is host-bound, profiles/README.md).
// Re-use across calls is safe because consecutive calls on one stream are stream-ordered; a second host thread that
// finds the arena busy falls back to stream-ordered allocations.
struct Arena {
    void *base = nullptr;
    size_t cap = 0, off = 0, peak = 0;
    int depth = 0;
    std::thread::id owner;
    uint64_t gen = 0;       // bumped whenever the block is re-allocated (captured CUDA graphs hold its addresses)
    uint64_t spills = 0;    // allocations that did not fit and went to cudaMallocAsync
};

struct Scratch {
    cudaStream_t st;
    std::vector<void *> ptrs;  // stream-ordered allocations that did not come from an arena
    Arena *ar = nullptr;
    std::mutex *amu = nullptr;
    size_t mark = 0;
    explicit Scratch(cudaStream_t s) : st(s) {}
    Scratch(cudaStream_t s, std::map<cudaStream_t, Arena> &arenas, std::mutex &mu) : st(s) {
        std::lock_guard<std::mutex> lk(mu);
        Arena &a = arenas[s];
        if (a.depth == 0 || a.owner == std::this_thread::get_id()) {
            a.owner = std::this_thread::get_id();
            a.depth++;
            ar = &a;  // std::map nodes are stable
            amu = &mu;
            mark = a.off;
        }
    }
    Scratch(const Scratch &) = delete;
    Scratch &operator=(const Scratch &) = delete;
    ~Scratch() {
        for (void *p : ptrs) cudaFreeAsync(p, st);
        if (!ar) return;
        std::lock_guard<std::mutex> lk(*amu);
        ar->off = mark;
        if (--ar->depth == 0 && ar->peak > ar->cap) {  // outermost scope: grow for the next call
            if (ar->base) cudaFreeAsync(ar->base, st);
            ar->base = nullptr;
            ar->cap = 0;
            ar->gen++;
            const size_t want = ar->peak + ar->peak / 4;
            void *p = nullptr;
            if (cudaMallocAsync(&p, want, st) == cudaSuccess) {
                ar->base = p;
                ar->cap = want;
            } else {
                cudaGetLastError();
            }
            ar->peak = 0;
        }
    }
    template <typename T>
    int get(T **out, size_t count) {
        size_t bytes = (std::max<size_t>(count * sizeof(T), 16) + 255) & ~(size_t)255;
        if (ar) {
            const size_t o = ar->off;
            ar->off += bytes;
            ar->peak = std::max(ar->peak, ar->off);
            if (ar->off <= ar->cap) {
                *out = reinterpret_cast<T *>(static_cast<unsigned char *>(ar->base) + o);
                return MMIDX_OK;
            }
            ar->spills++;
        }
        void *p = nullptr;
        CK(cudaMallocAsync(&p, bytes, st));
        ptrs.push_back(p);
        *out = reinterpret_cast<T *>(p);
        return MMIDX_OK;
    }
};

struct StageTimer {
    // events bracket stages of every chunk; summed lazily by mmidx_last_timings
    struct Span {
        cudaEvent_t a, b;
        int stage;
    };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    bool enabled = false;
    bool capturing = false;  // the stream is being captured into a CUDA graph: events become external record nodes
    void record(cudaEvent_t e, cudaStream_t st) const {
        if (capturing)
            cudaEventRecordWithFlags(e, st, cudaEventRecordExternal);
        else
            cudaEventRecord(e, st);
    }
    cudaEvent_t get() {
        if (used == pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            pool.push_back(e);
        }
        return pool[used++];
    }
    void reset() {
        spans.clear();
        used = 0;
    }
    ~StageTimer() {
        for (auto e : pool) cudaEventDestroy(e);
    }
};

struct mmidx_index {
    mmidx_params p;
    int S = 0, code_bytes = 0, device = 0;
    int shard_rank = 0, shard_count = 1;
    bool has_P = false, has_C = false, has_perm = false, has_rot = false;
    DevBuf dP, dC, dCt, dperm, dR;  // dR: RandomRotation matrix [d][d] (mmidx_set_transform)
    int64_t n = 0;        // loadCounter: vectors offered to the index (global iid counter)
    int64_t n_local = 0;  // vectors stored on this shard
    // Linear
    DevBuf dXb;
    // PQ flat / IVFPQ append log
    DevBuf dcodes;
    std::vector<int32_t> h_list;  // IVFPQ: list id of every stored entry, log order
    std::vector<int32_t> h_iid;   // IVFPQ: global iid of every stored entry, log order
    // IVFPQ CSR
    bool sealed = false;
    DevBuf csr_codes, csr_iids, dlist_off, dlist_len;
    DevBuf csr_ocodes, csr_oiids, csr_orank;  // fast path: conflict-aware order inside each list
    bool reorder = true;                      // MMIDX_REORDER=0 keeps insertion order (for A/B measurements)
    std::vector<int32_t> h_list_len;
    std::vector<int64_t> h_list_off;
    cudaStream_t stream = nullptr;
    std::map<cudaStream_t, Arena> arenas;  // search scratch per launching stream (Scratch)
    std::mutex arena_mu;
    // mmidx_search: copies overlap the kernels of other chunks; chunks alternate between `stream` and `comp2_stream`
    // so that the tail of one chunk's scan overlaps the next chunk's kernels
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr, comp2_stream = nullptr;
    std::mutex mu;
    StageTimer timer;
    int last_launches = 0;
    // fast path tables (fast_scan.cuh), rebuilt when a quantizer or the permutation changes
    DevBuf dT1, dP32t, dt1max, dpmax, dstats;
    DevBuf dPneg;             // flat PQ on the fused path: -P (the index is an IVFPQ with one zero centroid, fast_scan.cuh)
    int flat_nlist = 0;       // ... and its pseudo lists of equal length
    DevBuf dC32, dc2, dcmax;  // coarse_fast.cuh: fp32 copy of the coarse quantizer, ||C||^2, max ||C||
    DevBuf dCh, dCl;          // ... and its bf16 hi / lo split for the tensor-core filter (k_coarse_mma)
    DevBuf dPf32, dPb2, dPbmax;  // argmin_filter.cuh: fp32 product codebooks [m][ks][S], ||P_j,c||^2, max norm per sub-quantizer
    int dpad = 0;             // d rounded up to 16
    bool coarse_mma_ok = false;   // the tensor-core filter passed its measurement against binary64 on this quantizer
    double coarse_mma_measured = 0.0;  // measured error coefficient (of ||q|| Cmax), for mmidx_debug / logs
    std::vector<int32_t> shard_map;  // optional list -> owning shard (default l % shard_count)
    bool fast_ready = false;
    bool fast_len_ok = true;   // every list shorter than 2^22 entries (packed payload of the fp32 collector)
    bool long_lists = false;   // some list exceeds FAST_SEG entries: segmented sweeps (k_ivfpq_scan_fast LONG)
    bool fast_range_ok = true;    // quantizer magnitudes inside the fp32-safe window (fast_scan.cuh); else exact kernels
    bool coarse_range_ok = true;  // same for the fp32 coarse filter (coarse_fast.cuh)
    bool force_exact = false;  // MMIDX_MODE=exact
    bool want_stats = false;   // MMIDX_STATS=1
    size_t lut_chunk_bytes = (size_t)1024 << 20;  // ADC-table scratch per query chunk (MMIDX_LUT_CHUNK_MB)
    size_t smem_per_sm = 228 * 1024;  // shared memory of one SM (cudaDevAttrMaxSharedMemoryPerMultiprocessor)
    int64_t host_chunk = 2560; // queries per pipelined chunk of mmidx_search (MMIDX_CHUNK; sweep in profiles/README.md)
    int sm_count = 148;        // SMs of this device: grids are sized in CTA slots = sm_count x resident CTAs per SM
    uint64_t gen = 0;          // bumped by everything that invalidates captured search graphs
    bool use_graph = true;     // MMIDX_GRAPH=0: always enqueue kernel by kernel
    std::map<cudaStream_t, struct GraphCache *> graphs;  // captured search calls, per launching stream
    std::mutex graph_mu;                                 // guards the map itself; every cache has its own lock
    std::map<std::thread::id, cudaStream_t> thread_streams;  // mmidx_search: one stream per calling host thread
    std::mutex ts_mu;
    std::vector<cudaEvent_t> ev_pool;  // mmidx_search's pipeline events, re-used across calls
    std::mutex ev_mu;
    struct Comm *comm = nullptr;       // multi-GPU exchange (mmidx_comm_*)
};

// CTA slots of the device for kernels that keep 4 CTAs per SM resident (the fused scan's launch bound)
static inline int cta_slots(const mmidx_index *ix) { return ix->sm_count * 4; }
#ifndef MMIDX_VERIFY_VB
#define MMIDX_VERIFY_VB 0  // survivors whose squared terms are staged per batch of k_coarse_verify (0: chosen from w)
#endif
#ifndef MMIDX_VERIFY_REGKEYS
#define MMIDX_VERIFY_REGKEYS 1  // A/B switch: k_coarse_verify keeps its filter keys in registers when nlist <= 8192
#endif

static bool fast_eligible(const mmidx_index *ix);
static void drop_graphs(mmidx_index *ix);
static void free_graph_caches(mmidx_index *ix);  // GraphCache is complete only further down
static void comm_release(mmidx_index *ix);

static int check_device(int device) {
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0)
        return fail(MMIDX_ERR_CUDA, "no CUDA device: libmmidx has no CPU path (%s)", cudaGetErrorString(e));
    if (device >= cnt) return fail(MMIDX_ERR_CUDA, "device %d out of range (%d devices)", device, cnt);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(MMIDX_ERR_CUDA, "device %d is sm_%d%d; libmmidx is built for sm_100a only", device, prop.major,
                    prop.minor);
    return MMIDX_OK;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

// opt-in dynamic shared memory of a kernel; the attribute call takes a driver lock (~10 us), so the largest size set
// per (device, kernel) is remembered and the call is skipped on the hot path
template <typename K>
static int set_smem(K kernel, size_t bytes) {
    if (bytes > 227 * 1024) return fail(MMIDX_ERR_UNSUPPORTED, "kernel needs %zu bytes of shared memory", bytes);
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, size_t> done;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    const auto key = std::make_pair(dev, reinterpret_cast<const void *>(kernel));
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = done.find(key);
        if (it != done.end() && it->second >= bytes) return MMIDX_OK;
    }
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    std::lock_guard<std::mutex> lk(mu);
    size_t &v = done[key];
    v = std::max(v, bytes);
    return MMIDX_OK;
}

// SM count of the current device (for entry points that take no index)
static int current_sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
        return 148;
    return n;
}

static int post_launch(const char *name, int *launches) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(MMIDX_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e));
    if (launches) ++*launches;
    return MMIDX_OK;
}

// Tensor-core coarse filter of n rows of X against the index's split centroid tables: A32[n][nlist].  The rows are split into
// bf16 hi / lo once (Xh, Xl: [n][dpad] scratch of the caller), then k_coarse_mma runs on pre-split operands only.
static int launch_coarse_mma(mmidx_index *ix, const double *X, int64_t n, float *A32, unsigned short *Xh, unsigned short *Xl,
                             cudaStream_t st, int *launches) {
    const int d = ix->p.d, nlist = ix->p.nlist, dpad = ix->dpad;
    const size_t ne = (size_t)n * dpad;
    k_coarse_split_tables<<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(X, (int)n, d, dpad, Xh, Xl);
    RET(post_launch("k_coarse_split_tables", launches));
    RET(set_smem(k_coarse_mma, TM_SMEM));
    dim3 gg((unsigned)((nlist + TM_BN - 1) / TM_BN), (unsigned)((n + TM_BM - 1) / TM_BM));
    k_coarse_mma<<<gg, MMIDX_NT, TM_SMEM, st>>>(Xh, Xl, ix->dCh.as<unsigned short>(), ix->dCl.as<unsigned short>(), ix->dc2.as<float>(), n,
                                                nlist, dpad, A32);
    return post_launch("k_coarse_mma", launches);
}

// ---------------------------------------------------------------------------------------------------------
// lifecycle
// ---------------------------------------------------------------------------------------------------------
extern "C" int mmidx_create(const mmidx_params *pp, mmidx_t **out) {
    if (!pp || !out) return fail(MMIDX_ERR_INVALID, "null argument");
    mmidx_params p = *pp;
    if (p.type < MMIDX_LINEAR || p.type > MMIDX_IVFPQ) return fail(MMIDX_ERR_INVALID, "unknown index type %d", p.type);
    if (p.d < 1) return fail(MMIDX_ERR_DIM, "vectorLength must be >= 1");
    if (p.max_n < 0) return fail(MMIDX_ERR_INVALID, "maxNumVectors < 0");
    if (p.type != MMIDX_LINEAR) {
        // PQ.java:148-150: "The given number of subvectors is not valid!"
        if (p.m < 1 || p.d % p.m != 0) return fail(MMIDX_ERR_DIM, "The given number of subvectors is not valid!");
        if (p.ks < 1 || p.ks > 65536) return fail(MMIDX_ERR_UNSUPPORTED, "numProductCentroids must be in 1..65536");
    }
    if (p.type == MMIDX_IVFPQ) {
        if (p.nlist < 1) return fail(MMIDX_ERR_INVALID, "numCoarseCentroids must be >= 1");
        if (p.w <= 0) p.w = (int)(p.nlist * 0.1);  // IVFPQ.java:188
    }
    int device = p.device;
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    RET(check_device(device));
    DeviceGuard g(device);
    mmidx_index *ix = new mmidx_index();
    ix->p = p;
    ix->device = device;
    ix->S = (p.type == MMIDX_LINEAR) ? 0 : p.d / p.m;
    ix->code_bytes = (p.type == MMIDX_LINEAR) ? 0 : (p.ks <= 256 ? p.m : 2 * p.m);
    ix->shard_count = p.shard_count > 1 ? p.shard_count : 1;
    ix->shard_rank = p.shard_count > 1 ? p.shard_rank : 0;
    if (ix->shard_rank < 0 || ix->shard_rank >= ix->shard_count) {
        delete ix;
        return fail(MMIDX_ERR_INVALID, "shard_rank out of range");
    }
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) == cudaSuccess && prop.multiProcessorCount > 0) {
            ix->sm_count = prop.multiProcessorCount;
            if (prop.sharedMemPerMultiprocessor > 0) ix->smem_per_sm = prop.sharedMemPerMultiprocessor;
        }
    }
    if (const char *e = getenv("MMIDX_GRAPH")) ix->use_graph = atoi(e) != 0;
    if (const char *e = getenv("MMIDX_CHUNK")) ix->host_chunk = std::max(64, atoi(e));
    if (const char *e = getenv("MMIDX_MODE")) ix->force_exact = strcmp(e, "exact") == 0;
    if (const char *e = getenv("MMIDX_STATS")) ix->want_stats = atoi(e) != 0;
    if (const char *e = getenv("MMIDX_REORDER")) ix->reorder = atoi(e) != 0;
    if (const char *e = getenv("MMIDX_LUT_CHUNK_MB")) {
        long v = atol(e);
        if (v >= 1) ix->lut_chunk_bytes = (size_t)v << 20;
    }
    cudaError_t e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ix;
        return fail(MMIDX_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    // keep freed scratch in the pool: search calls re-use it without going back to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    if (p.type == MMIDX_IVFPQ) {
        ix->h_list_len.assign(p.nlist, 0);
        ix->h_list_off.assign(p.nlist, 0);
    }
    *out = ix;
    return MMIDX_OK;
}

extern "C" int mmidx_destroy(mmidx_t *ix) {
    if (!ix) return MMIDX_OK;
    {
        DeviceGuard g(ix->device);
        if (ix->want_stats && ix->dstats.p) {  // MMIDX_STATS=1: debug counters of the fast scan kernel
            unsigned long long h[4] = {0, 0, 0, 0};
            cudaDeviceSynchronize();
            cudaMemcpy(h, ix->dstats.p, sizeof(h), cudaMemcpyDeviceToHost);
            fprintf(stderr, "[mmidx stats] candidates=%llu rescanned_lists=%llu survivors=%llu direct_fallbacks=%llu\n", h[0], h[1], h[2], h[3]);
        }
        cudaDeviceSynchronize();
        free_graph_caches(ix);
        for (auto &kv : ix->thread_streams) cudaStreamDestroy(kv.second);
        for (cudaEvent_t e : ix->ev_pool) cudaEventDestroy(e);
        comm_release(ix);
        for (auto &kv : ix->arenas)
            if (kv.second.base) cudaFree(kv.second.base);
        if (ix->stream) {
            cudaStreamSynchronize(ix->stream);
            cudaStreamDestroy(ix->stream);
            if (ix->h2d_stream) cudaStreamDestroy(ix->h2d_stream);
            if (ix->d2h_stream) cudaStreamDestroy(ix->d2h_stream);
            if (ix->comp2_stream) cudaStreamDestroy(ix->comp2_stream);
        }
    }
    DeviceGuard g(ix->device);
    delete ix;
    return MMIDX_OK;
}

__global__ void k_identity_order(const int64_t *__restrict__ list_off, const int32_t *__restrict__ list_len,
                                 int32_t *__restrict__ src) {
    const int64_t start = list_off[blockIdx.x];
    for (int i = threadIdx.x; i < list_len[blockIdx.x]; i += blockDim.x) src[start + i] = i;
}

__global__ void k_transpose(const double *__restrict__ A, int rows, int cols, double *__restrict__ At) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)rows * cols) return;
    int r = (int)(e / cols), c = (int)(e - (int64_t)r * cols);
    At[(int64_t)c * rows + r] = A[e];
}

extern "C" int mmidx_set_product_quantizer(mmidx_t *ix, const double *P) {
    if (!ix || !P) return fail(MMIDX_ERR_INVALID, "null argument");
    if (ix->p.type == MMIDX_LINEAR) return fail(MMIDX_ERR_INVALID, "Linear index has no product quantizer");
    DeviceGuard g(ix->device);
    std::lock_guard<std::mutex> lk(ix->mu);
    size_t bytes = sizeof(double) * (size_t)ix->p.m * ix->p.ks * ix->S;
    RET(ix->dP.reserve(bytes, 0, ix->stream));
    CK(cudaMemcpyAsync(ix->dP.p, P, bytes, cudaMemcpyHostToDevice, ix->stream));
    {
        // fp32 tables of the encode filter (argmin_filter.cuh)
        const int m = ix->p.m, ks = ix->p.ks, S = ix->S;
        RET(ix->dPf32.reserve(sizeof(float) * (size_t)m * ks * S, 0, ix->stream));
        RET(ix->dPb2.reserve(sizeof(float) * (size_t)m * ks, 0, ix->stream));
        RET(ix->dPbmax.reserve(sizeof(float) * (size_t)m, 0, ix->stream));
        CK(cudaMemsetAsync(ix->dPbmax.p, 0, sizeof(float) * (size_t)m, ix->stream));
        k_argmin_tables<<<dim3(ks, m), MMIDX_NT, 0, ix->stream>>>(ix->dP.as<double>(), ks, S, ix->dPf32.as<float>(), ix->dPb2.as<float>(),
                                                                 ix->dPbmax.as<float>());
        RET(post_launch("k_argmin_tables", nullptr));
    }
    CK(cudaStreamSynchronize(ix->stream));
    ix->has_P = true;
    ix->fast_ready = false;
    ix->gen++;
    return MMIDX_OK;
}

extern "C" int mmidx_set_coarse_quantizer(mmidx_t *ix, const double *C) {
    if (!ix || !C) return fail(MMIDX_ERR_INVALID, "null argument");
    if (ix->p.type != MMIDX_IVFPQ) return fail(MMIDX_ERR_INVALID, "only IVFPQ has a coarse quantizer");
    DeviceGuard g(ix->device);
    std::lock_guard<std::mutex> lk(ix->mu);
    size_t cnt = (size_t)ix->p.nlist * ix->p.d;
    RET(ix->dC.reserve(cnt * sizeof(double), 0, ix->stream));
    RET(ix->dCt.reserve(cnt * sizeof(double), 0, ix->stream));
    CK(cudaMemcpyAsync(ix->dC.p, C, cnt * sizeof(double), cudaMemcpyHostToDevice, ix->stream));
    k_transpose<<<(unsigned)((cnt + 255) / 256), 256, 0, ix->stream>>>(ix->dC.as<double>(), ix->p.nlist, ix->p.d,
                                                                      ix->dCt.as<double>());
    RET(post_launch("k_transpose", nullptr));
    // fp32 filter tables of the coarse stage (coarse_fast.cuh)
    RET(ix->dC32.reserve(cnt * sizeof(float), 0, ix->stream));
    RET(ix->dc2.reserve((size_t)ix->p.nlist * sizeof(float), 0, ix->stream));
    RET(ix->dcmax.reserve(sizeof(float), 0, ix->stream));
    CK(cudaMemsetAsync(ix->dcmax.p, 0, sizeof(float), ix->stream));
    k_coarse_tables<<<ix->p.nlist, MMIDX_NT, 0, ix->stream>>>(ix->dC.as<double>(), ix->p.d, ix->dC32.as<float>(),
                                                             ix->dc2.as<float>(), ix->dcmax.as<float>());
    RET(post_launch("k_coarse_tables", nullptr));
    CK(cudaStreamSynchronize(ix->stream));
    float cm = 0.f;
    CK(cudaMemcpy(&cm, ix->dcmax.p, sizeof(float), cudaMemcpyDeviceToHost));
    ix->coarse_range_ok = fast_mag_ok((double)cm);  // false for inf / NaN too
    // tensor-core filter: bf16 split tables, then MEASURE the kernel against binary64 on this quantizer (queries = the
    // first centroids) before trusting the error model of coarse_fast.cuh
    ix->coarse_mma_ok = false;
    ix->dpad = (ix->p.d + 15) & ~15;
    const char *cmode = getenv("MMIDX_COARSE");
    if (ix->coarse_range_ok && cm > 0.f && !(cmode && strcmp(cmode, "ffma") == 0)) {
        const int d = ix->p.d, nlist = ix->p.nlist, dpad = ix->dpad;
        const size_t ce = (size_t)nlist * dpad;
        RET(ix->dCh.reserve(ce * sizeof(unsigned short), 0, ix->stream));
        RET(ix->dCl.reserve(ce * sizeof(unsigned short), 0, ix->stream));
        k_coarse_split_tables<<<(unsigned)((ce + 255) / 256), 256, 0, ix->stream>>>(ix->dC.as<double>(), nlist, d, dpad,
                                                                                   ix->dCh.as<unsigned short>(), ix->dCl.as<unsigned short>());
        RET(post_launch("k_coarse_split_tables", nullptr));
        const int nsamp = std::min(nlist, 256);
        Scratch sc(ix->stream);
        float *A32;
        double *dmax;
        RET(sc.get(&A32, (size_t)nsamp * nlist));
        RET(sc.get(&dmax, 1));
        CK(cudaMemsetAsync(dmax, 0, sizeof(double), ix->stream));
        unsigned short *Xh, *Xl;
        RET(sc.get(&Xh, (size_t)nsamp * dpad));
        RET(sc.get(&Xl, (size_t)nsamp * dpad));
        RET(launch_coarse_mma(ix, ix->dC.as<double>(), nsamp, A32, Xh, Xl, ix->stream, nullptr));
        k_coarse_filter_error<<<(unsigned)(((size_t)nsamp * nlist + 255) / 256), 256, 0, ix->stream>>>(
            ix->dC.as<double>(), ix->dC.as<double>(), A32, nsamp, nlist, d, (double)cm, dmax);
        RET(post_launch("k_coarse_filter_error", nullptr));
        double measured = 0.0;
        CK(cudaMemcpyAsync(&measured, dmax, sizeof(double), cudaMemcpyDeviceToHost, ix->stream));
        CK(cudaStreamSynchronize(ix->stream));
        ix->coarse_mma_measured = measured;
        ix->coarse_mma_ok = measured == measured && measured <= coarse_coef_mma(d) / 4.0;
        if (getenv("MMIDX_VERBOSE"))
            fprintf(stderr, "[mmidx] tensor-core coarse filter: measured error %.3g ||q|| Cmax, modelled radius %.3g -> %s\n", measured,
                    coarse_coef_mma(d), ix->coarse_mma_ok ? "enabled" : "disabled (FFMA filter)");
    }
    ix->has_C = true;
    ix->fast_ready = false;
    ix->gen++;
    return MMIDX_OK;
}

extern "C" int mmidx_set_permutation(mmidx_t *ix, const int32_t *perm) {
    if (!ix) return fail(MMIDX_ERR_INVALID, "null argument");
    if (ix->p.type == MMIDX_LINEAR) return fail(MMIDX_ERR_INVALID, "Linear index takes no transformation");
    DeviceGuard g(ix->device);
    std::lock_guard<std::mutex> lk(ix->mu);
    ix->fast_ready = false;
    ix->gen++;
    ix->has_rot = false;  // TransformationType is one of None / RandomRotation / RandomPermutation (PQ.java:30-32)
    if (!perm) {
        ix->has_perm = false;
        return MMIDX_OK;
    }
    std::vector<char> seen(ix->p.d, 0);
    for (int i = 0; i < ix->p.d; ++i) {
        if (perm[i] < 0 || perm[i] >= ix->p.d || seen[perm[i]]) return fail(MMIDX_ERR_INVALID, "perm is not a permutation of 0..d-1");
        seen[perm[i]] = 1;
    }
    RET(ix->dperm.reserve(sizeof(int32_t) * (size_t)ix->p.d, 0, ix->stream));
    CK(cudaMemcpyAsync(ix->dperm.p, perm, sizeof(int32_t) * (size_t)ix->p.d, cudaMemcpyHostToDevice, ix->stream));
    CK(cudaStreamSynchronize(ix->stream));
    ix->has_perm = true;
    return MMIDX_OK;
}

extern "C" int mmidx_set_transform(mmidx_t *ix, int32_t kind, const int32_t *perm, const double *R) {
    if (!ix) return fail(MMIDX_ERR_INVALID, "null argument");
    if (ix->p.type == MMIDX_LINEAR) return fail(MMIDX_ERR_INVALID, "Linear index takes no transformation");
    if (kind == MMIDX_TRANSFORM_NONE) return mmidx_set_permutation(ix, nullptr);
    if (kind == MMIDX_TRANSFORM_PERMUTATION) {
        if (!perm) return fail(MMIDX_ERR_INVALID, "RandomPermutation needs perm[d]");
        return mmidx_set_permutation(ix, perm);
    }
    if (kind != MMIDX_TRANSFORM_ROTATION) return fail(MMIDX_ERR_INVALID, "unknown transformation %d", kind);
    if (!R) return fail(MMIDX_ERR_INVALID, "RandomRotation needs the matrix R[d][d] (EJML's generator is not reproduced)");
    DeviceGuard g(ix->device);
    std::lock_guard<std::mutex> lk(ix->mu);
    const size_t bytes = sizeof(double) * (size_t)ix->p.d * ix->p.d;
    RET(ix->dR.reserve(bytes, 0, ix->stream));
    CK(cudaMemcpyAsync(ix->dR.p, R, bytes, cudaMemcpyHostToDevice, ix->stream));
    CK(cudaStreamSynchronize(ix->stream));
    ix->has_perm = false;
    ix->has_rot = true;
    ix->fast_ready = false;
    ix->gen++;
    return MMIDX_OK;
}

extern "C" int mmidx_set_w(mmidx_t *ix, int32_t w) {
    if (!ix) return fail(MMIDX_ERR_INVALID, "null argument");
    if (ix->p.type != MMIDX_IVFPQ) return fail(MMIDX_ERR_INVALID, "setW applies to IVFPQ only");
    ix->gen++;
    ix->p.w = w;  // validated at search time, like the reference (IVFPQ.java:95-97 stores it blindly)
    return MMIDX_OK;
}

static inline bool owns_list(const mmidx_index *ix, int32_t l) {
    if (ix->shard_count <= 1) return true;
    return (ix->shard_map.empty() ? l % ix->shard_count : ix->shard_map[l]) == ix->shard_rank;
}

extern "C" int mmidx_set_shard_map(mmidx_t *ix, const int32_t *owner) {
    if (!ix) return fail(MMIDX_ERR_INVALID, "null argument");
    if (ix->p.type != MMIDX_IVFPQ) return fail(MMIDX_ERR_INVALID, "only IVFPQ is sharded by list");
    std::lock_guard<std::mutex> lk(ix->mu);
    if (ix->n != 0) return fail(MMIDX_ERR_STATE, "the shard map must be set before any vector is indexed");
    if (!owner) {
        ix->shard_map.clear();
        return MMIDX_OK;
    }
    for (int l = 0; l < ix->p.nlist; ++l)
        if (owner[l] < 0 || owner[l] >= ix->shard_count) return fail(MMIDX_ERR_INVALID, "owner[%d] = %d out of range", l, owner[l]);
    ix->shard_map.assign(owner, owner + ix->p.nlist);
    return MMIDX_OK;
}

extern "C" int mmidx_enable_timings(mmidx_t *ix, int32_t on) {
    if (!ix) return fail(MMIDX_ERR_INVALID, "null argument");
    ix->timer.enabled = on != 0;
    return MMIDX_OK;
}

// ---------------------------------------------------------------------------------------------------------
// encode (K1c + K6)
// ---------------------------------------------------------------------------------------------------------
static int launch_assign(const double *dA, const double *dBt, int64_t na, int nb, int d, int32_t *dout, cudaStream_t st,
                         int *launches) {
    if (na == 0) return MMIDX_OK;
    k_assign_nearest<<<(unsigned)((na + QT - 1) / QT), MMIDX_NT, 0, st>>>(dA, dBt, na, nb, d, dout);
    return post_launch("k_assign_nearest", launches);
}

static bool env_exact() {
    static const bool v = [] {
        const char *e = getenv("MMIDX_MODE");
        return e && strcmp(e, "exact") == 0;
    }();
    return v;
}

// IVFPQ.computeNearestCoarseIndex (IVFPQ.java:547-564) for n vectors: filter matrix on the tensor cores (or FFMA), row minimum
// with the filter's error radius, binary64 only for the vectors the filter cannot decide (argmin_filter.cuh)
static int coarse_assign_dev(mmidx_index *ix, const double *dX, int64_t n, int32_t *dlist, cudaStream_t st, int *launches, Scratch &sc) {
    const int nlist = ix->p.nlist, d = ix->p.d;
    if (ix->force_exact || !ix->coarse_range_ok)
        return launch_assign(dX, ix->dCt.as<double>(), n, nlist, d, dlist, st, launches);
    const int64_t QC = std::max<int64_t>(1, (int64_t)(((size_t)512 << 20) / ((size_t)nlist * sizeof(float))));
    const int64_t cb = std::min(QC, n);
    float *A32, *amb_thr;
    int64_t *amb_list;
    int32_t *amb_count;
    RET(sc.get(&A32, (size_t)cb * nlist));
    RET(sc.get(&amb_list, (size_t)cb));
    RET(sc.get(&amb_thr, (size_t)cb));
    RET(sc.get(&amb_count, 1));
    unsigned short *Xh = nullptr, *Xl = nullptr;
    if (ix->coarse_mma_ok) {
        RET(sc.get(&Xh, (size_t)cb * ix->dpad));
        RET(sc.get(&Xl, (size_t)cb * ix->dpad));
    }
    for (int64_t q0 = 0; q0 < n; q0 += cb) {
        const int64_t nb = std::min(cb, n - q0);
        const double *xq = dX + q0 * d;
        double coef;
        if (ix->coarse_mma_ok) {
            RET(launch_coarse_mma(ix, xq, nb, A32, Xh, Xl, st, launches));
            coef = coarse_coef_mma(d);
        } else {
            dim3 gg((unsigned)((nlist + CG_BN - 1) / CG_BN), (unsigned)((nb + CG_BM - 1) / CG_BM));
            k_coarse_f32<<<gg, MMIDX_NT, 0, st>>>(xq, ix->dC32.as<float>(), ix->dc2.as<float>(), nb, nlist, d, A32);
            RET(post_launch("k_coarse_f32", launches));
            coef = coarse_coef_ffma(d);
        }
        CK(cudaMemsetAsync(amb_count, 0, sizeof(int32_t), st));
        k_rowmin_filter<<<(unsigned)((nb * 32 + MMIDX_NT - 1) / MMIDX_NT), MMIDX_NT, 0, st>>>(A32, xq, nb, nlist, d, ix->dcmax.as<float>(), coef,
                                                                                                dlist + q0, amb_list, amb_thr, amb_count);
        RET(post_launch("k_rowmin_filter", launches));
        k_rowmin_exact_list<<<2 * ix->sm_count, MMIDX_NT, 0, st>>>(A32, xq, ix->dC.as<double>(), nlist, d, amb_list, amb_thr, amb_count, dlist + q0);
        RET(post_launch("k_rowmin_exact_list", launches));
    }
    return MMIDX_OK;
}

static int launch_pq_encode(mmidx_index *ix, const double *dX, const int32_t *dlist, int64_t n, uint8_t *dout,
                            cudaStream_t st, int *launches, Scratch &sc) {
    if (n == 0) return MMIDX_OK;
    const int S = ix->S, m = ix->p.m, ks = ix->p.ks, d = ix->p.d;
    const double *C = dlist ? ix->dC.as<double>() : nullptr;
    const int32_t *perm = ix->has_perm ? ix->dperm.as<int32_t>() : nullptr;
    if (ix->has_rot) {
        // RandomRotation: rotate the vector (PQ.java:237-241) / the residual (IVFPQ.java:316-323) first, then plain encode
        double *dV;
        RET(sc.get(&dV, (size_t)n * d));
        k_rotate_vectors<<<(unsigned)n, MMIDX_NT, sizeof(double) * (size_t)d, st>>>(dX, C, dlist, 1, ix->dR.as<double>(), d, dV);
        RET(post_launch("k_rotate_vectors", launches));
        dX = dV;
        C = nullptr;
        dlist = nullptr;
        perm = nullptr;
    }
    if (!ix->force_exact) {
        // fp32 filter per sub-quantizer; binary64 only for the (vector, sub-quantizer) pairs it cannot decide
        int64_t *amb_list;
        int32_t *amb_count;
        RET(sc.get(&amb_list, (size_t)n * m));
        RET(sc.get(&amb_count, 1));
        CK(cudaMemsetAsync(amb_count, 0, sizeof(int32_t), st));
        ArgminRows rows{dX, d, C, dlist, d, perm, 0};
        uint8_t *o8 = ks <= 256 ? dout : nullptr;
        uint16_t *o16 = ks <= 256 ? nullptr : reinterpret_cast<uint16_t *>(dout);
        dim3 gf((unsigned)((n + AF_TV - 1) / AF_TV), m);
        k_argmin_filter<<<gf, MMIDX_NT, 0, st>>>(rows, ix->dPf32.as<float>(), ix->dPb2.as<float>(), ix->dPbmax.as<float>(), n, ks, S, m, o8, o16,
                                                 nullptr, m, amb_list, amb_count);
        RET(post_launch("k_argmin_filter", launches));
        k_argmin_exact_list<<<2 * ix->sm_count, MMIDX_NT, 0, st>>>(rows, ix->dP.as<double>(), ks, S, m, amb_list, amb_count, o8, o16, nullptr, m);
        return post_launch("k_argmin_exact_list", launches);
    }
    const double *P = ix->dP.as<double>();
    dim3 grid((unsigned)((n + MMIDX_NT - 1) / MMIDX_NT), m);
    int cchunk = std::min(ks, std::max(1, (32 * 1024) / (S * 8)));
    size_t smem = (size_t)cchunk * S * sizeof(double);
#define ENC(SV)                                                                                              \
    case SV:                                                                                                 \
        k_pq_encode<SV><<<grid, MMIDX_NT, smem, st>>>(dX, C, dlist, perm, P, n, d, m, ks, cchunk, dout, ix->code_bytes); \
        break;
    switch (S) {
        ENC(1) ENC(2) ENC(4) ENC(8) ENC(16) ENC(32)
        default:
            k_pq_encode_generic<<<grid, MMIDX_NT, 0, st>>>(dX, C, dlist, perm, P, n, d, m, ks, S, dout, ix->code_bytes);
    }
#undef ENC
    return post_launch("k_pq_encode", launches);
}

// encode n device-resident vectors: dlist (IVFPQ) and dcodes receive the assignment
static int encode_dev(mmidx_index *ix, const double *dX, int64_t n, int32_t *dlist, uint8_t *dcodes, cudaStream_t st,
                      int *launches, Scratch &sc) {
    if (ix->p.type == MMIDX_IVFPQ) {
        RET(coarse_assign_dev(ix, dX, n, dlist, st, launches, sc));
        RET(launch_pq_encode(ix, dX, dlist, n, dcodes, st, launches, sc));
    } else {
        RET(launch_pq_encode(ix, dX, nullptr, n, dcodes, st, launches, sc));
    }
    return MMIDX_OK;
}

static int require_quantizers(mmidx_index *ix) {
    if (ix->p.type != MMIDX_LINEAR && !ix->has_P) return fail(MMIDX_ERR_STATE, "product quantizer not loaded");
    if (ix->p.type == MMIDX_IVFPQ && !ix->has_C) return fail(MMIDX_ERR_STATE, "coarse quantizer not loaded");
    return MMIDX_OK;
}

static const int64_t ADD_BATCH = 1 << 16;

// a failed add leaves the index as it was: the host log is truncated back to what the device log holds
struct AddRollback {
    mmidx_index *ix;
    size_t hl;
    int64_t nl;
    bool ok = false;
    explicit AddRollback(mmidx_index *i) : ix(i), hl(i->h_list.size()), nl(i->n_local) {}
    ~AddRollback() {
        if (ok) return;
        ix->h_list.resize(hl);
        ix->h_iid.resize(hl);
        ix->n_local = nl;
    }
};

// dev: X, out_list and out_codes are DEVICE pointers (mmidx_add_dev); otherwise host pointers
static int add_or_encode(mmidx_index *ix, int64_t n, const double *X, int32_t *out_list, void *out_codes, bool store,
                         bool dev = false) {
    if (!ix || (n > 0 && !X)) return fail(MMIDX_ERR_INVALID, "null argument");
    if (n < 0) return fail(MMIDX_ERR_INVALID, "n < 0");
    RET(require_quantizers(ix));
    DeviceGuard g(ix->device);
    std::lock_guard<std::mutex> lk(ix->mu);
    AddRollback rollback(ix);
    const cudaMemcpyKind kin = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const cudaMemcpyKind kout = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    // ASS.java:232-235: indexVector returns false once loadCounter >= maxNumVectors
    if (store && ix->n + n > ix->p.max_n) return fail(MMIDX_ERR_FULL, "Maximum index capacity reached");
    cudaStream_t st = ix->stream;
    const int d = ix->p.d;
    int launches = 0;
    if (ix->p.type == MMIDX_LINEAR) {
        if (!store) return fail(MMIDX_ERR_INVALID, "Linear index has no codes to encode");
        int64_t total = ix->n + n;
        int64_t cap_vec = ((total + 31) / 32) * 32;
        RET(ix->dXb.reserve((size_t)cap_vec * d * sizeof(double), (size_t)(((ix->n + 31) / 32) * 32) * d * sizeof(double), st));
        Scratch sc(st);
        double *dX;
        RET(sc.get(&dX, (size_t)std::min(n, ADD_BATCH) * d));
        for (int64_t b = 0; b < n; b += ADD_BATCH) {
            int64_t nb = std::min(ADD_BATCH, n - b);
            const double *xb = X + b * d;
            if (!dev) {
                CK(cudaMemcpyAsync(dX, xb, sizeof(double) * (size_t)nb * d, kin, st));
                xb = dX;
            }
            k_linear_pack<<<(unsigned)((nb * d + 255) / 256), 256, 0, st>>>(xb, nb, d, ix->n + b, ix->dXb.as<double>());
            RET(post_launch("k_linear_pack", &launches));
            CK(cudaStreamSynchronize(st));
        }
        ix->n += n;
        ix->n_local = ix->n;
        ix->gen++;
        ix->last_launches = launches;
        rollback.ok = true;
        return MMIDX_OK;
    }
    const int cb = ix->code_bytes;
    Scratch sc(st);
    double *dX;
    int32_t *dlist = nullptr;
    uint8_t *dcodes;
    int64_t *dsel = nullptr;
    const int64_t bmax = std::min(std::max<int64_t>(n, 1), ADD_BATCH);
    RET(sc.get(&dX, (size_t)bmax * d));
    RET(sc.get(&dcodes, (size_t)bmax * cb));
    if (ix->p.type == MMIDX_IVFPQ) RET(sc.get(&dlist, (size_t)bmax));
    if (ix->shard_count > 1) RET(sc.get(&dsel, (size_t)bmax));
    std::vector<int32_t> hl;
    std::vector<int64_t> hsel;
    if (store) {
        // one growth step for the whole call (a shard keeps at most all n): no re-allocation inside the batch loop
        RET(ix->dcodes.reserve((size_t)(ix->n_local + n) * cb, (size_t)ix->n_local * cb, st));
        if (ix->p.type == MMIDX_IVFPQ) {
            ix->h_list.reserve(ix->h_list.size() + (size_t)n);
            ix->h_iid.reserve(ix->h_iid.size() + (size_t)n);
        }
    }
    for (int64_t b = 0; b < n; b += ADD_BATCH) {
        int64_t nb = std::min(ADD_BATCH, n - b);
        const double *xb = X + b * d;
        if (!dev) {
            CK(cudaMemcpyAsync(dX, xb, sizeof(double) * (size_t)nb * d, kin, st));
            xb = dX;
        }
        {
            Scratch sce(st);  // the rotated copy of a batch, when RandomRotation is set
            RET(encode_dev(ix, xb, nb, dlist, dcodes, st, &launches, sce));
        }
        if (out_codes) CK(cudaMemcpyAsync((uint8_t *)out_codes + b * cb, dcodes, (size_t)nb * cb, kout, st));
        if (ix->p.type == MMIDX_IVFPQ) {
            hl.resize(nb);
            CK(cudaMemcpyAsync(hl.data(), dlist, sizeof(int32_t) * (size_t)nb, cudaMemcpyDeviceToHost, st));
            if (out_list && dev) CK(cudaMemcpyAsync(out_list + b, dlist, sizeof(int32_t) * (size_t)nb, cudaMemcpyDeviceToDevice, st));
            CK(cudaStreamSynchronize(st));
            if (out_list && !dev) memcpy(out_list + b, hl.data(), sizeof(int32_t) * (size_t)nb);
        }
        if (store) {
            if (ix->p.type == MMIDX_PQ) {
                RET(ix->dcodes.reserve((size_t)(ix->n_local + nb) * cb, (size_t)ix->n_local * cb, st));
                CK(cudaMemcpyAsync(ix->dcodes.as<uint8_t>() + ix->n_local * cb, dcodes, (size_t)nb * cb,
                                   cudaMemcpyDeviceToDevice, st));
                ix->n_local += nb;
            } else if (ix->shard_count == 1) {
                RET(ix->dcodes.reserve((size_t)(ix->n_local + nb) * cb, (size_t)ix->n_local * cb, st));
                CK(cudaMemcpyAsync(ix->dcodes.as<uint8_t>() + ix->n_local * cb, dcodes, (size_t)nb * cb,
                                   cudaMemcpyDeviceToDevice, st));
                for (int64_t i = 0; i < nb; ++i) {
                    ix->h_list.push_back(hl[i]);
                    ix->h_iid.push_back((int32_t)(ix->n + b + i));
                }
                ix->n_local += nb;
            } else {
                // keep only the lists this shard owns (l % shard_count == shard_rank), iids stay global
                hsel.clear();
                for (int64_t i = 0; i < nb; ++i)
                    if (owns_list(ix, hl[i])) {
                        hsel.push_back(i);
                        ix->h_list.push_back(hl[i]);
                        ix->h_iid.push_back((int32_t)(ix->n + b + i));
                    }
                int64_t ns = (int64_t)hsel.size();
                if (ns) {
                    RET(ix->dcodes.reserve((size_t)(ix->n_local + ns) * cb, (size_t)ix->n_local * cb, st));
                    CK(cudaMemcpyAsync(dsel, hsel.data(), sizeof(int64_t) * (size_t)ns, cudaMemcpyHostToDevice, st));
                    k_gather_rows<<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(dcodes, dsel, ns, cb,
                                                                                 ix->dcodes.as<uint8_t>() + ix->n_local * cb);
                    RET(post_launch("k_gather_rows", &launches));
                    ix->n_local += ns;
                }
            }
            ix->sealed = false, ix->gen++;
        }
        CK(cudaStreamSynchronize(st));
    }
    if (store) ix->n += n;
    ix->last_launches = launches;
    rollback.ok = true;
    return MMIDX_OK;
}

extern "C" int mmidx_add_dev(mmidx_t *ix, int64_t n, const double *dX, int32_t *d_out_list, void *d_out_codes) {
    return add_or_encode(ix, n, dX, d_out_list, d_out_codes, true, true);
}

extern "C" int mmidx_add(mmidx_t *ix, int64_t n, const double *X, int32_t *out_list, void *out_codes) {
    return add_or_encode(ix, n, X, out_list, out_codes, true);
}

extern "C" int mmidx_encode(mmidx_t *ix, int64_t n, const double *X, int32_t *out_list, void *out_codes) {
    return add_or_encode(ix, n, X, out_list, out_codes, false);
}

extern "C" int mmidx_add_codes(mmidx_t *ix, int64_t n, const int32_t *list_ids, const void *codes) {
    if (!ix || (n > 0 && !codes)) return fail(MMIDX_ERR_INVALID, "null argument");
    if (ix->p.type == MMIDX_LINEAR) return fail(MMIDX_ERR_INVALID, "Linear index stores vectors, not codes");
    if (ix->p.type == MMIDX_IVFPQ && n > 0 && !list_ids) return fail(MMIDX_ERR_INVALID, "list_ids required for IVFPQ");
    DeviceGuard g(ix->device);
    std::lock_guard<std::mutex> lk(ix->mu);
    if (ix->n + n > ix->p.max_n) return fail(MMIDX_ERR_FULL, "Maximum index capacity reached");
    AddRollback rollback(ix);
    const int cb = ix->code_bytes;
    cudaStream_t st = ix->stream;
    // validate: code values < ks, list ids in range
    if (ix->p.ks <= 256) {
        if (ix->p.ks < 256) {
            const uint8_t *c = (const uint8_t *)codes;
            for (int64_t i = 0; i < n * ix->p.m; ++i)
                if (c[i] >= ix->p.ks) return fail(MMIDX_ERR_INVALID, "code value %d >= ks", (int)c[i]);
        }
    } else {
        const uint16_t *c = (const uint16_t *)codes;
        for (int64_t i = 0; i < n * ix->p.m; ++i)
            if (c[i] >= ix->p.ks) return fail(MMIDX_ERR_INVALID, "code value %d >= ks", (int)c[i]);
    }
    if (ix->p.type == MMIDX_PQ) {
        RET(ix->dcodes.reserve((size_t)(ix->n_local + n) * cb, (size_t)ix->n_local * cb, st));
        CK(cudaMemcpyAsync(ix->dcodes.as<uint8_t>() + ix->n_local * cb, codes, (size_t)n * cb, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        ix->n_local += n;
        ix->n += n;
        ix->sealed = false, ix->gen++;
        rollback.ok = true;
        return MMIDX_OK;
    }
    for (int64_t i = 0; i < n; ++i)
        if (list_ids[i] < 0 || list_ids[i] >= ix->p.nlist) return fail(MMIDX_ERR_INVALID, "list id %d out of range", list_ids[i]);
    std::vector<uint8_t> sel;
    const uint8_t *src = (const uint8_t *)codes;
    int64_t ns = n;
    if (ix->shard_count > 1) {
        sel.reserve((size_t)n * cb / ix->shard_count + 64);
        ns = 0;
        for (int64_t i = 0; i < n; ++i)
            if (owns_list(ix, list_ids[i])) {
                sel.insert(sel.end(), src + i * cb, src + (i + 1) * cb);
                ix->h_list.push_back(list_ids[i]);
                ix->h_iid.push_back((int32_t)(ix->n + i));
                ++ns;
            }
        src = sel.data();
    } else {
        for (int64_t i = 0; i < n; ++i) {
            ix->h_list.push_back(list_ids[i]);
            ix->h_iid.push_back((int32_t)(ix->n + i));
        }
    }
    if (ns) {
        RET(ix->dcodes.reserve((size_t)(ix->n_local + ns) * cb, (size_t)ix->n_local * cb, st));
        CK(cudaMemcpyAsync(ix->dcodes.as<uint8_t>() + ix->n_local * cb, src, (size_t)ns * cb, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
    }
    ix->n_local += ns;
    ix->n += n;
    ix->sealed = false, ix->gen++;
    rollback.ok = true;
    return MMIDX_OK;
}

// group the append log by list (stable => insertion order inside a list, IVFPQ.java:337-346)
static int seal(mmidx_index *ix) {
    if (ix->sealed) return MMIDX_OK;
    cudaStream_t st = ix->stream;
    const int nlist = ix->p.nlist, cb = ix->code_bytes;
    const int64_t nl = ix->n_local;
    std::fill(ix->h_list_len.begin(), ix->h_list_len.end(), 0);
    for (int64_t i = 0; i < nl; ++i) ix->h_list_len[ix->h_list[i]]++;
    int64_t pos = 0;
    ix->fast_len_ok = true;
    for (int l = 0; l < nlist; ++l) {
        if (ix->h_list_len[l] >= (1 << FAST_POS_BITS)) ix->fast_len_ok = false;
        ix->h_list_off[l] = pos;
        pos += (ix->h_list_len[l] + 15) & ~15;  // 16-entry aligned starts: 128-bit loads for any code width
    }
    const int64_t total = std::max<int64_t>(pos, 16);
    std::vector<int64_t> dst(nl), cur(ix->h_list_off);
    for (int64_t i = 0; i < nl; ++i) dst[i] = cur[ix->h_list[i]]++;
    RET(ix->csr_codes.reserve((size_t)total * cb, 0, st));
    RET(ix->csr_iids.reserve((size_t)total * sizeof(int32_t), 0, st));
    RET(ix->dlist_off.reserve(sizeof(int64_t) * (size_t)nlist, 0, st));
    RET(ix->dlist_len.reserve(sizeof(int32_t) * (size_t)nlist, 0, st));
    CK(cudaMemsetAsync(ix->csr_codes.p, 0, (size_t)total * cb, st));
    CK(cudaMemsetAsync(ix->csr_iids.p, 0xff, (size_t)total * sizeof(int32_t), st));
    CK(cudaMemcpyAsync(ix->dlist_off.p, ix->h_list_off.data(), sizeof(int64_t) * (size_t)nlist, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ix->dlist_len.p, ix->h_list_len.data(), sizeof(int32_t) * (size_t)nlist, cudaMemcpyHostToDevice, st));
    if (nl) {
        Scratch sc(st);
        int64_t *ddst;
        int32_t *diid;
        RET(sc.get(&ddst, (size_t)nl));
        RET(sc.get(&diid, (size_t)nl));
        CK(cudaMemcpyAsync(ddst, dst.data(), sizeof(int64_t) * (size_t)nl, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(diid, ix->h_iid.data(), sizeof(int32_t) * (size_t)nl, cudaMemcpyHostToDevice, st));
        k_scatter_codes<<<(unsigned)((nl + 255) / 256), 256, 0, st>>>(ix->dcodes.as<uint8_t>(), diid, ddst, nl, cb,
                                                                     ix->csr_codes.as<uint8_t>(), ix->csr_iids.as<int32_t>());
        RET(post_launch("k_scatter_codes", nullptr));
        CK(cudaStreamSynchronize(st));
    }
    CK(cudaStreamSynchronize(st));
    if (fast_eligible(ix)) {
        // second copy of the lists in a bank-conflict-aware order (fast_scan.cuh); offer order kept in orank
        const int m = ix->p.m;
        RET(ix->csr_ocodes.reserve((size_t)total * cb, 0, st));
        RET(ix->csr_oiids.reserve((size_t)total * sizeof(int32_t), 0, st));
        RET(ix->csr_orank.reserve((size_t)total * sizeof(int32_t), 0, st));
        CK(cudaMemsetAsync(ix->csr_ocodes.p, 0, (size_t)total * cb, st));
        CK(cudaMemsetAsync(ix->csr_oiids.p, 0xff, (size_t)total * sizeof(int32_t), st));
        Scratch sc(st);
        int32_t *src;
        RET(sc.get(&src, (size_t)total));
        if (ix->reorder) {
            const size_t rsm = (size_t)RCH * m;  // the pool's codes
            if (m == 8) {
                RET(set_smem(k_reorder_lists<8>, rsm));
                k_reorder_lists<8><<<nlist, MMIDX_NT, rsm, st>>>(ix->csr_codes.as<uint8_t>(), ix->dlist_off.as<int64_t>(),
                                                                ix->dlist_len.as<int32_t>(), src);
            } else {
                RET(set_smem(k_reorder_lists<16>, rsm));
                k_reorder_lists<16><<<nlist, MMIDX_NT, rsm, st>>>(ix->csr_codes.as<uint8_t>(), ix->dlist_off.as<int64_t>(),
                                                                 ix->dlist_len.as<int32_t>(), src);
            }
            RET(post_launch("k_reorder_lists", nullptr));
        } else {
            k_identity_order<<<nlist, MMIDX_NT, 0, st>>>(ix->dlist_off.as<int64_t>(), ix->dlist_len.as<int32_t>(), src);
            RET(post_launch("k_identity_order", nullptr));
        }
        k_apply_order<<<nlist, MMIDX_NT, 0, st>>>(ix->csr_codes.as<uint8_t>(), ix->csr_iids.as<int32_t>(),
                                                 ix->dlist_off.as<int64_t>(), ix->dlist_len.as<int32_t>(), src, m,
                                                 ix->csr_ocodes.as<uint8_t>(), ix->csr_oiids.as<int32_t>(),
                                                 ix->csr_orank.as<int32_t>());
        RET(post_launch("k_apply_order", nullptr));
        CK(cudaStreamSynchronize(st));
    }
    ix->long_lists = false;
    for (int32_t len : ix->h_list_len)
        if (len > FAST_SEG) ix->long_lists = true;
    ix->sealed = true;
    return MMIDX_OK;
}

// ---------------------------------------------------------------------------------------------------------
// search
// ---------------------------------------------------------------------------------------------------------
static inline int cap_for(int k) { return k <= 512 ? 1024 : 2048; }

template <int CAP>
static size_t topk_bytes() {
    return (sizeof(TopK<CAP>) + 127) & ~(size_t)127;
}

struct StageMark {
    mmidx_index *ix;
    cudaStream_t st;
    cudaEvent_t a = nullptr;
    int stage;
    StageMark(mmidx_index *i, cudaStream_t s, int stg) : ix(i), st(s), stage(stg) {
        if (ix->timer.enabled && stg >= 0) {
            a = ix->timer.get();
            ix->timer.record(a, st);
        }
    }
    void end() {
        if (a) {
            cudaEvent_t b = ix->timer.get();
            ix->timer.record(b, st);
            ix->timer.spans.push_back({a, b, stage});
            a = nullptr;
        }
    }
    ~StageMark() { end(); }
};

// coarse stage for a chunk: D (scratch) and probes[nq][w] in queue order
static int coarse_probe_dev(mmidx_index *ix, const double *dQ, int64_t nq, int w, int32_t *dprobes, Scratch &sc,
                            cudaStream_t st, int *launches, const PeerSink *psink = nullptr) {
    const int nlist = ix->p.nlist, d = ix->p.d;
    double *pd;
    int32_t *pcnt, *amb_list, *amb_count;
    RET(sc.get(&pd, (size_t)nq * w));
    RET(sc.get(&pcnt, (size_t)nq));
    RET(sc.get(&amb_list, (size_t)nq));
    RET(sc.get(&amb_count, 1));
    CK(cudaMemsetAsync(amb_count, 0, sizeof(int32_t), st));
    TopkOut o{};
    o.iids = dprobes;
    o.dist = pd;
    o.seq = nullptr;
    o.cnt = pcnt;
    o.tie = nullptr;
    o.amb_list = amb_list;
    o.amb_count = amb_count;
    o.nparts = 1;
    if (psink) o.sink = *psink;  // multi-GPU: the probe rows also land in the windows of the group peers
    // verification collector: the survivors are w plus the few centroids inside the error band; a small collector keeps
    // the CTA's shared memory low (8 CTAs/SM at w <= 128).  More survivors than ccap -> the kernel's exact sweep.
    const int ccap = w <= 128 ? 256 : cap_for(w);
    const size_t tkb = ccap == 256 ? topk_bytes<256>() : (ccap == 1024 ? topk_bytes<1024>() : topk_bytes<2048>());
    // survivors evaluated per batch: up to 16 / 32 rows of squared terms, at most 32 / 64 KB (at least one row)
    // (measured, profiles/README.md: 16 rows are best for w = 32, 32 rows for w = 64 -- about w / 2 survivors per batch)
    const size_t vcap = MMIDX_VERIFY_VB ? (size_t)MMIDX_VERIFY_VB : (w > 32 ? 32 : 16);
    const int vb = (int)std::max<size_t>(1, std::min<size_t>(vcap, (2 * vcap << 10) / ((size_t)(d + 1) * 8)));
    // filter keys in registers (k_coarse_verify<256, KPT>): no key array in shared memory
    const int kpt = (MMIDX_VERIFY_REGKEYS && ccap == 256 && nlist <= 32 * MMIDX_NT) ? (nlist <= 4 * MMIDX_NT ? 4 : (nlist <= 16 * MMIDX_NT ? 16 : 32)) : 0;
    const size_t vsm = tkb + (size_t)d * 8 + (kpt ? 0 : (((size_t)nlist * 4 + 7) & ~(size_t)7)) + (size_t)ccap * 4 + (size_t)vb * (d + 1) * 8;
    const bool fastc = !ix->force_exact && ix->coarse_range_ok && vsm <= 160 * 1024;
    TieLists tl;
    RET(sc.get(&tl.seq, (size_t)nq * w));
    RET(sc.get(&tl.pay, (size_t)nq * w));
    RET(sc.get(&tl.eq, (size_t)nq * w));
    RET(sc.get(&tl.cnt, (size_t)nq));
    const int tg = (int)std::min<int64_t>(nq, 2 * ix->sm_count);
    if (fastc) {
        // fp32 filter (tiled FFMA GEMM) + exact binary64 verification of the few centroids inside the error band
        float *A32;
        RET(sc.get(&A32, (size_t)nq * nlist));
        double coef;
        if (ix->coarse_mma_ok) {
            unsigned short *Xh, *Xl;
            RET(sc.get(&Xh, (size_t)nq * ix->dpad));
            RET(sc.get(&Xl, (size_t)nq * ix->dpad));
            RET(launch_coarse_mma(ix, dQ, nq, A32, Xh, Xl, st, launches));
            coef = coarse_coef_mma(d);
        } else {
            dim3 gg((unsigned)((nlist + CG_BN - 1) / CG_BN), (unsigned)((nq + CG_BM - 1) / CG_BM));
            k_coarse_f32<<<gg, MMIDX_NT, 0, st>>>(dQ, ix->dC32.as<float>(), ix->dc2.as<float>(), nq, nlist, d, A32);
            RET(post_launch("k_coarse_f32", launches));
            coef = coarse_coef_ffma(d);
        }
#define MMIDX_LAUNCH_VERIFY(CAPV, KPTV)                                                                                             \
    do {                                                                                                                            \
        RET(set_smem(k_coarse_verify<CAPV, KPTV>, vsm));                                                                            \
        k_coarse_verify<CAPV, KPTV><<<(unsigned)nq, MMIDX_NT, vsm, st>>>(dQ, ix->dC.as<double>(), A32, ix->dcmax.as<float>(), nlist, d, \
                                                                         w, vb, coef, o);                                           \
    } while (0)
        if (ccap == 256 && kpt == 4)
            MMIDX_LAUNCH_VERIFY(256, 4);
        else if (ccap == 256 && kpt == 16)
            MMIDX_LAUNCH_VERIFY(256, 16);
        else if (ccap == 256 && kpt == 32)
            MMIDX_LAUNCH_VERIFY(256, 32);
        else if (ccap == 256)
            MMIDX_LAUNCH_VERIFY(256, 0);
        else if (ccap == 1024)
            MMIDX_LAUNCH_VERIFY(1024, 0);
        else
            MMIDX_LAUNCH_VERIFY(2048, 0);
#undef MMIDX_LAUNCH_VERIFY
        RET(post_launch("k_coarse_verify", launches));
        // only rows ranked by the kernel's exact sweep (band wider than the collector) can be flagged here
        k_tie_collect_rows_direct<<<tg, MMIDX_NT, 0, st>>>(dQ, ix->dC.as<double>(), nlist, d, w, pd, amb_list, amb_count, tl);
        RET(post_launch("k_tie_collect_rows_direct", launches));
    } else {
        double *D;
        RET(sc.get(&D, (size_t)nq * nlist));
        dim3 g1((unsigned)((nlist + MMIDX_NT - 1) / MMIDX_NT), (unsigned)((nq + QT - 1) / QT));
        k_sqdist_matrix<<<g1, MMIDX_NT, 0, st>>>(dQ, ix->dCt.as<double>(), nq, nlist, d, D);
        RET(post_launch("k_sqdist_matrix", launches));
        if (cap_for(w) == 1024) {
            size_t smem = topk_bytes<1024>();
            RET(set_smem(k_select_rows<1024>, smem));
            k_select_rows<1024><<<(unsigned)nq, MMIDX_NT, smem, st>>>(D, nlist, w, o);
        } else {
            size_t smem = topk_bytes<2048>();
            RET(set_smem(k_select_rows<2048>, smem));
            k_select_rows<2048><<<(unsigned)nq, MMIDX_NT, smem, st>>>(D, nlist, w, o);
        }
        RET(post_launch("k_select_rows", launches));
        // ordered tie pass for rows whose w-th boundary was an exact tie
        k_tie_collect_rows<<<tg, MMIDX_NT, 0, st>>>(D, nlist, w, pd, amb_list, amb_count, tl);
        RET(post_launch("k_tie_collect_rows", launches));
    }
    size_t fs = (size_t)w * 16;
    RET(set_smem(k_tie_finish, std::max<size_t>(fs, 1024 * 16)));
    k_tie_finish<<<tg, MMIDX_NT, fs, st>>>(1, nq, w, tl.seq, tl.pay, tl.eq, tl.cnt, amb_list, amb_count, dprobes, pd, nullptr,
                                           psink ? *psink : PeerSink{});
    RET(post_launch("k_tie_finish", launches));
    return MMIDX_OK;
}

static inline int64_t lut_stride_of(const mmidx_index *ix) { return ((int64_t)ix->p.m * ix->p.ks + 1) & ~(int64_t)1; }

static int launch_lut(mmidx_index *ix, const double *dQ, const int32_t *dprobes, int64_t npairs, int w, double *dlut,
                      cudaStream_t st, int *launches, Scratch &sc, bool use_transform = true) {
    if (npairs == 0) return MMIDX_OK;
    const int S = ix->S, m = ix->p.m, ks = ix->p.ks, d = ix->p.d;
    const int32_t *perm = (use_transform && ix->has_perm) ? ix->dperm.as<int32_t>() : nullptr;
    const double *C = dprobes ? ix->dC.as<double>() : nullptr;
    if (use_transform && ix->has_rot) {
        // RandomRotation of the query (PQ.java:294-298) / of every probe's residual (IVFPQ.java:417-424), then plain tables
        double *dV;
        RET(sc.get(&dV, (size_t)npairs * d));
        k_rotate_vectors<<<(unsigned)npairs, MMIDX_NT, sizeof(double) * (size_t)d, st>>>(dQ, C, dprobes, w, ix->dR.as<double>(), d, dV);
        RET(post_launch("k_rotate_vectors", launches));
        dQ = dV;
        C = nullptr;
        dprobes = nullptr;
        w = 1;
    }
    dim3 grid((unsigned)((npairs + LUT_PT - 1) / LUT_PT), m);
    size_t smem = (size_t)LUT_PT * S * sizeof(double);
#define LUTK(SV)                                                                                                       \
    case SV:                                                                                                           \
        k_lut_build<SV><<<grid, MMIDX_NT, smem, st>>>(dQ, C, dprobes, perm, ix->dP.as<double>(), npairs, w, d, m, ks, S, lut_stride_of(ix), dlut); \
        break;
    switch (S) {
        LUTK(2) LUTK(4) LUTK(8) LUTK(16) LUTK(32)
        default:
            if (smem > 48 * 1024) RET(set_smem(k_lut_build<0>, smem));
            k_lut_build<0><<<grid, MMIDX_NT, smem, st>>>(dQ, C, dprobes, perm, ix->dP.as<double>(), npairs, w, d, m, ks, S, lut_stride_of(ix), dlut);
    }
#undef LUTK
    return post_launch("k_lut_build", launches);
}

struct ResultBufs {
    int32_t *iids;
    double *dist;
    unsigned long long *seq;  // may be NULL
    int32_t *cnt;
    PeerSink sink{};  // multi-GPU: where the final rows go besides / instead of the arrays above (kernels.cuh)
};

// merge [nq][nparts][k] partial results into res, flagging ambiguous queries
template <int CAP>
static int launch_merge(const TopkOut &part, int nparts, int64_t nq, int k, const ResultBufs &res, int32_t *amb_list,
                        int32_t *amb_count, double *out_tie, int64_t part_stride, int64_t q_stride, cudaStream_t st,
                        int *launches) {
    MergeArgs a{};
    a.iids = part.iids;
    a.dist = part.dist;
    a.seq = part.seq;
    a.cnt = part.cnt;
    a.tie = part.tie;
    a.part_stride = part_stride;
    a.q_stride = q_stride;
    a.nparts = nparts;
    a.k = k;
    TopkOut o{};
    o.iids = res.iids;
    o.dist = res.dist;
    o.seq = res.seq;
    o.cnt = res.cnt;
        o.sink = res.sink;
    o.tie = out_tie;
    o.amb_list = amb_list;
    o.amb_count = amb_count;
    o.nparts = 1;
    size_t smem = topk_bytes<CAP>() + (size_t)CAP * sizeof(int);  // + scratch of the in-kernel tie rule
    RET(set_smem(k_merge_topk<CAP>, smem));
    k_merge_topk<CAP><<<(unsigned)nq, MMIDX_NT, smem, st>>>(a, o);
    return post_launch("k_merge_topk", launches);
}

static int smem_limit_for_luts() { return 200 * 1024; }

// one chunk of IVFPQ queries.  want_seq: keep offer sequence numbers (sharded search).
// local_only_ties: true on a sharded index (ties are resolved after the cross-shard merge instead).
template <int CAP>
static int ivfpq_chunk(mmidx_index *ix, const double *dQ, int64_t nq, int k, int w, const ResultBufs &res,
                       double *res_tie, int32_t *amb_list, int32_t *amb_count, bool resolve_ties, cudaStream_t st,
                       int *launches, const int32_t *given_probes) {
    Scratch sc(st);
    const int m = ix->p.m, ks = ix->p.ks;
    int32_t *dprobes = const_cast<int32_t *>(given_probes);
    double *dlut;
    RET(sc.get(&dlut, (size_t)nq * w * lut_stride_of(ix)));
    if (!dprobes) {
        RET(sc.get(&dprobes, (size_t)nq * w));
        StageMark sm(ix, st, 0);
        RET(coarse_probe_dev(ix, dQ, nq, w, dprobes, sc, st, launches));
    }
    {
        StageMark sm(ix, st, 1);
        RET(launch_lut(ix, dQ, dprobes, nq * w, w, dlut, st, launches, sc));
    }
    IvfScanArgs a{};
    a.probes = dprobes;
    a.luts = dlut;
    a.lut_stride = lut_stride_of(ix);
    a.codes = ix->csr_codes.as<uint8_t>();
    a.iids = ix->csr_iids.as<int32_t>();
    a.list_off = ix->dlist_off.as<int64_t>();
    a.list_len = ix->dlist_len.as<int32_t>();
    a.w = w;
    a.k = k;
    a.L.m = m;
    a.L.ks = ks;
    a.L.code_bytes = ix->code_bytes;
    // enough CTAs to fill the machine: split a query's probes over `nsplit` CTAs when the chunk is small
    int nsplit = (int)std::min<int64_t>(w, std::max<int64_t>(1, (cta_slots(ix) + nq - 1) / nq));
    a.nsplit = nsplit;
    const size_t lut_bytes = (size_t)lut_stride_of(ix) * sizeof(double);
    const size_t smem = topk_bytes<CAP>() + 2 * lut_bytes + 64;
    if (smem > (size_t)smem_limit_for_luts())
        return fail(MMIDX_ERR_UNSUPPORTED, "m*ks = %d: ADC table (%zu bytes) does not fit shared memory", m * ks, lut_bytes);
    RET(set_smem(k_ivfpq_scan<CAP>, smem));
    TopkOut o{};
    o.nparts = nsplit;
    unsigned long long *pseq = nullptr;
    if (nsplit == 1) {
        o.iids = res.iids;
        o.dist = res.dist;
        o.seq = res.seq;
        o.cnt = res.cnt;
        o.sink = res.sink;
        o.tie = res_tie;
        o.amb_list = amb_list;
        o.amb_count = amb_count;
    } else {
        RET(sc.get(&o.iids, (size_t)nq * nsplit * k));
        RET(sc.get(&o.dist, (size_t)nq * nsplit * k));
        RET(sc.get(&pseq, (size_t)nq * nsplit * k));
        o.seq = pseq;
        RET(sc.get(&o.cnt, (size_t)nq * nsplit));
        RET(sc.get(&o.tie, (size_t)nq * nsplit));
        o.amb_list = nullptr;
        o.amb_count = nullptr;
    }
    {
        StageMark sm(ix, st, 2);
        k_ivfpq_scan<CAP><<<dim3(nsplit, (unsigned)nq), MMIDX_NT, smem, st>>>(a, o);
        RET(post_launch("k_ivfpq_scan", launches));
    }
    StageMark sm3(ix, st, 3);
    if (nsplit > 1)
        RET(launch_merge<CAP>(o, nsplit, nq, k, res, amb_list, amb_count, res_tie, 1, nsplit, st, launches));
    if (resolve_ties) {
        TieLists tl;
        RET(sc.get(&tl.seq, (size_t)nq * k));
        RET(sc.get(&tl.pay, (size_t)nq * k));
        RET(sc.get(&tl.eq, (size_t)nq * k));
        RET(sc.get(&tl.cnt, (size_t)nq));
        TieCodeArgs t{};
        t.luts = dlut;
    t.lut_stride = lut_stride_of(ix);
        t.codes = a.codes;
        t.iids = a.iids;
        t.probes = dprobes;
        t.list_off = a.list_off;
        t.list_len = a.list_len;
        t.w = w;
        t.m = m;
        t.ks = ks;
        t.code_bytes = ix->code_bytes;
        t.k = k;
        const int tg = (int)std::min<int64_t>(nq, 2 * ix->sm_count);
        k_tie_collect_ivfpq<<<tg, MMIDX_NT, 0, st>>>(t, res.dist, amb_list, amb_count, tl);
        RET(post_launch("k_tie_collect_ivfpq", launches));
        k_tie_finish<<<tg, MMIDX_NT, (size_t)k * 16, st>>>(1, nq, k, tl.seq, tl.pay, tl.eq, tl.cnt, amb_list, amb_count,
                                                           res.iids, res.dist, res.seq, res.sink);
        RET(post_launch("k_tie_finish", launches));
    }
    return MMIDX_OK;
}

template <int CAP>
static int pq_chunk(mmidx_index *ix, const double *dQ, int64_t nq, int k, const ResultBufs &res, int32_t *amb_list,
                    int32_t *amb_count, cudaStream_t st, int *launches) {
    Scratch sc(st);
    const int m = ix->p.m, ks = ix->p.ks;
    double *dlut;
    RET(sc.get(&dlut, (size_t)nq * lut_stride_of(ix)));
    {
        StageMark sm(ix, st, 1);
        RET(launch_lut(ix, dQ, nullptr, nq, 1, dlut, st, launches, sc));
    }
    constexpr int ROUND = TopK<CAP>::ROUND;
    const int64_t n = ix->n_local;
    int64_t rounds = std::max<int64_t>(1, (n + ROUND - 1) / ROUND);
    int64_t want = std::max<int64_t>(1, (cta_slots(ix) + nq - 1) / nq);
    int nsplit = (int)std::max<int64_t>(1, std::min<int64_t>(want, rounds));
    int64_t chunk = ((rounds + nsplit - 1) / nsplit) * ROUND;
    nsplit = (int)std::max<int64_t>(1, (n + chunk - 1) / chunk);
    PqScanArgs a{};
    a.luts = dlut;
    a.lut_stride = lut_stride_of(ix);
    a.codes = ix->dcodes.as<uint8_t>();
    a.n = n;
    a.chunk = chunk;
    a.k = k;
    a.L.m = m;
    a.L.ks = ks;
    a.L.code_bytes = ix->code_bytes;
    const size_t lut_bytes = (size_t)lut_stride_of(ix) * sizeof(double);
    const size_t smem = topk_bytes<CAP>() + lut_bytes + 64;
    if (smem > (size_t)smem_limit_for_luts())
        return fail(MMIDX_ERR_UNSUPPORTED, "m*ks = %d: ADC table (%zu bytes) does not fit shared memory", m * ks, lut_bytes);
    RET(set_smem(k_pq_scan<CAP>, smem));
    TopkOut o{};
    o.nparts = nsplit;
    if (nsplit == 1) {
        o.iids = res.iids;
        o.dist = res.dist;
        o.seq = res.seq;
        o.cnt = res.cnt;
        o.sink = res.sink;
        o.tie = nullptr;
        o.amb_list = amb_list;
        o.amb_count = amb_count;
    } else {
        unsigned long long *pseq;
        RET(sc.get(&o.iids, (size_t)nq * nsplit * k));
        RET(sc.get(&o.dist, (size_t)nq * nsplit * k));
        RET(sc.get(&pseq, (size_t)nq * nsplit * k));
        o.seq = pseq;
        RET(sc.get(&o.cnt, (size_t)nq * nsplit));
        RET(sc.get(&o.tie, (size_t)nq * nsplit));
    }
    {
        StageMark sm(ix, st, 2);
        k_pq_scan<CAP><<<dim3(nsplit, (unsigned)nq), MMIDX_NT, smem, st>>>(a, o);
        RET(post_launch("k_pq_scan", launches));
    }
    StageMark sm3(ix, st, 3);
    if (nsplit > 1)
        RET(launch_merge<CAP>(o, nsplit, nq, k, res, amb_list, amb_count, nullptr, 1, nsplit, st, launches));
    TieLists tl;
    RET(sc.get(&tl.seq, (size_t)nq * k));
    RET(sc.get(&tl.pay, (size_t)nq * k));
    RET(sc.get(&tl.eq, (size_t)nq * k));
    RET(sc.get(&tl.cnt, (size_t)nq));
    TieCodeArgs t{};
    t.luts = dlut;
    t.lut_stride = lut_stride_of(ix);
    t.codes = a.codes;
    t.n = n;
    t.w = 1;
    t.m = m;
    t.ks = ks;
    t.code_bytes = ix->code_bytes;
    t.k = k;
    const int tg = (int)std::min<int64_t>(nq, 2 * ix->sm_count);
    k_tie_collect_pq<<<tg, MMIDX_NT, 0, st>>>(t, res.dist, amb_list, amb_count, tl);
    RET(post_launch("k_tie_collect_pq", launches));
    k_tie_finish<<<tg, MMIDX_NT, (size_t)k * 16, st>>>(1, nq, k, tl.seq, tl.pay, tl.eq, tl.cnt, amb_list, amb_count, res.iids,
                                                       res.dist, res.seq, res.sink);
    RET(post_launch("k_tie_finish", launches));
    return MMIDX_OK;
}

template <int CAP>
static int linear_chunk(mmidx_index *ix, const double *dQ, int64_t nq, int k, const ResultBufs &res, int32_t *amb_list,
                        int32_t *amb_count, cudaStream_t st, int *launches) {
    Scratch sc(st);
    constexpr int ROUND = TopK<CAP>::ROUND;
    const int64_t n = ix->n_local;
    const int d = ix->p.d;
    int64_t rounds = std::max<int64_t>(1, (n + ROUND - 1) / ROUND);
    int64_t want = std::max<int64_t>(1, (cta_slots(ix) + nq - 1) / nq);
    int nsplit = (int)std::max<int64_t>(1, std::min<int64_t>(want, rounds));
    int64_t chunk = ((rounds + nsplit - 1) / nsplit) * ROUND;
    nsplit = (int)std::max<int64_t>(1, (n + chunk - 1) / chunk);
    LinearArgs a{};
    a.Q = dQ;
    a.Xb = ix->dXb.as<double>();
    a.n = n;
    a.chunk = chunk;
    a.d = d;
    a.k = k;
    const size_t smem = topk_bytes<CAP>() + (size_t)d * sizeof(double) + 64;
    RET(set_smem(k_linear_scan<CAP>, smem));
    TopkOut o{};
    o.nparts = nsplit;
    if (nsplit == 1) {
        o.iids = res.iids;
        o.dist = res.dist;
        o.seq = res.seq;
        o.cnt = res.cnt;
        o.sink = res.sink;
        o.tie = nullptr;
        o.amb_list = amb_list;
        o.amb_count = amb_count;
    } else {
        unsigned long long *pseq;
        RET(sc.get(&o.iids, (size_t)nq * nsplit * k));
        RET(sc.get(&o.dist, (size_t)nq * nsplit * k));
        RET(sc.get(&pseq, (size_t)nq * nsplit * k));
        o.seq = pseq;
        RET(sc.get(&o.cnt, (size_t)nq * nsplit));
        RET(sc.get(&o.tie, (size_t)nq * nsplit));
    }
    {
        StageMark sm(ix, st, 2);
        k_linear_scan<CAP><<<dim3(nsplit, (unsigned)nq), MMIDX_NT, smem, st>>>(a, o);
        RET(post_launch("k_linear_scan", launches));
    }
    StageMark sm3(ix, st, 3);
    if (nsplit > 1)
        RET(launch_merge<CAP>(o, nsplit, nq, k, res, amb_list, amb_count, nullptr, 1, nsplit, st, launches));
    TieLists tl;
    RET(sc.get(&tl.seq, (size_t)nq * k));
    RET(sc.get(&tl.pay, (size_t)nq * k));
    RET(sc.get(&tl.eq, (size_t)nq * k));
    RET(sc.get(&tl.cnt, (size_t)nq));
    const int tg = (int)std::min<int64_t>(nq, 2 * ix->sm_count);
    k_tie_collect_linear<<<tg, MMIDX_NT, 0, st>>>(dQ, a.Xb, n, d, k, res.dist, amb_list, amb_count, tl);
    RET(post_launch("k_tie_collect_linear", launches));
    k_tie_finish<<<tg, MMIDX_NT, (size_t)k * 16, st>>>(1, nq, k, tl.seq, tl.pay, tl.eq, tl.cnt, amb_list, amb_count, res.iids,
                                                       res.dist, res.seq, res.sink);
    RET(post_launch("k_tie_finish", launches));
    return MMIDX_OK;
}

// ---------------------------------------------------------------------------------------------------------
// fast path (fast_scan.cuh): fp32 filter + exact verification, no ADC tables in HBM
// ---------------------------------------------------------------------------------------------------------
static bool fast_eligible(const mmidx_index *ix) {
    if (ix->force_exact) return false;
    if (ix->has_rot) return false;  // a rotated residual goes through the binary64 ADC-table kernels
    // a flat PQ index runs the same kernels as an IVFPQ with ONE zero centroid and the negated codebook:
    // (0 - q) - (-P) = -(q - P) exactly, so every squared term has the bits of PQ.computeLookupADC (PQ.java:387-399)
    if (ix->p.type == MMIDX_PQ && ix->shard_count > 1) return false;
    if (ix->p.type != MMIDX_IVFPQ && ix->p.type != MMIDX_PQ) return false;
    // the fused kernel is specialised for byte codes with full 256-entry sub-tables and 8 or 16 sub-quantizers
    if (ix->p.ks != 256 || (ix->p.m != 8 && ix->p.m != 16)) return false;
    if (ix->p.d > 2048) return false;  // query + one survivor's squared terms must fit next to the collectors
    return true;
}

static int prepare_fast(mmidx_index *ix) {
    if (ix->fast_ready) return MMIDX_OK;
    cudaStream_t st = ix->stream;
    const bool flat = ix->p.type == MMIDX_PQ;
    const int m = ix->p.m, ks = ix->p.ks, S = ix->S, nlist = flat ? 1 : ix->p.nlist, d = ix->p.d;
    const double *P = ix->dP.as<double>();
    if (flat) {
        const int64_t np = (int64_t)m * ks * S;
        RET(ix->dPneg.reserve(sizeof(double) * (size_t)np, 0, st));
        RET(ix->dC.reserve(sizeof(double) * (size_t)d, 0, st));
        CK(cudaMemsetAsync(ix->dC.p, 0, sizeof(double) * (size_t)d, st));
        k_iota_negate<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(0, nullptr, np, P, ix->dPneg.as<double>());
        RET(post_launch("k_iota_negate", nullptr));
        P = ix->dPneg.as<double>();
    }
    RET(ix->dT1.reserve(sizeof(float) * (size_t)nlist * m * ks, 0, st));
    RET(ix->dP32t.reserve(sizeof(float) * (size_t)m * ks * S, 0, st));
    RET(ix->dt1max.reserve(sizeof(float) * (size_t)nlist * m, 0, st));
    RET(ix->dpmax.reserve(sizeof(float) * (size_t)m, 0, st));
    RET(ix->dstats.reserve(sizeof(unsigned long long) * 4, 0, st));
    CK(cudaMemsetAsync(ix->dt1max.p, 0, sizeof(float) * (size_t)nlist * m, st));
    CK(cudaMemsetAsync(ix->dpmax.p, 0, sizeof(float) * (size_t)m, st));
    CK(cudaMemsetAsync(ix->dstats.p, 0, sizeof(unsigned long long) * 4, st));
    const int32_t *perm = ix->has_perm ? ix->dperm.as<int32_t>() : nullptr;
    k_build_t1<<<dim3(nlist, m), MMIDX_NT, sizeof(double) * (size_t)S, st>>>(ix->dC.as<double>(), P, perm, d, m, ks, S,
                                                                          ix->dT1.as<float>(), ix->dt1max.as<float>());
    RET(post_launch("k_build_t1", nullptr));
    k_build_p32t<<<m, MMIDX_NT, 0, st>>>(P, m, ks, S, ix->dP32t.as<float>(), ix->dpmax.as<float>());
    RET(post_launch("k_build_p32t", nullptr));
    CK(cudaStreamSynchronize(st));
    {
        // magnitude window of the fp32 filter: max ||P_j,c||, max ||C_l|| and every T1 maximum (a squared norm)
        std::vector<float> pm((size_t)m), tm((size_t)nlist * m);
        CK(cudaMemcpy(pm.data(), ix->dpmax.p, sizeof(float) * pm.size(), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(tm.data(), ix->dt1max.p, sizeof(float) * tm.size(), cudaMemcpyDeviceToHost));
        bool ok = flat || ix->coarse_range_ok;
        for (float v : pm) ok = ok && fast_mag_ok((double)v);
        for (float v : tm) ok = ok && (v == 0.f || ((double)v >= FAST_MAG_MIN * FAST_MAG_MIN && (double)v <= 4.0 * FAST_MAG_MAX * FAST_MAG_MAX));
        ix->fast_range_ok = ok;
    }
    ix->fast_ready = true;
    return MMIDX_OK;
}

// Flat PQ index -> pseudo inverted lists of equal length L (iid order, so probe rank then list position is the
// reference's offer order, PQ.java:303-319) plus the bank-conflict-aware copy the fused kernel scans.
static int seal_pq_fast(mmidx_index *ix) {
    if (ix->sealed) return MMIDX_OK;
    cudaStream_t st = ix->stream;
    const int64_t n = ix->n_local;
    const int m = ix->p.m, cb = ix->code_bytes;
    const int64_t L = std::max<int64_t>(16384, (((n + MMIDX_MAX_K - 1) / MMIDX_MAX_K) + 15) & ~(int64_t)15);
    const int nl = (int)((n + L - 1) / L);
    ix->h_list_off.assign((size_t)nl + 1, 0);
    ix->h_list_len.assign((size_t)std::max(nl, 1), 0);
    for (int l = 0; l < nl; ++l) {
        ix->h_list_off[l] = (int64_t)l * L;
        ix->h_list_len[l] = (int32_t)std::min<int64_t>(L, n - (int64_t)l * L);
    }
    ix->h_list_off[nl] = n;
    ix->fast_len_ok = L < ((int64_t)1 << FAST_POS_BITS);
    ix->flat_nlist = nl;
    if (nl > 0 && ix->fast_len_ok) {
        RET(ix->dlist_off.reserve(sizeof(int64_t) * (size_t)(nl + 1), 0, st));
        RET(ix->dlist_len.reserve(sizeof(int32_t) * (size_t)nl, 0, st));
        CK(cudaMemcpyAsync(ix->dlist_off.p, ix->h_list_off.data(), sizeof(int64_t) * (size_t)(nl + 1), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ix->dlist_len.p, ix->h_list_len.data(), sizeof(int32_t) * (size_t)nl, cudaMemcpyHostToDevice, st));
        RET(ix->csr_iids.reserve(sizeof(int32_t) * (size_t)n, 0, st));
        k_iota_negate<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, ix->csr_iids.as<int32_t>(), 0, nullptr, nullptr);
        RET(post_launch("k_iota_negate", nullptr));
        RET(ix->csr_ocodes.reserve((size_t)n * cb, 0, st));
        RET(ix->csr_oiids.reserve((size_t)n * sizeof(int32_t), 0, st));
        RET(ix->csr_orank.reserve((size_t)n * sizeof(int32_t), 0, st));
        Scratch sc(st);
        int32_t *src;
        RET(sc.get(&src, (size_t)n));
        if (ix->reorder) {
            const size_t rsm = (size_t)RCH * m;
            if (m == 8) {
                RET(set_smem(k_reorder_lists<8>, rsm));
                k_reorder_lists<8><<<nl, MMIDX_NT, rsm, st>>>(ix->dcodes.as<uint8_t>(), ix->dlist_off.as<int64_t>(),
                                                             ix->dlist_len.as<int32_t>(), src);
            } else {
                RET(set_smem(k_reorder_lists<16>, rsm));
                k_reorder_lists<16><<<nl, MMIDX_NT, rsm, st>>>(ix->dcodes.as<uint8_t>(), ix->dlist_off.as<int64_t>(),
                                                              ix->dlist_len.as<int32_t>(), src);
            }
            RET(post_launch("k_reorder_lists", nullptr));
        } else {
            k_identity_order<<<nl, MMIDX_NT, 0, st>>>(ix->dlist_off.as<int64_t>(), ix->dlist_len.as<int32_t>(), src);
            RET(post_launch("k_identity_order", nullptr));
        }
        k_apply_order<<<nl, MMIDX_NT, 0, st>>>(ix->dcodes.as<uint8_t>(), ix->csr_iids.as<int32_t>(), ix->dlist_off.as<int64_t>(),
                                              ix->dlist_len.as<int32_t>(), src, m, ix->csr_ocodes.as<uint8_t>(),
                                              ix->csr_oiids.as<int32_t>(), ix->csr_orank.as<int32_t>());
        RET(post_launch("k_apply_order", nullptr));
        CK(cudaStreamSynchronize(st));
    }
    ix->long_lists = false;
    for (int32_t len : ix->h_list_len)
        if (len > FAST_SEG) ix->long_lists = true;
    ix->sealed = true;
    return MMIDX_OK;
}

// CTAs per query of the fused scan (each takes every nsplit-th probe): one, unless the batch cannot fill the 592
// resident CTA slots (148 SMs x 4).  Splitting larger batches to even out the last wave was measured and rejected: the
// partial results need k_merge_topk (0.36 ms for 2500 queries x 2 parts), more than the tail it saves.
// set while mmidx_search's pipelined chunks are being enqueued: they alternate between two streams, so the tail of one
// chunk's grid is filled by the next chunk's kernels and re-ordering the queries of a chunk only costs a kernel
static thread_local bool g_overlapped_chunks = false;

static int fast_nsplit(const mmidx_index *ix, int64_t nq, int w) {
    return (int)std::min<int64_t>(w, std::max<int64_t>(1, (cta_slots(ix) + nq - 1) / nq));
}

template <int CAP32, int M>
static size_t fast_smem_bytes(int ks, int S, int d) {
    constexpr int ECAP = FastExactCap<CAP32, M>::value;
    const size_t c32b = (sizeof(TopK32<CAP32>) + 127) & ~(size_t)127;
    const size_t tkb = (sizeof(TopK<ECAP>) + 127) & ~(size_t)127;
    const size_t finb = tkb + ECAP * sizeof(int) + (size_t)M * (S + 1) * sizeof(double);  // final phase of k_ivfpq_scan_fast
    size_t regA = std::max((size_t)3 * M * ks * sizeof(float), finb);
    regA = (regA + 15) & ~(size_t)15;
    return c32b + regA + (size_t)d * 8 + 16 + 64;
}

template <int CAP32, int M>
static int ivfpq_chunk_fast(mmidx_index *ix, const double *dQ, int64_t nq, int k, int w, const ResultBufs &res,
                            double *res_tie, int32_t *amb_list, int32_t *amb_count, bool resolve_ties, cudaStream_t st,
                            int *launches, const int32_t *given_probes) {
    Scratch sc(st, ix->arenas, ix->arena_mu);
    int32_t *dprobes = const_cast<int32_t *>(given_probes);
    if (!dprobes) {
        RET(sc.get(&dprobes, (size_t)nq * w));
        StageMark sm(ix, st, 0);
        RET(coarse_probe_dev(ix, dQ, nq, w, dprobes, sc, st, launches));
    }
    const int nsplit = fast_nsplit(ix, nq, w);
    FastArgs a{};
    a.Q = dQ;
    a.C = ix->dC.as<double>();
    const bool flat = ix->p.type == MMIDX_PQ;
    a.flat = flat ? 1 : 0;
    a.P = flat ? ix->dPneg.as<double>() : ix->dP.as<double>();
    a.T1 = ix->dT1.as<float>();
    a.P32t = ix->dP32t.as<float>();
    a.t1max = ix->dt1max.as<float>();
    a.pmax = ix->dpmax.as<float>();
    a.perm = ix->has_perm ? ix->dperm.as<int32_t>() : nullptr;
    a.probes = dprobes;
    a.codes = flat ? ix->dcodes.as<uint8_t>() : ix->csr_codes.as<uint8_t>();  // flat: the append log is the one list order
    a.iids = ix->csr_iids.as<int32_t>();
    a.ocodes = ix->csr_ocodes.as<uint8_t>();
    a.oiids = ix->csr_oiids.as<int32_t>();
    a.orank = ix->csr_orank.as<int32_t>();
    a.list_off = ix->dlist_off.as<int64_t>();
    a.list_len = ix->dlist_len.as<int32_t>();
    a.d = ix->p.d;
    a.m = M;
    a.ks = ix->p.ks;
    a.S = ix->S;
    a.w = w;
    a.k = k;
    a.nsplit = nsplit;
    a.resolve_ties = (resolve_ties && nsplit == 1) ? 1 : 0;
    a.stats = ix->want_stats ? ix->dstats.as<unsigned long long>() : nullptr;
    {
        // per-(query, probe) terms of the table decomposition and the per-query error radius
        unsigned char *desc;
        double *bq;
        int32_t *oprobes, *ocnt, *work = nullptr, *qorder = nullptr;
        float *T2;
        RET(sc.get(&T2, (size_t)nq * M * 256));
        RET(sc.get(&desc, (size_t)nq * w * fast_desc_stride(M)));
        RET(sc.get(&bq, (size_t)nq));
        RET(sc.get(&oprobes, (size_t)nq * w));
        RET(sc.get(&ocnt, (size_t)nq));
        // a batch of 1 .. 7 waves of CTAs is launched heaviest query first (the tail of the last wave gets short)
        const bool lpt = nsplit == 1 && nq > cta_slots(ix) && nq <= ORDER_MAX && !g_overlapped_chunks;
        if (lpt) {
            RET(sc.get(&work, (size_t)nq));
            RET(sc.get(&qorder, (size_t)nq));
        }
        StageMark sm(ix, st, 1);
        const size_t psm = (size_t)(a.d + M) * 8 + (size_t)M * 8 + (size_t)w * 8 + 16;
        if (M == 8 && a.S == 16)
            k_fast_prep<8, 16><<<(unsigned)nq, MMIDX_NT, psm, st>>>(dQ, a.C, a.perm, dprobes, a.t1max, a.pmax, a.list_off, a.list_len, a.d, M, a.S, w, a.flat, desc, bq, oprobes, ocnt, work);
        else if (M == 16 && a.S == 8)
            k_fast_prep<16, 8><<<(unsigned)nq, MMIDX_NT, psm, st>>>(dQ, a.C, a.perm, dprobes, a.t1max, a.pmax, a.list_off, a.list_len, a.d, M, a.S, w, a.flat, desc, bq, oprobes, ocnt, work);
        else
            k_fast_prep<0, 0><<<(unsigned)nq, MMIDX_NT, psm, st>>>(dQ, a.C, a.perm, dprobes, a.t1max, a.pmax, a.list_off, a.list_len, a.d, M, a.S, w, a.flat, desc, bq, oprobes, ocnt, work);
        RET(post_launch("k_fast_prep", launches));
        if (lpt) {
            k_order_queries<<<1, ORDER_NT, 0, st>>>((int)nq, work, qorder);
            RET(post_launch("k_order_queries", launches));
        }
        a.qorder = qorder;
        {
            dim3 g2((unsigned)((nq + T2_QB - 1) / T2_QB), M);
            const size_t sm2 = (size_t)T2_QB * a.S * sizeof(float);
            if (a.S == 16)
                k_fast_t2<16><<<g2, MMIDX_NT, sm2, st>>>(dQ, a.perm, a.P32t, nq, a.d, M, a.S, T2);
            else if (a.S == 8)
                k_fast_t2<8><<<g2, MMIDX_NT, sm2, st>>>(dQ, a.perm, a.P32t, nq, a.d, M, a.S, T2);
            else
                k_fast_t2<0><<<g2, MMIDX_NT, sm2, st>>>(dQ, a.perm, a.P32t, nq, a.d, M, a.S, T2);
            RET(post_launch("k_fast_t2", launches));
        }
        a.T2 = T2;
        a.desc = desc;
        a.bq = bq;
        a.oprobes = oprobes;
        a.ocnt = ocnt;
    }
    RET(sc.get(&a.fb_list, (size_t)nq * nsplit));
    RET(sc.get(&a.fb_count, 1));
    CK(cudaMemsetAsync(a.fb_count, 0, sizeof(int32_t), st));
    size_t smem = fast_smem_bytes<CAP32, M>(ix->p.ks, ix->S, ix->p.d);
    {
        // room for the query's descriptors in shared memory, if it does not cost a resident CTA (4 per SM by registers)
        const size_t per_sm = ix->smem_per_sm, blk = 1024;  // 1 KB per CTA is reserved by the driver
        const size_t ctas = std::min<size_t>(4, per_sm / (smem + blk));
        const size_t desc_bytes = (size_t)w * fast_desc_stride(M);
        a.desc_stage = (MMIDX_SCAN_STAGE && ctas > 0 && smem + desc_bytes + blk <= per_sm / ctas) ? 1 : 0;
        if (a.desc_stage) smem += desc_bytes;
    }
    // lists longer than FAST_SEG entries (pseudo lists of a large flat PQ index) take the segmented sweep
    auto scan_kernel = ix->long_lists ? k_ivfpq_scan_fast<CAP32, M, true> : k_ivfpq_scan_fast<CAP32, M, false>;
    RET(set_smem(scan_kernel, smem));
    TopkOut o{};
    o.nparts = nsplit;
    if (nsplit == 1) {
        o.iids = res.iids;
        o.dist = res.dist;
        o.seq = res.seq;
        o.cnt = res.cnt;
        o.sink = res.sink;
        o.tie = res_tie;
        o.amb_list = amb_list;
        o.amb_count = amb_count;
    } else {
        unsigned long long *pseq;
        RET(sc.get(&o.iids, (size_t)nq * nsplit * k));
        RET(sc.get(&o.dist, (size_t)nq * nsplit * k));
        RET(sc.get(&pseq, (size_t)nq * nsplit * k));
        o.seq = pseq;
        RET(sc.get(&o.cnt, (size_t)nq * nsplit));
        RET(sc.get(&o.tie, (size_t)nq * nsplit));
    }
    {
        StageMark sm(ix, st, 2);
        scan_kernel<<<dim3(nsplit, (unsigned)nq), MMIDX_NT, smem, st>>>(a, o);
        RET(post_launch("k_ivfpq_scan_fast", launches));
    }
    StageMark sm3(ix, st, 3);
    {
        // items whose error band overflowed the fp32 collector (massive exact duplicates): table-free exact scan
        constexpr int DCAP = 1024;  // k <= 256 on this path
        const size_t dsmem = topk_bytes<DCAP>();
        RET(set_smem(k_ivfpq_scan_direct<DCAP>, dsmem));
        const int dg = (int)std::min<int64_t>(nq * nsplit, 2 * ix->sm_count);
        k_ivfpq_scan_direct<DCAP><<<dg, MMIDX_NT, dsmem, st>>>(a, o);
        RET(post_launch("k_ivfpq_scan_direct", launches));
    }
    if (nsplit > 1) {
        if (cap_for(k) == 2048)
            RET(launch_merge<2048>(o, nsplit, nq, k, res, amb_list, amb_count, res_tie, 1, nsplit, st, launches));
        else
            RET(launch_merge<1024>(o, nsplit, nq, k, res, amb_list, amb_count, res_tie, 1, nsplit, st, launches));
    }
    if (resolve_ties) {
        TieLists tl;
        RET(sc.get(&tl.seq, (size_t)nq * k));
        RET(sc.get(&tl.pay, (size_t)nq * k));
        RET(sc.get(&tl.eq, (size_t)nq * k));
        RET(sc.get(&tl.cnt, (size_t)nq));
        TieDirectArgs t{};
        t.Q = dQ;
        t.C = a.C;
        t.P = a.P;
        t.perm = a.perm;
        t.probes = dprobes;
        t.codes = a.codes;
        t.iids = a.iids;
        t.list_off = a.list_off;
        t.list_len = a.list_len;
        t.d = a.d;
        t.m = M;
        t.ks = a.ks;
        t.S = a.S;
        t.w = w;
        t.k = k;
        t.code_bytes = ix->code_bytes;
        t.flat = a.flat;
        const int tg = (int)std::min<int64_t>(nq, 2 * ix->sm_count);
        k_tie_collect_ivfpq_direct<<<tg, MMIDX_NT, 0, st>>>(t, res.dist, amb_list, amb_count, tl);
        RET(post_launch("k_tie_collect_ivfpq_direct", launches));
        k_tie_finish<<<tg, MMIDX_NT, (size_t)k * 16, st>>>(1, nq, k, tl.seq, tl.pay, tl.eq, tl.cnt, amb_list, amb_count,
                                                           res.iids, res.dist, res.seq, res.sink);
        RET(post_launch("k_tie_finish", launches));
    }
    return MMIDX_OK;
}

static int ivfpq_chunk_fast_dispatch(mmidx_index *ix, const double *dQ, int64_t nq, int k, int w, const ResultBufs &res,
                                     double *res_tie, int32_t *amb_list, int32_t *amb_count, bool resolve_ties,
                                     cudaStream_t st, int *launches, const int32_t *given_probes) {
#define FASTCALL(CAPV, MV) \
    return ivfpq_chunk_fast<CAPV, MV>(ix, dQ, nq, k, w, res, res_tie, amb_list, amb_count, resolve_ties, st, launches, given_probes)
    // fp32 collector capacity: 1024 entries measured 10 % faster than 2048 at m = 8 (profiles/README.md); m = 16 stages
    // 16-byte survivor codes in the dead key array and needs the larger one for k up to 256
    if (ix->p.m == 8) FASTCALL(1024, 8);
    FASTCALL(2048, 16);
#undef FASTCALL
}

static int validate_search(mmidx_index *ix, int64_t nq, int k, int *w_out) {
    if (nq < 0) return fail(MMIDX_ERR_INVALID, "nq < 0");
    if (k < 1) return fail(MMIDX_ERR_INVALID, "k must be >= 1 (BoundedPriorityQueue max size)");
    if (k > MMIDX_MAX_K) return fail(MMIDX_ERR_UNSUPPORTED, "k = %d exceeds MMIDX_MAX_K = %d", k, MMIDX_MAX_K);
    RET(require_quantizers(ix));
    if (ix->p.type == MMIDX_IVFPQ) {
        int w = ix->p.w;
        if (w < 1) return fail(MMIDX_ERR_W, "w = %d: BoundedPriorityQueue needs a positive size", w);
        if (w > ix->p.nlist) return fail(MMIDX_ERR_W, "w = %d exceeds the number of coarse centroids %d", w, ix->p.nlist);
        if (w > MMIDX_MAX_K) return fail(MMIDX_ERR_UNSUPPORTED, "w = %d exceeds %d", w, MMIDX_MAX_K);
        *w_out = w;
    }
    return MMIDX_OK;
}

// device-side search over all chunks.  d_seq/d_tie non-NULL => sharded mode (no local tie resolution).
// Everything a search needs that is not per call: CSR lists sealed, fast-path tables built.  Host-synchronising, so it
// runs before (never inside) a CUDA-graph capture.
static int prepare_search(mmidx_index *ix, int k) {
    if (ix->p.type == MMIDX_IVFPQ && !ix->sealed) {
        std::lock_guard<std::mutex> lk(ix->mu);
        RET(seal(ix));
    }
    if (ix->p.type == MMIDX_PQ && fast_eligible(ix) && k <= 256 && !ix->sealed) {
        std::lock_guard<std::mutex> lk(ix->mu);
        RET(seal_pq_fast(ix));
    }
    const bool fast = fast_eligible(ix) && ix->fast_len_ok && k <= 256 && (ix->p.type != MMIDX_PQ || ix->flat_nlist > 0);
    if (fast && !ix->fast_ready) {
        std::lock_guard<std::mutex> lk(ix->mu);
        RET(prepare_fast(ix));  // also decides fast_range_ok
    }
    return MMIDX_OK;
}

static int search_dev_impl(mmidx_index *ix, int64_t nq, const double *dQ, int k, int32_t *d_iids, double *d_dist,
                           unsigned long long *d_seq, double *d_tie, int32_t *d_count, cudaStream_t st, bool sharded,
                           const int32_t *d_probes = nullptr, const PeerSink *sink = nullptr, bool reset_timer = true,
                           int *launches_out = nullptr) {
    int w = 0;
    RET(validate_search(ix, nq, k, &w));
    if (nq == 0) return MMIDX_OK;
    RET(prepare_search(ix, k));
    // otherwise: exact ADC-table kernels
    const bool fast = fast_eligible(ix) && ix->fast_len_ok && ix->fast_range_ok && k <= 256 &&
                      (ix->p.type != MMIDX_PQ || ix->flat_nlist > 0);
    if (ix->p.type == MMIDX_PQ && fast) w = ix->flat_nlist;  // every pseudo list is probed, in iid order
    int launches = 0;
    if (reset_timer) ix->timer.reset();
    StageMark whole(ix, st, reset_timer ? 4 : -1);  // inside a multi-GPU step the caller brackets the whole call
    Scratch sc(st, ix->arenas, ix->arena_mu);
    int32_t *amb_list, *amb_count;
    const int d = ix->p.d;
    int64_t qchunk = nq;
    if (ix->p.type == MMIDX_IVFPQ && fast) {
        // scratch is the coarse distance matrix only: [nq][nlist] binary64, bounded to 1 GiB
        qchunk = std::max<int64_t>(1, (int64_t)(((size_t)1 << 30) / ((size_t)ix->p.nlist * sizeof(double))));
    } else if (ix->p.type == MMIDX_PQ && fast) {
        qchunk = 32768;  // T2 (m KB) + w probe descriptors per query
    } else if (ix->p.type == MMIDX_IVFPQ) {
        size_t per_q = (size_t)w * ix->p.m * ix->p.ks * sizeof(double);
        qchunk = std::max<int64_t>(1, (int64_t)(ix->lut_chunk_bytes / per_q));
    } else if (ix->p.type == MMIDX_PQ) {
        size_t per_q = (size_t)ix->p.m * ix->p.ks * sizeof(double);
        qchunk = std::max<int64_t>(1, (int64_t)(((size_t)256 << 20) / per_q));
    } else {
        qchunk = 16384;
    }
    qchunk = std::min<int64_t>(qchunk, 32768);  // grid.y limit is 65535
    qchunk = std::min(qchunk, nq);
    RET(sc.get(&amb_list, (size_t)qchunk));
    RET(sc.get(&amb_count, 1));
    for (int64_t q0 = 0; q0 < nq; q0 += qchunk) {
        int64_t nb = std::min(qchunk, nq - q0);
        CK(cudaMemsetAsync(amb_count, 0, sizeof(int32_t), st));
        ResultBufs res{d_iids ? d_iids + q0 * k : nullptr, d_dist ? d_dist + q0 * k : nullptr,
                       d_seq ? d_seq + q0 * k : nullptr, d_count ? d_count + q0 : nullptr};
        if (sink) {
            res.sink = *sink;
            res.sink.q0 += q0;
        }
        const double *dq = dQ + q0 * d;
        int r;
        const bool big = cap_for(k) == 2048;
        switch (ix->p.type) {
            case MMIDX_IVFPQ:
            {
                const int32_t *gp = d_probes ? d_probes + q0 * w : nullptr;
                if (fast) {
                    r = ivfpq_chunk_fast_dispatch(ix, dq, nb, k, w, res, d_tie ? d_tie + q0 : nullptr, amb_list, amb_count,
                                                  !sharded, st, &launches, gp);
                    break;
                }
                r = big ? ivfpq_chunk<2048>(ix, dq, nb, k, w, res, d_tie ? d_tie + q0 : nullptr, amb_list, amb_count, !sharded, st, &launches, gp)
                        : ivfpq_chunk<1024>(ix, dq, nb, k, w, res, d_tie ? d_tie + q0 : nullptr, amb_list, amb_count, !sharded, st, &launches, gp);
            }
                break;
            case MMIDX_PQ:
                if (fast) {
                    int32_t *fp;  // probes[q][p] = p
                    r = sc.get(&fp, (size_t)nb * w);
                    if (r != MMIDX_OK) break;
                    k_fill_flat_probes<<<(unsigned)((nb * w + 255) / 256), 256, 0, st>>>(nb, w, fp);
                    r = post_launch("k_fill_flat_probes", &launches);
                    if (r != MMIDX_OK) break;
                    r = ivfpq_chunk_fast_dispatch(ix, dq, nb, k, w, res, nullptr, amb_list, amb_count, true, st, &launches, fp);
                    break;
                }
                r = big ? pq_chunk<2048>(ix, dq, nb, k, res, amb_list, amb_count, st, &launches)
                        : pq_chunk<1024>(ix, dq, nb, k, res, amb_list, amb_count, st, &launches);
                break;
            default:
                r = big ? linear_chunk<2048>(ix, dq, nb, k, res, amb_list, amb_count, st, &launches)
                        : linear_chunk<1024>(ix, dq, nb, k, res, amb_list, amb_count, st, &launches);
        }
        RET(r);
    }
    whole.end();
    ix->last_launches = launches;
    if (launches_out) *launches_out += launches;
    return MMIDX_OK;
}

// ---------------------------------------------------------------------------------------------------------
// CUDA graphs of whole search calls.  A call is ~12 kernel launches plus memsets; for small batches (one query, or
// the 1/8 slice of a batch on 8 GPUs) the host side of those launches is longer than the kernels.  The third call with
// the same shape and buffers is captured (the first two warm the scratch arena, whose addresses the graph holds) and
// later calls replay it with one cudaGraphLaunch.  Anything that changes the index or re-allocates the arena
// invalidates the captures.  parity: the multi-GPU step alternates between two window halves (comm.cuh).
// ---------------------------------------------------------------------------------------------------------
struct GraphKey {
    int kind;  // 0: mmidx_search_dev, 1: multi-GPU step
    int64_t nq;
    int k, w, flags;
    const void *p0, *p1, *p2, *p3;
    cudaStream_t st;
    bool operator<(const GraphKey &o) const {
        return std::tie(kind, nq, k, w, flags, p0, p1, p2, p3, st) < std::tie(o.kind, o.nq, o.k, o.w, o.flags, o.p0, o.p1, o.p2, o.p3, o.st);
    }
};
struct GraphEntry {
    cudaGraphExec_t exec[2] = {nullptr, nullptr};
    int seen = 0, launches = 0;
    bool disabled = false;
    uint64_t gen = 0, arena_gen = 0;
};

struct GraphCache {
    std::mutex mu;  // held for a whole call: calls on one stream are serialised anyway, calls on different streams are not
    std::map<GraphKey, GraphEntry> m;
};

static void clear_cache(GraphCache &c) {
    for (auto &kv : c.m)
        for (auto &e : kv.second.exec)
            if (e) cudaGraphExecDestroy(e);
    c.m.clear();
}

static void drop_graphs(mmidx_index *ix) {
    std::lock_guard<std::mutex> lk(ix->graph_mu);
    for (auto &kv : ix->graphs) {
        std::lock_guard<std::mutex> lk2(kv.second->mu);
        clear_cache(*kv.second);
    }
}

static void free_graph_caches(mmidx_index *ix) {
    drop_graphs(ix);
    std::lock_guard<std::mutex> lk(ix->graph_mu);
    for (auto &kv : ix->graphs) delete kv.second;
    ix->graphs.clear();
}

// The stream a host-pointer search call of this thread runs on.  One stream per calling thread: concurrent
// computeNearestNeighbors callers (the reference leaves search unsynchronised, ASS.java:281) do not serialise on one stream,
// and a thread's CUDA-graph capture never sees another thread's work on its stream.
static int host_stream(mmidx_index *ix, cudaStream_t *out) {
    std::lock_guard<std::mutex> lk(ix->ts_mu);
    auto it = ix->thread_streams.find(std::this_thread::get_id());
    if (it != ix->thread_streams.end()) {
        *out = it->second;
        return MMIDX_OK;
    }
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    ix->thread_streams[std::this_thread::get_id()] = st;
    *out = st;
    return MMIDX_OK;
}

// enqueue(int *launches) issues the call on `st`; run_graphed decides between eager issue, capture and replay
template <typename F>
static int run_graphed(mmidx_index *ix, GraphKey key, int parity, cudaStream_t st, F enqueue) {
    int launches = 0;
    // the legacy default stream (what a caller passes as stream 0) and the per-thread default stream cannot be captured
    if (!ix->use_graph || st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread) {
        RET(enqueue(&launches));
        ix->last_launches = launches;
        return MMIDX_OK;
    }
    GraphCache *cache;
    {
        std::lock_guard<std::mutex> lk0(ix->graph_mu);
        GraphCache *&slot = ix->graphs[st];
        if (!slot) slot = new GraphCache();
        cache = slot;
    }
    std::lock_guard<std::mutex> lk(cache->mu);
    if (cache->m.size() > 64) clear_cache(*cache);  // callers that never repeat a shape: do not accumulate
    key.flags = (key.flags << 1) | (ix->timer.enabled ? 1 : 0);
    GraphEntry &e = cache->m[key];
    uint64_t agen = 0, spills0 = 0;
    {
        std::lock_guard<std::mutex> lk2(ix->arena_mu);
        Arena &a = ix->arenas[st];
        agen = a.gen;
        spills0 = a.spills;
        if (a.depth != 0 && a.owner != std::this_thread::get_id()) e.disabled = true;  // another thread is mid-call on this stream
    }
    if (e.gen != ix->gen || e.arena_gen != agen) {
        for (auto &x : e.exec)
            if (x) cudaGraphExecDestroy(x);
        e = GraphEntry();
        e.gen = ix->gen;
        e.arena_gen = agen;
    }
    if (e.exec[parity] && !e.disabled) {
        CK(cudaGraphLaunch(e.exec[parity], st));
        ix->last_launches = e.launches;
        return MMIDX_OK;
    }
    if (e.disabled || e.seen < 2) {
        e.seen++;
        RET(enqueue(&launches));
        ix->last_launches = launches;
        return MMIDX_OK;
    }
    // capture
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        e.disabled = true;
        RET(enqueue(&launches));
        ix->last_launches = launches;
        return MMIDX_OK;
    }
    ix->timer.capturing = true;
    const int rc = enqueue(&launches);
    ix->timer.capturing = false;
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    bool spilled;
    {
        std::lock_guard<std::mutex> lk2(ix->arena_mu);
        Arena &a = ix->arenas[st];
        spilled = a.spills != spills0 || a.gen != agen;
    }
    if (rc != MMIDX_OK || ce != cudaSuccess || !graph || spilled) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        e.disabled = true;  // this shape does not capture cleanly: stay eager
        if (rc != MMIDX_OK) return rc;
        launches = 0;
        RET(enqueue(&launches));
        ix->last_launches = launches;
        return MMIDX_OK;
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess || !exec) {
        cudaGetLastError();
        e.disabled = true;
        launches = 0;
        RET(enqueue(&launches));
        ix->last_launches = launches;
        return MMIDX_OK;
    }
    e.exec[parity] = exec;
    e.launches = launches;
    CK(cudaGraphLaunch(exec, st));
    ix->last_launches = launches;
    return MMIDX_OK;
}

extern "C" int mmidx_search_dev(mmidx_t *ix, int64_t nq, const double *dQ, int32_t k, int32_t *d_iids, double *d_dist,
                                int32_t *d_count, void *stream) {
    if (!ix || (nq > 0 && (!dQ || !d_iids || !d_dist || !d_count))) return fail(MMIDX_ERR_INVALID, "null argument");
    DeviceGuard g(ix->device);
    if (ix->shard_count > 1)
        return fail(MMIDX_ERR_STATE, "sharded index: use mmidx_search_multi_dev (or mmidx_search_shard_dev + mmidx_merge_topk_dev)");
    int w = 0;
    RET(validate_search(ix, nq, k, &w));
    if (nq == 0) return MMIDX_OK;
    RET(prepare_search(ix, k));
    cudaStream_t st = (cudaStream_t)stream;
    GraphKey key{0, nq, k, w, 0, dQ, d_iids, d_dist, d_count, st};
    return run_graphed(ix, key, 0, st, [&](int *launches) {
        return search_dev_impl(ix, nq, dQ, k, d_iids, d_dist, nullptr, nullptr, d_count, st, false, nullptr, nullptr, true, launches);
    });
}

extern "C" int mmidx_search_shard_dev(mmidx_t *ix, int64_t nq, const double *dQ, int32_t k, const int32_t *d_probes,
                                      int32_t *d_iids, double *d_dist, int64_t *d_seq, double *d_tie, int32_t *d_count,
                                      void *stream) {
    if (!ix || (nq > 0 && (!dQ || !d_iids || !d_dist || !d_seq || !d_tie || !d_count)))
        return fail(MMIDX_ERR_INVALID, "null argument");
    if (ix->p.type != MMIDX_IVFPQ) return fail(MMIDX_ERR_INVALID, "sharded search is defined for IVFPQ");
    DeviceGuard g(ix->device);
    return search_dev_impl(ix, nq, dQ, k, d_iids, d_dist, (unsigned long long *)d_seq, d_tie, d_count, (cudaStream_t)stream, true,
                           d_probes);
}

extern "C" int mmidx_merge_topk_dev(int64_t nq, int32_t k, int32_t nparts, const int32_t *d_iids, const double *d_dist,
                                    const int64_t *d_seq, const double *d_tie, const int32_t *d_count, int32_t *d_out_iids,
                                    double *d_out_dist, int64_t *d_out_seq, int32_t *d_out_count, int32_t *d_amb_list,
                                    int32_t *d_amb_count, void *stream) {
    if (nq < 0 || k < 1 || k > MMIDX_MAX_K || nparts < 1) return fail(MMIDX_ERR_INVALID, "bad merge geometry");
    if (nq == 0) return MMIDX_OK;
    if (!d_iids || !d_dist || !d_seq || !d_count || !d_out_iids || !d_out_dist || !d_out_count || !d_amb_list || !d_amb_count)
        return fail(MMIDX_ERR_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemsetAsync(d_amb_count, 0, sizeof(int32_t), st));
    TopkOut part{};
    part.iids = const_cast<int32_t *>(d_iids);
    part.dist = const_cast<double *>(d_dist);
    part.seq = (unsigned long long *)const_cast<int64_t *>(d_seq);
    part.cnt = const_cast<int32_t *>(d_count);
    part.tie = const_cast<double *>(d_tie);
    ResultBufs res{d_out_iids, d_out_dist, (unsigned long long *)d_out_seq, d_out_count};
    int launches = 0;
    // all-gather layout [nparts][nq][k]: row(part, q) = part*nq + q
    if (nq > 32768) return fail(MMIDX_ERR_UNSUPPORTED, "merge handles at most 32768 queries per call");
    if (cap_for(k) == 2048)
        RET(launch_merge<2048>(part, nparts, nq, k, res, d_amb_list, d_amb_count, nullptr, nq, 1, st, &launches));
    else
        RET(launch_merge<1024>(part, nparts, nq, k, res, d_amb_list, d_amb_count, nullptr, nq, 1, st, &launches));
    return MMIDX_OK;
}

// sharded tie pass, step 1 (per shard): first-k entries in offer order with dist <= T of every ambiguous query.
// T is read from the merged result d_res_dist[q][k-1].  Needs the LUTs again, so it re-runs coarse + LUT for
// the ambiguous queries only (rare path).
extern "C" int mmidx_tie_collect_shard_dev(mmidx_t *ix, int64_t nq, const double *dQ, int32_t k, const double *d_res_dist,
                                           const int32_t *d_amb_list, const int32_t *d_amb_count, int64_t *d_l_seq,
                                           int32_t *d_l_iid, int32_t *d_l_eq, int32_t *d_l_cnt, void *stream) {
    if (!ix || !dQ || !d_res_dist || !d_amb_list || !d_amb_count || !d_l_seq || !d_l_iid || !d_l_eq || !d_l_cnt)
        return fail(MMIDX_ERR_INVALID, "null argument");
    if (ix->p.type != MMIDX_IVFPQ) return fail(MMIDX_ERR_INVALID, "sharded search is defined for IVFPQ");
    DeviceGuard g(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    int w = 0;
    RET(validate_search(ix, nq, k, &w));
    // host needs the ambiguous count to size the recomputation
    int32_t na = 0;
    CK(cudaMemcpyAsync(&na, d_amb_count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaMemsetAsync(d_l_cnt, 0, sizeof(int32_t) * (size_t)nq, st));
    if (na == 0) return MMIDX_OK;
    std::vector<int32_t> amb(na);
    CK(cudaMemcpyAsync(amb.data(), d_amb_list, sizeof(int32_t) * (size_t)na, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    int launches = 0;
    const int m = ix->p.m, ks = ix->p.ks, d = ix->p.d;
    // process ambiguous queries in small groups: gather their vectors, rebuild probes + LUTs, collect
    const int64_t G = std::max<int64_t>(1, (int64_t)(ix->lut_chunk_bytes / ((size_t)w * m * ks * sizeof(double))));
    for (int64_t a0 = 0; a0 < na; a0 += G) {
        int64_t nb = std::min<int64_t>(G, na - a0);
        Scratch sc(st);
        double *gQ, *dlut, *gT;
        int32_t *dprobes, *gl, *gcount;
        int64_t *gidx;
        RET(sc.get(&gQ, (size_t)nb * d));
        RET(sc.get(&dlut, (size_t)nb * w * lut_stride_of(ix)));
        RET(sc.get(&dprobes, (size_t)nb * w));
        RET(sc.get(&gidx, (size_t)nb));
        RET(sc.get(&gT, (size_t)nb * k));
        RET(sc.get(&gl, (size_t)nb));
        RET(sc.get(&gcount, 1));
        std::vector<int64_t> hidx(nb);
        std::vector<int32_t> hl(nb);
        for (int64_t i = 0; i < nb; ++i) {
            hidx[i] = amb[a0 + i];
            hl[i] = (int32_t)i;
        }
        int32_t nb32 = (int32_t)nb;
        CK(cudaMemcpyAsync(gidx, hidx.data(), sizeof(int64_t) * (size_t)nb, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(gl, hl.data(), sizeof(int32_t) * (size_t)nb, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(gcount, &nb32, sizeof(int32_t), cudaMemcpyHostToDevice, st));
        k_gather_rows<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>((const uint8_t *)dQ, gidx, nb, d * (int)sizeof(double), (uint8_t *)gQ);
        RET(post_launch("k_gather_rows", &launches));
        k_gather_rows<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>((const uint8_t *)d_res_dist, gidx, nb, k * (int)sizeof(double), (uint8_t *)gT);
        RET(post_launch("k_gather_rows", &launches));
        RET(coarse_probe_dev(ix, gQ, nb, w, dprobes, sc, st, &launches));
        RET(launch_lut(ix, gQ, dprobes, nb * w, w, dlut, st, &launches, sc));
        TieLists tl;
        RET(sc.get(&tl.seq, (size_t)nb * k));
        RET(sc.get(&tl.pay, (size_t)nb * k));
        RET(sc.get(&tl.eq, (size_t)nb * k));
        RET(sc.get(&tl.cnt, (size_t)nb));
        TieCodeArgs t{};
        t.luts = dlut;
    t.lut_stride = lut_stride_of(ix);
        t.codes = ix->csr_codes.as<uint8_t>();
        t.iids = ix->csr_iids.as<int32_t>();
        t.probes = dprobes;
        t.list_off = ix->dlist_off.as<int64_t>();
        t.list_len = ix->dlist_len.as<int32_t>();
        t.w = w;
        t.m = m;
        t.ks = ks;
        t.code_bytes = ix->code_bytes;
        t.k = k;
        k_tie_collect_ivfpq<<<(unsigned)std::min<int64_t>(nb, 2 * ix->sm_count), MMIDX_NT, 0, st>>>(t, gT, gl, gcount, tl);
        RET(post_launch("k_tie_collect_ivfpq", &launches));
        // scatter the group's lists back to rows indexed by the original query id
        for (int64_t i = 0; i < nb; ++i) {
            int64_t q = hidx[i];
            CK(cudaMemcpyAsync(d_l_seq + q * k, tl.seq + i * k, sizeof(int64_t) * (size_t)k, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(d_l_iid + q * k, tl.pay + i * k, sizeof(int32_t) * (size_t)k, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(d_l_eq + q * k, tl.eq + i * k, sizeof(int32_t) * (size_t)k, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(d_l_cnt + q, tl.cnt + i, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        }
        CK(cudaStreamSynchronize(st));
    }
    return MMIDX_OK;
}

// sharded tie pass, step 2 (after all-gathering the lists): patch the merged result in place
extern "C" int mmidx_tie_finish_dev(int64_t nq, int32_t k, int32_t nparts, const int64_t *d_l_seq, const int32_t *d_l_iid,
                                    const int32_t *d_l_eq, const int32_t *d_l_cnt, const int32_t *d_amb_list,
                                    const int32_t *d_amb_count, int32_t *d_res_iids, double *d_res_dist, void *stream) {
    if (nq <= 0) return MMIDX_OK;
    if (k < 1 || k > MMIDX_MAX_K || nparts < 1) return fail(MMIDX_ERR_INVALID, "bad geometry");
    cudaStream_t st = (cudaStream_t)stream;
    RET(set_smem(k_tie_finish, (size_t)1024 * 16));
    k_tie_finish<<<(unsigned)std::min<int64_t>(nq, 2 * current_sm_count()), MMIDX_NT, (size_t)k * 16, st>>>(
        nparts, nq, k, (const unsigned long long *)d_l_seq, d_l_iid, d_l_eq, d_l_cnt, d_amb_list, d_amb_count, d_res_iids,
        d_res_dist, nullptr, PeerSink{});
    return post_launch("k_tie_finish", nullptr);
}

extern "C" int mmidx_search(mmidx_t *ix, int64_t nq, const double *Q, int32_t k, int32_t *out_iids, double *out_dist,
                            int32_t *out_count) {
    if (!ix || (nq > 0 && (!Q || !out_iids || !out_dist))) return fail(MMIDX_ERR_INVALID, "null argument");
    if (ix->shard_count > 1) return fail(MMIDX_ERR_STATE, "sharded index: use the *_shard_dev entry points");
    int w = 0;
    RET(validate_search(ix, nq, k, &w));
    if (nq == 0) return MMIDX_OK;
    DeviceGuard g(ix->device);
    cudaStream_t st;
    RET(host_stream(ix, &st));
    Scratch sc(st, ix->arenas, ix->arena_mu);
    double *dQ, *ddist;
    int32_t *diids, *dcnt;
    RET(sc.get(&dQ, (size_t)nq * ix->p.d));
    RET(sc.get(&ddist, (size_t)nq * k));
    RET(sc.get(&diids, (size_t)nq * k));
    RET(sc.get(&dcnt, (size_t)nq));
    // Query chunks are pipelined over three streams: the host->device copy of chunk c+1 and the device->host copy of
    // chunk c-1 run under the kernels of chunk c (pinned host buffers; pageable ones still work, just serialised).
    const int64_t CH = ix->host_chunk;
    if (nq <= 2 * CH) {
        CK(cudaMemcpyAsync(dQ, Q, sizeof(double) * (size_t)nq * ix->p.d, cudaMemcpyHostToDevice, st));
        RET(prepare_search(ix, k));
        // the device part of a small call is one graph launch (the arena hands every call the same scratch addresses)
        GraphKey key{0, nq, k, w, 0, dQ, diids, ddist, dcnt, st};
        RET(run_graphed(ix, key, 0, st, [&](int *launches) {
            return search_dev_impl(ix, nq, dQ, k, diids, ddist, nullptr, nullptr, dcnt, st, false, nullptr, nullptr, true, launches);
        }));
        CK(cudaMemcpyAsync(out_iids, diids, sizeof(int32_t) * (size_t)nq * k, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(out_dist, ddist, sizeof(double) * (size_t)nq * k, cudaMemcpyDeviceToHost, st));
        if (out_count) CK(cudaMemcpyAsync(out_count, dcnt, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return MMIDX_OK;
    }
    if (!ix->h2d_stream) {
        std::lock_guard<std::mutex> lk(ix->mu);
        if (!ix->h2d_stream) {
            CK(cudaStreamCreateWithFlags(&ix->h2d_stream, cudaStreamNonBlocking));
            CK(cudaStreamCreateWithFlags(&ix->d2h_stream, cudaStreamNonBlocking));
            CK(cudaStreamCreateWithFlags(&ix->comp2_stream, cudaStreamNonBlocking));
        }
    }
    cudaStream_t sh = ix->h2d_stream, sd = ix->d2h_stream, sc2 = ix->comp2_stream;
    const int nch = (int)((nq + CH - 1) / CH);
    // events come from a per-index pool (no cudaEventCreate / Destroy on the hot path)
    std::vector<cudaEvent_t> ev((size_t)2 * nch + 1, nullptr);
    int rc = MMIDX_OK;
    {
        std::lock_guard<std::mutex> lk(ix->ev_mu);
        for (auto &e : ev) {
            if (!ix->ev_pool.empty()) {
                e = ix->ev_pool.back();
                ix->ev_pool.pop_back();
            } else if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
                e = nullptr;
                rc = fail(MMIDX_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(cudaGetLastError()));
                break;
            }
        }
    }
    auto mk = [&](cudaEvent_t &e, cudaStream_t s_) {
        if (rc == MMIDX_OK && cudaEventRecord(e, s_) != cudaSuccess)
            rc = fail(MMIDX_ERR_CUDA, "cudaEventRecord: %s", cudaGetErrorString(cudaGetLastError()));
    };
    auto ckc = [&](cudaError_t e_, const char *what) {
        if (rc == MMIDX_OK && e_ != cudaSuccess) rc = fail(MMIDX_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e_));
    };
    mk(ev[2 * nch], st);  // the scratch buffers exist (stream-ordered allocation on st)
    if (rc == MMIDX_OK && cudaStreamWaitEvent(sh, ev[2 * nch], 0) != cudaSuccess) rc = fail(MMIDX_ERR_CUDA, "cudaStreamWaitEvent");
    if (rc == MMIDX_OK && cudaStreamWaitEvent(sd, ev[2 * nch], 0) != cudaSuccess) rc = fail(MMIDX_ERR_CUDA, "cudaStreamWaitEvent");
    if (rc == MMIDX_OK && cudaStreamWaitEvent(sc2, ev[2 * nch], 0) != cudaSuccess) rc = fail(MMIDX_ERR_CUDA, "cudaStreamWaitEvent");
    for (int c = 0; c < nch && rc == MMIDX_OK; ++c) {
        const int64_t q0 = (int64_t)c * CH, nb = std::min(CH, nq - q0);
        if (cudaMemcpyAsync(dQ + q0 * ix->p.d, Q + q0 * ix->p.d, sizeof(double) * (size_t)nb * ix->p.d, cudaMemcpyHostToDevice, sh) != cudaSuccess) {
            rc = fail(MMIDX_ERR_CUDA, "cudaMemcpyAsync H2D");
            break;
        }
        mk(ev[2 * c], sh);
        if (rc != MMIDX_OK) break;
        cudaStream_t cs = (c & 1) ? sc2 : st;
        ckc(cudaStreamWaitEvent(cs, ev[2 * c], 0), "cudaStreamWaitEvent");
        if (rc != MMIDX_OK) break;
        g_overlapped_chunks = true;
        rc = search_dev_impl(ix, nb, dQ + q0 * ix->p.d, k, diids + q0 * k, ddist + q0 * k, nullptr, nullptr, dcnt + q0, cs, false);
        g_overlapped_chunks = false;
        if (rc != MMIDX_OK) break;
        mk(ev[2 * c + 1], cs);
        if (rc != MMIDX_OK) break;
        ckc(cudaStreamWaitEvent(sd, ev[2 * c + 1], 0), "cudaStreamWaitEvent");
        ckc(cudaMemcpyAsync(out_iids + q0 * k, diids + q0 * k, sizeof(int32_t) * (size_t)nb * k, cudaMemcpyDeviceToHost, sd), "cudaMemcpyAsync D2H");
        ckc(cudaMemcpyAsync(out_dist + q0 * k, ddist + q0 * k, sizeof(double) * (size_t)nb * k, cudaMemcpyDeviceToHost, sd), "cudaMemcpyAsync D2H");
        if (out_count) ckc(cudaMemcpyAsync(out_count + q0, dcnt + q0, sizeof(int32_t) * (size_t)nb, cudaMemcpyDeviceToHost, sd), "cudaMemcpyAsync D2H");
    }
    cudaError_t e1 = cudaStreamSynchronize(sh), e2 = cudaStreamSynchronize(st), e3 = cudaStreamSynchronize(sd);
    if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(sc2);
    {
        std::lock_guard<std::mutex> lk(ix->ev_mu);
        for (cudaEvent_t e : ev)
            if (e) ix->ev_pool.push_back(e);
    }
    if (rc != MMIDX_OK) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
        return fail(MMIDX_ERR_CUDA, "mmidx_search: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
    return MMIDX_OK;
}

extern "C" int mmidx_coarse_probe(mmidx_t *ix, int64_t nq, const double *Q, int32_t w, int32_t *out) {
    if (!ix || (nq > 0 && (!Q || !out))) return fail(MMIDX_ERR_INVALID, "null argument");
    if (ix->p.type != MMIDX_IVFPQ) return fail(MMIDX_ERR_INVALID, "coarse probe applies to IVFPQ only");
    if (!ix->has_C) return fail(MMIDX_ERR_STATE, "coarse quantizer not loaded");
    if (w < 1 || w > ix->p.nlist) return fail(MMIDX_ERR_W, "w = %d out of 1..nlist", w);
    if (w > MMIDX_MAX_K) return fail(MMIDX_ERR_UNSUPPORTED, "w = %d exceeds %d", w, MMIDX_MAX_K);
    if (nq == 0) return MMIDX_OK;
    DeviceGuard g(ix->device);
    cudaStream_t st = ix->stream;
    int launches = 0;
    const int64_t QC = 8192;
    for (int64_t q0 = 0; q0 < nq; q0 += QC) {
        int64_t nb = std::min(QC, nq - q0);
        Scratch sc(st);
        double *dQ;
        int32_t *dpr;
        RET(sc.get(&dQ, (size_t)nb * ix->p.d));
        RET(sc.get(&dpr, (size_t)nb * w));
        CK(cudaMemcpyAsync(dQ, Q + q0 * ix->p.d, sizeof(double) * (size_t)nb * ix->p.d, cudaMemcpyHostToDevice, st));
        RET(coarse_probe_dev(ix, dQ, nb, w, dpr, sc, st, &launches));
        CK(cudaMemcpyAsync(out + q0 * w, dpr, sizeof(int32_t) * (size_t)nb * w, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    ix->last_launches = launches;
    return MMIDX_OK;
}

extern "C" int mmidx_coarse_probe_dev(mmidx_t *ix, int64_t nq, const double *dQ, int32_t w, int32_t *d_out, void *stream) {
    if (!ix || (nq > 0 && (!dQ || !d_out))) return fail(MMIDX_ERR_INVALID, "null argument");
    if (ix->p.type != MMIDX_IVFPQ) return fail(MMIDX_ERR_INVALID, "coarse probe applies to IVFPQ only");
    if (!ix->has_C) return fail(MMIDX_ERR_STATE, "coarse quantizer not loaded");
    if (w < 1 || w > ix->p.nlist) return fail(MMIDX_ERR_W, "w = %d out of 1..nlist", w);
    if (w > MMIDX_MAX_K) return fail(MMIDX_ERR_UNSUPPORTED, "w = %d exceeds %d", w, MMIDX_MAX_K);
    DeviceGuard g(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    int launches = 0;
    const int64_t QC = std::max<int64_t>(1, (int64_t)(((size_t)1 << 30) / ((size_t)ix->p.nlist * sizeof(double))));
    for (int64_t q0 = 0; q0 < nq; q0 += QC) {
        int64_t nb = std::min(QC, nq - q0);
        Scratch sc(st);
        RET(coarse_probe_dev(ix, dQ + q0 * ix->p.d, nb, w, d_out + q0 * w, sc, st, &launches));
    }
    return MMIDX_OK;
}

extern "C" int mmidx_pq_lut(mmidx_t *ix, int64_t nq, const double *V, double *out) {
    if (!ix || (nq > 0 && (!V || !out))) return fail(MMIDX_ERR_INVALID, "null argument");
    if (ix->p.type == MMIDX_LINEAR) return fail(MMIDX_ERR_INVALID, "Linear index has no ADC table");
    if (!ix->has_P) return fail(MMIDX_ERR_STATE, "product quantizer not loaded");
    if (nq == 0) return MMIDX_OK;
    DeviceGuard g(ix->device);
    cudaStream_t st = ix->stream;
    Scratch sc(st);
    double *dV, *dl;
    const size_t row = (size_t)ix->p.m * ix->p.ks, stride = (size_t)lut_stride_of(ix);
    RET(sc.get(&dV, (size_t)nq * ix->p.d));
    RET(sc.get(&dl, (size_t)nq * stride));
    CK(cudaMemcpyAsync(dV, V, sizeof(double) * (size_t)nq * ix->p.d, cudaMemcpyHostToDevice, st));
    int launches = 0;
    // computeLookupADC takes the ALREADY transformed vector (PQ.java:387): no permutation here
    RET(launch_lut(ix, dV, nullptr, nq, 1, dl, st, &launches, sc, false));
    CK(cudaMemcpy2DAsync(out, row * sizeof(double), dl, stride * sizeof(double), row * sizeof(double), (size_t)nq,
                         cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ix->last_launches = launches;
    return MMIDX_OK;
}

// ---------------------------------------------------------------------------------------------------------
// introspection
// ---------------------------------------------------------------------------------------------------------
extern "C" int mmidx_size(mmidx_t *ix, int64_t *out) {
    if (!ix || !out) return fail(MMIDX_ERR_INVALID, "null argument");
    *out = ix->n;
    return MMIDX_OK;
}

extern "C" int mmidx_list_sizes(mmidx_t *ix, int32_t *out) {
    if (!ix || !out) return fail(MMIDX_ERR_INVALID, "null argument");
    if (ix->p.type != MMIDX_IVFPQ) return fail(MMIDX_ERR_INVALID, "only IVFPQ has inverted lists");
    std::lock_guard<std::mutex> lk(ix->mu);
    std::fill(out, out + ix->p.nlist, 0);
    for (int32_t l : ix->h_list) out[l]++;
    return MMIDX_OK;
}

extern "C" int mmidx_get_vector(mmidx_t *ix, int64_t iid, double *out) {
    if (!ix || !out) return fail(MMIDX_ERR_INVALID, "null argument");
    if (ix->p.type != MMIDX_LINEAR) return fail(MMIDX_ERR_INVALID, "getVector applies to Linear only");
    // Linear.java:254-256: "Internal id does not exist!"
    if (iid < 0 || iid >= ix->n) return fail(MMIDX_ERR_INVALID, "Internal id does not exist!");
    DeviceGuard g(ix->device);
    cudaStream_t st = ix->stream;
    Scratch sc(st);
    double *dv;
    RET(sc.get(&dv, (size_t)ix->p.d));
    k_linear_unpack_row<<<(ix->p.d + 255) / 256, 256, 0, st>>>(ix->dXb.as<double>(), ix->p.d, iid, dv);
    RET(post_launch("k_linear_unpack_row", nullptr));
    CK(cudaMemcpyAsync(out, dv, sizeof(double) * (size_t)ix->p.d, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return MMIDX_OK;
}

extern "C" int mmidx_scan_bytes(mmidx_t *ix, int64_t nq, const double *Q, int64_t *out_total) {
    if (!ix || !out_total) return fail(MMIDX_ERR_INVALID, "null argument");
    if (ix->p.type == MMIDX_LINEAR) {
        *out_total = nq * ix->n_local * ix->p.d * 8;
        return MMIDX_OK;
    }
    if (ix->p.type == MMIDX_PQ) {
        *out_total = nq * ix->n_local * ix->code_bytes;
        return MMIDX_OK;
    }
    int w = 0;
    RET(validate_search(ix, nq, 1, &w));
    std::vector<int32_t> probes((size_t)nq * w);
    RET(mmidx_coarse_probe(ix, nq, Q, w, probes.data()));
    std::vector<int32_t> len(ix->p.nlist);
    RET(mmidx_list_sizes(ix, len.data()));
    int64_t tot = 0;
    for (size_t i = 0; i < probes.size(); ++i) tot += (int64_t)len[probes[i]] * (ix->code_bytes + 4);
    *out_total = tot;
    return MMIDX_OK;
}

static int last_timings_n(mmidx_t *ix, float *out, int n) {
    if (!ix || !out) return fail(MMIDX_ERR_INVALID, "null argument");
    DeviceGuard g(ix->device);
    for (int i = 0; i < n; ++i) out[i] = 0.f;
    for (auto &s : ix->timer.spans) {
        CK(cudaEventSynchronize(s.b));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, s.a, s.b));
        if (s.stage < n) out[s.stage] += ms;
    }
    return MMIDX_OK;
}

extern "C" int mmidx_last_timings(mmidx_t *ix, float *out5) { return last_timings_n(ix, out5, 5); }
extern "C" int mmidx_last_timings_multi(mmidx_t *ix, float *out8) { return last_timings_n(ix, out8, 8); }

extern "C" int mmidx_debug_stats(mmidx_t *ix, uint64_t *out4) {
    if (!ix || !out4) return fail(MMIDX_ERR_INVALID, "null argument");
    if (!ix->want_stats || !ix->dstats.p) return fail(MMIDX_ERR_STATE, "MMIDX_STATS=1 was not set or no fast search ran yet");
    DeviceGuard g(ix->device);
    CK(cudaStreamSynchronize(ix->stream));
    CK(cudaMemcpy(out4, ix->dstats.p, sizeof(uint64_t) * 4, cudaMemcpyDeviceToHost));
    return MMIDX_OK;
}

extern "C" int mmidx_last_launches(mmidx_t *ix, int32_t *out) {
    if (!ix || !out) return fail(MMIDX_ERR_INVALID, "null argument");
    *out = ix->last_launches;
    return MMIDX_OK;
}

// ---------------------------------------------------------------------------------------------------------
// multi-GPU: exchange windows over CUDA IPC + the whole sharded search step (comm.cuh)
// ---------------------------------------------------------------------------------------------------------
struct Comm {
    int rank = 0, world = 1, S = 1, R = 1, shard = 0, group = 0;
    int64_t max_gq = 0;
    int k_max = 0, w_max = 0;
    WinLayout L;
    unsigned char *win = nullptr;
    unsigned char *peer[MMIDX_MAX_PEERS] = {};
    bool ipc_opened[MMIDX_MAX_PEERS] = {};
    bool attached = false;
    uint64_t calls = 0;
    unsigned long long timeout_ns = 60ull * 1000000000ull;
};

struct CommHandle {  // what mmidx_comm_create hands out for the rendezvous (MMIDX_COMM_HANDLE_BYTES)
    cudaIpcMemHandle_t ipc;
    uint64_t ptr;
    int64_t pid;
    int32_t device, rank, world, S;
    uint64_t bytes;
};
static_assert(sizeof(CommHandle) <= MMIDX_COMM_HANDLE_BYTES, "handle size");

static void comm_release(mmidx_index *ix) {
    Comm *c = ix->comm;
    if (!c) return;
    for (int r = 0; r < c->world && r < MMIDX_MAX_PEERS; ++r)
        if (c->ipc_opened[r] && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
    if (c->win) cudaFree(c->win);
    delete c;
    ix->comm = nullptr;
}

extern "C" int mmidx_comm_create(mmidx_t *ix, int32_t rank, int32_t world, int32_t list_shards, int64_t max_gq, int32_t k_max,
                                 void *handle_out) {
    if (!ix || !handle_out) return fail(MMIDX_ERR_INVALID, "null argument");
    if (ix->p.type != MMIDX_IVFPQ) return fail(MMIDX_ERR_INVALID, "the multi-GPU exchange is defined for IVFPQ");
    if (world < 1 || world > MMIDX_MAX_PEERS) return fail(MMIDX_ERR_UNSUPPORTED, "world must be 1..%d (one NVSwitch node)", MMIDX_MAX_PEERS);
    if (rank < 0 || rank >= world) return fail(MMIDX_ERR_INVALID, "rank out of range");
    if (list_shards < 1 || world % list_shards != 0) return fail(MMIDX_ERR_INVALID, "list_shards must divide world");
    if (ix->shard_count != list_shards || ix->shard_rank != rank % list_shards)
        return fail(MMIDX_ERR_STATE, "index was created as shard %d of %d; rank %d of a job with %d list shards needs shard %d of %d",
                    ix->shard_rank, ix->shard_count, rank, list_shards, rank % list_shards, list_shards);
    if (max_gq < 1 || k_max < 1 || k_max > MMIDX_MAX_K) return fail(MMIDX_ERR_INVALID, "bad window geometry");
    DeviceGuard g(ix->device);
    std::lock_guard<std::mutex> lk(ix->mu);
    comm_release(ix);
    ix->gen++;
    Comm *c = new Comm();
    c->rank = rank;
    c->world = world;
    c->S = list_shards;
    c->R = world / list_shards;
    c->shard = rank % list_shards;
    c->group = rank / list_shards;
    c->max_gq = max_gq;
    c->k_max = k_max;
    c->w_max = std::min<int>(ix->p.nlist, MMIDX_MAX_K);
    if (const char *e = getenv("MMIDX_COMM_TIMEOUT_S")) {
        long v = atol(e);
        if (v >= 1) c->timeout_ns = (unsigned long long)v * 1000000000ull;
    }
    c->L = make_layout(c->S, c->R, max_gq, k_max, c->w_max);
    cudaError_t e = cudaMalloc((void **)&c->win, c->L.total);  // plain cudaMalloc: legacy CUDA IPC cannot export pool memory
    if (e != cudaSuccess) {
        delete c;
        return fail(MMIDX_ERR_CUDA, "cudaMalloc of the %zu-byte exchange window: %s", c->L.total, cudaGetErrorString(e));
    }
    ix->comm = c;
    CK(cudaMemset(c->win, 0, c->L.data0));  // flags and epoch
    CK(cudaDeviceSynchronize());
    c->peer[rank] = c->win;
    CommHandle h;
    memset(&h, 0, sizeof(h));
    if (world > 1) CK(cudaIpcGetMemHandle(&h.ipc, c->win));
    h.ptr = (uint64_t)(uintptr_t)c->win;
    h.pid = (int64_t)getpid();
    h.device = ix->device;
    h.rank = rank;
    h.world = world;
    h.S = list_shards;
    h.bytes = c->L.total;
    memset(handle_out, 0, MMIDX_COMM_HANDLE_BYTES);
    memcpy(handle_out, &h, sizeof(h));
    if (world == 1) c->attached = true;
    return MMIDX_OK;
}

extern "C" int mmidx_comm_attach(mmidx_t *ix, const void *handles) {
    if (!ix || !handles) return fail(MMIDX_ERR_INVALID, "null argument");
    Comm *c = ix->comm;
    if (!c) return fail(MMIDX_ERR_STATE, "mmidx_comm_create first");
    DeviceGuard g(ix->device);
    std::lock_guard<std::mutex> lk(ix->mu);
    for (int r = 0; r < c->world; ++r) {
        CommHandle h;
        memcpy(&h, (const unsigned char *)handles + (size_t)r * MMIDX_COMM_HANDLE_BYTES, sizeof(h));
        if (h.rank != r || h.world != c->world || h.S != c->S || h.bytes != c->L.total)
            return fail(MMIDX_ERR_INVALID, "handle %d does not belong to this job (rank %d, world %d, shards %d, %llu bytes)", r,
                        h.rank, h.world, h.S, (unsigned long long)h.bytes);
        if (r == c->rank) continue;
        if (h.pid == (int64_t)getpid()) {
            // same process (one thread per GPU): plain peer access to the other device's allocation
            if (h.device != ix->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(MMIDX_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", h.device, cudaGetErrorString(e));
                cudaGetLastError();
            }
            c->peer[r] = (unsigned char *)(uintptr_t)h.ptr;
        } else {
            void *p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, h.ipc, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess)
                return fail(MMIDX_ERR_CUDA, "cudaIpcOpenMemHandle of rank %d's window: %s", r, cudaGetErrorString(e));
            c->peer[r] = (unsigned char *)p;
            c->ipc_opened[r] = true;
        }
    }
    c->attached = true;
    return MMIDX_OK;
}

extern "C" int mmidx_comm_destroy(mmidx_t *ix) {
    if (!ix) return MMIDX_OK;
    DeviceGuard g(ix->device);
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(ix->mu);
    drop_graphs(ix);
    comm_release(ix);
    ix->gen++;
    return MMIDX_OK;
}

// one whole step of the job on this rank, enqueued on st (see comm.cuh for the stages)
static int multi_enqueue(mmidx_index *ix, int64_t gq, const double *dQ, int k, int w, bool gather, int par, cudaStream_t st,
                         int *launches) {
    Comm &c = *ix->comm;
    const WinLayout &L = c.L;
    const int S = c.S, s = c.shard, g = c.group, d = ix->p.d;
    const int64_t sl = (gq + S - 1) / S, gqp = sl * S;
    auto pbase = [&](int r) { return c.peer[r] + L.data0 + (size_t)par * L.parity_stride; };
    unsigned char *mine = pbase(c.rank);
    unsigned *epoch = reinterpret_cast<unsigned *>(c.win + L.epoch);
    ix->timer.reset();
    StageMark whole(ix, st, 4);
    k_comm_begin<<<1, 1, 0, st>>>(epoch);
    RET(post_launch("k_comm_begin", launches));
    auto sync = [&](int stage, bool all_ranks) -> int {
        CommSync cs{};
        cs.epoch = epoch;
        cs.timeout_ns = c.timeout_ns;
        int n = 0;
        for (int r = 0; r < c.world; ++r) {
            if (r == c.rank || (!all_ranks && r / S != g)) continue;
            cs.remote[n] = reinterpret_cast<unsigned *>(c.peer[r] + L.flags) + stage * MMIDX_MAX_PEERS + c.rank;
            cs.local[n] = reinterpret_cast<const unsigned *>(c.win + L.flags) + stage * MMIDX_MAX_PEERS + r;
            ++n;
        }
        cs.n = n;
        if (n == 0) return MMIDX_OK;
        StageMark smx(ix, st, 5);  // exchange point: flag stores + wait for the peers (includes their skew)
        k_comm_sync<<<1, 32, 0, st>>>(cs);
        return post_launch("k_comm_sync", launches);
    };
    // final rows of this rank: [row0, row0 + nrows) of the job-wide result arrays in (every) window
    const int64_t row0 = (int64_t)g * gqp + (S > 1 ? (int64_t)s * sl : 0);
    int32_t *o_iids = reinterpret_cast<int32_t *>(mine + L.o_iids) + row0 * k;
    double *o_dist = reinterpret_cast<double *>(mine + L.o_dist) + row0 * k;
    int32_t *o_cnt = reinterpret_cast<int32_t *>(mine + L.o_cnt) + row0;
    PeerSink fin{};
    if (gather && c.world > 1) {
        fin.mode = 2;
        fin.fields = SINK_IIDS | SINK_DIST | SINK_CNT;
        fin.row0 = row0;
        fin.off_iids = (long long)L.o_iids;
        fin.off_dist = (long long)L.o_dist;
        fin.off_cnt = (long long)L.o_cnt;
        for (int r = 0; r < c.world; ++r)
            if (r != c.rank) fin.base[fin.npeer++] = pbase(r);
    }
    if (S == 1) {
        // replica group: the whole index is here; only the final rows travel
        RET(search_dev_impl(ix, gq, dQ, k, o_iids, o_dist, nullptr, nullptr, o_cnt, st, false, nullptr, &fin, false, launches));
    } else {
        Scratch sc(st, ix->arenas, ix->arena_mu);
        unsigned char *gb[MMIDX_MAX_PEERS];  // windows of the S members of my group, by shard
        for (int t = 0; t < S; ++t) gb[t] = pbase(g * S + t);
        // ---- stage 0: coarse stage of my query slice; the probe rows land in every group member's window ----
        int32_t *probes = reinterpret_cast<int32_t *>(mine + L.probes);
        const int64_t q0 = std::min<int64_t>(gq, (int64_t)s * sl), n0 = std::min<int64_t>(gq, q0 + sl) - q0;
        {
            StageMark sm(ix, st, 0);
            PeerSink ps{};
            ps.mode = 2;
            ps.fields = SINK_IIDS;
            ps.off_iids = (long long)L.probes;
            for (int t = 0; t < S; ++t)
                if (t != s) ps.base[ps.npeer++] = gb[t];
            const int64_t QC = std::max<int64_t>(1, (int64_t)(((size_t)1 << 30) / ((size_t)ix->p.nlist * sizeof(double))));
            for (int64_t b0 = 0; b0 < n0; b0 += QC) {
                const int64_t nb = std::min(QC, n0 - b0);
                ps.q0 = q0 + b0;
                Scratch sc0(st, ix->arenas, ix->arena_mu);
                RET(coarse_probe_dev(ix, dQ + (q0 + b0) * d, nb, w, probes + (q0 + b0) * w, sc0, st, launches, &ps));
            }
            RET(sync(0, false));
        }
        // ---- stage 1: scan the lists stored here for ALL group queries; row of query q -> window of shard q / sl ----
        PeerSink pr{};
        pr.mode = 1;
        pr.npeer = S;
        pr.sl = (int)sl;
        pr.fields = SINK_IIDS | SINK_DIST | SINK_SEQ | SINK_CNT | SINK_TIE;
        pr.row0 = (long long)s * sl;
        pr.off_iids = (long long)L.p_iids;
        pr.off_dist = (long long)L.p_dist;
        pr.off_seq = (long long)L.p_seq;
        pr.off_cnt = (long long)L.p_cnt;
        pr.off_tie = (long long)L.p_tie;
        for (int t = 0; t < S; ++t) pr.base[t] = gb[t];
        RET(search_dev_impl(ix, gq, dQ, k, nullptr, nullptr, nullptr, nullptr, nullptr, st, true, probes, &pr, false, launches));
        StageMark sm3(ix, st, 3);
        RET(sync(1, false));
        // ---- merge my slice (one queue for all probed lists, IVFPQ.java:409,445) ----
        const int64_t nslice = std::max<int64_t>(0, std::min<int64_t>(gq, q0 + sl) - q0);
        int32_t *amb_list, *amb_count;
        RET(sc.get(&amb_list, (size_t)sl));
        RET(sc.get(&amb_count, 1));
        CK(cudaMemsetAsync(amb_count, 0, sizeof(int32_t), st));
        if (nslice > 0) {
            StageMark smm(ix, st, 6);
            TopkOut part{};
            part.iids = reinterpret_cast<int32_t *>(mine + L.p_iids);
            part.dist = reinterpret_cast<double *>(mine + L.p_dist);
            part.seq = reinterpret_cast<unsigned long long *>(mine + L.p_seq);
            part.cnt = reinterpret_cast<int32_t *>(mine + L.p_cnt);
            part.tie = reinterpret_cast<double *>(mine + L.p_tie);
            ResultBufs res{o_iids, o_dist, nullptr, o_cnt};
            res.sink = fin;
            if (cap_for(k) == 2048)
                RET(launch_merge<2048>(part, S, nslice, k, res, amb_list, amb_count, nullptr, sl, 1, st, launches));
            else
                RET(launch_merge<1024>(part, S, nslice, k, res, amb_list, amb_count, nullptr, sl, 1, st, launches));
        }
        // ---- stages 2 + 3: exact ties cut at the k-th boundary (normally none: the kernels below find empty lists) ----
        StageMark smt(ix, st, 7);  // publish + collect + finish of the cross-shard tie pass (their exchange points: stage 5)
        AmbPublish ap{};
        ap.amb_list = amb_list;
        ap.amb_count = amb_count;
        ap.res_dist = o_dist;
        for (int t = 0; t < S; ++t) ap.base[t] = gb[t];
        ap.off_cnt = (long long)L.a_cnt;
        ap.off_q = (long long)L.a_q;
        ap.off_T = (long long)L.a_T;
        ap.S = S;
        ap.shard = s;
        ap.sl = (int)sl;
        ap.k = k;
        k_comm_publish_ties<<<1, MMIDX_NT, 0, st>>>(ap);
        RET(post_launch("k_comm_publish_ties", launches));
        RET(sync(2, false));
        TieMultiArgs tm{};
        tm.t.Q = dQ;
        tm.t.C = ix->dC.as<double>();
        tm.t.P = ix->dP.as<double>();
        tm.t.perm = ix->has_perm ? ix->dperm.as<int32_t>() : nullptr;
        tm.t.probes = probes;
        tm.t.codes = ix->csr_codes.as<uint8_t>();
        tm.t.iids = ix->csr_iids.as<int32_t>();
        tm.t.list_off = ix->dlist_off.as<int64_t>();
        tm.t.list_len = ix->dlist_len.as<int32_t>();
        tm.t.d = d;
        tm.t.m = ix->p.m;
        tm.t.ks = ix->p.ks;
        tm.t.S = ix->S;
        tm.t.w = w;
        tm.t.k = k;
        tm.t.code_bytes = ix->code_bytes;
        tm.t.flat = 0;
        tm.a_cnt = reinterpret_cast<const int32_t *>(mine + L.a_cnt);
        tm.a_q = reinterpret_cast<const int32_t *>(mine + L.a_q);
        tm.a_T = reinterpret_cast<const double *>(mine + L.a_T);
        for (int t = 0; t < S; ++t) tm.base[t] = gb[t];
        tm.off_seq = (long long)L.t_seq;
        tm.off_pay = (long long)L.t_pay;
        tm.off_eq = (long long)L.t_eq;
        tm.off_cnt = (long long)L.t_cnt;
        tm.S = S;
        tm.shard = s;
        tm.sl = (int)sl;
        {
            const size_t tsm = ((size_t)ix->p.m * ix->p.ks + (size_t)d) * sizeof(double);
            const int use_lut = tsm <= 96 * 1024 ? 1 : 0;
            if (use_lut) RET(set_smem(k_tie_collect_multi, tsm));
            k_tie_collect_multi<<<ix->sm_count, MMIDX_NT, use_lut ? tsm : 0, st>>>(tm, use_lut);
            RET(post_launch("k_tie_collect_multi", launches));
        }
        RET(sync(3, false));
        RET(set_smem(k_tie_finish, (size_t)1024 * 16));
        k_tie_finish<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>(nslice, 2 * ix->sm_count)), MMIDX_NT, (size_t)k * 16, st>>>(
            S, sl, k, reinterpret_cast<const unsigned long long *>(mine + L.t_seq), reinterpret_cast<const int32_t *>(mine + L.t_pay),
            reinterpret_cast<const int32_t *>(mine + L.t_eq), reinterpret_cast<const int32_t *>(mine + L.t_cnt), amb_list, amb_count,
            o_iids, o_dist, nullptr, fin);
        RET(post_launch("k_tie_finish", launches));
    }
    if (gather && c.world > 1) {
        StageMark sm3(ix, st, 3);
        RET(sync(4, true));
    }
    return MMIDX_OK;
}

static int multi_validate(mmidx_index *ix, int64_t gq, int k, int *w_out) {
    Comm *c = ix->comm;
    if (!c || !c->attached) return fail(MMIDX_ERR_STATE, "no attached communicator: mmidx_comm_create + mmidx_comm_attach first");
    if (ix->has_rot) return fail(MMIDX_ERR_UNSUPPORTED, "RandomRotation is not available in the multi-GPU step (its tie pass evaluates residuals without tables)");
    RET(validate_search(ix, gq, k, w_out));
    if (gq < 1 || gq > c->max_gq) return fail(MMIDX_ERR_INVALID, "gq = %lld outside 1..%lld (the window was sized by mmidx_comm_create)", (long long)gq, (long long)c->max_gq);
    if (k > c->k_max) return fail(MMIDX_ERR_INVALID, "k = %d exceeds the window's k_max = %d", k, c->k_max);
    if (*w_out > c->w_max) return fail(MMIDX_ERR_UNSUPPORTED, "w = %d exceeds %d", *w_out, c->w_max);
    return MMIDX_OK;
}

extern "C" int mmidx_search_multi_dev(mmidx_t *ix, int64_t gq, const double *dQ, int32_t k, int32_t gather_all,
                                      const int32_t **d_iids, const double **d_dist, const int32_t **d_count, int64_t *row0,
                                      int64_t *nrows, void *stream) {
    if (!ix || !dQ) return fail(MMIDX_ERR_INVALID, "null argument");
    DeviceGuard g(ix->device);
    int w = 0;
    RET(multi_validate(ix, gq, k, &w));
    RET(prepare_search(ix, k));
    Comm &c = *ix->comm;
    cudaStream_t st = (cudaStream_t)stream;
    const int par = (int)(c.calls & 1);
    GraphKey key{1, gq, k, w, gather_all ? 1 : 0, dQ, nullptr, nullptr, nullptr, st};
    RET(run_graphed(ix, key, par, st, [&](int *launches) { return multi_enqueue(ix, gq, dQ, k, w, gather_all != 0, par, st, launches); }));
    c.calls++;
    const int64_t sl = (gq + c.S - 1) / c.S, gqp = sl * c.S;
    unsigned char *mine = c.win + c.L.data0 + (size_t)par * c.L.parity_stride;
    if (d_iids) *d_iids = reinterpret_cast<const int32_t *>(mine + c.L.o_iids);
    if (d_dist) *d_dist = reinterpret_cast<const double *>(mine + c.L.o_dist);
    if (d_count) *d_count = reinterpret_cast<const int32_t *>(mine + c.L.o_cnt);
    const int64_t q0 = std::min<int64_t>(gq, (int64_t)c.shard * sl);
    if (row0) *row0 = (int64_t)c.group * gqp + (c.S > 1 ? (int64_t)c.shard * sl : 0);
    if (nrows) *nrows = c.S > 1 ? std::max<int64_t>(0, std::min<int64_t>(gq, q0 + sl) - q0) : gq;
    return MMIDX_OK;
}

extern "C" int mmidx_search_multi(mmidx_t *ix, int64_t gq, const double *Q, int32_t k, int32_t *out_iids, double *out_dist,
                                  int32_t *out_count, int64_t *first_query, int64_t *nrows) {
    if (!ix || !Q || !out_iids || !out_dist) return fail(MMIDX_ERR_INVALID, "null argument");
    DeviceGuard g(ix->device);
    int w = 0;
    RET(multi_validate(ix, gq, k, &w));
    cudaStream_t st = ix->stream;
    Scratch sc(st, ix->arenas, ix->arena_mu);
    double *dQ;
    RET(sc.get(&dQ, (size_t)gq * ix->p.d));
    CK(cudaMemcpyAsync(dQ, Q, sizeof(double) * (size_t)gq * ix->p.d, cudaMemcpyHostToDevice, st));
    const int32_t *di;
    const double *dd;
    const int32_t *dc;
    int64_t r0 = 0, nr = 0;
    RET(mmidx_search_multi_dev(ix, gq, dQ, k, 0, &di, &dd, &dc, &r0, &nr, st));
    if (nr > 0) {
        CK(cudaMemcpyAsync(out_iids, di + r0 * k, sizeof(int32_t) * (size_t)nr * k, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(out_dist, dd + r0 * k, sizeof(double) * (size_t)nr * k, cudaMemcpyDeviceToHost, st));
        if (out_count) CK(cudaMemcpyAsync(out_count, dc + r0, sizeof(int32_t) * (size_t)nr, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    Comm &c = *ix->comm;
    const int64_t sl = (gq + c.S - 1) / c.S;
    if (first_query) *first_query = c.S > 1 ? std::min<int64_t>(gq, (int64_t)c.shard * sl) : 0;
    if (nrows) *nrows = nr;
    return MMIDX_OK;
}

// ---------------------------------------------------------------------------------------------------------
// VLAD (K7)
// ---------------------------------------------------------------------------------------------------------
// one vocabulary: out rows `ld` doubles apart (the caller offsets d_out to the vocabulary's first column)
static int vlad_dev_impl(const double *d_codebook, int32_t K, int32_t D, int64_t n_img, const int64_t *d_offsets, int64_t n_desc,
                         const double *d_desc, double *d_out, int64_t ld, int32_t *d_assign, cudaStream_t st) {
    if (K < 1 || D < 1 || n_img < 0 || n_desc < 0) return fail(MMIDX_ERR_INVALID, "bad VLAD geometry");
    if (n_img == 0) return MMIDX_OK;
    if (!d_codebook || !d_offsets || !d_out || (n_desc > 0 && !d_desc)) return fail(MMIDX_ERR_INVALID, "null argument");
    Scratch sc(st);
    double *dBt;
    int32_t *assign = d_assign, *order, *cstart;
    RET(sc.get(&dBt, (size_t)K * D));
    if (!assign) RET(sc.get(&assign, (size_t)n_desc));
    RET(sc.get(&order, (size_t)n_desc));
    RET(sc.get(&cstart, (size_t)n_img * (K + 1)));
    if (env_exact() || n_desc == 0) {
        k_transpose<<<(unsigned)(((size_t)K * D + 255) / 256), 256, 0, st>>>(d_codebook, K, D, dBt);
        RET(post_launch("k_transpose", nullptr));
        RET(launch_assign(d_desc, dBt, n_desc, K, D, assign, st, nullptr));
    } else {
        // computeNearestCentroid through the fp32 filter (argmin_filter.cuh): tables of this codebook, filter, exact leftovers
        float *B32, *b2, *bmax;
        int64_t *amb_list;
        int32_t *amb_count;
        RET(sc.get(&B32, (size_t)K * D));
        RET(sc.get(&b2, (size_t)K));
        RET(sc.get(&bmax, 1));
        RET(sc.get(&amb_list, (size_t)n_desc));
        RET(sc.get(&amb_count, 1));
        CK(cudaMemsetAsync(bmax, 0, sizeof(float), st));
        CK(cudaMemsetAsync(amb_count, 0, sizeof(int32_t), st));
        k_argmin_tables<<<dim3(K, 1), MMIDX_NT, 0, st>>>(d_codebook, K, D, B32, b2, bmax);
        RET(post_launch("k_argmin_tables", nullptr));
        ArgminRows rows{d_desc, D, nullptr, nullptr, 0, nullptr, 0};
        for (int64_t v0 = 0; v0 < n_desc; v0 += (int64_t)AF_TV * 2000000) {  // grid.x limit
            const int64_t nb = std::min<int64_t>((int64_t)AF_TV * 2000000, n_desc - v0);
            ArgminRows rc = rows;
            rc.X = d_desc + v0 * D;
            k_argmin_filter<<<dim3((unsigned)((nb + AF_TV - 1) / AF_TV), 1), MMIDX_NT, 0, st>>>(rc, B32, b2, bmax, nb, K, D, 1, nullptr, nullptr,
                                                                                              assign + v0, 1, amb_list, amb_count);
            RET(post_launch("k_argmin_filter", nullptr));
            k_argmin_exact_list<<<2 * current_sm_count(), MMIDX_NT, 0, st>>>(rc, d_codebook, K, D, 1, amb_list, amb_count, nullptr, nullptr,
                                                                             assign + v0, 1);
            RET(post_launch("k_argmin_exact_list", nullptr));
            if (v0 + nb < n_desc) CK(cudaMemsetAsync(amb_count, 0, sizeof(int32_t), st));
        }
    }
    size_t smem = sizeof(int) * (size_t)(K + 1);
    if (smem > 48 * 1024) RET(set_smem(k_vlad_order, smem));
    for (int64_t i0 = 0; i0 < n_img; i0 += 1 << 30) {
        int64_t nb = std::min<int64_t>(1 << 30, n_img - i0);
        k_vlad_order<<<(unsigned)nb, MMIDX_NT, smem, st>>>(assign, d_offsets + i0, K, order, cstart + i0 * (K + 1));
        RET(post_launch("k_vlad_order", nullptr));
        k_vlad_accumulate<<<(unsigned)nb, MMIDX_NT, 0, st>>>(d_codebook, d_desc, d_offsets + i0, order, cstart + i0 * (K + 1), K, D,
                                                            d_out + i0 * ld, ld);
        RET(post_launch("k_vlad_accumulate", nullptr));
    }
    return MMIDX_OK;
}

extern "C" int mmidx_vlad_dev(const double *d_codebook, int32_t K, int32_t D, int64_t n_img, const int64_t *d_offsets,
                              int64_t n_desc, const double *d_desc, double *d_out, int32_t *d_assign, void *stream) {
    return vlad_dev_impl(d_codebook, K, D, n_img, d_offsets, n_desc, d_desc, d_out, (int64_t)K * D, d_assign, (cudaStream_t)stream);
}

extern "C" int mmidx_normalize_rows_dev(double *dX, int64_t rows, int64_t ld, int32_t len, int32_t do_power, double a,
                                        int32_t do_l2, void *stream) {
    if (rows < 0 || len < 1 || ld < len) return fail(MMIDX_ERR_INVALID, "bad row geometry");
    if (rows == 0 || (!do_power && !do_l2)) return MMIDX_OK;
    if (!dX) return fail(MMIDX_ERR_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    for (int64_t r0 = 0; r0 < rows; r0 += 1 << 30) {
        const int64_t nb = std::min<int64_t>(1 << 30, rows - r0);
        k_rows_normalize<<<(unsigned)nb, MMIDX_NT, 0, st>>>(dX + r0 * ld, ld, len, do_power, a, do_l2);
        RET(post_launch("k_rows_normalize", nullptr));
    }
    return MMIDX_OK;
}

// VladAggregatorMultipleVocabularies.aggregate (VAMV.java:84-101): every vocabulary aggregates the SAME descriptors;
// power(0.5) + L2 per sub-VLAD, concatenation, one more L2 over the whole vector when there are several vocabularies
extern "C" int mmidx_vlad_multi_dev(const double *d_codebooks, int32_t nvoc, const int32_t *Ks, int32_t D, int64_t n_img,
                                    const int64_t *d_offsets, int64_t n_desc, const double *d_desc, int32_t normalize,
                                    double *d_out, void *stream) {
    if (nvoc < 1 || !Ks) return fail(MMIDX_ERR_INVALID, "bad vocabulary list");
    int64_t total = 0;
    for (int v = 0; v < nvoc; ++v) {
        if (Ks[v] < 1) return fail(MMIDX_ERR_INVALID, "vocabulary %d has no centroids", v);
        total += (int64_t)Ks[v] * D;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int64_t col = 0, krow = 0;
    for (int v = 0; v < nvoc; ++v) {
        RET(vlad_dev_impl(d_codebooks + krow * D, Ks[v], D, n_img, d_offsets, n_desc, d_desc, d_out + col, total, nullptr, st));
        if (normalize) RET(mmidx_normalize_rows_dev(d_out + col, n_img, total, Ks[v] * D, 1, 0.5, 1, st));
        col += (int64_t)Ks[v] * D;
        krow += Ks[v];
    }
    if (normalize && nvoc > 1) RET(mmidx_normalize_rows_dev(d_out, n_img, total, (int32_t)total, 0, 0.0, 1, st));
    return MMIDX_OK;
}

extern "C" int mmidx_pca_project_dev(const double *d_Vt, const double *d_means, int32_t nc, int32_t ss, int64_t n,
                                     const double *dX, int32_t l2_normalize, double *d_out, void *stream) {
    if (nc < 1 || ss < 1 || n < 0) return fail(MMIDX_ERR_INVALID, "bad PCA geometry");
    if (n == 0) return MMIDX_OK;
    if (!d_Vt || !d_means || !dX || !d_out) return fail(MMIDX_ERR_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    for (int64_t r0 = 0; r0 < n; r0 += (int64_t)PCA_T * 65535) {
        const int64_t nb = std::min<int64_t>((int64_t)PCA_T * 65535, n - r0);
        dim3 grid((unsigned)((nc + PCA_T - 1) / PCA_T), (unsigned)((nb + PCA_T - 1) / PCA_T));
        k_pca_project<<<grid, MMIDX_NT, 0, st>>>(dX + r0 * ss, d_means, d_Vt, nb, nc, ss, d_out + r0 * nc);
        RET(post_launch("k_pca_project", nullptr));
    }
    if (l2_normalize) RET(mmidx_normalize_rows_dev(d_out, n, nc, nc, 0, 0.0, 1, st));  // PCA.java:203-204 (whitening)
    return MMIDX_OK;
}

// host-buffer wrappers of the three entry points above: copy in, run, copy out, synchronise
struct HostCall {
    cudaStream_t st = nullptr;
    int rc = MMIDX_OK;
    explicit HostCall(int32_t &device) {
        if (device < 0 && cudaGetDevice(&device) != cudaSuccess) device = 0;
        rc = check_device(device);
    }
};

extern "C" int mmidx_normalize_rows(double *X, int64_t rows, int64_t len, int32_t do_power, double a, int32_t do_l2, int32_t device) {
    if (rows < 0 || len < 1 || len > INT32_MAX) return fail(MMIDX_ERR_INVALID, "bad row geometry");
    if (rows == 0) return MMIDX_OK;
    if (!X) return fail(MMIDX_ERR_INVALID, "null argument");
    HostCall hc(device);
    RET(hc.rc);
    DeviceGuard g(device);
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int rc;
    {
        Scratch sc(st);
        auto body = [&]() -> int {
            double *dX;
            RET(sc.get(&dX, (size_t)rows * len));
            CK(cudaMemcpyAsync(dX, X, sizeof(double) * (size_t)rows * len, cudaMemcpyHostToDevice, st));
            RET(mmidx_normalize_rows_dev(dX, rows, len, (int32_t)len, do_power, a, do_l2, st));
            CK(cudaMemcpyAsync(X, dX, sizeof(double) * (size_t)rows * len, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            return MMIDX_OK;
        };
        rc = body();
    }
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return rc;
}

extern "C" int mmidx_vlad_multi(const double *codebooks, int32_t nvoc, const int32_t *Ks, int32_t D, int64_t n_img,
                                const int64_t *offsets, const double *desc, int32_t normalize, double *out, int32_t device) {
    if (nvoc < 1 || !Ks || D < 1 || n_img < 0) return fail(MMIDX_ERR_INVALID, "bad VLAD geometry");
    if (n_img == 0) return MMIDX_OK;
    if (!codebooks || !offsets || !out) return fail(MMIDX_ERR_INVALID, "null argument");
    for (int64_t i = 0; i < n_img; ++i)
        if (offsets[i + 1] < offsets[i]) return fail(MMIDX_ERR_INVALID, "offsets must be non-decreasing");
    if (offsets[0] != 0) return fail(MMIDX_ERR_INVALID, "offsets[0] must be 0");
    const int64_t n_desc = offsets[n_img];
    if (n_desc > 0 && !desc) return fail(MMIDX_ERR_INVALID, "null argument");
    int64_t ksum = 0;
    for (int v = 0; v < nvoc; ++v) ksum += Ks[v] > 0 ? Ks[v] : 0;
    HostCall hc(device);
    RET(hc.rc);
    DeviceGuard g(device);
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int rc;
    {
        Scratch sc(st);
        auto body = [&]() -> int {
            double *dcb, *ddesc, *dout;
            int64_t *doff;
            RET(sc.get(&dcb, (size_t)ksum * D));
            RET(sc.get(&ddesc, (size_t)std::max<int64_t>(n_desc, 1) * D));
            RET(sc.get(&dout, (size_t)n_img * ksum * D));
            RET(sc.get(&doff, (size_t)n_img + 1));
            CK(cudaMemcpyAsync(dcb, codebooks, sizeof(double) * (size_t)ksum * D, cudaMemcpyHostToDevice, st));
            if (n_desc) CK(cudaMemcpyAsync(ddesc, desc, sizeof(double) * (size_t)n_desc * D, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(doff, offsets, sizeof(int64_t) * (size_t)(n_img + 1), cudaMemcpyHostToDevice, st));
            RET(mmidx_vlad_multi_dev(dcb, nvoc, Ks, D, n_img, doff, n_desc, ddesc, normalize, dout, st));
            CK(cudaMemcpyAsync(out, dout, sizeof(double) * (size_t)n_img * ksum * D, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            return MMIDX_OK;
        };
        rc = body();
    }
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return rc;
}

extern "C" int mmidx_pca_project(const double *Vt, const double *means, int32_t nc, int32_t ss, int64_t n, const double *X,
                                 int32_t l2_normalize, double *out, int32_t device) {
    if (nc < 1 || ss < 1 || n < 0) return fail(MMIDX_ERR_INVALID, "bad PCA geometry");
    if (n == 0) return MMIDX_OK;
    if (!Vt || !means || !X || !out) return fail(MMIDX_ERR_INVALID, "null argument");
    HostCall hc(device);
    RET(hc.rc);
    DeviceGuard g(device);
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int rc;
    {
        Scratch sc(st);
        auto body = [&]() -> int {
            double *dV, *dm, *dX, *dY;
            RET(sc.get(&dV, (size_t)nc * ss));
            RET(sc.get(&dm, (size_t)ss));
            RET(sc.get(&dX, (size_t)n * ss));
            RET(sc.get(&dY, (size_t)n * nc));
            CK(cudaMemcpyAsync(dV, Vt, sizeof(double) * (size_t)nc * ss, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(dm, means, sizeof(double) * (size_t)ss, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(dX, X, sizeof(double) * (size_t)n * ss, cudaMemcpyHostToDevice, st));
            RET(mmidx_pca_project_dev(dV, dm, nc, ss, n, dX, l2_normalize, dY, st));
            CK(cudaMemcpyAsync(out, dY, sizeof(double) * (size_t)n * nc, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            return MMIDX_OK;
        };
        rc = body();
    }
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return rc;
}

extern "C" int mmidx_vlad(const double *codebook, int32_t K, int32_t D, int64_t n_img, const int64_t *offsets,
                          const double *desc, double *out, int32_t *out_assign, int32_t device) {
    if (K < 1 || D < 1 || n_img < 0) return fail(MMIDX_ERR_INVALID, "bad VLAD geometry");
    if (n_img == 0) return MMIDX_OK;
    if (!codebook || !offsets || !out) return fail(MMIDX_ERR_INVALID, "null argument");
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    RET(check_device(device));
    DeviceGuard g(device);
    for (int64_t i = 0; i < n_img; ++i)
        if (offsets[i + 1] < offsets[i]) return fail(MMIDX_ERR_INVALID, "offsets must be non-decreasing");
    if (offsets[0] != 0) return fail(MMIDX_ERR_INVALID, "offsets[0] must be 0");
    const int64_t n_desc = offsets[n_img];
    if (n_desc > 0 && !desc) return fail(MMIDX_ERR_INVALID, "null argument");
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int rc = MMIDX_OK;
    {
        Scratch sc(st);
        double *dcb, *ddesc, *dout;
        int64_t *doff;
        int32_t *dassign;
        auto body = [&]() -> int {
            RET(sc.get(&dcb, (size_t)K * D));
            RET(sc.get(&ddesc, (size_t)std::max<int64_t>(n_desc, 1) * D));
            RET(sc.get(&dout, (size_t)n_img * K * D));
            RET(sc.get(&doff, (size_t)n_img + 1));
            RET(sc.get(&dassign, (size_t)std::max<int64_t>(n_desc, 1)));
            CK(cudaMemcpyAsync(dcb, codebook, sizeof(double) * (size_t)K * D, cudaMemcpyHostToDevice, st));
            if (n_desc) CK(cudaMemcpyAsync(ddesc, desc, sizeof(double) * (size_t)n_desc * D, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(doff, offsets, sizeof(int64_t) * (size_t)(n_img + 1), cudaMemcpyHostToDevice, st));
            RET(mmidx_vlad_dev(dcb, K, D, n_img, doff, n_desc, ddesc, dout, dassign, st));
            CK(cudaMemcpyAsync(out, dout, sizeof(double) * (size_t)n_img * K * D, cudaMemcpyDeviceToHost, st));
            if (out_assign && n_desc)
                CK(cudaMemcpyAsync(out_assign, dassign, sizeof(int32_t) * (size_t)n_desc, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            return MMIDX_OK;
        };
        rc = body();
    }
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return rc;
}
