// mmsig_api.cu -- the C ABI of libmmsig.so (include/mmsig.h): handle, device memory, launch
// plans, NCCL exchange, fit loops.  Host side of the MMCTM / CTM / LDA variational-EM path.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <sched.h>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/mmsig.h"
#include "det_math.cuh"
#include "mmctm_kernels.cuh"
#include "theta_tile.cuh"
#include "mmctm_wide.cuh"
#include "mmctm_pack.cuh"
#include "mmctm_lean.cuh"
#include "elbo_kernels.cuh"
#include "lda_kernels.cuh"
#include "lda_tile.cuh"
#include "tile_f32.cuh"
#include "ingest_kernels.cuh"

using namespace mmsig;

// ---- NCCL, bound at run time (dlopen) so that a process that already holds torch's libnccl
// shares it and single-GPU users need no NCCL at all ---------------------------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef int (*pfn_ncclGetUniqueId)(ncclUniqueId_t *);
typedef int (*pfn_ncclCommInitRank)(ncclComm_t *, int, ncclUniqueId_t, int);
typedef int (*pfn_ncclCommDestroy)(ncclComm_t);
typedef int (*pfn_ncclAllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
typedef const char *(*pfn_ncclGetErrorString)(int);
static const int kNcclInt8 = 0, kNcclFloat64 = 8;

struct NcclApi {
    void *lib = nullptr;
    pfn_ncclGetUniqueId GetUniqueId = nullptr;
    pfn_ncclCommInitRank CommInitRank = nullptr;
    pfn_ncclCommDestroy CommDestroy = nullptr;
    pfn_ncclAllGather AllGather = nullptr;
    pfn_ncclGetErrorString GetErrorString = nullptr;
};
static NcclApi g_nccl;
static std::string g_last_error;

static bool load_nccl(std::string &err) {
    if (g_nccl.lib) return true;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { err = std::string("cannot dlopen libnccl: ") + dlerror(); return false; }
    g_nccl.GetUniqueId = (pfn_ncclGetUniqueId)dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (pfn_ncclCommInitRank)dlsym(lib, "ncclCommInitRank");
    g_nccl.CommDestroy = (pfn_ncclCommDestroy)dlsym(lib, "ncclCommDestroy");
    g_nccl.AllGather = (pfn_ncclAllGather)dlsym(lib, "ncclAllGather");
    g_nccl.GetErrorString = (pfn_ncclGetErrorString)dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy) {
        err = "libnccl lacks a required symbol";
        return false;
    }
    g_nccl.lib = lib;
    return true;
}

// ---- handle -------------------------------------------------------------------------------
struct KernelTime { const char *name; double ms; int64_t n; };
struct PendingEvent { int idx; cudaEvent_t a, b; };

// device-side storage of one modality's counts; reused across set_data calls of the same shape
struct CountBuf {
    long long D = -1, nnz = -1;
    long long *rowptr = nullptr;
    int2 *rec = nullptr;
    int *term = nullptr, *count = nullptr;      // staging for the packing kernel
    int *flags = nullptr;                       // [0] flags, [2..3] 8 bytes: sum of counts
    int *cnt = nullptr;                         // dense tiles [ceil(D / 32) * 32][V] when the modality is dense (tile_stage.cuh)
    int V = 0;
};

struct MmctmHost {
    bool has_data = false, has_state = false, estep_done = false, last_unsmoothed = false;
    CountBuf cb[MAXM];
    double *props_scratch = nullptr;
    MmctmDev p{};
    int G = 0;
    std::vector<long long> nnz;
    int grid_theta[MAXM] = {0}, W_theta[MAXM] = {0}, grid_ll[MAXM] = {0};
    size_t smem_theta[MAXM] = {0}, smem_ll[MAXM] = {0};
    int grid_solve = 0, grid_post = 0, grid_mom = 0, grid_zeta = 0;
    int solve_lean = 0;                // sum(K) <= 32: k_solve_lean, one kernel per LD_MMA phase, G = 4 or 8 lanes per sample (mmctm_lean.cuh); 0: one coordinate per lane (k_solve / k_solve_pack)
    bool wide = false;                 // 32 < sum(K) <= 64: two coordinates per lane (mmctm_wide.cuh)
    size_t smem_solve = 0;
    double2 *part_mom = nullptr;
    size_t smem_post = 0;
    double2 *part_theta[MAXM] = {nullptr};
    double2 *part_solve = nullptr, *part_post = nullptr, *part_elbo = nullptr;
    double2 *rank_p1 = nullptr, *gath_p1 = nullptr, *rank_p2 = nullptr, *gath_p2 = nullptr;
    double *d_ll = nullptr;
    int *d_status = nullptr;
    double *lamA = nullptr, *lamB = nullptr;
    double *sumtheta_alt = nullptr;     // second sumθ buffer: the θ pass of iteration t+1 may run (side-stream overlap) before
                                        // iteration t's stopping rule is known; it must not overwrite the sumθ the ELBO reads
    int *d_ctl = nullptr;               // MmctmDev::ctl
    double *d_llprev = nullptr;         // the previous iteration's log-likelihoods, for the stopping rule on the device
    int cur_iter = 0;                   // > 0 while mmctm_run_iterations enqueues iteration cur_iter of a batch
    double cur_tol = 0.0;
    double *snap[16] = {nullptr};       // best-restart snapshot of mmsig_mmctm_restarts, kept with the plan
    std::vector<double> alphaf_host;    // IMMCTM: per-(modality, feature) alpha
    std::vector<int> row_len_host, row_m_host;   // IMMCTM: J and modality of every feature-table row
    std::vector<double> alpha_host;
};

struct LdaHost {
    bool has_data = false, has_state = false, iterated = false, last_unsmoothed = false, last_frozen = false;
    CountBuf cb;
    LdaDev p{};
    long long nnz = 0;
    int grid = 0, W = 0, grid_ll = 0, grid_row = 0, grid_tile = 0, NW = 0, grid_t32 = 0, grid_llt = 0;
    bool tile = false, t32 = false;
    size_t smem = 0, smem_ll = 0, smem_elbo = 0, smem_tile = 0, smem_t32 = 0, smem_llt = 0;
    double2 *part = nullptr, *part_ll = nullptr, *rank_p = nullptr, *gath_p = nullptr, *rank_ll = nullptr, *gath_ll = nullptr;
    double *d_ll = nullptr;
    double *gamA = nullptr, *gamB = nullptr;
    std::vector<int> J_host;            // ILDA: values per feature
    std::vector<double> etaf_host;      // ILDA: eta per feature
};

// One process, several GPUs (include/mmsig.h "mmsig_group"): a group owns one handle per device and drives
// each from its own host thread.  Handles of a group exchange their packed partial sums through peer memory:
// a rank pushes its buffer into a slot of every member's exchange arena with plain stores over NVLink
// (k_group_push), records an event, and every member's stream waits for every other member's event.  No
// NCCL, no spin-waits; the only host-side coupling is a pthread barrier (events must be recorded before a
// peer can wait on them).  The arena is double-buffered by call parity, so a slot is rewritten only after
// the exchange that followed its readers.
struct mmsig_group {
    int n = 0;
    std::vector<mmsig_handle *> h;
    std::vector<int> devices;
    struct GroupBarrier *bar = nullptr;
    std::vector<double2 *> arena;               // [rank] current exchange arena of that member (device memory of its GPU)
    std::vector<size_t> arena_cap;              // [rank] capacity in double2
    std::vector<cudaEvent_t> ev;                // [rank * 2 + parity]
    std::vector<long long> scratch;             // [rank * MAXM8 + i] host-side all-sum
    std::vector<int> status;                    // [rank] result of the member's part of a group call
    std::vector<long long> cut;                 // row boundaries of the shards of the resident corpus
    int replica_best = -1;                      // after mmsig_group_mmctm_restarts: the member that holds the best restart
    std::string err;
};

struct mmsig_handle {
    mmsig_group *grp = nullptr;                        // member of a single-process multi-GPU group
    int xparity = 0;                                   // exchange arena buffer of the next group gather
    std::vector<void *> old_arenas;                    // outgrown arenas, freed with the handle
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;      // copy streams of mmsig_mmctm_fit_host
    // second half of the M-step (moments, log-likelihood pass, their exchange, Σ / invΣ) of iteration t on a side stream,
    // concurrent with the θ pass of iteration t + 1 (which needs none of its results) inside the sync-free loop of fit
    cudaStream_t s_aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool aux_pending = false;                          // the solve of the next E-step has to wait for ev_join
    long long *allsum_buf = nullptr;                   // scratch of allsum_ll (multi-rank totals)
    double *ll_pinned = nullptr;                       // page-locked landing zone of the log-likelihoods of the sync-free iterations
    int2 *fmt_rec = nullptr;                           // records of the last mmsig_format_counts
    long long fmt_nnz = -1;
    bool own_stream = false;
    std::string err;
    int stop_rule = 0;
    int precision = 0;                                 // MMSIG_PRECISION_*: 1 = the tile passes in float (tile_f32.cuh)
    bool profile = false;
    int numSM = 0;
    size_t smem_optin = 0;
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    int64_t launches = 0;
    std::vector<KernelTime> kt;
    std::vector<PendingEvent> pending;
    std::vector<void *> allocs_mm, allocs_lda;
    MmctmHost mm;
    LdaHost lda;
};

static int fail(mmsig_handle *h, int code, const std::string &msg) {
    if (h) h->err = msg;
    g_last_error = msg;
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(h, MMSIG_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));       \
    } while (0)
#define NEED(cond, msg)                                                                            \
    do {                                                                                           \
        if (!(cond)) return fail(h, MMSIG_EINVAL, msg);                                            \
    } while (0)

template <typename T>
static int dev_alloc(mmsig_handle *h, std::vector<void *> &pool, T **out, size_t count) {
    void *ptr = nullptr;
    cudaError_t e = cudaMalloc(&ptr, std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) return fail(h, MMSIG_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    pool.push_back(ptr);
    *out = (T *)ptr;
    return 0;
}
static void free_pool(std::vector<void *> &pool) {
    for (void *p : pool) cudaFree(p);
    pool.clear();
}

// kernel launch bookkeeping: count, optional event timing
static int kt_index(mmsig_handle *h, const char *name) {
    for (size_t i = 0; i < h->kt.size(); ++i)
        if (h->kt[i].name == name || !strcmp(h->kt[i].name, name)) return (int)i;
    h->kt.push_back({name, 0.0, 0});
    return (int)h->kt.size() - 1;
}
struct LaunchScope {
    mmsig_handle *h;
    PendingEvent pe{};
    bool timed;
    LaunchScope(mmsig_handle *h_, const char *name) : h(h_), timed(h_->profile) {
        h->launches++;
        if (timed) {
            pe.idx = kt_index(h, name);
            cudaEventCreate(&pe.a);
            cudaEventCreate(&pe.b);
            cudaEventRecord(pe.a, h->stream);
        }
    }
    ~LaunchScope() {
        if (timed) {
            cudaEventRecord(pe.b, h->stream);
            h->pending.push_back(pe);
        }
    }
};
static void resolve_pending(mmsig_handle *h) {
    for (auto &pe : h->pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, pe.a, pe.b) == cudaSuccess) {
            h->kt[pe.idx].ms += ms;
            h->kt[pe.idx].n += 1;
        }
        cudaEventDestroy(pe.a);
        cudaEventDestroy(pe.b);
    }
    h->pending.clear();
}

// ---- exchange inside a single-process group (peer memory) ------------------------------------
// spin-then-yield barrier over the group's host threads that can be aborted: a member that fails
// releases the others with an error instead of leaving them waiting
struct GroupBarrier {
    std::atomic<int> count{0}, gen{0};
    std::atomic<bool> aborted{false};
    int n = 1;
    bool wait() {
        if (aborted.load(std::memory_order_acquire)) return false;
        const int g = gen.load(std::memory_order_acquire);
        if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == n) {
            count.store(0, std::memory_order_relaxed);
            gen.fetch_add(1, std::memory_order_acq_rel);
            return true;
        }
        for (int spin = 0; gen.load(std::memory_order_acquire) == g; ++spin) {
            if (aborted.load(std::memory_order_acquire)) return false;
            if (spin > 4000) sched_yield();
        }
        return !aborted.load(std::memory_order_acquire);
    }
    void abort() { aborted.store(true, std::memory_order_release); }
    void reset(int n_) { n = n_; count.store(0); aborted.store(false); }
};

struct PeerSlots { double2 *p[16]; int n; };
__global__ void k_group_push(const double2 *__restrict__ src, size_t n, PeerSlots dst) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double2 v = src[i];
        for (int r = 0; r < dst.n; ++r) dst.p[r][i] = v;          // peer stores over NVLink (own slot: local)
    }
}

static int group_gather(mmsig_handle *h, const double2 *rank_buf, size_t n, const double2 **out);

static int gather(mmsig_handle *h, const double2 *rank_buf, double2 *gath_buf, size_t n, const double2 **out, cudaStream_t st = nullptr) {
    if (h->nranks == 1) { *out = rank_buf; return 0; }
    if (h->grp) return group_gather(h, rank_buf, n, out);
    LaunchScope ls(h, "ncclAllGather");
    int rc = g_nccl.AllGather(rank_buf, gath_buf, n * 2, kNcclFloat64, h->comm, st ? st : h->stream);
    if (rc != 0)
        return fail(h, MMSIG_ENCCL, std::string("ncclAllGather: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
    *out = gath_buf;
    return 0;
}


// opt a kernel in to the largest dynamic shared memory the device allows (minus its static part)
template <typename F>
static cudaError_t allow_max_smem(mmsig_handle *h, F kernel) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(h->smem_optin - fa.sharedSizeBytes));
}

// ---- generic ------------------------------------------------------------------------------
extern "C" int32_t mmsig_version(void) { return 111; }     // 111: mmsig_ilda_*

extern "C" int32_t mmsig_limits(int32_t *max_modalities, int32_t *max_sum_K, int32_t *max_K, int32_t *max_V_mmctm,
                                int32_t *max_V_lda) {
    if (max_modalities) *max_modalities = MAXM;
    if (max_sum_K) *max_sum_K = MAXMK;
    if (max_K) *max_K = 32;
    if (max_V_mmctm) *max_V_mmctm = 1024;
    if (max_V_lda) *max_V_lda = 65535;
    return 0;
}

extern "C" const char *mmsig_last_error(const mmsig_handle *h) { return h ? h->err.c_str() : g_last_error.c_str(); }

extern "C" int32_t mmsig_create(const mmsig_config *cfg, mmsig_handle **out) {
    mmsig_handle *h = nullptr;
    if (!cfg || !out) return fail(nullptr, MMSIG_EINVAL, "mmsig_create: null argument");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, MMSIG_ENODEV, std::string("no CUDA device (there is no CPU fallback): ") +
                                               (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0"));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, MMSIG_EINVAL, "mmsig_create: bad device ordinal");
    if (cfg->precision != MMSIG_PRECISION_FP64 && cfg->precision != MMSIG_PRECISION_FP32)
        return fail(nullptr, MMSIG_EINVAL, "mmsig_create: precision must be MMSIG_PRECISION_FP64 or MMSIG_PRECISION_FP32");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return fail(nullptr, MMSIG_ECUDA, "cudaGetDeviceProperties failed");
    if (prop.major < 10)
        return fail(nullptr, MMSIG_ENODEV, "device is not sm_100 class; libmmsig carries sm_100a code only");
    h = new mmsig_handle();
    h->device = cfg->device;
    h->stop_rule = cfg->stop_rule == MMSIG_STOP_NLOPT26 ? 1 : 0;
    h->precision = cfg->precision == MMSIG_PRECISION_FP32 ? 1 : 0;
    h->profile = cfg->profile != 0;
    h->numSM = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    if (cudaSetDevice(h->device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return fail(nullptr, MMSIG_ECUDA, "cannot create stream");
    }
    h->own_stream = true;
    // page-locked landing zone of the log-likelihoods of the sync-free iterations (allocated here: cudaHostAlloc
    // costs milliseconds and serialises across the host threads of a group)
    if (cudaHostAlloc((void **)&h->ll_pinned, (size_t)12 * (MAXM + 1) * sizeof(double), cudaHostAllocDefault) != cudaSuccess) {
        cudaStreamDestroy(h->stream);
        delete h;
        return fail(nullptr, MMSIG_ENOMEM, "cudaHostAlloc failed");
    }
    *out = h;
    return 0;
}

extern "C" int32_t mmsig_destroy(mmsig_handle *h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    resolve_pending(h);
    free_pool(h->allocs_mm);
    free_pool(h->allocs_lda);
    if (h->comm) g_nccl.CommDestroy(h->comm);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    cudaFree(h->fmt_rec);
    cudaFree(h->allsum_buf);
    if (h->ll_pinned) cudaFreeHost(h->ll_pinned);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    if (h->s_aux) cudaStreamDestroy(h->s_aux);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    delete h;
    return 0;
}

extern "C" int32_t mmsig_set_stream(mmsig_handle *h, void *cuda_stream) {
    NEED(h, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    if (h->own_stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)cuda_stream;
    h->own_stream = false;
    return 0;
}

extern "C" int32_t mmsig_synchronize(mmsig_handle *h) {
    NEED(h, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int32_t mmsig_set_profile(mmsig_handle *h, int32_t on) {
    NEED(h, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    resolve_pending(h);
    h->profile = on != 0;
    return 0;
}

// page-locked host memory for the caller's buffers: copies from / to it run at link speed and overlap
// with kernels (pageable buffers are staged by the driver, chunk by chunk)
extern "C" int32_t mmsig_host_alloc(uint64_t bytes, void **out) {
    mmsig_handle *h = nullptr;
    NEED(out, "null argument");
    void *p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, std::max<uint64_t>(bytes, 1), cudaHostAllocPortable);
    if (e != cudaSuccess) return fail(h, e == cudaErrorMemoryAllocation ? MMSIG_ENOMEM : MMSIG_ENODEV, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
    *out = p;
    return 0;
}
extern "C" int32_t mmsig_host_free(void *p) {
    mmsig_handle *h = nullptr;
    if (!p) return 0;
    CU(cudaFreeHost(p));
    return 0;
}

extern "C" int32_t mmsig_comm_unique_id(uint8_t id_out[128]) {
    mmsig_handle *h = nullptr;
    std::string err;
    if (!load_nccl(err)) return fail(h, MMSIG_ENCCL, err);
    ncclUniqueId_t id;
    int rc = g_nccl.GetUniqueId(&id);
    if (rc != 0) return fail(h, MMSIG_ENCCL, "ncclGetUniqueId failed");
    memcpy(id_out, id.internal, 128);
    return 0;
}

extern "C" int32_t mmsig_comm_init(mmsig_handle *h, const uint8_t id[128], int32_t rank, int32_t nranks) {
    NEED(h && id, "null argument");
    NEED(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
    NEED(!h->mm.has_data && !h->lda.has_data, "mmsig_comm_init must precede set_data");
    std::string err;
    if (!load_nccl(err)) return fail(h, MMSIG_ENCCL, err);
    CU(cudaSetDevice(h->device));
    ncclUniqueId_t uid;
    memcpy(uid.internal, id, 128);
    int rc = g_nccl.CommInitRank(&h->comm, nranks, uid, rank);
    if (rc != 0) return fail(h, MMSIG_ENCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
    h->rank = rank;
    h->nranks = nranks;
    return 0;
}

extern "C" int64_t mmsig_launch_count(const mmsig_handle *h) { return h ? h->launches : 0; }

extern "C" int32_t mmsig_kernel_times(mmsig_handle *h, int32_t n_max, const char **names_out, double *ms_total_out,
                                      int64_t *launches_out, int32_t reset) {
    NEED(h, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    resolve_pending(h);
    int n = std::min<int>(n_max, (int)h->kt.size());
    for (int i = 0; i < n; ++i) {
        if (names_out) names_out[i] = h->kt[i].name;
        if (ms_total_out) ms_total_out[i] = h->kt[i].ms;
        if (launches_out) launches_out[i] = h->kt[i].n;
    }
    if (reset)
        for (auto &k : h->kt) { k.ms = 0.0; k.n = 0; }
    return n;
}

// sum of per-rank int64 totals (Σ_d N_dm) over ranks, on the host
static int group_allsum(mmsig_handle *h, long long *vals, int n);
static int allsum_ll(mmsig_handle *h, long long *vals, int n) {
    if (h->nranks == 1) return 0;
    NEED(n <= MAXM, "allsum_ll: too many values");
    if (h->grp) return group_allsum(h, vals, n);
    if (!h->allsum_buf) CU(cudaMalloc(&h->allsum_buf, (size_t)MAXM * (h->nranks + 1) * sizeof(long long)));   // kept: no malloc / free per call
    long long *d_in = h->allsum_buf, *d_out = h->allsum_buf + MAXM;
    CU(cudaMemcpyAsync(d_in, vals, n * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    int rc = g_nccl.AllGather(d_in, d_out, (size_t)n * sizeof(long long), kNcclInt8, h->comm, h->stream);
    if (rc != 0) return fail(h, MMSIG_ENCCL, "ncclAllGather (totals) failed");
    std::vector<long long> all((size_t)n * h->nranks);
    CU(cudaMemcpyAsync(all.data(), d_out, all.size() * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < n; ++i) {
        long long s = 0;
        for (int r = 0; r < h->nranks; ++r) s += all[(size_t)r * n + i];
        vals[i] = s;
    }
    return 0;
}

// ---- count ingest: (term, count) -> packed records, row totals, validation ------------------
// flags: bit0 term out of range, bit1 count <= 0, bit2 terms of a row not strictly ascending
// tag != 0 (MMCTM): bits 16..20 of rec.x carry (tag_base + d) mod 32, the sample's slot in its
// 32-sample tile, so that the tile kernels scatter a record without searching the row pointers
// cnt != nullptr (a dense modality, tile_stage.cuh): the row is also written as V dense int32 cells
__global__ void k_pack_rows(const long long *rowptr, const int *term, const int *count, long long D, int V,
                            int2 *rec, double *N, int M, int m, int *flags, unsigned long long *ntot, int tag,
                            long long tag_base, int *cnt) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    int bad = 0;
    unsigned long long tot = 0;
    for (long long d = (long long)blockIdx.x * (blockDim.x >> 5) + warp; d < D; d += nw) {
        const long long beg = rowptr[d], end = rowptr[d + 1];
        long long s = 0;
        if (cnt) {
            for (int v = lane; v < V; v += 32) cnt[d * V + v] = 0;
            __syncwarp();
        }
        for (long long w = beg + lane; w < end; w += 32) {
            // count == nullptr: term[] holds packed records (bits 0-9 term, bits 10-31 count: mmsig_pack_records)
            int t = term[w], c;
            int tprev = w > beg ? term[w - 1] : -1;
            if (count) c = count[w];
            else {
                c = (int)((unsigned)t >> 10);
                t &= 1023;
                tprev = w > beg ? (tprev & 1023) : -1;
            }
            if (w > beg && tprev >= t) bad |= 4;
            // a bad entry is flagged AND neutralised (term 0, count 0): mmsig_mmctm_fit_host launches the
            // E-step of a chunk before the flags are read, and the tile kernels index shared memory with
            // the term (a term < 0 or >= 65536 would also alias into the slot-tag bits)
            if (t < 0 || t >= V) { bad |= 1; t = 0; c = 0; }
            if (c <= 0) { bad |= 2; c = 0; }
            rec[w] = make_int2(tag ? (t | (int)(((tag_base + d) & 31) << 16)) : t, c);
            if (cnt && c > 0) cnt[d * V + t] = c;
            s += c;
        }
        for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(FULLMASK, s, off);
        if (lane == 0) { N[d * M + m] = (double)s; tot += (unsigned long long)s; }
    }
    if (bad) atomicOr(flags, bad);
    if (lane == 0 && tot) atomicAdd(ntot, tot);
}

static int check_rowptr(mmsig_handle *h, const int64_t *rowptr, long long D) {
    NEED(rowptr && rowptr[0] == 0, "rowptr[0] must be 0");
    for (long long d = 0; d < D; ++d)
        if (rowptr[d + 1] < rowptr[d]) return fail(h, MMSIG_EINVAL, "rowptr not monotone");
    return 0;
}
// (re)allocate the device side of one modality's counts when its shape changed
// does a modality of this shape keep dense count tiles?  MMSIG_TILES=csr | dense overrides the density rule (A/B)
static bool want_dense_tiles(const mmsig_handle *h, long long D, long long nnz, int V) {
    const char *e = getenv("MMSIG_TILES");
    if (V < 1 || V > 1024) return false;
    if (h->precision) return true;                  // the FP32 kernels read dense tiles only
    if (e && !strcmp(e, "csr")) return false;
    if (e && !strcmp(e, "dense")) return true;
    return (double)nnz >= kDenseFrac * (double)D * (double)V;
}
// (re)allocate the device side of one modality's counts when its shape changed
static int ensure_countbuf(mmsig_handle *h, std::vector<void *> &pool, CountBuf &cb, long long D, long long nnz, int V = 0) {
    if (cb.D == D && cb.nnz == nnz && cb.V == V) return 0;
    int rc;
    if ((rc = dev_alloc(h, pool, &cb.rowptr, D + 1))) return rc;
    if ((rc = dev_alloc(h, pool, &cb.rec, nnz))) return rc;
    if ((rc = dev_alloc(h, pool, &cb.term, nnz))) return rc;
    if ((rc = dev_alloc(h, pool, &cb.count, nnz))) return rc;
    if ((rc = dev_alloc(h, pool, &cb.flags, (size_t)4))) return rc;
    cb.cnt = nullptr;
    if (V > 0 && want_dense_tiles(h, D, nnz, V)) {
        const size_t cells = (size_t)((D + 31) / 32) * 32 * V;
        if ((rc = dev_alloc(h, pool, &cb.cnt, cells))) return rc;
        CU(cudaMemsetAsync(cb.cnt, 0, cells * sizeof(int), h->stream));      // the padding rows of the last tile stay zero
    }
    cb.D = D;
    cb.nnz = nnz;
    cb.V = V;
    return 0;
}
// records of rows [d0, d1) -> the dense tiles of a dense modality
static void densify_launch(mmsig_handle *h, CountBuf &cb, long long d0, long long d1) {
    if (!cb.cnt || d1 <= d0) return;
    cudaMemsetAsync(cb.cnt + (size_t)d0 * cb.V, 0, (size_t)(d1 - d0) * cb.V * sizeof(int), h->stream);
    LaunchScope ls(h, "k_densify");
    const long long Dc = d1 - d0;
    const int grid = (int)std::max<long long>(1, std::min<long long>((Dc + 7) / 8, (long long)h->numSM * 8));
    k_densify<<<grid, 256, 0, h->stream>>>(cb.rowptr + d0, cb.rec, Dc, cb.V, cb.cnt + (size_t)d0 * cb.V);
}
// rows [d0, d1) of one modality: validation, (term, count) -> records, row totals
static void pack_launch(mmsig_handle *h, CountBuf &cb, long long d0, long long d1, int V, int M, int m, double *d_N,
                        int tag = 0, bool packed = false) {
    LaunchScope ls(h, "k_pack_rows");
    const long long Dc = d1 - d0;
    int grid = (int)std::min<long long>((Dc + 7) / 8, (long long)h->numSM * 8);
    k_pack_rows<<<std::max(grid, 1), 256, 0, h->stream>>>(cb.rowptr + d0, cb.term, packed ? nullptr : cb.count, Dc, V, cb.rec, d_N + d0 * M, M, m,
                                                           cb.flags, (unsigned long long *)(cb.flags + 2), tag, d0,
                                                           cb.cnt ? cb.cnt + (size_t)d0 * cb.V : nullptr);
}
// packing also writes the dense rows of a dense modality
static void pack_and_densify(mmsig_handle *h, CountBuf &cb, long long d0, long long d1, int V, int M, int m, double *d_N,
                             int tag = 0, bool packed = false) {
    pack_launch(h, cb, d0, d1, V, M, m, d_N, tag, packed);
}
static int flags_verdict(mmsig_handle *h, const int *hf, long long *ntot_out) {
    if (hf[0] & 1) return fail(h, MMSIG_EINVAL, "term index out of range [0, V)");
    if (hf[0] & 2) return fail(h, MMSIG_EINVAL, "count must be > 0 (zeros are dropped by format_counts_*)");
    if (hf[0] & 4) return fail(h, MMSIG_EINVAL, "terms of a row must be strictly ascending (as format_counts_* produces)");
    unsigned long long nt;
    memcpy(&nt, hf + 2, 8);
    *ntot_out = (long long)nt;
    return 0;
}

static int upload_counts(mmsig_handle *h, std::vector<void *> &pool, CountBuf &cb, long long D, int V, int M, int m,
                         const int64_t *rowptr, const int32_t *term, const int32_t *count, double *d_N,
                         long long *ntot_out, int tag = 0) {
    int rc;
    if ((rc = check_rowptr(h, rowptr, D))) return rc;
    const long long nnz = rowptr[D];
    NEED(nnz == 0 || (term && count), "null term / count");
    if ((rc = ensure_countbuf(h, pool, cb, D, nnz, tag ? V : 0))) return rc;
    CU(cudaMemsetAsync(cb.flags, 0, 4 * sizeof(int), h->stream));
    CU(cudaMemcpyAsync(cb.rowptr, rowptr, (D + 1) * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    if (nnz) {
        CU(cudaMemcpyAsync(cb.term, term, nnz * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(cb.count, count, nnz * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    }
    pack_and_densify(h, cb, 0, D, V, M, m, d_N, tag);
    int hf[4] = {0, 0, 0, 0};
    CU(cudaMemcpyAsync(hf, cb.flags, sizeof(hf), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return flags_verdict(h, hf, ntot_out);
}

// ===========================================================================================
// MMCTM
// ===========================================================================================
// launch plan of k_theta_tile for one modality: one thread per term, tiles of TILE_S samples
template <typename F>
static int pick_tile_plan(mmsig_handle *h, F kernel, int KP, bool ereg, bool dense, int V, long long D, int *grid_out, size_t *smem_out) {
    const int VP = V | 1, NW = (V + 31) / 32;
    const size_t base = (size_t)V * KP + (ereg ? 0 : (size_t)KP * VP) + (size_t)TILE_S * VP + (size_t)TILE_S * KP;
    const size_t smem = (dense ? tile_stage_offset(base) + tile_stage_doubles(V) : base + TILE_S + 2) * sizeof(double);
    if (smem > h->smem_optin) return fail(h, MMSIG_ELIMIT, "V too large for the theta tile (shared memory)");
    CU(allow_max_smem(h, kernel));
    int nb = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, NW * 32, smem));
    if (nb < 1) return fail(h, MMSIG_ELIMIT, "theta tile kernel does not fit on an SM for this K, V");
    *smem_out = smem;
    const long long ntiles = (D + TILE_S - 1) / TILE_S;
    *grid_out = (int)std::max<long long>(1, std::min<long long>((long long)h->numSM * nb, ntiles));
    return 0;
}

// launch plan of an FP32 tile kernel (tile_f32.cuh): one thread per term, tiles of TILE_S samples
template <typename F>
static int pick_f32_plan(mmsig_handle *h, F kernel, size_t smem, int V, long long D, int *grid_out, size_t *smem_out) {
    const int NW = (V + 31) / 32;
    if (smem > h->smem_optin) return fail(h, MMSIG_ELIMIT, "V too large for the FP32 tile kernels (shared memory)");
    CU(allow_max_smem(h, kernel));
    int nb = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, NW * 32, smem));
    if (nb < 1) return fail(h, MMSIG_ELIMIT, "FP32 tile kernel does not fit on an SM for this K, V");
    *smem_out = smem;
    const long long ntiles = (D + TILE_S - 1) / TILE_S;
    *grid_out = (int)std::max<long long>(1, std::min<long long>((long long)h->numSM * nb, ntiles));
    return 0;
}

#define THETA_DISPATCH(K, EXPR)                                   \
    do {                                                          \
        if ((K) <= 8) { constexpr int KP = 8, NP = 8; EXPR; }     \
        else if ((K) <= 12) { constexpr int KP = 12, NP = 16; EXPR; } \
        else if ((K) <= 16) { constexpr int KP = 16, NP = 16; EXPR; } \
        else if ((K) <= 20) { constexpr int KP = 20, NP = 32; EXPR; } \
        else if ((K) <= 24) { constexpr int KP = 24, NP = 32; EXPR; } \
        else { constexpr int KP = 32, NP = 32; EXPR; }            \
    } while (0)
template <typename F>
static int pick_ll_plan(mmsig_handle *h, F kernel, int KP, bool preg, bool dense, int V, long long D, int *grid_out, size_t *smem_out) {
    const int VP = V | 1, NW = (V + 31) / 32;
    const size_t base = (preg ? 0 : (size_t)KP * VP) + (size_t)TILE_S * VP + (size_t)TILE_S * KP + TILE_S + (size_t)NW * 32;
    const size_t smem = (dense ? tile_stage_offset(base) + tile_stage_doubles(V) : base + TILE_S + 2) * sizeof(double);
    if (smem > h->smem_optin) return fail(h, MMSIG_ELIMIT, "V too large for the log-likelihood tile (shared memory)");
    CU(allow_max_smem(h, kernel));
    int nb = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, NW * 32, smem));
    if (nb < 1) return fail(h, MMSIG_ELIMIT, "log-likelihood tile kernel does not fit on an SM for this K, V");
    *smem_out = smem;
    const long long ntiles = (D + TILE_S - 1) / TILE_S;
    *grid_out = (int)std::max<long long>(1, std::min<long long>((long long)h->numSM * nb, ntiles));
    return 0;
}

// k_theta_tile instance for (K, V): KP, EREG, NWT
#define TILE_DISPATCH_NW(V, EXPR)                                          \
    do {                                                                   \
        if ((V) <= 128) { constexpr int NWT = 4; EXPR; }                   \
        else if ((V) <= 256) { constexpr int NWT = 8; EXPR; }              \
        else if ((V) <= 512) { constexpr int NWT = 16; EXPR; }             \
        else { constexpr int NWT = 32; EXPR; }                             \
    } while (0)
#define TILE_DISPATCH(K, V, EXPR)                                          \
    do {                                                                   \
        if ((K) <= 8) { constexpr int KP = 8; constexpr bool EREG = true; TILE_DISPATCH_NW(V, EXPR); }         \
        else if ((K) <= 12) { constexpr int KP = 12; constexpr bool EREG = true; TILE_DISPATCH_NW(V, EXPR); }  \
        else if ((K) <= 16) { constexpr int KP = 16; constexpr bool EREG = true; TILE_DISPATCH_NW(V, EXPR); }  \
        else if ((K) <= 24) { constexpr int KP = 24; constexpr bool EREG = false; TILE_DISPATCH_NW(V, EXPR); } \
        else { constexpr int KP = 32; constexpr bool EREG = false; TILE_DISPATCH_NW(V, EXPR); }                \
    } while (0)
// ... and DENSE (the modality keeps dense count tiles)
#define TILE_DISPATCH_D(K, V, DENSE_RT, EXPR)                                                  \
    do {                                                                                       \
        if (DENSE_RT) { constexpr bool DENSE = true; TILE_DISPATCH(K, V, EXPR); }              \
        else { constexpr bool DENSE = false; TILE_DISPATCH(K, V, EXPR); }                      \
    } while (0)
#define MK_DISPATCH(MK, EXPR)                                     \
    do {                                                          \
        if ((MK) <= 8) { constexpr int MKP = 8; EXPR; }           \
        else if ((MK) <= 16) { constexpr int MKP = 16; EXPR; }    \
        else if ((MK) <= 24) { constexpr int MKP = 24; EXPR; }    \
        else { constexpr int MKP = 32; EXPR; }                    \
    } while (0)

// k_solve_lean instance for (lanes per sample G, sum(K)): CPL = ceil(sum(K) / G) coordinates per lane; FULL when no
// coordinate is padding (8 lanes only: the tuned shapes sum(K) = 8, 16, 24, 32)
#define LEAN_CASE(G_, C_, MK_, PH_, EXPR)                                                                        \
    do {                                                                                                         \
        if ((G_) == 8 && (MK_) == 8 * (C_)) { constexpr int LG = G_, LC = C_, LP = PH_; constexpr bool LF = true; EXPR; } \
        else { constexpr int LG = G_, LC = C_, LP = PH_; constexpr bool LF = false; EXPR; }                     \
    } while (0)
#define LEAN_DISPATCH(G_, MK_, PH_, EXPR)                                                        \
    do {                                                                                         \
        if ((G_) == 8) {                                                                         \
            if ((MK_) <= 8) LEAN_CASE(8, 1, MK_, PH_, EXPR);                                     \
            else if ((MK_) <= 16) LEAN_CASE(8, 2, MK_, PH_, EXPR);                               \
            else if ((MK_) <= 24) LEAN_CASE(8, 3, MK_, PH_, EXPR);                               \
            else LEAN_CASE(8, 4, MK_, PH_, EXPR);                                                \
        } else {                                                                                 \
            if ((MK_) <= 8) { constexpr int LG = 4, LC = 2, LP = PH_; constexpr bool LF = false; EXPR; }        \
            else if ((MK_) <= 12) { constexpr int LG = 4, LC = 3, LP = PH_; constexpr bool LF = false; EXPR; }  \
            else { constexpr int LG = 4, LC = 4, LP = PH_; constexpr bool LF = false; EXPR; }                   \
        }                                                                                        \
    } while (0)

// Shape-dependent part of set_data: allocations and launch plans for (D, D_total, M, K, V, nnz_m).
// *same_out: the resident shape matched and everything was kept.  Counts are not touched.
static int mmctm_prepare(mmsig_handle *h, int64_t D, int64_t D_total, int32_t M, const int32_t *K, const int32_t *V,
                         const long long *nnz, bool *same_out) {
    NEED(K && V && nnz, "null argument");
    NEED(D >= 1 && D_total >= D, "need 1 <= D <= D_total");
    NEED(h->nranks > 1 || D_total == D, "D_total != D without mmsig_comm_init");
    if (M < 1 || M > MAXM) return fail(h, MMSIG_ELIMIT, "1 <= M <= 8 modalities supported");
    int MKsum = 0;
    for (int m = 0; m < M; ++m) {
        NEED(K[m] >= 1 && V[m] >= 1, "K[m], V[m] must be >= 1");
        if (K[m] > 32) return fail(h, MMSIG_ELIMIT, "K[m] <= 32 supported");
        MKsum += K[m];
    }
    if (MKsum > MAXMK) return fail(h, MMSIG_ELIMIT, "sum(K) <= 64 supported");
    int rc;
    // same shape as what is already resident (a repeated fit! on the same corpus): keep every
    // allocation and launch plan
    bool same = h->mm.has_data && h->mm.p.M == M && h->mm.p.D == D && h->mm.p.D_total == D_total;
    for (int m = 0; same && m < M; ++m)
        same = h->mm.p.K[m] == K[m] && h->mm.p.V[m] == V[m] && h->mm.cb[m].nnz == nnz[m];
    *same_out = same;
    if (same) {
        h->mm.has_state = false;
        return 0;
    }
    free_pool(h->allocs_mm);
    h->mm = MmctmHost();
    MmctmHost &mm = h->mm;
    MmctmDev &p = mm.p;
    p.M = M;
    p.D = D;
    p.D_total = D_total;
    p.stop_rule = h->stop_rule;
    p.koff[0] = 0;
    p.goff[0] = 0;
    for (int m = 0; m < M; ++m) {
        p.K[m] = K[m];
        p.V[m] = V[m];
        p.koff[m + 1] = p.koff[m] + K[m];
        p.goff[m + 1] = p.goff[m] + K[m] * V[m];
    }
    p.MK = p.koff[M];
    mm.G = p.goff[M];
    double *dN = nullptr;
    if ((rc = dev_alloc(h, h->allocs_mm, &dN, (size_t)D * M))) return rc;
    p.N = dN;
    mm.nnz.resize(M);
    for (int m = 0; m < M; ++m) {
        if ((rc = ensure_countbuf(h, h->allocs_mm, mm.cb[m], D, nnz[m], V[m]))) return rc;
        p.rowptr[m] = mm.cb[m].rowptr;
        p.rec[m] = mm.cb[m].rec;
        p.cnt[m] = mm.cb[m].cnt;
        mm.nnz[m] = mm.cb[m].nnz;
    }
    const size_t DMK = (size_t)D * p.MK;
    if ((rc = dev_alloc(h, h->allocs_mm, &mm.lamA, DMK))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &mm.lamB, DMK))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &p.nu, DMK))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &p.sumtheta, DMK))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &mm.sumtheta_alt, DMK))) return rc;
    CU(cudaMemsetAsync(mm.sumtheta_alt, 0, DMK * sizeof(double), h->stream));
    if ((rc = dev_alloc(h, h->allocs_mm, &p.zeta, (size_t)D * M))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &p.nev_nu, (size_t)D))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &p.nev_lam, (size_t)D))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &p.work, (size_t)2))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &mm.d_ctl, (size_t)4))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &mm.d_llprev, (size_t)MAXM))) return rc;
    CU(cudaMemsetAsync(mm.d_ctl, 0, 4 * sizeof(int), h->stream));
    CU(cudaMemsetAsync(mm.d_llprev, 0, MAXM * sizeof(double), h->stream));
    p.ctl = mm.d_ctl;
    p.lam = mm.lamA;
    p.lam_prev = mm.lamB;
    for (double **t : {&p.gamma, &p.Elnphi, &p.Elnphi_prev, &p.phi, &p.stats})
        if ((rc = dev_alloc(h, h->allocs_mm, t, (size_t)mm.G))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &p.alpha, (size_t)M))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &p.mu, (size_t)p.MK))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &p.Sigma, (size_t)p.MK * p.MK))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &p.invSigma, (size_t)p.MK * p.MK))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &p.nusum, (size_t)p.MK))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &mm.d_ll, (size_t)M + 16))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &mm.d_status, (size_t)1))) return rc;
    CU(cudaMemsetAsync(mm.d_status, 0, sizeof(int), h->stream));
    CU(cudaMemsetAsync(p.nev_nu, 0, D * sizeof(int), h->stream));
    CU(cudaMemsetAsync(p.nev_lam, 0, D * sizeof(int), h->stream));
    CU(cudaMemsetAsync(p.sumtheta, 0, DMK * sizeof(double), h->stream));

    // launch plans
    for (int m = 0; m < M; ++m) {
        const int KV = K[m] * V[m];
        if (V[m] > 1024) return fail(h, MMSIG_ELIMIT, "V[m] <= 1024 supported");
        const bool dn = mm.cb[m].cnt != nullptr;
        if (h->precision) {
            TILE_DISPATCH(K[m], V[m], rc = pick_f32_plan(h, k_theta_tile_f32<KP, NWT>, f32_theta_smem(KP, V[m]), V[m], D, &mm.grid_theta[m],
                                                         &mm.smem_theta[m]));
            if (rc) return rc;
            TILE_DISPATCH(K[m], V[m], rc = pick_f32_plan(h, k_loglik_tile_f32<KP, NWT>, f32_ll_smem(KP, V[m]), V[m], D, &mm.grid_ll[m],
                                                         &mm.smem_ll[m]));
            if (rc) return rc;
        } else {
            TILE_DISPATCH_D(K[m], V[m], dn, rc = pick_tile_plan(h, k_theta_tile<KP, EREG, NWT, DENSE>, KP, EREG, DENSE, V[m], D, &mm.grid_theta[m],
                                                                &mm.smem_theta[m]));
            if (rc) return rc;
            TILE_DISPATCH_D(K[m], V[m], dn, rc = pick_ll_plan(h, k_loglik_tile<KP, EREG, NWT, DENSE>, KP, EREG, DENSE, V[m], D, &mm.grid_ll[m], &mm.smem_ll[m]));
            if (rc) return rc;
        }
        mm.W_theta[m] = TILE_S;
        if ((rc = dev_alloc(h, h->allocs_mm, &mm.part_theta[m], (size_t)mm.grid_theta[m] * KV))) return rc;
        CU(cudaMemsetAsync(mm.part_theta[m], 0, (size_t)mm.grid_theta[m] * KV * sizeof(double2), h->stream));
    }
    mm.wide = p.MK > 32;
    {
        // Lane layout of the LD_MMA kernels for sum(K) <= 32 (all bit-identical; tests/test_gpu_mmctm.py::test_solver_layouts_are_bit_identical).
        // Default: k_solve_lean (mmctm_lean.cuh), 4 lanes per sample up to sum(K) = 12, 8 above.  Measured on B200 at
        // D = 1e6 (profiles/r02_solve_summary.md): sum(K) = 24: 18.8 ms against 33.7 ms for one coordinate per lane
        // (k_solve); sum(K) = 10: 7.2 ms (4 lanes) / 9.8 ms (8 lanes) against 14.1 ms (k_solve_pack).
        // MMSIG_SOLVE=warp | lean8 | lean4 overrides (A/B measurements).
        const char *e = getenv("MMSIG_SOLVE");
        mm.solve_lean = p.MK > 32 ? 0 : (p.MK <= 12 ? 4 : 8);
        if (e && p.MK <= 32) {
            if (!strcmp(e, "warp")) mm.solve_lean = 0;
            else if (!strcmp(e, "lean8")) mm.solve_lean = 8;
            else if (!strcmp(e, "lean4")) mm.solve_lean = p.MK <= 16 ? 4 : 8;     // 4 lanes x (> 4 coordinates) spills: 43 ms against 23 ms at sum(K) = 24
        }
    }
    CU(allow_max_smem(h, k_mstep2));
    auto grid_for = [&](int nb) {
        return (int)std::max<long long>(1, std::min<long long>((long long)h->numSM * std::max(nb, 1), (D + 7) / 8));
    };
    if (!mm.wide) {
        int nb = 0;
        if (mm.solve_lean) {
            LEAN_DISPATCH(mm.solve_lean, p.MK, PH_NU, CU(allow_max_smem(h, k_solve_lean<LG, LC, LP, LF>)));
            LEAN_DISPATCH(mm.solve_lean, p.MK, PH_LAM, CU(allow_max_smem(h, k_solve_lean<LG, LC, LP, LF>)));
            LEAN_DISPATCH(mm.solve_lean, p.MK, PH_LAM, CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_solve_lean<LG, LC, LP, LF>, 128,
                                                                                                         lean_smem_doubles<LG, LC, LP>() * sizeof(double))));
        }
        else if (p.MK <= 8) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_solve_pack<8>, 256, 0));
        else if (p.MK <= 16) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_solve_pack<16>, 256, 0));
        else MK_DISPATCH(p.MK, CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_solve<MKP>, 256, 0)));
        {
            // samples per block: 4 warps x 4 or 8 (lean), 8 warps x 1, 2 or 4 (one coordinate per lane)
            const int spb = mm.solve_lean ? 4 * (32 / mm.solve_lean) : (p.MK <= 8 ? 32 : (p.MK <= 16 ? 16 : 8));
            mm.grid_solve = (int)std::max<long long>(1, std::min<long long>((long long)h->numSM * std::max(nb, 1),
                                                                            (D + spb - 1) / spb));
        }
        mm.grid_zeta = grid_for(nb);
        mm.smem_post = (size_t)512 * sizeof(double);
        MK_DISPATCH(p.MK, CU(allow_max_smem(h, k_moments<MKP>)));
        MK_DISPATCH(p.MK, CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_moments<MKP>, 256, mm.smem_post)));
        mm.grid_mom = grid_for(nb);
    } else {
        int nb = 0;
        mm.smem_solve = (size_t)(WMK * WSTRIDE + 8 * WMK) * sizeof(double) + (size_t)8 * 4 * 32 * sizeof(double2);
        CU(allow_max_smem(h, k_solve_wide));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_solve_wide, 256, mm.smem_solve));
        mm.grid_solve = grid_for(nb);
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_moments_wide, 256, 0));
        mm.grid_mom = grid_for(nb);
    }
    // partial buffer of the log-likelihood tiles (one dd per block and modality) and grid of the ELBO kernels
    mm.grid_post = grid_for(4);
    for (int m = 0; m < M; ++m) mm.grid_post = std::max(mm.grid_post, mm.grid_ll[m]);
    const int P1 = mm.G + 2 * p.MK, P2 = p.MK * p.MK + M;
    if ((rc = dev_alloc(h, h->allocs_mm, &mm.part_solve, (size_t)mm.grid_solve * 2 * p.MK))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &mm.part_post, (size_t)mm.grid_post * P2))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &mm.part_mom, (size_t)mm.grid_mom * P2))) return rc;
    CU(cudaMemsetAsync(mm.part_post, 0, (size_t)mm.grid_post * P2 * sizeof(double2), h->stream));
    CU(cudaMemsetAsync(mm.part_mom, 0, (size_t)mm.grid_mom * P2 * sizeof(double2), h->stream));
    if ((rc = dev_alloc(h, h->allocs_mm, &mm.part_elbo, (size_t)M * mm.grid_post * 4 + 8))) return rc;   // ELBO partials [M][grid_post][4] + 4 table terms
    if ((rc = dev_alloc(h, h->allocs_mm, &mm.rank_p1, (size_t)P1))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &mm.gath_p1, (size_t)P1 * h->nranks))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &mm.rank_p2, (size_t)P2 + 16))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &mm.gath_p2, (size_t)(P2 + 16) * h->nranks))) return rc;
    return 0;
}

extern "C" int32_t mmsig_mmctm_set_data(mmsig_handle *h, int64_t D, int64_t D_total, int32_t M, const int32_t *K,
                                        const int32_t *V, const int64_t *const *rowptr, const int32_t *const *term,
                                        const int32_t *const *count) {
    NEED(h, "null handle");
    NEED(K && V && rowptr && term && count, "null argument");
    CU(cudaSetDevice(h->device));
    if (M < 1 || M > MAXM) return fail(h, MMSIG_ELIMIT, "1 <= M <= 8 modalities supported");
    NEED(D >= 1, "need 1 <= D <= D_total");
    bool same = false;
    int rc;
    long long nnz[MAXM];
    for (int m = 0; m < M; ++m) {
        if ((rc = check_rowptr(h, rowptr[m], D))) return rc;
        nnz[m] = rowptr[m][D];
    }
    if ((rc = mmctm_prepare(h, D, D_total, M, K, V, nnz, &same))) return rc;
    MmctmHost &mm = h->mm;
    mm.has_data = false;
    long long ntot[MAXM];
    for (int m = 0; m < M; ++m)
        if ((rc = upload_counts(h, h->allocs_mm, mm.cb[m], D, V[m], M, m, rowptr[m], term[m], count[m],
                                const_cast<double *>(mm.p.N), &ntot[m], 1)))
            return rc;
    if ((rc = allsum_ll(h, ntot, M))) return rc;
    for (int m = 0; m < M; ++m) mm.p.Ntot[m] = (double)ntot[m];
    mm.has_data = true;
    return 0;
}

__global__ void k_fill(double *x, size_t n, double v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = v;
}
static void fill(mmsig_handle *h, double *x, size_t n, double v) {
    LaunchScope ls(h, "k_fill");
    int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)h->numSM * 8);
    k_fill<<<std::max(grid, 1), 256, 0, h->stream>>>(x, n, v);
}

extern "C" int32_t mmsig_mmctm_set_state(mmsig_handle *h, const double *alpha, const double *gamma, const double *lambda,
                                         const double *nu, const double *mu, const double *Sigma, const double *invSigma) {
    NEED(h, "null handle");
    MmctmHost &mm = h->mm;
    NEED(mm.has_data, "mmsig_mmctm_set_data first");
    NEED(alpha && gamma, "alpha and gamma are required");
    CU(cudaSetDevice(h->device));
    MmctmDev &p = mm.p;
    const size_t DMK = (size_t)p.D * p.MK, MK2 = (size_t)p.MK * p.MK;
    for (int m = 0; m < p.M; ++m) NEED(alpha[m] > 0, "alpha must be > 0");
    mm.alpha_host.assign(alpha, alpha + p.M);
    CU(cudaMemcpyAsync(p.alpha, alpha, p.M * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(p.gamma, gamma, mm.G * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (lambda) CU(cudaMemcpyAsync(p.lam, lambda, DMK * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    else CU(cudaMemsetAsync(p.lam, 0, DMK * sizeof(double), h->stream));
    CU(cudaMemcpyAsync(p.lam_prev, p.lam, DMK * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (nu) CU(cudaMemcpyAsync(p.nu, nu, DMK * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    else fill(h, p.nu, DMK, 1.0);
    if (mu) CU(cudaMemcpyAsync(p.mu, mu, p.MK * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    else CU(cudaMemsetAsync(p.mu, 0, p.MK * sizeof(double), h->stream));
    std::vector<double> eye(MK2, 0.0);
    for (int j = 0; j < p.MK; ++j) eye[(size_t)j * p.MK + j] = 1.0;
    CU(cudaMemcpyAsync(p.Sigma, Sigma ? Sigma : eye.data(), MK2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(p.invSigma, invSigma ? invSigma : eye.data(), MK2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    {
        LaunchScope ls(h, "k_elnphi");
        k_elnphi<<<1, 1024, 0, h->stream>>>(p);
    }
    {
        LaunchScope ls(h, "k_zeta");
        if (mm.wide) k_zeta_props_wide<<<mm.grid_solve, 256, 0, h->stream>>>(p, nullptr, 1);
        else k_zeta<<<mm.grid_solve, 256, 0, h->stream>>>(p);
    }
    CU(cudaMemsetAsync(p.stats, 0, mm.G * sizeof(double), h->stream));
    CU(cudaStreamSynchronize(h->stream));      // eye / caller buffers may go away
    CU(cudaGetLastError());
    mm.has_state = true;
    mm.estep_done = false;
    return 0;
}

extern "C" int32_t mmsig_mmctm_set_phi(mmsig_handle *h, const double *phi) {
    NEED(h && phi, "null argument");
    NEED(h->mm.has_state, "mmsig_mmctm_set_state first");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(h->mm.p.phi, phi, h->mm.G * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

// ---- update_α! (src/MMCTM.jl:252-269): a 1-D LD_MMA per modality on table sums, on the host.
// Plain libm arithmetic in the reference's operation order (α_objective, src/common.jl:38-46;
// digamma per SpecialFunctions.jl; NLopt LD_MMA with lb = 1e-7, xtol_rel = xtol_abs = 1e-5).
static double host_digamma(double x) {
    double psi = 0.0;
    if (x <= 0.0) { psi = -M_PI / std::tan(M_PI * x); x = 1.0 - x; }
    if (x < 7.0) {
        int n = 7 - (int)std::floor(x);
        for (int v = 1; v <= n - 1; ++v) psi -= 1.0 / (x + (double)v);
        psi -= 1.0 / x;
        x += (double)n;
    }
    double t = 1.0 / x;
    psi += std::log(x) - 0.5 * t;
    t *= t;
    static const double c[8] = {0.08333333333333333, -0.008333333333333333, 0.003968253968253968, -0.004166666666666667,
                                0.007575757575757576, -0.021092796092796094, 0.08333333333333333, -0.4432598039215686};
    double p = c[7];
    for (int i = 6; i >= 0; --i) p = std::fma(p, t, c[i]);
    psi -= t * p;
    return psi;
}
static double host_lgamma(double x) { int sg; return lgamma_r(x, &sg); }
// returns -objective and -gradient (NLopt minimises the negated max_objective)
static double neg_alpha_objective(double a, double *g, double sumE, int K, int V) {
    *g = -(K * V * (host_digamma(V * a) - host_digamma(a)) + sumE);
    return -(K * (host_lgamma(V * a) - V * host_lgamma(a)) + a * sumE);
}
static double mma_alpha(double x, double sumE, int K, int V, int stop_rule) {
    const double lb = 1e-7, xtol = 1e-5;
    double sigma = 1.0, rho = 1.0, g, gcur, fcur;
    double fmin = neg_alpha_objective(x, &g, sumE, K, V);
    double xcur = x, xprev = x, xprevprev = x;
    int k = 0, nev = 1;
    while (true) {
        if (++k > 1) xprevprev = xprev;
        xprev = xcur;
        while (true) {
            double u = g;
            const double v = std::fabs(g) * sigma + 0.5 * rho, sigma2 = sigma * sigma;
            u *= sigma2;
            const double r = u / (v * sigma);
            double dx = (u / v) / (-1 - std::sqrt(std::fabs(1 - r * r)));
            double xc = x + dx;
            if (xc > x + 0.9 * sigma) xc = x + 0.9 * sigma;
            else if (xc < x - 0.9 * sigma) xc = x - 0.9 * sigma;
            if (xc < lb) xc = lb;
            dx = xc - x;
            const double dx2 = dx * dx, denominv = 1.0 / (sigma2 - dx2), cc = sigma2 * dx;
            double gval = fmin, wval = 0.0;
            gval += (g * cc + (std::fabs(g) * sigma + 0.5 * rho) * dx2) * denominv;
            wval += 0.5 * dx2 * denominv;
            xcur = xc;
            fcur = neg_alpha_objective(xcur, &gcur, sumE, K, V);
            ++nev;
            const bool inner_done = gval >= fcur;
            if (fcur < fmin) { fmin = fcur; x = xcur; g = gcur; }
            if (nev >= 10000) return x;
            if (inner_done) break;
            if (fcur > gval) rho = std::min(10 * rho, 1.1 * (rho + (fcur - gval) / wval));
        }
        const double ad = std::fabs(xcur - xprev);
        bool stop;
        if (stop_rule == 1) stop = ad < xtol || ad < xtol * (std::fabs(xcur) + std::fabs(xprev)) * 0.5 || xcur == xprev;
        else stop = (ad <= xtol * std::fabs(xcur)) || !(ad > xtol);
        if (stop) break;
        rho = 0.1 * rho > 1e-5 ? 0.1 * rho : 1e-5;
        if (k > 1) {
            const double s2 = (xcur - xprev) * (xprev - xprevprev);
            sigma *= s2 < 0 ? 0.7 : (s2 > 0 ? 1.2 : 1.0);
        }
    }
    return x;
}
static int mmctm_update_alpha(mmsig_handle *h) {
    MmctmHost &mm = h->mm;
    MmctmDev &p = mm.p;
    if (p.factored) {            // src/IMMCTM.jl:223-241: one alpha per (modality, feature), on the feature tables
        std::vector<double> E(p.T);
        CU(cudaMemcpyAsync(E.data(), p.Elnphif, p.T * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        std::vector<int> roff(p.R);
        CU(cudaMemcpyAsync(roff.data(), p.row_off, p.R * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        int r0 = 0;
        for (int m = 0; m < p.M; ++m) {
            const int nf = p.nfeat[m], K = p.K[m];
            for (int f = 0; f < nf; ++f) {
                double s = 0.0;
                const int J = mm.row_len_host[r0 + f];
                for (int k = 0; k < K; ++k)
                    for (int j = 0; j < J; ++j) s += E[roff[r0 + k * nf + f] + j];
                double &a = mm.alphaf_host[p.aoff[m] + f];
                a = mma_alpha(a, s, K, J, h->stop_rule);
            }
            r0 += K * nf;
        }
        CU(cudaMemcpyAsync(p.alphaf, mm.alphaf_host.data(), mm.alphaf_host.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        return 0;
    }
    std::vector<double> E(mm.G);
    CU(cudaMemcpyAsync(E.data(), p.Elnphi, mm.G * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (int m = 0; m < p.M; ++m) {
        double s = 0.0;
        for (int t = p.goff[m]; t < p.goff[m + 1]; ++t) s += E[t];
        mm.alpha_host[m] = mma_alpha(mm.alpha_host[m], s, p.K[m], p.V[m], h->stop_rule);
    }
    CU(cudaMemcpyAsync(p.alpha, mm.alpha_host.data(), p.M * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int32_t mmsig_mmctm_get_alpha(mmsig_handle *h, double *alpha_out) {
    NEED(h && alpha_out, "null argument");
    NEED(h->mm.has_state, "mmsig_mmctm_set_state first");
    memcpy(alpha_out, h->mm.alpha_host.data(), h->mm.p.M * sizeof(double));
    return 0;
}

// view of samples [d0, d1) of the resident shard (rowptr entries are absolute offsets into rec)
static MmctmDev chunk_view(const MmctmDev &p, long long d0, long long d1, int accum) {
    MmctmDev q = p;
    q.D = d1 - d0;
    for (int m = 0; m < p.M; ++m) {
        q.rowptr[m] = p.rowptr[m] + d0;
        if (p.cnt[m]) q.cnt[m] = p.cnt[m] + (size_t)d0 * p.V[m];        // d0 is a multiple of 32: the chunk's tiles are the shard's
    }
    q.N = p.N + d0 * p.M;
    q.lam = p.lam + d0 * p.MK;
    q.lam_prev = p.lam_prev + d0 * p.MK;
    q.nu = p.nu + d0 * p.MK;
    q.sumtheta = p.sumtheta + d0 * p.MK;
    q.zeta = p.zeta + d0 * p.M;
    q.nev_nu = p.nev_nu + d0;
    q.nev_lam = p.nev_lam + d0;
    q.accum = accum;
    return q;
}

// E-step kernels (θ pass per modality, then ζ / ν / λ) over the samples of view q
static void mmctm_estep_launch(mmsig_handle *h, const MmctmDev &q, uint32_t flags) {
    MmctmHost &mm = h->mm;
    const int freeze_topics = (flags & MMSIG_FLAG_FREEZE_TOPICS) ? 1 : 0, unsm = (flags & MMSIG_FLAG_UNSMOOTHED) ? 1 : 0;
    auto cap = [&](int grid, int per_block) { return (int)std::max<long long>(1, std::min<long long>(grid, (q.D + per_block - 1) / per_block)); };
    for (int m = 0; m < q.M; ++m) {
        LaunchScope ls(h, "k_theta_tile");
        const int nthr = 32 * ((q.V[m] + 31) / 32);
        if (h->precision)
            TILE_DISPATCH(q.K[m], q.V[m], (k_theta_tile_f32<KP, NWT><<<cap(mm.grid_theta[m], TILE_S), nthr, mm.smem_theta[m], h->stream>>>(
                                              q, m, mm.part_theta[m], unsm, !freeze_topics)));
        else
        TILE_DISPATCH_D(q.K[m], q.V[m], q.cnt[m] != nullptr, (k_theta_tile<KP, EREG, NWT, DENSE><<<cap(mm.grid_theta[m], TILE_S), nthr, mm.smem_theta[m], h->stream>>>(
                                          q, m, mm.part_theta[m], unsm, !freeze_topics)));
    }
    if (h->aux_pending) {                  // Σ, invΣ of the previous iteration come from the side stream (mmctm_mstep_launch)
        cudaStreamWaitEvent(h->stream, h->ev_join, 0);
        h->aux_pending = false;
    }
    cudaMemsetAsync(q.work, 0, sizeof(unsigned long long), h->stream);      // the solve kernels draw samples from this counter
    {
        LaunchScope ls(h, "k_solve");
        if (mm.wide) k_solve_wide<<<cap(mm.grid_solve, 8), 256, mm.smem_solve, h->stream>>>(q, mm.part_solve);
        else if (mm.solve_lean) {
            // ν for every sample, then λ (which reads the new ν and the ζ the first kernel stored)
            const int gs = cap(mm.grid_solve, 4 * (32 / mm.solve_lean));
            // (Handing the samples out longest-first -- a counting sort by the previous iteration's evaluation counts -- was
            // measured in round 2 and bought nothing: 2.703 against 2.707 ms at D = 125000, profiles/notes/r02k_order_ab.txt.)
            LEAN_DISPATCH(mm.solve_lean, q.MK, PH_NU, (k_solve_lean<LG, LC, LP, LF><<<gs, 128, lean_smem_doubles<LG, LC, LP>() * sizeof(double), h->stream>>>(q, mm.part_solve)));
            cudaMemsetAsync(q.work, 0, sizeof(unsigned long long), h->stream);
            h->launches++;
            LEAN_DISPATCH(mm.solve_lean, q.MK, PH_LAM, (k_solve_lean<LG, LC, LP, LF><<<gs, 128, lean_smem_doubles<LG, LC, LP>() * sizeof(double), h->stream>>>(q, mm.part_solve)));
        }
        else if (q.MK <= 8) k_solve_pack<8><<<cap(mm.grid_solve, 32), 256, 0, h->stream>>>(q, mm.part_solve);
        else if (q.MK <= 16) k_solve_pack<16><<<cap(mm.grid_solve, 16), 256, 0, h->stream>>>(q, mm.part_solve);
        else MK_DISPATCH(q.MK, (k_solve<MKP><<<cap(mm.grid_solve, 8), 256, 0, h->stream>>>(q, mm.part_solve)));
    }
}

static int mmctm_mstep_launch(mmsig_handle *h, uint32_t flags, bool overlap = false);

static int mmctm_iterate_async(mmsig_handle *h, uint32_t flags, bool overlap = false) {
    MmctmHost &mm = h->mm;
    MmctmDev &p = mm.p;
    mm.last_unsmoothed = (flags & MMSIG_FLAG_UNSMOOTHED) != 0;
    std::swap(p.lam, p.lam_prev);              // lam_prev = λ of the previous iteration (what θ uses)
    std::swap(p.sumtheta, mm.sumtheta_alt);    // this iteration's θ pass writes the other buffer (see sumtheta_alt)
    mmctm_estep_launch(h, p, flags);
    return mmctm_mstep_launch(h, flags, overlap);
}

// everything after the per-sample E-step: combine, exchange, μ, γ, Elnϕ, ϕ, [α], [Σ, invΣ], props / LL
static int mmctm_mstep_launch(mmsig_handle *h, uint32_t flags, bool overlap) {
    MmctmHost &mm = h->mm;
    MmctmDev &p = mm.p;
    const int do_sigma = (flags & MMSIG_FLAG_UPDATE_SIGMA) ? 1 : 0;
    const int freeze_topics = (flags & MMSIG_FLAG_FREEZE_TOPICS) ? 1 : 0, freeze_mu = (flags & MMSIG_FLAG_FREEZE_MU) ? 1 : 0;
    const int P1 = mm.G + 2 * p.MK, P2 = p.MK * p.MK + p.M;
    {
        CombineSegs s{};
        s.nseg = p.M + 1;
        for (int m = 0; m < p.M; ++m) {
            s.src[m] = mm.part_theta[m];
            s.nparts[m] = mm.grid_theta[m];
            s.n[m] = p.K[m] * p.V[m];
            s.dst_off[m] = p.goff[m];
        }
        s.src[p.M] = mm.part_solve;
        s.nparts[p.M] = mm.grid_solve;
        s.n[p.M] = 2 * p.MK;
        s.dst_off[p.M] = mm.G;
        LaunchScope ls(h, "k_combine");
        k_combine<<<(P1 + 7) / 8, 256, 0, h->stream>>>(s, mm.rank_p1, p.ctl);
    }
    const double2 *g1 = nullptr;
    int rc;
    if ((rc = gather(h, mm.rank_p1, mm.gath_p1, P1, &g1))) return rc;
    {
        LaunchScope ls(h, "k_mstep1");
        if (p.factored) k_imstep1<<<1, 1024, (size_t)2 * p.R * sizeof(double), h->stream>>>(p, g1, h->nranks, freeze_topics, freeze_mu);
        else k_mstep1<<<1, 1024, 0, h->stream>>>(p, g1, h->nranks, freeze_topics, freeze_mu, (flags & MMSIG_FLAG_UNSMOOTHED) ? 1 : 0);
    }
    if ((flags & MMSIG_FLAG_AUTO_ALPHA) && !freeze_topics)       // src/MMCTM.jl:472-474, after update_γ!
        if ((rc = mmctm_update_alpha(h))) return rc;
    // From here on nothing feeds the next θ pass (it reads λ and the Elnϕ that k_mstep1 just wrote): inside fit's
    // sync-free loop this half runs on a side stream, so that its serial chain combine -> exchange -> LU (0.15 ms that
    // every rank repeats) hides behind the next iteration's θ pass; the next solve waits for ev_join.  Not with
    // per-kernel event timing (events are recorded on the main stream) and not inside a single-process group (its
    // exchange protocol is ordered on the main stream).
    cudaStream_t st = h->stream;
    if (overlap && !h->profile && !h->grp && !(flags & MMSIG_FLAG_AUTO_ALPHA)) {
        if (!h->s_aux) {
            CU(cudaStreamCreateWithFlags(&h->s_aux, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
        }
        CU(cudaEventRecord(h->ev_fork, h->stream));
        CU(cudaStreamWaitEvent(h->s_aux, h->ev_fork, 0));
        st = h->s_aux;
    }
    if (do_sigma) {
        if (mm.wide) {
            for (int r0 = 0; r0 < p.MK; r0 += 16) {
                LaunchScope ls(h, "k_moments");
                k_moments_wide<<<mm.grid_mom, 256, 0, st>>>(p, mm.part_mom, r0);
            }
        } else {
            LaunchScope ls(h, "k_moments");
            MK_DISPATCH(p.MK, (k_moments<MKP><<<mm.grid_mom, 256, mm.smem_post, st>>>(p, mm.part_mom)));
        }
    }
    for (int m = 0; m < p.M; ++m) {
        LaunchScope ls(h, "k_loglik_tile");
        const int nthr = 32 * ((p.V[m] + 31) / 32);
        if (h->precision)
            TILE_DISPATCH(p.K[m], p.V[m], (k_loglik_tile_f32<KP, NWT><<<mm.grid_ll[m], nthr, mm.smem_ll[m], st>>>(
                                              p, m, mm.part_post + p.MK * p.MK + m, P2)));
        else
        TILE_DISPATCH_D(p.K[m], p.V[m], p.cnt[m] != nullptr, (k_loglik_tile<KP, EREG, NWT, DENSE><<<mm.grid_ll[m], nthr, mm.smem_ll[m], st>>>(
                                          p, m, mm.part_post + p.MK * p.MK + m, P2)));
    }
    {
        // moments (first MK*MK entries) from the moments pass, LL (last M) from the LL pass
        CombineSegs s{};
        s.nseg = 1 + p.M;
        s.src[0] = mm.part_mom;
        s.nparts[0] = do_sigma ? mm.grid_mom : 0;
        s.n[0] = p.MK * p.MK;
        s.dst_off[0] = 0;
        s.stride[0] = P2;
        for (int m = 0; m < p.M; ++m) {
            s.src[1 + m] = mm.part_post + p.MK * p.MK + m;
            s.nparts[1 + m] = mm.grid_ll[m];
            s.n[1 + m] = 1;
            s.dst_off[1 + m] = p.MK * p.MK + m;
            s.stride[1 + m] = P2;
        }
        LaunchScope ls(h, "k_combine");
        k_combine<<<(P2 + 7) / 8, 256, 0, st>>>(s, mm.rank_p2, p.ctl);
    }
    const double2 *g2 = nullptr;
    if ((rc = gather(h, mm.rank_p2, mm.gath_p2, P2, &g2, st))) return rc;
    {
        LaunchScope ls(h, "k_mstep2");
        const size_t lus = (size_t)2 * p.MK * p.MK * sizeof(double) + p.MK * sizeof(int);
        k_mstep2<<<1, 256, lus, st>>>(p, g2, h->nranks, do_sigma, mm.d_ll, mm.d_status, mm.cur_iter, mm.cur_tol, mm.d_llprev);
    }
    if (st != h->stream) {
        CU(cudaEventRecord(h->ev_join, st));
        h->aux_pending = true;
    }
    mm.estep_done = true;
    return 0;
}

extern "C" int32_t mmsig_mmctm_iterate(mmsig_handle *h, uint32_t flags, double *ll_out) {
    NEED(h, "null handle");
    NEED(h->mm.has_state, "mmsig_mmctm_set_state first");
    NEED(!(h->mm.p.factored && (flags & MMSIG_FLAG_UNSMOOTHED)), "the IMMCTM has no unsmoothed E-step (no transform in src/IMMCTM.jl)");
    CU(cudaSetDevice(h->device));
    int rc = mmctm_iterate_async(h, flags);
    if (rc) return rc;
    double ll[MAXM];
    int status = 0;
    CU(cudaMemcpyAsync(ll, h->mm.d_ll, h->mm.p.M * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(&status, h->mm.d_status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    if (status) return fail(h, MMSIG_EINVAL, "Sigma is singular (inv failed)");
    if (ll_out) memcpy(ll_out, ll, h->mm.p.M * sizeof(double));
    return 0;
}

// src/common.jl:48-51 on the last two LL vectors
#ifndef MMSIG_DEVICE_RULE_DEFAULT
#define MMSIG_DEVICE_RULE_DEFAULT true      // the batched loop below; MMSIG_DEVICE_RULE=0 keeps one host round trip per iteration (A/B)
#endif
static bool converged_vec(const double *prev, const double *cur, int M, double tol) {
    double r = 0.0;
    for (int i = 0; i < M; ++i) {
        double v = std::fabs(prev[i] - cur[i]) / std::fabs(cur[i]);
        if (v > r || v != v) r = v;
    }
    return r < tol;
}

// Iterations first..maxiter of fit!'s loop (src/MMCTM.jl:462-489).  The convergence rule can only fire once the
// history holds more than 10 entries (:485), so the iterations before that are enqueued back to back, their
// log-likelihoods landing in page-locked memory, and the host synchronises once; from then on once per iteration
// (the decision needs that iteration's LL).  hist_len: entries of ll_hist filled before `first`.
static int mmctm_run_iterations(mmsig_handle *h, int first, int32_t maxiter, double tol, uint32_t flags, double *ll_hist,
                                int *it_out, int *conv_out) {
    MmctmHost &mm = h->mm;
    const int M = mm.p.M;
    int it = first - 1, conv = 0;
    int iter = first;
    if (!(flags & MMSIG_FLAG_AUTO_ALPHA)) {            // update_alpha! synchronises inside every iteration anyway
        const int nfree = std::min<int>(maxiter, 10) - first + 1;
        if (nfree > 1) {
            int *status = reinterpret_cast<int *>(h->ll_pinned + (size_t)12 * MAXM);
            for (int i = 0; i < nfree; ++i) {
                int rc = mmctm_iterate_async(h, flags, true);
                if (rc) return rc;
                cudaStream_t cs = h->aux_pending ? h->s_aux : h->stream;       // the stream that ran k_mstep2
                CU(cudaMemcpyAsync(h->ll_pinned + (size_t)i * MAXM, mm.d_ll, M * sizeof(double), cudaMemcpyDeviceToHost, cs));
                CU(cudaMemcpyAsync(status + i, mm.d_status, sizeof(int), cudaMemcpyDeviceToHost, cs));
            }
            CU(cudaStreamSynchronize(h->stream));
            if (h->s_aux) CU(cudaStreamSynchronize(h->s_aux));
            h->aux_pending = false;
            CU(cudaGetLastError());
            for (int i = 0; i < nfree; ++i) {
                if (status[i]) return fail(h, MMSIG_EINVAL, "Sigma is singular (inv failed)");
                memcpy(ll_hist + (size_t)(first - 1 + i) * M, h->ll_pinned + (size_t)i * MAXM, M * sizeof(double));
            }
            iter = first + nfree;
            it = iter - 1;
        }
    }
    // From iteration 11 on the rule of src/MMCTM.jl:485 can end the loop.  It is evaluated on the device (k_mstep2), so
    // the iterations are still enqueued without host round trips, in batches of kBatch: when the rule fires in iteration
    // j every kernel of the later iterations of the batch returns at once (MmctmDev::ctl), and the state is exactly what
    // the reference leaves -- that of iteration j.  The host reads the log-likelihoods, the flag and the count after
    // each batch.  (autoα updates α on the host inside every iteration: one iteration per round trip there.)
    constexpr int kBatch = 8;
    const char *edr = getenv("MMSIG_DEVICE_RULE");
    const bool device_rule = edr ? atoi(edr) != 0 : MMSIG_DEVICE_RULE_DEFAULT;
    if (device_rule && !(flags & MMSIG_FLAG_AUTO_ALPHA) && iter > 10) {
        int *status = reinterpret_cast<int *>(h->ll_pinned + (size_t)12 * MAXM);
        int *ctl_host = status + 16;
        while (iter <= maxiter && !conv) {
            const int nb = std::min<int>(kBatch, maxiter - iter + 1);
            for (int i = 0; i < nb; ++i) {
                mm.cur_iter = iter + i;
                mm.cur_tol = tol;
                int rc = mmctm_iterate_async(h, flags, true);
                mm.cur_iter = 0;
                if (rc) return rc;
                cudaStream_t cs = h->aux_pending ? h->s_aux : h->stream;
                CU(cudaMemcpyAsync(h->ll_pinned + (size_t)i * MAXM, mm.d_ll, M * sizeof(double), cudaMemcpyDeviceToHost, cs));
                CU(cudaMemcpyAsync(status + i, mm.d_status, sizeof(int), cudaMemcpyDeviceToHost, cs));
                if (i == nb - 1) CU(cudaMemcpyAsync(ctl_host, mm.d_ctl, 2 * sizeof(int), cudaMemcpyDeviceToHost, cs));
            }
            CU(cudaStreamSynchronize(h->stream));
            if (h->s_aux) CU(cudaStreamSynchronize(h->s_aux));
            h->aux_pending = false;
            CU(cudaGetLastError());
            const int done = ctl_host[0];
            const int n_exec = done ? ctl_host[1] - (iter - 1) : nb;         // iterations of this batch that ran
            if (n_exec < 1 || n_exec > nb) return fail(h, MMSIG_ECUDA, "internal: iteration count of a batch out of range");
            for (int i = 0; i < n_exec; ++i) {
                if (status[i]) return fail(h, MMSIG_EINVAL, "Sigma is singular (inv failed)");
                memcpy(ll_hist + (size_t)(iter - 1 + i) * M, h->ll_pinned + (size_t)i * MAXM, M * sizeof(double));
            }
            if ((nb - n_exec) & 1) {                                         // the skipped iterations' host-side buffer swaps
                std::swap(mm.p.lam, mm.p.lam_prev);
                std::swap(mm.p.sumtheta, mm.sumtheta_alt);
            }
            it = iter + n_exec - 1;
            iter += nb;
            if (done) {
                conv = 1;
                CU(cudaMemsetAsync(mm.d_ctl, 0, 2 * sizeof(int), h->stream));
                CU(cudaStreamSynchronize(h->stream));
            }
        }
    }
    for (; iter <= maxiter && !conv; ++iter) {
        double *ll = ll_hist + (size_t)(iter - 1) * M;
        int rc = mmsig_mmctm_iterate(h, flags, ll);
        if (rc) return rc;
        it = iter;
        if (iter > 10 && converged_vec(ll - M, ll, M, tol)) { conv = 1; break; }   // src/MMCTM.jl:485
    }
    *it_out = it;
    *conv_out = conv;
    return 0;
}

extern "C" int32_t mmsig_mmctm_fit(mmsig_handle *h, int32_t maxiter, double tol, uint32_t flags, double *ll_hist,
                                   int32_t *n_iter, int32_t *converged) {
    NEED(h, "null handle");
    NEED(h->mm.has_state, "mmsig_mmctm_set_state first");
    NEED(maxiter >= 1 && ll_hist, "maxiter >= 1 and ll_hist required");
    NEED(!(h->mm.p.factored && (flags & MMSIG_FLAG_UNSMOOTHED)), "the IMMCTM has no unsmoothed E-step (no transform in src/IMMCTM.jl)");
    CU(cudaSetDevice(h->device));
    int it = 0, conv = 0;
    int rc = mmctm_run_iterations(h, 1, maxiter, tol, flags, ll_hist, &it, &conv);
    if (rc) return rc;
    if (n_iter) *n_iter = it;
    if (converged) *converged = conv;
    return 0;
}

extern "C" int32_t mmsig_mmctm_get_state(mmsig_handle *h, double *lambda, double *nu, double *zeta, double *mu,
                                         double *Sigma, double *invSigma, double *gamma, double *Elnphi, double *phi,
                                         double *props) {
    NEED(h, "null handle");
    MmctmHost &mm = h->mm;
    NEED(mm.has_state, "mmsig_mmctm_set_state first");
    CU(cudaSetDevice(h->device));
    MmctmDev &p = mm.p;
    const size_t DMK = (size_t)p.D * p.MK, MK2 = (size_t)p.MK * p.MK;
    auto d2h = [&](double *dst, const double *src, size_t n) -> cudaError_t {
        return dst ? cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream) : cudaSuccess;
    };
    CU(d2h(lambda, p.lam, DMK));
    CU(d2h(nu, p.nu, DMK));
    CU(d2h(zeta, p.zeta, (size_t)p.D * p.M));
    CU(d2h(mu, p.mu, p.MK));
    CU(d2h(Sigma, p.Sigma, MK2));
    CU(d2h(invSigma, p.invSigma, MK2));
    CU(d2h(gamma, p.gamma, mm.G));
    CU(d2h(Elnphi, p.Elnphi, mm.G));
    CU(d2h(phi, p.phi, mm.G));
    if (props) {
        if (!mm.props_scratch) {
            int rc = dev_alloc(h, h->allocs_mm, &mm.props_scratch, DMK);
            if (rc) return rc;
        }
        {
            LaunchScope ls(h, "k_props");
            if (mm.wide) k_zeta_props_wide<<<mm.grid_solve, 256, 0, h->stream>>>(p, mm.props_scratch, 0);
            else k_props<<<mm.grid_solve, 256, 0, h->stream>>>(p, mm.props_scratch);
        }
        CU(cudaMemcpyAsync(props, mm.props_scratch, DMK * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return 0;
}

// ---- fit! from / to host buffers in one call, transfers overlapped with the E-step -----------
// = set_data + set_state + fit + get_state, bit for bit.  The samples are cut into chunks; chunk
// c's counts, λ, ν travel on a copy stream while chunk c-1 runs k_pack_rows / k_theta_tile /
// k_solve (block partials accumulate over the chunks), and when the loop is known to end with
// this iteration (iter == maxiter) the chunk's λ, ν, ζ, props leave on a second copy stream while
// the next chunk computes.  MMSIG_PIPE_CHUNKS overrides the chunk count (default ~150k samples).
// MMSIG_TRACE=1: host-side wall-clock marks of mmsig_mmctm_fit_host on stderr (where an end-to-end call spends its time)
struct HostTrace {
    bool on;
    double t0, last;
    static double now() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }
    HostTrace() : on(getenv("MMSIG_TRACE") != nullptr), t0(now()), last(t0) {}
    void mark(const char *what) {
        if (!on) return;
        const double t = now();
        fprintf(stderr, "[mmsig trace] %-28s +%8.3f ms  (%8.3f ms)\n", what, t - last, t - t0);
        last = t;
    }
};

static int pipe_chunks(long long D) {
    const char *e = getenv("MMSIG_PIPE_CHUNKS");
    // ~150k samples per chunk, but at least 4 chunks once a shard is worth pipelining at all (a rank
    // of an 8-GPU run holds 125k samples of the 1M corpus)
    long long c = e ? atoll(e) : std::max<long long>((D + 75000) / 150000, std::min<long long>(4, D / 16384));
    return (int)std::max<long long>(1, std::min<long long>({c, 64LL, (D + 31) / 32}));     // a chunk holds at least one tile of 32 samples
}

static int mmctm_fit_host_impl(mmsig_handle *h, int64_t D, int64_t D_total, int32_t M, const int32_t *K,
                                        const int32_t *V, const int64_t *const *rowptr, const int32_t *const *term,
                                        const int32_t *const *count, const double *alpha, const double *gamma,
                                        const double *lambda, const double *nu, const double *mu, const double *Sigma,
                                        const double *invSigma, int32_t maxiter, double tol, uint32_t flags,
                                        double *ll_hist, int32_t *n_iter, int32_t *converged, double *lambda_out,
                                        double *nu_out, double *zeta_out, double *mu_out, double *Sigma_out,
                                        double *invSigma_out, double *gamma_out, double *Elnphi_out, double *phi_out,
                                        double *props_out, bool packed) {
    NEED(h, "null handle");
    NEED(K && V && rowptr && term && (count || packed), "null argument");
    NEED(alpha && gamma, "alpha and gamma are required");
    NEED(maxiter >= 1 && ll_hist, "maxiter >= 1 and ll_hist required");
    CU(cudaSetDevice(h->device));
    HostTrace trace;
    if (M < 1 || M > MAXM) return fail(h, MMSIG_ELIMIT, "1 <= M <= 8 modalities supported");
    NEED(D >= 1, "need 1 <= D <= D_total");
    bool same = false;
    int rc;
    long long nnz_m[MAXM];
    for (int m = 0; m < M; ++m) {
        // the row pointers are validated chunk by chunk, right before each chunk's copies are sized
        // from them (2.7 ms of host time at D = 1e6 that would otherwise precede the first copy)
        NEED(rowptr[m] && rowptr[m][0] == 0 && rowptr[m][D] >= 0, "rowptr[0] must be 0");
        nnz_m[m] = rowptr[m][D];
    }
    if ((rc = mmctm_prepare(h, D, D_total, M, K, V, nnz_m, &same))) return rc;
    trace.mark(same ? "plan reused" : "plan + allocations");
    MmctmHost &mm = h->mm;
    MmctmDev &p = mm.p;
    NEED(!p.factored, "mmsig_mmctm_fit_host takes the MMCTM's K x V state; use set_data / mmsig_immctm_set_state / fit for the IMMCTM");
    mm.has_data = false;
    mm.has_state = false;
    for (int m = 0; m < M; ++m) {
        NEED(alpha[m] > 0, "alpha must be > 0");
        NEED(rowptr[m][D] == 0 || (term[m] && (packed || count[m])), "null term / count");
        NEED(!packed || V[m] <= 1024, "packed records carry 10-bit terms");
    }
    if (!h->s_in) CU(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    if (!h->s_out) CU(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    const size_t DMK = (size_t)D * p.MK, MK2 = (size_t)p.MK * p.MK;
    const int MK = p.MK;
    if (props_out && !mm.props_scratch)
        if ((rc = dev_alloc(h, h->allocs_mm, &mm.props_scratch, DMK))) return rc;

    // small state, as mmsig_mmctm_set_state
    mm.alpha_host.assign(alpha, alpha + M);
    CU(cudaMemcpyAsync(p.alpha, alpha, M * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(p.gamma, gamma, mm.G * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (mu) CU(cudaMemcpyAsync(p.mu, mu, MK * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    else CU(cudaMemsetAsync(p.mu, 0, MK * sizeof(double), h->stream));
    std::vector<double> eye(MK2, 0.0);
    for (int j = 0; j < MK; ++j) eye[(size_t)j * MK + j] = 1.0;
    CU(cudaMemcpyAsync(p.Sigma, Sigma ? Sigma : eye.data(), MK2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(p.invSigma, invSigma ? invSigma : eye.data(), MK2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    {
        LaunchScope ls(h, "k_elnphi");
        k_elnphi<<<1, 1024, 0, h->stream>>>(p);
    }
    CU(cudaMemsetAsync(p.stats, 0, mm.G * sizeof(double), h->stream));
    for (int m = 0; m < M; ++m) CU(cudaMemsetAsync(mm.cb[m].flags, 0, 4 * sizeof(int), h->stream));
    // the first E-step reads the caller's λ as lam_prev and writes lam
    p.lam = mm.lamA;
    p.lam_prev = mm.lamB;
    if (!lambda) CU(cudaMemsetAsync(p.lam_prev, 0, DMK * sizeof(double), h->stream));
    if (!nu) fill(h, p.nu, DMK, 1.0);

    const int C = pipe_chunks(D);
    std::vector<long long> cut(C + 1);
    {
        // equal chunks by default; MMSIG_PIPE_RATIO > 1 makes them grow geometrically (a smaller first
        // upload).  Measured at D = 1e6 (profiles/chunks_sweep.sh): 6-10 chunks of ratio 1.0-1.3 all
        // give 51.6-52.5 ms per call against 45.0 ms resident; 16 / 24 chunks 54.2 / 58.6 ms.
        const char *e = getenv("MMSIG_PIPE_RATIO");
        const double ratio = e ? std::max(1.0, atof(e)) : 1.0;
        std::vector<double> w(C);
        double tot = 0.0, x = 1.0;
        for (int c = 0; c < C; ++c) { w[c] = x; tot += x; x *= ratio; }
        double acc = 0.0;
        cut[0] = 0;
        for (int c = 0; c < C; ++c) {
            acc += w[c];
            // multiples of 32: a chunk's tiles coincide with the shard's (the slot tag of a record is d mod 32)
            cut[c + 1] = c + 1 == C ? D : std::min<long long>(D, std::max<long long>(cut[c] + 32, ((long long)(D * (acc / tot)) + 31) / 32 * 32));
        }
    }
    std::vector<cudaEvent_t> ev_in(C, nullptr), ev_out(C, nullptr);
    auto free_events = [&]() {
        cudaStreamSynchronize(h->s_in);        // no copy may outlive the caller's buffers, whatever the exit path
        cudaStreamSynchronize(h->s_out);
        for (auto e : ev_in) if (e) cudaEventDestroy(e);
        for (auto e : ev_out) if (e) cudaEventDestroy(e);
    };
    struct EvGuard { decltype(free_events) &f; ~EvGuard() { f(); } } evguard{free_events};
    for (int c = 0; c < C; ++c) {
        CU(cudaEventCreateWithFlags(&ev_in[c], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ev_out[c], cudaEventDisableTiming));
    }
    // the copy stream must not run ahead of the memsets / fills above
    {
        cudaEvent_t ev0;
        CU(cudaEventCreateWithFlags(&ev0, cudaEventDisableTiming));
        CU(cudaEventRecord(ev0, h->stream));
        CU(cudaStreamWaitEvent(h->s_in, ev0, 0));
        cudaEventDestroy(ev0);
    }
    auto zero_partials = [&]() -> cudaError_t {
        for (int m = 0; m < M; ++m) {
            cudaError_t e = cudaMemsetAsync(mm.part_theta[m], 0, (size_t)mm.grid_theta[m] * p.K[m] * p.V[m] * sizeof(double2), h->stream);
            if (e != cudaSuccess) return e;
        }
        return cudaMemsetAsync(mm.part_solve, 0, (size_t)mm.grid_solve * 2 * MK * sizeof(double2), h->stream);
    };
    // E-step of chunk c done -> its outputs leave on s_out (enqueued after all compute, so that a
    // pageable destination, whose copy blocks the host, cannot hold back the launches)
    auto after_chunk = [&](int c) -> cudaError_t {
        const long long d0 = cut[c], d1 = cut[c + 1];
        if (props_out) {
            MmctmDev q = chunk_view(p, d0, d1, 0);
            LaunchScope ls(h, "k_props");
            const int grid = (int)std::max<long long>(1, std::min<long long>(mm.grid_solve, (q.D + 7) / 8));
            if (mm.wide) k_zeta_props_wide<<<grid, 256, 0, h->stream>>>(q, mm.props_scratch + d0 * MK, 0);
            else k_props<<<grid, 256, 0, h->stream>>>(q, mm.props_scratch + d0 * MK);
        }
        return cudaEventRecord(ev_out[c], h->stream);
    };
    auto drain_chunk = [&](int c) -> cudaError_t {
        const long long d0 = cut[c], d1 = cut[c + 1];
        const size_t n = (size_t)(d1 - d0) * MK * sizeof(double);
        cudaError_t e = cudaStreamWaitEvent(h->s_out, ev_out[c], 0);
        if (e == cudaSuccess && lambda_out) e = cudaMemcpyAsync(lambda_out + d0 * MK, p.lam + d0 * MK, n, cudaMemcpyDeviceToHost, h->s_out);
        if (e == cudaSuccess && nu_out) e = cudaMemcpyAsync(nu_out + d0 * MK, p.nu + d0 * MK, n, cudaMemcpyDeviceToHost, h->s_out);
        if (e == cudaSuccess && zeta_out)
            e = cudaMemcpyAsync(zeta_out + d0 * M, p.zeta + d0 * M, (size_t)(d1 - d0) * M * sizeof(double), cudaMemcpyDeviceToHost, h->s_out);
        if (e == cudaSuccess && props_out)
            e = cudaMemcpyAsync(props_out + d0 * MK, mm.props_scratch + d0 * MK, n, cudaMemcpyDeviceToHost, h->s_out);
        return e;
    };

    // ---- iteration 1: upload chunk c while chunk c-1 computes
    mm.last_unsmoothed = (flags & MMSIG_FLAG_UNSMOOTHED) != 0;
    bool streamed_out = false;
    if (C > 1) CU(zero_partials());
    for (int c = 0; c < C; ++c) {
        const long long d0 = cut[c], d1 = cut[c + 1];
        for (int m = 0; m < M; ++m) {
            for (long long d = d0; d < d1; ++d)
                if (rowptr[m][d + 1] < rowptr[m][d] || rowptr[m][d + 1] > nnz_m[m]) return fail(h, MMSIG_EINVAL, "rowptr not monotone");
            CountBuf &cb = mm.cb[m];
            const long long r0 = c == 0 ? 0 : d0 + 1;          // entry d0 came with the previous chunk
            CU(cudaMemcpyAsync(cb.rowptr + r0, rowptr[m] + r0, (size_t)(d1 + 1 - r0) * sizeof(long long), cudaMemcpyHostToDevice, h->s_in));
            const long long w0 = rowptr[m][d0], w1 = rowptr[m][d1];
            if (w1 > w0) {
                CU(cudaMemcpyAsync(cb.term + w0, term[m] + w0, (size_t)(w1 - w0) * sizeof(int), cudaMemcpyHostToDevice, h->s_in));
                if (!packed) CU(cudaMemcpyAsync(cb.count + w0, count[m] + w0, (size_t)(w1 - w0) * sizeof(int), cudaMemcpyHostToDevice, h->s_in));
            }
        }
        const size_t n = (size_t)(d1 - d0) * MK * sizeof(double);
        if (lambda) CU(cudaMemcpyAsync(p.lam_prev + d0 * MK, lambda + d0 * MK, n, cudaMemcpyHostToDevice, h->s_in));
        if (nu) CU(cudaMemcpyAsync(p.nu + d0 * MK, nu + d0 * MK, n, cudaMemcpyHostToDevice, h->s_in));
        CU(cudaEventRecord(ev_in[c], h->s_in));
        CU(cudaStreamWaitEvent(h->stream, ev_in[c], 0));
        for (int m = 0; m < M; ++m) pack_and_densify(h, mm.cb[m], d0, d1, V[m], M, m, const_cast<double *>(p.N), 1, packed);
        mmctm_estep_launch(h, chunk_view(p, d0, d1, C > 1), flags);
        if (maxiter == 1) CU(after_chunk(c));
    }
    trace.mark("chunks enqueued");
    if (maxiter == 1) {
        for (int c = 0; c < C; ++c) CU(drain_chunk(c));
        streamed_out = true;
    }
    // validation verdict and Σ_d N_dm (the LL denominators) before the M-step
    {
        int hf[MAXM][4];
        long long ntot[MAXM];
        for (int m = 0; m < M; ++m) CU(cudaMemcpyAsync(hf[m], mm.cb[m].flags, sizeof(hf[m]), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaGetLastError());
        for (int m = 0; m < M; ++m)
            if ((rc = flags_verdict(h, hf[m], &ntot[m]))) {
                cudaStreamSynchronize(h->s_out);
                return rc;
            }
        if ((rc = allsum_ll(h, ntot, M))) return rc;
        for (int m = 0; m < M; ++m) p.Ntot[m] = (double)ntot[m];
    }
    trace.mark("E-step of iteration 1 done");
    mm.has_data = true;
    auto finish_iteration = [&](double *ll) -> int {
        int r = mmctm_mstep_launch(h, flags);
        if (r) return r;
        int status = 0;
        CU(cudaMemcpyAsync(ll, mm.d_ll, M * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(&status, mm.d_status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaGetLastError());
        if (status) return fail(h, MMSIG_EINVAL, "Sigma is singular (inv failed)");
        return 0;
    };
    if ((rc = finish_iteration(ll_hist))) { cudaStreamSynchronize(h->s_out); return rc; }
    trace.mark("M-step of iteration 1 done");
    mm.has_state = true;

    // ---- iterations 2 .. maxiter (src/MMCTM.jl:462-487)
    int it = 1, conv = 0;
    int iter0 = 2;
    {
        // iterations 2 .. min(10, maxiter - 1) cannot end the loop (src/MMCTM.jl:485): enqueue them without host round trips
        const int last_free = std::min<int>(10, maxiter - (C > 1 ? 1 : 0));
        if (last_free >= 3 && !(flags & MMSIG_FLAG_AUTO_ALPHA)) {
            int cv = 0;
            if ((rc = mmctm_run_iterations(h, 2, last_free, tol, flags, ll_hist, &it, &cv))) return rc;
            iter0 = last_free + 1;
        }
    }
    for (int iter = iter0; iter <= maxiter; ++iter) {
        double *ll = ll_hist + (size_t)(iter - 1) * M;
        if (iter == maxiter && C > 1) {
            // the loop ends here whatever the LL says: let the outputs leave chunk by chunk
            std::swap(p.lam, p.lam_prev);
            CU(zero_partials());
            for (int c = 0; c < C; ++c) {
                mmctm_estep_launch(h, chunk_view(p, cut[c], cut[c + 1], 1), flags);
                CU(after_chunk(c));
            }
            for (int c = 0; c < C; ++c) CU(drain_chunk(c));
            streamed_out = true;
            if ((rc = finish_iteration(ll))) { cudaStreamSynchronize(h->s_out); return rc; }
        } else if ((rc = mmsig_mmctm_iterate(h, flags, ll))) return rc;
        it = iter;
        if (iter > 10 && converged_vec(ll - M, ll, M, tol)) { conv = 1; break; }   // src/MMCTM.jl:485
    }
    if (n_iter) *n_iter = it;
    if (converged) *converged = conv;
    CU(cudaStreamSynchronize(h->s_out));
    trace.mark("loop done, outputs drained");
    if (streamed_out)
        return mmsig_mmctm_get_state(h, nullptr, nullptr, nullptr, mu_out, Sigma_out, invSigma_out, gamma_out, Elnphi_out, phi_out, nullptr);
    return mmsig_mmctm_get_state(h, lambda_out, nu_out, zeta_out, mu_out, Sigma_out, invSigma_out, gamma_out, Elnphi_out,
                                 phi_out, props_out);
}

extern "C" int32_t mmsig_mmctm_fit_host(mmsig_handle *h, int64_t D, int64_t D_total, int32_t M, const int32_t *K,
                                        const int32_t *V, const int64_t *const *rowptr, const int32_t *const *term,
                                        const int32_t *const *count, const double *alpha, const double *gamma,
                                        const double *lambda, const double *nu, const double *mu, const double *Sigma,
                                        const double *invSigma, int32_t maxiter, double tol, uint32_t flags,
                                        double *ll_hist, int32_t *n_iter, int32_t *converged, double *lambda_out,
                                        double *nu_out, double *zeta_out, double *mu_out, double *Sigma_out,
                                        double *invSigma_out, double *gamma_out, double *Elnphi_out, double *phi_out,
                                        double *props_out) {
    return mmctm_fit_host_impl(h, D, D_total, M, K, V, rowptr, term, count, alpha, gamma, lambda, nu, mu, Sigma, invSigma, maxiter, tol,
                               flags, ll_hist, n_iter, converged, lambda_out, nu_out, zeta_out, mu_out, Sigma_out, invSigma_out,
                               gamma_out, Elnphi_out, phi_out, props_out, false);
}

// the same with 4-byte records (half the host -> device traffic of the counts): rec[m][w] = term | count << 10
extern "C" int32_t mmsig_mmctm_fit_host_packed(mmsig_handle *h, int64_t D, int64_t D_total, int32_t M, const int32_t *K,
                                               const int32_t *V, const int64_t *const *rowptr, const uint32_t *const *rec,
                                               const double *alpha, const double *gamma, const double *lambda, const double *nu,
                                               const double *mu, const double *Sigma, const double *invSigma, int32_t maxiter,
                                               double tol, uint32_t flags, double *ll_hist, int32_t *n_iter, int32_t *converged,
                                               double *lambda_out, double *nu_out, double *zeta_out, double *mu_out,
                                               double *Sigma_out, double *invSigma_out, double *gamma_out, double *Elnphi_out,
                                               double *phi_out, double *props_out) {
    return mmctm_fit_host_impl(h, D, D_total, M, K, V, rowptr, reinterpret_cast<const int32_t *const *>(rec), nullptr, alpha, gamma,
                               lambda, nu, mu, Sigma, invSigma, maxiter, tol, flags, ll_hist, n_iter, converged, lambda_out, nu_out,
                               zeta_out, mu_out, Sigma_out, invSigma_out, gamma_out, Elnphi_out, phi_out, props_out, true);
}

// (term, count) -> packed records for mmsig_mmctm_fit_host_packed; MMSIG_ELIMIT when a term needs more than 10 bits
// or a count more than 22 (use the unpacked entry then).  Host only.
extern "C" int32_t mmsig_pack_records(int64_t nnz, const int32_t *term, const int32_t *count, uint32_t *rec) {
    mmsig_handle *h = nullptr;
    NEED(nnz >= 0 && (nnz == 0 || (term && count && rec)), "null argument");
    unsigned bad = 0;
    for (int64_t w = 0; w < nnz; ++w) {
        const unsigned t = (unsigned)term[w], c = (unsigned)count[w];
        bad |= (t >> 10) | (c >> 22);
        rec[w] = t | (c << 10);
    }
    if (bad) return fail(h, MMSIG_ELIMIT, "packed records hold terms < 1024 and counts < 4194304");
    return 0;
}

extern "C" int32_t mmsig_mmctm_get_theta(mmsig_handle *h, int32_t m, double *theta_out) {
    NEED(h && theta_out, "null argument");
    MmctmHost &mm = h->mm;
    NEED(mm.has_state, "mmsig_mmctm_set_state first");
    NEED(m >= 0 && m < mm.p.M, "bad modality");
    CU(cudaSetDevice(h->device));
    MmctmDev &p = mm.p;
    const size_t n = (size_t)mm.nnz[m] * p.K[m];
    double *d = nullptr;
    CU(cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(double)));
    const size_t smem = (size_t)p.K[m] * p.V[m] * sizeof(double);
    CU(allow_max_smem(h, k_theta_out));
    {
        LaunchScope ls(h, "k_theta_out");
        k_theta_out<<<mm.grid_solve, 256, smem, h->stream>>>(p, m, d, mm.last_unsmoothed ? 1 : 0);
    }
    CU(cudaMemcpyAsync(theta_out, d, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(d);
    CU(cudaGetLastError());
    return 0;
}

extern "C" int32_t mmsig_mmctm_get_evals(mmsig_handle *h, int32_t *nev_nu, int32_t *nev_lambda) {
    NEED(h, "null handle");
    NEED(h->mm.has_state, "mmsig_mmctm_set_state first");
    CU(cudaSetDevice(h->device));
    MmctmDev &p = h->mm.p;
    if (nev_nu) CU(cudaMemcpyAsync(nev_nu, p.nev_nu, p.D * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (nev_lambda) CU(cudaMemcpyAsync(nev_lambda, p.nev_lam, p.D * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

#include "elbo_api.inl"

extern "C" int32_t mmsig_mmctm_elbo(mmsig_handle *h, double *elbo, double *terms) {
    NEED(h, "null handle");
    MmctmHost &mm = h->mm;
    NEED(mm.has_state, "mmsig_mmctm_set_state first");
    CU(cudaSetDevice(h->device));
    return mmctm_elbo_impl(h, elbo, terms);
}

// ---- test hook -----------------------------------------------------------------------------------
__global__ void k_debug_math(int fn, long long n, const double *x, double *y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = fn == 0 ? det_exp(x[i]) : fn == 1 ? det_log(x[i]) : fn == 2 ? det_digamma(x[i])
             : fn == 3 ? fast_div(x[i], x[n + i]) : fn == 4 ? fast_rcp(x[i]) : fast_sqrt(x[i]);
}
extern "C" int32_t mmsig_debug_math(mmsig_handle *h, int32_t fn, int64_t n, const double *x, double *y) {
    NEED(h && x && y && n >= 0 && fn >= 0 && fn <= 5, "bad argument");
    CU(cudaSetDevice(h->device));
    double *dx = nullptr, *dy = nullptr;
    const int64_t nin = fn == 3 ? 2 * n : n;             // fn 3 reads numerators x[0..n) and denominators x[n..2n)
    CU(cudaMalloc(&dx, std::max<int64_t>(nin, 1) * sizeof(double)));
    CU(cudaMalloc(&dy, std::max<int64_t>(n, 1) * sizeof(double)));
    CU(cudaMemcpyAsync(dx, x, nin * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    {
        LaunchScope ls(h, "k_debug_math");
        k_debug_math<<<h->numSM, 256, 0, h->stream>>>(fn, n, dx, dy);
    }
    CU(cudaMemcpyAsync(y, dy, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(dx);
    cudaFree(dy);
    CU(cudaGetLastError());
    return 0;
}

// ---- IMMCTM (reference src/IMMCTM.jl): feature-factorised topics on the MMCTM path -----------------
// Every per-sample kernel is the MMCTM's over composite K x V tables; only the M-step over the
// feature tables (k_imstep1) and two table terms of the ELBO differ.
extern "C" int32_t mmsig_immctm_set_features(mmsig_handle *h, const int32_t *nfeat, const int32_t *const *features) {
    NEED(h && nfeat && features, "null argument");
    MmctmHost &mm = h->mm;
    NEED(mm.has_data, "mmsig_mmctm_set_data first");
    NEED(!mm.p.factored, "features are already set for this corpus");
    CU(cudaSetDevice(h->device));
    MmctmDev &p = mm.p;
    std::vector<int> ent_row, row_off, row_len, row_alpha, row_m;
    p.foff[0] = 0;
    p.aoff[0] = 0;
    int rc;
    for (int m = 0; m < p.M; ++m) {
        const int nf = nfeat[m], V = p.V[m];
        NEED(nf >= 1 && nf <= 16 && features[m], "1 <= features per modality <= 16");
        std::vector<int> J(nf, 0);
        for (int v = 0; v < V; ++v)
            for (int f = 0; f < nf; ++f) {
                const int x = features[m][(size_t)v * nf + f];
                NEED(x >= 0, "feature values are 0-based and non-negative");
                J[f] = std::max(J[f], x + 1);               // J = maximum(features, dims=1), src/IMMCTM.jl:44
            }
        int *dfeat = nullptr;
        if ((rc = dev_alloc(h, h->allocs_mm, &dfeat, (size_t)V * nf))) return rc;
        CU(cudaMemcpyAsync(dfeat, features[m], (size_t)V * nf * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        p.feat[m] = dfeat;
        p.nfeat[m] = nf;
        int t = p.foff[m];
        for (int k = 0; k < p.K[m]; ++k)
            for (int f = 0; f < nf; ++f) {
                row_off.push_back(t);
                row_len.push_back(J[f]);
                row_alpha.push_back(p.aoff[m] + f);
                row_m.push_back(m);
                for (int j = 0; j < J[f]; ++j) ent_row.push_back((int)row_off.size() - 1);
                t += J[f];
            }
        p.foff[m + 1] = t;
        p.aoff[m + 1] = p.aoff[m] + nf;
    }
    p.T = p.foff[p.M];
    p.R = (int)row_off.size();
    if ((size_t)2 * p.R * sizeof(double) > 48 * 1024) return fail(h, MMSIG_ELIMIT, "too many feature-table rows");
    auto up = [&](const std::vector<int> &v, const int **dst) -> int {
        int *d = nullptr;
        int r = dev_alloc(h, h->allocs_mm, &d, v.size());
        if (r) return r;
        if (cudaMemcpyAsync(d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream) != cudaSuccess)
            return fail(h, MMSIG_ECUDA, "cudaMemcpyAsync (feature index tables)");
        *dst = d;
        return 0;
    };
    if ((rc = up(ent_row, &p.ent_row)) || (rc = up(row_off, &p.row_off)) || (rc = up(row_len, &p.row_len)) ||
        (rc = up(row_alpha, &p.row_alpha)))
        return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &p.gammaf, (size_t)p.T))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &p.Elnphif, (size_t)p.T))) return rc;
    if ((rc = dev_alloc(h, h->allocs_mm, &p.alphaf, (size_t)p.aoff[p.M]))) return rc;
    CU(cudaStreamSynchronize(h->stream));            // the index vectors go out of scope
    mm.row_len_host = row_len;
    mm.row_m_host = row_m;
    p.factored = 1;
    mm.has_state = false;
    return 0;
}

// model.α ([m][i]), γ ([m][k][i][j] flat), λ, ν, μ, Σ, invΣ; NULL => the constructor's value (src/IMMCTM.jl:48-77)
extern "C" int32_t mmsig_immctm_set_state(mmsig_handle *h, const double *alphaf, const double *gammaf, const double *lambda,
                                          const double *nu, const double *mu, const double *Sigma, const double *invSigma) {
    NEED(h && alphaf && gammaf, "alpha and gamma are required");
    MmctmHost &mm = h->mm;
    NEED(mm.has_data && mm.p.factored, "mmsig_mmctm_set_data and mmsig_immctm_set_features first");
    CU(cudaSetDevice(h->device));
    MmctmDev &p = mm.p;
    const int nA = p.aoff[p.M];
    for (int i = 0; i < nA; ++i) NEED(alphaf[i] > 0, "alpha must be > 0");
    mm.alphaf_host.assign(alphaf, alphaf + nA);
    // the K x V gamma of the MMCTM state call is a placeholder here: k_icompose overwrites every table it seeds
    std::vector<double> a0(p.M, 1.0), g0((size_t)mm.G, 1.0);
    p.factored = 0;                                   // the plain call derives Elnphi from the placeholder ...
    int rc = mmsig_mmctm_set_state(h, a0.data(), g0.data(), lambda, nu, mu, Sigma, invSigma);
    p.factored = 1;
    if (rc) return rc;
    mm.has_state = false;
    CU(cudaMemcpyAsync(p.alphaf, alphaf, nA * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(p.gammaf, gammaf, p.T * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    {
        LaunchScope ls(h, "k_icompose");              // ... and this replaces it by the composite tables (:69-70)
        k_icompose<<<1, 1024, (size_t)2 * p.R * sizeof(double), h->stream>>>(p);
    }
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    mm.has_state = true;
    return 0;
}

extern "C" int32_t mmsig_immctm_get_tables(mmsig_handle *h, double *gammaf, double *Elnphif, double *alphaf) {
    NEED(h, "null handle");
    MmctmHost &mm = h->mm;
    NEED(mm.has_state && mm.p.factored, "mmsig_immctm_set_state first");
    CU(cudaSetDevice(h->device));
    MmctmDev &p = mm.p;
    if (gammaf) CU(cudaMemcpyAsync(gammaf, p.gammaf, p.T * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (Elnphif) CU(cudaMemcpyAsync(Elnphif, p.Elnphif, p.T * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (alphaf) memcpy(alphaf, mm.alphaf_host.data(), mm.alphaf_host.size() * sizeof(double));
    return 0;
}

// ---- restarts (config 5) ----------------------------------------------------------------------
extern "C" int32_t mmsig_mmctm_restarts(mmsig_handle *h, int32_t R, const double *gamma0, int32_t maxiter, double tol,
                                        uint32_t flags, double *elbo_out, double *ll_out, int32_t *n_iter_out,
                                        int32_t *best) {
    NEED(h, "null handle");
    MmctmHost &mm = h->mm;
    NEED(mm.has_state, "mmsig_mmctm_set_state first (it provides alpha)");
    NEED(R >= 1 && gamma0 && maxiter >= 1, "R >= 1, gamma0 and maxiter >= 1 required");
    CU(cudaSetDevice(h->device));
    MmctmDev &p = mm.p;
    const size_t DMK = (size_t)p.D * p.MK, G = mm.G, MK2 = (size_t)p.MK * p.MK, DM = (size_t)p.D * p.M;
    const size_t GT = p.factored ? (size_t)p.T : G;       // a restart's gamma0: K x V tables, or the IMMCTM's [m][k][i][j] tables
    // snapshot buffers for the best restart
    struct Snap { double **live; size_t n; double *copy; };
    std::vector<Snap> snaps = {{&p.lam, DMK, nullptr}, {&p.lam_prev, DMK, nullptr}, {&p.nu, DMK, nullptr},
                               {&p.sumtheta, DMK, nullptr}, {&p.zeta, DM, nullptr}, {&p.gamma, G, nullptr},
                               {&p.Elnphi, G, nullptr}, {&p.Elnphi_prev, G, nullptr}, {&p.phi, G, nullptr},
                               {&p.stats, G, nullptr}, {&p.mu, (size_t)p.MK, nullptr}, {&p.Sigma, MK2, nullptr},
                               {&p.invSigma, MK2, nullptr}};
    if (p.factored) {
        snaps.push_back({&p.gammaf, (size_t)p.T, nullptr});
        snaps.push_back({&p.Elnphif, (size_t)p.T, nullptr});
    }
    // the snapshot buffers live with the plan: a cudaMalloc / cudaFree pair per call costs up to
    // 0.9 s of driver time for the frees alone (measured, MMSIG_TRACE) against 0.6 s of fitting
    if (R > 1)
        for (size_t i = 0; i < snaps.size(); ++i) {
            if (!mm.snap[i]) {
                int rca = dev_alloc(h, h->allocs_mm, &mm.snap[i], snaps[i].n);
                if (rca) return rca;
            }
            snaps[i].copy = mm.snap[i];
        }
    std::vector<double> hist((size_t)maxiter * p.M);
    std::vector<double> alpha = mm.alpha_host, alphaf = mm.alphaf_host, best_alpha = alpha, best_alphaf = alphaf;
    HostTrace trace;
    trace.mark("restarts: snapshot buffers");
    int best_r = -1;
    double best_e = 0.0;
    int rc = 0;
    for (int r = 0; r < R && !rc; ++r) {
        rc = p.factored ? mmsig_immctm_set_state(h, alphaf.data(), gamma0 + (size_t)r * GT, nullptr, nullptr, nullptr, nullptr, nullptr)
                        : mmsig_mmctm_set_state(h, alpha.data(), gamma0 + (size_t)r * GT, nullptr, nullptr, nullptr, nullptr, nullptr);
        if (rc) break;
        int nit = 0, conv = 0;
        trace.mark("restart: state set");
        rc = mmsig_mmctm_fit(h, maxiter, tol, flags, hist.data(), &nit, &conv);
        if (rc) break;
        trace.mark("restart: fit");
        double e = 0.0;
        rc = mmsig_mmctm_elbo(h, &e, nullptr);
        if (rc) break;
        trace.mark("restart: elbo");
        if (elbo_out) elbo_out[r] = e;
        if (n_iter_out) n_iter_out[r] = nit;
        if (ll_out) memcpy(ll_out + (size_t)r * p.M, hist.data() + (size_t)(nit - 1) * p.M, p.M * sizeof(double));
        if (best_r < 0 || e > best_e || best_e != best_e) {           // arg-max ELBO, first wins ties
            best_r = r;
            best_e = e;
            best_alpha = mm.alpha_host;                               // autoα: every restart re-optimises α from the caller's
            best_alphaf = mm.alphaf_host;
            if (R > 1)
                for (auto &s : snaps)
                    cudaMemcpyAsync(s.copy, *s.live, s.n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
        }
    }
    if (!rc && R > 1 && best_r != R - 1) {
        for (auto &s : snaps) cudaMemcpyAsync(*s.live, s.copy, s.n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
        // ... and the best restart's α with its tables (mmsig_mmctm_get_alpha / _elbo / _iterate afterwards)
        mm.alpha_host = best_alpha;
        mm.alphaf_host = best_alphaf;
        if (!best_alpha.empty()) cudaMemcpyAsync(p.alpha, mm.alpha_host.data(), std::min<size_t>(best_alpha.size(), p.M) * sizeof(double), cudaMemcpyHostToDevice, h->stream);
        if (p.factored && p.alphaf && !best_alphaf.empty())
            cudaMemcpyAsync(p.alphaf, mm.alphaf_host.data(), best_alphaf.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    }
    cudaStreamSynchronize(h->stream);
    trace.mark("restarts: best state restored");
    if (rc) return rc;
    CU(cudaGetLastError());
    if (best) *best = best_r;
    return 0;
}

#include "ingest_api.inl"
#include "group_api.inl"

extern "C" int32_t mmsig_lda_iterate_flags(mmsig_handle *h, uint32_t flags, double *ll_out);
#include "lda_api.inl"
