// runtime.cu -- context, error channel, timers, pinned/device buffers, L2 flush and the NCCL count reduction.
#include <dlfcn.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdlib.h>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace bxg {

static Context g_ctx;
static thread_local char g_err[1024] = "";

Context &ctx() { return g_ctx; }

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int ensure_init() {
    if (g_ctx.device >= 0) return BXG_OK;
    return bxg_init(0);
}

int scratch(int slot, size_t bytes, void **out) {
    Context &c = ctx();
    if (bytes > c.scratch_cap[slot]) {
        if (c.scratch[slot]) {
            BXG_CUDA(cudaStreamSynchronize(c.stream));
            BXG_CUDA(cudaFree(c.scratch[slot]));
            c.scratch[slot] = nullptr;
            c.scratch_cap[slot] = 0;
        }
        size_t cap = bytes + bytes / 4 + 256;
        BXG_CUDA(cudaMalloc(&c.scratch[slot], cap));
        c.scratch_cap[slot] = cap;
    }
    *out = c.scratch[slot];
    return BXG_OK;
}

int stage_in(int slot, const void *src, size_t bytes, int loc, const void **dptr) {
    if (loc == BXG_DEVICE || bytes == 0) {
        *dptr = src;
        return BXG_OK;
    }
    void *d = nullptr;
    BXG_TRY(scratch(slot, bytes, &d));
    BXG_CUDA(cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, ctx().stream));
    *dptr = d;
    return BXG_OK;
}

int zc_wait(long long seq) {
    Context &c = ctx();
    volatile long long *flag = (volatile long long *)(c.zc + ZC_FLAG_OFFSET);
    for (int spin = 0; spin < 200000; spin++) {
        if (*flag == seq) return BXG_OK;
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    BXG_CUDA(cudaStreamSynchronize(c.stream));          // slow kernel or a launch error: let the runtime tell
    if (*flag != seq) return set_error(BXG_ERR_CUDA, "scalar-call kernel did not complete");
    return BXG_OK;
}

// ---- per-kernel event profiler ------------------------------------------------------------------------------------
struct ProfRec {
    const char *name;
    cudaEvent_t a, b;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_event_pool;

static cudaEvent_t prof_event() {
    if (!g_event_pool.empty()) {
        cudaEvent_t e = g_event_pool.back();
        g_event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

bool prof_enabled() { return g_prof_on; }

void prof_begin(const char *name) {
    if (!g_prof_on) return;
    ProfRec r{name, prof_event(), prof_event()};
    cudaEventRecord(r.a, ctx().stream);
    g_prof.push_back(r);
}
void prof_end() {
    if (!g_prof_on || g_prof.empty()) return;
    cudaEventRecord(g_prof.back().b, ctx().stream);
}

}  // namespace bxg

using namespace bxg;

extern "C" {

const char *bxg_last_error(void) { return g_err; }
const char *bxg_version(void) { return "bxb200 0.1.0 (sm_100a)"; }

int bxg_device_count(int *n) {
    int k = 0;
    cudaError_t e = cudaGetDeviceCount(&k);
    if (e != cudaSuccess) {
        *n = 0;
        return set_error(BXG_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    }
    *n = k;
    return BXG_OK;
}

int bxg_init(int device) {
    Context &c = g_ctx;
    int n = 0;
    BXG_TRY(bxg_device_count(&n));
    if (n <= 0) return set_error(BXG_ERR_CUDA, "no CUDA device visible (libbxb200 has no CPU fallback)");
    if (device < 0 || device >= n) return set_error(BXG_ERR_ARG, "device %d out of range [0,%d)", device, n);
    if (c.device == device && c.stream) return BXG_OK;
    if (c.device >= 0 && c.device != device)
        return set_error(BXG_ERR_STATE, "library already bound to device %d (one process per GPU)", c.device);
    BXG_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    BXG_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major < 10)
        return set_error(BXG_ERR_CUDA, "device %d is sm_%d%d; libbxb200 is built for sm_100a only", device, p.major, p.minor);
    c.sm_count = p.multiProcessorCount;
    c.l2_bytes = p.l2CacheSize;
    {
        // The query kernels of this library (find, count_range, set_range) read single 32-byte sectors at random from
        // working sets larger than L2, so the smallest L2 fetch granularity is asked for.  Measured on B200 (r02j): the
        // limit is accepted (32 / 64 / 128 read back) and changes neither time nor DRAM sectors of those kernels;
        // kept because it is the documented intent.  BXB200_L2_FETCH=64|128 for A/B runs, BXB200_DEBUG prints the value.
        const char *e = getenv("BXB200_L2_FETCH");
        const size_t g = e ? (size_t)atoi(e) : 32;
        if (g == 32 || g == 64 || g == 128) {
            cudaError_t le = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g);
            if (le != cudaSuccess) cudaGetLastError();
            if (getenv("BXB200_DEBUG")) {
                size_t got = 0;
                cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
                fprintf(stderr, "[bxb200] cudaLimitMaxL2FetchGranularity: asked %zu -> %s, now %zu\n", g,
                        cudaGetErrorString(le), got);
            }
        }
    }
    BXG_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    BXG_CUDA(cudaMallocHost(&c.mailbox, 64 * sizeof(int64_t)));
    BXG_CUDA(cudaMalloc(&c.d_mailbox, 64 * sizeof(int64_t)));
    BXG_CUDA(cudaMemset(c.d_mailbox, 0, 64 * sizeof(int64_t)));
    if (cudaHostAlloc((void **)&c.zc, 4096, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void **)&c.zc_dev, c.zc, 0) != cudaSuccess) {
        cudaGetLastError();                    // no mapped memory on this platform: scalar calls use the staged path
        c.zc = c.zc_dev = nullptr;
    }
    if (c.zc) memset(c.zc, 0, 4096);
    c.device = device;
    c.launches = 0;
    return BXG_OK;
}

int bxg_device_info(char *name, int name_cap, int *sm_count, int64_t *total_mem, int *cc_major, int *cc_minor) {
    BXG_TRY(ensure_init());
    cudaDeviceProp p;
    BXG_CUDA(cudaGetDeviceProperties(&p, g_ctx.device));
    if (name && name_cap > 0) {
        strncpy(name, p.name, (size_t)name_cap - 1);
        name[name_cap - 1] = 0;
    }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (total_mem) *total_mem = (int64_t)p.totalGlobalMem;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return BXG_OK;
}

int bxg_device_pci_bus_id(char *out, int cap) {
    BXG_TRY(ensure_init());
    if (!out || cap < 16) return set_error(BXG_ERR_ARG, "buffer too small");
    BXG_CUDA(cudaDeviceGetPCIBusId(out, cap, g_ctx.device));
    return BXG_OK;
}

int bxg_sync(void) {
    BXG_TRY(ensure_init());
    BXG_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return BXG_OK;
}

int bxg_launch_count(int64_t *n) {
    *n = g_ctx.launches;
    return BXG_OK;
}
int bxg_launch_count_reset(void) {
    g_ctx.launches = 0;
    return BXG_OK;
}

int bxg_host_alloc(int64_t bytes, void **ptr) {
    BXG_TRY(ensure_init());
    BXG_CUDA(cudaMallocHost(ptr, (size_t)(bytes > 0 ? bytes : 1)));
    return BXG_OK;
}
int bxg_host_free(void *ptr) {
    if (ptr) BXG_CUDA(cudaFreeHost(ptr));
    return BXG_OK;
}
int bxg_dev_alloc(int64_t bytes, void **dptr) {
    BXG_TRY(ensure_init());
    BXG_CUDA(cudaMalloc(dptr, (size_t)(bytes > 0 ? bytes : 1)));
    return BXG_OK;
}
int bxg_dev_free(void *dptr) {
    if (dptr) {
        BXG_CUDA(cudaStreamSynchronize(g_ctx.stream));
        BXG_CUDA(cudaFree(dptr));
    }
    return BXG_OK;
}
int bxg_dev_memset(void *dptr, int value, int64_t bytes) {
    BXG_TRY(ensure_init());
    BXG_CUDA(cudaMemsetAsync(dptr, value, (size_t)bytes, g_ctx.stream));
    return BXG_OK;
}
int bxg_memcpy_h2d(void *dptr, const void *hptr, int64_t bytes) {
    BXG_TRY(ensure_init());
    BXG_CUDA(cudaMemcpyAsync(dptr, hptr, (size_t)bytes, cudaMemcpyHostToDevice, g_ctx.stream));
    return BXG_OK;
}
int bxg_memcpy_d2h(void *hptr, const void *dptr, int64_t bytes) {
    BXG_TRY(ensure_init());
    BXG_CUDA(cudaMemcpyAsync(hptr, dptr, (size_t)bytes, cudaMemcpyDeviceToHost, g_ctx.stream));
    return BXG_OK;
}

struct bxg_timer {
    cudaEvent_t a, b;
};

int bxg_timer_create(bxg_timer_t **t) {
    BXG_TRY(ensure_init());
    bxg_timer *x = new bxg_timer();
    BXG_CUDA(cudaEventCreate(&x->a));
    BXG_CUDA(cudaEventCreate(&x->b));
    *t = x;
    return BXG_OK;
}
int bxg_timer_free(bxg_timer_t *t) {
    if (!t) return BXG_OK;
    cudaEventDestroy(t->a);
    cudaEventDestroy(t->b);
    delete t;
    return BXG_OK;
}
int bxg_timer_start(bxg_timer_t *t) {
    BXG_CUDA(cudaEventRecord(t->a, g_ctx.stream));
    return BXG_OK;
}
int bxg_timer_stop(bxg_timer_t *t) {
    BXG_CUDA(cudaEventRecord(t->b, g_ctx.stream));
    return BXG_OK;
}
int bxg_timer_elapsed_ms(bxg_timer_t *t, float *ms) {
    BXG_CUDA(cudaEventSynchronize(t->b));
    BXG_CUDA(cudaEventElapsedTime(ms, t->a, t->b));
    return BXG_OK;
}

int bxg_profile_enable(int on) {
    BXG_TRY(ensure_init());
    BXG_CUDA(cudaStreamSynchronize(g_ctx.stream));
    for (auto &r : g_prof) {
        g_event_pool.push_back(r.a);
        g_event_pool.push_back(r.b);
    }
    g_prof.clear();
    g_prof_on = on != 0;
    return BXG_OK;
}

// one line per kernel: "<name>\t<launches>\t<total_ms>\n"
int bxg_profile_report(char *buf, int64_t cap) {
    BXG_TRY(ensure_init());
    BXG_CUDA(cudaStreamSynchronize(g_ctx.stream));
    std::map<std::string, std::pair<int64_t, double>> agg;
    std::vector<std::string> order;
    for (auto &r : g_prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) ms = 0.f;
        auto it = agg.find(r.name);
        if (it == agg.end()) {
            order.push_back(r.name);
            agg[r.name] = {1, (double)ms};
        } else {
            it->second.first++;
            it->second.second += ms;
        }
    }
    std::string out;
    char line[512];
    for (auto &n : order) {
        snprintf(line, sizeof(line), "%s\t%lld\t%.6f\n", n.c_str(), (long long)agg[n].first, agg[n].second);
        out += line;
    }
    if ((int64_t)out.size() + 1 > cap) return set_error(BXG_ERR_ARG, "profile buffer too small (%zu needed)", out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return BXG_OK;
}

// Pinned-memory copy rates of this process's GPU: host-to-device alone, device-to-host alone, and both directions at once
// (two streams).  bench.py calls it on every rank at the same time (after a barrier) to measure what the box's PCIe /
// host-memory side can sustain when all GPUs copy together -- the ceiling of the end-to-end figure.
int bxg_copy_probe(int64_t bytes, int reps, double *h2d_gbs, double *d2h_gbs, double *bidir_gbs) {
    BXG_TRY(ensure_init());
    if (bytes < 4096 || reps < 1) return set_error(BXG_ERR_ARG, "bad probe size");
    void *h0 = nullptr, *h1 = nullptr, *d0 = nullptr, *d1 = nullptr;
    cudaStream_t s0 = nullptr, s1 = nullptr;
    cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
    int rc = BXG_OK;
    auto fail = [&](cudaError_t err, const char *what) {
        rc = set_error(BXG_ERR_CUDA, "copy probe: %s failed: %s", what, cudaGetErrorString(err));
    };
    cudaError_t err;
    do {
        if ((err = cudaMallocHost(&h0, (size_t)bytes)) != cudaSuccess) { fail(err, "cudaMallocHost"); break; }
        if ((err = cudaMallocHost(&h1, (size_t)bytes)) != cudaSuccess) { fail(err, "cudaMallocHost"); break; }
        if ((err = cudaMalloc(&d0, (size_t)bytes)) != cudaSuccess) { fail(err, "cudaMalloc"); break; }
        if ((err = cudaMalloc(&d1, (size_t)bytes)) != cudaSuccess) { fail(err, "cudaMalloc"); break; }
        memset(h0, 1, (size_t)bytes);                 // first touch on this process's node
        memset(h1, 2, (size_t)bytes);
        cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking);
        cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
        for (auto &x : e) cudaEventCreate(&x);
        cudaMemcpyAsync(d0, h0, (size_t)bytes, cudaMemcpyHostToDevice, s0);      // warm-up both directions
        cudaMemcpyAsync(h1, d1, (size_t)bytes, cudaMemcpyDeviceToHost, s1);
        cudaStreamSynchronize(s0);
        cudaStreamSynchronize(s1);
        float ms = 0.f;
        cudaEventRecord(e[0], s0);
        for (int r = 0; r < reps; r++) cudaMemcpyAsync(d0, h0, (size_t)bytes, cudaMemcpyHostToDevice, s0);
        cudaEventRecord(e[1], s0);
        cudaEventSynchronize(e[1]);
        cudaEventElapsedTime(&ms, e[0], e[1]);
        if (h2d_gbs) *h2d_gbs = (double)bytes * reps / (ms * 1e-3) / 1e9;
        cudaEventRecord(e[0], s1);
        for (int r = 0; r < reps; r++) cudaMemcpyAsync(h1, d1, (size_t)bytes, cudaMemcpyDeviceToHost, s1);
        cudaEventRecord(e[1], s1);
        cudaEventSynchronize(e[1]);
        cudaEventElapsedTime(&ms, e[0], e[1]);
        if (d2h_gbs) *d2h_gbs = (double)bytes * reps / (ms * 1e-3) / 1e9;
        // both directions at once: wall time from the common start to the later finish
        cudaEventRecord(e[0], s0);
        cudaStreamWaitEvent(s1, e[0], 0);
        for (int r = 0; r < reps; r++) {
            cudaMemcpyAsync(d0, h0, (size_t)bytes, cudaMemcpyHostToDevice, s0);
            cudaMemcpyAsync(h1, d1, (size_t)bytes, cudaMemcpyDeviceToHost, s1);
        }
        cudaEventRecord(e[1], s0);
        cudaEventRecord(e[2], s1);
        cudaEventSynchronize(e[1]);
        cudaEventSynchronize(e[2]);
        float m0 = 0.f, m1 = 0.f;
        cudaEventElapsedTime(&m0, e[0], e[1]);
        cudaEventElapsedTime(&m1, e[0], e[2]);
        if (bidir_gbs) *bidir_gbs = 2.0 * (double)bytes * reps / ((m0 > m1 ? m0 : m1) * 1e-3) / 1e9;
        if ((err = cudaGetLastError()) != cudaSuccess) fail(err, "copies");
    } while (0);
    for (auto &x : e) if (x) cudaEventDestroy(x);
    if (s0) cudaStreamDestroy(s0);
    if (s1) cudaStreamDestroy(s1);
    cudaFree(d0); cudaFree(d1);
    if (h0) cudaFreeHost(h0);
    if (h1) cudaFreeHost(h1);
    return rc;
}

int bxg_l2_flush(void) {
    BXG_TRY(ensure_init());
    Context &c = g_ctx;
    if (!c.l2_flush_buf) {
        c.l2_flush_bytes = (size_t)(c.l2_bytes > 0 ? c.l2_bytes : (128ll << 20)) * 2;
        BXG_CUDA(cudaMalloc(&c.l2_flush_buf, c.l2_flush_bytes));
    }
    BXG_CUDA(cudaMemsetAsync(c.l2_flush_buf, 0x5a, c.l2_flush_bytes, c.stream));
    return BXG_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// NCCL (dlopen on first use: the soname resolves to whatever libnccl.so.2 the process already has, else the system one)
// ------------------------------------------------------------------------------------------------------------------
namespace {
struct Nccl {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclComm_t comm = nullptr;
    int nranks = 0, rank = 0;
    void *d_buf = nullptr;
    size_t d_cap = 0;
} g_nccl;

int nccl_load() {
    if (g_nccl.h) return BXG_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return set_error(BXG_ERR_NCCL, "dlopen(libnccl.so.2) failed: %s", dlerror());
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
        return set_error(BXG_ERR_NCCL, "libnccl.so.2 lacks required symbols");
    g_nccl.h = h;
    return BXG_OK;
}

#define BXG_NCCL(call)                                                                               \
    do {                                                                                             \
        ncclResult_t r__ = (call);                                                                   \
        if (r__ != ncclSuccess)                                                                      \
            return set_error(BXG_ERR_NCCL, "%s failed: %s", #call,                                   \
                             g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?");              \
    } while (0)

int allreduce(void *host, size_t n, size_t elt, ncclDataType_t dt, ncclRedOp_t op) {
    if (!g_nccl.comm) {
        if (g_nccl.nranks <= 1) return BXG_OK;   // single rank: identity
        return set_error(BXG_ERR_STATE, "bxg_comm_init not called");
    }
    size_t bytes = n * elt;
    if (bytes > g_nccl.d_cap) {
        if (g_nccl.d_buf) BXG_CUDA(cudaFree(g_nccl.d_buf));
        BXG_CUDA(cudaMalloc(&g_nccl.d_buf, bytes + 256));
        g_nccl.d_cap = bytes + 256;
    }
    cudaStream_t s = ctx().stream;
    BXG_CUDA(cudaMemcpyAsync(g_nccl.d_buf, host, bytes, cudaMemcpyHostToDevice, s));
    BXG_NCCL(g_nccl.AllReduce(g_nccl.d_buf, g_nccl.d_buf, n, dt, op, g_nccl.comm, s));
    BXG_CUDA(cudaMemcpyAsync(host, g_nccl.d_buf, bytes, cudaMemcpyDeviceToHost, s));
    BXG_CUDA(cudaStreamSynchronize(s));
    return BXG_OK;
}
}  // namespace

int bxg_comm_unique_id(char id[BXG_UNIQUE_ID_BYTES]) {
    BXG_TRY(nccl_load());
    static_assert(sizeof(ncclUniqueId) == BXG_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    BXG_NCCL(g_nccl.GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return BXG_OK;
}

int bxg_comm_init(const char id[BXG_UNIQUE_ID_BYTES], int nranks, int rank) {
    BXG_TRY(ensure_init());
    g_nccl.nranks = nranks;
    g_nccl.rank = rank;
    if (nranks <= 1) return BXG_OK;
    BXG_TRY(nccl_load());
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    BXG_NCCL(g_nccl.CommInitRank(&g_nccl.comm, nranks, u, rank));
    return BXG_OK;
}

int bxg_comm_allreduce_i64(int64_t *buf, int64_t n) { return allreduce(buf, (size_t)n, 8, ncclInt64, ncclSum); }
int bxg_comm_allreduce_max_f64(double *buf, int64_t n) { return allreduce(buf, (size_t)n, 8, ncclFloat64, ncclMax); }
// in place on a DEVICE buffer, enqueued on the library stream, no host synchronisation: the reduction of the
// per-chromosome counters stays inside the device-timed step (single rank: identity)
int bxg_comm_allreduce_i64_dev(int64_t *dbuf, int64_t n) {
    if (!g_nccl.comm) {
        if (g_nccl.nranks <= 1) return BXG_OK;
        return set_error(BXG_ERR_STATE, "bxg_comm_init not called");
    }
    if (!dbuf || n <= 0) return set_error(BXG_ERR_ARG, "bad buffer");
    BXG_NCCL(g_nccl.AllReduce(dbuf, dbuf, (size_t)n, ncclInt64, ncclSum, g_nccl.comm, ctx().stream));
    return BXG_OK;
}
int bxg_comm_barrier(void) {
    int64_t x = 1;
    return bxg_comm_allreduce_i64(&x, 1);
}
int bxg_comm_destroy(void) {
    if (g_nccl.comm) {
        g_nccl.CommDestroy(g_nccl.comm);
        g_nccl.comm = nullptr;
    }
    if (g_nccl.d_buf) {
        cudaFree(g_nccl.d_buf);
        g_nccl.d_buf = nullptr;
        g_nccl.d_cap = 0;
    }
    g_nccl.nranks = 0;
    return BXG_OK;
}

}  // extern "C"
