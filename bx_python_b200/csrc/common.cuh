// common.cuh -- shared runtime pieces of libbxb200.so (context, error channel, scratch, launch accounting).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/bxb200.h"

namespace bxg {

struct Context {
    int device = -1;
    int sm_count = 0;
    int64_t l2_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream_override = nullptr;   // when set, BXG_LAUNCH issues on this stream (overlapped find pipeline)
    int cta_cap = 0;                          // when > 0, grid_for() allows at most this many CTAs per SM
    int64_t launches = 0;
    // grow-only device scratch (CUB temp storage, staged host arrays, partials)
    void *scratch[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t scratch_cap[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // small pinned host mailbox for scalar results
    int64_t *mailbox = nullptr;     // 64 x int64, pinned
    int64_t *d_mailbox = nullptr;   // 64 x int64, device
    // 4 KB of mapped pinned memory: tiny host arrays of scalar-style calls are read / written by the kernel in place
    // (no staging copies) -- see zc_small()
    char *zc = nullptr, *zc_dev = nullptr;
    long long zc_seq = 0;
    void *l2_flush_buf = nullptr;
    size_t l2_flush_bytes = 0;
    void *bucket_buf = nullptr;     // second record buffer of the bucketed set_ranges_multi (grow-only)
    size_t bucket_cap = 0;
};

Context &ctx();
int set_error(int code, const char *fmt, ...);
int ensure_init();
// returns device scratch slot `slot` with at least `bytes` capacity (contents not preserved on growth)
int scratch(int slot, size_t bytes, void **out);

#define BXG_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return bxg::set_error(BXG_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                                  __FILE__, __LINE__);                                              \
    } while (0)

#define BXG_TRY(call)             \
    do {                          \
        int r__ = (call);         \
        if (r__ != BXG_OK) return r__; \
    } while (0)

// optional per-kernel CUDA-event timing (bxg_profile_enable / bxg_profile_report); no-ops when disabled
void prof_begin(const char *name);
void prof_end();
bool prof_enabled();
static inline cudaStream_t launch_stream() { return ctx().stream_override ? ctx().stream_override : ctx().stream; }

// every kernel launch of the library goes through this so bench.py can report gpu_launches
#define BXG_LAUNCH(kernel, grid, block, smem, ...)                                   \
    do {                                                                             \
        bxg::prof_begin(#kernel);                                                    \
        kernel<<<(grid), (block), (smem), bxg::launch_stream()>>>(__VA_ARGS__);      \
        bxg::prof_end();                                                             \
        bxg::ctx().launches++;                                                       \
        BXG_CUDA(cudaGetLastError());                                                \
    } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid size for a grid-stride kernel: enough CTAs to cover `work_items / per_cta`, capped at `waves` full waves
static inline int grid_for(int64_t ctas_needed, int ctas_per_sm) {
    if (ctx().cta_cap > 0 && ctas_per_sm > ctx().cta_cap) ctas_per_sm = ctx().cta_cap;
    int64_t cap = (int64_t)ctx().sm_count * ctas_per_sm;
    if (ctas_needed < 1) ctas_needed = 1;
    return (int)(ctas_needed < cap ? ctas_needed : cap);
}

// Scalar-style calls (a handful of positions in host memory, answer needed at once) skip the staging copies: inputs
// are memcpy'd into the mapped pinned page, the kernel reads and writes it over PCIe, one synchronise ends the call.
constexpr int64_t ZC_MAX_ITEMS = 64;          // per array; the page holds 4 arrays of 64 x 8 bytes + slack
static inline bool zc_small(int loc, int64_t n) { return loc == BXG_HOST && n <= ZC_MAX_ITEMS && ctx().zc != nullptr; }
static inline void *zc_host(int k) { return ctx().zc + (size_t)k * 512; }
static inline void *zc_device(int k) { return ctx().zc_dev + (size_t)k * 512; }
// completion word of the scalar-style launches (last 8 bytes of the page): the kernel stores the call's sequence number
// there after its results (system-scope fence), the host spins on it -- a few hundred nanoseconds after the kernel's
// last write instead of the several microseconds a cudaStreamSynchronize adds
constexpr size_t ZC_FLAG_OFFSET = 4096 - 8;
static inline volatile long long *zc_flag_device() { return (volatile long long *)(ctx().zc_dev + ZC_FLAG_OFFSET); }
static inline long long zc_next_seq() { return ++ctx().zc_seq; }
int zc_wait(long long seq);

// Stage a caller array onto the device if it lives on the host; returns the device pointer to use.
int stage_in(int slot, const void *src, size_t bytes, int loc, const void **dptr);

}  // namespace bxg
