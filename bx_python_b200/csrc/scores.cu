// scores.cu -- score sources for the aggregate path.
//
// * BinnedArray (lib/bx/binned_array.py:72-136) as one dense float32 device array: `bins[b] is None` reads as
//   `default`, so a dense array pre-filled with `default` is observably the same thing; the bin geometry stays on the
//   host shim (it only shows through get_bin_offset / nbins).
// * Wiggle load (scripts/aggregate_scores_in_intervals.py:60-70 over lib/bx/wiggle.py:16-85): the reference assigns
//   `scores[chrom][pos] = val` one base at a time in file order, so when spans overlap the LAST one in the file wins.
//   k_spans_write stores every span directly and checks in the same pass whether the batch is sorted and disjoint
//   (every real wiggle file is) -- then each cell was written once and the job is done.  Otherwise an `owner` array
//   over the batch's bounding range takes atomicMax(span index + 1) per base and a second sweep copies the winner's
//   value over the speculative stores: same result as the sequential loop, no dependence on thread order.
// * bigWig summaries (lib/bx/bbi/bbi_file.pyx:66-111, SummarizedData.accumulate_interval_value): one warp per
//   summary bin weighs the intervals that reach it 32 at a time and folds them into the bin in file order with the
//   reference's exact float64 expression sequence (no fma contraction), so valid_count / sum / sum_squares / min /
//   max are bit-identical.
#include <math.h>

#include "scores.cuh"

using namespace bxg;

// ---- dense track ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scores_fill(float *__restrict__ v, int64_t a, int64_t b, float fill) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = a + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < b; i += stride) v[i] = fill;
}

// flags[0] |= 1 when the batch is not (sorted by start, pairwise disjoint); flags[1] = min start, flags[2] = max end
// over the non-empty spans (as offsets from `origin`, biased so that 0 means "no span")
__global__ void __launch_bounds__(256)
k_spans_check(const int32_t *__restrict__ start, const int32_t *__restrict__ end, int64_t n, int64_t origin, int64_t len,
              int no_empty, unsigned long long *__restrict__ flags) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int bad = 0, oob = 0;
    long long lo = INT64_MAX, hi = INT64_MIN;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t s = __ldg(start + i), e = end ? (int64_t)__ldg(end + i) : s + 1;
        if (e <= s) {                                  // range(start, end) is empty: nothing is assigned
            if (no_empty) bad = 1;                     // ... but a caller that bisects on `end` cannot have it
            continue;
        }
        if (s - origin < 0 || e - origin > len) oob = 1;
        lo = s < lo ? s : lo;
        hi = e > hi ? e : hi;
        // the previous NON-EMPTY span is what matters; an empty one in between hides it, so be conservative
        if (i > 0) {
            const int64_t ps = __ldg(start + i - 1), pe = end ? (int64_t)__ldg(end + i - 1) : ps + 1;
            if (pe <= ps || s < pe) bad = 1;
        }
    }
    bad = __any_sync(0xffffffffu, bad);
    oob = __any_sync(0xffffffffu, oob);
    for (int o = 16; o; o >>= 1) {
        const long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        if (bad | oob) atomicOr(flags, (unsigned long long)(bad | (oob << 1)));
        if (lo != INT64_MAX) {
            atomicMin((long long *)flags + 1, lo);
            atomicMax((long long *)flags + 2, hi);
        }
    }
}

// The common case in one pass: write every span as if the batch were sorted and disjoint (then each cell is written
// exactly once and order is moot) WHILE checking that it is -- same flags as k_spans_check.  If the check fails the
// host runs the owner / apply passes afterwards: they rewrite exactly the cells touched here with the correct winner,
// so the speculative stores do no harm.  Spans outside the track are flagged and not written.
// One lane per span for the first 8 bases; longer spans are finished by the whole warp with coalesced stores.
__global__ void __launch_bounds__(256)
k_spans_write(float *__restrict__ v, int64_t origin, int64_t len, const int32_t *__restrict__ start,
              const int32_t *__restrict__ end, const float *__restrict__ val, int64_t n,
              unsigned long long *__restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int bad = 0, oob = 0;
    long long lo = INT64_MAX, hi = INT64_MIN;
    for (int64_t base = ((((int64_t)blockIdx.x * blockDim.x) + threadIdx.x) >> 5) << 5; base < n; base += nwarps << 5) {
        const int64_t i = base + lane;
        int64_t a = 0, b = 0;
        float x = 0.0f;
        if (i < n) {
            const int64_t s = __ldg(start + i), e = end ? (int64_t)__ldg(end + i) : s + 1;
            if (e > s) {
                if (s - origin < 0 || e - origin > len) {
                    oob = 1;
                } else {
                    a = s - origin;
                    b = e - origin;
                    x = __ldg(val + i);
                    lo = s < lo ? s : lo;
                    hi = e > hi ? e : hi;
                }
                if (i > 0) {                       // previous span empty or reaching into this one: not the easy case
                    const int64_t ps = __ldg(start + i - 1), pe = end ? (int64_t)__ldg(end + i - 1) : ps + 1;
                    if (pe <= ps || s < pe) bad = 1;
                }
            }
        }
        const int64_t head = b - a < 8 ? b : a + 8;
        for (int64_t p = a; p < head; p++) v[p] = x;
        unsigned longm = __ballot_sync(0xffffffffu, b - a > 8);
        while (longm) {
            const int src = __ffs(longm) - 1;
            longm &= longm - 1;
            const int64_t la = __shfl_sync(0xffffffffu, a, src) + 8, lb = __shfl_sync(0xffffffffu, b, src);
            const float lx = __shfl_sync(0xffffffffu, x, src);
            for (int64_t p = la + lane; p < lb; p += 32) v[p] = lx;
        }
    }
    bad = __any_sync(0xffffffffu, bad);
    oob = __any_sync(0xffffffffu, oob);
    for (int o = 16; o; o >>= 1) {
        const long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if (lane == 0) {
        if (bad | oob) atomicOr(flags, (unsigned long long)(bad | (oob << 1)));
        if (lo != INT64_MAX) {
            atomicMin((long long *)flags + 1, lo);
            atomicMax((long long *)flags + 2, hi);
        }
    }
}

// general batch, pass 1: owner[p - lo] = max(span index + 1) over the spans covering p
__global__ void __launch_bounds__(256)
k_spans_owner(uint32_t *__restrict__ owner, int64_t lo, const int32_t *__restrict__ start,
              const int32_t *__restrict__ end, int64_t n) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = ((((int64_t)blockIdx.x * blockDim.x) + threadIdx.x) >> 5) << 5; base < n; base += nwarps << 5) {
        const int64_t i = base + lane;
        int64_t a = 0, b = 0;
        if (i < n) {
            a = (int64_t)__ldg(start + i) - lo;
            b = end ? (int64_t)__ldg(end + i) - lo : a + 1;
        }
        const int64_t head = b - a < 8 ? b : a + 8;
        for (int64_t p = a; p < head; p++) atomicMax(owner + p, (uint32_t)(i + 1));
        unsigned longm = __ballot_sync(0xffffffffu, b - a > 8);
        while (longm) {
            const int src = __ffs(longm) - 1;
            longm &= longm - 1;
            const int64_t la = __shfl_sync(0xffffffffu, a, src) + 8, lb = __shfl_sync(0xffffffffu, b, src);
            const uint32_t id = (uint32_t)(base + src + 1);
            for (int64_t p = la + lane; p < lb; p += 32) atomicMax(owner + p, id);
        }
    }
}

// pass 2: the winner's value lands in the track
__global__ void __launch_bounds__(256)
k_spans_apply(float *__restrict__ v, const uint32_t *__restrict__ owner, int64_t m, const float *__restrict__ val) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < m; p += stride) {
        const uint32_t o = owner[p];
        if (o) v[p] = __ldg(val + (o - 1));
    }
}

__global__ void __launch_bounds__(256)
k_scores_gather(const float *__restrict__ v, int64_t n, int64_t origin, float fill, const int32_t *__restrict__ pos,
                int64_t np, float *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += stride) {
        const int64_t p = (int64_t)__ldg(pos + i) - origin;
        out[i] = (p >= 0 && p < n) ? __ldg(v + p) : fill;
    }
}

// ---- bigWig summary ------------------------------------------------------------------------------------------------
// SummarizedData.accumulate_interval_value (bbi_file.pyx:80-111), bin j of `size` over [rs, re), for every interval
// in file order.  All arithmetic in the reference's order and types: `overlap` int, overlap_factor = overlap /
// interval_size (double), interval_weight = interval_size * overlap_factor, val a C float (so val * val is rounded to
// float before it is widened).  SORTED: the batch is sorted by start and disjoint (what a bigWig file holds), so
// the intervals reaching a bin are one contiguous run found by binary search; otherwise every bin scans the batch.
//
// One WARP per bin: the 32 lanes load and weigh 32 consecutive intervals at a time (coalesced loads, the divisions and
// products in parallel -- they do not depend on the running sums), park the three terms in shared memory, and lane 0
// then folds them into the bin's state strictly in file order (the float64 sums and the `<` / `>` updates of min / max
// are order-dependent, down to the sign of a zero).  The next batch's loads are issued before the fold.
constexpr int SUMM_WARPS = 4;
template <bool SORTED>
__global__ void __launch_bounds__(32 * SUMM_WARPS)
k_summarize(const int32_t *__restrict__ start, const int32_t *__restrict__ end, const float *__restrict__ val, int64_t n,
            int64_t rs, int64_t re, int32_t size, double *__restrict__ valid, double *__restrict__ mn,
            double *__restrict__ mx, double *__restrict__ sum, double *__restrict__ sq) {
    __shared__ double s_w[SUMM_WARPS][32], s_t1[SUMM_WARPS][32], s_t2[SUMM_WARPS][32], s_v[SUMM_WARPS][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int j = blockIdx.x * SUMM_WARPS + wib;
    if (j >= size) return;                          // whole warp leaves together
    const int64_t step = (re - rs) / size;
    const int64_t b0 = rs + step * j, b1 = b0 + step;
    double vc = 0.0, lo = 0.0, hi = 0.0, sm = 0.0, ss = 0.0;
    if (lane == 0) { vc = valid[j]; lo = mn[j]; hi = mx[j]; sm = sum[j]; ss = sq[j]; }   // accumulate INTO the caller's state
    int64_t i0 = 0;
    if (SORTED) {                                   // first interval with end > b0 (ends ascend too)
        int64_t a = 0, b = n;
        while (a < b) {
            const int64_t m = (a + b) >> 1;
            if ((int64_t)(uint32_t)__ldg(end + m) > b0) b = m; else a = m + 1;
        }
        i0 = a & ~31ll;                             // aligned batches: 128-byte coalesced loads
    }
    int64_t ns = 0, ne = 0;
    float nv = 0.0f;
    if (i0 + lane < n) { ns = (uint32_t)__ldg(start + i0 + lane); ne = (uint32_t)__ldg(end + i0 + lane); nv = __ldg(val + i0 + lane); }
    for (int64_t base = i0; base < n; base += 32) {
        int64_t s = ns, e = ne;                     // bits32 in the reference
        const float vf = nv;
        const bool live = base + lane < n;
        if (SORTED && __shfl_sync(0xffffffffu, s, 0) >= b1) break;      // every later interval starts past the bin
        if (base + 32 + lane < n) {                 // prefetch the next batch
            ns = (uint32_t)__ldg(start + base + 32 + lane); ne = (uint32_t)__ldg(end + base + 32 + lane); nv = __ldg(val + base + 32 + lane);
        }
        bool has = false;
        if (live) {
            if (s < rs) s = rs;
            if (e > re) e = re;
            if (s < e) {
                const int64_t ov = (e < b1 ? e : b1) - (s > b0 ? s : b0);
                if (ov > 0) {
                    has = true;
                    const double v = (double)vf;
                    const double isz = (double)(int32_t)(e - s);
                    const double w = __dmul_rn(isz, __ddiv_rn((double)(int32_t)ov, isz));
                    s_w[wib][lane] = w;
                    s_t1[wib][lane] = __dmul_rn(v, w);
                    s_t2[wib][lane] = __dmul_rn((double)__fmul_rn(vf, vf), w);
                    s_v[wib][lane] = v;
                }
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, has);
        __syncwarp();
        if (lane == 0) {
            while (m) {
                const int k = __ffs(m) - 1;
                m &= m - 1;
                const double v = s_v[wib][k];
                vc = __dadd_rn(vc, s_w[wib][k]);
                sm = __dadd_rn(sm, s_t1[wib][k]);
                ss = __dadd_rn(ss, s_t2[wib][k]);
                if (hi < v) hi = v;
                if (lo > v) lo = v;
            }
        }
        __syncwarp();
    }
    if (lane == 0) { valid[j] = vc; mn[j] = lo; mx[j] = hi; sum[j] = sm; sq[j] = ss; }
}

static int fill_range(bxg_scores *s, int64_t a, int64_t b) {
    if (b <= a) return BXG_OK;
    BXG_LAUNCH(k_scores_fill, grid_for(cdiv(b - a, 256 * 8), 8), 256, 0, s->v, a, b, s->fill);
    return BXG_OK;
}

extern "C" {

int bxg_scores_alloc(int64_t n, int32_t origin, float fill, bxg_scores_t **out) {
    BXG_TRY(ensure_init());
    if (!out || n < 0) return set_error(BXG_ERR_ARG, "bad arguments");
    bxg_scores *s = new bxg_scores();
    s->n = n;
    s->cap = n > 0 ? n : 1;
    s->origin = origin;
    s->fill = fill;
    if (cudaMalloc(&s->v, (size_t)s->cap * 4) != cudaSuccess) {
        delete s;
        cudaGetLastError();
        return set_error(BXG_ERR_CUDA, "cudaMalloc of %lld score cells failed", (long long)n);
    }
    int r = fill_range(s, 0, n);
    if (r != BXG_OK) {
        cudaFree(s->v);
        delete s;
        return r;
    }
    *out = s;
    return BXG_OK;
}

int bxg_scores_info(const bxg_scores_t *s, int64_t *n, int32_t *origin, float *fill) {
    if (!s) return set_error(BXG_ERR_ARG, "null scores handle");
    if (n) *n = s->n;
    if (origin) *origin = s->origin;
    if (fill) *fill = s->fill;
    return BXG_OK;
}

int bxg_scores_reserve(bxg_scores_t *s, int64_t n) {
    if (!s) return set_error(BXG_ERR_ARG, "null scores handle");
    if (n <= s->n) return BXG_OK;
    Context &c = ctx();
    if (n > s->cap) {
        int64_t cap = s->cap * 2 > n ? s->cap * 2 : n;
        float *nv = nullptr;
        if (cudaMalloc(&nv, (size_t)cap * 4) != cudaSuccess) {          // doubling may not fit: take the exact size
            cudaGetLastError();
            cap = n;
            BXG_CUDA(cudaMalloc(&nv, (size_t)cap * 4));
        }
        if (s->n) BXG_CUDA(cudaMemcpyAsync(nv, s->v, (size_t)s->n * 4, cudaMemcpyDeviceToDevice, c.stream));
        BXG_CUDA(cudaStreamSynchronize(c.stream));
        BXG_CUDA(cudaFree(s->v));
        s->v = nv;
        s->cap = cap;
    }
    const int64_t old = s->n;
    s->n = n;
    return fill_range(s, old, n);
}

int bxg_scores_set_spans(bxg_scores_t *s, const int32_t *start, const int32_t *end, const float *val, int64_t n,
                         int loc) {
    if (!s) return set_error(BXG_ERR_ARG, "null scores handle");
    if (n <= 0) return BXG_OK;
    if (n >= 0xffffffffll) return set_error(BXG_ERR_ARG, "at most 2^32-2 spans per batch");
    Context &c = ctx();
    const void *ds, *de = nullptr, *dv;
    BXG_TRY(stage_in(0, start, (size_t)n * 4, loc, &ds));
    if (end) BXG_TRY(stage_in(1, end, (size_t)n * 4, loc, &de));
    BXG_TRY(stage_in(5, val, (size_t)n * 4, loc, &dv));
    unsigned long long *flags = (unsigned long long *)(c.d_mailbox + 16);
    c.mailbox[16] = 0;
    c.mailbox[17] = INT64_MAX;
    c.mailbox[18] = INT64_MIN;
    BXG_CUDA(cudaMemcpyAsync(flags, c.mailbox + 16, 24, cudaMemcpyHostToDevice, c.stream));
    const int g = grid_for(cdiv(n, 256), 8);
    BXG_LAUNCH(k_spans_write, g, 256, 0, s->v, (int64_t)s->origin, s->n, (const int32_t *)ds, (const int32_t *)de,
               (const float *)dv, n, flags);
    BXG_CUDA(cudaMemcpyAsync(c.mailbox + 16, flags, 24, cudaMemcpyDeviceToHost, c.stream));
    BXG_CUDA(cudaStreamSynchronize(c.stream));
    const int64_t fl = c.mailbox[16], lo = c.mailbox[17], hi = c.mailbox[18];
    if (fl & 2) return set_error(BXG_ERR_ARG, "span outside the track [%lld, %lld) (spans inside it were written)",
                                 (long long)s->origin, (long long)(s->origin + s->n));
    if (lo == INT64_MAX || !(fl & 1)) return BXG_OK;    // nothing to write, or sorted and disjoint: done in one pass
    const int64_t m = hi - lo;
    void *owner;
    BXG_TRY(scratch(2, (size_t)m * 4, &owner));
    BXG_CUDA(cudaMemsetAsync(owner, 0, (size_t)m * 4, c.stream));
    BXG_LAUNCH(k_spans_owner, g, 256, 0, (uint32_t *)owner, lo, (const int32_t *)ds, (const int32_t *)de, n);
    BXG_LAUNCH(k_spans_apply, grid_for(cdiv(m, 256 * 4), 8), 256, 0, s->v + (lo - s->origin), (const uint32_t *)owner, m,
               (const float *)dv);
    return BXG_OK;
}

int bxg_scores_write(bxg_scores_t *s, int64_t start, const float *vals, int64_t n, int loc) {
    if (!s || n < 0) return set_error(BXG_ERR_ARG, "bad arguments");
    if (n == 0) return BXG_OK;
    const int64_t a = start - s->origin;
    if (a < 0 || a + n > s->n)
        return set_error(BXG_ERR_ARG, "cells [%lld, %lld) outside the track [%lld, %lld)", (long long)start,
                         (long long)(start + n), (long long)s->origin, (long long)(s->origin + s->n));
    Context &c = ctx();
    BXG_CUDA(cudaMemcpyAsync(s->v + a, vals, (size_t)n * 4,
                             loc == BXG_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, c.stream));
    if (loc == BXG_HOST) BXG_CUDA(cudaStreamSynchronize(c.stream));      // the caller may reuse `vals` right away
    return BXG_OK;
}

int bxg_scores_get(const bxg_scores_t *s, const int32_t *pos, int64_t n, float *out, int loc) {
    if (!s) return set_error(BXG_ERR_ARG, "null scores handle");
    if (n <= 0) return BXG_OK;
    Context &c = ctx();
    const void *dp;
    BXG_TRY(stage_in(0, pos, (size_t)n * 4, loc, &dp));
    float *dout = out;
    if (loc == BXG_HOST) {
        void *o;
        BXG_TRY(scratch(2, (size_t)n * 4, &o));
        dout = (float *)o;
    }
    BXG_LAUNCH(k_scores_gather, grid_for(cdiv(n, 256), 8), 256, 0, s->v, s->n, (int64_t)s->origin, s->fill,
               (const int32_t *)dp, n, dout);
    if (loc == BXG_HOST) {
        BXG_CUDA(cudaMemcpyAsync(out, dout, (size_t)n * 4, cudaMemcpyDeviceToHost, c.stream));
        BXG_CUDA(cudaStreamSynchronize(c.stream));
    }
    return BXG_OK;
}

int bxg_scores_get_range(const bxg_scores_t *s, int64_t start, int64_t end, float *out) {
    if (!s || end < start) return set_error(BXG_ERR_ARG, "bad arguments");
    Context &c = ctx();
    const int64_t a = start - s->origin, b = end - s->origin;
    const int64_t ca = a < 0 ? 0 : (a > s->n ? s->n : a), cb = b < 0 ? 0 : (b > s->n ? s->n : b);
    for (int64_t i = a; i < ca && i < b; i++) out[i - a] = s->fill;      // before the track
    if (cb > ca) BXG_CUDA(cudaMemcpyAsync(out + (ca - a), s->v + ca, (size_t)(cb - ca) * 4, cudaMemcpyDeviceToHost, c.stream));
    for (int64_t i = (cb > a ? cb : a); i < b; i++) out[i - a] = s->fill;   // past the track
    BXG_CUDA(cudaStreamSynchronize(c.stream));
    return BXG_OK;
}

int bxg_scores_device(const bxg_scores_t *s, const float **dptr, int64_t *n) {
    if (!s) return set_error(BXG_ERR_ARG, "null scores handle");
    if (dptr) *dptr = s->v;
    if (n) *n = s->n;
    return BXG_OK;
}

int bxg_summarize(const int32_t *start, const int32_t *end, const float *val, int64_t n, int loc, uint32_t rstart,
                  uint32_t rend, int32_t size, double *valid_count, double *min_val, double *max_val, double *sum_data,
                  double *sum_squares) {
    BXG_TRY(ensure_init());
    if (size <= 0 || rend <= rstart) return set_error(BXG_ERR_ARG, "summarize needs size > 0 and start < end");
    if (n < 0) return set_error(BXG_ERR_ARG, "bad arguments");
    Context &c = ctx();
    const void *ds = nullptr, *de = nullptr, *dv = nullptr;
    bool sorted = true;
    if (n) {
        BXG_TRY(stage_in(0, start, (size_t)n * 4, loc, &ds));
        BXG_TRY(stage_in(1, end, (size_t)n * 4, loc, &de));
        BXG_TRY(stage_in(5, val, (size_t)n * 4, loc, &dv));
        unsigned long long *flags = (unsigned long long *)(c.d_mailbox + 16);
        c.mailbox[16] = 0;
        c.mailbox[17] = INT64_MAX;
        c.mailbox[18] = INT64_MIN;
        BXG_CUDA(cudaMemcpyAsync(flags, c.mailbox + 16, 24, cudaMemcpyHostToDevice, c.stream));
        BXG_LAUNCH(k_spans_check, grid_for(cdiv(n, 256), 8), 256, 0, (const int32_t *)ds, (const int32_t *)de, n,
                   (int64_t)INT32_MIN, (int64_t)1 << 33, 1, flags);
        BXG_CUDA(cudaMemcpyAsync(c.mailbox + 16, flags, 8, cudaMemcpyDeviceToHost, c.stream));
        BXG_CUDA(cudaStreamSynchronize(c.stream));
        sorted = (c.mailbox[16] & 1) == 0;
    }
    double *d = valid_count;
    double *dmn = min_val, *dmx = max_val, *dsm = sum_data, *dsq = sum_squares;
    if (loc == BXG_HOST) {
        void *o;
        BXG_TRY(scratch(2, (size_t)size * 40, &o));
        d = (double *)o;
        dmn = d + size; dmx = dmn + size; dsm = dmx + size; dsq = dsm + size;
        cudaStream_t st = c.stream;
        BXG_CUDA(cudaMemcpyAsync(d, valid_count, (size_t)size * 8, cudaMemcpyHostToDevice, st));
        BXG_CUDA(cudaMemcpyAsync(dmn, min_val, (size_t)size * 8, cudaMemcpyHostToDevice, st));
        BXG_CUDA(cudaMemcpyAsync(dmx, max_val, (size_t)size * 8, cudaMemcpyHostToDevice, st));
        BXG_CUDA(cudaMemcpyAsync(dsm, sum_data, (size_t)size * 8, cudaMemcpyHostToDevice, st));
        BXG_CUDA(cudaMemcpyAsync(dsq, sum_squares, (size_t)size * 8, cudaMemcpyHostToDevice, st));
    }
    if (sorted)
        BXG_LAUNCH(k_summarize<true>, (int)cdiv(size, SUMM_WARPS), 32 * SUMM_WARPS, 0, (const int32_t *)ds, (const int32_t *)de,
                   (const float *)dv, n, (int64_t)rstart, (int64_t)rend, size, d, dmn, dmx, dsm, dsq);
    else
        BXG_LAUNCH(k_summarize<false>, (int)cdiv(size, SUMM_WARPS), 32 * SUMM_WARPS, 0, (const int32_t *)ds, (const int32_t *)de,
                   (const float *)dv, n, (int64_t)rstart, (int64_t)rend, size, d, dmn, dmx, dsm, dsq);
    if (loc == BXG_HOST) {
        cudaStream_t st = c.stream;
        BXG_CUDA(cudaMemcpyAsync(valid_count, d, (size_t)size * 8, cudaMemcpyDeviceToHost, st));
        BXG_CUDA(cudaMemcpyAsync(min_val, dmn, (size_t)size * 8, cudaMemcpyDeviceToHost, st));
        BXG_CUDA(cudaMemcpyAsync(max_val, dmx, (size_t)size * 8, cudaMemcpyDeviceToHost, st));
        BXG_CUDA(cudaMemcpyAsync(sum_data, dsm, (size_t)size * 8, cudaMemcpyDeviceToHost, st));
        BXG_CUDA(cudaMemcpyAsync(sum_squares, dsq, (size_t)size * 8, cudaMemcpyDeviceToHost, st));
        BXG_CUDA(cudaStreamSynchronize(st));
    }
    return BXG_OK;
}

}  // extern "C"
