// aggregate.cu -- the per-base inner loop of scripts/aggregate_scores_in_intervals.py:107-134 as one kernel.
//
// One thread per window walks its positions left to right (the float32 sum is order-dependent and must match the
// reference's sequential `total += score`), skipping score == 0.0 (:115 truthiness), masked bases (:117-119) and NaN
// (:122).  Scores are a dense float32 array (NaN = unset, lib/bx/binned_array.py:73,89-94); the mask is a dense
// LSB-first bitmap (bits.cu).
#include <math.h>

#include "scores.cuh"

using namespace bxg;

struct bxg_bits;
const uint64_t *bxg_bits_words_internal(const bxg_bits *b);
int32_t bxg_bits_size_internal(const bxg_bits *b);

// one aligned 32-byte sector of scores (8 floats) in a single LDG.256
__device__ __forceinline__ void ld_scores8(const float *p, float (&x)[8]) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3]), "=f"(x[4]), "=f"(x[5]), "=f"(x[6]), "=f"(x[7]) : "l"(p));
}

// one window: strict left-to-right float32 accumulation over positions [ws,we) of one track.
//
// The arithmetic is sequential by definition (the reference's `total += score`), but the LOADS need not be: the lanes of a
// warp work on 32 unrelated windows, so a 4-byte load per base is 32 L1 wavefronts per instruction for 32 useful floats --
// the kernel was bound by exactly that (0.59 ms per 5 M windows = one wavefront per base).  Each lane now fetches its
// strip sector by sector (one 256-bit load = 8 scores per wavefront) and walks the registers in order; the mask bitmap is
// read one 64-bit word per 64 positions.
__device__ __forceinline__ void aggregate_window(const float *__restrict__ v, int64_t n, int64_t origin, float fill,
                                                 const uint64_t *__restrict__ mask, int64_t mask_size, int64_t ws, int64_t we,
                                                 float &sum, float &avg, int32_t &cnt, float &mn, float &mx) {
    int64_t a = ws - origin, b = we - origin;
    // positions outside the track read as the array's default (BinnedArray.get, binned_array.py:89-94): NaN or 0.0 --
    // both skipped by the script -- for every array the script itself builds, so the window is clipped to the track;
    // any other default (a FileBinnedArray written with one) counts, and the window is walked in full
    const bool fill_counts = !(fill == 0.0f || fill != fill);
    if (!fill_counts) {
        if (a < 0) a = 0;
        if (b > n) b = n;
    }
    float total = 0.0f, lo = 100000000.0f, hi = -100000000.0f;   // script sentinels (:112-113), exact in float32
    int32_t c = 0;
    int64_t mw = -1;                                             // mask word currently held
    unsigned long long mbits = 0;
    auto take = [&](float s, int64_t i) {
        if (s == 0.0f || s != s) return;
        if (mask) {
            const int64_t p = i + origin;
            if (p >= 0 && p < mask_size) {
                if ((p >> 6) != mw) {
                    mw = p >> 6;
                    mbits = __ldg((const unsigned long long *)mask + mw);
                }
                if ((mbits >> (p & 63)) & 1ull) return;
            }
        }
        total = __fadd_rn(total, s);      // strict left-to-right float32 (no fma contraction possible, but be explicit)
        c++;
        hi = s > hi ? s : hi;
        lo = s < lo ? s : lo;
    };
    int64_t i = a;
    while (i < b) {
        const int64_t base = i & ~(int64_t)7;
        if (base >= 0 && base + 8 <= n) {                        // a whole sector inside the track: one 256-bit load
            float x[8];
            ld_scores8(v + base, x);
            const int k0 = (int)(i - base), k1 = b - base < 8 ? (int)(b - base) : 8;
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (k >= k0 && k < k1) take(x[k], base + k);
            i = base + 8;
        } else {                                                 // track edge (or a counting default outside it): per base
            const int64_t stop = (base + 8 < b) ? base + 8 : b;
            for (; i < stop; i++) take((i >= 0 && i < n) ? __ldg(v + i) : fill, i);
        }
    }
    cnt = c;
    sum = total;
    if (c > 0) {
        avg = __fdiv_rn(total, (float)c);
        mn = lo;
        mx = hi;
    } else {
        const float qnan = __int_as_float(0x7fc00000);
        avg = mn = mx = qnan;
    }
}

__global__ void __launch_bounds__(256)
k_aggregate(const float *__restrict__ v, int64_t n, int64_t origin, float fill, const uint64_t *__restrict__ mask, int64_t mask_size,
            const int32_t *__restrict__ ws, const int32_t *__restrict__ we, int64_t nw,
            float *__restrict__ sum, float *__restrict__ avg, int32_t *__restrict__ cnt, float *__restrict__ mn,
            float *__restrict__ mx) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nw; w += stride)
        aggregate_window(v, n, origin, fill, mask, mask_size, __ldg(ws + w), __ldg(we + w), sum[w], avg[w], cnt[w], mn[w], mx[w]);
}

// genome-wide form: every window names its track (chromosome); one launch for the whole BED file
struct TrackDesc {
    const float *v;
    int64_t n, origin;
    float fill;
    const uint64_t *mask;
    int64_t mask_size;
};

__global__ void __launch_bounds__(256)
k_aggregate_multi(const TrackDesc *__restrict__ tracks, int ntracks, const int32_t *__restrict__ wt,
                  const int32_t *__restrict__ ws, const int32_t *__restrict__ we, int64_t nw,
                  float *__restrict__ sum, float *__restrict__ avg, int32_t *__restrict__ cnt, float *__restrict__ mn,
                  float *__restrict__ mx) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nw; w += stride) {
        const int32_t t = __ldg(wt + w);
        if (t >= 0 && t < ntracks) {
            const TrackDesc d = tracks[t];
            aggregate_window(d.v, d.n, d.origin, d.fill, d.mask, d.mask_size, __ldg(ws + w), __ldg(we + w), sum[w], avg[w], cnt[w],
                             mn[w], mx[w]);
        } else {                          // `chrom not in scores_by_chrom` (:115): nothing counted
            const float qnan = __int_as_float(0x7fc00000);
            sum[w] = 0.0f; cnt[w] = 0; avg[w] = mn[w] = mx[w] = qnan;
        }
    }
}

extern "C" {

int bxg_scores_create(const float *scores, int64_t n, int32_t origin, int loc, bxg_scores_t **out) {
    BXG_TRY(ensure_init());
    if (!out || n < 0) return set_error(BXG_ERR_ARG, "bad arguments");
    bxg_scores *s = new bxg_scores();
    s->n = n;
    s->cap = n > 0 ? n : 1;
    s->origin = origin;
    s->fill = __builtin_nanf("");
    BXG_CUDA(cudaMalloc(&s->v, (size_t)s->cap * 4));
    if (n)
        BXG_CUDA(cudaMemcpyAsync(s->v, scores, (size_t)n * 4,
                                 loc == BXG_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, ctx().stream));
    if (loc == BXG_HOST) BXG_CUDA(cudaStreamSynchronize(ctx().stream));
    *out = s;
    return BXG_OK;
}

int bxg_scores_free(bxg_scores_t *s) {
    if (!s) return BXG_OK;
    cudaStreamSynchronize(ctx().stream);
    cudaFree(s->v);
    delete s;
    return BXG_OK;
}

int bxg_aggregate(const bxg_scores_t *s, const bxg_bits_t *mask, const int32_t *ws, const int32_t *we, int64_t nw,
                  int loc, float *sum, float *avg, int32_t *count, float *mn, float *mx) {
    if (!s) return set_error(BXG_ERR_ARG, "null scores handle");
    if (nw <= 0) return BXG_OK;
    Context &c = ctx();
    const void *dws, *dwe;
    BXG_TRY(stage_in(0, ws, (size_t)nw * 4, loc, &dws));
    BXG_TRY(stage_in(1, we, (size_t)nw * 4, loc, &dwe));
    float *dsum = sum, *davg = avg, *dmn = mn, *dmx = mx;
    int32_t *dcnt = count;
    if (loc == BXG_HOST) {
        void *o;
        BXG_TRY(scratch(2, (size_t)nw * 20, &o));
        dsum = (float *)o;
        davg = dsum + nw;
        dmn = davg + nw;
        dmx = dmn + nw;
        dcnt = (int32_t *)(dmx + nw);
    }
    BXG_LAUNCH(k_aggregate, grid_for(cdiv(nw, 256), 8), 256, 0, s->v, s->n, (int64_t)s->origin, s->fill,
               bxg_bits_words_internal((const bxg_bits *)mask), (int64_t)bxg_bits_size_internal((const bxg_bits *)mask),
               (const int32_t *)dws, (const int32_t *)dwe, nw, dsum, davg, dcnt, dmn, dmx);
    if (loc == BXG_HOST) {
        cudaStream_t st = c.stream;
        BXG_CUDA(cudaMemcpyAsync(sum, dsum, (size_t)nw * 4, cudaMemcpyDeviceToHost, st));
        BXG_CUDA(cudaMemcpyAsync(avg, davg, (size_t)nw * 4, cudaMemcpyDeviceToHost, st));
        BXG_CUDA(cudaMemcpyAsync(mn, dmn, (size_t)nw * 4, cudaMemcpyDeviceToHost, st));
        BXG_CUDA(cudaMemcpyAsync(mx, dmx, (size_t)nw * 4, cudaMemcpyDeviceToHost, st));
        BXG_CUDA(cudaMemcpyAsync(count, dcnt, (size_t)nw * 4, cudaMemcpyDeviceToHost, st));
        BXG_CUDA(cudaStreamSynchronize(st));
    }
    return BXG_OK;
}

int bxg_aggregate_multi(const bxg_scores_t *const *tracks, const bxg_bits_t *const *masks, int32_t ntracks,
                        const int32_t *wtrack, const int32_t *ws, const int32_t *we, int64_t nw, int loc,
                        float *sum, float *avg, int32_t *count, float *mn, float *mx) {
    BXG_TRY(ensure_init());
    if (ntracks <= 0 || ntracks > 4096) return set_error(BXG_ERR_ARG, "ntracks must be in [1, 4096]");
    if (nw <= 0) return BXG_OK;
    Context &c = ctx();
    static TrackDesc h_desc[4096];
    for (int t = 0; t < ntracks; t++) {
        if (!tracks[t]) {                             // `chrom in scores_by_chrom` is false (:115): nothing is counted
            h_desc[t] = TrackDesc{nullptr, 0, 0, __builtin_nanf(""), nullptr, 0};
            continue;
        }
        const bxg_bits *m = masks ? (const bxg_bits *)masks[t] : nullptr;
        h_desc[t] = TrackDesc{tracks[t]->v, tracks[t]->n, (int64_t)tracks[t]->origin, tracks[t]->fill, bxg_bits_words_internal(m),
                              (int64_t)bxg_bits_size_internal(m)};
    }
    void *d_desc;
    BXG_TRY(scratch(3, sizeof(TrackDesc) * (size_t)ntracks, &d_desc));
    BXG_CUDA(cudaMemcpyAsync(d_desc, h_desc, sizeof(TrackDesc) * (size_t)ntracks, cudaMemcpyHostToDevice, c.stream));
    const void *dwt, *dws, *dwe;
    BXG_TRY(stage_in(0, wtrack, (size_t)nw * 4, loc, &dwt));
    BXG_TRY(stage_in(1, ws, (size_t)nw * 4, loc, &dws));
    BXG_TRY(stage_in(5, we, (size_t)nw * 4, loc, &dwe));
    float *dsum = sum, *davg = avg, *dmn = mn, *dmx = mx;
    int32_t *dcnt = count;
    if (loc == BXG_HOST) {
        void *o;
        BXG_TRY(scratch(2, (size_t)nw * 20, &o));
        dsum = (float *)o;
        davg = dsum + nw;
        dmn = davg + nw;
        dmx = dmn + nw;
        dcnt = (int32_t *)(dmx + nw);
    }
    // (8 CTAs per SM with 5 resident: sizing the grid to the resident CTAs measured slower, 0.347 -> 0.365 ms, r02t)
    BXG_LAUNCH(k_aggregate_multi, grid_for(cdiv(nw, 256), 8), 256, 0, (const TrackDesc *)d_desc, ntracks,
               (const int32_t *)dwt, (const int32_t *)dws, (const int32_t *)dwe, nw, dsum, davg, dcnt, dmn, dmx);
    if (loc == BXG_HOST) {
        cudaStream_t st = c.stream;
        BXG_CUDA(cudaMemcpyAsync(sum, dsum, (size_t)nw * 4, cudaMemcpyDeviceToHost, st));
        BXG_CUDA(cudaMemcpyAsync(avg, davg, (size_t)nw * 4, cudaMemcpyDeviceToHost, st));
        BXG_CUDA(cudaMemcpyAsync(mn, dmn, (size_t)nw * 4, cudaMemcpyDeviceToHost, st));
        BXG_CUDA(cudaMemcpyAsync(mx, dmx, (size_t)nw * 4, cudaMemcpyDeviceToHost, st));
        BXG_CUDA(cudaMemcpyAsync(count, dcnt, (size_t)nw * 4, cudaMemcpyDeviceToHost, st));
        BXG_CUDA(cudaStreamSynchronize(st));
    }
    return BXG_OK;
}

}  // extern "C"
