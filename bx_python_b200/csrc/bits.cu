// bits.cu -- BinnedBitSet / BitSet kernels: uint64 bit-parallel set algebra, popcount, rank table, run extraction.
//
// HBM layout (DESIGN.md "bitsets"): dense LSB-first uint64 words over [0,size) -- tail bits of the last word and the
// padding words up to a 32-byte multiple are kept zero -- plus uint8 state[nbins] tracking the reference's lazy-bin
// state machine (src/binBits.c:5-6; transitions :67-128,:230-317).  All word kernels are state-oblivious; the state
// only feeds the strict count_range correction (binBits.c:155,161).
#include <cub/cub.cuh>
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

using namespace bxg;

enum : uint8_t { BZ = 0, BO = 1, BA = 2 };

struct bxg_bits {
    int32_t size = 0, bin_size = 0, nbins = 0, flat = 0;
    int64_t nwords = 0;        // ceil(size/64)
    int64_t nwords_alloc = 0;  // nwords rounded up to a multiple of 4 (32 B)
    uint64_t *words = nullptr;
    uint8_t *state = nullptr;
    uint32_t *rank = nullptr;  // rank lines: 32-byte lines {rank, 7 x 32 bitmap bits}, nlines + 1 of them (lazy; see count_range)
    int64_t nlines = 0;
    bool maybe_one = false;    // an ALL_ONE sentinel bin may exist (only then does strict count_range need the bin states)
    bool rank_valid = false;
    int32_t *run_s = nullptr, *run_e = nullptr;  // run extraction output (device)
    int64_t run_cap = 0, nruns = -1;
    int32_t *rr_s = nullptr, *rr_e = nullptr;    // runs-in-ranges output (device)
    int64_t rr_cap = 0, rr_total = 0;
};

// ------------------------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum -> one atomicAdd per CTA (warp __popcll partials, shuffle tree, smem across warps)
__device__ __forceinline__ void block_accumulate(unsigned long long v, unsigned long long *out) {
    __shared__ unsigned long long s_part[32];
    v = warp_sum(v);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) s_part[wid] = v;
    __syncthreads();
    if (wid == 0) {
        int nw = (blockDim.x + 31) >> 5;
        v = lane < nw ? s_part[lane] : 0ull;
        v = warp_sum(v);
        if (lane == 0 && v) atomicAdd(out, v);
    }
}

__device__ __forceinline__ ulonglong2 ld_stream(const ulonglong2 *p) {   // read-once operand: non-coherent, no L1 allocate
    ulonglong2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ ulonglong2 ld_rw(const ulonglong2 *p) {       // in-place operand: coherent path, evict-first
    ulonglong2 r;
    asm volatile("ld.global.cs.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(ulonglong2 *p, ulonglong2 v) {
    asm volatile("st.global.cs.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.x), "l"(v.y) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// a op= b (and / or / xor), optional fused popcount; also applies the per-bin sentinel algebra to state[]
// ------------------------------------------------------------------------------------------------------------------
enum { OP_AND = 0, OP_OR = 1, OP_XOR = 2 };

template <int OP>
__device__ __forceinline__ unsigned long long apply(unsigned long long a, unsigned long long b) {
    return OP == OP_AND ? (a & b) : OP == OP_OR ? (a | b) : (a ^ b);
}

// binBits.c:230-296 restated on the state byte
template <int OP>
__device__ __forceinline__ uint8_t state_op(uint8_t a, uint8_t b) {
    if (OP == OP_AND) {
        if (a == BZ) return BZ;
        if (b == BZ) return BZ;
        if (b == BO) return a;
        return BA;                       // a in {O,A}, b == A : clone or byte-and
    } else if (OP == OP_OR) {
        if (a == BO) return BO;
        if (b == BO) return BO;
        if (b == BZ) return a;
        return BA;
    }
    return BA;
}

constexpr int BINOP_THREADS = 256;
constexpr int BINOP_UNROLL = 4;

template <int OP, bool COUNT>
__global__ void __launch_bounds__(BINOP_THREADS)
k_binop(ulonglong2 *__restrict__ a, const ulonglong2 *__restrict__ b, int64_t nvec,
        uint8_t *__restrict__ sa, const uint8_t *__restrict__ sb, int nbins, unsigned long long *count) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long pc = 0;
    int64_t i = tid;
    for (; i + (BINOP_UNROLL - 1) * stride < nvec; i += BINOP_UNROLL * stride) {
        ulonglong2 va[BINOP_UNROLL], vb[BINOP_UNROLL];
#pragma unroll
        for (int u = 0; u < BINOP_UNROLL; u++) {
            va[u] = ld_rw(a + i + u * stride);
            vb[u] = ld_stream(b + i + u * stride);
        }
#pragma unroll
        for (int u = 0; u < BINOP_UNROLL; u++) {
            va[u].x = apply<OP>(va[u].x, vb[u].x);
            va[u].y = apply<OP>(va[u].y, vb[u].y);
            if (COUNT) pc += __popcll(va[u].x) + __popcll(va[u].y);
            st_stream(a + i + u * stride, va[u]);
        }
    }
    for (; i < nvec; i += stride) {
        ulonglong2 va = ld_rw(a + i), vb = ld_stream(b + i);
        va.x = apply<OP>(va.x, vb.x);
        va.y = apply<OP>(va.y, vb.y);
        if (COUNT) pc += __popcll(va.x) + __popcll(va.y);
        st_stream(a + i, va);
    }
    if (sa != nullptr)
        for (int64_t k = tid; k < nbins; k += stride) sa[k] = state_op<OP>(sa[k], sb[k]);
    if (COUNT) block_accumulate(pc, count);
}

// ------------------------------------------------------------------------------------------------------------------
// Genome-wide form: one persistent launch applies a[p] op= b[p] to every pair p (the reference loop
// `for chrom in bits1: bits1[chrom].iand(bits2[chrom])`, scripts/bed_intersect_basewise.py:25-28).
// The pairs are cut into 16 KB chunks; each CTA owns one contiguous chunk range (DRAM-page friendly, and a CTA
// flushes its popcount once per pair it touches).
// ------------------------------------------------------------------------------------------------------------------
struct BatchDesc {
    ulonglong2 *a;
    const ulonglong2 *b;
    int64_t nvec;
    uint8_t *sa;
    const uint8_t *sb;
    int64_t chunk0;      // global index of this pair's first chunk
    int32_t nbins, flat;
};
constexpr int BATCH_CHUNK_VEC = 1024;   // 16 KB of a per chunk: 4 x 128-bit per thread
constexpr int BATCH_MAX_PAIRS = 1024;

template <int OP, bool COUNT>
__global__ void __launch_bounds__(BINOP_THREADS)
k_binop_batch(const BatchDesc *__restrict__ descs, int npairs, int64_t nchunks, int64_t chunks_per_cta,
              unsigned long long *__restrict__ counts) {
    __shared__ int64_t s_chunk0[BATCH_MAX_PAIRS + 1];
    for (int p = threadIdx.x; p <= npairs; p += blockDim.x) s_chunk0[p] = p < npairs ? descs[p].chunk0 : nchunks;
    __syncthreads();
    int64_t c = (int64_t)blockIdx.x * chunks_per_cta;
    const int64_t c_end = min(c + chunks_per_cta, nchunks);
    if (c >= c_end) return;
    int p = 0;
    {   // pair of the first chunk: largest p with chunk0[p] <= c
        int lo = 0, hi = npairs;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (s_chunk0[mid] <= c) lo = mid; else hi = mid;
        }
        p = lo;
    }
    unsigned long long pc = 0;
    while (c < c_end) {
        const BatchDesc d = descs[p];
        const int64_t pair_end = min(s_chunk0[p + 1], c_end);
        if (c == d.chunk0 && d.sa != nullptr)       // the CTA that owns a pair's first chunk applies the bin-state algebra
            for (int k = threadIdx.x; k < d.nbins; k += blockDim.x) d.sa[k] = state_op<OP>(d.sa[k], d.sb[k]);
        for (; c < pair_end; c++) {
            const int64_t base = (c - d.chunk0) * BATCH_CHUNK_VEC + threadIdx.x;
            ulonglong2 va[4], vb[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                int64_t i = base + u * BINOP_THREADS;
                if (i < d.nvec) {
                    va[u] = ld_rw(d.a + i);
                    vb[u] = ld_stream(d.b + i);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                int64_t i = base + u * BINOP_THREADS;
                if (i < d.nvec) {
                    va[u].x = apply<OP>(va[u].x, vb[u].x);
                    va[u].y = apply<OP>(va[u].y, vb[u].y);
                    if (COUNT) pc += __popcll(va[u].x) + __popcll(va[u].y);
                    st_stream(d.a + i, va[u]);
                }
            }
        }
        if (COUNT) {
            block_accumulate(pc, counts + p);       // uniform across the CTA: every thread walks the same chunks
            pc = 0;
            __syncthreads();
        }
        p++;
    }
}

// a = ~a over [0,size); tail bits stay zero; state: Z <-> O (binBits.c:298-317)
__global__ void __launch_bounds__(BINOP_THREADS)
k_not(ulonglong2 *__restrict__ a, int64_t nvec, int64_t nwords, uint64_t tail_mask, uint8_t *__restrict__ sa, int nbins) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < nvec; i += stride) {
        ulonglong2 v = ld_rw(a + i);
        v.x = ~v.x;
        v.y = ~v.y;
        int64_t w = 2 * i;
        if (w + 1 >= nwords - 1) {     // touches the last real word or padding
            if (w == nwords - 1) v.x &= tail_mask; else if (w >= nwords) v.x = 0;
            if (w + 1 == nwords - 1) v.y &= tail_mask; else if (w + 1 >= nwords) v.y = 0;
        }
        st_stream(a + i, v);
    }
    if (sa != nullptr)
        for (int64_t k = tid; k < nbins; k += stride) {
            uint8_t s = sa[k];
            sa[k] = s == BZ ? BO : s == BO ? BZ : BA;
        }
}

__global__ void __launch_bounds__(BINOP_THREADS)
k_popcount(const ulonglong2 *__restrict__ a, int64_t nvec, unsigned long long *count) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long pc = 0;
    int64_t i = tid;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        ulonglong2 v0 = ld_stream(a + i), v1 = ld_stream(a + i + stride), v2 = ld_stream(a + i + 2 * stride),
                   v3 = ld_stream(a + i + 3 * stride);
        pc += __popcll(v0.x) + __popcll(v0.y) + __popcll(v1.x) + __popcll(v1.y) + __popcll(v2.x) + __popcll(v2.y) +
              __popcll(v3.x) + __popcll(v3.y);
    }
    for (; i < nvec; i += stride) {
        ulonglong2 v = ld_stream(a + i);
        pc += __popcll(v.x) + __popcll(v.y);
    }
    block_accumulate(pc, count);
}

// ------------------------------------------------------------------------------------------------------------------
// set_range x n  (binBits.c:98-128): one lane per range for the two edge words (64-bit atomicOr), the whole warp
// sweeps each lane's interior words with coalesced stores of ~0.
// ------------------------------------------------------------------------------------------------------------------
constexpr int64_t LONG_WORDS = 1 << 12;          // interiors above 32 KB are swept by the whole CTA instead of one warp
constexpr int LONG_SLOTS = 64;                   // per-CTA queue of such interiors (per 256-range tile)

__global__ void __launch_bounds__(256)
k_set_ranges(uint64_t *__restrict__ words, uint8_t *__restrict__ state, int bin_size, int flat,
             const int32_t *__restrict__ start, const int32_t *__restrict__ count, int64_t n) {
    __shared__ int64_t s_long[2 * LONG_SLOTS];
    __shared__ unsigned int s_nlong;
    const int lane = threadIdx.x & 31;
    // block-uniform tile loop (every warp of a CTA runs the same number of iterations: the tile ends with barriers)
    for (int64_t tile = (int64_t)blockIdx.x * blockDim.x; tile < n; tile += (int64_t)gridDim.x * blockDim.x) {
        if (threadIdx.x == 0) s_nlong = 0;
        __syncthreads();
        const int64_t i = tile + threadIdx.x;
        int64_t mb = 0, me = 0;
        if (i < n) {
            int32_t s = __ldg(start + i), c = __ldg(count + i);
            if (c > 0) {
                int32_t last = s + c - 1;
                int64_t w0 = s >> 6, w1 = last >> 6;
                unsigned long long m0 = ~0ull << (s & 63), m1 = ~0ull >> (63 - (last & 63));
                if (w0 == w1) {
                    atomicOr((unsigned long long *)words + w0, m0 & m1);
                } else {
                    atomicOr((unsigned long long *)words + w0, m0);
                    atomicOr((unsigned long long *)words + w1, m1);
                }
                if (!flat) {
                    int b1 = last / bin_size;
                    for (int b = s / bin_size; b <= b1; b++)
                        if (state[b] == BZ) state[b] = BA;     // benign race: every writer stores BA
                }
                mb = w0 + 1;
                me = w1;
            }
        }
        // a very long interior (a chromosome arm) is not worth one warp's time: hand it to the whole CTA
        if (me - mb > LONG_WORDS) {
            unsigned int slot = atomicAdd(&s_nlong, 1u);
            if (slot < LONG_SLOTS) {
                s_long[2 * slot] = mb;
                s_long[2 * slot + 1] = me;
                me = mb;
            }
        }
        // ordinary interiors (10-30 words): the warp sweeps them one after the other with coalesced stores of ~0
        unsigned has = __ballot_sync(0xffffffffu, me > mb);
        while (has) {
            int src = __ffs(has) - 1;
            has &= has - 1;
            int64_t b = __shfl_sync(0xffffffffu, mb, src), e = __shfl_sync(0xffffffffu, me, src);
            for (int64_t w = b + lane; w < e; w += 32) words[w] = ~0ull;
        }
        __syncthreads();
        const unsigned int nl = min(s_nlong, (unsigned int)LONG_SLOTS);
        for (unsigned int r = 0; r < nl; r++) {
            const int64_t b = s_long[2 * r], e = s_long[2 * r + 1];
            const int64_t b2 = (b + 1) & ~1ll, e2 = e & ~1ll;      // 16-byte aligned core, 128-bit streaming stores
            if (threadIdx.x == 0 && b < b2) words[b] = ~0ull;
            if (threadIdx.x == 1 && e2 < e) words[e2] = ~0ull;
            ulonglong2 *v = reinterpret_cast<ulonglong2 *>(words);
            const ulonglong2 ones = make_ulonglong2(~0ull, ~0ull);
            for (int64_t k = (b2 >> 1) + threadIdx.x; k < (e2 >> 1); k += blockDim.x) st_stream(v + k, ones);
        }
        __syncthreads();
    }
}

// set / clear single bits (binBits.c:67-96)
__global__ void k_set_bits(uint64_t *__restrict__ words, uint8_t *__restrict__ state, int bin_size, int flat,
                           const int32_t *__restrict__ pos, int64_t n, int value) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int32_t p = pos[i];
        unsigned long long m = 1ull << (p & 63);
        if (value) {
            atomicOr((unsigned long long *)words + (p >> 6), m);
            if (!flat && state[p / bin_size] == BZ) state[p / bin_size] = BA;
        } else {
            atomicAnd((unsigned long long *)words + (p >> 6), ~m);
            if (!flat && state[p / bin_size] == BO) state[p / bin_size] = BA;
        }
    }
}

// Genome-wide form: range i goes into bitmap descs[which[i]] -- one launch for a whole BED file instead of one per
// chromosome (24 launches of ~260 k ranges each barely fill one wave and pay 24 launch gaps and tails).  Entries whose
// `which` is outside the table, or that do not fit their bitmap, are skipped (the shim has raised before).
struct SetDesc {
    uint64_t *words;
    uint8_t *state;
    int32_t bin_size, flat, size;
    int32_t cellbase;      // first locality bucket of this set (bucketed form only)
};
struct RangeTriple {                    // 12 bytes (a 16-byte aligned record made the radix pass slower: 0.67 -> 0.78 ms, r02p)
    int32_t w, s, c;
};

// Locality bucket of every range: (set, start >> kshift) flattened over the genome, at most 256 buckets of a few MB of
// bitmap each.  A file in random order touches the whole genome's bitmaps (386 MB for hg38, three times L2) at random:
// every edge word's atomicOr pulls its sector from DRAM and every rewritten line goes back to DRAM (ncu r02b: 6.0 GB read +
// 7.6 GB written for 50 M ranges).  Processed bucket by bucket -- one 8-bit radix pass over 13-byte (key + record) pairs -- the same
// atomics and stores hit lines that are still in L2.
__global__ void __launch_bounds__(256)
k_range_buckets(const SetDesc *__restrict__ descs, int nsets, int kshift, int group, const int32_t *__restrict__ which,
                const int32_t *__restrict__ start, const int32_t *__restrict__ count, int64_t n, uint8_t *__restrict__ keys,
                RangeTriple *__restrict__ vals) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int32_t w = __ldcs(which + i), s = __ldcs(start + i), c = __ldcs(count + i);
        int k = 255;
        if (w >= 0 && w < nsets) {
            k = group > 0 ? w / group : descs[w].cellbase + (int)((uint32_t)(s < 0 ? 0 : s) >> kshift);
            k = k < 0 ? 0 : (k > 255 ? 255 : k);
        }
        keys[i] = (uint8_t)k;
        vals[i] = RangeTriple{w, s, c};
    }
}

__device__ __forceinline__ void st_ones256(uint64_t *p) {          // one full 32-byte sector of ones (p 32-byte aligned)
    asm volatile("st.global.v4.u64 [%0], {%1, %1, %1, %1};" ::"l"(p), "l"(~0ull) : "memory");
}

// interior words [b, e) of one range := ~0, by the lane that owns the range: single words up to the first 32-byte boundary,
// whole sectors (one 256-bit store each), single words after the last boundary.  A BED-sized interior is ~16 words: four
// full-sector stores per lane, no shuffles, no per-range loop over the warp.
__device__ __forceinline__ void fill_interior(uint64_t *__restrict__ w, int64_t b, int64_t e) {
    while (b < e && (b & 3)) w[b++] = ~0ull;
    for (; b + 4 <= e; b += 4) st_ones256(w + b);
    while (b < e) w[b++] = ~0ull;
}
constexpr int64_t LANE_WORDS = 64;               // interiors up to 512 bytes are written by their own lane

// DEFER_STATE: the bin states are not touched here; the caller runs k_mark_touched_bins afterwards.  (The per-range form
// needs two integer divisions by the run-time bin size; on a whole file those were most of the kernel's instructions.)
// Bucketed form (records from k_range_buckets + the radix pass; bin states deferred to k_mark_touched_bins).  No barriers,
// no shared memory: every thread takes SET_ILP consecutive tiles' records at once (their loads and the descriptor lookups
// are independent, so four ranges per thread are in flight), does the edge words with atomicOr and writes BED-sized
// interiors itself with full-sector stores; only an interior above 512 bytes is swept by the whole warp.
constexpr int SET_ILP = 4;
__global__ void __launch_bounds__(256)
k_set_ranges_bucketed(const SetDesc *__restrict__ descs, int nsets, const RangeTriple *__restrict__ recs, int64_t n) {
    const int lane = threadIdx.x & 31;
    const int64_t tile = (int64_t)blockDim.x * SET_ILP;
    for (int64_t base = (int64_t)blockIdx.x * tile; base < n; base += (int64_t)gridDim.x * tile) {     // warp-uniform trip count
        RangeTriple r[SET_ILP];
#pragma unroll
        for (int j = 0; j < SET_ILP; j++) {
            const int64_t i = base + (int64_t)j * blockDim.x + threadIdx.x;
            r[j] = i < n ? recs[i] : RangeTriple{-1, 0, 0};
        }
#pragma unroll
        for (int j = 0; j < SET_ILP; j++) {
            int64_t mb = 0, me = 0;
            uint64_t *words = nullptr;
            const int32_t t = r[j].w, s = r[j].s, c = r[j].c;
            if (t >= 0 && t < nsets && c > 0 && s >= 0) {
                const SetDesc d = descs[t];
                if ((int64_t)s + c <= d.size) {
                    words = d.words;
                    const int32_t last = s + c - 1;
                    const int64_t w0 = s >> 6, w1 = last >> 6;
                    const unsigned long long m0 = ~0ull << (s & 63), m1 = ~0ull >> (63 - (last & 63));
                    if (w0 == w1) {
                        atomicOr((unsigned long long *)words + w0, m0 & m1);
                    } else {
                        atomicOr((unsigned long long *)words + w0, m0);
                        atomicOr((unsigned long long *)words + w1, m1);
                    }
                    mb = w0 + 1;
                    me = w1;
                    if (me - mb <= LANE_WORDS) {
                        fill_interior(words, mb, me);
                        me = mb;
                    }
                }
            }
            unsigned has = __ballot_sync(0xffffffffu, me > mb);          // long interiors: the warp sweeps them together
            while (has) {
                const int src = __ffs(has) - 1;
                has &= has - 1;
                const int64_t b = __shfl_sync(0xffffffffu, mb, src), e = __shfl_sync(0xffffffffu, me, src);
                uint64_t *w = (uint64_t *)__shfl_sync(0xffffffffu, (unsigned long long)words, src);
                for (int64_t k = b + lane; k < e; k += 32) w[k] = ~0ull;
            }
        }
    }
}

// AOS (bucketed, L2-resident targets): every lane writes its own interior -- the kernel is instruction-bound there and this
// form needs no per-range loop over the warp (4.1 -> 3.2 ms per 50 M ranges, r02g).  Unbucketed (DRAM-bound) input keeps
// the warp-wide coalesced sweep: scattered sector stores cost it more DRAM traffic than they save instructions (5.1 -> 6.3 ms).
template <bool AOS, bool DEFER_STATE>
__global__ void __launch_bounds__(256)
k_set_ranges_multi(const SetDesc *__restrict__ descs, int nsets, const int32_t *__restrict__ which,
                   const int32_t *__restrict__ start, const int32_t *__restrict__ count, int64_t n) {
    __shared__ int64_t s_long[2 * LONG_SLOTS];
    __shared__ uint64_t *s_long_w[LONG_SLOTS];
    __shared__ unsigned int s_nlong;
    const int lane = threadIdx.x & 31;
    for (int64_t tile = (int64_t)blockIdx.x * blockDim.x; tile < n; tile += (int64_t)gridDim.x * blockDim.x) {
        if (threadIdx.x == 0) s_nlong = 0;
        __syncthreads();
        const int64_t i = tile + threadIdx.x;
        int64_t mb = 0, me = 0;
        uint64_t *words = nullptr;
        if (i < n) {
            int32_t t, s, c;
            if (AOS) {                                  // bucketed records (k_range_buckets + radix pass): `which` points at them
                const RangeTriple r = reinterpret_cast<const RangeTriple *>(which)[i];
                t = r.w; s = r.s; c = r.c;
            } else {
                t = __ldg(which + i); s = __ldg(start + i); c = __ldg(count + i);
            }
            if (t >= 0 && t < nsets && c > 0 && s >= 0) {
                const SetDesc d = descs[t];
                if ((int64_t)s + c <= d.size) {
                    words = d.words;
                    int32_t last = s + c - 1;
                    int64_t w0 = s >> 6, w1 = last >> 6;
                    unsigned long long m0 = ~0ull << (s & 63), m1 = ~0ull >> (63 - (last & 63));
                    if (w0 == w1) {
                        atomicOr((unsigned long long *)words + w0, m0 & m1);
                    } else {
                        atomicOr((unsigned long long *)words + w0, m0);
                        atomicOr((unsigned long long *)words + w1, m1);
                    }
                    if (!DEFER_STATE && !d.flat) {
                        int b1 = last / d.bin_size;
                        for (int b = s / d.bin_size; b <= b1; b++)
                            if (d.state[b] == BZ) d.state[b] = BA;     // benign race: every writer stores BA
                    }
                    mb = w0 + 1;
                    me = w1;
                    if (AOS && me - mb <= LANE_WORDS) {     // the usual case: this lane writes its own interior
                        fill_interior(words, mb, me);
                        me = mb;
                    }
                }
            }
        }
        if (me - mb > LONG_WORDS) {
            unsigned int slot = atomicAdd(&s_nlong, 1u);
            if (slot < LONG_SLOTS) {
                s_long[2 * slot] = mb;
                s_long[2 * slot + 1] = me;
                s_long_w[slot] = words;
                me = mb;
            }
        }
        // medium interiors (512 B .. 32 KB): the warp sweeps them one after the other with coalesced stores
        unsigned has = __ballot_sync(0xffffffffu, me > mb);
        while (has) {
            int src = __ffs(has) - 1;
            has &= has - 1;
            const int64_t b = __shfl_sync(0xffffffffu, mb, src), e = __shfl_sync(0xffffffffu, me, src);
            uint64_t *w = (uint64_t *)__shfl_sync(0xffffffffu, (unsigned long long)words, src);
            for (int64_t k = b + lane; k < e; k += 32) w[k] = ~0ull;
        }
        __syncthreads();
        const unsigned int nl = min(s_nlong, (unsigned int)LONG_SLOTS);
        for (unsigned int r = 0; r < nl; r++) {
            const int64_t b = s_long[2 * r], e = s_long[2 * r + 1];
            uint64_t *w = s_long_w[r];
            for (int64_t k = b + threadIdx.x; k < e; k += blockDim.x) w[k] = ~0ull;
        }
        __syncthreads();
    }
}

// Deferred form of binBitsSetRange's bin allocation (binBits.c:104-126) after a batch of set_range calls: an ALL_ZERO bin
// becomes allocated iff one of the ranges touched it, and -- since set_range only ever sets bits and an ALL_ZERO bin had
// none -- iff it now holds a set bit.  One warp per bin ORs the bin's words (edges masked to the bin); the whole genome is
// read once, streaming.  Bins in any other state are left alone (an ALL_ONE sentinel stays a sentinel, binBits.c:112-113).
struct MarkDesc {
    const uint64_t *words;
    uint8_t *state;
    int32_t bin_size, size, nbins, bin0;      // bin0: index of this set's first bin in the launch's global bin numbering
};
__global__ void __launch_bounds__(256)
k_mark_touched_bins(const MarkDesc *__restrict__ descs, int nsets, int64_t nbins_total) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < nbins_total; g += nwarps) {
        int lo = 0, hi = nsets - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((int64_t)descs[mid].bin0 <= g) lo = mid; else hi = mid - 1;
        }
        const MarkDesc d = descs[lo];
        const int b = (int)(g - d.bin0);
        if (b >= d.nbins || d.state[b] != BZ) continue;
        const int64_t p0 = (int64_t)b * d.bin_size, p1 = min(p0 + d.bin_size, (int64_t)d.size);
        if (p0 >= p1) continue;
        const int64_t w0 = p0 >> 6, w1 = (p1 - 1) >> 6;
        unsigned long long any = 0;
        for (int64_t w = w0 + lane; w <= w1; w += 32) {
            unsigned long long x = __ldg((const unsigned long long *)d.words + w);
            if (w == w0) x &= ~0ull << (p0 & 63);
            if (w == w1 && (p1 & 63)) x &= (1ull << (p1 & 63)) - 1ull;
            any |= x;
        }
        if (__ballot_sync(0xffffffffu, any != 0) && lane == 0) d.state[b] = BA;
    }
}

// `flag` (single-CTA launches of the scalar calls only, else null): completion word in mapped host memory, written after
// the results with a system-scope fence -- the host spins on it instead of synchronising the stream (common.cuh zc_wait)
__device__ __forceinline__ void zc_signal(volatile long long *flag, long long seq) {
    if (flag == nullptr) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *flag = seq;
}

__global__ void k_read_bits(const uint64_t *__restrict__ words, const int32_t *__restrict__ pos, int64_t n,
                            uint8_t *__restrict__ out, volatile long long *flag, long long seq) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int32_t p = pos[i];
        out[i] = (uint8_t)((words[p >> 6] >> (p & 63)) & 1ull);
    }
    zc_signal(flag, seq);
}

// ------------------------------------------------------------------------------------------------------------------
// count_range x n  (binBits.c:130-178): rank table lookup, O(1) per query
// ------------------------------------------------------------------------------------------------------------------
// Rank lines.  A count_range query reads at two random positions of a genome of bitmaps (386 MB for hg38: three times
// L2), so what it costs is DRAM sectors and the latency of fetching them.  A separate rank table (one entry per word or
// per 256-bit block) means two sectors per position -- the table's and the bitmap's.  Here the rank is stored NEXT TO the
// bits it belongs to: the count structure is a copy of the bitmap cut into 32-byte lines, one DRAM sector each,
//     line i = { rank_i : set bits before bit 224 i ;  the bitmap's 32-bit pieces 7 i .. 7 i + 6 }          (8 x uint32)
// (the bitmap is LSB-first, so as a uint32 array piece j holds bits 32 j .. 32 j + 31 and 224 = 7 x 32 keeps lines piece-
// aligned).  rank(p) = line[p / 224].rank + popcount of that line's bits below p % 224: ONE sector and ONE independent
// 256-bit load per position, two per query.  Built lazily (line popcounts -> one scan -> one interleaving pass, all
// streaming; a whole genome in three launches) and invalidated by every mutation; costs 8/7 of the bitmap in HBM.
// [r02f/r02g: 64-byte lines {rank, 7 x uint64} lost to the table they replaced -- 3.2 / 4.1 ms against 2.69 ms per 50 M
//  queries -- because half the positions needed the line's second sector: a dependent load, or twice the DRAM sectors.]
constexpr int RL_PIECES = 7;                   // 32-bit bitmap pieces per rank line
constexpr uint32_t RL_BITS = 32u * RL_PIECES;  // 224

// popcount of line i's pieces (scan input); the bitmap is read as uint32 pieces, npieces = 2 * nwords_alloc
__device__ __forceinline__ uint32_t line_popc(const uint32_t *__restrict__ pieces, int64_t npieces, int64_t nlines, int64_t i) {
    uint32_t c = 0;
    if (i < nlines) {
#pragma unroll
        for (int k = 0; k < RL_PIECES; k++) {
            const int64_t j = i * RL_PIECES + k;
            if (j < npieces) c += (uint32_t)__popc(__ldg(pieces + j));
        }
    }
    return c;
}

// every set's lines (its nlines + 1, sentinel included) are numbered consecutively across the call's sets: line popcounts
// -> ONE 64-bit exclusive scan over all of them -> lines, with each set's own first prefix subtracted.  Three launches for
// a whole genome (three per chromosome were launch-bound: 0.93 ms for hg38, r02f; 0.30 ms now).
struct LineDesc {
    const uint32_t *pieces;
    uint32_t *lines;
    int64_t npieces, nlines, line0;          // line0: global number of this set's line 0
};
__device__ __forceinline__ int line_owner(const LineDesc *__restrict__ d, int nsets, int64_t g) {
    int lo = 0, hi = nsets - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (d[mid].line0 <= g) lo = mid; else hi = mid - 1;
    }
    return lo;
}
__global__ void __launch_bounds__(256)
k_line_popc_multi(const LineDesc *__restrict__ descs, int nsets, int64_t total, uint32_t *__restrict__ pc) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const LineDesc d = descs[line_owner(descs, nsets, g)];
        pc[g] = line_popc(d.pieces, d.npieces, d.nlines, g - d.line0);
    }
}
__global__ void __launch_bounds__(256)
k_rank_lines_multi(const LineDesc *__restrict__ descs, int nsets, int64_t total, const unsigned long long *__restrict__ prefix) {
    // one thread per 16-byte half of a line: coalesced 128-bit stores
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total * 2; q += stride) {
        const int64_t g = q >> 1;
        const LineDesc d = descs[line_owner(descs, nsets, g)];
        const int64_t i = g - d.line0;
        const int half = (int)(q & 1);
        uint32_t v[4];
#pragma unroll
        for (int h = 0; h < 4; h++) {
            const int k = 4 * half + h;                     // 0 = rank, 1..7 = pieces 7i .. 7i+6
            if (k == 0) {
                v[h] = (uint32_t)(prefix[g] - prefix[d.line0]);
            } else {
                const int64_t j = i * RL_PIECES + (k - 1);
                v[h] = (i < d.nlines && j < d.npieces) ? __ldg(d.pieces + j) : 0u;
            }
        }
        reinterpret_cast<uint4 *>(d.lines)[i * 2 + half] = make_uint4(v[0], v[1], v[2], v[3]);
    }
}

// rank(p): set bits in [0, p) -- one sector, one load
__device__ __forceinline__ uint32_t rank_at(const uint32_t *__restrict__ lines, uint32_t p) {
    const uint32_t i = p / RL_BITS, o = p - i * RL_BITS;
    uint32_t d[8];
    // (r02i/r02j: the load flavour -- .nc, L1::no_allocate, L2 evict_last / evict_first policies, .cg -- and
    //  cudaLimitMaxL2FetchGranularity 32/64/128 leave time and DRAM traffic unchanged: ~3 DRAM sectors per missed sector)
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7])
        : "l"(lines + 8 * (size_t)i));
    uint32_t c = d[0];
#pragma unroll
    for (int k = 0; k < RL_PIECES; k++) {
        const uint32_t lo = 32u * k;
        const uint32_t m = o >= lo + 32u ? ~0u : (o > lo ? (1u << (o - lo)) - 1u : 0u);
        c += (uint32_t)__popc(d[k + 1] & m);
    }
    return c;
}

// popcount of [s, s + c), c > 0
__device__ __forceinline__ int32_t count_span(const uint32_t *__restrict__ lines, uint32_t s, uint32_t c) {
    return (int32_t)(rank_at(lines, s + c) - rank_at(lines, s));
}

__global__ void __launch_bounds__(256)
k_count_ranges(const uint32_t *__restrict__ lines, const uint8_t *__restrict__ state,
               int bin_size, int strict, const int32_t *__restrict__ start, const int32_t *__restrict__ count, int64_t n,
               int32_t *__restrict__ out, volatile long long *flag, long long seq) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int32_t s = __ldg(start + i), c = __ldg(count + i), r = 0;
        if (c > 0) {
            r = count_span(lines, (uint32_t)s, (uint32_t)c);
            // binBits.c:155,161: an ALL_ONE *sentinel* first bin contributes (k - offset) instead of k
            if (strict && state[s / bin_size] == BO) r -= s % bin_size;
        }
        out[i] = r;
    }
    zc_signal(flag, seq);
}

// genome-wide form: query i addresses bit set which[i] (the dict lookup `bitsets[chrom]` of scripts/bed_intersect.py:46-53)
struct CountDesc {
    const uint32_t *lines;       // rank lines (see above)
    const uint8_t *state;
    int32_t bin_size, strict, size;
};

__global__ void __launch_bounds__(256, 8)
k_count_ranges_multi(const CountDesc *__restrict__ descs, int nsets, const int32_t *__restrict__ which,
                     const int32_t *__restrict__ start, const int32_t *__restrict__ count, int64_t n,
                     int32_t *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int32_t w = __ldg(which + i), s = __ldg(start + i), c = __ldg(count + i);
        int32_t r = 0;
        if (w >= 0 && w < nsets && c > 0) {
            const CountDesc d = descs[w];
            if (s >= 0 && (int64_t)s + c <= d.size) {          // out-of-range queries are the host shim's IndexError
                r = count_span(d.lines, (uint32_t)s, (uint32_t)c);
                if (d.strict && d.state[s / d.bin_size] == BO) r -= s % d.bin_size;
            }
        }
        out[i] = r;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// next_set / next_clear  (binBits.c:180-228, bits.c:143-190): first p in [start,end) with bit == val, else end
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_next(const uint64_t *__restrict__ words, int64_t start, int64_t end, int val, unsigned long long *result) {
    const int64_t w0 = start >> 6, w1 = (end - 1) >> 6;   // end > start guaranteed by the caller
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w = w0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w <= w1; w += stride) {
        if ((unsigned long long)(w << 6) >= *(volatile unsigned long long *)result) break;   // someone found an earlier one
        unsigned long long x = words[w];
        if (!val) x = ~x;
        if (w == w0) x &= ~0ull << (start & 63);
        if (w == w1 && (end & 63)) x &= (1ull << (end & 63)) - 1ull;
        if (x) atomicMin(result, (unsigned long long)((w << 6) + __ffsll((long long)x) - 1));
    }
}

// The scalar next_set / next_clear almost always finds its bit close by (the run-extraction idiom of the scripts walks
// from run to run): one warp scans the first NEAR_WORDS words and writes the answer (or -1) straight into mapped host
// memory, completion word last -- no copies, no stream synchronise.  Only a miss pays for the grid-wide kernel above.
constexpr int NEAR_WORDS = 32 * 16;

__global__ void __launch_bounds__(32)
k_next_near(const uint64_t *__restrict__ words, int64_t start, int64_t end, int val, long long *__restrict__ out,
            volatile long long *flag, long long seq) {
    const int64_t w0 = start >> 6, w1 = (end - 1) >> 6;
    long long found = -1;
    for (int j = 0; j < NEAR_WORDS / 32 && found < 0; j++) {
        const int64_t w = w0 + j * 32 + threadIdx.x;
        unsigned long long x = 0;
        if (w <= w1) {
            x = words[w];
            if (!val) x = ~x;
            if (w == w0) x &= ~0ull << (start & 63);
            if (w == w1 && (end & 63)) x &= (1ull << (end & 63)) - 1ull;
        }
        const unsigned any = __ballot_sync(0xffffffffu, x != 0);
        if (any) {
            const int src = __ffs((int)any) - 1;
            const unsigned long long xs = __shfl_sync(0xffffffffu, x, src);
            found = ((w0 + j * 32 + src) << 6) + __ffsll((long long)xs) - 1;
        } else if (w0 + (j + 1) * 32 > w1) {
            found = end;                                   // scanned to the end of the range: nothing there
        }
    }
    if (threadIdx.x == 0) *out = found;
    zc_signal(flag, seq);
}

// ------------------------------------------------------------------------------------------------------------------
// run extraction: transition bits t = w ^ (w<<1 | carry); transitions alternate run start / run end.
// pass A counts transitions per CTA tile, a device scan turns them into offsets, pass B writes positions in order.
// ------------------------------------------------------------------------------------------------------------------
constexpr int RUN_THREADS = 256;
constexpr int RUN_WPT = 4;                       // words per thread (32 B contiguous)
constexpr int RUN_TILE = RUN_THREADS * RUN_WPT;  // words per CTA

__device__ __forceinline__ void load_transitions(const uint64_t *__restrict__ words, int64_t nwords_alloc, int64_t w,
                                                 unsigned long long t[RUN_WPT]) {
    unsigned long long v[RUN_WPT];
    unsigned long long prev = 0;
    if (w < nwords_alloc) {
        const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(words + w);
        ulonglong2 a = __ldg(p), b = __ldg(p + 1);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
        if (w > 0) prev = __ldg((const unsigned long long *)words + w - 1);
    } else {
        v[0] = v[1] = v[2] = v[3] = 0;
    }
#pragma unroll
    for (int k = 0; k < RUN_WPT; k++) {
        t[k] = v[k] ^ ((v[k] << 1) | (prev >> 63));
        prev = v[k];
    }
}

__global__ void __launch_bounds__(RUN_THREADS)
k_runs_count(const uint64_t *__restrict__ words, int64_t nwords_alloc, uint32_t *__restrict__ tile_count) {
    typedef cub::BlockReduce<uint32_t, RUN_THREADS> BR;
    __shared__ typename BR::TempStorage tmp;
    int64_t w = ((int64_t)blockIdx.x * RUN_THREADS + threadIdx.x) * RUN_WPT;
    unsigned long long t[RUN_WPT];
    load_transitions(words, nwords_alloc, w, t);
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < RUN_WPT; k++) c += __popcll(t[k]);
    uint32_t tot = BR(tmp).Sum(c);
    if (threadIdx.x == 0) tile_count[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(RUN_THREADS)
k_runs_fill(const uint64_t *__restrict__ words, int64_t nwords_alloc, const uint32_t *__restrict__ tile_offset,
            int32_t *__restrict__ run_s, int32_t *__restrict__ run_e) {
    typedef cub::BlockScan<uint32_t, RUN_THREADS> BS;
    __shared__ typename BS::TempStorage tmp;
    int64_t w = ((int64_t)blockIdx.x * RUN_THREADS + threadIdx.x) * RUN_WPT;
    unsigned long long t[RUN_WPT];
    load_transitions(words, nwords_alloc, w, t);
    uint32_t c = 0, off;
#pragma unroll
    for (int k = 0; k < RUN_WPT; k++) c += __popcll(t[k]);
    BS(tmp).ExclusiveSum(c, off);
    off += tile_offset[blockIdx.x];
#pragma unroll
    for (int k = 0; k < RUN_WPT; k++) {
        unsigned long long x = t[k];
        while (x) {
            int b = __ffsll((long long)x) - 1;
            x &= x - 1;
            int32_t p = (int32_t)(((w + k) << 6) + b);
            if (off & 1) run_e[off >> 1] = p; else run_s[off >> 1] = p;
            off++;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Runs inside ranges: for range i the maximal runs of bits == val inside [start[i], end[i]), clipped to the range -- the
// generators bits_set_in_range / bits_clear_in_range of lib/bx/intervals/operations/__init__.py:10-33 (the per-interval
// `pieces` of operations/intersect.py:62-70 and subtract.py:66-72) for a whole file.  One thread per range walks its
// words (BED intervals span a few dozen words); FILL = false counts the runs, FILL = true writes them at the CSR offsets.
// ------------------------------------------------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(256)
k_runs_in_ranges(const uint64_t *__restrict__ words, int32_t size, const int32_t *__restrict__ start,
                 const int32_t *__restrict__ end, int64_t n, int val, int32_t *__restrict__ cnt,
                 const int64_t *__restrict__ off, int32_t *__restrict__ out_s, int32_t *__restrict__ out_e) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int64_t s = start[i], e = end[i];
        if (s < 0) s = 0;
        if (e > size) e = size;
        int32_t c = 0;
        int64_t o = FILL ? off[i] : 0;
        if (s < e) {
            const int64_t w0 = s >> 6, w1 = (e - 1) >> 6;
            unsigned long long prev = 0;                      // the bit in front of the range counts as "not val"
            int64_t run_start = -1;
            for (int64_t w = w0; w <= w1; w++) {
                unsigned long long x = __ldg((const unsigned long long *)words + w);
                if (!val) x = ~x;
                if (w == w0) x &= ~0ull << (s & 63);
                if (w == w1 && (e & 63)) x &= (1ull << (e & 63)) - 1ull;
                unsigned long long rise = x & ~((x << 1) | prev);          // run starts in this word
                if (!FILL) {
                    c += __popcll(rise);
                } else {
                    unsigned long long fall = ~x & ((x << 1) | prev);      // first non-val bit after a run
                    while (rise | fall) {
                        const int br = rise ? __ffsll((long long)rise) - 1 : 64, bf = fall ? __ffsll((long long)fall) - 1 : 64;
                        if (br < bf) {
                            run_start = (w << 6) + br;
                            rise &= rise - 1;
                        } else {
                            out_s[o] = (int32_t)run_start;
                            out_e[o] = (int32_t)((w << 6) + bf);
                            o++;
                            fall &= fall - 1;
                        }
                    }
                }
                prev = x >> 63;
            }
            if (FILL && prev) {                                // the last run reaches the end of the range
                out_s[o] = (int32_t)run_start;
                out_e[o] = (int32_t)e;
            }
        }
        if (!FILL) cnt[i] = c;
    }
}

__global__ void k_fill_u8(uint8_t *p, int64_t n, uint8_t v) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}
__global__ void k_mask_tail(uint64_t *words, int64_t nwords, int64_t nwords_alloc, uint64_t tail_mask) {
    int64_t i = nwords - 1 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == nwords - 1) words[i] &= tail_mask; else if (i < nwords_alloc) words[i] = 0;
}

// ------------------------------------------------------------------------------------------------------------------
// Per-chromosome counters of a whole file (the numbers the multi-GPU runs reduce with NCCL): for every key k
//   stats[2k]   += #(entries with key == k and val >= threshold)      e.g. BED lines that overlap (bed_intersect.py:53)
//   stats[2k+1] += sum of val over the entries with key == k          e.g. overlapping bases / counted positions
// Warp-aggregated (match_any groups the lanes of one key: one shared-memory atomic per key and warp), one global atomic
// per key and CTA.  Bytes: 8 per entry.
// ------------------------------------------------------------------------------------------------------------------
constexpr int STATS_MAX_KEYS = 1024;
constexpr int STATS_FLUSH = 256;          // tiles between flushes: 256 x 256 entries keep the 32-bit partial sums exact

__global__ void __launch_bounds__(256)
k_group_stats(const int32_t *__restrict__ key, const int32_t *__restrict__ val, int64_t n, int nkeys, int32_t threshold,
              unsigned long long *__restrict__ stats) {
    // per-CTA partial counters; a value is accumulated as (v >> 16, v & 0xffff) so that 32-bit shared atomics suffice
    __shared__ unsigned int s_cnt[STATS_MAX_KEYS], s_lo[STATS_MAX_KEYS];
    __shared__ int s_hi[STATS_MAX_KEYS];
    auto flush = [&]() {
        __syncthreads();
        for (int k = threadIdx.x; k < nkeys; k += blockDim.x) {
            const long long sum = (long long)s_hi[k] * 65536ll + (long long)s_lo[k];
            if (s_cnt[k]) atomicAdd(stats + 2 * k, (unsigned long long)s_cnt[k]);
            if (sum) atomicAdd(stats + 2 * k + 1, (unsigned long long)sum);
            s_cnt[k] = 0; s_lo[k] = 0; s_hi[k] = 0;
        }
        __syncthreads();
    };
    for (int k = threadIdx.x; k < nkeys; k += blockDim.x) { s_cnt[k] = 0; s_lo[k] = 0; s_hi[k] = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // four consecutive entries per thread and iteration (two 128-bit loads): the loop is a dependent chain of warp votes and
    // shared atomics per entry, so more entries in flight per thread is what hides it (0.23 ms per 50 M entries before)
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    int tiles = 0;
    auto add = [&](int32_t k, int32_t v) {
        if (k < 0 || k >= nkeys) k = -1;
        const int k0 = __shfl_sync(0xffffffffu, k, 0);
        if (__all_sync(0xffffffffu, k == k0)) {
            // sorted input: the whole warp holds one key -- reduce in registers, one lane updates the counters
            if (k0 >= 0) {
                const unsigned c = (unsigned)__popc(__ballot_sync(0xffffffffu, v >= threshold));
                const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)v & 0xffffu);
                const int hi = __reduce_add_sync(0xffffffffu, v >> 16);
                if (lane == 0) {
                    if (c) atomicAdd(&s_cnt[k0], c);
                    if (lo) atomicAdd(&s_lo[k0], lo);
                    if (hi) atomicAdd(&s_hi[k0], hi);
                }
            }
        } else if (k >= 0) {
            if (v >= threshold) atomicAdd(&s_cnt[k], 1u);
            if (v & 0xffff) atomicAdd(&s_lo[k], (unsigned)v & 0xffffu);
            if (v >> 16) atomicAdd(&s_hi[k], v >> 16);
        }
    };
    const bool vec = ((reinterpret_cast<uintptr_t>(key) | reinterpret_cast<uintptr_t>(val)) & 15) == 0;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x * 4; base < n; base += stride) {             // CTA-uniform
        const int64_t i = base + 4 * (int64_t)threadIdx.x;
        int32_t k[4] = {-1, -1, -1, -1}, v[4] = {0, 0, 0, 0};
        if (vec && i + 3 < n) {
            const int4 kk = __ldcs(reinterpret_cast<const int4 *>(key + i)), vv = __ldcs(reinterpret_cast<const int4 *>(val + i));
            k[0] = kk.x; k[1] = kk.y; k[2] = kk.z; k[3] = kk.w;
            v[0] = vv.x; v[1] = vv.y; v[2] = vv.z; v[3] = vv.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (i + j < n) { k[j] = __ldcs(key + i + j); v[j] = __ldcs(val + i + j); }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) add(k[j], v[j]);
        if (++tiles == STATS_FLUSH / 4) {
            flush();
            tiles = 0;
        }
    }
    flush();
}

// out[k] = popcount of set k, read off the last entry of its rank table
struct TotalDesc {
    const uint32_t *rank_last;       // rank of the sentinel line = popcount of the whole bitmap
};
__global__ void k_gather_totals(const TotalDesc *__restrict__ d, int n, long long *__restrict__ out, int64_t out_stride) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[(int64_t)k * out_stride] = d[k].rank_last ? (long long)*d[k].rank_last : 0ll;
}

// ------------------------------------------------------------------------------------------------------------------
// host side of the C ABI
// ------------------------------------------------------------------------------------------------------------------
static inline uint64_t tail_mask_of(int32_t size) { return (size & 63) ? ((1ull << (size & 63)) - 1ull) : ~0ull; }
static inline int stream_grid(int64_t nvec, int unroll) {
    return grid_for(cdiv(nvec, (int64_t)BINOP_THREADS * unroll), 8);
}

static int check_pair(const bxg_bits *a, const bxg_bits *b) {
    if (!a || !b) return set_error(BXG_ERR_ARG, "null bitset handle");
    if (a->size != b->size) return set_error(BXG_ERR_MISMATCH, "BitSets must have the same size");
    if (a->flat != b->flat || a->bin_size != b->bin_size)
        return set_error(BXG_ERR_MISMATCH, "bitsets have different bin geometry (%d vs %d)", a->bin_size, b->bin_size);
    return BXG_OK;
}

static void invalidate(bxg_bits *b) {
    b->rank_valid = false;
    b->nruns = -1;
}

template <int OP, bool COUNT>
static int binop(bxg_bits *a, const bxg_bits *b, unsigned long long *d_count) {
    int64_t nvec = a->nwords_alloc / 2;
    BXG_LAUNCH((k_binop<OP, COUNT>), stream_grid(nvec, BINOP_UNROLL), BINOP_THREADS, 0,
               (ulonglong2 *)a->words, (const ulonglong2 *)b->words, nvec,
               a->flat ? nullptr : a->state, b->state, a->nbins, d_count);
    invalidate(a);
    return BXG_OK;
}

extern "C" {

int bxg_bits_create(int32_t size, int32_t granularity, bxg_bits_t **out) {
    BXG_TRY(ensure_init());
    if (!out) return set_error(BXG_ERR_ARG, "out is null");
    if (size <= 0) return set_error(BXG_ERR_ARG, "bitset size must be in [1, 2^31-1], got %d", size);
    if (granularity < 0) return set_error(BXG_ERR_ARG, "granularity must be >= 0");
    bxg_bits *b = new bxg_bits();
    b->size = size;
    if (granularity == 0) {
        b->flat = 1;
        b->bin_size = size;
        b->nbins = 1;
    } else {
        // binBits.c:8-17 -- float32 division, double ceil
        b->bin_size = (int32_t)ceil(size / (float)granularity);
        b->nbins = (int32_t)ceil(size / (float)b->bin_size);
        // the reference would index past its bins array if float rounding made nbins*bin_size < size (UB there);
        // keep our state[] large enough for every position instead
        int64_t need = ((int64_t)size + b->bin_size - 1) / b->bin_size;
        if (need > b->nbins) b->nbins = (int32_t)need;
    }
    b->nwords = ((int64_t)size + 63) >> 6;
    b->nwords_alloc = (b->nwords + 3) & ~3ll;
    Context &c = ctx();
    cudaError_t e = cudaMalloc(&b->words, (size_t)b->nwords_alloc * 8);
    if (e == cudaSuccess) e = cudaMalloc(&b->state, (size_t)b->nbins);
    if (e != cudaSuccess) {
        cudaFree(b->words);
        delete b;
        return set_error(BXG_ERR_CUDA, "cudaMalloc(bitset %d bits) failed: %s", size, cudaGetErrorString(e));
    }
    BXG_CUDA(cudaMemsetAsync(b->words, 0, (size_t)b->nwords_alloc * 8, c.stream));
    BXG_CUDA(cudaMemsetAsync(b->state, BZ, (size_t)b->nbins, c.stream));
    *out = b;
    return BXG_OK;
}

int bxg_bits_free(bxg_bits_t *b) {
    if (!b) return BXG_OK;
    cudaStreamSynchronize(ctx().stream);
    cudaFree(b->words);
    cudaFree(b->state);
    cudaFree(b->rank);
    cudaFree(b->run_s);
    cudaFree(b->run_e);
    cudaFree(b->rr_s);
    cudaFree(b->rr_e);
    delete b;
    return BXG_OK;
}

int bxg_bits_geometry(const bxg_bits_t *b, int32_t *size, int32_t *bin_size, int32_t *nbins) {
    if (!b) return set_error(BXG_ERR_ARG, "null bitset handle");
    if (size) *size = b->size;
    if (bin_size) *bin_size = b->bin_size;
    if (nbins) *nbins = b->nbins;
    return BXG_OK;
}

int bxg_bits_clone(const bxg_bits_t *b, bxg_bits_t **out) {
    if (!b) return set_error(BXG_ERR_ARG, "null bitset handle");
    bxg_bits *n = new bxg_bits();
    n->size = b->size; n->bin_size = b->bin_size; n->nbins = b->nbins; n->flat = b->flat;
    n->nwords = b->nwords; n->nwords_alloc = b->nwords_alloc; n->maybe_one = b->maybe_one;
    BXG_CUDA(cudaMalloc(&n->words, (size_t)n->nwords_alloc * 8));
    BXG_CUDA(cudaMalloc(&n->state, (size_t)n->nbins));
    BXG_CUDA(cudaMemcpyAsync(n->words, b->words, (size_t)n->nwords_alloc * 8, cudaMemcpyDeviceToDevice, ctx().stream));
    BXG_CUDA(cudaMemcpyAsync(n->state, b->state, (size_t)n->nbins, cudaMemcpyDeviceToDevice, ctx().stream));
    *out = n;
    return BXG_OK;
}

int bxg_bits_set_ranges(bxg_bits_t *b, const int32_t *start, const int32_t *count, int64_t n, int loc) {
    if (!b) return set_error(BXG_ERR_ARG, "null bitset handle");
    if (n <= 0) return BXG_OK;
    const void *ds, *dc;
    BXG_TRY(stage_in(0, start, (size_t)n * 4, loc, &ds));
    BXG_TRY(stage_in(1, count, (size_t)n * 4, loc, &dc));
    BXG_LAUNCH(k_set_ranges, grid_for(cdiv(n, 256), 8), 256, 0, b->words, b->state, b->bin_size, b->flat,
               (const int32_t *)ds, (const int32_t *)dc, n);
    invalidate(b);
    if (loc == BXG_HOST) BXG_CUDA(cudaStreamSynchronize(ctx().stream));   // caller may reuse its arrays on return
    return BXG_OK;
}

int bxg_bits_set_ranges_multi(bxg_bits_t *const *sets, int32_t nsets, const int32_t *which, const int32_t *start,
                              const int32_t *count, int64_t n, int loc) {
    BXG_TRY(ensure_init());
    if (nsets <= 0 || nsets > BATCH_MAX_PAIRS) return set_error(BXG_ERR_ARG, "nsets must be in [1, %d]", BATCH_MAX_PAIRS);
    if (n <= 0) return BXG_OK;
    static SetDesc h_desc[BATCH_MAX_PAIRS];
    int64_t total_bits = 0;
    for (int k = 0; k < nsets; k++) {
        bxg_bits *b = sets[k];
        if (!b) {                                     // a chromosome this process holds no bitmap for: its ranges are skipped
            h_desc[k] = SetDesc{nullptr, nullptr, 1, 1, 0, 0};
            continue;
        }
        h_desc[k] = SetDesc{b->words, b->state, b->bin_size, b->flat ? 1 : 0, b->size, 0};
        total_bits += b->size;
    }
    Context &c = ctx();
    // bucket the ranges by position first when the bitmaps do not fit L2 and there is enough work to pay for the pass
    static const int bucket_mode = [] {
        const char *e = getenv("BXB200_SET_BUCKETS");      // 0: never, 1: always, unset: by size
        return e ? atoi(e) : -1;
    }();
    const bool bucketed = bucket_mode == 1 || (bucket_mode != 0 && n >= (1 << 20) && total_bits / 8 > c.l2_bytes / 2);
    // a whole file at once: mark the touched bins afterwards from the bitmaps (one streaming pass) instead of per range
    const bool defer = bucketed || (int64_t)n * 64 >= total_bits / 64;
    int kshift = 0, group = 0;
    if (bucketed) {
        if (nsets > 128) {
            group = (nsets + 255) / 256;               // many small sets: neighbouring sets share a bucket
        } else {
            for (kshift = 16;; kshift++) {             // the coarsest split that still gives every set its own buckets
                int64_t cells = 0;
                for (int k = 0; k < nsets; k++)
                    if (sets[k]) cells += ((int64_t)sets[k]->size >> kshift) + 1;
                if (cells <= 256) break;
            }
            int32_t base = 0;
            for (int k = 0; k < nsets; k++) {
                h_desc[k].cellbase = base;
                if (sets[k]) base += (int32_t)(((int64_t)sets[k]->size >> kshift) + 1);
            }
        }
    }
    void *d_desc;
    BXG_TRY(scratch(3, (sizeof(SetDesc) + sizeof(MarkDesc)) * (size_t)nsets, &d_desc));      // [SetDesc x nsets | MarkDesc x <= nsets]
    BXG_CUDA(cudaMemcpyAsync(d_desc, h_desc, sizeof(SetDesc) * (size_t)nsets, cudaMemcpyHostToDevice, c.stream));
    const void *dw, *ds, *dc;
    BXG_TRY(stage_in(0, which, (size_t)n * 4, loc, &dw));
    BXG_TRY(stage_in(1, start, (size_t)n * 4, loc, &ds));
    BXG_TRY(stage_in(5, count, (size_t)n * 4, loc, &dc));
    if (bucketed) {
        void *d_k0, *d_k1, *d_v0, *d_v1, *tmp;
        BXG_TRY(scratch(2, (size_t)n, &d_k0));
        BXG_TRY(scratch(4, (size_t)n, &d_k1));
        BXG_TRY(scratch(6, (size_t)n * sizeof(RangeTriple), &d_v0));
        if ((size_t)n * sizeof(RangeTriple) > c.bucket_cap) {            // second value buffer: its own grow-only slot
            BXG_CUDA(cudaStreamSynchronize(c.stream));
            cudaFree(c.bucket_buf);
            c.bucket_buf = nullptr;
            c.bucket_cap = 0;
            BXG_CUDA(cudaMalloc(&c.bucket_buf, (size_t)n * sizeof(RangeTriple) + 256));
            c.bucket_cap = (size_t)n * sizeof(RangeTriple) + 256;
        }
        d_v1 = c.bucket_buf;
        BXG_LAUNCH(k_range_buckets, grid_for(cdiv(n, 256), 8), 256, 0, (const SetDesc *)d_desc, nsets, kshift, group,
                   (const int32_t *)dw, (const int32_t *)ds, (const int32_t *)dc, n, (uint8_t *)d_k0, (RangeTriple *)d_v0);
        size_t tmp_bytes = 0;
        BXG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const uint8_t *)d_k0, (uint8_t *)d_k1,
                                                 (const RangeTriple *)d_v0, (RangeTriple *)d_v1, n, 0, 8, c.stream));
        BXG_TRY(scratch(7, tmp_bytes, &tmp));
        prof_begin("cub::DeviceRadixSort(range buckets)");
        BXG_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, (const uint8_t *)d_k0, (uint8_t *)d_k1,
                                                 (const RangeTriple *)d_v0, (RangeTriple *)d_v1, n, 0, 8, c.stream));
        prof_end();
        c.launches += 3;
        // (8 CTAs per SM although 6 are resident: a grid of exactly the resident CTAs measured slower, 2.06 -> 2.37 ms, r02t --
        //  ranges differ in length, late CTAs even out the tail)
        BXG_LAUNCH(k_set_ranges_bucketed, grid_for(cdiv(n, 256 * SET_ILP), 8), 256, 0, (const SetDesc *)d_desc, nsets,
                   (const RangeTriple *)d_v1, n);
    } else if (defer) {
        BXG_LAUNCH((k_set_ranges_multi<false, true>), grid_for(cdiv(n, 256), 8), 256, 0, (const SetDesc *)d_desc, nsets,
                   (const int32_t *)dw, (const int32_t *)ds, (const int32_t *)dc, n);
    } else {
        BXG_LAUNCH((k_set_ranges_multi<false, false>), grid_for(cdiv(n, 256), 8), 256, 0, (const SetDesc *)d_desc, nsets,
                   (const int32_t *)dw, (const int32_t *)ds, (const int32_t *)dc, n);
    }
    if (defer) {                                    // bin states of the whole batch in one streaming pass over the bitmaps
        static MarkDesc h_mark[BATCH_MAX_PAIRS];
        int64_t nb = 0;
        int m = 0;
        for (int k = 0; k < nsets; k++) {
            bxg_bits *b = sets[k];
            if (!b || b->flat) continue;
            h_mark[m++] = MarkDesc{b->words, b->state, b->bin_size, b->size, b->nbins, (int32_t)nb};
            nb += b->nbins;
        }
        if (m > 0) {
            void *d_mark = (char *)d_desc + sizeof(SetDesc) * (size_t)nsets;
            BXG_CUDA(cudaMemcpyAsync(d_mark, h_mark, sizeof(MarkDesc) * (size_t)m, cudaMemcpyHostToDevice, c.stream));
            BXG_LAUNCH(k_mark_touched_bins, grid_for(cdiv(nb * 32, 256), 8), 256, 0, (const MarkDesc *)d_mark, m, nb);
        }
    }
    for (int k = 0; k < nsets; k++)
        if (sets[k]) invalidate(sets[k]);
    // (h_desc is pageable: cudaMemcpyAsync has staged it before returning, so the static table may be reused at once)
    if (loc == BXG_HOST) BXG_CUDA(cudaStreamSynchronize(c.stream));         // the caller may reuse its arrays on return
    return BXG_OK;
}

int bxg_bits_set_bits(bxg_bits_t *b, const int32_t *pos, int64_t n, int value, int loc) {
    if (!b) return set_error(BXG_ERR_ARG, "null bitset handle");
    if (n <= 0) return BXG_OK;
    const void *dp;
    BXG_TRY(stage_in(0, pos, (size_t)n * 4, loc, &dp));
    BXG_LAUNCH(k_set_bits, grid_for(cdiv(n, 256), 8), 256, 0, b->words, b->state, b->bin_size, b->flat,
               (const int32_t *)dp, n, value);
    invalidate(b);
    if (loc == BXG_HOST) BXG_CUDA(cudaStreamSynchronize(ctx().stream));
    return BXG_OK;
}

int bxg_bits_read(const bxg_bits_t *b, const int32_t *pos, int64_t n, uint8_t *out, int loc) {
    if (!b) return set_error(BXG_ERR_ARG, "null bitset handle");
    if (n <= 0) return BXG_OK;
    if (zc_small(loc, n)) {                 // __getitem__ of one position: no staging copies (common.cuh)
        memcpy(zc_host(0), pos, (size_t)n * 4);
        const long long seq = zc_next_seq();
        BXG_LAUNCH(k_read_bits, 1, 64, 0, b->words, (const int32_t *)zc_device(0), n, (uint8_t *)zc_device(1), zc_flag_device(), seq);
        BXG_TRY(zc_wait(seq));
        memcpy(out, zc_host(1), (size_t)n);
        return BXG_OK;
    }
    const void *dp;
    BXG_TRY(stage_in(0, pos, (size_t)n * 4, loc, &dp));
    uint8_t *dout = out;
    if (loc == BXG_HOST) {
        void *t;
        BXG_TRY(scratch(1, (size_t)n, &t));
        dout = (uint8_t *)t;
    }
    BXG_LAUNCH(k_read_bits, grid_for(cdiv(n, 256), 8), 256, 0, b->words, (const int32_t *)dp, n, dout,
               (volatile long long *)nullptr, 0ll);
    if (loc == BXG_HOST) {
        BXG_CUDA(cudaMemcpyAsync(out, dout, (size_t)n, cudaMemcpyDeviceToHost, ctx().stream));
        BXG_CUDA(cudaStreamSynchronize(ctx().stream));
    }
    return BXG_OK;
}

int bxg_bits_and(bxg_bits_t *a, const bxg_bits_t *b) {
    BXG_TRY(check_pair(a, b));
    if (a == b) return BXG_OK;              // x & x == x, states unchanged
    return binop<OP_AND, false>(a, b, nullptr);
}
int bxg_bits_or(bxg_bits_t *a, const bxg_bits_t *b) {
    BXG_TRY(check_pair(a, b));
    if (a == b) return BXG_OK;
    a->maybe_one = a->maybe_one || b->maybe_one;          // a takes over b's ALL_ONE sentinels (binBits.c:264-296)
    return binop<OP_OR, false>(a, b, nullptr);
}
int bxg_bits_xor(bxg_bits_t *a, const bxg_bits_t *b) {
    BXG_TRY(check_pair(a, b));
    if (a == b) {                           // x ^ x == 0 (the in-place kernel must not alias its read-only operand)
        BXG_CUDA(cudaMemsetAsync(a->words, 0, (size_t)a->nwords_alloc * 8, ctx().stream));
        invalidate(a);
        return BXG_OK;
    }
    return binop<OP_XOR, false>(a, b, nullptr);
}

int bxg_bits_not(bxg_bits_t *a) {
    if (!a) return set_error(BXG_ERR_ARG, "null bitset handle");
    int64_t nvec = a->nwords_alloc / 2;
    BXG_LAUNCH(k_not, stream_grid(nvec, 1), BINOP_THREADS, 0, (ulonglong2 *)a->words, nvec, a->nwords,
               tail_mask_of(a->size), a->flat ? nullptr : a->state, a->nbins);
    if (!a->flat) a->maybe_one = true;          // ALL_ZERO bins turn into ALL_ONE sentinels (binBits.c:298-317)
    invalidate(a);
    return BXG_OK;
}

static int fetch_counter(int64_t *out) {
    Context &c = ctx();
    BXG_CUDA(cudaMemcpyAsync(c.mailbox, c.d_mailbox, 8, cudaMemcpyDeviceToHost, c.stream));
    BXG_CUDA(cudaStreamSynchronize(c.stream));
    *out = c.mailbox[0];
    return BXG_OK;
}

int bxg_bits_and_count(bxg_bits_t *a, const bxg_bits_t *b, int64_t *count) {
    BXG_TRY(check_pair(a, b));
    if (a == b) return bxg_bits_count_all(a, count);
    Context &c = ctx();
    BXG_CUDA(cudaMemsetAsync(c.d_mailbox, 0, 8, c.stream));
    BXG_TRY((binop<OP_AND, true>(a, b, (unsigned long long *)c.d_mailbox)));
    if (count) BXG_TRY(fetch_counter(count));
    return BXG_OK;
}

int bxg_bits_binop_batch(int op, bxg_bits_t *const *a, const bxg_bits_t *const *b, int32_t n, int64_t *counts) {
    BXG_TRY(ensure_init());
    if (n <= 0) return BXG_OK;
    if (n > BATCH_MAX_PAIRS) return set_error(BXG_ERR_ARG, "at most %d pairs per batch", BATCH_MAX_PAIRS);
    if (op < OP_AND || op > OP_XOR) return set_error(BXG_ERR_ARG, "op must be 0 (and), 1 (or) or 2 (xor)");
    if (counts && op != OP_AND) return set_error(BXG_ERR_ARG, "fused counts are provided for op 0 (and) only");
    Context &c = ctx();
    static BatchDesc h_desc[BATCH_MAX_PAIRS];
    int64_t nchunks = 0;
    for (int p = 0; p < n; p++) {
        BXG_TRY(check_pair(a[p], b[p]));
        if (a[p] == b[p]) return set_error(BXG_ERR_ARG, "pair %d aliases one bitset", p);
        for (int q = 0; q < p; q++)
            if (a[q] == a[p]) return set_error(BXG_ERR_ARG, "bitset appears twice as a destination (pairs %d, %d)", q, p);
        BatchDesc &d = h_desc[p];
        d.a = (ulonglong2 *)a[p]->words;
        d.b = (const ulonglong2 *)b[p]->words;
        d.nvec = a[p]->nwords_alloc / 2;
        d.sa = a[p]->flat ? nullptr : a[p]->state;
        d.sb = b[p]->state;
        d.nbins = a[p]->nbins;
        d.flat = a[p]->flat;
        d.chunk0 = nchunks;
        nchunks += cdiv(d.nvec, BATCH_CHUNK_VEC);
        if (op == OP_OR) a[p]->maybe_one = a[p]->maybe_one || b[p]->maybe_one;
        invalidate(a[p]);
    }
    void *d_desc, *d_cnt;
    BXG_TRY(scratch(3, sizeof(BatchDesc) * (size_t)n, &d_desc));
    BXG_TRY(scratch(4, 8 * (size_t)n, &d_cnt));
    BXG_CUDA(cudaMemcpyAsync(d_desc, h_desc, sizeof(BatchDesc) * (size_t)n, cudaMemcpyHostToDevice, c.stream));
    if (counts) BXG_CUDA(cudaMemsetAsync(d_cnt, 0, 8 * (size_t)n, c.stream));
    // persistent grid: a whole number of CTAs per SM, each owning one contiguous chunk range
    int grid = (int)std::min<int64_t>((int64_t)c.sm_count * 8, nchunks);
    int64_t per_cta = cdiv(nchunks, grid);
    grid = (int)cdiv(nchunks, per_cta);
    const BatchDesc *dd = (const BatchDesc *)d_desc;
    unsigned long long *dc = (unsigned long long *)d_cnt;
    if (counts) {
        BXG_LAUNCH((k_binop_batch<OP_AND, true>), grid, BINOP_THREADS, 0, dd, n, nchunks, per_cta, dc);
        BXG_CUDA(cudaMemcpyAsync(counts, d_cnt, 8 * (size_t)n, cudaMemcpyDeviceToHost, c.stream));
        BXG_CUDA(cudaStreamSynchronize(c.stream));
    } else if (op == OP_AND) {
        BXG_LAUNCH((k_binop_batch<OP_AND, false>), grid, BINOP_THREADS, 0, dd, n, nchunks, per_cta, dc);
    } else if (op == OP_OR) {
        BXG_LAUNCH((k_binop_batch<OP_OR, false>), grid, BINOP_THREADS, 0, dd, n, nchunks, per_cta, dc);
    } else {
        BXG_LAUNCH((k_binop_batch<OP_XOR, false>), grid, BINOP_THREADS, 0, dd, n, nchunks, per_cta, dc);
    }
    return BXG_OK;
}

int bxg_bits_count_all(const bxg_bits_t *b, int64_t *count) {
    if (!b) return set_error(BXG_ERR_ARG, "null bitset handle");
    Context &c = ctx();
    int64_t nvec = b->nwords_alloc / 2;
    BXG_CUDA(cudaMemsetAsync(c.d_mailbox, 0, 8, c.stream));
    BXG_LAUNCH(k_popcount, stream_grid(nvec, 4), BINOP_THREADS, 0, (const ulonglong2 *)b->words, nvec,
               (unsigned long long *)c.d_mailbox);
    return fetch_counter(count);
}

struct CastU32ToU64 {
    __device__ __forceinline__ unsigned long long operator()(uint32_t v) const { return (unsigned long long)v; }
};

// rank lines of every set of the call that needs them, in three launches (see k_line_popc_multi)
static int build_rank_multi(bxg_bits_t *const *sets, int32_t nsets) {
    static LineDesc h_line[BATCH_MAX_PAIRS];
    int m = 0;
    int64_t total = 0;
    for (int k = 0; k < nsets; k++) {
        bxg_bits *b = sets[k];
        if (!b || b->rank_valid) continue;
        bool dup = false;
        for (int q = 0; q < m && !dup; q++) dup = h_line[q].pieces == (const uint32_t *)b->words;   // listed twice: built once
        if (dup) continue;
        const int64_t npieces = 2 * b->nwords_alloc, nlines = cdiv(2 * b->nwords, RL_PIECES);
        if (!b->rank) BXG_CUDA(cudaMalloc(&b->rank, (size_t)(nlines + 1) * 32));
        b->nlines = nlines;
        h_line[m++] = LineDesc{(const uint32_t *)b->words, b->rank, npieces, nlines, total};
        total += nlines + 1;
    }
    if (m == 0) return BXG_OK;
    Context &c = ctx();
    void *d_line, *d_pc, *d_prefix, *tmp;
    BXG_TRY(scratch(3, sizeof(LineDesc) * (size_t)m + 64 * (size_t)BATCH_MAX_PAIRS, &d_line));
    d_line = (char *)d_line + 64 * (size_t)BATCH_MAX_PAIRS;       // (the caller's own descriptor table sits in front)
    BXG_TRY(scratch(6, (size_t)total * 4, &d_pc));
    BXG_TRY(scratch(4, (size_t)total * 8, &d_prefix));
    BXG_CUDA(cudaMemcpyAsync(d_line, h_line, sizeof(LineDesc) * (size_t)m, cudaMemcpyHostToDevice, c.stream));
    BXG_LAUNCH(k_line_popc_multi, grid_for(cdiv(total, 256), 8), 256, 0, (const LineDesc *)d_line, m, total, (uint32_t *)d_pc);
    cub::TransformInputIterator<unsigned long long, CastU32ToU64, const uint32_t *> it((const uint32_t *)d_pc, CastU32ToU64());
    size_t tmp_bytes = 0;
    BXG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, it, (unsigned long long *)d_prefix, total, c.stream));
    BXG_TRY(scratch(7, tmp_bytes, &tmp));
    prof_begin("cub::DeviceScan::ExclusiveSum(rank lines)");
    BXG_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, it, (unsigned long long *)d_prefix, total, c.stream));
    prof_end();
    c.launches += 2;
    BXG_LAUNCH(k_rank_lines_multi, grid_for(cdiv(total * 2, 256), 8), 256, 0, (const LineDesc *)d_line, m, total,
               (const unsigned long long *)d_prefix);
    for (int k = 0; k < nsets; k++)
        if (sets[k]) sets[k]->rank_valid = true;
    return BXG_OK;
}

static int build_rank(bxg_bits *b) {
    if (b->rank_valid) return BXG_OK;
    bxg_bits_t *one[1] = {b};
    return build_rank_multi(one, 1);
}

int bxg_bits_count_ranges(bxg_bits_t *b, const int32_t *start, const int32_t *count, int64_t n, int32_t *out,
                          int strict, int loc) {
    if (!b) return set_error(BXG_ERR_ARG, "null bitset handle");
    if (n <= 0) return BXG_OK;
    BXG_TRY(build_rank(b));
    if (zc_small(loc, n)) {                 // the scalar count_range(start, count): no staging copies (common.cuh)
        memcpy(zc_host(0), start, (size_t)n * 4);
        memcpy(zc_host(1), count, (size_t)n * 4);
        const long long seq = zc_next_seq();
        BXG_LAUNCH(k_count_ranges, 1, 64, 0, b->rank, b->state, b->bin_size, (strict && !b->flat && b->maybe_one) ? 1 : 0,
                   (const int32_t *)zc_device(0), (const int32_t *)zc_device(1), n, (int32_t *)zc_device(2), zc_flag_device(), seq);
        BXG_TRY(zc_wait(seq));
        memcpy(out, zc_host(2), (size_t)n * 4);
        return BXG_OK;
    }
    const void *ds, *dc;
    BXG_TRY(stage_in(0, start, (size_t)n * 4, loc, &ds));
    BXG_TRY(stage_in(1, count, (size_t)n * 4, loc, &dc));
    int32_t *dout = out;
    if (loc == BXG_HOST) {
        void *t;
        BXG_TRY(scratch(2, (size_t)n * 4, &t));
        dout = (int32_t *)t;
    }
    BXG_LAUNCH(k_count_ranges, grid_for(cdiv(n, 256), 8), 256, 0, b->rank, b->state, b->bin_size,
               (strict && !b->flat && b->maybe_one) ? 1 : 0, (const int32_t *)ds, (const int32_t *)dc, n, dout, (volatile long long *)nullptr, 0ll);
    if (loc == BXG_HOST) {
        BXG_CUDA(cudaMemcpyAsync(out, dout, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx().stream));
        BXG_CUDA(cudaStreamSynchronize(ctx().stream));
    }
    return BXG_OK;
}

int bxg_bits_count_ranges_multi(bxg_bits_t *const *sets, int32_t nsets, const int32_t *which, const int32_t *start,
                                const int32_t *count, int64_t n, int32_t *out, int strict, int loc) {
    BXG_TRY(ensure_init());
    if (nsets <= 0 || nsets > BATCH_MAX_PAIRS) return set_error(BXG_ERR_ARG, "nsets must be in [1, %d]", BATCH_MAX_PAIRS);
    if (n <= 0) return BXG_OK;
    static CountDesc h_desc[BATCH_MAX_PAIRS];
    BXG_TRY(build_rank_multi(sets, nsets));
    for (int k = 0; k < nsets; k++) {
        bxg_bits *b = sets[k];
        if (!b) {                                     // `fields[0] in bitsets` is false (bed_intersect.py:53): count 0
            h_desc[k] = CountDesc{nullptr, nullptr, 1, 0, 0};
            continue;
        }
        h_desc[k] = CountDesc{b->rank, b->state, b->bin_size, (strict && !b->flat && b->maybe_one) ? 1 : 0, b->size};
    }
    Context &c = ctx();
    void *d_desc;
    BXG_TRY(scratch(3, sizeof(CountDesc) * (size_t)nsets, &d_desc));
    BXG_CUDA(cudaMemcpyAsync(d_desc, h_desc, sizeof(CountDesc) * (size_t)nsets, cudaMemcpyHostToDevice, c.stream));
    const void *dw, *ds, *dc;
    BXG_TRY(stage_in(0, which, (size_t)n * 4, loc, &dw));
    BXG_TRY(stage_in(1, start, (size_t)n * 4, loc, &ds));
    BXG_TRY(stage_in(5, count, (size_t)n * 4, loc, &dc));
    int32_t *dout = out;
    if (loc == BXG_HOST) {
        void *t;
        BXG_TRY(scratch(2, (size_t)n * 4, &t));
        dout = (int32_t *)t;
    }
    BXG_LAUNCH(k_count_ranges_multi, grid_for(cdiv(n, 256), 8), 256, 0, (const CountDesc *)d_desc, nsets,
               (const int32_t *)dw, (const int32_t *)ds, (const int32_t *)dc, n, dout);
    if (loc == BXG_HOST) {
        BXG_CUDA(cudaMemcpyAsync(out, dout, (size_t)n * 4, cudaMemcpyDeviceToHost, c.stream));
        BXG_CUDA(cudaStreamSynchronize(c.stream));
    }
    return BXG_OK;
}

int bxg_bits_clear(bxg_bits_t *b) {
    if (!b) return set_error(BXG_ERR_ARG, "null bitset handle");
    Context &c = ctx();
    BXG_CUDA(cudaMemsetAsync(b->words, 0, (size_t)b->nwords_alloc * 8, c.stream));
    BXG_CUDA(cudaMemsetAsync(b->state, BZ, (size_t)b->nbins, c.stream));
    b->maybe_one = false;
    invalidate(b);
    return BXG_OK;
}

int bxg_bits_count_all_multi(bxg_bits_t *const *sets, int32_t nsets, int64_t *out, int64_t out_stride, int loc) {
    BXG_TRY(ensure_init());
    if (nsets <= 0) return BXG_OK;
    if (nsets > BATCH_MAX_PAIRS) return set_error(BXG_ERR_ARG, "nsets must be in [1, %d]", BATCH_MAX_PAIRS);
    if (!out || out_stride < 1 || (loc == BXG_HOST && out_stride != 1))
        return set_error(BXG_ERR_ARG, "bad output (host output needs out_stride == 1)");
    static TotalDesc h_desc[BATCH_MAX_PAIRS];
    BXG_TRY(build_rank_multi(sets, nsets));
    for (int k = 0; k < nsets; k++) {
        bxg_bits *b = sets[k];
        if (!b) {
            h_desc[k] = TotalDesc{nullptr};
            continue;
        }
        h_desc[k] = TotalDesc{b->rank + 8 * (size_t)b->nlines};        // word 0 of the sentinel line
    }
    Context &c = ctx();
    void *d_desc;
    BXG_TRY(scratch(3, sizeof(TotalDesc) * (size_t)nsets, &d_desc));
    BXG_CUDA(cudaMemcpyAsync(d_desc, h_desc, sizeof(TotalDesc) * (size_t)nsets, cudaMemcpyHostToDevice, c.stream));
    long long *dout = (long long *)out;
    if (loc == BXG_HOST) {
        void *t;
        BXG_TRY(scratch(4, 8 * (size_t)nsets, &t));
        dout = (long long *)t;
    }
    BXG_LAUNCH(k_gather_totals, (nsets + 127) / 128, 128, 0, (const TotalDesc *)d_desc, (int)nsets, dout,
               loc == BXG_HOST ? (int64_t)1 : out_stride);
    if (loc == BXG_HOST) {
        BXG_CUDA(cudaMemcpyAsync(out, dout, 8 * (size_t)nsets, cudaMemcpyDeviceToHost, c.stream));
        BXG_CUDA(cudaStreamSynchronize(c.stream));
    }
    return BXG_OK;
}

int bxg_group_stats_i32(const int32_t *key, const int32_t *val, int64_t n, int32_t nkeys, int32_t threshold, int64_t *stats,
                        int loc) {
    BXG_TRY(ensure_init());
    if (nkeys <= 0 || nkeys > STATS_MAX_KEYS) return set_error(BXG_ERR_ARG, "nkeys must be in [1, %d]", STATS_MAX_KEYS);
    if (!stats) return set_error(BXG_ERR_ARG, "stats is null");
    if (n <= 0) return BXG_OK;
    Context &c = ctx();
    const void *dk, *dv;
    BXG_TRY(stage_in(0, key, (size_t)n * 4, loc, &dk));
    BXG_TRY(stage_in(1, val, (size_t)n * 4, loc, &dv));
    unsigned long long *dst = (unsigned long long *)stats;
    if (loc == BXG_HOST) {
        void *t;
        BXG_TRY(scratch(4, 16 * (size_t)nkeys, &t));
        dst = (unsigned long long *)t;
        BXG_CUDA(cudaMemcpyAsync(dst, stats, 16 * (size_t)nkeys, cudaMemcpyHostToDevice, c.stream));
    }
    BXG_LAUNCH(k_group_stats, grid_for(cdiv(n, 256 * 8), 4), 256, 0, (const int32_t *)dk, (const int32_t *)dv, n, (int)nkeys,
               threshold, dst);
    if (loc == BXG_HOST) {
        BXG_CUDA(cudaMemcpyAsync(stats, dst, 16 * (size_t)nkeys, cudaMemcpyDeviceToHost, c.stream));
        BXG_CUDA(cudaStreamSynchronize(c.stream));
    }
    return BXG_OK;
}

int bxg_bits_next(const bxg_bits_t *b, int32_t start, int32_t end, int val, int32_t *out) {
    if (!b) return set_error(BXG_ERR_ARG, "null bitset handle");
    if (start < 0 || end > b->size || end < start) return set_error(BXG_ERR_ARG, "next: bad range [%d,%d)", start, end);
    if (end == start) {
        *out = end;
        return BXG_OK;
    }
    Context &c = ctx();
    if (c.zc) {
        const long long seq = zc_next_seq();
        BXG_LAUNCH(k_next_near, 1, 32, 0, b->words, (int64_t)start, (int64_t)end, val, (long long *)zc_device(3), zc_flag_device(), seq);
        BXG_TRY(zc_wait(seq));
        const long long r = *(volatile long long *)zc_host(3);
        if (r >= 0) {
            *out = (int32_t)r;
            return BXG_OK;
        }
        start = (int32_t)((((int64_t)start >> 6) + NEAR_WORDS) << 6);     // continue behind the scanned words
    }
    c.mailbox[1] = end;
    BXG_CUDA(cudaMemcpyAsync(c.d_mailbox + 1, c.mailbox + 1, 8, cudaMemcpyHostToDevice, c.stream));
    int64_t nw = ((int64_t)(end - 1) >> 6) - (start >> 6) + 1;
    BXG_LAUNCH(k_next, grid_for(cdiv(nw, 256), 4), 256, 0, b->words, (int64_t)start, (int64_t)end, val,
               (unsigned long long *)(c.d_mailbox + 1));
    BXG_CUDA(cudaMemcpyAsync(c.mailbox + 1, c.d_mailbox + 1, 8, cudaMemcpyDeviceToHost, c.stream));
    BXG_CUDA(cudaStreamSynchronize(c.stream));
    *out = (int32_t)c.mailbox[1];
    return BXG_OK;
}

int bxg_bits_runs_count(bxg_bits_t *b, int64_t *nruns) {
    if (!b) return set_error(BXG_ERR_ARG, "null bitset handle");
    if (b->nruns >= 0) {
        *nruns = b->nruns;
        return BXG_OK;
    }
    Context &c = ctx();
    int64_t ntiles = cdiv(b->nwords_alloc, RUN_TILE);
    void *d_cnt, *d_off, *tmp;
    BXG_TRY(scratch(0, (size_t)(ntiles + 1) * 4, &d_cnt));
    BXG_TRY(scratch(1, (size_t)(ntiles + 1) * 4, &d_off));
    BXG_CUDA(cudaMemsetAsync((uint32_t *)d_cnt + ntiles, 0, 4, c.stream));
    BXG_LAUNCH(k_runs_count, (int)ntiles, RUN_THREADS, 0, b->words, b->nwords_alloc, (uint32_t *)d_cnt);
    size_t tmp_bytes = 0;
    BXG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (uint32_t *)d_cnt, (uint32_t *)d_off, ntiles + 1, c.stream));
    BXG_TRY(scratch(7, tmp_bytes, &tmp));
    BXG_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, (uint32_t *)d_cnt, (uint32_t *)d_off, ntiles + 1, c.stream));
    c.launches += 2;
    uint32_t *m32 = (uint32_t *)(c.mailbox + 2);
    BXG_CUDA(cudaMemcpyAsync(m32, (uint32_t *)d_off + ntiles, 4, cudaMemcpyDeviceToHost, c.stream));
    BXG_CUDA(cudaStreamSynchronize(c.stream));
    int64_t ntrans = *m32;
    int64_t nr = (ntrans + 1) / 2;
    if (nr > b->run_cap) {
        cudaFree(b->run_s);
        cudaFree(b->run_e);
        b->run_s = b->run_e = nullptr;
        b->run_cap = nr + nr / 4 + 16;
        BXG_CUDA(cudaMalloc(&b->run_s, (size_t)b->run_cap * 4));
        BXG_CUDA(cudaMalloc(&b->run_e, (size_t)b->run_cap * 4));
    }
    if (nr > 0) {
        BXG_LAUNCH(k_runs_fill, (int)ntiles, RUN_THREADS, 0, b->words, b->nwords_alloc, (const uint32_t *)d_off,
                   b->run_s, b->run_e);
        if (ntrans & 1) {   // last run reaches `size` exactly on a word boundary: no closing transition exists
            int32_t sz = b->size;
            BXG_CUDA(cudaMemcpyAsync(b->run_e + (nr - 1), &sz, 4, cudaMemcpyHostToDevice, c.stream));
            BXG_CUDA(cudaStreamSynchronize(c.stream));
        }
    }
    b->nruns = nr;
    *nruns = nr;
    return BXG_OK;
}

int bxg_bits_runs_fetch(bxg_bits_t *b, int32_t *starts, int32_t *ends, int64_t nruns) {
    if (!b) return set_error(BXG_ERR_ARG, "null bitset handle");
    if (b->nruns < 0) return set_error(BXG_ERR_STATE, "bxg_bits_runs_count must be called first");
    if (nruns != b->nruns) return set_error(BXG_ERR_ARG, "nruns mismatch (%lld vs %lld)", (long long)nruns, (long long)b->nruns);
    if (nruns == 0) return BXG_OK;
    BXG_CUDA(cudaMemcpyAsync(starts, b->run_s, (size_t)nruns * 4, cudaMemcpyDeviceToHost, ctx().stream));
    BXG_CUDA(cudaMemcpyAsync(ends, b->run_e, (size_t)nruns * 4, cudaMemcpyDeviceToHost, ctx().stream));
    BXG_CUDA(cudaStreamSynchronize(ctx().stream));
    return BXG_OK;
}

struct CastI32ToI64 {
    __device__ __forceinline__ int64_t operator()(int32_t v) const { return (int64_t)v; }
};

int bxg_bits_runs_in_ranges(bxg_bits_t *b, const int32_t *start, const int32_t *end, int64_t n, int val, int loc,
                            int64_t *offsets, int64_t *total) {
    if (!b) return set_error(BXG_ERR_ARG, "null bitset handle");
    if (n < 0 || !offsets) return set_error(BXG_ERR_ARG, "bad arguments");
    Context &c = ctx();
    offsets[0] = 0;
    b->rr_total = 0;
    if (total) *total = 0;
    if (n == 0) return BXG_OK;
    const void *ds, *de;
    BXG_TRY(stage_in(0, start, (size_t)n * 4, loc, &ds));
    BXG_TRY(stage_in(1, end, (size_t)n * 4, loc, &de));
    void *d_cnt, *d_off, *tmp;
    BXG_TRY(scratch(2, (size_t)(n + 1) * 4, &d_cnt));
    BXG_TRY(scratch(4, (size_t)(n + 1) * 8, &d_off));
    BXG_CUDA(cudaMemsetAsync((int32_t *)d_cnt + n, 0, 4, c.stream));
    const int grid = grid_for(cdiv(n, 256), 8);
    BXG_LAUNCH((k_runs_in_ranges<false>), grid, 256, 0, b->words, b->size, (const int32_t *)ds, (const int32_t *)de, n,
               val ? 1 : 0, (int32_t *)d_cnt, (const int64_t *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr);
    cub::TransformInputIterator<int64_t, CastI32ToI64, const int32_t *> it((const int32_t *)d_cnt, CastI32ToI64());
    size_t tmp_bytes = 0;
    BXG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, it, (int64_t *)d_off, n + 1, c.stream));
    BXG_TRY(scratch(7, tmp_bytes, &tmp));
    BXG_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, it, (int64_t *)d_off, n + 1, c.stream));
    c.launches += 2;
    BXG_CUDA(cudaMemcpyAsync(offsets, d_off, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, c.stream));
    BXG_CUDA(cudaStreamSynchronize(c.stream));
    const int64_t nr = offsets[n];
    if (nr > b->rr_cap) {
        cudaFree(b->rr_s);
        cudaFree(b->rr_e);
        b->rr_s = b->rr_e = nullptr;
        b->rr_cap = nr + nr / 4 + 16;
        BXG_CUDA(cudaMalloc(&b->rr_s, (size_t)b->rr_cap * 4));
        BXG_CUDA(cudaMalloc(&b->rr_e, (size_t)b->rr_cap * 4));
    }
    if (nr > 0)
        BXG_LAUNCH((k_runs_in_ranges<true>), grid, 256, 0, b->words, b->size, (const int32_t *)ds, (const int32_t *)de, n,
                   val ? 1 : 0, (int32_t *)nullptr, (const int64_t *)d_off, b->rr_s, b->rr_e);
    b->rr_total = nr;
    if (total) *total = nr;
    return BXG_OK;
}

int bxg_bits_runs_in_ranges_fetch(bxg_bits_t *b, int32_t *starts, int32_t *ends, int64_t total) {
    if (!b) return set_error(BXG_ERR_ARG, "null bitset handle");
    if (total != b->rr_total) return set_error(BXG_ERR_ARG, "total mismatch (%lld vs %lld)", (long long)total, (long long)b->rr_total);
    if (total == 0) return BXG_OK;
    BXG_CUDA(cudaMemcpyAsync(starts, b->rr_s, (size_t)total * 4, cudaMemcpyDeviceToHost, ctx().stream));
    BXG_CUDA(cudaMemcpyAsync(ends, b->rr_e, (size_t)total * 4, cudaMemcpyDeviceToHost, ctx().stream));
    BXG_CUDA(cudaStreamSynchronize(ctx().stream));
    return BXG_OK;
}

int bxg_bits_states(const bxg_bits_t *b, uint8_t *out) {
    if (!b) return set_error(BXG_ERR_ARG, "null bitset handle");
    BXG_CUDA(cudaMemcpyAsync(out, b->state, (size_t)b->nbins, cudaMemcpyDeviceToHost, ctx().stream));
    BXG_CUDA(cudaStreamSynchronize(ctx().stream));
    return BXG_OK;
}

int bxg_bits_export_words(const bxg_bits_t *b, uint64_t *out) {
    if (!b) return set_error(BXG_ERR_ARG, "null bitset handle");
    BXG_CUDA(cudaMemcpyAsync(out, b->words, (size_t)b->nwords * 8, cudaMemcpyDeviceToHost, ctx().stream));
    BXG_CUDA(cudaStreamSynchronize(ctx().stream));
    return BXG_OK;
}

int bxg_bits_import_words(bxg_bits_t *b, const uint64_t *in) {
    if (!b) return set_error(BXG_ERR_ARG, "null bitset handle");
    Context &c = ctx();
    BXG_CUDA(cudaMemcpyAsync(b->words, in, (size_t)b->nwords * 8, cudaMemcpyHostToDevice, c.stream));
    BXG_LAUNCH(k_mask_tail, 1, 32, 0, b->words, b->nwords, b->nwords_alloc, tail_mask_of(b->size));
    BXG_LAUNCH(k_fill_u8, grid_for(cdiv(b->nbins, 256), 1), 256, 0, b->state, (int64_t)b->nbins, (uint8_t)BA);
    b->maybe_one = false;
    BXG_CUDA(cudaStreamSynchronize(c.stream));
    invalidate(b);
    return BXG_OK;
}

int bxg_bits_device_words(const bxg_bits_t *b, const uint64_t **dptr, int64_t *nwords) {
    if (!b) return set_error(BXG_ERR_ARG, "null bitset handle");
    if (dptr) *dptr = b->words;
    if (nwords) *nwords = b->nwords;
    return BXG_OK;
}

}  // extern "C"

// used by aggregate.cu
const uint64_t *bxg_bits_words_internal(const bxg_bits *b) { return b ? b->words : nullptr; }
int32_t bxg_bits_size_internal(const bxg_bits *b) { return b ? b->size : 0; }
