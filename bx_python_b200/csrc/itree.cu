// itree.cu -- device interval index: radix-sorted implicit layout + batched find (count / scan / fill).
//
// Replaces IntervalNode.insert / _intersect of /root/reference/lib/bx/intervals/intersection.pyx:103-138,180-189.
// The reference treap's in-order sequence is the total order (start, end>start, +/-insertion index) (DESIGN.md
// "interval index"); find() is a filter over that sequence:  end > qs && start < qe.  So the index is
//   S[], E[], I[]  : starts, ends, item ids in that order (per tree, trees concatenated; toff[] delimits them)
//   PM[]           : per-tree running max of E  -> first possible hit   lo = first k with PM[k] >  qs
//                                                   end of candidates   hi = first k with S[k]  >= qe
//   M[l][]         : 32-ary max-of-E hierarchy  -> skip runs of non-hits inside [lo,hi) in O(log) steps
//   spS[], spPM[]  : every `stride`-th element of S / PM; staged to shared memory by one 1-D TMA bulk copy per CTA
// and find is: count pass (searches + scan of [lo,hi)), exclusive scan to int64 CSR offsets, fill pass.
#include <cub/cub.cuh>

#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "itree_search.cuh"

using namespace bxg;

constexpr int MAX_LEVELS = 6;          // 32^7 > 2^31
constexpr int MAX_SPLIT = 4096;        // splitter entries per array (16 KB each in shared memory)
constexpr int SMEM_TREES = 256;        // toff entries cached in shared memory
constexpr int FIND_THREADS = 256;
#ifndef BXB200_FIND_SERVER_DEFAULT
#define BXB200_FIND_SERVER_DEFAULT 1   // scalar find through the lingering server kernel (BXB200_FIND_SERVER=0/1 overrides)
#endif
#ifndef FIND_MIN_CTAS
#define FIND_MIN_CTAS 7            // co-resident CTAs per SM the find kernels are compiled for (register budget 36; the
                                   // direct-address search needs 38-40 without a cap and no shared-memory table, so occupancy is
                                   // the registers' to give: 6 / 7 / 8 CTAs -> count 0.312 / 0.302 / 0.302 ms, profiles/r02c).  History:
                                   // since the 8-ary probe search (7.4 sectors per query) the count kernel is bound by
                                   // latency, not by L1 wavefronts, and occupancy pays even with a few spills:
                                   // 4 / 5 / 6 CTAs -> 0.507 / 0.457 / 0.431 ms (profiles/r01s, r01t); shared memory
                                   // (19.8 KB per CTA) allows no more than 6.  4 was right for the 16-ary searches
#endif

constexpr int MAX_KLEV = 5;            // 16-ary sampled search levels: strides 1, 16, 256, 4096, 65536
constexpr int MAX_QLEV = 8;            // 8-ary sampled search levels: strides 1, 8, 64, 512, ... (one 32-byte sector per round)

struct IndexView {
    const int32_t *S, *E, *I, *PM;
    const int32_t *KS[MAX_KLEV], *KP[MAX_KLEV];   // KS[j][i] = S[i << 4j], KP likewise for PM; padded with INT32_MAX
    int32_t nk;                                   // levels in use (K*[0] are S / PM themselves)
    const int32_t *QS[MAX_QLEV], *QP[MAX_QLEV];   // QS[j][i] = S[i << 3j], QP likewise for PM (8-ary levels)
    int32_t n8;
    const int32_t *WE, *WI;                       // arrays the walk / the emitter read in 16-item groups (E, I)
    int32_t mul;                                  // their group pitch in 16-int units: 1 (plain arrays; an interleaved
                                                  // [a x16 | b x16] layout, pitch 2, measured slower -- profiles/r01k)
    const int32_t *M[MAX_LEVELS];
    const int64_t *toff;
    const bxs::GridRec *G;       // direct-address grid records (all trees, see itree_search.cuh: search_walk_grid)
    const bxs::GridRec16 *G16;   // the same with the cell's starts packed in (search_walk_grid16)
    const unsigned char *dir;    // tree directory: toff[ntrees+1] (int64), padded to 16 B, then GridDir[ntrees]
    uint32_t dir_bytes, dir_grid_off;
    const int32_t *spS, *spPM;   // contiguous: spS[nsplit_pad] then spPM[nsplit_pad]
    uint32_t n;
    int32_t ntrees, nlev, nsplit, nsplit_pad, shift;   // stride = 1 << shift
};

struct bxg_itree {
    int64_t n = 0;
    int32_t ntrees = 0;
    bool built = false;
    int32_t *S = nullptr, *E = nullptr, *I = nullptr, *PM = nullptr;
    int32_t *M[MAX_LEVELS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int64_t mlen[MAX_LEVELS] = {0, 0, 0, 0, 0, 0};
    int nlev = 0;
    int64_t *toff = nullptr;
    bxs::GridRec *G = nullptr;            // grid records of all trees
    bxs::GridRec16 *G16 = nullptr;
    unsigned char *dir = nullptr;         // [toff | pad | GridDir x ntrees], one TMA bulk copy per CTA
    uint32_t dir_bytes = 0, dir_grid_off = 0;
    int64_t ncells_total = 0;
    int32_t *split = nullptr;
    int nsplit = 0, nsplit_pad = 0, shift = 0;
    int32_t *KS[MAX_KLEV] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // [0] aliases S
    int32_t *KP[MAX_KLEV] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // [0] aliases PM
    int nk = 1;
    int32_t *QS[MAX_QLEV] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // [0] aliases S
    int32_t *QP[MAX_QLEV] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // [0] aliases PM
    int n8 = 1;
    int32_t *es_end = nullptr, *es_k = nullptr;   // per-tree (end, in-order position) ordering for before(); lazy
    // query-side buffers (grow-only)
    int32_t *d_cnt = nullptr, *d_lo = nullptr, *d_hi = nullptr;
    unsigned long long *d_mask = nullptr;         // hit masks of the first four groups of every query (count -> fill)
    int64_t *d_off = nullptr;
    int32_t *d_hits = nullptr;
    int64_t q_cap = 0, hits_cap = 0;
    int64_t nq = -1, total = 0;
    // pipelined host path (bxg_itree_find_host): copy-in / copy-out streams, per-chunk events, pinned result buffers
    static constexpr int MAX_CHUNKS = 16;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[MAX_CHUNKS], ev_scan[MAX_CHUNKS], ev_fill[MAX_CHUNKS];
    int64_t *h_off = nullptr, *h_tot = nullptr;
    int32_t *h_hits = nullptr;
    int64_t h_off_cap = 0, h_hits_cap = 0;
    // single-pass find: tile-state words, per-chunk tickets and {end offset, overflow} results
    unsigned long long *d_tiles = nullptr;
    int64_t tiles_cap = 0;
    unsigned int *d_ticket = nullptr;       // MAX_CHUNKS + 1
    long long *d_result = nullptr;          // 2 x (MAX_CHUNKS + 1)
    long long *h_result = nullptr;          // pinned mirror
    // small-batch path (bxg_itree_find_small): results land in mapped pinned memory, written by the kernel itself
    long long *m_off = nullptr;             // SMALL_Q + 1 offsets, an overflow flag, the completion sequence number
    int32_t *m_hits = nullptr;              // SMALL_CAP hit ids
    long long *md_off = nullptr;            // their device addresses
    int32_t *md_hits = nullptr;
    long long m_seq = 0;
    int srv_slot = -1;                      // descriptor slot of the lingering find server (-1: none yet)
    bool srv_dirty = true;                  // the slot's descriptor does not describe the current arrays
    // the staged query arrays of the last find (device pointers valid until the next call)
    IndexView view() const {
        IndexView v;
        v.S = S; v.E = E; v.I = I; v.PM = PM;
        for (int l = 0; l < MAX_LEVELS; l++) v.M[l] = M[l];
        for (int j = 0; j < MAX_KLEV; j++) { v.KS[j] = KS[j]; v.KP[j] = KP[j]; }
        v.nk = nk;
        for (int j = 0; j < MAX_QLEV; j++) { v.QS[j] = QS[j]; v.QP[j] = QP[j]; }
        v.QS[0] = S; v.QP[0] = PM; v.n8 = n8;
        v.KS[0] = S; v.KP[0] = PM; v.WE = E; v.WI = I; v.mul = 1;
        v.toff = toff; v.spS = split; v.spPM = split + nsplit_pad;
        v.G = G; v.G16 = G16; v.dir = dir; v.dir_bytes = dir_bytes; v.dir_grid_off = dir_grid_off;
        v.n = (uint32_t)n; v.ntrees = ntrees; v.nlev = nlev; v.nsplit = nsplit; v.nsplit_pad = nsplit_pad; v.shift = shift;
        return v;
    }
};

// ------------------------------------------------------------------------------------------------------------------
// build kernels
// ------------------------------------------------------------------------------------------------------------------
// key = (start biased to unsigned) : (end > start) : (proper ? index : 2^31-1-index)   -- intersection.pyx:110-116
__global__ void k_make_keys(const int32_t *__restrict__ tree, const int32_t *__restrict__ start,
                            const int32_t *__restrict__ end, int64_t n, int32_t ntrees,
                            uint64_t *__restrict__ keys, int32_t *__restrict__ vals, int *bad) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int32_t s = start[i], e = end[i];
        uint64_t c = e > s ? 1u : 0u;
        uint64_t tb = c ? (uint64_t)i : (uint64_t)(0x7fffffff - i);
        keys[i] = ((uint64_t)((uint32_t)s ^ 0x80000000u) << 32) | (c << 31) | tb;
        vals[i] = (int32_t)i;
        if (tree && (tree[i] < 0 || tree[i] >= ntrees)) *bad = 1;
    }
}

__global__ void k_gather_tree(const int32_t *__restrict__ tree, const int32_t *__restrict__ vals, int64_t n,
                              uint32_t *__restrict__ tk) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) tk[i] = (uint32_t)tree[vals[i]];
}

// S,E in sorted order; TE = (tree : biased end) for the segmented running max
__global__ void k_gather_items(const int32_t *__restrict__ tree, const int32_t *__restrict__ start,
                               const int32_t *__restrict__ end, const int32_t *__restrict__ I, int64_t n,
                               int32_t *__restrict__ S, int32_t *__restrict__ E, uint64_t *__restrict__ TE) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        int32_t i = I[k];
        int32_t e = end[i];
        S[k] = start[i];
        E[k] = e;
        uint64_t t = tree ? (uint64_t)(uint32_t)tree[i] : 0ull;
        TE[k] = (t << 32) | (uint64_t)((uint32_t)e ^ 0x80000000u);
    }
}

// toff[t] = first sorted position whose tree id >= t   (TE is sorted by tree id in its high word)
__global__ void k_tree_offsets(const uint64_t *__restrict__ TE, int64_t n, int32_t ntrees, int64_t *__restrict__ toff) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntrees) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)(TE[mid] >> 32) < t) lo = mid + 1; else hi = mid;
    }
    toff[t] = lo;
}

__global__ void k_unpack_pm(const uint64_t *__restrict__ TEmax, int64_t n, int32_t *__restrict__ PM) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride)
        PM[k] = (int32_t)((uint32_t)TEmax[k] ^ 0x80000000u);
}

// one warp per 32-entry block: out[b] = max(in[32b .. 32b+31])
__global__ void k_block_max(const int32_t *__restrict__ in, int64_t n, int32_t *__restrict__ out, int64_t nout) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < nout; b += nwarps) {
        int64_t k = b * 32 + lane;
        int32_t v = k < n ? in[k] : INT32_MIN;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0) out[b] = v;
    }
}

// out[i] = A[i << ss] for i < nout, INT32_MAX padding up to nout_pad (sampled 16-ary search level)
__global__ void k_sample_level(const int32_t *__restrict__ A, int64_t n, int ss, int32_t *__restrict__ out,
                               int64_t nout, int64_t nout_pad) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nout_pad; i += stride)
        out[i] = (i < nout && (i << ss) < n) ? A[i << ss] : INT32_MAX;
}
__global__ void k_fill_i32(int32_t *p, int64_t n, int32_t v) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

__global__ void k_sample(const int32_t *__restrict__ S, const int32_t *__restrict__ PM, int64_t n, int shift, int nsplit,
                         int nsplit_pad, int32_t *__restrict__ split) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nsplit_pad) return;
    int64_t p = (int64_t)k << shift;
    split[k] = (k < nsplit && p < n) ? S[p] : INT32_MAX;
    split[nsplit_pad + k] = (k < nsplit && p < n) ? PM[p] : INT32_MAX;
}

// per tree: {smallest start, largest start} (S is sorted inside a tree)
__global__ void k_tree_extent(const int32_t *__restrict__ S, const int64_t *__restrict__ toff, int ntrees, int32_t *__restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntrees) return;
    const int64_t a = toff[t], b = toff[t + 1];
    out[2 * t] = a < b ? S[a] : 0;
    out[2 * t + 1] = a < b ? S[b - 1] : 0;
}

// grid records: one thread per record (tree t, cell c in 0..ncells): x = lower_bound(S, cell_start), y = upper_bound(PM, cell_start)
// inside the tree's segment; cell_start = base + (c << shift) in 64 bits (the sentinel cell may lie beyond INT32_MAX)
__global__ void k_build_grid(const int32_t *__restrict__ S, const int32_t *__restrict__ PM, const int64_t *__restrict__ toff,
                             const bxs::GridDir *__restrict__ gd, int ntrees, int64_t nrec, bxs::GridRec *__restrict__ G) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrec; r += stride) {
        int lo = 0, hi = ntrees - 1;                     // tree of record r: largest t with coff[t] <= r
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((int64_t)gd[mid].coff <= r) lo = mid; else hi = mid - 1;
        }
        const bxs::GridDir d = gd[lo];
        const int64_t c = r - d.coff;
        const int64_t v = (int64_t)d.base + (c << d.shift);
        const int64_t a = toff[lo], b = toff[lo + 1];
        int64_t l = a, h = b;
        while (l < h) {                                  // first k with S[k] >= v
            const int64_t m = (l + h) >> 1;
            if ((int64_t)S[m] < v) l = m + 1; else h = m;
        }
        const uint32_t x = (uint32_t)l;
        l = a;
        h = b;
        while (l < h) {                                  // first k with PM[k] > v
            const int64_t m = (l + h) >> 1;
            if ((int64_t)PM[m] <= v) l = m + 1; else h = m;
        }
        G[r] = bxs::GridRec{x, (uint32_t)l};
    }
}

// 16-byte records: (x, y) of G plus the cell's item count and the offsets of its first 56 / shift items (search_walk_grid16)
__global__ void k_build_grid16(const int32_t *__restrict__ S, const bxs::GridRec *__restrict__ G, const bxs::GridDir *__restrict__ gd,
                               int ntrees, int64_t nrec, bxs::GridRec16 *__restrict__ G16) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrec; r += stride) {
        int lo = 0, hi = ntrees - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((int64_t)gd[mid].coff <= r) lo = mid; else hi = mid - 1;
        }
        const bxs::GridDir d = gd[lo];
        const int64_t c = r - d.coff;
        const bxs::GridRec g = G[r];
        unsigned long long p = 0;
        if (c < (int64_t)d.ncells) {
            const uint32_t cnt = G[r + 1].x - g.x;
            p = cnt < 255u ? cnt : 255u;
            const uint32_t kmax = (uint32_t)bxs::grid16_fields(d.shift), stored = cnt < kmax ? cnt : kmax;
            const int64_t v = (int64_t)d.base + (c << d.shift);
            for (uint32_t i = 0; i < stored; i++)
                p |= (unsigned long long)((int64_t)S[g.x + i] - v) << (8 + i * d.shift);
        }
        G16[r] = bxs::GridRec16{g.x, g.y, (uint32_t)p, (uint32_t)(p >> 32)};
    }
}

// ------------------------------------------------------------------------------------------------------------------
// 1-D TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP + SYNCS)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// find
// ------------------------------------------------------------------------------------------------------------------
struct SmemIndex {
    const int32_t *spS, *spPM;   // shared
    const int64_t *toff;         // shared (ntrees <= SMEM_TREES) or global
    const bxs::GridDir *gdir;    // likewise (PROBE 3 only)
};

// The search (dual_search) and walk (walk_hits) arithmetic lives in itree_search.cuh so that the CPU fuzz harness
// (tests/search_fuzz.cpp) compiles exactly the code the kernels run.
// One aligned 64-byte group per thread as two 256-bit loads (LDG.E.256, new with sm_100).  The searches are divergent --
// every lane reads its own line -- and the kernel is bound by L1 data-pipe wavefronts, one per lane-sector whatever the
// access width: 4 x LDG.128 cost 4 wavefronts per lane and group (each 32-byte sector visited twice), 2 x LDG.256 cost 2.
struct Ld4 {
    __device__ __forceinline__ void operator()(const int4 *p, int4 &a, int4 &b, int4 &c, int4 &d) const {
        asm("ld.global.nc.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
        asm("ld.global.nc.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w), "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w) : "l"(p + 2));
    }
};
struct Ld1 {
    __device__ __forceinline__ int32_t operator()(const int32_t *p) const { return __ldg(p); }
};
// one aligned 32-byte sector (8 entries): a single LDG.E.256
struct Ld8 {
    __device__ __forceinline__ void operator()(const int4 *p, int4 &a, int4 &b) const {
        asm("ld.global.nc.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
    }
};
// L2 residency.  The count pass streams 12 B/query in and 20 B/query out (320 MB per 10 M queries) past an index of
// ~85 MB (S, E, the sampled levels) that every query reads at random; with default caching the streams keep pushing
// index lines out of the 126 MB L2 and the kernel re-reads them from HBM (r01q: 854 MB of DRAM reads against ~205 MB
// of distinct data).  So index loads carry an evict_last policy, and the per-query streams use .cs (evict-first)
// accesses.  -DFIND_L2_HINTS=0 restores plain accesses for A/B runs.
#ifndef FIND_L2_HINTS
#define FIND_L2_HINTS 1
#endif
#if FIND_L2_HINTS
__device__ __forceinline__ uint64_t l2_keep_policy() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
struct LdK4 {
    uint64_t pol;
    __device__ __forceinline__ LdK4() : pol(l2_keep_policy()) {}
    __device__ __forceinline__ void operator()(const int4 *p, int4 &a, int4 &b, int4 &c, int4 &d) const {
        asm("ld.global.nc.L2::cache_hint.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
            : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p), "l"(pol));
        asm("ld.global.nc.L2::cache_hint.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
            : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w), "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w) : "l"(p + 2), "l"(pol));
    }
};
// L1: the top sampled level (QS[n8-1], ~80 KB for 10 M items) is read by every query -- LdTop8 asks L1 to keep it
// (evict_last) while every other index sector is touched once per query at random and would only push it out, so
// LdK8 does not allocate in L1 (-DFIND_L1_POLICY=0: default L1 behaviour for both).
#ifndef FIND_L1_POLICY
#define FIND_L1_POLICY 1
#endif
struct LdK8 {
    uint64_t pol;
    __device__ __forceinline__ LdK8() : pol(l2_keep_policy()) {}
    __device__ __forceinline__ void operator()(const int4 *p, int4 &a, int4 &b) const {
#if FIND_L1_POLICY
        asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
#else
        asm("ld.global.nc.L2::cache_hint.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
#endif
            : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p), "l"(pol));
    }
};
struct LdTop8 {
    uint64_t pol;
    __device__ __forceinline__ LdTop8() : pol(l2_keep_policy()) {}
    __device__ __forceinline__ void operator()(const int4 *p, int4 &a, int4 &b) const {
#if FIND_L1_POLICY
        asm("ld.global.nc.L1::evict_last.L2::cache_hint.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
#else
        asm("ld.global.nc.L2::cache_hint.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
#endif
            : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p), "l"(pol));
    }
};
__device__ __forceinline__ int32_t ld_stream(const int32_t *p) { return __ldcs(p); }
__device__ __forceinline__ unsigned long long ld_stream(const unsigned long long *p) { return __ldcs(p); }
__device__ __forceinline__ long long ld_stream(const long long *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream_q(int32_t *p, int32_t v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream_q(unsigned long long *p, unsigned long long v) { __stcs(p, v); }
#else
typedef Ld4 LdK4;
typedef Ld8 LdK8;
typedef Ld8 LdTop8;
__device__ __forceinline__ int32_t ld_stream(const int32_t *p) { return __ldg(p); }
__device__ __forceinline__ unsigned long long ld_stream(const unsigned long long *p) { return *p; }
__device__ __forceinline__ long long ld_stream(const long long *p) { return *p; }
__device__ __forceinline__ void st_stream_q(int32_t *p, int32_t v) { *p = v; }
__device__ __forceinline__ void st_stream_q(unsigned long long *p, unsigned long long v) { *p = v; }
#endif

// PROBE selects the search: 2 = search_walk_probe8 (one search over 8-ary levels + backward probe of QP[1], half-group
// walk: ~8 sectors per query), 1 = search_walk_probe (the same over the 16-ary levels: ~10), 0 = search_walk (two
// lock-step searches: ~13); identical results, kept switchable (BXB200_FIND_PROBE) for A/B runs
// one 8-byte grid record; the records of neighbouring cells share sectors (the qs and qe cells of a BED-sized query are
// 0-2 cells apart), so they may stay in L1; L2 keeps them with the rest of the index
#if FIND_L2_HINTS
struct LdRec {
    uint64_t pol;
    __device__ __forceinline__ LdRec() : pol(l2_keep_policy()) {}
    __device__ __forceinline__ bxs::GridRec operator()(const bxs::GridRec *p) const {
        bxs::GridRec r;
        asm("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(pol));
        return r;
    }
};
struct LdRec16 {
    uint64_t pol;
    __device__ __forceinline__ LdRec16() : pol(l2_keep_policy()) {}
    __device__ __forceinline__ bxs::GridRec16 operator()(const bxs::GridRec16 *p) const {
        bxs::GridRec16 r;
        asm("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.p0), "=r"(r.p1) : "l"(p), "l"(pol));
        return r;
    }
};
#else
struct LdRec {
    __device__ __forceinline__ bxs::GridRec operator()(const bxs::GridRec *p) const {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
        return bxs::GridRec{v.x, v.y};
    }
};
struct LdRec16 {
    __device__ __forceinline__ bxs::GridRec16 operator()(const bxs::GridRec16 *p) const {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
        return bxs::GridRec16{v.x, v.y, v.z, v.w};
    }
};
#endif

template <int PROBE, typename SP, typename F>
__device__ __forceinline__ void query_search_walk(const IndexView &ix, const SP &spS, const SP &spPM, const bxs::GridDir *gd,
                                                  uint32_t seg_lo, uint32_t seg_hi, int32_t qe, int32_t qs, uint32_t &hi,
                                                  uint32_t &lo, F &&f) {
    if (PROBE == 4)
        bxs::search_walk_grid16(ix.G16, *gd, ix.S, seg_lo, seg_hi, qe, qs, ix.E, ix.M, ix.nlev, LdRec16(), LdK8(), Ld1(), hi, lo, f);
    else if (PROBE == 3)
        bxs::search_walk_grid(ix.G, *gd, ix.S, seg_lo, seg_hi, qe, qs, ix.E, ix.M, ix.nlev, LdRec(), LdK8(), Ld1(), hi, lo, f);
    else if (PROBE == 2)
        bxs::search_walk_probe8(ix.QS, ix.QP, ix.n8, spS, spPM, ix.shift, seg_lo, seg_hi, qe, qs, ix.E, ix.M, ix.nlev, LdTop8(),
                                LdK8(), Ld1(), hi, lo, f);
    else if (PROBE == 1)
        bxs::search_walk_probe(ix.KS, ix.KP, ix.nk, spS, spPM, ix.shift, seg_lo, seg_hi, qe, qs, ix.E, ix.M, ix.nlev, LdK4(),
                               LdK8(), Ld1(), hi, lo, f);
    else
        bxs::search_walk(ix.KS, ix.KP, ix.nk, spS, spPM, ix.shift, seg_lo, seg_hi, qe, qs, ix.E, ix.M, ix.nlev, LdK4(), Ld1(),
                         hi, lo, f);
}
__device__ __forceinline__ const SmemIndex stage_index(const IndexView &ix, unsigned char *smem_raw) {
    // layout: [mbarrier 8 B][pad 8 B][spS nsplit_pad x 4][spPM nsplit_pad x 4][toff (ntrees+1) x 8]
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    int32_t *sp = reinterpret_cast<int32_t *>(smem_raw + 16);
    int64_t *stoff = reinterpret_cast<int64_t *>(smem_raw + 16 + (size_t)ix.nsplit_pad * 8);
    const uint32_t bytes = (uint32_t)ix.nsplit_pad * 8u;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        if (bytes) {
            mbar_expect_tx(bar, bytes);
            tma_load_1d(sp, ix.spS, bytes, bar);
        }
    }
    const bool toff_smem = ix.ntrees <= SMEM_TREES;
    if (toff_smem)
        for (int t = threadIdx.x; t <= ix.ntrees; t += blockDim.x) stoff[t] = ix.toff[t];
    __syncthreads();                       // barrier init + toff visible to every thread
    if (bytes) mbar_wait(bar, 0);
    SmemIndex s;
    s.spS = sp;
    s.spPM = sp + ix.nsplit_pad;
    s.toff = toff_smem ? stoff : ix.toff;
    s.gdir = nullptr;
    return s;
}

// PROBE 3 (direct addressing) needs no splitter table: the only per-CTA state is the tree directory -- segment offsets and
// grid geometry of every tree, 24 bytes per chromosome -- staged by ONE 1-D TMA bulk copy (it is laid out for that:
// [toff | pad to 16 B | GridDir x ntrees], csrc/itree.cu build).  Forests with more than SMEM_TREES trees read it from L2.
__device__ __forceinline__ const SmemIndex stage_dir(const IndexView &ix, unsigned char *smem_raw) {
    SmemIndex s;
    s.spS = s.spPM = nullptr;
    if (ix.ntrees > SMEM_TREES || ix.dir_bytes == 0) {
        s.toff = ix.toff;
        s.gdir = reinterpret_cast<const bxs::GridDir *>(ix.dir + ix.dir_grid_off);
        return s;
    }
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    unsigned char *dst = smem_raw + 16;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, ix.dir_bytes);
        tma_load_1d(dst, ix.dir, ix.dir_bytes, bar);
    }
    __syncthreads();                       // barrier init visible to every thread
    mbar_wait(bar, 0);
    s.toff = reinterpret_cast<const int64_t *>(dst);
    s.gdir = reinterpret_cast<const bxs::GridDir *>(dst + ix.dir_grid_off);
    return s;
}

// The count pass stashes the hit masks of four consecutive 16-item groups, starting at the first group that has a hit
// (64 bits per query), and records that group's position in lo_: the fill pass then emits straight from the masks -- it
// reads the aligned groups of I, but never E, qs or hi again.  Queries whose hits span more than four groups set
// WALK_AGAIN in lo_ and are re-walked by the fill from that position (everything in front of the first hit has E <= qs).
constexpr uint32_t WALK_AGAIN = 0x80000000u;

struct MaskStash {
    uint32_t base = 0;        // 16-aligned position of the first group that has a hit (valid once c > 0)
    unsigned long long m = 0;
    bool overflow = false;
    int32_t c = 0;
    __device__ __forceinline__ void operator()(uint32_t k0, unsigned mask) {
        if (c == 0) base = k0;
        c += __popc(mask);
        const uint32_t j = (k0 - base) >> 4;
        if (j < 4u) m |= (unsigned long long)mask << (16u * j); else overflow = true;
    }
};

template <bool FILL, int PROBE>
__global__ void __launch_bounds__(FIND_THREADS, FIND_MIN_CTAS)
k_find(const __grid_constant__ IndexView ix, const int32_t *__restrict__ qtree, const int32_t *__restrict__ qs_, const int32_t *__restrict__ qe_,
       int64_t nq, int32_t *__restrict__ cnt, int32_t *__restrict__ lo_, int32_t *__restrict__ hi_,
       unsigned long long *__restrict__ mask_, const int64_t *__restrict__ off, int32_t *__restrict__ hits,
       unsigned long long *total) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemIndex sm;
    if (!FILL) sm = PROBE >= 3 ? stage_dir(ix, smem_raw) : stage_index(ix, smem_raw);
    unsigned long long local = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // Fill only: the inputs of the next grid-stride iteration are loaded at the top of the current one (0.350 -> 0.338 ms).
    // The same pipelining of qs/qe in the count kernel, by registers or by prefetch.global.L1, measured no gain.
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int32_t n_lo = 0;
    unsigned long long n_m = 0;
    int64_t n_off = 0;
    if (FILL && q < nq) { n_lo = lo_[q]; n_m = mask_[q]; n_off = off[q]; }
    for (; q < nq; q += stride) {
        const int32_t c_a = n_lo;
        const unsigned long long c_m = n_m;
        const int64_t c_off = n_off;
        if (FILL && q + stride < nq) { n_lo = lo_[q + stride]; n_m = mask_[q + stride]; n_off = off[q + stride]; }
        if (!FILL) {
            const int32_t qs = ld_stream(qs_ + q), qe = ld_stream(qe_ + q), t = qtree ? ld_stream(qtree + q) : 0;
            uint32_t lo = 0, hi = 0;
            MaskStash st;
            if (t >= 0 && t < ix.ntrees) {
                const uint32_t seg_lo = (uint32_t)sm.toff[t], seg_hi = (uint32_t)sm.toff[t + 1];
                // hi: candidates (start < qe) end here; lo: coarse start of the walk (running max end > qs from here on)
                query_search_walk<PROBE>(ix, sm.spS, sm.spPM, PROBE >= 3 ? sm.gdir + t : nullptr, seg_lo, seg_hi, qe, qs, hi, lo, st);
            }
            cnt[q] = st.c;                                 // read again right away by the scan: keep it cached
            st_stream_q(lo_ + q, (int32_t)((st.c ? st.base : lo) | (st.overflow ? WALK_AGAIN : 0u)));
            st_stream_q(hi_ + q, (int32_t)hi);
            st_stream_q(mask_ + q, st.m);
            local += (unsigned long long)st.c;
        } else {
            const uint32_t lo_raw = (uint32_t)c_a;
            int32_t *dst = hits + c_off;
            if (!(lo_raw & WALK_AGAIN)) {
                unsigned long long m = c_m;
                uint32_t k0 = lo_raw;                          // 16-aligned position of the first group with a hit
                while (m) {                                    // at most four groups
                    const unsigned mk = (unsigned)(m & 0xffffull);
                    if (mk) dst = bxs::emit_group(ix.WI, k0, mk, dst, Ld4(), ix.mul);
                    m >>= 16;
                    k0 += 16;
                }
            } else {
                const uint32_t lo = lo_raw & ~WALK_AGAIN, hi = (uint32_t)hi_[q];
                const int32_t qs = __ldg(qs_ + q);
                bxs::walk_hits(ix.WE, ix.M, ix.nlev, lo, hi, qs, Ld4(), Ld1(),
                               [&](uint32_t k0, unsigned mask) { dst = bxs::emit_group(ix.WI, k0, mask, dst, Ld4(), ix.mul); },
                               bxs::NoPrefetch(), ix.mul);
            }
        }
    }
    if (!FILL && total) {
        // warp-aggregated: one atomic per warp
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if ((threadIdx.x & 31) == 0 && local) atomicAdd(total, local);
    }
}

// ---- fill, staged: the hits of the 32 consecutive queries of one warp form ONE contiguous span of the output
// (off[] is monotone), on average 32 x hits/query ints.  Each lane drops its hits into a per-warp shared-memory window
// at (off[q] - off[first query of the warp]) and the warp then writes the window out with coalesced stores: the
// direct fill's ~27 scattered 4-byte store instructions per warp (7 partial sectors each -- LG-throttle stalls and
// L2 tag traffic) become ~27 shared stores plus span/32 full-line stores.  Warps whose span exceeds the window
// (FILL_STAGE ints) write directly, as k_find<true> does.
constexpr int FILL_STAGE = 512;
#ifndef FILL_MIN_CTAS
#define FILL_MIN_CTAS 6            // 4 / 5 / 6 CTAs per SM -> 0.208 / 0.202 / 0.198 ms (profiles/r01t)
#endif

struct SharedSink {
    uint32_t a;                                           // shared-window byte address
    __device__ __forceinline__ void put(int32_t v) {
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
        a += 4;
    }
};

template <typename SINK>
__device__ __forceinline__ void fill_query(const IndexView &ix, uint32_t lo_raw, unsigned long long m, int64_t q,
                                           const int32_t *__restrict__ qs_, const int32_t *__restrict__ hi_, SINK &out) {
    if (!(lo_raw & WALK_AGAIN)) {
        uint32_t k0 = lo_raw;                                  // 16-aligned position of the first group with a hit
        while (m) {                                            // at most four groups
            const unsigned mk = (unsigned)(m & 0xffffull);
            if (mk) bxs::emit_group_halves_to(ix.WI, k0, mk, out, Ld8());
            m >>= 16;
            k0 += 16;
        }
    } else {
        const uint32_t lo = lo_raw & ~WALK_AGAIN, hi = (uint32_t)hi_[q];
        const int32_t qs = __ldg(qs_ + q);
        bxs::walk_hits(ix.WE, ix.M, ix.nlev, lo, hi, qs, Ld4(), Ld1(),
                       [&](uint32_t k0, unsigned mask) { bxs::emit_group_to(ix.WI, k0, mask, out, Ld4(), ix.mul); },
                       bxs::NoPrefetch(), ix.mul);
    }
}

__global__ void __launch_bounds__(FIND_THREADS, FILL_MIN_CTAS)
k_fill_staged(const __grid_constant__ IndexView ix, const int32_t *__restrict__ qs_, int64_t nq, const int32_t *__restrict__ lo_,
              const int32_t *__restrict__ hi_, const unsigned long long *__restrict__ mask_, const int64_t *__restrict__ off,
              int32_t *__restrict__ hits, int64_t hits_cap) {
    __shared__ int32_t stage[FIND_THREADS / 32][FILL_STAGE];
    const int lane = threadIdx.x & 31;
    int32_t *buf = stage[threadIdx.x >> 5];
    const uint32_t buf_a = (uint32_t)__cvta_generic_to_shared(buf);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t off_end = off[nq];
    // launched speculatively, before the host knows the total: if the hits do not fit the buffer, do nothing -- the host
    // sees the total a moment later, grows the buffer and launches the fill again
    if (off_end > hits_cap) return;
    for (int64_t qw = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); qw < nq; qw += stride) {   // warp-uniform
        const int64_t q = qw + lane;
        const bool valid = q < nq;
        uint32_t lo_raw = 0;
        unsigned long long m = 0;
        int64_t o = off_end;
        if (valid) { lo_raw = (uint32_t)ld_stream(lo_ + q); m = ld_stream(mask_ + q); o = ld_stream((const long long *)off + q); }
        int64_t o_next = __shfl_down_sync(0xffffffffu, o, 1);
        if (lane == 31) o_next = valid ? off[q + 1] : off_end;
        const int64_t span0 = __shfl_sync(0xffffffffu, o, 0);
        const int64_t span = __shfl_sync(0xffffffffu, o_next, 31) - span0;
        if (span == 0) continue;
        if (span <= FILL_STAGE) {
            if (o_next > o) {
                SharedSink out{buf_a + 4u * (uint32_t)(o - span0)};
                fill_query(ix, lo_raw, m, q, qs_, hi_, out);
            }
            __syncwarp();
            int32_t *dst = hits + span0;
            for (int i = lane; i < (int)span; i += 32) st_stream_q(dst + i, buf[i]);
            __syncwarp();
        } else if (o_next > o) {
            bxs::PtrSink out{hits + o};
            fill_query(ix, lo_raw, m, q, qs_, hi_, out);
        }
    }
}

struct CastI64 {
    __device__ __forceinline__ int64_t operator()(int32_t v) const { return (int64_t)v; }
};

// ------------------------------------------------------------------------------------------------------------------
// Single-pass find: count, CSR offsets and fill in ONE launch.
//
// Tiles of 256 queries are handed out in order by an atomic ticket to a persistent grid.  A tile counts its hits,
// block-scans them, publishes its total in a 64-bit tile-state word (2-bit flag + 62-bit value), obtains its global
// base with a warp-parallel decoupled look-back over its predecessors' words (as in single-pass prefix scans), writes
// offsets[q] and immediately re-walks its queries to emit the hits while their E / I lines are still in L1.
// Compared with count + cub scan + fill this drops two launches, the cnt/lo/hi round trip (24 B/query written and read)
// and the second DRAM fetch of every candidate window.  Hit order inside a query is the walk order, so the lists are
// the same ordered lists; no atomics touch the output.
// If the hit buffer is too small the kernel still finishes the scan (offsets and total are exact), raises result[1]
// and the host re-runs it with a larger buffer.
// ------------------------------------------------------------------------------------------------------------------
constexpr int FUSED_THREADS = 256;
constexpr unsigned long long TS_PARTIAL = 1ull << 62, TS_INCLUSIVE = 2ull << 62, TS_VALUE = (1ull << 62) - 1ull;

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// hits of one query into `out`, from the masks the count walk kept in registers (E is not read again; only the halves
// of the I groups that hold a hit are loaded), or by re-walking when the walk was too long for the stash
template <typename SINK>
__device__ __forceinline__ void emit_query(const IndexView &ix, const MaskStash &st, uint32_t hi, int32_t qs, SINK &out) {
    if (!st.overflow) {
        unsigned long long m = st.m;
        uint32_t k0 = st.base;
        while (m) {
            const unsigned mk = (unsigned)(m & 0xffffull);
            if (mk) bxs::emit_group_halves_to(ix.WI, k0, mk, out, Ld8());
            m >>= 16;
            k0 += 16;
        }
    } else {
        bxs::walk_hits_halves(ix.WE, ix.M, ix.nlev, st.base, hi, qs, Ld8(), Ld1(),
                              [&](uint32_t k0, unsigned mask) { bxs::emit_group_halves_to(ix.WI, k0, mask, out, Ld8()); });
    }
}

template <int PROBE>
__global__ void __launch_bounds__(FUSED_THREADS, FIND_MIN_CTAS)
k_find_fused(const __grid_constant__ IndexView ix, const int32_t *__restrict__ qtree, const int32_t *__restrict__ qs_,
             const int32_t *__restrict__ qe_, int64_t nq, int64_t *__restrict__ off, int32_t *__restrict__ hits,
             int64_t hits_cap, const int64_t *__restrict__ base_ptr, unsigned long long *__restrict__ tile_state,
             unsigned int *__restrict__ ticket, long long *__restrict__ result) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef cub::BlockScan<long long, FUSED_THREADS> BS;
    __shared__ typename BS::TempStorage scan_tmp;
    __shared__ unsigned int s_tile;
    __shared__ long long s_base;
    const SmemIndex sm = PROBE >= 3 ? stage_dir(ix, smem_raw) : stage_index(ix, smem_raw);
    const unsigned int ntiles = (unsigned int)((nq + FUSED_THREADS - 1) / FUSED_THREADS);
    const int lane = threadIdx.x & 31;
    while (true) {
        if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const unsigned int tile = s_tile;
        if (tile >= ntiles) break;
        const int64_t q = (int64_t)tile * FUSED_THREADS + threadIdx.x;
        uint32_t lo = 0, hi = 0;
        int32_t qs = 0;
        long long c = 0;
        MaskStash st;
        if (q < nq) {
            qs = __ldg(qs_ + q);
            const int32_t qe = __ldg(qe_ + q);
            const int32_t t = qtree ? __ldg(qtree + q) : 0;
            if (t >= 0 && t < ix.ntrees) {
                const uint32_t seg_lo = (uint32_t)sm.toff[t], seg_hi = (uint32_t)sm.toff[t + 1];
                query_search_walk<PROBE>(ix, sm.spS, sm.spPM, PROBE >= 3 ? sm.gdir + t : nullptr, seg_lo, seg_hi, qe, qs, hi, lo, st);
                c = st.c;
            }
        }
        long long excl, tile_total;
        BS(scan_tmp).ExclusiveSum(c, excl, tile_total);
        if (threadIdx.x < 32) {
            if (lane == 0) atomicExch(tile_state + tile, TS_PARTIAL | (unsigned long long)tile_total);
            // decoupled look-back: 32 predecessors per step; tile -1 is a virtual INCLUSIVE holding the chunk base
            unsigned long long base = 0;
            long long p0 = (long long)tile - 1;
            while (true) {
                const long long p = p0 - lane;
                unsigned long long s;
                if (p >= 0) {
                    do { s = ld_volatile_u64(tile_state + p); } while ((s >> 62) == 0);
                } else {
                    s = TS_INCLUSIVE | (p == -1 ? (unsigned long long)*base_ptr : 0ull);
                }
                const unsigned incl = __ballot_sync(0xffffffffu, (s >> 62) == 2);
                const int first = __ffs((int)incl) - 1;
                unsigned long long v = (first < 0 || lane <= first) ? (s & TS_VALUE) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                base += v;
                if (incl) break;
                p0 -= 32;
            }
            if (lane == 0) {
                atomicExch(tile_state + tile, TS_INCLUSIVE | (base + (unsigned long long)tile_total));
                s_base = (long long)base;
            }
        }
        __syncthreads();
        const long long base = s_base;
        const bool fits = base + tile_total <= hits_cap;
        if (q < nq) off[q] = base + excl;
        if (fits && c > 0) {
            // direct stores: staging the warp's span through shared memory, as k_fill_staged does, was measured SLOWER
            // here (single-pass 0.81 -> 1.05 ms per 10 M queries, profiles/r01g2): the tile loop is already barrier-bound
            bxs::PtrSink out{hits + base + excl};
            emit_query(ix, st, hi, qs, out);
        }
        if (threadIdx.x == 0) {
            if (!fits) result[1] = 1;
            if (tile == ntiles - 1) {
                off[nq] = base + tile_total;
                result[0] = base + tile_total;
            }
        }
        __syncthreads();      // s_tile / s_base / scan_tmp are reused by the next tile
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
// the lingering find server (see k_find_server below)
struct FindDesc {
    IndexView ix;
    int32_t *out_hits;
};
constexpr int SRV_SLOTS = 256;
constexpr uint32_t SRV_KILL = 0xffffffffu, SRV_OVERFLOW = 0xffffffffu;

struct FindServer {
    FindDesc *d_table = nullptr;
    unsigned long long *m = nullptr, *md = nullptr;   // the six mailbox words (host / device address)
    cudaStream_t stream = nullptr;
    uint32_t seq = 0;
    unsigned long long gen = 0;                        // generation of the last launch (0: never launched)
    bxg_itree *owner[SRV_SLOTS] = {};
    int enabled = -1;
    unsigned idle_us = 100;
    long long launches = 0, requests = 0;
};
static FindServer g_srv;

static void server_stop();

static void free_index(bxg_itree *t) {
    server_stop();                          // a lingering find server may hold descriptors of these arrays
    t->srv_dirty = true;
    cudaFree(t->S); cudaFree(t->E); cudaFree(t->I); cudaFree(t->PM); cudaFree(t->toff); cudaFree(t->split);
    cudaFree(t->G); cudaFree(t->G16); cudaFree(t->dir);
    t->G = nullptr; t->G16 = nullptr; t->dir = nullptr; t->dir_bytes = t->dir_grid_off = 0; t->ncells_total = 0;
    for (int l = 0; l < MAX_LEVELS; l++) { cudaFree(t->M[l]); t->M[l] = nullptr; t->mlen[l] = 0; }
    for (int j = 1; j < MAX_KLEV; j++) { cudaFree(t->KS[j]); cudaFree(t->KP[j]); }
    for (int j = 0; j < MAX_KLEV; j++) t->KS[j] = t->KP[j] = nullptr;
    t->nk = 1;
    for (int j = 1; j < MAX_QLEV; j++) { cudaFree(t->QS[j]); cudaFree(t->QP[j]); }
    for (int j = 0; j < MAX_QLEV; j++) t->QS[j] = t->QP[j] = nullptr;
    t->n8 = 1;
    cudaFree(t->es_end);
    cudaFree(t->es_k);
    t->es_end = t->es_k = nullptr;
    t->S = t->E = t->I = t->PM = nullptr;
    t->toff = nullptr;
    t->split = nullptr;
    t->nlev = 0;
    t->built = false;
}

static size_t find_smem_bytes(const bxg_itree *t, int probe) {
    if (probe >= 3) return 16 + (t->ntrees <= SMEM_TREES ? (size_t)t->dir_bytes : 0);
    return 16 + (size_t)t->nsplit_pad * 8 + (t->ntrees <= SMEM_TREES ? (size_t)(t->ntrees + 1) * 8 : 0);
}

static int ensure_query_buffers(bxg_itree *t, int64_t nq) {
    if (nq + 1 <= t->q_cap) return BXG_OK;
    BXG_CUDA(cudaStreamSynchronize(ctx().stream));
    cudaFree(t->d_cnt); cudaFree(t->d_lo); cudaFree(t->d_hi); cudaFree(t->d_off); cudaFree(t->d_mask);
    t->d_cnt = t->d_lo = t->d_hi = nullptr;
    t->d_off = nullptr;
    t->d_mask = nullptr;
    t->q_cap = 0;                    // a failed allocation below must not leave a stale capacity over null buffers
    int64_t cap = nq + 1 + nq / 8;
    BXG_CUDA(cudaMalloc(&t->d_cnt, (size_t)cap * 4));
    BXG_CUDA(cudaMalloc(&t->d_lo, (size_t)cap * 4));
    BXG_CUDA(cudaMalloc(&t->d_hi, (size_t)cap * 4));
    BXG_CUDA(cudaMalloc(&t->d_off, (size_t)cap * 8));
    BXG_CUDA(cudaMalloc(&t->d_mask, (size_t)cap * 8));
    t->q_cap = cap;
    return BXG_OK;
}

// 4 (default): direct-address grid with packed 16-byte records; 3: 8-byte records + S; 2: 8-ary levels + probe;
// 1: 16-ary levels + probe; 0: two lock-step searches
static int find_probe() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("BXB200_FIND_PROBE");
        v = (e && e[0] >= '0' && e[0] <= '4') ? e[0] - '0' : 4;
    }
    return v;
}

// pass A over queries [q0, q0+nq): searches + per-query hit counts into d_cnt/d_lo/d_hi[q0..]
template <int PROBE>
static int launch_count_as(bxg_itree *t, const int32_t *dqt, const int32_t *dqs, const int32_t *dqe, int64_t nq,
                           unsigned long long *d_total, int64_t q0) {
    size_t smem = find_smem_bytes(t, PROBE);
    static bool attr_set = false;
    if (!attr_set) {
        BXG_CUDA(cudaFuncSetAttribute(k_find<false, PROBE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_set = true;
    }
    // persistent grid: exactly the CTAs that are co-resident (whole waves only), grid-stride over the queries
    int occ = 0;
    BXG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_find<false, PROBE>, FIND_THREADS, smem));
    int grid = grid_for(cdiv(nq, FIND_THREADS), occ > 0 ? occ : 1);
    BXG_LAUNCH((k_find<false, PROBE>), grid, FIND_THREADS, smem, t->view(), dqt ? dqt + q0 : nullptr, dqs + q0, dqe + q0, nq,
               t->d_cnt + q0, t->d_lo + q0, t->d_hi + q0, t->d_mask + q0, (const int64_t *)nullptr, (int32_t *)nullptr,
               d_total);
    return BXG_OK;
}

static int launch_count(bxg_itree *t, const int32_t *dqt, const int32_t *dqs, const int32_t *dqe, int64_t nq,
                        unsigned long long *d_total, int64_t q0 = 0) {
    switch (find_probe()) {
        case 0: return launch_count_as<0>(t, dqt, dqs, dqe, nq, d_total, q0);
        case 1: return launch_count_as<1>(t, dqt, dqs, dqe, nq, d_total, q0);
        case 2: return launch_count_as<2>(t, dqt, dqs, dqe, nq, d_total, q0);
        case 3: return launch_count_as<3>(t, dqt, dqs, dqe, nq, d_total, q0);
        default: return launch_count_as<4>(t, dqt, dqs, dqe, nq, d_total, q0);
    }
}

// pass B over queries [q0, q0+nq): writes hits at the (global) CSR offsets d_off[q0..]
static bool fill_is_staged() {
    static const bool staged = [] {
        const char *e = getenv("BXB200_FILL_STAGED");     // 0: direct stores (k_find<true>), for A/B measurements
        return !(e && e[0] == '0');
    }();
    return staged;
}

static int launch_fill(bxg_itree *t, const int32_t *dqs, int64_t nq, int64_t q0 = 0) {
    int occ = 0;
    const bool staged = fill_is_staged();
    if (staged) {
        BXG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fill_staged, FIND_THREADS, 0));
        int grid = grid_for(cdiv(nq, FIND_THREADS), occ > 0 ? occ : 1);
        BXG_LAUNCH(k_fill_staged, grid, FIND_THREADS, 0, t->view(), dqs + q0, nq, (const int32_t *)(t->d_lo + q0),
                   (const int32_t *)(t->d_hi + q0), (const unsigned long long *)(t->d_mask + q0),
                   (const int64_t *)(t->d_off + q0), t->d_hits, t->hits_cap);
        return BXG_OK;
    }
    BXG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_find<true, 0>, FIND_THREADS, 0));
    int grid = grid_for(cdiv(nq, FIND_THREADS), occ > 0 ? occ : 1);
    BXG_LAUNCH((k_find<true, 0>), grid, FIND_THREADS, 0, t->view(), (const int32_t *)nullptr, dqs + q0,
               (const int32_t *)nullptr, nq, t->d_cnt + q0, t->d_lo + q0, t->d_hi + q0, t->d_mask + q0,
               (const int64_t *)(t->d_off + q0), t->d_hits, (unsigned long long *)nullptr);
    return BXG_OK;
}

// counts of one chunk followed by a single 0, so that an exclusive scan also emits the chunk's end offset
struct ChunkCount {
    const int32_t *cnt;
    int64_t n;
    __device__ __forceinline__ int64_t operator()(int64_t i) const { return i < n ? (int64_t)cnt[i] : 0ll; }
};

static int grow_hits(bxg_itree *t, int64_t need, bool preserve) {
    if (need <= t->hits_cap) return BXG_OK;
    int64_t cap = need + need / 4 + 1024;
    int32_t *nbuf = nullptr;
    BXG_CUDA(cudaStreamSynchronize(ctx().stream));
    if (t->s_out) BXG_CUDA(cudaStreamSynchronize(t->s_out));
    BXG_CUDA(cudaMalloc(&nbuf, (size_t)cap * 4));
    if (preserve && t->d_hits && t->hits_cap)
        BXG_CUDA(cudaMemcpy(nbuf, t->d_hits, (size_t)t->hits_cap * 4, cudaMemcpyDeviceToDevice));
    cudaFree(t->d_hits);
    t->d_hits = nbuf;
    t->hits_cap = cap;
    return BXG_OK;
}

extern "C" {

int bxg_itree_create(bxg_itree_t **out) {
    BXG_TRY(ensure_init());
    if (!out) return set_error(BXG_ERR_ARG, "out is null");
    *out = new bxg_itree();
    return BXG_OK;
}

int bxg_itree_free(bxg_itree_t *t) {
    if (!t) return BXG_OK;
    cudaStreamSynchronize(ctx().stream);
    free_index(t);                          // (stops a lingering find server)
    if (t->srv_slot >= 0) g_srv.owner[t->srv_slot] = nullptr;
    cudaFree(t->d_cnt); cudaFree(t->d_lo); cudaFree(t->d_hi); cudaFree(t->d_off); cudaFree(t->d_hits);
    cudaFree(t->d_mask);
    if (t->s_in) {
        cudaStreamSynchronize(t->s_in);
        cudaStreamSynchronize(t->s_out);
        for (int k = 0; k < bxg_itree::MAX_CHUNKS; k++) {
            cudaEventDestroy(t->ev_in[k]);
            cudaEventDestroy(t->ev_scan[k]);
            cudaEventDestroy(t->ev_fill[k]);
        }
        cudaStreamDestroy(t->s_in);
        cudaStreamDestroy(t->s_out);
        cudaFreeHost(t->h_tot);
    }
    if (t->h_off) cudaFreeHost(t->h_off);
    if (t->h_hits) cudaFreeHost(t->h_hits);
    cudaFree(t->d_tiles);
    cudaFree(t->d_ticket);
    cudaFree(t->d_result);
    if (t->h_result) cudaFreeHost(t->h_result);
    if (t->m_off) cudaFreeHost(t->m_off);
    if (t->m_hits) cudaFreeHost(t->m_hits);
    delete t;
    return BXG_OK;
}

int bxg_itree_size(const bxg_itree_t *t, int64_t *n, int32_t *ntrees) {
    if (!t) return set_error(BXG_ERR_ARG, "null index handle");
    if (n) *n = t->n;
    if (ntrees) *ntrees = t->ntrees;
    return BXG_OK;
}

int bxg_itree_build(bxg_itree_t *t, const int32_t *tree, const int32_t *start, const int32_t *end, int64_t n,
                    int32_t ntrees, int loc) {
    if (!t) return set_error(BXG_ERR_ARG, "null index handle");
    if (n < 0 || n > 0x7fffff00ll) return set_error(BXG_ERR_ARG, "item count %lld out of range", (long long)n);
    if (ntrees < 1) return set_error(BXG_ERR_ARG, "ntrees must be >= 1");
    Context &c = ctx();
    BXG_CUDA(cudaStreamSynchronize(c.stream));
    free_index(t);
    t->n = n;
    t->ntrees = ntrees;
    t->nq = -1;
    BXG_CUDA(cudaMalloc(&t->toff, (size_t)(ntrees + 1) * 8));
    BXG_CUDA(cudaMemsetAsync(t->toff, 0, (size_t)(ntrees + 1) * 8, c.stream));
    if (n == 0) {
        t->nsplit = t->nsplit_pad = 0;
        t->built = true;
        return BXG_OK;
    }
    const void *dt = nullptr, *ds, *de;
    if (tree) BXG_TRY(stage_in(0, tree, (size_t)n * 4, loc, &dt));
    BXG_TRY(stage_in(1, start, (size_t)n * 4, loc, &ds));
    BXG_TRY(stage_in(2, end, (size_t)n * 4, loc, &de));
    const int32_t *d_tree = (const int32_t *)dt, *d_start = (const int32_t *)ds, *d_end = (const int32_t *)de;

    uint64_t *k0 = nullptr, *k1 = nullptr;
    int32_t *v0 = nullptr, *v1 = nullptr;
    uint32_t *tk0 = nullptr, *tk1 = nullptr;
    auto cleanup = [&]() { cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(tk0); cudaFree(tk1); };
#define BUILD_CUDA(call)                                                                                   \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess) {                                                                          \
            cleanup();                                                                                     \
            return set_error(BXG_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
        }                                                                                                  \
    } while (0)
    BUILD_CUDA(cudaMalloc(&k0, (size_t)n * 8));
    BUILD_CUDA(cudaMalloc(&k1, (size_t)n * 8));
    BUILD_CUDA(cudaMalloc(&v0, (size_t)n * 4));
    BUILD_CUDA(cudaMalloc(&v1, (size_t)n * 4));
    BUILD_CUDA(cudaMemsetAsync(c.d_mailbox + 4, 0, 8, c.stream));
    int g = grid_for(cdiv(n, 256), 8);
    BXG_LAUNCH(k_make_keys, g, 256, 0, d_tree, d_start, d_end, n, ntrees, k0, v0, (int *)(c.d_mailbox + 4));

    // sort 1: (start, end>start, tie) -- 64-bit LSD radix sort
    size_t tmp_bytes = 0, tb2 = 0;
    BUILD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k0, k1, v0, v1, n, 0, 64, c.stream));
    void *tmp;
    {
        int r = scratch(7, tmp_bytes, &tmp);
        if (r != BXG_OK) { cleanup(); return r; }
    }
    BUILD_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v0, v1, n, 0, 64, c.stream));
    c.launches += 17;   // CUB onesweep: histogram + 8 x (scan, onesweep) kernels
    int32_t *order = v1;
    if (d_tree && ntrees > 1) {
        // sort 2 (stable): bring each tree's items together, keeping the order of sort 1 inside a tree
        int bits = 1;
        while ((1ll << bits) < ntrees) bits++;
        BUILD_CUDA(cudaMalloc(&tk0, (size_t)n * 4));
        BUILD_CUDA(cudaMalloc(&tk1, (size_t)n * 4));
        BXG_LAUNCH(k_gather_tree, g, 256, 0, d_tree, v1, n, tk0);
        BUILD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb2, tk0, tk1, v1, v0, n, 0, bits, c.stream));
        {
            int r = scratch(7, tb2, &tmp);
            if (r != BXG_OK) { cleanup(); return r; }
        }
        BUILD_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb2, tk0, tk1, v1, v0, n, 0, bits, c.stream));
        c.launches += 1 + 2 * ((bits + 7) / 8);
        order = v0;
    }
    const int64_t npad = ((n + 15) & ~15ll) + 16;     // S / PM are read in aligned 16-entry groups
    BUILD_CUDA(cudaMalloc(&t->S, (size_t)npad * 4));
    BUILD_CUDA(cudaMalloc(&t->E, (size_t)npad * 4));       // walked in aligned 16-item groups; pad never hits
    BUILD_CUDA(cudaMalloc(&t->I, (size_t)npad * 4));       // read in aligned 16-item groups by the fill
    BUILD_CUDA(cudaMalloc(&t->PM, (size_t)npad * 4));
    BXG_LAUNCH(k_fill_i32, 1, 64, 0, t->S + n, npad - n, INT32_MAX);
    BXG_LAUNCH(k_fill_i32, 1, 64, 0, t->PM + n, npad - n, INT32_MAX);
    BXG_LAUNCH(k_fill_i32, 1, 64, 0, t->E + n, npad - n, INT32_MIN);
    BXG_LAUNCH(k_fill_i32, 1, 64, 0, t->I + n, npad - n, -1);
    BUILD_CUDA(cudaMemcpyAsync(t->I, order, (size_t)n * 4, cudaMemcpyDeviceToDevice, c.stream));
    // k0/k1 are free now: reuse as TE / running max
    BXG_LAUNCH(k_gather_items, g, 256, 0, d_tree && ntrees > 1 ? d_tree : nullptr, d_start, d_end, t->I, n, t->S, t->E, k0);
    BXG_LAUNCH(k_tree_offsets, (ntrees + 1 + 127) / 128, 128, 0, k0, n, ntrees, t->toff);
    BUILD_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tb2, k0, k1, cub::Max(), n, c.stream));
    {
        int r = scratch(7, tb2, &tmp);
        if (r != BXG_OK) { cleanup(); return r; }
    }
    BUILD_CUDA(cub::DeviceScan::InclusiveScan(tmp, tb2, k0, k1, cub::Max(), n, c.stream));
    c.launches += 2;
    BXG_LAUNCH(k_unpack_pm, g, 256, 0, k1, n, t->PM);

    // 32-ary max hierarchy over E
    const int32_t *src = t->E;
    int64_t len = n;
    t->nlev = 0;
    while (t->nlev < MAX_LEVELS) {
        int64_t nout = cdiv(len, 32);
        BUILD_CUDA(cudaMalloc(&t->M[t->nlev], (size_t)nout * 4));
        BXG_LAUNCH(k_block_max, grid_for(cdiv(nout * 32, 256), 8), 256, 0, src, len, t->M[t->nlev], nout);
        t->mlen[t->nlev] = nout;
        src = t->M[t->nlev];
        len = nout;
        t->nlev++;
        if (nout <= 1) break;
    }
    // splitters
    t->shift = 0;
    while (cdiv(n, 1ll << t->shift) > MAX_SPLIT) t->shift++;
    t->nsplit = (int)cdiv(n, 1ll << t->shift);
    t->nsplit_pad = (t->nsplit + 3) & ~3;
    BUILD_CUDA(cudaMalloc(&t->split, (size_t)t->nsplit_pad * 8));
    BXG_LAUNCH(k_sample, (t->nsplit_pad + 255) / 256, 256, 0, t->S, t->PM, n, t->shift, t->nsplit, t->nsplit_pad, t->split);
    // 16-ary sampled levels between the shared-memory splitters (stride 2^shift) and the arrays themselves
    t->KS[0] = t->S;
    t->KP[0] = t->PM;

    t->nk = std::max(1, (t->shift + 3) / 4);
    for (int j = 1; j < t->nk; j++) {
        const int ss = 4 * j;
        const int64_t nout = cdiv(n, 1ll << ss), nout_pad = ((nout + 15) & ~15ll) + 16;
        BUILD_CUDA(cudaMalloc(&t->KS[j], (size_t)nout_pad * 4));
        BUILD_CUDA(cudaMalloc(&t->KP[j], (size_t)nout_pad * 4));
        int gk = grid_for(cdiv(nout_pad, 256), 8);
        BXG_LAUNCH(k_sample_level, gk, 256, 0, t->S, n, ss, t->KS[j], nout, nout_pad);
        BXG_LAUNCH(k_sample_level, gk, 256, 0, t->PM, n, ss, t->KP[j], nout, nout_pad);
    }
    // 8-ary levels for the single-sector rounds of search_walk_probe8 (strides 8^j; +18 % of S / PM in total)
    t->QS[0] = t->S;
    t->QP[0] = t->PM;
    t->n8 = std::min(MAX_QLEV, std::max(1, (t->shift + 2) / 3));
    for (int j = 1; j < t->n8; j++) {
        const int ss = 3 * j;
        const int64_t nout = cdiv(n, 1ll << ss), nout_pad = ((nout + 7) & ~7ll) + 8;
        BUILD_CUDA(cudaMalloc(&t->QS[j], (size_t)nout_pad * 4));
        BUILD_CUDA(cudaMalloc(&t->QP[j], (size_t)nout_pad * 4));
        int gk = grid_for(cdiv(nout_pad, 256), 8);
        BXG_LAUNCH(k_sample_level, gk, 256, 0, t->S, n, ss, t->QS[j], nout, nout_pad);
        BXG_LAUNCH(k_sample_level, gk, 256, 0, t->PM, n, ss, t->QP[j], nout, nout_pad);
    }

    // direct-address grid (search_walk_grid): per tree a uniform grid over its start coordinates with 2-4 items per cell
    {
        void *d_ext;
        {
            int r = scratch(6, (size_t)ntrees * 8 + 16, &d_ext);
            if (r != BXG_OK) { cleanup(); return r; }
        }
        BXG_LAUNCH(k_tree_extent, (ntrees + 127) / 128, 128, 0, t->S, t->toff, ntrees, (int32_t *)d_ext);
        std::vector<int32_t> ext((size_t)ntrees * 2);
        std::vector<int64_t> h_toff((size_t)ntrees + 1);
        BUILD_CUDA(cudaMemcpyAsync(ext.data(), d_ext, (size_t)ntrees * 8, cudaMemcpyDeviceToHost, c.stream));
        BUILD_CUDA(cudaMemcpyAsync(h_toff.data(), t->toff, (size_t)(ntrees + 1) * 8, cudaMemcpyDeviceToHost, c.stream));
        BUILD_CUDA(cudaStreamSynchronize(c.stream));
        static const int dens = [] {                     // log2 of the target items per cell (lower bound); default 1 -> 2-4 items
            const char *e = getenv("BXB200_GRID_LOG2");
            return e ? std::max(0, std::min(8, atoi(e))) : 1;
        }();
        const size_t toff_bytes = (((size_t)ntrees + 1) * 8 + 15) & ~(size_t)15;
        std::vector<unsigned char> h_dir(toff_bytes + (size_t)ntrees * sizeof(bxs::GridDir), 0);
        memcpy(h_dir.data(), h_toff.data(), ((size_t)ntrees + 1) * 8);
        bxs::GridDir *gd = reinterpret_cast<bxs::GridDir *>(h_dir.data() + toff_bytes);
        int64_t nrec = 0;
        for (int k = 0; k < ntrees; k++) {
            const int64_t nt = h_toff[k + 1] - h_toff[k];
            const uint64_t span = nt > 0 ? (uint64_t)((int64_t)ext[2 * k + 1] - (int64_t)ext[2 * k]) : 0;
            const uint64_t limit = (uint64_t)std::max<int64_t>(1, nt >> dens);
            int shift = 0;
            while ((span >> shift) + 1 > limit) shift++;
            gd[k].base = ext[2 * k];
            gd[k].shift = shift;
            gd[k].ncells = nt > 0 ? (uint32_t)((span >> shift) + 1) : 0u;
            gd[k].coff = (uint32_t)nrec;
            nrec += (int64_t)gd[k].ncells + 1;
        }
        t->dir_grid_off = (uint32_t)toff_bytes;
        t->dir_bytes = (uint32_t)h_dir.size();
        t->ncells_total = nrec;
        BUILD_CUDA(cudaMalloc(&t->dir, h_dir.size()));
        BUILD_CUDA(cudaMalloc(&t->G, (size_t)nrec * sizeof(bxs::GridRec)));
        BUILD_CUDA(cudaMemcpyAsync(t->dir, h_dir.data(), h_dir.size(), cudaMemcpyHostToDevice, c.stream));
        BUILD_CUDA(cudaMalloc(&t->G16, (size_t)nrec * sizeof(bxs::GridRec16)));
        BXG_LAUNCH(k_build_grid, grid_for(cdiv(nrec, 256), 8), 256, 0, t->S, t->PM, t->toff,
                   (const bxs::GridDir *)(t->dir + toff_bytes), (int)ntrees, nrec, t->G);
        BXG_LAUNCH(k_build_grid16, grid_for(cdiv(nrec, 256), 8), 256, 0, t->S, (const bxs::GridRec *)t->G,
                   (const bxs::GridDir *)(t->dir + toff_bytes), (int)ntrees, nrec, t->G16);
        BUILD_CUDA(cudaStreamSynchronize(c.stream));     // h_dir is a local
    }

    BUILD_CUDA(cudaMemcpyAsync(c.mailbox + 4, c.d_mailbox + 4, 8, cudaMemcpyDeviceToHost, c.stream));
    BUILD_CUDA(cudaStreamSynchronize(c.stream));
    cleanup();
#undef BUILD_CUDA
    if ((int)c.mailbox[4] != 0) {
        free_index(t);
        t->n = 0;
        return set_error(BXG_ERR_ARG, "tree id out of range [0,%d)", ntrees);
    }
    t->built = true;
    return BXG_OK;
}

int bxg_itree_order(const bxg_itree_t *t, int32_t *perm, int64_t *tree_offsets) {
    if (!t || !t->built) return set_error(BXG_ERR_STATE, "index not built");
    Context &c = ctx();
    if (perm && t->n) BXG_CUDA(cudaMemcpyAsync(perm, t->I, (size_t)t->n * 4, cudaMemcpyDeviceToHost, c.stream));
    if (tree_offsets)
        BXG_CUDA(cudaMemcpyAsync(tree_offsets, t->toff, (size_t)(t->ntrees + 1) * 8, cudaMemcpyDeviceToHost, c.stream));
    BXG_CUDA(cudaStreamSynchronize(c.stream));
    return BXG_OK;
}

static int stage_queries(bxg_itree *t, const int32_t *qtree, const int32_t *qs, const int32_t *qe, int64_t nq, int loc,
                         const int32_t **dqt, const int32_t **dqs, const int32_t **dqe) {
    const void *a = nullptr, *b, *d;
    if (qtree && t->ntrees > 1) BXG_TRY(stage_in(0, qtree, (size_t)nq * 4, loc, &a));
    BXG_TRY(stage_in(1, qs, (size_t)nq * 4, loc, &b));
    BXG_TRY(stage_in(2, qe, (size_t)nq * 4, loc, &d));
    *dqt = (const int32_t *)a;
    *dqs = (const int32_t *)b;
    *dqe = (const int32_t *)d;
    return BXG_OK;
}

static int find_three_pass(bxg_itree_t *t, const int32_t *qtree, const int32_t *qs, const int32_t *qe, int64_t nq, int loc,
                           int64_t *total) {
    Context &c = ctx();
    BXG_TRY(ensure_query_buffers(t, nq));
    t->nq = nq;
    t->total = 0;
    if (nq == 0) {
        BXG_CUDA(cudaMemsetAsync(t->d_off, 0, 8, c.stream));
        if (total) *total = 0;
        return BXG_OK;
    }
    const int32_t *dqt, *dqs, *dqe;
    BXG_TRY(stage_queries(t, qtree, qs, qe, nq, loc, &dqt, &dqs, &dqe));
    // pass A: searches + per-query hit counts
    BXG_CUDA(cudaMemsetAsync(t->d_cnt + nq, 0, 4, c.stream));
    BXG_TRY(launch_count(t, dqt, dqs, dqe, nq, nullptr));
    // exclusive scan -> int64 CSR offsets (offsets[nq] = total)
    cub::TransformInputIterator<int64_t, CastI64, const int32_t *> it(t->d_cnt, CastI64());
    size_t tmp_bytes = 0;
    void *tmp;
    BXG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, it, t->d_off, nq + 1, c.stream));
    BXG_TRY(scratch(7, tmp_bytes, &tmp));
    prof_begin("cub::DeviceScan::ExclusiveSum(offsets)");
    BXG_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, it, t->d_off, nq + 1, c.stream));
    prof_end();
    c.launches += 2;
    // pass B: fill (reuses lo / masks of pass A).  It is launched BEFORE the host reads the total -- the kernel checks the
    // total against the buffer it was given and backs off if it does not fit -- so the usual step has no host round trip
    // between the scan and the fill; only a call that outgrows the hit buffer pays the sync + regrow + second launch.
    const bool staged = fill_is_staged();
    const int64_t cap_before = t->hits_cap;
    if (staged && cap_before > 0) BXG_TRY(launch_fill(t, dqs, nq));
    BXG_CUDA(cudaMemcpyAsync(c.mailbox + 5, t->d_off + nq, 8, cudaMemcpyDeviceToHost, c.stream));
    BXG_CUDA(cudaStreamSynchronize(c.stream));
    t->total = c.mailbox[5];
    if (!(staged && cap_before > 0) || t->total > cap_before) {
        BXG_TRY(grow_hits(t, t->total, false));
        if (t->total > 0) BXG_TRY(launch_fill(t, dqs, nq));
    }
    if (total) *total = t->total;
    return BXG_OK;
}

// Device-resident find with the count and the fill of different chunks OVERLAPPED.  The two kernels are bound by different
// things -- the count pass by the latency of dependent random sectors (L1 wavefront pipe ~40 % busy), the staged fill by
// L1 wavefronts (84 %) -- so run back to back each leaves most of the SM idle half the time.  Here the queries are cut into
// a few chunks; chunk k is counted and scanned on the library stream (the scan starts from the previous chunk's end
// offset, read on the device) while chunk k-1 is filled on a second stream, each kernel launched with a share of the SM's
// CTA slots so that both are resident.  Same CSR, same order; the fill checks the hit buffer's capacity itself (see
// k_fill_staged), the host learns the total at the end and re-runs the fills only if the buffer was too small.
static int ensure_pipeline(bxg_itree *t);

static void overlap_shares(int *count_ctas, int *fill_ctas, int *nchunks) {
    static int cc = 0, fc = 0, nc = 0;
    if (!cc) {
        cc = 4; fc = 3; nc = 4;
        const char *e = getenv("BXB200_OVERLAP");                 // "count_ctas,fill_ctas,chunks"
        if (e) sscanf(e, "%d,%d,%d", &cc, &fc, &nc);
        cc = std::max(1, std::min(8, cc));
        fc = std::max(1, std::min(8, fc));
        nc = std::max(2, std::min((int)bxg_itree::MAX_CHUNKS, nc));
    }
    *count_ctas = cc; *fill_ctas = fc; *nchunks = nc;
}

static int find_three_pass_overlap(bxg_itree_t *t, const int32_t *qtree, const int32_t *qs, const int32_t *qe, int64_t nq,
                                   int loc, int64_t *total) {
    Context &c = ctx();
    BXG_TRY(ensure_pipeline(t));
    BXG_TRY(ensure_query_buffers(t, nq));
    t->nq = nq;
    t->total = 0;
    const int32_t *dqt, *dqs, *dqe;
    BXG_TRY(stage_queries(t, qtree, qs, qe, nq, loc, &dqt, &dqs, &dqe));
    int count_ctas, fill_ctas, nchunks;
    overlap_shares(&count_ctas, &fill_ctas, &nchunks);
    const int64_t per = cdiv(cdiv(nq, nchunks), FIND_THREADS) * FIND_THREADS;
    nchunks = (int)cdiv(nq, per);
    size_t tmp_bytes = 0;
    {
        cub::CountingInputIterator<int64_t> idx(0);
        cub::TransformInputIterator<int64_t, ChunkCount, cub::CountingInputIterator<int64_t>> it(idx, ChunkCount{t->d_cnt, per});
        BXG_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, tmp_bytes, it, t->d_off, cub::Sum(),
                                                cub::FutureValue<int64_t>(t->d_off), per + 1, c.stream));
    }
    void *tmp;
    BXG_TRY(scratch(7, tmp_bytes, &tmp));
    BXG_CUDA(cudaMemsetAsync(t->d_off, 0, 8, c.stream));
    if (t->hits_cap == 0) BXG_TRY(grow_hits(t, 8 * nq + 1024, false));      // first guess; exact after one overflow
    auto fills = [&](cudaStream_t st, bool wait_scans) -> int {
        for (int k = 0; k < nchunks; k++) {
            const int64_t q0 = k * per, n = std::min(per, nq - q0);
            if (wait_scans) BXG_CUDA(cudaStreamWaitEvent(st, t->ev_scan[k], 0));
            c.stream_override = st;
            c.cta_cap = wait_scans ? fill_ctas : 0;
            const int r = launch_fill(t, dqs, n, q0);
            c.stream_override = nullptr;
            c.cta_cap = 0;
            BXG_TRY(r);
        }
        return BXG_OK;
    };
    for (int k = 0; k < nchunks; k++) {
        const int64_t q0 = k * per, n = std::min(per, nq - q0);
        c.cta_cap = count_ctas;
        const int r = launch_count(t, dqt, dqs, dqe, n, nullptr, q0);
        c.cta_cap = 0;
        BXG_TRY(r);
        cub::CountingInputIterator<int64_t> idx(0);
        cub::TransformInputIterator<int64_t, ChunkCount, cub::CountingInputIterator<int64_t>> it(idx, ChunkCount{t->d_cnt + q0, n});
        size_t tb = tmp_bytes;
        BXG_CUDA(cub::DeviceScan::ExclusiveScan(tmp, tb, it, t->d_off + q0, cub::Sum(),
                                                cub::FutureValue<int64_t>(t->d_off + q0), n + 1, c.stream));
        c.launches += 2;
        BXG_CUDA(cudaEventRecord(t->ev_scan[k], c.stream));
        // fill of this chunk on the second stream, as soon as its offsets exist (it overlaps the next chunk's count)
        BXG_CUDA(cudaStreamWaitEvent(t->s_out, t->ev_scan[k], 0));
        c.stream_override = t->s_out;
        c.cta_cap = fill_ctas;
        const int rf = launch_fill(t, dqs, n, q0);
        c.stream_override = nullptr;
        c.cta_cap = 0;
        BXG_TRY(rf);
    }
    BXG_CUDA(cudaEventRecord(t->ev_fill[0], t->s_out));
    BXG_CUDA(cudaStreamWaitEvent(c.stream, t->ev_fill[0], 0));            // later work on the library stream sees the hits
    BXG_CUDA(cudaMemcpyAsync(c.mailbox + 5, t->d_off + nq, 8, cudaMemcpyDeviceToHost, c.stream));
    BXG_CUDA(cudaStreamSynchronize(c.stream));
    t->total = c.mailbox[5];
    if (t->total > t->hits_cap) {                                         // the fills backed off (at least the last one)
        BXG_TRY(grow_hits(t, t->total, false));
        BXG_TRY(fills(c.stream, false));
    }
    if (total) *total = t->total;
    return BXG_OK;
}

// Host-array find with the copies overlapped: queries are cut into chunks; chunk c+1 is uploaded (copy-in stream) and
// counted while chunk c is filled and its hits / offsets travel back (copy-out stream).  The CSR offsets stay global:
// each chunk's exclusive scan starts from the previous chunk's end offset, read on the device (cub::FutureValue).
// Results land in pinned host buffers owned by the index, valid until its next find / free.
static int ensure_pipeline(bxg_itree *t);

// queries per pipeline chunk of the host path (env BXB200_CHUNK_QUERIES, default 1 Mi): large enough that a chunk's
// kernel and copies dwarf the per-chunk launch/event overhead, small enough that the first D2H starts early
static int64_t chunk_queries() {
    static int64_t v = 0;
    if (!v) {
        const char *e = getenv("BXB200_CHUNK_QUERIES");
        v = e ? atoll(e) : (1 << 20);
        if (v < 1024) v = 1024;
    }
    return v;
}

static int find_host_three_pass(bxg_itree_t *t, const int32_t *qtree, const int32_t *qs, const int32_t *qe, int64_t nq,
                                const int64_t **offsets, const int32_t **hits, int64_t *total) {
    Context &c = ctx();
    BXG_TRY(ensure_pipeline(t));
    BXG_TRY(ensure_query_buffers(t, nq));
    if (nq + 1 > t->h_off_cap) {
        BXG_CUDA(cudaStreamSynchronize(t->s_out));
        if (t->h_off) BXG_CUDA(cudaFreeHost(t->h_off));
        t->h_off_cap = nq + 1 + nq / 8;
        BXG_CUDA(cudaMallocHost(&t->h_off, (size_t)t->h_off_cap * 8));
    }
    t->nq = nq;
    t->total = 0;
    t->h_off[0] = 0;
    if (offsets) *offsets = t->h_off;
    if (hits) *hits = t->h_hits;
    if (total) *total = 0;
    if (nq == 0) return BXG_OK;

    const bool has_tree = qtree && t->ntrees > 1;
    void *p0 = nullptr, *p1, *p2;
    if (has_tree) BXG_TRY(scratch(0, (size_t)nq * 4, &p0));
    BXG_TRY(scratch(1, (size_t)nq * 4, &p1));
    BXG_TRY(scratch(2, (size_t)nq * 4, &p2));
    int32_t *dqt = (int32_t *)p0, *dqs = (int32_t *)p1, *dqe = (int32_t *)p2;
    int nchunks = (int)std::min<int64_t>(bxg_itree::MAX_CHUNKS, std::max<int64_t>(1, nq / chunk_queries()));
    const int64_t per = cdiv(nq, nchunks);
    nchunks = (int)cdiv(nq, per);
    size_t tmp_bytes = 0;
    {
        cub::CountingInputIterator<int64_t> idx(0);
        cub::TransformInputIterator<int64_t, ChunkCount, cub::CountingInputIterator<int64_t>> it(idx, ChunkCount{t->d_cnt, per});
        BXG_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, tmp_bytes, it, t->d_off, cub::Sum(),
                                                cub::FutureValue<int64_t>(t->d_off), per + 1, c.stream));
    }
    void *tmp;
    BXG_TRY(scratch(7, tmp_bytes, &tmp));
    BXG_CUDA(cudaMemsetAsync(t->d_off, 0, 8, c.stream));

    auto stage_a = [&](int k) -> int {      // upload + count + scan of chunk k
        const int64_t q0 = k * per, n = std::min(per, nq - q0);
        if (has_tree) BXG_CUDA(cudaMemcpyAsync(dqt + q0, qtree + q0, (size_t)n * 4, cudaMemcpyHostToDevice, t->s_in));
        BXG_CUDA(cudaMemcpyAsync(dqs + q0, qs + q0, (size_t)n * 4, cudaMemcpyHostToDevice, t->s_in));
        BXG_CUDA(cudaMemcpyAsync(dqe + q0, qe + q0, (size_t)n * 4, cudaMemcpyHostToDevice, t->s_in));
        BXG_CUDA(cudaEventRecord(t->ev_in[k], t->s_in));
        BXG_CUDA(cudaStreamWaitEvent(c.stream, t->ev_in[k], 0));
        BXG_TRY(launch_count(t, has_tree ? dqt : nullptr, dqs, dqe, n, nullptr, q0));
        cub::CountingInputIterator<int64_t> idx(0);
        cub::TransformInputIterator<int64_t, ChunkCount, cub::CountingInputIterator<int64_t>> it(idx, ChunkCount{t->d_cnt + q0, n});
        size_t tb = tmp_bytes;
        prof_begin("cub::DeviceScan::ExclusiveScan(offsets)");
        BXG_CUDA(cub::DeviceScan::ExclusiveScan(tmp, tb, it, t->d_off + q0, cub::Sum(),
                                                cub::FutureValue<int64_t>(t->d_off + q0), n + 1, c.stream));
        prof_end();
        c.launches += 2;
        BXG_CUDA(cudaMemcpyAsync(t->h_tot + k + 1, t->d_off + q0 + n, 8, cudaMemcpyDeviceToHost, c.stream));
        BXG_CUDA(cudaEventRecord(t->ev_scan[k], c.stream));
        return BXG_OK;
    };
    auto stage_b = [&](int k) -> int {      // fill + download of chunk k
        const int64_t q0 = k * per, n = std::min(per, nq - q0);
        BXG_CUDA(cudaEventSynchronize(t->ev_scan[k]));
        const int64_t base = k ? t->h_tot[k] : 0, end = t->h_tot[k + 1];
        BXG_TRY(grow_hits(t, end, true));
        if (end > t->h_hits_cap) {
            BXG_CUDA(cudaStreamSynchronize(t->s_out));
            int64_t cap = std::max<int64_t>(end + end / 4, (int64_t)((double)end * nq / (q0 + n) * 1.05)) + 1024;
            int32_t *nh = nullptr;
            BXG_CUDA(cudaMallocHost(&nh, (size_t)cap * 4));
            if (t->h_hits) {
                if (base) memcpy(nh, t->h_hits, (size_t)base * 4);
                BXG_CUDA(cudaFreeHost(t->h_hits));
            }
            t->h_hits = nh;
            t->h_hits_cap = cap;
        }
        if (end > base) BXG_TRY(launch_fill(t, dqs, n, q0));
        BXG_CUDA(cudaEventRecord(t->ev_fill[k], c.stream));
        BXG_CUDA(cudaStreamWaitEvent(t->s_out, t->ev_fill[k], 0));
        if (end > base)
            BXG_CUDA(cudaMemcpyAsync(t->h_hits + base, t->d_hits + base, (size_t)(end - base) * 4, cudaMemcpyDeviceToHost, t->s_out));
        BXG_CUDA(cudaMemcpyAsync(t->h_off + q0 + 1, t->d_off + q0 + 1, (size_t)n * 8, cudaMemcpyDeviceToHost, t->s_out));
        return BXG_OK;
    };
    BXG_TRY(stage_a(0));
    for (int k = 0; k < nchunks; k++) {
        if (k + 1 < nchunks) BXG_TRY(stage_a(k + 1));
        BXG_TRY(stage_b(k));
    }
    BXG_CUDA(cudaStreamSynchronize(t->s_out));
    BXG_CUDA(cudaStreamSynchronize(c.stream));
    t->total = t->h_tot[nchunks];
    if (offsets) *offsets = t->h_off;
    if (hits) *hits = t->h_hits;
    if (total) *total = t->total;
    return BXG_OK;
}

static int ensure_pipeline(bxg_itree *t) {
    if (t->s_in) return BXG_OK;
    BXG_CUDA(cudaStreamCreateWithFlags(&t->s_in, cudaStreamNonBlocking));
    BXG_CUDA(cudaStreamCreateWithFlags(&t->s_out, cudaStreamNonBlocking));
    for (int k = 0; k < bxg_itree::MAX_CHUNKS; k++) {
        BXG_CUDA(cudaEventCreateWithFlags(&t->ev_in[k], cudaEventDisableTiming));
        BXG_CUDA(cudaEventCreateWithFlags(&t->ev_scan[k], cudaEventDisableTiming));
        BXG_CUDA(cudaEventCreateWithFlags(&t->ev_fill[k], cudaEventDisableTiming));
    }
    BXG_CUDA(cudaMallocHost(&t->h_tot, (bxg_itree::MAX_CHUNKS + 1) * 8));
    return BXG_OK;
}

// ---- single-pass path --------------------------------------------------------------------------------------------
static int ensure_fused_state(bxg_itree *t, int64_t nq) {
    const int64_t need = cdiv(nq, FUSED_THREADS) + bxg_itree::MAX_CHUNKS + 1;
    if (!t->d_ticket) {
        BXG_CUDA(cudaMalloc(&t->d_ticket, (bxg_itree::MAX_CHUNKS + 1) * sizeof(unsigned int)));
        BXG_CUDA(cudaMalloc(&t->d_result, 2 * (bxg_itree::MAX_CHUNKS + 1) * sizeof(long long)));
        BXG_CUDA(cudaMallocHost(&t->h_result, 2 * (bxg_itree::MAX_CHUNKS + 1) * sizeof(long long)));
    }
    if (need > t->tiles_cap) {
        BXG_CUDA(cudaStreamSynchronize(ctx().stream));
        cudaFree(t->d_tiles);
        t->d_tiles = nullptr;
        t->tiles_cap = need + need / 4;
        BXG_CUDA(cudaMalloc(&t->d_tiles, (size_t)t->tiles_cap * 8));
    }
    return BXG_OK;
}

// one fused launch over queries [q0, q0+n); slot selects the ticket / result pair, tile0 the tile-state region;
// the chunk's base offset is read on the device from d_off[q0]
extern "C++" {
template <int PROBE>
static int launch_fused_as(bxg_itree *t, const int32_t *dqt, const int32_t *dqs, const int32_t *dqe, int64_t n, int64_t q0,
                           int slot, int64_t tile0) {
    Context &c = ctx();
    const int64_t ntiles = cdiv(n, FUSED_THREADS);
    BXG_CUDA(cudaMemsetAsync(t->d_tiles + tile0, 0, (size_t)ntiles * 8, c.stream));
    BXG_CUDA(cudaMemsetAsync(t->d_ticket + slot, 0, sizeof(unsigned int), c.stream));
    BXG_CUDA(cudaMemsetAsync(t->d_result + 2 * slot, 0, 2 * sizeof(long long), c.stream));
    size_t smem = find_smem_bytes(t, PROBE);
    static bool attr_set = false;
    if (!attr_set) {
        BXG_CUDA(cudaFuncSetAttribute(k_find_fused<PROBE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_set = true;
    }
    int occ = 0;
    BXG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_find_fused<PROBE>, FUSED_THREADS, smem));
    int grid = grid_for(ntiles, occ > 0 ? occ : 1);     // every CTA is resident: the look-back cannot starve
    BXG_LAUNCH(k_find_fused<PROBE>, grid, FUSED_THREADS, smem, t->view(), dqt ? dqt + q0 : nullptr, dqs + q0, dqe + q0, n,
               t->d_off + q0, t->d_hits, t->hits_cap, (const int64_t *)(t->d_off + q0), t->d_tiles + tile0,
               t->d_ticket + slot, t->d_result + 2 * slot);
    return BXG_OK;
}
}  // extern "C++"

static int launch_fused(bxg_itree *t, const int32_t *dqt, const int32_t *dqs, const int32_t *dqe, int64_t n, int64_t q0,
                        int slot, int64_t tile0) {
    switch (find_probe()) {
        case 0: return launch_fused_as<0>(t, dqt, dqs, dqe, n, q0, slot, tile0);
        case 1: return launch_fused_as<1>(t, dqt, dqs, dqe, n, q0, slot, tile0);
        case 2: return launch_fused_as<2>(t, dqt, dqs, dqe, n, q0, slot, tile0);
        case 3: return launch_fused_as<3>(t, dqt, dqs, dqe, n, q0, slot, tile0);
        default: return launch_fused_as<4>(t, dqt, dqs, dqe, n, q0, slot, tile0);
    }
}

static int find_fused(bxg_itree_t *t, const int32_t *qtree, const int32_t *qs, const int32_t *qe, int64_t nq, int loc,
                      int64_t *total) {
    Context &c = ctx();
    BXG_TRY(ensure_query_buffers(t, nq));
    BXG_TRY(ensure_fused_state(t, nq));
    t->nq = nq;
    t->total = 0;
    BXG_CUDA(cudaMemsetAsync(t->d_off, 0, 8, c.stream));
    if (nq == 0) {
        if (total) *total = 0;
        return BXG_OK;
    }
    const int32_t *dqt, *dqs, *dqe;
    BXG_TRY(stage_queries(t, qtree, qs, qe, nq, loc, &dqt, &dqs, &dqe));
    if (t->hits_cap == 0) BXG_TRY(grow_hits(t, 8 * nq + 1024, false));      // first guess; exact after one overflow
    for (int attempt = 0; attempt < 2; attempt++) {
        BXG_TRY(launch_fused(t, dqt, dqs, dqe, nq, 0, 0, 0));
        BXG_CUDA(cudaMemcpyAsync(t->h_result, t->d_result, 2 * sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
        BXG_CUDA(cudaStreamSynchronize(c.stream));
        t->total = t->h_result[0];
        if (!t->h_result[1]) break;
        if (attempt == 1) return set_error(BXG_ERR_STATE, "hit buffer overflow persisted after regrowth");
        BXG_TRY(grow_hits(t, t->total, false));                                // too small: offsets/total are exact, redo
        BXG_CUDA(cudaMemsetAsync(t->d_off, 0, 8, c.stream));
    }
    if (total) *total = t->total;
    return BXG_OK;
}

static int grow_host_hits(bxg_itree *t, int64_t end, int64_t keep, double progress) {
    if (end <= t->h_hits_cap) return BXG_OK;
    BXG_CUDA(cudaStreamSynchronize(t->s_out));
    int64_t cap = std::max<int64_t>(end + end / 4, (int64_t)((double)end / progress * 1.05)) + 1024;
    int32_t *nh = nullptr;
    BXG_CUDA(cudaMallocHost(&nh, (size_t)cap * 4));
    if (t->h_hits) {
        if (keep) memcpy(nh, t->h_hits, (size_t)keep * 4);
        BXG_CUDA(cudaFreeHost(t->h_hits));
    }
    t->h_hits = nh;
    t->h_hits_cap = cap;
    return BXG_OK;
}

// Host arrays, copies overlapped with the fused kernel: chunk k+1 is uploaded and searched while chunk k's hits and
// offsets travel back.  Each chunk's kernel reads its base offset from d_off[q0], written by the previous chunk.
// narrow = true: the offsets travel back as int32 (a tiny kernel narrows each chunk's int64 offsets into d_cnt, unused
// on this path) -- 4 instead of 8 bytes per query on the PCIe-bound leg; valid while the total stays below 2^31.
__global__ void k_narrow_offsets(const int64_t *__restrict__ src, int32_t *__restrict__ dst, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (int32_t)src[i];
}

static int find_host_fused(bxg_itree_t *t, const int32_t *qtree, const int32_t *qs, const int32_t *qe, int64_t nq,
                           const int64_t **offsets, const int32_t **hits, int64_t *total, bool narrow = false) {
    Context &c = ctx();
    BXG_TRY(ensure_pipeline(t));
    BXG_TRY(ensure_query_buffers(t, nq));
    BXG_TRY(ensure_fused_state(t, nq));
    if (nq + 1 > t->h_off_cap) {
        BXG_CUDA(cudaStreamSynchronize(t->s_out));
        if (t->h_off) BXG_CUDA(cudaFreeHost(t->h_off));
        t->h_off_cap = nq + 1 + nq / 8;
        BXG_CUDA(cudaMallocHost(&t->h_off, (size_t)t->h_off_cap * 8));
    }
    t->nq = nq;
    t->total = 0;
    t->h_off[0] = 0;                                   // (also the int32 view's element 0)
    int32_t *h_off32 = (int32_t *)t->h_off;
    if (offsets) *offsets = t->h_off;
    if (hits) *hits = t->h_hits;
    if (total) *total = 0;
    if (nq == 0) return BXG_OK;
    const bool has_tree = qtree && t->ntrees > 1;
    void *p0 = nullptr, *p1, *p2;
    if (has_tree) BXG_TRY(scratch(0, (size_t)nq * 4, &p0));
    BXG_TRY(scratch(1, (size_t)nq * 4, &p1));
    BXG_TRY(scratch(2, (size_t)nq * 4, &p2));
    int32_t *dqt = (int32_t *)p0, *dqs = (int32_t *)p1, *dqe = (int32_t *)p2;
    int nchunks = (int)std::min<int64_t>(bxg_itree::MAX_CHUNKS, std::max<int64_t>(1, nq / chunk_queries()));
    const int64_t per = cdiv(cdiv(nq, nchunks), FUSED_THREADS) * FUSED_THREADS;     // whole tiles per chunk
    nchunks = (int)cdiv(nq, per);
    const int64_t tiles_per = per / FUSED_THREADS;
    if (t->hits_cap == 0) BXG_TRY(grow_hits(t, 8 * nq + 1024, false));
    BXG_CUDA(cudaMemsetAsync(t->d_off, 0, 8, c.stream));

    auto launch_chunk = [&](int k) -> int {
        const int64_t q0 = k * per, n = std::min(per, nq - q0);
        BXG_TRY(launch_fused(t, has_tree ? dqt : nullptr, dqs, dqe, n, q0, k, k * tiles_per));
        if (narrow)
            BXG_LAUNCH(k_narrow_offsets, grid_for(cdiv(n, 1024), 4), 256, 0, (const int64_t *)(t->d_off + q0 + 1),
                       t->d_cnt + q0 + 1, n);
        BXG_CUDA(cudaMemcpyAsync(t->h_result + 2 * k, t->d_result + 2 * k, 2 * sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
        BXG_CUDA(cudaEventRecord(t->ev_scan[k], c.stream));
        return BXG_OK;
    };
    auto stage_a = [&](int k) -> int {      // upload + fused find of chunk k
        const int64_t q0 = k * per, n = std::min(per, nq - q0);
        if (has_tree) BXG_CUDA(cudaMemcpyAsync(dqt + q0, qtree + q0, (size_t)n * 4, cudaMemcpyHostToDevice, t->s_in));
        BXG_CUDA(cudaMemcpyAsync(dqs + q0, qs + q0, (size_t)n * 4, cudaMemcpyHostToDevice, t->s_in));
        BXG_CUDA(cudaMemcpyAsync(dqe + q0, qe + q0, (size_t)n * 4, cudaMemcpyHostToDevice, t->s_in));
        BXG_CUDA(cudaEventRecord(t->ev_in[k], t->s_in));
        BXG_CUDA(cudaStreamWaitEvent(c.stream, t->ev_in[k], 0));
        return launch_chunk(k);
    };
    int64_t base = 0;
    auto stage_b = [&](int k) -> int {      // download of chunk k (re-running it first if the hit buffer was too small)
        const int64_t q0 = k * per, n = std::min(per, nq - q0);
        BXG_CUDA(cudaEventSynchronize(t->ev_scan[k]));
        int64_t end = t->h_result[2 * k];
        if (t->h_result[2 * k + 1]) {
            BXG_TRY(grow_hits(t, std::max<int64_t>(end, (int64_t)((double)end * nq / (q0 + n) * 1.05)), true));
            BXG_TRY(launch_chunk(k));        // same base (d_off[q0] is intact), now with room
            BXG_CUDA(cudaEventSynchronize(t->ev_scan[k]));
            end = t->h_result[2 * k];
            if (t->h_result[2 * k + 1]) return set_error(BXG_ERR_STATE, "hit buffer overflow persisted after regrowth");
        }
        BXG_TRY(grow_host_hits(t, end, base, (double)(q0 + n) / nq));
        BXG_CUDA(cudaStreamWaitEvent(t->s_out, t->ev_scan[k], 0));
        if (end > base)
            BXG_CUDA(cudaMemcpyAsync(t->h_hits + base, t->d_hits + base, (size_t)(end - base) * 4, cudaMemcpyDeviceToHost, t->s_out));
        if (narrow)
            BXG_CUDA(cudaMemcpyAsync(h_off32 + q0 + 1, t->d_cnt + q0 + 1, (size_t)n * 4, cudaMemcpyDeviceToHost, t->s_out));
        else
            BXG_CUDA(cudaMemcpyAsync(t->h_off + q0 + 1, t->d_off + q0 + 1, (size_t)n * 8, cudaMemcpyDeviceToHost, t->s_out));
        base = end;
        return BXG_OK;
    };
    BXG_TRY(stage_a(0));
    for (int k = 0; k < nchunks; k++) {
        if (k + 1 < nchunks) BXG_TRY(stage_a(k + 1));
        BXG_TRY(stage_b(k));
    }
    BXG_CUDA(cudaStreamSynchronize(t->s_out));
    BXG_CUDA(cudaStreamSynchronize(c.stream));
    if (narrow && base > 0x7fffffffll) {
        t->nq = -1;
        return set_error(BXG_ERR_MISMATCH, "%lld hits do not fit int32 offsets: use bxg_itree_find_host", (long long)base);
    }
    t->total = base;
    if (offsets) *offsets = t->h_off;
    if (hits) *hits = t->h_hits;
    if (total) *total = t->total;
    return BXG_OK;
}

// find implementation: -1 = auto (default), 0 = count / scan / fill everywhere, 1 = single-pass kernel everywhere.
// Measured on B200 (profiles/r01g): with queries resident in HBM the three independent passes win (1.53 vs 1.62 ms per
// 10 M queries: no per-tile barriers, more loads in flight); on the chunk-pipelined host path the single-pass kernel
// wins (1.4e9 vs 1.2e9 queries/s end to end: one launch and one host round trip per chunk).  Auto picks accordingly.
static int g_find_mode = -2;
static int find_mode() {
    if (g_find_mode == -2) {
        const char *e = getenv("BXB200_FIND_MODE");
        g_find_mode = !e ? -1 : (e[0] == '0' ? 0 : (e[0] == '1' ? 1 : (e[0] == '2' ? 2 : -1)));
    }
    return g_find_mode;
}

// Measured on B200 (profiles/r02m): the overlapped pipeline LOSES -- 0.77-0.90 ms per 10 M queries against 0.554 ms for the
// serial passes, for every split of the CTA slots (count,fill = 4,3 / 5,2 / 3,3 / 4,4 / 3,4) and 4-8 chunks: with a share
// of the SM each kernel slows down by more than the overlap wins, and every chunk adds a scan and three launches.  So auto
// mode keeps the serial form; BXB200_FIND_OVERLAP=1 (or find mode 2) selects the pipeline for A/B runs.
static bool overlap_default() {
    static const bool on = [] {
        const char *e = getenv("BXB200_FIND_OVERLAP");
        return e && e[0] == '1';
    }();
    return on;
}

int bxg_set_find_mode(int mode) {
    if (mode < -1 || mode > 2)
        return set_error(BXG_ERR_ARG, "find mode must be -1 (auto), 0 (three-pass), 1 (single-pass) or 2 (three-pass, count/fill overlapped)");
    g_find_mode = mode;
    return BXG_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Small batches (the scalar IntervalTree.find of the reference API, one Python call per query): the general path costs
// two staged uploads, three launches and two downloads over three streams -- ~70 us for ONE query.  Here the queries
// travel as kernel arguments, one warp answers up to 32 of them (one lane each: count walk, warp scan, emit walk) and
// writes the CSR straight into mapped pinned host memory: one launch, one stream synchronise, no copies.
// ------------------------------------------------------------------------------------------------------------------
constexpr int SMALL_Q = 32;
constexpr int SMALL_CAP = 1 << 16;
struct SmallQueries {
    int32_t t[SMALL_Q], qs[SMALL_Q], qe[SMALL_Q];
};

// out_off layout (mapped pinned host memory): [0..SMALL_Q] offsets, [SMALL_Q+1] overflow flag, [SMALL_Q+2] completion
// sequence number -- written LAST, after a system-scope fence, so the host can spin on it instead of paying a
// cudaStreamSynchronize (the kernel's few PCIe writes are the whole result).
__global__ void __launch_bounds__(32)
k_find_small(IndexView ix, SmallQueries a, int nq, long long seq, long long *__restrict__ out_off,
             int32_t *__restrict__ out_hits) {
    // direct addressing (search_walk_grid): cell record -> S sector -> E sectors, no shared-memory staging at all; the
    // tree directory entries come straight from L2 (one lane per query, up to 32 queries)
    const int q = threadIdx.x;
    uint32_t lo = 0, hi = 0;
    int32_t qs = 0;
    int c = 0;
    if (q < nq) {
        const int32_t t = a.t[q], qe = a.qe[q];
        qs = a.qs[q];
        if (t >= 0 && t < ix.ntrees) {
            const uint32_t seg_lo = (uint32_t)ix.toff[t], seg_hi = (uint32_t)ix.toff[t + 1];
            const bxs::GridDir gd = reinterpret_cast<const bxs::GridDir *>(ix.dir + ix.dir_grid_off)[t];
            bxs::search_walk_grid16(ix.G16, gd, ix.S, seg_lo, seg_hi, qe, qs, ix.E, ix.M, ix.nlev,
                                    [](const bxs::GridRec16 *p) { return *p; }, Ld8(), Ld1(), hi, lo,
                                    [&](uint32_t, unsigned m) { c += __popc(m); });
        }
    }
    int incl = c;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (q >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (q < nq) out_off[q] = incl - c;
    if (q == 0) {
        out_off[nq] = total;
        out_off[SMALL_Q + 1] = total > SMALL_CAP;             // overflow: the host re-runs the general path
    }
    if (total <= SMALL_CAP && c > 0) {
        int32_t *dst = out_hits + (incl - c);
        bxs::walk_hits_halves(ix.E, ix.M, ix.nlev, lo, hi, qs, Ld8(), Ld1(), [&](uint32_t k0, unsigned m) {
            bxs::PtrSink out{dst};
            bxs::emit_group_halves_to(ix.I, k0, m, out, Ld8());
            dst = out.p;
        });
    }
    __threadfence_system();                                   // every lane's result writes before the completion word
    __syncwarp();
    if (q == 0) *(volatile long long *)(out_off + SMALL_Q + 2) = seq;
}

// wait for the completion word of a small launch: spin on the mapped host word (the kernel finishes in a few
// microseconds); if it does not show up quickly fall back to a stream synchronise, which also surfaces launch errors
static int wait_small(bxg_itree *t, long long seq) {
    volatile long long *flag = (volatile long long *)(t->m_off + SMALL_Q + 2);
    for (int spin = 0; spin < 200000; spin++) {
        if (*flag == seq) return BXG_OK;
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    BXG_CUDA(cudaStreamSynchronize(ctx().stream));
    if (*flag != seq) return set_error(BXG_ERR_CUDA, "small find kernel did not complete");
    return BXG_OK;
}

static int small_buffers(bxg_itree *t) {
    if (t->m_off) return BXG_OK;
    BXG_CUDA(cudaHostAlloc((void **)&t->m_off, (SMALL_Q + 3) * sizeof(long long), cudaHostAllocMapped));
    BXG_CUDA(cudaHostAlloc((void **)&t->m_hits, (size_t)SMALL_CAP * 4, cudaHostAllocMapped));
    BXG_CUDA(cudaHostGetDevicePointer((void **)&t->md_off, t->m_off, 0));
    BXG_CUDA(cudaHostGetDevicePointer((void **)&t->md_hits, t->m_hits, 0));
    t->m_off[SMALL_Q + 2] = 0;
    t->m_seq = 0;
    return BXG_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Lingering find server.  The loop `for line in file: tree.find(start, end)` (scripts/bed_count_overlapping.py:27-33,
// bed_intersect.py without -b ...) pays one kernel launch per call on the path above: ~2.5 us of driver work on the host
// plus ~3 us until the GPU front end starts the kernel, more than the query itself.  The server is a one-warp kernel that,
// once launched for a scalar find, stays resident and polls a 32-byte request line in mapped host memory; the next calls
// are a few plain stores into that line and a spin on the response word -- no launch.  It leaves by itself after
// `idle_us` without a request (so cudaFree / cudaDeviceSynchronize wait at most that long), when told to (any index
// rebuild or free), and is relaunched on demand with the request that found it gone.
//
//   request  words 0..3 : {slot, start, end, tree} each as (seq << 32 | value) -- every 8-byte word carries the sequence
//                         number, so a poll that catches the line half-written sees mixed sequence numbers and retries
//                         (host stores of aligned 8-byte words are atomic; no ordering between them is needed)
//   response word 4     : (seq << 32 | number of hits) or (seq << 32 | 0xffffffff) when the hits exceed SMALL_CAP;
//                         written after a system fence behind the hit ids (mapped memory of the index)
//   exit     word 5     : the generation the host gave this launch, written last; after that the kernel touches nothing
//
// The indexes are reached through a device table of descriptors (slot = one index); a descriptor only changes while no
// server runs (server_stop first), so the server may cache it.  One host thread at a time, like the rest of the library.
// ------------------------------------------------------------------------------------------------------------------
extern "C++" {
static bool server_enabled() {
    if (g_srv.enabled < 0) {
        const char *e = getenv("BXB200_FIND_SERVER");
        g_srv.enabled = e ? (e[0] == '1') : BXB200_FIND_SERVER_DEFAULT;
        const char *i = getenv("BXB200_FIND_SERVER_IDLE_US");
        if (i && atoi(i) > 0) g_srv.idle_us = (unsigned)atoi(i);
    }
    return g_srv.enabled == 1;
}

static inline unsigned long long srv_word(const volatile unsigned long long *p) { return *p; }
static bool server_alive() { return g_srv.gen != 0 && srv_word(g_srv.m + 5) != g_srv.gen; }

static void server_post(uint32_t slot, int32_t a, int32_t b, int32_t c) {
    const unsigned long long k = (unsigned long long)(++g_srv.seq) << 32;
    __atomic_store_n(g_srv.m + 0, k | slot, __ATOMIC_RELAXED);
    __atomic_store_n(g_srv.m + 1, k | (uint32_t)a, __ATOMIC_RELAXED);
    __atomic_store_n(g_srv.m + 2, k | (uint32_t)b, __ATOMIC_RELAXED);
    __atomic_store_n(g_srv.m + 3, k | (uint32_t)c, __ATOMIC_RELEASE);
}

__device__ __forceinline__ unsigned long long srv_timer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(32)
k_find_server(const FindDesc *__restrict__ table, unsigned long long *mbox, uint32_t last_seq, unsigned long long gen,
              unsigned long long idle_ns) {
    const int lane = threadIdx.x;
    unsigned long long t_last = srv_timer();
    // the poll count is a second, clock-independent bound on the kernel's life (a poll is a PCIe read, >= ~0.5 us)
    for (unsigned long long polls = 0; polls < (idle_ns >> 4) + (1u << 16); polls++) {
        unsigned long long w = 0;
        if (lane < 4) w = *(volatile unsigned long long *)(mbox + lane);
        const uint32_t sq = (uint32_t)(w >> 32), val = (uint32_t)w;
        const uint32_t s0 = __shfl_sync(0xffffffffu, sq, 0);
        const bool whole = __all_sync(0xffffffffu, lane >= 4 || sq == s0);
        if (whole && s0 != last_seq) {
            last_seq = s0;
            const uint32_t slot = __shfl_sync(0xffffffffu, val, 0);
            const int32_t qs = (int32_t)__shfl_sync(0xffffffffu, val, 1), qe = (int32_t)__shfl_sync(0xffffffffu, val, 2);
            const int32_t t = (int32_t)__shfl_sync(0xffffffffu, val, 3);
            if (slot == SRV_KILL || slot >= SRV_SLOTS) break;
            if (lane == 0) {
                const IndexView &ix = table[slot].ix;
                uint32_t lo = 0, hi = 0;
                int c = 0;
                if (t >= 0 && t < ix.ntrees) {
                    const uint32_t seg_lo = (uint32_t)ix.toff[t], seg_hi = (uint32_t)ix.toff[t + 1];
                    const bxs::GridDir gd = reinterpret_cast<const bxs::GridDir *>(ix.dir + ix.dir_grid_off)[t];
                    bxs::search_walk_grid16(ix.G16, gd, ix.S, seg_lo, seg_hi, qe, qs, ix.E, ix.M, ix.nlev,
                                            [](const bxs::GridRec16 *p) { return *p; }, Ld8(), Ld1(), hi, lo,
                                            [&](uint32_t, unsigned m) { c += __popc(m); });
                }
                if (c > 0 && c <= SMALL_CAP) {
                    int32_t *dst = table[slot].out_hits;
                    bxs::walk_hits_halves(ix.E, ix.M, ix.nlev, lo, hi, qs, Ld8(), Ld1(), [&](uint32_t k0, unsigned m) {
                        bxs::PtrSink out{dst};
                        bxs::emit_group_halves_to(ix.I, k0, m, out, Ld8());
                        dst = out.p;
                    });
                }
                __threadfence_system();                       // the hit ids before the response word
                *(volatile unsigned long long *)(mbox + 4) =
                    ((unsigned long long)s0 << 32) | (c <= SMALL_CAP ? (uint32_t)c : SRV_OVERFLOW);
            }
            __syncwarp();
            t_last = srv_timer();
            polls = 0;
        } else if (__any_sync(0xffffffffu, srv_timer() - t_last > idle_ns)) {   // (one decision for the whole warp)
            break;
        }
    }
    __syncwarp();
    if (lane == 0) {
        __threadfence_system();
        *(volatile unsigned long long *)(mbox + 5) = gen;      // last: the host may relaunch from here on
    }
}

static int server_launch() {
    FindServer &s = g_srv;
    ++s.gen;
    ++s.launches;
    k_find_server<<<1, 32, 0, s.stream>>>(s.d_table, s.md, s.seq - 1, s.gen, (unsigned long long)s.idle_us * 1000ull);
    BXG_CUDA(cudaGetLastError());
    return BXG_OK;
}

// tell a lingering server to leave and wait until it has (a few microseconds; bounded by its idle time-out in any case)
static void server_stop() {
    FindServer &s = g_srv;
    if (!s.m || !server_alive()) return;
    server_post(SRV_KILL, 0, 0, 0);
    for (int spin = 0; spin < 2000000; spin++) {
        if (srv_word(s.m + 5) == s.gen) return;
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    cudaStreamSynchronize(s.stream);
}

static int server_init() {
    FindServer &s = g_srv;
    if (s.m) return BXG_OK;
    BXG_CUDA(cudaMalloc((void **)&s.d_table, sizeof(FindDesc) * SRV_SLOTS));
    BXG_CUDA(cudaHostAlloc((void **)&s.m, 64, cudaHostAllocMapped));
    BXG_CUDA(cudaHostGetDevicePointer((void **)&s.md, s.m, 0));
    for (int i = 0; i < 8; i++) s.m[i] = 0;
    BXG_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    return BXG_OK;
}

// -> number of hits (>= 0), a negative status, or FIND1_NO_SERVER when this index cannot use the server
constexpr int64_t FIND1_NO_SERVER = INT64_MIN;
static int64_t find1_server(bxg_itree *t, int32_t tree, int32_t start, int32_t end, const int32_t **hits) {
    FindServer &s = g_srv;
    if (int rc = server_init()) return rc;
    if (!t->m_off) server_stop();           // (pinned allocations wait for resident kernels)
    if (int rc = small_buffers(t)) return rc;
    if (t->srv_slot < 0) {
        for (int k = 0; k < SRV_SLOTS && t->srv_slot < 0; k++)
            if (!s.owner[k]) {
                s.owner[k] = t;
                t->srv_slot = k;
            }
        if (t->srv_slot < 0) return FIND1_NO_SERVER;
        t->srv_dirty = true;
    }
    if (t->srv_dirty) {
        server_stop();
        const FindDesc d{t->view(), t->md_hits};
        if (cudaMemcpyAsync(s.d_table + t->srv_slot, &d, sizeof d, cudaMemcpyHostToDevice, ctx().stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx().stream) != cudaSuccess)      // also: the build's kernels have finished
            return set_error(BXG_ERR_CUDA, "find server: descriptor upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        t->srv_dirty = false;
    }
    ++s.requests;
    server_post((uint32_t)t->srv_slot, start, end, tree);
    const uint32_t want = s.seq;
    if (!server_alive())
        if (int rc = server_launch()) return rc;
    for (int relaunch = 0; relaunch < 4; relaunch++) {
        for (long long spin = 0; spin < 4000000; spin++) {
            unsigned long long r = srv_word(s.m + 4);
            if ((uint32_t)(r >> 32) != want && srv_word(s.m + 5) == s.gen) {
                r = srv_word(s.m + 4);                         // it left: the response, if any, was written before the exit word
                if ((uint32_t)(r >> 32) != want) break;        // left without serving this request -> start another
            }
            if ((uint32_t)(r >> 32) == want) {
                const uint32_t c = (uint32_t)r;
                if (c == SRV_OVERFLOW) return FIND1_NO_SERVER;
                if (hits) *hits = t->m_hits;
                return (int64_t)c;
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        if (server_alive()) {                                  // no answer and no exit within the spin budget
            if (cudaStreamSynchronize(s.stream) != cudaSuccess)
                return set_error(BXG_ERR_CUDA, "find server: %s", cudaGetErrorString(cudaGetLastError()));
        }
        if ((uint32_t)(srv_word(s.m + 4) >> 32) == want) continue;
        if (int rc = server_launch()) return rc;
    }
    return set_error(BXG_ERR_CUDA, "find server did not answer");
}
}  // extern "C++"

int bxg_set_find_server(int on) {
    if (on < -1 || on > 1) return set_error(BXG_ERR_ARG, "find server: 0 (off), 1 (on) or -1 (the default / environment setting)");
    if (on < 0) {
        g_srv.enabled = -1;
        on = server_enabled();
    }
    server_enabled();                       // (reads the idle time-out from the environment once)
    if (!on) server_stop();
    g_srv.enabled = on;
    return BXG_OK;
}

int bxg_find_server_stats(int64_t *launches, int64_t *requests, int32_t *alive) {
    if (launches) *launches = g_srv.launches;
    if (requests) *requests = g_srv.requests;
    if (alive) *alive = g_srv.m && server_alive();
    return BXG_OK;
}

int bxg_itree_find_small(bxg_itree_t *t, const int32_t *qtree, const int32_t *qs, const int32_t *qe, int32_t nq,
                         const int64_t **offsets, const int32_t **hits, int64_t *total) {
    if (!t || !t->built) return set_error(BXG_ERR_STATE, "index not built");
    if (nq < 0 || nq > SMALL_Q) return set_error(BXG_ERR_ARG, "bxg_itree_find_small takes 0..%d queries", SMALL_Q);
    BXG_TRY(small_buffers(t));
    if (nq == 0 || t->n == 0) {
        for (int q = 0; q <= nq; q++) t->m_off[q] = 0;
    } else {
        SmallQueries a;
        for (int q = 0; q < nq; q++) {
            a.t[q] = qtree ? qtree[q] : 0;
            a.qs[q] = qs[q];
            a.qe[q] = qe[q];
        }
        const long long seq = ++t->m_seq;
        BXG_LAUNCH(k_find_small, 1, 32, 0, t->view(), a, (int)nq, seq, t->md_off, t->md_hits);
        BXG_TRY(wait_small(t, seq));
        if (t->m_off[SMALL_Q + 1])                   // more than SMALL_CAP hits: the general path has no such limit
            return bxg_itree_find_host(t, qtree, qs, qe, nq, offsets, hits, total);
    }
    static_assert(sizeof(long long) == sizeof(int64_t), "offset width");
    if (offsets) *offsets = (const int64_t *)t->m_off;
    if (hits) *hits = t->m_hits;
    if (total) *total = t->m_off[nq];
    return BXG_OK;
}

// The scalar IntervalTree.find(start, end) (intersection.pyx:400-406) with the thinnest possible call: plain integers in,
// the number of hits (>= 0) or a negative status out; *hits points into the index's mapped result buffer.
int64_t bxg_itree_find1(bxg_itree_t *t, int32_t tree, int32_t start, int32_t end, const int32_t **hits) {
    if (!t || !t->built) return set_error(BXG_ERR_STATE, "index not built");
    if (t->n > 0 && server_enabled() && !prof_enabled()) {
        const int64_t r = find1_server(t, tree, start, end, hits);
        if (r != FIND1_NO_SERVER) return r;
    }
    const int64_t *off = nullptr;
    int64_t total = 0;
    const int rc = bxg_itree_find_small(t, &tree, &start, &end, 1, &off, hits, &total);
    return rc == BXG_OK ? total : (int64_t)rc;
}

int bxg_itree_find(bxg_itree_t *t, const int32_t *qtree, const int32_t *qs, const int32_t *qe, int64_t nq, int loc,
                   int64_t *total) {
    if (!t || !t->built) return set_error(BXG_ERR_STATE, "index not built");
    if (nq < 0) return set_error(BXG_ERR_ARG, "nq < 0");
    const int mode = find_mode();
    if (mode == 1) return find_fused(t, qtree, qs, qe, nq, loc, total);
    // overlapped count / fill for large batches (mode 2, or auto); the serial three-pass form when per-kernel profiling is
    // on (each kernel timed alone is what the roofline entries quote) and for the direct-store fill variant
    const bool overlap = (mode == 2 || (mode == -1 && overlap_default())) && nq >= (1 << 20) && !prof_enabled() && fill_is_staged();
    return overlap ? find_three_pass_overlap(t, qtree, qs, qe, nq, loc, total) : find_three_pass(t, qtree, qs, qe, nq, loc, total);
}

int bxg_itree_find_host(bxg_itree_t *t, const int32_t *qtree, const int32_t *qs, const int32_t *qe, int64_t nq,
                        const int64_t **offsets, const int32_t **hits, int64_t *total) {
    if (!t || !t->built) return set_error(BXG_ERR_STATE, "index not built");
    if (nq < 0) return set_error(BXG_ERR_ARG, "nq < 0");
    return find_mode() != 0 ? find_host_fused(t, qtree, qs, qe, nq, offsets, hits, total)
                            : find_host_three_pass(t, qtree, qs, qe, nq, offsets, hits, total);
}

int bxg_itree_find_host32(bxg_itree_t *t, const int32_t *qtree, const int32_t *qs, const int32_t *qe, int64_t nq,
                          const int32_t **offsets, const int32_t **hits, int64_t *total) {
    if (!t || !t->built) return set_error(BXG_ERR_STATE, "index not built");
    if (nq < 0) return set_error(BXG_ERR_ARG, "nq < 0");
    const int64_t *off64 = nullptr;
    BXG_TRY(find_host_fused(t, qtree, qs, qe, nq, &off64, hits, total, true));
    if (offsets) *offsets = (const int32_t *)off64;     // the same pinned buffer, filled as int32
    return BXG_OK;
}

int bxg_itree_fetch(bxg_itree_t *t, int64_t *offsets, int32_t *hits) {
    if (!t || t->nq < 0) return set_error(BXG_ERR_STATE, "bxg_itree_find must be called first");
    Context &c = ctx();
    if (offsets) BXG_CUDA(cudaMemcpyAsync(offsets, t->d_off, (size_t)(t->nq + 1) * 8, cudaMemcpyDeviceToHost, c.stream));
    if (hits && t->total) BXG_CUDA(cudaMemcpyAsync(hits, t->d_hits, (size_t)t->total * 4, cudaMemcpyDeviceToHost, c.stream));
    BXG_CUDA(cudaStreamSynchronize(c.stream));
    return BXG_OK;
}

int bxg_itree_result_dev(const bxg_itree_t *t, const int64_t **d_offsets, const int32_t **d_hits, int64_t *nq, int64_t *total) {
    if (!t || t->nq < 0) return set_error(BXG_ERR_STATE, "bxg_itree_find must be called first");
    if (d_offsets) *d_offsets = t->d_off;
    if (d_hits) *d_hits = t->d_hits;
    if (nq) *nq = t->nq;
    if (total) *total = t->total;
    return BXG_OK;
}

// len(find(...)) for HOST arrays with the copies overlapped (the whole of scripts/bed_count_overlapping.py:27-33): chunk
// k+1 is uploaded while chunk k is counted and chunk k-1's counts travel back.  16 bytes per query cross PCIe (12 up, 4
// down) instead of the ~42 of the full CSR.
static int count_host_pipelined(bxg_itree_t *t, const int32_t *qtree, const int32_t *qs, const int32_t *qe, int64_t nq,
                                int32_t *counts, int64_t *total) {
    Context &c = ctx();
    BXG_TRY(ensure_pipeline(t));
    const bool has_tree = qtree && t->ntrees > 1;
    void *p0 = nullptr, *p1, *p2;
    if (has_tree) BXG_TRY(scratch(0, (size_t)nq * 4, &p0));
    BXG_TRY(scratch(1, (size_t)nq * 4, &p1));
    BXG_TRY(scratch(2, (size_t)nq * 4, &p2));
    int32_t *dqt = (int32_t *)p0, *dqs = (int32_t *)p1, *dqe = (int32_t *)p2;
    int nchunks = (int)std::min<int64_t>(bxg_itree::MAX_CHUNKS, std::max<int64_t>(1, nq / chunk_queries()));
    const int64_t per = cdiv(nq, nchunks);
    nchunks = (int)cdiv(nq, per);
    unsigned long long *d_total = (unsigned long long *)(c.d_mailbox + 6);
    BXG_CUDA(cudaMemsetAsync(d_total, 0, 8, c.stream));
    for (int k = 0; k < nchunks; k++) {
        const int64_t q0 = k * per, n = std::min(per, nq - q0);
        if (has_tree) BXG_CUDA(cudaMemcpyAsync(dqt + q0, qtree + q0, (size_t)n * 4, cudaMemcpyHostToDevice, t->s_in));
        BXG_CUDA(cudaMemcpyAsync(dqs + q0, qs + q0, (size_t)n * 4, cudaMemcpyHostToDevice, t->s_in));
        BXG_CUDA(cudaMemcpyAsync(dqe + q0, qe + q0, (size_t)n * 4, cudaMemcpyHostToDevice, t->s_in));
        BXG_CUDA(cudaEventRecord(t->ev_in[k], t->s_in));
        BXG_CUDA(cudaStreamWaitEvent(c.stream, t->ev_in[k], 0));
        BXG_TRY(launch_count(t, has_tree ? dqt : nullptr, dqs, dqe, n, d_total, q0));
        BXG_CUDA(cudaEventRecord(t->ev_scan[k], c.stream));
        if (counts) {
            BXG_CUDA(cudaStreamWaitEvent(t->s_out, t->ev_scan[k], 0));
            BXG_CUDA(cudaMemcpyAsync(counts + q0, t->d_cnt + q0, (size_t)n * 4, cudaMemcpyDeviceToHost, t->s_out));
        }
    }
    BXG_CUDA(cudaMemcpyAsync(c.mailbox + 6, d_total, 8, cudaMemcpyDeviceToHost, c.stream));
    BXG_CUDA(cudaStreamSynchronize(c.stream));
    BXG_CUDA(cudaStreamSynchronize(t->s_out));
    if (total) *total = c.mailbox[6];
    return BXG_OK;
}

int bxg_itree_count(bxg_itree_t *t, const int32_t *qtree, const int32_t *qs, const int32_t *qe, int64_t nq, int loc,
                    int32_t *counts, int64_t *total) {
    if (!t || !t->built) return set_error(BXG_ERR_STATE, "index not built");
    if (nq < 0) return set_error(BXG_ERR_ARG, "nq < 0");
    Context &c = ctx();
    if (nq == 0) {
        if (total) *total = 0;
        return BXG_OK;
    }
    BXG_TRY(ensure_query_buffers(t, nq));
    t->nq = -1;
    if (loc == BXG_HOST && nq >= 2 * chunk_queries()) return count_host_pipelined(t, qtree, qs, qe, nq, counts, total);
    const int32_t *dqt, *dqs, *dqe;
    BXG_TRY(stage_queries(t, qtree, qs, qe, nq, loc, &dqt, &dqs, &dqe));
    BXG_CUDA(cudaMemsetAsync(c.d_mailbox + 6, 0, 8, c.stream));
    BXG_TRY(launch_count(t, dqt, dqs, dqe, nq, (unsigned long long *)(c.d_mailbox + 6)));
    if (counts)
        BXG_CUDA(cudaMemcpyAsync(counts, t->d_cnt, (size_t)nq * 4,
                                 loc == BXG_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c.stream));
    if (total || loc == BXG_HOST) {
        BXG_CUDA(cudaMemcpyAsync(c.mailbox + 6, c.d_mailbox + 6, 8, cudaMemcpyDeviceToHost, c.stream));
        BXG_CUDA(cudaStreamSynchronize(c.stream));
        if (total) *total = c.mailbox[6];
    }
    return BXG_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// Neighbour queries: IntervalNode.left / right (intersection.pyx:192-260) = IntervalTree.before / after (:408-426).
//
// after(pos, n, md):  p = pos+1; candidates = items with 0 <= start-p < md, collected in in-order; if exactly n they
//   are returned as collected, else stable-sorted by start and cut to n.  In-order IS start order, so the answer is
//   always the first min(n, count) items of the contiguous range [lb(S,p), lb(S,p+md)).
// before(pos, n, md): p = pos-1; candidates = items with 0 <= p-end < md collected in REVERSED in-order; if exactly n
//   they are returned as collected (descending in-order position k), else stable-sorted by end descending and cut to n.
//   With a second per-tree ordering by (end, k) (built lazily, one stable radix sort) the candidates are the range
//   [first end > p-md, first end > p) and "stable sort of the reversed sequence by -end" is exactly that range walked
//   backwards; the count == n case sorts the n emitted entries by k descending.
// ------------------------------------------------------------------------------------------------------------------
struct EndOrder {
    int32_t *end = nullptr;   // ends sorted per tree
    int32_t *k = nullptr;     // in-order position of each entry
};

template <bool GT>   // first j in [lo,hi) with A[j] >= v (GT: A[j] > v); 64-bit compare so p +/- max_dist cannot wrap
__device__ __forceinline__ uint32_t bound64(const int32_t *__restrict__ A, uint32_t lo, uint32_t hi, int64_t v) {
    while (lo < hi) {
        uint32_t m = (lo + hi) >> 1;
        int64_t a = A[m];
        if (GT ? (a <= v) : (a < v)) lo = m + 1; else hi = m;
    }
    return lo;
}

template <bool FILL>
__global__ void k_neighbors(const __grid_constant__ IndexView ix, const int32_t *__restrict__ ES_end,
                            const int32_t *__restrict__ ES_k, int dir, const int32_t *__restrict__ qtree,
                            const int32_t *__restrict__ pos, const int32_t *__restrict__ nmax,
                            const int32_t *__restrict__ maxd, int64_t nq, int32_t *__restrict__ cnt,
                            const int64_t *__restrict__ off, int32_t *__restrict__ hits) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride) {
        const int32_t t = qtree ? qtree[q] : 0;
        int32_t c = 0;
        if (t >= 0 && t < ix.ntrees) {
            const uint32_t seg_lo = (uint32_t)ix.toff[t], seg_hi = (uint32_t)ix.toff[t + 1];
            const int64_t md = maxd[q];
            const int32_t want = nmax[q] > 0 ? nmax[q] : 0;
            if (dir == 1) {                                   // after / right
                const int64_t p = (int64_t)pos[q] + 1;
                const uint32_t a = bound64<false>(ix.S, seg_lo, seg_hi, p);
                const uint32_t b = md > 0 ? bound64<false>(ix.S, a, seg_hi, p + md) : a;
                const uint32_t have = b - a;
                c = (int32_t)(have < (uint32_t)want ? have : (uint32_t)want);
                if (FILL) {
                    int32_t *dst = hits + off[q];
                    for (int32_t j = 0; j < c; j++) dst[j] = ix.I[a + j];
                }
            } else {                                          // before / left
                const int64_t p = (int64_t)pos[q] - 1;
                const uint32_t b = bound64<true>(ES_end, seg_lo, seg_hi, p);             // first end > p
                const uint32_t a = md > 0 ? bound64<true>(ES_end, seg_lo, b, p - md) : b;  // first end > p - md
                const uint32_t have = b - a;
                c = (int32_t)(have < (uint32_t)want ? have : (uint32_t)want);
                if (FILL && c > 0) {
                    int32_t *dst = hits + off[q];
                    if (have != (uint32_t)want) {
                        for (int32_t j = 0; j < c; j++) dst[j] = ix.I[ES_k[b - 1 - j]];    // by end desc, then k desc
                    } else {
                        // exactly n candidates: the reference returns them as collected = by in-order position, descending
                        for (int32_t j = 0; j < c; j++) {                                   // insertion sort on k
                            int32_t kk = ES_k[a + j];
                            int32_t i = j;
                            while (i > 0 && dst[i - 1] < kk) { dst[i] = dst[i - 1]; i--; }
                            dst[i] = kk;
                        }
                        for (int32_t j = 0; j < c; j++) dst[j] = ix.I[dst[j]];
                    }
                }
            }
        }
        if (!FILL) cnt[q] = c;
    }
}

// (tree, end) keys in in-order sequence; a stable sort then keeps ties in ascending in-order position
__global__ void k_end_keys(const int32_t *__restrict__ E, const int64_t *__restrict__ toff, int32_t ntrees, int64_t n,
                           uint64_t *__restrict__ keys, int32_t *__restrict__ vals) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        int lo = 0, hi = ntrees;                 // tree of position k: largest t with toff[t] <= k
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (toff[mid] <= k) lo = mid; else hi = mid;
        }
        keys[k] = ((uint64_t)lo << 32) | (uint64_t)((uint32_t)E[k] ^ 0x80000000u);
        vals[k] = (int32_t)k;
    }
}
__global__ void k_end_unpack(const uint64_t *__restrict__ keys, int64_t n, int32_t *__restrict__ end) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride)
        end[k] = (int32_t)((uint32_t)keys[k] ^ 0x80000000u);
}

static int ensure_end_order(bxg_itree *t) {
    if (t->es_end || t->n == 0) return BXG_OK;
    Context &c = ctx();
    const int64_t n = t->n;
    uint64_t *k0 = nullptr, *k1 = nullptr;
    int32_t *v0 = nullptr;
    BXG_CUDA(cudaMalloc(&k0, (size_t)n * 8));
    BXG_CUDA(cudaMalloc(&k1, (size_t)n * 8));
    BXG_CUDA(cudaMalloc(&v0, (size_t)n * 4));
    BXG_CUDA(cudaMalloc(&t->es_end, (size_t)n * 4));
    BXG_CUDA(cudaMalloc(&t->es_k, (size_t)n * 4));
    int g = grid_for(cdiv(n, 256), 8);
    BXG_LAUNCH(k_end_keys, g, 256, 0, t->E, t->toff, t->ntrees, n, k0, v0);
    size_t tb = 0;
    void *tmp;
    BXG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, k0, k1, v0, t->es_k, n, 0, 64, c.stream));
    BXG_TRY(scratch(7, tb, &tmp));
    BXG_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, k0, k1, v0, t->es_k, n, 0, 64, c.stream));
    c.launches += 17;
    BXG_LAUNCH(k_end_unpack, g, 256, 0, k1, n, t->es_end);
    BXG_CUDA(cudaStreamSynchronize(c.stream));
    cudaFree(k0); cudaFree(k1); cudaFree(v0);
    return BXG_OK;
}

extern "C" int bxg_itree_neighbors(bxg_itree_t *t, const int32_t *qtree, const int32_t *pos, const int32_t *n,
                                   const int32_t *max_dist, int64_t nq, int dir, int loc, int64_t *total) {
    if (!t || !t->built) return set_error(BXG_ERR_STATE, "index not built");
    if (nq < 0 || (dir != 0 && dir != 1)) return set_error(BXG_ERR_ARG, "bad arguments");
    Context &c = ctx();
    BXG_TRY(ensure_query_buffers(t, nq));
    t->nq = nq;
    t->total = 0;
    if (nq == 0 || t->n == 0) {
        BXG_CUDA(cudaMemsetAsync(t->d_off, 0, (size_t)(nq + 1) * 8, c.stream));
        if (total) *total = 0;
        return BXG_OK;
    }
    if (dir == 0) BXG_TRY(ensure_end_order(t));
    const void *a = nullptr, *b, *d, *e;
    if (qtree && t->ntrees > 1) BXG_TRY(stage_in(0, qtree, (size_t)nq * 4, loc, &a));
    BXG_TRY(stage_in(1, pos, (size_t)nq * 4, loc, &b));
    BXG_TRY(stage_in(2, n, (size_t)nq * 4, loc, &d));
    BXG_TRY(stage_in(5, max_dist, (size_t)nq * 4, loc, &e));
    int grid = grid_for(cdiv(nq, 256), 8);
    BXG_CUDA(cudaMemsetAsync(t->d_cnt + nq, 0, 4, c.stream));
    BXG_LAUNCH((k_neighbors<false>), grid, 256, 0, t->view(), t->es_end, t->es_k, dir, (const int32_t *)a,
               (const int32_t *)b, (const int32_t *)d, (const int32_t *)e, nq, t->d_cnt, (const int64_t *)nullptr,
               (int32_t *)nullptr);
    cub::TransformInputIterator<int64_t, CastI64, const int32_t *> it(t->d_cnt, CastI64());
    size_t tmp_bytes = 0;
    void *tmp;
    BXG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, it, t->d_off, nq + 1, c.stream));
    BXG_TRY(scratch(7, tmp_bytes, &tmp));
    BXG_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, it, t->d_off, nq + 1, c.stream));
    c.launches += 2;
    BXG_CUDA(cudaMemcpyAsync(c.mailbox + 5, t->d_off + nq, 8, cudaMemcpyDeviceToHost, c.stream));
    BXG_CUDA(cudaStreamSynchronize(c.stream));
    t->total = c.mailbox[5];
    BXG_TRY(grow_hits(t, t->total, false));
    if (t->total > 0)
        BXG_LAUNCH((k_neighbors<true>), grid, 256, 0, t->view(), t->es_end, t->es_k, dir, (const int32_t *)a,
                   (const int32_t *)b, (const int32_t *)d, (const int32_t *)e, nq, t->d_cnt, (const int64_t *)t->d_off,
                   t->d_hits);
    if (loc == BXG_HOST) BXG_CUDA(cudaStreamSynchronize(c.stream));
    if (total) *total = t->total;
    return BXG_OK;
}
