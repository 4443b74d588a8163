// itree_search.cuh -- the search / walk logic of the find kernels, written once for device and host.
//
// Compiled by nvcc into k_find (csrc/itree.cu) and by g++ into the CPU fuzz harness (tests/search_fuzz.cpp), so the
// index arithmetic that cannot be watched on a GPU-less box is still exercised against std::lower_bound there.
//
//   S-search : hi = first k in the tree segment with S[k]  >= qe      (items with start <  qe are candidates)
//   PM-search: lo = first k in the tree segment with PM[k] >  qs      (first item whose running max end > qs)
// Both run in lock-step so their loads overlap: a fixed-trip binary-lifting search over the shared-memory splitters,
// then 16-ary rounds over the sampled levels (one aligned 64-byte group per round and search).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BXG_HD __host__ __device__ __forceinline__
#else
#define BXG_HD inline
struct int4 {
    int x, y, z, w;
};
#endif

namespace bxs {

struct Win {
    uint32_t lo, hi;   // answer lies in [lo, hi]; everything below lo is "before", position hi is not (or is seg end)
};

template <bool LESS_EQ>
BXG_HD bool before(int32_t v, int32_t key) {
    return LESS_EQ ? (v <= key) : (v < key);
}

// number of splitters in [k0,k1) that are before `key`, as k0 + count; fixed trip count for a given (k0,k1)
template <bool LESS_EQ, typename SP>
BXG_HD uint32_t lift_step(const SP &sp, uint32_t a, uint32_t step, uint32_t k1, int32_t key) {
    return (a + step <= k1 && before<LESS_EQ>(sp[a + step - 1], key)) ? a + step : a;
}

// window after the splitter phase: a = k0 + (#splitters in [k0,k1) before key)
BXG_HD Win splitter_window(uint32_t a, uint32_t k0, uint32_t k1, int shift, uint32_t seg_lo, uint32_t seg_hi) {
    Win w{seg_lo, seg_hi};
    if (k0 < k1) {
        if (a == k0) {
            w.hi = k0 << shift;                     // A[k0<<shift] is not before key
        } else {
            w.lo = ((a - 1) << shift) + 1;          // A[(a-1)<<shift] is before key
            uint32_t h = a << shift;
            if (a < k1 && h < w.hi) w.hi = h;
        }
    }
    return w;
}

struct Round {
    uint32_t m0, m1, g;
    bool active;
};

// samples m of level stride 2^ss with lo <= m<<ss < hi; they live in one aligned 16-entry group starting at g
BXG_HD Round round_prepare(const Win &w, int ss) {
    Round r{0, 0, 0, false};
    if (w.lo >= w.hi) return r;
    r.m0 = (w.lo + (1u << ss) - 1) >> ss;
    r.m1 = ((w.hi - 1) >> ss) + 1;
    r.g = r.m0 & ~15u;
    r.active = r.m0 < r.m1 && r.m1 - r.g <= 16u;
    return r;
}

template <bool LESS_EQ>
BXG_HD unsigned group_mask(const int4 &v0, const int4 &v1, const int4 &v2, const int4 &v3, int32_t key) {
    return (unsigned)before<LESS_EQ>(v0.x, key) | (unsigned)before<LESS_EQ>(v0.y, key) << 1 |
           (unsigned)before<LESS_EQ>(v0.z, key) << 2 | (unsigned)before<LESS_EQ>(v0.w, key) << 3 |
           (unsigned)before<LESS_EQ>(v1.x, key) << 4 | (unsigned)before<LESS_EQ>(v1.y, key) << 5 |
           (unsigned)before<LESS_EQ>(v1.z, key) << 6 | (unsigned)before<LESS_EQ>(v1.w, key) << 7 |
           (unsigned)before<LESS_EQ>(v2.x, key) << 8 | (unsigned)before<LESS_EQ>(v2.y, key) << 9 |
           (unsigned)before<LESS_EQ>(v2.z, key) << 10 | (unsigned)before<LESS_EQ>(v2.w, key) << 11 |
           (unsigned)before<LESS_EQ>(v3.x, key) << 12 | (unsigned)before<LESS_EQ>(v3.y, key) << 13 |
           (unsigned)before<LESS_EQ>(v3.z, key) << 14 | (unsigned)before<LESS_EQ>(v3.w, key) << 15;
}

BXG_HD int popc32(unsigned x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
BXG_HD int ffs32(unsigned x) {   // 1-based index of the lowest set bit
#if defined(__CUDA_ARCH__)
    return __ffs((int)x);
#else
    return __builtin_ffs((int)x);
#endif
}

// entries are sorted inside the segment, so the before-entries among the valid samples form a prefix
BXG_HD void round_apply(Win &w, const Round &r, int ss, unsigned mask) {
    const unsigned valid = ((1u << (r.m1 - r.g)) - 1u) & ~((1u << (r.m0 - r.g)) - 1u);
    const uint32_t c = (uint32_t)popc32(mask & valid);
    if (c == 0) {
        w.hi = r.m0 << ss;
    } else {
        w.lo = ((r.m0 + c - 1) << ss) + 1;
        if (r.m0 + c < r.m1) w.hi = (r.m0 + c) << ss;
    }
}

// Plain binary search on level 0: the safety net when a window does not fit one aligned group (cannot happen for
// the aligned windows the splitter phase produces) and the finisher used by the host-only callers.
// `mul` is the level-0 group pitch in 16-int units: 1 for a plain array, 2 when two arrays are interleaved group by
// group ([S x16 | PM x16] ..., [E x16 | I x16] ...) so that the lines the two searches / the walk and the emitter touch
// are adjacent in HBM.
BXG_HD const int32_t *group_ptr(const int32_t *A, uint32_t g, int mul) { return A + (size_t)g * (size_t)mul; }
BXG_HD const int32_t *elem_ptr(const int32_t *A, uint32_t m, int mul) {
    return A + (size_t)(m & ~15u) * (size_t)mul + (m & 15u);
}

template <bool LESS_EQ, typename LD>
BXG_HD uint32_t finish_binary(const int32_t *A, Win w, int32_t key, const LD &ld, int mul = 1) {
    while (w.lo < w.hi) {
        uint32_t m = (w.lo + w.hi) >> 1;
        if (before<LESS_EQ>(ld(elem_ptr(A, m, mul)), key)) w.lo = m + 1; else w.hi = m;
    }
    return w.lo;
}

// Both searches of one query, in lock-step.  SP: shared-memory splitter arrays; KS/KP: sampled levels (K*[0] = S / PM,
// padded to 16-entry groups with INT32_MAX); ld4(ptr, a, b, c, d) loads one aligned 64-byte group (16 entries).
struct NoPrefetch {
    BXG_HD void operator()(uint32_t) const {}
};

// pf(g) is called with the 16-aligned positions the two searches are about to resolve (final round): the walk that
// follows reads E (and the fill reads I) at exactly those groups, so the kernels prefetch them one DRAM latency early.
template <typename SP, typename LD4, typename LD, typename PF = NoPrefetch>
BXG_HD void dual_search(const int32_t *const *KS, const int32_t *const *KP, int nk, const SP &spS, const SP &spPM,
                        int shift, uint32_t seg_lo, uint32_t seg_hi, int32_t qe, int32_t qs, const LD4 &ld4,
                        const LD &ld, uint32_t &hi_out, uint32_t &lo_out, const PF &pf = PF(), int mul0 = 1,
                        bool coarse_lo = false) {
    // coarse_lo: stop the PM search one round early and return the START of its last (<= 16-item) window instead of
    // the exact first hit.  Every item in front of the true `lo` has running max <= qs, hence E <= qs, and the walk
    // masks it out by itself -- so the hit lists are identical, while the 4 B/item PM array is never touched (one
    // DRAM line per query less, and 1/3 less working set competing for L2).
    if (seg_lo >= seg_hi) {
        hi_out = lo_out = seg_hi;
        return;
    }
    const uint32_t k0 = (seg_lo + (1u << shift) - 1) >> shift, k1 = ((seg_hi - 1) >> shift) + 1;
    uint32_t a_s = k0, a_p = k0;
    if (k0 < k1) {
        uint32_t step = 1;
        while (step * 2 <= k1 - k0) step *= 2;
        for (; step > 0; step >>= 1) {               // same trip count for both searches: the loads interleave
            a_s = lift_step<false>(spS, a_s, step, k1, qe);
            a_p = lift_step<true>(spPM, a_p, step, k1, qs);
        }
    }
    Win ws = splitter_window(a_s, k0, k1, shift, seg_lo, seg_hi);
    Win wp = splitter_window(a_p, k0, k1, shift, seg_lo, seg_hi);
    for (int j = nk - 1; j >= 0; j--) {
        const int ss = 4 * j;
        const Round rs = round_prepare(ws, ss);
        Round rp = round_prepare(wp, ss);
        if (coarse_lo && j == 0) rp.active = false;
        int4 s0{}, s1{}, s2{}, s3{}, p0{}, p1{}, p2{}, p3{};
        if (j == 0) {
            if (rp.active) pf(rp.g);
            if (rs.active && (!rp.active || rs.g != rp.g)) pf(rs.g);
        }
        if (rs.active) {
            const int4 *p = reinterpret_cast<const int4 *>(j ? KS[j] + rs.g : group_ptr(KS[0], rs.g, mul0));
            ld4(p, s0, s1, s2, s3);
        }
        if (rp.active) {
            const int4 *p = reinterpret_cast<const int4 *>(j ? KP[j] + rp.g : group_ptr(KP[0], rp.g, mul0));
            ld4(p, p0, p1, p2, p3);
        }
        if (rs.active) round_apply(ws, rs, ss, group_mask<false>(s0, s1, s2, s3, qe));
        if (rp.active) round_apply(wp, rp, ss, group_mask<true>(p0, p1, p2, p3, qs));
    }
    hi_out = finish_binary<false>(KS[0], ws, qe, ld, mul0);   // no-ops when the rounds converged (lo == hi)
    lo_out = coarse_lo ? wp.lo : finish_binary<true>(KP[0], wp, qs, ld, mul0);
}

// Walk [lo,hi) in aligned 16-item groups of E (padded with INT32_MIN): f(k0, mask) gets the bit mask of hits
// (bit i <=> item k0+i has E > qs and lies in [lo,hi)).  After an empty group, 32-aligned all-miss blocks are skipped
// through the max hierarchy M[l] (M[l][b] = max E over 32^(l+1) items) -- O(32 log n) per hit in the worst case.
template <typename LD4, typename LD, typename F, typename PF = NoPrefetch>
BXG_HD void walk_hits(const int32_t *E, const int32_t *const *M, int nlev, uint32_t lo, uint32_t hi, int32_t qs,
                      const LD4 &ld4, const LD &ld, F &&f, const PF &pf = PF(), int mul = 1,
                      uint32_t k_start = 0xffffffffu, bool prev_empty0 = false) {
    if (lo >= hi) return;
    uint32_t k = k_start == 0xffffffffu ? (lo & ~15u) : k_start;   // resume point of a walk whose first group is done
    bool prev_empty = prev_empty0;
    while (k < hi) {
        if (prev_empty && (k & 31u) == 0 && k + 32u <= hi && ld(M[0] + (k >> 5)) <= qs) {
            uint32_t idx = k >> 5;
            int lvl = 0;
            while (lvl + 1 < nlev && (idx & 31u) == 0) {
                const uint32_t up = idx >> 5;
                const uint64_t span_end = ((uint64_t)up + 1) << (5 * (lvl + 2));
                if (span_end > hi || ld(M[lvl + 1] + up) > qs) break;
                idx = up;
                lvl++;
            }
            k = (idx + 1u) << (5 * (lvl + 1));
            continue;
        }
        pf(k);          // the emitter will read the same group of I: start that fetch together with the E loads
        const int4 *p = reinterpret_cast<const int4 *>(group_ptr(E, k, mul));
        int4 v0, v1, v2, v3;
        ld4(p, v0, v1, v2, v3);
        unsigned mask = 0xffffu & ~group_mask<true>(v0, v1, v2, v3, qs);      // E > qs
        if (k < lo) mask &= ~0u << (lo - k);
        if (k + 16u > hi) mask &= (1u << (hi - k)) - 1u;
        prev_empty = mask == 0;
        if (mask) f(k, mask);
        k += 16;
    }
}

// Search + walk of one query with the LAST search round and the FIRST walk group overlapped.
//
// After the sampled levels (j >= 1) the PM search has its coarse `lo` (see dual_search: coarse_lo), i.e. the walk's
// first E group is already known while the S search still needs its final round on S itself.  Both loads (two DRAM
// lines) are issued together, so the dependent chain of a typical query is
//     splitters (smem) -> level 2 -> level 1 -> { S group || E group } -> (second E group if the span needs it)
// instead of ... -> S group -> E group -> E group.  Plain (non-interleaved) level-0 arrays only.
template <typename SP, typename LD4, typename LD, typename F>
BXG_HD void search_walk(const int32_t *const *KS, const int32_t *const *KP, int nk, const SP &spS, const SP &spPM,
                        int shift, uint32_t seg_lo, uint32_t seg_hi, int32_t qe, int32_t qs, const int32_t *E,
                        const int32_t *const *M, int nlev, const LD4 &ld4, const LD &ld, uint32_t &hi_out,
                        uint32_t &lo_out, F &&f) {
    hi_out = lo_out = seg_hi;
    if (seg_lo >= seg_hi) return;
    const uint32_t k0 = (seg_lo + (1u << shift) - 1) >> shift, k1 = ((seg_hi - 1) >> shift) + 1;
    uint32_t a_s = k0, a_p = k0;
    if (k0 < k1) {
        uint32_t step = 1;
        while (step * 2 <= k1 - k0) step *= 2;
        for (; step > 0; step >>= 1) {
            a_s = lift_step<false>(spS, a_s, step, k1, qe);
            a_p = lift_step<true>(spPM, a_p, step, k1, qs);
        }
    }
    Win ws = splitter_window(a_s, k0, k1, shift, seg_lo, seg_hi);
    Win wp = splitter_window(a_p, k0, k1, shift, seg_lo, seg_hi);
    for (int j = nk - 1; j >= 1; j--) {
        const int ss = 4 * j;
        const Round rs = round_prepare(ws, ss), rp = round_prepare(wp, ss);
        int4 s0{}, s1{}, s2{}, s3{}, p0{}, p1{}, p2{}, p3{};
        if (rs.active) {
            const int4 *p = reinterpret_cast<const int4 *>(KS[j] + rs.g);
            ld4(p, s0, s1, s2, s3);
        }
        if (rp.active) {
            const int4 *p = reinterpret_cast<const int4 *>(KP[j] + rp.g);
            ld4(p, p0, p1, p2, p3);
        }
        if (rs.active) round_apply(ws, rs, ss, group_mask<false>(s0, s1, s2, s3, qe));
        if (rp.active) round_apply(wp, rp, ss, group_mask<true>(p0, p1, p2, p3, qs));
    }
    // final S round and first E group together
    const uint32_t lo_c = wp.lo;                       // coarse lo: <= the true lo, possibly > the true hi
    const uint32_t g0 = lo_c & ~15u;
    const Round rs = round_prepare(ws, 0);
    int4 s0{}, s1{}, s2{}, s3{}, e0{}, e1{}, e2{}, e3{};
    if (rs.active) {
        const int4 *p = reinterpret_cast<const int4 *>(KS[0] + rs.g);
        ld4(p, s0, s1, s2, s3);
    }
    const bool spec = lo_c < seg_hi;
    if (spec) {
        const int4 *p = reinterpret_cast<const int4 *>(E + g0);
        ld4(p, e0, e1, e2, e3);
    }
    if (rs.active) round_apply(ws, rs, 0, group_mask<false>(s0, s1, s2, s3, qe));
    const uint32_t hi = finish_binary<false>(KS[0], ws, qe, ld);
    const uint32_t lo = lo_c < hi ? lo_c : hi;
    hi_out = hi;
    lo_out = lo;
    if (lo >= hi) return;                              // (then lo == lo_c and g0 == lo & ~15)
    unsigned mask = 0xffffu & ~group_mask<true>(e0, e1, e2, e3, qs);
    if (g0 < lo) mask &= ~0u << (lo - g0);
    if (g0 + 16u > hi) mask &= (1u << (hi - g0)) - 1u;
    if (mask) f(g0, mask);
    if (g0 + 16u < hi) walk_hits(E, M, nlev, lo, hi, qs, ld4, ld, f, NoPrefetch(), 1, g0 + 16u, mask == 0);
}

// ---- search + walk, second form: ONE search, then a backward probe for the walk's start ---------------------------
//
// The count kernels are bound by L1 data-pipe wavefronts -- one per 32-byte sector a lane touches (ncu, profiles/).
// search_walk spends 13.3 sectors per query: three 64-byte rounds for `hi` (6), two for the coarse `lo` (4), and the E
// groups of the walk (~3.3).  But `lo` is almost always within a few groups of `hi`: a query reaches back at most by
// the length of the longest interval still "open" there.  So instead of searching PM from the top, look at the sampled
// running maxima right in front of hi: KP[1][j] = PM[16 j].  If PM[16 j] <= qs, every item up to 16 j ends at or
// before qs, i.e. the true lo is beyond 16 j and 16 j is a valid (coarse, aligned) start for the walk.  One 32-byte
// sector of KP[1] holds 8 such samples = 128 items of reach, and its address is known BEFORE the last S round
// resolves (hi - 1 lies in the <= 16-item window that round looks at), so the probe is issued together with the
// final S group.  If no sample of the sector qualifies the probe steps back one sector at a time (twice), then falls
// back to the full PM search -- long intervals in front pay that, short-interval data never does.
// Sectors per query: 6 (S) + 1 (probe) + the walk; the walk itself loads only the 32-byte halves of an E group that
// intersect [lo, hi) (walk_hits_halves).
template <typename LD8>
BXG_HD unsigned sector_mask_le(const int32_t *p, int32_t key, const LD8 &ld8) {
    int4 a, b;
    ld8(reinterpret_cast<const int4 *>(p), a, b);
    return (unsigned)(a.x <= key) | (unsigned)(a.y <= key) << 1 | (unsigned)(a.z <= key) << 2 | (unsigned)(a.w <= key) << 3 |
           (unsigned)(b.x <= key) << 4 | (unsigned)(b.y <= key) << 5 | (unsigned)(b.z <= key) << 6 | (unsigned)(b.w <= key) << 7;
}

BXG_HD int fls8(unsigned x) {   // 0-based index of the highest set bit of a non-zero 8-bit mask
    int i = 0;
    if (x & 0xf0u) { i += 4; x >>= 4; }
    if (x & 0x0cu) { i += 2; x >>= 2; }
    if (x & 0x02u) i += 1;
    return i;
}

// walk_hits with half-group loads: of each aligned 16-item group only the 8-item halves that intersect [lo, hi) are
// read (one 32-byte sector each); a half that is not read contributes no hits.
// E > qs mask of the halves of group k that intersect [lo, hi) (a half that is not read contributes no hits)
template <typename LD8>
BXG_HD unsigned group_hits_halves(const int32_t *E, uint32_t k, uint32_t lo, uint32_t hi, int32_t qs, const LD8 &ld8) {
    unsigned mask = 0;
    if (lo < k + 8u) mask |= 0xffu & ~sector_mask_le(E + k, qs, ld8);                             // items k..k+7
    if (hi > k + 8u && lo < k + 16u) mask |= (0xffu & ~sector_mask_le(E + k + 8, qs, ld8)) << 8;
    if (k < lo) mask &= ~0u << (lo - k);
    if (k + 16u > hi) mask &= (1u << (hi - k)) - 1u;
    return mask;
}

template <typename LD8, typename LD, typename F>
BXG_HD void walk_hits_halves(const int32_t *E, const int32_t *const *M, int nlev, uint32_t lo, uint32_t hi, int32_t qs,
                             const LD8 &ld8, const LD &ld, F &&f) {
    if (lo >= hi) return;
    uint32_t k = lo & ~15u;
    bool prev_empty = false;
#ifdef BXS_WALK_PREFETCH
    // Most walks span two groups.  The loop below decides about group 2 only after group 1's mask is known (a branch
    // on loaded data), i.e. two dependent memory latencies; issuing both groups' loads up front makes it one.
    // Measured SLOWER on B200 (count kernel 0.427 -> 0.460 ms, profiles/r01z): kept as an opt-in experiment.
    if (k + 16u < hi) {
        const unsigned m0 = group_hits_halves(E, k, lo, hi, qs, ld8);
        const unsigned m1 = group_hits_halves(E, k + 16u, lo, hi, qs, ld8);
        if (m0) f(k, m0);
        if (m1) f(k + 16u, m1);
        prev_empty = m1 == 0;
        k += 32u;
    }
#endif
    while (k < hi) {
        if (prev_empty && (k & 31u) == 0 && k + 32u <= hi && ld(M[0] + (k >> 5)) <= qs) {
            uint32_t idx = k >> 5;
            int lvl = 0;
            while (lvl + 1 < nlev && (idx & 31u) == 0) {
                const uint32_t up = idx >> 5;
                const uint64_t span_end = ((uint64_t)up + 1) << (5 * (lvl + 2));
                if (span_end > hi || ld(M[lvl + 1] + up) > qs) break;
                idx = up;
                lvl++;
            }
            k = (idx + 1u) << (5 * (lvl + 1));
            continue;
        }
        const unsigned mask = group_hits_halves(E, k, lo, hi, qs, ld8);
        prev_empty = mask == 0;
        if (mask) f(k, mask);
        k += 16;
    }
}

template <typename SP, typename LD4, typename LD8, typename LD, typename F>
BXG_HD void search_walk_probe(const int32_t *const *KS, const int32_t *const *KP, int nk, const SP &spS, const SP &spPM,
                              int shift, uint32_t seg_lo, uint32_t seg_hi, int32_t qe, int32_t qs, const int32_t *E,
                              const int32_t *const *M, int nlev, const LD4 &ld4, const LD8 &ld8, const LD &ld,
                              uint32_t &hi_out, uint32_t &lo_out, F &&f) {
    if (nk < 2) {                                      // small index: no sampled level to probe
        search_walk(KS, KP, nk, spS, spPM, shift, seg_lo, seg_hi, qe, qs, E, M, nlev, ld4, ld, hi_out, lo_out, f);
        return;
    }
    hi_out = lo_out = seg_hi;
    if (seg_lo >= seg_hi) return;
    const uint32_t k0 = (seg_lo + (1u << shift) - 1) >> shift, k1 = ((seg_hi - 1) >> shift) + 1;
    uint32_t a_s = k0;
    uint32_t step0 = 0;
    if (k0 < k1) {
        step0 = 1;
        while (step0 * 2 <= k1 - k0) step0 *= 2;
        for (uint32_t step = step0; step > 0; step >>= 1) a_s = lift_step<false>(spS, a_s, step, k1, qe);
    }
    Win ws = splitter_window(a_s, k0, k1, shift, seg_lo, seg_hi);
    for (int j = nk - 1; j >= 1; j--) {
        const int ss = 4 * j;
        const Round rs = round_prepare(ws, ss);
        if (rs.active) {
            int4 s0, s1, s2, s3;
            ld4(reinterpret_cast<const int4 *>(KS[j] + rs.g), s0, s1, s2, s3);
            round_apply(ws, rs, ss, group_mask<false>(s0, s1, s2, s3, qe));
        }
    }
    // final S round; the probe sector is chosen from the window's upper end (hi - 1 <= ws.hi - 1) and loaded with it
    const Round rs = round_prepare(ws, 0);
    uint32_t sb = ((ws.hi ? ws.hi - 1u : 0u) >> 4) & ~7u;
    int4 s0{}, s1{}, s2{}, s3{};
    if (rs.active) ld4(reinterpret_cast<const int4 *>(KS[0] + rs.g), s0, s1, s2, s3);
    unsigned pm_le = sector_mask_le(KP[1] + sb, qs, ld8);
    if (rs.active) round_apply(ws, rs, 0, group_mask<false>(s0, s1, s2, s3, qe));
    const uint32_t hi = finish_binary<false>(KS[0], ws, qe, ld);
    hi_out = lo_out = hi;
    if (hi <= seg_lo) return;                          // no item starts before qe
    const uint32_t m_hi = (hi - 1u) >> 4;              // group of the last candidate
    if (m_hi < sb) {                                   // hi resolved into the group in front of the probed sector
        sb = m_hi & ~7u;
        pm_le = sector_mask_le(KP[1] + sb, qs, ld8);
    }
    uint32_t lo_c = 0;
    bool found = false;
    for (int tries = 0; tries < 3; tries++) {
        // samples of this sector that belong to this tree (16 j >= seg_lo) and are not past the last candidate
        unsigned valid = 0xffu;
        if (m_hi < sb + 7u) valid &= (2u << (m_hi - sb)) - 1u;
        const uint32_t jmin = (seg_lo + 15u) >> 4;     // first sample position inside the segment
        if (jmin > sb) valid &= (jmin - sb >= 8u) ? 0u : (~0u << (jmin - sb));
        const unsigned ok = pm_le & valid;
        if (ok) {
            lo_c = (sb + (uint32_t)fls8(ok)) << 4;
            found = true;
            break;
        }
        if (jmin >= sb) {                              // ran off the front of the tree: everything before is another tree
            lo_c = seg_lo;
            found = true;
            break;
        }
        if (tries == 2) break;
        sb -= 8u;
        pm_le = sector_mask_le(KP[1] + sb, qs, ld8);
    }
    if (!found) {                                      // something long is open in front: the full PM search (coarse)
        uint32_t a_p = k0;
        for (uint32_t step = step0; step > 0; step >>= 1) a_p = lift_step<true>(spPM, a_p, step, k1, qs);
        Win wp = splitter_window(a_p, k0, k1, shift, seg_lo, seg_hi);
        for (int j = nk - 1; j >= 1; j--) {
            const int ss = 4 * j;
            const Round rp = round_prepare(wp, ss);
            if (rp.active) {
                int4 p0, p1, p2, p3;
                ld4(reinterpret_cast<const int4 *>(KP[j] + rp.g), p0, p1, p2, p3);
                round_apply(wp, rp, ss, group_mask<true>(p0, p1, p2, p3, qs));
            }
        }
        lo_c = wp.lo;
    }
    const uint32_t lo = lo_c < hi ? lo_c : hi;
    lo_out = lo;
    walk_hits_halves(E, M, nlev, lo, hi, qs, ld8, ld, f);
}

// ---- third form: the same single search + backward probe over 8-ARY levels --------------------------------------
//
// A 16-entry round costs two 32-byte sectors and resolves 4 bits; an 8-entry round costs one and resolves 3.  With the
// count kernels bound by sectors, the levels Q[j][i] = A[i << 3j] (strides 8, 64, 512, ...) take the window from the
// splitters (2^shift items) down to one item in ceil(shift / 3) single-sector rounds: 4 sectors for shift = 12 where
// the 16-ary levels need 6.  The probe reads one sector of QP[1] (8 samples of the running max = 64 items of reach).
BXG_HD Round round_prepare8(const Win &w, int ss) {
    Round r{0, 0, 0, false};
    if (w.lo >= w.hi) return r;
    r.m0 = (w.lo + (1u << ss) - 1) >> ss;
    r.m1 = ((w.hi - 1) >> ss) + 1;
    r.g = r.m0 & ~7u;
    r.active = r.m0 < r.m1 && r.m1 - r.g <= 8u;
    return r;
}

template <bool LESS_EQ, typename LD8>
BXG_HD unsigned sector_mask(const int32_t *p, int32_t key, const LD8 &ld8) {
    int4 a, b;
    ld8(reinterpret_cast<const int4 *>(p), a, b);
    return (unsigned)before<LESS_EQ>(a.x, key) | (unsigned)before<LESS_EQ>(a.y, key) << 1 |
           (unsigned)before<LESS_EQ>(a.z, key) << 2 | (unsigned)before<LESS_EQ>(a.w, key) << 3 |
           (unsigned)before<LESS_EQ>(b.x, key) << 4 | (unsigned)before<LESS_EQ>(b.y, key) << 5 |
           (unsigned)before<LESS_EQ>(b.z, key) << 6 | (unsigned)before<LESS_EQ>(b.w, key) << 7;
}

// one search over the 8-ary levels Q[n8-1] .. Q[0] (Q[0] = the array itself), starting from a splitter window
// ld8_top loads the top level (a few tens of KB that every query reads: worth keeping in L1), ld8 the others
template <bool LESS_EQ, typename LD8T, typename LD8>
BXG_HD void rounds8(const int32_t *const *Q, int n8, Win &w, int32_t key, const LD8T &ld8_top, const LD8 &ld8,
                    int last_level) {
    for (int j = n8 - 1; j >= last_level; j--) {
        const int ss = 3 * j;
        const Round r = round_prepare8(w, ss);
        if (!r.active) continue;
        const unsigned m = (j == n8 - 1 && j >= 2) ? sector_mask<LESS_EQ>(Q[j] + r.g, key, ld8_top)
                                                   : sector_mask<LESS_EQ>(Q[j] + r.g, key, ld8);
        round_apply(w, r, ss, m);
    }
}

template <typename SP, typename LD8T, typename LD8, typename LD, typename F>
BXG_HD void search_walk_probe8(const int32_t *const *QS, const int32_t *const *QP, int n8, const SP &spS, const SP &spPM,
                               int shift, uint32_t seg_lo, uint32_t seg_hi, int32_t qe, int32_t qs, const int32_t *E,
                               const int32_t *const *M, int nlev, const LD8T &ld8_top, const LD8 &ld8, const LD &ld,
                               uint32_t &hi_out, uint32_t &lo_out, F &&f) {
    hi_out = lo_out = seg_hi;
    if (seg_lo >= seg_hi) return;
    const uint32_t k0 = (seg_lo + (1u << shift) - 1) >> shift, k1 = ((seg_hi - 1) >> shift) + 1;
    uint32_t a_s = k0, step0 = 0;
    if (k0 < k1) {
        step0 = 1;
        while (step0 * 2 <= k1 - k0) step0 *= 2;
        for (uint32_t step = step0; step > 0; step >>= 1) a_s = lift_step<false>(spS, a_s, step, k1, qe);
    }
    Win ws = splitter_window(a_s, k0, k1, shift, seg_lo, seg_hi);
    rounds8<false>(QS, n8, ws, qe, ld8_top, ld8, 1);
    // final S round (one sector of S) together with the probe sector of QP[1]; without a sampled level (tiny index) the
    // walk simply starts at the front of the segment
    const Round rs = round_prepare8(ws, 0);
    uint32_t sb = ((ws.hi ? ws.hi - 1u : 0u) >> 3) & ~7u;
    unsigned s_lt = 0;
    if (rs.active) s_lt = sector_mask<false>(QS[0] + rs.g, qe, ld8);
    unsigned pm_le = n8 >= 2 ? sector_mask<true>(QP[1] + sb, qs, ld8) : 0u;
    if (rs.active) round_apply(ws, rs, 0, s_lt);
    const uint32_t hi = finish_binary<false>(QS[0], ws, qe, ld);
    hi_out = lo_out = hi;
    if (hi <= seg_lo) return;                          // no item starts before qe
    uint32_t lo_c = seg_lo;
    bool found = false;
    if (n8 >= 2) {
        const uint32_t m_hi = (hi - 1u) >> 3;          // sample block of the last candidate
        if (m_hi < sb) {
            sb = m_hi & ~7u;
            pm_le = sector_mask<true>(QP[1] + sb, qs, ld8);
        }
        const uint32_t jmin = (seg_lo + 7u) >> 3;      // first sample position inside the segment
        for (int tries = 0; tries < 3; tries++) {
            unsigned valid = 0xffu;
            if (m_hi < sb + 7u) valid &= (2u << (m_hi - sb)) - 1u;
            if (jmin > sb) valid &= (jmin - sb >= 8u) ? 0u : (~0u << (jmin - sb));
            const unsigned ok = pm_le & valid;
            if (ok) {
                lo_c = (sb + (uint32_t)fls8(ok)) << 3;
                found = true;
                break;
            }
            if (jmin >= sb) {                          // ran off the front of the tree
                lo_c = seg_lo;
                found = true;
                break;
            }
            if (tries == 2) break;
            sb -= 8u;
            pm_le = sector_mask<true>(QP[1] + sb, qs, ld8);
        }
    }
    if (!found) {   // something long is open in front, or an index too small for sampled levels: the PM search (coarse)
        uint32_t a_p = k0;
        for (uint32_t step = step0; step > 0; step >>= 1) a_p = lift_step<true>(spPM, a_p, step, k1, qs);
        Win wp = splitter_window(a_p, k0, k1, shift, seg_lo, seg_hi);
        rounds8<true>(QP, n8, wp, qs, ld8, ld8, 1);
        lo_c = wp.lo;
    }
    const uint32_t lo = lo_c < hi ? lo_c : hi;
    lo_out = lo;
    walk_hits_halves(E, M, nlev, lo, hi, qs, ld8, ld, f);
}

// ---- fourth form: direct addressing ------------------------------------------------------------------------------
//
// The three forms above SEARCH for hi: splitters in shared memory, then 3-4 dependent single-sector rounds.  A count
// kernel bound by dependent sector reads wants neither.  Here every tree carries a uniform grid over its start
// coordinates -- cell c covers [base + (c << shift), base + ((c + 1) << shift)), 2-4 items per cell on average -- and per
// cell one 8-byte record
//     G[c].x = first k with S[k]  >= cell_start(c)          (where the candidates of a query ENDING in this cell stop)
//     G[c].y = first k with PM[k] >  cell_start(c)          (where the hits of a query STARTING in this cell can begin)
// so that     hi in [G[ce].x, G[ce + 1].x]   for ce = cell(qe)   -- resolved by the one or two S sectors of that cell,
//             lo >= G[cs].y                  for cs = cell(qs)   -- a coarse start, the walk masks the rest (E <= qs).
// Dependent chain per query: cell record(s) -> S sector -> E sectors: 3 rounds instead of 5, ~5.5 sectors instead of
// 7.4, and no shared-memory table at all.  Any distribution stays correct: a crowded cell falls back to a binary search
// inside [G[ce].x, G[ce+1].x], a long interval in front only lengthens the walk (which skips through M[]).
struct GridDir {
    int32_t base;        // smallest start of the tree
    int32_t shift;       // cell width = 1 << shift
    uint32_t ncells;     // cells 0 .. ncells (ncells + 1 records; record ncells is the end sentinel)
    uint32_t coff;       // index of the tree's record 0 in G
};
struct GridRec {
    uint32_t x, y;
};

// cell of coordinate q (q > base), clamped to ncells
BXG_HD uint32_t grid_cell(const GridDir &d, int32_t q) {
    const uint32_t c = ((uint32_t)q - (uint32_t)d.base) >> d.shift;
    return c < d.ncells ? c : d.ncells;
}

// a + #{k in [a,b) : S[k] < qe}; S is sorted inside the segment, so those form a prefix of [a,b): one or two S sectors
// when the candidates fit them, a binary search otherwise (a crowded cell)
template <typename LD8, typename LD>
BXG_HD uint32_t grid_resolve_hi(const int32_t *S, uint32_t a, uint32_t b, int32_t qe, const LD8 &ld8, const LD &ld) {
    if (b <= a) return a;
    const uint32_t s0 = a & ~7u, s1 = (b - 1u) & ~7u;
    if (s1 - s0 <= 8u) {
        unsigned m = sector_mask<false>(S + s0, qe, ld8);
        if (s1 != s0) m |= sector_mask<false>(S + s1, qe, ld8) << 8;
        m &= ~0u << (a - s0);
        if (b - s0 < 32u) m &= (1u << (b - s0)) - 1u;
        return a + (uint32_t)popc32(m);
    }
    Win w{a, b};
    return finish_binary<false>(S, w, qe, ld);
}

template <typename LDR, typename LD8, typename LD, typename F>
BXG_HD void search_walk_grid(const GridRec *G, const GridDir &d, const int32_t *S, uint32_t seg_lo, uint32_t seg_hi,
                             int32_t qe, int32_t qs, const int32_t *E, const int32_t *const *M, int nlev, const LDR &ldr,
                             const LD8 &ld8, const LD &ld, uint32_t &hi_out, uint32_t &lo_out, F &&f) {
    hi_out = lo_out = seg_hi;
    if (seg_lo >= seg_hi) return;
    if (qe <= d.base) {                                // nothing starts before qe
        hi_out = lo_out = seg_lo;
        return;
    }
    const GridRec *g = G + d.coff;
    const uint32_t ce = grid_cell(d, qe);
    uint32_t a, b, lo_c = seg_lo;
    {
        const GridRec r0 = ldr(g + ce);
        a = r0.x;
        b = ce < d.ncells ? ldr(g + ce + 1).x : seg_hi;
        if (qs >= d.base) {
            const uint32_t cs = grid_cell(d, qs);
            lo_c = cs == ce ? r0.y : ldr(g + cs).y;
        }
    }
    const uint32_t hi = grid_resolve_hi(S, a, b, qe, ld8, ld);
    hi_out = hi;
    const uint32_t lo = lo_c < hi ? lo_c : hi;
    lo_out = lo;
    walk_hits_halves(E, M, nlev, lo, hi, qs, ld8, ld, f);
}

// ---- fifth form: direct addressing with the cell's starts packed into its record ----------------------------------------
//
// The grid form still reads S to place qe among the 2-4 items of its cell.  A 16-byte record has room for them:
//     x, y      as above
//     p         bits 0..7: items in the cell (saturating at 255); then up to K = 56 / shift fields of `shift` bits, the
//               offsets S[x + i] - cell_start(c) of the cell's first K items
// so hi = x + #{stored offsets < qe - cell_start} with NO access to S unless the cell holds more items than fit (then the
// remainder is resolved from S as before).  One 16-byte record per query end (+ 4 bytes of the start's record, same
// sector half the time) instead of record + neighbour + S sector: ~4 sectors per query instead of ~5.5, and S (4 bytes per
// item) drops out of the working set that competes for L2.
struct GridRec16 {
    uint32_t x, y, p0, p1;
};

BXG_HD int grid16_fields(int shift) { return shift > 0 ? 56 / shift : 0; }

template <typename LDR16, typename LD8, typename LD, typename F>
BXG_HD void search_walk_grid16(const GridRec16 *G, const GridDir &d, const int32_t *S, uint32_t seg_lo, uint32_t seg_hi,
                               int32_t qe, int32_t qs, const int32_t *E, const int32_t *const *M, int nlev,
                               const LDR16 &ldr, const LD8 &ld8, const LD &ld, uint32_t &hi_out, uint32_t &lo_out, F &&f) {
    hi_out = lo_out = seg_hi;
    if (seg_lo >= seg_hi) return;
    if (qe <= d.base) {
        hi_out = lo_out = seg_lo;
        return;
    }
    const GridRec16 *g = G + d.coff;
    const uint32_t ce = grid_cell(d, qe);
    const GridRec16 r = ldr(g + ce);
    uint32_t lo_c = seg_lo;
    if (qs >= d.base) {
        const uint32_t cs = grid_cell(d, qs);
        lo_c = cs == ce ? r.y : (uint32_t)ld(reinterpret_cast<const int32_t *>(&g[cs].y));
    }
    uint32_t hi = r.x;
    if (ce < d.ncells && d.shift > 0) {                // (shift 0: every item of the cell starts AT qe -> none before it)
        const uint64_t p = (uint64_t)r.p0 | ((uint64_t)r.p1 << 32);
        const uint32_t cnt = (uint32_t)(p & 255u);
        const uint32_t dq = ((uint32_t)qe - (uint32_t)d.base) - (ce << d.shift);
        const uint32_t kmax = (uint32_t)grid16_fields(d.shift), stored = cnt < kmax ? cnt : kmax;
        const uint64_t fmask = (1ull << d.shift) - 1ull;
        uint32_t below = 0;
        for (uint32_t i = 0; i < stored; i++) below += (uint32_t)(((p >> (8 + i * d.shift)) & fmask) < dq);
        hi = r.x + below;
        if (below == stored && cnt > stored) {         // more items than the record describes: the rest from S
            const uint32_t b = cnt < 255u ? r.x + cnt : (uint32_t)ld(reinterpret_cast<const int32_t *>(&g[ce + 1].x));
            hi = grid_resolve_hi(S, r.x + stored, b, qe, ld8, ld);
        }
    }
    hi_out = hi;
    const uint32_t lo = lo_c < hi ? lo_c : hi;
    lo_out = lo;
    walk_hits_halves(E, M, nlev, lo, hi, qs, ld8, ld, f);
}

// Write the item ids of the hits of one 16-item group: the whole aligned group of I is fetched with four 16-byte
// loads issued back to back (one 64-byte line), THEN the selected ids are stored -- a load per hit interleaved with
// the stores would serialise on every store (the compiler must assume hits[] may alias I[]).
template <typename LD4, typename SINK>
BXG_HD void emit_group_to(const int32_t *I, uint32_t k0, unsigned mask, SINK &out, const LD4 &ld4, int mul = 1) {
    const int4 *p = reinterpret_cast<const int4 *>(group_ptr(I, k0, mul));
    int4 a, b, c, d;
    ld4(p, a, b, c, d);
    if (mask & 0x0001u) out.put(a.x);
    if (mask & 0x0002u) out.put(a.y);
    if (mask & 0x0004u) out.put(a.z);
    if (mask & 0x0008u) out.put(a.w);
    if (mask & 0x0010u) out.put(b.x);
    if (mask & 0x0020u) out.put(b.y);
    if (mask & 0x0040u) out.put(b.z);
    if (mask & 0x0080u) out.put(b.w);
    if (mask & 0x0100u) out.put(c.x);
    if (mask & 0x0200u) out.put(c.y);
    if (mask & 0x0400u) out.put(c.z);
    if (mask & 0x0800u) out.put(c.w);
    if (mask & 0x1000u) out.put(d.x);
    if (mask & 0x2000u) out.put(d.y);
    if (mask & 0x4000u) out.put(d.z);
    if (mask & 0x8000u) out.put(d.w);
}

// The same with half-group loads: only the 32-byte halves of the I group that hold a selected item are read.
template <typename LD8, typename SINK>
BXG_HD void emit_group_halves_to(const int32_t *I, uint32_t k0, unsigned mask, SINK &out, const LD8 &ld8) {
    int4 a{}, b{}, c{}, d{};
    if (mask & 0x00ffu) ld8(reinterpret_cast<const int4 *>(I + k0), a, b);
    if (mask & 0xff00u) ld8(reinterpret_cast<const int4 *>(I + k0 + 8), c, d);
    if (mask & 0x0001u) out.put(a.x);
    if (mask & 0x0002u) out.put(a.y);
    if (mask & 0x0004u) out.put(a.z);
    if (mask & 0x0008u) out.put(a.w);
    if (mask & 0x0010u) out.put(b.x);
    if (mask & 0x0020u) out.put(b.y);
    if (mask & 0x0040u) out.put(b.z);
    if (mask & 0x0080u) out.put(b.w);
    if (mask & 0x0100u) out.put(c.x);
    if (mask & 0x0200u) out.put(c.y);
    if (mask & 0x0400u) out.put(c.z);
    if (mask & 0x0800u) out.put(c.w);
    if (mask & 0x1000u) out.put(d.x);
    if (mask & 0x2000u) out.put(d.y);
    if (mask & 0x4000u) out.put(d.z);
    if (mask & 0x8000u) out.put(d.w);
}

struct PtrSink {
    int32_t *p;
    BXG_HD void put(int32_t v) { *p++ = v; }
};

template <typename LD4>
BXG_HD int32_t *emit_group(const int32_t *I, uint32_t k0, unsigned mask, int32_t *dst, const LD4 &ld4, int mul = 1) {
    PtrSink out{dst};
    emit_group_to(I, k0, mask, out, ld4, mul);
    return out.p;
}

}  // namespace bxs
