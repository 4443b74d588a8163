// search_fuzz.cpp -- CPU fuzz of the find kernels' index arithmetic (bx_python_b200/csrc/itree_search.cuh compiled
// by g++): dual_search against std::lower_bound / std::upper_bound over multi-tree segments, and walk_hits (with the
// max-hierarchy skipping) against a naive scan.  Built and run by tests/test_search_logic.py.
#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <vector>

#ifdef FUZZ_WALK_PREFETCH
#define BXS_WALK_PREFETCH
#endif
#include "../bx_python_b200/csrc/itree_search.cuh"

// -DFUZZ_OFFSET=<n> shifts every coordinate by n with saturation: FUZZ_OFFSET=2147400000 exercises the top of the int32
// range (items and queries that touch INT32_MAX, where the padding values of the sampled levels live), a negative
// offset the bottom.
#ifndef FUZZ_OFFSET
#define FUZZ_OFFSET 0
#endif
static int32_t shifted(long long v) {
    v += (long long)FUZZ_OFFSET;
    return (int32_t)std::min<long long>(INT_MAX, std::max<long long>(INT_MIN, v));
}
static uint32_t rnd_state = 12345;
static uint32_t rnd() {
    rnd_state = rnd_state * 1664525u + 1013904223u;
    return rnd_state >> 8;
}

int main(int argc, char **argv) {
    const int trials = argc > 1 ? atoi(argv[1]) : 1500;
    auto ld4 = [](const int4 *p, int4 &a, int4 &b, int4 &c, int4 &d) { a = p[0]; b = p[1]; c = p[2]; d = p[3]; };
    auto ld = [](const int32_t *p) { return *p; };
    auto ld8 = [](const int4 *p, int4 &a, int4 &b) { a = p[0]; b = p[1]; };
    long checks = 0;
    for (int trial = 0; trial < trials; trial++) {
        const int n = 1 + (int)(rnd() % (trial % 4 == 0 ? 200000 : (trial % 7 == 3 ? 40 : 3000)));
        const int maxsplit = (trial % 2) ? 4096 : 8;      // a tiny splitter budget exercises the deep levels
        const int ntrees = 1 + (int)(rnd() % 6);
        std::vector<uint32_t> toff(ntrees + 1);
        for (int t = 1; t < ntrees; t++) toff[t] = rnd() % (uint32_t)(n + 1);
        toff[0] = 0;
        toff[ntrees] = (uint32_t)n;
        std::sort(toff.begin(), toff.end());
        // every 7th trial spreads few items over 16 M positions: wide grid cells (shift > 20, one or two packed offsets)
        const int range = (trial % 7 == 3) ? 16000000 : 1 + (int)(rnd() % 100000);
        const long npad = ((n + 15) & ~15l) + 16;
        std::vector<int32_t> S(npad, INT_MAX), PM(npad, INT_MAX), E(npad, INT_MIN);
        for (int t = 0; t < ntrees; t++) {
            for (uint32_t i = toff[t]; i < toff[t + 1]; i++) S[i] = shifted((long long)(rnd() % (uint32_t)range) - 50);
            std::sort(S.begin() + toff[t], S.begin() + toff[t + 1]);
            int pm = INT_MIN;
            const bool adversarial = trial % 5 == 0;
            for (uint32_t i = toff[t]; i < toff[t + 1]; i++) {
                int len = (int)(rnd() % 40) - 2;
                if (adversarial && i < toff[t] + 2) len = range * 2;     // chromosome-long items in front
                E[i] = (int32_t)std::min<long long>(INT_MAX, std::max<long long>(INT_MIN, (long long)S[i] + len));
                pm = std::max(pm, E[i]);
                PM[i] = pm;
            }
        }
        int shift = 0;
        while ((n + (1ll << shift) - 1) / (1ll << shift) > maxsplit) shift++;
        const int nsplit = (int)((n + (1ll << shift) - 1) >> shift);
        std::vector<int32_t> spS(nsplit + 4, INT_MAX), spPM(nsplit + 4, INT_MAX);
        for (int k = 0; k < nsplit; k++) {
            spS[k] = S[(size_t)k << shift];
            spPM[k] = PM[(size_t)k << shift];
        }
        const int nk = std::max(1, (shift + 3) / 4);
        std::vector<std::vector<int32_t>> ls(nk), lp(nk);
        std::vector<const int32_t *> KS(nk), KP(nk);
        for (int j = 0; j < nk; j++) {
            const int ss = 4 * j;
            const long nout = (n + (1l << ss) - 1) >> ss, pad = ((nout + 15) & ~15l) + 16;
            ls[j].assign(pad, INT_MAX);
            lp[j].assign(pad, INT_MAX);
            for (long i = 0; i < nout; i++) {
                ls[j][i] = S[i << ss];
                lp[j][i] = PM[i << ss];
            }
            KS[j] = ls[j].data();
            KP[j] = lp[j].data();
        }
        // 8-ary levels (strides 8^j) for search_walk_probe8
        const int n8 = std::max(1, (shift + 2) / 3);
        std::vector<std::vector<int32_t>> qs8(n8), qp8(n8);
        std::vector<const int32_t *> QS(n8), QP(n8);
        for (int j = 0; j < n8; j++) {
            const int ss = 3 * j;
            const long nout = (n + (1l << ss) - 1) >> ss, pad = ((nout + 7) & ~7l) + 8;
            qs8[j].assign(pad, INT_MAX);
            qp8[j].assign(pad, INT_MAX);
            for (long i = 0; i < nout; i++) {
                qs8[j][i] = S[i << ss];
                qp8[j][i] = PM[i << ss];
            }
            QS[j] = qs8[j].data();
            QP[j] = qp8[j].data();
        }
        QS[0] = S.data();
        QP[0] = PM.data();
        // max hierarchy
        std::vector<std::vector<int32_t>> M;
        {
            const int32_t *src = E.data();
            long len = n;
            while (true) {
                long nout = (len + 31) / 32;
                std::vector<int32_t> lvl(nout, INT_MIN);
                for (long b = 0; b < nout; b++)
                    for (long k = b * 32; k < std::min(len, b * 32 + 32); k++) lvl[b] = std::max(lvl[b], src[k]);
                M.push_back(lvl);
                src = M.back().data();
                len = nout;
                if (nout <= 1 || M.size() >= 6) break;
            }
        }
        std::vector<const int32_t *> Mp;
        for (auto &m : M) Mp.push_back(m.data());
        // interleaved level-0 layouts, as the kernels use them: [S x16 | PM x16] and [E x16 | I x16] per 16-item group
        std::vector<int32_t> SPi(2 * npad), EIi(2 * npad);
        for (long g = 0; g < npad; g += 16)
            for (int i = 0; i < 16; i++) {
                SPi[2 * g + i] = S[g + i];
                SPi[2 * g + 16 + i] = PM[g + i];
                EIi[2 * g + i] = E[g + i];
                EIi[2 * g + 16 + i] = (int32_t)(g + i) * 7 + 1;      // "item id" of position g+i
            }
        const int mul = (trial % 3 == 0) ? 1 : 2;
        if (mul == 2) {
            KS[0] = SPi.data();
            KP[0] = SPi.data() + 16;
        }
        // direct-address grid, built exactly as csrc/itree.cu does (k_tree_extent / host geometry / k_build_grid)
        const int dens = trial % 3;                       // 1, 2-4 or 4-8 items per cell
        std::vector<bxs::GridDir> gd(ntrees);
        std::vector<bxs::GridRec> G;
        for (int t = 0; t < ntrees; t++) {
            const long nt = (long)toff[t + 1] - (long)toff[t];
            const unsigned long long span = nt > 0 ? (unsigned long long)((long long)S[toff[t + 1] - 1] - (long long)S[toff[t]]) : 0ull;
            const unsigned long long limit = (unsigned long long)std::max<long>(1, nt >> dens);
            int gshift = 0;
            while ((span >> gshift) + 1 > limit) gshift++;
            gd[t].base = nt > 0 ? S[toff[t]] : 0;
            gd[t].shift = gshift;
            gd[t].ncells = nt > 0 ? (uint32_t)((span >> gshift) + 1) : 0u;
            gd[t].coff = (uint32_t)G.size();
            for (uint32_t c = 0; c <= gd[t].ncells; c++) {
                const long long v = (long long)gd[t].base + ((long long)c << gshift);
                const auto less = [](int32_t e, long long key) { return (long long)e < key; };
                const auto leq = [](long long key, int32_t e) { return key < (long long)e; };
                const uint32_t x = (uint32_t)(std::lower_bound(S.begin() + toff[t], S.begin() + toff[t + 1], v, less) - S.begin());
                const uint32_t y = (uint32_t)(std::upper_bound(PM.begin() + toff[t], PM.begin() + toff[t + 1], v, leq) - PM.begin());
                G.push_back(bxs::GridRec{x, y});
            }
        }
        auto ldr = [](const bxs::GridRec *p) { return *p; };
        // 16-byte records, packed as k_build_grid16 does
        std::vector<bxs::GridRec16> G16(G.size());
        for (int t = 0; t < ntrees; t++)
            for (uint32_t c = 0; c <= gd[t].ncells; c++) {
                const size_t r = gd[t].coff + c;
                unsigned long long p = 0;
                if (c < gd[t].ncells) {
                    const uint32_t cnt = G[r + 1].x - G[r].x;
                    p = cnt < 255u ? cnt : 255u;
                    const uint32_t kmax = (uint32_t)bxs::grid16_fields(gd[t].shift), stored = std::min(cnt, kmax);
                    const long long v = (long long)gd[t].base + ((long long)c << gd[t].shift);
                    for (uint32_t i = 0; i < stored; i++)
                        p |= (unsigned long long)((long long)S[G[r].x + i] - v) << (8 + i * gd[t].shift);
                }
                G16[r] = bxs::GridRec16{G[r].x, G[r].y, (uint32_t)p, (uint32_t)(p >> 32)};
            }
        auto ldr16 = [](const bxs::GridRec16 *p) { return *p; };
        for (int q = 0; q < 150; q++) {
            const int t = (int)(rnd() % (uint32_t)ntrees);
            const int32_t qs = shifted((long long)(rnd() % (uint32_t)(range + 40)) - 70);
            const int32_t qe = (int32_t)std::min<long long>(INT_MAX, std::max<long long>(INT_MIN, (long long)qs + (int)(rnd() % 60) - 5));
            uint32_t hi, lo;
            const bool coarse = (q % 2) == 1;      // coarse_lo: lo may be up to 15 items early, never late
            bxs::dual_search(KS.data(), KP.data(), nk, spS.data(), spPM.data(), shift, toff[t], toff[t + 1], qe, qs, ld4,
                             ld, hi, lo, bxs::NoPrefetch(), mul, coarse);
            const uint32_t ehi = (uint32_t)(std::lower_bound(S.begin() + toff[t], S.begin() + toff[t + 1], qe) - S.begin());
            const uint32_t elo = (uint32_t)(std::upper_bound(PM.begin() + toff[t], PM.begin() + toff[t + 1], qs) - PM.begin());
            const bool lo_ok = coarse ? (lo <= elo && lo + 16 > elo && lo >= toff[t]) : (lo == elo);
            if (hi != ehi || !lo_ok) {
                printf("SEARCH MISMATCH trial %d n=%d shift=%d nk=%d seg=[%u,%u) qs=%d qe=%d hi=%u/%u lo=%u/%u\n", trial, n,
                       shift, nk, toff[t], toff[t + 1], qs, qe, hi, ehi, lo, elo);
                return 1;
            }
            if (lo > hi) lo = hi;
            std::vector<uint32_t> got, want;
            std::vector<int32_t> ids(hi - lo + 32), want_ids;
            int32_t *dst = ids.data();
            const int32_t *Eb = mul == 2 ? EIi.data() : E.data();
            bxs::walk_hits(Eb, Mp.data(), (int)Mp.size(), lo, hi, qs, ld4, ld, [&](uint32_t k0, unsigned mask) {
                if (mul == 2) dst = bxs::emit_group(EIi.data() + 16, k0, mask, dst, ld4, 2);
                while (mask) {
                    int b = bxs::ffs32(mask) - 1;
                    mask &= mask - 1;
                    got.push_back(k0 + b);
                }
            }, bxs::NoPrefetch(), mul);
            for (uint32_t k = lo; k < hi; k++)
                if (E[k] > qs) {
                    want.push_back(k);
                    want_ids.push_back((int32_t)k * 7 + 1);
                }
            if (mul == 2 && std::vector<int32_t>(ids.data(), dst) != want_ids) {
                printf("EMIT MISMATCH trial %d lo=%u hi=%u\n", trial, lo, hi);
                return 1;
            }
            if (mul == 1) {           // the overlapped search+walk the count kernels use (plain level-0 arrays)
                std::vector<uint32_t> got2;
                uint32_t hi2, lo2;
                bxs::search_walk(KS.data(), KP.data(), nk, spS.data(), spPM.data(), shift, toff[t], toff[t + 1], qe, qs,
                                 E.data(), Mp.data(), (int)Mp.size(), ld4, ld, hi2, lo2, [&](uint32_t k0, unsigned mask) {
                                     while (mask) {
                                         int b = bxs::ffs32(mask) - 1;
                                         mask &= mask - 1;
                                         got2.push_back(k0 + b);
                                     }
                                 });
                // the single-search form with the backward probe for the walk start (what the count kernels run)
                std::vector<uint32_t> got3;
                uint32_t hi3, lo3;
                bxs::search_walk_probe(KS.data(), KP.data(), nk, spS.data(), spPM.data(), shift, toff[t], toff[t + 1], qe,
                                       qs, E.data(), Mp.data(), (int)Mp.size(), ld4, ld8, ld, hi3, lo3,
                                       [&](uint32_t k0, unsigned mask) {
                                           while (mask) {
                                               int b = bxs::ffs32(mask) - 1;
                                               mask &= mask - 1;
                                               got3.push_back(k0 + b);
                                           }
                                       });
                // the naive expectation over the whole candidate range (the probe may start earlier than dual_search's lo)
                std::vector<uint32_t> want3;
                for (uint32_t k = toff[t]; k < ehi; k++)
                    if (E[k] > qs) want3.push_back(k);
                if (hi3 != ehi || lo3 > std::min(elo, ehi) || lo3 < toff[t] || got3 != want3) {
                    printf("SEARCH_WALK_PROBE MISMATCH trial %d n=%d nk=%d seg=[%u,%u) qs=%d qe=%d hi=%u/%u lo=%u/%u got=%zu want=%zu\n",
                           trial, n, nk, toff[t], toff[t + 1], qs, qe, hi3, ehi, lo3, elo, got3.size(), want3.size());
                    return 1;
                }
                std::vector<uint32_t> got4;
                uint32_t hi4, lo4;
                bxs::search_walk_probe8(QS.data(), QP.data(), n8, spS.data(), spPM.data(), shift, toff[t], toff[t + 1], qe,
                                        qs, E.data(), Mp.data(), (int)Mp.size(), ld8, ld8, ld, hi4, lo4,
                                        [&](uint32_t k0, unsigned mask) {
                                            while (mask) {
                                                int b = bxs::ffs32(mask) - 1;
                                                mask &= mask - 1;
                                                got4.push_back(k0 + b);
                                            }
                                        });
                if (hi4 != ehi || lo4 > std::min(elo, ehi) || lo4 < toff[t] || got4 != want3) {
                    printf("SEARCH_WALK_PROBE8 MISMATCH trial %d n=%d n8=%d shift=%d seg=[%u,%u) qs=%d qe=%d hi=%u/%u lo=%u/%u got=%zu want=%zu\n",
                           trial, n, n8, shift, toff[t], toff[t + 1], qs, qe, hi4, ehi, lo4, elo, got4.size(), want3.size());
                    return 1;
                }
                std::vector<uint32_t> got5;
                uint32_t hi5, lo5;
                bxs::search_walk_grid(G.data(), gd[t], S.data(), toff[t], toff[t + 1], qe, qs, E.data(), Mp.data(),
                                      (int)Mp.size(), ldr, ld8, ld, hi5, lo5, [&](uint32_t k0, unsigned mask) {
                                          while (mask) {
                                              int b = bxs::ffs32(mask) - 1;
                                              mask &= mask - 1;
                                              got5.push_back(k0 + b);
                                          }
                                      });
                if (hi5 != ehi || lo5 > std::min(elo, ehi) || lo5 < toff[t] || got5 != want3) {
                    printf("SEARCH_WALK_GRID MISMATCH trial %d n=%d dens=%d seg=[%u,%u) qs=%d qe=%d hi=%u/%u lo=%u/%u got=%zu want=%zu\n",
                           trial, n, dens, toff[t], toff[t + 1], qs, qe, hi5, ehi, lo5, elo, got5.size(), want3.size());
                    return 1;
                }
                std::vector<uint32_t> got6;
                uint32_t hi6, lo6;
                bxs::search_walk_grid16(G16.data(), gd[t], S.data(), toff[t], toff[t + 1], qe, qs, E.data(), Mp.data(),
                                        (int)Mp.size(), ldr16, ld8, ld, hi6, lo6, [&](uint32_t k0, unsigned mask) {
                                            while (mask) {
                                                int b = bxs::ffs32(mask) - 1;
                                                mask &= mask - 1;
                                                got6.push_back(k0 + b);
                                            }
                                        });
                if (hi6 != ehi || lo6 > std::min(elo, ehi) || lo6 < toff[t] || got6 != want3) {
                    printf("SEARCH_WALK_GRID16 MISMATCH trial %d n=%d dens=%d shift=%d seg=[%u,%u) qs=%d qe=%d hi=%u/%u lo=%u/%u got=%zu want=%zu\n",
                           trial, n, dens, gd[t].shift, toff[t], toff[t + 1], qs, qe, hi6, ehi, lo6, elo, got6.size(), want3.size());
                    return 1;
                }
                if (hi2 != ehi || lo2 > std::min(elo, ehi) || got2 != want) {
                    printf("SEARCH_WALK MISMATCH trial %d n=%d seg=[%u,%u) qs=%d qe=%d hi=%u/%u lo=%u/%u got=%zu want=%zu\n", trial,
                           n, toff[t], toff[t + 1], qs, qe, hi2, ehi, lo2, elo, got2.size(), want.size());
                    return 1;
                }
            }
            if (got != want) {
                printf("WALK MISMATCH trial %d n=%d lo=%u hi=%u qs=%d got=%zu want=%zu\n", trial, n, lo, hi, qs, got.size(),
                       want.size());
                return 1;
            }
            checks++;
        }
    }
    printf("ok %ld\n", checks);
    return 0;
}
