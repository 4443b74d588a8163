/* abi_client.c -- a plain C program that uses libbxb200.so through include/bxb200.h only (no Python, no CUDA headers):
 * what a C caller of src/binBits.h / a Cython binding would do.  Prints results for tests/test_gpu_abi_client.py. */
#include <stdio.h>
#include <stdlib.h>

#include "bxb200.h"

#define CHK(x)                                                         \
    do {                                                               \
        int rc__ = (x);                                                \
        if (rc__ != BXG_OK) {                                          \
            fprintf(stderr, "%s -> %d: %s\n", #x, rc__, bxg_last_error()); \
            return 1;                                                  \
        }                                                              \
    } while (0)

int main(void) {
    CHK(bxg_init(0));
    /* bit sets: two 10 kb bitmaps, and + count, strict count_range after invert, runs */
    bxg_bits_t *a, *b;
    CHK(bxg_bits_create(10000, 10, &a));
    CHK(bxg_bits_create(10000, 10, &b));
    int32_t sa[3] = {0, 100, 5000}, ca[3] = {10, 900, 2500}, sb[2] = {50, 6000}, cb[2] = {500, 100};
    CHK(bxg_bits_set_ranges(a, sa, ca, 3, BXG_HOST));
    CHK(bxg_bits_set_ranges(b, sb, cb, 2, BXG_HOST));
    int64_t n_and = 0;
    CHK(bxg_bits_and_count(a, b, &n_and));
    printf("and_count %lld\n", (long long)n_and);
    int64_t nruns = 0;
    CHK(bxg_bits_runs_count(a, &nruns));
    int32_t *rs = malloc(sizeof(int32_t) * (size_t)(nruns + 1)), *re = malloc(sizeof(int32_t) * (size_t)(nruns + 1));
    CHK(bxg_bits_runs_fetch(a, rs, re, nruns));
    for (int64_t i = 0; i < nruns; i++) printf("run %d %d\n", rs[i], re[i]);
    CHK(bxg_bits_not(a));
    int32_t qs[3] = {1500, 1000, 0}, qc[3] = {100, 1000, 10000}, out[3];
    CHK(bxg_bits_count_ranges(a, qs, qc, 3, out, 1, BXG_HOST));
    printf("strict_counts %d %d %d\n", out[0], out[1], out[2]);
    int32_t nx = -1;
    CHK(bxg_bits_next(a, 100, 10000, 1, &nx));
    printf("next_set %d\n", nx);
    CHK(bxg_bits_free(a));
    CHK(bxg_bits_free(b));

    /* interval index: the SURVEY probe tree, ordered hit lists through the pipelined host entry point */
    bxg_itree_t *t;
    CHK(bxg_itree_create(&t));
    int32_t s[5] = {10, 15, 30, -5, 20}, e[5] = {20, 12, 30, 3, 25};
    CHK(bxg_itree_build(t, NULL, s, e, 5, 1, BXG_HOST));
    int32_t fs[4] = {-100, 18, 19, 29}, fe[4] = {100, 11, 21, 31};
    const int64_t *off;
    const int32_t *hits;
    int64_t total = 0;
    CHK(bxg_itree_find_host(t, NULL, fs, fe, 4, &off, &hits, &total));
    for (int q = 0; q < 4; q++) {
        printf("find %d:", q);
        for (int64_t k = off[q]; k < off[q + 1]; k++) printf(" %d", hits[k]);
        printf("\n");
    }
    printf("total %lld\n", (long long)total);
    /* the same through the small-batch entry point (scalar IntervalTree.find) and with int32 offsets */
    const int64_t *soff;
    const int32_t *shits;
    CHK(bxg_itree_find_small(t, NULL, fs, fe, 4, &soff, &shits, &total));
    for (int q = 0; q < 4; q++) {
        printf("small %d:", q);
        for (int64_t k = soff[q]; k < soff[q + 1]; k++) printf(" %d", shits[k]);
        printf("\n");
    }
    const int32_t *off32;
    CHK(bxg_itree_find_host32(t, NULL, fs, fe, 4, &off32, &hits, &total));
    printf("off32 %d %d %d %d %d\n", off32[0], off32[1], off32[2], off32[3], off32[4]);
    /* join: left intervals against the tree's items, mincols = 3 (join.py:35-50) */
    int32_t ls[2] = {0, 18}, le[2] = {16, 40};
    int64_t npairs = 0;
    CHK(bxg_itree_join(t, NULL, ls, le, 2, s, e, 3, BXG_HOST, &npairs));
    int64_t poff[3];
    int32_t pitems[16];
    uint8_t visited[5];
    CHK(bxg_itree_join_fetch(poff, pitems, visited));
    printf("join %lld:", (long long)npairs);
    for (int q = 0; q < 2; q++) {
        for (int64_t k = poff[q]; k < poff[q + 1]; k++) printf(" %d>%d", q, pitems[k]);
    }
    printf(" visited %d%d%d%d%d\n", visited[0], visited[1], visited[2], visited[3], visited[4]);
    CHK(bxg_itree_free(t));

    /* score sources: a 64-cell track, ordered span writes (the last record wins), reads, a bigWig-style summary */
    bxg_scores_t *sc;
    CHK(bxg_scores_alloc(64, 0, -1.0f, &sc));
    int32_t ps[3] = {2, 4, 3}, pe[3] = {6, 5, 4};
    float pv[3] = {1.0f, 2.0f, 3.0f};
    CHK(bxg_scores_set_spans(sc, ps, pe, pv, 3, BXG_HOST));
    float cells[8];
    CHK(bxg_scores_get_range(sc, 0, 8, cells));
    printf("cells");
    for (int i = 0; i < 8; i++) printf(" %g", cells[i]);
    printf("\n");
    CHK(bxg_scores_free(sc));
    int32_t is_[3] = {0, 5, 9}, ie[3] = {5, 8, 20};
    float iv[3] = {1.5f, 2.5f, -1.0f};
    double vc[3] = {0, 0, 0}, mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0}, sm[3] = {0, 0, 0}, sq[3] = {0, 0, 0};
    CHK(bxg_summarize(is_, ie, iv, 3, BXG_HOST, 0, 21, 3, vc, mn, mx, sm, sq));
    printf("summary %g %g %g | %g %g %g | %g %g %g\n", vc[0], vc[1], vc[2], sm[0], sm[1], sm[2], sq[0], sq[1], sq[2]);

    /* genome-wide set_range: two bitmaps, one launch */
    bxg_bits_t *g[2];
    CHK(bxg_bits_create(1000, 10, &g[0]));
    CHK(bxg_bits_create(500, 10, &g[1]));
    int32_t gw[3] = {1, 0, 1}, gs[3] = {10, 0, 400}, gc[3] = {20, 300, 100};
    CHK(bxg_bits_set_ranges_multi(g, 2, gw, gs, gc, 3, BXG_HOST));
    int64_t c0 = 0, c1 = 0;
    CHK(bxg_bits_count_all(g[0], &c0));
    CHK(bxg_bits_count_all(g[1], &c1));
    printf("multi %lld %lld\n", (long long)c0, (long long)c1);
    /* the bed_intersect counting pass in three calls: per-line count_range on the genome, the per-chromosome counters
     * (lines with >= mincols covered bases, covered bases), whole-bitmap counts; then pieces inside ranges and clear */
    int32_t qw[4] = {0, 1, 1, 5}, qst[4] = {0, 0, 390, 0}, qct[4] = {100, 500, 100, 10}, per_line[4];
    CHK(bxg_bits_count_ranges_multi(g, 2, qw, qst, qct, 4, per_line, 0, BXG_HOST));
    printf("per_line %d %d %d %d\n", per_line[0], per_line[1], per_line[2], per_line[3]);
    int64_t stats[4] = {0, 0, 0, 0}, all2[2] = {0, 0};
    CHK(bxg_group_stats_i32(qw, per_line, 4, 2, 50, stats, BXG_HOST));
    printf("stats %lld %lld %lld %lld\n", (long long)stats[0], (long long)stats[1], (long long)stats[2], (long long)stats[3]);
    CHK(bxg_bits_count_all_multi(g, 2, all2, 1, BXG_HOST));
    printf("all_multi %lld %lld\n", (long long)all2[0], (long long)all2[1]);
    int32_t rrs[2] = {0, 350}, rre[2] = {450, 500};
    int64_t roff[3], rtotal = 0;
    CHK(bxg_bits_runs_in_ranges(g[1], rrs, rre, 2, 1, BXG_HOST, roff, &rtotal));
    int32_t ps2[8], pe2[8];
    CHK(bxg_bits_runs_in_ranges_fetch(g[1], ps2, pe2, rtotal));
    printf("pieces %lld |", (long long)rtotal);
    for (int q = 0; q < 2; q++) {
        for (int64_t k = roff[q]; k < roff[q + 1]; k++) printf(" %d:%d-%d", q, ps2[k], pe2[k]);
    }
    printf("\n");
    CHK(bxg_bits_clear(g[1]));
    CHK(bxg_bits_count_all(g[1], &c1));
    printf("cleared %lld\n", (long long)c1);
    CHK(bxg_bits_free(g[0]));
    CHK(bxg_bits_free(g[1]));

    /* scalar IntervalTree.find through the one-call entry point (hit list in mapped host memory) */
    bxg_itree_t *t1;
    CHK(bxg_itree_create(&t1));
    CHK(bxg_itree_build(t1, NULL, s, e, 5, 1, BXG_HOST));
    const int32_t *h1;
    int64_t n1 = bxg_itree_find1(t1, 0, 19, 21, &h1);
    printf("find1 %lld:", (long long)n1);
    for (int64_t k = 0; k < n1; k++) printf(" %d", h1[k]);
    printf("\n");
    CHK(bxg_itree_free(t1));
    return 0;
}
