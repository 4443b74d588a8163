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
    CHK(bxg_itree_free(t));
    return 0;
}
