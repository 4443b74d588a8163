/*
 * bx_oracle.c -- CPU restatement ("port") of the bx-python hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker for the CUDA library, not a product path: only tests/, bench.py's cpu_baseline /
 * --impl reference legs and __graft_entry__.smoke() may load it.  bx_python_b200/ never does.
 *
 * Parity status: PINNED.  Every function below is validated (tests/test_oracle_vs_ref.py, run in the build
 * container where /root/reference exists) against the unmodified reference compiled by oracle/Makefile into
 * oracle/_ref/, and against the committed golden vectors in tests/golden/ (generated from that compiled
 * reference by tests/golden/make_golden.py) everywhere else.
 *
 * What is restated (citations are relative to /root/reference):
 *   orc_itree_order / orc_itree_find_*  <- lib/bx/intervals/intersection.pyx:103-138 (insert tie rule),
 *                                          :180-189 (_intersect predicate + in-order emission)
 *   orc_itree_before / orc_itree_after  <- intersection.pyx:192-260 (_seek_left/_seek_right + left/right)
 *   orc_bb_*                            <- src/binBits.c:8-317 over src/kent/bits.c:51-263 (MSB-first bytes,
 *                                          per-bin ZERO/ONE sentinels, count bug at binBits.c:155,161)
 *   orc_bits_* (flat BitSet)            <- src/kent/bits.c:51-263 via lib/bx/bitset.pyx:107-173
 *   orc_aggregate                       <- scripts/aggregate_scores_in_intervals.py:107-134 with
 *                                          lib/bx/binned_array.py:89-94 float32 semantics
 *   orc_scores_set_spans                <- scripts/aggregate_scores_in_intervals.py:60-70 over lib/bx/wiggle.py:71-85
 *                                          (per-base BinnedArray assignment in file order)
 *   orc_summarize                       <- lib/bx/bbi/bbi_file.pyx:80-111 (SummarizedData.accumulate_interval_value)
 *   orc_join                            <- lib/bx/intervals/operations/join.py:30-61 over quicksect.py:115-121
 *
 * Written as a closed-form restatement, not a transliteration: the treap is replaced by the total order its
 * in-order traversal realises; the sentinel pointers are replaced by a per-bin state byte.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------------------
 * Interval index
 * ------------------------------------------------------------------------------------------------------------
 * intersection.pyx:110-116: a new (start,end) descends right iff  (start==node.start ? end : start) > node.start.
 * Rotations (:140-152) preserve in-order, so the in-order sequence is the unique total order
 *      key(i) = (start[i], c[i], c[i] ? +i : -i),   c[i] = (end[i] > start[i])
 * i.e. among equal starts: non-proper items (end<=start) first, newest first; then proper items, oldest first.
 */
typedef struct { int32_t s; int32_t c; int64_t t; int32_t i; } okey_t;

static int okey_cmp(const void *a, const void *b)
{
    const okey_t *x = (const okey_t *)a, *y = (const okey_t *)b;
    if (x->s != y->s) return x->s < y->s ? -1 : 1;
    if (x->c != y->c) return x->c < y->c ? -1 : 1;
    if (x->t != y->t) return x->t < y->t ? -1 : 1;
    return 0;
}

/* perm[k] = insertion index of the k-th item of the in-order traversal. */
int orc_itree_order(const int32_t *start, const int32_t *end, int64_t n, int32_t *perm)
{
    okey_t *k = (okey_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(okey_t));
    if (!k) return -1;
    for (int64_t i = 0; i < n; i++) {
        k[i].s = start[i];
        k[i].c = end[i] > start[i];
        k[i].t = k[i].c ? i : -i;
        k[i].i = (int32_t)i;
    }
    qsort(k, (size_t)n, sizeof(okey_t), okey_cmp);
    for (int64_t i = 0; i < n; i++) perm[i] = k[i].i;
    free(k);
    return 0;
}

/* Sorted-order arrays + prefix max of end, built once per index. */
typedef struct {
    int64_t n;
    int32_t *S, *E, *I, *PM;
} orc_itree_t;

orc_itree_t *orc_itree_build(const int32_t *start, const int32_t *end, int64_t n)
{
    orc_itree_t *t = (orc_itree_t *)calloc(1, sizeof(*t));
    size_t m = (size_t)(n > 0 ? n : 1);
    t->n = n;
    t->S = (int32_t *)malloc(m * 4); t->E = (int32_t *)malloc(m * 4);
    t->I = (int32_t *)malloc(m * 4); t->PM = (int32_t *)malloc(m * 4);
    orc_itree_order(start, end, n, t->I);
    int32_t pm = INT32_MIN;
    for (int64_t k = 0; k < n; k++) {
        t->S[k] = start[t->I[k]];
        t->E[k] = end[t->I[k]];
        if (t->E[k] > pm) pm = t->E[k];
        t->PM[k] = pm;
    }
    return t;
}

void orc_itree_free(orc_itree_t *t)
{
    if (!t) return;
    free(t->S); free(t->E); free(t->I); free(t->PM); free(t);
}

/* first k with A[k] >= v  (A non-decreasing) */
static int64_t lower_bound_i32(const int32_t *A, int64_t n, int32_t v)
{
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (A[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}
/* first k with A[k] > v */
static int64_t upper_bound_i32(const int32_t *A, int64_t n, int32_t v)
{
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (A[mid] <= v) lo = mid + 1; else hi = mid; }
    return lo;
}

/*
 * intersection.pyx:180-189: emit, in in-order sequence, every item with  end > qs  &&  start < qe.
 * (The maxend / start<qe tests on subtrees only prune; they never change the result.)
 * offsets[q]..offsets[q+1] delimit query q's hits; pass hits==NULL to only count.
 */
int orc_itree_find(const orc_itree_t *t, const int32_t *qs, const int32_t *qe, int64_t nq,
                   int64_t *offsets /* nq+1 */, int32_t *hits /* may be NULL */)
{
    int64_t total = 0;
    for (int64_t q = 0; q < nq; q++) {
        offsets[q] = total;
        int64_t hi = lower_bound_i32(t->S, t->n, qe[q]);       /* items with start < qe  */
        int64_t lo = upper_bound_i32(t->PM, t->n, qs[q]);      /* first item whose prefix-max end > qs */
        for (int64_t k = lo; k < hi; k++)
            if (t->E[k] > qs[q]) { if (hits) hits[total] = t->I[k]; total++; }
    }
    offsets[nq] = total;
    return 0;
}

/*
 * intersection.pyx:192-245 left(position,n,max_dist): p = position-1; collect, in REVERSED in-order, every
 * item with 0 <= p - end < max_dist (the n argument never stops the walk, :192-210); if exactly n were found
 * return them as collected, else stable-sort by end descending and keep the first n.
 * out receives insertion indices; returns the number written (<= cap) or -1.
 */
typedef struct { int32_t key; int64_t pos; int32_t idx; } nb_t;
static int nb_cmp(const void *a, const void *b)
{
    const nb_t *x = (const nb_t *)a, *y = (const nb_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->pos < y->pos ? -1 : (x->pos > y->pos);
}

int64_t orc_itree_before(const orc_itree_t *t, int32_t position, int32_t n, int32_t max_dist,
                         int32_t *out, int64_t cap)
{
    int64_t p = (int64_t)position - 1, m = 0;
    nb_t *c = (nb_t *)malloc((size_t)(t->n > 0 ? t->n : 1) * sizeof(nb_t));
    for (int64_t k = t->n - 1; k >= 0; k--) {
        int64_t d = p - (int64_t)t->E[k];
        if (d > -1 && d < (int64_t)max_dist) { c[m].key = -t->E[k]; c[m].pos = m; c[m].idx = t->I[k]; m++; }
    }
    if (m != (int64_t)n) { qsort(c, (size_t)m, sizeof(nb_t), nb_cmp); if (m > n) m = n < 0 ? 0 : n; }
    if (m > cap) { free(c); return -1; }
    for (int64_t k = 0; k < m; k++) out[k] = c[k].idx;
    free(c);
    return m;
}

/* intersection.pyx:212-260 right(position,n,max_dist): p = position+1; forward in-order; 0 <= start-p < max_dist */
int64_t orc_itree_after(const orc_itree_t *t, int32_t position, int32_t n, int32_t max_dist,
                        int32_t *out, int64_t cap)
{
    int64_t p = (int64_t)position + 1, m = 0;
    nb_t *c = (nb_t *)malloc((size_t)(t->n > 0 ? t->n : 1) * sizeof(nb_t));
    for (int64_t k = 0; k < t->n; k++) {
        int64_t d = (int64_t)t->S[k] - p;
        if (d > -1 && d < (int64_t)max_dist) { c[m].key = t->S[k]; c[m].pos = m; c[m].idx = t->I[k]; m++; }
    }
    if (m != (int64_t)n) { qsort(c, (size_t)m, sizeof(nb_t), nb_cmp); if (m > n) m = n < 0 ? 0 : n; }
    if (m > cap) { free(c); return -1; }
    for (int64_t k = 0; k < m; k++) out[k] = c[k].idx;
    free(c);
    return m;
}

/* ------------------------------------------------------------------------------------------------------------
 * BinnedBitSet  (binBits.c over kent/bits.c)
 * ------------------------------------------------------------------------------------------------------------
 * Each bin is in one of three states (binBits.c:5-6): Z = NULL sentinel, O = &"ONE" sentinel, A = allocated
 * MSB-first byte array of (bin_size+7)>>3 bytes (bits.c:12-14,51-56).
 */
enum { BZ = 0, BO = 1, BA = 2 };

typedef struct {
    int32_t size, bin_size, nbins;
    uint8_t *state;     /* nbins */
    uint8_t **bin;      /* nbins byte arrays, valid iff state==BA */
} orc_bb_t;

static int bin_bytes(const orc_bb_t *b) { return (b->bin_size + 7) >> 3; }

/* binBits.c:8-17 -- NOTE the float32 division and the double ceil. */
void orc_bb_geometry(int32_t size, int32_t granularity, int32_t *bin_size, int32_t *nbins)
{
    *bin_size = (int32_t)ceil(size / (float)granularity);
    *nbins = (int32_t)ceil(size / (float)(*bin_size));
}

orc_bb_t *orc_bb_new(int32_t size, int32_t granularity)
{
    orc_bb_t *b = (orc_bb_t *)calloc(1, sizeof(*b));
    b->size = size;
    orc_bb_geometry(size, granularity, &b->bin_size, &b->nbins);
    b->state = (uint8_t *)calloc((size_t)b->nbins, 1);
    b->bin = (uint8_t **)calloc((size_t)b->nbins, sizeof(uint8_t *));
    return b;
}

void orc_bb_free(orc_bb_t *b)
{
    if (!b) return;
    for (int i = 0; i < b->nbins; i++) free(b->bin[i]);
    free(b->bin); free(b->state); free(b);
}

int32_t orc_bb_size(const orc_bb_t *b) { return b->size; }
int32_t orc_bb_bin_size(const orc_bb_t *b) { return b->bin_size; }
int32_t orc_bb_nbins(const orc_bb_t *b) { return b->nbins; }
int32_t orc_bb_state(const orc_bb_t *b, int32_t bin) { return b->state[bin]; }

/* MSB-first helpers (bits.c:12-14): bit k of a byte array is byte k>>3, mask 0x80>>(k&7). */
static inline int  getbit(const uint8_t *a, int k) { return (a[k >> 3] >> (7 - (k & 7))) & 1; }
static inline void setbit(uint8_t *a, int k) { a[k >> 3] |= (uint8_t)(0x80u >> (k & 7)); }
static inline void clrbit(uint8_t *a, int k) { a[k >> 3] &= (uint8_t)~(0x80u >> (k & 7)); }

/* bits.c:86-109 -- first/last byte masked, 0xff in between. */
static void bytes_set_range(uint8_t *a, int off, int cnt)
{
    if (cnt <= 0) return;
    int last = off + cnt - 1, b0 = off >> 3, b1 = last >> 3;
    uint8_t lm = (uint8_t)(0xffu >> (off & 7)), rm = (uint8_t)(0xffu << (7 - (last & 7)));
    if (b0 == b1) { a[b0] |= (uint8_t)(lm & rm); return; }
    a[b0] |= lm;
    if (b1 > b0 + 1) memset(a + b0 + 1, 0xff, (size_t)(b1 - b0 - 1));
    a[b1] |= rm;
}

/* bits.c:118-141 -- the reference uses a 256-entry LUT; any exact popcount is equivalent. */
static int bytes_count_range(const uint8_t *a, int off, int cnt)
{
    if (cnt <= 0) return 0;
    int last = off + cnt - 1, b0 = off >> 3, b1 = last >> 3, c = 0;
    uint8_t lm = (uint8_t)(0xffu >> (off & 7)), rm = (uint8_t)(0xffu << (7 - (last & 7)));
    if (b0 == b1) return __builtin_popcount(a[b0] & lm & rm);
    c = __builtin_popcount(a[b0] & lm);
    for (int i = b0 + 1; i < b1; i++) c += __builtin_popcount(a[i]);
    return c + __builtin_popcount(a[b1] & rm);
}

/* bits.c:143-176 -- first position in [off, nbits) holding val, else nbits. */
static int bytes_find(const uint8_t *a, int off, int val, int nbits)
{
    uint8_t skip = val ? 0x00 : 0xff;
    int k = off;
    while (k < nbits) {
        if ((k & 7) == 0 && k + 8 <= nbits && a[k >> 3] == skip) { k += 8; continue; }
        if (getbit(a, k) == val) return k;
        k++;
    }
    return nbits;
}

static void materialise(orc_bb_t *b, int i, int ones)
{
    b->bin[i] = (uint8_t *)calloc((size_t)bin_bytes(b), 1);
    if (ones) bytes_set_range(b->bin[i], 0, b->bin_size);   /* binBits.c:91-92 */
    b->state[i] = BA;
}

/* binBits.c:49-65 */
int orc_bb_get(const orc_bb_t *b, int32_t pos)
{
    int i = pos / b->bin_size;
    if (b->state[i] == BZ) return 0;
    if (b->state[i] == BO) return 1;
    return getbit(b->bin[i], pos % b->bin_size);
}
/* binBits.c:67-80 */
void orc_bb_set(orc_bb_t *b, int32_t pos)
{
    int i = pos / b->bin_size;
    if (b->state[i] == BO) return;
    if (b->state[i] == BZ) materialise(b, i, 0);
    setbit(b->bin[i], pos % b->bin_size);
}
/* binBits.c:82-96 */
void orc_bb_clear(orc_bb_t *b, int32_t pos)
{
    int i = pos / b->bin_size;
    if (b->state[i] == BZ) return;
    if (b->state[i] == BO) materialise(b, i, 1);
    clrbit(b->bin[i], pos % b->bin_size);
}
/* binBits.c:98-128: every touched Z bin is allocated, O bins are left alone. */
void orc_bb_set_range(orc_bb_t *b, int32_t start, int32_t count)
{
    while (count > 0) {
        int i = start / b->bin_size, off = start % b->bin_size;
        int room = b->bin_size - off, k = room < count ? room : count;
        if (b->state[i] == BZ) materialise(b, i, 0);
        if (b->state[i] != BO) bytes_set_range(b->bin[i], off, k);
        start += k; count -= k;
    }
}
/* binBits.c:130-178, INCLUDING the reference's ALL_ONE arithmetic (:155 `delta - offset`, :161 `size - offset`). */
int32_t orc_bb_count_range(const orc_bb_t *b, int32_t start, int32_t count)
{
    int32_t total = 0;
    while (count > 0) {
        int i = start / b->bin_size, off = start % b->bin_size;
        int room = b->bin_size - off, k = room < count ? room : count;
        if (b->state[i] == BO) total += k - off;
        else if (b->state[i] == BA) total += bytes_count_range(b->bin[i], off, k);
        start += k; count -= k;
    }
    return total;
}
/* binBits.c:180-228: val=1 -> next_set, val=0 -> next_clear; returns size when the bins run out. */
int32_t orc_bb_next(const orc_bb_t *b, int32_t start, int val)
{
    int i = start / b->bin_size, off = start % b->bin_size;
    for (; i < b->nbins; i++, off = 0) {
        int st = b->state[i];
        if (st == (val ? BO : BZ)) return i * b->bin_size + off;
        if (st == BA) {
            int k = bytes_find(b->bin[i], off, val, b->bin_size);
            if (k < b->bin_size) return i * b->bin_size + k;
        }
    }
    return b->size;
}
/* binBits.c:230-262 */
void orc_bb_and(orc_bb_t *a, const orc_bb_t *o)
{
    int nb = bin_bytes(a);
    for (int i = 0; i < a->nbins; i++) {
        if (a->state[i] == BZ) continue;
        if (o->state[i] == BZ) { free(a->bin[i]); a->bin[i] = NULL; a->state[i] = BZ; continue; }
        if (o->state[i] == BO) continue;
        if (a->state[i] == BO) { materialise(a, i, 0); memcpy(a->bin[i], o->bin[i], (size_t)nb); continue; }
        for (int k = 0; k < nb; k++) a->bin[i][k] &= o->bin[i][k];
    }
}
/* binBits.c:264-296 */
void orc_bb_or(orc_bb_t *a, const orc_bb_t *o)
{
    int nb = bin_bytes(a);
    for (int i = 0; i < a->nbins; i++) {
        if (a->state[i] == BO) continue;
        if (o->state[i] == BO) { free(a->bin[i]); a->bin[i] = NULL; a->state[i] = BO; continue; }
        if (o->state[i] == BZ) continue;
        if (a->state[i] == BZ) { materialise(a, i, 0); memcpy(a->bin[i], o->bin[i], (size_t)nb); continue; }
        for (int k = 0; k < nb; k++) a->bin[i][k] |= o->bin[i][k];
    }
}
/* binBits.c:298-317 -- allocated bins are complemented over ALL their bytes (pad bits included). */
void orc_bb_not(orc_bb_t *a)
{
    int nb = bin_bytes(a);
    for (int i = 0; i < a->nbins; i++) {
        if (a->state[i] == BO) a->state[i] = BZ;
        else if (a->state[i] == BZ) a->state[i] = BO;
        else for (int k = 0; k < nb; k++) a->bin[i][k] = (uint8_t)~a->bin[i][k];
    }
}

/* Batched conveniences used by the parity tests (loops over the scalar calls above). */
void orc_bb_set_ranges(orc_bb_t *b, const int32_t *start, const int32_t *count, int64_t n)
{ for (int64_t i = 0; i < n; i++) orc_bb_set_range(b, start[i], count[i]); }
void orc_bb_count_ranges(const orc_bb_t *b, const int32_t *start, const int32_t *count, int64_t n, int32_t *out)
{ for (int64_t i = 0; i < n; i++) out[i] = orc_bb_count_range(b, start[i], count[i]); }
void orc_bb_read(const orc_bb_t *b, const int32_t *pos, int64_t n, uint8_t *out)
{ for (int64_t i = 0; i < n; i++) out[i] = (uint8_t)orc_bb_get(b, pos[i]); }

/* Export bit p as LSB-first uint64 words over [0,size): the layout the CUDA library uses in HBM. */
void orc_bb_export_words(const orc_bb_t *b, uint64_t *words /* ceil(size/64) */)
{
    int64_t nw = ((int64_t)b->size + 63) >> 6;
    memset(words, 0, (size_t)nw * 8);
    for (int64_t p = 0; p < b->size; p++)
        if (orc_bb_get(b, (int32_t)p)) words[p >> 6] |= 1ull << (p & 63);
}

/* Run extraction idiom (scripts/bed_intersect_basewise.py:30-38, lib/bx/bitset_utils.py:34-43):
 *   end = 0; loop: start = next_set(end); if start == size: break; end = next_clear(start); emit (start,end)
 * The caller's next next_set(end) raises IndexError when end == size, which terminates the idiom as well.
 * Returns the number of runs; writes at most cap of them. */
int64_t orc_bb_runs(const orc_bb_t *b, int32_t *rs, int32_t *re, int64_t cap)
{
    int64_t n = 0;
    int32_t end = 0;
    while (end < b->size) {
        int32_t s = orc_bb_next(b, end, 1);
        if (s >= b->size) break;
        end = orc_bb_next(b, s, 0);
        if (n < cap) { rs[n] = s; re[n] = end; }
        n++;
    }
    return n;
}

/* ------------------------------------------------------------------------------------------------------------
 * Flat BitSet (bits.c through bitset.pyx:107-173) -- one MSB-first byte array of (n+7)>>3 bytes.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct { int32_t n; uint8_t *a; } orc_bits_t;

orc_bits_t *orc_bits_new(int32_t n)
{
    orc_bits_t *b = (orc_bits_t *)calloc(1, sizeof(*b));
    b->n = n; b->a = (uint8_t *)calloc((size_t)((n + 7) >> 3) + 1, 1);
    return b;
}
void orc_bits_free(orc_bits_t *b) { if (b) { free(b->a); free(b); } }
int  orc_bits_get(const orc_bits_t *b, int32_t p) { return getbit(b->a, p); }
void orc_bits_set(orc_bits_t *b, int32_t p) { setbit(b->a, p); }
void orc_bits_clear(orc_bits_t *b, int32_t p) { clrbit(b->a, p); }
void orc_bits_set_range(orc_bits_t *b, int32_t s, int32_t c) { bytes_set_range(b->a, s, c); }
int32_t orc_bits_count_range(const orc_bits_t *b, int32_t s, int32_t c) { return bytes_count_range(b->a, s, c); }
/* bitset.pyx:141-150: search [start,end) */
int32_t orc_bits_next(const orc_bits_t *b, int32_t s, int32_t end, int val) { return bytes_find(b->a, s, val, end); }
/* op: 0 and, 1 or, 2 xor (bits.c:222-252) ; not (bits.c:254-263) */
void orc_bits_binop(orc_bits_t *a, const orc_bits_t *o, int op)
{
    int nb = (a->n + 7) >> 3;
    for (int k = 0; k < nb; k++)
        a->a[k] = (uint8_t)(op == 0 ? a->a[k] & o->a[k] : op == 1 ? a->a[k] | o->a[k] : a->a[k] ^ o->a[k]);
}
void orc_bits_not(orc_bits_t *a) { int nb = (a->n + 7) >> 3; for (int k = 0; k < nb; k++) a->a[k] = (uint8_t)~a->a[k]; }

/* ------------------------------------------------------------------------------------------------------------
 * aggregate_scores_in_intervals inner loop (scripts/aggregate_scores_in_intervals.py:107-134)
 * ------------------------------------------------------------------------------------------------------------
 * For each window [ws,we): walk positions in order; v = float32 score (NaN = unset, binned_array.py:73,84-94);
 * skip v == 0.0 (Python truthiness, :115), skip masked (:117-119), skip NaN (:122); total += v in float32
 * (NumPy >= 2: python-int 0 + np.float32 stays float32), count, min, max; avg = total / count in float32.
 * min/max start from the script's sentinels +/-100000000 (:112-113) compared as the script does
 * (max(python_int, np.float32) / min(...) -> keeps whichever compares larger/smaller).
 * Outputs for count==0 windows: sum=0, avg/min/max = NaN (the script prints "nan nan nan").
 * mask: LSB-first uint64 words or NULL.  Positions outside [0,n) read as NaN (unset).
 */
void orc_aggregate(const float *scores, int64_t n, const uint64_t *mask,
                   const int32_t *ws, const int32_t *we, int64_t nw,
                   float *sum, float *avg, int32_t *cnt, float *mn, float *mx)
{
    for (int64_t w = 0; w < nw; w++) {
        volatile float total = 0.0f;
        int32_t c = 0;
        double hi = -100000000.0, lo = 100000000.0;   /* python ints, compared exactly against float32 values */
        for (int64_t p = ws[w]; p < we[w]; p++) {
            if (p < 0 || p >= n) continue;
            float v = scores[p];
            if (v == 0.0f) continue;                              /* :115  `if not s: continue` (NaN is truthy) */
            if (mask && ((mask[p >> 6] >> (p & 63)) & 1)) continue; /* :117-119 */
            if (v != v) continue;                                 /* :122 isNaN */
            total = total + v;
            c++;
            if ((double)v > hi) hi = (double)v;                   /* max(max_score, score) */
            if ((double)v < lo) lo = (double)v;                   /* min(min_score, score) */
        }
        cnt[w] = c;
        sum[w] = total;
        if (c > 0) { avg[w] = total / (float)c; mn[w] = (float)lo; mx[w] = (float)hi; }
        else { avg[w] = NAN; mn[w] = NAN; mx[w] = NAN; }
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * Score sources
 * ------------------------------------------------------------------------------------------------------------
 * load_scores_wiggle: `scores_by_chrom[chrom][pos] = val` for every base of every record, in file order; a later
 * record overwrites an earlier one.  track[] is the dense array (position p at index p - origin), pre-filled by the
 * caller with the BinnedArray default.  end == NULL means single positions.  Returns -1 if a base falls outside.
 */
int orc_scores_set_spans(float *track, int64_t n, int64_t origin, const int32_t *start, const int32_t *end,
                         const float *val, int64_t nspans)
{
    for (int64_t i = 0; i < nspans; i++) {
        int64_t a = start[i], b = end ? end[i] : a + 1;
        for (int64_t p = a; p < b; p++) {
            if (p - origin < 0 || p - origin >= n) return -1;
            track[p - origin] = val[i];
        }
    }
    return 0;
}

/*
 * SummarizedData.accumulate_interval_value for every interval in order.  The caller initialises the five arrays
 * (zeros in SummarizedData.__init__, +inf/-inf min/max in SummarizingBlockHandler).  Types follow the Cython
 * declarations: s, e are bits32 arguments clipped against the region; base_*, overlap, interval_size are C ints;
 * overlap_factor / interval_weight doubles; val a C float (val * val is a float product, then widened).
 * volatile keeps gcc from contracting a*b+c into an fma.
 */
void orc_summarize(const int32_t *start, const int32_t *end, const float *val, int64_t n,
                   uint32_t rstart, uint32_t rend, int32_t size,
                   double *valid_count, double *min_val, double *max_val, double *sum_data, double *sum_squares)
{
    for (int64_t i = 0; i < n; i++) {
        uint32_t s = (uint32_t)start[i], e = (uint32_t)end[i];
        if (s < rstart) s = rstart;
        if (e > rend) e = rend;
        if (s >= e) continue;
        int base_step = (int)((rend - rstart) / (uint32_t)size);
        float v = val[i];
        for (int j = 0; j < size; j++) {
            int base_start = (int)(rstart + (uint32_t)(base_step * j));
            int base_end = base_start + base_step;
            int lo = base_start > (int)s ? base_start : (int)s, hi = base_end < (int)e ? base_end : (int)e;
            int overlap = hi - lo;
            if (overlap > 0) {
                int interval_size = (int)(e - s);
                volatile double overlap_factor = (double)overlap / interval_size;
                volatile double interval_weight = interval_size * overlap_factor;
                volatile double t1 = v * interval_weight;
                volatile float vv = v * v;
                volatile double t2 = vv * interval_weight;
                valid_count[j] += interval_weight;
                sum_data[j] += t1;
                sum_squares[j] += t2;
                if (max_val[j] < v) max_val[j] = v;
                if (min_val[j] > v) min_val[j] = v;
            }
        }
    }
}

/*
 * join.py:30-61: for left interval q, the right items (same tree) with qs < item.end && qe > item.start
 * (quicksect.py:115-121) whose overlap (:35-50, inclusive membership tests) is >= mincols.  Emits, per left interval,
 * the kept item ids in ascending id order (the reference's own order is a random treap pre-order) and visited[].
 * pair_items may be NULL to only count.  Returns the number of kept pairs.
 */
int64_t orc_join(const int32_t *itree, const int32_t *istart, const int32_t *iend, int64_t n,
                 const int32_t *qtree, const int32_t *qs, const int32_t *qe, int64_t nq, int32_t mincols,
                 int64_t *pair_offsets, int32_t *pair_items, uint8_t *visited)
{
    int64_t w = 0;
    for (int64_t q = 0; q < nq; q++) {
        pair_offsets[q] = w;
        int64_t ls = qs[q], le = qe[q];
        for (int64_t i = 0; i < n; i++) {
            if (itree[i] != qtree[q]) continue;
            int64_t s = istart[i], e = iend[i];
            if (!(ls < e && le > s)) continue;
            int in_s = s >= ls && s <= le, in_e = e >= ls && e <= le;
            int64_t ov;
            if (in_s && !in_e) ov = le - s;
            else if (in_e && !in_s) ov = e - ls;
            else if (in_s && in_e) ov = e - s;
            else ov = le - ls;
            if (ov < mincols) continue;
            if (pair_items) pair_items[w] = (int32_t)i;
            if (visited) visited[i] = 1;
            w++;
        }
    }
    pair_offsets[nq] = w;
    return w;
}
