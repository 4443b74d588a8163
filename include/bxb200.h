/*
 * bxb200.h -- C ABI of libbxb200.so: the B200-native (sm_100a) replacement for bx-python's interval-intersection /
 * binned-bitset hot path.
 *
 * The reference has no FFI registry for this path; the boundary it offers is
 *   (1) the C ABI that lib/bx/bitset.pyx consumes:  src/binBits.h:7-26  (struct BinBits + binBits*) and
 *       src/kent/bits.h:13-59 (bit*), declared to Cython at lib/bx/bitset.pyx:14-69, and
 *   (2) the extension-module API of lib/bx/intervals/intersection.pyx:61-488 (IntervalNode/IntervalTree), which has
 *       no C layer underneath it.
 * Each entry point below names the reference interface it replaces.  Differences by design: calls are BATCHED
 * (arrays of positions/ranges/queries instead of one scalar per call), objects are opaque handles that own device
 * memory, and every function returns a status (0 = ok, <0 = error; message via bxg_last_error()) because the
 * reference C layer has no error channel at all (src/kent/common.c:3-16 exits on OOM).
 *
 * Conventions: plain C types only; `loc` arguments say where the caller's arrays live (BXG_HOST: host memory,
 * pageable or pinned; BXG_DEVICE: device memory of the current device).  All work is issued on the library's own
 * CUDA stream; functions that return data to host memory synchronise that stream before returning.
 * Not thread-safe per handle (the reference holds the GIL for every call, lib/bx/bitset.pyx has no nogil).
 */
#ifndef BXB200_H
#define BXB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BXG_HOST   0
#define BXG_DEVICE 1

#define BXG_OK            0
#define BXG_ERR_CUDA     -1   /* CUDA runtime / no device */
#define BXG_ERR_ARG      -2   /* invalid argument */
#define BXG_ERR_MISMATCH -3   /* operands of different size / geometry */
#define BXG_ERR_NCCL     -4
#define BXG_ERR_STATE    -5   /* call sequence error (e.g. fetch before find) */

/* ------------------------------------------------------------------------------------------------------------------
 * Runtime
 * ------------------------------------------------------------------------------------------------------------------ */
int         bxg_init(int device);                 /* bind this thread/process to `device`, create the stream       */
int         bxg_device_count(int *n);
int         bxg_device_info(char *name, int name_cap, int *sm_count, int64_t *total_mem, int *cc_major, int *cc_minor);
int         bxg_device_pci_bus_id(char *out, int cap);   /* "0000:1b:00.0": lets the host pin itself to the GPU's NUMA node */
const char *bxg_last_error(void);                 /* thread-local message of the last failing call                 */
const char *bxg_version(void);
int         bxg_sync(void);                       /* cudaStreamSynchronize(library stream)                          */
int         bxg_launch_count(int64_t *n);         /* kernels launched by this library since bxg_init / last reset   */
int         bxg_launch_count_reset(void);

/* pinned host memory + raw device buffers (used by the host shim for staging and by bench.py for resident inputs) */
int bxg_host_alloc(int64_t bytes, void **ptr);
int bxg_host_free(void *ptr);
int bxg_dev_alloc(int64_t bytes, void **dptr);
int bxg_dev_free(void *dptr);
int bxg_dev_memset(void *dptr, int value, int64_t bytes);           /* async on the library stream */
int bxg_memcpy_h2d(void *dptr, const void *hptr, int64_t bytes);   /* async on the library stream */
int bxg_memcpy_d2h(void *hptr, const void *dptr, int64_t bytes);   /* async; call bxg_sync() before reading */

/* device timing on the library stream (CUDA events) */
typedef struct bxg_timer bxg_timer_t;
int bxg_timer_create(bxg_timer_t **t);
int bxg_timer_free(bxg_timer_t *t);
int bxg_timer_start(bxg_timer_t *t);
int bxg_timer_stop(bxg_timer_t *t);
int bxg_timer_elapsed_ms(bxg_timer_t *t, float *ms);               /* synchronises on the stop event */
int bxg_l2_flush(void);                            /* overwrite a buffer larger than L2 (timing hygiene) */
/* pinned-memory copy rates of this process's GPU in GB/s: H2D alone, D2H alone, both directions at once (two streams);
 * `bytes` per copy, `reps` copies per measurement.  bench.py runs it on all ranks at once: the box's e2e ceiling. */
int bxg_copy_probe(int64_t bytes, int reps, double *h2d_gbs, double *d2h_gbs, double *bidir_gbs);
/* per-kernel CUDA-event timing of every launch the library makes: enable(1) clears and starts recording,
 * report() synchronises and writes "<kernel>\t<launches>\t<total_ms>\n" lines (bench.py's roofline uses it) */
int bxg_profile_enable(int on);
int bxg_profile_report(char *buf, int64_t cap);

/* ------------------------------------------------------------------------------------------------------------------
 * Bit sets.   Replaces struct BinBits + binBits* (src/binBits.h:7-26) and Bits + bit* (src/kent/bits.h:13-59).
 *
 * HBM layout: one dense LSB-first uint64 bitmap of exactly `size` bits (bit p -> word p>>6, bit p&63; tail bits of the
 * last word are kept 0), plus uint8 state[nbins] in {0=ALL_ZERO sentinel, 1=ALL_ONE sentinel, 2=allocated} that
 * tracks the reference's lazy-bin state machine (binBits.c:5-6, transitions at :67-128, :230-317) -- observable only
 * through count_range's ALL_ONE arithmetic (binBits.c:155,161), reproduced when strict != 0.
 * granularity == 0 creates a flat BitSet (bits.h semantics: no bins, no sentinel states).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct bxg_bits bxg_bits_t;

/* binBitsAlloc (binBits.c:8-17; float32 bin geometry) / bitAlloc (bits.c:51-56).  size in [1, 2^31-1]. */
int bxg_bits_create(int32_t size, int32_t granularity, bxg_bits_t **out);
int bxg_bits_free(bxg_bits_t *b);                                                   /* binBitsFree / bitFree   */
int bxg_bits_geometry(const bxg_bits_t *b, int32_t *size, int32_t *bin_size, int32_t *nbins);
int bxg_bits_clone(const bxg_bits_t *b, bxg_bits_t **out);                          /* bitClone (bits.c:58-66)  */
int bxg_bits_clear(bxg_bits_t *b);        /* bitClear (bits.c:190-195) / a fresh binBitsAlloc: all bits 0, all bins ALL_ZERO */

/* binBitsSetRange (binBits.c:98-128) / bitSetRange (bits.c:86-109), n ranges per call; ranges must satisfy
 * 0 <= start, 0 <= count, start+count <= size (the host shim raises the reference's IndexError first). */
int bxg_bits_set_ranges(bxg_bits_t *b, const int32_t *start, const int32_t *count, int64_t n, int loc);
/* binBitsSetOne / binBitsClearOne (binBits.c:67-96), n positions per call, value 1 = set, 0 = clear */
/* Genome-wide set_range: range i goes into sets[which[i]] -- the per-line `bitsets[chrom].set_range(start, end - start)`
 * of lib/bx/bitset_builders.py:40-53 for a whole file in one launch.  Entries with `which` outside [0, nsets), count <= 0
 * or start + count > size are skipped (the caller validates first, as bitset.pyx:184-189 does). */
int bxg_bits_set_ranges_multi(bxg_bits_t *const *sets, int32_t nsets, const int32_t *which, const int32_t *start,
                              const int32_t *count, int64_t n, int loc);
int bxg_bits_set_bits(bxg_bits_t *b, const int32_t *pos, int64_t n, int value, int loc);
/* binBitsReadOne (binBits.c:49-65) for n positions */
int bxg_bits_read(const bxg_bits_t *b, const int32_t *pos, int64_t n, uint8_t *out, int loc);

/* binBitsAnd / binBitsOr / binBitsNot (binBits.c:230-317); bitAnd/bitOr/bitXor/bitNot (bits.c:222-263). In place on a. */
int bxg_bits_and(bxg_bits_t *a, const bxg_bits_t *b);
int bxg_bits_or(bxg_bits_t *a, const bxg_bits_t *b);
int bxg_bits_xor(bxg_bits_t *a, const bxg_bits_t *b);      /* flat BitSet only in the reference (bitset.pyx:157-159) */
int bxg_bits_not(bxg_bits_t *a);
/* fused a &= b ; *count = popcount(a)   (bed_intersect_basewise.py:25-28 followed by a coverage count) */
int bxg_bits_and_count(bxg_bits_t *a, const bxg_bits_t *b, int64_t *count);
/* Genome-wide form of the loop `for chrom in bits1: bits1[chrom].iand(bits2[chrom])` (bed_intersect_basewise.py:25-28):
 * a[p] op= b[p] for n independent pairs in ONE persistent kernel launch.  op: 0 and, 1 or, 2 xor.
 * counts (host, n entries, may be NULL; op 0 only) receives popcount(a[p]) of each result. */
int bxg_bits_binop_batch(int op, bxg_bits_t *const *a, const bxg_bits_t *const *b, int32_t n, int64_t *counts);

/* binBitsCountRange (binBits.c:130-178) / bitCountRange (bits.c:118-141) for n (start,count) pairs -> int32 counts.
 * strict != 0 reproduces the reference's ALL_ONE-sentinel arithmetic; strict == 0 returns the true popcount. */
int bxg_bits_count_ranges(bxg_bits_t *b, const int32_t *start, const int32_t *count, int64_t n,
                          int32_t *out, int strict, int loc);
int bxg_bits_count_all(const bxg_bits_t *b, int64_t *count);                  /* count_range(0, size)          */
/* count_range(0, size) of every set of a genome in one launch (read off the rank tables, built on demand):
 * out[k * out_stride] = popcount(sets[k]) (0 for a NULL entry); device output (loc == BXG_DEVICE) is not synchronised. */
int bxg_bits_count_all_multi(bxg_bits_t *const *sets, int32_t nsets, int64_t *out, int64_t out_stride, int loc);
/* Genome-wide form of `bitsets[chrom].count_range(start, end-start)` per BED line (scripts/bed_intersect.py:46-53):
 * query i addresses sets[which[i]]; which outside [0,nsets), a NULL sets[k] (chromosome without a bitset: `fields[0] in
 * bitsets` is false, :53) or an out-of-range span yields 0.  One launch. */
int bxg_bits_count_ranges_multi(bxg_bits_t *const *sets, int32_t nsets, const int32_t *which, const int32_t *start,
                                const int32_t *count, int64_t n, int32_t *out, int strict, int loc);

/* Per-chromosome counters of a whole file -- what the multi-GPU runs reduce: for every key k in [0, nkeys)
 *   stats[2k] += #(i : key[i] == k && val[i] >= threshold),   stats[2k+1] += sum(val[i] : key[i] == k).
 * With val = the counts of bxg_bits_count_ranges_multi and threshold = mincols this is the number of lines
 * scripts/bed_intersect.py:53 prints per chromosome; with val = the window counts of bxg_aggregate_multi and threshold 1
 * the number of windows aggregate_scores_in_intervals.py:126 gives an average.  stats is ACCUMULATED into (zero it first);
 * entries with a key outside [0, nkeys) are ignored.  nkeys <= 1024. */
int bxg_group_stats_i32(const int32_t *key, const int32_t *val, int64_t n, int32_t nkeys, int32_t threshold,
                        int64_t *stats /* 2 * nkeys */, int loc);

/* binBitsFindSet / binBitsFindClear (binBits.c:180-228) ; bitFindSet/Clear with an end bound (bits.c:143-190).
 * first position p in [start, end) whose bit == val, else `end`; end <= size. */
int bxg_bits_next(const bxg_bits_t *b, int32_t start, int32_t end, int val, int32_t *out);
/* The run-extraction idiom  start=next_set(end); end=next_clear(start)  (bed_intersect_basewise.py:30-38,
 * lib/bx/bitset_utils.py:34-43) for the whole bitmap: first call counts, second fills (host arrays). */
int bxg_bits_runs_count(bxg_bits_t *b, int64_t *nruns);
int bxg_bits_runs_fetch(bxg_bits_t *b, int32_t *starts, int32_t *ends, int64_t nruns);

/* The generators bits_set_in_range / bits_clear_in_range (lib/bx/intervals/operations/__init__.py:10-33) for n ranges at
 * once: the maximal runs of bits == val inside [start[i], end[i]) (clipped to [0, size)), in order -- the `pieces` that
 * operations/intersect.py:62-70 / subtract.py:66-72 emit per interval.  offsets (HOST, n+1 entries) receives the CSR
 * offsets of the runs of each range; the runs stay on the device until ..._fetch copies them to host arrays. */
int bxg_bits_runs_in_ranges(bxg_bits_t *b, const int32_t *start, const int32_t *end, int64_t n, int val, int loc,
                            int64_t *offsets, int64_t *total);
int bxg_bits_runs_in_ranges_fetch(bxg_bits_t *b, int32_t *starts, int32_t *ends, int64_t total);

/* test / interchange helpers */
int bxg_bits_states(const bxg_bits_t *b, uint8_t *out /* nbins */);
int bxg_bits_export_words(const bxg_bits_t *b, uint64_t *out /* ceil(size/64) */);
int bxg_bits_import_words(bxg_bits_t *b, const uint64_t *in /* ceil(size/64) */);   /* marks every bin allocated */
int bxg_bits_device_words(const bxg_bits_t *b, const uint64_t **dptr, int64_t *nwords);

/* ------------------------------------------------------------------------------------------------------------------
 * Interval index.   Replaces IntervalNode/IntervalTree of lib/bx/intervals/intersection.pyx.
 *
 * One index holds `ntrees` independent trees (one per chromosome: the reference keeps dict[chrom] -> IntervalTree,
 * scripts/bed_count_overlapping.py:17-25); ntrees == 1 is a plain IntervalTree.  Items are identified by their
 * position in the arrays passed to bxg_itree_build (== insertion order, which fixes the reference's tie order,
 * intersection.pyx:110-116).
 * HBM layout: items radix-sorted by (tree, start, end>start, +/-index) into S[], E[], I[] (= in-order traversal of the
 * reference treap), a per-tree prefix-max of E (first possible hit), a 32-ary max-of-E hierarchy for skipping, and a
 * sampled splitter table that the find kernels stage into shared memory with a 1-D TMA bulk copy.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct bxg_itree bxg_itree_t;

int bxg_itree_create(bxg_itree_t **out);
int bxg_itree_free(bxg_itree_t *t);
/* IntervalTree.insert x n (intersection.pyx:388-395 -> IntervalNode.insert :103-138).  tree == NULL means all items
 * belong to tree 0.  Rebuilds the whole index (the host shim queues inserts and builds lazily). */
int bxg_itree_build(bxg_itree_t *t, const int32_t *tree, const int32_t *start, const int32_t *end,
                    int64_t n, int32_t ntrees, int loc);
int bxg_itree_size(const bxg_itree_t *t, int64_t *n, int32_t *ntrees);
/* in-order traversal (IntervalNode.traverse, intersection.pyx:262-268): perm[k] = item index of the k-th node;
 * tree_offsets[ntrees+1] (may be NULL) delimits each tree's slice of perm. */
int bxg_itree_order(const bxg_itree_t *t, int32_t *perm, int64_t *tree_offsets);

/* IntervalTree.find x nq (intersection.pyx:400-406 -> _intersect :180-189): for query q every item of tree qtree[q]
 * with  end > qs[q] && start < qe[q], in in-order sequence.  Results stay in device buffers owned by the index
 * (CSR: int64 offsets[nq+1], int32 hits[total]) until the next find; *total receives the number of hits.
 * qtree == NULL means tree 0 for every query. */
int bxg_itree_find(bxg_itree_t *t, const int32_t *qtree, const int32_t *qs, const int32_t *qe,
                   int64_t nq, int loc, int64_t *total);
int bxg_itree_fetch(bxg_itree_t *t, int64_t *offsets /* nq+1 */, int32_t *hits /* total */);   /* to host */
/* find implementation: 1 = single-pass kernel (count + decoupled look-back scan + fill in one launch), 0 = count kernel /
 * CUB scan / fill kernel back to back, 2 = the same three passes with the count of chunk k+1 and the fill of chunk k
 * overlapped on two streams (measured slower, kept for A/B), -1 (default) = auto: three-pass for bxg_itree_find,
 * single-pass for the chunk-pipelined bxg_itree_find_host (the measured winners).  Same results in every mode; env
 * BXB200_FIND_MODE overrides. */
int bxg_set_find_mode(int mode);
/* The same find for HOST query arrays with the PCIe copies overlapped with the kernels (queries are processed in
 * chunks on copy-in / compute / copy-out streams).  *offsets (nq+1 int64) and *hits (*total int32) point into pinned
 * host buffers owned by the index and stay valid until its next find or bxg_itree_free. */
int bxg_itree_find_host(bxg_itree_t *t, const int32_t *qtree, const int32_t *qs, const int32_t *qe, int64_t nq,
                        const int64_t **offsets, const int32_t **hits, int64_t *total);
/* bxg_itree_find_host with int32 CSR offsets (4 instead of 8 bytes per query on the PCIe-bound return leg; ~12 % less
 * device-to-host traffic at 6.5 hits per query).  Fails with BXG_ERR_MISMATCH when the hits do not fit (total >= 2^31):
 * call the 64-bit form then. */
int bxg_itree_find_host32(bxg_itree_t *t, const int32_t *qtree, const int32_t *qs, const int32_t *qe, int64_t nq,
                          const int32_t **offsets, const int32_t **hits, int64_t *total);

/* Up to 32 queries with the latency of one launch and one synchronise: the scalar `IntervalTree.find(start, end)`
 * (intersection.pyx:400-406) called once per line by scripts such as bed_count_overlapping.py:27-33.  Host arrays only;
 * the CSR is written by the kernel into mapped pinned memory owned by the index (valid until its next find).  Falls
 * back to bxg_itree_find_host when the queries have more than 65536 hits in total. */
int bxg_itree_find_small(bxg_itree_t *t, const int32_t *qtree, const int32_t *qs, const int32_t *qe, int32_t nq,
                         const int64_t **offsets, const int32_t **hits, int64_t *total);
/* One query, plain integers in: returns the number of hits (>= 0) or a negative status; *hits points into the index's
 * mapped result buffer (valid until its next find).  The thinnest form of IntervalTree.find (intersection.pyx:400-406). */
int64_t bxg_itree_find1(bxg_itree_t *t, int32_t tree, int32_t start, int32_t end, const int32_t **hits);
/* bxg_itree_find1 through the lingering find server: a one-warp kernel that stays resident after a scalar find and takes
 * the next ones from a request line in mapped host memory (no launch per call); it leaves after ~100 us without a request,
 * on any index rebuild / free, or when switched off here.  Same results either way; env BXB200_FIND_SERVER=0/1 overrides
 * the built-in default.  stats: launches of the server kernel, requests it was sent, whether one is resident now. */
int bxg_set_find_server(int on);       /* 1 on, 0 off, -1 back to the default */
int bxg_find_server_stats(int64_t *launches, int64_t *requests, int32_t *alive);
int bxg_itree_result_dev(const bxg_itree_t *t, const int64_t **d_offsets, const int32_t **d_hits, int64_t *nq,
                         int64_t *total);
/* len(find(...)) only (scripts/bed_count_overlapping.py:27-33): int32 counts[nq] written to `counts` (host or device
 * per loc); *total (may be NULL) receives the sum.  Host arrays of two or more pipeline chunks are copied in, counted and
 * copied out chunk by chunk on three streams (16 bytes per query over PCIe instead of the ~42 of the full CSR). */
int bxg_itree_count(bxg_itree_t *t, const int32_t *qtree, const int32_t *qs, const int32_t *qe,
                    int64_t nq, int loc, int32_t *counts, int64_t *total);

/* IntervalNode.left / right (intersection.pyx:192-260) = IntervalTree.before / after (:408-426): for query q the
 * reference-ordered list of at most n[q] neighbours of position pos[q] within max_dist[q].
 * dir = 0: before (left), 1: after (right).  CSR result fetched like find. */
int bxg_itree_neighbors(bxg_itree_t *t, const int32_t *qtree, const int32_t *pos, const int32_t *n,
                        const int32_t *max_dist, int64_t nq, int dir, int loc, int64_t *total);

/* Overlap join (lib/bx/intervals/operations/join.py:14-75 over operations/quicksect.py:11-125): for every left
 * interval q the items of the index with start < item.end && end > item.start (quicksect.py:115-121) whose `overlap`
 * by the case analysis of join.py:35-50 is >= mincols.  istart/iend are the index's items in INSERTION order (the ids
 * `find` returns index them).  Results stay on the device, in ONE library-wide buffer, until bxg_itree_join_fetch (fetch
 * before the next join; like every entry point this is not thread-safe): CSR pair_offsets[nq+1] /
 * pair_items[total] (item ids, index order within a left interval -- the reference's own order is a random treap
 * walk) and visited[n] = 1 for every item kept at least once (join.py:54, drives the left-fill pass :62-75). */
int bxg_itree_join(bxg_itree_t *t, const int32_t *qtree, const int32_t *qs, const int32_t *qe, int64_t nq,
                   const int32_t *istart, const int32_t *iend, int32_t mincols, int loc, int64_t *total_pairs);
int bxg_itree_join_fetch(int64_t *pair_offsets, int32_t *pair_items, uint8_t *visited);

/* ------------------------------------------------------------------------------------------------------------------
 * aggregate_scores_in_intervals inner loop (scripts/aggregate_scores_in_intervals.py:107-134 over
 * BinnedArray.get, lib/bx/binned_array.py:89-94).  scores: dense float32, position origin+i, NaN = unset.
 * Per window [ws,we): strict left-to-right float32 sum of the scores that are non-zero, non-NaN and not masked;
 * count; min; max; avg = sum/count (float32).  count==0 -> avg/min/max = NaN.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct bxg_scores bxg_scores_t;
int bxg_scores_create(const float *scores, int64_t n, int32_t origin, int loc, bxg_scores_t **out);
int bxg_scores_free(bxg_scores_t *s);
int bxg_aggregate(const bxg_scores_t *s, const bxg_bits_t *mask /* or NULL */,
                  const int32_t *ws, const int32_t *we, int64_t nw, int loc,
                  float *sum, float *avg, int32_t *count, float *mn, float *mx);
/* Genome-wide form: window w reads track tracks[wtrack[w]] (the script's `scores_by_chrom[chrom]`, :115; a track id
 * outside [0,ntracks), or a NULL tracks[k], behaves like a chromosome without scores).  masks may be NULL, or hold NULL
 * entries.  One launch. */
int bxg_aggregate_multi(const bxg_scores_t *const *tracks, const bxg_bits_t *const *masks, int32_t ntracks,
                        const int32_t *wtrack, const int32_t *ws, const int32_t *we, int64_t nw, int loc,
                        float *sum, float *avg, int32_t *count, float *mn, float *mx);

/* ------------------------------------------------------------------------------------------------------------------
 * Score sources.
 *
 * BinnedArray (lib/bx/binned_array.py:72-136: get :89-94, set :96-100, get_range :102-127) as a dense device track:
 * a bin that was never written reads as `default`, so a track pre-filled with `fill` is observably identical.
 * bxg_scores_set_spans replaces the per-base assignment loop of load_scores_wiggle
 * (scripts/aggregate_scores_in_intervals.py:60-70 over wiggle.Reader, lib/bx/wiggle.py:71-85): span i assigns val[i]
 * to every position of [start[i], end[i]) (end == NULL: single positions), spans are applied IN ORDER, i.e. where
 * spans overlap the last one wins; empty spans (end <= start) assign nothing.  Spans must lie inside the track
 * (bxg_scores_reserve first).
 * ------------------------------------------------------------------------------------------------------------------ */
int bxg_scores_alloc(int64_t n, int32_t origin, float fill, bxg_scores_t **out);       /* BinnedArray.__init__        */
int bxg_scores_info(const bxg_scores_t *s, int64_t *n, int32_t *origin, float *fill);
int bxg_scores_reserve(bxg_scores_t *s, int64_t n);            /* grow to n cells; new cells = fill (init_bin, :84-87) */
int bxg_scores_set_spans(bxg_scores_t *s, const int32_t *start, const int32_t *end /* or NULL */, const float *val,
                         int64_t n, int loc);
int bxg_scores_write(bxg_scores_t *s, int64_t start, const float *vals, int64_t n, int loc);     /* contiguous cells  */
int bxg_scores_get(const bxg_scores_t *s, const int32_t *pos, int64_t n, float *out, int loc);   /* get, batched      */
int bxg_scores_get_range(const bxg_scores_t *s, int64_t start, int64_t end, float *out /* host, end-start floats */);
int bxg_scores_device(const bxg_scores_t *s, const float **dptr, int64_t *n);

/* bigWig summary (lib/bx/bbi/bbi_file.pyx:66-111, SummarizedData.accumulate_interval_value applied to every interval
 * of the batch in order; lib/bx/bbi/bigwig_file.pyx:93-108,176-185 is the caller that feeds it the file's intervals):
 * `size` bins of (rend - rstart) / size bases over [rstart, rend); per bin float64 valid_count, sum, sum of squares
 * (weights and products in the reference's expression order), min and max.  The five arrays are IN/OUT: the batch
 * is accumulated into the state they hold (zeros from SummarizedData.__init__, :70-78; +inf / -inf min / max from
 * SummarizingBlockHandler, bigwig_file.pyx:98-105), so several calls continue one summary exactly like repeated
 * accumulate_interval_value calls.  Coordinates are bits32 in the reference; values < 2^31 here. */
int bxg_summarize(const int32_t *start, const int32_t *end, const float *val, int64_t n, int loc, uint32_t rstart,
                  uint32_t rend, int32_t size, double *valid_count, double *min_val, double *max_val, double *sum_data,
                  double *sum_squares);

/* ------------------------------------------------------------------------------------------------------------------
 * Multi-GPU: one process per GPU; chromosomes are sharded with no data-path exchange; only the final per-chromosome
 * counters are summed (NCCL all-reduce over NVLink).  NCCL is dlopen()ed on first use.
 * ------------------------------------------------------------------------------------------------------------------ */
#define BXG_UNIQUE_ID_BYTES 128
int bxg_comm_unique_id(char id[BXG_UNIQUE_ID_BYTES]);                 /* rank 0; ship to the other ranks      */
int bxg_comm_init(const char id[BXG_UNIQUE_ID_BYTES], int nranks, int rank);
int bxg_comm_allreduce_i64(int64_t *buf, int64_t n);                   /* in-place sum over ranks (host buffer) */
int bxg_comm_allreduce_max_f64(double *buf, int64_t n);                /* in-place max over ranks (timings)     */
/* the same sum on a DEVICE buffer, enqueued on the library stream without a host synchronisation (the reduce of the
 * per-chromosome counters inside a device-timed step) */
int bxg_comm_allreduce_i64_dev(int64_t *dbuf, int64_t n);
int bxg_comm_barrier(void);
int bxg_comm_destroy(void);

#ifdef __cplusplus
}
#endif
#endif /* BXB200_H */
