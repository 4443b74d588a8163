// join.cu -- the overlap join of lib/bx/intervals/operations/join.py:14-75 on top of the device interval index.
//
// The reference builds a quicksect treap of the right-hand set (operations/quicksect.py:11-125), and for every left
// interval collects the nodes with `start < node.end and end > node.start` (:115-121 -- the same predicate as
// IntervalTree.find), computes an `overlap` per node by the four-way case analysis of join.py:35-50 (note the
// INCLUSIVE `range(interval.start, interval.end + 1)` membership tests), keeps the pairs with overlap >= mincols and
// marks kept nodes `visited` (:54) for the left-fill pass (:62-75).
//
// Here: bxg_itree_find produces the CSR of candidate pairs; k_join<false> counts the kept pairs per left interval,
// one CUB scan turns the counts into offsets, k_join<true> writes the kept item ids and sets visited[item].  The
// treap's report order (a pre-order walk of a randomly balanced tree) is not reproducible even between two runs of
// the reference, so kept items come out in index order; the per-left SETS are identical.
#include <cub/cub.cuh>

#include "common.cuh"

using namespace bxg;

namespace {

struct JoinState {
    int32_t *d_kept = nullptr;      // kept pairs per left interval
    int64_t *d_poff = nullptr;      // nq + 1
    int32_t *d_items = nullptr;     // kept item ids
    uint8_t *d_visited = nullptr;   // per item
    int64_t kept_cap = 0, poff_cap = 0, items_cap = 0, vis_cap = 0;
    int64_t nq = -1, n = 0, total = 0;
};
JoinState g_join;

// join.py:35-50 with L = [ls, le] inclusive on both sides
__device__ __forceinline__ int64_t join_overlap(int64_t ls, int64_t le, int64_t s, int64_t e) {
    const bool in_s = s >= ls && s <= le, in_e = e >= ls && e <= le;
    if (in_s && !in_e) return le - s;
    if (in_e && !in_s) return e - ls;
    if (in_s && in_e) return e - s;
    return le - ls;
}

template <bool FILL>
__global__ void __launch_bounds__(256)
k_join(const int64_t *__restrict__ off, const int32_t *__restrict__ hits, const int32_t *__restrict__ qs,
       const int32_t *__restrict__ qe, const int32_t *__restrict__ istart, const int32_t *__restrict__ iend, int64_t nq,
       int64_t mincols, int32_t *__restrict__ kept, const int64_t *__restrict__ poff, int32_t *__restrict__ items,
       uint8_t *__restrict__ visited) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride) {
        const int64_t a = off[q], b = off[q + 1];
        const int64_t ls = __ldg(qs + q), le = __ldg(qe + q);
        int64_t w = FILL ? poff[q] : 0;
        int32_t k = 0;
        for (int64_t h = a; h < b; h++) {
            const int32_t it = __ldg(hits + h);
            if (join_overlap(ls, le, __ldg(istart + it), __ldg(iend + it)) >= mincols) {
                if (FILL) {
                    items[w++] = it;
                    visited[it] = 1;            // same value from every writer: no ordering needed
                } else {
                    k++;
                }
            }
        }
        if (!FILL) kept[q] = k;
    }
}

int grow(void **p, int64_t *cap, int64_t need, size_t elt) {
    if (need <= *cap) return BXG_OK;
    if (*p) {
        BXG_CUDA(cudaStreamSynchronize(ctx().stream));
        BXG_CUDA(cudaFree(*p));
        *p = nullptr;
        *cap = 0;
    }
    const int64_t c = need + need / 4 + 64;
    BXG_CUDA(cudaMalloc(p, (size_t)c * elt));
    *cap = c;
    return BXG_OK;
}

}  // namespace

extern "C" {

int bxg_itree_join(bxg_itree_t *t, const int32_t *qtree, const int32_t *qs, const int32_t *qe, int64_t nq,
                   const int32_t *istart, const int32_t *iend, int32_t mincols, int loc, int64_t *total_pairs) {
    BXG_TRY(ensure_init());
    JoinState &j = g_join;
    j.nq = -1;
    int64_t n = 0;
    int32_t ntrees = 0;
    BXG_TRY(bxg_itree_size(t, &n, &ntrees));
    if (nq < 0) return set_error(BXG_ERR_ARG, "nq < 0");
    Context &c = ctx();
    BXG_TRY(grow((void **)&j.d_visited, &j.vis_cap, n > 0 ? n : 1, 1));
    BXG_CUDA(cudaMemsetAsync(j.d_visited, 0, (size_t)(n > 0 ? n : 1), c.stream));
    j.n = n;
    j.total = 0;
    if (nq == 0 || n == 0) {
        j.nq = nq;
        BXG_TRY(grow((void **)&j.d_poff, &j.poff_cap, nq + 1, 8));
        BXG_CUDA(cudaMemsetAsync(j.d_poff, 0, (size_t)(nq + 1) * 8, c.stream));
        if (total_pairs) *total_pairs = 0;
        return BXG_OK;
    }
    int64_t cand = 0;
    BXG_TRY(bxg_itree_find(t, qtree, qs, qe, nq, loc, &cand));
    const int64_t *d_off;
    const int32_t *d_hits;
    BXG_TRY(bxg_itree_result_dev(t, &d_off, &d_hits, nullptr, nullptr));
    // the left coordinates and the right-hand items in insertion order (hit ids index these)
    const void *dqs, *dqe, *dis, *die;
    BXG_TRY(stage_in(4, qs, (size_t)nq * 4, loc, &dqs));
    BXG_TRY(stage_in(5, qe, (size_t)nq * 4, loc, &dqe));
    BXG_TRY(stage_in(6, istart, (size_t)n * 4, loc, &dis));
    BXG_TRY(stage_in(7, iend, (size_t)n * 4, loc, &die));
    BXG_TRY(grow((void **)&j.d_kept, &j.kept_cap, nq + 1, 4));
    BXG_TRY(grow((void **)&j.d_poff, &j.poff_cap, nq + 1, 8));
    const int g = grid_for(cdiv(nq, 256), 8);
    BXG_LAUNCH(k_join<false>, g, 256, 0, d_off, d_hits, (const int32_t *)dqs, (const int32_t *)dqe, (const int32_t *)dis,
               (const int32_t *)die, nq, (int64_t)mincols, j.d_kept, (const int64_t *)nullptr, (int32_t *)nullptr,
               (uint8_t *)nullptr);
    BXG_CUDA(cudaMemsetAsync(j.d_kept + nq, 0, 4, c.stream));           // the scan reads nq + 1 counts
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, j.d_kept, j.d_poff, nq + 1, c.stream);
    void *tmp;
    BXG_TRY(scratch(3, tmp_bytes, &tmp));
    prof_begin("cub::DeviceScan::ExclusiveSum(join)");
    BXG_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, j.d_kept, j.d_poff, nq + 1, c.stream));
    prof_end();
    c.launches++;
    BXG_CUDA(cudaMemcpyAsync(c.mailbox + 20, j.d_poff + nq, 8, cudaMemcpyDeviceToHost, c.stream));
    BXG_CUDA(cudaStreamSynchronize(c.stream));
    j.total = c.mailbox[20];
    BXG_TRY(grow((void **)&j.d_items, &j.items_cap, j.total > 0 ? j.total : 1, 4));
    if (j.total)
        BXG_LAUNCH(k_join<true>, g, 256, 0, d_off, d_hits, (const int32_t *)dqs, (const int32_t *)dqe,
                   (const int32_t *)dis, (const int32_t *)die, nq, (int64_t)mincols, (int32_t *)nullptr,
                   (const int64_t *)j.d_poff, j.d_items, j.d_visited);
    j.nq = nq;
    if (total_pairs) *total_pairs = j.total;
    return BXG_OK;
}

int bxg_itree_join_fetch(int64_t *pair_offsets, int32_t *pair_items, uint8_t *visited) {
    JoinState &j = g_join;
    if (j.nq < 0) return set_error(BXG_ERR_STATE, "bxg_itree_join must be called first");
    Context &c = ctx();
    if (pair_offsets) BXG_CUDA(cudaMemcpyAsync(pair_offsets, j.d_poff, (size_t)(j.nq + 1) * 8, cudaMemcpyDeviceToHost, c.stream));
    if (pair_items && j.total) BXG_CUDA(cudaMemcpyAsync(pair_items, j.d_items, (size_t)j.total * 4, cudaMemcpyDeviceToHost, c.stream));
    if (visited && j.n) BXG_CUDA(cudaMemcpyAsync(visited, j.d_visited, (size_t)j.n, cudaMemcpyDeviceToHost, c.stream));
    BXG_CUDA(cudaStreamSynchronize(c.stream));
    return BXG_OK;
}

}  // extern "C"
