// osp_kway.cuh -- k-way merge of the pre-sorted ways of an output row.
//
// The reference's earlier design merged the ways of a row instead of sorting their concatenation: `merge2way`
// (simulator/SimSpGEMM.cpp:306-327) walks two sorted lists, `mergeHardware` (:411-441) stacks it into a 64-way, six-layer
// merge tree, `merge` (:445-517) schedules it.  The ways ARE sorted by construction -- way r of row i is A(i,k_r) * B(k_r,:),
// emitted in the order of B's row (SimOuterSPACE.cpp:85-92), ascending k -- which is what this kernel uses for the rows
// where neither a register sort nor a dense accumulator fits: 4 097 .. 32 768 partial products made of a few long runs
// (config 3: 188 879 rows, ~10 runs each, compression 1.2 -- DESIGN.md section 10).
//
// One CTA per row, no tree and no intermediate buffers: every partial product computes its final RANK in the merged
// row directly -- its index in its own run plus, for every other run, the number of elements that sort before it
// (binary searches over the runs, which sit in L1/L2: the row's bin is 32-256 KB): upper bound in the runs of smaller k,
// lower bound in the runs of larger k, so equal columns keep their k order (the order `merge2way` would give a chain of
// left-deep merges).  The element goes to its rank in a per-CTA scratch row; a second phase left-folds equal columns in
// that order with separately rounded adds (deduplicateCOO's fold, SimSpGEMM.cpp:519-535) and compacts the row over the
// start of its bin, uniq[row] = survivors: what k_merge_chain expects of a long row.
// Divergence from `merge2way`, on purpose: it adds equal keys INSIDE the tree ((r0 + r1) + (r2 + r3)); the oracle's values are
// the k-ordered left fold (((r0 + r1) + r2) + r3), so nothing is added before the ranks are final.
#pragma once
#include "osp_device.cuh"

namespace osp {

constexpr int KW_THREADS = 512;
constexpr uint32_t KW_MAX_LEN = 32768;      // partial products of a row (scratch row per CTA: 256 KB of global memory, L2-resident)
constexpr uint32_t KW_MAX_WAYS = 64;        // MAX_MERGE_K of the reference's merge tree (SimSpGEMM.cpp:409-410)

// Does the k-way merge take this row?  (shared with k_merge_xl, which then leaves it alone)
__device__ __forceinline__ bool kway_takes(uint64_t len, uint64_t ways) { return len <= KW_MAX_LEN && ways >= 1 && ways <= KW_MAX_WAYS; }

__global__ void __launch_bounds__(KW_THREADS)
k_merge_ways(const uint64_t *__restrict__ a_pos, const uint64_t *__restrict__ run_off, const uint64_t *__restrict__ row_bin,
             uint64_t bin_base, Elem *bins, uint32_t *uniq, const uint32_t *__restrict__ xl_list, DevScalars *sc,
             Elem *scratch_all, uint64_t row_lo, uint64_t row_hi) {
    __shared__ uint32_t s_start[KW_MAX_WAYS + 1];          // start of every way inside the row (s_start[R] = len)
    __shared__ uint32_t warp_sums[33];
    __shared__ uint32_t s_x;
    const uint32_t tid = threadIdx.x;
    Elem *scratch = scratch_all + uint64_t(blockIdx.x) * KW_MAX_LEN;
    const uint32_t n_xl = sc->n_xl;
    while (true) {
        __syncthreads();
        if (tid == 0) {
            uint32_t x;
            while (true) {                                 // next listed row of this row block that the merge takes
                x = atomicAdd(&sc->kw_ticket, 1u);
                if (x >= n_xl) break;
                const uint64_t r = xl_list[x];
                if (r >= row_lo && r < row_hi && kway_takes(row_bin[r + 1] - row_bin[r], a_pos[r + 1] - a_pos[r])) break;
            }
            s_x = x;
        }
        __syncthreads();
        const uint32_t x = s_x;
        if (x >= n_xl) break;
        const uint64_t row = xl_list[x];
        const uint64_t p0 = a_pos[row];
        const uint32_t R = uint32_t(a_pos[row + 1] - p0);
        const uint64_t b0 = row_bin[row];
        const uint32_t len = uint32_t(row_bin[row + 1] - b0);
        Elem *bin = bins + (b0 - bin_base);
        if (tid <= R) s_start[tid] = uint32_t(run_off[p0 + tid] - b0);
        __syncthreads();
        // ---- phase 1: every partial product to its rank ----
        for (uint32_t p = tid; p < len; p += KW_THREADS) {
            uint32_t r = 0;                                // the way of position p: s_start[r] <= p < s_start[r + 1]
            {
                uint32_t hi = R;
                while (hi - r > 1) { const uint32_t mid = (r + hi) >> 1; if (s_start[mid] <= p) r = mid; else hi = mid; }
            }
            const Elem e = bin[p];
            uint32_t rank = p - s_start[r];
            for (uint32_t q = 0; q < R; q++) {
                if (q == r) continue;
                uint32_t lo = s_start[q], hi = s_start[q + 1];
                const uint32_t base = lo;
                // earlier ways: elements <= e.idx come first (upper bound); later ways: elements < e.idx (lower bound)
                const uint32_t key = q < r ? e.idx : e.idx - 1u;       // count of elements <= key; e.idx == 0 and q > r: none
                if (q > r && e.idx == 0) continue;
                while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (bin[mid].idx <= key) lo = mid + 1; else hi = mid; }
                rank += lo - base;
            }
            scratch[rank] = e;
        }
        __threadfence_block();
        __syncthreads();
        // ---- phase 2: left fold of equal columns in rank (= k) order, compaction over the start of the bin ----
        uint32_t produced = 0;
        for (uint32_t s0 = 0; s0 < len; s0 += KW_THREADS) {
            const uint32_t s = s0 + tid;
            bool head = false;
            Elem e; e.idx = 0; e.val = 0.f;
            if (s < len) {
                e = scratch[s];
                head = s == 0 || scratch[s - 1].idx != e.idx;
            }
            float sum = e.val;
            if (head)
                for (uint32_t u = s + 1; u < len; u++) {
                    const Elem n = scratch[u];
                    if (n.idx != e.idx) break;
                    sum = __fadd_rn(sum, n.val);
                }
            uint32_t total;
            const uint32_t rank = block_exclusive_scan(head ? 1u : 0u, warp_sums, total);
            if (head) { Elem o; o.idx = e.idx; o.val = sum; bin[produced + rank] = o; }
            produced += total;
        }
        if (tid == 0) uniq[row] = produced;
    }
}

}  // namespace osp
