// osp_longrows.cuh -- fused band sweep for long output rows over wide column ranges.
//
// STATUS: logic validated on the CPU emulation of the execution model (tests/cusim: tests/test_longrows_sim.py with small
// template parameters, tests/test_engine_sim.py and tools/fuzz_engine_sim.py end to end through osp_spgemm with the
// engine's own instantiation -- bit-exact against the oracle, also under AddressSanitizer); compiled for sm_100a with the rest of the
// engine and launched by osp_spgemm ONLY when the caller opts in (OSP_LONGROW_SWEEP flag or environment variable, see
// include/osp_b200.h): the round's GPU budget was spent before it could be run on a B200, so the default path still
// sends long rows through k_multiply + k_merge_xl.  GPU parity tests for the opt-in path: tests/test_gpu_zzz_sweep.py.
//
// Why: a long row of C = A*B (config 3: 295 801 rows carry 98.7 % of the 2.09e10 partial products) is made of FEW LONG
// SORTED runs -- run r = A(i,k_r) * B(k_r,:), ascending k, columns ascending and distinct inside a run.  Writing those
// partial products to bins and folding them through a per-CTA accumulator of `cols` floats in global memory costs a
// random DRAM sector per product (profiles/README.md).  Here nothing is materialised: the row is swept one band of
// LR_BAND columns at a time with the accumulator in shared memory; a run contributes to a band one contiguous
// segment (runs are sorted) whose bounds come from a band index of B built once per call (k_long_bands: for every
// row k of B the position of the first column of every band -- two adjacent 4-byte loads per (run, band), no search
// and no per-run state during the sweep), so every element of B is read once per use, from L2.
//   k_long_count: the sweep with a bitmap only -> the exact number of non-zeros of every long row, so that C can be
//                 allocated exactly and every long row's place known before any value is computed (not used by the
//                 engine yet: the first integration hands the merged row over through the start of its bin).
//   k_long_fill:  the sweep with values.  Products of a chunk that hit the same column are serialised by the
//                 racing-minimum arbitration of k_merge_dense (lowest flat position = lowest (k, position) first), so
//                 every column is summed in ascending k with separately rounded products and adds: the bits of the
//                 reference's left fold.  acc[] starts at -0.0f (x + -0 = x for every x): no first-touch test on the add.
// Runs are taken in groups of LR_RUNS (the prefix of their segment lengths lives in shared memory); groups of a band
// are visited in ascending k, which keeps the order.  The elements of the next chunk are loaded before the current one
// is arbitrated.
#pragma once
#include "osp_device.cuh"

namespace osp {

// Band index of B: bandptr[k * (n_bands + 1) + b] = position in b_data of the first element of row k whose column is
// >= b * band (b = n_bands: the end of the row).  One thread per entry, a binary search inside the row.
__global__ void k_long_bands(const uint64_t *__restrict__ b_pos, const Elem *__restrict__ b_data, uint64_t n_k, uint32_t band,
                             uint32_t n_bands, uint32_t *__restrict__ bandptr) {
    const uint64_t t = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    const uint64_t nb1 = uint64_t(n_bands) + 1;
    if (t >= n_k * nb1) return;
    const uint64_t k = t / nb1, b = t % nb1;
    uint64_t lo = b_pos[k], hi = b_pos[k + 1];
    if (b < n_bands) {
        const uint64_t target = b * uint64_t(band);
        while (lo < hi) {
            const uint64_t mid = lo + (hi - lo) / 2;
            if (b_data[mid].idx < target) lo = mid + 1; else hi = mid;
        }
    } else {
        lo = hi;
    }
    bandptr[t] = uint32_t(lo);                                             // nnz(B) < 2^32
}

// Shared-memory layout of one CTA: band accumulator | owner | presence bitmap | run group | scan scratch
template <int BAND, int RUNS, bool VALUES> struct LongRowSmem {
    static constexpr size_t acc = 0;                                              // float[BAND]      (VALUES)
    static constexpr size_t owner = acc + (VALUES ? size_t(BAND) * 4 : 0);        // uint16_t[BAND]   (VALUES)
    static constexpr size_t bits = owner + (VALUES ? size_t(BAND) * 2 : 0);       // uint32_t[BAND/32]
    static constexpr size_t pre = bits + size_t(BAND) / 8;                        // uint32_t[RUNS+1] exclusive prefix of the segment lengths
    static constexpr size_t lo = pre + size_t(RUNS + 1) * 4;                      // uint32_t[RUNS]   segment start inside row k of B (absolute in b_data)
    static constexpr size_t aval = lo + size_t(RUNS) * 4;                         // float[RUNS]      A(i,k)  (VALUES)
    static constexpr size_t sums = aval + (VALUES ? size_t(RUNS) * 4 : 0);        // uint32_t[34] scan scratch | ticket, more | row (uint64)
    static constexpr size_t bytes = ((sums + 36 * 4 + 7) & ~size_t(7)) + 8;
};

// Where the rows come from and where they go.
//   next(x, row): called by thread 0 -- the next listed row, false when the list is exhausted;
//   out(x, row):  where the merged row is written (VALUES);  done(x, row, n): called by thread 0 with the row's nnz.
struct LongRowsListed {                 // an explicit list with explicit places (tests; exact allocation of C)
    const uint32_t *rows; uint32_t n_rows; unsigned int *ticket;
    uint32_t *count; const uint64_t *out_off; Elem *base;
    __device__ __forceinline__ bool next(uint32_t &x, uint64_t &row) const {
        x = atomicAdd(ticket, 1u);
        if (x >= n_rows) return false;
        row = rows[x];
        return true;
    }
    __device__ __forceinline__ Elem *out(uint32_t x, uint64_t) const { return base + out_off[x]; }
    __device__ __forceinline__ void done(uint32_t x, uint64_t, uint64_t n) const { if (count) count[x] = uint32_t(n); }
};
struct LongRowsInBins {                 // the engine: rows of the plan's xl list inside [row_lo, row_hi) with at least
    const uint32_t *xl_list;            // min_len partial products; the merged row goes to the start of the row's
    DevScalars *sc;                     // (untouched) bin and uniq[row] = nnz, which is where k_merge_chain expects a
    const uint64_t *row_bin;            // long row -- the same hand-over as k_merge_xl
    uint64_t bin_base; Elem *bins; uint32_t *uniq;
    uint64_t row_lo, row_hi, min_len;
    uint64_t huge_len;                  // rows of at least this many partial products are handed out first: the list is in
                                        // no particular order, and a hub row that starts last is the tail of the launch
    __device__ __forceinline__ bool next(uint32_t &x, uint64_t &row) const {
        const uint32_t n = sc->n_xl;
        while (true) {
            x = atomicAdd(&sc->xl_ticket, 1u);                 // tickets 0 .. n-1: the huge rows; n .. 2n-1: the others
            if (x >= 2 * uint64_t(n)) return false;
            const bool first_pass = x < n;
            row = xl_list[first_pass ? x : x - n];
            const uint64_t len = row_bin[row + 1] - row_bin[row];
            if (row >= row_lo && row < row_hi && len >= min_len && (len >= huge_len) == first_pass) return true;
        }
    }
    __device__ __forceinline__ Elem *out(uint32_t, uint64_t row) const { return bins + (row_bin[row] - bin_base); }
    __device__ __forceinline__ void done(uint32_t, uint64_t row, uint64_t n) const { uniq[row] = uint32_t(n); }
};

// One CTA per listed row, rows handed out by ticket.  bandptr: the band index of B for bands of BAND columns.
template <int THREADS, int BAND, int RUNS, bool VALUES, class Rows>
__device__ __forceinline__ void long_rows_sweep(const uint64_t *__restrict__ a_pos, const Elem *__restrict__ a_data,
                                                const Elem *__restrict__ b_data, const uint32_t *__restrict__ bandptr,
                                                const uint64_t cols, const Rows rows) {
    static_assert(BAND % 32 == 0 && BAND <= 65536 && RUNS >= 1, "the band is a bitmap of whole words");
    static_assert(THREADS <= 1024 && THREADS % 32 == 0, "whole warps");
    using L = LongRowSmem<BAND, RUNS, VALUES>;
    OSP_EXTERN_SMEM(smem);
    float *acc = reinterpret_cast<float *>(smem + L::acc);
    uint16_t *owner = reinterpret_cast<uint16_t *>(smem + L::owner);
    uint32_t *bits = reinterpret_cast<uint32_t *>(smem + L::bits);
    uint32_t *s_pre = reinterpret_cast<uint32_t *>(smem + L::pre);
    uint32_t *s_lo = reinterpret_cast<uint32_t *>(smem + L::lo);
    float *s_a = reinterpret_cast<float *>(smem + L::aval);
    uint32_t *warp_sums = reinterpret_cast<uint32_t *>(smem + L::sums);
    uint32_t *s_x = warp_sums + 34;                                        // [0] ticket, [1] 1 while there is a row
    uint64_t *s_row = reinterpret_cast<uint64_t *>(smem + L::bytes - 8);
    const uint32_t tid = threadIdx.x;
    constexpr uint32_t WORDS = BAND / 32;
    for (uint32_t w = tid; w < WORDS; w += THREADS) bits[w] = 0u;
    if (VALUES)
        for (uint32_t c = tid; c < BAND; c += THREADS) { acc[c] = -0.0f; owner[c] = 0xFFFF; }
    const uint32_t n_bands = uint32_t((cols + BAND - 1) / BAND);
    const uint64_t nb1 = uint64_t(n_bands) + 1;
    while (true) {
        __syncthreads();
        if (tid == 0) {
            uint32_t tx = 0; uint64_t trow = 0;
            s_x[1] = rows.next(tx, trow) ? 1u : 0u;
            s_x[0] = tx; *s_row = trow;
        }
        __syncthreads();
        if (!s_x[1]) break;
        const uint32_t x = s_x[0];
        const uint64_t row = *s_row;
        const uint64_t p0 = a_pos[row];
        const uint64_t R = a_pos[row + 1] - p0;
        Elem *out = VALUES ? rows.out(x, row) : nullptr;
        // Rows whose runs fit one group and one pass (the common case: <= min(RUNS, THREADS) non-zeros in the row of A)
        // keep their run in registers: A's element, the row of the band index, and the segment bounds of the current
        // band -- the bound of the band after is loaded one band ahead, so no load sits between two bands.
        const bool one_group = R <= uint64_t(RUNS < THREADS ? RUNS : THREADS);
        Elem my_ak; my_ak.idx = 0; my_ak.val = 0.f;
        const uint32_t *my_bp = bandptr;
        uint32_t cur_lo = 0, cur_hi = 0;
        if (one_group && tid < R) {
            my_ak = a_data[p0 + tid];
            my_bp = bandptr + uint64_t(my_ak.idx) * nb1;
            cur_lo = my_bp[0];
            cur_hi = n_bands ? my_bp[1] : cur_lo;
        }
        uint32_t my_count = 0;                      // VALUES=false: columns seen, summed over this thread's bitmap words
        uint64_t produced = 0;                      // VALUES=true: elements of the row already written
        for (uint32_t band = 0; band < n_bands; band++) {
            const uint64_t band_lo = uint64_t(band) * BAND;
            bool touched = false;                                          // uniform: some run has an element in this band
            uint32_t nxt_hi = 0;
            if (one_group && tid < R && band + 2 <= n_bands) nxt_hi = my_bp[band + 2];
            for (uint64_t g0 = 0; g0 < R; g0 += RUNS) {
                const uint32_t G = uint32_t(min(uint64_t(RUNS), R - g0));
                // ---- the segment of every run of the group inside this band ----
                uint32_t carry = 0;
                for (uint32_t r0 = 0; r0 < G; r0 += THREADS) {
                    const uint32_t r = r0 + tid;
                    uint32_t seg = 0;
                    if (r < G && one_group) {
                        seg = cur_hi - cur_lo;
                        s_lo[r] = cur_lo;
                        if (VALUES) s_a[r] = my_ak.val;
                    } else if (r < G) {
                        const Elem ak = a_data[p0 + g0 + r];
                        const uint32_t *bp = bandptr + uint64_t(ak.idx) * nb1 + band;
                        const uint32_t lo = bp[0];
                        seg = bp[1] - lo;
                        s_lo[r] = lo;
                        if (VALUES) s_a[r] = ak.val;
                    }
                    uint32_t total;
                    const uint32_t rank = block_exclusive_scan(seg, warp_sums, total);
                    if (r < G) s_pre[r] = carry + rank;
                    carry += total;
                    __syncthreads();                                       // warp_sums is reused by the next pass
                }
                if (tid == 0) s_pre[G] = carry;
                __syncthreads();
                const uint32_t T = carry;                                  // elements of the group inside the band, in (k, column) order
                touched |= T > 0;
                // ---- consume them THREADS at a time; the next chunk's elements are in flight during the arbitration ----
                // element f of the group: run a with s_pre[a] <= f < s_pre[a + 1], element f - s_pre[a] of its segment
                auto locate = [&](uint32_t f, Elem &e, uint32_t &a) {
                    a = 0;
                    uint32_t b = G;
                    while (b - a > 1) {
                        const uint32_t mid = (a + b) >> 1;
                        if (s_pre[mid] <= f) a = mid; else b = mid;
                    }
                    e = b_data[uint64_t(s_lo[a]) + (f - s_pre[a])];
                };
                Elem e_next; e_next.idx = 0; e_next.val = 0.f;
                uint32_t a_next = 0;
                if (tid < T) locate(tid, e_next, a_next);
                for (uint32_t c0 = 0; c0 < T; c0 += THREADS) {
                    const uint32_t f = c0 + tid;
                    bool pending = f < T;
                    const Elem e = e_next;
                    const uint32_t a = a_next;
                    if (f + THREADS < T) locate(f + THREADS, e_next, a_next);
                    const uint32_t col = pending ? uint32_t(e.idx - band_lo) : 0u;
                    float val = 0.f;
                    if (VALUES && pending) val = __fmul_rn(s_a[a], e.val);
                    if (!VALUES) {
                        if (pending) atomicOr(&bits[col >> 5], 1u << (col & 31));
                    } else {
                        // A chunk holds at least one product (c0 < T) and follows a barrier, so the first round needs no
                        // vote: three barriers for a chunk without a repeated column.
                        do {
                            // the lowest pending position of every column wins this round (racing minimum, re-checked)
                            bool want = pending;
                            do {
                                if (want && owner[col] > tid) owner[col] = uint16_t(tid);
                                __syncthreads();
                                want = pending && owner[col] > tid;
                            } while (__syncthreads_or(want));
                            if (pending && owner[col] == tid) {
                                acc[col] = __fadd_rn(acc[col], val);
                                atomicOr(&bits[col >> 5], 1u << (col & 31));
                                owner[col] = 0xFFFF;
                                pending = false;
                            }
                        } while (__syncthreads_or(pending));
                    }
                }
                __syncthreads();                                           // s_pre / s_lo are rewritten by the next group
            }
            cur_lo = cur_hi; cur_hi = nxt_hi;                              // (one_group) the next band's segment
            // ---- the band is complete: count or emit its columns, reset it ----
            if (!touched) continue;                                        // nothing was set: bitmap and accumulator are still clean
            if (!VALUES) {
                for (uint32_t w = tid; w < WORDS; w += THREADS) {
                    my_count += __popc(bits[w]);
                    bits[w] = 0u;
                }
            } else {
                for (uint32_t w0 = 0; w0 < WORDS; w0 += THREADS) {
                    const uint32_t w = w0 + tid;
                    uint32_t bm = w < WORDS ? bits[w] : 0u;
                    uint32_t total;
                    const uint32_t rank = block_exclusive_scan(uint32_t(__popc(bm)), warp_sums, total);
                    uint64_t o = produced + rank;
                    if (bm) bits[w] = 0u;
                    while (bm) {
                        const uint32_t bit = __ffs(bm) - 1;
                        bm &= bm - 1;
                        const uint32_t c = w * 32 + bit;
                        Elem e; e.idx = uint32_t(band_lo + c); e.val = acc[c];
                        out[o++] = e;
                        acc[c] = -0.0f;
                    }
                    produced += total;
                    __syncthreads();                                       // warp_sums is reused
                }
            }
            __syncthreads();
        }
        if (!VALUES) {
            uint32_t total;
            block_exclusive_scan(my_count, warp_sums, total);
            if (tid == 0) rows.done(x, row, total);
        } else if (tid == 0) {
            rows.done(x, row, produced);
        }
    }
}

template <int THREADS, int BAND, int RUNS, class Rows>
__global__ void __launch_bounds__(THREADS)
k_long_count(const uint64_t *__restrict__ a_pos, const Elem *__restrict__ a_data, const Elem *__restrict__ b_data,
             const uint32_t *__restrict__ bandptr, uint64_t cols, Rows rows) {
    long_rows_sweep<THREADS, BAND, RUNS, false>(a_pos, a_data, b_data, bandptr, cols, rows);
}

template <int THREADS, int BAND, int RUNS, class Rows>
__global__ void __launch_bounds__(THREADS)
k_long_fill(const uint64_t *__restrict__ a_pos, const Elem *__restrict__ a_data, const Elem *__restrict__ b_data,
            const uint32_t *__restrict__ bandptr, uint64_t cols, Rows rows) {
    long_rows_sweep<THREADS, BAND, RUNS, true>(a_pos, a_data, b_data, bandptr, cols, rows);
}

// The tasks (non-zeros of A) of the rows the sweep takes: one bit per task, read by k_multiply (TaskSrcSoASwept), which
// then emits nothing for them -- their bins stay untouched until k_long_fill writes the merged row there.
// One warp per listed row; `swept` is zeroed by the host.
__global__ void k_mark_swept(const uint64_t *__restrict__ a_pos, const uint32_t *__restrict__ xl_list, const DevScalars *sc,
                             const uint64_t *__restrict__ row_bin, uint64_t min_len, uint32_t *swept) {
    const uint32_t n = sc->n_xl;
    const uint64_t warp = (blockIdx.x * uint64_t(blockDim.x) + threadIdx.x) >> 5, nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    for (uint64_t x = warp; x < n; x += nwarps) {
        const uint64_t row = xl_list[x];
        if (row_bin[row + 1] - row_bin[row] < min_len) continue;
        for (uint64_t e = a_pos[row] + lane_id(); e < a_pos[row + 1]; e += 32) atomicOr(&swept[e >> 5], 1u << (e & 31));
    }
}

// OSP_FUSED_SHORT: the tasks whose partial products still go to the bins -- those of the medium rows (long_list) and of
// the xl rows the sweep does not take (fewer than sweep_min partial products; ~0 without the sweep).
__global__ void k_mark_binned(const uint64_t *__restrict__ a_pos, const uint32_t *__restrict__ xl_list,
                              const uint32_t *__restrict__ long_list, const DevScalars *sc, const uint64_t *__restrict__ row_bin,
                              uint64_t sweep_min, uint32_t *bits) {
    const uint32_t n_xl = sc->n_xl, n = n_xl + sc->n_long;
    const uint64_t warp = (blockIdx.x * uint64_t(blockDim.x) + threadIdx.x) >> 5, nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    for (uint64_t x = warp; x < n; x += nwarps) {
        const uint64_t row = x < n_xl ? xl_list[x] : long_list[x - n_xl];
        if (x < n_xl && row_bin[row + 1] - row_bin[row] >= sweep_min) continue;
        for (uint64_t e = a_pos[row] + lane_id(); e < a_pos[row + 1]; e += 32) atomicOr(&bits[e >> 5], 1u << (e & 31));
    }
}

}  // namespace osp
