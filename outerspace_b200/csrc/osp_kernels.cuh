// osp_kernels.cuh -- sm_100a kernels of the outer-product SpGEMM engine.
//
// Phase map (reference file:line it replaces, relative to simulator/):
//   k_scan<SymIn,...>       symbolic pass: flop count SimSpGEMM.cpp:884-891 and the implicit
//                           push_back sizing of SimOuterSPACE.cpp:87-92 (+ column histogram of A)
//   k_scan<U32In,..>/k_scatter_tasks
//                           CSR->CSC conversion of A: coo2csr<true> SimSpGEMM.cpp:111-117,128-141
//   k_plan                  partitions the output rows into merge tiles and queues long rows
//   k_multiply              TaskProvider::multiplyPhase SimOuterSPACE.cpp:74-97 with the intended
//                           semantics of cscMulcsr SimSpGEMM.cpp:265-281 (true column ids)
//   k_merge_tiles/long/xl   TaskProvider::mergePhase SimOuterSPACE.cpp:98-132 with the intended
//                           semantics of deduplicateCOO SimSpGEMM.cpp:519-535 (sum equal columns),
//                           writing CSRMatrix mergedResult (SimOuterSPACE.cpp:140) directly
#pragma once
#include "osp_device.cuh"

namespace osp {

// =====================================================================================
// Generic single-pass exclusive scan (decoupled look-back).
//   In : (idx, valid) -> uint64 contribution; called by ALL lanes of every warp (valid = idx < n),
//        so functors may use warp collectives.
//   Out: (idx, exclusive prefix, own contribution) for idx < n, and once (n, total, 0).
// =====================================================================================
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

template <class In, class Out>
__global__ void __launch_bounds__(SCAN_BLOCK)
k_scan(In in, Out out, uint64_t n, uint64_t *tile_state, unsigned int *ticket) {
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_warp[SCAN_BLOCK / 32];
    __shared__ uint64_t s_tile_excl;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const unsigned int lane = lane_id(), warp = threadIdx.x >> 5;
    const uint64_t wbase = uint64_t(tile) * SCAN_TILE + uint64_t(warp) * (32 * SCAN_ITEMS);

    uint64_t excl[SCAN_ITEMS], own[SCAN_ITEMS];
    uint64_t running = 0;
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; it++) {
        uint64_t idx = wbase + it * 32 + lane;
        uint64_t x = in(idx, idx < n);
        uint64_t incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t y = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += y;
        }
        own[it] = x;
        excl[it] = running + incl - x;
        running += __shfl_sync(FULL, incl, 31);
    }
    if (lane == 0) s_warp[warp] = running;
    __syncthreads();
    if (warp == 0) {
        uint64_t w = lane < SCAN_BLOCK / 32 ? s_warp[lane] : 0;
        uint64_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t y = __shfl_up_sync(FULL, wi, o);
            if (lane >= o) wi += y;
        }
        uint64_t aggregate = __shfl_sync(FULL, wi, 31);
        uint64_t tile_excl = lookback_exclusive(tile_state, tile, aggregate);
        if (lane < SCAN_BLOCK / 32) s_warp[lane] = wi - w;
        if (lane == 0) {
            s_tile_excl = tile_excl;
            uint64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
            if (tile == ntiles - 1) out(n, tile_excl + aggregate, 0);
        }
    }
    __syncthreads();
    const uint64_t offset = s_tile_excl + s_warp[warp];
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; it++) {
        uint64_t idx = wbase + it * 32 + lane;
        if (idx < n) out(idx, offset + excl[it], own[it]);
    }
}

// ---- functors ---------------------------------------------------------------------------
// Symbolic pass over the non-zeros of A in row order: entry p contributes nnz(B(k_p,:)) partial
// products; the exclusive prefix is the offset of its run inside the row bins (bins are laid out
// row after row, runs inside a row in ascending k: the order in which multiplyPhase appends to
// multResults[rowId], SimOuterSPACE.cpp:91).  Side effect: column histogram of A (first step of
// the CSR->CSC conversion), skipped when col_cnt is null.
struct SymIn {
    const Elem *a;
    const uint64_t *b_pos;
    uint64_t n_k;
    uint32_t *col_cnt;
    DevScalars *sc;
    __device__ uint64_t operator()(uint64_t p, bool valid) const {
        if (!valid) return 0;
        uint32_t k = a[p].idx;
        if (k >= n_k) { atomicMax(&sc->err, 4u); return 0; }   // OSP_ERR_INDEX
        if (col_cnt) atomicAdd(&col_cnt[k], 1u);
        return b_pos[k + 1] - b_pos[k];
    }
};
struct RunOffOut {
    uint64_t *run_off;   // [nnzA+1]
    DevScalars *sc;
    uint64_t n;
    __device__ void operator()(uint64_t p, uint64_t v, uint64_t) const {
        run_off[p] = v;
        if (p == n) sc->products = v;
    }
};
struct U32In {
    const uint32_t *x;
    __device__ uint64_t operator()(uint64_t i, bool valid) const { return valid ? x[i] : 0; }
};
struct U64Out {
    uint64_t *y;
    __device__ void operator()(uint64_t i, uint64_t v, uint64_t) const { y[i] = v; }
};
struct U32Out {           // column pointers of the task list (nnzA < 2^32 is checked on the host)
    uint32_t *y;
    __device__ void operator()(uint64_t i, uint64_t v, uint64_t) const { y[i] = uint32_t(v); }
};

// =====================================================================================
// Small utility kernels
// =====================================================================================
__global__ void k_max_idx(const Elem *d, uint64_t nnz, DevScalars *sc) {
    uint32_t m = 0;
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < nnz; i += uint64_t(gridDim.x) * blockDim.x)
        m = max(m, d[i].idx);
    m = __reduce_max_sync(FULL, m);
    if (lane_id() == 0 && m) atomicMax(&sc->max_idx, m);
}

// =====================================================================================
// Merge plan.  Tiles are runs of consecutive output rows; row i opens a tile when
//   i == 0, i % MT_RMAX == 0, its bin start crosses a multiple of MT_CAP, or it or its
//   predecessor is longer than MT_LONG (long rows are tiles of their own).
// So a tile of short rows holds < MT_CAP + MT_LONG partial products and <= MT_RMAX rows.
// Also: row_bin[] (bin start of every row), the queues of long rows, the upper bound of nnz(C)
// and the reference's row count rule numRows = maxRowId + 1 (SimOuterSPACE.cpp:49-53).
// =====================================================================================
constexpr uint32_t MT_CAP = 3072;      // partial products per tile (soft)
constexpr uint32_t MT_LONG = 512;      // longest row sorted in registers by one warp
constexpr uint32_t MT_STAGE = MT_CAP + MT_LONG;
constexpr uint32_t MT_RMAX = 1024;     // rows per tile
constexpr uint32_t MT_XL = 4096;       // longest row sorted in shared memory by one CTA

constexpr int PLAN_BLOCK = 256;
constexpr int PLAN_ITEMS = 4;
constexpr int PLAN_TILE = PLAN_BLOCK * PLAN_ITEMS;

struct RowBinFromRuns {        // row i starts where the run of its first non-zero starts
    const uint64_t *a_pos;
    uint64_t m_a;
    const uint64_t *run_off;
    uint64_t nnz_a;
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const { return run_off[i <= m_a ? a_pos[i] : nnz_a]; }
    // a row counts for numRows = maxRowId + 1 (SimOuterSPACE.cpp:49-53) when A holds a non-zero in it
    __device__ __forceinline__ bool nonempty(uint64_t i) const { return i < m_a && a_pos[i + 1] > a_pos[i]; }
};
struct RowBinDirect {
    const uint64_t *pos;
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const { return pos[i]; }
    __device__ __forceinline__ bool nonempty(uint64_t i) const { return pos[i + 1] > pos[i]; }
};

template <class RB>
__global__ void __launch_bounds__(PLAN_BLOCK)
k_plan(RB rb, uint64_t rows, uint64_t cols_hint, uint64_t *row_bin, uint32_t *tile_row, uint32_t *long_list,
       uint32_t *xl_list, uint64_t *tile_state, DevScalars *sc, int ticket_slot) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp[33];
    __shared__ uint64_t s_bound[PLAN_BLOCK / 32];
    __shared__ uint32_t s_last[PLAN_BLOCK / 32];
    __shared__ uint64_t s_excl;
    if (threadIdx.x == 0) s_tile = atomicAdd(&sc->scan_ticket[ticket_slot], 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const unsigned int lane = lane_id(), warp = threadIdx.x >> 5;
    const uint64_t i0 = uint64_t(tile) * PLAN_TILE + uint64_t(threadIdx.x) * PLAN_ITEMS;

    uint64_t s[PLAN_ITEMS + 1];
#pragma unroll
    for (int it = 0; it <= PLAN_ITEMS; it++) s[it] = rb(min(i0 + it, rows));
    uint64_t sp = i0 > 0 && i0 <= rows ? rb(i0 - 1) : 0;   // start of the previous row
    bool flag[PLAN_ITEMS];
    uint32_t nflags = 0, last = 0;
    uint64_t bound = 0;
#pragma unroll
    for (int it = 0; it < PLAN_ITEMS; it++) {
        const uint64_t i = i0 + it;
        flag[it] = false;
        if (i < rows) {
            const uint64_t len = s[it + 1] - s[it];
            const uint64_t plen = s[it] - sp;
            flag[it] = i == 0 || (i % MT_RMAX) == 0 || len > MT_LONG || plen > MT_LONG || (sp / MT_CAP) != (s[it] / MT_CAP);
            row_bin[i] = s[it];
            if (len > MT_XL) xl_list[atomicAdd(&sc->n_xl, 1u)] = uint32_t(i);
            else if (len > MT_LONG) long_list[atomicAdd(&sc->n_long, 1u)] = uint32_t(i);
            if (rb.nonempty(i)) last = uint32_t(i + 1);
            bound += cols_hint ? min(len, cols_hint) : len;
            nflags += flag[it];
        }
        sp = s[it];
    }
    // block reductions of bound / last
    last = __reduce_max_sync(FULL, last);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bound += __shfl_xor_sync(FULL, bound, o);
    if (lane == 0) { s_bound[warp] = bound; s_last[warp] = last; }
    uint32_t total;
    uint32_t rank = block_exclusive_scan(nflags, s_warp, total);
    if (warp == 0) {
        uint64_t b = lane < PLAN_BLOCK / 32 ? s_bound[lane] : 0;
        uint32_t l = lane < PLAN_BLOCK / 32 ? s_last[lane] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(FULL, b, o);
        l = __reduce_max_sync(FULL, l);
        if (lane == 0) {
            if (b) atomicAdd(&sc->cap_bound, (unsigned long long)b);
            if (l) atomicMax(&sc->last_nonempty, (unsigned long long)l);
        }
        uint64_t excl = lookback_exclusive(tile_state, tile, total);
        if (lane == 0) {
            s_excl = excl;
            const uint64_t ntiles = (rows + PLAN_TILE - 1) / PLAN_TILE;
            if (tile == ntiles - 1) {
                sc->n_tiles = uint32_t(excl + total);
                tile_row[excl + total] = uint32_t(rows);
                row_bin[rows] = rb(rows);
            }
        }
    }
    __syncthreads();
    uint64_t o = s_excl + rank;
#pragma unroll
    for (int it = 0; it < PLAN_ITEMS; it++)
        if (flag[it]) tile_row[o++] = uint32_t(i0 + it);
}

// =====================================================================================
// CSR -> CSC conversion of A.  Task list in k-slice (CSC) order: every non-zero of A lands in the
// slot range of its column k together with the bin offset of its run.  The order of the tasks
// inside a column is immaterial (each task owns a distinct run), so slots are claimed with an
// atomic; `col_cnt` (the histogram built by the symbolic pass) is counted down.
// =====================================================================================
__global__ void k_scatter_tasks(const Elem *a, const uint64_t *run_off, uint64_t e0, uint64_t e1,
                                const uint32_t *col_ptr, uint32_t *col_cnt, Task *tasks) {
    for (uint64_t p = e0 + blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; p < e1; p += uint64_t(gridDim.x) * blockDim.x) {
        Elem e = a[p];
        uint32_t c = atomicSub(&col_cnt[e.idx], 1u);
        Task t;
        t.k = e.idx; t.a = e.val; t.off = run_off[p];
        tasks[col_ptr[e.idx] + c - 1] = t;
    }
}

// Public stable conversion, step 1: bucket scatter of {source slice id, val} by minor index.
// One warp walks one source slice; the order inside a bucket is fixed up by the merge kernels.
__global__ void k_scatter_elems(const uint64_t *pos, const Elem *d, uint64_t n_major, uint64_t n_minor,
                                const uint64_t *pos_out, uint32_t *cursor, Elem *out, DevScalars *sc) {
    uint64_t warp = (blockIdx.x * uint64_t(blockDim.x) + threadIdx.x) >> 5;
    uint64_t nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    for (uint64_t r = warp; r < n_major; r += nwarps) {
        uint64_t b = pos[r], e = pos[r + 1];
        for (uint64_t p = b + lane_id(); p < e; p += 32) {
            Elem x = d[p];
            if (x.idx >= n_minor) { atomicMax(&sc->err, 4u); continue; }
            uint32_t slot = atomicAdd(&cursor[x.idx], 1u);
            Elem y; y.idx = uint32_t(r); y.val = x.val;
            out[pos_out[x.idx] + slot] = y;
        }
    }
}
__global__ void k_hist_elems(const Elem *d, uint64_t nnz, uint64_t n_minor, uint32_t *cnt, DevScalars *sc) {
    for (uint64_t p = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; p < nnz; p += uint64_t(gridDim.x) * blockDim.x) {
        uint32_t c = d[p].idx;
        if (c >= n_minor) { atomicMax(&sc->err, 4u); continue; }
        atomicAdd(&cnt[c], 1u);
    }
}

// =====================================================================================
// Multiply phase: every task streams row k of B, scales it by A(i,k) and writes the run of
// (col, a*b) partial products into row i's bin.  G lanes cooperate on one run (G | 32), so a
// warp works on 32/G runs at once; G is picked from the mean row length of B.
// Task sources: Task[] (k-slice order) or {Elem[], run_off[]} (row order of A).
// =====================================================================================
struct TaskSrcAoS {
    const Task *t;
    __device__ __forceinline__ void load(uint64_t i, uint32_t &k, float &a, uint64_t &off) const {
        const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(t + i));   // one 128-bit load
        k = raw.x; a = __uint_as_float(raw.y);
        off = (uint64_t(raw.w) << 32) | raw.z;
    }
};
struct TaskSrcSoA {
    const Elem *a_data;
    const uint64_t *run_off;
    __device__ __forceinline__ void load(uint64_t i, uint32_t &k, float &a, uint64_t &off) const {
        Elem e = a_data[i];
        k = e.idx; a = e.val; off = run_off[i];
    }
};

template <int G, class Src>
__global__ void __launch_bounds__(256)
k_multiply(Src src, uint64_t t0, uint64_t t1, const uint64_t *__restrict__ b_pos, const Elem *__restrict__ b_data,
           Elem *__restrict__ bins, uint64_t bin_base) {
    const unsigned int lane = lane_id();
    const uint64_t warp = (blockIdx.x * uint64_t(blockDim.x) + threadIdx.x) >> 5;
    const uint64_t nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    constexpr int RUNS = 32 / G;
    const unsigned int g = lane / G, tl = lane % G;
    for (uint64_t base = t0 + warp * 32; base < t1; base += nwarps * 32) {
        uint32_t k = 0; float a = 0.f; uint64_t off = 0, bs = 0; uint32_t bl = 0;
        if (base + lane < t1) {
            src.load(base + lane, k, a, off);
            bs = b_pos[k];
            bl = uint32_t(b_pos[k + 1] - bs);
        }
#pragma unroll 1
        for (int j0 = 0; j0 < 32; j0 += RUNS) {
            const int srcl = j0 + g;
            const uint64_t bs_j = __shfl_sync(FULL, bs, srcl);
            const uint32_t bl_j = __shfl_sync(FULL, bl, srcl);
            const float a_j = __shfl_sync(FULL, a, srcl);
            const uint64_t off_j = __shfl_sync(FULL, off, srcl) - bin_base;
            if (__ballot_sync(FULL, bl_j != 0) == 0) continue;
            for (uint32_t t = tl; t < bl_j; t += G) {
                Elem b = b_data[bs_j + t];
                Elem o; o.idx = b.idx; o.val = __fmul_rn(a_j, b.val);     // rounded on its own: no FMA
                bins[off_j + t] = o;
            }
        }
    }
}

// =====================================================================================
// Merge, long rows (MT_LONG < len <= MT_XL): one CTA sorts one row's partial products by
// (col, arrival position) with a bitonic network in shared memory, left-folds equal columns in
// arrival (= k) order with separately rounded adds, and writes the compacted row back to the
// start of its bin; uniq[row] = surviving entries.  k_merge_tiles copies it into C.
// Shared: uint64 keys[cap] | float vals[cap] | uint32 warp_sums[33]
// =====================================================================================
__device__ __forceinline__ void bitonic_sort_shared(uint64_t *keys, uint32_t N) {
    for (uint32_t k = 2; k <= N; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < (N >> 1); t += blockDim.x) {
                uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                uint32_t p = i | j;
                uint64_t x = keys[i], y = keys[p];
                bool up = (i & k) == 0;
                if ((x > y) == up) { keys[i] = y; keys[p] = x; }
            }
            __syncthreads();
        }
    }
}

// Sorted keys (col<<32 | pos) + vals[pos] -> folded, compacted row at out[0..uniq).  Returns uniq
// (valid in every thread).  When `acc` is non-null the fold of a column starts from acc[col] if its
// bit in `bits` is set (long-row accumulator) and the result goes to acc instead of `out`.
__device__ __forceinline__ uint32_t fold_sorted(const uint64_t *keys, const float *vals, uint32_t len, uint32_t N,
                                                Elem *out, uint32_t *warp_sums, float *acc, uint32_t *bits) {
    uint32_t produced = 0;
    for (uint32_t s0 = 0; s0 < N; s0 += blockDim.x) {
        uint32_t s = s0 + threadIdx.x;
        bool head = false;
        uint32_t col = 0;
        if (s < len) {
            col = uint32_t(keys[s] >> 32);
            head = (s == 0) || (uint32_t(keys[s - 1] >> 32) != col);
        }
        float sum = 0.f;
        if (head) {
            sum = vals[uint32_t(keys[s])];
            if (acc) {
                bool seen = (bits[col >> 5] >> (col & 31)) & 1u;
                if (seen) sum = __fadd_rn(acc[col], sum);
            }
            for (uint32_t u = s + 1; u < len && uint32_t(keys[u] >> 32) == col; u++)
                sum = __fadd_rn(sum, vals[uint32_t(keys[u])]);
        }
        uint32_t total;
        uint32_t rank = block_exclusive_scan(head ? 1u : 0u, warp_sums, total);
        if (head) {
            if (acc) {
                acc[col] = sum;
                atomicOr(&bits[col >> 5], 1u << (col & 31));
            } else {
                Elem o; o.idx = col; o.val = sum;
                out[produced + rank] = o;
            }
        }
        produced += total;
    }
    return produced;
}

__global__ void __launch_bounds__(256)
k_merge_long(const uint64_t *__restrict__ row_bin, uint64_t bin_base, Elem *bins, uint32_t *uniq,
             const uint32_t *long_list, const DevScalars *sc, uint64_t row_lo, uint64_t row_hi) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem);
    float *vals = reinterpret_cast<float *>(keys + MT_XL);
    uint32_t *warp_sums = reinterpret_cast<uint32_t *>(vals + MT_XL);
    const uint32_t n_long = sc->n_long;
    for (uint32_t x = blockIdx.x; x < n_long; x += gridDim.x) {
        const uint64_t row = long_list[x];
        if (row < row_lo || row >= row_hi) continue;
        const uint64_t start = row_bin[row] - bin_base;
        const uint32_t len = uint32_t(row_bin[row + 1] - row_bin[row]), N = pow2ceil(len);
        Elem *bin = bins + start;
        for (uint32_t p = threadIdx.x; p < N; p += blockDim.x) {
            if (p < len) {
                Elem e = bin[p];
                keys[p] = (uint64_t(e.idx) << 32) | p;
                vals[p] = e.val;
            } else {
                keys[p] = ~0ull;
            }
        }
        __syncthreads();
        bitonic_sort_shared(keys, N);
        uint32_t u = fold_sorted(keys, vals, len, N, bin, warp_sums, nullptr, nullptr);
        if (threadIdx.x == 0) uniq[row] = u;
        __syncthreads();
    }
}

// Rows longer than MT_XL: the row is consumed in chunks of MT_XL partial products.  Each chunk is
// sorted by (col, arrival position) and folded INTO a dense per-CTA accumulator acc[cols_b]
// (presence in bits[]), which keeps the left fold in arrival order across chunks.  The accumulator
// is then compacted in ascending column order over the start of the row's bin; bits[] is cleared.
__global__ void __launch_bounds__(256)
k_merge_xl(const uint64_t *__restrict__ row_bin, uint64_t bin_base, Elem *bins, uint32_t *uniq,
           const uint32_t *xl_list, const DevScalars *sc, float *acc_all, uint32_t *bits_all, uint64_t cols_b,
           uint64_t row_lo, uint64_t row_hi) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem);
    float *vals = reinterpret_cast<float *>(keys + MT_XL);
    uint32_t *warp_sums = reinterpret_cast<uint32_t *>(vals + MT_XL);
    const uint64_t words = (cols_b + 31) >> 5;
    float *acc = acc_all + uint64_t(blockIdx.x) * cols_b;
    uint32_t *bits = bits_all + uint64_t(blockIdx.x) * words;
    const uint32_t n_xl = sc->n_xl;
    for (uint32_t x = blockIdx.x; x < n_xl; x += gridDim.x) {
        const uint64_t row = xl_list[x];
        if (row < row_lo || row >= row_hi) continue;
        const uint64_t start = row_bin[row] - bin_base;
        const uint64_t len = row_bin[row + 1] - row_bin[row];
        Elem *bin = bins + start;
        for (uint64_t c0 = 0; c0 < len; c0 += MT_XL) {
            const uint32_t n = uint32_t(min(uint64_t(MT_XL), len - c0)), N = pow2ceil(n);
            for (uint32_t p = threadIdx.x; p < N; p += blockDim.x) {
                if (p < n) {
                    Elem e = bin[c0 + p];
                    keys[p] = (uint64_t(e.idx) << 32) | p;
                    vals[p] = e.val;
                } else {
                    keys[p] = ~0ull;
                }
            }
            __syncthreads();
            bitonic_sort_shared(keys, N);
            fold_sorted(keys, vals, n, N, nullptr, warp_sums, acc, bits);
            __syncthreads();
        }
        // ordered compaction of the accumulator over the (fully consumed) bin
        uint64_t produced = 0;
        for (uint64_t w0 = 0; w0 < words; w0 += blockDim.x) {
            uint64_t w = w0 + threadIdx.x;
            uint32_t b = w < words ? bits[w] : 0u;
            uint32_t total;
            uint32_t rank = block_exclusive_scan(__popc(b), warp_sums, total);
            uint64_t o = produced + rank;
            if (b) bits[w] = 0u;
            while (b) {
                uint32_t bit = __ffs(b) - 1;
                b &= b - 1;
                uint32_t col = uint32_t(w * 32 + bit);
                Elem e; e.idx = col; e.val = acc[col];
                bin[o++] = e;
            }
            produced += total;
        }
        if (threadIdx.x == 0) uniq[row] = uint32_t(produced);
        __syncthreads();
    }
}

// =====================================================================================
// Merge, the main kernel.  A persistent grid takes merge tiles in order (dynamic ticket).
// Per tile:
//   1. rows of 0/1 partial products are handled one per thread; every other row is taken by a
//      warp, which loads the row's bin (coalesced), sorts 32*E keys (col << pb | arrival position)
//      in registers with a bitonic network over shuffles, left-folds equal columns in arrival (= k)
//      order with separately rounded adds and leaves the compacted row in shared memory;
//   2. a block scan of the surviving counts + one decoupled look-back across tiles give the tile's
//      offset in C, so rows go straight from shared memory to their final place in C.data and
//      C.pos is written on the way -- no second pass over the merged rows, no scan kernel.
// A long row (a tile of its own) was merged in place by k_merge_long / k_merge_xl; its compacted
// prefix is copied from the bin.
// K = uint32_t when col << 9 fits 32 bits (cols <= 2^23), else uint64_t.
// =====================================================================================
template <int E, class K>
__device__ __forceinline__ void bitonic_regs(K (&x)[E], const unsigned int lane) {
    constexpr int N = 32 * E;
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= E) {                       // partner in lane ^ (j / E), same register
                const int lj = j / E;
                const bool lower = (lane & lj) == 0;
                const bool up = ((lane * E) & k) == 0;      // k >= 2j >= 2E here: direction is per lane
                const bool keep_min = lower == up;
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const K y = __shfl_xor_sync(FULL, x[e], lj);
                    x[e] = keep_min ? min(x[e], y) : max(x[e], y);
                }
            } else {                            // both elements in this lane
#pragma unroll
                for (int e = 0; e < E; e++) {
                    if ((e & j) == 0) {
                        const K lo = min(x[e], x[e | j]), hi = max(x[e], x[e | j]);
                        const bool up = k < E ? ((e & k) == 0) : (((lane * E) & k) == 0);   // i = lane*E + e
                        x[e] = up ? lo : hi;
                        x[e | j] = up ? hi : lo;
                    }
                }
            }
        }
    }
}

template <int V> struct ILog2 { static constexpr int value = 1 + ILog2<V / 2>::value; };
template <> struct ILog2<1> { static constexpr int value = 0; };

// One warp merges one row of 2 <= len <= 32*E partial products.  `src` = the row's bin in global
// memory, `region` = the row's len-element slot of the tile's staging buffer.  Returns the number of
// surviving entries, left at region[0 .. uniq).
template <int E, class K>
__device__ __forceinline__ uint32_t merge_row_regs(const Elem *__restrict__ src, Elem *region, const uint32_t len,
                                                   const unsigned int lane) {
    constexpr int N = 32 * E;
    constexpr int PB = ILog2<N>::value;
    K key[E];
    float *fstage = reinterpret_cast<float *>(region);
#pragma unroll
    for (int e = 0; e < E; e++) {
        const uint32_t p = e * 32 + lane;
        key[e] = ~K(0);
        if (p < len) {
            const Elem el = src[p];
            key[e] = (K(el.idx) << PB) | K(p);
            fstage[p] = el.val;
        }
    }
    __syncwarp();
    bitonic_regs<E, K>(key, lane);
    // lane now holds sorted positions lane*E .. lane*E+E-1
    uint32_t col[E];
    float v[E];
#pragma unroll
    for (int e = 0; e < E; e++) {
        const uint32_t s = lane * E + e;
        col[e] = uint32_t(key[e] >> PB);
        v[e] = s < len ? fstage[uint32_t(key[e]) & (N - 1)] : 0.f;
    }
    __syncwarp();
    uint32_t *scol = reinterpret_cast<uint32_t *>(region);
    float *sval = fstage + len;
#pragma unroll
    for (int e = 0; e < E; e++) {
        const uint32_t s = lane * E + e;
        if (s < len) { scol[s] = col[e]; sval[s] = v[e]; }
    }
    __syncwarp();
    // heads and their left folds
    uint32_t prev = __shfl_up_sync(FULL, col[E - 1], 1);
    const uint32_t next_first = __shfl_down_sync(FULL, col[0], 1);   // padding (all ones >> PB) past the end
    bool head[E];
    uint32_t nheads = 0;
#pragma unroll
    for (int e = 0; e < E; e++) {
        const uint32_t s = lane * E + e;
        head[e] = s < len && (s == 0 || prev != col[e]);
        prev = col[e];
        if (head[e]) {
            nheads++;
            const uint32_t nxt = e + 1 < E ? col[e + 1 < E ? e + 1 : 0] : next_first;
            const bool more = s + 1 < len && nxt == col[e];
            if (more) {
                float sum = v[e];
                for (uint32_t u = s + 1; u < len && scol[u] == col[e]; u++) sum = __fadd_rn(sum, sval[u]);
                v[e] = sum;
            }
        }
    }
    const uint32_t incl = warp_inclusive_scan(nheads);
    const uint32_t total = __shfl_sync(FULL, incl, 31);
    uint32_t r = incl - nheads;
    __syncwarp();
#pragma unroll
    for (int e = 0; e < E; e++) {
        if (head[e]) {
            Elem o; o.idx = col[e]; o.val = v[e];
            region[r++] = o;
        }
    }
    return total;
}

constexpr int MT_THREADS = 256;

template <class K>
__global__ void __launch_bounds__(MT_THREADS, 3)
k_merge_tiles(const uint64_t *__restrict__ row_bin, const uint64_t bin_base, const Elem *__restrict__ bins,
              const uint32_t *__restrict__ tile_row, const uint32_t t_first, const uint32_t t0, const uint32_t t1,
              const uint32_t n_tiles, const uint32_t *__restrict__ uniq_long, uint64_t *tile_state,
              uint64_t *__restrict__ c_pos, Elem *__restrict__ c_data, const uint64_t c_cap, const uint64_t rows,
              DevScalars *sc) {
    __shared__ __align__(16) Elem stage[MT_STAGE];
    __shared__ uint32_t s_start[MT_RMAX + 1];
    __shared__ uint32_t s_off[MT_RMAX + 1];
    __shared__ uint16_t s_work[MT_RMAX];
    __shared__ uint32_t s_warp[33];
    __shared__ uint32_t s_tile, s_nwork, s_next;
    __shared__ uint64_t s_base;
    const unsigned int lane = lane_id(), warp = threadIdx.x >> 5;

    while (true) {
        __syncthreads();                        // previous tile fully retired (stage, s_* reusable)
        if (threadIdx.x == 0) {
            s_tile = t0 + atomicAdd(&sc->tile_ticket, 1u);
            s_nwork = 0;
            s_next = 0;
        }
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= t1) break;
        const uint64_t r0 = tile_row[tile], r1 = tile_row[tile + 1];
        const uint32_t R = uint32_t(r1 - r0);
        const uint64_t bin0 = row_bin[r0];
        const uint64_t tile_len = row_bin[r1] - bin0;
        const bool single_long = R == 1 && tile_len > MT_LONG;
        uint32_t total = 0;
        if (single_long) {
            total = uniq_long[r0];
            if (threadIdx.x == 0) { s_off[0] = 0; s_off[1] = total; }
        } else {
            for (uint32_t j = threadIdx.x; j <= R; j += MT_THREADS) s_start[j] = uint32_t(row_bin[r0 + j] - bin0);
            __syncthreads();
            // rows with 0 / 1 partial products: one thread each; the others are queued for the warps
            uint32_t cnt[MT_RMAX / MT_THREADS];
#pragma unroll
            for (int q = 0; q < int(MT_RMAX / MT_THREADS); q++) {
                const uint32_t j = threadIdx.x * (MT_RMAX / MT_THREADS) + q;
                cnt[q] = 0;
                if (j < R) {
                    const uint32_t st = s_start[j], len = s_start[j + 1] - st;
                    if (len == 1) {
                        stage[st] = bins[bin0 - bin_base + st];
                        cnt[q] = 1;
                    } else if (len > 1) {
                        s_work[atomicAdd(&s_nwork, 1u)] = uint16_t(j);
                    }
                }
            }
            __syncthreads();
            const uint32_t nwork = s_nwork;
            while (true) {
                uint32_t w = 0;
                if (lane == 0) w = atomicAdd(&s_next, 1u);
                w = __shfl_sync(FULL, w, 0);
                if (w >= nwork) break;
                const uint32_t j = s_work[w];
                const uint32_t st = s_start[j], len = s_start[j + 1] - st;
                const Elem *src = bins + (bin0 - bin_base + st);
                Elem *region = stage + st;
                uint32_t u;
                if (len <= 32) u = merge_row_regs<1, K>(src, region, len, lane);
                else if (len <= 64) u = merge_row_regs<2, K>(src, region, len, lane);
                else if (len <= 128) u = merge_row_regs<4, K>(src, region, len, lane);
                else if (len <= 256) u = merge_row_regs<8, K>(src, region, len, lane);
                else u = merge_row_regs<16, K>(src, region, len, lane);
                if (lane == 0) s_off[j] = u;          // surviving count, turned into an offset below
            }
            __syncthreads();
            // exclusive scan of the surviving counts over the tile's rows (4 consecutive rows per thread)
            uint32_t mine = 0;
#pragma unroll
            for (int q = 0; q < int(MT_RMAX / MT_THREADS); q++) {
                const uint32_t j = threadIdx.x * (MT_RMAX / MT_THREADS) + q;
                if (j < R) {
                    const uint32_t len = s_start[j + 1] - s_start[j];
                    if (len > 1) cnt[q] = s_off[j];
                    mine += cnt[q];
                }
            }
            uint32_t ex = block_exclusive_scan(mine, s_warp, total);
#pragma unroll
            for (int q = 0; q < int(MT_RMAX / MT_THREADS); q++) {
                const uint32_t j = threadIdx.x * (MT_RMAX / MT_THREADS) + q;
                if (j < R) { s_off[j] = ex; ex += cnt[q]; }
            }
            if (threadIdx.x == 0) s_off[R] = total;
        }
        if (warp == 0) {
            const uint64_t base = lookback_exclusive(tile_state, tile, total, t_first, 0);
            if (lane == 0) {
                s_base = base;
                if (base + total > c_cap) atomicMax(&sc->err, 3u);       // cannot happen: c_cap is an upper bound
                if (tile == n_tiles - 1) {
                    c_pos[rows] = base + total;
                    sc->nnz_c = base + total;
                }
            }
        }
        __syncthreads();
        const uint64_t base = s_base;
        if (base + total > c_cap) continue;
        for (uint32_t j = threadIdx.x; j < R; j += MT_THREADS) c_pos[r0 + j] = base + s_off[j];
        Elem *dst = c_data + base;
        if (single_long) {
            const Elem *src = bins + (bin0 - bin_base);
            for (uint32_t i = threadIdx.x; i < total; i += MT_THREADS) dst[i] = src[i];
        } else if (total >= 16u * R) {
            for (uint32_t j = warp; j < R; j += MT_THREADS / 32) {
                const uint32_t o = s_off[j], n = s_off[j + 1] - o, st = s_start[j];
                for (uint32_t i = lane; i < n; i += 32) dst[o + i] = stage[st + i];
            }
        } else {
            for (uint32_t i = threadIdx.x; i < total; i += MT_THREADS) {
                uint32_t lo = 0, hi = R;                     // last j with s_off[j] <= i
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (s_off[mid] <= i) lo = mid; else hi = mid;
                }
                dst[i] = stage[s_start[lo] + (i - s_off[lo])];
            }
        }
    }
}

// Duplicate check of the stable conversion: a bucket that shrank while folding held a duplicate.
__global__ void k_check_same(const uint64_t *a, const uint64_t *b, uint64_t n, DevScalars *sc) {
    uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (i < n && a[i] != b[i]) atomicMax(&sc->err, 233u);
}

// =====================================================================================
// k-sharded multi-GPU path (DESIGN.md "Multi-GPU"): each rank multiplies its k-range and the
// partial products travel to the owner of their output row (contiguous row blocks), where the
// segments of the G sources are regrouped row by row in ascending source (= ascending k) order.
// =====================================================================================
// Bin start of every row of the shard's product and the per-row partial-product counts the owners need.
__global__ void k_shard_rows(const uint64_t *a_pos, uint64_t m_a, const uint64_t *run_off, uint64_t nnz_a, uint64_t rows,
                             uint64_t *row_bin, uint32_t *lens, DevScalars *sc) {
    uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (i > rows) return;
    const uint64_t s = run_off[i <= m_a ? a_pos[i] : nnz_a];
    row_bin[i] = s;
    if (i < rows) {
        const uint64_t e = run_off[i + 1 <= m_a ? a_pos[i + 1] : nnz_a];
        if (e - s >= (1ull << 32)) atomicMax(&sc->err, 6u);      // OSP_ERR_UNSUPPORTED
        lens[i] = uint32_t(e - s);
    }
}
__global__ void k_pick_u64(const uint64_t *src, const uint64_t *index, uint32_t n, uint64_t *dst) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[index[i]];
}
// lens[s * RL + i] read in (i, s) order
struct TransposedIn {
    const uint32_t *lens;
    uint64_t RL;
    uint64_t G;
    __device__ uint64_t operator()(uint64_t j, bool valid) const { return valid ? lens[(j % G) * RL + j / G] : 0; }
};
struct RowBinStrided {         // row i of the owner starts at dst_off[i * G]
    const uint64_t *off;
    uint64_t G;
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const { return off[i * G]; }
    __device__ __forceinline__ bool nonempty(uint64_t i) const { return off[(i + 1) * G] > off[i * G]; }
};
// One warp per owned row: copies the row's segment of every source (received source-major) into the
// row-major bins, sources in ascending order.
__global__ void k_regroup(const Elem *__restrict__ recv, const uint64_t *__restrict__ src_off, const uint64_t *__restrict__ dst_off,
                          const uint32_t *__restrict__ lens, uint64_t RL, uint32_t G, Elem *__restrict__ bins) {
    uint64_t warp = (blockIdx.x * uint64_t(blockDim.x) + threadIdx.x) >> 5;
    uint64_t nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    for (uint64_t i = warp; i < RL; i += nwarps) {
        for (uint32_t s = 0; s < G; s++) {
            const uint32_t n = lens[s * RL + i];
            const Elem *src = recv + src_off[s * RL + i];
            Elem *dst = bins + dst_off[i * G + s];
            for (uint32_t t = lane_id(); t < n; t += 32) dst[t] = src[t];
        }
    }
}

}  // namespace osp
