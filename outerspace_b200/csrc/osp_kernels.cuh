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
//   In : load(idx, valid) -> key, then value(key, valid) -> uint64 contribution (two stages so that the
//        first-level loads of all the items of a thread are in flight together).
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
    // two-stage input: all the first-level loads of a thread are issued before anything depends on them
    uint64_t key[SCAN_ITEMS];
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; it++) {
        const uint64_t idx = wbase + it * 32 + lane;
        key[it] = in.load(idx, idx < n);
    }
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; it++) {
        const uint64_t idx = wbase + it * 32 + lane;
        own[it] = in.value(key[it], idx < n);
    }
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; it++) {
        uint64_t x = own[it];
        uint64_t incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t y = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += y;
        }
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
    __device__ uint64_t load(uint64_t p, bool valid) const { return valid ? a[p].idx : ~0ull; }
    __device__ uint64_t value(uint64_t k, bool valid) const {
        if (!valid) return 0;
        if (k >= n_k) { atomicMax(&sc->err, 4u); return 0; }   // OSP_ERR_INDEX
        const uint64_t len = b_pos[k + 1] - b_pos[k];
        if (col_cnt) atomicAdd(&col_cnt[k], 1u);
        if (len >> TASK_LEN_BITS) { atomicMax(&sc->err, 6u); return 0; }   // OSP_ERR_UNSUPPORTED
        return len;
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
    __device__ uint64_t load(uint64_t i, bool valid) const { return valid ? x[i] : 0; }
    __device__ uint64_t value(uint64_t v, bool) const { return v; }
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
constexpr uint32_t MT_CAP = 512;       // partial products per tile (soft, = 2^9): one warp merges a tile
constexpr uint32_t MT_LONG = 512;      // longest row sorted in registers by one warp
constexpr uint32_t MT_STAGE = MT_CAP + MT_LONG;
constexpr uint32_t MT_RMAX = 32;       // rows per tile (one lane per row)
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
    // partial products per tile: MT_CAP, less when the whole product is small, so that there are enough
    // tiles (one warp each) to fill the machine: 2^cap_shift ~ P / 4096 within [32, MT_CAP]
    int cap_shift = 5;
    while (cap_shift < 9 && (sc->products >> (cap_shift + 1)) >= 4096) cap_shift++;

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
            flag[it] = i == 0 || (i % MT_RMAX) == 0 || len > MT_LONG || plen > MT_LONG || (sp >> cap_shift) != (s[it] >> cap_shift);
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
__global__ void k_scatter_tasks(const Elem *a, const uint64_t *run_off, const uint64_t *b_pos, uint64_t e0, uint64_t e1,
                                const uint32_t *col_ptr, uint32_t *col_cnt, Task *tasks, DevScalars *sc) {
    for (uint64_t p = e0 + blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; p < e1; p += uint64_t(gridDim.x) * blockDim.x) {
        const Elem e = a[p];
        const uint64_t off = run_off[p], len = run_off[p + 1] - off;
        const uint32_t c = atomicSub(&col_cnt[e.idx], 1u);
        if (len >> TASK_LEN_BITS) { atomicMax(&sc->err, 6u); continue; }     // OSP_ERR_UNSUPPORTED
        Task t;
        t.bs = uint32_t(b_pos[e.idx]); t.a = e.val; t.offlen = (off << TASK_LEN_BITS) | len;
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
// Multiply phase (warp-flat).  A warp takes 32 tasks, scans their run lengths and then walks the
// concatenation of the 32 runs 32 partial products at a time: every lane finds the task of its
// element with a 5-step search over the scanned lengths (shuffles), reads the element of B(k,:),
// multiplies by A(i,k) and writes (col, a*b) into row i's bin.  Every lane is busy whatever the run
// lengths are, loads and stores are contiguous inside a run, and there is no per-run loop.
// Task sources: packed Task[] (k-slice order: the CSC of A) or {Elem[], run_off[]} (row order of A).
// =====================================================================================
struct TaskSrcAoS {
    const Task *t;
    __device__ __forceinline__ void load(uint64_t i, uint32_t &bs, float &a, uint64_t &off, uint32_t &len) const {
        const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(t + i));   // one 128-bit load
        bs = raw.x; a = __uint_as_float(raw.y);
        const uint64_t offlen = (uint64_t(raw.w) << 32) | raw.z;
        off = offlen >> TASK_LEN_BITS;
        len = uint32_t(offlen) & ((1u << TASK_LEN_BITS) - 1);
    }
};
struct TaskSrcSoA {
    const Elem *a_data;
    const uint64_t *run_off;
    const uint64_t *b_pos;
    __device__ __forceinline__ void load(uint64_t i, uint32_t &bs, float &a, uint64_t &off, uint32_t &len) const {
        const Elem e = a_data[i];
        a = e.val; off = run_off[i];
        len = uint32_t(run_off[i + 1] - off);
        bs = uint32_t(b_pos[e.idx]);
    }
};

template <class Src>
__global__ void __launch_bounds__(256)
k_multiply(Src src, uint64_t t0, uint64_t t1, const Elem *__restrict__ b_data, Elem *__restrict__ bins, uint64_t bin_base) {
    const unsigned int lane = lane_id();
    const uint64_t warp = (blockIdx.x * uint64_t(blockDim.x) + threadIdx.x) >> 5;
    const uint64_t nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    for (uint64_t base = t0 + warp * 32; base < t1; base += nwarps * 32) {
        uint32_t bs = 0, len = 0; float a = 0.f; uint64_t off = 0;
        if (base + lane < t1) src.load(base + lane, bs, a, off, len);
        const uint32_t incl = warp_inclusive_scan(len);
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        const uint32_t excl = incl - len;
        const uint32_t dbs = bs - excl;                       // B index of element e of this task: dbs + e
        const uint64_t doff = off - bin_base - excl;          // bin index of element e of this task: doff + e
        for (uint32_t e0 = 0; e0 < total; e0 += 64) {          // two independent 32-element chunks per turn
            const uint32_t e[2] = {e0 + lane, e0 + 32 + lane};
            uint32_t t[2] = {0, 0};                            // number of tasks that end at or before e
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const uint32_t v = __shfl_sync(FULL, incl, t[u] + step - 1);
                    if (v <= e[u]) t[u] += step;
                }
            }
            float a_t[2]; uint32_t dbs_t[2]; uint64_t doff_t[2]; Elem b[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                a_t[u] = __shfl_sync(FULL, a, t[u] & 31);
                dbs_t[u] = __shfl_sync(FULL, dbs, t[u] & 31);
                doff_t[u] = __shfl_sync(FULL, doff, t[u] & 31);
            }
#pragma unroll
            for (int u = 0; u < 2; u++)
                if (e[u] < total) b[u] = b_data[dbs_t[u] + e[u]];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                if (e[u] < total) {
                    Elem o; o.idx = b[u].idx; o.val = __fmul_rn(a_t[u], b[u].val);     // rounded on its own: no FMA
                    bins[doff_t[u] + e[u]] = o;
                }
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
// Merge, long rows over a small column range (cols <= DENSE_MAX_COLS): one CTA per row folds the
// partial products into a dense accumulator in shared memory -- acc[col], seen[col] -- and then
// emits the seen columns in ascending order.  The row is consumed T partial products at a time in
// arrival order; products of one chunk that hit the same column are serialised by an arbitration
// on owner[col] (the lowest position goes first), so every column is still summed in ascending
// arrival (= k) order with separately rounded adds.  No sort, no global scratch.
// Shared: float acc[cols] | uint16 owner[cols] | uint8 seen[cols] | uint32 warp_sums[33]
// =====================================================================================
constexpr uint32_t DENSE_MAX_COLS = 16384;
constexpr int DENSE_THREADS = 512;
__host__ __device__ inline size_t dense_smem(uint64_t cols) { return size_t((cols + 15) & ~15ull) * 7 + 34 * 4; }

__global__ void __launch_bounds__(DENSE_THREADS)
k_merge_dense(const uint64_t *__restrict__ row_bin, uint64_t bin_base, Elem *bins, uint32_t *uniq,
              const uint32_t *long_list, const uint32_t *xl_list, const DevScalars *sc, uint32_t cols,
              uint64_t row_lo, uint64_t row_hi) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t cpad = (cols + 15) & ~15u;
    float *acc = reinterpret_cast<float *>(smem);
    uint16_t *owner = reinterpret_cast<uint16_t *>(smem + size_t(cpad) * 4);
    unsigned char *seen = smem + size_t(cpad) * 6;
    uint32_t *warp_sums = reinterpret_cast<uint32_t *>(smem + size_t(cpad) * 7);
    const uint32_t tid = threadIdx.x;
    for (uint32_t c = tid; c < cpad; c += DENSE_THREADS) { owner[c] = 0xFFFF; seen[c] = 0; }
    __syncthreads();
    const uint32_t n_long = sc->n_long, n_all = n_long + sc->n_xl;
    for (uint32_t x = blockIdx.x; x < n_all; x += gridDim.x) {
        const uint64_t row = x < n_long ? long_list[x] : xl_list[x - n_long];
        if (row < row_lo || row >= row_hi) continue;
        const uint64_t len = row_bin[row + 1] - row_bin[row];
        Elem *bin = bins + (row_bin[row] - bin_base);
        for (uint64_t c0 = 0; c0 < len; c0 += DENSE_THREADS) {
            const uint64_t p = c0 + tid;
            bool pending = p < len;
            Elem e; e.idx = 0; e.val = 0.f;
            if (pending) e = bin[p];
            while (__syncthreads_or(pending)) {
                // the lowest pending position of every column wins this round (racing minimum, re-checked)
                bool want = pending;
                do {
                    if (want && owner[e.idx] > tid) owner[e.idx] = uint16_t(tid);
                    __syncthreads();
                    want = pending && owner[e.idx] > tid;
                } while (__syncthreads_or(want));
                if (pending && owner[e.idx] == tid) {
                    acc[e.idx] = seen[e.idx] ? __fadd_rn(acc[e.idx], e.val) : e.val;
                    seen[e.idx] = 1;
                    owner[e.idx] = 0xFFFF;
                    pending = false;
                }
            }
        }
        __syncthreads();
        // emit the seen columns in ascending order over the (fully consumed) bin
        const uint32_t per = (cols + DENSE_THREADS - 1) / DENSE_THREADS;
        const uint32_t cb = min(tid * per, cols), ce = min(cb + per, cols);
        uint32_t cnt = 0;
        for (uint32_t c = cb; c < ce; c++) cnt += seen[c];
        uint32_t total;
        uint32_t o = block_exclusive_scan(cnt, warp_sums, total);
        for (uint32_t c = cb; c < ce; c++) {
            if (seen[c]) {
                Elem r; r.idx = c; r.val = acc[c];
                bin[o++] = r;
                seen[c] = 0;
            }
        }
        if (tid == 0) uniq[row] = total;
        __syncthreads();
    }
}

// =====================================================================================
// Merge of one short row by one warp, building blocks:
//   bitonic_regs      sorts 32*E keys (col << pb | arrival position) held E per lane in registers
//                     (compare-exchange over shuffles);
//   merge_row_regs    loads a row's bin (coalesced), sorts, left-folds equal columns in arrival (= k)
//                     order with separately rounded adds, writes the compacted row;
//   merge_row_bitmap  the same result without a sort, for small column ranges.
// =====================================================================================
template <int E, class K>
__device__ __forceinline__ void bitonic_regs(K (&x)[E], const unsigned int lane) {
    constexpr int N = 32 * E;
    if constexpr (E <= 4) {
        // short rows: the whole network unrolled (a loop would cost as much as its body here)
#pragma unroll
        for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                if (j >= E) {                       // partner in lane ^ (j / E), same register
                    const int lj = j / E;
                    const bool keep_min = ((lane & lj) == 0) == (((lane * E) & k) == 0);
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
    } else {
        // levels k = 2 .. E: both partners in this lane (compile-time register pairs and directions)
#pragma unroll
        for (int k = 2; k <= E; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
                for (int e = 0; e < E; e++) {
                    if ((e & j) == 0) {
                        const K lo = min(x[e], x[e | j]), hi = max(x[e], x[e | j]);
                        const bool up = k < E ? ((e & k) == 0) : ((lane & 1) == 0);   // i = lane*E + e
                        x[e] = up ? lo : hi;
                        x[e | j] = up ? hi : lo;
                    }
                }
            }
        }
        // levels k = 2E .. N: a runtime loop (keeps the code small enough for the instruction cache):
        // partners in lane ^ (j / E) while j >= E, then the in-lane tail j = E/2 .. 1
#pragma unroll 1
        for (int k = 2 * E; k <= N; k <<= 1) {
            const bool up = ((lane * E) & k) == 0;
#pragma unroll 1
            for (int lj = k / (2 * E); lj > 0; lj >>= 1) {
                const bool keep_min = ((lane & lj) == 0) == up;
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const K y = __shfl_xor_sync(FULL, x[e], lj);
                    x[e] = keep_min ? min(x[e], y) : max(x[e], y);
                }
            }
#pragma unroll
            for (int j = E >> 1; j > 0; j >>= 1) {
#pragma unroll
                for (int e = 0; e < E; e++) {
                    if ((e & j) == 0) {
                        const K lo = min(x[e], x[e | j]), hi = max(x[e], x[e | j]);
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
// memory, `region` = a len-element scratch in shared memory, `out` = where the compacted row goes (may
// be `src`: every input is in registers before the first output is written).  Returns the number of
// surviving entries.
template <int E, class K>
__device__ __forceinline__ uint32_t merge_row_regs(const Elem *src, Elem *region, Elem *out, const uint32_t len,
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
    // heads (first arrival of a column); most rows hold no duplicate column at all: then the sorted
    // row IS the result
    uint32_t prev = __shfl_up_sync(FULL, col[E - 1], 1);
    const uint32_t next_first = __shfl_down_sync(FULL, col[0], 1);   // padding (all ones >> PB) past the end
    {
        bool dup = false;
        uint32_t pc = prev;
#pragma unroll
        for (int e = 0; e < E; e++) {
            const uint32_t s = lane * E + e;
            dup |= s < len && s > 0 && pc == col[e];
            pc = col[e];
        }
        if (!__any_sync(FULL, dup)) {
            __syncwarp();
#pragma unroll
            for (int e = 0; e < E; e++) {
                const uint32_t s = lane * E + e;
                if (s < len) { Elem o; o.idx = col[e]; o.val = v[e]; out[s] = o; }
            }
            return len;
        }
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
    // left folds of the heads
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
            out[r++] = o;
        }
    }
    return total;
}

// ---- bitmap-rank merge of one row (column range <= 32 * BM_WORDS) -----------------------------------
// No sort: every partial product sets the bit of its column in a per-warp bitmap, a prefix popcount over
// the bitmap words turns a column into its rank among the row's distinct columns (= its place in the
// sorted, folded row), the earliest arrival of a column opens that place and later arrivals are added in
// arrival (= k) order.  Shared-memory read-modify-writes are plain loads/stores re-checked after a
// __syncwarp and retried by the lanes that lost a race (shared atomics cost ~2 cycles per lane; races
// here are rare).  Runtime loops over the 32-element slots of the row keep the code small: the unrolled
// per-E variants of an earlier version thrashed the instruction cache (profiles/README.md).
constexpr uint32_t BM_WORDS = 512;                 // bitmap words per warp: columns < 16384
// per-warp scratch: bitmap | word prefixes (u16) | first arrival per rank (u16) | rank per element (u16)
constexpr uint32_t BM_SCRATCH = BM_WORDS * 4 + BM_WORDS * 2 + MT_LONG * 2 + MT_LONG * 2;

__device__ __noinline__ uint32_t merge_row_bitmap(Elem *bin, Elem *region, const uint32_t len, const uint32_t words,
                                                  unsigned char *scratch, const unsigned int lane) {
    uint32_t *bm = reinterpret_cast<uint32_t *>(scratch);
    uint16_t *pre = reinterpret_cast<uint16_t *>(scratch + BM_WORDS * 4);
    uint16_t *first = reinterpret_cast<uint16_t *>(scratch + BM_WORDS * 4 + BM_WORDS * 2);
    uint16_t *rnk = first + MT_LONG;
    const uint32_t wpl = (((words + 31) >> 5) + 3) & ~3u;        // bitmap words per lane, multiple of 4, <= 16
    // lane owns the words {4*lane + 32*q .. +3}: conflict-free 128-bit accesses
    for (uint32_t q = 0; q < wpl; q += 4) *reinterpret_cast<uint4 *>(bm + 4 * lane + 32 * q) = make_uint4(0, 0, 0, 0);
    for (uint32_t q = lane; q < (len + 1) / 2; q += 32) reinterpret_cast<uint32_t *>(first)[q] = 0xFFFFFFFFu;
    for (uint32_t p = lane; p < len; p += 32) region[p] = bin[p];          // stage the row (the bin is overwritten below)
    __syncwarp();
    // set the bits
    for (uint32_t p0 = 0; p0 < len; p0 += 32) {
        const uint32_t p = p0 + lane;
        bool pend = p < len;
        const uint32_t col = pend ? region[p].idx : 0;
        const uint32_t w = col >> 5, bit = 1u << (col & 31);
        do {
            if (pend) { const uint32_t cur = bm[w]; if (!(cur & bit)) bm[w] = cur | bit; }
            __syncwarp();
            if (pend) pend = !(bm[w] & bit);
        } while (__any_sync(FULL, pend));
    }
    __syncwarp();
    // exclusive prefix popcount over the words, in word order (chunk q of every lane, then chunk q+4, ...)
    uint32_t uniq = 0;
    for (uint32_t q = 0; q < wpl; q += 4) {
        const uint4 v = *reinterpret_cast<const uint4 *>(bm + 4 * lane + 32 * q);
        const uint32_t c0 = __popc(v.x), c1 = __popc(v.y), c2 = __popc(v.z), c3 = __popc(v.w);
        const uint32_t cnt = c0 + c1 + c2 + c3;
        const uint32_t incl = warp_inclusive_scan(cnt);
        const uint32_t b0 = uniq + incl - cnt;
        uint2 pk;
        pk.x = b0 | ((b0 + c0) << 16);
        pk.y = (b0 + c0 + c1) | ((b0 + c0 + c1 + c2) << 16);
        *reinterpret_cast<uint2 *>(pre + 4 * lane + 32 * q) = pk;
        uniq += __shfl_sync(FULL, incl, 31);
    }
    __syncwarp();
    // rank of every partial product; the smallest position of a rank opens its place (slots ascend in position)
    for (uint32_t p0 = 0; p0 < len; p0 += 32) {
        const uint32_t p = p0 + lane;
        const bool valid = p < len;
        const uint32_t col = valid ? region[p].idx : 0;
        const uint32_t w = col >> 5;
        const uint32_t r = valid ? pre[w] + __popc(bm[w] & ((1u << (col & 31)) - 1)) : 0;
        if (valid) rnk[p] = uint16_t(r);
        bool want = valid;
        do {
            if (want && first[r] > p) first[r] = uint16_t(p);
            __syncwarp();
            want = valid && first[r] > p;
        } while (__any_sync(FULL, want));
    }
    __syncwarp();
    // first arrivals write their entry; later arrivals are flagged (top bit of rnk)
    bool any_loser = false;
    for (uint32_t p0 = 0; p0 < len; p0 += 32) {
        const uint32_t p = p0 + lane;
        if (p < len) {
            const uint32_t r = rnk[p];
            if (first[r] == p) bin[r] = region[p];
            else { rnk[p] = uint16_t(r | 0x8000u); any_loser = true; }
        }
    }
    __syncwarp();
    // later arrivals are added in arrival order: slot after slot, inside a slot lowest lane first
    if (__any_sync(FULL, any_loser)) {
        for (uint32_t p0 = 0; p0 < len; p0 += 32) {
            const uint32_t p = p0 + lane;
            const uint32_t rr = p < len ? rnk[p] : 0;
            bool pending = (rr & 0x8000u) != 0;
            const uint32_t r = rr & 0x7FFFu;
            while (__any_sync(FULL, pending)) {
                if (pending) { const uint32_t cur = first[r]; if (!(cur & 0x8000u) || (cur & 0x7FFFu) > lane) first[r] = uint16_t(0x8000u | lane); }
                __syncwarp();
                const bool mine = pending && first[r] == (0x8000u | lane);
                __syncwarp();
                if (mine) {
                    bin[r].val = __fadd_rn(__ldcg(&bin[r].val), region[p].val);
                    first[r] = 0;
                    pending = false;
                }
                __syncwarp();
            }
        }
    }
    return uniq;
}

// =====================================================================================
// Merge, the main kernel.  Every warp takes tiles of consecutive short rows (<= 32 rows, one lane
// per row; < MT_CAP + MT_LONG partial products) and merges the rows one after the other: register
// bitonic sort or bitmap rank, then the k-ordered left fold.  The compacted row is written back over
// the start of its own bin and uniq[row] = surviving entries.  Tiles are independent: no ordering,
// no look-back, no CTA barrier.  (A chained single-pass variant that wrote C directly was measured
// and dropped: with thousands of tiles in flight the look-back wave and in-order retirement cost more
// than the extra pass -- profiles/README.md.)  k_scan over uniq[] then gives C.pos and k_gather_rows
// moves the rows into C.data.
// K = uint32_t when col << 9 fits 32 bits (cols <= 2^23), else uint64_t; BM enables the bitmap method.
// =====================================================================================
constexpr int MW_THREADS = 256;

template <class K, bool BM>
__global__ void __launch_bounds__(MW_THREADS, BM ? 3 : 4)
k_merge_tiles(const uint64_t *__restrict__ row_bin, const uint64_t bin_base, Elem *bins,
              const uint32_t *__restrict__ tile_row, const uint32_t bm_words, const uint32_t t0, const uint32_t t1,
              uint32_t *__restrict__ uniq) {
    extern __shared__ __align__(16) unsigned char smem[];
    const unsigned int lane = lane_id(), warp = threadIdx.x >> 5;
    constexpr uint32_t PER_WARP = MT_LONG * 8 + (BM ? BM_SCRATCH : 0);
    Elem *region = reinterpret_cast<Elem *>(smem + warp * PER_WARP);
    unsigned char *scratch = smem + warp * PER_WARP + MT_LONG * 8;
    const uint32_t gwarp = blockIdx.x * (MW_THREADS / 32) + warp, nwarps = gridDim.x * (MW_THREADS / 32);

    for (uint32_t tile = t0 + gwarp; tile < t1; tile += nwarps) {
        const uint64_t r0 = tile_row[tile], r1 = tile_row[tile + 1];
        const uint32_t R = uint32_t(r1 - r0);                          // <= 32
        const uint64_t st_j = row_bin[min(r0 + lane, r1)];
        const uint64_t en_j = row_bin[min(r0 + lane + 1, r1)];
        const uint64_t len64 = en_j - st_j;
        if (R == 1 && __shfl_sync(FULL, len64, 0) > MT_LONG) continue;   // long row: k_merge_long / k_merge_xl
        const uint32_t len_j = lane < R ? uint32_t(len64) : 0;
        uint32_t uniq_j = len_j;                                        // rows of 0 / 1 partial products stay as they are
        unsigned int todo = __ballot_sync(FULL, len_j > 1);
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t len = __shfl_sync(FULL, len_j, j);
            Elem *bin = bins + (__shfl_sync(FULL, st_j, j) - bin_base);
            uint32_t u;
            if constexpr (BM) {
                if (len > 32) u = merge_row_bitmap(bin, region, len, bm_words, scratch, lane);
                else u = merge_row_regs<1, uint32_t>(bin, region, bin, len, lane);
            } else {
                if (len <= 32) u = merge_row_regs<1, K>(bin, region, bin, len, lane);
                else if (len <= 64) u = merge_row_regs<2, K>(bin, region, bin, len, lane);
                else if (len <= 128) u = merge_row_regs<4, K>(bin, region, bin, len, lane);
                else if (len <= 256) u = merge_row_regs<8, K>(bin, region, bin, len, lane);
                else u = merge_row_regs<16, K>(bin, region, bin, len, lane);
            }
            if (int(lane) == j) uniq_j = u;
            __syncwarp();
        }
        if (lane < R) uniq[r0 + lane] = uniq_j;
    }
}

// Moves the merged rows (prefix of each bin) into the CSR data array of C.  Same tiles as the merge:
// lane j of a warp reads the bounds of row j, then the warp copies the rows one after the other
// (consecutive rows are consecutive in C, so the stores of a tile form one contiguous stream).
__global__ void __launch_bounds__(256)
k_gather_rows(const uint64_t *__restrict__ row_bin, const uint64_t bin_base, const Elem *__restrict__ bins,
              const uint32_t *__restrict__ tile_row, const uint32_t t0, const uint32_t t1,
              const uint32_t *__restrict__ uniq, const uint64_t *__restrict__ c_pos, Elem *__restrict__ c_data) {
    const unsigned int lane = lane_id();
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t tile = t0 + gwarp; tile < t1; tile += nwarps) {
        const uint64_t r0 = tile_row[tile], r1 = tile_row[tile + 1];
        const uint32_t R = uint32_t(r1 - r0);
        const uint64_t st_j = lane < R ? row_bin[r0 + lane] - bin_base : 0;
        const uint64_t dst_j = lane < R ? c_pos[r0 + lane] : 0;
        const uint32_t n_j = lane < R ? uniq[r0 + lane] : 0;
        const uint32_t n_max = __reduce_max_sync(FULL, n_j);
        if (n_max <= 4) {                       // tiny rows: one lane per row
            for (uint32_t i = 0; i < n_j; i++) c_data[dst_j + i] = bins[st_j + i];
            continue;
        }
        unsigned int todo = __ballot_sync(FULL, n_j > 0);
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t n = __shfl_sync(FULL, n_j, j);
            const Elem *src = bins + __shfl_sync(FULL, st_j, j);
            Elem *dst = c_data + __shfl_sync(FULL, dst_j, j);
            for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
        }
    }
}

// uniq[] -> C.pos for rows [r_lo, r_hi): exclusive scan plus the running total of the earlier row blocks
// (carry_in), whose successor is left in carry_out; the last block also closes C.pos.
struct U64OutCarry {
    uint64_t *y;
    const unsigned long long *carry_in;
    unsigned long long *carry_out;
    uint64_t n;
    __device__ void operator()(uint64_t i, uint64_t v, uint64_t) const {
        const uint64_t c = *carry_in;
        y[i] = c + v;
        if (i == n) *carry_out = c + v;
    }
};

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
    __device__ uint64_t load(uint64_t j, bool valid) const { return valid ? lens[(j % G) * RL + j / G] : 0; }
    __device__ uint64_t value(uint64_t v, bool) const { return v; }
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
