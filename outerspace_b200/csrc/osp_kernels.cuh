// osp_kernels.cuh -- sm_100a kernels of the outer-product SpGEMM engine.
//
// Phase map (reference file:line it replaces, relative to simulator/):
//   k_scan<RunLenIn,...>    symbolic pass: flop count SimSpGEMM.cpp:884-891 and the implicit
//                           push_back sizing of SimOuterSPACE.cpp:87-92
//   k_col_hist/k_scatter_*  CSR->CSC conversion of A: coo2csr<true> SimSpGEMM.cpp:111-117,128-141
//   k_multiply              TaskProvider::multiplyPhase SimOuterSPACE.cpp:74-97 with the intended
//                           semantics of cscMulcsr SimSpGEMM.cpp:265-281 (true column ids)
//   k_merge_*               TaskProvider::mergePhase SimOuterSPACE.cpp:98-132 with the intended
//                           semantics of deduplicateCOO SimSpGEMM.cpp:519-535 (sum equal columns)
//   k_gather_rows           the CSR output container CSRMatrix mergedResult SimOuterSPACE.cpp:140
#pragma once
#include "osp_device.cuh"

namespace osp {

// =====================================================================================
// Generic single-pass exclusive scan (decoupled look-back).  In: idx -> uint64 contribution,
// Out: (idx, exclusive prefix); Out is also called once with (n, total).
// =====================================================================================
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

template <class In, class Out>
__global__ void __launch_bounds__(SCAN_BLOCK)
k_scan(In in, Out out, uint64_t n, uint64_t *tile_state, unsigned int *tile_counter) {
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_warp[SCAN_BLOCK / 32];
    __shared__ uint64_t s_tile_excl;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const unsigned int lane = lane_id(), warp = threadIdx.x >> 5;
    const uint64_t wbase = uint64_t(tile) * SCAN_TILE + uint64_t(warp) * (32 * SCAN_ITEMS);

    uint64_t excl[SCAN_ITEMS];
    uint64_t running = 0;
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; it++) {
        uint64_t idx = wbase + it * 32 + lane;
        uint64_t x = idx < n ? in(idx) : 0;
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
            if (tile == ntiles - 1) out(n, tile_excl + aggregate);
        }
    }
    __syncthreads();
    const uint64_t offset = s_tile_excl + s_warp[warp];
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; it++) {
        uint64_t idx = wbase + it * 32 + lane;
        if (idx < n) out(idx, offset + excl[it]);
    }
}

// ---- functors ---------------------------------------------------------------------------
// Symbolic pass over the non-zeros of A in row order: entry p contributes nnz(B(k_p,:)) partial
// products; the exclusive prefix is the offset of its run inside the row bins (bins are laid out
// row after row, runs inside a row in ascending k: the order in which multiplyPhase appends to
// multResults[rowId], SimOuterSPACE.cpp:91).
struct RunLenIn {
    const Elem *a;
    const uint64_t *b_pos;
    uint64_t n_k;
    DevScalars *sc;
    __device__ uint64_t operator()(uint64_t p) const {
        uint32_t k = a[p].idx;
        if (k >= n_k) { atomicMax(&sc->err, 4u); return 0; }   // OSP_ERR_INDEX
        return b_pos[k + 1] - b_pos[k];
    }
};
struct RunOffOut {
    uint64_t *run_off;   // [nnzA+1]
    DevScalars *sc;
    uint64_t n;
    __device__ void operator()(uint64_t p, uint64_t v) const {
        run_off[p] = v;
        if (p == n) sc->products = v;
    }
};
struct U32In {
    const uint32_t *x;
    __device__ uint64_t operator()(uint64_t i) const { return x[i]; }
};
struct U64Out {
    uint64_t *y;
    uint64_t carry;
    __device__ void operator()(uint64_t i, uint64_t v) const { y[i] = carry + v; }
};
struct U64OutTotal {      // like U64Out, and reports the total into DevScalars::block_nnz
    uint64_t *y;
    uint64_t carry;
    uint64_t n;
    DevScalars *sc;
    __device__ void operator()(uint64_t i, uint64_t v) const {
        y[i] = carry + v;
        if (i == n) sc->block_nnz = v;
    }
};
struct U32OutFromU64 {    // column pointers of the task list (nnzA < 2^32 is checked on the host)
    uint32_t *y;
    __device__ void operator()(uint64_t i, uint64_t v) const { y[i] = uint32_t(v); }
};

// =====================================================================================
// Small utility kernels
// =====================================================================================
__global__ void k_max_idx(const Elem *d, uint64_t nnz, DevScalars *sc) {
    uint32_t m = 0;
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < nnz; i += uint64_t(gridDim.x) * blockDim.x)
        m = max(m, d[i].idx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(FULL, m, o));
    if (lane_id() == 0 && m) atomicMax(&sc->max_idx, m);
}

// rows of C under the reference's rule numRows = maxRowId+1 (SimOuterSPACE.cpp:49-53) when A is
// given row-compressed: index of the last non-empty row + 1 (0 when A has no non-zeros).
__global__ void k_last_nonempty(const uint64_t *pos, uint64_t m, DevScalars *sc) {
    uint64_t nnz = pos[m];
    uint64_t lo = 0, hi = m;            // first i in [0,m] with pos[i] == nnz
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (pos[mid] >= nnz) hi = mid; else lo = mid + 1;
    }
    sc->last_nonempty = lo;             // rows 0..lo-1 cover every non-zero
}

// row_bin[i - r0] = offset of row i's bin relative to nothing (absolute), i in [r0, r1]; rows
// past the end of A's row pointer are empty.
__global__ void k_row_bins(const uint64_t *a_pos, uint64_t m_a, const uint64_t *run_off, uint64_t nnz_a,
                           uint64_t rows, uint64_t *row_bin) {
    uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (i > rows) return;
    uint64_t e = i <= m_a ? a_pos[i] : nnz_a;
    row_bin[i] = run_off[e];
}

// =====================================================================================
// CSR -> CSC conversion of A (column histogram -> exclusive scan -> scatter)
// =====================================================================================
__global__ void k_col_hist(const Elem *a, uint64_t e0, uint64_t e1, uint32_t *col_cnt) {
    for (uint64_t p = e0 + blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; p < e1; p += uint64_t(gridDim.x) * blockDim.x)
        atomicAdd(&col_cnt[a[p].idx], 1u);
}

// Task list in k-slice (CSC) order: every non-zero of A lands in the slot range of its column k
// together with the bin offset of its run.  The order of the tasks inside a column is immaterial
// (each task owns a distinct run), so slots are claimed with an atomic; `col_cnt` is counted down.
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
// One warp walks one source slice; the order inside a bucket is fixed up by the sort kernels.
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
// Merge phase, generic CTA path: one CTA sorts one row's partial products by (col, arrival
// position) with a bitonic network in shared memory, left-folds equal columns in arrival (= k)
// order with separately rounded adds, and writes the compacted row back to the start of its bin.
// Rows longer than `cap` are queued for k_merge_xl.
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
k_merge_cta(const uint64_t *__restrict__ row_bin, uint64_t bin_base, Elem *bins, uint32_t *uniq, uint64_t rows,
            uint32_t cap, uint32_t *xl_list, DevScalars *sc) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem);
    float *vals = reinterpret_cast<float *>(keys + cap);
    uint32_t *warp_sums = reinterpret_cast<uint32_t *>(vals + cap);
    for (uint64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const uint64_t start = row_bin[row] - bin_base;
        const uint64_t len64 = row_bin[row + 1] - row_bin[row];
        if (len64 <= 1) {
            if (threadIdx.x == 0) uniq[row] = uint32_t(len64);
            continue;
        }
        if (len64 > cap) {
            if (threadIdx.x == 0) xl_list[atomicAdd(&sc->xl_count, 1u)] = uint32_t(row);
            continue;
        }
        const uint32_t len = uint32_t(len64), N = pow2ceil(len);
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

// Long rows: the row is consumed in chunks of `cap` partial products.  Each chunk is sorted by
// (col, arrival position) and folded INTO a dense per-CTA accumulator acc[cols_b] (presence in
// bits[]), which keeps the left fold in arrival order across chunks.  The accumulator is then
// compacted in ascending column order over the start of the row's bin and bits[] is cleared.
__global__ void __launch_bounds__(256)
k_merge_xl(const uint64_t *__restrict__ row_bin, uint64_t bin_base, Elem *bins, uint32_t *uniq,
           const uint32_t *xl_list, const DevScalars *sc, uint32_t cap, float *acc_all, uint32_t *bits_all,
           uint64_t cols_b) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem);
    float *vals = reinterpret_cast<float *>(keys + cap);
    uint32_t *warp_sums = reinterpret_cast<uint32_t *>(vals + cap);
    const uint64_t words = (cols_b + 31) >> 5;
    float *acc = acc_all + uint64_t(blockIdx.x) * cols_b;
    uint32_t *bits = bits_all + uint64_t(blockIdx.x) * words;
    const uint32_t n_xl = sc->xl_count;
    for (uint32_t x = blockIdx.x; x < n_xl; x += gridDim.x) {
        const uint64_t row = xl_list[x];
        const uint64_t start = row_bin[row] - bin_base;
        const uint64_t len = row_bin[row + 1] - row_bin[row];
        Elem *bin = bins + start;
        for (uint64_t c0 = 0; c0 < len; c0 += cap) {
            const uint32_t n = uint32_t(min(uint64_t(cap), len - c0)), N = pow2ceil(n);
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

// Compacts the merged rows (prefix of each bin) into the CSR data array of C: one warp per row.
__global__ void k_gather_rows(const uint64_t *__restrict__ row_bin, uint64_t bin_base, const Elem *__restrict__ bins,
                              const uint32_t *__restrict__ uniq, const uint64_t *__restrict__ row_ptr, uint64_t rows,
                              Elem *__restrict__ c_data) {
    uint64_t warp = (blockIdx.x * uint64_t(blockDim.x) + threadIdx.x) >> 5;
    uint64_t nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    for (uint64_t r = warp; r < rows; r += nwarps) {
        const Elem *src = bins + (row_bin[r] - bin_base);
        Elem *dst = c_data + row_ptr[r];
        uint32_t n = uniq[r];
        for (uint32_t i = lane_id(); i < n; i += 32) dst[i] = src[i];
    }
}

// Duplicate check after the stable conversion: a bucket that shrank while folding held a duplicate.
__global__ void k_check_full(const uint64_t *pos, const uint32_t *uniq, uint64_t n, DevScalars *sc) {
    uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (i < n && uniq[i] != pos[i + 1] - pos[i]) atomicMax(&sc->err, 233u);
}

}  // namespace osp
