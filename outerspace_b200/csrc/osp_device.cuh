// osp_device.cuh -- device-side building blocks shared by every kernel of the engine.
//
// Data layout in HBM (DESIGN.md "Data layout"):
//   Elem   8 B  {uint32 idx; float val}   == reference CSRElement (simulator/common.h:10-16)
//   pos    8 B  uint64                    == reference CSRMatrix::pos (common.h:41)
//   Task  16 B  {uint32 bs; float a; uint64 off<<24|len}: one non-zero A(i,k) annotated with where row k
//               of B starts, and with the offset and length of its run of partial products inside the
//               output-row bins (one LDG.128 per task, no further look-ups in the multiply).
#pragma once
#ifdef OSP_CUSIM                      // tests/cusim: the same source on the CPU emulation of the execution model (tests only)
#ifdef __CUDACC__
#error "OSP_CUSIM is for the g++ test build under tests/cusim only: the product is compiled by nvcc for sm_100a and has no CPU path"
#endif
#include "cusim.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

// Dynamic shared memory of a kernel.  (tests/cusim: the block's emulated shared memory.)
#ifdef OSP_CUSIM
#define OSP_EXTERN_SMEM(name) unsigned char *name = cusim::dyn_smem()
#else
#define OSP_EXTERN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace osp {

struct __align__(8) Elem {
    uint32_t idx;
    float val;
};
static_assert(sizeof(Elem) == 8, "Elem must match the reference's packed CSRElement");

constexpr int TASK_LEN_BITS = 24;     // a run (row of B) holds < 2^24 partial products; offsets < 2^40
struct __align__(16) Task {
    uint32_t bs;       // offset of row k of B inside B's data array (nnz(B) < 2^32)
    float a;           // A(i,k)
    uint64_t offlen;   // (absolute offset of the run inside the partial-product bins << 24) | run length
};
static_assert(sizeof(Task) == 16, "Task is one 128-bit load");

// Scalars the kernels report back to the host (one pinned-mirror copy per sync).  Lives at the start
// of the per-call zeroed arena, so every counter starts at 0.
struct DevScalars {
    unsigned long long products;       // P
    unsigned long long cap_bound;      // sum_i min(len_i, cols) : upper bound of nnz(C)
    unsigned long long nnz_c[2];       // running nnz(C) over the row blocks (ping-pong carry of the C.pos scan)
    unsigned long long last_nonempty;  // 1 + index of the last row with partial products / non-zeros
    unsigned int err;                  // first OSP_ERR_* raised on the device
    unsigned int max_idx;              // reduction result of k_max_idx
    unsigned int n_tiles;              // merge tiles planned
    unsigned int n_long;               // rows queued for the CTA shared-memory sort
    unsigned int n_xl;                 // rows queued for the long-row (dense accumulator) kernel
    unsigned int tile_ticket;          // dynamic tile ids of the merge kernel
    unsigned int scan_ticket[4];       // dynamic tile ids of the look-back scans (one per scan of a call)
    unsigned int n_dups;               // rows that shrank while folding (duplicate check of csr2csc)
    unsigned int xl_ticket;            // dynamic row ids of k_merge_xl (zeroed before every launch)
    // k_validate (operand preconditions): positions p with d[p].idx <= d[p-1].idx, and how many of them are the first
    // element of a slice; the same for d[p].idx == d[p-1].idx.  Slices are ascending and duplicate-free exactly when
    // every descent sits on a slice boundary.
    unsigned long long v_desc, v_bdesc, v_eq, v_beq;
    unsigned long long v_bad_pos;      // pos[] entries that decrease or exceed nnz
    unsigned int kw_ticket;            // dynamic row ids of k_merge_ways (zeroed before every launch)
    unsigned int cut_tile;             // k_plan: index of the tile that starts at the requested cut row (k-sharded path: two halves)
    unsigned long long fl_quads;       // k_fl_count: quads (4 groups of 32 slots) handed out to the rows of B (osp_fusedlanes.cuh)
};

constexpr unsigned int FULL = 0xffffffffu;

__device__ __forceinline__ unsigned int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t pow2ceil(uint32_t v) {
    return v <= 1 ? 1u : 1u << (32 - __clz(v - 1));
}

// ---- decoupled look-back (single-pass chained scan) --------------------------------------
// One 64-bit status word per tile: flag in the top two bits, value in the low 62, so a single
// relaxed load sees a consistent (flag, value) pair.
constexpr uint64_t LB_FLAG_AGG = 1ull << 62;
constexpr uint64_t LB_FLAG_PREFIX = 2ull << 62;
constexpr uint64_t LB_VALUE_MASK = (1ull << 62) - 1;

#ifdef OSP_CUSIM
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t *p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
__device__ __forceinline__ void st_relaxed_u64(uint64_t *p, uint64_t v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
#else
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
#endif

// Called by every lane of ONE warp.  Publishes this tile's aggregate, walks back over the
// predecessors 32 at a time until an inclusive prefix is found, publishes the tile's own
// inclusive prefix and returns the exclusive prefix (same value in every lane).
// `first` is the id of the first tile of the chain (its exclusive prefix is `carry`).
__device__ __forceinline__ uint64_t lookback_exclusive(uint64_t *state, uint32_t tile, uint64_t aggregate,
                                                       uint32_t first = 0, uint64_t carry = 0) {
    const unsigned int lane = lane_id();
    if (tile == first) {
        if (lane == 0) st_relaxed_u64(state + tile, LB_FLAG_PREFIX | (carry + aggregate));
        return carry;
    }
    if (lane == 0) st_relaxed_u64(state + tile, LB_FLAG_AGG | aggregate);
    uint64_t exclusive = 0;
    int64_t base = int64_t(tile) - 1;
    while (true) {
        int64_t idx = base - lane;
        uint64_t word = LB_FLAG_PREFIX;  // tiles before `first` contribute an inclusive prefix of 0
        if (idx >= int64_t(first)) {
            word = ld_relaxed_u64(state + idx);
            unsigned int ns = 32;
            while ((word >> 62) == 0) {           // predecessor still merging: back off, do not hammer L2
                __nanosleep(ns);
                if (ns < 1024) ns <<= 1;
                word = ld_relaxed_u64(state + idx);
            }
        }
        unsigned int has_prefix = __ballot_sync(FULL, (word >> 62) == 2);
        unsigned int firstp = has_prefix ? (__ffs(has_prefix) - 1) : 31;
        uint64_t v = (lane <= firstp) ? (word & LB_VALUE_MASK) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        exclusive += v;
        if (has_prefix) break;
        base -= 32;
    }
    if (lane == 0) st_relaxed_u64(state + tile, LB_FLAG_PREFIX | (exclusive + aggregate));
    return exclusive;
}

// ---- warp / block scans -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t x) {
    const unsigned int lane = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(FULL, x, o);
        if (lane >= o) x += y;
    }
    return x;
}

// Block-wide exclusive scan of a small per-thread count (blockDim.x <= 1024).
// `warp_sums` is a 33-entry shared array.  Contains two __syncthreads().
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t x, uint32_t *warp_sums, uint32_t &total) {
    const unsigned int lane = lane_id(), warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    uint32_t incl = warp_inclusive_scan(x);
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < nwarps ? warp_sums[lane] : 0;
        uint32_t wi = warp_inclusive_scan(w);
        warp_sums[lane] = wi - w;                 // exclusive offsets of the warps
        if (lane == 31) warp_sums[32] = wi;       // block total
    }
    __syncthreads();
    total = warp_sums[32];
    uint32_t r = warp_sums[warp] + incl - x;
    return r;
}

}  // namespace osp
