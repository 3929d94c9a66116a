// osp_chain2.cuh -- k_chain2: multiply + merge of the short rows in ONE warp-specialised kernel (round 2).
//
// Replaces, for the rows of at most MT_LONG (MT_LONG_BM) partial products, TaskProvider::multiplyPhase AND mergePhase
// (simulator/SimOuterSPACE.cpp:74-132; intended semantics cscMulcsr + deduplicateCOO, SimSpGEMM.cpp:265-281, 519-535):
// the partial products of a tile never exist in HBM.  Same tiles, same look-back chain and the same register sort /
// bitmap-rank merge as k_merge_chain (osp_kernels.cuh); what changes is who does what, and when:
//
//   * 4 MEMORY warps open tiles two ahead and fill their stage while the 8 SORT warps work (and, round 2 second version,
//     also stream the finished tiles out to C):
//       - tickets are taken two tiles ahead and the tile's start record (TileStart, written by k_plan: first row, first
//         task, first partial product -- entries t and t+1 are ONE 32-byte load) one tile ahead, so no dependent global
//         load sits on the path of a tile;
//       - the producers also classify the tile's rows by length and lay out the batches of the register sort
//         (what k_merge_chain's consumers did behind three CTA barriers and a single-thread section);
//       - the fill is the warp-flat walk of k_multiply with shared-memory stores: 32 tasks (non-zeros of A) per warp
//         and turn, the concatenation of their runs A(i,k) * B(k,:) walked 256 products at a time -- eight independent
//         loads in flight per lane, `ld.global.L2::64B` (measured: a random gather of 64-byte rows moves 111 B per row
//         from DRAM with it, 159 B without, profiles/r02_gather_probe.md).
//   * THREE stages in flight and NO barrier among the sort warps: a sort warp pulls batches of the current tile until
//     they run out, then moves on to the next tile on its own (a tile holds only ~1.5 batches per warp, so waiting at a
//     per-tile barrier for the slowest warp cost a third of the sort time -- measured on the first version, 32 % barrier
//     stalls).  The warp that finishes a tile's last batch scans the rows' survivor counts (one warp, 8 rows per lane),
//     publishes the tile's aggregate for the look-back and hands the tile to the memory warps (sorted[] mbarrier).
//   * the memory warps resolve the tile's offset in C and stream its output stage out two fills later -- its
//     predecessors have had that long to publish.
// Long rows (tiles of their own) still come merged from the bins (k_merge_xl / k_merge_long / k_merge_dense / k_long_fill).
// The generic source (`Src`) lets the k-sharded path reuse the kernel: there a "task" is one received segment
// (row, source rank) and the multiplier is exactly 1.
#pragma once
#include "osp_kernels.cuh"

namespace osp {

constexpr int C2_CW = 8, C2_PW = 4;                          // consumer / producer warps
constexpr int C2_CONS = C2_CW * 32, C2_PROD = C2_PW * 32, C2_THREADS = C2_CONS + C2_PROD;
constexpr int C2_FW = 4;                                     // 32-product chunks whose searches are interleaved in the fill
constexpr int C2_NS = 3;                                     // tiles in flight per CTA: being filled / sorted / streamed out
constexpr uint32_t C2_CAP = 1024, C2_CAP_BM = 640;            // partial products per tile (soft): 3 + 3 stages (+ multipliers) of two CTAs fit an SM
constexpr uint32_t C2_STAGE_ELEMS = C2_CAP + MT_LONG + 16, C2_STAGE_ELEMS_BM = C2_CAP_BM + MT_LONG_BM + 16;
constexpr int C2_NCLS = 9;                                   // batch kinds in processing order: sort classes 7 .. 0, then the rows of one product

// ---- hand-over primitives -------------------------------------------------------------------------------------------
#ifdef OSP_CUSIM   // tests/cusim: fibers of one OS thread; a waiting thread yields
__device__ __forceinline__ void mbar2_init(uint64_t *bar, uint32_t count) {
    uint32_t *w = reinterpret_cast<uint32_t *>(bar); w[0] = count; w[1] = count << 1;
}
__device__ __forceinline__ void mbar2_arrive(uint64_t *bar) {
    uint32_t *w = reinterpret_cast<uint32_t *>(bar);
    if (--w[0] == 0) { w[0] = w[1] >> 1; w[1] ^= 1u; cusim::g_progress++; }
}
__device__ __forceinline__ void mbar2_wait(uint64_t *bar, uint32_t parity) {
    volatile uint32_t *w = reinterpret_cast<volatile uint32_t *>(bar);
    while ((w[1] & 1u) == parity) cusim::yield();
}
__device__ __forceinline__ void named_bar(uint32_t id, uint32_t n, uint32_t *state) {      // state: [arrived, generation] per id
    volatile uint32_t *st = state + 2 * id;
    const uint32_t gen = st[1];
    if (st[0] + 1 == n) { st[0] = 0; st[1] = gen + 1; cusim::g_progress++; }
    else { st[0] = st[0] + 1; while (st[1] == gen) cusim::yield(); }
}
__device__ __forceinline__ void cp_async8(Elem *dst_smem, const Elem *src) { *dst_smem = *src; }     // completes at issue
__device__ __forceinline__ void mbar2_arrive_async(uint64_t *bar) { mbar2_arrive(bar); }
#else
__device__ __forceinline__ void mbar2_init(uint64_t *bar, uint32_t count) { mbar_init(bar, count); }
__device__ __forceinline__ void mbar2_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar2_wait(uint64_t *bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void named_bar(uint32_t id, uint32_t n, uint32_t *) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
// asynchronous 8-byte copy global -> shared (LDGSTS): no register, no waiting warp; 64-byte fetch granule
__device__ __forceinline__ void cp_async8(Elem *dst_smem, const Elem *src) {
    asm volatile("cp.async.ca.shared.global.L2::64B [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
// the barrier gets one arrival from this thread once all its earlier cp.async copies have landed
__device__ __forceinline__ void mbar2_arrive_async(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#endif

// ---- task sources ---------------------------------------------------------------------------------------------------
// task i -> where its run comes from (`src`), its multiplier, its length and its absolute offset in the (virtual) bins.
struct C2SrcProduct {              // single GPU: task = non-zero A(i,k) in row order of A, run = a * B(k,:)
    const Elem *a_data;            // CSR(A)
    const uint64_t *run_off;       // [nnzA + 1] offset of every task's run (k_scan<SymIn, RunOffOut>)
    const uint32_t *task_bs;       // b_pos[k] of every task
    const Elem *b_data;
    __device__ __forceinline__ void load(uint64_t i, const Elem *&src, float &a, uint64_t &off, uint32_t &len) const {
        const uint64_t ro = run_off[i];
        a = a_data[i].val; off = ro;
        len = uint32_t(run_off[i + 1] - ro);
        src = b_data + task_bs[i];
    }
    static constexpr bool UNIT = false;
};
struct C2SrcSegments {             // k-sharded owner: task = the segment (row i, source s) of the landing buffer, multiplier 1
    const Elem *recv;              // landing buffer: G regions, one per source, each row-major over the owner's rows
    const uint64_t *src_off;       // [s * RL + i] start of the segment inside recv
    const uint64_t *dst_off;       // [i * G + s] offset of the segment in row-major order (= the plan's row_bin at s = 0)
    const uint32_t *lens;          // [s * RL + i]
    uint64_t RL; uint32_t G;
    __device__ __forceinline__ void load(uint64_t t, const Elem *&src, float &a, uint64_t &off, uint32_t &len) const {
        const uint64_t i = t / G, s = t - i * G;
        const uint64_t j = s * RL + i;
        len = lens[j]; src = recv + src_off[j]; off = dst_off[t]; a = 1.0f;
    }
    static constexpr bool UNIT = true;
};

template <bool BM>
struct __align__(16) Chain2Smem {
    static constexpr uint32_t STAGE_ELEMS = BM ? C2_STAGE_ELEMS_BM : C2_STAGE_ELEMS;
    Elem stage[C2_NS][STAGE_ELEMS];    // partial products of the tiles in flight
    Elem ostage[C2_NS][STAGE_ELEMS];   // their merged rows (XOR-swizzled), until they stream out
    float stage_a[C2_NS][STAGE_ELEMS]; // multiplier of every element of the stage: the A(i,k) of its run (stage holds raw rows of B)
    uint32_t rstart[C2_NS][MT_RMAX + 1];   // bin start of every row relative to the tile
    uint32_t rout[C2_NS][MT_RMAX];     // survivors per row, then their exclusive scan
    uint16_t order[C2_NS][MT_RMAX];    // the tile's rows grouped by batch kind
    uint32_t cls_cnt[C2_NS][C2_NCLS], cls_off[C2_NS][C2_NCLS], cls_b0[C2_NS][C2_NCLS + 1];   // per kind: rows, start in order[], first batch
    uint32_t next_batch[C2_NS], done_warps[C2_NS];
    uint64_t d_r0[C2_NS], d_g0[C2_NS]; // tile descriptors (written by the memory warps; n_out by the tile's last sort warp)
    uint32_t d_idx[C2_NS], d_R[C2_NS], d_nin[C2_NS], d_long[C2_NS], d_nout[C2_NS];
    uint64_t full[C2_NS], sorted[C2_NS];   // mbarriers: stage filled (memory -> sort), tile merged (sort -> memory)
    uint64_t base;                     // offset in C of the tile being streamed out
    uint64_t p_r0, p_b0, p_e0, p_e1;   // memory warps: the record of the tile they open next (broadcast from warp 0)
    uint32_t p_idx, p_R, p_nin, p_long;
    uint32_t nb_state[8];              // tests/cusim only: state of the emulated named barriers
    __align__(16) unsigned char bm_scratch[BM ? C2_CW * BM_SCRATCH : 16];
};

template <class K, bool BM, class Src>
__global__ void __launch_bounds__(C2_THREADS, 2)
k_chain2(const TileStart *__restrict__ tiles, const uint64_t *__restrict__ row_bin, const uint64_t bin_base,
         const Elem *__restrict__ bins, const uint32_t t0, const uint32_t n_chain, const uint32_t *__restrict__ uniq,
         uint64_t *tile_state, DevScalars *sc, const int carry_slot, uint64_t *__restrict__ c_pos, Elem *__restrict__ c_data,
         const uint32_t bm_wpl, const uint32_t long_thresh, const Src src) {
    using Smem = Chain2Smem<BM>;
    constexpr bool MUL = !Src::UNIT;
    Smem &sm = *reinterpret_cast<Smem *>(osp_smem);
    const unsigned int tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    if (tid == 0) {
        // full: every memory thread arrives twice -- once for what it wrote itself, once when its asynchronous copies have landed
        for (int i = 0; i < C2_NS; i++) { mbar2_init(&sm.full[i], 2 * C2_PROD); mbar2_init(&sm.sorted[i], 1); }
        for (int i = 0; i < 8; i++) sm.nb_state[i] = 0;
    }
    __syncthreads();
    const uint64_t carry = sc->nnz_c[carry_slot];

    if (warp >= C2_CW) {
        // ============================================= memory warps ==============================================
        const unsigned int ptid = tid - C2_CONS, pw = warp - C2_CW;
        auto prod_bar = [&] { named_bar(2, C2_PROD, sm.nb_state); };
        // warp 0, lane 0 keeps two tickets in flight: `tk1` (taken last turn; its record is loaded this turn) and the
        // one taken now.  Nothing a tile depends on is loaded in the turn that uses it.
        uint32_t tk1 = 0;
        TileStart rec0{0, 0, 0}, rec1{0, 0, 0};                      // record of the tile opened next turn, and its successor
        if (pw == 0 && lane == 0) {
            const uint32_t tk0 = atomicAdd(&sc->tile_ticket, 1u);
            tk1 = atomicAdd(&sc->tile_ticket, 1u);
            if (tk0 < n_chain) { rec0 = tiles[t0 + tk0]; rec1 = tiles[t0 + tk0 + 1]; }
            sm.p_idx = tk0;
            sm.p_r0 = rec0.r0; sm.p_b0 = rec0.b0; sm.p_e0 = rec0.e0; sm.p_e1 = rec1.e0;
            sm.p_R = rec1.r0 - rec0.r0;
            const uint64_t n = rec1.b0 - rec0.b0;
            sm.p_long = (sm.p_R == 1 && n > long_thresh) ? 1u : 0u;
            sm.p_nin = sm.p_long ? 0u : uint32_t(n);
        }
        // Streams tile t (slot s) to its place in C once its last sort warp has handed it over.
        auto drain = [&](uint32_t t) {
            const uint32_t s = t % C2_NS;
            mbar2_wait(&sm.sorted[s], (t / C2_NS) & 1u);
            const uint32_t idx = sm.d_idx[s], R = sm.d_R[s], n_in = sm.d_nin[s], n_out = sm.d_nout[s];
            const bool is_long = sm.d_long[s] != 0;
            const uint64_t r0 = sm.d_r0[s], g0 = sm.d_g0[s];
            if (pw == 0) {
                const uint64_t excl = lb_resolve(tile_state, idx, n_out, carry);
                if (lane == 0) sm.base = excl;
            }
            prod_bar();
            const uint64_t base = sm.base;
            if (ptid == 0 && idx + 1 == n_chain) { c_pos[r0 + R] = base + n_out; sc->nnz_c[carry_slot ^ 1] = base + n_out; }
            if (is_long) {
                const Elem *srcp = bins + g0;
                uint32_t i = ptid;
                for (; i + 7 * C2_PROD < n_out; i += 8 * C2_PROD) {     // eight loads in flight per thread
                    Elem e[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) e[u] = srcp[i + u * C2_PROD];
#pragma unroll
                    for (int u = 0; u < 8; u++) c_data[base + i + u * C2_PROD] = e[u];
                }
                for (; i < n_out; i += C2_PROD) c_data[base + i] = srcp[i];
                if (ptid == 0) c_pos[r0] = base;
                return;
            }
            const uint32_t *rout = sm.rout[s];
            for (uint32_t j = ptid; j < R; j += C2_PROD) c_pos[r0 + j] = base + rout[j];
            const Elem *ostage = sm.ostage[s];
            if (n_out == n_in) {                                 // no duplicate column in the tile: rows are back to back
                for (uint32_t p = ptid; p < n_out; p += C2_PROD) c_data[base + p] = ostage[swz(p)];
            } else {
                const uint32_t *rstart = sm.rstart[s];
                for (uint32_t j = pw; j < R; j += C2_PW) {
                    const uint32_t rs = rstart[j], o = rout[j];
                    const uint32_t u = (j + 1 < R ? rout[j + 1] : n_out) - o;
                    for (uint32_t i = lane; i < u; i += 32) c_data[base + o + i] = ostage[swz(rs + i)];
                }
            }
        };
        // What a thread prefetches for the tile opened next turn: the starts of up to three of its rows and its (up to two)
        // tasks -- loads issued a turn ahead, so that nothing a tile depends on is waited for in the turn that uses it.
        uint64_t n_rb[3] = {0, 0, 0};
        const Elem *n_sp[2] = {nullptr, nullptr}; float n_a[2] = {0.f, 0.f}; uint64_t n_off[2] = {0, 0}; uint32_t n_len[2] = {0, 0};
        auto prefetch = [&](bool live, uint32_t is_long, uint64_t r0, uint32_t R, uint64_t e0, uint64_t e1) {
#pragma unroll
            for (int u = 0; u < 2; u++) { n_sp[u] = nullptr; n_a[u] = 0.f; n_off[u] = 0; n_len[u] = 0; }
            if (!live || is_long) return;
#pragma unroll
            for (int u = 0; u < 3; u++) {
                const uint32_t j = ptid + u * C2_PROD;
                if (j <= R) n_rb[u] = row_bin[r0 + j];
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const uint64_t i = e0 + ptid + u * C2_PROD;
                if (i < e1) src.load(i, n_sp[u], n_a[u], n_off[u], n_len[u]);
            }
        };
        prod_bar();
        prefetch(sm.p_idx < n_chain, sm.p_long, sm.p_r0, sm.p_R, sm.p_e0, sm.p_e1);
        for (uint32_t f = 0;; f++) {
            const uint32_t s = f % C2_NS;
            prod_bar();                                              // the record of tile f is in shared memory; the drain before is over
            const uint32_t idx = sm.p_idx, R = sm.p_R, n_in = sm.p_nin, is_long = sm.p_long;
            const uint64_t r0 = sm.p_r0, b0 = sm.p_b0, e0 = sm.p_e0, e1 = sm.p_e1;
            const bool live = idx < n_chain;
            prod_bar();                                              // everybody has read it
            // ---- the record after next: ticket now, loads now, consumed at the end of this turn ----
            uint32_t tk2 = 0;
            if (pw == 0 && lane == 0 && live) {
                tk2 = atomicAdd(&sc->tile_ticket, 1u);
                if (tk1 < n_chain) { rec0 = tiles[t0 + tk1]; rec1 = tiles[t0 + tk1 + 1]; }
            }
            // slot s is free: tile f - 3 was streamed out in the turn before
            if (ptid == 0) {
                sm.d_idx[s] = live ? idx : n_chain; sm.d_R[s] = R; sm.d_nin[s] = n_in; sm.d_long[s] = is_long;
                sm.d_r0[s] = r0; sm.d_g0[s] = b0 - bin_base;
                sm.next_batch[s] = 0; sm.done_warps[s] = 0;
                sm.cls_b0[s][C2_NCLS] = 0;
            }
            if (ptid < C2_NCLS) sm.cls_cnt[s][ptid] = 0;
            if (live && !is_long) {
                uint32_t *rstart = sm.rstart[s], *rout = sm.rout[s];
#pragma unroll
                for (int u = 0; u < 3; u++) {
                    const uint32_t j = ptid + u * C2_PROD;
                    if (j <= R) rstart[j] = uint32_t(n_rb[u] - b0);
                }
                prod_bar();                                          // rstart complete, cls_cnt zeroed
                // batch kinds, in the order they are handed out: kind k < 8 sorts rows of <= 8 << (7 - k) partial products,
                // 32 >> (7 - k) rows per warp at a time; kind 8 copies rows of ONE partial product, 32 per batch
                uint32_t my_k[2] = {C2_NCLS, C2_NCLS}, my_pos[2] = {0, 0};
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const uint32_t j = ptid + u * C2_PROD;
                    if (j < R) {
                        const uint32_t len = rstart[j + 1] - rstart[j];
                        if (len <= 1) rout[j] = len;                 // nothing to merge
                        if (len >= 1) {
                            const uint32_t c = len <= 8 ? 0u : 29u - uint32_t(__clz(len - 1));
                            my_k[u] = len == 1 ? 8u : 7u - c;
                            my_pos[u] = atomicAdd(&sm.cls_cnt[s][my_k[u]], 1u);
                        }
                    }
                }
                prod_bar();
                if (ptid == 0) {
                    uint32_t off = 0, b = 0;
#pragma unroll
                    for (int k = 0; k < C2_NCLS; k++) {
                        const uint32_t n = sm.cls_cnt[s][k];
                        const int c = 7 - k;                          // (k == 8: c = -1, the copy batches)
                        sm.cls_off[s][k] = off; sm.cls_b0[s][k] = b;
                        off += n;
                        b += k == 8 ? (n + 31) >> 5 : c >= 6 || (BM && c == 5) ? n : (n + (32u >> c) - 1) >> (5 - c);
                    }
                    sm.cls_b0[s][C2_NCLS] = b;
                }
                prod_bar();
#pragma unroll
                for (int u = 0; u < 2; u++)
                    if (my_k[u] < C2_NCLS) sm.order[s][sm.cls_off[s][my_k[u]] + my_pos[u]] = uint16_t(ptid + u * C2_PROD);
                // ---- fill: tasks [e0, e1), 32 per warp and round; their runs are walked 32 products per lane-step and every
                //      element goes from B to the stage as an asynchronous 8-byte copy, its multiplier beside it ----
                Elem *stage = sm.stage[s];
                float *stage_a = sm.stage_a[s];
                uint32_t round = 0;
                for (uint64_t base = e0; base < e1; base += C2_PROD, round++) {
                    const Elem *sp = nullptr; uint32_t len = 0, off = 0; float a = 0.f;
                    if (round == 0) {                                // prefetched a turn ago
                        sp = n_sp[0]; a = n_a[0]; len = n_len[0]; off = uint32_t(n_off[0] - b0);
                    } else if (round == 1) {
                        sp = n_sp[1]; a = n_a[1]; len = n_len[1]; off = uint32_t(n_off[1] - b0);
                    } else if (base + ptid < e1) {                   // (tiles of many very short runs: more than 256 tasks)
                        uint64_t o64;
                        src.load(base + ptid, sp, a, o64, len);
                        off = uint32_t(o64 - b0);
                    }
                    const uint32_t incl = warp_inclusive_scan(len);
                    const uint32_t total = __shfl_sync(FULL, incl, 31);
                    const uint32_t excl = incl - len;
                    const uint64_t dsp = uint64_t(reinterpret_cast<uintptr_t>(sp)) - uint64_t(excl) * 8;   // element q of the concatenation lives at dsp + 8 q
                    const uint32_t doff = off - excl;                // ... and lands at stage[doff + q]
                    for (uint32_t q0 = 0; q0 < total; q0 += 32 * C2_FW) {
                        uint32_t q[C2_FW], t[C2_FW];
#pragma unroll
                        for (int u = 0; u < C2_FW; u++) { q[u] = q0 + 32 * u + lane; t[u] = 0; }
#pragma unroll
                        for (int step = 16; step > 0; step >>= 1) {
#pragma unroll
                            for (int u = 0; u < C2_FW; u++) {
                                const uint32_t v = __shfl_sync(FULL, incl, t[u] + step - 1);
                                if (v <= q[u]) t[u] += step;
                            }
                        }
#pragma unroll
                        for (int u = 0; u < C2_FW; u++) {
                            const uint64_t p = __shfl_sync(FULL, dsp, t[u] & 31);
                            const float a_t = __shfl_sync(FULL, a, t[u] & 31);
                            const uint32_t d = __shfl_sync(FULL, doff, t[u] & 31) + q[u];
                            if (q[u] < total) {
                                cp_async8(stage + d, reinterpret_cast<const Elem *>(uintptr_t(p)) + q[u]);
                                if (!Src::UNIT) stage_a[d] = a_t;
                            }
                        }
                    }
                }
            }
            mbar2_arrive(&sm.full[s]);                               // descriptor, row tables, batches, multipliers: written
            mbar2_arrive_async(&sm.full[s]);                         // ... and the copies of this thread: landed
            // ---- the next record goes round; its rows and tasks are fetched while the tile two back streams out ----
            if (pw == 0 && lane == 0 && live) {
                const bool nl = tk1 < n_chain;
                sm.p_idx = tk1;
                sm.p_r0 = rec0.r0; sm.p_b0 = rec0.b0; sm.p_e0 = rec0.e0; sm.p_e1 = rec1.e0;
                const uint32_t nR = nl ? rec1.r0 - rec0.r0 : 0u;
                const uint64_t n = nl ? rec1.b0 - rec0.b0 : 0ull;
                sm.p_R = nR;
                sm.p_long = (nR == 1 && n > long_thresh) ? 1u : 0u;
                sm.p_nin = sm.p_long ? 0u : uint32_t(n);
                tk1 = tk2;
            }
            prod_bar();
            if (live) prefetch(sm.p_idx < n_chain, sm.p_long, sm.p_r0, sm.p_R, sm.p_e0, sm.p_e1);
            if (f >= 2) drain(f - 2);
            if (!live) {                                             // tile f does not exist: f - 1 is this CTA's last one
                if (f >= 1) { prod_bar(); drain(f - 1); }
                break;
            }
        }
        return;
    }

    // ================================================ sort warps =================================================
    for (uint32_t ct = 0;; ct++) {
        const uint32_t s = ct % C2_NS;
        mbar2_wait(&sm.full[s], (ct / C2_NS) & 1u);
        const uint32_t idx = sm.d_idx[s];
        if (idx >= n_chain) break;
        const uint32_t R = sm.d_R[s];
        const bool is_long = sm.d_long[s] != 0;
        uint32_t *rstart = sm.rstart[s], *rout = sm.rout[s];
        if (!is_long) {
            Elem *ostage = sm.ostage[s];
            const Elem *stage = sm.stage[s];
            const uint32_t stage_off = uint32_t(offsetof(Smem, stage)) + s * uint32_t(sizeof(Elem) * Smem::STAGE_ELEMS);
            const uint32_t ost_off = uint32_t(offsetof(Smem, ostage)) + s * uint32_t(sizeof(Elem) * Smem::STAGE_ELEMS);
            const uint32_t sa_off = uint32_t(offsetof(Smem, stage_a)) + s * uint32_t(sizeof(float) * Smem::STAGE_ELEMS);
            const uint32_t *cls_b0 = sm.cls_b0[s], *cls_cnt = sm.cls_cnt[s], *cls_off = sm.cls_off[s];
            const uint32_t n_batches = cls_b0[C2_NCLS];
            while (true) {
                uint32_t b = 0;
                if (lane == 0) b = atomicAdd(&sm.next_batch[s], 1u);
                b = __shfl_sync(FULL, b, 0);
                if (b >= n_batches) break;
                int k = 0;
                while (k < C2_NCLS - 1 && b >= cls_b0[k + 1]) k++;
                if (k == 8) {                                    // rows of one partial product: sorted = merged
                    const uint32_t i = ((b - cls_b0[8]) << 5) + lane;
                    if (i < cls_cnt[8]) {
                        const uint32_t rs = rstart[sm.order[s][cls_off[8] + i]];
                        Elem e = stage[rs];
                        if (MUL) e.val = __fmul_rn(sm.stage_a[s][rs], e.val);
                        ostage[swz(rs)] = e;
                    }
                    continue;
                }
                const int c = 7 - k;
                const uint32_t T = c >= 6 ? 5u : uint32_t(c);
                const uint32_t i = c >= 6 || (BM && c == 5) ? b - cls_b0[k] : ((b - cls_b0[k]) << (5 - c)) + (lane >> T);
                const bool valid = i < cls_cnt[k];
                uint32_t j = 0, rs = 0, len = 0;
                if (valid) {
                    j = sm.order[s][cls_off[k] + i];
                    rs = rstart[j];
                    len = rstart[j + 1] - rs;
                }
                const uint32_t row_off = stage_off + rs * 8, ar = sa_off + rs * 4;
                uint32_t u;
                if (BM && c >= 5) {       // small column range: rows of > 128 partial products skip the sort
                    const uint32_t scr_off = uint32_t(offsetof(Smem, bm_scratch)) + warp * BM_SCRATCH;
                    if (c <= 6) u = merge_row_bitmap<16, MUL>(row_off, ost_off, rs, len, bm_wpl, scr_off, lane, ar);
                    else u = merge_row_bitmap<MT_LONG_BM / 32, MUL>(row_off, ost_off, rs, len, bm_wpl, scr_off, lane, ar);
                } else
                if (c <= 5) u = merge_rows_grouped<8, K, MUL>(c, row_off, ost_off, rs, len, lane, ar);
                else u = merge_rows_grouped<16, K, MUL>(5, row_off, ost_off, rs, len, lane, ar);
                if (valid && (lane & ((1u << T) - 1)) == 0) rout[j] = u;
            }
        }
        // ---- this warp is through with the tile; the last of the eight closes it ----
        __threadfence_block();
        __syncwarp();
        uint32_t last = 0;
        if (lane == 0) last = atomicAdd(&sm.done_warps[s], 1u) == uint32_t(C2_CW - 1) ? 1u : 0u;
        last = __shfl_sync(FULL, last, 0);
        if (last) {
            __threadfence_block();
            uint32_t n_out;
            if (is_long) {
                n_out = uniq[sm.d_r0[s]];
            } else {
                // offsets of the rows inside the tile: lane l owns rows 8 l .. 8 l + 7
                uint32_t v[8], sum = 0;
#pragma unroll
                for (int e = 0; e < 8; e++) { const uint32_t j = lane * 8 + e; v[e] = j < R ? rout[j] : 0u; sum += v[e]; }
                const uint32_t incl = warp_inclusive_scan(sum);
                uint32_t o = incl - sum;
#pragma unroll
                for (int e = 0; e < 8; e++) { const uint32_t j = lane * 8 + e; if (j < R) rout[j] = o; o += v[e]; }
                n_out = __shfl_sync(FULL, incl, 31);
            }
            lb_publish(tile_state, idx, n_out, carry);
            if (lane == 0) sm.d_nout[s] = n_out;
            __threadfence_block();
            __syncwarp();
            if (lane == 0) mbar2_arrive(&sm.sorted[s]);
        }
    }
}

}  // namespace osp
