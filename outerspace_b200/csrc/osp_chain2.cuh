// osp_chain2.cuh -- k_chain2: multiply + merge of the short rows in ONE warp-specialised kernel (round 2).
//
// Replaces, for the rows of at most MT_LONG (MT_LONG_BM) partial products, TaskProvider::multiplyPhase AND mergePhase
// (simulator/SimOuterSPACE.cpp:74-132; intended semantics cscMulcsr + deduplicateCOO, SimSpGEMM.cpp:265-281, 519-535):
// the partial products of a tile never exist in HBM.  Same tiles, same look-back chain and the same register sort /
// bitmap-rank merge as k_merge_chain (osp_kernels.cuh); what changes is who does what, and when:
//
//   * 4 PRODUCER warps open the tile after next and fill its stage while the 8 CONSUMER warps sort the current one:
//       - tickets are taken two tiles ahead and the tile's start record (TileStart, written by k_plan: first row, first
//         task, first partial product -- entries t and t+1 are ONE 32-byte load) one tile ahead, so no dependent global
//         load sits on the path of a tile;
//       - the producers also classify the tile's rows by length and lay out the batches of the register sort
//         (what k_merge_chain's consumers did behind three CTA barriers and a single-thread section);
//       - the fill is the warp-flat walk of k_multiply with shared-memory stores: 32 tasks (non-zeros of A) per warp
//         and turn, the concatenation of their runs A(i,k) * B(k,:) walked 256 products at a time -- eight independent
//         loads in flight per lane, `ld.global.L2::64B` (measured: a random gather of 64-byte rows moves 111 B per row
//         from DRAM with it, 159 B without, profiles/r02_gather_probe.md).
//   * full[2] / empty[2] mbarriers hand the two stages back and forth; the consumers synchronise among themselves with
//     a named barrier (bar.sync 1, 256): three per tile instead of nine CTA-wide ones, no single-thread section.
//   * the look-back of the previous tile is resolved by consumer warp 0 at the START of a tile (its predecessors have
//     had a whole tile of slack) while the other warps already pull batches; the tile's output stage streams to C
//     one tile later, as before.
// Long rows (tiles of their own) still come merged from the bins (k_merge_xl / k_merge_long / k_merge_dense / k_long_fill).
// The generic source (`Src`) lets the k-sharded path reuse the kernel: there a "task" is one received segment
// (row, source rank) and the multiplier is exactly 1.
#pragma once
#include "osp_kernels.cuh"

namespace osp {

constexpr int C2_CW = 8, C2_PW = 4;                          // consumer / producer warps
constexpr int C2_CONS = C2_CW * 32, C2_PROD = C2_PW * 32, C2_THREADS = C2_CONS + C2_PROD;
constexpr int C2_FW = 8;                                     // 32-product chunks (= loads in flight per lane) per turn of the fill
constexpr int C2_META = 4;                                   // tiles whose row tables are alive: previous, current, next (+1 slack)
static_assert(MT_RMAX == C2_CONS, "one consumer thread per row of a tile");
constexpr int C2_CAP_SHIFT_BM = 10;                          // bitmap variant: tiles of < 1024 + 640 partial products (shared memory)
constexpr uint32_t C2_STAGE_ELEMS_BM = (1u << C2_CAP_SHIFT_BM) + MT_LONG_BM + 16;

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
__device__ __forceinline__ Elem ld_gather(const Elem *p) { return *p; }
#else
__device__ __forceinline__ void mbar2_init(uint64_t *bar, uint32_t count) { mbar_init(bar, count); }
__device__ __forceinline__ void mbar2_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar2_wait(uint64_t *bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void named_bar(uint32_t id, uint32_t n, uint32_t *) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
// gathered rows of B: fetch 64-byte granules, not whole 128-byte lines
__device__ __forceinline__ Elem ld_gather(const Elem *p) {
    Elem e; uint32_t v;
    asm volatile("ld.global.L2::64B.v2.u32 {%0, %1}, [%2];" : "=r"(e.idx), "=r"(v) : "l"(p));
    e.val = __uint_as_float(v);
    return e;
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
    static constexpr uint32_t STAGE_ELEMS = BM ? C2_STAGE_ELEMS_BM : MC_STAGE_ELEMS;
    Elem stage[2][STAGE_ELEMS];        // partial products of the tile being sorted / being filled
    Elem ostage[2][STAGE_ELEMS];       // output of the tile being sorted / of the tile awaiting its offset in C
    uint32_t rstart[C2_META][MT_RMAX + 1];   // bin start of every row relative to the tile
    uint32_t rout[C2_META][MT_RMAX];   // survivors per row, then their exclusive scan
    uint16_t order[2][MT_RMAX];        // per stage: the tile's rows grouped by size class, longest class first
    uint32_t cls_cnt[2][8], cls_off[2][8], cls_b0[2][9];
    uint32_t next_batch[2];
    uint64_t d_r0[2], d_g0[2];         // tile descriptors per stage (written by the producers)
    uint32_t d_idx[2], d_R[2], d_nin[2], d_long[2];
    uint64_t full[2], empty[2];        // mbarriers
    uint64_t base;                     // offset in C of the tile being retired
    uint64_t p_r0, p_b0, p_e0, p_e1;   // producers: the record of the tile they open next (broadcast from warp 0)
    uint32_t p_idx, p_R, p_nin, p_long;
    uint32_t warp_sums[C2_CW + 1];
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
    Smem &sm = *reinterpret_cast<Smem *>(osp_smem);
    const unsigned int tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    if (tid == 0) {
        mbar2_init(&sm.full[0], C2_PROD); mbar2_init(&sm.full[1], C2_PROD);
        mbar2_init(&sm.empty[0], 1); mbar2_init(&sm.empty[1], 1);
        for (int i = 0; i < 8; i++) sm.nb_state[i] = 0;
    }
    __syncthreads();

    if (warp >= C2_CW) {
        // =============================================== producers ===============================================
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
        for (uint32_t it = 0;; it++) {
            const uint32_t s = it & 1, m = it & (C2_META - 1);
            prod_bar();                                              // the record of tile `it` is in shared memory
            const uint32_t idx = sm.p_idx, R = sm.p_R, n_in = sm.p_nin, is_long = sm.p_long;
            const uint64_t r0 = sm.p_r0, b0 = sm.p_b0, e0 = sm.p_e0, e1 = sm.p_e1;
            const bool live = idx < n_chain;
            prod_bar();                                              // everybody has read it
            // ---- next turn's record: loads issued now, consumed after this tile's fill ----
            uint32_t tk2 = 0;
            if (pw == 0 && lane == 0 && live) {
                tk2 = atomicAdd(&sc->tile_ticket, 1u);
                if (tk1 < n_chain) { rec0 = tiles[t0 + tk1]; rec1 = tiles[t0 + tk1 + 1]; }
            }
            // ---- the rows' starts (global loads in flight while we wait for the stage) ----
            uint64_t rb[3] = {0, 0, 0};
            if (live && !is_long) {
#pragma unroll
                for (int u = 0; u < 3; u++) {
                    const uint32_t j = ptid + u * C2_PROD;
                    if (j <= R) rb[u] = row_bin[r0 + j];
                }
            }
            if (it >= 2) mbar2_wait(&sm.empty[s], ((it >> 1) & 1u) ^ 1u);     // the consumers are done with tile it - 2 of this stage
            if (ptid == 0) {
                sm.d_idx[s] = live ? idx : n_chain; sm.d_R[s] = R; sm.d_nin[s] = n_in; sm.d_long[s] = is_long;
                sm.d_r0[s] = r0; sm.d_g0[s] = b0 - bin_base;
                sm.next_batch[s] = 0;
            }
            if (ptid < 8) sm.cls_cnt[s][ptid] = 0;
            if (live && !is_long) {
                uint32_t *rstart = sm.rstart[m], *rout = sm.rout[m];
#pragma unroll
                for (int u = 0; u < 3; u++) {
                    const uint32_t j = ptid + u * C2_PROD;
                    if (j <= R) rstart[j] = uint32_t(rb[u] - b0);
                }
                prod_bar();                                          // rstart complete, cls_cnt zeroed
                // size classes: class c sorts rows of <= 8 << c partial products, 32 >> c rows per warp at a time
                uint32_t my_cls[2] = {8, 8}, my_pos[2] = {0, 0};
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const uint32_t j = ptid + u * C2_PROD;
                    if (j < R) {
                        const uint32_t len = rstart[j + 1] - rstart[j];
                        if (len <= 1) rout[j] = len;                 // rows of 0 / 1 partial products need no merge
                        else {
                            my_cls[u] = len <= 8 ? 0u : 29u - uint32_t(__clz(len - 1));
                            my_pos[u] = atomicAdd(&sm.cls_cnt[s][my_cls[u]], 1u);
                        }
                    }
                }
                prod_bar();
                if (ptid == 0) {
                    uint32_t off = 0, b = 0;
#pragma unroll
                    for (int c = 7; c >= 0; c--) {
                        const uint32_t n = sm.cls_cnt[s][c];
                        sm.cls_off[s][c] = off; sm.cls_b0[s][c] = b;
                        off += n;
                        b += c >= 6 || (BM && c == 5) ? n : (n + (32u >> c) - 1) >> (5 - c);
                    }
                    sm.cls_b0[s][8] = b;
                }
                prod_bar();
#pragma unroll
                for (int u = 0; u < 2; u++)
                    if (my_cls[u] < 8) sm.order[s][sm.cls_off[s][my_cls[u]] + my_pos[u]] = uint16_t(ptid + u * C2_PROD);
                // ---- fill: tasks [e0, e1), 32 per warp and turn, their runs walked C2_FW * 32 products at a time ----
                Elem *stage = sm.stage[s];
                for (uint64_t base = e0 + pw * 32; base < e1; base += C2_PW * 32) {
                    const Elem *sp = nullptr; uint32_t len = 0, off = 0; float a = 0.f;
                    if (base + lane < e1) {
                        uint64_t o64;
                        src.load(base + lane, sp, a, o64, len);
                        off = uint32_t(o64 - b0);
                    }
                    const uint32_t incl = warp_inclusive_scan(len);
                    const uint32_t total = __shfl_sync(FULL, incl, 31);
                    const uint32_t excl = incl - len;
                    const uint64_t dsp = uint64_t(reinterpret_cast<uintptr_t>(sp)) - uint64_t(excl) * 8;   // element q of the concatenation lives at dsp + 8 q
                    const uint32_t doff = off - excl;                // ... lands at stage[doff + q]
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
                        Elem b[C2_FW]; float a_t[C2_FW]; uint32_t doff_t[C2_FW];
#pragma unroll
                        for (int u = 0; u < C2_FW; u++) {
                            const uint64_t p = __shfl_sync(FULL, dsp, t[u] & 31);
                            a_t[u] = __shfl_sync(FULL, a, t[u] & 31);
                            doff_t[u] = __shfl_sync(FULL, doff, t[u] & 31);
                            if (q[u] < total) b[u] = ld_gather(reinterpret_cast<const Elem *>(uintptr_t(p)) + q[u]);
                        }
#pragma unroll
                        for (int u = 0; u < C2_FW; u++)
                            if (q[u] < total) {
                                Elem o; o.idx = b[u].idx;
                                o.val = Src::UNIT ? b[u].val : __fmul_rn(a_t[u], b[u].val);      // rounded on its own: no FMA
                                stage[doff_t[u] + q[u]] = o;
                            }
                    }
                }
            }
            mbar2_arrive(&sm.full[s]);                               // descriptor, row tables, batches and stage are complete
            if (!live) break;
            // ---- hand the next record to the other producer warps ----
            if (pw == 0 && lane == 0) {
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
        }
        return;
    }

    // ================================================= consumers =================================================
    auto cons_bar = [&] { named_bar(1, C2_CONS, sm.nb_state); };
    const uint64_t carry = sc->nnz_c[carry_slot];
    TileDesc prev; prev.idx = 0xFFFFFFFFu; prev.R = prev.n_in = prev.n_out = 0; prev.r0 = prev.g0 = 0; prev.is_long = prev.last = false;
    uint32_t prev_m = 0, prev_s = 0;
    // Streams tile `d` (row tables in slot pm, output stage ob) to its place in C; sm.base holds its offset.
    auto write_out = [&](const TileDesc &d, uint32_t pm, uint32_t ob) {
        const uint64_t base = sm.base;
        if (tid == 0 && d.last) { c_pos[d.r0 + d.R] = base + d.n_out; sc->nnz_c[carry_slot ^ 1] = base + d.n_out; }
        if (d.is_long) {
            const Elem *srcp = bins + d.g0;
            uint32_t i = tid;
            for (; i + 7 * C2_CONS < d.n_out; i += 8 * C2_CONS) {      // eight loads in flight per thread
                Elem e[8];
#pragma unroll
                for (int u = 0; u < 8; u++) e[u] = srcp[i + u * C2_CONS];
#pragma unroll
                for (int u = 0; u < 8; u++) c_data[base + i + u * C2_CONS] = e[u];
            }
            for (; i < d.n_out; i += C2_CONS) c_data[base + i] = srcp[i];
            if (tid == 0) c_pos[d.r0] = base;
            return;
        }
        const uint32_t *rout = sm.rout[pm];
        if (tid < d.R) c_pos[d.r0 + tid] = base + rout[tid];
        const Elem *ostage = sm.ostage[ob];
        if (d.n_out == d.n_in) {                             // no duplicate column in the tile: rows are back to back
            for (uint32_t p = tid; p < d.n_out; p += C2_CONS) c_data[base + p] = ostage[swz(p)];
        } else {
            const uint32_t *rstart = sm.rstart[pm];
            for (uint32_t j = warp; j < d.R; j += C2_CW) {
                const uint32_t s = rstart[j], o = rout[j];
                const uint32_t u = (j + 1 < d.R ? rout[j + 1] : d.n_out) - o;
                for (uint32_t i = lane; i < u; i += 32) c_data[base + o + i] = ostage[swz(s + i)];
            }
        }
    };

    for (uint32_t it = 0;; it++) {
        const uint32_t s = it & 1, m = it & (C2_META - 1);
        mbar2_wait(&sm.full[s], (it >> 1) & 1u);
        TileDesc cur;
        cur.idx = sm.d_idx[s]; cur.R = sm.d_R[s]; cur.n_in = sm.d_nin[s]; cur.is_long = sm.d_long[s] != 0;
        cur.r0 = sm.d_r0[s]; cur.g0 = sm.d_g0[s];
        cur.last = cur.idx + 1 == n_chain; cur.n_out = 0;
        if (cur.idx >= n_chain) break;
        // ---- the previous tile's place in C: its predecessors have had this long to publish ----
        if (warp == 0 && prev.idx != 0xFFFFFFFFu) {
            const uint64_t excl = lb_resolve(tile_state, prev.idx, prev.n_out, carry);
            if (lane == 0) sm.base = excl;
        }
        uint32_t *rstart = sm.rstart[m], *rout = sm.rout[m];
        if (!cur.is_long) {
            Elem *ostage = sm.ostage[s];
            const Elem *stage = sm.stage[s];
            const uint32_t stage_off = uint32_t(offsetof(Smem, stage)) + s * uint32_t(sizeof(Elem) * Smem::STAGE_ELEMS);
            const uint32_t ost_off = uint32_t(offsetof(Smem, ostage)) + s * uint32_t(sizeof(Elem) * Smem::STAGE_ELEMS);
            if (tid < cur.R && rstart[tid + 1] - rstart[tid] == 1) ostage[swz(rstart[tid])] = stage[rstart[tid]];
            const uint32_t n_batches = sm.cls_b0[s][8];
            const uint32_t *cls_b0 = sm.cls_b0[s], *cls_cnt = sm.cls_cnt[s], *cls_off = sm.cls_off[s];
            while (true) {
                uint32_t b = 0;
                if (lane == 0) b = atomicAdd(&sm.next_batch[s], 1u);
                b = __shfl_sync(FULL, b, 0);
                if (b >= n_batches) break;
                int c = 7;
                while (c > 0 && b >= cls_b0[c - 1]) c--;        // cls_b0 ascends from class 7 down to class 0
                const uint32_t T = c >= 6 ? 5u : uint32_t(c);
                const uint32_t idx = c >= 6 || (BM && c == 5) ? b - cls_b0[c] : ((b - cls_b0[c]) << (5 - c)) + (lane >> T);
                const bool valid = idx < cls_cnt[c];
                uint32_t j = 0, rs = 0, len = 0;
                if (valid) {
                    j = sm.order[s][cls_off[c] + idx];
                    rs = rstart[j];
                    len = rstart[j + 1] - rs;
                }
                const uint32_t row_off = stage_off + rs * 8;
                uint32_t u;
                if (BM && c >= 5) {       // small column range: rows of > 128 partial products skip the sort
                    const uint32_t scr_off = uint32_t(offsetof(Smem, bm_scratch)) + warp * BM_SCRATCH;
                    if (c <= 6) u = merge_row_bitmap<16>(row_off, ost_off, rs, len, bm_wpl, scr_off, lane);
                    else u = merge_row_bitmap<MT_LONG_BM / 32>(row_off, ost_off, rs, len, bm_wpl, scr_off, lane);
                } else
                switch (c) {
                    case 0: u = merge_rows_grouped<8, 0, K>(row_off, ost_off, rs, len, lane); break;
                    case 1: u = merge_rows_grouped<8, 1, K>(row_off, ost_off, rs, len, lane); break;
                    case 2: u = merge_rows_grouped<8, 2, K>(row_off, ost_off, rs, len, lane); break;
                    case 3: u = merge_rows_grouped<8, 3, K>(row_off, ost_off, rs, len, lane); break;
                    case 4: u = merge_rows_grouped<8, 4, K>(row_off, ost_off, rs, len, lane); break;
                    case 5: u = merge_rows_grouped<8, 5, K>(row_off, ost_off, rs, len, lane); break;
                    default: u = merge_rows_grouped<16, 5, K>(row_off, ost_off, rs, len, lane); break;
                }
                if (valid && (lane & ((1u << T) - 1)) == 0) rout[j] = u;
            }
        }
        cons_bar();                                              // A: every row is merged, the stage is consumed, sm.base is set
        if (tid == 0) mbar2_arrive(&sm.empty[s]);
        // ---- offsets of the rows inside the tile, the tile's aggregate ----
        if (cur.is_long) {
            cur.n_out = uniq[cur.r0];
        } else {
            const uint32_t my_u = tid < cur.R ? rout[tid] : 0u;   // one row per consumer thread
            const uint32_t incl = warp_inclusive_scan(my_u);
            if (lane == 31) sm.warp_sums[warp] = incl;
            cons_bar();
            uint32_t woff = 0, total = 0;
#pragma unroll
            for (int w = 0; w < C2_CW; w++) {
                const uint32_t x = sm.warp_sums[w];
                if (uint32_t(w) < warp) woff += x;
                total += x;
            }
            if (tid < cur.R) rout[tid] = woff + incl - my_u;
            cur.n_out = total;
        }
        if (warp == 0) lb_publish(tile_state, cur.idx, cur.n_out, carry);
        // ---- the previous tile leaves ----
        if (prev.idx != 0xFFFFFFFFu) write_out(prev, prev_m, s ^ 1);
        cons_bar();                                              // B: ostage[s ^ 1], sm.base and warp_sums are free again
        prev = cur; prev_m = m; prev_s = s;
    }
    if (prev.idx != 0xFFFFFFFFu) {                               // the last tile of this CTA
        if (warp == 0) {
            const uint64_t excl = lb_resolve(tile_state, prev.idx, prev.n_out, carry);
            if (lane == 0) sm.base = excl;
        }
        cons_bar();
        write_out(prev, prev_m, prev_s);
    }
}

}  // namespace osp
