// osp_dist.inl -- k-sharded multi-GPU SpGEMM (one process per GPU) on top of the single-GPU engine.
//
// BASELINE.json north_star: "each GPU runs the outer products for its k-range, then partial products
// are exchanged by output-row ownership with an NCCL all-to-allv over NVLink and merged locally".
// NCCL is reached through dlopen("libnccl.so.2") so that the library has no link-time dependency on it
// (the CPU-only ABI test loads the library without NCCL) and shares the copy torch already loaded.
#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

NcclApi *nccl_api(std::string &err) {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        // 1. the copy this process already holds (torch's bundled one when torch was imported first);
        // 2. OSP_NCCL_LIB (the Python front end points it at torch's bundled copy, so that a later
        //    `import torch` finds the version it was built against under the same soname); 3. the system copy.
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h)
            if (const char *path = std::getenv("OSP_NCCL_LIB")) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (h) {
            api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
            api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
            api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
            api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(h, "ncclGroupStart"));
            api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
            api.Send = reinterpret_cast<decltype(api.Send)>(dlsym(h, "ncclSend"));
            api.Recv = reinterpret_cast<decltype(api.Recv)>(dlsym(h, "ncclRecv"));
            api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
            api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
            if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send &&
                api.Recv && api.AllGather && api.GetErrorString)
                api.handle = h;
        }
    }
    if (!api.handle) { err = "NCCL (libnccl.so.2) could not be loaded"; return nullptr; }
    return &api;
}

}  // namespace

struct osp_dist {
    osp_ctx *ctx = nullptr;
    NcclApi *nccl = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    DevBuf lens_send, lens_recv, recv_buf, bins2, src_off, dst_off, bounds_idx, bounds_dev, bounds_all;
    uint64_t *h_bounds = nullptr;      // pinned [world * (world + 2)]: per rank its G+1 bounds and its landing capacity
    // peer-memory exchange (the default): every rank maps the owners' bins (CUDA IPC) and its multiply
    // writes the partial products straight into them over NVLink
    bool p2p = true, mapped_once = false;
    DevBuf handles_dev, flags_dev;
    cudaIpcMemHandle_t *h_handles = nullptr;          // pinned [world]: the bins' IPC handles, all-gathered every call
    uint32_t *h_flags = nullptr;                      // pinned [world]
    cudaIpcMemHandle_t peer_handle[MAX_PEERS];        // what is currently mapped
    void *peer_ptr[MAX_PEERS] = {};
    bool peer_open[MAX_PEERS] = {};
    void *own_exported = nullptr;                     // allocation h_handles[rank] describes
    void *retired = nullptr;                          // previous bins allocation, freed once the peers have remapped
};

namespace {

#define NC(ctx, d, expr)                                                                        \
    do {                                                                                        \
        ncclResult_t _r = (expr);                                                               \
        if (_r != ncclSuccess)                                                                  \
            return fail(ctx, OSP_ERR_CUDA, std::string(#expr) + ": " + (d)->nccl->GetErrorString(_r)); \
    } while (0)

uint64_t block_begin(uint64_t rows, int world, int r) { return rows * uint64_t(r) / uint64_t(world); }

// Makes this rank's landing buffer hold `bytes` and maps every owner's landing buffer into this process.  The previous
// allocation of a grown buffer is kept until the next growth: peers may still have it mapped until they
// see the new handle in this call's all-gather.  Returns OSP_OK with d->p2p cleared when some rank could
// not map a peer (every rank then takes the NCCL path, consistently).
int map_peer_bins(osp_dist *d, uint64_t bytes) {
    osp_ctx *ctx = d->ctx;
    const int G = d->world, me = d->rank;
    (void)bytes;                                   // (the caller has grown the landing buffer, and every rank knows it succeeded)
    d->peer_ptr[me] = d->recv_buf.p;
    if (G == 1) return OSP_OK;
    if (d->own_exported != d->recv_buf.p) {
        CU(ctx, cudaIpcGetMemHandle(&d->h_handles[me], d->recv_buf.p));
        d->own_exported = d->recv_buf.p;
    }
    CU(ctx, d->handles_dev.reserve(size_t(G) * sizeof(cudaIpcMemHandle_t)));
    CU(ctx, d->flags_dev.reserve(size_t(G) * 4 + 4));
    unsigned char *hd = d->handles_dev.as<unsigned char>();
    CU(ctx, cudaMemcpyAsync(hd + size_t(me) * sizeof(cudaIpcMemHandle_t), &d->h_handles[me], sizeof(cudaIpcMemHandle_t),
                            cudaMemcpyHostToDevice, ctx->stream));
    NC(ctx, d, d->nccl->AllGather(hd + size_t(me) * sizeof(cudaIpcMemHandle_t), hd, sizeof(cudaIpcMemHandle_t), ncclUint8, d->comm,
                                  ctx->stream));
    CU(ctx, cudaMemcpyAsync(d->h_handles, hd, size_t(G) * sizeof(cudaIpcMemHandle_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    uint32_t ok = 1;
    for (int r = 0; r < G; r++) {
        if (r == me) continue;
        if (d->peer_open[r] && std::memcmp(&d->peer_handle[r], &d->h_handles[r], sizeof(cudaIpcMemHandle_t)) == 0) continue;
        if (d->peer_open[r]) { cudaIpcCloseMemHandle(d->peer_ptr[r]); d->peer_open[r] = false; }
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, d->h_handles[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
            continue;
        }
        d->peer_ptr[r] = p; d->peer_handle[r] = d->h_handles[r]; d->peer_open[r] = true;
    }
    // every rank learns whether every mapping exists (a rank that changed nothing still takes part)
    d->h_flags[me] = ok;
    uint32_t *fd = d->flags_dev.as<uint32_t>();
    CU(ctx, cudaMemcpyAsync(fd + me, &d->h_flags[me], 4, cudaMemcpyHostToDevice, ctx->stream));
    NC(ctx, d, d->nccl->AllGather(fd + me, fd, 1, ncclUint32, d->comm, ctx->stream));
    CU(ctx, cudaMemcpyAsync(d->h_flags, fd, size_t(G) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < G; r++)
        if (!d->h_flags[r]) d->p2p = false;
    if (!d->p2p)                                   // staged exchange from now on: nobody writes through a mapping again
        for (int r = 0; r < G; r++)
            if (d->peer_open[r]) { cudaIpcCloseMemHandle(d->peer_ptr[r]); d->peer_open[r] = false; }
    return OSP_OK;
}

}  // namespace

extern "C" {

int osp_dist_unique_id(void *id128) {
    if (!id128) return fail(nullptr, OSP_ERR_INVALID, "osp_dist_unique_id: NULL argument");
    std::string err;
    NcclApi *api = nccl_api(err);
    if (!api) return fail(nullptr, OSP_ERR_UNSUPPORTED, err);
    ncclUniqueId id;
    ncclResult_t r = api->GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, OSP_ERR_CUDA, std::string("ncclGetUniqueId: ") + api->GetErrorString(r));
    std::memcpy(id128, &id, sizeof(id));
    return OSP_OK;
}

int osp_dist_create(osp_ctx *ctx, const void *id128, int rank, int world, osp_dist **out) {
    if (!ctx || !id128 || !out || world < 1 || rank < 0 || rank >= world)
        return fail(ctx, OSP_ERR_INVALID, "osp_dist_create: bad argument");
    *out = nullptr;
    std::string err;
    NcclApi *api = nccl_api(err);
    if (!api) return fail(ctx, OSP_ERR_UNSUPPORTED, err);
    CU(ctx, cudaSetDevice(ctx->device));
    osp_dist *d = new osp_dist();
    d->ctx = ctx; d->nccl = api; d->rank = rank; d->world = world;
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclResult_t r = api->CommInitRank(&d->comm, world, id, rank);
    if (r != ncclSuccess) {
        delete d;
        return fail(ctx, OSP_ERR_CUDA, std::string("ncclCommInitRank: ") + api->GetErrorString(r));
    }
    cudaMallocHost(reinterpret_cast<void **>(&d->h_bounds), size_t(world) * (world + 6) * 8);
    // the buffers of the per-call agreement exist from here on: a call never has to allocate before it can tell its peers
    if (!d->h_bounds || d->bounds_idx.reserve((world + 6) * 8) != cudaSuccess || d->bounds_dev.reserve((world + 6) * 8) != cudaSuccess ||
        d->bounds_all.reserve(size_t(world) * (world + 6) * 8) != cudaSuccess || d->flags_dev.reserve(size_t(world) * 4 + 4) != cudaSuccess) {
        cudaGetLastError();
        osp_dist_destroy(d);
        return fail(ctx, OSP_ERR_OOM, "osp_dist_create: buffers of the per-call agreement");
    }
    cudaMallocHost(reinterpret_cast<void **>(&d->h_handles), size_t(world) * sizeof(cudaIpcMemHandle_t));
    cudaMallocHost(reinterpret_cast<void **>(&d->h_flags), size_t(world) * 4);
    std::memset(d->peer_handle, 0, sizeof(d->peer_handle));
    const char *mode = std::getenv("OSP_DIST_EXCHANGE");     // "nccl": staged all-to-allv + regroup instead of peer stores
    d->p2p = world <= MAX_PEERS && !(mode && std::string(mode) == "nccl");
    *out = d;
    return OSP_OK;
}

void osp_dist_destroy(osp_dist *d) {
    if (!d) return;
    cudaSetDevice(d->ctx->device);
    cudaStreamSynchronize(d->ctx->stream);
    if (d->comm) d->nccl->CommDestroy(d->comm);
    for (int r = 0; r < d->world && r < MAX_PEERS; r++)
        if (d->peer_open[r]) cudaIpcCloseMemHandle(d->peer_ptr[r]);
    for (DevBuf *b : {&d->lens_send, &d->lens_recv, &d->recv_buf, &d->bins2, &d->src_off, &d->dst_off, &d->bounds_idx,
                      &d->bounds_dev, &d->bounds_all, &d->handles_dev, &d->flags_dev})
        b->release();
    if (d->retired) cudaFree(d->retired);
    if (d->h_bounds) cudaFreeHost(d->h_bounds);
    if (d->h_handles) cudaFreeHost(d->h_handles);
    if (d->h_flags) cudaFreeHost(d->h_flags);
    delete d;
}

int osp_dist_rows(const osp_dist *d, uint64_t rows_c, uint64_t *row_begin, uint64_t *row_end) {
    if (!d) return fail(nullptr, OSP_ERR_INVALID, "osp_dist_rows: NULL argument");
    if (row_begin) *row_begin = block_begin(rows_c, d->world, d->rank);
    if (row_end) *row_end = block_begin(rows_c, d->world, d->rank + 1);
    return OSP_OK;
}

// This rank's shard: CSR(A[:, k0:k1]) with k ids relative to k0 (args->a_slices rows, OSP_A_IS_CSR
// required), CSR(B[k0:k1, :]) with args->n_k = k1 - k0 rows; args->rows_c = rows of C (the same on
// every rank, required) and args->cols_b = columns of C (required).  The result holds this rank's
// block of rows of C = sum over ranks of A_g * B_g: rows [row_begin, row_end) of osp_dist_rows.
int osp_dist_spgemm(osp_dist *d, const osp_spgemm_args *args, osp_result **out) {
    if (!d || !args || !out) return fail(d ? d->ctx : nullptr, OSP_ERR_INVALID, "osp_dist_spgemm: NULL argument");
    osp_ctx *ctx = d->ctx;
    *out = nullptr;
    if (cudaSetDevice(ctx->device) != cudaSuccess) { cudaGetLastError(); return fail(ctx, OSP_ERR_CUDA, "osp_dist_spgemm: cudaSetDevice"); }
    ctx->launches = 0;
    ctx->events_used = 0;
    ctx->call_id++;
    ctx->marks.clear();
    ctx->profile_kernels = args->flags & OSP_PROFILE_KERNELS;
    const int G = d->world, me = d->rank;
    const uint32_t W = uint32_t(G) + 6;       // record of a rank: G+1 bounds | 3 capacities | host status | device status
    const uint64_t m = args->rows_c, m_a = args->a_slices, n_k = args->n_k, cols_b = args->cols_b;
    const uint64_t R0 = block_begin(m, G, me), R1 = block_begin(m, G, me + 1), RL = R1 - R0;
    const bool validate = !(args->flags & OSP_NO_VALIDATE);
    int rc;
    Operands op;
    Arena ar;
    uint64_t st[4] = {0, 0, 0, 0};
    uint64_t *run_off = nullptr, *row_bin = nullptr;
    uint32_t *task_bs = nullptr;
    cudaEvent_t ev_begin = nullptr;

    // ---- everything a rank can fail at on its own happens in here; the verdict travels in the record, so that either every
    //      rank goes on to the exchange or none does (a rank that returned early used to leave its peers in the collectives)
    const int pre_rc = [&]() -> int {
        if (!(args->flags & OSP_A_IS_CSR) || !args->rows_c || !args->cols_b || !args->a_pos || !args->b_pos)
            return fail(ctx, OSP_ERR_INVALID, "osp_dist_spgemm: needs CSR(A) shards, rows_c and cols_b");
        if (args->a_slices > args->rows_c) return fail(ctx, OSP_ERR_INVALID, "osp_dist_spgemm: shard has more rows than C");
        if (args->n_k >= (1ull << 32) || args->rows_c >= (1ull << 32) || args->cols_b >= (1ull << 32) || RL * uint64_t(G) >= (1ull << 32))
            return fail(ctx, OSP_ERR_INVALID, "osp_dist_spgemm: dimensions must fit index_t (uint32)");
        int r = stage_operands(ctx, args, op);
        if (r) return r;
        ev_begin = next_event(ctx);
        st[0] = scan_tiles(std::max<uint64_t>(op.nnz_a, 1)); st[1] = plan_tiles(std::max<uint64_t>(RL, 1));
        st[2] = scan_tiles(std::max<uint64_t>(RL * G, 1)); st[3] = st[2];
        r = prepare_arena(ctx, st, 0, ar);
        if (r) return r;
        CU(ctx, ctx->run_off.reserve((op.nnz_a + 1) * 8));
        CU(ctx, ctx->task_bs.reserve(std::max<uint64_t>(op.nnz_a, 1) * 4));
        CU(ctx, ctx->row_bin.reserve((std::max(m, RL) + 1) * 8));
        CU(ctx, d->lens_send.reserve(std::max<uint64_t>(m, 1) * 4));
        CU(ctx, d->lens_recv.reserve(std::max<uint64_t>(RL * G, 1) * 4));
        CU(ctx, d->src_off.reserve((RL * G + 1) * 8));
        CU(ctx, d->dst_off.reserve((RL * G + 1) * 8));
        r = reserve_plan(ctx, std::max<uint64_t>(RL, 1), std::max<uint64_t>(RL, 1));
        if (r) return r;
        run_off = ctx->run_off.as<uint64_t>(); task_bs = ctx->task_bs.as<uint32_t>(); row_bin = ctx->row_bin.as<uint64_t>();
        if (validate) {      // the shard's operands: ascending duplicate-free slices, column ids below cols_b (k < n_k: symbolic pass)
            const ValidateOp opa{op.a_pos, op.a_data, m_a, op.nnz_a, 0, nullptr}, opb{op.b_pos, op.b_data, n_k, op.nnz_b, cols_b, nullptr};
            r = launch_validate(ctx, opa, opb, 2);
            if (r) return r;
        }
        if (op.nnz_a)
            LAUNCH(ctx, (k_scan<SymIn, RunOffOut>), unsigned(st[0]), SCAN_BLOCK, 0, SymIn{op.a_data, op.b_pos, n_k, nullptr, ctx->d_sc, task_bs, nullptr},
                   RunOffOut{run_off, ctx->d_sc, op.nnz_a}, op.nnz_a, ar.state[0], &ctx->d_sc->scan_ticket[0]);
        else
            CU(ctx, cudaMemsetAsync(run_off, 0, 8, ctx->stream));
        LAUNCH(ctx, k_shard_rows, grid_for(m + 1, 256, 1u << 30), 256, 0, op.a_pos, m_a, run_off, op.nnz_a, m, row_bin,
               d->lens_send.as<uint32_t>(), ctx->d_sc);
        return OSP_OK;
    }();
    const std::string pre_msg = pre_rc ? ctx->err : std::string();
    const uint64_t nnz_a = op.nnz_a;

    // ---- the call's all-gather: bounds_all[s][dst] = row_bin_s[R_dst], capacities, status --------------------------------
    {
        for (int r = 0; r <= G; r++) d->h_bounds[r] = block_begin(m, G, r);
        d->h_bounds[G + 1] = uint64_t(d->recv_buf.cap);
        d->h_bounds[G + 2] = uint64_t(d->bins2.cap);
        d->h_bounds[G + 3] = uint64_t(ctx->bins.cap);
        d->h_bounds[G + 4] = uint64_t(pre_rc);
        d->h_bounds[G + 5] = 0;
        CU(ctx, cudaMemcpyAsync(d->bounds_idx.p, d->h_bounds, W * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (pre_rc == OSP_OK) {
            OSP_KERNEL_LAUNCH(k_dist_record, 1, 256, 0, ctx->stream, row_bin, d->bounds_idx.as<uint64_t>(), uint32_t(G), W, ctx->d_sc,
                              d->bounds_dev.as<uint64_t>());
            ctx->launches++;
        } else {
            CU(ctx, cudaMemcpyAsync(d->bounds_dev.p, d->bounds_idx.p, W * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        NC(ctx, d, d->nccl->AllGather(d->bounds_dev.p, d->bounds_all.p, W, ncclUint64, d->comm, ctx->stream));
        CU(ctx, cudaMemcpyAsync(d->h_bounds, d->bounds_all.p, uint64_t(G) * W * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    auto word = [&](int s, uint32_t i) { return d->h_bounds[uint64_t(s) * W + i]; };
    auto bound = [&](int s, int r) { return word(s, uint32_t(r)); };
    for (int r = 0; r < G; r++) {
        const uint64_t host_code = word(r, uint32_t(G) + 4), dev_code = word(r, uint32_t(G) + 5);
        if (!host_code && !dev_code) continue;
        if (r == me) {
            if (host_code) return fail(ctx, int(host_code), pre_msg);
            if (dev_code == 233) return fail(ctx, OSP_ERR_DUPLICATE, "osp_dist_spgemm: duplicate (row,col) entry in this rank's shard");
            if (dev_code == 1) return fail(ctx, OSP_ERR_INVALID, "osp_dist_spgemm: a slice of this rank's shard is not ascending, or a pos array is broken");
            if (dev_code == 6) return fail(ctx, OSP_ERR_UNSUPPORTED, "osp_dist_spgemm: a row of one shard holds >= 2^32 partial products");
            return fail(ctx, OSP_ERR_INDEX, "osp_dist_spgemm: index out of range in this rank's shard (k beyond the shard's inner dimension or a column beyond cols_b)");
        }
    }
    for (int r = 0; r < G; r++)
        if (word(r, uint32_t(G) + 4) || word(r, uint32_t(G) + 5))
            return fail(ctx, OSP_ERR_INVALID, "osp_dist_spgemm: rank " + std::to_string(r) + " reported an error (code " +
                                              std::to_string(word(r, uint32_t(G) + 4) ? word(r, uint32_t(G) + 4) : word(r, uint32_t(G) + 5)) +
                                              "); no rank entered the exchange");
    const uint64_t P_local = bound(me, G);
    std::vector<uint64_t> recv_cnt(G), recv_off(G + 1, 0), p_owned(G, 0);
    for (int s = 0; s < G; s++) { recv_cnt[s] = bound(s, me + 1) - bound(s, me); recv_off[s + 1] = recv_off[s] + recv_cnt[s]; }
    const uint64_t P_owned = recv_off[G];
    for (int r = 0; r < G; r++)
        for (int s2 = 0; s2 < G; s2++) p_owned[r] += bound(s2, r + 1) - bound(s2, r);

    // ---- buffers that depend on the exchanged sizes.  Every rank knows every rank's sizes and capacities, so all agree,
    //      without talking, on whether anybody allocates; if so the outcome of the allocations goes round before anything else
    bool direct = d->p2p;
    {
        bool any_grow = direct && !d->mapped_once;
        for (int r = 0; r < G; r++) {
            const uint64_t need = std::max<uint64_t>(p_owned[r], 1) * 8 + 16;
            any_grow |= need > word(r, uint32_t(G) + 1) || need > word(r, uint32_t(G) + 2);
            if (!direct) any_grow |= std::max<uint64_t>(bound(r, G), 1) * 8 > word(r, uint32_t(G) + 3);
        }
        if (any_grow) {
            const uint64_t need = std::max<uint64_t>(P_owned, 1) * 8 + 16;
            int grow_rc = OSP_OK;
            if (need > d->bins2.cap && d->bins2.reserve(need) != cudaSuccess) { cudaGetLastError(); grow_rc = OSP_ERR_OOM; }
            if (!direct) {
                if (!grow_rc && d->recv_buf.reserve(need) != cudaSuccess) { cudaGetLastError(); grow_rc = OSP_ERR_OOM; }
                if (!grow_rc && ctx->bins.reserve(std::max<uint64_t>(P_local, 1) * 8) != cudaSuccess) { cudaGetLastError(); grow_rc = OSP_ERR_OOM; }
            } else if (!grow_rc && need > d->recv_buf.cap) {       // landing buffer: the old one stays until the peers have remapped
                if (d->retired) { cudaFree(d->retired); d->retired = nullptr; }
                d->retired = d->recv_buf.p;
                d->recv_buf.p = nullptr; d->recv_buf.cap = 0;
                if (d->recv_buf.reserve(need) != cudaSuccess) { cudaGetLastError(); grow_rc = OSP_ERR_OOM; }
            }
            // the verdicts go round
            d->h_flags[me] = uint32_t(grow_rc);
            CU(ctx, d->flags_dev.reserve(size_t(G) * 4 + 4));
            uint32_t *fd = d->flags_dev.as<uint32_t>();
            CU(ctx, cudaMemcpyAsync(fd + me, &d->h_flags[me], 4, cudaMemcpyHostToDevice, ctx->stream));
            NC(ctx, d, d->nccl->AllGather(fd + me, fd, 1, ncclUint32, d->comm, ctx->stream));
            CU(ctx, cudaMemcpyAsync(d->h_flags, fd, size_t(G) * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CU(ctx, cudaStreamSynchronize(ctx->stream));
            for (int r = 0; r < G; r++)
                if (d->h_flags[r])
                    return fail(ctx, r == me ? OSP_ERR_OOM : OSP_ERR_INVALID,
                                r == me ? std::string("osp_dist_spgemm: device memory for the exchange buffers") :
                                          "osp_dist_spgemm: rank " + std::to_string(r) + " could not allocate its exchange buffers; no rank entered the exchange");
            if (direct) {
                rc = map_peer_bins(d, need);                  // (handles and success flags go round: two more all-gathers)
                if (rc) return rc;
                direct = d->p2p;
                d->mapped_once = true;
                if (!direct) {                                // some peer is not mappable: every rank takes the staged exchange from now on
                    const bool ok = ctx->bins.reserve(std::max<uint64_t>(P_local, 1) * 8) == cudaSuccess;
                    if (!ok) cudaGetLastError();
                    d->h_flags[me] = ok ? 0u : 1u;
                    CU(ctx, cudaMemcpyAsync(fd + me, &d->h_flags[me], 4, cudaMemcpyHostToDevice, ctx->stream));
                    NC(ctx, d, d->nccl->AllGather(fd + me, fd, 1, ncclUint32, d->comm, ctx->stream));
                    CU(ctx, cudaMemcpyAsync(d->h_flags, fd, size_t(G) * 4, cudaMemcpyDeviceToHost, ctx->stream));
                    CU(ctx, cudaStreamSynchronize(ctx->stream));
                    for (int r = 0; r < G; r++)
                        if (d->h_flags[r]) return fail(ctx, OSP_ERR_OOM, "osp_dist_spgemm: device memory for the staged exchange");
                }
            }
        }
    }

    // ---- per-row counts to the owners --------------------------------------------------------------------------------------
    NC(ctx, d, d->nccl->GroupStart());
    for (int r = 0; r < G; r++) {
        const uint64_t b0 = block_begin(m, G, r), b1 = block_begin(m, G, r + 1);
        if (b1 > b0) NC(ctx, d, d->nccl->Send(d->lens_send.as<uint32_t>() + b0, b1 - b0, ncclUint32, r, d->comm, ctx->stream));
        if (RL) NC(ctx, d, d->nccl->Recv(d->lens_recv.as<uint32_t>() + uint64_t(r) * RL, RL, ncclUint32, r, d->comm, ctx->stream));
    }
    NC(ctx, d, d->nccl->GroupEnd());
    cudaEvent_t ev_sym = next_event(ctx), ev_mul = nullptr, ev_xchg = nullptr;

    // ---- the owner's plan needs the counts only: it runs on the second stream BESIDE the multiply, whose stores are bound by
    //      NVLink and leave the SMs mostly idle (scans + plan used to follow the exchange: ~0.35 ms that did not shrink with G).
    //      The exchange itself goes in TWO HALVES of every owner's rows: while the sources send the second half, the owners
    //      already regroup and merge the first (row blocks of the merge chain, carried nnz(C) like the single-GPU row blocks).
    const uint32_t *lens = d->lens_recv.as<uint32_t>();
    const uint64_t n_seg = RL * G;
    cudaStream_t main_stream = ctx->stream;
    const bool side = RL && !ctx->profile_kernels;               // (per-kernel event pairs assume one stream)
    // (measured: worth 0.07 ms at G = 8, nothing at G = 2, where half of the products stay local and the multiply is not NVLink-bound)
    const int H = (direct && G > 1 && (G >= 4 || std::getenv("OSP_DIST_TWO_PHASE")) && m >= uint64_t(4 * G) &&
                   !std::getenv("OSP_DIST_ONE_PHASE")) ? 2 : 1;
    const uint64_t cut_local = H == 2 ? RL / 2 : ~0ull;          // first row (of this owner's block) of the second half
    if (RL) {
        if (side) {
            CU(ctx, cudaEventRecord(ctx->ev_fork, main_stream));
            CU(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
            ctx->stream = ctx->stream2;                          // LAUNCH and sync_scalars use ctx->stream
        }
        const int plan_rc = [&]() -> int {
            LAUNCH(ctx, (k_scan<U32In, U64Out>), unsigned(st[2]), SCAN_BLOCK, 0, U32In{lens}, U64Out{d->src_off.as<uint64_t>()}, n_seg,
                   ar.state[2], &ctx->d_sc->scan_ticket[2]);
            LAUNCH(ctx, (k_scan<TransposedIn, U64Out>), unsigned(st[3]), SCAN_BLOCK, 0, TransposedIn{lens, RL, uint64_t(G)},
                   U64Out{d->dst_off.as<uint64_t>()}, n_seg, ar.state[3], &ctx->d_sc->scan_ticket[3]);
            LAUNCH(ctx, k_plan<RowBinStrided>, unsigned(st[1]), PLAN_BLOCK, 0, RowBinStrided{d->dst_off.as<uint64_t>(), uint64_t(G)}, RL,
                   cols_b, ctx->row_bin.as<uint64_t>(), ctx->tile_row.as<uint32_t>(), ctx->long_list.as<uint32_t>(),
                   ctx->xl_list.as<uint32_t>(), ar.state[1], ctx->d_sc, 1, plan_long_thresh(cols_b), ctx->tile_state.as<uint64_t>(),
                   static_cast<TileStart *>(nullptr), MT_CAP, cut_local);
            return OSP_OK;
        }();
        ctx->stream = main_stream;
        if (plan_rc) return plan_rc;
    }

    // ---- multiply; partial products to the owners --------------------------------------------------------------------------
    cudaEvent_t ev_half[2] = {nullptr, nullptr};
    if (direct) {
        PeerDst dst;
        std::memset(&dst, 0, sizeof(dst));
        dst.world = G;
        for (int r = 0; r < G; r++) {
            dst.base[r] = static_cast<Elem *>(d->peer_ptr[r]);
            dst.bound[r] = bound(me, r);
            uint64_t region = 0;                                   // where source `me` lands inside owner r's buffer
            for (int s2 = 0; s2 < me; s2++) region += bound(s2, r + 1) - bound(s2, r);
            dst.delta[r] = int64_t(region) - int64_t(bound(me, r));
        }
        dst.bound[G] = bound(me, G);
        uint32_t *fd = d->flags_dev.as<uint32_t>();
        for (int h = 0; h < H; h++) {
            if (nnz_a && P_local) {
                PeerRows rows;
                std::memset(&rows, 0, sizeof(rows));
                for (int r = 0; r < G; r++) {
                    const uint64_t b0 = block_begin(m, G, r), b1 = block_begin(m, G, r + 1), cut = b0 + (b1 - b0) / 2;
                    rows.lo[r] = H == 1 || h == 0 ? b0 : cut;
                    rows.hi[r] = H == 1 || h == 1 ? b1 : cut;
                }
                LAUNCH(ctx, k_multiply_peer, grid_for(nnz_a / H + 1, 256, unsigned(ctx->sm_count) * 32u), 256, 0, op.a_data, run_off, task_bs,
                       op.b_data, dst, op.a_pos, m_a, rows, (me + 1) % G);
            }
            if (h == H - 1) ev_mul = next_event(ctx);
            // every rank's stores (of this half) have landed once every rank's multiply has retired: one tiny collective
            if (G > 1) NC(ctx, d, d->nccl->AllGather(fd + me, fd, 1, ncclUint32, d->comm, ctx->stream));
            if (side && H == 2) {
                ev_half[h] = ctx->ev_half[h];
                CU(ctx, cudaEventRecord(ev_half[h], ctx->stream));
            }
        }
        ev_xchg = next_event(ctx);
    } else {
        rc = launch_multiply(ctx, TaskSrcSoA{op.a_data, run_off, task_bs}, 0, nnz_a, P_local, op.b_data, ctx->bins.as<Elem>(), 0);
        if (rc) return rc;
        ev_mul = next_event(ctx);
        NC(ctx, d, d->nccl->GroupStart());
        for (int r = 0; r < G; r++) {
            const uint64_t cnt = bound(me, r + 1) - bound(me, r);
            if (cnt) NC(ctx, d, d->nccl->Send(ctx->bins.as<Elem>() + bound(me, r), cnt * 8, ncclUint8, r, d->comm, ctx->stream));
            if (recv_cnt[r]) NC(ctx, d, d->nccl->Recv(d->recv_buf.as<Elem>() + recv_off[r], recv_cnt[r] * 8, ncclUint8, r, d->comm, ctx->stream));
        }
        NC(ctx, d, d->nccl->GroupEnd());
        ev_xchg = next_event(ctx);
    }
    // From here on there is no collective left: a rank that fails below returns on its own.

    // ---- regroup source-major -> row-major, merge ---------------------------------------------------------------------------
    osp_result *res = new osp_result();
    res->ctx = ctx;
    std::memset(&res->stats, 0, sizeof(res->stats));
    ResultGuard guard{res};                      // every early return below (LAUNCH / CU included) frees the result
    auto bail = [&](int code) { ctx->stream = main_stream; return code; };
    MergeJob job;
    job.rows = std::max<uint64_t>(RL, 1); job.idx_range = cols_b; job.long_thresh = plan_long_thresh(cols_b);
    uint64_t nnz_c = 0;
    int n_blocks = 1;
    if (RL) {
        // the plan's scalars (second stream): sizes of C and of the merge scratch
        if (side) ctx->stream = ctx->stream2;
        rc = sync_scalars(ctx);
        if (rc) return bail(rc);
        job.n_tiles = ctx->h_sc->n_tiles; job.n_long = ctx->h_sc->n_long; job.n_xl = ctx->h_sc->n_xl;
        const uint32_t cut_tile = ctx->h_sc->cut_tile;
        const bool halves = H == 2 && cut_local > 0 && cut_local < RL && cut_tile > 0 && cut_tile < job.n_tiles;
        const uint64_t cap = std::max<uint64_t>(ctx->h_sc->cap_bound, 1);
        cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&res->d_pos), (RL + 1) * 8, ctx->stream);
        if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void **>(&res->d_data), cap * 8, ctx->stream);
        if (e != cudaSuccess) {
            cudaGetLastError();
            bail(0);
            return fail(ctx, OSP_ERR_OOM, std::string("result allocation: ") + cudaGetErrorString(e));
        }
        job.c_pos = res->d_pos; job.c_data = res->d_data; job.c_cap = cap;
        unsigned int xl_ctas = 0;
        rc = reserve_merge(ctx, job, xl_ctas);
        if (rc) return bail(rc);
        // owner side, on the second stream when there is one: half h as soon as its partial products have landed
        n_blocks = halves ? 2 : 1;
        for (int h = 0; h < n_blocks; h++) {
            const uint64_t r_lo = halves && h == 1 ? cut_local : 0, r_hi = halves && h == 0 ? cut_local : RL;
            const uint32_t t_lo = halves && h == 1 ? cut_tile : 0u, t_hi = halves && h == 0 ? cut_tile : job.n_tiles;
            rc = [&]() -> int {
                if (side) {
                    // (one phase: everything; two phases whose plan could not be cut: wait for the second half as well)
                    cudaEvent_t need = H == 2 ? ev_half[halves ? h : 1] : nullptr;
                    if (need) CU(ctx, cudaStreamWaitEvent(ctx->stream, need, 0));
                    else { CU(ctx, cudaEventRecord(ctx->ev_half[0], main_stream)); CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_half[0], 0)); }
                }
                LAUNCH(ctx, k_regroup, grid_for(r_hi - r_lo, 8, unsigned(ctx->sm_count) * 32u), 256, 0, d->recv_buf.as<Elem>(),
                       d->src_off.as<uint64_t>(), d->dst_off.as<uint64_t>(), lens, RL, uint32_t(G), d->bins2.as<Elem>(), r_lo, r_hi);
                // the first half is merged while the peers' second multiply runs.  Three resident chain CTAs per SM hold every
                // register of the SM; capping them (OSP_DIST_CHAIN_OCC) lets that multiply in earlier but slows the merge by as
                // much: both kernels are bound by the SMs at G = 4 (profiles/r02_dist.md)
                MergeJob jb = job;
                if (halves && h == 0 && side) {
                    const char *env = std::getenv("OSP_DIST_CHAIN_OCC");
                    jb.chain_ctas_per_sm = env ? std::atoi(env) : 0;      // measured at G = 4: 3.92 ms with 3 or 2, 4.78 ms with 1 -- no cap
                }
                return launch_merge(ctx, jb, xl_ctas, d->bins2.as<Elem>(), 0, t_lo, t_hi, r_lo, r_hi, unsigned(h));
            }();
            if (rc) return bail(rc);
        }
        if (side) {
            rc = [&]() -> int {
                CU(ctx, cudaEventRecord(ctx->ev_join, ctx->stream2));
                CU(ctx, cudaStreamWaitEvent(main_stream, ctx->ev_join, 0));
                return OSP_OK;
            }();
            if (rc) return bail(rc);
        }
        ctx->stream = main_stream;
    } else {
        cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&res->d_pos), 8, ctx->stream);
        if (e != cudaSuccess) { cudaGetLastError(); bail(0); return fail(ctx, OSP_ERR_OOM, "result allocation"); }
        cudaMemsetAsync(res->d_pos, 0, 8, ctx->stream);
    }
    cudaEvent_t ev_end = next_event(ctx);
    rc = sync_scalars(ctx);
    if (rc) return bail(rc);
    if (ctx->h_sc->err) return bail(fail(ctx, OSP_ERR_CUDA, "osp_dist_spgemm: internal capacity check failed on the device"));
    if (RL) nnz_c = ctx->h_sc->nnz_c[n_blocks & 1];
    res->rows = RL;
    res->nnz = nnz_c;
    osp_stats &stt = res->stats;
    stt.rows_c = RL; stt.cols_b = cols_b; stt.n_k = n_k;
    stt.nnz_a = nnz_a; stt.nnz_b = op.nnz_b; stt.nnz_c = nnz_c; stt.products = P_local;
    // this rank's share of the algorithmic bytes: its multiplies (8P_local out) + its merge (8P_owned in, C out)
    stt.algorithmic_bytes = 8 * P_local + 8 * P_owned + 8 * nnz_c + 24 * nnz_a + 8 * op.nnz_b + 8 * (2 * RL + 3 * n_k + 5);
    stt.merge_tiles = job.n_tiles; stt.rows_medium = job.n_long; stt.rows_long = job.n_xl;
    stt.kernel_launches = ctx->launches;
    stt.row_chunks = 1;
    stt.ms_h2d = op.ms_h2d;
    res->call_id = ctx->call_id;
    res->spans.push_back({&stt.ms_total, nullptr, ev_begin, ev_end});
    res->spans.push_back({&stt.ms_convert, nullptr, ev_begin, ev_sym});
    res->spans.push_back({&stt.ms_multiply, nullptr, ev_sym, ev_mul});
    res->spans.push_back({&stt.ms_exchange, nullptr, ev_mul, ev_xchg});
    res->spans.push_back({&stt.ms_merge, nullptr, ev_xchg, ev_end});
    stt.exchange_bytes_out = 8 * (P_local - (bound(me, me + 1) - bound(me, me)));
    for (const auto &mk : ctx->marks) res->spans.push_back({nullptr, mk.name, mk.e0, mk.e1});
    ctx->profile_kernels = false;
    guard.r = nullptr;
    *out = res;
    return OSP_OK;
}

}  // extern "C"
