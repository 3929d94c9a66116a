// osp_engine.cu -- C ABI (include/osp_b200.h) and host orchestration of the sm_100a kernels.
//
// There is no CPU fallback in this file: every entry point that computes needs a CUDA device and
// fails with OSP_ERR_NO_DEVICE / OSP_ERR_CUDA otherwise.
#include "../../include/osp_b200.h"
#include "osp_kernels.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace osp;

namespace {

thread_local std::string g_last_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { cudaGetLastError(); e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want; else p = nullptr;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

}  // namespace

struct osp_ctx {
    int device = 0;
    int sm_count = 148;
    size_t total_mem = 0;
    cudaStream_t stream = nullptr;
    uint64_t ws_limit = 0;          // bytes of partial-product bins per row block
    uint64_t launches = 0;
    std::string err;
    DevScalars *d_sc = nullptr;
    DevScalars *h_sc = nullptr;     // pinned mirror
    // operand staging (host-pointer calls) and converted operands
    DevBuf op_a_pos, op_a_data, op_b_pos, op_b_data, conv_pos, conv_data;
    // symbolic / conversion scratch
    DevBuf run_off, row_bin, col_cnt, col_ptr, tasks, scan_state, uniq, xl_list, xl_acc, xl_bits;
    DevBuf bins;
    std::vector<cudaEvent_t> events;
    size_t events_used = 0;
    // OSP_PROFILE_KERNELS: one event pair per launch
    bool profile_kernels = false;
    struct KernelMark { const char *name; cudaEvent_t e0, e1; };
    std::vector<KernelMark> marks;
};

struct osp_result {
    osp_ctx *ctx = nullptr;
    uint64_t *d_pos = nullptr;
    Elem *d_data = nullptr;
    size_t data_cap = 0;   // elements
    uint64_t rows = 0, nnz = 0;
    osp_stats stats;
    std::vector<std::pair<const char *, float>> kernel_ms;   // OSP_PROFILE_KERNELS
};

namespace {

int fail(osp_ctx *ctx, int code, const std::string &msg) {
    g_last_error = msg;
    if (ctx) ctx->err = msg;
    return code;
}

cudaEvent_t next_event(osp_ctx *ctx) {
    if (ctx->events_used == ctx->events.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ctx->events.push_back(e);
    }
    cudaEvent_t e = ctx->events[ctx->events_used++];
    cudaEventRecord(e, ctx->stream);
    return e;
}

#define CU(ctx, expr)                                                                           \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            cudaGetLastError();                                                                 \
            return fail(ctx, _e == cudaErrorMemoryAllocation ? OSP_ERR_OOM : OSP_ERR_CUDA,      \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));                    \
        }                                                                                       \
    } while (0)

#define LAUNCH(ctx, kernel, grid, block, smem, ...)                                             \
    do {                                                                                        \
        cudaEvent_t _m0 = nullptr;                                                              \
        if ((ctx)->profile_kernels) _m0 = next_event(ctx);                                      \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                        \
        (ctx)->launches++;                                                                      \
        CU(ctx, cudaGetLastError());                                                            \
        if (_m0) (ctx)->marks.push_back({#kernel, _m0, next_event(ctx)});                       \
    } while (0)

constexpr uint32_t MERGE_CAP = 4096;
constexpr size_t MERGE_SMEM = size_t(MERGE_CAP) * 12 + 34 * 4;

unsigned int grid_for(uint64_t items, unsigned int per_block, unsigned int max_blocks) {
    uint64_t b = (items + per_block - 1) / per_block;
    if (b < 1) b = 1;
    return unsigned(std::min<uint64_t>(b, max_blocks));
}

int sync_scalars(osp_ctx *ctx) {
    CU(ctx, cudaMemcpyAsync(ctx->h_sc, ctx->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return OSP_OK;
}

int reset_scalars(osp_ctx *ctx) {
    CU(ctx, cudaMemsetAsync(ctx->d_sc, 0, sizeof(DevScalars), ctx->stream));
    return OSP_OK;
}

template <class In, class Out>
int run_scan(osp_ctx *ctx, In in, Out out, uint64_t n) {
    // n >= 1
    uint64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    CU(ctx, ctx->scan_state.reserve(tiles * 8));
    CU(ctx, cudaMemsetAsync(ctx->scan_state.p, 0, tiles * 8, ctx->stream));
    CU(ctx, cudaMemsetAsync(&ctx->d_sc->tile_counter, 0, 4, ctx->stream));
    LAUNCH(ctx, (k_scan<In, Out>), unsigned(tiles), SCAN_BLOCK, 0, in, out, n, ctx->scan_state.as<uint64_t>(),
           &ctx->d_sc->tile_counter);
    return OSP_OK;
}

// Sorts every bucket [pos[i], pos[i+1]) of `data` by idx (stable w.r.t. nothing: keys are unique
// unless the operand held duplicates) and folds equal idx.  uniq[i] = surviving entries.
int sort_buckets(osp_ctx *ctx, const uint64_t *d_pos, Elem *d_data, uint64_t n_buckets, uint64_t idx_range,
                 uint64_t bin_base, uint64_t *n_long) {
    CU(ctx, ctx->uniq.reserve(std::max<uint64_t>(n_buckets, 1) * 4));
    CU(ctx, ctx->xl_list.reserve(std::max<uint64_t>(n_buckets, 1) * 4));
    CU(ctx, cudaMemsetAsync(&ctx->d_sc->xl_count, 0, 4, ctx->stream));
    unsigned int grid = grid_for(n_buckets, 1, unsigned(ctx->sm_count) * 64u);
    LAUNCH(ctx, k_merge_cta, grid, 256, MERGE_SMEM, d_pos, bin_base, d_data, ctx->uniq.as<uint32_t>(), n_buckets,
           MERGE_CAP, ctx->xl_list.as<uint32_t>(), ctx->d_sc);
    int rc = sync_scalars(ctx);
    if (rc) return rc;
    uint32_t n_xl = ctx->h_sc->xl_count;
    if (n_long) *n_long = n_xl;
    if (n_xl) {
        uint64_t words = (idx_range + 31) / 32;
        uint64_t per_cta = idx_range * 4 + words * 4;
        uint64_t budget = std::max<uint64_t>(ctx->total_mem / 16, 1ull << 28);
        uint64_t max_ctas = budget / std::max<uint64_t>(per_cta, 1);
        if (max_ctas < 1) return fail(ctx, OSP_ERR_UNSUPPORTED, "long-row accumulator does not fit: column range too large");
        unsigned int ctas = unsigned(std::min<uint64_t>({max_ctas, uint64_t(ctx->sm_count) * 2, uint64_t(n_xl)}));
        CU(ctx, ctx->xl_acc.reserve(uint64_t(ctas) * idx_range * 4));
        size_t old_bits_cap = ctx->xl_bits.cap;
        CU(ctx, ctx->xl_bits.reserve(uint64_t(ctas) * words * 4));
        if (ctx->xl_bits.cap != old_bits_cap || true)   // the kernel leaves bits cleared, but layouts change with idx_range
            CU(ctx, cudaMemsetAsync(ctx->xl_bits.p, 0, uint64_t(ctas) * words * 4, ctx->stream));
        LAUNCH(ctx, k_merge_xl, ctas, 256, MERGE_SMEM, d_pos, bin_base, d_data, ctx->uniq.as<uint32_t>(),
               ctx->xl_list.as<uint32_t>(), ctx->d_sc, MERGE_CAP, ctx->xl_acc.as<float>(), ctx->xl_bits.as<uint32_t>(),
               idx_range);
    }
    return OSP_OK;
}

// Stable transposition on the device (all pointers are device pointers).
int csr2csc_device(osp_ctx *ctx, uint64_t n_major, uint64_t n_minor, const uint64_t *d_pos, const Elem *d_data,
                   uint64_t nnz, uint64_t *d_pos_out, Elem *d_data_out) {
    if (nnz >= (1ull << 32)) return fail(ctx, OSP_ERR_UNSUPPORTED, "operands with >= 2^32 non-zeros are not supported");
    if (n_minor == 0) {
        CU(ctx, cudaMemsetAsync(d_pos_out, 0, 8, ctx->stream));
        return OSP_OK;
    }
    CU(ctx, ctx->col_cnt.reserve(n_minor * 4));
    uint32_t *cnt = ctx->col_cnt.as<uint32_t>();
    CU(ctx, cudaMemsetAsync(cnt, 0, n_minor * 4, ctx->stream));
    unsigned int g = grid_for(nnz, 256, unsigned(ctx->sm_count) * 16u);
    if (nnz) LAUNCH(ctx, k_hist_elems, g, 256, 0, d_data, nnz, n_minor, cnt, ctx->d_sc);
    int rc = run_scan(ctx, U32In{cnt}, U64Out{d_pos_out, 0}, n_minor);
    if (rc) return rc;
    if (!nnz) return OSP_OK;
    CU(ctx, cudaMemsetAsync(cnt, 0, n_minor * 4, ctx->stream));
    unsigned int gw = grid_for(n_major, 8, unsigned(ctx->sm_count) * 16u);
    LAUNCH(ctx, k_scatter_elems, gw, 256, 0, d_pos, d_data, n_major, n_minor, d_pos_out, cnt, d_data_out, ctx->d_sc);
    rc = sort_buckets(ctx, d_pos_out, d_data_out, n_minor, std::max<uint64_t>(n_major, 1), 0, nullptr);
    if (rc) return rc;
    LAUNCH(ctx, k_check_full, grid_for(n_minor, 256, 1u << 30), 256, 0, d_pos_out, ctx->uniq.as<uint32_t>(), n_minor,
           ctx->d_sc);
    rc = sync_scalars(ctx);
    if (rc) return rc;
    if (ctx->h_sc->err == 233) return fail(ctx, OSP_ERR_DUPLICATE, "duplicate (row,col) entry in operand");
    if (ctx->h_sc->err) return fail(ctx, OSP_ERR_INDEX, "index out of range in operand");
    return OSP_OK;
}

template <class Src>
int launch_multiply(osp_ctx *ctx, Src src, uint64_t t0, uint64_t t1, uint64_t products, const uint64_t *b_pos,
                    const Elem *b_data, Elem *bins, uint64_t bin_base) {
    uint64_t n = t1 - t0;
    if (!n || !products) return OSP_OK;
    double avg = double(products) / double(n);
    int G = avg <= 4.0 ? 4 : avg <= 8.0 ? 8 : avg <= 16.0 ? 16 : 32;
    unsigned int grid = grid_for(n, 256, unsigned(ctx->sm_count) * 32u);
    switch (G) {
        case 4: LAUNCH(ctx, (k_multiply<4, Src>), grid, 256, 0, src, t0, t1, b_pos, b_data, bins, bin_base); break;
        case 8: LAUNCH(ctx, (k_multiply<8, Src>), grid, 256, 0, src, t0, t1, b_pos, b_data, bins, bin_base); break;
        case 16: LAUNCH(ctx, (k_multiply<16, Src>), grid, 256, 0, src, t0, t1, b_pos, b_data, bins, bin_base); break;
        default: LAUNCH(ctx, (k_multiply<32, Src>), grid, 256, 0, src, t0, t1, b_pos, b_data, bins, bin_base); break;
    }
    return OSP_OK;
}

int grow_result(osp_result *r, uint64_t need_elems, uint64_t hint_elems) {
    osp_ctx *ctx = r->ctx;
    if (need_elems <= r->data_cap) return OSP_OK;
    uint64_t want = std::max<uint64_t>({need_elems, hint_elems, r->data_cap + r->data_cap / 2, 1});
    Elem *nd = nullptr;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&nd), want * sizeof(Elem), ctx->stream);
    if (e != cudaSuccess && want > need_elems) {
        cudaGetLastError();
        want = need_elems;
        e = cudaMallocAsync(reinterpret_cast<void **>(&nd), want * sizeof(Elem), ctx->stream);
    }
    CU(ctx, e);
    if (r->d_data) {
        if (r->nnz) CU(ctx, cudaMemcpyAsync(nd, r->d_data, r->nnz * sizeof(Elem), cudaMemcpyDeviceToDevice, ctx->stream));
        CU(ctx, cudaFreeAsync(r->d_data, ctx->stream));
    }
    r->d_data = nd;
    r->data_cap = want;
    return OSP_OK;
}

}  // namespace

// ========================================================================================
// C ABI
// ========================================================================================
extern "C" {

const char *osp_version(void) { return "outerspace_b200 0.1 (sm_100a)"; }

int osp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *osp_last_error(const osp_ctx *ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int osp_create(int device, osp_ctx **out) {
    if (!out) return fail(nullptr, OSP_ERR_INVALID, "osp_create: out is NULL");
    *out = nullptr;
    int n = osp_device_count();
    if (n <= 0) return fail(nullptr, OSP_ERR_NO_DEVICE, "no CUDA device visible: this engine has no CPU fallback");
    if (device < 0 || device >= n) return fail(nullptr, OSP_ERR_INVALID, "osp_create: device index out of range");
    osp_ctx *ctx = new osp_ctx();
    ctx->device = device;
    CU(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(nullptr, cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    ctx->total_mem = prop.totalGlobalMem;
    CU(nullptr, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CU(nullptr, cudaMalloc(reinterpret_cast<void **>(&ctx->d_sc), sizeof(DevScalars)));
    CU(nullptr, cudaMallocHost(reinterpret_cast<void **>(&ctx->h_sc), sizeof(DevScalars)));
    CU(nullptr, cudaFuncSetAttribute(k_merge_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, int(MERGE_SMEM)));
    CU(nullptr, cudaFuncSetAttribute(k_merge_xl, cudaFuncAttributeMaxDynamicSharedMemorySize, int(MERGE_SMEM)));
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    ctx->ws_limit = uint64_t(double(ctx->total_mem) * 0.35);
    if (const char *env = std::getenv("OSP_WORKSPACE_LIMIT_MB")) {
        uint64_t mb = std::strtoull(env, nullptr, 10);
        if (mb) ctx->ws_limit = mb << 20;
    }
    *out = ctx;
    return OSP_OK;
}

void osp_destroy(osp_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (DevBuf *b : {&ctx->op_a_pos, &ctx->op_a_data, &ctx->op_b_pos, &ctx->op_b_data, &ctx->conv_pos, &ctx->conv_data,
                      &ctx->run_off, &ctx->row_bin, &ctx->col_cnt, &ctx->col_ptr, &ctx->tasks, &ctx->scan_state,
                      &ctx->uniq, &ctx->xl_list, &ctx->xl_acc, &ctx->xl_bits, &ctx->bins})
        b->release();
    for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
    if (ctx->d_sc) cudaFree(ctx->d_sc);
    if (ctx->h_sc) cudaFreeHost(ctx->h_sc);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int osp_set_workspace_limit(osp_ctx *ctx, uint64_t bytes) {
    if (!ctx || bytes < 4096) return fail(ctx, OSP_ERR_INVALID, "osp_set_workspace_limit: bad argument");
    ctx->ws_limit = bytes;
    return OSP_OK;
}

void *osp_stream(osp_ctx *ctx) { return ctx ? static_cast<void *>(ctx->stream) : nullptr; }

int osp_spgemm(osp_ctx *ctx, const osp_spgemm_args *args, osp_result **out) {
    if (!ctx || !args || !out) return fail(ctx, OSP_ERR_INVALID, "osp_spgemm: NULL argument");
    *out = nullptr;
    if (!args->a_pos || !args->b_pos) return fail(ctx, OSP_ERR_INVALID, "osp_spgemm: NULL pos array");
    const bool a_is_csr = args->flags & OSP_A_IS_CSR;
    const bool on_device = args->flags & OSP_DEVICE_POINTERS;
    const bool rowwise = args->flags & OSP_ROWWISE_ORDER;
    const bool profile = args->flags & OSP_PROFILE_PHASES;
    // k-dimension check: lmat.NRow() == rmat.NRow(), SimOuterSPACE.cpp:47
    if (!a_is_csr && args->a_slices != args->n_k)
        return fail(ctx, OSP_ERR_INVALID, "osp_spgemm: CSC(A) and CSR(B) must have the same number of slices");
    if (args->n_k >= (1ull << 32) || args->a_slices >= (1ull << 32) || args->rows_c >= (1ull << 32) ||
        args->cols_b >= (1ull << 32))
        return fail(ctx, OSP_ERR_INVALID, "osp_spgemm: dimensions must fit index_t (uint32)");
    CU(ctx, cudaSetDevice(ctx->device));
    ctx->launches = 0;
    ctx->events_used = 0;
    ctx->marks.clear();
    ctx->profile_kernels = args->flags & OSP_PROFILE_KERNELS;
    int rc = reset_scalars(ctx);
    if (rc) return rc;

    const uint64_t n_k = args->n_k;
    uint64_t nnz_a = 0, nnz_b = 0;
    const uint64_t *dA_pos, *dB_pos;
    const Elem *dA_data, *dB_data;
    float ms_h2d = 0.f;
    if (on_device) {
        dA_pos = args->a_pos; dB_pos = args->b_pos;
        dA_data = static_cast<const Elem *>(args->a_data);
        dB_data = static_cast<const Elem *>(args->b_data);
        CU(ctx, cudaMemcpyAsync(&nnz_a, dA_pos + args->a_slices, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaMemcpyAsync(&nnz_b, dB_pos + n_k, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        nnz_a = args->a_pos[args->a_slices];
        nnz_b = args->b_pos[n_k];
        CU(ctx, ctx->op_a_pos.reserve((args->a_slices + 1) * 8));
        CU(ctx, ctx->op_b_pos.reserve((n_k + 1) * 8));
        CU(ctx, ctx->op_a_data.reserve(std::max<uint64_t>(nnz_a, 1) * 8));
        CU(ctx, ctx->op_b_data.reserve(std::max<uint64_t>(nnz_b, 1) * 8));
        cudaEvent_t e0 = next_event(ctx);
        CU(ctx, cudaMemcpyAsync(ctx->op_a_pos.p, args->a_pos, (args->a_slices + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(ctx->op_b_pos.p, args->b_pos, (n_k + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (nnz_a) CU(ctx, cudaMemcpyAsync(ctx->op_a_data.p, args->a_data, nnz_a * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (nnz_b) CU(ctx, cudaMemcpyAsync(ctx->op_b_data.p, args->b_data, nnz_b * 8, cudaMemcpyHostToDevice, ctx->stream));
        cudaEvent_t e1 = next_event(ctx);
        CU(ctx, cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms_h2d, e0, e1);
        dA_pos = ctx->op_a_pos.as<uint64_t>(); dB_pos = ctx->op_b_pos.as<uint64_t>();
        dA_data = ctx->op_a_data.as<Elem>(); dB_data = ctx->op_b_data.as<Elem>();
    }
    if ((nnz_a && !args->a_data) || (nnz_b && !args->b_data))
        return fail(ctx, OSP_ERR_INVALID, "osp_spgemm: NULL data array");
    if (nnz_a >= (1ull << 32) || nnz_b >= (1ull << 32))
        return fail(ctx, OSP_ERR_UNSUPPORTED, "operands with >= 2^32 non-zeros are not supported");

    cudaEvent_t ev_begin = next_event(ctx);

    // ---- bring A into row-compressed form (needed for k-ordered bin offsets) ---------------------
    uint64_t m_a = args->a_slices;
    if (!a_is_csr) {
        // rows of A = max row id + 1 (SimOuterSPACE.cpp:49-53)
        CU(ctx, cudaMemsetAsync(&ctx->d_sc->max_idx, 0, 4, ctx->stream));
        if (nnz_a) LAUNCH(ctx, k_max_idx, grid_for(nnz_a, 1024, unsigned(ctx->sm_count) * 8u), 256, 0, dA_data, nnz_a, ctx->d_sc);
        rc = sync_scalars(ctx);
        if (rc) return rc;
        m_a = uint64_t(ctx->h_sc->max_idx) + 1;
        CU(ctx, ctx->conv_pos.reserve((m_a + 1) * 8));
        CU(ctx, ctx->conv_data.reserve(std::max<uint64_t>(nnz_a, 1) * 8));
        rc = csr2csc_device(ctx, n_k, m_a, dA_pos, dA_data, nnz_a, ctx->conv_pos.as<uint64_t>(), ctx->conv_data.as<Elem>());
        if (rc) return rc;
        dA_pos = ctx->conv_pos.as<uint64_t>();
        dA_data = ctx->conv_data.as<Elem>();
    }

    // ---- dimensions of C -------------------------------------------------------------------------
    uint64_t cols_b = args->cols_b;
    if (!cols_b) {
        CU(ctx, cudaMemsetAsync(&ctx->d_sc->max_idx, 0, 4, ctx->stream));
        if (nnz_b) LAUNCH(ctx, k_max_idx, grid_for(nnz_b, 1024, unsigned(ctx->sm_count) * 8u), 256, 0, dB_data, nnz_b, ctx->d_sc);
    }
    LAUNCH(ctx, k_last_nonempty, 1, 1, 0, dA_pos, m_a, ctx->d_sc);

    // ---- symbolic pass: run offsets of every non-zero of A, P --------------------------------------
    CU(ctx, ctx->run_off.reserve((nnz_a + 1) * 8));
    uint64_t *run_off = ctx->run_off.as<uint64_t>();
    if (nnz_a) {
        rc = run_scan(ctx, RunLenIn{dA_data, dB_pos, n_k, ctx->d_sc}, RunOffOut{run_off, ctx->d_sc, nnz_a}, nnz_a);
        if (rc) return rc;
    } else {
        CU(ctx, cudaMemsetAsync(run_off, 0, 8, ctx->stream));
    }
    rc = sync_scalars(ctx);
    if (rc) return rc;
    if (ctx->h_sc->err) return fail(ctx, OSP_ERR_INDEX, "osp_spgemm: index of A out of range of the inner dimension");
    const uint64_t P = ctx->h_sc->products;
    if (!cols_b) cols_b = uint64_t(ctx->h_sc->max_idx) + 1;
    const uint64_t min_rows = std::max<uint64_t>(ctx->h_sc->last_nonempty, 1);
    uint64_t rows_c = args->rows_c ? args->rows_c : min_rows;
    if (rows_c < ctx->h_sc->last_nonempty)
        return fail(ctx, OSP_ERR_INDEX, "osp_spgemm: rows_c is smaller than the largest row id of A + 1");

    CU(ctx, ctx->row_bin.reserve((rows_c + 1) * 8));
    uint64_t *row_bin = ctx->row_bin.as<uint64_t>();
    LAUNCH(ctx, k_row_bins, grid_for(rows_c + 1, 256, 1u << 30), 256, 0, dA_pos, m_a, run_off, nnz_a, rows_c, row_bin);

    // ---- result object ------------------------------------------------------------------------------
    osp_result *res = new osp_result();
    res->ctx = ctx;
    res->rows = rows_c;
    std::memset(&res->stats, 0, sizeof(res->stats));
    {
        cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&res->d_pos), (rows_c + 1) * 8, ctx->stream);
        if (e != cudaSuccess) { delete res; CU(ctx, e); }
    }
    auto bail = [&](int code) { osp_result_free(res); return code; };

    // ---- row blocks ------------------------------------------------------------------------------------
    const uint64_t limit_elems = std::max<uint64_t>(ctx->ws_limit / 8, 1);
    std::vector<uint64_t> block_rows;   // boundaries r0 < r1 < ...
    std::vector<uint64_t> h_row_bin, h_a_pos;
    block_rows.push_back(0);
    if (P <= limit_elems) {
        block_rows.push_back(rows_c);
    } else {
        h_row_bin.resize(rows_c + 1);
        CU(ctx, cudaMemcpyAsync(h_row_bin.data(), row_bin, (rows_c + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        uint64_t r0 = 0;
        while (r0 < rows_c) {
            uint64_t target = h_row_bin[r0] + limit_elems;
            uint64_t r1 = std::upper_bound(h_row_bin.begin() + r0, h_row_bin.end(), target) - h_row_bin.begin() - 1;
            if (r1 <= r0) r1 = r0 + 1;
            if (r1 > rows_c) r1 = rows_c;
            block_rows.push_back(r1);
            r0 = r1;
        }
    }
    const size_t n_blocks = block_rows.size() - 1;
    if (n_blocks > 1) {
        h_a_pos.resize(m_a + 1);
        CU(ctx, cudaMemcpyAsync(h_a_pos.data(), dA_pos, (m_a + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }

    cudaEvent_t ev_sym = next_event(ctx);
    std::vector<cudaEvent_t> ev_blocks;   // 4 per block: start, after convert, after multiply, after merge
    uint64_t nnz_c = 0, rows_long = 0;

    for (size_t b = 0; b < n_blocks; b++) {
        const uint64_t r0 = block_rows[b], r1 = block_rows[b + 1], rows = r1 - r0;
        uint64_t e0, e1, bin0, bin1;
        if (n_blocks == 1) {
            e0 = 0; e1 = nnz_a; bin0 = 0; bin1 = P;
        } else {
            e0 = h_a_pos[std::min(r0, m_a)]; e1 = h_a_pos[std::min(r1, m_a)];
            bin0 = h_row_bin[r0]; bin1 = h_row_bin[r1];
        }
        const uint64_t p_block = bin1 - bin0;
        ev_blocks.push_back(next_event(ctx));
        CU(ctx, ctx->bins.reserve(std::max<uint64_t>(p_block, 1) * 8));
        Elem *bins = ctx->bins.as<Elem>();

        if (!rowwise && e1 > e0 && p_block) {
            // CSR -> CSC of this block of A: histogram, scan, scatter into the k-ordered task list
            CU(ctx, ctx->col_cnt.reserve(n_k * 4));
            CU(ctx, ctx->col_ptr.reserve((n_k + 1) * 4));
            CU(ctx, ctx->tasks.reserve((e1 - e0) * sizeof(Task)));
            uint32_t *cnt = ctx->col_cnt.as<uint32_t>();
            CU(ctx, cudaMemsetAsync(cnt, 0, n_k * 4, ctx->stream));
            unsigned int g = grid_for(e1 - e0, 256, unsigned(ctx->sm_count) * 16u);
            LAUNCH(ctx, k_col_hist, g, 256, 0, dA_data, e0, e1, cnt);
            rc = run_scan(ctx, U32In{cnt}, U32OutFromU64{ctx->col_ptr.as<uint32_t>()}, n_k);
            if (rc) return bail(rc);
            LAUNCH(ctx, k_scatter_tasks, g, 256, 0, dA_data, run_off, e0, e1, ctx->col_ptr.as<uint32_t>(), cnt,
                   ctx->tasks.as<Task>());
        }
        if (profile) CU(ctx, cudaStreamSynchronize(ctx->stream));
        ev_blocks.push_back(next_event(ctx));

        if (rowwise) rc = launch_multiply(ctx, TaskSrcSoA{dA_data, run_off}, e0, e1, p_block, dB_pos, dB_data, bins, bin0);
        else rc = launch_multiply(ctx, TaskSrcAoS{ctx->tasks.as<Task>()}, 0, e1 - e0, p_block, dB_pos, dB_data, bins, bin0);
        if (rc) return bail(rc);
        if (profile) CU(ctx, cudaStreamSynchronize(ctx->stream));
        ev_blocks.push_back(next_event(ctx));

        uint64_t n_long = 0;
        rc = sort_buckets(ctx, row_bin + r0, bins, rows, cols_b, bin0, &n_long);
        if (rc) return bail(rc);
        rows_long += n_long;
        rc = run_scan(ctx, U32In{ctx->uniq.as<uint32_t>()},
                      U64OutTotal{res->d_pos + r0, nnz_c, rows, ctx->d_sc}, rows);
        if (rc) return bail(rc);
        rc = sync_scalars(ctx);
        if (rc) return bail(rc);
        const uint64_t block_nnz = ctx->h_sc->block_nnz;
        uint64_t hint = n_blocks == 1 ? block_nnz
                                      : uint64_t(double(nnz_c + block_nnz) * double(rows_c) / double(r1) * 1.05);
        rc = grow_result(res, nnz_c + block_nnz, hint);
        if (rc) return bail(rc);
        if (block_nnz)
            LAUNCH(ctx, k_gather_rows, grid_for(rows, 8, unsigned(ctx->sm_count) * 32u), 256, 0, row_bin + r0, bin0, bins,
                   ctx->uniq.as<uint32_t>(), res->d_pos + r0, rows, res->d_data);
        nnz_c += block_nnz;
        res->nnz = nnz_c;
        ev_blocks.push_back(next_event(ctx));
    }
    // U64OutTotal wrote d_pos[r1] = carry + total for the last block, i.e. d_pos[rows_c] = nnz_c.
    cudaEvent_t ev_end = next_event(ctx);
    CU(ctx, cudaStreamSynchronize(ctx->stream));

    osp_stats &st = res->stats;
    st.rows_c = rows_c; st.cols_b = cols_b; st.n_k = n_k;
    st.nnz_a = nnz_a; st.nnz_b = nnz_b; st.nnz_c = nnz_c; st.products = P;
    st.algorithmic_bytes = 16 * P + 8 * nnz_c + 24 * nnz_a + 8 * nnz_b + 8 * (2 * rows_c + 3 * n_k + 5);
    st.rows_long = rows_long;
    st.kernel_launches = ctx->launches;
    st.row_chunks = n_blocks;
    st.ms_h2d = ms_h2d;
    cudaEventElapsedTime(&st.ms_total, ev_begin, ev_end);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev_begin, ev_sym);
    st.ms_convert = ms;
    for (size_t b = 0; b < n_blocks; b++) {
        cudaEventElapsedTime(&ms, ev_blocks[4 * b], ev_blocks[4 * b + 1]); st.ms_convert += ms;
        cudaEventElapsedTime(&ms, ev_blocks[4 * b + 1], ev_blocks[4 * b + 2]); st.ms_multiply += ms;
        cudaEventElapsedTime(&ms, ev_blocks[4 * b + 2], ev_blocks[4 * b + 3]); st.ms_merge += ms;
    }
    for (const auto &m : ctx->marks) {
        cudaEventElapsedTime(&ms, m.e0, m.e1);
        res->kernel_ms.emplace_back(m.name, ms);
    }
    ctx->profile_kernels = false;
    *out = res;
    return OSP_OK;
}

int osp_result_kernels(const osp_result *r, uint64_t *n, const char **names, float *ms) {
    if (!r || !n) return fail(nullptr, OSP_ERR_INVALID, "osp_result_kernels: NULL argument");
    if (names && ms)
        for (size_t i = 0; i < r->kernel_ms.size() && i < *n; i++) { names[i] = r->kernel_ms[i].first; ms[i] = r->kernel_ms[i].second; }
    *n = r->kernel_ms.size();
    return OSP_OK;
}

int osp_result_dims(const osp_result *r, uint64_t *rows, uint64_t *nnz) {
    if (!r) return fail(nullptr, OSP_ERR_INVALID, "osp_result_dims: NULL result");
    if (rows) *rows = r->rows;
    if (nnz) *nnz = r->nnz;
    return OSP_OK;
}

int osp_result_copy(osp_result *r, uint64_t *pos, void *data) {
    if (!r || !pos) return fail(nullptr, OSP_ERR_INVALID, "osp_result_copy: NULL argument");
    osp_ctx *ctx = r->ctx;
    if (r->nnz && !data) return fail(ctx, OSP_ERR_INVALID, "osp_result_copy: NULL data");
    CU(ctx, cudaSetDevice(ctx->device));
    cudaEvent_t e0 = next_event(ctx);
    CU(ctx, cudaMemcpyAsync(pos, r->d_pos, (r->rows + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (r->nnz) CU(ctx, cudaMemcpyAsync(data, r->d_data, r->nnz * 8, cudaMemcpyDeviceToHost, ctx->stream));
    cudaEvent_t e1 = next_event(ctx);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&r->stats.ms_d2h, e0, e1);
    return OSP_OK;
}

int osp_result_device(const osp_result *r, const uint64_t **d_pos, const void **d_data) {
    if (!r) return fail(nullptr, OSP_ERR_INVALID, "osp_result_device: NULL result");
    if (d_pos) *d_pos = r->d_pos;
    if (d_data) *d_data = r->d_data;
    return OSP_OK;
}

int osp_result_stats(const osp_result *r, osp_stats *stats) {
    if (!r || !stats) return fail(nullptr, OSP_ERR_INVALID, "osp_result_stats: NULL argument");
    *stats = r->stats;
    return OSP_OK;
}

void osp_result_free(osp_result *r) {
    if (!r) return;
    osp_ctx *ctx = r->ctx;
    cudaSetDevice(ctx->device);
    if (r->d_pos) cudaFreeAsync(r->d_pos, ctx->stream);
    if (r->d_data) cudaFreeAsync(r->d_data, ctx->stream);
    delete r;
}

// Sizes the reference's timing models read from TaskProvider (SimOuterSPACE.cpp:34-42,59-64), derived
// from the operand structure and the result's row pointer (host arithmetic on sizes only; no values).
int osp_task_sizes(osp_ctx *ctx, const osp_spgemm_args *args, const osp_result *r, uint64_t *n_multiply,
                   uint32_t *multiply_nnzc_nnzr, uint64_t *n_merge, uint32_t *merge_ways_out) {
    if (!ctx || !args || !r || !n_multiply || !n_merge) return fail(ctx, OSP_ERR_INVALID, "osp_task_sizes: NULL argument");
    const bool a_is_csr = args->flags & OSP_A_IS_CSR;
    const bool on_device = args->flags & OSP_DEVICE_POINTERS;
    const uint64_t n_k = args->n_k, a_slices = args->a_slices;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<uint64_t> a_pos(a_slices + 1), b_pos(n_k + 1), c_pos(r->rows + 1);
    cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost;
    CU(ctx, cudaMemcpy(a_pos.data(), args->a_pos, (a_slices + 1) * 8, kind));
    CU(ctx, cudaMemcpy(b_pos.data(), args->b_pos, (n_k + 1) * 8, kind));
    CU(ctx, cudaMemcpy(c_pos.data(), r->d_pos, (r->rows + 1) * 8, cudaMemcpyDeviceToHost));
    const uint64_t nnz_a = a_pos[a_slices];
    std::vector<Elem> a_data(nnz_a);
    if (nnz_a) CU(ctx, cudaMemcpy(a_data.data(), args->a_data, nnz_a * 8, kind));
    std::vector<uint32_t> nnzc(n_k, 0), ways(r->rows, 0);
    if (a_is_csr) {
        for (uint64_t i = 0; i < a_slices; i++)
            for (uint64_t e = a_pos[i]; e < a_pos[i + 1]; e++) {
                uint32_t k = a_data[e].idx;
                if (k >= n_k) return fail(ctx, OSP_ERR_INDEX, "osp_task_sizes: index out of range");
                nnzc[k]++;
                if (b_pos[k + 1] > b_pos[k] && i < r->rows) ways[i]++;
            }
    } else {
        for (uint64_t k = 0; k < n_k; k++) {
            nnzc[k] = uint32_t(a_pos[k + 1] - a_pos[k]);
            if (b_pos[k + 1] > b_pos[k])
                for (uint64_t e = a_pos[k]; e < a_pos[k + 1]; e++)
                    if (a_data[e].idx < r->rows) ways[a_data[e].idx]++;
        }
    }
    uint64_t nm = 0;
    for (uint64_t k = 0; k < n_k; k++) {
        uint64_t nnzr = b_pos[k + 1] - b_pos[k];
        if (!nnzc[k] || !nnzr) continue;                       // skipped slices, SimOuterSPACE.cpp:82-83
        if (multiply_nnzc_nnzr) { multiply_nnzc_nnzr[2 * nm] = nnzc[k]; multiply_nnzc_nnzr[2 * nm + 1] = uint32_t(nnzr); }
        nm++;
    }
    *n_multiply = nm;
    *n_merge = r->rows;                                        // one MergeTask per row, empty rows included (:128)
    if (merge_ways_out)
        for (uint64_t i = 0; i < r->rows; i++) {
            merge_ways_out[2 * i] = ways[i];
            merge_ways_out[2 * i + 1] = uint32_t(c_pos[i + 1] - c_pos[i]);
        }
    return OSP_OK;
}

int osp_csr2csc(osp_ctx *ctx, uint64_t n_major, uint64_t n_minor, const uint64_t *pos, const void *data, uint32_t flags,
                uint64_t *pos_out, void *data_out) {
    if (!ctx || !pos || !pos_out) return fail(ctx, OSP_ERR_INVALID, "osp_csr2csc: NULL argument");
    if (n_major >= (1ull << 32) || n_minor >= (1ull << 32))
        return fail(ctx, OSP_ERR_INVALID, "osp_csr2csc: dimensions must fit index_t (uint32)");
    CU(ctx, cudaSetDevice(ctx->device));
    ctx->launches = 0;
    ctx->events_used = 0;
    ctx->profile_kernels = false;
    int rc = reset_scalars(ctx);
    if (rc) return rc;
    if (flags & OSP_DEVICE_POINTERS) {
        uint64_t nnz = 0;
        CU(ctx, cudaMemcpyAsync(&nnz, pos + n_major, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        return csr2csc_device(ctx, n_major, n_minor, pos, static_cast<const Elem *>(data), nnz, pos_out,
                              static_cast<Elem *>(data_out));
    }
    uint64_t nnz = pos[n_major];
    if (nnz && (!data || !data_out)) return fail(ctx, OSP_ERR_INVALID, "osp_csr2csc: NULL data");
    CU(ctx, ctx->op_a_pos.reserve((n_major + 1) * 8));
    CU(ctx, ctx->op_a_data.reserve(std::max<uint64_t>(nnz, 1) * 8));
    CU(ctx, ctx->conv_pos.reserve((n_minor + 1) * 8));
    CU(ctx, ctx->conv_data.reserve(std::max<uint64_t>(nnz, 1) * 8));
    CU(ctx, cudaMemcpyAsync(ctx->op_a_pos.p, pos, (n_major + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (nnz) CU(ctx, cudaMemcpyAsync(ctx->op_a_data.p, data, nnz * 8, cudaMemcpyHostToDevice, ctx->stream));
    rc = csr2csc_device(ctx, n_major, n_minor, ctx->op_a_pos.as<uint64_t>(), ctx->op_a_data.as<Elem>(), nnz,
                        ctx->conv_pos.as<uint64_t>(), ctx->conv_data.as<Elem>());
    if (rc) return rc;
    CU(ctx, cudaMemcpyAsync(pos_out, ctx->conv_pos.p, (n_minor + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (nnz) CU(ctx, cudaMemcpyAsync(data_out, ctx->conv_data.p, nnz * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return OSP_OK;
}

}  // extern "C"
