// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C-ABI wrapper around the UNMODIFIED reference sources, compiled where they
// lie (-I/root/reference/simulator) into oracle/_ref/libosp_ref.so by
// oracle/Makefile.  No reference source is copied into this repository: the two
// translation units are textually included at build time only.
//
//   ref_a.so part (this file, REF_PART_SPGEMM): SimSpGEMM.cpp gives readcoo,
//       dupcheck, coo2csr<>, cscMulcsr, compareCOO verbatim (its main() is
//       renamed away).  deduplicateCOO sits inside `#if 0` in the reference
//       (SimSpGEMM.cpp:519-535) so it cannot be compiled from there; its fold
//       is applied here on the reference's own cscMulcsr output with
//       std::stable_sort (the canonical k-ordered merge).
//   ref_b.so part (REF_PART_TASKPROVIDER): SimOuterSPACE.cpp with the stub
//       ramulator header gives the as-written TaskProvider (bug-compatible:
//       positional column index at :89, inverted duplicate test at :120).
//
// The two parts are separate shared objects because both reference files
// define overlapping global symbols when put in one TU.

#include <cstdint>
#include <cstring>
#include <list>
#include <tuple>
#include <functional>
#include <string>
#include <sstream>
#include <chrono>

#if defined(REF_PART_SPGEMM)

#define main ref_main_unused
#include "SimSpGEMM.cpp"
#undef main

// satisfies the forward declaration at SimSpGEMM.cpp:816 (never called here)
size_t simulateOuterSPACE(const CSRMatrix &, const CSRMatrix &) { return 1; }

namespace {
struct RefResult {
    std::vector<size_t> pos;
    std::vector<CSRElement> data;
    uint64_t products = 0;
    double seconds = 0;
};
struct RefMtx {
    size_t nrow = 0, ncol = 0;
    COOMatrix coo;
};
CSRMatrix wrap(uint64_t n, const uint64_t *pos, const void *data) {
    CSRMatrix m;
    m.pos.assign(pos, pos + n + 1);
    const CSRElement *d = static_cast<const CSRElement *>(data);
    m.data.assign(d, d + pos[n]);
    return m;
}
}  // namespace

extern "C" {

void *ref_mtx_open(const char *path, int symmetric) {
    std::ifstream fin(path);
    if (!fin) return nullptr;
    auto *m = new RefMtx();
    m->coo = readcoo(fin, m->nrow, m->ncol, symmetric != 0);
    return m;
}
void ref_mtx_dims(void *h, uint64_t *nrow, uint64_t *ncol, uint64_t *nnz) {
    auto *m = static_cast<RefMtx *>(h);
    *nrow = m->nrow; *ncol = m->ncol; *nnz = m->coo.size();
}
void ref_mtx_copy(void *h, uint32_t *rows, uint32_t *cols, float *vals) {
    auto *m = static_cast<RefMtx *>(h);
    for (size_t i = 0; i < m->coo.size(); i++) {
        rows[i] = m->coo[i].row; cols[i] = m->coo[i].col; vals[i] = m->coo[i].val;
    }
}
void ref_mtx_free(void *h) { delete static_cast<RefMtx *>(h); }

// reference coo2csr<transpose> (SimSpGEMM.cpp:102-152); 233 on duplicates.
int ref_coo2csr(uint64_t nnz, const uint32_t *rows, const uint32_t *cols, const float *vals,
                uint64_t N, int transpose, uint64_t *pos, void *data_out) {
    COOMatrix coo(nnz);
    for (uint64_t i = 0; i < nnz; i++) coo[i] = COOElement{rows[i], cols[i], vals[i]};
    CSRMatrix r;
    try {
        r = transpose ? coo2csr<true>(coo, N) : coo2csr<false>(coo, N);
    } catch (int code) {
        return code;
    }
    std::memcpy(pos, r.pos.data(), (N + 1) * sizeof(size_t));
    if (nnz) std::memcpy(data_out, r.data.data(), nnz * sizeof(CSRElement));
    return 0;
}

// reference csr2compact (SimSpGEMM.cpp:154-219) / csc2rawcompact (:221-243), unmodified.  First call with
// group_pos == nullptr returns the number of groups + 1 (pos.size()).
uint64_t ref_compact(int raw, uint64_t n, const uint64_t *pos, const void *data, uint64_t *group_pos,
                     uint32_t *rows, uint32_t *cols, float *vals) {
    CSRMatrix m = wrap(n, pos, data);
    // csr2compact reports its run time on std::cout (TIMER, SimSpGEMM.cpp:38,162): muted for the call -- the stream of
    // a dlopen'ed library built with a newer libstdc++ than the host process may not have its locale set up
    const std::ios_base::iostate saved = std::cout.rdstate();
    std::cout.setstate(std::ios_base::failbit);
    CompactCOOMatrix c = raw ? csc2rawcompact(m) : csr2compact(m);
    std::cout.clear(saved);
    if (!group_pos) return c.pos.size();
    for (size_t i = 0; i < c.pos.size(); i++) group_pos[i] = c.pos[i];
    for (size_t i = 0; i < c.data.size(); i++) { rows[i] = c.data[i].row; cols[i] = c.data[i].col; vals[i] = c.data[i].val; }
    return c.pos.size();
}

// reference compactMulcsr (SimSpGEMM.cpp:247-263) on a compact operand given as group_pos + triplets: the partial
// products of all groups in group order, then the deduplicateCOO fold with a stable sort (a row's entries appear in
// ascending k over the groups, so this is the k-ordered merge).  233 when its dupcheck throws.
void *ref_compact_spgemm(uint64_t n_groups, const uint64_t *group_pos, const uint32_t *rows, const uint32_t *cols,
                         const float *vals, uint64_t n_k, const uint64_t *b_pos, const void *b_data, int *status) {
    CompactCOOMatrix c;
    c.pos.assign(group_pos, group_pos + n_groups + 1);
    c.data.resize(group_pos[n_groups]);
    for (size_t i = 0; i < c.data.size(); i++) c.data[i] = COOElement{rows[i], cols[i], vals[i]};
    CSRMatrix csr = wrap(n_k, b_pos, b_data);
    auto *res = new RefResult();
    std::vector<COOMatrix> per_group;
    try {
        per_group = compactMulcsr(c, csr);
    } catch (int code) {
        *status = code;
        delete res;
        return nullptr;
    }
    *status = 0;
    COOMatrix flat;
    for (auto &m : per_group) flat.insert(flat.end(), m.begin(), m.end());
    res->products = flat.size();
    std::stable_sort(flat.begin(), flat.end());
    index_t maxRow = 0;
    for (auto &e : c.data) maxRow = std::max(maxRow, e.row);
    const size_t nrows = c.data.empty() ? 0 : size_t(maxRow) + 1;
    res->pos.assign(nrows + 1, 0);
    for (size_t i = 0; i < flat.size();) {
        size_t j = i;
        value_t sum = flat[i].val;
        for (j = i + 1; j < flat.size() && flat[j].row == flat[i].row && flat[j].col == flat[i].col; j++) sum += flat[j].val;
        res->data.push_back(CSRElement{flat[i].col, sum});
        res->pos[flat[i].row + 1]++;
        i = j;
    }
    for (size_t r = 0; r < nrows; r++) res->pos[r + 1] += res->pos[r];
    return res;
}

// reference cscMulcsr (SimSpGEMM.cpp:265-281), flattened in k order, then the
// deduplicateCOO fold with a stable sort, rows = maxRowId+1.
void *ref_spgemm(uint64_t n_k, const uint64_t *a_pos, const void *a_data,
                 const uint64_t *b_pos, const void *b_data) {
    CSRMatrix csc = wrap(n_k, a_pos, a_data), csr = wrap(n_k, b_pos, b_data);
    auto *res = new RefResult();
    auto t0 = std::chrono::high_resolution_clock::now();
    std::vector<COOMatrix> per_k = cscMulcsr(csc, csr);
    COOMatrix flat;
    for (auto &m : per_k) flat.insert(flat.end(), m.begin(), m.end());
    res->products = flat.size();
    std::stable_sort(flat.begin(), flat.end());          // COOElement::operator<, common.h:29-32
    index_t maxRow = 0;
    for (auto &e : csc.data) maxRow = std::max(maxRow, e.idx);
    size_t nrows = size_t(maxRow) + 1;
    res->pos.assign(nrows + 1, 0);
    for (size_t i = 0; i < flat.size(); i++) {
        if (i > 0 && flat[i].row == flat[i - 1].row && flat[i].col == flat[i - 1].col)
            res->data.back().val += flat[i].val;
        else {
            res->data.push_back(CSRElement{flat[i].col, flat[i].val});
            res->pos[flat[i].row + 1]++;
        }
    }
    for (size_t r = 0; r < nrows; r++) res->pos[r + 1] += res->pos[r];
    res->seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    return res;
}

// reference compareCOO (SimSpGEMM.cpp:283-297) on two CSR results expanded to COO.
int ref_compare(uint64_t rows_a, const uint64_t *pos_a, const void *data_a,
                uint64_t rows_b, const uint64_t *pos_b, const void *data_b) {
    auto expand = [](uint64_t rows, const uint64_t *pos, const void *data) {
        COOMatrix c;
        const CSRElement *d = static_cast<const CSRElement *>(data);
        for (uint64_t r = 0; r < rows; r++)
            for (uint64_t e = pos[r]; e < pos[r + 1]; e++) c.push_back(COOElement(index_t(r), d[e]));
        return c;
    };
    return compareCOO(expand(rows_a, pos_a, data_a), expand(rows_b, pos_b, data_b)) ? 1 : 0;
}

void ref_result_dims(void *h, uint64_t *rows, uint64_t *nnz, uint64_t *products, double *seconds) {
    auto *r = static_cast<RefResult *>(h);
    *rows = r->pos.size() - 1; *nnz = r->data.size(); *products = r->products; *seconds = r->seconds;
}
void ref_result_copy(void *h, uint64_t *pos, void *data) {
    auto *r = static_cast<RefResult *>(h);
    std::memcpy(pos, r->pos.data(), r->pos.size() * sizeof(size_t));
    if (!r->data.empty()) std::memcpy(data, r->data.data(), r->data.size() * sizeof(CSRElement));
}
void ref_result_free(void *h) { delete static_cast<RefResult *>(h); }

}  // extern "C"

#elif defined(REF_PART_TASKPROVIDER)

#define private public
#include "SimOuterSPACE.cpp"
#undef private
std::list<Module *> Module::listModules;   // lives in the reference's missing SimCycle.cpp
// SimCycle.h declares these statics too; define whatever the header leaves undefined.

namespace {
struct TPResult {
    std::vector<size_t> pos;
    std::vector<CSRElement> data;
    std::vector<uint32_t> mult_task_sizes;   // per MultiplyTask: nnzc, nnzr
    std::vector<uint32_t> merge_task_ways;   // per MergeTask: #ways, output size
    double seconds = 0;
};
}  // namespace

extern "C" {

// Runs the reference's TaskProvider constructor (multiplyPhase + mergePhase,
// SimOuterSPACE.cpp:46-132) exactly as the simulator does and returns its
// private mergedResult plus the task-size lists the timing models consume.
void *ref_taskprovider(uint64_t n_k, const uint64_t *a_pos, const void *a_data,
                       const uint64_t *b_pos, const void *b_data) {
    CSRMatrix csc, csr;
    csc.pos.assign(a_pos, a_pos + n_k + 1);
    csr.pos.assign(b_pos, b_pos + n_k + 1);
    const CSRElement *ad = static_cast<const CSRElement *>(a_data);
    const CSRElement *bd = static_cast<const CSRElement *>(b_data);
    csc.data.assign(ad, ad + a_pos[n_k]);
    csr.data.assign(bd, bd + b_pos[n_k]);
    auto *res = new TPResult();
    auto t0 = std::chrono::high_resolution_clock::now();
    {
        TaskProvider provider(csc, csr);
        res->seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
        res->pos = provider.mergedResult.pos;
        res->data = provider.mergedResult.data;
        for (auto &t : provider.getMultiplyTasks()) {
            res->mult_task_sizes.push_back(t.lmatCol.size);
            res->mult_task_sizes.push_back(t.rmatRow.size);
        }
        for (auto &t : provider.getMergeTasks()) {
            res->merge_task_ways.push_back(uint32_t(t.inputs.size()));
            res->merge_task_ways.push_back(t.output.size);
        }
    }
    return res;
}
void ref_tp_dims(void *h, uint64_t *rows, uint64_t *nnz, uint64_t *n_mult, uint64_t *n_merge, double *seconds) {
    auto *r = static_cast<TPResult *>(h);
    *rows = r->pos.size() - 1; *nnz = r->data.size();
    *n_mult = r->mult_task_sizes.size() / 2; *n_merge = r->merge_task_ways.size() / 2;
    *seconds = r->seconds;
}
void ref_tp_copy(void *h, uint64_t *pos, void *data, uint32_t *mult_sizes, uint32_t *merge_ways) {
    auto *r = static_cast<TPResult *>(h);
    std::memcpy(pos, r->pos.data(), r->pos.size() * sizeof(size_t));
    if (!r->data.empty()) std::memcpy(data, r->data.data(), r->data.size() * sizeof(CSRElement));
    if (mult_sizes && !r->mult_task_sizes.empty())
        std::memcpy(mult_sizes, r->mult_task_sizes.data(), r->mult_task_sizes.size() * 4);
    if (merge_ways && !r->merge_task_ways.empty())
        std::memcpy(merge_ways, r->merge_task_ways.data(), r->merge_task_ways.size() * 4);
}
void ref_tp_free(void *h) { delete static_cast<TPResult *>(h); }

}  // extern "C"

#else
#error "define REF_PART_SPGEMM or REF_PART_TASKPROVIDER"
#endif
