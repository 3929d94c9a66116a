// oracle/spgemm_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference's functional outer-product SpGEMM path, used
// as the parity checker by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs.  Nothing under outerspace_b200/ may
// link, import or call this file: the product path is CUDA only.
//
// Parity pin: the reference (anneouyang/OuterSPACE) ships NO tests, golden
// vectors or fixtures for this path.  The restatement is
// therefore pinned against outputs of the reference's own, unmodified source
// compiled in this container (oracle/_ref, built by oracle/Makefile from
// /root/reference/simulator) -- see tests/test_oracle_vs_ref.py and the
// committed fixtures tests/golden/*.npz made by tests/golden/make_golden.py.
//
// What each function restates (reference file:line, relative to simulator/):
//   orc_mtx_*      readcoo                      SimSpGEMM.cpp:55-100
//   orc_coo2csr    coo2csr<false|true>+dupcheck SimSpGEMM.cpp:43-53,102-152
//   orc_spgemm     cscMulcsr                    SimSpGEMM.cpp:265-281
//                  + deduplicateCOO             SimSpGEMM.cpp:519-535 (in #if 0)
//                  + row count = maxRowId+1     SimOuterSPACE.cpp:49-53,99-102,131
//   orc_csr2csc    coo2csr<true> applied to an already-CSR operand
//   orc_flops      mulflops_ref                 SimSpGEMM.cpp:884-891
//
// Canonical value semantics: the product a*b is rounded to fp32
// on its own, then duplicates of one (row,col) are left-folded in ascending k
// (the reference sorts with an unstable std::sort, which leaves that order
// implementation-defined for >2 duplicates; the oracle fixes it with a stable
// order, i.e. the "deterministic k-ordered merge mode" of BASELINE.json).
//
// Build: g++ -O3 -std=c++17 -shared -fPIC (flags of simulator/Makefile:10-14:
// -O3, no -ffast-math, no -march => no FMA contraction of a*b + acc).

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

namespace {

#pragma pack(push, 1)
struct Elem {            // mirrors CSRElement, common.h:10-16 (8 bytes, packed)
    uint32_t idx;
    float val;
};
#pragma pack(pop)
static_assert(sizeof(Elem) == 8, "element must be 8 bytes");

struct Triplet {         // mirrors COOElement, common.h:18-33 (12 bytes)
    uint32_t row, col;
    float val;
};

struct MtxFile {
    uint64_t nrow = 0, ncol = 0, nnz_header = 0;
    std::vector<Triplet> coo;
};

struct Result {
    std::vector<uint64_t> pos;
    std::vector<Elem> data;
    uint64_t products = 0;
};

// A line is a comment/blank when its first character other than ' ' or '\t'
// is '%' or does not exist (SimSpGEMM.cpp:66-77).
bool is_skippable(const std::string &line) {
    size_t p = line.find_first_not_of(" \t");
    return p == std::string::npos || line[p] == '%';
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------- readcoo ----
// Returns NULL when the file cannot be opened.
void *orc_mtx_open(const char *path, int symmetric) {
    std::ifstream in(path);
    if (!in) return nullptr;
    auto *m = new MtxFile();
    bool have_header = false;
    std::string line;
    while (std::getline(in, line)) {
        if (is_skippable(line)) continue;
        if (!have_header) {
            size_t r = 0, c = 0, z = 0;
            std::sscanf(line.c_str(), "%zu %zu %zu", &r, &c, &z);
            m->nrow = r; m->ncol = c; m->nnz_header = z;
            m->coo.reserve(symmetric ? 2 * z : z);
            have_header = true;
            continue;
        }
        size_t r = 0, c = 0;
        double v = 0.0;
        int got = std::sscanf(line.c_str(), "%zu %zu %lf", &r, &c, &v);
        if (got < 3) v = 1.0;                      // pattern matrices: value 1
        m->coo.push_back(Triplet{uint32_t(r - 1), uint32_t(c - 1), float(v)});
        if (symmetric && r != c)
            m->coo.push_back(Triplet{uint32_t(c - 1), uint32_t(r - 1), float(v)});
    }
    return m;
}

void orc_mtx_dims(void *h, uint64_t *nrow, uint64_t *ncol, uint64_t *nnz) {
    auto *m = static_cast<MtxFile *>(h);
    *nrow = m->nrow; *ncol = m->ncol; *nnz = m->coo.size();
}

void orc_mtx_copy(void *h, uint32_t *rows, uint32_t *cols, float *vals) {
    auto *m = static_cast<MtxFile *>(h);
    for (size_t i = 0; i < m->coo.size(); i++) {
        rows[i] = m->coo[i].row; cols[i] = m->coo[i].col; vals[i] = m->coo[i].val;
    }
}

void orc_mtx_free(void *h) { delete static_cast<MtxFile *>(h); }

// ---------------------------------------------------------------- coo2csr ----
// transpose=0: CSR (major=row, minor=col); transpose=1: CSC (major=col,
// minor=row).  pos has N+1 entries, data has nnz entries holding the minor
// index.  Returns 0, or 233 on a duplicate (row,col) -- the value the
// reference throws (SimSpGEMM.cpp:49) -- or 1 when a major index is >= N.
// `single_row_quirk` != 0 reproduces the reference's fix-up corner
// (SimSpGEMM.cpp:143-148): when every non-zero lies in one major slice, all of
// pos becomes nnz.  The product path does NOT reproduce that corner.
int orc_coo2csr(uint64_t nnz, const uint32_t *rows, const uint32_t *cols, const float *vals,
                uint64_t N, int transpose, int single_row_quirk, uint64_t *pos, void *data_out) {
    const uint32_t *major = transpose ? cols : rows;
    const uint32_t *minor = transpose ? rows : cols;
    std::vector<uint64_t> order(nnz);
    for (uint64_t i = 0; i < nnz; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) {
        if (major[a] != major[b]) return major[a] < major[b];
        return minor[a] < minor[b];
    });
    for (uint64_t i = 0; i + 1 < nnz; i++) {
        uint64_t a = order[i], b = order[i + 1];
        if (major[a] == major[b] && minor[a] == minor[b]) return 233;
    }
    for (uint64_t i = 0; i <= N; i++) pos[i] = 0;
    Elem *data = static_cast<Elem *>(data_out);
    for (uint64_t i = 0; i < nnz; i++) {
        uint64_t s = order[i];
        if (major[s] >= N) return 1;
        pos[major[s] + 1]++;
        data[i] = Elem{minor[s], vals[s]};
    }
    for (uint64_t i = 0; i < N; i++) pos[i + 1] += pos[i];
    if (single_row_quirk && nnz > 0) {
        bool one_slice = true;
        for (uint64_t i = 1; i < nnz; i++) one_slice = one_slice && (major[i] == major[0]);
        if (one_slice)
            for (uint64_t i = 0; i <= N; i++) pos[i] = nnz;
    }
    return 0;
}

// ---------------------------------------------------------------- csr2csc ----
// Stable transposition of a compressed operand with `n_major` slices whose
// minor indices are < n_minor: the output has n_minor slices and, inside each,
// the original major indices in ascending order (what coo2csr<true> yields by
// comparison sort, SimSpGEMM.cpp:111-117).  Returns 1 on an out-of-range index.
int orc_csr2csc(uint64_t n_major, uint64_t n_minor, const uint64_t *pos, const void *data_in,
                uint64_t *pos_out, void *data_out) {
    const Elem *in = static_cast<const Elem *>(data_in);
    Elem *out = static_cast<Elem *>(data_out);
    uint64_t nnz = pos[n_major];
    for (uint64_t i = 0; i <= n_minor; i++) pos_out[i] = 0;
    for (uint64_t e = 0; e < nnz; e++) {
        if (in[e].idx >= n_minor) return 1;
        pos_out[in[e].idx + 1]++;
    }
    for (uint64_t i = 0; i < n_minor; i++) pos_out[i + 1] += pos_out[i];
    std::vector<uint64_t> cursor(pos_out, pos_out + n_minor);
    for (uint64_t r = 0; r < n_major; r++)
        for (uint64_t e = pos[r]; e < pos[r + 1]; e++)
            out[cursor[in[e].idx]++] = Elem{uint32_t(r), in[e].val};
    return 0;
}

// ------------------------------------------------------------------ flops ----
uint64_t orc_flops(uint64_t n_k, const uint64_t *a_pos, const uint64_t *b_pos) {
    uint64_t p = 0;
    for (uint64_t k = 0; k < n_k; k++)
        p += (a_pos[k + 1] - a_pos[k]) * (b_pos[k + 1] - b_pos[k]);
    return p;
}

// ----------------------------------------------------------------- spgemm ----
// C = A*B with A as CSC (idx = row ids) and B as CSR (idx = col ids), both with
// n_k slices.  Literal two-phase restatement: (1) every k with both slices
// non-empty emits (row, col, a*b) in (k, j, t) order; (2) stable sort by
// (row, col), left fold of equal keys, rows = max row id + 1.
// rows_override > 0 forces that many rows instead (must be > max row id).
void *orc_spgemm(uint64_t n_k, const uint64_t *a_pos, const void *a_data_in,
                 const uint64_t *b_pos, const void *b_data_in, uint64_t rows_override) {
    const Elem *a = static_cast<const Elem *>(a_data_in);
    const Elem *b = static_cast<const Elem *>(b_data_in);
    auto *res = new Result();

    std::vector<Triplet> partial;
    partial.reserve(orc_flops(n_k, a_pos, b_pos));
    uint32_t max_row = 0;
    for (uint64_t e = 0; e < a_pos[n_k]; e++) max_row = std::max(max_row, a[e].idx);
    for (uint64_t k = 0; k < n_k; k++) {
        if (a_pos[k] == a_pos[k + 1] || b_pos[k] == b_pos[k + 1]) continue;
        for (uint64_t j = a_pos[k]; j < a_pos[k + 1]; j++)
            for (uint64_t t = b_pos[k]; t < b_pos[k + 1]; t++) {
                float prod = a[j].val * b[t].val;           // rounded on its own
                partial.push_back(Triplet{a[j].idx, b[t].idx, prod});
            }
    }
    res->products = partial.size();
    std::stable_sort(partial.begin(), partial.end(), [](const Triplet &x, const Triplet &y) {
        return x.row == y.row ? x.col < y.col : x.row < y.row;
    });

    uint64_t nrows = rows_override ? rows_override : uint64_t(max_row) + 1;
    res->pos.assign(nrows + 1, 0);
    for (size_t i = 0; i < partial.size(); i++) {
        const Triplet &p = partial[i];
        if (i > 0 && partial[i - 1].row == p.row && partial[i - 1].col == p.col) {
            res->data.back().val += p.val;                  // left fold, ascending k
        } else {
            res->data.push_back(Elem{p.col, p.val});
            res->pos[p.row + 1]++;
        }
    }
    for (uint64_t r = 0; r < nrows; r++) res->pos[r + 1] += res->pos[r];
    return res;
}

// Same result as orc_spgemm, computed row block by row block so that the
// transient triplet list stays small: used for the CPU baseline on workloads
// whose full partial-product list would not fit in host memory.  Needs A as CSR
// (a_csr_*: m rows, idx = k) so that a row block is a contiguous slice.
void *orc_spgemm_rowblocks(uint64_t m, const uint64_t *a_csr_pos, const void *a_csr_data,
                           const uint64_t *b_pos, const void *b_data_in, uint64_t rows_per_block) {
    const Elem *a = static_cast<const Elem *>(a_csr_data);
    const Elem *b = static_cast<const Elem *>(b_data_in);
    auto *res = new Result();
    uint64_t last_nonempty = 0;
    for (uint64_t r = 0; r < m; r++)
        if (a_csr_pos[r + 1] > a_csr_pos[r]) last_nonempty = r + 1;
    uint64_t nrows = last_nonempty ? last_nonempty : 1;     // maxRowId+1 (0 nnz -> 1 row)
    res->pos.assign(nrows + 1, 0);
    std::vector<Triplet> partial;
    for (uint64_t r0 = 0; r0 < nrows; r0 += rows_per_block) {
        uint64_t r1 = std::min(nrows, r0 + rows_per_block);
        partial.clear();
        for (uint64_t r = r0; r < r1; r++)
            for (uint64_t e = a_csr_pos[r]; e < a_csr_pos[r + 1]; e++) {
                uint64_t k = a[e].idx;
                for (uint64_t t = b_pos[k]; t < b_pos[k + 1]; t++)
                    partial.push_back(Triplet{uint32_t(r), b[t].idx, a[e].val * b[t].val});
            }
        res->products += partial.size();
        std::stable_sort(partial.begin(), partial.end(), [](const Triplet &x, const Triplet &y) {
            return x.row == y.row ? x.col < y.col : x.row < y.row;
        });
        for (size_t i = 0; i < partial.size(); i++) {
            const Triplet &p = partial[i];
            if (i > 0 && partial[i - 1].row == p.row && partial[i - 1].col == p.col) {
                res->data.back().val += p.val;
            } else {
                res->data.push_back(Elem{p.col, p.val});
                res->pos[p.row + 1]++;
            }
        }
    }
    for (uint64_t r = 0; r < nrows; r++) res->pos[r + 1] += res->pos[r];
    return res;
}

void orc_result_dims(void *h, uint64_t *rows, uint64_t *nnz, uint64_t *products) {
    auto *r = static_cast<Result *>(h);
    *rows = r->pos.size() - 1; *nnz = r->data.size(); *products = r->products;
}

void orc_result_copy(void *h, uint64_t *pos, void *data) {
    auto *r = static_cast<Result *>(h);
    std::memcpy(pos, r->pos.data(), r->pos.size() * sizeof(uint64_t));
    if (!r->data.empty()) std::memcpy(data, r->data.data(), r->data.size() * sizeof(Elem));
}

void orc_result_free(void *h) { delete static_cast<Result *>(h); }

}  // extern "C"
