// osp_host.cpp -- host side of the drop-in surface: the .mtx loader and COO->CSR/CSC builders.
//
// north_star keeps "the same .mtx loaders, the same CSR/CSC operand structs": these are the
// C-ABI forms of readcoo (simulator/SimSpGEMM.cpp:55-100) and coo2csr<transpose> + dupcheck
// (SimSpGEMM.cpp:43-53,102-152).  They are host code in the reference and stay host code here
// (text parsing and one sort per operand; SURVEY.md 8f ranks their GPU versions as "next").
//
// Behaviour kept: blank and '%' lines are skipped; the first remaining line is "NRow NCol NNZ";
// entries are 1-based; a missing value means 1.0; values are parsed as double then narrowed to
// float; `symmetric` mirrors off-diagonal entries; duplicates make coo2csr fail with 233.
// Deliberate divergences (DESIGN.md "Divergences"): an index >= N is reported as OSP_ERR_INDEX
// instead of writing out of bounds, and the reference's trailing fix-up corner that turns every
// pos into nnz when all non-zeros share one slice (SimSpGEMM.cpp:143-148) is not reproduced.
#include "../../include/osp_b200.h"

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct osp_coo {
    uint64_t nrow = 0, ncol = 0, nnz_header = 0;
    std::vector<uint32_t> rows, cols;
    std::vector<float> vals;
};

namespace {

// Parses an unsigned decimal field the way sscanf("%zu") does: skips white space, accepts digits.
bool parse_u64(const char *&s, uint64_t &out) {
    while (*s == ' ' || *s == '\t' || *s == '\r' || *s == '\n' || *s == '\v' || *s == '\f') s++;
    char *end = nullptr;
    errno = 0;
    unsigned long long v = std::strtoull(s, &end, 10);
    if (end == s) return false;
    out = v;
    s = end;
    return true;
}
bool parse_f64(const char *&s, double &out) {
    char *end = nullptr;
    double v = std::strtod(s, &end);
    if (end == s) return false;
    out = v;
    s = end;
    return true;
}

#pragma pack(push, 1)
struct HostElem {
    uint32_t idx;
    float val;
};
#pragma pack(pop)

}  // namespace

extern "C" {

// Parses Matrix-Market text held in memory (the reference reads from a std::istream, SimSpGEMM.cpp:55).
int osp_readcoo_buffer(const char *text, uint64_t len, int symmetric, osp_coo **out) {
    if ((!text && len) || !out) return OSP_ERR_INVALID;
    *out = nullptr;
    osp_coo *c = new osp_coo();
    std::string line;
    bool header_seen = false;
    uint64_t at = 0;
    while (at < len) {
        uint64_t nl = at;
        while (nl < len && text[nl] != '\n') nl++;
        line.assign(text + at, nl - at);
        at = nl + 1;
        size_t first = line.find_first_not_of(" \t\r");
        if (first == std::string::npos || line[first] == '%') continue;
        const char *s = line.c_str();
        if (!header_seen) {
            uint64_t a = 0, b = 0, z = 0;
            if (parse_u64(s, a) && parse_u64(s, b)) parse_u64(s, z);
            c->nrow = a; c->ncol = b; c->nnz_header = z;
            size_t reserve = symmetric ? 2 * z : z;
            c->rows.reserve(reserve); c->cols.reserve(reserve); c->vals.reserve(reserve);
            header_seen = true;
            continue;
        }
        uint64_t r = 0, col = 0;
        double v = 1.0;
        if (!parse_u64(s, r) || !parse_u64(s, col)) {   // the reference reads garbage here; we refuse
            delete c;
            return OSP_ERR_INVALID;
        }
        double parsed;
        if (parse_f64(s, parsed)) v = parsed;
        c->rows.push_back(uint32_t(r - 1)); c->cols.push_back(uint32_t(col - 1)); c->vals.push_back(float(v));
        if (symmetric && r != col) {
            c->rows.push_back(uint32_t(col - 1)); c->cols.push_back(uint32_t(r - 1)); c->vals.push_back(float(v));
        }
    }
    *out = c;
    return OSP_OK;
}

int osp_readcoo(const char *path, int symmetric, osp_coo **out) {
    if (!path || !out) return OSP_ERR_INVALID;
    *out = nullptr;
    FILE *f = std::fopen(path, "rb");
    if (!f) return OSP_ERR_IO;
    std::string text;
    char buf[1 << 16];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
    std::fclose(f);
    return osp_readcoo_buffer(text.data(), text.size(), symmetric, out);
}

int osp_coo_dims(const osp_coo *c, uint64_t *nrow, uint64_t *ncol, uint64_t *nnz) {
    if (!c) return OSP_ERR_INVALID;
    if (nrow) *nrow = c->nrow;
    if (ncol) *ncol = c->ncol;
    if (nnz) *nnz = c->rows.size();
    return OSP_OK;
}

int osp_coo_copy(const osp_coo *c, uint32_t *rows, uint32_t *cols, float *vals) {
    if (!c) return OSP_ERR_INVALID;
    size_t n = c->rows.size();
    if (n && (!rows || !cols || !vals)) return OSP_ERR_INVALID;
    if (n) {
        std::memcpy(rows, c->rows.data(), n * 4);
        std::memcpy(cols, c->cols.data(), n * 4);
        std::memcpy(vals, c->vals.data(), n * 4);
    }
    return OSP_OK;
}

void osp_coo_free(osp_coo *c) { delete c; }

// Two stable counting passes (minor, then major) instead of the reference's comparison sort:
// the resulting order (major, minor) is the same because keys are unique or the call fails.
int osp_coo2csr(uint64_t nnz, const uint32_t *rows, const uint32_t *cols, const float *vals, uint64_t N, int transpose,
                uint64_t *pos, void *data) {
    if (!pos || (nnz && (!rows || !cols || !vals || !data))) return OSP_ERR_INVALID;
    const uint32_t *major = transpose ? cols : rows;
    const uint32_t *minor = transpose ? rows : cols;
    uint64_t minor_range = 0;
    for (uint64_t i = 0; i < nnz; i++) {
        if (major[i] >= N) return OSP_ERR_INDEX;
        if (uint64_t(minor[i]) + 1 > minor_range) minor_range = uint64_t(minor[i]) + 1;
    }
    // pass 1: order by minor
    std::vector<uint64_t> start(minor_range + 1, 0);
    for (uint64_t i = 0; i < nnz; i++) start[minor[i] + 1]++;
    for (uint64_t i = 0; i < minor_range; i++) start[i + 1] += start[i];
    std::vector<uint64_t> by_minor(nnz);
    for (uint64_t i = 0; i < nnz; i++) by_minor[start[minor[i]]++] = i;
    // pass 2: stable order by major
    for (uint64_t i = 0; i <= N; i++) pos[i] = 0;
    for (uint64_t i = 0; i < nnz; i++) pos[major[i] + 1]++;
    for (uint64_t i = 0; i < N; i++) pos[i + 1] += pos[i];
    std::vector<uint64_t> cursor(pos, pos + N);
    HostElem *out = static_cast<HostElem *>(data);
    for (uint64_t j = 0; j < nnz; j++) {
        uint64_t i = by_minor[j];
        out[cursor[major[i]]++] = HostElem{minor[i], vals[i]};
    }
    // dupcheck: equal neighbours inside a slice
    for (uint64_t s = 0; s < N; s++)
        for (uint64_t e = pos[s] + 1; e < pos[s + 1]; e++)
            if (out[e].idx == out[e - 1].idx) return OSP_ERR_DUPLICATE;
    return OSP_OK;
}

}  // extern "C"
