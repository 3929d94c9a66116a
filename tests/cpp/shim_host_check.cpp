// Host-only check of include/osp_b200.hpp (no GPU): readcoo + coo2csr<> through the C++ shim.
// usage: shim_host_check file.mtx  -> prints NRow NCol nnz, then csr pos/data, then csc pos/data, then the compact
// forms (csr2compact of the CSR, csc2rawcompact of the CSC) as "cpos ..." / "cdata row:col:val ..."
#include <cstdio>
#include <fstream>

#include "../../include/osp_b200.hpp"

using namespace osp_b200;

static void print(const CSRMatrix &m) {
    std::printf("pos");
    for (size_t p : m.pos) std::printf(" %zu", p);
    std::printf("\ndata");
    for (auto &e : m.data) std::printf(" %u:%a", e.idx, (double)e.val);
    std::printf("\n");
}

static void print(const CompactCOOMatrix &m) {
    std::printf("cpos");
    for (size_t p : m.pos) std::printf(" %zu", p);
    std::printf("\ncdata");
    for (auto &e : m.data) std::printf(" %u:%u:%a", e.row, e.col, (double)e.val);
    std::printf("\n");
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    std::ifstream fin(argv[1]);
    size_t NRow = 0, NCol = 0;
    COOMatrix coo = readcoo(fin, NRow, NCol, argc > 2);
    std::printf("%zu %zu %zu\n", NRow, NCol, coo.size());
    try {
        const CSRMatrix csr = coo2csr(coo, NRow), csc = coo2csr<true>(coo, NCol);
        print(csr);
        print(csc);
        print(csr2compact(csr));
        print(csc2rawcompact(csc));
    } catch (int code) {
        std::printf("throw %d\n", code);
    }
    return 0;
}
