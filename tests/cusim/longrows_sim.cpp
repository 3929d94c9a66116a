// longrows_sim.cpp -- runs k_long_count / k_long_fill (outerspace_b200/csrc/osp_longrows.cuh, the very source nvcc
// compiles for sm_100a) on the CPU emulation of the execution model.  Test infrastructure: built by
// tests/test_longrows_sim.py with g++ into a temporary .so, never part of the product.
#define OSP_CUSIM 1
#include "osp_longrows.cuh"

#include <vector>

using namespace osp;

namespace {

template <int BAND>
std::vector<uint32_t> band_index(const uint64_t *b_pos, const Elem *b_data, uint64_t n_k, uint64_t cols) {
    const uint32_t n_bands = uint32_t((cols + BAND - 1) / BAND);
    std::vector<uint32_t> bandptr(n_k * (uint64_t(n_bands) + 1) + 1, 0xDEADBEEFu);
    const uint64_t n = n_k * (uint64_t(n_bands) + 1);
    cusim::launch(unsigned((n + 63) / 64), 64, 0, [&] { k_long_bands(b_pos, b_data, n_k, BAND, n_bands, bandptr.data()); });
    return bandptr;
}

template <int THREADS, int BAND, int RUNS>
int run(const uint64_t *a_pos, const Elem *a_data, const uint64_t *b_pos, const Elem *b_data, uint64_t n_k, uint64_t cols, const uint32_t *rows,
        uint32_t n_rows, unsigned grid, uint32_t *count, uint64_t *out_off, Elem *out, uint64_t out_capacity) {
    const std::vector<uint32_t> bandptr = band_index<BAND>(b_pos, b_data, n_k, cols);
    unsigned int ticket = 0;
    const LongRowsListed count_rows{rows, n_rows, &ticket, count, nullptr, nullptr};
    cusim::launch(grid, THREADS, LongRowSmem<BAND, RUNS, false>::bytes, [&] {
        k_long_count<THREADS, BAND, RUNS>(a_pos, a_data, b_data, bandptr.data(), cols, count_rows);
    });
    uint64_t total = 0;
    for (uint32_t x = 0; x < n_rows; x++) { out_off[x] = total; total += count[x]; }
    out_off[n_rows] = total;
    if (total > out_capacity) return 1;
    ticket = 0;
    std::vector<uint32_t> written(n_rows ? n_rows : 1, 0xFFFFFFFFu);
    const LongRowsListed fill_rows{rows, n_rows, &ticket, written.data(), out_off, out};
    cusim::launch(grid, THREADS, LongRowSmem<BAND, RUNS, true>::bytes, [&] {
        k_long_fill<THREADS, BAND, RUNS>(a_pos, a_data, b_data, bandptr.data(), cols, fill_rows);
    });
    for (uint32_t x = 0; x < n_rows; x++)
        if (written[x] != count[x]) return 3;          // the two sweeps must agree on every row's length
    return 0;
}

// The engine's hand-over: rows of the plan's xl list, merged into the start of their own bins, uniq[row] = nnz;
// k_mark_swept flags their tasks for the multiply.
template <int THREADS, int BAND, int RUNS>
int run_bins(const uint64_t *a_pos, const Elem *a_data, const uint64_t *b_pos, const Elem *b_data, uint64_t n_k, uint64_t cols,
             const uint32_t *xl_list, uint32_t n_xl, const uint64_t *row_bin, uint64_t bin_base, Elem *bins, uint32_t *uniq, uint64_t row_lo,
             uint64_t row_hi, uint64_t min_len, uint32_t *swept, unsigned grid) {
    DevScalars sc;
    std::memset(&sc, 0, sizeof(sc));
    sc.n_xl = n_xl;
    cusim::launch(2, 64, 0, [&] { k_mark_swept(a_pos, xl_list, &sc, row_bin, min_len, swept); });
    const std::vector<uint32_t> bandptr = band_index<BAND>(b_pos, b_data, n_k, cols);
    const LongRowsInBins rows{xl_list, &sc, row_bin, bin_base, bins, uniq, row_lo, row_hi, min_len, 2 * min_len};   // two hand-out passes
    cusim::launch(grid, THREADS, LongRowSmem<BAND, RUNS, true>::bytes, [&] {
        k_long_fill<THREADS, BAND, RUNS>(a_pos, a_data, b_data, bandptr.data(), cols, rows);
    });
    return int(sc.err);
}

}  // namespace

extern "C" int lr_sim_bins(int config, const uint64_t *a_pos, const void *a_data, const uint64_t *b_pos, const void *b_data, uint64_t n_k,
                           uint64_t cols, const uint32_t *xl_list, uint32_t n_xl, const uint64_t *row_bin, uint64_t bin_base, void *bins,
                           uint32_t *uniq, uint64_t row_lo, uint64_t row_hi, uint64_t min_len, uint32_t *swept, unsigned grid) {
    const Elem *ad = static_cast<const Elem *>(a_data), *bd = static_cast<const Elem *>(b_data);
    Elem *bn = static_cast<Elem *>(bins);
    if (config == 0) return run_bins<64, 64, 8>(a_pos, ad, b_pos, bd, n_k, cols, xl_list, n_xl, row_bin, bin_base, bn, uniq, row_lo, row_hi, min_len, swept, grid);
    if (config == 1) return run_bins<128, 256, 32>(a_pos, ad, b_pos, bd, n_k, cols, xl_list, n_xl, row_bin, bin_base, bn, uniq, row_lo, row_hi, min_len, swept, grid);
    return -1;
}

// config: 0 = <64 threads, 64-column bands, 8 runs per group>, 1 = <128, 256, 32>, 2 = <32, 32, 64>, 3 = <96, 1024, 5>
extern "C" int lr_sim(int config, const uint64_t *a_pos, const void *a_data, const uint64_t *b_pos, const void *b_data, uint64_t n_k, uint64_t cols,
                      const uint32_t *rows, uint32_t n_rows, unsigned grid, uint32_t *count, uint64_t *out_off, void *out,
                      uint64_t out_capacity) {
    const Elem *ad = static_cast<const Elem *>(a_data), *bd = static_cast<const Elem *>(b_data);
    Elem *o = static_cast<Elem *>(out);
    switch (config) {
    case 0: return run<64, 64, 8>(a_pos, ad, b_pos, bd, n_k, cols, rows, n_rows, grid, count, out_off, o, out_capacity);
    case 1: return run<128, 256, 32>(a_pos, ad, b_pos, bd, n_k, cols, rows, n_rows, grid, count, out_off, o, out_capacity);
    case 2: return run<32, 32, 64>(a_pos, ad, b_pos, bd, n_k, cols, rows, n_rows, grid, count, out_off, o, out_capacity);
    case 3: return run<96, 1024, 5>(a_pos, ad, b_pos, bd, n_k, cols, rows, n_rows, grid, count, out_off, o, out_capacity);
    }
    return 2;
}

extern "C" unsigned long long lr_sim_switches(void) { return cusim::g_switches.load(); }
