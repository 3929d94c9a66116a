// longrows_sim.cpp -- runs k_long_count / k_long_fill (outerspace_b200/csrc/osp_longrows.cuh, the very source nvcc
// compiles for sm_100a) on the CPU emulation of the execution model.  Test infrastructure: built by
// tests/test_longrows_sim.py with g++ into a temporary .so, never part of the product.
#define OSP_CUSIM 1
#include "osp_longrows.cuh"

#include <vector>

using namespace osp;

namespace {

template <int THREADS, int BAND, int RUNS>
int run(const uint64_t *a_pos, const Elem *a_data, const uint64_t *b_pos, const Elem *b_data, uint64_t cols, const uint32_t *rows,
        uint32_t n_rows, unsigned grid, uint32_t *count, uint64_t *out_off, Elem *out, uint64_t out_capacity) {
    uint64_t stride = 1;
    for (uint32_t x = 0; x < n_rows; x++) stride = std::max<uint64_t>(stride, a_pos[rows[x] + 1] - a_pos[rows[x]]);
    std::vector<uint32_t> cursors(size_t(grid) * stride, 0xDEADBEEFu);
    unsigned int ticket = 0;
    cusim::launch(grid, THREADS, LongRowSmem<BAND, RUNS, false>::bytes, [&] {
        k_long_count<THREADS, BAND, RUNS>(a_pos, a_data, b_pos, b_data, cols, rows, n_rows, &ticket, cursors.data(), stride, count);
    });
    uint64_t total = 0;
    for (uint32_t x = 0; x < n_rows; x++) { out_off[x] = total; total += count[x]; }
    out_off[n_rows] = total;
    if (total > out_capacity) return 1;
    ticket = 0;
    std::fill(cursors.begin(), cursors.end(), 0xDEADBEEFu);
    cusim::launch(grid, THREADS, LongRowSmem<BAND, RUNS, true>::bytes, [&] {
        k_long_fill<THREADS, BAND, RUNS>(a_pos, a_data, b_pos, b_data, cols, rows, n_rows, &ticket, cursors.data(), stride, out_off, out);
    });
    return 0;
}

}  // namespace

// config: 0 = <64 threads, 64-column bands, 8 runs per group>, 1 = <128, 256, 32>, 2 = <32, 32, 64>, 3 = <96, 1024, 5>
extern "C" int lr_sim(int config, const uint64_t *a_pos, const void *a_data, const uint64_t *b_pos, const void *b_data, uint64_t cols,
                      const uint32_t *rows, uint32_t n_rows, unsigned grid, uint32_t *count, uint64_t *out_off, void *out,
                      uint64_t out_capacity) {
    const Elem *ad = static_cast<const Elem *>(a_data), *bd = static_cast<const Elem *>(b_data);
    Elem *o = static_cast<Elem *>(out);
    switch (config) {
    case 0: return run<64, 64, 8>(a_pos, ad, b_pos, bd, cols, rows, n_rows, grid, count, out_off, o, out_capacity);
    case 1: return run<128, 256, 32>(a_pos, ad, b_pos, bd, cols, rows, n_rows, grid, count, out_off, o, out_capacity);
    case 2: return run<32, 32, 64>(a_pos, ad, b_pos, bd, cols, rows, n_rows, grid, count, out_off, o, out_capacity);
    case 3: return run<96, 1024, 5>(a_pos, ad, b_pos, bd, cols, rows, n_rows, grid, count, out_off, o, out_capacity);
    }
    return 2;
}

extern "C" unsigned long long lr_sim_switches(void) { return cusim::g_switches; }
