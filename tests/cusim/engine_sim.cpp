// engine_sim.cpp -- the WHOLE single-GPU engine (outerspace_b200/csrc/osp_engine.cu: C ABI, host orchestration, every
// kernel) compiled by g++ against tests/cusim: kernels run on the CPU emulation of the execution model, the CUDA
// runtime calls on cusim_runtime.h.  Test infrastructure, built by tests/test_engine_sim.py into a temporary .so and
// loaded by that test only; the product library is built by nvcc from the same, unchanged sources and has no CPU path.
#define OSP_CUSIM 1
#include "osp_engine.cu"

// The multi-GPU entry points (peer memory + NCCL) are outside the emulated build: present so that the Python
// binding finds every symbol, refusing every call.
extern "C" {
int osp_dist_unique_id(void *) { return OSP_ERR_UNSUPPORTED; }
int osp_dist_create(osp_ctx *, const void *, int, int, osp_dist **) { return OSP_ERR_UNSUPPORTED; }
void osp_dist_destroy(osp_dist *) {}
int osp_dist_rows(const osp_dist *, uint64_t, uint64_t *, uint64_t *) { return OSP_ERR_UNSUPPORTED; }
int osp_dist_spgemm(osp_dist *, const osp_spgemm_args *, osp_result **) { return OSP_ERR_UNSUPPORTED; }
const char *cusim_marker(void) { return "tests/cusim build: CPU emulation, test infrastructure only"; }
unsigned long long cusim_device_bytes(void) { return cusim::device_bytes(); }      // live "device" allocations (leak checks)
}
