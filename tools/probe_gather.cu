// probe_gather.cu -- development probe (not part of the product): what does a random gather of short rows cost on a B200?
// The row-order multiply of config 4 gathers 6.7e7 rows of B of ~64 bytes at random 8-byte alignment; ncu showed
// 17.95 GB read for 4.8 GB asked (profiles/README.md).  This probe isolates the access pattern:
//   rows of `len` elements (8 B each) at random offsets of a 512 MiB buffer, offsets aligned to `align` bytes,
//   one group of 8 lanes per row, under different cudaLimitMaxL2FetchGranularity settings and load flavours.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probe_gather.bin tools/probe_gather.cu
// Run under ncu for dram__bytes_read.sum per launch (kernel names carry the variant).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE>
__device__ __forceinline__ uint2 ld8(const uint2 *p) {
    uint2 v;
    if (MODE == 0) v = *p;
    else if (MODE == 1) v = __ldg(p);
    else if (MODE == 2) asm volatile("ld.global.L2::64B.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    else if (MODE == 3) asm volatile("ld.global.L2::128B.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    else asm volatile("ld.global.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

// one group of 8 lanes per row; rows of `len` <= 16 elements
template <int MODE>
__global__ void __launch_bounds__(256) k_gather(const uint2 *__restrict__ buf, const uint32_t *__restrict__ start, uint32_t n_rows, uint32_t len, float *out) {
    const uint32_t g = (blockIdx.x * 256u + threadIdx.x) >> 3, l = threadIdx.x & 7;
    const uint32_t ngroups = (gridDim.x * 256u) >> 3;
    float acc = 0.f;
    for (uint32_t r = g; r < n_rows; r += ngroups * 4) {
        uint32_t s[4];
#pragma unroll
        for (int u = 0; u < 4; u++) s[u] = r + u * ngroups < n_rows ? start[r + u * ngroups] : 0xFFFFFFFFu;
        uint2 v[4][2];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            v[u][0] = make_uint2(0, 0); v[u][1] = make_uint2(0, 0);
            if (s[u] != 0xFFFFFFFFu) {
                if (l < len) v[u][0] = ld8<MODE>(buf + s[u] + l);
                if (l + 8 < len) v[u][1] = ld8<MODE>(buf + s[u] + l + 8);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) acc += __uint_as_float(v[u][0].y) + __uint_as_float(v[u][1].y);
    }
    if (acc == 123.456f) out[0] = acc;
}

// the mirror image: random scatter of rows (the outer-product order writes its partial products like this)
__global__ void __launch_bounds__(256) k_scatter(uint2 *__restrict__ buf, const uint32_t *__restrict__ start, uint32_t n_rows, uint32_t len) {
    const uint32_t g = (blockIdx.x * 256u + threadIdx.x) >> 3, l = threadIdx.x & 7;
    const uint32_t ngroups = (gridDim.x * 256u) >> 3;
    for (uint32_t r = g; r < n_rows; r += ngroups) {
        const uint32_t s = start[r];
        if (l < len) buf[s + l] = make_uint2(r, l);
        if (l + 8 < len) buf[s + l + 8] = make_uint2(r, l);
    }
}

int main(int argc, char **argv) {
    const size_t elems = size_t(64) << 20;                  // 512 MiB of 8-byte elements
    const uint32_t n_rows = argc > 1 ? atoi(argv[1]) : (64u << 20) / 8 * 8;   // 6.7e7 rows like config 4
    uint2 *buf; uint32_t *start; float *out;
    CK(cudaMalloc(&buf, elems * 8 + 4096));
    CK(cudaMemset(buf, 1, elems * 8 + 4096));
    CK(cudaMalloc(&start, size_t(n_rows) * 4));
    CK(cudaMalloc(&out, 4));
    std::vector<uint32_t> h(n_rows);
    size_t lim = 0;
    CK(cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity));
    printf("default cudaLimitMaxL2FetchGranularity = %zu\n", lim);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grid = 148 * 16;
    for (int gran : {0, 32, 128}) {
        if (gran) {
            cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
            CK(cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity));
            printf("set granularity %d -> %s, now %zu\n", gran, cudaGetErrorString(e), lim);
        }
        for (uint32_t len : {8u, 4u, 16u}) {
            for (uint32_t align : {8u, 32u, 64u, 128u}) {
                uint64_t x = 88172645463325252ull + align + len;
                for (uint32_t i = 0; i < n_rows; i++) {
                    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
                    uint64_t off = (x % (elems - 32)) * 8;
                    off &= ~uint64_t(align - 1);
                    h[i] = uint32_t(off / 8);
                }
                CK(cudaMemcpy(start, h.data(), size_t(n_rows) * 4, cudaMemcpyHostToDevice));
                float ms[5] = {0, 0, 0, 0, 0};
                for (int rep = 0; rep < 2; rep++) {
                    CK(cudaEventRecord(e0)); k_gather<0><<<grid, 256>>>(buf, start, n_rows, len, out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms[0], e0, e1));
                    CK(cudaEventRecord(e0)); k_gather<1><<<grid, 256>>>(buf, start, n_rows, len, out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms[1], e0, e1));
                    CK(cudaEventRecord(e0)); k_gather<2><<<grid, 256>>>(buf, start, n_rows, len, out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms[2], e0, e1));
                    CK(cudaEventRecord(e0)); k_gather<3><<<grid, 256>>>(buf, start, n_rows, len, out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms[3], e0, e1));
                    CK(cudaEventRecord(e0)); k_scatter<<<grid, 256>>>(buf, start, n_rows, len); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms[4], e0, e1));
                }
                const double gb = double(n_rows) * len * 8 / 1e9;
                printf("gran %3d len %2u align %3u: useful %.2f GB | ld %.3f ms (%.0f GB/s)  ldg %.3f  L2::64B %.3f  L2::128B %.3f | scatter %.3f ms (%.0f GB/s)\n",
                       gran, len, align, gb, ms[0], gb / ms[0] * 1e3, ms[1], ms[2], ms[3], ms[4], gb / ms[4] * 1e3);
                fflush(stdout);
            }
        }
    }
    return 0;
}
