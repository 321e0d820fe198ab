// Device-wide exclusive scan (three kernels, 1024 items per block) and the small host helpers shared by the
// graph-construction and graph-preparation translation units (internal, not part of the ABI).
#pragma once

#include <vector>

#include "cb_internal.cuh"

namespace cb {

// ---- exclusive scan int32 -> int64 (three kernels, 1024 items per block) ----------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t* total) {
    __shared__ int64_t warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int64_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    int64_t base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < SCAN_THREADS / 32; ++i) {
        const int64_t s = warp_sums[i];
        if (i < w) base += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return base + inc - v;
}

template <typename Map>
__global__ void k_scan_block_sums(int64_t n, Map map, int64_t* __restrict__ block_sums) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < n) s += map(base + i);
    int64_t tot;
    block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

static __global__ void k_scan_spine(int64_t n_blocks, int64_t* __restrict__ block_sums) {
    // single block; sequential over tiles of SCAN_THREADS entries, exclusive in place
    int64_t carry = 0;
    for (int64_t b = 0; b < n_blocks; b += SCAN_THREADS) {
        const int64_t i = b + threadIdx.x;
        const int64_t v = i < n_blocks ? block_sums[i] : 0;
        int64_t tot;
        const int64_t ex = block_exclusive_scan(v, &tot);
        if (i < n_blocks) block_sums[i] = carry + ex;
        carry += tot;
        __syncthreads();
    }
}

template <typename Map>
__global__ void k_scan_write(int64_t n, Map map, const int64_t* __restrict__ block_sums,
                             int64_t* __restrict__ out) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int64_t v[SCAN_ITEMS];
    int64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? map(base + i) : 0;
        s += v[i];
    }
    int64_t tot;
    int64_t run = block_exclusive_scan(s, &tot) + block_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
        if (base + i == n - 1) out[n] = run;  // grand total in the extra slot
    }
}

struct DegMap {
    const int32_t* deg;
    __device__ int64_t operator()(int64_t i) const { return deg[i]; }
};
struct ChunkCountMap {
    const int32_t* deg;
    int hub_chunk;
    __device__ int64_t operator()(int64_t i) const {
        const int d = deg[i];
        return d > hub_chunk ? (d + hub_chunk - 1) / hub_chunk : 0;
    }
};

template <typename Map>
static int exclusive_scan(int64_t n, Map map, int64_t* out /*[n+1]*/, int64_t* spine, cudaStream_t st) {
    if (n == 0) {
        CB_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t), st));
        return CB_OK;
    }
    const int64_t nb = ceil_div(n, SCAN_TILE);
    k_scan_block_sums<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(n, map, spine);
    CB_LAUNCH_CHECK();
    k_scan_spine<<<1, SCAN_THREADS, 0, st>>>(nb, spine);
    CB_LAUNCH_CHECK();
    k_scan_write<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(n, map, spine, out);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

static int grid_for(int64_t n, int threads) {
    const int64_t want = ceil_div(n > 0 ? n : 1, threads);
    const int64_t cap = (int64_t)sm_count() * 16;
    return (int)(want < cap ? want : cap);
}

static int bits_for(uint32_t max_value) {  // number of low bits that can be non-zero in [0, max_value]
    int b = 1;
    while (b < 32 && (max_value >> b) != 0) ++b;
    return b;
}

struct Scratch {  // frees whatever build() allocated for temporary use, on every exit path
    std::vector<void*> ptrs;
    ~Scratch() {
        for (void* p : ptrs) cudaFree(p);
    }
    template <typename T>
    cudaError_t alloc(T** p, int64_t n) {
        cudaError_t e = cudaMalloc((void**)p, (size_t)(n > 0 ? n : 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

}  // namespace cb
