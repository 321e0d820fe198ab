// Internal declarations shared by the translation units of libcoldbrew_b200.so (not part of the ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "coldbrew_b200.h"

namespace cb {

// One CSR view of the (possibly sliced) graph.  Rows are the owned nodes, columns are global ids.
struct Side {
    int64_t* rowptr = nullptr;     // [rows+1]
    int32_t* col = nullptr;        // [n_edges]
    int32_t* perm = nullptr;       // [n_edges] position of the stored edge in the caller's edge list
    int32_t* deg = nullptr;        // [rows]
    int64_t n_edges = 0;
    // hub rows (deg > hub_chunk) are cut into chunks; chunk c covers col[chunk_beg[c] .. +hub_chunk)
    int32_t* chunk_row = nullptr;  // [n_chunks] local row of the chunk
    int64_t* chunk_beg = nullptr;  // [n_chunks]
    int64_t n_chunks = 0;
};

}  // namespace cb

struct cb_graph {
    int64_t n_nodes = 0;          // global N (bound of every column id)
    int64_t row_begin = 0, row_end = 0;
    int64_t rows = 0;
    int hub_chunk = CB_DEFAULT_HUB_CHUNK;
    int has_zero_in_deg = 0;
    cb::Side by_dst, by_src;
    float* din_is = nullptr;      // [rows]
    float* dout_is = nullptr;     // [rows]
    int device = 0;
};

namespace cb {

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define CB_CUDA(expr)                                                             \
    do {                                                                          \
        cudaError_t _e = (expr);                                                  \
        if (_e != cudaSuccess) return cb::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define CB_REQUIRE(cond, code, msg)        \
    do {                                   \
        if (!(cond)) {                     \
            cb::set_error(msg);            \
            return (code);                 \
        }                                  \
    } while (0)

// checks the launch that was just issued (configuration errors surface here, not asynchronously)
#define CB_LAUNCH_CHECK()                  \
    do {                                   \
        cb::count_launch();                \
        CB_CUDA(cudaGetLastError());       \
    } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Number of SMs of the current device (cached); grids of the streaming kernels are sized from it.
int sm_count();

}  // namespace cb
