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
    int64_t* rowptr_exp = nullptr; // [rows*src_panels+1] or null: a row's list grouped by source panel (see cb_graph)
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
    // src_panels S > 1: inside every row the stored neighbours are grouped by panel(c) = (c >> CB_PANEL_SHIFT) % S of
    // their column id c (stable inside a group), so that an aggregation can be run as S passes -- pass p needs only
    // the source rows of panel p -- whose in-order partial sums continue each other: the exchange of panel p+1
    // overlaps the aggregation of panel p at FULL row width.  The grouping does not depend on the slicing, so every
    // world size sums in the same order.
    int src_panels = 1;
    int has_zero_in_deg = 0;
    cb::Side by_dst, by_src;
    float* din_is = nullptr;      // [rows]
    float* dout_is = nullptr;     // [rows]
    int device = 0;
};

namespace cb {

// Live-column compaction of one CSR side (cb_graph_compact_live): the caller-supplied workspace holds a second
// CSR whose rows keep only the columns flagged live, in the stored order, plus the hub-chunk bounds mapped into
// it (a chunk still covers the same ORIGINAL edges, so every partial sum keeps its association).
struct LiveView {
    int64_t* rowptr;     // [rows+1] offsets into col
    int64_t* chunk_beg;  // [n_chunks]
    int64_t* chunk_end;  // [n_chunks]
    int32_t* col;        // [<= n_edges]
    uint32_t* bits;      // [ceil(n_edges/32)] live bit per stored edge
    int32_t* posw;       // [ceil(n_edges/32)] number of live edges before each 32-edge group
    int32_t* spine;      // [ceil(n_edges/LIVE_TILE)+1] exclusive prefix of the tile counts, total in the last slot
    int2* row_be;        // [rows] (begin, end) of every row in col; hub rows are EMPTY here (their chunks produce them)
    int64_t bytes;
};
constexpr int LIVE_TILE = 4096;
LiveView live_view(const cb_graph* g, int side, void* workspace);

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
