// Graph preparation either side of the TeacherGNN path, on the device.
//
// The reference prepares its graphs with O(E) Python loops over .tolist()-ed edge lists and numpy calls on the host:
//   utils.py:300-334  graph_analyze            edges per node as origin / as destination
//   utils.py:667-674  ensure_symmetric         coalesced indices of A + A^T
//   utils.py:910-943  get_partial_sorted_idx   repeated-median selection of the low / high degree share
//   utils.py:676-730  save_graph_analyze       Table-1 statistics, head / tail / isolated splits
//   utils.py:732-752  craft_isolation_v2       order-preserving removal of the edges of "isolated" nodes
// These are the same results -- values AND order -- as integer kernels: one atomic histogram pass, a 64-bit radix sort
// (cub::DeviceRadixSort, header library of the toolkit) + adjacent-difference compaction, order statistics read off
// one sorted copy, and flag / scan / scatter compactions with the scan of cb_scan.cuh.  All index work: bit-exact.
//
// Setup-time calls: scratch comes from the caller (cb_prep_graph_workspace_bytes), so that a framework's caching allocator
// serves it (cudaMalloc / cudaFree of the sort buffers cost more than the kernels); each call synchronises the stream
// once or twice to hand a count (or the key range that sizes the radix sort) back to the host.
#include <cub/device/device_radix_sort.cuh>

#include "cb_internal.cuh"
#include "cb_scan.cuh"

namespace cb {
namespace prep {

// ---- graph_analyze ---------------------------------------------------------------------------------------------
// ids >= num_nodes are ignored like the reference's dict lookup over range(N_nodes) ignores them; negative ids are an
// error (numpy would index from the end).
__global__ void k_degrees(const int64_t* __restrict__ ori, const int64_t* __restrict__ dst, int64_t E, int64_t N,
                          unsigned int* __restrict__ d_ori, unsigned int* __restrict__ d_dst, int* __restrict__ err) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
        const int64_t o = ori[e], t = dst[e];
        if (o < 0 || t < 0) {
            *err = 1;
            continue;
        }
        if (o < N) atomicAdd(d_ori + o, 1u);
        if (t < N) atomicAdd(d_dst + t, 1u);
    }
}

__global__ void k_widen(const unsigned int* __restrict__ a, const unsigned int* __restrict__ b, int64_t n,
                        int64_t* __restrict__ oa, int64_t* __restrict__ ob) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        oa[i] = a[i];
        ob[i] = b[i];
    }
}

// sort key of a value: its offset from the smallest one, so that the radix sort only walks the bits the range needs
__global__ void k_offset_keys(const int64_t* __restrict__ a, int64_t n, long long lo, uint64_t* __restrict__ key) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        key[i] = (uint64_t)((long long)a[i] - lo);
}

// ---- ensure_symmetric ------------------------------------------------------------------------------------------
__global__ void k_max_id(const int64_t* __restrict__ ids, int64_t n, long long* __restrict__ out /*[2]: max, min*/) {
    long long mx = LLONG_MIN, mn = LLONG_MAX;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long v = ids[i];
        mx = v > mx ? v : mx;
        mn = v < mn ? v : mn;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long a = __shfl_xor_sync(0xffffffffu, mx, o), b = __shfl_xor_sync(0xffffffffu, mn, o);
        mx = a > mx ? a : mx;
        mn = b < mn ? b : mn;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out, mx);
        atomicMin(out + 1, mn);
    }
}

// key of (row, col) and of (col, row): the order coalesce() yields is the order of row * n + col
__global__ void k_sym_keys(const int64_t* __restrict__ ori, const int64_t* __restrict__ dst, int64_t E, uint64_t n,
                           uint64_t* __restrict__ key) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
        const uint64_t o = (uint64_t)ori[e], t = (uint64_t)dst[e];
        key[e] = o * n + t;
        key[E + e] = t * n + o;
    }
}

struct FirstOfRunMap {    // 1 where a sorted key differs from its predecessor
    const uint64_t* key;
    __device__ int64_t operator()(int64_t i) const { return (i == 0 || key[i] != key[i - 1]) ? 1 : 0; }
};

__global__ void k_sym_decode(const uint64_t* __restrict__ key, const int64_t* __restrict__ pos, int64_t m, uint64_t n,
                             int64_t* __restrict__ out_row, int64_t* __restrict__ out_col) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const uint64_t k = key[i];
        if (i == 0 || k != key[i - 1]) {
            const int64_t p = pos[i];
            out_row[p] = (int64_t)(k / n);
            out_col[p] = (int64_t)(k % n);
        }
    }
}

// ---- get_partial_sorted_idx ------------------------------------------------------------------------------------
// np.median of an even count is the mean of the two middle values; everything is compared doubled (2 a <= a_lo + a_hi)
// so the arithmetic stays in integers.  The set {a <= m} is a prefix of the sorted copy, {a >= m} a suffix: each level
// is one median lookup and one binary search.  thr[0] = doubled threshold, thr[1] = 1 if the selection is empty.
__global__ void k_repeated_median(const uint64_t* __restrict__ s, int64_t n, int top, int levels, long long lo_value,
                                  long long* __restrict__ thr) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int64_t cnt = n;
    long long m2 = 0;
    for (int l = 0; l < levels && cnt > 0; ++l) {
        const int64_t base = top ? 0 : n - cnt;
        m2 = (long long)s[base + (cnt - 1) / 2] + (long long)s[base + cnt / 2];
        int64_t lo = 0, hi = n;
        if (top) {   // number of elements with 2 a <= m2
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (2 * (long long)s[mid] <= m2) lo = mid + 1; else hi = mid;
            }
            cnt = lo;
        } else {     // number of elements with 2 a >= m2
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (2 * (long long)s[mid] < m2) lo = mid + 1; else hi = mid;
            }
            cnt = n - lo;
        }
    }
    thr[0] = m2 + 2 * lo_value;      // back from offsets (a - lo_value) to values: 2 a <= m2 + 2 lo
    thr[1] = cnt == 0;
}

struct SelectMap {
    const int64_t* a;
    const long long* thr;
    int top;
    __device__ int64_t operator()(int64_t i) const {
        if (thr[1]) return 0;
        const long long v = 2 * (long long)a[i];
        return (top ? v <= thr[0] : v >= thr[0]) ? 1 : 0;
    }
};

template <typename Map>
__global__ void k_scatter_index(int64_t n, Map map, const int64_t* __restrict__ pos, int64_t* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (map(i)) out[pos[i]] = i;
}

// ---- degree statistics (gen_rec_for_table1_stats, utils.py:676-678) --------------------------------------------
__global__ void k_deg_stats(const int64_t* __restrict__ d, int64_t n, unsigned long long* __restrict__ acc /*sum, zeros*/,
                            long long* __restrict__ mx) {
    unsigned long long s = 0, z = 0;
    long long m = LLONG_MIN;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long v = d[i];
        s += (unsigned long long)v;
        z += v == 0;
        m = v > m ? v : m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        z += __shfl_xor_sync(0xffffffffu, z, o);
        const long long b = __shfl_xor_sync(0xffffffffu, m, o);
        m = b > m ? b : m;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(acc, s);
        atomicAdd(acc + 1, z);
        atomicMax(mx, m);
    }
}

// ---- stable sort of an index list by the value it points at ------------------------------------------------------
__global__ void k_gather_keys(const int64_t* __restrict__ arr, const int64_t* __restrict__ idx, int64_t m, int64_t n,
                              int64_t* __restrict__ key, int* __restrict__ err) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const int64_t j = idx[i];
        if (j < 0 || j >= n) {
            *err = 1;
            key[i] = 0;
        } else {
            key[i] = arr[j];
        }
    }
}

__global__ void k_mask_from_idx(const int64_t* __restrict__ idx, int64_t m, int64_t n, uint8_t* __restrict__ mask,
                                int* __restrict__ err) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const int64_t j = idx[i];
        if (j < 0 || j >= n) *err = 1; else mask[j] = 1;
    }
}

// ---- craft_isolation_v2 ------------------------------------------------------------------------------------------
struct KeepEdgeMap {    // an edge goes if it is not a self loop and touches a flagged node
    const int64_t* ori;
    const int64_t* dst;
    const uint8_t* mask;
    int64_t N;
    __device__ int64_t operator()(int64_t e) const {
        const int64_t o = ori[e], t = dst[e];
        const bool zo = o >= 0 && o < N && mask[o], zt = t >= 0 && t < N && mask[t];
        return ((o != t) && (zo || zt)) ? 0 : 1;
    }
};

__global__ void k_compact_edges(int64_t E, KeepEdgeMap keep, const int64_t* __restrict__ pos,
                                int64_t* __restrict__ out_ori, int64_t* __restrict__ out_dst) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
        if (keep(e)) {
            const int64_t p = pos[e];
            out_ori[p] = keep.ori[e];
            out_dst[p] = keep.dst[e];
        }
    }
}

static int bits64(uint64_t max_value) {
    int b = 1;
    while (b < 64 && (max_value >> b) != 0) ++b;
    return b;
}

// ---- caller-supplied scratch ----------------------------------------------------------------------------------------
struct Carver {     // base == nullptr: only measures
    char* base;
    int64_t off = 0;
    template <typename T>
    T* take(int64_t n) {
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += ((n > 0 ? n : 1) * (int64_t)sizeof(T) + 255) & ~(int64_t)255;
        return p;
    }
};

static int64_t cub_keys_bytes(int64_t n) {
    size_t bytes = 0;
    cub::DoubleBuffer<uint64_t> kb(nullptr, nullptr);
    cub::DeviceRadixSort::SortKeys(nullptr, bytes, kb, (int)n, 0, 64, (cudaStream_t)0);
    return (int64_t)bytes;
}

static int64_t cub_pairs_bytes(int64_t n) {
    size_t bytes = 0;
    cub::DoubleBuffer<uint64_t> kb(nullptr, nullptr);
    cub::DoubleBuffer<int64_t> vb(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, kb, vb, (int)n, 0, 64, (cudaStream_t)0);
    return (int64_t)bytes;
}

struct SortedCopy {     // scratch of "a sorted copy of n 64-bit keys"
    uint64_t *key, *alt;
    void* cub;
    int64_t cub_bytes;
    long long* mm;      // [2] max, min of the values
};

static SortedCopy carve_sorted(Carver& c, int64_t n) {
    SortedCopy s;
    s.key = c.take<uint64_t>(n);
    s.alt = c.take<uint64_t>(n);
    s.cub_bytes = cub_keys_bytes(n);
    s.cub = c.take<char>(s.cub_bytes);
    s.mm = c.take<long long>(2);
    return s;
}

// sorted copy of arr as offsets from its minimum: *sorted [n] ascending, *lo the minimum (host)
static int sorted_offsets(const int64_t* arr, int64_t n, const SortedCopy& w, const uint64_t** sorted, long long* lo,
                          cudaStream_t st) {
    const long long init[2] = {LLONG_MIN, LLONG_MAX};
    CB_CUDA(cudaMemcpyAsync(w.mm, init, sizeof(init), cudaMemcpyHostToDevice, st));
    k_max_id<<<grid_for(n, 256), 256, 0, st>>>(arr, n, w.mm);
    CB_LAUNCH_CHECK();
    long long h[2];
    CB_CUDA(cudaMemcpyAsync(h, w.mm, sizeof(h), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    *lo = h[1];
    k_offset_keys<<<grid_for(n, 256), 256, 0, st>>>(arr, n, h[1], w.key);
    CB_LAUNCH_CHECK();
    const int end_bit = bits64((uint64_t)h[0] - (uint64_t)h[1]);
    cub::DoubleBuffer<uint64_t> kb(w.key, w.alt);
    size_t bytes = (size_t)w.cub_bytes;
    CB_CUDA(cub::DeviceRadixSort::SortKeys(w.cub, bytes, kb, (int)n, 0, end_bit, st));
    count_launch(2 * ((end_bit + 7) / 8));
    *sorted = kb.Current();
    return CB_OK;
}

struct SymWs { long long* mm; uint64_t *key, *alt; void* cub; int64_t cub_bytes; int64_t *pos, *spine; };
static SymWs carve_sym(Carver& c, int64_t E) {
    SymWs w;
    const int64_t M = 2 * E;
    w.mm = c.take<long long>(2);
    w.key = c.take<uint64_t>(M);
    w.alt = c.take<uint64_t>(M);
    w.cub_bytes = cub_keys_bytes(M);
    w.cub = c.take<char>(w.cub_bytes);
    w.pos = c.take<int64_t>(M + 1);
    w.spine = c.take<int64_t>(ceil_div(M > 0 ? M : 1, SCAN_TILE) + 1);
    return w;
}

struct SelWs { SortedCopy s; long long* thr; int64_t *pos, *spine; };
static SelWs carve_sel(Carver& c, int64_t n) {
    SelWs w;
    w.s = carve_sorted(c, n);
    w.thr = c.take<long long>(2);
    w.pos = c.take<int64_t>(n + 1);
    w.spine = c.take<int64_t>(ceil_div(n > 0 ? n : 1, SCAN_TILE) + 1);
    return w;
}

struct StatWs { SortedCopy s; unsigned long long* acc; long long* mx; };
static StatWs carve_stat(Carver& c, int64_t n) {
    StatWs w;
    w.s = carve_sorted(c, n);
    w.acc = c.take<unsigned long long>(2);
    w.mx = c.take<long long>(1);
    return w;
}

struct PairWs { uint64_t *key, *key_alt; int64_t *val, *val_alt; void* cub; int64_t cub_bytes; long long* mm; int* err; };
static PairWs carve_pair(Carver& c, int64_t m) {
    PairWs w;
    w.key = c.take<uint64_t>(m);
    w.key_alt = c.take<uint64_t>(m);
    w.val = c.take<int64_t>(m);
    w.val_alt = c.take<int64_t>(m);
    w.cub_bytes = cub_pairs_bytes(m);
    w.cub = c.take<char>(w.cub_bytes);
    w.mm = c.take<long long>(2);
    w.err = c.take<int>(1);
    return w;
}

struct DropWs { int64_t *pos, *spine; };
static DropWs carve_drop(Carver& c, int64_t E) {
    DropWs w;
    w.pos = c.take<int64_t>(E + 1);
    w.spine = c.take<int64_t>(ceil_div(E > 0 ? E : 1, SCAN_TILE) + 1);
    return w;
}

struct DegWs { unsigned int *a, *b; int* err; };
static DegWs carve_deg(Carver& c, int64_t N) {
    DegWs w;
    w.a = c.take<unsigned int>(N);
    w.b = c.take<unsigned int>(N);
    w.err = c.take<int>(1);
    return w;
}

}  // namespace prep
}  // namespace cb

extern "C" {

using namespace cb;
using namespace cb::prep;

int64_t cb_prep_graph_workspace_bytes(int what, int64_t n) {
    if (n < 0 || n >= (int64_t)INT32_MAX / 2) return 0;
    Carver c{nullptr};
    switch (what) {
        case CB_PREP_DEGREES: carve_deg(c, n); break;
        case CB_PREP_SYMMETRIZE: carve_sym(c, n); break;
        case CB_PREP_PARTIAL_SORTED_IDX: carve_sel(c, n); break;
        case CB_PREP_DEGREE_STATS: carve_stat(c, n); break;
        case CB_PREP_SORT_IDX_BY_VALUE: carve_pair(c, n); break;
        case CB_PREP_MASK_FROM_IDX: c.take<int>(1); break;
        case CB_PREP_DROP_EDGES: carve_drop(c, n); break;
        default: return 0;
    }
    return c.off;
}

#define CB_PREP_WS(what, n)                                                                                   \
    CB_REQUIRE(workspace != nullptr && workspace_bytes >= cb_prep_graph_workspace_bytes(what, n), CB_E_WORKSPACE,    \
               "graph preparation: workspace missing or smaller than cb_prep_graph_workspace_bytes()");             \
    Carver carver { (char*)workspace }

int cb_prep_degrees(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes, int64_t* degs_ori,
                    int64_t* degs_dst, void* workspace, int64_t workspace_bytes, void* stream) {
    CB_REQUIRE(num_edges >= 0 && num_nodes >= 0, CB_E_INVALID, "cb_prep_degrees: negative size");
    CB_REQUIRE(num_edges == 0 || edge_index, CB_E_INVALID, "cb_prep_degrees: edge_index is NULL");
    CB_REQUIRE(num_nodes == 0 || (degs_ori && degs_dst), CB_E_INVALID, "cb_prep_degrees: an output is NULL");
    CB_REQUIRE(num_edges < (int64_t)UINT32_MAX, CB_E_UNSUPPORTED, "cb_prep_degrees: 2^32 edges or more");
    CB_REQUIRE(num_nodes < (int64_t)INT32_MAX / 2, CB_E_UNSUPPORTED, "cb_prep_degrees: 2^30 nodes or more");
    CB_PREP_WS(CB_PREP_DEGREES, num_nodes);
    const DegWs w = carve_deg(carver, num_nodes);
    cudaStream_t st = (cudaStream_t)stream;
    CB_CUDA(cudaMemsetAsync(w.err, 0, sizeof(int), st));
    if (num_nodes > 0) {
        CB_CUDA(cudaMemsetAsync(w.a, 0, (size_t)num_nodes * sizeof(unsigned int), st));
        CB_CUDA(cudaMemsetAsync(w.b, 0, (size_t)num_nodes * sizeof(unsigned int), st));
    }
    if (num_edges > 0) {
        k_degrees<<<grid_for(num_edges, 256), 256, 0, st>>>(edge_index, edge_index + num_edges, num_edges, num_nodes,
                                                            w.a, w.b, w.err);
        CB_LAUNCH_CHECK();
    }
    if (num_nodes > 0) {
        k_widen<<<grid_for(num_nodes, 256), 256, 0, st>>>(w.a, w.b, num_nodes, degs_ori, degs_dst);
        CB_LAUNCH_CHECK();
    }
    int h = 0;
    CB_CUDA(cudaMemcpyAsync(&h, w.err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    CB_REQUIRE(h == 0, CB_E_RANGE, "cb_prep_degrees: negative node id in the edge list");
    return CB_OK;
}

int cb_prep_symmetrize(const int64_t* edge_index, int64_t num_edges, int64_t* out, int64_t* count, void* workspace,
                       int64_t workspace_bytes, void* stream) {
    CB_REQUIRE(num_edges >= 0, CB_E_INVALID, "cb_prep_symmetrize: negative size");
    CB_REQUIRE(count != nullptr, CB_E_INVALID, "cb_prep_symmetrize: count is NULL");
    *count = 0;
    if (num_edges == 0) return CB_OK;
    CB_REQUIRE(edge_index && out, CB_E_INVALID, "cb_prep_symmetrize: a buffer is NULL");
    CB_REQUIRE(2 * num_edges < (int64_t)INT32_MAX, CB_E_UNSUPPORTED, "cb_prep_symmetrize: 2^30 edges or more");
    CB_PREP_WS(CB_PREP_SYMMETRIZE, num_edges);
    const SymWs w = carve_sym(carver, num_edges);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t E = num_edges, M = 2 * num_edges;
    const long long init[2] = {LLONG_MIN, LLONG_MAX};
    CB_CUDA(cudaMemcpyAsync(w.mm, init, sizeof(init), cudaMemcpyHostToDevice, st));
    k_max_id<<<grid_for(M, 256), 256, 0, st>>>(edge_index, M, w.mm);
    CB_LAUNCH_CHECK();
    long long h[2];
    CB_CUDA(cudaMemcpyAsync(h, w.mm, sizeof(h), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    CB_REQUIRE(h[1] >= 0, CB_E_RANGE, "cb_prep_symmetrize: negative node id in the edge list");
    const uint64_t n = (uint64_t)h[0] + 1;     // N = edge_index.max() + 1 (utils.py:669)
    CB_REQUIRE(n <= (1ull << 32), CB_E_UNSUPPORTED, "cb_prep_symmetrize: node ids must be below 2^32");
    k_sym_keys<<<grid_for(E, 256), 256, 0, st>>>(edge_index, edge_index + E, E, n, w.key);
    CB_LAUNCH_CHECK();
    const int end_bit = bits64(n * n - 1);
    cub::DoubleBuffer<uint64_t> kb(w.key, w.alt);
    size_t bytes = (size_t)w.cub_bytes;
    CB_CUDA(cub::DeviceRadixSort::SortKeys(w.cub, bytes, kb, (int)M, 0, end_bit, st));
    count_launch(2 * ((end_bit + 7) / 8));
    const uint64_t* sorted = kb.Current();
    int rc = exclusive_scan(M, FirstOfRunMap{sorted}, w.pos, w.spine, st);
    if (rc) return rc;
    k_sym_decode<<<grid_for(M, 256), 256, 0, st>>>(sorted, w.pos, M, n, out, out + M);
    CB_LAUNCH_CHECK();
    CB_CUDA(cudaMemcpyAsync(count, w.pos + M, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    return CB_OK;
}

int cb_prep_partial_sorted_idx(const int64_t* arr, int64_t n, int top, int levels, int64_t* idx_out, int64_t* count,
                               void* workspace, int64_t workspace_bytes, void* stream) {
    CB_REQUIRE(n >= 0 && levels >= 1 && levels <= 5, CB_E_INVALID, "cb_prep_partial_sorted_idx: bad size or level");
    CB_REQUIRE(count != nullptr, CB_E_INVALID, "cb_prep_partial_sorted_idx: count is NULL");
    *count = 0;
    if (n == 0) return CB_OK;
    CB_REQUIRE(arr && idx_out, CB_E_INVALID, "cb_prep_partial_sorted_idx: a buffer is NULL");
    CB_REQUIRE(n < (int64_t)INT32_MAX / 2, CB_E_UNSUPPORTED, "cb_prep_partial_sorted_idx: 2^30 values or more");
    CB_PREP_WS(CB_PREP_PARTIAL_SORTED_IDX, n);
    const SelWs w = carve_sel(carver, n);
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t* sorted = nullptr;
    long long lo = 0;
    int rc = sorted_offsets(arr, n, w.s, &sorted, &lo, st);
    if (rc) return rc;
    k_repeated_median<<<1, 32, 0, st>>>(sorted, n, top ? 1 : 0, levels, lo, w.thr);
    CB_LAUNCH_CHECK();
    const SelectMap sel{arr, w.thr, top ? 1 : 0};
    rc = exclusive_scan(n, sel, w.pos, w.spine, st);
    if (rc) return rc;
    k_scatter_index<<<grid_for(n, 256), 256, 0, st>>>(n, sel, w.pos, idx_out);
    CB_LAUNCH_CHECK();
    CB_CUDA(cudaMemcpyAsync(count, w.pos + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    return CB_OK;
}

int cb_prep_degree_stats(const int64_t* degs, int64_t n, double* stats, void* workspace, int64_t workspace_bytes,
                         void* stream) {
    CB_REQUIRE(n > 0 && degs && stats, CB_E_INVALID, "cb_prep_degree_stats: empty input or NULL buffer");
    CB_REQUIRE(n < (int64_t)INT32_MAX / 2, CB_E_UNSUPPORTED, "cb_prep_degree_stats: 2^30 values or more");
    CB_PREP_WS(CB_PREP_DEGREE_STATS, n);
    const StatWs w = carve_stat(carver, n);
    cudaStream_t st = (cudaStream_t)stream;
    CB_CUDA(cudaMemsetAsync(w.acc, 0, 2 * sizeof(unsigned long long), st));
    const long long lowest = LLONG_MIN;
    CB_CUDA(cudaMemcpyAsync(w.mx, &lowest, sizeof(lowest), cudaMemcpyHostToDevice, st));
    k_deg_stats<<<grid_for(n, 256), 256, 0, st>>>(degs, n, w.acc, w.mx);
    CB_LAUNCH_CHECK();
    const uint64_t* sorted = nullptr;
    long long lo = 0;
    int rc = sorted_offsets(degs, n, w.s, &sorted, &lo, st);
    if (rc) return rc;
    unsigned long long h_acc[2], mid[2];
    long long h_mx;
    CB_CUDA(cudaMemcpyAsync(h_acc, w.acc, sizeof(h_acc), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaMemcpyAsync(&h_mx, w.mx, sizeof(h_mx), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaMemcpyAsync(&mid[0], sorted + (n - 1) / 2, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaMemcpyAsync(&mid[1], sorted + n / 2, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    stats[0] = (double)n;
    stats[1] = (double)h_acc[0];
    stats[2] = (double)h_mx;
    stats[3] = (double)h_acc[0] / (double)n;
    stats[4] = ((double)((long long)mid[0] + lo) + (double)((long long)mid[1] + lo)) / 2.0;
    stats[5] = (double)h_acc[1] / (double)n * 100.0;
    return CB_OK;
}

int cb_prep_sort_idx_by_value(const int64_t* arr, int64_t n, const int64_t* idx, int64_t m, int64_t* idx_sorted,
                              void* workspace, int64_t workspace_bytes, void* stream) {
    CB_REQUIRE(n >= 0 && m >= 0, CB_E_INVALID, "cb_prep_sort_idx_by_value: negative size");
    if (m == 0) return CB_OK;
    CB_REQUIRE(arr && idx && idx_sorted, CB_E_INVALID, "cb_prep_sort_idx_by_value: a buffer is NULL");
    CB_REQUIRE(m < (int64_t)INT32_MAX / 2, CB_E_UNSUPPORTED, "cb_prep_sort_idx_by_value: 2^30 items or more");
    CB_PREP_WS(CB_PREP_SORT_IDX_BY_VALUE, m);
    const PairWs w = carve_pair(carver, m);
    cudaStream_t st = (cudaStream_t)stream;
    CB_CUDA(cudaMemsetAsync(w.err, 0, sizeof(int), st));
    // the gathered values first (as int64 in key_alt), their range, then keys = offsets from the minimum
    int64_t* gathered = reinterpret_cast<int64_t*>(w.key_alt);
    k_gather_keys<<<grid_for(m, 256), 256, 0, st>>>(arr, idx, m, n, gathered, w.err);
    CB_LAUNCH_CHECK();
    const long long init[2] = {LLONG_MIN, LLONG_MAX};
    CB_CUDA(cudaMemcpyAsync(w.mm, init, sizeof(init), cudaMemcpyHostToDevice, st));
    k_max_id<<<grid_for(m, 256), 256, 0, st>>>(gathered, m, w.mm);
    CB_LAUNCH_CHECK();
    long long h[2];
    int herr = 0;
    CB_CUDA(cudaMemcpyAsync(h, w.mm, sizeof(h), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaMemcpyAsync(&herr, w.err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    CB_REQUIRE(herr == 0, CB_E_RANGE, "cb_prep_sort_idx_by_value: index outside [0, n)");
    k_offset_keys<<<grid_for(m, 256), 256, 0, st>>>(gathered, m, h[1], w.key);
    CB_LAUNCH_CHECK();
    CB_CUDA(cudaMemcpyAsync(w.val, idx, (size_t)m * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    const int end_bit = bits64((uint64_t)h[0] - (uint64_t)h[1]);
    cub::DoubleBuffer<uint64_t> kb(w.key, w.key_alt);
    cub::DoubleBuffer<int64_t> vb(w.val, w.val_alt);
    size_t bytes = (size_t)w.cub_bytes;
    CB_CUDA(cub::DeviceRadixSort::SortPairs(w.cub, bytes, kb, vb, (int)m, 0, end_bit, st));     // LSD radix sort: stable
    count_launch(2 * ((end_bit + 7) / 8));
    CB_CUDA(cudaMemcpyAsync(idx_sorted, vb.Current(), (size_t)m * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    CB_CUDA(cudaStreamSynchronize(st));
    return CB_OK;
}

int cb_prep_mask_from_idx(const int64_t* idx, int64_t m, int64_t num_nodes, uint8_t* mask, void* workspace,
                          int64_t workspace_bytes, void* stream) {
    CB_REQUIRE(m >= 0 && num_nodes >= 0, CB_E_INVALID, "cb_prep_mask_from_idx: negative size");
    CB_REQUIRE(num_nodes == 0 || mask, CB_E_INVALID, "cb_prep_mask_from_idx: mask is NULL");
    CB_PREP_WS(CB_PREP_MASK_FROM_IDX, 0);
    int* err = carver.take<int>(1);
    cudaStream_t st = (cudaStream_t)stream;
    CB_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
    if (num_nodes > 0) CB_CUDA(cudaMemsetAsync(mask, 0, (size_t)num_nodes, st));
    if (m > 0) {
        CB_REQUIRE(idx != nullptr, CB_E_INVALID, "cb_prep_mask_from_idx: idx is NULL");
        k_mask_from_idx<<<grid_for(m, 256), 256, 0, st>>>(idx, m, num_nodes, mask, err);
        CB_LAUNCH_CHECK();
    }
    int h = 0;
    CB_CUDA(cudaMemcpyAsync(&h, err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    CB_REQUIRE(h == 0, CB_E_RANGE, "cb_prep_mask_from_idx: index outside [0, num_nodes)");
    return CB_OK;
}

int cb_prep_drop_edges(const int64_t* edge_index, int64_t num_edges, const uint8_t* node_mask, int64_t num_nodes,
                       int64_t* out, int64_t* kept, void* workspace, int64_t workspace_bytes, void* stream) {
    CB_REQUIRE(num_edges >= 0 && num_nodes >= 0, CB_E_INVALID, "cb_prep_drop_edges: negative size");
    CB_REQUIRE(num_edges < (int64_t)INT32_MAX / 2, CB_E_UNSUPPORTED, "cb_prep_drop_edges: 2^30 edges or more");
    CB_REQUIRE(kept != nullptr, CB_E_INVALID, "cb_prep_drop_edges: kept is NULL");
    *kept = 0;
    if (num_edges == 0) return CB_OK;
    CB_REQUIRE(edge_index && out && (node_mask || num_nodes == 0), CB_E_INVALID, "cb_prep_drop_edges: a buffer is NULL");
    CB_PREP_WS(CB_PREP_DROP_EDGES, num_edges);
    const DropWs w = carve_drop(carver, num_edges);
    cudaStream_t st = (cudaStream_t)stream;
    const KeepEdgeMap keep{edge_index, edge_index + num_edges, node_mask, num_nodes};
    int rc = exclusive_scan(num_edges, keep, w.pos, w.spine, st);
    if (rc) return rc;
    k_compact_edges<<<grid_for(num_edges, 256), 256, 0, st>>>(num_edges, keep, w.pos, out, out + num_edges);
    CB_LAUNCH_CHECK();
    CB_CUDA(cudaMemcpyAsync(kept, w.pos + num_edges, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    return CB_OK;
}

}  // extern "C"
