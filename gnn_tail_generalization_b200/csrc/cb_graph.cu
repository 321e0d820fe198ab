// Device-side graph construction: COO edge_index -> CSR by destination + CSR by source, degrees,
// degree^-1/2 vectors, zero-in-degree flag and the hub-chunk work lists.
//
// Replaces the reference's host-side graph build (GNN_model/GCN.py:92-94: edge_index -> Python lists ->
// dgl.graph), the per-forward degree recomputation (GCN.py:205-209, 242-246) and the per-forward
// zero-in-degree scan with its device->host sync (GCN.py:187-188).
//
// The stable key sort uses cub::DeviceRadixSort (header library shipped with the CUDA toolkit); every
// other step is a kernel in this file.
#include <cub/device/device_radix_sort.cuh>

#include <vector>

#include "cb_internal.cuh"
#include "cb_scan.cuh"

namespace cb {

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------

// One pass over the edge list: range-check both endpoints, build the 32-bit sort key of each side
// (local row, or `rows` as the "not owned" sentinel that sorts to the end) and count degrees.
__global__ void k_edge_keys(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E,
                            int64_t N, int64_t row_begin, int64_t row_end, uint32_t* __restrict__ key_dst,
                            uint32_t* __restrict__ key_src, int32_t* __restrict__ eid_a,
                            int32_t* __restrict__ eid_b, int32_t* __restrict__ in_deg,
                            int32_t* __restrict__ out_deg, int* __restrict__ err) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint32_t rows = (uint32_t)(row_end - row_begin);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
        const int64_t s = src[e], t = dst[e];
        if (s < 0 || s >= N || t < 0 || t >= N) {
            *err = 1;
            key_dst[e] = rows;
            key_src[e] = rows;
        } else {
            const bool own_t = (t >= row_begin && t < row_end);
            const bool own_s = (s >= row_begin && s < row_end);
            key_dst[e] = own_t ? (uint32_t)(t - row_begin) : rows;
            key_src[e] = own_s ? (uint32_t)(s - row_begin) : rows;
            if (own_t) atomicAdd(in_deg + (t - row_begin), 1);
            if (own_s) atomicAdd(out_deg + (s - row_begin), 1);
        }
        eid_a[e] = (int32_t)e;
        eid_b[e] = (int32_t)e;
    }
}

// One side only (cb_graph_create_local): sort key = local row of the owned endpoint, degree count, range check.
__global__ void k_edge_keys_side(const int64_t* __restrict__ own_end, const int64_t* __restrict__ other_end, int64_t E,
                                 int64_t N, int64_t row_begin, int64_t row_end, int panels, int filter,
                                 uint32_t* __restrict__ key, int32_t* __restrict__ eid, int32_t* __restrict__ deg,
                                 int32_t* __restrict__ deg_exp, int* __restrict__ err) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint32_t sentinel = (uint32_t)(row_end - row_begin) * (uint32_t)panels;   // sorts behind every owned edge
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
        const int64_t o = own_end[e], t = other_end[e];
        if (o < 0 || o >= N || t < 0 || t >= N) {
            err[0] = 1;
            key[e] = sentinel;
        } else if (o < row_begin || o >= row_end) {
            if (!filter) err[2] = 1;          // the caller promised edges of the owned rows only
            key[e] = sentinel;
        } else {
            const uint32_t r = (uint32_t)(o - row_begin);
            const uint32_t k = panels > 1 ? r * (uint32_t)panels + (uint32_t)((t >> CB_PANEL_SHIFT) % panels) : r;
            key[e] = k;
            atomicAdd(deg + r, 1);
            if (deg_exp) atomicAdd(deg_exp + k, 1);
        }
        eid[e] = (int32_t)e;
    }
}

// col[j] = other endpoint of the j-th stored edge
__global__ void k_gather_cols(const int64_t* __restrict__ other, const int32_t* __restrict__ perm,
                              int64_t n, int32_t* __restrict__ col) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
        col[j] = (int32_t)other[perm[j]];
}

// degree^-1/2 with the reference's clamp(min=1); computed in double and rounded once to fp32
__global__ void k_inv_sqrt_deg(const int32_t* __restrict__ deg, int64_t rows, float* __restrict__ out,
                               int* __restrict__ zero_flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
        const int d = deg[r];
        if (d == 0 && zero_flag) *zero_flag = 1;
        out[r] = (float)(1.0 / sqrt((double)(d < 1 ? 1 : d)));
    }
}

// expand hub rows into their chunk descriptors
__global__ void k_fill_chunks(const int32_t* __restrict__ deg, const int64_t* __restrict__ rowptr,
                              const int64_t* __restrict__ chunk_ofs, int64_t rows, int hub_chunk,
                              int32_t* __restrict__ chunk_row, int64_t* __restrict__ chunk_beg) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
        const int d = deg[r];
        if (d <= hub_chunk) continue;
        const int n = (d + hub_chunk - 1) / hub_chunk;
        const int64_t o = chunk_ofs[r], b = rowptr[r];
        for (int k = 0; k < n; ++k) {
            chunk_row[o + k] = (int32_t)r;
            chunk_beg[o + k] = b + (int64_t)k * hub_chunk;
        }
    }
}

static void free_side(Side& s) {
    cudaFree(s.rowptr);
    cudaFree(s.col);
    cudaFree(s.perm);
    cudaFree(s.deg);
    cudaFree(s.rowptr_exp);
    cudaFree(s.chunk_row);
    cudaFree(s.chunk_beg);
    s = Side();
}

// Finish one side given sorted edge ids (first n_edges entries are the owned edges in stable order).
static int finish_side(Side& side, const int64_t* other_endpoint, int64_t rows, int hub_chunk,
                       int64_t* spine, cudaStream_t st) {
    CB_CUDA(cudaMalloc((void**)&side.rowptr, (size_t)(rows + 1) * sizeof(int64_t)));
    int rc = exclusive_scan(rows, DegMap{side.deg}, side.rowptr, spine, st);
    if (rc) return rc;
    CB_CUDA(cudaMemcpyAsync(&side.n_edges, side.rowptr + rows, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    CB_CUDA(cudaMalloc((void**)&side.col, (size_t)(side.n_edges > 0 ? side.n_edges : 1) * sizeof(int32_t)));
    if (side.n_edges > 0) {
        k_gather_cols<<<grid_for(side.n_edges, 256), 256, 0, st>>>(other_endpoint, side.perm, side.n_edges,
                                                                    side.col);
        CB_LAUNCH_CHECK();
    }
    // hub chunk lists
    Scratch tmp;
    int64_t* chunk_ofs = nullptr;
    CB_CUDA(tmp.alloc(&chunk_ofs, rows + 1));
    rc = exclusive_scan(rows, ChunkCountMap{side.deg, hub_chunk}, chunk_ofs, spine, st);
    if (rc) return rc;
    CB_CUDA(cudaMemcpyAsync(&side.n_chunks, chunk_ofs + rows, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    if (side.n_chunks > 0) {
        CB_CUDA(cudaMalloc((void**)&side.chunk_row, (size_t)side.n_chunks * sizeof(int32_t)));
        CB_CUDA(cudaMalloc((void**)&side.chunk_beg, (size_t)side.n_chunks * sizeof(int64_t)));
        k_fill_chunks<<<grid_for(rows, 256), 256, 0, st>>>(side.deg, side.rowptr, chunk_ofs, rows, hub_chunk,
                                                           side.chunk_row, side.chunk_beg);
        CB_LAUNCH_CHECK();
        CB_CUDA(cudaStreamSynchronize(st));
    }
    return CB_OK;
}

static int build(cb_graph* g, const int64_t* edge_index, int64_t E, cudaStream_t st) {
    const int64_t rows = g->rows;
    const int64_t* src = edge_index;
    const int64_t* dst = edge_index + E;
    Scratch tmp;

    CB_CUDA(cudaMalloc((void**)&g->by_dst.deg, (size_t)(rows > 0 ? rows : 1) * sizeof(int32_t)));
    CB_CUDA(cudaMalloc((void**)&g->by_src.deg, (size_t)(rows > 0 ? rows : 1) * sizeof(int32_t)));
    CB_CUDA(cudaMemsetAsync(g->by_dst.deg, 0, (size_t)(rows > 0 ? rows : 1) * sizeof(int32_t), st));
    CB_CUDA(cudaMemsetAsync(g->by_src.deg, 0, (size_t)(rows > 0 ? rows : 1) * sizeof(int32_t), st));
    CB_CUDA(cudaMalloc((void**)&g->din_is, (size_t)(rows > 0 ? rows : 1) * sizeof(float)));
    CB_CUDA(cudaMalloc((void**)&g->dout_is, (size_t)(rows > 0 ? rows : 1) * sizeof(float)));

    uint32_t *key_dst = nullptr, *key_src = nullptr, *key_alt = nullptr;
    int32_t *eid_a = nullptr, *eid_b = nullptr, *eid_alt = nullptr;
    int* flags = nullptr;  // [0] = range error, [1] = zero in-degree
    int64_t* spine = nullptr;
    CB_CUDA(tmp.alloc(&key_dst, E));
    CB_CUDA(tmp.alloc(&key_src, E));
    CB_CUDA(tmp.alloc(&key_alt, E));
    CB_CUDA(tmp.alloc(&eid_a, E));
    CB_CUDA(tmp.alloc(&eid_b, E));
    CB_CUDA(tmp.alloc(&eid_alt, E));
    CB_CUDA(tmp.alloc(&flags, 2));
    CB_CUDA(tmp.alloc(&spine, ceil_div(rows > 0 ? rows : 1, SCAN_TILE) + 1));
    CB_CUDA(cudaMemsetAsync(flags, 0, 2 * sizeof(int), st));

    if (E > 0) {
        k_edge_keys<<<grid_for(E, 256), 256, 0, st>>>(src, dst, E, g->n_nodes, g->row_begin, g->row_end,
                                                      key_dst, key_src, eid_a, eid_b, g->by_dst.deg,
                                                      g->by_src.deg, flags);
        CB_LAUNCH_CHECK();
    }
    k_inv_sqrt_deg<<<grid_for(rows, 256), 256, 0, st>>>(g->by_dst.deg, rows, g->din_is, flags + 1);
    CB_LAUNCH_CHECK();
    k_inv_sqrt_deg<<<grid_for(rows, 256), 256, 0, st>>>(g->by_src.deg, rows, g->dout_is, nullptr);
    CB_LAUNCH_CHECK();

    int h_flags[2] = {0, 0};
    CB_CUDA(cudaMemcpyAsync(h_flags, flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    CB_REQUIRE(h_flags[0] == 0, CB_E_RANGE, "edge_index holds a node id outside [0, num_nodes)");
    g->has_zero_in_deg = h_flags[1];

    // stable LSD radix sort of (key, edge id); `rows` is the sentinel for edges of other slices
    const int end_bit = bits_for((uint32_t)rows);
    size_t cub_bytes = 0;
    {
        cub::DoubleBuffer<uint32_t> kb(key_dst, key_alt);
        cub::DoubleBuffer<int32_t> vb(eid_a, eid_alt);
        CB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, kb, vb, (int)E, 0, end_bit, st));
    }
    void* cub_tmp = nullptr;
    CB_CUDA(tmp.alloc((char**)&cub_tmp, (int64_t)cub_bytes));

    for (int which = 0; which < 2; ++which) {
        Side& side = which == 0 ? g->by_dst : g->by_src;
        cub::DoubleBuffer<uint32_t> kb(which == 0 ? key_dst : key_src, key_alt);
        cub::DoubleBuffer<int32_t> vb(which == 0 ? eid_a : eid_b, eid_alt);
        if (E > 0) {
            CB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, kb, vb, (int)E, 0, end_bit, st));
            count_launch(2 * ((end_bit + 7) / 8));
        }
        // the owned edges are the first sum(deg) entries of the sorted list; keep exactly those
        int rc = CB_OK;
        // temporarily point perm at the sorted buffer; finish_side reads it to gather columns
        int32_t* sorted = vb.Current();
        side.perm = sorted;
        rc = finish_side(side, which == 0 ? src : dst, rows, g->hub_chunk, spine, st);
        if (rc) {
            side.perm = nullptr;
            return rc;
        }
        int32_t* keep = nullptr;
        cudaError_t e = cudaMalloc((void**)&keep, (size_t)(side.n_edges > 0 ? side.n_edges : 1) * sizeof(int32_t));
        if (e != cudaSuccess) {
            side.perm = nullptr;
            return cuda_fail(e, "cudaMalloc(perm)", __FILE__, __LINE__);
        }
        side.perm = keep;
        CB_CUDA(cudaMemcpyAsync(keep, sorted, (size_t)side.n_edges * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        CB_CUDA(cudaStreamSynchronize(st));
    }
    return CB_OK;
}

// ---------------------------------------------------------------------------------------------
// live-column compaction (row-sparse gathers: the gradient under a loss over the train rows only)
// ---------------------------------------------------------------------------------------------
static inline int64_t align256(int64_t b) { return (b + 255) & ~(int64_t)255; }

LiveView live_view(const cb_graph* g, int side_id, void* workspace) {
    const Side& s = side_id == CB_BY_DST ? g->by_dst : g->by_src;
    const int64_t groups = ceil_div(s.n_edges > 0 ? s.n_edges : 1, 32);
    const int64_t tiles = ceil_div(s.n_edges > 0 ? s.n_edges : 1, LIVE_TILE);
    char* p = (char*)workspace;
    int64_t o = 0;
    LiveView v{};
    v.rowptr = (int64_t*)(p + o);    o += align256((g->rows + 1) * (int64_t)sizeof(int64_t));
    v.chunk_beg = (int64_t*)(p + o); o += align256((s.n_chunks + 1) * (int64_t)sizeof(int64_t));
    v.chunk_end = (int64_t*)(p + o); o += align256((s.n_chunks + 1) * (int64_t)sizeof(int64_t));
    v.col = (int32_t*)(p + o);       o += align256((s.n_edges + 1) * (int64_t)sizeof(int32_t));
    v.bits = (uint32_t*)(p + o);     o += align256(groups * (int64_t)sizeof(uint32_t));
    v.posw = (int32_t*)(p + o);      o += align256(groups * (int64_t)sizeof(int32_t));
    v.spine = (int32_t*)(p + o);     o += align256((tiles + 1) * (int64_t)sizeof(int32_t));
    v.row_be = (int2*)(p + o);       o += align256((g->rows + 1) * (int64_t)sizeof(int2));
    v.bytes = o;
    return v;
}

// pass 1: one live bit per stored edge (ballot of live[col[j]] over 32 consecutive edges), live count per tile
__global__ void __launch_bounds__(256) k_live_bits(const int32_t* __restrict__ col, int64_t E,
                                                   const uint8_t* __restrict__ live, uint32_t* __restrict__ bits,
                                                   int32_t* __restrict__ tile_count) {
    __shared__ int wsum[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t g0 = (int64_t)blockIdx.x * (LIVE_TILE / 32) + w * (LIVE_TILE / 32 / 8);
    int cnt = 0;
#pragma unroll 4
    for (int k = 0; k < LIVE_TILE / 32 / 8; ++k) {
        const int64_t j = (g0 + k) * 32 + lane;
        const bool on = j < E && __ldg(live + __ldg(col + j)) != 0;
        const uint32_t b = __ballot_sync(0xffffffffu, on);
        if (lane == 0 && (g0 + k) * 32 < E) bits[g0 + k] = b;
        cnt += __popc(b);
    }
    if (lane == 0) wsum[w] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += wsum[i];
        tile_count[blockIdx.x] = t;
    }
}

// exclusive prefix of the tile counts in place (one block), grand total into spine[n]
__global__ void __launch_bounds__(1024) k_live_spine(int32_t* __restrict__ spine, int64_t n) {
    __shared__ int wtot[32];
    __shared__ int carry_s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t b = 0; b < n; b += 1024) {
        const int64_t i = b + threadIdx.x;
        const int v = i < n ? spine[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wtot[w] = inc;
        __syncthreads();
        int base = carry_s;
        for (int k = 0; k < w; ++k) base += wtot[k];
        if (i < n) spine[i] = base + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = base + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) spine[n] = carry_s;
}

// pass 2: position of every 32-edge group in the compacted list, and the compacted column ids themselves
__global__ void __launch_bounds__(256) k_live_fill(const int32_t* __restrict__ col, int64_t E,
                                                   const uint32_t* __restrict__ bits, const int32_t* __restrict__ spine,
                                                   int32_t* __restrict__ posw, int32_t* __restrict__ col_c) {
    constexpr int GPW = LIVE_TILE / 32 / 8;   // groups per warp: 16
    __shared__ int wsum[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t g0 = (int64_t)blockIdx.x * (LIVE_TILE / 32) + w * GPW;
    const int64_t n_groups = (E + 31) >> 5;
    // lane k < GPW holds the bit word of the warp's k-th group
    const uint32_t mine = (lane < GPW && g0 + lane < n_groups) ? __ldg(bits + g0 + lane) : 0u;
    int inc = __popc(mine);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    const int excl = inc - __popc(mine);
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    int base = __ldg(spine + blockIdx.x);
    for (int k = 0; k < w; ++k) base += wsum[k];
    if (lane < GPW && g0 + lane < n_groups) posw[g0 + lane] = base + excl;
#pragma unroll 4
    for (int k = 0; k < GPW; ++k) {
        const uint32_t b = __shfl_sync(0xffffffffu, mine, k);
        const int off = __shfl_sync(0xffffffffu, excl, k);
        if ((b >> lane) & 1u) {
            const int64_t j = (g0 + k) * 32 + lane;
            col_c[base + off + __popc(b & ((1u << lane) - 1u))] = __ldg(col + j);
        }
    }
}

__device__ __forceinline__ int64_t live_pos(int64_t e, int64_t E, const uint32_t* __restrict__ bits,
                                            const int32_t* __restrict__ posw, int32_t total) {
    if (e >= E) return total;
    const uint32_t b = __ldg(bits + (e >> 5));
    return (int64_t)__ldg(posw + (e >> 5)) + __popc(b & ((1u << (e & 31)) - 1u));
}

// pass 3: row offsets and hub-chunk bounds of the compacted list
__global__ void __launch_bounds__(256) k_live_offsets(const int64_t* __restrict__ rowptr, int64_t rows, int64_t E,
                                                      const int32_t* __restrict__ chunk_row,
                                                      const int64_t* __restrict__ chunk_beg, int64_t n_chunks,
                                                      int hub_chunk, const uint32_t* __restrict__ bits,
                                                      const int32_t* __restrict__ posw, const int32_t* __restrict__ spine,
                                                      int64_t n_tiles, int64_t* __restrict__ rowptr_c,
                                                      int64_t* __restrict__ chunk_beg_c, int64_t* __restrict__ chunk_end_c,
                                                      int2* __restrict__ row_be) {
    const int32_t total = __ldg(spine + n_tiles);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= rows + n_chunks; i += stride) {
        if (i <= rows) {
            const int64_t b0 = __ldg(rowptr + i);
            const int64_t pb = live_pos(b0, E, bits, posw, total);
            rowptr_c[i] = pb;
            if (i < rows) {
                const int64_t e0 = __ldg(rowptr + i + 1);
                const bool hub = e0 - b0 > hub_chunk;       // a row is a hub by its ORIGINAL degree
                row_be[i] = make_int2((int)pb, hub ? (int)pb : (int)live_pos(e0, E, bits, posw, total));
            }
        } else {
            const int64_t c = i - rows - 1;
            const int64_t b = __ldg(chunk_beg + c);
            const int64_t rend = __ldg(rowptr + __ldg(chunk_row + c) + 1);
            const int64_t e = b + hub_chunk < rend ? b + hub_chunk : rend;
            chunk_beg_c[c] = live_pos(b, E, bits, posw, total);
            chunk_end_c[c] = live_pos(e, E, bits, posw, total);
        }
    }
}

// One CSR side from an edge list that holds exactly the edges of the owned rows on that side.
static int build_side_local(cb_graph* g, Side& side, const int64_t* own_end, const int64_t* other_end, int64_t E,
                            float* deg_is, int* zero_flag_host, cudaStream_t st, int filter = 0) {
    const int64_t rows = g->rows;
    const int panels = g->src_panels;
    Scratch tmp;
    int32_t* deg_exp = nullptr;
    if (panels > 1) {
        CB_CUDA(tmp.alloc(&deg_exp, rows * panels));
        CB_CUDA(cudaMemsetAsync(deg_exp, 0, (size_t)(rows * panels > 0 ? rows * panels : 1) * sizeof(int32_t), st));
    }
    CB_CUDA(cudaMalloc((void**)&side.deg, (size_t)(rows > 0 ? rows : 1) * sizeof(int32_t)));
    CB_CUDA(cudaMemsetAsync(side.deg, 0, (size_t)(rows > 0 ? rows : 1) * sizeof(int32_t), st));
    uint32_t *key = nullptr, *key_alt = nullptr;
    int32_t *eid = nullptr, *eid_alt = nullptr;
    int* flags = nullptr;     // [0] range error, [1] zero degree, [2] edge of a row that is not owned
    int64_t* spine = nullptr;
    CB_CUDA(tmp.alloc(&key, E));
    CB_CUDA(tmp.alloc(&key_alt, E));
    CB_CUDA(tmp.alloc(&eid, E));
    CB_CUDA(tmp.alloc(&eid_alt, E));
    CB_CUDA(tmp.alloc(&flags, 3));
    CB_CUDA(tmp.alloc(&spine, ceil_div(rows * panels > 0 ? rows * panels : 1, SCAN_TILE) + 1));
    CB_CUDA(cudaMemsetAsync(flags, 0, 3 * sizeof(int), st));
    if (E > 0) {
        k_edge_keys_side<<<grid_for(E, 256), 256, 0, st>>>(own_end, other_end, E, g->n_nodes, g->row_begin, g->row_end,
                                                           panels, filter, key, eid, side.deg, deg_exp, flags);
        CB_LAUNCH_CHECK();
    }
    if (panels > 1) {
        CB_CUDA(cudaMalloc((void**)&side.rowptr_exp, (size_t)(rows * panels + 1) * sizeof(int64_t)));
        int rc_ = exclusive_scan(rows * panels, DegMap{deg_exp}, side.rowptr_exp, spine, st);
        if (rc_) return rc_;
    }
    k_inv_sqrt_deg<<<grid_for(rows, 256), 256, 0, st>>>(side.deg, rows, deg_is, flags + 1);
    CB_LAUNCH_CHECK();
    int h_flags[3] = {0, 0, 0};
    CB_CUDA(cudaMemcpyAsync(h_flags, flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    CB_REQUIRE(h_flags[0] == 0, CB_E_RANGE, "edge list holds a node id outside [0, num_nodes)");
    CB_REQUIRE(h_flags[2] == 0, CB_E_RANGE, "cb_graph_create_local: an edge does not belong to the owned row range");
    if (zero_flag_host) *zero_flag_host = h_flags[1];
    const int end_bit = bits_for((uint32_t)(rows * panels));
    cub::DoubleBuffer<uint32_t> kb(key, key_alt);
    cub::DoubleBuffer<int32_t> vb(eid, eid_alt);
    if (E > 0) {
        size_t cub_bytes = 0;
        CB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, kb, vb, (int)E, 0, end_bit, st));
        void* cub_tmp = nullptr;
        CB_CUDA(tmp.alloc((char**)&cub_tmp, (int64_t)cub_bytes));
        CB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, kb, vb, (int)E, 0, end_bit, st));
        count_launch(2 * ((end_bit + 7) / 8));
    }
    int32_t* sorted = vb.Current();
    side.perm = sorted;
    int rc = finish_side(side, other_end, rows, g->hub_chunk, spine, st);
    if (rc) {
        side.perm = nullptr;
        return rc;
    }
    int32_t* keep = nullptr;
    cudaError_t e = cudaMalloc((void**)&keep, (size_t)(side.n_edges > 0 ? side.n_edges : 1) * sizeof(int32_t));
    if (e != cudaSuccess) {
        side.perm = nullptr;
        return cuda_fail(e, "cudaMalloc(perm)", __FILE__, __LINE__);
    }
    side.perm = keep;
    CB_CUDA(cudaMemcpyAsync(keep, sorted, (size_t)side.n_edges * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    CB_CUDA(cudaStreamSynchronize(st));
    return CB_OK;
}

}  // namespace cb

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int cb_graph_create_sliced(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes,
                           int64_t row_begin, int64_t row_end, int hub_chunk, void* stream,
                           cb_graph_t** out) {
    using namespace cb;
    CB_REQUIRE(out != nullptr, CB_E_INVALID, "cb_graph_create: out is NULL");
    *out = nullptr;
    CB_REQUIRE(num_edges >= 0 && num_nodes >= 0, CB_E_INVALID, "cb_graph_create: negative size");
    CB_REQUIRE(num_edges == 0 || edge_index != nullptr, CB_E_INVALID, "cb_graph_create: edge_index is NULL");
    CB_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= num_nodes, CB_E_INVALID,
               "cb_graph_create: row range outside [0, num_nodes]");
    CB_REQUIRE(num_nodes < (int64_t)INT32_MAX && num_edges < (int64_t)INT32_MAX, CB_E_UNSUPPORTED,
               "cb_graph_create: num_nodes and num_edges must be < 2^31 per handle");
    cb_graph* g = new cb_graph();
    g->n_nodes = num_nodes;
    g->row_begin = row_begin;
    g->row_end = row_end;
    g->rows = row_end - row_begin;
    g->hub_chunk = hub_chunk > 0 ? hub_chunk : CB_DEFAULT_HUB_CHUNK;
    cudaError_t e = cudaGetDevice(&g->device);
    int rc = e == cudaSuccess ? build(g, edge_index, num_edges, (cudaStream_t)stream)
                              : cuda_fail(e, "cudaGetDevice", __FILE__, __LINE__);
    if (rc != CB_OK) {
        cb_graph_destroy(g);
        return rc;
    }
    *out = g;
    return CB_OK;
}

int cb_graph_create_local(const int64_t* in_edges, int64_t num_in_edges, const int64_t* out_edges,
                          int64_t num_out_edges, int64_t num_nodes, int64_t row_begin, int64_t row_end, int hub_chunk,
                          void* stream, cb_graph_t** out) {
    using namespace cb;
    CB_REQUIRE(out != nullptr, CB_E_INVALID, "cb_graph_create_local: out is NULL");
    *out = nullptr;
    CB_REQUIRE(num_in_edges >= 0 && num_out_edges >= 0 && num_nodes >= 0, CB_E_INVALID,
               "cb_graph_create_local: negative size");
    CB_REQUIRE((num_in_edges == 0 || in_edges) && (num_out_edges == 0 || out_edges), CB_E_INVALID,
               "cb_graph_create_local: an edge list is NULL");
    CB_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= num_nodes, CB_E_INVALID,
               "cb_graph_create_local: row range outside [0, num_nodes]");
    CB_REQUIRE(num_nodes < (int64_t)INT32_MAX && num_in_edges < (int64_t)INT32_MAX &&
                   num_out_edges < (int64_t)INT32_MAX,
               CB_E_UNSUPPORTED, "cb_graph_create_local: num_nodes and the per-handle edge counts must be < 2^31");
    cb_graph* g = new cb_graph();
    g->n_nodes = num_nodes;
    g->row_begin = row_begin;
    g->row_end = row_end;
    g->rows = row_end - row_begin;
    g->hub_chunk = hub_chunk > 0 ? hub_chunk : CB_DEFAULT_HUB_CHUNK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rows1 = g->rows > 0 ? g->rows : 1;
    int rc = CB_OK;
    cudaError_t e = cudaGetDevice(&g->device);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaGetDevice", __FILE__, __LINE__);
    if (!rc && (e = cudaMalloc((void**)&g->din_is, (size_t)rows1 * sizeof(float))) != cudaSuccess)
        rc = cuda_fail(e, "cudaMalloc(din)", __FILE__, __LINE__);
    if (!rc && (e = cudaMalloc((void**)&g->dout_is, (size_t)rows1 * sizeof(float))) != cudaSuccess)
        rc = cuda_fail(e, "cudaMalloc(dout)", __FILE__, __LINE__);
    // in-edges: rows = destinations (row 1 of the list), columns = sources (row 0)
    if (!rc) rc = build_side_local(g, g->by_dst, in_edges + num_in_edges, in_edges, num_in_edges, g->din_is,
                                   &g->has_zero_in_deg, st);
    // out-edges: rows = sources (row 0), columns = destinations (row 1)
    if (!rc) rc = build_side_local(g, g->by_src, out_edges, out_edges + num_out_edges, num_out_edges, g->dout_is,
                                   nullptr, st);
    if (rc != CB_OK) {
        cb_graph_destroy(g);
        return rc;
    }
    *out = g;
    return CB_OK;
}

int cb_graph_create_panelled(const int64_t* in_edges, int64_t num_in_edges, const int64_t* out_edges,
                             int64_t num_out_edges, int64_t num_nodes, int64_t row_begin, int64_t row_end,
                             int hub_chunk, int src_panels, int filter, void* stream, cb_graph_t** out) {
    using namespace cb;
    CB_REQUIRE(out != nullptr, CB_E_INVALID, "cb_graph_create_panelled: out is NULL");
    *out = nullptr;
    CB_REQUIRE(src_panels == 1 || src_panels == 2 || src_panels == 4, CB_E_INVALID,
               "cb_graph_create_panelled: src_panels must be 1, 2 or 4");
    CB_REQUIRE(num_in_edges >= 0 && num_out_edges >= 0 && num_nodes >= 0, CB_E_INVALID,
               "cb_graph_create_panelled: negative size");
    CB_REQUIRE((num_in_edges == 0 || in_edges) && (num_out_edges == 0 || out_edges), CB_E_INVALID,
               "cb_graph_create_panelled: an edge list is NULL");
    CB_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= num_nodes, CB_E_INVALID,
               "cb_graph_create_panelled: row range outside [0, num_nodes]");
    CB_REQUIRE(num_nodes < (int64_t)INT32_MAX && num_in_edges < (int64_t)INT32_MAX &&
                   num_out_edges < (int64_t)INT32_MAX && (row_end - row_begin) * src_panels < (int64_t)INT32_MAX,
               CB_E_UNSUPPORTED, "cb_graph_create_panelled: sizes must be < 2^31 per handle");
    cb_graph* g = new cb_graph();
    g->n_nodes = num_nodes;
    g->row_begin = row_begin;
    g->row_end = row_end;
    g->rows = row_end - row_begin;
    g->hub_chunk = hub_chunk > 0 ? hub_chunk : CB_DEFAULT_HUB_CHUNK;
    g->src_panels = src_panels;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rows1 = g->rows > 0 ? g->rows : 1;
    int rc = CB_OK;
    cudaError_t e = cudaGetDevice(&g->device);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaGetDevice", __FILE__, __LINE__);
    if (!rc && (e = cudaMalloc((void**)&g->din_is, (size_t)rows1 * sizeof(float))) != cudaSuccess)
        rc = cuda_fail(e, "cudaMalloc(din)", __FILE__, __LINE__);
    if (!rc && (e = cudaMalloc((void**)&g->dout_is, (size_t)rows1 * sizeof(float))) != cudaSuccess)
        rc = cuda_fail(e, "cudaMalloc(dout)", __FILE__, __LINE__);
    if (!rc) rc = build_side_local(g, g->by_dst, in_edges + num_in_edges, in_edges, num_in_edges, g->din_is,
                                   &g->has_zero_in_deg, st, filter);
    if (!rc) rc = build_side_local(g, g->by_src, out_edges, out_edges + num_out_edges, num_out_edges, g->dout_is,
                                   nullptr, st, filter);
    if (rc != CB_OK) {
        cb_graph_destroy(g);
        return rc;
    }
    *out = g;
    return CB_OK;
}

int cb_graph_create(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes, int hub_chunk,
                    void* stream, cb_graph_t** out) {
    return cb_graph_create_sliced(edge_index, num_edges, num_nodes, 0, num_nodes, hub_chunk, stream, out);
}

int64_t cb_graph_live_workspace_bytes(const cb_graph_t* g, int side) {
    if (!g || (side != CB_BY_DST && side != CB_BY_SRC)) return 0;
    return cb::live_view(g, side, nullptr).bytes;
}

int cb_graph_compact_live(const cb_graph_t* g, int side_id, const uint8_t* row_live, void* live_ws,
                          int64_t live_ws_bytes, void* stream) {
    using namespace cb;
    CB_REQUIRE(g != nullptr, CB_E_INVALID, "cb_graph_compact_live: graph is NULL");
    CB_REQUIRE(side_id == CB_BY_DST || side_id == CB_BY_SRC, CB_E_INVALID, "cb_graph_compact_live: unknown side");
    CB_REQUIRE(row_live != nullptr, CB_E_INVALID, "cb_graph_compact_live: row_live is NULL");
    CB_REQUIRE(live_ws != nullptr && live_ws_bytes >= cb_graph_live_workspace_bytes(g, side_id), CB_E_WORKSPACE,
               "cb_graph_compact_live: workspace smaller than cb_graph_live_workspace_bytes()");
    CB_REQUIRE((reinterpret_cast<uintptr_t>(live_ws) & 255u) == 0, CB_E_INVALID,
               "cb_graph_compact_live: workspace must be 256-byte aligned");
    const Side& s = side_id == CB_BY_DST ? g->by_dst : g->by_src;
    const LiveView v = live_view(g, side_id, live_ws);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t E = s.n_edges;
    const int64_t tiles = ceil_div(E > 0 ? E : 1, LIVE_TILE);
    k_live_bits<<<(unsigned)tiles, 256, 0, st>>>(s.col, E, row_live, v.bits, v.spine);
    CB_LAUNCH_CHECK();
    k_live_spine<<<1, 1024, 0, st>>>(v.spine, tiles);
    CB_LAUNCH_CHECK();
    k_live_fill<<<(unsigned)tiles, 256, 0, st>>>(s.col, E, v.bits, v.spine, v.posw, v.col);
    CB_LAUNCH_CHECK();
    k_live_offsets<<<grid_for(g->rows + 1 + s.n_chunks, 256), 256, 0, st>>>(
        s.rowptr, g->rows, E, s.chunk_row, s.chunk_beg, s.n_chunks, g->hub_chunk, v.bits, v.posw, v.spine, tiles,
        v.rowptr, v.chunk_beg, v.chunk_end, v.row_be);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_graph_destroy(cb_graph_t* g) {
    if (!g) return CB_OK;
    cb::free_side(g->by_dst);
    cb::free_side(g->by_src);
    cudaFree(g->din_is);
    cudaFree(g->dout_is);
    delete g;
    return CB_OK;
}

int cb_graph_query(const cb_graph_t* g, int what, void* out) {
    using namespace cb;
    CB_REQUIRE(g != nullptr && out != nullptr, CB_E_INVALID, "cb_graph_query: NULL argument");
    int64_t* i = (int64_t*)out;
    const void** p = (const void**)out;
    switch (what) {
        case CB_Q_NUM_NODES: *i = g->n_nodes; break;
        case CB_Q_NUM_EDGES: *i = g->by_dst.n_edges; break;
        case CB_Q_ROW_BEGIN: *i = g->row_begin; break;
        case CB_Q_ROW_END: *i = g->row_end; break;
        case CB_Q_HAS_ZERO_IN_DEG: *i = g->has_zero_in_deg; break;
        case CB_Q_HUB_CHUNK: *i = g->hub_chunk; break;
        case CB_Q_SRC_PANELS: *i = g->src_panels; break;
        case CB_Q_DST_ROWPTR_EXP: *p = g->by_dst.rowptr_exp; break;
        case CB_Q_SRC_ROWPTR_EXP: *p = g->by_src.rowptr_exp; break;
        case CB_Q_DST_ROWPTR: *p = g->by_dst.rowptr; break;
        case CB_Q_DST_COL: *p = g->by_dst.col; break;
        case CB_Q_DST_PERM: *p = g->by_dst.perm; break;
        case CB_Q_DST_NUM_HUB_CHUNKS: *i = g->by_dst.n_chunks; break;
        case CB_Q_SRC_ROWPTR: *p = g->by_src.rowptr; break;
        case CB_Q_SRC_COL: *p = g->by_src.col; break;
        case CB_Q_SRC_PERM: *p = g->by_src.perm; break;
        case CB_Q_SRC_NUM_HUB_CHUNKS: *i = g->by_src.n_chunks; break;
        case CB_Q_SRC_NUM_EDGES: *i = g->by_src.n_edges; break;
        case CB_Q_DIN_INV_SQRT: *p = g->din_is; break;
        case CB_Q_DOUT_INV_SQRT: *p = g->dout_is; break;
        case CB_Q_IN_DEGREE: *p = g->by_dst.deg; break;
        case CB_Q_OUT_DEGREE: *p = g->by_src.deg; break;
        default: set_error("cb_graph_query: unknown selector"); return CB_E_INVALID;
    }
    return CB_OK;
}

}  // extern "C"
