// Neighbour gather-and-reduce over a CSR side of the graph, with the GCN layer's epilogue fused in.
//
// Replaces DGL's gSpMM('copy_lhs','sum') reached from GNN_model/GCN.py:238
// (graph.update_all(fn.copy_src('h','m'), fn.sum('m','h'))) and, fused behind it, GCN.py:242-253
// (in-degree scale, bias), GCN.py:127-128 (relu), res_tricks.py:14/23 (residual mix) and the next
// layer's GCN.py:205-213 (out-degree scale).
//
// Work decomposition.  A "task" is one CSR row, or one hub chunk (a run of hub_chunk consecutive
// entries of a row longer than hub_chunk).  A group of LPR lanes owns a task; a warp holds 32/LPR
// groups.  Every lane owns NCH column slots of VEC floats (column = (ch*LPR + sub)*VEC), so a group
// reads a feature row with full-width coalesced 16-byte loads, and every (row, column) accumulator is
// a strictly in-order fp32 sum over the stored neighbours -- no atomics, no cross-lane reduction, the
// result is bit-stable from run to run and equal to the sequential sum.  Hub chunks write partial rows
// that k_combine adds in chunk order (same association as oracle/spmm_sum_csr.c with hub_chunk).
//
// The neighbour ids of a task are read LPR at a time with one coalesced load and handed round with
// warp shuffles; UNROLL independent row loads are in flight per group before the first add.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "cb_internal.cuh"

namespace cb {

struct AggArgs {
    const int64_t* rowptr;
    const int32_t* col;
    int64_t n_rows;
    const int32_t* chunk_row;
    const int64_t* chunk_beg;
    int64_t n_chunks;
    int hub_chunk;
    const void* X;       // [n_src, x_ld], fp32 or bf16 (the kernel's storage type S)
    int64_t d;           // columns aggregated (logical width)
    int64_t x_ld;        // row pitch of X, elements
    int64_t o_ld;        // row pitch of x0 / out / out2 / mask, elements
    int64_t col0;        // first column handled by this launch
    float* partial;      // [n_chunks, d]
    // epilogue
    const float* row_scale;   // [rows] or null
    const float* bias;        // [d] or null
    const void* x0;           // [rows, d] or null (storage type S)
    float alpha, one_minus_alpha;
    int act;
    void* out;                // [rows, d] or null (storage type S)
    const float* out2_scale;  // [rows]
    void* out2;               // [rows, d] or null (storage type S)
    uint8_t* mask;            // [rows, d] or null
    const uint8_t* live;      // [n_src] or null: rows of X with live[s] == 0 are all-zero and are not gathered
    const float* ew;          // [n_edges] or null: weight of every stored edge (GCN.py:199-202 u_mul_e): sum of x * w
    // live-column compacted lists (cb_graph_compact_live): rowptr / col / chunk_beg point into the compacted CSR,
    // hub_rowptr is the ORIGINAL rowptr (a row is a hub by its original degree, so that every partial sum keeps
    // the association of the uncompacted walk) and chunk_end bounds each chunk explicitly
    const int64_t* hub_rowptr;  // null: rowptr
    const int64_t* chunk_end;   // null: min(chunk_beg + hub_chunk, row end)
    // label-propagation epilogue (cb_agg_propagate): out = clamp(one_minus_alpha * r + alpha * x0, lo, hi)
    int clamp;
    float clamp_lo, clamp_hi;
    int64_t task0;              // first task of a k_agg launch (0, or n_rows when the rows went to k_gather_sparse_rows)
    // source-panel pass (cb_agg_*_pass on a graph built with src_panels S > 1): this launch walks only the group of
    // panel `panel` of every row and continues the row's in-order sum through `carry` (fp32 [rows, d])
    const int64_t* rowptr_exp;  // [rows*S+1] or null
    int n_panels, panel;
    float* carry;
    const int2* row_be;         // compacted plain gather: (begin, end) per row with hub rows emptied, else null
    int prefetch_x0;            // pull the row of x0 into L2 before the neighbour walk (CB_AGG_PREFETCH=0: A/B switch)
};

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
    using T = float4;
    static __device__ __forceinline__ void load(float (&v)[4], const float* p) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void load_plain(float (&v)[4], const float* p) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
    static __device__ __forceinline__ void store_mask(uint8_t* p, const float (&z)[4]) {
        uchar4 m;
        m.x = z[0] > 0.f; m.y = z[1] > 0.f; m.z = z[2] > 0.f; m.w = z[3] > 0.f;
        *reinterpret_cast<uchar4*>(p) = m;
    }
};
template <>
struct Vec<1> {
    using T = float;
    static __device__ __forceinline__ void load(float (&v)[1], const float* p) { v[0] = __ldg(p); }
    static __device__ __forceinline__ void load_plain(float (&v)[1], const float* p) { v[0] = *p; }
    static __device__ __forceinline__ void store(float* p, const float (&v)[1]) { *p = v[0]; }
    static __device__ __forceinline__ void store_mask(uint8_t* p, const float (&z)[1]) { *p = z[0] > 0.f; }
};

template <>
struct Vec<8> {   // eight fp32 values (bias, hub partials) beside a bf16x8 feature access
    static __device__ __forceinline__ void load(float (&v)[8], const float* p) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void load_plain(float (&v)[8], const float* p) {
        const float4 a = *reinterpret_cast<const float4*>(p), b = *(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
        reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    static __device__ __forceinline__ void store_mask(uint8_t* p, const float (&z)[8]) {
        uint2 m;
        m.x = (z[0] > 0.f ? 1u : 0u) | (z[1] > 0.f ? 0x100u : 0u) | (z[2] > 0.f ? 0x10000u : 0u) | (z[3] > 0.f ? 0x1000000u : 0u);
        m.y = (z[4] > 0.f ? 1u : 0u) | (z[5] > 0.f ? 0x100u : 0u) | (z[6] > 0.f ? 0x10000u : 0u) | (z[7] > 0.f ? 0x1000000u : 0u);
        *reinterpret_cast<uint2*>(p) = m;
    }
};

// Feature-matrix access in the storage type S: fp32 (VEC = 4 or 1 per access) or bf16 (VEC = 8 or 1).  bf16 rows
// are widened to fp32 on load -- every sum and the whole epilogue stay fp32 -- and rounded to nearest-even on store.
template <typename S, int VEC>
struct Elem;
template <int VEC>
struct Elem<float, VEC> {
    struct Raw { float f[VEC]; };
    static __device__ __forceinline__ void load_raw(Raw& r, const void* base, int64_t off) {
        Vec<VEC>::load(r.f, reinterpret_cast<const float*>(base) + off);
    }
    static __device__ __forceinline__ void add(float (&acc)[VEC], const Raw& r) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] += r.f[i];
    }
    // u_mul_e: the product is rounded before it is added (no FMA), like DGL's `out += lhs * rhs` on the CPU
    static __device__ __forceinline__ void add_scaled(float (&acc)[VEC], const Raw& r, float w) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = __fadd_rn(acc[i], __fmul_rn(r.f[i], w));
    }
    static __device__ __forceinline__ void load(float (&v)[VEC], const void* base, int64_t off) {
        Vec<VEC>::load(v, reinterpret_cast<const float*>(base) + off);
    }
    static __device__ __forceinline__ void store(void* base, int64_t off, const float (&v)[VEC]) {
        Vec<VEC>::store(reinterpret_cast<float*>(base) + off, v);
    }
};
template <>
struct Elem<__nv_bfloat16, 8> {
    // gathered rows stay packed (4 registers per 16-byte load) until they are added, so that eight loads per lane
    // can be in flight
    struct Raw { uint4 u; };
    static __device__ __forceinline__ void load_raw(Raw& r, const void* base, int64_t off) {
        r.u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + off));
    }
    static __device__ __forceinline__ void widen(float (&v)[8], const uint4& t) {
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void add(float (&acc)[8], const Raw& r) {
        float v[8];
        widen(v, r.u);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
    static __device__ __forceinline__ void add_scaled(float (&acc)[8], const Raw& r, float w) {
        float v[8];
        widen(v, r.u);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = __fadd_rn(acc[i], __fmul_rn(v[i], w));
    }
    static __device__ __forceinline__ void load(float (&v)[8], const void* base, int64_t off) {
        widen(v, __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + off)));
    }
    static __device__ __forceinline__ void store(void* base, int64_t off, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + off) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};
template <>
struct Elem<__nv_bfloat16, 1> {
    struct Raw { float f; };
    static __device__ __forceinline__ void load_raw(Raw& r, const void* base, int64_t off) {
        r.f = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[off]);
    }
    static __device__ __forceinline__ void add(float (&acc)[1], const Raw& r) { acc[0] += r.f; }
    static __device__ __forceinline__ void add_scaled(float (&acc)[1], const Raw& r, float w) {
        acc[0] = __fadd_rn(acc[0], __fmul_rn(r.f, w));
    }
    static __device__ __forceinline__ void load(float (&v)[1], const void* base, int64_t off) {
        v[0] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[off]);
    }
    static __device__ __forceinline__ void store(void* base, int64_t off, const float (&v)[1]) {
        reinterpret_cast<__nv_bfloat16*>(base)[off] = __float2bfloat16_rn(v[0]);
    }
};

// z = rs*acc + b ; r = act(z) ; out = (1-a) r + a x0 ; out2 = s2 * out ; mask = z > 0
// Every product and sum is rounded separately (no FMA contraction), like the reference's chain of
// elementwise ops (GCN.py:250,253; res_tricks.py:14,23).
template <typename S, int VEC>
__device__ __forceinline__ void epilogue_store(const AggArgs& a, int64_t row, int64_t c, float (&acc)[VEC]) {
    const int64_t off = row * a.o_ld + c;
    float z[VEC], o[VEC];
    const float rs = a.row_scale ? __ldg(a.row_scale + row) : 1.f;
    float b[VEC], x0[VEC];
    if (a.bias) Vec<VEC>::load(b, a.bias + c);
    if (a.x0) Elem<S, VEC>::load(x0, a.x0, off);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        float t = a.row_scale ? __fmul_rn(acc[i], rs) : acc[i];
        if (a.bias) t = __fadd_rn(t, b[i]);
        z[i] = t;
        float r = (a.act == CB_ACT_RELU) ? fmaxf(t, 0.f) : t;
        if (a.x0) r = __fadd_rn(__fmul_rn(a.one_minus_alpha, r), __fmul_rn(a.alpha, x0[i]));
        if (a.clamp) r = fminf(fmaxf(r, a.clamp_lo), a.clamp_hi);
        o[i] = r;
    }
    if (a.out) Elem<S, VEC>::store(a.out, off, o);
    if (a.out2) {
        const float s2 = __ldg(a.out2_scale + row);
        float o2[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) o2[i] = __fmul_rn(o[i], s2);
        Elem<S, VEC>::store(a.out2, off, o2);
    }
    if (a.mask) Vec<VEC>::store_mask(a.mask + off, z);
}

// LIVE: X is row-sparse (e.g. the gradient arriving from a loss over the train rows only); a.live says which
// source rows can be non-zero.  Skipping an all-zero row leaves every fp32 sum unchanged (x + 0 = x).
// MODE 0: plain walk.  MODE 1 (LIVE): X is row-sparse, see above.  MODE 2 (EW): every stored edge carries a weight
// (a.ew, in stored order): out = sum of x * w, the product rounded before the in-order add.
// Register budget: 64 per thread (32 resident warps per SM) for everything up to two column slots per lane.  Without
// the bound the bf16 instantiations drifted from 64 to 76 registers as the argument block grew, which cost a quarter of
// the resident warps and 35 % of their throughput (caught by scripts/agg_bf16_bench.py against its round-1 numbers).
template <typename S, int VEC, int LPR, int NCH, int UNROLL, int MODE>
__global__ void __launch_bounds__(256, (NCH >= 4 ? 2 : 4)) k_agg(const AggArgs a) {
    constexpr bool LIVE = MODE == 1, EW = MODE == 2;
    constexpr int GROUPS = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR;
    const int grp = lane / LPR;
    const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (grp * LPR));
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t task = a.task0 + warp * GROUPS + grp;      // task0 = n_rows: hub chunks only
    if (task >= a.n_rows + a.n_chunks) return;

    const bool is_chunk = task >= a.n_rows;
    int64_t row, beg, end;
    if (!is_chunk) {
        row = task;
        if (a.rowptr_exp) {
            const int64_t* re = a.rowptr_exp + row * a.n_panels;
            // hub row (by its whole degree): its chunks and k_combine produce it, in the last pass
            if (__ldg(re + a.n_panels) - __ldg(re) > a.hub_chunk) return;
            beg = __ldg(re + a.panel);
            end = __ldg(re + a.panel + 1);
        } else {
            beg = __ldg(a.rowptr + row);
            end = __ldg(a.rowptr + row + 1);
            // hub row: its chunks and k_combine produce it
            if (a.hub_rowptr) {
                if (__ldg(a.hub_rowptr + row + 1) - __ldg(a.hub_rowptr + row) > a.hub_chunk) return;
            } else if (end - beg > a.hub_chunk) {
                return;
            }
        }
    } else {
        const int64_t c = task - a.n_rows;
        row = __ldg(a.chunk_row + c);
        beg = __ldg(a.chunk_beg + c);
        if (a.chunk_end) {
            end = __ldg(a.chunk_end + c);
        } else {
            const int64_t rend = __ldg(a.rowptr + row + 1);
            end = beg + a.hub_chunk < rend ? beg + a.hub_chunk : rend;
        }
    }

    float acc[NCH][VEC];
    int64_t cofs[NCH];
    bool cval[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        cofs[ch] = a.col0 + (int64_t)(ch * LPR + sub) * VEC;
        cval[ch] = cofs[ch] < a.d;
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[ch][i] = 0.f;
    }
    const bool pass_row = a.rowptr_exp != nullptr && !is_chunk;
    if (a.prefetch_x0 && a.x0 && !is_chunk) {
        // the epilogue's x0 row is the last dependent DRAM access of a row: request it now, read it from L2 later
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
            if (cval[ch])
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const S*>(a.x0) + row * a.o_ld + cofs[ch]));
    }
    if (pass_row && a.panel > 0) {       // continue the in-order sum of the earlier panels
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
            if (cval[ch]) Vec<VEC>::load_plain(acc[ch], a.carry + row * a.d + cofs[ch]);
    }

    for (int64_t base = beg; base < end; base += LPR) {
        const int n = (int)(end - base < LPR ? end - base : LPR);
        int my = sub < n ? __ldg(a.col + base + sub) : 0;
        if (LIVE && sub < n && __ldg(a.live + my) == 0) my |= (int)0x80000000;   // N < 2^31: the sign bit is free
        float myw = 0.f;
        if (EW && sub < n) myw = __ldg(a.ew + base + sub);
        for (int k = 0; k < n; k += UNROLL) {
            typename Elem<S, VEC>::Raw v[UNROLL][NCH];
            bool on[UNROLL];
            float wv[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int s = __shfl_sync(gmask, my, (k + u) & (LPR - 1), LPR);
                if (EW) wv[u] = __shfl_sync(gmask, myw, (k + u) & (LPR - 1), LPR);
                on[u] = (k + u < n) && (!LIVE || s >= 0);
                if (on[u]) {
                    const int64_t xr = (int64_t)s * a.x_ld;
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch)
                        if (cval[ch]) Elem<S, VEC>::load_raw(v[u][ch], a.X, xr + cofs[ch]);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (on[u]) {
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) {
                        if (cval[ch]) {
                            if constexpr (EW) Elem<S, VEC>::add_scaled(acc[ch], v[u][ch], wv[u]);
                            else Elem<S, VEC>::add(acc[ch], v[u][ch]);
                        }
                    }
                }
            }
        }
    }

    if (is_chunk) {
        float* pr = a.partial + (task - a.n_rows) * a.d;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
            if (cval[ch]) Vec<VEC>::store(pr + cofs[ch], acc[ch]);
    } else if (pass_row && a.panel < a.n_panels - 1) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
            if (cval[ch]) Vec<VEC>::store(a.carry + row * a.d + cofs[ch], acc[ch]);
    } else {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
            if (cval[ch]) epilogue_store<S, VEC>(a, row, cofs[ch], acc[ch]);
    }
}

// Row-sparse gathers over the live-column compacted lists: most rows keep 0-3 neighbours, so the per-row overhead is
// the whole cost (ncu, 10 % live rows at the bench shape: the general kernel executes ~310 warp instructions per row
// and sits at 44 % issue utilisation, 3 TB/s).  This kernel does only what a compacted gather needs: one 8-byte
// (begin, end) per row -- hub rows are empty there, their chunks and k_combine produce them afterwards -- the column
// ids, the rows, one scaled store.  A row's sum is the in-order sum of its live neighbours (bit-identical to k_agg).
template <typename S, int VEC, int LPR, int NCH>
__global__ void __launch_bounds__(256) k_gather_sparse_rows(const AggArgs a, const int2* __restrict__ row_be) {
    constexpr int GROUPS = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR;
    const int64_t row = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GROUPS + lane / LPR;
    if (row >= a.n_rows) return;
    const int2 be = __ldg(row_be + row);
    float acc[NCH][VEC];
    int64_t cofs[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        cofs[ch] = a.col0 + (int64_t)(ch * LPR + sub) * VEC;
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[ch][i] = 0.f;
    }
    for (int e = be.x; e < be.y; e += 2) {          // two rows in flight per group
        const int s0 = __ldg(a.col + e);
        const bool two = e + 1 < be.y;
        const int s1 = two ? __ldg(a.col + e + 1) : s0;
        typename Elem<S, VEC>::Raw v0[NCH], v1[NCH];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            if (cofs[ch] < a.d) {
                Elem<S, VEC>::load_raw(v0[ch], a.X, (int64_t)s0 * a.x_ld + cofs[ch]);
                if (two) Elem<S, VEC>::load_raw(v1[ch], a.X, (int64_t)s1 * a.x_ld + cofs[ch]);
            }
        }
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            if (cofs[ch] < a.d) {
                Elem<S, VEC>::add(acc[ch], v0[ch]);
                if (two) Elem<S, VEC>::add(acc[ch], v1[ch]);
            }
        }
    }
    const float rs = a.row_scale ? __ldg(a.row_scale + row) : 1.f;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        if (cofs[ch] < a.d) {
            float o[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) o[i] = a.row_scale ? __fmul_rn(acc[ch][i], rs) : acc[ch][i];
            Elem<S, VEC>::store(a.out, row * a.o_ld + cofs[ch], o);
        }
    }
}

// Adds the chunk partials of every hub row in chunk order and applies the epilogue.
template <typename S, int VEC, int LPR, int NCH>
__global__ void __launch_bounds__(256) k_combine(const AggArgs a) {
    constexpr int GROUPS = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR;
    const int grp = lane / LPR;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t c = warp * GROUPS + grp;
    if (c >= a.n_chunks) return;
    const int64_t row = a.chunk_row[c];
    if (c > 0 && a.chunk_row[c - 1] == row) return;  // only the first chunk of a row combines
    const int64_t* rp = a.hub_rowptr ? a.hub_rowptr : a.rowptr;
    const int64_t deg = rp[row + 1] - rp[row];
    const int64_t n = (deg + a.hub_chunk - 1) / a.hub_chunk;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        const int64_t co = a.col0 + (int64_t)(ch * LPR + sub) * VEC;
        if (co >= a.d) continue;
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
        for (int64_t k = 0; k < n; ++k) {
            float v[VEC];
            Vec<VEC>::load_plain(v, a.partial + (c + k) * a.d + co);
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] += v[i];
        }
        epilogue_store<S, VEC>(a, row, co, acc);
    }
}

template <typename S, int VEC, int LPR, int NCH>
static int launch_cfg(const AggArgs& a_in, cudaStream_t st) {
    constexpr int GROUPS = 32 / LPR;
    // Warps per block.  A block keeps its slot until its longest row is done, and on a power-law graph the longest of
    // 8 rows is several times the mean: with 8-warp blocks sm__warps_active was 35 % of a 50 % limit.  Measured at the
    // bench shape (profiles/r02y_*): fused forward 20.4 / 18.9 / 18.6 ms and transposed gather 17.5 / 16.6 / 16.4 ms
    // with 8 / 4 / 2 warps per block.  CB_AGG_WARPS overrides (A/B switch).
    static const int warps_env = getenv("CB_AGG_WARPS") ? atoi(getenv("CB_AGG_WARPS")) : 0;
    const int WARPS = (warps_env == 1 || warps_env == 2 || warps_env == 4 || warps_env == 8) ? warps_env : 2;
    constexpr int UNROLL = NCH >= 2 ? 4 : 8;   // >= 128 bytes of gathered rows in flight per lane
    constexpr int U = UNROLL < LPR ? UNROLL : LPR;
    AggArgs a = a_in;
    static const bool lean_rows = !(getenv("CB_SPARSE_LEAN") && atoi(getenv("CB_SPARSE_LEAN")) == 0);   // A/B switch
    if (lean_rows && a.row_be != nullptr && a.n_rows > 0 && LPR >= 16) {
        // compacted (row-sparse) plain gather: the lean per-row kernel, then the hub chunks through k_agg / k_combine
        const int64_t blocks = ceil_div(a.n_rows, (int64_t)WARPS * GROUPS);
        CB_REQUIRE(blocks < (int64_t)INT32_MAX, CB_E_UNSUPPORTED, "aggregation grid too large");
        k_gather_sparse_rows<S, VEC, LPR, NCH><<<(unsigned)blocks, WARPS * 32, 0, st>>>(a, a.row_be);
        CB_LAUNCH_CHECK();
        a.task0 = a.n_rows;      // k_agg below: the hub chunks only
    }
    if (a.rowptr_exp && a.panel < a.n_panels - 1) a.n_chunks = 0;    // hub rows: whole, in the last pass
    const int64_t tasks = a.n_rows + a.n_chunks - a.task0;
    if (tasks > 0) {
        const int64_t blocks = ceil_div(tasks, (int64_t)WARPS * GROUPS);
        CB_REQUIRE(blocks < (int64_t)INT32_MAX, CB_E_UNSUPPORTED, "aggregation grid too large");
        if (a.live)
            k_agg<S, VEC, LPR, NCH, U, 1><<<(unsigned)blocks, WARPS * 32, 0, st>>>(a);
        else if (a.ew)
            k_agg<S, VEC, LPR, NCH, U, 2><<<(unsigned)blocks, WARPS * 32, 0, st>>>(a);
        else
            k_agg<S, VEC, LPR, NCH, U, 0><<<(unsigned)blocks, WARPS * 32, 0, st>>>(a);
        CB_LAUNCH_CHECK();
    }
    if (a.n_chunks > 0) {
        const int64_t blocks = ceil_div(a.n_chunks, (int64_t)WARPS * GROUPS);
        k_combine<S, VEC, LPR, NCH><<<(unsigned)blocks, WARPS * 32, 0, st>>>(a);
        CB_LAUNCH_CHECK();
    }
    return CB_OK;
}


// ---------------------------------------------------------------------------------------------
// Experiment (round-1 verdict 3d, the north_star's wording): gathered rows staged through shared memory by the bulk-copy
// engine.  One warp per task as in k_agg; an elected lane issues one cp.async.bulk (global -> shared, a whole feature
// row, completion counted on an mbarrier) per neighbour into a per-warp ring of R row slots, the warp adds the slots in
// stored order -- the same in-order fp32 sums, bit-identical outputs.  R rows of a task are in flight without holding
// them in registers.  Selected with CB_AGG_BULK=R (4 or 8) for full-width fp32 / bf16 rows; measured, see DESIGN.md 9.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if (spin > (1u << 24)) __trap();      // a pipeline bug traps instead of hanging the device
    }
}

template <typename S, int VEC, int NCH, int R>
__global__ void __launch_bounds__(256) k_agg_bulk(const AggArgs a) {
    extern __shared__ __align__(128) unsigned char bulk_smem[];
    constexpr int LPR = 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
    const uint32_t row_bytes = (uint32_t)(a.d * sizeof(S));
    unsigned char* ring = bulk_smem + (size_t)w * R * row_bytes;
    const uint32_t bars = smem_addr(bulk_smem + (size_t)W * R * row_bytes) + (uint32_t)w * R * 8u;
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + 8u * r), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int64_t task = a.task0 + (int64_t)blockIdx.x * W + w;
    if (task >= a.n_rows + a.n_chunks) return;
    const bool is_chunk = task >= a.n_rows;
    int64_t row, beg, end;
    if (!is_chunk) {
        row = task;
        beg = __ldg(a.rowptr + row);
        end = __ldg(a.rowptr + row + 1);
        if (end - beg > a.hub_chunk) return;      // hub row: its chunks and k_combine produce it
    } else {
        const int64_t c = task - a.n_rows;
        row = __ldg(a.chunk_row + c);
        beg = __ldg(a.chunk_beg + c);
        const int64_t rend = __ldg(a.rowptr + row + 1);
        end = beg + a.hub_chunk < rend ? beg + a.hub_chunk : rend;
    }
    float acc[NCH][VEC];
    int64_t cofs[NCH];
    bool cval[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        cofs[ch] = (int64_t)(ch * LPR + lane) * VEC;
        cval[ch] = cofs[ch] < a.d;
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[ch][i] = 0.f;
    }
    if (a.prefetch_x0 && a.x0 && !is_chunk) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
            if (cval[ch])
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const S*>(a.x0) + row * a.o_ld + cofs[ch]));
    }
    uint32_t phases = 0;      // bit r: parity the next wait on slot r expects
    uint32_t head = 0;        // slot of the next row to consume
    for (int64_t base = beg; base < end; base += 32) {
        const int nb = (int)(end - base < 32 ? end - base : 32);
        const int my = lane < nb ? __ldg(a.col + base + lane) : 0;
        int issued = 0;
        auto issue = [&](int j) {
            const int s = __shfl_sync(0xffffffffu, my, j);
            if (lane == 0) {
                const uint32_t slot = (head + (uint32_t)j) % R;      // consumption k of this batch uses (head + k) % R
                const uint32_t bar = bars + 8u * slot;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes) : "memory");
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                        smem_addr(ring + (size_t)slot * row_bytes)),
                    "l"(reinterpret_cast<const S*>(a.X) + (int64_t)s * a.x_ld), "r"(row_bytes), "r"(bar)
                    : "memory");
            }
        };
        for (; issued < nb && issued < R; ++issued) issue(issued);
        for (int k = 0; k < nb; ++k) {
            const uint32_t slot = (head + (uint32_t)k) % R;
            bulk_wait(bars + 8u * slot, (phases >> slot) & 1u);
            phases ^= 1u << slot;
            const S* src = reinterpret_cast<const S*>(ring + (size_t)slot * row_bytes);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                if (cval[ch]) {
                    float v[VEC];
                    if constexpr (sizeof(S) == 4) {
                        const float4 t = *reinterpret_cast<const float4*>(src + cofs[ch]);
                        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                    } else {
                        Elem<S, VEC>::widen(v, *reinterpret_cast<const uint4*>(src + cofs[ch]));
                    }
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc[ch][i] += v[i];
                }
            }
            __syncwarp();      // every lane has read the slot before the copy engine may overwrite it
            if (issued < nb) {
                issue(issued);
                ++issued;
            }
        }
        head = (head + (uint32_t)nb) % R;
    }
    if (is_chunk) {
        float* pr = a.partial + (task - a.n_rows) * a.d;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
            if (cval[ch]) Vec<VEC>::store(pr + cofs[ch], acc[ch]);
    } else {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
            if (cval[ch]) epilogue_store<S, VEC>(a, row, cofs[ch], acc[ch]);
    }
}

template <typename S, int VEC, int NCH, int R>
static int launch_bulk(const AggArgs& a, cudaStream_t st) {
    constexpr int WARPS = 8;
    const size_t smem = (size_t)WARPS * R * (size_t)a.d * sizeof(S) + (size_t)WARPS * R * 8;
    CB_CUDA(cudaFuncSetAttribute(k_agg_bulk<S, VEC, NCH, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t tasks = a.n_rows + a.n_chunks;
    const int64_t blocks = ceil_div(tasks, (int64_t)WARPS);
    CB_REQUIRE(blocks < (int64_t)INT32_MAX, CB_E_UNSUPPORTED, "aggregation grid too large");
    k_agg_bulk<S, VEC, NCH, R><<<(unsigned)blocks, WARPS * 32, smem, st>>>(a);
    CB_LAUNCH_CHECK();
    if (a.n_chunks > 0) {
        k_combine<S, VEC, 32, NCH><<<(unsigned)ceil_div(a.n_chunks, (int64_t)WARPS), WARPS * 32, 0, st>>>(a);
        CB_LAUNCH_CHECK();
    }
    return CB_OK;
}

// the bulk-copy experiment covers: full-width rows of 128..256 (fp32) / 256..512 (bf16) elements, a contiguous 16-byte
// aligned source, the plain walk (no live flags, no compacted lists, no source-panel pass)
template <typename S, int VEC>
static bool try_bulk(const AggArgs& a, cudaStream_t st, int* rc) {
    static const int ring = getenv("CB_AGG_BULK") ? atoi(getenv("CB_AGG_BULK")) : 0;
    if (ring != 4 && ring != 8) return false;
    const int64_t units = ceil_div(a.d, VEC);
    if (a.live || a.ew || a.row_be || a.rowptr_exp || a.hub_rowptr || a.chunk_end || a.col0 != 0 || a.task0 != 0) return false;
    if (units <= 32 || units > 64 || a.d % VEC != 0 || (a.x_ld * (int64_t)sizeof(S)) % 16 != 0) return false;
    if ((size_t)8 * ring * a.d * sizeof(S) > 200 * 1024) return false;
    *rc = ring == 4 ? launch_bulk<S, VEC, 2, 4>(a, st) : launch_bulk<S, VEC, 2, 8>(a, st);
    return true;
}

template <typename S, int VEC>
static int launch_vec(AggArgs a, cudaStream_t st) {
    if constexpr (VEC > 1) {
        int rc = CB_OK;
        if (try_bulk<S, VEC>(a, st, &rc)) return rc;
    }
    const int64_t units = ceil_div(a.d, VEC);
    if (units <= 1) return launch_cfg<S, VEC, 1, 1>(a, st);
    if (units <= 2) return launch_cfg<S, VEC, 2, 1>(a, st);
    if (units <= 4) return launch_cfg<S, VEC, 4, 1>(a, st);
    if (units <= 8) return launch_cfg<S, VEC, 8, 1>(a, st);
    if (units <= 16) return launch_cfg<S, VEC, 16, 1>(a, st);
    if (units <= 32) return launch_cfg<S, VEC, 32, 1>(a, st);
    if (units <= 64) return launch_cfg<S, VEC, 32, 2>(a, st);
    // wider rows: column blocks of 128 vector slots, the neighbour list is walked once per block
    for (int64_t u0 = 0; u0 < units; u0 += 128) {
        a.col0 = u0 * VEC;
        const int64_t left = units - u0;
        int rc = left <= 32   ? launch_cfg<S, VEC, 32, 1>(a, st)
                 : left <= 64 ? launch_cfg<S, VEC, 32, 2>(a, st)
                              : launch_cfg<S, VEC, 32, 4>(a, st);
        if (rc) return rc;
    }
    return CB_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int run_agg(const cb_graph* g, int side_id, int dtype, AggArgs a, void* workspace, int64_t workspace_bytes,
                   cudaStream_t st, const void* live_ws = nullptr, int panel = -1, float* carry = nullptr) {
    const Side& s = side_id == CB_BY_DST ? g->by_dst : g->by_src;
    if (panel >= 0) {
        CB_REQUIRE(g->src_panels > 1 && s.rowptr_exp != nullptr, CB_E_INVALID,
                   "aggregation pass: the graph was not built with source panels (cb_graph_create_panelled)");
        CB_REQUIRE(panel < g->src_panels, CB_E_INVALID, "aggregation pass: panel out of range");
        CB_REQUIRE(carry != nullptr || g->rows == 0, CB_E_INVALID, "aggregation pass: carry buffer is NULL");
        CB_REQUIRE(live_ws == nullptr, CB_E_UNSUPPORTED, "aggregation pass over compacted lists is not supported");
        CB_REQUIRE(a.ew == nullptr, CB_E_UNSUPPORTED, "edge-weighted aggregation passes are not supported");
        a.rowptr_exp = s.rowptr_exp;
        a.n_panels = g->src_panels;
        a.panel = panel;
        a.carry = carry;
    }
    // measured at the bench shape: 21.5 -> 20.6 ms per fused forward aggregation (profiles/r02x_*)
    static const int prefetch_x0 = (getenv("CB_AGG_PREFETCH") && atoi(getenv("CB_AGG_PREFETCH")) == 0) ? 0 : 1;
    a.prefetch_x0 = prefetch_x0;
    a.rowptr = s.rowptr;
    a.col = s.col;
    a.n_rows = g->rows;
    a.chunk_row = s.chunk_row;
    a.chunk_beg = s.chunk_beg;
    a.n_chunks = s.n_chunks;
    CB_REQUIRE(!(a.ew && (live_ws || a.live)), CB_E_UNSUPPORTED, "edge weights on a row-sparse gather are not supported");
    if (live_ws) {
        const LiveView v = live_view(g, side_id, const_cast<void*>(live_ws));
        a.hub_rowptr = s.rowptr;
        a.rowptr = v.rowptr;
        a.col = v.col;
        a.chunk_beg = v.chunk_beg;
        a.chunk_end = v.chunk_end;
        // the lean kernel covers what cb_agg_gather_compacted can ask for: out = row_scale * sum
        if (!a.bias && !a.x0 && a.act == CB_ACT_NONE && !a.out2 && !a.mask && !a.clamp && a.out) a.row_be = v.row_be;
    }
    a.hub_chunk = g->hub_chunk;
    a.col0 = 0;
    a.partial = (float*)workspace;
    const int64_t need = s.n_chunks * a.d * (int64_t)sizeof(float);
    CB_REQUIRE(need == 0 || (workspace != nullptr && workspace_bytes >= need), CB_E_WORKSPACE,
               "aggregation workspace missing or smaller than cb_graph_workspace_bytes()");
    const bool al = aligned16(a.X) && aligned16(a.out) && aligned16(a.out2) && aligned16(a.x0) && aligned16(a.bias) &&
                    aligned16(a.partial);
    if (dtype == CB_BF16) {
        const bool vec8 = (a.d % 8 == 0) && (a.x_ld % 8 == 0) && (a.o_ld % 8 == 0) && al &&
                          (reinterpret_cast<uintptr_t>(a.mask) & 7u) == 0;
        return vec8 ? launch_vec<__nv_bfloat16, 8>(a, st) : launch_vec<__nv_bfloat16, 1>(a, st);
    }
    const bool vec4 = (a.d % 4 == 0) && (a.x_ld % 4 == 0) && (a.o_ld % 4 == 0) && al &&
                      (reinterpret_cast<uintptr_t>(a.mask) & 3u) == 0;
    return vec4 ? launch_vec<float, 4>(a, st) : launch_vec<float, 1>(a, st);
}

static int agg_forward_impl(const cb_graph* g, int dtype, const void* H, int64_t ld_h, int64_t d, const float* bias,
                            const void* x0, double alpha, int act, void* out, void* out_scaled, uint8_t* mask,
                            int64_t ld_out, void* workspace, int64_t workspace_bytes, void* stream, int panel = -1,
                            float* carry = nullptr) {
    CB_REQUIRE(g != nullptr, CB_E_INVALID, "cb_agg_forward: graph is NULL");
    CB_REQUIRE(d > 0, CB_E_INVALID, "cb_agg_forward: d must be positive");
    if (g->rows == 0) return CB_OK;       // a slice that owns no rows: nothing to aggregate (buffers may be NULL)
    CB_REQUIRE(H != nullptr, CB_E_INVALID, "cb_agg_forward: H is NULL");
    CB_REQUIRE(out != nullptr || out_scaled != nullptr, CB_E_INVALID, "cb_agg_forward: no output buffer");
    CB_REQUIRE(act == CB_ACT_NONE || act == CB_ACT_RELU, CB_E_INVALID, "cb_agg_forward: unknown activation");
    CB_REQUIRE((ld_h == 0 || ld_h >= d) && (ld_out == 0 || ld_out >= d), CB_E_INVALID,
               "cb_agg_forward: a row pitch is smaller than d");
    AggArgs a{};
    a.X = H;
    a.d = d;
    a.x_ld = ld_h ? ld_h : d;
    a.o_ld = ld_out ? ld_out : d;
    a.row_scale = g->din_is;
    a.bias = bias;
    a.x0 = x0;
    a.alpha = (float)alpha;
    a.one_minus_alpha = (float)(1.0 - alpha);
    a.act = act;
    a.out = out;
    a.out2_scale = g->dout_is;
    a.out2 = out_scaled;
    a.mask = mask;
    return run_agg(g, CB_BY_DST, dtype, a, workspace, workspace_bytes, (cudaStream_t)stream, nullptr, panel, carry);
}

static int agg_gather_impl(const cb_graph* g, int side, int dtype, const void* X, int64_t ld_x, int64_t d,
                           const float* row_scale, const uint8_t* row_live, void* out, int64_t ld_out, void* workspace,
                           int64_t workspace_bytes, void* stream, const void* live_ws = nullptr, int panel = -1,
                           float* carry = nullptr, const float* edge_val = nullptr) {
    CB_REQUIRE(g != nullptr, CB_E_INVALID, "cb_agg_gather: graph is NULL");
    CB_REQUIRE(side == CB_BY_DST || side == CB_BY_SRC, CB_E_INVALID, "cb_agg_gather: unknown side");
    CB_REQUIRE(d > 0, CB_E_INVALID, "cb_agg_gather: d must be positive");
    if (g->rows == 0) return CB_OK;
    CB_REQUIRE(X != nullptr && out != nullptr, CB_E_INVALID, "cb_agg_gather: NULL buffer");
    CB_REQUIRE((ld_x == 0 || ld_x >= d) && (ld_out == 0 || ld_out >= d), CB_E_INVALID,
               "cb_agg_gather: a row pitch is smaller than d");
    AggArgs a{};
    a.X = X;
    a.d = d;
    a.x_ld = ld_x ? ld_x : d;
    a.o_ld = ld_out ? ld_out : d;
    a.row_scale = row_scale;
    a.act = CB_ACT_NONE;
    a.out = out;
    a.live = row_live;
    a.ew = edge_val;
    return run_agg(g, side, dtype, a, workspace, workspace_bytes, (cudaStream_t)stream, live_ws, panel, carry);
}

// ---- edge-weighted aggregation (GCN.py:199-202): weights into stored order, per-edge dot products for dL/dw -------
__global__ void k_sort_edge_values(const float* __restrict__ values, const int32_t* __restrict__ perm, int64_t n,
                                   float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) out[j] = values[perm[j]];
}

template <typename S>
__device__ __forceinline__ float ld_elem(const S* p);
template <>
__device__ __forceinline__ float ld_elem<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld_elem<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// out[perm[j]] = <X[col[j], :], Y[row, :]> for every stored edge j of the side: one warp per task (a row, or one
// hub_chunk-long piece of a hub row), lanes stride over the columns, a fixed butterfly adds the 32 partials.
template <typename S>
__global__ void __launch_bounds__(256) k_edge_dot(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                  const int32_t* __restrict__ perm, const int32_t* __restrict__ chunk_row,
                                                  const int64_t* __restrict__ chunk_beg, int64_t n_rows, int64_t n_chunks,
                                                  int hub_chunk, const S* __restrict__ X, int64_t ld_x,
                                                  const S* __restrict__ Y, int64_t ld_y, int64_t d,
                                                  float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t task = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (task >= n_rows + n_chunks) return;
    int64_t row, beg, end;
    if (task < n_rows) {
        row = task;
        beg = rowptr[row];
        end = rowptr[row + 1];
        if (end - beg > hub_chunk) return;
    } else {
        const int64_t c = task - n_rows;
        row = chunk_row[c];
        beg = chunk_beg[c];
        const int64_t rend = rowptr[row + 1];
        end = beg + hub_chunk < rend ? beg + hub_chunk : rend;
    }
    const S* y = Y + row * ld_y;
    for (int64_t j = beg; j < end; ++j) {
        const S* x = X + (int64_t)col[j] * ld_x;
        float t = 0.f;
        for (int64_t k = lane; k < d; k += 32) t = __fadd_rn(t, __fmul_rn(ld_elem<S>(x + k), ld_elem<S>(y + k)));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) out[perm[j]] = t;
    }
}

}  // namespace cb

extern "C" {

int64_t cb_graph_workspace_bytes(const cb_graph_t* g, int side, int64_t d) {
    if (!g || d <= 0) return 0;
    const cb::Side& s = side == CB_BY_DST ? g->by_dst : g->by_src;
    return s.n_chunks * d * (int64_t)sizeof(float);
}

int cb_agg_forward(const cb_graph_t* g, const float* H, int64_t ld_h, int64_t d, const float* bias, const float* x0,
                   double alpha, int act, float* out, float* out_scaled, uint8_t* mask, int64_t ld_out,
                   void* workspace, int64_t workspace_bytes, void* stream) {
    return cb::agg_forward_impl(g, CB_F32, H, ld_h, d, bias, x0, alpha, act, out, out_scaled, mask, ld_out, workspace,
                                workspace_bytes, stream);
}

int cb_agg_forward_bf16(const cb_graph_t* g, const uint16_t* H, int64_t ld_h, int64_t d, const float* bias,
                        const uint16_t* x0, double alpha, int act, uint16_t* out, uint16_t* out_scaled, uint8_t* mask,
                        int64_t ld_out, void* workspace, int64_t workspace_bytes, void* stream) {
    return cb::agg_forward_impl(g, CB_BF16, H, ld_h, d, bias, x0, alpha, act, out, out_scaled, mask, ld_out, workspace,
                                workspace_bytes, stream);
}

int cb_agg_gather(const cb_graph_t* g, int side, const float* X, int64_t ld_x, int64_t d, const float* row_scale,
                  const uint8_t* row_live, float* out, int64_t ld_out, void* workspace, int64_t workspace_bytes,
                  void* stream) {
    return cb::agg_gather_impl(g, side, CB_F32, X, ld_x, d, row_scale, row_live, out, ld_out, workspace,
                               workspace_bytes, stream);
}

int cb_graph_sort_edge_values(const cb_graph_t* g, int side, const float* values, int64_t num_values, float* out,
                              void* stream) {
    CB_REQUIRE(g != nullptr, CB_E_INVALID, "cb_graph_sort_edge_values: graph is NULL");
    CB_REQUIRE(side == CB_BY_DST || side == CB_BY_SRC, CB_E_INVALID, "cb_graph_sort_edge_values: unknown side");
    const cb::Side& s = side == CB_BY_DST ? g->by_dst : g->by_src;
    if (s.n_edges == 0) return CB_OK;
    CB_REQUIRE(values != nullptr && out != nullptr, CB_E_INVALID, "cb_graph_sort_edge_values: NULL buffer");
    CB_REQUIRE(num_values >= s.n_edges, CB_E_INVALID,
               "cb_graph_sort_edge_values: fewer values than edges in the list the graph was built from");
    const int64_t blocks = cb::ceil_div(s.n_edges, 256);
    cb::k_sort_edge_values<<<(unsigned)(blocks < 65535 * 16 ? blocks : 65535 * 16), 256, 0, (cudaStream_t)stream>>>(
        values, s.perm, s.n_edges, out);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_agg_gather_weighted(const cb_graph_t* g, int side, int dtype, const void* X, int64_t ld_x, int64_t d,
                           const float* edge_val, const float* row_scale, void* out, int64_t ld_out, void* workspace,
                           int64_t workspace_bytes, void* stream) {
    CB_REQUIRE(dtype == CB_F32 || dtype == CB_BF16, CB_E_INVALID, "cb_agg_gather_weighted: unknown dtype");
    CB_REQUIRE(g == nullptr || edge_val != nullptr || (side == CB_BY_DST ? g->by_dst : g->by_src).n_edges == 0,
               CB_E_INVALID, "cb_agg_gather_weighted: edge_val is NULL");
    return cb::agg_gather_impl(g, side, dtype, X, ld_x, d, row_scale, nullptr, out, ld_out, workspace, workspace_bytes,
                               stream, nullptr, -1, nullptr, edge_val);
}

int cb_agg_edge_dot(const cb_graph_t* g, int side, int dtype, const void* X, int64_t ld_x, const void* Y, int64_t ld_y,
                    int64_t d, float* out, void* stream) {
    CB_REQUIRE(g != nullptr, CB_E_INVALID, "cb_agg_edge_dot: graph is NULL");
    CB_REQUIRE(side == CB_BY_DST || side == CB_BY_SRC, CB_E_INVALID, "cb_agg_edge_dot: unknown side");
    CB_REQUIRE(dtype == CB_F32 || dtype == CB_BF16, CB_E_INVALID, "cb_agg_edge_dot: unknown dtype");
    CB_REQUIRE(d > 0, CB_E_INVALID, "cb_agg_edge_dot: d must be positive");
    const cb::Side& s = side == CB_BY_DST ? g->by_dst : g->by_src;
    if (g->rows == 0 || s.n_edges == 0) return CB_OK;
    CB_REQUIRE(X && Y && out, CB_E_INVALID, "cb_agg_edge_dot: NULL buffer");
    ld_x = ld_x ? ld_x : d;
    ld_y = ld_y ? ld_y : d;
    const int64_t tasks = g->rows + s.n_chunks;
    const int64_t blocks = cb::ceil_div(tasks, 8);
    CB_REQUIRE(blocks < (int64_t)INT32_MAX, CB_E_UNSUPPORTED, "cb_agg_edge_dot: grid too large");
    if (dtype == CB_F32)
        cb::k_edge_dot<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            s.rowptr, s.col, s.perm, s.chunk_row, s.chunk_beg, g->rows, s.n_chunks, g->hub_chunk, (const float*)X, ld_x,
            (const float*)Y, ld_y, d, out);
    else
        cb::k_edge_dot<__nv_bfloat16><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            s.rowptr, s.col, s.perm, s.chunk_row, s.chunk_beg, g->rows, s.n_chunks, g->hub_chunk,
            (const __nv_bfloat16*)X, ld_x, (const __nv_bfloat16*)Y, ld_y, d, out);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_agg_gather_compacted(const cb_graph_t* g, int side, int dtype, const void* X, int64_t ld_x, int64_t d,
                            const float* row_scale, const void* live_ws, void* out, int64_t ld_out, void* workspace,
                            int64_t workspace_bytes, void* stream) {
    CB_REQUIRE(live_ws != nullptr, CB_E_INVALID, "cb_agg_gather_compacted: live_ws is NULL");
    CB_REQUIRE(dtype == CB_F32 || dtype == CB_BF16, CB_E_INVALID, "cb_agg_gather_compacted: unknown dtype");
    return cb::agg_gather_impl(g, side, dtype, X, ld_x, d, row_scale, nullptr, out, ld_out, workspace,
                               workspace_bytes, stream, live_ws);
}

int cb_agg_forward_pass(const cb_graph_t* g, int dtype, const void* H, int64_t ld_h, int64_t d, const float* bias,
                        const void* x0, double alpha, int act, void* out, void* out_scaled, uint8_t* mask, int64_t ld_out,
                        int panel, float* carry, void* workspace, int64_t workspace_bytes, void* stream) {
    CB_REQUIRE(dtype == CB_F32 || dtype == CB_BF16, CB_E_INVALID, "cb_agg_forward_pass: unknown dtype");
    CB_REQUIRE(panel >= 0, CB_E_INVALID, "cb_agg_forward_pass: negative panel");
    return cb::agg_forward_impl(g, dtype, H, ld_h, d, bias, x0, alpha, act, out, out_scaled, mask, ld_out, workspace,
                                workspace_bytes, stream, panel, carry);
}

int cb_agg_gather_pass(const cb_graph_t* g, int side, int dtype, const void* X, int64_t ld_x, int64_t d,
                       const float* row_scale, void* out, int64_t ld_out, int panel, float* carry, void* workspace,
                       int64_t workspace_bytes, void* stream) {
    CB_REQUIRE(dtype == CB_F32 || dtype == CB_BF16, CB_E_INVALID, "cb_agg_gather_pass: unknown dtype");
    CB_REQUIRE(panel >= 0, CB_E_INVALID, "cb_agg_gather_pass: negative panel");
    return cb::agg_gather_impl(g, side, dtype, X, ld_x, d, row_scale, nullptr, out, ld_out, workspace, workspace_bytes,
                               stream, nullptr, panel, carry);
}

int cb_agg_propagate(const cb_graph_t* g, int side, const float* X, int64_t d, const float* row_scale, const float* y,
                     double c_agg, double c_y, int clamp, double clamp_lo, double clamp_hi, float* out,
                     const float* out2_scale, float* out2, void* workspace, int64_t workspace_bytes, void* stream) {
    using namespace cb;
    CB_REQUIRE(g != nullptr, CB_E_INVALID, "cb_agg_propagate: graph is NULL");
    CB_REQUIRE(side == CB_BY_DST || side == CB_BY_SRC, CB_E_INVALID, "cb_agg_propagate: unknown side");
    CB_REQUIRE(d > 0, CB_E_INVALID, "cb_agg_propagate: d must be positive");
    if (g->rows == 0) return CB_OK;
    CB_REQUIRE(X != nullptr && y != nullptr && (out != nullptr || out2 != nullptr), CB_E_INVALID,
               "cb_agg_propagate: NULL buffer");
    CB_REQUIRE(!out2 || out2_scale, CB_E_INVALID, "cb_agg_propagate: out2 needs out2_scale");
    AggArgs a{};
    a.X = X;
    a.d = d;
    a.x_ld = d;
    a.o_ld = d;
    a.row_scale = row_scale;
    a.act = CB_ACT_NONE;
    a.x0 = y;
    a.one_minus_alpha = (float)c_agg;
    a.alpha = (float)c_y;
    a.clamp = clamp;
    a.clamp_lo = (float)clamp_lo;
    a.clamp_hi = (float)clamp_hi;
    a.out = out;
    a.out2_scale = out2_scale;
    a.out2 = out2;
    return run_agg(g, side, CB_F32, a, workspace, workspace_bytes, (cudaStream_t)stream);
}

int cb_agg_gather_bf16(const cb_graph_t* g, int side, const uint16_t* X, int64_t ld_x, int64_t d,
                       const float* row_scale, const uint8_t* row_live, uint16_t* out, int64_t ld_out, void* workspace,
                       int64_t workspace_bytes, void* stream) {
    return cb::agg_gather_impl(g, side, CB_BF16, X, ld_x, d, row_scale, row_live, out, ld_out, workspace,
                               workspace_bytes, stream);
}

}  // extern "C"
