// Dense transform of the GCN layer on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), fp32 in /
// fp32 out with "3xTF32" split operands so that the result keeps fp32-level accuracy:
//
//   cb_gemm_rows     C[M,N] = epilogue( A[M,K] . Bt[N,K]^T )
//       GNN_model/GCN.py:225   th.matmul(feat_src, weight)      (and its adjoint  dX = dH . W^T)
//       GCN.py:205-213         out-degree row scale             (epilogue: rs[row] * acc, since (D X) W = D (X W))
//       GCN.py:230-231         + self.le                        (epilogue: + add[row, col])
//       GCN.py:104-106,138     Linear + bias (+ relu)           (epilogue: + bias[col], relu)
//   cb_gemm_rows_grad  the adjoint dX = dH . W^T with the backward prologue of the layer below in its epilogue
//       autograd of GCN.py:242-253, 127-128, res_tricks.py:23 (relu gate, residual split, degree scale, bias
//       column sums) or of the input Linear's relu + bias (GCN.py:104-106); optional row-liveness flags
//   cb_peer_push_t     multi-GPU: finished rows are also stored into the peers that gather them (the exchange of
//                      the node-sliced path rides on the epilogue)
//   cb_gemm_split_weight       hi/lo TF32 split of the small weight operand (done once per call, 256 KB)
//
// Precision.  Every fp32 value x is split as x = hi + lo with hi = tf32_rna(x), lo = tf32_rna(x - hi)
// (x - hi is exact in fp32).  The kernel accumulates  A_lo.B_hi + A_hi.B_lo + A_hi.B_hi  in fp32 in
// tensor memory; the dropped term A_lo.B_lo and the rounding of lo are O(2^-21) relative per product,
// i.e. fp32-class accuracy (the reference GEMM is cuBLAS/MKL SGEMM with TF32 off, GCN.py:225).
//
// Structure (one CTA per SM, persistent over 128-row tiles, 320 threads):
//   warp 0      TMA producer: per 32-wide k-chunk one box of A (128 x 32 fp32, raw) and the matching
//               boxes of Bt_hi / Bt_lo (BN x 32) into a SWIZZLE_128B stage; mbarrier expect_tx
//   warps 2-5   split: read the raw A chunk from shared memory, write hi in place and lo beside it
//               (element-wise at identical offsets, so the TMA swizzle is preserved), fence.proxy.async
//   warp 1      MMA issuer: 3 x (BK/8) tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) per chunk into one of two
//               TMEM accumulators; tcgen05.commit frees the stage / publishes the accumulator
//   warps 6-9   epilogue: tcgen05.ld 32x32b.x32 -> registers -> per-warp padded smem transpose -> an 8 x float4
//               register tile per lane (8 rows x 4 columns of a 32-column slab) -> straight passes over that tile,
//               each behind a uniform branch -> coalesced 16-byte stores.  One warp per SM sub-partition issues
//               all of it: no per-element predicates, pointers stepped by constant strides.
// The weight operand is re-streamed from L2 per tile (it is 2 x N x K x 4 bytes = 512 KB at 256x256); measured, that
// is not the limiter (the kernel runs at ~78 % of the TF32 issue rate; see profiles/r01_summary_final.md).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "cb_internal.cuh"

namespace cb {
namespace tc {

constexpr int BM = 128;
constexpr int THREADS = 320;
constexpr int STG_LD = 36;  // floats per staging row (32 + 4 pad: 16-byte aligned, conflict-free)

// Storage type S of the streamed matrices (A, add, gate, d_x0, out, out2): float -> 3xTF32 split operands
// (kind::tf32), __nv_bfloat16 -> operands fed as stored (kind::f16, BASELINE.json configs[4]).  A k-chunk is
// always one 128-byte SWIZZLE_128B row; the accumulator and the whole epilogue are fp32 either way, bf16 results
// are rounded to nearest-even once, at the store.
template <typename S>
struct El;
template <>
struct El<float> {
    static constexpr int BKE = 32, UK = 8;       // elements per k-chunk, K of one MMA
    static constexpr bool SPLIT = true;
    static constexpr uint32_t FMT = 2u;          // TF32
    static constexpr CUtensorMapDataType TMA_T = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
};
template <>
struct El<__nv_bfloat16> {
    static constexpr int BKE = 64, UK = 16;
    static constexpr bool SPLIT = false;
    static constexpr uint32_t FMT = 1u;          // BF16
    static constexpr CUtensorMapDataType TMA_T = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
};

template <typename S, int BN>
struct Cfg {
    static constexpr bool SPLIT = El<S>::SPLIT;
    static constexpr int STAGES = SPLIT ? (BN >= 256 ? 2 : (BN >= 128 ? 3 : 4)) : 4;
    static constexpr int A_BYTES = BM * 128;      // 16 KB
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = SPLIT ? 2 * A_BYTES + 2 * B_BYTES : A_BYTES + B_BYTES;
    static constexpr int TX_BYTES = SPLIT ? A_BYTES + 2 * B_BYTES : A_BYTES + B_BYTES;
    static constexpr int STG_BYTES = 4 * 32 * STG_LD * 4;
    static constexpr int CS_BYTES = 4 * BN * 4;   // per-epilogue-warp column sums (gradient epilogue)
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES + STG_BYTES + CS_BYTES;
    static constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;  // barriers + alignment slack
    static constexpr int TMEM_COLS = 2 * BN;
};

struct GemmArgs {
    int64_t M;
    int N, K;
    const float* row_scale;  // [M] or null
    const float* bias;       // [N] or null
    const void* add;         // [M, ld_add] or null (storage type S)
    int64_t ld_add;
    int act;
    void* out;               // [M, ld_out] or null (S)
    int64_t ld_out;
    const float* out2_scale; // [M]
    void* out2;              // [M, ld_out2] or null (S)
    int64_t ld_out2;
    int64_t n_tiles_m;
    uint8_t* relu_mask;      // forward: [M, ld_mask] bytes, 1 where the relu output is positive, or null
    int64_t ld_mask;
    int tile_first, tile_step;   // this launch's 128-row tiles: tile_first, tile_first + tile_step, ... (0, 1 = all)
    // ---- gradient epilogue (k_gemm_rows<S, BN, true>): the backward prologue of the layer below ----
    const uint8_t* gate_u8;   // [M, ld_gate] relu mask bytes, or null
    const void* gate_f32;     // [M, ld_gate] relu output in S (gate = value > 0), or null
    int64_t ld_gate;
    int mixed;                // dz = (1-alpha) * dtot when the layer output was mixed with x0
    float alpha, one_minus_alpha;
    void* d_x0;               // [M, ld_dx0] (+)= alpha * dtot, or null (S)
    int64_t ld_dx0;
    int accumulate_x0;
    const float* post_scale;  // [M] scale of the stored value (din^-1/2), or null
    float* col_partial;       // [gridDim.x, N] per-CTA column sums of dz, or null
    uint8_t* row_live;        // [M] set to 1 where the stored row has a non-zero element, or null
    const uint8_t* a_live;    // [M] or null: rows with 0 have an all-zero A row (their dtot is 0): nothing of them is
                              // loaded or stored -- out / d_x0 keep whatever the buffers held (the consumers skip them)
    const uint8_t* x0_valid;  // [M] or null (with accumulate_x0): rows with 0 hold no valid d_x0 yet, read as 0
    // ---- multi-GPU: rows of `out` are also stored into the peers that gather them (cb_peer_push_t) ----
    cb_peer_push_t push;
};

// ---- PTX helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a pipeline bug traps instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
        if (spin > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// Pull [ptr, ptr + bytes) into L2 (bytes a multiple of 16): the epilogue's streamed operands are requested one
// tile ahead so that its loads hit L2 instead of waiting a DRAM round trip with few bytes in flight.
__device__ __forceinline__ void l2_prefetch_bulk(const void* ptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
}
// rows [r0, r0 + nrows) x cols [c0, c0 + ncols) of a row-major matrix with `ld` elements per row
__device__ __forceinline__ void l2_prefetch_rows(const void* base, int elem_bytes, int64_t ld, int64_t r0, int nrows,
                                                 int c0, int ncols, int lane) {
    const char* p = reinterpret_cast<const char*>(base) + (r0 * ld + c0) * elem_bytes;
    if (c0 == 0 && ncols == ld) {                       // the rows are one contiguous block: 16 KB pieces
        const int64_t total = (int64_t)nrows * ld * elem_bytes;
        for (int64_t o = (int64_t)lane * 16384; o < total; o += 32 * 16384)
            l2_prefetch_bulk(p + o, (uint32_t)(total - o < 16384 ? total - o : 16384));
    } else {
        for (int r = lane; r < nrows; r += 32) l2_prefetch_bulk(p + (int64_t)r * ld * elem_bytes, (uint32_t)(ncols * elem_bytes));
    }
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
template <typename S>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    if (El<S>::SPLIT) umma_tf32(d_tmem, adesc, bdesc, idesc, acc);
    else umma_f16(d_tmem, adesc, bdesc, idesc, acc);
}

// four consecutive elements of a streamed matrix <-> float4 (bf16: widened on load, rounded to nearest-even on store)
__device__ __forceinline__ float4 bf16x4_to_f32(uint2 u) {
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xffff0000u));
}
__device__ __forceinline__ uint2 f32_to_bf16x4(float4 v) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}
__device__ __forceinline__ float4 ld4_cs(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4_cs(const __nv_bfloat16* p) { return bf16x4_to_f32(__ldcs(reinterpret_cast<const uint2*>(p))); }
__device__ __forceinline__ float4 ld4_g(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4_g(const __nv_bfloat16* p) { return bf16x4_to_f32(__ldg(reinterpret_cast<const uint2*>(p))); }
__device__ __forceinline__ void st4_cs(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ void st4_cs(__nv_bfloat16* p, float4 v) { __stcs(reinterpret_cast<uint2*>(p), f32_to_bf16x4(v)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) { *reinterpret_cast<uint2*>(p) = f32_to_bf16x4(v); }

__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// x = hi + lo, both representable in TF32 (low 13 mantissa bits zero)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    const float rest = __fsub_rn(x, __uint_as_float(hi));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(rest));
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 128 bytes, 8-row atoms of 1024 bytes)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
    uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                  // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                  // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
    return d;
}

template <typename S, int BN>
__device__ __forceinline__ constexpr uint32_t instr_desc() {
    return (1u << 4)                 // D format  = F32
           | (El<S>::FMT << 7)       // A format  = TF32 / BF16
           | (El<S>::FMT << 10)      // B format
           | ((uint32_t)(BN >> 3) << 17)   // N >> 3
           | ((uint32_t)(BM >> 4) << 24);  // M >> 4   (A and B both K-major: bits 15/16 = 0)
}

// Stores the lane's 8 x float4 register tile (rows rbase + 4*itr, column col) into every peer whose bit is
// set in need[row]: the exchange step of the node-sliced path, riding on the epilogue.
template <typename S>
__device__ __forceinline__ void push_to_peers(const cb_peer_push_t& ps, int64_t rbase, int col, int nval,
                                              const float4 (&v)[8]) {
    uint32_t nd[8];
#pragma unroll
    for (int itr = 0; itr < 8; ++itr) {
        nd[itr] = itr < nval ? (uint32_t)__ldg(ps.need + rbase + itr * 4) : 0u;
        if (ps.row_live && itr < nval && __ldg(ps.row_live + rbase + itr * 4) == 0) nd[itr] = 0u;
    }
    for (int j = 0; j < ps.n_peers; ++j) {
        S* p = reinterpret_cast<S*>(ps.peer[j]) + (ps.row0 + rbase) * ps.ld + col;
#pragma unroll
        for (int itr = 0; itr < 8; ++itr)
            if ((nd[itr] >> j) & 1u) st4(p + (int64_t)itr * 4 * ps.ld, v[itr]);
    }
}

// ---- the kernel -----------------------------------------------------------------------------------
template <typename S, int BN, bool GRAD>
__global__ void __launch_bounds__(THREADS, 1)
k_gemm_rows(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
            const __grid_constant__ CUtensorMap map_blo, const GemmArgs g) {
    using C = Cfg<S, BN>;
    constexpr int STAGES = C::STAGES;
    constexpr bool SPLIT = C::SPLIT;
    constexpr int BKE = El<S>::BKE;     // elements per k-chunk (one 128-byte row)
    constexpr int UK = El<S>::UK;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const uint32_t bar_base = smem_base + C::BAR_OFF;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto ready_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (3 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (3 * STAGES + 2 + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + C::BAR_OFF + 8 * (3 * STAGES + 4));

    auto a_hi = [&](int s) { return smem_base + (uint32_t)s * C::STAGE_BYTES; };
    auto a_lo = [&](int s) { return smem_base + (uint32_t)s * C::STAGE_BYTES + C::A_BYTES; };
    auto b_hi = [&](int s) { return smem_base + (uint32_t)s * C::STAGE_BYTES + (SPLIT ? 2 : 1) * C::A_BYTES; };
    auto b_lo = [&](int s) { return smem_base + (uint32_t)s * C::STAGE_BYTES + 2 * C::A_BYTES + C::B_BYTES; };

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&map_a);
            tma_prefetch_desc(&map_bhi);
            if (SPLIT) tma_prefetch_desc(&map_blo);
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(full_bar(s), 1);
                mbar_init(ready_bar(s), 4);
                mbar_init(empty_bar(s), 1);
            }
            for (int a = 0; a < 2; ++a) {
                mbar_init(tfull_bar(a), 1);
                mbar_init(tempty_bar(a), 4);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n0 = blockIdx.y * BN;
    const int nk = (g.K + BKE - 1) / BKE;
    constexpr int ES = (int)sizeof(S);

    if (warp == 0) {
        // ===== TMA producer =====
        uint32_t it = 0;
        for (int64_t tile = g.tile_first + (int64_t)blockIdx.x * g.tile_step; tile < g.n_tiles_m; tile += (int64_t)gridDim.x * g.tile_step) {
            {   // epilogue operands of this tile -> L2 (the epilogue reads them about one tile later)
                const int64_t r0 = tile * BM;
                const int nr = (int)(g.M - r0 < BM ? g.M - r0 : BM);
                const int nc = g.N - n0 < BN ? g.N - n0 : BN;
                if (g.add) l2_prefetch_rows(g.add, ES, g.ld_add, r0, nr, n0, nc, lane);
                if (GRAD) {
                    if (g.d_x0 && g.accumulate_x0) l2_prefetch_rows(g.d_x0, ES, g.ld_dx0, r0, nr, n0, nc, lane);
                    if (g.gate_u8) l2_prefetch_rows(g.gate_u8, 1, g.ld_gate, r0, nr, n0, nc, lane);
                    if (g.gate_f32) l2_prefetch_rows(g.gate_f32, ES, g.ld_gate, r0, nr, n0, nc, lane);
                }
            }
            for (int kc = 0; kc < nk; ++kc, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1u);
                if (lane == 0) {
                    mbar_arrive_expect_tx(full_bar(s), C::TX_BYTES);
                    tma_load_2d(a_hi(s), &map_a, full_bar(s), kc * BKE, (int)(tile * BM));
                    tma_load_2d(b_hi(s), &map_bhi, full_bar(s), kc * BKE, n0);
                    if (SPLIT) tma_load_2d(b_lo(s), &map_blo, full_bar(s), kc * BKE, n0);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = instr_desc<S, BN>();
        uint32_t it = 0, tl = 0;
        for (int64_t tile = g.tile_first + (int64_t)blockIdx.x * g.tile_step; tile < g.n_tiles_m; tile += (int64_t)gridDim.x * g.tile_step, ++tl) {
            const int acc = tl & 1u;
            const uint32_t aph = (tl >> 1) & 1u;
            mbar_wait(tempty_bar(acc), aph ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
            for (int kc = 0; kc < nk; ++kc, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1u;
                mbar_wait(full_bar(s), ph);
                if (SPLIT) mbar_wait(ready_bar(s), ph);
                tc_fence_after();
                if (lane == 0) {
                    const uint64_t dah = smem_desc_sw128(a_hi(s));
                    const uint64_t dbh = smem_desc_sw128(b_hi(s));
                    if (SPLIT) {
                        const uint64_t dal = smem_desc_sw128(a_lo(s));
                        const uint64_t dbl = smem_desc_sw128(b_lo(s));
#pragma unroll
                        for (int k = 0; k < BKE / UK; ++k) {
                            const uint64_t adv = (uint64_t)((k * UK * ES) >> 4);  // 32 bytes per k-step
                            umma<S>(d_tmem, dal + adv, dbh + adv, idesc, (kc | k) != 0);
                            umma<S>(d_tmem, dah + adv, dbl + adv, idesc, 1u);
                            umma<S>(d_tmem, dah + adv, dbh + adv, idesc, 1u);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < BKE / UK; ++k) {
                            const uint64_t adv = (uint64_t)((k * UK * ES) >> 4);  // 32 bytes per k-step
                            umma<S>(d_tmem, dah + adv, dbh + adv, idesc, (kc | k) != 0);
                        }
                    }
                    umma_commit(empty_bar(s));
                    if (kc == nk - 1) umma_commit(tfull_bar(acc));
                }
                __syncwarp();
            }
        }
    } else if (warp < 6) {
        // ===== split warps: raw fp32 A chunk -> TF32 hi (in place) + lo  (idle for bf16 operands) =====
        const int t = threadIdx.x - 64;
        uint32_t it = 0;
        for (int64_t tile = g.tile_first + (int64_t)blockIdx.x * g.tile_step; SPLIT && tile < g.n_tiles_m; tile += (int64_t)gridDim.x * g.tile_step) {
            for (int kc = 0; kc < nk; ++kc, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1u;
                mbar_wait(full_bar(s), ph);
                uint8_t* hi_p = smem_gen + (size_t)s * C::STAGE_BYTES;
                uint8_t* lo_p = hi_p + C::A_BYTES;
#pragma unroll
                for (int i = 0; i < C::A_BYTES / 16 / 128; ++i) {
                    const int idx = t + 128 * i;
                    const float4 v = *reinterpret_cast<const float4*>(hi_p + 16 * idx);
                    uint4 h, l;
                    split_tf32(v.x, h.x, l.x);
                    split_tf32(v.y, h.y, l.y);
                    split_tf32(v.z, h.z, l.z);
                    split_tf32(v.w, h.w, l.w);
                    *reinterpret_cast<uint4*>(hi_p + 16 * idx) = h;
                    *reinterpret_cast<uint4*>(lo_p + 16 * idx) = l;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(ready_bar(s));
            }
        }
    } else {
        // ===== epilogue warps =====
        const int q = warp & 3;  // TMEM lane quarter this warp may read
        float* stg = reinterpret_cast<float*>(smem_gen + STAGES * C::STAGE_BYTES) + q * 32 * STG_LD;
        const int rsub = lane >> 3;  // 0..3
        const int c4 = lane & 7;     // 0..7
        // GRAD: running column sums of dz over all tiles of the CTA, one row of BN floats per epilogue warp
        float* csw = reinterpret_cast<float*>(smem_gen + STAGES * C::STAGE_BYTES + C::STG_BYTES) + q * BN;
        if (GRAD) {
            for (int c = lane; c < BN; c += 32) csw[c] = 0.f;
            __syncwarp();
        }
        uint32_t tl = 0;
        for (int64_t tile = g.tile_first + (int64_t)blockIdx.x * g.tile_step; tile < g.n_tiles_m; tile += (int64_t)gridDim.x * g.tile_step, ++tl) {
            const int acc = tl & 1u;
            const uint32_t aph = (tl >> 1) & 1u;
            mbar_wait(tfull_bar(acc), aph);
            tc_fence_after();
            const int64_t row0 = tile * BM + q * 32;
            const int n_slabs = min(BN, g.N - n0 + 31) / 32;
            // Per-row epilogue operands do not depend on the slab: loaded once per tile (they used to be re-loaded in
            // every slab, each time a dependent DRAM / L2 round trip in front of the slab's arithmetic).
            const int64_t rbase = row0 + rsub;
            const int64_t left = (g.M - rbase + 3) >> 2;
            const int nrow = left <= 0 ? 0 : (left < 8 ? (int)left : 8);
            float rsv[8], psv[8];         // GRAD: row_scale, post_scale; forward: row_scale, out2_scale
            uint32_t lm_rows = nrow >= 8 ? 0xffu : ((1u << nrow) - 1u), vm_rows = 0xffu;
            if (GRAD) {
                if (g.a_live) {
#pragma unroll
                    for (int itr = 0; itr < 8; ++itr)
                        if (((lm_rows >> itr) & 1u) && __ldg(g.a_live + rbase + itr * 4) == 0) lm_rows &= ~(1u << itr);
                }
                if (g.x0_valid) {
#pragma unroll
                    for (int itr = 0; itr < 8; ++itr)
                        if (itr < nrow && __ldg(g.x0_valid + rbase + itr * 4) == 0) vm_rows &= ~(1u << itr);
                }
                if (g.post_scale) {
#pragma unroll
                    for (int itr = 0; itr < 8; ++itr)
                        if (itr < nrow) psv[itr] = __ldg(g.post_scale + rbase + itr * 4);
                }
            } else if (g.out2) {
#pragma unroll
                for (int itr = 0; itr < 8; ++itr)
                    if (itr < nrow) psv[itr] = __ldg(g.out2_scale + rbase + itr * 4);
            }
            if (g.row_scale) {
#pragma unroll
                for (int itr = 0; itr < 8; ++itr)
                    if (itr < nrow) rsv[itr] = __ldg(g.row_scale + rbase + itr * 4);
            }
            for (int j = 0; j < n_slabs; ++j) {
                // the slab's streamed gate words are requested BEFORE the accumulator is read, so that their latency
                // overlaps the TMEM load and the transpose
                const int col_pre = n0 + j * 32 + c4 * 4;
                const uint32_t lm_pre = col_pre < g.N ? lm_rows : 0u;
                uint32_t gm[8];
                if (GRAD && g.gate_u8) {
                    const uint8_t* p = g.gate_u8 + rbase * g.ld_gate + col_pre;
#pragma unroll
                    for (int itr = 0; itr < 8; ++itr)
                        if ((lm_pre >> itr) & 1u) gm[itr] = __ldg(reinterpret_cast<const uint32_t*>(p + (int64_t)itr * 4 * g.ld_gate));
                }
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + j * 32), r);
                tmem_ld_wait();
                if (j == n_slabs - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(acc));
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    *reinterpret_cast<uint4*>(stg + lane * STG_LD + 4 * i) =
                        make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
                __syncwarp();
                const int col = n0 + j * 32 + c4 * 4;
                const bool col_ok = col < g.N;
                if (GRAD) {
                    // dtot = rs*acc + add ; d_x0 (+)= alpha*dtot ; dz = [(1-alpha)*] dtot * gate ; out = ps*dz
                    // (same operations, in the same order, as cb_gemm_rows followed by cb_agg_backward_prep).
                    // Every option is a uniform branch around a straight pass over the lane's 8 x float4
                    // register tile, so no per-element predicates are executed; rows are valid for itr < nval.
                    const int nval = col_ok ? nrow : 0;
                    float4 av[8], gy[8];      // add or old d_x0 (mutually exclusive); fp32 gate source
                    const bool acc_x0 = g.d_x0 && g.accumulate_x0;
                    // rows this lane really works on: inside M, and -- with a_live -- not known to be all-zero
                    const uint32_t lm = col_ok ? lm_rows : 0u;
                    if (g.add || acc_x0) {
                        const int64_t ld = g.add ? g.ld_add : g.ld_dx0;
                        const S* p = reinterpret_cast<const S*>(g.add ? g.add : g.d_x0) + rbase * ld + col;
                        const uint32_t vm = g.add ? lm : (lm & vm_rows);           // rows whose old d_x0 exists
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr)
                            av[itr] = ((vm >> itr) & 1u) ? ld4_cs(p + (int64_t)itr * 4 * ld) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    if (!g.gate_u8 && g.gate_f32) {
                        const S* p = reinterpret_cast<const S*>(g.gate_f32) + rbase * g.ld_gate + col;
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr)
                            if ((lm >> itr) & 1u) gy[itr] = ld4_g(p + (int64_t)itr * 4 * g.ld_gate);
                    }
                    float4 v[8];
#pragma unroll
                    for (int itr = 0; itr < 8; ++itr)
                        v[itr] = *reinterpret_cast<const float4*>(stg + (itr * 4 + rsub) * STG_LD + 4 * c4);
                    if (g.row_scale) {
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr) {
                            const float r_ = rsv[itr];
                            v[itr].x = __fmul_rn(r_, v[itr].x); v[itr].y = __fmul_rn(r_, v[itr].y);
                            v[itr].z = __fmul_rn(r_, v[itr].z); v[itr].w = __fmul_rn(r_, v[itr].w);
                        }
                    }
                    if (g.add) {
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr) {
                            v[itr].x = __fadd_rn(v[itr].x, av[itr].x); v[itr].y = __fadd_rn(v[itr].y, av[itr].y);
                            v[itr].z = __fadd_rn(v[itr].z, av[itr].z); v[itr].w = __fadd_rn(v[itr].w, av[itr].w);
                        }
                    }
                    if (g.d_x0) {
                        S* p = reinterpret_cast<S*>(g.d_x0) + rbase * g.ld_dx0 + col;
                        const float al = g.alpha;
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr) {
                            float4 x = make_float4(__fmul_rn(al, v[itr].x), __fmul_rn(al, v[itr].y),
                                                   __fmul_rn(al, v[itr].z), __fmul_rn(al, v[itr].w));
                            if (acc_x0) {
                                x.x = __fadd_rn(av[itr].x, x.x); x.y = __fadd_rn(av[itr].y, x.y);
                                x.z = __fadd_rn(av[itr].z, x.z); x.w = __fadd_rn(av[itr].w, x.w);
                            }
                            if ((lm >> itr) & 1u) st4_cs(p + (int64_t)itr * 4 * g.ld_dx0, x);
                        }
                    }
                    if (g.mixed) {
                        const float om = g.one_minus_alpha;
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr) {
                            v[itr].x = __fmul_rn(om, v[itr].x); v[itr].y = __fmul_rn(om, v[itr].y);
                            v[itr].z = __fmul_rn(om, v[itr].z); v[itr].w = __fmul_rn(om, v[itr].w);
                        }
                    }
                    if (g.gate_u8) {
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr) {
                            const uint32_t m = gm[itr];
                            v[itr].x = (m & 0xffu) ? v[itr].x : 0.f;       v[itr].y = (m & 0xff00u) ? v[itr].y : 0.f;
                            v[itr].z = (m & 0xff0000u) ? v[itr].z : 0.f;   v[itr].w = (m & 0xff000000u) ? v[itr].w : 0.f;
                        }
                    } else if (g.gate_f32) {
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr) {
                            v[itr].x = gy[itr].x > 0.f ? v[itr].x : 0.f; v[itr].y = gy[itr].y > 0.f ? v[itr].y : 0.f;
                            v[itr].z = gy[itr].z > 0.f ? v[itr].z : 0.f; v[itr].w = gy[itr].w > 0.f ? v[itr].w : 0.f;
                        }
                    }
                    if (g.col_partial) {
                        if (lm != 0xffu) {      // rows outside M or skipped rows (whose gate words were not loaded)
#pragma unroll
                            for (int itr = 0; itr < 8; ++itr)
                                if (!((lm >> itr) & 1u)) v[itr] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        float4 cs = v[0];
#pragma unroll
                        for (int itr = 1; itr < 8; ++itr) {
                            cs.x += v[itr].x; cs.y += v[itr].y; cs.z += v[itr].z; cs.w += v[itr].w;
                        }
                        // lanes with equal c4 (rsub = 0..3) are added in rsub order; lane rsub == 0 owns the
                        // warp's running sum of those columns
                        const float c_[4] = {cs.x, cs.y, cs.z, cs.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float t1 = __shfl_sync(0xffffffffu, c_[i], c4 + 8);
                            const float t2 = __shfl_sync(0xffffffffu, c_[i], c4 + 16);
                            const float t3 = __shfl_sync(0xffffffffu, c_[i], c4 + 24);
                            if (rsub == 0) csw[j * 32 + c4 * 4 + i] += ((c_[i] + t1) + t2) + t3;
                        }
                    }
                    if (g.post_scale) {
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr) {
                            const float p_ = psv[itr];
                            v[itr].x = __fmul_rn(p_, v[itr].x); v[itr].y = __fmul_rn(p_, v[itr].y);
                            v[itr].z = __fmul_rn(p_, v[itr].z); v[itr].w = __fmul_rn(p_, v[itr].w);
                        }
                    }
                    {
                        S* p = reinterpret_cast<S*>(g.out) + rbase * g.ld_out + col;
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr)
                            if ((lm >> itr) & 1u) st4_cs(p + (int64_t)itr * 4 * g.ld_out, v[itr]);
                    }
                    if (g.push.n_peers) push_to_peers<S>(g.push, rbase, col, nval, v);
                    if (g.row_live) {
                        // a row's 32 columns of this slab sit in the 8 lanes with equal rsub: one ballot per row group
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr) {
                            const uint32_t bits = (__float_as_uint(v[itr].x) | __float_as_uint(v[itr].y) |
                                                   __float_as_uint(v[itr].z) | __float_as_uint(v[itr].w)) << 1;  // -0 is 0
                            const uint32_t b = __ballot_sync(0xffffffffu, itr < nval && bits != 0u);
                            if (c4 == 0 && ((b >> (rsub * 8)) & 0xffu)) g.row_live[rbase + itr * 4] = 1;
                        }
                    }
                    __syncwarp();
                    continue;
                }
                {
                    // forward epilogue, same structure: v = act(rs*acc + bias + add); out = v; out2 = s2*v
                    const int nval = col_ok ? nrow : 0;
                    float4 av[8];
                    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (g.bias && col_ok) bv = __ldg(reinterpret_cast<const float4*>(g.bias + col));
                    if (g.add) {
                        const S* p = reinterpret_cast<const S*>(g.add) + rbase * g.ld_add + col;
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr)
                            if (itr < nval) av[itr] = ld4_cs(p + (int64_t)itr * 4 * g.ld_add);
                    }
                    float4 v[8];
#pragma unroll
                    for (int itr = 0; itr < 8; ++itr)
                        v[itr] = *reinterpret_cast<const float4*>(stg + (itr * 4 + rsub) * STG_LD + 4 * c4);
                    if (g.row_scale) {
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr) {
                            const float r_ = rsv[itr];
                            v[itr].x = __fmul_rn(v[itr].x, r_); v[itr].y = __fmul_rn(v[itr].y, r_);
                            v[itr].z = __fmul_rn(v[itr].z, r_); v[itr].w = __fmul_rn(v[itr].w, r_);
                        }
                    }
                    if (g.bias) {
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr) {
                            v[itr].x = __fadd_rn(v[itr].x, bv.x); v[itr].y = __fadd_rn(v[itr].y, bv.y);
                            v[itr].z = __fadd_rn(v[itr].z, bv.z); v[itr].w = __fadd_rn(v[itr].w, bv.w);
                        }
                    }
                    if (g.add) {
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr) {
                            v[itr].x = __fadd_rn(v[itr].x, av[itr].x); v[itr].y = __fadd_rn(v[itr].y, av[itr].y);
                            v[itr].z = __fadd_rn(v[itr].z, av[itr].z); v[itr].w = __fadd_rn(v[itr].w, av[itr].w);
                        }
                    }
                    if (g.act == CB_ACT_RELU) {
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr) {
                            v[itr].x = fmaxf(v[itr].x, 0.f); v[itr].y = fmaxf(v[itr].y, 0.f);
                            v[itr].z = fmaxf(v[itr].z, 0.f); v[itr].w = fmaxf(v[itr].w, 0.f);
                        }
                    }
                    if (g.out) {
                        S* p = reinterpret_cast<S*>(g.out) + rbase * g.ld_out + col;
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr)
                            if (itr < nval) st4_cs(p + (int64_t)itr * 4 * g.ld_out, v[itr]);
                        if (g.push.n_peers) push_to_peers<S>(g.push, rbase, col, nval, v);
                    }
                    if (g.relu_mask) {
                        // the backward's relu gate as bytes: a quarter (fp32) of what re-reading the activations costs
                        uint8_t* p = g.relu_mask + rbase * g.ld_mask + col;
                        // "positive as STORED": a bf16 store rounds fp32 values up to 2^-134 to zero
                        const float z = sizeof(S) == 2 ? __uint_as_float(0x00008000u) : 0.f;
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr) {
                            const uint32_t m = (v[itr].x > z ? 1u : 0u) | (v[itr].y > z ? 0x100u : 0u) |
                                               (v[itr].z > z ? 0x10000u : 0u) | (v[itr].w > z ? 0x1000000u : 0u);
                            if (itr < nval) *reinterpret_cast<uint32_t*>(p + (int64_t)itr * 4 * g.ld_mask) = m;
                        }
                    }
                    if (g.out2) {
                        S* p = reinterpret_cast<S*>(g.out2) + rbase * g.ld_out2 + col;
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr) {
                            const float s_ = psv[itr];
                            const float4 w = make_float4(__fmul_rn(v[itr].x, s_), __fmul_rn(v[itr].y, s_),
                                                         __fmul_rn(v[itr].z, s_), __fmul_rn(v[itr].w, s_));
                            if (itr < nval) st4_cs(p + (int64_t)itr * 4 * g.ld_out2, w);
                        }
                    }
                }
                __syncwarp();
            }
        }
        if (GRAD && g.col_partial) {
            // the four epilogue warps are added in quarter order: a fixed association
            asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps only
            const float* cs0 = reinterpret_cast<const float*>(smem_gen + STAGES * C::STAGE_BYTES + C::STG_BYTES);
            const int et = threadIdx.x - 6 * 32;             // 0..127
            for (int c = et; c < BN; c += 128) {
                float t = 0.f;
#pragma unroll
                for (int w = 0; w < 4; ++w) t += cs0[w * BN + c];
                if (n0 + c < g.N) g.col_partial[(int64_t)blockIdx.x * g.N + n0 + c] = t;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
    }
}

// col_sum[c] = sum over the CTA partials, in CTA order
__global__ void __launch_bounds__(256) k_col_final(const float* __restrict__ partial, int n_ctas, int n,
                                                   float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    float t = 0.f;
    for (int b = 0; b < n_ctas; ++b) t += partial[(int64_t)b * n + c];
    out[c] = t;
}

// hi/lo TF32 split of the weight operand, optionally transposed:  dst[n, k] = src[n, k] or src[k, n]
__global__ void __launch_bounds__(256) k_split_weight(const float* __restrict__ W, int n_rows, int k_cols,
                                                      int transpose, float* __restrict__ hi,
                                                      float* __restrict__ lo) {
    const int64_t total = (int64_t)n_rows * k_cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / k_cols), k = (int)(i % k_cols);
        const float x = transpose ? W[(int64_t)k * n_rows + n] : W[i];
        uint32_t h, l;
        split_tf32(x, h, l);
        hi[i] = __uint_as_float(h);
        lo[i] = __uint_as_float(l);
    }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows, cols] row-major matrix of S with a row pitch of ld elements; box = box_rows x one 128-byte row, SWIZZLE_128B
template <typename S>
static int make_map(CUtensorMap* map, const S* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    CB_REQUIRE(fn != nullptr, CB_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(S)};
    const cuuint32_t box[2] = {(cuuint32_t)El<S>::BKE, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, El<S>::TMA_T, 2, const_cast<S*>(base), dims, strides, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return CB_E_CUDA;
    }
    return CB_OK;
}

static int64_t gemm_grid_x(int64_t n_tiles_m, int max_ctas = 0) {
    int64_t g = n_tiles_m < sm_count() ? n_tiles_m : sm_count();
    return (max_ctas > 0 && max_ctas < g) ? max_ctas : g;
}

template <typename S, int BN, bool GRAD>
static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mh, const CUtensorMap& ml, const GemmArgs& g,
                       cudaStream_t st) {
    using C = Cfg<S, BN>;
    // the attribute is per device (and per context): one flag per device ordinal, set under a benign race
    static std::atomic<bool> configured[64];
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
        CB_CUDA(cudaFuncSetAttribute(k_gemm_rows<S, BN, GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     C::SMEM_BYTES));
        if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
    }
    const int64_t my_tiles = g.n_tiles_m > g.tile_first ? ceil_div(g.n_tiles_m - g.tile_first, g.tile_step) : 0;
    if (my_tiles == 0) return CB_OK;      // a panel without tiles on this rank
    dim3 grid((unsigned)gemm_grid_x(my_tiles, g.push.max_ctas), (unsigned)ceil_div(g.N, BN));
    k_gemm_rows<S, BN, GRAD><<<grid, THREADS, C::SMEM_BYTES, st>>>(ma, mh, ml, g);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

// fp32 weight -> bf16, optionally transposed:  dst[n, k] = src[n, k] or src[k, n]   (round to nearest even)
__global__ void __launch_bounds__(256) k_weight_to_bf16(const float* __restrict__ W, int n_rows, int k_cols,
                                                        int transpose, __nv_bfloat16* __restrict__ out) {
    const int64_t total = (int64_t)n_rows * k_cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / k_cols), k = (int)(i % k_cols);
        out[i] = __float2bfloat16_rn(transpose ? W[(int64_t)k * n_rows + n] : W[i]);
    }
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// elements per 16 bytes: pitches and widths must be multiples of it (TMA strides, vector accesses)
template <typename S>
constexpr int vec16() { return 16 / (int)sizeof(S); }

template <typename S>
static int rows_supported(int64_t M, int64_t N, int64_t K) {
    return M > 0 && N > 0 && K > 0 && N % 4 == 0 && K % vec16<S>() == 0 && M < ((int64_t)1 << 31) - 128 &&
           N < (1 << 20) && K < (1 << 20);
}

template <typename S>
static int gemm_rows_impl(const S* A, int64_t M, int64_t K, int64_t lda, const S* Bt_hi, const S* Bt_lo, int64_t N,
                          const float* row_scale, const float* bias, const S* add, int64_t ld_add, int act, S* out,
                          int64_t ld_out, const float* out2_scale, S* out2, int64_t ld_out2, const cb_peer_push_t* push,
                          void* stream, uint8_t* relu_mask = nullptr, int64_t ld_mask = 0) {
    CB_REQUIRE(!relu_mask || (ld_mask >= N && ld_mask % 4 == 0 && (reinterpret_cast<uintptr_t>(relu_mask) & 3u) == 0),
               CB_E_INVALID, "cb_gemm_rows_masked: the mask needs a 4-byte aligned base and pitch >= N");
    CB_REQUIRE(!push || (push->n_peers >= 0 && push->n_peers <= CB_MAX_PEERS && (push->n_peers == 0 ||
                         (out && push->need && push->ld >= N && push->ld % 4 == 0))), CB_E_INVALID,
               "cb_gemm_rows: bad cb_peer_push_t");
    CB_REQUIRE(A && Bt_hi && (Bt_lo || !El<S>::SPLIT), CB_E_INVALID, "cb_gemm_rows: NULL operand");
    CB_REQUIRE(out || out2, CB_E_INVALID, "cb_gemm_rows: no output buffer");
    CB_REQUIRE(!out2 || out2_scale, CB_E_INVALID, "cb_gemm_rows: out2 needs out2_scale");
    CB_REQUIRE(act == CB_ACT_NONE || act == CB_ACT_RELU, CB_E_INVALID, "cb_gemm_rows: unknown activation");
    CB_REQUIRE(rows_supported<S>(M, N, K), CB_E_UNSUPPORTED,
               "cb_gemm_rows: needs N % 4 == 0, K a multiple of 16 bytes and M < 2^31");
    CB_REQUIRE(lda >= K && lda % vec16<S>() == 0 && al16(A) && al16(Bt_hi) && al16(Bt_lo), CB_E_UNSUPPORTED,
               "cb_gemm_rows: operands must be 16-byte aligned with a row pitch that is a multiple of 16 bytes");
    // epilogue accesses are 4 elements wide: 16 bytes (fp32) or 8 bytes (bf16)
    auto al4 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & (4 * sizeof(S) - 1)) == 0; };
    CB_REQUIRE((!out || (al4(out) && ld_out % 4 == 0 && ld_out >= N)) &&
                   (!out2 || (al4(out2) && ld_out2 % 4 == 0 && ld_out2 >= N)) &&
                   (!add || (al4(add) && ld_add % 4 == 0 && ld_add >= N)) && (!bias || al16(bias)),
               CB_E_UNSUPPORTED, "cb_gemm_rows: epilogue buffers must be aligned to 4 elements, pitches multiples of 4");
    const int bn = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
    CUtensorMap ma, mh, ml;
    int rc = make_map<S>(&ma, A, M, K, lda, BM);
    if (rc) return rc;
    rc = make_map<S>(&mh, Bt_hi, N, K, K, bn);
    if (rc) return rc;
    ml = mh;
    if (El<S>::SPLIT) {
        rc = make_map<S>(&ml, Bt_lo, N, K, K, bn);
        if (rc) return rc;
    }
    GemmArgs g{};
    g.M = M; g.N = (int)N; g.K = (int)K;
    g.row_scale = row_scale; g.bias = bias; g.add = add; g.ld_add = ld_add; g.act = act;
    g.out = out; g.ld_out = ld_out; g.out2_scale = out2_scale; g.out2 = out2; g.ld_out2 = ld_out2;
    g.relu_mask = relu_mask; g.ld_mask = ld_mask;
    if (push) g.push = *push;
    g.n_tiles_m = ceil_div(M, BM);
    g.tile_first = (push && push->tile_step > 0) ? push->tile_first : 0;
    g.tile_step = (push && push->tile_step > 0) ? push->tile_step : 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (bn == 64) return launch_gemm<S, 64, false>(ma, mh, ml, g, st);
    if (bn == 128) return launch_gemm<S, 128, false>(ma, mh, ml, g, st);
    return launch_gemm<S, 256, false>(ma, mh, ml, g, st);
}

template <typename S>
static int gemm_rows_grad_impl(const S* A, int64_t M, int64_t K, int64_t lda, const S* Bt_hi, const S* Bt_lo, int64_t N,
                               const float* row_scale, const S* add, int64_t ld_add, const uint8_t* gate_u8,
                               const S* gate_val, int64_t ld_gate, int mixed, double alpha, S* d_x0, int64_t ld_dx0,
                               int accumulate_x0, const float* post_scale, S* out, int64_t ld_out, float* col_sum,
                               uint8_t* row_live, const uint8_t* a_live, const uint8_t* x0_valid, void* workspace,
                               int64_t workspace_bytes, const cb_peer_push_t* push, void* stream) {
    CB_REQUIRE(!push || (push->n_peers >= 0 && push->n_peers <= CB_MAX_PEERS && (push->n_peers == 0 ||
                         (push->need && push->ld >= N && push->ld % 4 == 0))), CB_E_INVALID,
               "cb_gemm_rows_grad: bad cb_peer_push_t");
    CB_REQUIRE(A && Bt_hi && (Bt_lo || !El<S>::SPLIT) && out, CB_E_INVALID, "cb_gemm_rows_grad: NULL operand");
    CB_REQUIRE(!(gate_u8 && gate_val), CB_E_INVALID, "cb_gemm_rows_grad: one gate at most");
    CB_REQUIRE(!(add && d_x0 && accumulate_x0), CB_E_UNSUPPORTED,
               "cb_gemm_rows_grad: `add` and an accumulating d_x0 cannot be combined");
    CB_REQUIRE(!(a_live && add), CB_E_INVALID,
               "cb_gemm_rows_grad: a_live promises all-zero output rows, which `add` would break");
    CB_REQUIRE(rows_supported<S>(M, N, K), CB_E_UNSUPPORTED,
               "cb_gemm_rows_grad: needs N % 4 == 0, K a multiple of 16 bytes and M < 2^31");
    CB_REQUIRE(lda >= K && lda % vec16<S>() == 0 && al16(A) && al16(Bt_hi) && al16(Bt_lo), CB_E_UNSUPPORTED,
               "cb_gemm_rows_grad: operands must be 16-byte aligned with a row pitch that is a multiple of 16 bytes");
    auto al4 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & (4 * sizeof(S) - 1)) == 0; };
    CB_REQUIRE(al4(out) && ld_out % 4 == 0 && ld_out >= N && (!add || (al4(add) && ld_add % 4 == 0 && ld_add >= N)) &&
                   (!d_x0 || (al4(d_x0) && ld_dx0 % 4 == 0 && ld_dx0 >= N)) &&
                   (!gate_val || (al4(gate_val) && ld_gate % 4 == 0 && ld_gate >= N)) &&
                   (!gate_u8 || ((reinterpret_cast<uintptr_t>(gate_u8) & 3u) == 0 && ld_gate % 4 == 0 && ld_gate >= N)),
               CB_E_UNSUPPORTED, "cb_gemm_rows_grad: epilogue buffers must be aligned, pitches multiples of 4");
    CB_REQUIRE(!col_sum || (workspace && workspace_bytes >= cb_gemm_rows_grad_workspace_bytes(M, N)), CB_E_WORKSPACE,
               "cb_gemm_rows_grad: workspace smaller than cb_gemm_rows_grad_workspace_bytes()");
    const int bn = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
    CUtensorMap ma, mh, ml;
    int rc = make_map<S>(&ma, A, M, K, lda, BM);
    if (rc) return rc;
    rc = make_map<S>(&mh, Bt_hi, N, K, K, bn);
    if (rc) return rc;
    ml = mh;
    if (El<S>::SPLIT) {
        rc = make_map<S>(&ml, Bt_lo, N, K, K, bn);
        if (rc) return rc;
    }
    GemmArgs g{};
    g.M = M; g.N = (int)N; g.K = (int)K;
    g.row_scale = row_scale; g.add = add; g.ld_add = ld_add;
    g.out = out; g.ld_out = ld_out;
    g.n_tiles_m = ceil_div(M, BM);
    g.gate_u8 = gate_u8; g.gate_f32 = gate_val; g.ld_gate = ld_gate;
    g.mixed = mixed; g.alpha = (float)alpha; g.one_minus_alpha = (float)(1.0 - alpha);
    g.d_x0 = d_x0; g.ld_dx0 = ld_dx0; g.accumulate_x0 = d_x0 ? accumulate_x0 : 0;
    g.post_scale = post_scale;
    g.col_partial = col_sum ? (float*)workspace : nullptr;
    g.row_live = row_live;
    g.a_live = a_live;
    g.x0_valid = (d_x0 && accumulate_x0) ? x0_valid : nullptr;
    if (push) g.push = *push;
    g.tile_first = (push && push->tile_step > 0) ? push->tile_first : 0;
    g.tile_step = (push && push->tile_step > 0) ? push->tile_step : 1;
    cudaStream_t st = (cudaStream_t)stream;
    rc = bn == 64 ? launch_gemm<S, 64, true>(ma, mh, ml, g, st)
                  : (bn == 128 ? launch_gemm<S, 128, true>(ma, mh, ml, g, st)
                               : launch_gemm<S, 256, true>(ma, mh, ml, g, st));
    if (rc) return rc;
    if (col_sum) {
        const int64_t my_tiles = g.n_tiles_m > g.tile_first ? ceil_div(g.n_tiles_m - g.tile_first, g.tile_step) : 0;
        if (my_tiles == 0) {
            CB_CUDA(cudaMemsetAsync(col_sum, 0, (size_t)N * sizeof(float), st));
            return CB_OK;
        }
        k_col_final<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(
            (const float*)workspace, (int)gemm_grid_x(my_tiles, g.push.max_ctas), (int)N, col_sum);
        CB_LAUNCH_CHECK();
    }
    return CB_OK;
}

// =====================================================================================================
// Weight gradient:  C[Ka, Nb] = A[M, Ka]^T . B[M, Nb]   (reduction over the M node rows)
//   autograd of GCN.py:225 (dW = (D X)^T dH) and of the Linear layers (dW = dY^T X).
// Both operands are read as stored (row-major, feature-contiguous = "MN-major" for the tensor core):
// a k-chunk is 16 node rows; TMA lands it as [feature group of 32][16 nodes][128 bytes]
// (SWIZZLE_128B_ATOM_32B = UMMA SWIZZLE_128B_BASE32B, the one layout MN-major TF32 operands may use),
// the canonical MN-major UMMA layout with leading byte offset = 16*128 between feature groups and
// stride byte offset = 512 between 4-node swizzle atoms.  Both operands are split hi/lo in shared
// memory by warps 2-9.
// Accumulation order.  The tensor core adds into its fp32 accumulator with truncation, so the error of
// one TMEM accumulation chain grows linearly with its length.  The node rows are therefore cut into
// segments of seg_chunks k-chunks (<= 1024 rows; measured bias 3e-8 per accumulation, 6 per chunk): a CTA accumulates one segment in TMEM (two 128-lane
// halves), warps 10-13 drain it to its own slot of the workspace, and k_reduce_partials adds the slots
// in segment order in double precision (fixed association, bit-stable run to run).
// =====================================================================================================
// Per storage type: fp32 operands are split hi/lo in shared memory (TF32, SWIZZLE_128B_BASE32B, 16 node rows per
// chunk, 32 features per 128-byte group); bf16 operands are fed as stored (kind::f16, plain SWIZZLE_128B, 32 node rows
// per chunk, 64 features per group).  Either way one operand chunk is 16 KB for 256 features.
template <typename S>
struct Tn;
template <>
struct Tn<float> {
    static constexpr int BKR = 16, FEAT = 32, UKR = 8;    // node rows per chunk, features per group, rows per MMA
    static constexpr int STAGES = 3, OPS = 4;             // A raw | B raw | A lo | B lo
    static constexpr uint32_t SBO = 512, SWZ = 1;         // 4-row swizzle atoms, SWIZZLE_128B_BASE32B
    static constexpr CUtensorMapSwizzle TMA_SWZ = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
};
template <>
struct Tn<__nv_bfloat16> {
    static constexpr int BKR = 32, FEAT = 64, UKR = 16;
    static constexpr int STAGES = 6, OPS = 2;             // A | B
    static constexpr uint32_t SBO = 1024, SWZ = 2;        // 8-row swizzle atoms, SWIZZLE_128B
    static constexpr CUtensorMapSwizzle TMA_SWZ = CU_TENSOR_MAP_SWIZZLE_128B;
};
constexpr int TN_OP_BYTES = 16384;               // up to 256 features of one chunk: 16 KB per operand
template <typename S>
struct TnCfg {
    static constexpr int GROUP_BYTES = Tn<S>::BKR * 128;          // one feature group of a chunk
    static constexpr int STAGE_BYTES = Tn<S>::OPS * TN_OP_BYTES;
    static constexpr int BAR_OFF = Tn<S>::STAGES * STAGE_BYTES;
    static constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;
};
constexpr int TN_THREADS = 448;              // warp 0 TMA, warp 1 MMA, warps 2-9 split, warps 10-13 drain
constexpr int TN_MAX_SEG_CHUNKS = 64;   // 1024 rows: 384 accumulations per chain, ~1e-5 worst-case relative bias

struct TnArgs {
    int64_t n_chunks;   // ceil(M / TN_BK)
    int64_t n_segs;     // ceil(n_chunks / seg_chunks)
    int seg_chunks;
    int ka, nb;         // features of A / B handled by this launch (ka <= 256, nb <= 256, multiples of 32)
    int ga, gb;         // ka/32, nb/32
    float* partial;     // [n_segs, ka, nb]
    uint32_t lbo, sbo;  // descriptor byte offsets (host-computed)
    const float* row_scale;  // [M] or null: per-node scale applied to one operand while it is split
    int scale_op;            // 0 = A, 1 = B
    int64_t M;
};

// MN-major operand: swz = 1 SWIZZLE_128B_BASE32B (the only shared-memory layout for MN-major TF32 operands),
// swz = 2 SWIZZLE_128B (16-bit operands)
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t swz) {
    uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)swz << 61;
    return d;
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

template <typename S>
__global__ void __launch_bounds__(TN_THREADS, 1)
k_gemm_tn(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TnArgs g) {
    constexpr int TN_BK = Tn<S>::BKR;
    constexpr int TN_STAGES = Tn<S>::STAGES;
    constexpr int TN_GROUP_BYTES = TnCfg<S>::GROUP_BYTES;
    constexpr int TN_STAGE_BYTES = TnCfg<S>::STAGE_BYTES;
    constexpr int TN_BAR_OFF = TnCfg<S>::BAR_OFF;
    constexpr bool SPLIT = El<S>::SPLIT;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const uint32_t bar_base = smem_base + TN_BAR_OFF;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto ready_bar = [&](int s) { return bar_base + 8u * (TN_STAGES + s); };
    auto empty_bar = [&](int s) { return bar_base + 8u * (2 * TN_STAGES + s); };
    const uint32_t done_bar = bar_base + 8u * (3 * TN_STAGES);         // segment accumulated (tcgen05.commit)
    const uint32_t drained_bar = bar_base + 8u * (3 * TN_STAGES + 1);  // segment read out of TMEM (4 warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + TN_BAR_OFF + 8 * (3 * TN_STAGES + 2));

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&map_a);
            tma_prefetch_desc(&map_b);
            for (int s = 0; s < TN_STAGES; ++s) {
                mbar_init(full_bar(s), 1);
                mbar_init(ready_bar(s), 8);
                mbar_init(empty_bar(s), 1);
            }
            mbar_init(done_bar, 1);
            mbar_init(drained_bar, 4);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // this CTA's contiguous range of segments (host guarantees gridDim.x <= n_segs)
    const int64_t s_beg = g.n_segs * blockIdx.x / gridDim.x;
    const int64_t s_end = g.n_segs * (blockIdx.x + 1) / gridDim.x;
    const int64_t c_beg = s_beg * g.seg_chunks;
    const int64_t c_end = s_end * g.seg_chunks < g.n_chunks ? s_end * g.seg_chunks : g.n_chunks;
    const int a_bytes = g.ga * TN_GROUP_BYTES, b_bytes = g.gb * TN_GROUP_BYTES;
    const int halves = g.ka > 128 ? 2 : 1;

    if (warp == 0) {
        uint32_t it = 0;
        for (int64_t c = c_beg; c < c_end; ++c, ++it) {
            const int s = it % TN_STAGES;
            const uint32_t ph = (it / TN_STAGES) & 1u;
            mbar_wait(empty_bar(s), ph ^ 1u);
            if (lane == 0) {
                const uint32_t st = smem_base + (uint32_t)s * TN_STAGE_BYTES;
                mbar_arrive_expect_tx(full_bar(s), (uint32_t)(a_bytes + b_bytes));
                tma_load_3d(st, &map_a, full_bar(s), 0, (int)(c * TN_BK), 0);
                tma_load_3d(st + TN_OP_BYTES, &map_b, full_bar(s), 0, (int)(c * TN_BK), 0);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // A and B both MN-major (bits 15 / 16), N = nb
        const uint32_t idesc = (1u << 4) | (El<S>::FMT << 7) | (El<S>::FMT << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(g.nb >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        uint32_t it = 0, seg = 0;
        for (int64_t c = c_beg; c < c_end; ++c, ++it) {
            const int s = it % TN_STAGES;
            const uint32_t ph = (it / TN_STAGES) & 1u;
            const int pos = (int)((c - c_beg) % g.seg_chunks);
            const bool seg_last = pos == g.seg_chunks - 1 || c == c_end - 1;
            if (pos == 0 && seg > 0) {  // the previous segment must have left TMEM
                mbar_wait(drained_bar, (seg - 1) & 1u);
                tc_fence_after();
            }
            mbar_wait(full_bar(s), ph);
            if (SPLIT) mbar_wait(ready_bar(s), ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t st = smem_base + (uint32_t)s * TN_STAGE_BYTES;
                constexpr uint32_t SWZ = Tn<S>::SWZ;
                for (int h = 0; h < halves; ++h) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(h * 256);
                    const uint32_t a_off = (uint32_t)h * 8192u;     // 128 features further: half an operand chunk
#pragma unroll
                    for (int k = 0; k < TN_BK / Tn<S>::UKR; ++k) {
                        const uint32_t koff = (uint32_t)k * (uint32_t)(Tn<S>::UKR * 128);  // next node rows of one MMA
                        const uint64_t dah = smem_desc_mn_sw128(st + a_off + koff, g.lbo, g.sbo, SWZ);
                        const uint64_t dbh = smem_desc_mn_sw128(st + TN_OP_BYTES + koff, g.lbo, g.sbo, SWZ);
                        if (SPLIT) {
                            const uint64_t dal = smem_desc_mn_sw128(st + 2 * TN_OP_BYTES + a_off + koff, g.lbo, g.sbo, SWZ);
                            const uint64_t dbl = smem_desc_mn_sw128(st + 3 * TN_OP_BYTES + koff, g.lbo, g.sbo, SWZ);
                            umma<S>(d_tmem, dal, dbh, idesc, (pos | k) != 0);
                            umma<S>(d_tmem, dah, dbl, idesc, 1u);
                            umma<S>(d_tmem, dah, dbh, idesc, 1u);
                        } else {
                            umma<S>(d_tmem, dah, dbh, idesc, (pos | k) != 0);
                        }
                    }
                }
                umma_commit(empty_bar(s));
                if (seg_last) umma_commit(done_bar);
            }
            __syncwarp();
            if (seg_last) ++seg;
        }
    } else if (warp < 10) {
        // ===== split warps (2..9): hi in place, lo at +2*TN_OP_BYTES =====
        const int t = threadIdx.x - 64;  // 0..255
        uint32_t it = 0;
        for (int64_t c = c_beg; SPLIT && c < c_end; ++c, ++it) {      // bf16 operands: nothing to split
            const int s = it % TN_STAGES;
            const uint32_t ph = (it / TN_STAGES) & 1u;
            mbar_wait(full_bar(s), ph);
            uint8_t* st = smem_gen + (size_t)s * TN_STAGE_BYTES;
            for (int op = 0; op < 2; ++op) {
                uint8_t* hi_p = st + op * TN_OP_BYTES;
                uint8_t* lo_p = hi_p + 2 * TN_OP_BYTES;
                const int n16 = (op == 0 ? a_bytes : b_bytes) >> 4;
                const bool scaled = g.row_scale != nullptr && g.scale_op == op;
                for (int idx = t; idx < n16; idx += 256) {
                    float4 v = *reinterpret_cast<const float4*>(hi_p + 16 * idx);
                    if (scaled) {  // 16-byte chunk idx sits in node row (idx / 8) % 16 of the chunk (swizzle keeps rows)
                        const int64_t node = c * TN_BK + ((idx >> 3) & (TN_BK - 1));
                        const float sc = node < g.M ? __ldg(g.row_scale + node) : 0.f;
                        v.x = __fmul_rn(v.x, sc); v.y = __fmul_rn(v.y, sc);
                        v.z = __fmul_rn(v.z, sc); v.w = __fmul_rn(v.w, sc);
                    }
                    uint4 h, l;
                    split_tf32(v.x, h.x, l.x);
                    split_tf32(v.y, h.y, l.y);
                    split_tf32(v.z, h.z, l.z);
                    split_tf32(v.w, h.w, l.w);
                    *reinterpret_cast<uint4*>(hi_p + 16 * idx) = h;
                    *reinterpret_cast<uint4*>(lo_p + 16 * idx) = l;
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(ready_bar(s));
        }
    } else {
        // ===== drain warps (10..13): one finished segment [ka, nb] -> its workspace slot =====
        const int q = warp & 3;
        uint32_t seg = 0;
        for (int64_t sg = s_beg; sg < s_end; ++sg, ++seg) {
            mbar_wait(done_bar, seg & 1u);
            tc_fence_after();
            float* part = g.partial + (size_t)sg * g.ka * g.nb;
            for (int h = 0; h < halves; ++h) {
                const int row = h * 128 + q * 32 + lane;
                const int n_slabs = g.nb >> 5;          // 32 accumulator columns per tcgen05.ld
                for (int j = 0; j < n_slabs; ++j) {
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 256 + j * 32), r);
                    tmem_ld_wait();
                    if (h == halves - 1 && j == n_slabs - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(drained_bar);
                    }
                    if (row < g.ka) {
                        float* dst = part + (size_t)row * g.nb + j * 32;
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            __stcs(reinterpret_cast<uint4*>(dst + 4 * i),
                                   make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]));
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// Two-stage, fixed-order reduction of the segment partials (double accumulators):
//   stage 1: slice y of the partials -> mid[y, i]      stage 2: out[i] = sum_y mid[y, i]
constexpr int TN_REDUCE_SLICES = 32;

__global__ void __launch_bounds__(256) k_reduce_partials_1(const float* __restrict__ partial, int64_t n_part,
                                                           int total, double* __restrict__ mid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t p_beg = n_part * blockIdx.y / gridDim.y, p_end = n_part * (blockIdx.y + 1) / gridDim.y;
    double acc = 0.0;
    int64_t p = p_beg;
    for (; p + 8 <= p_end; p += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcs(partial + (size_t)(p + u) * total + i);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += (double)v[u];
    }
    for (; p < p_end; ++p) acc += (double)__ldcs(partial + (size_t)p * total + i);
    mid[(size_t)blockIdx.y * total + i] = acc;
}

__global__ void __launch_bounds__(256) k_reduce_partials_2(const double* __restrict__ mid, int n_slices, int ka,
                                                           int nb, float* __restrict__ out, int64_t ld_out) {
    const int total = ka * nb;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    double acc = 0.0;
    for (int y = 0; y < n_slices; ++y) acc += mid[(size_t)y * total + i];
    out[(int64_t)(i / nb) * ld_out + (i % nb)] = (float)acc;
}

// [rows, feat0 .. feat0 + FEAT*groups) of a row-major matrix of S viewed as {FEAT features (128 bytes), rows, groups}
template <typename S>
static int make_map_tn(CUtensorMap* map, const S* base, int64_t rows, int64_t ld, int groups) {
    EncodeTiledFn fn = encode_fn();
    CB_REQUIRE(fn != nullptr, CB_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t dims[3] = {(cuuint64_t)Tn<S>::FEAT, (cuuint64_t)rows, (cuuint64_t)groups};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(S), 128};
    const cuuint32_t box[3] = {(cuuint32_t)Tn<S>::FEAT, (cuuint32_t)Tn<S>::BKR, (cuuint32_t)groups};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(map, El<S>::TMA_T, 3, const_cast<S*>(base), dims, strides, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, Tn<S>::TMA_SWZ,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (3-D) failed with CUresult " + std::to_string((int)r));
        return CB_E_CUDA;
    }
    return CB_OK;
}

}  // namespace tc
}  // namespace cb

extern "C" {

int cb_gemm_split_weight(const float* W, int64_t n_rows, int64_t k_cols, int transpose, float* hi, float* lo,
                         void* stream) {
    using namespace cb;
    CB_REQUIRE(W && hi && lo, CB_E_INVALID, "cb_gemm_split_weight: NULL buffer");
    CB_REQUIRE(n_rows > 0 && k_cols > 0 && n_rows * k_cols < (int64_t)1 << 31, CB_E_INVALID,
               "cb_gemm_split_weight: bad shape");
    const int64_t total = n_rows * k_cols;
    const int blocks = (int)(ceil_div(total, 256) < 1184 ? ceil_div(total, 256) : 1184);
    tc::k_split_weight<<<blocks, 256, 0, (cudaStream_t)stream>>>(W, (int)n_rows, (int)k_cols, transpose, hi, lo);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_gemm_rows_supported(int64_t M, int64_t N, int64_t K) { return cb::tc::rows_supported<float>(M, N, K); }
int cb_gemm_rows_supported_bf16(int64_t M, int64_t N, int64_t K) {
    return cb::tc::rows_supported<__nv_bfloat16>(M, N, K);
}

int cb_gemm_rows(const float* A, int64_t M, int64_t K, int64_t lda, const float* Bt_hi, const float* Bt_lo,
                 int64_t N, const float* row_scale, const float* bias, const float* add, int64_t ld_add, int act,
                 float* out, int64_t ld_out, const float* out2_scale, float* out2, int64_t ld_out2,
                 const cb_peer_push_t* push, void* stream) {
    return cb::tc::gemm_rows_impl<float>(A, M, K, lda, Bt_hi, Bt_lo, N, row_scale, bias, add, ld_add, act, out, ld_out,
                                         out2_scale, out2, ld_out2, push, stream);
}

int cb_gemm_rows_bf16(const uint16_t* A, int64_t M, int64_t K, int64_t lda, const uint16_t* Bt, int64_t N,
                      const float* row_scale, const float* bias, const uint16_t* add, int64_t ld_add, int act,
                      uint16_t* out, int64_t ld_out, const float* out2_scale, uint16_t* out2, int64_t ld_out2,
                      const cb_peer_push_t* push, void* stream) {
    using B = __nv_bfloat16;
    return cb::tc::gemm_rows_impl<B>((const B*)A, M, K, lda, (const B*)Bt, nullptr, N, row_scale, bias, (const B*)add,
                                     ld_add, act, (B*)out, ld_out, out2_scale, (B*)out2, ld_out2, push, stream);
}

int cb_gemm_rows_masked(int dtype, const void* A, int64_t M, int64_t K, int64_t lda, const void* Bt_hi, const void* Bt_lo,
                        int64_t N, const float* row_scale, const float* bias, const void* add, int64_t ld_add, int act,
                        void* out, int64_t ld_out, const float* out2_scale, void* out2, int64_t ld_out2,
                        uint8_t* relu_mask, int64_t ld_mask, const cb_peer_push_t* push, void* stream) {
    using B = __nv_bfloat16;
    CB_REQUIRE(dtype == CB_F32 || dtype == CB_BF16, CB_E_INVALID, "cb_gemm_rows_masked: unknown dtype");
    if (dtype == CB_F32)
        return cb::tc::gemm_rows_impl<float>((const float*)A, M, K, lda, (const float*)Bt_hi, (const float*)Bt_lo, N,
                                             row_scale, bias, (const float*)add, ld_add, act, (float*)out, ld_out,
                                             out2_scale, (float*)out2, ld_out2, push, stream, relu_mask, ld_mask);
    return cb::tc::gemm_rows_impl<B>((const B*)A, M, K, lda, (const B*)Bt_hi, nullptr, N, row_scale, bias, (const B*)add,
                                     ld_add, act, (B*)out, ld_out, out2_scale, (B*)out2, ld_out2, push, stream,
                                     relu_mask, ld_mask);
}

int cb_gemm_weight_to_bf16(const float* W, int64_t n_rows, int64_t k_cols, int transpose, uint16_t* out, void* stream) {
    using namespace cb;
    CB_REQUIRE(W && out, CB_E_INVALID, "cb_gemm_weight_to_bf16: NULL buffer");
    CB_REQUIRE(n_rows > 0 && k_cols > 0 && n_rows * k_cols < (int64_t)1 << 31, CB_E_INVALID,
               "cb_gemm_weight_to_bf16: bad shape");
    const int64_t total = n_rows * k_cols;
    const int blocks = (int)(ceil_div(total, 256) < 1184 ? ceil_div(total, 256) : 1184);
    tc::k_weight_to_bf16<<<blocks, 256, 0, (cudaStream_t)stream>>>(W, (int)n_rows, (int)k_cols, transpose,
                                                                   (__nv_bfloat16*)out);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int64_t cb_gemm_rows_grad_workspace_bytes(int64_t M, int64_t N) {
    if (M <= 0 || N <= 0) return 0;
    return cb::tc::gemm_grid_x(cb::ceil_div(M, cb::tc::BM)) * N * (int64_t)sizeof(float);
}

int cb_gemm_rows_grad(const float* A, int64_t M, int64_t K, int64_t lda, const float* Bt_hi, const float* Bt_lo,
                      int64_t N, const float* row_scale, const float* add, int64_t ld_add, const uint8_t* gate_u8,
                      const float* gate_f32, int64_t ld_gate, int mixed, double alpha, float* d_x0, int64_t ld_dx0,
                      int accumulate_x0, const float* post_scale, float* out, int64_t ld_out, float* col_sum,
                      uint8_t* row_live, const uint8_t* a_live, const uint8_t* x0_valid, void* workspace,
                      int64_t workspace_bytes, const cb_peer_push_t* push, void* stream) {
    return cb::tc::gemm_rows_grad_impl<float>(A, M, K, lda, Bt_hi, Bt_lo, N, row_scale, add, ld_add, gate_u8, gate_f32,
                                              ld_gate, mixed, alpha, d_x0, ld_dx0, accumulate_x0, post_scale, out,
                                              ld_out, col_sum, row_live, a_live, x0_valid, workspace, workspace_bytes,
                                              push, stream);
}

int cb_gemm_rows_grad_bf16(const uint16_t* A, int64_t M, int64_t K, int64_t lda, const uint16_t* Bt, int64_t N,
                           const float* row_scale, const uint16_t* add, int64_t ld_add, const uint8_t* gate_u8,
                           const uint16_t* gate_val, int64_t ld_gate, int mixed, double alpha, uint16_t* d_x0,
                           int64_t ld_dx0, int accumulate_x0, const float* post_scale, uint16_t* out, int64_t ld_out,
                           float* col_sum, uint8_t* row_live, const uint8_t* a_live, const uint8_t* x0_valid,
                           void* workspace, int64_t workspace_bytes, const cb_peer_push_t* push, void* stream) {
    using B = __nv_bfloat16;
    return cb::tc::gemm_rows_grad_impl<B>((const B*)A, M, K, lda, (const B*)Bt, nullptr, N, row_scale, (const B*)add,
                                          ld_add, gate_u8, (const B*)gate_val, ld_gate, mixed, alpha, (B*)d_x0, ld_dx0,
                                          accumulate_x0, post_scale, (B*)out, ld_out, col_sum, row_live, a_live, x0_valid,
                                          workspace, workspace_bytes, push, stream);
}

}  // extern "C" (reopened below)

namespace cb {
namespace tc {

template <typename S>
static int tn_supported(int64_t M, int64_t Ka, int64_t Nb) {
    return M > 0 && Ka > 0 && Nb > 0 && Ka % Tn<S>::FEAT == 0 && Nb % Tn<S>::FEAT == 0 &&
           M < ((int64_t)1 << 31) - 64 && Ka <= 4096 && Nb <= 4096;
}

template <typename S>
static void tn_plan(int64_t M, int64_t* chunks, int* seg_chunks, int64_t* n_segs, int* ctas) {
    const int64_t ch = ceil_div(M, Tn<S>::BKR);
    const int sms = sm_count();
    int64_t sc = ceil_div(ch, sms);
    // Chain length vs workspace.  The accumulator error grows with the number of MMAs chained into one TMEM
    // accumulation (measured on B200, fp32 operands, M = 1.7e5: max error / sum|a||b| = 3.2e-7 / 6.8e-7 / 4.4e-6 for
    // segments of 4 / 16 / 64 chunks; cuBLAS sgemm: 2.4e-6), every segment costs one [Ka, Nb] partial in HBM.  Up to
    // 8 segments per SM the segments stay at <= 16 chunks (256 rows); beyond that they grow to the cap of 64 chunks
    // (the partials of BASELINE configs[3], 10^7 rows, are then 12 % of the operand bytes).
    static const int64_t forced = getenv("CB_TN_SEG_CHUNKS") ? atoll(getenv("CB_TN_SEG_CHUNKS")) : 0;
    int64_t max_sc = forced > 0 ? forced : ceil_div(ch, (int64_t)8 * sms);
    if (forced <= 0) max_sc = max_sc < 16 ? 16 : (max_sc > TN_MAX_SEG_CHUNKS ? TN_MAX_SEG_CHUNKS : max_sc);
    if (sc > max_sc) sc = max_sc;
    if (sc < 1) sc = 1;
    *chunks = ch;
    *seg_chunks = (int)sc;
    *n_segs = ceil_div(ch, sc);
    *ctas = (int)(*n_segs < sms ? *n_segs : sms);
}

template <typename S>
static int64_t tn_workspace_bytes(int64_t M, int64_t Ka, int64_t Nb) {
    if (!tn_supported<S>(M, Ka, Nb)) return 0;
    int64_t chunks, n_segs;
    int sc, ctas;
    tn_plan<S>(M, &chunks, &sc, &n_segs, &ctas);
    const int64_t ka = Ka < 256 ? Ka : 256, nb = Nb < 256 ? Nb : 256;
    return n_segs * ka * nb * (int64_t)sizeof(float) + TN_REDUCE_SLICES * ka * nb * (int64_t)sizeof(double);
}

template <typename S>
static int gemm_tn_impl(const S* A, int64_t lda, const S* B, int64_t ldb, int64_t M, int64_t Ka, int64_t Nb,
                        const float* row_scale, int scale_b, float* out, int64_t ld_out, void* workspace,
                        int64_t workspace_bytes, void* stream) {
    CB_REQUIRE(A && B && out, CB_E_INVALID, "cb_gemm_tn: NULL buffer");
    CB_REQUIRE(tn_supported<S>(M, Ka, Nb), CB_E_UNSUPPORTED,
               "cb_gemm_tn: needs Ka and Nb multiples of one 128-byte feature group and M < 2^31");
    CB_REQUIRE(!row_scale || El<S>::SPLIT, CB_E_UNSUPPORTED, "cb_gemm_tn: the bf16 variant takes pre-scaled operands");
    CB_REQUIRE(al16(A) && al16(B) && lda % vec16<S>() == 0 && ldb % vec16<S>() == 0 && lda >= Ka && ldb >= Nb &&
                   ld_out >= Nb,
               CB_E_UNSUPPORTED, "cb_gemm_tn: operands must be 16-byte aligned, pitches multiples of 16 bytes");
    CB_REQUIRE(workspace && workspace_bytes >= tn_workspace_bytes<S>(M, Ka, Nb), CB_E_WORKSPACE,
               "cb_gemm_tn: workspace missing or smaller than cb_gemm_tn_workspace_bytes()");
    static std::atomic<bool> configured[64];   // per device ordinal (the attribute is per device)
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
        CB_CUDA(cudaFuncSetAttribute(k_gemm_tn<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, TnCfg<S>::SMEM_BYTES));
        if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
    }
    cudaStream_t st = (cudaStream_t)stream;
    int64_t chunks, n_segs;
    int seg_chunks, ctas;
    tn_plan<S>(M, &chunks, &seg_chunks, &n_segs, &ctas);
    for (int64_t a0 = 0; a0 < Ka; a0 += 256) {
        for (int64_t b0 = 0; b0 < Nb; b0 += 256) {
            TnArgs g{};
            g.n_chunks = chunks;
            g.n_segs = n_segs;
            g.seg_chunks = seg_chunks;
            g.ka = (int)(Ka - a0 < 256 ? Ka - a0 : 256);
            g.nb = (int)(Nb - b0 < 256 ? Nb - b0 : 256);
            g.ga = g.ka / Tn<S>::FEAT;
            g.gb = g.nb / Tn<S>::FEAT;
            g.partial = (float*)workspace;
            g.row_scale = row_scale;
            g.scale_op = scale_b ? 1 : 0;
            g.M = M;
            g.lbo = (uint32_t)TnCfg<S>::GROUP_BYTES;   // between feature groups
            g.sbo = Tn<S>::SBO;                        // between swizzle atoms of node rows
            CUtensorMap ma, mb;
            int rc = make_map_tn<S>(&ma, A + a0, M, lda, g.ga);
            if (rc) return rc;
            rc = make_map_tn<S>(&mb, B + b0, M, ldb, g.gb);
            if (rc) return rc;
            k_gemm_tn<S><<<ctas, TN_THREADS, TnCfg<S>::SMEM_BYTES, st>>>(ma, mb, g);
            CB_LAUNCH_CHECK();
            const int total = g.ka * g.nb;
            const int slices = (int)(n_segs < TN_REDUCE_SLICES ? n_segs : TN_REDUCE_SLICES);
            double* mid = reinterpret_cast<double*>(g.partial + (size_t)n_segs * total);   // 8-byte aligned: total % 1024 == 0
            k_reduce_partials_1<<<dim3((total + 255) / 256, slices), 256, 0, st>>>(g.partial, n_segs, total, mid);
            CB_LAUNCH_CHECK();
            k_reduce_partials_2<<<(total + 255) / 256, 256, 0, st>>>(mid, slices, g.ka, g.nb,
                                                                    out + a0 * ld_out + b0, ld_out);
            CB_LAUNCH_CHECK();
        }
    }
    return CB_OK;
}

}  // namespace tc
}  // namespace cb

extern "C" {

int cb_gemm_tn_supported(int64_t M, int64_t Ka, int64_t Nb) { return cb::tc::tn_supported<float>(M, Ka, Nb); }
int cb_gemm_tn_supported_bf16(int64_t M, int64_t Ka, int64_t Nb) {
    return cb::tc::tn_supported<__nv_bfloat16>(M, Ka, Nb);
}
int64_t cb_gemm_tn_workspace_bytes(int64_t M, int64_t Ka, int64_t Nb) {
    return cb::tc::tn_workspace_bytes<float>(M, Ka, Nb);
}
int64_t cb_gemm_tn_workspace_bytes_bf16(int64_t M, int64_t Ka, int64_t Nb) {
    return cb::tc::tn_workspace_bytes<__nv_bfloat16>(M, Ka, Nb);
}

int cb_gemm_tn(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t Ka, int64_t Nb,
               const float* row_scale, int scale_b, float* out, int64_t ld_out, void* workspace,
               int64_t workspace_bytes, void* stream) {
    return cb::tc::gemm_tn_impl<float>(A, lda, B, ldb, M, Ka, Nb, row_scale, scale_b, out, ld_out, workspace,
                                       workspace_bytes, stream);
}

int cb_gemm_tn_bf16(const uint16_t* A, int64_t lda, const uint16_t* B, int64_t ldb, int64_t M, int64_t Ka, int64_t Nb,
                    float* out, int64_t ld_out, void* workspace, int64_t workspace_bytes, void* stream) {
    using T = __nv_bfloat16;
    return cb::tc::gemm_tn_impl<T>((const T*)A, lda, (const T*)B, ldb, M, Ka, Nb, nullptr, 0, out, ld_out, workspace,
                                   workspace_bytes, stream);
}

}  // extern "C"
