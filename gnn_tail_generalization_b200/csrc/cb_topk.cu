// SEMLP "virtual neighbour" replacement (MLP_model/__init__.py:143-156): for every query row the K
// largest scores <le_guess[i], teacherSE[j]> over ALL teacher rows j, their soft-max, and the soft-max-weighted sum of
// those K teacher rows.  The [B, N] score matrix is never materialised: the caller computes it tile by tile with
// cb_gemm_rows (tcgen05, 3xTF32: fp32-class scores, so the selection is the reference's) into a small, L2-sized
// buffer and k_topk_merge folds each tile into a running per-row top-K list.
//
//   k_topk_merge        one warp per query row; lane l holds the l-th largest (score, teacher id) so far.  A tile is
//                       scanned 32 scores at a time against the current K-th largest; the rare survivors are inserted
//                       with one ballot (position) and one shuffle (shift) each.
//   k_topk_softmax_mix  one warp per query row: soft-max over the K kept scores and the weighted sum of the K rows,
//                       added in ascending score order like the reference's [1, K] x [K, d] product.
#include <math.h>

#include "cb_internal.cuh"

namespace cb {

__global__ void __launch_bounds__(256) k_topk_merge(const float* __restrict__ scores, int64_t B, int width, int64_t ld,
                                                    int64_t col_base, int K, float* __restrict__ top_val,
                                                    int32_t* __restrict__ top_idx, int first) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= B) return;
    float lv = -INFINITY;
    int32_t li = -1;
    if (!first) {
        lv = top_val[row * 32 + lane];
        li = top_idx[row * 32 + lane];
    }
    float tau = __shfl_sync(0xffffffffu, lv, K - 1);
    const float* srow = scores + row * ld;
    for (int base = 0; base < width; base += 32) {
        const int c = base + lane;
        const float v = c < width ? __ldcs(srow + c) : -INFINITY;
        unsigned m = __ballot_sync(0xffffffffu, v > tau);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float cv = __shfl_sync(0xffffffffu, v, src);
            if (cv > tau) {   // warp-uniform: tau may have risen since the ballot
                const unsigned ge = __ballot_sync(0xffffffffu, lane < K && lv >= cv);
                const int pos = __popc(ge);                   // after every kept score >= cv
                const float pv = __shfl_up_sync(0xffffffffu, lv, 1);
                const int32_t pi = __shfl_up_sync(0xffffffffu, li, 1);
                if (lane > pos && lane < K) { lv = pv; li = pi; }
                if (lane == pos) { lv = cv; li = (int32_t)(col_base + base + src); }
                tau = __shfl_sync(0xffffffffu, lv, K - 1);
            }
        }
    }
    top_val[row * 32 + lane] = lv;
    top_idx[row * 32 + lane] = li;
}

__global__ void __launch_bounds__(256) k_topk_softmax_mix(const float* __restrict__ top_val,
                                                          const int32_t* __restrict__ top_idx, int64_t B, int K,
                                                          const float* __restrict__ table, int64_t d, int64_t ld_table,
                                                          float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= B) return;
    const float v = lane < K ? top_val[row * 32 + lane] : -INFINITY;
    const int32_t id = lane < K ? top_idx[row * 32 + lane] : -1;
    const float mx = __shfl_sync(0xffffffffu, v, 0);          // lane 0 holds the largest
    const float e = (lane < K && id >= 0) ? expf(v - mx) : 0.f;
    float sum = 0.f;
    for (int k = K - 1; k >= 0; --k) sum += __shfl_sync(0xffffffffu, e, k);    // ascending score order
    const float p = e / sum;
    for (int64_t c0 = 0; c0 < d; c0 += 32) {       // warp-uniform trip count: the shuffles below need every lane
        const int64_t c = c0 + lane;
        float acc = 0.f;
        for (int k = K - 1; k >= 0; --k) {
            const float pk = __shfl_sync(0xffffffffu, p, k);
            const int32_t ik = __shfl_sync(0xffffffffu, id, k);
            if (ik >= 0 && c < d) acc += pk * __ldg(table + (int64_t)ik * ld_table + c);
        }
        if (c < d) out[row * d + c] = acc;
    }
}

}  // namespace cb

extern "C" {

int cb_topk_merge(const float* scores, int64_t B, int64_t width, int64_t ld, int64_t col_base, int K, float* top_val,
                  int32_t* top_idx, int first, void* stream) {
    using namespace cb;
    CB_REQUIRE(B >= 0 && width >= 0 && ld >= width, CB_E_INVALID, "cb_topk_merge: bad shape");
    CB_REQUIRE(K >= 1 && K <= 32, CB_E_UNSUPPORTED, "cb_topk_merge: K must be 1..32");
    CB_REQUIRE(width < ((int64_t)1 << 31) && col_base + width < ((int64_t)1 << 31), CB_E_UNSUPPORTED,
               "cb_topk_merge: teacher ids must fit int32");
    if (B == 0) return CB_OK;
    CB_REQUIRE(scores && top_val && top_idx, CB_E_INVALID, "cb_topk_merge: NULL buffer");
    k_topk_merge<<<(unsigned)ceil_div(B, 8), 256, 0, (cudaStream_t)stream>>>(scores, B, (int)width, ld, col_base, K,
                                                                             top_val, top_idx, first);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_topk_softmax_mix(const float* top_val, const int32_t* top_idx, int64_t B, int K, const float* table, int64_t d,
                        int64_t ld_table, float* out, void* stream) {
    using namespace cb;
    CB_REQUIRE(B >= 0 && d > 0 && ld_table >= d, CB_E_INVALID, "cb_topk_softmax_mix: bad shape");
    CB_REQUIRE(K >= 1 && K <= 32, CB_E_UNSUPPORTED, "cb_topk_softmax_mix: K must be 1..32");
    if (B == 0) return CB_OK;
    CB_REQUIRE(top_val && top_idx && table && out, CB_E_INVALID, "cb_topk_softmax_mix: NULL buffer");
    k_topk_softmax_mix<<<(unsigned)ceil_div(B, 8), 256, 0, (cudaStream_t)stream>>>(top_val, top_idx, B, K, table, d,
                                                                                   ld_table, out);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

}  // extern "C"
