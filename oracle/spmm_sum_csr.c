/*
 * C restatement of the CPU aggregation the reference reaches through DGL  --  TEST INFRASTRUCTURE.
 *
 * Reference call site: /root/reference/GNN_model/GCN.py:198,238
 *     graph.update_all(fn.copy_src('h','m'), fn.sum('m','h'))
 * which in dgl==0.7.0 (requirements.txt:19; not vendored, not installable here) dispatches to
 * gSpMM('copy_lhs','sum') and, on CPU, to SpMMSumCsr (dgl/src/array/cpu/spmm.h): an OpenMP
 * row-parallel loop; for each destination row, for each stored in-edge in CSR order, for each
 * feature k:  out[row,k] += X[col,k].  Every (row,k) accumulator is therefore a strictly
 * sequential fp32 sum in CSR order, independent of the thread count, which is what this file does.
 *
 * hub_chunk > 0 additionally reproduces the association used by the CUDA path for rows longer
 * than hub_chunk (chunk partials in order, then partials in order) so that parity can be
 * checked bit-for-bit instead of within a tolerance.
 *
 * cb_oracle_spmm_mul_sum_csr restates the edge-weighted form (GCN.py:199-202: fn.u_mul_e('h', '_edge_weight', 'm') ->
 * gSpMM('mul', 'sum'), SpMMSumCsr with a binary op): out[row,k] += X[col,k] * w[edge], the product rounded before it
 * is added (compiled with -ffp-contract=off), same hub-chunk association.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int cb_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static void row_sum(const int64_t* cols, int64_t beg, int64_t end, const float* X, int64_t d, float* acc) {
    for (int64_t k = 0; k < d; ++k) acc[k] = 0.0f;
    for (int64_t j = beg; j < end; ++j) {
        const float* x = X + cols[j] * d;
        for (int64_t k = 0; k < d; ++k) acc[k] += x[k];
    }
}

/* out[r,:] = sum_{j in [rowptr[r], rowptr[r+1])} X[cols[j],:]   (fp32, in order) */
int cb_oracle_spmm_sum_csr(const int64_t* rowptr, const int64_t* cols, const float* X, int64_t n_src,
                           int64_t d, float* out, int64_t n_rows, int64_t hub_chunk, int threads) {
    if (!rowptr || !X || !out || d <= 0) return -1;
    int bad = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
#pragma omp parallel
    {
        float* part = (float*)malloc(sizeof(float) * (size_t)d);
#pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < n_rows; ++r) {
            const int64_t beg = rowptr[r], end = rowptr[r + 1];
            float* o = out + r * d;
            for (int64_t j = beg; j < end; ++j)
                if (cols[j] < 0 || cols[j] >= n_src) bad = 1;
            if (bad) continue;
            if (hub_chunk <= 0 || end - beg <= hub_chunk) {
                row_sum(cols, beg, end, X, d, o);
            } else {
                for (int64_t k = 0; k < d; ++k) o[k] = 0.0f;
                for (int64_t b = beg; b < end; b += hub_chunk) {
                    const int64_t e = (b + hub_chunk < end) ? b + hub_chunk : end;
                    row_sum(cols, b, e, X, d, part);
                    for (int64_t k = 0; k < d; ++k) o[k] += part[k];
                }
            }
        }
        free(part);
    }
    return bad ? -2 : 0;
}

static void row_mul_sum(const int64_t* cols, const float* vals, int64_t beg, int64_t end, const float* X, int64_t d,
                        float* acc) {
    for (int64_t k = 0; k < d; ++k) acc[k] = 0.0f;
    for (int64_t j = beg; j < end; ++j) {
        const float* x = X + cols[j] * d;
        const float w = vals[j];
        for (int64_t k = 0; k < d; ++k) {
            const float m = x[k] * w;
            acc[k] += m;
        }
    }
}

/* out[r,:] = sum_{j in [rowptr[r], rowptr[r+1])} X[cols[j],:] * vals[j]   (fp32, product rounded, summed in order) */
int cb_oracle_spmm_mul_sum_csr(const int64_t* rowptr, const int64_t* cols, const float* vals, const float* X,
                               int64_t n_src, int64_t d, float* out, int64_t n_rows, int64_t hub_chunk, int threads) {
    if (!rowptr || !X || !out || !vals || d <= 0) return -1;
    int bad = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
#pragma omp parallel
    {
        float* part = (float*)malloc(sizeof(float) * (size_t)d);
#pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < n_rows; ++r) {
            const int64_t beg = rowptr[r], end = rowptr[r + 1];
            float* o = out + r * d;
            for (int64_t j = beg; j < end; ++j)
                if (cols[j] < 0 || cols[j] >= n_src) bad = 1;
            if (bad) continue;
            if (hub_chunk <= 0 || end - beg <= hub_chunk) {
                row_mul_sum(cols, vals, beg, end, X, d, o);
            } else {
                for (int64_t k = 0; k < d; ++k) o[k] = 0.0f;
                for (int64_t b = beg; b < end; b += hub_chunk) {
                    const int64_t e = (b + hub_chunk < end) ? b + hub_chunk : end;
                    row_mul_sum(cols, vals, b, e, X, d, part);
                    for (int64_t k = 0; k < d; ++k) o[k] += part[k];
                }
            }
        }
        free(part);
    }
    return bad ? -2 : 0;
}
