// Streaming (one pass, HBM-bound) kernels either side of the aggregation:
//   cb_row_scale          GCN.py:205-213   feat_src * out_deg^-1/2          (and its adjoint)
//   cb_sumsq              GCN.py:232       th.norm(self.le)                 (sum of squares; sqrt on the caller side)
//   cb_agg_backward_prep  autograd of GCN.py:242-253, 127-128 and res_tricks.py:14/23
// plus the library-wide bookkeeping (error text, launch counter).
#include "cb_internal.cuh"

namespace cb {

static thread_local std::string t_error;
std::atomic<int64_t> g_launches{0};

void set_error(const std::string& msg) { t_error = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    t_error = std::string(what) + " failed: " + cudaGetErrorString(e) + " (" + file + ":" + std::to_string(line) + ")";
    return CB_E_CUDA;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

// ---------------------------------------------------------------------------------------------
// y = s[r] * x
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) k_row_scale(const float* __restrict__ x, const float* __restrict__ s,
                                                   int64_t rows, int64_t units, float* __restrict__ y) {
    const int64_t total = rows * units;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const float sc = __ldg(s + i / units);
        if (VEC == 4) {
            float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
            v.x = __fmul_rn(v.x, sc); v.y = __fmul_rn(v.y, sc); v.z = __fmul_rn(v.z, sc); v.w = __fmul_rn(v.w, sc);
            reinterpret_cast<float4*>(y)[i] = v;
        } else {
            y[i] = __fmul_rn(__ldg(x + i), sc);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// sum of squares, fixed association: thread-strided partial -> block tree -> one final block
// ---------------------------------------------------------------------------------------------
constexpr int SUMSQ_MAX_BLOCKS = 2048;

__device__ __forceinline__ float block_sum_256(float v) {
    __shared__ float sm[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 32) {
        t = threadIdx.x < 8 ? sm[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    __syncthreads();
    return t;  // valid in thread 0
}

__global__ void __launch_bounds__(256) k_sumsq_partial(const float* __restrict__ x, int64_t n,
                                                       float* __restrict__ partial) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.f;
    const bool vec = (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
    const int64_t n4 = vec ? n / 4 : 0;
    for (int64_t i = tid; i < n4; i += stride) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int64_t i = n4 * 4 + tid; i < n; i += stride) {
        const float v = __ldg(x + i);
        acc += v * v;
    }
    const float t = block_sum_256(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void __launch_bounds__(256) k_sumsq_final(const float* __restrict__ partial, int nblocks,
                                                     float* __restrict__ out) {
    // double accumulation over <= 2048 partials: removes the last-stage rounding from the picture
    double acc = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) acc += (double)partial[i];
    __shared__ double sm[256];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (float)sm[0];
}

// ---------------------------------------------------------------------------------------------
// backward prologue
// ---------------------------------------------------------------------------------------------
struct PrepArgs {
    const float* d_out;
    const float* d_out2;
    const float* s2;       // dout^-1/2 (scale of d_out2)
    const float* rs;       // din^-1/2
    const uint8_t* mask;
    const float* relu_out;
    int act, mixed;
    float alpha, one_minus_alpha;
    float* G;
    float* d_x0;
    int accumulate_x0;
    float* bias_partial;   // [gridDim.x, d] or null
    int64_t rows, d;
    int64_t rows_per_block;
};

// Block layout: CU = min(units, 256) column slots side by side, RL = 256 / CU row lanes.  Every thread
// keeps a running bias-gradient sum for its column slot over the rows it visits; the row lanes of a
// block are then added in lane order, giving one partial row per block.
template <int VEC>
__global__ void __launch_bounds__(256) k_prep(const PrepArgs a) {
    extern __shared__ float sm[];  // [RL][CU*VEC] when bias partials are wanted
    const int64_t units = (a.d + VEC - 1) / VEC;
    const int cu = (int)(units < 256 ? units : 256);
    const int rl = 256 / cu;
    const int cslot = threadIdx.x % cu;
    const int rlane = threadIdx.x / cu;
    const bool active = rlane < rl;
    const int64_t r0 = (int64_t)blockIdx.x * a.rows_per_block;
    const int64_t r1 = r0 + a.rows_per_block < a.rows ? r0 + a.rows_per_block : a.rows;

    for (int64_t ub = 0; ub < units; ub += cu) {
        const int64_t c = (ub + cslot) * VEC;
        const bool cok = active && c < a.d;
        float bsum[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) bsum[i] = 0.f;
        if (cok) {
            for (int64_t r = r0 + rlane; r < r1; r += rl) {
                const int64_t off = r * a.d + c;
                float dt[VEC];
#pragma unroll
                for (int i = 0; i < VEC; ++i) dt[i] = 0.f;
                if (a.d_out) {
                    if (VEC == 4) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(a.d_out + off));
                        dt[0] = v.x; dt[1 % VEC] = v.y; dt[2 % VEC] = v.z; dt[3 % VEC] = v.w;
                    } else {
                        dt[0] = __ldg(a.d_out + off);
                    }
                }
                if (a.d_out2) {
                    const float s2 = __ldg(a.s2 + r);
                    if (VEC == 4) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(a.d_out2 + off));
                        const float t[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int i = 0; i < VEC; ++i)
                            dt[i] = a.d_out ? __fadd_rn(dt[i], __fmul_rn(s2, t[i])) : __fmul_rn(s2, t[i]);
                    } else {
                        const float t = __ldg(a.d_out2 + off);
                        dt[0] = a.d_out ? __fadd_rn(dt[0], __fmul_rn(s2, t)) : __fmul_rn(s2, t);
                    }
                }
                if (a.d_x0) {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const float g0 = __fmul_rn(a.alpha, dt[i]);
                        a.d_x0[off + i] = a.accumulate_x0 ? __fadd_rn(a.d_x0[off + i], g0) : g0;
                    }
                }
                bool m[VEC];
#pragma unroll
                for (int i = 0; i < VEC; ++i) m[i] = true;
                if (a.act == CB_ACT_RELU) {
                    if (a.mask) {
                        if (VEC == 4) {
                            const uchar4 mm = *reinterpret_cast<const uchar4*>(a.mask + off);
                            m[0] = mm.x; m[1 % VEC] = mm.y; m[2 % VEC] = mm.z; m[3 % VEC] = mm.w;
                        } else {
                            m[0] = a.mask[off] != 0;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) m[i] = __ldg(a.relu_out + off + i) > 0.f;
                    }
                }
                const float rs = __ldg(a.rs + r);
                float g[VEC];
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    float dz = a.mixed ? __fmul_rn(a.one_minus_alpha, dt[i]) : dt[i];
                    dz = m[i] ? dz : 0.f;
                    bsum[i] += dz;
                    g[i] = __fmul_rn(rs, dz);
                }
                if (VEC == 4) {
                    *reinterpret_cast<float4*>(a.G + off) = make_float4(g[0], g[1 % VEC], g[2 % VEC], g[3 % VEC]);
                } else {
                    a.G[off] = g[0];
                }
            }
        }
        if (a.bias_partial) {
            if (active) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) sm[(rlane * cu + cslot) * VEC + i] = bsum[i];
            }
            __syncthreads();
            if (rlane == 0 && c < a.d) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    float t = 0.f;
                    for (int l = 0; l < rl; ++l) t += sm[(l * cu + cslot) * VEC + i];
                    a.bias_partial[(int64_t)blockIdx.x * a.d + c + i] = t;
                }
            }
            __syncthreads();
        }
    }
}

// d_bias[c] = sum over block partials, in block order
__global__ void __launch_bounds__(256) k_bias_final(const float* __restrict__ partial, int nblocks, int64_t d,
                                                    float* __restrict__ d_bias) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d) return;
    float t = 0.f;
    for (int b = 0; b < nblocks; ++b) t += partial[(int64_t)b * d + c];
    d_bias[c] = t;
}

static int prep_blocks(int64_t rows) {
    const int64_t cap = (int64_t)sm_count() * 8;
    const int64_t want = ceil_div(rows > 0 ? rows : 1, 64);
    return (int)(want < cap ? want : cap);
}

}  // namespace cb

extern "C" {

const char* cb_last_error(void) { return cb::t_error.c_str(); }
int cb_abi_version(void) { return CB_ABI_VERSION; }
int64_t cb_launch_count(void) { return cb::g_launches.load(); }

int cb_row_scale(const float* x, const float* s, int64_t rows, int64_t d, float* y, void* stream) {
    using namespace cb;
    CB_REQUIRE(rows >= 0 && d > 0, CB_E_INVALID, "cb_row_scale: bad shape");
    if (rows == 0) return CB_OK;
    CB_REQUIRE(x && s && y, CB_E_INVALID, "cb_row_scale: NULL buffer");
    const bool vec = d % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0;
    const int64_t units = vec ? d / 4 : d;
    const int64_t total = rows * units;
    const int64_t cap = (int64_t)sm_count() * 32;
    const int64_t want = ceil_div(total, 256);
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (vec)
        k_row_scale<4><<<grid, 256, 0, (cudaStream_t)stream>>>(x, s, rows, units, y);
    else
        k_row_scale<1><<<grid, 256, 0, (cudaStream_t)stream>>>(x, s, rows, units, y);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int64_t cb_sumsq_workspace_bytes(void) { return cb::SUMSQ_MAX_BLOCKS * (int64_t)sizeof(float); }

int cb_sumsq(const float* x, int64_t n, float* out, void* workspace, int64_t workspace_bytes, void* stream) {
    using namespace cb;
    CB_REQUIRE(n >= 0 && out != nullptr, CB_E_INVALID, "cb_sumsq: bad argument");
    CB_REQUIRE(n == 0 || x != nullptr, CB_E_INVALID, "cb_sumsq: x is NULL");
    CB_REQUIRE(workspace != nullptr && workspace_bytes >= cb_sumsq_workspace_bytes(), CB_E_WORKSPACE,
               "cb_sumsq: workspace smaller than cb_sumsq_workspace_bytes()");
    // the grid depends on n only, so the association (and the result bits) do too
    int64_t blocks = ceil_div(n > 0 ? n : 1, 256 * 16);
    if (blocks > SUMSQ_MAX_BLOCKS) blocks = SUMSQ_MAX_BLOCKS;
    k_sumsq_partial<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, (float*)workspace);
    CB_LAUNCH_CHECK();
    k_sumsq_final<<<1, 256, 0, (cudaStream_t)stream>>>((const float*)workspace, (int)blocks, out);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int64_t cb_prep_workspace_bytes(int64_t rows, int64_t d) {
    if (rows < 0 || d <= 0) return 0;
    return (int64_t)cb::prep_blocks(rows) * d * (int64_t)sizeof(float);
}

int cb_agg_backward_prep(const cb_graph_t* g, const float* d_out, const float* d_out_scaled, int64_t d,
                         const uint8_t* mask, const float* relu_out, int act, int mixed, double alpha,
                         float* G, float* d_bias, float* d_x0, int accumulate_x0, void* workspace,
                         int64_t workspace_bytes, void* stream) {
    using namespace cb;
    CB_REQUIRE(g != nullptr, CB_E_INVALID, "cb_agg_backward_prep: graph is NULL");
    CB_REQUIRE(d > 0, CB_E_INVALID, "cb_agg_backward_prep: d must be positive");
    CB_REQUIRE(d_out != nullptr || d_out_scaled != nullptr, CB_E_INVALID, "cb_agg_backward_prep: no incoming gradient");
    CB_REQUIRE(G != nullptr, CB_E_INVALID, "cb_agg_backward_prep: G is NULL");
    CB_REQUIRE(act == CB_ACT_NONE || mask != nullptr || relu_out != nullptr, CB_E_INVALID,
               "cb_agg_backward_prep: relu needs mask or relu_out");
    const int64_t rows = g->rows;
    const int blocks = prep_blocks(rows);
    if (d_bias) {
        CB_REQUIRE(workspace != nullptr && workspace_bytes >= cb_prep_workspace_bytes(rows, d), CB_E_WORKSPACE,
                   "cb_agg_backward_prep: workspace smaller than cb_prep_workspace_bytes()");
    }
    PrepArgs a{};
    a.d_out = d_out;
    a.d_out2 = d_out_scaled;
    a.s2 = g->dout_is;
    a.rs = g->din_is;
    a.mask = mask;
    a.relu_out = relu_out;
    a.act = act;
    a.mixed = mixed;
    a.alpha = (float)alpha;
    a.one_minus_alpha = (float)(1.0 - alpha);
    a.G = G;
    a.d_x0 = d_x0;
    a.accumulate_x0 = accumulate_x0;
    a.bias_partial = d_bias ? (float*)workspace : nullptr;
    a.rows = rows;
    a.d = d;
    a.rows_per_block = ceil_div(rows > 0 ? rows : 1, blocks);
    auto al = [](const void* p, uintptr_t m) { return (reinterpret_cast<uintptr_t>(p) & m) == 0; };
    const bool vec = d % 4 == 0 && al(d_out, 15) && al(d_out_scaled, 15) && al(G, 15) && al(d_x0, 15) &&
                     al(relu_out, 15) && al(mask, 3);
    const int64_t units = vec ? d / 4 : d;
    const int cu = (int)(units < 256 ? units : 256);
    const int rl = 256 / cu;
    const size_t smem = d_bias ? (size_t)rl * cu * (vec ? 4 : 1) * sizeof(float) : 0;
    if (vec)
        k_prep<4><<<blocks, 256, smem, (cudaStream_t)stream>>>(a);
    else
        k_prep<1><<<blocks, 256, smem, (cudaStream_t)stream>>>(a);
    CB_LAUNCH_CHECK();
    if (d_bias) {
        k_bias_final<<<(unsigned)ceil_div(d, 256), 256, 0, (cudaStream_t)stream>>>((const float*)workspace, blocks, d, d_bias);
        CB_LAUNCH_CHECK();
    }
    return CB_OK;
}

}  // extern "C"
