// Streaming (one pass, HBM-bound) kernels either side of the aggregation:
//   cb_row_scale          GCN.py:205-213   feat_src * out_deg^-1/2          (and its adjoint)
//   cb_sumsq              GCN.py:232       th.norm(self.le)                 (sum of squares; sqrt on the caller side)
//   cb_agg_backward_prep  autograd of GCN.py:242-253, 127-128 and res_tricks.py:14/23
// plus the library-wide bookkeeping (error text, launch counter).
#include <cuda_bf16.h>
#include <math.h>

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
// element access in the storage type S of the streamed matrices (fp32, or bf16 widened on load / rounded on store)
template <typename S, int VEC>
struct PIo;
template <>
struct PIo<float, 4> {
    static __device__ __forceinline__ void ld(float (&v)[4], const float* p) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void st(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <>
struct PIo<float, 1> {
    static __device__ __forceinline__ void ld(float (&v)[1], const float* p) { v[0] = __ldg(p); }
    static __device__ __forceinline__ void st(float* p, const float (&v)[1]) { *p = v[0]; }
};
template <>
struct PIo<__nv_bfloat16, 4> {
    static __device__ __forceinline__ void ld(float (&v)[4], const __nv_bfloat16* p) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
        v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
        v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&v)[4]) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    }
};
template <>
struct PIo<__nv_bfloat16, 1> {
    static __device__ __forceinline__ void ld(float (&v)[1], const __nv_bfloat16* p) { v[0] = __bfloat162float(*p); }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&v)[1]) { *p = __float2bfloat16_rn(v[0]); }
};

struct PrepArgs {
    const void* d_out;     // storage type S
    const void* d_out2;    // S
    const float* s2;       // dout^-1/2 (scale of d_out2)
    const float* rs;       // din^-1/2
    const uint8_t* mask;
    const void* relu_out;  // S
    int act, mixed;
    float alpha, one_minus_alpha;
    void* G;               // S
    void* d_x0;            // S
    int accumulate_x0;
    float* bias_partial;   // [gridDim.x, d] or null
    int64_t rows, d;
    int64_t rows_per_block;
    // cb_agg_backward_prep_ex: d_out is the gradient of dropout(out): dt = keep ? drop_scale * d_out : 0 first
    // (native_dropout_backward's grad * mask * scale), and the rows of G holding a non-zero are flagged
    const uint8_t* drop_keep;  // [rows, d] bytes (a bool tensor), or null
    float drop_scale;
    uint8_t* row_live;         // [rows] zeroed by the caller, or null
};

// Block layout: CU = min(units, 256) column slots side by side, RL = 256 / CU row lanes.  Every thread
// keeps a running bias-gradient sum for its column slot over the rows it visits; the row lanes of a
// block are then added in lane order, giving one partial row per block.
template <typename S, int VEC>
__global__ void __launch_bounds__(256) k_prep(const PrepArgs a) {
    using IO = PIo<S, VEC>;
    const S* p_out = reinterpret_cast<const S*>(a.d_out);
    const S* p_out2 = reinterpret_cast<const S*>(a.d_out2);
    const S* p_relu = reinterpret_cast<const S*>(a.relu_out);
    S* p_G = reinterpret_cast<S*>(a.G);
    S* p_x0 = reinterpret_cast<S*>(a.d_x0);
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
                if (a.d_out) IO::ld(dt, p_out + off);
                if (a.drop_keep) {
                    bool k[VEC];
                    if (VEC == 4) {
                        const uchar4 kk = *reinterpret_cast<const uchar4*>(a.drop_keep + off);
                        k[0] = kk.x; k[1 % VEC] = kk.y; k[2 % VEC] = kk.z; k[3 % VEC] = kk.w;
                    } else {
                        k[0] = a.drop_keep[off] != 0;
                    }
#pragma unroll
                    for (int i = 0; i < VEC; ++i) dt[i] = k[i] ? __fmul_rn(dt[i], a.drop_scale) : 0.f;
                }
                if (a.d_out2) {
                    const float s2 = __ldg(a.s2 + r);
                    float t[VEC];
                    IO::ld(t, p_out2 + off);
#pragma unroll
                    for (int i = 0; i < VEC; ++i)
                        dt[i] = a.d_out ? __fadd_rn(dt[i], __fmul_rn(s2, t[i])) : __fmul_rn(s2, t[i]);
                }
                if (a.d_x0) {
                    float old_[VEC], g0[VEC];
                    if (a.accumulate_x0) IO::ld(old_, p_x0 + off);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        g0[i] = __fmul_rn(a.alpha, dt[i]);
                        if (a.accumulate_x0) g0[i] = __fadd_rn(old_[i], g0[i]);
                    }
                    IO::st(p_x0 + off, g0);
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
                        float ro[VEC];
                        IO::ld(ro, p_relu + off);
#pragma unroll
                        for (int i = 0; i < VEC; ++i) m[i] = ro[i] > 0.f;
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
                IO::st(p_G + off, g);
                if (a.row_live) {
                    bool nz = false;
#pragma unroll
                    for (int i = 0; i < VEC; ++i) nz |= g[i] != 0.f;
                    if (nz) a.row_live[r] = 1;       // every writer stores the same byte
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

// ---------------------------------------------------------------------------------------------
// Structural-Embedding optimizer step: Adam + weight decay + the ||E||_F regulariser's gradient, one pass
// ---------------------------------------------------------------------------------------------
struct SeAdamArgs {
    float* E;              // [n] fp32 master of GCNConv.le
    const void* grad;      // [n] dL/dh of the layer's transform (= dL/dE of the additive term), fp32 or bf16, or null
    float* m;
    float* v;
    __nv_bfloat16* shadow; // [n] bf16 copy the forward pass reads, or null
    int64_t n;
    float one_minus_b1, b2, one_minus_b2, eps, wd, step_size, inv_bc2_sqrt;
    const float* sumsq;    // device scalar sum(E^2) over every rank's rows (this step's forward), or null
    float reg_coef;        // se_reg coefficient of the loss
};

template <typename GS>
__global__ void __launch_bounds__(256) k_se_adam(const SeAdamArgs a) {
    // d/dE [ coef * ||E||_F ] = coef * E / ||E||_F  (0 at E = 0, like torch.norm's subgradient)
    float reg = 0.f;
    if (a.sumsq) {
        const float ss = __ldg(a.sumsq);
        reg = ss > 0.f ? a.reg_coef / sqrtf(ss) : 0.f;
    }
    const int64_t n4 = a.n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const GS* gp = reinterpret_cast<const GS*>(a.grad);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float e[4], g[4], m[4], v[4];
        PIo<float, 4>::ld(e, a.E + 4 * i);
        PIo<float, 4>::ld(m, a.m + 4 * i);
        PIo<float, 4>::ld(v, a.v + 4 * i);
        if (gp) PIo<GS, 4>::ld(g, gp + 4 * i);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float gr = gp ? g[k] : 0.f;
            gr = fmaf(reg, e[k], gr);
            gr = fmaf(a.wd, e[k], gr);                                   // grad.add(param, alpha=weight_decay)
            m[k] = fmaf(a.one_minus_b1, gr - m[k], m[k]);                // exp_avg.lerp_(grad, 1 - beta1)
            v[k] = fmaf(a.one_minus_b2 * gr, gr, a.b2 * v[k]);           // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
            const float denom = sqrtf(v[k]) * a.inv_bc2_sqrt + a.eps;
            e[k] = e[k] - a.step_size * (m[k] / denom);                  // param.addcdiv_(exp_avg, denom, -step_size)
        }
        PIo<float, 4>::st(a.E + 4 * i, e);
        PIo<float, 4>::st(a.m + 4 * i, m);
        PIo<float, 4>::st(a.v + 4 * i, v);
        if (a.shadow) PIo<__nv_bfloat16, 4>::st(a.shadow + 4 * i, e);
    }
    // tail (n not a multiple of 4)
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        float gr = 0.f;
        if (gp) {
            float t[1];
            PIo<GS, 1>::ld(t, gp + i);
            gr = t[0];
        }
        const float e = a.E[i];
        gr = fmaf(reg, e, gr);
        gr = fmaf(a.wd, e, gr);
        const float m = fmaf(a.one_minus_b1, gr - a.m[i], a.m[i]);
        const float v = fmaf(a.one_minus_b2 * gr, gr, a.b2 * a.v[i]);
        const float en = e - a.step_size * (m / (sqrtf(v) * a.inv_bc2_sqrt + a.eps));
        a.E[i] = en; a.m[i] = m; a.v[i] = v;
        if (a.shadow) a.shadow[i] = __float2bfloat16_rn(en);
    }
}

// fp32 -> bf16 copy (initialises the shadow table)
__global__ void __launch_bounds__(256) k_to_bf16(const float* __restrict__ x, int64_t n, __nv_bfloat16* __restrict__ y) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float t[4];
        PIo<float, 4>::ld(t, x + 4 * i);
        PIo<__nv_bfloat16, 4>::st(y + 4 * i, t);
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        y[i] = __float2bfloat16_rn(x[i]);
}

// flags[r] = any(x[r, :] != 0): LPR lanes per row (so that narrow rows fill the warp), 16 bytes per lane per step
// (-0.0 counts as zero)
template <int ES, int LPR>
__global__ void __launch_bounds__(256) k_row_any_nonzero(const void* __restrict__ x, int64_t rows, int64_t d, int64_t ld,
                                                         uint8_t* __restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR;
    const int64_t row = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (32 / LPR) + lane / LPR;
    uint32_t any = 0;
    if (row < rows) {
        const char* p = reinterpret_cast<const char*>(x) + row * ld * ES;
        const int64_t bytes = d * ES;
        const bool vec = ((reinterpret_cast<uintptr_t>(p) & 15u) == 0);
        const int64_t n16 = vec ? bytes / 16 : 0;
        for (int64_t i = sub; i < n16; i += LPR) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + i);
            if (ES == 4) any |= (v.x | v.y | v.z | v.w) << 1;                      // drop the sign bits
            else any |= (v.x | v.y | v.z | v.w) & 0x7fff7fffu;
        }
        for (int64_t b = n16 * 16 + sub * ES; b < bytes; b += LPR * ES) {
            if (ES == 4) any |= *reinterpret_cast<const uint32_t*>(p + b) << 1;
            else any |= (uint32_t)(*reinterpret_cast<const uint16_t*>(p + b) & 0x7fffu);
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, any != 0u);
    const unsigned grp = LPR == 32 ? 0xffffffffu : (((1u << LPR) - 1u) << ((lane / LPR) * LPR));
    if (sub == 0 && row < rows) flags[row] = (m & grp) != 0u;
}

template <int ES>
static void launch_row_any(const void* x, int64_t rows, int64_t d, int64_t ld, uint8_t* flags, cudaStream_t st) {
    const int64_t n16 = (d * ES + 15) / 16;      // 16-byte pieces per row
    if (n16 <= 4)
        k_row_any_nonzero<ES, 4><<<(unsigned)ceil_div(rows, 8 * 8), 256, 0, st>>>(x, rows, d, ld, flags);
    else if (n16 <= 8)
        k_row_any_nonzero<ES, 8><<<(unsigned)ceil_div(rows, 8 * 4), 256, 0, st>>>(x, rows, d, ld, flags);
    else if (n16 <= 16)
        k_row_any_nonzero<ES, 16><<<(unsigned)ceil_div(rows, 8 * 2), 256, 0, st>>>(x, rows, d, ld, flags);
    else
        k_row_any_nonzero<ES, 32><<<(unsigned)ceil_div(rows, 8), 256, 0, st>>>(x, rows, d, ld, flags);
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

int cb_se_adam_step(float* E, const void* grad, int grad_dtype, float* m, float* v, uint16_t* shadow_bf16, int64_t n,
                    double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step,
                    const float* sumsq, double reg_coef, void* stream) {
    using namespace cb;
    CB_REQUIRE(n >= 0 && step >= 1, CB_E_INVALID, "cb_se_adam_step: bad size or step");
    if (n == 0) return CB_OK;
    CB_REQUIRE(E && m && v, CB_E_INVALID, "cb_se_adam_step: NULL state");
    CB_REQUIRE(grad_dtype == CB_F32 || grad_dtype == CB_BF16, CB_E_INVALID, "cb_se_adam_step: unknown gradient dtype");
    auto al = [](const void* p, uintptr_t msk) { return (reinterpret_cast<uintptr_t>(p) & msk) == 0; };
    CB_REQUIRE(al(E, 15) && al(m, 15) && al(v, 15) && al(shadow_bf16, 7) && al(grad, grad_dtype == CB_BF16 ? 7 : 15),
               CB_E_UNSUPPORTED, "cb_se_adam_step: buffers must be aligned to 4 elements");
    SeAdamArgs a{};
    a.E = E; a.grad = grad; a.m = m; a.v = v; a.shadow = (__nv_bfloat16*)shadow_bf16; a.n = n;
    a.one_minus_b1 = (float)(1.0 - beta1);
    a.b2 = (float)beta2;
    a.one_minus_b2 = (float)(1.0 - beta2);
    a.eps = (float)eps;
    a.wd = (float)weight_decay;
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    a.step_size = (float)(lr / bc1);
    a.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    a.sumsq = sumsq;
    a.reg_coef = (float)reg_coef;
    const int64_t want = ceil_div(ceil_div(n, 4), 256);
    const int64_t cap = (int64_t)sm_count() * 16;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (grad_dtype == CB_BF16)
        k_se_adam<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    else
        k_se_adam<float><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_row_any_nonzero(const void* x, int dtype, int64_t rows, int64_t d, int64_t ld, uint8_t* flags, void* stream) {
    using namespace cb;
    CB_REQUIRE(rows >= 0 && d > 0 && ld >= d, CB_E_INVALID, "cb_row_any_nonzero: bad shape");
    CB_REQUIRE(dtype == CB_F32 || dtype == CB_BF16, CB_E_INVALID, "cb_row_any_nonzero: unknown dtype");
    if (rows == 0) return CB_OK;
    CB_REQUIRE(x && flags, CB_E_INVALID, "cb_row_any_nonzero: NULL buffer");
    if (dtype == CB_BF16) launch_row_any<2>(x, rows, d, ld, flags, (cudaStream_t)stream);
    else launch_row_any<4>(x, rows, d, ld, flags, (cudaStream_t)stream);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_to_bf16(const float* x, int64_t n, uint16_t* y, void* stream) {
    using namespace cb;
    CB_REQUIRE(n >= 0, CB_E_INVALID, "cb_to_bf16: negative size");
    if (n == 0) return CB_OK;
    CB_REQUIRE(x && y, CB_E_INVALID, "cb_to_bf16: NULL buffer");
    CB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15u) == 0 && (reinterpret_cast<uintptr_t>(y) & 7u) == 0,
               CB_E_UNSUPPORTED, "cb_to_bf16: buffers must be aligned to 4 elements");
    const int64_t want = ceil_div(ceil_div(n, 4), 256);
    const int64_t cap = (int64_t)sm_count() * 16;
    k_to_bf16<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(x, n, (__nv_bfloat16*)y);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int64_t cb_prep_workspace_bytes(int64_t rows, int64_t d) {
    if (rows < 0 || d <= 0) return 0;
    return (int64_t)cb::prep_blocks(rows) * d * (int64_t)sizeof(float);
}

static int backward_prep_impl(const cb_graph_t* g, int dtype, const void* d_out, const void* d_out_scaled, int64_t d,
                              const uint8_t* mask, const void* relu_out, int act, int mixed, double alpha,
                              void* G, float* d_bias, void* d_x0, int accumulate_x0, void* workspace,
                              int64_t workspace_bytes, void* stream, const uint8_t* drop_keep = nullptr,
                              double drop_scale = 1.0, uint8_t* row_live = nullptr) {
    using namespace cb;
    CB_REQUIRE(g != nullptr, CB_E_INVALID, "cb_agg_backward_prep: graph is NULL");
    CB_REQUIRE(d > 0, CB_E_INVALID, "cb_agg_backward_prep: d must be positive");
    if (g->rows == 0) {                   // a slice that owns no rows: the bias gradient is zero, nothing else exists
        if (d_bias) CB_CUDA(cudaMemsetAsync(d_bias, 0, (size_t)d * sizeof(float), (cudaStream_t)stream));
        return CB_OK;
    }
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
    CB_REQUIRE(drop_keep == nullptr || (d_out != nullptr && d_out_scaled == nullptr), CB_E_INVALID,
               "cb_agg_backward_prep_ex: the dropout mask applies to the plain output's gradient only");
    a.drop_keep = drop_keep;
    a.drop_scale = (float)drop_scale;
    a.row_live = row_live;
    a.rows = rows;
    a.d = d;
    a.rows_per_block = ceil_div(rows > 0 ? rows : 1, blocks);
    auto al = [](const void* p, uintptr_t m) { return (reinterpret_cast<uintptr_t>(p) & m) == 0; };
    const uintptr_t am = dtype == CB_BF16 ? 7 : 15;    // 4 elements per access
    const bool vec = d % 4 == 0 && al(d_out, am) && al(d_out_scaled, am) && al(G, am) && al(d_x0, am) &&
                     al(relu_out, am) && al(mask, 3) && al(drop_keep, 3);
    const int64_t units = vec ? d / 4 : d;
    const int cu = (int)(units < 256 ? units : 256);
    const int rl = 256 / cu;
    const size_t smem = d_bias ? (size_t)rl * cu * (vec ? 4 : 1) * sizeof(float) : 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CB_BF16) {
        if (vec) k_prep<__nv_bfloat16, 4><<<blocks, 256, smem, st>>>(a);
        else k_prep<__nv_bfloat16, 1><<<blocks, 256, smem, st>>>(a);
    } else {
        if (vec) k_prep<float, 4><<<blocks, 256, smem, st>>>(a);
        else k_prep<float, 1><<<blocks, 256, smem, st>>>(a);
    }
    CB_LAUNCH_CHECK();
    if (d_bias) {
        k_bias_final<<<(unsigned)ceil_div(d, 256), 256, 0, (cudaStream_t)stream>>>((const float*)workspace, blocks, d, d_bias);
        CB_LAUNCH_CHECK();
    }
    return CB_OK;
}

int cb_agg_backward_prep(const cb_graph_t* g, const float* d_out, const float* d_out_scaled, int64_t d,
                         const uint8_t* mask, const float* relu_out, int act, int mixed, double alpha,
                         float* G, float* d_bias, float* d_x0, int accumulate_x0, void* workspace,
                         int64_t workspace_bytes, void* stream) {
    return backward_prep_impl(g, CB_F32, d_out, d_out_scaled, d, mask, relu_out, act, mixed, alpha, G, d_bias, d_x0,
                              accumulate_x0, workspace, workspace_bytes, stream);
}

int cb_agg_backward_prep_ex(const cb_graph_t* g, int dtype, const void* d_out, const void* d_out_scaled, int64_t d,
                            const uint8_t* mask, const void* relu_out, int act, int mixed, double alpha,
                            const uint8_t* drop_keep, double drop_scale, void* G, float* d_bias, void* d_x0,
                            int accumulate_x0, uint8_t* row_live, void* workspace, int64_t workspace_bytes, void* stream) {
    CB_REQUIRE(dtype == CB_F32 || dtype == CB_BF16, CB_E_INVALID, "cb_agg_backward_prep_ex: unknown dtype");
    return backward_prep_impl(g, dtype, d_out, d_out_scaled, d, mask, relu_out, act, mixed, alpha, G, d_bias, d_x0,
                              accumulate_x0, workspace, workspace_bytes, stream, drop_keep, drop_scale, row_live);
}

int cb_agg_backward_prep_bf16(const cb_graph_t* g, const uint16_t* d_out, const uint16_t* d_out_scaled, int64_t d,
                              const uint8_t* mask, const uint16_t* relu_out, int act, int mixed, double alpha,
                              uint16_t* G, float* d_bias, uint16_t* d_x0, int accumulate_x0, void* workspace,
                              int64_t workspace_bytes, void* stream) {
    return backward_prep_impl(g, CB_BF16, d_out, d_out_scaled, d, mask, relu_out, act, mixed, alpha, G, d_bias, d_x0,
                              accumulate_x0, workspace, workspace_bytes, stream);
}

}  // extern "C"
