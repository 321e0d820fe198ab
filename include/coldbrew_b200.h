/*
 * coldbrew_b200 -- C ABI of the B200-native TeacherGNN aggregation path.
 *
 * This is the "lower seam" of Cold Brew's GCN layer: everything the reference reaches through DGL's C
 * API from GNN_model/GCN.py is replaced by the entry points below (reference file:line cited per
 * function).  Plain C: raw device pointers, sizes and a cudaStream_t passed as void*.  No torch types.
 *
 * Conventions
 *   - every function returns 0 on success or a negative CB_E_* code; cb_last_error() gives the text
 *     of the last failure on the calling thread;
 *   - the caller owns every tensor buffer; the library owns only what cb_graph_create allocates;
 *   - no call synchronises the device except cb_graph_create / cb_graph_create_sliced (one-time build);
 *   - all launches go to the stream passed in; no internal threads; no hidden allocations after
 *     graph creation (per-call scratch is the caller-supplied `workspace`, sized by
 *     cb_graph_workspace_bytes);
 *   - feature matrices are row-major, contiguous, fp32; row r of an [n, d] matrix starts at r*d.
 *   - node ids inside a graph handle are int32 (N < 2^31), edge offsets int64.
 */
#ifndef COLDBREW_B200_H
#define COLDBREW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_ABI_VERSION 6   /* 3: live-column compaction, bf16 transform, SE optimizer step, local-edge graph build; 4: a_live / x0_valid; 5: source panels; 6: graph preparation, cb_agg_backward_prep_ex, edge weights, cb_gemm_rows_masked */

enum {
    CB_OK = 0,
    CB_E_INVALID = -1,     /* bad argument (null pointer, negative size, d <= 0 ...) */
    CB_E_RANGE = -2,       /* a node id in edge_index is outside [0, N) */
    CB_E_CUDA = -3,        /* a CUDA runtime call failed; see cb_last_error() */
    CB_E_UNSUPPORTED = -4, /* shape outside what the kernels cover (N or E >= 2^31) */
    CB_E_WORKSPACE = -5    /* workspace missing or too small */
};

/* activation selector for the fused epilogue (GCN.py:127-128 applies F.relu after the layer) */
enum { CB_ACT_NONE = 0, CB_ACT_RELU = 1 };

/* storage type of the feature matrices an aggregation reads / writes (sums and epilogue are always fp32) */
enum { CB_F32 = 0, CB_BF16 = 1 };

/* which side of the graph an aggregation walks */
enum {
    CB_BY_DST = 0, /* rows = destination nodes, gathers over in-edges   (forward,  GCN.py:238) */
    CB_BY_SRC = 1  /* rows = source nodes,      gathers over out-edges  (backward, autograd of :238) */
};

/* cb_graph_query selectors.  Integers are written as int64_t, pointers as const void* (device). */
enum {
    CB_Q_NUM_NODES = 0,      /* int64: N (global) */
    CB_Q_NUM_EDGES = 1,      /* int64: edges stored in the BY_DST structure */
    CB_Q_ROW_BEGIN = 2,      /* int64: first owned node */
    CB_Q_ROW_END = 3,        /* int64: one past the last owned node */
    CB_Q_HAS_ZERO_IN_DEG = 4,/* int64: 1 if an owned node has no in-edge (GCN.py:187-188) */
    CB_Q_HUB_CHUNK = 5,      /* int64: rows longer than this are summed chunk-wise */
    CB_Q_SRC_PANELS = 6,     /* int64: source panels the rows are grouped by (1 = plain order) */
    CB_Q_DST_ROWPTR_EXP = 14,/* const int64_t* [rows*panels+1] or NULL: offsets of every (row, panel) group */
    CB_Q_SRC_ROWPTR_EXP = 25,
    CB_Q_DST_ROWPTR = 10,    /* const int64_t* [rows+1] */
    CB_Q_DST_COL = 11,       /* const int32_t* [E_dst]  source id of every stored in-edge */
    CB_Q_DST_PERM = 12,      /* const int32_t* [E_dst]  position in the caller's edge list */
    CB_Q_DST_NUM_HUB_CHUNKS = 13, /* int64 */
    CB_Q_SRC_ROWPTR = 20,    /* const int64_t* [rows+1] */
    CB_Q_SRC_COL = 21,       /* const int32_t* [E_src]  destination id of every stored out-edge */
    CB_Q_SRC_PERM = 22,      /* const int32_t* [E_src] */
    CB_Q_SRC_NUM_HUB_CHUNKS = 23, /* int64 */
    CB_Q_SRC_NUM_EDGES = 24, /* int64: edges stored in the BY_SRC structure */
    CB_Q_DIN_INV_SQRT = 30,  /* const float* [rows]  clamp(in_degree,1)^-1/2   (GCN.py:242-246) */
    CB_Q_DOUT_INV_SQRT = 31, /* const float* [rows]  clamp(out_degree,1)^-1/2  (GCN.py:205-209) */
    CB_Q_IN_DEGREE = 32,     /* const int32_t* [rows] */
    CB_Q_OUT_DEGREE = 33     /* const int32_t* [rows] */
};

typedef struct cb_graph cb_graph_t;

/* text of the last error raised on this thread ("" if none) */
const char* cb_last_error(void);
int cb_abi_version(void);

/*
 * Build the device-side graph structure from a COO edge list.
 * Replaces GCN.py:92-94 (edge_index -> .tolist() -> dgl.graph(...).to(device)), the per-call degree
 * computations of GCN.py:205-209,242-246 and the per-call zero-in-degree scan of GCN.py:187-188:
 * all three are properties of the static graph and are computed once here.
 *
 *   edge_index  device, int64, [2, E] row-major: row 0 = source ids, row 1 = destination ids.
 *               Multigraph semantics: duplicate edges are kept and count twice.
 *   hub_chunk   rows with more than hub_chunk stored edges are summed as consecutive chunks of
 *               hub_chunk entries (in order) whose partials are then added in order; <= 0 picks the
 *               default (CB_DEFAULT_HUB_CHUNK).
 * Result: CSR by destination and CSR by source.  Inside a row the stored order is the order of the
 * caller's edge list (stable), so an in-order row sum equals a sequential pass over the COO list.
 * Synchronises `stream` before returning.
 */
#define CB_DEFAULT_HUB_CHUNK 256
int cb_graph_create(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes, int hub_chunk,
                    void* stream, cb_graph_t** out);

/*
 * Same, for one slice of a 1-D node partition (no counterpart in the reference, which is single
 * device): keeps the in-edges of destination nodes in [row_begin, row_end) (BY_DST structure) and the
 * out-edges of source nodes in the same range (BY_SRC structure).  Column ids stay global; row
 * indices, degrees and scale vectors are local (node v is row v - row_begin).
 */
int cb_graph_create_sliced(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes,
                           int64_t row_begin, int64_t row_end, int hub_chunk, void* stream,
                           cb_graph_t** out);

/*
 * Same slice, built from the rank's OWN edges only (BASELINE.json configs[4]: 10^9 edges never exist as one list):
 *   in_edges   [2, E_in]  every edge whose destination is in [row_begin, row_end)  -> BY_DST structure
 *   out_edges  [2, E_out] every edge whose source is in [row_begin, row_end)       -> BY_SRC structure
 * (row 0 = sources, row 1 = destinations, global ids).  For a symmetric graph out_edges is in_edges with its two
 * rows swapped.  Stored order inside a row = order in the respective list; the perm arrays index that list.  An
 * edge outside the owned range is an error (CB_E_RANGE).  Scratch is ~16 bytes per edge of one list at a time.
 */
int cb_graph_create_local(const int64_t* in_edges, int64_t num_in_edges, const int64_t* out_edges,
                          int64_t num_out_edges, int64_t num_nodes, int64_t row_begin, int64_t row_end, int hub_chunk,
                          void* stream, cb_graph_t** out);

/*
 * Same again, with the stored neighbours of every row grouped by SOURCE PANEL: panel(c) = (c >> CB_PANEL_SHIFT) %
 * src_panels of the column id c (128-row blocks dealt round-robin; inside a group the order of the list is kept).
 * An aggregation can then run as src_panels passes (cb_agg_forward_pass / cb_agg_gather_pass): pass p touches only
 * source rows of panel p and continues the in-order partial sums of pass p-1, so on a node-sliced graph the exchange
 * of panel p+1 (the producing GEMM launched on the row tiles of that panel, cb_peer_push_t.tile_first / tile_step)
 * overlaps the aggregation of panel p at full row width.  The grouping is a property of the graph, not of the
 * slicing: every world size, 1 included, sums in the same order.  src_panels in {1, 2, 4}; 1 = cb_graph_create_local.
 * filter != 0: the lists may hold edges of other slices (e.g. the whole edge list twice), which are dropped.
 */
#define CB_PANEL_SHIFT 7
int cb_graph_create_panelled(const int64_t* in_edges, int64_t num_in_edges, const int64_t* out_edges,
                             int64_t num_out_edges, int64_t num_nodes, int64_t row_begin, int64_t row_end,
                             int hub_chunk, int src_panels, int filter, void* stream, cb_graph_t** out);

int cb_graph_destroy(cb_graph_t* g);
int cb_graph_query(const cb_graph_t* g, int what, void* out);

/* bytes of scratch cb_agg_forward / cb_agg_backward need for feature width d (may be 0) */
int64_t cb_graph_workspace_bytes(const cb_graph_t* g, int side, int64_t d);

/*
 * Fused forward aggregation (replaces, for one layer, GCN.py:238 update_all(copy_src,sum),
 * GCN.py:242-250 in-degree scale, GCN.py:252-253 bias, GCN.py:127-128 relu and
 * res_tricks.py:14/23 residual mix, plus GCN.py:205-213 out-degree scale of the NEXT layer's input):
 *
 *   z[v,:]    = din^-1/2[v] * sum_{(u->v)} H[u,:] + bias            (rows v owned by the handle)
 *   r         = act == RELU ? max(z,0) : z
 *   out[v,:]  = x0 ? (1-alpha) * r + alpha * x0[v,:] : r
 *   out_scaled[v,:] = dout^-1/2[v] * out[v,:]                        (if out_scaled != NULL)
 *   mask[v,c] = z[v,c] > 0                                            (if mask != NULL; 1 byte each)
 *
 *   H          [N_global, d]: every source row (after the halo exchange on a sliced graph), row pitch ld_h
 *   bias       [d] or NULL;  x0, out, out_scaled, mask: [rows, d] local, row pitch ld_out (elements);  any of
 *              out/out_scaled may be NULL but not both.  A pitch of 0 means d.  Pitches wider than d let the
 *              caller aggregate one column panel of wider matrices (pointers pre-offset to the panel).
 */
int cb_agg_forward(const cb_graph_t* g, const float* H, int64_t ld_h, int64_t d, const float* bias,
                   const float* x0, double alpha, int act, float* out, float* out_scaled,
                   uint8_t* mask, int64_t ld_out, void* workspace, int64_t workspace_bytes, void* stream);

/*
 * bf16 storage (BASELINE.json configs[4]: 128-dim bf16 features): H, x0, out, out_scaled hold bf16 bit patterns; every
 * row is widened to fp32 when gathered, the in-order fp32 sum and the whole epilogue are those of cb_agg_forward, and
 * the results are rounded to nearest-even on store.  Half the gather traffic of the fp32 path.  Pitches in elements.
 */
int cb_agg_forward_bf16(const cb_graph_t* g, const uint16_t* H, int64_t ld_h, int64_t d, const float* bias,
                        const uint16_t* x0, double alpha, int act, uint16_t* out, uint16_t* out_scaled, uint8_t* mask,
                        int64_t ld_out, void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Plain gather-reduce over one side of the graph, optional per-row scale of the result:
 *   out[r,:] = (row_scale ? row_scale[r] : 1) * sum_{j in row r} X[col[j],:]
 * side = CB_BY_SRC is the autograd transpose of GCN.py:238 (dH[u] = sum_{(u->v)} G[v]).
 * row_live [N_global] bytes or NULL: X rows with row_live[s] == 0 are known to be all-zero and are skipped (the
 * gradient of a loss over the train rows only is row-sparse after the output head; x + 0 = x, so the sums are
 * unchanged).  cb_gemm_rows_grad can produce the flags of its output.
 */
int cb_agg_gather(const cb_graph_t* g, int side, const float* X, int64_t ld_x, int64_t d, const float* row_scale,
                  const uint8_t* row_live, float* out, int64_t ld_out, void* workspace, int64_t workspace_bytes,
                  void* stream);

int cb_agg_gather_bf16(const cb_graph_t* g, int side, const uint16_t* X, int64_t ld_x, int64_t d,
                       const float* row_scale, const uint8_t* row_live, uint16_t* out, int64_t ld_out, void* workspace,
                       int64_t workspace_bytes, void* stream);

/*
 * One source-panel pass of cb_agg_forward / cb_agg_gather on a graph built by cb_graph_create_panelled (dtype CB_F32 /
 * CB_BF16 of H, x0, out, out_scaled / of X, out): pass `panel` adds the neighbours of that panel, in stored order, to the
 * row sums left in `carry` (fp32 [rows, d]; pass 0 starts them) and writes them back; the LAST pass (panel = src_panels -
 * 1) applies the epilogue and stores the outputs, and also produces the hub rows (chunks over their whole lists, which
 * need every panel).  After the last pass the outputs are bit-identical to the one-pass calls.  The caller runs the
 * passes of one aggregation in order on one stream, each after the rows of its panel have arrived.
 */
int cb_agg_forward_pass(const cb_graph_t* g, int dtype, const void* H, int64_t ld_h, int64_t d, const float* bias,
                        const void* x0, double alpha, int act, void* out, void* out_scaled, uint8_t* mask, int64_t ld_out,
                        int panel, float* carry, void* workspace, int64_t workspace_bytes, void* stream);
int cb_agg_gather_pass(const cb_graph_t* g, int side, int dtype, const void* X, int64_t ld_x, int64_t d,
                       const float* row_scale, void* out, int64_t ld_out, int panel, float* carry, void* workspace,
                       int64_t workspace_bytes, void* stream);

/*
 * Edge-weighted aggregation (GCN.py:199-202: graph.edata['_edge_weight'] = edge_weight; fn.u_mul_e('h', '_edge_weight',
 * 'm') then fn.sum): out[r] = row_scale[r] * sum_j X[col[j]] * w[j], every product rounded before the in-order add.
 *   cb_graph_sort_edge_values  the caller's per-edge values [num_values] (positions of the edge list the graph was built
 *                              from) -> the stored order of `side`: out[j] = values[perm[j]], out [E_side]
 *   cb_agg_gather_weighted     the gather with edge_val in stored order (by destination: the forward sum; by source:
 *                              its autograd transpose on the same weights sorted for that side)
 *   cb_agg_edge_dot            dL/dw: out[perm[j]] = <X[col[j], :], Y[row, :]> for every stored edge of `side` (X holds
 *                              every gathered row, Y the owned rows; out in the caller's edge positions, fp32)
 * dtype CB_F32 / CB_BF16 of X, Y and the gathered output; weights and dot products are fp32.
 */
int cb_graph_sort_edge_values(const cb_graph_t* g, int side, const float* values, int64_t num_values, float* out,
                              void* stream);
int cb_agg_gather_weighted(const cb_graph_t* g, int side, int dtype, const void* X, int64_t ld_x, int64_t d,
                           const float* edge_val, const float* row_scale, void* out, int64_t ld_out, void* workspace,
                           int64_t workspace_bytes, void* stream);
int cb_agg_edge_dot(const cb_graph_t* g, int side, int dtype, const void* X, int64_t ld_x, const void* Y, int64_t ld_y,
                    int64_t d, float* out, void* stream);

/*
 * One label-propagation / outcome-correlation iteration in one kernel (Label_propagation_model/
 * outcome_correlation.py:139-145: result = alpha * (adj @ result); result += (1 - alpha) * y; result = post_step(result)
 * with adj = D^-1/2 A D^-1/2 | D^-1 A | A D^-1, outcome_correlation.py:50-54).  The edge values of adj are products of
 * per-node factors, so adj @ x is the unit-weight gather between two row scalings and no edge-value array exists:
 *   v[r,:]    = c_agg * (row_scale ? row_scale[r] : 1) * sum_{j in row r} X[col[j],:]  +  c_y * y[r,:]
 *   v         = clamp ? min(max(v, clamp_lo), clamp_hi) : v
 *   out[r,:]  = v                                   (if out)
 *   out2[r,:] = out2_scale[r] * v                   (if out2: the pre-scaled iterate the next step gathers)
 * X: [N_global, d] fp32, y / out / out2: [rows, d] fp32, contiguous.
 */
int cb_agg_propagate(const cb_graph_t* g, int side, const float* X, int64_t d, const float* row_scale, const float* y,
                     double c_agg, double c_y, int clamp, double clamp_lo, double clamp_hi, float* out,
                     const float* out2_scale, float* out2, void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Row-sparse gather without walking the dead columns.  cb_graph_compact_live builds, in the caller's workspace
 * (cb_graph_live_workspace_bytes, 256-byte aligned), a second CSR of one side that keeps only the columns s with
 * row_live[s] != 0, in the stored order; cb_agg_gather_compacted then gathers over it (dtype CB_F32 / CB_BF16 of X
 * and out).  Hub rows keep the chunk bounds of the full list, so every partial sum has the same association as the
 * full walk and the result is bit-identical to cb_agg_gather with the same flags (and, x + 0 = x, without them).
 * Use: the gradient under a loss over the train rows only (trainer_node_classification.py:390-391) is non-zero on
 * those rows alone; the transposed aggregation (autograd of GCN.py:238) of the last layer is then a gather over
 * ~|train|/N of the edges.  The compaction is 4 small kernels reading the column ids twice.
 */
int64_t cb_graph_live_workspace_bytes(const cb_graph_t* g, int side);
int cb_graph_compact_live(const cb_graph_t* g, int side, const uint8_t* row_live, void* live_ws,
                          int64_t live_ws_bytes, void* stream);
int cb_agg_gather_compacted(const cb_graph_t* g, int side, int dtype, const void* X, int64_t ld_x, int64_t d,
                            const float* row_scale, const void* live_ws, void* out, int64_t ld_out, void* workspace,
                            int64_t workspace_bytes, void* stream);

/*
 * Backward prologue of cb_agg_forward: from the gradient(s) arriving at the layer output build the
 * matrix that the transposed aggregation gathers, the bias gradient and the residual gradient.
 *   dtot = (d_out ? d_out : 0) + (d_out_scaled ? dout^-1/2[v] * d_out_scaled : 0)
 *   dz   = (x0 was mixed ? (1-alpha) : 1) * dtot * (act == RELU ? m : 1),  m = mask ? mask!=0 : relu_out>0
 *   G[v,:]   = din^-1/2[v] * dz                         [rows, d]
 *   d_bias   = sum_v dz[v,:]                            [d]   (if d_bias != NULL; two-stage, fixed order)
 *   d_x0[v,:] (+)= alpha * dtot                         (if d_x0 != NULL; accumulate_x0 selects += vs =)
 * workspace: cb_prep_workspace_bytes(rows, d).
 */
int64_t cb_prep_workspace_bytes(int64_t rows, int64_t d);
int cb_agg_backward_prep(const cb_graph_t* g, const float* d_out, const float* d_out_scaled, int64_t d,
                         const uint8_t* mask, const float* relu_out, int act, int mixed, double alpha,
                         float* G, float* d_bias, float* d_x0, int accumulate_x0, void* workspace,
                         int64_t workspace_bytes, void* stream);

/* bf16 storage of every streamed matrix (d_out, d_out_scaled, relu_out, G, d_x0); same arithmetic in fp32 */
int cb_agg_backward_prep_bf16(const cb_graph_t* g, const uint16_t* d_out, const uint16_t* d_out_scaled, int64_t d,
                              const uint8_t* mask, const uint16_t* relu_out, int act, int mixed, double alpha,
                              uint16_t* G, float* d_bias, uint16_t* d_x0, int accumulate_x0, void* workspace,
                              int64_t workspace_bytes, void* stream);

/*
 * The same prologue when a dropout sat between the aggregation's output and its consumer (GCN.py:104,110,133 with the
 * reference's training defaults, base_options.py:190-220): d_out is the gradient of dropout(out) and
 * drop_keep [rows, d] (bytes of the boolean mask torch.native_dropout returned) / drop_scale = 1 / (1 - p) turn it into
 * the gradient of out first -- keep ? drop_scale * d_out : 0, what native_dropout_backward computes -- in the same pass
 * (d_out_scaled must be NULL then).  row_live [rows] (zeroed by the caller) or NULL: receives 1 for every row of G that
 * holds a non-zero, so that the transposed gather can run over the compacted lists (cb_graph_compact_live) without a
 * separate pass over G.  dtype CB_F32 / CB_BF16 of d_out, d_out_scaled, relu_out, G, d_x0.
 */
int cb_agg_backward_prep_ex(const cb_graph_t* g, int dtype, const void* d_out, const void* d_out_scaled, int64_t d,
                            const uint8_t* mask, const void* relu_out, int act, int mixed, double alpha,
                            const uint8_t* drop_keep, double drop_scale, void* G, float* d_bias, void* d_x0,
                            int accumulate_x0, uint8_t* row_live, void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Optimizer step of one Structural-Embedding table (GCN.py:181-182 `self.le`; trainer_node_classification.py:310
 * Adam(lr, weight_decay), :393-394 loss += se_reg * ||E||_F, :428-430 zero_grad / backward / step), one pass over
 * the table instead of autograd's norm backward + gradient add + torch.optim.Adam's passes:
 *   g  = grad[i] + reg_coef * E[i] / sqrt(*sumsq) + weight_decay * E[i]        (grad = dL/dh of the layer's transform:
 *                                                                               h = (D X) W + E  =>  dL/dE = dL/dh)
 *   m  = m + (1-beta1) (g - m);  v = beta2 v + (1-beta2) g^2                   (torch.optim.Adam, amsgrad off)
 *   E -= lr / (1-beta1^step) * m / (sqrt(v) / sqrt(1-beta2^step) + eps)
 *   shadow_bf16[i] = bf16(E[i])                                                (the copy a bf16 forward reads)
 * grad: fp32 (CB_F32) or bf16 (CB_BF16) or NULL; sumsq: device scalar holding sum(E^2) over EVERY rank's rows as
 * computed by this step's forward (cb_sumsq + all-reduce), or NULL for no regulariser; step counts from 1.
 */
int cb_se_adam_step(float* E, const void* grad, int grad_dtype, float* m, float* v, uint16_t* shadow_bf16, int64_t n,
                    double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step,
                    const float* sumsq, double reg_coef, void* stream);
/* y = bf16(x), round to nearest even (initialises a shadow table; casts fp32 inputs of a bf16 forward) */
int cb_to_bf16(const float* x, int64_t n, uint16_t* y, void* stream);

/* y[r,:] = s[r] * x[r,:]   (GCN.py:205-213 `feat_src * norm`; also its adjoint) */
int cb_row_scale(const float* x, const float* s, int64_t rows, int64_t d, float* y, void* stream);

/*
 * sum of squares of n floats -> out[0] (device), fixed-order two-stage reduction (GCN.py:232
 * th.norm(self.le) = sqrt of this).  workspace: cb_sumsq_workspace_bytes().
 */
int64_t cb_sumsq_workspace_bytes(void);
int cb_sumsq(const float* x, int64_t n, float* out, void* workspace, int64_t workspace_bytes,
             void* stream);

/*
 * Multi-GPU exchange (1-D node partition, one process per GPU on one node; no counterpart in the
 * single-device reference).  Every aggregation needs the source rows its owned rows gather from.  Instead
 * of an all-gather after the kernel that produced the local row block, that kernel stores each finished
 * row into its own copy AND into the copy of every peer that gathers it, through NVLink peer memory:
 *
 *   cb_peer_alloc / cb_peer_open   a device buffer other processes can map (CUDA IPC; the 64-byte handle is
 *                                  what the ranks exchange).  cb_peer_close unmaps, cb_peer_free releases.
 *   cb_peer_push_t                 where the rows of a [M, N] kernel output go besides `out`:
 *                                  peer[j] + (row0 + m) * ld + n   for every j with bit j of need[m] set.
 * The caller orders "all pushes done" before "anyone gathers" with a barrier on the stream (NCCL).
 */
#define CB_PEER_HANDLE_BYTES 64
#define CB_MAX_PEERS 7
typedef struct {
    int32_t n_peers;             /* 0 .. CB_MAX_PEERS */
    int32_t max_ctas;            /* > 0: cap the kernel's grid (a pushing kernel is NVLink-bound; the SMs it
                                    leaves free run the aggregation of the previous column panel) */
    void* peer[CB_MAX_PEERS];    /* peer-mapped base of each remote [N_global, ld] buffer (element type of `out`) */
    const uint8_t* need;         /* [M] device: bit j set = remote j gathers local row m */
    int64_t row0;                /* global index of local row 0 */
    int64_t ld;                  /* row pitch of the remote buffers, elements */
    const uint8_t* row_live;     /* [M] device or NULL: rows with 0 are known to be all-zero and are not pushed
                                    (the receivers skip them with the same flags, cb_agg_gather row_live) */
    int32_t tile_first;          /* source-panel exchange (cb_graph_create_panelled): this launch computes and pushes */
    int32_t tile_step;           /* only the 128-row tiles tile_first, tile_first + tile_step, ...; 0 / 0 = every tile */
} cb_peer_push_t;
int cb_peer_alloc(int64_t bytes, void** ptr, void* handle_out /* CB_PEER_HANDLE_BYTES */);
int cb_peer_open(const void* handle, void** ptr);
int cb_peer_close(void* ptr);
int cb_peer_free(void* ptr);

/*
 * Dense transform on the tcgen05 tensor cores, fp32 in / fp32 out, "3xTF32" split operands (fp32-class
 * accuracy; the reference GEMM is th.matmul in fp32 with TF32 off, GCN.py:225).
 *
 * cb_gemm_split_weight: hi/lo TF32 split of the small weight operand into the K-major layout the kernel
 *   streams:  hi[n,k] + lo[n,k] ~= (transpose ? W[k,n] : W[n,k]);  hi, lo: [n_rows, k_cols] each.
 *   GCNConv forward (X @ W, W is [in,out], GCN.py:225): transpose=1, n_rows=out, k_cols=in.
 *   Its adjoint dX = dH @ W^T: transpose=0 on the same W (n_rows=in, k_cols=out).
 *   nn.Linear forward (x @ weight^T, weight is [out,in], GCN.py:43,106): transpose=0.
 *
 * cb_gemm_rows:  acc = A[M,K] . Bt[N,K]^T
 *                v   = act( (row_scale ? row_scale[m] : 1) * acc + bias[n] + add[m,n] )
 *                out[m,n] = v ;  out2[m,n] = out2_scale[m] * v
 *   fuses GCN.py:205-213 (out-degree scale, (D X) W = D (X W)), GCN.py:230-231 (+ le), the Linear bias
 *   and relu of GCN.py:104-106, and the next layer's out-degree scale (out2).
 *   A: row pitch lda floats; out/out2/add: row pitches ld_*; every pointer 16-byte aligned, pitches,
 *   N and K multiples of 4 (else CB_E_UNSUPPORTED: the caller then uses a library GEMM).
 */
int cb_gemm_split_weight(const float* W, int64_t n_rows, int64_t k_cols, int transpose, float* hi, float* lo,
                         void* stream);
int cb_gemm_rows_supported(int64_t M, int64_t N, int64_t K);
int cb_gemm_rows(const float* A, int64_t M, int64_t K, int64_t lda, const float* Bt_hi, const float* Bt_lo,
                 int64_t N, const float* row_scale, const float* bias, const float* add, int64_t ld_add, int act,
                 float* out, int64_t ld_out, const float* out2_scale, float* out2, int64_t ld_out2,
                 const cb_peer_push_t* push /* rows of `out` also go to the peers; NULL on one GPU */, void* stream);

/*
 * Adjoint transform with the backward prologue of the layer BELOW fused into its epilogue: one kernel
 * instead of cb_gemm_rows followed by cb_agg_backward_prep (autograd of GCN.py:242-253, 127-128,
 * res_tricks.py:23), or followed by the relu / bias backward of a Linear layer (GCN.py:104-106).
 * Same operations in the same order as the two-kernel path, so G and d_x0 are bit-identical to it:
 *   acc      = A[M,K] . Bt[N,K]^T                                   (dX = dH . W^T)
 *   dtot     = (row_scale ? row_scale[m] : 1) * acc + (add ? add[m,n] : 0)
 *   d_x0[m,n] (+)= alpha * dtot                                     (if d_x0; accumulate_x0: += vs =)
 *   dz       = (mixed ? (1-alpha) : 1) * dtot * gate[m,n]           gate = gate_u8 != 0 | gate_f32 > 0 | 1
 *   out[m,n] = (post_scale ? post_scale[m] : 1) * dz                (G of the layer below; [M, ld_out])
 *   col_sum[n] = sum_m dz[m,n]                                       (bias gradient; per-CTA partials in
 *                                                                     `workspace`, added in CTA order)
 *   row_live[m] = 1 if any out[m,:] != 0                             (if row_live; the caller zeroes it first;
 *                                                                     feeds cb_agg_gather's row_live)
 *   a_live [M] or NULL (not with `add`): 0 = the A row is all-zero (cb_row_any_nonzero), so dtot, dz, out and the d_x0
 *            contribution of that row are zero: nothing of the row is loaded or STORED -- out / d_x0 keep whatever the
 *            buffers held.  The gradient under a loss over the train rows only has ~|train|/N live rows; the consumers
 *            of `out` (cb_agg_gather_compacted with the same flags, the peer pushes) never touch the others.
 *   x0_valid [M] or NULL (with accumulate_x0): 0 = d_x0[m,:] was left unwritten by such a call; it is read as 0.
 * workspace: cb_gemm_rows_grad_workspace_bytes(M, N), needed when col_sum != NULL.
 */
int64_t cb_gemm_rows_grad_workspace_bytes(int64_t M, int64_t N);
int cb_gemm_rows_grad(const float* A, int64_t M, int64_t K, int64_t lda, const float* Bt_hi, const float* Bt_lo,
                      int64_t N, const float* row_scale, const float* add, int64_t ld_add, const uint8_t* gate_u8,
                      const float* gate_f32, int64_t ld_gate, int mixed, double alpha, float* d_x0, int64_t ld_dx0,
                      int accumulate_x0, const float* post_scale, float* out, int64_t ld_out, float* col_sum,
                      uint8_t* row_live, const uint8_t* a_live, const uint8_t* x0_valid, void* workspace,
                      int64_t workspace_bytes, const cb_peer_push_t* push, void* stream);

/* flags[r] = 1 if row r of x [rows, d] (row pitch ld elements, dtype CB_F32 / CB_BF16) holds a non-zero element */
int cb_row_any_nonzero(const void* x, int dtype, int64_t rows, int64_t d, int64_t ld, uint8_t* flags, void* stream);

/*
 * bf16 storage (BASELINE.json configs[4]): A, add, gate_val, d_x0, out, out2 and the weight operand Bt [N, K] hold
 * bf16; the MMA is tcgen05.mma.kind::f16 on the operands as stored (no split: twice the TF32 rate, half the bytes),
 * accumulation in fp32 in tensor memory, the same fp32 epilogues, one round-to-nearest-even at each store.
 * K, lda multiples of 8; N and the epilogue pitches multiples of 4.  cb_gemm_weight_to_bf16 prepares Bt.
 */
int cb_gemm_weight_to_bf16(const float* W, int64_t n_rows, int64_t k_cols, int transpose, uint16_t* out, void* stream);
int cb_gemm_rows_supported_bf16(int64_t M, int64_t N, int64_t K);
int cb_gemm_rows_bf16(const uint16_t* A, int64_t M, int64_t K, int64_t lda, const uint16_t* Bt, int64_t N,
                      const float* row_scale, const float* bias, const uint16_t* add, int64_t ld_add, int act,
                      uint16_t* out, int64_t ld_out, const float* out2_scale, uint16_t* out2, int64_t ld_out2,
                      const cb_peer_push_t* push, void* stream);

/*
 * The same transform (dtype CB_F32: Bt_hi / Bt_lo from cb_gemm_split_weight; CB_BF16: Bt_hi from cb_gemm_weight_to_bf16,
 * Bt_lo NULL) that also writes relu_mask [M, ld_mask] bytes: 1 where the activated output is positive.  The backward's
 * relu gate (cb_gemm_rows_grad gate_u8) then reads a byte per element instead of re-reading the activations
 * (GCN.py:104-106: the input Linear + relu, whose gate the first layer's adjoint GEMM applies).
 */
int cb_gemm_rows_masked(int dtype, const void* A, int64_t M, int64_t K, int64_t lda, const void* Bt_hi, const void* Bt_lo,
                        int64_t N, const float* row_scale, const float* bias, const void* add, int64_t ld_add, int act,
                        void* out, int64_t ld_out, const float* out2_scale, void* out2, int64_t ld_out2,
                        uint8_t* relu_mask, int64_t ld_mask, const cb_peer_push_t* push, void* stream);
int cb_gemm_rows_grad_bf16(const uint16_t* A, int64_t M, int64_t K, int64_t lda, const uint16_t* Bt, int64_t N,
                           const float* row_scale, const uint16_t* add, int64_t ld_add, const uint8_t* gate_u8,
                           const uint16_t* gate_val, int64_t ld_gate, int mixed, double alpha, uint16_t* d_x0,
                           int64_t ld_dx0, int accumulate_x0, const float* post_scale, uint16_t* out, int64_t ld_out,
                           float* col_sum, uint8_t* row_live, const uint8_t* a_live, const uint8_t* x0_valid,
                           void* workspace, int64_t workspace_bytes, const cb_peer_push_t* push, void* stream);

/*
 * Weight gradient on the tcgen05 tensor cores (3xTF32): out[Ka, Nb] = A[M, Ka]^T . B[M, Nb], the reduction
 * running over the M node rows (autograd of GCN.py:225: dW = (D X)^T dH; and of the Linear layers).
 * Split-K over the SMs with a fixed-order second pass, so the result is bit-stable from run to run.
 * row_scale [M] or NULL: out = sum_m row_scale[m] * A[m,:]^T B[m,:] (the out-degree scale of GCN.py:205-213
 * folded into the gradient); it is applied to A's rows, or to B's when scale_b != 0, while the operand is
 * split in shared memory.
 * Needs Ka % 32 == 0, Nb % 32 == 0, 16-byte aligned operands (else CB_E_UNSUPPORTED).
 * workspace: cb_gemm_tn_workspace_bytes(M, Ka, Nb).
 */
int cb_gemm_tn_supported(int64_t M, int64_t Ka, int64_t Nb);
int64_t cb_gemm_tn_workspace_bytes(int64_t M, int64_t Ka, int64_t Nb);
int cb_gemm_tn(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t Ka, int64_t Nb,
               const float* row_scale, int scale_b, float* out, int64_t ld_out, void* workspace,
               int64_t workspace_bytes, void* stream);

/* bf16 operands as stored (MN-major, SWIZZLE_128B, kind::f16), fp32 result; Ka, Nb multiples of 64; no row scale
 * (a bf16 forward hands the pre-scaled copy on).  Same segmenting and fixed-order fp64 second pass. */
int cb_gemm_tn_supported_bf16(int64_t M, int64_t Ka, int64_t Nb);
int64_t cb_gemm_tn_workspace_bytes_bf16(int64_t M, int64_t Ka, int64_t Nb);
int cb_gemm_tn_bf16(const uint16_t* A, int64_t lda, const uint16_t* B, int64_t ldb, int64_t M, int64_t Ka, int64_t Nb,
                    float* out, int64_t ld_out, void* workspace, int64_t workspace_bytes, void* stream);

/*
 * SEMLP "virtual neighbour" replacement (MLP_model/__init__.py:143-156: per node, le_guess[i] . teacherSE^T -> the K
 * largest scores -> soft-max -> weighted sum of those K teacher rows) without the [B, N] score matrix:
 *   cb_topk_merge        folds one score tile [B, width] (row pitch ld; column c is teacher row col_base + c; computed
 *                        by cb_gemm_rows with the teacher table as the weight operand) into the running per-row lists
 *                        top_val / top_idx [B, 32] (descending; first != 0 starts them).  K <= 32.
 *   cb_topk_softmax_mix  out[b,:] = sum_k softmax(top_val[b,:K])[k] * table[top_idx[b,k],:], added in ascending score
 *                        order like the reference's [1, K] x [K, d] product.
 */
int cb_topk_merge(const float* scores, int64_t B, int64_t width, int64_t ld, int64_t col_base, int K, float* top_val,
                  int32_t* top_idx, int first, void* stream);
int cb_topk_softmax_mix(const float* top_val, const int32_t* top_idx, int64_t B, int K, const float* table, int64_t d,
                        int64_t ld_table, float* out, void* stream);

/*
 * Graph preparation either side of the path: the reference's host loops over .tolist()-ed edge lists as
 * integer kernels, same values and the same ORDER.  Setup-time calls: scratch is the caller's
 * (workspace of cb_prep_graph_workspace_bytes(what, n) bytes, device memory), the stream is synchronised once or twice (a count,
 * or the key range that sizes the radix sort, goes back to the host).  Device pointers unless marked host.
 *
 *   cb_prep_degrees            utils.py:300-334 graph_analyze: edges per node as origin / as destination, int64 [N]
 *                              (ids >= N are ignored like the reference's range(N_nodes) read-out; negative: CB_E_RANGE)
 *   cb_prep_symmetrize         utils.py:667-674 ensure_symmetric: the coalesced indices of A + A^T, N = max id + 1,
 *                              sorted by (row, col).  out: [2, 2 * num_edges] (row pitch 2 * num_edges), *count (host)
 *                              columns are valid
 *   cb_prep_partial_sorted_idx utils.py:910-943 get_partial_sorted_idx on an int64 array: `levels` rounds (1..5 = the
 *                              50 / 25 / 12 / 6 / 3 modes) of "keep what is <= (top != 0) or >= (top == 0) the numpy
 *                              median of what was kept", then the indices in ascending order: idx_out [n], *count (host)
 *   cb_prep_degree_stats       utils.py:676-678 gen_rec_for_table1_stats: stats (host, 6 doubles) = N, sum, max, mean,
 *                              numpy median, percentage of zeros
 *   cb_prep_sort_idx_by_value  idx[np.argsort(arr[idx], kind='stable')] (utils.py:703-704; numpy's default sort there
 *                              is not stable, ties are resolved by position in idx here)
 *   cb_prep_mask_from_idx      mask[N] bytes: 1 at the listed ids, 0 elsewhere (utils.py:694-697, 711-717)
 *   cb_prep_drop_edges         utils.py:732-752 craft_isolation_v2: drops every edge that is not a self loop and touches
 *                              a node with node_mask != 0, keeping the order.  out: [2, num_edges] (row pitch num_edges),
 *                              *kept (host) columns are valid
 */
enum {  /* `what` of cb_prep_graph_workspace_bytes; n = the size named beside it */
    CB_PREP_DEGREES = 0,            /* n = num_nodes */
    CB_PREP_SYMMETRIZE = 1,         /* n = num_edges */
    CB_PREP_PARTIAL_SORTED_IDX = 2, /* n = length of arr */
    CB_PREP_DEGREE_STATS = 3,       /* n = length of degs */
    CB_PREP_SORT_IDX_BY_VALUE = 4,  /* n = m (length of idx) */
    CB_PREP_MASK_FROM_IDX = 5,      /* n ignored */
    CB_PREP_DROP_EDGES = 6          /* n = num_edges */
};
int64_t cb_prep_graph_workspace_bytes(int what, int64_t n);
int cb_prep_degrees(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes, int64_t* degs_ori,
                    int64_t* degs_dst, void* workspace, int64_t workspace_bytes, void* stream);
int cb_prep_symmetrize(const int64_t* edge_index, int64_t num_edges, int64_t* out, int64_t* count, void* workspace,
                       int64_t workspace_bytes, void* stream);
int cb_prep_partial_sorted_idx(const int64_t* arr, int64_t n, int top, int levels, int64_t* idx_out, int64_t* count,
                               void* workspace, int64_t workspace_bytes, void* stream);
int cb_prep_degree_stats(const int64_t* degs, int64_t n, double* stats, void* workspace, int64_t workspace_bytes,
                         void* stream);
int cb_prep_sort_idx_by_value(const int64_t* arr, int64_t n, const int64_t* idx, int64_t m, int64_t* idx_sorted,
                              void* workspace, int64_t workspace_bytes, void* stream);
int cb_prep_mask_from_idx(const int64_t* idx, int64_t m, int64_t num_nodes, uint8_t* mask, void* workspace,
                          int64_t workspace_bytes, void* stream);
int cb_prep_drop_edges(const int64_t* edge_index, int64_t num_edges, const uint8_t* node_mask, int64_t num_nodes,
                       int64_t* out, int64_t* kept, void* workspace, int64_t workspace_bytes, void* stream);

/* number of kernel launches issued by this library since load (bench.py's gpu_launches claim) */
int64_t cb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* COLDBREW_B200_H */
