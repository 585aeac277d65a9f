/* gnnome_b200.h — C ABI of the B200-native GatedGCN message-passing engine.
 *
 * The reference (lvrcek/GNNome-assembly) has no FFI: its hot path is Python calling PyTorch and DGL.
 * This header IS the new seam.  Each entry point names the reference interface it replaces
 * (file:line relative to the reference repo).  All pointers are raw DEVICE pointers to fp32
 * row-major contiguous arrays unless stated otherwise; `stream` is a cudaStream_t passed as void*.
 * The caller owns every buffer (inputs, outputs, saved activations, workspace); the library owns
 * only gg_plan_t.  No hidden allocation, no hidden stream, no device synchronisation inside the
 * compute calls.  Return value: 0 = ok, <0 = error (see GG_ERR_*), message via gg_last_error().
 *
 * Supported sizes: hidden d in {64,128,256}; predictor hidden H = 64; indices int32.
 */
#ifndef GNNOME_B200_H
#define GNNOME_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GG_OK 0
#define GG_ERR_ARG (-1)         /* null pointer, bad size, index out of range */
#define GG_ERR_UNSUPPORTED (-2) /* hidden size / option the kernels are not built for */
#define GG_ERR_CUDA (-3)        /* a CUDA runtime call failed; gg_last_error() has the string */

#define GG_NORM_BATCH 0 /* nn.BatchNorm1d(track_running_stats=False), gated_gcn_full.py:55-56 */
#define GG_NORM_LAYER 1 /* nn.LayerNorm, gated_gcn_full.py:58-59 */

typedef struct gg_plan gg_plan_t;

int gg_version(void);          /* 100 = round-1 entry points; 200 = + whole-model calls, gg_edge_mlp_fwd, plan flags (a superset) */
const char* gg_last_error(void);

/* ---- instrumentation ------------------------------------------------------------------------
 * gg_launch_count: number of kernels this library has launched in this process (bench.py's gpu_launches).
 * gg_profile_enable(1): bracket every launch with a CUDA event pair on the launching stream;
 * gg_profile_report: synchronise those events, write {"kernel": [launches, total_ms], ...} as JSON text
 * into buf and clear the records.  The reference has torch.profiler imported but unused (train.py:16). */
/* gg_set_tc_mode(1) (default): projections with N % 128 == 0 run on the tcgen05 3xTF32 tensor-core kernel;
 * gg_set_tc_mode(0): true-fp32 FFMA kernel everywhere.  Returns the previous mode. */
int gg_set_tc_mode(int mode);
/* experiment switches for the tensor-core GEMM epilogue (tools/epi_experiment.py): bit 0 skips the column
 * statistics, bit 1 skips the epilogue operand prefetch (WRONG RESULTS; timing experiments only); bit 3 (8) runs
 * the forward edge-gate pass with its streamed operands staged through shared memory by cp.async.bulk (correct
 * results; measured slower than the default register-staged kernel, kept as an experiment); bit 6 (64) switches the
 * zig-zag traversal of consecutive E-sized kernels off, bit 7 (128) keeps the weight-gradient GEMMs of gg_model_bwd on
 * the main stream (same results, A/B timing). Returns old. */
int gg_debug_flags(int flags);
/* Trace builds only (-DGG_TC_TRACE, tools/ab_build.sh + tools/tc_trace.py): per-role clock64 totals of the tcgen05
 * GEMM (cycles inside each mbarrier wait / per role), summed over CTAs since the last reset; synchronises the
 * device.  A default build returns zeros.  Slot order: gg_gemm_tc.cuh, enum TraceSlot. */
int gg_debug_trace(unsigned long long* out, int n, int reset);
int64_t gg_launch_count(void);
int gg_profile_enable(int on);
int gg_profile_report(char* buf, size_t cap);

/* ---- graph plan -------------------------------------------------------------------------
 * Replaces the structure side of the DGLGraph argument of GraphGatedGCNModel.forward
 * (models/full_graph.py:22) and dgl.reverse (layers/gated_gcn_full.py:115).
 * src/dst: int32[E] in the caller's edge-id order, HOST or DEVICE memory.  Builds the internal edge order
 * (stable sort by dst = CSR over in-edges) and a CSR over out-edges that points into it, once per graph, ON THE
 * DEVICE (radix sorts + a two-level locality relabelling of the nodes, csrc/gg_plan_device.cu) into one
 * library-owned device allocation; synchronises `stream` once (index check).  Internal position p holds caller
 * edge perm[p].
 * (Per-batch sub-graph plans are built on the device: gg_subplan_* below.) */
int gg_plan_create(const int32_t* src, const int32_t* dst, int64_t num_nodes, int64_t num_edges,
                   void* stream, gg_plan_t** out);
/* flags: GG_PLAN_RELABEL (default of gg_plan_create) renumbers the nodes breadth-first so that the
 * neighbours of a node are close in memory (assembly graphs are near-linear, read ids are arbitrary).
 * Node features then live in INTERNAL node order: row p of h holds caller node node_perm[p]. */
#define GG_PLAN_RELABEL 1
/* GG_PLAN_HOST_BUILD: build on the host (the round-1 builder: D2H of the edge list, counting sort, exact breadth-first
 * relabelling on one core, one upload) instead of on the device; also selected by the environment variable
 * GG_PLAN_HOST=1.  Same arrays except for the node order under GG_PLAN_RELABEL (both are locality orders; the
 * device one is a two-level region order, see csrc/gg_plan_device.cu). */
#define GG_PLAN_HOST_BUILD 2
int gg_plan_create_ex(const int32_t* src, const int32_t* dst, int64_t num_nodes, int64_t num_edges, int flags,
                      void* stream, gg_plan_t** out);
int gg_plan_destroy(gg_plan_t* plan);
int64_t gg_plan_num_nodes(const gg_plan_t* plan);
int64_t gg_plan_num_edges(const gg_plan_t* plan);
/* device int32[E]: perm (internal position -> caller edge id) and its inverse */
const int32_t* gg_plan_perm(const gg_plan_t* plan);
const int32_t* gg_plan_inv_perm(const gg_plan_t* plan);
/* device int32 arrays in internal order: src[E], dst[E], in_ptr[N+1], out_ptr[N+1], out_eid[E] */
const int32_t* gg_plan_src(const gg_plan_t* plan);
const int32_t* gg_plan_dst(const gg_plan_t* plan);
const int32_t* gg_plan_in_ptr(const gg_plan_t* plan);
const int32_t* gg_plan_out_ptr(const gg_plan_t* plan);
const int32_t* gg_plan_out_eid(const gg_plan_t* plan);
/* copy one of those arrays into a caller-owned device buffer (async on `stream`);
 * which: 0 perm, 1 inv_perm, 2 src, 3 dst, 4 in_ptr, 5 out_ptr, 6 out_eid, 7 node_perm[N] (internal ->
 * caller node id), 8 node_inv[N]; on plans made by gg_subplan_fill also 9 parent_eid[E] (sub-graph edge id
 * -> parent edge id, DGL's edata[dgl.EID]), 10 / 11 the sub-graph's own edge list src[E] / dst[E] in
 * sub-graph edge-id order and sub-graph node ids (what sub_g.edges() returns) */
int gg_plan_copy_array(const gg_plan_t* plan, int which, int32_t* out, void* stream);

/* ---- node-induced sub-graph plan (mini-batch path) ------------------------------------------------
 * Replaces dgl.dataloading.ClusterGCNSampler.sample -> g.subgraph(node_ids) as the reference's mini-batch
 * branch uses it (train.py:292-296 and :434-438; dgl.node_subgraph at inference.py:271) together with the
 * plan the engine needs for the result.  nodes: DEVICE int64[n] parent node ids (the reference runs this
 * branch on g.long(), train.py:290), unique; sub-graph node j is parent node nodes[j] and the sub-graph's
 * edges are all parent edges with both ends selected, in increasing parent edge id (DGL's convention).
 * Built entirely on the device as a stream compaction of the parent plan (4 exclusive scans + scatters,
 * no sort): the selected nodes keep the parent's internal (breadth-first) order, so every edge array stays
 * sorted and the gather locality of the parent carries over.
 * Two steps, because the caller owns every buffer and the result's size is data dependent:
 *   gg_subplan_count  marks the nodes and runs the scans into scratch owned by the parent plan, synchronises
 *                     `stream` (one int has to reach the host) and returns the sub-graph's edge count;
 *   gg_subplan_fill   scatters the result into `slab` (DEVICE int32[gg_subplan_slab_words(n, num_edges)],
 *                     caller-owned and kept alive as long as the plan) and returns the plan.
 * The pair must be issued on ONE stream, count then fill, with no other gg_subplan_* call on the same parent
 * in between (the scratch belongs to the parent).  The result is an ordinary plan: every gg_layer_* /
 * gg_score_* / gg_prep_* entry point and gg_subplan_* itself accept it; gg_plan_destroy releases the handle
 * (not the slab). */
size_t gg_subplan_slab_words(int64_t n, int64_t num_edges);
int gg_subplan_count(const gg_plan_t* parent, const int64_t* nodes, int64_t n, void* stream, int64_t* num_edges);
int gg_subplan_fill(const gg_plan_t* parent, int32_t* slab, void* stream, gg_plan_t** out);

/* ---- dense linear ( nn.Linear ) -----------------------------------------------------------
 * Replaces torch.nn.functional.linear at models/full_graph.py:23-26 (linear_pe, linear1_edge,
 * linear2_edge).  Y[M,N] = X[M,K] * W[N,K]^T + b (b may be NULL); relu!=0 applies max(.,0).
 * K and all leading dimensions must be multiples of 4 (callers zero-pad K = 18 -> 20, 2 -> 4). */
int gg_linear_fwd(int64_t M, int N, int K, const float* X, const float* W, const float* b, int relu,
                  float* Y, void* stream);
/* dX[M,K] = dY[M,N] * W[N,K]  (+ addend[M,K] if not NULL); if relu_mask != NULL, dX is zeroed where
 * relu_mask[M,K] <= 0 (relu_mask is the post-ReLU activation that produced X) */
int gg_linear_bwd_data(int64_t M, int N, int K, const float* dY, const float* W, const float* addend,
                       const float* relu_mask, float* dX, void* stream);
/* dW[N,K] = dY[M,N]^T * X[M,K];  db[N] = column sums of dY (db may be NULL).
 * dW and db are overwritten (zeroed inside, then accumulated split-K with fp32 atomics). */
int gg_linear_bwd_weight(int64_t M, int N, int K, const float* dY, const float* X, float* dW, float* db,
                         void* stream);

/* ---- one GatedGCN layer ---------------------------------------------------------------------
 * Replaces GatedGCN_1d.forward, layers/gated_gcn_full.py:99-157 (K1-K13 of SURVEY.md §2b).
 * h_in[N,d], e_in[E,d] (e in INTERNAL edge order).  Wn[5d,d] = rows of A_1,A_2,A_3,B_1,B_2 stacked,
 * bn[5d] their biases; B3[d,d], b3[d]; gamma/beta of bn_e and bn_h.
 * Outputs h_out[N,d], e_out[E,d].  Saved for backward / scratch (caller allocated):
 *   P[N,5d] projections, t[E,d] pre-norm edge gate, z[N,d] pre-norm node update,
 *   agg[5,N,d] = hf | hb | 1/(den_f+eps) | 1/(den_b+eps) | sum over in-edges of xhat_e (batch norm),
 *   stats[4d] doubles: sum_t | sum_t^2 | sum_z | sum_z^2 per channel (batch-norm only). */
int gg_layer_fwd(const gg_plan_t* plan, int d, int norm_kind, int residual, const float* h_in,
                 const float* e_in, const float* Wn, const float* bn, const float* B3, const float* b3,
                 const float* gamma_e, const float* beta_e, const float* gamma_h, const float* beta_h,
                 float* h_out, float* e_out, float* P, float* t, float* z, float* agg, double* stats,
                 void* stream);

/* Backward of the layer (the reference has no code for it: torch.autograd replays K16 of SURVEY §2b).
 * g_h[N,d], g_e[E,d]: gradients w.r.t. h_out / e_out (either may be NULL = zero).
 * Outputs: g_h_in[N,d], g_e_in[E,d], dWn[5d,d], dbn[5d], dB3[d,d], db3[d], dgamma/dbeta (4 x [d]).
 * Workspace: gP[N,5d], G[2,N,2d], g_eo[E,d], g_t[E,d], bstats[4d] doubles.  Pass g_e_in distinct from g_eo:
 * with batch norm on the tensor-core path g_t is then produced inside the g_e_in GEMM (no separate pass);
 * an aliased g_e_in still works but takes the unfused path. */
int gg_layer_bwd(const gg_plan_t* plan, int d, int norm_kind, int residual, const float* h_in,
                 const float* e_in, const float* e_out, const float* Wn, const float* B3,
                 const float* gamma_e, const float* beta_e, const float* gamma_h, const float* beta_h,
                 const float* P, const float* t, const float* z, const float* agg, const double* stats,
                 const float* g_h, const float* g_e, float* g_h_in, float* g_e_in, float* dWn, float* dbn,
                 float* dB3, float* db3, float* dgamma_e, float* dbeta_e, float* dgamma_h, float* dbeta_h,
                 float* gP, float* G, float* g_eo, float* g_t, double* bstats, void* stream);

/* ---- edge score predictor -----------------------------------------------------------------
 * Replaces ScorePredictor.forward, layers/score_predictor.py:12-25 (K15), without materialising
 * the E x 3d concat: W1 = [W1s | W1d | W1e] (columns), Wq[2H,d] = [W1s ; W1d] stacked by rows,
 * bq[2H] = [b1 ; 0], W1e[H,d], w2[H], b2[1].  x[N,d], e[E,d] internal order.
 * Outputs score[E] (internal order); Q[N,2H] scratch; hid[E,H] saved post-ReLU hidden (NULL to skip). */
int gg_score_fwd(const gg_plan_t* plan, int d, int H, const float* x, const float* e, const float* Wq,
                 const float* bq, const float* W1e, const float* w2, const float* b2, float* score,
                 float* Q, float* hid, void* stream);
/* g_score[E].  Outputs g_x[N,d], g_e[E,d], dWq[2H,d], dbq[2H] (second half zero), dW1e[H,d], dw2[H],
 * db2[1].  Workspace g_pre[E,H] (may alias hid), gQ[N,2H], red[2H+1] doubles. */
int gg_score_bwd(const gg_plan_t* plan, int d, int H, const float* x, const float* e, const float* Wq,
                 const float* W1e, const float* w2, const float* g_score, const float* hid, float* g_x,
                 float* g_e, float* dWq, float* dbq, float* dW1e, float* dw2, float* db2, float* g_pre,
                 float* gQ, double* red, void* stream);

/* ---- whole model in one call -------------------------------------------------------------------
 * Replaces GraphGatedGCNModel.forward, models/full_graph.py:22-29 (linear_pe, linear1_edge -> ReLU -> linear2_edge,
 * L x GatedGCN_1d, ScorePredictor) and its autograd: a host-side sequencer over the entry points above, so that one
 * forward (and one backward) is ONE call issuing ~60 launches back to back instead of ~150 bound calls — what makes the
 * reference's default cluster mini-batch path (train.py:282-312: batches of 25-40 k edges) GPU-bound instead of
 * host-bound.  `params` is the model's FLAT fp32 parameter buffer, `offs` an int64 table of offsets (in floats) into it,
 * n_offs = 10 + 8 * layers entries:
 *   0 linear_pe.weight[d, node_in] 1 linear_pe.bias  2 linear1_edge.weight[he, edge_in] 3 .bias  4 linear2_edge.weight[d, he]
 *   5 .bias  6 predictor.W1.weight[H, 3d] 7 .bias  8 predictor.W2.weight[1, H] 9 .bias
 *   10 + 8 l + {0 Wn[5d, d] = A_1,A_2,A_3,B_1,B_2 stacked, 1 bn[5d], 2 B_3.weight, 3 B_3.bias, 4 bn_e.weight, 5 bn_e.bias,
 *               6 bn_h.weight, 7 bn_h.bias}
 * Every offset is a multiple of 4 floats and both buffers are 16-byte aligned (vector and TMA accesses; GG_ERR_ARG otherwise).
 * `grads` (gg_model_bwd) is a flat buffer addressed by the SAME table.  e[E, edge_in], pe[N, node_in] and scores[E] are in
 * the CALLER's edge / node order (the permutation to the plan's internal order happens inside).
 * Workspaces are caller-owned: gg_model_workspace_floats(plan, m, which) floats, which = 0 forward that keeps what the
 * backward reads, 1 forward for inference (layer buffers reused), 2 backward scratch.  gg_model_bwd runs the phases
 * [phase_begin, phase_end) of: 0 predictor, 1 + k layer L-1-k, L + 1 encoders — a caller that all-reduces gradients
 * per layer issues one call per phase; 0 .. L + 2 does everything.  A pass starts with phase 0: it zeroes every accumulator
 * of the pass in two memsets (backward scratch + the span of `grads` the table covers, when the table's only gaps are
 * alignment padding of < 4 floats; any other table makes each op zero its own outputs instead).
 * side_stream (may be NULL): a second caller-owned stream.  The weight-gradient GEMMs of every layer (dB3, dWn: nothing
 * downstream reads them) are forked onto it and joined back into `stream` by the call that runs the last phase, so
 * after a whole backward everything is ordered on `stream` as usual; a caller that consumes a layer's gradients between
 * phases must order that consumer after BOTH streams. */
typedef struct gg_model_desc {
  int32_t d;            /* hidden_features: 64, 128 or 256 */
  int32_t layers;       /* num_layers */
  int32_t hidden_edge;  /* hidden_edge_features (16) */
  int32_t hidden_score; /* hidden_edge_scores (64) */
  int32_t norm_kind;    /* GG_NORM_BATCH / GG_NORM_LAYER */
  int32_t node_in;      /* nb_pos_enc + 2 */
  int32_t edge_in;      /* edge_features (2) */
  int32_t reserved;
} gg_model_desc_t;
int64_t gg_model_workspace_floats(const gg_plan_t* plan, const gg_model_desc_t* m, int which);
int gg_model_fwd(const gg_plan_t* plan, const gg_model_desc_t* m, const float* params, const int64_t* offs, int n_offs,
                 const float* e, const float* pe, int training, float* ws, float* scores, void* stream);
int gg_model_bwd(const gg_plan_t* plan, const gg_model_desc_t* m, const float* params, const int64_t* offs, int n_offs,
                 const float* g_scores, const float* ws, float* bws, float* grads, int phase_begin, int phase_end,
                 void* stream, void* side_stream);

/* ---- input preparation ("next" row 1 of SURVEY.md 8f) ----------------------------------------------
 * gg_prep_edge_features replaces utils.preprocess_graph, utils.py:70-74: e[E,2] = z-scored overlap_length,
 * overlap_similarity (mean / unbiased std per graph), caller edge order in and out.  ws: 4 doubles.
 * gg_prep_pe replaces utils.add_positional_encoding, utils.py:102-138 (type_pe == 'PR') and the concat of
 * train.py:249-251: pe[N, 2 + pe_dim] = in_degree | out_degree | pe_dim PageRank iterates (alpha = 0.95 in the
 * reference), fp64 internally like scipy, rows in CALLER node order.  ws: 3N doubles. */
int gg_prep_edge_features(int64_t E, const float* overlap_length, const float* overlap_similarity, float* e_out,
                          double* ws, void* stream);
int gg_prep_pe(const gg_plan_t* plan, int pe_dim, double alpha, float* pe_out, double* ws, void* stream);

/* ---- fused loss + metrics ("next" row 2 of SURVEY.md 8f) -----------------------------------------------
 * Replaces BCEWithLogitsLoss(pos_weight) (train.py:211,255) and utils.calculate_tfpn (utils.py:217-223),
 * four .item() syncs per step in the reference.  out5 (doubles): sum of per-edge losses | TP | TN | FP | FN.
 * gg_bce_bwd: g_scores[i] = g_loss[0] / E * d(loss_i)/d(score_i)  (mean reduction). */
int gg_bce_metrics_fwd(int64_t E, const float* scores, const float* y, float pos_weight, double* out5, void* stream);
int gg_bce_bwd(int64_t E, const float* scores, const float* y, float pos_weight, const float* g_loss,
               float* g_scores, void* stream);

/* ---- edge-order helpers ------------------------------------------------------------------------
 * out[p, :] = in[idx[p], :]  (rows of `width` floats).  Used to move e[E,2] / scores[E] between the
 * caller's edge-id order (the contract at the model boundary) and the internal order. */
int gg_gather_rows(int64_t rows, int width, const float* in, const int32_t* idx, float* out, void* stream);

/* ---- edge encoder backward ------------------------------------------------------------------
 * Autograd of linear2_edge(relu(linear1_edge(e))) (models/full_graph.py:24-26) in one pass over the E x d
 * gradient g: dW2[d,hidden] = g^T hid, db2[d] = colsum g, g_hid = (g W2) * [hid > 0] (never stored),
 * dW1[hidden,K] = g_hid^T e, db1[hidden] = colsum g_hid.  hid[E,hidden] = relu(W1 e + b1) from the forward,
 * e[E,K] the (zero-padded) raw edge features.  Built for hidden = 16, K = 4, d in {64,128}; other shapes return
 * GG_ERR_UNSUPPORTED and the caller uses gg_linear_bwd_weight / gg_linear_bwd_data.  Outputs are zeroed here. */
int gg_edge_mlp_bwd(int64_t E, int d, int hidden, int K, const float* g, const float* hid, const float* e,
                    const float* W2, float* dW1, float* db1, float* dW2, float* db2, void* stream);
/* The forward of the same two layers (models/full_graph.py:24-26) in one pass: hid[E,hidden] = relu(W1 e + b1) (kept for
 * the backward), out[E,d] = W2 hid + b2.  e[E,K] zero-padded raw edge features, W1[hidden,K] padded alike.  Same shape
 * limits and fallback (two gg_linear_fwd calls) as gg_edge_mlp_bwd; e, W1, W2, hid, out 16-byte aligned. */
int gg_edge_mlp_fwd(int64_t E, int d, int hidden, int K, const float* e, const float* W1, const float* b1,
                    const float* W2, const float* b2, float* hid, float* out, void* stream);

/* ---- greedy contig decoding -------------------------------------------------------------------
 * Replaces the body of get_contigs (inference.py:182-259): per decoding iteration the reference samples
 * nb_paths start edges and runs walk_forwards / walk_backwards (inference.py:31-77) for each of them one after
 * the other in Python.  Here all walks of an iteration run concurrently, one warp per walk.
 * Adjacency is given in CALLER node ids (the reference's succs / preds dictionaries flattened to CSR, list
 * order kept: ties in the arg-max and the single-neighbour shortcut depend on it): *_ptr int32[N+1], *_node
 * int32[E], *_eid int32[E] = id of the edge (current -> neighbour) for successors, (neighbour -> current) for
 * predecessors (the reference's `edges` dictionary).  Node 2k and 2k+1 are the two strands of one read
 * (`current ^ 1`, inference.py:39).  The graph must have no self loops (the reference drops them, :187).
 * scores float32[E] (edata['score'], or overlap_length / overlap_similarity for the baselines :134-141),
 * prefix_length int64[E], read_length int64[N] (get_contig_length, inference.py:20-28), all by caller ids.
 * visited: uint32[(N+31)/32] bitmap of the nodes of finished contigs.
 *
 * gg_decode_walks: for walk w, forward from start_dst[w] then backward from start_src[w] (with the forward
 * walk's nodes masked, :238); writes the node sequence to walk_buf[w*2N + out_beg[w] .. + out_len[w]), the
 * walk's own visited set (its nodes and their strand mates) to local_visited[w*words ..], and the length of
 * the reconstructed sequence to out_seq_len[w].  *err (device int) is set to 1 if a walk ran into a cycle of
 * single-neighbour nodes (the reference never returns in that case; the walk is cut at N nodes).
 * Buffers (caller-owned, device): local_visited uint32[n_walks*words], walk_buf int32[n_walks*2N]. */
int gg_decode_walks(int64_t N, const int32_t* succ_ptr, const int32_t* succ_node, const int32_t* succ_eid,
                    const int32_t* pred_ptr, const int32_t* pred_node, const int32_t* pred_eid, const float* scores,
                    const int64_t* prefix_length, const int64_t* read_length, const uint32_t* visited, int n_walks,
                    const int32_t* start_src, const int32_t* start_dst, const int32_t* start_eid,
                    uint32_t* local_visited, int32_t* walk_buf, int32_t* out_beg, int32_t* out_len,
                    int64_t* out_seq_len, int* err, void* stream);
/* gg_decode_commit: inference.py:223-234,241 for the chosen walk: visited |= walk_visited | {t, t^1 : t in
 * succs[ss] & preds[dd] for consecutive walk nodes ss, dd}.  walk: device int32[len]. */
int gg_decode_commit(int64_t N, const int32_t* succ_ptr, const int32_t* succ_node, const int32_t* pred_ptr,
                     const int32_t* pred_node, const int32_t* walk, int len, const uint32_t* walk_visited,
                     uint32_t* visited, void* stream);
/* gg_decode_edge_weights: the (unnormalised) sampling weights of sample_edges (inference.py:279-286) on the
 * graph without the visited nodes (get_subgraph, :262-275): max(sigmoid(score), 1e-9) for an edge whose two
 * ends are unvisited and distinct, 0 otherwise.  src/dst int32[E] in edge-id order. */
int gg_decode_edge_weights(int64_t E, const int32_t* src, const int32_t* dst, const float* scores,
                           const uint32_t* visited, float* weights, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GNNOME_B200_H */
