// gg_api.cu — C ABI entry points (include/gnnome_b200.h): dense linear, GatedGCN layer forward /
// backward, edge score predictor forward / backward.  Host code only sequences kernel launches on the
// caller's stream; every buffer is caller-owned.
#include "gg_common.cuh"
#include "gg_gemm_ffma.cuh"
#include "gg_gemm_tc.cuh"
#include "gg_layer_kernels.cuh"
#include "gg_layer_bulk.cuh"

namespace gg {

// per-device caches (one process may drive several GPUs: the reference's default device is 'cuda:3',
// hyperparameters.py:25): everything cached about "the GPU" is keyed by the current device ordinal
static int sm_count() {
  static int n[kMaxDevices] = {};
  const int dev = current_device();
  if (n[dev] == 0) {
    if (cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

// one warp per node, 8 warps per CTA, grid-stride over the nodes.  The grid is exactly ONE wave: as many CTAs as
// are resident at once for this kernel's register use (occupancy API, cached per kernel) x the SM count -- a
// grid of 8 CTAs per SM ran as 2.7 waves of 3 resident CTAs, the last one a third empty.
template <class Kern>
static unsigned node_grid(Kern kern, int64_t n) {
  static int per_sm_dev[kMaxDevices] = {};    // one static per kernel instantiation, per device
  int& per_sm = per_sm_dev[current_device()];
  if (per_sm == 0) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kNodeThreads, 0) != cudaSuccess || occ < 1) occ = 2;
    per_sm = occ;
  }
  int64_t blocks = (n + (kNodeThreads / 32) - 1) / (kNodeThreads / 32);
  const int64_t cap = (int64_t)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

// Zig-zag traversal.  An E x d tensor (191 MB on the bench graph) does not fit the 126 MB L2, and a producer that
// writes it front to back followed by a consumer that reads it front to back is the worst case for a recency-managed
// cache: by the time the consumer starts, the front has been evicted and the back — which IS resident — is evicted
// before it is reached.  So consecutive E-sized kernels walk their rows / nodes in OPPOSITE directions: the consumer
// starts where the producer stopped.  Forward of layer l: gemm_edge_gate b, edge_gate_fwd !b, node_agg_fwd b, with
// b = l & 1 (the next layer's GEMM reads e_out where node_agg_fwd has just finished).  Backward: edge_bwd_a forward,
// gemm_bwd_e_in reverse, edge_bwd_src forward (every layer: the upstream g_e was written by a reverse GEMM).
// Per-tile / per-node arithmetic is unchanged.  gg_debug_flags bit 6 (64) switches it off (A/B).
int gg_debug_flags_peek() { return tc::tc_dbg_ref(); }
static thread_local int g_layer_parity = 0;
void set_layer_parity(int p) { g_layer_parity = p & 1; }
// gg_model_* zeroes every accumulator of a pass (statistics, split-K weight gradients) with ONE memset up front and sets
// this; the entry points below then skip their own memsets (60 memset nodes per training step otherwise).
static thread_local bool g_prezeroed = false;
void set_prezeroed(bool on) { g_prezeroed = on; }
bool prezeroed() { return g_prezeroed; }
static inline int zig(int dir) { return (tc::tc_dbg_ref() & 64) ? 0 : (dir & 1); }

// Weight-gradient GEMMs off the critical path.  dB3 = g_t^T e_in and dWn = gP^T h_in feed nothing but the optimizer: the
// next layer's backward needs only g_h_in and g_e_in.  When the whole-model sequencer (gg_model_bwd) provides a side
// stream, gg_layer_bwd forks after edge_bwd_src, runs those two split-K GEMMs there and records `done`; the sequencer
// alternates the buffers they read (gP, g_t) between layers and joins at the end.  On a full-size graph they then
// overlap the next layer's node kernels (latency-bound, far from filling the GPU); on mini-batch sub-graphs they leave
// the kernel-latency chain altogether.
struct SideCtx { cudaStream_t side = nullptr; cudaEvent_t fork = nullptr, done = nullptr; };
static thread_local SideCtx g_side;
void set_layer_bwd_side(cudaStream_t side, cudaEvent_t fork, cudaEvent_t done) { g_side.side = side; g_side.fork = fork; g_side.done = done; }

// 1: dense projections with N % 128 == 0 run on the tcgen05 3xTF32 kernel; 0: FFMA everywhere
static int g_tc_mode = 1;
#ifdef GG_NO_WRES             // A/B builds: per-stage B operand everywhere
#define GG_WRES(D) false
#else
#define GG_WRES(D) ((D) <= 128)
#endif
namespace tc { int& tc_dbg_ref() { static int v = 0; return v; } }

static int pick_splits(int64_t K, int64_t tiles) {
  int64_t want = (2LL * sm_count() + tiles - 1) / tiles;
  int64_t max_by_k = (K + 511) / 512;          // at least 512 reduction rows per split
  if (want > max_by_k) want = max_by_k;
  if (want < 1) want = 1;
  return (int)want;
}

// ------------------------------------------------------------------ generic GEMM front-ends
// Y[M,N] = X[M,K] W[N,K]^T (+b)(relu)
static int linear_fwd(const char* tag, int64_t M, int N, int K, const float* X, int64_t ldx, const float* W, int64_t ldw,
                      const float* b, int relu, float* Y, int64_t ldy, cudaStream_t st) {
  GemmArgs g{};
  g.A = X; g.lda = ldx; g.B = W; g.ldb = ldw; g.M = M; g.N = N; g.K = K;
  EpiBias epi{Y, ldy, b, relu};
  if (g_tc_mode && tc::eligible(false, false, M, N, K, ldx, ldw, X, W))
    return tc::launch<false, false, false, false>(tag, X, ldx, W, ldw, M, N, K, 1, nullptr, nullptr, epi, sm_count(), st);
  if (N >= 128) return launch_gemm<128, false, true, false, false>(tag, g, epi, 1, st);
  if (N > 32) return launch_gemm<64, false, true, false, false>(tag, g, epi, 1, st);
  return launch_gemm<32, false, true, false, false>(tag, g, epi, 1, st);
}

// dX[M,K] = dY[M,N] W[N,K] (+addend)(mask)
template <bool kAdd, bool kMask>
static int linear_bwd_data_t(const char* tag, int64_t M, int N, int K, const float* dY, int64_t lddy, const float* W,
                             int64_t ldw, const float* addend, const float* mask, float* dX, int64_t lddx,
                             cudaStream_t st) {
  GemmArgs g{};
  g.A = dY; g.lda = lddy; g.B = W; g.ldb = ldw; g.M = M; g.N = K; g.K = N;
  EpiAddMaskT<kAdd, kMask> epi{dX, lddx, addend, mask};
  if (g_tc_mode && tc::eligible(false, true, M, K, N, lddy, ldw, dY, W))
    return tc::launch<false, true, false, false>(tag, dY, lddy, W, ldw, M, K, N, 1, nullptr, nullptr, epi, sm_count(), st);
  if (K >= 128) return launch_gemm<128, false, false, false, false>(tag, g, epi, 1, st);
  if (K > 32) return launch_gemm<64, false, false, false, false>(tag, g, epi, 1, st);
  return launch_gemm<32, false, false, false, false>(tag, g, epi, 1, st);
}
static int linear_bwd_data(const char* tag, int64_t M, int N, int K, const float* dY, int64_t lddy, const float* W,
                           int64_t ldw, const float* addend, const float* mask, float* dX, int64_t lddx, cudaStream_t st) {
  if (addend && mask) return linear_bwd_data_t<true, true>(tag, M, N, K, dY, lddy, W, ldw, addend, mask, dX, lddx, st);
  if (addend) return linear_bwd_data_t<true, false>(tag, M, N, K, dY, lddy, W, ldw, addend, mask, dX, lddx, st);
  if (mask) return linear_bwd_data_t<false, true>(tag, M, N, K, dY, lddy, W, ldw, addend, mask, dX, lddx, st);
  return linear_bwd_data_t<false, false>(tag, M, N, K, dY, lddy, W, ldw, addend, mask, dX, lddx, st);
}

// dW[N,K] = dY[M,N]^T X[M,K] ; db[N] = colsum(dY)       (dW, db zeroed here)
static int linear_bwd_weight(const char* tag, int64_t M, int N, int K, const float* dY, int64_t lddy, const float* X, int64_t ldx,
                             float* dW, float* db, cudaStream_t st) {
  if (!g_prezeroed) {
    GG_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)N * K, st));
    if (db) GG_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * (size_t)N, st));
  }
  if (M <= 0) return GG_OK;
  GemmArgs g{};
  g.A = dY; g.lda = lddy; g.B = X; g.ldb = ldx; g.M = N; g.N = K; g.K = M;
  g.bias_grad = db;
  EpiAtomic epi{dW, (int64_t)K};
  const int64_t m_tiles = (N + kBM - 1) / kBM;
  if (g_tc_mode && tc::eligible(true, true, N, K, M, lddy, ldx, dY, X)) {
    const int64_t tiles = m_tiles * ((K + tc::BN - 1) / tc::BN);
    // persistent CTAs, one work item (tile x split) at a time: the items must fit ONE wave.  Rounding the split count
    // up (5 tiles x 30 splits = 150 items on 148 SMs) made two CTAs run a second item while 146 idled: floor.
    int64_t want = sm_count() / tiles;
    // at least 256 reduction rows (8 K-blocks) per split: a mini-batch sub-graph (20-40 k edges) then still spreads over
    // ~100 CTAs instead of ~25 (each K-block is a TMA round trip: 32 of them in a row were 33 us for a 23 k-edge batch)
    const int64_t max_by_k = (M + 255) / 256;
    if (want > max_by_k) want = max_by_k;
    if (want < 1) want = 1;
    return tc::launch<true, true, false, true>(tag, dY, lddy, X, ldx, N, K, M, (int)want, nullptr, db, epi, sm_count(), st);
  }
  if (K >= 128) return launch_gemm<128, true, false, false, true>(tag, g, epi, pick_splits(M, m_tiles * ((K + 127) / 128)), st);
  if (K > 32) return launch_gemm<64, true, false, false, true>(tag, g, epi, pick_splits(M, m_tiles), st);
  return launch_gemm<32, true, false, false, true>(tag, g, epi, pick_splits(M, m_tiles), st);
}

// ------------------------------------------------------------------ layer forward / backward
template <int D, int NORM>
static int layer_fwd_impl(const Plan* pl, int residual, const float* h_in, const float* e_in, const float* Wn,
                          const float* bn, const float* B3, const float* b3, const float* gamma_e,
                          const float* beta_e, const float* gamma_h, const float* beta_h, float* h_out,
                          float* e_out, float* P, float* t, float* z, float* agg, double* stats, cudaStream_t st) {
  const int64_t N = pl->N, E = pl->E;
  const int b = g_layer_parity;                   // zig-zag base direction of this layer (see zig())
  if (!g_prezeroed) GG_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 4 * D, st));
  // node projections P = h [A1|A2|A3|B1|B2]^T + b      (gated_gcn_full.py:107-112)
  int rc = linear_fwd("gemm_node_proj", N, 5 * D, D, h_in, D, Wn, D, bn, 0, P, 5 * D, st);
  if (rc) return rc;
  // t = B3 e + b3 + B1h[src] + B2h[dst], with batch statistics  (:113,120-121)
  {
    GemmArgs g{};
    g.A = e_in; g.lda = D; g.B = B3; g.ldb = D; g.M = E; g.N = D; g.K = D;
    g.col_stats = stats;
    EpiEdgeGate epi{t, D, b3, P, pl->src, pl->dst};
    constexpr int BN = D >= 128 ? 128 : 64;
    if (g_tc_mode && tc::eligible(false, false, E, D, D, D, D, e_in, B3))
      rc = tc::launch<false, false, NORM == GG_NORM_BATCH, false, EpiEdgeGate, tc::NoATx, GG_WRES(D)>(
          "gemm_edge_gate", e_in, D, B3, D, E, D, D, 1, stats, nullptr, epi, sm_count(), st, tc::NoATx{}, nullptr, zig(b));
    else
      rc = launch_gemm<BN, false, true, NORM == GG_NORM_BATCH, false>("gemm_edge_gate", g, epi, 1, st);
    if (rc) return rc;
  }
  if (E > 0 && (tc::tc_dbg_ref() & 8)) {
    // gg_debug_flags(8): streamed operands staged through shared memory by bulk copies (gg_layer_bulk.cuh).
    // Parity-green, but measured SLOWER than the register-staged kernel on the bench graph (154 vs 141 us at
    // d = 128: 16 consumer warps per SM instead of 24, every row crosses shared memory), so it is not the default.
    auto kern = edge_gate_fwd_bulk_kernel<D, NORM>;
    constexpr int kSmem = 6 * kBulkStageBytes;
    static int per_sm_dev[kMaxDevices] = {};
    int& per_sm = per_sm_dev[current_device()];
    if (per_sm == 0) {
      GG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
      int occ = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBulkThreads, kSmem) != cudaSuccess || occ < 1) occ = 1;
      per_sm = occ;
    }
    int64_t blocks = (N + kBulkNodes - 1) / kBulkNodes;
    if (blocks > (int64_t)sm_count() * per_sm) blocks = (int64_t)sm_count() * per_sm;
    GG_KERNEL_BEGIN("edge_gate_fwd_kernel", st);
    kern<<<(unsigned)blocks, kBulkThreads, kSmem, st>>>(N, E, pl->in_ptr, pl->src, t, e_in, P, stats, gamma_e, beta_e,
                                                        residual, e_out, agg);
    GG_KERNEL_END("edge_gate_fwd_kernel", st);
  } else {
    GG_KERNEL_BEGIN("edge_gate_fwd_kernel", st);
    edge_gate_fwd_kernel<D, NORM><<<node_grid(edge_gate_fwd_kernel<D, NORM>, N), kNodeThreads, 0, st>>>(
        N, E, pl->in_ptr, pl->src, t, e_in, P, stats, gamma_e, beta_e, residual, e_out, agg, zig(b ^ 1));
    GG_KERNEL_END("edge_gate_fwd_kernel", st);
  }
  GG_KERNEL_BEGIN("node_agg_fwd_kernel", st);
  node_agg_fwd_kernel<D, NORM><<<node_grid(node_agg_fwd_kernel<D, NORM>, N), kNodeThreads, 0, st>>>(
      N, pl->out_ptr, pl->out_eid, pl->out_dst, e_out, P, agg, z, stats + 2 * D, zig(b));
  GG_KERNEL_END("node_agg_fwd_kernel", st);
  GG_KERNEL_BEGIN("node_update_fwd_kernel", st);
  node_update_fwd_kernel<D, NORM><<<node_grid(node_update_fwd_kernel<D, NORM>, N), kNodeThreads, 0, st>>>(
      N, z, h_in, stats + 2 * D, gamma_h, beta_h, residual, h_out);
  GG_KERNEL_END("node_update_fwd_kernel", st);
  return GG_OK;
}

template <int D, int NORM>
static int layer_bwd_impl(const Plan* pl, int residual, const float* h_in, const float* e_in, const float* e_out,
                          const float* Wn, const float* B3, const float* gamma_e, const float* beta_e,
                          const float* gamma_h, const float* beta_h, const float* P, const float* t,
                          const float* z, const float* agg, const double* stats, const float* g_h,
                          const float* g_e, float* g_h_in, float* g_e_in, float* dWn, float* dbn, float* dB3,
                          float* db3, float* dgamma_e, float* dbeta_e, float* dgamma_h, float* dbeta_h,
                          float* gP, float* G, float* g_eo, float* g_t, double* bstats, cudaStream_t st) {
  const int64_t N = pl->N, E = pl->E;
  if (!g_prezeroed) GG_CUDA(cudaMemsetAsync(bstats, 0, sizeof(double) * 4 * D, st));
  GG_KERNEL_BEGIN("node_bwd_reduce_kernel", st);
  node_bwd_reduce_kernel<D, NORM><<<node_grid(node_bwd_reduce_kernel<D, NORM>, N), kNodeThreads, 0, st>>>(N, z, g_h, stats + 2 * D, gamma_h, beta_h, bstats);
  GG_KERNEL_END("node_bwd_reduce_kernel", st);
  GG_KERNEL_BEGIN("node_bwd_apply_kernel", st);
  node_bwd_apply_kernel<D, NORM><<<node_grid(node_bwd_apply_kernel<D, NORM>, N), kNodeThreads, 0, st>>>(N, z, g_h, stats + 2 * D, bstats, gamma_h, beta_h,
                                                               agg, gP, G, dgamma_h, dbeta_h);
  GG_KERNEL_END("node_bwd_apply_kernel", st);
  GG_KERNEL_BEGIN("edge_bwd_a_kernel", st);
  edge_bwd_a_kernel<D, NORM><<<node_grid(edge_bwd_a_kernel<D, NORM>, N), kNodeThreads, 0, st>>>(N, E, pl->in_ptr, pl->src, t, e_in, g_e, P, G, stats,
                                                           gamma_e, beta_e, residual, g_eo, gP, bstats + 2 * D, zig(0));
  GG_KERNEL_END("edge_bwd_a_kernel", st);
  int rc;
  // Batch norm + tensor-core path: g_t is produced inside the bwd-data GEMM (A-operand transform), so the
  // edge_bwd_b pass (read t, g_eo; write g_t) does not exist; gB2h is finished per node in edge_bwd_src.
  const bool fused_gt = (NORM == GG_NORM_BATCH) && g_tc_mode && E > 0 && D <= 256 && residual &&
                        g_e_in != g_eo &&      // the GEMM re-reads g_eo rows as its A operand: no in-place output
                        tc::eligible(false, true, E, D, D, D, D, g_eo, B3);
  if (fused_gt) {
    EpiAddMaskT<true, false> epi{g_e_in, (int64_t)D, g_eo, nullptr};
    tc::BnBwdATx atx{stats, bstats + 2 * D, gamma_e, beta_e, 1.0 / (double)E, g_t, (int64_t)D};
    rc = tc::launch<false, true, false, false, EpiAddMaskT<true, false>, tc::BnBwdATx, GG_WRES(D)>(
        "gemm_bwd_e_in", g_eo, D, B3, D, E, D, D, 1, nullptr, nullptr, epi, sm_count(), st, atx, t, zig(1));
    if (rc) return rc;
  } else {
    GG_KERNEL_BEGIN("edge_bwd_b_kernel", st);
    edge_bwd_b_kernel<D, NORM><<<node_grid(edge_bwd_b_kernel<D, NORM>, N), kNodeThreads, 0, st>>>(N, E, pl->in_ptr, t, g_eo, stats, bstats + 2 * D,
                                                             gamma_e, beta_e, g_t, gP);
    GG_KERNEL_END("edge_bwd_b_kernel", st);
  }
  GG_KERNEL_BEGIN("edge_bwd_src_kernel", st);
  edge_bwd_src_kernel<D><<<node_grid(edge_bwd_src_kernel<D>, N), kNodeThreads, 0, st>>>(N, pl->out_ptr, pl->out_eid, pl->out_dst, g_t, e_out, G, gP,
                                                       fused_gt ? 1 : 0, E, pl->in_ptr, agg + 4 * N * D, stats,
                                                       bstats + 2 * D, gamma_e, dgamma_e, dbeta_e, zig(0));
  GG_KERNEL_END("edge_bwd_src_kernel", st);
  // g_e_in = g_eo (residual) + g_t B3 ; dB3 = g_t^T e_in ; db3 = colsum g_t
  if (!fused_gt) {
    rc = linear_bwd_data("gemm_bwd_e_in", E, D, D, g_t, D, B3, D, residual ? g_eo : nullptr, nullptr, g_e_in, D, st);
    if (rc) return rc;
  }
  // weight gradients: on the side stream when the sequencer gave one (see SideCtx), else in line
  cudaStream_t wst = st;
  const SideCtx sc = g_side;
  if (sc.side != nullptr) {
    GG_CUDA(cudaEventRecord(sc.fork, st));
    GG_CUDA(cudaStreamWaitEvent(sc.side, sc.fork, 0));
    wst = sc.side;
  }
  rc = linear_bwd_weight("gemm_dB3", E, D, D, g_t, D, e_in, D, dB3, db3, wst);
  if (rc) return rc;
  rc = linear_bwd_weight("gemm_dWn", N, 5 * D, D, gP, 5 * D, h_in, D, dWn, dbn, wst);
  if (rc) return rc;
  if (sc.side != nullptr) GG_CUDA(cudaEventRecord(sc.done, sc.side));
  // g_h_in = g_h (residual) + gP Wn
  rc = linear_bwd_data("gemm_bwd_h_in", N, 5 * D, D, gP, 5 * D, Wn, D, residual ? g_h : nullptr, nullptr, g_h_in, D, st);
  if (rc) return rc;
  // (dgamma / dbeta of both norms are written by node_bwd_apply_kernel and edge_bwd_src_kernel: no launches of their own)
  return GG_OK;
}

// ------------------------------------------------------------------ predictor backward helpers
// g_pre = g_score * w2 * [hid > 0] (in place over hid); red = [dw2 (H) | db1 (H) | db2 (1)] in fp64
__global__ void __launch_bounds__(kNodeThreads)
score_bwd_pre_kernel(int64_t E, const float* __restrict__ g_score, const float* __restrict__ w2,
                     const float* hid, float* g_pre, double* __restrict__ red) {
  constexpr int H = 64, VPL = 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t gw = ((int64_t)blockIdx.x * kNodeThreads + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * kNodeThreads) >> 5;
  Row<H> w;
  w.load(w2, lane);
  double s1[VPL] = {0.0, 0.0}, s2[VPL] = {0.0, 0.0}, sg = 0.0;
  for (int64_t i = gw; i < E; i += nw) {
    const float gs = __ldg(g_score + i);
    Row<H> h;
    h.load_stream(hid + i * H, lane);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      s1[k] += (double)gs * (double)h.v[k];
      const float gp = h.v[k] > 0.f ? gs * w.v[k] : 0.f;
      s2[k] += (double)gp;
      h.v[k] = gp;
    }
    h.store(g_pre + i * H, lane);
    sg += (double)gs;
  }
  block_flush_stats<H>(s1, s2, red, red + H);
  __shared__ double shg[kNodeThreads / 32];
  if (lane == 0) shg[warp] = sg;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int wi = 0; wi < kNodeThreads / 32; ++wi) a += shg[wi];
    atomicAdd(red + 2 * H, a);
  }
}

__global__ void score_small_grads_kernel(int H, const double* __restrict__ red, float* __restrict__ dw2,
                                         float* __restrict__ dbq, float* __restrict__ db2) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < H) {
    dw2[c] = (float)red[c];
    dbq[c] = (float)red[H + c];
    dbq[H + c] = 0.f;
  }
  if (c == 0) db2[0] = (float)red[2 * H];
}

__global__ void gather_rows_kernel(int64_t rows, int width, const float* __restrict__ in,
                                   const int32_t* __restrict__ idx, float* __restrict__ out) {
  const int64_t total = rows * width;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = k / width;
    const int c = (int)(k - r * width);
    out[k] = __ldg(in + (int64_t)__ldg(idx + r) * width + c);
  }
}

}  // namespace gg

using namespace gg;

#define GG_DISPATCH(d, norm, CALL)                                                  \
  do {                                                                              \
    if (norm == GG_NORM_BATCH) {                                                    \
      if (d == 64) return CALL(64, GG_NORM_BATCH);                                  \
      if (d == 128) return CALL(128, GG_NORM_BATCH);                                \
      if (d == 256) return CALL(256, GG_NORM_BATCH);                                \
    } else if (norm == GG_NORM_LAYER) {                                             \
      if (d == 64) return CALL(64, GG_NORM_LAYER);                                  \
      if (d == 128) return CALL(128, GG_NORM_LAYER);                                \
      if (d == 256) return CALL(256, GG_NORM_LAYER);                                \
    }                                                                               \
    set_error("gnnome_b200: unsupported hidden size / norm kind (d in {64,128,256})"); \
    return GG_ERR_UNSUPPORTED;                                                      \
  } while (0)

extern "C" {

int gg_debug_flags(int flags) { const int old = tc::tc_dbg_ref(); tc::tc_dbg_ref() = flags; return old; }

int gg_debug_trace(unsigned long long* out, int n, int reset) {
  GG_REQUIRE(out != nullptr && n >= 0, "debug_trace: bad arguments");
  unsigned long long host[tc::TR_SLOTS] = {};
  GG_CUDA(cudaDeviceSynchronize());
  GG_CUDA(cudaMemcpyFromSymbol(host, tc::gg_tc_trace, sizeof host));
  for (int i = 0; i < n; ++i) out[i] = i < (int)tc::TR_SLOTS ? host[i] : 0ull;
  if (reset) {
    unsigned long long zero[tc::TR_SLOTS] = {};
    GG_CUDA(cudaMemcpyToSymbol(tc::gg_tc_trace, zero, sizeof zero));
  }
  return GG_OK;
}

int gg_set_tc_mode(int mode) {
  const int old = g_tc_mode;
  g_tc_mode = mode ? 1 : 0;
  return old;
}

int gg_linear_fwd(int64_t M, int N, int K, const float* X, const float* W, const float* b, int relu, float* Y,
                  void* stream) {
  GG_REQUIRE(M >= 0 && N > 0 && K > 0, "linear_fwd: bad sizes");
  GG_REQUIRE(N % 4 == 0 && K % 4 == 0, "linear_fwd: N and K must be multiples of 4");
  if (M == 0) return GG_OK;
  GG_REQUIRE(X && W && Y, "linear_fwd: null pointer");
  return linear_fwd("gemm_linear_fwd", M, N, K, X, K, W, K, b, relu, Y, N, (cudaStream_t)stream);
}

int gg_linear_bwd_data(int64_t M, int N, int K, const float* dY, const float* W, const float* addend,
                       const float* relu_mask, float* dX, void* stream) {
  GG_REQUIRE(M >= 0 && N > 0 && K > 0, "linear_bwd_data: bad sizes");
  GG_REQUIRE(N % 4 == 0 && K % 4 == 0, "linear_bwd_data: N and K must be multiples of 4");
  if (M == 0) return GG_OK;
  GG_REQUIRE(dY && W && dX, "linear_bwd_data: null pointer");
  return linear_bwd_data("gemm_linear_bwd_data", M, N, K, dY, N, W, K, addend, relu_mask, dX, K, (cudaStream_t)stream);
}

int gg_linear_bwd_weight(int64_t M, int N, int K, const float* dY, const float* X, float* dW, float* db,
                         void* stream) {
  GG_REQUIRE(M >= 0 && N > 0 && K > 0, "linear_bwd_weight: bad sizes");
  GG_REQUIRE(N % 4 == 0 && K % 4 == 0, "linear_bwd_weight: N and K must be multiples of 4");
  GG_REQUIRE(dW && (M == 0 || (dY && X)), "linear_bwd_weight: null pointer");
  return linear_bwd_weight("gemm_linear_bwd_weight", M, N, K, dY, N, X, K, dW, db, (cudaStream_t)stream);
}

int gg_layer_fwd(const gg_plan_t* plan, int d, int norm_kind, int residual, const float* h_in, const float* e_in,
                 const float* Wn, const float* bn, const float* B3, const float* b3, const float* gamma_e,
                 const float* beta_e, const float* gamma_h, const float* beta_h, float* h_out, float* e_out,
                 float* P, float* t, float* z, float* agg, double* stats, void* stream) {
  GG_REQUIRE(plan, "layer_fwd: null plan");
  const Plan* pl = reinterpret_cast<const Plan*>(plan);
  GG_REQUIRE(Wn && bn && B3 && b3 && gamma_e && beta_e && gamma_h && beta_h && stats, "layer_fwd: null parameter");
  GG_REQUIRE(pl->N == 0 || (h_in && h_out && P && z && agg), "layer_fwd: null node buffer");
  GG_REQUIRE(pl->E == 0 || (e_in && e_out && t), "layer_fwd: null edge buffer");
#define CALL_FWD(DD, NN)                                                                                  \
  layer_fwd_impl<DD, NN>(pl, residual, h_in, e_in, Wn, bn, B3, b3, gamma_e, beta_e, gamma_h, beta_h, h_out, \
                         e_out, P, t, z, agg, stats, (cudaStream_t)stream)
  GG_DISPATCH(d, norm_kind, CALL_FWD);
#undef CALL_FWD
}

int gg_layer_bwd(const gg_plan_t* plan, int d, int norm_kind, int residual, const float* h_in, const float* e_in,
                 const float* e_out, const float* Wn, const float* B3, const float* gamma_e, const float* beta_e,
                 const float* gamma_h, const float* beta_h, const float* P, const float* t, const float* z,
                 const float* agg, const double* stats, const float* g_h, const float* g_e, float* g_h_in,
                 float* g_e_in, float* dWn, float* dbn, float* dB3, float* db3, float* dgamma_e, float* dbeta_e,
                 float* dgamma_h, float* dbeta_h, float* gP, float* G, float* g_eo, float* g_t, double* bstats,
                 void* stream) {
  GG_REQUIRE(plan, "layer_bwd: null plan");
  const Plan* pl = reinterpret_cast<const Plan*>(plan);
  GG_REQUIRE(Wn && B3 && gamma_e && beta_e && gamma_h && beta_h && stats && bstats, "layer_bwd: null parameter");
  GG_REQUIRE(dWn && dbn && dB3 && db3 && dgamma_e && dbeta_e && dgamma_h && dbeta_h, "layer_bwd: null output");
  GG_REQUIRE(pl->N == 0 || (h_in && P && z && agg && g_h_in && gP && G), "layer_bwd: null node buffer");
  GG_REQUIRE(pl->N == 0 || g_h, "layer_bwd: g_h must be given (pass zeros when only e_out has a gradient)");
  GG_REQUIRE(pl->E == 0 || (e_in && e_out && t && g_e_in && g_eo && g_t), "layer_bwd: null edge buffer");
#define CALL_BWD(DD, NN)                                                                                      \
  layer_bwd_impl<DD, NN>(pl, residual, h_in, e_in, e_out, Wn, B3, gamma_e, beta_e, gamma_h, beta_h, P, t, z, agg, \
                         stats, g_h, g_e, g_h_in, g_e_in, dWn, dbn, dB3, db3, dgamma_e, dbeta_e, dgamma_h,   \
                         dbeta_h, gP, G, g_eo, g_t, bstats, (cudaStream_t)stream)
  GG_DISPATCH(d, norm_kind, CALL_BWD);
#undef CALL_BWD
}

int gg_score_fwd(const gg_plan_t* plan, int d, int H, const float* x, const float* e, const float* Wq,
                 const float* bq, const float* W1e, const float* w2, const float* b2, float* score, float* Q,
                 float* hid, void* stream) {
  GG_REQUIRE(plan, "score_fwd: null plan");
  const Plan* pl = reinterpret_cast<const Plan*>(plan);
  GG_REQUIRE(Wq && bq && W1e && w2 && b2, "score_fwd: null parameter");
  GG_REQUIRE(pl->N == 0 || (x && Q), "score_fwd: null node buffer");
  GG_REQUIRE(pl->E == 0 || (e && score), "score_fwd: null edge buffer");
  if (H != 64 || !(d == 64 || d == 128 || d == 256)) {
    set_error("gnnome_b200: score predictor supports H = 64 and d in {64,128,256}");
    return GG_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int rc = linear_fwd("gemm_score_q", pl->N, 2 * H, d, x, d, Wq, d, bq, 0, Q, 2 * H, st);
  if (rc) return rc;
  if (pl->E == 0) return GG_OK;
  if (g_tc_mode && d % tc::BK == 0 && reinterpret_cast<uintptr_t>(e) % 16 == 0 && reinterpret_cast<uintptr_t>(W1e) % 16 == 0) {
    EpiScoreTC epi{score, hid, Q, w2, b2, pl->src, pl->dst};
    return tc::launch<false, false, false, false>("gemm_score", e, d, W1e, d, pl->E, H, d, 1, nullptr, nullptr, epi,
                                                  sm_count(), st);
  }
  GemmArgs g{};
  g.A = e; g.lda = d; g.B = W1e; g.ldb = d; g.M = pl->E; g.N = H; g.K = d;
  EpiScore epi{score, hid, Q, w2, b2, pl->src, pl->dst, pl->E};
  return launch_gemm<64, false, true, false, false>("gemm_score", g, epi, 1, st);
}

int gg_score_bwd(const gg_plan_t* plan, int d, int H, const float* x, const float* e, const float* Wq,
                 const float* W1e, const float* w2, const float* g_score, const float* hid, float* g_x,
                 float* g_e, float* dWq, float* dbq, float* dW1e, float* dw2, float* db2, float* g_pre,
                 float* gQ, double* red, void* stream) {
  GG_REQUIRE(plan, "score_bwd: null plan");
  const Plan* pl = reinterpret_cast<const Plan*>(plan);
  GG_REQUIRE(Wq && W1e && w2 && dWq && dbq && dW1e && dw2 && db2 && red, "score_bwd: null parameter / output");
  GG_REQUIRE(pl->N == 0 || (x && g_x && gQ), "score_bwd: null node buffer");
  GG_REQUIRE(pl->E == 0 || (e && g_score && hid && g_e && g_pre), "score_bwd: null edge buffer");
  if (H != 64 || !(d == 64 || d == 128 || d == 256)) {
    set_error("gnnome_b200: score predictor supports H = 64 and d in {64,128,256}");
    return GG_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t N = pl->N, E = pl->E;
  if (!g_prezeroed) GG_CUDA(cudaMemsetAsync(red, 0, sizeof(double) * (2 * H + 1), st));
  {
    int64_t blocks = (E + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    GG_KERNEL_BEGIN("score_bwd_pre_kernel", st);
    score_bwd_pre_kernel<<<(unsigned)blocks, kNodeThreads, 0, st>>>(E, g_score, w2, hid, g_pre, red);
    GG_KERNEL_END("score_bwd_pre_kernel", st);
  }
  GG_KERNEL_BEGIN("score_small_grads_kernel", st);
  score_small_grads_kernel<<<1, 64, 0, st>>>(H, red, dw2, dbq, db2);
  GG_KERNEL_END("score_small_grads_kernel", st);
  int rc = linear_bwd_data("gemm_score_bwd_e", E, H, d, g_pre, H, W1e, d, nullptr, nullptr, g_e, d, st);     // g_e = g_pre W1e
  if (rc) return rc;
  rc = linear_bwd_weight("gemm_score_dW1e", E, H, d, g_pre, H, e, d, dW1e, nullptr, st);                      // dW1e = g_pre^T e
  if (rc) return rc;
  GG_KERNEL_BEGIN("edge_to_node_sums_kernel", st);
  edge_to_node_sums_kernel<64><<<node_grid(edge_to_node_sums_kernel<64>, N), kNodeThreads, 0, st>>>(N, pl->in_ptr, pl->out_ptr, pl->out_eid, g_pre, gQ);
  GG_KERNEL_END("edge_to_node_sums_kernel", st);
  rc = linear_bwd_data("gemm_score_bwd_x", N, 2 * H, d, gQ, 2 * H, Wq, d, nullptr, nullptr, g_x, d, st);    // g_x = gQ Wq
  if (rc) return rc;
  return linear_bwd_weight("gemm_score_dWq", N, 2 * H, d, gQ, 2 * H, x, d, dWq, nullptr, st);              // dWq = gQ^T x
}

int gg_gather_rows(int64_t rows, int width, const float* in, const int32_t* idx, float* out, void* stream) {
  GG_REQUIRE(rows >= 0 && width > 0, "gather_rows: bad sizes");
  if (rows == 0) return GG_OK;
  GG_REQUIRE(in && idx && out, "gather_rows: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = (rows * width + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  GG_KERNEL_BEGIN("gather_rows_kernel", st);
  gather_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(rows, width, in, idx, out);
  GG_KERNEL_END("gather_rows_kernel", st);
  return GG_OK;
}

}  // extern "C"
