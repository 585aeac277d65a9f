// gg_model.cu — the whole GraphGatedGCNModel forward / backward as ONE C-ABI call each.
//
// Replaces models/full_graph.py:22-29 (linear_pe, edge MLP, L x GatedGCN_1d, ScorePredictor) and its autograd as a
// host-side sequencer over the entry points of include/gnnome_b200.h: no arithmetic of its own beyond row padding /
// weight re-packing kernels.  Why it exists: through the per-op bindings one training step is ~300 Python-level calls
// (ctypes + allocation + autograd node each); on the reference's DEFAULT path — cluster mini-batches of ~25-40 k edges
// (train.py:282-312, hyperparameters.py:15-18) — the step is bound by that host overhead (4.9 ms per 23 k-edge batch in
// round 1, 7.5x worse per edge than the full graph).  Here the ~120 launches of a step are issued back to back from
// C++; the caller passes the FLAT parameter buffer (flat.py) with an offset table, one workspace, and gets gradients
// written into a flat arena with the same offsets.
//
// Offset table (int64, in floats, into `params` / `grads`), n = 10 + 8 L entries:
//   0 linear_pe.weight [d, node_in]   1 linear_pe.bias [d]
//   2 linear1_edge.weight [he, edge_in]   3 linear1_edge.bias [he]   4 linear2_edge.weight [d, he]   5 linear2_edge.bias [d]
//   6 predictor.W1.weight [H, 3d]   7 predictor.W1.bias [H]   8 predictor.W2.weight [1, H]   9 predictor.W2.bias [1]
//   10 + 8 l + {0: Wn [5d, d] (A_1,A_2,A_3,B_1,B_2 stacked), 1: bn [5d], 2: B_3.weight, 3: B_3.bias,
//               4: bn_e.weight, 5: bn_e.bias, 6: bn_h.weight, 7: bn_h.bias}
#include <vector>

#include "gg_common.cuh"

namespace gg {

constexpr int64_t kAlignFloats = 64;       // 256-byte alignment of every workspace slice (TMA / float4)

struct Bump {
  int64_t off = 0;
  int64_t take(int64_t n) {
    const int64_t o = off;
    off += (n + kAlignFloats - 1) / kAlignFloats * kAlignFloats;
    return o;
  }
};

struct ModelDims {
  int64_t N, E;
  int d, L, he, H, node_in, edge_in, node_k4, edge_k4;
};

// forward workspace (training: everything the backward reads stays; inference: layer buffers ping-pong)
struct FwdLayout {
  int64_t e4, pe4, Wpe4, W1e4, hid_e, Wq, bq, W1e, Q, hid_s, score_int;
  std::vector<int64_t> h, e, P, t, z, agg, stats;      // h / e: L + 1 entries (layer inputs, last = outputs)
  int64_t stats_all, stats_floats;                     // the per-layer statistics as one block
  int64_t total;
};

static FwdLayout fwd_layout(const ModelDims& m, bool training) {
  FwdLayout w;
  Bump b;
  const int64_t N = m.N > 0 ? m.N : 1, E = m.E > 0 ? m.E : 1, d = m.d;
  w.e4 = b.take(E * m.edge_k4);
  w.pe4 = b.take(N * m.node_k4);
  w.Wpe4 = b.take((int64_t)d * m.node_k4);
  w.W1e4 = b.take((int64_t)m.he * m.edge_k4);
  w.hid_e = b.take(E * m.he);
  w.Wq = b.take(2LL * m.H * d);
  w.bq = b.take(2LL * m.H);
  w.W1e = b.take((int64_t)m.H * d);
  w.Q = b.take(N * 2 * m.H);
  w.hid_s = training ? b.take(E * m.H) : -1;
  w.score_int = b.take(E);
  const int slots = training ? m.L + 1 : 2;
  std::vector<int64_t> hs(slots), es(slots);
  for (int i = 0; i < slots; ++i) { hs[i] = b.take(N * d); es[i] = b.take(E * d); }
  const int sets = training ? m.L : 1;
  std::vector<int64_t> P(sets), t(sets), z(sets), agg(sets);
  for (int i = 0; i < sets; ++i) {
    P[i] = b.take(N * 5 * d); t[i] = b.take(E * d); z[i] = b.take(N * d); agg[i] = b.take(5 * N * d);
  }
  // batch statistics: 4d doubles per layer, all layers contiguous (one memset per pass zeroes them)
  w.stats_all = b.take(2LL * 4 * d * (m.L > 0 ? m.L : 1));
  w.stats_floats = 2LL * 4 * d * m.L;
  for (int l = 0; l <= m.L; ++l) { w.h.push_back(hs[training ? l : (l & 1)]); w.e.push_back(es[training ? l : (l & 1)]); }
  for (int l = 0; l < m.L; ++l) {
    const int s = training ? l : 0;
    w.P.push_back(P[s]); w.t.push_back(t[s]); w.z.push_back(z[s]); w.agg.push_back(agg[s]);
    w.stats.push_back(w.stats_all + 2LL * 4 * d * l);
  }
  w.total = b.off;
  return w;
}

struct BwdLayout {
  int64_t g_int, g_h[2], g_e[2], gP[2], G, g_eo, g_t[2], bstats, g_pre, gQ, red, dWq, dbq, dW1e, dWpe4, dW1e4, g_hid;
  int64_t zero_begin, zero_end;                        // accumulators (bstats per layer .. dW1e4): one memset per pass
  int64_t total;
};

static BwdLayout bwd_layout(const ModelDims& m) {
  BwdLayout w;
  Bump b;
  const int64_t N = m.N > 0 ? m.N : 1, E = m.E > 0 ? m.E : 1, d = m.d;
  w.g_int = b.take(E);
  for (int i = 0; i < 2; ++i) { w.g_h[i] = b.take(N * d); w.g_e[i] = b.take(E * d); }
  // gP and g_t are read by the weight-gradient GEMMs on the side stream while the next layer overwrites them: two sets
  for (int i = 0; i < 2; ++i) { w.gP[i] = b.take(N * 5 * d); w.g_t[i] = b.take(E * d); }
  w.G = b.take(4 * N * d);
  w.g_eo = b.take(E * d);
  w.g_pre = b.take(E * m.H);
  w.gQ = b.take(N * 2 * m.H);
  w.zero_begin = b.off;
  w.bstats = b.take(2LL * 4 * d * (m.L > 0 ? m.L : 1));       // 4d doubles per layer
  w.red = b.take(2LL * (2 * m.H + 1));
  w.dWq = b.take(2LL * m.H * d);
  w.dbq = b.take(2LL * m.H);
  w.dW1e = b.take((int64_t)m.H * d);
  w.dWpe4 = b.take((int64_t)d * m.node_k4);
  w.dW1e4 = b.take((int64_t)m.he * m.edge_k4);
  w.zero_end = b.off;
  w.g_hid = b.take(E * m.he);
  w.total = b.off;
  return w;
}

// out[r, 0:w_out] = in[idx ? idx[r] : r, 0:w_in] zero-padded (w_out >= w_in) or truncated (w_out < w_in)
__global__ void copy_rows_kernel(int64_t rows, int w_in, int w_out, const float* __restrict__ in,
                                 const int32_t* __restrict__ idx, float* __restrict__ out) {
  const int64_t total = rows * w_out;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = k / w_out;
    const int c = (int)(k - r * w_out);
    const int64_t s = idx ? (int64_t)__ldg(idx + r) : r;
    out[k] = c < w_in ? __ldg(in + s * w_in + c) : 0.f;
  }
}

// predictor weights: W1 [H, 3d] = [W1s | W1d | W1e]  ->  Wq [2H, d] = [W1s ; W1d], bq [2H] = [b1 ; 0], W1e [H, d]
__global__ void score_split_kernel(int H, int d, const float* __restrict__ W1, const float* __restrict__ b1,
                                   float* __restrict__ Wq, float* __restrict__ bq, float* __restrict__ W1e) {
  const int total = H * 3 * d;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
    const int r = k / (3 * d), c = k - r * 3 * d;
    const float v = __ldg(W1 + k);
    if (c < d) Wq[r * d + c] = v;
    else if (c < 2 * d) Wq[(H + r) * d + (c - d)] = v;
    else W1e[r * d + (c - 2 * d)] = v;
  }
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < 2 * H; k += gridDim.x * blockDim.x)
    bq[k] = k < H ? __ldg(b1 + k) : 0.f;
}

// the inverse for the gradients: dW1 [H, 3d] <- dWq [2H, d], dW1e [H, d];  db1 [H] <- dbq[0:H]
__global__ void score_merge_kernel(int H, int d, const float* __restrict__ dWq, const float* __restrict__ dbq,
                                   const float* __restrict__ dW1e, float* __restrict__ dW1, float* __restrict__ db1) {
  const int total = H * 3 * d;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
    const int r = k / (3 * d), c = k - r * 3 * d;
    dW1[k] = c < d ? __ldg(dWq + r * d + c) : (c < 2 * d ? __ldg(dWq + (H + r) * d + (c - d)) : __ldg(dW1e + r * d + (c - 2 * d)));
  }
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < H; k += gridDim.x * blockDim.x) db1[k] = __ldg(dbq + k);
}

static int copy_rows(const char* tag, int64_t rows, int w_in, int w_out, const float* in, const int32_t* idx, float* out,
                     cudaStream_t st) {
  if (rows <= 0) return GG_OK;
  int64_t blocks = (rows * w_out + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  GG_KERNEL_BEGIN(tag, st);
  copy_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(rows, w_in, w_out, in, idx, out);
  GG_KERNEL_END(tag, st);
  return GG_OK;
}

static int check_desc(const gg_plan_t* plan, const gg_model_desc_t* m, const int64_t* offs, int n_offs, ModelDims* out) {
  GG_REQUIRE(plan && m && offs, "model: null plan / descriptor / offset table");
  GG_REQUIRE(m->layers >= 0 && n_offs == 10 + 8 * m->layers, "model: offset table must hold 10 + 8 L entries");
  if (!(m->d == 64 || m->d == 128 || m->d == 256) || m->hidden_score != 64) {
    set_error("gnnome_b200: model kernels are built for d in {64,128,256}, hidden_edge_scores = 64");
    return GG_ERR_UNSUPPORTED;
  }
  GG_REQUIRE(m->hidden_edge > 0 && m->hidden_edge % 4 == 0, "model: hidden_edge_features must be a multiple of 4");
  GG_REQUIRE(m->node_in > 0 && m->edge_in > 0, "model: bad input widths");
  for (int i = 0; i < n_offs; ++i)      // float4 / TMA accesses: every tensor starts on a 16-byte boundary of the buffer
    GG_REQUIRE(offs[i] >= 0 && offs[i] % 4 == 0, "model: every offset must be a non-negative multiple of 4 floats");
  const Plan* pl = reinterpret_cast<const Plan*>(plan);
  out->N = pl->N; out->E = pl->E;
  out->d = m->d; out->L = m->layers; out->he = m->hidden_edge; out->H = m->hidden_score;
  out->node_in = m->node_in; out->edge_in = m->edge_in;
  out->node_k4 = (m->node_in + 3) & ~3; out->edge_k4 = (m->edge_in + 3) & ~3;
  return GG_OK;
}

// size (floats) of parameter i of the offset table (include/gnnome_b200.h: gg_model_fwd)
static int64_t param_floats(const ModelDims& m, int i) {
  const int64_t d = m.d;
  switch (i) {
    case 0: return d * m.node_in;
    case 1: return d;
    case 2: return (int64_t)m.he * m.edge_in;
    case 3: return m.he;
    case 4: return d * m.he;
    case 5: return d;
    case 6: return (int64_t)m.H * 3 * d;
    case 7: return m.H;
    case 8: return m.H;
    case 9: return 1;
    default: break;
  }
  switch ((i - 10) % 8) {
    case 0: return 5 * d * d;
    case 1: return 5 * d;
    case 2: return d * d;
    default: return d;
  }
}

// The gradient arena mirrors the parameter buffer.  When the table describes one span whose only gaps are alignment
// padding (< 4 floats after a tensor: FlatLayout's always does) the whole span is zeroed with one memset at the start of
// a pass; otherwise every op zeroes its own outputs as before.
static bool arena_span(const ModelDims& m, const int64_t* offs, int n_offs, int64_t* lo, int64_t* floats) {
  int64_t mn = INT64_MAX, mx = 0, sum = 0;
  for (int i = 0; i < n_offs; ++i) {
    const int64_t n = param_floats(m, i);
    if (offs[i] < mn) mn = offs[i];
    if (offs[i] + n > mx) mx = offs[i] + n;
    sum += (n + 3) & ~3LL;
  }
  *lo = mn; *floats = mx - mn;
  return n_offs > 0 && mx - mn <= sum && mx - mn > sum - 4;
}

}  // namespace gg

using namespace gg;

#define GG_TRY(call)          \
  do {                        \
    int _rc = (call);         \
    if (_rc) return _rc;      \
  } while (0)

extern "C" {

int64_t gg_model_workspace_floats(const gg_plan_t* plan, const gg_model_desc_t* m, int which) {
  if (!plan || !m) return -1;
  ModelDims dm;
  std::vector<int64_t> dummy(10 + 8 * (m->layers > 0 ? m->layers : 0), 0);
  if (check_desc(plan, m, dummy.data(), (int)dummy.size(), &dm)) return -1;
  if (which == 0) return fwd_layout(dm, true).total;
  if (which == 1) return fwd_layout(dm, false).total;
  if (which == 2) return bwd_layout(dm).total;
  return -1;
}

int gg_model_fwd(const gg_plan_t* plan, const gg_model_desc_t* m, const float* params, const int64_t* offs, int n_offs,
                 const float* e, const float* pe, int training, float* ws, float* scores, void* stream) {
  ModelDims dm;
  GG_TRY(check_desc(plan, m, offs, n_offs, &dm));
  GG_REQUIRE(params && ws, "model_fwd: null parameter buffer / workspace");
  GG_REQUIRE(reinterpret_cast<uintptr_t>(params) % 16 == 0 && reinterpret_cast<uintptr_t>(ws) % 16 == 0,
             "model_fwd: parameter buffer / workspace must be 16-byte aligned");
  GG_REQUIRE(dm.E == 0 || (e && scores), "model_fwd: null edge input / output");
  GG_REQUIRE(dm.N == 0 || pe, "model_fwd: null node input");
  const Plan* pl = reinterpret_cast<const Plan*>(plan);
  cudaStream_t st = (cudaStream_t)stream;
  const FwdLayout w = fwd_layout(dm, training != 0);
  const int d = dm.d;
  auto P = [&](int i) { return params + offs[i]; };
  // inputs: caller edge / node order -> internal order, K padded to a multiple of 4          (full_graph.py:22)
  GG_TRY(copy_rows("gather_rows_kernel", dm.E, dm.edge_in, dm.edge_k4, e, pl->perm, ws + w.e4, st));
  GG_TRY(copy_rows("gather_rows_kernel", dm.N, dm.node_in, dm.node_k4, pe, pl->node_perm, ws + w.pe4, st));
  GG_TRY(copy_rows("pad_weight_kernel", d, dm.node_in, dm.node_k4, P(0), nullptr, ws + w.Wpe4, st));
  GG_TRY(copy_rows("pad_weight_kernel", dm.he, dm.edge_in, dm.edge_k4, P(2), nullptr, ws + w.W1e4, st));
  // encoders                                                                               (full_graph.py:23-26)
  GG_TRY(gg_linear_fwd(dm.N, d, dm.node_k4, ws + w.pe4, ws + w.Wpe4, P(1), 0, ws + w.h[0], stream));
  {
    int rc = gg_edge_mlp_fwd(dm.E, d, dm.he, dm.edge_k4, ws + w.e4, ws + w.W1e4, P(3), P(4), P(5), ws + w.hid_e, ws + w.e[0],
                             stream);
    if (rc == GG_ERR_UNSUPPORTED) {                         // other widths: the two layers as two GEMM calls
      GG_TRY(gg_linear_fwd(dm.E, dm.he, dm.edge_k4, ws + w.e4, ws + w.W1e4, P(3), 1, ws + w.hid_e, stream));
      GG_TRY(gg_linear_fwd(dm.E, d, dm.he, ws + w.hid_e, P(4), P(5), 0, ws + w.e[0], stream));
    } else if (rc) {
      return rc;
    }
  }
  // L x GatedGCN                                                                           (processor.py:15-20)
  if (w.stats_floats > 0) GG_CUDA(cudaMemsetAsync(ws + w.stats_all, 0, sizeof(float) * w.stats_floats, st));
  struct Prezeroed { Prezeroed() { set_prezeroed(true); } ~Prezeroed() { set_prezeroed(false); } };
  for (int l = 0; l < dm.L; ++l) {
    const int o = 10 + 8 * l;
    set_layer_parity(l);                                  // zig-zag traversal: layer l starts where layer l-1 stopped
    Prezeroed hint;
    GG_TRY(gg_layer_fwd(plan, d, m->norm_kind, 1, ws + w.h[l], ws + w.e[l], P(o), P(o + 1), P(o + 2), P(o + 3), P(o + 4),
                        P(o + 5), P(o + 6), P(o + 7), ws + w.h[l + 1], ws + w.e[l + 1], ws + w.P[l], ws + w.t[l],
                        ws + w.z[l], ws + w.agg[l], reinterpret_cast<double*>(ws + w.stats[l]), stream));
  }
  set_layer_parity(0);
  // predictor                                                                              (score_predictor.py:12-25)
  GG_KERNEL_BEGIN("score_split_kernel", st);
  score_split_kernel<<<32, 256, 0, st>>>(dm.H, d, P(6), P(7), ws + w.Wq, ws + w.bq, ws + w.W1e);
  GG_KERNEL_END("score_split_kernel", st);
  GG_TRY(gg_score_fwd(plan, d, dm.H, ws + w.h[dm.L], ws + w.e[dm.L], ws + w.Wq, ws + w.bq, ws + w.W1e, P(8), P(9),
                      ws + w.score_int, ws + w.Q, w.hid_s >= 0 ? ws + w.hid_s : nullptr, stream));
  // internal -> caller edge order
  GG_TRY(gg_gather_rows(dm.E, 1, ws + w.score_int, pl->inv_perm, scores, stream));
  return GG_OK;
}

// events of the side-stream protocol: created once per host thread, reused (recording re-arms them)
struct SideEvents {
  cudaEvent_t fork = nullptr, done[2] = {nullptr, nullptr};
  bool ok = false;
  bool pending[2] = {false, false};      // side GEMMs reading buffer set k are in flight (persists across phase calls)
  bool init() {
    if (ok) return true;
    ok = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming) == cudaSuccess;
    return ok;
  }
};
static thread_local SideEvents g_side_events[kMaxDevices];

int gg_model_bwd(const gg_plan_t* plan, const gg_model_desc_t* m, const float* params, const int64_t* offs, int n_offs,
                 const float* g_scores, const float* ws, float* bws, float* grads, int phase_begin, int phase_end,
                 void* stream, void* side_stream) {
  ModelDims dm;
  GG_TRY(check_desc(plan, m, offs, n_offs, &dm));
  GG_REQUIRE(params && ws && bws && grads, "model_bwd: null buffer");
  GG_REQUIRE(reinterpret_cast<uintptr_t>(params) % 16 == 0 && reinterpret_cast<uintptr_t>(grads) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(ws) % 16 == 0 && reinterpret_cast<uintptr_t>(bws) % 16 == 0,
             "model_bwd: parameter / gradient buffers and workspaces must be 16-byte aligned");
  GG_REQUIRE(dm.E == 0 || g_scores, "model_bwd: null upstream gradient");
  const Plan* pl = reinterpret_cast<const Plan*>(plan);
  cudaStream_t st = (cudaStream_t)stream;
  const FwdLayout w = fwd_layout(dm, true);
  const BwdLayout b = bwd_layout(dm);
  const int d = dm.d, L = dm.L;
  auto P = [&](int i) { return params + offs[i]; };
  auto Gr = [&](int i) { return grads + offs[i]; };
  // phases: 0 = predictor, 1 + k = layer L-1-k, L + 1 = encoders.  The gradient of the layer stack's output alternates
  // between the two g_h / g_e buffers; layer l reads buffer (L - l) & 1... and writes the other one.
  if (phase_end > L + 2) phase_end = L + 2;
  // side stream for the weight-gradient GEMMs.  A caller that goes phase by phase (data-parallel all-reduce per layer
  // bucket) must make its collective wait for `side_stream` as well as `stream`; the join into `stream` happens in the
  // call that runs the last phase.
  cudaStream_t side = (cudaStream_t)side_stream;
  SideEvents& ev = g_side_events[current_device()];
  const bool use_side = side != nullptr && side != st && L > 0 && !(gg_debug_flags_peek() & 128) && ev.init();   // bit 7: in line (A/B)
  bool (&pending)[2] = ev.pending;
  if (phase_begin <= 0) pending[0] = pending[1] = false;
  int64_t arena_lo = 0, arena_floats = 0;
  const bool one_memset = arena_span(dm, offs, n_offs, &arena_lo, &arena_floats);
  struct Prezeroed { explicit Prezeroed(bool on) { set_prezeroed(on); } ~Prezeroed() { set_prezeroed(false); } } hint(one_memset);
  for (int ph = phase_begin < 0 ? 0 : phase_begin; ph < phase_end; ++ph) {
    if (ph == 0) {
      if (one_memset) {                                   // every accumulator of the pass: two memsets instead of ~50
        GG_CUDA(cudaMemsetAsync(grads + arena_lo, 0, sizeof(float) * arena_floats, st));
        GG_CUDA(cudaMemsetAsync(bws + b.zero_begin, 0, sizeof(float) * (b.zero_end - b.zero_begin), st));
      }
      GG_TRY(gg_gather_rows(dm.E, 1, g_scores, pl->perm, bws + b.g_int, stream));
      GG_TRY(gg_score_bwd(plan, d, dm.H, ws + w.h[L], ws + w.e[L], ws + w.Wq, ws + w.W1e, P(8), bws + b.g_int, ws + w.hid_s,
                          bws + b.g_h[0], bws + b.g_e[0], bws + b.dWq, bws + b.dbq, bws + b.dW1e, Gr(8), Gr(9),
                          bws + b.g_pre, bws + b.gQ, reinterpret_cast<double*>(bws + b.red), stream));
      GG_KERNEL_BEGIN("score_merge_kernel", st);
      score_merge_kernel<<<32, 256, 0, st>>>(dm.H, d, bws + b.dWq, bws + b.dbq, bws + b.dW1e, Gr(6), Gr(7));
      GG_KERNEL_END("score_merge_kernel", st);
    } else if (ph <= L) {
      const int l = L - ph, o = 10 + 8 * l;
      const int in = (ph - 1) & 1, out = ph & 1;
      const int k = ph & 1;                               // buffer set (gP, g_t) of this layer
      if (use_side) {
        if (pending[k]) GG_CUDA(cudaStreamWaitEvent(st, ev.done[k], 0));      // the GEMMs of two layers ago have read set k
        set_layer_bwd_side(side, ev.fork, ev.done[k]);
        pending[k] = true;
      }
      const int rc = gg_layer_bwd(plan, d, m->norm_kind, 1, ws + w.h[l], ws + w.e[l], ws + w.e[l + 1], P(o), P(o + 2), P(o + 4),
                                  P(o + 5), P(o + 6), P(o + 7), ws + w.P[l], ws + w.t[l], ws + w.z[l], ws + w.agg[l],
                                  reinterpret_cast<const double*>(ws + w.stats[l]), bws + b.g_h[in], bws + b.g_e[in],
                                  bws + b.g_h[out], bws + b.g_e[out], Gr(o), Gr(o + 1), Gr(o + 2), Gr(o + 3), Gr(o + 4),
                                  Gr(o + 5), Gr(o + 6), Gr(o + 7), bws + b.gP[k], bws + b.G, bws + b.g_eo, bws + b.g_t[k],
                                  reinterpret_cast<double*>(bws + b.bstats + 2LL * 4 * d * l), stream);
      set_layer_bwd_side(nullptr, nullptr, nullptr);
      if (rc) return rc;
    } else {
      const int in = L & 1;                              // where the gradient w.r.t. (h0, e0) ended up
      const float* g_h0 = bws + b.g_h[in];
      const float* g_e0 = bws + b.g_e[in];
      // edge encoder: dW2, db2, ReLU-masked hidden gradient, dW1 (K padded), db1          (full_graph.py:24-26)
      int rc = gg_edge_mlp_bwd(dm.E, d, dm.he, dm.edge_k4, g_e0, ws + w.hid_e, ws + w.e4, P(4), bws + b.dW1e4, Gr(3), Gr(4),
                               Gr(5), stream);
      if (rc == GG_ERR_UNSUPPORTED) {
        GG_TRY(gg_linear_bwd_weight(dm.E, d, dm.he, g_e0, ws + w.hid_e, Gr(4), Gr(5), stream));
        GG_TRY(gg_linear_bwd_data(dm.E, d, dm.he, g_e0, P(4), nullptr, ws + w.hid_e, bws + b.g_hid, stream));
        GG_TRY(gg_linear_bwd_weight(dm.E, dm.he, dm.edge_k4, bws + b.g_hid, ws + w.e4, bws + b.dW1e4, Gr(3), stream));
      } else if (rc) {
        return rc;
      }
      GG_TRY(copy_rows("unpad_weight_kernel", dm.he, dm.edge_k4, dm.edge_in, bws + b.dW1e4, nullptr, Gr(2), st));
      // node encoder                                                                        (full_graph.py:23)
      GG_TRY(gg_linear_bwd_weight(dm.N, d, dm.node_k4, g_h0, ws + w.pe4, bws + b.dWpe4, Gr(1), stream));
      GG_TRY(copy_rows("unpad_weight_kernel", d, dm.node_k4, dm.node_in, bws + b.dWpe4, nullptr, Gr(0), st));
    }
  }
  if (use_side && phase_end == L + 2) {                    // join: every weight gradient is complete on `stream`
    for (int k = 0; k < 2; ++k)
      if (pending[k]) { GG_CUDA(cudaStreamWaitEvent(st, ev.done[k], 0)); pending[k] = false; }
  }
  return GG_OK;
}

}  // extern "C"
