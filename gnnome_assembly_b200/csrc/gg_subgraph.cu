// gg_subgraph.cu — node-induced sub-graph plans for the mini-batch path, built on the device.
//
// The reference's mini-batch branch (train.py:282-312, :428-456) cuts every training graph into METIS
// clusters (dgl.dataloading.ClusterGCNSampler) and steps on g.subgraph(union of batch_size clusters).  The
// engine needs a plan per sub-graph; rebuilding one on the host (D2H of the edge list, counting sort, H2D)
// for each of the thousands of batches of an epoch would dominate the step.  Because the parent plan is
// already dst-sorted and breadth-first relabelled, the sub-graph's plan is a STREAM COMPACTION of the
// parent's arrays: keep the selected nodes in parent-internal order (monotone renumbering keeps every
// edge array sorted and inherits the parent's gather locality), keep the edges whose two ends survive.
// Four exclusive scans (nodes, in-edge order, out-edge slots, caller edge-id order) + four scatters; all
// integer HBM traffic, O(N + E), no sort, no host round trip except the one int64 that sizes the result.
// Two calls because the caller owns every buffer: gg_subplan_count (mark + scans, returns the edge count),
// then gg_subplan_fill scatters into a caller-allocated slab (PyTorch's caching allocator: no cudaMalloc per batch).
#include <algorithm>

#include "gg_common.cuh"

namespace gg {

constexpr int kScanThreads = 256, kScanItems = 8, kScanTile = kScanThreads * kScanItems;

// ---- flag functors (flag(i) in {0,1}) ---------------------------------------------------------
struct NodeFlag {       // internal node i selected?   (mark[i] = caller sub id + 1, 0 = not selected)
  const int32_t* mark;
  __device__ int operator()(int64_t i) const { return mark[i] != 0; }
};
struct EdgeFlag {       // internal edge p kept?
  const int32_t* mark; const int32_t* src; const int32_t* dst;
  __device__ int operator()(int64_t p) const { return (mark[src[p]] != 0) & (mark[dst[p]] != 0); }
};
struct IndirectEdgeFlag {   // slot k -> internal edge map[k] kept?  (out-edge slots, caller edge ids)
  EdgeFlag f; const int32_t* map;
  __device__ int operator()(int64_t k) const { return f(map[k]); }
};

template <class F>
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(int64_t n, F flag, int32_t* __restrict__ tsum) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = base + (int64_t)k * kScanThreads + threadIdx.x;      // coalesced
    if (i < n) s += flag(i);
  }
  __shared__ int sh[kScanThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int a = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) a += sh[w];
    tsum[blockIdx.x] = a;
  }
}

// exclusive scan of the tile sums in place by ONE block (tiles <= a few thousand); total -> tsum[tiles]
__global__ void __launch_bounds__(1024) scan_tile_offsets_kernel(int tiles, int32_t* __restrict__ tsum) {
  __shared__ int sh[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < tiles; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < tiles ? tsum[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) sh[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = sh[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
      sh[lane] = w;                                      // inclusive over warps
    }
    __syncthreads();
    const int incl = x + (warp ? sh[warp - 1] : 0) + carry;
    if (i < tiles) tsum[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) tsum[tiles] = carry;
}

// out[i] = number of set flags before i, i in [0, n]; thread owns kScanItems CONSECUTIVE items
template <class F>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(int64_t n, F flag, const int32_t* __restrict__ tsum,
                                                                  int32_t* __restrict__ out) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int f[kScanItems], s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) { f[k] = (base + k < n) ? flag(base + k) : 0; s += f[k]; }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int x = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  __shared__ int sh[kScanThreads / 32];
  if (lane == 31) sh[warp] = x;
  __syncthreads();
  int pre = tsum[blockIdx.x] + x - s;
  for (int w = 0; w < warp; ++w) pre += sh[w];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k <= n) out[base + k] = pre;              // includes the one-past-the-end total
    pre += f[k];
  }
}

template <class F>
static int exclusive_scan(const char* tag, int64_t n, F flag, int32_t* tsum, int32_t* out, cudaStream_t st) {
  const int tiles = (int)((n + 1 + kScanTile - 1) / kScanTile);          // n + 1 outputs
  GG_KERNEL_BEGIN(tag, st);
  scan_tile_sums_kernel<<<tiles, kScanThreads, 0, st>>>(n, flag, tsum);
  GG_KERNEL_END(tag, st);
  GG_KERNEL_BEGIN("scan_tile_offsets_kernel", st);
  scan_tile_offsets_kernel<<<1, 1024, 0, st>>>(tiles, tsum);
  GG_KERNEL_END("scan_tile_offsets_kernel", st);
  GG_KERNEL_BEGIN(tag, st);
  scan_apply_kernel<<<tiles, kScanThreads, 0, st>>>(n, flag, tsum, out);
  GG_KERNEL_END(tag, st);
  return GG_OK;
}

// ---- mark / scatter kernels --------------------------------------------------------------------
// mark[node_inv[nodes[j]]] = j + 1 ; err |= 1 for an id out of range, 2 for a duplicate
__global__ void sub_mark_kernel(int64_t n, const int64_t* __restrict__ nodes, int64_t N,
                                const int32_t* __restrict__ node_inv, int32_t* __restrict__ mark, int* __restrict__ err) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = nodes[j];
    if (u < 0 || u >= N) { atomicOr(err, 1); continue; }
    if (atomicExch(mark + node_inv[u], (int32_t)(j + 1)) != 0) atomicOr(err, 2);
  }
}

struct SubArrays {
  int32_t *src, *dst, *in_ptr, *out_ptr, *out_eid, *out_dst, *perm, *inv_perm, *node_perm, *node_inv;
  int32_t *parent_eid, *csrc, *cdst;
};

__global__ void sub_edges_kernel(int64_t E, const int32_t* __restrict__ mark, const int32_t* __restrict__ newid,
                                 const int32_t* __restrict__ epos, const int32_t* __restrict__ cpos,
                                 const int32_t* __restrict__ psrc, const int32_t* __restrict__ pdst,
                                 const int32_t* __restrict__ pperm, SubArrays o) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < E; p += (int64_t)gridDim.x * blockDim.x) {
    const int32_t s = psrc[p], d = pdst[p];
    const int32_t ms = mark[s], md = mark[d];
    if (ms == 0 || md == 0) continue;
    const int32_t q = epos[p], eid = pperm[p], c = cpos[eid];
    o.src[q] = newid[s];
    o.dst[q] = newid[d];
    o.perm[q] = c;
    o.inv_perm[c] = q;
    o.parent_eid[c] = eid;
    o.csrc[c] = ms - 1;
    o.cdst[c] = md - 1;
  }
}

__global__ void sub_out_kernel(int64_t E, const int32_t* __restrict__ mark, const int32_t* __restrict__ newid,
                               const int32_t* __restrict__ epos, const int32_t* __restrict__ opos,
                               const int32_t* __restrict__ psrc, const int32_t* __restrict__ pdst,
                               const int32_t* __restrict__ pout_eid, SubArrays o) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < E; k += (int64_t)gridDim.x * blockDim.x) {
    const int32_t p = pout_eid[k];
    const int32_t d = pdst[p];
    if (mark[psrc[p]] == 0 || mark[d] == 0) continue;
    const int32_t slot = opos[k];
    o.out_eid[slot] = epos[p];
    o.out_dst[slot] = newid[d];
  }
}

__global__ void sub_nodes_kernel(int64_t N, int64_t E, const int32_t* __restrict__ mark,
                                 const int32_t* __restrict__ newid, const int32_t* __restrict__ epos,
                                 const int32_t* __restrict__ opos, const int32_t* __restrict__ pin_ptr,
                                 const int32_t* __restrict__ pout_ptr, SubArrays o) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= N; i += (int64_t)gridDim.x * blockDim.x) {
    if (i == N) {                                        // closing entries
      o.in_ptr[newid[N]] = epos[E];
      o.out_ptr[newid[N]] = opos[E];
      continue;
    }
    const int32_t m = mark[i];
    if (m == 0) continue;
    const int32_t ni = newid[i];
    o.in_ptr[ni] = epos[pin_ptr[i]];
    o.out_ptr[ni] = opos[pout_ptr[i]];
    o.node_perm[ni] = m - 1;
    o.node_inv[m - 1] = ni;
  }
}

// scratch of a parent plan for building sub-plans (sized by the parent; reused by every gg_subplan_count)
struct SubScratch {
  int32_t* mark = nullptr;     // [N]
  int32_t* newid = nullptr;    // [N+1]
  int32_t* epos = nullptr;     // [E+1]
  int32_t* cpos = nullptr;     // [E+1]
  int32_t* opos = nullptr;     // [E+1]
  int32_t* tsum = nullptr;     // tile sums
  int* err = nullptr;
  int32_t* host = nullptr;     // pinned [4]: N_sub, E_sub, E_out, err
  int64_t pending_n = -1, pending_e = 0;   // result of the last gg_subplan_count, consumed by gg_subplan_fill
};

void free_sub_scratch(SubScratch* s) {
  if (!s) return;
  cudaFree(s->mark);           // one slab
  cudaFreeHost(s->host);
  delete s;
}

static unsigned flat_grid(int64_t n, int threads, int num_sms) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = (int64_t)num_sms * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace gg

using namespace gg;

extern "C" {

// words (int32) of the caller-owned slab that backs a sub-graph plan of n nodes and num_edges edges
size_t gg_subplan_slab_words(int64_t n, int64_t num_edges) {
  const size_t e1 = (size_t)(num_edges > 0 ? num_edges : 1), n1 = (size_t)(n > 0 ? n : 0) + 1;
  return 9 * e1 + 4 * n1;
}

int gg_subplan_count(const gg_plan_t* parent_, const int64_t* nodes, int64_t n, void* stream_, int64_t* num_edges) {
  GG_REQUIRE(parent_ != nullptr && num_edges != nullptr, "subplan_count: null argument");
  Plan* par = const_cast<Plan*>(reinterpret_cast<const Plan*>(parent_));
  GG_REQUIRE(n >= 0 && n <= par->N, "subplan_count: more nodes than the parent has");
  GG_REQUIRE(n == 0 || nodes != nullptr, "subplan_count: null node list");
  cudaStream_t st = (cudaStream_t)stream_;
  const int64_t N = par->N, E = par->E;

  if (par->sub_scratch == nullptr) {
    SubScratch* s = new SubScratch();
    const int64_t tiles = (std::max(N, E) + 1 + kScanTile - 1) / kScanTile + 2;
    const size_t words = (size_t)N + (N + 1) + 3 * (size_t)(E + 1) + (size_t)tiles + 8;
    cudaError_t e = cudaMalloc((void**)&s->mark, words * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&s->host, 4 * sizeof(int32_t));
    if (e != cudaSuccess) { free_sub_scratch(s); return cuda_fail(e, "subplan_count scratch"); }
    s->newid = s->mark + N;
    s->epos = s->newid + (N + 1);
    s->cpos = s->epos + (E + 1);
    s->opos = s->cpos + (E + 1);
    s->tsum = s->opos + (E + 1);
    s->err = reinterpret_cast<int*>(s->tsum + tiles);
    par->sub_scratch = s;
  }
  SubScratch* sc = reinterpret_cast<SubScratch*>(par->sub_scratch);
  sc->pending_n = -1;
  GG_CUDA(cudaMemsetAsync(sc->mark, 0, sizeof(int32_t) * (size_t)N, st));
  GG_CUDA(cudaMemsetAsync(sc->err, 0, sizeof(int), st));
  if (n > 0) {
    GG_KERNEL_BEGIN("sub_mark_kernel", st);
    sub_mark_kernel<<<flat_grid(n, 256, par->num_sms), 256, 0, st>>>(n, nodes, N, par->node_inv, sc->mark, sc->err);
    GG_KERNEL_END("sub_mark_kernel", st);
  }
  const EdgeFlag ef{sc->mark, par->src, par->dst};
  int rc = exclusive_scan("sub_scan_nodes", N, NodeFlag{sc->mark}, sc->tsum, sc->newid, st);
  if (rc) return rc;
  rc = exclusive_scan("sub_scan_in_edges", E, ef, sc->tsum, sc->epos, st);
  if (rc) return rc;
  rc = exclusive_scan("sub_scan_edge_ids", E, IndirectEdgeFlag{ef, par->inv_perm}, sc->tsum, sc->cpos, st);
  if (rc) return rc;
  rc = exclusive_scan("sub_scan_out_slots", E, IndirectEdgeFlag{ef, par->out_eid}, sc->tsum, sc->opos, st);
  if (rc) return rc;
  GG_CUDA(cudaMemcpyAsync(sc->host + 0, sc->newid + N, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  GG_CUDA(cudaMemcpyAsync(sc->host + 1, sc->epos + E, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  GG_CUDA(cudaMemcpyAsync(sc->host + 2, sc->opos + E, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  GG_CUDA(cudaMemcpyAsync(sc->host + 3, sc->err, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  GG_CUDA(cudaStreamSynchronize(st));                      // the one host round trip: the size of the result
  GG_REQUIRE((sc->host[3] & 1) == 0, "subplan_count: node id out of range");
  GG_REQUIRE((sc->host[3] & 2) == 0 && sc->host[0] == n, "subplan_count: duplicate node ids");
  GG_REQUIRE(sc->host[2] == sc->host[1], "subplan_count: internal error (out-edge count differs)");
  sc->pending_n = n;
  sc->pending_e = sc->host[1];
  *num_edges = sc->host[1];
  return GG_OK;
}

int gg_subplan_fill(const gg_plan_t* parent_, int32_t* slab, void* stream_, gg_plan_t** out) {
  GG_REQUIRE(out != nullptr, "subplan_fill: out is null");
  *out = nullptr;
  GG_REQUIRE(parent_ != nullptr && slab != nullptr, "subplan_fill: null argument");
  const Plan* par = reinterpret_cast<const Plan*>(parent_);
  SubScratch* sc = reinterpret_cast<SubScratch*>(par->sub_scratch);
  GG_REQUIRE(sc != nullptr && sc->pending_n >= 0, "subplan_fill: no gg_subplan_count result pending on this parent");
  cudaStream_t st = (cudaStream_t)stream_;
  const int64_t N = par->N, E = par->E, Ns = sc->pending_n, Es = sc->pending_e;
  sc->pending_n = -1;

  Plan* pl = new Plan();
  pl->N = Ns; pl->E = Es; pl->num_sms = par->num_sms;
  pl->slab = slab;                                         // caller-owned: gg_plan_destroy leaves it alone
  const size_t e1 = (size_t)(Es ? Es : 1), n1 = (size_t)Ns + 1;
  int32_t* w = slab;
  auto take = [&](size_t k) { int32_t* r = w; w += k; return r; };
  pl->src = take(e1); pl->dst = take(e1); pl->out_eid = take(e1); pl->out_dst = take(e1);
  pl->perm = take(e1); pl->inv_perm = take(e1);
  pl->parent_eid = take(e1); pl->csrc = take(e1); pl->cdst = take(e1);
  pl->in_ptr = take(n1); pl->out_ptr = take(n1); pl->node_perm = take(n1); pl->node_inv = take(n1);
  SubArrays o{pl->src, pl->dst, pl->in_ptr, pl->out_ptr, pl->out_eid, pl->out_dst, pl->perm, pl->inv_perm,
              pl->node_perm, pl->node_inv, pl->parent_eid, pl->csrc, pl->cdst};
  if (E > 0) {
    kernel_begin("sub_edges_kernel", st);
    sub_edges_kernel<<<flat_grid(E, 256, par->num_sms), 256, 0, st>>>(E, sc->mark, sc->newid, sc->epos, sc->cpos,
                                                                      par->src, par->dst, par->perm, o);
    kernel_end("sub_edges_kernel", st);
    kernel_begin("sub_out_kernel", st);
    sub_out_kernel<<<flat_grid(E, 256, par->num_sms), 256, 0, st>>>(E, sc->mark, sc->newid, sc->epos, sc->opos,
                                                                    par->src, par->dst, par->out_eid, o);
    kernel_end("sub_out_kernel", st);
  }
  kernel_begin("sub_nodes_kernel", st);
  sub_nodes_kernel<<<flat_grid(N + 1, 256, par->num_sms), 256, 0, st>>>(N, E, sc->mark, sc->newid, sc->epos, sc->opos,
                                                                        par->in_ptr, par->out_ptr, o);
  kernel_end("sub_nodes_kernel", st);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { delete pl; return cuda_fail(e, "subplan_fill scatter"); }
  *out = reinterpret_cast<gg_plan_t*>(pl);
  return GG_OK;
}

}  // extern "C"
