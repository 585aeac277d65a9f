// gg_plan.cu — graph plan: internal edge order (CSR over in-edges), out-edge CSR, permutations.
// Replaces the structural role of the DGLGraph argument and of dgl.reverse
// (models/full_graph.py:22, layers/gated_gcn_full.py:115): dgl.reverse keeps edge ids and swaps
// src/dst, which here is simply "walk the out-edge CSR instead of the in-edge CSR".
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <numeric>
#include <vector>

#include "gg_common.cuh"

namespace gg {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char* what) {
  g_last_error = std::string("gnnome_b200: CUDA error in ") + what + ": " + cudaGetErrorString(e);
  return GG_ERR_CUDA;
}

const char* last_error_cstr() { return g_last_error.c_str(); }

// ---- launch counter + optional per-launch CUDA-event timing (gg_profile_*) ----------------------
struct ProfRec { const char* name; cudaEvent_t a, b; };
static std::atomic<int64_t> g_launches{0};
static bool g_prof_on = false;
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof_recs;
static thread_local cudaEvent_t g_open_a = nullptr;

void kernel_begin(const char* name, cudaStream_t st) {
  (void)name;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_prof_on) return;
  cudaEvent_t a;
  if (cudaEventCreate(&a) != cudaSuccess) { g_open_a = nullptr; return; }
  cudaEventRecord(a, st);
  g_open_a = a;
}

void kernel_end(const char* name, cudaStream_t st) {
  if (!g_prof_on || g_open_a == nullptr) return;
  cudaEvent_t b;
  if (cudaEventCreate(&b) != cudaSuccess) return;
  cudaEventRecord(b, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_recs.push_back({name, g_open_a, b});
  g_open_a = nullptr;
}

static bool is_device_ptr(const void* p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

static void free_plan(Plan* p) {
  if (!p) return;
  free_sub_scratch(reinterpret_cast<SubScratch*>(p->sub_scratch));
  if (p->slab) { delete p; return; }      // sub-graph plan: the arrays live in a caller-owned slab
  cudaFree(p->host_slab);                 // gg_plan_create: one library-owned slab
  delete p;
}

}  // namespace gg

using gg::Plan;

extern "C" {

int gg_version(void) { return 200; }   // 100: round-1 ABI; 200: + gg_model_*, gg_edge_mlp_fwd, device-built plans, plan flags

const char* gg_last_error(void) { return gg::last_error_cstr(); }

int64_t gg_launch_count(void) { return gg::g_launches.load(); }

int gg_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(gg::g_prof_mu);
  gg::g_prof_on = on != 0;
  return GG_OK;
}

int gg_profile_report(char* buf, size_t cap) {
  GG_REQUIRE(buf && cap > 2, "profile_report: bad buffer");
  std::lock_guard<std::mutex> lk(gg::g_prof_mu);
  std::map<std::string, std::pair<int64_t, double>> agg;
  for (auto& r : gg::g_prof_recs) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      auto& e = agg[r.name];
      e.first += 1;
      e.second += ms;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  gg::g_prof_recs.clear();
  std::string out = "{";
  bool first = true;
  for (auto& kv : agg) {
    char line[256];
    std::snprintf(line, sizeof line, "%s\"%s\": [%lld, %.6f]", first ? "" : ", ", kv.first.c_str(),
                  (long long)kv.second.first, kv.second.second);
    out += line;
    first = false;
  }
  out += "}";
  GG_REQUIRE(out.size() + 1 <= cap, "profile_report: buffer too small");
  std::memcpy(buf, out.c_str(), out.size() + 1);
  return GG_OK;
}

int gg_plan_create(const int32_t* src, const int32_t* dst, int64_t N, int64_t E, void* stream_, gg_plan_t** out) {
  return gg_plan_create_ex(src, dst, N, E, GG_PLAN_RELABEL, stream_, out);
}

int gg_plan_create_ex(const int32_t* src, const int32_t* dst, int64_t N, int64_t E, int flags, void* stream_,
                      gg_plan_t** out) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GG_REQUIRE(out != nullptr, "plan_create: out is null");
  *out = nullptr;
  GG_REQUIRE(N >= 0 && E >= 0, "plan_create: negative size");
  GG_REQUIRE(N < (1LL << 31) && E < (1LL << 31), "plan_create: int32 indices only");
  GG_REQUIRE(E == 0 || (src && dst), "plan_create: null edge list");
  // Device builder (gg_plan_device.cu): sorts and relabelling on the GPU, one synchronisation.  A host-resident edge
  // list is uploaded first (one H2D, E x 8 bytes).  GG_PLAN_HOST=1 selects the round-1 host builder below (kept as
  // the reference the device builder is tested against, and for A/B timing).
  if (std::getenv("GG_PLAN_HOST") == nullptr && !(flags & GG_PLAN_HOST_BUILD)) {
    const int32_t* dsrc = src;
    const int32_t* ddst = dst;
    int32_t* up = nullptr;
    if (E > 0 && !gg::is_device_ptr(src)) {
      GG_CUDA(cudaMalloc((void**)&up, 2 * (size_t)E * sizeof(int32_t)));
      GG_CUDA(cudaMemcpyAsync(up, src, E * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
      GG_CUDA(cudaMemcpyAsync(up + E, dst, E * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
      dsrc = up; ddst = up + E;
    }
    Plan* pl = nullptr;
    const int rc = gg::plan_create_device(dsrc, ddst, N, E, flags, stream, &pl);
    if (up) {
      cudaStreamSynchronize(stream);        // host-resident edge list only: the staging copy is released synchronously
      cudaFree(up);
    }
    if (rc) return rc;
    *out = reinterpret_cast<gg_plan_t*>(pl);
    return GG_OK;
  }
  const bool timing = std::getenv("GG_PLAN_TIMING") != nullptr;
  auto t_start = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "gg_plan_create: %-14s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_start).count());
    t_start = now;
  };
  std::vector<int32_t> hs((size_t)E), hd((size_t)E);
  if (E > 0) {
    if (gg::is_device_ptr(src)) {
      GG_CUDA(cudaMemcpyAsync(hs.data(), src, E * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
      GG_CUDA(cudaMemcpyAsync(hd.data(), dst, E * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
      GG_CUDA(cudaStreamSynchronize(stream));
    } else {
      std::memcpy(hs.data(), src, E * sizeof(int32_t));
      std::memcpy(hd.data(), dst, E * sizeof(int32_t));
    }
  }
  for (int64_t i = 0; i < E; ++i)
    GG_REQUIRE(hs[i] >= 0 && hs[i] < N && hd[i] >= 0 && hd[i] < N, "plan_create: node index out of range");

  // ---- internal node order.  Assembly graphs are near-linear (an edge joins reads that overlap on the
  // genome) but read ids are arbitrary, so row gathers of node features are random 512-byte DRAM accesses.
  // A breadth-first (Cuthill-McKee style) relabelling over the undirected graph makes neighbours close in
  // memory; legal because the model boundary exposes per-edge logits only.
  std::vector<int32_t> node_perm((size_t)N), node_inv((size_t)N);      // internal -> caller, caller -> internal
  if ((flags & GG_PLAN_RELABEL) && N > 0) {
    std::vector<int32_t> uptr((size_t)N + 1, 0);
    for (int64_t i = 0; i < E; ++i) { uptr[hs[i] + 1]++; uptr[hd[i] + 1]++; }
    for (int64_t v = 0; v < N; ++v) uptr[v + 1] += uptr[v];
    std::vector<int32_t> uadj((size_t)2 * E);
    {
      std::vector<int32_t> cur(uptr.begin(), uptr.end() - 1);
      for (int64_t i = 0; i < E; ++i) { uadj[cur[hs[i]]++] = hd[i]; uadj[cur[hd[i]]++] = hs[i]; }
    }
    std::vector<int32_t> order((size_t)N);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
      return (uptr[a + 1] - uptr[a]) < (uptr[b + 1] - uptr[b]);       // component seeds: lowest degree first
    });
    std::vector<char> seen((size_t)N, 0);
    int64_t filled = 0;
    for (int64_t oi = 0; oi < N; ++oi) {
      const int32_t seed = order[oi];
      if (seen[seed]) continue;
      seen[seed] = 1;
      int64_t head = filled;
      node_perm[filled++] = seed;
      while (head < filled) {
        const int32_t u = node_perm[head++];
        for (int32_t k = uptr[u]; k < uptr[u + 1]; ++k) {
          const int32_t v = uadj[k];
          if (!seen[v]) { seen[v] = 1; node_perm[filled++] = v; }
        }
      }
    }
  } else {
    std::iota(node_perm.begin(), node_perm.end(), 0);
  }
  for (int64_t p = 0; p < N; ++p) node_inv[node_perm[p]] = (int32_t)p;
  for (int64_t i = 0; i < E; ++i) { hs[i] = node_inv[hs[i]]; hd[i] = node_inv[hd[i]]; }
  lap("copy+relabel");

  // stable counting sort by dst -> internal order
  std::vector<int32_t> in_ptr((size_t)N + 1, 0), out_ptr((size_t)N + 1, 0);
  for (int64_t i = 0; i < E; ++i) { in_ptr[hd[i] + 1]++; out_ptr[hs[i] + 1]++; }
  for (int64_t v = 0; v < N; ++v) { in_ptr[v + 1] += in_ptr[v]; out_ptr[v + 1] += out_ptr[v]; }
  std::vector<int32_t> perm((size_t)E), inv((size_t)E), isrc((size_t)E), idst((size_t)E);
  {
    std::vector<int32_t> cur(in_ptr.begin(), in_ptr.end() - 1);
    for (int64_t i = 0; i < E; ++i) {
      const int32_t p = cur[hd[i]]++;
      perm[p] = (int32_t)i;
      inv[i] = p;
    }
  }
  for (int64_t p = 0; p < E; ++p) { isrc[p] = hs[perm[p]]; idst[p] = hd[perm[p]]; }
  // out-edge CSR over internal ids (stable: increasing internal id within a source)
  std::vector<int32_t> out_eid((size_t)E), out_dst((size_t)E);
  {
    std::vector<int32_t> cur(out_ptr.begin(), out_ptr.end() - 1);
    for (int64_t p = 0; p < E; ++p) {
      const int32_t slot = cur[isrc[p]]++;
      out_eid[slot] = (int32_t)p;
      out_dst[slot] = idst[p];
    }
  }

  lap("sort+csr");
  Plan* pl = new Plan();
  pl->N = N; pl->E = E;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&pl->num_sms, cudaDevAttrMultiProcessorCount, dev);
  // one device slab, one H2D copy: [src | dst | out_eid | out_dst | perm | inv_perm | in_ptr | out_ptr | node_perm | node_inv]
  const size_t e1 = (size_t)(E ? E : 1), n1 = (size_t)N + 1;
  const size_t words = 6 * e1 + 4 * n1;
  std::vector<int32_t> stage(words, 0);
  int32_t* slab = nullptr;
  cudaError_t e = cudaMalloc((void**)&slab, words * sizeof(int32_t));
  if (e != cudaSuccess) { delete pl; return gg::cuda_fail(e, "plan_create alloc"); }
  pl->host_slab = slab;
  size_t off = 0;
  auto put = [&](int32_t** d, const std::vector<int32_t>& h, size_t cap) {
    *d = slab + off;
    if (!h.empty()) std::memcpy(stage.data() + off, h.data(), h.size() * sizeof(int32_t));
    off += cap;
  };
  put(&pl->src, isrc, e1); put(&pl->dst, idst, e1); put(&pl->out_eid, out_eid, e1); put(&pl->out_dst, out_dst, e1);
  put(&pl->perm, perm, e1); put(&pl->inv_perm, inv, e1);
  put(&pl->in_ptr, in_ptr, n1); put(&pl->out_ptr, out_ptr, n1); put(&pl->node_perm, node_perm, n1); put(&pl->node_inv, node_inv, n1);
  e = cudaMemcpyAsync(slab, stage.data(), words * sizeof(int32_t), cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);   // the host staging vector dies at return
  if (e != cudaSuccess) {
    gg::free_plan(pl);
    return gg::cuda_fail(e, "plan_create upload");
  }
  lap("alloc+upload");
  *out = reinterpret_cast<gg_plan_t*>(pl);
  return GG_OK;
}

int gg_plan_destroy(gg_plan_t* plan) {
  gg::free_plan(reinterpret_cast<Plan*>(plan));
  return GG_OK;
}

#define GG_PLAN_GET(name, field, type, dflt)                              \
  type name(const gg_plan_t* plan) {                                      \
    return plan ? reinterpret_cast<const Plan*>(plan)->field : dflt;      \
  }
GG_PLAN_GET(gg_plan_num_nodes, N, int64_t, -1)
GG_PLAN_GET(gg_plan_num_edges, E, int64_t, -1)
GG_PLAN_GET(gg_plan_perm, perm, const int32_t*, nullptr)
GG_PLAN_GET(gg_plan_inv_perm, inv_perm, const int32_t*, nullptr)
GG_PLAN_GET(gg_plan_src, src, const int32_t*, nullptr)
GG_PLAN_GET(gg_plan_dst, dst, const int32_t*, nullptr)
GG_PLAN_GET(gg_plan_in_ptr, in_ptr, const int32_t*, nullptr)
GG_PLAN_GET(gg_plan_out_ptr, out_ptr, const int32_t*, nullptr)
GG_PLAN_GET(gg_plan_out_eid, out_eid, const int32_t*, nullptr)


int gg_plan_copy_array(const gg_plan_t* plan, int which, int32_t* out, void* stream) {
  GG_REQUIRE(plan && out, "plan_copy_array: null pointer");
  const Plan* p = reinterpret_cast<const Plan*>(plan);
  const int32_t* srcs[12] = {p->perm, p->inv_perm, p->src, p->dst, p->in_ptr, p->out_ptr, p->out_eid,
                             p->node_perm, p->node_inv, p->parent_eid, p->csrc, p->cdst};
  GG_REQUIRE(which >= 0 && which < 12, "plan_copy_array: bad selector");
  GG_REQUIRE(which < 9 || p->slab != nullptr, "plan_copy_array: selectors 9-11 exist on sub-graph plans only");
  const int64_t n = (which == 4 || which == 5) ? p->N + 1 : ((which == 7 || which == 8) ? p->N : p->E);
  if (n > 0)
    GG_CUDA(cudaMemcpyAsync(out, srcs[which], n * sizeof(int32_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return GG_OK;
}

}  // extern "C"
