// gg_plan_device.cu — graph plan built ON THE DEVICE (whole graphs): no D2H of the edge list, no host sort, one
// stream synchronisation (the error flag of the index check, read right after the first kernel).
//
// Replaces the host builder of round 1 (D2H + counting sort + breadth-first relabelling on one core + H2D: 27 ms for a
// chr19 graph, 64 ms measured for chr21 on the GPU box — 5x the forward it precedes).  Same contract as before
// (gg_plan_create in gg_plan.cu, which now calls this whenever the edge list is device-resident or a device exists):
//   internal edge order = STABLE sort of the caller's edges by (relabelled) destination   -> perm / inv_perm / in_ptr
//   out-edge CSR        = STABLE sort of the internal edge ids by (relabelled) source      -> out_ptr / out_eid / out_dst
// both by cub::DeviceRadixSort (LSD radix sort, stable), CSR pointers by binary search in the sorted keys.
//
// Node relabelling (GG_PLAN_RELABEL).  Goal: neighbours close in memory (assembly graphs are near-linear, read ids are
// arbitrary).  A level-synchronous breadth-first search is hopeless here: the graph's diameter is thousands of levels.
// Two-level ordering instead, all data-parallel except one warp:
//   1. REGIONS: one node in 64 (by a hash of its id) seeds a region; (distance, seed) keys are relaxed over the edges with 64-bit
//      atomicMin until nothing changes (a cooperative kernel, grid sync per sweep; ~10 sweeps: a region's radius).  The
//      fixed point — nearest seed, ties to the smaller seed id — does not depend on execution order.  Components that
//      hold no seed are re-seeded (every 4th remaining node, then every remaining node).
//   2. REGION GRAPH: region pairs joined by an edge, sorted + uniqued (radix sort of 2E 64-bit keys) into a CSR of
//      ~N/64 vertices; ONE WARP walks it breadth-first (sequential over the queue, 32 neighbours at a time), which is
//      exact, deterministic, and ~1.5 us per region: 1 ms for chr19.
//   3. nodes sorted by (rank of their region, distance to its seed), ties by node id (stable radix sort).
// Neighbouring nodes end up within a few regions (a few hundred rows) of each other; isolated nodes go last.
#include <cooperative_groups.h>
#include <cub/cub.cuh>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gg_common.cuh"

namespace cg = cooperative_groups;

namespace gg {

constexpr unsigned long long kInfKey = ~0ull;

__global__ void plan_check_degree_kernel(int64_t E, int64_t N, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                         int32_t* __restrict__ udeg, int* __restrict__ err) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) {
    const int s = src[i], d = dst[i];
    if (s < 0 || s >= N || d < 0 || d >= N) { atomicOr(err, 1); continue; }
    if (udeg != nullptr && s != d) { atomicAdd(udeg + s, 1); atomicAdd(udeg + d, 1); }
  }
}

// regions: (dist << 32 | seed) relaxed to the fixed point, three seeding rounds (see the header)
__global__ void __launch_bounds__(256) plan_regions_kernel(int64_t E, int64_t N, const int32_t* __restrict__ src,
                                                           const int32_t* __restrict__ dst, const int32_t* __restrict__ udeg,
                                                           unsigned long long* __restrict__ key, int32_t* __restrict__ is_seed,
                                                           int* __restrict__ changed, const int* __restrict__ err) {
  cg::grid_group grid = cg::this_grid();
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  if (*err) return;                                          // bad indices: every thread leaves before the first sync
  for (int round = 0; round < 3; ++round) {
    for (int64_t v = tid; v < N; v += nth) {
      const bool connected = udeg[v] > 0;
      // seeds are drawn by a multiplicative hash of the id, NOT by id modulo: assembly graphs put the two strands of
      // read k at ids 2k / 2k+1 and the strands are separate components, so "every 64th id" seeds one strand only
      const unsigned hv = (unsigned)v * 2654435761u;
      if (round == 0) {
        const bool seed = connected && (hv >> 26) == 0;             // 1 in 64
        key[v] = seed ? (unsigned long long)v : kInfKey;
        is_seed[v] = seed ? 1 : 0;
      } else if (connected && key[v] == kInfKey && (round == 2 || ((hv >> 8) & 3u) == 0)) {
        key[v] = (unsigned long long)v;
        is_seed[v] = 1;
      }
    }
    grid.sync();
    if (round == 2) break;                                   // the last round's seeds are singletons among seeds
    for (int sweep = 0;; ++sweep) {
      if (tid == 0) changed[sweep & 1] = 0;
      grid.sync();
      int any = 0;
      for (int64_t i = tid; i < E; i += nth) {
        const int s = src[i], d = dst[i];
        if (s == d) continue;
        const unsigned long long ks = key[s], kd = key[d];
        if (ks != kInfKey && ks + (1ull << 32) < kd) { atomicMin(key + d, ks + (1ull << 32)); any = 1; }
        if (kd != kInfKey && kd + (1ull << 32) < ks) { atomicMin(key + s, kd + (1ull << 32)); any = 1; }
      }
      if (any) changed[sweep & 1] = 1;
      if (tid == 0) atomicAdd(changed + 2, 1);              // sweep counter (diagnostics)
      grid.sync();
      if (!changed[sweep & 1]) break;
      grid.sync();                                           // (the flag of this parity is reset two sweeps later)
    }
  }
}

// region id of every node's seed; emits the region pairs of the cross-region edges (both directions) as sort keys
__global__ void plan_region_pairs_kernel(int64_t E, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                         const unsigned long long* __restrict__ key, const int32_t* __restrict__ seed_rank,
                                         unsigned long long* __restrict__ pairs) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) {
    const int s = src[i], d = dst[i];
    unsigned long long a = kInfKey, b = kInfKey;
    if (s != d) {
      const unsigned ra = (unsigned)seed_rank[(unsigned)(key[s] & 0xffffffffu)];
      const unsigned rb = (unsigned)seed_rank[(unsigned)(key[d] & 0xffffffffu)];
      if (ra != rb) { a = ((unsigned long long)ra << 32) | rb; b = ((unsigned long long)rb << 32) | ra; }
    }
    pairs[2 * i] = a;
    pairs[2 * i + 1] = b;
  }
}

__global__ void plan_pair_flags_kernel(int64_t n, const unsigned long long* __restrict__ sorted, int32_t* __restrict__ flag) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    flag[i] = (sorted[i] != kInfKey && (i == 0 || sorted[i] != sorted[i - 1])) ? 1 : 0;
}

__global__ void plan_pair_compact_kernel(int64_t n, const unsigned long long* __restrict__ sorted, const int32_t* __restrict__ flag,
                                         const int32_t* __restrict__ pos, int32_t* __restrict__ radj_src, int32_t* __restrict__ radj) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (flag[i]) { radj_src[pos[i]] = (int32_t)(sorted[i] >> 32); radj[pos[i]] = (int32_t)(sorted[i] & 0xffffffffu); }
}

// out[v] = first index i in [0, n) with keys[i] >= v   (v in [0, nv]); n may live on the device (n_dev)
__global__ void plan_lower_bound_kernel(int64_t nv, const int32_t* __restrict__ keys, int64_t n, const int32_t* __restrict__ n_dev,
                                        int32_t* __restrict__ out) {
  const int64_t cnt = n_dev ? (int64_t)*n_dev : n;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v <= nv; v += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = 0, hi = cnt;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (keys[mid] < v) lo = mid + 1; else hi = mid;
    }
    out[v] = (int32_t)lo;
  }
}

// breadth-first order of the region graph by ONE warp: rank[r] = position of region r in the visiting order
__global__ void __launch_bounds__(32) plan_region_bfs_kernel(const int32_t* __restrict__ n_regions, const int32_t* __restrict__ rptr,
                                                             const int32_t* __restrict__ radj, int32_t* __restrict__ queue,
                                                             int32_t* __restrict__ rank) {
  const int lane = threadIdx.x;
  const int R = *n_regions;
  for (int r = lane; r < R; r += 32) rank[r] = -1;
  __syncwarp();
  int filled = 0, next_seed = 0;
  while (filled < R) {
    while (rank[next_seed] >= 0) ++next_seed;                // uniform across the warp (same memory, after __syncwarp)
    if (lane == 0) { queue[filled] = next_seed; rank[next_seed] = filled; }
    __syncwarp();
    int head = filled;
    ++filled;
    while (head < filled) {
      const int u = queue[head++];
      const int beg = rptr[u], end = rptr[u + 1];
      for (int base = beg; base < end; base += 32) {
        const int k = base + lane;
        const int v = k < end ? radj[k] : -1;
        const bool fresh = v >= 0 && rank[v] < 0;            // unique neighbour lists: no two lanes hold the same v
        const unsigned m = __ballot_sync(0xffffffffu, fresh);
        if (fresh) {
          const int pos = filled + __popc(m & ((1u << lane) - 1));
          queue[pos] = v;
          rank[v] = pos;
        }
        filled += __popc(m);
        __syncwarp();
      }
    }
  }
}

__global__ void plan_node_keys_kernel(int64_t N, const int32_t* __restrict__ udeg, const unsigned long long* __restrict__ key,
                                      const int32_t* __restrict__ seed_rank, const int32_t* __restrict__ rank,
                                      unsigned long long* __restrict__ nkey, int32_t* __restrict__ ids) {
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < N; v += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long k = (1ull << 47);                     // isolated nodes: after every region, by id
    if (udeg[v] > 0) {
      const unsigned long long kk = key[v];
      const unsigned dist = (unsigned)(kk >> 32);
      const int r = rank[seed_rank[(unsigned)(kk & 0xffffffffu)]];
      k = ((unsigned long long)(unsigned)r << 16) | (dist < 0xffffu ? dist : 0xffffu);
    }
    nkey[v] = k;
    ids[v] = (int32_t)v;
  }
}

__global__ void plan_invert_kernel(int64_t n, const int32_t* __restrict__ perm, int32_t* __restrict__ inv) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) inv[perm[p]] = (int32_t)p;
}

__global__ void plan_iota_kernel(int64_t n, int32_t* __restrict__ a, int32_t* __restrict__ b) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    a[p] = (int32_t)p;
    if (b) b[p] = (int32_t)p;
  }
}

// keys for the dst sort: relabelled destination of caller edge i; values: i
__global__ void plan_dst_keys_kernel(int64_t E, const int32_t* __restrict__ dst, const int32_t* __restrict__ node_inv,
                                     int32_t* __restrict__ keys, int32_t* __restrict__ vals) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) {
    keys[i] = node_inv[dst[i]];
    vals[i] = (int32_t)i;
  }
}

// internal order known (perm): internal src, inverse permutation, keys / values of the out-edge sort
__global__ void plan_internal_src_kernel(int64_t E, const int32_t* __restrict__ src, const int32_t* __restrict__ node_inv,
                                         const int32_t* __restrict__ perm, int32_t* __restrict__ isrc, int32_t* __restrict__ inv_perm,
                                         int32_t* __restrict__ ids) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < E; p += (int64_t)gridDim.x * blockDim.x) {
    const int i = perm[p];
    isrc[p] = node_inv[src[i]];
    inv_perm[i] = (int32_t)p;
    ids[p] = (int32_t)p;
  }
}

__global__ void plan_gather_kernel(int64_t n, const int32_t* __restrict__ in, const int32_t* __restrict__ idx, int32_t* __restrict__ out) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) out[k] = in[idx[k]];
}

static unsigned grid_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (unsigned)b;
}

static int bits_for(int64_t n) {
  int b = 1;
  while (b < 63 && (1LL << b) < n) ++b;
  return b;
}

// Stream-ordered scratch from a memory pool of the library's own (one per device) that keeps up to 1 GiB cached between
// calls: with the default pool every synchronisation hands the memory back to the driver, and the ~25 scratch arrays of
// one build then cost ~8 ms of allocation calls — more than all the kernels together.
static cudaMemPool_t scratch_pool() {
  static cudaMemPool_t pools[kMaxDevices] = {};
  const int dev = current_device();
  if (pools[dev] == nullptr) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool = nullptr;
    if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    unsigned long long keep = 1ull << 30;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    pools[dev] = pool;
  }
  return pools[dev];
}

struct TempPool {                      // released (back to the pool) when the builder returns
  cudaStream_t st;
  cudaMemPool_t pool;
  std::vector<void*> ptrs;
  explicit TempPool(cudaStream_t s) : st(s), pool(scratch_pool()) {}
  ~TempPool() { for (void* p : ptrs) cudaFreeAsync(p, st); }
  template <class T> T* get(size_t n) {
    void* p = nullptr;
    const size_t bytes = (n ? n : 1) * sizeof(T);
    const cudaError_t e = pool ? cudaMallocFromPoolAsync(&p, bytes, pool, st) : cudaMallocAsync(&p, bytes, st);
    if (e != cudaSuccess) { cudaGetLastError(); return nullptr; }
    ptrs.push_back(p);
    return reinterpret_cast<T*>(p);
  }
};

#define GG_PLAN_LAUNCH(name, kern, n, ...)                     \
  do {                                                         \
    GG_KERNEL_BEGIN(name, st);                                 \
    kern<<<grid_for(n), 256, 0, st>>>(__VA_ARGS__);            \
    GG_KERNEL_END(name, st);                                   \
  } while (0)

template <class K, class V>
static int sort_pairs(TempPool& tp, const K* kin, K* kout, const V* vin, V* vout, int64_t n, int end_bit, cudaStream_t st) {
  size_t bytes = 0;
  GG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, (int)n, 0, end_bit, st));
  void* tmp = tp.get<char>(bytes);
  if (!tmp) { set_error("gnnome_b200: plan: out of device memory (sort scratch)"); return GG_ERR_CUDA; }
  GG_KERNEL_BEGIN("plan_radix_sort", st);
  GG_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, kin, kout, vin, vout, (int)n, 0, end_bit, st));
  GG_KERNEL_END("plan_radix_sort", st);
  return GG_OK;
}

static int exclusive_sum(TempPool& tp, const int32_t* in, int32_t* out, int64_t n, cudaStream_t st) {
  size_t bytes = 0;
  GG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, st));
  void* tmp = tp.get<char>(bytes);
  if (!tmp) { set_error("gnnome_b200: plan: out of device memory (scan scratch)"); return GG_ERR_CUDA; }
  GG_KERNEL_BEGIN("plan_scan", st);
  GG_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, (int)n, st));
  GG_KERNEL_END("plan_scan", st);
  return GG_OK;
}

// src / dst: DEVICE int32[E] in caller edge order.  On success *out owns one device slab with every array.
int plan_create_device(const int32_t* src, const int32_t* dst, int64_t N, int64_t E, int flags, cudaStream_t st, Plan** out) {
  *out = nullptr;
  TempPool tp(st);
  Plan* pl = new Plan();
  pl->N = N; pl->E = E;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&pl->num_sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t e1 = (size_t)(E ? E : 1), n1 = (size_t)N + 1;
  const size_t words = 6 * e1 + 4 * n1;
  int32_t* slab = nullptr;
  cudaError_t ce = cudaMalloc((void**)&slab, words * sizeof(int32_t));
  if (ce != cudaSuccess) { delete pl; return cuda_fail(ce, "plan_create alloc"); }
  pl->host_slab = slab;
  struct Guard { Plan* p; bool armed = true; ~Guard() { if (armed) { cudaFree(p->host_slab); delete p; } } } guard{pl};
  size_t off = 0;
  auto carve = [&](size_t cap) { int32_t* p = slab + off; off += cap; return p; };
  pl->src = carve(e1); pl->dst = carve(e1); pl->out_eid = carve(e1); pl->out_dst = carve(e1);
  pl->perm = carve(e1); pl->inv_perm = carve(e1);
  pl->in_ptr = carve(n1); pl->out_ptr = carve(n1); pl->node_perm = carve(n1); pl->node_inv = carve(n1);

  int* err = tp.get<int>(4);                                  // [0] error flag, [1..2] sweep flags
  if (!err) return cuda_fail(cudaErrorMemoryAllocation, "plan scratch");
  GG_CUDA(cudaMemsetAsync(err, 0, 4 * sizeof(int), st));
  const bool relabel = (flags & GG_PLAN_RELABEL) && N > 0 && E > 0;
  int32_t* udeg = nullptr;
  if (relabel) {
    udeg = tp.get<int32_t>((size_t)N);
    if (!udeg) return cuda_fail(cudaErrorMemoryAllocation, "plan scratch");
    GG_CUDA(cudaMemsetAsync(udeg, 0, (size_t)N * sizeof(int32_t), st));
  }
  if (E >= (1LL << 30)) { set_error("gnnome_b200: plan_create: more than 2^30 edges"); return GG_ERR_UNSUPPORTED; }
  if (E > 0) GG_PLAN_LAUNCH("plan_check_degree_kernel", plan_check_degree_kernel, E, E, N, src, dst, udeg, err);
  // The ONE synchronisation of the builder, taken early: everything below indexes arrays with the node ids, so they
  // must be known to be in range first.  Nothing after this point waits for the device.
  {
    int herr = 0;
    GG_CUDA(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, st));
    GG_CUDA(cudaStreamSynchronize(st));
    if (herr) { set_error("gnnome_b200: plan_create: node index out of range"); return GG_ERR_ARG; }
  }

  // ------------------------------------------------------------------ node order
  if (!relabel) {
    if (N > 0) GG_PLAN_LAUNCH("plan_iota_kernel", plan_iota_kernel, N, N, pl->node_perm, pl->node_inv);
  } else {
    unsigned long long* key = tp.get<unsigned long long>((size_t)N);
    int32_t* is_seed = tp.get<int32_t>((size_t)N + 1);
    int32_t* seed_rank = tp.get<int32_t>((size_t)N + 1);       // exclusive scan of is_seed; [N] = number of regions
    unsigned long long* pairs = tp.get<unsigned long long>(2 * (size_t)E);
    unsigned long long* pairs_sorted = tp.get<unsigned long long>(2 * (size_t)E);
    int32_t* pflag = tp.get<int32_t>(2 * (size_t)E + 1);
    int32_t* ppos = tp.get<int32_t>(2 * (size_t)E + 1);        // [2E] = number of unique region pairs
    int32_t* radj_src = tp.get<int32_t>(2 * (size_t)E);
    int32_t* radj = tp.get<int32_t>(2 * (size_t)E);
    int32_t* rptr = tp.get<int32_t>((size_t)N + 2);
    int32_t* rqueue = tp.get<int32_t>((size_t)N);
    int32_t* rrank = tp.get<int32_t>((size_t)N);
    unsigned long long* nkey = tp.get<unsigned long long>((size_t)N);
    unsigned long long* nkey_sorted = tp.get<unsigned long long>((size_t)N);
    int32_t* ids = tp.get<int32_t>((size_t)N);
    if (!key || !is_seed || !seed_rank || !pairs || !pairs_sorted || !pflag || !ppos || !radj_src || !radj || !rptr || !rqueue ||
        !rrank || !nkey || !nkey_sorted || !ids)
      return cuda_fail(cudaErrorMemoryAllocation, "plan scratch");
    // 1. regions (cooperative kernel: one resident wave)
    {
      int per_sm = 0;
      GG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, plan_regions_kernel, 256, 0));
      if (per_sm < 1) per_sm = 1;
      if (per_sm > 4) per_sm = 4;
      int64_t blocks = (int64_t)pl->num_sms * per_sm;
      const int64_t need = (std::max<int64_t>(E, N) + 255) / 256;
      if (blocks > need) blocks = need;
      if (blocks < 1) blocks = 1;
      int* changed = err + 1;
      const int* err_c = err;
      void* args[] = {(void*)&E, (void*)&N, (void*)&src, (void*)&dst, (void*)&udeg, (void*)&key, (void*)&is_seed, (void*)&changed,
                      (void*)&err_c};
      GG_KERNEL_BEGIN("plan_regions_kernel", st);
      GG_CUDA(cudaLaunchCooperativeKernel((void*)plan_regions_kernel, dim3((unsigned)blocks), dim3(256), args, 0, st));
      GG_KERNEL_END("plan_regions_kernel", st);
    }
    GG_CUDA(cudaMemsetAsync(is_seed + N, 0, sizeof(int32_t), st));
    GG_TRY_RC(exclusive_sum(tp, is_seed, seed_rank, N + 1, st));
    // 2. region graph: unique cross-region pairs -> CSR
    GG_PLAN_LAUNCH("plan_region_pairs_kernel", plan_region_pairs_kernel, E, E, src, dst, key, seed_rank, pairs);
    {
      size_t bytes = 0;
      GG_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, pairs, pairs_sorted, (int)(2 * E), 0, 64, st));
      void* tmp = tp.get<char>(bytes);
      if (!tmp) return cuda_fail(cudaErrorMemoryAllocation, "plan scratch");
      GG_KERNEL_BEGIN("plan_radix_sort", st);
      GG_CUDA(cub::DeviceRadixSort::SortKeys(tmp, bytes, pairs, pairs_sorted, (int)(2 * E), 0, 64, st));
      GG_KERNEL_END("plan_radix_sort", st);
    }
    GG_PLAN_LAUNCH("plan_pair_flags_kernel", plan_pair_flags_kernel, 2 * E, 2 * E, pairs_sorted, pflag);
    GG_CUDA(cudaMemsetAsync(pflag + 2 * E, 0, sizeof(int32_t), st));
    GG_TRY_RC(exclusive_sum(tp, pflag, ppos, 2 * E + 1, st));
    GG_PLAN_LAUNCH("plan_pair_compact_kernel", plan_pair_compact_kernel, 2 * E, 2 * E, pairs_sorted, pflag, ppos, radj_src, radj);
    GG_PLAN_LAUNCH("plan_lower_bound_kernel", plan_lower_bound_kernel, N + 1, N, radj_src, 0, ppos + 2 * E, rptr);
    // 3. breadth-first order of the regions (one warp), then the nodes
    GG_KERNEL_BEGIN("plan_region_bfs_kernel", st);
    plan_region_bfs_kernel<<<1, 32, 0, st>>>(seed_rank + N, rptr, radj, rqueue, rrank);
    GG_KERNEL_END("plan_region_bfs_kernel", st);
    if (std::getenv("GG_PLAN_DEBUG")) {          // diagnostics (synchronises): region / region-edge counts, relaxation sweeps
      int h[3] = {0, 0, 0};
      cudaMemcpyAsync(&h[0], seed_rank + N, sizeof(int), cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(&h[1], ppos + 2 * E, sizeof(int), cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(&h[2], err + 3, sizeof(int), cudaMemcpyDeviceToHost, st);
      cudaStreamSynchronize(st);
      std::fprintf(stderr, "gg_plan_device: N=%lld E=%lld regions=%d region_edges=%d sweeps=%d\n", (long long)N, (long long)E, h[0], h[1], h[2]);
    }
    GG_PLAN_LAUNCH("plan_node_keys_kernel", plan_node_keys_kernel, N, N, udeg, key, seed_rank, rrank, nkey, ids);
    GG_TRY_RC(sort_pairs(tp, nkey, nkey_sorted, ids, pl->node_perm, N, 48, st));
    GG_PLAN_LAUNCH("plan_invert_kernel", plan_invert_kernel, N, N, pl->node_perm, pl->node_inv);
  }

  // ------------------------------------------------------------------ edge order + CSRs
  if (E > 0) {
    int32_t* k1 = tp.get<int32_t>((size_t)E);
    int32_t* v1 = tp.get<int32_t>((size_t)E);
    int32_t* src_sorted = tp.get<int32_t>((size_t)E);
    if (!k1 || !v1 || !src_sorted) return cuda_fail(cudaErrorMemoryAllocation, "plan scratch");
    const int nbits = bits_for(N);
    GG_PLAN_LAUNCH("plan_dst_keys_kernel", plan_dst_keys_kernel, E, E, dst, pl->node_inv, k1, v1);
    GG_TRY_RC(sort_pairs(tp, k1, pl->dst, v1, pl->perm, E, nbits, st));                 // stable: ties keep caller edge order
    GG_PLAN_LAUNCH("plan_internal_src_kernel", plan_internal_src_kernel, E, E, src, pl->node_inv, pl->perm, pl->src, pl->inv_perm, v1);
    GG_TRY_RC(sort_pairs(tp, pl->src, src_sorted, v1, pl->out_eid, E, nbits, st));      // stable: increasing internal id per source
    GG_PLAN_LAUNCH("plan_gather_kernel", plan_gather_kernel, E, E, pl->dst, pl->out_eid, pl->out_dst);
    GG_PLAN_LAUNCH("plan_lower_bound_kernel", plan_lower_bound_kernel, N + 1, N, pl->dst, E, nullptr, pl->in_ptr);
    GG_PLAN_LAUNCH("plan_lower_bound_kernel", plan_lower_bound_kernel, N + 1, N, src_sorted, E, nullptr, pl->out_ptr);
  } else {
    GG_CUDA(cudaMemsetAsync(pl->in_ptr, 0, n1 * sizeof(int32_t), st));
    GG_CUDA(cudaMemsetAsync(pl->out_ptr, 0, n1 * sizeof(int32_t), st));
  }
  guard.armed = false;
  *out = pl;
  return GG_OK;
}

}  // namespace gg
