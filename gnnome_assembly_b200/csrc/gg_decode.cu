// gg_decode.cu — greedy contig decoding on the device.
//
// The reference decodes one contig per iteration (inference.py:182-259, get_contigs): it samples nb_paths
// (50) start edges, runs a greedy forward walk from the edge's head and a greedy backward walk from its tail
// for each of them (walk_forwards / walk_backwards, inference.py:31-77) ONE AFTER THE OTHER in Python,
// keeps the walk that reconstructs the longest sequence (get_contig_length, :20-28), marks it and the nodes
// it jumps over (:227-234) as visited, and repeats.  A walk is an inherently sequential pointer chase, but
// the nb_paths walks of an iteration are independent: here one warp owns one walk (lanes evaluate the
// candidate neighbours of the current node in parallel: visited bits, edge score, warp arg-max), all walks
// of the iteration run concurrently, the sequence length is accumulated on the way, and the visited sets
// are bitmaps (one private bitmap per walk, one shared bitmap for the nodes of finished contigs).
// Everything is integer / comparison work: results are bit-identical to the reference given the same start
// edges (ties in the arg-max go to the first neighbour in list order).
#include <limits.h>

#include <algorithm>

#include "gg_common.cuh"

namespace gg {

struct Adj {                      // adjacency in CALLER node ids (the n ^ 1 strand pairing lives there)
  const int32_t* ptr;             // [N+1]
  const int32_t* node;            // [E] neighbour, in the reference's list order
  const int32_t* eid;             // [E] caller edge id of (current -> neighbour) resp. (neighbour -> current)
};

__device__ __forceinline__ bool bit_get(const uint32_t* __restrict__ bm, int i) { return (__ldg(bm + (i >> 5)) >> (i & 31)) & 1u; }
// the walk's own bitmap is written by lane 0 while the walk runs: read it through L2, never the read-only path
__device__ __forceinline__ bool bit_get_rw(const uint32_t* bm, int i) { return (__ldcg(bm + (i >> 5)) >> (i & 31)) & 1u; }

// one greedy walk (inference.py:31-54 forwards with adj = successors, :57-77 backwards with adj = predecessors);
// returns the number of nodes written; nodes go to buf[pos0], buf[pos0 + dir], ...
__device__ int greedy_walk(int start, const Adj adj, const float* __restrict__ scores,
                           const int64_t* __restrict__ prefix_length, const uint32_t* __restrict__ visited_old,
                           uint32_t* loc, int* buf, int pos0, int dir, int max_len, int64_t& total, int& last,
                           int* err) {
  const int lane = threadIdx.x & 31;
  int cur = start, len = 0;
  while (true) {
    if (len >= max_len) {                      // a cycle of single-neighbour nodes: the reference would never return
      if (lane == 0) atomicOr(err, 1);
      break;
    }
    if (lane == 0) {
      buf[pos0 + dir * len] = cur;
      const int mate = cur ^ 1;                // same 32-bit word: 2k and 2k + 1
      __stcg(loc + (cur >> 5), __ldcg(loc + (cur >> 5)) | (1u << (cur & 31)) | (1u << (mate & 31)));
    }
    __syncwarp();
    ++len;
    last = cur;
    const int beg = __ldg(adj.ptr + cur), end = __ldg(adj.ptr + cur + 1);
    if (end == beg) break;                                          // :40-41
    if (end - beg == 1) {                                           // :42-44 (no visited check on purpose)
      total += prefix_length[__ldg(adj.eid + beg)];
      cur = __ldg(adj.node + beg);
      continue;
    }
    float best = -INFINITY;
    int best_pos = INT_MAX;
    for (int base = beg; base < end; base += 32) {                  // :45-51
      const int i = base + lane;
      if (i < end) {
        const int nb = __ldg(adj.node + i);
        if (!bit_get(visited_old, nb) && !bit_get_rw(loc, nb)) {
          const float s = scores[__ldg(adj.eid + i)];
          if (best_pos == INT_MAX || s > best) { best = s; best_pos = i; }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int op = __shfl_xor_sync(0xffffffffu, best_pos, o);
      if (op != INT_MAX && (best_pos == INT_MAX || ob > best || (ob == best && op < best_pos))) { best = ob; best_pos = op; }
    }
    if (best_pos == INT_MAX) break;                                 // :48-49
    total += prefix_length[__ldg(adj.eid + best_pos)];
    cur = __ldg(adj.node + best_pos);
  }
  return len;
}

// one warp per start edge: forward walk from its head, then backward walk from its tail (inference.py:236-241)
__global__ void __launch_bounds__(32)
decode_walks_kernel(int N, Adj succ, Adj pred, const float* __restrict__ scores,
                    const int64_t* __restrict__ prefix_length, const int64_t* __restrict__ read_length,
                    const uint32_t* __restrict__ visited, const int32_t* __restrict__ start_src,
                    const int32_t* __restrict__ start_dst, const int32_t* __restrict__ start_eid, int words,
                    uint32_t* __restrict__ local_visited, int32_t* __restrict__ walk_buf,
                    int32_t* __restrict__ out_beg, int32_t* __restrict__ out_len, int64_t* __restrict__ out_seq_len,
                    int* __restrict__ err) {
  const int w = blockIdx.x;
  uint32_t* loc = local_visited + (size_t)w * words;
  int32_t* buf = walk_buf + (size_t)w * 2 * N;
  int64_t total = 0;
  int last_f = 0, last_b = 0;
  const int len_f = greedy_walk(start_dst[w], succ, scores, prefix_length, visited, loc, buf, N, +1, N, total, last_f, err);
  const int len_b = greedy_walk(start_src[w], pred, scores, prefix_length, visited, loc, buf, N - 1, -1, N, total, last_b, err);
  if (threadIdx.x == 0) {
    total += prefix_length[start_eid[w]] + read_length[last_f];    // get_contig_length, :20-28
    out_beg[w] = N - len_b;
    out_len[w] = len_b + len_f;
    out_seq_len[w] = total;
  }
}

// after the best walk is chosen (inference.py:223-234,241): nodes jumped over by each step ss -> dd
// (succs[ss] & preds[dd]) and their strand mates, plus the walk's own visited set, join `visited`
__global__ void decode_commit_kernel(int len, const int32_t* __restrict__ walk, Adj succ, Adj pred, int words,
                                     const uint32_t* __restrict__ loc, uint32_t* __restrict__ visited) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int i = tid; i + 1 < len; i += nth) {
    const int ss = walk[i], dd = walk[i + 1];
    const int pb = __ldg(pred.ptr + dd), pe = __ldg(pred.ptr + dd + 1);
    for (int k = __ldg(succ.ptr + ss); k < __ldg(succ.ptr + ss + 1); ++k) {
      const int t = __ldg(succ.node + k);
      bool hit = false;
      for (int q = pb; q < pe && !hit; ++q) hit = __ldg(pred.node + q) == t;
      if (hit) {
        atomicOr(visited + (t >> 5), 1u << (t & 31));
        atomicOr(visited + ((t ^ 1) >> 5), 1u << ((t ^ 1) & 31));
      }
    }
  }
  for (int wd = tid; wd < words; wd += nth) {
    const uint32_t v = loc[wd];
    if (v) atomicOr(visited + wd, v);
  }
}

// sampling weights of inference.py:279-286 on the remaining graph (:262-275 + dgl.remove_self_loop, :187):
// p = max(sigmoid(score), 1e-9) for an edge whose two ends are unvisited, 0 otherwise (normalisation is the sampler's)
__global__ void decode_edge_weights_kernel(int64_t E, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                           const float* __restrict__ scores, const uint32_t* __restrict__ visited,
                                           float* __restrict__ weights) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) {
    const int s = src[i], d = dst[i];
    float p = 0.f;
    if (s != d && !bit_get(visited, s) && !bit_get(visited, d)) {
      p = 1.0f / (1.0f + expf(-scores[i]));
      p = p < 1e-9f ? 1e-9f : p;
    }
    weights[i] = p;
  }
}

}  // namespace gg

using namespace gg;

extern "C" {

int gg_decode_walks(int64_t N, const int32_t* succ_ptr, const int32_t* succ_node, const int32_t* succ_eid,
                    const int32_t* pred_ptr, const int32_t* pred_node, const int32_t* pred_eid, const float* scores,
                    const int64_t* prefix_length, const int64_t* read_length, const uint32_t* visited, int n_walks,
                    const int32_t* start_src, const int32_t* start_dst, const int32_t* start_eid,
                    uint32_t* local_visited, int32_t* walk_buf, int32_t* out_beg, int32_t* out_len,
                    int64_t* out_seq_len, int* err, void* stream) {
  GG_REQUIRE(N > 0 && N < (1LL << 30), "decode_walks: bad node count");
  GG_REQUIRE(n_walks >= 0, "decode_walks: negative walk count");
  if (n_walks == 0) return GG_OK;
  GG_REQUIRE(succ_ptr && succ_node && succ_eid && pred_ptr && pred_node && pred_eid, "decode_walks: null adjacency");
  GG_REQUIRE(scores && prefix_length && read_length && visited, "decode_walks: null graph data");
  GG_REQUIRE(start_src && start_dst && start_eid, "decode_walks: null start edges");
  GG_REQUIRE(local_visited && walk_buf && out_beg && out_len && out_seq_len && err, "decode_walks: null output");
  cudaStream_t st = (cudaStream_t)stream;
  const int words = (int)((N + 31) / 32);
  GG_CUDA(cudaMemsetAsync(local_visited, 0, sizeof(uint32_t) * (size_t)words * n_walks, st));
  GG_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
  GG_KERNEL_BEGIN("decode_walks_kernel", st);
  decode_walks_kernel<<<n_walks, 32, 0, st>>>((int)N, Adj{succ_ptr, succ_node, succ_eid}, Adj{pred_ptr, pred_node, pred_eid},
                                             scores, prefix_length, read_length, visited, start_src, start_dst, start_eid,
                                             words, local_visited, walk_buf, out_beg, out_len, out_seq_len, err);
  GG_KERNEL_END("decode_walks_kernel", st);
  return GG_OK;
}

int gg_decode_commit(int64_t N, const int32_t* succ_ptr, const int32_t* succ_node, const int32_t* pred_ptr,
                     const int32_t* pred_node, const int32_t* walk, int len, const uint32_t* walk_visited,
                     uint32_t* visited, void* stream) {
  GG_REQUIRE(N > 0 && len >= 0, "decode_commit: bad sizes");
  GG_REQUIRE(succ_ptr && succ_node && pred_ptr && pred_node && walk_visited && visited, "decode_commit: null pointer");
  GG_REQUIRE(len == 0 || walk, "decode_commit: null walk");
  cudaStream_t st = (cudaStream_t)stream;
  const int words = (int)((N + 31) / 32);
  int blocks = (std::max(len, words) + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  GG_KERNEL_BEGIN("decode_commit_kernel", st);
  decode_commit_kernel<<<blocks, 256, 0, st>>>(len, walk, Adj{succ_ptr, succ_node, nullptr}, Adj{pred_ptr, pred_node, nullptr},
                                              words, walk_visited, visited);
  GG_KERNEL_END("decode_commit_kernel", st);
  return GG_OK;
}

int gg_decode_edge_weights(int64_t E, const int32_t* src, const int32_t* dst, const float* scores,
                           const uint32_t* visited, float* weights, void* stream) {
  GG_REQUIRE(E >= 0, "decode_edge_weights: negative size");
  if (E == 0) return GG_OK;
  GG_REQUIRE(src && dst && scores && visited && weights, "decode_edge_weights: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = (E + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  GG_KERNEL_BEGIN("decode_edge_weights_kernel", st);
  decode_edge_weights_kernel<<<(unsigned)blocks, 256, 0, st>>>(E, src, dst, scores, visited, weights);
  GG_KERNEL_END("decode_edge_weights_kernel", st);
  return GG_OK;
}

}  // extern "C"
