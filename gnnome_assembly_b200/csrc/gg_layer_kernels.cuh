// gg_layer_kernels.cuh — the message-passing kernels of one GatedGCN layer (forward and backward).
//
// All kernels are node-centric: one warp owns one node and walks its in-edges (contiguous in the
// internal dst-sorted edge order, so E x d tensors are streamed exactly once, in order, in whole
// 128-byte lines) or its out-edges (through the out-edge CSR).  Per-node sums are accumulated in
// registers and written once: no atomics on the data path, deterministic summation order.
// Math being reproduced: layers/gated_gcn_full.py:120-154 (see SURVEY.md §8a "Exact forward spec").
#pragma once
#include "gg_common.cuh"

namespace gg {

constexpr int kNodeThreads = 256;   // 8 warps per CTA
// resident CTAs per SM the streaming kernels are compiled for (register cap 64 / 85 / 128 per thread): a kernel that
// drifts from 80 to 86 registers loses a third of its warps (edge_gate_fwd: 143 -> 167 us when that happened)
template <int D> constexpr int node_min_blocks() { return D <= 64 ? 4 : (D <= 128 ? 3 : 2); }

// ------------------------------------------------------------------ normalisation coefficients
// NORM == GG_NORM_BATCH: per-channel batch statistics from fp64 sums (biased variance, eps 1e-5)
// NORM == GG_NORM_LAYER: per-row statistics computed on the fly (whole row lives in one warp)
template <int D, int NORM>
struct Norm {
  static constexpr int VPL = D / 32;
  float mean[VPL], rstd[VPL], gamma[VPL], beta[VPL];

  __device__ __forceinline__ void init(const double* __restrict__ sums, double inv_count,
                                       const float* __restrict__ g, const float* __restrict__ b, int lane) {
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = Row<D>::channel(k, lane);
      gamma[k] = __ldg(g + c);
      beta[k] = __ldg(b + c);
      if constexpr (NORM == GG_NORM_BATCH) {
        const double m = sums[c] * inv_count;
        double var = sums[D + c] * inv_count - m * m;
        var = var > 0.0 ? var : 0.0;
        mean[k] = (float)m;
        rstd[k] = (float)(1.0 / sqrt(var + (double)kNormEps));
      } else {
        mean[k] = 0.f; rstd[k] = 1.f;
      }
    }
  }
  // x -> xhat (in place); for layer norm returns the row rstd (needed by the backward formulas)
  __device__ __forceinline__ float normalize(Row<D>& x) const {
    if constexpr (NORM == GG_NORM_BATCH) {
#pragma unroll
      for (int k = 0; k < VPL; ++k) x.v[k] = (x.v[k] - mean[k]) * rstd[k];
      return 1.f;
    } else {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) s += x.v[k];
      const float mu = warp_sum(s) * (1.0f / D);
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) { x.v[k] -= mu; q += x.v[k] * x.v[k]; }
      const float r = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + kNormEps);
#pragma unroll
      for (int k = 0; k < VPL; ++k) x.v[k] *= r;
      return r;
    }
  }
  // backward of the normalisation: g = dL/d(gamma*xhat+beta) -> dL/dx, given xhat.
  // Batch norm needs the global means m1 = mean(g), m2 = mean(g*xhat) per channel (passed in);
  // layer norm computes its row means here.
  __device__ __forceinline__ void backward(Row<D>& g, const Row<D>& xhat, float row_rstd,
                                           const float (&m1)[VPL], const float (&m2)[VPL]) const {
    if constexpr (NORM == GG_NORM_BATCH) {
#pragma unroll
      for (int k = 0; k < VPL; ++k) g.v[k] = gamma[k] * rstd[k] * (g.v[k] - m1[k] - xhat.v[k] * m2[k]);
    } else {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) { g.v[k] *= gamma[k]; a += g.v[k]; b += g.v[k] * xhat.v[k]; }
      a = warp_sum(a) * (1.0f / D);
      b = warp_sum(b) * (1.0f / D);
#pragma unroll
      for (int k = 0; k < VPL; ++k) g.v[k] = row_rstd * (g.v[k] - a - xhat.v[k] * b);
    }
  }
};

// Tried in round 2 and withdrawn: a CSR cursor with a two-node look-ahead (CSR range of the next two nodes and the first
// 32 indices of the next node fetched while the current node is processed, to take the ptr -> index -> row dependency
// chain off the critical path).  Same-box A/B on the bench graph (profiles/r2_ab_lookahead.txt): 1-2 % SLOWER in all
// four streaming kernels — other resident warps already hide that chain, the extra registers only add spills.
// per-thread fp64 column accumulators -> one fp64 atomic per channel per CTA
template <int D>
__device__ __forceinline__ void block_flush_stats(const double (&s1)[D / 32], const double (&s2)[D / 32],
                                                  double* __restrict__ g1, double* __restrict__ g2) {
  __shared__ double sh[2][kNodeThreads / 32][D];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < D / 32; ++k) {
    const int c = Row<D>::channel(k, lane);
    sh[0][warp][c] = s1[k];
    sh[1][warp][c] = s2[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += kNodeThreads) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int w = 0; w < kNodeThreads / 32; ++w) { a += sh[0][w][c]; b += sh[1][w][c]; }
    atomicAdd(g1 + c, a);
    atomicAdd(g2 + c, b);
  }
}

// =====================================================================================  FORWARD
// F3: per dst node v, over its in-edges i (s_i -> v):
//   n = norm_e(t_i); e_out_i = relu(n) + e_in_i; sigma = sigmoid(e_out_i)           (:122-127)
//   num_f[v] += sigma * A2h[s_i]; den_f[v] += sigma; hf = num_f / (den_f + 1e-6)      (:128-130)
template <int D, int NORM>
__global__ void __launch_bounds__(kNodeThreads, node_min_blocks<D>())
edge_gate_fwd_kernel(int64_t N, int64_t E, const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ src,
                     const float* __restrict__ t, const float* __restrict__ e_in, const float* __restrict__ P,
                     const double* __restrict__ stats, const float* __restrict__ gamma,
                     const float* __restrict__ beta, int residual, float* __restrict__ e_out,
                     float* __restrict__ agg, int rev) {
  constexpr int VPL = D / 32;
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * kNodeThreads + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * kNodeThreads) >> 5;
  Norm<D, NORM> nrm;
  nrm.init(stats, E > 0 ? 1.0 / (double)E : 0.0, gamma, beta, lane);
  float* hf = agg;
  float* invden_f = agg + 2 * N * D;
  float* sum_xhat = agg + 4 * N * D;      // sum over in-edges of xhat_e: lets the backward rebuild gB2h per node
  for (int64_t vi = gw; vi < N; vi += nw) {
    const int64_t v = rev ? N - 1 - vi : vi;     // zig-zag traversal, see layer_fwd_impl
    const int beg = __ldg(in_ptr + v), end = __ldg(in_ptr + v + 1);
    Row<D> num, den, sx;
    num.fill(0.f); den.fill(0.f); sx.fill(0.f);
    for (int base = beg; base < end; base += 32) {
      const int cnt = min(32, end - base);
      const int my_s = (lane < cnt) ? __ldg(src + base + lane) : 0;
      for (int j = 0; j < cnt; j += 2) {
        const bool two = (j + 1) < cnt;
        const int64_t i0 = base + j, i1 = two ? i0 + 1 : i0;
        const int64_t sa = __shfl_sync(0xffffffffu, my_s, j);
        const int64_t sb = __shfl_sync(0xffffffffu, my_s, two ? j + 1 : j);
        Row<D> x0, ein0, a20, x1, ein1, a21;
        x0.load_stream(t + i0 * D, lane);
        ein0.load_stream(e_in + i0 * D, lane);
        a20.load(P + sa * (5 * D) + D, lane);
        x1.load_stream(t + i1 * D, lane);
        ein1.load_stream(e_in + i1 * D, lane);
        a21.load(P + sb * (5 * D) + D, lane);
        nrm.normalize(x0);
        nrm.normalize(x1);
        if constexpr (NORM == GG_NORM_BATCH) {
#pragma unroll
          for (int k = 0; k < VPL; ++k) sx.v[k] += x0.v[k] + (two ? x1.v[k] : 0.f);
        }
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          const float nv = x0.v[k] * nrm.gamma[k] + nrm.beta[k];
          const float eo = fmaxf(nv, 0.f) + (residual ? ein0.v[k] : 0.f);
          x0.v[k] = eo;
          const float sg = sigmoidf_(eo);
          num.v[k] = fmaf(sg, a20.v[k], num.v[k]);
          den.v[k] += sg;
        }
        x0.store(e_out + i0 * D, lane);
        if (two) {
#pragma unroll
          for (int k = 0; k < VPL; ++k) {
            const float nv = x1.v[k] * nrm.gamma[k] + nrm.beta[k];
            const float eo = fmaxf(nv, 0.f) + (residual ? ein1.v[k] : 0.f);
            x1.v[k] = eo;
            const float sg = sigmoidf_(eo);
            num.v[k] = fmaf(sg, a21.v[k], num.v[k]);
            den.v[k] += sg;
          }
          x1.store(e_out + i1 * D, lane);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      den.v[k] = 1.0f / (den.v[k] + kAggEps);
      num.v[k] *= den.v[k];
    }
    num.store(hf + v * D, lane);
    den.store(invden_f + v * D, lane);
    if constexpr (NORM == GG_NORM_BATCH) sx.store(sum_xhat + v * D, lane);
  }
}

// F4: per src node u, over its out-edges i (u -> v_i):  (reverse-graph branch, :133-145)
//   num_b[u] += sigma_i * A3h[v_i]; den_b[u] += sigma_i; hb = num_b / (den_b + 1e-6)
//   z[u] = A1h[u] + hf[u] + hb[u]   and per-channel sum / sum of squares of z for bn_h
template <int D, int NORM>
__global__ void __launch_bounds__(kNodeThreads, node_min_blocks<D>())
node_agg_fwd_kernel(int64_t N, const int32_t* __restrict__ out_ptr, const int32_t* __restrict__ out_eid,
                    const int32_t* __restrict__ out_dst, const float* __restrict__ e_out,
                    const float* __restrict__ P, float* __restrict__ agg, float* __restrict__ z,
                    double* __restrict__ stats_h, int rev) {
  constexpr int VPL = D / 32;
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * kNodeThreads + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * kNodeThreads) >> 5;
  const float* hf = agg;
  float* hb = agg + N * D;
  float* invden_b = agg + 3 * N * D;
  double s1[VPL], s2[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) { s1[k] = 0.0; s2[k] = 0.0; }
  for (int64_t ui = gw; ui < N; ui += nw) {
    const int64_t u = rev ? N - 1 - ui : ui;     // zig-zag traversal, see layer_fwd_impl
    const int beg = __ldg(out_ptr + u), end = __ldg(out_ptr + u + 1);
    Row<D> num, den;
    num.fill(0.f); den.fill(0.f);
    for (int base = beg; base < end; base += 32) {
      const int cnt = min(32, end - base);
      const int my_i = (lane < cnt) ? __ldg(out_eid + base + lane) : 0;
      const int my_v = (lane < cnt) ? __ldg(out_dst + base + lane) : 0;
      for (int j = 0; j < cnt; j += 4) {
        // four out-edges in flight: all row gathers are issued before any math
        Row<D> eo[4], a3[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int jj = (j + q < cnt) ? j + q : j;
          const int64_t i = __shfl_sync(0xffffffffu, my_i, jj);
          const int64_t v = __shfl_sync(0xffffffffu, my_v, jj);
          eo[q].load(e_out + i * D, lane);
          a3[q].load(P + v * (5 * D) + 2 * D, lane);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (j + q < cnt) {
#pragma unroll
            for (int k = 0; k < VPL; ++k) {
              const float sg = sigmoidf_(eo[q].v[k]);
              num.v[k] = fmaf(sg, a3[q].v[k], num.v[k]);
              den.v[k] += sg;
            }
          }
        }
      }
    }
    Row<D> a1, f;
    a1.load(P + u * (5 * D), lane);
    f.load(hf + u * D, lane);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      den.v[k] = 1.0f / (den.v[k] + kAggEps);
      num.v[k] *= den.v[k];
      const float zz = a1.v[k] + f.v[k] + num.v[k];      // (A1h + h_forward) + h_backward, :145
      a1.v[k] = zz;
      s1[k] += (double)zz;
      s2[k] += (double)zz * (double)zz;
    }
    num.store(hb + u * D, lane);
    den.store(invden_b + u * D, lane);
    a1.store(z + u * D, lane);
  }
  if constexpr (NORM == GG_NORM_BATCH) block_flush_stats<D>(s1, s2, stats_h, stats_h + D);
}

// F5: h_out = relu(norm_h(z)) + h_in     (:147-152)
template <int D, int NORM>
__global__ void __launch_bounds__(kNodeThreads)
node_update_fwd_kernel(int64_t N, const float* __restrict__ z, const float* __restrict__ h_in,
                       const double* __restrict__ stats_h, const float* __restrict__ gamma,
                       const float* __restrict__ beta, int residual, float* __restrict__ h_out) {
  constexpr int VPL = D / 32;
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * kNodeThreads + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * kNodeThreads) >> 5;
  Norm<D, NORM> nrm;
  nrm.init(stats_h, N > 0 ? 1.0 / (double)N : 0.0, gamma, beta, lane);
  for (int64_t u = gw; u < N; u += nw) {
    Row<D> x, hin;
    x.load(z + u * D, lane);
    hin.load(h_in + u * D, lane);
    nrm.normalize(x);
#pragma unroll
    for (int k = 0; k < VPL; ++k)
      x.v[k] = fmaxf(x.v[k] * nrm.gamma[k] + nrm.beta[k], 0.f) + (residual ? hin.v[k] : 0.f);
    x.store(h_out + u * D, lane);
  }
}

// =====================================================================================  BACKWARD
// B1: g_y = g_h * [y > 0];  bstats_h = [sum g_y | sum g_y * xhat_h]   (= dbeta_h | dgamma_h)
template <int D, int NORM>
__global__ void __launch_bounds__(kNodeThreads)
node_bwd_reduce_kernel(int64_t N, const float* __restrict__ z, const float* __restrict__ g_h,
                       const double* __restrict__ stats_h, const float* __restrict__ gamma,
                       const float* __restrict__ beta, double* __restrict__ bstats_h) {
  constexpr int VPL = D / 32;
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * kNodeThreads + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * kNodeThreads) >> 5;
  Norm<D, NORM> nrm;
  nrm.init(stats_h, N > 0 ? 1.0 / (double)N : 0.0, gamma, beta, lane);
  double s1[VPL], s2[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) { s1[k] = 0.0; s2[k] = 0.0; }
  for (int64_t u = gw; u < N; u += nw) {
    Row<D> x, g;
    x.load(z + u * D, lane);
    g.load(g_h + u * D, lane);
    nrm.normalize(x);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const float y = x.v[k] * nrm.gamma[k] + nrm.beta[k];
      const float gy = y > 0.f ? g.v[k] : 0.f;
      s1[k] += (double)gy;
      s2[k] += (double)gy * (double)x.v[k];
    }
  }
  block_flush_stats<D>(s1, s2, bstats_h, bstats_h + D);
}

// B2: g_z = norm_h'(g_y);  gP[:, 0:d] = g_z (grad of A1h);
//     Gf[u] = [ g_z*invden_f | -g_z*invden_f*hf ],  Gb[u] = [ g_z*invden_b | -g_z*invden_b*hb ]
template <int D, int NORM>
__global__ void __launch_bounds__(kNodeThreads)
node_bwd_apply_kernel(int64_t N, const float* __restrict__ z, const float* __restrict__ g_h,
                      const double* __restrict__ stats_h, const double* __restrict__ bstats_h,
                      const float* __restrict__ gamma, const float* __restrict__ beta,
                      const float* __restrict__ agg, float* __restrict__ gP, float* __restrict__ G,
                      float* __restrict__ dgamma_h, float* __restrict__ dbeta_h) {
  constexpr int VPL = D / 32;
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * kNodeThreads + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * kNodeThreads) >> 5;
  // the affine gradients of bn_h are the finished fp64 sums of B1 (dbeta = sum g_y, dgamma = sum g_y xhat): written
  // here by one CTA instead of by a kernel of their own
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < D; c += kNodeThreads) {
      dbeta_h[c] = (float)bstats_h[c];
      dgamma_h[c] = (float)bstats_h[D + c];
    }
  }
  Norm<D, NORM> nrm;
  const double inv_n = N > 0 ? 1.0 / (double)N : 0.0;
  nrm.init(stats_h, inv_n, gamma, beta, lane);
  float m1[VPL], m2[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int c = Row<D>::channel(k, lane);
    m1[k] = (float)(bstats_h[c] * inv_n);
    m2[k] = (float)(bstats_h[D + c] * inv_n);
  }
  const float* hf = agg;
  const float* hb = agg + N * D;
  const float* idf = agg + 2 * N * D;
  const float* idb = agg + 3 * N * D;
  float* Gf = G;
  float* Gb = G + N * 2 * D;
  for (int64_t u = gw; u < N; u += nw) {
    Row<D> x, g;
    x.load(z + u * D, lane);
    g.load(g_h + u * D, lane);
    const float rr = nrm.normalize(x);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const float y = x.v[k] * nrm.gamma[k] + nrm.beta[k];
      g.v[k] = y > 0.f ? g.v[k] : 0.f;
    }
    nrm.backward(g, x, rr, m1, m2);           // g = dL/dz
    g.store(gP + u * (5 * D), lane);
    Row<D> a, b, o1, o2;
    a.load(idf + u * D, lane);
    b.load(hf + u * D, lane);
#pragma unroll
    for (int k = 0; k < VPL; ++k) { o1.v[k] = g.v[k] * a.v[k]; o2.v[k] = -o1.v[k] * b.v[k]; }
    o1.store(Gf + u * (2 * D), lane);
    o2.store(Gf + u * (2 * D) + D, lane);
    a.load(idb + u * D, lane);
    b.load(hb + u * D, lane);
#pragma unroll
    for (int k = 0; k < VPL; ++k) { o1.v[k] = g.v[k] * a.v[k]; o2.v[k] = -o1.v[k] * b.v[k]; }
    o1.store(Gb + u * (2 * D), lane);
    o2.store(Gb + u * (2 * D) + D, lane);
  }
}

// B3: edge pass A, per dst node v over in-edges i (s -> v):
//   recompute xhat_e, n, e_out, sigma from (t, e_in);
//   g_sigma = gnf[v]*A2h[s] + gdf[v] + gnb[s]*A3h[v] + gdb[s]
//   g_eo = g_e + g_sigma*sigma*(1-sigma)            -> stored
//   g_n  = g_eo*[n>0];  bstats_e += [g_n | g_n*xhat_e]
//   gA3h[v] = sum_i sigma_i * gnb[s_i]              -> gP[v, 2d:3d]
#ifndef GG_BWD_A_BLOCKS          // A/B builds: resident CTAs per SM edge_bwd_a is compiled for at d <= 128
#define GG_BWD_A_BLOCKS 2
#endif
template <int D, int NORM>
__global__ void __launch_bounds__(kNodeThreads, D <= 128 ? GG_BWD_A_BLOCKS : 1)
edge_bwd_a_kernel(int64_t N, int64_t E, const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ src,
                  const float* __restrict__ t, const float* __restrict__ e_in, const float* __restrict__ g_e,
                  const float* __restrict__ P, const float* __restrict__ G, const double* __restrict__ stats_e,
                  const float* __restrict__ gamma, const float* __restrict__ beta, int residual,
                  float* __restrict__ g_eo, float* __restrict__ gP, double* __restrict__ bstats_e, int rev) {
  constexpr int VPL = D / 32;
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * kNodeThreads + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * kNodeThreads) >> 5;
  Norm<D, NORM> nrm;
  nrm.init(stats_e, E > 0 ? 1.0 / (double)E : 0.0, gamma, beta, lane);
  const float* Gf = G;
  const float* Gb = G + N * 2 * D;
  // gradient statistics (dbeta_e, dgamma_e): fp32 per thread over its few nodes, fp64 across the grid
  float f1[VPL], f2[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) { f1[k] = 0.f; f2[k] = 0.f; }
  for (int64_t vi = gw; vi < N; vi += nw) {
    const int64_t v = rev ? N - 1 - vi : vi;     // zig-zag traversal, see layer_fwd_impl
    const int beg = __ldg(in_ptr + v), end = __ldg(in_ptr + v + 1);
    Row<D> gnf, gdf, a3, acc, sgn;
    acc.fill(0.f); sgn.fill(0.f);
    if (beg < end) {
      gnf.load(Gf + v * (2 * D), lane);
      gdf.load(Gf + v * (2 * D) + D, lane);
      a3.load(P + v * (5 * D) + 2 * D, lane);
    }
    auto edge_math = [&](Row<D>& x, const Row<D>& ein, Row<D>& ge, const Row<D>& gnb, const Row<D>& gdb,
                         const Row<D>& a2) {
      nrm.normalize(x);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const float nv = x.v[k] * nrm.gamma[k] + nrm.beta[k];
        const float eo = fmaxf(nv, 0.f) + (residual ? ein.v[k] : 0.f);
        const float sg = sigmoidf_(eo);
        const float gs = gnf.v[k] * a2.v[k] + gdf.v[k] + gnb.v[k] * a3.v[k] + gdb.v[k];
        const float geo = ge.v[k] + gs * sg * (1.0f - sg);
        ge.v[k] = geo;
        const float gn = nv > 0.f ? geo : 0.f;
        f1[k] += gn;
        f2[k] = fmaf(gn, x.v[k], f2[k]);
        sgn.v[k] += gn;
        acc.v[k] = fmaf(sg, gnb.v[k], acc.v[k]);
      }
    };
    for (int base = beg; base < end; base += 32) {
      const int cnt = min(32, end - base);
      const int my_s = (lane < cnt) ? __ldg(src + base + lane) : 0;
      for (int j = 0; j < cnt; j += 2) {
        const bool two = (j + 1) < cnt;
        const int64_t i0 = base + j, i1 = two ? i0 + 1 : i0;
        const int64_t sa = __shfl_sync(0xffffffffu, my_s, j);
        const int64_t sb = __shfl_sync(0xffffffffu, my_s, two ? j + 1 : j);
        Row<D> x0, ein0, ge0, gnb0, gdb0, a20, x1, ein1, ge1, gnb1, gdb1, a21;
        x0.load_stream(t + i0 * D, lane);
        ein0.load_stream(e_in + i0 * D, lane);
        if (g_e) ge0.load_stream(g_e + i0 * D, lane); else ge0.fill(0.f);
        gnb0.load(Gb + sa * (2 * D), lane);
        gdb0.load(Gb + sa * (2 * D) + D, lane);
        a20.load(P + sa * (5 * D) + D, lane);
        x1.load_stream(t + i1 * D, lane);
        ein1.load_stream(e_in + i1 * D, lane);
        if (g_e) ge1.load_stream(g_e + i1 * D, lane); else ge1.fill(0.f);
        gnb1.load(Gb + sb * (2 * D), lane);
        gdb1.load(Gb + sb * (2 * D) + D, lane);
        a21.load(P + sb * (5 * D) + D, lane);
        edge_math(x0, ein0, ge0, gnb0, gdb0, a20);
        ge0.store(g_eo + i0 * D, lane);
        if (two) {
          edge_math(x1, ein1, ge1, gnb1, gdb1, a21);
          ge1.store(g_eo + i1 * D, lane);
        }
      }
    }
    acc.store(gP + v * (5 * D) + 2 * D, lane);
    // sum over in-edges of g_n: with the per-node sum of xhat from the forward, gB2h[v] = sum_i g_t_i follows
    // without a pass over g_t (edge_bwd_src_kernel finishes it; edge_bwd_b_kernel overwrites it otherwise)
    sgn.store(gP + v * (5 * D) + 4 * D, lane);
  }
  double s1[VPL], s2[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) { s1[k] = (double)f1[k]; s2[k] = (double)f2[k]; }
  block_flush_stats<D>(s1, s2, bstats_e, bstats_e + D);
}

// B4: edge pass B, per dst node v over in-edges:
//   g_n = g_eo*[n>0]; g_t = norm_e'(g_n) -> stored;  gB2h[v] = sum_i g_t_i -> gP[v, 4d:5d]
template <int D, int NORM>
__global__ void __launch_bounds__(kNodeThreads)
edge_bwd_b_kernel(int64_t N, int64_t E, const int32_t* __restrict__ in_ptr, const float* __restrict__ t,
                  const float* __restrict__ g_eo, const double* __restrict__ stats_e,
                  const double* __restrict__ bstats_e, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float* __restrict__ g_t, float* __restrict__ gP) {
  constexpr int VPL = D / 32;
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * kNodeThreads + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * kNodeThreads) >> 5;
  Norm<D, NORM> nrm;
  const double inv_e = E > 0 ? 1.0 / (double)E : 0.0;
  nrm.init(stats_e, inv_e, gamma, beta, lane);
  float m1[VPL], m2[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int c = Row<D>::channel(k, lane);
    m1[k] = (float)(bstats_e[c] * inv_e);
    m2[k] = (float)(bstats_e[D + c] * inv_e);
  }
  for (int64_t v = gw; v < N; v += nw) {
    const int beg = __ldg(in_ptr + v), end = __ldg(in_ptr + v + 1);
    Row<D> acc;
    acc.fill(0.f);
    for (int64_t i = beg; i < end; i += 2) {
      const bool two = (i + 1) < end;
      const int64_t i1 = two ? i + 1 : i;
      Row<D> x0, g0, x1, g1;
      x0.load_stream(t + i * D, lane);
      g0.load_stream(g_eo + i * D, lane);
      x1.load_stream(t + i1 * D, lane);
      g1.load_stream(g_eo + i1 * D, lane);
      const float r0 = nrm.normalize(x0);
      const float r1 = nrm.normalize(x1);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        g0.v[k] = (x0.v[k] * nrm.gamma[k] + nrm.beta[k]) > 0.f ? g0.v[k] : 0.f;
        g1.v[k] = (x1.v[k] * nrm.gamma[k] + nrm.beta[k]) > 0.f ? g1.v[k] : 0.f;
      }
      nrm.backward(g0, x0, r0, m1, m2);
      g0.store(g_t + i * D, lane);
#pragma unroll
      for (int k = 0; k < VPL; ++k) acc.v[k] += g0.v[k];
      if (two) {
        nrm.backward(g1, x1, r1, m1, m2);
        g1.store(g_t + i1 * D, lane);
#pragma unroll
        for (int k = 0; k < VPL; ++k) acc.v[k] += g1.v[k];
      }
    }
    acc.store(gP + v * (5 * D) + 4 * D, lane);
  }
}

// B5: per src node u over its out-edges i (u -> v):
//   gB1h[u] = sum_i g_t_i -> gP[u, 3d:4d];   gA2h[u] = sum_i sigma_i * gnf[v_i] -> gP[u, d:2d]
// With fix != 0 (batch norm, g_t produced inside the bwd-data GEMM): gP[u, 4d:5d] arrives holding
// S = sum_{in(u)} g_n and is finished here as gB2h[u] = gamma rstd (S - indeg m1 - m2 sum_{in(u)} xhat).
template <int D>
__global__ void __launch_bounds__(kNodeThreads, node_min_blocks<D>())
edge_bwd_src_kernel(int64_t N, const int32_t* __restrict__ out_ptr, const int32_t* __restrict__ out_eid,
                    const int32_t* __restrict__ out_dst, const float* __restrict__ g_t,
                    const float* __restrict__ e_out, const float* __restrict__ G, float* __restrict__ gP,
                    int fix, int64_t E, const int32_t* __restrict__ in_ptr, const float* __restrict__ sum_xhat,
                    const double* __restrict__ stats_e, const double* __restrict__ bstats_e,
                    const float* __restrict__ gamma_e, float* __restrict__ dgamma_e, float* __restrict__ dbeta_e, int rev) {
  constexpr int VPL = D / 32;
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * kNodeThreads + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * kNodeThreads) >> 5;
  if (blockIdx.x == 0) {                  // affine gradients of bn_e from the finished sums of edge pass A
    for (int c = threadIdx.x; c < D; c += kNodeThreads) {
      dbeta_e[c] = (float)bstats_e[c];
      dgamma_e[c] = (float)bstats_e[D + c];
    }
  }
  float fscale[VPL], fm1[VPL], fm2[VPL];
  if (fix) {
    const double inv_e = E > 0 ? 1.0 / (double)E : 0.0;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = Row<D>::channel(k, lane);
      const double m = stats_e[c] * inv_e;
      double var = stats_e[D + c] * inv_e - m * m;
      var = var > 0.0 ? var : 0.0;
      fscale[k] = __ldg(gamma_e + c) * (float)(1.0 / sqrt(var + (double)kNormEps));
      fm1[k] = (float)(bstats_e[c] * inv_e);
      fm2[k] = (float)(bstats_e[D + c] * inv_e);
    }
  }
  const float* Gf = G;
  for (int64_t ui = gw; ui < N; ui += nw) {
    const int64_t u = rev ? N - 1 - ui : ui;     // zig-zag traversal, see layer_fwd_impl
    const int beg = __ldg(out_ptr + u), end = __ldg(out_ptr + u + 1);
    Row<D> acc1, acc2;
    acc1.fill(0.f); acc2.fill(0.f);
    for (int base = beg; base < end; base += 32) {
      const int cnt = min(32, end - base);
      const int my_i = (lane < cnt) ? __ldg(out_eid + base + lane) : 0;
      const int my_v = (lane < cnt) ? __ldg(out_dst + base + lane) : 0;
      for (int j = 0; j < cnt; j += 4) {
        Row<D> gt[4], eo[4], gnf[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int jj = (j + q < cnt) ? j + q : j;
          const int64_t i = __shfl_sync(0xffffffffu, my_i, jj);
          const int64_t v = __shfl_sync(0xffffffffu, my_v, jj);
          gt[q].load(g_t + i * D, lane);
          eo[q].load(e_out + i * D, lane);
          gnf[q].load(Gf + v * (2 * D), lane);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (j + q < cnt) {
#pragma unroll
            for (int k = 0; k < VPL; ++k) {
              acc1.v[k] += gt[q].v[k];
              acc2.v[k] = fmaf(sigmoidf_(eo[q].v[k]), gnf[q].v[k], acc2.v[k]);
            }
          }
        }
      }
    }
    acc1.store(gP + u * (5 * D) + 3 * D, lane);
    acc2.store(gP + u * (5 * D) + D, lane);
    if (fix) {
      Row<D> sg, sx;
      sg.load(gP + u * (5 * D) + 4 * D, lane);
      sx.load(sum_xhat + u * D, lane);
      const float indeg = (float)(__ldg(in_ptr + u + 1) - __ldg(in_ptr + u));
#pragma unroll
      for (int k = 0; k < VPL; ++k) sg.v[k] = fscale[k] * (sg.v[k] - indeg * fm1[k] - fm2[k] * sx.v[k]);
      sg.store(gP + u * (5 * D) + 4 * D, lane);
    }
  }
}

// segmented row sums over in- and out-edges for the predictor backward:
//   gQ[u, 0:W] = sum_{out-edges i of u} g[i, :]     gQ[v, W:2W] = sum_{in-edges i of v} g[i, :]
template <int W>
__global__ void __launch_bounds__(kNodeThreads)
edge_to_node_sums_kernel(int64_t N, const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ out_ptr,
                         const int32_t* __restrict__ out_eid, const float* __restrict__ g,
                         float* __restrict__ gQ) {
  constexpr int VPL = W / 32;
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * kNodeThreads + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * kNodeThreads) >> 5;
  for (int64_t u = gw; u < N; u += nw) {
    Row<W> acc;
    acc.fill(0.f);
    int beg = __ldg(out_ptr + u), end = __ldg(out_ptr + u + 1);
    for (int base = beg; base < end; base += 32) {
      const int cnt = min(32, end - base);
      const int my_i = (lane < cnt) ? __ldg(out_eid + base + lane) : 0;
      for (int j = 0; j < cnt; ++j) {
        const int64_t i = __shfl_sync(0xffffffffu, my_i, j);
        Row<W> r;
        r.load(g + i * W, lane);
#pragma unroll
        for (int k = 0; k < VPL; ++k) acc.v[k] += r.v[k];
      }
    }
    acc.store(gQ + u * (2 * W), lane);
    acc.fill(0.f);
    beg = __ldg(in_ptr + u); end = __ldg(in_ptr + u + 1);
    for (int64_t i = beg; i < end; ++i) {
      Row<W> r;
      r.load(g + i * W, lane);
#pragma unroll
      for (int k = 0; k < VPL; ++k) acc.v[k] += r.v[k];
    }
    acc.store(gQ + u * (2 * W) + W, lane);
  }
}

// dgamma = S2, dbeta = S1 (fp64 sums -> fp32 outputs)
__global__ void affine_grads_kernel(int d, const double* __restrict__ bstats, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < d) {
    dbeta[c] = (float)bstats[c];
    dgamma[c] = (float)bstats[d + c];
  }
}

}  // namespace gg
