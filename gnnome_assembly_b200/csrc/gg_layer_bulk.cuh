// gg_layer_bulk.cuh — dst-ordered edge kernels with the streamed E x d operands staged through shared memory by the
// bulk-copy engine (cp.async.bulk global -> shared, completion on an mbarrier).
//
// In the internal edge order the in-edges of consecutive nodes are consecutive rows, so what a CTA reads of an
// E x d tensor is ONE contiguous byte range.  The register-staged kernels (gg_layer_kernels.cuh) keep two edges
// per warp in flight; their bytes in flight are bounded by registers (edge_bwd_a sits at 128 registers, 16 warps
// per SM), and ncu shows them at 56-63 % of DRAM bandwidth with half the issue slots idle.  Here a producer warp
// streams the CTA's range in chunks of kChunk rows per operand into a ring of shared-memory stages (up to ~190 KB
// in flight per SM, no registers), and the 8 consumer warps — one node each, as before — read whole 512-byte rows
// from shared memory (conflict free).  Gathered node rows (A2h[src] ...) still come straight from L2.
//
// Synchronisation: full[s] (1 arrival + transaction bytes) is completed by the copies; empty[s] counts one
// arrival per consumer warp: a warp releases chunk c once its next edge lies beyond it (edges are visited in
// increasing order), whether or not it owned a row of c — but only after c has landed, so that its arrival for
// chunk c + S can never fall into the phase of chunk c.  A warp releases every older chunk before it waits on a
// newer one, so the producer (which needs all 8 releases of chunk q - S to refill its stage) can always serve the
// slowest warp: no cycle.
#pragma once
#include "gg_common.cuh"
#include "gg_gemm_tc.cuh"
#include "gg_layer_kernels.cuh"

namespace gg {

constexpr int kBulkConsumers = 8;                          // consumer warps per CTA (one node each)
constexpr int kBulkThreads = 32 * (kBulkConsumers + 1);    // + the producer warp
constexpr int kBulkNodes = 16;                             // nodes per work block (2 per consumer warp): ~10 blocks per CTA
constexpr int kBulkStageBytes = 16 * 1024;                 // per stage, all streamed operands together

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cnt(uint32_t bar) { tc::mbar_arrive(bar); }

// ring bookkeeping shared by producer and consumers: chunk sequence number q -> (stage, phase)
template <int S>
struct Ring {
  uint32_t full0, empty0;                                  // shared addresses of full[0], empty[0]
  __device__ __forceinline__ uint32_t full(int q) const { return full0 + 8u * (q % S); }
  __device__ __forceinline__ uint32_t empty(int q) const { return empty0 + 8u * (q % S); }
  __device__ __forceinline__ uint32_t phase(int q) const { return (uint32_t)((q / S) & 1); }
};

// F3 (edge_gate_fwd_kernel) with t and e_in staged by bulk copies.  Same arithmetic, same per-node order.
template <int D, int NORM>
__global__ void __launch_bounds__(kBulkThreads, 2)
edge_gate_fwd_bulk_kernel(int64_t N, int64_t E, const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ src,
                          const float* __restrict__ t, const float* __restrict__ e_in, const float* __restrict__ P,
                          const double* __restrict__ stats, const float* __restrict__ gamma,
                          const float* __restrict__ beta, int residual, float* __restrict__ e_out,
                          float* __restrict__ agg) {
  constexpr int VPL = D / 32;
  constexpr int kOps = 2;                                  // streamed operands: t, e_in
  constexpr int CH = kBulkStageBytes / (kOps * D * 4);     // rows per chunk: 16 at d = 128
  constexpr int S = 6;                                     // stages: 96 KB per CTA, two CTAs per SM
  extern __shared__ __align__(128) uint8_t bulk_smem[];
  __shared__ __align__(8) uint64_t bars[2 * S];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t smem0 = tc::smem_u32(bulk_smem);
  Ring<S> ring{tc::smem_u32(bars), tc::smem_u32(bars + S)};
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { tc::mbar_init(ring.full0 + 8u * s, 1); tc::mbar_init(ring.empty0 + 8u * s, kBulkConsumers); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t nblocks = (N + kBulkNodes - 1) / kBulkNodes;

  if (warp == kBulkConsumers) {
    // ================================================================ producer (one lane)
    if (lane == 0) {
      int q = 0;                                           // chunk sequence number of this CTA
      for (int64_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const int64_t v0 = b * kBulkNodes, v1 = min(v0 + (int64_t)kBulkNodes, N);
        const int eb = __ldg(in_ptr + v0), ee = __ldg(in_ptr + v1);
        for (int c0 = eb; c0 < ee; c0 += CH, ++q) {
          const int rows = min(CH, ee - c0);
          if (q >= S) tc::mbar_wait(ring.empty(q), ring.phase(q) ^ 1u);
          const uint32_t st = smem0 + (uint32_t)(q % S) * kBulkStageBytes;
          const uint32_t bytes = (uint32_t)rows * D * 4;
          tc::mbar_expect_tx(ring.full(q), kOps * bytes);
          bulk_load(st, t + (int64_t)c0 * D, bytes, ring.full(q));
          bulk_load(st + CH * D * 4, e_in + (int64_t)c0 * D, bytes, ring.full(q));
        }
      }
    }
    return;
  }

  // ================================================================ consumers: warp = node
  Norm<D, NORM> nrm;
  nrm.init(stats, E > 0 ? 1.0 / (double)E : 0.0, gamma, beta, lane);
  float* hf = agg;
  float* invden_f = agg + 2 * N * D;
  float* sum_xhat = agg + 4 * N * D;
  auto smem_row = [&](Row<D>& r, int q, int op, int row) {
    const float* p = reinterpret_cast<const float*>(bulk_smem + (size_t)(q % S) * kBulkStageBytes + (size_t)op * CH * D * 4) + row * D;
    if constexpr (D == 64) {
      const float2 x = reinterpret_cast<const float2*>(p)[lane];
      r.v[0] = x.x; r.v[1] = x.y;
    } else {
#pragma unroll
      for (int j = 0; j < D / 128; ++j) {
        const float4 x = reinterpret_cast<const float4*>(p + 128 * j)[lane];
        r.v[4 * j] = x.x; r.v[4 * j + 1] = x.y; r.v[4 * j + 2] = x.z; r.v[4 * j + 3] = x.w;
      }
    }
  };
  int qbase = 0;                                           // sequence number of the current block's first chunk
  for (int64_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
    const int64_t v0 = b * kBulkNodes, v1 = min(v0 + (int64_t)kBulkNodes, N);
    const int eb = __ldg(in_ptr + v0), ee = __ldg(in_ptr + v1);
    const int nchunks = (ee - eb + CH - 1) / CH;
    int released = 0, ready = 0;                           // chunks of this block released / known to have landed
    auto release_to = [&](int c) {                         // this warp is done with every chunk < c
      if (c > released) {
        __syncwarp();
        for (int k = released; k < c; ++k) {
          // a chunk is released only once it has LANDED (even if this warp owns no row of it): stage k % S is
          // refilled for chunk k only after chunk k - S was released by all warps, so this wait keeps a warp
          // that skips ahead from arriving twice in one phase of the same barrier
          if (k >= ready) tc::mbar_wait(ring.full(qbase + k), ring.phase(qbase + k));
          if (lane == 0) mbar_arrive_cnt(ring.empty(qbase + k));
        }
        if (c > ready) ready = c;
        released = c;
      }
    };
    for (int64_t v = v0 + warp; v < v1; v += kBulkConsumers) {
      const int beg = __ldg(in_ptr + v), end = __ldg(in_ptr + v + 1);
      Row<D> num, den, sx;
      num.fill(0.f); den.fill(0.f); sx.fill(0.f);
      for (int base = beg; base < end; base += 32) {
        const int cnt = min(32, end - base);
        const int my_s = (lane < cnt) ? __ldg(src + base + lane) : 0;
        for (int j = 0; j < cnt; j += 4) {
          // the gathered rows of four edges are requested together (L2 latency), then the four edges are served
          // from shared memory one after the other
          Row<D> a2[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int64_t sa = __shfl_sync(0xffffffffu, my_s, min(j + u, cnt - 1));
            a2[u].load(P + sa * (5 * D) + D, lane);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (j + u < cnt) {
              const int i = base + j + u;
              const int c = (i - eb) / CH, r = (i - eb) - c * CH;
              release_to(c);
              if (c >= ready) { tc::mbar_wait(ring.full(qbase + c), ring.phase(qbase + c)); ready = c + 1; }
              Row<D> x, ein;
              smem_row(x, qbase + c, 0, r);
              smem_row(ein, qbase + c, 1, r);
              nrm.normalize(x);
              if constexpr (NORM == GG_NORM_BATCH) {
#pragma unroll
                for (int k = 0; k < VPL; ++k) sx.v[k] += x.v[k];
              }
#pragma unroll
              for (int k = 0; k < VPL; ++k) {
                const float nv = x.v[k] * nrm.gamma[k] + nrm.beta[k];
                const float eo = fmaxf(nv, 0.f) + (residual ? ein.v[k] : 0.f);
                x.v[k] = eo;
                const float sg = sigmoidf_(eo);
                num.v[k] = fmaf(sg, a2[u].v[k], num.v[k]);
                den.v[k] += sg;
              }
              x.store(e_out + (int64_t)i * D, lane);
            }
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
    release_to(nchunks);
    qbase += nchunks;
  }
}

}  // namespace gg
