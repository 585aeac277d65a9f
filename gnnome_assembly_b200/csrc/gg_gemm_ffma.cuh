// gg_gemm_ffma.cuh — true-fp32 (FFMA) tiled GEMM with pluggable epilogues.
//
// C[m,n] = sum_k A(m,k) * B(k,n), fp32 accumulate, 128 x BN x 16 tiles, 256 threads, 8 x (BN/16)
// register micro-tile, double-buffered shared memory with register prefetch.
//   A_T = false : A stored [M, K] row-major (lda)         A_T = true : A stored [K, M] (lda)   (weight grads)
//   B_T = true  : B stored [N, K] row-major (nn.Linear W) B_T = false: B stored [K, N] (ldb)
// The reference runs these projections as cuBLAS sgemm in true fp32 (torch allow_tf32=False,
// layers/gated_gcn_full.py:44,107-113); plain TF32 misses the 1e-4 logit tolerance,
// so this kernel is the exact-fp32 path for every hidden size; d >= 128 additionally has the
// tcgen05 3xTF32 path (gg_gemm_tc.cuh) where enabled.
//
// Epilogue functor contract:   template<int GW> __device__ void apply(int64_t m, int n, float (&acc)[GW], bool valid)
// is called by ALL threads uniformly (so it may use warp shuffles among the 16 lanes that share a row);
// it must write its own outputs when `valid`, and leave the final values in acc[] (used for column stats).
#pragma once
#include <type_traits>

#include "gg_common.cuh"

namespace gg {

struct GemmArgs {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  int64_t M; int N; int64_t K;
  int n_tiles;       // ceil(N / BN)
  int splits;        // split-K factor (gridDim.y)
  int64_t k_chunk;   // multiple of 16
  double* col_stats; // [2N]: sum, sum of squares of the epilogue output per column (kStats)
  float* bias_grad;  // [M]: sum over k of A(m,k)  (kBiasGrad, A_T only)
};

template <int BN> struct TileCfg;
template <> struct TileCfg<128> { static constexpr int NG = 2, GW = 4; };
template <> struct TileCfg<64>  { static constexpr int NG = 1, GW = 4; };
template <> struct TileCfg<32>  { static constexpr int NG = 1, GW = 2; };

constexpr int kBM = 128, kBK = 16, kGemmThreads = 256;

template <int BN, bool A_T, bool B_T, bool kStats, bool kBiasGrad, class Epi>
__global__ void __launch_bounds__(kGemmThreads, 2) gemm_ffma_kernel(GemmArgs g, Epi epi) {
  using Cfg = TileCfg<BN>;
  constexpr int NG = Cfg::NG, GW = Cfg::GW, TN = NG * GW;
  constexpr int LDA_S = kBM + 4, LDB_S = BN + 4;
  constexpr int A_STAGE = kBK * LDA_S, B_STAGE = kBK * LDB_S;
  constexpr int SMEM_FLOATS = 2 * (A_STAGE + B_STAGE);
  constexpr int STATS_FLOATS = kStats ? (2 * 16 * BN * 2) : 0;   // doubles counted as 2 floats
  __shared__ __align__(16) float smem[SMEM_FLOATS > STATS_FLOATS ? SMEM_FLOATS : STATS_FLOATS];
  float* As = smem;
  float* Bs = smem + 2 * A_STAGE;

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n_tile = blockIdx.x % g.n_tiles;
  const int64_t m_tile = blockIdx.x / g.n_tiles;
  const int64_t m0 = m_tile * kBM;
  const int n0 = n_tile * BN;
  const int64_t kbeg = (int64_t)blockIdx.y * g.k_chunk;
  const int64_t kend = (kbeg + g.k_chunk < g.K) ? (kbeg + g.k_chunk) : g.K;

  constexpr int A_F4 = kBM * kBK / 4, A_IT = A_F4 / kGemmThreads;              // 512 / 256 = 2
  constexpr int B_F4 = BN * kBK / 4, B_IT = (B_F4 + kGemmThreads - 1) / kGemmThreads;
  float4 ra[A_IT], rb[B_IT];

  auto load_tiles = [&](int64_t k0) {
#pragma unroll
    for (int it = 0; it < A_IT; ++it) {
      const int idx = tid + it * kGemmThreads;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (!A_T) {
        const int row = idx >> 2, k4 = idx & 3;
        const int64_t m = m0 + row, k = k0 + 4 * k4;
        if (m < g.M && k < kend) v = __ldg(reinterpret_cast<const float4*>(g.A + m * g.lda + k));
      } else {
        const int kk = idx >> 5, m4 = idx & 31;
        const int64_t m = m0 + 4 * m4, k = k0 + kk;
        if (m < g.M && k < kend) v = __ldg(reinterpret_cast<const float4*>(g.A + k * g.lda + m));
      }
      ra[it] = v;
    }
#pragma unroll
    for (int it = 0; it < B_IT; ++it) {
      const int idx = tid + it * kGemmThreads;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < B_F4) {
        if constexpr (B_T) {
          const int row = idx >> 2, k4 = idx & 3;
          const int n = n0 + row;
          const int64_t k = k0 + 4 * k4;
          if (n < g.N && k < kend) v = __ldg(reinterpret_cast<const float4*>(g.B + (int64_t)n * g.ldb + k));
        } else {
          const int kk = idx / (BN / 4), n4 = idx % (BN / 4);
          const int n = n0 + 4 * n4;
          const int64_t k = k0 + kk;
          if (n < g.N && k < kend) v = __ldg(reinterpret_cast<const float4*>(g.B + k * g.ldb + n));
        }
      }
      rb[it] = v;
    }
  };
  auto store_tiles = [&](int buf) {
    float* as = As + buf * A_STAGE;
    float* bs = Bs + buf * B_STAGE;
#pragma unroll
    for (int it = 0; it < A_IT; ++it) {
      const int idx = tid + it * kGemmThreads;
      if constexpr (!A_T) {
        const int row = idx >> 2, k4 = idx & 3;
        as[(4 * k4 + 0) * LDA_S + row] = ra[it].x;
        as[(4 * k4 + 1) * LDA_S + row] = ra[it].y;
        as[(4 * k4 + 2) * LDA_S + row] = ra[it].z;
        as[(4 * k4 + 3) * LDA_S + row] = ra[it].w;
      } else {
        const int kk = idx >> 5, m4 = idx & 31;
        *reinterpret_cast<float4*>(as + kk * LDA_S + 4 * m4) = ra[it];
      }
    }
#pragma unroll
    for (int it = 0; it < B_IT; ++it) {
      const int idx = tid + it * kGemmThreads;
      if (idx < B_F4) {
        if constexpr (B_T) {
          const int row = idx >> 2, k4 = idx & 3;
          bs[(4 * k4 + 0) * LDB_S + row] = rb[it].x;
          bs[(4 * k4 + 1) * LDB_S + row] = rb[it].y;
          bs[(4 * k4 + 2) * LDB_S + row] = rb[it].z;
          bs[(4 * k4 + 3) * LDB_S + row] = rb[it].w;
        } else {
          const int kk = idx / (BN / 4), n4 = idx % (BN / 4);
          *reinterpret_cast<float4*>(bs + kk * LDB_S + 4 * n4) = rb[it];
        }
      }
    }
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;

  int buf = 0;
  if (kbeg < kend) {
    load_tiles(kbeg);
    store_tiles(0);
  }
  __syncthreads();
  for (int64_t k0 = kbeg; k0 < kend; k0 += kBK) {
    const bool has_next = (k0 + kBK) < kend;
    if (has_next) load_tiles(k0 + kBK);
    const float* as = As + buf * A_STAGE;
    const float* bs = Bs + buf * B_STAGE;
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      float a[8], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(as + kk * LDA_S + 4 * ty);
      const float4 a1 = *reinterpret_cast<const float4*>(as + kk * LDA_S + 64 + 4 * ty);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      if constexpr (GW == 4) {
#pragma unroll
        for (int gi = 0; gi < NG; ++gi) {
          const float4 bv = *reinterpret_cast<const float4*>(bs + kk * LDB_S + 64 * gi + 4 * tx);
          b[4 * gi + 0] = bv.x; b[4 * gi + 1] = bv.y; b[4 * gi + 2] = bv.z; b[4 * gi + 3] = bv.w;
        }
      } else {
        const float2 bv = *reinterpret_cast<const float2*>(bs + kk * LDB_S + 2 * tx);
        b[0] = bv.x; b[1] = bv.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if constexpr (kBiasGrad) {
      if (g.bias_grad != nullptr && n_tile == 0 && tid < kBM) {
#pragma unroll
        for (int kk = 0; kk < kBK; ++kk) bsum += as[kk * LDA_S + tid];
      }
    }
    if (has_next) store_tiles(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  if constexpr (kBiasGrad) {
    if (g.bias_grad != nullptr && n_tile == 0 && tid < kBM && m0 + tid < g.M && kbeg < kend)
      atomicAdd(g.bias_grad + m0 + tid, bsum);
  }

  // ---------------- epilogue
  double s1[kStats ? TN : 1], s2[kStats ? TN : 1];
  if constexpr (kStats) {
#pragma unroll
    for (int j = 0; j < TN; ++j) { s1[j] = 0.0; s2[j] = 0.0; }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? 4 * ty + i : 64 + 4 * ty + (i - 4));
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) {
      const int n = n0 + (GW == 4 ? 64 * gi + 4 * tx : 2 * tx);
      float vals[GW];
#pragma unroll
      for (int j = 0; j < GW; ++j) vals[j] = acc[i][gi * GW + j];
      const bool valid = (m < g.M) && (n < g.N);
      epi.template apply<GW>(m, n, vals, valid);
      if constexpr (kStats) {
        if (valid) {
#pragma unroll
          for (int j = 0; j < GW; ++j) {
            s1[gi * GW + j] += (double)vals[j];
            s2[gi * GW + j] += (double)vals[j] * (double)vals[j];
          }
        }
      }
    }
  }
  if constexpr (kStats) {
    double* red1 = reinterpret_cast<double*>(smem);   // [16][BN]
    double* red2 = red1 + 16 * BN;
#pragma unroll
    for (int gi = 0; gi < NG; ++gi)
#pragma unroll
      for (int j = 0; j < GW; ++j) {
        const int nl = (GW == 4 ? 64 * gi + 4 * tx : 2 * tx) + j;
        red1[ty * BN + nl] = s1[gi * GW + j];
        red2[ty * BN + nl] = s2[gi * GW + j];
      }
    __syncthreads();
    if (tid < BN && n0 + tid < g.N) {
      double a = 0.0, b = 0.0;
#pragma unroll
      for (int r = 0; r < 16; ++r) { a += red1[r * BN + tid]; b += red2[r * BN + tid]; }
      atomicAdd(g.col_stats + n0 + tid, a);
      atomicAdd(g.col_stats + g.N + n0 + tid, b);
    }
  }
}

// ------------------------------------------------------------------------------------ epilogues
// Y = acc + bias (optional) ; optional ReLU ; store
// (tensor-core kernel: PreD / PreN hold the epilogue's global operands for one float4 of output, fetched a
//  whole tile ahead (deep: random gathers, streams) or one chunk ahead (near: cache-friendly operands);
//  apply_pre() finishes the element; kIdx = the functor needs the row's src/dst node ids)
struct EpiBias {
  float* C; int64_t ldc; const float* bias; int relu;
  static constexpr bool kIdx = false;
  static constexpr bool kWideStaging = true;      // tensor-core kernel: 64-column staging tile, 3 stages
  static constexpr bool kMidStaging = false;
  static constexpr bool kRowReduce = false;
  struct PreD {};
  struct PreN {};
  __device__ __forceinline__ void prefetch_deep(PreD&, int64_t, int, int, int) const {}
  __device__ __forceinline__ void prefetch_near(PreN&, int64_t, int, int, int) const {}
  struct ChunkC { float4 b; };     // the bias of this thread's four columns
  __device__ __forceinline__ ChunkC chunk_const(int n, bool ok) const {
    ChunkC c;
    c.b = (ok && bias) ? __ldg(reinterpret_cast<const float4*>(bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
    return c;
  }
  __device__ __forceinline__ void apply_pre(int64_t m, int n, float (&acc)[4], const PreD&, const PreN&, const ChunkC& c,
                                            bool valid) const {
    if (!valid) return;
    acc[0] += c.b.x; acc[1] += c.b.y; acc[2] += c.b.z; acc[3] += c.b.w;
    if (relu) { acc[0] = fmaxf(acc[0], 0.f); acc[1] = fmaxf(acc[1], 0.f); acc[2] = fmaxf(acc[2], 0.f); acc[3] = fmaxf(acc[3], 0.f); }
    *reinterpret_cast<float4*>(C + m * ldc + n) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
  template <int GW>
  __device__ __forceinline__ void apply(int64_t m, int n, float (&acc)[GW], bool valid) const {
    if (!valid) return;
#pragma unroll
    for (int j = 0; j < GW; ++j) {
      float v = acc[j] + (bias ? __ldg(bias + n + j) : 0.f);
      acc[j] = relu ? fmaxf(v, 0.f) : v;
    }
    if constexpr (GW == 4) *reinterpret_cast<float4*>(C + m * ldc + n) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else *reinterpret_cast<float2*>(C + m * ldc + n) = make_float2(acc[0], acc[1]);
  }
};

// dX = acc (+ addend) ; zero where mask <= 0 ; store          (bwd-data, ReLU backward fused)
// kAdd / kMask select at compile time which operands exist (so the tensor-core epilogue only holds prefetch
// registers for what it uses)
template <bool kAdd, bool kMask>
struct EpiAddMaskT {
  float* C; int64_t ldc; const float* addend; const float* mask;
  template <int GW>
  __device__ __forceinline__ void apply(int64_t m, int n, float (&acc)[GW], bool valid) const {
    if (!valid) return;
#pragma unroll
    for (int j = 0; j < GW; ++j) {
      float v = acc[j];
      if constexpr (kAdd) v += __ldg(addend + m * ldc + n + j);
      if constexpr (kMask) { if (!(__ldg(mask + m * ldc + n + j) > 0.f)) v = 0.f; }
      acc[j] = v;
    }
    if constexpr (GW == 4) *reinterpret_cast<float4*>(C + m * ldc + n) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else *reinterpret_cast<float2*>(C + m * ldc + n) = make_float2(acc[0], acc[1]);
  }
  static constexpr bool kIdx = false;
  static constexpr bool kWideStaging = false;
#ifdef GG_MID_ADD       // A/B builds: 32-column chunks (whole lines per row piece) + 3 stages for the bwd-data GEMMs
  static constexpr bool kMidStaging = true;
#else
  static constexpr bool kMidStaging = false;
#endif
  static constexpr bool kRowReduce = false;
  struct Empty {};
  struct Val { float4 v; };
  using PreD = typename std::conditional<kAdd, Val, Empty>::type;
  using PreN = typename std::conditional<kMask, Val, Empty>::type;
  __device__ __forceinline__ void prefetch_deep(PreD& p, int64_t m, int n, int, int) const {
    if constexpr (kAdd) p.v = __ldg(reinterpret_cast<const float4*>(addend + m * ldc + n));
  }
  __device__ __forceinline__ void prefetch_near(PreN& p, int64_t m, int n, int, int) const {
    if constexpr (kMask) p.v = __ldg(reinterpret_cast<const float4*>(mask + m * ldc + n));
  }
  struct ChunkC {};
  __device__ __forceinline__ ChunkC chunk_const(int, bool) const { return ChunkC{}; }
  __device__ __forceinline__ void apply_pre(int64_t m, int n, float (&acc)[4], const PreD& d, const PreN& p, const ChunkC&,
                                            bool valid) const {
    if (!valid) return;
    if constexpr (kAdd) { acc[0] += d.v.x; acc[1] += d.v.y; acc[2] += d.v.z; acc[3] += d.v.w; }
    if constexpr (kMask) {
      acc[0] = p.v.x > 0.f ? acc[0] : 0.f; acc[1] = p.v.y > 0.f ? acc[1] : 0.f;
      acc[2] = p.v.z > 0.f ? acc[2] : 0.f; acc[3] = p.v.w > 0.f ? acc[3] : 0.f;
    }
    *reinterpret_cast<float4*>(C + m * ldc + n) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
};

// split-K weight gradient: C += acc with fp32 atomics (C zeroed by the caller)
struct EpiAtomic {
  float* C; int64_t ldc;
  static constexpr bool kIdx = false;
  static constexpr bool kWideStaging = false;
  static constexpr bool kMidStaging = false;
  static constexpr bool kRowReduce = false;
  struct PreD {};
  struct PreN {};
  __device__ __forceinline__ void prefetch_deep(PreD&, int64_t, int, int, int) const {}
  __device__ __forceinline__ void prefetch_near(PreN&, int64_t, int, int, int) const {}
  struct ChunkC {};
  __device__ __forceinline__ ChunkC chunk_const(int, bool) const { return ChunkC{}; }
  __device__ __forceinline__ void apply_pre(int64_t m, int n, float (&acc)[4], const PreD&, const PreN&, const ChunkC&,
                                            bool valid) const {
    apply<4>(m, n, acc, valid);
  }
  template <int GW>
  __device__ __forceinline__ void apply(int64_t m, int n, float (&acc)[GW], bool valid) const {
    if (!valid) return;
#pragma unroll
    for (int j = 0; j < GW; ++j) atomicAdd(C + m * ldc + n + j, acc[j]);
  }
};

// edge gate pre-activation (layers/gated_gcn_full.py:120-121):
//   t[i, c] = (B3 e_i + b3)[c] + B1h[src_i, c] + B2h[dst_i, c];  P rows are [A1h|A2h|A3h|B1h|B2h]
struct EpiEdgeGate {
  float* t; int d; const float* b3; const float* P; const int32_t* src; const int32_t* dst;
  static constexpr bool kIdx = true;
  static constexpr bool kWideStaging = false;     // (the 64-column tile does not fit next to the resident W at d = 128 and
                                                  // measured 3-4 % slower at d = 256, same-box A/B, tools/ab_sweep.sh)
  static constexpr bool kMidStaging = true;       // 32-column chunks: every gathered / stored row piece is a whole line
  static constexpr bool kRowReduce = false;
  struct PreD { float4 p1; };     // B1h[src]: random row gather -> a whole tile ahead
  struct PreN { float4 p2; };     // B2h[dst]: edges are dst-sorted, consecutive rows share it -> one chunk ahead
  __device__ __forceinline__ void prefetch_deep(PreD& p, int64_t, int n, int s, int) const {
    p.p1 = __ldg(reinterpret_cast<const float4*>(P + (int64_t)s * (5 * d) + 3 * d + n));
  }
  __device__ __forceinline__ void prefetch_near(PreN& p, int64_t, int n, int, int v) const {
    p.p2 = __ldg(reinterpret_cast<const float4*>(P + (int64_t)v * (5 * d) + 4 * d + n));
  }
  struct ChunkC { float4 b; };     // b3 of this thread's four columns
  __device__ __forceinline__ ChunkC chunk_const(int n, bool ok) const {
    ChunkC c;
    c.b = ok ? __ldg(reinterpret_cast<const float4*>(b3 + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
    return c;
  }
  __device__ __forceinline__ void apply_pre(int64_t m, int n, float (&acc)[4], const PreD& dp, const PreN& np, const ChunkC& c,
                                            bool valid) const {
    if (!valid) return;
    const float4 b = c.b;
    acc[0] = (dp.p1.x + np.p2.x) + (acc[0] + b.x);
    acc[1] = (dp.p1.y + np.p2.y) + (acc[1] + b.y);
    acc[2] = (dp.p1.z + np.p2.z) + (acc[2] + b.z);
    acc[3] = (dp.p1.w + np.p2.w) + (acc[3] + b.w);
    *reinterpret_cast<float4*>(t + m * d + n) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
  template <int GW>
  __device__ __forceinline__ void apply(int64_t m, int n, float (&acc)[GW], bool valid) const {
    static_assert(GW == 4, "edge gate epilogue runs on 128/64-wide tiles");
    if (!valid) return;
    const int64_t s = __ldg(src + m), v = __ldg(dst + m);
    const float4 b = __ldg(reinterpret_cast<const float4*>(b3 + n));
    const float4 p1 = __ldg(reinterpret_cast<const float4*>(P + s * (5 * d) + 3 * d + n));
    const float4 p2 = __ldg(reinterpret_cast<const float4*>(P + v * (5 * d) + 4 * d + n));
    // same association as the reference: (B1h[s] + B2h[v]) + B3e, with B3e = acc + b3
    acc[0] = (p1.x + p2.x) + (acc[0] + b.x);
    acc[1] = (p1.y + p2.y) + (acc[1] + b.y);
    acc[2] = (p1.z + p2.z) + (acc[2] + b.z);
    acc[3] = (p1.w + p2.w) + (acc[3] + b.w);
    *reinterpret_cast<float4*>(t + m * d + n) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
};

// score predictor (layers/score_predictor.py:13-17) on a 64-wide tile:
//   hid = relu(acc + Q[src, n] + Q[dst, H + n]) ; score = sum_n hid*w2[n] + b2 ; optional save of hid
struct EpiScore {
  float* score; float* hid; const float* Q; const float* w2; const float* b2; const int32_t* src; const int32_t* dst;
  int64_t M;
  template <int GW>
  __device__ __forceinline__ void apply(int64_t m, int n, float (&acc)[GW], bool valid) const {
    static_assert(GW == 4, "score epilogue runs on the 64-wide tile");
    constexpr int H = 64;
    const int64_t mc = m < M ? m : M - 1;     // clamp so that all 16 lanes of the row stay converged
    const int64_t s = __ldg(src + mc), v = __ldg(dst + mc);
    const float4 q1 = __ldg(reinterpret_cast<const float4*>(Q + s * (2 * H) + n));
    const float4 q2 = __ldg(reinterpret_cast<const float4*>(Q + v * (2 * H) + H + n));
    const float4 w = __ldg(reinterpret_cast<const float4*>(w2 + n));
    acc[0] = fmaxf(acc[0] + q1.x + q2.x, 0.f);
    acc[1] = fmaxf(acc[1] + q1.y + q2.y, 0.f);
    acc[2] = fmaxf(acc[2] + q1.z + q2.z, 0.f);
    acc[3] = fmaxf(acc[3] + q1.w + q2.w, 0.f);
    float part = acc[0] * w.x + acc[1] * w.y + acc[2] * w.z + acc[3] * w.w;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);   // 16 lanes share a row
    if (valid) {
      if (hid) *reinterpret_cast<float4*>(hid + m * H + n) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      if ((threadIdx.x & 15) == 0) score[m] = part + __ldg(b2);
    }
  }
};

// the same on the tensor-core kernel (wide staging: the 16 threads that share a row sit in one half-warp).
// N = 64 < BN: the TMA box of the B operand reaches past W1e's 64 rows and is zero-filled, the upper 64
// accumulator columns are never valid.
struct EpiScoreTC {
  float* score; float* hid; const float* Q; const float* w2; const float* b2; const int32_t* src; const int32_t* dst;
  static constexpr bool kIdx = true;
  static constexpr bool kWideStaging = true;
  static constexpr bool kMidStaging = false;
  static constexpr bool kRowReduce = true;
  static constexpr int H = 64;
  struct PreD { float4 q1; };     // Q[src, n]
  struct PreN { float4 q2; };     // Q[dst, H + n]
  __device__ __forceinline__ void prefetch_deep(PreD& p, int64_t, int n, int s, int) const {
    if (n < H) p.q1 = __ldg(reinterpret_cast<const float4*>(Q + (int64_t)s * (2 * H) + n));
  }
  __device__ __forceinline__ void prefetch_near(PreN& p, int64_t, int n, int, int v) const {
    if (n < H) p.q2 = __ldg(reinterpret_cast<const float4*>(Q + (int64_t)v * (2 * H) + H + n));
  }
  struct ChunkC {};
  __device__ __forceinline__ ChunkC chunk_const(int, bool) const { return ChunkC{}; }
  __device__ __forceinline__ void apply_pre(int64_t m, int n, float (&acc)[4], const PreD& d, const PreN& p, const ChunkC&,
                                            bool valid) const {
    if (!valid) { acc[0] = acc[1] = acc[2] = acc[3] = 0.f; return; }
    acc[0] = fmaxf(acc[0] + d.q1.x + p.q2.x, 0.f);
    acc[1] = fmaxf(acc[1] + d.q1.y + p.q2.y, 0.f);
    acc[2] = fmaxf(acc[2] + d.q1.z + p.q2.z, 0.f);
    acc[3] = fmaxf(acc[3] + d.q1.w + p.q2.w, 0.f);
    if (hid) *reinterpret_cast<float4*>(hid + m * H + n) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
  // this thread's share of the row's dot product with w2 (0 for invalid elements: acc was zeroed)
  __device__ __forceinline__ float row_partial(const float (&acc)[4], int n) const {
    if (n >= H) return 0.f;
    const float4 w = __ldg(reinterpret_cast<const float4*>(w2 + n));
    return acc[0] * w.x + acc[1] * w.y + acc[2] * w.z + acc[3] * w.w;
  }
  __device__ __forceinline__ void row_finish(int64_t m, float sum) const { score[m] = sum + __ldg(b2); }
};

// ------------------------------------------------------------------------------------ launcher
template <int BN, bool A_T, bool B_T, bool kStats, bool kBiasGrad, class Epi>
int launch_gemm(const char* tag, GemmArgs g, const Epi& epi, int splits, cudaStream_t stream) {
  if (g.M <= 0 || g.N <= 0) return GG_OK;
  g.n_tiles = (g.N + BN - 1) / BN;
  const int64_t m_tiles = (g.M + kBM - 1) / kBM;
  if (splits < 1) splits = 1;
  int64_t chunk = (g.K + splits - 1) / splits;
  chunk = ((chunk + kBK - 1) / kBK) * kBK;
  if (chunk < kBK) chunk = kBK;
  splits = (int)((g.K + chunk - 1) / chunk);
  if (splits < 1) splits = 1;
  g.splits = splits;
  g.k_chunk = chunk;
  const int64_t blocks = m_tiles * g.n_tiles;
  if (blocks > 2147483647LL) { set_error("gemm: grid too large"); return GG_ERR_ARG; }
  dim3 grid((unsigned)blocks, (unsigned)splits, 1);
  GG_KERNEL_BEGIN(tag, stream);
  gemm_ffma_kernel<BN, A_T, B_T, kStats, kBiasGrad, Epi><<<grid, kGemmThreads, 0, stream>>>(g, epi);
  GG_KERNEL_END(tag, stream);
  return GG_OK;
}

}  // namespace gg
