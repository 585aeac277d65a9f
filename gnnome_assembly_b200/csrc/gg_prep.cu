// gg_prep.cu — "next" rows of SURVEY.md §8f on the GPU, on the same GraphPlan:
//   (1) input preparation: z-scored edge features (utils.py:70-74), in/out degrees and the 16-step PageRank
//       positional encoding (utils.py:102-138, type_pe == 'PR'), i.e. pe_dim sparse mat-vecs on the plan's CSR;
//   (2) BCEWithLogitsLoss(pos_weight) (train.py:211,255) fused with the TP/TN/FP/FN counts of
//       utils.calculate_tfpn (utils.py:217-223): one pass and one D2H instead of five .item() syncs.
// Everything the reference computes in float64 (scipy) is computed in fp64 here and cast to fp32 at the end.
#include "gg_common.cuh"

namespace gg {

__device__ __forceinline__ double warp_sum_d(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// block-wide sum of up to K doubles per thread -> atomicAdd into out[K]
template <int K>
__device__ __forceinline__ void block_atomic_sums(double (&v)[K], double* out) {
  __shared__ double sh[K][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double s = warp_sum_d(v[k]);
    if (lane == 0) sh[k][warp] = s;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double s = lane < nwarp ? sh[k][lane] : 0.0;
      s = warp_sum_d(s);
      if (lane == 0) atomicAdd(out + k, s);
    }
  }
}

// ---- z-score of two edge feature columns: two passes (mean, then centred sum of squares), unbiased std
__global__ void zscore_sum_kernel(int64_t E, const float* __restrict__ a, const float* __restrict__ b, double* ws) {
  double v[2] = {0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) {
    v[0] += (double)a[i];
    v[1] += (double)b[i];
  }
  block_atomic_sums<2>(v, ws);
}
__global__ void zscore_sq_kernel(int64_t E, const float* __restrict__ a, const float* __restrict__ b, double* ws) {
  const double ma = ws[0] / (double)E, mb = ws[1] / (double)E;
  double v[2] = {0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) {
    const double da = (double)a[i] - ma, db = (double)b[i] - mb;
    v[0] += da * da;
    v[1] += db * db;
  }
  block_atomic_sums<2>(v, ws + 2);
}
__global__ void zscore_apply_kernel(int64_t E, const float* __restrict__ a, const float* __restrict__ b,
                                    const double* __restrict__ ws, float* __restrict__ out) {
  // torch: (x - x.mean()) / x.std(), std unbiased (utils.py:72-73).  The reference does this in fp32, where
  // x - mean cancels badly for overlap_similarity (values in [0.99, 1]); here the subtraction and division
  // are done in fp64 and rounded once, so the result is the correctly rounded z-score.
  const double ma = ws[0] / (double)E, mb = ws[1] / (double)E;
  const double ia = 1.0 / sqrt(ws[2] / (double)(E - 1)), ib = 1.0 / sqrt(ws[3] / (double)(E - 1));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) {
    out[2 * i + 0] = (float)(((double)a[i] - ma) * ia);
    out[2 * i + 1] = (float)(((double)b[i] - mb) * ib);
  }
}

// ---- PageRank positional encoding on the plan (internal node order), output rows in CALLER node order
__global__ void pe_init_kernel(int64_t N, const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ out_ptr,
                               const int32_t* __restrict__ node_perm, int width, float* __restrict__ pe,
                               double* __restrict__ x, double* __restrict__ dinv) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  const int indeg = in_ptr[v + 1] - in_ptr[v], outdeg = out_ptr[v + 1] - out_ptr[v];
  const int64_t row = node_perm[v];
  pe[row * width + 0] = (float)indeg;                       // utils.py:102
  pe[row * width + 1] = (float)outdeg;                      // utils.py:103
  x[v] = 1.0 / (double)N;                                   // utils.py:131
  dinv[v] = outdeg > 0 ? 1.0 / ((double)outdeg + 1e-9) : 0.0;   // utils.py:126
}
__global__ void pe_step_kernel(int64_t N, const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ src,
                               const int32_t* __restrict__ node_perm, double alpha, int width, int col,
                               const double* __restrict__ x, const double* __restrict__ dinv,
                               double* __restrict__ x_new, float* __restrict__ pe) {
  // x_new = alpha * (Dinv A)^T x + (1 - alpha) / n    (utils.py:128,135): row v of (Dinv A)^T = in-edges of v
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  double s = 0.0;
  for (int i = in_ptr[v]; i < in_ptr[v + 1]; ++i) {
    const int u = src[i];
    s += x[u] * dinv[u];
  }
  const double r = alpha * s + (1.0 - alpha) / (double)N;
  x_new[v] = r;
  pe[(int64_t)node_perm[v] * width + col] = (float)r;
}

// ---- fused BCE-with-logits (pos_weight) + confusion counts
__global__ void bce_metrics_kernel(int64_t E, const float* __restrict__ s, const float* __restrict__ y, float pw,
                                   double* __restrict__ out) {
  double v[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = s[i], t = y[i];
    // ATen: (1 - y) * x + (1 + (pw - 1) * y) * (log1p(exp(-|x|)) + max(-x, 0))
    const float lw = 1.0f + (pw - 1.0f) * t;
    v[0] += (double)((1.0f - t) * x + lw * (log1pf(expf(-fabsf(x))) + fmaxf(-x, 0.0f)));
    // utils.py:218: round(sigmoid(x)) -> 1 iff sigmoid(x) > 0.5 (round half to even sends exactly 0.5 to 0)
    const float p = 1.0f / (1.0f + expf(-x));
    const bool pos = rintf(p) == 1.0f;
    const bool lab1 = t == 1.0f, lab0 = t == 0.0f;
    v[1] += (pos && lab1) ? 1.0 : 0.0;       // TP
    v[2] += (!pos && lab0) ? 1.0 : 0.0;      // TN
    v[3] += (pos && lab0) ? 1.0 : 0.0;       // FP
    v[4] += (!pos && lab1) ? 1.0 : 0.0;      // FN
  }
  block_atomic_sums<5>(v, out);
}
__global__ void bce_bwd_kernel(int64_t E, const float* __restrict__ s, const float* __restrict__ y, float pw,
                               const float* __restrict__ g_loss, float* __restrict__ g) {
  const float scale = g_loss[0] / (float)E;     // mean reduction
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = s[i], t = y[i];
    const float p = 1.0f / (1.0f + expf(-x));
    g[i] = scale * (p * (1.0f - t + pw * t) - pw * t);
  }
}

static unsigned flat_grid(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace gg

using namespace gg;

extern "C" {

int gg_prep_edge_features(int64_t E, const float* overlap_length, const float* overlap_similarity, float* e_out,
                          double* ws, void* stream) {
  GG_REQUIRE(E >= 0, "prep_edge_features: bad size");
  if (E == 0) return GG_OK;
  GG_REQUIRE(overlap_length && overlap_similarity && e_out && ws, "prep_edge_features: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  GG_CUDA(cudaMemsetAsync(ws, 0, 4 * sizeof(double), st));
  const unsigned grid = flat_grid(E, 256);
  GG_KERNEL_BEGIN("zscore_sum_kernel", st);
  zscore_sum_kernel<<<grid, 256, 0, st>>>(E, overlap_length, overlap_similarity, ws);
  GG_KERNEL_END("zscore_sum_kernel", st);
  GG_KERNEL_BEGIN("zscore_sq_kernel", st);
  zscore_sq_kernel<<<grid, 256, 0, st>>>(E, overlap_length, overlap_similarity, ws);
  GG_KERNEL_END("zscore_sq_kernel", st);
  GG_KERNEL_BEGIN("zscore_apply_kernel", st);
  zscore_apply_kernel<<<grid, 256, 0, st>>>(E, overlap_length, overlap_similarity, ws, e_out);
  GG_KERNEL_END("zscore_apply_kernel", st);
  return GG_OK;
}

int gg_prep_pe(const gg_plan_t* plan, int pe_dim, double alpha, float* pe_out, double* ws, void* stream) {
  GG_REQUIRE(plan, "prep_pe: null plan");
  GG_REQUIRE(pe_dim >= 0, "prep_pe: bad pe_dim");
  const Plan* pl = reinterpret_cast<const Plan*>(plan);
  const int64_t N = pl->N;
  if (N == 0) return GG_OK;
  GG_REQUIRE(pe_out && ws, "prep_pe: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int width = pe_dim + 2;
  double* x = ws;
  double* x2 = ws + N;
  double* dinv = ws + 2 * N;
  const unsigned grid = (unsigned)((N + 255) / 256);
  GG_KERNEL_BEGIN("pe_init_kernel", st);
  pe_init_kernel<<<grid, 256, 0, st>>>(N, pl->in_ptr, pl->out_ptr, pl->node_perm, width, pe_out, x, dinv);
  GG_KERNEL_END("pe_init_kernel", st);
  for (int k = 0; k < pe_dim; ++k) {
    GG_KERNEL_BEGIN("pe_step_kernel", st);
    pe_step_kernel<<<grid, 256, 0, st>>>(N, pl->in_ptr, pl->src, pl->node_perm, alpha, width, 2 + k, x, dinv, x2, pe_out);
    GG_KERNEL_END("pe_step_kernel", st);
    double* tmp = x; x = x2; x2 = tmp;
  }
  return GG_OK;
}

int gg_bce_metrics_fwd(int64_t E, const float* scores, const float* y, float pos_weight, double* out5, void* stream) {
  GG_REQUIRE(E >= 0 && out5, "bce_metrics_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  GG_CUDA(cudaMemsetAsync(out5, 0, 5 * sizeof(double), st));
  if (E == 0) return GG_OK;
  GG_REQUIRE(scores && y, "bce_metrics_fwd: null pointer");
  GG_KERNEL_BEGIN("bce_metrics_kernel", st);
  bce_metrics_kernel<<<flat_grid(E, 256), 256, 0, st>>>(E, scores, y, pos_weight, out5);
  GG_KERNEL_END("bce_metrics_kernel", st);
  return GG_OK;
}

int gg_bce_bwd(int64_t E, const float* scores, const float* y, float pos_weight, const float* g_loss, float* g_scores,
               void* stream) {
  GG_REQUIRE(E >= 0, "bce_bwd: bad size");
  if (E == 0) return GG_OK;
  GG_REQUIRE(scores && y && g_loss && g_scores, "bce_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  GG_KERNEL_BEGIN("bce_bwd_kernel", st);
  bce_bwd_kernel<<<flat_grid(E, 256), 256, 0, st>>>(E, scores, y, pos_weight, g_loss, g_scores);
  GG_KERNEL_END("bce_bwd_kernel", st);
  return GG_OK;
}

}  // extern "C"
