// gg_edge_mlp.cu — the edge encoder  e0 = W2 relu(W1 e + b1) + b2  (models/full_graph.py:17-18,24-26): its backward
// in ONE pass over the E x d gradient, and (below) its forward in one pass as well.
//
// As three GEMM calls (dW2 = g^T hid, g_hid = (g W2) [hid > 0], dW1 = g_hid^T e) the E x d gradient is read
// twice by narrow-N FFMA kernels (N = 16) that run far from both rooflines (150 + 200 us on the chr19 graph,
// HBM time of one read: 30 us).  Here one warp owns a row at a time: lane l holds the row's channels
// Row<D>::channel(., l), the lane's rows of W2 live in registers for the whole kernel, dW2 / db2 accumulate in
// registers, the 16 hidden-unit sums are reduced across the warp with a transpose-reduce (16 shuffles instead
// of 80) and land one per lane pair, so g_hid is never written: dW1 / db1 are accumulated in place.
#include "gg_common.cuh"

namespace gg {

constexpr int kMlpThreads = 256;         // 8 warps, TWO rows per warp and iteration (see below)
constexpr int kMlpHid = 16, kMlpK = 4;       // hidden_edge_features = 16 (hyperparameters.py:11), edge_features 2 padded to 4

// Round 2: the kernel was bound by the shared-memory reads of W2 (ncu: l1tex 87 % busy, DRAM 13 %): every row re-read the
// lane's 16 x VPL weights.  Each warp now takes TWO rows per iteration and reads every W2 value once for both.
template <int D>
__global__ void __launch_bounds__(kMlpThreads, 1)
edge_mlp_bwd_kernel(int64_t E, const float* __restrict__ g, const float* __restrict__ hid, const float* __restrict__ e,
                    const float* __restrict__ W2, float* __restrict__ dW1, float* __restrict__ db1,
                    float* __restrict__ dW2, float* __restrict__ db2) {
  constexpr int VPL = D / 32, H = kMlpHid;
  constexpr int kOut = D * H + D + H * kMlpK + H;            // dW2 | db2 | dW1 | db1
  __shared__ float red[kOut];
  __shared__ __align__(16) float wsh[H * D];                 // W2 transposed: wsh[k][c], lane reads its channels as float4/float2
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * kMlpThreads + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * kMlpThreads) >> 5;
  for (int i = threadIdx.x; i < kOut; i += kMlpThreads) red[i] = 0.f;
  for (int i = threadIdx.x; i < D * H; i += kMlpThreads) wsh[(i % H) * D + i / H] = __ldg(W2 + i);   // W2[c][k] -> wsh[k][c]
  __syncthreads();

  float a2[VPL][H], ab2[VPL], a1[kMlpK], ab1 = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    ab2[i] = 0.f;
#pragma unroll
    for (int k = 0; k < H; ++k) a2[i][k] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < kMlpK; ++j) a1[j] = 0.f;
  const int my_k = lane >> 1;                                // the hidden unit this lane pair ends up owning

  // rows r0 = it, r1 = it + nw of this warp's stride-2nw sequence; a missing second row contributes zeros.
  // Software pipeline: the next pair's operands are requested before the current pair is processed (8 warps per SM
  // cannot hide a DRAM round trip per pair on their own).
  Row<D> g0n, g1n;
  float4 h0n[H / 4], h1n[H / 4], e0n, e1n;
  float hk0n, hk1n;
  auto fetch = [&](int64_t r0) {
    const int64_t r1 = r0 + nw;
    const bool two = r1 < E;
    g0n.load_stream(g + r0 * D, lane);
    if (two) g1n.load_stream(g + r1 * D, lane); else g1n.fill(0.f);
#pragma unroll
    for (int k = 0; k < H / 4; ++k) {                        // same address in every lane (broadcast)
      h0n[k] = __ldg(reinterpret_cast<const float4*>(hid + r0 * H) + k);
      h1n[k] = two ? __ldg(reinterpret_cast<const float4*>(hid + r1 * H) + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    e0n = __ldg(reinterpret_cast<const float4*>(e + r0 * kMlpK));
    e1n = two ? __ldg(reinterpret_cast<const float4*>(e + r1 * kMlpK)) : make_float4(0.f, 0.f, 0.f, 0.f);
    hk0n = __ldg(hid + r0 * H + my_k);
    hk1n = two ? __ldg(hid + r1 * H + my_k) : 0.f;
  };
  if (gw < E) fetch(gw);
  for (int64_t r0 = gw; r0 < E; r0 += 2 * nw) {
    const Row<D> g0 = g0n, g1 = g1n;
    float h0[H], h1[H];
#pragma unroll
    for (int k = 0; k < H / 4; ++k) {
      h0[4 * k] = h0n[k].x; h0[4 * k + 1] = h0n[k].y; h0[4 * k + 2] = h0n[k].z; h0[4 * k + 3] = h0n[k].w;
      h1[4 * k] = h1n[k].x; h1[4 * k + 1] = h1n[k].y; h1[4 * k + 2] = h1n[k].z; h1[4 * k + 3] = h1n[k].w;
    }
    const float4 e0 = e0n, e1 = e1n;
    const float hk0 = hk0n, hk1 = hk1n;
    if (r0 + 2 * nw < E) fetch(r0 + 2 * nw);
    float v0[H], v1[H];
#pragma unroll
    for (int k = 0; k < H; ++k) { v0[k] = 0.f; v1[k] = 0.f; }
#pragma unroll
    for (int i = 0; i < VPL; ++i) ab2[i] += g0.v[i] + g1.v[i];
#pragma unroll
    for (int k = 0; k < H; ++k) {
      Row<D> wk;                                             // W2[c][k] for this lane's channels (conflict-free LDS), read ONCE for both rows
      if constexpr (D == 64) {
        const float2 x = reinterpret_cast<const float2*>(wsh + k * D)[lane];
        wk.v[0] = x.x; wk.v[1] = x.y;
      } else {
#pragma unroll
        for (int j = 0; j < D / 128; ++j) {
          const float4 x = reinterpret_cast<const float4*>(wsh + k * D + 128 * j)[lane];
          wk.v[4 * j] = x.x; wk.v[4 * j + 1] = x.y; wk.v[4 * j + 2] = x.z; wk.v[4 * j + 3] = x.w;
        }
      }
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        v0[k] = fmaf(g0.v[i], wk.v[i], v0[k]);               // partial (g W2)[k] over this lane's channels
        v1[k] = fmaf(g1.v[i], wk.v[i], v1[k]);
        a2[i][k] = fmaf(g0.v[i], h0[k], a2[i][k]);           // dW2[c][k] += g[c] hid[k]
        a2[i][k] = fmaf(g1.v[i], h1[k], a2[i][k]);
      }
    }
    // transpose-reduce: after the step with offset o, a lane keeps the half of its values selected by (lane & o)
#pragma unroll
    for (int o = 16, n = H / 2; o >= 2; o >>= 1, n >>= 1) {
      const bool up = (lane & o) != 0;
#pragma unroll
      for (int j = 0; j < n; ++j) {
        const float s0 = up ? v0[j] : v0[j + n], k0 = up ? v0[j + n] : v0[j];
        const float s1 = up ? v1[j] : v1[j + n], k1 = up ? v1[j + n] : v1[j];
        v0[j] = k0 + __shfl_xor_sync(0xffffffffu, s0, o);
        v1[j] = k1 + __shfl_xor_sync(0xffffffffu, s1, o);
      }
    }
    v0[0] += __shfl_xor_sync(0xffffffffu, v0[0], 1);         // both lanes of the pair: (g W2)[lane >> 1]
    v1[0] += __shfl_xor_sync(0xffffffffu, v1[0], 1);
    const float gh0 = hk0 > 0.f ? v0[0] : 0.f;               // ReLU backward
    const float gh1 = hk1 > 0.f ? v1[0] : 0.f;
    a1[0] = fmaf(gh0, e0.x, a1[0]); a1[1] = fmaf(gh0, e0.y, a1[1]);
    a1[2] = fmaf(gh0, e0.z, a1[2]); a1[3] = fmaf(gh0, e0.w, a1[3]);
    a1[0] = fmaf(gh1, e1.x, a1[0]); a1[1] = fmaf(gh1, e1.y, a1[1]);
    a1[2] = fmaf(gh1, e1.z, a1[2]); a1[3] = fmaf(gh1, e1.w, a1[3]);
    ab1 += gh0 + gh1;
  }

  __syncthreads();
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = Row<D>::channel(i, lane);
#pragma unroll
    for (int k = 0; k < H; ++k) atomicAdd(&red[c * H + k], a2[i][k]);
    atomicAdd(&red[D * H + c], ab2[i]);
  }
  if ((lane & 1) == 0) {
#pragma unroll
    for (int j = 0; j < kMlpK; ++j) atomicAdd(&red[D * H + D + my_k * kMlpK + j], a1[j]);
    atomicAdd(&red[D * H + D + H * kMlpK + my_k], ab1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kOut; i += kMlpThreads) {
    const float x = red[i];
    float* dst = i < D * H ? dW2 + i : i < D * H + D ? db2 + (i - D * H)
               : i < D * H + D + H * kMlpK ? dW1 + (i - D * H - D) : db1 + (i - D * H - D - H * kMlpK);
    atomicAdd(dst, x);
  }
}


// Forward of the same encoder in one pass: hid = relu(W1 e + b1) [E, 16], out = W2 hid + b2 [E, d].  As two GEMM calls
// (N = 16 and K = 16: FFMA kernels, 29 + 83 us on the chr19 graph) the E x 16 intermediate is written and re-read and
// the second call runs at a third of the write bandwidth.  Here a warp owns 32 CONSECUTIVE rows at a time: lane i loads
// row i's (padded) features with one coalesced access and computes that row's 16 hidden units (W1 / b1 are shared-memory
// broadcasts), stores them as one contiguous 2 KB block, and then the warp walks the 32 rows: the row's hidden units
// are broadcast from lane i (16 shuffles) into the W2 rows every lane keeps in registers for its Row<D> channels.  The
// next block's features are requested before the current one is processed.  (A first version that gave every warp
// single rows in a grid-stride loop paid a DRAM round trip per row pair: 136 us, slower than the two GEMMs.)
template <int D>
__global__ void __launch_bounds__(kMlpThreads, 2)
edge_mlp_fwd_kernel(int64_t E, const float* __restrict__ e, const float* __restrict__ W1, const float* __restrict__ b1,
                    const float* __restrict__ W2, const float* __restrict__ b2, float* __restrict__ hid,
                    float* __restrict__ out) {
  constexpr int VPL = D / 32, H = kMlpHid;
  __shared__ float4 sw1[H];
  __shared__ float sb1[H];
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * kMlpThreads + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * kMlpThreads) >> 5;
  if (threadIdx.x < H) {
    sw1[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(W1) + threadIdx.x);   // W1[k][0..3] (K padded to 4 with zeros)
    sb1[threadIdx.x] = __ldg(b1 + threadIdx.x);
  }
  __syncthreads();
  float w2[VPL][H], bo[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = Row<D>::channel(i, lane);
    bo[i] = __ldg(b2 + c);
#pragma unroll
    for (int k = 0; k < H / 4; ++k) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(W2 + (int64_t)c * H) + k);
      w2[i][4 * k] = t.x; w2[i][4 * k + 1] = t.y; w2[i][4 * k + 2] = t.z; w2[i][4 * k + 3] = t.w;
    }
  }
  const int64_t blocks32 = (E + 31) / 32;
  auto fetch = [&](int64_t blk) {
    const int64_t r = blk * 32 + lane;
    return (blk < blocks32 && r < E) ? __ldg(reinterpret_cast<const float4*>(e) + r) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  float4 xn = fetch(gw);
  for (int64_t blk = gw; blk < blocks32; blk += nw) {
    const float4 x = xn;
    xn = fetch(blk + nw);
    const int64_t base = blk * 32;
    const int rows = (int)(E - base < 32 ? E - base : 32);
    float h[H];
#pragma unroll
    for (int k = 0; k < H; ++k) {                            // same order as the GEMM path: sum over k, then the bias
      const float4 w = sw1[k];
      h[k] = fmaxf(fmaf(x.w, w.w, fmaf(x.z, w.z, fmaf(x.y, w.y, x.x * w.x))) + sb1[k], 0.f);
    }
    if (lane < rows) {
      float4* hp = reinterpret_cast<float4*>(hid + (base + lane) * H);
#pragma unroll
      for (int k = 0; k < H / 4; ++k) hp[k] = make_float4(h[4 * k], h[4 * k + 1], h[4 * k + 2], h[4 * k + 3]);
    }
    for (int j = 0; j < rows; j += 2) {                      // two rows per step: independent shuffle / FMA chains
      const int j1 = j + 1 < rows ? j + 1 : j;
      Row<D> o0, o1;
#pragma unroll
      for (int i = 0; i < VPL; ++i) { o0.v[i] = 0.f; o1.v[i] = 0.f; }
#pragma unroll
      for (int k = 0; k < H; ++k) {
        const float a0 = __shfl_sync(0xffffffffu, h[k], j), a1 = __shfl_sync(0xffffffffu, h[k], j1);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          o0.v[i] = fmaf(a0, w2[i][k], o0.v[i]);
          o1.v[i] = fmaf(a1, w2[i][k], o1.v[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < VPL; ++i) { o0.v[i] += bo[i]; o1.v[i] += bo[i]; }
      o0.store(out + (base + j) * D, lane);
      if (j + 1 < rows) o1.store(out + (base + j + 1) * D, lane);
    }
  }
}

}  // namespace gg

using namespace gg;

extern "C" int gg_edge_mlp_fwd(int64_t E, int d, int hidden, int K, const float* e, const float* W1, const float* b1,
                               const float* W2, const float* b2, float* hid, float* out, void* stream) {
  GG_REQUIRE(E >= 0, "edge_mlp_fwd: negative size");
  if (hidden != kMlpHid || K != kMlpK || !(d == 64 || d == 128)) {
    set_error("gnnome_b200: edge_mlp_fwd is built for hidden_edge_features = 16, K = 4 (2 padded), d in {64,128}");
    return GG_ERR_UNSUPPORTED;
  }
  GG_REQUIRE(W1 && b1 && W2 && b2, "edge_mlp_fwd: null parameter");
  GG_REQUIRE(E == 0 || (e && hid && out), "edge_mlp_fwd: null edge buffer");
  GG_REQUIRE(reinterpret_cast<uintptr_t>(e) % 16 == 0 && reinterpret_cast<uintptr_t>(W1) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(W2) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(hid) % 16 == 0,
             "edge_mlp_fwd: e, W1, W2, hid and out must be 16-byte aligned");
  if (E == 0) return GG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  int64_t blocks = (E + 255) / 256;                           // 8 warps x 32 rows per CTA and iteration
  if (blocks > 2LL * sms) blocks = 2LL * sms;
  GG_KERNEL_BEGIN("edge_mlp_fwd_kernel", st);
  if (d == 64) edge_mlp_fwd_kernel<64><<<(unsigned)blocks, kMlpThreads, 0, st>>>(E, e, W1, b1, W2, b2, hid, out);
  else edge_mlp_fwd_kernel<128><<<(unsigned)blocks, kMlpThreads, 0, st>>>(E, e, W1, b1, W2, b2, hid, out);
  GG_KERNEL_END("edge_mlp_fwd_kernel", st);
  return GG_OK;
}

extern "C" int gg_edge_mlp_bwd(int64_t E, int d, int hidden, int K, const float* g, const float* hid, const float* e,
                               const float* W2, float* dW1, float* db1, float* dW2, float* db2, void* stream) {
  GG_REQUIRE(E >= 0, "edge_mlp_bwd: negative size");
  if (hidden != kMlpHid || K != kMlpK || !(d == 64 || d == 128)) {
    set_error("gnnome_b200: edge_mlp_bwd is built for hidden_edge_features = 16, K = 4 (2 padded), d in {64,128}");
    return GG_ERR_UNSUPPORTED;
  }
  GG_REQUIRE(W2 && dW1 && db1 && dW2 && db2, "edge_mlp_bwd: null parameter / output");
  GG_REQUIRE(E == 0 || (g && hid && e), "edge_mlp_bwd: null edge buffer");
  cudaStream_t st = (cudaStream_t)stream;
  if (!prezeroed()) {                                   // gg_model_bwd zeroes the whole gradient arena in one memset
    GG_CUDA(cudaMemsetAsync(dW2, 0, sizeof(float) * (size_t)d * hidden, st));
    GG_CUDA(cudaMemsetAsync(db2, 0, sizeof(float) * (size_t)d, st));
    GG_CUDA(cudaMemsetAsync(dW1, 0, sizeof(float) * (size_t)hidden * K, st));
    GG_CUDA(cudaMemsetAsync(db1, 0, sizeof(float) * (size_t)hidden, st));
  }
  if (E == 0) return GG_OK;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  int64_t blocks = (E + 7) / 8;
  if (blocks > sms) blocks = sms;                             // one persistent CTA per SM (register-heavy)
  GG_KERNEL_BEGIN("edge_mlp_bwd_kernel", st);
  if (d == 64) edge_mlp_bwd_kernel<64><<<(unsigned)blocks, kMlpThreads, 0, st>>>(E, g, hid, e, W2, dW1, db1, dW2, db2);
  else edge_mlp_bwd_kernel<128><<<(unsigned)blocks, kMlpThreads, 0, st>>>(E, g, hid, e, W2, dW1, db1, dW2, db2);
  GG_KERNEL_END("edge_mlp_bwd_kernel", st);
  return GG_OK;
}
