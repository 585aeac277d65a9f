// gg_common.cuh — shared device/host helpers for the GatedGCN engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/gnnome_b200.h"

namespace gg {

// ---------------------------------------------------------------- host-side error plumbing
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

#define GG_CUDA(call)                                              \
  do {                                                             \
    cudaError_t _e = (call);                                       \
    if (_e != cudaSuccess) return gg::cuda_fail(_e, #call);        \
  } while (0)

// every kernel launch is bracketed by these two: they count launches (gg_launch_count) and, when
// profiling is on (gg_profile_enable), time the launch with a CUDA event pair on the launching stream
void kernel_begin(const char* name, cudaStream_t st);
void kernel_end(const char* name, cudaStream_t st);

#define GG_KERNEL_BEGIN(name, st) gg::kernel_begin(name, st)
#define GG_KERNEL_END(name, st)                                    \
  do {                                                             \
    gg::kernel_end(name, st);                                      \
    cudaError_t _e = cudaGetLastError();                           \
    if (_e != cudaSuccess) return gg::cuda_fail(_e, name);         \
  } while (0)

#define GG_REQUIRE(cond, msg)                                      \
  do {                                                             \
    if (!(cond)) {                                                 \
      gg::set_error(std::string("gnnome_b200: ") + msg);           \
      return GG_ERR_ARG;                                           \
    }                                                              \
  } while (0)

// device ordinal of the calling thread, clamped: index of the per-device host caches
constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  return dev;
}

struct Plan {
  int64_t N = 0, E = 0;
  int32_t* src = nullptr;      // [E] internal (dst-sorted) order
  int32_t* dst = nullptr;      // [E]
  int32_t* in_ptr = nullptr;   // [N+1] CSR over in-edges: internal ids [in_ptr[v], in_ptr[v+1])
  int32_t* out_ptr = nullptr;  // [N+1] CSR over out-edges
  int32_t* out_eid = nullptr;  // [E] internal edge id of the j-th out-edge slot
  int32_t* out_dst = nullptr;  // [E] dst of that edge (saves one indirection)
  int32_t* perm = nullptr;     // [E] internal position -> caller edge id
  int32_t* inv_perm = nullptr; // [E]
  int32_t* node_perm = nullptr; // [N] internal node id -> caller node id
  int32_t* node_inv = nullptr;  // [N] caller node id -> internal node id
  int num_sms = 148;
  // sub-graph plans (gg_subplan_fill): every array above lives in ONE caller-owned slab; plus the caller-facing view
  int32_t* slab = nullptr;
  int32_t* host_slab = nullptr;   // gg_plan_create: the library-owned allocation behind all arrays
  int32_t* parent_eid = nullptr;  // [E] caller edge id -> parent caller edge id (dgl.EID)
  int32_t* csrc = nullptr;        // [E] caller-order edge list in sub-graph node ids (sub_g.edges())
  int32_t* cdst = nullptr;
  void* sub_scratch = nullptr;    // SubScratch of a plan that has been used as a parent
};

// gg_plan_device.cu: the whole-graph plan built on the device (src / dst: device int32[E], caller edge order)
int plan_create_device(const int32_t* src, const int32_t* dst, int64_t N, int64_t E, int flags, cudaStream_t st, Plan** out);

#define GG_TRY_RC(call)       \
  do {                        \
    int _rc = (call);         \
    if (_rc) return _rc;      \
  } while (0)

// gg_api.cu: zig-zag base direction of the layer whose forward is issued next (set by the whole-model sequencer)
void set_layer_parity(int p);
void set_prezeroed(bool on);    // accumulators were zeroed by the caller (gg_model_*): skip the per-op memsets
bool prezeroed();

int gg_debug_flags_peek();      // current gg_debug_flags value (gg_api.cu)
// gg_api.cu: side stream + events for the weight-gradient GEMMs of the gg_layer_bwd call issued next (null = in line)
void set_layer_bwd_side(cudaStream_t side, cudaEvent_t fork, cudaEvent_t done);

struct SubScratch;
void free_sub_scratch(SubScratch* s);

constexpr float kAggEps = 1e-6f;   // gated_gcn_full.py:130,143
constexpr float kNormEps = 1e-5f;  // nn.BatchNorm1d / nn.LayerNorm default eps

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

// A feature row of D floats spread over one warp: VPL = D/32 values per lane.
//   D = 64 : lane holds channels {2*lane, 2*lane+1}                       (one 64-bit access)
//   D >= 128: lane holds channels {128*j + 4*lane .. +3}, j < D/128       (128-bit accesses)
// Every warp-wide access touches whole 128-byte lines.
template <int D>
struct Row {
  static constexpr int VPL = D / 32;
  float v[VPL];

  __device__ __forceinline__ void load(const float* __restrict__ row, int lane) {
    if constexpr (D == 64) {
      float2 t = __ldg(reinterpret_cast<const float2*>(row) + lane);
      v[0] = t.x; v[1] = t.y;
    } else {
#pragma unroll
      for (int j = 0; j < D / 128; ++j) {
        float4 t = __ldg(reinterpret_cast<const float4*>(row + 128 * j) + lane);
        v[4 * j + 0] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
      }
    }
  }
  // streaming variant for E x d tensors that are touched once per kernel (do not pollute L1)
  __device__ __forceinline__ void load_stream(const float* __restrict__ row, int lane) {
    if constexpr (D == 64) {
      float2 t = __ldcs(reinterpret_cast<const float2*>(row) + lane);
      v[0] = t.x; v[1] = t.y;
    } else {
#pragma unroll
      for (int j = 0; j < D / 128; ++j) {
        float4 t = __ldcs(reinterpret_cast<const float4*>(row + 128 * j) + lane);
        v[4 * j + 0] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
      }
    }
  }
  __device__ __forceinline__ void store(float* __restrict__ row, int lane) const {
    if constexpr (D == 64) {
      reinterpret_cast<float2*>(row)[lane] = make_float2(v[0], v[1]);
    } else {
#pragma unroll
      for (int j = 0; j < D / 128; ++j)
        reinterpret_cast<float4*>(row + 128 * j)[lane] =
            make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
  __device__ __forceinline__ void fill(float x) {
#pragma unroll
    for (int k = 0; k < VPL; ++k) v[k] = x;
  }
  // channel index of element k held by `lane`
  __device__ __forceinline__ static int channel(int k, int lane) {
    if constexpr (D == 64) return 2 * lane + k;
    else return 128 * (k >> 2) + 4 * lane + (k & 3);
  }
};

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

#endif  // __CUDACC__
}  // namespace gg
