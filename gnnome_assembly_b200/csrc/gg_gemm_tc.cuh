// gg_gemm_tc.cuh — 3xTF32 GEMM on the Blackwell tensor cores (tcgen05.mma kind::tf32, accumulators AND the
// A operand in TMEM, B operand and raw A tiles staged by TMA with 128-byte swizzle), with the same pluggable
// epilogues as the FFMA GEMM.
//
// Why 3xTF32: the reference runs its projections in true fp32 (torch allow_tf32=False,
// layers/gated_gcn_full.py:44,107-113) and plain TF32 misses the 1e-4 logit tolerance.
// Each fp32 operand x is split as x = hi + lo, hi = x with the low 13 mantissa bits dropped (what kind::tf32
// reads anyway), lo = x - hi (exact in fp32); D = A_lo B_hi + A_hi B_lo + A_hi B_hi drops only lo*lo (~2^-20).
//
// Data flow per 128 x 128 x 32 K-block (32 fp32 = one 128-byte swizzle row), 4-stage mbarrier ring:
//   TMA      : raw A tile (16 KB) and raw B tile (16 KB) -> shared memory
//   converter: 4 warps, thread = A row: row -> registers -> (hi, lo) -> TMEM via tcgen05.st  (A never goes
//              back to shared memory; the MMA reads it from TMEM, "TS" form) ; B_lo = B - trunc(B) is written
//              next to B (same swizzled byte layout, so the split is layout agnostic)
//   MMA      : one thread, 3 x 4 tcgen05.mma per K-block, fp32 accumulators in TMEM (2 x 128 columns, so the
//              epilogue of tile i overlaps the main loop of tile i+1)
//   epilogue : 8 warps in two groups of 64 columns: TMEM -> registers -> shared staging -> row-wise coalesced
//              global traffic; the fused epilogue's gathered operands are prefetched into registers a whole
//              tile ahead ("deep") / one 16-column chunk ahead ("near"), register budget via setmaxnreg.
// Shared-memory traffic per K-block: 32 KB TMA + 16 KB (A read) + 32 KB (B split) + 48 KB (MMA reads B) —
// the SS form (A_hi/A_lo in shared memory) needed 192 KB and was bound by shared-memory bandwidth
// (profiles/r1_ncu_full_fwd_kernels_tc_ss.txt: l1tex 62-73 %, tensor pipe 16-30 %).
// Persistent CTAs, one per SM.  Operand layouts (template flags): K-major = reduction index contiguous in
// global memory (A[M,K] row-major, nn.Linear W[N,K]); MN-major = the M / N index contiguous (B[K,N] in
// bwd-data, both operands in weight gradients).  Descriptor encodings follow CUTLASS
// cute/arch/mma_sm100_desc.hpp.
#pragma once
#include <cuda.h>

#include "gg_common.cuh"
#include "gg_gemm_ffma.cuh"

namespace gg {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 32;        // BK fp32 = 128 bytes = one swizzle row
constexpr int STAGES = 4;                         // barrier-slot stride (3 stages are in use)
constexpr int TILE_BYTES = BM * BK * 4;           // 16 KB (A tile == B tile size since BM == BN)
constexpr int COEF_BYTES = 8 * 256 * 4;           // A-transform coefficients: 2 float4 per channel, K <= 256 channels
constexpr int STATS_BYTES = 4 * BN * 2 * 8;       // [4 lane quarters][128 cols][sum, sumsq] doubles
constexpr int IDX_BYTES = 2 * 2 * 2 * BM * 4;     // per epilogue group, double buffered: src / dst node ids of 128 rows
constexpr int BAR_BYTES = 256;
// Shared-memory budget (227 KB), three layouts:
//   plain      : 4 stages of A raw | B hi (raw) | B lo = 4 x 48 KB, staging 2 x 8 KB (16-column chunks)
//   wide       : 3 stages (144 KB), which leaves room for a staging tile of the epilogue group's WHOLE 64 columns
//                (2 x 32 KB): one barrier pair per tile instead of one per 16-column chunk, 256-byte row segments
//                in every global access, column statistics reduced once per tile
//   A transform: 3 stages that also hold the second streamed A tile = 3 x 64 KB; staging 2 x 8 KB
template <class ATx, class Epi, bool kWRes = false>
struct Cfg {
  // wide staging pays for plain stores of wide outputs and is required by the row-reducing score epilogue: measured
  // gemm_node_proj 67 -> 61 us (gemm_edge_gate 159 -> 154 us at d = 128, where W-residency is used instead; at
  // d = 256 it lost 3-4 % in a same-box A/B and is off); the split-K weight gradients and the bwd-data GEMMs have
  // long K loops and lose more from the fourth stage than they gain (79 -> 88 us)
#ifdef GG_NO_WIDE            // A/B builds (tools/ab_build.sh): the layout before the per-epilogue configuration
  static constexpr bool kWide = !ATx::kActive && Epi::kRowReduce;      // (the row reduction needs the wide tile)
#else
  static constexpr bool kWide = !ATx::kActive && Epi::kWideStaging && !kWRes;
#endif
  // W-resident (kWRes): the whole B operand (a d x d weight, K <= 128: 4 K-blocks of B_hi and of B_lo = 128 KB)
  // is loaded and split ONCE per CTA; the ring then carries only A tiles.  The kernel is bound by the LSU /
  // shared-memory pipe (ncu: l1tex 66-70 % busy, everything else < 40 %), and per K-block the B side was 16 KB of
  // TMA writes + 32 KB of split traffic out of ~128 KB.
  // mid staging (fused gather epilogues: EpiEdgeGate): 32-column chunks, i.e. one whole 128-byte line per row and chunk.
  // Per-role cycle traces (profiles/r2_tc_trace_d128.txt) show this kernel paced by its epilogue, and the epilogue by
  // the LSU wavefront count: with 16-column chunks every gathered / stored row piece is HALF a line (64 bytes), so the
  // 128 KB of B1h[src] / B2h[dst] gathers and the 64 KB of t stores of a tile cost twice the wavefronts they need.
#ifdef GG_NO_MID
  static constexpr bool kMid = false;
#else
  static constexpr bool kMid = !ATx::kActive && !kWide && Epi::kMidStaging;
#endif
  static constexpr int kStages = kWRes ? (ATx::kActive ? 2 : (kMid ? 3 : 4)) : ((ATx::kActive || kWide || kMid) ? 3 : 4);
  static constexpr int kATiles = ATx::kActive ? 2 : 1;
  static constexpr int kStageBytes = (kWRes ? kATiles : kATiles + 2) * TILE_BYTES;
  static constexpr int kOffA2 = kWRes ? TILE_BYTES : 3 * TILE_BYTES;        // second streamed A tile (A transform)
  static constexpr int kBResBytes = kWRes ? 8 * TILE_BYTES : 0;            // [4 K-blocks hi][4 K-blocks lo]
  static constexpr int kPipeBytes = kStages * kStageBytes + kBResBytes;
  static constexpr int kEC = kWide ? 64 : (kMid ? 32 : 16);       // epilogue chunk width in accumulator columns
  static constexpr int kStagingBytes = 2 * BM * kEC * 4;
};
template <bool kStats, class Epi, class ATx, bool kWRes = false>
constexpr int smem_bytes() {
  return 1024 /*align slack*/ + Cfg<ATx, Epi, kWRes>::kPipeBytes + Cfg<ATx, Epi, kWRes>::kStagingBytes + (kStats ? STATS_BYTES : 0) +
         (Epi::kIdx ? IDX_BYTES : 0) + BAR_BYTES + (ATx::kActive ? COEF_BYTES : 0);
}
constexpr int THREADS = 512;                      // 4 control + 4 converter + 8 epilogue warps
constexpr int THREADS_ATX = 640;                  // an A transform doubles the converter: 4 + 8 + 8 warps
template <class ATx> constexpr int threads() { return ATx::kActive ? THREADS_ATX : THREADS; }
constexpr int TMEM_COLS = 512;                    // 2 x 128 accumulator columns + 4 stages x (32 hi + 32 lo) A columns
constexpr int TMEM_A0 = 256;

int& tc_dbg_ref();                                // experiment switches, see Args::dbg

// ---------------------------------------------------------------------------------- role tracing (-DGG_TC_TRACE)
// Where does a tile's time go?  ncu's stall sampling cannot tell a role that waits from a role that works slowly.
// A trace build (tools/ab_build.sh trace "-DGG_TC_TRACE", tools/tc_trace.py) accumulates, per role, the clock64
// cycles spent inside each mbarrier wait and in total, summed over CTAs into gg_tc_trace[] (read and reset through
// gg_debug_trace).  In the default build every call below is empty.
#ifdef GG_TC_TRACE
constexpr bool kTrace = true;
#else
constexpr bool kTrace = false;
#endif
enum TraceSlot { TR_PROD_WAIT_EMPTY = 0, TR_PROD_TOTAL, TR_MMA_WAIT_AB, TR_MMA_WAIT_ACC, TR_MMA_TOTAL, TR_CONV_WAIT_RAW,
                 TR_CONV_TOTAL, TR_EPI_WAIT_ACC, TR_EPI_CHUNKS, TR_EPI_FETCH, TR_EPI_TOTAL, TR_CTAS, TR_TILES, TR_SLOTS };
__device__ unsigned long long gg_tc_trace[TR_SLOTS];      // this header is compiled into exactly one translation unit (gg_api.cu)
struct Tracer {
  unsigned long long t0 = 0, t_role = 0;
  bool on = true;      // runtime filter (gg_debug_flags bit 4: kernels with an A transform only; bit 5: edge-gate epilogue only)
  __device__ __forceinline__ void role_begin() { if constexpr (kTrace) if (on) t_role = clock64(); }
  __device__ __forceinline__ void role_end(int slot) { if constexpr (kTrace) if (on) atomicAdd(&gg_tc_trace[slot], clock64() - t_role); }
  __device__ __forceinline__ void begin() { if constexpr (kTrace) if (on) t0 = clock64(); }
  __device__ __forceinline__ void end(int slot) { if constexpr (kTrace) if (on) atomicAdd(&gg_tc_trace[slot], clock64() - t0); }
  __device__ __forceinline__ void count(int slot, unsigned long long n) { if constexpr (kTrace) if (on) atomicAdd(&gg_tc_trace[slot], n); }
};

struct Args {
  int64_t M; int N; int64_t K;
  int m_tiles, n_tiles, splits;
  int64_t k_chunk;           // multiple of BK
  double* col_stats;         // kStats: [2N]
  float* bias_grad;          // kBiasGrad (A MN-major only): [M] sums of A over k
  int dbg;                   // experiment switches (tools/epi_experiment.py): 1 = skip column stats, 2 = skip epilogue prefetches
  int rev;                   // walk the row tiles from the last to the first (zig-zag traversal, gg_api.cu)
};

// ---------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand from TMEM ("TS" form): lane = row, one tf32 per 32-bit column
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// split form: issue the load now, wait once for several of them.  The wait carries the destination registers as
// read-write operands so that no use of them can be scheduled above it.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_fence(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :: "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address [0,14), LBO [16,30),
// SBO [32,46) (all >> 4), version = 1 at [46,48), layout type SWIZZLE_128B = 2 at [61,64)
// K-major operands use SWIZZLE_128B (16-byte chunks permuted over 8 rows).  MN-major TF32 operands
// have exactly one legal layout, SWIZZLE_128B_BASE32B = 1 (32-byte chunks permuted over 4 rows, CuTe
// Layout_MN_SW128_32B_Atom; TMA mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).
template <bool MN>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  // K-major : rows 128 B apart, 8-row groups 1 KB apart (SBO); LBO unused (1)
  // MN-major: 32-wide MN blocks 4 KB apart (LBO, one TMA box of 32 k-rows each); 4-row K groups 512 B apart (SBO)
  constexpr uint32_t lbo_bytes = MN ? 4096 : 16, sbo_bytes = MN ? 512 : 1024;
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(MN ? 1 : 2) << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10),
// a_major bit 15, b_major bit 16 (1 = MN-major), N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ---------------------------------------------------------------------------------- A-operand transforms
// The converter owns whole A rows in registers on their way to TMEM, so an element-wise producer of the A
// operand can be fused there instead of materialising A in HBM first.
struct NoATx { static constexpr bool kActive = false; };

// Batch-norm backward of the edge gate (layers/gated_gcn_full.py:122-123 differentiated): the A operand of
// g_e_in = g_eo + g_t B3 is  g_t = gamma rstd (g_n - m1 - xhat m2),  g_n = g_eo [n > 0],  xhat = (t - mean) rstd,
// n = gamma xhat + beta, built from the streamed tiles of g_eo (TMA map A) and t (TMA map A2).  The converter
// also stores g_t (the weight-gradient GEMM and the out-edge pass read it), so the separate edge_bwd_b pass
// (read t, g_eo; write g_t) disappears.
struct BnBwdATx {
  static constexpr bool kActive = true;
  const double* stats;      // [2K] sum t, sum t^2
  const double* bstats;     // [2K] sum g_n, sum g_n xhat
  const float* gamma; const float* beta;
  double inv_count;         // 1 / E
  float* g_t; int64_t ld;
};

// ---------------------------------------------------------------------------------- the kernel
template <bool A_MN, bool B_MN, bool kStats, bool kBiasGrad, class Epi, class ATx, bool kWRes = false, bool kNarrow = false>
__global__ void __launch_bounds__(threads<ATx>(), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmOut2, Args g, Epi epi, ATx atx) {
  static_assert(!ATx::kActive || !A_MN, "A transforms are written for K-major A");
  constexpr int kStages = Cfg<ATx, Epi, kWRes>::kStages;
  constexpr int kStageBytes = Cfg<ATx, Epi, kWRes>::kStageBytes;
  constexpr int kOffA2 = Cfg<ATx, Epi, kWRes>::kOffA2;
  static_assert(!kWRes || !A_MN, "W-resident mode is for K-major A (forward / bwd-data)");
  // warp roles: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 3 idle | converters | 8 epilogue warps.
  // With an A transform the converter is the longest stage of the pipeline (ncu: its warps never wait), so it
  // gets 8 warps: warps w and w + 4 share a TMEM lane quarter and each takes 16 of the K-block's 32 columns.
  constexpr int kConvWarps = ATx::kActive ? 8 : 4;
  constexpr int kConvThreads = 32 * kConvWarps;
  constexpr int kEpiWarp0 = 4 + kConvWarps;          // multiple of 4: epilogue warp w reads TMEM lanes 32 (w % 4) ..
  constexpr int kEpiThread0 = 32 * kEpiWarp0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;                 // swizzle atoms need 1 KB alignment
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t stage0 = base;
  constexpr int OFF_STG = Cfg<ATx, Epi, kWRes>::kPipeBytes, OFF_STAT = OFF_STG + Cfg<ATx, Epi, kWRes>::kStagingBytes, OFF_IDX = OFF_STAT + (kStats ? STATS_BYTES : 0),
                OFF_BAR = OFF_IDX + (Epi::kIdx ? IDX_BYTES : 0), OFF_COEF = OFF_BAR + BAR_BYTES;
  float4* coef = reinterpret_cast<float4*>(gen + OFF_COEF);   // [K][2]: {p0, p1, gamma, beta}, {q0, q1, q2, -} (ATx only)
  float* staging = reinterpret_cast<float*>(gen + OFF_STG);
  double* sstat = reinterpret_cast<double*>(gen + OFF_STAT);
  int* sidx = reinterpret_cast<int*>(gen + OFF_IDX);
  const uint32_t bars = base + OFF_BAR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + OFF_BAR + 128);
  auto full_raw = [&](int s) { return bars + 8u * s; };
  auto full_ab = [&](int s) { return bars + 8u * (STAGES + s); };
  auto empty = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto tmem_full = [&](int a) { return bars + 8u * (3 * STAGES + a); };
  auto tmem_empty = [&](int a) { return bars + 8u * (3 * STAGES + 2 + a); };
  const uint32_t bres_full = bars + 8u * 17, bres_ready = bars + 8u * 18;     // W-resident B: loaded / split
  auto gt_full = [&](int s) { return bars + 8u * (19 + s); };                   // A transform: g_t tile of stage s is in smem
  const uint32_t bres0 = stage0 + kStages * kStageBytes;                         // B_hi K-blocks, then B_lo K-blocks
  const int nkb_total = (int)((g.K + BK - 1) / BK);                              // kWRes: <= 4

  // register split (setmaxnreg): a fused epilogue that holds a whole tile of prefetched operands gets more
  constexpr bool kHeavyEpi = (sizeof(typename Epi::PreD) + sizeof(typename Epi::PreN)) >= 32;
  constexpr int kCtrlRegs = ATx::kActive ? 24 : (kHeavyEpi ? 24 : 40), kConvRegs = ATx::kActive ? 72 : (kHeavyEpi ? 96 : 104),
                kEpiRegs = ATx::kActive ? 152 : (kHeavyEpi ? 192 : 184);
  // setmaxnreg moves registers inside the pool the CTA was LAUNCHED with (threads x launch registers), not the
  // whole register file: 640 x 96 = 61440 >= 128 x 24 + 256 x 72 + 256 x 152; 512 x 128 = 65536 >= 128 x 24 + 128 x 96 + 256 x 192
  static_assert(128 * kCtrlRegs + kConvThreads * kConvRegs + 256 * kEpiRegs <= (ATx::kActive ? 640 * 96 : 512 * 128),
                "setmaxnreg split exceeds the CTA's register pool: setmaxnreg.inc would block forever");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t total_work = (int64_t)g.m_tiles * g.n_tiles * g.splits;
  const bool trace_this = kTrace && (!(g.dbg & 16) || ATx::kActive) && (!(g.dbg & 32) || (Epi::kIdx && !Epi::kRowReduce));

  if (threadIdx.x == 0) {
    // a stage is free when its MMAs have completed AND (A transform) the bulk store of its g_t tile has read shared memory
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_raw(s), 1); mbar_init(full_ab(s), kConvThreads); mbar_init(empty(s), ATx::kActive ? 2 : 1);
      if constexpr (ATx::kActive) mbar_init(gt_full(s), kConvThreads);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(tmem_full(a), 1); mbar_init(tmem_empty(a), 256); }
    if constexpr (kWRes) { mbar_init(bres_full, 1); mbar_init(bres_ready, kConvThreads); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if constexpr (kStats) {
    if (warp >= kEpiWarp0) {
      for (int i = threadIdx.x - kEpiThread0; i < 4 * BN * 2; i += 256) sstat[i] = 0.0;
    }
  }
  if constexpr (ATx::kActive) {
    const int K = (int)g.K;
    for (int c = threadIdx.x; c < K; c += (int)blockDim.x) {
      const double m = atx.stats[c] * atx.inv_count;
      double var = atx.stats[K + c] * atx.inv_count - m * m;
      var = var > 0.0 ? var : 0.0;
      // xhat = t p0 + p1 ; n = xhat gamma + beta ; g_t = q0 g_n - q1 - xhat q2
      const float rs = (float)(1.0 / sqrt(var + (double)kNormEps)), gm = __ldg(atx.gamma + c);
      coef[2 * c + 0] = make_float4(rs, -(float)m * rs, gm, __ldg(atx.beta + c));
      coef[2 * c + 1] = make_float4(gm * rs, gm * rs * (float)(atx.bstats[c] * atx.inv_count),
                                    gm * rs * (float)(atx.bstats[K + c] * atx.inv_count), 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work item -> (m_tile, n_tile, split): n fastest so that consecutive CTAs share the A rows in L2
  auto decode = [&](int64_t w, int& mt, int& nt, int& sp) {
    nt = (int)(w % g.n_tiles);
    const int64_t r = w / g.n_tiles;
    mt = (int)(r % g.m_tiles);
    if (g.rev) mt = g.m_tiles - 1 - mt;
    sp = (int)(r / g.m_tiles);
  };
  auto k_range = [&](int sp, int64_t& kbeg, int& nkb) {
    kbeg = (int64_t)sp * g.k_chunk;
    int64_t kend = kbeg + g.k_chunk;
    if (kend > g.K) kend = g.K;
    nkb = (int)((kend - kbeg + BK - 1) / BK);
  };

  // W-resident B: split the whole operand once (element-wise at identical swizzled offsets, like the per-stage split)
  auto split_resident_b = [&](int ct) {
    if (blockIdx.x >= total_work) return;
    mbar_wait(bres_full, 0u);
    const float4* bh = reinterpret_cast<const float4*>(gen + kStages * kStageBytes);
    float4* bl = reinterpret_cast<float4*>(gen + kStages * kStageBytes + 4 * TILE_BYTES);
    const int total = nkb_total * (TILE_BYTES / 16);
    for (int q = ct; q < total; q += kConvThreads) {
      const float4 x = bh[q];
      float4 l;
      l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
      l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
      l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
      l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
      bl[q] = l;
    }
    proxy_fence_async();                               // generic-proxy smem writes -> visible to the tensor core
    mbar_arrive(bres_ready);
  };

  if (warp < 4) {
    reg_dec<kCtrlRegs>();
    if (warp == 0 && lane == 0) {
      // ================================================================ TMA producer
      int s = 0; uint32_t ph = 0;
      Tracer tr;
      tr.on = trace_this;
      tr.role_begin();
      tr.count(TR_CTAS, 1);
      if constexpr (kWRes) {
        if (blockIdx.x < total_work) {
          mbar_expect_tx(bres_full, (uint32_t)nkb_total * TILE_BYTES);
          for (int kb = 0; kb < nkb_total; ++kb) {
            if constexpr (!B_MN) {
              tma_load_2d(bres0 + kb * TILE_BYTES, &tmB, bres_full, kb * BK, 0);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) tma_load_2d(bres0 + kb * TILE_BYTES + j * 4096, &tmB, bres_full, 32 * j, kb * BK);
            }
          }
        }
      }
      for (int64_t w = blockIdx.x; w < total_work; w += gridDim.x) {
        int mt, nt, sp; decode(w, mt, nt, sp);
        int64_t kbeg; int nkb; k_range(sp, kbeg, nkb);
        for (int kb = 0; kb < nkb; ++kb) {
          tr.begin();
          mbar_wait(empty(s), ph ^ 1u);
          tr.end(TR_PROD_WAIT_EMPTY);
          const uint32_t st = stage0 + s * kStageBytes;
          mbar_expect_tx(full_raw(s), (Cfg<ATx, Epi, kWRes>::kATiles + (kWRes ? 0 : 1)) * TILE_BYTES);
          if constexpr (ATx::kActive) tma_load_2d(st + kOffA2, &tmA2, full_raw(s), (int)(kbeg + (int64_t)kb * BK), mt * BM);
          const int k0 = (int)(kbeg + (int64_t)kb * BK);
          if constexpr (!A_MN) {
            tma_load_2d(st, &tmA, full_raw(s), k0, mt * BM);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_load_2d(st + j * 4096, &tmA, full_raw(s), mt * BM + 32 * j, k0);
          }
          if constexpr (!kWRes) {
            if constexpr (!B_MN) {
              tma_load_2d(st + TILE_BYTES, &tmB, full_raw(s), k0, nt * BN);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) tma_load_2d(st + TILE_BYTES + j * 4096, &tmB, full_raw(s), nt * BN + 32 * j, k0);
            }
          }
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
        tr.count(TR_TILES, 1);
      }
      tr.role_end(TR_PROD_TOTAL);
    } else if (warp == 1 && lane == 0) {
      // ================================================================ MMA issuer (one thread)
      constexpr uint32_t idesc = make_idesc(false, B_MN);          // A comes from TMEM: always K-major there
      int s = 0; uint32_t ph = 0; int acc = 0; uint32_t acc_ph = 0;
      Tracer tr;
      tr.on = trace_this;
      tr.role_begin();
      if constexpr (kWRes) {
        if (blockIdx.x < total_work) { mbar_wait(bres_ready, 0u); tc_fence_after(); }
      }
      for (int64_t w = blockIdx.x; w < total_work; w += gridDim.x) {
        int mt, nt, sp; decode(w, mt, nt, sp);
        int64_t kbeg; int nkb; k_range(sp, kbeg, nkb);
        tr.begin();
        mbar_wait(tmem_empty(acc), acc_ph ^ 1u);
        tr.end(TR_MMA_WAIT_ACC);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < nkb; ++kb) {
          tr.begin();
          mbar_wait(full_ab(s), ph);
          tr.end(TR_MMA_WAIT_AB);
          tc_fence_after();
          const uint32_t st = stage0 + s * kStageBytes;
          const uint32_t a_hi = tmem_base + (uint32_t)(TMEM_A0 + 64 * s), a_lo = a_hi + 32;
          const uint64_t b_hi = make_desc<B_MN>(kWRes ? bres0 + kb * TILE_BYTES : st + TILE_BYTES);
          const uint64_t b_lo = make_desc<B_MN>(kWRes ? bres0 + (4 + kb) * TILE_BYTES : st + 2 * TILE_BYTES);
          constexpr uint64_t b_step = B_MN ? (1024 >> 4) : (32 >> 4);   // 8 tf32 along K: 8 k-rows / 32 bytes
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            const uint32_t first = (kb == 0 && ks == 0) ? 0u : 1u;
            umma_tf32_ts(d_tmem, a_lo + 8 * ks, b_hi + ks * b_step, idesc, first);   // small terms first
            umma_tf32_ts(d_tmem, a_hi + 8 * ks, b_lo + ks * b_step, idesc, 1u);
            umma_tf32_ts(d_tmem, a_hi + 8 * ks, b_hi + ks * b_step, idesc, 1u);
          }
          umma_commit(empty(s));                       // frees the stage (smem B, TMEM A) when these MMAs are done
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
        umma_commit(tmem_full(acc));                   // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
      }
      tr.role_end(TR_MMA_TOTAL);
    } else if (ATx::kActive && warp == 3 && lane == 0) {
      // ================================================================ g_t store (A transform only)
      // The converter threads own one A row each, so a direct global store of g_t is 32 scattered 16-byte pieces per
      // warp instruction (32 LSU wavefronts per 512 bytes: the single largest item on the kernel's busiest pipe).  They
      // write g_t back into the stage's `t` slot instead (same swizzled positions they just read), and this thread
      // sends the finished 128 x 32 tile to global memory with ONE bulk tensor store; the stage is released to the
      // producer when the store has finished reading shared memory.
      if constexpr (ATx::kActive) {
        int s = 0; uint32_t ph = 0;
        for (int64_t w = blockIdx.x; w < total_work; w += gridDim.x) {
          int mt, nt, sp; decode(w, mt, nt, sp);
          int64_t kbeg; int nkb; k_range(sp, kbeg, nkb);
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(gt_full(s), ph);
            if (nt == 0) {
              const uint32_t st = stage0 + s * kStageBytes + kOffA2;
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                           ::"l"(&tmOut2), "r"(st), "r"((int)(kbeg + (int64_t)kb * BK)), "r"(mt * BM) : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            mbar_arrive(empty(s));
            if (++s == kStages) { s = 0; ph ^= 1u; }
          }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");       // the stores have landed before the CTA exits
      }
    }
  } else if (ATx::kActive && warp < kEpiWarp0) {
    // ================================================================ converter with an A transform
    // 256 threads: thread = (A row t, half of the K-block's 32 columns)
    reg_dec<kConvRegs>();
    if constexpr (ATx::kActive) {
      const int ct = threadIdx.x - 128;
      const int cw = ct >> 5;
      const int t = 32 * (cw & 3) + lane;
      const int half = cw >> 2;
      const uint32_t lane_base = (uint32_t)(32 * (cw & 3)) << 16;
      int s = 0; uint32_t ph = 0;
      Tracer tr;
      tr.on = trace_this;
      const bool tr_on = kTrace && ct == 0;
      if (tr_on) tr.role_begin();
      if constexpr (kWRes) split_resident_b(ct);
      for (int64_t w = blockIdx.x; w < total_work; w += gridDim.x) {
        int mt, nt, sp; decode(w, mt, nt, sp);
        int64_t kbeg; int nkb; k_range(sp, kbeg, nkb);
        const int64_t m = (int64_t)mt * BM + t;
        for (int kb = 0; kb < nkb; ++kb) {
          if (tr_on) tr.begin();
          mbar_wait(full_raw(s), ph);
          if (tr_on) tr.end(TR_CONV_WAIT_RAW);
          const uint8_t* st = gen + s * kStageBytes;
          // g_eo[row, c0 + 16 half ..+15] (tile 0) and t (tile 3) -> g_t (stored, and the A operand)
          const float4* row = reinterpret_cast<const float4*>(st + t * 128);
          float4* row2 = reinterpret_cast<float4*>(const_cast<uint8_t*>(st) + kOffA2 + t * 128);
          const int c0 = (int)(kbeg + (int64_t)kb * BK);
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int cc = 4 * half + c;                                // logical 16-byte chunk of the 128-byte row
            const float4 gv = row[cc ^ (t & 7)];
            const float4 tv = row2[cc ^ (t & 7)];
            const float ge[4] = {gv.x, gv.y, gv.z, gv.w};
            const float tt[4] = {tv.x, tv.y, tv.z, tv.w};
            float gt[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int ch = c0 + 4 * cc + j;
              const float4 ca = coef[2 * ch], cb = coef[2 * ch + 1];     // warp-uniform address: smem broadcast
              const float xh = fmaf(tt[j], ca.x, ca.y);
              const float gn = fmaf(xh, ca.z, ca.w) > 0.f ? ge[j] : 0.f;
              gt[j] = fmaf(cb.x, gn, -cb.y) - xh * cb.z;
              hi[4 * c + j] = __float_as_uint(gt[j]);
            }
            // g_t replaces t in the stage (this thread's own 16-byte chunk): stored by the bulk-store thread
            row2[cc ^ (t & 7)] = make_float4(gt[0], gt[1], gt[2], gt[3]);
          }
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float x = __uint_as_float(hi[k]);
            lo[k] = __float_as_uint(x - __uint_as_float(hi[k] & 0xffffe000u));
          }
          const uint32_t ta = tmem_base + lane_base + (uint32_t)(TMEM_A0 + 64 * s + 16 * half);
          tmem_st16(ta, hi);
          tmem_st16(ta + 32, lo);
          if constexpr (!kWRes) {
            const float4* bh = reinterpret_cast<const float4*>(st + TILE_BYTES);
            float4* bl = reinterpret_cast<float4*>(const_cast<uint8_t*>(st) + 2 * TILE_BYTES);
#pragma unroll
            for (int i = 0; i < TILE_BYTES / 16 / kConvThreads; ++i) {
              const int q = ct + kConvThreads * i;
              const float4 x = bh[q];
              float4 l;
              l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
              l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
              l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
              l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
              bl[q] = l;
            }
          }
          tmem_st_wait();
          tc_fence_before();
          proxy_fence_async();                           // generic-proxy smem writes (g_t, B_lo) -> async proxy (MMA, bulk store)
          mbar_arrive(full_ab(s));
          mbar_arrive(gt_full(s));
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
      }
      if (tr_on) tr.role_end(TR_CONV_TOTAL);
    }
  } else if (!ATx::kActive && warp < 8) {
    // ================================================================ converter (128 threads, thread = A row)
    reg_dec<kConvRegs>();
    const int t = threadIdx.x - 128;
    const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
    int s = 0; uint32_t ph = 0;
    Tracer tr;
    tr.on = trace_this;
    const bool tr_on = kTrace && t == 0;
    if (tr_on) tr.role_begin();
    if constexpr (kWRes) split_resident_b(t);
    for (int64_t w = blockIdx.x; w < total_work; w += gridDim.x) {
      int mt, nt, sp; decode(w, mt, nt, sp);
      int64_t kbeg; int nkb; k_range(sp, kbeg, nkb);
      float bsum = 0.f;
      for (int kb = 0; kb < nkb; ++kb) {
        if (tr_on) tr.begin();
        mbar_wait(full_raw(s), ph);
        if (tr_on) tr.end(TR_CONV_WAIT_RAW);
        const uint8_t* st = gen + s * kStageBytes;
        // ---- A row t of this K-block -> registers (k order), hi = raw bits, lo = x - trunc(x)
        uint32_t hi[32], lo[32];
        if constexpr (!A_MN) {
          // K-major SWIZZLE_128B tile: row t at t*128 B, logical 16-byte chunk c stored at chunk c ^ (t % 8)
          const float4* row = reinterpret_cast<const float4*>(st + t * 128);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 x = row[c ^ (t & 7)];
            hi[4 * c + 0] = __float_as_uint(x.x); hi[4 * c + 1] = __float_as_uint(x.y);
            hi[4 * c + 2] = __float_as_uint(x.z); hi[4 * c + 3] = __float_as_uint(x.w);
          }
        } else {
          // MN-major 128B_ATOM_32B tile: box t/32 (4 KB), k-row k at k*128 B, m' = t%32 lives in 32-byte chunk
          // (m'/8) ^ (k % 4) at word m' % 8  -> for a fixed k a warp reads one whole 128-byte row: conflict free
          const uint8_t* box = st + (t >> 5) * 4096;
          const int mp = t & 31;
#pragma unroll
          for (int k = 0; k < 32; ++k)
            hi[k] = *reinterpret_cast<const uint32_t*>(box + k * 128 + ((((mp >> 3) ^ (k & 3)) << 5) | ((mp & 7) << 2)));
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float x = __uint_as_float(hi[k]);
          lo[k] = __float_as_uint(x - __uint_as_float(hi[k] & 0xffffe000u));
          if constexpr (kBiasGrad) bsum += x;
        }
        const uint32_t ta = tmem_base + lane_base + (uint32_t)(TMEM_A0 + 64 * s);
        tmem_st32(ta, hi);
        tmem_st32(ta + 32, lo);
        // ---- B_lo = B - trunc(B), element-wise at identical (swizzled) byte offsets
        if constexpr (!kWRes) {
          const float4* bh = reinterpret_cast<const float4*>(st + TILE_BYTES);
          float4* bl = reinterpret_cast<float4*>(const_cast<uint8_t*>(st) + 2 * TILE_BYTES);
#pragma unroll
          for (int i = 0; i < TILE_BYTES / 16 / 128; ++i) {
            const int q = t + 128 * i;
            const float4 x = bh[q];
            float4 l;
            l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
            l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
            l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
            l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
            bl[q] = l;
          }
        }
        tmem_st_wait();
        tc_fence_before();
        proxy_fence_async();                           // generic-proxy smem writes -> visible to the tensor core
        mbar_arrive(full_ab(s));
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
      if constexpr (kBiasGrad) {
        const int64_t m = (int64_t)mt * BM + t;
        if (g.bias_grad != nullptr && nt == 0 && m < g.M) atomicAdd(g.bias_grad + m, bsum);
      }
    }
    if (tr_on) tr.role_end(TR_CONV_TOTAL);
  } else {
    // ================================================================ epilogue (2 groups x 128 threads)
    // group grp owns accumulator columns [64 grp, 64 grp + 64), 4 chunks of 16 columns.  Per chunk: every
    // thread pulls its row (TMEM lane) out of TMEM, parks it in the group's staging tile, then the group
    // walks the tile row-wise (4 threads x float4 = one 64-byte row segment) so that all global traffic of
    // the fused epilogue is coalesced.
    reg_inc<kEpiRegs>();
    const int grp = (warp - kEpiWarp0) >> 2;
    const int ew = warp & 3;                            // TMEM lanes 32*ew .. 32*ew+31
    const int tg = (threadIdx.x - kEpiThread0) & 127;   // thread in group == accumulator row of the tile
    const uint32_t bar_id = 1 + grp;
    // chunk geometry: EC accumulator columns per chunk; a chunk row is TPR float4s, so TPR threads share a row and
    // the group's 128 threads cover RPP rows per pass, PPC passes per chunk; QN chunks per 64-column group
    constexpr int EC = Cfg<ATx, Epi, kWRes>::kEC, TPR = EC / 4, RPP = 128 / TPR, PPC = 128 / RPP, QN = 64 / EC;
    float4* stg = reinterpret_cast<float4*>(staging) + grp * (BM * EC / 4);
    int* idx_base = sidx + grp * 4 * BM;                // [2 buffers][src | dst][128]
    const int tcol = tg % TPR, trow = tg / TPR;             // this thread's float4 column / row inside a pass
    // N <= 64 with chunks narrower than 64 columns (d = 64): the two groups share the 64 valid columns, 32 each, instead
    // of group 1 idling.  (64-column chunks — the score predictor's wide staging — keep the idle group below.)
    // (kNarrow is chosen by launch(): a compile-time switch, so the N = 128 kernels are the code they were.)
    static_assert(!kNarrow || QN >= 2, "narrow mode needs chunks of at most 32 columns");
    constexpr bool narrow = kNarrow;
    auto cbase = [&]() { return narrow ? 32 * grp : 64 * grp; };   // first accumulator column of this group (re-derived at
                                                                  // each use: a live variable changed the N = 128 kernels' allocation)
    constexpr int qn = narrow ? QN / 2 : QN;              // chunks this group walks per tile
    const int c4 = tcol * 4;                              // column offset of this thread inside a chunk
    auto swz = [](int row) { return EC == 16 ? ((row >> 1) & 3) : (row & (TPR - 1)); };   // staging xor-swizzle
    auto group_bar = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); };
    using PreD = typename Epi::PreD;
    using PreN = typename Epi::PreN;
    // The fused epilogue's global operands (gathered projection rows / streamed addends) for the WHOLE next
    // tile are fetched into registers in one batch right after the current tile has been processed, and
    // nothing is fetched while a tile is being consumed: a warp has only six load scoreboards, so a load
    // issued between two uses would make the older, already-landed operands wait for it.  The batch lands
    // while this group waits for the next accumulator (the MMA main loop of a tile is longer than a DRAM trip).
    PreD deep[QN][PPC];                                 // [chunk][pass]: 16 float4 slots either way
    PreN near[QN][PPC];

    auto row_of = [&](int mt_, int p) {
      int64_t m = (int64_t)mt_ * BM + p * RPP + trow;
      return m < g.M ? m : g.M - 1;
    };
    auto col_of = [&](int nt_, int q) { return nt_ * BN + cbase() + EC * q + c4; };
    auto fetch_tile = [&](int mt_, int nt_, int buf) {
      if (g.dbg & 2) return;
#pragma unroll
      for (int p = 0; p < PPC; ++p) {
        const int r = p * RPP + trow;
        int sv = 0, dv = 0;
        if constexpr (Epi::kIdx) { sv = idx_base[buf * 2 * BM + r]; dv = idx_base[buf * 2 * BM + BM + r]; }
        const int64_t m = row_of(mt_, p);
#pragma unroll
        for (int q = 0; q < QN; ++q) {
          if (q < qn) {
            epi.prefetch_deep(deep[q][p], m, col_of(nt_, q), sv, dv);
            epi.prefetch_near(near[q][p], m, col_of(nt_, q), sv, dv);
          }
        }
      }
    };
    auto load_idx = [&](int mt_, int& sv, int& dv) {
      if constexpr (Epi::kIdx) {
        int64_t m = (int64_t)mt_ * BM + tg;
        if (m >= g.M) m = g.M - 1;
        sv = __ldg(epi.src + m);
        dv = __ldg(epi.dst + m);
      }
    };
    auto store_idx = [&](int buf, int sv, int dv) {
      if constexpr (Epi::kIdx) { idx_base[buf * 2 * BM + tg] = sv; idx_base[buf * 2 * BM + BM + tg] = dv; }
    };

    Tracer tr;
    tr.on = trace_this;
    const bool tr_on = kTrace && tg == 0 && grp == 0;
    if (tr_on) tr.role_begin();
    int acc = 0; uint32_t acc_ph = 0;
    int cur = 0;                                        // idx buffer of the current tile
    int64_t w = blockIdx.x;
    if (!narrow && g.n_tiles == 1 && 64 * grp >= g.N) {
      // N <= 64 (the score predictor's hidden layer): this group's 64 columns are never valid.  It only hands the
      // accumulators back; its group barriers, staging and idx buffers are its own, so skipping them is consistent.
      for (; w < total_work; w += gridDim.x) {
        mbar_wait(tmem_full(acc), acc_ph);
        tc_fence_after();
        tc_fence_before();
        mbar_arrive(tmem_empty(acc));
        if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
      }
    }
    int mt = 0, nt = 0, sp = 0, mtn = 0, ntn = 0, spn = 0;
    if (w < total_work) {
      decode(w, mt, nt, sp);
      int sv = 0, dv = 0;
      load_idx(mt, sv, dv);
      store_idx(0, sv, dv);
      if (w + gridDim.x < total_work) {
        decode(w + gridDim.x, mtn, ntn, spn);
        load_idx(mtn, sv, dv);
        store_idx(1, sv, dv);
      }
      group_bar();
      fetch_tile(mt, nt, 0);
    }
    for (; w < total_work; w += gridDim.x) {
      const int64_t wn = w + gridDim.x, wnn = wn + gridDim.x;
      const bool has_next = wn < total_work;
      if (tr_on) tr.begin();
      mbar_wait(tmem_full(acc), acc_ph);
      if (tr_on) { tr.end(TR_EPI_WAIT_ACC); tr.begin(); }
      tc_fence_after();
      const int64_t m0 = (int64_t)mt * BM;
#pragma unroll
      for (int q = 0; q < QN; ++q) {
        if constexpr (narrow) { if (q >= qn) break; }
        if constexpr (!kHeavyEpi) {
          // all of the chunk's TMEM loads are issued before the first wait (one round trip instead of EC / 16).  Only
          // for epilogues without a tile of prefetched operands in registers: with them (edge gate: 128 registers of
          // prefetch) the EC extra live registers spill — measured 138 -> 180 us on gemm_edge_gate.
          uint32_t v[EC / 16][16];
#pragma unroll
          for (int sub = 0; sub < EC / 16; ++sub)
            tmem_ld16_issue(tmem_base + ((uint32_t)(32 * ew) << 16) + (uint32_t)(acc * BN + cbase() + EC * q + 16 * sub), v[sub]);
#pragma unroll
          for (int sub = 0; sub < EC / 16; ++sub) tmem_ld16_fence(v[sub]);
          if (q == qn - 1) {                            // accumulator fully read: hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(tmem_empty(acc));
          }
#pragma unroll
          for (int sub = 0; sub < EC / 16; ++sub)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              stg[tg * TPR + ((4 * sub + j) ^ swz(tg))] =
                  make_float4(__uint_as_float(v[sub][4 * j]), __uint_as_float(v[sub][4 * j + 1]),
                              __uint_as_float(v[sub][4 * j + 2]), __uint_as_float(v[sub][4 * j + 3]));
        } else {
#pragma unroll
          for (int sub = 0; sub < EC / 16; ++sub) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(32 * ew) << 16) + (uint32_t)(acc * BN + cbase() + EC * q + 16 * sub), v);
            if (q == qn - 1 && sub == EC / 16 - 1) {    // accumulator fully read: hand it back to the MMA warp
              tc_fence_before();
              mbar_arrive(tmem_empty(acc));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              stg[tg * TPR + ((4 * sub + j) ^ swz(tg))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
        group_bar();
        // operands that depend on the column only (bias rows): once per chunk, not once per pass
        const typename Epi::ChunkC cconst = epi.chunk_const(col_of(nt, q), col_of(nt, q) < g.N);
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int p = 0; p < PPC; ++p) {
          const int r = p * RPP + trow;
          const float4 x = stg[r * TPR + (tcol ^ swz(r))];
          float a[4] = {x.x, x.y, x.z, x.w};
          const int64_t m = m0 + r;
          const int n = col_of(nt, q);
          const bool valid = (m < g.M) && (n < g.N);
          epi.apply_pre(m, n, a, deep[q][p], near[q][p], cconst, valid);
          if constexpr (Epi::kRowReduce) {
            // one scalar per output row: the TPR threads of the row are consecutive lanes (wide staging only)
            static_assert(!Epi::kRowReduce || EC == 64, "row reductions need the 64-column staging tile");
            float part = epi.row_partial(a, n);
#pragma unroll
            for (int o = TPR / 2; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (tcol == 0 && grp == 0 && m < g.M) epi.row_finish(m, part);
          }
          if constexpr (kStats) {
            if (valid && !(g.dbg & 1)) {
#pragma unroll
              for (int j = 0; j < 4; ++j) { s1[j] += a[j]; s2[j] = fmaf(a[j], a[j], s2[j]); }
            }
          }
        }
        if (kStats && !(g.dbg & 1)) {
          // column statistics: fp32 over this chunk's 128 rows (PPC per thread, then the 32 / TPR lanes that own
          // the same 4 columns), fp64 from there on
#pragma unroll
          for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int o = TPR; o < 32; o <<= 1) {
              s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
              s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
            }
          }
          if (lane < TPR) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              sstat[(ew * BN + cbase() + EC * q + c4 + j) * 2 + 0] += (double)s1[j];
              sstat[(ew * BN + cbase() + EC * q + c4 + j) * 2 + 1] += (double)s2[j];
            }
          }
        }
        group_bar();
      }
      if constexpr (kStats) {
        // flush this tile's column statistics (the n tile can change between work items)
        const int col = narrow ? 32 * grp + (tg & 31) : 64 * grp + (tg & 63);
        if (tg < (narrow ? 32 : 64)) {
          double a = 0.0, b = 0.0;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) { a += sstat[(qq * BN + col) * 2]; b += sstat[(qq * BN + col) * 2 + 1]; }
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) { sstat[(qq * BN + col) * 2] = 0.0; sstat[(qq * BN + col) * 2 + 1] = 0.0; }
          const int n = nt * BN + col;
          if (n < g.N) { atomicAdd(g.col_stats + n, a); atomicAdd(g.col_stats + g.N + n, b); }
        }
      }
      if (tr_on) { tr.end(TR_EPI_CHUNKS); tr.begin(); }
      // one batch: every operand of the next tile, then the node ids of the tile after it
      if (has_next) fetch_tile(mtn, ntn, cur ^ 1);
      int mtt = 0, ntt = 0, spt = 0;
      if (wnn < total_work) {
        int nsv = 0, ndv = 0;
        decode(wnn, mtt, ntt, spt);
        load_idx(mtt, nsv, ndv);
        store_idx(cur, nsv, ndv);                       // buffer of the tile just finished
      }
      group_bar();
      mt = mtn; nt = ntn; sp = spn;
      mtn = mtt; ntn = ntt; spn = spt;
      cur ^= 1;
      if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
      if (tr_on) tr.end(TR_EPI_FETCH);
    }
    if (tr_on) tr.role_end(TR_EPI_TOTAL);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor map: inner (contiguous) extent `inner`, outer extent `outer`, row pitch `ld` floats,
// box = 32 floats (128 B, SWIZZLE_128B) x box_rows
inline int make_map(CUtensorMap* map, const float* ptr, int64_t inner, int64_t outer, int64_t ld, int box_rows,
                    bool mn_major) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("gnnome_b200: cuTensorMapEncodeTiled not available"); return GG_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("gnnome_b200: cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")"); return GG_ERR_CUDA; }
  return GG_OK;
}

// K-major operands need K % 32 == 0 (whole 128-byte swizzle rows); MN-major ones take any K (zero fill).
// N is a multiple of 64: a last n tile with 64 valid columns reads a zero-filled B box and its second epilogue group
// never stores (d = 64: every layer GEMM is such a tile).
inline bool eligible(bool a_mn, bool b_mn, int64_t M, int N, int64_t K, int64_t lda, int64_t ldb, const void* A,
                     const void* B) {
  if (M <= 0 || K <= 0 || N % 64 != 0 || lda % 4 != 0 || ldb % 4 != 0) return false;
  if ((!a_mn || !b_mn) && K % BK != 0) return false;
  if (a_mn && M % 4 != 0) return false;
  return (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (reinterpret_cast<uintptr_t>(B) % 16 == 0);
}

// C[M,N] = sum_k A(m,k) B(k,n).  A_MN: A stored [K, M] (lda) else [M, K];  B_MN: B stored [K, N] (ldb) else [N, K].
// With an A transform, A2 is a second [M, K] operand streamed next to A (same lda).
template <bool A_MN, bool B_MN, bool kStats, bool kBiasGrad, class Epi, class ATx = NoATx, bool kWRes = false>
int launch(const char* tag, const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int N, int64_t K,
           int splits, double* col_stats, float* bias_grad, const Epi& epi, int num_sms, cudaStream_t st,
           const ATx& atx = ATx{}, const float* A2 = nullptr, int rev = 0) {
  CUtensorMap tmA, tmB, tmA2;
  int rc;
  if (A_MN) rc = make_map(&tmA, A, M, K, lda, 32, true); else rc = make_map(&tmA, A, K, M, lda, BM, false);
  if (rc) return rc;
  if (B_MN) rc = make_map(&tmB, B, N, K, ldb, 32, true); else rc = make_map(&tmB, B, K, N, ldb, BN, false);
  if (rc) return rc;
  CUtensorMap tmOut2;
  if constexpr (ATx::kActive) {
    if (K > 256) { set_error("gnnome_b200: A transform supports K <= 256"); return GG_ERR_UNSUPPORTED; }
    rc = make_map(&tmA2, A2, K, M, lda, BM, false);
    if (rc) return rc;
    if (atx.ld % 4 != 0 || reinterpret_cast<uintptr_t>(atx.g_t) % 16 != 0) {
      set_error("gnnome_b200: g_t must be 16-byte aligned with a leading dimension that is a multiple of 4");
      return GG_ERR_ARG;
    }
    rc = make_map(&tmOut2, atx.g_t, K, M, atx.ld, BM, false);       // same geometry as the t tiles it replaces in the stage
    if (rc) return rc;
  } else {
    tmA2 = tmA;
    tmOut2 = tmA;
  }
  Args g{};
  g.M = M; g.N = N; g.K = K;
  g.m_tiles = (int)((M + BM - 1) / BM);
  g.n_tiles = (N + BN - 1) / BN;                 // N < BN: the B box is zero-filled past N, columns >= N are never valid
  if (splits < 1) splits = 1;
  int64_t chunk = (K + splits - 1) / splits;
  chunk = ((chunk + BK - 1) / BK) * BK;
  g.k_chunk = chunk;
  g.splits = (int)((K + chunk - 1) / chunk);
  g.col_stats = col_stats;
  g.bias_grad = bias_grad;
  g.dbg = tc_dbg_ref();
  g.rev = rev;
  const int64_t work = (int64_t)g.m_tiles * g.n_tiles * g.splits;
  const int grid = (int)(work < num_sms ? work : num_sms);
  if (kWRes && (g.n_tiles != 1 || K > 4 * BK || g.splits != 1)) {
    set_error("gnnome_b200: W-resident GEMM needs one n tile, K <= 128, no split-K");
    return GG_ERR_UNSUPPORTED;
  }
  constexpr int kSmem = smem_bytes<kStats, Epi, ATx, kWRes>();
  static_assert(kSmem <= 227 * 1024, "shared-memory layout exceeds 227 KB");
  // N <= 64 (every layer GEMM at d = 64): the variant whose two epilogue groups share the 64 valid columns.  A separate
  // instantiation, so the N = 128 kernels keep their code (and register allocation) to the instruction.
  constexpr bool kCanNarrow = Cfg<ATx, Epi, kWRes>::kEC <= 32;
  const bool narrow = kCanNarrow && g.n_tiles == 1 && N <= 64;
  auto launch_variant = [&](auto kern, bool& attr_set) -> int {
    if (!attr_set) {
      GG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
      attr_set = true;
    }
    GG_KERNEL_BEGIN(tag, st);
    kern<<<grid, threads<ATx>(), kSmem, st>>>(tmA, tmB, tmA2, tmOut2, g, epi, atx);
    GG_KERNEL_END(tag, st);
    return GG_OK;
  };
  static bool attr_set_dev[2][kMaxDevices] = {};   // one static per template instantiation: [narrow][device]
  if constexpr (kCanNarrow) {
    if (narrow)
      return launch_variant(gemm_tc_kernel<A_MN, B_MN, kStats, kBiasGrad, Epi, ATx, kWRes, true>, attr_set_dev[1][current_device()]);
  }
  return launch_variant(gemm_tc_kernel<A_MN, B_MN, kStats, kBiasGrad, Epi, ATx, kWRes, false>, attr_set_dev[0][current_device()]);
}

}  // namespace tc
}  // namespace gg
