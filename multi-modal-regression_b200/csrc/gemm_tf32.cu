// (a) Grouped TF32 GEMM for the bin-delta heads on sm_100a: tcgen05.mma with TMEM accumulators,
// operands streamed by TMA straight from the fp32 master weights (no bf16 repack pass: at the
// reference's batch sizes the head is bound by streaming the weights once).
//
//   D_g[m, n] = sum_k A_g(m, k) * B_g(n, k)          g = 0..G-1
//
// Either operand may be K-major (row-major [rows, K]) or MN-major (row-major [K, rows]); both are
// legal tcgen05 TF32 shared-memory layouts, so one kernel serves fprop (W x^T), dgrad (W^T dy) and
// wgrad (dy a^T) of nn.Linear (binDeltaModels.py:71-75, 87-91) without any transpose pass:
//   rows of D (m)  -> the 128 TMEM lanes  (fprop / dgrad: the batch, only the rows that exist are fetched)
//   cols of D (n)  -> TMEM columns, BN <= 256 per tile (fprop / dgrad: the streamed weight rows)
//
// Warp roles (512 threads, persistent over tiles): warp 0 = TMA producer, warp 1 = MMA issuer
// (one elected lane), warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld -> global),
// warps 8-15 = operand splitters (precise mode only).  mbarrier pipelines: smem full/empty
// (TMA <-> MMA), split-ready (splitters -> MMA), TMEM full/empty (MMA <-> epilogue, two accumulator
// stages so the epilogue of tile i overlaps the mainloop of tile i+1).
//
// Precision.  precise = 0: one TF32 MMA per k-step (operands truncated to 10 mantissa bits by the
// tensor core, ~1e-3 relative).  precise = 1 ("3xTF32"): every fp32 operand tile is split in shared
// memory into hi = tf32(x) (the raw fp32 itself: the tensor core ignores the low 13 mantissa bits)
// and lo = tf32(x - hi), and D += Ahi*Bhi + Ahi*Blo + Alo*Bhi; the dropped
// terms are ~2^-20 relative, i.e. fp32-class results (needed for parity: at 1e-3 a few ReLU masks
// flip and gradients drift by percents).  The weights still stream from HBM exactly once.
// Stacked form (K-major A of <= 64 rows, i.e. every fprop / dgrad at the reference's batch sizes):
// the A_lo rows are written directly under the A_hi rows of the same operand tile, so ONE MMA
// computes Ahi*B (TMEM lanes 0..a_rows-1) and Alo*B (lanes a_rows..2*a_rows-1); a k-step is
// [Ahi;Alo]*Bhi + [Ahi;Alo]*Blo = 2 MMAs instead of 3, and the epilogue adds the lo lanes to the hi
// lanes.  The precise mode is bound by shared-memory traffic (every MMA re-reads its operands), so
// one MMA less per k-step is ~30 % of the kernel time.
//
// Accumulator chains.  The tensor core adds each K=8 partial sum into the fp32 accumulator with
// truncation, so one long chain drifts by ~(#k-steps) * ulp(acc)/2 (measured: 1.6e-5 of the result
// scale at K=2048).  When the tile is narrow the spare TMEM columns hold `chains` independent
// accumulators per tile: k-steps go round-robin over `chains_hi` of them, the small hi*lo cross terms
// go to the remaining ones (where ulp is ~2^-11 smaller), and the epilogue adds the chains in
// registers with round-to-nearest.  BN = 32 -> 8 chains, 64 -> 4, 128 -> 2, 256 -> 1.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kBM = 128;            // tile rows = TMEM lanes
constexpr int kBK = 32;             // fp32 per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 8;           // K of one tcgen05.mma.kind::tf32
constexpr int kGemmThreads = 512;
constexpr int kSplitThreads = 256;   // warps 8-15 (precise mode)
constexpr int kABytes = kBM * kBK * 4;     // 16 KB per stage
constexpr int kMaxStages = 10;
constexpr int kEpiPitch = 36;        // floats per staged row (32 + 4: conflict-free float4 rows)
constexpr int kEpiBytes = 4 * 32 * kEpiPitch * 4;   // 4 epilogue warps x 32 rows
constexpr int kStackEpiBytes = 64 * kEpiPitch * 4;  // stacked mode: the lo rows' partial sums of one 32-column chunk

struct GemmParams {
  int M, N, K, G;
  int a_mn, b_mn;                   // 1 = MN-major operand
  int BN;                           // tile columns: multiple of 32, <= 256
  int m_tiles, n_tiles, splits, kb_per_split, kb_total, stages;
  int a_g, b_g;                     // 1: operand has a group axis, 0: shared by all groups
  float* C;
  long long ldc, c_gstride, c_sstride;
  int c_nm;                         // 0: C[m*ldc + n], 1: C[n*ldc + m]
  int precise;                      // 1: 3xTF32 split-operand accumulation
  int chains, chains_hi;            // TMEM accumulator chains per tile (see below)
  int cstride;                      // TMEM columns between accumulator chains (BN rounded up to 32)
  int k_rotate;                     // 1: stagger the K walk of the tiles
  int a_rows;                       // A rows actually fetched per stage (small-M problems)
  int a_bytes;                      // smem reserved for the A tile of one stage (1 KB multiple)
  int acc_stages;                   // TMEM accumulator stages (2 only when a CTA runs >1 tile)
  int stack;                        // precise mode, batch-side A with <= 64 rows: A_lo rows stacked under A_hi (2 MMAs per k-step)
  int n_epi;                        // epilogue warps per TMEM lane quarter: warps 4-7, plus the warps after the splitters
  int stage_out;                    // 1: epilogue transposes through smem so stores are whole 128-byte row segments
};

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t it = 0;; ++it) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (it > (1u << 27)) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// One lane of a converged warp (always the same one for a full mask): the warp keeps running
// uniform code, so descriptors and addresses stay in uniform registers for UTCHMMA / UTMALDG.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// `issue` is the elected-lane flag: the instructions are predicated, not branched around, so the
// warp stays converged and the operands stay in uniform registers.
__device__ __forceinline__ void umma_commit(uint32_t bar, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(bar), "r"(issue) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate, uint32_t issue) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %6, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z), "r"(issue)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
      "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 columns, or the 16-column tail of a tile whose width is an odd multiple of 16
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[32], bool full) {
  if (full) {
    tmem_ld32(taddr, v);
  } else {
    tmem_ld16(taddr, v);
#pragma unroll
    for (int i = 16; i < 32; ++i) v[i] = 0u;
  }
}

// UMMA shared-memory descriptor, 128-byte swizzle (cute::UMMA::SmemDescriptor, version 1).
//   K-major : rows are 128 B apart inside an 8-row group, groups SBO = 1024 B apart; the K
//             position inside the 128 B row is selected by advancing the start address.
//   MN-major: 32-bit operands only exist in the SWIZZLE_128B_BASE32B layout (type 1; TMA mode
//             SWIZZLE_128B_ATOM_32B): rows of 128 B hold 32 consecutive MN elements, 4 K-rows form
//             a swizzle atom (SBO = 512 B between atoms, one K=8 MMA spans two), the next 32 MN
//             elements live LBO bytes further.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;     // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(layout_type) << 61;
  return d;
}

// One kernel per mode (MODE 0: tf32, 1: 3xTF32, 2: 3xTF32 with stacked A): each carries only its own
// roles and loops.  The producer / issuer / splitter / epilogue warps run different code at the same
// time, and both the uniform mode tests inside the k-block loop and the sheer code size of a single
// generic kernel showed up in the timings (tf32 fc1 45 us generic vs 37 us specialised).
template <int MODE>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmParams P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte aligned operand ring (swizzle-128B atoms are 1 KB), barriers after it
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_bytes = static_cast<uint32_t>(P.BN) * kBK * 4;
  const uint32_t a_bytes = static_cast<uint32_t>(P.a_bytes);
  const uint32_t tile_bytes = a_bytes + b_bytes;                 // operand footprint of one stage
  const uint32_t a_tx = static_cast<uint32_t>(P.a_rows) * kBK * 4;   // A bytes TMA really delivers
  // lo copies (precise mode): of the whole [A | B] footprint right after it, or — stacked mode — the
  // A_lo rows directly under the A_hi rows (inside a_bytes) and B_lo after B
  constexpr bool kPrecise = MODE != 0, kStack = MODE == 2;
  const uint32_t stage_bytes = kStack ? a_bytes + 2u * b_bytes : kPrecise ? 2u * tile_bytes : tile_bytes;
  const uint32_t a_lo_off = kStack ? a_tx : tile_bytes;      // from the A tile
  const uint32_t b_lo_off = kStack ? b_bytes : tile_bytes;   // from the B tile
  const uint32_t bar_base = base + P.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto split_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (3 * kMaxStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (3 * kMaxStages + 2 + s); };
  const uint32_t epi_base = bar_base + 8u * (3 * kMaxStages + 4);      // 16-byte aligned
  __shared__ uint32_t s_tmem_base;

  // warp index through a shuffle: provably warp-uniform, so the role branches are uniform branches
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int need_cols = P.acc_stages * P.chains * P.cstride;
  const int tmem_cols = (need_cols <= 32) ? 32 : (need_cols <= 64) ? 64 : (need_cols <= 128) ? 128
                        : (need_cols <= 256) ? 256 : 512;
  const int chains_x = P.chains - P.chains_hi;       // chains reserved for the cross terms
  // tf32 mode: the splitter warps have nothing to split and join the epilogue (three warps per TMEM
  // lane quarter take the 32-column chunks round-robin; the epilogue is a latency chain per chunk)
  const int n_epi = P.n_epi;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(split_bar(s), kSplitThreads / 32);
    }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4 * n_epi); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&s_tmem_base)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  const int tiles_per_group = P.m_tiles * P.n_tiles * P.splits;
  const int total_tiles = P.G * tiles_per_group;

  if (warp == 0) {
    // ===== TMA producer (whole warp runs the loop, one elected lane issues) =====
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int g = t / tiles_per_group;
      int r = t - g * tiles_per_group;
      const int sp = r % P.splits; r /= P.splits;
      const int nt = r % P.n_tiles;
      const int mt = r / P.n_tiles;
      const int kb0 = sp * P.kb_per_split;
      const int kb1 = min(P.kb_total, kb0 + P.kb_per_split);
      const int m0 = mt * kBM, n0 = nt * P.BN;
      // Each tile walks its K range from a different starting block (the sum is order-free): with
      // a power-of-two row pitch (2048 floats) all CTAs would otherwise request the same column
      // block of 148 x BN different rows at the same time — addresses 8 KB apart, which pile onto a
      // few DRAM channels.
      const int nkb = kb1 - kb0;
      const int rot = P.k_rotate ? static_cast<int>((static_cast<unsigned>(t) * 13u) % static_cast<unsigned>(nkb)) : 0;
      int kb = kb0 + rot;
      for (int i = 0; i < nkb; ++i, ++kb) {
        if (kb == kb1) kb = kb0;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = base + stage * stage_bytes;
        const uint32_t sb = sa + a_bytes;
        const int k0 = kb * kBK;
        if (elect_one()) {
          mbar_expect_tx(full_bar(stage), a_tx + b_bytes);
          if (!P.a_mn) {
            tma_load_3d(sa, &tmA, full_bar(stage), k0, m0, g * P.a_g);    // box = a_rows x 32
          } else {
            for (int j = 0; j < P.a_rows / 32; ++j)
              tma_load_3d(sa + j * 4096, &tmA, full_bar(stage), m0 + 32 * j, k0, g * P.a_g);
          }
          if (!P.b_mn) {
            tma_load_3d(sb, &tmB, full_bar(stage), k0, n0, g * P.b_g);
          } else {
            for (int j = 0; j < P.BN / 32; ++j)
              tma_load_3d(sb + j * 4096, &tmB, full_bar(stage), n0 + 32 * j, k0, g * P.b_g);
          }
        }
        __syncwarp();
        if (++stage == P.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp runs the loop, one elected lane issues) =====
    // cute::UMMA::InstrDescriptor: c_format F32 (bit 4), a/b_format TF32 = 2 (bits 7, 10),
    // a/b major (bits 15, 16), N >> 3 (bits 17..22), M >> 4 (bits 24..28)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) |
                           (static_cast<uint32_t>(P.a_mn) << 15) |
                           (static_cast<uint32_t>(P.b_mn) << 16) |
                           (static_cast<uint32_t>(P.BN >> 3) << 17) |
                           (static_cast<uint32_t>(kBM >> 4) << 24);
    // Everything that does not change per k-step is hoisted: up to 12 MMAs are issued per k-block and
    // the scalar instruction stream around them paces the precise mode (runtime modulos and 64-bit
    // descriptor assembly per MMA cost ~2500 cycles per k-block; a single divergent issuing lane
    // ~145 cycles per MMA in R2UR moves and elect loops — hence the converged warp).
    const uint64_t a_desc0 = umma_desc(0, P.a_mn ? 4096u : 16u, P.a_mn ? 512u : 1024u, P.a_mn ? 1u : 2u);
    const uint64_t b_desc0 = umma_desc(0, P.b_mn ? 4096u : 16u, P.b_mn ? 512u : 1024u, P.b_mn ? 1u : 2u);
    const uint32_t a_kstep16 = (P.a_mn ? 1024u : kUmmaK * 4u) >> 4;   // descriptor address units
    const uint32_t b_kstep16 = (P.b_mn ? 1024u : kUmmaK * 4u) >> 4;
    const uint32_t lo16 = tile_bytes >> 4;               // hi -> lo copy of the same operand
    const uint32_t b_lo16 = b_lo_off >> 4;
    const uint32_t cstride = static_cast<uint32_t>(P.cstride);
    const uint32_t hi_span = static_cast<uint32_t>(P.chains_hi) * cstride;
    const uint32_t x_span = static_cast<uint32_t>(chains_x) * cstride;
    const uint32_t leader = elect_one() ? 1u : 0u;
    {
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int r = t % tiles_per_group;
        const int sp = r % P.splits;
        const int kb0 = sp * P.kb_per_split;
        const int kb1 = min(P.kb_total, kb0 + P.kb_per_split);
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_t = tmem_base + static_cast<uint32_t>(as * P.chains) * cstride;
        // chains are visited round-robin; a chain accumulates from its second visit on
        uint32_t hi_col = 0, x_col = 0;
        int hi_fresh = P.chains_hi, x_fresh = chains_x;   // chains not yet written in this tile
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(kPrecise ? split_bar(stage) : full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = base + stage * stage_bytes;
          uint64_t ad = a_desc0 | static_cast<uint64_t>((sa & 0x3FFFFu) >> 4);
          uint64_t bd = b_desc0 | static_cast<uint64_t>(((sa + a_bytes) & 0x3FFFFu) >> 4);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            const uint32_t d_hi = tmem_t + hi_col;
            if (kStack) {
              // A = [A_hi rows ; A_lo rows]: lanes < a_rows get A_hi*B, lanes a_rows.. get A_lo*B
              if (x_span != 0) {
                const uint32_t d_x = tmem_t + hi_span + x_col;
                umma_tf32(d_hi, ad, bd, idesc, hi_fresh <= 0, leader);
                umma_tf32(d_x, ad, bd + b_lo16, idesc, x_fresh <= 0, leader);
                --x_fresh;
                x_col += cstride;
                if (x_col == x_span) x_col = 0;
              } else {
                umma_tf32(d_hi, ad, bd + b_lo16, idesc, hi_fresh <= 0, leader);
                umma_tf32(d_hi, ad, bd, idesc, 1u, leader);
              }
            } else if (kPrecise) {
              if (x_span != 0) {
                const uint32_t d_x = tmem_t + hi_span + x_col;
                umma_tf32(d_x, ad + lo16, bd, idesc, x_fresh <= 0, leader);
                umma_tf32(d_x, ad, bd + lo16, idesc, 1u, leader);
                umma_tf32(d_hi, ad, bd, idesc, hi_fresh <= 0, leader);
                --x_fresh;
                x_col += cstride;
                if (x_col == x_span) x_col = 0;
              } else {                                     // a single chain takes everything
                umma_tf32(d_hi, ad + lo16, bd, idesc, hi_fresh <= 0, leader);
                umma_tf32(d_hi, ad, bd + lo16, idesc, 1u, leader);
                umma_tf32(d_hi, ad, bd, idesc, 1u, leader);
              }
            } else {
              umma_tf32(d_hi, ad, bd, idesc, hi_fresh <= 0, leader);
            }
            --hi_fresh;
            hi_col += cstride;
            if (hi_col == hi_span) hi_col = 0;
            ad += a_kstep16;
            bd += b_kstep16;
          }
          umma_commit(empty_bar(stage), leader);                 // frees the smem slot when the MMAs retire
          if (kb == kb1 - 1) umma_commit(tfull_bar(as), leader); // accumulator complete -> epilogue
          if (++stage == P.stages) { stage = 0; phase ^= 1u; }
        }
        if (++as == P.acc_stages) { as = 0; aphase ^= 1u; }
      }
    }
  } else if (kPrecise && warp >= 8) {
    // ===== operand splitters (precise mode): x -> hi = tf32(x) in place, lo = tf32(x - hi) =====
    {
      const int tid = threadIdx.x - 8 * 32;
      int stage = 0;
      uint32_t phase = 0;
      // A rows beyond a_rows are never fetched (their D rows are discarded), so only the fetched
      // part of A and the B tile are split; K-major A: the first a_rows*128 bytes, MN-major A: whole
      const int a_vec = static_cast<int>((P.a_mn ? a_bytes : a_tx) / 16);
      const int b_vec = static_cast<int>(b_bytes / 16);
      const int nvec = a_vec + b_vec;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int sp = (t % tiles_per_group) % P.splits;
        const int kb0 = sp * P.kb_per_split;
        const int kb1 = min(P.kb_total, kb0 + P.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          const uint32_t sa = base + stage * stage_bytes;
          // four vectors per thread in flight: the loop is latency-bound (shared-memory round trip
          // per iteration), not bandwidth-bound
          for (int i0 = tid; i0 < nvec; i0 += 4 * kSplitThreads) {
            uint32_t x[4][4];
            uint32_t off[4], lo_off[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int ii = i0 + u * kSplitThreads;
              const int i = ii < a_vec ? ii : ii - a_vec + static_cast<int>(a_bytes / 16);
              off[u] = 16u * static_cast<uint32_t>(i);
              lo_off[u] = off[u] + (ii < a_vec ? a_lo_off : b_lo_off);
              if (ii < nvec)
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(x[u][0]), "=r"(x[u][1]), "=r"(x[u][2]), "=r"(x[u][3]) : "r"(sa + off[u]));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (i0 + u * kSplitThreads < nvec) {
                // the tensor core ignores the low 13 mantissa bits of a TF32 operand, so the raw fp32
                // left in place IS the hi part; only lo = tf32(x - hi) is written
                uint32_t l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint32_t h = x[u][j] & 0xFFFFE000u;
                  l[j] = __float_as_uint(__uint_as_float(x[u][j]) - __uint_as_float(h)) & 0xFFFFE000u;
                }
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sa + lo_off[u]),
                             "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
              }
            }
          }
          // generic-proxy writes -> visible to the tensor core's async-proxy reads
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(split_bar(stage));
          if (++stage == P.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if ((warp >= 4 && warp < 8) || (!kPrecise && warp >= 8 && warp < 8 + 4 * (P.n_epi - 1))) {
    // ===== epilogue: TMEM -> registers -> global =====
    const int q = warp & 3;                            // TMEM lane quarter of this warp
    const int e = warp < 8 ? 0 : 1 + ((warp - 8) >> 2);   // which of the quarter's n_epi warps
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int g = t / tiles_per_group;
      int r = t - g * tiles_per_group;
      const int sp = r % P.splits; r /= P.splits;
      const int nt = r % P.n_tiles;
      const int mt = r / P.n_tiles;
      const int m = mt * kBM + q * 32 + lane;
      const int n0 = nt * P.BN;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      float* Cg = P.C + sp * P.c_sstride + g * P.c_gstride;
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                            static_cast<uint32_t>(as * P.chains * P.cstride);
      // chains that received at least one k-step of this tile
      const int nsteps = (min(P.kb_total, sp * P.kb_per_split + P.kb_per_split) - sp * P.kb_per_split) *
                         (kBK / kUmmaK);
      const int used_hi = min(P.chains_hi, nsteps);
      const int used_x = (kPrecise && chains_x > 0) ? min(chains_x, nsteps) : 0;
      for (int c = e; c * 32 < P.BN; c += n_epi) {
        if (n0 + c * 32 >= P.N) break;                 // warp-uniform
        const bool full = c * 32 + 32 <= P.BN;         // else: 16-column tail (BN = odd multiple of 16)
        uint32_t v[32];
        tmem_ld_cols(trow + c * 32, v, full);
        for (int chn = 1; chn < used_hi + used_x; ++chn) {
          const int cc = chn < used_hi ? chn : P.chains_hi + (chn - used_hi);
          uint32_t w[32];
          tmem_ld_cols(trow + cc * P.cstride + c * 32, w, full);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(w[i]));
        }
        const int nb = n0 + c * 32;
        const int wcols = full ? 32 : 16;
        if (kStack) {
          // rows a_rows .. 2*a_rows-1 hold A_lo * (B_hi + B_lo): hand them to the rows above through
          // shared memory (the two row sets can sit in different warps)
          const int rr = q * 32 + lane;
          if (rr >= P.a_rows && rr < 2 * P.a_rows) {
            const uint32_t dst = epi_base + static_cast<uint32_t>((rr - P.a_rows) * kEpiPitch) * 4;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 16 * j),
                           "r"(v[4 * j]), "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3]) : "memory");
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (rr < P.a_rows) {
            const uint32_t src = epi_base + static_cast<uint32_t>(rr * kEpiPitch) * 4;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint32_t o0, o1, o2, o3;
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(o0), "=r"(o1), "=r"(o2), "=r"(o3) : "r"(src + 16 * j));
              v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + __uint_as_float(o0));
              v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + __uint_as_float(o1));
              v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + __uint_as_float(o2));
              v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + __uint_as_float(o3));
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");   // the buffer is rewritten by the next chunk
        }
        if (P.stage_out) {
          // Big outputs (wgrad: 197 MB): lane r holds 32 columns of ROW r, so a direct store writes
          // 32 scattered 16-byte pieces per instruction and every 32-byte sector twice.  Transposed
          // through shared memory each instruction writes four whole 128-byte row segments.
          const uint32_t tile = epi_base + static_cast<uint32_t>(e * 4 + q) * (32 * kEpiPitch * 4);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(tile + (lane * kEpiPitch + 4 * j) * 4),
                         "r"(v[4 * j]), "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3]) : "memory");
          __syncwarp();
          const int sub = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + sub;
            const int mr = mt * kBM + q * 32 + rr;
            uint32_t o0, o1, o2, o3;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(o0), "=r"(o1), "=r"(o2), "=r"(o3) : "r"(tile + (rr * kEpiPitch + c4) * 4));
            if (mr < P.M && c4 < wcols && nb + c4 < P.N)          // N % 4 == 0 on this path
              *reinterpret_cast<float4*>(Cg + static_cast<long long>(mr) * P.ldc + nb + c4) =
                  make_float4(__uint_as_float(o0), __uint_as_float(o1), __uint_as_float(o2), __uint_as_float(o3));
          }
          __syncwarp();
        } else if (!P.c_nm) {
          if (m < P.M) {
            float* dst = Cg + static_cast<long long>(m) * P.ldc + nb;
            const bool vec = (nb + wcols <= P.N) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
            if (vec) {
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                if (i < wcols)
                  *reinterpret_cast<float4*>(dst + i) =
                      make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                                  __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < wcols && nb + i < P.N) dst[i] = __uint_as_float(v[i]);
            }
          }
        } else {
          if (m < P.M) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < wcols && nb + i < P.N)
                Cg[static_cast<long long>(nb + i) * P.ldc + m] = __uint_as_float(v[i]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      if (++as == P.acc_stages) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(tmem_cols) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// rank-3 fp32 map over [dim2][dim1][dim0] (dim0 contiguous), 128-byte swizzle (16 B atoms for
// K-major tiles, 32 B atoms for MN-major tiles), zero OOB fill
// Encoded maps are cached: a training step issues the same 8 GEMMs every iteration (same pointers
// when the caching allocator hands the same blocks back, always the same weights), and one
// cuTensorMapEncodeTiled costs ~2 us of host time on a step that is host-bound.
struct MapKey {
  const float* ptr;
  uint64_t dim0, dim1, dim2, s1, s2;
  uint32_t box0, box1, mn;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && dim0 == o.dim0 && dim1 == o.dim1 && dim2 == o.dim2 && s1 == o.s1 &&
           s2 == o.s2 && box0 == o.box0 && box1 == o.box1 && mn == o.mn;
  }
};
constexpr int kMapCache = 64;
thread_local MapKey t_keys[kMapCache];
thread_local CUtensorMap t_maps[kMapCache];
thread_local bool t_valid[kMapCache];

int make_map(CUtensorMap* map, const float* ptr, uint64_t dim0, uint64_t dim1, uint64_t dim2,
             uint64_t stride1_elems, uint64_t stride2_elems, uint32_t box0, uint32_t box1,
             bool mn_major) {
  const MapKey key = {ptr, dim0, dim1, dim2, stride1_elems, stride2_elems, box0, box1,
                      mn_major ? 1u : 0u};
  uint64_t h = reinterpret_cast<uint64_t>(ptr) >> 4;
  h = (h ^ (dim0 * 0x9E3779B97F4A7C15ull) ^ (dim1 * 0xC2B2AE3D27D4EB4Full) ^ (dim2 << 7) ^
       (stride1_elems << 17) ^ (stride2_elems << 29) ^ ((uint64_t)box1 << 41) ^ key.mn) *
      0xD6E8FEB86659FD93ull;
  const int slot = (int)((h >> 32) % kMapCache);
  if (t_valid[slot] && t_keys[slot] == key) {
    *map = t_maps[slot];
    return BDP_OK;
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) { bdp_set_error("gemm_tf32: cuTensorMapEncodeTiled entry point unavailable"); return BDP_ERR_CUDA; }
  cuuint64_t dims[3] = {dim0, dim1, dim2};
  cuuint64_t strides[2] = {stride1_elems * 4, stride2_elems * 4};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    bdp_set_error("gemm_tf32: cuTensorMapEncodeTiled failed (%d) dims=[%llu,%llu,%llu] strides=[%llu,%llu]B "
                  "box=[%u,%u]", (int)r, (unsigned long long)dim0, (unsigned long long)dim1,
                  (unsigned long long)dim2, (unsigned long long)strides[0],
                  (unsigned long long)strides[1], box0, box1);
    return BDP_ERR_CUDA;
  }
  t_keys[slot] = key;
  t_maps[slot] = *map;
  t_valid[slot] = true;
  return BDP_OK;
}

}  // namespace

extern "C" int bdp_gemm_tf32(const float* A, int a_major, int64_t a_ld, int64_t a_gstride,
                             const float* B, int b_major, int64_t b_ld, int64_t b_gstride, float* C,
                             int c_layout, int64_t ldc, int64_t c_gstride, int64_t M, int64_t N,
                             int64_t K, int G, int splits, int64_t c_sstride, int precise,
                             void* stream) {
  BDP_REQUIRE(A && B && C, "gemm_tf32: NULL operand");
  BDP_REQUIRE(M > 0 && N > 0 && K > 0 && G > 0, "gemm_tf32: empty problem M=%lld N=%lld K=%lld G=%d",
              (long long)M, (long long)N, (long long)K, G);
  BDP_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "gemm_tf32: dimension too large");
  BDP_REQUIRE((a_major == 0 || a_major == 1) && (b_major == 0 || b_major == 1) &&
                  (c_layout == 0 || c_layout == 1), "gemm_tf32: bad major/layout flag");
  BDP_REQUIRE(a_ld % 4 == 0 && b_ld % 4 == 0 && a_gstride % 4 == 0 && b_gstride % 4 == 0,
              "gemm_tf32: operand strides must be multiples of 4 floats (TMA: 16 bytes)");
  BDP_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
              "gemm_tf32: operands must be 16-byte aligned");
  BDP_REQUIRE(splits >= 1, "gemm_tf32: splits must be >= 1");

  GemmParams P = {};
  P.M = (int)M; P.N = (int)N; P.K = (int)K; P.G = G;
  P.a_mn = a_major; P.b_mn = b_major;
  P.m_tiles = (int)((M + kBM - 1) / kBM);
  P.kb_total = (int)((K + kBK - 1) / kBK);
  if (splits > P.kb_total) splits = P.kb_total;
  P.kb_per_split = (P.kb_total + splits - 1) / splits;
  P.splits = (P.kb_total + P.kb_per_split - 1) / P.kb_per_split;   // no empty split
  // Tile width: the B operand carries the big matrix (the weights) in fprop / dgrad, and every CTA
  // streams its BN rows of it exactly once, so the best width is the one that keeps the most SMs
  // streaming for the fewest waves: minimise ceil(tiles / SMs) * (BN + 16) (ties: wider tile).
  const int sms = bdp_num_sms();
  const int bn_step = b_major ? 32 : 16;              // MN-major B arrives in 32-wide TMA boxes
  // Precise mode bounds the number of k-steps one TMEM accumulator chain absorbs (each MMA adds its
  // partial sum with truncation: ~64 steps keep the drift near 3e-6 of the result scale), which caps
  // the tile width: `need` chains of roundup32(BN) columns must fit the 512 TMEM columns.
  const int steps = P.kb_per_split * (kBK / kUmmaK);
  int need = 1;
  if (precise && 3 * steps > 64) {
    int need_hi = (steps + 63) / 64;
    if (need_hi > 6) need_hi = 6;
    need = need_hi + 1;                               // + one chain for the hi*lo cross terms
  }
  // (short K — the wgrads, K = batch: all 3 * steps MMAs fit one chain's drift budget, and one chain
  // of 256 columns leaves room for two accumulator stages, so epilogue and mainloop overlap)
  int bn = 0;
  long long best_cost = 0;
  for (int cand = 256; cand >= bn_step; cand -= bn_step) {
    if (cand > (int)((N + bn_step - 1) / bn_step * bn_step)) continue;
    if (((cand + 31) / 32 * 32) * need > 512 && cand > bn_step) continue;
    const long long tiles = (long long)G * P.m_tiles * ((N + cand - 1) / cand) * P.splits;
    const long long cost = ((tiles + sms - 1) / sms) * (cand + 16);
    if (bn == 0 || cost < best_cost) { bn = cand; best_cost = cost; }
  }
  { const char* e = getenv("BDP_GEMM_BN"); if (e && atoi(e) >= bn_step && atoi(e) <= 256) bn = atoi(e) / bn_step * bn_step; }   // timing experiments
  P.BN = bn;
  P.n_tiles = (int)((N + bn - 1) / bn);
  P.a_g = (a_gstride != 0 || G == 1) ? 1 : 0;
  P.b_g = (b_gstride != 0 || G == 1) ? 1 : 0;
  P.C = C; P.ldc = ldc; P.c_gstride = c_gstride; P.c_sstride = c_sstride; P.c_nm = c_layout;
  P.stage_out = (c_layout == 0 && M >= 64 && N % 4 == 0 && ldc % 4 == 0 && c_gstride % 4 == 0 &&
                 c_sstride % 4 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0) ? 1 : 0;
  P.precise = precise ? 1 : 0;
  {
    static const int no_rotate = [] { const char* e = getenv("BDP_GEMM_NO_ROTATE"); return (e && e[0] == '1') ? 1 : 0; }();
    P.k_rotate = no_rotate ? 0 : 1;
  }
  // small-M problems (M = batch) fetch only the rows that exist
  if (P.m_tiles == 1) P.a_rows = a_major ? (int)((M + 31) / 32 * 32) : (int)((M + 7) / 8 * 8);
  else P.a_rows = kBM;
  // 3xTF32 with a batch-side A of <= 64 rows: the A_lo rows ride in the same MMA as the A_hi rows
  // (TMEM lanes a_rows..2*a_rows-1), so a k-step is 2 MMAs instead of 3 — the precise mode is bound
  // by the tensor core's shared-memory operand reads (every MMA re-reads 128 A rows + BN B rows)
  P.stack = (precise && !a_major && P.m_tiles == 1 && P.a_rows <= 64) ? 1 : 0;
  { const char* e = getenv("BDP_GEMM_NO_STACK"); if (e && e[0] == '1') P.stack = 0; }
  P.a_bytes = a_major ? (P.a_rows / 32) * 4096
                      : ((P.stack ? 2 : 1) * P.a_rows * kBK * 4 + 1023) / 1024 * 1024;
  if (P.stack) P.stage_out = 0;
  // Warps 8-15: operand splitters in precise mode, otherwise free.  The epilogue is a latency chain
  // per 32-column chunk (tcgen05.ld -> shared-memory transpose -> stores), so in tf32 mode four of the
  // free warps take every second chunk of their TMEM lane quarter (fc1 wgrad 52 -> 37 us on the same
  // box; a third set did not add anything).  In precise mode the second staging buffer does not fit
  // next to two 96 KB operand stages, so the epilogue stays with warps 4-7.
  P.n_epi = precise ? 1 : 2;
  { const char* e = getenv("BDP_GEMM_EPI_WARPS"); if (e && e[0] == '1') P.n_epi = 1; }   // experiments
  const long long total = (long long)G * P.m_tiles * P.n_tiles * P.splits;
  long long grid = sms;
  if (grid > total) grid = total;
  P.cstride = (bn + 31) / 32 * 32;
  // two accumulator stages (epilogue of tile i overlaps the mainloop of tile i+1) when a CTA runs
  // several tiles and the chains still fit
  P.acc_stages = (total > grid && 2 * need * P.cstride <= 512) ? 2 : 1;
  P.chains = 512 / (P.acc_stages * P.cstride);
  if (P.chains < 1) P.chains = 1;
  if (P.chains > 8) P.chains = 8;
  // tf32 mode rounds the operands to 10 mantissa bits (~1e-3): the accumulator drift of a single
  // chain (1.6e-5 of the result scale at K = 2048) is irrelevant there, and one chain means one
  // tcgen05.ld per chunk in the epilogue instead of up to five
  { const char* e = getenv("BDP_GEMM_TF32_CHAINS"); if (!precise && !(e && e[0] == '1')) P.chains = 1; }
  P.chains_hi = (precise && P.chains >= 2) ? P.chains - (P.chains >= 4 ? P.chains / 4 : 1) : P.chains;

  const size_t stage_bytes = P.stack ? (size_t)P.a_bytes + 2 * (size_t)bn * kBK * 4
                                     : ((size_t)P.a_bytes + (size_t)bn * kBK * 4) * (precise ? 2 : 1);
  // The MMA always reads 128 A rows from shared memory (rows >= a_rows produce discarded D rows), so
  // a shrunk A reservation is followed by `slack` bytes that keep those reads inside the allocation.
  const size_t slack = (size_t)(kABytes - P.a_bytes);
  const size_t fixed = 1024 + 8 * (3 * kMaxStages + 4) + slack + (P.stage_out ? kEpiBytes * P.n_epi : P.stack ? kStackEpiBytes : 0);
  int stages = (int)((224 * 1024 - fixed) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  { const char* e = getenv("BDP_GEMM_STAGES"); if (e && atoi(e) >= 2 && atoi(e) < stages) stages = atoi(e); }
  if (stages < 2) stages = 2;
  P.stages = stages;
  const size_t smem = stages * stage_bytes + fixed;

  CUtensorMap tmA, tmB;
  const uint64_t ga = P.a_g ? (uint64_t)G : 1, gb = P.b_g ? (uint64_t)G : 1;
  const uint64_t a_gs = a_gstride ? (uint64_t)a_gstride : (uint64_t)a_ld * (uint64_t)(a_major ? K : M);
  const uint64_t b_gs = b_gstride ? (uint64_t)b_gstride : (uint64_t)b_ld * (uint64_t)(b_major ? K : N);
  int st;
  if (!a_major) st = make_map(&tmA, A, (uint64_t)K, (uint64_t)M, ga, (uint64_t)a_ld, a_gs, kBK, (uint32_t)P.a_rows, false);
  else st = make_map(&tmA, A, (uint64_t)M, (uint64_t)K, ga, (uint64_t)a_ld, a_gs, 32, kBK, true);
  if (st != BDP_OK) return st;
  if (!b_major) st = make_map(&tmB, B, (uint64_t)K, (uint64_t)N, gb, (uint64_t)b_ld, b_gs, kBK, (uint32_t)bn, false);
  else st = make_map(&tmB, B, (uint64_t)N, (uint64_t)K, gb, (uint64_t)b_ld, b_gs, 32, kBK, true);
  if (st != BDP_OK) return st;

  const int mode = !precise ? 0 : P.stack ? 2 : 1;
  void (*kern)(CUtensorMap, CUtensorMap, GemmParams) =
      mode == 0 ? gemm_tf32_kernel<0> : mode == 1 ? gemm_tf32_kernel<1> : gemm_tf32_kernel<2>;
  // the opt-in is per device: a process that drives several GPUs sets it once on each
  static bool attr_set[3][64] = {};
  int dev = 0;
  BDP_CUDA_CALL(cudaGetDevice(&dev));
  const bool tracked = dev >= 0 && dev < 64;
  if (!tracked || !attr_set[mode][dev]) {
    cudaFuncAttributes fa;
    BDP_CUDA_CALL(cudaFuncGetAttributes(&fa, kern));
    // opt-in limit is 227 KB per block INCLUDING the kernel's static shared memory
    BDP_CUDA_CALL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       227 * 1024 - (int)fa.sharedSizeBytes));
    if (tracked) attr_set[mode][dev] = true;
  }
  kern<<<(unsigned)grid, kGemmThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tmA, tmB, P);
  BDP_CUDA_CHECK_LAUNCH("gemm_tf32_kernel");
  return BDP_OK;
}

// number of K splits bdp_gemm_tf32 will actually use for (K, splits) — callers size the partial buffer
extern "C" int bdp_gemm_tf32_splits(int64_t K, int splits) {
  int kb_total = (int)((K + kBK - 1) / kBK);
  if (splits < 1) splits = 1;
  if (splits > kb_total) splits = kb_total;
  const int per = (kb_total + splits - 1) / splits;
  return (kb_total + per - 1) / per;
}
