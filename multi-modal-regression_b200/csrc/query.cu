// The pruned nearest-key query through the key grid (binDeltaGenerators.py:27-30 kmeans.predict +
// residual; the E-step / M-step accumulation of learnKmeansDictionary.py:41-42), second generation.
//
// What bounded the first kernel (assign_grid_kernel, profiles/r1d_ncu_assign.csv): not DRAM (26 %)
// but the L1/LSU path (84 %) and the issue slots (58 %) — per point it issued 25 memory instructions
// (coalesced loads staged through shared memory by hand, scattered 64-bit loads of the fp64 centre,
// per-lane stores) and ~317 instructions in all.  This kernel moves every streaming byte with the TMA
// engine instead of the LSU:
//   * each WARP owns a private ring of shared-memory stages; one lane issues `cp.async.bulk` copies
//     of whole warp tiles (128 fp32 / 64 fp64 rotations = 1.5 KB, and the previous labels in Lloyd
//     mode) that complete on the stage's mbarrier — 3 stages in flight per warp, no block barrier in
//     the steady state, warps run decoupled;
//   * a lane reads its 4 (fp32) / 2 (fp64) consecutive rotations with three 128-bit shared loads;
//   * residuals and labels are written to a shared staging buffer (128-bit stores) and leave with one
//     `cp.async.bulk` store per tile and array;
//   * the fp64 dictionary lives in shared memory next to the fp32 screening records, so the exact
//     residual is three 64-bit shared loads instead of three scattered global loads;
//   * cell index arithmetic is 32-bit, the point -> cell map one FMA per coordinate.
// The arithmetic that decides a label is unchanged (fp32 screen of the cell's candidate keys with a
// rigorous error bound, fp64 exact pass on near ties, ascending key order, warp-cooperative scan for
// points outside the grid), so labels and residuals are bit-identical to the brute-force kernel.
#include "assign_common.cuh"

using namespace bdp_assign;

namespace {

constexpr int kMaxStages = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t it = 0;; ++it) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
    if (it > (1u << 27)) __trap();
  }
}
// global -> shared bulk copy (TMA engine), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// shared -> global bulk copy, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy shared writes -> visible to the async proxy (the bulk store that follows)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__host__ __device__ constexpr size_t a16(size_t b) { return (b + 15) & ~(size_t)15; }

struct QueryCfg {
  int nst;         // ring stages per warp (2..4)
  int use_tma;     // every streamed array is 16-byte aligned
  int cd_smem;     // fp64 dictionary staged in shared memory
  int acc_smem;    // Lloyd: chunk accumulators in shared memory
};

template <typename T, int D> struct Geo {
  static constexpr int PTS = 2;                                 // consecutive rotations per lane
  static constexpr int WPTS = 32 * PTS;                         // rotations per warp tile
  static constexpr int XB = WPTS * D * (int)sizeof(T);          // bytes of a tile of rotations
  static constexpr int XV = PTS * D * (int)sizeof(T) / 16;      // 128-bit words per lane
  static constexpr int RB = WPTS * D * 4;                       // bytes of a tile of fp32 residuals
  static constexpr int RV = PTS * D * 4 / 8;                    // 64-bit words of residuals per lane
};

// Shared chunk accumulators of a block, per cluster: 4 x D counters of 16-bit chunks of the two
// fixed-point limbs (hi biased by 2^31), the number of operations (to remove the bias) and the SIGNED
// net member count (a rotation that leaves a cluster is added as the negated two-limb number).
template <int D> struct AccLayout { static constexpr int W = 4 * D + 2; };

template <int D, int kQThreads>
__device__ __forceinline__ void flush_acc(unsigned* s_acc32, int K, unsigned long long* acc) {
  constexpr int W = AccLayout<D>::W;
  __syncthreads();
  for (int idx = threadIdx.x; idx < K * (D + 1); idx += kQThreads) {
    const int j = idx / (D + 1), k = idx % (D + 1);
    const unsigned* a = s_acc32 + (size_t)j * W;
    const unsigned nb = a[4 * D];
    if (nb == 0u) continue;
    unsigned long long* g = acc + (size_t)j * (2 * D + 1);
    if (k == D) {
      const long long members = (long long)(int)a[4 * D + 1];
      if (members) atomicAdd(g + 2 * D, (unsigned long long)members);
    } else {
      const long long hi = (long long)a[4 * k] + ((long long)a[4 * k + 1] << 16) -
                           (long long)nb * 2147483648LL;           // remove the +2^31 bias
      const unsigned long long lo = (unsigned long long)a[4 * k + 2] +
                                    ((unsigned long long)a[4 * k + 3] << 16);
      if (hi) atomicAdd(g + 2 * k, (unsigned long long)hi);
      if (lo) atomicAdd(g + 2 * k + 1, lo);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * W; i += kQThreads) s_acc32[i] = 0u;
  __syncthreads();
}

// M-step contribution of ONE rotation (kept out of line: after the first iterations few rotations
// move, and the hot loop keeps its registers): take it out of cluster `from` (>= 0: incremental mode)
// and add it to cluster `to` — shared 32-bit chunk counters, or global 64-bit atomics for dictionaries
// too large for shared memory.  Leaving a cluster = adding the NEGATED two-limb number
// (-(hi 2^32 + lo) = (-hi - [lo != 0]) 2^32 + (2^32 - lo) mod 2^32), so one code path serves both.
template <int D>
__device__ __forceinline__ void acc_one(long long hi, unsigned lo, int k, unsigned* a) {
  const unsigned ub = (unsigned)(hi + 2147483648LL);      // |hi| < 2^30 by the choice of the scale
  atomicAdd(a + 4 * k, ub & 0xFFFFu);
  atomicAdd(a + 4 * k + 1, ub >> 16);
  atomicAdd(a + 4 * k + 2, lo & 0xFFFFu);
  atomicAdd(a + 4 * k + 3, lo >> 16);
}

// The same additions for a GROUP of lanes that target the same cluster (rows sorted by cell: most
// lanes of a warp do): the 16-bit chunks are summed across the group with the REDUX unit (32 x 2^16
// fits easily) and the group's leader issues ONE atomic per counter — without this, 32 lanes hammering
// the same 14 shared counters serialise (first iteration of a sorted fit: 451 us instead of 212).
// Smallest group worth aggregating: the groups of a warp take the REDUX path one after the other
// (~150 cycles each), while direct atomics of all lanes run together and only serialise g-fold on a
// counter shared by g lanes (14 g cycles): aggregation pays for the big groups of the first iterations
// of a sorted fit, not for the two or three rows that cross a cell boundary later on.
constexpr int kAggMin = 16;
template <int D>
__device__ __forceinline__ void acc_group(unsigned grp, bool leader, const long long hi[D],
                                          const unsigned lo[D], unsigned member_delta, unsigned* a) {
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const unsigned ub = (unsigned)(hi[k] + 2147483648LL);
    const unsigned s0 = __reduce_add_sync(grp, ub & 0xFFFFu), s1 = __reduce_add_sync(grp, ub >> 16);
    const unsigned s2 = __reduce_add_sync(grp, lo[k] & 0xFFFFu), s3 = __reduce_add_sync(grp, lo[k] >> 16);
    if (leader) {
      atomicAdd(a + 4 * k, s0);
      atomicAdd(a + 4 * k + 1, s1);
      atomicAdd(a + 4 * k + 2, s2);
      atomicAdd(a + 4 * k + 3, s3);
    }
  }
  if (leader) {
    const unsigned n = (unsigned)__popc(grp);
    atomicAdd(a + 4 * D, n);
    atomicAdd(a + 4 * D + 1, member_delta * n);
  }
}

// (arguments by value: a pointer to the caller's copy of the rotation forces it through local memory)
template <int D>
__device__ __noinline__ void lloyd_move(double x0, double x1, double x2, double x3, int to, int from,
                                        double scale_hi, unsigned* s_acc32, unsigned long long* acc) {
  constexpr int W = AccLayout<D>::W;
  const double x[4] = {x0, x1, x2, x3};
  long long hi[D], lw[D];
#pragma unroll
  for (int k = 0; k < D; ++k) to_limbs(x[k], scale_hi, hi[k], lw[k]);
  if (s_acc32 != nullptr) {
    const unsigned act = __activemask();
    const unsigned lane = threadIdx.x & 31u;
    {
      unsigned lo[D];
#pragma unroll
      for (int k = 0; k < D; ++k) lo[k] = (unsigned)lw[k];
      unsigned* a = s_acc32 + (size_t)to * W;
      const unsigned grp = __match_any_sync(act, to);
      if (__popc(grp) < kAggMin) {
#pragma unroll
        for (int k = 0; k < D; ++k) acc_one<D>(hi[k], lo[k], k, a);
        atomicAdd(a + 4 * D, 1u);
        atomicAdd(a + 4 * D + 1, 1u);
      } else {
        acc_group<D>(grp, lane == (unsigned)(__ffs(grp) - 1), hi, lo, 1u, a);
      }
    }
    const unsigned actf = __ballot_sync(act, from >= 0);
    if (from >= 0) {
      // leaving a cluster = adding the NEGATED two-limb number
      long long nh[D];
      unsigned nl[D];
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const unsigned lo = (unsigned)lw[k];
        nh[k] = -hi[k] - (lo != 0u ? 1 : 0);
        nl[k] = 0u - lo;
      }
      unsigned* b = s_acc32 + (size_t)from * W;
      const unsigned grp = __match_any_sync(actf, from);
      if (__popc(grp) < kAggMin) {
#pragma unroll
        for (int k = 0; k < D; ++k) acc_one<D>(nh[k], nl[k], k, b);
        atomicAdd(b + 4 * D, 1u);
        atomicAdd(b + 4 * D + 1, 0xFFFFFFFFu);             // member count - 1
      } else {
        acc_group<D>(grp, lane == (unsigned)(__ffs(grp) - 1), nh, nl, 0xFFFFFFFFu, b);
      }
    }
  } else {
    if (from >= 0) {
      unsigned long long* a = acc + (size_t)from * (2 * D + 1);
#pragma unroll
      for (int k = 0; k < D; ++k) {
        atomicAdd(a + 2 * k, (unsigned long long)(-hi[k]));
        atomicAdd(a + 2 * k + 1, (unsigned long long)(-lw[k]));
      }
      atomicAdd(a + 2 * D, ~0ull);                        // count - 1
    }
    unsigned long long* a = acc + (size_t)to * (2 * D + 1);
#pragma unroll
    for (int k = 0; k < D; ++k) {
      atomicAdd(a + 2 * k, (unsigned long long)hi[k]);
      atomicAdd(a + 2 * k + 1, (unsigned long long)lw[k]);
    }
    atomicAdd(a + 2 * D, 1ull);
  }
}

// LAB64: assign mode writes int64 labels (else int32); Lloyd labels are int32
template <typename T, int D, bool LLOYD, bool LAB64, int kQThreads>
__global__ void __launch_bounds__(kQThreads, 1) query_kernel(const AssignParams P, const QueryCfg C) {
  constexpr int kQWarps = kQThreads / 32;
  if (P.stop != nullptr && *reinterpret_cast<const volatile int*>(P.stop) != 0) return;
  using G_ = Geo<T, D>;
  constexpr int PTS = G_::PTS, WPTS = G_::WPTS, XB = G_::XB, RB = G_::RB, RV = G_::RV;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long s_full[kQWarps][kMaxStages];
  __shared__ double s_red[2][kQWarps];
  __shared__ float s_cmax[kQWarps];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int K = P.K, nst = C.nst;
  // ---- shared-memory carve-up ---------------------------------------------------------------------
  unsigned char* sp = smem_raw;
  float4* s_rec = reinterpret_cast<float4*>(sp); sp += (size_t)K * 16;
  float* s_cn = reinterpret_cast<float*>(sp); sp += (D == 4 ? a16((size_t)K * 4) : 0);
  double* s_cd = reinterpret_cast<double*>(sp); sp += C.cd_smem ? a16((size_t)K * D * 8) : 0;
  unsigned* s_acc32 = reinterpret_cast<unsigned*>(sp);
  sp += (LLOYD && C.acc_smem) ? a16((size_t)K * (4 * D + 2) * 4) : 0;
  // Lloyd mode: the previous / new labels of a tile are one coalesced 8-byte load / store per lane
  // (256 contiguous bytes per warp) — fewer instructions than a second bulk copy per tile plus the
  // staging store, proxy fence and bulk store on the way out (with rows in cell order the kernel is
  // bound by issue slots, not by the L1 any more); the next tile's labels are fetched while the
  // current tile is processed.
  constexpr int LB_IN = 0;
  constexpr int LB_OUT = LLOYD ? 0 : WPTS * (LAB64 ? 8 : 4);            // labels out (assign mode)
  constexpr int RB_OUT = LLOYD ? 0 : RB;
  const int per_warp = nst * (XB + LB_IN) + 2 * (LB_OUT + RB_OUT);
  unsigned char* wbase = sp + (size_t)warp * per_warp;
  unsigned char* w_x = wbase;                                           // [nst][XB]
  unsigned char* w_lin = w_x + nst * XB;                                // [nst][LB_IN]
  unsigned char* w_lout = w_lin + nst * LB_IN;                          // [2][LB_OUT]
  unsigned char* w_rout = w_lout + 2 * LB_OUT;                          // [2][RB_OUT]

  const bool acc_in_smem = LLOYD && P.update && C.acc_smem;
  if (LLOYD && C.acc_smem) {
    for (int i = threadIdx.x; i < K * (4 * D + 2); i += kQThreads) s_acc32[i] = 0u;
  }
  // stage the fp32 screening records (+ fp64 keys) of the whole dictionary, max ||c||^2 for the bound
  float cmax2 = 0.f;
  for (int j = threadIdx.x; j < K; j += kQThreads) {
    const double* c = P.centers + (int64_t)j * D;
    double cn = 0.0;
    float m2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const double ck = __ldg(c + k);
      cn += ck * ck;
      m2[k] = (float)(-2.0 * ck);
      if (C.cd_smem) s_cd[j * D + k] = ck;
    }
    if (D == 3) s_rec[j] = make_float4(m2[0], m2[1], m2[2], (float)cn);
    else { s_rec[j] = make_float4(m2[0], m2[1], m2[2], m2[3]); s_cn[j] = (float)cn; }
    cmax2 = fmaxf(cmax2, (float)cn * 1.000001f);
  }
  cmax2 = warp_max(cmax2);
  if (lane == 0) {
    s_cmax[warp] = cmax2;
    for (int s = 0; s < nst; ++s) mbar_init(smem_u32(&s_full[warp][s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // multi-GPU build: the grid is complete once every rank has published its slab
  if (P.gflags != nullptr && threadIdx.x < P.gworld) {
    while (ld_acquire_sys(P.gflags + threadIdx.x) < P.gflag_value) __nanosleep(40);
  }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < kQWarps; ++w) cmax2 = fmaxf(cmax2, s_cmax[w]);
  // tau = err_coef * (|x| + cmax)^2 <= 2 * err_coef * (|x|^2 + cmax^2): no square root per point
  const float coef2 = 2.f * P.err_coef;
  const float tau0 = coef2 * cmax2;

  // point -> cell: t = x * inv + (-origin * inv), one FMA per coordinate in the input type (the
  // rounding is covered by kBoxEps, see assign_common.cuh)
  float g_inv[D], g_off[D];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    g_inv[k] = (float)P.ghdr->inv_cell[k];
    g_off[k] = (float)(-P.ghdr->origin[k] * P.ghdr->inv_cell[k]);
  }
  const int G = P.ghdr->G;
  const bool g_on = P.ghdr->enabled != 0;
  const bool want_sq = LLOYD ? (P.inertia != nullptr) : (P.min_sqdist != nullptr);
  const char* recs = reinterpret_cast<const char*>(s_rec);
  const uint4* gfine = P.gfine;
  const T* __restrict__ xg = reinterpret_cast<const T*>(P.x);

  const int64_t N = P.N;
  const int n_tiles = (int)((N + WPTS - 1) / WPTS);         // N < 2^31 rotations per launch
  const int n_full = C.use_tma ? (int)(N / WPTS) : 0;       // tiles the TMA path moves
  const int TW = (int)gridDim.x * kQWarps;
  const int gw = (int)blockIdx.x * kQWarps + warp;
  const int n_it = (n_tiles + TW - 1) / TW;                 // the same trip count for every warp
  // TMA tiles of this warp: tile = gw + k * TW < n_full
  const int m_tma = n_full > gw ? (n_full - gw + TW - 1) / TW : 0;

  auto issue = [&](int k, int s) {                          // lane 0: this warp's k-th TMA tile -> stage s
    const int64_t tile = gw + (int64_t)k * TW;
    const uint32_t bar = smem_u32(&s_full[warp][s]);
    mbar_expect_tx(bar, (uint32_t)XB);
    bulk_load(smem_u32(w_x + s * XB), xg + tile * (int64_t)(WPTS * D), XB, bar);
  };
  // previous labels of tile t for this lane (Lloyd; the 16-byte aligned path only, else read per tile)
  const bool lab_vec = LLOYD && C.use_tma;
  auto fetch_labels = [&](int t) -> int2 {
    if (lab_vec && t < n_full) return __ldcs(reinterpret_cast<const int2*>(P.labels32 + (int64_t)t * WPTS) + lane);
    return make_int2(0, 0);
  };
  int2 lab_next = fetch_labels(gw);
  if (lane == 0) {
    for (int k = 0; k < m_tma && k < nst; ++k) issue(k, k);
  }
  int stage = 0;                                            // ring position of the next TMA tile
  uint32_t phase = 0;

  int changed = 0;
  double inertia = 0.0;
  const int flush_every = 65536 / (kQWarps * WPTS);
  int since_flush = 0;
  int64_t done_tiles = 0;                                   // tiles this warp has processed

  for (int it = 0; it < n_it; ++it) {
    const int tile = gw + it * TW;
    if (tile < n_tiles) {
      const int64_t base = (int64_t)tile * WPTS;
      const int64_t rem = N - base;
      const int nval = rem < WPTS ? (int)rem : WPTS;
      const bool tma = tile < n_full;
      const int s = tma ? stage : 0;
      unsigned char* xs = w_x + s * XB;
      int2 lab_cur = lab_next;
      if (tma) {
        if (LLOYD) lab_next = fetch_labels(tile + TW);       // in flight while this tile is processed
        mbar_wait(smem_u32(&s_full[warp][s]), phase);
      } else {
        // generic fill (unaligned arrays, or the partial last tile): coalesced loads, zero padding
        T* xw = reinterpret_cast<T*>(xs);
        for (int i = lane; i < WPTS * D; i += 32) xw[i] = i < nval * D ? __ldcs(xg + base * D + i) : (T)0;
        if (LLOYD) {
          lab_cur.x = lane * PTS < nval ? __ldcs(P.labels32 + base + lane * PTS) : 0;
          lab_cur.y = lane * PTS + 1 < nval ? __ldcs(P.labels32 + base + lane * PTS + 1) : 0;
        }
        __syncwarp();
      }
      // The stage stays valid for the whole tile (its reload is issued at the end): registers only
      // hold the fp32 copies of the lane's PTS consecutive rotations; the few places that need the
      // original values (exact re-check, residual, M-step of a moved rotation) read the stage again.
      const T* xst = reinterpret_cast<const T*>(xs) + lane * (PTS * D);
      int olab[PTS];
      if (LLOYD) { olab[0] = lab_cur.x; olab[PTS - 1] = lab_cur.y; }   // Lloyd data is fp64: PTS == 2

      // phase 1: cells and the 16-byte record of every rotation (all loads in flight together).  The
      // point -> cell map runs in fp32 for fp64 rotations too (the fp32 copy is needed for the screen
      // anyway; its rounding moves the cell coordinate by < 1e-5 cells, covered by kBoxEps)
      uint4 first[PTS];
      float xf[PTS * D];
      bool valid[PTS], slow[PTS];
#pragma unroll
      for (int p = 0; p < PTS; ++p) {
        valid[p] = lane * PTS + p < nval;
        bool ok = g_on && valid[p];
        unsigned cidx = 0, mul = 1;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          xf[p * D + k] = (float)xst[p * D + k];
          const float t = fmaf(xf[p * D + k], g_inv[k], g_off[k]);
          // floor + one unsigned compare: negative, too large and infinite coordinates fail it; a NaN
          // converts to 0 and flows through the candidates of a valid cell — every comparison with
          // it is false, which ends at label 0 exactly like the scan (argmin of an all-NaN row)
          const int ck = __float2int_rd(t);
          ok = ok && ((unsigned)ck < (unsigned)G);
          cidx += (unsigned)ck * mul;
          mul *= (unsigned)G;
        }
        first[p] = __ldg(gfine + (ok ? cidx : 0u));
        slow[p] = valid[p] && !ok;
      }
      // phase 2: fp32 screen over the candidates, fp64 exact pass on near ties
      int label[PTS];
#pragma unroll
      for (int p = 0; p < PTS; ++p) {
        label[p] = 0;
        const unsigned cnt = first[p].x & 0xFFFFu;
        if (valid[p] && cnt == kGridOverflow) slow[p] = true;
        if (!valid[p] || slow[p]) continue;
        const float* xfp = xf + p * D;
        float n2 = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) n2 = fmaf(xfp[k], xfp[k], n2);
        float best = INFINITY, second = INFINITY;
        unsigned boff = 0;                                 // byte offset (id * 16) of the best record
        const unsigned fw[4] = {first[p].x, first[p].y, first[p].z, first[p].w};
        auto screen = [&](unsigned off) {                  // off = key id * 16 (byte offset of its record)
          const float4 cr = *reinterpret_cast<const float4*>(recs + off);
          const float cn = (D == 4) ? s_cn[off >> 4] : 0.f;
          const float d = screen_dist<D>(xfp, cr, cn);
          const bool lt = d < best;
          second = fminf(second, lt ? best : d);
          boff = lt ? off : boff;
          best = fminf(best, d);
        };
        // keys 1..6 sit inline; h7 is key 7 of a short list or the side slot of a long one
#pragma unroll
        for (int e = 1; e <= kFineInline - 1; ++e) {
          if ((unsigned)e > cnt) break;
          const unsigned w = fw[e >> 1];
          screen((e & 1) ? ((w >> 12) & 0xFFFF0u) : ((w << 4) & 0xFFFF0u));
        }
        if (cnt == (unsigned)kFineInline) screen((fw[3] >> 12) & 0xFFFF0u);
        const unsigned short* srec = nullptr;
        if (cnt > (unsigned)kFineInline) {
          // long list (3 % of the cells): keys 7..14 in the first 16 bytes of the side record
          srec = P.gside + (size_t)(fw[3] >> 16) * kSideWidth;
          const uint4 more = __ldg(reinterpret_cast<const uint4*>(srec));
          const unsigned mw[4] = {more.x, more.y, more.z, more.w};
#pragma unroll
          for (int e = 7; e <= 14; ++e) {
            if ((unsigned)e > cnt) break;
            const unsigned w = mw[(e - 7) >> 1];
            screen(((e - 7) & 1) ? ((w >> 12) & 0xFFFF0u) : ((w << 4) & 0xFFFF0u));
          }
          for (unsigned e = 15; e <= cnt; ++e) screen((unsigned)__ldg(srec + (e - 7)) << 4);
        }
        int bidx = (int)(boff >> 4);
        const float tau = fmaf(coef2, n2, tau0);
        if (!(second - best > tau) && cnt > 1u) {
          // near tie: exact fp64 pass over the same candidates (ascending ids: lowest index wins)
          double bd = INFINITY;
          for (unsigned e = 1; e <= cnt; ++e) {
            int id;
            if (e < (unsigned)kFineInline || cnt == (unsigned)kFineInline) {
              const unsigned w = e < 2 ? fw[0] : (e < 4 ? fw[1] : (e < 6 ? fw[2] : fw[3]));
              id = (int)((e & 1) ? (w >> 16) : (w & 0xFFFFu));
            } else {
              id = (int)__ldg(srec + (e - kFineInline));
            }
            const double* c = P.centers + (int64_t)id * D;
            double sq = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) {
              const double df = (double)xst[p * D + k] - __ldg(c + k);
              sq += df * df;
            }
            if (sq < bd) { bd = sq; bidx = id; }
          }
        }
        label[p] = bidx;
      }
      // slow path: points outside the grid / in overflowed cells, the whole warp scans the dictionary
#pragma unroll
      for (int p = 0; p < PTS; ++p) {
        unsigned m = __ballot_sync(BDP_FULL_MASK, slow[p]);
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          double xe[D];
          const T* xsrc = reinterpret_cast<const T*>(xs) + (src * PTS + p) * D;
#pragma unroll
          for (int k = 0; k < D; ++k) xe[k] = (double)xsrc[k];
          double bd = INFINITY;
          int bi = 0x7fffffff;
          for (int j = lane; j < K; j += 32) {
            const double* c = P.centers + (int64_t)j * D;
            double sq = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) {
              const double df = xe[k] - __ldg(c + k);
              sq += df * df;
            }
            if (sq < bd) { bd = sq; bi = j; }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(BDP_FULL_MASK, bd, o);
            const int oi = __shfl_xor_sync(BDP_FULL_MASK, bi, o);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
          }
          if (bi == 0x7fffffff) bi = 0;                    // all distances NaN: argmin gives 0
          if (lane == src) label[p] = bi;
        }
      }
      // emit
      const int ob = (int)(done_tiles & 1);
      unsigned char* lo = w_lout + ob * LB_OUT;
      unsigned char* ro = w_rout + ob * RB_OUT;
      if (!LLOYD) {
        if (lane == 0) bulk_wait_read<1>();                // the store that last used this buffer is out
        __syncwarp();
      }
      float rout[LLOYD ? 1 : PTS * D];
#pragma unroll
      for (int p = 0; p < PTS; ++p) {
        double diff[D], sq = 0.0;
        if (!LLOYD || want_sq) {
          if (C.cd_smem) {
#pragma unroll
            for (int k = 0; k < D; ++k) diff[k] = (double)xst[p * D + k] - s_cd[label[p] * D + k];
          } else {
            const double* c = P.centers + (int64_t)label[p] * D;
#pragma unroll
            for (int k = 0; k < D; ++k) diff[k] = (double)xst[p * D + k] - __ldg(c + k);
          }
#pragma unroll
          for (int k = 0; k < D; ++k) sq += diff[k] * diff[k];
        }
        if (LLOYD) {
          if (valid[p]) {
            const bool moved = olab[p] != label[p];
            changed += moved;
            inertia += sq;
            // incremental M-step: the integer sums are exact, so "remove from the old cluster, add
            // to the new one" for the rotations that moved equals a recomputation bit for bit —
            // and after the first iterations most rotations stay where they are
            if (P.update && (moved || !P.incremental)) {
              lloyd_move<D>((double)xst[p * D], (double)xst[p * D + 1], (double)xst[p * D + 2],
                            D == 4 ? (double)xst[p * D + D - 1] : 0.0, label[p],
                            (P.incremental && moved) ? olab[p] : -1, P.scale_hi,
                            acc_in_smem ? s_acc32 : nullptr, P.acc);
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < D; ++k) rout[p * D + k] = (float)diff[k];
          if (P.min_sqdist && valid[p]) __stcs(P.min_sqdist + base + lane * PTS + p, sq);
        }
      }
      // labels (and residuals) of the lane -> staging buffer, vector stores
      if (LLOYD) {
        if (tma && lab_vec) {
          __stcs(reinterpret_cast<int2*>(P.labels32 + base) + lane, make_int2(label[0], label[PTS - 1]));
        } else {
          if (valid[0]) __stcs(P.labels32 + base + lane * PTS, label[0]);
          if (valid[PTS - 1]) __stcs(P.labels32 + base + lane * PTS + (PTS - 1), label[PTS - 1]);
        }
      } else if (LAB64) {
        if (PTS == 4) {
          uint4* d4 = reinterpret_cast<uint4*>(lo) + lane * 2;
          d4[0] = make_uint4((unsigned)label[0], 0u, (unsigned)label[1], 0u);
          d4[1] = make_uint4((unsigned)label[PTS - 2], 0u, (unsigned)label[PTS - 1], 0u);
        } else {
          *(reinterpret_cast<uint4*>(lo) + lane) = make_uint4((unsigned)label[0], 0u, (unsigned)label[PTS - 1], 0u);
        }
      } else {
        if (PTS == 4)
          *(reinterpret_cast<int4*>(lo) + lane) = make_int4(label[0], label[1], label[PTS - 2], label[PTS - 1]);
        else
          *(reinterpret_cast<int2*>(lo) + lane) = make_int2(label[0], label[PTS - 1]);
      }
      if (!LLOYD && P.residual) {
        if (RV % 2 == 0) {
          uint4* dst = reinterpret_cast<uint4*>(ro) + lane * (RV / 2);
#pragma unroll
          for (int v = 0; v < RV / 2; ++v)
            dst[v] = make_uint4(__float_as_uint(rout[4 * v]), __float_as_uint(rout[4 * v + 1]),
                                __float_as_uint(rout[4 * v + 2]), __float_as_uint(rout[4 * v + 3]));
        } else {                                            // fp64 rotations, d = 3: 24 bytes per lane
          uint2* dst = reinterpret_cast<uint2*>(ro) + lane * RV;
#pragma unroll
          for (int v = 0; v < RV; ++v)
            dst[v] = make_uint2(__float_as_uint(rout[2 * v]), __float_as_uint(rout[2 * v + 1]));
        }
      }
      if (LLOYD) {
        // (labels already stored from registers)
      } else if (tma) {
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (LAB64) bulk_store(P.labels64 + base, smem_u32(lo), LB_OUT);
          else bulk_store(P.labels32 + base, smem_u32(lo), LB_OUT);
          if (P.residual) bulk_store(P.residual + base * D, smem_u32(ro), RB_OUT);
          bulk_commit();
        }
      } else {
        __syncwarp();
        if (!LAB64) {
          const int* l32 = reinterpret_cast<const int*>(lo);
          for (int i = lane; i < nval; i += 32) __stcs(P.labels32 + base + i, l32[i]);
        } else {
          const long long* l64 = reinterpret_cast<const long long*>(lo);
          for (int i = lane; i < nval; i += 32)
            __stcs(reinterpret_cast<long long*>(P.labels64) + base + i, l64[i]);
        }
        if (!LLOYD && P.residual) {
          const float* r32 = reinterpret_cast<const float*>(ro);
          for (int i = lane; i < nval * D; i += 32) __stcs(P.residual + base * D + i, r32[i]);
        }
        __syncwarp();
      }
      ++done_tiles;
      __syncwarp();                                         // every lane is done with the stage
      if (tma) {
        if (lane == 0 && it + nst < m_tma) issue(it + nst, s);
        if (++stage == nst) { stage = 0; phase ^= 1u; }
      }
    }
    if (acc_in_smem && ++since_flush == flush_every) {
      flush_acc<D, kQThreads>(s_acc32, K, P.acc);
      since_flush = 0;
    }
  }
  if (lane == 0) bulk_wait_all();                           // staged stores have left shared memory

  if (LLOYD) {
    if (acc_in_smem) flush_acc<D, kQThreads>(s_acc32, K, P.acc);
    changed = warp_sum(changed);
    inertia = warp_sum(inertia);
    if (lane == 0) { s_red[0][warp] = (double)changed; s_red[1][warp] = inertia; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double ch = 0.0, in = 0.0;
      for (int w = 0; w < kQWarps; ++w) { ch += s_red[0][w]; in += s_red[1][w]; }
      if (ch != 0.0) atomicAdd(P.stats, (unsigned long long)ch);
      if (P.inertia) atomicAdd(P.inertia, in);
    }
  }
}

template <typename T, int D, bool LLOYD, bool LAB64, int kQThreads>
int launch_query_nt(const AssignParams& P, cudaStream_t st) {
  constexpr int kQWarps = kQThreads / 32;
  using G_ = Geo<T, D>;
  QueryCfg C = {};
  const int K = P.K;
  auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  C.use_tma = aligned(P.x) && aligned(P.labels32) && aligned(P.labels64) && aligned(P.residual) ? 1 : 0;
  const size_t lim = 227 * 1024 - 2048;                     // static shared memory + margin
  const size_t rec = (size_t)K * 16 + (D == 4 ? a16((size_t)K * 4) : 0);
  const size_t cd = a16((size_t)K * D * 8);
  const size_t acc = a16((size_t)K * (4 * D + 2) * 4);
  const int lb_in = 0;                                        // (Lloyd labels travel through registers)
  const int lb_out = LLOYD ? 0 : G_::WPTS * (LAB64 ? 8 : 4);
  const int rb_out = LLOYD ? 0 : G_::RB;
  auto warp_bytes = [&](int nst) { return (size_t)nst * (G_::XB + lb_in) + 2 * (size_t)(lb_out + rb_out); };
  // what goes to shared memory, in order of value: ring stages, Lloyd accumulators, fp64 keys
  C.nst = 3;
  C.acc_smem = (LLOYD && P.update && K <= kMaxSmemAccK) ? 1 : 0;
  C.cd_smem = (!LLOYD || P.inertia != nullptr) ? 1 : 0;
  auto total = [&]() {
    return rec + (C.cd_smem ? cd : 0) + (C.acc_smem ? acc : 0) + kQWarps * warp_bytes(C.nst);
  };
  if (total() > lim) C.cd_smem = 0;
  if (total() > lim) C.nst = 2;
  if (total() > lim) C.acc_smem = 0;
  if (total() > lim) return BDP_ERR_UNSUPPORTED;           // huge dictionary: first-generation kernel
  if (C.nst < kMaxStages && total() + kQWarps * (size_t)(G_::XB + lb_in) <= lim) C.nst += 1;
  const size_t smem = total();
  auto kern = query_kernel<T, D, LLOYD, LAB64, kQThreads>;
  BDP_CUDA_CALL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_tiles = ceil_div64(P.N, (int64_t)G_::WPTS);
  int64_t blocks = bdp_num_sms();
  const int64_t need = ceil_div64(n_tiles, kQWarps);
  if (blocks > need) blocks = need;
  if (blocks < 1) blocks = 1;
  kern<<<(unsigned)blocks, kQThreads, smem, st>>>(P, C);
  BDP_CUDA_CHECK_LAUNCH("query_kernel");
  return BDP_OK;
}

// One CTA per SM.  Threads per CTA trade registers against warps.  Label generation (rows in arbitrary
// order: bound by L1 wavefronts and latency): 1024 threads leave 64 registers and the compiler
// rematerialises tile indices and parameters all over the loop (137 us for 10 M rotations), 512 / 128
// starve the schedulers (170 us), 768 / 80 measured 133 us and 896 / 72 129 us.  The Lloyd step runs on
// rows in cell order (cellsort.cu) and is bound by issue slots: there the 32 warps of the 1024-thread
// form win (E+M mean 123 us against 127 us at 896 and 138 us at 768).  Dictionaries too large for this
// form's shared memory take the brute-force scan.
template <typename T, int D, bool LLOYD, bool LAB64>
int launch_query(const AssignParams& P, cudaStream_t st) {
  return launch_query_nt<T, D, LLOYD, LAB64, LLOYD ? 1024 : 896>(P, st);
}

}  // namespace

// Returns BDP_ERR_UNSUPPORTED when the request needs the first-generation kernel (both label widths
// at once, or neither).
int bdpi_query_grid(const AssignParams& P, int x_dtype, int d, bool lloyd, cudaStream_t st) {
  if (lloyd) {
    return d == 3 ? launch_query<double, 3, true, false>(P, st) : launch_query<double, 4, true, false>(P, st);
  }
  if ((P.labels64 != nullptr) == (P.labels32 != nullptr)) return BDP_ERR_UNSUPPORTED;
  const bool l64 = P.labels64 != nullptr;
  if (x_dtype == BDP_F32) {
    if (d == 3) return l64 ? launch_query<float, 3, false, true>(P, st) : launch_query<float, 3, false, false>(P, st);
    return l64 ? launch_query<float, 4, false, true>(P, st) : launch_query<float, 4, false, false>(P, st);
  }
  if (d == 3) return l64 ? launch_query<double, 3, false, true>(P, st) : launch_query<double, 3, false, false>(P, st);
  return l64 ? launch_query<double, 4, false, true>(P, st) : launch_query<double, 4, false, false>(P, st);
}
