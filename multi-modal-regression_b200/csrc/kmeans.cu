// k-means dictionary learning (learnKmeansDictionary.py:41-42, scikit-learn Lloyd) — the iteration
// driver: M-step finalisation fused with the cross-GPU exchange of the cluster sums.
//
// Every rank accumulates its shard's per-cluster sums into a double-buffered int64 fixed-point
// accumulator that lives in peer-accessible (symmetric) memory.  The exchange is ONE kernel per
// iteration and rank (`kmeans_xfin_kernel`):
//   1. the rank raises its flag in every peer's flag array (release, system scope) — the E+M kernel
//      before it on the stream has finished, so its accumulator is complete;
//   2. every block waits until all peers' flags show this iteration (acquire, system scope);
//   3. each thread owns one cluster, loads that cluster's 2d+1 accumulator words from EVERY rank over
//      NVLink (or one in-switch multimem.ld_reduce per word) and adds them: integer sums, so every
//      rank derives bit-identical centres whatever the world size;
//   4. the new centre, its squared shift, the empty-cluster census and the convergence decision are
//      computed in the same kernel (the last block to finish reduces the block partials in a fixed
//      order), and the accumulator of the NEXT iteration is zeroed (all peers have finished reading it:
//      their flag for this iteration was raised after their previous exchange kernel completed).
// There is no NCCL call, no memset and no host round trip inside the iteration loop; a stopped fit
// (converged / needs the host for an empty-cluster relocation) turns the remaining launches of the
// loop into no-ops through a device flag.
#include <stddef.h>

#include "assign_common.cuh"

// internal entry points of assign.cu (same shared object, not part of the C ABI)
int bdpi_keygrid_build(const double* centers, int K, int d, void* grid, int64_t grid_bytes,
                       const int* stop, const bdpi_grid_peers* peers, cudaStream_t st,
                       const int* cells, int n_cells, int header_mode);
void bdpi_keygrid_parts(void* grid, int K, int d, void** hdr, void** cf32);
int bdpi_lloyd_step_grid(const double* x, int64_t N, int d, const double* centers, int K,
                         const void* grid, int64_t grid_bytes, int32_t* labels, int64_t* acc,
                         int fix_hi_bits, int64_t* stats, double* inertia, int update,
                         const int* stop, const unsigned long long* gflags, int gworld,
                         unsigned long long gflag_value, int incremental, cudaStream_t st);
int bdpi_lloyd_step(const double* x, int64_t N, int d, const double* centers, int K,
                    int32_t* labels, int64_t* acc, int fix_hi_bits, int64_t* stats, double* inertia,
                    int update, const int* stop, cudaStream_t st);

using bdp_assign::ld_acquire_sys;
using bdp_assign::st_release_sys;

namespace {

constexpr int kMaxWorld = BDP_KMEANS_MAX_RANKS;
constexpr int kXfThreads = 128;
constexpr int kMaxBlocks = 64;
constexpr int kMaxK = 8192;

// Control block of a fit (device memory, zero-initialised by the caller).  The first 40 bytes are the
// public part (struct bdp_kmeans_status in include/bdpose.h).
struct KmCtl {
  int state;               // BDP_KMEANS_RUNNING / _STRICT / _TOL / _NEEDS_HOST
  int reserved;
  long long iter_done;     // iterations completed so far
  long long changed;       // labels changed in the last completed iteration (all ranks)
  long long n_empty;       // empty clusters in the last completed iteration
  double shift2;           // sum ||c_new - c_old||^2 of the last completed iteration
  unsigned int ticket;
  unsigned int reserved2;
  long long p_cnt[kMaxBlocks];
  int p_idx[kMaxBlocks];
  int p_empty[kMaxBlocks];
  long long cnt[kMaxK];    // global member counts of the last exchange
  double sh[kMaxK];        // per-cluster squared shift
};
static_assert(offsetof(KmCtl, ticket) == 40, "public part of KmCtl");

struct XfinParams {
  unsigned long long* xchg[kMaxWorld];   // exchange buffer of every rank: [acc0 A][acc1 A][flags 8][gflags 8]
  const unsigned long long* mc;          // multicast mapping of the exchange buffer, or NULL
  int world, rank;
  int K;
  int A;                                 // K*(2d+1) + 2 accumulator words per buffer
  int parity;
  int check;                             // apply scikit-learn's stopping rules
  int incremental;                       // the accumulators persist: carry them over instead of zeroing
  unsigned long long flag_value;         // global iteration index + 1
  double inv_scale_lo;
  double tol_abs;
  const double* c_old;
  double* c_new;
  KmCtl* ctl;
  // fixed-geometry key grid of the fit (NULL: the build's own header kernel does this): the exchange
  // kernel renews the fp32 copy of the keys and the build counters for the next iteration's build
  float4* cf32;
  bdp_assign::GridHdr* ghdr;
};

__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// in-switch sum over every rank's copy of one accumulator word (NVLS)
__device__ __forceinline__ unsigned long long multimem_add_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// (hi, lo) limbs -> correctly rounded double of  hi*2^32 + lo  (|value| < 2^95), then * scale.
__device__ __forceinline__ double limbs_to_double(long long hi, long long lo, double inv_scale_lo) {
  // value = hi*2^32 + lo as a signed 128-bit integer; lo is a sum of non-negative 32-bit limbs
  __int128 v = ((__int128)hi << 32) + (__int128)lo;
  const bool neg = v < 0;
  unsigned __int128 m = neg ? (unsigned __int128)(-v) : (unsigned __int128)v;
  if (m == 0) return 0.0;
  const unsigned long long top = (unsigned long long)(m >> 64);
  double r;
  if (top == 0) {
    r = __ull2double_rn((unsigned long long)m);       // u64 -> double rounds to nearest even once
  } else {
    const int shift = 64 - __clzll((long long)top);   // bits to drop so that 64 remain
    unsigned long long kept = (unsigned long long)(m >> shift);
    const unsigned __int128 dropped = m & ((((unsigned __int128)1) << shift) - 1);
    if (dropped != 0) kept |= 1ull;                   // sticky bit, 11 bits below the double mantissa
    r = __ull2double_rn(kept) * __longlong_as_double((long long)(1023 + shift) << 52);   // * 2^shift
  }
  r *= inv_scale_lo;                                  // power of two: exact
  return neg ? -r : r;
}

// sum over ranks of accumulator word `w` of the current parity buffer
__device__ __forceinline__ unsigned long long xsum(const XfinParams& P, int w) {
  const size_t off = (size_t)P.parity * P.A + w;
  if (P.mc != nullptr) return multimem_add_u64(P.mc + off);
  unsigned long long v[kMaxWorld];
#pragma unroll
  for (int r = 0; r < kMaxWorld; ++r) v[r] = r < P.world ? ld_relaxed_sys(P.xchg[r] + off) : 0ull;
  unsigned long long s = 0ull;
#pragma unroll
  for (int r = 0; r < kMaxWorld; ++r) s += v[r];
  return s;
}

template <int D>
__global__ void __launch_bounds__(kXfThreads) kmeans_xfin_kernel(const XfinParams P) {
  KmCtl* ctl = P.ctl;
  if (*reinterpret_cast<const volatile int*>(&ctl->state) != 0) return;
  constexpr int W = 2 * D + 1;
  __shared__ long long s_cnt[kXfThreads / 32];
  __shared__ int s_idx[kXfThreads / 32], s_emp[kXfThreads / 32];
  __shared__ double s_red[kXfThreads];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned long long* local = P.xchg[P.rank];
  unsigned long long* flags = local + 2 * (size_t)P.A;

  if (P.world > 1) {
    // 1. publish: this rank's accumulator for the iteration is complete (the E+M kernel precedes us)
    if (blockIdx.x == 0 && tid < P.world && tid != P.rank) {
      __threadfence_system();
      st_release_sys(P.xchg[tid] + 2 * (size_t)P.A + P.rank, P.flag_value);
    }
    // 2. wait for every peer's accumulator
    if (tid < P.world && tid != P.rank) {
      while (ld_acquire_sys(flags + tid) < P.flag_value) __nanosleep(40);
    }
    __syncthreads();
  }

  // 3. one cluster per thread: global sums -> centre, shift
  const int j = blockIdx.x * kXfThreads + tid;
  long long cnt = -1;
  int empty = 0;
  if (j < P.K) {
    // all loads of the cluster's words from all ranks are issued before the first one is consumed:
    // a dependent add after every load would serialise 8 x (2d+1) NVLink round trips (~2 us each)
    unsigned long long a[W];
    if (P.mc != nullptr) {
#pragma unroll
      for (int w = 0; w < W; ++w) a[w] = multimem_add_u64(P.mc + (size_t)P.parity * P.A + j * W + w);
    } else {
      unsigned long long v[kMaxWorld][W];
#pragma unroll
      for (int r = 0; r < kMaxWorld; ++r) {
        if (r < P.world) {
#pragma unroll
          for (int w = 0; w < W; ++w) v[r][w] = ld_relaxed_sys(P.xchg[r] + (size_t)P.parity * P.A + j * W + w);
        }
      }
#pragma unroll
      for (int w = 0; w < W; ++w) a[w] = 0ull;
#pragma unroll
      for (int r = 0; r < kMaxWorld; ++r) {
        if (r < P.world) {
#pragma unroll
          for (int w = 0; w < W; ++w) a[w] += v[r][w];
        }
      }
    }
    cnt = (long long)a[2 * D];
    ctl->cnt[j] = cnt;
    double sh = 0.0;
    if (cnt > 0) {
      const double alpha = 1.0 / (double)cnt;        // sklearn _average_centers: c *= 1/weight
      float cf[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const double c = limbs_to_double((long long)a[2 * k], (long long)a[2 * k + 1], P.inv_scale_lo) * alpha;
        P.c_new[(size_t)j * D + k] = c;
        cf[k] = (float)c;
        const double df = c - P.c_old[(size_t)j * D + k];
        sh += df * df;
      }
      if (P.cf32 != nullptr) P.cf32[j] = make_float4(cf[0], cf[1], cf[2], cf[3]);
    } else {
      empty = 1;                                      // the last block fills these in
    }
    ctl->sh[j] = sh;
  }
  // block partial: heaviest cluster (first maximum, np.argmax) and the empty census
  long long bc = cnt;
  int bi = j < P.K ? j : 0x7fffffff, ne = empty;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long oc = __shfl_xor_sync(BDP_FULL_MASK, bc, o);
    const int oi = __shfl_xor_sync(BDP_FULL_MASK, bi, o);
    if (oc > bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
  }
  ne = warp_sum(ne);
  if (lane == 0) { s_cnt[warp] = bc; s_idx[warp] = bi; s_emp[warp] = ne; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kXfThreads / 32; ++w) {
      if (s_cnt[w] > bc || (s_cnt[w] == bc && s_idx[w] < bi)) { bc = s_cnt[w]; bi = s_idx[w]; }
      ne += s_emp[w];
    }
    ctl->p_cnt[blockIdx.x] = bc; ctl->p_idx[blockIdx.x] = bi; ctl->p_empty[blockIdx.x] = ne;
  }
  // 4. prepare the OTHER parity buffer for the next iteration's E+M kernel (every peer has raised its
  //    flag for this iteration, i.e. finished reading that buffer in its previous exchange kernel):
  //    zero it, or — incremental M-step — carry this rank's sums over (the next E+M kernel only moves
  //    the rotations whose label changes); the {changed, unused} counters always restart at 0
  {
    const unsigned long long* cur = local + (size_t)P.parity * P.A;
    unsigned long long* nxt = local + (size_t)(P.parity ^ 1) * P.A;
    for (int i = blockIdx.x * kXfThreads + tid; i < P.A; i += gridDim.x * kXfThreads)
      nxt[i] = (P.incremental && i < P.A - 2) ? cur[i] : 0ull;
  }
  // 5. the last block to arrive closes the iteration (block barrier, then ONE cumulative fence)
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = (atomicAdd(&ctl->ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const volatile KmCtl* vc = ctl;
  // block partials -> heaviest cluster / empty census: one entry per lane (all loads in flight at
  // once; a serial loop of dependent L2 round trips cost more than the rest of the kernel), reduced
  // with shuffles.  The result does not depend on the order: (max count, lowest index), integer sum.
  long long big_cnt = -1;
  int big = 0x7fffffff, n_empty = 0;
  {
    long long c0 = -1, c1 = -1;
    int i0 = 0x7fffffff, i1 = 0x7fffffff, e0 = 0, e1 = 0;
    if (lane < (int)gridDim.x) { c0 = vc->p_cnt[lane]; i0 = vc->p_idx[lane]; e0 = vc->p_empty[lane]; }
    if (lane + 32 < (int)gridDim.x) { c1 = vc->p_cnt[lane + 32]; i1 = vc->p_idx[lane + 32]; e1 = vc->p_empty[lane + 32]; }
    if (c1 > c0 || (c1 == c0 && i1 < i0)) { c0 = c1; i0 = i1; }
    e0 += e1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const long long oc = __shfl_xor_sync(BDP_FULL_MASK, c0, o);
      const int oi = __shfl_xor_sync(BDP_FULL_MASK, i0, o);
      if (oc > c0 || (oc == c0 && oi < i0)) { c0 = oc; i0 = oi; }
      e0 += __shfl_xor_sync(BDP_FULL_MASK, e0, o);
    }
    big_cnt = c0; big = i0; n_empty = e0;               // every warp computes the same values
  }
  if (n_empty > 0) {
    // sklearn _average_centers on a cluster that stayed empty: it copies row `big` as it stands when
    // the in-order loop reaches j — already averaged if big < j, still the raw SUM if big > j (only
    // reachable when relocation bails out; kept for fidelity).  Rare path: one thread per cluster.
    for (int jj = tid; jj < P.K; jj += kXfThreads) {
      if (vc->cnt[jj] != 0) continue;
      const double alpha = (big < jj && big_cnt > 0) ? 1.0 / (double)big_cnt : 1.0;
      double sh = 0.0;
      float cf[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = 0; k < D; ++k) {
        const unsigned long long hi = xsum(P, big * W + 2 * k), lo = xsum(P, big * W + 2 * k + 1);
        const double c = limbs_to_double((long long)hi, (long long)lo, P.inv_scale_lo) * alpha;
        P.c_new[(size_t)jj * D + k] = c;
        cf[k] = (float)c;
        const double df = c - P.c_old[(size_t)jj * D + k];
        sh += df * df;
      }
      if (P.cf32 != nullptr) P.cf32[jj] = make_float4(cf[0], cf[1], cf[2], cf[3]);
      ctl->sh[jj] = sh;
    }
    __syncthreads();
  }
  // shift^2 in a FIXED order (every rank must take the same stopping decision)
  // (per thread: up to 8 independent loads per pass, then a fixed left-to-right sum)
  double acc = 0.0;
  for (int j0 = tid; j0 < P.K; j0 += 8 * kXfThreads) {
    double v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int jj = j0 + q * kXfThreads;
      v[q] = jj < P.K ? vc->sh[jj] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) acc += v[q];
  }
  s_red[tid] = acc;
  __syncthreads();
  for (int o = kXfThreads / 2; o > 0; o >>= 1) {
    if (tid < o) s_red[tid] += s_red[tid + o];
    __syncthreads();
  }
  if (tid == 0) {
    const double shift2 = s_red[0];
    const long long changed = (long long)xsum(P, P.A - 2);
    ctl->shift2 = shift2;
    ctl->changed = changed;
    ctl->n_empty = n_empty;
    ctl->iter_done = (long long)P.flag_value;
    ctl->ticket = 0u;
    if (P.ghdr != nullptr) { P.ghdr->side_next = 0u; P.ghdr->ticket = 0u; }
    int state = BDP_KMEANS_RUNNING;
    if (n_empty > 0) state = BDP_KMEANS_NEEDS_HOST;          // relocation runs on the host (rare)
    else if (P.check) {
      if (changed == 0) state = BDP_KMEANS_STRICT;
      else if (shift2 <= P.tol_abs) state = BDP_KMEANS_TOL;
    }
    __threadfence();
    ctl->state = state;
  }
}

template <int D>
int launch_xfin(const XfinParams& P, cudaStream_t st) {
  const int blocks = (P.K + kXfThreads - 1) / kXfThreads;
  kmeans_xfin_kernel<D><<<blocks, kXfThreads, 0, st>>>(P);
  BDP_CUDA_CHECK_LAUNCH("kmeans_xfin_kernel");
  return BDP_OK;
}

int check_common(int K, int d, int world, int rank, const char* who) {
  BDP_REQUIRE(d == 3 || d == 4, "%s: d must be 3 or 4 (got %d)", who, d);
  BDP_REQUIRE(K >= 1 && K <= kMaxK, "%s: K must be in 1..%d (got %d)", who, kMaxK, K);
  BDP_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world,
              "%s: rank %d of %d (at most %d ranks)", who, rank, world, kMaxWorld);
  return BDP_OK;
}

}  // namespace

extern "C" int64_t bdp_kmeans_ctl_bytes(void) { return (int64_t)sizeof(KmCtl); }

extern "C" int64_t bdp_kmeans_xchg_bytes(int K, int d) {
  if ((d != 3 && d != 4) || K < 1 || K > kMaxK) return -1;
  return (int64_t)(2 * ((int64_t)K * (2 * d + 1) + 2) + 2 * kMaxWorld) * 8;
}

namespace {
int exchange_finalize(void* const* xchg, const void* xchg_multicast, int world, int rank, int K, int d,
                      int fix_hi_bits, int parity, int64_t flag_value, int check, int incremental,
                      double tol_abs, const double* centers_old, double* centers_new, void* ctl,
                      void* fixed_grid, void* stream) {
  int rc = check_common(K, d, world, rank, "kmeans_exchange_finalize");
  if (rc != BDP_OK) return rc;
  BDP_REQUIRE(xchg && centers_old && centers_new && ctl, "kmeans_exchange_finalize: NULL buffer");
  BDP_REQUIRE(fix_hi_bits >= 0 && fix_hi_bits <= 30, "kmeans_exchange_finalize: fix_hi_bits %d", fix_hi_bits);
  XfinParams P = {};
  for (int r = 0; r < world; ++r) {
    BDP_REQUIRE(xchg[r] != nullptr, "kmeans_exchange_finalize: exchange buffer of rank %d is NULL", r);
    P.xchg[r] = reinterpret_cast<unsigned long long*>(xchg[r]);
  }
  P.mc = reinterpret_cast<const unsigned long long*>(xchg_multicast);
  P.world = world; P.rank = rank; P.K = K; P.A = K * (2 * d + 1) + 2;
  P.parity = parity & 1; P.check = check; P.incremental = incremental;
  P.flag_value = (unsigned long long)flag_value;
  P.inv_scale_lo = ldexp(1.0, -(fix_hi_bits + 32));
  P.tol_abs = tol_abs; P.c_old = centers_old; P.c_new = centers_new;
  P.ctl = reinterpret_cast<KmCtl*>(ctl);
  if (fixed_grid != nullptr) {
    void *h, *c;
    bdpi_keygrid_parts(fixed_grid, K, d, &h, &c);
    P.ghdr = reinterpret_cast<bdp_assign::GridHdr*>(h);
    P.cf32 = reinterpret_cast<float4*>(c);
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return d == 3 ? launch_xfin<3>(P, st) : launch_xfin<4>(P, st);
}
}  // namespace

extern "C" int bdp_kmeans_exchange_finalize(void* const* xchg, const void* xchg_multicast, int world,
                                            int rank, int K, int d, int fix_hi_bits, int parity,
                                            int64_t flag_value, int check, int incremental,
                                            double tol_abs, const double* centers_old,
                                            double* centers_new, void* ctl, void* stream) {
  return exchange_finalize(xchg, xchg_multicast, world, rank, K, d, fix_hi_bits, parity, flag_value,
                           check, incremental, tol_abs, centers_old, centers_new, ctl, nullptr, stream);
}

// n_iters Lloyd iterations as one launch sequence: [key-grid build, E+M step, exchange+finalise] x n.
extern "C" int bdp_kmeans_run(const double* x, int64_t N, int d, double* centers2, int K, void* grid,
                              int64_t grid_bytes, void* const* grid_peers, int32_t* labels,
                              void* const* xchg, const void* xchg_multicast, int world, int rank,
                              int fix_hi_bits, int64_t iter0, int n_iters, int check, int incremental,
                              double tol_abs, void* ctl, void* const* em_events,
                              const int32_t* cells, int n_cells, void* stream) {
  int rc = check_common(K, d, world, rank, "kmeans_run");
  if (rc != BDP_OK) return rc;
  BDP_REQUIRE(centers2 && labels && xchg && ctl, "kmeans_run: NULL buffer");
  BDP_REQUIRE(N >= 0 && (x != nullptr || N == 0), "kmeans_run: x is NULL");
  BDP_REQUIRE(iter0 >= 0 && n_iters >= 0, "kmeans_run: negative iteration range");
  BDP_REQUIRE(cells == nullptr || (grid != nullptr && n_cells >= 0), "kmeans_run: a cell list needs the key grid");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int A = K * (2 * d + 1) + 2;
  const int* stop = reinterpret_cast<const int*>(ctl);          // KmCtl::state
  // sharded key-grid build: every rank builds one slab and stores it into every rank's grid
  const bool shard = grid != nullptr && grid_peers != nullptr && world > 1;
  bdpi_grid_peers gp = {};
  if (shard) {
    gp.world = world; gp.rank = rank;
    for (int r = 0; r < world; ++r) {
      BDP_REQUIRE(grid_peers[r] && xchg[r], "kmeans_run: peer buffers of rank %d are NULL", r);
      gp.grid[r] = grid_peers[r];
      gp.gflags[r] = reinterpret_cast<unsigned long long*>(xchg[r]) + 2 * (size_t)A + kMaxWorld;
    }
  }
  for (int it = 0; it < n_iters; ++it) {
    const int64_t gi = iter0 + it;
    const int cur = (int)(gi & 1);
    double* c_cur = centers2 + (size_t)cur * K * d;
    double* c_new = centers2 + (size_t)(cur ^ 1) * K * d;
    int64_t* acc = reinterpret_cast<int64_t*>(xchg[rank]) + (size_t)cur * A;
    auto mark = [&](int k) -> int {                 // benchmark instrumentation: 4 events per iteration
      if (em_events) BDP_CUDA_CALL(cudaEventRecord(reinterpret_cast<cudaEvent_t>(em_events[4 * it + k]), st));
      return BDP_OK;
    };
    if ((rc = mark(0)) != BDP_OK) return rc;
    if (grid) {
      // (a rank without rows still builds its slab: the peers' queries read it)
      gp.flag_value = (unsigned long long)(gi + 1);
      if (N > 0 || shard) {
        // fixed-geometry grid (cells != NULL): the header kernel runs once per call, after that the
        // exchange kernel of the previous iteration has renewed the counters and the fp32 keys
        rc = bdpi_keygrid_build(c_cur, K, d, grid, grid_bytes, stop, shard ? &gp : nullptr, st, cells,
                                n_cells, cells == nullptr ? 0 : (it == 0 ? 1 : 2));
        if (rc != BDP_OK) return rc;
      }
      if ((rc = mark(1)) != BDP_OK) return rc;
      if (N > 0) {
        rc = bdpi_lloyd_step_grid(x, N, d, c_cur, K, grid, grid_bytes, labels, acc, fix_hi_bits,
                                  acc + A - 2, nullptr, 1, stop, shard ? gp.gflags[rank] : nullptr,
                                  world, gp.flag_value, incremental, st);
      }
      if (rc == BDP_OK) rc = mark(2);
    } else if (N > 0) {
      BDP_REQUIRE(!incremental, "kmeans_run: the incremental M-step needs the key grid");
      rc = bdpi_lloyd_step(x, N, d, c_cur, K, labels, acc, fix_hi_bits, acc + A - 2, nullptr, 1,
                           stop, st);
    }
    if (rc != BDP_OK) return rc;
    rc = exchange_finalize(xchg, xchg_multicast, world, rank, K, d, fix_hi_bits, cur, gi + 1, check,
                           incremental, tol_abs, c_cur, c_new, ctl,
                           (grid != nullptr && cells != nullptr && (N > 0 || shard)) ? grid : nullptr, stream);
    if (rc != BDP_OK) return rc;
    if ((rc = mark(3)) != BDP_OK) return rc;
  }
  return BDP_OK;
}

// ---- single-buffer forms kept for callers that drive the iteration themselves -------------------
namespace {
struct FinParams {
  const long long* acc;
  int K;
  double inv_scale_lo;
  const double* c_old;
  double* c_new;
  double* shift2;
  long long* n_empty;
};

// M-step finalisation from an already-summed accumulator (one block; the rare host-driven path:
// empty-cluster relocation, gloo-driven loops).
template <int D>
__global__ void __launch_bounds__(256) kmeans_finalize_kernel(const FinParams P) {
  __shared__ long long s_best_cnt[8];
  __shared__ int s_best_idx[8];
  __shared__ double s_red[8];
  __shared__ int s_cnt_empty[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long bc = -1;
  int bi = 0x7fffffff, ne = 0;
  for (int j = threadIdx.x; j < P.K; j += blockDim.x) {
    const long long c = P.acc[(size_t)j * (2 * D + 1) + 2 * D];
    if (c > bc) { bc = c; bi = j; }
    ne += (c == 0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long oc = __shfl_xor_sync(BDP_FULL_MASK, bc, o);
    const int oi = __shfl_xor_sync(BDP_FULL_MASK, bi, o);
    if (oc > bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
  }
  ne = warp_sum(ne);
  if (lane == 0) { s_best_cnt[warp] = bc; s_best_idx[warp] = bi; s_cnt_empty[warp] = ne; }
  __syncthreads();
  bc = s_best_cnt[0]; bi = s_best_idx[0]; ne = s_cnt_empty[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
    if (s_best_cnt[w] > bc || (s_best_cnt[w] == bc && s_best_idx[w] < bi)) {
      bc = s_best_cnt[w]; bi = s_best_idx[w];
    }
    ne += s_cnt_empty[w];
  }
  const int big = bi;
  double sh = 0.0;
  for (int j = threadIdx.x; j < P.K; j += blockDim.x) {
    const long long* a = P.acc + (size_t)j * (2 * D + 1);
    const long long cnt = a[2 * D];
    const long long* src = cnt > 0 ? a : P.acc + (size_t)big * (2 * D + 1);
    const long long scnt = src[2 * D];
    double alpha = 1.0;
    if (cnt > 0) alpha = 1.0 / (double)cnt;
    else if (big < j && scnt > 0) alpha = 1.0 / (double)scnt;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const double c = limbs_to_double(src[2 * k], src[2 * k + 1], P.inv_scale_lo) * alpha;
      P.c_new[(size_t)j * D + k] = c;
      const double df = c - P.c_old[(size_t)j * D + k];
      sh += df * df;
    }
  }
  sh = warp_sum(sh);
  __syncthreads();
  if (lane == 0) s_red[warp] = sh;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
    if (P.shift2) *P.shift2 = t;
    if (P.n_empty) *P.n_empty = ne;
  }
}
}  // namespace

extern "C" int bdp_kmeans_finalize(const int64_t* acc, int K, int d, int fix_hi_bits,
                                   const double* centers_old, double* centers_new, double* shift2,
                                   int64_t* n_empty, void* stream) {
  BDP_REQUIRE(acc && centers_old && centers_new, "kmeans_finalize: NULL buffer");
  BDP_REQUIRE(d == 3 || d == 4, "kmeans_finalize: d must be 3 or 4");
  BDP_REQUIRE(K >= 1, "kmeans_finalize: K must be >= 1");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  FinParams P = {reinterpret_cast<const long long*>(acc), K, ldexp(1.0, -(fix_hi_bits + 32)),
                 centers_old, centers_new, shift2, reinterpret_cast<long long*>(n_empty)};
  if (d == 3) kmeans_finalize_kernel<3><<<1, 256, 0, st>>>(P);
  else kmeans_finalize_kernel<4><<<1, 256, 0, st>>>(P);
  BDP_CUDA_CHECK_LAUNCH("kmeans_finalize_kernel");
  return BDP_OK;
}

// One Lloyd iteration as a single launch sequence on ONE caller-owned accumulator: zero it, rebuild
// the key grid for the current centres, E+M step, and (single rank) the M-step finalisation.
extern "C" int bdp_kmeans_iteration(const double* x, int64_t N, int d, const double* centers, int K,
                                    void* grid, int64_t grid_bytes, int32_t* labels,
                                    int64_t* acc_stats, int fix_hi_bits, double* inertia, int update,
                                    double* centers_new, double* shift2, int64_t* n_empty,
                                    void* stream) {
  BDP_REQUIRE(acc_stats != nullptr, "kmeans_iteration: acc_stats is NULL");
  BDP_REQUIRE(d == 3 || d == 4, "kmeans_iteration: d must be 3 or 4 (got %d)", d);
  BDP_REQUIRE(K >= 1, "kmeans_iteration: K must be >= 1");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t n_acc = (size_t)K * (2 * d + 1);
  BDP_CUDA_CALL(cudaMemsetAsync(acc_stats, 0, (n_acc + 2) * sizeof(int64_t), st));
  int rc;
  if (grid) {
    rc = bdpi_keygrid_build(centers, K, d, grid, grid_bytes, nullptr, nullptr, st, nullptr, 0, 0);
    if (rc != BDP_OK) return rc;
    rc = bdpi_lloyd_step_grid(x, N, d, centers, K, grid, grid_bytes, labels, acc_stats, fix_hi_bits,
                              acc_stats + n_acc, inertia, update, nullptr, nullptr, 0, 0ull, 0, st);
  } else {
    rc = bdpi_lloyd_step(x, N, d, centers, K, labels, acc_stats, fix_hi_bits, acc_stats + n_acc,
                         inertia, update, nullptr, st);
  }
  if (rc != BDP_OK) return rc;
  if (centers_new)
    return bdp_kmeans_finalize(acc_stats, K, d, fix_hi_bits, centers, centers_new, shift2, n_empty,
                               stream);
  return BDP_OK;
}
