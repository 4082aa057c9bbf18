// (c) Nearest-dictionary-key assignment, residuals, and the k-means Lloyd step for sm_100a.
//
// Reference call sites: kmeans.predict + residual (binDeltaGenerators.py:27-30, 78-82),
// KMeans.fit (learnKmeansDictionary.py:41-42; scikit-learn 1.9.0 Lloyd E/M step in
// sklearn/cluster/_k_means_lloyd.pyx:_update_chunk_dense), argmax|K.q| (learnObjectnetModel.py:108),
// Riemannian residual (binDeltaGenerators.py:131-137).
//
// Brute-force tiled distance/argmin: the dictionary chunk is staged in shared memory as
// (-2c, ||c||^2) fp32 records and broadcast to every lane; each thread owns PTS points in
// registers and tracks (best, second-best, argmin).  fp32 is only a SCREEN: a point whose
// best/second gap is inside the rigorous fp32 error bound is pushed to a per-block list and
// re-assigned by a whole warp in fp64 (direct sum of squared differences, lowest index on ties),
// so the labels are those of an fp64 evaluation — what scikit-learn computes on float64 data.
//
// Lloyd accumulation is exact: every coordinate is split into two int64 fixed-point limbs and
// added with integer atomics (shared memory per block, then global), so the cluster sums do not
// depend on the order of accumulation, the grid, or how the points are sharded over GPUs.
#include <stdlib.h>

#include "assign_common.cuh"

using namespace bdp_assign;

namespace {

template <typename T, int D>
__device__ __forceinline__ void load_point(const T* __restrict__ x, int64_t i, double xd[D],
                                           float xf[D]) {
#pragma unroll
  for (int k = 0; k < D; ++k) {
    xd[k] = (double)x[i * D + k];
    xf[k] = (float)xd[k];
  }
}

// Writes every output of one resolved point; in LLOYD mode also accumulates it.
template <int D, bool LLOYD>
__device__ __forceinline__ void emit_point(const AssignParams& P, int64_t i, const double xd[D],
                                           int label, unsigned long long* s_acc, bool acc_in_smem,
                                           int& changed, double& inertia) {
  const double* c = P.centers + (int64_t)label * D;
  double diff[D], sq = 0.0;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    diff[k] = xd[k] - __ldg(c + k);
    sq += diff[k] * diff[k];
  }
  if (LLOYD) {
    changed += (P.labels32[i] != label);
    P.labels32[i] = label;
    inertia += sq;
    if (P.update) {
      unsigned long long* a = (acc_in_smem ? s_acc : P.acc) + (size_t)label * (2 * D + 1);
#pragma unroll
      for (int k = 0; k < D; ++k) {
        long long hi, lo;
        to_limbs(xd[k], P.scale_hi, hi, lo);
        atomicAdd(a + 2 * k, (unsigned long long)hi);
        atomicAdd(a + 2 * k + 1, (unsigned long long)lo);
      }
      atomicAdd(a + 2 * D, 1ull);
    }
  } else {
    if (P.labels32) P.labels32[i] = label;
    if (P.labels64) P.labels64[i] = (int64_t)label;
    if (P.residual) {
#pragma unroll
      for (int k = 0; k < D; ++k) P.residual[i * D + k] = (float)diff[k];
    }
    if (P.min_sqdist) P.min_sqdist[i] = sq;
  }
}

template <typename T, int D, bool LLOYD>
__global__ void __launch_bounds__(kThreads) assign_kernel(const AssignParams P) {
  if (P.stop != nullptr && *reinterpret_cast<const volatile int*>(P.stop) != 0) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [chunk float4 recs][chunk float norms (D==4)][recheck list ints][acc (lloyd)]
  float4* s_rec = reinterpret_cast<float4*>(smem_raw);
  const int chunk = P.K < kChunk ? P.K : kChunk;
  float* s_cn = reinterpret_cast<float*>(s_rec + chunk);
  int* s_list = reinterpret_cast<int*>(s_cn + (D == 4 ? chunk : 0));
  unsigned long long* s_acc = reinterpret_cast<unsigned long long*>(
      reinterpret_cast<unsigned char*>(s_list) + (((size_t)kThreads * kPts * 4 + 15) & ~(size_t)15));
  __shared__ int s_nlist;
  __shared__ double s_red[2][kThreads / 32];
  const bool acc_in_smem = LLOYD && P.update && P.K <= kMaxSmemAccK;

  const T* __restrict__ x = reinterpret_cast<const T*>(P.x);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kTile = kThreads * kPts;
  const int64_t n_tiles = (P.N + kTile - 1) / kTile;

  if (acc_in_smem) {
    for (int i = threadIdx.x; i < P.K * (2 * D + 1); i += kThreads) s_acc[i] = 0ull;
  }
  // max_k ||c_k|| (rounded up) for the screening bound: K*D loads per block, nothing next to
  // the N*K distance work, and it keeps the entry point free of host round trips.
  __shared__ float s_cmax[kThreads / 32];
  float cmax = 0.f;
  for (int j = threadIdx.x; j < P.K; j += kThreads) {
    double n2 = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const double ck = __ldg(P.centers + (int64_t)j * D + k);
      n2 += ck * ck;
    }
    cmax = fmaxf(cmax, (float)sqrt(n2) * 1.000001f);
  }
  cmax = warp_max(cmax);
  if (lane == 0) s_cmax[warp] = cmax;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) cmax = fmaxf(cmax, s_cmax[w]);

  int changed = 0;
  double inertia = 0.0;
  bool staged = false;   // a dictionary that fits one chunk is staged once per block

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * kTile;
    double xd[kPts][D];
    float xf[kPts][D];
    float best[kPts], second[kPts], tau[kPts];
    int bidx[kPts];
#pragma unroll
    for (int p = 0; p < kPts; ++p) {
      const int64_t i = base + threadIdx.x + p * kThreads;
      if (i < P.N) {
        load_point<T, D>(x, i, xd[p], xf[p]);
      } else {
#pragma unroll
        for (int k = 0; k < D; ++k) { xd[p][k] = 0.0; xf[p][k] = 0.f; }
      }
      float n2 = 0.f;
#pragma unroll
      for (int k = 0; k < D; ++k) n2 += xf[p][k] * xf[p][k];
      const float s = sqrtf(n2) * 1.0000002f + cmax;
      tau[p] = P.err_coef * s * s;
      best[p] = INFINITY; second[p] = INFINITY; bidx[p] = 0;
    }
    if (threadIdx.x == 0) s_nlist = 0;

    for (int k0 = 0; k0 < P.K; k0 += chunk) {
      const int kc = min(chunk, P.K - k0);
      __syncthreads();   // previous chunk fully consumed (and s_nlist reset visible)
      if (!staged)
      for (int j = threadIdx.x; j < kc; j += kThreads) {
        const double* c = P.centers + (int64_t)(k0 + j) * D;
        double cn = 0.0;
        float m2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < D; ++k) {
          const double ck = __ldg(c + k);
          cn += ck * ck;
          m2[k] = (float)(-2.0 * ck);
        }
        if (D == 3) s_rec[j] = make_float4(m2[0], m2[1], m2[2], (float)cn);
        else { s_rec[j] = make_float4(m2[0], m2[1], m2[2], m2[3]); s_cn[j] = (float)cn; }
      }
      staged = (P.K <= chunk);
      __syncthreads();
#pragma unroll 4
      for (int j = 0; j < kc; ++j) {
        const float4 c = s_rec[j];
        const float cn = (D == 4) ? s_cn[j] : 0.f;
#pragma unroll
        for (int p = 0; p < kPts; ++p) {
          const float d = screen_dist<D>(xf[p], c, cn);
          const bool lt = d < best[p];
          second[p] = fminf(second[p], lt ? best[p] : d);
          bidx[p] = lt ? (k0 + j) : bidx[p];
          best[p] = fminf(best[p], d);
        }
      }
    }

    // resolved points are emitted by their owner; near-ties go to the block list
#pragma unroll
    for (int p = 0; p < kPts; ++p) {
      const int64_t i = base + threadIdx.x + p * kThreads;
      if (i >= P.N) continue;
      if (second[p] - best[p] > tau[p] || P.K == 1) {
        emit_point<D, LLOYD>(P, i, xd[p], bidx[p], s_acc, acc_in_smem, changed, inertia);
      } else {
        const int slot = atomicAdd(&s_nlist, 1);
        s_list[slot] = threadIdx.x + p * kThreads;
      }
    }
    __syncthreads();
    const int nlist = s_nlist;
    for (int e = warp; e < nlist; e += kThreads / 32) {
      const int64_t i = base + s_list[e];
      double xe[D];
#pragma unroll
      for (int k = 0; k < D; ++k) xe[k] = (double)x[i * D + k];
      double bd = INFINITY;
      int bi = 0x7fffffff;
      for (int j = lane; j < P.K; j += 32) {
        const double* c = P.centers + (int64_t)j * D;
        double sq = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          const double df = xe[k] - __ldg(c + k);
          sq += df * df;
        }
        if (sq < bd) { bd = sq; bi = j; }     // ascending j per lane: first minimum kept
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(BDP_FULL_MASK, bd, o);
        const int oi = __shfl_xor_sync(BDP_FULL_MASK, bi, o);
        if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
      }
      if (bi == 0x7fffffff) bi = 0;                      // all distances NaN: argmin gives 0
      if (lane == 0) emit_point<D, LLOYD>(P, i, xe, bi, s_acc, acc_in_smem, changed, inertia);
    }
    // next tile's first __syncthreads() (chunk loop) orders the list reuse
  }

  if (LLOYD) {
    __syncthreads();
    if (acc_in_smem) {
      for (int i = threadIdx.x; i < P.K * (2 * D + 1); i += kThreads) {
        const unsigned long long v = s_acc[i];
        if (v) atomicAdd(P.acc + i, v);
      }
    }
    changed = warp_sum(changed);
    inertia = warp_sum(inertia);
    if (lane == 0) { s_red[0][warp] = (double)changed; s_red[1][warp] = inertia; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double ch = 0.0, in = 0.0;
      for (int w = 0; w < kThreads / 32; ++w) { ch += s_red[0][w]; in += s_red[1][w]; }
      if (ch != 0.0) atomicAdd(P.stats, (unsigned long long)ch);
      if (P.inertia) atomicAdd(P.inertia, in);
    }
  }
}

template <int D>
size_t assign_smem_bytes(int K, bool lloyd_smem_acc) {
  const int chunk = K < kChunk ? K : kChunk;
  size_t b = (size_t)chunk * 16 + (D == 4 ? (size_t)chunk * 4 : 0);
  b += ((size_t)kThreads * kPts * 4 + 15) & ~(size_t)15;
  if (lloyd_smem_acc) b += (size_t)K * (2 * D + 1) * 8;
  return b;
}

template <typename T, int D, bool LLOYD>
int launch_assign(const AssignParams& P, cudaStream_t st) {
  const bool smem_acc = LLOYD && P.update && P.K <= kMaxSmemAccK;
  const size_t smem = assign_smem_bytes<D>(P.K, smem_acc);
  auto kern = assign_kernel<T, D, LLOYD>;
  if (smem > 48 * 1024) {
    BDP_CUDA_CALL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  int per_sm = 0;
  BDP_CUDA_CALL(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t n_tiles = ceil_div64(P.N, (int64_t)kThreads * kPts);
  int64_t blocks = (int64_t)bdp_num_sms() * per_sm;
  if (blocks > n_tiles) blocks = n_tiles;
  if (blocks < 1) blocks = 1;
  kern<<<(unsigned)blocks, kThreads, smem, st>>>(P);
  BDP_CUDA_CHECK_LAUNCH("assign_kernel");
  return BDP_OK;
}



template <bool LLOYD>
int dispatch_assign(AssignParams& P, int x_dtype, int d, cudaStream_t st) {
  P.err_coef = screen_err_coef(d);
  if (x_dtype == BDP_F32) {
    if (d == 3) return launch_assign<float, 3, LLOYD>(P, st);
    return launch_assign<float, 4, LLOYD>(P, st);
  }
  if (d == 3) return launch_assign<double, 3, LLOYD>(P, st);
  return launch_assign<double, 4, LLOYD>(P, st);
}

// ---- key grid: candidate pruning for the nearest-key query -------------------------------------
// The brute-force scan costs N*K pair evaluations (10^10 at the benchmark size) although a point
// can only be nearest to the handful of keys around it.  The key grid is a uniform grid over the
// bounding box of the dictionary (plus a margin); every cell stores the ascending list of keys that
// can be nearest to SOME point of the cell:
//     U(cell)    = min_k maxdist^2(cell, c_k)          (an upper bound on the nearest distance),
//     p(cell)    = the key attaining it (the pivot)
//     cand(cell) = { k : mindist^2(cell, c_k) <= U  and  k beats the pivot SOMEWHERE in the cell }
// (the second test is exact: |x-c_k|^2 - |x-c_p|^2 is linear in x, its minimum over the box sits at
// a corner).  The true nearest key of any point of the cell — and every key tied with it — is in it,
// built in two levels (coarse cells of 4^d fine cells filter the dictionary once, fine cells filter
// their parent's list), ~3*10^7 box tests for K=1000 instead of 5*10^8.  A query then evaluates
// ~5 candidates per point instead of K — exactly the same fp32-screen / fp64-exact arithmetic as the
// brute-force kernel, on a superset of the keys that matter, so the labels are identical.  Points
// outside the grid and cells whose list overflows the 64-byte record take the brute-force slow path.
__host__ __device__ inline int64_t ipow64(int64_t b, int e) {
  int64_t r = 1;
  for (int i = 0; i < e; ++i) r *= b;
  return r;
}

int keygrid_G(int K, int d) {
  const double r = pow((double)K, 1.0 / d);
  int gc = (int)ceil((d == 3 ? 1.6 : 1.0) * r);
  const int hi = d == 3 ? 16 : 6;
  if (gc < 2) gc = 2;
  if (gc > hi) gc = hi;
  return 4 * gc;
}

// Grid geometry from the bounding box of the dictionary; every thread of the (256-thread) block
// takes part, thread 0 writes *hdr.
template <int D>
__device__ __forceinline__ void compute_header(const double* __restrict__ centers, int K, int G,
                                               double margin_frac, GridHdr* hdr) {
  __shared__ double s_lo[8][D], s_hi[8][D];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double lo[D], hi[D];
#pragma unroll
  for (int k = 0; k < D; ++k) { lo[k] = INFINITY; hi[k] = -INFINITY; }
  bool finite = true;
  for (int j = threadIdx.x; j < K; j += blockDim.x) {
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const double c = centers[(int64_t)j * D + k];
      finite = finite && isfinite(c);
      lo[k] = fmin(lo[k], c);
      hi[k] = fmax(hi[k], c);
    }
  }
  const int all_finite = __syncthreads_and(finite ? 1 : 0);
#pragma unroll
  for (int k = 0; k < D; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[k] = fmin(lo[k], __shfl_xor_sync(BDP_FULL_MASK, lo[k], o));
      hi[k] = fmax(hi[k], __shfl_xor_sync(BDP_FULL_MASK, hi[k], o));
    }
    if (lane == 0) { s_lo[warp][k] = lo[k]; s_hi[warp][k] = hi[k]; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ext = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      for (int w = 1; w < 8; ++w) { s_lo[0][k] = fmin(s_lo[0][k], s_lo[w][k]); s_hi[0][k] = fmax(s_hi[0][k], s_hi[w][k]); }
      ext = fmax(ext, s_hi[0][k] - s_lo[0][k]);
    }
    const bool ok = all_finite && ext > 0.0 && isfinite(ext);
    const double margin = ext * margin_frac;
    for (int k = 0; k < 4; ++k) {
      double o = 0.0, c = 1.0;
      if (k < D && ok) {
        o = s_lo[0][k] - margin;
        c = (s_hi[0][k] - s_lo[0][k] + 2.0 * margin) / (double)G;
      }
      hdr->origin[k] = o; hdr->cell[k] = c; hdr->inv_cell[k] = 1.0 / c;
      hdr->origin32[k] = (float)o; hdr->inv_cell32[k] = (float)(1.0 / c);
    }
    hdr->G = G;
    hdr->enabled = ok ? 1 : 0;
  }
  __syncthreads();
}

// squared distance bounds between key c and the box [lo, hi]
template <int D>
__device__ __forceinline__ void box_bounds(const double* __restrict__ c, const double lo[D],
                                           const double hi[D], double& mind2, double& maxd2) {
  mind2 = 0.0; maxd2 = 0.0;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const double ck = __ldg(c + k);
    const double a = ck - lo[k], b = hi[k] - ck;      // >= 0 when inside along this axis
    const double far = fmax(fabs(a), fabs(b));
    const double near = fmax(fmax(-a, -b), 0.0);
    maxd2 += far * far;
    mind2 += near * near;
  }
}

// min over the box of |x - c_k|^2 - |x - c_p|^2 (linear in x: attained at a corner).  Positive means
// key k loses to the pivot key p everywhere in the box, so k can never be the nearest key there.
template <int D>
__device__ __forceinline__ double bisector_min(const double* __restrict__ ck, const double cp[D],
                                               const double lo[D], const double hi[D], double& scale) {
  double f = 0.0, nk = 0.0, np = 0.0;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const double c = __ldg(ck + k);
    const double dlt = c - cp[k];
    f -= 2.0 * dlt * (dlt > 0.0 ? hi[k] : lo[k]);
    nk += c * c;
    np += cp[k] * cp[k];
  }
  scale = 1.0 + nk + np;
  return f + (nk - np);
}

// A block of 8 warps filters the WHOLE dictionary against one box; the warps take contiguous key
// ranges so the compacted list stays in ascending key order.  rec has room for K ids.
template <int D>
__device__ __forceinline__ void filter_box_block(const double* __restrict__ centers, int K,
                                                 const double lo[D], const double hi[D],
                                                 unsigned short* __restrict__ rec) {
  __shared__ double s_u[8];
  __shared__ int s_p[8], s_c[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = (((K + 7) / 8) + 31) & ~31;
  const int b0 = min(K, warp * per), b1 = min(K, b0 + per);
  double u = INFINITY;
  int piv = 0x7fffffff;
  for (int k = b0 + lane; k < b1; k += 32) {
    double mn, mx;
    box_bounds<D>(centers + (int64_t)k * D, lo, hi, mn, mx);
    if (mx < u) { u = mx; piv = k; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ou = __shfl_xor_sync(BDP_FULL_MASK, u, o);
    const int op = __shfl_xor_sync(BDP_FULL_MASK, piv, o);
    if (ou < u || (ou == u && op < piv)) { u = ou; piv = op; }
  }
  if (lane == 0) { s_u[warp] = u; s_p[warp] = piv; }
  __syncthreads();
  u = s_u[0]; piv = s_p[0];
#pragma unroll
  for (int w = 1; w < 8; ++w)
    if (s_u[w] < u || (s_u[w] == u && s_p[w] < piv)) { u = s_u[w]; piv = s_p[w]; }
  if (piv == 0x7fffffff) piv = 0;
  const double thr = u * (1.0 + 1e-9) + 1e-300;
  double cp[D];
#pragma unroll
  for (int k = 0; k < D; ++k) cp[k] = __ldg(centers + (int64_t)piv * D + k);
  auto keeps = [&](int k) {
    double mn, mx, sc;
    box_bounds<D>(centers + (int64_t)k * D, lo, hi, mn, mx);
    const double bm = bisector_min<D>(centers + (int64_t)k * D, cp, lo, hi, sc);
    return mn <= thr && bm <= 1e-9 * sc;
  };
  int cnt = 0;
  for (int k0 = b0; k0 < b1; k0 += 32) {
    const int k = k0 + lane;
    cnt += __popc(__ballot_sync(BDP_FULL_MASK, k < b1 && keeps(k)));
  }
  if (lane == 0) s_c[warp] = cnt;
  __syncthreads();
  int off = 0, total = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) { off += w < warp ? s_c[w] : 0; total += s_c[w]; }
  for (int k0 = b0; k0 < b1; k0 += 32) {
    const int k = k0 + lane;
    const bool keep = k < b1 && keeps(k);
    const unsigned m = __ballot_sync(BDP_FULL_MASK, keep);
    if (keep) rec[1 + off + __popc(m & ((1u << lane) - 1u))] = (unsigned short)k;
    off += __popc(m);
  }
  if (threadIdx.x == 0) rec[0] = (unsigned short)total;
  __syncthreads();
}

// The coarse filter of the fused build, in fp32: keys of the WHOLE dictionary that can be nearest to
// some point of the box [lo, hi], ascending, into rec[0] = count, rec[1..] = ids, with their fp32
// coordinates next to them in s_cf (both shared memory; at most `cap` keys are stored, the count is
// the true one).  Same two tests as box_bounds / bisector_min, on a box widened by 1e-6 and with
// thresholds relaxed by far more than the fp32 rounding — a list may only grow.
template <int D>
__device__ __forceinline__ void coarse_filter_f32(const float4* __restrict__ cf32, int K,
                                                  const float lo[D], const float hi[D],
                                                  unsigned short* __restrict__ rec,
                                                  float4* __restrict__ s_cf, int cap) {
  __shared__ float s_u[8];
  __shared__ int s_p[8], s_c[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = (((K + 7) / 8) + 31) & ~31;          // contiguous key range per warp (<= 512)
  const int b0 = min(K, warp * per), b1 = min(K, b0 + per);
  float u = INFINITY;
  int piv = 0x7fffffff;
  for (int k = b0 + lane; k < b1; k += 32) {
    const float4 c4 = __ldg(cf32 + k);
    const float c[4] = {c4.x, c4.y, c4.z, c4.w};
    float mx = 0.f;
#pragma unroll
    for (int q = 0; q < D; ++q) {
      const float far = fmaxf(fabsf(c[q] - lo[q]), fabsf(hi[q] - c[q]));
      mx = fmaf(far, far, mx);
    }
    if (mx < u) { u = mx; piv = k; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ou = __shfl_xor_sync(BDP_FULL_MASK, u, o);
    const int op = __shfl_xor_sync(BDP_FULL_MASK, piv, o);
    if (ou < u || (ou == u && op < piv)) { u = ou; piv = op; }
  }
  if (lane == 0) { s_u[warp] = u; s_p[warp] = piv; }
  __syncthreads();
  u = s_u[0]; piv = s_p[0];
#pragma unroll
  for (int w = 1; w < 8; ++w)
    if (s_u[w] < u || (s_u[w] == u && s_p[w] < piv)) { u = s_u[w]; piv = s_p[w]; }
  if (piv == 0x7fffffff) piv = 0;
  const float thr = u * (1.f + 1e-4f) + 1e-5f;
  const float4 p4 = __ldg(cf32 + piv);
  const float cp[4] = {p4.x, p4.y, p4.z, p4.w};
  float np = 0.f;
#pragma unroll
  for (int q = 0; q < D; ++q) np = fmaf(cp[q], cp[q], np);
  // keep flags of this warp's range, one ballot per 32 keys (at most 16 chunks)
  unsigned masks[16];
  int cnt = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    masks[i] = 0u;
    const int k0 = b0 + 32 * i;
    if (k0 < b1) {                                       // warp-uniform
      const int k = k0 + lane;
      bool keep = false;
      if (k < b1) {
        const float4 c4 = __ldg(cf32 + k);
        const float c[4] = {c4.x, c4.y, c4.z, c4.w};
        float mn = 0.f, f = 0.f, nk = 0.f;
#pragma unroll
        for (int q = 0; q < D; ++q) {
          const float near = fmaxf(fmaxf(lo[q] - c[q], c[q] - hi[q]), 0.f);
          mn = fmaf(near, near, mn);
          const float dlt = c[q] - cp[q];
          f = fmaf(-2.f * dlt, dlt > 0.f ? hi[q] : lo[q], f);
          nk = fmaf(c[q], c[q], nk);
        }
        keep = mn <= thr && f + (nk - np) <= 1e-3f * (1.f + nk + np);
      }
      masks[i] = __ballot_sync(BDP_FULL_MASK, keep);
      cnt += __popc(masks[i]);
    }
  }
  if (lane == 0) s_c[warp] = cnt;
  __syncthreads();
  int off = 0, total = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) { off += w < warp ? s_c[w] : 0; total += s_c[w]; }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const unsigned m = masks[i];
    if (m & (1u << lane)) {
      const int pos = off + __popc(m & ((1u << lane) - 1u));
      if (pos < cap) {
        const int k = b0 + 32 * i + lane;
        rec[1 + pos] = (unsigned short)k;
        s_cf[pos] = __ldg(cf32 + k);
      }
    }
    off += __popc(m);
  }
  if (threadIdx.x == 0) rec[0] = (unsigned short)(total > 0xFFFF ? 0xFFFF : total);
  __syncthreads();
}

// Where a build writes its cells: the fine / side arrays of every rank of a multi-GPU fit (peer
// addresses; world == 1: this grid only) and the flags that tell the peers the slab is complete.
struct BuildOut {
  uint4* fine[BDP_KMEANS_MAX_RANKS];
  unsigned short* side[BDP_KMEANS_MAX_RANKS];
  unsigned long long* gflags[BDP_KMEANS_MAX_RANKS];
  int world, rank;
  unsigned side_per_rank;        // side records each rank may hand out
  unsigned long long flag_value;
};

// Grid geometry + reset of the build counters: one block.
template <int D>
__global__ void __launch_bounds__(256) keygrid_header_kernel(const double* __restrict__ centers, int K,
                                                             int G, double margin_frac,
                                                             GridHdr* __restrict__ hdr,
                                                             float4* __restrict__ cf32, const int* stop,
                                                             int keep_geometry) {
  if (stop != nullptr && *reinterpret_cast<const volatile int*>(stop) != 0) return;
  if (keep_geometry) {
    // fixed-geometry grid (bdp_keygrid_prepare): only the build counters and the fp32 keys are renewed
    if (threadIdx.x == 0) { hdr->side_next = 0u; hdr->ticket = 0u; }
  } else {
    __shared__ GridHdr s_hdr;
    compute_header<D>(centers, K, G, margin_frac, &s_hdr);
    if (threadIdx.x == 0) { s_hdr.side_next = 0u; s_hdr.ticket = 0u; *hdr = s_hdr; }
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x) {   // fp32 copy of the keys for the box tests
    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < D; ++q) c[q] = (float)centers[(int64_t)k * D + q];
    cf32[k] = make_float4(c[0], c[1], c[2], c[3]);
  }
}

// One BLOCK per coarse cell (4^D fine cells).  Phase (a): all 256 threads filter the whole dictionary
// against the coarse box into an ascending shared-memory list (the keys that can be nearest somewhere
// in the coarse cell).  Phase (b): the fine children filter that list against their own boxes — four
// lanes per child for d = 3 (64 children), one for d = 4 (256 children) — and one lane per child
// assembles the 16-byte record (+ side record) and stores it to every rank's grid (posted 128-bit
// stores over NVLink).  The last block to finish raises this rank's flag on every peer.  No level of
// the hierarchy leaves the block, so a rank that builds 1/8 of the cells pays 1/8 of the time.
// The box tests run in fp32 with relaxed thresholds (a candidate list has to CONTAIN every key that
// can be nearest; a few extra keys only cost query time — the query itself is exact).
constexpr int kCoarseListCap = 1024;
// (62 registers: 4 blocks = 32 warps per SM, 74 % of the issue slots busy, profiles/r2g_cell_kernel_*.
// Capping the registers at 48 / 40 for 5 / 6 resident blocks was measured: 43-44 us instead of 39 —
// the extra warps do not make up for the rematerialised index and box arithmetic.)
template <int D>
__global__ void __launch_bounds__(256) keygrid_cell_kernel(const float4* __restrict__ cf32, int K,
                                                           GridHdr* __restrict__ hdr, const BuildOut out,
                                                           int64_t coarse0, int64_t n_coarse, const int* stop,
                                                           const int* __restrict__ cells) {
  if (stop != nullptr && *reinterpret_cast<const volatile int*>(stop) != 0) return;
  if (n_coarse <= 0) {                                 // a rank with an empty slab only publishes its flag
    if (out.world > 1 && threadIdx.x < out.world) {
      __threadfence_system();
      st_release_sys(out.gflags[threadIdx.x] + out.rank, out.flag_value);
    }
    return;
  }
  constexpr int kChildren = D == 3 ? 64 : 256;
  constexpr int kSubs = 256 / kChildren;               // lanes per child
  extern __shared__ __align__(16) unsigned char s_build[];
  const int cap = K < kCoarseListCap ? K : kCoarseListCap;
  float4* s_cf = reinterpret_cast<float4*>(s_build);                                // [cap] fp32 keys of the list
  unsigned short* s_list = reinterpret_cast<unsigned short*>(s_cf + cap);           // [1 + cap]: count, ids
  unsigned short* s_ids = s_list + ((cap + 1 + 7) & ~7);                            // [kChildren][32]
  const int G = hdr->G, Gc = G / 4;
  const int lane = threadIdx.x & 31;
  // A block walks the slab with the grid's stride: one block per coarse cell on a single GPU; in a
  // sharded build a few cells per block, so that the system-scope fence that orders a block's peer
  // stores before its ticket (one NVLink round trip with the block's resources held) is paid once per
  // block instead of once per cell.
  // (32-bit cell arithmetic: a grid has at most 24^4 fine cells; the 64-bit form paid three emulated
  // divisions per level and block)
  float org[D], cel[D];
#pragma unroll
  for (int k = 0; k < D; ++k) { org[k] = (float)hdr->origin[k]; cel[k] = (float)hdr->cell[k]; }
  const int c0 = (int)coarse0, nc = (int)n_coarse;
  for (int ci = blockIdx.x; ci < nc; ci += gridDim.x) {
  // cells != NULL: the build covers a LIST of coarse cells (those that hold rows of the fit)
  const unsigned parent = cells != nullptr ? (unsigned)__ldg(cells + c0 + ci) : (unsigned)(c0 + ci);
  // boxes in fp32, widened by kBoxEps of a cell (query rounding) + 1e-5 of a cell (their own rounding)
  constexpr float kEps = (float)kBoxEps + 1e-5f;
  int pc[D];                                           // coarse coordinates of the parent
  {
    unsigned pr = parent;
#pragma unroll
    for (int k = 0; k < D; ++k) { pc[k] = (int)(pr % (unsigned)Gc); pr /= (unsigned)Gc; }
  }
  // (a) coarse box -> s_list, s_cf
  {
    float lo[D], hi[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const int ck = pc[k];
      lo[k] = fmaf((float)(4 * ck) - kEps, cel[k], org[k]) - 1e-6f * fabsf(org[k]);
      hi[k] = fmaf((float)(4 * ck + 4) + kEps, cel[k], org[k]) + 1e-6f * fabsf(org[k]);
    }
    coarse_filter_f32<D>(cf32, K, lo, hi, s_list, s_cf, cap);   // ends with __syncthreads()
  }
  const int n_all = (int)s_list[0];
  const bool too_long = n_all > cap;                   // never seen; the children then take the scan
  const int n_src = too_long ? 0 : n_all;
  // (b) fine children
  const int child = threadIdx.x / kSubs, sub = threadIdx.x % kSubs;
  float lo[D], hi[D];
  unsigned cell = 0, mul = 1;
  int ch = child;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const int ck = 4 * pc[k] + (ch & 3);
    ch >>= 2;
    lo[k] = fmaf((float)ck - kEps, cel[k], org[k]) - 1e-6f * fabsf(org[k]);
    hi[k] = fmaf((float)(ck + 1) + kEps, cel[k], org[k]) + 1e-6f * fabsf(org[k]);
    cell += (unsigned)ck * mul;
    mul *= (unsigned)G;
  }
  float u = INFINITY;
  int pj = 0x7fffffff;                                  // pivot: position in the list
  for (int j = sub; j < n_src; j += kSubs) {
    const float4 c4 = s_cf[j];
    const float c[4] = {c4.x, c4.y, c4.z, c4.w};
    float mx = 0.f;
#pragma unroll
    for (int q = 0; q < D; ++q) {
      const float far = fmaxf(fabsf(c[q] - lo[q]), fabsf(hi[q] - c[q]));
      mx = fmaf(far, far, mx);
    }
    if (mx < u) { u = mx; pj = j; }
  }
#pragma unroll
  for (int o = kSubs / 2; o > 0; o >>= 1) {
    const float ou = __shfl_xor_sync(BDP_FULL_MASK, u, o);
    const int op = __shfl_xor_sync(BDP_FULL_MASK, pj, o);
    if (ou < u || (ou == u && op < pj)) { u = ou; pj = op; }
  }
  if (pj == 0x7fffffff) pj = 0;
  const float thr = u * (1.f + 1e-5f) + 1e-5f;
  const float4 p4 = n_src > 0 ? s_cf[pj] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float cp[4] = {p4.x, p4.y, p4.z, p4.w};
  float np = 0.f;
#pragma unroll
  for (int q = 0; q < D; ++q) np = fmaf(cp[q], cp[q], np);
  unsigned short* ids = s_ids + child * kSideWidth;
  int cnt = 0;
  for (int j0 = 0; j0 < n_src; j0 += kSubs) {           // block-uniform trip count: ballots are safe
    const int j = j0 + sub;
    bool keep = false;
    if (j < n_src) {
      const float4 c4 = s_cf[j];
      const float c[4] = {c4.x, c4.y, c4.z, c4.w};
      float mn = 0.f, f = 0.f, nk = 0.f;
#pragma unroll
      for (int q = 0; q < D; ++q) {
        const float near = fmaxf(fmaxf(lo[q] - c[q], c[q] - hi[q]), 0.f);
        mn = fmaf(near, near, mn);
        const float dlt = c[q] - cp[q];
        f = fmaf(-2.f * dlt, dlt > 0.f ? hi[q] : lo[q], f);
        nk = fmaf(c[q], c[q], nk);
      }
      // fp32 rounding of f is < 2e-5 here (|c| < 4, keys within ~2 of the cell): 2e-5 (1 + nk + np)
      keep = mn <= thr && f + (nk - np) <= 2e-5f * (1.f + nk + np);
    }
    const unsigned m = __ballot_sync(BDP_FULL_MASK, keep);
    const unsigned gm = (m >> (lane & ~(kSubs - 1))) & ((1u << kSubs) - 1u);   // this child's lanes
    const int pos = cnt + __popc(gm & ((1u << sub) - 1u));
    if (keep && pos < kSideWidth - 1) ids[pos] = s_list[1 + j];                // ascending key order
    cnt += __popc(gm);
  }
  __syncwarp();
  if (sub == 0) {
    unsigned long long w0 = 0ull, w1 = 0ull;            // halfwords h0..h3, h4..h7
    const unsigned short* srec = nullptr;               // this cell's side record in the LOCAL grid
    bool overflow = cnt > kGridCap || too_long;
    const int n_inline = cnt <= kFineInline ? cnt : kFineInline - 1;
    for (int e = 1; e <= n_inline && !overflow; ++e) {
      const unsigned long long id = ids[e - 1];
      if (e < 4) w0 |= id << (16 * e);
      else w1 |= id << (16 * (e - 4));
    }
    if (!overflow && cnt > kFineInline) {
      // long list: keys 1..6 inline, h7 = slot of a side record with keys 7..cnt
      const unsigned sl = atomicAdd(&hdr->side_next, 1u);
      if (sl < out.side_per_rank) {
        const unsigned slot = (unsigned)out.rank * out.side_per_rank + sl;
        unsigned short* dst = out.side[out.rank] + (size_t)slot * kSideWidth;
        for (int e = kFineInline; e <= cnt; ++e) dst[e - kFineInline] = ids[e - 1];
        w1 |= (unsigned long long)slot << 48;
        srec = dst;
      } else {
        overflow = true;
      }
    }
    w0 = (w0 & ~0xFFFFull) | (overflow ? (unsigned long long)kGridOverflow : (unsigned long long)cnt);
    const uint4 rec = make_uint4((unsigned)w0, (unsigned)(w0 >> 32), (unsigned)w1, (unsigned)(w1 >> 32));
#pragma unroll
    for (int r = 0; r < BDP_KMEANS_MAX_RANKS; ++r)
      if (r < out.world) out.fine[r][cell] = rec;
    if (srec != nullptr && out.world > 1) {
      // the side record goes to the peers as it stands in the local grid (this thread wrote it)
      const size_t off = (size_t)(srec - out.side[out.rank]);
      const uint4* src = reinterpret_cast<const uint4*>(srec);
      uint4 q[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) q[v] = __ldcv(src + v);
      for (int r = 0; r < out.world; ++r) {
        if (r == out.rank) continue;
        uint4* dst = reinterpret_cast<uint4*>(out.side[r] + off);
#pragma unroll
        for (int v = 0; v < 4; ++v) dst[v] = q[v];
      }
    }
  }
  __syncthreads();                                     // the shared lists are reused by the next cell
  }
  if (out.world > 1) {
    // Every store of this block is ordered before its ticket: the block barrier makes them visible to
    // thread 0, whose system-scope fence is cumulative (one fence per block — a fence in every thread
    // made the sharded build slower than the local one).  The last block publishes the slab.
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0) {
      __threadfence_system();
      s_last = (atomicAdd(&hdr->ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last && threadIdx.x < out.world) {
      __threadfence_system();
      st_release_sys(out.gflags[threadIdx.x] + out.rank, out.flag_value);
    }
  }
}

template <bool LLOYD>
int dispatch_assign_grid(AssignParams& P, int x_dtype, int d, cudaStream_t st) {
  P.err_coef = screen_err_coef(d);
  const int rc = bdpi_query_grid(P, x_dtype, d, LLOYD, st);      // the pruned query (query.cu)
  if (rc != BDP_ERR_UNSUPPORTED) return rc;
  return dispatch_assign<LLOYD>(P, x_dtype, d, st);              // e.g. both label widths at once
}

int keygrid_check(const void* grid, int64_t grid_bytes, int K, int d, const char* who) {
  BDP_REQUIRE(grid != nullptr, "%s: grid is NULL", who);
  BDP_REQUIRE(K >= 1 && K <= kGridMaxK, "%s: the key grid supports 1 <= K <= %d (got %d)", who, kGridMaxK, K);
  BDP_REQUIRE(grid_bytes >= bdp_keygrid_bytes(K, d), "%s: grid buffer too small (%lld < %lld)", who,
              (long long)grid_bytes, (long long)bdp_keygrid_bytes(K, d));
  BDP_REQUIRE((reinterpret_cast<uintptr_t>(grid) & 15) == 0, "%s: grid must be 16-byte aligned", who);
  return BDP_OK;
}

// layout: [GridHdr][fine records 16 B][side records 64 B][fp32 copy of the keys, float4 x kGridMaxK]
// side records of a grid: one per 8 fine cells (3 % of the cells have lists longer than 7 keys),
// at most 65535 (the slot is a halfword of the cell record), divisible among up to 8 ranks
int64_t keygrid_side_records(int64_t n_fine) {
  int64_t n = n_fine / 8;
  if (n < 1024) n = 1024;
  if (n > 65528) n = 65528;
  return n & ~(int64_t)7;
}

struct GridPtrs {
  GridHdr* hdr;
  uint4* fine;
  unsigned short* side;
  float4* cf32;
  int G;
  int64_t n_fine, n_coarse, n_side;
};

GridPtrs keygrid_pointers(const void* grid, int K, int d) {
  unsigned char* b = const_cast<unsigned char*>(reinterpret_cast<const unsigned char*>(grid));
  GridPtrs g;
  g.G = keygrid_G(K, d);
  g.n_coarse = ipow64(g.G / 4, d);
  g.n_fine = ipow64(g.G, d);
  g.n_side = keygrid_side_records(g.n_fine);
  g.hdr = reinterpret_cast<GridHdr*>(b);
  g.fine = reinterpret_cast<uint4*>(b + sizeof(GridHdr));
  g.side = reinterpret_cast<unsigned short*>(g.fine + g.n_fine);
  g.cf32 = reinterpret_cast<float4*>(g.side + g.n_side * kSideWidth);
  return g;
}

// ---- key-grid statistics of a query batch (diagnostics: benchmark / DESIGN numbers) ------------------
// stats[0] = points, [1] = points outside the grid (slow path), [2] = points in overflowed cells (slow
// path), [3] = sum of candidate-list lengths over the in-grid points, [4] = max list length seen.
template <typename T, int D>
__global__ void __launch_bounds__(256) keygrid_stats_kernel(const T* __restrict__ x, int64_t N,
                                                            const GridHdr* __restrict__ hdr,
                                                            const uint4* __restrict__ fine,
                                                            unsigned long long* __restrict__ stats) {
  const int G = hdr->G;
  unsigned long long outside = 0, over = 0, len = 0, mx = 0, n = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (int64_t)gridDim.x * blockDim.x) {
    bool ok = hdr->enabled != 0;
    int64_t cidx = 0, mul = 1;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const T org = sizeof(T) == 4 ? (T)hdr->origin32[k] : (T)hdr->origin[k];
      const T inv = sizeof(T) == 4 ? (T)hdr->inv_cell32[k] : (T)hdr->inv_cell[k];
      const T t = (x[i * D + k] - org) * inv;
      ok = ok && (t >= (T)0) && (t < (T)G);
      cidx += (int64_t)(ok ? (int)t : 0) * mul;
      mul *= G;
    }
    ++n;
    if (!ok) { ++outside; continue; }
    const unsigned c = fine[cidx].x & 0xFFFFu;
    if (c == kGridOverflow) { ++over; continue; }
    len += c;
    mx = c > mx ? c : mx;
  }
  atomicAdd(stats + 0, n); atomicAdd(stats + 1, outside); atomicAdd(stats + 2, over);
  atomicAdd(stats + 3, len); atomicMax(stats + 4, mx);
}

// ---- argmax |<key, q>| -----------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) quatdot_kernel(const T* __restrict__ q, int64_t N,
                                                      const double* __restrict__ keys, int K,
                                                      int64_t* __restrict__ bin,
                                                      float* __restrict__ residual) {
  extern __shared__ double s_keys[];   // [K,4]
  for (int i = threadIdx.x; i < K * 4; i += blockDim.x) s_keys[i] = keys[i];
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    double v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (double)q[i * 4 + k];
    double best = -1.0;
    int bi = 0;
    for (int j = 0; j < K; ++j) {
      // np.dot order: ((k0*q0 + k1*q1) + k2*q2) + k3*q3, no fma contraction across the adds
      double d = __dmul_rn(s_keys[j * 4 + 0], v[0]);
      d = __dadd_rn(d, __dmul_rn(s_keys[j * 4 + 1], v[1]));
      d = __dadd_rn(d, __dmul_rn(s_keys[j * 4 + 2], v[2]));
      d = __dadd_rn(d, __dmul_rn(s_keys[j * 4 + 3], v[3]));
      d = fabs(d);
      if (d > best) { best = d; bi = j; }   // np.argmax: first maximum
    }
    if (bin) bin[i] = (int64_t)bi;
    if (residual) {
#pragma unroll
      for (int k = 0; k < 4; ++k) residual[i * 4 + k] = (float)(v[k] - s_keys[bi * 4 + k]);
    }
  }
}

// ---- Rodrigues exp / log in fp64 ----------------------------------------------------------------
// axisAngle.get_R (axisAngle.py:33-41): identity when theta < eps.
__device__ __forceinline__ void aa_exp(const double v[3], double R[9]) {
  const double th = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (th < 1e-6) {
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
    return;
  }
  const double a[3] = {v[0] / th, v[1] / th, v[2] / th};
  double s, c;
  sincos(th, &s, &c);
  const double omc = 1.0 - c;
  const double V[9] = {0.0, -a[2], a[1], a[2], 0.0, -a[0], -a[1], a[0], 0.0};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double VV = V[i * 3 + 0] * V[0 * 3 + j] + V[i * 3 + 1] * V[1 * 3 + j] +
                        V[i * 3 + 2] * V[2 * 3 + j];
      R[i * 3 + j] = (i == j ? 1.0 : 0.0) + s * V[i * 3 + j] + omc * VV;
    }
}

// axisAngle.get_y (axisAngle.py:19-29): zero vector when ||vee((R-R^T)/2)|| <= eps.
__device__ __forceinline__ void rot_log(const double R[9], double y[3]) {
  const double tR = 0.5 * (R[0] + R[4] + R[8] - 1.0);
  const double th = acos(fmin(fmax(tR, -1.0), 1.0));
  const double v[3] = {0.5 * (R[7] - R[5]), 0.5 * (R[2] - R[6]), 0.5 * (R[3] - R[1])};
  const double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (n > 1e-6) {
    y[0] = th * (v[0] / n); y[1] = th * (v[1] / n); y[2] = th * (v[2] / n);
  } else {
    y[0] = y[1] = y[2] = 0.0;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) riem_residual_kernel(const T* __restrict__ x, int64_t N,
                                                            const double* __restrict__ key_rot,
                                                            int K, const int64_t* __restrict__ bin,
                                                            float* __restrict__ rot,
                                                            float* __restrict__ residual) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    const double v[3] = {(double)x[i * 3 + 0], (double)x[i * 3 + 1], (double)x[i * 3 + 2]};
    double R[9];
    aa_exp(v, R);
    if (rot) {
#pragma unroll
      for (int k = 0; k < 9; ++k) rot[i * 9 + k] = (float)R[k];
    }
    if (residual) {
      int64_t b = bin[i];
      b = b < 0 ? 0 : (b >= K ? K - 1 : b);
      const double* Kr = key_rot + b * 9;
      double M[9];   // Key^T R
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          M[r * 3 + c] = __ldg(Kr + 0 * 3 + r) * R[0 * 3 + c] + __ldg(Kr + 1 * 3 + r) * R[1 * 3 + c] +
                         __ldg(Kr + 2 * 3 + r) * R[2 * 3 + c];
      double y[3];
      rot_log(M, y);
#pragma unroll
      for (int k = 0; k < 3; ++k) residual[i * 3 + k] = (float)y[k];
    }
  }
}

__global__ void __launch_bounds__(128) convert_aa_kernel(const double* __restrict__ aa, int64_t N,
                                                         double* __restrict__ rotmat,
                                                         double* __restrict__ quat) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double v[3] = {aa[i * 3 + 0], aa[i * 3 + 1], aa[i * 3 + 2]};
  if (rotmat) {
    double R[9];
    aa_exp(v, R);
#pragma unroll
    for (int k = 0; k < 9; ++k) rotmat[i * 9 + k] = R[k];
  }
  if (quat) {
    // quaternion.convert_dictionary (quaternion.py:79-92): axis = 0 when angle <= eps; the
    // result is renormalised.
    const double ang = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    double a[3] = {0.0, 0.0, 0.0};
    if (ang > 1e-6) { a[0] = v[0] / ang; a[1] = v[1] / ang; a[2] = v[2] / ang; }
    double s, c;
    sincos(ang / 2.0, &s, &c);
    double y[4] = {c, s * a[0], s * a[1], s * a[2]};
    const double n = sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2] + y[3] * y[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k) quat[i * 4 + k] = y[k] / n;
  }
}

// ---- pose targets from Euler angles -------------------------------------------
// helperFunctions.rotation_matrix (37-48): R = Rz(ct) Rx(el) Rz(az), degrees; then axisAngle.get_y
// (19-29) and / or quaternion.get_y (18-29).  The reference does this per image in python
// (learnKmeansDictionary.py:31-37, dataGenerators.py:55-69).
__global__ void __launch_bounds__(128) euler_pose_kernel(const double* __restrict__ euler, int64_t N,
                                                         double* __restrict__ aa,
                                                         double* __restrict__ quat) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double d2r = 0.017453292519943295;
  double sa, ca, sb, cb, sc, cc;
  sincos(euler[i * 3 + 0] * d2r, &sa, &ca);
  sincos(euler[i * 3 + 1] * d2r, &sb, &cb);
  sincos(euler[i * 3 + 2] * d2r, &sc, &cc);
  // Rb Ra, then Rc (Rb Ra): the association order of np.dot(np.dot(Rc, Rb), Ra) differs only by rounding
  const double Ra[9] = {ca, -sa, 0.0, sa, ca, 0.0, 0.0, 0.0, 1.0};
  const double Rb[9] = {1.0, 0.0, 0.0, 0.0, cb, -sb, 0.0, sb, cb};
  const double Rc[9] = {cc, -sc, 0.0, sc, cc, 0.0, 0.0, 0.0, 1.0};
  double T[9], R[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      T[r * 3 + c] = Rc[r * 3 + 0] * Rb[0 * 3 + c] + Rc[r * 3 + 1] * Rb[1 * 3 + c] + Rc[r * 3 + 2] * Rb[2 * 3 + c];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      R[r * 3 + c] = T[r * 3 + 0] * Ra[0 * 3 + c] + T[r * 3 + 1] * Ra[1 * 3 + c] + T[r * 3 + 2] * Ra[2 * 3 + c];
  if (aa) {
    double y[3];
    rot_log(R, y);
#pragma unroll
    for (int k = 0; k < 3; ++k) aa[i * 3 + k] = y[k];
  }
  if (quat) {
    const double tR = 0.5 * (R[0] + R[4] + R[8] - 1.0);
    double th = acos(fmin(fmax(tR, -1.0), 1.0));
    double v[3] = {0.5 * (R[7] - R[5]), 0.5 * (R[2] - R[6]), 0.5 * (R[3] - R[1])};
    const double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (n > 1e-6) { v[0] /= n; v[1] /= n; v[2] /= n; }
    else { th = 0.0; v[0] = v[1] = v[2] = 0.0; }
    double s, c;
    sincos(0.5 * th, &s, &c);
    quat[i * 4 + 0] = c; quat[i * 4 + 1] = s * v[0]; quat[i * 4 + 2] = s * v[1]; quat[i * 4 + 3] = s * v[2];
  }
}

}  // namespace

float bdp_assign::screen_err_coef(int D) {
  // |fl32(dist) - dist| <= (D + 2.1) * 2^-24 * (||x|| + ||c||)^2 (input rounding of x, -2c, ||c||^2
  // plus one rounding per fma); the best/second GAP carries twice that.  x2 safety on top.
  return (float)(2.0 * 2.0 * (D + 5) * 5.9604644775390625e-08);
}

extern "C" int bdp_assign_nearest(const void* x, int x_dtype, int64_t N, int d,
                                  const double* centers, int K, int32_t* labels32,
                                  int64_t* labels64, float* residual, double* min_sqdist,
                                  void* stream) {
  BDP_REQUIRE(N >= 0, "assign_nearest: N < 0");
  if (N == 0) return BDP_OK;
  BDP_REQUIRE(x && centers, "assign_nearest: NULL input");
  BDP_REQUIRE(d == 3 || d == 4, "assign_nearest: d must be 3 or 4 (got %d)", d);
  BDP_REQUIRE(K >= 1, "assign_nearest: K must be >= 1 (got %d)", K);
  BDP_REQUIRE(x_dtype == BDP_F32 || x_dtype == BDP_F64, "assign_nearest: x_dtype %d", x_dtype);
  AssignParams P = {};
  P.x = x; P.N = N; P.centers = centers; P.K = K;
  P.labels32 = labels32; P.labels64 = labels64; P.residual = residual; P.min_sqdist = min_sqdist;
  return dispatch_assign<false>(P, x_dtype, d, reinterpret_cast<cudaStream_t>(stream));
}

int bdpi_lloyd_step(const double* x, int64_t N, int d, const double* centers, int K,
                    int32_t* labels, int64_t* acc, int fix_hi_bits, int64_t* stats, double* inertia,
                    int update, const int* stop, cudaStream_t st) {
  BDP_REQUIRE(N >= 0 && N < (1ll << 30), "kmeans_lloyd_step: N out of range");
  if (N == 0) return BDP_OK;
  BDP_REQUIRE(x && centers && labels && stats, "kmeans_lloyd_step: NULL buffer");
  BDP_REQUIRE(d == 3 || d == 4, "kmeans_lloyd_step: d must be 3 or 4 (got %d)", d);
  BDP_REQUIRE(K >= 1, "kmeans_lloyd_step: K must be >= 1");
  BDP_REQUIRE(!update || acc, "kmeans_lloyd_step: acc is NULL with update=1");
  BDP_REQUIRE(fix_hi_bits >= 0 && fix_hi_bits <= 30, "kmeans_lloyd_step: fix_hi_bits %d",
              fix_hi_bits);
  AssignParams P = {};
  P.x = x; P.N = N; P.centers = centers; P.K = K; P.labels32 = labels;
  P.acc = reinterpret_cast<unsigned long long*>(acc);
  P.scale_hi = ldexp(1.0, fix_hi_bits);
  P.stats = reinterpret_cast<unsigned long long*>(stats);
  P.inertia = inertia; P.update = update; P.stop = stop;
  return dispatch_assign<true>(P, BDP_F64, d, st);
}

extern "C" int bdp_kmeans_lloyd_step(const double* x, int64_t N, int d, const double* centers,
                                     int K, int32_t* labels, int64_t* acc, int fix_hi_bits,
                                     int64_t* stats, double* inertia, int update, void* stream) {
  return bdpi_lloyd_step(x, N, d, centers, K, labels, acc, fix_hi_bits, stats, inertia, update,
                         nullptr, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int bdp_assign_quatdot(const void* q, int q_dtype, int64_t N, const double* keys, int K,
                                  int64_t* bin, float* residual, void* stream) {
  BDP_REQUIRE(N >= 0, "assign_quatdot: N < 0");
  if (N == 0) return BDP_OK;
  BDP_REQUIRE(q && keys, "assign_quatdot: NULL input");
  BDP_REQUIRE(K >= 1 && K <= 4096, "assign_quatdot: K out of range (%d)", K);
  BDP_REQUIRE(q_dtype == BDP_F32 || q_dtype == BDP_F64, "assign_quatdot: q_dtype %d", q_dtype);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int64_t blocks = ceil_div64(N, 256);
  const int64_t cap = (int64_t)bdp_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  const size_t smem = (size_t)K * 4 * sizeof(double);
  if (smem > 48 * 1024) {
    BDP_CUDA_CALL(cudaFuncSetAttribute(quatdot_kernel<float>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BDP_CUDA_CALL(cudaFuncSetAttribute(quatdot_kernel<double>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (q_dtype == BDP_F32)
    quatdot_kernel<float><<<(unsigned)blocks, 256, smem, st>>>(
        reinterpret_cast<const float*>(q), N, keys, K, bin, residual);
  else
    quatdot_kernel<double><<<(unsigned)blocks, 256, smem, st>>>(
        reinterpret_cast<const double*>(q), N, keys, K, bin, residual);
  BDP_CUDA_CHECK_LAUNCH("quatdot_kernel");
  return BDP_OK;
}

extern "C" int bdp_riemannian_residual(const void* x, int x_dtype, int64_t N,
                                       const double* key_rot, int K, const int64_t* bin,
                                       float* rot, float* residual, void* stream) {
  BDP_REQUIRE(N >= 0, "riemannian_residual: N < 0");
  if (N == 0) return BDP_OK;
  BDP_REQUIRE(x != nullptr, "riemannian_residual: x is NULL");
  BDP_REQUIRE(!residual || (key_rot && bin && K >= 1),
              "riemannian_residual: residual needs key_rot, bin and K >= 1");
  BDP_REQUIRE(x_dtype == BDP_F32 || x_dtype == BDP_F64, "riemannian_residual: x_dtype %d", x_dtype);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int64_t blocks = ceil_div64(N, 256);
  const int64_t cap = (int64_t)bdp_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (x_dtype == BDP_F32)
    riem_residual_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(
        reinterpret_cast<const float*>(x), N, key_rot, K, bin, rot, residual);
  else
    riem_residual_kernel<double><<<(unsigned)blocks, 256, 0, st>>>(
        reinterpret_cast<const double*>(x), N, key_rot, K, bin, rot, residual);
  BDP_CUDA_CHECK_LAUNCH("riem_residual_kernel");
  return BDP_OK;
}

extern "C" int bdp_convert_axis_angle(const double* aa, int64_t N, double* rotmat, double* quat,
                                      void* stream) {
  BDP_REQUIRE(N >= 0, "convert_axis_angle: N < 0");
  if (N == 0) return BDP_OK;
  BDP_REQUIRE(aa != nullptr, "convert_axis_angle: aa is NULL");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  convert_aa_kernel<<<(unsigned)ceil_div64(N, 128), 128, 0, st>>>(aa, N, rotmat, quat);
  BDP_CUDA_CHECK_LAUNCH("convert_aa_kernel");
  return BDP_OK;
}

extern "C" int64_t bdp_keygrid_bytes(int K, int d) {
  if ((d != 3 && d != 4) || K < 1 || K > kGridMaxK) return -1;
  const int G = keygrid_G(K, d);
  const int64_t n_fine = ipow64(G, d);
  return (int64_t)sizeof(GridHdr) + n_fine * 16 + keygrid_side_records(n_fine) * kSideWidth * 2 +
         (int64_t)kGridMaxK * 16;
}

// Build for `centers`.  peers == NULL (or one rank): the whole grid, locally.  Several ranks: every
// rank derives the same grid geometry, builds the coarse cells of ITS slab only (whole layers along
// the last axis) and stores their fine / side records into every rank's grid; the query waits for all
// slabs (gflags).
// cells != NULL (fixed-geometry grid of a k-means fit, bdp_keygrid_prepare): only the n_cells listed
// coarse cells are built — the ones that hold rows — and a rank's slab is its share of that list.
// header_mode: 0 = geometry from the dictionary's bounding box (header kernel), 1 = keep the geometry,
// renew counters + fp32 keys (header kernel), 2 = nothing (the exchange kernel of the previous
// iteration has renewed them).
int bdpi_keygrid_build(const double* centers, int K, int d, void* grid, int64_t grid_bytes,
                       const int* stop, const bdpi_grid_peers* peers, cudaStream_t st,
                       const int* cells, int n_cells, int header_mode) {
  BDP_REQUIRE(centers != nullptr, "keygrid_build: NULL centers");
  BDP_REQUIRE(d == 3 || d == 4, "keygrid_build: d must be 3 or 4 (got %d)", d);
  BDP_REQUIRE(cells != nullptr || header_mode == 0, "keygrid_build: header_mode %d needs a cell list", header_mode);
  int rc = keygrid_check(grid, grid_bytes, K, d, "keygrid_build");
  if (rc != BDP_OK) return rc;
  const GridPtrs g = keygrid_pointers(grid, K, d);
  const int G = g.G, Gc = G / 4;
  const double r = pow((double)K, 1.0 / d);
  const double margin_frac = r > 2.0 ? 1.0 / r : 0.5;
  const int world = (peers && peers->world > 1) ? peers->world : 1;
  const int rank = world > 1 ? peers->rank : 0;
  BDP_REQUIRE(world <= BDP_KMEANS_MAX_RANKS && rank >= 0 && rank < world, "keygrid_build: rank %d of %d", rank, world);
  BuildOut out = {};
  out.world = world; out.rank = rank;
  out.side_per_rank = (unsigned)(g.n_side / world);
  out.flag_value = world > 1 ? peers->flag_value : 0ull;
  for (int q = 0; q < world; ++q) {
    BDP_REQUIRE(world == 1 || (peers->grid[q] != nullptr && peers->gflags[q] != nullptr),
                "keygrid_build: buffers of rank %d are NULL", q);
    const GridPtrs gq = keygrid_pointers(world > 1 ? peers->grid[q] : grid, K, d);
    out.fine[q] = gq.fine; out.side[q] = gq.side;
    out.gflags[q] = world > 1 ? peers->gflags[q] : nullptr;
  }
  int64_t coarse0, n_coarse;
  if (cells != nullptr) {
    BDP_REQUIRE(n_cells >= 0 && n_cells <= g.n_coarse, "keygrid_build: %d cells of %lld", n_cells, (long long)g.n_coarse);
    coarse0 = (int64_t)rank * n_cells / world;
    n_coarse = (int64_t)(rank + 1) * n_cells / world - coarse0;
  } else {
    // slab of this rank: coarse layers [zc0, zc1) along the last axis
    const int zc0 = (int)((int64_t)rank * Gc / world), zc1 = (int)((int64_t)(rank + 1) * Gc / world);
    const int64_t layer_c = ipow64(Gc, d - 1);
    coarse0 = zc0 * layer_c; n_coarse = (zc1 - zc0) * layer_c;
  }
  const int cap = K < kCoarseListCap ? K : kCoarseListCap;
  const size_t smem = (size_t)cap * 16 + (size_t)(((cap + 1 + 7) & ~7) + (d == 3 ? 64 : 256) * kSideWidth) * 2;
  // sharded build: ~4 resident blocks per SM walk the slab (see the kernel); local build: one block per cell
  int64_t nb = n_coarse > 0 ? n_coarse : 1;
  if (world > 1 && nb > 4 * (int64_t)bdp_num_sms()) nb = 4 * (int64_t)bdp_num_sms();
  const unsigned blocks = (unsigned)nb;
  const bool active = n_coarse > 0;
  if (d == 3) {
    if (header_mode != 2)
      keygrid_header_kernel<3><<<1, 256, 0, st>>>(centers, K, G, margin_frac, g.hdr, g.cf32, stop, header_mode);
    if (active || world > 1)
      keygrid_cell_kernel<3><<<blocks, 256, smem, st>>>(g.cf32, K, g.hdr, out, coarse0, n_coarse, stop, cells);
  } else {
    if (header_mode != 2)
      keygrid_header_kernel<4><<<1, 256, 0, st>>>(centers, K, G, margin_frac, g.hdr, g.cf32, stop, header_mode);
    if (active || world > 1)
      keygrid_cell_kernel<4><<<blocks, 256, smem, st>>>(g.cf32, K, g.hdr, out, coarse0, n_coarse, stop, cells);
  }
  BDP_CUDA_CHECK_LAUNCH("keygrid kernels");
  return BDP_OK;
}

// header and fp32-key array of a grid buffer (kmeans.cu: the exchange kernel renews them)
void bdpi_keygrid_parts(void* grid, int K, int d, void** hdr, void** cf32) {
  const GridPtrs g = keygrid_pointers(grid, K, d);
  *hdr = g.hdr; *cf32 = g.cf32;
}

namespace {
// Fixed geometry over the caller's box (widened by 1e-4 of its largest extent); counters reset.
__global__ void keygrid_prepare_kernel(GridHdr* __restrict__ hdr, int d, int G, const double* __restrict__ lo,
                                       const double* __restrict__ hi) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double ext = 0.0;
  bool finite = true;
  for (int k = 0; k < d; ++k) {
    finite = finite && isfinite(lo[k]) && isfinite(hi[k]) && hi[k] >= lo[k];
    ext = fmax(ext, hi[k] - lo[k]);
  }
  const bool ok = finite && ext > 0.0 && isfinite(ext);
  const double margin = ext * 1e-4;
  for (int k = 0; k < 4; ++k) {
    double o = 0.0, c = 1.0;
    if (k < d && ok) {
      o = lo[k] - margin;
      c = (hi[k] - lo[k] + 2.0 * margin) / (double)G;
    }
    hdr->origin[k] = o; hdr->cell[k] = c; hdr->inv_cell[k] = 1.0 / c;
    hdr->origin32[k] = (float)o; hdr->inv_cell32[k] = (float)(1.0 / c);
  }
  hdr->G = G;
  hdr->enabled = ok ? 1 : 0;
  hdr->side_next = 0u; hdr->ticket = 0u;
}

// Coarse cells that hold at least one row: the point -> cell map is the query kernel's, expression
// for expression (fp32 FMA + floor on the fp32 copy of the coordinate), so a row marks exactly the
// coarse parent of the fine cell its query will read.
template <int D>
__global__ void __launch_bounds__(256) keygrid_occupancy_kernel(const double* __restrict__ x, int64_t N,
                                                                const GridHdr* __restrict__ hdr,
                                                                int* __restrict__ occ) {
  if (hdr->enabled == 0) return;
  float g_inv[D], g_off[D];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    g_inv[k] = (float)hdr->inv_cell[k];
    g_off[k] = (float)(-hdr->origin[k] * hdr->inv_cell[k]);
  }
  const int G = hdr->G, Gc = G / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    bool ok = true;
    int cidx = 0, mul = 1;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const float xf = (float)x[i * D + k];
      const float t = fmaf(xf, g_inv[k], g_off[k]);
      const int ck = __float2int_rd(t);
      ok = ok && ((unsigned)ck < (unsigned)G);
      cidx += (ck >> 2) * mul;
      mul *= Gc;
    }
    // every writer stores the same value; test first — ten million stores to a few thousand words
    // serialise in the L2 (350 us for the pass instead of 60)
    if (ok && __ldcg(occ + cidx) == 0) occ[cidx] = 1;
  }
}
}  // namespace

extern "C" int64_t bdp_keygrid_coarse_cells(int K, int d) {
  if ((d != 3 && d != 4) || K < 1 || K > kGridMaxK) return -1;
  return ipow64(keygrid_G(K, d) / 4, d);
}

extern "C" int bdp_keygrid_prepare(const double* box_lo, const double* box_hi, int K, int d, void* grid,
                                   int64_t grid_bytes, void* stream) {
  BDP_REQUIRE(box_lo && box_hi, "keygrid_prepare: NULL box");
  BDP_REQUIRE(d == 3 || d == 4, "keygrid_prepare: d must be 3 or 4 (got %d)", d);
  int rc = keygrid_check(grid, grid_bytes, K, d, "keygrid_prepare");
  if (rc != BDP_OK) return rc;
  const GridPtrs g = keygrid_pointers(grid, K, d);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // every cell starts as "overflowed": a query that reaches a cell no build has written scans the
  // dictionary (exact), so the occupancy list is an optimisation, never a correctness condition
  BDP_CUDA_CALL(cudaMemsetAsync(g.fine, 0xFF, (size_t)g.n_fine * 16, st));
  keygrid_prepare_kernel<<<1, 32, 0, st>>>(g.hdr, d, g.G, box_lo, box_hi);
  BDP_CUDA_CHECK_LAUNCH("keygrid_prepare_kernel");
  return BDP_OK;
}

extern "C" int bdp_keygrid_occupancy(const double* x, int64_t N, int d, int K, const void* grid,
                                     int64_t grid_bytes, int32_t* occ, void* stream) {
  BDP_REQUIRE(N >= 0 && occ, "keygrid_occupancy: bad arguments");
  BDP_REQUIRE(d == 3 || d == 4, "keygrid_occupancy: d must be 3 or 4 (got %d)", d);
  int rc = keygrid_check(grid, grid_bytes, K, d, "keygrid_occupancy");
  if (rc != BDP_OK) return rc;
  if (N == 0) return BDP_OK;
  BDP_REQUIRE(x != nullptr, "keygrid_occupancy: x is NULL");
  const GridPtrs g = keygrid_pointers(grid, K, d);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned blocks = (unsigned)(ceil_div64(N, 256) < 2368 ? ceil_div64(N, 256) : 2368);
  if (d == 3) keygrid_occupancy_kernel<3><<<blocks, 256, 0, st>>>(x, N, g.hdr, occ);
  else keygrid_occupancy_kernel<4><<<blocks, 256, 0, st>>>(x, N, g.hdr, occ);
  BDP_CUDA_CHECK_LAUNCH("keygrid_occupancy_kernel");
  return BDP_OK;
}

extern "C" int bdp_keygrid_build(const double* centers, int K, int d, void* grid, int64_t grid_bytes,
                                 void* stream) {
  return bdpi_keygrid_build(centers, K, d, grid, grid_bytes, nullptr, nullptr,
                            reinterpret_cast<cudaStream_t>(stream), nullptr, 0, 0);
}

extern "C" int bdp_keygrid_stats(const void* x, int x_dtype, int64_t N, int d, int K, const void* grid,
                                 int64_t grid_bytes, int64_t* stats, void* stream) {
  BDP_REQUIRE(N >= 0 && stats, "keygrid_stats: bad arguments");
  BDP_REQUIRE(d == 3 || d == 4, "keygrid_stats: d must be 3 or 4 (got %d)", d);
  BDP_REQUIRE(x_dtype == BDP_F32 || x_dtype == BDP_F64, "keygrid_stats: x_dtype %d", x_dtype);
  int rc = keygrid_check(grid, grid_bytes, K, d, "keygrid_stats");
  if (rc != BDP_OK) return rc;
  if (N == 0) return BDP_OK;
  const GridPtrs g = keygrid_pointers(grid, K, d);
  const GridHdr* hdr = g.hdr;
  const uint4* fine = g.fine;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned long long* s = reinterpret_cast<unsigned long long*>(stats);
  const unsigned blocks = (unsigned)(ceil_div64(N, 256) < 1184 ? ceil_div64(N, 256) : 1184);
  if (x_dtype == BDP_F32) {
    if (d == 3) keygrid_stats_kernel<float, 3><<<blocks, 256, 0, st>>>((const float*)x, N, hdr, fine, s);
    else keygrid_stats_kernel<float, 4><<<blocks, 256, 0, st>>>((const float*)x, N, hdr, fine, s);
  } else {
    if (d == 3) keygrid_stats_kernel<double, 3><<<blocks, 256, 0, st>>>((const double*)x, N, hdr, fine, s);
    else keygrid_stats_kernel<double, 4><<<blocks, 256, 0, st>>>((const double*)x, N, hdr, fine, s);
  }
  BDP_CUDA_CHECK_LAUNCH("keygrid_stats_kernel");
  return BDP_OK;
}

extern "C" int bdp_assign_nearest_grid(const void* x, int x_dtype, int64_t N, int d,
                                       const double* centers, int K, const void* grid,
                                       int64_t grid_bytes, int32_t* labels32, int64_t* labels64,
                                       float* residual, double* min_sqdist, void* stream) {
  BDP_REQUIRE(N >= 0, "assign_nearest_grid: N < 0");
  if (N == 0) return BDP_OK;
  BDP_REQUIRE(x && centers, "assign_nearest_grid: NULL input");
  BDP_REQUIRE(d == 3 || d == 4, "assign_nearest_grid: d must be 3 or 4 (got %d)", d);
  BDP_REQUIRE(x_dtype == BDP_F32 || x_dtype == BDP_F64, "assign_nearest_grid: x_dtype %d", x_dtype);
  int rc = keygrid_check(grid, grid_bytes, K, d, "assign_nearest_grid");
  if (rc != BDP_OK) return rc;
  AssignParams P = {};
  P.x = x; P.N = N; P.centers = centers; P.K = K;
  P.labels32 = labels32; P.labels64 = labels64; P.residual = residual; P.min_sqdist = min_sqdist;
  const GridPtrs g = keygrid_pointers(grid, K, d);
  P.ghdr = g.hdr; P.gfine = g.fine; P.gside = g.side;
  return dispatch_assign_grid<false>(P, x_dtype, d, reinterpret_cast<cudaStream_t>(stream));
}

int bdpi_lloyd_step_grid(const double* x, int64_t N, int d, const double* centers, int K,
                         const void* grid, int64_t grid_bytes, int32_t* labels, int64_t* acc,
                         int fix_hi_bits, int64_t* stats, double* inertia, int update,
                         const int* stop, const unsigned long long* gflags, int gworld,
                         unsigned long long gflag_value, int incremental, cudaStream_t st) {
  BDP_REQUIRE(N >= 0 && N < (1ll << 30), "kmeans_lloyd_step_grid: N out of range");
  if (N == 0) return BDP_OK;
  BDP_REQUIRE(x && centers && labels && stats, "kmeans_lloyd_step_grid: NULL buffer");
  BDP_REQUIRE(d == 3 || d == 4, "kmeans_lloyd_step_grid: d must be 3 or 4 (got %d)", d);
  BDP_REQUIRE(!update || acc, "kmeans_lloyd_step_grid: acc is NULL with update=1");
  BDP_REQUIRE(fix_hi_bits >= 0 && fix_hi_bits <= 30, "kmeans_lloyd_step_grid: fix_hi_bits %d",
              fix_hi_bits);
  int rc = keygrid_check(grid, grid_bytes, K, d, "kmeans_lloyd_step_grid");
  if (rc != BDP_OK) return rc;
  AssignParams P = {};
  P.x = x; P.N = N; P.centers = centers; P.K = K; P.labels32 = labels;
  P.acc = reinterpret_cast<unsigned long long*>(acc);
  P.scale_hi = ldexp(1.0, fix_hi_bits);
  P.stats = reinterpret_cast<unsigned long long*>(stats);
  P.inertia = inertia; P.update = update; P.stop = stop;
  P.gflags = gflags; P.gworld = gworld; P.gflag_value = gflag_value;
  P.incremental = incremental;
  const GridPtrs g = keygrid_pointers(grid, K, d);
  P.ghdr = g.hdr; P.gfine = g.fine; P.gside = g.side;
  return dispatch_assign_grid<true>(P, BDP_F64, d, st);
}

extern "C" int bdp_kmeans_lloyd_step_grid(const double* x, int64_t N, int d, const double* centers,
                                          int K, const void* grid, int64_t grid_bytes,
                                          int32_t* labels, int64_t* acc, int fix_hi_bits,
                                          int64_t* stats, double* inertia, int update,
                                          void* stream) {
  return bdpi_lloyd_step_grid(x, N, d, centers, K, grid, grid_bytes, labels, acc, fix_hi_bits, stats,
                              inertia, update, nullptr, nullptr, 0, 0ull, 0,
                              reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int bdp_euler_to_pose(const double* euler_deg, int64_t N, double* aa, double* quat,
                                 void* stream) {
  BDP_REQUIRE(N >= 0, "euler_to_pose: N < 0");
  if (N == 0) return BDP_OK;
  BDP_REQUIRE(euler_deg != nullptr && (aa || quat), "euler_to_pose: NULL buffer");
  euler_pose_kernel<<<(unsigned)ceil_div64(N, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      euler_deg, N, aa, quat);
  BDP_CUDA_CHECK_LAUNCH("euler_pose_kernel");
  return BDP_OK;
}

// ---- (c4) soft assignment -------------------------------------------------------------------------
// p[n,k] = exp(-gamma ||x_n - c_k||^2) / sum_j exp(-gamma ||x_n - c_j||^2), res[n] = x_n - sum_k p[n,k] c_k,
// all in fp64 like numpy/scipy in the reference (exp/normalise without max subtraction:
// binDeltaGenerators.py:104-108; dataGenerators.py:155-157, 166), stored as fp32 (`.float()`).
// One warp per row, lanes over the keys: the [N,K] output goes out as whole 128-byte lines, the
// weights are recomputed in the second pass instead of being parked (K can be 4096).
namespace {

template <typename T, int D>
__global__ void __launch_bounds__(256) soft_assign_kernel(const T* __restrict__ x,
                                                          const double* __restrict__ centers,
                                                          int64_t N, int K, double gamma, int c_in_smem,
                                                          float* __restrict__ p, float* __restrict__ res) {
  extern __shared__ __align__(16) unsigned char soft_smem[];
  double* s_c = reinterpret_cast<double*>(soft_smem);
  if (c_in_smem) {
    for (int i = threadIdx.x; i < K * D; i += blockDim.x) s_c[i] = __ldg(centers + i);
    __syncthreads();
  }
  const double* c = c_in_smem ? s_c : centers;
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < N; row += warps) {
    double y[D];
#pragma unroll
    for (int k = 0; k < D; ++k) y[k] = (double)x[row * D + k];
    double sum = 0.0;
    for (int j = lane; j < K; j += 32) {
      double d2 = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const double df = y[k] - c[(int64_t)j * D + k];
        d2 += df * df;
      }
      sum += exp(-gamma * d2);
    }
    sum = warp_sum(sum);                                // butterfly: the same value in every lane
    double acc[D];
#pragma unroll
    for (int k = 0; k < D; ++k) acc[k] = 0.0;
    for (int j = lane; j < K; j += 32) {
      double d2 = 0.0;
      double cj[D];
#pragma unroll
      for (int k = 0; k < D; ++k) {
        cj[k] = c[(int64_t)j * D + k];
        const double df = y[k] - cj[k];
        d2 += df * df;
      }
      const double pk = exp(-gamma * d2) / sum;
      if (p) p[row * K + j] = (float)pk;
#pragma unroll
      for (int k = 0; k < D; ++k) acc[k] += pk * cj[k];
    }
#pragma unroll
    for (int k = 0; k < D; ++k) acc[k] = warp_sum(acc[k]);
    if (res && lane == 0) {
#pragma unroll
      for (int k = 0; k < D; ++k) res[row * D + k] = (float)(y[k] - acc[k]);
    }
  }
}

template <typename T, int D>
int launch_soft_assign(const void* x, const double* centers, int64_t N, int K, double gamma, float* p,
                       float* res, cudaStream_t st) {
  int64_t blocks = ceil_div64(N, 8);                    // 8 warps (rows) per block
  const int64_t cap = (int64_t)bdp_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  const size_t bytes = (size_t)K * D * sizeof(double);
  const int in_smem = bytes <= 48 * 1024 ? 1 : 0;       // larger dictionaries are read through L1/L2
  soft_assign_kernel<T, D><<<(unsigned)blocks, 256, in_smem ? bytes : 0, st>>>(
      reinterpret_cast<const T*>(x), centers, N, K, gamma, in_smem, p, res);
  BDP_CUDA_CHECK_LAUNCH("soft_assign_kernel");
  return BDP_OK;
}

}  // namespace

extern "C" int bdp_assign_soft(const void* x, int x_dtype, int64_t N, int d, const double* centers, int K,
                               double gamma, float* p, float* residual, void* stream) {
  BDP_REQUIRE(N >= 0, "assign_soft: N < 0");
  if (N == 0) return BDP_OK;
  BDP_REQUIRE(x && centers, "assign_soft: NULL input");
  BDP_REQUIRE(d == 3 || d == 4, "assign_soft: d must be 3 or 4 (got %d)", d);
  BDP_REQUIRE(K >= 1 && K <= 65536, "assign_soft: K out of range (%d)", K);
  BDP_REQUIRE(x_dtype == BDP_F32 || x_dtype == BDP_F64, "assign_soft: x_dtype %d", x_dtype);
  BDP_REQUIRE(p || residual, "assign_soft: no output requested");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (x_dtype == BDP_F32)
    return d == 3 ? launch_soft_assign<float, 3>(x, centers, N, K, gamma, p, residual, st)
                  : launch_soft_assign<float, 4>(x, centers, N, K, gamma, p, residual, st);
  return d == 3 ? launch_soft_assign<double, 3>(x, centers, N, K, gamma, p, residual, st)
                : launch_soft_assign<double, 4>(x, centers, N, K, gamma, p, residual, st);
}
