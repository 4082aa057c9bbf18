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
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kPts = 4;                 // points per thread
constexpr int kChunk = 2048;            // dictionary records staged per shared-memory pass
constexpr int kMaxSmemAccK = 2048;      // largest K whose int64 accumulators live in shared memory

template <int D> struct CenterRec;      // fp32 screening record
template <> struct CenterRec<3> { float4 v; };                 // (-2c0,-2c1,-2c2,|c|^2)
template <> struct CenterRec<4> { float4 v; float n; };       // (-2c0..-2c3), |c|^2

template <int D>
__device__ __forceinline__ float screen_dist(const float* x, const float4& c, float cn) {
  if (D == 3) return fmaf(x[0], c.x, fmaf(x[1], c.y, fmaf(x[2], c.z, c.w)));
  return fmaf(x[0], c.x, fmaf(x[1], c.y, fmaf(x[2], c.z, fmaf(x[3], c.w, cn))));
}

// exact split of x*2^hi_bits into an integer part and 32 fractional bits (truncated below)
__device__ __forceinline__ void to_limbs(double x, double scale_hi, long long& hi, long long& lo) {
  const double xs = x * scale_hi;           // power-of-two scale: exact
  const double f = floor(xs);
  hi = (long long)f;
  lo = (long long)((xs - f) * 4294967296.0);  // (xs-f) in [0,1) exact; product exact; trunc
}

struct AssignParams {
  const void* x;
  int64_t N;
  const double* centers;   // [K, D] fp64
  int K;
  int32_t* labels32;       // out (assign) / in-out (lloyd)
  int64_t* labels64;
  float* residual;
  double* min_sqdist;
  // lloyd
  unsigned long long* acc; // [K, 2D+1]
  double scale_hi;
  unsigned long long* stats;
  double* inertia;
  int update;
  float err_coef;          // 2^-24 * 2(D+5) * safety
};

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


float screen_err_coef(int D) {
  // |fl32(dist) - dist| <= (D + 2.1) * 2^-24 * (||x|| + ||c||)^2 (input rounding of x, -2c, ||c||^2
  // plus one rounding per fma); the best/second GAP carries twice that.  x2 safety on top.
  return (float)(2.0 * 2.0 * (D + 5) * 5.9604644775390625e-08);
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

// ---- k-means M-step finalisation ------------------------------------------------------------------
// (hi, lo) limbs -> correctly rounded double of  hi*2^32 + lo  (|value| < 2^95), then * scale.
__device__ __forceinline__ double limbs_to_double(long long hi, long long lo, double inv_scale_lo) {
  // value = hi*2^32 + lo as a signed 128-bit integer (lo >= 0)
  __int128 v = ((__int128)hi << 32) + (__int128)lo;
  const bool neg = v < 0;
  unsigned __int128 m = neg ? (unsigned __int128)(-v) : (unsigned __int128)v;
  if (m == 0) return 0.0;
  // normalise to 64 significant bits + sticky, then let the u64 -> double conversion round once
  int shift = 0;
  unsigned long long top = (unsigned long long)(m >> 64);
  double r;
  if (top == 0) {
    const unsigned long long low = (unsigned long long)m;
    // u64 -> double is correctly rounded (round to nearest even) by the hardware conversion
    r = __ull2double_rn(low);
  } else {
    const int lz = __clzll((long long)top);
    shift = 64 - lz;                                  // bits to drop so that 64 remain
    unsigned long long kept = (unsigned long long)(m >> shift);
    const unsigned __int128 dropped = m & ((((unsigned __int128)1) << shift) - 1);
    if (dropped != 0) kept |= 1ull;                   // sticky: kept has 64 bits, 11 of them are
                                                      // below the double mantissa, so OR-ing the
                                                      // lowest one cannot change a tie wrongly
    r = __ull2double_rn(kept) * exp2((double)shift);
  }
  r *= inv_scale_lo;                                  // power of two: exact
  return neg ? -r : r;
}

template <int D>
__global__ void __launch_bounds__(256) kmeans_finalize_kernel(const long long* __restrict__ acc,
                                                              int K, double inv_scale_lo,
                                                              const double* __restrict__ c_old,
                                                              double* __restrict__ c_new,
                                                              double* __restrict__ shift2,
                                                              long long* __restrict__ n_empty) {
  // single block: K is a dictionary size (<= a few thousand)
  __shared__ long long s_best_cnt[8];
  __shared__ int s_best_idx[8];
  __shared__ double s_red[8];
  __shared__ int s_cnt_empty[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // pass 1: argmax count (first maximum, np.argmax) and number of empty clusters
  long long bc = -1;
  int bi = 0x7fffffff, ne = 0;
  for (int j = threadIdx.x; j < K; j += blockDim.x) {
    const long long c = acc[(size_t)j * (2 * D + 1) + 2 * D];
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
  // pass 2: centres.  sklearn _average_centers: alpha = 1/w; c *= alpha.  An empty cluster j copies
  // row `big` as it stands when the in-order loop reaches j: already averaged if big < j, still the
  // raw SUM if big > j (only reachable when relocation bailed out; kept for fidelity).
  double sh = 0.0;
  for (int j = threadIdx.x; j < K; j += blockDim.x) {
    const long long* a = acc + (size_t)j * (2 * D + 1);
    const long long cnt = a[2 * D];
    const long long* src = cnt > 0 ? a : acc + (size_t)big * (2 * D + 1);
    const long long scnt = src[2 * D];
    double alpha = 1.0;
    if (cnt > 0) alpha = 1.0 / (double)cnt;
    else if (big < j && scnt > 0) alpha = 1.0 / (double)scnt;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const double sum = limbs_to_double(src[2 * k], src[2 * k + 1], inv_scale_lo);
      const double c = sum * alpha;
      c_new[(size_t)j * D + k] = c;
      const double df = c - c_old[(size_t)j * D + k];
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
    if (shift2) *shift2 = t;
    if (n_empty) *n_empty = ne;
  }
}

}  // namespace

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

extern "C" int bdp_kmeans_lloyd_step(const double* x, int64_t N, int d, const double* centers,
                                     int K, int32_t* labels, int64_t* acc, int fix_hi_bits,
                                     int64_t* stats, double* inertia, int update, void* stream) {
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
  P.inertia = inertia; P.update = update;
  return dispatch_assign<true>(P, BDP_F64, d, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int bdp_kmeans_finalize(const int64_t* acc, int K, int d, int fix_hi_bits,
                                   const double* centers_old, double* centers_new, double* shift2,
                                   int64_t* n_empty, void* stream) {
  BDP_REQUIRE(acc && centers_old && centers_new, "kmeans_finalize: NULL buffer");
  BDP_REQUIRE(d == 3 || d == 4, "kmeans_finalize: d must be 3 or 4");
  BDP_REQUIRE(K >= 1, "kmeans_finalize: K must be >= 1");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const double inv = ldexp(1.0, -(fix_hi_bits + 32));
  if (d == 3)
    kmeans_finalize_kernel<3><<<1, 256, 0, st>>>(reinterpret_cast<const long long*>(acc), K, inv,
                                                 centers_old, centers_new, shift2,
                                                 reinterpret_cast<long long*>(n_empty));
  else
    kmeans_finalize_kernel<4><<<1, 256, 0, st>>>(reinterpret_cast<const long long*>(acc), K, inv,
                                                 centers_old, centers_new, shift2,
                                                 reinterpret_cast<long long*>(n_empty));
  BDP_CUDA_CHECK_LAUNCH("kmeans_finalize_kernel");
  return BDP_OK;
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
