// (d) Evaluation reductions for sm_100a: batched geodesic error in degrees, Acc@30deg count,
// max, and np.median-exact (per-class) medians by MSB-first radix select on the fp64 bit patterns.
//
// Reference: axisAngle.get_error / get_error2 (axisAngle.py:45-95) and quaternion.get_error /
// get_error2 (quaternion.py:33-76) — per-sample Python loops on the host.
//
// Error kernel: one thread per pair.  The arithmetic restates the reference line by line in fp64
// (inputs are widened first): R = I + sin(t) V + (1 - cos t) V V with V the skew of v/|v| (identity
// when |v| < 1e-6), tr(R1^T R2) = sum_ij R1_ij R2_ij, acos(clip((tr - 1)/2, -1, 1)), rad2deg.  Keeping
// the reference's own (acos-of-trace) form, rather than a better conditioned atan2 form, keeps the
// per-sample errors within rounding noise of what the reference prints, including its exact 0 for
// identical inputs whose trace rounds above 3.
#include "common.cuh"

namespace {

constexpr double kRad2Deg = 57.295779513082320876798154814105;

template <typename T>
__device__ __forceinline__ void aa_to_rot(const T* v, double R[9]) {
  const double x = (double)v[0], y = (double)v[1], z = (double)v[2];
  const double th = sqrt(x * x + y * y + z * z);
  if (th < 1e-6) {            // axisAngle.get_R: identity when theta < eps (axisAngle.py:35)
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
    return;
  }
  const double a[3] = {x / th, y / th, z / th};
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

template <typename T, int REPR>
__global__ void __launch_bounds__(256) geodesic_error_kernel(const T* __restrict__ a,
                                                            const T* __restrict__ b, int64_t N,
                                                            double* __restrict__ err) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    double e;
    if (REPR == BDP_REPR_AXIS_ANGLE) {
      T va[3], vb[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) { va[k] = a[i * 3 + k]; vb[k] = b[i * 3 + k]; }
      double R1[9], R2[9];
      aa_to_rot(va, R1);
      aa_to_rot(vb, R2);
      // trace(R1^T R2): diagonal entry j of the product is sum_i R1[i][j] * R2[i][j]
      double tr = 0.0;
#pragma unroll
      for (int j = 0; j < 3; ++j)
        tr += R1[0 * 3 + j] * R2[0 * 3 + j] + R1[1 * 3 + j] * R2[1 * 3 + j] +
              R1[2 * 3 + j] * R2[2 * 3 + j];
      e = fabs(acos(fmin(fmax(0.5 * (tr - 1.0), -1.0), 1.0)));   // axisAngle.py:58-59
    } else {
      // quaternion.get_error: 2*acos(|clip(c1*c2 + sum(v1*v2), -1, 1)|)   (quaternion.py:43-45)
      const double c = (double)a[i * 4 + 0] * (double)b[i * 4 + 0];
      const double v = ((double)a[i * 4 + 1] * (double)b[i * 4 + 1] +
                        (double)a[i * 4 + 2] * (double)b[i * 4 + 2]) +
                       (double)a[i * 4 + 3] * (double)b[i * 4 + 3];
      const double d = fmin(fmax(c + v, -1.0), 1.0);
      e = 2.0 * acos(fabs(d));
    }
    err[i] = e * kRad2Deg;
  }
}

// ---- statistics -----------------------------------------------------------------------------
constexpr int kSelPerChunk = 24;      // selectors (2 per class) whose histograms share one CTA
constexpr int kDigitBits = 8;
constexpr int kBins = 1 << kDigitBits;

struct StatsWs {
  unsigned long long* prefix;   // [S] bits decided so far (high bits), low bits zero
  long long* krem;              // [S] remaining rank inside the current prefix bucket; <0 = empty
  unsigned int* hist;           // [S][kBins]
  long long* count;             // [C]
};

__global__ void __launch_bounds__(256) stats_count_kernel(const double* __restrict__ err,
                                                          const int64_t* __restrict__ labels,
                                                          int64_t N, int C, long long* count,
                                                          long long* below30,
                                                          unsigned long long* max_bits) {
  extern __shared__ unsigned int s_cnt[];   // [C]
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_cnt[c] = 0;
  __syncthreads();
  int nb30 = 0;
  unsigned long long mx = 0ull;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    const double e = err[i];
    nb30 += (e < 30.0);
    const unsigned long long bits = (unsigned long long)__double_as_longlong(e);
    mx = bits > mx ? bits : mx;        // err >= 0: the bit pattern is order preserving
    const int c = labels ? (int)labels[i] : 0;
    if (c >= 0 && c < C) atomicAdd(&s_cnt[c], 1u);
  }
  nb30 = warp_sum(nb30);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(BDP_FULL_MASK, mx, o);
    mx = other > mx ? other : mx;
  }
  if ((threadIdx.x & 31) == 0) {
    if (nb30) atomicAdd((unsigned long long*)below30, (unsigned long long)nb30);
    atomicMax(max_bits, mx);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    if (s_cnt[c]) atomicAdd((unsigned long long*)&count[c], (unsigned long long)s_cnt[c]);
}

__global__ void stats_init_selectors(StatsWs ws, int C) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= 2 * C) return;
  const long long n = ws.count[s >> 1];
  ws.prefix[s] = 0ull;
  // np.median: mean of order statistics (n-1)/2 and n/2 (0-based)
  ws.krem[s] = n > 0 ? ((s & 1) ? n / 2 : (n - 1) / 2) : -1;
}

// One radix pass: histogram the digit at `shift` of every key that matches its selector's prefix.
__global__ void __launch_bounds__(256) stats_hist_kernel(const double* __restrict__ err,
                                                         const int64_t* __restrict__ labels,
                                                         int64_t N, int C, int shift, StatsWs ws) {
  __shared__ unsigned int s_hist[kSelPerChunk][kBins];
  __shared__ unsigned long long s_prefix[kSelPerChunk];
  __shared__ int s_live[kSelPerChunk];
  const int sel0 = blockIdx.y * kSelPerChunk;
  const int nsel = min(kSelPerChunk, 2 * C - sel0);
  for (int i = threadIdx.x; i < kSelPerChunk * kBins; i += blockDim.x) (&s_hist[0][0])[i] = 0;
  if (threadIdx.x < kSelPerChunk) {
    const bool ok = threadIdx.x < nsel;
    s_prefix[threadIdx.x] = ok ? ws.prefix[sel0 + threadIdx.x] : 0ull;
    s_live[threadIdx.x] = ok && ws.krem[sel0 + threadIdx.x] >= 0;
  }
  __syncthreads();
  const unsigned long long hi_mask = (shift + kDigitBits >= 64) ? 0ull : (~0ull << (shift + kDigitBits));
  const int c0 = sel0 >> 1, c1 = (sel0 + nsel + 1) >> 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    const int c = labels ? (int)labels[i] : 0;
    if (c < c0 || c >= c1 || c >= C) continue;
    const unsigned long long key = (unsigned long long)__double_as_longlong(err[i]);
    const int digit = (int)((key >> shift) & (kBins - 1));
    const unsigned long long hi = key & hi_mask;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int ls = 2 * c + h - sel0;
      if (ls >= 0 && ls < nsel && s_live[ls] && hi == s_prefix[ls]) atomicAdd(&s_hist[ls][digit], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nsel * kBins; i += blockDim.x) {
    const unsigned int v = (&s_hist[0][0])[i];
    if (v) atomicAdd(&ws.hist[(size_t)sel0 * kBins + i], v);
  }
}

// Per selector: walk the 256 bins in order, pick the digit holding rank krem, zero the histogram.
__global__ void __launch_bounds__(kBins) stats_scan_kernel(StatsWs ws, int shift) {
  const int s = blockIdx.x;
  __shared__ unsigned int s_h[kBins];
  unsigned int* h = ws.hist + (size_t)s * kBins;
  s_h[threadIdx.x] = h[threadIdx.x];
  h[threadIdx.x] = 0u;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long k = ws.krem[s];
    if (k >= 0) {
      int d = 0;
      for (; d < kBins - 1; ++d) {
        if (k < (long long)s_h[d]) break;
        k -= s_h[d];
      }
      ws.krem[s] = k;
      ws.prefix[s] |= ((unsigned long long)d << shift);
    }
  }
}

__global__ void stats_finish_kernel(StatsWs ws, int C, double* median, int64_t* count_out,
                                    const unsigned long long* max_bits, double* max_err) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && max_err) *max_err = __longlong_as_double((long long)*max_bits);
  if (c >= C) return;
  const long long n = ws.count[c];
  if (count_out) count_out[c] = n;
  if (n <= 0) {
    median[c] = __longlong_as_double(0x7ff8000000000000LL);   // NaN: np.median of an empty slice
  } else {
    const double lo = __longlong_as_double((long long)ws.prefix[2 * c]);
    const double hi = __longlong_as_double((long long)ws.prefix[2 * c + 1]);
    median[c] = (lo + hi) / 2.0;
  }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct StatsLayout {
  size_t prefix, krem, hist, count, below30, maxbits, total;
};
StatsLayout stats_layout(int C) {
  StatsLayout L;
  const size_t S = 2 * (size_t)C;
  size_t o = 0;
  L.prefix = o; o = align_up(o + S * 8, 256);
  L.krem = o; o = align_up(o + S * 8, 256);
  L.hist = o; o = align_up(o + S * kBins * 4, 256);
  L.count = o; o = align_up(o + (size_t)C * 8, 256);
  L.below30 = o; o += 8;
  L.maxbits = o; o = align_up(o + 8, 256);
  L.total = o;
  return L;
}

}  // namespace

extern "C" int bdp_geodesic_error_deg(const void* y_gt, const void* y_hat, int dtype, int repr,
                                      int64_t N, double* err_deg, void* stream) {
  BDP_REQUIRE(N >= 0, "geodesic_error: N < 0");
  if (N == 0) return BDP_OK;
  BDP_REQUIRE(y_gt && y_hat && err_deg, "geodesic_error: NULL buffer");
  BDP_REQUIRE(dtype == BDP_F32 || dtype == BDP_F64, "geodesic_error: dtype %d", dtype);
  BDP_REQUIRE(repr == BDP_REPR_AXIS_ANGLE || repr == BDP_REPR_QUATERNION, "geodesic_error: repr %d",
              repr);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int threads = 256;
  int64_t blocks = ceil_div64(N, threads);
  const int64_t cap = (int64_t)bdp_num_sms() * 16;
  if (blocks > cap) blocks = cap;
#define LAUNCH(T, R)                                                                     \
  geodesic_error_kernel<T, R><<<(unsigned)blocks, threads, 0, st>>>(                      \
      reinterpret_cast<const T*>(y_gt), reinterpret_cast<const T*>(y_hat), N, err_deg)
  if (dtype == BDP_F32) {
    if (repr == BDP_REPR_AXIS_ANGLE) LAUNCH(float, BDP_REPR_AXIS_ANGLE);
    else LAUNCH(float, BDP_REPR_QUATERNION);
  } else {
    if (repr == BDP_REPR_AXIS_ANGLE) LAUNCH(double, BDP_REPR_AXIS_ANGLE);
    else LAUNCH(double, BDP_REPR_QUATERNION);
  }
#undef LAUNCH
  BDP_CUDA_CHECK_LAUNCH("geodesic_error_kernel");
  return BDP_OK;
}

extern "C" int64_t bdp_error_stats_workspace_bytes(int64_t /*N*/, int num_classes) {
  return (int64_t)stats_layout(num_classes < 1 ? 1 : num_classes).total;
}

extern "C" int bdp_error_stats(const double* err_deg, const int64_t* labels, int64_t N,
                               int num_classes, double* median, int64_t* count, int64_t* below30,
                               double* max_err, void* workspace, int64_t workspace_bytes,
                               void* stream) {
  BDP_REQUIRE(num_classes >= 1 && num_classes <= 4096, "error_stats: num_classes %d", num_classes);
  BDP_REQUIRE(N >= 0, "error_stats: N < 0");
  BDP_REQUIRE(median != nullptr, "error_stats: median is NULL");
  BDP_REQUIRE(N == 0 || err_deg != nullptr, "error_stats: err is NULL");
  const StatsLayout L = stats_layout(num_classes);
  BDP_REQUIRE(workspace && workspace_bytes >= (int64_t)L.total, "error_stats: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* base = reinterpret_cast<char*>(workspace);
  StatsWs ws;
  ws.prefix = reinterpret_cast<unsigned long long*>(base + L.prefix);
  ws.krem = reinterpret_cast<long long*>(base + L.krem);
  ws.hist = reinterpret_cast<unsigned int*>(base + L.hist);
  ws.count = reinterpret_cast<long long*>(base + L.count);
  long long* d_below = reinterpret_cast<long long*>(base + L.below30);
  unsigned long long* d_max = reinterpret_cast<unsigned long long*>(base + L.maxbits);
  BDP_CUDA_CALL(cudaMemsetAsync(workspace, 0, L.total, st));

  const int C = num_classes;
  const int threads = 256;
  int64_t blocks = N > 0 ? ceil_div64(N, threads * 4) : 1;
  const int64_t cap = (int64_t)bdp_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  stats_count_kernel<<<(unsigned)blocks, threads, C * sizeof(unsigned int), st>>>(
      err_deg, labels, N, C, ws.count, d_below, d_max);
  BDP_CUDA_CHECK_LAUNCH("stats_count_kernel");
  stats_init_selectors<<<(2 * C + 127) / 128, 128, 0, st>>>(ws, C);
  BDP_CUDA_CHECK_LAUNCH("stats_init_selectors");
  const int chunks = (2 * C + kSelPerChunk - 1) / kSelPerChunk;
  int64_t hblocks = blocks;
  if (hblocks * chunks > cap * 2) hblocks = (cap * 2 + chunks - 1) / chunks;
  if (hblocks < 1) hblocks = 1;
  for (int shift = 64 - kDigitBits; shift >= 0; shift -= kDigitBits) {
    stats_hist_kernel<<<dim3((unsigned)hblocks, chunks), threads, 0, st>>>(err_deg, labels, N, C,
                                                                          shift, ws);
    BDP_CUDA_CHECK_LAUNCH("stats_hist_kernel");
    stats_scan_kernel<<<2 * C, kBins, 0, st>>>(ws, shift);
    BDP_CUDA_CHECK_LAUNCH("stats_scan_kernel");
  }
  stats_finish_kernel<<<(C + 127) / 128, 128, 0, st>>>(ws, C, median, count, d_max, max_err);
  BDP_CUDA_CHECK_LAUNCH("stats_finish_kernel");
  if (below30)
    BDP_CUDA_CALL(cudaMemcpyAsync(below30, d_below, sizeof(long long), cudaMemcpyDeviceToDevice, st));
  return BDP_OK;
}
