// Small ops around the hot path (SURVEY §8 rows d3, f2, f4 and the autograd glue of row b):
//   bdp_compose_prediction  test-time pose composition  dict[argmax(score)] (+) residual
//                           (learnGeodesicBDModel.py:217-219, _quaternion.py:217-218,
//                           learnRiemannianBDModel.py:247)
//   bdp_min_key_gap         min_{i != j} ||k_i - k_j||^2   (helperFunctions.get_gamma, 51-58)
//   bdp_sgd_step            multi-tensor step of the cyclical-LR SGD (helperFunctions.mySGD, 62-120)
//   bdp_scale_inplace       buf *= *g with an early exit when *g == 1 (upstream scalar of a loss)
#include "common.cuh"

namespace {

// ---- Rodrigues exp / log in fp64 (axisAngle.get_R 33-41, get_y 19-29) --------------------------
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

// One warp per prediction: argmax over the K scores (first maximum, np.argmax), then lane 0 composes.
__global__ void __launch_bounds__(256) compose_kernel(const float* __restrict__ score, int64_t N, int K,
                                                      int64_t ld, const float* __restrict__ res,
                                                      int ndim, const double* __restrict__ dict,
                                                      int mode, double* __restrict__ out,
                                                      int64_t* __restrict__ bin_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < N; row += warps) {
    const float* s = score + row * ld;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int j = lane; j < K; j += 32) {
      const float v = __ldcs(s + j);
      if (v > bv || (bi == 0x7fffffff && !(v < bv))) { bv = v; bi = j; }   // first max; NaN-free rows
    }
    warp_argmax(bv, bi);
    if (bi == 0x7fffffff) bi = 0;
    if (lane != 0) continue;
    if (bin_out) bin_out[row] = bi;
    if (mode == BDP_COMPOSE_RIEMANNIAN) {
      // get_y(R_key[bin] . get_R(res))
      const double v[3] = {(double)res[row * 3 + 0], (double)res[row * 3 + 1], (double)res[row * 3 + 2]};
      double R[9], M[9], y[3];
      aa_exp(v, R);
      const double* Kr = dict + (int64_t)bi * 9;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          M[r * 3 + c] = Kr[r * 3 + 0] * R[0 * 3 + c] + Kr[r * 3 + 1] * R[1 * 3 + c] + Kr[r * 3 + 2] * R[2 * 3 + c];
      rot_log(M, y);
      out[row * 3 + 0] = y[0]; out[row * 3 + 1] = y[1]; out[row * 3 + 2] = y[2];
    } else {
      double y[4], n2 = 0.0;
      for (int k = 0; k < ndim; ++k) {
        y[k] = dict[(int64_t)bi * ndim + k] + (double)res[row * ndim + k];
        n2 += y[k] * y[k];
      }
      const double inv = mode == BDP_COMPOSE_ADD_NORMALIZE ? 1.0 / fmax(sqrt(n2), 1e-10) : 1.0;
      for (int k = 0; k < ndim; ++k) out[row * ndim + k] = mode == BDP_COMPOSE_ADD_NORMALIZE ? y[k] * inv : y[k];
    }
  }
}

// min over i != j of ||k_i - k_j||^2 (no fma contraction: the reference's cdist adds plain products)
__global__ void __launch_bounds__(256) min_gap_kernel(const double* __restrict__ keys, int K, int d,
                                                      double* __restrict__ out) {
  __shared__ double s_min[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double best = INFINITY;
  for (int i = warp; i < K; i += 8) {
    for (int j = lane; j < K; j += 32) {
      if (j == i) continue;
      double sq = 0.0;
      for (int k = 0; k < d; ++k) {
        const double df = __dsub_rn(keys[(int64_t)i * d + k], keys[(int64_t)j * d + k]);
        sq = __dadd_rn(sq, __dmul_rn(df, df));
      }
      best = fmin(best, sq);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_xor_sync(BDP_FULL_MASK, best, o));
  if (lane == 0) s_min[warp] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) best = fmin(best, s_min[w]);
    *out = best;
  }
}

struct SgdTensor {      // one row of the device table of bdp_sgd_step (mirrors bdp_sgd_tensor)
  float* p;
  const float* g;
  float* buf;
  int64_t n;
  float step_size;
  int first;            // no momentum buffer yet: buf = d_p
};

__global__ void __launch_bounds__(256) sgd_kernel(const SgdTensor* __restrict__ table, float weight_decay,
                                                  float momentum, float dampening, int nesterov) {
  const SgdTensor t = table[blockIdx.y];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t.n; i += stride) {
    float p = t.p[i];
    float d_p = t.g[i];
    if (weight_decay != 0.f) d_p = fmaf(weight_decay, p, d_p);
    if (momentum != 0.f) {
      float b = t.first ? d_p : fmaf(1.f - dampening, d_p, momentum * t.buf[i]);
      t.buf[i] = b;
      d_p = nesterov ? fmaf(momentum, b, d_p) : b;
    }
    t.p[i] = fmaf(-t.step_size, d_p, p);
  }
}

__global__ void __launch_bounds__(256) scale_kernel(float* __restrict__ buf, int64_t n,
                                                    const float* __restrict__ g) {
  const float s = *g;
  if (s == 1.0f) return;                  // the usual upstream gradient of a loss: nothing to move
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n >> 2;
  float4* b4 = reinterpret_cast<float4*>(buf);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = b4[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    b4[i] = v;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) buf[i] *= s;
}

}  // namespace

extern "C" int bdp_compose_prediction(const float* score, int64_t N, int K, int64_t ld_score,
                                      const float* residual, int ndim, const double* dict, int mode,
                                      double* out, int64_t* bin_out, void* stream) {
  BDP_REQUIRE(N >= 0, "compose_prediction: N < 0");
  if (N == 0) return BDP_OK;
  BDP_REQUIRE(score && residual && dict && out, "compose_prediction: NULL buffer");
  BDP_REQUIRE(K >= 1 && ld_score >= K, "compose_prediction: K %d, ld %lld", K, (long long)ld_score);
  BDP_REQUIRE(mode == BDP_COMPOSE_ADD || mode == BDP_COMPOSE_ADD_NORMALIZE || mode == BDP_COMPOSE_RIEMANNIAN,
              "compose_prediction: mode %d", mode);
  BDP_REQUIRE(mode == BDP_COMPOSE_RIEMANNIAN ? ndim == 3 : (ndim == 3 || ndim == 4),
              "compose_prediction: ndim %d for mode %d", ndim, mode);
  int64_t blocks = ceil_div64(N, 8);
  const int64_t cap = (int64_t)bdp_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  compose_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      score, N, K, ld_score, residual, ndim, dict, mode, out, bin_out);
  BDP_CUDA_CHECK_LAUNCH("compose_kernel");
  return BDP_OK;
}

extern "C" int bdp_min_key_gap(const double* keys, int K, int d, double* out, void* stream) {
  BDP_REQUIRE(keys && out, "min_key_gap: NULL buffer");
  BDP_REQUIRE(K >= 2 && d >= 1 && d <= 16, "min_key_gap: K %d, d %d", K, d);
  min_gap_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(keys, K, d, out);
  BDP_CUDA_CHECK_LAUNCH("min_gap_kernel");
  return BDP_OK;
}

static_assert(sizeof(SgdTensor) == sizeof(bdp_sgd_tensor), "bdp_sgd_tensor layout");

extern "C" int bdp_sgd_step(const bdp_sgd_tensor* table_dev, int n_tensors, int64_t max_numel,
                            float weight_decay, float momentum, float dampening, int nesterov,
                            void* stream) {
  BDP_REQUIRE(n_tensors >= 0 && n_tensors <= 65535, "sgd_step: %d tensors", n_tensors);
  if (n_tensors == 0) return BDP_OK;
  BDP_REQUIRE(table_dev != nullptr && max_numel >= 0, "sgd_step: NULL table");
  int64_t bx = ceil_div64(max_numel, 256 * 4);
  const int64_t cap = (int64_t)bdp_num_sms() * 4;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)n_tensors);
  sgd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const SgdTensor*>(table_dev), weight_decay, momentum, dampening, nesterov);
  BDP_CUDA_CHECK_LAUNCH("sgd_kernel");
  return BDP_OK;
}

extern "C" int bdp_scale_inplace(float* buf, int64_t n, const float* scale_dev, void* stream) {
  BDP_REQUIRE(n >= 0, "scale_inplace: n < 0");
  if (n == 0) return BDP_OK;
  BDP_REQUIRE(buf && scale_dev, "scale_inplace: NULL buffer");
  BDP_REQUIRE((reinterpret_cast<uintptr_t>(buf) & 15) == 0, "scale_inplace: buffer must be 16-byte aligned");
  int64_t blocks = ceil_div64(n, 256 * 16);
  const int64_t cap = (int64_t)bdp_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  scale_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(buf, n, scale_dev);
  BDP_CUDA_CHECK_LAUNCH("scale_kernel");
  return BDP_OK;
}

// ---- fit preprocessing (scikit-learn KMeans.fit: X -= X.mean(0); tol = mean(var(X)) * tol) -------------
// Column statistics in fixed point, so that the mean / variance — and with them the centred data and
// every label — do not depend on how the rows are split over blocks or GPUs.
namespace {

// mode 0: stats[0] = max |x| (as the bit pattern of a non-negative double, atomicMax)
// mode 1: limbs[2][d] += two-limb fixed-point column sums of x * 2^hi_bits
// mode 2: y = x - mean (written), stats[0] = max (x-mean)^2, stats[1] = max |x - mean|
// mode 3: q[d] += round((x)^2 * 2^sh) column sums (x already centred)
template <int MODE>
__global__ void __launch_bounds__(256) fit_stats_kernel(const double* __restrict__ x, int64_t N, int d,
                                                        const double* __restrict__ mean, double scale,
                                                        double* __restrict__ y,
                                                        unsigned long long* __restrict__ out) {
  __shared__ unsigned long long s_a[8][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t total = N * d;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // every thread keeps its column fixed: the stride is a multiple of d
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double m0 = 0.0, m1 = 0.0;
  long long a_hi = 0, a_lo = 0;
  const int col = (int)(i0 % d);
  const double mu = (MODE == 2) ? mean[col] : 0.0;
  for (int64_t i = i0; i < total; i += stride) {
    const double v = x[i];
    if (MODE == 0) {
      m0 = fmax(m0, fabs(v));
    } else if (MODE == 1) {
      const double xs = v * scale;
      const double f = floor(xs);
      a_hi += (long long)f;
      a_lo += (long long)((xs - f) * 4294967296.0);
    } else if (MODE == 2) {
      const double c = v - mu;
      y[i] = c;
      m0 = fmax(m0, c * c);
      m1 = fmax(m1, fabs(c));
    } else {
      a_hi += llrint(v * v * scale);
    }
  }
  if (MODE == 0 || MODE == 2) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m0 = fmax(m0, __shfl_xor_sync(BDP_FULL_MASK, m0, o));
      m1 = fmax(m1, __shfl_xor_sync(BDP_FULL_MASK, m1, o));
    }
    if (lane == 0) {
      atomicMax(out + 0, (unsigned long long)__double_as_longlong(m0));
      if (MODE == 2) atomicMax(out + 1, (unsigned long long)__double_as_longlong(m1));
    }
  } else {
    // per-column sums: lanes of a warp hold different columns (lane % d pattern differs per warp), so
    // reduce through shared memory by column
    if (threadIdx.x < 64) s_a[threadIdx.x >> 3][threadIdx.x & 7] = 0ull;
    __syncthreads();
    atomicAdd(&s_a[0][col], (unsigned long long)a_hi);
    if (MODE == 1) atomicAdd(&s_a[1][col], (unsigned long long)a_lo);
    __syncthreads();
    if (threadIdx.x < d) {
      atomicAdd(out + threadIdx.x, s_a[0][threadIdx.x]);
      if (MODE == 1) atomicAdd(out + d + threadIdx.x, s_a[1][threadIdx.x]);
    }
    (void)warp;
  }
}

}  // namespace

extern "C" int bdp_fit_stats(const double* x, int64_t N, int d, int mode, const double* mean,
                             double scale, double* y, void* out, void* stream) {
  BDP_REQUIRE(N >= 0 && d >= 1 && d <= 8, "fit_stats: N %lld, d %d", (long long)N, d);
  if (N == 0) return BDP_OK;
  BDP_REQUIRE(x && out, "fit_stats: NULL buffer");
  BDP_REQUIRE(mode >= 0 && mode <= 3, "fit_stats: mode %d", mode);
  BDP_REQUIRE(mode != 2 || (mean && y), "fit_stats: mode 2 needs mean and y");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // grid stride = blocks * 256 must be a multiple of d so that a thread stays in one column
  int64_t blocks = (int64_t)bdp_num_sms() * 8;
  blocks -= blocks % d;
  const int64_t need = ceil_div64(N * d, 256);
  if (blocks > need) blocks = need - need % d;
  if (blocks < d) blocks = d;
  unsigned long long* o = reinterpret_cast<unsigned long long*>(out);
  switch (mode) {
    case 0: fit_stats_kernel<0><<<(unsigned)blocks, 256, 0, st>>>(x, N, d, mean, scale, y, o); break;
    case 1: fit_stats_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(x, N, d, mean, scale, y, o); break;
    case 2: fit_stats_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(x, N, d, mean, scale, y, o); break;
    default: fit_stats_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(x, N, d, mean, scale, y, o); break;
  }
  BDP_CUDA_CHECK_LAUNCH("fit_stats_kernel");
  return BDP_OK;
}
