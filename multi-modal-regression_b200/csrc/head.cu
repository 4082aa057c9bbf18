// (a) The CUDA-core pieces around the tcgen05 GEMMs of the bin-delta heads: BatchNorm1d+ReLU
// forward/backward, the label-selected / soft-mixed fc3, and the split-K slab sum.  Activations are
// batch-major [B, F] (the torch layout; the GEMM puts the batch on the TMEM lanes and streams the
// weights on the N side): a BatchNorm block owns 32 adjacent features (coalesced 128-byte rows) and
// its 8 warps split the batch.
//
// Reference: binDeltaModels.py:62-91 (layers), 112-121 (stack + one-hot bmm select),
// learnJointCatPoseModel_weighted.py:107-115 (softmax mixing); nn.BatchNorm1d defaults.
#include "common.cuh"

namespace {

constexpr int kBnWarps = 8;

// sum over the 8 warps of a block of per-thread partials; result broadcast to every warp
__device__ __forceinline__ double block_col_sum(double v, double (*s)[32], int wy, int lane) {
  __syncthreads();
  s[wy][lane] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kBnWarps; ++w) t += s[w][lane];
  return t;
}

// ---- BatchNorm + ReLU ----------------------------------------------------------------------------
__global__ void __launch_bounds__(kBnWarps * 32)
bn_relu_fwd_kernel(const float* __restrict__ h, int64_t F, int B, int64_t ld,
                   const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ running_mean, float* __restrict__ running_var,
                   float* __restrict__ save_mean, float* __restrict__ save_invstd, float eps,
                   float momentum, int training, float* __restrict__ a) {
  __shared__ double s_red[kBnWarps][32];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int64_t f = (int64_t)blockIdx.x * 32 + lane;
  const bool ok = f < F;
  double mean = 0.0, invstd = 1.0;
  if (training) {
    // statistics in double (what torch's CPU kernel accumulates in; its CUDA kernel's Welford in
    // float agrees to ~1e-7)
    double sacc = 0.0;
    if (ok) for (int b = wy; b < B; b += kBnWarps) sacc += (double)h[(int64_t)b * ld + f];
    mean = block_col_sum(sacc, s_red, wy, lane) / (double)B;
    double vacc = 0.0;
    if (ok) for (int b = wy; b < B; b += kBnWarps) {
      const double d = (double)h[(int64_t)b * ld + f] - mean;
      vacc += d * d;
    }
    const double v = block_col_sum(vacc, s_red, wy, lane);
    const double var = v / (double)B;
    invstd = 1.0 / sqrt(var + (double)eps);
    if (ok && wy == 0) {
      if (save_mean) save_mean[f] = (float)mean;
      if (save_invstd) save_invstd[f] = (float)invstd;
      if (running_mean) running_mean[f] = (1.f - momentum) * running_mean[f] + momentum * (float)mean;
      if (running_var) {
        const double unbiased = B > 1 ? v / (double)(B - 1) : var;
        running_var[f] = (1.f - momentum) * running_var[f] + momentum * (float)unbiased;
      }
    }
  } else if (ok) {
    mean = (double)running_mean[f];
    invstd = 1.0 / sqrt((double)running_var[f] + (double)eps);
    if (wy == 0) {                       // the eval-mode backward reads them like batch statistics
      if (save_mean) save_mean[f] = (float)mean;
      if (save_invstd) save_invstd[f] = (float)invstd;
    }
  }
  if (!ok) return;
  const float m = (float)mean, is = (float)invstd, g = gamma[f], bt = beta[f];
  for (int b = wy; b < B; b += kBnWarps) {
    const int64_t i = (int64_t)b * ld + f;
    a[i] = fmaxf((h[i] - m) * is * g + bt, 0.f);
  }
}

__global__ void __launch_bounds__(kBnWarps * 32)
bn_relu_bwd_kernel(const float* __restrict__ da, const float* __restrict__ a,
                   const float* __restrict__ h, const float* __restrict__ gamma,
                   const float* __restrict__ save_mean, const float* __restrict__ save_invstd,
                   int64_t F, int B, int64_t ld, int training, float* __restrict__ dh,
                   float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ double s_red[kBnWarps][32];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int64_t f = (int64_t)blockIdx.x * 32 + lane;
  const bool ok = f < F;
  const float m = ok ? save_mean[f] : 0.f, is = ok ? save_invstd[f] : 1.f, g = ok ? gamma[f] : 0.f;
  double a1 = 0.0, a2 = 0.0;
  if (ok) for (int b = wy; b < B; b += kBnWarps) {
    const int64_t i = (int64_t)b * ld + f;
    const float dy = a[i] > 0.f ? da[i] : 0.f;        // relu'(x) = 1[x > 0]
    const float xh = (h[i] - m) * is;
    a1 += (double)dy;
    a2 += (double)dy * (double)xh;
  }
  const double s1 = block_col_sum(a1, s_red, wy, lane);
  const double s2 = block_col_sum(a2, s_red, wy, lane);
  if (!ok) return;
  if (wy == 0) {
    if (dgamma) dgamma[f] = (float)s2;
    if (dbeta) dbeta[f] = (float)s1;
  }
  const float k1 = training ? (float)(s1 / (double)B) : 0.f;
  const float k2 = training ? (float)(s2 / (double)B) : 0.f;
  for (int b = wy; b < B; b += kBnWarps) {
    const int64_t i = (int64_t)b * ld + f;
    const float dy = a[i] > 0.f ? da[i] : 0.f;
    const float xh = (h[i] - m) * is;
    dh[i] = g * is * (dy - k1 - xh * k2);
  }
}

// ---- fc3 + mixing ---------------------------------------------------------------------------------
// a2 is [B, ld] batch-major; the activations of head hd are columns [hd*N2, (hd+1)*N2).  These
// kernels move little data (w3 is 4.8 MB for the Pascal bin heads) and are latency-bound: each one
// is shaped so that a thread's loads are independent and issued together (one memory round trip per
// output instead of a serial chain of them).  N2 % 4 == 0 and 16-byte aligned rows (checked by the
// callers), so rows are read as float4.
constexpr int kFc3Warps = 8;

// one warp per (sample, output): y[b,o] = sum_h mix[b,h] * (b3[h,o] + w3[h,o,:] . a2[b,h,:])
__global__ void __launch_bounds__(kFc3Warps * 32)
fc3_fwd_kernel(const float* __restrict__ a2, int64_t ld, const float* __restrict__ w3,
               const float* __restrict__ b3, const float* __restrict__ mix, int H, int O, int N2,
               float* __restrict__ y) {
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * kFc3Warps + warp;
  if (o >= O) return;
  const int nv = N2 >> 2;
  float acc = 0.f;
  for (int hd = 0; hd < H; ++hd) {
    const float p = __ldg(mix + (int64_t)b * H + hd);
    if (p == 0.f) continue;                           // warp-uniform
    const float4* w = reinterpret_cast<const float4*>(w3 + ((int64_t)hd * O + o) * N2);
    const float4* a = reinterpret_cast<const float4*>(a2 + (int64_t)b * ld + (int64_t)hd * N2);
    float d = 0.f;
    for (int j0 = 0; j0 < nv; j0 += 128) {            // 4 independent float4 pairs per lane in flight
      float4 wv[4], av[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * 32 + lane;
        const bool ok = j < nv;
        wv[u] = ok ? __ldg(w + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        av[u] = ok ? __ldg(a + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        d = fmaf(wv[u].x, av[u].x, fmaf(wv[u].y, av[u].y, fmaf(wv[u].z, av[u].z, fmaf(wv[u].w, av[u].w, d))));
    }
    d = warp_sum(d);
    acc += p * (d + __ldg(b3 + (int64_t)hd * O + o));
  }
  if (lane == 0) y[(int64_t)b * O + o] = acc;
}

// da2[b, hd*N2 + j] = mix[b,hd] * sum_o dy[b,o] w3[hd,o,j]   (zeros where the mixing weight is 0).
// One block per (head, sample); the O outputs are split over G thread groups, each thread keeps a
// float4 of columns, partial sums meet in shared memory.
constexpr int kFc3ActThreads = 512;

__global__ void __launch_bounds__(kFc3ActThreads)
fc3_bwd_act_kernel(const float* __restrict__ dy, const float* __restrict__ w3,
                   const float* __restrict__ mix, int64_t ld, int H, int O, int N2,
                   float* __restrict__ da2) {
  extern __shared__ __align__(16) float s_fc3[];      // [O] scaled dy, then [G][N2] partial sums
  const int hd = blockIdx.x, b = blockIdx.y;
  const float p = mix[(int64_t)b * H + hd];
  const int nv = N2 >> 2;
  float4* out = reinterpret_cast<float4*>(da2 + (int64_t)b * ld + (int64_t)hd * N2);
  if (p == 0.f) {
    for (int j = threadIdx.x; j < nv; j += blockDim.x) out[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  float* s_dy = s_fc3;
  float4* s_part = reinterpret_cast<float4*>(s_fc3 + ((O + 3) & ~3));
  for (int o = threadIdx.x; o < O; o += blockDim.x) s_dy[o] = p * dy[(int64_t)b * O + o];
  __syncthreads();
  const int W = nv < (int)blockDim.x ? nv : (int)blockDim.x;     // threads per output group
  const int G = (int)blockDim.x / W;                              // output groups
  const int og = threadIdx.x / W, jt = threadIdx.x % W;
  if (og < G) {
    const int o0 = (int)(((int64_t)O * og) / G), o1 = (int)(((int64_t)O * (og + 1)) / G);
    for (int j = jt; j < nv; j += W) {
      const float4* w = reinterpret_cast<const float4*>(w3 + (int64_t)hd * O * N2) + j;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int o = o0;
      for (; o + 8 <= o1; o += 8) {                   // 8 independent row loads in flight
        float4 wv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) wv[u] = __ldg(w + (int64_t)(o + u) * nv);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float s = s_dy[o + u];
          acc.x = fmaf(s, wv[u].x, acc.x); acc.y = fmaf(s, wv[u].y, acc.y);
          acc.z = fmaf(s, wv[u].z, acc.z); acc.w = fmaf(s, wv[u].w, acc.w);
        }
      }
      for (; o < o1; ++o) {
        const float4 wv = __ldg(w + (int64_t)o * nv);
        const float s = s_dy[o];
        acc.x = fmaf(s, wv.x, acc.x); acc.y = fmaf(s, wv.y, acc.y);
        acc.z = fmaf(s, wv.z, acc.z); acc.w = fmaf(s, wv.w, acc.w);
      }
      s_part[(int64_t)og * nv + j] = acc;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nv; j += blockDim.x) {
    float4 acc = s_part[j];
    for (int g = 1; g < G; ++g) {
      const float4 q = s_part[(int64_t)g * nv + j];
      acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
    }
    out[j] = acc;
  }
}

constexpr int kFc3WOut = 8;   // outputs per block in the weight-gradient kernel

// dw3[hd,o,j] = sum_b mix[b,hd] dy[b,o] a2[b,hd*N2+j], db3[hd,o] = sum_b mix[b,hd] dy[b,o]; only the
// samples with a non-zero mixing weight for this head are visited (one-hot: B/H of them on average)
__global__ void __launch_bounds__(256)
fc3_bwd_w_kernel(const float* __restrict__ dy, const float* __restrict__ a2, int64_t ld,
                 const float* __restrict__ mix, int B, int H, int O, int N2,
                 float* __restrict__ dw3, float* __restrict__ db3) {
  extern __shared__ __align__(16) float s_fc3[];      // [B][kFc3WOut] mix * dy, then [B] active list
  __shared__ int s_nact;
  const int hd = blockIdx.x;
  const int o0 = blockIdx.y * kFc3WOut;
  float* s_pd = s_fc3;
  int* s_act = reinterpret_cast<int*>(s_fc3 + (size_t)B * kFc3WOut);
  // ordered compaction of the active samples (ascending b: the accumulation order is fixed, so the
  // result is launch-invariant): ballot inside a warp, warp counts through shared memory
  __shared__ int s_wcnt[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int base = 0;
  for (int b0 = 0; b0 < B; b0 += blockDim.x) {
    const int b = b0 + threadIdx.x;
    const float p = b < B ? mix[(int64_t)b * H + hd] : 0.f;
    const bool act = p != 0.f;
    const unsigned m = __ballot_sync(BDP_FULL_MASK, act);
    if (lane == 0) s_wcnt[warp] = __popc(m);
    __syncthreads();
    int off = base, tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { off += w < warp ? s_wcnt[w] : 0; tot += s_wcnt[w]; }
    if (act) {
      const int slot = off + __popc(m & ((1u << lane) - 1u));
      s_act[slot] = b;
#pragma unroll
      for (int oo = 0; oo < kFc3WOut; ++oo)
        s_pd[slot * kFc3WOut + oo] = (o0 + oo < O) ? p * dy[(int64_t)b * O + o0 + oo] : 0.f;
    }
    base += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) s_nact = base;
  __syncthreads();
  const int nact = s_nact;
  for (int j = threadIdx.x; j < N2; j += blockDim.x) {
    const float* ac = a2 + (int64_t)hd * N2 + j;
    float acc[kFc3WOut];
#pragma unroll
    for (int oo = 0; oo < kFc3WOut; ++oo) acc[oo] = 0.f;
    for (int i = 0; i < nact; ++i) {
      const float av = ac[(int64_t)s_act[i] * ld];    // coalesced over j
#pragma unroll
      for (int oo = 0; oo < kFc3WOut; ++oo) acc[oo] = fmaf(s_pd[i * kFc3WOut + oo], av, acc[oo]);
    }
#pragma unroll
    for (int oo = 0; oo < kFc3WOut; ++oo)
      if (o0 + oo < O) dw3[((int64_t)hd * O + o0 + oo) * N2 + j] = acc[oo];
  }
  if (threadIdx.x < kFc3WOut && o0 + threadIdx.x < O) {
    float sacc = 0.f;
    for (int i = 0; i < nact; ++i) sacc += s_pd[i * kFc3WOut + threadIdx.x];
    db3[(int64_t)hd * O + o0 + threadIdx.x] = sacc;
  }
}

// dmix[b, h] = sum_o dy[b,o] * (b3[h,o] + w3[h,o,:] . a2[b, h*N2 : (h+1)*N2])
__global__ void __launch_bounds__(256)
fc3_bwd_mix_kernel(const float* __restrict__ dy, const float* __restrict__ a2, int64_t ld,
                   const float* __restrict__ w3, const float* __restrict__ b3, int H, int O, int N2,
                   float* __restrict__ dmix) {
  extern __shared__ float s_col[];                    // [N2]
  __shared__ float s_part[8];
  const int hd = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = threadIdx.x; j < N2; j += blockDim.x)
    s_col[j] = a2[(int64_t)b * ld + (int64_t)hd * N2 + j];
  __syncthreads();
  float acc = 0.f;
  for (int o = warp; o < O; o += 8) {
    const float* w = w3 + ((int64_t)hd * O + o) * N2;
    float d = 0.f;
    for (int j = lane; j < N2; j += 32) d = fmaf(w[j], s_col[j], d);
    d = warp_sum(d);
    acc += dy[(int64_t)b * O + o] * (d + b3[(int64_t)hd * O + o]);
  }
  if (lane == 0) s_part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += s_part[w];
    dmix[(int64_t)b * H + hd] = s;
  }
}

__global__ void __launch_bounds__(256)
sum_slabs_kernel(const float* __restrict__ parts, int64_t n, int S, int64_t stride,
                 float* __restrict__ out) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    float s = 0.f;
    for (int k = 0; k < S; ++k) s += parts[(int64_t)k * stride + i];
    out[i] = s;
  }
}

}  // namespace

extern "C" int bdp_bn_relu_fwd(const float* h, int64_t F, int64_t B, int64_t ld,
                               const float* gamma, const float* beta, float* running_mean,
                               float* running_var, float* save_mean, float* save_invstd, float eps,
                               float momentum, int training, float* a, void* stream) {
  BDP_REQUIRE(h && gamma && beta && a, "bn_relu_fwd: NULL buffer");
  BDP_REQUIRE(F > 0 && B > 0 && ld >= F, "bn_relu_fwd: bad sizes F=%lld B=%lld ld=%lld",
              (long long)F, (long long)B, (long long)ld);
  BDP_REQUIRE(training || (running_mean && running_var), "bn_relu_fwd: eval mode needs running stats");
  const unsigned blocks = (unsigned)ceil_div64(F, 32);
  bn_relu_fwd_kernel<<<blocks, kBnWarps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      h, F, (int)B, ld, gamma, beta, running_mean, running_var, save_mean, save_invstd, eps,
      momentum, training, a);
  BDP_CUDA_CHECK_LAUNCH("bn_relu_fwd_kernel");
  return BDP_OK;
}

extern "C" int bdp_bn_relu_bwd(const float* da, const float* a, const float* h, const float* gamma,
                               const float* save_mean, const float* save_invstd, int64_t F,
                               int64_t B, int64_t ld, int training, float* dh, float* dgamma,
                               float* dbeta, void* stream) {
  BDP_REQUIRE(da && a && h && gamma && save_mean && save_invstd && dh, "bn_relu_bwd: NULL buffer");
  BDP_REQUIRE(F > 0 && B > 0 && ld >= F, "bn_relu_bwd: bad sizes");
  const unsigned blocks = (unsigned)ceil_div64(F, 32);
  bn_relu_bwd_kernel<<<blocks, kBnWarps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      da, a, h, gamma, save_mean, save_invstd, F, (int)B, ld, training, dh, dgamma, dbeta);
  BDP_CUDA_CHECK_LAUNCH("bn_relu_bwd_kernel");
  return BDP_OK;
}

extern "C" int bdp_head_fc3_fwd(const float* a2, int64_t ld, const float* w3, const float* b3,
                                const float* mix, int64_t B, int H, int O, int N2, float* y,
                                void* stream) {
  BDP_REQUIRE(a2 && w3 && b3 && mix && y, "head_fc3_fwd: NULL buffer");
  BDP_REQUIRE(B > 0 && B <= 65535 && H > 0 && O > 0 && N2 > 0 && N2 <= 12000,
              "head_fc3_fwd: bad sizes B=%lld H=%d O=%d N2=%d", (long long)B, H, O, N2);
  BDP_REQUIRE(N2 % 4 == 0 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(a2) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(w3) & 15) == 0,
              "head_fc3_fwd: N2 and ld must be multiples of 4 and a2 / w3 16-byte aligned");
  dim3 grid((unsigned)((O + kFc3Warps - 1) / kFc3Warps), (unsigned)B);
  fc3_fwd_kernel<<<grid, kFc3Warps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      a2, ld, w3, b3, mix, H, O, N2, y);
  BDP_CUDA_CHECK_LAUNCH("fc3_fwd_kernel");
  return BDP_OK;
}

extern "C" int bdp_head_fc3_bwd(const float* dy, const float* a2, int64_t ld, const float* w3,
                                const float* b3, const float* mix, int64_t B, int H, int O, int N2,
                                float* da2, float* dw3, float* db3, float* dmix, void* stream) {
  BDP_REQUIRE(dy && a2 && w3 && b3 && mix, "head_fc3_bwd: NULL buffer");
  BDP_REQUIRE(B > 0 && B <= 65535 && H > 0 && H <= 65535 && O > 0 && O <= 12000 && N2 > 0 &&
                  N2 <= 12000, "head_fc3_bwd: bad sizes");
  BDP_REQUIRE((size_t)B * (kFc3WOut + 1) * 4 <= 48 * 1024, "head_fc3_bwd: batch too large (%lld)", (long long)B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  BDP_REQUIRE(N2 % 4 == 0 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(w3) & 15) == 0 &&
                  (!da2 || (reinterpret_cast<uintptr_t>(da2) & 15) == 0),
              "head_fc3_bwd: N2 and ld must be multiples of 4 and w3 / da2 16-byte aligned");
  if (da2) {
    const int nv = N2 / 4;
    const int W = nv < kFc3ActThreads ? nv : kFc3ActThreads;
    const int G = kFc3ActThreads / W;
    const size_t smem = ((size_t)((O + 3) & ~3) + (size_t)G * N2) * sizeof(float);
    if (smem > 48 * 1024)
      BDP_CUDA_CALL(cudaFuncSetAttribute(fc3_bwd_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fc3_bwd_act_kernel<<<dim3(H, (unsigned)B), kFc3ActThreads, smem, st>>>(dy, w3, mix, ld, H, O, N2, da2);
    BDP_CUDA_CHECK_LAUNCH("fc3_bwd_act_kernel");
  }
  if (dw3) {
    BDP_REQUIRE(db3 != nullptr, "head_fc3_bwd: db3 is NULL");
    fc3_bwd_w_kernel<<<dim3(H, (unsigned)((O + kFc3WOut - 1) / kFc3WOut)), 256,
                       (size_t)B * (kFc3WOut + 1) * sizeof(float), st>>>(dy, a2, ld, mix, (int)B, H, O,
                                                                         N2, dw3, db3);
    BDP_CUDA_CHECK_LAUNCH("fc3_bwd_w_kernel");
  }
  if (dmix) {
    fc3_bwd_mix_kernel<<<dim3(H, (unsigned)B), 256, N2 * sizeof(float), st>>>(dy, a2, ld, w3, b3, H,
                                                                             O, N2, dmix);
    BDP_CUDA_CHECK_LAUNCH("fc3_bwd_mix_kernel");
  }
  return BDP_OK;
}

extern "C" int bdp_sum_slabs(const float* parts, int64_t n, int S, int64_t stride, float* out,
                             void* stream) {
  BDP_REQUIRE(parts && out && n >= 0 && S >= 1, "sum_slabs: bad arguments");
  if (n == 0) return BDP_OK;
  int64_t blocks = ceil_div64(n, 256);
  const int64_t cap = (int64_t)bdp_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  sum_slabs_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      parts, n, S, stride, out);
  BDP_CUDA_CHECK_LAUNCH("sum_slabs_kernel");
  return BDP_OK;
}
