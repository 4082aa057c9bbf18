// (b) Fused bin-delta loss for sm_100a: softmax cross-entropy over the K pose bins, argmax key
// gather, pose composition (additive or Rodrigues-on-key) and the geodesic / MSE pose loss, with
// the closed-form backward written in the same pass.
//
// Layout: one warp walks R consecutive rows.  For each row the 32 lanes cooperate on the logits
// (128-bit streaming loads held in registers, shuffle max/argmax/sum, gradient written straight
// back with 128-bit streaming stores: the logits are read once and the gradient written once);
// lane r then keeps (ce, argmax) of row r, and after the R rows all lanes do their row's pose
// math in parallel (so the transcendental tail costs one instruction stream per 32 rows instead
// of one per row).  R adapts to B so that small training batches still spread over the chip.
//
// Algorithmic traffic per row (K=200, fp32): 2*K*4 (logits in, grad out) + 8 (bin) + 12 (delta)
// + 12..36 (target) + 12 (grad delta) = 1644..1668 B  — the kernel is HBM-bound.
#include "common.cuh"

namespace {

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}


struct LossParams {
  const float* logits;
  int64_t B, K, ld;
  const int64_t* bin_true;
  const float* pred;
  int ndim;
  const float* keys;
  int use_keys;
  const float* target;
  int tdim;
  int pose_mode;
  float* out_loss;
  float* row_ce;
  float* row_pose;
  float* grad_logits;
  float* grad_pred;
  int64_t* argmax_out;
  double* partials;       // [gridDim.x * 2]
  unsigned int* ticket;   // [1], zero on entry, zero on exit
  int rows_per_warp;      // R in [1,32]
  float inv_B;              // gradient scale: 1/B (mean) or 1 (per-sample)
};

// ---- pose losses: value + gradient w.r.t. the composed prediction ---------------------------
// The per-row pose math is O(1) next to the O(K) logits, and its conditioning is poor exactly where
// training ends up (small angles: acos near 1, 1/sqrt(1 - w^2)), so it is evaluated in fp64 from
// the fp32 inputs and rounded once.  The result is the correctly rounded value of the reference's
// formula; the reference's own fp32 evaluation differs from it by its rounding noise
// (~2^-24 / (1 - w^2) relative).  eps / clamp conventions are the reference's, cited inline.
#define BDP_EPS_D 1e-6
#define BDP_NORM_EPS_D 1e-12

// axisAngle.geodesic_loss.forward, axisAngle.py:110-120.  p = predicted axis-angle (key + delta),
// t = ground truth.  Returns theta; g = d theta / d p.
__device__ __forceinline__ double pose_geodesic_aa(const double p[3], const double t[3],
                                                   double g[3]) {
  const double ap = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  const double at = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
  const double ipn = 1.0 / fmax(ap, BDP_NORM_EPS_D);   // F.normalize: v / max(||v||, 1e-12)
  const double itn = 1.0 / fmax(at, BDP_NORM_EPS_D);
  const double ph[3] = {p[0] * ipn, p[1] * ipn, p[2] * ipn};
  const double th[3] = {t[0] * itn, t[1] * itn, t[2] * itn};
  const double d = th[0] * ph[0] + th[1] * ph[1] + th[2] * ph[2];
  double sp, cp, st, ct;
  sincos(0.5 * ap, &sp, &cp);
  sincos(0.5 * at, &st, &ct);
  const double w = ct * cp + st * sp * d;
  const double c = fabs(w);
  const double cc = fmin(c, 1.0 - BDP_EPS_D);          // clamp(|w|, -1+eps, 1-eps); |w| >= 0
  const double theta = 2.0 * acos(cc);
  // backward: zero where the clamp saturates (torch.clamp passes the
  // gradient on the closed interval), sign(0) = 0 from torch.abs.
  double dth_dw = 0.0;
  if (c <= 1.0 - BDP_EPS_D && w != 0.0) dth_dw = (w > 0.0 ? -2.0 : 2.0) / sqrt(1.0 - c * c);
  const double dw_dap = 0.5 * (st * cp * d - ct * sp);
  const double k = st * sp;                             // dw/d p_hat = k * t_hat
  if (ap > BDP_NORM_EPS_D) {
    // d ap/dp = p_hat ; d p_hat/dp = (I - p_hat p_hat^T)/ap
    const double proj = k * d;                          // p_hat . (k t_hat)
    const double ia = 1.0 / ap;
#pragma unroll
    for (int i = 0; i < 3; ++i)
      g[i] = dth_dw * (dw_dap * ph[i] + (k * th[i] - proj * ph[i]) * ia);
  } else {
    // ||p|| <= 1e-12: normalize is p/eps (linear), torch.norm's subgradient at 0 is 0
    const double s = (ap > 0.0) ? dw_dap / ap : 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) g[i] = dth_dw * (s * p[i] + k * th[i] * (1.0 / BDP_NORM_EPS_D));
  }
  return theta;
}

// quaternion.geodesic_loss.forward, quaternion.py:156-163.
__device__ __forceinline__ double pose_geodesic_quat(const double q[4], const double t[4],
                                                     double g[4]) {
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double in = 1.0 / fmax(n, BDP_NORM_EPS_D);
  const double qh[4] = {q[0] * in, q[1] * in, q[2] * in, q[3] * in};
  const double w = t[0] * qh[0] + t[1] * qh[1] + t[2] * qh[2] + t[3] * qh[3];
  const double c = fabs(w);
  const double theta = 2.0 * acos(fmin(c, 1.0 - BDP_EPS_D));
  double dth_dw = 0.0;
  if (c <= 1.0 - BDP_EPS_D && w != 0.0) dth_dw = (w > 0.0 ? -2.0 : 2.0) / sqrt(1.0 - c * c);
  if (n > BDP_NORM_EPS_D) {
#pragma unroll
    for (int i = 0; i < 4; ++i) g[i] = dth_dw * (t[i] - qh[i] * w) * in;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) g[i] = dth_dw * t[i] * (1.0 / BDP_NORM_EPS_D);
  }
  return theta;
}

// RiemannianLoss, binDeltaLosses.py:221-239: R_hat = Key * exp([r]x), phi = acos(clamp((tr(R_hat^T
// R) - 1)/2)).  With M = Key^T R:  tr = sum_ij E_ij M_ij,  E = I + sin(th)[a]x + (1-cos th)[a]x^2.
__device__ __forceinline__ double pose_riemannian(const double r[3], const float* __restrict__ key,
                                                  const double R[9], double g[3]) {
  double Kd[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) Kd[i] = (double)__ldg(key + i);
  double M[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      M[i * 3 + j] = Kd[0 * 3 + i] * R[0 * 3 + j] + Kd[1 * 3 + i] * R[1 * 3 + j] +
                     Kd[2 * 3 + i] * R[2 * 3 + j];
  const double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  const double in = 1.0 / fmax(th, BDP_NORM_EPS_D);
  const double a[3] = {r[0] * in, r[1] * in, r[2] * in};
  double s, c;
  sincos(th, &s, &c);
  const double omc = 1.0 - c;
  // skew from the reference's `proj` rows (binDeltaLosses.py:217): [[0,-a3,a2],[a3,0,-a1],[-a2,a1,0]]
  const double A[9] = {0.0, -a[2], a[1], a[2], 0.0, -a[0], -a[1], a[0], 0.0};
  double tr = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double AA = A[i * 3 + 0] * A[0 * 3 + j] + A[i * 3 + 1] * A[1 * 3 + j] +
                        A[i * 3 + 2] * A[2 * 3 + j];
      const double E = (i == j ? 1.0 : 0.0) + s * A[i * 3 + j] + omc * AA;
      tr += E * M[i * 3 + j];
    }
  const double u = 0.5 * (tr - 1.0);
  const double uc = fmin(fmax(u, -1.0 + BDP_EPS_D), 1.0 - BDP_EPS_D);
  const double phi = acos(uc);
  double dphi_dt = 0.0;
  if (u >= -1.0 + BDP_EPS_D && u <= 1.0 - BDP_EPS_D) dphi_dt = -0.5 / sqrt(1.0 - u * u);
  if (th > BDP_NORM_EPS_D) {
    const double trM = M[0] + M[4] + M[8];
    const double m[3] = {M[7] - M[5], M[2] - M[6], M[3] - M[1]};
    const double am = a[0] * m[0] + a[1] * m[1] + a[2] * m[2];
    double Sa[3];  // (M + M^T) a
#pragma unroll
    for (int i = 0; i < 3; ++i)
      Sa[i] = (M[i * 3 + 0] + M[0 * 3 + i]) * a[0] + (M[i * 3 + 1] + M[1 * 3 + i]) * a[1] +
              (M[i * 3 + 2] + M[2 * 3 + i]) * a[2];
    const double aMa = 0.5 * (a[0] * Sa[0] + a[1] * Sa[1] + a[2] * Sa[2]);
    const double dt_dth = c * am + s * (aMa - trM);
    double v[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = s * m[i] + omc * Sa[i];
    const double av = a[0] * v[0] + a[1] * v[1] + a[2] * v[2];
    const double ith = 1.0 / th;
#pragma unroll
    for (int i = 0; i < 3; ++i) g[i] = dphi_dt * (dt_dth * a[i] + (v[i] - av * a[i]) * ith);
  } else {
    g[0] = g[1] = g[2] = 0.0;
  }
  return phi;
}

// RiemannianLoss.my_loss on explicit matrices (binDeltaLosses.py:221-225): P predicted, T truth.
__device__ __forceinline__ double pose_rotmat(const double P[9], const double T[9], double g[9]) {
  double tr = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) tr += P[i] * T[i];
  const double u = 0.5 * (tr - 1.0);
  const double uc = fmin(fmax(u, -1.0 + BDP_EPS_D), 1.0 - BDP_EPS_D);
  double dphi_dt = 0.0;
  if (u >= -1.0 + BDP_EPS_D && u <= 1.0 - BDP_EPS_D) dphi_dt = -0.5 / sqrt(1.0 - u * u);
#pragma unroll
  for (int i = 0; i < 9; ++i) g[i] = dphi_dt * T[i];
  return acos(uc);
}

// Per-row pose phase. Returns the row's pose-loss value, writes the row's gradient.
__device__ __forceinline__ float pose_row(const LossParams& P, int64_t row, int ind) {
  const int nd = P.ndim;
  float p[9], t[9];
  const float* pr = P.pred + row * nd;
#pragma unroll
  for (int i = 0; i < 9; ++i) p[i] = (i < nd) ? __ldg(pr + i) : 0.f;
  const float* tg = P.target + row * P.tdim;
#pragma unroll
  for (int i = 0; i < 9; ++i) t[i] = (i < P.tdim) ? __ldg(tg + i) : 0.f;
  if (P.use_keys && P.pose_mode != BDP_POSE_RIEMANNIAN) {
    // centers[argmax] + delta: one fp32 add, as the reference composes it
    // (learnGeodesicBDModel.py:176-177, binDeltaLosses.py:48-49)
    const float* key = P.keys + (int64_t)ind * nd;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < nd) p[i] = __fadd_rn(p[i], __ldg(key + i));
  }
  float val;
  float gf[9];
  if (P.pose_mode == BDP_POSE_MSE) {   // nn.MSELoss: mean over B*ndim elements (fp32, as torch)
    const float ind_nd = 1.f / (float)nd;
    val = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const float e = (i < nd) ? p[i] - t[i] : 0.f;
      val += e * e;
      gf[i] = 2.f * e * ind_nd;
    }
    val *= ind_nd;
  } else {
    double pd[9], td[9], g[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { pd[i] = (double)p[i]; td[i] = (double)t[i]; g[i] = 0.0; }
    double v;
    if (P.pose_mode == BDP_POSE_RIEMANNIAN) v = pose_riemannian(pd, P.keys + (int64_t)ind * 9, td, g);
    else if (P.pose_mode == BDP_POSE_GEODESIC_AA) v = pose_geodesic_aa(pd, td, g);
    else if (P.pose_mode == BDP_POSE_GEODESIC_Q) v = pose_geodesic_quat(pd, td, g);
    else v = pose_rotmat(pd, td, g);
    val = (float)v;
#pragma unroll
    for (int i = 0; i < 9; ++i) gf[i] = (float)(g[i] * (double)P.inv_B);
  }
  if (P.grad_pred) {
    float* gp = P.grad_pred + row * nd;
    const float sc = (P.pose_mode == BDP_POSE_MSE) ? P.inv_B : 1.f;
#pragma unroll
    for (int i = 0; i < 9; ++i)
      if (i < nd) gp[i] = gf[i] * sc;
  }
  if (P.row_pose) P.row_pose[row] = val;
  return val;
}

// ---- logits phase, vector path: K % 4 == 0, rows 16-byte aligned, K <= KV*128 -----------------
template <int KV>
__device__ __forceinline__ void row_load(const LossParams& P, int64_t row, int lane, int nvec,
                                         float4 v[KV], int& tgt) {
  // the target bin is fetched with the row: read right before its use it exposed one full memory
  // latency per row (15 % of all stall samples)
  tgt = (int)__ldg(P.bin_true + row);
  const float4* src = reinterpret_cast<const float4*>(P.logits + row * P.ld);
#pragma unroll
  for (int j = 0; j < KV; ++j) {
    const int i = lane + j * 32;
    v[j] = (i < nvec) ? ldg_stream(src + i)
                      : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  }
}

template <int KV>
__device__ __forceinline__ void row_softmax_ce(const LossParams& P, int64_t row, int lane,
                                               int nvec, float4 v[KV], int tgt, float& ce, int& amax) {
  // max, then argmax (lowest index wins).  The warp maximum goes through the integer REDUX unit on
  // an order-preserving key (one instruction instead of a 5-step shuffle tree); only the lanes that
  // hold the maximum look for its position.
  float lm = -INFINITY;
#pragma unroll
  for (int j = 0; j < KV; ++j) lm = fmaxf(lm, fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w)));
  const unsigned lb = __float_as_uint(lm);
  const unsigned key = (lb & 0x80000000u) ? ~lb : (lb | 0x80000000u);      // monotone float -> uint
  const unsigned wk = __reduce_max_sync(BDP_FULL_MASK, key);
  const float m = __uint_as_float((wk & 0x80000000u) ? (wk & 0x7fffffffu) : ~wk);
  int mi = 0x7fffffff;
  if (lm == m) {                                   // float compare: -0.0 and +0.0 tie, as in torch.max
#pragma unroll
    for (int j = KV - 1; j >= 0; --j) {
      const int base = (lane + j * 32) * 4;
      if (v[j].w == m) mi = base + 3;
      if (v[j].z == m) mi = base + 2;
      if (v[j].y == m) mi = base + 1;
      if (v[j].x == m) mi = base;
    }
  }
  mi = __reduce_min_sync(BDP_FULL_MASK, mi);
  // the target logit sits in exactly one lane's registers.  tgt is warp-uniform, so the vector index
  // and the component are selected with uniform predicates (7 selects for KV = 2 instead of a
  // compare + select per element), then the owning lane broadcasts.
  float xt;
  {
    const int tj = tgt >> 7, tq = tgt & 3;
    float4 w = v[0];
#pragma unroll
    for (int j = 1; j < KV; ++j)
      if (tj == j) w = v[j];
    const float a = (tq & 1) ? w.y : w.x, b = (tq & 1) ? w.w : w.z;
    xt = __shfl_sync(BDP_FULL_MASK, (tq & 2) ? b : a, (tgt >> 2) & 31);
  }
  // 2^(x*log2e - m*log2e): one FMA + ex2.approx per element.  The product m*log2e is rounded once
  // (mh); its rounding error ml = m*log2e - mh (exact, one more FMA per row) multiplies every term by
  // the same 2^ml, which cancels in the probabilities and is taken out of log(s) below.
  const float kLog2e = 1.4426950408889634f;
  const float mh = m * kLog2e;
  const float ml = fmaf(m, kLog2e, -mh);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < KV; ++j) {
    float e[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      e[q] = ex2_approx(fmaf(e[q], kLog2e, -mh));              // exp(-inf) = 0 for the padding lanes
      s += e[q];
    }
    v[j] = make_float4(e[0], e[1], e[2], e[3]);
  }
  s = warp_sum(s);
  ce = (m - xt) + fmaf(-ml, 0.6931471805599453f, logf(s));
  amax = mi;
  if (P.grad_logits) {
    const float sc = __fdividef(P.inv_B, s);                   // s in [1, K]: 2 ulp
    float4* dst = reinterpret_cast<float4*>(P.grad_logits + row * P.ld);
    const int tv = tgt >> 2, tq = tgt & 3;
#pragma unroll
    for (int j = 0; j < KV; ++j) {
      const int i = lane + j * 32;
      if (i < nvec) {
        const float sub = (i == tv) ? P.inv_B : 0.f;
        float4 o;
        o.x = fmaf(v[j].x, sc, tq == 0 ? -sub : 0.f);
        o.y = fmaf(v[j].y, sc, tq == 1 ? -sub : 0.f);
        o.z = fmaf(v[j].z, sc, tq == 2 ? -sub : 0.f);
        o.w = fmaf(v[j].w, sc, tq == 3 ? -sub : 0.f);
        stg_stream(dst + i, o);
      }
    }
  }
}

// scalar path: any K / alignment; three passes over the row (the 2nd and 3rd hit L1/L2).
__device__ __forceinline__ void row_softmax_ce_scalar(const LossParams& P, int64_t row, int lane,
                                                      float& ce, int& amax) {
  const float* x = P.logits + row * P.ld;
  const int K = (int)P.K;
  float m = -INFINITY;
  int mi = 0x7fffffff;
  for (int i = lane; i < K; i += 32) {
    const float e = x[i];
    if (e > m) { m = e; mi = i; }
  }
  warp_argmax(m, mi);
  const int tgt = (int)__ldg(P.bin_true + row);
  float s = 0.f;
  for (int i = lane; i < K; i += 32) s += expf(x[i] - m);
  s = warp_sum(s);
  const float xt = x[tgt];
  ce = (m - xt) + logf(s);
  amax = mi;
  if (P.grad_logits) {
    const float sc = P.inv_B / s;
    float* dst = P.grad_logits + row * P.ld;
    for (int i = lane; i < K; i += 32)
      dst[i] = expf(x[i] - m) * sc - (i == tgt ? P.inv_B : 0.f);
  }
}

template <int KV>   // KV == 0: scalar path
__global__ void __launch_bounds__(128, 5) bd_loss_kernel(const LossParams P) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int R = P.rows_per_warp;
  const int64_t n_groups = (P.B + R - 1) / R;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int nvec = (int)(P.K >> 2);
  const bool has_ce = P.logits != nullptr;
  const bool has_pose = P.pose_mode != BDP_POSE_NONE;

  double acc_ce = 0.0, acc_pose = 0.0;

  for (int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; g < n_groups;
       g += warps_total) {
    const int64_t row0 = g * R;
    const int nrows = (int)min((int64_t)R, P.B - row0);
    float my_ce = 0.f;
    int my_ind = 0;
    if (has_ce) {
      if (KV > 0) {
        // Software pipeline over row pairs: the loads of the NEXT pair are issued before the current
        // pair is reduced, so a warp keeps up to four rows (3.2 KB at K=200) in flight — the kernel
        // is bound by bytes in flight per SM, not by issue slots or DRAM bandwidth.
        constexpr int KVV = KV > 0 ? KV : 1;
        int r = 0;
        float4 va[KVV], vb[KVV], na[KVV], nb[KVV];
        int ta = 0, tb = 0, tna = 0, tnb = 0;
        if (nrows >= 2) {
          row_load<KVV>(P, row0, lane, nvec, va, ta);
          row_load<KVV>(P, row0 + 1, lane, nvec, vb, tb);
        }
        for (; r + 1 < nrows; r += 2) {
          const bool more = r + 3 < nrows;
          if (more) {
            row_load<KVV>(P, row0 + r + 2, lane, nvec, na, tna);
            row_load<KVV>(P, row0 + r + 3, lane, nvec, nb, tnb);
          }
          float ce_a, ce_b;
          int ia, ib;
          row_softmax_ce<KVV>(P, row0 + r, lane, nvec, va, ta, ce_a, ia);
          row_softmax_ce<KVV>(P, row0 + r + 1, lane, nvec, vb, tb, ce_b, ib);
          if (lane == r) { my_ce = ce_a; my_ind = ia; }
          if (lane == r + 1) { my_ce = ce_b; my_ind = ib; }
          if (more) {
#pragma unroll
            for (int j = 0; j < KVV; ++j) { va[j] = na[j]; vb[j] = nb[j]; }
            ta = tna; tb = tnb;
          }
        }
        if (r < nrows) {
          row_load<KVV>(P, row0 + r, lane, nvec, va, ta);
          float ce_a;
          int ia;
          row_softmax_ce<KVV>(P, row0 + r, lane, nvec, va, ta, ce_a, ia);
          if (lane == r) { my_ce = ce_a; my_ind = ia; }
        }
      } else {
        for (int r = 0; r < nrows; ++r) {
          float ce_a;
          int ia;
          row_softmax_ce_scalar(P, row0 + r, lane, ce_a, ia);
          if (lane == r) { my_ce = ce_a; my_ind = ia; }
        }
      }
    }
    if (lane < nrows) {
      const int64_t row = row0 + lane;
      if (has_ce) {
        acc_ce += (double)my_ce;
        if (P.row_ce) P.row_ce[row] = my_ce;
        if (P.argmax_out) P.argmax_out[row] = (int64_t)my_ind;
      }
      if (has_pose) acc_pose += (double)pose_row(P, row, my_ind);
    }
  }

  // deterministic reduction: warp -> block (fixed order) -> per-block partial -> last block sums
  // the partials in block order.
  __shared__ double s_part[2][8];
  __shared__ bool s_last;
  acc_ce = warp_sum(acc_ce);
  acc_pose = warp_sum(acc_pose);
  if (lane == 0) { s_part[0][warp] = acc_ce; s_part[1][warp] = acc_pose; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += s_part[0][w]; b += s_part[1][w]; }
    P.partials[2 * blockIdx.x + 0] = a;
    P.partials[2 * blockIdx.x + 1] = b;
    __threadfence();
    const unsigned int t = atomicAdd(P.ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && warp == 0) {
    __threadfence();
    double a = 0.0, b = 0.0;
    // lanes take interleaved blocks, then a fixed-shape shuffle tree: order is launch-invariant
    for (int i = lane; i < (int)gridDim.x; i += 32) {
      a += __ldcg(P.partials + 2 * i + 0);
      b += __ldcg(P.partials + 2 * i + 1);
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) {
      P.out_loss[0] = (float)(a / (double)P.B);
      P.out_loss[1] = (float)(b / (double)P.B);
      *P.ticket = 0u;   // leave the workspace reusable
    }
  }
}

// ---- expected pose loss of the soft-bin family -------------------------------
// E_b = sum_k softmax(score_b)_k * L(ydata_b, key_k + delta_b[k])      binDeltaLosses.py:119-123,
// 141-146, 162-167, 176-180, 311-316, 327-333: the reference loops over the K bins in python (K pose-loss
// launches forward, K backward); here one warp owns a row and its lanes the bins.  L is the
// axis-angle geodesic distance (symmetric in its arguments) or the quaternion one with the prediction
// in the un-normalised `ytrue` slot, as the reference calls it.
//   dE/dscore_j = p_j (L_j - E),   dE/ddelta = sum_k p_k dL_k/dpose  (shared delta)  |  p_k dL_k/dpose (per bin)
__device__ __forceinline__ double quat_second_arg(const double pose[4], const double t[4], double g[4]) {
  // quaternion.py:156-163 with ypred = t (normalised), ytrue = pose (as is)
  const double n = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2] + t[3] * t[3]);
  const double in = 1.0 / fmax(n, BDP_NORM_EPS_D);
  const double th[4] = {t[0] * in, t[1] * in, t[2] * in, t[3] * in};
  const double w = pose[0] * th[0] + pose[1] * th[1] + pose[2] * th[2] + pose[3] * th[3];
  const double c = fabs(w);
  const double theta = 2.0 * acos(fmin(c, 1.0 - BDP_EPS_D));
  double dth_dw = 0.0;
  if (c <= 1.0 - BDP_EPS_D && w != 0.0) dth_dw = (w > 0.0 ? -2.0 : 2.0) / sqrt(1.0 - c * c);
#pragma unroll
  for (int i = 0; i < 4; ++i) g[i] = dth_dw * th[i];
  return theta;
}

struct ExpParams {
  const float* logits; int64_t B, K, ld;
  const float* delta; int per_bin; int nd;
  const float* keys; const float* target; int mode;
  float* rows; float* g_logits; float* g_delta;
};

__global__ void __launch_bounds__(128) expected_pose_kernel(const ExpParams P) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= P.B) return;
  const float* s = P.logits + row * P.ld;
  const int K = (int)P.K, nd = P.nd;
  float m = -INFINITY;
  for (int k = lane; k < K; k += 32) m = fmaxf(m, s[k]);
  m = warp_max(m);
  float z = 0.f;
  for (int k = lane; k < K; k += 32) z += expf(s[k] - m);
  z = warp_sum(z);
  const float iz = 1.f / z;
  double t[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = 0; i < nd; ++i) t[i] = (double)__ldg(P.target + row * nd + i);
  double e_acc = 0.0, gs[4] = {0.0, 0.0, 0.0, 0.0};
  float* gl = P.g_logits + row * K;                   // holds L_k between the two passes
  for (int k = lane; k < K; k += 32) {
    const float* d = P.per_bin ? P.delta + (row * K + k) * nd : P.delta + row * nd;
    double pose[4] = {0.0, 0.0, 0.0, 0.0}, g[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = 0; i < nd; ++i) pose[i] = (double)__fadd_rn(__ldg(d + i), __ldg(P.keys + (int64_t)k * nd + i));
    const double Lk = P.mode == BDP_POSE_GEODESIC_AA ? pose_geodesic_aa(pose, t, g) : quat_second_arg(pose, t, g);
    const double p = (double)(expf(s[k] - m) * iz);
    e_acc += p * Lk;
    gl[k] = (float)Lk;
    if (P.per_bin) {
      float* gd = P.g_delta + (row * K + k) * nd;
      for (int i = 0; i < nd; ++i) gd[i] = (float)(p * g[i]);
    } else {
      for (int i = 0; i < nd; ++i) gs[i] += p * g[i];
    }
  }
  e_acc = warp_sum(e_acc);
  if (!P.per_bin) {
    for (int i = 0; i < nd; ++i) gs[i] = warp_sum(gs[i]);
    if (lane == 0) for (int i = 0; i < nd; ++i) P.g_delta[row * nd + i] = (float)gs[i];
  }
  if (lane == 0) P.rows[row] = (float)e_acc;
  __syncwarp();
  for (int k = lane; k < K; k += 32) {
    const float p = expf(s[k] - m) * iz;
    gl[k] = p * (gl[k] - (float)e_acc);
  }
}

constexpr int kLossThreads = 128;
constexpr int kLossMaxBlocks = 148 * 16;

}  // namespace

extern "C" int64_t bdp_bd_loss_workspace_bytes(int64_t /*B*/) {
  return (int64_t)kLossMaxBlocks * 2 * sizeof(double) + 64;
}

extern "C" int bdp_bd_loss_fwd_bwd(const float* logits, int64_t B, int64_t K, int64_t ld_logits,
                                   const int64_t* bin_true, const float* pred, int ndim,
                                   const float* keys, int use_keys, const float* target,
                                   int pose_mode, float* out_loss, float* row_ce, float* row_pose,
                                   float* grad_logits, float* grad_pred, float grad_scale,
                                   int64_t* argmax_out, void* workspace, int64_t workspace_bytes, void* stream) {
  BDP_REQUIRE(B > 0, "bd_loss: B must be positive (got %lld)", (long long)B);
  BDP_REQUIRE(out_loss != nullptr, "bd_loss: out_loss is NULL");
  BDP_REQUIRE(workspace != nullptr && workspace_bytes >= bdp_bd_loss_workspace_bytes(B),
              "bd_loss: workspace too small (%lld < %lld)", (long long)workspace_bytes,
              (long long)bdp_bd_loss_workspace_bytes(B));
  BDP_REQUIRE(pose_mode >= BDP_POSE_NONE && pose_mode <= BDP_POSE_ROTMAT,
              "bd_loss: unknown pose_mode %d", pose_mode);
  if (logits) {
    BDP_REQUIRE(K > 0 && ld_logits >= K, "bd_loss: bad K=%lld ld=%lld", (long long)K,
                (long long)ld_logits);
    BDP_REQUIRE(bin_true != nullptr, "bd_loss: bin_true is NULL with logits given");
  } else {
    BDP_REQUIRE(!use_keys && pose_mode != BDP_POSE_RIEMANNIAN,
                "bd_loss: key gather needs logits (argmax)");
    BDP_REQUIRE(pose_mode != BDP_POSE_NONE, "bd_loss: nothing to compute");
  }
  int tdim = ndim;
  if (pose_mode != BDP_POSE_NONE) {
    BDP_REQUIRE(pred != nullptr && target != nullptr, "bd_loss: pred/target is NULL");
    switch (pose_mode) {
      case BDP_POSE_MSE: BDP_REQUIRE(ndim >= 1 && ndim <= 9, "bd_loss: MSE ndim %d", ndim); break;
      case BDP_POSE_GEODESIC_AA: BDP_REQUIRE(ndim == 3, "bd_loss: axis-angle needs ndim=3"); break;
      case BDP_POSE_GEODESIC_Q: BDP_REQUIRE(ndim == 4, "bd_loss: quaternion needs ndim=4"); break;
      case BDP_POSE_RIEMANNIAN:
        BDP_REQUIRE(ndim == 3 && keys != nullptr, "bd_loss: riemannian needs ndim=3 and keys");
        tdim = 9;
        break;
      case BDP_POSE_ROTMAT: BDP_REQUIRE(ndim == 9 && !use_keys, "bd_loss: rotmat needs ndim=9"); break;
    }
    if (use_keys) BDP_REQUIRE(keys != nullptr && ndim <= 4, "bd_loss: use_keys needs keys, ndim<=4");
  }

  LossParams P;
  P.logits = logits; P.B = B; P.K = K; P.ld = ld_logits; P.bin_true = bin_true;
  P.pred = pred; P.ndim = ndim; P.keys = keys; P.use_keys = use_keys; P.target = target;
  P.tdim = tdim; P.pose_mode = pose_mode; P.out_loss = out_loss; P.row_ce = row_ce;
  P.row_pose = row_pose; P.grad_logits = grad_logits; P.grad_pred = grad_pred;
  P.argmax_out = argmax_out;
  P.partials = reinterpret_cast<double*>(workspace);
  P.ticket = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(workspace) +
                                             (size_t)kLossMaxBlocks * 2 * sizeof(double));
  P.inv_B = grad_scale > 0.f ? grad_scale : 1.f / (float)B;

  // rows per warp: keep >= ~16 warps per SM busy before making warps walk several rows
  const int sms = bdp_num_sms();
  const int64_t want_warps = (int64_t)sms * 16;
  int R = (int)((B + want_warps - 1) / want_warps);
  R = R < 1 ? 1 : (R > 32 ? 32 : R);
  if (R > 1 && (R & 1)) ++R;          // even, so the 2-rows-in-flight loop has no tail
  if (R > 32) R = 32;
  P.rows_per_warp = R;
  const int64_t groups = (B + R - 1) / R;
  const int warps_per_block = kLossThreads / 32;
  int64_t blocks = (groups + warps_per_block - 1) / warps_per_block;
  if (blocks > kLossMaxBlocks) blocks = kLossMaxBlocks;

  const bool vec_ok = logits && (K % 4 == 0) && (ld_logits % 4 == 0) && K <= 1024 &&
                      ((reinterpret_cast<uintptr_t>(logits) & 15) == 0) &&
                      (!grad_logits || (reinterpret_cast<uintptr_t>(grad_logits) & 15) == 0);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!logits || !vec_ok) {
    bd_loss_kernel<0><<<(unsigned)blocks, kLossThreads, 0, st>>>(P);
  } else if (K <= 128) {
    bd_loss_kernel<1><<<(unsigned)blocks, kLossThreads, 0, st>>>(P);
  } else if (K <= 256) {
    bd_loss_kernel<2><<<(unsigned)blocks, kLossThreads, 0, st>>>(P);
  } else if (K <= 512) {
    bd_loss_kernel<4><<<(unsigned)blocks, kLossThreads, 0, st>>>(P);
  } else {
    bd_loss_kernel<8><<<(unsigned)blocks, kLossThreads, 0, st>>>(P);
  }
  BDP_CUDA_CHECK_LAUNCH("bd_loss_kernel");
  return BDP_OK;
}

extern "C" int bdp_expected_pose_loss(const float* logits, int64_t B, int64_t K, int64_t ld_logits,
                                      const float* delta, int per_bin, int ndim, const float* keys,
                                      const float* target, int pose_mode, float* rows,
                                      float* grad_logits, float* grad_delta, void* stream) {
  BDP_REQUIRE(B > 0 && K > 0 && ld_logits >= K, "expected_pose_loss: bad sizes");
  BDP_REQUIRE(logits && delta && keys && target && rows && grad_logits && grad_delta,
              "expected_pose_loss: NULL buffer");
  BDP_REQUIRE((pose_mode == BDP_POSE_GEODESIC_AA && ndim == 3) ||
                  (pose_mode == BDP_POSE_GEODESIC_Q && ndim == 4),
              "expected_pose_loss: axis-angle (ndim 3) or quaternion (ndim 4) geodesic loss only");
  ExpParams P;
  P.logits = logits; P.B = B; P.K = K; P.ld = ld_logits; P.delta = delta; P.per_bin = per_bin ? 1 : 0;
  P.nd = ndim; P.keys = keys; P.target = target; P.mode = pose_mode; P.rows = rows;
  P.g_logits = grad_logits; P.g_delta = grad_delta;
  const unsigned blocks = (unsigned)((B + 3) / 4);
  expected_pose_kernel<<<blocks, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(P);
  BDP_CUDA_CHECK_LAUNCH("expected_pose_kernel");
  return BDP_OK;
}
