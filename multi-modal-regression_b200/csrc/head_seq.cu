// (a) Whole-head launch sequences: one C call runs every kernel of the forward (or backward) pass of
// a stack of H three-layer heads, so the host cost of a training step is two FFI calls instead of
// ~25 (at the reference's batch sizes the step is ~20 kernels of 10-80 us each: per-launch Python
// overhead used to dominate the GPU time 5:1).
//
// Reference: OneBinDeltaModel.forward (binDeltaModels.py:112-121) and its autograd backward;
// bin_3layer / res_3layer (binDeltaModels.py:62-91).
#include "common.cuh"

namespace {

struct Dims {
  int64_t B, F1, F2;
  int H, N0, N1, N2;
};

Dims dims_of(const bdp_head_desc* d, int64_t B) {
  Dims D;
  D.B = B; D.H = d->H; D.N0 = d->N0; D.N1 = d->N1; D.N2 = d->N2;
  D.F1 = (int64_t)d->H * d->N1;
  D.F2 = (int64_t)d->H * d->N2;
  return D;
}

int check_desc(const bdp_head_desc* d, int64_t B, const char* who) {
  BDP_REQUIRE(d != nullptr, "%s: NULL descriptor", who);
  BDP_REQUIRE(d->H > 0 && d->N0 > 0 && d->N1 > 0 && d->N2 > 0, "%s: bad head sizes", who);
  BDP_REQUIRE(d->N0 % 4 == 0 && d->N1 % 4 == 0 && d->N2 % 4 == 0,
              "%s: layer widths must be multiples of 4 (got %d, %d, %d)", who, d->N0, d->N1, d->N2);
  BDP_REQUIRE(d->n_groups >= 1 && d->n_groups <= BDP_HEAD_MAX_GROUPS, "%s: n_groups %d", who,
              d->n_groups);
  int heads = 0;
  for (int g = 0; g < d->n_groups; ++g) {
    BDP_REQUIRE(d->group_heads[g] == d->group_heads[0], "%s: every fc3 group must have the same "
                "number of heads (they share the mixing weights)", who);
    BDP_REQUIRE(d->group_out[g] > 0 && d->w3[g] && d->b3[g], "%s: fc3 group %d incomplete", who, g);
    heads += d->group_heads[g];
  }
  BDP_REQUIRE(heads == d->H, "%s: fc3 groups cover %d heads of %d", who, heads, d->H);
  BDP_REQUIRE(B >= 1 && B <= 65535, "%s: batch %lld out of range", who, (long long)B);
  BDP_REQUIRE(d->w1 && d->g1 && d->be1 && d->w2 && d->g2 && d->be2, "%s: NULL parameter", who);
  return BDP_OK;
}

// saved-activation layout (floats): h1 | a1 | h2 | a2 | m1 | is1 | m2 | is2
struct Saved {
  float *h1, *a1, *h2, *a2, *m1, *is1, *m2, *is2;
};
Saved carve_saved(float* base, const Dims& D) {
  Saved s;
  s.h1 = base;
  s.a1 = s.h1 + D.B * D.F1;
  s.h2 = s.a1 + D.B * D.F1;
  s.a2 = s.h2 + D.B * D.F2;
  s.m1 = s.a2 + D.B * D.F2;
  s.is1 = s.m1 + D.F1;
  s.m2 = s.is1 + D.F1;
  s.is2 = s.m2 + D.F2;
  return s;
}

int dgrad_splits(const Dims& D) {
  const int n_tiles = (D.N0 + 255) / 256;
  int want = bdp_num_sms() / n_tiles;
  if (want < 1) want = 1;
  if (want > 64) want = 64;
  return bdp_gemm_tf32_splits(D.F1, want);
}

}  // namespace

#define BDP_TRY(expr)            \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != BDP_OK) return rc__; \
  } while (0)

extern "C" int64_t bdp_head_saved_floats(const bdp_head_desc* d, int64_t B) {
  if (!d || B < 1) return -1;
  const Dims D = dims_of(d, B);
  return 2 * D.B * D.F1 + 2 * D.B * D.F2 + 2 * D.F1 + 2 * D.F2;
}

extern "C" int64_t bdp_head_bwd_workspace_floats(const bdp_head_desc* d, int64_t B) {
  if (!d || B < 1) return -1;
  const Dims D = dims_of(d, B);
  const int64_t parts = (int64_t)dgrad_splits(D) * D.B * D.N0;
  const int64_t dmix_parts = (int64_t)d->n_groups * D.B * d->group_heads[0];
  return 2 * D.B * D.F1 + 2 * D.B * D.F2 + parts + dmix_parts;
}

extern "C" int bdp_head_forward(const bdp_head_desc* d, const float* x, const float* mix, int64_t B,
                                float* saved, float* const* y, void* stream) {
  BDP_TRY(check_desc(d, B, "head_forward"));
  BDP_REQUIRE(x && mix && saved && y, "head_forward: NULL buffer");
  BDP_REQUIRE(!d->training || B >= 2, "head_forward: training-mode BatchNorm needs a batch of >= 2");
  const Dims D = dims_of(d, B);
  const Saved s = carve_saved(saved, D);
  // fc1: H1 [B, H*N1] = X [B, N0] (lanes, K-major) x W1 [H*N1, N0] (streamed, K-major)
  BDP_TRY(bdp_gemm_tf32(x, 0, D.N0, 0, d->w1, 0, D.N0, 0, s.h1, 0, D.F1, 0, D.B, D.F1, D.N0, 1, 1, 0,
                        d->precise, stream));
  BDP_TRY(bdp_bn_relu_fwd(s.h1, D.F1, D.B, D.F1, d->g1, d->be1, d->rm1, d->rv1, s.m1, s.is1, d->eps,
                          d->momentum, d->training, s.a1, stream));
  // fc2 (grouped): H2_g [B, N2] = A1_g [B, N1] (columns g*N1.. of a1) x W2_g [N2, N1]
  BDP_TRY(bdp_gemm_tf32(s.a1, 0, D.F1, D.N1, d->w2, 0, D.N1, (int64_t)D.N2 * D.N1, s.h2, 0, D.F2,
                        D.N2, D.B, D.N2, D.N1, D.H, 1, 0, d->precise, stream));
  BDP_TRY(bdp_bn_relu_fwd(s.h2, D.F2, D.B, D.F2, d->g2, d->be2, d->rm2, d->rv2, s.m2, s.is2, d->eps,
                          d->momentum, d->training, s.a2, stream));
  int64_t off = 0;
  for (int g = 0; g < d->n_groups; ++g) {
    BDP_REQUIRE(y[g] != nullptr, "head_forward: y[%d] is NULL", g);
    BDP_TRY(bdp_head_fc3_fwd(s.a2 + off * D.N2, D.F2, d->w3[g], d->b3[g], mix, D.B,
                             d->group_heads[g], d->group_out[g], D.N2, y[g], stream));
    off += d->group_heads[g];
  }
  return BDP_OK;
}

extern "C" int bdp_head_backward(const bdp_head_desc* d, const float* x, const float* mix, int64_t B,
                                 const float* saved, const float* const* dy, float* ws, float* dw1,
                                 float* dg1, float* dbe1, float* dw2, float* dg2, float* dbe2,
                                 float* const* dw3, float* const* db3, float* dmix, float* dx,
                                 void* stream) {
  BDP_TRY(check_desc(d, B, "head_backward"));
  BDP_REQUIRE(x && mix && saved && dy && ws && dw1 && dg1 && dbe1 && dw2 && dg2 && dbe2 && dw3 && db3,
              "head_backward: NULL buffer");
  const Dims D = dims_of(d, B);
  const Saved s = carve_saved(const_cast<float*>(saved), D);
  float* da2 = ws;
  float* dh2 = da2 + D.B * D.F2;
  float* da1 = dh2 + D.B * D.F2;
  float* dh1 = da1 + D.B * D.F1;
  float* parts = dh1 + D.B * D.F1;
  const int splits = dgrad_splits(D);
  float* dmix_parts = parts + (int64_t)splits * D.B * D.N0;
  const int Hg = d->group_heads[0];
  // fc3 backward, group by group, each filling its column window of da2
  int64_t off = 0;
  for (int g = 0; g < d->n_groups; ++g) {
    BDP_REQUIRE(dy[g] && dw3[g] && db3[g], "head_backward: group %d buffer is NULL", g);
    float* dm = dmix ? (d->n_groups == 1 ? dmix : dmix_parts + (int64_t)g * D.B * Hg) : nullptr;
    BDP_TRY(bdp_head_fc3_bwd(dy[g], s.a2 + off * D.N2, D.F2, d->w3[g], d->b3[g], mix, D.B,
                             d->group_heads[g], d->group_out[g], D.N2, da2 + off * D.N2, dw3[g],
                             db3[g], dm, stream));
    off += d->group_heads[g];
  }
  if (dmix && d->n_groups > 1)
    BDP_TRY(bdp_sum_slabs(dmix_parts, D.B * Hg, d->n_groups, D.B * Hg, dmix, stream));
  // bn2 backward
  BDP_TRY(bdp_bn_relu_bwd(da2, s.a2, s.h2, d->g2, s.m2, s.is2, D.F2, D.B, D.F2, d->training, dh2, dg2,
                          dbe2, stream));
  // fc2 wgrad: dW2_g [N2, N1] = sum_b dH2_g[b, :]^T A1_g[b, :]   (both MN-major, K = batch)
  BDP_TRY(bdp_gemm_tf32(dh2, 1, D.F2, D.N2, s.a1, 1, D.F1, D.N1, dw2, 0, D.N1,
                        (int64_t)D.N2 * D.N1, D.N2, D.N1, D.B, D.H, 1, 0, d->precise, stream));
  // fc2 dgrad: dA1_g [B, N1] = dH2_g [B, N2] (K-major) x W2_g ([k = N2 rows, n = N1]: MN-major)
  BDP_TRY(bdp_gemm_tf32(dh2, 0, D.F2, D.N2, d->w2, 1, D.N1, (int64_t)D.N2 * D.N1, da1, 0, D.F1, D.N1,
                        D.B, D.N1, D.N2, D.H, 1, 0, d->precise, stream));
  // bn1 backward
  BDP_TRY(bdp_bn_relu_bwd(da1, s.a1, s.h1, d->g1, s.m1, s.is1, D.F1, D.B, D.F1, d->training, dh1, dg1,
                          dbe1, stream));
  // fc1 wgrad: dW1 [H*N1, N0] = dH1^T X   (both MN-major, K = batch)
  BDP_TRY(bdp_gemm_tf32(dh1, 1, D.F1, 0, x, 1, D.N0, 0, dw1, 0, D.N0, 0, D.F1, D.N0, D.B, 1, 1, 0,
                        d->precise, stream));
  if (dx) {
    // fc1 dgrad: dX [B, N0] = dH1 [B, H*N1] (K-major) x W1 ([k = H*N1 rows, n = N0]: MN-major), split-K
    BDP_TRY(bdp_gemm_tf32(dh1, 0, D.F1, 0, d->w1, 1, D.N0, 0, parts, 0, D.N0, 0, D.B, D.N0, D.F1, 1,
                          splits, D.B * D.N0, d->precise, stream));
    BDP_TRY(bdp_sum_slabs(parts, D.B * D.N0, splits, D.B * D.N0, dx, stream));
  }
  return BDP_OK;
}
