// Rows of a k-means fit in key-grid cell order (learnKmeansDictionary.py:41-42: one data set, up to
// 300 Lloyd iterations over it).
//
// The E+M kernel (query.cu) is bound by L1 / shared-memory wavefronts, not by DRAM: with rows in
// arbitrary order the 32 lanes of a warp read 32 different cell records (32 tag look-ups per load),
// then 32 different candidate records from shared memory (bank conflicts), and run the candidate loop
// for the LONGEST list among them.  The fit's grid geometry is fixed (bdp_keygrid_prepare), so the rows
// can be sorted by fine cell ONCE: a warp's 64 rows then sit in one or two cells — one cell-record
// line, broadcast candidate records, uniform list lengths — and the rows that change cluster sit
// together, which the M-step's warp-aggregated atomics (query.cu) turn into a handful of additions.
// Labels are computed per row and the cluster sums are integers, so the order of the rows changes
// neither; the caller scatters the labels back through the permutation at the end of the fit.
//
// The sort itself is CUB's radix sort on (cell id, row index) pairs — a library call outside the
// iteration loop — followed by a row gather.
#include <cub/device/device_radix_sort.cuh>

#include "assign_common.cuh"

using namespace bdp_assign;

void bdpi_keygrid_parts(void* grid, int K, int d, void** hdr, void** cf32);

namespace {

// cell id of every row (the query kernel's point -> cell map, expression for expression) and the row
// indices; rows outside the grid get the id one past the last cell.
template <int D>
__global__ void __launch_bounds__(256) cell_keys_kernel(const double* __restrict__ x, int64_t N,
                                                        const GridHdr* __restrict__ hdr,
                                                        unsigned* __restrict__ keys,
                                                        unsigned* __restrict__ idx) {
  float g_inv[D], g_off[D];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    g_inv[k] = (float)hdr->inv_cell[k];
    g_off[k] = (float)(-hdr->origin[k] * hdr->inv_cell[k]);
  }
  const int G = hdr->G;
  const bool on = hdr->enabled != 0;
  unsigned n_fine = 1;
#pragma unroll
  for (int k = 0; k < D; ++k) n_fine *= (unsigned)G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    bool ok = on;
    unsigned cidx = 0, mul = 1;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const float xf = (float)x[i * D + k];               // (L1 serves the row's other coordinates)
      const float t = fmaf(xf, g_inv[k], g_off[k]);
      const int ck = __float2int_rd(t);
      ok = ok && ((unsigned)ck < (unsigned)G);
      cidx += (unsigned)ck * mul;
      mul *= (unsigned)G;
    }
    keys[i] = ok ? cidx : n_fine;
    idx[i] = (unsigned)i;
  }
}

// Occupied coarse cells from the SORTED cell ids: a row marks its coarse cell only where the coarse id
// differs from the previous row's — a few ten thousand stores instead of one per row (ten million
// stores to a few thousand words serialise in the L2: 350 us for the pass).
template <int D>
__global__ void __launch_bounds__(256) occ_from_sorted_kernel(const unsigned* __restrict__ keys, int64_t N,
                                                              const GridHdr* __restrict__ hdr,
                                                              int* __restrict__ occ) {
  const int G = hdr->G, Gc = G / 4;
  unsigned n_fine = 1;
#pragma unroll
  for (int k = 0; k < D; ++k) n_fine *= (unsigned)G;
  auto coarse_of = [&](unsigned key) -> int {
    if (key >= n_fine) return -1;
    int c = 0, cmul = 1;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      c += (int)((key % (unsigned)G) >> 2) * cmul;
      key /= (unsigned)G;
      cmul *= Gc;
    }
    return c;
  };
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = coarse_of(keys[i]);
    const int p = i > 0 ? coarse_of(keys[i - 1]) : -2;
    if (c >= 0 && c != p) occ[c] = 1;
  }
}

template <int D>
__global__ void __launch_bounds__(256) gather_rows_kernel(const double* __restrict__ x, int64_t N,
                                                          const unsigned* __restrict__ perm,
                                                          double* __restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const double* src = x + (int64_t)perm[i] * D;
    double v[D];
#pragma unroll
    for (int k = 0; k < D; ++k) v[k] = __ldg(src + k);
#pragma unroll
    for (int k = 0; k < D; ++k) __stcs(y + i * D + k, v[k]);
  }
}

__global__ void __launch_bounds__(256) scatter_i32_kernel(const int* __restrict__ src,
                                                          const unsigned* __restrict__ perm, int64_t N,
                                                          int* __restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    dst[perm[i]] = __ldcs(src + i);
}

int key_bits(int K, int d) {
  const int64_t cells = bdp_keygrid_coarse_cells(K, d) * (d == 3 ? 64 : 256) + 1;   // fine cells + "outside"
  int b = 1;
  while (((int64_t)1 << b) < cells) ++b;
  return b;
}

size_t a256(size_t b) { return (b + 255) & ~(size_t)255; }

size_t cub_temp_bytes(int64_t N, int bits) {
  size_t t = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t, (const unsigned*)nullptr, (unsigned*)nullptr,
                                  (const unsigned*)nullptr, (unsigned*)nullptr, (int)N, 0, bits);
  return t;
}

unsigned grid_for(int64_t N) {
  const int64_t cap = (int64_t)bdp_num_sms() * 16;
  int64_t b = (N + 255) / 256;
  if (b > cap) b = cap;
  return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" int64_t bdp_cellsort_workspace_bytes(int64_t N, int K, int d) {
  if (N < 0 || N >= ((int64_t)1 << 31) || bdp_keygrid_coarse_cells(K, d) < 0) return -1;
  return (int64_t)(3 * a256((size_t)N * 4) + a256(cub_temp_bytes(N, key_bits(K, d))));
}

extern "C" int bdp_cellsort(const double* x, int64_t N, int d, int K, void* grid, int64_t grid_bytes,
                            int32_t* occ, void* workspace, int64_t workspace_bytes, int32_t* perm,
                            double* x_sorted, void* stream) {
  BDP_REQUIRE(N >= 0 && N < ((int64_t)1 << 31), "cellsort: N out of range");
  BDP_REQUIRE(d == 3 || d == 4, "cellsort: d must be 3 or 4 (got %d)", d);
  BDP_REQUIRE(grid != nullptr && bdp_keygrid_bytes(K, d) > 0 && grid_bytes >= bdp_keygrid_bytes(K, d),
              "cellsort: grid buffer missing or too small");
  if (N == 0) return BDP_OK;
  BDP_REQUIRE(x && workspace && perm && x_sorted, "cellsort: NULL buffer");
  BDP_REQUIRE(workspace_bytes >= bdp_cellsort_workspace_bytes(N, K, d), "cellsort: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  void *h, *c;
  bdpi_keygrid_parts(grid, K, d, &h, &c);
  const GridHdr* hdr = reinterpret_cast<const GridHdr*>(h);
  unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
  const size_t nb = a256((size_t)N * 4);
  unsigned* keys_in = reinterpret_cast<unsigned*>(w);
  unsigned* keys_out = reinterpret_cast<unsigned*>(w + nb);
  unsigned* idx_in = reinterpret_cast<unsigned*>(w + 2 * nb);
  void* temp = w + 3 * nb;
  const int bits = key_bits(K, d);
  size_t temp_bytes = cub_temp_bytes(N, bits);
  if (d == 3) cell_keys_kernel<3><<<grid_for(N), 256, 0, st>>>(x, N, hdr, keys_in, idx_in);
  else cell_keys_kernel<4><<<grid_for(N), 256, 0, st>>>(x, N, hdr, keys_in, idx_in);
  BDP_CUDA_CHECK_LAUNCH("cell_keys_kernel");
  BDP_CUDA_CALL(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, idx_in,
                                                reinterpret_cast<unsigned*>(perm), (int)N, 0, bits, st));
  if (d == 3) gather_rows_kernel<3><<<grid_for(N), 256, 0, st>>>(x, N, reinterpret_cast<unsigned*>(perm), x_sorted);
  else gather_rows_kernel<4><<<grid_for(N), 256, 0, st>>>(x, N, reinterpret_cast<unsigned*>(perm), x_sorted);
  BDP_CUDA_CHECK_LAUNCH("gather_rows_kernel");
  if (occ != nullptr) {
    if (d == 3) occ_from_sorted_kernel<3><<<grid_for(N), 256, 0, st>>>(keys_out, N, hdr, occ);
    else occ_from_sorted_kernel<4><<<grid_for(N), 256, 0, st>>>(keys_out, N, hdr, occ);
    BDP_CUDA_CHECK_LAUNCH("occ_from_sorted_kernel");
  }
  return BDP_OK;
}

extern "C" int bdp_scatter_i32(const int32_t* src, const int32_t* perm, int64_t N, int32_t* dst,
                               void* stream) {
  BDP_REQUIRE(N >= 0, "scatter_i32: N < 0");
  if (N == 0) return BDP_OK;
  BDP_REQUIRE(src && perm && dst, "scatter_i32: NULL buffer");
  scatter_i32_kernel<<<grid_for(N), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, reinterpret_cast<const unsigned*>(perm), N, dst);
  BDP_CUDA_CHECK_LAUNCH("scatter_i32_kernel");
  return BDP_OK;
}
