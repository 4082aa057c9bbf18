// Definitions shared by assign.cu (brute-force scan, key-grid build, entry points) and query.cu (the
// pruned query kernel).  Internal to libbdpose.so.
#pragma once
#include "common.cuh"

namespace bdp_assign {

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

struct GridHdr;
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
  int incremental;         // Lloyd: acc holds the sums of the PREVIOUS labels; only rotations whose
                           // label changed move (subtracted from the old cluster, added to the new)
  float err_coef;          // 2^-24 * 2(D+5) * safety
  // key grid (candidate pruning); NULL for the brute-force kernel
  const struct GridHdr* ghdr;
  const uint4* gfine;            // [n_fine] 16-byte cell records
  const unsigned short* gside;   // [n_side][32] overflow lists of the long cells
  // device flag of a k-means run (bdp_kmeans_run): non-zero = the fit has stopped, do nothing
  const int* stop;
  // multi-GPU build: the grid is complete once every rank has raised gflags[r] to gflag_value
  const unsigned long long* gflags;
  int gworld;
  unsigned long long gflag_value;
};

// Fine cell record (16 bytes = 8 halfwords): h0 = number of candidate keys (0xFFFF: overflow, the
// query scans the dictionary), then the ascending key ids.  Up to 7 keys sit inline (97 % of the
// cells); a longer list keeps keys 1..6 inline, h7 = slot of a 64-byte side record that holds keys
// 7..31.  16.7 MB -> 4.2 MB for the 64^3 grid of a K=1000 dictionary: two cells per 32-byte sector,
// and small enough to be broadcast to the peers of a multi-GPU build.
constexpr int kFineInline = 7;
constexpr int kSideWidth = 32;               // halfwords per side record
constexpr int kGridCap = 31;                 // ids per fine record  (u16 count + 31 u16 ids = 64 B)
constexpr int kGridMaxK = 4096;              // fp32 screening records of the whole dictionary in smem
constexpr unsigned kGridOverflow = 0xFFFFu;
// Cell boxes are grown by this fraction of a cell on every side before the bounds are taken, which
// covers the rounding of the point -> cell mapping in the query (fp32 for fp32 rotations: the cell
// coordinate is off by < 3e-5 cells; fp64: < 1e-13).
constexpr double kBoxEps = 1e-3;

struct GridHdr {
  double origin[4];
  double cell[4];
  double inv_cell[4];
  int G;            // fine cells per dimension (multiple of 4)
  int enabled;      // 0: degenerate dictionary -> every point takes the slow path
  unsigned side_next;   // side records handed out by this build (reset with the header)
  unsigned ticket;      // blocks of the fine kernel that have finished
  int pad[4];
  float origin32[4];
  float inv_cell32[4];
};
static_assert(sizeof(GridHdr) == 160, "GridHdr layout");

float screen_err_coef(int D);

}  // namespace bdp_assign

namespace bdp_assign {
// system-scope accesses for the peer-memory handshakes
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
}  // namespace bdp_assign

// multi-GPU key-grid build (kmeans.cu -> assign.cu): this process's addresses of every rank's grid
// buffer and grid-flag array
struct bdpi_grid_peers {
  void* grid[BDP_KMEANS_MAX_RANKS];
  unsigned long long* gflags[BDP_KMEANS_MAX_RANKS];
  int world, rank;
  unsigned long long flag_value;
};

// the pruned query (query.cu): assign (labels + residual) or Lloyd E+M step through a built key grid
int bdpi_query_grid(const bdp_assign::AssignParams& P, int x_dtype, int d, bool lloyd, cudaStream_t st);
