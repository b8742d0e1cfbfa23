// Multi-sweep accumulation on the GPU (SURVEY §8 f rank 2): the step right before the hot path.
// Replaces, for point clouds already in device memory, the reference's per-frame NumPy work
//   det3d/datasets/pipelines/loading.py:37-46   remove_close (|x| < r and |y| < r, sensor frame)
//   det3d/datasets/pipelines/loading.py:49-61   read_sweep: remove_close, 4x4 transform (float64), time lag
//   det3d/datasets/pipelines/loading.py:118-141 key frame + sweeps concatenated, time channel appended
// Order preserving (key frame first, then the sweeps in the order given, points in file order): keep bits per
// point (one ballot per warp = one mask word, no atomics), popcount scan (mask_scan), ranked scatter.  The output
// row base and the running total are device scalars, so the frames of a batch chain without a host sync and the
// result feeds pn_pillarize directly.
#include "common.cuh"
#include "mask_scan.cuh"

namespace {

constexpr int kMaxSweeps = 16;

struct SweepParams {
  int n_sweeps;
  int off[kMaxSweeps + 1];      // first raw row of each sweep; off[n_sweeps] = n_raw
  double T[kMaxSweeps][12];     // row-major 3x4 (rotation | translation)
  int has_T[kMaxSweeps];
  float lag[kMaxSweeps];
  float radius;
  int in_dim, n_feat;           // raw row width, leading columns kept (x,y,z,+features); output width n_feat+1
};

__device__ __forceinline__ int sweep_of(const SweepParams& P, int i) {
  int s = 0;
#pragma unroll 1
  for (int k = 1; k < P.n_sweeps; ++k) s = (i >= P.off[k]) ? k : s;
  return s;
}

__global__ void __launch_bounds__(256)
k_sweep_mark(const __grid_constant__ SweepParams P, const float* __restrict__ raw, int n_raw, long long n_words,
             uint32_t* __restrict__ words) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((i >> 5) >= n_words) return;   // whole warps only: every lane of a live word takes part in the ballot
  bool keep = false;
  if (i < n_raw) {
    const int s = sweep_of(P, (int)i);
    const float x = __ldg(raw + i * P.in_dim), y = __ldg(raw + i * P.in_dim + 1);
    // the key frame (sweep 0) is taken whole; sweeps drop the points close to the sensor (loading.py:51-52)
    keep = s == 0 || !(fabsf(x) < P.radius && fabsf(y) < P.radius);
  }
  const unsigned bits = __ballot_sync(0xffffffffu, keep);
  if ((threadIdx.x & 31) == 0) words[i >> 5] = bits;
}

__global__ void __launch_bounds__(256)
k_sweep_emit(const __grid_constant__ SweepParams P, const float* __restrict__ raw, int n_raw,
             const uint32_t* __restrict__ words, const int* __restrict__ prefix, const int* __restrict__ kept,
             const int* __restrict__ out_base, float* __restrict__ out, int out_cap, int* __restrict__ n_total) {
  const int base = out_base ? *out_base : 0;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *n_total = min(base + *kept, out_cap);
  if (i >= n_raw) return;
  const uint32_t w = words[i >> 5];
  const int bit = i & 31;
  if (!((w >> bit) & 1u)) return;
  const int row = base + prefix[i >> 5] + __popc(w & ((1u << bit) - 1u));
  if (row >= out_cap) return;
  const int s = sweep_of(P, i);
  const float* q = raw + (long long)i * P.in_dim;
  float* o = out + (long long)row * (P.n_feat + 1);
  float x = q[0], y = q[1], z = q[2];
  if (P.has_T[s]) {
    // float64 product rounded once to fp32, as the assignment into the float32 array does (loading.py:55-58)
    const double* T = P.T[s];
    const double dx = x, dy = y, dz = z;
    x = (float)(T[0] * dx + T[1] * dy + T[2] * dz + T[3]);
    y = (float)(T[4] * dx + T[5] * dy + T[6] * dz + T[7]);
    z = (float)(T[8] * dx + T[9] * dy + T[10] * dz + T[11]);
  }
  o[0] = x; o[1] = y; o[2] = z;
  for (int k = 3; k < P.n_feat; ++k) o[k] = q[k];
  o[P.n_feat] = P.lag[s];
}

}  // namespace

extern "C" {

size_t pn_merge_sweeps_scratch_bytes(int n_raw) {
  const long long nw = pn_detail::n_words(n_raw > 0 ? n_raw : 1);
  // mask words + prefix + kept count + scan scratch
  return (size_t)nw * 8 + 256 + pn_detail::scan_scratch_bytes(nw);
}

int pn_merge_sweeps(const float* raw, int in_dim, int n_feat, const int* sweep_offsets, int n_sweeps,
                    const double* transforms, const float* time_lag, float min_distance, const int* out_base,
                    float* out, int out_cap, int* n_total, void* scratch, size_t scratch_bytes,
                    pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(n_sweeps >= 1 && n_sweeps <= kMaxSweeps && sweep_offsets && time_lag && out && n_total && scratch);
  PN_REQUIRE(in_dim >= 3 && n_feat >= 3 && n_feat <= in_dim && out_cap >= 0);
  const int n_raw = sweep_offsets[n_sweeps];
  PN_REQUIRE(n_raw >= 0 && sweep_offsets[0] == 0);
  PN_REQUIRE(n_raw == 0 || raw);
  if (scratch_bytes < pn_merge_sweeps_scratch_bytes(n_raw)) return PN_ERR_WORKSPACE;
  SweepParams P;
  P.n_sweeps = n_sweeps;
  for (int k = 0; k <= kMaxSweeps; ++k) P.off[k] = k <= n_sweeps ? sweep_offsets[k] : n_raw;
  for (int k = 0; k < kMaxSweeps; ++k) {
    const bool live = transforms && k < n_sweeps;
    for (int j = 0; j < 12; ++j) P.T[k][j] = live ? transforms[k * 12 + j] : 0.0;
    P.has_T[k] = (live && transforms[k * 12] == transforms[k * 12]) ? 1 : 0;   // a row of NaNs = no transform
    P.lag[k] = k < n_sweeps ? time_lag[k] : 0.f;
  }
  P.radius = min_distance;
  P.in_dim = in_dim;
  P.n_feat = n_feat;
  const long long nw = pn_detail::n_words(n_raw > 0 ? n_raw : 1);
  uint32_t* words = reinterpret_cast<uint32_t*>(scratch);
  int* prefix = reinterpret_cast<int*>(words + nw);
  int* kept = prefix + nw;
  void* scan_scratch = reinterpret_cast<char*>(scratch) + nw * 8 + 256;
  const unsigned blocks = (unsigned)PN_DIVUP(nw * 32, 256ll);
  k_sweep_mark<<<blocks, 256, 0, stream>>>(P, raw, n_raw, nw, words);
  PN_CHECK_LAUNCH();
  int rc = pn_detail::mask_scan_emit(words, prefix, nw, 1 << 30, 1, nullptr, 0, kept, scan_scratch,
                                     pn_detail::scan_scratch_bytes(nw), stream);
  if (rc != PN_OK) return rc;
  k_sweep_emit<<<blocks, 256, 0, stream>>>(P, raw, n_raw, words, prefix, kept, out_base, out, out_cap, n_total);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

}  // extern "C"
