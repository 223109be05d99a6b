// Peak picking on the device: the k highest local maxima of each row of a [rows, n] array.
//
// Replaces, for large periodograms, the host-side `find_peaks` + `pmax` chain behind
// `FSeries.period_at_highest_peak` / `psort_by_peak` (reference src/periodicity/core.py:283-317,
// 938-955), which calls `scipy.signal.find_peaks(values, prominence=0.0)`.  Peak definition
// reproduced from scipy's `_local_maxima_1d`: sample j is a peak iff it is the midpoint
// (left + right) // 2 of a maximal run of equal values values[left..right] with
// values[left-1] < values[left] and values[right+1] < values[right]; the first and the last
// sample are never peaks; comparisons with NaN are false.  With prominence = 0.0 every local
// maximum qualifies, so "k highest peaks" = the k largest such samples, ties broken by the
// lower index (np.nanargmax picks the first occurrence).
//
// Kernels: peaks_block_kernel (each block scans PEAK_ITEMS consecutive samples of one row, collects
// its peaks in shared memory and extracts its k best by repeated block arg-max),
// peaks_merge_kernel (one block per row merges the per-block candidates).
#include "pdc_common.cuh"

namespace pdc {

constexpr int PEAK_THREADS = 256;
constexpr int PEAK_PER_THREAD = 16;
constexpr int PEAK_ITEMS = PEAK_THREADS * PEAK_PER_THREAD;  // samples per block
constexpr int PEAK_KMAX = 64;

// k rounds of block arg-max over a candidate list; winners are written out and removed.
__device__ void select_topk(double* cv, long long* ci, int count, int k, double* out_v, long long* out_i,
                            double* sv, long long* si, long long* win) {
  for (int r = 0; r < k; ++r) {
    double bv = 0.0;
    long long bi = -1;
    int bslot = -1;
    for (int c = threadIdx.x; c < count; c += blockDim.x) {
      const long long i = ci[c];
      if (i >= 0 && better<+1>(cv[c], i, bv, bi)) { bv = cv[c]; bi = i; bslot = c; }
    }
    double v = bv;
    long long i = bi;
    block_argext<+1>(v, i, sv, si);
    if (threadIdx.x == 0) {
      *win = i;
      out_v[r] = i >= 0 ? v : nan("");
      out_i[r] = i;
    }
    __syncthreads();
    if (bslot >= 0 && bi == *win) ci[bslot] = -1;  // remove the winner (indices are unique)
    __syncthreads();
  }
}

__global__ void __launch_bounds__(PEAK_THREADS)
peaks_block_kernel(const double* __restrict__ values, long long n, int k, double* __restrict__ cand_v,
                   long long* __restrict__ cand_i) {
  __shared__ double cv[PEAK_ITEMS / 2 + 1];
  __shared__ long long ci[PEAK_ITEMS / 2 + 1];
  __shared__ int count;
  __shared__ double sv[32];
  __shared__ long long si[32];
  __shared__ long long win;
  const int row = blockIdx.y;
  const double* v = values + (long long)row * n;
  if (threadIdx.x == 0) count = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * PEAK_ITEMS;
  for (int e = 0; e < PEAK_PER_THREAD; ++e) {
    const long long j = base + (long long)e * PEAK_THREADS + threadIdx.x;  // coalesced
    if (j >= 1 && j < n - 1) {
      const double x = v[j];
      if (v[j - 1] < x) {  // left edge of a (possibly one-sample) plateau
        long long r = j;
        while (r + 1 < n && v[r + 1] == x) ++r;
        if (r + 1 < n && v[r + 1] < x) {
          const int slot = atomicAdd(&count, 1);  // at most every other sample starts a plateau
          cv[slot] = x;
          ci[slot] = (j + r) / 2;
        }
      }
    }
  }
  __syncthreads();
  double* ov = cand_v + ((long long)row * gridDim.x + blockIdx.x) * k;
  long long* oi = cand_i + ((long long)row * gridDim.x + blockIdx.x) * k;
  select_topk(cv, ci, count, k, ov, oi, sv, si, &win);
}

__global__ void __launch_bounds__(PEAK_THREADS)
peaks_merge_kernel(double* __restrict__ cand_v, long long* __restrict__ cand_i, int per_row, int k,
                   double* __restrict__ out_v, long long* __restrict__ out_i) {
  __shared__ double sv[32];
  __shared__ long long si[32];
  __shared__ long long win;
  const int row = blockIdx.x;
  select_topk(cand_v + (long long)row * per_row, cand_i + (long long)row * per_row, per_row, k,
              out_v + (long long)row * k, out_i + (long long)row * k, sv, si, &win);
}

// Half-maximum crossings of given peaks: `FSeries.periods_at_half_max` (reference core.py:957-972).
// For a peak at index p the level is half = v[p] - height / 2 with height = v[p] (or the caller's per-peak
// height, e.g. the prominence: use_prominence=True, core.py:960-963); d[j] = v[j] - half.
//   left  = the LAST  j in [0, p - 2]      with signbit(d[j]) != signbit(d[j + 1])  (core.py:967, find_zero_crossings :362)
//   right = the FIRST j in [p, n - 2]      with signbit(d[j]) != signbit(d[j + 1])  (core.py:968)
// -1 if there is none.  One warp per peak scans outwards 32 samples at a time.
__device__ __forceinline__ bool sl_signbit(double x) { return (__double2hiint(x) >> 31) != 0; }

__global__ void __launch_bounds__(128)
peaks_halfmax_kernel(const double* __restrict__ values, long long n, int k, const long long* __restrict__ peak_idx,
                     const double* __restrict__ height, long long total, long long* __restrict__ left_out,
                     long long* __restrict__ right_out) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= total) return;
  const long long row = w / k;
  const double* v = values + row * n;
  const long long p = peak_idx[w];
  long long left = -1, right = -1;
  if (p >= 0 && p < n) {
    const double half = v[p] - (height ? height[w] : v[p]) / 2;
    // rightwards: pairs (j, j + 1), j = p .. n - 2
    for (long long j0 = p; j0 <= n - 2; j0 += 32) {
      const long long j = j0 + lane;
      const bool hit = j <= n - 2 && sl_signbit(v[j] - half) != sl_signbit(v[j + 1] - half);
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) { right = j0 + (__ffs(m) - 1); break; }
    }
    // leftwards: pairs (j, j + 1), j = p - 2 .. 0
    for (long long j0 = p - 2; j0 >= 0; j0 -= 32) {
      const long long j = j0 - lane;
      const bool hit = j >= 0 && sl_signbit(v[j] - half) != sl_signbit(v[j + 1] - half);
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) { left = j0 - (__ffs(m) - 1); break; }
    }
  }
  if (lane == 0) {
    left_out[w] = left;
    right_out[w] = right;
  }
}

int peaks_halfmax_run(pdc_ctx* ctx, const double* values, int64_t rows, int64_t n, int k, const int64_t* peak_idx,
                      const double* height, int64_t* left_out, int64_t* right_out, cudaStream_t st) {
  if (rows < 1 || n < 1 || k < 1) { set_error("pdc_peaks_halfmax: empty input"); return PDC_EINVAL; }
  const long long total = (long long)rows * k;
  const long long blocks = (total * 32 + 127) / 128;
  if (blocks > 0x7fffffffLL) { set_error("pdc_peaks_halfmax: too many peaks for one call"); return PDC_EINVAL; }
  peaks_halfmax_kernel<<<(unsigned)blocks, 128, 0, st>>>(values, (long long)n, k, (const long long*)peak_idx, height,
                                                         total, (long long*)left_out, (long long*)right_out);
  PDC_CUDA(cudaGetLastError());
  ctx->launches++;
  return PDC_OK;
}

int peaks_run(pdc_ctx* ctx, const double* values, int64_t rows, int64_t n, int k, int64_t* idx_out,
              double* val_out, cudaStream_t st) {
  if (rows < 1 || n < 1) { set_error("pdc_peaks_topk: empty input"); return PDC_EINVAL; }
  if (k < 1 || k > PEAK_KMAX) { set_error("pdc_peaks_topk: k must be in 1..%d", PEAK_KMAX); return PDC_EINVAL; }
  if (rows > 65535) { set_error("pdc_peaks_topk: at most 65535 rows per call"); return PDC_EINVAL; }
  const long long nblk = (n + PEAK_ITEMS - 1) / PEAK_ITEMS;
  const size_t ncand = (size_t)rows * nblk * k;
  ScratchScope scratch(ctx, st);
  PDC_TRY(scratch.acquire());
  PDC_TRY(ctx->peak_cand.reserve(ncand * (sizeof(double) + sizeof(long long))));
  double* cand_v = ctx->peak_cand.as<double>();
  long long* cand_i = reinterpret_cast<long long*>(cand_v + ncand);
  dim3 grid((unsigned)nblk, (unsigned)rows);
  peaks_block_kernel<<<grid, PEAK_THREADS, 0, st>>>(values, n, k, cand_v, cand_i);
  PDC_CUDA(cudaGetLastError());
  ctx->launches++;
  peaks_merge_kernel<<<(unsigned)rows, PEAK_THREADS, 0, st>>>(cand_v, cand_i, (int)(nblk * k), k, val_out,
                                                              (long long*)idx_out);
  PDC_CUDA(cudaGetLastError());
  ctx->launches++;
  PDC_TRY(scratch.release());
  return PDC_OK;
}

}  // namespace pdc
