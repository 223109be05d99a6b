// Shared declarations for the periodicity_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <functional>
#include <utility>
#include <vector>

#include "../../include/periodicity_b200.h"

namespace pdc {

// ---- error plumbing -------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define PDC_CUDA(expr)                                                       \
  do {                                                                       \
    cudaError_t e__ = (expr);                                                \
    if (e__ != cudaSuccess) return ::pdc::cuda_fail(e__, #expr, __FILE__, __LINE__); \
  } while (0)

#define PDC_TRY(expr)                 \
  do {                                \
    int rc__ = (expr);                \
    if (rc__ != PDC_OK) return rc__;  \
  } while (0)

// ---- grow-only device / pinned buffers -------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes);  // keeps contents only if no growth is needed
  void release();
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinnedBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes);
  void release();
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace pdc

namespace pdc { struct MultiState; }

// The context: one device, one stream, scratch that only grows.  A multi-device ctx (pdc_ctx_create_multi) is the
// ctx of its first device plus `multi`: one child ctx and one host worker thread per further device (multi.cu).
struct pdc_ctx {
  pdc::MultiState* multi = nullptr;
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_fence = nullptr;                     // caller-stream -> scratch reuse fence
  cudaEvent_t ev_done = nullptr;                      // end of the previous call: the scratch buffers are shared, so
                                                      // a call on another stream first waits for it (scratch_acquire)
  int scratch_acquire(cudaStream_t st);               // order `st` after the previous call on this ctx
  int scratch_release(cudaStream_t st);               // mark the end of this call
  int64_t launches = 0;
  int gls_occ[32][2] = {};  // cached blocks/SM of each strip-kernel variant [geom][weighted]
  int gls_nsplit_override = 0;
  bool gls_geom_forced = false;  // PDC_GLS_GEOM given: no automatic small-problem geometry
  bool gls_three_term = true;  // env PDC_GLS_THREE_TERM=0 forces the rotation form of the strip step (tuning aid)
  int pdm_ppt_override = 0;  // env PDC_PDM_PPT=1|2 forces the trial periods per thread of pdm_hist_kernel (tuning aid)
  int last_gls_path = 0;       // hot kernel of the most recent GLS call: 0 gls_strip_kernel (or free-frequency), 1 gls_umma_kernel,
                               // 2 gls_umma_kernel with the precomputed fine operand, 3 gls_umma2_kernel (pairs of CTAs)
  int gls_umma = -1;           // tensor-core formulation of the GLS sums (gls_umma.cu): -1 automatic, 0 off, 1 whenever eligible
                               // (env PDC_GLS_UMMA)
  int gls_umma_chunk = 0;      // env PDC_GLS_UMMA_CHUNK: stages of 16 samples per TMEM accumulation run (default 4)
  int gls_umma_cg2 = -1;       // env PDC_GLS_UMMA_CG2: one long curve on pairs of CTAs (tcgen05 cta_group::2, gls_umma2.cu): -1 automatic
                               // (>= 16384 frequencies), 0 never, 1 from 4096 frequencies on
  int gls_umma_fine = -1;      // env PDC_GLS_UMMA_FINE: fine operand precomputed per curve: -1 automatic, 0 never, 1 whenever B == 1
  pdc::DevBuf umma_fine;       // its shared-memory images [2 types][stage][16 KB]
  int gls_umma_rzcomp = 1;     // env PDC_GLS_UMMA_RZCOMP=0: no compensation of the TMEM truncation bias (diagnostic)
  int gls_umma_dbg = 0;        // env PDC_GLS_UMMA_DBG: timing experiments (results are wrong when non-zero)
  int gls_umma_nsplit = 0;     // env PDC_GLS_UMMA_NSPLIT: sample splits of the tensor-core kernel (tuning aid)
  pdc::DevBuf umma_status;     // int[2]: word (call & 1) is set by the tensor-core kernel on a protocol time-out (the epilogue then
                               // writes NaN); every call clears the other word for the next one
  int64_t umma_calls = 0;
  int* umma_status_cur = nullptr;   // this call's word (what the epilogue reads)
  bool umma_status_clean = false;
  bool umma_prof_on = false;   // env PDC_GLS_UMMA_PROF=1
  pdc::DevBuf umma_prof;       // long long [jobs][4] clock stamps of the last gls_umma_kernel launch
  int64_t umma_prof_jobs = 0;
  int gls_geom = 0;  // index into kGlsGeoms (gls.cu); env PDC_GLS_GEOM overrides at ctx creation (tuning aid)

  // CUDA-event timing of the dominant kernel (GLS strip / PDM histogram), recorded on the
  // launching stream; pairs are resolved lazily so recording never synchronises.
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pending, ev_free;
  cudaEvent_t ev_cur_begin = nullptr, ev_cur_end = nullptr;
  double main_ms_total = 0.0, main_ms_last = -1.0;
  int64_t main_count = 0;
  int main_begin(cudaStream_t st);
  int main_end(cudaStream_t st);
  int main_resolve();

  // host-pointer entry points: device copies of the caller's arrays
  pdc::DevBuf in_a, in_b, in_c, in_d;
  pdc::DevBuf out_a;           // power / theta
  pdc::DevBuf out_small;       // argmax/max records
  pdc::PinnedBuf pin_small;    // pinned landing zone for the small records
  pdc::PinnedBuf pin_out;      // pinned staging of large results on their way to the caller's (pageable) buffer
  cudaEvent_t ev_chunk[8] = {};
  // large batches through the host entry point: uploads on a second stream, chunk c+1 travels while chunk c is computed
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_up[8] = {};
  size_t pipe_min_bytes = (size_t)32 << 20;  // env PDC_BATCH_PIPE_BYTES: smallest input size that is uploaded in chunks

  // GLS scratch
  pdc::DevBuf gls_curves;      // GlsCurve[B]
  pdc::DevBuf gls_part;        // GlsPart[B][G]: per-block partial statistics
  pdc::DevBuf gls_rec1;        // double2[n]  (t - tmin, frac(df (t - tmin)))
  pdc::DevBuf gls_rec2;        // float4[n]   (cos, sin of the per-index rotation, y or w*y, w)
  pdc::DevBuf glsm_y;          // float [groups][n][R]: scaled values of the shared-time series (glsm.cu)
  int glsm_occ[2] = {0, 0};    // cached blocks/SM of glsm_strip_kernel [weighted]
  pdc::DevBuf gls_low;         // gls.cu: fixed-point plane of the sub-cycle bins' FP64 sums [6][B*low_cap] (clean between calls,
                               // gls_low_dirty); glsm.cu: float64 [chunk][...] scratch -- glsm_run marks it dirty
  pdc::DevBuf gls_cnt;         // completion counters of the last-block-done reductions (gls.cu), self-resetting
  pdc::DevBuf partial;         // glsm.cu: float64 partial sums [nsplit][rows][units]; strlen.cu: sort scratch
  // Planes of partial sums shared by ALL sample splits of a call: gls.cu 64-bit fixed point [6][B*nf]; pdm.cu counts +
  // 64-bit fixed-point sums [m0][np]; ce.cu counts [cells][np].  The epilogue that reads a plane clears it, so a plane
  // is all zeros between calls; it is memset only when (re)allocated or when a call failed before its epilogue.
  pdc::DevBuf gls_plane, hist_plane;
  bool gls_plane_dirty = false, gls_low_dirty = false, hist_plane_dirty = false;
  pdc::DevBuf blockred;        // per-block (value, index) candidates
  pdc::PinnedBuf pin_meta;     // host staging for per-curve metadata

  pdc::DevBuf peak_cand;       // per-block peak candidates (peaks.cu)

  // PDM scratch
  pdc::DevBuf pdm_meta;        // PdmMeta
  pdc::DevBuf gl_acc;          // Gregory-Loredo: running ln sum_m O_m per trial period (ce.cu)
  pdc::DevBuf pdm_cnt;         // completion counters of the last-block-done reductions (pdm.cu), self-resetting
};

namespace pdc {

// Scope of one call's use of the ctx's shared scratch: orders `st` after the previous call on construction
// (acquire) and records the end-of-call event when the scope ends -- on the success path through release(), on
// ANY early error return through the destructor, so the ordering contract of the header (every call is ordered
// after the previous call on the same ctx) also holds after a failed call.
struct ScratchScope {
  pdc_ctx* ctx;
  cudaStream_t st;
  bool open = false;
  ScratchScope(pdc_ctx* c, cudaStream_t s) : ctx(c), st(s) {}
  int acquire() {
    int rc = ctx->scratch_acquire(st);
    open = rc == PDC_OK;
    return rc;
  }
  int release() {
    open = false;
    return ctx->scratch_release(st);
  }
  ~ScratchScope() {
    if (open) cudaEventRecord(ctx->ev_done, st);
  }
};

// statistic computed from the phase-bin histograms of pdm.cu
enum { PDC_STAT_PDM = 0, PDC_STAT_AOV = 1 };

// multi-device dispatch of the host entry points (multi.cu)
void multi_destroy(pdc_ctx* ctx);
int multi_device_count(const pdc_ctx* ctx);
int multi_gls(pdc_ctx* ctx, const double* t, const double* y, const double* w, int64_t n, double fmin, double df,
              int64_t j0, int64_t nf, unsigned flags, double psd_scale, double* power_out, int64_t* argmax_out,
              double* max_out);
int multi_gls_batch(pdc_ctx* ctx, const double* t, const double* y, const double* w, const int64_t* offsets, int64_t B,
                    const double* fmin, const double* df, int64_t nf, unsigned flags, const double* psd_scale,
                    double* power_out, int64_t* argmax_out, double* max_out);
int multi_gls_multi(pdc_ctx* ctx, const double* t, const double* Y, const double* w, int64_t n, int64_t S, double fmin,
                    double df, int64_t j0, int64_t nf, unsigned flags, double psd_scale, double* power_out,
                    int64_t* argmax_out, double* max_out);

// launchers implemented in gls.cu / pdm.cu; all device pointers, stream ordered
int gls_run(pdc_ctx* ctx, const double* t, const double* y, const double* w,
            const int64_t* offsets_host, int64_t B, const double* fmin_host, const double* df_host,
            int64_t j0, int64_t nf, unsigned flags, const double* psd_scale_host,
            double* power_out, int64_t* argmax_out, double* max_out, cudaStream_t stream,
            const pdc_fanout* fanout = nullptr,
            const double* freqs_dev = nullptr);  // non-NULL: evaluate at this device list of nf frequencies (pdc_gls_freqs);
                                                 // fmin_host / df_host must then be {0}, {0}

int glsm_run(pdc_ctx* ctx, const double* t, const double* Y, const double* w, int64_t n, int64_t S,
             double fmin, double df, int64_t j0, int64_t nf, unsigned flags, double psd_scale,
             double* power_out, int64_t* argmax_out, double* max_out, cudaStream_t stream);

int pdm_run(pdc_ctx* ctx, const double* t, const double* x, int64_t n, const double* periods,
            int64_t np, int nb, int nc, double* theta_out, int64_t* argmin_out, double* min_out,
            cudaStream_t stream, const pdc_fanout* fanout = nullptr, int64_t fan_offset = 0,
            int statistic = PDC_STAT_PDM);   // PDC_STAT_AOV: theta_out = AoV statistic, argmin/min_out = its arg-MAX / max

// sign: -1 arg-min (PDM, String Length, conditional entropy), +1 arg-max (AoV)
int multi_period_grid(pdc_ctx* ctx, int64_t n, const double* periods, int64_t np, int sign, double* out,
                      int64_t* arg_out, double* best_out,
                      const std::function<int(pdc_ctx*, const double*, int64_t, double*, int64_t*, double*)>& call);

int ce_run(pdc_ctx* ctx, const double* t, const double* x, int64_t n, const double* periods, int64_t np, int nphi,
           int nm, double* h_out, int64_t* argmin_out, double* min_out, cudaStream_t stream);

int gl_run(pdc_ctx* ctx, const double* t, int64_t n, const double* periods, int64_t np, int m_max, int nc,
           double* lnodds_out, int64_t* argmax_out, double* max_out, cudaStream_t stream);

int strlen_run(pdc_ctx* ctx, const double* t, const double* m, int64_t n, const double* periods, int64_t np,
               double* ell_out, int64_t* argmin_out, double* min_out, cudaStream_t stream);

int peaks_run(pdc_ctx* ctx, const double* values, int64_t rows, int64_t n, int k, int64_t* idx_out,
              double* val_out, cudaStream_t stream);

int peaks_halfmax_run(pdc_ctx* ctx, const double* values, int64_t rows, int64_t n, int k, const int64_t* peak_idx,
                      const double* height, int64_t* left_out, int64_t* right_out, cudaStream_t stream);

// ---- small device helpers ---------------------------------------------------
#ifdef __CUDACC__

// 1.5 * 2^52: adding it rounds a |x| < 2^51 double to the nearest integer.
__device__ __forceinline__ double round_magic(double x) {
  const double M = 6755399441055744.0;
  return __dadd_rn(__dadd_rn(x, M), -M);
}

// frac(a*b) in [-0.5, 0.5], using the exact product a*b = p + e (FMA residual) so
// the result carries the full precision of a and b even when a*b is ~1e7 cycles.
__device__ __forceinline__ double frac_of_product(double a, double b) {
  double p = __dmul_rn(a, b);
  double e = __fma_rn(a, b, -p);
  double r = __dadd_rn(p, -round_magic(p));
  return __dadd_rn(r, e);
}

template <class T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of doubles; result valid in every thread. `scratch` >= 33 doubles.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double s = lane < nw ? scratch[lane] : 0.0;
    s = warp_sum(s);
    if (lane == 0) scratch[32] = s;
  }
  __syncthreads();
  return scratch[32];
}

// NS sums and NM maxima at once with ONE pair of barriers (the statistics kernels reduce 5-8 quantities: one
// reduction each costs three barriers apiece).  Fixed xor-tree order: deterministic.  Results valid in thread 0.
// `scratch` holds 32 * (NS + NM) doubles.
template <int NS, int NM>
__device__ __forceinline__ void block_reduce_many(double (&s)[NS], double (&m)[NM], double* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int i = 0; i < NS; ++i) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
#pragma unroll
    for (int i = 0; i < NM; ++i) m[i] = fmax(m[i], __shfl_xor_sync(0xffffffffu, m[i], o));
  }
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NS; ++i) scratch[i * 32 + wid] = s[i];
#pragma unroll
    for (int i = 0; i < NM; ++i) scratch[(NS + i) * 32 + wid] = m[i];
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int i = 0; i < NS; ++i) s[i] = lane < nw ? scratch[i * 32 + lane] : 0.0;
#pragma unroll
    for (int i = 0; i < NM; ++i) m[i] = lane < nw ? scratch[(NS + i) * 32 + lane] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int i = 0; i < NS; ++i) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
#pragma unroll
      for (int i = 0; i < NM; ++i) m[i] = fmax(m[i], __shfl_xor_sync(0xffffffffu, m[i], o));
    }
  }
}

__device__ __forceinline__ double block_min(double v, double* scratch) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double s = lane < nw ? scratch[lane] : v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = fmin(s, __shfl_xor_sync(0xffffffffu, s, o));
    if (lane == 0) scratch[32] = s;
  }
  __syncthreads();
  return scratch[32];
}

// NaN-ignoring "better" test for arg-extremum with first-occurrence tie break.
// SIGN = +1: maximum, -1: minimum.  idx < 0 means "no candidate yet".
template <int SIGN>
__device__ __forceinline__ bool better(double v, long long i, double bv, long long bi) {
  if (i < 0 || v != v) return false;
  if (bi < 0 || bv != bv) return true;  // a NaN incumbent is no candidate
  if (SIGN > 0 ? (v > bv) : (v < bv)) return true;
  return v == bv && i < bi;
}

// Block-wide arg-extremum; result in thread 0.  sv/si: shared scratch of 32 each.
template <int SIGN>
__device__ __forceinline__ void block_argext(double& v, long long& i, double* sv, long long* si) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, v, o);
    long long oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (better<SIGN>(ov, oi, v, i)) { v = ov; i = oi; }
  }
  __syncthreads();
  if (lane == 0) { sv[wid] = v; si[wid] = i; }
  __syncthreads();
  if (wid == 0) {
    double bv = lane < nw ? sv[lane] : 0.0;
    long long bi = lane < nw ? si[lane] : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better<SIGN>(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    v = bv; i = bi;
  }
}

// Final reduction of per-block (value, index) candidates: one block per unit
// (GLS curve / PDM call).  SIGN = +1 arg-max, -1 arg-min; NaN ignored, first
// occurrence wins (np.nanargmax / np.nanargmin, reference core.py:202-210).
template <int SIGN>
__global__ void __launch_bounds__(256)
argext_final_kernel(const double* __restrict__ red_val, const long long* __restrict__ red_idx,
                    int nblk, long long* __restrict__ arg_out, double* __restrict__ val_out) {
  __shared__ double sv[32];
  __shared__ long long si[32];
  const int unit = blockIdx.x;
  double bv = 0.0;
  long long bi = -1;
  for (int k = threadIdx.x; k < nblk; k += blockDim.x) {
    double v = red_val[(long long)unit * nblk + k];
    long long i = red_idx[(long long)unit * nblk + k];
    if (better<SIGN>(v, i, bv, bi)) { bv = v; bi = i; }
  }
  block_argext<SIGN>(bv, bi, sv, si);
  if (threadIdx.x == 0) {
    if (arg_out) arg_out[unit] = bi;
    if (val_out) val_out[unit] = bi >= 0 ? bv : nan("");
  }
}

#endif  // __CUDACC__

}  // namespace pdc
