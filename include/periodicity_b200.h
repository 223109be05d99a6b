/*
 * periodicity_b200 -- C ABI of the B200-native trial-frequency search.
 *
 * The reference (dioph/periodicity, pure Python) has no FFI of its own; the
 * seams these entry points replace are Python call sites:
 *
 *   pdc_gls*   replaces the three `_trig_sum` calls plus the numpy epilogue of
 *              `GLS.__call__`          (src/periodicity/spectral.py:109-132)
 *   pdc_pdm*   replaces `pool.map(self._pdm, self.periods)` in `PDM.__call__`
 *                                      (src/periodicity/phase.py:185-187,128-149)
 *   pdc_stringlength*  replaces `pool.map(self._stringlength, periods)` in
 *              `StringLength.__call__` (src/periodicity/phase.py:68-70,45-51)
 *
 * Everything before those lines (signal coercion, grid derivation, weight
 * normalisation) and after them (FSeries wrap, sub-harmonic averaging) stays in
 * the Python drop-in classes `periodicity_b200.spectral.GLS` and
 * `periodicity_b200.phase.PDM`, which call this library through ctypes
 * (see INTEGRATION.md for the binding).
 *
 * Conventions
 *   - All functions return 0 (PDC_OK) or a pdc_status error; nothing throws or
 *     aborts across the ABI.  The message for the last error on the calling
 *     thread is returned by pdc_last_error().
 *   - The caller owns every buffer passed in; the library never retains a
 *     caller pointer past return.  A pdc_ctx owns one CUDA device, one stream
 *     and grow-only device scratch (pdc_ctx_create), or several devices with one
 *     worker thread each (pdc_ctx_create_multi).  A ctx is not thread-safe; distinct
 *     ctxs may be used concurrently.
 *   - Host entry points (`pdc_gls`, `pdc_gls_batch`, `pdc_pdm`) take host
 *     pointers and are synchronous: inputs are copied to the device, results
 *     are back in the output buffers on return.
 *   - Device entry points (`*_dev`) take device pointers on the ctx's device,
 *     are ordered on `stream` (a cudaStream_t passed as void*; NULL is CUDA's
 *     legacy default stream, as everywhere in CUDA -- it is what
 *     torch.cuda.current_stream().cuda_stream yields by default; pass
 *     PDC_STREAM_CTX to use the ctx's own stream) and return without synchronising.
 *     Calls on one ctx share its scratch buffers: each call is ordered (by a CUDA
 *     event) after the previous call on the same ctx, whatever streams they use.
 *   - All arrays are C-contiguous float64 (what TSeries.time / .values yield,
 *     core.py:60-66,479-481); FP32 is an internal detail of the kernels.
 *   - There is no CPU fallback: without a usable CUDA device the ctx cannot be
 *     created (PDC_ENODEVICE).
 */
#ifndef PERIODICITY_B200_H_
#define PERIODICITY_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDC_VERSION 200 /* 0.2.0 */

#if defined(__GNUC__)
#define PDC_API __attribute__((visibility("default")))
#else
#define PDC_API
#endif

typedef enum pdc_status {
  PDC_OK = 0,
  PDC_EINVAL = 1,    /* bad argument (Python raises ValueError, cf. core.py:469-470) */
  PDC_ECUDA = 2,     /* CUDA runtime error */
  PDC_ENOMEM = 3,    /* device or host allocation failed */
  PDC_ENODEVICE = 4  /* no usable CUDA device */
} pdc_status;

/* pdc_gls flags */
#define PDC_GLS_FIT_MEAN 1u /* fit_mean=True   (spectral.py:104-108,112-113,125-127) */
#define PDC_GLS_PSD 2u      /* psd=True        (spectral.py:129-130): power *= psd_scale */

typedef struct pdc_ctx pdc_ctx;

/* `stream` value selecting the ctx's own non-blocking stream in the *_dev calls. */
#define PDC_STREAM_CTX ((void*)(intptr_t)-1)

PDC_API int pdc_version(void);
PDC_API const char* pdc_last_error(void);

/* Create / destroy a context bound to CUDA device `device` (ordinal as seen by
 * the process).  Creation initialises the device, a non-blocking stream and
 * reads the SM count used to size grids. */
PDC_API int pdc_ctx_create(pdc_ctx** out, int device);
PDC_API int pdc_ctx_destroy(pdc_ctx* ctx);

/*
 * Multi-device context (the single-process counterpart of the reference's transparent
 * multiprocessing.Pool fan-out, phase.py:182-187).  `device_ids[0..ndev)` are CUDA ordinals of this
 * process, ndev <= PDC_MAX_PEERS (an ordinal may repeat: each entry gets its own stream, scratch and worker).  HOST-pointer entry points called on such a ctx shard the work over the devices --
 * pdc_gls: contiguous slices of the frequency grid; pdc_pdm / pdc_aov / pdc_ce / pdc_stringlength: slices of the
 * period grid; pdc_gls_batch: contiguous groups of curves balanced by sample count; pdc_gls_multi: groups of series
 * -- with one host worker thread per device: inputs are uploaded to every device concurrently, each device copies
 * its slice of the result straight into the caller's host buffer, and the (best value, best index) candidates are
 * reduced on the host (NaN ignored, first occurrence).  Values are identical to the single-device call's for the same
 * slice.  Work too small to be worth cutting (below ~5e8 evaluations per device; env PDC_MULTI_MIN_EVALS) uses fewer
 * devices.  Device-pointer (`*_dev`) entry points act on device_ids[0].  ndev == 1 gives an ordinary ctx.
 * pdc_ctx_destroy releases everything.
 */
PDC_API int pdc_ctx_create_multi(pdc_ctx** out, const int* device_ids, int ndev);
/* Number of devices of the ctx (1 for pdc_ctx_create) and the CUDA ordinal of its index-th device (-1 if out of range). */
PDC_API int pdc_ctx_device_count(pdc_ctx* ctx);
PDC_API int pdc_ctx_device_id(pdc_ctx* ctx, int index);
/* Block until all work queued on the ctx's stream has finished. */
PDC_API int pdc_ctx_synchronize(pdc_ctx* ctx);
/* SM count of the ctx's device (148 on B200); negative on error. */
PDC_API int pdc_ctx_sm_count(pdc_ctx* ctx);

/*
 * Generalised Lomb-Scargle power on the uniform grid f_j = fmin + (j0 + j)*df,
 * j = 0..nf-1, with exact trigonometric sums (the quantities `_trig_sum`
 * approximates, spectral.py:11-16) and the tau-offset algebra of
 * spectral.py:113-128 evaluated in float64 per frequency.
 *
 *   t, y        float64[n]  sample times and values (times need not be sorted)
 *   w           float64[n]  sample weights err**-2 (any positive scale; they are
 *                           normalised to sum 1 as in spectral.py:102-103), or
 *                           NULL for uniform weights (err=None, spectral.py:99-100)
 *   fmin, df    grid origin and spacing (spectral.py:88-97; Python owns the grid)
 *   j0          index of the first frequency this call evaluates (0 for a whole
 *               grid; the shard offset when the grid is split across GPUs)
 *   nf          number of frequencies evaluated by this call
 *   flags       PDC_GLS_FIT_MEAN | PDC_GLS_PSD
 *   psd_scale   0.5 * sum(err**-2) (spectral.py:130); ignored without PDC_GLS_PSD,
 *               in which case power is divided by YY = sum w y^2 (spectral.py:132)
 *   power_out   float64[nf]  periodogram values
 *   argmax_out  index (0-based within this call's nf) of the largest non-NaN
 *               power, first occurrence (np.nanargmax, core.py:202-205); -1 if
 *               every value is NaN.  May be NULL.
 *   max_out     that power (np.nanmax, core.py:212-215).  May be NULL.
 */
PDC_API int pdc_gls(pdc_ctx* ctx, const double* t, const double* y, const double* w, int64_t n,
            double fmin, double df, int64_t j0, int64_t nf, unsigned flags, double psd_scale,
            double* power_out, int64_t* argmax_out, double* max_out);

/* Same, device pointers, stream-ordered, no synchronisation. `argmax_out` /
 * `max_out` are device pointers too (or NULL). */
PDC_API int pdc_gls_dev(pdc_ctx* ctx, const double* t, const double* y, const double* w, int64_t n,
                double fmin, double df, int64_t j0, int64_t nf, unsigned flags, double psd_scale,
                double* power_out, int64_t* argmax_out, double* max_out, void* stream);

/*
 * GLS power at an ARBITRARY list of frequencies (non-uniform or user-supplied grids; the reference lists flexible
 * grids among its TODOs, phase.py:11-15, and its GLS cannot offer them because `_trig_sum` is an FFT on a uniform
 * grid, spectral.py:11-40).  Same formula and outputs as pdc_gls (spectral.py:113-132 with exact sums); every
 * (sample, frequency) pair is seeded exactly, there is no recurrence along the frequency axis.
 *
 *   freqs       float64[nfreq]  frequencies in any order (host pointer in pdc_gls_freqs, device pointer in the _dev twin)
 *   power_out   float64[nfreq]  in the order of `freqs`
 *   argmax_out / max_out        index into `freqs` of the largest non-NaN power (first occurrence) and that power
 */
PDC_API int pdc_gls_freqs(pdc_ctx* ctx, const double* t, const double* y, const double* w, int64_t n,
                          const double* freqs, int64_t nfreq, unsigned flags, double psd_scale,
                          double* power_out, int64_t* argmax_out, double* max_out);
PDC_API int pdc_gls_freqs_dev(pdc_ctx* ctx, const double* t, const double* y, const double* w, int64_t n,
                              const double* freqs, int64_t nfreq, unsigned flags, double psd_scale,
                              double* power_out, int64_t* argmax_out, double* max_out, void* stream);

/*
 * Frequency-grid-sharded GLS with the all-gather FUSED into the epilogue: rank `rank` of `world`
 * evaluates frequencies [j0, j0 + nf) like pdc_gls_dev, and its epilogue kernel stores every power
 * value straight into ALL ranks' result buffers through NVLink peer mappings (one coalesced 8-byte
 * store per destination), and its arg-max kernel stores (max, global argmax) into slot `rank` of
 * every rank's candidate table.  No NCCL call, no staging buffer: the "collective" is part of
 * the kernel that produces the data.  The caller provides buffers mapped into this process for
 * every rank (e.g. torch.distributed._symmetric_memory: `handle.buffer_ptrs`) and must end every call
 * with a cross-rank barrier (`handle.barrier()`) before reading the result; with two alternating sets of
 * buffers no barrier is needed before the call (periodicity_b200/dist.py::_symm_buffer explains why).
 *
 *   power[r]  base of rank r's full-grid power array (float64[nf_total]); element j0 + j is written
 *   best[r]   rank r's candidate table (float64[2*world]); elements 2*rank, 2*rank+1 are written
 */
#define PDC_MAX_PEERS 16
typedef struct pdc_fanout {
  int32_t world;
  int32_t rank;
  double* power[PDC_MAX_PEERS];
  double* best[PDC_MAX_PEERS];
} pdc_fanout;

PDC_API int pdc_gls_dev_fanout(pdc_ctx* ctx, const double* t, const double* y, const double* w, int64_t n,
                               double fmin, double df, int64_t j0, int64_t nf, unsigned flags,
                               double psd_scale, const pdc_fanout* dst, void* stream);

/*
 * Batched GLS: B independent light curves stored back to back
 * (curve b = samples offsets[b] .. offsets[b+1]-1), each with its own grid
 * origin fmin[b] and spacing df[b] but a common number of frequencies nf
 * (the survey workload, and GLS.bootstrap, spectral.py:140-152).
 *
 *   offsets     int64[B+1]   (host pointer in both variants)
 *   fmin, df    float64[B]   (host pointers in both variants)
 *   psd_scale   float64[B] or NULL (host pointer; required with PDC_GLS_PSD)
 *   power_out   float64[B*nf] row-major, or NULL to keep only the peaks
 *   argmax_out  int64[B], max_out float64[B]  (either may be NULL)
 */
PDC_API int pdc_gls_batch(pdc_ctx* ctx, const double* t, const double* y, const double* w,
                  const int64_t* offsets, int64_t B, const double* fmin, const double* df,
                  int64_t nf, unsigned flags, const double* psd_scale,
                  double* power_out, int64_t* argmax_out, double* max_out);

PDC_API int pdc_gls_batch_dev(pdc_ctx* ctx, const double* t, const double* y, const double* w,
                      const int64_t* offsets, int64_t B, const double* fmin, const double* df,
                      int64_t nf, unsigned flags, const double* psd_scale,
                      double* power_out, int64_t* argmax_out, double* max_out, void* stream);

/*
 * GLS of S series sampled at the SAME times (shared timestamps): `GLS.bootstrap` with
 * err=None (spectral.py:140-152: the values are resampled, the times are kept and all
 * weights are equal), or a survey sector whose light curves share one time axis.  The
 * rotation and the window sums are computed once per group of 8 series, which makes a
 * series-evaluation ~3x cheaper than in pdc_gls_batch; results are the same quantities.
 *
 *   t          float64[n]     common sample times
 *   Y          float64[S*n]   row-major, series s = Y[s*n .. s*n + n - 1]
 *   w          float64[n]     weights shared by all series, or NULL (uniform)
 *   fmin, df, j0, nf, flags, psd_scale   as pdc_gls (one grid for all series)
 *   power_out  float64[S*nf] or NULL;  argmax_out int64[S], max_out float64[S] (may be NULL)
 */
PDC_API int pdc_gls_multi(pdc_ctx* ctx, const double* t, const double* Y, const double* w, int64_t n,
                          int64_t S, double fmin, double df, int64_t j0, int64_t nf, unsigned flags,
                          double psd_scale, double* power_out, int64_t* argmax_out, double* max_out);
PDC_API int pdc_gls_multi_dev(pdc_ctx* ctx, const double* t, const double* Y, const double* w, int64_t n,
                              int64_t S, double fmin, double df, int64_t j0, int64_t nf, unsigned flags,
                              double psd_scale, double* power_out, int64_t* argmax_out, double* max_out,
                              void* stream);

/*
 * Phase Dispersion Minimisation: theta statistic of `PDM._pdm`
 * (phase.py:128-149) for each trial period.
 *
 *   t, x        float64[n]   sample times and values
 *   periods     float64[np]  trial periods (phase.py:180; any order, all != 0)
 *   nb, nc      bins and covers (phase.py:108-112); m0 = nb*nc fine bins
 *   theta_out   float64[np]  theta in the order of `periods`
 *   argmin_out  index of the smallest non-NaN theta, first occurrence; -1 if all
 *               NaN.  May be NULL.
 *   min_out     that theta.  May be NULL.
 */
PDC_API int pdc_pdm(pdc_ctx* ctx, const double* t, const double* x, int64_t n,
            const double* periods, int64_t np, int nb, int nc,
            double* theta_out, int64_t* argmin_out, double* min_out);

PDC_API int pdc_pdm_dev(pdc_ctx* ctx, const double* t, const double* x, int64_t n,
                const double* periods, int64_t np, int nb, int nc,
                double* theta_out, int64_t* argmin_out, double* min_out, void* stream);

/*
 * Analysis of Variance periodogram (Schwarzenberg-Czerny 1989).  The reference only lists it as a
 * TODO (phase.py:11); it is computed from the same per-period phase-bin histograms as PDM
 * (phase.py:131,137-141 with nc = 1: phi = (t / P) % 1, bin k = [k/nb, (k+1)/nb)):
 *     Theta(P) = [(N - r) / (r - 1)] * sum_b n_b (mean_b - mean)^2 / sum_b sum_{i in b} (x_i - mean_b)^2
 * over the r populated bins -- the F statistic of a one-way ANOVA of the values grouped by phase bin.
 *
 *   theta_out   float64[np]  statistic in the order of `periods` (NaN for period 0 / inf / NaN or r < 2)
 *   argmax_out  index of the LARGEST non-NaN value, first occurrence; -1 if all NaN.  May be NULL.
 *   max_out     that value.  May be NULL.
 */
PDC_API int pdc_aov(pdc_ctx* ctx, const double* t, const double* x, int64_t n,
            const double* periods, int64_t np, int nb,
            double* theta_out, int64_t* argmax_out, double* max_out);

PDC_API int pdc_aov_dev(pdc_ctx* ctx, const double* t, const double* x, int64_t n,
                const double* periods, int64_t np, int nb,
                double* theta_out, int64_t* argmax_out, double* max_out, void* stream);

/*
 * Conditional-entropy periodogram (Graham et al. 2013).  The reference only lists it as a TODO (phase.py:13); it is
 * computed from per-period COUNT histograms over nphi phase bins x nm magnitude bins, with the reference's phase and
 * bin-edge conventions (phase.py:131,138-140 with nc = 1: phi = (t / P) % 1, phase bin k = [k/nphi, (k+1)/nphi)) and
 * magnitude bin min(int(nm (x - min x) / (max x - min x)), nm - 1):
 *     H_c(P) = sum_jk p(phi_j, m_k) ln( p(phi_j) / p(phi_j, m_k) ),   p = cell occupation / N, over occupied cells.
 *
 *   h_out       float64[np]  conditional entropy in the order of `periods` (NaN for period 0 / inf / NaN)
 *   argmin_out  index of the SMALLEST non-NaN value, first occurrence; -1 if all NaN.  May be NULL.
 *   min_out     that value.  May be NULL.
 */
PDC_API int pdc_ce(pdc_ctx* ctx, const double* t, const double* x, int64_t n,
           const double* periods, int64_t np, int nphi, int nm,
           double* h_out, int64_t* argmin_out, double* min_out);

PDC_API int pdc_ce_dev(pdc_ctx* ctx, const double* t, const double* x, int64_t n,
               const double* periods, int64_t np, int nphi, int nm,
               double* h_out, int64_t* argmin_out, double* min_out, void* stream);

/*
 * Gregory-Loredo (1992) periodogram for EVENT ARRIVAL TIMES (no values): a TODO of the reference (phase.py:14).  For
 * every trial period the odds ratio of a stepwise periodic model with m phase bins against the constant model,
 * marginalised over the bin rates and over the phase offset (nc offsets per bin: events are counted in m * nc fine
 * phase bins with the reference's phase and edge conventions, phase.py:131,138-140, and the m bins at offset c are
 * circular unions of nc consecutive fine bins),
 *     O_m(P) = [(m - 1)! / (N + m - 1)!] m^N < prod_j n_j! >_offsets ,
 * averaged over m = 2 .. m_max with equal weights.  One count-histogram pass per m.
 *
 *   t           float64[n]   event arrival times
 *   lnodds_out  float64[np]  ln O(P) in the order of `periods` (NaN for period 0 / inf / NaN)
 *   argmax_out / max_out     index and value of the LARGEST non-NaN ln O (first occurrence).  May be NULL.
 *   m_max >= 2, nc >= 1, m_max * nc <= 4096 (and the fine bins must fit a shared-memory histogram).
 */
PDC_API int pdc_gl(pdc_ctx* ctx, const double* t, int64_t n, const double* periods, int64_t np, int m_max, int nc,
                   double* lnodds_out, int64_t* argmax_out, double* max_out);
PDC_API int pdc_gl_dev(pdc_ctx* ctx, const double* t, int64_t n, const double* periods, int64_t np, int m_max, int nc,
                       double* lnodds_out, int64_t* argmax_out, double* max_out, void* stream);

/*
 * String Length (Dworetsky 1983): `StringLength._stringlength` (phase.py:45-51) for each trial
 * period, replacing `pool.map(self._stringlength, periods)` (phase.py:68-70).
 *
 *   t         float64[n]   sample times
 *   m         float64[n]   the scaled signal `self.m` of phase.py:64-65 (computed by the caller)
 *   periods   float64[np]  trial periods (phase.py:67; any order, all != 0)
 *   ell_out   float64[np]  string length in the order of `periods`:
 *               phi = (t / P) % 1, samples stably sorted by phi (core.py:543-544,473-477),
 *               sum_j hypot(m[j+1] - m[j], phi[j+1] - phi[j]) with indices mod n (np.roll)
 *   argmin_out / min_out   shortest string (NaN ignored, first occurrence).  May be NULL.
 */
PDC_API int pdc_stringlength(pdc_ctx* ctx, const double* t, const double* m, int64_t n,
                             const double* periods, int64_t np,
                             double* ell_out, int64_t* argmin_out, double* min_out);

PDC_API int pdc_stringlength_dev(pdc_ctx* ctx, const double* t, const double* m, int64_t n,
                                 const double* periods, int64_t np,
                                 double* ell_out, int64_t* argmin_out, double* min_out, void* stream);

/*
 * The k highest local maxima of each row of a row-major float64 [rows, n] array
 * (periodograms): what `FSeries.find_peaks()` + sorting by height gives
 * (core.py:283-317,944-955 -> scipy.signal.find_peaks(values, prominence=0.0)):
 * strict local maxima, the midpoint of flat tops, never the first or last sample,
 * NaN never a peak; ties go to the lower index.
 *
 *   idx_out  int64[rows*k]    peak positions, highest first; -1 where a row has fewer than k peaks
 *   val_out  float64[rows*k]  their values (NaN where idx is -1)
 * k <= 64.  `pdc_peaks_topk` takes host pointers (the values are copied to the device),
 * `pdc_peaks_topk_dev` device pointers, e.g. the power_out of pdc_gls_dev, so a 1e7-point
 * periodogram never has to leave the GPU to find its peaks.
 */
PDC_API int pdc_peaks_topk(pdc_ctx* ctx, const double* values, int64_t rows, int64_t n, int k,
                           int64_t* idx_out, double* val_out);
PDC_API int pdc_peaks_topk_dev(pdc_ctx* ctx, const double* values, int64_t rows, int64_t n, int k,
                               int64_t* idx_out, double* val_out, void* stream);

/*
 * Half-maximum crossings of given peaks, the index arithmetic of `FSeries.periods_at_half_max`
 * (core.py:957-972): for peak p of row r, level = v[p] - height / 2 and d = v - level, where
 * height = v[p], or height[r, j] if `height` (float64[rows, k], e.g. prominences, core.py:960-963) is not NULL;
 *   left_out   last  j <= p - 2 with signbit(d[j]) != signbit(d[j + 1])   (core.py:967; -1 if none)
 *   right_out  first j >= p     with signbit(d[j]) != signbit(d[j + 1])   (core.py:968; -1 if none)
 * `peak_idx` is int64[rows, k] (e.g. the idx_out of pdc_peaks_topk; entries < 0 give -1, -1).
 * The `_dev` twin takes device pointers for all four arrays.
 */
PDC_API int pdc_peaks_halfmax(pdc_ctx* ctx, const double* values, int64_t rows, int64_t n, int k,
                              const int64_t* peak_idx, const double* height, int64_t* left_out, int64_t* right_out);
PDC_API int pdc_peaks_halfmax_dev(pdc_ctx* ctx, const double* values, int64_t rows, int64_t n, int k,
                                  const int64_t* peak_idx, const double* height, int64_t* left_out,
                                  int64_t* right_out, void* stream);

/* Period-grid-sharded PDM with the all-gather fused into the epilogue (see pdc_gls_dev_fanout):
 * this rank evaluates `periods[0..np)`, which are elements [offset, offset + np) of the full period
 * grid; theta goes to power[r][offset + i] and (min, global argmin) to best[r][2*rank..] of every rank. */
PDC_API int pdc_pdm_dev_fanout(pdc_ctx* ctx, const double* t, const double* x, int64_t n,
                               const double* periods, int64_t np, int nb, int nc, int64_t offset,
                               const pdc_fanout* dst, void* stream);

/* Number of kernels this ctx has launched since creation (bench.py's
 * `gpu_launches` claim is the difference across the timed region). */
PDC_API int64_t pdc_ctx_launch_count(pdc_ctx* ctx);

/* Duration in milliseconds of the dominant kernel (GLS strip kernel or PDM
 * histogram kernel) of the most recent call on this ctx, from CUDA events
 * recorded on the launching stream.  Synchronises on the end event.
 * Negative on error / if no call has been made. */
PDC_API double pdc_ctx_last_main_kernel_ms(pdc_ctx* ctx);

/* Cumulative duration in milliseconds of every dominant-kernel launch made
 * through this ctx since creation; `count_out` (may be NULL) receives the number
 * of launches.  Synchronises on the last recorded end event.  bench.py takes the
 * difference across its timed region: average launch duration = d(ms) / d(count). */
PDC_API double pdc_ctx_main_kernel_ms_total(pdc_ctx* ctx, int64_t* count_out);

/* Which hot kernel the most recent GLS call on this ctx (primary device) used: 0 = gls_strip_kernel (FP32 SIMT; also the
 * free-frequency kernel), 1 = gls_umma_kernel (tcgen05 tensor cores), 2 = gls_umma_kernel with the fine operand of the
 * curve precomputed once (one long curve), 3 = gls_umma2_kernel (one long curve, a pair of CTAs per tile, cta_group::2).  The choice is automatic (problem size, weights, grid direction); the
 * environment variable PDC_GLS_UMMA=0|1 read at ctx creation forces it off / on whenever eligible.  < 0 on error. */
PDC_API int pdc_ctx_last_gls_path(pdc_ctx* ctx);

/* Work decomposition the tensor-core GLS kernels would use for a call of B curves of at most nmax samples on nf
 * frequencies on a device with sm_count SMs (pure host arithmetic, needs no GPU; knobs fine / cg2 / nsplit / chunk as the
 * environment variables PDC_GLS_UMMA_FINE / _CG2 / _NSPLIT / _CHUNK, -1 or 0 = automatic).  out[0..10] = {path (1 one CTA per
 * tile, 2 the same with the fine operand precomputed, 3 a pair of CTAs per tile), fine indices per tile, coarse blocks per
 * curve, type-1 tiles, coarse blocks per type-1 tile, type-2 tiles, coarse blocks per type-2 tile, sample splits, stages of
 * 16 samples per accumulation run, jobs, bytes of fine-operand scratch}. */
PDC_API int pdc_debug_umma_plan(int sm_count, int64_t B, int64_t nf, int64_t nmax, int fine, int cg2, int nsplit, int chunk,
                                int64_t* out);

/* Diagnostics of the tensor-core GLS kernel (gls_umma.cu), filled only when the ctx was created with the environment
 * variable PDC_GLS_UMMA_PROF=1: per work item (thread block) of the most recent launch four SM clock stamps
 * {start, main loop begin, main loop end, end of flush} are copied to `out` (up to `cap` items x 4 values).
 * Returns the number of work items of that launch (0 if profiling is off or the kernel has not run), < 0 on error.
 * Synchronises the ctx. */
PDC_API int64_t pdc_debug_umma_prof(pdc_ctx* ctx, int64_t* out, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* PERIODICITY_B200_H_ */
