// Phase Dispersion Minimisation: theta statistic for a grid of trial periods.
//
// Replaces `pool.map(self._pdm, self.periods)` of `PDM.__call__`
// (reference src/periodicity/phase.py:185-187) and `PDM._pdm` (phase.py:128-149).
//
// Restatement implemented here (SURVEY.md §8a "PDM restatement"): with
// m0 = nb*nc (phase.py:130) every sample has phase phi = (t/P) % 1 (phase.py:131)
// and lies in exactly one FINE bin q, thr[q] <= phi < thr[q+1], thr[k] = k/m0
// being the very float64 thresholds the reference compares against
// (phase.py:138-140).  The reference's coarse bin k (phase.py:137-141) is the
// circular union of fine bins k .. k+nc-1.  With x' = (x - mean(x)) / std(x, ddof=1)
// (so that sum_i x'_i^2 = N - 1 and the division by sigma^2, phase.py:148,165, is folded in)
// and every sample lying in exactly nc coarse bins,
//     sum_{k good} (n_k - 1) s_k^2 = sum_k (Q_k - S_k^2 / n_k) = nc (N - 1) - sum_{k: n_k >= 1} S_k^2 / n_k,
// because a bin with n_k = 1 has Q_k = S_k^2 (it contributes 0 whether it is dropped, phase.py:142,
// or not) and an empty bin contributes nothing.  Hence only the count n and the sum S = sum x'
// per fine bin are histogrammed -- no per-bin sum of squares -- and
//     theta = [nc (N - 1) - sum_{n_k >= 1} S_k^2 / n_k] / sum_{n_k > 1} (n_k - 1)      (phase.py:145-149).
// The argsort of phase.py:132-134 does not influence the result and is dropped.
//
// Mapping (north_star: "per-trial-period phase-bin variance histograms in shared
// memory with warp-aggregated atomics"): every trial period has a PRIVATE histogram column in shared
// memory, hist[bin][column] -- bank = column % 32 whatever the bin, so updates are conflict free and nothing
// needs aggregating.  A thread owns PPT = 2 columns (long curves; 1 for short ones): each sample read from the
// tile feeds both periods, and a trip of 16 samples keeps 32 independent phase -> bin -> update chains in flight.
// Two levels:
//   level 1  one 32-bit word per bin, (count << 23) + sum rint(x' 2^q): one sample is ONE native
//            shared-memory integer atomic without return value (ATOMS.ADD, fire-and-forget; the
//            column is private, the atomic is used for its single-instruction read-modify-write,
//            measured 13.7 updates/clk/SM against 9.5 for LDS + IADD + STS, profiles/r01/pipes_r01.json);
//   level 2  integer planes count[bin][column], sum[bin][column] (fixed point, exact) over the memory of the
//            float2 columns, fed from level 1 every 256 samples with three atomics per bin (fetch-and-clear of
//            the packed word, two adds), merged into the FP64 partials in global memory every 8192 samples.
// Level 1 needs 256 max|x'| 2^q < 2^22 with a fine enough step 2^-q (q >= 11) and is skipped for
// short curves, non-finite samples or huge |t / P|: then samples go straight to float2 (count, sum x') columns
// (one LDS.64 + 2 FADD + STS.64).  A shared-memory FLOAT atomicAdd is a CAS loop on sm_100a
// (ATOMS.CAST.SPIN, 0.96 updates/clk/SM for warp-shared histograms) and so is a 64-bit integer
// one, which is why the packed word is 32 bits wide.
//
// Bin fidelity: t/P is the correctly rounded quotient (one Newton correction of
// t * (1/P) with the exact FMA residual), phi = q - floor(q) is exact, and the
// bin index from phi*m0 is re-checked against thr[] whenever phi*m0 is within
// 2^-30 of an integer, so ties on bin edges (integer times, rational periods)
// fall exactly where the reference puts them.
//
// Kernels (two launches per call since round 2, seven in round 1):
//   pdm_stats_kernel  mean, 1/std, packing exponent in one pass; multi-block partials, the last block finalises
//   pdm_hist_kernel   the hot kernel; x' is formed while a tile is staged (no centring pass); every 8192 samples a
//                     thread adds its columns to ONE count plane (RED.ADD.32) and ONE 64-bit fixed-point sum plane
//                     (RED.ADD.64) shared by all sample splits -- integer addition is associative, so the totals do
//                     not depend on arrival order, and round 1's nine FP64 planes (288 MB on C3) shrink to 24 MB
//   pdm_epilogue_kernel  FP64 theta (phase.py:145-149) per trial period, clears the planes for the next call, stores
//                     theta (to every rank's buffer in the fan-out variant), arg-min; the last block finalises.
#include <cstring>
#include <type_traits>

#include "pdc_common.cuh"
#include "phase_common.cuh"

namespace pdc {

struct PdmMeta {
  double mean, inv_sd;
  double q_binned;  // sum of x'^2 over the samples with a finite time stamp (== N - 1 when all are)
  double t_absmax;  // max |t| over the finite time stamps: sizes the guard band of the fixed-point phase
  double t0, t_span;  // smallest finite stamp and the span of the finite stamps: the fixed-point phase works on t - t0
  int pack_q;       // >= 0: x' is also provided as 2^23 + rint(x' 2^pack_q) for the packed 32-bit histogram; -1: not usable
  int bad;          // some t or x is NaN / inf: the histogram kernel then runs its guarded variant
};

// Packed level-1 word = (count << PDM_SUM_BITS) + sum of fixed-point x' (two's complement in the low PDM_SUM_BITS bits).
// A window of PDM_PACK_FLUSH samples is fed into level 2 at once.  PDM_PACK_FLUSH = 256 (9 + 23 bits) holds ANY 256 samples
// by the choice of the exponent q (256 max|x'| 2^q < 2^22).  PDM_PACK_FLUSH = 512 (10 + 22 bits) halves the number of feeds
// (the feed is what separates the kernel from its ATOMS.ADD floor: 3 shared atomics per bin per window): a window is fed
// whole only if the block has verified sum |increment| < 2^21 over it while staging the tile (true for anything but a
// burst of outliers: Gaussian values reach 40 % of the bound); otherwise that window is fed in PDM_PACK_SUB = 128-sample
// pieces, which the same rule for q guarantees (128 max|x'| 2^q < 2^21).  Both are exact.
#ifndef PDM_PACK_FLUSH
#define PDM_PACK_FLUSH 512
#endif
constexpr int PDM_CNT_BITS = PDM_PACK_FLUSH > 256 ? 10 : 9;
constexpr int PDM_SUM_BITS = 32 - PDM_CNT_BITS;
constexpr int PDM_PACK_SUB = 128;     // guaranteed sub-window of the 512-sample layout
static_assert(PDM_PACK_FLUSH == 256 || PDM_PACK_FLUSH == 512, "PDM_PACK_FLUSH is 256 or 512");
constexpr int PDM_PACK_MIN_Q = 11;    // coarsest usable quantisation of x' (2^-11 sigma)
constexpr int PDM_PACK_MIN_N = 4096;  // shorter curves keep the FP32 columns (they are accurate to 1e-7 there)
constexpr int PDM_TILE = 1024;        // samples per shared-memory tile
#ifndef PDM_TRIP_CHAINS
#define PDM_TRIP_CHAINS 32            // independent (sample, period) chains per trip of the packed loop = PDM_TRIP_CHAINS / PPT samples
#endif
// Build-time switches of the packed loop; the defaults are the best of the sweep in profiles/r01/tune_pdm_r01.txt.
#ifndef PDM_EDGE_FIXUP
#define PDM_EDGE_FIXUP 1             // 1: updates go out with the fast bins at once, the rare edge sample is moved afterwards
#endif
#ifndef PDM_PREFETCH
#define PDM_PREFETCH 1               // 1: time stamps of the next trip are loaded before this trip's atomics
#endif
#ifndef PDM_FLUSH_BINS
#define PDM_FLUSH_BINS 10              // the level-1 -> level-2 feed handles this many bins of all the thread's columns at once: every
                                       // fetch-and-clear is in flight before the first dependent add (C3, B200: 2.950 ms with 1, 2.912 with 4,
                                       // 2.896 with 10, 2.903 with 20; profiles/tune_pdm_r02.txt)
#endif
#ifndef PDM_L2_INT
#define PDM_L2_INT 1                 // 1: second level of the packed path = integer planes fed with shared-memory atomics
#endif
#ifndef PDM_FEED_PLAIN
#define PDM_FEED_PLAIN 0             // 1: the feed uses plain LDS / STS on the (private) columns instead of atomics (needs PDM_L2_INT)
#endif
constexpr int PDM_FLUSH_TILES = 8;    // FP32 histograms are merged into FP64 every 8192 samples
constexpr int PDM_TILE_PAD = PDM_PREFETCH ? PDM_TRIP_CHAINS : 0;  // the prefetch of the packed loop reads one trip past the tile

// Statistics in ONE launch: a long curve is read by many SMs (one block would take ~100 us for 1e5 samples); every
// block writes its partial moments (about the first value x[0], one pass) in a fixed layout and the last block to
// finish reduces them with a fixed tree, so the result is deterministic.
struct PdmPart {
  double s1, s2;          // sum d, sum d^2 with d = x - x[0]
  double s1b, s2b, nb;    // the same over the samples with a finite time stamp (used only if some stamp is not finite)
  double xneg, xmax;      // -min x, max x
  double tabs;            // max |t| over the finite stamps
  double tneg, tmax;      // -min t, max t over the finite stamps
  int bad, pad_;
};
constexpr int PDM_STATS_THREADS = 256;
constexpr int PDM_STATS_MAXBLK = 128;

__device__ __forceinline__ double block_max(double v, double* scratch) { return -block_min(-v, scratch); }

__global__ void __launch_bounds__(PDM_STATS_THREADS)
pdm_stats_kernel(const double* __restrict__ t, const double* __restrict__ x, long long n, PdmPart* part,
                 unsigned* done, PdmMeta* meta) {
  __shared__ double scratch[32 * 10];
  __shared__ int s_last;
  const int G = gridDim.x;
  const double x0 = x[0];
  double s1 = 0.0, s2 = 0.0, s1b = 0.0, s2b = 0.0, nb = 0.0, xneg = -INFINITY, xmax = -INFINITY, tabs = 0.0;
  double tneg = -INFINITY, tmax = -INFINITY;
  int bad = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)G * blockDim.x) {
    const double xi = x[i], ti = t[i], d = xi - x0;
    s1 += d;
    s2 = fma(d, d, s2);
    xneg = fmax(xneg, -xi);
    xmax = fmax(xmax, xi);
    bad |= !isfinite(xi) || !isfinite(ti);
    if (isfinite(ti)) {
      tabs = fmax(tabs, fabs(ti));
      tneg = fmax(tneg, -ti);
      tmax = fmax(tmax, ti);
      s1b += d;
      s2b = fma(d, d, s2b);
      nb += 1.0;
    }
  }
  bad = __syncthreads_or(bad);
  {
    double sums[5] = {s1, s2, s1b, s2b, nb}, maxs[5] = {xneg, xmax, tabs, tneg, tmax};
    block_reduce_many<5, 5>(sums, maxs, scratch);
    s1 = sums[0]; s2 = sums[1]; s1b = sums[2]; s2b = sums[3]; nb = sums[4]; xneg = maxs[0]; xmax = maxs[1]; tabs = maxs[2];
    tneg = maxs[3]; tmax = maxs[4];
  }
  if (threadIdx.x == 0) {
    PdmPart& p = part[blockIdx.x];
    p.s1 = s1; p.s2 = s2; p.s1b = s1b; p.s2b = s2b; p.nb = nb; p.xneg = xneg; p.xmax = xmax; p.tabs = tabs; p.tneg = tneg; p.tmax = tmax; p.bad = bad;
    __threadfence();
    s_last = atomicAdd(done, 1u) == (unsigned)(G - 1);
  }
  __syncthreads();
  if (!s_last) return;
  // last block: thread k holds partial k (G <= PDM_STATS_MAXBLK <= blockDim), fixed-tree block reductions
  __threadfence();
  const bool has = (int)threadIdx.x < G;
  const PdmPart* p = part + threadIdx.x;
  {
    double sums[5] = {has ? __ldcg(&p->s1) : 0.0, has ? __ldcg(&p->s2) : 0.0, has ? __ldcg(&p->s1b) : 0.0,
                      has ? __ldcg(&p->s2b) : 0.0, has ? __ldcg(&p->nb) : 0.0};
    double maxs[5] = {has ? __ldcg(&p->xneg) : -INFINITY, has ? __ldcg(&p->xmax) : -INFINITY, has ? __ldcg(&p->tabs) : 0.0,
                      has ? __ldcg(&p->tneg) : -INFINITY, has ? __ldcg(&p->tmax) : -INFINITY};
    block_reduce_many<5, 5>(sums, maxs, scratch);
    s1 = sums[0]; s2 = sums[1]; s1b = sums[2]; s2b = sums[3]; nb = sums[4]; xneg = maxs[0]; xmax = maxs[1]; tabs = maxs[2];
    tneg = maxs[3]; tmax = maxs[4];
  }
  bad = __syncthreads_or(has ? __ldcg(&p->bad) : 0);
  if (threadIdx.x == 0) {
    const double dn = (double)n;
    const double dm = s1 / dn;                       // mean - x0
    const double mean = x0 + dm;
    double q = s2 - s1 * dm;                         // sum (x - mean)^2
    if (q < 0.0) q = 0.0;
    const double var = q / (dn - 1.0);               // phase.py:165  np.var(values, ddof=1)
    const double dmax = fmax(xmax - mean, mean + xneg);
    meta->mean = mean;
    meta->inv_sd = 1.0 / sqrt(var);
    // a sample whose phase is NaN fails every mask of phase.py:138-140 and is in no bin, but still
    // counts in sigma^2: the epilogue then needs sum x'^2 over the binned samples only
    double qb = s2b - 2.0 * dm * s1b + nb * dm * dm;   // sum (x - mean)^2 over the finite stamps
    if (qb < 0.0) qb = 0.0;
    meta->q_binned = bad ? qb / var : dn - 1.0;
    meta->bad = bad;
    meta->t_absmax = tabs;
    meta->t0 = tmax >= -tneg ? -tneg : 0.0;            // no finite stamp at all: any origin will do
    meta->t_span = tmax >= -tneg ? tmax + tneg : 0.0;
    // Packed first-level histogram (pdm_hist_kernel): one 32-bit word per (bin, period) holds the count in
    // its top PDM_CNT_BITS bits and sum rint(x' 2^q) in the rest (two's complement) for one feed window,
    // so 256 * max|x'| * 2^q must stay below 2^22 (9 + 23 bits; with 10 + 22 bits: 128 * max|x'| * 2^q < 2^21, the
    // same condition, for the guaranteed sub-window).  Worth it only if the quantisation step 2^-q is fine
    // enough (q >= PDM_PACK_MIN_Q: relative theta error ~ 0.4 * 2^-q / sqrt(nc N) / theta) and the curve is long.
    int pq = -1;
    const double xm = dmax / sqrt(var);
    if (!bad && xm > 0.0 && isfinite(xm) && n >= PDM_PACK_MIN_N) {
      pq = (int)floor(log2(16000.0 / xm));
      if (pq > 20) pq = 20;
      if (pq < PDM_PACK_MIN_Q) pq = -1;
    }
    meta->pack_q = pq;
    *done = 0u;   // self-resetting
  }
}

struct PdmArgs {
  const double* t;
  const double* x;      // raw values: x' = (x - mean) / std is formed while a tile is staged (no separate pass, no copy)
  const double* periods;
  const PdmMeta* meta;
  unsigned* cnt_plane;           // [m0][np]  samples per fine bin, all sample splits add into it (RED.ADD.32)
  unsigned long long* sum_plane; // [m0][np]  sum x' per fine bin as 64-bit fixed point, 32 fraction bits (RED.ADD.64)
  long long n, np;
  int m0, nsplit;
  int allow_packed;     // 0: keep the float2 columns whatever the curve (AoV: its ratio of sums needs their accuracy)
};

// Shared-memory layout: hist[bin][VT] float2 = (count, sum x') and hist32[bin][VT] packed words, VT = THREADS * PPT
// period columns per block: a column is private to one thread and conflict free (bank = column % 32).
// PPT = trial periods per thread.  With PPT = 2 a thread owns columns tid and tid + THREADS: every sample read
// from the tile (warp-uniform LDS.128) feeds two independent phase -> bin -> ATOMS chains, which halves the
// shared-memory load traffic and the loop overhead per histogram update (the update itself stays one ATOMS.ADD).
#ifndef PDM_MINB
#define PDM_MINB 1   // tuning aid: minimum resident blocks per SM promised to ptxas for the two-periods-per-thread variants
#endif
template <int THREADS, int PPT>
__global__ void __launch_bounds__(THREADS, PPT == 2 ? PDM_MINB : 1)
pdm_hist_kernel(const PdmArgs a) {
  constexpr int VT = THREADS * PPT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int m0 = a.m0;
  double* s_t = reinterpret_cast<double*>(smem_raw);                 // [PDM_TILE + PDM_TILE_PAD]
  double* s_thr = s_t + PDM_TILE + PDM_TILE_PAD;                     // [m0 + 1]  (padded to even)
  float2* hist = reinterpret_cast<float2*>(s_thr + ((m0 + 2) & ~1)); // [m0][VT]
  float* s_x = reinterpret_cast<float*>(hist + (size_t)m0 * VT);     // [PDM_TILE]  (x' as float, or packed increments)
  unsigned* hist32 = reinterpret_cast<unsigned*>(s_x + PDM_TILE);    // [m0][VT] packed first-level columns
  __shared__ unsigned s_winabs[8][PDM_TILE / PDM_PACK_FLUSH];        // per warp: sum |increment| of each feed window of the tile

  const int split = blockIdx.x % a.nsplit;
  const long long pb = blockIdx.x / a.nsplit;
  long long pis[PPT];
  bool valids[PPT];
  double Ps[PPT], rPs[PPT];
  bool in_range = true, plain_ok = true;
#pragma unroll
  for (int s = 0; s < PPT; ++s) {
    pis[s] = pb * VT + s * THREADS + threadIdx.x;
    valids[s] = pis[s] < a.np;
    double P = valids[s] ? a.periods[pis[s]] : 1.0;
    if (!isfinite(1.0 / P) || !isfinite(P)) P = 1.0;  // invalid trial period: theta is set to NaN by the epilogue
    Ps[s] = P;
    rPs[s] = 1.0 / P;
    // the fixed-point phase works on t - t0: |(t - t0) / P| < 2^18 and |t / P| < 2^30 (guard band, see phase_common.cuh);
    // padding columns (P = 1) must not veto the block's fast path
    in_range = in_range && (!valids[s] || (fabs(rPs[s]) * a.meta->t_span < PDM_FAST_LIMIT &&
                                           fabs(rPs[s]) * a.meta->t_absmax < PDM_SHIFT_LIMIT));
    plain_ok = plain_ok && (!valids[s] || fabs(rPs[s]) * a.meta->t_absmax < PDM_FAST_LIMIT);
  }
  const bool clamp_bins = a.meta->bad != 0;        // block-uniform
  const double m0d = (double)m0;
  const unsigned kmax = (unsigned)(m0 - 1);

  for (int k = threadIdx.x; k <= m0; k += THREADS) s_thr[k] = (double)k / m0d;  // phase.py:138-140
  for (int k = threadIdx.x; k < PDM_TILE_PAD; k += THREADS) s_t[PDM_TILE + k] = 0.0;
  for (int k = threadIdx.x; k < m0 * VT; k += THREADS) {
    hist[k] = make_float2(0.f, 0.f);
    hist32[k] = 0u;
  }

  const long long per = (a.n + a.nsplit - 1) / a.nsplit;
  const long long sb = (long long)split * per;
  const long long se = sb + per < a.n ? sb + per : a.n;

  // block-uniform: every period of this block keeps |(t - t0) / P| small enough for the fixed-point phase
  const bool fast = __syncthreads_and(in_range) != 0;
  // block-uniform: even |t / P| is small for every period of the block (stamps near zero): no shift, the magic number is
  // a compile-time constant and the packed loop is the round-1 / round-2 tuned code; otherwise (Julian dates) the shifted
  // form with one magic number per period and a guard band sized by |t / P| (phase_common.cuh)
  const bool plain = __syncthreads_and(plain_ok) != 0;
  __shared__ unsigned s_guard;
  if (threadIdx.x == 0) s_guard = 0u;
  __syncthreads();
  {
    unsigned g = 0u;
#pragma unroll
    for (int s = 0; s < PPT; ++s)
      if (valids[s] && fast) g = max(g, pdm_guard_units(fabs(rPs[s]), a.meta->t_absmax));
    atomicMax(&s_guard, g);
  }
  __syncthreads();
  const unsigned guard_units = plain ? PDM_FAST_GUARD : s_guard;   // the widest any period of the block needs
  const double t0 = plain ? 0.0 : a.meta->t0;                      // pdm_fast_magic(0, rP, 4) == PDM_FAST_MAGIC_G exactly
  long long tile0 = 0;   // first sample of the tile being processed (the exact path re-reads the ORIGINAL stamps)
  const int pack_q = a.meta->pack_q;
  const bool packed = fast && !clamp_bins && pack_q >= 0 && a.allow_packed != 0;   // block-uniform
#if !PDM_L2_INT
  const float unpack = packed ? 1.0f / (float)(1u << pack_q) : 0.f;
#endif
  const unsigned* s_xq = reinterpret_cast<const unsigned*>(s_x);
  const unsigned m0u = (unsigned)m0, guard = guard_units * m0u;
  // x' = (x - mean) / std(ddof=1) is formed while the tile is staged: as a float for the FP32 columns, as the packed
  // increment (one unit of the count field + signed fixed-point x') for the packed ones
  const double x_mean = a.meta->mean, x_inv_sd = a.meta->inv_sd;
  const double x_scale = packed ? (double)(1u << pack_q) : 0.0;

  // With finite inputs the bin index is always in range (phi == 1.0 is caught by the ambiguity test and
  // fixed by pdm_fix_bin), so the guard is compiled in only for the SAFE variant used when some
  // sample is NaN / inf: a NaN phase is in no bin (every comparison of phase.py:138-140 is false).
  auto update = [&](auto safe, float2* col, unsigned k, double phi, float xv) {
    if (decltype(safe)::value) {
      if (!(phi == phi)) return;
      k = min(k, kmax);
    }
    float2 h = col[k * VT];
    h.x += 1.0f;
    h.y += xv;
    col[k * VT] = h;
  };
  auto exact_bin = [&](double P, double rP, double tv, double& phi) {
    unsigned e;
    int k = pdm_bin(tv, P, rP, m0d, phi, e);
    if (e < PDM_AMBIG) k = pdm_fix_bin(k, phi, s_thr, m0);
    return (unsigned)k;
  };
  // Packed path: the private column is a 32-bit word per bin, (count << 23) + sum of fixed-point x', so one
  // sample costs one shared-memory integer add; every PDM_PACK_FLUSH samples the words are unpacked into the
  // float2 columns.
#if PDM_L2_INT
  // Second level of the packed path: two integer planes over the memory of the float2 columns, count[bin][VT] and
  // sum[bin][VT] (fixed point, exact), fed with three shared-memory atomics per bin (fetch-and-clear of the packed
  // word, two adds without return value) instead of LDS.32 + LDS.64 + STS.64 + STS.32.
  unsigned* cnt2 = reinterpret_cast<unsigned*>(hist);
  int* sum2 = reinterpret_cast<int*>(hist) + (size_t)m0 * VT;
  auto flush32 = [&](int column, unsigned* col32) {
    for (int b = 0; b < m0; ++b) {
#if PDM_FEED_PLAIN
      const unsigned w = col32[b * VT];
      col32[b * VT] = 0u;
      const int sfix = ((int)(w << PDM_CNT_BITS)) >> PDM_CNT_BITS;
      const unsigned cnt = (w - (unsigned)sfix) >> PDM_SUM_BITS;
      cnt2[b * VT + column] += cnt;
      sum2[b * VT + column] += sfix;
#else
      const unsigned w = atomicExch(col32 + b * VT, 0u);
      const int sfix = ((int)(w << PDM_CNT_BITS)) >> PDM_CNT_BITS;   // low PDM_SUM_BITS bits, sign extended
      const unsigned cnt = (w - (unsigned)sfix) >> PDM_SUM_BITS;
      atomicAdd(cnt2 + b * VT + column, cnt);
      atomicAdd(sum2 + b * VT + column, sfix);
#endif
    }
  };
#else
  auto flush32 = [&](int column, unsigned* col32) {
    float2* col = hist + column;
    for (int b = 0; b < m0; ++b) {
      const unsigned w = col32[b * VT];
      const int sfix = ((int)(w << PDM_CNT_BITS)) >> PDM_CNT_BITS;
      const unsigned cnt = (w - (unsigned)sfix) >> PDM_SUM_BITS;
      float2 h = col[b * VT];
      h.x += (float)cnt;
      h.y = fmaf((float)sfix, unpack, h.y);
      col[b * VT] = h;
      col32[b * VT] = 0u;
    }
  };
#endif
#ifndef PDM_PACK_ATOMIC
#define PDM_PACK_ATOMIC 1
#endif
  // The column is private, so the integer add needs no atomicity -- but a native shared-memory integer
  // atomic without a return value (ATOMS.ADD) is ONE fire-and-forget instruction instead of the dependent
  // LDS -> IADD -> STS chain, and same-bin updates of consecutive samples stay ordered in the memory pipe.
  auto add32 = [&](unsigned* col32, unsigned k, unsigned inc) {
#if PDM_PACK_ATOMIC
    // keep the bin index (high word of the 32 x 32 -> 64 bit product) opaque: otherwise the compiler folds the
    // column scaling into the 64-bit product (SHF.R.U64 + LOP3 + IADD per sample) instead of one IMAD on the high word
    asm volatile("" : "+r"(k));
    atomicAdd(col32 + k * VT, inc);
#else
    col32[k * VT] += inc;
#endif
  };
  // One sample of the packed path with the edge test (tail of a chunk, and the rare slow trips).
  auto packed_one = [&](auto sc, int i, unsigned guard2) {
    constexpr int s = decltype(sc)::value;
    unsigned p0;
    unsigned k0 = pdm_bin_fast_m(s_t[i], rPs[s], pdm_fast_magic(t0, rPs[s], guard_units), m0u, p0);   // (rare path: recomputed, not kept live)
    double ph;
    if (p0 < guard2) k0 = exact_bin(Ps[s], rPs[s], a.t[tile0 + i], ph);
    add32(hist32 + s * THREADS + threadIdx.x, k0, s_xq[i]);
  };
  auto packed_one_all = [&](int i, unsigned guard2) {
    static_for<PPT>([&](auto sc) { packed_one(sc, i, guard2); });
  };
  // Rare fix-up of one sample whose fast-path update has already been issued: if the exact bin differs from the
  // fast one, move the increment (the packed word is a sum modulo 2^32, so adding -inc undoes the update exactly).
  auto fixup_one = [&](auto sc, int i, unsigned guard2) {
    constexpr int s = decltype(sc)::value;
    unsigned p0;
    const unsigned kf = pdm_bin_fast_m(s_t[i], rPs[s], pdm_fast_magic(t0, rPs[s], guard_units), m0u, p0);
    if (p0 < guard2) {
      double ph;
      const unsigned ke = exact_bin(Ps[s], rPs[s], a.t[tile0 + i], ph);
      if (ke != kf) {
        const unsigned inc = s_xq[i];
        add32(hist32 + s * THREADS + threadIdx.x, kf, 0u - inc);
        add32(hist32 + s * THREADS + threadIdx.x, ke, inc);
      }
    }
  };
  auto fixup_one_all = [&](int i, unsigned guard2) {
    static_for<PPT>([&](auto sc) { fixup_one(sc, i, guard2); });
  };
  auto tile_loop_packed = [&](auto shifted, int cnt) {
    constexpr int U = PPT == 1 ? 8 : PDM_TRIP_CHAINS / PPT;  // samples per trip: U * PPT independent DFMA -> IMAD.WIDE -> IMAD -> ATOMS chains, one edge test
    double mg[PPT];   // magic number per column: a compile-time constant unless the stamps are shifted
#pragma unroll
    for (int s = 0; s < PPT; ++s) mg[s] = decltype(shifted)::value ? pdm_fast_magic(t0, rPs[s], guard_units) : PDM_FAST_MAGIC_G;
    const unsigned guard2 = 2u * guard;
    unsigned* c32 = hist32 + threadIdx.x;
    for (int w0 = 0; w0 < cnt; w0 += PDM_PACK_FLUSH) {
     // block-uniform: may this window be fed whole?  (always, in the 256-sample layout)
     int piece = PDM_PACK_FLUSH;
     if (PDM_PACK_FLUSH > 256) {
       unsigned wabs = 0;
#pragma unroll
       for (int wp = 0; wp < THREADS / 32; ++wp) wabs += s_winabs[wp][w0 / PDM_PACK_FLUSH];
       if (wabs >= (1u << (PDM_SUM_BITS - 1))) piece = PDM_PACK_SUB;
     }
     const int w1 = w0 + PDM_PACK_FLUSH < cnt ? w0 + PDM_PACK_FLUSH : cnt;
     for (int c0 = w0; c0 < w1; c0 += piece) {
      const int c1 = c0 + piece < w1 ? c0 + piece : w1;
      int i = c0;
      double tv[U];
#if PDM_PREFETCH
      // the time stamps of the next trip are fetched before this trip's atomics (the compiler cannot move a
      // shared-memory load across them); the read past the last trip lands in the tile's padding and is unused
#pragma unroll
      for (int u = 0; u < U; u += 2) {
        const double2 tt = *reinterpret_cast<const double2*>(s_t + i + u);
        tv[u] = tt.x;
        tv[u + 1] = tt.y;
      }
#endif
      for (; i + U <= c1; i += U) {
        unsigned xv[U], k[PPT][U], pos;
#if !PDM_PREFETCH
#pragma unroll
        for (int u = 0; u < U; u += 2) {
          const double2 tt = *reinterpret_cast<const double2*>(s_t + i + u);
          tv[u] = tt.x;
          tv[u + 1] = tt.y;
        }
#endif
        unsigned pmin = 0xffffffffu;
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            k[s][u] = pdm_bin_fast_m(tv[u], rPs[s], mg[s], m0u, pos);
            pmin = min(pmin, pos);
          }
        }
#if PDM_PREFETCH
#pragma unroll
        for (int u = 0; u < U; u += 2) {
          const double2 tt = *reinterpret_cast<const double2*>(s_t + i + U + u);
          tv[u] = tt.x;
          tv[u + 1] = tt.y;
        }
#endif
#pragma unroll
        for (int u = 0; u < U; u += 4) {
          const uint4 xx = *reinterpret_cast<const uint4*>(s_xq + i + u);
          xv[u] = xx.x; xv[u + 1] = xx.y; xv[u + 2] = xx.z; xv[u + 3] = xx.w;
        }
#if PDM_EDGE_FIXUP
        // the updates go out at once with the fast bins (always in range); the rare trip with a sample on a bin
        // edge (about m0 * U * PPT * 4e-9 of them) re-bins exactly afterwards and moves the increment if needed
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int s = 0; s < PPT; ++s) add32(c32 + s * THREADS, k[s][u], xv[u]);
        }
        if (pmin < guard2) {
          for (int u = 0; u < U; ++u) fixup_one_all(i + u, guard2);
        }
#else
        if (pmin < guard2) {  // rare (about m0 * U * PPT * 4e-9 of the trips): some sample sits on a bin edge
          for (int u = 0; u < U; ++u) packed_one_all(i + u, guard2);
          continue;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int s = 0; s < PPT; ++s) add32(c32 + s * THREADS, k[s][u], xv[u]);
        }
#endif
      }
      for (; i < c1; ++i) packed_one_all(i, guard2);
#if PDM_PACK_ATOMIC
      __syncwarp();
#endif
#if PDM_FLUSH_BINS > 1 && PDM_L2_INT
      // all columns of the thread, PDM_FLUSH_BINS bins at a time: every fetch-and-clear is issued before the first add
      // that depends on one (the plain loop waits for each ATOMS.EXCH: 23 % of the kernel's stall samples, r01e capture)
      for (int b0 = 0; b0 < m0; b0 += PDM_FLUSH_BINS) {
        unsigned w[PPT][PDM_FLUSH_BINS];
#if PDM_FEED_PLAIN
        unsigned oc[PPT][PDM_FLUSH_BINS];
        int os[PPT][PDM_FLUSH_BINS];
#endif
#pragma unroll
        for (int j = 0; j < PDM_FLUSH_BINS; ++j) {
#pragma unroll
          for (int s = 0; s < PPT; ++s) {
#if PDM_FEED_PLAIN
            const bool in = b0 + j < m0;
            w[s][j] = in ? c32[s * THREADS + (b0 + j) * VT] : 0u;
            oc[s][j] = in ? cnt2[(b0 + j) * VT + s * THREADS + threadIdx.x] : 0u;
            os[s][j] = in ? sum2[(b0 + j) * VT + s * THREADS + threadIdx.x] : 0;
#else
            w[s][j] = b0 + j < m0 ? atomicExch(c32 + s * THREADS + (b0 + j) * VT, 0u) : 0u;
#endif
          }
        }
#pragma unroll
        for (int j = 0; j < PDM_FLUSH_BINS; ++j) {
#pragma unroll
          for (int s = 0; s < PPT; ++s) {
            if (b0 + j < m0) {
              const int sfix = ((int)(w[s][j] << PDM_CNT_BITS)) >> PDM_CNT_BITS;
              const unsigned cnt = (w[s][j] - (unsigned)sfix) >> PDM_SUM_BITS;
#if PDM_FEED_PLAIN
              c32[s * THREADS + (b0 + j) * VT] = 0u;
              cnt2[(b0 + j) * VT + s * THREADS + threadIdx.x] = oc[s][j] + cnt;
              sum2[(b0 + j) * VT + s * THREADS + threadIdx.x] = os[s][j] + sfix;
#else
              atomicAdd(cnt2 + (b0 + j) * VT + s * THREADS + threadIdx.x, cnt);
              atomicAdd(sum2 + (b0 + j) * VT + s * THREADS + threadIdx.x, sfix);
#endif
            }
          }
        }
      }
#else
#pragma unroll
      for (int s = 0; s < PPT; ++s) flush32(s * THREADS + threadIdx.x, c32 + s * THREADS);
#endif
     }
    }
  };
  // The unpacked paths serve one period column after the other (sc = which of the thread's PPT columns).
  auto tile_loop_fast = [&](auto safe, auto sc, int cnt) {
    constexpr bool SAFE = decltype(safe)::value;
    constexpr int s = decltype(sc)::value;
    const double P = Ps[s], rP = rPs[s], mg = pdm_fast_magic(t0, rP, guard_units);
    const unsigned guard2 = 2u * guard;
    float2* col = hist + s * THREADS + threadIdx.x;
    int i = 0;
    for (; i + 4 <= cnt; i += 4) {
      const double2 ta = *reinterpret_cast<const double2*>(s_t + i);
      const double2 tb = *reinterpret_cast<const double2*>(s_t + i + 2);
      const float4 xv = *reinterpret_cast<const float4*>(s_x + i);
      unsigned p0, p1, p2, p3;   // position inside the bin + guard (stamps are t - t0 here)
      unsigned k0 = pdm_bin_fast_m(ta.x, rP, mg, m0u, p0);
      unsigned k1 = pdm_bin_fast_m(ta.y, rP, mg, m0u, p1);
      unsigned k2 = pdm_bin_fast_m(tb.x, rP, mg, m0u, p2);
      unsigned k3 = pdm_bin_fast_m(tb.y, rP, mg, m0u, p3);
      double f0 = 0.0, f1 = 0.0, f2 = 0.0, f3 = 0.0;   // only the guarded variant looks at them (NaN = skip)
      if (SAFE) {
        f0 = ta.x - ta.x; f1 = ta.y - ta.y; f2 = tb.x - tb.x; f3 = tb.y - tb.y;   // NaN for NaN / inf stamps
      }
      if (min(min(p0, p1), min(p2, p3)) < guard2) {  // rare: a sample sits on a bin edge -> exact path, original stamp
        double ph;
        if (p0 < guard2) k0 = exact_bin(P, rP, a.t[tile0 + i], ph);
        if (p1 < guard2) k1 = exact_bin(P, rP, a.t[tile0 + i + 1], ph);
        if (p2 < guard2) k2 = exact_bin(P, rP, a.t[tile0 + i + 2], ph);
        if (p3 < guard2) k3 = exact_bin(P, rP, a.t[tile0 + i + 3], ph);
      }
      update(safe, col, k0, f0, xv.x);
      update(safe, col, k1, f1, xv.y);
      update(safe, col, k2, f2, xv.z);
      update(safe, col, k3, f3, xv.w);
    }
    for (; i < cnt; ++i) {
      unsigned p0;
      const double tv = s_t[i];
      unsigned k0 = pdm_bin_fast_m(tv, rP, mg, m0u, p0);
      double ph;
      if (p0 < guard2) k0 = exact_bin(P, rP, a.t[tile0 + i], ph);
      update(safe, col, k0, SAFE ? tv - tv : 0.0, s_x[i]);
    }
  };
  auto tile_loop = [&](auto safe, auto sc, int cnt) {
    constexpr int s = decltype(sc)::value;
    const double P = Ps[s], rP = rPs[s];
    float2* col = hist + s * THREADS + threadIdx.x;
    int i = 0;
    for (; i + 4 <= cnt; i += 4) {
      // four independent phase computations (FP64 chains overlap), then four updates in order
      const double2 ta = *reinterpret_cast<const double2*>(s_t + i);
      const double2 tb = *reinterpret_cast<const double2*>(s_t + i + 2);
      const float4 xv = *reinterpret_cast<const float4*>(s_x + i);
      double f0, f1, f2, f3;
      unsigned e0, e1, e2, e3;
      int k0 = pdm_bin(ta.x, P, rP, m0d, f0, e0);
      int k1 = pdm_bin(ta.y, P, rP, m0d, f1, e1);
      int k2 = pdm_bin(tb.x, P, rP, m0d, f2, e2);
      int k3 = pdm_bin(tb.y, P, rP, m0d, f3, e3);
      if (min(min(e0, e1), min(e2, e3)) < PDM_AMBIG) {  // rare: a sample sits on a bin edge
        if (e0 < PDM_AMBIG) k0 = pdm_fix_bin(k0, f0, s_thr, m0);
        if (e1 < PDM_AMBIG) k1 = pdm_fix_bin(k1, f1, s_thr, m0);
        if (e2 < PDM_AMBIG) k2 = pdm_fix_bin(k2, f2, s_thr, m0);
        if (e3 < PDM_AMBIG) k3 = pdm_fix_bin(k3, f3, s_thr, m0);
      }
      update(safe, col, (unsigned)k0, f0, xv.x);
      update(safe, col, (unsigned)k1, f1, xv.y);
      update(safe, col, (unsigned)k2, f2, xv.z);
      update(safe, col, (unsigned)k3, f3, xv.w);
    }
    for (; i < cnt; ++i) {
      double f0;
      unsigned e0;
      int k0 = pdm_bin(s_t[i], P, rP, m0d, f0, e0);
      if (e0 < PDM_AMBIG) k0 = pdm_fix_bin(k0, f0, s_thr, m0);
      update(safe, col, (unsigned)k0, f0, s_x[i]);
    }
  };
  auto tile_unpacked = [&](auto sc, int cnt) {
    if (fast) {
      if (clamp_bins) tile_loop_fast(std::true_type{}, sc, cnt);
      else tile_loop_fast(std::false_type{}, sc, cnt);
    } else {
      if (clamp_bins) tile_loop(std::true_type{}, sc, cnt);
      else tile_loop(std::false_type{}, sc, cnt);
    }
  };

  int tiles_since_flush = 0;
  tile0 = sb;
  do {
    long long left = se - tile0;
    const int cnt = left <= 0 ? 0 : (left < PDM_TILE ? (int)left : PDM_TILE);
    __syncthreads();
    if (packed && PDM_PACK_FLUSH > 256) {
      // staging + sum |increment| per feed window (decides, block-uniformly, whether the window may be fed whole)
      unsigned wabs[PDM_TILE / PDM_PACK_FLUSH];
#pragma unroll
      for (int k = 0; k < PDM_TILE / PDM_PACK_FLUSH; ++k) wabs[k] = 0u;
      static_assert(PDM_PACK_FLUSH % THREADS == 0 || THREADS > PDM_PACK_FLUSH, "a thread's staging stride stays inside windows");
#pragma unroll
      for (int k = 0; k < PDM_TILE / PDM_PACK_FLUSH; ++k) {
        for (int i = k * PDM_PACK_FLUSH + threadIdx.x; i < (k + 1) * PDM_PACK_FLUSH && i < cnt; i += THREADS) {
          s_t[i] = a.t[tile0 + i] - t0;   // packed implies fast: the fixed-point phase works on t - t0
          const int sfix = (int)rint((a.x[tile0 + i] - x_mean) * x_inv_sd * x_scale);
          reinterpret_cast<unsigned*>(s_x)[i] = (1u << PDM_SUM_BITS) + (unsigned)sfix;
          wabs[k] += (unsigned)(sfix < 0 ? -sfix : sfix);
        }
      }
#pragma unroll
      for (int k = 0; k < PDM_TILE / PDM_PACK_FLUSH; ++k) {
        unsigned v = wabs[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) s_winabs[threadIdx.x >> 5][k] = v;
      }
    } else {
      for (int i = threadIdx.x; i < cnt; i += THREADS) {
        s_t[i] = fast ? a.t[tile0 + i] - t0 : a.t[tile0 + i];   // the exact path keeps the original stamps
        const double v = (a.x[tile0 + i] - x_mean) * x_inv_sd;
        if (packed) reinterpret_cast<unsigned*>(s_x)[i] = (1u << PDM_SUM_BITS) + (unsigned)(int)rint(v * x_scale);
        else s_x[i] = (float)v;
      }
    }
    __syncthreads();

    if (packed) {
      if (plain) tile_loop_packed(std::false_type{}, cnt);
      else tile_loop_packed(std::true_type{}, cnt);
    } else {
      static_for<PPT>([&](auto sc) { tile_unpacked(sc, cnt); });
    }

    tile0 += PDM_TILE;
    ++tiles_since_flush;
    if (tiles_since_flush == PDM_FLUSH_TILES || tile0 >= se) {
      // merge this thread's columns into the planes shared by all sample splits: counts as 32-bit, sums as 64-bit fixed
      // point with 32 fraction bits.  Integer addition is associative, so the totals do not depend on the order in which
      // splits and flushes arrive (bit-reproducible) and one plane serves all splits (|sum x'| < n by Cauchy-Schwarz with
      // sum x'^2 = n - 1, so 2^32 n < 2^63 for any n this library can be given).
#if PDM_L2_INT && PDM_PACK_ATOMIC
      __syncwarp();   // the second level was fed with atomics without return value
#endif
#pragma unroll
      for (int s = 0; s < PPT; ++s) {
        if (!valids[s]) continue;
        const int column = s * THREADS + threadIdx.x;
        float2* col = hist + column;
        unsigned* gc = a.cnt_plane + pis[s];
        unsigned long long* gs = a.sum_plane + pis[s];
        for (int b = 0; b < m0; ++b) {
          unsigned hn;
          long long hs;
#if PDM_L2_INT
          if (packed) {   // integer planes (exact): count, fixed-point sum with pack_q fraction bits
            hn = cnt2[b * VT + column];
            hs = (long long)sum2[b * VT + column] << (32 - pack_q);
            cnt2[b * VT + column] = 0u;
            sum2[b * VT + column] = 0;
          } else
#endif
          {
            const float2 h = col[b * VT];
            hn = (unsigned)h.x;                                   // an integer <= 8192 held exactly by the float
            hs = __double2ll_rn((double)h.y * 4294967296.0);      // NaN -> 0x8000...: only with NaN values, where the epilogue writes NaN anyway
            col[b * VT] = make_float2(0.f, 0.f);
          }
          if (hn) atomicAdd(gc + (long long)b * a.np, hn);
          if (hs) atomicAdd(gs + (long long)b * a.np, (unsigned long long)hs);
        }
      }
      tiles_since_flush = 0;
    }
  } while (tile0 < se);
}

// FP64 epilogue, one thread per trial period: reads (and clears) the period's column of the count / sum planes,
// evaluates the statistic, stores it (to every rank's buffer in the fan-out variant), per-block arg-extremum; the last
// block reduces those to the call's (best, index).
//   PDC_STAT_PDM  theta of phase.py:145-149 (smaller is better)
//   PDC_STAT_AOV  the analysis-of-variance statistic of Schwarzenberg-Czerny (1989) -- a TODO of the reference
//                 (phase.py:11) -- from the same fine-bin histograms with nc = 1: Theta = [(N - r) / (r - 1)] * s1 / s2
//                 over the r populated bins, s1 = sum_b n_b (mean_b - mean)^2 (between bins), s2 = sum_b sum_i
//                 (x_i - mean_b)^2 (within bins); larger is better.
struct PdmEpiArgs {
  unsigned* cnt_plane;
  unsigned long long* sum_plane;
  const double* periods;
  const PdmMeta* meta;
  int m0, nc;
  long long np;
  double* theta_out;        // [np] or NULL
  double* red_val;          // [gridDim.x] per-block candidates
  long long* red_idx;
  unsigned* call_done;      // [1] epilogue blocks that have finished (self-resetting)
  long long* arg_out;       // or NULL
  double* best_out;         // or NULL
  pdc_fanout fan;           // fan.world == 0: no fan-out
  long long fan_offset;
};

template <int STAT>
__global__ void __launch_bounds__(256)
pdm_epilogue_kernel(const PdmEpiArgs a) {
  constexpr int SIGN = STAT == PDC_STAT_AOV ? +1 : -1;
  __shared__ double sv[32];
  __shared__ long long si[32];
  __shared__ int s_last;
  const long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int m0 = a.m0, nc = a.nc;
  const long long np = a.np;
  double theta = 0.0;
  long long idx = -1;
  if (pi < np) {
    unsigned* pn = a.cnt_plane + pi;
    unsigned long long* p1 = a.sum_plane + pi;
    const double P = a.periods[pi];
    const double q_binned = a.meta->q_binned;
    auto cnt_of = [&](int k) { return (double)__ldcg(pn + (long long)k * np); };
    auto sum_of = [&](int k) { return (double)(long long)__ldcg(p1 + (long long)k * np) * (1.0 / 4294967296.0); };
    if (STAT == PDC_STAT_AOV) {
      double sq = 0.0, ntot = 0.0, stot = 0.0;
      int r = 0;
      for (int k = 0; k < m0; ++k) {
        const double N = cnt_of(k), S = sum_of(k);
        if (N >= 1.0) {
          sq += S * S / N;
          ntot += N;
          stot += S;
          ++r;
        }
      }
      const double s1 = sq - stot * stot / ntot;   // between the bins, about the mean of the binned samples
      const double s2 = q_binned - sq;             // within the bins
      theta = ((ntot - (double)r) / (double)(r - 1)) * (s1 / s2);
      // fewer than two populated bins (r - 1 == 0) or no scatter inside the bins: 0/0-like, the same NaN class as
      // PDM's "every bin dropped" below -- never +-inf from accumulation residue
      if (r < 2 || !(s2 > 0.0)) theta = nan("");
      if (!isfinite(P) || !isfinite(1.0 / P)) theta = nan("");  // no phases: period 0, denormal, inf or NaN
    } else {
      double sq = 0.0, den = 0.0;
      for (int k = 0; k < m0; ++k) {
        double N = 0.0, S = 0.0;
        for (int c = 0; c < nc; ++c) {
          int q = k + c;
          if (q >= m0) q -= m0;
          N += cnt_of(q);
          S += sum_of(q);
        }
        if (N >= 1.0) sq += S * S / N;
        if (N > 1.0) den += N - 1.0;  // phase.py:142,147: bins with <= 1 sample are dropped
      }
      // sum_k (n_k - 1) s_k^2 = nc (N - 1) - sum S_k^2 / n_k in units of sigma^2 (see file header)
      theta = ((double)nc * q_binned - sq) / den;
      // every coarse bin holds <= 1 sample: the reference divides 0.0 by 0.0 (phase.py:145-147 with an empty `mj`)
      // and gets NaN, which its nan-aware reductions skip; accumulation residue in the numerator must not turn it into +-inf
      if (!(den > 0.0)) theta = nan("");
      if (isinf(P)) theta = 1.0;  // every phase is 0: one populated fine bin holding all samples
      else if (!isfinite(1.0 / P) || P != P) theta = nan("");  // period 0, denormal or NaN: phases are inf / NaN
    }
    if (!(a.meta->inv_sd == a.meta->inv_sd)) theta = nan("");   // NaN values poison sigma^2 and hence every theta (phase.py:165)
    for (int k = 0; k < m0; ++k) {                              // the planes are clean for the next call
      pn[(long long)k * np] = 0u;
      p1[(long long)k * np] = 0ull;
    }
    if (a.theta_out) a.theta_out[pi] = theta;
    // fused all-gather: the value goes to every rank's buffer over NVLink peer mappings
    for (int r = 0; r < a.fan.world; ++r) a.fan.power[r][a.fan_offset + pi] = theta;
    idx = pi;
  }
  block_argext<SIGN>(theta, idx, sv, si);
  const int nblk = gridDim.x;
  if (threadIdx.x == 0) {
    a.red_val[blockIdx.x] = theta;
    a.red_idx[blockIdx.x] = idx;
    __threadfence();
    s_last = atomicAdd(a.call_done, 1u) == (unsigned)(nblk - 1);
  }
  __syncthreads();
  if (!s_last) return;
  // the last block: final arg-extremum; NaN ignored, first occurrence (np.nanargmin / np.nanargmax, core.py:202-210)
  if (threadIdx.x == 0) *a.call_done = 0u;
  __threadfence();
  double best = 0.0;
  long long bidx = -1;
  for (int k = threadIdx.x; k < nblk; k += blockDim.x) {
    const double v = __ldcg(a.red_val + k);
    const long long i = __ldcg(a.red_idx + k);
    if (better<SIGN>(v, i, best, bidx)) { best = v; bidx = i; }
  }
  block_argext<SIGN>(best, bidx, sv, si);
  if (threadIdx.x == 0) {
    const double val = bidx >= 0 ? best : nan("");
    if (a.arg_out) *a.arg_out = bidx;
    if (a.best_out) *a.best_out = val;
    for (int r = 0; r < a.fan.world; ++r) {   // slot `rank` of every rank's candidate table: (best, GLOBAL index)
      a.fan.best[r][2 * a.fan.rank] = val;
      a.fan.best[r][2 * a.fan.rank + 1] = bidx >= 0 ? (double)(bidx + a.fan_offset) : -1.0;
    }
  }
}

static size_t pdm_smem_bytes(int m0, int threads) {
  return sizeof(double) * (PDM_TILE + PDM_TILE_PAD + ((m0 + 2) & ~1)) + sizeof(float) * PDM_TILE +
         (sizeof(float2) + sizeof(unsigned)) * (size_t)m0 * threads;
}

template <int THREADS, int PPT>
static int pdm_launch(pdc_ctx* ctx, const PdmArgs& a, size_t smem, long long blocks, cudaStream_t st) {
  PDC_CUDA(cudaFuncSetAttribute(pdm_hist_kernel<THREADS, PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pdm_hist_kernel<THREADS, PPT><<<(unsigned)blocks, THREADS, smem, st>>>(a);
  PDC_CUDA(cudaGetLastError());
  ctx->launches++;
  return PDC_OK;
}

int pdm_run(pdc_ctx* ctx, const double* t, const double* x, int64_t n, const double* periods,
            int64_t np, int nb, int nc, double* theta_out, int64_t* argmin_out, double* min_out,
            cudaStream_t st, const pdc_fanout* fanout, int64_t fan_offset, int statistic) {
  if (statistic != PDC_STAT_PDM && statistic != PDC_STAT_AOV) { set_error("pdc_pdm: unknown statistic %d", statistic); return PDC_EINVAL; }
  if (statistic == PDC_STAT_AOV && (nc != 1 || fanout)) { set_error("pdc_aov: needs nc == 1 and no fan-out"); return PDC_EINVAL; }
  if (fanout && (fanout->world < 1 || fanout->world > PDC_MAX_PEERS || fanout->rank < 0 ||
                 fanout->rank >= fanout->world)) {
    set_error("pdc_pdm_dev_fanout: needs 1 <= world <= %d", PDC_MAX_PEERS);
    return PDC_EINVAL;
  }
  if (n < 2) { set_error("pdc_pdm: need at least 2 samples"); return PDC_EINVAL; }
  if (np < 1) { set_error("pdc_pdm: need at least one trial period"); return PDC_EINVAL; }
  if (nb < 1 || nc < 1) { set_error("pdc_pdm: nb and nc must be >= 1"); return PDC_EINVAL; }
  const long long m0l = (long long)nb * nc;
  const size_t smem_max = 227 * 1024;
  if (m0l > 100000 || pdm_smem_bytes((int)m0l, 32) > smem_max) {
    set_error("pdc_pdm: nb*nc = %lld fine bins do not fit a shared-memory histogram (max %d)",
              m0l, (int)((smem_max - 13 * 1024) / (12 * 32)));
    return PDC_EINVAL;
  }
  const int m0 = (int)m0l;

  // period columns per block (VT): the candidate that keeps the most columns resident per SM
  int vt = 32, best_res = 0;
  const int cands[4] = {256, 128, 64, 32};
  for (int c = 0; c < 4; ++c) {
    size_t sm = pdm_smem_bytes(m0, cands[c]);
    if (sm > smem_max) continue;
    int blocks = (int)((228 * 1024) / (sm + 1024));
    if (blocks > 2048 / cands[c]) blocks = 2048 / cands[c];
    int res = blocks * cands[c];
    if (res > best_res) { best_res = res; vt = cands[c]; }
  }
  // Trial periods per thread: long curves run the packed path (decided on the device, PDM_PACK_MIN_N), where two
  // columns per thread share every tile read; short curves keep one column per thread (more threads per SM).
  // AoV never runs the packed path (see PdmArgs::allow_packed), so it keeps one column per thread as well.
  int ppt = (n >= PDM_PACK_MIN_N && vt >= 64 && statistic == PDC_STAT_PDM) ? 2 : 1;
  {
    const int o = ctx->pdm_ppt_override;
    if (o == 1 || (o == 2 && vt >= 64)) ppt = o;
  }
  const size_t smem = pdm_smem_bytes(m0, vt);
  const long long resident = (long long)ctx->sm_count * (best_res / vt);
  const long long npb = (np + vt - 1) / vt;

  // sample split: same cost model as GLS (per-item overhead ~ one histogram flush)
  int nsplit = 1;
  if (npb < 24 * resident) {
    long long cap = n / 512;
    if (cap < 1) cap = 1;
    if (cap > 1024) cap = 1024;
    double best = 1e300;   // (all splits share one pair of planes: no memory cost per split)
    for (long long s = 1; s <= cap; ++s) {
      long long items = npb * s;
      long long waves = (items + resident - 1) / resident;
      double cost = (double)waves * ((double)((n + s - 1) / s) + 2.0 * m0 + 64.0);
      if (cost < best * 0.999) { best = cost; nsplit = (int)s; }
      if (items > 64 * resident) break;
    }
  }
  const long long blocks = npb * nsplit;
  if (blocks > 0x7fffffffLL || npb > 0x3fffffffLL) { set_error("pdc_pdm: problem too large for one call"); return PDC_EINVAL; }

  ScratchScope scratch(ctx, st);
  PDC_TRY(scratch.acquire());
  PDC_TRY(ctx->pdm_meta.reserve(sizeof(PdmMeta) + 16 + sizeof(PdmPart) * PDM_STATS_MAXBLK));
  // count plane (32-bit) + sum plane (64-bit fixed point) shared by all sample splits; the epilogue leaves them
  // cleared, so they are zeroed only when (re)allocated or after a call that failed before its epilogue
  const size_t plane_elems = (size_t)m0 * np;
  {
    const void* before = ctx->hist_plane.p;
    const size_t cap_before = ctx->hist_plane.cap;
    PDC_TRY(ctx->hist_plane.reserve((sizeof(unsigned long long) + sizeof(unsigned)) * plane_elems));
    if (ctx->hist_plane.p != before || ctx->hist_plane.cap != cap_before || ctx->hist_plane_dirty) {
      PDC_CUDA(cudaMemsetAsync(ctx->hist_plane.p, 0, ctx->hist_plane.cap, st));
      ctx->hist_plane_dirty = false;
    }
  }
  const int eblk = (int)((np + 255) / 256);
  PDC_TRY(ctx->blockred.reserve((sizeof(double) + sizeof(long long)) * (size_t)eblk));
  // completion counters (stats blocks, epilogue blocks): zeroed when allocated, every kernel leaves them at zero again
  {
    const void* before = ctx->pdm_cnt.p;
    PDC_TRY(ctx->pdm_cnt.reserve(sizeof(unsigned) * 4));
    if (ctx->pdm_cnt.p != before) PDC_CUDA(cudaMemsetAsync(ctx->pdm_cnt.p, 0, ctx->pdm_cnt.cap, st));
  }
  unsigned* cnt_stats = ctx->pdm_cnt.as<unsigned>();
  unsigned* cnt_call = cnt_stats + 1;

  PdmMeta* meta = ctx->pdm_meta.as<PdmMeta>();
  {
    PdmPart* part = reinterpret_cast<PdmPart*>(reinterpret_cast<char*>(meta) + ((sizeof(PdmMeta) + 15) & ~(size_t)15));
    long long sblk = (n + 8 * PDM_STATS_THREADS - 1) / (8 * PDM_STATS_THREADS);
    if (sblk > PDM_STATS_MAXBLK) sblk = PDM_STATS_MAXBLK;
    pdm_stats_kernel<<<(unsigned)sblk, PDM_STATS_THREADS, 0, st>>>(t, x, n, part, cnt_stats, meta);
    PDC_CUDA(cudaGetLastError());
    ctx->launches++;
  }

  PdmArgs a;
  a.t = t;
  a.x = x;
  a.periods = periods;
  a.meta = meta;
  a.sum_plane = ctx->hist_plane.as<unsigned long long>();
  a.cnt_plane = reinterpret_cast<unsigned*>(a.sum_plane + plane_elems);
  a.n = n;
  a.np = np;
  a.m0 = m0;
  a.nsplit = nsplit;
  a.allow_packed = statistic == PDC_STAT_PDM ? 1 : 0;

  ctx->hist_plane_dirty = true;   // until the epilogue that clears the planes has been enqueued
  PDC_TRY(ctx->main_begin(st));
  switch (vt * 8 + ppt) {
    case 256 * 8 + 1: PDC_TRY((pdm_launch<256, 1>(ctx, a, smem, blocks, st))); break;
    case 256 * 8 + 2: PDC_TRY((pdm_launch<128, 2>(ctx, a, smem, blocks, st))); break;
    case 128 * 8 + 1: PDC_TRY((pdm_launch<128, 1>(ctx, a, smem, blocks, st))); break;
    case 128 * 8 + 2: PDC_TRY((pdm_launch<64, 2>(ctx, a, smem, blocks, st))); break;
    case 64 * 8 + 1: PDC_TRY((pdm_launch<64, 1>(ctx, a, smem, blocks, st))); break;
    case 64 * 8 + 2: PDC_TRY((pdm_launch<32, 2>(ctx, a, smem, blocks, st))); break;
    default: PDC_TRY((pdm_launch<32, 1>(ctx, a, smem, blocks, st))); break;
  }
  PDC_TRY(ctx->main_end(st));

  PdmEpiArgs e;
  e.cnt_plane = a.cnt_plane;
  e.sum_plane = a.sum_plane;
  e.periods = periods;
  e.meta = meta;
  e.m0 = m0;
  e.nc = nc;
  e.np = np;
  e.theta_out = theta_out;
  e.red_val = ctx->blockred.as<double>();
  e.red_idx = reinterpret_cast<long long*>(e.red_val + eblk);
  e.call_done = cnt_call;
  e.arg_out = (long long*)argmin_out;
  e.best_out = min_out;
  if (fanout) e.fan = *fanout;
  else memset(&e.fan, 0, sizeof(e.fan));
  e.fan_offset = fan_offset;
  if (statistic == PDC_STAT_AOV) pdm_epilogue_kernel<PDC_STAT_AOV><<<(unsigned)eblk, 256, 0, st>>>(e);
  else pdm_epilogue_kernel<PDC_STAT_PDM><<<(unsigned)eblk, 256, 0, st>>>(e);
  PDC_CUDA(cudaGetLastError());
  ctx->launches++;
  ctx->hist_plane_dirty = false;
  PDC_TRY(scratch.release());
  return PDC_OK;
}

}  // namespace pdc
