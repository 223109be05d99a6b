// Generalised Lomb-Scargle periodogram on a uniform frequency grid, exact sums.
//
// Replaces the three `_trig_sum` calls and the numpy epilogue of
// `GLS.__call__` (reference src/periodicity/spectral.py:109-132).  The reference
// approximates
//     S_j = sum_i w_i sin(2 pi f_j t_i),  C_j = sum_i w_i cos(2 pi f_j t_i)
// (spectral.py:12-16) with an extirpolation + FFT; this file evaluates the sums
// themselves.
//
// Interpretation of "angle-addition recurrence reseeded every K samples"
// (BASELINE.json north_star; SURVEY.md §7 hard part 5): samples are irregular, so
// there is no constant rotation along the sample axis.  The frequency grid IS
// uniform, f_j = fmin + j df (spectral.py:36,97), hence for a fixed sample i
//     exp(2 pi i f_{j+1} t_i) = exp(2 pi i f_j t_i) * exp(2 pi i df t_i)
// and the rotation exp(2 pi i df t_i) depends on the sample only.  A thread owns a
// strip of K consecutive frequencies; for every sample it seeds (cos, sin) at the
// first frequency of its strip EXACTLY (phase reduced mod 1 in FP64, then
// MUFU sin/cos) and rotates K-1 times in FP32.  "K" = strip length = reseed period.
//
// Kernels (launch order; three launches per call since round 2, eight in round 1):
//   gls_stats_kernel    per curve: t_min, t_max, sum w, weighted mean, YY in ONE pass (FP64, moments about the
//                       curve's first value), multi-block partials; the last block to finish a curve reduces
//                       them in a fixed order and writes the curve's derived parameters
//   gls_prep_kernel     two block roles in one launch: per sample (t - t_min, frac(df (t - t_min))) as double2 and
//                       (cos, sin of 2 pi df (t - t_min), y', w') as float4; FP64 direct sums for the
//                       frequencies with < 1 cycle over the baseline
//   gls_strip_kernel    the hot kernel: six FP32 sums per frequency {C, S, YC, YS, CC, CS}; every 1024-sample tile
//                       the sums are converted to 64-bit fixed point (2^-30) and added to ONE plane of partial sums
//                       with RED.ADD.64 -- integer addition is associative, so the result does not depend on
//                       the order in which sample splits and tiles arrive (bit-reproducible), and the partial
//                       planes of round 1 (one FP64 plane per sample split, 86 MB on C2) shrink to 4.8 MB
//   gls_epilogue_kernel FP64: tau-offset algebra (spectral.py:113-132) per frequency, power store (to every
//                       rank's buffer in the fan-out variant), clears the plane for the next call, per-block
//                       NaN-aware arg-max; the last block of a curve reduces those to the curve's (max, argmax)
//
// Algorithmic work of the hot kernel: per sample*frequency evaluation
// 4 FP32 instructions for the rotation + 6 for the sums (7 with weights).
#include <cstring>
#include <type_traits>

#include "gls_common.cuh"

// Three-term strip: source order of the eight statements of one (sample, frequency) step.  ptxas
// derives its register allocation and schedule from it, and on B200 the resulting register-file
// operand conflicts ("dispatch" stalls) move the kernel time between 1.90 and 2.47 ms on the C2
// shape.  The default is the best of ~250 orders timed on B200 for BOTH the plain and the weighted
// kernel (tools/build_variants.sh + tools/tune_strip.py; profiles/r01/tune_strip_r01.txt).
#ifndef PDC_TT_ORDER
#define PDC_TT_ORDER ST_CC ST_YC ST_RC ST_YS ST_RS ST_S ST_CS ST_C
#endif
#ifndef PDC_TT_ORDER_W  /* weighted kernel (all eight statements are FFMA) */
#define PDC_TT_ORDER_W PDC_TT_ORDER
#endif
#ifndef PDC_GEOM0_K
#define PDC_GEOM0_K 16  /* strip length of the default geometry */
#endif

namespace pdc {

// ---------------------------------------------------------------------------
// per-curve statistics: grid (G, B).  A long single curve is read by G blocks (one block would need
// ~30 us for 65,000 samples), a batch uses G = 1; partials live in a fixed layout and every consumer
// reduces them in the same order, so the result does not depend on scheduling.
// ---------------------------------------------------------------------------
struct GlsPart {
  double tmin, tneg, sw, swd, swdd;   // moments of d = y - y[first sample of the curve]
  int bad, pad_;                      // some t, y or w of the curve is NaN / inf
};
constexpr int GLS_STATS_THREADS = 256;

// `single`: a one-curve call passes its host-known fields by value (no metadata upload); batches read curves[].
__global__ void __launch_bounds__(GLS_STATS_THREADS)
gls_stats_kernel(const double* __restrict__ t, const double* __restrict__ y, const double* __restrict__ w,
                 GlsCurve* curves, GlsPart* part, unsigned* done, const GlsCurve single, int use_single,
                 unsigned flags, long long j0, long long nf, int allow_three_term, int low_cap) {
  __shared__ double scratch[32 * 5];
  __shared__ int s_last;
  const int curve = blockIdx.y, G = gridDim.x;
  const GlsCurve cin = use_single ? single : curves[curve];
  const long long b = cin.begin, n = cin.n;
  const double y0 = y[b];
  double tmin = INFINITY, tneg = INFINITY, sw = 0.0, swd = 0.0, swdd = 0.0;
  int bad = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)G * blockDim.x) {
    const double ti = t[b + i], d = y[b + i] - y0, wi = w ? w[b + i] : 1.0;
    bad |= !isfinite(ti) || !isfinite(d) || !isfinite(wi);
    tmin = fmin(tmin, ti);
    tneg = fmin(tneg, -ti);
    sw += wi;
    const double wd = wi * d;
    swd += wd;
    swdd = fma(wd, d, swdd);
  }
  {
    double sums[3] = {sw, swd, swdd}, maxs[2] = {-tmin, -tneg};
    block_reduce_many<3, 2>(sums, maxs, scratch);
    sw = sums[0]; swd = sums[1]; swdd = sums[2]; tmin = -maxs[0]; tneg = -maxs[1];
  }
  bad = __syncthreads_or(bad);
  if (threadIdx.x == 0) {
    GlsPart& p = part[(long long)curve * G + blockIdx.x];
    p.tmin = tmin; p.tneg = tneg; p.sw = sw; p.swd = swd; p.swdd = swdd; p.bad = bad;
    __threadfence();
    s_last = atomicAdd(done + curve, 1u) == (unsigned)(G - 1);
  }
  __syncthreads();
  if (!s_last || threadIdx.x >= 32) return;
  // the last block of this curve: fixed-order (tree over the block index) reduction of the G <= 32 partials
  __threadfence();
  const int lane = threadIdx.x;
  const GlsPart* p = part + (long long)curve * G;
  tmin = lane < G ? __ldcg(&p[lane].tmin) : INFINITY;
  tneg = lane < G ? __ldcg(&p[lane].tneg) : INFINITY;
  sw = lane < G ? __ldcg(&p[lane].sw) : 0.0;
  swd = lane < G ? __ldcg(&p[lane].swd) : 0.0;
  swdd = lane < G ? __ldcg(&p[lane].swdd) : 0.0;
  bad = __any_sync(0xffffffffu, lane < G ? __ldcg(&p[lane].bad) : 0);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tmin = fmin(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
    tneg = fmin(tneg, __shfl_xor_sync(0xffffffffu, tneg, o));
    sw += __shfl_xor_sync(0xffffffffu, sw, o);
    swd += __shfl_xor_sync(0xffffffffu, swd, o);
    swdd += __shfl_xor_sync(0xffffffffu, swdd, o);
  }
  if (lane == 0) {
    GlsCurve cv = cin;
    const double tmax = -tneg;
    // spectral.py:102-108: w /= w.sum(); y = values - dot(w, values) if fit_mean.  In terms of d = y - y0:
    // mean = y0 + swd / sw;  sum w (y - mean)^2 = swdd - swd^2 / sw;  sum w y^2 = swdd + 2 y0 swd + y0^2 sw.
    const bool fit_mean = flags & PDC_GLS_FIT_MEAN;
    const double ymean = fit_mean ? y0 + swd / sw : 0.0;
    double syy = fit_mean ? swdd - swd * (swd / sw) : fma(y0, fma(y0, sw, 2.0 * swd), swdd);
    if (syy < 0.0) syy = 0.0;
    const double yy = syy / sw;  // spectral.py:120  YY = dot(w, y**2)
    cv.tmin = tmin;
    cv.tmax = tmax;
    int lb, lc;
    gls_low_range(cv.fmin, cv.df, j0, nf, tmax - tmin, lb, lc, low_cap);
    cv.low_begin = lb;
    cv.low_count = lc;
    cv.wsum = sw;
    cv.ymean = ymean;
    cv.yy = yy;
    cv.inv_rms = yy > 0.0 ? rsqrt(yy) : 0.0;
    const double span = cv.df * (tmax - tmin);  // turns swept by the per-index step angle over the samples
    const bool tt = allow_three_term && cv.df > 0.0 && span <= GLS_TT_MAX_SPAN;
    cv.three_term = tt;
    cv.gamma = tt ? 0.25 - 0.5 * span : 0.0;
    // non-finite input: the sums leave the strip kernel as fixed point, which cannot carry a NaN, so the epilogue
    // writes the NaN the reference's float arithmetic would have produced for every frequency of this curve
    cv.bad = bad;
    curves[curve] = cv;
    done[curve] = 0u;  // self-resetting: the next call finds the counter at zero
  }
}

// ---------------------------------------------------------------------------
// per-sample records + FP64 evaluation of the sub-cycle frequencies: one launch, two block roles
// ---------------------------------------------------------------------------
// grid = (rec_blocks + GLS_LOW_LANES * nlowchunk, curves).
//  * blocks x < rec_blocks write the sample records (grid-stride over the curve's samples);
//  * block x = rec_blocks + lane * nlowchunk + chunk sums chunk `chunk` of the samples for the sub-cycle frequencies
//    slot = lane, lane + GLS_LOW_LANES, ... < low_count in FP64 and adds them to a small fixed-point plane of their
//    own, lowplane[6][curve * low_cap + slot], which the epilogue prefers over the strip kernel's sums for those bins.
constexpr int GLS_LOW_LANES = 16;

__global__ void __launch_bounds__(256)
gls_prep_kernel(const double* __restrict__ t, const double* __restrict__ y, const double* __restrict__ w,
                const GlsCurve* __restrict__ curves, double2* __restrict__ rec1, float4* __restrict__ rec2,
                unsigned long long* __restrict__ lowplane, int B, int low_cap, double fix_scale,
                long long j0, int rec_blocks, int nlowchunk) {
  __shared__ double s_red[6][8];
  const int curve = blockIdx.y;
  const GlsCurve cv = curves[curve];
  if ((int)blockIdx.x < rec_blocks) {
    const double wscale = (double)cv.n / cv.wsum;  // weights rescaled to mean 1 (O(1) in FP32)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cv.n;
         i += (long long)rec_blocks * blockDim.x) {
      const long long g = cv.begin + i;
      const double tt = t[g] - cv.tmin;  // power is shift invariant; the reference shifts too (spectral.py:19-21)
      // step of the phase per frequency index, in turns.  gamma is the same for every sample of the
      // curve, i.e. a per-frequency phase origin, which the power does not depend on (the tau offset
      // of spectral.py:113-119 absorbs it).
      const double b = frac_of_product(cv.df, tt) + cv.gamma;
      double sb, cb;
      sincospi(2.0 * b, &sb, &cb);
      const double yv = (y[g] - cv.ymean) * cv.inv_rms;  // unit weighted RMS before the FP32 cast
      float4 r;
      rec_set(r, rec_slot(REC_CR), (float)cb);
      rec_set(r, rec_slot(REC_SR), (float)sb);
      if (w && cv.three_term) {
        // the three-term strip carries (sqrt(w') cos, sqrt(w') sin): every sum is then one FFMA
        const double sw = sqrt(w[g] * wscale);
        rec_set(r, rec_slot(REC_Y), (float)(sw * yv));
        rec_set(r, rec_slot(REC_W), (float)sw);
      } else if (w) {
        const double wn = w[g] * wscale;
        rec_set(r, rec_slot(REC_Y), (float)(wn * yv));
        rec_set(r, rec_slot(REC_W), (float)wn);
      } else {
        rec_set(r, rec_slot(REC_Y), (float)yv);
        rec_set(r, rec_slot(REC_W), 1.0f);
      }
      rec1[g] = make_double2(tt, b);
      rec2[g] = r;
    }
    return;
  }
  // ---- sub-cycle frequencies, FP64 ----
  const int q = (int)blockIdx.x - rec_blocks;
  const int chunk = q % nlowchunk, lane0 = q / nlowchunk;
  if (lane0 >= cv.low_count) return;  // block-uniform
  const long long per = (cv.n + nlowchunk - 1) / nlowchunk;
  const long long sb = (long long)chunk * per;
  const long long se = sb + per < cv.n ? sb + per : cv.n;
  const double wscale = (double)cv.n / cv.wsum;   // weights of mean 1: the same convention as the strip kernel's sums
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int slot = lane0; slot < cv.low_count; slot += GLS_LOW_LANES) {
    const double f = cv.fmin + (double)(j0 + cv.low_begin + slot) * cv.df;
    double a[6] = {0, 0, 0, 0, 0, 0};
    for (long long i = sb + threadIdx.x; i < se; i += blockDim.x) {
      const long long g = cv.begin + i;
      const double ph = frac_of_product(f, t[g] - cv.tmin);
      double sn, cs;
      sincospi(2.0 * ph, &sn, &cs);
      const double wi = w ? w[g] * wscale : 1.0;
      const double wy = wi * ((y[g] - cv.ymean) * cv.inv_rms);
      const double wc = wi * cs;
      a[0] += wc;
      a[1] = fma(wi, sn, a[1]);
      a[2] = fma(wy, cs, a[2]);
      a[3] = fma(wy, sn, a[3]);
      a[4] = fma(wc, cs, a[4]);
      a[5] = fma(wc, sn, a[5]);
    }
    // six block sums with one pair of barriers (fixed tree: deterministic)
#pragma unroll
    for (int k = 0; k < 6; ++k) a[k] = warp_sum(a[k]);
    __syncthreads();  // previous slot's s_red fully consumed
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 6; ++k) s_red[k][wid] = a[k];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
      double tot = 0.0;
      for (int k = 0; k < 8; ++k) tot += s_red[threadIdx.x][k];
      // 64-bit fixed point, order-independent integer adds over the sample chunks: lowplane[6][curve * low_cap + slot]
      atomicAdd(lowplane + (long long)threadIdx.x * B * low_cap + (long long)curve * low_cap + slot,
                (unsigned long long)__double2ll_rn(tot * fix_scale));
    }
  }
}

// ---------------------------------------------------------------------------
// the hot kernel
// ---------------------------------------------------------------------------
struct GlsMainArgs {
  const GlsCurve* curves;
  const double2* rec1;
  const float4* rec2;
  unsigned long long* partial;   // [6][nf_tot] fixed-point (2^-GLS_FIX_BITS) sums, all sample splits add into it
  long long nf;       // frequencies per curve in this call
  long long nf_tot;   // B * nf
  long long j0;       // absolute index of this call's first frequency
  int nfb;            // frequency blocks per curve
  int nsplit;         // sample splits per curve
  float fix_scale;    // 2^fix_bits
};

// Tile sums leave the hot kernels as 64-bit fixed point with fix_bits fraction bits.  With the weights rescaled to
// mean 1 and y' to unit weighted RMS every one of the six sums is bounded by n in magnitude (Cauchy-Schwarz:
// sum w'|y'| <= sqrt(sum w') sqrt(sum w' y'^2) = n), so fix_bits = 61 - ceil(log2 n) (gls_run) uses the whole word:
// the resolution relative to the sums' scale n is 2^-61, finer than float64's -- the FP64 sums of the sub-cycle bins
// travel through the same plane without loss, and for the FP32 tile sums the conversion is exact.
__device__ __forceinline__ void gls_flush(unsigned long long* p, float v, float fix_scale) {
  atomicAdd(p, (unsigned long long)__float2ll_rn(v * fix_scale));   // RED.ADD.64, no return value
}

template <int K, int THREADS, int MINB, bool WEIGHTED>
__global__ void __launch_bounds__(THREADS, MINB)
gls_strip_kernel(const GlsMainArgs a) {
  __shared__ __align__(16) double2 s_ab[GLS_TILE + 2];  // (phase at block base frequency, phase step per index)
  __shared__ __align__(16) float4 s_r2[GLS_TILE + 2];   // (cos, sin of the rotation, y' or w'y', w'); +2 look-ahead pad

  const int item = blockIdx.x;
  const int split = item % a.nsplit;
  const int rest = item / a.nsplit;
  const int fb = rest % a.nfb;
  const int curve = rest / a.nfb;

  const GlsCurve* cvp = a.curves + curve;
  const long long cbegin = cvp->begin, cn = cvp->n;
  const double fmin = cvp->fmin, df = cvp->df;

  const long long per = (cn + a.nsplit - 1) / a.nsplit;
  const long long sb = (long long)split * per;
  const long long se = sb + per < cn ? sb + per : cn;

  const long long jB = (long long)fb * (THREADS * K);
  const double fB = fmin + (double)(a.j0 + jB) * df;  // spectral.py:36: f = fmin + df * arange(nf)
  const int lK = threadIdx.x * K;
  const double lKd = (double)lK;
  const bool three_term = cvp->three_term != 0;  // block-uniform
  double gB = (double)(a.j0 + jB) * cvp->gamma;  // phase origin of the block's first frequency (turns)
  gB -= floor(gB);

  unsigned long long* pbase = a.partial + (long long)curve * a.nf + jB + lK;
  const long long jrem = a.nf - (jB + lK);  // strip entries with k < jrem are real frequencies

  if (threadIdx.x < 2) {
    s_ab[GLS_TILE + threadIdx.x] = make_double2(0.0, 0.0);
    s_r2[GLS_TILE + threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  long long tile0 = sb;
  do {
    long long left = se - tile0;
    const int cnt = left <= 0 ? 0 : (left < GLS_TILE ? (int)left : GLS_TILE);
    __syncthreads();  // previous tile fully consumed
    for (int i = threadIdx.x; i < cnt; i += THREADS) {
      const double2 r1 = a.rec1[cbegin + tile0 + i];
      s_ab[i] = make_double2(frac_of_product(fB, r1.x) + gB, r1.y);
      s_r2[i] = a.rec2[cbegin + tile0 + i];
    }
    __syncthreads();

    float aC[K], aS[K], aYC[K], aYS[K], aCC[K], aCS[K];
#pragma unroll
    for (int k = 0; k < K; ++k) aC[k] = aS[k] = aYC[k] = aYS[k] = aCC[k] = aCS[k] = 0.f;

    // One sample: K accumulations and K-1 steps along the frequency axis.  The sums are the
    // per-frequency {C, S, YC, YS, CC, CS}; (c, s) enters as the exact seed at the strip's first
    // frequency.  Two forms of the step (block-uniform choice, GlsCurve::three_term):
    //  * rotation        (c, s) <- (c cr - s sr, s cr + c sr)                  2 FMUL + 2 FFMA
    //  * three-term      c[k+1] = 2 cr c[k] - c[k-1], same for s               2 FFMA
    //    (Chebyshev recurrence; rounding errors grow by 1/|sin step|, bounded because the step
    //    angles were centred on a quarter turn by GlsCurve::gamma).  Being linear, it also carries
    //    a factor sqrt(w') for free, which turns every weighted sum into a single FFMA.
    auto accumulate = [&](int k, float c, float s, float yv, float wv) {  // rotation form
      if (WEIGHTED) {
        const float wc = wv * c;
        aC[k] += wc;
        aS[k] = fmaf(wv, s, aS[k]);
        aCC[k] = fmaf(wc, c, aCC[k]);
        aCS[k] = fmaf(wc, s, aCS[k]);
      } else {
        aC[k] += c;
        aS[k] += s;
        aCC[k] = fmaf(c, c, aCC[k]);
        aCS[k] = fmaf(c, s, aCS[k]);
      }
      aYC[k] = fmaf(yv, c, aYC[k]);
      aYS[k] = fmaf(yv, s, aYS[k]);
    };
    auto strip = [&](auto tt_tag, float c, float s, const float4 r2) {
      constexpr bool TT = decltype(tt_tag)::value;
      const float cr = rec_get(r2, rec_slot(REC_CR)), sr = rec_get(r2, rec_slot(REC_SR));
      const float yv = rec_get(r2, rec_slot(REC_Y)), wv = rec_get(r2, rec_slot(REC_W));
      (void)wv;
      if (TT) {
        if (WEIGHTED) {  // (c, s) carry sqrt(w'); the record holds yv = sqrt(w') y', wv = sqrt(w')
          c *= wv;
          s *= wv;
        }
        const float tc = cr + cr;
        float cp = c, sp = s;                       // index k - 1
        c = fmaf(cp, cr, -(sp * sr));               // index 1 by one rotation
        s = fmaf(sp, cr, cp * sr);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const float cc = k == 0 ? cp : c, ss = k == 0 ? sp : s;
          float cn = 0.f, sn = 0.f;
          const bool step = k >= 1 && k + 1 < K;
#define ST_C  if (WEIGHTED) aC[k] = fmaf(cc, wv, aC[k]); else aC[k] += cc;
#define ST_S  if (WEIGHTED) aS[k] = fmaf(ss, wv, aS[k]); else aS[k] += ss;
#define ST_YC aYC[k] = fmaf(cc, yv, aYC[k]);
#define ST_YS aYS[k] = fmaf(ss, yv, aYS[k]);
#define ST_CC aCC[k] = fmaf(cc, cc, aCC[k]);
#define ST_CS aCS[k] = fmaf(ss, cc, aCS[k]);
#define ST_RC if (step) cn = fmaf(c, tc, -cp);
#define ST_RS if (step) sn = fmaf(s, tc, -sp);
          if (WEIGHTED) { PDC_TT_ORDER_W } else { PDC_TT_ORDER }
#undef ST_C
#undef ST_S
#undef ST_YC
#undef ST_YS
#undef ST_CC
#undef ST_CS
#undef ST_RC
#undef ST_RS
          if (step) {
            cp = c;
            sp = s;
            c = cn;
            s = sn;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) {
          accumulate(k, c, s, yv, wv);
          if (k + 1 < K) {
            const float c2 = fmaf(c, cr, -(s * sr));
            const float s2 = fmaf(s, cr, c * sr);
            c = c2;
            s = s2;
          }
        }
      }
    };

    // Software pipeline, two samples per trip (ping-pong registers, no moves): the exact
    // seed of the next sample is computed while the current strip runs.  The tile arrays
    // carry two pad entries so the look-ahead never needs a bounds check.
    auto run_tile = [&](auto tt_tag) {
      float c0, s0, c1, s1;
      float4 ra = s_r2[0], rb;
      {
        const double2 ab = s_ab[0];
        gls_seed(ab.x, ab.y, lKd, c0, s0);
      }
      int i = 0;
      for (; i + 1 < cnt; i += 2) {
        {
          const double2 ab = s_ab[i + 1];
          rb = s_r2[i + 1];
          gls_seed(ab.x, ab.y, lKd, c1, s1);
        }
        strip(tt_tag, c0, s0, ra);
        {
          const double2 ab = s_ab[i + 2];
          ra = s_r2[i + 2];
          gls_seed(ab.x, ab.y, lKd, c0, s0);
        }
        strip(tt_tag, c1, s1, rb);
      }
      if (i < cnt) strip(tt_tag, c0, s0, ra);
    };
    if (cnt > 0) {
      if (three_term) run_tile(std::true_type{});
      else run_tile(std::false_type{});
    }

    // flush this tile's FP32 sums into the fixed-point plane (RED.ADD.64: no load latency to wait for, and integer
    // addition makes the total independent of the order in which splits and tiles arrive)
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (k < jrem) {
        unsigned long long* p = pbase + k;
        gls_flush(p, aC[k], a.fix_scale);
        gls_flush(p + a.nf_tot, aS[k], a.fix_scale);
        gls_flush(p + 2 * a.nf_tot, aYC[k], a.fix_scale);
        gls_flush(p + 3 * a.nf_tot, aYS[k], a.fix_scale);
        gls_flush(p + 4 * a.nf_tot, aCC[k], a.fix_scale);
        gls_flush(p + 5 * a.nf_tot, aCS[k], a.fix_scale);
      }
    }
    tile0 += GLS_TILE;
  } while (tile0 < se);
}

// ---------------------------------------------------------------------------
// arbitrary (non-uniform, user-supplied) frequency lists: pdc_gls_freqs
// ---------------------------------------------------------------------------
// The formula of spectral.py:113-132 does not need a uniform grid -- only the reference's FFT in `_trig_sum` does
// (spectral.py:11-40), which is why its GLS has no such option.  Without a uniform grid there is no recurrence along
// the frequency axis: a thread owns ONE frequency and seeds (cos, sin) exactly for every sample -- phase f (t - t_min)
// reduced mod 1 by ONE DFMA against a magic number (the product is exact inside the FMA, the sum's ulp is 2^-32 turn),
// I2F, FMUL, MUFU.SIN / MUFU.COS -- and accumulates the same six FP32 sums per 1024-sample tile, flushed to the same
// fixed-point plane and read by the same epilogue.  Bound: the MUFU pipe (2 per evaluation, 16 lanes per clk per SM).
// Frequencies with |f| T >= 2^19 turns reduce the phase with the FMA-residual product instead; frequencies with less
// than one cycle over the baseline (|f| T < 1) keep FP64 sums (the cancellation described in gls_common.cuh).
struct GlsFreeArgs {
  const GlsCurve* curves;    // one curve
  const double2* rec1;       // (t - t_min, unused)
  const float4* rec2;        // (unused, unused, y' or w'y', w')  -- rotation-form records (df == 0)
  const double* freqs;       // [nf]
  unsigned long long* partial;
  long long nf;
  int nsplit;
  float fix_scale;
};

constexpr int GLS_FREE_THREADS = 128;

template <bool WEIGHTED>
__global__ void __launch_bounds__(GLS_FREE_THREADS)
gls_free_kernel(const GlsFreeArgs a) {
  __shared__ double s_t[GLS_TILE];
  __shared__ __align__(16) float2 s_yw[GLS_TILE];
  const int split = blockIdx.x % a.nsplit;
  const long long fb = blockIdx.x / a.nsplit;
  const GlsCurve* cvp = a.curves;
  const long long cn = cvp->n;
  const double T = cvp->tmax - cvp->tmin;
  const long long j = fb * GLS_FREE_THREADS + threadIdx.x;
  const bool active = j < a.nf;
  const double f = active ? a.freqs[j] : 0.0;
  const double fT = fabs(f) * T;
  const bool lowf = fT < GLS_LOW_CYCLES;          // FP64 sums
  const bool wide = !(fT < 262144.0);             // 2^18: beyond the magic-number reduction (also NaN / inf)
  const long long per = (cn + a.nsplit - 1) / a.nsplit;
  const long long sb = (long long)split * per;
  const long long se = sb + per < cn ? sb + per : cn;
  const int yslot = rec_slot(REC_Y), wslot = rec_slot(REC_W);
  const float TWO_PI_32 = 1.4629180792671596e-9f;  // 2 pi / 2^32

  long long tile0 = sb;
  do {
    long long left = se - tile0;
    const int cnt = left <= 0 ? 0 : (left < GLS_TILE ? (int)left : GLS_TILE);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += GLS_FREE_THREADS) {
      s_t[i] = a.rec1[tile0 + i].x;
      const float4 r = a.rec2[tile0 + i];
      s_yw[i] = make_float2(rec_get(r, yslot), rec_get(r, wslot));
    }
    __syncthreads();
    float aC = 0.f, aS = 0.f, aYC = 0.f, aYS = 0.f, aCC = 0.f, aCS = 0.f;
    double dC = 0.0, dS = 0.0, dYC = 0.0, dYS = 0.0, dCC = 0.0, dCS = 0.0;
    if (lowf || wide) {
      for (int i = 0; i < cnt; ++i) {
        const double ph = frac_of_product(f, s_t[i]);
        const float2 yw = s_yw[i];
        if (lowf) {
          double sn, cs;
          sincospi(2.0 * ph, &sn, &cs);
          const double wv = WEIGHTED ? (double)yw.y : 1.0, wc = wv * cs;
          dC += wc;
          dS = fma(wv, sn, dS);
          dYC = fma((double)yw.x, cs, dYC);
          dYS = fma((double)yw.x, sn, dYS);
          dCC = fma(wc, cs, dCC);
          dCS = fma(wc, sn, dCS);
        } else {
          float sn, cs;
          __sincosf((float)ph * 6.2831853071795864f, &sn, &cs);
          const float wc = WEIGHTED ? yw.y * cs : cs;
          aC += wc;
          aS = WEIGHTED ? fmaf(yw.y, sn, aS) : aS + sn;
          aYC = fmaf(yw.x, cs, aYC);
          aYS = fmaf(yw.x, sn, aYS);
          aCC = fmaf(wc, cs, aCC);
          aCS = fmaf(wc, sn, aCS);
        }
      }
    } else {
#pragma unroll 4
      for (int i = 0; i < cnt; ++i) {
        const double v = __fma_rn(f, s_t[i], 1572864.0);      // 1.5 * 2^20: ulp(v) = 2^-32 turn
        const float x = (float)__double2loint(v) * TWO_PI_32;   // two's-complement fraction -> radians in [-pi, pi)
        float sn, cs;
        __sincosf(x, &sn, &cs);
        const float2 yw = s_yw[i];
        const float wc = WEIGHTED ? yw.y * cs : cs;
        aC += wc;
        aS = WEIGHTED ? fmaf(yw.y, sn, aS) : aS + sn;
        aYC = fmaf(yw.x, cs, aYC);
        aYS = fmaf(yw.x, sn, aYS);
        aCC = fmaf(wc, cs, aCC);
        aCS = fmaf(wc, sn, aCS);
      }
    }
    if (active && cnt > 0) {
      unsigned long long* p = a.partial + j;
      if (lowf) {
        const double sc = (double)a.fix_scale;
        atomicAdd(p, (unsigned long long)__double2ll_rn(dC * sc));
        atomicAdd(p + a.nf, (unsigned long long)__double2ll_rn(dS * sc));
        atomicAdd(p + 2 * a.nf, (unsigned long long)__double2ll_rn(dYC * sc));
        atomicAdd(p + 3 * a.nf, (unsigned long long)__double2ll_rn(dYS * sc));
        atomicAdd(p + 4 * a.nf, (unsigned long long)__double2ll_rn(dCC * sc));
        atomicAdd(p + 5 * a.nf, (unsigned long long)__double2ll_rn(dCS * sc));
      } else {
        gls_flush(p, aC, a.fix_scale);
        gls_flush(p + a.nf, aS, a.fix_scale);
        gls_flush(p + 2 * a.nf, aYC, a.fix_scale);
        gls_flush(p + 3 * a.nf, aYS, a.fix_scale);
        gls_flush(p + 4 * a.nf, aCC, a.fix_scale);
        gls_flush(p + 5 * a.nf, aCS, a.fix_scale);
      }
    }
    tile0 += GLS_TILE;
  } while (tile0 < se);
}

// ---------------------------------------------------------------------------
// FP64 epilogue: spectral.py:113-132 per frequency + arg-max
// ---------------------------------------------------------------------------
struct GlsEpiArgs {
  const GlsCurve* curves;
  unsigned long long* partial;  // [6][nf_tot]; read and cleared
  unsigned long long* lowplane; // [6][B * low_cap] FP64-accurate sums of the sub-cycle bins; read and cleared
  int low_cap, B;
  double inv_fix;               // 2^-fix_bits
  unsigned flags;
  long long nf, nf_tot, j0;
  double* power_out;            // [B * nf] or NULL
  double* red_val;              // [B * gridDim.x] per-block arg-max candidates
  long long* red_idx;
  unsigned* curve_done;         // [B] epilogue blocks of this curve that have finished (self-resetting)
  long long* arg_out;           // [B] or NULL
  double* max_out;              // [B] or NULL
  pdc_fanout fan;               // fan.world == 0: no fan-out
  int double_angle;             // planes 4, 5 hold sum w cos 2x, sum w sin 2x (gls_umma_kernel) instead of sum w cos^2 x, sum w cos x sin x
  const int* umma_status;       // non-NULL: gls_umma_kernel's protocol status; non-zero poisons the result with NaN
};

__global__ void __launch_bounds__(256)
gls_epilogue_kernel(const GlsEpiArgs a) {
  __shared__ double sv[32];
  __shared__ long long si[32];
  __shared__ int s_last;
  const int curve = blockIdx.y;
  const GlsCurve cv = a.curves[curve];
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double power = 0.0;
  long long idx = -1;
  if (j < a.nf) {
    double sums[6];
    double inv_n;
    unsigned long long* p = a.partial + (long long)curve * a.nf + j;
    unsigned long long raw[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) raw[q] = __ldcg(p + (long long)q * a.nf_tot);   // written with RED at L2
#pragma unroll
    for (int q = 0; q < 6; ++q) p[(long long)q * a.nf_tot] = 0ull;              // the plane is clean for the next call
    bool c2_direct = a.double_angle != 0;
    if (j >= cv.low_begin && j < cv.low_begin + cv.low_count) {
      // sub-cycle frequency: the FP64 sums of gls_prep_kernel replace the strip kernel's FP32 ones
      c2_direct = false;
      unsigned long long* lp = a.lowplane + (long long)curve * a.low_cap + (j - cv.low_begin);
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        raw[q] = __ldcg(lp + (long long)q * a.B * a.low_cap);
        lp[(long long)q * a.B * a.low_cap] = 0ull;
      }
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) sums[q] = (double)(long long)raw[q] * a.inv_fix;
    inv_n = 1.0 / (double)cv.n;
    power = gls_power_from_sums(sums, inv_n, a.flags, cv.yy, cv.psd_scale, c2_direct);
    if (cv.bad || (a.umma_status && *a.umma_status)) power = nan("");
    if (a.power_out) a.power_out[(long long)curve * a.nf + j] = power;
    // fused all-gather: the value goes to every rank's buffer over NVLink peer mappings
    for (int r = 0; r < a.fan.world; ++r) a.fan.power[r][a.j0 + j] = power;
    idx = j;
  }
  block_argext<+1>(power, idx, sv, si);
  const int nblk = gridDim.x;
  if (threadIdx.x == 0) {
    a.red_val[(long long)curve * nblk + blockIdx.x] = power;
    a.red_idx[(long long)curve * nblk + blockIdx.x] = idx;
    __threadfence();
    s_last = atomicAdd(a.curve_done + curve, 1u) == (unsigned)(nblk - 1);
  }
  __syncthreads();
  if (!s_last) return;
  // the last block of this curve: final (max, argmax); NaN ignored, first occurrence (np.nanargmax, core.py:202-205)
  if (threadIdx.x == 0) a.curve_done[curve] = 0u;
  __threadfence();
  double best = 0.0;
  long long bidx = -1;
  for (int k = threadIdx.x; k < nblk; k += blockDim.x) {
    const double v = __ldcg(a.red_val + (long long)curve * nblk + k);
    const long long i = __ldcg(a.red_idx + (long long)curve * nblk + k);
    if (better<+1>(v, i, best, bidx)) { best = v; bidx = i; }
  }
  block_argext<+1>(best, bidx, sv, si);
  if (threadIdx.x == 0) {
    const double val = bidx >= 0 ? best : nan("");
    if (a.arg_out) a.arg_out[curve] = bidx;
    if (a.max_out) a.max_out[curve] = val;
    for (int r = 0; r < a.fan.world; ++r) {   // slot `rank` of every rank's candidate table: (max, GLOBAL argmax)
      a.fan.best[r][2 * a.fan.rank] = val;
      a.fan.best[r][2 * a.fan.rank + 1] = bidx >= 0 ? (double)(bidx + a.j0) : -1.0;
    }
  }
}

// ---------------------------------------------------------------------------
// host-side launcher
// ---------------------------------------------------------------------------
// tensor-core formulation of the same sums (gls_umma.cu)
bool gls_umma_eligible(const pdc_ctx* ctx, int64_t B, int64_t nf, long long ntot, long long nmax, bool weighted,
                       const double* df_host);
int gls_umma_launch(pdc_ctx* ctx, const GlsCurve* curves, const double2* rec1, const float4* rec2,
                    unsigned long long* plane, int64_t B, int64_t nf, int64_t j0, long long nmax, bool weighted,
                    float fix_scale, cudaStream_t st);

// Strip-kernel geometries: K frequencies per thread, THREADS per block, MINB blocks per SM.
struct GlsGeom {
  int K, threads, minb;
};
static const GlsGeom kGlsGeoms[] = {
    {PDC_GEOM0_K, 128, 2},  // 0: default -- 142 registers, 3 blocks/SM; best of the 16 geometries tried in round 1
    {8, 64, 8},    // 1: small problems -- 512 frequencies per block so that tiny grids still fill the SMs
    {16, 128, 4},  // 2: capped at 128 registers (4 blocks/SM): ~4 % slower, more bank conflicts
    {20, 128, 2},  // 3: longer strips: fewer seeds per evaluation, ~2 % slower overall
    {12, 128, 2},  // 4
    {8, 256, 4},   // 5: 64 registers, 32 warps/SM: occupancy does not help this kernel
};
constexpr int kGlsNumGeoms = sizeof(kGlsGeoms) / sizeof(kGlsGeoms[0]);

template <int K, int THREADS, int MINB>
static int strip_blocks_per_sm(bool weighted) {
  int nb = 0;
  cudaError_t e = weighted
      ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gls_strip_kernel<K, THREADS, MINB, true>, THREADS, 0)
      : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gls_strip_kernel<K, THREADS, MINB, false>, THREADS, 0);
  if (e != cudaSuccess) { cudaGetLastError(); return MINB; }
  return nb > 0 ? nb : 1;
}

template <int K, int THREADS, int MINB>
static int launch_strip_t(pdc_ctx* ctx, const GlsMainArgs& a, bool weighted, long long items,
                          cudaStream_t st) {
  if (weighted) gls_strip_kernel<K, THREADS, MINB, true><<<(unsigned)items, THREADS, 0, st>>>(a);
  else gls_strip_kernel<K, THREADS, MINB, false><<<(unsigned)items, THREADS, 0, st>>>(a);
  PDC_CUDA(cudaGetLastError());
  ctx->launches++;
  return PDC_OK;
}

#define PDC_GLS_GEOM_CASES(X) \
  X(0, PDC_GEOM0_K, 128, 2) X(1, 8, 64, 8) X(2, 16, 128, 4) X(3, 20, 128, 2) X(4, 12, 128, 2) X(5, 8, 256, 4)

#ifdef PDC_ONLY_DEFAULT_GEOM
#undef PDC_GLS_GEOM_CASES
#define PDC_GLS_GEOM_CASES(X) X(0, PDC_GEOM0_K, 128, 2)
#endif

static int strip_occupancy(int geom, bool weighted) {
  switch (geom) {
#define X(i, k, t, m) case i: return strip_blocks_per_sm<k, t, m>(weighted);
    PDC_GLS_GEOM_CASES(X)
#undef X
  }
  return 1;
}

static int launch_strip(int geom, pdc_ctx* ctx, const GlsMainArgs& a, bool weighted, long long items,
                        cudaStream_t st) {
  switch (geom) {
#define X(i, k, t, m) case i: return launch_strip_t<k, t, m>(ctx, a, weighted, items, st);
    PDC_GLS_GEOM_CASES(X)
#undef X
  }
  set_error("bad strip geometry %d", geom);
  return PDC_EINVAL;
}

// Pick the sample split so that (curves * frequency blocks * nsplit) work items
// fill whole waves of resident blocks.  Cost model: every wave costs the samples
// of one item plus a fixed per-item overhead (staging, flush) worth ~48 samples.
static int choose_nsplit(long long base_items, long long nmax, long long resident, long long plane_bytes,
                         int min_samples) {
  if (base_items >= 24 * resident) return 1;
  long long cap = nmax / min_samples;
  if (cap < 1) cap = 1;
  if (cap > 4096) cap = 4096;
  if (plane_bytes > 0) {   // a caller whose splits own one plane of partial sums each: keep the scratch below 2 GiB
    const long long mem_cap = ((long long)2 << 30) / plane_bytes;
    if (cap > mem_cap) cap = mem_cap < 1 ? 1 : mem_cap;
  }
  double best = 1e300;
  int best_s = 1;
  for (long long s = 1; s <= cap; ++s) {
    long long items = base_items * s;
    long long waves = (items + resident - 1) / resident;
    double per = (double)((nmax + s - 1) / s) + 48.0;
    // measured on C2: a single wave loses ~2-4 % to uneven SM finish times; finer items even it out
    double cost = (double)waves * per * (1.0 + 0.04 / (double)waves);
    if (cost < best * 0.999) { best = cost; best_s = (int)s; }
    if (items > 64 * resident) break;
  }
  return best_s;
}

int gls_run(pdc_ctx* ctx, const double* t, const double* y, const double* w,
            const int64_t* offsets_host, int64_t B, const double* fmin_host, const double* df_host,
            int64_t j0, int64_t nf, unsigned flags, const double* psd_scale_host,
            double* power_out, int64_t* argmax_out, double* max_out, cudaStream_t st,
            const pdc_fanout* fanout, const double* freqs_dev) {
  if (freqs_dev && (B != 1 || fanout || j0 != 0)) { set_error("pdc_gls_freqs: one curve, no fan-out"); return PDC_EINVAL; }
  if (fanout && (B != 1 || fanout->world < 1 || fanout->world > PDC_MAX_PEERS || fanout->rank < 0 ||
                 fanout->rank >= fanout->world)) {
    set_error("pdc_gls_dev_fanout: needs one curve and 1 <= world <= %d", PDC_MAX_PEERS);
    return PDC_EINVAL;
  }
  if (B <= 0 || nf <= 0) { set_error("pdc_gls: need at least one curve and one frequency"); return PDC_EINVAL; }
  const long long ntot = offsets_host[B] - offsets_host[0];
  long long nmax = 0;
  for (int64_t b = 0; b < B; ++b) {
    long long nb = offsets_host[b + 1] - offsets_host[b];
    if (nb < 1) { set_error("pdc_gls: curve %lld is empty", (long long)b); return PDC_EINVAL; }
    if (!(df_host[b] == df_host[b]) || !(fmin_host[b] == fmin_host[b])) {
      set_error("pdc_gls: fmin/df of curve %lld is NaN", (long long)b);
      return PDC_EINVAL;
    }
    if (nb > nmax) nmax = nb;
  }
  if ((flags & PDC_GLS_PSD) && !psd_scale_host) { set_error("pdc_gls: PSD flag needs psd_scale"); return PDC_EINVAL; }
  const long long nf_tot = (long long)B * nf;

  // geometry of the hot kernel
  int geom = ctx->gls_geom;
  if (geom < 0 || geom >= kGlsNumGeoms) geom = 0;
  if (!ctx->gls_geom_forced) {
    // tiny problems: with 2048 frequencies per block the grid cannot fill 148 SMs even after
    // splitting the sample axis; use 512-frequency blocks of 64 threads instead
    const long long big_items = (long long)B * ((nf + 2047) / 2048) * (nmax / 256 > 0 ? nmax / 256 : 1);
    if (big_items < 2LL * ctx->sm_count) geom = 1;
  }
  const int K = kGlsGeoms[geom].K, THREADS = kGlsGeoms[geom].threads, MINB = kGlsGeoms[geom].minb;
  const long long fpb = (long long)K * THREADS;
  const long long nfb = (nf + fpb - 1) / fpb;
  (void)MINB;
  if (ctx->gls_occ[geom][w != nullptr] == 0) ctx->gls_occ[geom][w != nullptr] = strip_occupancy(geom, w != nullptr);
  const long long resident = (long long)ctx->sm_count * ctx->gls_occ[geom][w != nullptr];
  int nsplit = choose_nsplit((long long)B * nfb, nmax, resident, 0 /* splits share one plane: no memory cost */,
                             geom == 1 ? 96 : 256);
  if (ctx->gls_nsplit_override > 0) nsplit = ctx->gls_nsplit_override;  // tuning aid (env PDC_GLS_NSPLIT)
  long long free_blocks = 0;
  if (freqs_dev) {   // one frequency per thread, 128 per block, 8 blocks per SM
    free_blocks = (nf + GLS_FREE_THREADS - 1) / GLS_FREE_THREADS;
    nsplit = choose_nsplit(free_blocks, nmax, (long long)ctx->sm_count * 8, 0, 256);
  }
  const long long items = freqs_dev ? free_blocks * nsplit : (long long)B * nfb * nsplit;
  if (items > 0x7fffffffLL) { set_error("pdc_gls: problem too large for one call (%lld work items)", items); return PDC_EINVAL; }

  if (B > 65535) { set_error("pdc_gls_batch: at most 65535 curves per call"); return PDC_EINVAL; }

  // scratch is shared by all calls on this ctx: order this stream after the previous call
  ScratchScope scratch(ctx, st);
  PDC_TRY(scratch.acquire());

  // sub-cycle (FP64) bins: sized from the actual count, up to GLS_NLOW_CAP per curve (beyond that the surplus is FP32)
  long long nlowchunk = (nmax + GLS_LOW_CHUNK - 1) / GLS_LOW_CHUNK;
  if (nlowchunk > GLS_LOW_MAXCHUNKS) nlowchunk = GLS_LOW_MAXCHUNKS;
  long long low_cap = ((long long)1 << 22) / B;   // keeps the low plane below 200 MB for the largest batches
  if (low_cap > GLS_NLOW_CAP) low_cap = GLS_NLOW_CAP;
  if (low_cap < GLS_NLOW_MAX) low_cap = GLS_NLOW_MAX;
  // fixed-point scale of the plane of partial sums: |sum| <= n, one bit of slack
  int fix_bits = 61;
  while (fix_bits > 8 && ((long long)1 << (61 - fix_bits)) < nmax) --fix_bits;
  const double fix_scale = ldexp(1.0, fix_bits);

  // scratch
  PDC_TRY(ctx->gls_curves.reserve(sizeof(GlsCurve) * B));
  PDC_TRY(ctx->gls_rec1.reserve(sizeof(double2) * ntot));
  PDC_TRY(ctx->gls_rec2.reserve(sizeof(float4) * ntot));
  // ONE fixed-point plane of partial sums for all sample splits; the epilogue leaves it cleared, so it is zeroed only
  // when it is (re)allocated or when a previous call failed between the strip kernel and the epilogue
  {
    const void* before = ctx->gls_plane.p;
    const size_t cap_before = ctx->gls_plane.cap;
    PDC_TRY(ctx->gls_plane.reserve(sizeof(unsigned long long) * 6 * (size_t)nf_tot));
    if (ctx->gls_plane.p != before || ctx->gls_plane.cap != cap_before || ctx->gls_plane_dirty) {
      PDC_CUDA(cudaMemsetAsync(ctx->gls_plane.p, 0, ctx->gls_plane.cap, st));
      ctx->gls_plane_dirty = false;
    }
  }
  {   // the small plane of the sub-cycle bins: same clean-between-calls contract (and the same dirty flag)
    const void* before = ctx->gls_low.p;
    const size_t cap_before = ctx->gls_low.cap;
    PDC_TRY(ctx->gls_low.reserve(sizeof(unsigned long long) * 6 * (size_t)B * low_cap));
    if (ctx->gls_low.p != before || ctx->gls_low.cap != cap_before || ctx->gls_low_dirty) {
      PDC_CUDA(cudaMemsetAsync(ctx->gls_low.p, 0, ctx->gls_low.cap, st));
      ctx->gls_low_dirty = false;
    }
  }
  const int eblk = (int)((nf + 255) / 256);
  PDC_TRY(ctx->blockred.reserve((sizeof(double) + sizeof(long long)) * (size_t)eblk * B));
  // completion counters (stats blocks per curve, sample splits per frequency block, frequency blocks per curve):
  // zeroed when the buffer is (re)allocated, every kernel leaves them at zero again
  {
    const size_t need = sizeof(unsigned) * ((size_t)2 * B);
    const void* before = ctx->gls_cnt.p;
    const size_t cap_before = ctx->gls_cnt.cap;
    PDC_TRY(ctx->gls_cnt.reserve(need));
    if (ctx->gls_cnt.p != before || ctx->gls_cnt.cap != cap_before)
      PDC_CUDA(cudaMemsetAsync(ctx->gls_cnt.p, 0, ctx->gls_cnt.cap, st));
  }
  unsigned* cnt_stats = ctx->gls_cnt.as<unsigned>();
  unsigned* cnt_curve = cnt_stats + B;

  GlsCurve* dc = ctx->gls_curves.as<GlsCurve>();
  const long long off0 = offsets_host[0];
  GlsCurve single;
  memset(&single, 0, sizeof(single));
  if (B == 1) {
    // one curve: its host-known fields travel as a kernel argument (no metadata upload, no host-side fence)
    single.begin = 0;
    single.n = offsets_host[1] - offsets_host[0];
    single.fmin = fmin_host[0];
    single.df = df_host[0];
    single.psd_scale = psd_scale_host ? psd_scale_host[0] : 1.0;
  } else {
    // the pinned staging buffer is reused by every call: wait for the previous upload
    PDC_CUDA(cudaEventSynchronize(ctx->ev_fence));
    PDC_TRY(ctx->pin_meta.reserve(sizeof(GlsCurve) * B));
    GlsCurve* hc = ctx->pin_meta.as<GlsCurve>();
    for (int64_t b = 0; b < B; ++b) {
      memset(&hc[b], 0, sizeof(GlsCurve));
      hc[b].begin = offsets_host[b] - off0;
      hc[b].n = offsets_host[b + 1] - offsets_host[b];
      hc[b].fmin = fmin_host[b];
      hc[b].df = df_host[b];
      hc[b].psd_scale = psd_scale_host ? psd_scale_host[b] : 1.0;
    }
    PDC_CUDA(cudaMemcpyAsync(dc, hc, sizeof(GlsCurve) * B, cudaMemcpyHostToDevice, st));
    PDC_CUDA(cudaEventRecord(ctx->ev_fence, st));
  }

  const double* tt = t + off0;
  const double* yy = y + off0;
  const double* ww = w ? w + off0 : nullptr;

  {
    long long g = 1;
    if (B < 64) {
      g = (nmax + 8 * GLS_STATS_THREADS - 1) / (8 * GLS_STATS_THREADS);
      if (g > 32) g = 32;
      if (g < 1) g = 1;
    }
    PDC_TRY(ctx->gls_part.reserve(sizeof(GlsPart) * (size_t)g * B));
    dim3 grid((unsigned)g, (unsigned)B);
    gls_stats_kernel<<<grid, GLS_STATS_THREADS, 0, st>>>(tt, yy, ww, dc, ctx->gls_part.as<GlsPart>(), cnt_stats, single,
                                                        B == 1 ? 1 : 0, flags, (long long)j0, (long long)nf,
                                                        ctx->gls_three_term ? 1 : 0, (int)low_cap);
    PDC_CUDA(cudaGetLastError());
    ctx->launches++;
  }

  {
    long long bx = (nmax + 255) / 256;
    if (bx > 1024) bx = 1024;
    dim3 grid((unsigned)(bx + GLS_LOW_LANES * nlowchunk), (unsigned)B);
    gls_prep_kernel<<<grid, 256, 0, st>>>(tt, yy, ww, dc, ctx->gls_rec1.as<double2>(), ctx->gls_rec2.as<float4>(),
                                          ctx->gls_low.as<unsigned long long>(), (int)B, (int)low_cap, fix_scale,
                                          (long long)j0, (int)bx, (int)nlowchunk);
    PDC_CUDA(cudaGetLastError());
    ctx->launches++;
  }

  GlsMainArgs a;
  a.curves = dc;
  a.rec1 = ctx->gls_rec1.as<double2>();
  a.rec2 = ctx->gls_rec2.as<float4>();
  a.partial = ctx->gls_plane.as<unsigned long long>();
  a.nf = nf;
  a.nf_tot = nf_tot;
  a.j0 = j0;
  a.nfb = (int)nfb;
  a.nsplit = nsplit;
  a.fix_scale = (float)fix_scale;

  ctx->gls_plane_dirty = ctx->gls_low_dirty = true;   // until the epilogue that clears the planes has been enqueued
  const bool use_umma = !freqs_dev && gls_umma_eligible(ctx, B, nf, ntot, nmax, w != nullptr, df_host);
  if (!use_umma) PDC_TRY(ctx->main_begin(st));   // (the tensor-core path times its main kernel itself, after its operand pre-pass)
  ctx->last_gls_path = 0;
  if (freqs_dev) {
    GlsFreeArgs fa;
    fa.curves = dc;
    fa.rec1 = a.rec1;
    fa.rec2 = a.rec2;
    fa.freqs = freqs_dev;
    fa.partial = a.partial;
    fa.nf = nf;
    fa.nsplit = nsplit;
    fa.fix_scale = (float)fix_scale;
    if (w) gls_free_kernel<true><<<(unsigned)items, GLS_FREE_THREADS, 0, st>>>(fa);
    else gls_free_kernel<false><<<(unsigned)items, GLS_FREE_THREADS, 0, st>>>(fa);
    PDC_CUDA(cudaGetLastError());
    ctx->launches++;
  } else if (use_umma) {
    PDC_TRY(gls_umma_launch(ctx, dc, a.rec1, a.rec2, a.partial, B, nf, j0, nmax, w != nullptr, (float)fix_scale, st));
  } else {
    PDC_TRY(launch_strip(geom, ctx, a, w != nullptr, items, st));
  }
  if (!use_umma) PDC_TRY(ctx->main_end(st));

  {
    GlsEpiArgs e;
    e.curves = dc;
    e.partial = a.partial;
    e.lowplane = ctx->gls_low.as<unsigned long long>();
    e.low_cap = (int)low_cap;
    e.B = (int)B;
    e.inv_fix = 1.0 / fix_scale;
    e.flags = flags;
    e.nf = nf;
    e.nf_tot = nf_tot;
    e.j0 = j0;
    e.power_out = power_out;
    e.red_val = ctx->blockred.as<double>();
    e.red_idx = reinterpret_cast<long long*>(e.red_val + (size_t)eblk * B);
    e.curve_done = cnt_curve;
    e.arg_out = (long long*)argmax_out;
    e.max_out = max_out;
    if (fanout) e.fan = *fanout;
    else memset(&e.fan, 0, sizeof(e.fan));
    e.double_angle = use_umma ? 1 : 0;
    e.umma_status = use_umma ? ctx->umma_status_cur : nullptr;
    dim3 grid((unsigned)eblk, (unsigned)B);
    gls_epilogue_kernel<<<grid, 256, 0, st>>>(e);
    PDC_CUDA(cudaGetLastError());
    ctx->launches++;
    ctx->gls_plane_dirty = ctx->gls_low_dirty = false;
  }
  PDC_TRY(scratch.release());
  return PDC_OK;
}

}  // namespace pdc
