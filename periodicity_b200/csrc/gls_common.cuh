// Device code shared by the GLS kernels (gls.cu: one series per curve; glsm.cu: many series on
// common sampling times).
#pragma once

#include "pdc_common.cuh"

namespace pdc {

struct GlsCurve {
  long long begin, n;
  double fmin, df;
  double psd_scale;
  // filled on the device by gls_stats_kernel
  double tmin, tmax, wsum, ymean, yy, inv_rms;
  int low_begin, low_count;  // frequencies [low_begin, low_begin + low_count) of this call go through FP64
  // Three-term recurrence (see gls.cu): per-frequency-index phase offset in cycles that centres the
  // per-sample step angles 2 pi (df (t - tmin) + gamma) on a quarter turn; three_term == 0 when the step
  // angles span too much of the circle and the strip kernel has to use the plain rotation.
  double gamma;
  int three_term;
  int bad;       // some t, y or w of the curve is NaN / inf: every power of the curve is NaN (gls.cu)
};

// Order of (cos, sin, y', w') inside the float4 sample record.  The record is loaded with one
// LDS.128, which pins its four components to register banks 0..3 (the register file behaves as
// 4 banks, reg % 4, one read per bank per cycle -- tools/bank_model.py reproduces ncu's issue
// utilisation from the SASS with exactly that model); the order decides which components collide
// with the rotating (c, s) state.
#ifndef PDC_REC_CODE
#define PDC_REC_CODE 123  /* decimal digits = float4 slot of cos, sin, y', w' (0123 = x, y, z, w) */
#endif
__host__ __device__ constexpr int rec_slot(int which) {
  constexpr int order[4] = {(PDC_REC_CODE / 1000) % 10, (PDC_REC_CODE / 100) % 10, (PDC_REC_CODE / 10) % 10,
                            PDC_REC_CODE % 10};
  return order[which];
}
__host__ __device__ __forceinline__ float rec_get(const float4& r, int slot) {
  return slot == 0 ? r.x : (slot == 1 ? r.y : (slot == 2 ? r.z : r.w));
}
__host__ __device__ __forceinline__ void rec_set(float4& r, int slot, float v) {
  if (slot == 0) r.x = v;
  else if (slot == 1) r.y = v;
  else if (slot == 2) r.z = v;
  else r.w = v;
}
constexpr int REC_CR = 0, REC_SR = 1, REC_Y = 2, REC_W = 3;

constexpr int GLS_TILE = 1024;  // samples per shared-memory tile == FP32 flush interval

// The three-term recurrence c[k+1] = 2 cos(d) c[k] - c[k-1] amplifies rounding errors by 1 / |sin d|;
// it is used when every step angle d_i can be brought within +-GLS_TT_MAX_SPAN/2 turns of a quarter
// turn (|sin d| >= 0.48), i.e. when df * (tmax - tmin) <= GLS_TT_MAX_SPAN (n >= 3 samples per peak).
constexpr double GLS_TT_MAX_SPAN = 0.34;

// Frequencies with |f| * (tmax - tmin) < GLS_LOW_CYCLES see less than one cycle over the
// baseline: there CC - C^2 and SS - S^2 (spectral.py:125-127) cancel almost completely
// (a slow cosine is nearly degenerate with the floating mean) and amplify FP32 rounding
// by 1/var(cos) ~ 300x at f*T = 0.1.  Those few bins (at most GLS_NLOW_MAX per curve) are
// evaluated entirely in FP64 (gls.cu: the second block role of gls_prep_kernel; glsm.cu: glsm_lowfreq_kernel).
// Their number is ~ the grid's samples per peak `n` (spectral.py:88) when fmin is at its default: 5 by default, 100
// for GLS(n=100).  gls.cu sizes the range from the actual count up to GLS_NLOW_CAP per curve, glsm.cu keeps GLS_NLOW_MAX.
constexpr double GLS_LOW_CYCLES = 1.0;
constexpr int GLS_NLOW_MAX = 16;
constexpr int GLS_NLOW_CAP = 1024;
constexpr int GLS_LOW_CHUNK = 2048;   // samples per FP64 block (gls_prep_kernel's second role / glsm_lowfreq_kernel)
constexpr int GLS_LOW_MAXCHUNKS = 256;


#ifdef __CUDACC__

// Exact seed: phase = A + lK*b cycles (FP64), reduced mod 1 by a magic-number add
// whose low mantissa word is the fraction in units of 2^-32 cycle, then MUFU sin/cos.
__device__ __forceinline__ void gls_seed(double A, double b, double lKd, float& c, float& s) {
  const double ph = __fma_rn(lKd, b, A);
  const double v = __dadd_rn(ph, 1572864.0);  // 1.5 * 2^20: ulp(v) = 2^-32
  const int fx = __double2loint(v);           // two's-complement fraction, [-0.5, 0.5) cycle
  const float x = (float)fx * 1.4629180792671596e-9f;  // 2 pi / 2^32 -> radians in [-pi, pi)
  __sincosf(x, &s, &c);
}

__device__ __forceinline__ double np_sign(double x) {
  if (x != x) return x;
  return (double)((x > 0.0) - (x < 0.0));
}

// Low-frequency range of a call: indices j in [0, nf) with |fmin + (j0 + j) df| * T < GLS_LOW_CYCLES.
__device__ __forceinline__ void gls_low_range(double fmin, double df, long long j0, long long nf, double T,
                                              int& low_begin, int& low_count, int cap = GLS_NLOW_MAX) {
  low_begin = 0;
  low_count = 0;
  if (T > 0.0 && df != 0.0 && df == df) {
    const double flim = GLS_LOW_CYCLES / T;
    double lo = (-flim - fmin) / df - (double)j0, hi = (flim - fmin) / df - (double)j0;
    if (df < 0.0) { const double tmp = lo; lo = hi; hi = tmp; }   // a backward grid crosses the range the other way
    double ja = ceil(lo);
    double jb = floor(hi);
    if (ja < 0.0) ja = 0.0;
    if (jb > (double)(nf - 1)) jb = (double)(nf - 1);
    if (jb >= ja) {
      const double cnt = jb - ja + 1.0;
      low_begin = (int)ja;
      low_count = cnt > (double)cap ? cap : (int)cnt;
    }
  }
}

// FP64 epilogue for one frequency: spectral.py:113-132 literally.
//   sums = {sum w c, sum w s, sum w y c, sum w y s, sum w c^2, sum w c s} * (1 / inv_n)
// y was pre-scaled to unit weighted RMS, so YY == 1 (spectral.py:120,132).
// c2_direct: sums[4], sums[5] are sum w cos 2x, sum w sin 2x themselves (gls_umma.cu) instead of sum w cos^2 x, sum w cos x sin x.
__device__ __forceinline__ double gls_power_from_sums(const double* sums, double inv_n, unsigned flags,
                                                      double yy, double psd_scale, bool c2_direct = false) {
  const double C = sums[0] * inv_n, S = sums[1] * inv_n;
  const double Ch = sums[2] * inv_n, Sh = sums[3] * inv_n;
  // sum w cos(2x) = 2 sum w cos^2 x - 1,  sum w sin(2x) = 2 sum w sin x cos x   (sum w = 1)
  const double C2 = c2_direct ? sums[4] * inv_n : 2.0 * sums[4] * inv_n - 1.0;
  const double S2 = c2_direct ? sums[5] * inv_n : 2.0 * sums[5] * inv_n;
  const bool fit_mean = flags & PDC_GLS_FIT_MEAN;
  double tan2;
  if (fit_mean) tan2 = (S2 - 2.0 * S * C) / (C2 - (C * C - S * S));  // spectral.py:113
  else tan2 = S2 / C2;                                               // spectral.py:115
  const double hyp = sqrt(1.0 + tan2 * tan2);
  const double S2w = tan2 / hyp;
  const double C2w = 1.0 / hyp;
  const double Cw = sqrt(0.5) * sqrt(1.0 + C2w);
  const double Sw = sqrt(0.5) * np_sign(S2w) * sqrt(1.0 - C2w);
  const double YC = Ch * Cw + Sh * Sw;
  const double YS = Sh * Cw - Ch * Sw;
  double CC = 0.5 * (1.0 + C2 * C2w + S2 * S2w);
  double SS = 0.5 * (1.0 - C2 * C2w - S2 * S2w);
  if (fit_mean) {
    const double a1 = C * Cw + S * Sw, a2 = S * Cw - C * Sw;
    CC -= a1 * a1;
    SS -= a2 * a2;
  }
  double power = YC * YC / CC + YS * YS / SS;          // spectral.py:128
  if (flags & PDC_GLS_PSD) power *= yy * psd_scale;    // spectral.py:130
  else if (!(yy > 0.0)) power = nan("");              // spectral.py:132 with YY == 0
  return power;
}

#endif  // __CUDACC__

}  // namespace pdc
