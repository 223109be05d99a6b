// Phase -> fine-bin helpers shared by the phase-histogram kernels (pdm.cu: PDM / AoV; ce.cu: conditional entropy).
// Conventions of the reference: phi = (t / P) % 1 (phase.py:131), bin k = [k/m0, (k+1)/m0) against the very float64
// thresholds k/m0 the reference compares with (phase.py:138-140).
#pragma once

#include <type_traits>

#include "pdc_common.cuh"

namespace pdc {

// Fine-bin index of one sample for trial period P (rP = 1/P), plus an "ambiguity key":
// key < PDM_AMBIG means phi*m0 is within 2^-20 of an integer and the bin has to be decided
// against the reference's own thresholds (pdm_fix_bin).
constexpr unsigned PDM_AMBIG = 2u << 12;

__device__ __forceinline__ int pdm_bin(double tv, double P, double rP, double m0d, double& phi, unsigned& key) {
  // correctly rounded t / P: q0 = t * (1/P), exact FMA residual, one correction (phase.py:131)
  const double q0 = __dmul_rn(tv, rP);
  const double r = __fma_rn(-q0, P, tv);
  const double q1 = __fma_rn(r, rP, q0);
  phi = __dadd_rn(q1, -floor(q1));                 // exact; == np.remainder(q1, 1)
  // phi * m0 + 1.5 * 2^32 in ONE rounding: ulp = 2^-20, low word = rint(phi * m0 * 2^20)
  const double v = __fma_rn(phi, m0d, 6442450944.0);
  const int lo = __double2loint(v);
  key = (unsigned)(lo + 1) << 12;                  // fraction bits of u (+1 ulp), top-aligned
  return lo >> 20;                                 // floor(u) unless ambiguous
}

__device__ __forceinline__ int pdm_fix_bin(int k, double phi, const double* s_thr, int m0) {
  k = k < 0 ? 0 : (k > m0 - 1 ? m0 - 1 : k);
  if (phi < s_thr[k]) --k;
  else if (k < m0 - 1 && phi >= s_thr[k + 1]) ++k;
  return k;
}

// Fast path: |t / P| < 2^19 for every sample.  fma(t, 1/P, 1.5 * 2^20) leaves frac(t / P) in units of 2^-32
// turn in the low mantissa word (one DFMA, as in the GLS seed), a 32 x 32 -> 64 bit multiply by m0
// then gives the fine bin (high word) and the position inside the bin (low word).  The fixed-point
// phase is within 2^-32 + |t/P| 2^-52 of the reference's (t / P) % 1, so the bin can differ only
// if the phase lies within PDM_FAST_GUARD * 2^-32 of a bin edge: those samples (about m0 * 4e-9 of
// them) are re-binned by the exact path below.
constexpr double PDM_FAST_MAGIC = 1572864.0;   // 1.5 * 2^20
constexpr double PDM_FAST_LIMIT = 262144.0;    // |t / P| < 2^18 keeps the sum inside [2^20, 2^21)
constexpr unsigned PDM_FAST_GUARD = 4u;

// The guard is folded into the magic constant: the phase is shifted up by PDM_FAST_GUARD units of
// 2^-32 turn (exact: ulp of the sum is 2^-32), so `pos` = position inside the bin + guard, and
// pos < 2 guard  <=>  the unshifted phase is within guard of a bin edge (then the bin index may be off by
// one and the caller re-bins exactly); otherwise the shift cannot have carried into the bin index.
constexpr double PDM_FAST_MAGIC_G = PDM_FAST_MAGIC + PDM_FAST_GUARD * (1.0 / 4294967296.0);

// Shifted form, usable for time stamps far from zero (Julian dates: t ~ 2.45e6 d with periods of a day make
// |t / P| ~ 2^21, beyond the magic-number reduction).  With t0 = the smallest finite stamp,
//     (t / P) mod 1 = ((t - t0) / P + c) mod 1,   c = frac(t0 / P)  (one constant per trial period),
// so the kernel stages t' = t - t0 (exact in float64 for stamps of one series) and evaluates
//     fma(t', 1/P, magic),   magic = 1.5 * 2^20 + c + G * 2^-32 ,
// which needs only |t' / P| < 2^18.  The guard G (units of 2^-32 turn, folded into the magic number as above) must
// now cover the distance between this phase and the REFERENCE's own (t / P) % 1, whose quotient is rounded at the
// magnitude of t / P: |difference| <= 2 |t / P| 2^-53 + 2 * 2^-32, i.e. G = 4 + ceil(|t / P|_max 2^-20) units.  Samples
// inside the guard are re-binned by the exact path from the ORIGINAL stamp, so they land where the reference puts them.
constexpr double PDM_SHIFT_LIMIT = 1073741824.0;   // |t / P| < 2^30: G stays <= 1028 units (2.4e-7 turn)
__device__ __forceinline__ unsigned pdm_guard_units(double rP_abs, double t_absmax) {
  return 4u + (unsigned)ceil(rP_abs * t_absmax * (1.0 / 1048576.0));
}
__device__ __forceinline__ double pdm_fast_magic(double t0, double rP, unsigned guard_units) {
  const double p = __dmul_rn(t0, rP);
  const double e = __fma_rn(t0, rP, -p);
  const double c = __dadd_rn(__dadd_rn(p, -rint(p)), e);     // frac(t0 / P) in [-0.5, 0.5], exact product
  return __dadd_rn(__dadd_rn(PDM_FAST_MAGIC, c), (double)guard_units * (1.0 / 4294967296.0));
}
__device__ __forceinline__ unsigned pdm_bin_fast_m(double tv, double rP, double magic, unsigned m0u, unsigned& pos) {
  const unsigned u = (unsigned)__double2loint(__fma_rn(tv, rP, magic));
  const unsigned long long w = (unsigned long long)u * m0u;
  pos = (unsigned)w;
  return (unsigned)(w >> 32);
}

template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (N > 0) {
    static_for<N - 1>(f);
    f(std::integral_constant<int, N - 1>{});
  }
}

}  // namespace pdc
