// GLS of many series sampled at the SAME times (shared timestamps).
//
// Use cases: `GLS.bootstrap` with err=None (reference src/periodicity/spectral.py:140-152 --
// B resamples of the VALUES at fixed times; with uniform errors every replicate has the same
// weights), and survey sectors whose light curves share one time axis.  For S series on one
// time axis the rotation exp(2 pi i f_j t_i) and the window sums {C, S, CC, CS} (they depend on
// t and w only) are common to all series; only YC_s = sum w y_s cos and YS_s = sum w y_s sin
// differ.  A thread therefore owns K consecutive frequencies of R series at once:
//   per (sample, frequency): 2 FP32 instr for the step (three-term form, see gls.cu; 4 in the rotation
//   form) + 2R for the R series, and only in the first group of series 4 (5 weighted) for the window
//   sums = (2 + 2R)/R instructions per series-evaluation: 2.25 at R = 8 instead of 8 in gls_strip_kernel.
// Everything else -- exact FP64 seeding per strip, FP32 tile sums flushed to FP64 partials, FP64
// sub-cycle bins, FP64 epilogue (spectral.py:113-132), NaN-aware argmax -- is as in gls.cu.
//
// Kernels: glsm_stats_kernel (per series: mean, YY; shared: t range, sum w),
// glsm_records_kernel (shared rotation records + per-series scaled values, group-interleaved),
// glsm_lowfreq_kernel (FP64), glsm_strip_kernel (hot), glsm_epilogue_kernel, argext_final_kernel.
#include <type_traits>

#include "gls_common.cuh"

namespace pdc {

struct GlsmShared {  // one per call
  long long n;
  double fmin, df, psd_scale;
  double tmin, tmax, wsum;
  int low_begin, low_count;
  double gamma;    // per-frequency-index phase origin of the three-term step (see gls.cu / GlsCurve)
  int three_term, pad_;
};

struct GlsmSeries {  // one per series
  double ymean, yy, inv_rms;
};

constexpr int GLSM_TILE = 512;

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
glsm_stats_kernel(const double* __restrict__ t, const double* __restrict__ Y, const double* __restrict__ w,
                  GlsmShared* sh, GlsmSeries* series, unsigned flags, long long j0, long long nf,
                  int allow_three_term) {
  __shared__ double scratch[33];
  const long long n = sh->n;
  const double* y = Y + (long long)blockIdx.x * n;
  double tmin = INFINITY, tneg = INFINITY, sw = 0.0, swy = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const double ti = t[i], wi = w ? w[i] : 1.0;
    tmin = fmin(tmin, ti);
    tneg = fmin(tneg, -ti);
    sw += wi;
    swy = fma(wi, y[i], swy);
  }
  tmin = block_min(tmin, scratch);
  const double tmax = -block_min(tneg, scratch);
  sw = block_sum(sw, scratch);
  swy = block_sum(swy, scratch);
  const double ymean = (flags & PDC_GLS_FIT_MEAN) ? swy / sw : 0.0;  // spectral.py:104-108
  double syy = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = y[i] - ymean, wi = w ? w[i] : 1.0;
    syy = fma(wi * d, d, syy);
  }
  syy = block_sum(syy, scratch);
  if (threadIdx.x == 0) {
    const double yy = syy / sw;  // spectral.py:120
    series[blockIdx.x].ymean = ymean;
    series[blockIdx.x].yy = yy;
    series[blockIdx.x].inv_rms = yy > 0.0 ? rsqrt(yy) : 0.0;
    if (blockIdx.x == 0) {
      sh->tmin = tmin;
      sh->tmax = tmax;
      sh->wsum = sw;
      int lb, lc;
      gls_low_range(sh->fmin, sh->df, j0, nf, tmax - tmin, lb, lc);
      sh->low_begin = lb;
      sh->low_count = lc;
      const double span = sh->df * (tmax - tmin);
      const bool tt = allow_three_term && sh->df > 0.0 && span <= GLS_TT_MAX_SPAN;
      sh->three_term = tt;
      sh->gamma = tt ? 0.25 - 0.5 * span : 0.0;
    }
  }
}

// rec1[i] = (t - tmin, frac(df (t - tmin))); rot[i] = (cos, sin of the per-index rotation, w', 0);
// yrec[(g*n + i)*R + r] = w'_i * y'_{g*R+r, i}  (series padded with zeros up to a multiple of R)
template <int R>
__global__ void __launch_bounds__(256)
glsm_records_kernel(const double* __restrict__ t, const double* __restrict__ Y, const double* __restrict__ w,
                    const GlsmShared* __restrict__ shp, const GlsmSeries* __restrict__ series, long long S,
                    double2* __restrict__ rec1, float4* __restrict__ rot, float* __restrict__ yrec) {
  const GlsmShared sh = *shp;
  const long long n = sh.n;
  const int g = blockIdx.y;
  const double wscale = (double)n / sh.wsum;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double wn = w ? w[i] * wscale : 1.0;
    if (g == 0) {
      const double tt = t[i] - sh.tmin;
      const double b = frac_of_product(sh.df, tt) + sh.gamma;
      double sb, cb;
      sincospi(2.0 * b, &sb, &cb);
      rec1[i] = make_double2(tt, b);
      rot[i] = make_float4((float)cb, (float)sb, (float)wn, 0.f);
    }
    float* out = yrec + ((long long)g * n + i) * R;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long s = (long long)g * R + r;
      float v = 0.f;
      if (s < S) v = (float)(wn * ((Y[s * n + i] - series[s].ymean) * series[s].inv_rms));
      out[r] = v;
    }
  }
}

// FP64 sums for the sub-cycle bins: grid (chunk, low bin, group); sincos shared by the R series.
// lowwin[chunk][4][NLOW] (window sums, written by group 0) and lowys[chunk][2][S_pad][NLOW].
template <int R>
__global__ void __launch_bounds__(256)
glsm_lowfreq_kernel(const double* __restrict__ t, const double* __restrict__ Y, const double* __restrict__ w,
                    const GlsmShared* __restrict__ shp, const GlsmSeries* __restrict__ series, long long S,
                    long long S_pad, double* __restrict__ lowwin, double* __restrict__ lowys, long long j0) {
  __shared__ double scratch[33];
  const GlsmShared sh = *shp;
  const int chunk = blockIdx.x, nchunk = gridDim.x, slot = blockIdx.y, g = blockIdx.z;
  if (slot >= sh.low_count) return;
  const long long n = sh.n;
  const double f = sh.fmin + (double)(j0 + sh.low_begin + slot) * sh.df;
  const long long per = (n + nchunk - 1) / nchunk;
  const long long sb = (long long)chunk * per;
  const long long se = sb + per < n ? sb + per : n;
  const double winv = 1.0 / sh.wsum;
  double aw[4] = {0, 0, 0, 0}, ayc[R], ays[R], ym[R], ir[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const long long s = (long long)g * R + r;
    ayc[r] = ays[r] = 0.0;
    ym[r] = s < S ? series[s].ymean : 0.0;
    ir[r] = s < S ? series[s].inv_rms : 0.0;
  }
  for (long long i = sb + threadIdx.x; i < se; i += blockDim.x) {
    const double ph = frac_of_product(f, t[i] - sh.tmin);
    double sn, cs;
    sincospi(2.0 * ph, &sn, &cs);
    const double wi = (w ? w[i] : 1.0) * winv;
    const double wc = wi * cs, ws = wi * sn;
    aw[0] += wc;
    aw[1] += ws;
    aw[2] = fma(wc, cs, aw[2]);
    aw[3] = fma(wc, sn, aw[3]);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long s = (long long)g * R + r;
      const double yv = s < S ? (Y[s * n + i] - ym[r]) * ir[r] : 0.0;
      ayc[r] = fma(wc, yv, ayc[r]);
      ays[r] = fma(ws, yv, ays[r]);
    }
  }
  if (g == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double tot = block_sum(aw[q], scratch);
      if (threadIdx.x == 0) lowwin[((long long)chunk * 4 + q) * GLS_NLOW_MAX + slot] = tot;
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const long long s = (long long)g * R + r;
    const double tc = block_sum(ayc[r], scratch);
    const double ts = block_sum(ays[r], scratch);
    if (threadIdx.x == 0) {
      lowys[(((long long)chunk * 2 + 0) * S_pad + s) * GLS_NLOW_MAX + slot] = tc;
      lowys[(((long long)chunk * 2 + 1) * S_pad + s) * GLS_NLOW_MAX + slot] = ts;
    }
  }
}

// ---------------------------------------------------------------------------
struct GlsmArgs {
  const GlsmShared* sh;
  const double2* rec1;
  const float4* rot;
  const float* yrec;
  double* win;   // [nsplit][4][nf]
  double* ys;    // [nsplit][2][S_pad][nf]
  long long nf, j0, S_pad;
  int nfb, nsplit, ngroups;
};

template <int K, int R, int THREADS, bool WEIGHTED>
__global__ void __launch_bounds__(THREADS)
glsm_strip_kernel(const GlsmArgs a) {
  __shared__ __align__(16) double2 s_ab[GLSM_TILE + 2];
  __shared__ __align__(16) float4 s_rot[GLSM_TILE + 2];
  __shared__ __align__(16) float s_y[(GLSM_TILE + 2) * R];

  const int item = blockIdx.x;
  const int split = item % a.nsplit;
  const int rest = item / a.nsplit;
  const int fb = rest % a.nfb;
  const int g = rest / a.nfb;

  const long long n = a.sh->n;
  const double fmin = a.sh->fmin, df = a.sh->df;
  const long long per = (n + a.nsplit - 1) / a.nsplit;
  const long long sb = (long long)split * per;
  const long long se = sb + per < n ? sb + per : n;

  const long long jB = (long long)fb * (THREADS * K);
  const double fB = fmin + (double)(a.j0 + jB) * df;
  const int lK = threadIdx.x * K;
  const double lKd = (double)lK;
  const long long jrem = a.nf - (jB + lK);
  const bool three_term = a.sh->three_term != 0;  // block-uniform
  double gB = (double)(a.j0 + jB) * a.sh->gamma;   // phase origin of the block's first frequency (turns)
  gB -= floor(gB);

  double* pwin = a.win + (long long)split * 4 * a.nf + jB + lK;
  double* pys = a.ys + ((long long)split * 2 * a.S_pad + (long long)g * R) * a.nf + jB + lK;
  const long long ys_stat = a.S_pad * a.nf;  // distance between the YC and YS planes

  if (threadIdx.x < 2) {
    s_ab[GLSM_TILE + threadIdx.x] = make_double2(0.0, 0.0);
    s_rot[GLSM_TILE + threadIdx.x] = make_float4(1.f, 0.f, 0.f, 0.f);
  }
  for (int k = threadIdx.x; k < 2 * R; k += THREADS) s_y[GLSM_TILE * R + k] = 0.f;

  bool first = true;
  long long tile0 = sb;
  do {
    long long left = se - tile0;
    const int cnt = left <= 0 ? 0 : (left < GLSM_TILE ? (int)left : GLSM_TILE);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += THREADS) {
      const double2 r1 = a.rec1[tile0 + i];
      s_ab[i] = make_double2(frac_of_product(fB, r1.x) + gB, r1.y);
      s_rot[i] = a.rot[tile0 + i];
    }
    {
      const float4* src = reinterpret_cast<const float4*>(a.yrec + ((long long)g * n + tile0) * R);
      float4* dst = reinterpret_cast<float4*>(s_y);
      for (int i = threadIdx.x; i < cnt * R / 4; i += THREADS) dst[i] = src[i];
    }
    __syncthreads();

    float aC[K], aS[K], aCC[K], aCS[K], aYC[R][K], aYS[R][K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      aC[k] = aS[k] = aCC[k] = aCS[k] = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) aYC[r][k] = aYS[r][k] = 0.f;
    }

    // One sample: K accumulations and K-1 steps along the frequency axis (rotation or three-term
    // form, as in gls_strip_kernel).  The window sums {C, S, CC, CS} are the same in every group of
    // series: only group 0 (WIN) spends instructions on them.
    auto strip = [&](auto win_tag, auto tt_tag, float c, float s, const float4 rt, const float* yv) {
      constexpr bool WIN = decltype(win_tag)::value, TT = decltype(tt_tag)::value;
      const float cr = rt.x, sr = rt.y;
      const float tc = cr + cr;
      float cp = 0.f, sp = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (WIN) {
          if (WEIGHTED) {
            const float wc = rt.z * c;
            aC[k] += wc;
            aS[k] = fmaf(rt.z, s, aS[k]);
            aCC[k] = fmaf(wc, c, aCC[k]);
            aCS[k] = fmaf(wc, s, aCS[k]);
          } else {
            aC[k] += c;
            aS[k] += s;
            aCC[k] = fmaf(c, c, aCC[k]);
            aCS[k] = fmaf(c, s, aCS[k]);
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          aYC[r][k] = fmaf(yv[r], c, aYC[r][k]);
          aYS[r][k] = fmaf(yv[r], s, aYS[r][k]);
        }
        if (k + 1 < K) {
          float c2, s2;
          if (TT && k >= 1) {  // c[k+1] = 2 cos(d) c[k] - c[k-1]
            c2 = fmaf(tc, c, -cp);
            s2 = fmaf(tc, s, -sp);
          } else {
            c2 = fmaf(c, cr, -(s * sr));
            s2 = fmaf(s, cr, c * sr);
          }
          cp = c;
          sp = s;
          c = c2;
          s = s2;
        }
      }
    };

    auto run_tile = [&](auto win_tag, auto tt_tag) {
      float c0, s0, c1, s1;
      {
        const double2 ab = s_ab[0];
        gls_seed(ab.x, ab.y, lKd, c0, s0);
      }
      for (int i = 0; i < cnt; ++i) {
        {  // exact seed of the next sample overlaps this sample's strip
          const double2 ab = s_ab[i + 1];
          gls_seed(ab.x, ab.y, lKd, c1, s1);
        }
        float yv[R];
        const float4* yp = reinterpret_cast<const float4*>(s_y + i * R);
#pragma unroll
        for (int q = 0; q < R / 4; ++q) {
          const float4 v = yp[q];
          yv[4 * q] = v.x; yv[4 * q + 1] = v.y; yv[4 * q + 2] = v.z; yv[4 * q + 3] = v.w;
        }
        strip(win_tag, tt_tag, c0, s0, s_rot[i], yv);
        c0 = c1;
        s0 = s1;
      }
    };
    if (cnt > 0) {
      if (g == 0) {
        if (three_term) run_tile(std::true_type{}, std::true_type{});
        else run_tile(std::true_type{}, std::false_type{});
      } else {
        if (three_term) run_tile(std::false_type{}, std::true_type{});
        else run_tile(std::false_type{}, std::false_type{});
      }
    }

    // flush: the window sums are identical in every group; group 0 publishes them
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (k < jrem) {
        if (g == 0) {
          double* p = pwin + k;
          if (first) {
            p[0] = (double)aC[k];
            p[a.nf] = (double)aS[k];
            p[2 * a.nf] = (double)aCC[k];
            p[3 * a.nf] = (double)aCS[k];
          } else {
            atomicAdd(p, (double)aC[k]);
            atomicAdd(p + a.nf, (double)aS[k]);
            atomicAdd(p + 2 * a.nf, (double)aCC[k]);
            atomicAdd(p + 3 * a.nf, (double)aCS[k]);
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          double* p = pys + (long long)r * a.nf + k;
          if (first) {
            p[0] = (double)aYC[r][k];
            p[ys_stat] = (double)aYS[r][k];
          } else {
            atomicAdd(p, (double)aYC[r][k]);
            atomicAdd(p + ys_stat, (double)aYS[r][k]);
          }
        }
      }
    }
    first = false;
    tile0 += GLSM_TILE;
  } while (tile0 < se);
}

__global__ void __launch_bounds__(256)
glsm_epilogue_kernel(const GlsmShared* __restrict__ shp, const GlsmSeries* __restrict__ series,
                     const double* __restrict__ win, const double* __restrict__ ys,
                     const double* __restrict__ lowwin, const double* __restrict__ lowys, int nlowchunk,
                     int nsplit, long long nf, long long S_pad, unsigned flags,
                     double* __restrict__ power_out, double* __restrict__ red_val, long long* __restrict__ red_idx) {
  __shared__ double sv[32];
  __shared__ long long si[32];
  const GlsmShared sh = *shp;
  const long long s = blockIdx.y;
  const GlsmSeries se = series[s];
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double power = 0.0;
  long long idx = -1;
  if (j < nf) {
    double sums[6] = {0, 0, 0, 0, 0, 0};
    double inv_n;
    if (j >= sh.low_begin && j < sh.low_begin + sh.low_count) {
      const int slot = (int)(j - sh.low_begin);
      for (int c = 0; c < nlowchunk; ++c) {
        sums[0] += lowwin[((long long)c * 4 + 0) * GLS_NLOW_MAX + slot];
        sums[1] += lowwin[((long long)c * 4 + 1) * GLS_NLOW_MAX + slot];
        sums[4] += lowwin[((long long)c * 4 + 2) * GLS_NLOW_MAX + slot];
        sums[5] += lowwin[((long long)c * 4 + 3) * GLS_NLOW_MAX + slot];
        sums[2] += lowys[(((long long)c * 2 + 0) * S_pad + s) * GLS_NLOW_MAX + slot];
        sums[3] += lowys[(((long long)c * 2 + 1) * S_pad + s) * GLS_NLOW_MAX + slot];
      }
      inv_n = 1.0;
    } else {
      for (int p = 0; p < nsplit; ++p) {
        sums[0] += win[((long long)p * 4 + 0) * nf + j];
        sums[1] += win[((long long)p * 4 + 1) * nf + j];
        sums[4] += win[((long long)p * 4 + 2) * nf + j];
        sums[5] += win[((long long)p * 4 + 3) * nf + j];
        sums[2] += ys[(((long long)p * 2 + 0) * S_pad + s) * nf + j];
        sums[3] += ys[(((long long)p * 2 + 1) * S_pad + s) * nf + j];
      }
      inv_n = 1.0 / (double)sh.n;
    }
    power = gls_power_from_sums(sums, inv_n, flags, se.yy, sh.psd_scale);
    if (power_out) power_out[s * nf + j] = power;
    idx = j;
  }
  block_argext<+1>(power, idx, sv, si);
  if (threadIdx.x == 0) {
    red_val[s * gridDim.x + blockIdx.x] = power;
    red_idx[s * gridDim.x + blockIdx.x] = idx;
  }
}

// ---------------------------------------------------------------------------
constexpr int GLSM_K = 8, GLSM_R = 8, GLSM_THREADS = 128;

int glsm_run(pdc_ctx* ctx, const double* t, const double* Y, const double* w, int64_t n, int64_t S,
             double fmin, double df, int64_t j0, int64_t nf, unsigned flags, double psd_scale,
             double* power_out, int64_t* argmax_out, double* max_out, cudaStream_t st) {
  if (n < 1 || S < 1 || nf < 1) { set_error("pdc_gls_multi: need n, S, nf >= 1"); return PDC_EINVAL; }
  if (S > 65535) { set_error("pdc_gls_multi: at most 65535 series per call"); return PDC_EINVAL; }
  if (!(df == df) || !(fmin == fmin)) { set_error("pdc_gls_multi: fmin/df is NaN"); return PDC_EINVAL; }
  constexpr int K = GLSM_K, R = GLSM_R, THREADS = GLSM_THREADS;
  const long long G = (S + R - 1) / R, S_pad = G * R;
  const long long fpb = (long long)K * THREADS;
  const long long nfb = (nf + fpb - 1) / fpb;

  if (ctx->glsm_occ[w != nullptr] == 0) {
    int nb = 0;
    cudaError_t e = w ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, glsm_strip_kernel<K, R, THREADS, true>, THREADS, 0)
                      : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, glsm_strip_kernel<K, R, THREADS, false>, THREADS, 0);
    if (e != cudaSuccess) { cudaGetLastError(); nb = 1; }
    ctx->glsm_occ[w != nullptr] = nb > 0 ? nb : 1;
  }
  const long long resident = (long long)ctx->sm_count * ctx->glsm_occ[w != nullptr];
  // sample split (same wave model as gls.cu); scratch per split: window 4 planes + 2 planes per series
  int nsplit = 1;
  {
    const long long base = G * nfb;
    long long cap = n / 256;
    if (cap < 1) cap = 1;
    if (cap > 1024) cap = 1024;
    const long long plane = (long long)sizeof(double) * (4 + 2 * S_pad) * nf;
    const long long mem_cap = ((long long)2 << 30) / (plane > 0 ? plane : 1);
    if (cap > mem_cap) cap = mem_cap < 1 ? 1 : mem_cap;
    if (base < 24 * resident) {
      double best = 1e300;
      for (long long s = 1; s <= cap; ++s) {
        const long long items = base * s;
        const long long waves = (items + resident - 1) / resident;
        const double cost = (double)waves * ((double)((n + s - 1) / s) + 64.0) * (1.0 + 0.04 / (double)waves);
        if (cost < best * 0.999) { best = cost; nsplit = (int)s; }
        if (items > 64 * resident) break;
      }
    }
  }
  const long long items = G * nfb * nsplit;
  if (items > 0x7fffffffLL) { set_error("pdc_gls_multi: problem too large for one call"); return PDC_EINVAL; }

  long long nlowchunk = (n + GLS_LOW_CHUNK - 1) / GLS_LOW_CHUNK;
  if (nlowchunk > GLS_LOW_MAXCHUNKS) nlowchunk = GLS_LOW_MAXCHUNKS;
  const int eblk = (int)((nf + 255) / 256);

  ScratchScope scratch(ctx, st);
  PDC_TRY(scratch.acquire());
  PDC_CUDA(cudaEventSynchronize(ctx->ev_fence));
  PDC_TRY(ctx->gls_curves.reserve(sizeof(GlsmShared) + sizeof(GlsmSeries) * S));
  PDC_TRY(ctx->pin_meta.reserve(sizeof(GlsmShared)));
  PDC_TRY(ctx->gls_rec1.reserve(sizeof(double2) * n));
  PDC_TRY(ctx->gls_rec2.reserve(sizeof(float4) * n));
  PDC_TRY(ctx->glsm_y.reserve(sizeof(float) * (size_t)S_pad * n));
  PDC_TRY(ctx->partial.reserve(sizeof(double) * (size_t)nsplit * (4 + 2 * S_pad) * nf));
  PDC_TRY(ctx->gls_low.reserve(sizeof(double) * (size_t)nlowchunk * (4 + 2 * S_pad) * GLS_NLOW_MAX));
  ctx->gls_low_dirty = true;   // gls.cu keeps this buffer as an all-zero fixed-point plane between its calls: it must re-clear it
  PDC_TRY(ctx->blockred.reserve((sizeof(double) + sizeof(long long)) * (size_t)eblk * S));

  GlsmShared* hs = ctx->pin_meta.as<GlsmShared>();
  hs->n = n; hs->fmin = fmin; hs->df = df; hs->psd_scale = psd_scale;
  hs->tmin = hs->tmax = hs->wsum = 0.0; hs->low_begin = hs->low_count = 0;
  hs->gamma = 0.0; hs->three_term = hs->pad_ = 0;
  GlsmShared* dsh = ctx->gls_curves.as<GlsmShared>();
  GlsmSeries* dser = reinterpret_cast<GlsmSeries*>(dsh + 1);
  PDC_CUDA(cudaMemcpyAsync(dsh, hs, sizeof(GlsmShared), cudaMemcpyHostToDevice, st));
  PDC_CUDA(cudaEventRecord(ctx->ev_fence, st));

  glsm_stats_kernel<<<(unsigned)S, 1024, 0, st>>>(t, Y, w, dsh, dser, flags, (long long)j0, (long long)nf,
                                                  ctx->gls_three_term ? 1 : 0);
  PDC_CUDA(cudaGetLastError());
  ctx->launches++;
  {
    long long bx = (n + 255) / 256;
    if (bx > 512) bx = 512;
    dim3 grid((unsigned)bx, (unsigned)G);
    glsm_records_kernel<R><<<grid, 256, 0, st>>>(t, Y, w, dsh, dser, S, ctx->gls_rec1.as<double2>(),
                                                ctx->gls_rec2.as<float4>(), ctx->glsm_y.as<float>());
    PDC_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  double* lowwin = ctx->gls_low.as<double>();
  double* lowys = lowwin + (size_t)nlowchunk * 4 * GLS_NLOW_MAX;
  {
    dim3 grid((unsigned)nlowchunk, GLS_NLOW_MAX, (unsigned)G);
    glsm_lowfreq_kernel<R><<<grid, 256, 0, st>>>(t, Y, w, dsh, dser, S, S_pad, lowwin, lowys, (long long)j0);
    PDC_CUDA(cudaGetLastError());
    ctx->launches++;
  }

  GlsmArgs a;
  a.sh = dsh;
  a.rec1 = ctx->gls_rec1.as<double2>();
  a.rot = ctx->gls_rec2.as<float4>();
  a.yrec = ctx->glsm_y.as<float>();
  a.win = ctx->partial.as<double>();
  a.ys = a.win + (size_t)nsplit * 4 * nf;
  a.nf = nf; a.j0 = j0; a.S_pad = S_pad;
  a.nfb = (int)nfb; a.nsplit = nsplit; a.ngroups = (int)G;

  PDC_TRY(ctx->main_begin(st));
  if (w) glsm_strip_kernel<K, R, THREADS, true><<<(unsigned)items, THREADS, 0, st>>>(a);
  else glsm_strip_kernel<K, R, THREADS, false><<<(unsigned)items, THREADS, 0, st>>>(a);
  PDC_CUDA(cudaGetLastError());
  ctx->launches++;
  PDC_TRY(ctx->main_end(st));

  double* red_val = ctx->blockred.as<double>();
  long long* red_idx = reinterpret_cast<long long*>(red_val + (size_t)eblk * S);
  {
    dim3 grid((unsigned)eblk, (unsigned)S);
    glsm_epilogue_kernel<<<grid, 256, 0, st>>>(dsh, dser, a.win, a.ys, lowwin, lowys, (int)nlowchunk, nsplit, nf, S_pad,
                                               flags, power_out, red_val, red_idx);
    PDC_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  if (argmax_out || max_out) {
    argext_final_kernel<+1><<<(unsigned)S, 256, 0, st>>>(red_val, red_idx, eblk, (long long*)argmax_out, max_out);
    PDC_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  PDC_TRY(scratch.release());
  return PDC_OK;
}

}  // namespace pdc
