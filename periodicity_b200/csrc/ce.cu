// Conditional-entropy periodogram (Graham et al. 2013) for a grid of trial periods.
//
// The reference only lists the method as a TODO (src/periodicity/phase.py:13, "conditional entropy"); there is no
// reference code, so PARITY IS UNPINNED BY THE REFERENCE (oracle: oracle/ce_numpy.py, pinned to np.histogram2d).
// Conventions follow the reference's PDM where they overlap: phase phi = (t / P) % 1 (phase.py:131), phase bin k
// selected by the float64 thresholds k / nphi (phase.py:138-140 with nc = 1).  Magnitudes are scaled to [0, 1] with
// the sample minimum and maximum and cut into nm equal bins (the maximum goes to the last bin):
//     mbin_i = min(int(nm * (x_i - min) / (max - min)), nm - 1)           -- period independent
//     H_c(P) = sum_jk p(phi_j, m_k) ln( p(phi_j) / p(phi_j, m_k) ),  p = cell occupation / N, over the occupied cells;
// the best period MINIMISES it.
//
// It needs COUNTS only, in nphi * nm cells: `cell = phase_bin * nm + mbin`.  Mapping = the packed loop of
// pdm_hist_kernel (pdm.cu) without its second level: every trial period owns a private column of 32-bit count words in
// shared memory (consecutive columns -> consecutive words, conflict free), a sample is ONE native shared-memory
// integer atomic without return value (ATOMS.ADD, fire-and-forget) and a 32-bit count cannot overflow, so there is no
// feed, no flush window: the columns go to global memory once, when the block has seen its share of the samples.
// Phase -> bin: the fixed-point fast path of phase_common.cuh (one DFMA + one 32 x 32 -> 64 bit multiply), samples
// within 4 * 2^-32 of a bin edge re-binned exactly and the increment moved; huge |t / P| or non-finite samples take
// the exact FP64 path for every sample.
//
// Kernels: ce_stats_kernel (min / max of x, max |t|, non-finite flag; multi-block, last block finalises),
// ce_hist_kernel (hot; at the end a thread adds its columns to ONE plane of 32-bit counts shared by all sample splits,
// RED.ADD.32 -- integer, hence order independent), ce_epilogue_kernel (FP64 entropy per trial period, clears the plane
// for the next call, arg-min; the last block finalises).
#include <cstring>

#include "pdc_common.cuh"
#include "phase_common.cuh"

namespace pdc {

struct CeMeta {
  double xmin, xmax;
  double t_absmax;
  double t0, t_span;   // smallest finite stamp, span of the finite stamps (the fixed-point phase works on t - t0)
  int bad, pad_;
};

struct CePart {
  double xneg, xmax, tabs, tneg, tmax;
  int bad, pad_;
};
constexpr int CE_STATS_THREADS = 256;
constexpr int CE_STATS_MAXBLK = 128;
constexpr int CE_TILE = 1024;
constexpr int CE_U = 16;            // samples per trip: CE_U x PPT independent DFMA -> IMAD.WIDE -> IMAD -> ATOMS chains
constexpr int CE_TILE_PAD = CE_U;   // the prefetch reads one trip past the tile

__global__ void __launch_bounds__(CE_STATS_THREADS)
ce_stats_kernel(const double* __restrict__ t, const double* __restrict__ x, long long n, CePart* part, unsigned* done,
                CeMeta* meta) {
  __shared__ double scratch[32 * 6];
  __shared__ int s_last;
  const int G = gridDim.x;
  double xneg = -INFINITY, xmax = -INFINITY, tabs = 0.0, tneg = -INFINITY, tmax = -INFINITY;
  int bad = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)G * blockDim.x) {
    const double xi = x[i], ti = t[i];
    xneg = fmax(xneg, -xi);     // fmax / fmin ignore NaN, as np.nanmin / np.nanmax do (oracle: magnitude_bins)
    xmax = fmax(xmax, xi);
    bad |= !isfinite(xi) || !isfinite(ti);
    if (isfinite(ti)) {
      tabs = fmax(tabs, fabs(ti));
      tneg = fmax(tneg, -ti);
      tmax = fmax(tmax, ti);
    }
  }
  bad = __syncthreads_or(bad);
  {
    double none[1] = {0.0}, maxs[5] = {xneg, xmax, tabs, tneg, tmax};
    block_reduce_many<1, 5>(none, maxs, scratch);
    xneg = maxs[0]; xmax = maxs[1]; tabs = maxs[2]; tneg = maxs[3]; tmax = maxs[4];
  }
  if (threadIdx.x == 0) {
    part[blockIdx.x].xneg = xneg;
    part[blockIdx.x].xmax = xmax;
    part[blockIdx.x].tabs = tabs;
    part[blockIdx.x].tneg = tneg;
    part[blockIdx.x].tmax = tmax;
    part[blockIdx.x].bad = bad;
    __threadfence();
    s_last = atomicAdd(done, 1u) == (unsigned)(G - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const bool has = (int)threadIdx.x < G;
  const CePart* p = part + threadIdx.x;
  {
    double none[1] = {0.0};
    double maxs[5] = {has ? __ldcg(&p->xneg) : -INFINITY, has ? __ldcg(&p->xmax) : -INFINITY, has ? __ldcg(&p->tabs) : 0.0,
                      has ? __ldcg(&p->tneg) : -INFINITY, has ? __ldcg(&p->tmax) : -INFINITY};
    block_reduce_many<1, 5>(none, maxs, scratch);
    xneg = maxs[0]; xmax = maxs[1]; tabs = maxs[2]; tneg = maxs[3]; tmax = maxs[4];
  }
  bad = __syncthreads_or(has ? __ldcg(&p->bad) : 0);
  if (threadIdx.x == 0) {
    meta->xmin = -xneg;
    meta->xmax = xmax;
    meta->t_absmax = tabs;
    meta->t0 = tmax >= -tneg ? -tneg : 0.0;
    meta->t_span = tmax >= -tneg ? tmax + tneg : 0.0;
    meta->bad = bad;
    *done = 0u;
  }
}

struct CeArgs {
  const double* t;
  const double* x;
  const double* periods;
  const CeMeta* meta;
  unsigned* plane;      // [cells][np] counts, all sample splits add into it
  long long n, np;
  int nphi, nm, cells, nsplit;
  int counts_only;      // 1: phase bins only (nm == 1, x is not read): the Gregory-Loredo passes
};

// numpy's magnitude bin: scaled = (x - lo) / (hi - lo); min(int(scaled * nm), nm - 1)  (oracle/ce_numpy.py)
__device__ __forceinline__ int ce_mbin(double xv, double lo, double range, int nm) {
  const double scaled = __ddiv_rn(__dadd_rn(xv, -lo), range);
  const double v = __dmul_rn(scaled, (double)nm);
  if (!(v == v)) return -1;                 // NaN value (or a constant signal: 0 / 0): in no cell
  int k = (int)v;                           // truncation towards zero, as astype(int64)
  return k < 0 ? 0 : (k > nm - 1 ? nm - 1 : k);
}

struct CeEpiArgs {
  unsigned* plane;      // [cells][np]; read and cleared
  const double* periods;
  long long np;
  int nphi, nm;
  double* h_out;        // [np]
  double* red_val;      // [gridDim.x]
  long long* red_idx;
  unsigned* call_done;  // [1] self-resetting
  long long* arg_out;
  double* best_out;
};

__global__ void __launch_bounds__(256)
ce_epilogue_kernel(const CeEpiArgs a) {
  __shared__ double sv[32];
  __shared__ long long si[32];
  __shared__ int s_last;
  const long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long np = a.np;
  double h = 0.0;
  long long idx = -1;
  if (pi < np) {
    unsigned* base = a.plane + pi;
    // H = (1 / N) sum_j [ r_j ln r_j - sum_k c_jk ln c_jk ],  r_j = sum_k c_jk,  N = sum_j r_j
    double acc = 0.0, ntot = 0.0;
    for (int j = 0; j < a.nphi; ++j) {
      double r = 0.0, cl = 0.0;
      for (int k = 0; k < a.nm; ++k) {
        unsigned* cell = base + (long long)(j * a.nm + k) * np;
        const double c = (double)__ldcg(cell);
        *cell = 0u;                                          // the plane is clean for the next call
        r += c;
        if (c > 0.0) cl += c * log(c);
      }
      if (r > 0.0) acc += r * log(r) - cl;
      ntot += r;
    }
    const double P = a.periods[pi];
    h = acc / ntot;                                          // no binned sample: 0 / 0 = NaN
    if (!isfinite(P) || !isfinite(1.0 / P)) h = nan("");     // period 0, denormal, inf or NaN: no phases
    if (a.h_out) a.h_out[pi] = h;
    idx = pi;
  }
  block_argext<-1>(h, idx, sv, si);
  const int nblk = gridDim.x;
  if (threadIdx.x == 0) {
    a.red_val[blockIdx.x] = h;
    a.red_idx[blockIdx.x] = idx;
    __threadfence();
    s_last = atomicAdd(a.call_done, 1u) == (unsigned)(nblk - 1);
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) *a.call_done = 0u;
  __threadfence();
  double best = 0.0;
  long long bidx = -1;
  for (int k = threadIdx.x; k < nblk; k += blockDim.x) {
    const double v = __ldcg(a.red_val + k);
    const long long i = __ldcg(a.red_idx + k);
    if (better<-1>(v, i, best, bidx)) { best = v; bidx = i; }
  }
  block_argext<-1>(best, bidx, sv, si);
  if (threadIdx.x == 0) {
    if (a.arg_out) *a.arg_out = bidx;
    if (a.best_out) *a.best_out = bidx >= 0 ? best : nan("");
  }
}

// Shared memory: s_t[CE_TILE + pad] time stamps, s_thr[nphi + 1] thresholds, s_m[CE_TILE] magnitude-bin word offsets
// (mbin * VT, or 0xffffffff for a value in no cell), cnt[cells][VT] private count columns (VT = THREADS * PPT).
template <int THREADS, int PPT>
__global__ void __launch_bounds__(THREADS)
ce_hist_kernel(const CeArgs a) {
  constexpr int VT = THREADS * PPT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nphi = a.nphi, nm = a.nm, cells = a.cells;
  double* s_t = reinterpret_cast<double*>(smem_raw);
  double* s_thr = s_t + CE_TILE + CE_TILE_PAD;
  unsigned* s_m = reinterpret_cast<unsigned*>(s_thr + ((nphi + 2) & ~1));
  unsigned* cnt = s_m + CE_TILE;

  const int split = blockIdx.x % a.nsplit;
  const long long pb = blockIdx.x / a.nsplit;
  long long pis[PPT];
  bool valids[PPT];
  double Ps[PPT], rPs[PPT];
  bool in_range = true;
#pragma unroll
  for (int s = 0; s < PPT; ++s) {
    pis[s] = pb * VT + s * THREADS + threadIdx.x;
    valids[s] = pis[s] < a.np;
    double P = valids[s] ? a.periods[pis[s]] : 1.0;
    if (!isfinite(1.0 / P) || !isfinite(P)) P = 1.0;  // invalid trial period: the tail writes NaN
    Ps[s] = P;
    rPs[s] = 1.0 / P;
    // padding columns (P = 1) must not veto the block's fast path
    in_range = in_range && (!valids[s] || (fabs(rPs[s]) * a.meta->t_span < PDM_FAST_LIMIT &&
                                           fabs(rPs[s]) * a.meta->t_absmax < PDM_SHIFT_LIMIT));
  }
  const bool bad = a.meta->bad != 0;   // block-uniform
  const double nphid = (double)nphi;
  const unsigned nphiu = (unsigned)nphi;
  const double xlo = a.meta->xmin, xrange = a.meta->xmax - a.meta->xmin;
  // Shared-memory addresses are formed in 32 bits, in BYTES: one IMAD per update (bin * stride + sample offset) and the
  // column of the thread as an immediate -- the generic-pointer form cost 1.3 instructions more per update (ncu r02i).
  const unsigned mstride = (unsigned)nm * VT * 4u;   // bytes between consecutive phase bins of one column
  const unsigned col0 = (unsigned)__cvta_generic_to_shared(cnt + threadIdx.x);   // this thread's first column

  for (int k = threadIdx.x; k <= nphi; k += THREADS) s_thr[k] = (double)k / nphid;  // phase.py:138-140
  for (int k = threadIdx.x; k < CE_TILE_PAD; k += THREADS) s_t[CE_TILE + k] = 0.0;
  for (int k = threadIdx.x; k < cells * VT; k += THREADS) cnt[k] = 0u;

  const long long per = (a.n + a.nsplit - 1) / a.nsplit;
  const long long sb = (long long)split * per;
  const long long se = sb + per < a.n ? sb + per : a.n;
  // guard band of the block (see phase_common.cuh: shifted fixed-point phase on t - t0)
  __shared__ unsigned s_guard;
  if (threadIdx.x == 0) s_guard = 0u;
  __syncthreads();
  {
    unsigned g = 0u;
#pragma unroll
    for (int s = 0; s < PPT; ++s)
      if (valids[s] && in_range) g = max(g, pdm_guard_units(fabs(rPs[s]), a.meta->t_absmax));
    atomicMax(&s_guard, g);
  }
  __syncthreads();
  const unsigned guard2 = 2u * s_guard * nphiu;
  const double t0 = a.meta->t0;
  double magic[PPT];
#pragma unroll
  for (int s = 0; s < PPT; ++s) magic[s] = pdm_fast_magic(t0, rPs[s], s_guard);
  // block-uniform; a constant signal (max == min) has no magnitude bins at all (0 / 0): every sample is in no cell
  const bool fast = __syncthreads_and(in_range) != 0 && !bad && (xrange > 0.0 || a.counts_only);

  auto exact_bin = [&](double P, double rP, double tv, double& phi) {
    unsigned e;
    int k = pdm_bin(tv, P, rP, nphid, phi, e);
    if (e < PDM_AMBIG) k = pdm_fix_bin(k, phi, s_thr, nphi);
    return (unsigned)k;
  };
  // `base` = col0 + byte offset of the sample's magnitude bin (computed once per sample, shared by the thread's columns);
  // S = which of the thread's columns.  The increment is a literal: ptxas emits ATOMS.POPC.INC for +1, measured faster
  // here than ATOMS.ADD with a register operand (3.28 vs 3.39 ms on the C3 shape).
  auto add1 = [&](auto sc, unsigned base, unsigned k) {
    constexpr int S = decltype(sc)::value;
    asm volatile("red.shared.add.u32 [%0+%1], 1;" :: "r"(base + k * mstride), "n"(S * THREADS * 4) : "memory");
  };
  auto sub1 = [&](auto sc, unsigned base, unsigned k) {
    constexpr int S = decltype(sc)::value;
    asm volatile("red.shared.add.u32 [%0+%1], 0xffffffff;" :: "r"(base + k * mstride), "n"(S * THREADS * 4) : "memory");
  };

  long long tile0 = sb;
  do {
    long long left = se - tile0;
    const int cntv = left <= 0 ? 0 : (left < CE_TILE ? (int)left : CE_TILE);
    __syncthreads();
    for (int i = threadIdx.x; i < cntv; i += THREADS) {
      s_t[i] = fast ? a.t[tile0 + i] - t0 : a.t[tile0 + i];
      const int mb = a.counts_only ? 0 : ce_mbin(a.x[tile0 + i], xlo, xrange, nm);
      s_m[i] = mb < 0 ? 0xffffffffu : (unsigned)mb * VT * 4u;   // byte offset of the magnitude bin inside a phase bin
    }
    __syncthreads();

    if (fast) {
      // every sample is finite and in range: CE_U samples x PPT periods per trip, atomics issued at once with the fast
      // bins, the rare trip with a sample on a bin edge re-bins those exactly afterwards and moves the increment
      int i = 0;
      double tv[CE_U];
#pragma unroll
      for (int u = 0; u < CE_U; u += 2) {
        const double2 tt = *reinterpret_cast<const double2*>(s_t + u);
        tv[u] = tt.x;
        tv[u + 1] = tt.y;
      }
      for (; i + CE_U <= cntv; i += CE_U) {
        unsigned k[PPT][CE_U], mo[CE_U], pos, pmin = 0xffffffffu;
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
#pragma unroll
          for (int u = 0; u < CE_U; ++u) {
            k[s][u] = pdm_bin_fast_m(tv[u], rPs[s], magic[s], nphiu, pos);
            pmin = min(pmin, pos);
          }
        }
#pragma unroll
        for (int u = 0; u < CE_U; u += 2) {   // next trip's time stamps before this trip's atomics
          const double2 tt = *reinterpret_cast<const double2*>(s_t + i + CE_U + u);
          tv[u] = tt.x;
          tv[u + 1] = tt.y;
        }
#pragma unroll
        for (int u = 0; u < CE_U; u += 4) {
          const uint4 mm = *reinterpret_cast<const uint4*>(s_m + i + u);
          mo[u] = mm.x; mo[u + 1] = mm.y; mo[u + 2] = mm.z; mo[u + 3] = mm.w;
        }
#pragma unroll
        for (int u = 0; u < CE_U; ++u) {
          const unsigned base = col0 + mo[u];
          static_for<PPT>([&](auto sc) { add1(sc, base, k[decltype(sc)::value][u]); });
        }
        if (pmin < guard2) {
          for (int u = 0; u < CE_U; ++u) {
            static_for<PPT>([&](auto sc) {
              constexpr int s = decltype(sc)::value;
              unsigned p0;
              const unsigned kf = pdm_bin_fast_m(s_t[i + u], rPs[s], magic[s], nphiu, p0);
              if (p0 < guard2) {
                double ph;
                const unsigned ke = exact_bin(Ps[s], rPs[s], a.t[tile0 + i + u], ph);   // the ORIGINAL stamp
                if (ke != kf) {
                  sub1(sc, col0 + s_m[i + u], kf);   // counts are sums modulo 2^32: -1 undoes the update
                  add1(sc, col0 + s_m[i + u], ke);
                }
              }
            });
          }
        }
      }
      for (; i < cntv; ++i) {
        static_for<PPT>([&](auto sc) {
          constexpr int s = decltype(sc)::value;
          unsigned p0;
          unsigned kf = pdm_bin_fast_m(s_t[i], rPs[s], magic[s], nphiu, p0);
          if (p0 < guard2) {
            double ph;
            kf = exact_bin(Ps[s], rPs[s], a.t[tile0 + i], ph);
          }
          add1(sc, col0 + s_m[i], kf);
        });
      }
    } else {
      // exact FP64 phase for every sample; samples whose phase is NaN (non-finite stamp) or whose value is NaN are in no cell
      for (int i = 0; i < cntv; ++i) {
        const double tvu = s_t[i];
        const unsigned mo = s_m[i];
        static_for<PPT>([&](auto sc) {
          constexpr int s = decltype(sc)::value;
          double ph;
          unsigned k = exact_bin(Ps[s], rPs[s], tvu, ph);
          if (ph == ph && mo != 0xffffffffu) add1(sc, col0 + mo, min(k, nphiu - 1u));
        });
      }
    }
    tile0 += CE_TILE;
  } while (tile0 < se);

  // the thread's columns -> the count plane shared by all sample splits (RED.ADD.32; integer, hence order independent)
  __syncwarp();
#pragma unroll
  for (int s = 0; s < PPT; ++s) {
    if (!valids[s]) continue;
    unsigned* pcol = a.plane + pis[s];
    const unsigned* col = cnt + s * THREADS + threadIdx.x;
    for (int c = 0; c < cells; ++c) {
      const unsigned v = col[c * VT];
      if (v) atomicAdd(pcol + (long long)c * a.np, v);
    }
  }
}

static size_t ce_smem_bytes(int nphi, int cells, int vt) {
  return sizeof(double) * (CE_TILE + CE_TILE_PAD + ((nphi + 2) & ~1)) + sizeof(unsigned) * CE_TILE +
         sizeof(unsigned) * (size_t)cells * vt;
}

template <int THREADS, int PPT>
static int ce_launch(pdc_ctx* ctx, const CeArgs& a, size_t smem, long long blocks, cudaStream_t st) {
  PDC_CUDA(cudaFuncSetAttribute(ce_hist_kernel<THREADS, PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ce_hist_kernel<THREADS, PPT><<<(unsigned)blocks, THREADS, smem, st>>>(a);
  PDC_CUDA(cudaGetLastError());
  ctx->launches++;
  return PDC_OK;
}

// statistics + count histograms of one call (or one Gregory-Loredo pass): leaves the counts in ctx->hist_plane
// ([nphi * nm][np], marked dirty until an epilogue has cleared it).  x == NULL: phase bins only.
static int ce_hist_pass(pdc_ctx* ctx, const double* t, const double* x, int64_t n, const double* periods, int64_t np,
                        int nphi, int nm, cudaStream_t st) {
  if (n < 1) { set_error("pdc_ce: need at least 1 sample"); return PDC_EINVAL; }
  if (np < 1) { set_error("pdc_ce: need at least one trial period"); return PDC_EINVAL; }
  if (nphi < 1 || nm < 1) { set_error("pdc_ce: nphi and nm must be >= 1"); return PDC_EINVAL; }
  const long long cells_l = (long long)nphi * nm;
  const size_t smem_max = 227 * 1024;
  if (cells_l > 100000 || ce_smem_bytes(nphi, (int)cells_l, 32) > smem_max) {
    set_error("pdc_ce: nphi*nm = %lld cells do not fit a shared-memory histogram", cells_l);
    return PDC_EINVAL;
  }
  const int cells = (int)cells_l;
  // period columns per block: the candidate that keeps the most columns resident per SM
  int vt = 32, best_res = 0;
  const int cands[4] = {256, 128, 64, 32};
  for (int c = 0; c < 4; ++c) {
    size_t sm = ce_smem_bytes(nphi, cells, cands[c]);
    if (sm > smem_max) continue;
    int blocks = (int)((228 * 1024) / (sm + 1024));
    if (blocks > 2048 / cands[c]) blocks = 2048 / cands[c];
    int res = blocks * cands[c];
    if (res > best_res) { best_res = res; vt = cands[c]; }
  }
  const int ppt = vt >= 64 ? 2 : 1;
  const size_t smem = ce_smem_bytes(nphi, cells, vt);
  const long long resident = (long long)ctx->sm_count * (best_res / vt);
  const long long npb = (np + vt - 1) / vt;
  int nsplit = 1;
  if (npb < 24 * resident) {
    long long cap = n / 512;
    if (cap < 1) cap = 1;
    if (cap > 1024) cap = 1024;
    double best = 1e300;
    for (long long s = 1; s <= cap; ++s) {
      long long items = npb * s;
      long long waves = (items + resident - 1) / resident;
      double cost = (double)waves * ((double)((n + s - 1) / s) + 2.0 * cells + 64.0);
      if (cost < best * 0.999) { best = cost; nsplit = (int)s; }
      if (items > 64 * resident) break;
    }
  }
  const long long blocks = npb * nsplit;
  if (blocks > 0x7fffffffLL || npb > 0x3fffffffLL) { set_error("pdc_ce: problem too large for one call"); return PDC_EINVAL; }

  PDC_TRY(ctx->pdm_meta.reserve(sizeof(CeMeta) + 16 + sizeof(CePart) * CE_STATS_MAXBLK));
  {
    const void* before = ctx->hist_plane.p;
    const size_t cap_before = ctx->hist_plane.cap;
    PDC_TRY(ctx->hist_plane.reserve(sizeof(unsigned) * (size_t)cells * np));
    if (ctx->hist_plane.p != before || ctx->hist_plane.cap != cap_before || ctx->hist_plane_dirty) {
      PDC_CUDA(cudaMemsetAsync(ctx->hist_plane.p, 0, ctx->hist_plane.cap, st));
      ctx->hist_plane_dirty = false;
    }
  }
  const int eblk = (int)((np + 255) / 256);
  PDC_TRY(ctx->blockred.reserve((sizeof(double) + sizeof(long long)) * (size_t)eblk));
  {
    const void* before = ctx->pdm_cnt.p;
    PDC_TRY(ctx->pdm_cnt.reserve(sizeof(unsigned) * 4));
    if (ctx->pdm_cnt.p != before) PDC_CUDA(cudaMemsetAsync(ctx->pdm_cnt.p, 0, ctx->pdm_cnt.cap, st));
  }
  unsigned* cnt_stats = ctx->pdm_cnt.as<unsigned>();
  CeMeta* meta = ctx->pdm_meta.as<CeMeta>();
  {
    CePart* part = reinterpret_cast<CePart*>(reinterpret_cast<char*>(meta) + ((sizeof(CeMeta) + 15) & ~(size_t)15));
    long long sblk = (n + 8 * CE_STATS_THREADS - 1) / (8 * CE_STATS_THREADS);
    if (sblk > CE_STATS_MAXBLK) sblk = CE_STATS_MAXBLK;
    ce_stats_kernel<<<(unsigned)sblk, CE_STATS_THREADS, 0, st>>>(t, x ? x : t, n, part, cnt_stats, meta);
    PDC_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  CeArgs a;
  a.t = t;
  a.x = x;
  a.periods = periods;
  a.meta = meta;
  a.plane = ctx->hist_plane.as<unsigned>();
  a.n = n;
  a.np = np;
  a.nphi = nphi;
  a.nm = nm;
  a.cells = cells;
  a.nsplit = nsplit;
  a.counts_only = x == nullptr;

  ctx->hist_plane_dirty = true;
  PDC_TRY(ctx->main_begin(st));
  switch (vt * 8 + ppt) {
    case 256 * 8 + 2: PDC_TRY((ce_launch<128, 2>(ctx, a, smem, blocks, st))); break;
    case 128 * 8 + 2: PDC_TRY((ce_launch<64, 2>(ctx, a, smem, blocks, st))); break;
    case 64 * 8 + 2: PDC_TRY((ce_launch<32, 2>(ctx, a, smem, blocks, st))); break;
    default: PDC_TRY((ce_launch<32, 1>(ctx, a, smem, blocks, st))); break;
  }
  PDC_TRY(ctx->main_end(st));
  return PDC_OK;
}

int ce_run(pdc_ctx* ctx, const double* t, const double* x, int64_t n, const double* periods, int64_t np, int nphi,
           int nm, double* h_out, int64_t* argmin_out, double* min_out, cudaStream_t st) {
  if (!x) { set_error("pdc_ce: x is NULL"); return PDC_EINVAL; }
  ScratchScope scratch(ctx, st);
  PDC_TRY(scratch.acquire());
  PDC_TRY(ce_hist_pass(ctx, t, x, n, periods, np, nphi, nm, st));
  const int eblk = (int)((np + 255) / 256);
  CeEpiArgs e;
  e.plane = ctx->hist_plane.as<unsigned>();
  e.periods = periods;
  e.np = np;
  e.nphi = nphi;
  e.nm = nm;
  e.h_out = h_out;
  e.red_val = ctx->blockred.as<double>();
  e.red_idx = reinterpret_cast<long long*>(e.red_val + eblk);
  e.call_done = ctx->pdm_cnt.as<unsigned>() + 1;
  e.arg_out = (long long*)argmin_out;
  e.best_out = min_out;
  ce_epilogue_kernel<<<(unsigned)eblk, 256, 0, st>>>(e);
  PDC_CUDA(cudaGetLastError());
  ctx->launches++;
  ctx->hist_plane_dirty = false;
  PDC_TRY(scratch.release());
  return PDC_OK;
}

// ---------------------------------------------------------------------------
// Gregory-Loredo (1992) periodogram for event arrival times
// ---------------------------------------------------------------------------
// A TODO of the reference (phase.py:14); PARITY UNPINNED BY THE REFERENCE (oracle: oracle/gl_numpy.py).  For N events,
// trial period P and a stepwise light-curve model with m phase bins, the likelihood marginalised over the bin rates is
// inversely proportional to the multiplicity W_m(P, phi) = N! / prod_j n_j! of the binned events, and the odds ratio of
// the m-bin periodic model against the constant model, marginalised over the phase offset phi, is
//     O_m(P) = [ (m - 1)! / (N + m - 1)! ] m^N  < prod_j n_j(P, phi)! >_phi                       (GL 1992, eq. 5.25 ff.)
// The phase offset is integrated on a grid of nc offsets per bin: events are counted in m * nc fine phase bins (the
// reference's phase and edge conventions, phase.py:131,138-140) and the m coarse bins at offset c are circular unions
// of nc consecutive fine bins starting at c -- the "covers" structure of PDM.  One count-histogram pass (the
// conditional-entropy kernel with phase bins only) per m = 2 .. m_max; this kernel turns the counts of pass m into
// ln O_m and folds it into the running ln sum_m O_m per trial period.  The periodogram is
//     ln O(P) = ln [ (1 / (m_max - 1)) sum_{m=2}^{m_max} O_m(P) ]      (equal prior weight for every m),
// to be MAXIMISED.
struct GlEpiArgs {
  unsigned* plane;        // [m * nc][np] counts of this pass; read and cleared
  const double* periods;
  long long np;
  int m, nc, last;        // last: this is the m_max pass -> normalise, store, arg-max
  double ln_prior;        // lgamma(m) - lgamma(N + m) + N ln m   (N = number of events)
  double ln_norm;         // ln(m_max - 1)
  double* acc;            // [np] running ln sum_m O_m (read-modify-write by the thread that owns the period)
  double* out;            // [np]
  double* red_val;
  long long* red_idx;
  unsigned* call_done;
  long long* arg_out;
  double* best_out;
};

__global__ void __launch_bounds__(256)
gl_epilogue_kernel(const GlEpiArgs a) {
  __shared__ double sv[32];
  __shared__ long long si[32];
  __shared__ int s_last;
  const long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long np = a.np;
  const int m = a.m, nc = a.nc, F = m * nc;
  double val = 0.0;
  long long idx = -1;
  if (pi < np) {
    unsigned* base = a.plane + pi;
    // ln < prod_j n_j! >_phi  =  logsumexp_c( sum_j lgamma(n_j(c) + 1) ) - ln nc
    double mx = -INFINITY, se = 0.0;
    for (int c = 0; c < nc; ++c) {
      double L = 0.0;
      for (int j = 0; j < m; ++j) {
        unsigned nj = 0;
        for (int k = 0; k < nc; ++k) {
          int f = c + j * nc + k;
          if (f >= F) f -= F;
          nj += __ldcg(base + (long long)f * np);
        }
        L += lgamma((double)nj + 1.0);
      }
      if (L > mx) { se = se * exp(mx - L) + 1.0; mx = L; }
      else se += exp(L - mx);
    }
    for (int f = 0; f < F; ++f) base[(long long)f * np] = 0u;    // the plane is clean for the next pass / call
    const double ln_om = a.ln_prior + mx + log(se) - log((double)nc);
    double acc = a.m == 2 ? ln_om : a.acc[pi];
    if (a.m != 2) {   // logaddexp
      const double hi = fmax(acc, ln_om), lo = fmin(acc, ln_om);
      acc = hi + log1p(exp(lo - hi));
    }
    a.acc[pi] = acc;
    if (a.last) {
      const double P = a.periods[pi];
      val = acc - a.ln_norm;
      if (!isfinite(P) || !isfinite(1.0 / P)) val = nan("");
      a.out[pi] = val;
      idx = pi;
    }
  }
  if (!a.last) return;
  block_argext<+1>(val, idx, sv, si);
  const int nblk = gridDim.x;
  if (threadIdx.x == 0) {
    a.red_val[blockIdx.x] = val;
    a.red_idx[blockIdx.x] = idx;
    __threadfence();
    s_last = atomicAdd(a.call_done, 1u) == (unsigned)(nblk - 1);
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) *a.call_done = 0u;
  __threadfence();
  double best = 0.0;
  long long bidx = -1;
  for (int k = threadIdx.x; k < nblk; k += blockDim.x) {
    const double v = __ldcg(a.red_val + k);
    const long long i = __ldcg(a.red_idx + k);
    if (better<+1>(v, i, best, bidx)) { best = v; bidx = i; }
  }
  block_argext<+1>(best, bidx, sv, si);
  if (threadIdx.x == 0) {
    if (a.arg_out) *a.arg_out = bidx;
    if (a.best_out) *a.best_out = bidx >= 0 ? best : nan("");
  }
}

int gl_run(pdc_ctx* ctx, const double* t, int64_t n, const double* periods, int64_t np, int m_max, int nc,
           double* lnodds_out, int64_t* argmax_out, double* max_out, cudaStream_t st) {
  if (n < 1 || np < 1) { set_error("pdc_gl: need n >= 1 events and np >= 1 periods"); return PDC_EINVAL; }
  if (m_max < 2 || nc < 1 || (long long)m_max * nc > 4096) {
    set_error("pdc_gl: need m_max >= 2, nc >= 1 and m_max * nc <= 4096");
    return PDC_EINVAL;
  }
  ScratchScope scratch(ctx, st);
  PDC_TRY(scratch.acquire());
  PDC_TRY(ctx->gl_acc.reserve(sizeof(double) * (size_t)np));
  {   // size the count plane for the largest pass once (the passes must not reallocate it between them)
    const void* before = ctx->hist_plane.p;
    const size_t cap_before = ctx->hist_plane.cap;
    PDC_TRY(ctx->hist_plane.reserve(sizeof(unsigned) * (size_t)m_max * nc * np));
    if (ctx->hist_plane.p != before || ctx->hist_plane.cap != cap_before || ctx->hist_plane_dirty) {
      PDC_CUDA(cudaMemsetAsync(ctx->hist_plane.p, 0, ctx->hist_plane.cap, st));
      ctx->hist_plane_dirty = false;
    }
  }
  const int eblk = (int)((np + 255) / 256);
  for (int m = 2; m <= m_max; ++m) {
    PDC_TRY(ce_hist_pass(ctx, t, nullptr, n, periods, np, m * nc, 1, st));
    GlEpiArgs e;
    e.plane = ctx->hist_plane.as<unsigned>();
    e.periods = periods;
    e.np = np;
    e.m = m;
    e.nc = nc;
    e.last = m == m_max;
    e.ln_prior = lgamma((double)m) - lgamma((double)n + (double)m) + (double)n * log((double)m);
    e.ln_norm = log((double)(m_max - 1));
    e.acc = ctx->gl_acc.as<double>();
    e.out = lnodds_out;
    e.red_val = ctx->blockred.as<double>();
    e.red_idx = reinterpret_cast<long long*>(e.red_val + eblk);
    e.call_done = ctx->pdm_cnt.as<unsigned>() + 1;
    e.arg_out = (long long*)argmax_out;
    e.best_out = max_out;
    gl_epilogue_kernel<<<(unsigned)eblk, 256, 0, st>>>(e);
    PDC_CUDA(cudaGetLastError());
    ctx->launches++;
    ctx->hist_plane_dirty = false;
  }
  PDC_TRY(scratch.release());
  return PDC_OK;
}

}  // namespace pdc
