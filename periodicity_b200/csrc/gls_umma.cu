// GLS trial-frequency sums on the tcgen05 tensor cores (sm_100a): the second formulation of the hot kernel.
//
// Same sums as gls_strip_kernel (gls.cu) -- the exact C_j = sum_i w_i cos(2 pi f_j t_i), ... that replace the
// reference's `_trig_sum` (src/periodicity/spectral.py:11-40,109-111) -- but the angle addition runs across BLOCKS of
// the uniform frequency grid instead of along a thread's strip.  Write the grid index as j = 128 cb + k (coarse block
// cb, fine index k < 128).  For sample i the phase splits into psi = f_{128 cb} tau_i (+ gauge) and phi = k df tau_i, and
//     w cos(psi + phi) = [cos phi, sin phi] . [ w cos psi, -w sin psi]
//     w sin(psi + phi) = [cos phi, sin phi] . [ w sin psi,  w cos psi]
// so every sum over the samples is a GEMM over the sample axis: a FINE operand of 128 rows (one per k, two K-slots
// per sample) against a COARSE operand with one row per (coarse block, sum).  The six sums of spectral.py:109-111 are
//   type-1 tiles: {C, S, YC, YS}[cb][k]            coarse rows (w c, -w s), (w s, w c), (w y c, -w y s), (w y s, w y c)
//   type-2 tiles: {C2, S2}[cb][k] at the doubled angle (the reference's second `_trig_sum` at 2 f, spectral.py:110)
// Operands cost O((128 + nf / 128) N) sincos instead of O(nf N) FP32 steps; they never exist in global memory: producer
// warps compute them from the per-sample records (phase reduced mod 1 in FP64, MUFU sin/cos) straight into shared memory
// in the tcgen05 no-swizzle K-major layout.
//
// Precision.  fp16 inputs with FP32 accumulation, every operand split x = hi + lo (two fp16 numbers, 22 significant bits)
// and the product taken as hi hi + hi lo + lo hi: three tcgen05.mma per K-step.  The TMEM accumulator TRUNCATES toward
// zero (measured: tools/microbench/umma_probe.cu, profiles/r02/umma_probe.txt), a bias of up to one ulp per instruction,
// so an accumulation run in TMEM is only UM_CHUNK_STAGES * 16 = 64 samples long (24 instructions): epilogue warps drain
// the tile (double-buffered in TMEM) and add it to FP32 master accumulators in registers with round-to-nearest.  The
// masters leave the kernel once per job as 64-bit fixed point through RED.ADD.64 into the same plane gls_strip_kernel
// uses, so everything downstream (FP64 sub-cycle bins, FP64 epilogue, arg-max, fan-out) is shared.
//
// Roles in a CTA of 640 threads, one CTA per SM:
//   warps 0-7    epilogue: TMEM -> registers (+=), final flush
//   warps 8-15   producers: records -> fp16 hi/lo operand tiles, 16 samples per stage, 4 stages
//   warp 16      one elected lane issues tcgen05.mma / tcgen05.commit; the warp owns the TMEM allocation
// All waits are bounded (clock-based): a protocol error ends the kernel with a status word, it cannot hang the GPU.
#include "gls_common.cuh"
#include "umma.cuh"

namespace pdc {

using namespace umma;

constexpr int UM_FINE = 128;           // fine indices per tile = MMA M = TMEM lanes
constexpr int UM_STAGE_SAMPLES = 16;   // 32 K-slots = two K = 16 steps
constexpr int UM_NSTAGES = 4;
constexpr int UM_CHUNK_STAGES = 4;     // one TMEM accumulation run
constexpr int UM_BLOCK = 256;          // samples per block of staged records
constexpr int UM_THREADS = 640;
constexpr int UM_MAX_T1 = 64, UM_MAX_T2 = 128;   // coarse blocks per tile (N = 4 * 64 = 2 * 128 = 256 columns)
// one stage in shared memory: [K-chunk of 8 slots][row][8 halves]; 16-byte rows, 8-row groups contiguous (SBO = 128 B)
constexpr uint32_t UM_FINE_HI = 0, UM_FINE_LO = 8192, UM_COARSE_HI = 16384, UM_COARSE_LO = 32768;
constexpr uint32_t UM_STAGE_BYTES = 49152;
constexpr uint32_t UM_LBO_FINE = 128 * 16, UM_LBO_COARSE = 256 * 16, UM_SBO = 128;
constexpr uint32_t UM_REC_BYTES = 2 * UM_BLOCK * (8 + 8 + 4 + 4);
constexpr uint32_t UM_SMEM_BYTES = UM_NSTAGES * UM_STAGE_BYTES + UM_REC_BYTES + 256;
constexpr long long UM_WAIT_CLOCKS = 4000000000LL;   // ~2 s: far beyond any legitimate wait

struct GlsUmmaArgs {
  const GlsCurve* curves;
  const double2* rec1;
  const float4* rec2;
  unsigned long long* partial;   // [6][nf_tot] fixed point; planes 4, 5 receive sum w cos 2x, sum w sin 2x
  long long nf, nf_tot, j0;
  int nC;                        // coarse blocks per curve
  int nt1, nt2, cpt1, cpt2;      // tiles per curve and coarse blocks per tile of each type
  int nsplit;
  int weighted;
  float fix_scale;
  int* status;                   // set non-zero on a protocol time-out
};

__device__ __forceinline__ bool um_wait(uint32_t bar, uint32_t parity, volatile int* s_abort, long long t_start) {
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 256; ++i)
      if (mbar_try_wait(bar, parity)) return true;
    if (*s_abort || clock64() - t_start > UM_WAIT_CLOCKS) {
      *s_abort = 1;
      return false;
    }
  }
}

// x = hi + lo with hi, lo fp16 (hi rounded to nearest, lo the rounded remainder): packed (c, s) pair
__device__ __forceinline__ void um_split2(float c, float s, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(c, s);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(c - hf.x, s - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void um_sincos_turns(double v, float& c, float& s) {
  // v = phase + 1.5 * 2^20: the low mantissa word is the fraction in units of 2^-32 turn (two's complement)
  const float x = (float)__double2loint(v) * 1.4629180792671596e-9f;   // 2 pi / 2^32
  __sincosf(x, &s, &c);
}
__device__ __forceinline__ void um_sts128(uint32_t addr, const uint32_t (&v)[4]) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(UM_THREADS, 1)
gls_umma_kernel(const GlsUmmaArgs a) {
  extern __shared__ __align__(1024) unsigned char um_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long t_start = clock64();

  // ---- job ----
  int job = blockIdx.x;
  const int split = job % a.nsplit;
  job /= a.nsplit;
  const int ntile = a.nt1 + a.nt2;
  const int tile = job % ntile;
  const int curve = job / ntile;
  const bool type2 = tile >= a.nt1;
  const int cpt = type2 ? a.cpt2 : a.cpt1;                       // coarse blocks per tile (padded to 4 / 8)
  const int cb0 = (type2 ? tile - a.nt1 : tile) * cpt;           // first coarse block of this tile
  const int ncb = min(cpt, a.nC - cb0);                          // real coarse blocks (> 0 by construction)
  const int rows_per_cb = type2 ? 2 : 4;
  const int N = cpt * rows_per_cb;                               // MMA N: multiple of 16, <= 256

  const GlsCurve* cvp = a.curves + curve;
  const long long cbegin = cvp->begin, cn = cvp->n;
  const long long per = (cn + a.nsplit - 1) / a.nsplit;
  const long long sb = (long long)split * per;
  const long long se = sb + per < cn ? sb + per : cn;
  const long long ns = se > sb ? se - sb : 0;
  const int nchunks = (int)((ns + UM_CHUNK_STAGES * UM_STAGE_SAMPLES - 1) / (UM_CHUNK_STAGES * UM_STAGE_SAMPLES));
  const int nstages = nchunks * UM_CHUNK_STAGES;                 // padded with zero-weight samples
  if (nchunks == 0) return;                                      // block-uniform: nothing to add

  // ---- shared memory ----
  const uint32_t smem0 = smem_u32(um_smem);
  unsigned char* recs = um_smem + UM_NSTAGES * UM_STAGE_BYTES;
  double* s_b = reinterpret_cast<double*>(recs);                               // [2][UM_BLOCK] phase step per index (turns)
  double* s_A = s_b + 2 * UM_BLOCK;                                            // [2][UM_BLOCK] phase at the tile's base frequency
  float* s_wy = reinterpret_cast<float*>(s_A + 2 * UM_BLOCK);                  // [2][UM_BLOCK] w' y'
  float* s_w = s_wy + 2 * UM_BLOCK;                                            // [2][UM_BLOCK] w' (0 for padding samples)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + 2 * UM_BLOCK);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * UM_NSTAGES;
  const uint32_t bar_tfull = bar_empty + 8 * UM_NSTAGES, bar_tempty = bar_tfull + 16;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * UM_NSTAGES + 4);
  volatile int* s_abort = reinterpret_cast<volatile int*>(s_tmem + 1);

  if (tid == 0) {
    for (int s = 0; s < UM_NSTAGES; ++s) {
      mbar_init(bar_full + 8 * s, 8);     // one arrival per producer warp
      mbar_init(bar_empty + 8 * s, 1);    // tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);    // tcgen05.commit
      mbar_init(bar_tempty + 8 * s, 8);   // one arrival per epilogue warp
    }
    mbar_init_fence();
    *s_abort = 0;
  }
  if (warp == 16) {
    tmem_alloc(smem_u32(s_tmem), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp < 8) {
    // =====================================================================================================
    // epilogue warps: lane quarter q = warp & 3 (TMEM lanes 32 q ..), column half h = warp >> 2
    // =====================================================================================================
    setmaxnreg_inc<168>();
    const int q = warp & 3, h = warp >> 2;
    float m[128];
#pragma unroll
    for (int c = 0; c < 128; ++c) m[c] = 0.f;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16) + h * 128;
    bool ok = true;
    for (int ch = 0; ch < nchunks && ok; ++ch) {
      const int acc = ch & 1;
      ok = um_wait(bar_tfull + 8 * acc, (ch >> 1) & 1, s_abort, t_start);
      tc_fence_after();
      if (ok) {
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 16) {
          if (h * 128 + c0 < N) {   // warp-uniform
            uint32_t r[16];
            tmem_ld16(tlane + acc * 256 + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 16; ++u) m[c0 + u] += __uint_as_float(r[u]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
    }
    if (ok) {
      // flush: column n = 128 h + c  ->  (sum s, coarse block cb) = (n / cpt, n % cpt); row = fine index k
      const int k = q * 32 + lane;
      int s = (h * 128) / cpt, cb = (h * 128) % cpt;
      const int plane0 = type2 ? 4 : 0;
#pragma unroll
      for (int c = 0; c < 128; ++c) {
        const long long j = (long long)(cb0 + cb) * UM_FINE + k;
        if (h * 128 + c < N && cb < ncb && j < a.nf) {
          unsigned long long* p = a.partial + (long long)(plane0 + s) * a.nf_tot + (long long)curve * a.nf + j;
          atomicAdd(p, (unsigned long long)__float2ll_rn(m[c] * a.fix_scale));
        }
        if (++cb == cpt) { cb = 0; ++s; }
      }
    }
  } else if (warp < 16) {
    // =====================================================================================================
    // producers
    // =====================================================================================================
    setmaxnreg_dec<56>();
    const int p = tid - 256;
    const bool weighted_tt = a.weighted && cvp->three_term;
    const int yslot = rec_slot(REC_Y), wslot = rec_slot(REC_W);
    const double kmul = type2 ? 2.0 : 1.0;
    // tile base frequency and its gauge (the per-index phase origin gamma of the records, see gls.cu)
    const long long jT = a.j0 + (long long)cb0 * UM_FINE;
    const double fT = cvp->fmin + (double)jT * cvp->df;
    double gT = (double)jT * cvp->gamma;
    gT -= floor(gT);
    const double MAGIC = 1572864.0;   // 1.5 * 2^20
    // fine task: row prow, samples 8 phalf .. 8 phalf + 7 of every stage
    const int prow = p & 127, phalf = p >> 7;
    const double kd = kmul * (double)prow;
    // coarse task(s)
    const int ccb = type2 ? (p & 127) : (p & 63);
    const double cd = kmul * (double)(ccb * UM_FINE);
    const bool cactive = ccb < cpt;

    auto load_block = [&](int blk, double2& r1, float4& r2, bool& in) {
      const long long i = sb + (long long)blk * UM_BLOCK + p;
      in = i < se;
      if (in) {
        r1 = a.rec1[cbegin + i];
        r2 = a.rec2[cbegin + i];
      }
    };
    auto store_block = [&](int blk, const double2& r1, const float4& r2, bool in) {
      const int o = (blk & 1) * UM_BLOCK + p;
      if (in) {
        const float yv = rec_get(r2, yslot), wv = rec_get(r2, wslot);
        s_b[o] = r1.y;
        s_A[o] = kmul * (frac_of_product(fT, r1.x) + gT) + MAGIC;
        s_wy[o] = weighted_tt ? wv * yv : yv;      // three-term records carry sqrt(w'), sqrt(w') y'
        s_w[o] = weighted_tt ? wv * wv : wv;
      } else {
        s_b[o] = 0.0;
        s_A[o] = MAGIC;
        s_wy[o] = 0.f;
        s_w[o] = 0.f;
      }
    };
    const int nblk = (nstages * UM_STAGE_SAMPLES + UM_BLOCK - 1) / UM_BLOCK;
    {
      double2 r1 = make_double2(0.0, 0.0);
      float4 r2 = make_float4(0.f, 0.f, 0.f, 0.f);
      bool in;
      load_block(0, r1, r2, in);
      store_block(0, r1, r2, in);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    bool ok = true;
    int g = 0;   // global stage counter
    for (int blk = 0; blk < nblk && ok; ++blk) {
      double2 n1 = make_double2(0.0, 0.0);
      float4 n2 = make_float4(0.f, 0.f, 0.f, 0.f);
      bool nin = false;
      if (blk + 1 < nblk) load_block(blk + 1, n1, n2, nin);
      const int rbase = (blk & 1) * UM_BLOCK;
      for (int st = 0; st < UM_BLOCK / UM_STAGE_SAMPLES && g < nstages; ++st, ++g) {
        const int slot = g % UM_NSTAGES;
        ok = um_wait(bar_empty + 8 * slot, ((g / UM_NSTAGES) & 1) ^ 1, s_abort, t_start);
        if (!ok) break;
        const uint32_t sbase = smem0 + slot * UM_STAGE_BYTES;
        const int ro = rbase + st * UM_STAGE_SAMPLES;
        // ---- fine operand: (cos, sin)(kmul k b_i) ----
#pragma unroll
        for (int hq = 0; hq < 2; ++hq) {
          const int quad = 2 * phalf + hq;
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const double v = __fma_rn(kd, s_b[ro + quad * 4 + u], MAGIC);
            float c, s;
            um_sincos_turns(v, c, s);
            um_split2(c, s, hi[u], lo[u]);
          }
          const uint32_t off = quad * UM_LBO_FINE + prow * 16;
          um_sts128(sbase + UM_FINE_HI + off, hi);
          um_sts128(sbase + UM_FINE_LO + off, lo);
        }
        // ---- coarse operand ----
        if (cactive) {
          if (!type2) {
            const int quad = p >> 6;
            float c[4], s[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const double v = __fma_rn(cd, s_b[ro + quad * 4 + u], s_A[ro + quad * 4 + u]);
              um_sincos_turns(v, c[u], s[u]);
            }
            const uint32_t off = quad * UM_LBO_COARSE + ccb * 16;
#pragma unroll
            for (int yy = 0; yy < 2; ++yy) {
              uint32_t ph[4], pl[4], rw[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float wv = yy ? s_wy[ro + quad * 4 + u] : s_w[ro + quad * 4 + u];
                um_split2(wv * c[u], wv * s[u], ph[u], pl[u]);
              }
              const uint32_t rowc = off + (uint32_t)((2 * yy) * cpt) * 16, rows = off + (uint32_t)((2 * yy + 1) * cpt) * 16;
#pragma unroll
              for (int u = 0; u < 4; ++u) rw[u] = ph[u] ^ 0x80000000u;          // (w c, -w s)
              um_sts128(sbase + UM_COARSE_HI + rowc, rw);
#pragma unroll
              for (int u = 0; u < 4; ++u) rw[u] = pl[u] ^ 0x80000000u;
              um_sts128(sbase + UM_COARSE_LO + rowc, rw);
#pragma unroll
              for (int u = 0; u < 4; ++u) rw[u] = __byte_perm(ph[u], 0, 0x1032);  // (w s, w c)
              um_sts128(sbase + UM_COARSE_HI + rows, rw);
#pragma unroll
              for (int u = 0; u < 4; ++u) rw[u] = __byte_perm(pl[u], 0, 0x1032);
              um_sts128(sbase + UM_COARSE_LO + rows, rw);
            }
          } else {
#pragma unroll
            for (int hq = 0; hq < 2; ++hq) {
              const int quad = 2 * (p >> 7) + hq;
              uint32_t ph[4], pl[4], rw[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const double v = __fma_rn(cd, s_b[ro + quad * 4 + u], s_A[ro + quad * 4 + u]);
                float c, s;
                um_sincos_turns(v, c, s);
                const float wv = s_w[ro + quad * 4 + u];
                um_split2(wv * c, wv * s, ph[u], pl[u]);
              }
              const uint32_t rowc = quad * UM_LBO_COARSE + ccb * 16, rows = rowc + (uint32_t)cpt * 16;
#pragma unroll
              for (int u = 0; u < 4; ++u) rw[u] = ph[u] ^ 0x80000000u;
              um_sts128(sbase + UM_COARSE_HI + rowc, rw);
#pragma unroll
              for (int u = 0; u < 4; ++u) rw[u] = pl[u] ^ 0x80000000u;
              um_sts128(sbase + UM_COARSE_LO + rowc, rw);
#pragma unroll
              for (int u = 0; u < 4; ++u) rw[u] = __byte_perm(ph[u], 0, 0x1032);
              um_sts128(sbase + UM_COARSE_HI + rows, rw);
#pragma unroll
              for (int u = 0; u < 4; ++u) rw[u] = __byte_perm(pl[u], 0, 0x1032);
              um_sts128(sbase + UM_COARSE_LO + rows, rw);
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * slot);
      }
      if (blk + 1 < nblk) store_block(blk + 1, n1, n2, nin);
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  } else {
    // =====================================================================================================
    // MMA issuer (warp 16); warps 17-19 only give their registers away
    // =====================================================================================================
    setmaxnreg_dec<24>();
    if (warp == 16 && lane == 0) {
      const uint32_t idesc = idesc_f16_f32(UM_FINE, N);
      bool ok = true;
      for (int g = 0; g < nstages && ok; ++g) {
        const int slot = g % UM_NSTAGES, ch = g / UM_CHUNK_STAGES, acc = ch & 1, first = (g % UM_CHUNK_STAGES) == 0;
        if (first) {
          ok = um_wait(bar_tempty + 8 * acc, ((ch >> 1) & 1) ^ 1, s_abort, t_start);
          if (!ok) break;
        }
        ok = um_wait(bar_full + 8 * slot, (g / UM_NSTAGES) & 1, s_abort, t_start);
        if (!ok) break;
        tc_fence_after();
        const uint32_t sbase = smem0 + slot * UM_STAGE_BYTES;
        const uint32_t d = tmem + acc * 256;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t ah = smem_desc(sbase + UM_FINE_HI + ks * 2 * UM_LBO_FINE, UM_LBO_FINE, UM_SBO);
          const uint64_t al = smem_desc(sbase + UM_FINE_LO + ks * 2 * UM_LBO_FINE, UM_LBO_FINE, UM_SBO);
          const uint64_t bh = smem_desc(sbase + UM_COARSE_HI + ks * 2 * UM_LBO_COARSE, UM_LBO_COARSE, UM_SBO);
          const uint64_t bl = smem_desc(sbase + UM_COARSE_LO + ks * 2 * UM_LBO_COARSE, UM_LBO_COARSE, UM_SBO);
          mma_f16_ss(d, al, bh, idesc, !(first && ks == 0));
          mma_f16_ss(d, ah, bl, idesc, 1);
          mma_f16_ss(d, ah, bh, idesc, 1);
        }
        mma_commit(bar_empty + 8 * slot);
        if ((g % UM_CHUNK_STAGES) == UM_CHUNK_STAGES - 1) mma_commit(bar_tfull + 8 * acc);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, 512);
  if (tid == 0 && *s_abort) *a.status = 1;
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
bool gls_umma_eligible(const pdc_ctx* ctx, int64_t B, int64_t nf, long long ntot, long long nmax, bool weighted,
                       const double* df_host) {
  if (ctx->gls_umma == 0) return false;
  if (nf < 2 * UM_FINE) return false;
  if (weighted && nmax > 60000) return false;   // |w' y'| <= n must stay inside fp16's range (gls_umma.cu, DESIGN)
  for (int64_t b = 0; b < B; ++b)
    if (!(df_host[b] > 0.0)) return false;      // the per-index phase step must be a forward grid
  if (ctx->gls_umma == 1) return true;           // forced on
  return (double)ntot * (double)nf >= 2.0e8;     // below that the strip kernel's launch is already ~60 us
}

int gls_umma_launch(pdc_ctx* ctx, const GlsCurve* curves, const double2* rec1, const float4* rec2,
                    unsigned long long* plane, int64_t B, int64_t nf, int64_t j0, long long nmax, bool weighted,
                    float fix_scale, cudaStream_t st) {
  GlsUmmaArgs a;
  a.curves = curves;
  a.rec1 = rec1;
  a.rec2 = rec2;
  a.partial = plane;
  a.nf = nf;
  a.nf_tot = (long long)B * nf;
  a.j0 = j0;
  a.nC = (int)((nf + UM_FINE - 1) / UM_FINE);
  a.nt1 = (a.nC + UM_MAX_T1 - 1) / UM_MAX_T1;
  a.cpt1 = (((a.nC + a.nt1 - 1) / a.nt1) + 3) & ~3;
  a.nt2 = (a.nC + UM_MAX_T2 - 1) / UM_MAX_T2;
  a.cpt2 = (((a.nC + a.nt2 - 1) / a.nt2) + 7) & ~7;
  // rounding up the tile size can leave the last tile(s) of a type empty: drop them
  while (a.nt1 > 1 && (long long)(a.nt1 - 1) * a.cpt1 >= a.nC) --a.nt1;
  while (a.nt2 > 1 && (long long)(a.nt2 - 1) * a.cpt2 >= a.nC) --a.nt2;
  a.weighted = weighted ? 1 : 0;
  a.fix_scale = fix_scale;
  PDC_TRY(ctx->umma_status.reserve(sizeof(int)));
  a.status = ctx->umma_status.as<int>();

  // sample splits: fill whole waves of one CTA per SM; a job should keep >= 1024 samples (its flush is 32768 REDs)
  const long long base_jobs = (long long)B * (a.nt1 + a.nt2);
  int nsplit = 1;
  if (ctx->gls_umma_nsplit > 0) nsplit = ctx->gls_umma_nsplit;
  else if (base_jobs < 6LL * ctx->sm_count) {
    long long cap = nmax / 1024;
    if (cap < 1) cap = 1;
    double best = 1e300;
    for (long long s = 1; s <= cap && s <= 1024; ++s) {
      const long long jobs = base_jobs * s;
      const long long waves = (jobs + ctx->sm_count - 1) / ctx->sm_count;
      const double per = (double)((nmax + s - 1) / s) + 768.0;   // per-job fixed cost (set-up + flush) in samples
      const double cost = (double)waves * per;
      if (cost < best * 0.999) { best = cost; nsplit = (int)s; }
      if (jobs > 16LL * ctx->sm_count) break;
    }
  }
  a.nsplit = nsplit;
  const long long jobs = base_jobs * nsplit;
  if (jobs > 0x7fffffffLL) { set_error("pdc_gls: problem too large for one call (%lld jobs)", jobs); return PDC_EINVAL; }

  static bool attr_set[64] = {};
  if (ctx->device >= 0 && ctx->device < 64 && !attr_set[ctx->device]) {
    PDC_CUDA(cudaFuncSetAttribute(gls_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UM_SMEM_BYTES));
    attr_set[ctx->device] = true;
  }
  if (!ctx->umma_status_clean) {
    PDC_CUDA(cudaMemsetAsync(a.status, 0, sizeof(int), st));
    ctx->umma_status_clean = true;
  }
  gls_umma_kernel<<<(unsigned)jobs, UM_THREADS, UM_SMEM_BYTES, st>>>(a);
  PDC_CUDA(cudaGetLastError());
  ctx->launches++;
  return PDC_OK;
}

}  // namespace pdc
