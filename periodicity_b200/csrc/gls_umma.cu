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
// warps compute them from the per-sample records (phases in 2^-32-turn integer arithmetic, MUFU sin/cos) straight into
// shared memory in the tcgen05 no-swizzle K-major layout.  (One long curve: the fine operand is the same for all its coarse
// tiles and is computed once per call, gls_umma_fine_kernel; from 16384 frequencies on a PAIR of CTAs works a tile of 256
// fine indices, gls_umma2.cu.  This file is the one-CTA kernel: batches, short grids.)
//
// Precision.  fp16 inputs with FP32 accumulation, every operand split x = hi + lo (two fp16 numbers, 22 significant bits)
// and the product taken as hi hi + hi lo + lo hi: three tcgen05.mma per K-step.  The TMEM accumulator TRUNCATES toward
// zero (measured: tools/microbench/umma_probe.cu, profiles/r02/umma_probe.txt), a bias of up to one ulp per instruction,
// so an accumulation run in TMEM is only chunk_stages * 16 = 64 ... 256 samples long: worker warps drain the tile
// (double-buffered in TMEM), multiply it by 1 + the expected truncation loss of the run (see rz_comp in gls_umma_launch) and
// add it to FP32 master accumulators in registers with round-to-nearest.  The masters leave the kernel once per job (at
// most 16384 samples) as 64-bit fixed point through RED.ADD.64 into the same plane gls_strip_kernel
// uses, so everything downstream (FP64 sub-cycle bins, FP64 epilogue, arg-max, fan-out) is shared.
//
// One CTA of 640 threads per SM: 16 identical worker warps (records -> fp16 hi/lo operand tiles, 16 samples per stage, 4
// stages; one accumulation run behind, TMEM -> FP32 masters in registers; final flush) and one warp group whose first
// warp issues the tcgen05.mma + tcgen05.commit of every stage.  Issuing blocks while the tensor core is busy, so the
// issuer cannot be a worker (tried: fixed worker warp 0.78 ms on C2, rotating duty 0.74 ms -- each issuer became the next
// pair's straggler); a 17th warp caps the launch at 96 registers per thread, so the group gives its registers to the
// workers (setmaxnreg 24 / 112).
// (A first version had 8 dedicated epilogue warps with 128 masters each and 8 producer warps squeezed into 56
// registers: 0.73 ms on C2, the producers latency-bound at a quarter of the issue rate.)
// All waits are bounded (clock-based): a protocol error ends the kernel with a status word, it cannot hang the GPU.
#include "gls_umma_common.cuh"

namespace pdc {

// Fine operand of a whole curve, once per call, when many coarse tiles share it (one long curve: C2 has 13 + 7 tiles per
// sample split, C5 1221 + 611): (cos, sin)(kmul k b_i) for k < 128 as fp16 hi / lo, written as the 16 KB shared-memory
// image of every 16-sample stage ([hi | lo][K-chunk of 4 samples][row][8 halves]) so that a CTA fetches a stage with one
// bulk copy.  Same integer phase arithmetic as the in-kernel path: both give bit-identical operands.
// grid = (stages, 2 types), block = 128 (row).
__global__ void __launch_bounds__(128)
gls_umma_fine_kernel(const double2* __restrict__ rec1, long long n, unsigned char* __restrict__ img, long long stages) {
  const long long stg = blockIdx.x;
  const unsigned kfine = (blockIdx.y + 1u) * threadIdx.x;
  unsigned char* out = img + ((long long)blockIdx.y * stages + stg) * 16384 + threadIdx.x * 16;
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = stg * UM_STAGE_SAMPLES + q * 4 + u;
      unsigned long long b64 = 0ull;
      if (i < n) {
        const double b = rec1[i].y;
        b64 = __double2ull_rn((b - floor(b)) * 18446744073709551616.0);
      }
      float c, s;
      um_sincos_fx(kfine * (unsigned)(b64 >> 32) + __umulhi(kfine, (unsigned)b64), c, s);
      um_split2(c, s, hi[u], lo[u]);
    }
    *reinterpret_cast<uint4*>(out + q * UM_LBO_FINE) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(out + 8192 + q * UM_LBO_FINE) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}


// A worker warp
//   * PRODUCES: per stage of 16 samples one fine task (row = 32 (warp & 3) + lane, sample quad = warp >> 2: four phases)
//     and one coarse task of the same quad;
//   * DRAINS: one accumulation run behind the producers it adds its 32 lanes x 64 columns of the finished TMEM
//     accumulator to 64 FP32 masters in registers (TMEM lanes 32 (warp & 3).., columns 64 (warp >> 2)..);
//   * FLUSHES the masters at the end of the job.
template <bool FINE_PRE>
__global__ void __launch_bounds__(UM_THREADS, 1)
gls_umma_kernel(const GlsUmmaArgs a) {
  extern __shared__ __align__(1024) unsigned char um_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long t_start = clock64();

  // ---- job ----
  // tile fastest: the CTAs that run together then work on the same sample split and share its fine-operand images in L2
  int job = blockIdx.x;
  const int ntile = a.nt1 + a.nt2;
  const int tile = job % ntile;
  job /= ntile;
  const int split = job % a.nsplit;
  const int curve = job / a.nsplit;
  const bool type2 = tile >= a.nt1;
  const int cpt = type2 ? a.cpt2 : a.cpt1;                       // coarse blocks per tile (padded to 4 / 8)
  const int cb0 = (type2 ? tile - a.nt1 : tile) * cpt;           // first coarse block of this tile
  const int ncb = min(cpt, a.nC - cb0);                          // real coarse blocks (> 0 by construction)
  const int N = cpt * (type2 ? 2 : 4);                           // MMA N: multiple of 16, <= 256
  const int CS = a.chunk_stages;

  const GlsCurve* cvp = a.curves + curve;
  const long long cbegin = cvp->begin, cn = cvp->n;
  const long long per = (((cn + a.nsplit - 1) / a.nsplit) + 63) & ~63LL;   // whole stages (the fine images are per stage)
  const long long sb = (long long)split * per;
  const long long se = sb + per < cn ? sb + per : cn;
  const long long ns = se > sb ? se - sb : 0;
  const int nchunks = (int)((ns + CS * UM_STAGE_SAMPLES - 1) / (CS * UM_STAGE_SAMPLES));
  const int nstages = nchunks * CS;                              // padded with zero-weight samples
  if (nchunks == 0) return;                                      // block-uniform: nothing to add

  // ---- shared memory ----
  const uint32_t smem0 = smem_u32(um_smem);
  unsigned char* recs = um_smem + UM_NSTAGES * UM_STAGE_BYTES;
  // per-sample records, two buffers of UM_BLOCK samples.  Phases are 2^-32-turn fixed point inside the loop: the per-index
  // step b_i as a 64-bit fraction (so that (integer index) * b_i keeps 2^-32 turn after 15 bits of index), the phase at
  // the tile's base frequency as a 32-bit fraction -- no FP64 arithmetic per stage.
  unsigned long long* s_b64 = reinterpret_cast<unsigned long long*>(recs);     // [2][UM_BLOCK] frac(b_i) * 2^64
  unsigned* s_A32 = reinterpret_cast<unsigned*>(s_b64 + 2 * UM_BLOCK);         // [2][UM_BLOCK] frac(kmul A_i) * 2^32
  float* s_wy = reinterpret_cast<float*>(s_A32 + 2 * UM_BLOCK);                // [2][UM_BLOCK] w' y'
  float* s_w = s_wy + 2 * UM_BLOCK;                                            // [2][UM_BLOCK] w' (0 for padding samples)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + 2 * UM_BLOCK);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * UM_NSTAGES;
  const uint32_t bar_tfull = bar_empty + 8 * UM_NSTAGES, bar_tempty = bar_tfull + 16;
  const uint32_t bar_rec = bar_tempty + 16;   // [2] record buffer `b` staged by all 16 warps
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * UM_NSTAGES + 6);
  volatile int* s_abort = reinterpret_cast<volatile int*>(s_tmem + 1);

  if (tid == 0) {
    for (int s = 0; s < UM_NSTAGES; ++s) {
      mbar_init(bar_full + 8 * s, UM_WORKERS / 32 + (FINE_PRE ? 1 : 0));   // one arrival per worker warp (+ the bulk copy's expect_tx)
      mbar_init(bar_empty + 8 * s, 1);                // tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);                // tcgen05.commit
      mbar_init(bar_tempty + 8 * s, UM_WORKERS / 32); // one arrival per worker warp
      mbar_init(bar_rec + 8 * s, UM_WORKERS / 32);
    }
    mbar_init_fence();
    *s_abort = 0;
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(s_tmem), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  long long t_loop0 = 0, t_loop1 = 0;

  const uint32_t idesc = idesc_f16_f32(UM_FINE, N);
  if (warp >= 16) {
    // =====================================================================================================
    // MMA warp.  Issuing a tcgen05.mma blocks the issuing thread while the tensor core is busy (measured: the 12
    // instructions of a pair of stages take ~1470 clocks to issue), so the issuer must have nothing else to do.
    // =====================================================================================================
    setmaxnreg_dec<UM_MMA_REGS>();
    if (FINE_PRE && warp == 17) {
      // copy warp: the fine operand of both stages of a pair comes as two 16 KB bulk copies from the per-curve images
      bool ok = true;
      const unsigned char* src = a.fine_img + ((type2 ? a.fine_stages : 0) + (cbegin + sb) / UM_STAGE_SAMPLES) * 16384;
      for (int g = 0; g < nstages && ok; g += 2) {
        const int pslot = (g >> 1) & 1;
        ok = um_wait(bar_empty + 8 * pslot, ((g >> 2) & 1) ^ 1, s_abort, t_start);
        if (!ok) break;
        if (elect_one()) {
          const uint32_t dst = smem0 + pslot * 2 * UM_STAGE_BYTES + UM_FINE_HI;
          um_arrive_expect_tx(bar_full + 8 * pslot, 32768);
          um_bulk_g2s(dst, src + (long long)g * 16384, 16384, bar_full + 8 * pslot);
          um_bulk_g2s(dst + UM_STAGE_BYTES, src + (long long)(g + 1) * 16384, 16384, bar_full + 8 * pslot);
        }
        __syncwarp();
      }
    }
    if (warp == 16) {
      bool ok = true;
      int cpos = 0, ch = 0;
      long long* trace = (a.prof && (a.dbg & 16) && blockIdx.x == 0) ? a.prof + 4LL * gridDim.x : nullptr;
      for (int g = 0; g < nstages && ok; g += 2) {
        const int pslot = (g >> 1) & 1, acc = ch & 1;     // pair slot = stages (2 pslot, 2 pslot + 1) of the ring
        if (trace && lane == 0 && g < 2 * 1024) trace[(g >> 1) * 8 + 2] = clock64();
        if (cpos == 0) ok = um_wait(bar_tempty + 8 * acc, ((ch >> 1) & 1) ^ 1, s_abort, t_start);
        if (ok) ok = um_wait(bar_full + 8 * pslot, (g >> 2) & 1, s_abort, t_start);
        if (!ok) break;
        if (trace && lane == 0 && g < 2 * 1024) trace[(g >> 1) * 8 + 3] = clock64();
        tc_fence_after();
        const uint32_t sbase = smem0 + pslot * 2 * UM_STAGE_BYTES;
        const uint32_t d = tmem + acc * 256;
        const uint64_t ah = smem_desc(sbase + UM_FINE_HI, UM_LBO_FINE, UM_SBO);
        const uint64_t al = smem_desc(sbase + UM_FINE_LO, UM_LBO_FINE, UM_SBO);
        const uint64_t bh = smem_desc(sbase + UM_COARSE_HI, UM_LBO_COARSE, UM_SBO);
        const uint64_t bl = smem_desc(sbase + UM_COARSE_LO, UM_LBO_COARSE, UM_SBO);
        // start-address field steps: second K-step of a stage, second stage of the pair
        constexpr uint64_t KA = (2 * UM_LBO_FINE) >> 4, KB = (2 * UM_LBO_COARSE) >> 4, ST = UM_STAGE_BYTES >> 4;
        const bool last = cpos + 2 == CS;
        if (a.dbg & 1) {
          if (elect_one()) {
            mbar_arrive(bar_empty + 8 * pslot);
            if (last) mbar_arrive(bar_tfull + 8 * acc);
          }
        } else if (elect_one()) {
          // per K-step of eight samples: hi lo + lo hi + hi hi
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint64_t oa = (q >> 1) * ST + (q & 1) * KA, ob = (q >> 1) * ST + (q & 1) * KB;
            mma_f16_ss(d, al + oa, bh + ob, idesc, (cpos != 0 || q != 0) ? 1u : 0u);
            mma_f16_ss(d, ah + oa, bl + ob, idesc, 1);
            mma_f16_ss(d, ah + oa, bh + ob, idesc, 1);
          }
          mma_commit(bar_empty + 8 * pslot);
          if (last) mma_commit(bar_tfull + 8 * acc);
        }
        __syncwarp();
        if (trace && lane == 0 && g < 2 * 1024) trace[(g >> 1) * 8 + 4] = clock64();
        cpos += 2;
        if (cpos == CS) { cpos = 0; ++ch; }
      }
    }
  } else {
    setmaxnreg_inc<UM_WORKER_REGS>();
    const int p = tid;
    const int wq = warp & 3, quad = warp >> 2;
    const int row = wq * 32 + lane;
    const bool weighted_tt = a.weighted && cvp->three_term;
    const int yslot = rec_slot(REC_Y), wslot = rec_slot(REC_W);
    const unsigned kmul = type2 ? 2u : 1u;
    // tile base frequency and its gauge (the per-index phase origin gamma of the records, see gls.cu)
    const long long jT = a.j0 + (long long)cb0 * UM_FINE;
    const double fT = cvp->fmin + (double)jT * cvp->df;
    double gT = (double)jT * cvp->gamma;
    gT -= floor(gT);
    const unsigned kfine = kmul * (unsigned)row;                  // phase of the fine operand = kfine * b_i
    // coarse task: type 1 (coarse block, plain | y-weighted pair of rows), type 2 (coarse block)
    const int ccb = type2 ? row : wq * 16 + (lane & 15);
    const int yy = type2 ? 0 : lane >> 4;
    const unsigned kcoarse = kmul * (unsigned)(ccb * UM_FINE);    // phase of the coarse operand = A_i + kcoarse * b_i
    const bool cactive = ccb < cpt;
    const float* s_wsel = yy ? s_wy : s_w;
    // byte offsets of this thread's operand rows inside a stage
    const uint32_t fine_off = quad * UM_LBO_FINE + row * 16;
    const uint32_t rowc_off = quad * UM_LBO_COARSE + (uint32_t)((type2 ? 0 : 2 * yy) * cpt + ccb) * 16;
    const uint32_t rows_off = rowc_off + (uint32_t)cpt * 16;

    float m[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) m[c] = 0.f;
    const uint32_t tlane = tmem + ((uint32_t)(wq * 32) << 16) + quad * 64;
    bool ok = true;

    // compensation factors of a full run and of the job's last run: 1 + loss per instruction * (instructions of the run
    // that added real samples; zero padding adds exact zeros, which lose nothing)
    const float comp_full = 1.0f + a.rz_comp * (float)(6 * CS);
    const float comp_last = 1.0f + a.rz_comp * (float)(3 * (int)((ns - (long long)(nchunks - 1) * CS * UM_STAGE_SAMPLES + 7) >> 3));
    auto drain = [&](int ch) {
      const int acc = ch & 1;
      const float comp = ch == nchunks - 1 ? comp_last : comp_full;
      ok = um_wait(bar_tfull + 8 * acc, (ch >> 1) & 1, s_abort, t_start);
      tc_fence_after();
      if (ok && !(a.dbg & 4)) {
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          if (quad * 64 + c0 < N) {   // warp-uniform (N is a multiple of 16: a partial group reads columns nobody flushes)
            uint32_t r[32];
            tmem_ld32(tlane + acc * 256 + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 32; ++u) m[c0 + u] = fmaf(__uint_as_float(r[u]), comp, m[c0 + u]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
    };
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
        const double bf = r1.y - floor(r1.y);                                        // [0, 1)
        const double A = (double)kmul * (frac_of_product(fT, r1.x) + gT);            // |A| < 4 turns
        s_b64[o] = __double2ull_rn(bf * 18446744073709551616.0);                     // 2^64: exact scaling, < 2^64
        s_A32[o] = (unsigned)__double2ll_rn(A * 4294967296.0);                       // wraps mod 1 turn
        s_wy[o] = weighted_tt ? wv * yv : yv;      // three-term records carry sqrt(w'), sqrt(w') y'
        s_w[o] = weighted_tt ? wv * wv : wv;
      } else {
        s_b64[o] = 0ull;
        s_A32[o] = 0u;
        s_wy[o] = 0.f;
        s_w[o] = 0.f;
      }
    };
    // one stage (16 samples) of this thread's tasks: fine row `row` and coarse block `ccb`, sample quad `quad`;
    // packed (c, s) pairs of four samples, fp16 hi and lo
    auto compute_fine = [&](int ro, uint32_t (&hi)[4], uint32_t (&lo)[4]) {
      const uint4 b01 = *reinterpret_cast<const uint4*>(s_b64 + ro), b23 = *reinterpret_cast<const uint4*>(s_b64 + ro + 2);
      const unsigned blo[4] = {b01.x, b01.z, b23.x, b23.z}, bhi[4] = {b01.y, b01.w, b23.y, b23.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float c, s;
        um_sincos_fx(kfine * bhi[u] + __umulhi(kfine, blo[u]), c, s);
        um_split2(c, s, hi[u], lo[u]);
      }
    };
    auto compute_coarse = [&](int ro, uint32_t (&ph)[4], uint32_t (&pl)[4]) {
      const uint4 b01 = *reinterpret_cast<const uint4*>(s_b64 + ro), b23 = *reinterpret_cast<const uint4*>(s_b64 + ro + 2);
      const unsigned blo[4] = {b01.x, b01.z, b23.x, b23.z}, bhi[4] = {b01.y, b01.w, b23.y, b23.w};
      const uint4 a4 = *reinterpret_cast<const uint4*>(s_A32 + ro);
      const unsigned aq[4] = {a4.x, a4.y, a4.z, a4.w};
      const float4 wq4 = *reinterpret_cast<const float4*>(s_wsel + ro);
      const float wv[4] = {wq4.x, wq4.y, wq4.z, wq4.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float c, s;
        um_sincos_fx(aq[u] + kcoarse * bhi[u] + __umulhi(kcoarse, blo[u]), c, s);
        um_split2(wv[u] * c, wv[u] * s, ph[u], pl[u]);
      }
    };
    auto store_fine = [&](uint32_t sbase, const uint32_t (&hi)[4], const uint32_t (&lo)[4]) {
      um_sts128(sbase + UM_FINE_HI + fine_off, hi);
      um_sts128(sbase + UM_FINE_LO + fine_off, lo);
    };
    auto store_coarse = [&](uint32_t sbase, const uint32_t (&ph)[4], const uint32_t (&pl)[4]) {
      if (cactive) {
        uint32_t rw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) rw[u] = ph[u] ^ 0x80000000u;             // (w c, -w s)
        um_sts128(sbase + UM_COARSE_HI + rowc_off, rw);
#pragma unroll
        for (int u = 0; u < 4; ++u) rw[u] = pl[u] ^ 0x80000000u;
        um_sts128(sbase + UM_COARSE_LO + rowc_off, rw);
#pragma unroll
        for (int u = 0; u < 4; ++u) rw[u] = __byte_perm(ph[u], 0, 0x1032);   // (w s, w c)
        um_sts128(sbase + UM_COARSE_HI + rows_off, rw);
#pragma unroll
        for (int u = 0; u < 4; ++u) rw[u] = __byte_perm(pl[u], 0, 0x1032);
        um_sts128(sbase + UM_COARSE_LO + rows_off, rw);
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
    // (mbarriers instead of bar.sync: every wait in this kernel must be able to time out)
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_rec);
    ok = um_wait(bar_rec, 0, s_abort, t_start);
    t_loop0 = clock64();
    // optional trace of block 0 (dbg & 16): per pair 8 clock stamps behind the per-job records
    long long* trace = (a.prof && (a.dbg & 16) && blockIdx.x == 0) ? a.prof + 4LL * gridDim.x : nullptr;
    int g = 0;        // global stage counter; stages are processed in pairs (nstages is even: CS is)
    int cpos = 0;     // position of stage g inside its accumulation run
    int ch = 0;       // accumulation run of stage g
    for (int blk = 0; blk < nblk && ok; ++blk) {
      double2 n1 = make_double2(0.0, 0.0);
      float4 n2 = make_float4(0.f, 0.f, 0.f, 0.f);
      bool nin = false;
      const bool more = blk + 1 < nblk;
      if (more) load_block(blk + 1, n1, n2, nin);
      const int rbase = (blk & 1) * UM_BLOCK + quad * 4;
      for (int st = 0; st < UM_BLOCK / UM_STAGE_SAMPLES && g < nstages; st += 2, g += 2) {
        if (st == 8 && more) {
          // stage the next block of records early.  Buffer (blk + 1) & 1 was last read for the last pair of block blk - 1; this
          // warp is four pairs into block blk and could only store those pairs after the tensor core had consumed the pair
          // two (three in the 2-CTA kernel) before each, i.e. after EVERY warp had written that one: nobody is still in block
          // blk - 1.  (Ordering by mbarriers only: compute-sanitizer's racecheck, which models bar.sync, reports these buffers.)
          // The wait for THIS arrival is at the end of the block.
          store_block(blk + 1, n1, n2, nin);
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_rec + 8 * ((blk + 1) & 1));
        }
        // operands go to registers before the wait for the pair slot (the wait overlaps the arithmetic): both stages'
        // coarse rows when the fine operand comes by bulk copy, else the first stage (more would not fit beside the 64
        // masters)
        const int pslot = (g >> 1) & 1;
        const int ro0 = rbase + st * UM_STAGE_SAMPLES, ro1 = ro0 + UM_STAGE_SAMPLES;
        const uint32_t sbase = smem0 + pslot * 2 * UM_STAGE_BYTES;
        uint32_t h0[4], l0[4], p0[4], q0[4];
        if (!(a.dbg & 2)) {
          if (FINE_PRE) {
            compute_coarse(ro0, h0, l0);
            compute_coarse(ro1, p0, q0);
          } else {
            compute_fine(ro0, h0, l0);
            compute_coarse(ro0, p0, q0);
          }
        }
        ok = um_wait(bar_empty + 8 * pslot, ((g >> 2) & 1) ^ 1, s_abort, t_start);
        if (!ok) break;
        if (trace && warp == 3 && lane == 0 && g < 2 * 1024) trace[(g >> 1) * 8 + 0] = clock64();
        if (!(a.dbg & 2)) {
          if (FINE_PRE) {
            store_coarse(sbase, h0, l0);
            store_coarse(sbase + UM_STAGE_BYTES, p0, q0);
          } else {
            store_fine(sbase, h0, l0);
            store_coarse(sbase, p0, q0);
            compute_fine(ro1, h0, l0);
            store_fine(sbase + UM_STAGE_BYTES, h0, l0);
            compute_coarse(ro1, p0, q0);
            store_coarse(sbase + UM_STAGE_BYTES, p0, q0);
          }
        }
        if (!(a.dbg & 8)) fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * pslot);
        if (trace && warp == 3 && lane == 0 && g < 2 * 1024) trace[(g >> 1) * 8 + 1] = clock64();
        cpos += 2;
        if (cpos == CS) {
          // the run `ch` has all its stages; one run behind, add the finished accumulator of run ch - 1 to the masters
          // (its last instruction was issued CS stages ago)
          if (ch >= 1) {
            if (trace && warp == 3 && lane == 0 && g < 2 * 1024) trace[(g >> 1) * 8 + 6] = clock64();
            drain(ch - 1);
            if (!ok) break;
            if (trace && warp == 3 && lane == 0 && g < 2 * 1024) trace[(g >> 1) * 8 + 7] = clock64();
          }
          cpos = 0;
          ++ch;
        }
      }
      if (more && ok) ok = um_wait(bar_rec + 8 * ((blk + 1) & 1), ((blk + 1) >> 1) & 1, s_abort, t_start);
    }
    if (ok) drain(nchunks - 1);
    t_loop1 = clock64();
    if (ok) {
      // flush: column n = 64 quad + c  ->  (sum s, coarse block cb) = (n / cpt, n % cpt); row = fine index
      int s = (quad * 64) / cpt, cb = (quad * 64) % cpt;
      const int plane0 = type2 ? 4 : 0;
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        const long long j = (long long)(cb0 + cb) * UM_FINE + row;
        if (quad * 64 + c < N && cb < ncb && j < a.nf) {
          unsigned long long* pp = a.partial + (long long)(plane0 + s) * a.nf_tot + (long long)curve * a.nf + j;
          atomicAdd(pp, (unsigned long long)__float2ll_rn(m[c] * a.fix_scale));
        }
        if (++cb == cpt) { cb = 0; ++s; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
  if (tid == 0) {
    if (blockIdx.x == 0) *a.status_next = 0;
    if (*s_abort) *a.status = 1;
    if (a.prof) {
      long long* pr = a.prof + 4LL * blockIdx.x;
      pr[0] = t_start;
      pr[1] = t_loop0;
      pr[2] = t_loop1;
      pr[3] = clock64();
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// work decomposition (pure host arithmetic)
// ---------------------------------------------------------------------------------------------------------------
void gls_umma_plan(int sm_count, long long B, long long nf, long long nmax, const GlsUmmaKnobs& k, GlsUmmaPlan* out) {
  GlsUmmaPlan p;
  // One long curve from 16384 frequencies on (at least one full tile of 64 coarse blocks; knob cg2 = 1: from 4096, 0: never):
  // a pair of CTAs per tile of 256 fine indices (gls_umma2.cu), fine operand precomputed at 64 KB per 16 samples.  Beyond
  // UM_MAX_FINE_BYTES of scratch the fine operand is computed in the kernel instead (one-CTA kernel).
  const bool pair = B == 1 && k.fine != 0 && k.cg2 != 0 && nf >= (k.cg2 > 0 ? 4096 : 16384) &&
                    (nmax / UM_STAGE_SAMPLES + 64) * 65536 <= UM_MAX_FINE_BYTES;
  p.fine = pair ? 256 : UM_FINE;
  p.nC = (int)((nf + p.fine - 1) / p.fine);
  // tiles of (nearly) equal size; the number of coarse blocks per tile is rounded so that the MMA's N (4 cpt1, 2 cpt2) is a
  // multiple of 16 -- and, in the pair kernel, the type-2 half per CTA (cpt2 / 2) a multiple of 8
  const int r2 = pair ? 16 : 8;
  p.nt1 = (p.nC + UM_MAX_T1 - 1) / UM_MAX_T1;
  p.cpt1 = (((p.nC + p.nt1 - 1) / p.nt1) + 3) & ~3;
  p.nt2 = (p.nC + UM_MAX_T2 - 1) / UM_MAX_T2;
  p.cpt2 = (((p.nC + p.nt2 - 1) / p.nt2) + r2 - 1) & ~(r2 - 1);
  // rounding up the tile size can leave the last tile(s) of a type empty: drop them
  while (p.nt1 > 1 && (long long)(p.nt1 - 1) * p.cpt1 >= p.nC) --p.nt1;
  while (p.nt2 > 1 && (long long)(p.nt2 - 1) * p.cpt2 >= p.nC) --p.nt2;

  // Sample splits.  (i) A job keeps its 32768 sums in FP32 registers until its end: at most UM_MAX_JOB_SAMPLES samples per
  // job bound the rounding of those masters (64 additions of 256-sample runs: ~2e-7 of their magnitude; C5 with one job
  // per tile showed 4.7e-6 on weak bins).  (ii) Few tiles: fill whole waves of one CTA (pair) per SM (pair), with >= 1024
  // samples per job (its set-up and its flush of 32768 REDs cost about as much as 300 samples).
  const long long base_jobs = B * (p.nt1 + p.nt2);
  const long long slots = pair ? sm_count / 2 : sm_count;
  const long long smin = (nmax + UM_MAX_JOB_SAMPLES - 1) / UM_MAX_JOB_SAMPLES;
  int nsplit = (int)smin;
  if (k.nsplit > 0) nsplit = k.nsplit;
  else if (base_jobs * smin < 6LL * slots) {
    long long cap = nmax / 1024;
    if (cap < smin) cap = smin;
    double best = 1e300;
    for (long long s = smin; s <= cap && s <= 4096; ++s) {
      const long long jobs = base_jobs * s;
      const long long waves = (jobs + slots - 1) / slots;
      const double per = (double)((nmax + s - 1) / s) + 300.0;   // per-job fixed cost (set-up + flush) in samples
      const double cost = (double)waves * per;
      if (cost < best * 0.999) { best = cost; nsplit = (int)s; }
      if (jobs > 16LL * slots) break;
    }
  }
  p.nsplit = nsplit;
  // stages (16 samples) per accumulation run in TMEM; even, because stages go in pairs.  Longer runs mean fewer drains (C2,
  // one-CTA kernel: 0.48 ms at 4, 0.43 ms at 16; pair kernel with 5,000-sample jobs: 0.389 ms at 8, 0.367 ms at 16); short
  // jobs keep short runs (see rz_comp in gls_umma_launch).
  const long long per_job = (nmax + nsplit - 1) / nsplit;
  int cs = per_job >= (pair ? 4096 : 8192) ? 16 : (per_job >= 2048 ? 8 : 4);
  if (k.chunk > 0) cs = (k.chunk + 1) & ~1;
  p.chunk_stages = cs;
  p.jobs = base_jobs * nsplit;
  // One curve with several coarse tiles per sample split: the fine operand is the same for all of them, compute it once
  // (32 KB per 16 samples on the one-CTA kernel: 133 MB for C2, 2 GB for C5; twice that on the pair kernel)
  bool fine_pre = pair || (B == 1 && p.nt1 >= 3 && k.fine != 0);
  if (k.fine == 1 && B == 1) fine_pre = true;
  if (!pair && (nmax / UM_STAGE_SAMPLES + 64) * 32768 > UM_MAX_FINE_BYTES) fine_pre = false;
  const long long per = ((((long long)nmax + nsplit - 1) / nsplit) + 63) & ~63LL;      // samples per split: whole stages
  const long long stages = per * nsplit / UM_STAGE_SAMPLES + 16;                       // the last run of a split may read past it
  p.fine_bytes = fine_pre ? stages * (pair ? 65536 : 32768) : 0;
  p.path = pair ? 3 : (fine_pre ? 2 : 1);
  *out = p;
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
int gls_umma2_launch(pdc_ctx* ctx, GlsUmmaArgs a, const GlsUmmaPlan& plan, long long nmax, cudaStream_t st);   // gls_umma2.cu

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
  GlsUmmaPlan plan;
  gls_umma_plan(ctx->sm_count, B, nf, nmax, GlsUmmaKnobs{ctx->gls_umma_fine, ctx->gls_umma_cg2, ctx->gls_umma_nsplit, ctx->gls_umma_chunk},
                &plan);
  a.nC = plan.nC;
  a.nt1 = plan.nt1;
  a.cpt1 = plan.cpt1;
  a.nt2 = plan.nt2;
  a.cpt2 = plan.cpt2;
  a.nsplit = plan.nsplit;
  a.chunk_stages = plan.chunk_stages;
  a.weighted = weighted ? 1 : 0;
  // The TMEM accumulator truncates toward zero after every instruction: an expected loss of 0.5 ulp(acc) = 0.5 * ln 2 *
  // 2^-23 |acc| per instruction.  Over a run of n instructions with |acc| growing about linearly that is a relative loss
  // of 2.07e-8 * n of the run's sum (measured on C2: the power error grows from 1.0e-6 at 24 instructions to 1.9e-6 at 48
  // and 4.0e-6 at 96 without the correction, and stays at 1-3e-7 with it).  The drain multiplies the run's sum back by
  // 1 + 2.07e-8 * (instructions of the run that added real samples).  The correction is exact in expectation for a sum
  // that grows steadily (the peaks); for a partial sum that oscillates inside a run it is not, which shows as a relative
  // error of up to ~1e-8 * n on weak bins of SHORT curves -- hence the shorter runs chosen for them in gls_umma_launch.
  a.rz_comp = ctx->gls_umma_rzcomp ? 2.07e-8f : 0.0f;
  a.fix_scale = fix_scale;
  a.prof = nullptr;
  a.dbg = ctx->gls_umma_dbg;
  PDC_TRY(ctx->umma_status.reserve(2 * sizeof(int)));
  a.status = ctx->umma_status.as<int>() + (ctx->umma_calls & 1);
  a.status_next = ctx->umma_status.as<int>() + ((ctx->umma_calls + 1) & 1);
  ctx->umma_status_cur = a.status;
  ctx->umma_calls++;

  if (plan.path == 3) return gls_umma2_launch(ctx, a, plan, nmax, st);   // a pair of CTAs per tile (gls_umma2.cu)
  const long long jobs = plan.jobs;
  const int nsplit = plan.nsplit;
  if (jobs > 0x7fffffffLL) { set_error("pdc_gls: problem too large for one call (%lld jobs)", jobs); return PDC_EINVAL; }

  static bool attr_set[64] = {};
  if (ctx->device >= 0 && ctx->device < 64 && !attr_set[ctx->device]) {
    PDC_CUDA(cudaFuncSetAttribute(gls_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UM_SMEM_BYTES));
    PDC_CUDA(cudaFuncSetAttribute(gls_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UM_SMEM_BYTES));
    attr_set[ctx->device] = true;
  }
  const bool fine_pre = plan.path == 2;
  a.fine_img = nullptr;
  a.fine_stages = 0;
  if (fine_pre) {
    const long long per = ((((long long)nmax + nsplit - 1) / nsplit) + 63) & ~63LL;
    const long long stages = per * nsplit / UM_STAGE_SAMPLES + 16;   // the last accumulation run of a split may read past it
    PDC_TRY(ctx->umma_fine.reserve((size_t)stages * 2 * 16384));
    a.fine_img = ctx->umma_fine.as<unsigned char>();
    a.fine_stages = stages;
    dim3 grid((unsigned)stages, 2);
    gls_umma_fine_kernel<<<grid, 128, 0, st>>>(rec1, (long long)nmax, ctx->umma_fine.as<unsigned char>(), stages);
    PDC_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  if (!ctx->umma_status_clean) {
    PDC_CUDA(cudaMemsetAsync(ctx->umma_status.p, 0, 2 * sizeof(int), st));
    ctx->umma_status_clean = true;
  }
  if (ctx->umma_prof_on) {
    PDC_TRY(ctx->umma_prof.reserve(sizeof(long long) * (4 * (size_t)jobs + 8 * 1024)));
    PDC_CUDA(cudaMemsetAsync(ctx->umma_prof.p, 0, sizeof(long long) * (4 * (size_t)jobs + 8 * 1024), st));
    a.prof = ctx->umma_prof.as<long long>();
    ctx->umma_prof_jobs = jobs;
  }
  PDC_TRY(ctx->main_begin(st));
  if (fine_pre) gls_umma_kernel<true><<<(unsigned)jobs, UM_THREADS, UM_SMEM_BYTES, st>>>(a);
  else gls_umma_kernel<false><<<(unsigned)jobs, UM_THREADS, UM_SMEM_BYTES, st>>>(a);
  PDC_CUDA(cudaGetLastError());
  PDC_TRY(ctx->main_end(st));
  ctx->launches++;
  ctx->last_gls_path = fine_pre ? 2 : 1;
  return PDC_OK;
}

}  // namespace pdc
