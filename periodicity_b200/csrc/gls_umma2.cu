// GLS trial-frequency sums on the tensor cores, two CTAs per tile (tcgen05 `cta_group::2`).
//
// Same formulation and numerics as gls_umma_kernel (gls_umma.cu) for ONE LONG CURVE (fine operand precomputed per call),
// but a tile is 256 fine indices x up to 256 coarse rows worked by a PAIR of CTAs on the two SMs of a TPC:
//   * every CTA holds 128 fine rows (its half of the MMA's M = 256) and the accumulator rows that belong to them;
//   * the coarse operand (the MMA's N rows) is split between the two CTAs' shared memories -- each CTA computes only HALF
//     of it, which halves the work of the worker warps (they were the bottleneck of the one-CTA kernel: 8 sincos chains per
//     thread and pair of stages, now 4) and the B-operand traffic per SM; a stage is 32 KB instead of 48, so the ring holds
//     three pairs of stages instead of two;
//   * the leader CTA's warp 16 issues `tcgen05.mma.cta_group::2`; its commits are multicast to both CTAs' barriers; the
//     peer CTA's warp 16 forwards "my half of the pair is written" / "my accumulator is drained" to the leader with
//     remote mbarrier arrivals.
// Grid index j = 256 cb + 128 rank + k.  Type-1 tiles: coarse rows {C, S} in the leader, {YC, YS} in the peer (the
// one-CTA column layout n = s * cpt + cb, whose halves are exactly that); type-2 tiles: the coarse blocks are split, each
// CTA holds {C2, S2} of its half (n = h * cpt + s * cpt/2 + cbl).
// All waits are bounded; a protocol error ends both CTAs with a status word.
#include "gls_umma_common.cuh"

namespace pdc {

constexpr int U2_NSTAGES = 6;                                  // three pairs of 16-sample stages
constexpr uint32_t U2_FINE_HI = 0, U2_FINE_LO = 8192, U2_COARSE_HI = 16384, U2_COARSE_LO = 24576;
constexpr uint32_t U2_STAGE_BYTES = 32768;
constexpr uint32_t U2_LBO = 128 * 16;                          // both operands are 128 rows per CTA
constexpr uint32_t U2_SMEM_BYTES = U2_NSTAGES * U2_STAGE_BYTES + UM_REC_BYTES + 256;
constexpr int U2_CL_FINE = 256;                                // fine indices per cluster

// ---- cluster / cta_group::2 PTX ----
__device__ __forceinline__ uint32_t u2_cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void u2_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void u2_remote_arrive(uint32_t local_bar, uint32_t target_rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(local_bar), "r"(target_rank) : "memory");
}
__device__ __forceinline__ bool u2_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool u2_wait_cluster(uint32_t bar, uint32_t parity, volatile int* s_abort, long long t_start) {
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 256; ++i)
      if (u2_try_wait_cluster(bar, parity)) return true;
    if (*s_abort || clock64() - t_start > UM_WAIT_CLOCKS) {
      *s_abort = 1;
      return false;
    }
  }
}
__device__ __forceinline__ void u2_tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void u2_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void u2_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// completion of all tcgen05.mma issued so far by this thread -> one arrival on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void u2_commit_both(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((unsigned short)3) : "memory");
}
__device__ __forceinline__ void u2_sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void u2_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }

// fine operand images for the pair kernel: [type][stage][rank][hi 8 KB | lo 8 KB]; grid = (stages, 2 types, 2 ranks)
__global__ void __launch_bounds__(128)
gls_umma2_fine_kernel(const double2* __restrict__ rec1, long long n, unsigned char* __restrict__ img, long long stages) {
  const long long stg = blockIdx.x;
  const unsigned kfine = (blockIdx.y + 1u) * (blockIdx.z * 128u + threadIdx.x);
  unsigned char* out = img + (((long long)blockIdx.y * stages + stg) * 2 + blockIdx.z) * 16384 + threadIdx.x * 16;
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
    *reinterpret_cast<uint4*>(out + q * U2_LBO) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(out + 8192 + q * U2_LBO) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(UM_THREADS, 1)
gls_umma2_kernel(const GlsUmmaArgs a) {
  extern __shared__ __align__(1024) unsigned char um_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long t_start = clock64();
  const uint32_t rank = u2_cta_rank();

  // ---- job (one per cluster; tile fastest) ----
  int job = blockIdx.x >> 1;
  const int ntile = a.nt1 + a.nt2;
  const int tile = job % ntile;
  const int split = job / ntile;
  const bool type2 = tile >= a.nt1;
  const int cpt = type2 ? a.cpt2 : a.cpt1;                       // coarse blocks (of 256 frequencies) per tile
  const int cb0 = (type2 ? tile - a.nt1 : tile) * cpt;
  const int ncb = min(cpt, a.nC - cb0);
  const int N = cpt * (type2 ? 2 : 4);                           // MMA N (both CTAs together): multiple of 16, <= 256
  const int cptl = type2 ? cpt / 2 : cpt;                        // coarse blocks whose rows THIS CTA computes
  const int CS = a.chunk_stages;

  const GlsCurve* cvp = a.curves;
  const long long cn = cvp->n;
  const long long per = (((cn + a.nsplit - 1) / a.nsplit) + 63) & ~63LL;
  const long long sb = (long long)split * per;
  const long long se = sb + per < cn ? sb + per : cn;
  const long long ns = se > sb ? se - sb : 0;
  const int nchunks = (int)((ns + CS * UM_STAGE_SAMPLES - 1) / (CS * UM_STAGE_SAMPLES));
  const int nstages = nchunks * CS;
  if (nchunks == 0) return;                                      // cluster-uniform

  // ---- shared memory (same layout in both CTAs: remote arrivals and multicast commits address it by offset) ----
  const uint32_t smem0 = smem_u32(um_smem);
  unsigned char* recs = um_smem + U2_NSTAGES * U2_STAGE_BYTES;
  unsigned long long* s_b64 = reinterpret_cast<unsigned long long*>(recs);
  unsigned* s_A32 = reinterpret_cast<unsigned*>(s_b64 + 2 * UM_BLOCK);
  float* s_wy = reinterpret_cast<float*>(s_A32 + 2 * UM_BLOCK);
  float* s_w = s_wy + 2 * UM_BLOCK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + 2 * UM_BLOCK);
  const uint32_t bar_full = smem_u32(bars);          // [3] this CTA's half of a pair of stages is written (16 warps + bulk copy)
  const uint32_t bar_empty = bar_full + 24;          // [3] the pair's instructions have completed (multicast commit)
  const uint32_t bar_tfull = bar_empty + 24;         // [2] accumulator ready (multicast commit)
  const uint32_t bar_tempty = bar_tfull + 16;        // [2] this CTA's 16 warps have drained the accumulator
  const uint32_t bar_rec = bar_tempty + 16;          // [2] record buffer staged
  const uint32_t bar_pfull = bar_rec + 16;           // [3] leader only: the peer's half is written (forwarded)
  const uint32_t bar_ptempty = bar_pfull + 24;       // [2] leader only: the peer has drained (forwarded)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 17);
  volatile int* s_abort = reinterpret_cast<volatile int*>(s_tmem + 1);

  if (tid == 0) {
    for (int s = 0; s < 3; ++s) {
      mbar_init(bar_full + 8 * s, UM_WORKERS / 32 + 1);
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_pfull + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, UM_WORKERS / 32);
      mbar_init(bar_rec + 8 * s, UM_WORKERS / 32);
      mbar_init(bar_ptempty + 8 * s, 1);
    }
    mbar_init_fence();
    *s_abort = 0;
  }
  if (warp == 0) u2_tmem_alloc(smem_u32(s_tmem), 512);
  tc_fence_before();
  __syncthreads();
  u2_cluster_sync();            // both CTAs' barriers exist before any remote arrival or multicast commit
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  long long t_loop0 = 0, t_loop1 = 0;
  const uint32_t idesc = idesc_f16_f32(256, N);
  const int npairs = nstages >> 1;

  if (warp >= 16) {
    setmaxnreg_dec<UM_MMA_REGS>();
    if (warp == 17) {
      // copy warp: this CTA's 128 fine rows of both stages of a pair, two 16 KB bulk copies
      bool ok = true;
      const unsigned char* src = a.fine_img + (((type2 ? a.fine_stages : 0) + sb / UM_STAGE_SAMPLES) * 2 + rank) * 16384;
      int pslot = 0, use = 0;
      for (int p = 0; p < npairs && ok; ++p) {
        ok = um_wait(bar_empty + 8 * pslot, (use & 1) ^ 1, s_abort, t_start);
        if (!ok) break;
        if (elect_one()) {
          const uint32_t dst = smem0 + pslot * 2 * U2_STAGE_BYTES + U2_FINE_HI;
          um_arrive_expect_tx(bar_full + 8 * pslot, 32768);
          um_bulk_g2s(dst, src + (long long)(2 * p) * 32768, 16384, bar_full + 8 * pslot);
          um_bulk_g2s(dst + U2_STAGE_BYTES, src + (long long)(2 * p + 1) * 32768, 16384, bar_full + 8 * pslot);
        }
        __syncwarp();
        if (++pslot == 3) { pslot = 0; ++use; }
      }
    } else if (warp == 16 && rank == 0) {
      // leader: issues the 12 instructions of every pair for both CTAs
      bool ok = true;
      int cpos = 0, ch = 0, pslot = 0, use = 0;
      for (int p = 0; p < npairs && ok; ++p) {
        const int acc = ch & 1;
        if (cpos == 0) {
          const uint32_t par = ((ch >> 1) & 1) ^ 1;
          ok = um_wait(bar_tempty + 8 * acc, par, s_abort, t_start) && u2_wait_cluster(bar_ptempty + 8 * acc, par, s_abort, t_start);
        }
        if (ok) ok = um_wait(bar_full + 8 * pslot, use & 1, s_abort, t_start) &&
                     u2_wait_cluster(bar_pfull + 8 * pslot, use & 1, s_abort, t_start);
        if (!ok) break;
        tc_fence_after();
        const uint32_t sbase = smem0 + pslot * 2 * U2_STAGE_BYTES;
        const uint32_t d = tmem + acc * 256;
        const uint64_t ah = smem_desc(sbase + U2_FINE_HI, U2_LBO, UM_SBO), al = smem_desc(sbase + U2_FINE_LO, U2_LBO, UM_SBO);
        const uint64_t bh = smem_desc(sbase + U2_COARSE_HI, U2_LBO, UM_SBO), bl = smem_desc(sbase + U2_COARSE_LO, U2_LBO, UM_SBO);
        constexpr uint64_t KS = (2 * U2_LBO) >> 4, ST = U2_STAGE_BYTES >> 4;
        const bool last = cpos + 2 == CS;
        if (elect_one()) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint64_t o = (q >> 1) * ST + (q & 1) * KS;
            u2_mma(d, al + o, bh + o, idesc, (cpos != 0 || q != 0) ? 1u : 0u);
            u2_mma(d, ah + o, bl + o, idesc, 1);
            u2_mma(d, ah + o, bh + o, idesc, 1);
          }
          u2_commit_both(bar_empty + 8 * pslot);
          if (last) u2_commit_both(bar_tfull + 8 * acc);
        }
        __syncwarp();
        cpos += 2;
        if (cpos == CS) { cpos = 0; ++ch; }
        if (++pslot == 3) { pslot = 0; ++use; }
      }
    } else if (warp == 16) {
      // peer: forwards its barriers to the leader in the order the leader waits for them
      bool ok = true;
      int cpos = 0, ch = 0, pslot = 0, use = 0;
      for (int p = 0; p < npairs && ok; ++p) {
        ok = um_wait(bar_full + 8 * pslot, use & 1, s_abort, t_start);
        if (!ok) break;
        if (elect_one()) u2_remote_arrive(bar_pfull + 8 * pslot, 0);
        __syncwarp();
        cpos += 2;
        if (cpos == CS) {
          // the workers drain run ch - 1 right after writing the last pair of run ch
          if (ch >= 1) {
            ok = um_wait(bar_tempty + 8 * ((ch - 1) & 1), ((ch - 1) >> 1) & 1, s_abort, t_start);
            if (!ok) break;
            if (elect_one()) u2_remote_arrive(bar_ptempty + 8 * ((ch - 1) & 1), 0);
            __syncwarp();
          }
          cpos = 0;
          ++ch;
        }
        if (++pslot == 3) { pslot = 0; ++use; }
      }
    }
  } else {
    setmaxnreg_inc<UM_WORKER_REGS>();
    const int p_ = tid;
    const int wq = warp & 3, quad = warp >> 2;                     // TMEM lane quarter / sample quad and column quarter
    const int row = wq * 32 + lane;                                // fine row of this CTA = TMEM lane
    const bool weighted_tt = a.weighted && cvp->three_term;
    const int yslot = rec_slot(REC_Y), wslot = rec_slot(REC_W);
    const unsigned kmul = type2 ? 2u : 1u;
    const long long jT = a.j0 + (long long)cb0 * U2_CL_FINE;
    const double fT = cvp->fmin + (double)jT * cvp->df;
    double gT = (double)jT * cvp->gamma;
    gT -= floor(gT);
    // coarse half-task: local coarse block cbl, sample quad `quad`, samples 2 half, 2 half + 1 of the quad
    const int cbl = wq * 16 + (lane >> 1), half = lane & 1;
    const int cbg = type2 ? (int)rank * cptl + cbl : cbl;          // coarse block inside the tile
    const unsigned kcoarse = kmul * (unsigned)(cbg * U2_CL_FINE);
    const bool cactive = cbl < cptl;
    const float* s_wsel = (!type2 && rank) ? s_wy : s_w;           // type 1: the peer computes the y-weighted rows
    const uint32_t rowc_addr = smem0 + quad * U2_LBO + (uint32_t)cbl * 16 + half * 8;
    const uint32_t rows_step = (uint32_t)cptl * 16;

    float m[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) m[c] = 0.f;
    const uint32_t tlane = tmem + ((uint32_t)(wq * 32) << 16) + quad * 64;
    bool ok = true;
    const float comp_full = 1.0f + a.rz_comp * (float)(6 * CS);
    const float comp_last = 1.0f + a.rz_comp * (float)(3 * (int)((ns - (long long)(nchunks - 1) * CS * UM_STAGE_SAMPLES + 7) >> 3));

    auto drain = [&](int ch) {
      const int acc = ch & 1;
      const float comp = ch == nchunks - 1 ? comp_last : comp_full;
      ok = um_wait(bar_tfull + 8 * acc, (ch >> 1) & 1, s_abort, t_start);
      tc_fence_after();
      if (ok) {
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          if (quad * 64 + c0 < N) {
            uint32_t r[32];
            u2_tmem_ld32(tlane + acc * 256 + c0, r);
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
      const long long i = sb + (long long)blk * UM_BLOCK + p_;
      in = i < se;
      if (in) {
        r1 = a.rec1[i];
        r2 = a.rec2[i];
      }
    };
    auto store_block = [&](int blk, const double2& r1, const float4& r2, bool in) {
      const int o = (blk & 1) * UM_BLOCK + p_;
      if (in) {
        const float yv = rec_get(r2, yslot), wv = rec_get(r2, wslot);
        const double bf = r1.y - floor(r1.y);
        const double A = (double)kmul * (frac_of_product(fT, r1.x) + gT);
        s_b64[o] = __double2ull_rn(bf * 18446744073709551616.0);
        s_A32[o] = (unsigned)__double2ll_rn(A * 4294967296.0);
        s_wy[o] = weighted_tt ? wv * yv : yv;
        s_w[o] = weighted_tt ? wv * wv : wv;
      } else {
        s_b64[o] = 0ull;
        s_A32[o] = 0u;
        s_wy[o] = 0.f;
        s_w[o] = 0.f;
      }
    };
    // two samples of one stage: packed (w c, w s) pairs as fp16 hi / lo
    auto compute = [&](int ro, uint32_t (&ph)[2], uint32_t (&pl)[2]) {
      const uint4 b = *reinterpret_cast<const uint4*>(s_b64 + ro);             // two 64-bit fractions
      const uint2 a2 = *reinterpret_cast<const uint2*>(s_A32 + ro);
      const float2 w2 = *reinterpret_cast<const float2*>(s_wsel + ro);
      float c, s;
      um_sincos_fx(a2.x + kcoarse * b.y + __umulhi(kcoarse, b.x), c, s);
      um_split2(w2.x * c, w2.x * s, ph[0], pl[0]);
      um_sincos_fx(a2.y + kcoarse * b.w + __umulhi(kcoarse, b.z), c, s);
      um_split2(w2.y * c, w2.y * s, ph[1], pl[1]);
    };
    auto store = [&](uint32_t addr, const uint32_t (&ph)[2], const uint32_t (&pl)[2]) {
      if (cactive) {
        u2_sts64(addr + U2_COARSE_HI, ph[0] ^ 0x80000000u, ph[1] ^ 0x80000000u);             // (w c, -w s)
        u2_sts64(addr + U2_COARSE_LO, pl[0] ^ 0x80000000u, pl[1] ^ 0x80000000u);
        u2_sts64(addr + rows_step + U2_COARSE_HI, __byte_perm(ph[0], 0, 0x1032), __byte_perm(ph[1], 0, 0x1032));   // (w s, w c)
        u2_sts64(addr + rows_step + U2_COARSE_LO, __byte_perm(pl[0], 0, 0x1032), __byte_perm(pl[1], 0, 0x1032));
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
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_rec);
    ok = um_wait(bar_rec, 0, s_abort, t_start);
    t_loop0 = clock64();
    int g = 0, cpos = 0, ch = 0, pslot = 0, use = 0;
    for (int blk = 0; blk < nblk && ok; ++blk) {
      double2 n1 = make_double2(0.0, 0.0);
      float4 n2 = make_float4(0.f, 0.f, 0.f, 0.f);
      bool nin = false;
      const bool more = blk + 1 < nblk;
      if (more) load_block(blk + 1, n1, n2, nin);
      const int rbase = (blk & 1) * UM_BLOCK + quad * 4 + half * 2;
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
        const int ro0 = rbase + st * UM_STAGE_SAMPLES, ro1 = ro0 + UM_STAGE_SAMPLES;
        uint32_t h0[2], l0[2], h1[2], l1[2];
        compute(ro0, h0, l0);
        compute(ro1, h1, l1);
        ok = um_wait(bar_empty + 8 * pslot, (use & 1) ^ 1, s_abort, t_start);
        if (!ok) break;
        const uint32_t addr = rowc_addr + pslot * 2 * U2_STAGE_BYTES;
        store(addr, h0, l0);
        store(addr + U2_STAGE_BYTES, h1, l1);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * pslot);
        if (++pslot == 3) { pslot = 0; ++use; }
        cpos += 2;
        if (cpos == CS) {
          if (ch >= 1) {
            drain(ch - 1);
            if (!ok) break;
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
      // flush: column n = 64 quad + c -> (sum s, coarse block cb); row = fine index 128 rank + row of block cb
      const int plane0 = type2 ? 4 : 0;
      // type 1: n = s * cpt + cb (wrap = cpt, one block of columns per sum);
      // type 2: n = h * cpt + s * cpt/2 + cbl (wrap = cpt/2; the sum alternates, the half h advances every second wrap)
      const int wrap = type2 ? cpt >> 1 : cpt;
      const int n0 = quad * 64;
      int blkc = n0 / wrap, cw = n0 - blkc * wrap;     // block of `wrap` columns, position inside it
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        const int sidx = type2 ? (blkc & 1) : blkc;
        const int cb = type2 ? (blkc >> 1) * wrap + cw : cw;
        const long long j = (long long)(cb0 + cb) * U2_CL_FINE + 128 * (int)rank + row;
        if (n0 + c < N && cb < ncb && j < a.nf) {
          unsigned long long* pp = a.partial + (long long)(plane0 + sidx) * a.nf_tot + j;
          atomicAdd(pp, (unsigned long long)__float2ll_rn(m[c] * a.fix_scale));
        }
        if (++cw == wrap) { cw = 0; ++blkc; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  u2_cluster_sync();            // nobody leaves while the other CTA's instructions may still read this shared memory
  if (warp == 0) u2_tmem_dealloc(tmem, 512);
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
// host side (one curve, forward grid; called from gls_umma_launch when the pair kernel is selected)
// ---------------------------------------------------------------------------------------------------------------
int gls_umma2_launch(pdc_ctx* ctx, GlsUmmaArgs a, const GlsUmmaPlan& plan, long long nmax, cudaStream_t st) {
  // (a already carries the plan's tiles, splits and run length; gls_umma_plan in gls_umma.cu)
  const int nsplit = plan.nsplit;
  const long long jobs = plan.jobs;                              // clusters
  if (2 * jobs > 0x7fffffffLL) { set_error("pdc_gls: problem too large for one call (%lld jobs)", jobs); return PDC_EINVAL; }
  static bool attr_set[64] = {};
  if (ctx->device >= 0 && ctx->device < 64 && !attr_set[ctx->device]) {
    PDC_CUDA(cudaFuncSetAttribute(gls_umma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)U2_SMEM_BYTES));
    attr_set[ctx->device] = true;
  }
  const long long per = ((((long long)nmax + nsplit - 1) / nsplit) + 63) & ~63LL;
  const long long stages = per * nsplit / UM_STAGE_SAMPLES + 16;
  PDC_TRY(ctx->umma_fine.reserve((size_t)stages * 2 * 32768));
  a.fine_img = ctx->umma_fine.as<unsigned char>();
  a.fine_stages = stages;
  {
    dim3 grid((unsigned)stages, 2, 2);
    gls_umma2_fine_kernel<<<grid, 128, 0, st>>>(a.rec1, (long long)nmax, ctx->umma_fine.as<unsigned char>(), stages);
    PDC_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  if (!ctx->umma_status_clean) {
    PDC_CUDA(cudaMemsetAsync(ctx->umma_status.p, 0, 2 * sizeof(int), st));
    ctx->umma_status_clean = true;
  }
  a.prof = nullptr;
  if (ctx->umma_prof_on) {
    PDC_TRY(ctx->umma_prof.reserve(sizeof(long long) * (4 * (size_t)(2 * jobs) + 8 * 1024)));
    PDC_CUDA(cudaMemsetAsync(ctx->umma_prof.p, 0, sizeof(long long) * (4 * (size_t)(2 * jobs) + 8 * 1024), st));
    a.prof = ctx->umma_prof.as<long long>();
    ctx->umma_prof_jobs = 2 * jobs;
  }
  PDC_TRY(ctx->main_begin(st));
  gls_umma2_kernel<<<(unsigned)(2 * jobs), UM_THREADS, U2_SMEM_BYTES, st>>>(a);
  PDC_CUDA(cudaGetLastError());
  PDC_TRY(ctx->main_end(st));
  ctx->launches++;
  ctx->last_gls_path = 3;
  return PDC_OK;
}

}  // namespace pdc
