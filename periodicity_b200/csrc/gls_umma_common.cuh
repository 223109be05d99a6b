// Pieces shared by the two tensor-core GLS kernels (gls_umma.cu: one CTA per tile; gls_umma2.cu: a pair of CTAs per tile).
#pragma once

#include "gls_common.cuh"
#include "umma.cuh"

namespace pdc {

using namespace umma;

constexpr int UM_FINE = 128;           // fine indices per tile = MMA M = TMEM lanes
constexpr int UM_STAGE_SAMPLES = 16;   // 32 K-slots = two K = 16 steps
constexpr int UM_NSTAGES = 4;
constexpr int UM_BLOCK = 512;          // samples per block of staged records (one per worker thread)
constexpr int UM_WORKERS = 512;        // 16 worker warps
constexpr int UM_THREADS = UM_WORKERS + 128;   // + one warp group: warp 16 issues the tcgen05.mma, warps 17-19 only donate registers
constexpr int UM_WORKER_REGS = 112, UM_MMA_REGS = 24;   // setmaxnreg: 640 x 96 at launch -> 512 x 112 + 128 x 24
constexpr int UM_MAX_T1 = 64, UM_MAX_T2 = 128;   // coarse blocks per tile (N = 4 * 64 = 2 * 128 = 256 columns)
// one stage in shared memory: [K-chunk of 8 slots][row][8 halves]; 16-byte rows, 8-row groups contiguous (SBO = 128 B)
constexpr uint32_t UM_FINE_HI = 0, UM_FINE_LO = 8192, UM_COARSE_HI = 16384, UM_COARSE_LO = 32768;
constexpr uint32_t UM_STAGE_BYTES = 49152;
constexpr uint32_t UM_LBO_FINE = 128 * 16, UM_LBO_COARSE = 256 * 16, UM_SBO = 128;
constexpr uint32_t UM_REC_BYTES = 2 * UM_BLOCK * (8 + 4 + 4 + 4);
constexpr uint32_t UM_SMEM_BYTES = UM_NSTAGES * UM_STAGE_BYTES + UM_REC_BYTES + 256;
constexpr long long UM_MAX_JOB_SAMPLES = 16384;   // FP32 masters are flushed to the fixed-point plane at least this often
constexpr long long UM_MAX_FINE_BYTES = 16LL << 30;   // scratch for the precomputed fine operand of one curve (C5: 4 GB)
constexpr long long UM_WAIT_CLOCKS = 4000000000LL;   // ~2 s: far beyond any legitimate wait

// Work decomposition of one call (pure host arithmetic; gls_umma_plan in gls_umma.cu, also exported for the CPU tests as
// pdc_debug_umma_plan).  path: 1 one CTA per tile, 2 the same with the fine operand precomputed, 3 a pair of CTAs per tile.
struct GlsUmmaPlan {
  int path;
  int fine;            // fine indices per tile: 128 (paths 1, 2) or 256 (path 3)
  int nC;              // coarse blocks of `fine` frequencies per curve
  int nt1, cpt1;       // type-1 tiles {C, S, YC, YS}: count and coarse blocks per tile (4 cpt1 <= 256 columns)
  int nt2, cpt2;       // type-2 tiles {C2, S2}: count and coarse blocks per tile (2 cpt2 <= 256 columns)
  int nsplit;          // sample splits per curve
  int chunk_stages;    // stages of 16 samples per accumulation run in TMEM (even)
  long long jobs;      // tiles x splits x curves (path 3: clusters)
  long long fine_bytes;   // scratch for the precomputed fine operand (0 on path 1)
};
struct GlsUmmaKnobs {   // the ctx's tuning knobs (environment), -1 / 0 = automatic as documented in pdc_common.cuh
  int fine, cg2, nsplit, chunk;
};
void gls_umma_plan(int sm_count, long long B, long long nf, long long nmax, const GlsUmmaKnobs& k, GlsUmmaPlan* out);

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
  int chunk_stages;              // stages (of 16 samples) per accumulation run in TMEM
  float rz_comp;                 // expected relative truncation loss of the accumulator per tcgen05.mma of a run (see the drain)
  float fix_scale;
  const unsigned char* fine_img; // FINE_PRE: fine operand of the whole curve as shared-memory images, [2 types][stage][16 KB] (gls_umma_fine_kernel)
  long long fine_stages;         // stages per type in fine_img
  int* status;                   // set non-zero on a protocol time-out (this call's word; the epilogue reads it)
  int* status_next;              // the next call's word: cleared here, so that a time-out poisons one call, not the ctx
  int dbg;                       // timing experiments (env PDC_GLS_UMMA_DBG): 1 no MMA, 2 no operand production, 4 no drain, 8 no proxy fence, 16 trace
  long long* prof;               // optional [jobs][4] clock64 stamps (start, main loop begin, main loop end, flush end)
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
__device__ __forceinline__ void um_sts128(uint32_t addr, const uint32_t (&v)[4]) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
// global -> shared bulk copy (TMA without a tensor map): completion is counted in bytes on `bar`
__device__ __forceinline__ void um_bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void um_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// phase (2^-32 turn, two's complement) -> (cos, sin)
__device__ __forceinline__ void um_sincos_fx(unsigned fx, float& c, float& s) {
  const float x = (float)(int)fx * 1.4629180792671596e-9f;   // 2 pi / 2^32
  __sincosf(x, &s, &c);
}

}  // namespace pdc
