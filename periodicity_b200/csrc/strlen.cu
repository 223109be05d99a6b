// String Length (Dworetsky 1983) over a grid of trial periods.
//
// Replaces `pool.map(self._stringlength, periods)` of `StringLength.__call__`
// (reference src/periodicity/phase.py:68-70) and `StringLength._stringlength`
// (phase.py:45-51):
//     fold = self.m.fold(period)          phi = (t / P) % 1, samples re-sorted by phi
//                                         (core.py:543-544; the TSeries constructor sorts, core.py:473-477)
//     ll   = hypot(roll(m, -1) - m, roll(phi, -1) - phi).sum()
// i.e. the length of the closed polygon through the folded light curve, the closing segment
// being taken literally as (phi[0] - phi[N-1], m[0] - m[N-1]).  `m` (the signal scaled to
// [-0.25, 0.25], phase.py:64-65) is computed by the Python front end.
//
// Mapping: one thread block per trial period.  The block computes the N phases (correctly
// rounded t / P, exact q - floor(q), both as numpy does), sorts (phase bits, sample index) with a
// bitonic network -- in shared memory when the padded curve fits (N <= 16384), else in a per-block
// slice of global scratch that stays L2 resident, with all stages that fit a chunk done in shared memory
// (sl_hybrid_kernel) -- and sums the N segment lengths in FP64 with a
// fixed reduction tree.  The composite key makes the order total, so equal phases keep their
// time order exactly like the stable sort behind `sortby` (core.py:477).
//
// Bound: shared-memory bandwidth / barrier latency of the sorting network
// (log2(Np) (log2(Np) + 1) / 2 compare-exchange stages over Np keys).
#include "pdc_common.cuh"

namespace pdc {

constexpr int SL_THREADS = 256;
constexpr int SL_SMEM_MAX_PAD = 16384;  // 16384 * (8 + 4) B = 192 KB of keys + indices

__device__ __forceinline__ bool sl_less(unsigned long long ka, unsigned ia, unsigned long long kb, unsigned ib) {
  return ka < kb || (ka == kb && ia < ib);
}

template <bool SMEM>
__global__ void __launch_bounds__(SL_THREADS)
sl_kernel(const double* __restrict__ t, const double* __restrict__ m, int n, int npad,
          const double* __restrict__ periods, long long np, unsigned long long* __restrict__ gkeys,
          unsigned* __restrict__ gidx, double* __restrict__ ell_out, long long* __restrict__ idx_out) {
  extern __shared__ __align__(16) unsigned char sl_smem[];
  __shared__ double scratch[33];
  unsigned long long* keys;
  unsigned* idx;
  if (SMEM) {
    keys = reinterpret_cast<unsigned long long*>(sl_smem);
    idx = reinterpret_cast<unsigned*>(keys + npad);
  } else {
    keys = gkeys + (size_t)blockIdx.x * npad;
    idx = gidx + (size_t)blockIdx.x * npad;
  }

  for (long long p = blockIdx.x; p < np; p += gridDim.x) {
    const double P = periods[p];
    // phases: (t / P) % 1 as numpy computes it (correctly rounded quotient, exact remainder)
    for (int i = threadIdx.x; i < npad; i += SL_THREADS) {
      unsigned long long k = ~0ull;  // padding sorts last
      if (i < n) {
        const double q = __ddiv_rn(t[i], P);
        const double phi = __dadd_rn(q, -floor(q));
        k = (unsigned long long)__double_as_longlong(phi);  // phi >= 0: bit pattern is monotone; NaN sorts after every number
      }
      keys[i] = k;
      idx[i] = (unsigned)i;
    }
    __syncthreads();

    // bitonic sort, ascending in (key, index)
    for (int k = 2; k <= npad; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = threadIdx.x; i < (npad >> 1); i += SL_THREADS) {
          const int l = ((i / j) * (j << 1)) + (i % j);
          const int r = l + j;
          const unsigned long long kl = keys[l], kr = keys[r];
          const unsigned il = idx[l], ir = idx[r];
          const bool up = (l & k) == 0;
          if (sl_less(kr, ir, kl, il) == up) {
            keys[l] = kr;
            keys[r] = kl;
            idx[l] = ir;
            idx[r] = il;
          }
        }
        __syncthreads();
      }
    }

    // ll = sum_j hypot(m[j+1] - m[j], phi[j+1] - phi[j]), indices mod N (np.roll, phase.py:50)
    double acc = 0.0;
    for (int j = threadIdx.x; j < n; j += SL_THREADS) {
      const int jn = j + 1 < n ? j + 1 : 0;
      const double dphi = __dadd_rn(__longlong_as_double((long long)keys[jn]), -__longlong_as_double((long long)keys[j]));
      const double dm = __dadd_rn(m[idx[jn]], -m[idx[j]]);
      acc += hypot(dm, dphi);
    }
    const double ll = block_sum(acc, scratch);
    if (threadIdx.x == 0) {
      ell_out[p] = ll;
      idx_out[p] = p;
    }
    __syncthreads();  // keys / idx are rewritten by the next period
  }
}

// Curves too long for shared memory (padded length > SL_SMEM_MAX_PAD): the array lives in a per-block slice of global
// scratch, but only the compare-exchange stages whose partner distance j spans chunks (j >= SL_CHUNK) run there; all
// stages with j < SL_CHUNK of one merge step are done on a chunk loaded into shared memory.  For 20,000 samples (32,768
// padded) that is 7 passes over the L2-resident slice instead of 120 (round 1 ran every stage in global memory).
constexpr int SL_CHUNK = 8192;   // records per shared-memory chunk: 8192 * 12 B = 96 KB, two blocks per SM

__global__ void __launch_bounds__(SL_THREADS)
sl_hybrid_kernel(const double* __restrict__ t, const double* __restrict__ m, int n, int npad,
                 const double* __restrict__ periods, long long np, unsigned long long* __restrict__ gkeys,
                 unsigned* __restrict__ gidx, double* __restrict__ ell_out, long long* __restrict__ idx_out) {
  extern __shared__ __align__(16) unsigned char sl_smem[];
  __shared__ double scratch[33];
  unsigned long long* skey = reinterpret_cast<unsigned long long*>(sl_smem);
  unsigned* sidx = reinterpret_cast<unsigned*>(skey + SL_CHUNK);
  unsigned long long* keys = gkeys + (size_t)blockIdx.x * npad;
  unsigned* idx = gidx + (size_t)blockIdx.x * npad;
  const int nchunk = npad / SL_CHUNK;

  // stages j = jtop .. 1 of merge step k on the chunk in shared memory whose first record has global index `base`
  auto chunk_stages = [&](int k, int jtop, int base) {
    for (int j = jtop; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < (SL_CHUNK >> 1); i += SL_THREADS) {
        const int l = ((i / j) * (j << 1)) + (i % j);
        const int r = l + j;
        const unsigned long long kl = skey[l], kr = skey[r];
        const unsigned il = sidx[l], ir = sidx[r];
        const bool up = ((base + l) & k) == 0;
        if (sl_less(kr, ir, kl, il) == up) {
          skey[l] = kr;
          skey[r] = kl;
          sidx[l] = ir;
          sidx[r] = il;
        }
      }
      __syncthreads();
    }
  };

  for (long long p = blockIdx.x; p < np; p += gridDim.x) {
    const double P = periods[p];
    // phase A: phases of one chunk, full bitonic sort of the chunk (merge steps k = 2 .. SL_CHUNK), store
    for (int c = 0; c < nchunk; ++c) {
      const int base = c * SL_CHUNK;
      for (int i = threadIdx.x; i < SL_CHUNK; i += SL_THREADS) {
        unsigned long long k = ~0ull;  // padding sorts last
        const int g = base + i;
        if (g < n) {
          const double q = __ddiv_rn(t[g], P);
          const double phi = __dadd_rn(q, -floor(q));
          k = (unsigned long long)__double_as_longlong(phi);
        }
        skey[i] = k;
        sidx[i] = (unsigned)g;
      }
      __syncthreads();
      for (int k = 2; k <= SL_CHUNK; k <<= 1) chunk_stages(k, k >> 1, base);
      for (int i = threadIdx.x; i < SL_CHUNK; i += SL_THREADS) {
        keys[base + i] = skey[i];
        idx[base + i] = sidx[i];
      }
      __syncthreads();
    }
    // phase B: merge steps that span chunks
    for (int k = SL_CHUNK << 1; k <= npad; k <<= 1) {
      for (int j = k >> 1; j >= SL_CHUNK; j >>= 1) {   // partners in different chunks: global memory
        for (int i = threadIdx.x; i < (npad >> 1); i += SL_THREADS) {
          const int l = ((i / j) * (j << 1)) + (i % j);
          const int r = l + j;
          const unsigned long long kl = keys[l], kr = keys[r];
          const unsigned il = idx[l], ir = idx[r];
          const bool up = (l & k) == 0;
          if (sl_less(kr, ir, kl, il) == up) {
            keys[l] = kr;
            keys[r] = kl;
            idx[l] = ir;
            idx[r] = il;
          }
        }
        __syncthreads();
      }
      for (int c = 0; c < nchunk; ++c) {               // the remaining stages of this merge step, chunk by chunk
        const int base = c * SL_CHUNK;
        for (int i = threadIdx.x; i < SL_CHUNK; i += SL_THREADS) {
          skey[i] = keys[base + i];
          sidx[i] = idx[base + i];
        }
        __syncthreads();
        chunk_stages(k, SL_CHUNK >> 1, base);
        for (int i = threadIdx.x; i < SL_CHUNK; i += SL_THREADS) {
          keys[base + i] = skey[i];
          idx[base + i] = sidx[i];
        }
        __syncthreads();
      }
    }
    // ll = sum_j hypot(m[j+1] - m[j], phi[j+1] - phi[j]), indices mod N (np.roll, phase.py:50)
    double acc = 0.0;
    for (int j = threadIdx.x; j < n; j += SL_THREADS) {
      const int jn = j + 1 < n ? j + 1 : 0;
      const double dphi = __dadd_rn(__longlong_as_double((long long)keys[jn]), -__longlong_as_double((long long)keys[j]));
      const double dm = __dadd_rn(m[idx[jn]], -m[idx[j]]);
      acc += hypot(dm, dphi);
    }
    const double ll = block_sum(acc, scratch);
    if (threadIdx.x == 0) {
      ell_out[p] = ll;
      idx_out[p] = p;
    }
    __syncthreads();
  }
}

int strlen_run(pdc_ctx* ctx, const double* t, const double* m, int64_t n, const double* periods, int64_t np,
               double* ell_out, int64_t* argmin_out, double* min_out, cudaStream_t st) {
  if (n < 1) { set_error("pdc_stringlength: need at least one sample"); return PDC_EINVAL; }
  if (np < 1) { set_error("pdc_stringlength: need at least one trial period"); return PDC_EINVAL; }
  if (n > (1 << 26)) { set_error("pdc_stringlength: at most 2^26 samples per curve"); return PDC_EINVAL; }
  if (np > 0x7fffffffLL) { set_error("pdc_stringlength: at most 2^31-1 trial periods per call"); return PDC_EINVAL; }
  int npad = 2;
  while (npad < n) npad <<= 1;
  const bool smem = npad <= SL_SMEM_MAX_PAD;
  const size_t per_block = (size_t)npad * (sizeof(unsigned long long) + sizeof(unsigned));

  ScratchScope scratch(ctx, st);
  PDC_TRY(scratch.acquire());
  long long grid;
  unsigned long long* gkeys = nullptr;
  unsigned* gidx = nullptr;
  if (smem) {
    // resident blocks per SM by shared memory (227 KB usable) and threads (2048)
    long long per_sm = (long long)((227 * 1024) / (per_block + 1024));
    if (per_sm > 2048 / SL_THREADS) per_sm = 2048 / SL_THREADS;
    if (per_sm < 1) per_sm = 1;
    grid = (long long)ctx->sm_count * per_sm;
  } else {
    // global scratch: one slice per block, capped at 2 GiB in total
    grid = (long long)(((size_t)2 << 30) / per_block);
    if (grid > 2LL * ctx->sm_count) grid = 2LL * ctx->sm_count;
    if (grid < 1) grid = 1;
    PDC_TRY(ctx->partial.reserve(per_block * (size_t)grid));
    gkeys = ctx->partial.as<unsigned long long>();
    gidx = reinterpret_cast<unsigned*>(gkeys + (size_t)grid * npad);
  }
  if (grid > np) grid = np;
  PDC_TRY(ctx->blockred.reserve(sizeof(long long) * (size_t)np));
  long long* idx_out = ctx->blockred.as<long long>();

  PDC_TRY(ctx->main_begin(st));
  if (smem) {
    PDC_CUDA(cudaFuncSetAttribute(sl_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)per_block));
    sl_kernel<true><<<(unsigned)grid, SL_THREADS, per_block, st>>>(t, m, (int)n, npad, periods, (long long)np,
                                                                    nullptr, nullptr, ell_out, idx_out);
  } else {
    const size_t chunk_bytes = (size_t)SL_CHUNK * (sizeof(unsigned long long) + sizeof(unsigned));
    PDC_CUDA(cudaFuncSetAttribute(sl_hybrid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chunk_bytes));
    sl_hybrid_kernel<<<(unsigned)grid, SL_THREADS, chunk_bytes, st>>>(t, m, (int)n, npad, periods, (long long)np, gkeys,
                                                                      gidx, ell_out, idx_out);
  }
  PDC_CUDA(cudaGetLastError());
  ctx->launches++;
  PDC_TRY(ctx->main_end(st));
  if (argmin_out || min_out) {
    argext_final_kernel<-1><<<1, 256, 0, st>>>(ell_out, idx_out, (int)np, (long long*)argmin_out, min_out);
    PDC_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  PDC_TRY(scratch.release());
  return PDC_OK;
}

}  // namespace pdc
