// Hardware probe for the tcgen05 path used by periodicity_b200/csrc/gls_umma.cu (B200, sm_100a).
//   1. layout:    D = A B^T with exactly representable fp16 data, no-swizzle K-major operand tiles written with plain
//                 st.shared; checks the shared-memory descriptor (LBO / SBO meaning), the instruction descriptor, the
//                 K-step advance, the accumulate flag and the tcgen05.ld 32x32b lane/column mapping against the host.
//   2. rounding:  how the FP32 accumulator in TMEM rounds (nearest or truncation), one product per instruction and
//                 sixteen tiny products per instruction.
//   3. rate:      back-to-back tcgen05.mma 128 x N x 16 per SM (N = 128, 192, 256), all SMs busy.
//   4. tmem read: tcgen05.ld rate of 8 warps draining a 128 x 256 FP32 tile.
// Every wait is bounded; a protocol error prints FAIL instead of hanging.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_probe umma_probe.cu && ./umma_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#include "../../periodicity_b200/csrc/umma.cuh"

using namespace pdc::umma;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

// ---------------------------------------------------------------------------------------------------------------
// generic one-CTA GEMM: A [128][K], B [N][K] given in global memory as plain row-major fp16; the CTA arranges them in
// shared memory as [K/8][rows][8] (core matrices of 8 rows x 16 bytes, row groups contiguous), issues `nrep` rounds
// of K/16 instructions and writes D [128][N] (FP32).
// `swap` exchanges the LBO / SBO fields of the descriptors (to find out which is which if the first guess is wrong).
// `second`: optional second operand pair (A2, B2) issued `nrep2` times after the first, accumulating.
// ---------------------------------------------------------------------------------------------------------------
struct GemmArgs {
  const __half *A, *B, *A2, *B2;
  float* D;
  int N, K, nrep2, swap;
  int* status;
};

__global__ void __launch_bounds__(128, 1) gemm_probe_kernel(GemmArgs g) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int K = g.K, N = g.N;
  __half* sA = reinterpret_cast<__half*>(smem);
  __half* sB = sA + 128 * K;
  __half* sA2 = sB + 256 * K;
  __half* sB2 = sA2 + 128 * K;
  for (int e = tid; e < 128 * K; e += 128) {
    const int r = e / K, k = e % K;
    sA[((k >> 3) * 128 + r) * 8 + (k & 7)] = g.A[e];
    if (g.A2) sA2[((k >> 3) * 128 + r) * 8 + (k & 7)] = g.A2[e];
  }
  for (int e = tid; e < 256 * K; e += 128) {
    const int r = e / K, k = e % K;
    const bool in = r < N;
    sB[((k >> 3) * 256 + r) * 8 + (k & 7)] = in ? g.B[r * K + k] : __float2half(0.f);
    if (g.B2) sB2[((k >> 3) * 256 + r) * 8 + (k & 7)] = in ? g.B2[r * K + k] : __float2half(0.f);
  }
  fence_proxy_async_smem();
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_base_s), 256);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init_fence();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = idesc_f16_f32(128, N);
    const uint32_t lboA = 128 * 16, lboB = 256 * 16, sbo = 128;
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint32_t a0 = smem_u32(sA) + ks * 2 * lboA, b0 = smem_u32(sB) + ks * 2 * lboB;
      const uint64_t da = g.swap ? smem_desc(a0, sbo, lboA) : smem_desc(a0, lboA, sbo);
      const uint64_t db = g.swap ? smem_desc(b0, sbo, lboB) : smem_desc(b0, lboB, sbo);
      mma_f16_ss(tmem, da, db, idesc, ks > 0);
    }
    for (int rep = 0; rep < g.nrep2; ++rep)
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint32_t a0 = smem_u32(sA2) + ks * 2 * lboA, b0 = smem_u32(sB2) + ks * 2 * lboB;
        mma_f16_ss(tmem, smem_desc(a0, lboA, sbo), smem_desc(b0, lboB, sbo), idesc, 1);
      }
    mma_commit(smem_u32(&bar));
  }
  const bool ok = mbar_wait(smem_u32(&bar), 0, 1u << 20);
  tc_fence_after();
  if (!ok) {
    if (tid == 0) *g.status = 1;
  } else {
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j)
        if (c0 + j < N) g.D[(warp * 32 + (tid & 31)) * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

static int run_gemm(const std::vector<__half>& A, const std::vector<__half>& B, const std::vector<__half>* A2,
                    const std::vector<__half>* B2, int N, int K, int nrep2, int swap, std::vector<float>& D) {
  __half *dA, *dB, *dA2 = nullptr, *dB2 = nullptr;
  float* dD;
  int* dst;
  CK(cudaMalloc(&dA, A.size() * 2));
  CK(cudaMalloc(&dB, B.size() * 2));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  if (A2) {
    CK(cudaMalloc(&dA2, A2->size() * 2));
    CK(cudaMalloc(&dB2, B2->size() * 2));
    CK(cudaMemcpy(dA2, A2->data(), A2->size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB2, B2->data(), B2->size() * 2, cudaMemcpyHostToDevice));
  }
  CK(cudaMalloc(&dD, sizeof(float) * 128 * N));
  CK(cudaMemset(dD, 0xff, sizeof(float) * 128 * N));
  CK(cudaMalloc(&dst, sizeof(int)));
  CK(cudaMemset(dst, 0, sizeof(int)));
  GemmArgs g{dA, dB, dA2, dB2, dD, N, K, nrep2, swap, dst};
  const size_t smem = (size_t)2 * (128 + 256) * K * 2 + 1024;
  CK(cudaFuncSetAttribute(gemm_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gemm_probe_kernel<<<1, 128, smem>>>(g);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("  kernel error: %s\n", cudaGetErrorString(e));
    exit(3);   // the context is gone
  }
  int st = 0;
  CK(cudaMemcpy(&st, dst, sizeof(int), cudaMemcpyDeviceToHost));
  D.resize((size_t)128 * N);
  CK(cudaMemcpy(D.data(), dD, sizeof(float) * 128 * N, cudaMemcpyDeviceToHost));
  cudaFree(dA); cudaFree(dB); cudaFree(dA2); cudaFree(dB2); cudaFree(dD); cudaFree(dst);
  return st;
}

static void test_layout() {
  printf("== 1. layout ==\n");
  for (int K : {16, 64}) {
    for (int N : {256, 160, 64}) {
      std::vector<__half> A((size_t)128 * K), B((size_t)N * K);
      for (int r = 0; r < 128; ++r)
        for (int k = 0; k < K; ++k) A[(size_t)r * K + k] = __float2half((float)((r * 3 + k * 5) % 17 - 8) / 8.f);
      for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) B[(size_t)n * K + k] = __float2half((float)((n * 7 + k * 11) % 13 - 6) / 4.f);
      for (int swap = 0; swap < 1; ++swap) {   // swapped fields fault (illegal address): the documented meaning is right
        std::vector<float> D;
        const int st = run_gemm(A, B, nullptr, nullptr, N, K, 0, swap, D);
        double maxerr = 0;
        long bad = 0;
        for (int r = 0; r < 128; ++r)
          for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k)
              ref += (double)__half2float(A[(size_t)r * K + k]) * (double)__half2float(B[(size_t)n * K + k]);
            const double err = fabs(ref - (double)D[(size_t)r * N + n]);
            if (!(err <= 1e-6)) ++bad;
            if (err > maxerr || err != err) maxerr = err;
          }
        printf("  K=%3d N=%3d fields %s: status=%d max|err|=%.3g mismatches=%ld  %s\n", K, N,
               swap ? "swapped(LBO<->SBO)" : "as documented   ", st, maxerr, bad, (st == 0 && bad == 0) ? "OK" : "WRONG");
      }
    }
  }
}

static void test_rounding() {
  printf("== 2. accumulator rounding (units: ulp of 1.5 = 2^-23; 64 accumulating instructions) ==\n");
  const int K = 16, N = 16;
  std::vector<__half> A1((size_t)128 * K, __float2half(0.f)), B1((size_t)N * K, __float2half(0.f));
  std::vector<__half> A2((size_t)128 * K, __float2half(0.f)), B2((size_t)N * K, __float2half(0.f));
  for (int r = 0; r < 128; ++r) {
    A1[(size_t)r * K] = __float2half((r & 1) ? -1.5f : 1.5f);
    for (int k = 0; k < K; ++k) A2[(size_t)r * K + k] = __float2half(ldexpf(1.f, -12));
  }
  const float c[4] = {1.5f, -1.5f, 0.5f, -0.5f};
  for (int n = 0; n < N; ++n) {
    B1[(size_t)n * K] = __float2half(1.f);
    if (n < 4) B2[(size_t)n * K] = __float2half(c[n] * ldexpf(1.f, -12));                  // one product of c 2^-24
    else if (n < 8)
      for (int k = 0; k < K; ++k) B2[(size_t)n * K + k] = __float2half(c[n - 4] * ldexpf(1.f, -16));  // 16 x c 2^-28
  }
  std::vector<float> D;
  const int st = run_gemm(A1, B1, &A2, &B2, N, K, 64, 0, D);
  printf("  status=%d\n", st);
  const char* names[8] = {"+0.75 ulp x1", "-0.75 ulp x1", "+0.25 ulp x1", "-0.25 ulp x1",
                          "+0.75 ulp as 16 products", "-0.75 ulp as 16", "+0.25 ulp as 16", "-0.25 ulp as 16"};
  for (int row = 0; row < 2; ++row)
    for (int n = 0; n < 8; ++n) {
      const double base = row ? -1.5 : 1.5;
      const double drift = ((double)D[(size_t)row * N + n] - base) / ldexp(1.0, -23);
      const double exact = 64.0 * (double)c[n & 3] * 0.5;
      printf("  base %+4.1f addend %-26s: drift %+8.2f ulp (exact %+6.1f; nearest-even would give %+6.1f)\n", base,
             names[n], drift, exact, (fabs(c[n & 3]) > 1.0 ? 64.0 * (c[n & 3] > 0 ? 1 : -1) : 0.0));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 3. instruction rate
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, long long* cycles, int* status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 96 * 1024 / 4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0x3c003c00u;  // 1.0h
  fence_proxy_async_smem();
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_base_s), 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init_fence();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = idesc_f16_f32(128, N);
    const uint32_t base = smem_u32(smem);
    uint32_t parity = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // 6 instructions (2 K-steps of three products) out of a 48 KB stage = one 16-sample stage of the GLS kernel
      const uint32_t st = base + (it & 1) * 48 * 1024;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint64_t ah = smem_desc(st + ks * 4096, 2048, 128), al = smem_desc(st + 8192 + ks * 4096, 2048, 128);
        const uint64_t bh = smem_desc(st + 16384 + ks * 8192, 4096, 128), bl = smem_desc(st + 32768 + ks * 8192, 4096, 128);
        mma_f16_ss(tmem + (it & 1) * 256, al, bh, idesc, 1);
        mma_f16_ss(tmem + (it & 1) * 256, ah, bl, idesc, 1);
        mma_f16_ss(tmem + (it & 1) * 256, ah, bh, idesc, 1);
      }
      if ((it & 15) == 15) {
        mma_commit(smem_u32(&bar));
        if (!mbar_wait(smem_u32(&bar), parity, 1u << 20)) { *status = 1; break; }
        parity ^= 1;
      }
    }
    mma_commit(smem_u32(&bar));
    if (!mbar_wait(smem_u32(&bar), parity, 1u << 20)) *status = 1;
    cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static void test_rate(int nsm) {
  printf("== 3. tcgen05.mma rate (kind::f16, M=128, K=16, cta_group::1; 6 instructions per iteration) ==\n");
  long long* dcyc;
  int* dst;
  CK(cudaMalloc(&dcyc, sizeof(long long) * nsm));
  CK(cudaMalloc(&dst, sizeof(int)));
  const int smem = 96 * 1024 + 1024;
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int grid : {1, nsm})
    for (int N : {128, 192, 256}) {
      const int iters = 8192;
      CK(cudaMemset(dst, 0, sizeof(int)));
      rate_kernel<<<grid, 128, smem>>>(N, 64, dcyc, dst);   // warm-up
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      rate_kernel<<<grid, 128, smem>>>(N, iters, dcyc, dst);
      CK(cudaEventRecord(e1));
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("  kernel error: %s\n", cudaGetErrorString(e)); exit(3); }
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      std::vector<long long> cyc(grid);
      int st;
      CK(cudaMemcpy(cyc.data(), dcyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(&st, dst, sizeof(int), cudaMemcpyDeviceToHost));
      double cmax = 0;
      for (long long c : cyc) cmax = c > cmax ? c : cmax;
      const double macs = (double)iters * 6.0 * 128.0 * N * 16.0;
      printf("  grid=%3d N=%3d: status=%d  %.1f clk per instruction, %.0f MAC/clk/SM, %.1f us -> %.0f TFLOP/s chip (dense fp16)\n",
             grid, N, st, cmax / (iters * 6.0), macs / cmax, ms * 1e3, 2.0 * macs * grid / (ms * 1e-3) * 1e-12);
    }
  cudaFree(dcyc);
  cudaFree(dst);
}

// ---------------------------------------------------------------------------------------------------------------
// 4. tcgen05.ld rate: 8 warps (two per lane quarter) drain 128 lanes x 256 columns, `iters` times, adding into registers
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) tmem_read_kernel(int iters, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_base_s), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  float acc[128];
#pragma unroll
  for (int j = 0; j < 128; ++j) acc[j] = 0.f;
  const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t r[32];
      tmem_ld32(taddr + q * 32 + (it & 1) * 256, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[q * 32 + j] += __uint_as_float(r[j]);
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 128; ++j) s += acc[j];
  sink[blockIdx.x * 256 + tid] = s;
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static void test_tmem_read() {
  printf("== 4. tcgen05.ld: 8 warps drain a 128 x 256 FP32 tile (128 KB) and add it into registers ==\n");
  long long* dcyc;
  float* sink;
  CK(cudaMalloc(&dcyc, sizeof(long long)));
  CK(cudaMalloc(&sink, sizeof(float) * 256));
  const int iters = 1024;
  tmem_read_kernel<<<1, 256>>>(8, dcyc, sink);
  tmem_read_kernel<<<1, 256>>>(iters, dcyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  kernel error: %s\n", cudaGetErrorString(e)); exit(3); }
  long long cyc;
  CK(cudaMemcpy(&cyc, dcyc, sizeof(long long), cudaMemcpyDeviceToHost));
  printf("  %.0f clk per tile drain (%.0f B/clk/SM)\n", (double)cyc / iters, 131072.0 * iters / (double)cyc);
  cudaFree(dcyc);
  cudaFree(sink);
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("device: %s, %d SMs, cc %d.%d\n", p.name, p.multiProcessorCount, p.major, p.minor);
  test_layout();
  test_rounding();
  test_rate(p.multiProcessorCount);
  test_tmem_read();
  printf("done\n");
  return 0;
}
