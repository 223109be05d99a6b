// extern "C" surface of libperiodicity_b200.so (see include/periodicity_b200.h).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>

#include "pdc_common.cuh"
#include "gls_umma_common.cuh"

namespace pdc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) in `%s` at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();  // clear the sticky-less OOM so the ctx stays usable
    return PDC_ENOMEM;
  }
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice)
    return PDC_ENODEVICE;
  return PDC_ECUDA;
}

int DevBuf::reserve(size_t bytes) {
  if (bytes <= cap && p) return PDC_OK;
  if (bytes == 0) bytes = 16;
  // grow geometrically below 1 GiB to limit reallocations, exactly above
  size_t want = bytes;
  if (cap && bytes < ((size_t)1 << 30) && bytes < cap * 2) want = cap * 2;
  if (p) {
    cudaError_t e = cudaFree(p);  // implicit device sync: no kernel can still be using it
    p = nullptr;
    cap = 0;
    if (e != cudaSuccess) return cuda_fail(e, "cudaFree", __FILE__, __LINE__);
  }
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess && want != bytes) {
    cudaGetLastError();
    want = bytes;
    e = cudaMalloc(&p, want);
  }
  if (e != cudaSuccess) {
    p = nullptr;
    return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__);
  }
  cap = want;
  return PDC_OK;
}

void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
}

int PinnedBuf::reserve(size_t bytes) {
  if (bytes <= cap && p) return PDC_OK;
  if (bytes == 0) bytes = 16;
  if (p) {
    cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
  cudaError_t e = cudaMallocHost(&p, bytes);
  if (e != cudaSuccess) {
    p = nullptr;
    return cuda_fail(e, "cudaMallocHost", __FILE__, __LINE__);
  }
  cap = bytes;
  return PDC_OK;
}

void PinnedBuf::release() {
  if (p) cudaFreeHost(p);
  p = nullptr;
  cap = 0;
}

}  // namespace pdc

int pdc_ctx::scratch_acquire(cudaStream_t st) {
  PDC_CUDA(cudaStreamWaitEvent(st, ev_done, 0));  // no-op until the event has been recorded once
  return PDC_OK;
}

int pdc_ctx::scratch_release(cudaStream_t st) {
  PDC_CUDA(cudaEventRecord(ev_done, st));
  return PDC_OK;
}

int pdc_ctx::main_begin(cudaStream_t st) {
  if (ev_pending.size() >= 1024) PDC_TRY(main_resolve());
  if (ev_free.empty()) {
    cudaEvent_t a, b;
    PDC_CUDA(cudaEventCreate(&a));
    PDC_CUDA(cudaEventCreate(&b));
    ev_free.emplace_back(a, b);
  }
  ev_cur_begin = ev_free.back().first;
  ev_cur_end = ev_free.back().second;
  ev_free.pop_back();
  PDC_CUDA(cudaEventRecord(ev_cur_begin, st));
  return PDC_OK;
}

int pdc_ctx::main_end(cudaStream_t st) {
  PDC_CUDA(cudaEventRecord(ev_cur_end, st));
  ev_pending.emplace_back(ev_cur_begin, ev_cur_end);
  return PDC_OK;
}

int pdc_ctx::main_resolve() {
  for (auto& pr : ev_pending) {
    PDC_CUDA(cudaEventSynchronize(pr.second));
    float ms = 0.f;
    PDC_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
    main_ms_total += ms;
    main_ms_last = ms;
    main_count++;
    ev_free.push_back(pr);
  }
  ev_pending.clear();
  return PDC_OK;
}

namespace pdc {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

}  // namespace pdc

using namespace pdc;

extern "C" {

int pdc_version(void) { return PDC_VERSION; }

const char* pdc_last_error(void) { return g_err; }

int pdc_ctx_create(pdc_ctx** out, int device) {
  if (!out) { set_error("pdc_ctx_create: out is NULL"); return PDC_EINVAL; }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("pdc_ctx_create: no usable CUDA device (%s); this library has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return PDC_ENODEVICE;
  }
  if (device < 0 || device >= ndev) {
    set_error("pdc_ctx_create: device %d out of range (0..%d)", device, ndev - 1);
    return PDC_EINVAL;
  }
  DeviceGuard guard(device);
  if (!guard.ok) { set_error("pdc_ctx_create: cudaSetDevice(%d) failed", device); return PDC_ECUDA; }
  cudaDeviceProp prop;
  PDC_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_error("pdc_ctx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
              device, prop.major, prop.minor);
    return PDC_ENODEVICE;
  }
  pdc_ctx* ctx = new (std::nothrow) pdc_ctx();
  if (!ctx) { set_error("pdc_ctx_create: out of host memory"); return PDC_ENOMEM; }
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  if (const char* g = getenv("PDC_GLS_GEOM")) { ctx->gls_geom = atoi(g); ctx->gls_geom_forced = true; }
  if (const char* g = getenv("PDC_GLS_THREE_TERM")) ctx->gls_three_term = atoi(g) != 0;
  if (const char* g = getenv("PDC_GLS_NSPLIT")) ctx->gls_nsplit_override = atoi(g);
  if (const char* g = getenv("PDC_GLS_UMMA")) ctx->gls_umma = atoi(g);
  if (const char* g = getenv("PDC_GLS_UMMA_NSPLIT")) ctx->gls_umma_nsplit = atoi(g);
  if (const char* g = getenv("PDC_GLS_UMMA_CHUNK")) ctx->gls_umma_chunk = atoi(g);
  if (const char* g = getenv("PDC_GLS_UMMA_DBG")) ctx->gls_umma_dbg = atoi(g);
  if (const char* g = getenv("PDC_GLS_UMMA_FINE")) ctx->gls_umma_fine = atoi(g);
  if (const char* g = getenv("PDC_GLS_UMMA_CG2")) ctx->gls_umma_cg2 = atoi(g);
  if (const char* g = getenv("PDC_GLS_UMMA_RZCOMP")) ctx->gls_umma_rzcomp = atoi(g);
  if (const char* g = getenv("PDC_GLS_UMMA_PROF")) ctx->umma_prof_on = atoi(g) != 0;
  if (const char* g = getenv("PDC_PDM_PPT")) ctx->pdm_ppt_override = atoi(g);
  if (const char* g = getenv("PDC_BATCH_PIPE_BYTES")) ctx->pipe_min_bytes = (size_t)atoll(g);
  cudaError_t e1 = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  cudaError_t e4 = cudaEventCreateWithFlags(&ctx->ev_fence, cudaEventDisableTiming);
  if (e4 == cudaSuccess) e4 = cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming);
  if (e1 != cudaSuccess || e4 != cudaSuccess) {
    pdc_ctx_destroy(ctx);
    return cuda_fail(e1 != cudaSuccess ? e1 : e4, "stream/event creation", __FILE__, __LINE__);
  }
  *out = ctx;
  return PDC_OK;
}

int pdc_ctx_destroy(pdc_ctx* ctx) {
  if (!ctx) return PDC_OK;
  multi_destroy(ctx);   // stops the worker threads and destroys the child ctxs of a multi-device ctx
  DeviceGuard guard(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  ctx->in_a.release(); ctx->in_b.release(); ctx->in_c.release(); ctx->in_d.release();
  ctx->out_a.release(); ctx->out_small.release(); ctx->pin_small.release(); ctx->pin_out.release();
  for (auto& e : ctx->ev_chunk) if (e) cudaEventDestroy(e);
  for (auto& e : ctx->ev_up) if (e) cudaEventDestroy(e);
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  ctx->gls_curves.release(); ctx->gls_part.release(); ctx->gls_rec1.release(); ctx->gls_rec2.release(); ctx->gls_low.release(); ctx->gls_cnt.release(); ctx->glsm_y.release();
  ctx->partial.release(); ctx->gls_plane.release(); ctx->hist_plane.release(); ctx->blockred.release(); ctx->pin_meta.release();
  ctx->pdm_meta.release(); ctx->pdm_cnt.release(); ctx->gl_acc.release(); ctx->peak_cand.release(); ctx->umma_status.release(); ctx->umma_prof.release(); ctx->umma_fine.release();
  ctx->main_resolve();
  for (auto& pr : ctx->ev_free) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  if (ctx->ev_fence) cudaEventDestroy(ctx->ev_fence);
  if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return PDC_OK;
}

int pdc_ctx_synchronize(pdc_ctx* ctx) {
  if (!ctx) { set_error("ctx is NULL"); return PDC_EINVAL; }
  DeviceGuard guard(ctx->device);
  PDC_CUDA(cudaStreamSynchronize(ctx->stream));
  return PDC_OK;
}

int pdc_ctx_sm_count(pdc_ctx* ctx) { return ctx ? ctx->sm_count : -1; }

int64_t pdc_ctx_launch_count(pdc_ctx* ctx) { return ctx ? ctx->launches : -1; }  /* primary device; see pdc_ctx_device_count */

double pdc_ctx_last_main_kernel_ms(pdc_ctx* ctx) {
  if (!ctx) return -1.0;
  DeviceGuard guard(ctx->device);
  if (ctx->main_resolve() != PDC_OK) return -1.0;
  return ctx->main_ms_last;
}

double pdc_ctx_main_kernel_ms_total(pdc_ctx* ctx, int64_t* count_out) {
  if (!ctx) return -1.0;
  DeviceGuard guard(ctx->device);
  if (ctx->main_resolve() != PDC_OK) return -1.0;
  if (count_out) *count_out = ctx->main_count;
  return ctx->main_ms_total;
}

int pdc_ctx_last_gls_path(pdc_ctx* ctx) { return ctx ? ctx->last_gls_path : -1; }

int pdc_debug_umma_plan(int sm_count, int64_t B, int64_t nf, int64_t nmax, int fine, int cg2, int nsplit, int chunk, int64_t* out) {
  if (!out || sm_count < 2 || B < 1 || nf < 1 || nmax < 1) return PDC_EINVAL;
  pdc::GlsUmmaPlan p;
  pdc::gls_umma_plan(sm_count, B, nf, nmax, pdc::GlsUmmaKnobs{fine, cg2, nsplit, chunk}, &p);
  out[0] = p.path; out[1] = p.fine; out[2] = p.nC; out[3] = p.nt1; out[4] = p.cpt1; out[5] = p.nt2; out[6] = p.cpt2;
  out[7] = p.nsplit; out[8] = p.chunk_stages; out[9] = p.jobs; out[10] = p.fine_bytes;
  return PDC_OK;
}

int64_t pdc_debug_umma_prof(pdc_ctx* ctx, int64_t* out, int64_t cap) {
  if (!ctx) return -1;
  DeviceGuard guard(ctx->device);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return -1;
  /* the per-job records are followed by the 1024 x 8 trace stamps of block 0 (PDC_GLS_UMMA_DBG & 16) */
  const int64_t have = ctx->umma_prof_jobs > 0 ? ctx->umma_prof_jobs + 2 * 1024 : 0;
  const int64_t n = have < cap ? have : cap;
  if (n > 0 && out &&
      cudaMemcpy(out, ctx->umma_prof.p, sizeof(long long) * 4 * (size_t)n, cudaMemcpyDeviceToHost) != cudaSuccess)
    return -1;
  return have;
}

struct SmallRec {
  long long arg;
  double val;
};

// ---------------------------------------------------------------------------
// GLS
// ---------------------------------------------------------------------------
int pdc_gls_batch_dev(pdc_ctx* ctx, const double* t, const double* y, const double* w,
                      const int64_t* offsets, int64_t B, const double* fmin, const double* df,
                      int64_t nf, unsigned flags, const double* psd_scale,
                      double* power_out, int64_t* argmax_out, double* max_out, void* stream) {
  if (!ctx || !t || !y || !offsets || !fmin || !df) { set_error("pdc_gls_batch_dev: NULL argument"); return PDC_EINVAL; }
  DeviceGuard guard(ctx->device);
  cudaStream_t st = stream == PDC_STREAM_CTX ? ctx->stream : (cudaStream_t)stream;
  return gls_run(ctx, t, y, w, offsets, B, fmin, df, 0, nf, flags, psd_scale, power_out, argmax_out, max_out, st);
}

int pdc_gls_dev(pdc_ctx* ctx, const double* t, const double* y, const double* w, int64_t n,
                double fmin, double df, int64_t j0, int64_t nf, unsigned flags, double psd_scale,
                double* power_out, int64_t* argmax_out, double* max_out, void* stream) {
  if (!ctx || !t || !y) { set_error("pdc_gls_dev: NULL argument"); return PDC_EINVAL; }
  if (n < 1) { set_error("pdc_gls: n must be >= 1"); return PDC_EINVAL; }
  if (j0 < 0) { set_error("pdc_gls: j0 must be >= 0"); return PDC_EINVAL; }
  DeviceGuard guard(ctx->device);
  cudaStream_t st = stream == PDC_STREAM_CTX ? ctx->stream : (cudaStream_t)stream;
  const int64_t offsets[2] = {0, n};
  return gls_run(ctx, t, y, w, offsets, 1, &fmin, &df, j0, nf, flags, &psd_scale, power_out, argmax_out, max_out, st);
}

int pdc_gls_dev_fanout(pdc_ctx* ctx, const double* t, const double* y, const double* w, int64_t n,
                       double fmin, double df, int64_t j0, int64_t nf, unsigned flags, double psd_scale,
                       const pdc_fanout* dst, void* stream) {
  if (!ctx || !t || !y || !dst) { set_error("pdc_gls_dev_fanout: NULL argument"); return PDC_EINVAL; }
  if (n < 1) { set_error("pdc_gls: n must be >= 1"); return PDC_EINVAL; }
  if (j0 < 0) { set_error("pdc_gls: j0 must be >= 0"); return PDC_EINVAL; }
  for (int r = 0; r < dst->world && r < PDC_MAX_PEERS; ++r)
    if (!dst->power[r] || !dst->best[r]) { set_error("pdc_gls_dev_fanout: NULL destination for rank %d", r); return PDC_EINVAL; }
  DeviceGuard guard(ctx->device);
  cudaStream_t st = stream == PDC_STREAM_CTX ? ctx->stream : (cudaStream_t)stream;
  const int64_t offsets[2] = {0, n};
  return gls_run(ctx, t, y, w, offsets, 1, &fmin, &df, j0, nf, flags, &psd_scale, nullptr, nullptr, nullptr, st, dst);
}

// Device -> caller's host buffer.  A cudaMemcpyAsync into pageable memory is staged by the driver at ~10 GB/s
// (0.8 MB periodogram: 86 us); going through our own pinned buffer in a few chunks, each copied out by the CPU
// while the next one is in flight, takes about half of that.  Returns with the data in `dst`.
static int staged_d2h(pdc_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (bytes < ((size_t)128 << 10) || bytes > ((size_t)256 << 20)) {
    PDC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    PDC_CUDA(cudaStreamSynchronize(st));
    return PDC_OK;
  }
  PDC_TRY(ctx->pin_out.reserve(bytes));
  int nchunk = (int)(bytes / ((size_t)256 << 10));
  if (nchunk < 2) nchunk = 2;
  if (nchunk > 8) nchunk = 8;
  const size_t step = (((bytes + nchunk - 1) / nchunk) + 4095) & ~(size_t)4095;  // ceil: at most nchunk <= 8 chunks
  char* pin = ctx->pin_out.as<char>();
  int used = 0;
  for (size_t off = 0; off < bytes && used < 8; off += step, ++used) {
    const size_t len = off + step < bytes ? step : bytes - off;
    if (!ctx->ev_chunk[used]) PDC_CUDA(cudaEventCreateWithFlags(&ctx->ev_chunk[used], cudaEventDisableTiming));
    PDC_CUDA(cudaMemcpyAsync(pin + off, (const char*)src + off, len, cudaMemcpyDeviceToHost, st));
    PDC_CUDA(cudaEventRecord(ctx->ev_chunk[used], st));
  }
  int c = 0;
  for (size_t off = 0; off < bytes; off += step, ++c) {
    const size_t len = off + step < bytes ? step : bytes - off;
    PDC_CUDA(cudaEventSynchronize(ctx->ev_chunk[c]));
    memcpy((char*)dst + off, pin + off, len);
  }
  return PDC_OK;
}

// Upload + compute of a batch in `nrun` runs of whole curves (boundaries balanced by sample count).  The upload of run
// c+1 is issued after the kernels of run c have been enqueued: with pinned host memory it is asynchronous anyway, with
// pageable memory cudaMemcpyAsync keeps the host busy staging while the GPU computes.  `off` is rebased to 0.
static int gls_upload_and_run_pipelined(pdc_ctx* ctx, const double* t, const double* y, const double* w,
                                        double* d_t, double* d_y, double* d_w, const int64_t* off, int64_t B, int nrun,
                                        const double* fmin, const double* df, int64_t nf, unsigned flags,
                                        const double* psd_scale, double* d_power, long long* d_arg, double* d_val,
                                        cudaStream_t st) {
  if (!ctx->copy_stream) PDC_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (int c = 0; c < nrun; ++c)
    if (!ctx->ev_up[c]) PDC_CUDA(cudaEventCreateWithFlags(&ctx->ev_up[c], cudaEventDisableTiming));
  // first curve of each run
  int64_t first[9];
  first[0] = 0;
  const int64_t ntot = off[B];
  for (int c = 1; c < nrun; ++c) {
    const int64_t target = ntot / nrun * c;
    int64_t b = first[c - 1] + 1;
    while (b < B - (nrun - c) && off[b] < target) ++b;
    first[c] = b;
  }
  first[nrun] = B;
  auto upload = [&](int c) -> int {
    const int64_t s0 = off[first[c]], s1 = off[first[c + 1]];
    const size_t bytes = sizeof(double) * (size_t)(s1 - s0);
    PDC_CUDA(cudaMemcpyAsync(d_t + s0, t + s0, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    PDC_CUDA(cudaMemcpyAsync(d_y + s0, y + s0, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    if (w) PDC_CUDA(cudaMemcpyAsync(d_w + s0, w + s0, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    PDC_CUDA(cudaEventRecord(ctx->ev_up[c], ctx->copy_stream));
    return PDC_OK;
  };
  int rc = upload(0);
  for (int c = 0; c < nrun && rc == PDC_OK; ++c) {
    const int64_t b0 = first[c], nb = first[c + 1] - first[c];
    cudaError_t e = cudaStreamWaitEvent(st, ctx->ev_up[c], 0);
    if (e != cudaSuccess) { rc = cuda_fail(e, "cudaStreamWaitEvent", __FILE__, __LINE__); break; }
    rc = gls_run(ctx, d_t, d_y, d_w, off + b0, nb, fmin + b0, df + b0, 0, nf, flags, psd_scale ? psd_scale + b0 : nullptr,
                 d_power ? d_power + (size_t)b0 * nf : nullptr, (int64_t*)(d_arg + b0), d_val + b0, st);
    if (rc == PDC_OK && c + 1 < nrun) rc = upload(c + 1);
  }
  if (rc != PDC_OK) cudaStreamSynchronize(ctx->copy_stream);  // never return with a copy from the caller's memory in flight
  return rc;
}

static int gls_host_common(pdc_ctx* ctx, const double* t, const double* y, const double* w,
                           const int64_t* offsets, int64_t B, const double* fmin, const double* df,
                           int64_t j0, int64_t nf, unsigned flags, const double* psd_scale,
                           double* power_out, int64_t* argmax_out, double* max_out) {
  if (B < 1 || nf < 1) { set_error("pdc_gls: need at least one curve and one frequency"); return PDC_EINVAL; }
  const int64_t off0 = offsets[0];
  const int64_t ntot = offsets[B] - off0;
  if (ntot < 1) { set_error("pdc_gls: no samples"); return PDC_EINVAL; }
  cudaStream_t st = ctx->stream;
  const size_t nbytes = sizeof(double) * (size_t)ntot;
  PDC_TRY(ctx->in_a.reserve(nbytes));
  PDC_TRY(ctx->in_b.reserve(nbytes));
  if (w) PDC_TRY(ctx->in_c.reserve(nbytes));
  if (power_out) PDC_TRY(ctx->out_a.reserve(sizeof(double) * (size_t)nf * B));
  PDC_TRY(ctx->out_small.reserve((sizeof(long long) + sizeof(double)) * (size_t)B));
  PDC_TRY(ctx->pin_small.reserve((sizeof(long long) + sizeof(double)) * (size_t)B));
  // offsets rebased to the device copies
  int64_t local_off[2];
  const int64_t* use_off = offsets;
  std::unique_ptr<int64_t[]> heap_off;
  if (off0 != 0) {
    if (B == 1) { local_off[0] = 0; local_off[1] = ntot; use_off = local_off; }
    else {
      heap_off.reset(new (std::nothrow) int64_t[B + 1]);
      if (!heap_off) { set_error("out of host memory"); return PDC_ENOMEM; }
      for (int64_t b = 0; b <= B; ++b) heap_off[b] = offsets[b] - off0;
      use_off = heap_off.get();
    }
  }
  long long* d_arg = ctx->out_small.as<long long>();
  double* d_val = reinterpret_cast<double*>(d_arg + B);
  double* d_t = ctx->in_a.as<double>();
  double* d_y = ctx->in_b.as<double>();
  double* d_w = w ? ctx->in_c.as<double>() : nullptr;

  // Survey-sized batches: the curves are independent, so the batch is cut into up to 8 runs of whole curves;
  // run c+1 is uploaded on a second stream while the kernels of run c execute.
  int nrun = 1;
  const size_t in_bytes = nbytes * (w ? 3 : 2);
  if (B >= 16 && in_bytes >= ctx->pipe_min_bytes) {
    nrun = (int)(in_bytes / ((size_t)16 << 20));
    if (nrun > 8) nrun = 8;
    if (nrun > B / 8) nrun = (int)(B / 8);
    if (nrun < 2) nrun = 2;
  }
  int rc = PDC_OK;
  if (nrun == 1) {
    PDC_CUDA(cudaMemcpyAsync(d_t, t + off0, nbytes, cudaMemcpyHostToDevice, st));
    PDC_CUDA(cudaMemcpyAsync(d_y, y + off0, nbytes, cudaMemcpyHostToDevice, st));
    if (w) PDC_CUDA(cudaMemcpyAsync(d_w, w + off0, nbytes, cudaMemcpyHostToDevice, st));
    rc = gls_run(ctx, d_t, d_y, d_w, use_off, B, fmin, df, j0, nf, flags, psd_scale,
                 power_out ? ctx->out_a.as<double>() : nullptr, (int64_t*)d_arg, d_val, st);
  } else {
    rc = gls_upload_and_run_pipelined(ctx, t + off0, y + off0, w ? w + off0 : nullptr, d_t, d_y, d_w, use_off, B, nrun,
                                      fmin, df, nf, flags, psd_scale, power_out ? ctx->out_a.as<double>() : nullptr,
                                      d_arg, d_val, st);
  }
  PDC_TRY(rc);
  PDC_CUDA(cudaMemcpyAsync(ctx->pin_small.p, ctx->out_small.p, (sizeof(long long) + sizeof(double)) * (size_t)B,
                           cudaMemcpyDeviceToHost, st));
  if (power_out) PDC_TRY(staged_d2h(ctx, power_out, ctx->out_a.p, sizeof(double) * (size_t)nf * B, st));
  PDC_CUDA(cudaStreamSynchronize(st));
  const long long* h_arg = ctx->pin_small.as<long long>();
  const double* h_val = reinterpret_cast<const double*>(h_arg + B);
  for (int64_t b = 0; b < B; ++b) {
    if (argmax_out) argmax_out[b] = h_arg[b];
    if (max_out) max_out[b] = h_val[b];
  }
  return PDC_OK;
}

int pdc_gls(pdc_ctx* ctx, const double* t, const double* y, const double* w, int64_t n,
            double fmin, double df, int64_t j0, int64_t nf, unsigned flags, double psd_scale,
            double* power_out, int64_t* argmax_out, double* max_out) {
  if (!ctx || !t || !y) { set_error("pdc_gls: NULL argument"); return PDC_EINVAL; }
  if (n < 1) { set_error("pdc_gls: n must be >= 1"); return PDC_EINVAL; }
  if (j0 < 0) { set_error("pdc_gls: j0 must be >= 0"); return PDC_EINVAL; }
  if (nf < 1) { set_error("pdc_gls: need at least one curve and one frequency"); return PDC_EINVAL; }
  if (ctx->multi) return multi_gls(ctx, t, y, w, n, fmin, df, j0, nf, flags, psd_scale, power_out, argmax_out, max_out);
  DeviceGuard guard(ctx->device);
  const int64_t offsets[2] = {0, n};
  return gls_host_common(ctx, t, y, w, offsets, 1, &fmin, &df, j0, nf, flags, &psd_scale, power_out, argmax_out, max_out);
}

int pdc_gls_batch(pdc_ctx* ctx, const double* t, const double* y, const double* w,
                  const int64_t* offsets, int64_t B, const double* fmin, const double* df,
                  int64_t nf, unsigned flags, const double* psd_scale,
                  double* power_out, int64_t* argmax_out, double* max_out) {
  if (!ctx || !t || !y || !offsets || !fmin || !df) { set_error("pdc_gls_batch: NULL argument"); return PDC_EINVAL; }
  if (ctx->multi && B >= 1 && nf >= 1 && offsets[B] > offsets[0])
    return multi_gls_batch(ctx, t, y, w, offsets, B, fmin, df, nf, flags, psd_scale, power_out, argmax_out, max_out);
  DeviceGuard guard(ctx->device);
  return gls_host_common(ctx, t, y, w, offsets, B, fmin, df, 0, nf, flags, psd_scale, power_out, argmax_out, max_out);
}

// GLS at an arbitrary list of frequencies (non-uniform / user-supplied grids)
int pdc_gls_freqs_dev(pdc_ctx* ctx, const double* t, const double* y, const double* w, int64_t n,
                      const double* freqs, int64_t nfreq, unsigned flags, double psd_scale,
                      double* power_out, int64_t* argmax_out, double* max_out, void* stream) {
  if (!ctx || !t || !y || !freqs) { set_error("pdc_gls_freqs_dev: NULL argument"); return PDC_EINVAL; }
  if (n < 1 || nfreq < 1) { set_error("pdc_gls_freqs: need n >= 1 samples and nfreq >= 1 frequencies"); return PDC_EINVAL; }
  DeviceGuard guard(ctx->device);
  cudaStream_t st = stream == PDC_STREAM_CTX ? ctx->stream : (cudaStream_t)stream;
  const int64_t offsets[2] = {0, n};
  const double zero = 0.0;
  return gls_run(ctx, t, y, w, offsets, 1, &zero, &zero, 0, nfreq, flags, &psd_scale, power_out, argmax_out, max_out, st,
                 nullptr, freqs);
}

int pdc_gls_freqs(pdc_ctx* ctx, const double* t, const double* y, const double* w, int64_t n,
                  const double* freqs, int64_t nfreq, unsigned flags, double psd_scale,
                  double* power_out, int64_t* argmax_out, double* max_out) {
  if (!ctx || !t || !y || !freqs || !power_out) { set_error("pdc_gls_freqs: NULL argument"); return PDC_EINVAL; }
  if (n < 1 || nfreq < 1) { set_error("pdc_gls_freqs: need n >= 1 samples and nfreq >= 1 frequencies"); return PDC_EINVAL; }
  if (ctx->multi)
    return multi_period_grid(ctx, n, freqs, nfreq, +1, power_out, argmax_out, max_out,
                             [=](pdc_ctx* c, const double* f, int64_t k, double* o, int64_t* a, double* b) {
                               return pdc_gls_freqs(c, t, y, w, n, f, k, flags, psd_scale, o, a, b);
                             });
  DeviceGuard guard(ctx->device);
  cudaStream_t st = ctx->stream;
  const size_t nbytes = sizeof(double) * (size_t)n;
  PDC_TRY(ctx->in_a.reserve(nbytes));
  PDC_TRY(ctx->in_b.reserve(nbytes));
  if (w) PDC_TRY(ctx->in_c.reserve(nbytes));
  PDC_TRY(ctx->in_d.reserve(sizeof(double) * (size_t)nfreq));
  PDC_TRY(ctx->out_a.reserve(sizeof(double) * (size_t)nfreq));
  PDC_TRY(ctx->out_small.reserve(sizeof(SmallRec)));
  PDC_TRY(ctx->pin_small.reserve(sizeof(SmallRec)));
  PDC_CUDA(cudaMemcpyAsync(ctx->in_a.p, t, nbytes, cudaMemcpyHostToDevice, st));
  PDC_CUDA(cudaMemcpyAsync(ctx->in_b.p, y, nbytes, cudaMemcpyHostToDevice, st));
  if (w) PDC_CUDA(cudaMemcpyAsync(ctx->in_c.p, w, nbytes, cudaMemcpyHostToDevice, st));
  PDC_CUDA(cudaMemcpyAsync(ctx->in_d.p, freqs, sizeof(double) * (size_t)nfreq, cudaMemcpyHostToDevice, st));
  SmallRec* d_rec = ctx->out_small.as<SmallRec>();
  const int64_t offsets[2] = {0, n};
  const double zero = 0.0;
  PDC_TRY(gls_run(ctx, ctx->in_a.as<double>(), ctx->in_b.as<double>(), w ? ctx->in_c.as<double>() : nullptr, offsets, 1,
                  &zero, &zero, 0, nfreq, flags, &psd_scale, ctx->out_a.as<double>(), (int64_t*)&d_rec->arg, &d_rec->val,
                  st, nullptr, ctx->in_d.as<double>()));
  PDC_CUDA(cudaMemcpyAsync(ctx->pin_small.p, d_rec, sizeof(SmallRec), cudaMemcpyDeviceToHost, st));
  PDC_TRY(staged_d2h(ctx, power_out, ctx->out_a.p, sizeof(double) * (size_t)nfreq, st));
  PDC_CUDA(cudaStreamSynchronize(st));
  const SmallRec* h = ctx->pin_small.as<SmallRec>();
  if (argmax_out) *argmax_out = h->arg;
  if (max_out) *max_out = h->val;
  return PDC_OK;
}

int pdc_gls_multi_dev(pdc_ctx* ctx, const double* t, const double* Y, const double* w, int64_t n, int64_t S,
                      double fmin, double df, int64_t j0, int64_t nf, unsigned flags, double psd_scale,
                      double* power_out, int64_t* argmax_out, double* max_out, void* stream) {
  if (!ctx || !t || !Y) { set_error("pdc_gls_multi_dev: NULL argument"); return PDC_EINVAL; }
  if (j0 < 0) { set_error("pdc_gls_multi: j0 must be >= 0"); return PDC_EINVAL; }
  DeviceGuard guard(ctx->device);
  cudaStream_t st = stream == PDC_STREAM_CTX ? ctx->stream : (cudaStream_t)stream;
  return glsm_run(ctx, t, Y, w, n, S, fmin, df, j0, nf, flags, psd_scale, power_out, argmax_out, max_out, st);
}

int pdc_gls_multi(pdc_ctx* ctx, const double* t, const double* Y, const double* w, int64_t n, int64_t S,
                  double fmin, double df, int64_t j0, int64_t nf, unsigned flags, double psd_scale,
                  double* power_out, int64_t* argmax_out, double* max_out) {
  if (!ctx || !t || !Y) { set_error("pdc_gls_multi: NULL argument"); return PDC_EINVAL; }
  if (n < 1 || S < 1 || nf < 1) { set_error("pdc_gls_multi: need n, S, nf >= 1"); return PDC_EINVAL; }
  if (j0 < 0) { set_error("pdc_gls_multi: j0 must be >= 0"); return PDC_EINVAL; }
  if (ctx->multi)
    return multi_gls_multi(ctx, t, Y, w, n, S, fmin, df, j0, nf, flags, psd_scale, power_out, argmax_out, max_out);
  DeviceGuard guard(ctx->device);
  cudaStream_t st = ctx->stream;
  const size_t tb = sizeof(double) * (size_t)n, yb = tb * (size_t)S;
  PDC_TRY(ctx->in_a.reserve(tb));
  PDC_TRY(ctx->in_b.reserve(yb));
  if (w) PDC_TRY(ctx->in_c.reserve(tb));
  if (power_out) PDC_TRY(ctx->out_a.reserve(sizeof(double) * (size_t)nf * S));
  PDC_TRY(ctx->out_small.reserve((sizeof(long long) + sizeof(double)) * (size_t)S));
  PDC_TRY(ctx->pin_small.reserve((sizeof(long long) + sizeof(double)) * (size_t)S));
  PDC_CUDA(cudaMemcpyAsync(ctx->in_a.p, t, tb, cudaMemcpyHostToDevice, st));
  PDC_CUDA(cudaMemcpyAsync(ctx->in_b.p, Y, yb, cudaMemcpyHostToDevice, st));
  if (w) PDC_CUDA(cudaMemcpyAsync(ctx->in_c.p, w, tb, cudaMemcpyHostToDevice, st));
  long long* d_arg = ctx->out_small.as<long long>();
  double* d_val = reinterpret_cast<double*>(d_arg + S);
  PDC_TRY(glsm_run(ctx, ctx->in_a.as<double>(), ctx->in_b.as<double>(), w ? ctx->in_c.as<double>() : nullptr, n, S,
                   fmin, df, j0, nf, flags, psd_scale, power_out ? ctx->out_a.as<double>() : nullptr,
                   (int64_t*)d_arg, d_val, st));
  PDC_CUDA(cudaMemcpyAsync(ctx->pin_small.p, ctx->out_small.p, (sizeof(long long) + sizeof(double)) * (size_t)S,
                           cudaMemcpyDeviceToHost, st));
  if (power_out) PDC_TRY(staged_d2h(ctx, power_out, ctx->out_a.p, sizeof(double) * (size_t)nf * S, st));
  PDC_CUDA(cudaStreamSynchronize(st));
  const long long* h_arg = ctx->pin_small.as<long long>();
  const double* h_val = reinterpret_cast<const double*>(h_arg + S);
  for (int64_t s = 0; s < S; ++s) {
    if (argmax_out) argmax_out[s] = h_arg[s];
    if (max_out) max_out[s] = h_val[s];
  }
  return PDC_OK;
}

// ---------------------------------------------------------------------------
// PDM
// ---------------------------------------------------------------------------
int pdc_pdm_dev(pdc_ctx* ctx, const double* t, const double* x, int64_t n,
                const double* periods, int64_t np, int nb, int nc,
                double* theta_out, int64_t* argmin_out, double* min_out, void* stream) {
  if (!ctx || !t || !x || !periods || !theta_out) { set_error("pdc_pdm_dev: NULL argument"); return PDC_EINVAL; }
  DeviceGuard guard(ctx->device);
  cudaStream_t st = stream == PDC_STREAM_CTX ? ctx->stream : (cudaStream_t)stream;
  return pdm_run(ctx, t, x, n, periods, np, nb, nc, theta_out, argmin_out, min_out, st);
}

int pdc_pdm_dev_fanout(pdc_ctx* ctx, const double* t, const double* x, int64_t n, const double* periods,
                       int64_t np, int nb, int nc, int64_t offset, const pdc_fanout* dst, void* stream) {
  if (!ctx || !t || !x || !periods || !dst) { set_error("pdc_pdm_dev_fanout: NULL argument"); return PDC_EINVAL; }
  if (offset < 0) { set_error("pdc_pdm_dev_fanout: offset must be >= 0"); return PDC_EINVAL; }
  for (int r = 0; r < dst->world && r < PDC_MAX_PEERS; ++r)
    if (!dst->power[r] || !dst->best[r]) { set_error("pdc_pdm_dev_fanout: NULL destination for rank %d", r); return PDC_EINVAL; }
  DeviceGuard guard(ctx->device);
  cudaStream_t st = stream == PDC_STREAM_CTX ? ctx->stream : (cudaStream_t)stream;
  return pdm_run(ctx, t, x, n, periods, np, nb, nc, nullptr, nullptr, nullptr, st, dst, offset);
}

// host-pointer body shared by pdc_pdm and pdc_aov: upload, histogram kernels + epilogue for `statistic`, download
static int phase_hist_host(pdc_ctx* ctx, const double* t, const double* x, int64_t n, const double* periods,
                           int64_t np, int nb, int nc, int statistic, double* stat_out, int64_t* arg_out,
                           double* best_out) {
  DeviceGuard guard(ctx->device);
  cudaStream_t st = ctx->stream;
  PDC_TRY(ctx->in_a.reserve(sizeof(double) * (size_t)n));
  PDC_TRY(ctx->in_b.reserve(sizeof(double) * (size_t)n));
  PDC_TRY(ctx->in_d.reserve(sizeof(double) * (size_t)np));
  PDC_TRY(ctx->out_a.reserve(sizeof(double) * (size_t)np));
  PDC_TRY(ctx->out_small.reserve(sizeof(SmallRec)));
  PDC_TRY(ctx->pin_small.reserve(sizeof(SmallRec)));
  PDC_CUDA(cudaMemcpyAsync(ctx->in_a.p, t, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
  PDC_CUDA(cudaMemcpyAsync(ctx->in_b.p, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
  PDC_CUDA(cudaMemcpyAsync(ctx->in_d.p, periods, sizeof(double) * (size_t)np, cudaMemcpyHostToDevice, st));
  SmallRec* d_rec = ctx->out_small.as<SmallRec>();
  PDC_TRY(pdm_run(ctx, ctx->in_a.as<double>(), ctx->in_b.as<double>(), n, ctx->in_d.as<double>(), np, nb, nc,
                  ctx->out_a.as<double>(), (int64_t*)&d_rec->arg, &d_rec->val, st, nullptr, 0, statistic));
  PDC_CUDA(cudaMemcpyAsync(ctx->pin_small.p, d_rec, sizeof(SmallRec), cudaMemcpyDeviceToHost, st));
  PDC_TRY(staged_d2h(ctx, stat_out, ctx->out_a.p, sizeof(double) * (size_t)np, st));
  PDC_CUDA(cudaStreamSynchronize(st));
  const SmallRec* h = ctx->pin_small.as<SmallRec>();
  if (arg_out) *arg_out = h->arg;
  if (best_out) *best_out = h->val;
  return PDC_OK;
}

int pdc_pdm(pdc_ctx* ctx, const double* t, const double* x, int64_t n,
            const double* periods, int64_t np, int nb, int nc,
            double* theta_out, int64_t* argmin_out, double* min_out) {
  if (!ctx || !t || !x || !periods || !theta_out) { set_error("pdc_pdm: NULL argument"); return PDC_EINVAL; }
  if (n < 2 || np < 1) { set_error("pdc_pdm: need n >= 2 samples and np >= 1 periods"); return PDC_EINVAL; }
  if (ctx->multi)
    return multi_period_grid(ctx, n, periods, np, -1, theta_out, argmin_out, min_out,
                             [=](pdc_ctx* c, const double* p, int64_t k, double* o, int64_t* a, double* b) {
                               return pdc_pdm(c, t, x, n, p, k, nb, nc, o, a, b);
                             });
  return phase_hist_host(ctx, t, x, n, periods, np, nb, nc, PDC_STAT_PDM, theta_out, argmin_out, min_out);
}

// ---------------------------------------------------------------------------
// analysis of variance (same histograms as PDM, nc = 1)
// ---------------------------------------------------------------------------
int pdc_aov_dev(pdc_ctx* ctx, const double* t, const double* x, int64_t n, const double* periods, int64_t np,
                int nb, double* theta_out, int64_t* argmax_out, double* max_out, void* stream) {
  if (!ctx || !t || !x || !periods || !theta_out) { set_error("pdc_aov_dev: NULL argument"); return PDC_EINVAL; }
  if (nb < 2) { set_error("pdc_aov: needs at least 2 phase bins"); return PDC_EINVAL; }
  DeviceGuard guard(ctx->device);
  cudaStream_t st = stream == PDC_STREAM_CTX ? ctx->stream : (cudaStream_t)stream;
  return pdm_run(ctx, t, x, n, periods, np, nb, 1, theta_out, argmax_out, max_out, st, nullptr, 0, PDC_STAT_AOV);
}

int pdc_aov(pdc_ctx* ctx, const double* t, const double* x, int64_t n, const double* periods, int64_t np,
            int nb, double* theta_out, int64_t* argmax_out, double* max_out) {
  if (!ctx || !t || !x || !periods || !theta_out) { set_error("pdc_aov: NULL argument"); return PDC_EINVAL; }
  if (n < 2 || np < 1) { set_error("pdc_aov: need n >= 2 samples and np >= 1 periods"); return PDC_EINVAL; }
  if (nb < 2) { set_error("pdc_aov: needs at least 2 phase bins"); return PDC_EINVAL; }
  if (ctx->multi)
    return multi_period_grid(ctx, n, periods, np, +1, theta_out, argmax_out, max_out,
                             [=](pdc_ctx* c, const double* p, int64_t k, double* o, int64_t* a, double* b) {
                               return pdc_aov(c, t, x, n, p, k, nb, o, a, b);
                             });
  return phase_hist_host(ctx, t, x, n, periods, np, nb, 1, PDC_STAT_AOV, theta_out, argmax_out, max_out);
}

// ---------------------------------------------------------------------------
// conditional entropy (count histograms over phase x magnitude cells)
// ---------------------------------------------------------------------------
int pdc_ce_dev(pdc_ctx* ctx, const double* t, const double* x, int64_t n, const double* periods, int64_t np,
               int nphi, int nm, double* h_out, int64_t* argmin_out, double* min_out, void* stream) {
  if (!ctx || !t || !x || !periods || !h_out) { set_error("pdc_ce_dev: NULL argument"); return PDC_EINVAL; }
  DeviceGuard guard(ctx->device);
  cudaStream_t st = stream == PDC_STREAM_CTX ? ctx->stream : (cudaStream_t)stream;
  return ce_run(ctx, t, x, n, periods, np, nphi, nm, h_out, argmin_out, min_out, st);
}

int pdc_ce(pdc_ctx* ctx, const double* t, const double* x, int64_t n, const double* periods, int64_t np,
           int nphi, int nm, double* h_out, int64_t* argmin_out, double* min_out) {
  if (!ctx || !t || !x || !periods || !h_out) { set_error("pdc_ce: NULL argument"); return PDC_EINVAL; }
  if (n < 1 || np < 1) { set_error("pdc_ce: need n >= 1 samples and np >= 1 periods"); return PDC_EINVAL; }
  if (nphi < 1 || nm < 1) { set_error("pdc_ce: nphi and nm must be >= 1"); return PDC_EINVAL; }
  if (ctx->multi)
    return multi_period_grid(ctx, n, periods, np, -1, h_out, argmin_out, min_out,
                             [=](pdc_ctx* c, const double* p, int64_t k, double* o, int64_t* a, double* b) {
                               return pdc_ce(c, t, x, n, p, k, nphi, nm, o, a, b);
                             });
  DeviceGuard guard(ctx->device);
  cudaStream_t st = ctx->stream;
  PDC_TRY(ctx->in_a.reserve(sizeof(double) * (size_t)n));
  PDC_TRY(ctx->in_b.reserve(sizeof(double) * (size_t)n));
  PDC_TRY(ctx->in_d.reserve(sizeof(double) * (size_t)np));
  PDC_TRY(ctx->out_a.reserve(sizeof(double) * (size_t)np));
  PDC_TRY(ctx->out_small.reserve(sizeof(SmallRec)));
  PDC_TRY(ctx->pin_small.reserve(sizeof(SmallRec)));
  PDC_CUDA(cudaMemcpyAsync(ctx->in_a.p, t, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
  PDC_CUDA(cudaMemcpyAsync(ctx->in_b.p, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
  PDC_CUDA(cudaMemcpyAsync(ctx->in_d.p, periods, sizeof(double) * (size_t)np, cudaMemcpyHostToDevice, st));
  SmallRec* d_rec = ctx->out_small.as<SmallRec>();
  PDC_TRY(ce_run(ctx, ctx->in_a.as<double>(), ctx->in_b.as<double>(), n, ctx->in_d.as<double>(), np, nphi, nm,
                 ctx->out_a.as<double>(), (int64_t*)&d_rec->arg, &d_rec->val, st));
  PDC_CUDA(cudaMemcpyAsync(ctx->pin_small.p, d_rec, sizeof(SmallRec), cudaMemcpyDeviceToHost, st));
  PDC_TRY(staged_d2h(ctx, h_out, ctx->out_a.p, sizeof(double) * (size_t)np, st));
  PDC_CUDA(cudaStreamSynchronize(st));
  const SmallRec* h = ctx->pin_small.as<SmallRec>();
  if (argmin_out) *argmin_out = h->arg;
  if (min_out) *min_out = h->val;
  return PDC_OK;
}

// ---------------------------------------------------------------------------
// Gregory-Loredo (event arrival times; count histograms for m = 2 .. m_max phase bins)
// ---------------------------------------------------------------------------
int pdc_gl_dev(pdc_ctx* ctx, const double* t, int64_t n, const double* periods, int64_t np, int m_max, int nc,
               double* lnodds_out, int64_t* argmax_out, double* max_out, void* stream) {
  if (!ctx || !t || !periods || !lnodds_out) { set_error("pdc_gl_dev: NULL argument"); return PDC_EINVAL; }
  DeviceGuard guard(ctx->device);
  cudaStream_t st = stream == PDC_STREAM_CTX ? ctx->stream : (cudaStream_t)stream;
  return gl_run(ctx, t, n, periods, np, m_max, nc, lnodds_out, argmax_out, max_out, st);
}

int pdc_gl(pdc_ctx* ctx, const double* t, int64_t n, const double* periods, int64_t np, int m_max, int nc,
           double* lnodds_out, int64_t* argmax_out, double* max_out) {
  if (!ctx || !t || !periods || !lnodds_out) { set_error("pdc_gl: NULL argument"); return PDC_EINVAL; }
  if (n < 1 || np < 1) { set_error("pdc_gl: need n >= 1 events and np >= 1 periods"); return PDC_EINVAL; }
  if (m_max < 2 || nc < 1 || (long long)m_max * nc > 4096) {
    set_error("pdc_gl: need m_max >= 2, nc >= 1 and m_max * nc <= 4096");
    return PDC_EINVAL;
  }
  if (ctx->multi)
    return multi_period_grid(ctx, n * (int64_t)(m_max - 1), periods, np, +1, lnodds_out, argmax_out, max_out,
                             [=](pdc_ctx* c, const double* p, int64_t k, double* o, int64_t* a, double* b) {
                               return pdc_gl(c, t, n, p, k, m_max, nc, o, a, b);
                             });
  DeviceGuard guard(ctx->device);
  cudaStream_t st = ctx->stream;
  PDC_TRY(ctx->in_a.reserve(sizeof(double) * (size_t)n));
  PDC_TRY(ctx->in_d.reserve(sizeof(double) * (size_t)np));
  PDC_TRY(ctx->out_a.reserve(sizeof(double) * (size_t)np));
  PDC_TRY(ctx->out_small.reserve(sizeof(SmallRec)));
  PDC_TRY(ctx->pin_small.reserve(sizeof(SmallRec)));
  PDC_CUDA(cudaMemcpyAsync(ctx->in_a.p, t, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
  PDC_CUDA(cudaMemcpyAsync(ctx->in_d.p, periods, sizeof(double) * (size_t)np, cudaMemcpyHostToDevice, st));
  SmallRec* d_rec = ctx->out_small.as<SmallRec>();
  PDC_TRY(gl_run(ctx, ctx->in_a.as<double>(), n, ctx->in_d.as<double>(), np, m_max, nc, ctx->out_a.as<double>(),
                 (int64_t*)&d_rec->arg, &d_rec->val, st));
  PDC_CUDA(cudaMemcpyAsync(ctx->pin_small.p, d_rec, sizeof(SmallRec), cudaMemcpyDeviceToHost, st));
  PDC_TRY(staged_d2h(ctx, lnodds_out, ctx->out_a.p, sizeof(double) * (size_t)np, st));
  PDC_CUDA(cudaStreamSynchronize(st));
  const SmallRec* h = ctx->pin_small.as<SmallRec>();
  if (argmax_out) *argmax_out = h->arg;
  if (max_out) *max_out = h->val;
  return PDC_OK;
}

// ---------------------------------------------------------------------------
// string length
// ---------------------------------------------------------------------------
int pdc_stringlength_dev(pdc_ctx* ctx, const double* t, const double* m, int64_t n, const double* periods,
                         int64_t np, double* ell_out, int64_t* argmin_out, double* min_out, void* stream) {
  if (!ctx || !t || !m || !periods || !ell_out) { set_error("pdc_stringlength_dev: NULL argument"); return PDC_EINVAL; }
  DeviceGuard guard(ctx->device);
  cudaStream_t st = stream == PDC_STREAM_CTX ? ctx->stream : (cudaStream_t)stream;
  return strlen_run(ctx, t, m, n, periods, np, ell_out, argmin_out, min_out, st);
}

int pdc_stringlength(pdc_ctx* ctx, const double* t, const double* m, int64_t n, const double* periods,
                     int64_t np, double* ell_out, int64_t* argmin_out, double* min_out) {
  if (!ctx || !t || !m || !periods || !ell_out) { set_error("pdc_stringlength: NULL argument"); return PDC_EINVAL; }
  if (n < 1 || np < 1) { set_error("pdc_stringlength: need n >= 1 samples and np >= 1 periods"); return PDC_EINVAL; }
  if (ctx->multi)   // the sort makes a sample*period ~ log2(n)^2 / 2 times dearer than a PDM update: weigh the split accordingly
    return multi_period_grid(ctx, n * 64, periods, np, -1, ell_out, argmin_out, min_out,
                             [=](pdc_ctx* c, const double* p, int64_t k, double* o, int64_t* a, double* b) {
                               return pdc_stringlength(c, t, m, n, p, k, o, a, b);
                             });
  DeviceGuard guard(ctx->device);
  cudaStream_t st = ctx->stream;
  PDC_TRY(ctx->in_a.reserve(sizeof(double) * (size_t)n));
  PDC_TRY(ctx->in_b.reserve(sizeof(double) * (size_t)n));
  PDC_TRY(ctx->in_d.reserve(sizeof(double) * (size_t)np));
  PDC_TRY(ctx->out_a.reserve(sizeof(double) * (size_t)np));
  PDC_TRY(ctx->out_small.reserve(sizeof(SmallRec)));
  PDC_TRY(ctx->pin_small.reserve(sizeof(SmallRec)));
  PDC_CUDA(cudaMemcpyAsync(ctx->in_a.p, t, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
  PDC_CUDA(cudaMemcpyAsync(ctx->in_b.p, m, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
  PDC_CUDA(cudaMemcpyAsync(ctx->in_d.p, periods, sizeof(double) * (size_t)np, cudaMemcpyHostToDevice, st));
  SmallRec* d_rec = ctx->out_small.as<SmallRec>();
  PDC_TRY(strlen_run(ctx, ctx->in_a.as<double>(), ctx->in_b.as<double>(), n, ctx->in_d.as<double>(), np,
                     ctx->out_a.as<double>(), (int64_t*)&d_rec->arg, &d_rec->val, st));
  PDC_CUDA(cudaMemcpyAsync(ctx->pin_small.p, d_rec, sizeof(SmallRec), cudaMemcpyDeviceToHost, st));
  PDC_TRY(staged_d2h(ctx, ell_out, ctx->out_a.p, sizeof(double) * (size_t)np, st));
  PDC_CUDA(cudaStreamSynchronize(st));
  const SmallRec* h = ctx->pin_small.as<SmallRec>();
  if (argmin_out) *argmin_out = h->arg;
  if (min_out) *min_out = h->val;
  return PDC_OK;
}

// ---------------------------------------------------------------------------
// peaks
// ---------------------------------------------------------------------------
int pdc_peaks_topk_dev(pdc_ctx* ctx, const double* values, int64_t rows, int64_t n, int k,
                       int64_t* idx_out, double* val_out, void* stream) {
  if (!ctx || !values || !idx_out || !val_out) { set_error("pdc_peaks_topk_dev: NULL argument"); return PDC_EINVAL; }
  DeviceGuard guard(ctx->device);
  cudaStream_t st = stream == PDC_STREAM_CTX ? ctx->stream : (cudaStream_t)stream;
  return peaks_run(ctx, values, rows, n, k, idx_out, val_out, st);
}

int pdc_peaks_topk(pdc_ctx* ctx, const double* values, int64_t rows, int64_t n, int k,
                   int64_t* idx_out, double* val_out) {
  if (!ctx || !values || !idx_out || !val_out) { set_error("pdc_peaks_topk: NULL argument"); return PDC_EINVAL; }
  if (rows < 1 || n < 1 || k < 1) { set_error("pdc_peaks_topk: empty input"); return PDC_EINVAL; }
  DeviceGuard guard(ctx->device);
  cudaStream_t st = ctx->stream;
  const size_t nb = sizeof(double) * (size_t)rows * n, ob = (size_t)rows * k;
  PDC_TRY(ctx->out_a.reserve(nb));
  PDC_TRY(ctx->out_small.reserve(ob * (sizeof(double) + sizeof(int64_t))));
  PDC_CUDA(cudaMemcpyAsync(ctx->out_a.p, values, nb, cudaMemcpyHostToDevice, st));
  int64_t* d_idx = ctx->out_small.as<int64_t>();
  double* d_val = reinterpret_cast<double*>(d_idx + ob);
  PDC_TRY(peaks_run(ctx, ctx->out_a.as<double>(), rows, n, k, d_idx, d_val, st));
  PDC_CUDA(cudaMemcpyAsync(idx_out, d_idx, ob * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  PDC_CUDA(cudaMemcpyAsync(val_out, d_val, ob * sizeof(double), cudaMemcpyDeviceToHost, st));
  PDC_CUDA(cudaStreamSynchronize(st));
  return PDC_OK;
}

int pdc_peaks_halfmax_dev(pdc_ctx* ctx, const double* values, int64_t rows, int64_t n, int k, const int64_t* peak_idx,
                          const double* height, int64_t* left_out, int64_t* right_out, void* stream) {
  if (!ctx || !values || !peak_idx || !left_out || !right_out) { set_error("pdc_peaks_halfmax_dev: NULL argument"); return PDC_EINVAL; }
  DeviceGuard guard(ctx->device);
  cudaStream_t st = stream == PDC_STREAM_CTX ? ctx->stream : (cudaStream_t)stream;
  return peaks_halfmax_run(ctx, values, rows, n, k, peak_idx, height, left_out, right_out, st);
}

int pdc_peaks_halfmax(pdc_ctx* ctx, const double* values, int64_t rows, int64_t n, int k, const int64_t* peak_idx,
                      const double* height, int64_t* left_out, int64_t* right_out) {
  if (!ctx || !values || !peak_idx || !left_out || !right_out) { set_error("pdc_peaks_halfmax: NULL argument"); return PDC_EINVAL; }
  if (rows < 1 || n < 1 || k < 1) { set_error("pdc_peaks_halfmax: empty input"); return PDC_EINVAL; }
  DeviceGuard guard(ctx->device);
  cudaStream_t st = ctx->stream;
  const size_t nb = sizeof(double) * (size_t)rows * n, ob = (size_t)rows * k;
  PDC_TRY(ctx->out_a.reserve(nb));
  PDC_TRY(ctx->out_small.reserve(ob * (3 * sizeof(int64_t) + sizeof(double))));
  int64_t* d_idx = ctx->out_small.as<int64_t>();
  int64_t* d_left = d_idx + ob;
  int64_t* d_right = d_left + ob;
  double* d_h = reinterpret_cast<double*>(d_right + ob);
  PDC_CUDA(cudaMemcpyAsync(ctx->out_a.p, values, nb, cudaMemcpyHostToDevice, st));
  PDC_CUDA(cudaMemcpyAsync(d_idx, peak_idx, ob * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  if (height) PDC_CUDA(cudaMemcpyAsync(d_h, height, ob * sizeof(double), cudaMemcpyHostToDevice, st));
  PDC_TRY(peaks_halfmax_run(ctx, ctx->out_a.as<double>(), rows, n, k, d_idx, height ? d_h : nullptr, d_left, d_right, st));
  PDC_CUDA(cudaMemcpyAsync(left_out, d_left, ob * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  PDC_CUDA(cudaMemcpyAsync(right_out, d_right, ob * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  PDC_CUDA(cudaStreamSynchronize(st));
  return PDC_OK;
}

}  // extern "C"
