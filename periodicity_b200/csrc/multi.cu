// Single-process multi-device contexts (SURVEY.md section 8b `pdc_ctx_create(out, device_ids, ndev)` / 8e).
//
// The trial-frequency search shards trivially: frequencies / trial periods / light curves are independent
// (the reference's counterpart is the multiprocessing.Pool fan-out of phase.py:182-187, transparent to the caller).
// A multi-device ctx owns one ordinary single-device ctx per CUDA device plus one host worker thread per extra
// device.  A HOST-pointer entry point called on it
//   * cuts the grid (pdc_gls, pdc_pdm, pdc_aov, pdc_ce, pdc_stringlength: contiguous slices of the frequency / period
//     grid) or the batch (pdc_gls_batch: contiguous groups of curves balanced by sample count; pdc_gls_multi: groups of
//     series) into one piece per device,
//   * runs the single-device entry point for every piece concurrently -- each worker uploads the inputs to its device,
//     runs the kernels with the slice offset j0, and copies its slice of the result STRAIGHT INTO THE CALLER'S HOST
//     BUFFER, so there is no device-to-device exchange at all (the result is wanted on the host),
//   * reduces the per-device (best value, best index) candidates on the host, NaN-aware with first-occurrence ties,
//     exactly like the device arg-extremum (np.nanargmax / np.nanargmin, core.py:202-210).
// Results are the same values the single-device call produces for each slice (tests/test_multi_device.py).  Problems
// too small to be worth cutting (configs[0]-sized) run on the first device only.
// Device-pointer (`*_dev`) entry points given a multi-device ctx act on its first device: a device pointer belongs to
// one device.  Multi-process sharding (one rank per GPU, NVLink fan-out stores) stays in periodicity_b200/dist.py.
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>

#include "pdc_common.cuh"

namespace pdc {

// One worker thread bound to one child ctx; runs one job at a time.
struct Worker {
  pdc_ctx* ctx = nullptr;
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<int()> job;
  bool has_job = false, done = true, quit = false;
  int rc = PDC_OK;
  std::string err;

  void loop() {
    for (;;) {
      std::function<int()> j;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return has_job || quit; });
        if (quit) return;
        j = std::move(job);
        has_job = false;
      }
      int r = j();
      std::string e = r == PDC_OK ? std::string() : std::string(pdc_last_error());  // thread-local message of THIS thread
      {
        std::lock_guard<std::mutex> lk(mu);
        rc = r;
        err = std::move(e);
        done = true;
      }
      cv.notify_all();
    }
  }
  void submit(std::function<int()> j) {
    {
      std::lock_guard<std::mutex> lk(mu);
      job = std::move(j);
      has_job = true;
      done = false;
    }
    cv.notify_all();
  }
  int wait() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return done; });
    return rc;
  }
};

struct MultiState {
  // One ordinary single-device ctx per entry of device_ids.  devs[0] is a child too (NOT the owning ctx: the owner's
  // `multi` pointer is what routes a host call here, a child must take the single-device path); its piece of the work
  // runs on the calling thread.
  std::vector<pdc_ctx*> devs;
  std::vector<Worker*> workers;    // workers[d - 1] serves devs[d], d >= 1
  long long min_evals_per_device = 500000000LL;  // env PDC_MULTI_MIN_EVALS: below this much work per device, use fewer devices
};

void multi_destroy(pdc_ctx* ctx) {
  MultiState* m = ctx->multi;
  if (!m) return;
  for (Worker* w : m->workers) {
    {
      std::lock_guard<std::mutex> lk(w->mu);
      w->quit = true;
    }
    w->cv.notify_all();
    if (w->th.joinable()) w->th.join();
    delete w;
  }
  for (size_t d = 0; d < m->devs.size(); ++d) pdc_ctx_destroy(m->devs[d]);
  delete m;
  ctx->multi = nullptr;
}

int multi_device_count(const pdc_ctx* ctx) { return ctx->multi ? (int)ctx->multi->devs.size() : 1; }

// Devices to use for `evals` sample*frequency evaluations cut into at most `units` pieces.
static int devices_for(const MultiState* m, double evals, long long units) {
  long long k = (long long)(evals / (double)m->min_evals_per_device);
  if (k < 1) k = 1;
  if (k > (long long)m->devs.size()) k = (long long)m->devs.size();
  if (k > units) k = units;
  return (int)k;
}

// Run fn(d) for d = 0..k-1: d = 0 on the calling thread (primary device), the others on their workers.
// Returns the first failure (its message becomes this thread's pdc_last_error()).
static int fan_out(MultiState* m, int k, const std::function<int(int)>& fn) {
  for (int d = 1; d < k; ++d) m->workers[d - 1]->submit([&fn, d] { return fn(d); });
  int rc = fn(0);
  std::string msg = rc == PDC_OK ? std::string() : std::string(pdc_last_error());
  for (int d = 1; d < k; ++d) {
    const int r = m->workers[d - 1]->wait();
    if (r != PDC_OK && rc == PDC_OK) {
      rc = r;
      msg = "device " + std::to_string(m->devs[d]->device) + ": " + m->workers[d - 1]->err;
    }
  }
  if (rc != PDC_OK) set_error("%s", msg.c_str());
  return rc;
}

// contiguous slice [start, stop) of piece d of k over n units
static inline void slice_of(long long n, int k, int d, long long& start, long long& stop) {
  const long long L = (n + k - 1) / k;
  start = (long long)d * L < n ? (long long)d * L : n;
  stop = start + L < n ? start + L : n;
}

// host reduction of per-device candidates; SIGN +1 max, -1 min; NaN ignored, first occurrence wins
static void reduce_candidates(int k, const long long* arg, const double* val, const long long* offset, int sign,
                              int64_t* arg_out, double* val_out) {
  long long bi = -1;
  double bv = 0.0;
  for (int d = 0; d < k; ++d) {
    if (arg[d] < 0 || val[d] != val[d]) continue;
    const long long gi = arg[d] + offset[d];
    if (bi < 0 || (sign > 0 ? val[d] > bv : val[d] < bv) || (val[d] == bv && gi < bi)) {
      bi = gi;
      bv = val[d];
    }
  }
  if (arg_out) *arg_out = bi;
  if (val_out) *val_out = bi >= 0 ? bv : nan("");
}

int multi_gls(pdc_ctx* ctx, const double* t, const double* y, const double* w, int64_t n, double fmin, double df,
              int64_t j0, int64_t nf, unsigned flags, double psd_scale, double* power_out, int64_t* argmax_out,
              double* max_out) {
  MultiState* m = ctx->multi;
  const int k = devices_for(m, (double)n * (double)nf, nf);
  if (k <= 1) return pdc_gls(m->devs[0], t, y, w, n, fmin, df, j0, nf, flags, psd_scale, power_out, argmax_out, max_out);
  std::vector<long long> arg(k, -1), off(k, 0);
  std::vector<double> val(k, 0.0);
  int rc = fan_out(m, k, [&](int d) {
    long long a, b;
    slice_of(nf, k, d, a, b);
    off[d] = a;
    if (b <= a) return (int)PDC_OK;
    int64_t ai = -1;
    double av = 0.0;
    int r = pdc_gls(m->devs[d], t, y, w, n, fmin, df, j0 + a, b - a, flags, psd_scale,
                    power_out ? power_out + a : nullptr, &ai, &av);
    arg[d] = ai;
    val[d] = av;
    return r;
  });
  if (rc != PDC_OK) return rc;
  reduce_candidates(k, arg.data(), val.data(), off.data(), +1, argmax_out, max_out);
  return PDC_OK;
}

int multi_gls_batch(pdc_ctx* ctx, const double* t, const double* y, const double* w, const int64_t* offsets, int64_t B,
                    const double* fmin, const double* df, int64_t nf, unsigned flags, const double* psd_scale,
                    double* power_out, int64_t* argmax_out, double* max_out) {
  MultiState* m = ctx->multi;
  const double ntot = (double)(offsets[B] - offsets[0]);
  const int k = devices_for(m, ntot * (double)nf, B);
  if (k <= 1)
    return pdc_gls_batch(m->devs[0], t, y, w, offsets, B, fmin, df, nf, flags, psd_scale, power_out, argmax_out, max_out);
  // contiguous groups of whole curves, boundaries balanced by sample count
  std::vector<int64_t> first(k + 1, 0);
  first[k] = B;
  for (int d = 1; d < k; ++d) {
    const int64_t target = offsets[0] + (int64_t)(ntot / k * d);
    int64_t b = first[d - 1] + 1;
    while (b < B - (k - d) && offsets[b] < target) ++b;
    first[d] = b;
  }
  return fan_out(m, k, [&](int d) {
    const int64_t b0 = first[d], nb = first[d + 1] - first[d];
    if (nb <= 0) return (int)PDC_OK;
    return pdc_gls_batch(m->devs[d], t, y, w, offsets + b0, nb, fmin + b0, df + b0, nf, flags,
                         psd_scale ? psd_scale + b0 : nullptr, power_out ? power_out + (size_t)b0 * nf : nullptr,
                         argmax_out ? argmax_out + b0 : nullptr, max_out ? max_out + b0 : nullptr);
  });
}

int multi_gls_multi(pdc_ctx* ctx, const double* t, const double* Y, const double* w, int64_t n, int64_t S, double fmin,
                    double df, int64_t j0, int64_t nf, unsigned flags, double psd_scale, double* power_out,
                    int64_t* argmax_out, double* max_out) {
  MultiState* m = ctx->multi;
  const int k = devices_for(m, (double)n * (double)nf * (double)S, (S + 7) / 8);  // whole groups of 8 series per device
  if (k <= 1)
    return pdc_gls_multi(m->devs[0], t, Y, w, n, S, fmin, df, j0, nf, flags, psd_scale, power_out, argmax_out, max_out);
  return fan_out(m, k, [&](int d) {
    long long ga, gb;
    slice_of((S + 7) / 8, k, d, ga, gb);
    const int64_t s0 = ga * 8, s1 = gb * 8 < S ? gb * 8 : S;
    if (s1 <= s0) return (int)PDC_OK;
    return pdc_gls_multi(m->devs[d], t, Y + (size_t)s0 * n, w, n, s1 - s0, fmin, df, j0, nf, flags, psd_scale,
                         power_out ? power_out + (size_t)s0 * nf : nullptr, argmax_out ? argmax_out + s0 : nullptr,
                         max_out ? max_out + s0 : nullptr);
  });
}

// period-grid methods: `call(dev_ctx, periods, np, out, arg, best)` runs the single-device entry point on a slice
int multi_period_grid(pdc_ctx* ctx, int64_t n, const double* periods, int64_t np, int sign, double* out,
                      int64_t* arg_out, double* best_out,
                      const std::function<int(pdc_ctx*, const double*, int64_t, double*, int64_t*, double*)>& call) {
  MultiState* m = ctx->multi;
  const int k = devices_for(m, (double)n * (double)np, np);
  if (k <= 1) return call(m->devs[0], periods, np, out, arg_out, best_out);
  std::vector<long long> arg(k, -1), off(k, 0);
  std::vector<double> val(k, 0.0);
  int rc = fan_out(m, k, [&](int d) {
    long long a, b;
    slice_of(np, k, d, a, b);
    off[d] = a;
    if (b <= a) return (int)PDC_OK;
    int64_t ai = -1;
    double av = 0.0;
    int r = call(m->devs[d], periods + a, b - a, out + a, &ai, &av);
    arg[d] = ai;
    val[d] = av;
    return r;
  });
  if (rc != PDC_OK) return rc;
  reduce_candidates(k, arg.data(), val.data(), off.data(), sign, arg_out, best_out);
  return PDC_OK;
}

}  // namespace pdc

using namespace pdc;

extern "C" {

int pdc_ctx_create_multi(pdc_ctx** out, const int* device_ids, int ndev) {
  if (!out) { set_error("pdc_ctx_create_multi: out is NULL"); return PDC_EINVAL; }
  *out = nullptr;
  if (!device_ids || ndev < 1 || ndev > PDC_MAX_PEERS) {
    set_error("pdc_ctx_create_multi: need 1 <= ndev <= %d device ordinals", PDC_MAX_PEERS);
    return PDC_EINVAL;
  }
  // An ordinal may repeat: every entry gets its own child ctx (stream, scratch, worker thread).  That buys no speed,
  // but it runs the whole sharded path -- slicing, concurrent workers, host reduction -- on a single-GPU box
  // (tests/test_multi_device.py does exactly that).
  pdc_ctx* primary = nullptr;
  PDC_TRY(pdc_ctx_create(&primary, device_ids[0]));
  if (ndev == 1) { *out = primary; return PDC_OK; }
  MultiState* m = new (std::nothrow) MultiState();
  if (!m) { pdc_ctx_destroy(primary); set_error("pdc_ctx_create_multi: out of host memory"); return PDC_ENOMEM; }
  if (const char* g = getenv("PDC_MULTI_MIN_EVALS")) m->min_evals_per_device = atoll(g) > 0 ? atoll(g) : 1;
  primary->multi = m;
  for (int d = 0; d < ndev; ++d) {
    pdc_ctx* child = nullptr;
    int rc = pdc_ctx_create(&child, device_ids[d]);
    if (rc != PDC_OK) { pdc_ctx_destroy(primary); return rc; }   // destroys the children created so far too
    m->devs.push_back(child);
    if (d == 0) continue;   // the first device's piece runs on the calling thread
    Worker* w = new (std::nothrow) Worker();
    if (!w) { pdc_ctx_destroy(primary); set_error("pdc_ctx_create_multi: out of host memory"); return PDC_ENOMEM; }
    w->ctx = child;
    m->workers.push_back(w);
    w->th = std::thread([w] { w->loop(); });
  }
  *out = primary;
  return PDC_OK;
}

int pdc_ctx_device_count(pdc_ctx* ctx) { return ctx ? multi_device_count(ctx) : -1; }

int pdc_ctx_device_id(pdc_ctx* ctx, int index) {
  if (!ctx || index < 0 || index >= multi_device_count(ctx)) return -1;
  return ctx->multi ? ctx->multi->devs[index]->device : ctx->device;
}

}  // extern "C"
