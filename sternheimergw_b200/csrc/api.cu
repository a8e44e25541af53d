// api.cu -- C-ABI entry points (include/sgw_b200.h): context, operator installation, linear_op and
// the batched select_solver replacement.  Host arrays in, host arrays out; all device memory is owned here.
#include "internal.cuh"

#include <mutex>
#include <thread>

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>

using namespace sgw;

namespace sgw {

namespace {
struct DevPool {
  std::multimap<std::pair<int, size_t>, void *> parked;   // (device, bytes) -> block
  std::map<void *, std::pair<int, size_t>> live;
  int contexts = 0;
  std::mutex mu;                                           // contexts of several host threads share the pool
};
DevPool &pool() { static DevPool p; return p; }
void pool_trim_locked(DevPool &dp) {
  for (auto &kv : dp.parked) cudaFree(kv.second);
  dp.parked.clear();
}
}  // namespace

// A parked block is handed out again without a stream dependency: callers free table buffers only after synchronising their
// stream (upload() and the sgw_set_* entry points do), so no queued work can still touch a parked block.
cudaError_t dev_malloc(void **p, size_t bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  DevPool &dp = pool();
  std::lock_guard<std::mutex> lock(dp.mu);
  auto it = dp.parked.find({dev, bytes});
  if (it != dp.parked.end()) {
    *p = it->second;
    dp.parked.erase(it);
  } else {
    const cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {          // give parked memory back and retry once
      cudaGetLastError();
      pool_trim_locked(dp);
      const cudaError_t e2 = cudaMalloc(p, bytes);
      if (e2 != cudaSuccess) return e2;
    }
  }
  dp.live[*p] = {dev, bytes};
  return cudaSuccess;
}

void dev_free(void *p) {
  if (!p) return;
  DevPool &dp = pool();
  std::lock_guard<std::mutex> lock(dp.mu);
  auto it = dp.live.find(p);
  if (it == dp.live.end()) { cudaFree(p); return; }
  dp.parked.insert({it->second, p});
  dp.live.erase(it);
}

void dev_pool_trim() {
  DevPool &dp = pool();
  std::lock_guard<std::mutex> lock(dp.mu);
  pool_trim_locked(dp);
}

namespace {
struct CopyLane {
  void *pin = nullptr;
  cudaStream_t st = nullptr;
  cudaEvent_t ev = nullptr;
};
struct CopyPool {
  std::mutex mu;                       // one staged transfer at a time per process
  int device = -1;
  std::vector<CopyLane> lanes;
  size_t chunk = 0;
};
CopyPool &copy_pool() { static CopyPool p; return p; }
int copy_threads() {
  const char *e = getenv("SGW_COPY_THREADS");
  const int n = e ? atoi(e) : 4;
  return std::max(0, std::min(n, 16));
}
constexpr size_t COPY_CHUNK = (size_t)8 << 20, COPY_MIN = (size_t)32 << 20;

// the pool's lanes for this device (created on first use; nullptr when page-locked memory cannot be had)
CopyPool *copy_pool_for(int device, int nl) {
  CopyPool &cp = copy_pool();
  if (cp.device == device && (int)cp.lanes.size() == nl) return &cp;
  for (auto &l : cp.lanes) { if (l.pin) cudaFreeHost(l.pin); if (l.st) cudaStreamDestroy(l.st); if (l.ev) cudaEventDestroy(l.ev); }
  cp.lanes.assign(nl, CopyLane());
  cp.device = device; cp.chunk = COPY_CHUNK;
  for (auto &l : cp.lanes) {
    if (cudaMallocHost(&l.pin, cp.chunk) != cudaSuccess || cudaStreamCreateWithFlags(&l.st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&l.ev, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      for (auto &k : cp.lanes) { if (k.pin) cudaFreeHost(k.pin); if (k.st) cudaStreamDestroy(k.st); if (k.ev) cudaEventDestroy(k.ev); }
      cp.lanes.clear(); cp.device = -1;
      return nullptr;
    }
  }
  return &cp;
}
}  // namespace

static int staged_copy(sgw_ctx *ctx, void *dst, const void *src, size_t bytes, bool to_device) {
  const int nl = copy_threads();
  const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
  CopyPool *cp = nullptr;
  std::unique_lock<std::mutex> lock(copy_pool().mu, std::defer_lock);
  if (nl > 0 && bytes >= COPY_MIN) { lock.lock(); cp = copy_pool_for(ctx->device, nl); }
  if (!cp) {
    SGW_CUDA(cudaMemcpyAsync(dst, src, bytes, kind, ctx->stream));
    if (!to_device) SGW_CUDA(cudaStreamSynchronize(ctx->stream));
    return SGW_OK;
  }
  SGW_CUDA(cudaStreamSynchronize(ctx->stream));      // earlier work on the context's stream (producers of src / consumers of dst)
  const size_t nchunk = (bytes + cp->chunk - 1) / cp->chunk;
  std::vector<int> rc(nl, 0);
  auto work = [&](int t) {
    cudaSetDevice(ctx->device);
    CopyLane &l = cp->lanes[t];
    for (size_t c = t; c < nchunk; c += nl) {
      const size_t off = c * cp->chunk, len = std::min(cp->chunk, bytes - off);
      if (to_device) {
        memcpy(l.pin, (const char *)src + off, len);
        if (cudaMemcpyAsync((char *)dst + off, l.pin, len, kind, l.st) != cudaSuccess) { rc[t] = 1; return; }
        if (cudaStreamSynchronize(l.st) != cudaSuccess) { rc[t] = 1; return; }   // the buffer is re-used by the next chunk
      } else {
        if (cudaMemcpyAsync(l.pin, (const char *)src + off, len, kind, l.st) != cudaSuccess) { rc[t] = 1; return; }
        if (cudaStreamSynchronize(l.st) != cudaSuccess) { rc[t] = 1; return; }
        memcpy((char *)dst + off, l.pin, len);
      }
    }
  };
  {
    std::vector<std::thread> th;
    for (int t = 1; t < nl; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &t : th) t.join();
  }
  for (int t = 0; t < nl; ++t)
    if (rc[t]) { cudaGetLastError(); ctx->err = "staged host/device copy failed"; return SGW_E_CUDA; }
  return SGW_OK;
}

int h2d_large(sgw_ctx *ctx, void *d_dst, const void *h_src, size_t bytes) { return staged_copy(ctx, d_dst, h_src, bytes, true); }
int d2h_large(sgw_ctx *ctx, void *h_dst, const void *d_src, size_t bytes) { return staged_copy(ctx, h_dst, d_src, bytes, false); }

int ws_get(sgw_ctx *ctx, const char *name, size_t bytes, void **out) {
  if (bytes == 0) bytes = 16;
  auto it = ctx->ws.bufs.find(name);
  if (it != ctx->ws.bufs.end() && it->second.second >= bytes) {
    *out = it->second.first;
    return SGW_OK;
  }
  if (it != ctx->ws.bufs.end()) {
    SGW_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(it->second.first);
    ctx->ws.bufs.erase(it);
  }
  void *p = nullptr;
  size_t want = bytes + bytes / 8;   // a little slack so that slowly growing batches do not reallocate every call
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = bytes;
    e = cudaMalloc(&p, want);
  }
  if (e != cudaSuccess) {          // give the parked table buffers back to the driver and try once more
    cudaGetLastError();
    dev_pool_trim();
    e = cudaMalloc(&p, want);
  }
  if (e != cudaSuccess) {
    char buf[256];
    snprintf(buf, sizeof buf, "cudaMalloc of %zu bytes for workspace '%s' failed: %s", bytes, name, cudaGetErrorString(e));
    ctx->err = buf;
    return SGW_E_CUDA;
  }
  ctx->ws.bufs[name] = std::make_pair(p, want);
  *out = p;
  return SGW_OK;
}

void ws_free_all(sgw_ctx *ctx) {
  for (auto &kv : ctx->ws.bufs) cudaFree(kv.second.first);
  ctx->ws.bufs.clear();
}

static void free_slot(KSlot &k) {
  free_sphere(&k.sph);
  if (k.d_g2kin) dev_free(k.d_g2kin);
  if (k.d_P) dev_free(k.d_P);
  if (k.d_dion) dev_free(k.d_dion);
  if (k.d_dion_ptr) dev_free(k.d_dion_ptr);
  if (k.d_dion_col) dev_free(k.d_dion_col);
  if (k.d_dion_val) dev_free(k.d_dion_val);
  if (k.d_A) dev_free(k.d_A);
  k = KSlot();
}

static void free_pair(KPair &p) {
  free_sphere(&p.sph_k);
  if (p.d_evc) dev_free(p.d_evc);
  if (p.d_evq_all) dev_free(p.d_evq_all);
  p = KPair();
}

static cudaEvent_t pool_event(sgw_ctx *ctx) {
  if (!ctx->ev_pool.empty()) {
    cudaEvent_t e = ctx->ev_pool.back();
    ctx->ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

void prof_begin(sgw_ctx *ctx, int cls) {
  ProfRec r;
  r.cls = cls; r.a = pool_event(ctx); r.b = pool_event(ctx);
  cudaEventRecord(r.a, ctx->stream);
  ctx->prof_recs.push_back(r);
}

void prof_end(sgw_ctx *ctx) { cudaEventRecord(ctx->prof_recs.back().b, ctx->stream); }

static void prof_collect(sgw_ctx *ctx) {
  for (auto &r : ctx->prof_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { ctx->prof_ms[r.cls] += ms; ctx->prof_n[r.cls] += 1; }
    ctx->ev_pool.push_back(r.a);
    ctx->ev_pool.push_back(r.b);
  }
  ctx->prof_recs.clear();
  ctx->stats.ms_linear_op = ctx->prof_ms[PC_FFT_Z] + ctx->prof_ms[PC_FFT_PLANE] + ctx->prof_ms[PC_GEMM_PROJ] + ctx->prof_ms[PC_GEMM_OUT];
}

void begin_call(sgw_ctx *ctx) {
  // records left behind by a call that returned early with an error (its end_call never ran) must not be billed to this one
  for (auto &r : ctx->prof_recs) { ctx->ev_pool.push_back(r.a); ctx->ev_pool.push_back(r.b); }
  ctx->prof_recs.clear();
  memset(&ctx->stats, 0, sizeof(ctx->stats));
  ctx->launches = 0;
  memset(ctx->prof_ms, 0, sizeof(ctx->prof_ms));
  memset(ctx->prof_n, 0, sizeof(ctx->prof_n));
  cudaEventRecord(ctx->ev0, ctx->stream);
}

void lanes_destroy(sgw_ctx *ctx) {
  for (sgw_ctx *l : ctx->lanes) {
    cudaStreamSynchronize(l->own_stream);
    ws_free_all(l);
    cudaEventDestroy(l->ev0); cudaEventDestroy(l->ev1); cudaEventDestroy(l->ev2); cudaEventDestroy(l->ev3);
    for (auto &r : l->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : l->ev_pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) if (l->ev_iter[i]) cudaEventDestroy(l->ev_iter[i]);
    if (l->h_flags) cudaFreeHost(l->h_flags);
    cudaStreamDestroy(l->own_stream);
    delete l;                                   // the tables it points to belong to the parent
  }
  ctx->lanes.clear();
}

void end_call(sgw_ctx *ctx) {
  cudaEventRecord(ctx->ev1, ctx->stream);
  cudaEventSynchronize(ctx->ev1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->stats.ms_total = ms;
  ctx->stats.n_kernel_launch = ctx->launches;
  if (ctx->profiling) prof_collect(ctx);
}

}  // namespace sgw

extern "C" {

int sgw_create(int device, sgw_ctx **out) {
  if (!out) return SGW_E_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return SGW_E_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return SGW_E_CUDA;
  sgw_ctx *ctx = new sgw_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return SGW_E_CUDA; }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return SGW_E_CUDA; }
  ctx->own_stream = ctx->stream;
  cudaEventCreate(&ctx->ev0); cudaEventCreate(&ctx->ev1); cudaEventCreate(&ctx->ev2); cudaEventCreate(&ctx->ev3);
  memset(&ctx->stats, 0, sizeof(ctx->stats));
  { std::lock_guard<std::mutex> lock(pool().mu); pool().contexts++; }
  *out = ctx;
  return SGW_OK;
}

int sgw_destroy(sgw_ctx *ctx) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  lanes_destroy(ctx);
  for (auto &k : ctx->slots) free_slot(k);
  for (auto &p : ctx->pairs) free_pair(p);
  for (auto &kv : ctx->rho_spheres) free_sphere(&kv.second);
  free_sphere(&ctx->rho_sph_c);
  for (auto &sp : ctx->pair_k_c) free_sphere(&sp);
  for (auto &sp : ctx->pair_kq_c) free_sphere(&sp);
  free_fft_grid(&ctx->rho_grid);
  ws_free_all(ctx);
  if (ctx->d_twx) dev_free(ctx->d_twx);
  if (ctx->d_twy) dev_free(ctx->d_twy);
  if (ctx->d_twz) dev_free(ctx->d_twz);
  if (ctx->d_vperm) dev_free(ctx->d_vperm);
  if (ctx->d_vperm_t) dev_free(ctx->d_vperm_t);
  cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1); cudaEventDestroy(ctx->ev2); cudaEventDestroy(ctx->ev3);
  for (auto &r : ctx->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : ctx->ev_pool) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) if (ctx->ev_iter[i]) cudaEventDestroy(ctx->ev_iter[i]);
  if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
  if (ctx->corr.d_Ec) dev_free(ctx->corr.d_Ec);
  if (ctx->corr.d_ET) dev_free(ctx->corr.d_ET);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  bool last = false;
  {
    std::lock_guard<std::mutex> lock(pool().mu);
    last = --pool().contexts <= 0;
    if (last) pool().contexts = 0;
  }
  if (last) dev_pool_trim();
  return SGW_OK;
}

const char *sgw_last_error(const sgw_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int sgw_get_stats(const sgw_ctx *ctx, sgw_stats *out) {
  if (!ctx || !out) return SGW_E_ARG;
  *out = ctx->stats;
  return SGW_OK;
}

int sgw_set_profiling(sgw_ctx *ctx, int on) {
  if (!ctx) return SGW_E_ARG;
  ctx->profiling = on != 0;
  return SGW_OK;
}

int sgw_release_workspace(sgw_ctx *ctx) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  ws_free_all(ctx);
  lanes_destroy(ctx);
  dev_pool_trim();
  return SGW_OK;
}

int sgw_get_profile(const sgw_ctx *ctx, int max_classes, double *ms, int64_t *regions, int *nclasses) {
  if (!ctx || !ms || !regions || !nclasses) return SGW_E_ARG;
  *nclasses = PC_N;
  for (int i = 0; i < PC_N && i < max_classes; ++i) { ms[i] = ctx->prof_ms[i]; regions[i] = ctx->prof_n[i]; }
  return SGW_OK;
}

const char *sgw_profile_class_name(int cls) {
  static const char *names[PC_N] = {"fft_zpass", "fft_plane", "gemm_project", "gemm_expand", "shift_fused", "seed_blas1",
                                    "rho_plane", "other", "gw_product", "shift_gemm"};
  return cls >= 0 && cls < PC_N ? names[cls] : "";
}

int sgw_set_stream(sgw_ctx *ctx, void *cuda_stream) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return SGW_OK;
}

int sgw_set_message_callback(sgw_ctx *ctx, sgw_message_fn fn, void *user) {
  if (!ctx) return SGW_E_ARG;
  ctx->msg_fn = fn;
  ctx->msg_user = user;
  return SGW_OK;
}

int sgw_device_synchronize(sgw_ctx *ctx) {
  if (!ctx) return SGW_E_ARG;
  SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  return SGW_OK;
}

static int upload_twiddle(sgw_ctx *ctx, int n, cplx **d) {
  std::vector<cplx> tw(n);
  for (int m = 0; m < n; ++m) {
    // exact quadrant symmetry via sincospi-like evaluation in long double
    const long double a = -2.0L * 3.141592653589793238462643383279502884L * (long double)m / (long double)n;
    tw[m].x = (double)cosl(a);
    tw[m].y = (double)sinl(a);
  }
  return upload(ctx, d, tw.data(), tw.size());
}

}  // extern "C"

namespace sgw {
int make_fft_grid(sgw_ctx *ctx, int n1, int n2, int n3, FftGrid *gr) {
  free_fft_grid(gr);
  if (!make_plan(n1, &gr->px) || !make_plan(n2, &gr->py) || !make_plan(n3, &gr->pz)) {
    ctx->err = "FFT dimension without a plan: it must be r1*r2 with radices from {1..12, 14, 15, 16, 18, 20, 21, 22, 24, 25, 27, 28, 30, 32} (every 2^a 3^b 5^c <= 960 is, and those times one 7 and/or 11 up to 350)";
    return SGW_E_UNSUPPORTED;
  }
  gr->n1 = n1; gr->n2 = n2; gr->n3 = n3;
  SGW_CHECK(upload_twiddle(ctx, n1, &gr->d_twx));
  SGW_CHECK(upload_twiddle(ctx, n2, &gr->d_twy));
  SGW_CHECK(upload_twiddle(ctx, n3, &gr->d_twz));
  return SGW_OK;
}
void free_fft_grid(FftGrid *gr) {
  if (gr->d_twx) dev_free(gr->d_twx);
  if (gr->d_twy) dev_free(gr->d_twy);
  if (gr->d_twz) dev_free(gr->d_twz);
  *gr = FftGrid();
}
}  // namespace sgw

extern "C" {

int sgw_set_grid(sgw_ctx *ctx, int nr1, int nr2, int nr3, int nr1x, int nr2x, int nr3x) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  ctx->tables_version++;
  SGW_ARG(nr1 > 0 && nr2 > 0 && nr3 > 0, "grid dimensions must be positive");
  SGW_ARG(nr1x >= nr1 && nr2x >= nr2 && nr3x >= nr3, "physical dimensions nr1x, nr2x, nr3x must not be smaller than the FFT dimensions");
  if (!make_plan(nr1, &ctx->px) || !make_plan(nr2, &ctx->py) || !make_plan(nr3, &ctx->pz)) {
    ctx->err = "FFT dimension without a plan: it must be r1*r2 with radices from {1..12, 14, 15, 16, 18, 20, 21, 22, 24, 25, 27, 28, 30, 32} (every 2^a 3^b 5^c <= 960 is, and those times one 7 and/or 11 up to 350)";
    return SGW_E_UNSUPPORTED;
  }
  ctx->nr1 = nr1; ctx->nr2 = nr2; ctx->nr3 = nr3;
  ctx->nr1x = nr1x; ctx->nr2x = nr2x; ctx->nr3x = nr3x;
  SGW_CHECK(upload_twiddle(ctx, nr1, &ctx->d_twx));
  SGW_CHECK(upload_twiddle(ctx, nr2, &ctx->d_twy));
  SGW_CHECK(upload_twiddle(ctx, nr3, &ctx->d_twz));
  auto mk = [](const Plan1D &p, std::vector<int> &perm) {
    perm.resize(p.n);
    for (int i = 0; i < p.n; ++i) perm[i] = perm_index(p.r1, p.r2, i);
  };
  mk(ctx->px, ctx->permx); mk(ctx->py, ctx->permy); mk(ctx->pz, ctx->permz);
  ctx->grid_set = true;
  ctx->vloc_set = false;
  for (auto &k : ctx->slots) free_slot(k);
  for (auto &p : ctx->pairs) free_pair(p);
  for (auto &kv : ctx->rho_spheres) free_sphere(&kv.second);
  ctx->rho_spheres.clear();
  return SGW_OK;
}

int sgw_set_vloc(sgw_ctx *ctx, const double *vrs) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_ARG(vrs != nullptr, "vrs is null");
  if (!ctx->grid_set) { ctx->err = "sgw_set_grid must be called first"; return SGW_E_STATE; }
  const int nx = ctx->nr1, ny = ctx->nr2, nz = ctx->nr3;
  std::vector<double> vp((size_t)nx * ny * nz);
  for (int pz = 0; pz < nz; ++pz)
    for (int py = 0; py < ny; ++py)
      for (int pxi = 0; pxi < nx; ++pxi)
        vp[((size_t)pz * ny + py) * nx + pxi] = vrs[ctx->permx[pxi] + (size_t)ctx->nr1x * (ctx->permy[py] + (size_t)ctx->nr2x * ctx->permz[pz])];
  SGW_CHECK(upload(ctx, &ctx->d_vperm, vp.data(), vp.size()));
  {
    // the same potential with y fastest, [pz][px][py]: in the fused middle x stage consecutive lanes work on consecutive
    // rows (y) of the same x group, so this layout makes their 8-byte loads of v coalesce (k_plane_vloc, fft.cu)
    std::vector<double> vt(vp.size());
    for (int pz = 0; pz < nz; ++pz)
      for (int py = 0; py < ny; ++py)
        for (int pxi = 0; pxi < nx; ++pxi) vt[((size_t)pz * nx + pxi) * ny + py] = vp[((size_t)pz * ny + py) * nx + pxi];
    SGW_CHECK(upload(ctx, &ctx->d_vperm_t, vt.data(), vt.size()));
  }
  ctx->vloc_set = true;
  return SGW_OK;
}

static KSlot *get_slot(sgw_ctx *ctx, int slot) {
  if (slot < 0 || slot > (1 << 20)) return nullptr;
  if ((int)ctx->slots.size() <= slot) ctx->slots.resize(slot + 1);
  return &ctx->slots[slot];
}

int sgw_set_kpoint(sgw_ctx *ctx, int slot, int npw, int npwx, const int32_t *nl_igk, const double *g2kin, int nkb,
                   const sgw_cplx *vkb, const double *dion, int nbnd_occ, const sgw_cplx *evq, double alpha_pv) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  if (!ctx->grid_set) { ctx->err = "sgw_set_grid must be called first"; return SGW_E_STATE; }
  SGW_ARG(npw > 0 && npwx >= npw, "need 0 < npw <= npwx");
  SGW_ARG(nl_igk && g2kin, "nl_igk / g2kin null");
  SGW_ARG(nkb >= 0 && nbnd_occ >= 0, "negative nkb / nbnd_occ");
  SGW_ARG(nkb == 0 || (vkb && dion), "vkb / dion null");
  SGW_ARG(nbnd_occ == 0 || evq, "evq null");
  KSlot *k = get_slot(ctx, slot);
  SGW_ARG(k != nullptr, "bad slot");
  free_slot(*k);
  ctx->tables_version++;
  {
    std::vector<int32_t> nlc(nl_igk, nl_igk + npw);
    if (grid_padded(ctx)) for (auto &v : nlc) v = unpad_index(ctx, v);
    SGW_CHECK(build_sphere(ctx, npw, nlc.data(), &k->sph));
  }
  k->npw = npw; k->npwx = npwx; k->nkb = nkb; k->nbnd = nbnd_occ; k->alpha_pv = alpha_pv;
  const std::vector<int> &perm = k->sph.perm;
  std::vector<double> g2(npwx, 0.0);
  for (int p = 0; p < npw; ++p) g2[p] = g2kin[perm[p]];
  SGW_CHECK(upload(ctx, &k->d_g2kin, g2.data(), g2.size()));
  const int m = nkb + nbnd_occ;
  if (m > 0) {
    // P = [vkb | evq] with rows in column order: the caller's arrays go up as they are and are permuted on the device
    cplx *stage = nullptr;
    const int mmax = std::max(nkb, nbnd_occ);
    SGW_CHECK(ws(ctx, "io_in", (size_t)npwx * mmax, &stage));
    if (k->d_P) { dev_free(k->d_P); k->d_P = nullptr; }
    SGW_CUDA(dev_malloc((void **)&k->d_P, sizeof(cplx) * (size_t)npwx * m));
    if (nkb > 0) {
      SGW_CHECK(h2d_large(ctx, stage, vkb, sizeof(cplx) * (size_t)npwx * nkb));
      SGW_CHECK(permute_in(ctx, k->sph, nkb, stage, npwx, k->d_P, npwx, npwx));
    }
    if (nbnd_occ > 0) {
      SGW_CHECK(h2d_large(ctx, stage, evq, sizeof(cplx) * (size_t)npwx * nbnd_occ));
      SGW_CHECK(permute_in(ctx, k->sph, nbnd_occ, stage, npwx, k->d_P + (size_t)nkb * npwx, npwx, npwx));
    }
    SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  if (nkb > 0) {
    SGW_CHECK(upload(ctx, &k->d_dion, dion, (size_t)nkb * nkb));
    // compressed rows of D (column-major input): used when at most a quarter of the entries is non-zero
    std::vector<int> ptr(nkb + 1, 0), col;
    std::vector<double> val;
    for (int i = 0; i < nkb; ++i) {
      for (int j = 0; j < nkb; ++j) {
        const double d = dion[i + (size_t)nkb * j];
        if (d != 0.0) { col.push_back(j); val.push_back(d); }
      }
      ptr[i + 1] = (int)col.size();
    }
    if (k->d_dion_ptr) { dev_free(k->d_dion_ptr); k->d_dion_ptr = nullptr; }
    if (k->d_dion_col) { dev_free(k->d_dion_col); k->d_dion_col = nullptr; }
    if (k->d_dion_val) { dev_free(k->d_dion_val); k->d_dion_val = nullptr; }
    if (col.size() * 4 <= (size_t)nkb * nkb && !col.empty()) {
      SGW_CHECK(upload(ctx, &k->d_dion_ptr, ptr.data(), ptr.size()));
      SGW_CHECK(upload(ctx, &k->d_dion_col, col.data(), col.size()));
      SGW_CHECK(upload(ctx, &k->d_dion_val, val.data(), val.size()));
    }
  }
  k->set = true;
  k->dense = false;
  return SGW_OK;
}

int sgw_set_dense_operator(sgw_ctx *ctx, int slot, int n, const sgw_cplx *A, int lda) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_ARG(n > 0 && A && lda >= n, "bad dense operator");
  KSlot *k = get_slot(ctx, slot);
  SGW_ARG(k != nullptr, "bad slot");
  free_slot(*k);
  std::vector<cplx> a((size_t)n * n);
  for (int j = 0; j < n; ++j) memcpy(&a[(size_t)j * n], (const cplx *)A + (size_t)j * lda, sizeof(cplx) * n);
  SGW_CHECK(upload(ctx, &k->d_A, a.data(), a.size()));
  k->npw = k->npwx = n;
  k->sph.npw = n;
  k->sph.perm.resize(n);
  std::vector<int> id(n);
  for (int i = 0; i < n; ++i) id[i] = k->sph.perm[i] = i;
  SGW_CHECK(upload(ctx, &k->sph.d_perm, id.data(), id.size()));
  k->dense = true;
  k->set = true;
  return SGW_OK;
}

int sgw_linear_op(sgw_ctx *ctx, int slot, int nvec, const sgw_cplx *omega, double alpha_pv, const sgw_cplx *psi, int ldpsi,
                  sgw_cplx *apsi, int ldapsi) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_ARG(nvec > 0 && omega && psi && apsi, "null argument");
  SGW_ARG(nvec <= 65535, "at most 65535 vectors per call (grid limit of the batched kernels): split the batch");
  if (slot < 0 || slot >= (int)ctx->slots.size() || !ctx->slots[slot].set) { ctx->err = "operator slot not set"; return SGW_E_STATE; }
  const KSlot &k = ctx->slots[slot];
  SGW_ARG(ldpsi >= k.npw && ldapsi >= k.npw, "leading dimension smaller than npw");
  begin_call(ctx);
  const long n = k.npwx;
  cplx *d_in = nullptr, *d_psi = nullptr, *d_out = nullptr, *d_om = nullptr;
  SGW_CHECK(ws(ctx, "io_in", (size_t)std::max(ldpsi, ldapsi) * nvec, &d_in));
  SGW_CHECK(ws(ctx, "lo_psi", (size_t)n * nvec, &d_psi));
  SGW_CHECK(ws(ctx, "lo_out", (size_t)n * nvec, &d_out));
  SGW_CHECK(ws(ctx, "lo_om", (size_t)nvec, &d_om));
  SGW_CUDA(cudaMemcpyAsync(d_in, psi, sizeof(cplx) * (size_t)ldpsi * nvec, cudaMemcpyHostToDevice, ctx->stream));
  SGW_CUDA(cudaMemcpyAsync(d_om, omega, sizeof(cplx) * nvec, cudaMemcpyHostToDevice, ctx->stream));
  SGW_CHECK(permute_in(ctx, k.sph, nvec, d_in, ldpsi, d_psi, n, (int)n));
  SGW_CHECK(apply_operator(ctx, slot, alpha_pv, nvec, d_psi, n, d_om, 1, d_out, n, nullptr));
  // A_psi is INTENT(OUT): rows beyond npw are defined (zero) like the reference's zero padded arrays
  SGW_CUDA(cudaMemsetAsync(d_in, 0, sizeof(cplx) * (size_t)ldapsi * nvec, ctx->stream));
  SGW_CHECK(permute_out(ctx, k.sph, nvec, d_out, n, d_in, ldapsi));
  SGW_CUDA(cudaMemcpyAsync(apsi, d_in, sizeof(cplx) * (size_t)ldapsi * nvec, cudaMemcpyDeviceToHost, ctx->stream));
  SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stats.n_linear_op = nvec;
  end_call(ctx);
  return SGW_OK;
}

int sgw_solve_multishift(sgw_ctx *ctx, int slot, const sgw_solver_cfg *cfg, int use_alpha_pv, int nrhs, const sgw_cplx *b,
                         int ldb, int nshift, const sgw_cplx *sigma, sgw_cplx *x, int64_t stride_shift, int64_t stride_rhs,
                         int32_t *ierr) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_ARG(cfg && b && sigma && x && ierr, "null argument");
  SGW_ARG(nrhs > 0 && nshift > 0, "need nrhs > 0 and nshift > 0");
  SGW_ARG(nrhs <= 65535, "at most 65535 right-hand sides per call (grid limit of the batched kernels): split the batch");
  SGW_ARG(cfg->npriority >= 1 && cfg->npriority <= 4, "priority of the solvers not specified");   // select_solver.f90:116-118
  if (slot < 0 || slot >= (int)ctx->slots.size() || !ctx->slots[slot].set) { ctx->err = "operator slot not set"; return SGW_E_STATE; }
  const KSlot &k = ctx->slots[slot];
  SGW_ARG(ldb >= k.npw, "ldb smaller than npw");
  SGW_ARG(stride_shift >= k.npw || nshift == 1, "stride_shift smaller than npw");
  begin_call(ctx);
  const long n = k.npwx;
  cplx *d_in = nullptr, *d_b = nullptr, *d_sig = nullptr, *d_x = nullptr;
  int *d_ierr = nullptr;
  SGW_CHECK(ws(ctx, "io_in", (size_t)ldb * nrhs, &d_in));
  SGW_CHECK(ws(ctx, "sv_b", (size_t)n * nrhs, &d_b));
  SGW_CHECK(ws(ctx, "sv_sig", (size_t)nshift * nrhs, &d_sig));
  SGW_CHECK(ws(ctx, "sv_x", (size_t)n * nshift * nrhs, &d_x));
  SGW_CHECK(ws(ctx, "sv_ierr", (size_t)nrhs, &d_ierr));
  SGW_CUDA(cudaMemcpyAsync(d_in, b, sizeof(cplx) * (size_t)ldb * nrhs, cudaMemcpyHostToDevice, ctx->stream));
  SGW_CUDA(cudaMemcpyAsync(d_sig, sigma, sizeof(cplx) * (size_t)nshift * nrhs, cudaMemcpyHostToDevice, ctx->stream));
  SGW_CHECK(permute_in(ctx, k.sph, nrhs, d_in, ldb, d_b, n, (int)n));
  SolveBatch sb;
  sb.slot = slot;
  sb.alpha_pv = use_alpha_pv ? k.alpha_pv : 0.0;
  sb.nrhs = nrhs; sb.nshift = nshift; sb.n = (int)n;
  sb.d_b = d_b; sb.ldb = n; sb.d_sigma = d_sig; sb.d_x = d_x; sb.d_ierr = d_ierr;
  cudaEventRecord(ctx->ev2, ctx->stream);
  SGW_CHECK(select_solver_batched(ctx, sb, cfg));
  cudaEventRecord(ctx->ev3, ctx->stream);
  // x is INTENT(OUT): un-permute every (rhs, shift) vector into the caller's strided layout
  cplx *d_xo = nullptr;
  SGW_CHECK(ws(ctx, "sv_xo", (size_t)k.npw * nshift * nrhs, &d_xo));
  SGW_CHECK(permute_out(ctx, k.sph, nrhs * nshift, d_x, n, d_xo, k.npw));
  std::vector<cplx> hx((size_t)k.npw * nshift * nrhs);
  SGW_CUDA(cudaMemcpyAsync(hx.data(), d_xo, sizeof(cplx) * hx.size(), cudaMemcpyDeviceToHost, ctx->stream));
  SGW_CUDA(cudaMemcpyAsync(ierr, d_ierr, sizeof(int) * nrhs, cudaMemcpyDeviceToHost, ctx->stream));
  SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int r = 0; r < nrhs; ++r)
    for (int s = 0; s < nshift; ++s)
      memcpy((cplx *)x + (size_t)r * stride_rhs + (size_t)s * stride_shift, &hx[((size_t)r * nshift + s) * k.npw],
             sizeof(cplx) * k.npw);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3);
  end_call(ctx);
  ctx->stats.ms_solver = ms;
  return SGW_OK;
}

int sgw_parallel_task(int nproc, int rank, int num_task_total, int32_t *first_task, int32_t *last_task, int32_t *num_task) {
  if (nproc <= 0 || rank < 0 || rank >= nproc || num_task_total < 0 || !first_task || !last_task || !num_task) return SGW_E_ARG;
  const int nmin = num_task_total / nproc;          // parallel.f90:121
  const int nrem = num_task_total % nproc;          // :124
  const int last_proc = nproc - nrem;               // :127  (the LAST ranks take the extra task)
  int sum = 0;
  for (int p = 0; p < nproc; ++p) {
    num_task[p] = p < last_proc ? nmin : nmin + 1;
    if (p <= rank) sum += num_task[p];
  }
  *last_task = sum;                                 // :133
  *first_task = sum - num_task[rank] + 1;           // :134
  return SGW_OK;
}

int sgw_bench_linear_op(sgw_ctx *ctx, int slot, int nvec, int reps, double *ms_total, double *ms_fft, double *ms_gemm) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_ARG(nvec > 0 && reps > 0, "nvec, reps must be positive");
  if (slot < 0 || slot >= (int)ctx->slots.size() || !ctx->slots[slot].set || ctx->slots[slot].dense) { ctx->err = "plane-wave slot not set"; return SGW_E_STATE; }
  const KSlot &k = ctx->slots[slot];
  const long n = k.npwx;
  cplx *d_psi = nullptr, *d_out = nullptr, *d_om = nullptr, *T1 = nullptr, *T2 = nullptr;
  SGW_CHECK(ws(ctx, "lo_psi", (size_t)n * nvec, &d_psi));
  SGW_CHECK(ws(ctx, "lo_out", (size_t)n * nvec, &d_out));
  SGW_CHECK(ws(ctx, "lo_om", (size_t)nvec, &d_om));
  {
    std::vector<cplx> h((size_t)n * nvec, cmake(0.0, 0.0));
    uint64_t st = 0x9E3779B97F4A7C15ull;
    for (int v = 0; v < nvec; ++v)
      for (int p = 0; p < k.npw; ++p) {
        st = st * 6364136223846793005ull + 1442695040888963407ull;
        const double a = (double)((st >> 11) & 0xFFFFF) / 1048576.0 - 0.5;
        st = st * 6364136223846793005ull + 1442695040888963407ull;
        const double b2 = (double)((st >> 11) & 0xFFFFF) / 1048576.0 - 0.5;
        h[(size_t)v * n + p] = cmake(a, b2);
      }
    SGW_CUDA(cudaMemcpyAsync(d_psi, h.data(), sizeof(cplx) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<cplx> om(nvec, cmake(0.1, 0.05));
    SGW_CUDA(cudaMemcpyAsync(d_om, om.data(), sizeof(cplx) * nvec, cudaMemcpyHostToDevice, ctx->stream));
    SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  const size_t tsz = (size_t)nvec * ctx->nr3 * k.sph.ncol;
  SGW_CHECK(ws(ctx, "fft_T1", tsz, &T1));
  SGW_CHECK(ws(ctx, "fft_T2", tsz, &T2));
  float ms = 0.f;
  begin_call(ctx);
  // warm-up
  for (int r = 0; r < 3; ++r) SGW_CHECK(apply_operator(ctx, slot, k.alpha_pv, nvec, d_psi, n, d_om, 1, d_out, n, nullptr));
  SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaEventRecord(ctx->ev2, ctx->stream);
  for (int r = 0; r < reps; ++r) SGW_CHECK(apply_operator(ctx, slot, k.alpha_pv, nvec, d_psi, n, d_om, 1, d_out, n, nullptr));
  cudaEventRecord(ctx->ev3, ctx->stream);
  SGW_CUDA(cudaEventSynchronize(ctx->ev3));
  cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3);
  if (ms_total) *ms_total = ms / reps;
  ZEpilogue epi;
  epi.mode = 1; epi.g2kin = k.d_g2kin; epi.psi = d_psi; epi.sigma = d_om; epi.sigma_stride = 1; epi.keep_out = 1;
  cudaEventRecord(ctx->ev2, ctx->stream);
  for (int r = 0; r < reps; ++r) {
    SGW_CHECK(fft_zpass_g2r(ctx, k.sph, nvec, d_psi, n, T1, nullptr));
    SGW_CHECK(fft_plane(ctx, PLANE_VLOC, &k.sph, &k.sph, nvec, T1, T2, nullptr, 1, nullptr, nullptr));
    SGW_CHECK(fft_zpass_r2g(ctx, k.sph, nvec, T2, d_out, n, epi, nullptr));
  }
  cudaEventRecord(ctx->ev3, ctx->stream);
  SGW_CUDA(cudaEventSynchronize(ctx->ev3));
  cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3);
  if (ms_fft) *ms_fft = ms / reps;
  cudaEventRecord(ctx->ev2, ctx->stream);
  for (int r = 0; r < reps; ++r) SGW_CHECK(nonlocal_apply(ctx, k, k.alpha_pv, nvec, d_psi, n, d_out, n, nullptr));
  cudaEventRecord(ctx->ev3, ctx->stream);
  SGW_CUDA(cudaEventSynchronize(ctx->ev3));
  cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3);
  if (ms_gemm) *ms_gemm = ms / reps;
  end_call(ctx);
  return SGW_OK;
}

}  // extern "C"
