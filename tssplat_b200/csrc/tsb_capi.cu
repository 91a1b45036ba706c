// C ABI (include/tssplat_b200.h) over the plan builder and the sm_100a kernels.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/tssplat_b200.h"
#include "tsb_kernels.cuh"
#include "tsb_plan.h"

struct tsb_handle_s {
  int device = 0;
  tsb::KParams kp{};
  const int32_t *slot_ptr = nullptr;
  float *stage_x = nullptr, *stage_grad = nullptr, *stage_energy = nullptr;   // tsb_energy_grad_host staging
  tsb_info_t info{};
  std::vector<void *> allocs;
  std::string err;
};

namespace {

thread_local std::string g_create_err;
std::mutex g_mu;
std::map<int, float *> g_limit_work;  // per-device scratch for tsb_grad_limit

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int fail(tsb_handle_t h, int code, const std::string &msg) {
  if (h) h->err = msg; else g_create_err = msg;
  return code;
}

template <class T>
int upload(tsb_handle_t h, const std::vector<T> &v, const T **out, size_t min_elems = 1) {
  const size_t bytes = std::max(v.size(), min_elems) * sizeof(T);
  void *d = nullptr;
  cudaError_t e = cudaMalloc(&d, bytes);
  if (e != cudaSuccess) return fail(h, TSB_E_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  h->allocs.push_back(d);
  h->info.device_bytes += int64_t(bytes);
  if (!v.empty()) {
    e = cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail(h, TSB_E_CUDA, std::string("cudaMemcpy: ") + cudaGetErrorString(e));
  }
  *out = static_cast<const T *>(d);
  return TSB_OK;
}

template <class T>
int alloc_zero(tsb_handle_t h, size_t elems, T **out) {
  const size_t bytes = std::max<size_t>(elems, 1) * sizeof(T);
  void *d = nullptr;
  cudaError_t e = cudaMalloc(&d, bytes);
  if (e != cudaSuccess) return fail(h, TSB_E_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  h->allocs.push_back(d);
  h->info.device_bytes += int64_t(bytes);
  e = cudaMemset(d, 0, bytes);
  if (e != cudaSuccess) return fail(h, TSB_E_CUDA, std::string("cudaMemset: ") + cudaGetErrorString(e));
  *out = static_cast<T *>(d);
  return TSB_OK;
}

}  // namespace

extern "C" {

int tsb_create(const float *rest_xyz, const int32_t *tets, int32_t n, int32_t nele, const tsb_options_t *opt,
               int device, tsb_handle_t *out) {
  if (!out) return fail(nullptr, TSB_E_INVALID, "out is null");
  *out = nullptr;
  tsb::PlanOptions po;
  if (opt) {
    if (opt->tile_tets != 0) po.tile_tets = opt->tile_tets;
    po.laplacian_scale = opt->laplacian_scale ? 1 : 0;
  }
  if (const char *env = std::getenv("TSSPLAT_B200_TILE_TETS")) {
    if (!(opt && opt->tile_tets != 0)) po.tile_tets = std::atoi(env);
  }
  po.max_local_vertices = tsb::nvmax_for(po.tile_tets);
  if (po.max_local_vertices == 0)
    return fail(nullptr, TSB_E_INVALID, "unsupported tile_tets " + std::to_string(po.tile_tets) + " (compiled: 256, 512, 1024)");

  DeviceGuard guard(device);
  if (!guard.ok) return fail(nullptr, TSB_E_CUDA, "cannot select CUDA device " + std::to_string(device));
  {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) po.balance_sms = sms;
    if (const char *env = std::getenv("TSSPLAT_B200_BALANCE")) po.balance_sms = std::atoi(env) ? po.balance_sms : 0;
  }

  tsb::HostPlan plan;
  std::string err;
  int rc = tsb::build_plan(rest_xyz, tets, n, nele, po, plan, err);
  if (rc != TSB_OK) return fail(nullptr, rc, err);

  cudaError_t ce = tsb::prepare_energy_grad(po.tile_tets);
  if (ce != cudaSuccess) return fail(nullptr, TSB_E_CUDA, std::string("kernel attribute setup: ") + cudaGetErrorString(ce));

  tsb_handle_t h = new tsb_handle_s();
  h->device = device;
  tsb::KParams &kp = h->kp;
#define TSB_TRY(expr) do { rc = (expr); if (rc != TSB_OK) { g_create_err = h->err; tsb_destroy(h); return rc; } } while (0)
  TSB_TRY(upload(h, plan.vblob, &kp.vblob, 16));
  TSB_TRY(upload(h, plan.tblob, &kp.tblob, 16));
  TSB_TRY(upload(h, plan.ell, &kp.ell, 8));
  TSB_TRY(upload(h, plan.slot_ptr, &h->slot_ptr, 2));
  {
    const int32_t *te = nullptr;
    TSB_TRY(upload(h, plan.tile_ell, &te, 2));
    kp.tile_ell = reinterpret_cast<const int2 *>(te);
  }
  TSB_TRY(alloc_zero(h, size_t(plan.n_slots) * 4, &kp.scratch));
  TSB_TRY(alloc_zero(h, size_t(plan.n_tiles) * 2, &kp.tile_energy));
#undef TSB_TRY
  kp.laplacian_scale = plan.laplacian_scale;
  kp.n_tiles = plan.n_tiles;
  kp.fill = plan.fill;

  tsb_info_t &I = h->info;
  I.n = plan.n; I.nele = plan.nele; I.n_tiles = plan.n_tiles; I.tile_tets = plan.tile_tets;
  I.n_components = plan.n_components; I.n_shared_vertices = plan.n_shared_vertices;
  I.n_local_vertices = plan.n_local_vertices; I.n_boundary_faces = plan.n_boundary_faces;
  I.max_local_vertices = plan.max_local_vertices;
  I.fill = plan.fill;
  // bytes one launch requests from the memory system: the fixed-size vertex-blob and tet-blob TMA
  // copies, the gather tables, x gathered per staged vertex (12 B), grad (12 B/vertex), the
  // shared-vertex partials written + read back, per-tile energies
  I.stream_bytes = int64_t(plan.n_tiles) * (tsb::vblob_bytes(plan.tile_tets, plan.max_local_vertices) + int64_t(52) * plan.fill) +
                   int64_t(plan.ell.size()) * 2 + plan.n_local_vertices * 12 + int64_t(plan.n) * (12 + 4) +
                   int64_t(plan.n_slots) * 32 + int64_t(plan.n_tiles) * 16;
  *out = h;
  return TSB_OK;
}

void tsb_destroy(tsb_handle_t h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  for (void *p : h->allocs) cudaFree(p);
  delete h;
}

const char *tsb_last_error(tsb_handle_t h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int tsb_get_info(tsb_handle_t h, tsb_info_t *info) {
  if (!h || !info) return TSB_E_INVALID;
  *info = h->info;
  return TSB_OK;
}

int tsb_energy_grad(tsb_handle_t h, const float *x_dev, float c1, float c2, int32_t order, float gradH,
                    const float *gradH_dev, float *energy_out_dev, float *grad_out_dev, void *stream) {
  if (!h) return TSB_E_INVALID;
  if (!x_dev || !energy_out_dev) return fail(h, TSB_E_INVALID, "x_dev and energy_out_dev must be non-null");
  if (order != 2 && order != 4)
    return fail(h, TSB_E_INVALID, "order must be 2 or 4 (the reference yields zeros for anything else: tet_spheres_cuda.cu:57-63)");
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(h, TSB_E_CUDA, "cannot select the handle's CUDA device");
  tsb::KParams kp = h->kp;
  kp.x = x_dev; kp.grad = grad_out_dev; kp.energy_out = energy_out_dev; kp.gradH_dev = gradH_dev;
  kp.c1 = c1; kp.c2 = c2; kp.gradH = gradH; kp.order = order;
  cudaError_t e = tsb::launch_energy_grad(kp, h->info.tile_tets, h->info.n, h->slot_ptr, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(h, TSB_E_CUDA, std::string("energy_grad launch: ") + cudaGetErrorString(e));
  return TSB_OK;
}

int tsb_energy_grad_host(tsb_handle_t h, const float *x_host, float c1, float c2, int32_t order, float gradH,
                         float *energy_out_host, float *grad_out_host, void *stream) {
  if (!h) return TSB_E_INVALID;
  if (!x_host || !energy_out_host) return fail(h, TSB_E_INVALID, "x_host and energy_out_host must be non-null");
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(h, TSB_E_CUDA, "cannot select the handle's CUDA device");
  const size_t nb = size_t(h->info.n) * 3 * sizeof(float);
  if (!h->stage_x) {
    int rc = alloc_zero(h, size_t(h->info.n) * 3, &h->stage_x);
    if (rc == TSB_OK) rc = alloc_zero(h, size_t(h->info.n) * 3, &h->stage_grad);
    if (rc == TSB_OK) rc = alloc_zero(h, 4, &h->stage_energy);
    if (rc != TSB_OK) return rc;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemcpyAsync(h->stage_x, x_host, nb, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return fail(h, TSB_E_CUDA, std::string("H2D copy: ") + cudaGetErrorString(e));
  const int rc = tsb_energy_grad(h, h->stage_x, c1, c2, order, gradH, nullptr, h->stage_energy,
                                 grad_out_host ? h->stage_grad : nullptr, stream);
  if (rc != TSB_OK) return rc;
  e = cudaMemcpyAsync(energy_out_host, h->stage_energy, 3 * sizeof(float), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && grad_out_host) e = cudaMemcpyAsync(grad_out_host, h->stage_grad, nb, cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return fail(h, TSB_E_CUDA, std::string("D2H copy: ") + cudaGetErrorString(e));
  return TSB_OK;
}

int tsb_scale(const float *g_dev, int64_t count, float gradH, const float *gradH_dev, float *out_dev, void *stream) {
  if (!g_dev || !out_dev || count < 0) return fail(nullptr, TSB_E_INVALID, "tsb_scale: null pointer or negative count");
  if (count == 0) return TSB_OK;
  cudaError_t e = tsb::launch_scale(g_dev, count, gradH, gradH_dev, out_dev, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(nullptr, TSB_E_CUDA, std::string("scale launch: ") + cudaGetErrorString(e));
  return TSB_OK;
}

int tsb_grad_limit(float *grad_dev, int64_t count, float s_threshold, float s, void *stream) {
  if (!grad_dev || count < 0) return fail(nullptr, TSB_E_INVALID, "tsb_grad_limit: null pointer or negative count");
  if (count == 0) return TSB_OK;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(nullptr, TSB_E_CUDA, "cudaGetDevice failed");
  float *work = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_limit_work.find(dev);
    if (it == g_limit_work.end()) {
      if (cudaMalloc(reinterpret_cast<void **>(&work), 4 * sizeof(float)) != cudaSuccess || cudaMemset(work, 0, 4 * sizeof(float)) != cudaSuccess)
        return fail(nullptr, TSB_E_NOMEM, "tsb_grad_limit: scratch allocation failed");
      g_limit_work[dev] = work;
    } else {
      work = it->second;
    }
  }
  cudaError_t e = tsb::launch_grad_limit(grad_dev, count, s_threshold, s, work, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(nullptr, TSB_E_CUDA, std::string("grad_limit launch: ") + cudaGetErrorString(e));
  return TSB_OK;
}

int tsb_adam_uniform_step(float *p_dev, const float *grad_dev, float *g1_dev, float *g2_dev, int64_t count, double lr,
                          double beta1, double beta2, int32_t step, double grad_limit, float *work_dev, void *stream) {
  if (!p_dev || !grad_dev || !g1_dev || !g2_dev || !work_dev || count < 0 || step < 1)
    return fail(nullptr, TSB_E_INVALID, "tsb_adam_uniform_step: null pointer, negative count or step < 1");
  if (count == 0) return TSB_OK;
  cudaError_t e = tsb::launch_adam_uniform(p_dev, grad_dev, g1_dev, g2_dev, count, lr, beta1, beta2, step, grad_limit,
                                           work_dev, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(nullptr, TSB_E_CUDA, std::string("adam_uniform launch: ") + cudaGetErrorString(e));
  return TSB_OK;
}

/* ---- host-plan inspection (tests only; no CUDA calls): lets the CPU test-suite check the tile
 * plan -- staging lists, gather tables, shared-vertex combine lists -- without a GPU. ---------- */
struct tsb_debug_plan_s { tsb::HostPlan plan; };

int tsb_debug_plan_build(const float *rest_xyz, const int32_t *tets, int32_t n, int32_t nele, int32_t tile_tets,
                         int32_t laplacian_scale, int32_t balance_sms, tsb_debug_plan_s **out) {
  if (!out) return TSB_E_INVALID;
  *out = nullptr;
  tsb::PlanOptions po;
  if (tile_tets) po.tile_tets = tile_tets;
  po.laplacian_scale = laplacian_scale;
  po.max_local_vertices = tsb::nvmax_for(po.tile_tets);
  if (po.max_local_vertices == 0) return fail(nullptr, TSB_E_INVALID, "unsupported tile_tets");
  po.balance_sms = balance_sms;
  tsb_debug_plan_s *d = new tsb_debug_plan_s();
  std::string err;
  const int rc = tsb::build_plan(rest_xyz, tets, n, nele, po, d->plan, err);
  if (rc != TSB_OK) { delete d; return fail(nullptr, rc, err); }
  *out = d;
  return TSB_OK;
}

/* name -> (pointer, element count, element bytes); returns TSB_E_INVALID for an unknown name */
int tsb_debug_plan_array(tsb_debug_plan_s *d, const char *name, const void **ptr, int64_t *count, int32_t *elem_bytes) {
  if (!d || !name || !ptr || !count || !elem_bytes) return TSB_E_INVALID;
  const tsb::HostPlan &P = d->plan;
  const std::string k(name);
#define TSB_ARR(nm, vec) if (k == nm) { *ptr = (vec).data(); *count = int64_t((vec).size()); *elem_bytes = int32_t(sizeof((vec)[0])); return TSB_OK; }
  TSB_ARR("vblob", P.vblob) TSB_ARR("tblob", P.tblob) TSB_ARR("ell", P.ell) TSB_ARR("slot_ptr", P.slot_ptr)
  TSB_ARR("tet_order", P.tet_order) TSB_ARR("tile_first", P.tile_first) TSB_ARR("tile_ell", P.tile_ell)
#undef TSB_ARR
  return TSB_E_INVALID;
}

int tsb_debug_plan_scalars(tsb_debug_plan_s *d, int32_t *out8) {  /* out8: 10 ints */
  if (!d || !out8) return TSB_E_INVALID;
  const tsb::HostPlan &P = d->plan;
  out8[0] = P.n; out8[1] = P.nele; out8[2] = P.tile_tets; out8[3] = P.max_local_vertices; out8[4] = P.n_tiles;
  out8[5] = P.n_components; out8[6] = P.n_shared_vertices; out8[7] = P.n_slots; out8[8] = P.fill; out8[9] = tsb::ell_cap(P.tile_tets, P.max_local_vertices);
  return TSB_OK;
}

void tsb_debug_plan_free(tsb_debug_plan_s *d) { delete d; }

/* Tuning hook (not part of the stable ABI): threads per CTA for the 512-tet variant. */
void tsb_debug_set_threads_512(int nt) { tsb::set_threads_512(nt); }
void tsb_debug_set_skip_combine(int v) { tsb::set_skip_combine(v); }
void tsb_debug_set_pdl_tile(int v) { tsb::set_pdl_tile(v); }
void tsb_debug_set_exp_flags(int v) { tsb::set_exp_flags(v); }


}  // extern "C"
