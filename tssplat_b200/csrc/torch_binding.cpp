// tssplat-b200: the autograd bridge in C++ (built for real).
//
// What the reference's pybind11 module + energies/smooth_barrier.py:9-31 do per iteration -- forward (energy) and
// backward (dE/dx * grad_output) of SmoothnessBarrierFunc -- as a torch::autograd::Function over the C ABI of
// include/tssplat_b200.h.  Same semantics as tssplat_b200.tet_spheres_ext.forward/backward (one fused launch in
// forward, the gradient kept for ONE backward while x is unchanged, tsb_scale for a CUDA grad_output), without the
// Python interpreter between autograd and the launches.  The C entry points are handed over as addresses by the
// Python side (`bind`), so this module never loads a second copy of libtssplat_b200.so.
#include <atomic>
#include <cstdint>
#include <stdexcept>
#include <string>

#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>

#include "../../include/tssplat_b200.h"

namespace {

using energy_grad_fn = int (*)(tsb_handle_t, const float *, float, float, int32_t, float, const float *, float *, float *, void *);
using scale_fn = int (*)(const float *, int64_t, float, const float *, float *, void *);
using last_error_fn = const char *(*)(tsb_handle_t);

energy_grad_fn g_energy_grad = nullptr;
scale_fn g_scale = nullptr;
last_error_fn g_last_error = nullptr;
std::atomic<int64_t> g_epoch{0};       // bumped when parameters change behind autograd's back (optimizer steps on p.data)

// per-TetSpheres state (owned by the Python object through a capsule-like integer handle)
struct State {
  tsb_handle_t h = nullptr;
  int64_t n = 0;
  int device = 0;
  torch::Tensor ring;          // [32, 4] energies of the last 32 launches (same role as the Python ring)
  int64_t ring_i = 0;
};

// keeps a Python object alive from inside an autograd node (IValue payload); released under the GIL
struct OwnerHolder final : c10::ivalue::PyObjectHolder {
  explicit OwnerHolder(py::object o) : obj(std::move(o)) {}
  PyObject *getPyObject() override { return obj.ptr(); }
  c10::InferredType tryToInferType() override { return c10::InferredType("tssplat_b200 owner object"); }
  c10::IValue toIValue(const c10::TypePtr &, std::optional<int32_t>) override { TORCH_CHECK(false, "not convertible"); }
  std::string toStr() override { return "tssplat_b200.TetSpheres"; }
  std::vector<at::Tensor> extractTensors() override { return {}; }
  ~OwnerHolder() override {
    py::gil_scoped_acquire gil;
    obj.dec_ref();
    obj.ptr() = nullptr;
  }
  py::object obj;
};

void check(int rc, const State *S, const char *what) {
  if (rc == TSB_OK) return;
  const char *msg = g_last_error ? g_last_error(S ? S->h : nullptr) : "";
  throw std::runtime_error(std::string(what) + ": " + (msg ? msg : "") + " (code " + std::to_string(rc) + ")");
}

const torch::Tensor &checked_x(const State *S, const torch::Tensor &x, torch::Tensor &holder) {
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == torch::kFloat32, "vertexPositions must be a float32 CUDA tensor");
  TORCH_CHECK(x.get_device() == S->device, "vertexPositions is on cuda:", x.get_device(), ", TetSpheres on cuda:", S->device);
  TORCH_CHECK(x.numel() == 3 * S->n, "vertexPositions has ", x.numel(), " entries, expected ", 3 * S->n);
  if (x.is_contiguous()) return x;
  holder = x.contiguous();
  return holder;
}

// one fused launch; returns (energy slot view [4], grad or undefined)
std::pair<torch::Tensor, torch::Tensor> launch(State *S, const torch::Tensor &x, double c1, double c2, int64_t order, float gradH,
                                               const float *gradH_dev, bool want_grad) {
  TORCH_CHECK(order == 2 || order == 4, "order must be 2 or 4");
  torch::Tensor holder;
  const torch::Tensor &xc = checked_x(S, x, holder);
  const int64_t i = S->ring_i;
  S->ring_i = (i + 1) & 31;
  torch::Tensor grad;
  if (want_grad) grad = torch::empty({S->n, 3}, xc.options());
  float *e = S->ring.data_ptr<float>() + 4 * i;
  void *st = c10::cuda::getCurrentCUDAStream(S->device).stream();
  check(g_energy_grad(S->h, xc.data_ptr<float>(), float(c1), float(c2), int32_t(order), gradH, gradH_dev, e,
                      want_grad ? grad.data_ptr<float>() : nullptr, st),
        S, "tssplat_b200 fused launch");
  return {S->ring.select(0, i), grad};
}

struct EnergyFunction : public torch::autograd::Function<EnergyFunction> {
  // `owner`: the Python TetSpheres that owns `state` and the C handle; the graph node keeps it alive (like
  // ctx.constants does on the Python route), so a backward that outlives the energy module stays valid
  static torch::Tensor forward(torch::autograd::AutogradContext *ctx, const torch::Tensor &x, int64_t state, double c1, double c2,
                               int64_t order, const py::object &owner) {
    State *S = reinterpret_cast<State *>(state);
    ctx->saved_data["owner"] = c10::IValue(c10::intrusive_ptr<c10::ivalue::PyObjectHolder>(c10::make_intrusive<OwnerHolder>(owner)));
    const bool want = x.requires_grad();
    auto eg = launch(S, x, c1, c2, order, 1.f, nullptr, want);
    ctx->save_for_backward({x});
    ctx->saved_data["state"] = state;
    ctx->saved_data["c1"] = c1;
    ctx->saved_data["c2"] = c2;
    ctx->saved_data["order"] = order;
    ctx->saved_data["ptr"] = int64_t(reinterpret_cast<intptr_t>(x.data_ptr()));
    ctx->saved_data["version"] = int64_t(x._version());
    ctx->saved_data["epoch"] = g_epoch.load();
    if (want) ctx->saved_data["grad"] = eg.second;
    return eg.first.select(0, 0);
  }

  static torch::autograd::tensor_list backward(torch::autograd::AutogradContext *ctx, torch::autograd::tensor_list grad_outputs) {
    const torch::Tensor &go = grad_outputs[0];
    if (!go.defined()) return {torch::Tensor(), torch::Tensor(), torch::Tensor(), torch::Tensor(), torch::Tensor(), torch::Tensor()};
    State *S = reinterpret_cast<State *>(ctx->saved_data["state"].toInt());
    const auto saved = ctx->get_saved_variables();
    const torch::Tensor &x = saved[0];
    const double c1 = ctx->saved_data["c1"].toDouble(), c2 = ctx->saved_data["c2"].toDouble();
    const int64_t order = ctx->saved_data["order"].toInt();
    // gradH: a CUDA scalar stays on the device (no sync), a host scalar is read here
    torch::Tensor go_dev;
    float gh = 1.f;
    if (go.is_cuda()) {
      go_dev = (go.scalar_type() == torch::kFloat32 && go.get_device() == S->device) ? go : go.to(x.options());
    } else {
      gh = go.item<float>();
    }
    const bool fresh = ctx->saved_data.count("grad") && ctx->saved_data["grad"].isTensor() &&
                       ctx->saved_data["ptr"].toInt() == int64_t(reinterpret_cast<intptr_t>(x.data_ptr())) &&
                       ctx->saved_data["version"].toInt() == int64_t(x._version()) && ctx->saved_data["epoch"].toInt() == g_epoch.load();
    torch::Tensor g;
    if (fresh) {
      g = ctx->saved_data["grad"].toTensor();
      ctx->saved_data.erase("grad");                 // single use: a second backward recomputes at the current x
      void *st = c10::cuda::getCurrentCUDAStream(S->device).stream();
      if (go_dev.defined())
        check(g_scale(g.data_ptr<float>(), g.numel(), 1.f, go_dev.data_ptr<float>(), g.data_ptr<float>(), st), nullptr, "tsb_scale");
      else if (gh != 1.f)
        check(g_scale(g.data_ptr<float>(), g.numel(), gh, nullptr, g.data_ptr<float>(), st), nullptr, "tsb_scale");
    } else {
      g = launch(S, x, c1, c2, order, gh, go_dev.defined() ? go_dev.data_ptr<float>() : nullptr, true).second;
    }
    return {g.view(x.sizes()), torch::Tensor(), torch::Tensor(), torch::Tensor(), torch::Tensor(), torch::Tensor()};
  }
};

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "tssplat_b200: C++ autograd bridge over the C ABI (include/tssplat_b200.h)";
  m.def("bind", [](int64_t energy_grad, int64_t scale, int64_t last_error) {
    g_energy_grad = reinterpret_cast<energy_grad_fn>(energy_grad);
    g_scale = reinterpret_cast<scale_fn>(scale);
    g_last_error = reinterpret_cast<last_error_fn>(last_error);
  }, "addresses of tsb_energy_grad, tsb_scale and tsb_last_error in the already loaded libtssplat_b200.so");
  m.def("state_new", [](int64_t handle, int64_t n, int64_t device) {
    TORCH_CHECK(g_energy_grad && g_scale, "bind() first");
    State *S = new State;
    S->h = reinterpret_cast<tsb_handle_t>(handle);
    S->n = n;
    S->device = int(device);
    S->ring = torch::zeros({32, 4}, torch::TensorOptions().dtype(torch::kFloat32).device(torch::kCUDA, int(device)));
    return int64_t(reinterpret_cast<intptr_t>(S));
  });
  m.def("state_free", [](int64_t state) { delete reinterpret_cast<State *>(state); });
  m.def("note_parameters_changed", []() { g_epoch.fetch_add(1); });
  m.def("energy", [](const torch::Tensor &x, int64_t state, double c1, double c2, int64_t order, const py::object &owner) {
    return EnergyFunction::apply(x, state, c1, c2, order, owner);
  }, "differentiable E(x) as a 0-dim tensor on x's device (SmoothnessBarrierFunc.apply); `owner` is kept alive by the graph");
}
