// torch_ext.cpp - PyTorch C++ extension over the C ABI (include/md2_loss.h).
//
// Validates tensors the way the reference never does (dtype / device /
// contiguity / shape -> RuntimeError), allocates outputs and the workspace through the
// caching allocator, takes the current CUDA stream and calls the extern "C" entry points
// of libmd2loss.so.  No host synchronisation, no compute here.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <vector>

#include "../../include/md2_loss.h"

namespace {

using torch::Tensor;

void check_f32(const Tensor& t, const char* name, const c10::Device& dev) {
  TORCH_CHECK(t.defined(), name, " is undefined");
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor (the fused loss has no CPU path)");
  TORCH_CHECK(t.device() == dev, name, " is on ", t.device(), ", expected ", dev);
  TORCH_CHECK(t.scalar_type() == torch::kFloat32, name, " must be float32, got ", t.scalar_type());
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous (NCHW)");
}

void check_shape(const Tensor& t, const char* name, std::initializer_list<int64_t> shape) {
  TORCH_CHECK(t.sizes() == c10::IntArrayRef(shape.begin(), shape.size()), name, " has shape ", t.sizes(),
              ", expected ", c10::IntArrayRef(shape.begin(), shape.size()));
}

struct Packed {
  md2_cfg cfg;
  md2_inputs in;
};

Packed pack(const Tensor& target, const std::vector<Tensor>& sources, const std::vector<Tensor>& disps,
            const std::vector<Tensor>& color_pyr, const Tensor& K, const Tensor& inv_K,
            const std::vector<Tensor>& Ts, const std::vector<Tensor>& noise, int64_t seed, bool automask,
            double min_depth, double max_depth, double disp_smoothness,
            const c10::optional<Tensor>& seed_tensor = c10::nullopt) {
  TORCH_CHECK(target.dim() == 4 && target.size(1) == 3, "target must be [B,3,H,W], got ", target.sizes());
  const auto dev = target.device();
  const int64_t B = target.size(0), H = target.size(2), W = target.size(3);
  const int64_t S = (int64_t)sources.size(), ns = (int64_t)disps.size();
  TORCH_CHECK(S >= 1 && S <= MD2_MAX_SOURCES, "1..", MD2_MAX_SOURCES, " source frames supported, got ", S);
  TORCH_CHECK(ns >= 1 && ns <= MD2_MAX_SCALES, "1..", MD2_MAX_SCALES, " scales supported, got ", ns);
  TORCH_CHECK((int64_t)Ts.size() == S, "need one transformation per source");
  TORCH_CHECK((int64_t)color_pyr.size() == ns, "need one colour pyramid level per scale");
  TORCH_CHECK(noise.empty() || (int64_t)noise.size() == ns, "noise must be empty or one tensor per scale");
  check_f32(target, "target", dev);
  check_f32(K, "K", dev);
  check_f32(inv_K, "inv_K", dev);
  check_shape(K, "K", {B, 4, 4});
  check_shape(inv_K, "inv_K", {B, 4, 4});
  Packed p;
  memset(&p, 0, sizeof(p));
  p.cfg.B = (int)B; p.cfg.H = (int)H; p.cfg.W = (int)W; p.cfg.S = (int)S; p.cfg.num_scales = (int)ns;
  p.cfg.automask = automask ? 1 : 0;
  p.cfg.min_depth = min_depth; p.cfg.max_depth = max_depth;
  p.cfg.disp_smoothness = disp_smoothness; p.cfg.eps_proj = 1e-7;
  p.in.target = target.data_ptr<float>();
  p.in.K = K.data_ptr<float>();
  p.in.inv_K = inv_K.data_ptr<float>();
  for (int64_t f = 0; f < S; ++f) {
    check_f32(sources[f], "source", dev);
    check_shape(sources[f], "source", {B, 3, H, W});
    check_f32(Ts[f], "T", dev);
    check_shape(Ts[f], "T", {B, 4, 4});
    p.in.sources[f] = sources[f].data_ptr<float>();
    p.in.T[f] = Ts[f].data_ptr<float>();
  }
  for (int64_t s = 0; s < ns; ++s) {
    TORCH_CHECK(H % (1 << s) == 0 && W % (1 << s) == 0, "H and W must be divisible by 2^(num_scales-1)");
    check_f32(disps[s], "disp", dev);
    check_shape(disps[s], "disp", {B, 1, H >> s, W >> s});
    check_f32(color_pyr[s], "color_pyr", dev);
    check_shape(color_pyr[s], "color_pyr", {B, 3, H >> s, W >> s});
    p.in.disp[s] = disps[s].data_ptr<float>();
    p.in.color_pyr[s] = color_pyr[s].data_ptr<float>();
    if (!noise.empty()) {
      check_f32(noise[s], "noise", dev);
      check_shape(noise[s], "noise", {B, S, H, W});
      p.in.noise[s] = noise[s].data_ptr<float>();
    }
  }
  p.in.seed = (uint64_t)seed;
  if (seed_tensor.has_value() && seed_tensor->defined()) {
    const Tensor& st = *seed_tensor;
    TORCH_CHECK(st.device() == dev && st.scalar_type() == torch::kInt64 && st.numel() == 1,
                "seed_tensor must be a one-element int64 tensor on the same device");
    p.in.seed_dev = (const uint64_t*)st.data_ptr<int64_t>();
  }
  return p;
}

void check_rc(int rc, const char* what) {
  TORCH_CHECK(rc <= 0, what, ": CUDA error ", rc, " (", cudaGetErrorString((cudaError_t)rc), ")");
  TORCH_CHECK(rc == 0, what, ": invalid argument (md2 error ", rc, ")");
}

Tensor workspace_for(const md2_cfg& cfg, const c10::Device& dev) {
  const size_t bytes = md2_workspace_bytes(&cfg);
  TORCH_CHECK(bytes > 0, "md2_workspace_bytes rejected the configuration");
  return torch::empty({(int64_t)bytes}, torch::TensorOptions().dtype(torch::kUInt8).device(dev));
}

// returns [loss(1), per_pixel, argmin, depth]
std::vector<Tensor> loss_forward(const Tensor& target, const std::vector<Tensor>& sources,
                                 const std::vector<Tensor>& disps, const std::vector<Tensor>& color_pyr,
                                 const Tensor& K, const Tensor& inv_K, const std::vector<Tensor>& Ts,
                                 const std::vector<Tensor>& noise, int64_t seed, bool automask, double min_depth,
                                 double max_depth, double disp_smoothness, bool want_per_pixel,
                                 const c10::optional<Tensor>& seed_tensor) {
  Packed p = pack(target, sources, disps, color_pyr, K, inv_K, Ts, noise, seed, automask, min_depth, max_depth,
                  disp_smoothness, seed_tensor);
  const c10::cuda::CUDAGuard guard(target.device());
  const auto f32 = target.options();
  const int64_t ns = p.cfg.num_scales, B = p.cfg.B, H = p.cfg.H, W = p.cfg.W;
  Tensor loss = torch::empty({1}, f32);
  Tensor per_px = want_per_pixel ? torch::empty({ns, B, H, W}, f32) : Tensor();
  Tensor argmin = torch::empty({ns, B, H, W}, f32.dtype(torch::kUInt8));
  Tensor depth = torch::empty({ns, B, 1, H, W}, f32);
  Tensor ws = workspace_for(p.cfg, target.device());
  md2_outputs out{loss.data_ptr<float>(), want_per_pixel ? per_px.data_ptr<float>() : nullptr,
                  argmin.data_ptr<uint8_t>(), depth.data_ptr<float>()};
  check_rc(md2_loss_forward(&p.cfg, &p.in, &out, ws.data_ptr(), at::cuda::getCurrentCUDAStream().stream()),
           "md2_loss_forward");
  return {loss, per_px, argmin, depth};
}

// returns [loss(1), per_pixel, argmin, depth, grad_disp..., grad_T...]
std::vector<Tensor> loss_forward_backward(const Tensor& target, const std::vector<Tensor>& sources,
                                          const std::vector<Tensor>& disps, const std::vector<Tensor>& color_pyr,
                                          const Tensor& K, const Tensor& inv_K, const std::vector<Tensor>& Ts,
                                          const std::vector<Tensor>& noise, int64_t seed, bool automask,
                                          double min_depth, double max_depth, double disp_smoothness,
                                          bool want_per_pixel, double grad_loss,
                                          const c10::optional<Tensor>& seed_tensor) {
  Packed p = pack(target, sources, disps, color_pyr, K, inv_K, Ts, noise, seed, automask, min_depth, max_depth,
                  disp_smoothness, seed_tensor);
  const c10::cuda::CUDAGuard guard(target.device());
  const auto f32 = target.options();
  const int64_t ns = p.cfg.num_scales, B = p.cfg.B, H = p.cfg.H, W = p.cfg.W, S = p.cfg.S;
  Tensor loss = torch::empty({1}, f32);
  Tensor per_px = want_per_pixel ? torch::empty({ns, B, H, W}, f32) : Tensor();
  Tensor argmin = torch::empty({ns, B, H, W}, f32.dtype(torch::kUInt8));
  Tensor depth = torch::empty({ns, B, 1, H, W}, f32);
  Tensor ws = workspace_for(p.cfg, target.device());
  md2_outputs out{loss.data_ptr<float>(), want_per_pixel ? per_px.data_ptr<float>() : nullptr,
                  argmin.data_ptr<uint8_t>(), depth.data_ptr<float>()};
  md2_grads g;
  memset(&g, 0, sizeof(g));
  std::vector<Tensor> ret{loss, per_px, argmin, depth};
  for (int64_t s = 0; s < ns; ++s) {
    Tensor gd = torch::empty_like(disps[s]);
    g.grad_disp[s] = gd.data_ptr<float>();
    ret.push_back(gd);
  }
  for (int64_t f = 0; f < S; ++f) {
    Tensor gt = torch::empty({B, 4, 4}, f32);
    g.grad_T[f] = gt.data_ptr<float>();
    ret.push_back(gt);
  }
  check_rc(md2_loss_forward_backward(&p.cfg, &p.in, &out, &g, (float)grad_loss, ws.data_ptr(),
                                     at::cuda::getCurrentCUDAStream().stream()),
           "md2_loss_forward_backward");
  return ret;
}

// returns [grad_disp..., grad_T...]
std::vector<Tensor> loss_backward(const Tensor& target, const std::vector<Tensor>& sources,
                                  const std::vector<Tensor>& disps, const std::vector<Tensor>& color_pyr,
                                  const Tensor& K, const Tensor& inv_K, const std::vector<Tensor>& Ts,
                                  bool automask, double min_depth, double max_depth, double disp_smoothness,
                                  const Tensor& argmin, const Tensor& grad_loss) {
  Packed p = pack(target, sources, disps, color_pyr, K, inv_K, Ts, {}, 0, automask, min_depth, max_depth,
                  disp_smoothness);
  const c10::cuda::CUDAGuard guard(target.device());
  const auto f32 = target.options();
  const int64_t ns = p.cfg.num_scales, B = p.cfg.B, H = p.cfg.H, W = p.cfg.W, S = p.cfg.S;
  TORCH_CHECK(argmin.is_cuda() && argmin.scalar_type() == torch::kUInt8 && argmin.is_contiguous(),
              "argmin must be a contiguous CUDA uint8 tensor");
  check_shape(argmin, "argmin", {ns, B, H, W});
  Tensor gl = grad_loss.to(f32).reshape({1}).contiguous();
  Tensor ws = workspace_for(p.cfg, target.device());
  md2_grads g;
  memset(&g, 0, sizeof(g));
  std::vector<Tensor> ret;
  for (int64_t s = 0; s < ns; ++s) {
    Tensor gd = torch::empty_like(disps[s]);
    g.grad_disp[s] = gd.data_ptr<float>();
    ret.push_back(gd);
  }
  for (int64_t f = 0; f < S; ++f) {
    Tensor gt = torch::empty({B, 4, 4}, f32);
    g.grad_T[f] = gt.data_ptr<float>();
    ret.push_back(gt);
  }
  check_rc(md2_loss_backward(&p.cfg, &p.in, argmin.data_ptr<uint8_t>(), gl.data_ptr<float>(), &g, ws.data_ptr(),
                             at::cuda::getCurrentCUDAStream().stream()),
           "md2_loss_backward");
  return ret;
}

Tensor pose_forward(const Tensor& axisangle, const Tensor& translation, bool invert) {
  const auto dev = axisangle.device();
  check_f32(axisangle, "axisangle", dev);
  check_f32(translation, "translation", dev);
  TORCH_CHECK(axisangle.numel() % 3 == 0 && axisangle.numel() == translation.numel(),
              "axisangle / translation must both be [n,1,3]");
  const c10::cuda::CUDAGuard guard(dev);
  const int64_t n = axisangle.numel() / 3;
  Tensor M = torch::empty({n, 4, 4}, axisangle.options());
  check_rc(md2_pose_forward((int)n, axisangle.data_ptr<float>(), translation.data_ptr<float>(), invert ? 1 : 0,
                            M.data_ptr<float>(), at::cuda::getCurrentCUDAStream().stream()),
           "md2_pose_forward");
  return M;
}

std::vector<Tensor> pose_backward(const Tensor& axisangle, const Tensor& translation, bool invert,
                                  const Tensor& grad_M) {
  const auto dev = axisangle.device();
  check_f32(axisangle, "axisangle", dev);
  check_f32(translation, "translation", dev);
  Tensor gM = grad_M.contiguous();
  check_f32(gM, "grad_M", dev);
  const c10::cuda::CUDAGuard guard(dev);
  const int64_t n = axisangle.numel() / 3;
  check_shape(gM, "grad_M", {n, 4, 4});
  Tensor ga = torch::empty_like(axisangle), gt = torch::empty_like(translation);
  check_rc(md2_pose_backward((int)n, axisangle.data_ptr<float>(), translation.data_ptr<float>(), invert ? 1 : 0,
                             gM.data_ptr<float>(), ga.data_ptr<float>(), gt.data_ptr<float>(),
                             at::cuda::getCurrentCUDAStream().stream()),
           "md2_pose_backward");
  return {ga, gt};
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "B200-native fused view-synthesis loss (PyTorch binding of libmd2loss.so)";
  m.def("loss_forward", &loss_forward);
  m.def("loss_forward_backward", &loss_forward_backward);
  m.def("loss_backward", &loss_backward);
  m.def("pose_forward", &pose_forward);
  m.def("pose_backward", &pose_backward);
  m.def("version", []() { return std::string(md2_version()); });
  m.def("launches_per_step", [](int B, int H, int W, int S, int ns, bool bwd) {
    md2_cfg c;
    memset(&c, 0, sizeof(c));
    c.B = B; c.H = H; c.W = W; c.S = S; c.num_scales = ns; c.min_depth = 0.1; c.max_depth = 100.0;
    return md2_launches_per_step(&c, bwd ? 1 : 0);
  });
}
