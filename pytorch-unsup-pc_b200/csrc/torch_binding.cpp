// C++ autograd binding of the whole-path projection (dpc_project_fwd / dpc_project_bwd and their
// replica-aware variants) -- the same C ABI the ctypes path calls (include/dpc_b200.h), bound as a
// torch::autograd::Function so that neither pass runs Python: the eager step costs the autograd
// engine and two C calls instead of ~250 us of interpreter work (the kernels take ~150 us).
//
// Optional accelerator: pytorch-unsup-pc_b200/ops.py uses it when lib/dpc_b200_torch.so is there
// and falls back to the ctypes Function (same kernels, same results) when it is not.
// Argument validation stays in the Python mirror (point_cloud.py); this file trusts contiguous
// fp32 CUDA tensors.
#include <torch/extension.h>

#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include "../../include/dpc_b200.h"

namespace {

using torch::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;
using OptTensor = c10::optional<Tensor>;

const float *fptr(const Tensor &t) { return t.defined() ? t.data_ptr<float>() : nullptr; }
float *fptr_mut(Tensor &t) { return t.defined() ? t.data_ptr<float>() : nullptr; }
Tensor unwrap(const OptTensor &t) { return (t.has_value() && t->defined()) ? *t : Tensor(); }

void check(int status, const char *what) {
  TORCH_CHECK(status == 0, "dpc_b200 ", what, " failed (status ", status, "): ", dpc_last_error());
}

// geom: P, N, Vz, V, camera_distance, focal_length, max_depth, drc_clip, drc_logsum, flip_y
dpc_params make_params(const std::vector<double> &g, bool want_voxels, bool want_probs) {
  dpc_params p;
  p.P = (int32_t)g[0]; p.N = (int32_t)g[1]; p.Vz = (int32_t)g[2]; p.V = (int32_t)g[3];
  p.camera_distance = g[4]; p.focal_length = g[5]; p.max_depth = g[6]; p.drc_clip = g[7];
  p.drc_logsum = (int32_t)g[8]; p.flip_y = (int32_t)g[9];
  p.outputs = (want_voxels ? DPC_OUT_VOXELS : 0) | (want_probs ? DPC_OUT_PROBS : 0);
  return p;
}

struct HostTaps {
  const float *p[3] = {nullptr, nullptr, nullptr};
  int n[3] = {0, 0, 0};
  HostTaps(const Tensor &kx, const Tensor &ky, const Tensor &kz) {
    const Tensor *k[3] = {&kx, &ky, &kz};
    for (int i = 0; i < 3; ++i)
      if (k[i]->defined()) {
        p[i] = k[i]->data_ptr<float>();
        n[i] = (int)k[i]->numel();
      }
  }
};

class ProjectFunction : public torch::autograd::Function<ProjectFunction> {
 public:
  static variable_list forward(AutogradContext *ctx, Tensor points, Tensor quat, OptTensor trans_,
                               OptTensor focal_, OptTensor scale_, std::vector<double> geom,
                               OptTensor kx_, OptTensor ky_, OptTensor kz_, bool want_voxels,
                               bool want_probs, int64_t mode, bool plane_local, int64_t replicas,
                               int64_t n_src, OptTensor sel_) {
    const Tensor trans = unwrap(trans_), focal = unwrap(focal_), scale = unwrap(scale_);
    const Tensor kx = unwrap(kx_), ky = unwrap(ky_), kz = unwrap(kz_), sel = unwrap(sel_);
    const dpc_params p = make_params(geom, want_voxels, want_probs);
    const c10::cuda::CUDAGuard guard(points.device());
    void *stream = c10::cuda::getCurrentCUDAStream(points.device().index()).stream();
    const auto f32 = points.options();
    const auto u8 = f32.dtype(torch::kUInt8);
    Tensor tr_pc = torch::empty({p.P, p.N, 3}, f32);
    Tensor mask = torch::empty({p.P, p.V, p.V}, f32), depth = torch::empty({p.P, p.V, p.V}, f32);
    Tensor voxels = want_voxels ? torch::empty({p.P, p.Vz, p.V, p.V}, f32) : Tensor();
    Tensor probs = want_probs ? torch::empty({p.Vz + 1, p.P, p.V, p.V}, f32) : Tensor();
    // saved state, one allocation: blurred grid | clamp bits | cell records + ray checkpoints
    const bool use_cells = plane_local;   // both scatter modes save the plane-local state
    const int64_t n_grid = (int64_t)p.P * p.Vz * p.V * p.V * 4;
    const int64_t n_bits = (int64_t)p.P * p.Vz * p.V * (p.V / 32) * 4;
    const int64_t n_cells = use_cells ? (int64_t)dpc_cells_bytes(&p) : 0;
    Tensor state = torch::empty({n_grid + n_bits + n_cells}, u8);
    uint8_t *base = state.data_ptr<uint8_t>();
    Tensor ws = torch::empty({(int64_t)dpc_workspace_bytes(&p)}, u8);
    const HostTaps k(kx, ky, kz);
    int st;
    if (replicas > 0)
      st = dpc_project_replicated_fwd(
          &p, (int)replicas, (int)n_src, sel.defined() ? sel.data_ptr<int32_t>() : nullptr,
          fptr(points), fptr(quat), fptr(trans), fptr(focal), fptr(scale), k.p[0], k.n[0], k.p[1],
          k.n[1], k.p[2], k.n[2], (int)mode, fptr_mut(tr_pc), (float *)base,
          (uint32_t *)(base + n_grid), use_cells ? base + n_grid + n_bits : nullptr, fptr_mut(mask),
          fptr_mut(depth), fptr_mut(voxels), fptr_mut(probs), ws.data_ptr(), (size_t)ws.numel(),
          stream);
    else
      st = dpc_project_fwd(&p, fptr(points), fptr(quat), fptr(trans), fptr(focal), fptr(scale),
                           k.p[0], k.n[0], k.p[1], k.n[1], k.p[2], k.n[2], (int)mode,
                           fptr_mut(tr_pc), (float *)base, (uint32_t *)(base + n_grid),
                           use_cells ? base + n_grid + n_bits : nullptr, fptr_mut(mask),
                           fptr_mut(depth), fptr_mut(voxels), fptr_mut(probs), ws.data_ptr(),
                           (size_t)ws.numel(), stream);
    check(st, "project_fwd");
    ctx->save_for_backward({points, quat, trans, focal, scale, state, sel, kx, ky, kz});
    ctx->saved_data["geom"] = geom;
    ctx->saved_data["flags"] = std::vector<int64_t>{want_voxels, want_probs, use_cells, replicas,
                                                    n_src, n_grid, n_bits};
    ctx->set_materialize_grads(false);
    // (an autograd output must be a defined tensor: the optional ones are appended when asked for)
    variable_list out = {mask, depth, tr_pc};
    if (want_voxels) out.push_back(voxels);
    if (want_probs) out.push_back(probs);
    return out;
  }

  static variable_list backward(AutogradContext *ctx, variable_list grads) {
    const auto saved = ctx->get_saved_variables();
    const Tensor &points = saved[0], &quat = saved[1], &trans = saved[2], &focal = saved[3],
                 &scale = saved[4], &state = saved[5], &sel = saved[6];
    const std::vector<double> geom = ctx->saved_data["geom"].toDoubleVector();
    const std::vector<int64_t> fl = ctx->saved_data["flags"].toIntVector();
    const bool use_cells = fl[2] != 0;
    const int64_t replicas = fl[3], n_src = fl[4], n_grid = fl[5], n_bits = fl[6];
    const dpc_params p = make_params(geom, fl[0] != 0, fl[1] != 0);
    const c10::cuda::CUDAGuard guard(points.device());
    void *stream = c10::cuda::getCurrentCUDAStream(points.device().index()).stream();
    auto grad = [&](size_t i) {
      return grads[i].defined() ? grads[i].to(torch::kFloat32).contiguous() : Tensor();
    };
    const Tensor g_mask = grad(0), g_depth = grad(1), g_trpc = grad(2);
    const Tensor g_voxels = fl[0] ? grad(3) : Tensor();
    const Tensor g_probs = fl[1] ? grad(fl[0] ? 4 : 3) : Tensor();
    const auto f32 = points.options();
    const auto u8 = f32.dtype(torch::kUInt8);
    Tensor g_grid = torch::empty({n_grid}, u8);
    Tensor ws = torch::empty({(int64_t)dpc_workspace_bytes(&p)}, u8);
    Tensor g_quat = torch::empty({p.P, 4}, f32);
    Tensor g_trans = trans.defined() ? torch::empty({p.P, 3}, f32) : Tensor();
    Tensor g_focal = focal.defined() ? torch::empty({p.P}, f32) : Tensor();
    Tensor g_scale = scale.defined() ? torch::empty({p.P}, f32) : Tensor();
    uint8_t *base = state.data_ptr<uint8_t>();
    const HostTaps k(saved[7], saved[8], saved[9]);
    Tensor g_points;
    int st;
    if (replicas > 0) {
      g_points = torch::empty({p.P / replicas, n_src, 3}, f32);
      Tensor g_rep = torch::empty({p.P, p.N, 3}, f32);
      Tensor inv = sel.defined() ? torch::empty({p.P, n_src}, f32.dtype(torch::kInt32)) : Tensor();
      st = dpc_project_replicated_bwd(
          &p, (int)replicas, (int)n_src, sel.defined() ? sel.data_ptr<int32_t>() : nullptr,
          fptr(points), fptr(quat), fptr(trans), fptr(focal), fptr(scale), k.p[0], k.n[0], k.p[1],
          k.n[1], k.p[2], k.n[2], (const float *)base, (const uint32_t *)(base + n_grid),
          use_cells ? base + n_grid + n_bits : nullptr, fptr(g_mask), fptr(g_depth), fptr(g_probs),
          fptr(g_voxels), fptr(g_trpc), (float *)g_grid.data_ptr(), fptr_mut(g_rep),
          inv.defined() ? inv.data_ptr<int32_t>() : nullptr, fptr_mut(g_points), fptr_mut(g_quat),
          fptr_mut(g_trans), fptr_mut(g_focal), fptr_mut(g_scale), ws.data_ptr(),
          (size_t)ws.numel(), stream);
    } else {
      g_points = torch::empty({p.P, p.N, 3}, f32);
      st = dpc_project_bwd(&p, fptr(points), fptr(quat), fptr(trans), fptr(focal), fptr(scale),
                           k.p[0], k.n[0], k.p[1], k.n[1], k.p[2], k.n[2], (const float *)base,
                           (const uint32_t *)(base + n_grid),
                           use_cells ? base + n_grid + n_bits : nullptr, fptr(g_mask),
                           fptr(g_depth), fptr(g_probs), fptr(g_voxels), fptr(g_trpc),
                           (float *)g_grid.data_ptr(), fptr_mut(g_points), fptr_mut(g_quat),
                           fptr_mut(g_trans), fptr_mut(g_focal), fptr_mut(g_scale), ws.data_ptr(),
                           (size_t)ws.numel(), stream);
    }
    check(st, "project_bwd");
    const Tensor none;
    return {g_points, g_quat, g_trans, g_focal, g_scale, none, none, none, none,
            none, none, none, none, none, none, none};
  }
};

std::vector<Tensor> project(Tensor points, Tensor quat, OptTensor trans, OptTensor focal,
                            OptTensor scale, std::vector<double> geom, OptTensor kx, OptTensor ky,
                            OptTensor kz, bool want_voxels, bool want_probs, int64_t mode,
                            bool plane_local, int64_t replicas, int64_t n_src, OptTensor sel) {
  return ProjectFunction::apply(points, quat, trans, focal, scale, geom, kx, ky, kz, want_voxels,
                                want_probs, mode, plane_local, replicas, n_src, sel);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "C++ autograd binding of dpc_project_fwd / dpc_project_bwd (include/dpc_b200.h)";
  m.def("project", &project);
  m.def("abi_version", []() { return dpc_version(); });
}
