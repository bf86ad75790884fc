// wl_torch_ops.cpp -- PyTorch-extension front over the C-ABI (include/wheeledlab_b200.h): `torch.ops.wheeledlab_b200.*`.
// Ops take at::Tensor arguments, check device / dtype / shape / contiguity with TORCH_CHECK, and enqueue the same C entry
// points on torch's CURRENT CUDA stream of the tensors' device (so they compose with torch streams and CUDA-graph capture).
// No kernel lives here and nothing is computed on the host: it is the binding a torch caller would use instead of ctypes.
// The handle (wl_sim*) travels as an int64, as it does through ctypes (WheeledSim._h).
//
// Reference surface these ops replace: ManagerBasedRLEnv.step / reset / get_observations of the registered gym ids
// (wheeledlab_tasks/__init__.py:14-63; isaaclab ManagerBasedRLEnv), see INTEGRATION.md section 2.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>

#include <tuple>

#include "../../include/wheeledlab_b200.h"

namespace {

wl_sim* sim_of(int64_t handle) {
    TORCH_CHECK(handle != 0, "wheeledlab_b200: null simulation handle");
    return reinterpret_cast<wl_sim*>(static_cast<intptr_t>(handle));
}

void check_cuda(const at::Tensor& t, at::ScalarType dt, const char* name) {
    TORCH_CHECK(t.is_cuda(), "wheeledlab_b200: `", name, "` must be a CUDA tensor (this library has no CPU path)");
    TORCH_CHECK(t.scalar_type() == dt, "wheeledlab_b200: `", name, "` must have dtype ", dt, ", got ", t.scalar_type());
    TORCH_CHECK(t.is_contiguous(), "wheeledlab_b200: `", name, "` must be contiguous");
}

void check_rc(int rc, const char* what) { TORCH_CHECK(rc == WL_OK, "wheeledlab_b200::", what, ": ", wl_last_error()); }

// env.step(): action [N,2] f32 -> obs [N,obs_dim] f32, rew [N] f32, terminated [N] u8, truncated [N] u8 (written in place);
// log [16] f32 optional (extras["log"] row).  step_counter as in wl_step (>= 0, or WL_DEVICE_COUNTER_PLUS(k)).
void step_out(int64_t handle, const at::Tensor& action, at::Tensor& obs, at::Tensor& rew, at::Tensor& terminated,
              at::Tensor& truncated, const c10::optional<at::Tensor>& log, int64_t step_counter) {
    wl_sim* sim = sim_of(handle);
    const int64_t d = wl_obs_dim(sim);
    check_cuda(action, at::kFloat, "action"); check_cuda(obs, at::kFloat, "obs"); check_cuda(rew, at::kFloat, "rew");
    check_cuda(terminated, at::kByte, "terminated"); check_cuda(truncated, at::kByte, "truncated");
    TORCH_CHECK(action.dim() == 2 && action.size(1) == 2, "wheeledlab_b200::step: action must be [N, 2]");
    const int64_t n = action.size(0);
    TORCH_CHECK(obs.dim() == 2 && obs.size(0) == n && obs.size(1) == d, "wheeledlab_b200::step: obs must be [N, ", d, "]");
    TORCH_CHECK(rew.numel() == n && terminated.numel() == n && truncated.numel() == n, "wheeledlab_b200::step: rew / terminated / truncated must have N elements");
    TORCH_CHECK(obs.get_device() == action.get_device() && rew.get_device() == action.get_device(), "wheeledlab_b200::step: tensors on different devices");
    float* lp = nullptr;
    if (log.has_value()) {
        check_cuda(*log, at::kFloat, "log");
        TORCH_CHECK(log->numel() >= 16, "wheeledlab_b200::step: log must hold 16 floats");
        lp = log->data_ptr<float>();
    }
    c10::cuda::CUDAGuard guard(action.device());
    check_rc(wl_step(sim, action.data_ptr<float>(), obs.data_ptr<float>(), rew.data_ptr<float>(), terminated.data_ptr<uint8_t>(),
                     truncated.data_ptr<uint8_t>(), lp, step_counter, at::cuda::getCurrentCUDAStream().stream()), "step");
}

// allocating form: returns (obs, rew, terminated, truncated)
std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor> step(int64_t handle, const at::Tensor& action, int64_t step_counter) {
    wl_sim* sim = sim_of(handle);
    check_cuda(action, at::kFloat, "action");
    TORCH_CHECK(action.dim() == 2 && action.size(1) == 2, "wheeledlab_b200::step: action must be [N, 2]");
    const int64_t n = action.size(0), d = wl_obs_dim(sim);
    auto f = action.options();
    at::Tensor obs = at::empty({n, d}, f);
    at::Tensor rew = at::empty({n}, f);
    at::Tensor term = at::empty({n}, f.dtype(at::kByte));
    at::Tensor trunc = at::empty({n}, f.dtype(at::kByte));
    step_out(handle, action, obs, rew, term, trunc, c10::nullopt, step_counter);
    return std::make_tuple(obs, rew, term, trunc);
}

// get_observations(): obs [N, obs_dim] written in place
void observe_out(int64_t handle, at::Tensor& obs, int64_t step_counter, int64_t call_idx) {
    wl_sim* sim = sim_of(handle);
    check_cuda(obs, at::kFloat, "obs");
    TORCH_CHECK(obs.dim() == 2 && obs.size(1) == wl_obs_dim(sim), "wheeledlab_b200::observe: obs must be [N, ", wl_obs_dim(sim), "]");
    c10::cuda::CUDAGuard guard(obs.device());
    check_rc(wl_observe(sim, obs.data_ptr<float>(), step_counter, static_cast<int32_t>(call_idx), at::cuda::getCurrentCUDAStream().stream()), "observe");
}

// reset(env_ids): env_ids int64 [M] on the device, or None for all envs; `like` only names the device / stream
void reset(int64_t handle, const c10::optional<at::Tensor>& env_ids, const at::Tensor& like, int64_t step_counter) {
    wl_sim* sim = sim_of(handle);
    TORCH_CHECK(like.is_cuda(), "wheeledlab_b200::reset: `like` must be a CUDA tensor of the simulation's device");
    const int64_t* ids = nullptr;
    int32_t n_ids = 0;
    if (env_ids.has_value()) {
        check_cuda(*env_ids, at::kLong, "env_ids");
        ids = env_ids->data_ptr<int64_t>();
        n_ids = static_cast<int32_t>(env_ids->numel());
    }
    c10::cuda::CUDAGuard guard(like.device());
    check_rc(wl_reset(sim, ids, n_ids, step_counter, at::cuda::getCurrentCUDAStream().stream()), "reset");
}

}  // namespace

TORCH_LIBRARY(wheeledlab_b200, m) {
    m.def("step_out(int handle, Tensor action, Tensor(a!) obs, Tensor(b!) rew, Tensor(c!) terminated, Tensor(d!) truncated, Tensor(e!)? log, int step_counter) -> ()");
    m.def("step(int handle, Tensor action, int step_counter) -> (Tensor, Tensor, Tensor, Tensor)");
    m.def("observe_out(int handle, Tensor(a!) obs, int step_counter, int call_idx) -> ()");
    m.def("reset(int handle, Tensor? env_ids, Tensor like, int step_counter) -> ()");
}

TORCH_LIBRARY_IMPL(wheeledlab_b200, CUDA, m) {
    m.impl("step_out", &step_out);
    m.impl("step", &step);
    m.impl("observe_out", &observe_out);
    m.impl("reset", &reset);
}
