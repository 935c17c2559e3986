// binding.cpp — native (pybind11 / libtorch) form of the reference's extension module
// `warprnnt_pytorch.warp_rnnt` (pytorch_binding/src/binding.cpp:12-19,84-91,157-162), bound to this
// repository's libwarprnnt.so.  Same function names, argument order and return value; the only
// changes against the reference source are the ones current PyTorch forces (no THC: the workspace
// comes from the caching allocator; C++17) and that a non-success status raises instead of being
// dropped.  The Python package works without this module (warp_rnnt.py binds the same C-ABI with
// ctypes); it exists for callers that want the compiled extension-module form.
#include <torch/extension.h>

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "rnnt.h"

namespace {

void check(const torch::Tensor& t, const char* name, bool want_cuda) {
    TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
    TORCH_CHECK(t.is_cuda() == want_cuda, name, want_cuda ? " must be a CUDA tensor" : " must be a CPU tensor");
}

}  // namespace

int gpu_rnnt(torch::Tensor acts, torch::Tensor labels, torch::Tensor input_lengths,
             torch::Tensor label_lengths, torch::Tensor costs, torch::Tensor grads, int blank_label,
             int num_threads) {
    check(acts, "acts", true);
    check(labels, "labels", true);
    check(input_lengths, "input_lengths", true);
    check(label_lengths, "label_lengths", true);
    TORCH_CHECK(acts.dim() == 4, "acts must be 4D");
    const int N = acts.size(0), T = acts.size(1), U = acts.size(2), V = acts.size(3);
    c10::cuda::CUDAGuard guard(acts.device());

    rnntOptions options{};
    options.maxT = T;
    options.maxU = U;
    options.blank_label = blank_label;
    options.loc = RNNT_GPU;
    options.stream = at::cuda::getCurrentCUDAStream();
    options.num_threads = num_threads;

    const bool f64 = acts.scalar_type() == torch::kFloat64;
    TORCH_CHECK(f64 || acts.scalar_type() == torch::kFloat32, "unsupported data type");
    size_t bytes = 0;
    TORCH_CHECK(get_workspace_size(T, U, N, true, &bytes, f64 ? sizeof(double) : sizeof(float)) ==
                    RNNT_STATUS_SUCCESS,
                "get_workspace_size failed");
    auto ws = torch::empty({(int64_t)bytes}, acts.options().dtype(torch::kUInt8));
    // U == 1: no labels at all, the ABI still wants a non-null pointer
    auto lab = labels.numel() ? labels : torch::zeros({1}, labels.options());
    void* g = grads.numel() ? grads.data_ptr() : nullptr;
    const rnntStatus_t st =
        f64 ? compute_rnnt_loss_fp64(acts.data_ptr<double>(), static_cast<double*>(g), lab.data_ptr<int>(),
                                     label_lengths.data_ptr<int>(), input_lengths.data_ptr<int>(), V, N,
                                     costs.data_ptr<double>(), ws.data_ptr(), options)
            : compute_rnnt_loss(acts.data_ptr<float>(), static_cast<float*>(g), lab.data_ptr<int>(),
                                label_lengths.data_ptr<int>(), input_lengths.data_ptr<int>(), V, N,
                                costs.data_ptr<float>(), ws.data_ptr(), options);
    TORCH_CHECK(st == RNNT_STATUS_SUCCESS, "compute_rnnt_loss: ", rnntGetStatusString(st));
    return 0;
}

int cpu_rnnt(torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, int,
             int) {
    TORCH_CHECK(false, "warprnnt_pytorch (B200 build): cpu_rnnt is not available - there is no CPU path");
    return -1;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("cpu_rnnt", &cpu_rnnt, "RNNT CPU version (not available in the B200 build)");
    m.def("gpu_rnnt", &gpu_rnnt, "RNNT GPU version");
}
