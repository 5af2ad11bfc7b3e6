// ORACLE glue (test infrastructure only): python bindings for the reference's correlation
// kernels (src/correlation_kernels.cu, src/altcorr_kernel.cu), compiled from /root/reference by
// oracle/build_ref.py.  Mirrors the entries of the reference's src/droid.cpp:280-327.
#include <torch/extension.h>
#include <vector>
std::vector<torch::Tensor> corr_index_cuda_forward(torch::Tensor volume, torch::Tensor coords, int radius);
std::vector<torch::Tensor> altcorr_cuda_forward(torch::Tensor fmap1, torch::Tensor fmap2, torch::Tensor coords, int radius);
PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("corr_index_forward", &corr_index_cuda_forward);
  m.def("altcorr_forward", &altcorr_cuda_forward);
}
