// TEST INFRASTRUCTURE ONLY.  A binding of our own for the reference's matching kernels so that
// VSLAM/backend/src/matching_kernels.cu can be compiled where it lies WITHOUT gn.cpp / gn_kernels.cu (those need Eigen, which
// is not vendored).  Declares the two host functions that file defines (matching_kernels.cu:86-116, 280-316) and exports them
// under the reference's Python names.  Built by oracle/build_ref.py into oracle/_ref/mast3r_matching_ref.so.
#include <torch/extension.h>
#include <vector>

std::vector<torch::Tensor> refine_matches_cuda(torch::Tensor D11, torch::Tensor D21, torch::Tensor p1, const int radius,
                                               const int dilation);
std::vector<torch::Tensor> iter_proj_cuda(torch::Tensor rays_img_with_grad, torch::Tensor pts_3d_norm, torch::Tensor p_init,
                                          const int max_iter, const float lambda_init, const float cost_thresh);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("iter_proj", &iter_proj_cuda, "reference iter_proj (matching_kernels.cu)");
    m.def("refine_matches", &refine_matches_cuda, "reference refine_matches (matching_kernels.cu)");
}
