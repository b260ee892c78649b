// Binding shim (NOT reference code): exposes only the CPU entry point of the
// reference's fill_voxels_cpu.cc (cc/module.cc:18-29 also binds the CUDA one).
#include <pybind11/pybind11.h>
#include <torch/torch.h>
torch::Tensor fill_inside_voxels_cpu(const torch::Tensor grid);
PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("fill_inside_voxels_cpu", &fill_inside_voxels_cpu, "reference CPU fill");
}
