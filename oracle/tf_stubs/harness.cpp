// C ABI over the reference's tf_interpolate.cpp compiled UNMODIFIED against the stand-in headers next to this file --
// TEST INFRASTRUCTURE (oracle/__init__.py).  Two ways in: the three plain loops directly (threenn_cpu,
// threeinterpolate_cpu, threeinterpolate_grad_cpu -- tf_interpolate.cpp:60,107,131) and the registered OpKernels'
// Compute() (shape checks of :163-168,197-206,231-243 included) via ref_run_kernel.
#include <cstring>

#include "tensorflow/core/framework/op_kernel.h"

void threenn_cpu(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist, int* idx);
void threeinterpolate_cpu(int b, int m, int c, int n, const float* points, const int* idx, const float* weight, float* out);
void threeinterpolate_grad_cpu(int b, int n, int c, int m, const float* grad_out, const int* idx, const float* weight, float* grad_points);

extern "C" {

void ref_threenn_cpu(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist, int* idx) {
  threenn_cpu(b, n, m, xyz1, xyz2, dist, idx);
}
void ref_threeinterpolate_cpu(int b, int m, int c, int n, const float* points, const int* idx, const float* weight, float* out) {
  threeinterpolate_cpu(b, m, c, n, points, idx, weight, out);
}
void ref_threeinterpolate_grad_cpu(int b, int n, int c, int m, const float* grad_out, const int* idx, const float* weight, float* grad_points) {
  threeinterpolate_grad_cpu(b, n, c, m, grad_out, idx, weight, grad_points);
}

// Runs the OpKernel registered under `name` ("ThreeNN", "ThreeInterpolate", "ThreeInterpolateGrad").  Input i is a rank-
// in_rank[i] tensor with dimensions in_shape[4*i ..]; outputs are written to the caller's buffers.  Returns 0, or -1 with the
// kernel's InvalidArgument message in err.
int ref_run_kernel(const char* name, int n_in, const void* const* in_data, const int* in_rank, const long long* in_shape, int n_out,
                   void* const* out_data, char* err, int errlen) {
  auto& reg = tensorflow::kernel_registry();
  auto it = reg.find(name);
  if (it == reg.end()) { strncpy(err, "kernel not registered", (size_t)errlen); return -2; }
  tensorflow::OpKernelContext ctx;
  for (int i = 0; i < n_in; ++i) {
    tensorflow::Tensor t;
    for (int d = 0; d < in_rank[i]; ++d) t.shp.d.push_back(in_shape[4 * i + d]);
    t.data = const_cast<void*>(in_data[i]);
    ctx.inputs.push_back(t);
  }
  for (int i = 0; i < n_out; ++i) ctx.out_buffers.push_back(out_data[i]);
  tensorflow::OpKernel* k = it->second();
  k->Compute(&ctx);
  delete k;
  if (!ctx.status.ok()) {
    strncpy(err, ctx.status.msg.c_str(), (size_t)errlen - 1);
    err[errlen - 1] = 0;
    return -1;
  }
  return 0;
}

}  // extern "C"
