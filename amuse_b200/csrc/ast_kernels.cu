// K5 placeholder translation unit: staging only; the encoder kernels land in a later milestone.
#include "ast_kernels.cuh"

#include "../../include/amuse_b200.h"

namespace amuse {
namespace ast {

int stage(Weights& w, const std::string& key, const void* data, const int64_t* shape, int ndim) {
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= shape[i];
  DevTensor t;
  t.shape.assign(shape, shape + ndim);
  if (cudaMalloc(&t.p, static_cast<size_t>(n) * sizeof(float)) != cudaSuccess) return AMUSE_E_CUDA;
  if (cudaMemcpy(t.p, data, static_cast<size_t>(n) * sizeof(float), cudaMemcpyDefault) != cudaSuccess) {
    cudaFree(t.p);
    return AMUSE_E_CUDA;
  }
  auto it = w.raw.find(key);
  if (it != w.raw.end()) cudaFree(it->second.p);
  w.raw[key] = t;
  w.is_ready = false;
  return AMUSE_OK;
}
bool staged(const Weights& w) { return !w.raw.empty(); }
bool ready(const Weights& w) { return w.is_ready; }
int finalize(Weights& w, cudaStream_t) {
  w.err = "AST encoders are not implemented yet";
  return AMUSE_E_UNSUPPORTED;
}
int forward(Weights& w, int, const float*, float*, float*, float*, cudaStream_t, int64_t*) {
  w.err = "AST encoders are not implemented yet";
  return AMUSE_E_UNSUPPORTED;
}
void release(Weights& w) {
  for (auto& kv : w.raw) cudaFree(kv.second.p);
  w.raw.clear();
  w.is_ready = false;
}
const char* last_error(const Weights& w) { return w.err.c_str(); }

}  // namespace ast
}  // namespace amuse
