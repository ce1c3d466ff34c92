// K5: the three AST (DeiT-base-distilled) audio encoders -- interface used by amuse_api.cu.
// Reference: models/audio/audio_main_new.py:174-204, models/audio/AST_EVP.py:84-90.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <vector>

namespace amuse {
namespace ast {

struct DevTensor {
  float* p = nullptr;
  std::vector<int64_t> shape;
};

struct Weights {
  std::unordered_map<std::string, DevTensor> raw;   // staged on the device under the reference key
  bool is_ready = false;
  std::string err;
  void* impl = nullptr;   // packed layouts + workspaces (ast_kernels.cu)
};

int stage(Weights& w, const std::string& key, const void* data, const int64_t* shape, int ndim);
bool staged(const Weights& w);
bool ready(const Weights& w);
int finalize(Weights& w, cudaStream_t st);
int forward(Weights& w, int B, const float* fbank, float* con, float* emo, float* sty, cudaStream_t st,
            int64_t* launches);
void release(Weights& w);
const char* last_error(const Weights& w);

}  // namespace ast
}  // namespace amuse
