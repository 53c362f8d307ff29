// zpq_aot.h -- registry of specialised lane-resident kernels (internal).
#pragma once
#include <stddef.h>

namespace zpq {

struct SpecKernels {
  const void* enc = nullptr;   // __global__ function pointer (ahead of time) or cudaKernel_t (NVRTC); time-skewed when the model allows
  const void* enc_lanes = nullptr;   // encoder that walks bit by bit (huge blocks, A/B measurements)
  const void* enc_duo = nullptr;     // two-role encoder (zpq_duo.cuh) when the model allows it
  int duo_g = 0;                     // lanes per block in its role warps (8, 16 or 32)
  const void* dec = nullptr;
  const void* dec_fast = nullptr;    // speculative decoder (zpq_fdec.cuh) when the model allows it
  const char* origin = "";     // "aot2 (HCOMP compiled)", "nvrtc", ...
};

// Ahead-of-time kernels register themselves at library load (generated files zpq_gen_aot*.cu).
struct AotRegistrar {
  AotRegistrar(const unsigned char* header, size_t len, const void* enc, const void* enc_lanes, const void* enc_duo, int duo_g, const void* dec,
               const void* dec_fast, const char* origin);
};

}  // namespace zpq
