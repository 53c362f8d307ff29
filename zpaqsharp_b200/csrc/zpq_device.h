// zpq_device.h -- kernel parameter blocks and the host-callable launch wrappers implemented in
// zpq_kernels.cu.  Internal to libzpaqb200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "zpq_plan.h"

namespace zpq {

// Offsets (bytes) of the CTA-common part of dynamic shared memory.
struct SmemLayout {
  uint32_t stretch, squash, dt, dt2k, ns, comp, order, steps, mix, hcomp;  // hcomp == kNoSmem: read from the plan
  uint32_t slices;        // first per-block slice
  uint32_t slice_bytes;   // == plan.smem_warp_bytes
  uint32_t total;         // dynamic shared bytes of the launch
};

struct CodecParams {
  const Plan* plan;
  const Tables* tab;
  uint8_t* arenas;            // resident_blocks * arena_stride bytes
  uint64_t arena_stride;
  const uint8_t* in;          // encode: (pre-processed) block bytes ; decode: archive bytes
  const uint8_t* preamble;    // encode: PCOMP preamble bytes coded before the data
  uint8_t* out;               // encode: slot buffer ; decode: restored bytes
  const EncJob* ejobs;
  const DecJob* djobs;
  const DecSeg* segs;
  BlockResult* results;
  uint32_t njobs;
  uint32_t resident;          // warps that take part
  uint32_t* queue;            // next block index (atomic)
  SmemLayout sm;
};

struct LaunchGeom {
  uint32_t grid, warps_per_cta;
  uint32_t lanes;   // 1: lane-resident kernels (n <= 32), 0: step-scheduled generic kernels
};

// Launch wrappers (all asynchronous on `s`).
cudaError_t launch_encode(const CodecParams& p, LaunchGeom g, cudaStream_t s);
cudaError_t launch_decode(const CodecParams& p, LaunchGeom g, cudaStream_t s);
cudaError_t codec_set_smem_limit(uint32_t bytes);

// SHA-1 of nb byte ranges (one thread per range): digests[20*nb].
cudaError_t launch_sha1(const uint8_t* data, const uint64_t* off, const uint32_t* len, uint32_t nb, uint8_t* digests,
                        cudaStream_t s);
// In-place E8E9 forward transform of nb ranges (LibZPAQ.cs:372-384).
cudaError_t launch_e8e9(uint8_t* data, const uint64_t* off, const uint32_t* len, uint32_t nb, cudaStream_t s);

// Frame assembly after encoding: for block i the slot holds, from slot_off[i]:
//   [prefix_len[i] bytes reserved][coded bytes ...]
// finish() copies the prefix (tag, zPQ header, segment header) in front, appends
// 00 00 00 00, (FD sha1 | FE), FF behind, and writes the frame length to frame_len[i].
struct FinishParams {
  uint8_t* slots;
  const uint64_t* slot_off;
  const uint8_t* prefix;        // concatenated prefixes
  const uint32_t* prefix_off;   // nb+1
  const BlockResult* results;
  const uint8_t* digests;       // 20*nb or null
  uint64_t* frame_len;          // nb
  uint32_t nb;
};
cudaError_t launch_finish(const FinishParams& p, cudaStream_t s);
// Exclusive prefix sum of len[0..nb) into off[0..nb] on the device.
cudaError_t launch_scan(const uint64_t* len, uint64_t* off, uint32_t nb, cudaStream_t s);
// out[dst_off[i] .. +len[i]) = src[src_off[i] ..): compaction of frames / restored blocks.
cudaError_t launch_gather(const uint8_t* src, const uint64_t* src_off, const uint64_t* len, uint8_t* dst,
                          const uint64_t* dst_off, uint64_t dst_cap, uint32_t nb, uint64_t max_len, cudaStream_t s);

}  // namespace zpq
