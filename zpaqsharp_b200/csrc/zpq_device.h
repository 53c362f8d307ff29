// zpq_device.h -- kernel parameter blocks and the host-callable launch wrappers implemented in
// zpq_kernels.cu.  Internal to libzpaqb200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "zpq_plan.h"

namespace zpq {

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

// Post-processing pass behind the speculative decoder: PostProcessor.write over every job's raw model stream.
cudaError_t launch_post(const PostParams& q, cudaStream_t s);
// Native post-processors (zpq_post.cu): classify every job's stored program, run PASS / LZ77 / BWT / E8E9 natively; run before
// launch_post, which then interprets only what is left (jobkind == PK_GENERIC).
cudaError_t launch_post_native(const PostParams& q, cudaStream_t s);

// ---- pre-processing (zpq_preproc.cu) ----
size_t sa_workspace_bytes(uint64_t n_total);
// Suffix arrays (block-local positions) and optionally inverse suffix arrays of every block of a
// batch; d_off = nb+1 offsets (only differences to d_off[0] matter), `in` = first byte of block 0.
cudaError_t build_suffix_arrays(const uint8_t* in, const uint64_t* d_off, uint32_t nb, uint64_t n_total, uint64_t max_len,
                                uint32_t* sa, uint32_t* isa, void* workspace, size_t workspace_bytes, cudaStream_t s,
                                int* rounds_out);
cudaError_t launch_bwt_emit(const uint8_t* in, const uint64_t* d_off, const uint32_t* sa, const uint64_t* d_out_off, uint8_t* out,
                            EncJob* jobs, uint32_t nb, uint64_t max_len, cudaStream_t s);

struct LzParams {            // LZBuffer constructor arguments, LZBuffer.cs:151-222
  const uint8_t* in;         // first byte of block 0 of this launch
  const uint64_t* off;       // nb+1 input offsets
  uint8_t* out;              // transformed streams
  const uint64_t* out_off;   // nb+1 output slot offsets
  EncJob* jobs;              // jobs[b].in_len receives the stream length (0xFFFFFFFF = slot overflow)
  uint32_t* ht;              // hash tables, nb * htsize entries (hash matcher)
  const uint32_t* sa;        // suffix arrays / inverse suffix arrays (suffix-array matcher)
  const uint32_t* isa;
  uint32_t nb, htsize;
  int level, use_sa, checkbits;
  uint32_t minMatch, minMatch2, maxMatch, maxLiteral, lookahead, bucket, shift1, shift2, minMatchBoth, rb;
};
cudaError_t launch_lz77(const LzParams& p, cudaStream_t s);
// Byte-gap histograms (4096 bins per block) for the data analysis of method levels 5..9 (LibZPAQ.cs:242-258).
cudaError_t launch_gap_hist(const uint8_t* in, const uint64_t* d_off, uint32_t nb, uint64_t max_len, int* gap, cudaStream_t s);

}  // namespace zpq
