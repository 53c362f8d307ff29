// zpq_plan.h -- structures shared by the host scheduler and the sm_100a kernels.
//
// A "plan" is the device-side form of one ZPAQ block header (ZPAQL.cs:112-156 layout): the
// component list of Predictor.init (Predictor.cs:84-171) flattened into fixed-size descriptors
// with table offsets inside a per-block state arena, an evaluation schedule (components grouped
// by dependency level so that one warp evaluates a level across its lanes), the table
// initialisation list and the HCOMP bytecode.
#pragma once
#ifdef __CUDACC_RTC__
typedef unsigned char uint8_t; typedef signed char int8_t; typedef unsigned short uint16_t; typedef short int16_t;
typedef unsigned int uint32_t; typedef int int32_t; typedef unsigned long long uint64_t; typedef long long int64_t;
#else
#include <stdint.h>
#endif

namespace zpq {

enum CompType : uint8_t { C_NONE = 0, C_CONS, C_CM, C_ICM, C_MATCH, C_AVG, C_MIX2, C_MIX, C_ISSE, C_SSE };

constexpr int kMaxComp = 255;
constexpr uint32_t kNoSmem = 0xFFFFFFFFu;

// One model component.  `a[]` are the descriptor bytes after the type (Component.cs:27-43).
struct CompDesc {
  uint8_t type;
  uint8_t a[5];
  uint8_t level;       // dependency depth: 0 = reads no other prediction
  uint8_t coop;        // 1 = evaluated by the whole warp (MIX)
  uint32_t mask;       // index mask of the primary table (entries - 1), or context-count mask for MIX/MIX2
  uint32_t mask2;      // MATCH: buffer mask
  uint64_t tab;        // arena byte offset of the primary table (cm / ht / a16 / MATCH index)
  uint64_t tab2;       // arena byte offset of the secondary table (ICM/ISSE cm when not in smem; MATCH buffer)
  uint32_t smem_cm;    // byte offset of the ICM/ISSE cm table in the warp's shared slice, or kNoSmem
  uint8_t delay;       // pipelined encoder: this component works `delay` bits behind the leading bit
  uint8_t pad[3];
};
static_assert(sizeof(CompDesc) == 40, "CompDesc layout");

// Table fill executed by the owning warp when it starts a block (Predictor.cs:96-165).
struct InitOp {
  uint64_t dst;        // arena byte offset, or shared-slice offset when to_smem
  uint64_t bytes;      // multiple of 16
  uint32_t value;      // fill word (kind 0) / SSE start count (kind 3)
  uint8_t kind;        // 0 = constant word, 1 = ICM cm template, 2 = ISSE cm template, 3 = SSE pattern, 4 = compact ICM cm (4-byte stride)
  uint8_t to_smem;
  uint8_t role;        // two-role encoder (zpq_duo.cuh): 0 = table of the context role (HCOMP arrays, MATCH), 3 = of the history role, 1 = of the coder role, 2 = of the mixer role
  uint8_t pad;
};

// One step of the per-bit prediction schedule.
struct alignas(4) Step {
  uint16_t first;      // index into order[]
  uint8_t count;       // components in this step (<= 32); lane l takes order[first+l]
  uint8_t coop;        // 1 = a single cooperative component
};

// A MIX component as the lane-resident kernel sees it (evaluated by the whole warp).
struct MixDesc {
  uint8_t lane;        // component index == owning lane
  uint8_t level;
  uint8_t j0, m;       // inputs p[j0 .. j0+m)
  uint8_t rate, cmask; // learning rate, mask applied to the partial byte c8
  uint16_t pad;
  uint32_t mask;       // contexts - 1
  uint32_t pad2;
  uint64_t tab;        // arena offset of the weight rows
};
constexpr int kMaxMix = 16;
constexpr int kMixRegs = 4;   // MIX components whose weights a specialised kernel keeps in registers
constexpr int kMaxPipeDelay = 23;   // the pipelined encoder keeps the last 32 coded bits in one register
constexpr int kPipeMixAhead = 2;    // MIX rows are loaded this many bits before they are used
constexpr int kDuoMixAhead = 6;     // ... in the two-role encoder, whose ticks are shorter than an L2 round trip

struct Plan {
  int32_t n;                    // components
  int32_t hh, hm, ph, pm;       // log2 sizes of H, M (HCOMP) and H, M (PCOMP)
  int32_t nsteps, nupd, ninit;
  int32_t hcomp_len;            // bytes of HCOMP program incl. END
  uint32_t smem_warp_bytes;     // shared memory slice per resident block
  uint32_t smem_p, smem_st, smem_h, smem_cm;  // slice offsets: p[], state[], H (or kNoSmem), cm tables
  uint64_t arena_bytes;         // per resident block
  uint64_t off_h, off_m, off_r; // HCOMP H (if not in smem), M, R[256]
  uint64_t off_ph, off_pm, off_pr, off_pcode;  // decode only: PCOMP H, M, R and program (<= 64 KB)
  int32_t lane_ok;              // 1: n <= 32 and few MIXes -> lane-resident kernel applies
  int32_t nmix, maxlevel;
  uint32_t smem_rows;           // slice offset of the 32 x 16-byte hash-row cache
  uint32_t smem_m;              // slice offset of the HCOMP M array when it is small, else kNoSmem
  uint32_t smem_chain;          // slice offset of 32 x {w0, w1*64} slots for warp-evaluated ISSE chains
  // pipelined encoder (zpq_pipe.cuh): rings indexed [bit time & (ring_slots-1)][component]
  int32_t pipe_ok;              // 1: the time-skewed encoder applies (lane_ok and coder_delay <= kMaxPipeDelay)
  int32_t pipe_maps;            // 1: with this shared-memory budget every ICM/ISSE map is in the shared slice (required at launch)
  int32_t coder_delay;          // delay of the arithmetic coder = delay of component n-1 plus one
  uint32_t ring_slots;          // power of two > coder_delay
  uint32_t ring_stride;         // entries per ring slot (>= n, multiple of 8)
  uint32_t smem_pring;          // int16 stretched predictions
  uint32_t smem_bhring;         // uint8 bit histories of ICM/ISSE components
  uint32_t smem_hsnap;          // uint32 contexts H[i] of the last 8 bytes (ring of 8 x ring_stride)
  // two-role encoder (zpq_duo.cuh): slice laid out for groups of duo_g lanes per block
  int32_t duo_g;                // 0: standard slice layout; 8/16/32: lanes per block in each role warp
  int32_t duo_ok;               // 1: the two-role encoder applies to this model with this slice layout
  int32_t duo_hdepth;           // deepest look-back into a lane's own prediction history (<= 8)
  int32_t duo_ldepth;           // deepest look-back of a lane-owned consumer (ISSE/AVG/MIX2/SSE)
  int32_t duo_split;            // 1: the MIX components run in a role warp of their own (no lane-owned component reads a MIX)
  uint32_t smem_sync;           // DuoSync words of the block
  uint32_t smem_pfring;         // int16 [64]: final stretched prediction per bit, for the arithmetic coder warp
  // speculative decoder (zpq_fdec.cuh)
  int32_t fd_layout;            // 1: the slice is laid out for it (rows, 32 slots of 16 bytes, compact ICM maps)
  uint32_t smem_fd_cm;          // 8 x 64 bytes: the 16-entry line a CM component works on during a nibble, or kNoSmem
  MixDesc mix[kMaxMix];
  CompDesc comp[kMaxComp];
  uint8_t order[kMaxComp];      // components sorted by (level, coop)
  uint8_t upd[kMaxComp];        // lane-parallel update list (non-coop components that learn)
  Step steps[kMaxComp * 2];
  InitOp init[kMaxComp * 2 + 8];
  uint8_t hcomp[65536 + 8];
};

// Per-block job for the coding kernels.
struct EncJob {
  uint64_t in_off;      // into the (pre-processed) input buffer
  uint32_t in_len;
  uint32_t pre_len;     // bytes of the preamble buffer to code first (PCOMP preamble, Compressor.cs:177-188)
  uint64_t out_off;     // into the slot buffer: where coded bytes start
  uint64_t out_cap;     // coded bytes allowed
};

struct DecSeg {         // one segment's coded byte range
  uint64_t in_off;
  uint64_t in_len;
};
struct DecJob {
  uint32_t seg_first, seg_count;  // into the DecSeg array
  uint64_t out_off, out_cap;      // restored bytes (decoders with the post-processor inside) or the raw model stream (zpq_fdec.cuh)
};
// Post-processing pass after the speculative decoder (PostProcessor.cs:37-86): raw model stream -> restored bytes.
struct PostJob {
  uint64_t out_off, out_cap;      // into the slot buffer of restored bytes
};

struct BlockResult {
  uint64_t out_len;     // coded (encode) or restored (decode) bytes
  uint32_t status;      // ZPQ_BLOCK_*
  uint32_t pad;
};

// Read-only tables, built on the host (Predictor.cs:54-67, StateTable.cs) and copied once.
struct Tables {           // the first `SmemLayout::tables_bytes` bytes are staged in shared memory (dt only for models that train a CM / SSE)
  int16_t stretch[32768];
  uint16_t squash[4096];
  uint16_t dt2k[256];
  uint8_t ns[1024];
  int32_t dt[1024];
  uint32_t icm_init[256];   // cminit(j)                       Predictor.cs:111-112
  uint32_t isse_init[512];  // {1<<15, clamp512k(stretch(cminit(j)>>8)*1024)}  Predictor.cs:150-155
  uint32_t sse_init[32];    // squash((j&31)*64-992)<<17       Predictor.cs:163-164
};

// Offsets (bytes) of the CTA-common part of dynamic shared memory.
struct SmemLayout {
  uint32_t stretch, squash, dt, dt2k, ns, comp, order, steps, mix, hcomp;  // hcomp == kNoSmem: read from the plan
  uint32_t tables_bytes;  // bytes of Tables staged: 79360 with dt, 75264 without
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
  uint32_t wb;                // two-role encoder: blocks per CTA
  uint64_t* seg_end;          // speculative decoder: raw stream position at the end of every segment (parallel to segs)
  SmemLayout sm;
};

// What the post-processing pass does with a job (PostParams::jobkind): kind | e8 << 4 | param << 8.
// PK_GENERIC = PostProcessor.write with the stored PCOMP program interpreted (any program); the others are native kernels
// for the PASS type and for the four programs makeConfig emits (LibZPAQ.cs:427-826), recognised by comparing the stored
// program with the host front end's own output for the block's (ph, pm).
enum PostKind : uint32_t { PK_GENERIC = 0, PK_PASS = 1, PK_LZ_BITS = 2, PK_LZ_BYTES = 3, PK_BWT = 4, PK_E8E9 = 5,
                           PK_COMPILED = 6 /* restored by the stored program compiled with NVRTC (zpq_codegen.cpp: generate_post_source) */ };
struct PostCand {             // one recognisable program
  uint32_t kind, e8, param;   // param: LZ_BITS rb (LibZPAQ.cs:427); LZ_BYTES: taken from the program byte at `wild` (minMatch, "$3")
  uint32_t off, len;          // program bytes at cand_bytes + off
  int32_t wild;               // index of the one byte that may differ, or -1
};

struct PostParams {
  const Plan* plan;           // ph, pm and the PCOMP offsets inside an arena
  uint8_t* arenas; uint64_t arena_stride;
  const uint8_t* raw;         // what the decoder wrote: per job at djobs[j].out_off
  const DecJob* djobs;
  const uint64_t* seg_end;
  uint64_t* seg_out_end;      // restored bytes of the block at the end of every segment (parallel to seg_end): the ranges the stored SHA-1s cover
  const BlockResult* raw_results;
  uint8_t* out;               // restored bytes: per job at pjobs[j].out_off
  const PostJob* pjobs;
  BlockResult* results;
  uint32_t njobs, resident;
  uint32_t* queue;            // job counter of the interpreter pass
  uint32_t* queue2;           // job counter of the native BWT pass
  uint32_t* jobkind;          // per job, written by the classifier; a native kernel that meets a stream it does not handle puts PK_GENERIC back
  const uint8_t* cand_bytes; const PostCand* cands; uint32_t ncand;
  uint32_t has_bwt;           // a PK_BWT candidate exists (the arenas then hold 4 << ph bytes at off_ph for the list)
  uint64_t max_raw;           // longest raw stream slot (grid sizing of the PASS copy)
};

struct LaunchGeom {
  uint32_t grid, warps_per_cta;
  uint32_t lanes;   // 1: lane-resident kernels (n <= 32), 0: step-scheduled generic kernels
};

}  // namespace zpq
