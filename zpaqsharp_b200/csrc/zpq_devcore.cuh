// zpq_devcore.cuh -- device core of the ZPAQ block codec for sm_100a.  NVRTC-safe: this header
// (plus zpq_plan.h) is embedded in the library and compiled at run time together with a generated
// model description, and it is also included by the ahead-of-time kernels.
//
// Execution model: ZPAQ archive blocks are independent (LICENSE:44-46), so every block is owned
// by ONE WARP for its whole life.  A coding kernel is launched once per batch with as many
// resident warps as per-block state arenas fit in HBM; each warp pulls block indices from an
// atomic queue, (re)initialises its arena and codes the block bit by bit.
//
// Lane-resident predictor (models with <= 32 components): lane i OWNS component i -- descriptor,
// table pointers, context and current stretched prediction live in that lane's registers.
//   phase A   table-dependent part of every component, all lanes at once
//   levels    inputs travel by SHFL between lanes, dependency level by level; a MIX is evaluated
//             by the whole warp (lane j: weight j x input j, REDUX adds)      [Predictor.cs:245-350]
//   update    every lane trains its own component; MIX rows one weight per lane [Predictor.cs:353-475]
//   nibble    ICM/ISSE hash rows (16 B) are cached in shared memory for the 4 bits they serve;
//             all lanes look up their next rows together so the DRAM misses overlap [Predictor.cs:550-567]
//   coder     32-bit arithmetic coder redundantly in registers of all lanes  [Encoder.cs:87-103, Decoder.cs:136-158]
//   HCOMP     per byte: interpreted by lane 0 (generic) or compiled to straight-line code and run
//             uniformly by all lanes (generated models)                       [ZPAQL.cs:1028-1265]
// The "Model" policy class says how phases are composed: GenericModel walks the plan at run time;
// generated models (zpq_codegen.cpp) unroll the component structure and HCOMP at compile time.
//
// All arithmetic is 32-bit integer; results are bit-identical to the reference semantics.
#pragma once
#include "zpq_plan.h"

namespace zpq {

#define ZPQ_FULL 0xFFFFFFFFu
constexpr int kCtaThreads = 512;
enum { BLK_OK = 0, BLK_OVERFLOW = 1, BLK_CORRUPT = 2, BLK_ZPAQL = 3, BLK_POSTPROC = 4 };  // == ZPQ_BLOCK_*

// ------------------------------------------------------------------------------------------
// Shared-memory resident read-only state of a CTA
// ------------------------------------------------------------------------------------------
struct Shared {
  const int16_t* stretch;
  const uint16_t* squash;
  const int32_t* dt;
  const uint16_t* dt2k;
  const uint8_t* ns;
  const CompDesc* comp;
  const uint8_t* order;
  const Step* steps;
  const uint8_t* hcomp;  // shared copy when it fits, else the plan's global copy
  const MixDesc* mix;
  int n, nsteps, hcomp_len, nmix, maxlevel;
};

// Per-block (per-warp) bindings.
struct Blk {
  uint8_t* arena;
  uint8_t* slice;
  int32_t* p;        // step-scheduled kernels: stretched predictions (shared)
  uint32_t* st;      // step-scheduled kernels: 5 words per component (shared)
  uint32_t* H; uint32_t hmask;
  uint8_t* M; uint32_t mmask;
  uint32_t* R;
  int c8, hmap4;
  uint32_t status;
};

struct VM { uint32_t b, c, d, f; };   // ZPAQL registers that persist between runs (A is the input)

struct VMEnv {
  const uint8_t* code; int len;             // program incl. END byte; pc is relative to code
  uint32_t* H; uint32_t hmask;
  uint8_t* M; uint32_t mmask;
  uint32_t* R;
  uint8_t* out; uint64_t out_pos, out_cap;  // OUT destination (decode post-processing only)
};

__device__ __forceinline__ int clamp2k(int x) { return max(-2048, min(2047, x)); }
__device__ __forceinline__ int clamp512k(int x) { return max(-(1 << 19), min((1 << 19) - 1, x)); }

// ------------------------------------------------------------------------------------------
// ZPAQL interpreter (ZPAQL.cs:1028-1265), decoded by instruction field.  Runs on one lane.
// Returns 0 on HALT, 1 on an execution error, -1 when the instruction budget is spent.
// ------------------------------------------------------------------------------------------
static __device__ __noinline__ int zpaql_run(VM& vm, VMEnv& e, uint32_t input, uint64_t budget) {
  uint32_t a = input, b = vm.b, c = vm.c, d = vm.d, f = vm.f;
  int pc = 0, rc = -1;
  const uint8_t* code = e.code;
  const int len = e.len;
#define MB e.M[b & e.mmask]
#define MC e.M[c & e.mmask]
#define HD e.H[d & e.hmask]
  while (budget--) {
    if ((unsigned)pc >= (unsigned)len) { rc = 1; break; }
    const int op = code[pc++];
    if (op < 64) {
      const int ddd = op >> 3, x = op & 7;
      if (ddd == 7) {
        if (x == 0) { rc = 0; break; }                                   // HALT
        else if (x == 1) {                                               // OUT
          if (e.out) { if (e.out_pos < e.out_cap) e.out[e.out_pos] = (uint8_t)a; ++e.out_pos; }
        } else if (x == 3) a = (a + MB + 512) * 773;                      // HASH
        else if (x == 4) HD = (HD + a + 512) * 773;                       // HASHD
        else if (x == 7) pc += ((code[pc] + 128) & 255) - 127;            // JMP
        else { rc = 1; break; }
        continue;
      }
      if (x == 7) {
        const int n = code[pc];
        if (ddd < 4) { const uint32_t v = e.R[n]; ++pc; if (ddd == 0) a = v; else if (ddd == 1) b = v; else if (ddd == 2) c = v; else d = v; }
        else if (ddd == 4) { if (f) pc += ((n + 128) & 255) - 127; else ++pc; }   // JT
        else if (ddd == 5) { if (!f) pc += ((n + 128) & 255) - 127; else ++pc; }  // JF
        else { e.R[n] = a; ++pc; }                                                 // R=A
        continue;
      }
      if (x > 4 || op == 0) { rc = 1; break; }
      uint32_t v;
      switch (ddd) {
        case 0: v = a; break; case 1: v = b; break; case 2: v = c; break; case 3: v = d; break;
        case 4: v = MB; break; case 5: v = MC; break; default: v = HD; break;
      }
      uint32_t w;
      if (x == 0) { w = a; a = (ddd == 4 || ddd == 5) ? ((a & ~255u) | v) : v; }  // swap (low byte only for M)
      else if (x == 1) w = v + 1;
      else if (x == 2) w = v - 1;
      else if (x == 3) w = ~v;
      else w = 0;
      switch (ddd) {
        case 0: if (x) a = w; break;
        case 1: b = w; break; case 2: c = w; break; case 3: d = w; break;
        case 4: MB = (uint8_t)w; break; case 5: MC = (uint8_t)w; break; default: HD = w; break;
      }
      continue;
    }
    if (op == 255) {                                                      // LJ
      pc = code[pc] + 256 * code[pc + 1];
      if (pc >= len) { rc = 1; break; }
      continue;
    }
    const int s = op & 7;
    uint32_t v;
    switch (s) {
      case 0: v = a; break; case 1: v = b; break; case 2: v = c; break; case 3: v = d; break;
      case 4: v = MB; break; case 5: v = MC; break; case 6: v = HD; break;
      default: v = code[pc++]; break;
    }
    if (op < 128) {                                                       // assignment
      const int ddd = (op >> 3) & 7;
      if (ddd == 7) { rc = 1; break; }
      switch (ddd) {
        case 0: a = v; break; case 1: b = v; break; case 2: c = v; break; case 3: d = v; break;
        case 4: MB = (uint8_t)v; break; case 5: MC = (uint8_t)v; break; default: HD = v; break;
      }
      continue;
    }
    const int x = (op >> 3) & 15;
    if (x > 13) { rc = 1; break; }
    switch (x) {
      case 0: a += v; break;
      case 1: a -= v; break;
      case 2: a *= v; break;
      case 3: a = v ? a / v : 0; break;
      case 4: a = v ? a % v : 0; break;
      case 5: a &= v; break;
      case 6: a &= ~v; break;
      case 7: a |= v; break;
      case 8: a ^= v; break;
      case 9: a <<= (v & 31); break;
      case 10: a >>= (v & 31); break;
      case 11: f = (a == v); break;
      case 12: f = (a < v); break;
      default: f = (a > v); break;
    }
  }
#undef MB
#undef MC
#undef HD
  vm.b = b; vm.c = c; vm.d = d; vm.f = f;
  return rc;
}

// ------------------------------------------------------------------------------------------
// Table initialisation by the owning warp (Predictor.cs:96-165)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void init_block_state(const Plan* plan, const Tables* tab, uint8_t* arena, uint8_t* slice, int lane) {
  const int nops = plan->ninit;
  for (int k = 0; k < nops; ++k) {
    const InitOp op = plan->init[k];
    uint8_t* dst = op.to_smem ? slice + op.dst : arena + op.dst;
    if (op.kind == 0) {
      const uint4 v = make_uint4(op.value, op.value, op.value, op.value);
      uint4* q = reinterpret_cast<uint4*>(dst);
      const uint64_t n16 = op.bytes >> 4;
      uint64_t i = lane;
      for (; i + 96 < n16; i += 128) { q[i] = v; q[i + 32] = v; q[i + 64] = v; q[i + 96] = v; }
      for (; i < n16; i += 32) q[i] = v;
    } else {
      uint32_t* q = reinterpret_cast<uint32_t*>(dst);
      const uint64_t nw = op.bytes >> 2;
      if (op.kind == 1) { for (uint64_t i = lane; i < nw; i += 32) q[i] = (i & 1) ? 0u : tab->icm_init[(i >> 1) & 255]; }
      else if (op.kind == 2) { for (uint64_t i = lane; i < nw; i += 32) q[i] = tab->isse_init[i & 511]; }
      else if (op.kind == 4) { for (uint64_t i = lane; i < nw; i += 32) q[i] = tab->icm_init[i & 255]; }
      else {
        const uint32_t w = tab->sse_init[lane] | op.value;  // period 32 == warp width
        for (uint64_t i = lane; i < nw; i += 32) q[i] = w;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// CTA prologue: stage tables and model descriptors into shared memory.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_shared(const CodecParams& P, uint8_t* smem, Shared& S) {
  const Plan* plan = P.plan;
  const SmemLayout& L = P.sm;
  {  // stretch, squash, dt2k, ns (and dt when the model trains a CM / SSE) are the first bytes of Tables, in this order
    const uint4* src = reinterpret_cast<const uint4*>(P.tab);
    uint4* dst = reinterpret_cast<uint4*>(smem + L.stretch);
    for (int i = threadIdx.x; i < (int)(L.tables_bytes / 16); i += blockDim.x) dst[i] = src[i];
  }
  const int n = plan->n, ns = plan->nsteps;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(plan->comp);
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem + L.comp);
    for (int i = threadIdx.x; i < n * (int)(sizeof(CompDesc) / 4); i += blockDim.x) dst[i] = src[i];
    for (int i = threadIdx.x; i < n; i += blockDim.x) smem[L.order + i] = plan->order[i];
    const uint32_t* ssrc = reinterpret_cast<const uint32_t*>(plan->steps);
    uint32_t* sdst = reinterpret_cast<uint32_t*>(smem + L.steps);
    for (int i = threadIdx.x; i < ns; i += blockDim.x) sdst[i] = ssrc[i];
    if (L.hcomp != kNoSmem)
      for (int i = threadIdx.x; i < plan->hcomp_len + 8; i += blockDim.x) smem[L.hcomp + i] = plan->hcomp[i];
    const uint32_t* msrc = reinterpret_cast<const uint32_t*>(plan->mix);
    uint32_t* mdst = reinterpret_cast<uint32_t*>(smem + L.mix);
    for (int i = threadIdx.x; i < plan->nmix * (int)(sizeof(MixDesc) / 4); i += blockDim.x) mdst[i] = msrc[i];
  }
  __syncthreads();
  S.stretch = reinterpret_cast<const int16_t*>(smem + L.stretch);
  S.squash = reinterpret_cast<const uint16_t*>(smem + L.squash);
  S.dt = reinterpret_cast<const int32_t*>(smem + L.dt);
  S.dt2k = reinterpret_cast<const uint16_t*>(smem + L.dt2k);
  S.ns = smem + L.ns;
  S.comp = reinterpret_cast<const CompDesc*>(smem + L.comp);
  S.order = smem + L.order;
  S.steps = reinterpret_cast<const Step*>(smem + L.steps);
  S.hcomp = L.hcomp != kNoSmem ? smem + L.hcomp : plan->hcomp;
  S.mix = reinterpret_cast<const MixDesc*>(smem + L.mix);
  S.n = n; S.nsteps = ns; S.hcomp_len = plan->hcomp_len; S.nmix = plan->nmix; S.maxlevel = plan->maxlevel;
}

__device__ __forceinline__ void bind_block(const CodecParams& P, uint8_t* smem, Blk& w, uint32_t gw, int warp) {
  const Plan* plan = P.plan;
  w.arena = P.arenas + (uint64_t)gw * P.arena_stride;
  w.slice = smem + P.sm.slices + (uint32_t)warp * P.sm.slice_bytes;
  w.p = reinterpret_cast<int32_t*>(w.slice + plan->smem_p);
  w.st = reinterpret_cast<uint32_t*>(w.slice + plan->smem_st);
  w.H = plan->smem_h != kNoSmem ? reinterpret_cast<uint32_t*>(w.slice + plan->smem_h)
                                : reinterpret_cast<uint32_t*>(w.arena + plan->off_h);
  w.hmask = (1u << plan->hh) - 1;
  w.M = plan->smem_m != kNoSmem ? w.slice + plan->smem_m : w.arena + plan->off_m;
  w.mmask = (uint32_t)((1ull << plan->hm) - 1);
  w.R = reinterpret_cast<uint32_t*>(w.arena + plan->off_r);
}

// restored train(), Predictor.cs:1031-1036
__device__ __forceinline__ void train(const Shared& S, uint32_t* cm, uint32_t limit, int y) {
  uint32_t pn = *cm;
  const uint32_t count = pn & 0x3ff;
  const int err = y * 32767 - (int)(pn >> 17);
  pn += ((uint32_t)err * (uint32_t)S.dt[count] & 0xFFFFFC00u) + (count < limit);
  *cm = pn;
}

// ==========================================================================================
// Lane-resident predictor
// ==========================================================================================
struct LaneRegs {
  int type, level, srcj, srck;
  int d;              // pipelined encoder: bits this component works behind the lead (zpq_pipe.cuh)
  uint32_t a1, a2, a3, a4, a5;
  uint32_t mask, mask2;
  uint8_t* tab;
  uint8_t* tab2;
  uint32_t* cm;       // ICM/ISSE probability / weight map (shared or arena)
  uint8_t* row;       // this lane's 16-byte row cache in shared memory
  // dynamic
  uint32_t cxt, c, ma, mb, mpos, h;
  uint32_t mbyte;     // MATCH: the byte the match predicts next
  uint32_t mbit;      // MATCH: bits of the current byte already coded
  int p, t0, t1;
  int2* chain;        // this lane's slot {w0, w1*64} for ISSE chains evaluated by the whole warp
  // MIX k held in registers: weight `lane` of the current row and of both possible next rows
  int mw[kMixRegs], mn0[kMixRegs], mn1[kMixRegs], mn2[kMixRegs];
  uint32_t moff[kMixRegs], mo0[kMixRegs], mo1[kMixRegs], mo2[kMixRegs];   // byte offsets of those weights in the table
  const uint8_t* mixtab[kMixRegs];
  // two-role encoder: rows requested kDuoMixAhead bits ahead (mq[0] = the current bit's) and the last trained weights
  int mq[kMixRegs][kDuoMixAhead], mt[kMixRegs][kDuoMixAhead];
  uint32_t mqo[kMixRegs][kDuoMixAhead], mto[kMixRegs][kDuoMixAhead];
};

struct WarpCtx {
  uint8_t* arena;
  uint32_t* H; uint32_t hmask;
  int c8, hmap4;
  uint32_t mixh[kMixRegs];   // context hash of MIX k for the current byte
};

__device__ __forceinline__ void lane_load(const Shared& S, const CodecParams& P, const Blk& w, LaneRegs& r, int lane) {
  r.type = C_NONE; r.level = 0; r.srcj = r.srck = 0; r.d = 0;
  r.a1 = r.a2 = r.a3 = r.a4 = r.a5 = 0; r.mask = r.mask2 = 0;
  r.tab = r.tab2 = nullptr;
  r.cm = reinterpret_cast<uint32_t*>(const_cast<int16_t*>(S.stretch));   // readable dummy for lanes without a map
  r.row = w.slice + P.plan->smem_rows + lane * 16;
  r.chain = reinterpret_cast<int2*>(w.slice + P.plan->smem_chain) + lane;
  for (int k = 0; k < kMixRegs; ++k) r.mixtab[k] = k < S.nmix ? w.arena + S.mix[k].tab : w.arena;
  if (lane < S.n) {
    const CompDesc& d = S.comp[lane];
    r.type = d.type; r.level = d.level; r.d = d.delay;
    r.a1 = d.a[0]; r.a2 = d.a[1]; r.a3 = d.a[2]; r.a4 = d.a[3]; r.a5 = d.a[4];
    r.mask = d.mask; r.mask2 = d.mask2;
    r.tab = w.arena + d.tab; r.tab2 = w.arena + d.tab2;
    if (d.type == C_ICM || d.type == C_ISSE)
      r.cm = d.smem_cm != kNoSmem ? reinterpret_cast<uint32_t*>(w.slice + d.smem_cm) : reinterpret_cast<uint32_t*>(w.arena + d.tab2);
    switch (d.type) {
      case C_ISSE: case C_SSE: r.srcj = d.a[1]; break;
      case C_AVG: r.srcj = d.a[0]; r.srck = d.a[1]; break;
      case C_MIX2: r.srcj = d.a[1]; r.srck = d.a[2]; break;
      default: break;
    }
  }
}

// Look the hash row for context `cxt` up (Predictor.cs:550-567) and bring it into the lane's
// shared row cache.  The three candidate rows share one 64-byte line, so their loads overlap.
__device__ __forceinline__ void lane_find(LaneRegs& r, uint32_t cxt) {
  const int sizebits = r.a1 + 2;
  const uint32_t chk = (cxt >> sizebits) & 255;
  const uint32_t h0 = (cxt * 16) & r.mask, h1 = h0 ^ 16, h2 = h0 ^ 32;
  const uint4 r0 = *reinterpret_cast<const uint4*>(r.tab + h0);
  const uint4 r1 = *reinterpret_cast<const uint4*>(r.tab + h1);
  const uint4 r2 = *reinterpret_cast<const uint4*>(r.tab + h2);
  uint4 v; uint32_t at;
  if ((r0.x & 255) == chk) { v = r0; at = h0; }
  else if ((r1.x & 255) == chk) { v = r1; at = h1; }
  else if ((r2.x & 255) == chk) { v = r2; at = h2; }
  else {
    const uint32_t p0 = (r0.x >> 8) & 255, p1 = (r1.x >> 8) & 255, p2 = (r2.x >> 8) & 255;
    at = (p0 <= p1 && p0 <= p2) ? h0 : (p1 < p2 ? h1 : h2);
    v = make_uint4(chk, 0, 0, 0);
  }
  r.c = at;
  *reinterpret_cast<uint4*>(r.row) = v;
}

// ---- phase A: the table-dependent part, per component type (Predictor.cs:263-340) ----------
__device__ __forceinline__ void pa_cm(const Shared& S, const WarpCtx& W, LaneRegs& r) {
  r.cxt = (r.h ^ W.hmap4) & r.mask;
  r.p = S.stretch[reinterpret_cast<const uint32_t*>(r.tab)[r.cxt] >> 17];
}
__device__ __forceinline__ void pa_icm(const Shared& S, const WarpCtx& W, LaneRegs& r) {
  r.cxt = r.row[W.hmap4 & 15];
  r.t0 = (int)r.cm[r.cxt * 2];          // ICM maps use the 8-byte stride of ISSE maps (second word unused)
  r.p = S.stretch[(uint32_t)r.t0 >> 8];
}
__device__ __forceinline__ void pa_isse(const Shared& S, const WarpCtx& W, LaneRegs& r) {
  r.cxt = r.row[W.hmap4 & 15];
  const int2 wt = *reinterpret_cast<const int2*>(r.cm + r.cxt * 2);
  r.t0 = wt.x; r.t1 = wt.y;
  *r.chain = make_int2(wt.x, wt.y << 6);
}
__device__ __forceinline__ void pa_match(const Shared& S, const WarpCtx& W, LaneRegs& r) {
  if (r.ma == 0) r.p = 0;
  else {
    const uint32_t bit = (r.mbyte >> (7 - r.mbit)) & 1;
    r.c = bit;
    r.p = S.stretch[(S.dt2k[r.ma] * (1 - 2 * (int)bit)) & 32767];
  }
}
__device__ __forceinline__ void pa_mix2(const Shared& S, const WarpCtx& W, LaneRegs& r) {
  r.cxt = (r.h + (W.c8 & r.a5)) & r.mask;
  r.t0 = reinterpret_cast<const uint16_t*>(r.tab)[r.cxt];
}
__device__ __forceinline__ void pa_sse(const Shared& S, const WarpCtx& W, LaneRegs& r) {
  r.t0 = (int)((r.h + W.c8) * 32);
}

// ---- level evaluation: the part that needs other components' predictions ---------------------
__device__ __forceinline__ int ev_isse(const LaneRegs& r, int pj) { return clamp2k((r.t0 * pj + r.t1 * 64) >> 16); }
__device__ __forceinline__ int ev_avg(const LaneRegs& r, int pj, int pk) { return (pj * (int)r.a3 + pk * (256 - (int)r.a3)) >> 8; }
__device__ __forceinline__ int ev_mix2(const LaneRegs& r, int pj, int pk) { return (r.t0 * pj + (65536 - r.t0) * pk) >> 16; }
__device__ __forceinline__ int ev_sse(const Shared& S, LaneRegs& r, int pj) {
  int pq = max(0, min(1983, pj + 992));
  const int wt = pq & 63;
  pq >>= 6;
  const uint32_t cx = (uint32_t)r.t0 + pq;
  const uint32_t* cm = reinterpret_cast<const uint32_t*>(r.tab);
  const int v = S.stretch[((cm[cx & r.mask] >> 10) * (64 - wt) + (cm[(cx + 1) & r.mask] >> 10) * wt) >> 13];
  r.cxt = (cx + (wt >> 5)) & r.mask;
  return v;
}

// ---- update per component type (Predictor.cs:365-459) -----------------------------------------
__device__ __forceinline__ void up_cm(const Shared& S, LaneRegs& r, int y) {
  train(S, reinterpret_cast<uint32_t*>(r.tab) + r.cxt, r.a2 * 4u, y);
}
__device__ __forceinline__ void up_icm(const Shared& S, const WarpCtx& W, LaneRegs& r, int y) {
  r.row[W.hmap4 & 15] = S.ns[r.cxt * 4 + y];
  const uint32_t pn = (uint32_t)r.t0;
  r.cm[r.cxt * 2] = pn + (uint32_t)(((int)(y * 32767 - (pn >> 8))) >> 2);
}
__device__ __forceinline__ void up_isse(const Shared& S, const WarpCtx& W, LaneRegs& r, int y, int pj) {
  const int err = y * 32767 - (int)S.squash[r.p + 2048];
  int2 wt;
  wt.x = clamp512k(r.t0 + ((err * pj + (1 << 12)) >> 13));
  wt.y = clamp512k(r.t1 + ((err + 16) >> 5));
  *reinterpret_cast<int2*>(r.cm + r.cxt * 2) = wt;
  r.row[W.hmap4 & 15] = S.ns[r.cxt * 4 + y];
}
__device__ __forceinline__ void match_byte(const WarpCtx& W, LaneRegs& r, int y);
// The history buffer is written once per byte (position `limit` is never read before its byte is
// complete: offsets are non-zero modulo the buffer size), and the predicted byte is read once.
__device__ __forceinline__ void up_match(const WarpCtx& W, LaneRegs& r, int y) {
  if ((int)r.c != y) r.ma = 0;
  if (++r.mbit == 8) match_byte(W, r, y);
}
__device__ __forceinline__ void match_byte(const WarpCtx& W, LaneRegs& r, int y) {
  {
    uint8_t* buf = r.tab2;
    buf[r.mpos] = (uint8_t)(W.c8 * 2 + y);
    r.mbit = 0;
    r.mpos = (r.mpos + 1) & r.mask2;
    uint32_t* idx = reinterpret_cast<uint32_t*>(r.tab) + (r.h & r.mask);
    if (r.ma == 0) {
      r.mb = r.mpos - *idx;
      if (r.mb & r.mask2)
        while (r.ma < 255 && buf[(r.mpos - r.ma - 1) & r.mask2] == buf[(r.mpos - r.ma - r.mb - 1) & r.mask2]) ++r.ma;
    } else r.ma += r.ma < 255;
    *idx = r.mpos;
    if (r.ma) r.mbyte = buf[(r.mpos - r.mb) & r.mask2];
  }
}
__device__ __forceinline__ void up_mix2(const Shared& S, LaneRegs& r, int y, int pj, int pk) {
  const int err = ((y * 32767 - (int)S.squash[r.p + 2048]) * (int)r.a4) >> 5;
  int wt = r.t0 + ((err * (pj - pk) + (1 << 12)) >> 13);
  wt = max(0, min(65535, wt));
  reinterpret_cast<uint16_t*>(r.tab)[r.cxt] = (uint16_t)wt;
}
__device__ __forceinline__ void up_sse(const Shared& S, LaneRegs& r, int y) {
  train(S, reinterpret_cast<uint32_t*>(r.tab) + r.cxt, r.a4 * 4u, y);
}

// ---- branch-free ICM + ISSE + MATCH (the shape of min/mid/max and most makeConfig models) -------
// Every lane executes the same straight-line code on its own registers; lanes of other types read
// harmless dummies and keep their state through selects, so the three per-type dependency chains
// (row byte -> map -> stretch; dt2k -> stretch) overlap instead of running one branch after another.
__device__ __forceinline__ void pa_unified(const Shared& S, const WarpCtx& W, LaneRegs& r) {
  const bool icm = r.type == C_ICM, isse = r.type == C_ISSE, mat = r.type == C_MATCH;
  const uint32_t bh = r.row[W.hmap4 & 15];
  const int2 wt = *reinterpret_cast<const int2*>(r.cm + bh * 2);
  const uint32_t bit = (r.mbyte >> (7 - r.mbit)) & 1;
  const int sp = S.stretch[((uint32_t)wt.x >> 8) & 32767];
  const int pm = S.stretch[(S.dt2k[r.ma] * (1 - 2 * (int)bit)) & 32767];
  if (icm || isse) { r.cxt = bh; r.t0 = wt.x; r.t1 = wt.y; }
  *r.chain = make_int2(wt.x, wt.y << 6);
  r.p = icm ? sp : r.p;
  if (mat) { r.c = r.ma ? bit : r.c; r.p = r.ma ? pm : 0; }
}
__device__ __forceinline__ void up_unified(const Shared& S, const WarpCtx& W, LaneRegs& r, int y, int pj) {
  const bool icm = r.type == C_ICM, isse = r.type == C_ISSE;
  const int err = y * 32767 - (int)S.squash[r.p + 2048];
  const uint32_t nxt = S.ns[(r.cxt & 255) * 4 + y];
  const uint32_t pn = (uint32_t)r.t0;
  int2 wt;
  wt.x = isse ? clamp512k(r.t0 + ((err * pj + (1 << 12)) >> 13)) : (int)(pn + (uint32_t)(((int)(y * 32767 - (pn >> 8))) >> 2));
  wt.y = clamp512k(r.t1 + ((err + 16) >> 5));
  if (icm || isse) {
    *reinterpret_cast<int2*>(r.cm + r.cxt * 2) = wt;
    r.row[W.hmap4 & 15] = (uint8_t)nxt;
  }
  if (r.type == C_MATCH) {
    if ((int)r.c != y) r.ma = 0;
    if (++r.mbit == 8) match_byte(W, r, y);
  }
}

// ---- MIX, evaluated by the whole warp (Predictor.cs:302-316, 427-439) --------------------------
// run-time described variant: weights are read and written in place each bit
__device__ __forceinline__ void mix_predict_rt(const MixDesc& md, const WarpCtx& W, LaneRegs& r, int lane) {
  const uint32_t hm = __shfl_sync(ZPQ_FULL, r.h, md.lane);
  const uint32_t rowi = ((hm + (W.c8 & md.cmask)) & md.mask) * md.m;
  const int pin = __shfl_sync(ZPQ_FULL, r.p, md.j0 + lane);
  int prod = 0;
  if (lane < md.m) prod = (reinterpret_cast<const int*>(W.arena + md.tab)[rowi + lane] >> 8) * pin;
  const int acc = __reduce_add_sync(ZPQ_FULL, prod);
  if (lane == md.lane) { r.p = clamp2k(acc >> 8); r.cxt = rowi; }
}
__device__ __forceinline__ void mix_update_rt(const Shared& S, const MixDesc& md, const WarpCtx& W, LaneRegs& r, int lane, int y) {
  const int pm = __shfl_sync(ZPQ_FULL, r.p, md.lane);
  const uint32_t rowi = __shfl_sync(ZPQ_FULL, r.cxt, md.lane);
  const int pin = __shfl_sync(ZPQ_FULL, r.p, md.j0 + lane);
  const int err = ((y * 32767 - (int)S.squash[pm + 2048]) * (int)md.rate) >> 4;
  if (lane < md.m) {
    int* wp = reinterpret_cast<int*>(W.arena + md.tab) + rowi + lane;
    *wp = clamp512k(*wp + ((err * pin + (1 << 12)) >> 13));
  }
}

// compile-time described variant with the weights of MIX K held in registers: r.mw[K] is weight
// `lane` of the row of the current partial byte; while the bit is being coded both possible next
// rows are fetched into r.mn0/mn1[K], so the row walk never waits for HBM inside a byte.
// Requires cmask == 255 and at least 256 contexts (rows of c8, 2*c8 and 2*c8+1 are then distinct).
template <int K, int MIXLANE, int J0, int M, int RATE, unsigned MASK>
struct MixCT {
  static __device__ __forceinline__ uint32_t off(const WarpCtx& W, uint32_t c8, int lane) {
    return ((W.mixh[K] + c8) & MASK) * (uint32_t)(M * 4) + (uint32_t)lane * 4u;
  }
  static __device__ __forceinline__ void load_current(const WarpCtx& W, LaneRegs& r, int lane) {
    r.moff[K] = off(W, (uint32_t)W.c8, lane);
    r.mw[K] = lane < M ? *reinterpret_cast<const int*>(r.mixtab[K] + r.moff[K]) : 0;
  }
  static __device__ __forceinline__ void predict(const WarpCtx& W, LaneRegs& r, int lane) {
    if (W.c8 < 128) {
      r.mo0[K] = off(W, (uint32_t)W.c8 * 2, lane);
      r.mo1[K] = off(W, (uint32_t)W.c8 * 2 + 1, lane);
      if (lane < M) {
        r.mn0[K] = *reinterpret_cast<const int*>(r.mixtab[K] + r.mo0[K]);
        r.mn1[K] = *reinterpret_cast<const int*>(r.mixtab[K] + r.mo1[K]);
      }
    }
    const int pin = __shfl_sync(ZPQ_FULL, r.p, J0 + lane);
    const int acc = __reduce_add_sync(ZPQ_FULL, (r.mw[K] >> 8) * pin);
    r.p = lane == MIXLANE ? clamp2k(acc >> 8) : r.p;
  }
  static __device__ __forceinline__ void update(const Shared& S, const WarpCtx& W, LaneRegs& r, int lane, int y) {
    const int pm = __shfl_sync(ZPQ_FULL, r.p, MIXLANE);
    const int pin = __shfl_sync(ZPQ_FULL, r.p, J0 + lane);
    const int err = ((y * 32767 - (int)S.squash[pm + 2048]) * RATE) >> 4;
    if (lane < M) {
      const int w = clamp512k(r.mw[K] + ((err * pin + (1 << 12)) >> 13));
      *reinterpret_cast<int*>(const_cast<uint8_t*>(r.mixtab[K]) + r.moff[K]) = w;
    }
  }
  // after the bit is known and before c8 changes
  static __device__ __forceinline__ void shift(LaneRegs& r, int y) {
    r.mw[K] = y ? r.mn1[K] : r.mn0[K];
    r.moff[K] = y ? r.mo1[K] : r.mo0[K];
  }
};

// ------------------------------------------------------------------------------------------
// Model policy that walks the plan at run time (any model with 1..32 components).
// ------------------------------------------------------------------------------------------
struct GenericModel {
  static __device__ __forceinline__ void phase_a(const Shared& S, const WarpCtx& W, LaneRegs& r, int lane) {
    switch (r.type) {
      case C_CM: pa_cm(S, W, r); break;
      case C_ICM: pa_icm(S, W, r); break;
      case C_ISSE: pa_isse(S, W, r); break;
      case C_MATCH: pa_match(S, W, r); break;
      case C_MIX2: pa_mix2(S, W, r); break;
      case C_SSE: pa_sse(S, W, r); break;
      default: break;
    }
  }
  static __device__ __forceinline__ void levels(const Shared& S, const WarpCtx& W, LaneRegs& r, int lane) {
    for (int L = 1; L <= S.maxlevel; ++L) {
      const int pj = __shfl_sync(ZPQ_FULL, r.p, r.srcj);
      const int pk = __shfl_sync(ZPQ_FULL, r.p, r.srck);
      if (r.level == L) {
        switch (r.type) {
          case C_ISSE: r.p = ev_isse(r, pj); break;
          case C_AVG: r.p = ev_avg(r, pj, pk); break;
          case C_MIX2: r.p = ev_mix2(r, pj, pk); break;
          case C_SSE: r.p = ev_sse(S, r, pj); break;
          default: break;
        }
      }
      for (int k = 0; k < S.nmix; ++k) {
        const MixDesc md = S.mix[k];
        if (md.level == L) mix_predict_rt(md, W, r, lane);
      }
    }
  }
  static __device__ __forceinline__ void update(const Shared& S, const WarpCtx& W, LaneRegs& r, int lane, int y) {
    const int pj = __shfl_sync(ZPQ_FULL, r.p, r.srcj);
    const int pk = __shfl_sync(ZPQ_FULL, r.p, r.srck);
    switch (r.type) {
      case C_CM: up_cm(S, r, y); break;
      case C_ICM: up_icm(S, W, r, y); break;
      case C_ISSE: up_isse(S, W, r, y, pj); break;
      case C_MATCH: up_match(W, r, y); break;
      case C_MIX2: up_mix2(S, r, y, pj, pk); break;
      case C_SSE: up_sse(S, r, y); break;
      default: break;
    }
    for (int k = 0; k < S.nmix; ++k) mix_update_rt(S, S.mix[k], W, r, lane, y);
  }
  // bit y is known, c8 not yet shifted
  static __device__ __forceinline__ void mix_shift(LaneRegs& r, int y) {}
  // a new byte begins (c8 == 1, H refreshed)
  static __device__ __forceinline__ void mix_new_byte(const Shared& S, WarpCtx& W, LaneRegs& r, int lane) {}
  static __device__ __forceinline__ int hcomp(const Shared& S, WarpCtx& W, VM& vm, VMEnv& env, uint32_t c, int lane) {
    int rc = 0;
    __syncwarp();
    if (lane == 0) rc = zpaql_run(vm, env, c, 1u << 22);
    rc = __shfl_sync(ZPQ_FULL, rc, 0);
    __syncwarp();
    return rc;
  }
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Probability (0..32767) that the next bit is 1.
template <class Model>
__device__ __forceinline__ int lane_predict(const Shared& S, const WarpCtx& W, LaneRegs& r, int lane) {
  Model::phase_a(S, W, r, lane);
  Model::levels(S, W, r, lane);
  return S.squash[__shfl_sync(ZPQ_FULL, r.p, S.n - 1) + 2048];
}

// Shift the coded bit into c8 / hmap4; at nibble boundaries write hash rows back and fetch the
// next ones; at byte boundaries run HCOMP first (Predictor.cs:463-474).
template <class Model>
__device__ __forceinline__ uint32_t lane_advance(const Shared& S, WarpCtx& W, LaneRegs& r, VM& vm, VMEnv& env, int lane, int y) {
  uint32_t status = BLK_OK;
  int c8 = W.c8 * 2 + y;
  const bool hashed = (r.type == C_ICM || r.type == C_ISSE);
  if (c8 >= 256) {
    if (hashed) *reinterpret_cast<uint4*>(r.tab + r.c) = *reinterpret_cast<const uint4*>(r.row);
    if (Model::hcomp(S, W, vm, env, (uint32_t)(c8 - 256), lane)) status = BLK_ZPAQL;
    r.h = W.H[lane & W.hmask];
    W.hmap4 = 1;
    W.c8 = 1;
    Model::mix_new_byte(S, W, r, lane);
    if (hashed) lane_find(r, r.h + 16);
    if (r.type == C_MATCH) prefetch_l2(reinterpret_cast<const uint32_t*>(r.tab) + (r.h & r.mask));   // used at the end of this byte
    return status;
  }
  Model::mix_shift(r, y);
  if (c8 >= 16 && c8 < 32) {
    W.hmap4 = (W.hmap4 & 0xf) << 5 | y << 4 | 1;
    if (hashed) {
      *reinterpret_cast<uint4*>(r.tab + r.c) = *reinterpret_cast<const uint4*>(r.row);
      lane_find(r, r.h + 16 * c8);
    }
  } else {
    W.hmap4 = (W.hmap4 & 0x1f0) | (((W.hmap4 & 0xf) * 2 + y) & 0xf);
    if (c8 >= 4 && c8 < 8 && hashed) {
      // two bits of the first nibble are known: pull the four rows the second nibble may use into L2
#pragma unroll
      for (uint32_t q = 0; q < 4; ++q) prefetch_l2(r.tab + (((r.h + 16u * ((uint32_t)c8 * 4u + q)) * 16u) & r.mask));
    }
  }
  W.c8 = c8;
  return status;
}

// Reset everything a new block needs (Predictor.init + ZPAQL.inith).
template <class Model>
__device__ __forceinline__ void lane_begin(const CodecParams& P, const Shared& S, Blk& w, WarpCtx& W, LaneRegs& r, VM& vm,
                                           VMEnv& env, int lane) {
  init_block_state(P.plan, P.tab, w.arena, w.slice, lane);
  __syncwarp();
  r.cxt = r.c = r.ma = r.mb = r.mpos = r.h = 0;
  r.t0 = r.t1 = 0;
  for (int k = 0; k < kMixRegs; ++k) { r.mw[k] = r.mn0[k] = r.mn1[k] = 0; r.moff[k] = r.mo0[k] = r.mo1[k] = 0; W.mixh[k] = 0; }
  r.mbyte = 0; r.mbit = 0;
  r.p = r.type == C_CONS ? ((int)r.a1 - 128) * 4 : 0;
  if (r.type == C_MATCH) r.tab2[0] = 1;                         // Predictor.cs:118
  W.arena = w.arena; W.H = w.H; W.hmask = w.hmask; W.c8 = 1; W.hmap4 = 1;
  vm.b = vm.c = vm.d = vm.f = 0;
  env.code = S.hcomp; env.len = S.hcomp_len;
  env.H = w.H; env.hmask = w.hmask; env.M = w.M; env.mmask = w.mmask; env.R = w.R;
  env.out = nullptr; env.out_pos = 0; env.out_cap = 0;
  __syncwarp();
  Model::mix_new_byte(S, W, r, lane);
  if (r.type == C_ICM || r.type == C_ISSE) lane_find(r, 16);     // h = 0, c8 = 1
  __syncwarp();
}

// ------------------------------------------------------------------------------------------
// Encoder body (Encoder.cs:39-103 as driven by Compressor.cs:156-248)
// ------------------------------------------------------------------------------------------
template <class Model>
__device__ __forceinline__ void encode_lanes_body(const CodecParams& P, uint8_t* smem) {
  Shared S;
  stage_shared(P, smem, S);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp;
  if (gw >= P.resident) return;
  Blk w;
  bind_block(P, smem, w, gw, warp);
  LaneRegs r;
  lane_load(S, P, w, r, lane);
  WarpCtx W;
  VM vm; VMEnv env;

  for (;;) {
    uint32_t job = 0;
    if (lane == 0) job = atomicAdd(P.queue, 1u);
    job = __shfl_sync(ZPQ_FULL, job, 0);
    if (job >= P.njobs) break;
    const EncJob J = P.ejobs[job];
    const uint8_t* in = P.in + J.in_off;
    uint8_t* out = P.out + J.out_off;
    const uint64_t total = (uint64_t)J.pre_len + J.in_len;
    uint64_t opos = 0;
    uint32_t status = BLK_OK;
    if (J.in_len == 0xFFFFFFFFu) {   // the pre-processing stage overflowed its slot
      if (lane == 0) { P.results[job].out_len = 0; P.results[job].status = BLK_OVERFLOW; }
      continue;
    }
    lane_begin<Model>(P, S, w, W, r, vm, env, lane);
    uint32_t low = 1, high = 0xFFFFFFFFu;
#define ZPQ_NORMALISE()                                                       \
    while ((high ^ low) < 0x1000000u) {                                       \
      if (lane == 0 && opos < J.out_cap) out[opos] = (uint8_t)(high >> 24);   \
      ++opos;                                                                 \
      high = high << 8 | 255; low <<= 8; low += (low == 0);                   \
    }
    for (uint64_t s = 0; s < total; ++s) {
      const int c = s < J.pre_len ? P.preamble[s] : in[s - J.pre_len];
      ++low;  // encode(0, 0), Encoder.cs:49
      ZPQ_NORMALISE();
      for (int i = 7; i >= 0; --i) {
        const uint32_t pr = (uint32_t)lane_predict<Model>(S, W, r, lane) * 2 + 1;
        const int y = (c >> i) & 1;
        const uint32_t mid = low + (uint32_t)(((uint64_t)(high - low) * pr) >> 16);
        if (y) high = mid; else low = mid + 1;
        ZPQ_NORMALISE();
        Model::update(S, W, r, lane, y);
        status |= lane_advance<Model>(S, W, r, vm, env, lane, y);
      }
      if (opos > J.out_cap || status) break;
    }
    high = low;  // encode(1, 0), Encoder.cs:46
    ZPQ_NORMALISE();
#undef ZPQ_NORMALISE
    if (opos > J.out_cap) status = BLK_OVERFLOW;
    __syncwarp();
    if (lane == 0) { P.results[job].out_len = opos; P.results[job].status = status; }
  }
}

// ------------------------------------------------------------------------------------------
// Decoder body (Decoder.cs:32-158); the post-processor (PostProcessor.cs:37-86) is a separate pass over the raw stream
// ------------------------------------------------------------------------------------------
template <class Model>
__device__ __forceinline__ void decode_lanes_body(const CodecParams& P, uint8_t* smem) {
  Shared S;
  stage_shared(P, smem, S);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp;
  if (gw >= P.resident) return;
  Blk w;
  bind_block(P, smem, w, gw, warp);
  const Plan* plan = P.plan;
  LaneRegs r;
  lane_load(S, P, w, r, lane);
  WarpCtx W;
  VM vm; VMEnv env;

  for (;;) {
    uint32_t job = 0;
    if (lane == 0) job = atomicAdd(P.queue, 1u);
    job = __shfl_sync(ZPQ_FULL, job, 0);
    if (job >= P.njobs) break;
    const DecJob J = P.djobs[job];
    uint8_t* raw = P.out + J.out_off;        // the model's byte stream: PCOMP preamble + transformed data; the post-processor is a separate pass
    uint64_t rpos = 0;
    lane_begin<Model>(P, S, w, W, r, vm, env, lane);
    uint32_t status = BLK_OK;

    uint32_t low = 1, high = 0xFFFFFFFFu, curr = 0;     // Decoder.init: once per block (Decompresser.cs:128-134), not per segment
    for (uint32_t sg = 0; sg < J.seg_count && status == BLK_OK; ++sg) {
      const DecSeg seg = P.segs[J.seg_first + sg];
      const uint8_t* in = P.in + seg.in_off;
      uint64_t ipos = 0;
      auto get = [&]() -> uint32_t {
        if (ipos < seg.in_len) return in[ipos++];
        status = BLK_CORRUPT;
        return 0;
      };
      if (curr == 0)                                      // segment initialisation, Decoder.cs:38-42
        for (int k = 0; k < 4; ++k) curr = curr << 8 | get();
#define ZPQ_DECODE(pr, y)                                                         \
      {                                                                           \
        if (curr < low || curr > high) status = BLK_CORRUPT;                      \
        const uint32_t mid = low + (uint32_t)(((uint64_t)(high - low) * (pr)) >> 16); \
        if (curr <= mid) { y = 1; high = mid; } else { y = 0; low = mid + 1; }     \
        while ((high ^ low) < 0x1000000u) {                                       \
          high = high << 8 | 255; low <<= 8; low += (low == 0);                   \
          curr = curr << 8 | get();                                               \
        }                                                                         \
      }
      while (status == BLK_OK) {
        int eos;
        ZPQ_DECODE(0u, eos);
        if (status) break;
        if (eos) {
          if (curr != 0) status = BLK_CORRUPT;
          break;
        }
        int c = 1;
        while (c < 256) {
          const uint32_t pr = (uint32_t)lane_predict<Model>(S, W, r, lane) * 2 + 1;
          int y;
          ZPQ_DECODE(pr, y);
          c += c + y;
          Model::update(S, W, r, lane, y);
          status |= lane_advance<Model>(S, W, r, vm, env, lane, y);
        }
        if (status) break;
        if (lane == 0 && rpos < J.out_cap) raw[rpos] = (uint8_t)(c - 256);
        ++rpos;
      }
#undef ZPQ_DECODE
      if (lane == 0) P.seg_end[J.seg_first + sg] = rpos;
    }
    if (status == BLK_OK && rpos > J.out_cap) status = BLK_OVERFLOW;
    __syncwarp();
    if (lane == 0) { P.results[job].out_len = rpos; P.results[job].status = status; }
  }
}

}  // namespace zpq
