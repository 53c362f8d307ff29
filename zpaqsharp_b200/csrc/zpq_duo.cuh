// zpq_duo.cuh -- role-split block encoder for sm_100a.  NVRTC-safe like zpq_devcore.cuh.
//
// The time-skewed encoder of zpq_pipe.cuh still runs every stage of a tick in ONE instruction stream, so a tick costs
// the SUM of the stage latencies (ncu: 1400 cycles per bit, 18 % issue utilisation, ~6 cycles per instruction), and 24
// of the 32 lanes idle for an 8-component model.  HBM caps the blocks in flight (mid.cfg: 11 per SM), so the only way
// to more throughput is a shorter dependent chain per block and bit.  This file cuts the per-bit work of a block into
// ROLES, each a warp of its own with a short chain, and packs several blocks into a warp:
//
//   context role   everything per BYTE that depends on the data only: HCOMP contexts (ZPAQL.cs:1253-1265), the MATCH
//                  component for all 8 bits of the byte at once (Predictor.cs:273-287, 382-411), L2 prefetch of every
//                  table line the next byte will touch; it owns the job queue.
//   history role   hash-row look-ups and bit histories of ICM/ISSE, nibble by nibble (Predictor.cs:550-567, 375-381),
//                  and the complete ICM / CM / CONS components -- their predictions never read another prediction
//                  (Predictor.cs:263-272, 365-373).
//   coder role     what depends on PREDICTIONS and is owned by one lane: ISSE weights, AVG, MIX2, SSE
//                  (Predictor.cs:288-340, 414-455), time-skewed by component delay as in zpq_pipe.cuh; predictions travel
//                  between lanes by SHFL from a per-lane history of the last 8 bits kept in registers, so a tick has no
//                  shared-memory read-after-write.
//   mixer role     every MIX (Predictor.cs:302-316, 427-439), when nothing lane-owned reads a MIX output (DM::SPLIT;
//                  otherwise the coder role evaluates the MIXes too): weight rows requested kDuoMixAhead bits ahead in
//                  a register pipeline, trained rows forwarded when a row repeats.
//   arithmetic coder  ONE warp per CTA, lane b codes block b of the CTA (Encoder.cs:39-103).
//
// The two lead roles hand one int16 per component and bit (a stretched prediction, or the bit history an ISSE will
// index its weights with) to the coder / mixer roles through a 64-bit-deep ring in shared memory, plus the HCOMP
// contexts of the last 8 bytes; the coder role replaces an entry by its own prediction once it has used it (the mixer
// reads that); the last component's prediction goes to a third ring for the arithmetic coder.  Flow control is one
// byte counter per role and block (DuoSync): a producer stays at most 6 bytes ahead of the last reader of its ring.
//
// A role warp serves 32/G blocks at once (G = 8, 16 or 32 lanes per block, the model's component count rounded up):
// lanes [g*G, g*G+G) of every role warp of a group own block g of the group.  Groups run in lockstep while they have
// work; a group that waits (ring full / ring empty / no job) sits the byte out under a branch, every collective uses
// the group's own member mask; a fast path without range tests and with whole-warp collectives runs whenever every
// group is in its steady state.
//
// Five instruction streams share an SM, so code size matters: with every byte loop unrolled the kernel ran 1.5x
// slower on most SMs (instruction fetch); the loops are rolled except the mixer's (see ZPQ_DUO_UNROLL).
//
// The arithmetic is exactly the reference's, per component in the same order, bit for bit; only the interleaving
// between components and blocks changes.  Decoding cannot be split this way.
#pragma once
#include "zpq_pipe.cuh"

namespace zpq {

struct DuoSync {              // per block, in its shared slice
  volatile uint32_t lpos;     // bytes of the current job the history role (hash rows, bit histories) has finished
  volatile uint32_t qpos;     // byte the coder role is working on
  volatile uint32_t jobseq;   // bumped by the lead role when it publishes a job
  volatile uint32_t jobid;    // that job; 0xFFFFFFFF = the queue is empty, retire
  volatile uint32_t cdone;    // jobseq of the last job the arithmetic coder finished
  volatile uint32_t lstatus;  // BLK_* raised by the lead role (ZPAQL error)
  volatile uint32_t qfin;     // bytes (ticks / 8) of the current job the coder role has finished
  volatile uint32_t cpos;     // byte the arithmetic coder is working on
  volatile uint32_t mfin;     // mixer role (when the model has one): bytes finished
  volatile uint32_t mpos;     // ... byte it is working on
  volatile uint32_t lafin;    // bytes of the current job the context role (HCOMP, MATCH) has finished
  uint32_t pad;
};

// ZPQ_DUO_TIMING: each role warp of CTA 0 prints the cycles it spent inside its byte loop bodies and in total
#ifdef ZPQ_DUO_TIMING
#include <stdio.h>
#define ZPQ_T_DECL long long zt_busy = 0, zt_iters = 0, zt_fast = 0; const long long zt_begin = clock64();
#define ZPQ_T_IN const long long zt_in = clock64();
#define ZPQ_T_OUT(isfast) { zt_busy += clock64() - zt_in; ++zt_iters; zt_fast += (isfast) ? 1 : 0; }
#define ZPQ_T_REPORT(name) if (blockIdx.x % 16 == 0 && (threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < 6) printf("cta %d %s warp %d: busy %lld cycles in %lld byte steps (%lld fast) = %.1f per step; total %lld\n", (int)blockIdx.x, name, (int)(threadIdx.x >> 5), zt_busy, zt_iters, zt_fast, zt_iters ? (double)zt_busy / zt_iters : 0.0, clock64() - zt_begin);
#else
#define ZPQ_T_DECL
#define ZPQ_T_IN
#define ZPQ_T_OUT(isfast)
#define ZPQ_T_REPORT(name)
#endif

constexpr uint32_t kDuoRing = 64;     // bits of lead-role output kept per block
constexpr uint32_t kDuoRetire = 0xFFFFFFFFu;

// table initialisation of one role (Predictor.cs:96-165), by all 32 lanes of the calling warp
// Out of line on purpose: inlined into the four role bodies it put 4 x 7.5 KB of cold code between their hot loops and
// the kernel ran 10-20 % slower (instruction cache); it runs once per block and role.
static __device__ __noinline__ void init_block_state_role(const Plan* plan, const Tables* tab, uint8_t* arena, uint8_t* slice, int lane, int role) {
  const int nops = plan->ninit;
  for (int k = 0; k < nops; ++k) {
    const InitOp op = plan->init[k];
    if (op.role != role) continue;
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
        const uint32_t w = tab->sse_init[lane] | op.value;
        for (uint64_t i = lane; i < nw; i += 32) q[i] = w;
      }
    }
  }
}

// Lockstep policy shared by the role warps.  The groups (lanes, in the arithmetic coder warp) of a warp are only
// cheap when they move together, and nothing else keeps them in phase, so a warp that finds some of its running
// groups ready and others not waits a bounded number of polls for the stragglers before it serves a partial set.
// A group that has been left behind `kLagMax` times in a row (table initialisation, a longer job) is not waited for.
constexpr int kLagMax = 3, kSpinMax = 48;
struct Lockstep {
  int lag = 0, spins = 0;
  // true: poll again (some recent straggler is not ready yet)
  __device__ __forceinline__ bool hold(bool running, bool ready) {
    if (__any_sync(ZPQ_FULL, running && !ready && lag < kLagMax) && spins < kSpinMax) { ++spins; return true; }
    spins = 0;
    lag = ready ? 0 : (running ? (lag < 255 ? lag + 1 : lag) : 0);
    return false;
  }
};

// sum over the G lanes of a group; every lane of the group gets the result
template <int G>
__device__ __forceinline__ int grp_sum(uint32_t gmask, int v) {
  if (G == 32) return __reduce_add_sync(ZPQ_FULL, v);
#pragma unroll
  for (int o = 1; o < G; o <<= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}

// the last 8 predictions of a lane, newest in the low 16 bits of h0
struct Hist {
  uint64_t h0, h1;
  __device__ __forceinline__ void push(int p) { h1 = (h1 << 16) | (h0 >> 48); h0 = (h0 << 16) | (uint64_t)(uint32_t)(p & 0xFFFF); }
  __device__ __forceinline__ int at(int idx) const {   // idx 0 = newest
    const uint64_t w = idx < 4 ? h0 : h1;
    return (int)(int16_t)(uint16_t)(w >> ((idx & 3) * 16));
  }
};

// ==========================================================================================
// Lead role
// ==========================================================================================
template <int G>
struct LeadCtx {
  WarpCtx w;                 // c8, hmap4, arena, H of the block
  DuoSync* sync;
  int16_t* lring;            // [bit & 63][G]
  uint32_t* hsnap;           // [byte & 7][G]
  const uint8_t* in; const uint8_t* preamble;
  uint32_t pre_len, total, s, seq, status, job;
  uint32_t cb0, cb1, cb2, hnext;
  int st;                    // 0 idle, 1 running, 2 all bytes done (waiting for the arithmetic coder), 3 retired
  __device__ __forceinline__ uint32_t fetch(uint32_t i) const {
    if (i < pre_len) return preamble[i];
    return i < total ? in[i - pre_len] : 0u;
  }
};

// ------------------------------------------------------------------------------------------
// Context role, one byte: HCOMP, the MATCH component for all 8 bits of the byte, and the L2 prefetches.
// Because the data are known, a MATCH component needs no per-bit work: its predictions for the byte follow from
// the predicted byte, the match length and the position of the first mispredicted bit (Predictor.cs:273-287),
// and its end-of-byte update (Predictor.cs:382-411) only needs the bytes coded so far.
// ------------------------------------------------------------------------------------------
template <class DM>
__device__ __forceinline__ void duo_context_byte(const Shared& S, LeadCtx<DM::G>& C, LaneRegs& r, VM& vm, VMEnv& env, int gl,
                                                  uint32_t gmask, int gbase) {
  constexpr int G = DM::G;
  const uint32_t s = C.s;
  const bool mat = DM::HAS_MATCH && r.type == C_MATCH;
  const uint32_t byte = C.cb0;
  C.cb2 = C.fetch(s + 2);
  r.h = C.hnext;                                              // context of byte s (computed when byte s-1 was here)
  // ---- MATCH, early part: store the byte, request what the end-of-byte update will look at ----
  uint8_t* buf = r.tab2;
  uint32_t* islot = reinterpret_cast<uint32_t*>(r.tab) + (r.h & r.mask);
  uint32_t a[8], b[8], npos = 0, cand = 0;
  if (mat) {
    buf[r.mpos] = (uint8_t)byte;
    npos = (r.mpos + 1) & r.mask2;
    // index slot of this byte's context: requested one byte ago; the previous byte's own update may have hit the same slot
    cand = islot == reinterpret_cast<uint32_t*>(r.tab) + r.cxt ? r.mb : (uint32_t)r.t0;   // r.cxt / r.mb(hi): slot and value written last
#pragma unroll
    for (int k = 0; k < 8; ++k) {       // the 8 most recent bytes and the 8 bytes before the remembered position
      a[k] = buf[(npos - 1 - k) & r.mask2];
      b[k] = buf[(cand - 1 - k) & r.mask2];
    }
  }
  // ---- contexts of byte s+1 (HCOMP sees byte s, ZPAQL.cs:1253-1265) and the lines that byte will touch ----
  if (DM::hcomp(S, C.w, vm, env, byte, gl, gmask)) C.status = BLK_ZPAQL;
  const uint32_t hn = C.w.H[gl & C.w.hmask];
  C.hsnap[((s + 1) & 7) * G + gl] = hn;
  C.hnext = hn;
  if (r.type == C_ICM || r.type == C_ISSE) {
    prefetch_l2(r.tab + (((hn + 16u) * 16u) & r.mask));
    prefetch_l2(r.tab + (((hn + 16u * (16u + (C.cb1 >> 4))) * 16u) & r.mask));
  }
  DM::prefetch(r, hn, C.cb1, gl, gmask, gbase);
  if (mat) {
    // ---- predictions of the 8 bits (Predictor.cs:273-287) ----
    const uint32_t len = r.ma;
    const int d2 = S.dt2k[len & 255u];
    const int sp0 = S.stretch[d2 & 32767], sp1 = S.stretch[(-d2) & 32767];
    const uint32_t diff = (r.mbyte ^ byte) & 255u;
    const int nmatch = diff ? (int)__clz(diff) - 24 : 8;      // bits 0..nmatch-1 are predicted right; bit nmatch (if < 8) ends the match
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int pk = (len && k <= nmatch) ? (((r.mbyte >> (7 - k)) & 1u) ? sp1 : sp0) : 0;
      C.lring[((s * 8u + k) & (kDuoRing - 1)) * G + gl] = (int16_t)pk;
    }
    if (nmatch < 8) r.ma = 0;
    // ---- end of the byte (Predictor.cs:382-411) ----
    if (r.ma == 0) {
      r.c = npos - cand;                                     // r.c: offset of the match (mb)
      if (r.c & r.mask2) {
        uint32_t n = 0;
        bool go = true;
#pragma unroll
        for (int k = 0; k < 8; ++k) { go = go && a[k] == b[k]; n += go ? 1u : 0u; }
        r.ma = n;
        while (n == 8 && r.ma < 255) {                        // longer than 8: keep comparing, 8 byte pairs per round trip
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            a[k] = buf[(npos - r.ma - 1 - k) & r.mask2];
            b[k] = buf[(npos - r.ma - r.c - 1 - k) & r.mask2];
          }
          n = 0; go = true;
#pragma unroll
          for (int k = 0; k < 8; ++k) { go = go && a[k] == b[k] && r.ma + n < 255; n += go ? 1u : 0u; }
          r.ma += n;
        }
      }
    } else r.ma += r.ma < 255;
    *islot = npos;
    r.cxt = r.h & r.mask; r.mb = npos;                         // slot and value just written (forwarded to the next byte if it hits the same slot)
    r.mpos = npos;
    if (r.ma) r.mbyte = buf[(npos - r.c) & r.mask2];
    // index slot of the NEXT byte's context: used one byte from now
    r.t0 = (int)reinterpret_cast<const uint32_t*>(r.tab)[hn & r.mask];
  }
  C.cb0 = C.cb1; C.cb1 = C.cb2;
}

// ------------------------------------------------------------------------------------------
// History role, one nibble: the four bit-history slots a nibble visits are known from the data, so they are read,
// advanced (StateTable.cs next()) and written back together (Predictor.cs:267-272, 375-381, 440-449); the maps indexed
// by these states belong to the coder role.
// ------------------------------------------------------------------------------------------
template <class DM>
__device__ __forceinline__ void duo_history_nibble(const Shared& S, LeadCtx<DM::G>& C, LaneRegs& r, int gl, uint32_t nib, uint32_t bit0,
                                                    uint32_t hmap_hi) {
  constexpr int G = DM::G;
  const bool hashed = r.type == C_ICM || r.type == C_ISSE;
  uint32_t idx[4], bh[4], y[4];
  y[0] = (nib >> 3) & 1; y[1] = (nib >> 2) & 1; y[2] = (nib >> 1) & 1; y[3] = nib & 1;
  idx[0] = 1; idx[1] = 2 + y[0]; idx[2] = 4 + y[0] * 2 + y[1]; idx[3] = 8 + y[0] * 4 + y[1] * 2 + y[2];
  if (DM::HAS_HASHED) {
#pragma unroll
    for (int k = 0; k < 4; ++k) bh[k] = r.row[idx[k]];
    uint32_t nx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) nx[k] = S.ns[bh[k] * 4 + y[k]];
#pragma unroll
    for (int k = 0; k < 4; ++k) if (hashed) r.row[idx[k]] = (uint8_t)nx[k];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int val = 0;
    bool mine = false;
    if (DM::HAS_HASHED) { val = (int)bh[k]; mine = hashed; }     // the bit history the ICM / ISSE map is indexed with
    if (DM::HAS_CM) {
      if (r.type == C_CM) {                                    // Predictor.cs:263-266, 365-373
        C.w.hmap4 = (int)(hmap_hi | idx[k]);
        pa_cm(S, C.w, r);
        val = r.p;
        up_cm(S, r, (int)y[k]);
        mine = true;
      }
    }
    if (DM::HAS_CONS) {
      if (r.type == C_CONS) { val = ((int)r.a1 - 128) * 4; mine = true; }   // Predictor.cs:96-98
    }
    if (mine) C.lring[((bit0 + k) & (kDuoRing - 1)) * G + gl] = (int16_t)val;
  }
}

// History role, one byte: two nibbles, the hash-row look-ups between them (Predictor.cs:550-567).
template <class DM>
__device__ __forceinline__ void duo_history_byte(const Shared& S, LeadCtx<DM::G>& C, LaneRegs& r, FindAhead& F, uint8_t*& row2, int gl) {
  constexpr int G = DM::G;
  const bool hashed = r.type == C_ICM || r.type == C_ISSE;
  const uint32_t s = C.s;
  const uint32_t byte = C.cb0;
  C.cb2 = C.fetch(s + 2);
  r.h = C.hsnap[(s & 7) * G + gl];
  C.hnext = C.hsnap[((s + 1) & 7) * G + gl];                   // written by the context role during its byte s
  if (hashed) {
    if (s == 0) lane_find(r, r.h + 16);
    else find_swap(r, F, row2, r.h + 16);
    find_issue(r, r.h + 16u * (16u + (byte >> 4)), F);         // second nibble of this byte
  }
  // hmap4 (Predictor.cs:463-474): first nibble = the slot index itself; second nibble = 256 + 16 * high nibble + slot index
  duo_history_nibble<DM>(S, C, r, gl, byte >> 4, s * 8u, 0u);
  if (hashed) {
    find_resolve(F, row2);
    find_swap(r, F, row2, r.h + 16u * (16u + (byte >> 4)));
    find_issue(r, C.hnext + 16u, F);                           // first nibble of the next byte
  }
  duo_history_nibble<DM>(S, C, r, gl, byte & 15u, s * 8u + 4u, 256u + ((byte >> 4) << 4));
  if (hashed) find_resolve(F, row2);                           // swapped in when the next byte begins
  C.cb0 = C.cb1; C.cb1 = C.cb2;
}

template <class DM, int LR>
__device__ __forceinline__ void duo_lead_body(const CodecParams& P, uint8_t* smem, const Shared& S, int pair) {
  constexpr int G = DM::G, B = 32 / G;
  const int lane = threadIdx.x & 31, gl = lane & (G - 1), grp = lane / G, gbase = lane & ~(G - 1);
  const uint32_t gmask = G == 32 ? ZPQ_FULL : (((1u << (G & 31)) - 1u) << gbase);
  const Plan* plan = P.plan;
  const uint32_t b = (uint32_t)pair * B + grp;
  const bool valid = b < P.wb && blockIdx.x * P.wb + b < P.resident;
  // groups without a block alias block 0 of the CTA for their (never used) pointers
  const uint32_t bb = valid ? b : 0u;
  Blk w;
  bind_block(P, smem, w, blockIdx.x * P.wb + bb, (int)bb);
  LaneRegs r;
  lane_load(S, P, w, r, gl < S.n ? gl : 0);
  if (gl >= S.n) r.type = C_NONE;
  r.row = w.slice + plan->smem_rows + gl * 16;
  uint8_t* row2 = r.row + 16 * G;
  // maps live in the shared slice (the host launches this kernel only then): keep the pointer provably shared
  r.cm = r.type == C_ICM ? reinterpret_cast<uint32_t*>(w.slice + S.comp[gl < S.n ? gl : 0].smem_cm)
                         : reinterpret_cast<uint32_t*>(const_cast<int16_t*>(S.stretch));
  LeadCtx<G> C;
  C.sync = reinterpret_cast<DuoSync*>(w.slice + plan->smem_sync);
  C.lring = reinterpret_cast<int16_t*>(w.slice + plan->smem_pring);
  C.hsnap = reinterpret_cast<uint32_t*>(w.slice + plan->smem_hsnap);
  C.preamble = P.preamble;
  C.in = P.in; C.pre_len = 0; C.total = 0; C.s = 0; C.seq = 0; C.status = BLK_OK; C.job = 0;
  C.cb0 = C.cb1 = C.cb2 = 0; C.hnext = 0;
  C.st = valid ? 0 : 3;
  C.w.arena = w.arena; C.w.H = w.H; C.w.hmask = w.hmask; C.w.c8 = 1; C.w.hmap4 = 1;
  VM vm; VMEnv env;
  vm.b = vm.c = vm.d = vm.f = 0;
  env.code = S.hcomp; env.len = S.hcomp_len;
  env.H = w.H; env.hmask = w.hmask; env.M = w.M; env.mmask = w.mmask; env.R = w.R;
  env.out = nullptr; env.out_pos = 0; env.out_cap = 0;
  FindAhead F;
  F.r0 = F.r1 = F.r2 = make_uint4(0, 0, 0, 0); F.h0 = F.chk = F.at = 0; F.hz = true;
  if (LR == 0 && valid && gl == 0) {
    C.sync->lpos = 0; C.sync->qpos = 0; C.sync->jobid = 0; C.sync->cdone = 0; C.sync->lstatus = 0; C.sync->jobseq = 0;
    C.sync->qfin = 0; C.sync->cpos = 0; C.sync->mfin = 0; C.sync->mpos = 0; C.sync->lafin = 0;
  }
  __syncthreads();   // the other roles read the sync words from here on

  constexpr uint32_t DB = ((uint32_t)DM::D + 7u) / 8u;   // bytes the coder / mixer role reads back behind its own byte
  Lockstep LS;
  ZPQ_T_DECL
  for (;;) {
    // ---- job management (warp-convergent) ----
    if (LR == 0) {
      // the context role owns the queue: a block is handed out when the arithmetic coder has finished the last one
      if (C.st == 2 && C.sync->cdone == C.seq) C.st = 0;
      uint32_t want = __ballot_sync(ZPQ_FULL, C.st == 0 && gl == 0);
      while (want) {
        const int src = __ffs(want) - 1;
        want &= want - 1;
        uint32_t job = 0;
        if (lane == src) job = atomicAdd(P.queue, 1u);
        job = __shfl_sync(ZPQ_FULL, job, src);
        const bool mine = gbase == src;
        if (job >= P.njobs) {
          if (mine) {
            C.st = 3;
            if (gl == 0) { C.sync->jobid = kDuoRetire; __threadfence_block(); C.sync->jobseq = C.seq + 1; }
          }
          continue;
        }
        const EncJob J = P.ejobs[job];
        uint8_t* arena_g = reinterpret_cast<uint8_t*>(__shfl_sync(ZPQ_FULL, (unsigned long long)w.arena, src));
        uint8_t* slice_g = reinterpret_cast<uint8_t*>(__shfl_sync(ZPQ_FULL, (unsigned long long)w.slice, src));
        if (J.in_len != 0xFFFFFFFFu) init_block_state_role(plan, P.tab, arena_g, slice_g, lane, 0);   // ZPAQL.inith, MATCH tables
        __syncwarp();
        if (mine) {
          C.in = P.in + J.in_off; C.pre_len = J.pre_len;
          C.total = J.in_len == 0xFFFFFFFFu ? 0u : J.pre_len + J.in_len;
          C.s = 0; C.status = BLK_OK; C.seq += 1;
          r.cxt = r.c = r.ma = r.mb = r.mpos = r.h = 0; r.t0 = r.t1 = 0; r.p = 0; r.mbyte = 0; r.mbit = 0;
          if (r.type == C_MATCH) r.tab2[0] = 1;                  // Predictor.cs:118
          C.w.c8 = 1; C.w.hmap4 = 1;
          vm.b = vm.c = vm.d = vm.f = 0;
          if (gl < G) C.hsnap[gl] = 0;                           // contexts of byte 0 are H == 0
          C.hnext = 0;
          r.cxt = 0xFFFFFFFFu;                                    // no index slot written yet
          r.t0 = 0;                                              // index slot of context 0 in the zeroed table
          C.cb0 = C.fetch(0); C.cb1 = C.fetch(1); C.cb2 = 0;
          C.st = C.total ? 1 : 2;
        }
        __syncwarp();
        if (mine && gl == 0) {
          C.sync->lpos = 0; C.sync->qpos = 0; C.sync->qfin = 0; C.sync->cpos = 0; C.sync->mfin = 0; C.sync->mpos = 0;
          C.sync->lafin = 0; C.sync->lstatus = 0; C.sync->jobid = job;
          __threadfence_block();
          C.sync->jobseq = C.seq;
        }
      }
    } else {
      bool fresh = false;
      if (C.st == 0 && C.sync->jobseq != C.seq) {
        __threadfence_block();
        C.seq += 1;
        C.job = C.sync->jobid;
        if (C.job == kDuoRetire) C.st = 3; else fresh = true;
      }
      uint32_t want = __ballot_sync(ZPQ_FULL, fresh && gl == 0);
      while (want) {
        const int src = __ffs(want) - 1;
        want &= want - 1;
        const uint32_t job = __shfl_sync(ZPQ_FULL, C.job, src);
        const EncJob J = P.ejobs[job];
        const bool mine = gbase == src;
        if (J.in_len == 0xFFFFFFFFu) continue;     // the pre-processing stage overflowed its slot: nothing to do
        uint8_t* arena_g = reinterpret_cast<uint8_t*>(__shfl_sync(ZPQ_FULL, (unsigned long long)w.arena, src));
        uint8_t* slice_g = reinterpret_cast<uint8_t*>(__shfl_sync(ZPQ_FULL, (unsigned long long)w.slice, src));
        init_block_state_role(plan, P.tab, arena_g, slice_g, lane, 3);   // hash tables, ICM maps, CM (Predictor.cs:99-113, 146-149)
        __syncwarp();
        if (mine) {
          C.in = P.in + J.in_off; C.pre_len = J.pre_len; C.total = J.pre_len + J.in_len;
          C.s = 0; C.status = BLK_OK;
          r.cxt = r.c = r.h = 0; r.t0 = r.t1 = 0; r.p = 0;
          C.w.c8 = 1; C.w.hmap4 = 1;
          F.hz = true;
          C.cb0 = C.fetch(0); C.cb1 = C.fetch(1); C.cb2 = 0;
          C.st = 1;
        }
      }
    }
    // ---- flow control: stay at most 6 - DB bytes ahead of the last role that reads the rings; the history
    //      role also needs the contexts of bytes s and s+1, i.e. the context role's byte s ----
    bool run = C.st == 1 && C.s + DB <= (DM::SPLIT ? C.sync->mpos : C.sync->qpos) + 6u;
    if (LR == 1) run = run && C.sync->lafin >= C.s + 1;
    if (!__any_sync(ZPQ_FULL, run)) {
      if (__all_sync(ZPQ_FULL, C.st == 3)) break;
      __nanosleep(LR == 0 ? 200 : 100);
      continue;
    }
    if (LS.hold(C.st == 1, run)) continue;
    ZPQ_T_IN
    if (run) {
      if (LR == 1) __threadfence_block();
      if (LR == 0) duo_context_byte<DM>(S, C, r, vm, env, gl, gmask, gbase);
      else duo_history_byte<DM>(S, C, r, F, row2, gl);
      ++C.s;
    }
    __syncwarp();
    ZPQ_T_OUT(__all_sync(ZPQ_FULL, run || C.st == 3))
    __threadfence_block();
    if (run) {
      if (LR == 0) {
        if (gl == 0) { if (C.status) C.sync->lstatus = C.status; C.sync->lafin = C.s; }
        if (C.s == C.total) C.st = 2;       // wait for the arithmetic coder to finish the block
      } else {
        if (gl == 0) C.sync->lpos = C.s;
        if (C.s == C.total) C.st = 0;
      }
    }
  }
  ZPQ_T_REPORT(LR == 0 ? "context" : "history")
}

// ==========================================================================================
// Coder role
// ==========================================================================================
template <int G>
struct CoderCtx {
  DuoSync* sync;
  int16_t* lring;            // [bit & 63][G]: lead-role output; the coder role overwrites an entry with its own prediction once it has used it
  const uint32_t* hsnap;
  int16_t* pfring;           // [bit & 63] stretched prediction of the last component, for the arithmetic coder warp
  uint8_t* arena;
  const uint8_t* in; const uint8_t* preamble;
  uint32_t pre_len, total, s, seq, job;
  uint32_t T, NB, bits;      // tick == bit time of the lead stage; bit k of `bits` = y(T-k)
  uint32_t cb0, cb1, cb2;
  int st;                    // 0 idle, 1 running, 3 retired
  __device__ __forceinline__ uint32_t fetch(uint32_t i) const {
    const uint8_t* q = i < pre_len ? preamble + i : in + (i - pre_len);
    return i < total ? (uint32_t)*q : 0u;
  }
  // partial byte (leading 1 + the bits already coded) of bit t as a stage d bits behind the lead sees it
  __device__ __forceinline__ uint32_t c8(uint32_t t, uint32_t d) const {
    const uint32_t k = t & 7;
    return ((bits >> (d + 1)) & ((1u << k) - 1)) | (1u << k);
  }
};

// per-lane constants of the coder role
struct CoderLane {
  int jL, kL;        // 1: input j / k is a lead-role component (read from the ring)
  int dj, dk;        // else: index into the producer's history (0 = its previous tick)
};

// MIX K evaluated by the group DM_ bits behind the lead (Predictor.cs:302-316, 427-439); see MixPipe.
// LMASK: bit i set = component i is a lead-role component.
// FINAL: this MIX is the model's last component (its prediction goes to the arithmetic coder).
template <int G, int K, int MIXLANE, int J0, int M, int RATE, unsigned MASK, unsigned CMASK, int DM_, unsigned LMASK, bool FINAL>
struct MixDuo {
  static __device__ __forceinline__ uint32_t rowoff(const CoderCtx<G>& C, uint32_t t, uint32_t d, int gl) {
    const uint32_t h = C.hsnap[((t >> 3) & 7) * G + MIXLANE];
    return ((h + (C.c8(t, d) & CMASK)) & MASK) * (uint32_t)(M * 4) + (uint32_t)gl * 4u;
  }
  // FAST: every group of the warp is inside its block with all stages (no range tests, whole-warp collectives);
  // `live` = this group really owns a running block (false: it rides along and must not store).
  // KT = bit of the byte the LEAD stage is at (the tick index inside the byte).
  // RING: mixer role -- every input is read from the ring (the coder role has replaced its entries by its
  // predictions), the output goes back into the ring (for a later MIX) and to the arithmetic coder's ring.
  template <bool FAST, int KC, bool RING>
  static __device__ __forceinline__ void tick(const Shared& S, const CoderCtx<G>& C, LaneRegs& r, const Hist& H, int& p, int& pmv, int gl,
                                               uint32_t gmask, int gbase, bool live, int krt) {
    const int KT = KC >= 0 ? KC : krt;
    constexpr int A = kDuoMixAhead;
    const uint32_t tm = C.T - (uint32_t)DM_, t2 = tm + A;
    const bool on2 = FAST || t2 < C.NB, onm = FAST || tm < C.NB;
    const uint32_t o2 = on2 ? rowoff(C, t2, (uint32_t)(DM_ - A), gl) : 0xFFFFFFFFu;
    int n2 = 0;
    if (on2 && gl < M) n2 = *reinterpret_cast<const int*>(r.mixtab[K] + o2);
    // input j0+gl of bit tm: from the ring (lead-role component) or from its lane's history
    int hv = 0;
    if (!RING) {
      hv = H.at((DM_ - r.d - 1) & 7);
      if (J0) hv = __shfl_sync(gmask, hv, gbase + ((J0 + gl) & (G - 1)));
    }
    const int lv = C.lring[(tm & (kDuoRing - 1)) * G + ((J0 + gl) & (G - 1))];
    const int pin = gl < M ? ((RING || ((LMASK >> ((J0 + gl) & 31)) & 1u)) ? lv : hv) : 0;
    // The row of bit tm was requested A ticks ago; a training of the last A bits that hit the same row has to be
    // forwarded (mt[j] = weight trained at bit tm-1-j).  With the whole partial byte in the row index
    // (CMASK == 255, >= 256 rows) rows of one byte are pairwise distinct, so only trainings of the PREVIOUS
    // byte can match: bit kBit of a byte checks j >= kBit only, and bits >= A of a byte check nothing.
    constexpr bool kDistinct = CMASK == 255u && MASK >= 255u;
    const int kBit = (KT - DM_) & 7;
    int wcur = r.mq[K][0];
    if (!(kDistinct && kBit >= A)) {
      // A hit is rare (it needs equal context hashes in consecutive bytes), and testing for it under a warp-uniform
      // branch keeps the weight independent of the previous ticks' training in the common case.
      bool hit = false;
#pragma unroll
      for (int j = 0; j < A; ++j)
        if (!(kDistinct && j < kBit)) hit = hit || r.mto[K][j] == r.mqo[K][0];
      if (__any_sync(gmask, hit)) {
#pragma unroll
        for (int j = A - 1; j >= 0; --j)
          if (!(kDistinct && j < kBit)) wcur = r.mto[K][j] == r.mqo[K][0] ? r.mt[K][j] : wcur;
      }
    }
    const int acc = grp_sum<G>(gmask, (wcur >> 8) * pin);
    const int pm = clamp2k(acc >> 8);
    const int y = (int)((C.bits >> DM_) & 1);
    const int err = ((y * 32767 - (int)S.squash[pm + 2048]) * RATE) >> 4;
    const int wn = clamp512k(wcur + ((err * pin + (1 << 12)) >> 13));
    if (onm && live && gl < M) *reinterpret_cast<int*>(const_cast<uint8_t*>(r.mixtab[K]) + r.mqo[K][0]) = wn;
    if (onm) {
#pragma unroll
      for (int j = A - 1; j > 0; --j) { r.mt[K][j] = r.mt[K][j - 1]; r.mto[K][j] = r.mto[K][j - 1]; }
      r.mt[K][0] = wn; r.mto[K][0] = r.mqo[K][0];
    }
#pragma unroll
    for (int j = 0; j < A - 1; ++j) { r.mq[K][j] = r.mq[K][j + 1]; r.mqo[K][j] = r.mqo[K][j + 1]; }
    r.mq[K][A - 1] = n2; r.mqo[K][A - 1] = o2;
    if (gl == MIXLANE) p = pm;
    pmv = pm;
    if (RING && gl == MIXLANE && onm && live) {
      C.lring[(tm & (kDuoRing - 1)) * G + MIXLANE] = (int16_t)pm;
      if (FINAL) C.pfring[tm & (kDuoRing - 1)] = (int16_t)pm;
    }
  }
  // lead role, start of byte s: pull the 8 rows byte s+1 will use into L2 (lane k: row of bit k)
  static __device__ __forceinline__ void prefetch(const LaneRegs& r, uint32_t hnext, uint32_t cnext, int gl, uint32_t gmask, int gbase) {
    const uint32_t h = __shfl_sync(gmask, hnext, gbase + MIXLANE);
    if (gl < 8) {
      const uint32_t c8 = (1u << gl) | (cnext >> (8 - gl));
      const uint8_t* row = r.mixtab[K] + ((h + (c8 & CMASK)) & MASK) * (uint32_t)(M * 4);
      prefetch_l2(row);
      if ((M * 4) & (M * 4 - 1)) prefetch_l2(row + M * 4 - 4);
    }
  }
};

// MIX described at run time (more than kMixRegs mixers): weights read and written in place.
template <int G, unsigned LMASK, bool RING>
__device__ __forceinline__ void duo_mix_rt(const Shared& S, const MixDesc& md, int dm, const CoderCtx<G>& C, LaneRegs& r, const Hist& H,
                                            int& p, int& pmv, int gl, uint32_t gmask, int gbase, bool live) {
  const uint32_t tm = C.T - (uint32_t)dm;
  const uint32_t h = C.hsnap[((tm >> 3) & 7) * G + md.lane];
  const uint32_t rowi = ((h + (C.c8(tm, (uint32_t)dm) & md.cmask)) & md.mask) * md.m;
  int* wp = reinterpret_cast<int*>(C.arena + md.tab) + rowi + gl;
  int hv = 0;
  if (!RING) {
    hv = H.at((dm - r.d - 1) & 7);
    hv = __shfl_sync(gmask, hv, gbase + ((md.j0 + gl) & (G - 1)));
  }
  const int lv = C.lring[(tm & (kDuoRing - 1)) * G + ((md.j0 + gl) & (G - 1))];
  const bool on = live && tm < C.NB && gl < md.m;
  const int pin = gl < md.m ? ((RING || ((LMASK >> ((md.j0 + gl) & 31)) & 1u)) ? lv : hv) : 0;
  const int wv = on ? *wp : 0;
  const int pm = clamp2k(grp_sum<G>(gmask, (wv >> 8) * pin) >> 8);
  const int y = (int)((C.bits >> dm) & 1);
  const int err = ((y * 32767 - (int)S.squash[pm + 2048]) * (int)md.rate) >> 4;
  if (on) *wp = clamp512k(wv + ((err * pin + (1 << 12)) >> 13));
  if (gl == md.lane) p = pm;
  pmv = pm;
  if (RING && gl == md.lane && live && tm < C.NB) {
    C.lring[(tm & (kDuoRing - 1)) * G + md.lane] = (int16_t)pm;
    if ((int)md.lane == (int)S.n - 1) C.pfring[tm & (kDuoRing - 1)] = (int16_t)pm;
  }
}

// One bit of the coder role.  DM (generated by zpq_codegen.cpp) supplies the constants
//   G, N, D, LDEPTH, HDEPTH, LMASK, FINAL_MIX, HAS_*, NEEDK and
//   DM::lanes(S, C, r, pj, pk, y, act, t, p, gl)                       AVG / MIX2 / SSE lanes
//   DM::mixes<FAST, K>(S, C, r, H, p, pmv, gl, gmask, gbase, live)      every MIX
// FAST: steady state of every group of the warp -- no range tests, whole-warp collectives, and the tick is one
// basic block, so consecutive ISSE chain links and the MIX overlap.
template <class DM, int KC, bool FAST>
__device__ __forceinline__ void duo_coder_tick(const Shared& S, CoderCtx<DM::G>& C, LaneRegs& r, const CoderLane& L, Hist& H, int gl,
                                                uint32_t gmask_rt, int gbase, bool live, int krt) {
  constexpr int G = DM::G;
  const int K = KC >= 0 ? KC : krt;
  constexpr uint32_t D = (uint32_t)DM::D;
  const uint32_t gmask = FAST ? ZPQ_FULL : gmask_rt;
  if (K == 0) C.cb2 = C.fetch(C.s + 2);
  C.bits = C.bits << 1 | ((C.cb0 >> (7 - K)) & 1u);
  const uint32_t t = C.T - (uint32_t)r.d;
  const bool act = live && gl < DM::N && (FAST || t < C.NB);
  const uint32_t slot = (t & (kDuoRing - 1)) * G;
  const int y = (int)((C.bits >> r.d) & 1);
  // ---- inputs of lane-owned components: ring (lead-role producer) or the producer lane's history ----
  int pj, pk = 0;
  {
    const int lj = C.lring[slot + r.srcj];
    int hj;
    if (DM::LDEPTH <= 1) hj = (int)(int16_t)(uint16_t)__shfl_sync(gmask, (uint32_t)H.h0, gbase + r.srcj);
    else {
      const uint64_t a0 = __shfl_sync(gmask, (unsigned long long)H.h0, gbase + r.srcj);
      const uint64_t a1 = DM::LDEPTH > 4 ? __shfl_sync(gmask, (unsigned long long)H.h1, gbase + r.srcj) : 0ull;
      hj = (int)(int16_t)(uint16_t)((L.dj < 4 ? a0 : a1) >> ((L.dj & 3) * 16));
    }
    pj = L.jL ? lj : hj;
    if (DM::NEEDK) {
      const int lk = C.lring[slot + r.srck];
      int hk;
      if (DM::LDEPTH <= 1) hk = (int)(int16_t)(uint16_t)__shfl_sync(gmask, (uint32_t)H.h0, gbase + r.srck);
      else {
        const uint64_t a0 = __shfl_sync(gmask, (unsigned long long)H.h0, gbase + r.srck);
        const uint64_t a1 = DM::LDEPTH > 4 ? __shfl_sync(gmask, (unsigned long long)H.h1, gbase + r.srck) : 0ull;
        hk = (int)(int16_t)(uint16_t)((L.dk < 4 ? a0 : a1) >> ((L.dk & 3) * 16));
      }
      pk = L.kL ? lk : hk;
    }
  }
  const int lown = C.lring[slot + gl];
  int p = ((DM::LMASK >> gl) & 1u) ? lown : 0;    // a lead-role component's prediction is its ring entry
  if (DM::HAS_ISSE) {
    // ---- ICM + ISSE, branch-free on all lanes (Predictor.cs:267-272, 317-326, 375-381, 440-449): both index a map with
    //      the bit history the history role found and advanced.  ISSE map: {weight, bias} pairs; ICM map: 22-bit
    //      probabilities with a 4-byte stride.
    const bool isse = r.type == C_ISSE, icm = r.type == C_ICM;
    const uint32_t bh = (uint32_t)lown & 255u;
    const uint32_t i0 = icm ? bh : bh * 2u;
    const int w0 = (int)r.cm[i0], w1 = (int)r.cm[isse ? i0 + 1u : i0];
    const int sp = S.stretch[((uint32_t)w0 >> 8) & 32767u];
    const int pe = clamp2k((w0 * pj + w1 * 64) >> 16);
    const int err = y * 32767 - (int)S.squash[pe + 2048];
    const int nw0 = isse ? clamp512k(w0 + ((err * pj + (1 << 12)) >> 13))
                         : (int)((uint32_t)w0 + (uint32_t)(((int)(y * 32767 - (int)((uint32_t)w0 >> 8))) >> 2));
    const int nw1 = clamp512k(w1 + ((err + 16) >> 5));
    if (act && (isse || icm)) r.cm[i0] = (uint32_t)nw0;
    if (act && isse) r.cm[i0 + 1u] = (uint32_t)nw1;
    p = isse ? pe : (icm ? sp : p);
  }
  DM::lanes(S, C, r, pj, pk, y, act, t, p, gl);
  int pmv = 0;
  if (!DM::SPLIT) DM::template mixes<FAST, KC, false>(S, C, r, H, p, pmv, gl, gmask, gbase, live, krt);
  H.push(p);
  // mixer role present: it reads this lane's prediction of bit t from the ring entry the lead role used for it
  if (DM::SPLIT && act && !((DM::LMASK >> gl) & 1u) && r.type != C_MIX) C.lring[slot + gl] = (int16_t)p;
  // the last component's prediction of bit t goes to the arithmetic coder warp
  if (!(DM::SPLIT && DM::FINAL_MIX) && gl == DM::N - 1 && act) C.pfring[t & (kDuoRing - 1)] = (int16_t)p;
  if (K == 7) { C.cb0 = C.cb1; C.cb1 = C.cb2; }
  ++C.T;
}

// lane-owned components other than ISSE, at bit t (see pipe_avg / pipe_mix2 / pipe_sse)
template <int G>
__device__ __forceinline__ void duo_avg(LaneRegs& r, int pj, int pk, bool act, int& p) {   // Predictor.cs:288-290
  if (act) p = ev_avg(r, pj, pk);
}
template <int G>
__device__ __forceinline__ void duo_mix2(const Shared& S, const CoderCtx<G>& C, LaneRegs& r, int pj, int pk, int y, bool act, uint32_t t,
                                          int& p, int gl) {   // Predictor.cs:291-301, 414-426
  if (!act) return;
  const uint32_t h = C.hsnap[((t >> 3) & 7) * G + gl];
  r.cxt = (h + (C.c8(t, (uint32_t)r.d) & r.a5)) & r.mask;
  r.t0 = reinterpret_cast<const uint16_t*>(r.tab)[r.cxt];
  r.p = ev_mix2(r, pj, pk);
  p = r.p;
  up_mix2(S, r, y, pj, pk);
}
template <int G>
__device__ __forceinline__ void duo_sse(const Shared& S, const CoderCtx<G>& C, LaneRegs& r, int pj, int y, bool act, uint32_t t, int& p,
                                         int gl) {   // Predictor.cs:327-340, 451-455
  if (!act) return;
  const uint32_t h = C.hsnap[((t >> 3) & 7) * G + gl];
  r.t0 = (int)((h + C.c8(t, (uint32_t)r.d)) * 32);
  r.p = ev_sse(S, r, pj);
  p = r.p;
  up_sse(S, r, y);
}

// One bit of the mixer role (models whose MIX components nothing lane-owned reads, DM::SPLIT): every MIX of the
// model, inputs from the ring, weights requested kDuoMixAhead bits ahead, output to the arithmetic coder's ring.
template <class DM, int KC, bool FAST>
__device__ __forceinline__ void duo_mix_tick(const Shared& S, CoderCtx<DM::G>& C, LaneRegs& r, int gl, uint32_t gmask_rt, int gbase, bool live, int krt) {
  const uint32_t gmask = FAST ? ZPQ_FULL : gmask_rt;
  const int K = KC >= 0 ? KC : krt;
  if (K == 0) C.cb2 = C.fetch(C.s + 2);
  C.bits = C.bits << 1 | ((C.cb0 >> (7 - K)) & 1u);
  Hist H; H.h0 = H.h1 = 0;
  int p = 0, pmv = 0;
  DM::template mixes<FAST, KC, true>(S, C, r, H, p, pmv, gl, gmask, gbase, live, krt);
  if (K == 7) { C.cb0 = C.cb1; C.cb1 = C.cb2; }
  ++C.T;
}

// ROLE 1: coder (prediction) role; ROLE 2: mixer role (only launched when DM::SPLIT).
template <class DM, int ROLE>
__device__ __forceinline__ void duo_coder_body(const CodecParams& P, uint8_t* smem, const Shared& S, int pair) {
  constexpr int G = DM::G, B = 32 / G;
  constexpr uint32_t D = (uint32_t)DM::D;
  const int lane = threadIdx.x & 31, gl = lane & (G - 1), grp = lane / G, gbase = lane & ~(G - 1);
  const uint32_t gmask = G == 32 ? ZPQ_FULL : (((1u << (G & 31)) - 1u) << gbase);
  const Plan* plan = P.plan;
  const uint32_t b = (uint32_t)pair * B + grp;
  const bool valid = b < P.wb && blockIdx.x * P.wb + b < P.resident;
  const uint32_t bb = valid ? b : 0u;
  Blk w;
  bind_block(P, smem, w, blockIdx.x * P.wb + bb, (int)bb);
  LaneRegs r;
  lane_load(S, P, w, r, gl < S.n ? gl : 0);
  if (gl >= S.n) { r.type = C_NONE; r.d = 0; r.srcj = r.srck = 0; }
  // maps live in the shared slice (the host launches this kernel only then): keep the pointer provably shared
  r.cm = (r.type == C_ISSE || r.type == C_ICM) ? reinterpret_cast<uint32_t*>(w.slice + S.comp[gl < S.n ? gl : 0].smem_cm)
                                               : reinterpret_cast<uint32_t*>(const_cast<int16_t*>(S.stretch));
  CoderLane L;
  {
    const int tj = S.comp[r.srcj].type, tk = S.comp[r.srck].type;
    L.jL = (tj == C_CONS || tj == C_CM || tj == C_MATCH) ? 1 : 0;
    L.kL = (tk == C_CONS || tk == C_CM || tk == C_MATCH) ? 1 : 0;
    L.dj = (r.d - (int)S.comp[r.srcj].delay - 1) & 7;
    L.dk = (r.d - (int)S.comp[r.srck].delay - 1) & 7;
  }
  CoderCtx<G> C;
  C.sync = reinterpret_cast<DuoSync*>(w.slice + plan->smem_sync);
  C.lring = reinterpret_cast<int16_t*>(w.slice + plan->smem_pring);
  C.hsnap = reinterpret_cast<const uint32_t*>(w.slice + plan->smem_hsnap);
  C.pfring = reinterpret_cast<int16_t*>(w.slice + plan->smem_pfring);
  C.arena = w.arena;
  C.preamble = P.preamble; C.in = P.in;
  C.pre_len = 0; C.total = 0; C.s = 0; C.seq = 0; C.job = 0;
  C.T = 0; C.NB = 0; C.bits = 0; C.cb0 = C.cb1 = C.cb2 = 0;
  C.st = valid ? 0 : 3;
  Hist H; H.h0 = H.h1 = 0;
  Lockstep LS;
  ZPQ_T_DECL
  __syncthreads();   // pairs with the lead role: sync words are initialised

  for (;;) {
    // ---- job management (warp-convergent) ----
    bool fresh = false;
    if (C.st == 0 && C.sync->jobseq != C.seq) {
      __threadfence_block();
      C.seq += 1;
      C.job = C.sync->jobid;
      if (C.job == kDuoRetire) C.st = 3; else fresh = true;
    }
    uint32_t want = __ballot_sync(ZPQ_FULL, fresh && gl == 0);
    while (want) {
      const int src = __ffs(want) - 1;
      want &= want - 1;
      const uint32_t job = __shfl_sync(ZPQ_FULL, C.job, src);
      const EncJob J = P.ejobs[job];
      const bool mine = gbase == src;
      if (J.in_len == 0xFFFFFFFFu) continue;     // the pre-processing stage overflowed its slot: the coder warp reports it
      uint8_t* arena_g = reinterpret_cast<uint8_t*>(__shfl_sync(ZPQ_FULL, (unsigned long long)w.arena, src));
      uint8_t* slice_g = reinterpret_cast<uint8_t*>(__shfl_sync(ZPQ_FULL, (unsigned long long)w.slice, src));
      init_block_state_role(plan, P.tab, arena_g, slice_g, lane, ROLE);
      __syncwarp();
      if (mine) {
        C.in = P.in + J.in_off;
        C.pre_len = J.pre_len; C.total = J.pre_len + J.in_len;
        C.s = 0; C.T = 0; C.NB = C.total * 8u; C.bits = 0;
        H.h0 = H.h1 = 0;
        r.cxt = 0; r.t0 = r.t1 = 0; r.p = 0;
        for (int k = 0; k < kMixRegs; ++k)
          for (int j = 0; j < kDuoMixAhead; ++j) { r.mq[k][j] = r.mt[k][j] = 0; r.mqo[k][j] = 0xFFFFFFFFu; r.mto[k][j] = 0xFFFFFFFEu; }
        C.cb0 = C.fetch(0); C.cb1 = C.fetch(1); C.cb2 = 0;
        C.st = 1;
      }
    }
    // ---- flow control ----
    // coder role: byte s needs the lead role's byte s (none while the pipeline drains); whoever reads what this
    // byte writes (mixer role: ring entries; arithmetic coder: final predictions) must be done with the slots.
    // mixer role: byte s needs the coder role's byte s; the arithmetic coder must be done with the slots.
    bool run = false;
    if (C.st == 1) {
      if (ROLE == 1) {
        const uint32_t need = C.s + 1 < C.total ? C.s + 1 : C.total;
        run = C.sync->lpos >= need;
        if (DM::SPLIT) run = run && C.s <= C.sync->mpos + 6u;
        if (!(DM::SPLIT && DM::FINAL_MIX)) run = run && C.s <= C.sync->cpos + 6u;
      } else {
        run = C.sync->qfin >= C.s + 1 && C.s <= C.sync->cpos + 6u;
      }
    }
    if (!__any_sync(ZPQ_FULL, run)) {
      if (__all_sync(ZPQ_FULL, C.st == 3)) break;
      __nanosleep(100);
      continue;
    }
    if (LS.hold(C.st == 1, run)) continue;
    // fast path: every group that owns a running block is ready and in its steady state (all stages inside
    // the block for the whole byte); groups without a running block ride along with their stores switched off
    constexpr uint32_t PRO = (D + 7u) / 8u;
    const bool live = C.st == 1;
    const bool steady = run && C.s >= PRO && C.s + 1 < C.total;
    if (run) { __threadfence_block(); if (gl == 0) { if (ROLE == 1) C.sync->qpos = C.s; else C.sync->mpos = C.s; } }
    __syncwarp();   // the lanes must be CONVERGED when they enter the fast path: its shuffles are whole-warp
    const bool fast = __all_sync(ZPQ_FULL, steady || !live);
    ZPQ_T_IN
    if (fast) {
#ifdef ZPQ_DUO_UNROLL
      if (ROLE == 1) duo_coder_tick<DM, 0, true>(S, C, r, L, H, gl, gmask, gbase, live, 0); else duo_mix_tick<DM, 0, true>(S, C, r, gl, gmask, gbase, live, 0);
      if (ROLE == 1) duo_coder_tick<DM, 1, true>(S, C, r, L, H, gl, gmask, gbase, live, 1); else duo_mix_tick<DM, 1, true>(S, C, r, gl, gmask, gbase, live, 1);
      if (ROLE == 1) duo_coder_tick<DM, 2, true>(S, C, r, L, H, gl, gmask, gbase, live, 2); else duo_mix_tick<DM, 2, true>(S, C, r, gl, gmask, gbase, live, 2);
      if (ROLE == 1) duo_coder_tick<DM, 3, true>(S, C, r, L, H, gl, gmask, gbase, live, 3); else duo_mix_tick<DM, 3, true>(S, C, r, gl, gmask, gbase, live, 3);
      if (ROLE == 1) duo_coder_tick<DM, 4, true>(S, C, r, L, H, gl, gmask, gbase, live, 4); else duo_mix_tick<DM, 4, true>(S, C, r, gl, gmask, gbase, live, 4);
      if (ROLE == 1) duo_coder_tick<DM, 5, true>(S, C, r, L, H, gl, gmask, gbase, live, 5); else duo_mix_tick<DM, 5, true>(S, C, r, gl, gmask, gbase, live, 5);
      if (ROLE == 1) duo_coder_tick<DM, 6, true>(S, C, r, L, H, gl, gmask, gbase, live, 6); else duo_mix_tick<DM, 6, true>(S, C, r, gl, gmask, gbase, live, 6);
      if (ROLE == 1) duo_coder_tick<DM, 7, true>(S, C, r, L, H, gl, gmask, gbase, live, 7); else duo_mix_tick<DM, 7, true>(S, C, r, gl, gmask, gbase, live, 7);
#else
      if (ROLE == 1) {
#pragma unroll 1
        for (int k = 0; k < 8; ++k) duo_coder_tick<DM, -1, true>(S, C, r, L, H, gl, gmask, gbase, live, k);
      } else {
        // the mixer's tick is short and rotates two register pipelines: unrolled, the rotation is free
        duo_mix_tick<DM, 0, true>(S, C, r, gl, gmask, gbase, live, 0);
        duo_mix_tick<DM, 1, true>(S, C, r, gl, gmask, gbase, live, 1);
        duo_mix_tick<DM, 2, true>(S, C, r, gl, gmask, gbase, live, 2);
        duo_mix_tick<DM, 3, true>(S, C, r, gl, gmask, gbase, live, 3);
        duo_mix_tick<DM, 4, true>(S, C, r, gl, gmask, gbase, live, 4);
        duo_mix_tick<DM, 5, true>(S, C, r, gl, gmask, gbase, live, 5);
        duo_mix_tick<DM, 6, true>(S, C, r, gl, gmask, gbase, live, 6);
        duo_mix_tick<DM, 7, true>(S, C, r, gl, gmask, gbase, live, 7);
      }
#endif
    } else if (run) {
#pragma unroll 1
      for (int k = 0; k < 8; ++k) {
        if (ROLE == 1) duo_coder_tick<DM, -1, false>(S, C, r, L, H, gl, gmask, gbase, true, k); else duo_mix_tick<DM, -1, false>(S, C, r, gl, gmask, gbase, true, k);
      }
    }
    __syncwarp();
    ZPQ_T_OUT(fast)
    __threadfence_block();
    if (run) {
      ++C.s;
      if (gl == 0) { if (ROLE == 1) C.sync->qfin = C.s; else C.sync->mfin = C.s; }
      if (C.T >= C.NB + D) C.st = 0;     // every prediction of the block is in the ring
    }
  }
  ZPQ_T_REPORT(ROLE == 1 ? "coder" : "mixer")
}

// ==========================================================================================
// Arithmetic coder warp (Encoder.cs:39-103): ONE warp per CTA, lane b codes block b of the CTA.  The coder is
// scalar per block and full of data-dependent branches (renormalisation, byte output), so it is kept out of
// the prediction warps: it reads the final stretched prediction of every bit from the block's ring.
// ==========================================================================================
template <class DM>
__device__ __forceinline__ void duo_arith_body(const CodecParams& P, uint8_t* smem, const Shared& S) {
  constexpr uint32_t D = (uint32_t)DM::D;
  constexpr uint32_t CD = (D + 7u + 7u) / 8u;     // byte c is complete in the ring once the coder role has finished c + CD bytes
  const int lane = threadIdx.x & 31;
  const Plan* plan = P.plan;
  const uint32_t b = (uint32_t)lane;
  const bool valid = b < P.wb && blockIdx.x * P.wb + b < P.resident;
  uint8_t* slice = smem + P.sm.slices + (valid ? b : 0u) * P.sm.slice_bytes;
  DuoSync* sync = reinterpret_cast<DuoSync*>(slice + plan->smem_sync);
  const int16_t* pfring = reinterpret_cast<const int16_t*>(slice + plan->smem_pfring);
  int st = valid ? 0 : 3;
  Lockstep LS;
  ZPQ_T_DECL
  uint32_t seq = 0, job = 0, c = 0, total = 0, pre_len = 0, cb = 0, cbn = 0, low = 1, high = 0xFFFFFFFFu;
  const uint8_t* in = P.in; uint8_t* out = P.out;
  uint64_t out_cap = 0, opos = 0;
  auto fetch = [&](uint32_t i) -> uint32_t {
    const uint8_t* q = i < pre_len ? P.preamble + i : in + (i - pre_len);
    return i < total ? (uint32_t)*q : 0u;
  };
  auto normalise = [&]() {                                       // Encoder.cs:95-102
    while ((high ^ low) < 0x1000000u) {
      if (opos < out_cap) out[opos] = (uint8_t)(high >> 24);
      ++opos;
      high = high << 8 | 255; low <<= 8; low += (low == 0);
    }
  };
  __syncthreads();   // pairs with the lead role: sync words are initialised

  for (;;) {
    if (st == 0 && sync->jobseq != seq) {
      __threadfence_block();
      seq += 1;
      job = sync->jobid;
      if (job == kDuoRetire) st = 3;
      else {
        const EncJob J = P.ejobs[job];
        if (J.in_len == 0xFFFFFFFFu) {     // the pre-processing stage overflowed its slot
          P.results[job].out_len = 0; P.results[job].status = BLK_OVERFLOW;
          __threadfence_block();
          sync->cdone = seq;
        } else {
          in = P.in + J.in_off; out = P.out + J.out_off;
          pre_len = J.pre_len; total = J.pre_len + J.in_len;
          out_cap = J.out_cap; opos = 0; low = 1; high = 0xFFFFFFFFu;
          c = 0; cb = fetch(0); cbn = fetch(1);
          st = 1;
        }
      }
    }
    const bool ready = st == 1 && ((DM::SPLIT && DM::FINAL_MIX) ? sync->mfin : sync->qfin) >= c + CD;
    if (!__any_sync(ZPQ_FULL, ready)) {
      if (__all_sync(ZPQ_FULL, st == 3)) break;
      __nanosleep(200);
      continue;
    }
    if (LS.hold(st == 1, ready)) continue;
    ZPQ_T_IN
    if (ready) {
      __threadfence_block();
      sync->cpos = c;
      ++low;                         // encode(0, 0) in front of every byte, Encoder.cs:49
      normalise();
#ifdef ZPQ_DUO_ARITH_UNROLL
#pragma unroll
#else
#pragma unroll 1     // rolled: this warp is never the slowest role, and 4 KB less hot code helps the other four
#endif
      for (int k = 0; k < 8; ++k) {
        const int pf = pfring[(c * 8u + k) & (kDuoRing - 1)];
        const uint32_t pr = (uint32_t)S.squash[pf + 2048] * 2 + 1;
        const uint32_t mid = low + (uint32_t)(((uint64_t)(high - low) * pr) >> 16);
        if ((cb >> (7 - k)) & 1) high = mid; else low = mid + 1;
        normalise();
      }
      ++c; cb = cbn; cbn = fetch(c + 1);
      if (c == total) {
        high = low;                  // encode(1, 0), Encoder.cs:46
        normalise();
        uint32_t status = sync->lstatus;
        if (opos > out_cap) status = BLK_OVERFLOW;
        P.results[job].out_len = opos; P.results[job].status = status;
        __threadfence_block();
        sync->cdone = seq;
        st = 0;
      }
    }
    __syncwarp();
    ZPQ_T_OUT(__all_sync(ZPQ_FULL, ready || st != 1))
  }
  ZPQ_T_REPORT("arith")
}

#ifndef ZPQ_DUO_ORDER
#define ZPQ_DUO_ORDER 4123     // decimal digits, first body first: 0,4,1,2,3 (a leading zero would make it octal)
#endif
// role 0 context, 1 history, 2 coder, 3 mixer (only in a SPLIT model), 4 arithmetic coder
template <class DM, int ROLE>
__device__ __forceinline__ void duo_role_body(const CodecParams& P, uint8_t* smem, const Shared& S, int pair) {
  if (ROLE == 0) duo_lead_body<DM, 0>(P, smem, S, pair);
  else if (ROLE == 1) duo_lead_body<DM, 1>(P, smem, S, pair);
  else if (ROLE == 2) duo_coder_body<DM, 1>(P, smem, S, pair);
  else if (ROLE == 3) { if (DM::SPLIT) duo_coder_body<DM, 2>(P, smem, S, pair); }
  else duo_arith_body<DM>(P, smem, S);
}

// Kernel body.  Per block group p: context role, history role, coder (prediction) role and, when the model's MIX
// components can run on their own, mixer role -- warp 4p+r, so role r of every group lands on SM sub-partition r and
// shares its instruction cache with copies of itself; the LAST warp is the arithmetic coder of every block of the CTA
// (it joins the context roles' sub-partition, the lightest).
template <class DM>
__device__ __forceinline__ void encode_duo_body(const CodecParams& P, uint8_t* smem) {
  Shared S;
  stage_shared(P, smem, S);
  constexpr int W = DM::SPLIT ? 4 : 3;
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int pair = warp / W, role = warp == nwarps - 1 ? 4 : warp % W;
  // The order of the role bodies in the kernel image is a tuning knob: the hot loops of the five roles together are about
  // as large as the SM's 32 KB instruction cache, and which of them collide depends on their addresses.
  constexpr int O0 = ZPQ_DUO_ORDER / 10000 % 10, O1 = ZPQ_DUO_ORDER / 1000 % 10, O2 = ZPQ_DUO_ORDER / 100 % 10,
                O3 = ZPQ_DUO_ORDER / 10 % 10, O4 = ZPQ_DUO_ORDER % 10;
  if (role == O0) duo_role_body<DM, O0>(P, smem, S, pair);
  else if (role == O1) duo_role_body<DM, O1>(P, smem, S, pair);
  else if (role == O2) duo_role_body<DM, O2>(P, smem, S, pair);
  else if (role == O3) duo_role_body<DM, O3>(P, smem, S, pair);
  else duo_role_body<DM, O4>(P, smem, S, pair);
}

}  // namespace zpq
