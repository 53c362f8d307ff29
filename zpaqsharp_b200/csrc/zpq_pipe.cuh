// zpq_pipe.cuh -- time-skewed ("pipelined") block encoder for sm_100a.  NVRTC-safe like zpq_devcore.cuh.
//
// When COMPRESSING, every bit of the block is known in advance, so the strict per-bit chain
//   predict(all components, in dependency order) -> code -> update          [Encoder.cs:50-57]
// does not have to be walked one bit at a time.  A ZPAQ model is feed-forward: component i reads
// only predictions of components j < i of the SAME bit and its own state left by EARLIER bits
// (Predictor.cs:245-475).  So component i may work `delay[i]` bits behind the leading bit, with
// delay[i] > delay[j] for every input j, and the arithmetic coder one bit behind component n-1:
//
//   tick T:   lead stage      bit T        HCOMP contexts, hash rows + bit histories (they depend on
//                                          the data only), CM, MATCH, ICM
//             lane i          bit T-d_i    ISSE / AVG / MIX2 / SSE: predict AND update in the same tick
//             MIX k (warp)    bit T-d_k    dot product by REDUX, weight update, rows requested 2 bits ahead
//             coder           bit T-D      32-bit arithmetic coder                [Encoder.cs:87-103]
//
// Predictions travel between stages through a small shared-memory ring indexed
// [bit time & (slots-1)][component]; bit histories of ISSE components travel from the lead stage
// to the owner's lagging stage through a second ring.  Every value a stage needs was produced
// in an earlier tick, so ONE __syncwarp per tick orders everything, the per-tick critical path is a
// single component (not the whole chain), and the stages of one tick are independent instruction
// streams the scheduler can overlap.  The arithmetic is exactly the reference's, bit for bit, in
// the same order per component; only the interleaving between components changes.
//
// Because the data are known, HCOMP runs one byte ahead and every table line the next byte will
// touch (first and second nibble hash rows, MATCH index slot, MIX rows) is prefetched into L2 a
// byte before it is needed.
//
// Decompression cannot be skewed (bit T is only known once all components of bit T are mixed);
// it keeps the lane-resident kernels of zpq_devcore.cuh.
#pragma once
#include "zpq_devcore.cuh"

namespace zpq {

struct PipeCtx {
  WarpCtx w;             // lead-stage partial byte state (c8, hmap4), arena, H
  uint32_t T;            // tick == bit time of the lead stage
  uint32_t NB;           // modeled bits of the block (8 x coded bytes)
  uint32_t bits;         // bit k = y(T-k)
  int16_t* pring;        // [slot][component] stretched predictions
  uint8_t* bhring;       // [slot][component] bit histories seen by the lead stage
  uint32_t* hsnap;       // [byte & 7][component] HCOMP contexts of the last 8 bytes
  uint32_t rmask, rstride;
};

// partial byte (leading 1 + the bits already coded) as a stage `d` bits behind the lead sees it
__device__ __forceinline__ uint32_t pipe_c8(const PipeCtx& W, uint32_t t, uint32_t d) {
  const uint32_t k = t & 7;
  return ((W.bits >> (d + 1)) & ((1u << k) - 1)) | (1u << k);
}

// ---- hash rows requested one nibble ahead (Predictor.cs:550-567) -----------------------------------
// The row a component needs for the NEXT nibble is known as soon as the current nibble starts (the data
// are known), so its three candidates are loaded then, the choice is made two ticks later into the
// lane's second row buffer, and the nibble boundary only swaps buffers.  The one thing that can
// invalidate the early choice is the write-back of the current row into the same 64-byte group;
// that case (hz) falls back to a plain look-up after the write-back.
struct FindAhead {
  uint4 r0, r1, r2;
  uint32_t h0, chk, at;
  bool hz;
};
__device__ __forceinline__ void find_issue(const LaneRegs& r, uint32_t cxt, FindAhead& f) {
  f.chk = (cxt >> (r.a1 + 2)) & 255;
  f.h0 = (cxt * 16) & r.mask;
  f.hz = ((f.h0 ^ r.c) & ~63u) == 0;
  f.r0 = *reinterpret_cast<const uint4*>(r.tab + f.h0);
  f.r1 = *reinterpret_cast<const uint4*>(r.tab + (f.h0 ^ 16));
  f.r2 = *reinterpret_cast<const uint4*>(r.tab + (f.h0 ^ 32));
}
__device__ __forceinline__ void find_resolve(FindAhead& f, uint8_t* rowbuf) {
  const uint32_t h0 = f.h0, h1 = h0 ^ 16, h2 = h0 ^ 32, chk = f.chk;
  uint4 v; uint32_t at;
  if ((f.r0.x & 255) == chk) { v = f.r0; at = h0; }
  else if ((f.r1.x & 255) == chk) { v = f.r1; at = h1; }
  else if ((f.r2.x & 255) == chk) { v = f.r2; at = h2; }
  else {
    const uint32_t p0 = (f.r0.x >> 8) & 255, p1 = (f.r1.x >> 8) & 255, p2 = (f.r2.x >> 8) & 255;
    at = (p0 <= p1 && p0 <= p2) ? h0 : (p1 < p2 ? h1 : h2);
    v = make_uint4(chk, 0, 0, 0);
  }
  f.at = at;
  *reinterpret_cast<uint4*>(rowbuf) = v;
}
// nibble boundary: the current row goes back to the table, the next one becomes current
__device__ __forceinline__ void find_swap(LaneRegs& r, const FindAhead& f, uint8_t*& row2, uint32_t cxt) {
  *reinterpret_cast<uint4*>(r.tab + r.c) = *reinterpret_cast<const uint4*>(r.row);
  if (f.hz) lane_find(r, cxt);
  else { uint8_t* t = r.row; r.row = row2; row2 = t; r.c = f.at; }
}

// ---- ICM + ISSE, branch-free on all lanes: predict and update bit T - r.d -----------------------
// (Predictor.cs:267-272, 317-326, 375-381, 440-449; the bit history itself was advanced by the lead stage)
template <bool CHECKED>
__device__ __forceinline__ void pipe_icm_isse(const Shared& S, const PipeCtx& W, LaneRegs& r, int lane) {
  const bool icm = r.type == C_ICM, isse = r.type == C_ISSE;
  const uint32_t t = W.T - (uint32_t)r.d;
  const bool act = (icm || isse) && (!CHECKED || t < W.NB);
  const uint32_t cell = (t & W.rmask) * W.rstride;
  const int y = (int)((W.bits >> r.d) & 1);
  const uint32_t bh = W.bhring[cell + lane];
  const int2 wt = *reinterpret_cast<const int2*>(r.cm + bh * 2);
  const int pin = W.pring[cell + r.srcj];
  const int sp = S.stretch[((uint32_t)wt.x >> 8) & 32767];
  const int pe = clamp2k((wt.x * pin + wt.y * 64) >> 16);
  const int p = icm ? sp : pe;
  const int err = y * 32767 - (int)S.squash[p + 2048];
  const uint32_t pn = (uint32_t)wt.x;
  int2 nw;
  nw.x = isse ? clamp512k(wt.x + ((err * pin + (1 << 12)) >> 13)) : (int)(pn + (uint32_t)(((int)(y * 32767 - (pn >> 8))) >> 2));
  nw.y = clamp512k(wt.y + ((err + 16) >> 5));
  if (act) {
    W.pring[cell + lane] = (int16_t)p;
    *reinterpret_cast<int2*>(r.cm + bh * 2) = nw;
  }
}

// ---- lead-stage components (delay 0) ---------------------------------------------------------------
__device__ __forceinline__ void pipe_cm(const Shared& S, const PipeCtx& W, LaneRegs& r, int lane, int y) {   // Predictor.cs:263-266, 365-373
  pa_cm(S, W.w, r);
  W.pring[(W.T & W.rmask) * W.rstride + lane] = (int16_t)r.p;
  up_cm(S, r, y);
}
__device__ __forceinline__ void pipe_match(const Shared& S, const PipeCtx& W, LaneRegs& r, int lane, int k, int y) {   // Predictor.cs:273-287, 382-411
  const uint32_t bit = (r.mbyte >> (7 - k)) & 1;
  const int pm = S.stretch[(S.dt2k[r.ma] * (1 - 2 * (int)bit)) & 32767];
  W.pring[(W.T & W.rmask) * W.rstride + lane] = (int16_t)(r.ma ? pm : 0);
  if ((int)bit != y) r.ma = 0;
  if (k == 7) match_byte(W.w, r, y);
}

// ---- lagging lane-owned components ----------------------------------------------------------------
template <bool CHECKED>
__device__ __forceinline__ void pipe_avg(const PipeCtx& W, LaneRegs& r, int lane) {   // Predictor.cs:288-290
  const uint32_t t = W.T - (uint32_t)r.d;
  if (CHECKED && t >= W.NB) return;
  const uint32_t cell = (t & W.rmask) * W.rstride;
  W.pring[cell + lane] = (int16_t)ev_avg(r, W.pring[cell + r.srcj], W.pring[cell + r.srck]);
}
template <bool CHECKED>
__device__ __forceinline__ void pipe_mix2(const Shared& S, const PipeCtx& W, LaneRegs& r, int lane) {   // Predictor.cs:291-301, 414-426
  const uint32_t t = W.T - (uint32_t)r.d;
  if (CHECKED && t >= W.NB) return;
  const uint32_t cell = (t & W.rmask) * W.rstride;
  const int y = (int)((W.bits >> r.d) & 1);
  const uint32_t h = W.hsnap[((t >> 3) & 7) * W.rstride + lane];
  r.cxt = (h + (pipe_c8(W, t, (uint32_t)r.d) & r.a5)) & r.mask;
  r.t0 = reinterpret_cast<const uint16_t*>(r.tab)[r.cxt];
  const int pj = W.pring[cell + r.srcj], pk = W.pring[cell + r.srck];
  r.p = ev_mix2(r, pj, pk);
  W.pring[cell + lane] = (int16_t)r.p;
  up_mix2(S, r, y, pj, pk);
}
template <bool CHECKED>
__device__ __forceinline__ void pipe_sse(const Shared& S, const PipeCtx& W, LaneRegs& r, int lane) {   // Predictor.cs:327-340, 451-455
  const uint32_t t = W.T - (uint32_t)r.d;
  if (CHECKED && t >= W.NB) return;
  const uint32_t cell = (t & W.rmask) * W.rstride;
  const int y = (int)((W.bits >> r.d) & 1);
  const uint32_t h = W.hsnap[((t >> 3) & 7) * W.rstride + lane];
  r.t0 = (int)((h + pipe_c8(W, t, (uint32_t)r.d)) * 32);
  r.p = ev_sse(S, r, W.pring[cell + r.srcj]);
  W.pring[cell + lane] = (int16_t)r.p;
  up_sse(S, r, y);
}

// ---- MIX K, evaluated by the whole warp DM bits behind the lead (Predictor.cs:302-316, 427-439) ----
// Lane j owns weight j of every row.  r.mw[K] is the weight of the row of bit tm = T - DM,
// r.mn0[K] the one of bit tm+1 (requested a tick ago), and the row of bit tm+2 is requested now;
// rows are known ahead because the data are.  A row that repeats within those three bits gets the
// freshly trained weights forwarded instead of the stale loaded ones.
template <int K, int MIXLANE, int J0, int M, int RATE, unsigned MASK, unsigned CMASK, int DM>
struct MixPipe {
  static __device__ __forceinline__ uint32_t rowoff(const PipeCtx& W, uint32_t t, uint32_t d, int lane) {
    const uint32_t h = W.hsnap[((t >> 3) & 7) * W.rstride + MIXLANE];
    return ((h + (pipe_c8(W, t, d) & CMASK)) & MASK) * (uint32_t)(M * 4) + (uint32_t)lane * 4u;
  }
  // Registers of MIX K in lane j: r.mw = loaded weight j of the row of bit tm, r.mn0 = of bit tm+1;
  // r.mn1 / r.mo1 = a weight trained after the row was requested and the row it belongs to (forwarding:
  // a load is never consumed in the tick that issued it, the select happens when the row is used).
  template <bool CHECKED>
  static __device__ __forceinline__ void tick(const Shared& S, const PipeCtx& W, LaneRegs& r, int lane) {
    const uint32_t tm = W.T - (uint32_t)DM, t2 = tm + kPipeMixAhead;
    uint32_t o2 = 0xFFFFFFFFu;
    int n2 = 0;
    if (!CHECKED || t2 < W.NB) {
      o2 = rowoff(W, t2, (uint32_t)(DM - kPipeMixAhead), lane);
      if (lane < M) n2 = *reinterpret_cast<const int*>(r.mixtab[K] + o2);
    }
    if (!CHECKED || tm < W.NB) {
      const uint32_t cell = (tm & W.rmask) * W.rstride;
      const int pin = lane < M ? (int)W.pring[cell + J0 + lane] : 0;
      // the two most recent trainings (bits tm-1, tm-2) may have hit this row after it was requested
      const int w = r.mo1[K] == r.moff[K] ? r.mn1[K] : (r.mo2[K] == r.moff[K] ? r.mn2[K] : r.mw[K]);
      const int acc = __reduce_add_sync(ZPQ_FULL, (w >> 8) * pin);
      const int pm = clamp2k(acc >> 8);
      if (lane == MIXLANE) W.pring[cell + MIXLANE] = (int16_t)pm;
      const int y = (int)((W.bits >> DM) & 1);
      const int err = ((y * 32767 - (int)S.squash[pm + 2048]) * RATE) >> 4;
      const int wn = clamp512k(w + ((err * pin + (1 << 12)) >> 13));
      if (lane < M) *reinterpret_cast<int*>(const_cast<uint8_t*>(r.mixtab[K]) + r.moff[K]) = wn;
      r.mn2[K] = r.mn1[K]; r.mo2[K] = r.mo1[K];
      r.mn1[K] = wn; r.mo1[K] = r.moff[K];
    }
    r.mw[K] = r.mn0[K]; r.moff[K] = r.mo0[K];
    r.mn0[K] = n2; r.mo0[K] = o2;
  }
  // lead stage, start of byte s: pull the 8 rows byte s+1 will use into L2 (lane k: row of bit k)
  static __device__ __forceinline__ void prefetch(const PipeCtx& W, const LaneRegs& r, uint32_t hnext, uint32_t cnext, int lane) {
    const uint32_t h = __shfl_sync(ZPQ_FULL, hnext, MIXLANE);
    if (lane < 8) {
      const uint32_t c8 = (1u << lane) | (cnext >> (8 - lane));
      const uint8_t* row = r.mixtab[K] + ((h + (c8 & CMASK)) & MASK) * (uint32_t)(M * 4);
      prefetch_l2(row);
      if ((M * 4) & (M * 4 - 1)) prefetch_l2(row + M * 4 - 4);   // rows that are not a power of two long may straddle a line
    }
  }
};

// MIX described at run time (more than kMixRegs mixers): weights read and written in place.
template <bool CHECKED>
__device__ __forceinline__ void pipe_mix_rt(const Shared& S, const MixDesc& md, int dm, const PipeCtx& W, LaneRegs& r, int lane) {
  const uint32_t tm = W.T - (uint32_t)dm;
  if (CHECKED && tm >= W.NB) return;
  const uint32_t cell = (tm & W.rmask) * W.rstride;
  const uint32_t h = W.hsnap[((tm >> 3) & 7) * W.rstride + md.lane];
  const uint32_t rowi = ((h + (pipe_c8(W, tm, (uint32_t)dm) & md.cmask)) & md.mask) * md.m;
  int* wp = reinterpret_cast<int*>(W.w.arena + md.tab) + rowi + lane;
  const int pin = lane < md.m ? (int)W.pring[cell + md.j0 + lane] : 0;
  const int w = lane < md.m ? *wp : 0;
  const int pm = clamp2k(__reduce_add_sync(ZPQ_FULL, (w >> 8) * pin) >> 8);
  if (lane == md.lane) W.pring[cell + md.lane] = (int16_t)pm;
  const int y = (int)((W.bits >> dm) & 1);
  const int err = ((y * 32767 - (int)S.squash[pm + 2048]) * (int)md.rate) >> 4;
  if (lane < md.m) *wp = clamp512k(w + ((err * pin + (1 << 12)) >> 13));
}

// ------------------------------------------------------------------------------------------
// Encoder body.  PM (generated by zpq_codegen.cpp) supplies:
//   PM::D, PM::N                  coder delay, components (compile time)
//   PM::lead0(S, W, r, lane, k, y)  CM / MATCH components at the lead bit
//   PM::lag<CHECKED>(S, W, r, lane) lane-owned components behind the lead (ICM/ISSE/AVG/MIX2/SSE)
//   PM::mixes<CHECKED>(S, W, r, lane) every MIX
//   PM::prefetch(W, r, hnext, cnext, lane)   MIX rows of the next byte
//   PM::hcomp(S, w, vm, env, c, lane)        HCOMP (compiled or interpreted)
// CHECKED = false is the steady state: every stage is inside the block, no range tests.
// ------------------------------------------------------------------------------------------
struct PipeEnc {            // per-block coder and input state, warp-uniform
  const uint8_t* in; const uint8_t* preamble; uint8_t* out;
  uint32_t pre_len, total;
  uint64_t out_cap, opos;
  uint32_t low, high, status;
  uint32_t cb0, cb1, cb2;   // bytes s, s+1, s+2 of the lead stage
  __device__ __forceinline__ uint32_t fetch(uint32_t s) const {
    if (s < pre_len) return preamble[s];
    return s < total ? in[s - pre_len] : 0u;
  }
  __device__ __forceinline__ void normalise(int lane) {     // Encoder.cs:95-102
    while ((high ^ low) < 0x1000000u) {
      if (lane == 0 && opos < out_cap) out[opos] = (uint8_t)(high >> 24);
      ++opos;
      high = high << 8 | 255; low <<= 8; low += (low == 0);
    }
  }
};

template <class PM, bool CHECKED>
__device__ __forceinline__ void pipe_tick(const Shared& S, PipeCtx& W, LaneRegs& r, PipeEnc& E, VM& vm, VMEnv& env, FindAhead& F,
                                          uint8_t*& row2, uint32_t& hnext, int lane, int k, bool lead, bool hashed) {
  constexpr uint32_t D = (uint32_t)PM::D;
  const uint32_t s = W.T >> 3;
  int yL = 0;
  if (!CHECKED || lead) {
    if (k == 0) {
      // ---- the lead stage enters byte s ----
      E.cb2 = E.fetch(s + 2);
      r.h = W.hsnap[(s & 7) * W.rstride + (lane < (int)W.rstride ? lane : 0)];
      W.w.c8 = 1; W.w.hmap4 = 1;
      if (hashed) {
        if (s == 0) lane_find(r, r.h + 16);
        else find_swap(r, F, row2, r.h + 16);          // the previous byte's last row goes back, the requested one comes in
        find_issue(r, r.h + 16u * (16u + (E.cb0 >> 4)), F);   // second nibble of this byte
      }
      // contexts of byte s+1 (HCOMP sees byte s, ZPAQL.cs:1253-1265) and the lines that byte will touch
      if (PM::hcomp(S, W.w, vm, env, E.cb0, lane)) E.status = BLK_ZPAQL;
      const uint32_t hn = W.w.H[lane & W.w.hmask];
      if (lane < (int)W.rstride) W.hsnap[((s + 1) & 7) * W.rstride + lane] = hn;
      hnext = hn;
      if (hashed) {
        prefetch_l2(r.tab + (((hn + 16u) * 16u) & r.mask));
        prefetch_l2(r.tab + (((hn + 16u * (16u + (E.cb1 >> 4))) * 16u) & r.mask));
      }
      if (r.type == C_MATCH) prefetch_l2(reinterpret_cast<const uint32_t*>(r.tab) + (hn & r.mask));
      PM::prefetch(W, r, hn, E.cb1, lane);
    }
    yL = (int)((E.cb0 >> (7 - k)) & 1);
  }
  W.bits = W.bits << 1 | (uint32_t)yL;
  if (!CHECKED || lead) {
    // bit histories depend on the data only: advance them at the lead, hand the value to the owner's stage
    const uint32_t si = (uint32_t)W.w.hmap4 & 15u;
    const uint32_t bh = r.row[si];
    if (hashed) {
      W.bhring[(W.T & W.rmask) * W.rstride + lane] = (uint8_t)bh;
      r.row[si] = S.ns[bh * 4 + yL];
    }
    PM::lead0(S, W, r, lane, k, yL);
  }
  PM::template lag<CHECKED>(S, W, r, lane);
  PM::template mixes<CHECKED>(S, W, r, lane);
  if (!CHECKED || (W.T >= D && W.T - D < W.NB)) {
    // ---- arithmetic coder, bit T-D (Encoder.cs:44-57, 87-103) ----
    if (((k - (int)D) & 7) == 0) { ++E.low; E.normalise(lane); }   // encode(0, 0) in front of every byte
    const int pf = W.pring[((W.T - D) & W.rmask) * W.rstride + (PM::N - 1)];
    const uint32_t pr = (uint32_t)S.squash[pf + 2048] * 2 + 1;
    const uint32_t mid = E.low + (uint32_t)(((uint64_t)(E.high - E.low) * pr) >> 16);
    if ((W.bits >> D) & 1) E.high = mid; else E.low = mid + 1;
    E.normalise(lane);
  }
  if (!CHECKED || lead) {
    // ---- shift the bit into the lead's c8 / hmap4 (Predictor.cs:463-474) ----
    const int c8 = W.w.c8 * 2 + yL;
    if (k == 7) {
      E.cb0 = E.cb1; E.cb1 = E.cb2;                    // (the row swap waits for the next byte's context)
    } else if (k == 3) {
      W.w.hmap4 = (W.w.hmap4 & 0xf) << 5 | yL << 4 | 1;
      if (hashed) {
        find_swap(r, F, row2, r.h + 16u * (uint32_t)c8);
        find_issue(r, hnext + 16u, F);                 // first nibble of the next byte
      }
    } else {
      W.w.hmap4 = (W.w.hmap4 & 0x1f0) | (((W.w.hmap4 & 0xf) * 2 + yL) & 0xf);
      if ((k == 2 || k == 5) && hashed) find_resolve(F, row2);
    }
    W.w.c8 = c8;
  }
  __syncwarp();
  ++W.T;
}

template <class PM>
__device__ __forceinline__ void encode_pipe_body(const CodecParams& P, uint8_t* smem) {
  Shared S;
  stage_shared(P, smem, S);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp;
  if (gw >= P.resident) return;
  const Plan* plan = P.plan;
  Blk w;
  bind_block(P, smem, w, gw, warp);
  LaneRegs r;
  lane_load(S, P, w, r, lane);
  const bool hashed = (r.type == C_ICM || r.type == C_ISSE);
  // the host launches this kernel only when every ICM/ISSE map lives in the shared slice
  r.cm = hashed ? reinterpret_cast<uint32_t*>(w.slice + S.comp[lane < S.n ? lane : 0].smem_cm)
                : reinterpret_cast<uint32_t*>(const_cast<int16_t*>(S.stretch));
  PipeCtx W;
  W.pring = reinterpret_cast<int16_t*>(w.slice + plan->smem_pring);
  W.bhring = w.slice + plan->smem_bhring;
  W.hsnap = reinterpret_cast<uint32_t*>(w.slice + plan->smem_hsnap);
  W.rmask = PM::RS - 1;
  W.rstride = PM::RSTRIDE;
  VM vm; VMEnv env;
  PipeEnc E;
  FindAhead F;
  F.r0 = F.r1 = F.r2 = make_uint4(0, 0, 0, 0); F.h0 = F.chk = F.at = 0; F.hz = true;
  uint8_t* row2 = r.row + 512;     // second row buffer of this lane (the plan reserves 2 x 512 bytes)
  uint32_t hnext = 0;
  constexpr uint32_t D = (uint32_t)PM::D;
  constexpr uint32_t PRO = (D + 7) / 8;      // lead bytes before every stage is inside the block

  for (;;) {
    uint32_t job = 0;
    if (lane == 0) job = atomicAdd(P.queue, 1u);
    job = __shfl_sync(ZPQ_FULL, job, 0);
    if (job >= P.njobs) break;
    const EncJob J = P.ejobs[job];
    if (J.in_len == 0xFFFFFFFFu) {   // the pre-processing stage overflowed its slot
      if (lane == 0) { P.results[job].out_len = 0; P.results[job].status = BLK_OVERFLOW; }
      continue;
    }
    E.in = P.in + J.in_off; E.preamble = P.preamble; E.out = P.out + J.out_off;
    E.pre_len = J.pre_len; E.total = J.pre_len + J.in_len;   // the host routes blocks >= 2^28 bytes to the unskewed kernel
    E.out_cap = J.out_cap; E.opos = 0; E.status = BLK_OK;
    E.low = 1; E.high = 0xFFFFFFFFu;

    // ---- Predictor.init + ZPAQL.inith ----
    init_block_state(plan, P.tab, w.arena, w.slice, lane);
    __syncwarp();
    r.cxt = r.c = r.ma = r.mb = r.mpos = r.h = 0;
    r.t0 = r.t1 = 0; r.p = 0;
    r.mbyte = 0; r.mbit = 0;
    for (int k = 0; k < kMixRegs; ++k) { r.mw[k] = r.mn0[k] = r.mn1[k] = r.mn2[k] = 0; r.moff[k] = r.mo0[k] = 0xFFFFFFFFu; r.mo1[k] = r.mo2[k] = 0xFFFFFFFEu; }
    if (r.type == C_MATCH) r.tab2[0] = 1;                         // Predictor.cs:118
    W.w.arena = w.arena; W.w.H = w.H; W.w.hmask = w.hmask; W.w.c8 = 1; W.w.hmap4 = 1;
    W.T = 0; W.NB = E.total * 8u; W.bits = 0;
    vm.b = vm.c = vm.d = vm.f = 0;
    env.code = S.hcomp; env.len = S.hcomp_len;
    env.H = w.H; env.hmask = w.hmask; env.M = w.M; env.mmask = w.mmask; env.R = w.R;
    env.out = nullptr; env.out_pos = 0; env.out_cap = 0;
    for (uint32_t i = lane; i < 8 * W.rstride; i += 32) W.hsnap[i] = 0;   // contexts of byte 0 are H == 0
    if (lane < (int)W.rstride) {
      const int16_t p0 = r.type == C_CONS ? (int16_t)(((int)r.a1 - 128) * 4) : (int16_t)0;   // Predictor.cs:96-98
      for (uint32_t q = 0; q <= W.rmask; ++q) { W.pring[q * W.rstride + lane] = p0; W.bhring[q * W.rstride + lane] = 0; }
    }
    __syncwarp();
    E.cb0 = E.fetch(0); E.cb1 = E.fetch(1); E.cb2 = 0;

    const uint32_t nbytes = E.total + PRO;    // lead bytes plus the ticks that drain the pipeline
    uint32_t s = 0;
    while (s < nbytes) {
      if (s >= PRO && s + 1 < E.total) {
        // steady state: whole bytes with every stage active
        do {
#pragma unroll
          for (int k = 0; k < 8; ++k) pipe_tick<PM, false>(S, W, r, E, vm, env, F, row2, hnext, lane, k, true, hashed);
          ++s;
        } while (s + 1 < E.total && E.opos <= E.out_cap && !E.status);
      } else {
        const bool lead = s < E.total;
#pragma unroll 1
        for (int k = 0; k < 8; ++k) pipe_tick<PM, true>(S, W, r, E, vm, env, F, row2, hnext, lane, k, lead, hashed);
        ++s;
      }
      if (E.opos > E.out_cap || E.status) break;
    }
    E.high = E.low;  // encode(1, 0), Encoder.cs:46
    E.normalise(lane);
    if (E.opos > E.out_cap) E.status = BLK_OVERFLOW;
    __syncwarp();
    if (lane == 0) { P.results[job].out_len = E.opos; P.results[job].status = E.status; }
  }
}

}  // namespace zpq
