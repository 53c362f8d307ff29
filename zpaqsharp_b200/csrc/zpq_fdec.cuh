// zpq_fdec.cuh -- speculative block decoder for sm_100a.  NVRTC-safe like zpq_devcore.cuh.
//
// Decoding is a strict chain per bit: bit t is only known when every component of bit t has been mixed and the
// arithmetic decoder has compared (Decoder.cs:136-158, Predictor.cs:245-475), so the encoder's time skew and role split
// do not apply.  HBM caps the blocks in flight (mid.cfg: ~1600), so decode throughput = blocks in flight / latency of
// one bit, and this kernel is built to shorten that latency.  One warp still owns a block, but nothing that can be known
// before the bit is decoded waits for it:
//
//   * both outcomes ahead.  While bit t is being mixed and decoded, every ICM / ISSE lane reads the bit-history slot
//     and the map entry bit t+1 will use under y_t = 0 AND under y_t = 1 (Predictor.cs:267-272, 317-326), trains the
//     entry of bit t under both outcomes (Predictor.cs:375-381, 440-449) and forwards the trained value when bit t+1
//     meets the same bit history.  When y_t arrives one select per value yields the state of bit t+1.
//   * every component in every lane.  The lanes publish {weight, bias} (ISSE) or the stretched prediction (CONS / ICM)
//     of both outcomes in shared memory; all lanes read all slots, so after the select the whole dependency chain
//     (ICM -> ISSE -> ISSE ... -> MIX) is evaluated redundantly in registers without a shuffle or a barrier on the path.
//   * hash rows a nibble ahead.  A model with n <= 8 components leaves 24 lanes idle: lane 8g+i is a helper of
//     component i.  When two bits of a nibble are known the four hash rows the NEXT nibble can need are fetched, one
//     per lane group (Predictor.cs:550-567: three candidates, check byte, replacement priority), and resolved into the
//     helper's row slot; at the nibble boundary the owner lane only switches its row pointer.  The first nibble of a
//     byte cannot be fetched ahead (its context is HCOMP of the byte being decoded) and is looked up directly.
//   * MIX rows two bits ahead.  Lane group g requests weight row 4*c8 + g, so a row has two bit times to arrive
//     (Predictor.cs:302-316); the trained row is written behind the decoder's back (Predictor.cs:427-439).
//   * MATCH is warp-uniform state: its prediction for all 8 bits follows from the predicted byte and the match length;
//     the index slot, the predicted bytes and 32 bytes of the candidate's history are requested early in the byte and
//     the end-of-byte verification (Predictor.cs:382-411) is one ballot.
//   * the 32-bit arithmetic decoder (Decoder.cs:136-158) runs redundantly in all lanes.
//
// The post-processor (PostProcessor.cs:37-86) is a separate pass (k_post* in zpq_kernels.cu): this kernel stores the
// bytes the model decodes -- the PCOMP preamble and the transformed data -- into the job's raw stream.
//
// Applies to models with n <= 8 components out of CONS, CM, ICM, ISSE, MATCH (one) and MIX (with the whole partial
// byte in its context and at most 8 inputs); everything else keeps the lane-resident decoder of zpq_devcore.cuh.
// The arithmetic is exactly the reference's, per component in the same order, bit for bit.
#pragma once
#include "zpq_pipe.cuh"

namespace zpq {

constexpr int kFdecThreads = 384;     // 12 warps per CTA at most: 170 registers per thread

// Loads the compiler must issue where they are written (it otherwise sinks them behind the branches of the row choice and a
// look-up becomes three dependent trips to HBM).  The addresses are global memory (arena tables).
__device__ __forceinline__ uint4 fd_ldg128(const void* p) {
  uint4 v;
  asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ int fd_ldg32(const void* p) {
  int v;
  asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t fd_ldg8(const void* p) {
  uint32_t v;
  asm volatile("ld.global.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// coded bytes of a segment, two bytes ahead of the decoder (Decoder.get, Decoder.cs:112-122); warp-uniform
struct FdIn {
  const uint8_t* p;
  uint32_t len, pos;
  uint32_t n0, n1;
  __device__ __forceinline__ void open(const uint8_t* q, uint64_t l) {
    p = q; len = (uint32_t)l; pos = 0;
    n0 = l > 0 ? q[0] : 0u; n1 = l > 1 ? q[1] : 0u;
  }
  __device__ __forceinline__ uint32_t take(uint32_t& status) {
    const uint32_t c = n0;
    if (pos >= len) status = BLK_CORRUPT;       // "unexpected end of file"
    ++pos; n0 = n1;
    n1 = pos + 1 < len ? (uint32_t)p[pos + 1] : 0u;
    return c;
  }
};

// MIX K_ of a model with 8 lanes per component group (Predictor.cs:302-316, 427-439).  Lane gl < M of group 0 owns
// weight gl of the current row (r.mw); r.mn0 / r.mn1 hold, per lane group, a candidate row of the next bit / the bit
// after (see the file header).  Requires CMASK == 255 and >= 256 contexts: the rows of one byte are pairwise distinct.
template <int K_, int MIXLANE, int J0, int M, int RATE, unsigned MASK>
struct FMix {
  static __device__ __forceinline__ uint32_t rowoff(uint32_t hm, uint32_t c8) { return ((hm + c8) & MASK) * (uint32_t)(M * 4); }
  static __device__ __forceinline__ void new_byte(LaneRegs& r, WarpCtx& W, uint32_t hm, int lane) {
    const uint32_t gl = (uint32_t)lane & 7u, grp = (uint32_t)lane >> 3, lo = gl < (uint32_t)M ? gl * 4u : 0u;
    W.mixh[K_] = hm;
    r.moff[K_] = rowoff(hm, 1u) + lo;
    r.mw[K_] = fd_ldg32(r.mixtab[K_] + r.moff[K_]);
    r.mn0[K_] = fd_ldg32(r.mixtab[K_] + rowoff(hm, 2u + (grp & 1u)) + lo);
    r.mn1[K_] = fd_ldg32(r.mixtab[K_] + rowoff(hm, 4u + grp) + lo);
  }
  static __device__ __forceinline__ int predict(const LaneRegs& r, int lane, int pin) {
    const int prod = lane < M ? (r.mw[K_] >> 8) * pin : 0;
    return clamp2k(__reduce_add_sync(ZPQ_FULL, prod) >> 8);
  }
  // bit KB of the byte is decoded: train the row (nothing reads it again inside this byte), move to the next one
  template <int KB>
  static __device__ __forceinline__ void advance(const Shared& S, LaneRegs& r, const WarpCtx& W, int lane, int y, int yprev, int pin, int pm,
                                                 uint32_t c8new) {
    const int err = ((y * 32767 - (int)S.squash[pm + 2048]) * RATE) >> 4;
    const int w = clamp512k(r.mw[K_] + ((err * pin + (1 << 12)) >> 13));
    if (lane < M) *reinterpret_cast<int*>(const_cast<uint8_t*>(r.mixtab[K_]) + r.moff[K_]) = w;
    if (KB < 7) {
      const uint32_t gl = (uint32_t)lane & 7u, grp = (uint32_t)lane >> 3, lo = gl < (uint32_t)M ? gl * 4u : 0u;
      const int g = KB == 0 ? y : 2 * yprev + y;
      r.mw[K_] = __shfl_sync(ZPQ_FULL, r.mn0[K_], (int)gl + 8 * g);
      r.moff[K_] = rowoff(W.mixh[K_], c8new) + lo;
      r.mn0[K_] = r.mn1[K_];
      if (KB <= 4) r.mn1[K_] = fd_ldg32(r.mixtab[K_] + rowoff(W.mixh[K_], c8new * 4u + grp) + lo);
    }
  }
};

template <class FD>
struct FdCtx {
  // ---- warp-uniform ----
  uint32_t low, high, curr, status;
  int c8, hmap4, yprev;
  FdIn in;
  WarpCtx W;
  // MATCH (Predictor.cs:273-287, 382-411), kept redundantly in every lane
  uint32_t ma, mb, mpos, mbyte, mh, idxv, lastslot, lastpos;
  uint32_t pbA, pbB;
  int mp0, mp1;
  uint64_t hist;            // the last 8 decoded bytes, newest in the low byte
  // ---- per lane ----
  uint32_t be, ae;          // MATCH verification: byte `lane` of the candidate's / of our own history
  uint8_t* rowp;            // current hash row (shared memory)
  uint32_t at;              // its byte offset in the component's table
  uint32_t nat, nhz;        // helper: offset and hazard flag of the resolved candidate row
  uint4 f0, f1, f2;         // helper: the three candidate rows in flight
  uint32_t fh0, fchk;
  int2 cw[8];               // selected slot of every component for the current bit
};

// Hash-row candidates (Predictor.cs:550-567): issue the loads ...
__device__ __forceinline__ void fd_find_issue(const LaneRegs& r, uint32_t cxt, uint4& f0, uint4& f1, uint4& f2, uint32_t& h0, uint32_t& chk) {
  chk = (cxt >> (r.a1 + 2)) & 255u;
  h0 = (cxt * 16u) & r.mask;
  f0 = fd_ldg128(r.tab + h0);
  f1 = fd_ldg128(r.tab + (h0 ^ 16u));
  f2 = fd_ldg128(r.tab + (h0 ^ 32u));
}
// ... and pick: the row whose check byte matches, else the lowest-priority one, emptied.  Branch-free.
__device__ __forceinline__ uint4 fd_find_pick(const uint4& f0, const uint4& f1, const uint4& f2, uint32_t h0, uint32_t chk, uint32_t& at) {
  const bool m0 = (f0.x & 255u) == chk, m1 = (f1.x & 255u) == chk, m2 = (f2.x & 255u) == chk;
  const uint32_t p0 = (f0.x >> 8) & 255u, p1 = (f1.x >> 8) & 255u, p2 = (f2.x >> 8) & 255u;
  const uint32_t repl = (p0 <= p1 && p0 <= p2) ? 0u : (p1 < p2 ? 16u : 32u);
  const bool hit = m0 || m1 || m2;
  const uint32_t sel = m0 ? 0u : m1 ? 16u : m2 ? 32u : repl;
  uint4 v = sel == 0u ? f0 : (sel == 16u ? f1 : f2);
  if (!hit) v = make_uint4(chk, 0u, 0u, 0u);
  at = h0 ^ sel;
  return v;
}

#define FD_NORMALISE()                                                    \
  while ((X.high ^ X.low) < 0x1000000u) {                                 \
    X.high = X.high << 8 | 255u; X.low <<= 8; X.low += (X.low == 0);      \
    X.curr = X.curr << 8 | X.in.take(X.status);                           \
  }

// per-lane constants of the kernel
struct FdLane {
  bool icm, isse, hashed, cm, cons;   // what this lane OWNS (lanes 0..7 only)
  bool hashedC;                       // component (lane & 7) is hashed (helper lanes mirror it)
  int consp;                          // CONS: its prediction (Predictor.cs:96-98)
  const uint32_t* mapr;               // ICM / ISSE map in the shared slice: 4-byte entries {p} (ICM) / 8-byte entries {weight, bias} (ISSE); a readable dummy elsewhere
  uint32_t* mapw;                     // the same for stores; lanes without a map write their own scratch words
  uint32_t rmul;                      // words per entry when reading: 1 (ICM) or 2
  uint32_t wmul;                      // ... when storing: 1, 2, or 0 for lanes without a map
  uint32_t ymul;                      // word offset of the second value of an entry: 1 (ISSE, scratch) or 0 (ICM: there is none)
  uint8_t* rows0; uint8_t* rows1;     // the block's row buffers: 32 x 16 bytes each
  int4* slots;                        // 32 x {x0, y0, x1, y1}
  uint8_t* cmline;                    // CM: this lane's 16-entry line in shared memory
  int lm[8];                          // lm[k] = lane == k ? -1 : 0
  // MATCH tables (warp-uniform)
  uint32_t* mtab; uint8_t* mbuf; uint32_t mmask, mmask2;
};

// map entry of bit history bh: {p, p} for an ICM, {weight, bias} for an ISSE
__device__ __forceinline__ int2 fd_map_read(const FdLane& L, uint32_t bh) {
  const uint32_t* q = L.mapr + bh * L.rmul;
  return make_int2((int)q[0], (int)q[L.ymul]);
}

// slot of this lane for one outcome: what the other lanes need to evaluate the component
__device__ __forceinline__ int2 fd_slot(const FdLane& L, int x, int y, int plead) {
  return L.isse ? make_int2(x, y << 6) : make_int2(plead, 0);
}

// ------------------------------------------------------------------------------------------
// Start of a byte: c8 == 1, r.h = contexts after HCOMP.  Everything here waits for HBM once.
// ------------------------------------------------------------------------------------------
template <class FD>
__device__ __forceinline__ void fd_begin_byte(const Shared& S, FdCtx<FD>& X, const FdLane& L, LaneRegs& r, int lane) {
  // first-nibble rows, looked up directly (helper lanes repeat their component's look-up: same lines, no extra traffic)
  uint4 a0, a1, a2; uint32_t h0, chk, at = 0;
  fd_find_issue(r, r.h + 16u, a0, a1, a2, h0, chk);
  FD::mix_new_byte(S, X.W, r, lane);
  uint32_t cmbase = 0;
  uint4 c0, c1, c2, c3;
  if (FD::M_CM) {
    cmbase = ((r.h ^ 1u) & r.mask) & ~15u;           // hmap4 == 1..15 in the first nibble: one 64-byte line (Predictor.cs:263-266)
    const uint32_t* q = reinterpret_cast<const uint32_t*>(r.tab) + cmbase;
    c0 = fd_ldg128(q); c1 = fd_ldg128(q + 4); c2 = fd_ldg128(q + 8); c3 = fd_ldg128(q + 12);
  }
  if (FD::ML >= 0) {
    X.mh = X.W.H[FD::ML & X.W.hmask];
    const uint32_t slot = X.mh & L.mmask;
    const uint32_t v = (uint32_t)fd_ldg32(L.mtab + slot);
    X.idxv = slot == X.lastslot ? X.lastpos : v;     // the slot written at the end of the last byte
    const int d2 = (int)S.dt2k[X.ma & 255u];
    X.mp0 = S.stretch[d2 & 32767];
    X.mp1 = S.stretch[(-d2) & 32767];
  }
  const uint4 v = fd_find_pick(a0, a1, a2, h0, chk, at);
  uint8_t* slot0 = L.rows0 + lane * 16;
  *reinterpret_cast<uint4*>(slot0) = v;                      // (lanes without a hashed component keep garbage in their own slot)
  X.rowp = slot0; X.at = at;
  int plead = L.cons ? L.consp : 0;
  if (FD::M_ICM | FD::M_ISSE) {
    const uint32_t bh = (v.x >> 8) & 255u;
    const int2 e = fd_map_read(L, bh);
    if (FD::M_CM == 0 || L.hashed) { r.cxt = bh; r.t0 = e.x; r.t1 = e.y; }
    if (L.icm) plead = S.stretch[((uint32_t)r.t0 >> 8) & 32767u];
  }
  if (FD::M_CM) {
    if (L.cm) {
      uint4* q = reinterpret_cast<uint4*>(L.cmline);
      q[0] = c0; q[1] = c1; q[2] = c2; q[3] = c3;
      r.c = cmbase;
      r.cxt = (r.h ^ 1u) & r.mask & 15u;
      r.t0 = (int)reinterpret_cast<const uint32_t*>(L.cmline)[r.cxt];
      plead = S.stretch[(uint32_t)r.t0 >> 17];
    }
  }
  const int2 s = fd_slot(L, r.t0, r.t1, plead);
  L.slots[lane] = make_int4(s.x, s.y, s.x, s.y);
  __syncwarp();
#pragma unroll
  for (int q = 0; q < FD::N; ++q)
    if ((FD::M_SLOT >> q) & 1u) { const int4 t = L.slots[q]; X.cw[q] = make_int2(t.x, t.y); }
  X.yprev = 0;
}

// ------------------------------------------------------------------------------------------
// One bit.  K = bit index inside the byte (compile time).
// ------------------------------------------------------------------------------------------
template <class FD, int K>
__device__ __forceinline__ void fd_bit(const Shared& S, FdCtx<FD>& X, const FdLane& L, LaneRegs& r, int lane) {
  constexpr bool NIB_END = K == 3, BYTE_END = K == 7;
  const uint32_t gl = (uint32_t)lane & 7u, grp = (uint32_t)lane >> 3;

  // ---- A: what bit K+1 will look at, under both outcomes of this bit ----
  uint32_t nb0 = 0, nb1 = 0;
  int2 e0 = make_int2(0, 0), e1 = make_int2(0, 0);
  uint32_t ce0 = 0, ce1 = 0; int cm0 = 0, cm1 = 0;
  if (!BYTE_END) {
    if (FD::M_ICM | FD::M_ISSE) {
      if (!NIB_END) {
        const uint32_t i0 = ((uint32_t)X.hmap4 * 2u) & 14u;
        const uint32_t v16 = *reinterpret_cast<const uint16_t*>(X.rowp + i0);
        nb0 = v16 & 255u; nb1 = v16 >> 8;
      } else {
        // first slot of the next nibble's row: one of the helper lanes' candidates, group 2*y2 + y3
        const uint32_t y2 = (uint32_t)X.c8 & 1u;
        const uint8_t* base = L.rows1 + (gl + 16u * y2) * 16u;
        nb0 = base[1]; nb1 = base[8 * 16 + 1];
      }
      e0 = fd_map_read(L, nb0);
      e1 = fd_map_read(L, nb1);
    }
    if (FD::M_CM) {
      if (!NIB_END) {
        // hmap4 of the next bit (Predictor.cs:463-474) differs from this bit's in its low 4 bits only: same line
        const uint32_t hm0 = ((uint32_t)X.hmap4 & 0x1f0u) | ((((uint32_t)X.hmap4 & 0xfu) * 2u) & 0xfu);
        ce0 = ((r.h ^ hm0) & r.mask) & 15u; ce1 = ((r.h ^ (hm0 | 1u)) & r.mask) & 15u;
        cm0 = (int)reinterpret_cast<const uint32_t*>(L.cmline)[ce0];
        cm1 = (int)reinterpret_cast<const uint32_t*>(L.cmline)[ce1];
      }
    }
  }

  // ---- B: on the path: every component of this bit, redundantly in all lanes ----
  int pj = 0, pin[FD::NMIX > 0 ? FD::NMIX : 1], pm[FD::NMIX > 0 ? FD::NMIX : 1];
  int vmatch = 0, ek = 0;
  if (FD::ML >= 0) {
    ek = (int)((X.mbyte >> (7 - K)) & 1u);
    vmatch = X.ma ? (ek ? X.mp1 : X.mp0) : 0;
  }
  const int pf = FD::eval(S, r, X.cw, L.lm, lane, vmatch, pj, pin, pm);     // sets r.p = this lane's own prediction
  const uint32_t prf = (uint32_t)S.squash[pf + 2048] * 2u + 1u;

  // ---- C: this bit's map entry trained under both outcomes (Predictor.cs:375-381, 440-449; 365-373 for CM) ----
  int t0x = 0, t0y = 0, t1x = 0, t1y = 0;
  uint32_t nx0 = 0, nx1 = 0;
  if (FD::M_ICM | FD::M_ISSE) {
    const int sq = (int)S.squash[r.p + 2048];
    const int er0 = 0 - sq, er1 = 32767 - sq;
    t0x = clamp512k(r.t0 + ((er0 * pj + (1 << 12)) >> 13));
    t1x = clamp512k(r.t0 + ((er1 * pj + (1 << 12)) >> 13));
    if (FD::M_ICM) {
      const uint32_t pn = (uint32_t)r.t0;
      const int i0 = (int)(pn + (uint32_t)(((int)(0 - (int)(pn >> 8))) >> 2));
      const int i1 = (int)(pn + (uint32_t)(((int)(32767 - (int)(pn >> 8))) >> 2));
      t0x = L.icm ? i0 : t0x; t1x = L.icm ? i1 : t1x;
    }
    t0y = clamp512k(r.t1 + ((er0 + 16) >> 5));
    t1y = clamp512k(r.t1 + ((er1 + 16) >> 5));
    const uint32_t n16 = *reinterpret_cast<const uint16_t*>(S.ns + (r.cxt & 255u) * 4u);
    nx0 = n16 & 255u; nx1 = n16 >> 8;
  }
  int ct0 = 0, ct1 = 0;
  if (FD::M_CM) {
    // restored train(), Predictor.cs:1031-1036
    const uint32_t pn = (uint32_t)r.t0, count = pn & 0x3ffu, lim = r.a2 * 4u;
    const int d = S.dt[count];
    ct0 = (int)(pn + (((uint32_t)(0 - (int)(pn >> 17)) * (uint32_t)d) & 0xFFFFFC00u) + (count < lim));
    ct1 = (int)(pn + (((uint32_t)(32767 - (int)(pn >> 17)) * (uint32_t)d) & 0xFFFFFC00u) + (count < lim));
  }

  // ---- D: candidates of bit K+1 with the training forwarded; publish both outcomes; read everybody's ----
  int c0x = 0, c0y = 0, c1x = 0, c1y = 0;
  if (!BYTE_END) {
    int pl0 = L.cons ? L.consp : 0, pl1 = pl0;
    if (FD::M_ICM | FD::M_ISSE) {
      const bool s0 = nb0 == r.cxt, s1 = nb1 == r.cxt;
      c0x = s0 ? t0x : e0.x; c0y = s0 ? t0y : e0.y;
      c1x = s1 ? t1x : e1.x; c1y = s1 ? t1y : e1.y;
      if (FD::M_ICM) {
        const int q0 = S.stretch[((uint32_t)c0x >> 8) & 32767u], q1 = S.stretch[((uint32_t)c1x >> 8) & 32767u];
        if (L.icm) { pl0 = q0; pl1 = q1; }
      }
    }
    if (FD::M_CM) {
      if (!NIB_END) {
        cm0 = ce0 == r.cxt ? ct0 : cm0; cm1 = ce1 == r.cxt ? ct1 : cm1;
        const int q0 = S.stretch[(uint32_t)cm0 >> 17], q1 = S.stretch[(uint32_t)cm1 >> 17];
        if (L.cm) { pl0 = q0; pl1 = q1; c0x = cm0; c1x = cm1; }
      }
    }
    const int2 s0 = fd_slot(L, c0x, c0y, pl0), s1 = fd_slot(L, c1x, c1y, pl1);
    L.slots[lane] = make_int4(s0.x, s0.y, s1.x, s1.y);
  }
  __syncwarp();
  int4 sl[FD::N];
  if (!BYTE_END) {
#pragma unroll
    for (int q = 0; q < FD::N; ++q)
      if ((FD::M_SLOT >> q) & 1u) sl[q] = L.slots[q];
  }

  // ---- E: the arithmetic decoder (Decoder.cs:136-158) ----
  if (X.curr < X.low || X.curr > X.high) X.status = BLK_CORRUPT;
  const uint32_t mid = X.low + (uint32_t)(((uint64_t)(X.high - X.low) * prf) >> 16);
  const int y = X.curr <= mid ? 1 : 0;
  if (y) X.high = mid; else X.low = mid + 1u;
  FD_NORMALISE();

  // ---- F: commit this bit, become bit K+1 ----
  if (FD::M_ICM | FD::M_ISSE) {
    // lanes without a map store into their own scratch (entry 0 of their row slot)
    {
      uint32_t* q = L.mapw + (r.cxt & 255u) * L.wmul;
      q[L.ymul] = (uint32_t)(y ? t1y : t0y);          // (an ICM entry has one word: the next store overwrites this one)
      q[0] = (uint32_t)(y ? t1x : t0x);
    }
    X.rowp[(uint32_t)X.hmap4 & 15u] = (uint8_t)(y ? nx1 : nx0);
  }
  if (FD::M_CM) {
    if (L.cm) reinterpret_cast<uint32_t*>(L.cmline)[r.cxt] = (uint32_t)(y ? ct1 : ct0);
  }
  const uint32_t c8new = (uint32_t)X.c8 * 2u + (uint32_t)y;
  FD::template mix_advance<K>(S, r, X.W, lane, y, X.yprev, pin, pm, c8new);
  if (FD::ML >= 0) { if (ek != y) X.ma = 0; }
  if (!BYTE_END) {
    if (FD::M_ICM | FD::M_ISSE) { if (FD::M_CM == 0 || L.hashed) { r.cxt = y ? nb1 : nb0; r.t0 = y ? c1x : c0x; r.t1 = y ? c1y : c0y; } }
    if (FD::M_CM) { if (L.cm && !NIB_END) { r.cxt = y ? ce1 : ce0; r.t0 = y ? c1x : c0x; } }
#pragma unroll
    for (int q = 0; q < FD::N; ++q)
      if ((FD::M_SLOT >> q) & 1u) X.cw[q] = y ? make_int2(sl[q].z, sl[q].w) : make_int2(sl[q].x, sl[q].y);
  }
  // Predictor.cs:463-474
  if (NIB_END) X.hmap4 = ((X.hmap4 & 0xf) << 5) | (y << 4) | 1;
  else X.hmap4 = (X.hmap4 & 0x1f0) | (((X.hmap4 & 0xf) * 2 + y) & 0xf);
  const uint32_t y2 = (uint32_t)X.c8 & 1u;
  X.c8 = (int)c8new;
  X.yprev = y;

  // ---- hash rows of the second nibble, a nibble ahead (helper lanes: one candidate per lane group) ----
  if (FD::M_ICM | FD::M_ISSE) {
#ifndef ZPQ_FDEC_NO_PF8
    // one bit known: pull the 8 rows the second nibble can use into L2.  (Measured, profiles/README.md: worth 1 - 2 % of the
    // kernel time at 12 blocks per SM -- 1110 .. 1125 ms with, 1138 ms without on 1776 x 200 kB -- for 2.35 kB of DRAM reads per
    // input byte, 38 % of the kernel's traffic; -DZPQ_FDEC_NO_PF8 turns it off.)
    if (K == 0) {
      const uint32_t cx = r.h + 16u * (c8new * 8u + grp * 2u);
      prefetch_l2(r.tab + ((cx * 16u) & r.mask));
      prefetch_l2(r.tab + (((cx + 16u) * 16u) & r.mask));
    }
#endif
    if (K == 1) {
      fd_find_issue(r, r.h + 16u * (c8new * 4u + grp), X.f0, X.f1, X.f2, X.fh0, X.fchk);
      const uint32_t oat = __shfl_sync(ZPQ_FULL, X.at, (int)gl);
      X.nhz = (((X.fh0 ^ oat) & ~63u) == 0u) ? 1u : 0u;     // same 64-byte group as the row in use: the copy in flight is stale
    }
    if (K == 2) {
      uint32_t at = 0;
      const uint4 v = fd_find_pick(X.f0, X.f1, X.f2, X.fh0, X.fchk, at);
      *reinterpret_cast<uint4*>(L.rows1 + lane * 16) = v;
      X.nat = at;
      __syncwarp();
    }
    if (NIB_END) {
      const int g = (int)(2u * y2 + (uint32_t)y);
      if (L.hashed) *reinterpret_cast<uint4*>(r.tab + X.at) = *reinterpret_cast<const uint4*>(X.rowp);
      uint8_t* nrow = L.rows1 + (gl + 8u * (uint32_t)g) * 16u;
      X.rowp = L.hashed ? nrow : X.rowp;                    // (the other lanes keep scribbling into their own slot)
      X.at = __shfl_sync(ZPQ_FULL, X.nat, (int)gl + 8 * g);
      const uint32_t hz = __shfl_sync(ZPQ_FULL, X.nhz, (int)gl + 8 * g);
      const bool fix = L.hashed && hz != 0u;
      if (__any_sync(ZPQ_FULL, fix)) {
        // rare: the chosen candidate shares its 64-byte group with the row just written back -- look it up again
        if (fix) {
          uint4 a0, a1, a2; uint32_t h0, chk, at = 0;
          fd_find_issue(r, r.h + 16u * c8new, a0, a1, a2, h0, chk);
          const uint4 v = fd_find_pick(a0, a1, a2, h0, chk, at);
          *reinterpret_cast<uint4*>(X.rowp) = v;
          X.at = at;
          const uint32_t bh = (v.x >> 8) & 255u;
          const int2 e = fd_map_read(L, bh);
          r.cxt = bh; r.t0 = e.x; r.t1 = e.y;
          const int plead = S.stretch[((uint32_t)r.t0 >> 8) & 32767u];
          const int2 s = fd_slot(L, r.t0, r.t1, plead);
          L.slots[lane] = make_int4(s.x, s.y, s.x, s.y);
        }
        __syncwarp();
        const uint32_t fm = __ballot_sync(ZPQ_FULL, fix);
#pragma unroll
        for (int q = 0; q < FD::N; ++q)
          if ((fm >> q) & 1u) { const int4 t = L.slots[q]; X.cw[q] = make_int2(t.x, t.y); }
      }
    }
  }
  if (FD::M_CM) {
    if (K == 1) {
      // two bits known: pull the 4 lines the second nibble can use into L2 (Predictor.cs:263-266: hmap4 = 256 + 16 * nibble + slot)
      const uint32_t hm = 256u + 16u * ((c8new * 4u + grp) & 15u);
      prefetch_l2(reinterpret_cast<const uint32_t*>(r.tab) + (((r.h ^ hm) & r.mask) & ~15u));
    }
    if (NIB_END) {
      // the 16 entries of the first nibble go back, the line of the second one comes in
      const uint32_t nbase = ((r.h ^ (uint32_t)X.hmap4) & r.mask) & ~15u;
      int plead = 0;
      if (L.cm) {
        uint4* q = reinterpret_cast<uint4*>(L.cmline);
        uint4* t = reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(r.tab) + r.c);
        t[0] = q[0]; t[1] = q[1]; t[2] = q[2]; t[3] = q[3];
        const uint4* n = reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(r.tab) + nbase);
        const uint4 c0 = n[0], c1 = n[1], c2 = n[2], c3 = n[3];
        q[0] = c0; q[1] = c1; q[2] = c2; q[3] = c3;
        r.c = nbase;
        r.cxt = ((r.h ^ (uint32_t)X.hmap4) & r.mask) & 15u;
        r.t0 = (int)reinterpret_cast<const uint32_t*>(L.cmline)[r.cxt];
        plead = S.stretch[(uint32_t)r.t0 >> 17];
        L.slots[lane] = make_int4(plead, 0, plead, 0);
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < FD::N; ++q)
        if ((FD::M_CM >> q) & 1u) { const int4 t = L.slots[q]; X.cw[q] = make_int2(t.x, t.y); }
    }
  }
  // ---- MATCH: what the end of the byte will look at (Predictor.cs:393-408), requested while the byte is decoded ----
  if (FD::ML >= 0) {
    if (K == 2) {
      X.be = L.mbuf[(X.idxv - 1u - (uint32_t)lane) & L.mmask2];
      X.ae = L.mbuf[(X.mpos - (uint32_t)lane) & L.mmask2];
      X.pbB = L.mbuf[X.idxv & L.mmask2];
      X.pbA = L.mbuf[(X.mpos + 1u - X.mb) & L.mmask2];
    }
  }
}

// ------------------------------------------------------------------------------------------
// End of a byte (c8 >= 256): MATCH update, rows back to their tables, HCOMP (Predictor.cs:382-411, 463-468)
// ------------------------------------------------------------------------------------------
template <class FD>
__device__ __forceinline__ void fd_end_byte(const Shared& S, FdCtx<FD>& X, const FdLane& L, LaneRegs& r, VM& vm, VMEnv& env, int lane, uint32_t c) {
  if (FD::ML >= 0) {
    const uint32_t mpos = X.mpos, npos = (mpos + 1u) & L.mmask2;
    if (lane == 0) L.mbuf[mpos] = (uint8_t)c;
    uint32_t pb;
    if (X.ma == 0) {
      X.mb = npos - X.idxv;
      if (X.mb & L.mmask2) {
        const uint32_t bpos = (X.idxv - 1u - (uint32_t)lane) & L.mmask2;
        const uint32_t a = lane == 0 ? c : lane <= 8 ? (uint32_t)(X.hist >> (8 * (lane - 1))) & 255u : X.ae;
        const uint32_t b = bpos == mpos ? c : X.be;
        const uint32_t neq = __ballot_sync(ZPQ_FULL, a != b);
        uint32_t n = neq ? (uint32_t)__ffs((int)neq) - 1u : 32u;
        if (n == 32u) {
          // more than 32 bytes agree: keep comparing, 32 byte pairs per round trip
          bool go = true;
          while (go && n < 255u) {
            const uint32_t ap = (npos - n - 1u - (uint32_t)lane) & L.mmask2, bp = (npos - n - X.mb - 1u - (uint32_t)lane) & L.mmask2;
            const uint32_t av = ap == mpos ? c : (uint32_t)L.mbuf[ap], bv = bp == mpos ? c : (uint32_t)L.mbuf[bp];
            const uint32_t q = __ballot_sync(ZPQ_FULL, av != bv);
            const uint32_t k = q ? (uint32_t)__ffs((int)q) - 1u : 32u;
            n += k; go = k == 32u;
          }
          n = n < 255u ? n : 255u;
        }
        X.ma = n;
      }
      pb = (X.idxv & L.mmask2) == mpos ? c : X.pbB;
    } else {
      X.ma += X.ma < 255u;
      pb = ((npos - X.mb) & L.mmask2) == mpos ? c : X.pbA;
    }
    const uint32_t slot = X.mh & L.mmask;
    if (lane == 0) L.mtab[slot] = npos;
    X.lastslot = slot; X.lastpos = npos;
    X.mpos = npos;
    if (X.ma) X.mbyte = pb;
    X.hist = X.hist << 8 | (uint64_t)c;
  }
  if (FD::M_ICM | FD::M_ISSE) {
    if (L.hashed) *reinterpret_cast<uint4*>(r.tab + X.at) = *reinterpret_cast<const uint4*>(X.rowp);
  }
  if (FD::M_CM) {
    if (L.cm) {
      const uint4* q = reinterpret_cast<const uint4*>(L.cmline);
      uint4* t = reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(r.tab) + r.c);
      t[0] = q[0]; t[1] = q[1]; t[2] = q[2]; t[3] = q[3];
    }
  }
  if (FD::hcomp(S, X.W, vm, env, c, lane)) X.status = BLK_ZPAQL;
  r.h = X.W.H[((uint32_t)lane & 7u) & X.W.hmask];
  X.c8 = 1; X.hmap4 = 1;
}

// ------------------------------------------------------------------------------------------
// Kernel body (Decoder.cs:32-68 driven as Decompresser.cs:121-153 does; the post-processor is a separate pass)
// ------------------------------------------------------------------------------------------
template <class FD>
__device__ __forceinline__ void fdec_body(const CodecParams& P, uint8_t* smem) {
  Shared S;
  stage_shared(P, smem, S);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gl = lane & 7;
  const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp;
  if (gw >= P.resident) return;
  Blk w;
  bind_block(P, smem, w, gw, warp);
  const Plan* plan = P.plan;
  // small H / M arrays live in the shared slice (the host plans this kernel only then): keep the pointers provably shared
  if (FD::H_SMEM) w.H = reinterpret_cast<uint32_t*>(smem + P.sm.slices + (uint32_t)warp * P.sm.slice_bytes + plan->smem_h);
  if (FD::M_SMEM) w.M = smem + P.sm.slices + (uint32_t)warp * P.sm.slice_bytes + plan->smem_m;
  LaneRegs r;
  lane_load(S, P, w, r, gl < FD::N ? gl : 0);
  FdLane L;
  {
    const int ctype = gl < FD::N ? r.type : (int)C_NONE;
    const bool owner = lane < 8;
    L.hashedC = ctype == C_ICM || ctype == C_ISSE;
    L.icm = owner && ctype == C_ICM; L.isse = owner && ctype == C_ISSE; L.hashed = L.icm || L.isse;
    L.cm = owner && ctype == C_CM; L.cons = owner && ctype == C_CONS;
    L.consp = ((int)r.a1 - 128) * 4;
    if (!L.hashedC && ctype != C_CM) { r.tab = w.arena; r.mask = 0; r.a1 = 0; }
    L.rows0 = w.slice + plan->smem_rows; L.rows1 = L.rows0 + 512;
    L.mapr = L.hashed ? reinterpret_cast<const uint32_t*>(w.slice + S.comp[gl].smem_cm)
                      : reinterpret_cast<const uint32_t*>(S.stretch);
    L.mapw = L.hashed ? reinterpret_cast<uint32_t*>(w.slice + S.comp[gl].smem_cm) : reinterpret_cast<uint32_t*>(L.rows0 + lane * 16);
    L.rmul = L.icm ? 1u : 2u;
    L.wmul = L.icm ? 1u : L.isse ? 2u : 0u;
    L.ymul = L.icm ? 0u : 1u;
    for (int k = 0; k < 8; ++k) { int m = lane == k ? -1 : 0; asm volatile("" : "+r"(m)); L.lm[k] = m; }   // opaque: kept in registers, not recomputed
    L.slots = reinterpret_cast<int4*>(w.slice + plan->smem_chain);
    L.cmline = w.slice + plan->smem_fd_cm + (uint32_t)gl * 64u;
    L.mtab = nullptr; L.mbuf = nullptr; L.mmask = L.mmask2 = 0;
    if (FD::ML >= 0) {
      const CompDesc& d = S.comp[FD::ML >= 0 ? FD::ML : 0];
      L.mtab = reinterpret_cast<uint32_t*>(w.arena + d.tab); L.mbuf = w.arena + d.tab2; L.mmask = d.mask; L.mmask2 = d.mask2;
    }
    if (!owner) r.type = C_NONE;
  }
  FdCtx<FD> X;
  VM vm; VMEnv env;

  for (;;) {
    uint32_t job = 0;
    if (lane == 0) job = atomicAdd(P.queue, 1u);
    job = __shfl_sync(ZPQ_FULL, job, 0);
    if (job >= P.njobs) break;
    const DecJob J = P.djobs[job];
    uint8_t* raw = P.out + J.out_off;
    uint64_t rpos = 0;

    // ---- Predictor.init + ZPAQL.inith ----
    init_block_state(plan, P.tab, w.arena, w.slice, lane);
    __syncwarp();
    r.cxt = r.c = r.h = 0; r.t0 = r.t1 = 0; r.p = 0;
    for (int k = 0; k < kMixRegs; ++k) { r.mw[k] = r.mn0[k] = r.mn1[k] = 0; r.moff[k] = 0; X.W.mixh[k] = 0; }
    X.W.arena = w.arena; X.W.H = w.H; X.W.hmask = w.hmask; X.W.c8 = 1; X.W.hmap4 = 1;
    X.low = 1; X.high = 0xFFFFFFFFu; X.curr = 0; X.status = BLK_OK;
    X.c8 = 1; X.hmap4 = 1; X.yprev = 0;
    X.ma = X.mb = X.mpos = X.mbyte = X.mh = X.idxv = 0; X.lastslot = 0xFFFFFFFFu; X.lastpos = 0;
    X.pbA = X.pbB = 0; X.mp0 = X.mp1 = 0; X.hist = 0; X.be = X.ae = 0;
    X.at = X.nat = X.nhz = 0; X.fh0 = X.fchk = 0;
    X.f0 = X.f1 = X.f2 = make_uint4(0, 0, 0, 0);
    X.rowp = L.rows0 + lane * 16;
    if (FD::ML >= 0 && lane == 0) L.mbuf[0] = 1;                    // Predictor.cs:118
    vm.b = vm.c = vm.d = vm.f = 0;
    env.code = S.hcomp; env.len = S.hcomp_len;
    env.H = w.H; env.hmask = w.hmask; env.M = w.M; env.mmask = w.mmask; env.R = w.R;
    env.out = nullptr; env.out_pos = 0; env.out_cap = 0;
    __syncwarp();
    fd_begin_byte<FD>(S, X, L, r, lane);

    for (uint32_t sg = 0; sg < J.seg_count && X.status == BLK_OK; ++sg) {
      const DecSeg seg = P.segs[J.seg_first + sg];
      X.in.open(P.in + seg.in_off, seg.in_len);
      if (X.curr == 0)                                              // segment initialisation, Decoder.cs:38-42
        for (int k = 0; k < 4; ++k) X.curr = X.curr << 8 | X.in.take(X.status);
      while (X.status == BLK_OK) {
        // decode(0): the end-of-segment flag in front of every byte (Decoder.cs:43-47)
        if (X.curr < X.low || X.curr > X.high) { X.status = BLK_CORRUPT; break; }
        const bool eos = X.curr <= X.low;
        if (eos) X.high = X.low; else X.low += 1u;
        FD_NORMALISE();
        if (eos) { if (X.curr != 0) X.status = BLK_CORRUPT; break; }
        fd_bit<FD, 0>(S, X, L, r, lane);
        fd_bit<FD, 1>(S, X, L, r, lane);
        fd_bit<FD, 2>(S, X, L, r, lane);
        fd_bit<FD, 3>(S, X, L, r, lane);
        fd_bit<FD, 4>(S, X, L, r, lane);
        fd_bit<FD, 5>(S, X, L, r, lane);
        fd_bit<FD, 6>(S, X, L, r, lane);
        fd_bit<FD, 7>(S, X, L, r, lane);
        const uint32_t c = (uint32_t)X.c8 - 256u;
        if (lane == 0 && rpos < J.out_cap) raw[rpos] = (uint8_t)c;
        ++rpos;
        fd_end_byte<FD>(S, X, L, r, vm, env, lane, c);
        fd_begin_byte<FD>(S, X, L, r, lane);
      }
      if (lane == 0) P.seg_end[J.seg_first + sg] = rpos;
    }
    if (X.status == BLK_OK && rpos > J.out_cap) X.status = BLK_OVERFLOW;
    __syncwarp();
    if (lane == 0) { P.results[job].out_len = rpos; P.results[job].status = X.status; }
  }
}

}  // namespace zpq
