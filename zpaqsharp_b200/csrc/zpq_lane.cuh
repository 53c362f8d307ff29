// zpq_lane.cuh -- lane-resident predictor: the hot kernels for models with <= 32 components.
// Included by zpq_kernels.cu after the shared helpers (Shared, zpaql_run, find helpers ...).
//
// Lane i of the block's warp OWNS component i: its descriptor, its table pointers, its context
// and its current stretched prediction live in that lane's registers for the whole block.
//
//   phase A   every lane computes the part of its prediction that depends on tables only
//             (CM / ICM / MATCH finish here; ISSE fetches its two weights, MIX2 its weight)
//   levels    for L = 1..maxlevel: inputs travel by SHFL from the producing lanes; lanes whose
//             component sits at level L finish.  A MIX at level L is evaluated by the whole
//             warp: lane j multiplies weight j by input j, REDUX adds     [Predictor.cs:245-350]
//   update    one pass, every lane trains its own component; MIX rows one weight per lane
//                                                                          [Predictor.cs:353-475]
//   nibble    ICM/ISSE hash rows (16 B) are cached in shared memory for the 4 bits they serve:
//             all lanes look their rows up together at the nibble boundary, so the DRAM misses
//             of the whole chain overlap; the row is written back once    [Predictor.cs:550-567]
#pragma once

namespace zpq {

struct LaneRegs {
  int type, level, srcj, srck;
  uint32_t a1, a2, a3, a4, a5;
  uint32_t mask, mask2;
  uint8_t* tab;
  uint8_t* tab2;
  uint32_t* cm;       // ICM/ISSE probability / weight map (shared or arena)
  uint8_t* row;       // this lane's 16-byte row cache in shared memory
  // dynamic
  uint32_t cxt, c, ma, mb, mpos, h;
  int p, t0, t1;
};

struct WarpCtx {
  uint8_t* arena;
  uint32_t* H; uint32_t hmask;
  int c8, hmap4;
};

__device__ __forceinline__ void lane_load(const Shared& S, const CodecParams& P, const Blk& w, LaneRegs& r, int lane) {
  r.type = C_NONE; r.level = 0; r.srcj = r.srck = 0;
  r.a1 = r.a2 = r.a3 = r.a4 = r.a5 = 0; r.mask = r.mask2 = 0;
  r.tab = r.tab2 = nullptr; r.cm = nullptr;
  r.row = w.slice + P.plan->smem_rows + lane * 16;
  if (lane < S.n) {
    const CompDesc& d = S.comp[lane];
    r.type = d.type; r.level = d.level;
    r.a1 = d.a[0]; r.a2 = d.a[1]; r.a3 = d.a[2]; r.a4 = d.a[3]; r.a5 = d.a[4];
    r.mask = d.mask; r.mask2 = d.mask2;
    r.tab = w.arena + d.tab; r.tab2 = w.arena + d.tab2;
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

// ---- phase A + levels -------------------------------------------------------------------------
__device__ __forceinline__ int lane_predict(const Shared& S, const WarpCtx& W, LaneRegs& r, int lane) {
  const int c8 = W.c8, hmap4 = W.hmap4;
  switch (r.type) {
    case C_CM:
      r.cxt = (r.h ^ hmap4) & r.mask;
      r.p = S.stretch[reinterpret_cast<const uint32_t*>(r.tab)[r.cxt] >> 17];
      break;
    case C_ICM:
      r.cxt = r.row[hmap4 & 15];
      r.p = S.stretch[r.cm[r.cxt] >> 8];
      break;
    case C_ISSE: {
      r.cxt = r.row[hmap4 & 15];
      const int2 wt = *reinterpret_cast<const int2*>(r.cm + r.cxt * 2);
      r.t0 = wt.x; r.t1 = wt.y;
      break;
    }
    case C_MATCH:
      if (r.ma == 0) r.p = 0;
      else {
        const uint32_t bit = (r.tab2[(r.mpos - r.mb) & r.mask2] >> (7 - r.cxt)) & 1;
        r.c = bit;
        r.p = S.stretch[(S.dt2k[r.ma] * (1 - 2 * (int)bit)) & 32767];
      }
      break;
    case C_MIX2:
      r.cxt = (r.h + (c8 & r.a5)) & r.mask;
      r.t0 = reinterpret_cast<const uint16_t*>(r.tab)[r.cxt];
      break;
    case C_SSE:
      r.t0 = (int)((r.h + c8) * 32);
      break;
    default: break;
  }
  for (int L = 1; L <= S.maxlevel; ++L) {
    const int pj = __shfl_sync(FULL, r.p, r.srcj);
    const int pk = __shfl_sync(FULL, r.p, r.srck);
    if (r.level == L) {
      switch (r.type) {
        case C_ISSE: r.p = clamp2k((r.t0 * pj + r.t1 * 64) >> 16); break;
        case C_AVG: r.p = (pj * (int)r.a3 + pk * (256 - (int)r.a3)) >> 8; break;
        case C_MIX2: r.p = (r.t0 * pj + (65536 - r.t0) * pk) >> 16; break;
        case C_SSE: {
          int pq = max(0, min(1983, pj + 992));
          const int wt = pq & 63;
          pq >>= 6;
          const uint32_t cx = (uint32_t)r.t0 + pq;
          const uint32_t* cm = reinterpret_cast<const uint32_t*>(r.tab);
          r.p = S.stretch[((cm[cx & r.mask] >> 10) * (64 - wt) + (cm[(cx + 1) & r.mask] >> 10) * wt) >> 13];
          r.cxt = (cx + (wt >> 5)) & r.mask;
          break;
        }
        default: break;
      }
    }
    for (int k = 0; k < S.nmix; ++k) {
      const MixDesc md = S.mix[k];
      if (md.level != L) continue;
      const uint32_t hm = __shfl_sync(FULL, r.h, md.lane);
      const uint32_t rowi = ((hm + (c8 & md.cmask)) & md.mask) * md.m;
      const int pin = __shfl_sync(FULL, r.p, md.j0 + lane);
      int prod = 0;
      if (lane < md.m) prod = (reinterpret_cast<const int*>(W.arena + md.tab)[rowi + lane] >> 8) * pin;
      const int acc = __reduce_add_sync(FULL, prod);
      if (lane == md.lane) { r.p = clamp2k(acc >> 8); r.cxt = rowi; }
    }
  }
  return S.squash[__shfl_sync(FULL, r.p, S.n - 1) + 2048];
}

// ---- update -----------------------------------------------------------------------------------
__device__ __forceinline__ void lane_update(const Shared& S, const WarpCtx& W, LaneRegs& r, int lane, int y) {
  const int pj = __shfl_sync(FULL, r.p, r.srcj);
  const int pk = __shfl_sync(FULL, r.p, r.srck);
  switch (r.type) {
    case C_CM:
      train(S, reinterpret_cast<uint32_t*>(r.tab) + r.cxt, r.a2 * 4u, y);
      break;
    case C_ICM: {
      uint8_t* slot = r.row + (W.hmap4 & 15);
      *slot = S.ns[r.cxt * 4 + y];
      const uint32_t pn = r.cm[r.cxt];
      r.cm[r.cxt] = pn + (uint32_t)(((int)(y * 32767 - (pn >> 8))) >> 2);
      break;
    }
    case C_ISSE: {
      const int err = y * 32767 - (int)S.squash[r.p + 2048];
      int2 wt;
      wt.x = clamp512k(r.t0 + ((err * pj + (1 << 12)) >> 13));
      wt.y = clamp512k(r.t1 + ((err + 16) >> 5));
      *reinterpret_cast<int2*>(r.cm + r.cxt * 2) = wt;
      r.row[W.hmap4 & 15] = S.ns[r.cxt * 4 + y];
      break;
    }
    case C_MATCH: {
      uint8_t* buf = r.tab2;
      if ((int)r.c != y) r.ma = 0;
      buf[r.mpos] = (uint8_t)(buf[r.mpos] * 2 + y);
      if (++r.cxt == 8) {
        r.cxt = 0;
        r.mpos = (r.mpos + 1) & r.mask2;
        uint32_t* idx = reinterpret_cast<uint32_t*>(r.tab) + (r.h & r.mask);
        if (r.ma == 0) {
          r.mb = r.mpos - *idx;
          if (r.mb & r.mask2)
            while (r.ma < 255 && buf[(r.mpos - r.ma - 1) & r.mask2] == buf[(r.mpos - r.ma - r.mb - 1) & r.mask2]) ++r.ma;
        } else r.ma += r.ma < 255;
        *idx = r.mpos;
      }
      break;
    }
    case C_MIX2: {
      const int err = ((y * 32767 - (int)S.squash[r.p + 2048]) * (int)r.a4) >> 5;
      int wt = r.t0 + ((err * (pj - pk) + (1 << 12)) >> 13);
      wt = max(0, min(65535, wt));
      reinterpret_cast<uint16_t*>(r.tab)[r.cxt] = (uint16_t)wt;
      break;
    }
    case C_SSE:
      train(S, reinterpret_cast<uint32_t*>(r.tab) + r.cxt, r.a4 * 4u, y);
      break;
    default: break;
  }
  for (int k = 0; k < S.nmix; ++k) {
    const MixDesc md = S.mix[k];
    const int pm = __shfl_sync(FULL, r.p, md.lane);
    const uint32_t rowi = __shfl_sync(FULL, r.cxt, md.lane);
    const int pin = __shfl_sync(FULL, r.p, md.j0 + lane);
    const int err = ((y * 32767 - (int)S.squash[pm + 2048]) * (int)md.rate) >> 4;
    if (lane < md.m) {
      int* wp = reinterpret_cast<int*>(W.arena + md.tab) + rowi + lane;
      *wp = clamp512k(*wp + ((err * pin + (1 << 12)) >> 13));
    }
  }
}

// Shift the coded bit into c8 / hmap4; at nibble boundaries write hash rows back and fetch the
// next ones; at byte boundaries run HCOMP first (Predictor.cs:463-474).
__device__ __forceinline__ uint32_t lane_advance(const Shared& S, WarpCtx& W, LaneRegs& r, VM& vm, VMEnv& env, int lane, int y) {
  uint32_t status = ZPQ_BLOCK_OK;
  int c8 = W.c8 * 2 + y;
  const bool hashed = (r.type == C_ICM || r.type == C_ISSE);
  if (c8 >= 256) {
    if (hashed) *reinterpret_cast<uint4*>(r.tab + r.c) = *reinterpret_cast<const uint4*>(r.row);
    int rc = 0;
    __syncwarp();
    if (lane == 0) rc = zpaql_run(vm, env, (uint32_t)(c8 - 256), 1u << 22);
    rc = __shfl_sync(FULL, rc, 0);
    __syncwarp();
    if (rc) status = ZPQ_BLOCK_ZPAQL;
    r.h = W.H[lane & W.hmask];
    W.hmap4 = 1;
    c8 = 1;
    if (hashed) lane_find(r, r.h + 16);
  } else if (c8 >= 16 && c8 < 32) {
    W.hmap4 = (W.hmap4 & 0xf) << 5 | y << 4 | 1;
    if (hashed) {
      *reinterpret_cast<uint4*>(r.tab + r.c) = *reinterpret_cast<const uint4*>(r.row);
      lane_find(r, r.h + 16 * c8);
    }
  } else W.hmap4 = (W.hmap4 & 0x1f0) | (((W.hmap4 & 0xf) * 2 + y) & 0xf);
  W.c8 = c8;
  return status;
}

__device__ __forceinline__ void lane_begin(const CodecParams& P, const Shared& S, Blk& w, WarpCtx& W, LaneRegs& r, VM& vm,
                                           VMEnv& env, int lane) {
  init_block_state(P.plan, P.tab, w.arena, w.slice, lane);
  __syncwarp();
  r.cxt = r.c = r.ma = r.mb = r.mpos = r.h = 0;
  r.t0 = r.t1 = 0;
  r.p = r.type == C_CONS ? ((int)r.a1 - 128) * 4 : 0;
  if (r.type == C_MATCH) r.tab2[0] = 1;                         // Predictor.cs:118
  W.arena = w.arena; W.H = w.H; W.hmask = w.hmask; W.c8 = 1; W.hmap4 = 1;
  vm.b = vm.c = vm.d = vm.f = 0;
  env.code = S.hcomp; env.len = S.hcomp_len;
  env.H = w.H; env.hmask = w.hmask; env.M = w.M; env.mmask = w.mmask; env.R = w.R;
  env.out = nullptr; env.out_pos = 0; env.out_cap = 0;
  if (r.type == C_ICM || r.type == C_ISSE) lane_find(r, 16);     // h = 0, c8 = 1
  __syncwarp();
}

// ------------------------------------------------------------------------------------------
// Encoder (Encoder.cs:39-103 as driven by Compressor.cs:156-248)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCtaThreads, 1) k_zpaq_encode_lanes(const CodecParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
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
    job = __shfl_sync(FULL, job, 0);
    if (job >= P.njobs) break;
    const EncJob J = P.ejobs[job];
    const uint8_t* in = P.in + J.in_off;
    uint8_t* out = P.out + J.out_off;
    const uint64_t total = (uint64_t)J.pre_len + J.in_len;
    uint64_t opos = 0;
    uint32_t status = ZPQ_BLOCK_OK;
    lane_begin(P, S, w, W, r, vm, env, lane);
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
        const uint32_t pr = (uint32_t)lane_predict(S, W, r, lane) * 2 + 1;
        const int y = (c >> i) & 1;
        const uint32_t mid = low + (uint32_t)(((uint64_t)(high - low) * pr) >> 16);
        if (y) high = mid; else low = mid + 1;
        ZPQ_NORMALISE();
        lane_update(S, W, r, lane, y);
        status |= lane_advance(S, W, r, vm, env, lane, y);
      }
      if (opos > J.out_cap || status) break;
    }
    high = low;  // encode(1, 0), Encoder.cs:46
    ZPQ_NORMALISE();
#undef ZPQ_NORMALISE
    if (opos > J.out_cap) status = ZPQ_BLOCK_OVERFLOW;
    __syncwarp();
    if (lane == 0) { P.results[job].out_len = opos; P.results[job].status = status; }
  }
}

// ------------------------------------------------------------------------------------------
// Decoder + post-processor (Decoder.cs:32-158, PostProcessor.cs:37-86)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCtaThreads, 1) k_zpaq_decode_lanes(const CodecParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
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
    job = __shfl_sync(FULL, job, 0);
    if (job >= P.njobs) break;
    const DecJob J = P.djobs[job];
    uint8_t* out = P.out + J.out_off;
    lane_begin(P, S, w, W, r, vm, env, lane);
    uint32_t status = ZPQ_BLOCK_OK;

    int pstate = 0;
    uint32_t psize = 0, ploaded = 0;
    uint8_t* pcode = w.arena + plan->off_pcode;
    VM pvm; pvm.b = pvm.c = pvm.d = pvm.f = 0;
    VMEnv penv;
    penv.code = pcode; penv.len = 0;
    penv.H = reinterpret_cast<uint32_t*>(w.arena + plan->off_ph); penv.hmask = (1u << plan->ph) - 1;
    penv.M = w.arena + plan->off_pm; penv.mmask = (uint32_t)((1ull << plan->pm) - 1);
    penv.R = reinterpret_cast<uint32_t*>(w.arena + plan->off_pr);
    penv.out = out; penv.out_pos = 0; penv.out_cap = J.out_cap;
    uint64_t opos = 0, consumed = 0;

    auto post = [&](int c) {
      switch (pstate) {
        case 0:
          if (c < 0 || c > 1) { status = ZPQ_BLOCK_POSTPROC; return; }
          pstate = c + 1;
          break;
        case 1:
          if (c >= 0) { if (lane == 0 && opos < J.out_cap) out[opos] = (uint8_t)c; ++opos; }
          break;
        case 2:
          if (c < 0) { status = ZPQ_BLOCK_POSTPROC; return; }
          psize = c; pstate = 3;
          break;
        case 3:
          if (c < 0) { status = ZPQ_BLOCK_POSTPROC; return; }
          psize += c * 256;
          if (psize < 1) { status = ZPQ_BLOCK_POSTPROC; return; }
          ploaded = 0; pstate = 4;
          break;
        case 4:
          if (c < 0) { status = ZPQ_BLOCK_POSTPROC; return; }
          if (lane == 0) pcode[ploaded] = (uint8_t)c;
          if (++ploaded == psize) {
            if (lane == 0) { pcode[psize] = 0; pcode[psize + 1] = 0; pcode[psize + 2] = 0; }
            penv.len = (int)psize;
            pstate = 5;
          }
          break;
        default: {
          int rc = 0;
          __syncwarp();
          if (lane == 0) {
            penv.out_pos = opos;
            rc = zpaql_run(pvm, penv, c < 0 ? 0xFFFFFFFFu : (uint32_t)c, 65536 + 512 * (consumed + J.out_cap));
          }
          rc = __shfl_sync(FULL, rc, 0);
          opos = __shfl_sync(FULL, (unsigned long long)penv.out_pos, 0);
          if (rc) status = ZPQ_BLOCK_ZPAQL;
        }
      }
    };

    for (uint32_t sg = 0; sg < J.seg_count && status == ZPQ_BLOCK_OK; ++sg) {
      const DecSeg seg = P.segs[J.seg_first + sg];
      const uint8_t* in = P.in + seg.in_off;
      uint64_t ipos = 0;
      auto get = [&]() -> uint32_t {
        if (ipos < seg.in_len) return in[ipos++];
        status = ZPQ_BLOCK_CORRUPT;
        return 0;
      };
      uint32_t low = 1, high = 0xFFFFFFFFu, curr = 0;
      for (int k = 0; k < 4; ++k) curr = curr << 8 | get();
#define ZPQ_DECODE(pr, y)                                                         \
      {                                                                           \
        if (curr < low || curr > high) status = ZPQ_BLOCK_CORRUPT;                \
        const uint32_t mid = low + (uint32_t)(((uint64_t)(high - low) * (pr)) >> 16); \
        if (curr <= mid) { y = 1; high = mid; } else { y = 0; low = mid + 1; }     \
        while ((high ^ low) < 0x1000000u) {                                       \
          high = high << 8 | 255; low <<= 8; low += (low == 0);                   \
          curr = curr << 8 | get();                                               \
        }                                                                         \
      }
      while (status == ZPQ_BLOCK_OK) {
        int eos;
        ZPQ_DECODE(0u, eos);
        if (status) break;
        if (eos) {
          if (curr != 0) status = ZPQ_BLOCK_CORRUPT;
          else post(-1);
          break;
        }
        int c = 1;
        while (c < 256) {
          const uint32_t pr = (uint32_t)lane_predict(S, W, r, lane) * 2 + 1;
          int y;
          ZPQ_DECODE(pr, y);
          c += c + y;
          lane_update(S, W, r, lane, y);
          status |= lane_advance(S, W, r, vm, env, lane, y);
        }
        if (status) break;
        post(c - 256);
        ++consumed;
      }
#undef ZPQ_DECODE
    }
    if (status == ZPQ_BLOCK_OK && opos > J.out_cap) status = ZPQ_BLOCK_OVERFLOW;
    __syncwarp();
    if (lane == 0) { P.results[job].out_len = opos; P.results[job].status = status; }
  }
}

}  // namespace zpq
