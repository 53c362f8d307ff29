// zpq_kernels.cu -- sm_100a kernels of the ZPAQ block codec.
//
// Execution model: ZPAQ archive blocks are independent (LICENSE:44-46), so every block is owned
// by ONE WARP for its whole life.  A coding kernel is launched once per batch with as many
// resident warps as per-block state arenas fit in HBM; each warp pulls block indices from an
// atomic queue, (re)initialises its arena, and codes the block bit by bit:
//
//   predict  components of equal dependency level are evaluated across lanes (lane l owns the
//            l-th component of the level); MIX dot products are lane-parallel multiplies reduced
//            with REDUX (__reduce_add_sync)                      [Predictor.cs:245-350]
//   code     32-bit carry-less arithmetic coder, kept redundantly in registers of all lanes;
//            lane 0 stores / all lanes broadcast-load stream bytes  [Encoder.cs:87-103, Decoder.cs:136-158]
//   update   every lane trains the component it predicted; MIX weight rows are updated one
//            weight per lane                                      [Predictor.cs:353-475]
//   per byte the ZPAQL HCOMP program recomputes the context hashes  [ZPAQL.cs:1028-1265]
//
// squash/stretch/dt/state tables (78 KB) and the model descriptors live in shared memory once
// per CTA; per-block small tables (ICM/ISSE probability maps, H[], predictions) live in the
// block's shared-memory slice; the large hashed tables live in the block's HBM arena.
//
// All arithmetic is 32-bit integer; results are bit-identical to the reference semantics.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/zpaqb200.h"
#include "zpq_device.h"
#include "zpq_devcore.cuh"

namespace zpq {

#define FULL ZPQ_FULL

// ------------------------------------------------------------------------------------------
// Hash-row lookup for ICM / ISSE (Predictor.cs:550-567).  Rows are 16 bytes:
// [check, 15 bit-history slots]; the three probed rows share one 64-byte line.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t find_row(uint8_t* ht, uint32_t lenm16, int sizebits, uint32_t cxt) {
  const uint32_t chk = (cxt >> sizebits) & 255;
  const uint32_t h0 = (cxt * 16) & lenm16;
  const uint32_t h1 = h0 ^ 16, h2 = h0 ^ 32;
  const uint32_t r0 = *reinterpret_cast<const uint16_t*>(ht + h0);
  if ((r0 & 255) == chk) return h0;
  const uint32_t r1 = *reinterpret_cast<const uint16_t*>(ht + h1);
  if ((r1 & 255) == chk) return h1;
  const uint32_t r2 = *reinterpret_cast<const uint16_t*>(ht + h2);
  if ((r2 & 255) == chk) return h2;
  const uint32_t p0 = r0 >> 8, p1 = r1 >> 8, p2 = r2 >> 8;
  const uint32_t r = (p0 <= p1 && p0 <= p2) ? h0 : (p1 < p2 ? h1 : h2);
  *reinterpret_cast<uint4*>(ht + r) = make_uint4(chk, 0, 0, 0);
  return r;
}

__device__ __forceinline__ uint32_t* cm_small(const CompDesc& d, const Blk& w) {
  return d.smem_cm != kNoSmem ? reinterpret_cast<uint32_t*>(w.slice + d.smem_cm)
                              : reinterpret_cast<uint32_t*>(w.arena + d.tab2);
}

// ------------------------------------------------------------------------------------------
// One lane-owned component: prediction (Predictor.cs:255-347)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void predict_one(const Shared& S, Blk& w, int i) {
  const CompDesc& d = S.comp[i];
  uint32_t* st = w.st + i * 5;
  const int c8 = w.c8, hmap4 = w.hmap4;
  int pr;
  switch (d.type) {
    case C_CM: {
      const uint32_t cxt = (w.H[i & w.hmask] ^ hmap4) & d.mask;
      st[0] = cxt;
      pr = S.stretch[reinterpret_cast<const uint32_t*>(w.arena + d.tab)[cxt] >> 17];
      break;
    }
    case C_ICM: {
      uint8_t* ht = w.arena + d.tab;
      if (c8 == 1 || (c8 & 0xf0) == 16) st[1] = find_row(ht, d.mask, d.a[0] + 2, w.H[i & w.hmask] + 16 * c8);
      const uint32_t bh = ht[st[1] + (hmap4 & 15)];
      st[0] = bh;
      pr = S.stretch[cm_small(d, w)[bh * 2] >> 8];
      break;
    }
    case C_MATCH: {
      if (st[2] == 0) pr = 0;
      else {
        const uint8_t* buf = w.arena + d.tab2;
        const uint32_t bit = (buf[(st[4] - st[3]) & d.mask2] >> (7 - st[0])) & 1;
        st[1] = bit;
        pr = S.stretch[(S.dt2k[st[2]] * (1 - 2 * (int)bit)) & 32767];
      }
      break;
    }
    case C_AVG:
      pr = (w.p[d.a[0]] * d.a[2] + w.p[d.a[1]] * (256 - d.a[2])) >> 8;
      break;
    case C_MIX2: {
      const uint32_t cxt = (w.H[i & w.hmask] + (c8 & d.a[4])) & d.mask;
      st[0] = cxt;
      const int wt = reinterpret_cast<const uint16_t*>(w.arena + d.tab)[cxt];
      pr = (wt * w.p[d.a[1]] + (65536 - wt) * w.p[d.a[2]]) >> 16;
      break;
    }
    case C_ISSE: {
      uint8_t* ht = w.arena + d.tab;
      if (c8 == 1 || (c8 & 0xf0) == 16) st[1] = find_row(ht, d.mask, d.a[0] + 2, w.H[i & w.hmask] + 16 * c8);
      const uint32_t bh = ht[st[1] + (hmap4 & 15)];
      st[0] = bh;
      const int* wt = reinterpret_cast<const int*>(cm_small(d, w)) + bh * 2;
      pr = clamp2k((wt[0] * w.p[d.a[1]] + wt[1] * 64) >> 16);
      break;
    }
    case C_SSE: {
      uint32_t cxt = (w.H[i & w.hmask] + c8) * 32;
      int pq = w.p[d.a[1]] + 992;
      pq = max(0, min(1983, pq));
      const int wt = pq & 63;
      pq >>= 6;
      cxt += pq;
      const uint32_t* cm = reinterpret_cast<const uint32_t*>(w.arena + d.tab);
      pr = S.stretch[((cm[cxt & d.mask] >> 10) * (64 - wt) + (cm[(cxt + 1) & d.mask] >> 10) * wt) >> 13];
      st[0] = (cxt + (wt >> 5)) & d.mask;
      break;
    }
    default: return;
  }
  w.p[i] = pr;
}

// One lane-owned component: update with coded bit y (Predictor.cs:365-459)
__device__ __forceinline__ void update_one(const Shared& S, Blk& w, int i, int y) {
  const CompDesc& d = S.comp[i];
  uint32_t* st = w.st + i * 5;
  switch (d.type) {
    case C_CM:
      train(S, reinterpret_cast<uint32_t*>(w.arena + d.tab) + st[0], d.a[1] * 4u, y);
      break;
    case C_ICM: {
      uint8_t* slot = w.arena + d.tab + st[1] + (w.hmap4 & 15);
      *slot = S.ns[*slot * 4 + y];
      uint32_t* cm = cm_small(d, w) + st[0] * 2;
      const uint32_t pn = *cm;
      *cm = pn + (uint32_t)(((int)(y * 32767 - (pn >> 8))) >> 2);
      break;
    }
    case C_MATCH: {
      uint8_t* buf = w.arena + d.tab2;
      uint32_t len = st[2], pos = st[4];
      if ((int)st[1] != y) len = 0;
      buf[pos] = (uint8_t)(buf[pos] * 2 + y);
      if (++st[0] == 8) {
        st[0] = 0;
        pos = (pos + 1) & d.mask2;
        uint32_t* idx = reinterpret_cast<uint32_t*>(w.arena + d.tab) + (w.H[i & w.hmask] & d.mask);
        if (len == 0) {
          const uint32_t off = pos - *idx;
          st[3] = off;
          if (off & d.mask2)
            while (len < 255 && buf[(pos - len - 1) & d.mask2] == buf[(pos - len - off - 1) & d.mask2]) ++len;
        } else len += len < 255;
        *idx = pos;
        st[4] = pos;
      }
      st[2] = len;
      break;
    }
    case C_MIX2: {
      const int err = ((y * 32767 - (int)S.squash[w.p[i] + 2048]) * d.a[3]) >> 5;
      uint16_t* a16 = reinterpret_cast<uint16_t*>(w.arena + d.tab) + st[0];
      int wt = *a16;
      wt += (err * (w.p[d.a[1]] - w.p[d.a[2]]) + (1 << 12)) >> 13;
      wt = max(0, min(65535, wt));
      *a16 = (uint16_t)wt;
      break;
    }
    case C_ISSE: {
      const int err = y * 32767 - (int)S.squash[w.p[i] + 2048];
      int* wt = reinterpret_cast<int*>(cm_small(d, w)) + st[0] * 2;
      wt[0] = clamp512k(wt[0] + ((err * w.p[d.a[1]] + (1 << 12)) >> 13));
      wt[1] = clamp512k(wt[1] + ((err + 16) >> 5));
      (w.arena + d.tab)[st[1] + (w.hmap4 & 15)] = S.ns[st[0] * 4 + y];
      break;
    }
    case C_SSE:
      train(S, reinterpret_cast<uint32_t*>(w.arena + d.tab) + st[0], d.a[3] * 4u, y);
      break;
    default: break;
  }
}

// Cooperative MIX (Predictor.cs:302-316 / 427-439): one weight per lane, REDUX for the sum.
__device__ __forceinline__ void mix_predict(const Shared& S, Blk& w, int i, int lane) {
  const CompDesc& d = S.comp[i];
  const int m = d.a[2], j0 = d.a[1];
  const uint32_t row = ((w.H[i & w.hmask] + (w.c8 & d.a[4])) & d.mask) * m;
  const int* wt = reinterpret_cast<const int*>(w.arena + d.tab) + row;
  int acc = 0;
  for (int j = lane; j < m; j += 32) acc += (wt[j] >> 8) * w.p[j0 + j];
  acc = __reduce_add_sync(FULL, acc);
  if (lane == 0) { w.p[i] = clamp2k(acc >> 8); w.st[i * 5] = row; }
}
__device__ __forceinline__ void mix_update(const Shared& S, Blk& w, int i, int lane, int y) {
  const CompDesc& d = S.comp[i];
  const int m = d.a[2], j0 = d.a[1];
  const int err = ((y * 32767 - (int)S.squash[w.p[i] + 2048]) * d.a[3]) >> 4;
  int* wt = reinterpret_cast<int*>(w.arena + d.tab) + w.st[i * 5];
  for (int j = lane; j < m; j += 32) wt[j] = clamp512k(wt[j] + ((err * w.p[j0 + j] + (1 << 12)) >> 13));
}

// Probability (0..32767) that the next bit is 1.
__device__ __forceinline__ int predict_bit(const Shared& S, Blk& w, int lane) {
  for (int s = 0; s < S.nsteps; ++s) {
    const Step stp = S.steps[s];
    if (stp.coop) mix_predict(S, w, S.order[stp.first], lane);
    else if (lane < stp.count) predict_one(S, w, S.order[stp.first + lane]);
    __syncwarp();
  }
  return S.squash[w.p[S.n - 1] + 2048];
}

// Train on bit y, then shift it into the partial-byte contexts; at a byte boundary run HCOMP.
__device__ __forceinline__ void update_bit(const Shared& S, Blk& w, VM& vm, VMEnv& env, int lane, int y) {
  for (int s = 0; s < S.nsteps; ++s) {
    const Step stp = S.steps[s];
    if (stp.coop) mix_update(S, w, S.order[stp.first], lane, y);
    else if (lane < stp.count) update_one(S, w, S.order[stp.first + lane], y);
  }
  __syncwarp();
  int c8 = w.c8 * 2 + y;
  if (c8 >= 256) {
    int rc = 0;
    if (lane == 0) rc = zpaql_run(vm, env, (uint32_t)(c8 - 256), 1u << 22);
    rc = __shfl_sync(FULL, rc, 0);
    if (rc) w.status = ZPQ_BLOCK_ZPAQL;
    w.hmap4 = 1;
    c8 = 1;
  } else if (c8 >= 16 && c8 < 32) w.hmap4 = (w.hmap4 & 0xf) << 5 | y << 4 | 1;
  else w.hmap4 = (w.hmap4 & 0x1f0) | (((w.hmap4 & 0xf) * 2 + y) & 0xf);
  w.c8 = c8;
  __syncwarp();
}

// Reset everything a new block needs (Predictor.init + ZPAQL.inith).
__device__ void begin_block(const CodecParams& P, const Shared& S, Blk& w, VM& vm, VMEnv& env, int lane) {
  const Plan* plan = P.plan;
  init_block_state(plan, P.tab, w.arena, w.slice, lane);
  for (int i = lane; i < S.n; i += 32) {
    w.p[i] = S.comp[i].type == C_CONS ? (S.comp[i].a[0] - 128) * 4 : 0;
    for (int k = 0; k < 5; ++k) w.st[i * 5 + k] = 0;
  }
  __syncwarp();
  if (lane == 0)
    for (int i = 0; i < S.n; ++i)
      if (S.comp[i].type == C_MATCH) (w.arena + S.comp[i].tab2)[0] = 1;   // Predictor.cs:118
  w.c8 = 1; w.hmap4 = 1; w.status = ZPQ_BLOCK_OK;
  vm.b = vm.c = vm.d = vm.f = 0;
  env.code = S.hcomp; env.len = S.hcomp_len;
  env.H = w.H; env.hmask = w.hmask; env.M = w.M; env.mmask = w.mmask; env.R = w.R;
  env.out = nullptr; env.out_pos = 0; env.out_cap = 0;
  __syncwarp();
}

// ------------------------------------------------------------------------------------------
// Encoder kernel (Encoder.cs:39-103 driven as Compressor.cs:156-248 does)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCtaThreads, 1) k_zpaq_encode(const CodecParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  Shared S;
  stage_shared(P, smem, S);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp;
  if (gw >= P.resident) return;
  Blk w;
  bind_block(P, smem, w, gw, warp);
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
    if (J.in_len == 0xFFFFFFFFu) {   // the pre-processing stage overflowed its slot
      if (lane == 0) { P.results[job].out_len = 0; P.results[job].status = ZPQ_BLOCK_OVERFLOW; }
      continue;
    }

    if (S.n == 0) {
      // stored mode (Encoder.cs:58-72): [len32 BE][bytes] per 64 KB of the stream
      for (uint64_t base = 0; base < total; base += 65536) {
        const uint32_t len = (uint32_t)min((uint64_t)65536, total - base);
        if (opos + 4 + len > J.out_cap) { status = ZPQ_BLOCK_OVERFLOW; break; }
        if (lane < 4) out[opos + lane] = (uint8_t)(len >> (24 - 8 * lane));
        for (uint32_t k = lane; k < len; k += 32) {
          const uint64_t s = base + k;
          out[opos + 4 + k] = s < J.pre_len ? P.preamble[s] : in[s - J.pre_len];
        }
        opos += 4 + len;
      }
    } else {
      begin_block(P, S, w, vm, env, lane);
      uint32_t low = 1, high = 0xFFFFFFFFu;
      // shift out the leading bytes low and high agree on (Encoder.cs:95-102)
#define ZPQ_NORMALISE()                                                       \
      while ((high ^ low) < 0x1000000u) {                                     \
        if (lane == 0 && opos < J.out_cap) out[opos] = (uint8_t)(high >> 24); \
        ++opos;                                                               \
        high = high << 8 | 255; low <<= 8; low += (low == 0);                 \
      }
      for (uint64_t s = 0; s < total; ++s) {
        const int c = s < J.pre_len ? P.preamble[s] : in[s - J.pre_len];
        ++low;  // encode(0, 0): mid = low, y = 0 -> low = mid + 1   (Encoder.cs:49)
        ZPQ_NORMALISE();
        for (int i = 7; i >= 0; --i) {
          const uint32_t pr = (uint32_t)predict_bit(S, w, lane) * 2 + 1;
          const int y = (c >> i) & 1;
          const uint32_t mid = low + (uint32_t)(((uint64_t)(high - low) * pr) >> 16);
          if (y) high = mid; else low = mid + 1;
          ZPQ_NORMALISE();
          update_bit(S, w, vm, env, lane, y);
        }
        if (opos > J.out_cap || w.status) break;
      }
      high = low;  // encode(1, 0): end of stream (Encoder.cs:46)
      ZPQ_NORMALISE();
#undef ZPQ_NORMALISE
      status = w.status;
      if (opos > J.out_cap) status = ZPQ_BLOCK_OVERFLOW;
    }
    __syncwarp();
    if (lane == 0) { P.results[job].out_len = opos; P.results[job].status = status; }
  }
}

// ------------------------------------------------------------------------------------------
// Decoder kernel: Decoder.cs:32-158; the post-processor (PostProcessor.cs:37-86) is a separate pass over the raw stream
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCtaThreads, 1) k_zpaq_decode(const CodecParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  Shared S;
  stage_shared(P, smem, S);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp;
  if (gw >= P.resident) return;
  Blk w;
  bind_block(P, smem, w, gw, warp);
  const Plan* plan = P.plan;
  VM vm; VMEnv env;

  for (;;) {
    uint32_t job = 0;
    if (lane == 0) job = atomicAdd(P.queue, 1u);
    job = __shfl_sync(FULL, job, 0);
    if (job >= P.njobs) break;
    const DecJob J = P.djobs[job];
    uint8_t* raw = P.out + J.out_off;        // the model's byte stream: PCOMP preamble + transformed data; the post-processor is a separate pass
    uint64_t rpos = 0;
    begin_block(P, S, w, vm, env, lane);
    uint32_t status = ZPQ_BLOCK_OK;

    uint32_t low = 1, high = 0xFFFFFFFFu, curr = 0;     // Decoder.init: once per block (Decompresser.cs:128-134), not per segment
    for (uint32_t sg = 0; sg < J.seg_count && status == ZPQ_BLOCK_OK; ++sg) {
      const DecSeg seg = P.segs[J.seg_first + sg];
      const uint8_t* in = P.in + seg.in_off;
      uint64_t ipos = 0;
      auto get = [&]() -> uint32_t {
        if (ipos < seg.in_len) return in[ipos++];
        status = ZPQ_BLOCK_CORRUPT;  // "unexpected end of file"
        return 0;
      };
      if (S.n == 0) {
        // stored mode (Decoder.cs:56-66): chunks of a 4-byte big-endian length and that many bytes, copied by the 32 lanes
        for (;;) {
          uint32_t len = 0;
          for (int k = 0; k < 4; ++k) len = len << 8 | get();
          if (len == 0 || status) break;
          const uint64_t have = seg.in_len - ipos, take = len < have ? len : have;
          const uint64_t room = rpos < J.out_cap ? J.out_cap - rpos : 0, put = take < room ? take : room;
          for (uint64_t i = lane; i < put; i += 32) raw[rpos + i] = in[ipos + i];
          rpos += take; ipos += take;
          if (take < len) status = ZPQ_BLOCK_CORRUPT;  // "unexpected end of file"
        }
        if (lane == 0) P.seg_end[J.seg_first + sg] = rpos;
        continue;
      }
      if (curr == 0)                                      // segment initialisation, Decoder.cs:38-42
        for (int k = 0; k < 4; ++k) curr = curr << 8 | get();
      // decode one bit with P(1) = pr/65536 (Decoder.cs:136-158)
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
          if (curr != 0) status = ZPQ_BLOCK_CORRUPT;  // "decoding end of stream"
          break;
        }
        int c = 1;
        while (c < 256) {
          const uint32_t pr = (uint32_t)predict_bit(S, w, lane) * 2 + 1;
          int y;
          ZPQ_DECODE(pr, y);
          c += c + y;
          update_bit(S, w, vm, env, lane, y);
        }
        if (w.status) { status = w.status; break; }
        if (lane == 0 && rpos < J.out_cap) raw[rpos] = (uint8_t)(c - 256);
        ++rpos;
      }
#undef ZPQ_DECODE
      if (lane == 0) P.seg_end[J.seg_first + sg] = rpos;
    }
    if (status == ZPQ_BLOCK_OK && rpos > J.out_cap) status = ZPQ_BLOCK_OVERFLOW;
    __syncwarp();
    if (lane == 0) { P.results[job].out_len = rpos; P.results[job].status = status; }
  }
}

// ------------------------------------------------------------------------------------------
// SHA-1 (FIPS 180-4), one thread per byte range.  The reference hashes every input block
// (LibZPAQ.cs:143-155) with a SHA1 class it does not ship; any conforming SHA-1 is identical.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rol32(uint32_t x, int n) { return __funnelshift_l(x, x, n); }

__device__ void sha1_chunk(uint32_t h[5], const uint32_t wbe[16]) {
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = wbe[i];
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4];
#pragma unroll
  for (int i = 0; i < 80; ++i) {
    uint32_t x;
    if (i < 16) x = w[i];
    else { x = rol32(w[(i + 13) & 15] ^ w[(i + 8) & 15] ^ w[(i + 2) & 15] ^ w[i & 15], 1); w[i & 15] = x; }
    uint32_t f, k;
    if (i < 20) { f = (b & c) | (~b & d); k = 0x5A827999u; }
    else if (i < 40) { f = b ^ c ^ d; k = 0x6ED9EBA1u; }
    else if (i < 60) { f = (b & c) | (b & d) | (c & d); k = 0x8F1BBCDCu; }
    else { f = b ^ c ^ d; k = 0xCA62C1D6u; }
    const uint32_t t = rol32(a, 5) + f + e + k + x;
    e = d; d = c; c = rol32(b, 30); b = a; a = t;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e;
}

__global__ void k_sha1(const uint8_t* data, const uint64_t* off, const uint32_t* len, uint32_t nb, uint8_t* digests) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const uint8_t* p = data + off[i];
  const uint64_t n = len[i];
  uint32_t h[5] = {0x67452301u, 0xEFCDAB89u, 0x98BADCFEu, 0x10325476u, 0xC3D2E1F0u};
  uint32_t w[16];
  uint64_t pos = 0;
  if (((uintptr_t)p & 15u) == 0) {
    // block starts and slots are 16-byte aligned: a 64-byte chunk is four 16-byte loads, byte order swapped by PRMT
    for (; pos + 64 <= n; pos += 64) {
      const uint4* q = reinterpret_cast<const uint4*>(p + pos);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint4 v = q[k];
        w[4 * k] = __byte_perm(v.x, 0, 0x0123); w[4 * k + 1] = __byte_perm(v.y, 0, 0x0123);
        w[4 * k + 2] = __byte_perm(v.z, 0, 0x0123); w[4 * k + 3] = __byte_perm(v.w, 0, 0x0123);
      }
      sha1_chunk(h, w);
    }
  }
  for (; pos + 64 <= n; pos += 64) {
#pragma unroll
    for (int k = 0; k < 16; ++k)
      w[k] = (uint32_t)p[pos + 4 * k] << 24 | (uint32_t)p[pos + 4 * k + 1] << 16 | (uint32_t)p[pos + 4 * k + 2] << 8 | p[pos + 4 * k + 3];
    sha1_chunk(h, w);
  }
  uint8_t tail[128];
  const int rem = (int)(n - pos);
  for (int k = 0; k < rem; ++k) tail[k] = p[pos + k];
  tail[rem] = 0x80;
  const int padded = rem < 56 ? 64 : 128;
  for (int k = rem + 1; k < padded - 8; ++k) tail[k] = 0;
  const uint64_t bits = n * 8;
  for (int k = 0; k < 8; ++k) tail[padded - 1 - k] = (uint8_t)(bits >> (8 * k));
  for (int q = 0; q < padded; q += 64) {
    for (int k = 0; k < 16; ++k)
      w[k] = (uint32_t)tail[q + 4 * k] << 24 | (uint32_t)tail[q + 4 * k + 1] << 16 | (uint32_t)tail[q + 4 * k + 2] << 8 | tail[q + 4 * k + 3];
    sha1_chunk(h, w);
  }
  for (int k = 0; k < 5; ++k) {
    digests[20 * i + 4 * k] = h[k] >> 24; digests[20 * i + 4 * k + 1] = h[k] >> 16;
    digests[20 * i + 4 * k + 2] = h[k] >> 8; digests[20 * i + 4 * k + 3] = h[k];
  }
}

cudaError_t launch_sha1(const uint8_t* data, const uint64_t* off, const uint32_t* len, uint32_t nb, uint8_t* digests,
                        cudaStream_t s) {
  if (!nb) return cudaSuccess;
  // one thread per block is the unit (SHA-1 is a serial chain); 8 threads per CTA spread them over the SMs so that each
  // thread's loads do not queue behind 31 others on one load-store unit
  k_sha1<<<(nb + 7) / 8, 8, 0, s>>>(data, off, len, nb, digests);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// E8E9 (LibZPAQ.cs:372-384): descending scan; a hit rewrites bytes i+1..i+3, which later tests
// at lower i read, so each block is scanned by one thread.
// ------------------------------------------------------------------------------------------
__global__ void k_e8e9(uint8_t* data, const uint64_t* off, const uint32_t* len, uint32_t nb) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  uint8_t* buf = data + off[b];
  const int n = (int)len[b];
  for (int i = n - 5; i >= 0; --i) {
    if ((buf[i] & 254) == 0xe8 && ((buf[i + 4] + 1) & 254) == 0) {
      const uint32_t a = (buf[i + 1] | buf[i + 2] << 8 | buf[i + 3] << 16) + (uint32_t)i;
      buf[i + 1] = (uint8_t)a; buf[i + 2] = (uint8_t)(a >> 8); buf[i + 3] = (uint8_t)(a >> 16);
    }
  }
}
cudaError_t launch_e8e9(uint8_t* data, const uint64_t* off, const uint32_t* len, uint32_t nb, cudaStream_t s) {
  if (!nb) return cudaSuccess;
  k_e8e9<<<(nb + 31) / 32, 32, 0, s>>>(data, off, len, nb);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Frame assembly (Compressor.cs:27-43,109-113,133-146,235-246,297): one warp per block.
// ------------------------------------------------------------------------------------------
__global__ void k_finish(const FinishParams P) {
  const uint32_t b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= P.nb) return;
  uint8_t* slot = P.slots + P.slot_off[b];
  const uint32_t p0 = P.prefix_off[b], plen = P.prefix_off[b + 1] - p0;
  for (uint32_t k = lane; k < plen; k += 32) slot[k] = P.prefix[p0 + k];
  uint8_t* t = slot + plen + P.results[b].out_len;
  uint64_t tail;
  if (P.digests) {
    if (lane < 4) t[lane] = 0;
    if (lane == 4) t[4] = 253;
    if (lane < 20) t[5 + lane] = P.digests[20 * b + lane];
    if (lane == 20) t[25] = 255;
    tail = 26;
  } else {
    if (lane < 4) t[lane] = 0;
    if (lane == 4) t[4] = 254;
    if (lane == 5) t[5] = 255;
    tail = 6;
  }
  if (lane == 0) P.frame_len[b] = plen + P.results[b].out_len + tail;
}
cudaError_t launch_finish(const FinishParams& p, cudaStream_t s) {
  if (!p.nb) return cudaSuccess;
  k_finish<<<(p.nb + 3) / 4, 128, 0, s>>>(p);
  return cudaGetLastError();
}

__global__ void k_scan(const uint64_t* len, uint64_t* off, uint32_t nb) {
  // nb is at most a few hundred thousand; a single warp walks it in 32-wide strides
  const int lane = threadIdx.x;
  uint64_t base = 0;
  for (uint32_t i0 = 0; i0 < nb; i0 += 32) {
    const uint32_t i = i0 + lane;
    uint64_t v = i < nb ? len[i] : 0, incl = v;
    for (int d = 1; d < 32; d <<= 1) {
      const uint64_t t = __shfl_up_sync(FULL, incl, d);
      if (lane >= d) incl += t;
    }
    if (i < nb) off[i] = base + incl - v;
    base += __shfl_sync(FULL, incl, 31);
  }
  if (lane == 0) off[nb] = base;
}
cudaError_t launch_scan(const uint64_t* len, uint64_t* off, uint32_t nb, cudaStream_t s) {
  k_scan<<<1, 32, 0, s>>>(len, off, nb);
  return cudaGetLastError();
}

// Copy variable-length ranges: grid (nb, chunks), 64 KB per CTA.
constexpr uint32_t kGatherChunk = 65536;
__global__ void k_gather(const uint8_t* src, const uint64_t* src_off, const uint64_t* len, uint8_t* dst,
                         const uint64_t* dst_off, uint64_t dst_cap) {
  const uint32_t b = blockIdx.x;
  const uint64_t n = len[b], begin = (uint64_t)blockIdx.y * kGatherChunk;
  if (begin >= n) return;
  const uint64_t end = min(n, begin + kGatherChunk);
  const uint8_t* s = src + src_off[b];
  const uint64_t d0 = dst_off[b];
  if (d0 + n > dst_cap) return;
  uint8_t* d = dst + d0;
  for (uint64_t i = begin + threadIdx.x; i < end; i += blockDim.x) d[i] = s[i];
}
cudaError_t launch_gather(const uint8_t* src, const uint64_t* src_off, const uint64_t* len, uint8_t* dst,
                          const uint64_t* dst_off, uint64_t dst_cap, uint32_t nb, uint64_t max_len, cudaStream_t s) {
  if (!nb || !max_len) return cudaSuccess;
  dim3 grid(nb, (unsigned)((max_len + kGatherChunk - 1) / kGatherChunk));
  k_gather<<<grid, 256, 0, s>>>(src, src_off, len, dst, dst_off, dst_cap);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Post-processing pass behind the speculative decoder (zpq_fdec.cuh): PostProcessor.write (PostProcessor.cs:37-86)
// over the raw model stream of a block -- first byte 0 = PASS (copy), 1 = PROG (length, PCOMP program, then run the
// program once per byte and once with 0xFFFFFFFF at the end of every segment).  One warp per resident arena (the PCOMP
// arrays H, M, R live there; the model tables in it are dead by now), jobs from an atomic queue.  Programs makeConfig
// emits are recognised by the host and run as native kernels instead (k_post_*); this kernel is the general case.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_post(const PostParams Q) {
  const int lane = threadIdx.x & 31;
  const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= Q.resident) return;
  const Plan* plan = Q.plan;
  uint8_t* arena = Q.arenas + (uint64_t)gw * Q.arena_stride;
  for (;;) {
    uint32_t job = 0;
    if (lane == 0) job = atomicAdd(Q.queue, 1u);
    job = __shfl_sync(FULL, job, 0);
    if (job >= Q.njobs) break;
    if (Q.jobkind && (Q.jobkind[job] & 15u) != PK_GENERIC) continue;      // restored by a native kernel (zpq_post.cu)
    const DecJob J = Q.djobs[job];
    const PostJob O = Q.pjobs[job];
    const BlockResult R = Q.raw_results[job];
    uint32_t status = R.status;
    uint64_t opos = 0;
    if (status == ZPQ_BLOCK_OK && J.seg_count) {
      const uint8_t* raw = Q.raw + J.out_off;
      uint8_t* out = Q.out + O.out_off;
      const uint64_t* send = Q.seg_end + J.seg_first;
      const uint64_t e0 = send[0];
      if (e0 < 1) status = ZPQ_BLOCK_POSTPROC;                       // "Unexpected EOS"
      else if (raw[0] == 0) {
        // PASS: every byte behind the type byte, segment ends are no-ops
        const uint64_t n = R.out_len - 1;
        if (n > O.out_cap) status = ZPQ_BLOCK_OVERFLOW;
        else for (uint64_t i = lane; i < n; i += 32) out[i] = raw[1 + i];
        opos = n;
        if (lane == 0) for (uint32_t k = 0; k < J.seg_count; ++k) Q.seg_out_end[J.seg_first + k] = send[k] - 1;
      } else if (raw[0] == 1) {
        uint32_t psize = 0;
        if (e0 < 3) status = ZPQ_BLOCK_POSTPROC;
        else {
          psize = raw[1] + 256u * raw[2];
          if (psize < 1 || 3ull + psize > e0) status = ZPQ_BLOCK_POSTPROC;   // "Empty PCOMP" / "Unexpected EOS"
        }
        if (status == ZPQ_BLOCK_OK) {
          uint8_t* pcode = arena + plan->off_pcode;
          for (uint32_t i = lane; i < psize + 3; i += 32) pcode[i] = i < psize ? raw[3 + i] : 0;
          // ZPAQL.initp: H, M, R zeroed
          uint32_t* H = reinterpret_cast<uint32_t*>(arena + plan->off_ph);
          uint8_t* M = arena + plan->off_pm;
          uint32_t* Rr = reinterpret_cast<uint32_t*>(arena + plan->off_pr);
          const uint64_t hn = 1ull << plan->ph, mn4 = ((1ull << plan->pm) + 3) / 4;
          for (uint64_t i = lane; i < hn; i += 32) H[i] = 0;
          for (uint64_t i = lane; i < mn4; i += 32) reinterpret_cast<uint32_t*>(M)[i] = 0;
          for (uint32_t i = lane; i < 256; i += 32) Rr[i] = 0;
          __syncwarp();
          if (lane == 0) {
            VM pvm; pvm.b = pvm.c = pvm.d = pvm.f = 0;
            VMEnv penv;
            penv.code = pcode; penv.len = (int)psize;
            penv.H = H; penv.hmask = (1u << plan->ph) - 1;
            penv.M = M; penv.mmask = (uint32_t)((1ull << plan->pm) - 1);
            penv.R = Rr;
            penv.out = out; penv.out_pos = 0; penv.out_cap = O.out_cap;
            uint64_t pos = 3ull + psize;
            for (uint32_t sg = 0; sg < J.seg_count && status == ZPQ_BLOCK_OK; ++sg) {
              const uint64_t end = send[sg];
              const uint64_t budget = 65536 + 512 * (end + O.out_cap);
              for (; pos < end; ++pos)
                if (zpaql_run(pvm, penv, raw[pos], budget)) { status = ZPQ_BLOCK_ZPAQL; break; }
              if (status == ZPQ_BLOCK_OK && zpaql_run(pvm, penv, 0xFFFFFFFFu, budget)) status = ZPQ_BLOCK_ZPAQL;
              Q.seg_out_end[J.seg_first + sg] = penv.out_pos;
            }
            opos = penv.out_pos;
          }
          status = __shfl_sync(FULL, status, 0);
          opos = __shfl_sync(FULL, (unsigned long long)opos, 0);
          if (status == ZPQ_BLOCK_OK && opos > O.out_cap) status = ZPQ_BLOCK_OVERFLOW;
        }
      } else status = ZPQ_BLOCK_POSTPROC;                             // "unknown post processing type"
    }
    __syncwarp();
    if (lane == 0) { Q.results[job].out_len = opos; Q.results[job].status = status; }
  }
}
cudaError_t launch_post(const PostParams& q, cudaStream_t s) {
  if (!q.njobs) return cudaSuccess;
  const uint32_t warps = 8;
  k_post<<<(q.resident + warps - 1) / warps, warps * 32, 0, s>>>(q);
  return cudaGetLastError();
}

}  // namespace zpq

namespace zpq {

// Lane-resident kernels with the run-time model walker (any model with 1..32 components).
__global__ void __launch_bounds__(kCtaThreads, 1) k_zpaq_encode_lanes(const CodecParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  encode_lanes_body<GenericModel>(P, smem);
}
__global__ void __launch_bounds__(kCtaThreads, 1) k_zpaq_decode_lanes(const CodecParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  decode_lanes_body<GenericModel>(P, smem);
}

cudaError_t codec_set_smem_limit(uint32_t bytes) {
  const void* ks[4] = {(const void*)k_zpaq_encode, (const void*)k_zpaq_decode, (const void*)k_zpaq_encode_lanes,
                       (const void*)k_zpaq_decode_lanes};
  for (const void* k : ks) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_encode(const CodecParams& p, LaunchGeom g, cudaStream_t s) {
  if (g.lanes) k_zpaq_encode_lanes<<<g.grid, g.warps_per_cta * 32, p.sm.total, s>>>(p);
  else k_zpaq_encode<<<g.grid, g.warps_per_cta * 32, p.sm.total, s>>>(p);
  return cudaGetLastError();
}
cudaError_t launch_decode(const CodecParams& p, LaunchGeom g, cudaStream_t s) {
  if (g.lanes) k_zpaq_decode_lanes<<<g.grid, g.warps_per_cta * 32, p.sm.total, s>>>(p);
  else k_zpaq_decode<<<g.grid, g.warps_per_cta * 32, p.sm.total, s>>>(p);
  return cudaGetLastError();
}

}  // namespace zpq
