// zpq_post.cu -- native post-processors for sm_100a: the decode-side inverse of LZ77 / BWT / E8E9.
//
// The reference restores a block by running the PCOMP program stored in the block once per decoded byte
// (PostProcessor.write, PostProcessor.cs:37-86; ZPAQL.run).  libzpaq's makeConfig only ever emits four programs
// (LibZPAQ.cs:427-826): "lazy2" (bit-packed LZ77), "lzpre" (byte-aligned LZ77), "bwtrle" (inverse BWT) -- each with or
// without the inverse E8E9 pass -- and "e8e9".  The decoder kernels leave the raw model stream (type byte, program,
// transformed data) in device memory; k_post_classify compares the stored program with the candidates the host front end
// generated for the block's (ph, pm) and the kernels below compute what the program would have written:
//
//   k_post_pass   type 0: copy (one CTA per 64 KB)
//   k_post_lz     lazy2 / lzpre / e8e9: one warp per block; codes are parsed warp-uniformly exactly as the program parses
//                 them (lazy2: the per-byte bit-buffer state machine of LibZPAQ.cs:440-571 restated), literals and match
//                 copies are done by the 32 lanes
//   k_post_bwt    bwtrle: one 1024-thread CTA per block: byte histogram, stable counting sort into the linked list the
//                 program builds in H (LibZPAQ.cs:664-695), then the list traversal (:700-712) split at every position
//                 that is a multiple of a stride: the sub-lists are walked in parallel, ranked, and walked again to emit
//   un_e8e9       the ascending E8E9 inverse (LibZPAQ.cs:444-463), 32 positions per step until a position hits
//
// A kernel that meets a stream it does not reproduce bit for bit (a match reaching in front of the buffer, more than one
// BWT segment, output beyond 2^pm, ...) hands the job back (PK_GENERIC) and the interpreter pass (k_post, zpq_kernels.cu)
// runs the stored program itself; unknown programs always go there.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "../../include/zpaqb200.h"
#include "zpq_device.h"

namespace zpq {

namespace {

constexpr uint32_t FULLM = 0xFFFFFFFFu;

// ---- geometry of a job's raw stream ----
struct RawJob {
  const uint8_t* raw;
  const uint64_t* send;     // raw position at the end of every segment
  uint32_t nseg;
  uint64_t first;           // raw position of the first data byte (behind type byte and program)
  uint8_t* out;
  uint64_t cap;
};

__device__ __forceinline__ void raw_job(const PostParams& Q, uint32_t job, RawJob& r) {
  const DecJob J = Q.djobs[job];
  const PostJob O = Q.pjobs[job];
  r.raw = Q.raw + J.out_off;
  r.send = Q.seg_end + J.seg_first;
  r.nseg = J.seg_count;
  r.out = Q.out + O.out_off;
  r.cap = O.out_cap;
  r.first = r.raw[0] == 1 ? 3ull + r.raw[1] + 256u * r.raw[2] : 1ull;
}

// ------------------------------------------------------------------------------------------
// Which program does the block carry?  One warp per job.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_post_classify(const PostParams Q) {
  const int lane = threadIdx.x & 31;
  const uint32_t job = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (job >= Q.njobs) return;
  const DecJob J = Q.djobs[job];
  const BlockResult R = Q.raw_results[job];
  uint32_t kind = PK_GENERIC;
  if (R.status == ZPQ_BLOCK_OK && J.seg_count) {
    const uint8_t* raw = Q.raw + J.out_off;
    const uint64_t e0 = Q.seg_end[J.seg_first];
    if (e0 >= 1 && raw[0] == 0) kind = PK_PASS;
    else if (e0 >= 3 && raw[0] == 1) {
      const uint32_t psize = raw[1] + 256u * raw[2];
      if (psize >= 1 && 3ull + psize <= e0) {
        for (uint32_t c = 0; c < Q.ncand && kind == PK_GENERIC; ++c) {
          const PostCand C = Q.cands[c];
          if (C.len != psize) continue;
          bool same = true;
          for (uint32_t i = lane; i < psize; i += 32) same = same && ((int)i == C.wild || raw[3 + i] == Q.cand_bytes[C.off + i]);
          if (__all_sync(FULLM, same)) {
            const uint32_t param = C.wild >= 0 ? (uint32_t)raw[3 + C.wild] : C.param;
            kind = C.kind | (C.e8 << 4) | (param << 8);
          }
        }
      }
    }
  }
  if (lane == 0) Q.jobkind[job] = kind;
}

// ------------------------------------------------------------------------------------------
// PASS (PostProcessor.cs:42-49, state 1): every byte behind the type byte; segment ends are no-ops.
// grid (njobs, 64 KB chunks)
// ------------------------------------------------------------------------------------------
constexpr uint32_t kPassChunk = 65536;
__global__ void __launch_bounds__(256) k_post_pass(const PostParams Q) {
  const uint32_t job = blockIdx.x;
  if ((Q.jobkind[job] & 15u) != PK_PASS) return;
  const DecJob J = Q.djobs[job];
  const PostJob O = Q.pjobs[job];
  const uint64_t n = Q.raw_results[job].out_len - 1;
  if (blockIdx.y == 0 && threadIdx.x == 0) {
    Q.results[job].out_len = n;
    Q.results[job].status = n > O.out_cap ? ZPQ_BLOCK_OVERFLOW : ZPQ_BLOCK_OK;
    for (uint32_t k = 0; k < J.seg_count; ++k) Q.seg_out_end[J.seg_first + k] = Q.seg_end[J.seg_first + k] - 1;
  }
  if (n > O.out_cap) return;
  const uint8_t* src = Q.raw + J.out_off + 1;
  uint8_t* dst = Q.out + O.out_off;
  for (uint64_t begin = (uint64_t)blockIdx.y * kPassChunk; begin < n; begin += (uint64_t)gridDim.y * kPassChunk) {
    const uint64_t end = min(n, begin + kPassChunk);
    // slots are 16-byte aligned and the chunk size is a multiple of 4: whole words from two aligned source words
    if (((uintptr_t)dst & 3u) == 0) {
      const uint32_t sh = (uint32_t)((uintptr_t)(src + begin) & 3u) * 8u;
      const uint32_t* s32 = reinterpret_cast<const uint32_t*>((uintptr_t)(src + begin) & ~(uintptr_t)3);
      uint32_t* d32 = reinterpret_cast<uint32_t*>(dst + begin);
      const uint64_t words = (end - begin) / 4;
      for (uint64_t i = threadIdx.x; i < words; i += blockDim.x) d32[i] = __funnelshift_r(s32[i], s32[i + 1], sh);
      for (uint64_t i = begin + words * 4 + threadIdx.x; i < end; i += blockDim.x) dst[i] = src[i];
    } else {
      for (uint64_t i = begin + threadIdx.x; i < end; i += blockDim.x) dst[i] = src[i];
    }
  }
}

// ------------------------------------------------------------------------------------------
// Inverse E8E9 in place (LibZPAQ.cs:444-463 inside lazy2 / lzpre / bwtrle, :801-826 as a program of its own): for i
// ascending with i + 4 < n: if x[i] is E8 or E9 and x[i+4] is 00 or FF, the 24-bit little-endian field behind x[i] loses
// i.  A hit rewrites bytes that later positions test, so the warp looks at 32 positions, applies the first hit and
// looks again behind it.
// ------------------------------------------------------------------------------------------
__device__ void un_e8e9(uint8_t* x, uint64_t n, int lane) {
  uint64_t base = 0;
  __syncwarp();
  while (base + 4 < n) {
    const uint64_t i = base + (uint32_t)lane;
    bool hit = false;
    if (i + 4 < n) hit = (x[i] & 254u) == 0xE8u && ((x[i + 4] + 1u) & 254u) == 0u;
    const uint32_t m = __ballot_sync(FULLM, hit);
    if (!m) { base += 32; continue; }
    const int k = __ffs((int)m) - 1;
    if (lane == k) {
      uint32_t a = (uint32_t)x[i + 1] | (uint32_t)x[i + 2] << 8 | (uint32_t)x[i + 3] << 16;
      a -= (uint32_t)i;
      x[i + 1] = (uint8_t)a; x[i + 2] = (uint8_t)(a >> 8); x[i + 3] = (uint8_t)(a >> 16);
    }
    __syncwarp();
    base += (uint32_t)k + 1u;
  }
  __syncwarp();
}

// M[b .. b+len) = M[src .. src+len) in ascending order, as the byte loop of the programs does (overlap repeats the
// period); src < b.  All lanes call.
__device__ __forceinline__ void lz_copy(uint8_t* M, uint64_t b, uint64_t src, uint64_t len, int lane) {
  __syncwarp();
  const uint64_t dist = b - src;
  if (dist >= len) {
    for (uint64_t i = (uint32_t)lane; i < len; i += 32) M[b + i] = M[src + i];
  } else if (dist >= 32) {
    for (uint64_t i0 = 0; i0 < len; i0 += 32) {
      const uint64_t i = i0 + (uint32_t)lane;
      if (i < len) M[b + i] = M[src + i];
      __syncwarp();
    }
  } else {
    const uint32_t d = (uint32_t)dist;
    uint32_t ph = (uint32_t)lane % d;
    const uint32_t step = 32u % d;
    for (uint64_t i = (uint32_t)lane; i < len; i += 32) {
      M[b + i] = M[src + ph];
      ph += step; if (ph >= d) ph -= d;
    }
  }
  __syncwarp();
}

enum { LZ_OK = 0, LZ_OVERFLOW = 1, LZ_ODD = 2 };

// "lzpre" (LibZPAQ.cs:577-638): code byte c < 64: c + 1 literals follow; else (c >> 6) + 1 offset bytes, most significant
// first, and the match is (c & 63) + minMatch bytes from offset + 1 back.
__device__ int lz_bytes_segment(const uint8_t* in, uint64_t n, uint8_t* M, uint64_t cap, uint64_t mlimit, uint32_t minMatch, int lane,
                                uint64_t& produced) {
  uint64_t pos = 0, b = 0;
  while (pos < n) {
    const uint32_t c = in[pos];
    if (c < 64u) {
      const uint64_t len = c + 1u, avail = min(len, n - pos - 1);
      if (b + avail > mlimit) return LZ_ODD;
      if (b + avail > cap) { produced = b + avail; return LZ_OVERFLOW; }
      for (uint64_t i = (uint32_t)lane; i < avail; i += 32) M[b + i] = in[pos + 1 + i];
      b += avail; pos += 1 + len;
    } else {
      const uint32_t nb = (c >> 6) + 1u;
      if (pos + 1 + nb > n) break;                   // the stream ends inside a code: nothing more is written
      uint32_t off = 0;
      for (uint32_t k = 0; k < nb; ++k) off = off << 8 | in[pos + 1 + k];
      const uint64_t len = (c & 63u) + minMatch;
      if (len == 0 || (uint64_t)off + 1 > b || b + len > mlimit) return LZ_ODD;
      if (b + len > cap) { produced = b + len; return LZ_OVERFLOW; }
      lz_copy(M, b, b - off - 1, len, lane);
      b += len; pos += 1 + nb;
    }
  }
  produced = b;
  return LZ_OK;
}

// "lazy2" (LibZPAQ.cs:430-571), restated state by state: the program is run once per stream byte; it appends the byte to
// a bit buffer (c = bits, d = n) and lets the states consume what is there.  r1 = state, r2 = len, r3 = m, r4 = ptr,
// r5 = low offset bits.  Registers are 32 bits wide and shifts take their count modulo 32, as in ZPAQL.
__device__ int lz_bits_segment(const uint8_t* in, uint64_t n, uint8_t* M, uint64_t cap, uint64_t mlimit, uint32_t rb, int lane,
                               uint64_t& produced) {
  uint32_t bits = 0, nbit = 0, state = 0, len = 0, m = 0, r5 = 0;
  uint64_t ptr = 0;
  for (uint64_t pos = 0; pos < n; ++pos) {
    bits += (uint32_t)in[pos] << (nbit & 31u);
    nbit += 8;
    if (state == 0) {                                          // expect a new code: mm,mmm (match) or 00 (literals)
      len = 1;
      if (bits & 3u) {
        m = ((bits & 3u) - 1u) * 8u; bits >>= 2;
        m += bits & 7u; bits >>= 3;
        nbit -= 5; state = 1;
      } else { bits >>= 2; nbit -= 2; state = 3; }
    }
    while (state == 1 && nbit > 2u) {                          // match length: (1 b)* 0 ll
      if (bits & 1u) { bits >>= 1; len = len + len + (bits & 1u); bits >>= 1; nbit -= 2; }
      else { bits >>= 1; len = (len << 2) + (bits & 3u); bits >>= 2; nbit -= 3; state = rb ? 5u : 2u; }
    }
    if (rb && state == 5 && nbit > rb - 1u) { r5 = bits & ((1u << rb) - 1u); bits >>= rb; nbit -= rb; state = 2; }
    if (state == 2 && !(m > nbit)) {                           // m offset bits
      uint32_t off = (bits & ((1u << m) - 1u)) + (1u << m);
      if (rb) off = (off << rb) + r5 - ((1u << rb) - 1u);
      if (off == 0 || (uint64_t)off > ptr || ptr + len > mlimit) return LZ_ODD;
      if (ptr + len > cap) { produced = ptr + len; return LZ_OVERFLOW; }
      lz_copy(M, ptr, ptr - off, len, lane);
      ptr += len;
      bits >>= m; nbit -= m; state = 0;
    }
    while (state == 3 && nbit > 1u) {                          // literal length: (1 b)* 0
      if (bits & 1u) { bits >>= 1; len = len + len + (bits & 1u); bits >>= 1; nbit -= 2; }
      else { bits >>= 1; nbit -= 1; state = 4; }
    }
    if (state == 4 && nbit > 7u) {                             // one literal per stream byte
      if (ptr + 1 > mlimit) return LZ_ODD;
      if (ptr + 1 > cap) { produced = ptr + 1; return LZ_OVERFLOW; }
      if (lane == 0) M[ptr] = (uint8_t)bits;
      ++ptr; bits >>= 8; nbit -= 8;
      if (--len == 0) state = 0;
    }
  }
  produced = ptr;
  return LZ_OK;
}

// One warp per job: lazy2, lzpre and e8e9.  Every segment restarts the program state (its end-of-segment branch resets
// the registers, LibZPAQ.cs:465, :599, :805-812), so a segment is restored into out + (bytes of the segments before it).
__global__ void __launch_bounds__(128) k_post_lz(const PostParams Q) {
  const int lane = threadIdx.x & 31;
  const uint32_t job = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (job >= Q.njobs) return;
  const uint32_t jk = Q.jobkind[job], kind = jk & 15u, e8 = (jk >> 4) & 1u, param = jk >> 8;
  if (kind != PK_LZ_BITS && kind != PK_LZ_BYTES && kind != PK_E8E9) return;
  RawJob r;
  raw_job(Q, job, r);
  const uint64_t mlimit = kind == PK_E8E9 ? ~0ull : (1ull << Q.plan->pm);
  uint64_t opos = 0, pos = r.first;
  int rc = LZ_OK;
  for (uint32_t sg = 0; sg < r.nseg && rc == LZ_OK; ++sg) {
    const uint64_t end = r.send[sg];
    if (end < pos) { rc = LZ_ODD; break; }
    const uint8_t* in = r.raw + pos;
    const uint64_t n = end - pos;
    uint8_t* M = r.out + opos;
    const uint64_t cap = r.cap - opos;
    uint64_t produced = 0;
    if (kind == PK_E8E9) {
      produced = n;
      if (n > cap) rc = LZ_OVERFLOW;
      else for (uint64_t i = (uint32_t)lane; i < n; i += 32) M[i] = in[i];
    } else if (kind == PK_LZ_BYTES) rc = lz_bytes_segment(in, n, M, cap, mlimit, param, lane, produced);
    else rc = lz_bits_segment(in, n, M, cap, mlimit, param, lane, produced);
    if (rc == LZ_OK && (e8 || kind == PK_E8E9)) un_e8e9(M, produced, lane);
    opos += produced;
    pos = end;
    if (lane == 0) Q.seg_out_end[Q.djobs[job].seg_first + sg] = opos;
  }
  __syncwarp();
  if (lane == 0) {
    if (rc == LZ_ODD) Q.jobkind[job] = PK_GENERIC;
    else { Q.results[job].out_len = opos; Q.results[job].status = rc == LZ_OK ? ZPQ_BLOCK_OK : ZPQ_BLOCK_OVERFLOW; }
  }
}

// ------------------------------------------------------------------------------------------
// "bwtrle" (LibZPAQ.cs:644-794).  The stream is the BWT of the block with a dummy byte at the position of the
// end-of-string symbol, then that position as 4 bytes, low byte first.  With C[v] = 1 + number of stream bytes < v the
// program files every position b != idx, in ascending order, under T[C[M[b]]++] = b, and then emits M[p] for
// p = T[idx], T[T[idx]], ... until p == 0.  T is injective and idx is not in its image, so the walk from idx is a simple
// path that ends at position 0.
// ------------------------------------------------------------------------------------------
constexpr int kBwtThreads = 1024;
constexpr uint32_t kBwtSplit = 2048;         // sub-lists per block at most
constexpr uint32_t kUnset = 0xFFFFFFFFu;

__global__ void __launch_bounds__(kBwtThreads) k_post_bwt(const PostParams Q) {
  __shared__ uint32_t wh[32 * 256];          // per-warp bucket cursors; later len / end / offs of the sub-lists
  __shared__ uint32_t cnt[256];
  __shared__ uint32_t misc[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Plan* plan = Q.plan;
  uint32_t* T = reinterpret_cast<uint32_t*>(Q.arenas + (uint64_t)blockIdx.x * Q.arena_stride + plan->off_ph);
  for (;;) {
    __syncthreads();
    if (tid == 0) misc[0] = atomicAdd(Q.queue2, 1u);
    __syncthreads();
    const uint32_t job = misc[0];
    if (job >= Q.njobs) break;
    const uint32_t jk = Q.jobkind[job];
    if ((jk & 15u) != PK_BWT) continue;
    const uint32_t e8 = (jk >> 4) & 1u;
    RawJob r;
    raw_job(Q, job, r);
    const uint64_t e0 = r.send[0];
    const uint64_t nin = e0 - r.first;
    const uint8_t* M = r.raw + r.first;
    bool odd = r.nseg != 1 || nin < 5 || nin - 4 + 256 > (1ull << plan->ph) || nin - 4 > 0xFFFFFF00ull;
    uint32_t size = 0, idx = 0;
    if (!odd) {
      size = (uint32_t)(nin - 4);
      idx = (uint32_t)M[size] | (uint32_t)M[size + 1] << 8 | (uint32_t)M[size + 2] << 16 | (uint32_t)M[size + 3] << 24;
      odd = idx >= size;
    }
    if (odd) { if (tid == 0) Q.jobkind[job] = PK_GENERIC; continue; }
    const bool packed = size <= (1u << 24);        // position and byte in one word, as the program does for blocks up to 16 MB (:697-704)

    // ---- bucket sizes per warp chunk (position idx is not filed, but it is counted, :664-669) ----
    for (int i = tid; i < 32 * 256; i += kBwtThreads) wh[i] = 0;
    __syncthreads();
    const uint32_t csz = (((size + 31u) / 32u) + 31u) & ~31u;
    const uint32_t c0 = (uint32_t)warp * csz, c1 = min(size, c0 + csz);
    for (uint32_t b0 = c0; b0 < c1; b0 += 32) {
      const uint32_t b = b0 + (uint32_t)lane;
      const bool valid = b < c1 && b != idx;
      const uint32_t v = valid ? (uint32_t)M[b] : 256u + (uint32_t)lane;
      const uint32_t peers = __match_any_sync(FULLM, v);
      if (valid && (peers >> lane) == 1u) wh[warp * 256 + v] += (uint32_t)__popc(peers);
      __syncwarp();
    }
    __syncthreads();
    if (tid < 256) {
      uint32_t s = 0;
      for (int w = 0; w < 32; ++w) s += wh[w * 256 + tid];
      cnt[tid] = s + ((uint32_t)M[idx] == (uint32_t)tid ? 1u : 0u);
    }
    __syncthreads();
    if (tid < 256) {
      uint32_t run = 1;                            // C[v] = 1 + bytes below v (:671-675)
      for (int u = 0; u < tid; ++u) run += cnt[u];
      for (int w = 0; w < 32; ++w) { const uint32_t t = wh[w * 256 + tid]; wh[w * 256 + tid] = run; run += t; }
    }
    if (tid == 0) T[0] = 0;
    __syncthreads();
    // ---- the list: stable counting sort of the positions by their byte (:677-695) ----
    for (uint32_t b0 = c0; b0 < c1; b0 += 32) {
      const uint32_t b = b0 + (uint32_t)lane;
      const bool valid = b < c1 && b != idx;
      const uint32_t v = valid ? (uint32_t)M[b] : 256u + (uint32_t)lane;
      const uint32_t peers = __match_any_sync(FULLM, v);
      uint32_t pos = 0;
      if (valid) pos = wh[warp * 256 + v] + (uint32_t)__popc(peers & ((1u << lane) - 1u));
      __syncwarp();
      if (valid) {
        T[pos] = b;
        if ((peers >> lane) == 1u) wh[warp * 256 + v] += (uint32_t)__popc(peers);
      }
      __syncwarp();
    }
    __syncthreads();
    if (packed) {
      for (uint32_t p = tid; p < size; p += kBwtThreads) T[p] = T[p] << 8 | (uint32_t)M[p];
      __syncthreads();
    }
    // ---- sub-lists: one starts at idx, one at every other multiple of `stride`; each ends at the next multiple ----
    uint32_t stride = 64;
    while ((size + stride - 1) / stride > kBwtSplit) stride <<= 1;
    const uint32_t nsplit = (size + stride - 1) / stride;
    uint32_t* slen = wh; uint32_t* send_ = wh + kBwtSplit; uint32_t* soff = wh + 2 * kBwtSplit;
    for (uint32_t j = tid; j < nsplit; j += kBwtThreads) {
      uint32_t p = j ? j * stride : idx, len = 0;
      if (p != 0) {
        do { p = packed ? T[p] >> 8 : T[p]; ++len; } while ((p & (stride - 1)) != 0 && len <= size);
      }
      slen[j] = len; send_[j] = (p & (stride - 1)) == 0 ? p / stride : 0u; soff[j] = kUnset;
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t cur = 0, off = 0;
      for (uint32_t it = 0; it <= nsplit; ++it) {
        if (soff[cur] != kUnset) break;
        soff[cur] = off; off += slen[cur];
        const uint32_t e = send_[cur];
        if (e == 0 || e >= nsplit) break;
        cur = e;
      }
      misc[1] = off;
    }
    __syncthreads();
    const uint32_t total = misc[1];
    if (total <= r.cap) {
      for (uint32_t j = tid; j < nsplit; j += kBwtThreads) {
        const uint32_t o = soff[j];
        if (o == kUnset) continue;
        uint32_t p = j ? j * stride : idx;
        const uint32_t len = slen[j];
        if (packed) {
          uint32_t v = T[p];
          for (uint32_t k = 0; k < len; ++k) { v = T[v >> 8]; r.out[o + k] = (uint8_t)v; }
        } else {
          for (uint32_t k = 0; k < len; ++k) { p = T[p]; r.out[o + k] = M[p]; }
        }
      }
      __syncthreads();
      if (e8 && warp == 0) un_e8e9(r.out, total, lane);
    }
    if (tid == 0) {
      Q.results[job].out_len = total; Q.results[job].status = total > r.cap ? ZPQ_BLOCK_OVERFLOW : ZPQ_BLOCK_OK;
      Q.seg_out_end[Q.djobs[job].seg_first] = total;
    }
  }
}

}  // namespace

cudaError_t launch_post_native(const PostParams& q, cudaStream_t s) {
  if (!q.njobs) return cudaSuccess;
  k_post_classify<<<(q.njobs + 7) / 8, 256, 0, s>>>(q);
  const uint32_t chunks = (uint32_t)std::min<uint64_t>(65535, std::max<uint64_t>(1, (q.max_raw + kPassChunk - 1) / kPassChunk));
  k_post_pass<<<dim3(q.njobs, chunks), 256, 0, s>>>(q);
  if (q.ncand) {
    k_post_lz<<<(q.njobs + 3) / 4, 128, 0, s>>>(q);
    const uint32_t ctas = std::min(q.resident, q.njobs);
    if (q.has_bwt) k_post_bwt<<<ctas, kBwtThreads, 0, s>>>(q);
  }
  return cudaGetLastError();
}

}  // namespace zpq
