// zpq_preproc.cu -- pre-processing kernels of the compress path (LZBuffer.cs:151-486):
//   suffix arrays for a whole batch of blocks by prefix doubling (rank pairs -> radix sort ->
//   re-rank), BWT emission, and the LZ77 parsers (hash-table matcher and suffix-array matcher)
//   with both code formats (bit-packed level 1, byte-aligned level 2).
//
// Parity is the constraint: the transformed stream must be byte-identical to what the
// reference's greedy parser produces, because it is the input of the context model.  The suffix
// array is unique, so any correct construction gives the same BWT; the LZ77 parse depends on
// running state (pending literals, best length so far, bucket replacement order), so each block
// is parsed by one warp that follows the reference's decisions exactly while its lanes share the
// candidate probing and match-length work.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "zpq_device.h"

namespace zpq {

#define FULL 0xFFFFFFFFu

// ------------------------------------------------------------------------------------------
// Suffix arrays of all blocks of a batch at once (replaces divsufsort, divsufsort.cs:1940).
// Positions are global indices g into the concatenated batch; blk[g] is the block of g.
// Round k sorts the keys  blk | rank[g] | rank[g+k]  (0 past the block end); equal keys share
// the index of their group head as the new rank.  Finished when every key is distinct.
// ------------------------------------------------------------------------------------------
__global__ void k_sa_blockids(const uint64_t* off, uint32_t nb, uint64_t n_total, uint32_t* blk) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_total) return;
  uint32_t lo = 0, hi = nb;  // largest b with off[b] - off[0] <= g
  const uint64_t base = off[0];
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (off[mid] - base <= g) lo = mid; else hi = mid;
  }
  blk[g] = lo;
}

__global__ void k_sa_init_rank(const uint8_t* in, uint64_t n_total, uint32_t* rank) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n_total) rank[g] = (uint32_t)in[g] + 1;
}

__global__ void k_sa_keys(const uint32_t* rank, const uint32_t* blk, const uint64_t* off, uint64_t n_total, uint32_t k,
                          int rbits, uint64_t* key, uint32_t* val) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_total) return;
  const uint32_t b = blk[g];
  const uint64_t end = off[b + 1] - off[0];
  const uint64_t r2 = g + k < end ? rank[g + k] : 0;
  key[g] = ((uint64_t)b << (2 * rbits)) | ((uint64_t)rank[g] << rbits) | r2;
  val[g] = (uint32_t)g;
}

__global__ void k_sa_heads(const uint64_t* key, uint64_t n_total, uint32_t* head, unsigned long long* distinct) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t is_head = 0;
  if (j < n_total) {
    is_head = (j == 0 || key[j] != key[j - 1]) ? 1u : 0u;
    head[j] = is_head ? (uint32_t)j : 0u;
  }
  const uint32_t cnt = __reduce_add_sync(FULL, is_head);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(distinct, (unsigned long long)cnt);
}

__global__ void k_sa_rerank(const uint32_t* head_scanned, const uint32_t* val, const uint32_t* blk, const uint64_t* off,
                            uint64_t n_total, uint32_t* rank) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_total) return;
  const uint32_t g = val[j];
  const uint64_t start = off[blk[g]] - off[0];   // the block's suffixes occupy sorted positions [start, start + n_b)
  rank[g] = (uint32_t)(head_scanned[j] - start) + 1;
}

// sa_local[j] = position inside its block of the j-th smallest suffix; isa[g] = rank inside the block
__global__ void k_sa_finish(const uint32_t* val, const uint32_t* blk, const uint64_t* off, uint64_t n_total, uint32_t* sa,
                            uint32_t* isa) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_total) return;
  const uint32_t g = val[j];
  const uint64_t start = off[blk[g]] - off[0];
  sa[j] = (uint32_t)(g - start);
  if (isa) isa[g] = (uint32_t)(j - start);
}

struct MaxOp {
  __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

size_t sa_workspace_bytes(uint64_t n_total) {
  size_t sort_tmp = 0, scan_tmp = 0;
  cub::DoubleBuffer<uint64_t> dk(nullptr, nullptr);
  cub::DoubleBuffer<uint32_t> dv(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, dk, dv, (int64_t)n_total, 0, 64);
  cub::DeviceScan::InclusiveScan(nullptr, scan_tmp, (uint32_t*)nullptr, (uint32_t*)nullptr, MaxOp(), (int64_t)n_total);
  const size_t tmp = (sort_tmp > scan_tmp ? sort_tmp : scan_tmp) + 256;
  // keys x2, vals x2, rank, blk, head, temp, counter
  return (size_t)n_total * (8 * 2 + 4 * 2 + 4 + 4 + 4) + tmp + 4096;
}

// Builds SA (and ISA if isa != nullptr) of every block.  d_off: device array of nb+1 absolute
// offsets into the batch (in = batch base + (off[b]-off[0])).  max_len = longest block.
// Synchronises the stream once per round to read the distinct-key count.
cudaError_t build_suffix_arrays(const uint8_t* in, const uint64_t* d_off, uint32_t nb, uint64_t n_total, uint64_t max_len,
                                uint32_t* sa, uint32_t* isa, void* workspace, size_t workspace_bytes, cudaStream_t s,
                                int* rounds_out) {
  if (rounds_out) *rounds_out = 0;
  if (!n_total) return cudaSuccess;
  if (workspace_bytes < sa_workspace_bytes(n_total)) return cudaErrorInvalidValue;
  int rbits = 1;
  while ((1ull << rbits) <= max_len + 1) ++rbits;
  int bbits = 1;
  while ((1ull << bbits) < nb) ++bbits;
  if (2 * rbits + bbits > 64) return cudaErrorInvalidValue;   // caller must split the batch
  uint8_t* w = static_cast<uint8_t*>(workspace);
  auto take = [&](size_t bytes) { uint8_t* p = w; w += (bytes + 255) / 256 * 256; return p; };
  uint64_t* key0 = (uint64_t*)take(8 * n_total);
  uint64_t* key1 = (uint64_t*)take(8 * n_total);
  uint32_t* val0 = (uint32_t*)take(4 * n_total);
  uint32_t* val1 = (uint32_t*)take(4 * n_total);
  uint32_t* rank = (uint32_t*)take(4 * n_total);
  uint32_t* blk = (uint32_t*)take(4 * n_total);
  uint32_t* head = (uint32_t*)take(4 * n_total);
  unsigned long long* counter = (unsigned long long*)take(256);
  size_t tmp_bytes = workspace_bytes - (size_t)(w - static_cast<uint8_t*>(workspace));
  void* tmp = w;

  const int T = 256;
  const unsigned G = (unsigned)((n_total + T - 1) / T);
  k_sa_blockids<<<G, T, 0, s>>>(d_off, nb, n_total, blk);
  k_sa_init_rank<<<G, T, 0, s>>>(in, n_total, rank);
  cudaError_t e;
  int rounds = 0;
  const uint32_t* final_vals = nullptr;
  for (uint64_t k = 1;; k <<= 1) {
    ++rounds;
    k_sa_keys<<<G, T, 0, s>>>(rank, blk, d_off, n_total, (uint32_t)(k > 0xFFFFFFFFull ? 0xFFFFFFFFu : k), rbits, key0, val0);
    cub::DoubleBuffer<uint64_t> dk(key0, key1);
    cub::DoubleBuffer<uint32_t> dv(val0, val1);
    size_t tb = tmp_bytes;
    if ((e = cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dv, (int64_t)n_total, 0, 2 * rbits + bbits, s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(counter, 0, 8, s)) != cudaSuccess) return e;
    k_sa_heads<<<G, T, 0, s>>>(dk.Current(), n_total, head, counter);
    tb = tmp_bytes;
    if ((e = cub::DeviceScan::InclusiveScan(tmp, tb, head, head, MaxOp(), (int64_t)n_total, s)) != cudaSuccess) return e;
    k_sa_rerank<<<G, T, 0, s>>>(head, dv.Current(), blk, d_off, n_total, rank);
    unsigned long long distinct = 0;
    if ((e = cudaMemcpyAsync(&distinct, counter, 8, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    final_vals = dv.Current();
    if (distinct == n_total || k >= max_len) break;
  }
  k_sa_finish<<<G, T, 0, s>>>(final_vals, blk, d_off, n_total, sa, isa);
  if (rounds_out) *rounds_out = rounds;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// BWT emission (LZBuffer.cs:229-241): in[n-1], then for every rank the byte before the suffix
// (255 where the suffix is the whole block, whose rank+1 is the index), then the index, 4 bytes
// LSB first.  Output n + 5 bytes.
// ------------------------------------------------------------------------------------------
__global__ void k_bwt_emit(const uint8_t* in, const uint64_t* off, const uint32_t* sa, const uint64_t* out_off, uint8_t* out,
                           EncJob* jobs) {
  const uint32_t b = blockIdx.y;
  const uint64_t base = off[b] - off[0];
  const uint64_t n = off[b + 1] - off[b];
  const uint8_t* src = in + base;
  const uint32_t* s = sa + base;
  uint8_t* dst = out + out_off[b];
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (uint64_t)gridDim.x * blockDim.x) {
    if (i == 0) {
      dst[0] = n > 0 ? src[n - 1] : 255;
      if (n == 0) { dst[1] = dst[2] = dst[3] = dst[4] = 0; }   // idx = 0
      jobs[b].in_len = (uint32_t)(n + 5);
    } else {
      const uint32_t p = s[i - 1];
      if (p == 0) {
        dst[i] = 255;
        dst[n + 1] = (uint8_t)i; dst[n + 2] = (uint8_t)(i >> 8); dst[n + 3] = (uint8_t)(i >> 16); dst[n + 4] = (uint8_t)(i >> 24);
      } else dst[i] = src[p - 1];
    }
  }
}

cudaError_t launch_bwt_emit(const uint8_t* in, const uint64_t* d_off, const uint32_t* sa, const uint64_t* d_out_off, uint8_t* out,
                            EncJob* jobs, uint32_t nb, uint64_t max_len, cudaStream_t s) {
  if (!nb) return cudaSuccess;
  unsigned gx = (unsigned)((max_len + 1 + 255) / 256);
  if (gx > 1024) gx = 1024;
  if (gx < 1) gx = 1;
  k_bwt_emit<<<dim3(gx, nb), 256, 0, s>>>(in, d_off, sa, d_out_off, out, jobs);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// LZ77 (LZBuffer.cs:243-486).  One warp per block.  Lane 0 carries the parser state and emits;
// all lanes probe candidates and measure match lengths.
// ------------------------------------------------------------------------------------------
struct LzOut {
  uint8_t* p; uint64_t pos, cap;
  uint32_t bits, nbits;
  __device__ void put(uint32_t c) { if (pos < cap) p[pos] = (uint8_t)c; ++pos; }
  __device__ void putb(uint32_t x, int k) {             // LZBuffer.cs:53-62
    x &= (1u << k) - 1;
    bits |= x << nbits;
    nbits += k;
    while (nbits > 7) { put(bits); bits >>= 8; nbits -= 8; }
  }
  __device__ void flush() { if (nbits > 0) put(bits); bits = nbits = 0; }
};

__device__ __forceinline__ int lg32(uint32_t x) { return 32 - __clz(x); }   // LZBuffer.cs:118-127

__device__ void lz_write_literal(const LzParams& P, LzOut& o, const uint8_t* in, uint32_t i, uint32_t& lit) {   // :387-419
  if (P.level == 1) {
    if (lit < 1) return;
    int ll = lg32(lit);
    o.putb(0, 2);
    --ll;
    while (--ll >= 0) { o.putb(1, 1); o.putb((lit >> ll) & 1, 1); }
    o.putb(0, 1);
    while (lit) o.putb(in[i - lit--], 8);
  } else {
    while (lit > 0) {
      const uint32_t lit1 = lit > 64 ? 64 : lit;
      o.put(lit1 - 1);
      for (uint32_t j = i - lit; j < i - lit + lit1; ++j) o.put(in[j]);
      lit -= lit1;
    }
  }
}

__device__ void lz_write_match(const LzParams& P, LzOut& o, uint32_t len, uint32_t off) {   // :422-486
  if (P.level == 1) {
    int ll = lg32(len) - 1;
    off += (1u << P.rb) - 1;
    const int lo = lg32(off) - 1 - P.rb;
    o.putb((lo + 8) >> 3, 2);
    o.putb(lo & 7, 3);
    while (--ll >= 2) { o.putb(1, 1); o.putb((len >> ll) & 1, 1); }
    o.putb(0, 1);
    o.putb(len & 3, 2);
    o.putb(off, P.rb);
    o.putb(off >> P.rb, lo);
  } else {
    --off;
    while (len > 0) {
      const uint32_t len1 = len > P.minMatch * 2 + 63 ? P.minMatch + 63 : len > P.minMatch + 63 ? len - P.minMatch : len;
      if (off < (1u << 16)) { o.put(64 + len1 - P.minMatch); o.put(off >> 8); o.put(off); }
      else if (off < (1u << 24)) { o.put(128 + len1 - P.minMatch); o.put(off >> 16); o.put(off >> 8); o.put(off); }
      else { o.put(192 + len1 - P.minMatch); o.put(off >> 24); o.put(off >> 16); o.put(off >> 8); o.put(off); }
      len -= len1;
    }
  }
}

// Bytes past the end of the block read as 0 (the reference reads its buffer's slack there).
__device__ __forceinline__ uint32_t lz_byte(const uint8_t* in, uint32_t n, uint32_t i) { return i < n ? in[i] : 0u; }

// length of the common prefix of in[p+from..] and in[i+from..], capped (forward match, LZBuffer.cs:264,298,318)
__device__ __forceinline__ uint32_t lz_match_fwd(const uint8_t* in, uint32_t n, uint32_t p, uint32_t i, uint32_t from, uint32_t maxMatch) {
  uint32_t l = from;
  // four byte pairs per step (eight independent loads in flight instead of two), then byte by byte up to the first difference
  const uint32_t lim = min(maxMatch, n - i);                  // l stays below both i + l < n and l < maxMatch
  while (l + 4 <= lim) {
    const bool e0 = in[p + l] == in[i + l], e1 = in[p + l + 1] == in[i + l + 1], e2 = in[p + l + 2] == in[i + l + 2],
               e3 = in[p + l + 3] == in[i + l + 3];
    if (!(e0 && e1 && e2 && e3)) break;
    l += 4;
  }
  while (l < lim && in[p + l] == in[i + l]) ++l;
  return l;
}

__global__ void __launch_bounds__(128) k_lz77(const LzParams P) {
  const int lane = threadIdx.x & 31;
  const uint32_t b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= P.nb) return;
  const uint64_t base = P.off[b] - P.off[0];
  const uint32_t n = (uint32_t)(P.off[b + 1] - P.off[b]);
  const uint8_t* in = P.in + base;
  LzOut o;
  o.p = P.out + P.out_off[b]; o.pos = 0; o.cap = P.out_off[b + 1] - P.out_off[b]; o.bits = o.nbits = 0;

  const uint32_t mask = (1u << P.checkbits) - 1;
  const uint32_t minMatch = P.minMatch, minMatch2 = P.minMatch2, lookahead = P.lookahead, bucket = P.bucket;
  const uint32_t maxMatch = P.maxMatch, maxLiteral = P.maxLiteral;
  uint32_t i = 0, lit = 0, h1 = 0, h2 = 0;

  if (P.use_sa) {
    // ---- suffix-array matcher, LZBuffer.cs:255-283 ----
    const uint32_t* sa = P.sa + base;
    const uint32_t* isa = P.isa + base;
    const uint32_t wmask = ~mask;          // window of the reference's partial ISA (2^checkbits positions)
    int64_t built = -1;                    // window for which the reference's ISA would be valid
    while (i < n) {
      uint32_t blen = minMatch - 1, bp = 0, blit = 0;
      int bscore = 0;
      // the reference rebuilds its windowed ISA when the entry of i is stale (LZBuffer.cs:256-259)
      if (built < 0) { if (sa[0] != i) built = i & wmask; }
      else if ((int64_t)(i & wmask) != built) built = i & wmask;
      // one neighbour of rank q in direction dir at distance k, seen from look-ahead position h (LZBuffer.cs:262-272)
      auto cand = [&](uint32_t h, uint32_t q, int dir, uint32_t k, bool on, uint32_t& p, uint32_t& l, uint32_t& l1) -> bool {
        p = l = l1 = 0;
        const int64_t at = (int64_t)q + (int64_t)dir * (int64_t)k;
        if (!on || k > bucket || at < 0 || at >= (int64_t)n) return false;
        p = sa[at] - h;                                  // unsigned wrap as in the reference when sa < h
        if (!(p < i)) return false;
        l = lz_match_fwd(in, n, p, i, h, maxMatch);
        for (l1 = h; l1 > 0 && in[p + l1 - 1] == in[i + l1 - 1]; --l1) {}
        return true;
      };
      // the reference's sequential scoring and early exit over the candidates of lanes [l0, l0 + cnt), in neighbour order
      // (LZBuffer.cs:273-280); candidates that do not lie in front of i are passed over.  True: the reference stops here.
      auto replay = [&](bool valid, uint32_t p, uint32_t l, uint32_t l1, uint32_t h, int l0, int cnt) -> bool {
        uint32_t vm = __ballot_sync(FULL, valid);
        vm &= cnt >= 32 ? 0xFFFFFFFFu : (((1u << cnt) - 1u) << l0);
        while (vm) {
          const int tl = __ffs((int)vm) - 1;
          vm &= vm - 1;
          const uint32_t cp = __shfl_sync(FULL, p, tl), cl = __shfl_sync(FULL, l, tl), cl1 = __shfl_sync(FULL, l1, tl);
          int score = (int)(cl - cl1) * 8 - lg32(i - cp) - 4 * (lit == 0 && cl1 > 0) - 11;
          for (uint32_t a = 0; a < h; ++a) score = score * 5 / 8;
          if (score > bscore) { blen = cl; bp = cp; blit = cl1; bscore = score; }
          if (cl < blen || cl < minMatch || cl > 255) return true;
        }
        return false;
      };
      // entry of t = h + i in the windowed ISA: valid only inside the built window (or the zero-filled initial array, which
      // maps everything to rank 0); does not depend on what the search has found so far
      auto entry = [&](uint32_t h, uint32_t& q) -> bool {
        const uint32_t t = h + i;
        if (built < 0) { q = 0; return sa[0] == t; }
        if (t >= n || (int64_t)(t & wmask) != built) return false;
        q = isa[t];
        return true;
      };
      if (lookahead <= 1 && bucket >= 16) {
        // Usually the reference stops within a few neighbours (the first one in front of i that matches worse than the best so
        // far ends a direction).  So the nearest neighbours of EVERY (look-ahead position, direction) pair are measured in one
        // warp round -- 32 / pairs lanes each -- and the pairs are replayed in the reference's order; a direction that is not
        // finished by then goes on 32 neighbours at a time.
        const uint32_t npair = 2 * (lookahead + 1), gsz = 32 / npair;
        uint32_t hq[2] = {0, 0};
        bool hon[2] = {false, false};
        for (uint32_t h = 0; h <= lookahead; ++h) hon[h] = entry(h, hq[h]);
        const uint32_t c = (uint32_t)lane / gsz, hc = c >> 1;
        uint32_t p, l, l1;
        const bool valid = cand(hc, hq[hc], (c & 1) ? 1 : -1, (uint32_t)lane % gsz + 1, hon[hc], p, l, l1);
        for (uint32_t h = 0; h <= lookahead; ++h) {
          if (!hon[h]) continue;
          for (uint32_t dd = 0; dd < 2; ++dd) {
            bool brk = replay(valid, p, l, l1, h, (int)((2 * h + dd) * gsz), (int)gsz);
            for (uint32_t k0 = gsz + 1; !brk && k0 <= bucket; k0 += 32) {
              uint32_t p2, l2, l12;
              const bool v2 = cand(h, hq[h], dd ? 1 : -1, k0 + (uint32_t)lane, true, p2, l2, l12);
              brk = replay(v2, p2, l2, l12, h, 0, 32);
            }
          }
          if (bscore <= 0 || blen < minMatch) break;
        }
      } else {
        for (uint32_t h = 0; h <= lookahead; ++h) {
          uint32_t q;
          if (!entry(h, q)) continue;
          for (int dir = -1; dir <= 1; dir += 2) {
            for (uint32_t k0 = 1; k0 <= bucket; k0 += 32) {
              // 32 neighbours at a time: every lane measures one candidate
              uint32_t p, l, l1;
              const bool valid = cand(h, q, dir, k0 + (uint32_t)lane, true, p, l, l1);
              if (replay(valid, p, l, l1, h, 0, 32)) break;
            }
          }
          if (bscore <= 0 || blen < minMatch) break;
        }
      }
      const uint32_t offv = i - bp;
      if (offv > 0 && bscore > 0 && blen - blit >= minMatch + (P.level == 2) * ((offv >= (1u << 16)) + (offv >= (1u << 24)))) {
        lit += blit;
        if (lane == 0) { uint32_t l2 = lit; lz_write_literal(P, o, in, i + blit, l2); lz_write_match(P, o, blen - blit, offv); }
        lit = 0;
      } else { blen = 1; ++lit; }
      i += blen;
      if (lit >= maxLiteral) { if (lane == 0) { uint32_t l2 = lit; lz_write_literal(P, o, in, i, l2); } lit = 0; }
    }
  } else {
    // ---- hash-table matcher, LZBuffer.cs:288-368 ----
    uint32_t* ht = P.ht + (uint64_t)b * P.htsize;
    const uint32_t hmask = P.htsize - 1;
    for (uint32_t k = lane; k < P.htsize; k += 32) ht[k] = 0;
    __syncwarp();
    const bool search = (P.level == 1 || minMatch <= 64);
    while (i < n) {
      uint32_t blen = minMatch - 1, bp = 0, blit = 0;
      int bscore = 0;
      if (search) {
        if (minMatch2 > 0) {
          for (uint32_t k0 = 0; k0 <= bucket; k0 += 32) {
            const uint32_t k = k0 + lane;
            uint32_t e = k <= bucket ? ht[h2 ^ k] : 0;
            const bool hit = e && (e & mask) == (lz_byte(in, n, i + 3) & mask);
            const uint32_t p = e >> P.checkbits;
            uint32_t l = 0; int l1 = 0;
            if (hit && p < i) {
              l = lz_match_fwd(in, n, p, i, lookahead, maxMatch);
              for (l1 = (int)lookahead; l1 > 0 && in[p + l1 - 1] == in[i + l1 - 1]; --l1) {}
            }
            bool brk = false;
            for (int tl = 0; tl < 32 && k0 + tl <= bucket; ++tl) {
              const bool v = __shfl_sync(FULL, hit, tl);
              const uint32_t cp = __shfl_sync(FULL, p, tl), cl = __shfl_sync(FULL, l, tl);
              const int cl1 = __shfl_sync(FULL, l1, tl);
              if (v && cp < i && i + blen <= n && lz_byte(in, n, cp + blen - 1) == lz_byte(in, n, i + blen - 1)) {
                if (cl >= minMatch2 + lookahead) {
                  const int score = (int)(cl - cl1) * 8 - lg32(i - cp) - 8 * (lit == 0 && cl1 > 0) - 11;
                  if (score > bscore) { blen = cl; bp = cp; blit = (uint32_t)cl1; bscore = score; }
                }
              }
              if (blen >= 128) { brk = true; break; }
            }
            if (brk) break;
          }
        }
        if (!minMatch2 || blen < minMatch2) {
          for (uint32_t k0 = 0; k0 <= bucket; k0 += 32) {
            const uint32_t k = k0 + lane;
            uint32_t e = k <= bucket ? ht[h1 ^ k] : 0;
            const bool hit = e && i + 3 < n && (e & mask) == (in[i + 3] & mask);
            const uint32_t p = e >> P.checkbits;
            uint32_t l = 0;
            if (hit && p < i) l = lz_match_fwd(in, n, p, i, 0, maxMatch);
            bool brk = false;
            for (int tl = 0; tl < 32 && k0 + tl <= bucket; ++tl) {
              const bool v = __shfl_sync(FULL, hit, tl);
              const uint32_t cp = __shfl_sync(FULL, p, tl), cl = __shfl_sync(FULL, l, tl);
              if (v && cp < i && i + blen <= n && lz_byte(in, n, cp + blen - 1) == lz_byte(in, n, i + blen - 1)) {
                const int score = (int)cl * 8 - lg32(i - cp) - 2 * (lit > 0) - 11;
                if (score > bscore) { blen = cl; bp = cp; blit = 0; bscore = score; }
              }
              if (blen >= 128) { brk = true; break; }
            }
            if (brk) break;
          }
        }
      }
      const uint32_t offv = i - bp;
      if (offv > 0 && bscore > 0 && blen - blit >= minMatch + (P.level == 2) * ((offv >= (1u << 16)) + (offv >= (1u << 24)))) {
        lit += blit;
        if (lane == 0) { uint32_t l2 = lit; lz_write_literal(P, o, in, i + blit, l2); lz_write_match(P, o, blen - blit, offv); }
        lit = 0;
      } else { blen = 1; ++lit; }
      // index the positions passed (LZBuffer.cs:349-368); sequential because later inserts may
      // replace earlier ones in the same bucket slot
      if (lane == 0) {
        uint32_t ii = i;
        for (uint32_t c = 0; c < blen; ++c, ++ii) {
          if (ii + P.minMatchBoth < n) {
            const uint32_t ih = ((ii * 1234547u) >> 19) & bucket;
            const uint32_t e = (ii << P.checkbits) | (in[ii + 3] & mask);
            if (minMatch2) {
              ht[h2 ^ ih] = e;
              h2 = (((h2 * 9) << P.shift2) + (in[ii + minMatch2 + lookahead] + 1u) * 23456789u) & hmask;
            }
            ht[h1 ^ ih] = e;
            h1 = (((h1 * 5) << P.shift1) + (in[ii + minMatch] + 1u) * 123456791u) & hmask;
          }
        }
      }
      h1 = __shfl_sync(FULL, h1, 0);
      h2 = __shfl_sync(FULL, h2, 0);
      __syncwarp();
      i += blen;
      if (lit >= maxLiteral) { if (lane == 0) { uint32_t l2 = lit; lz_write_literal(P, o, in, i, l2); } lit = 0; }
    }
  }
  if (lane == 0) {
    uint32_t l2 = lit;
    lz_write_literal(P, o, in, n, l2);
    o.flush();
    P.jobs[b].in_len = (uint32_t)(o.pos <= o.cap ? o.pos : 0xFFFFFFFFu);
  }
}

cudaError_t launch_lz77(const LzParams& p, cudaStream_t s) {
  if (!p.nb) return cudaSuccess;
  k_lz77<<<(p.nb + 3) / 4, 128, 0, s>>>(p);
  return cudaGetLastError();
}


// ------------------------------------------------------------------------------------------
// Byte-gap histogram of every block (the data analysis of method levels 5..9, LibZPAQ.cs:242-258):
// gap[b][k] = number of positions i of block b whose byte last occurred at i - k, 0 < k < 4096, where a byte that has not
// occurred yet counts as having occurred at position 0.  grid (blocks, 4096-position chunks): the chunk and the 4095 bytes in
// front of it sit in shared memory, every thread walks back from its positions to the previous occurrence.
// ------------------------------------------------------------------------------------------
constexpr int kGapBinsDev = 4096;
__global__ void __launch_bounds__(256) k_gap_hist(const uint8_t* in, const uint64_t* off, int* gap) {
  __shared__ uint8_t buf[2 * kGapBinsDev];
  __shared__ int hist[kGapBinsDev];
  const uint32_t b = blockIdx.x;
  const uint64_t n = off[b + 1] - off[b];
  const uint8_t* p = in + (off[b] - off[0]);
  for (uint64_t c0 = (uint64_t)blockIdx.y * kGapBinsDev; c0 < n; c0 += (uint64_t)gridDim.y * kGapBinsDev) {
    const uint64_t lo = c0 >= (uint64_t)kGapBinsDev ? c0 - kGapBinsDev : 0;       // first byte held
    const uint64_t hi = min(n, c0 + kGapBinsDev);
    __syncthreads();
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) buf[i - lo] = p[i];
    for (int k = threadIdx.x; k < kGapBinsDev; k += blockDim.x) hist[k] = 0;
    __syncthreads();
    for (uint64_t i = c0 + threadIdx.x; i < hi; i += blockDim.x) {
      const uint32_t at = (uint32_t)(i - lo);
      const uint8_t v = buf[at];
      const uint32_t reach = (uint32_t)min((uint64_t)(kGapBinsDev - 1), i);      // gaps 1 .. reach can be seen
      uint32_t k = 1;
      while (k <= reach && buf[at - k] != v) ++k;
      if (k <= reach) atomicAdd(&hist[k], 1);
      else if (i > 0 && i < (uint64_t)kGapBinsDev) atomicAdd(&hist[(uint32_t)i], 1);   // never seen: "last seen at 0"
    }
    __syncthreads();
    for (int k = threadIdx.x; k < kGapBinsDev; k += blockDim.x)
      if (hist[k]) atomicAdd(&gap[(uint64_t)b * kGapBinsDev + k], hist[k]);
  }
}
cudaError_t launch_gap_hist(const uint8_t* in, const uint64_t* d_off, uint32_t nb, uint64_t max_len, int* gap, cudaStream_t s) {
  if (!nb) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(gap, 0, (size_t)nb * kGapBinsDev * sizeof(int), s);
  if (e != cudaSuccess) return e;
  const uint32_t chunks = (uint32_t)std::min<uint64_t>(65535, (max_len + kGapBinsDev - 1) / kGapBinsDev);
  if (!chunks) return cudaSuccess;
  k_gap_hist<<<dim3(nb, chunks), 256, 0, s>>>(in, d_off, gap);
  return cudaGetLastError();
}

}  // namespace zpq
