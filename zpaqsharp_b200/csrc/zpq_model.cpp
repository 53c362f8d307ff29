// zpq_model.cpp -- host side: block header parsing, device plan construction, read-only tables,
// archive block framing.  Nothing here touches block data; it prepares < 100 KB of metadata per
// batch.
//
// Reference behaviour mirrored (file:line under /root/reference/ZPAQSharp):
//   parse_header   ZPAQL.read            ZPAQL.cs:112-156
//   header_memory  ZPAQL.memory          ZPAQL.cs:58-81
//   build_plan     Predictor.init        Predictor.cs:84-171 (sizes, limits, initial values)
//   build_tables   Predictor.init        Predictor.cs:54-78, StateTable.cs:21-162
//   parse_block    Decompresser          Decompresser.cs:29-108,163-194; Decoder.skip Decoder.cs:70-98
#include "zpq_host.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace zpq {

const uint8_t kLocatorTag[13] = {0x37, 0x6b, 0x53, 0x74, 0xa0, 0x31, 0x83, 0xd3, 0x8c, 0xb2, 0x28, 0xb0, 0xd3};

static inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

size_t parse_header(const uint8_t* p, size_t avail, Header& h) {
  auto need = [&](size_t k) { if (k > avail) throw Failure(ZPQ_E_CORRUPT, "unexpected end of file"); };
  need(7);
  const size_t hsize = p[0] + 256u * p[1];
  need(hsize + 2);
  h.hh = p[2]; h.hm = p[3]; h.ph = p[4]; h.pm = p[5]; h.n = p[6];
  size_t pos = 7;
  for (int i = 0; i < h.n; ++i) {
    need(pos + 1);
    int len = comp_len(p[pos]);
    if (len < 1) throw Failure(ZPQ_E_CORRUPT, "Invalid component type");
    if (pos + len > hsize) throw Failure(ZPQ_E_CORRUPT, "COMP overflows header");
    pos += len;
  }
  need(pos + 1);
  if (p[pos++] != 0) throw Failure(ZPQ_E_CORRUPT, "missing COMP END");
  h.cend = (int)pos;
  if (pos > hsize + 1) throw Failure(ZPQ_E_CORRUPT, "missing HCOMP");
  const size_t total = hsize + 2;
  if (p[total - 1] != 0) throw Failure(ZPQ_E_CORRUPT, "missing HCOMP END");
  h.wire.assign(p, p + total);
  return total;
}

double header_memory(const Header& h) {
  auto p2 = [](int x) { return std::ldexp(1.0, x); };
  // header.Length in the reference is hsize+300 (ZPAQL.cs:117)
  double mem = p2(h.hh + 2) + p2(h.hm) + p2(h.ph + 2) + p2(h.pm) + (double)(h.wire.size() - 2 + 300);
  size_t cp = 7;
  for (int i = 0; i < h.n; ++i) {
    const uint8_t* c = &h.wire[cp];
    double size = p2(c[1]);
    switch (c[0]) {
      case C_CM: mem += 4 * size; break;
      case C_ICM: mem += 64 * size + 1024; break;
      case C_MATCH: mem += 4 * size + p2(c[2]); break;
      case C_MIX2: mem += 2 * size; break;
      case C_MIX: mem += 4 * size * c[3]; break;
      case C_ISSE: mem += 64 * size + 2048; break;
      case C_SSE: mem += 128 * size; break;
    }
    cp += comp_len(c[0]);
  }
  return mem;
}

// ------------------------------------------------------------------------------------------
void build_tables(Tables& t) {
  for (int i = 0; i < 4096; ++i) {
    double v = 32768.0 / (1 + std::exp((i - 2048) * (-1.0 / 64)));
    t.squash[i] = i < 1376 ? 0 : i >= 2720 ? 32767 : (uint16_t)(int)v;
  }
  for (int i = 0; i < 32768; ++i)
    t.stretch[i] = (int16_t)((int)(std::log((i + 0.5) / (32767.5 - i)) * 64 + 0.5 + 100000) - 100000);
  uint32_t st = 0, sq = 0;  // the reference's own self-check, Predictor.cs:69-78
  for (int i = 32767; i >= 0; --i) st = st * 3 + (uint32_t)(int)t.stretch[i];
  for (int i = 4095; i >= 0; --i) sq = sq * 3 + t.squash[i];
  if (st != 3887533746u || sq != 2278286169u) throw Failure(ZPQ_E_CONFIG, "squash/stretch tables failed their checksum");
  for (int i = 0; i < 1024; ++i) t.dt[i] = (1 << 17) / (i * 2 + 3) * 2;
  t.dt2k[0] = 0;
  for (int i = 1; i < 256; ++i) t.dt2k[i] = (uint16_t)(2048 / i);

  // Bit-history state machine.  States are (n0, n1) count pairs, some in two flavours that
  // remember the last bit; counts are bounded and the opposite count decays on each update.
  struct Gen {
    static int flavours(int a, int b) {
      static const int cap[6] = {20, 48, 15, 8, 6, 5};
      if (a < b) std::swap(a, b);
      if (b < 0 || b > 5 || a > cap[b]) return 0;
      return (b > 0 && a + b <= 17) ? 2 : 1;
    }
    static int fade(int v) { return v < 5 ? v : v < 7 ? 5 : v < 8 ? 6 : 7; }
    static void observe(int& a, int& b, int bit) {  // a = count of zeros, b = count of ones
      if (a < b) { observe(b, a, bit ^ 1); return; }
      if (bit) { ++b; a = fade(a); } else { ++a; b = fade(b); }
      while (!flavours(a, b)) {
        if (b < 2) --a;
        else { a = (a * (b - 1) + b / 2) / b; --b; }
      }
    }
  };
  static uint8_t code[50][50][2];
  memset(code, 0, sizeof(code));
  int count = 0;
  for (int sum = 0; sum < 50; ++sum)
    for (int b = 0; b <= sum; ++b) {
      int a = sum - b, f = Gen::flavours(a, b);
      if (!f) continue;
      code[a][b][0] = (uint8_t)count;
      code[a][b][1] = (uint8_t)(count + f - 1);
      count += f;
    }
  memset(t.ns, 0, sizeof(t.ns));
  for (int a = 0; a < 50; ++a)
    for (int b = 0; b < 50; ++b)
      for (int f = 0; f < Gen::flavours(a, b); ++f) {
        uint8_t* row = &t.ns[code[a][b][f] * 4];
        int x = a, y = b;
        Gen::observe(x, y, 0); row[0] = code[x][y][0];
        x = a; y = b;
        Gen::observe(x, y, 1); row[1] = code[x][y][1];
        row[2] = (uint8_t)a; row[3] = (uint8_t)b;
      }
  auto clamp512k = [](int x) { return x < -(1 << 19) ? -(1 << 19) : x >= (1 << 19) ? (1 << 19) - 1 : x; };
  for (int j = 0; j < 256; ++j) {
    int init = ((t.ns[j * 4 + 3] * 2 + 1) << 22) / (t.ns[j * 4 + 2] + t.ns[j * 4 + 3] + 1);  // StateTable.cs:158
    t.icm_init[j] = (uint32_t)init;
    t.isse_init[j * 2] = 1 << 15;
    t.isse_init[j * 2 + 1] = (uint32_t)clamp512k(t.stretch[init >> 8] * 1024);
  }
  for (int j = 0; j < 32; ++j) t.sse_init[j] = (uint32_t)t.squash[j * 64 - 992 + 2048] << 17;
}

// ------------------------------------------------------------------------------------------
void build_plan(const Header& h, bool for_decode, uint32_t smem_budget, Plan& pl, int duo_g, bool fdec, uint64_t max_stream) {
  memset(&pl, 0, sizeof(Plan) - sizeof(pl.hcomp));
  pl.n = h.n; pl.hh = h.hh; pl.hm = h.hm; pl.ph = h.ph; pl.pm = h.pm;
  if (h.hh > 28 || h.hm > 30 || h.ph > 28 || h.pm > 30)
    throw Failure(ZPQ_E_UNSUPPORTED, "H/M array larger than the device build supports");
  pl.hcomp_len = (int)h.wire.size() - h.cend;
  memcpy(pl.hcomp, &h.wire[h.cend], pl.hcomp_len);
  memset(pl.hcomp + pl.hcomp_len, 0, 8);

  uint64_t arena = 0;
  auto take = [&](uint64_t bytes) { uint64_t o = arena; arena = align_up(arena + bytes, 256); return o; };
  int ninit = 0;
  // role: which warp of the two-role encoder owns (and initialises) the table -- 0 lead, 1 coder
  auto fill = [&](uint64_t dst, uint64_t bytes, uint8_t kind, uint32_t value, bool smem, uint8_t role = 0) {
    InitOp& op = pl.init[ninit++];
    op.dst = dst; op.bytes = align_up(bytes, 16); op.kind = kind; op.value = value; op.to_smem = smem; op.role = role; op.pad = 0;
  };

  // shared slice: p[n], state[n] (5 words each), then optionally H and the small cm tables
  uint32_t slice = 0;
  auto stake = [&](uint32_t bytes) { uint32_t o = slice; slice = (uint32_t)align_up(slice + bytes, 16); return o; };
  const int nn = std::max(h.n, 1);
  pl.duo_g = duo_g;
  pl.fd_layout = fdec ? 1 : 0;
  pl.smem_fd_cm = kNoSmem;
  if (fdec) {
    pl.smem_rows = stake(1024);   // 32 lanes x 16-byte hash row, twice (first nibble / candidates of the second)
    pl.smem_chain = stake(512);   // 32 slots of {x0, y0, x1, y1}
    const uint8_t* q = &h.wire[7];
    for (int i = 0; i < h.n; ++i) { if (q[0] == C_CM && pl.smem_fd_cm == kNoSmem) pl.smem_fd_cm = stake(512); q += comp_len(q[0]); }
  } else if (duo_g) {
    pl.smem_sync = stake(48);
    pl.smem_pfring = stake(128);
    pl.smem_rows = stake(32u * duo_g);   // duo_g lanes x 16-byte row cache, twice
  } else {
    pl.smem_p = stake(4u * nn);
    pl.smem_st = stake(20u * nn);
    pl.smem_rows = stake(1024);   // 32 lanes x 16-byte row cache, twice (the skewed encoder requests rows a nibble ahead)
    pl.smem_chain = stake(512);
  }
  // Pipelined encoder: component i works delay[i] bits behind the leading bit, strictly later than
  // everything it reads; a MIX additionally stays kPipeMixAhead bits back so that its weight rows
  // can be requested before they are needed.  The coder follows component n-1 by one bit.
  uint8_t delay[kMaxComp + 1] = {0};
  {
    const uint8_t* q = &h.wire[7];
    for (int i = 0; i < h.n; ++i) {
      int dl = 0;
      auto in = [&](int j) { if (j < i) dl = std::max(dl, delay[j] + 1); };
      switch (q[0]) {
        case C_AVG: in(q[1]); in(q[2]); break;
        case C_MIX2: in(q[2]); in(q[3]); break;
        case C_ISSE: case C_SSE: in(q[2]); break;
        case C_MIX: for (int j = 0; j < q[3]; ++j) in(q[2] + j); dl = std::max(dl, duo_g ? kDuoMixAhead : kPipeMixAhead); break;
        default: break;
      }
      delay[i] = (uint8_t)std::min(dl, 255);
      q += comp_len(q[0]);
    }
    pl.coder_delay = h.n ? delay[h.n - 1] + 1 : 1;
    pl.ring_slots = 8;
    while ((int)pl.ring_slots <= pl.coder_delay) pl.ring_slots *= 2;
    pl.ring_stride = (uint32_t)align_up(nn, 8);
    pl.pipe_ok = h.n >= 1 && h.n <= 32 && pl.coder_delay <= kMaxPipeDelay;
    // two-role encoder: predictions of lead-role components (CONS/CM/ICM/MATCH) are read from the ring at any
    // distance; a prediction made by the coder role is read from its lane's history of the last 8 bits
    {
      auto lead_type = [&](int t) { return t == C_CONS || t == C_CM || t == C_MATCH; };   // (an ICM's map is trained by the coder role)
      std::vector<int> types(h.n);
      const uint8_t* q2 = &h.wire[7];
      for (int i = 0; i < h.n; ++i) { types[i] = q2[0]; q2 += comp_len(q2[0]); }
      int hdepth = 1, ldepth = 1;
      q2 = &h.wire[7];
      for (int i = 0; i < h.n; ++i) {
        auto dep = [&](int j, bool lane_owned) {
          if (j >= i || lead_type(types[j])) return;
          hdepth = std::max(hdepth, delay[i] - delay[j]);
          if (lane_owned) ldepth = std::max(ldepth, delay[i] - delay[j]);
        };
        switch (q2[0]) {
          case C_AVG: dep(q2[1], true); dep(q2[2], true); break;
          case C_MIX2: dep(q2[2], true); dep(q2[3], true); break;
          case C_ISSE: case C_SSE: dep(q2[2], true); break;
          case C_MIX: for (int j = 0; j < q2[3]; ++j) dep(q2[2] + j, false); break;
          default: break;
        }
        q2 += comp_len(q2[0]);
      }
      pl.duo_hdepth = hdepth; pl.duo_ldepth = ldepth;
      // the MIX components get a role warp of their own when nothing lane-owned reads a MIX output
      bool any_mix = false, mix_read = false;
      q2 = &h.wire[7];
      for (int i = 0; i < h.n; ++i) {
        auto rd = [&](int j) { if (j < i && types[j] == C_MIX) mix_read = true; };
        switch (q2[0]) {
          case C_AVG: rd(q2[1]); rd(q2[2]); break;
          case C_MIX2: rd(q2[2]); rd(q2[3]); break;
          case C_ISSE: case C_SSE: rd(q2[2]); break;
          case C_MIX: any_mix = true; break;
          default: break;
        }
        q2 += comp_len(q2[0]);
      }
      pl.duo_split = duo_g && any_mix && !mix_read;
      pl.duo_ok = duo_g && pl.pipe_ok && h.n <= duo_g && hdepth <= 8;
    }
    if (duo_g) {
      pl.ring_slots = 64; pl.ring_stride = (uint32_t)duo_g;
      pl.smem_pring = stake(2u * 64 * duo_g);
      pl.smem_hsnap = stake(4u * 8 * duo_g);
    } else if (pl.pipe_ok && !fdec) {
      pl.smem_pring = stake(2u * pl.ring_slots * pl.ring_stride);
      pl.smem_bhring = stake(pl.ring_slots * pl.ring_stride);
      pl.smem_hsnap = stake(4u * 8 * pl.ring_stride);
    }
  }
  const uint64_t hbytes = 4ull << h.hh;
  if (h.n > 0 && hbytes <= 2048 && slice + hbytes <= smem_budget) { pl.smem_h = stake((uint32_t)hbytes); fill(pl.smem_h, hbytes, 0, 0, true); }
  else { pl.smem_h = kNoSmem; pl.off_h = take(hbytes); fill(pl.off_h, hbytes, 0, 0, false); }
  const uint64_t mbytes = 1ull << h.hm;
  if (h.n > 0 && mbytes <= 1024 && slice + mbytes <= smem_budget) { pl.smem_m = stake((uint32_t)std::max<uint64_t>(mbytes, 16)); fill(pl.smem_m, mbytes, 0, 0, true); }
  else { pl.smem_m = kNoSmem; pl.off_m = take(mbytes); fill(pl.off_m, mbytes, 0, 0, false); }
  pl.off_r = take(1024); fill(pl.off_r, 1024, 0, 0, false);
  pl.smem_cm = slice;

  const uint8_t* cp = &h.wire[7];
  for (int i = 0; i < h.n; ++i) {
    CompDesc& d = pl.comp[i];
    d.type = cp[0];
    const int len = comp_len(cp[0]);
    for (int k = 1; k < len && k <= 5; ++k) d.a[k - 1] = cp[k];
    d.smem_cm = kNoSmem;
    const int bits = cp[1];
    auto limit = [&](int maxbits, const char* what) {
      if (bits > 32) throw Failure(ZPQ_E_CONFIG, what);
      if (bits > maxbits) throw Failure(ZPQ_E_UNSUPPORTED, "component table larger than the device build supports");
    };
    auto lvl = [&](int j) { return (int)pl.comp[j].level; };
    int level = 0;
    switch (cp[0]) {
      case C_CONS: break;
      case C_CM:
        limit(29, "max size for CM is 32");
        d.mask = (1u << bits) - 1;
        d.tab = take(4ull << bits); fill(d.tab, 4ull << bits, 0, 0x80000000u, false, 3);
        break;
      case C_ICM:
        if (bits > 26) throw Failure(ZPQ_E_CONFIG, "max size for ICM is 26");
        d.mask = (uint32_t)((64ull << bits) - 16);
        d.tab = take(64ull << bits); fill(d.tab, 64ull << bits, 0, 0, false, 3);
        if (duo_g || fdec) {   // the coder role (zpq_duo.cuh) / the speculative decoder (zpq_fdec.cuh) address ICM maps with a 4-byte stride
          if (slice + 1024 <= smem_budget) { d.smem_cm = stake(1024); fill(d.smem_cm, 1024, 4, 0, true, 1); }
          else { d.tab2 = take(1024); fill(d.tab2, 1024, 4, 0, false, 1); }
        } else if (slice + 2048 <= smem_budget) { d.smem_cm = stake(2048); fill(d.smem_cm, 2048, 1, 0, true); }
        else { d.tab2 = take(2048); fill(d.tab2, 2048, 1, 0, false); }
        break;
      case C_MATCH:
        if (bits > 32 || cp[2] > 32) throw Failure(ZPQ_E_CONFIG, "max size for MATCH is 32 32");
        if (bits > 29 || cp[2] > 31) throw Failure(ZPQ_E_UNSUPPORTED, "MATCH larger than the device build supports");
        {
          // The buffer is indexed by the stream position (Predictor.cs:382-411: limit counts bytes and wraps at 2^cp[2]) and by
          // positions up to 256 + limit bytes in front of it, which wrap to the unwritten, zeroed top of the buffer.  When no
          // stream of the launch reaches 2^k - 1024 bytes, a 2^k-byte buffer therefore holds the same bytes at every index the
          // component can form: it is allocated (and cleared) at that size instead of 2^cp[2] (mid.cfg: 2 MB instead of 16 MB).
          int bb = cp[2];
          if (max_stream) {
            int need = 12;
            while ((1ull << need) < max_stream + 1024) ++need;
            bb = std::min(bb, need);
          }
          d.mask = (1u << bits) - 1; d.mask2 = (uint32_t)((1ull << bb) - 1);
          d.tab = take(4ull << bits); fill(d.tab, 4ull << bits, 0, 0, false);
          d.tab2 = take(1ull << bb); fill(d.tab2, 1ull << bb, 0, 0, false);
        }
        break;
      case C_AVG:
        if (cp[1] >= i) throw Failure(ZPQ_E_CONFIG, "AVG j >= i");
        if (cp[2] >= i) throw Failure(ZPQ_E_CONFIG, "AVG k >= i");
        level = 1 + std::max(lvl(cp[1]), lvl(cp[2]));
        break;
      case C_MIX2:
        limit(30, "max size for MIX2 is 32");
        if (cp[3] >= i) throw Failure(ZPQ_E_CONFIG, "MIX2 k >= i");
        if (cp[2] >= i) throw Failure(ZPQ_E_CONFIG, "MIX2 j >= i");
        d.mask = (1u << bits) - 1;
        d.tab = take(2ull << bits); fill(d.tab, 2ull << bits, 0, 0x80008000u, false, 1);
        level = 1 + std::max(lvl(cp[2]), lvl(cp[3]));
        break;
      case C_MIX: {
        limit(24, "max size for MIX is 32");
        if (cp[2] >= i) throw Failure(ZPQ_E_CONFIG, "MIX j >= i");
        if (cp[3] < 1 || cp[3] > i - cp[2]) throw Failure(ZPQ_E_CONFIG, "MIX m not in 1..i-j");
        const int m = cp[3];
        d.mask = (1u << bits) - 1;
        d.coop = 1;
        d.tab = take((4ull * m) << bits); fill(d.tab, (4ull * m) << bits, 0, (uint32_t)(65536 / m), false, pl.duo_split ? 2 : 1);
        for (int j = 0; j < m; ++j) level = std::max(level, 1 + lvl(cp[2] + j));
        break;
      }
      case C_ISSE:
        limit(26, "max size for ISSE is 32");
        if (cp[2] >= i) throw Failure(ZPQ_E_CONFIG, "ISSE j >= i");
        d.mask = (uint32_t)((64ull << bits) - 16);
        d.tab = take(64ull << bits); fill(d.tab, 64ull << bits, 0, 0, false, 3);
        if (slice + 2048 <= smem_budget) { d.smem_cm = stake(2048); fill(d.smem_cm, 2048, 2, 0, true, 1); }
        else { d.tab2 = take(2048); fill(d.tab2, 2048, 2, 0, false, 1); }
        level = 1 + lvl(cp[2]);
        break;
      case C_SSE:
        limit(24, "max size for SSE is 32");
        if (cp[2] >= i) throw Failure(ZPQ_E_CONFIG, "SSE j >= i");
        if (cp[3] > cp[4] * 4) throw Failure(ZPQ_E_CONFIG, "SSE start > limit*4");
        d.mask = (uint32_t)((32ull << bits) - 1);
        d.tab = take(128ull << bits); fill(d.tab, 128ull << bits, 3, cp[3], false, 1);
        level = 1 + lvl(cp[2]);
        break;
      default: throw Failure(ZPQ_E_CONFIG, "unknown component type");
    }
    d.level = (uint8_t)level;
    d.delay = delay[i];
    cp += len;
  }

  if (for_decode) {
    pl.off_ph = take(4ull << h.ph); fill(pl.off_ph, 4ull << h.ph, 0, 0, false);
    pl.off_pm = take(1ull << h.pm); fill(pl.off_pm, 1ull << h.pm, 0, 0, false);
    pl.off_pr = take(1024); fill(pl.off_pr, 1024, 0, 0, false);
    pl.off_pcode = take(65536 + 512);
  }
  pl.ninit = ninit;
  pl.arena_bytes = align_up(arena, 4096);
  pl.smem_warp_bytes = (uint32_t)align_up(slice, 128);

  // Evaluation schedule: by level; inside a level lane-owned components first (<= 32 per step),
  // then every cooperative component on its own.
  int maxlevel = 0;
  for (int i = 0; i < h.n; ++i) maxlevel = std::max(maxlevel, (int)pl.comp[i].level);
  int no = 0, ns = 0, nu = 0;
  for (int L = 0; L <= maxlevel; ++L) {
    int start = no;
    for (int i = 0; i < h.n; ++i)
      if (pl.comp[i].level == L && !pl.comp[i].coop && pl.comp[i].type != C_CONS) pl.order[no++] = (uint8_t)i;
    for (int s = start; s < no; s += 32) {
      Step& st = pl.steps[ns++];
      st.first = (uint16_t)s; st.count = (uint8_t)std::min(32, no - s); st.coop = 0;
    }
    for (int i = 0; i < h.n; ++i)
      if (pl.comp[i].level == L && pl.comp[i].coop) {
        Step& st = pl.steps[ns++];
        st.first = (uint16_t)no; st.count = 1; st.coop = 1;
        pl.order[no++] = (uint8_t)i;
      }
  }
  pl.nsteps = ns;
  pl.maxlevel = maxlevel;
  // the time-skewed encoder addresses the ICM/ISSE maps as shared memory
  pl.pipe_maps = 1;
  for (int i = 0; i < h.n; ++i)
    if ((pl.comp[i].type == C_ICM || pl.comp[i].type == C_ISSE) && pl.comp[i].smem_cm == kNoSmem) pl.pipe_maps = 0;
  pl.nmix = 0;
  pl.lane_ok = h.n >= 1 && h.n <= 32;
  for (int i = 0; i < h.n; ++i) {
    const CompDesc& d = pl.comp[i];
    if (d.type != C_MIX) continue;
    if (pl.nmix == kMaxMix) { pl.lane_ok = 0; break; }
    MixDesc& m = pl.mix[pl.nmix++];
    m.lane = (uint8_t)i; m.level = d.level; m.j0 = d.a[1]; m.m = d.a[2]; m.rate = d.a[3]; m.cmask = d.a[4];
    m.pad = 0; m.pad2 = 0; m.mask = d.mask; m.tab = d.tab;
  }
  for (int i = 0; i < h.n; ++i) {
    const int t = pl.comp[i].type;
    if (t == C_CM || t == C_ICM || t == C_MATCH || t == C_MIX2 || t == C_ISSE || t == C_SSE) pl.upd[nu++] = (uint8_t)i;
  }
  pl.nupd = nu;
}

// ------------------------------------------------------------------------------------------
// Archive framing (host).
// ------------------------------------------------------------------------------------------
void parse_block(const uint8_t* p, uint64_t avail, BlockRef& out) {
  uint64_t pos = 0;
  auto need = [&](uint64_t k) { if (pos + k > avail) throw Failure(ZPQ_E_CORRUPT, "unexpected EOF"); };
  if (avail >= 13 && memcmp(p, kLocatorTag, 13) == 0) pos = 13;
  need(5);
  if (p[pos] != 'z' || p[pos + 1] != 'P' || p[pos + 2] != 'Q') throw Failure(ZPQ_E_CORRUPT, "block does not start with zPQ");
  const int level = p[pos + 3];
  if (level != 1 && level != 2) throw Failure(ZPQ_E_CORRUPT, "unsupported ZPAQ level");
  if (p[pos + 4] != 1) throw Failure(ZPQ_E_CORRUPT, "unsupported ZPAQL type");
  pos += 5;
  pos += parse_header(p + pos, avail - pos, out.hdr);
  if (level == 1 && out.hdr.n == 0) throw Failure(ZPQ_E_CORRUPT, "ZPAQ level 1 requires at least 1 component");
  out.segs.clear();
  for (;;) {
    need(1);
    int c = p[pos++];
    if (c == 255) break;
    if (c != 1) throw Failure(ZPQ_E_CORRUPT, "missing segment or end of block");
    while (true) { need(1); if (p[pos++] == 0) break; }      // filename
    SegmentRef seg;
    seg.size_hint = -1;
    {
      uint64_t cstart = pos;
      while (true) { need(1); if (p[pos++] == 0) break; }    // comment
      int64_t v = 0; bool any = false;
      for (uint64_t k = cstart; k + 1 < pos && p[k] >= '0' && p[k] <= '9' && v < (1ll << 40); ++k) { v = v * 10 + (p[k] - '0'); any = true; }
      if (any) seg.size_hint = v;
    }
    need(1);
    if (p[pos++] != 0) throw Failure(ZPQ_E_CORRUPT, "missing reserved byte");
    seg.data_off = pos;
    if (out.hdr.n > 0) {
      // coded data ends with the first run of >= 4 zero bytes (Decoder.skip)
      // (a run of four zeros contains one of every four consecutive positions: probe every fourth byte)
      uint64_t k = pos + 3, end4 = 0;
      while (k < avail) {
        if (p[k] != 0) { k += 4; continue; }
        uint64_t lo = k, hi = k;
        while (lo > pos && p[lo - 1] == 0) --lo;
        while (hi + 1 < avail && p[hi + 1] == 0) ++hi;
        if (hi - lo + 1 >= 4) { end4 = lo + 4; break; }
        k = hi + 4;
      }
      if (!end4) throw Failure(ZPQ_E_CORRUPT, "unexpected EOF");
      pos = end4;
      while (pos < avail && p[pos] == 0) ++pos;
    } else {
      // stored: [len32 BE][bytes]... terminated by a zero length (Decoder.cs:56-66)
      while (true) {
        need(4);
        uint64_t len = (uint64_t)p[pos] << 24 | (uint64_t)p[pos + 1] << 16 | (uint64_t)p[pos + 2] << 8 | p[pos + 3];
        pos += 4;
        if (len == 0) break;
        need(len);
        pos += len;
      }
    }
    seg.data_len = pos - seg.data_off;
    need(1);
    c = p[pos++];
    seg.has_sha1 = false;
    if (c == 253) { need(20); memcpy(seg.sha1, p + pos, 20); pos += 20; seg.has_sha1 = true; }
    else if (c != 254) throw Failure(ZPQ_E_CORRUPT, "missing end of segment marker");
    out.segs.push_back(seg);
  }
  out.end = pos;
}

int64_t find_blocks(const uint8_t* p, uint64_t n, uint64_t* offsets, uint64_t max_blocks) {
  int64_t found = 0;
  uint64_t pos = 0;
  while (pos < n) {
    uint32_t h1 = 0x3D49B113, h2 = 0x29EB7F93, h3 = 0x2614BE13, h4 = 0x3828EB13;  // Decompresser.cs:34
    bool hit = false;
    for (; pos < n; ++pos) {
      const uint32_t c = p[pos];
      h1 = h1 * 12 + c; h2 = h2 * 20 + c; h3 = h3 * 28 + c; h4 = h4 * 44 + c;
      if (h1 == 0xB16B88F1 && h2 == 0xFF5376F1 && h3 == 0x72AC5BF1 && h4 == 0x2F909AF1) { hit = true; ++pos; break; }
    }
    if (!hit) break;
    const uint64_t start = pos - 3;  // the 'z' of "zPQ"
    BlockRef b;
    parse_block(p + start, n - start, b);
    if ((uint64_t)found < max_blocks && offsets) { offsets[2 * found] = start; offsets[2 * found + 1] = start + b.end; }
    ++found;
    pos = start + b.end;
  }
  return found;
}

}  // namespace zpq
