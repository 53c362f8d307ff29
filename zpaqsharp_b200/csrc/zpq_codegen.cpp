// zpq_codegen.cpp -- model specialisation: turns one block header into CUDA source for the
// lane-resident kernels.  The reference does the same thing for x86 (ZPAQL.assemble,
// ZPAQL.cs:353-1008; Predictor.assemble_p, Predictor.cs:579-1356): translate the component list
// and the HCOMP program once per model into straight-line native code.  Here the target is
// sm_100a: the text produced below is compiled ahead of time for the three built-in models
// (zpq_gen tool, run by build.py) and at run time through NVRTC for every other header.
//
// What gets fixed at compile time: which component types exist and on which lanes, the dependency
// levels (one SHFL + one predicated evaluation per level instead of a run-time walk), MIX shapes
// (weights held in registers, next rows prefetched) and the HCOMP program (ZPAQL -> C with
// gotos, executed uniformly by all lanes).  Table offsets stay run-time (they depend on the
// shared-memory budget of the launch) and reach the lanes' registers through lane_load().
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <sstream>

#include "zpq_host.h"

namespace zpq {

namespace {

const char* kReg[4] = {"a", "b", "c", "d"};

std::string M_(const char* idx) { return std::string("M[(") + idx + ") & MMASK]"; }
std::string H_(const char* idx) { return std::string("H[(") + idx + ") & HMASK]"; }

// rvalue of source field sss (7 = immediate n)
std::string src_expr(int s, int n) {
  switch (s) {
    case 0: case 1: case 2: case 3: return kReg[s];
    case 4: return "(uint32_t)" + M_("b");
    case 5: return "(uint32_t)" + M_("c");
    case 6: return H_("d");
    default: return std::to_string(n) + "u";
  }
}
// statement assigning expression e to destination field ddd
std::string store_stmt(int ddd, const std::string& e) {
  switch (ddd) {
    case 0: case 1: case 2: case 3: return std::string(kReg[ddd]) + " = " + e + ";";
    case 4: return M_("b") + " = (uint8_t)(" + e + ");";
    case 5: return M_("c") + " = (uint8_t)(" + e + ");";
    default: return H_("d") + " = " + e + ";";
  }
}

// Translate a ZPAQL program (incl. END byte) to the body of a C function.  Returns false when
// the program cannot be compiled faithfully (jump into the middle of an instruction, undefined
// opcode on a reachable path is still fine: it becomes `goto Lerr`).
bool translate_zpaql(const uint8_t* code, int len, std::ostringstream& o, bool with_out = false) {
  // linear sweep: instruction starts
  std::vector<int> start(len + 1, 0);
  std::vector<int> ilen(len + 1, 0);
  for (int pc = 0; pc < len;) {
    const int op = code[pc];
    const int l = op == 255 ? 3 : (op & 7) == 7 ? 2 : 1;
    if (pc + l > len) break;   // trailing partial instruction: never reached in a well formed program
    start[pc] = 1; ilen[pc] = l;
    pc += l;
  }
  std::set<int> targets;
  for (int pc = 0; pc < len; ++pc) {
    if (!start[pc]) continue;
    const int op = code[pc];
    int t = -1;
    if (op == 39 || op == 47 || op == 63) t = pc + 2 + (((code[pc + 1] + 128) & 255) - 128);
    else if (op == 255) t = code[pc + 1] + 256 * code[pc + 2];
    if (t >= 0) {
      if (t < len && !start[t]) return false;
      targets.insert(t);
    }
  }
  auto jump = [&](int from, int t) {
    std::string g;
    if (t < 0 || t >= len) return std::string("goto Lerr;");
    if (t <= from) g = "if (--budget == 0) goto Lerr; ";
    return g + "goto L" + std::to_string(t) + ";";
  };
  static const char* bin[14] = {"+", "-", "*", "/", "%", "&", "&~", "|", "^", "<<", ">>", "==", "<", ">"};
  for (int pc = 0; pc < len; ++pc) {
    if (!start[pc]) continue;
    if (targets.count(pc)) o << "  L" << pc << ":;\n";
    const int op = code[pc], n = ilen[pc] > 1 ? code[pc + 1] : 0;
    o << "  ";
    if (op < 64) {
      const int ddd = op >> 3, x = op & 7;
      if (ddd == 7) {
        if (x == 0) o << "goto Lhalt;";
        else if (x == 1) { if (with_out) o << "{ if (opos < ocap) outp[opos] = (uint8_t)a; ++opos; }"; else o << "/* out: no destination in HCOMP */;"; }
        else if (x == 3) o << "a = (a + " << M_("b") << " + 512u) * 773u;";
        else if (x == 4) o << H_("d") << " = (" << H_("d") << " + a + 512u) * 773u;";
        else if (x == 7) o << jump(pc, pc + 2 + (((n + 128) & 255) - 128));
        else o << "goto Lerr;";
      } else if (x == 7) {
        if (ddd < 4) o << kReg[ddd] << " = R[" << n << "];";
        else if (ddd == 4) o << "if (f) { " << jump(pc, pc + 2 + (((n + 128) & 255) - 128)) << " }";
        else if (ddd == 5) o << "if (!f) { " << jump(pc, pc + 2 + (((n + 128) & 255) - 128)) << " }";
        else o << "R[" << n << "] = a;";
      } else if (x > 4 || op == 0) o << "goto Lerr;";
      else {
        const std::string v = src_expr(ddd, 0);
        if (x == 0) {  // swap with a; byte cells exchange the low 8 bits only
          if (ddd == 0) o << ";";
          else if (ddd == 4 || ddd == 5) o << "{ const uint32_t t = " << v << "; " << store_stmt(ddd, "a") << " a = (a & ~255u) | t; }";
          else o << "{ const uint32_t t = " << v << "; " << store_stmt(ddd, "a") << " a = t; }";
        } else if (x == 1) o << store_stmt(ddd, v + " + 1u");
        else if (x == 2) o << store_stmt(ddd, v + " - 1u");
        else if (x == 3) o << store_stmt(ddd, "~" + v);
        else o << store_stmt(ddd, "0u");
      }
    } else if (op < 128) {
      const int ddd = (op >> 3) & 7;
      if (ddd == 7) o << "goto Lerr;";
      else o << store_stmt(ddd, src_expr(op & 7, n));
    } else if (op == 255) {
      o << jump(pc, code[pc + 1] + 256 * code[pc + 2]);
    } else {
      const int x = (op >> 3) & 15;
      const std::string v = src_expr(op & 7, n);
      if (x > 13) o << "goto Lerr;";
      else if (x == 3) o << "{ const uint32_t t = " << v << "; a = t ? a / t : 0u; }";
      else if (x == 4) o << "{ const uint32_t t = " << v << "; a = t ? a % t : 0u; }";
      else if (x == 6) o << "a &= ~(" << v << ");";
      else if (x == 9) o << "a <<= ((" << v << ") & 31u);";
      else if (x == 10) o << "a >>= ((" << v << ") & 31u);";
      else if (x >= 11) o << "f = (a " << bin[x] << " " << v << ");";
      else o << "a " << bin[x] << "= " << v << ";";
    }
    o << "\n";
  }
  o << "  goto Lerr;\n";  // ran off the end (the reference would execute the zero guard: error)
  return true;
}

std::string lane_test(uint32_t mask) {
  if (mask && !(mask & (mask - 1))) {
    int l = 0;
    while (!((mask >> l) & 1)) ++l;
    return "lane == " + std::to_string(l);
  }
  char buf[64];
  snprintf(buf, sizeof buf, "(0x%xu >> lane) & 1u", mask);
  return buf;
}

}  // namespace

// Source of one specialised model.  `name` becomes the model struct name; the kernels are
// extern "C" <enc_kernel> / <dec_kernel>.
// The post-processing pass for ONE PCOMP program, compiled: PostProcessor.write (PostProcessor.cs:37-86) with ZPAQL.run
// (ZPAQL.cs:1028-1265) replaced by the program translated to straight C with gotos -- the device analogue of the reference's
// x86 JIT for PCOMP (ZPAQL.cs:353-1008).  The kernel takes the jobs the native kernels left (jobkind == PK_GENERIC) whose stored
// program is exactly `prog`, and marks them PK_COMPILED; everything else stays for the interpreter pass.  Returns "" when the
// program cannot be translated faithfully.
std::string generate_post_source(const Bytes& prog, int ph, int pm) {
  std::ostringstream body;
  if (prog.empty() || !translate_zpaql(prog.data(), (int)prog.size(), body, true)) return std::string();
  std::ostringstream o;
  o << "// generated by zpq_codegen.cpp: PCOMP program of " << prog.size() << " bytes, ph " << ph << ", pm " << pm << "\n"
    << "#include \"zpq_plan.h\"\n"
    << "namespace zpq {\n"
    << "struct PVm { uint32_t b, c, d, f; };\n"
    << "static __device__ __noinline__ int zpq_prog_run(uint32_t input, PVm& vm, uint32_t* const H, uint8_t* const M, uint32_t* const R,\n"
    << "                                                uint8_t* const outp, uint64_t& opos_, const uint64_t ocap, long long budget) {\n"
    << "  const uint32_t HMASK = " << ((1u << ph) - 1) << "u, MMASK = " << (uint32_t)((1ull << pm) - 1) << "u;\n"
    << "  uint32_t a = input, b = vm.b, c = vm.c, d = vm.d, f = vm.f;\n"
    << "  uint64_t opos = opos_;\n"
    << "  int rc = 0;\n"
    << "  (void)HMASK; (void)MMASK; (void)budget;\n"
    << body.str()
    << "  Lerr: rc = 1;\n"
    << "  Lhalt:\n"
    << "  vm.b = b; vm.c = c; vm.d = d; vm.f = f; opos_ = opos;\n"
    << "  return rc;\n"
    << "}\n"
    << "}  // namespace zpq\n"
    << R"ZPQK(
extern "C" __global__ void __launch_bounds__(128) zpq_post_rt(const zpq::PostParams Q, const uint8_t* prog, uint32_t plen, uint32_t* counter) {
  using namespace zpq;
  const int lane = threadIdx.x & 31;
  const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= Q.resident) return;
  const Plan* plan = Q.plan;
  uint8_t* arena = Q.arenas + (uint64_t)gw * Q.arena_stride;
  for (;;) {
    uint32_t job = 0;
    if (lane == 0) job = atomicAdd(counter, 1u);
    job = __shfl_sync(0xFFFFFFFFu, job, 0);
    if (job >= Q.njobs) break;
    if ((Q.jobkind[job] & 15u) != PK_GENERIC) continue;
    const DecJob J = Q.djobs[job];
    const PostJob O = Q.pjobs[job];
    const BlockResult R0 = Q.raw_results[job];
    if (R0.status != 0 || !J.seg_count) continue;                    // the interpreter pass reports these
    const uint8_t* raw = Q.raw + J.out_off;
    const uint64_t* send = Q.seg_end + J.seg_first;
    const uint64_t e0 = send[0];
    if (e0 < 3 || raw[0] != 1) continue;
    const uint32_t psize = raw[1] + 256u * raw[2];
    if (psize != plen || 3ull + psize > e0) continue;
    bool same = true;
    for (uint32_t i = lane; i < psize; i += 32) same = same && raw[3 + i] == prog[i];
    if (!__all_sync(0xFFFFFFFFu, same)) continue;
    // ZPAQL.initp: H, M, R zeroed
    uint32_t* H = reinterpret_cast<uint32_t*>(arena + plan->off_ph);
    uint8_t* M = arena + plan->off_pm;
    uint32_t* Rr = reinterpret_cast<uint32_t*>(arena + plan->off_pr);
    const uint64_t hn = 1ull << plan->ph, mn4 = ((1ull << plan->pm) + 3) / 4;
    for (uint64_t i = lane; i < hn; i += 32) H[i] = 0;
    for (uint64_t i = lane; i < mn4; i += 32) reinterpret_cast<uint32_t*>(M)[i] = 0;
    for (uint32_t i = lane; i < 256; i += 32) Rr[i] = 0;
    __syncwarp();
    uint32_t status = 0;
    uint64_t opos = 0;
    if (lane == 0) {
      PVm vm; vm.b = vm.c = vm.d = vm.f = 0;
      uint8_t* out = Q.out + O.out_off;
      uint64_t pos = 3ull + psize;
      for (uint32_t sg = 0; sg < J.seg_count && status == 0; ++sg) {
        const uint64_t end = send[sg];
        const long long budget = 65536 + 512 * (long long)(end + O.out_cap);
        for (; pos < end; ++pos)
          if (zpq_prog_run(raw[pos], vm, H, M, Rr, out, opos, O.out_cap, budget)) { status = 3; break; }       // ZPQ_BLOCK_ZPAQL
        if (status == 0 && zpq_prog_run(0xFFFFFFFFu, vm, H, M, Rr, out, opos, O.out_cap, budget)) status = 3;
        Q.seg_out_end[J.seg_first + sg] = opos;
      }
      if (status == 0 && opos > O.out_cap) status = 1;                                                          // ZPQ_BLOCK_OVERFLOW
      Q.results[job].out_len = opos; Q.results[job].status = status;
      Q.jobkind[job] = PK_COMPILED;
    }
    __syncwarp();
  }
}
)ZPQK";
  return o.str();
}

std::string generate_model_source(const Header& hdr, const std::string& name, const std::string& enc_kernel,
                                  const std::string& dec_kernel, bool* compiled_hcomp, int* duo_g, bool* fdec) {
  std::unique_ptr<Plan> plp(new Plan);
  Plan& pl = *plp;
  build_plan(hdr, false, 48 * 1024, pl);
  if (!pl.lane_ok) throw Failure(ZPQ_E_UNSUPPORTED, "model has more than 32 components: no lane-resident specialisation");
  std::ostringstream o;
  o << "// Generated by zpq_codegen from block header";
  for (size_t i = 0; i < hdr.wire.size() && i < 48; ++i) { char b[8]; snprintf(b, sizeof b, " %02x", hdr.wire[i]); o << b; }
  o << (hdr.wire.size() > 48 ? " ...\n" : "\n");
  o << "#include \"zpq_duo.cuh\"\n#include \"zpq_fdec.cuh\"\nnamespace zpq {\nstruct " << name << " {\n";

  uint32_t tmask[10] = {0};
  for (int i = 0; i < pl.n; ++i) tmask[pl.comp[i].type] |= 1u << i;
  const bool mix_regs = pl.nmix <= kMixRegs;
  std::vector<bool> mix_ct(pl.nmix);
  for (int k = 0; k < pl.nmix; ++k) {
    mix_ct[k] = mix_regs && pl.mix[k].cmask == 255 && pl.mix[k].mask >= 255;
    if (mix_ct[k])
      o << "  typedef MixCT<" << k << ", " << (int)pl.mix[k].lane << ", " << (int)pl.mix[k].j0 << ", " << (int)pl.mix[k].m << ", "
        << (int)pl.mix[k].rate << ", " << pl.mix[k].mask << "u> Mix" << k << ";\n";
  }

  // ---- phase A ----
  o << "  static __device__ __forceinline__ void phase_a(const Shared& S, const WarpCtx& W, LaneRegs& r, int lane) {\n";
  static const struct { int t; const char* fn; } pa[] = {{C_CM, "pa_cm"}, {C_ICM, "pa_icm"}, {C_ISSE, "pa_isse"},
                                                         {C_MATCH, "pa_match"}, {C_MIX2, "pa_mix2"}, {C_SSE, "pa_sse"}};
  const uint32_t umask = tmask[C_ICM] | tmask[C_ISSE] | tmask[C_MATCH];
  if (umask) o << "    pa_unified(S, W, r);   // ICM + ISSE + MATCH, branch-free on all lanes\n";
  for (auto& e : pa)
    if (tmask[e.t] && e.t != C_ICM && e.t != C_ISSE && e.t != C_MATCH) o << "    if (" << lane_test(tmask[e.t]) << ") " << e.fn << "(S, W, r);\n";
  if (tmask[C_ISSE]) o << "    __syncwarp();   // ISSE chain slots are read by every lane\n";
  o << "  }\n";

  // ---- levels ----
  o << "  static __device__ __forceinline__ void levels(const Shared& S, const WarpCtx& W, LaneRegs& r, int lane) {\n";
  // An ISSE chain (every level holds one ISSE fed by the component of the level before, the usual
  // shape of ICM-ISSE-ISSE... models) is evaluated redundantly by all lanes from the {w0, w1*64}
  // slots the owners published in phase A: no SHFL and no branch per level.
  auto sole_isse = [&](int L) {   // lane of the only non-MIX component at level L if it is an ISSE, else -1
    int found = -1, count = 0;
    for (int i = 0; i < pl.n; ++i)
      if (pl.comp[i].level == L && pl.comp[i].type != C_MIX && pl.comp[i].type != C_CONS) { found = i; ++count; }
    return (count == 1 && pl.comp[found].type == C_ISSE) ? found : -1;
  };
  auto mixes_at = [&](int L) {
    for (int k = 0; k < pl.nmix; ++k)
      if (pl.mix[k].level == L) {
        if (mix_ct[k]) o << "    Mix" << k << "::predict(W, r, lane);\n";
        else o << "    mix_predict_rt(S.mix[" << k << "], W, r, lane);\n";
      }
  };
  bool any_chain = false; (void)any_chain;
  for (int L = 1; L <= pl.maxlevel;) {
    // try to start a chain at level L
    std::vector<int> chain;
    int lane0 = sole_isse(L);
    if (lane0 >= 0) {
      chain.push_back(lane0);
      int LL = L + 1;
      while (LL <= pl.maxlevel) {
        bool mix_between = false;
        for (int k = 0; k < pl.nmix; ++k) if (pl.mix[k].level == LL - 1) mix_between = true;
        int nx = sole_isse(LL);
        if (mix_between || nx < 0 || pl.comp[nx].a[1] != chain.back()) break;
        chain.push_back(nx);
        ++LL;
      }
    }
    if (chain.size() >= 2) {
      any_chain = true;
      const int root = pl.comp[chain[0]].a[1];
      o << "    {  // ISSE chain, levels " << L << ".." << (L + (int)chain.size() - 1) << "\n";
      o << "      const int2* cs = r.chain - lane;\n";
      for (size_t q = 0; q < chain.size(); ++q) o << "      const int2 w" << q << " = cs[" << chain[q] << "];\n";
      o << "      int v = __shfl_sync(ZPQ_FULL, r.p, " << root << ");\n";
      for (size_t q = 0; q < chain.size(); ++q)
        o << "      v = clamp2k((w" << q << ".x * v + w" << q << ".y) >> 16); r.p = lane == " << chain[q] << " ? v : r.p;\n";
      o << "    }\n";
      L += (int)chain.size();
      mixes_at(L - 1);
      continue;
    }
    uint32_t lm[10] = {0};
    for (int i = 0; i < pl.n; ++i)
      if (pl.comp[i].level == L && pl.comp[i].type != C_MIX) lm[pl.comp[i].type] |= 1u << i;
    const bool any = lm[C_ISSE] || lm[C_AVG] || lm[C_MIX2] || lm[C_SSE];
    if (any) {
      o << "    {  // level " << L << "\n      const int pj = __shfl_sync(ZPQ_FULL, r.p, r.srcj);\n";
      if (lm[C_AVG] || lm[C_MIX2]) o << "      const int pk = __shfl_sync(ZPQ_FULL, r.p, r.srck);\n";
      if (lm[C_ISSE]) o << "      if (" << lane_test(lm[C_ISSE]) << ") r.p = ev_isse(r, pj);\n";
      if (lm[C_AVG]) o << "      if (" << lane_test(lm[C_AVG]) << ") r.p = ev_avg(r, pj, pk);\n";
      if (lm[C_MIX2]) o << "      if (" << lane_test(lm[C_MIX2]) << ") r.p = ev_mix2(r, pj, pk);\n";
      if (lm[C_SSE]) o << "      if (" << lane_test(lm[C_SSE]) << ") r.p = ev_sse(S, r, pj);\n";
      o << "    }\n";
    }
    mixes_at(L);
    ++L;
  }
  o << "  }\n";

  // ---- update ----
  o << "  static __device__ __forceinline__ void update(const Shared& S, const WarpCtx& W, LaneRegs& r, int lane, int y) {\n";
  if (tmask[C_ISSE] || tmask[C_MIX2]) o << "    const int pj = __shfl_sync(ZPQ_FULL, r.p, r.srcj);\n";
  if (tmask[C_MIX2]) o << "    const int pk = __shfl_sync(ZPQ_FULL, r.p, r.srck);\n";
  if (tmask[C_CM]) o << "    if (" << lane_test(tmask[C_CM]) << ") up_cm(S, r, y);\n";
  if (tmask[C_ICM] | tmask[C_ISSE] | tmask[C_MATCH]) o << "    up_unified(S, W, r, y, " << ((tmask[C_ISSE] || tmask[C_MIX2]) ? "pj" : "0") << ");\n";
  if (tmask[C_MIX2]) o << "    if (" << lane_test(tmask[C_MIX2]) << ") up_mix2(S, r, y, pj, pk);\n";
  if (tmask[C_SSE]) o << "    if (" << lane_test(tmask[C_SSE]) << ") up_sse(S, r, y);\n";
  for (int k = 0; k < pl.nmix; ++k) {
    if (mix_ct[k]) o << "    Mix" << k << "::update(S, W, r, lane, y);\n";
    else o << "    mix_update_rt(S, S.mix[" << k << "], W, r, lane, y);\n";
  }
  o << "  }\n";
  o << "  static __device__ __forceinline__ void mix_shift(LaneRegs& r, int y) {\n";
  for (int k = 0; k < pl.nmix; ++k) if (mix_ct[k]) o << "    Mix" << k << "::shift(r, y);\n";
  o << "  }\n";
  o << "  static __device__ __forceinline__ void mix_new_byte(const Shared& S, WarpCtx& W, LaneRegs& r, int lane) {\n";
  for (int k = 0; k < pl.nmix; ++k)
    if (mix_ct[k])
      o << "    W.mixh[" << k << "] = __shfl_sync(ZPQ_FULL, r.h, " << (int)pl.mix[k].lane << "); Mix" << k << "::load_current(W, r, lane);\n";
  o << "  }\n";

  // ---- HCOMP ----
  std::ostringstream body;
  const bool ok = translate_zpaql(pl.hcomp, pl.hcomp_len, body);
  if (compiled_hcomp) *compiled_hcomp = ok;
  // hcomp_m: `smask` names the lanes that run this HCOMP together (the whole warp, or one block's lanes in the two-role encoder)
  o << "  static __device__ __forceinline__ int hcomp_m(const Shared& S, WarpCtx& W, VM& vm, VMEnv& env, uint32_t input, int lane, uint32_t smask) {\n";
  if (ok) {
    o << "    // ZPAQL -> C, run uniformly by all lanes (every lane stores the same values)\n"
      << "    const uint32_t HMASK = " << ((1u << pl.hh) - 1) << "u, MMASK = " << (uint32_t)((1ull << pl.hm) - 1) << "u;\n"
      << "    uint32_t* const H = env.H; uint8_t* const M = env.M; uint32_t* const R = env.R;\n"
      << "    uint32_t a = input, b = vm.b, c = vm.c, d = vm.d, f = vm.f;\n"
      << "    int budget = 1 << 20, rc = 0;\n"
      << "    (void)HMASK; (void)MMASK; (void)H; (void)M; (void)R; (void)budget;\n"
      << "    __syncwarp(smask);\n"
      << body.str()
      << "  Lerr: rc = 1;\n"
      << "  Lhalt:\n"
      << "    vm.b = b; vm.c = c; vm.d = d; vm.f = f;\n"
      << "    __syncwarp(smask);\n"
      << "    return rc;\n";
  } else {
    o << "    return GenericModel::hcomp(S, W, vm, env, input, lane);\n";
  }
  o << "  }\n";
  o << "  static __device__ __forceinline__ int hcomp(const Shared& S, WarpCtx& W, VM& vm, VMEnv& env, uint32_t input, int lane) {\n"
    << "    return hcomp_m(S, W, vm, env, input, lane, ZPQ_FULL);\n";
  o << "  }\n};\n";

  // ---- time-skewed encoder policy (zpq_pipe.cuh) ----
  if (pl.pipe_ok) {
    o << "struct Pipe_" << name << " {\n"
      << "  static constexpr int D = " << pl.coder_delay << ", N = " << pl.n << ", RS = " << pl.ring_slots << ", RSTRIDE = " << pl.ring_stride << ";\n";
    for (int k = 0; k < pl.nmix; ++k)
      if (mix_regs)
        o << "  typedef MixPipe<" << k << ", " << (int)pl.mix[k].lane << ", " << (int)pl.mix[k].j0 << ", " << (int)pl.mix[k].m << ", "
          << (int)pl.mix[k].rate << ", " << pl.mix[k].mask << "u, " << (int)pl.mix[k].cmask << "u, " << (int)pl.comp[pl.mix[k].lane].delay
          << "> PMix" << k << ";\n";
    o << "  static __device__ __forceinline__ void lead0(const Shared& S, const PipeCtx& W, LaneRegs& r, int lane, int k, int y) {\n";
    if (tmask[C_CM]) o << "    if (" << lane_test(tmask[C_CM]) << ") pipe_cm(S, W, r, lane, y);\n";
    if (tmask[C_MATCH]) o << "    if (" << lane_test(tmask[C_MATCH]) << ") pipe_match(S, W, r, lane, k, y);\n";
    o << "  }\n";
    o << "  template <bool CHECKED> static __device__ __forceinline__ void lag(const Shared& S, const PipeCtx& W, LaneRegs& r, int lane) {\n";
    if (tmask[C_ICM] | tmask[C_ISSE]) o << "    pipe_icm_isse<CHECKED>(S, W, r, lane);   // branch-free on all lanes\n";
    if (tmask[C_AVG]) o << "    if (" << lane_test(tmask[C_AVG]) << ") pipe_avg<CHECKED>(W, r, lane);\n";
    if (tmask[C_MIX2]) o << "    if (" << lane_test(tmask[C_MIX2]) << ") pipe_mix2<CHECKED>(S, W, r, lane);\n";
    if (tmask[C_SSE]) o << "    if (" << lane_test(tmask[C_SSE]) << ") pipe_sse<CHECKED>(S, W, r, lane);\n";
    o << "  }\n";
    o << "  template <bool CHECKED> static __device__ __forceinline__ void mixes(const Shared& S, const PipeCtx& W, LaneRegs& r, int lane) {\n";
    for (int k = 0; k < pl.nmix; ++k) {
      if (mix_regs) o << "    PMix" << k << "::template tick<CHECKED>(S, W, r, lane);\n";
      else o << "    pipe_mix_rt<CHECKED>(S, S.mix[" << k << "], " << (int)pl.comp[pl.mix[k].lane].delay << ", W, r, lane);\n";
    }
    o << "  }\n";
    o << "  static __device__ __forceinline__ void prefetch(const PipeCtx& W, const LaneRegs& r, uint32_t hnext, uint32_t cnext, int lane) {\n";
    for (int k = 0; k < pl.nmix; ++k)
      if (mix_regs) o << "    PMix" << k << "::prefetch(W, r, hnext, cnext, lane);\n";
    o << "  }\n";
    o << "  static __device__ __forceinline__ int hcomp(const Shared& S, WarpCtx& W, VM& vm, VMEnv& env, uint32_t input, int lane) {\n"
      << "    return " << name << "::hcomp(S, W, vm, env, input, lane);\n  }\n};\n";
  }
  // ---- two-role encoder policy (zpq_duo.cuh) ----
  const int G = pl.n <= 8 ? 8 : pl.n <= 16 ? 16 : 32;
  std::unique_ptr<Plan> pdp(new Plan);
  Plan& pd = *pdp;
  build_plan(hdr, false, 48 * 1024, pd, G);
  const bool duo = ok && pd.duo_ok;
  if (duo_g) *duo_g = duo ? G : 0;
  if (duo) {
    uint32_t lmask = tmask[C_CONS] | tmask[C_CM] | tmask[C_MATCH];   // components the lead roles predict completely
    o << "struct Duo_" << name << " {\n"
      << "  static constexpr int G = " << G << ", N = " << pl.n << ", D = " << pd.coder_delay << ", LDEPTH = " << pd.duo_ldepth
      << ", HDEPTH = " << pd.duo_hdepth << ";\n"
      << "  static constexpr unsigned LMASK = 0x" << std::hex << lmask << std::dec << "u;\n"
      << "  static constexpr bool HAS_HASHED = " << ((tmask[C_ICM] | tmask[C_ISSE]) ? "true" : "false") << ", HAS_MATCH = " << (tmask[C_MATCH] ? "true" : "false")
      << ", HAS_CM = " << (tmask[C_CM] ? "true" : "false") << ", HAS_CONS = " << (tmask[C_CONS] ? "true" : "false")
      << ", HAS_ISSE = " << ((tmask[C_ISSE] | tmask[C_ICM]) ? "true" : "false") << ", NEEDK = " << ((tmask[C_AVG] | tmask[C_MIX2]) ? "true" : "false")
      << ", FINAL_MIX = " << (pl.comp[pl.n - 1].type == C_MIX ? "true" : "false") << ", SPLIT = " << (pd.duo_split ? "true" : "false") << ";\n";
    for (int k = 0; k < pl.nmix; ++k)
      if (mix_regs)
        o << "  typedef MixDuo<" << G << ", " << k << ", " << (int)pl.mix[k].lane << ", " << (int)pl.mix[k].j0 << ", " << (int)pl.mix[k].m << ", "
          << (int)pl.mix[k].rate << ", " << pl.mix[k].mask << "u, " << (int)pl.mix[k].cmask << "u, " << (int)pd.comp[pl.mix[k].lane].delay
          << ", LMASK, " << ((int)pl.mix[k].lane == pl.n - 1 ? "true" : "false") << "> QMix" << k << ";\n";
    o << "  static __device__ __forceinline__ void lanes(const Shared& S, const CoderCtx<G>& C, LaneRegs& r, int pj, int pk, int y, bool act, uint32_t t, int& p, int lane) {\n";
    if (tmask[C_AVG]) o << "    if (" << lane_test(tmask[C_AVG]) << ") duo_avg<G>(r, pj, pk, act, p);\n";
    if (tmask[C_MIX2]) o << "    if (" << lane_test(tmask[C_MIX2]) << ") duo_mix2<G>(S, C, r, pj, pk, y, act, t, p, lane);\n";
    if (tmask[C_SSE]) o << "    if (" << lane_test(tmask[C_SSE]) << ") duo_sse<G>(S, C, r, pj, y, act, t, p, lane);\n";
    o << "  }\n";
    o << "  template <bool FAST, int KT, bool RING> static __device__ __forceinline__ void mixes(const Shared& S, const CoderCtx<G>& C, LaneRegs& r, const Hist& H, int& p, int& pmv, int lane, uint32_t gmask, int gbase, bool live, int krt) {\n";
    for (int k = 0; k < pl.nmix; ++k) {
      if (mix_regs) o << "    QMix" << k << "::template tick<FAST, KT, RING>(S, C, r, H, p, pmv, lane, gmask, gbase, live, krt);\n";
      else o << "    duo_mix_rt<G, LMASK, RING>(S, S.mix[" << k << "], " << (int)pd.comp[pl.mix[k].lane].delay << ", C, r, H, p, pmv, lane, gmask, gbase, live);\n";
    }
    o << "  }\n";
    o << "  static __device__ __forceinline__ void prefetch(const LaneRegs& r, uint32_t hnext, uint32_t cnext, int lane, uint32_t gmask, int gbase) {\n";
    for (int k = 0; k < pl.nmix; ++k)
      if (mix_regs) o << "    QMix" << k << "::prefetch(r, hnext, cnext, lane, gmask, gbase);\n";
    o << "  }\n";
    o << "  static __device__ __forceinline__ int hcomp(const Shared& S, WarpCtx& W, VM& vm, VMEnv& env, uint32_t input, int lane, uint32_t smask) {\n"
      << "    return " << name << "::hcomp_m(S, W, vm, env, input, lane, smask);\n  }\n};\n";
  }
  // ---- speculative decoder policy (zpq_fdec.cuh): every component evaluated redundantly in all lanes from the slots ----
  bool fd = pl.n >= 1 && pl.n <= 8 && pl.nmix <= kMixRegs;
  int ml = -1;
  for (int i = 0; i < pl.n && fd; ++i) {
    const int t = pl.comp[i].type;
    if (t == C_MATCH) { if (ml >= 0) fd = false; ml = i; }
    else if (t != C_CONS && t != C_CM && t != C_ICM && t != C_ISSE && t != C_MIX) fd = false;
  }
  for (int k = 0; k < pl.nmix && fd; ++k)
    if (pl.mix[k].cmask != 255 || pl.mix[k].mask < 255 || pl.mix[k].m > 8) fd = false;
  if (fdec) *fdec = fd;
  if (fd) {
    o << "struct Fdec_" << name << " {\n"
      << "  static constexpr int N = " << pl.n << ", ML = " << ml << ", NMIX = " << pl.nmix << ";\n"
      << "  static constexpr bool H_SMEM = " << ((4ull << pl.hh) <= 2048 ? "true" : "false") << ", M_SMEM = " << ((1ull << pl.hm) <= 1024 ? "true" : "false") << ";\n"
      << "  static constexpr unsigned M_ICM = 0x" << std::hex << tmask[C_ICM] << "u, M_ISSE = 0x" << tmask[C_ISSE] << "u, M_CM = 0x" << tmask[C_CM]
      << "u, M_CONS = 0x" << tmask[C_CONS] << ", M_SLOT = 0x" << (tmask[C_ICM] | tmask[C_ISSE] | tmask[C_CM] | tmask[C_CONS]) << std::dec << "u;\n";
    for (int k = 0; k < pl.nmix; ++k)
      o << "  typedef FMix<" << k << ", " << (int)pl.mix[k].lane << ", " << (int)pl.mix[k].j0 << ", " << (int)pl.mix[k].m << ", "
        << (int)pl.mix[k].rate << ", " << pl.mix[k].mask << "u> FMix" << k << ";\n";
    o << "  static __device__ __forceinline__ int eval(const Shared& S, LaneRegs& r, const int2* cw, const int* lm, int lane, int vmatch, int& pj, int* pin, int* pm) {\n";
    int mk = 0;
    for (int i = 0; i < pl.n; ++i) {
      const CompDesc& d = pl.comp[i];
      switch (d.type) {
        case C_MATCH: o << "    const int v" << i << " = vmatch;\n"; break;
        case C_ISSE: o << "    const int v" << i << " = clamp2k((cw[" << i << "].x * v" << (int)d.a[1] << " + cw[" << i << "].y) >> 16);\n"; break;
        case C_MIX: {
          o << "    const int q" << i << " = 0";
          for (int j = 0; j < d.a[2]; ++j) o << " | (v" << (d.a[1] + j) << " & lm[" << j << "])";
          o << ";\n";
          o << "    pin[" << mk << "] = q" << i << "; pm[" << mk << "] = FMix" << mk << "::predict(r, lane, q" << i << ");\n"
            << "    const int v" << i << " = pm[" << mk << "];\n";
          ++mk;
          break;
        }
        default: o << "    const int v" << i << " = cw[" << i << "].x;\n"; break;   // CONS, CM, ICM: the owner's published prediction
      }
    }
    o << "    int own = 0; pj = 0;\n";
    for (int i = 0; i < pl.n; ++i)
      if (pl.comp[i].type == C_ISSE)
        o << "    own |= v" << i << " & lm[" << i << "]; pj |= v" << (int)pl.comp[i].a[1] << " & lm[" << i << "];\n";
    o << "    r.p = own;\n    return v" << (pl.n - 1) << ";\n  }\n";
    o << "  static __device__ __forceinline__ void mix_new_byte(const Shared& S, WarpCtx& W, LaneRegs& r, int lane) {\n";
    for (int k = 0; k < pl.nmix; ++k)
      o << "    FMix" << k << "::new_byte(r, W, W.H[" << (int)pl.mix[k].lane << "u & W.hmask], lane);\n";
    o << "  }\n";
    o << "  template <int KB> static __device__ __forceinline__ void mix_advance(const Shared& S, LaneRegs& r, const WarpCtx& W, int lane, int y, int yprev, const int* pin, const int* pm, uint32_t c8new) {\n";
    for (int k = 0; k < pl.nmix; ++k)
      o << "    FMix" << k << "::template advance<KB>(S, r, W, lane, y, yprev, pin[" << k << "], pm[" << k << "], c8new);\n";
    o << "  }\n";
    o << "  static __device__ __forceinline__ int hcomp(const Shared& S, WarpCtx& W, VM& vm, VMEnv& env, uint32_t input, int lane) {\n"
      << "    return " << name << "::hcomp(S, W, vm, env, input, lane);\n  }\n};\n";
  }
  o << "}  // namespace zpq\n\n";
  if (fd)
    o << "extern \"C\" __global__ void __launch_bounds__(zpq::kFdecThreads, 1) " << dec_kernel << "_f(const zpq::CodecParams P) {\n"
      << "  extern __shared__ __align__(128) uint8_t smem[];\n  zpq::fdec_body<zpq::Fdec_" << name << ">(P, smem);\n}\n";
  if (duo)
    o << "extern \"C\" __global__ void __launch_bounds__(zpq::kCtaThreads, 1) " << enc_kernel << "_d(const zpq::CodecParams P) {\n"
      << "  extern __shared__ __align__(128) uint8_t smem[];\n  zpq::encode_duo_body<zpq::Duo_" << name << ">(P, smem);\n}\n";
  // <enc_kernel>: time-skewed encoder when the model allows it; <enc_kernel>_l: lane-resident encoder
  // that walks bit by bit (blocks of 256 MB and more, A/B measurements)
  o << "extern \"C\" __global__ void __launch_bounds__(zpq::kCtaThreads, 1) " << enc_kernel << "(const zpq::CodecParams P) {\n"
    << "  extern __shared__ __align__(128) uint8_t smem[];\n";
  if (pl.pipe_ok) o << "  zpq::encode_pipe_body<zpq::Pipe_" << name << ">(P, smem);\n}\n";
  else o << "  zpq::encode_lanes_body<zpq::" << name << ">(P, smem);\n}\n";
  o << "extern \"C\" __global__ void __launch_bounds__(zpq::kCtaThreads, 1) " << enc_kernel << "_l(const zpq::CodecParams P) {\n"
    << "  extern __shared__ __align__(128) uint8_t smem[];\n  zpq::encode_lanes_body<zpq::" << name << ">(P, smem);\n}\n";
  o << "extern \"C\" __global__ void __launch_bounds__(zpq::kCtaThreads, 1) " << dec_kernel << "(const zpq::CodecParams P) {\n"
    << "  extern __shared__ __align__(128) uint8_t smem[];\n  zpq::decode_lanes_body<zpq::" << name << ">(P, smem);\n}\n";
  return o.str();
}

}  // namespace zpq

#ifdef ZPQ_GEN_MAIN
// zpq_gen: emits the ahead-of-time specialisations of the built-in models (build.py runs it).
int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: zpq_gen <outdir>\n"); return 2; }
  try {
    for (int level = 1; level <= 3; ++level) {
      zpq::Bytes wire;
      zpq::builtin_model(level, wire);
      zpq::Header h;
      zpq::parse_header(wire.data(), wire.size(), h);
      const std::string id = "aot" + std::to_string(level);
      bool compiled = false;
      int duo_g = 0;
      bool fd = false;
      std::string src = zpq::generate_model_source(h, "Model_" + id, "zpq_enc_" + id, "zpq_dec_" + id, &compiled, &duo_g, &fd);
      std::ostringstream reg;
      reg << "\n#include \"zpq_aot.h\"\nnamespace {\nconst unsigned char kHeader[] = {";
      for (size_t i = 0; i < wire.size(); ++i) reg << (int)wire[i] << (i + 1 < wire.size() ? "," : "");
      reg << "};\nconst zpq::AotRegistrar kReg(kHeader, sizeof kHeader, (const void*)zpq_enc_" << id << ", (const void*)zpq_enc_" << id
          << "_l, " << (duo_g ? "(const void*)zpq_enc_" + id + "_d" : std::string("nullptr")) << ", " << duo_g << ", (const void*)zpq_dec_" << id << ", " << (fd ? "(const void*)zpq_dec_" + id + "_f" : std::string("nullptr")) << ", \"" << id << (compiled ? " (HCOMP compiled)" : " (HCOMP interpreted)") << "\");\n}\n";
      const std::string path = std::string(argv[1]) + "/zpq_gen_" + id + ".cu";
      FILE* f = fopen(path.c_str(), "w");
      if (!f) { perror(path.c_str()); return 1; }
      fputs(src.c_str(), f);
      fputs(reg.str().c_str(), f);
      fclose(f);
    }
  } catch (const std::exception& e) { fprintf(stderr, "zpq_gen: %s\n", e.what()); return 1; }
  return 0;
}
#endif
