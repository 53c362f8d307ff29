// zpq_frontend.cpp -- host-side model front end of libzpaqb200: method strings -> ZPAQL config
// text -> block header / PCOMP bytecode.  Runs once per batch on the host and produces < 1 KB;
// the device only ever sees the resulting bytes.
//
// Mirrors (behaviour, not code) of the reference at /root/reference/ZPAQSharp:
//   expand_method   LibZPAQ.cs:128-283   (inside compressBlock)
//   make_config     LibZPAQ.cs:388-1044  (makeConfig)
//   compile_config  Compiler.cs:13-478   (Compiler ctor + CompileComp; helpers per SURVEY app. B)
//   builtin_model   Compressor.cs:45-83  (startBlock(int))
#include "zpq_host.h"

#include <cctype>
#include <cstdio>
#include <cstring>

namespace zpq {

static const int kCompLen[10] = {0, 2, 3, 2, 3, 4, 6, 6, 3, 5};  // Component.cs:27-43
int comp_len(int type) { return type >= 0 && type < 10 ? kCompLen[type] : 0; }

static int bit_length(uint32_t x) { int r = 0; for (; x; x >>= 1) ++r; return r; }   // lg(), LZBuffer.cs:118
static int popcount(uint32_t x) { int r = 0; for (; x; x >>= 1) r += x & 1; return r; }  // nbits(), :130

static std::string num(long v) { return std::to_string(v); }

// ------------------------------------------------------------------------------------------
// ZPAQL assembler
// ------------------------------------------------------------------------------------------
namespace {

enum Word {  // values >= 256 are the assembler's structure words (Compiler.cs:123-154)
  W_POST = 256, W_PCOMP, W_END, W_IF, W_IFNOT, W_ELSE, W_ENDIF, W_DO, W_WHILE, W_UNTIL, W_FOREVER,
  W_IFL, W_IFNOTL, W_ELSEL, W_SEMI
};
enum { OP_JT = 39, OP_JF = 47, OP_JMP = 63, OP_LJ = 255 };

// Mnemonic for opcode byte `op` built from the instruction fields (ISA: ZPAQL.cs:256-321);
// empty for undefined opcodes.
std::string mnemonic(int op) {
  static const char* R[8] = {"a", "b", "c", "d", "*b", "*c", "*d", ""};
  if (op == 0) return "error";
  if (op == 255) return "lj";
  if (op < 64) {
    int d = op >> 3, x = op & 7;
    if (d == 7) {
      static const char* sp[8] = {"halt", "out", "", "hash", "hashd", "", "", "jmp"};
      return sp[x];
    }
    switch (x) {
      case 0: return std::string(R[d]) + "<>a";
      case 1: return std::string(R[d]) + "++";
      case 2: return std::string(R[d]) + "--";
      case 3: return std::string(R[d]) + "!";
      case 4: return std::string(R[d]) + "=0";
      case 7: return d < 4 ? std::string(R[d]) + "=r" : d == 4 ? "jt" : d == 5 ? "jf" : "r=a";
      default: return "";
    }
  }
  if (op < 128) {
    int d = (op >> 3) & 7;
    if (d == 7) return "";
    return std::string(R[d]) + "=" + R[op & 7];
  }
  static const char* O[16] = {"+=", "-=", "*=", "/=", "%=", "&=", "&~", "|=", "^=", "<<=", ">>=", "==", "<", ">", 0, 0};
  const char* o = O[(op >> 3) & 15];
  if (!o) return "";
  return std::string("a") + o + R[op & 7];
}

const std::vector<std::string>& word_list() {
  static std::vector<std::string> w;
  if (w.empty()) {
    for (int i = 0; i < 256; ++i) w.push_back(mnemonic(i));
    const char* extra[] = {"post", "pcomp", "end", "if", "ifnot", "else", "endif", "do", "while", "until",
                           "forever", "ifl", "ifnotl", "elsel", ";"};
    for (const char* e : extra) w.push_back(e);
  }
  return w;
}

struct Scanner {
  const std::string& s;  // config text
  size_t pos;            // start of the current token
  int nest;              // 0 between tokens, -1 inside a token, >0 comment depth
  int line;
  const int* args;
  explicit Scanner(const std::string& text, const int* a) : s(text), pos(0), nest(0), line(1), args(a) {}
  char at(size_t i) const { return i < s.size() ? s[i] : '\0'; }

  // Move to the first character of the next token (skipping the rest of the current one,
  // white space and nested parenthesised comments).
  void advance() {
    for (; at(pos); ++pos) {
      char ch = at(pos);
      if (ch == '\n') ++line;
      if (ch == '(') nest += 1 + (nest < 0);
      else if (nest > 0 && ch == ')') --nest;
      else if (nest < 0 && (unsigned char)ch <= ' ') nest = 0;
      else if (nest == 0 && (unsigned char)ch > ' ') { nest = -1; return; }
    }
    throw Failure(ZPQ_E_CONFIG, "unexpected end of config");
  }
  bool is(const std::string& w) const {
    if (w.empty()) return false;
    size_t a = pos, k = 0;
    for (; (unsigned char)at(a) > ' ' && at(a) != '(' && k < w.size(); ++a, ++k)
      if (tolower((unsigned char)at(a)) != tolower((unsigned char)w[k])) return false;
    return k == w.size() && ((unsigned char)at(a) <= ' ' || at(a) == '(');
  }
  [[noreturn]] void bad(const std::string& what) const {
    std::string tok;
    for (size_t a = pos; (unsigned char)at(a) > ' ' && tok.size() < 40; ++a) tok += at(a);
    throw Failure(ZPQ_E_CONFIG, "Config line " + num(line) + " at " + tok + ": " + what);
  }
  int number(int lo, int hi) {
    advance();
    long r = 0;
    const char* p = s.c_str() + pos;
    if (p[0] == '$' && p[1] >= '1' && p[1] <= '9') {
      if (p[2] == '+') r = atol(p + 3);
      if (args) r += args[p[1] - '1'];
    } else if (p[0] == '-' || isdigit((unsigned char)p[0])) r = atol(p);
    else bad("expected a number");
    if (r < lo) bad("number too low");
    if (r > hi) bad("number too high");
    return (int)r;
  }
  void keyword(const char* w) {
    advance();
    if (!is(w)) bad(std::string("expected ") + w);
  }
  int lookup(const std::vector<std::string>& list) {
    advance();
    for (size_t i = 0; i < list.size(); ++i)
      if (is(list[i])) return (int)i;
    bad("unexpected");
  }
};

// Assemble one HCOMP/PCOMP body; returns the word that ended it (post / pcomp / end).
int assemble(Scanner& sc, Bytes& code) {
  std::vector<int> ifs, dos;
  auto pop = [&](std::vector<int>& st) {
    if (st.empty()) throw Failure(ZPQ_E_CONFIG, "unmatched IF or DO");
    int v = st.back(); st.pop_back(); return v;
  };
  auto push = [&](std::vector<int>& st, int v) {
    if (st.size() >= 1000) throw Failure(ZPQ_E_CONFIG, "IF or DO nested too deep");
    st.push_back(v);
  };
  const std::vector<std::string>& words = word_list();
  for (;;) {
    int op = sc.lookup(words);
    if (op == W_POST || op == W_PCOMP || op == W_END) { code.push_back(0); return op; }
    int arg1 = -1, arg2 = -1;
    const int here = (int)code.size();
    switch (op) {
      case W_IF: op = OP_JF; arg1 = 0; push(ifs, here + 1); break;
      case W_IFNOT: op = OP_JT; arg1 = 0; push(ifs, here + 1); break;
      case W_IFL: case W_IFNOTL:
        code.push_back(op == W_IFL ? OP_JT : OP_JF);
        code.push_back(3);
        op = OP_LJ; arg1 = arg2 = 0;
        push(ifs, (int)code.size() + 1);
        break;
      case W_ELSE: case W_ELSEL: {
        const bool longjump = (op == W_ELSEL);
        op = longjump ? OP_LJ : OP_JMP;
        arg1 = 0;
        if (longjump) arg2 = 0;
        int a = pop(ifs);
        if (code[a - 1] != OP_LJ) {
          int j = here - a + 1 + (longjump ? 1 : 0);
          if (j > 127) sc.bad("IF too big, try IFL, IFNOTL");
          code[a] = (uint8_t)j;
        } else {
          int j = here + 2 + (longjump ? 1 : 0);
          code[a] = j & 255; code[a + 1] = (j >> 8) & 255;
        }
        push(ifs, here + 1);
        break;
      }
      case W_ENDIF: {
        int a = pop(ifs);
        if (code[a - 1] != OP_LJ) {
          int j = here - a - 1;
          if (j > 127) sc.bad("IF too big, try IFL, IFNOTL, ELSEL");
          code[a] = (uint8_t)j;
        } else { code[a] = here & 255; code[a + 1] = (here >> 8) & 255; }
        break;
      }
      case W_DO: push(dos, here); break;
      case W_WHILE: case W_UNTIL: case W_FOREVER: {
        int a = pop(dos);
        int j = a - here - 2;
        if (j >= -127) {
          op = op == W_WHILE ? OP_JT : op == W_UNTIL ? OP_JF : OP_JMP;
          arg1 = j & 255;
        } else {
          if (op == W_WHILE) { code.push_back(OP_JF); code.push_back(3); }
          if (op == W_UNTIL) { code.push_back(OP_JT); code.push_back(3); }
          op = OP_LJ; arg1 = a & 255; arg2 = a >> 8;
        }
        break;
      }
      default:
        if ((op & 7) == 7) {
          if (op == OP_LJ) { int v = sc.number(0, 65535); arg1 = v & 255; arg2 = v >> 8; }
          else if (op == OP_JT || op == OP_JF || op == OP_JMP) arg1 = sc.number(-128, 127) & 255;
          else arg1 = sc.number(0, 255);
        }
    }
    if (op >= 0 && op <= 255) code.push_back((uint8_t)op);
    if (arg1 >= 0) code.push_back((uint8_t)arg1);
    if (arg2 >= 0) code.push_back((uint8_t)arg2);
    if (code.size() > 65535 - 8) sc.bad("program too big");
  }
}

}  // namespace

void compile_config(const std::string& text, const int* args, Bytes& hdr, Bytes& pcomp, std::string* pcomp_cmd) {
  static const std::vector<std::string> kinds = {"", "const", "cm", "icm", "match", "avg", "mix2", "mix", "isse", "sse"};
  Scanner sc(text, args);
  hdr.assign(7, 0);
  pcomp.clear();
  sc.keyword("comp");
  for (int k = 2; k < 7; ++k) hdr[k] = (uint8_t)sc.number(0, 255);
  const int n = hdr[6];
  for (int i = 0; i < n; ++i) {
    sc.number(i, i);
    int t = sc.lookup(kinds);
    int len = comp_len(t);
    if (len < 1) sc.bad("invalid component");
    hdr.push_back((uint8_t)t);
    for (int j = 1; j < len; ++j) hdr.push_back((uint8_t)sc.number(0, 255));
  }
  hdr.push_back(0);
  sc.keyword("hcomp");
  Bytes prog;
  int end = assemble(sc, prog);
  hdr.insert(hdr.end(), prog.begin(), prog.end());
  size_t hsize = hdr.size() - 2;
  if (hsize > 65535) throw Failure(ZPQ_E_CONFIG, "program too big");
  hdr[0] = hsize & 255; hdr[1] = (uint8_t)(hsize >> 8);
  if (end == W_POST) {
    sc.number(0, 0);
    sc.keyword("end");
  } else if (end == W_PCOMP) {
    sc.advance();
    size_t j = sc.pos;
    while (sc.at(j) && sc.at(j) != ';') ++j;
    if (pcomp_cmd) *pcomp_cmd = text.substr(sc.pos, j - sc.pos);
    sc.pos = j + (sc.at(j) ? 1 : 0);
    if (assemble(sc, pcomp) != W_END) sc.bad("expected END");
  } else if (end != W_END) sc.bad("expected END or POST 0 END or PCOMP cmd ; ... END");
}

// ------------------------------------------------------------------------------------------
// Built-in models (Compressor.cs:45-83 stores them as bytecode; kept here as config source,
// tests compare the assembled bytes with the reference's).
// ------------------------------------------------------------------------------------------
static const char* kMin =
    "comp 1 2 0 0 2 0 icm 16 1 isse 19 0 "
    "hcomp *b=a a=0 d=0 hash b-- hash *d=a d++ b-- hash b-- hash *d=a halt end ";
static const char* kMid =
    "comp 3 3 0 0 8 0 icm 5 1 isse 13 0 2 isse 17 1 3 isse 18 2 4 isse 18 3 5 isse 19 4 6 match 22 24 "
    "7 mix 16 0 7 24 255 "
    "hcomp c++ *c=a b=c a=0 d= 1 hash *d=a b-- d++ hash *d=a b-- d++ hash *d=a b-- d++ hash *d=a "
    "b-- d++ hash *d=a b-- d++ hash b-- hash *d=a d++ a=*c a<<= 8 *d=a halt end ";
static const char* kMax =
    "comp 5 9 0 0 22 0 const 160 1 icm 5 2 isse 13 1 3 isse 16 2 4 isse 18 3 5 isse 19 4 6 isse 19 5 "
    "7 isse 20 6 8 match 22 24 9 icm 17 10 isse 19 9 11 icm 13 12 icm 13 13 icm 13 14 icm 14 "
    "15 mix 16 0 15 24 255 16 mix 8 0 16 10 255 17 mix2 0 15 16 24 0 18 sse 8 17 32 255 "
    "19 mix2 8 17 18 16 255 20 sse 16 19 32 255 21 mix2 0 19 20 16 0 "
    "hcomp c++ *c=a b=c a=0 d= 2 hash *d=a b-- d++ hash *d=a b-- d++ hash *d=a b-- d++ hash *d=a b-- "
    "d++ hash *d=a b-- d++ hash b-- hash *d=a b-- d++ hash *d=a b-- d++ a=*c a&~ 32 "
    "a> 64 if a< 91 if d++ hashd d-- *d<>a a+=*d a*= 20 *d=a jmp 9 endif endif "
    "a=*d a== 0 ifnot d++ *d=a d-- endif *d=0 "
    "d++ d++ b=c b-- a=0 hash *d=a d++ b-- a=0 hash *d=a d++ b-- a=0 hash *d=a "
    "d++ a=b a-= 212 b=a a=0 hash *d=a b<>a a-= 216 b<>a a=*b a&= 60 hashd "
    "d++ a=*c a<<= 9 *d=a d++ d++ d++ d++ d++ *d=a halt end ";

void builtin_model(int level, Bytes& hdr) {
  const char* src = level == 1 ? kMin : level == 2 ? kMid : level == 3 ? kMax : nullptr;
  if (level < 1) throw Failure(ZPQ_E_ARG, "compression level must be at least 1");
  if (!src) throw Failure(ZPQ_E_ARG, "compression level too high");
  Bytes pc;
  compile_config(src, nullptr, hdr, pc, nullptr);
}

// ------------------------------------------------------------------------------------------
// Level digit -> method string
// ------------------------------------------------------------------------------------------
int block_arg0(uint64_t n) {  // LibZPAQ.cs:125
  int v = bit_length((uint32_t)(n + 4095)) - 20;
  return v > 0 ? v : 0;
}

// Byte-gap histogram of a block (LibZPAQ.cs:242-258): gap[k] = positions whose byte last occurred k positions earlier, the
// "last occurrence" of a byte not seen yet being position 0.  Host form; k_gap_hist (zpq_preproc.cu) is the device form.
void gap_histogram(const uint8_t* data, uint64_t n, int* gap) {
  for (int k = 0; k < kGapBins; ++k) gap[k] = 0;
  int last[256] = {0};
  for (uint64_t i = 0; i < n; ++i) {
    const int k = (int)i - last[data[i]];
    if (k > 0 && k < kGapBins) ++gap[k];
    last[data[i]] = (int)i;
  }
}

bool method_needs_analysis(const std::string& method) { return !method.empty() && method[0] >= '5' && method[0] <= '9'; }

std::string expand_method(const std::string& method, const uint8_t* data, uint64_t n) {
  if (!method_needs_analysis(method)) return expand_method_gaps(method, n, nullptr);
  std::vector<int> gap(kGapBins, 0);
  gap_histogram(data, n, gap.data());
  return expand_method_gaps(method, n, gap.data());
}

std::string expand_method_gaps(const std::string& method, uint64_t n, const int* gaps) {
  if (method.empty()) throw Failure(ZPQ_E_ARG, "empty method");
  if (!isdigit((unsigned char)method[0])) return method;
  const int arg0 = block_arg0(n);
  int commas = 0, arg[4] = {0, 0, 0, 0};
  for (size_t i = 1; i < method.size() && commas < 4; ++i) {
    if (method[i] == ',' || method[i] == '.') ++commas;
    else if (isdigit((unsigned char)method[i])) arg[commas] = arg[commas] * 10 + method[i] - '0';
  }
  const unsigned type = commas == 0 ? 512u : (unsigned)(arg[1] * 4 + arg[2]);
  const int level = method[0] - '0';
  const int doe8 = (type & 2) * 2;
  const std::string hashsz = "," + num(19 + arg0 + (arg0 <= 6));
  const std::string sufsz = "," + num(21 + arg0);
  std::string m = "x" + num(arg0);
  const std::string lz1 = "," + num(1 + doe8) + ",";

  if (level == 0) return "0" + num(arg0) + ",0";
  if (level == 1) {
    if (type < 40) return m + ",0";
    m += lz1;
    if (type < 80) m += "4,0,1,15";
    else if (type < 128) m += "4,0,2,16";
    else if (type < 256) m += "4,0,2" + hashsz;
    else if (type < 960) m += "5,0,3" + hashsz;
    else m += "6,0,3" + hashsz;
    return m;
  }
  if (level == 2) {
    if (type < 32) return m + ",0";
    m += lz1;
    m += type < 64 ? "4,0,3" + hashsz : "4,0,7" + sufsz + ",1";
    return m;
  }
  if (level == 3) {
    if (type < 20) return m + ",0";
    if (type < 48) return m + lz1 + "4,0,3" + hashsz;
    if (type >= 640 || (type & 1)) return m + "," + num(3 + doe8) + "ci1";
    return m + "," + num(2 + doe8) + ",12,0,7" + sufsz + ",1c0,0,511i2";
  }
  if (level == 4) {
    if (type < 12) return m + ",0";
    if (type < 24) return m + lz1 + "4,0,3" + hashsz;
    if (type < 48) return m + "," + num(2 + doe8) + ",5,0,7" + sufsz + "1c0,0,511";
    if (type < 900) {
      m += "," + num(doe8) + "ci1,1,1,1,2a";
      if (type & 1) m += "w";
      return m + "m";
    }
    return m + "," + num(3 + doe8) + "ci1";
  }
  // levels 5..9: many models, periodic ones chosen from the byte-gap histogram
  m += "," + num(doe8);
  m += (type & 1) ? "w2c0,1010,255i1" : "w1i1";
  m += "c256ci1,1,1,1,1,1,2a";
  const int NR = kGapBins;
  if (!gaps) throw Failure(ZPQ_E_ARG, "method levels 5..9 need the block's byte-gap histogram");
  std::vector<int> gap(gaps, gaps + NR);
  int rest = (int)n - gap[1] - gap[2] - gap[3];
  for (int pass = 0; pass < 2; ++pass) {
    int period = 0, seen = 0;
    double best = 0;
    for (int j = 5; j < NR && seen < rest; ++j) {
      const double sc = gap[j] / (256.0 + rest - seen);
      if (sc > best) { best = sc; period = j; }
      seen += gap[j];
    }
    if (!(period > 4 && best > 0.1)) break;
    m += "c0,0," + num(999 + period) + ",255i1";
    if (period <= 255) m += "c0," + num(period) + "i1";
    rest -= gap[period];
    gap[period] = 0;
  }
  return m + "c0,2,0,255i1c0,3,0,0,255i1c0,4,0,0,0,255i1mm16ts19t0";
}

// ------------------------------------------------------------------------------------------
// makeConfig: post-processor programs (ZPAQL source, data of the archive format) and the
// context-model generator.
// ------------------------------------------------------------------------------------------
namespace {

// inverse E8E9 over M[0..d) run at end of input by the LZ77/BWT post-processors
const char* kUnE8 =
    " a=b a==d ifnot a+= 4 a<d if a=*b a&= 254 a== 232 if c=b b++ b++ b++ b++ a=*b a++ a&= 254 a== 0 if"
    " b-- a=*b b-- a<<= 8 a+=*b b-- a<<= 8 a+=*b a-=b a++ *b=a a>>= 8 b++ *b=a a>>= 8 b++ *b=a b++"
    " endif b=c endif endif a=*b out b++ forever endif\n";

std::string post_lz_bits(const int* args, bool e8) {  // "lazy2", LibZPAQ.cs:427-571
  const int rb = args[0] > 4 ? args[0] - 4 : 0;
  const char* emit = e8 ? "" : " out";
  std::string p = "pcomp lazy2 3 ;\n a> 255 if\n";
  if (e8) p += std::string(" b=0 d=r 4 do") + kUnE8;
  p += " a=0 b=0 c=0 d=0 r=a 1 r=a 2 r=a 3 r=a 4 halt endif\n"
       " a<<=d a+=c c=a a= 8 a+=d d=a\n"
       " a=r 1 a== 0 if a= 1 r=a 2 a=c a&= 3 a> 0 if"
       " a-- a<<= 3 r=a 3 a=c a>>= 2 c=a b=r 3 a&= 7 a+=b r=a 3 a=c a>>= 3 c=a a=d a-= 5 d=a a= 1 r=a 1"
       " else a=c a>>= 2 c=a d-- d-- a= 3 r=a 1 endif endif\n"
       " do a=r 1 a== 1 if a=d a> 2 if a=c a&= 1 a== 1 if"
       " a=c a>>= 1 c=a b=r 2 a=c a&= 1 a+=b a+=b r=a 2 a=c a>>= 1 c=a d-- d--"
       " else a=c a>>= 1 c=a a=r 2 a<<= 2 b=a a=c a&= 3 a+=b r=a 2 a=c a>>= 2 c=a d-- d-- d--";
  p += rb ? " a= 5 r=a 1" : " a= 2 r=a 1";
  p += " endif forever endif endif\n";
  if (rb)
    p += " a=r 1 a== 5 if a=d a> " + num(rb - 1) + " if a=c a&= " + num((1 << rb) - 1) + " r=a 5 a=c a>>= " +
         num(rb) + " c=a a=d a-= " + num(rb) + " d=a a= 2 r=a 1 endif endif\n";
  p += " a=r 1 a== 2 if a=r 3 a>d ifnot a=c r=a 6 a=d r=a 7 b=r 3 a= 1 a<<=b d=a a-- a&=c a+=d";
  if (rb) p += " a<<= " + num(rb) + " d=r 5 a+=d a-= " + num((1 << rb) - 1);
  p += " d=a b=r 4 a=b a-=d c=a d=r 2 do a=d a> 0 if d-- a=*c *b=a c++ b++";
  p += emit;
  p += " forever endif a=b r=a 4 a=r 6 b=r 3 a>>=b c=a a=r 7 a-=b d=a a=0 r=a 1 endif endif\n"
       " do a=r 1 a== 3 if a=d a> 1 if a=c a&= 1 a== 1 if"
       " a=c a>>= 1 c=a b=r 2 a&= 1 a+=b a+=b r=a 2 a=c a>>= 1 c=a d-- d--"
       " else a=c a>>= 1 c=a d-- a= 4 r=a 1 endif forever endif endif\n"
       " a=r 1 a== 4 if a=d a> 7 if b=r 4 a=c *b=a";
  p += emit;
  p += " b++ a=b r=a 4 a=c a>>= 8 c=a a=d a-= 8 d=a a=r 2 a-- r=a 2 a== 0 if a=0 r=a 1 endif endif endif\n"
       " halt end\n";
  return p;
}

std::string post_lz_bytes(bool e8) {  // "lzpre", LibZPAQ.cs:574-638
  const char* emit = e8 ? "" : " out";
  std::string p = "pcomp lzpre c ;\n a> 255 if\n";
  if (e8) p += std::string(" d=b b=0 do") + kUnE8;
  p += " b=0 c=0 d=0 a=0 r=a 1 r=a 2 halt endif\n"
       " c=a a=d a== 0 if a=c a>>= 6 a++ d=a a== 1 if a+=c r=a 1 a=0 r=a 2"
       " else d++ a=c a&= 63 a+= $3 r=a 1 a=0 r=a 2 endif\n"
       " else a== 1 if a=c *b=a b++";
  p += emit;
  p += " a=r 1 a-- a== 0 if d=0 endif r=a 1\n"
       " else a> 2 if a=r 2 a<<= 8 a|=c r=a 2 d--"
       " else a=r 2 a<<= 8 a|=c c=a a=b a-=c a-- c=a d=r 1 do a=*c *b=a c++ b++";
  p += emit;
  p += " d-- a=d a> 0 while endif endif endif\n halt end\n";
  return p;
}

std::string post_bwt(const int* args, bool e8) {  // "bwtrle", LibZPAQ.cs:641-795
  std::string p =
      "pcomp bwtrle c ;\n a> 255 ifnot *b=a b++ elsel\n"
      " b-- a=*b b-- a<<= 8 a+=*b b-- a<<= 8 a+=*b b-- a<<= 8 a+=*b c=a r=a 1 a=b r=a 2\n"
      " do a=b a> 0 if b-- a=*b a++ a&= 255 d=a d! *d++ forever endif\n"
      " d=0 d! *d= 1 a=0 do a+=*d *d=a d-- d<>a a! a> 255 a! d<>a until\n"
      " b=0 do a=c a>b if d=*b d! *d++ d=*d d-- *d=b b++ forever endif\n"
      " b=c b++ c=r 2 do a=c a>b if d=*b d! *d++ d=*d d-- *d=b b++ forever endif\n";
  if (args[0] <= 4) {
    p += " b=0 do a=c a>b if d=b a=*d a<<= 8 a+=*b *d=a b++ forever endif\n"
         " d=r 1 b=0 do a=d a== 0 ifnot a=*d a>>= 8 d=a";
    p += e8 ? " *b=*d b++" : " a=*d out";
    p += " forever endif\n";
    if (e8) p += std::string(" d=b b=0 do") + kUnE8;
    p += " endif halt end\n";
  } else if (e8) {
    p += " a=r 2 a-- r=a 2 c=0 d=r 1 do a=d a== 0 ifnot d=*d b=d a=*b a<<= 24 b=a"
         " a=r 4 r=a 5 a>>= 8 a|=b r=a 4 a=c a> 3 if a=r 5 a&= 254 a== 232 if"
         " a=r 4 a>>= 24 b=a a++ a&= 254 a< 2 if a=r 4 a-=c a+= 4 a<<= 8 a>>= 8 b<>a a<<= 24 a+=b r=a 4"
         " endif endif endif a=c a> 3 if a=r 5 out endif c++ forever endif\n"
         " b=r 4 a=c a> 3 a=b if out endif a>>= 8 b=a a=c a> 2 a=b if out endif a>>= 8 b=a"
         " a=c a> 1 a=b if out endif a>>= 8 b=a a=c a> 0 a=b if out endif\n"
         " endif halt end\n";
  } else {
    p += " d=r 1 do a=d a== 0 ifnot d=*d b=d a=*b out forever endif endif halt end\n";
  }
  return p;
}

const char* kPostE8 =  // "e8e9", LibZPAQ.cs:798-826
    "pcomp e8e9 d ;\n a> 255 if a=c a> 4 if c= 4 else a! a+= 5 a<<= 3 d=a a=b a>>=d b=a endif"
    " do a=c a> 0 if a=b out a>>= 8 b=a c-- forever endif\n"
    " else *b=b a<<= 24 d=a a=b a>>= 8 a+=d b=a c++ a=c a> 4 if a=*b out a&= 254 a== 232 if"
    " a=b a>>= 24 a++ a&= 254 a== 0 if a=b a>>= 24 a<<= 24 d=a a=b a-=c a+= 5 a<<= 8 a>>= 8 a|=d b=a"
    " endif endif endif endif halt end\n";

// Split "C[N1[,N2]...]" off the front of *p.
std::vector<int> take_command(const char*& p) {
  std::vector<int> v;
  v.push_back((unsigned char)*p++);
  if (isdigit((unsigned char)*p)) {
    v.push_back(*p++ - '0');
    while (isdigit((unsigned char)*p) || *p == ',' || *p == '.') {
      if (isdigit((unsigned char)*p)) v.back() = v.back() * 10 + *p - '0';
      else v.push_back(0);
      ++p;
    }
  }
  return v;
}

}  // namespace

// The PCOMP programs make_config can emit for a block whose header says (ph, pm), assembled: what the decoder's native
// post-processors (zpq_post.cu) recognise.  LZ77: ph = 0, pm = arg0 + 20 (LibZPAQ.cs:429,576); BWT: ph = pm = arg0 + 20 (:643);
// E8E9 alone: ph = pm = 0 (:799).  lzpre carries minMatch ("$3") in one operand byte: `wild` is its index.
void post_candidates(int ph, int pm, std::vector<PostCandidate>& out) {
  out.clear();
  auto assemble_post = [](const std::string& text, const int* args) {
    Bytes hdr, pc;
    compile_config("comp 0 0 0 0 0 hcomp halt " + text, args, hdr, pc, nullptr);
    return pc;
  };
  int args[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (ph == 0 && pm == 0) {
    PostCandidate c; c.kind = PK_E8E9; c.e8 = 1; c.param = 0; c.wild = -1; c.prog = assemble_post(kPostE8, args);
    out.push_back(c);
  }
  if (pm >= 20 && pm <= 31) {
    args[0] = pm - 20;
    for (int e8 = 0; e8 < 2; ++e8) {
      if (ph == 0) {
        PostCandidate c; c.kind = PK_LZ_BITS; c.e8 = e8; c.param = args[0] > 4 ? args[0] - 4 : 0; c.wild = -1;
        c.prog = assemble_post(post_lz_bits(args, e8 != 0), args);
        out.push_back(c);
        PostCandidate d; d.kind = PK_LZ_BYTES; d.e8 = e8; d.param = 0; d.wild = -1;
        args[2] = 1; d.prog = assemble_post(post_lz_bytes(e8 != 0), args);
        args[2] = 254; const Bytes other = assemble_post(post_lz_bytes(e8 != 0), args);
        args[2] = 0;
        int ndiff = 0;
        if (other.size() == d.prog.size())
          for (size_t i = 0; i < other.size(); ++i) if (other[i] != d.prog[i]) { d.wild = (int)i; ++ndiff; }
        if (ndiff == 1) out.push_back(d);
      }
      if (ph == pm) {
        PostCandidate c; c.kind = PK_BWT; c.e8 = e8; c.param = 0; c.wild = -1;
        c.prog = assemble_post(post_bwt(args, e8 != 0), args);
        out.push_back(c);
      }
    }
  }
}

std::string make_config(const std::string& method, int args[9]) {
  if (method.empty() || !strchr("xs0i", method[0])) throw Failure(ZPQ_E_CONFIG, "method must start with x, s, i or 0");
  const char kind = method[0];
  for (int i = 0; i < 9; ++i) args[i] = 0;
  const char* p = method.c_str() + 1;
  for (int i = 0; i < 9 && (isdigit((unsigned char)*p) || *p == ',' || *p == '.'); ++p) {
    if (isdigit((unsigned char)*p)) args[i] = args[i] * 10 + *p - '0';
    else if (++i < 9) args[i] = 0;
  }
  if (kind == '0') return "comp 0 0 0 0 0 hcomp end\n";

  const int pre = args[1] & 3;
  const bool e8 = args[1] >= 4 && args[1] <= 7;
  std::string head, post;
  switch (pre) {
    case 1: head = "comp 9 16 0 $1+20 "; post = post_lz_bits(args, e8); break;
    case 2: head = "comp 9 16 0 $1+20 "; post = post_lz_bytes(e8); break;
    case 3: head = "comp 9 16 $1+20 $1+20 "; post = post_bwt(args, e8); break;
    default: head = "comp 9 16 0 0 "; post = e8 ? kPostE8 : "end\n";
  }

  // Context model: H[0..254] contexts, H[255..511] last position of byte value, M = last 64 KB
  // written backward, C -> newest byte, R1/R2 = byte-LZ77 parse state (LibZPAQ.cs:835-847).
  int ncomp = 0, ctxbits = 5;
  const int membits = args[0] + 20;
  std::string comps, ctx = "hcomp\nc-- *c=a a+= 255 d=a *d=c\n";
  if (pre == 2)
    ctx += " a=r 1 a== 0 if a= " + num(111 + 57 * (e8 ? 1 : 0)) +
           " else a== 1 if a=*c r=a 2 a> 63 if a>>= 6 a++ a++ else a++ a++ endif else a-- endif endif r=a 1\n";

  while (*p && ncomp < 254) {
    std::vector<int> v = take_command(p);
    const int cmd = v[0];
    auto want = [&](size_t k, int dflt) { if (v.size() <= k) v.push_back(dflt); };

    if (cmd == 'c') {  // context model: ICM (N1%1000==0) or CM with limit N1%1000-1
      while (v.size() < 3) v.push_back(0);
      ctxbits = 11 + (v[2] < 256 ? bit_length(v[2]) : 6);
      for (size_t i = 3; i < v.size(); ++i)
        if (v[i] < 512) ctxbits += popcount(v[i]) * 3 / 4;
      if (ctxbits > membits) ctxbits = membits;
      comps += num(ncomp) + " ";
      if (v[1] % 1000 == 0) comps += "icm " + num(ctxbits - 6 - v[1] / 1000) + "\n";
      else comps += "cm " + num(ctxbits - 2 - v[1] / 1000) + " " + num(v[1] % 1000 - 1) + "\n";
      ctx += "d= " + num(ncomp) + " *d=0\n";
      if (v[2] > 1 && v[2] <= 255) {
        if (bit_length(v[2]) != bit_length(v[2] - 1)) ctx += "a=c a&= " + num(v[2] - 1) + " hashd\n";
        else ctx += "a=c a%= " + num(v[2]) + " hashd\n";
      } else if (v[2] >= 1000 && v[2] <= 1255)
        ctx += "a= 255 a+= " + num(v[2] - 1000) + " d=a a=*d a-=c a> 255 if a= 255 endif d= " + num(ncomp) + " hashd\n";
      for (size_t i = 3; i < v.size(); ++i) {
        if (i == 3) ctx += "b=c ";
        if (v[i] == 255) ctx += "a=*b hashd\n";
        else if (v[i] > 0 && v[i] < 255) ctx += "a=*b a&= " + num(v[i]) + " hashd\n";
        else if (v[i] >= 256 && v[i] < 512) {
          ctx += "a=r 1 a> 1 if a=r 2 a< 64 if a=*b ";
          if (v[i] < 511) ctx += "a&= " + num(v[i] - 256);
          ctx += " hashd else a>>= 6 hashd a=r 1 hashd endif else a= 255 hashd a=r 2 hashd endif\n";
        } else if (v[i] >= 1256)
          ctx += "a= " + num(((v[i] - 1000) >> 8) & 255) + " a<<= 8 a+= " + num((v[i] - 1000) & 255) + " a+=b b=a\n";
        else if (v[i] > 1000) ctx += "a= " + num(v[i] - 1000) + " a+=b b=a\n";
        if (v[i] < 512 && i < v.size() - 1) ctx += "b++ ";
      }
      ++ncomp;
    }

    if ((cmd == 'm' || cmd == 't' || cmd == 's') && ncomp > (cmd == 't')) {  // MIX / MIX2 / SSE
      want(1, 8);
      want(2, 24 + 8 * (cmd == 's'));
      if (cmd == 's') want(3, 255);
      ctxbits = 5 + v[1] * 3 / 4;
      comps += num(ncomp);
      if (cmd == 'm') comps += " mix " + num(v[1]) + " 0 " + num(ncomp) + " " + num(v[2]) + " 255\n";
      else if (cmd == 't') comps += " mix2 " + num(v[1]) + " " + num(ncomp - 1) + " " + num(ncomp - 2) + " " + num(v[2]) + " 255\n";
      else comps += " sse " + num(v[1]) + " " + num(ncomp - 1) + " " + num(v[2]) + " " + num(v[3]) + "\n";
      if (v[1] > 8) {
        ctx += "d= " + num(ncomp) + " *d=0 b=c a=0\n";
        int bitsleft = v[1];
        for (; bitsleft >= 16; bitsleft -= 8) {
          ctx += "a<<= 8 a+=*b";
          if (bitsleft > 16) ctx += " b++";
          ctx += "\n";
        }
        if (bitsleft > 8) ctx += "a<<= 8 a+=*b a>>= " + num(16 - bitsleft) + "\n";
        ctx += "a<<= 8 *d=a\n";
      }
      ++ncomp;
    }

    if (cmd == 'i' && ncomp > 0) {  // ISSE chain, each order raised by N%10 bytes
      ctx += "d= " + num(ncomp - 1) + " b=c a=*d d++\n";
      for (size_t i = 1; i < v.size() && ncomp < 254; ++i) {
        for (int j = 0; j < v[i] % 10; ++j) {
          ctx += "hash ";
          if (i < v.size() - 1 || j < v[i] % 10 - 1) ctx += "b++ ";
          ctxbits += 6;
        }
        ctx += "*d=a";
        if (i < v.size() - 1) ctx += " d++";
        ctx += "\n";
        if (ctxbits > membits) ctxbits = membits;
        comps += num(ncomp) + " isse " + num(ctxbits - 6 - v[i] / 10) + " " + num(ncomp - 1) + "\n";
        ++ncomp;
      }
    }

    if (cmd == 'a') {  // MATCH
      want(1, 24);
      while (v.size() < 4) v.push_back(0);
      comps += num(ncomp) + " match " + num(membits - v[3] - 2) + " " + num(membits - v[2]) + "\n";
      ctx += "d= " + num(ncomp) + " a=*d a*= " + num(v[1]) + " a+=*c a++ *d=a\n";
      ctxbits = 5 + (membits - v[2]) * 3 / 4;
      ++ncomp;
    }

    if (cmd == 'w') {  // word-model ICM-ISSE chain
      want(1, 1); want(2, 65); want(3, 26); want(4, 223); want(5, 20); want(6, 0);
      comps += num(ncomp) + " icm " + num(membits - 6 - v[6]) + "\n";
      for (int i = 1; i < v[1]; ++i)
        comps += num(ncomp + i) + " isse " + num(membits - 6 - v[6]) + " " + num(ncomp + i - 1) + "\n";
      ctx += "a=*c a&= " + num(v[4]) + " a-= " + num(v[2]) + " a&= 255 a< " + num(v[3]) + " if\n";
      for (int i = 0; i < v[1]; ++i) {
        ctx += i == 0 ? " d= " + num(ncomp) : std::string(" d++");
        ctx += " a=*d a*= " + num(v[5]) + " a+=*c a++ *d=a\n";
      }
      ctx += "else\n";
      for (int i = v[1] - 1; i > 0; --i) ctx += " d= " + num(ncomp + i - 1) + " a=*d d++ *d=a\n";
      ctx += " d= " + num(ncomp) + " *d=0\nendif\n";
      ncomp += v[1] - 1;
      ctxbits = membits - v[6];
      ++ncomp;
    }
  }
  return head + num(ncomp) + "\n" + comps + ctx + "halt\n" + post;
}

}  // namespace zpq
