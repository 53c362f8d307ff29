"""Builds the fragments of the reference that are compilable here:

  * the body of /root/reference/ZPAQSharp/divsufsort.cs (libdivsufsort-lite, still C text inside a C# class)
    -> oracle/_ref/libdivsufsort_ref.so                                   (build())
  * the body of /root/reference/ZPAQSharp/LZBuffer.cs (the LZ77 / BWT pre-processor, still C++ text) together with
    `e8e9` from LibZPAQ.cs:371-384, around a harness of the three library classes the reference does not contain
    (Array<T> as documented in LICENSE:618-634, Reader, StringBuffer) -> oracle/_ref/liblzbuffer_ref.so  (build_lzbuffer())
  * the bodies of Encoder.encode (Encoder.cs:86-103) and Decoder.decode (Decoder.cs:136-158) -- the 32-bit arithmetic
    coder, valid C++ once `uint` / `ulong` are typedefs -- as members of two harness structs that supply low / high / curr
    and the byte sink / source -> oracle/_ref/libcoder_ref.so                                            (build_coder())
  * the body of ZPAQL.execute (ZPAQL.cs:1028-1251), the 256-way interpreter of the HCOMP / PCOMP virtual machine, as a
    member of a harness struct that supplies the registers, H / M / R, the C#-ified helpers (swap, div, mod, err,
    outc) and the header layout of ZPAQL.cs:112-156 -> oracle/_ref/libzpaql_ref.so                       (build_zpaql())
  * the bodies of Predictor.init, predict0, update0 and find (Predictor.cs:39-172, 245-350, 353-475, 550-567) -- the
    component formulas of the bit predictor -- on top of the interpreter above.  Textual fixes only: `Array.Resize(ref x,`
    -> `x.resize(`, `.Length` -> `.size()`, `@` escapes, C# access modifiers, find()'s C#-ified signature.  The five
    one-line helpers whose C# text is wrong (train, squash, stretch, clamp2k, clamp512k) are supplied by the
    harness as the reference's own JIT comments state them (Predictor.cs:916-921, 1031-1036, 1116-1121); the static
    tables come from the literals in the reference (tests/golden/reference_kat.json) -> libpredictor_ref.so (build_predictor())

Test infrastructure only.  Nothing is copied into the repository: the C text is read where it lies, three
mechanical repairs of formatter damage are applied in memory (blank lines inside macro continuations, `budget.`
for `budget->` in trbudget_init / trbudget_check, C# `@` escapes), and the result is compiled into
oracle/_ref/libdivsufsort_ref.so (git-ignored; it travels to the GPU box with the snapshot).  The tests use it to
check the oracle's and the device's suffix arrays / BWT against the reference's own suffix sorter.
The rest of the reference is not valid C# and no .NET toolchain exists here: reference unbuildable."""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/ZPAQSharp/divsufsort.cs"
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "libdivsufsort_ref.so")


def repaired_text() -> str:
    lines = open(SRC, encoding="utf-8-sig").read().split("\n")
    # the C text sits between the class opening (line 10) and the two closing braces of class / namespace
    end = len(lines)
    closes = 0
    while end > 0 and closes < 2:
        end -= 1
        if lines[end].strip() == "}":
            closes += 1
    body = lines[10:end]
    out = []
    in_macro = False
    for ln in body:
        if in_macro:
            # multi-line macros here are all `do { ... } while(0)`: the formatter dropped some continuation
            # backslashes and inserted blank lines; restore until the closing `while(0)`
            if ln.strip() == "":
                continue
            if re.search(r"while\s*\(0\)", ln):
                in_macro = False
                out.append(ln.rstrip().rstrip("\\"))
            else:
                out.append(ln.rstrip().rstrip("\\") + " \\")
            continue
        out.append(ln)
        if re.match(r"\s*#\s*define\b", ln) and ln.rstrip().endswith("\\"):
            in_macro = True
    text = "\n".join(out)
    text = text.replace("@", "")          # C# identifier escapes (budget.@incval)
    # pointer parameters written with '.', only inside the definitions of the two trbudget helpers
    def fix(m):
        return m.group(0).replace("budget.", "budget->")
    text = re.sub(r"trbudget_init\(trbudget_t\s*\*\s*budget.*?\n}\n", fix, text, flags=re.S)
    text = re.sub(r"trbudget_check\(trbudget_t\s*\*\s*budget.*?\n}\n", fix, text, flags=re.S)
    return "#include <assert.h>\n#include <stdlib.h>\n#include <stdio.h>\n" + text + "\n"


def build(force: bool = False) -> str | None:
    if not os.path.exists(SRC):
        return OUT if os.path.exists(OUT) else None
    if os.path.exists(OUT) and not force and os.path.getmtime(OUT) >= os.path.getmtime(__file__):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    p = subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-w", "-x", "c", "-", "-o", OUT], input=repaired_text().encode(),
                       capture_output=True)
    if p.returncode != 0:
        sys.stderr.write(p.stderr.decode()[:4000])
        return None
    return OUT


LZ_SRC = "/root/reference/ZPAQSharp/LZBuffer.cs"
LIB_SRC = "/root/reference/ZPAQSharp/LibZPAQ.cs"
LZ_OUT = os.path.join(OUT_DIR, "liblzbuffer_ref.so")

# What the reference's LZBuffer text needs from classes the reference does not contain.  Array<T>: "resize(n, e) = n << e
# elements, zeroed; a[i]" (LICENSE:618-634); Reader: get()/read() (Reader.cs:7-28); StringBuffer: size()/data()
# (StringBuffer.cs:17-179); error() must not return (LICENSE:41-46).
_HARNESS_HEAD = r"""
#include <assert.h>
#include <string.h>
#include <stdlib.h>
#include <stdexcept>
#include <vector>
#define MAX(a, b) ((a) > (b) ? (a) : (b))
static void error(const char* msg) { throw std::runtime_error(msg); }
extern "C" int divsufsort(const unsigned char* T, int* SA, int n);
template <class T> class Array {
  std::vector<T> v;
 public:
  explicit Array(size_t n = 0) : v(n, T()) {}
  size_t size() const { return v.size(); }
  T& operator[](size_t i) { return v[i]; }
  const T& operator[](size_t i) const { return v[i]; }
};
class Reader {
 public:
  virtual int get() = 0;
  virtual int read(char* buf, int n) { int i = 0, c; while (i < n && (c = get()) >= 0) buf[i++] = c; return i; }
  virtual ~Reader() {}
};
class StringBuffer {
  std::vector<unsigned char> v;
 public:
  StringBuffer(const unsigned char* p, size_t n) : v(p, p + n) { v.push_back(0); v.pop_back(); }
  size_t size() const { return v.size(); }
  unsigned char* data() { return v.empty() ? (unsigned char*)"" : &v[0]; }
};
"""

_HARNESS_TAIL = r"""
extern "C" long long ref_lzbuffer(const unsigned char* in, unsigned n, const int* args, unsigned char* out, unsigned long long cap) {
  try {
    StringBuffer sb(in, n);
    int a[9];
    for (int i = 0; i < 9; ++i) a[i] = args[i];
    LZBuffer lz(sb, a);
    unsigned long long k = 0;
    int c;
    while ((c = lz.get()) >= 0) { if (k < cap) out[k] = (unsigned char)c; ++k; }
    return (long long)k;
  } catch (const std::exception&) { return -1; }
}
extern "C" void ref_e8e9(unsigned char* buf, int n) { e8e9(buf, n); }
"""


def lzbuffer_text() -> str:
    lines = open(LZ_SRC, encoding="utf-8-sig").read().split("\n")
    first = next(i for i, l in enumerate(lines) if l.strip().startswith("class LZBuffer"))
    end = len(lines)
    closes = 0
    while end > 0 and closes < 2:          # closing braces of class and namespace
        end -= 1
        if lines[end].strip() == "}":
            closes += 1
    body = lines[first + 2:end]            # after "class LZBuffer : Reader" and its "{"
    # the class text holds, in this order: member declarations, the free functions lg() / nbits() (libzpaq has them at
    # namespace scope), and the out-of-line member definitions (`LZBuffer::...`).  Split it back into those three.
    i_lg = next(i for i, l in enumerate(body) if "floor(log2(x)) + 1" in l)
    i_def = next(i for i, l in enumerate(body) if "Read n bytes of compressed output into p" in l)
    decl, free, defs = body[:i_lg], body[i_lg:i_def], body[i_def:]
    fix = lambda t: t.replace("inbuf.Length", "inbuf.size()").replace("ht.Length", "ht.size()").replace("libzpaq::Array", "Array")
    lib = open(LIB_SRC, encoding="utf-8-sig").read().split("\n")
    j = next(i for i, l in enumerate(lib) if l.strip().startswith("void e8e9(unsigned char* buf, int n)"))
    k = j
    depth = 0
    while True:                            # the function's own braces
        depth += lib[k].count("{") - lib[k].count("}")
        k += 1
        if depth == 0 and "{" in "".join(lib[j:k]):
            break
    e8 = "\n".join(lib[j:k])
    return (_HARNESS_HEAD + e8 + "\n" + fix("\n".join(free)) + "\nclass LZBuffer : public Reader {\n" + fix("\n".join(decl)) + "\n};\n"
            + fix("\n".join(defs)) + "\n" + _HARNESS_TAIL)


def build_lzbuffer(force: bool = False) -> str | None:
    if not (os.path.exists(LZ_SRC) and os.path.exists(SRC) and os.path.exists(LIB_SRC)):
        return LZ_OUT if os.path.exists(LZ_OUT) else None
    if os.path.exists(LZ_OUT) and not force and os.path.getmtime(LZ_OUT) >= os.path.getmtime(__file__):
        return LZ_OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    ds_o = os.path.join(OUT_DIR, "divsufsort_ref.o")
    p = subprocess.run(["gcc", "-O2", "-fPIC", "-w", "-c", "-x", "c", "-", "-o", ds_o], input=repaired_text().encode(), capture_output=True)
    if p.returncode != 0:
        sys.stderr.write(p.stderr.decode()[:4000])
        return None
    lz_o = os.path.join(OUT_DIR, "lzbuffer_ref.o")
    p = subprocess.run(["g++", "-O2", "-fPIC", "-w", "-fpermissive", "-std=c++14", "-c", "-x", "c++", "-", "-o", lz_o],
                       input=lzbuffer_text().encode(), capture_output=True)
    if p.returncode != 0:
        sys.stderr.write(p.stderr.decode()[:6000])
        return None
    p = subprocess.run(["g++", "-shared", "-o", LZ_OUT, lz_o, ds_o], capture_output=True)
    if p.returncode != 0:
        sys.stderr.write(p.stderr.decode()[:4000])
        return None
    return LZ_OUT


ENC_SRC = "/root/reference/ZPAQSharp/Encoder.cs"
DEC_SRC = "/root/reference/ZPAQSharp/Decoder.cs"
CODER_OUT = os.path.join(OUT_DIR, "libcoder_ref.so")


def _method_text(path: str, signature: str) -> str:
    """The text of one method, from its signature line to its closing brace, access modifier dropped."""
    lines = open(path, encoding="utf-8-sig").read().split("\n")
    j = next(i for i, l in enumerate(lines) if signature in l)
    k, depth, seen = j, 0, False
    while True:
        depth += lines[k].count("{") - lines[k].count("}")
        seen = seen or "{" in lines[k]
        k += 1
        if seen and depth == 0:
            break
    text = "\n".join(lines[j:k])
    return re.sub(r"^\s*(private|public)\s+(unsafe\s+)?", "", text)


def coder_text() -> str:
    enc = _method_text(ENC_SRC, "void encode(int y, int p)")
    dec = _method_text(DEC_SRC, "int decode(int p)")
    return r"""
#include <stdexcept>
#include <vector>
#define assert(x) ((void)0)
#include <sys/types.h>   // uint = unsigned int, ulong = unsigned long (64 bits here), as in C#
static_assert(sizeof(uint) == 4 && sizeof(ulong) == 8, "C# uint / ulong");
static void error(const char* msg) { throw std::runtime_error(msg); }
struct Sink { std::vector<unsigned char> v; void put(int c) { v.push_back((unsigned char)c); } };
struct RefEncoder {
  uint low, high; Sink out;
""" + enc + r"""
};
struct RefDecoder {
  uint low, high, curr; const unsigned char* p; unsigned long long n, pos;
  int get() { return pos < n ? p[pos++] : -1; }
""" + dec + r"""
};
extern "C" long long ref_arith_encode(const unsigned char* bits, const unsigned short* probs, unsigned n, unsigned char* out, unsigned long long cap) {
  RefEncoder e; e.low = 1; e.high = 0xFFFFFFFFu;
  for (unsigned i = 0; i < n; ++i) e.encode(bits[i] & 1, probs[i]);
  e.encode(1, 0);
  for (unsigned long long i = 0; i < e.out.v.size() && i < cap; ++i) out[i] = e.out.v[i];
  return (long long)e.out.v.size();
}
extern "C" int ref_arith_decode(const unsigned char* in, unsigned long long len, const unsigned short* probs, unsigned n, unsigned char* bits) {
  try {
    RefDecoder d; d.low = 1; d.high = 0xFFFFFFFFu; d.curr = 0; d.p = in; d.n = len; d.pos = 0;
    for (int i = 0; i < 4; ++i) d.curr = d.curr << 8 | (uint)(d.get() & 255);
    for (unsigned i = 0; i < n; ++i) bits[i] = (unsigned char)d.decode(probs[i]);
    return 0;
  } catch (const std::exception&) { return -1; }
}
"""


def build_coder(force: bool = False) -> str | None:
    if not (os.path.exists(ENC_SRC) and os.path.exists(DEC_SRC)):
        return CODER_OUT if os.path.exists(CODER_OUT) else None
    if os.path.exists(CODER_OUT) and not force and os.path.getmtime(CODER_OUT) >= os.path.getmtime(__file__):
        return CODER_OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    p = subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-w", "-std=c++14", "-x", "c++", "-", "-o", CODER_OUT],
                       input=coder_text().encode(), capture_output=True)
    if p.returncode != 0:
        sys.stderr.write(p.stderr.decode()[:6000])
        return None
    return CODER_OUT


ZPAQL_SRC = "/root/reference/ZPAQSharp/ZPAQL.cs"
ZPAQL_OUT = os.path.join(OUT_DIR, "libzpaql_ref.so")


def zpaql_text() -> str:
    exe = _method_text(ZPAQL_SRC, "int execute()")
    return r"""
#include <stdexcept>
#include <vector>
#include <string.h>
#include <sys/types.h>
#define assert(x) ((void)0)
typedef unsigned char byte;
struct RefVM {
  std::vector<byte> header;        // hsize(2) hh hm ph pm n COMP 0 [128-byte gap] HCOMP 0 ..., ZPAQL.cs:112-156
  int cend, hbegin, hend;
  uint a, b, c, d; int f, pc;
  std::vector<uint> hv; std::vector<byte> mv; uint r[256];
  std::vector<byte> out;
  uint& h(uint i) { return hv[i & (hv.size() - 1)]; }           // Array<T>::operator(): index masked to the size, LICENSE:618-634
  byte& m(uint i) { return mv[i & (mv.size() - 1)]; }
  void swap(uint& x) { a ^= x; x ^= a; a ^= x; }                  // ZPAQL.cs:1288-1300
  void swap(byte& x) { a ^= x; x = (byte)(x ^ a); a ^= x; }
  void div(uint x) { if (x != 0) a /= x; else a = 0; }            // ZPAQL.cs:1266-1286
  void mod(uint x) { if (x != 0) a %= x; else a = 0; }
  void err() { throw std::runtime_error("ZPAQL execution error"); }
  void outc(int ch) { out.push_back((byte)ch); }
""" + exe + r"""
};
// Runs `prog` (incl. its END byte) once per input value with persistent state, like ZPAQL.run0 (ZPAQL.cs:1253-1265).
// Returns the number of OUT bytes, -1 on a ZPAQL execution error, -2 when the instruction budget is spent.
extern "C" long long ref_zpaql_run(const unsigned char* prog, int proglen, int hbits, int mbits, const unsigned* inputs, int ninputs,
                                   unsigned* hout, int hn, unsigned char* outbuf, unsigned long long cap, unsigned long long budget) {
  RefVM z;
  z.cend = 8; z.hbegin = z.cend + 128; z.hend = z.hbegin + proglen;
  z.header.assign(z.hend + 304, 0);
  memcpy(&z.header[z.hbegin], prog, proglen);
  z.hv.assign((size_t)1 << hbits, 0); z.mv.assign((size_t)1 << mbits, 0);
  memset(z.r, 0, sizeof z.r);
  z.a = z.b = z.c = z.d = 0; z.f = 0; z.pc = 0;
  try {
    for (int i = 0; i < ninputs; ++i) {
      z.pc = z.hbegin; z.a = inputs[i];
      while (z.execute()) if (budget-- == 0) return -2;
    }
  } catch (const std::exception&) { return -1; }
  for (int i = 0; i < hn; ++i) hout[i] = z.h((uint)i);
  for (unsigned long long i = 0; i < z.out.size() && i < cap; ++i) outbuf[i] = z.out[i];
  return (long long)z.out.size();
}
"""


def build_zpaql(force: bool = False) -> str | None:
    if not os.path.exists(ZPAQL_SRC):
        return ZPAQL_OUT if os.path.exists(ZPAQL_OUT) else None
    if os.path.exists(ZPAQL_OUT) and not force and os.path.getmtime(ZPAQL_OUT) >= os.path.getmtime(__file__):
        return ZPAQL_OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    p = subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-w", "-std=c++14", "-x", "c++", "-", "-o", ZPAQL_OUT],
                       input=zpaql_text().encode(), capture_output=True)
    if p.returncode != 0:
        sys.stderr.write(p.stderr.decode()[:6000])
        return None
    return ZPAQL_OUT


PRED_SRC = "/root/reference/ZPAQSharp/Predictor.cs"
PRED_OUT = os.path.join(OUT_DIR, "libpredictor_ref.so")


def predictor_text() -> str:
    exe = _method_text(ZPAQL_SRC, "int execute()")
    enc = _method_text(ENC_SRC, "void encode(int y, int p)")
    fix = lambda t: re.sub(r"Array\.Resize\(ref\s+([\w\.]+),\s*", r"\1.resize(", t).replace(".Length", ".size()").replace("@", "")
    init = fix(_method_text(PRED_SRC, "void init() // build model"))
    pred = fix(_method_text(PRED_SRC, "int predict0() // default"))
    upd = fix(_method_text(PRED_SRC, "void update0(int y) // default"))
    find = fix(_method_text(PRED_SRC, "ulong find(byte[] ht, int sizebits, uint cxt)"))
    find = find.replace("ulong find(byte[] ht, int sizebits, uint cxt)", "size_t find(Arr<byte>& ht, int sizebits, uint cxt)")
    return r"""
#include <stdexcept>
#include <vector>
#include <string.h>
#include <sys/types.h>
#define NDEBUG 1
#define assert(x) ((void)0)
#define ssert(x) ((void)0)
#define allocx(a, b, c) ((void)0)
typedef unsigned char byte;
static void error(const char* msg) { throw std::runtime_error(msg); }
// Array<T> of libzpaq (LICENSE:618-634): resize(n, e) = n << e zeroed elements, a[i] plain, a(i) index masked to the size
template <class T> struct Arr {
  std::vector<T> v;
  void resize(size_t n, int e = 0) { v.assign(n << e, T()); }
  size_t size() const { return v.size(); }
  T& operator[](size_t i) { return v[i]; }
  T& operator()(size_t i) { return v[i & (v.size() - 1)]; }
};
struct RefVM {
  Arr<byte> header; int cend, hbegin, hend;
  uint a, b, c, d; int f, pc;
  Arr<uint> hv; Arr<byte> mv; uint r[256];
  uint& h(uint i) { return hv(i); }
  byte& m(uint i) { return mv(i); }
  void swap(uint& x) { a ^= x; x ^= a; a ^= x; }
  void swap(byte& x) { a ^= x; x = (byte)(x ^ a); a ^= x; }
  void div(uint x) { if (x != 0) a /= x; else a = 0; }
  void mod(uint x) { if (x != 0) a %= x; else a = 0; }
  void err() { throw std::runtime_error("ZPAQL execution error"); }
  void outc(int) {}
""" + exe + r"""
  void inith() {                                     // ZPAQL.inith / init, ZPAQL.cs:1010-1026
    hv.resize(1, header[2]); mv.resize(1, header[3]); memset(r, 0, sizeof r);
    a = b = c = d = 0; f = 0; pc = 0;
  }
  void run(uint input) { pc = hbegin; a = input; while (execute()) ; }   // run0, ZPAQL.cs:1253-1265
  uint H(int i) { return hv(i); }
};
enum { NONE, CONS, CM, ICM, MATCH, AVG, MIX2, MIX, ISSE, SSE };
static const int compsize[256] = {0, 2, 3, 2, 3, 4, 6, 6, 3, 5};          // Component.cs:27-43
struct CompState { size_t limit, cxt, a, b, c; Arr<uint> cm; Arr<byte> ht; Arr<unsigned short> a16; };
// a C# class variable is a reference: `Component cr = comp[i];` must alias comp[i]
struct Component {
  size_t &limit, &cxt, &a, &b, &c; Arr<uint>& cm; Arr<byte>& ht; Arr<unsigned short>& a16;
  Component(CompState& s) : limit(s.limit), cxt(s.cxt), a(s.a), b(s.b), c(s.c), cm(s.cm), ht(s.ht), a16(s.a16) {}
  void init() { limit = cxt = a = b = c = 0; cm.resize(0); ht.resize(0); a16.resize(0); }   // Component.cs:45-51
};
struct CompArray { CompState st[256]; Component operator[](int i) { return Component(st[i]); } };
static int sdt2k[256], sdt[1024]; static unsigned short ssquasht[1344]; static int stdt[712]; static byte sns[1024];
struct StateTable {                                                          // StateTable.cs:151-162
  int next(int state, int y) { return sns[state * 4 + y]; }
  int cminit(int state) { return ((sns[state * 4 + 3] * 2 + 1) << 22) / (sns[state * 4 + 2] + sns[state * 4 + 3] + 1); }
};
struct RefPredictor {
  int c8, hmap4; int p[256]; uint h[256]; RefVM z; CompArray comp; bool initTables; StateTable st;
  int dt2k[256]; int dt[1024]; unsigned short squasht[4096]; short stretcht[32768];
  byte* pcode; int pcode_size;
  bool isModeled() { return z.header[6] != 0; }
  // ---- restored helpers (the C# bodies are wrong, SURVEY 8c; stated as in the reference's JIT comments) ----
  void train(Component cr, int y) { uint& pn = cr.cm(cr.cxt); uint count = pn & 0x3ff; int err = y * 32767 - (pn >> 17);
                                    pn += (err * dt[count] & -1024) + (count < cr.limit); }          // Predictor.cs:1031-1036
  int squash(int x) { return squasht[x + 2048]; }
  int stretch(int x) { return stretcht[x]; }
  int clamp2k(int x) { return x < -2048 ? -2048 : x > 2047 ? 2047 : x; }                              // Predictor.cs:916-921
  int clamp512k(int x) { return x < -(1 << 19) ? -(1 << 19) : x >= (1 << 19) ? (1 << 19) - 1 : x; }    // Predictor.cs:1116-1121
  // ---- reference text ----
""" + find + "\n" + init + "\n" + pred + "\n" + upd + r"""
};
struct RefSink { unsigned char* p; unsigned long long cap, n; void put(int c) { if (n < cap) p[n] = (unsigned char)c; ++n; } };
struct RefBlockEncoder {          // Encoder (Encoder.cs:26-103) around the reference predictor: encode() is reference text
  uint low, high; RefSink out; RefPredictor* pr;
""" + enc + r"""
  void compress(int c) {          // Encoder.compress, Encoder.cs:39-57 (modeled branch)
    if (c == -1) encode(1, 0);
    else {
      encode(0, 0);
      for (int i = 7; i >= 0; --i) { int p = pr->predict0() * 2 + 1; int y = c >> i & 1; encode(y, p); pr->update0(y); }
    }
  }
};
extern "C" void ref_predictor_tables(const int* a_sdt2k, const int* a_sdt, const unsigned short* a_ssquasht, const int* a_stdt, const unsigned char* a_sns) {
  memcpy(sdt2k, a_sdt2k, sizeof sdt2k); memcpy(sdt, a_sdt, sizeof sdt); memcpy(ssquasht, a_ssquasht, sizeof ssquasht);
  memcpy(stdt, a_stdt, sizeof stdt); memcpy(sns, a_sns, sizeof sns);
}
// Feeds `input` through the model of block header `hdr` (as stored in an archive) and records the 16-bit probability handed
// to the coder for every bit: predict() * 2 + 1 (Encoder.cs:51).  Returns bits written or -1.
extern "C" long long ref_predict_trace(const unsigned char* hdr, unsigned long long hlen, const unsigned char* input, unsigned long long n,
                                       unsigned short* probs) {
  try {
    RefPredictor* P = new RefPredictor();
    P->initTables = false; P->pcode = 0; P->pcode_size = 0; P->c8 = 1; P->hmap4 = 1;
    int ncomp = hdr[6], pos = 7;
    for (int i = 0; i < ncomp; ++i) pos += compsize[hdr[pos]];
    int cend = pos + 1;                                      // one past the COMP END byte
    int hlen2 = (int)hlen - cend;                            // HCOMP incl. END
    RefVM& z = P->z;
    z.cend = cend; z.hbegin = cend + 128; z.hend = z.hbegin + hlen2 - 1;
    z.header.resize(z.hend + 304);
    memcpy(&z.header[0], hdr, cend);
    memcpy(&z.header[z.hbegin], hdr + cend, hlen2);
    P->init();
    unsigned long long k = 0;
    for (unsigned long long i = 0; i < n; ++i)
      for (int b = 7; b >= 0; --b) { probs[k++] = (unsigned short)(P->predict0() * 2 + 1); P->update0(input[i] >> b & 1); }
    delete P;
    return (long long)k;
  } catch (const std::exception&) { return -1; }
}
// The coded stream of one block as Compressor.compress produces it (Compressor.cs:156-221, 224-232): `preamble` (the
// PCOMP bytes, Compressor.cs:177-188) and the data go through Encoder.compress, then EOS.  Returns coded bytes or -1.
extern "C" long long ref_code_block(const unsigned char* hdr, unsigned long long hlen, const unsigned char* preamble, unsigned long long npre,
                                    const unsigned char* input, unsigned long long n, unsigned char* out, unsigned long long cap) {
  try {
    RefPredictor* P = new RefPredictor();
    P->initTables = false; P->pcode = 0; P->pcode_size = 0; P->c8 = 1; P->hmap4 = 1;
    int ncomp = hdr[6], pos = 7;
    for (int i = 0; i < ncomp; ++i) pos += compsize[hdr[pos]];
    int cend = pos + 1, hlen2 = (int)hlen - cend;
    RefVM& z = P->z;
    z.cend = cend; z.hbegin = cend + 128; z.hend = z.hbegin + hlen2 - 1;
    z.header.resize(z.hend + 304);
    memcpy(&z.header[0], hdr, cend);
    memcpy(&z.header[z.hbegin], hdr + cend, hlen2);
    P->init();
    RefBlockEncoder e; e.low = 1; e.high = 0xFFFFFFFFu; e.out.p = out; e.out.cap = cap; e.out.n = 0; e.pr = P;
    for (unsigned long long i = 0; i < npre; ++i) e.compress(preamble[i]);
    for (unsigned long long i = 0; i < n; ++i) e.compress(input[i]);
    e.compress(-1);
    delete P;
    return (long long)e.out.n;
  } catch (const std::exception&) { return -1; }
}
"""


def build_predictor(force: bool = False) -> str | None:
    if not (os.path.exists(PRED_SRC) and os.path.exists(ZPAQL_SRC)):
        return PRED_OUT if os.path.exists(PRED_OUT) else None
    if os.path.exists(PRED_OUT) and not force and os.path.getmtime(PRED_OUT) >= os.path.getmtime(__file__):
        return PRED_OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    p = subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-w", "-fpermissive", "-std=c++14", "-x", "c++", "-", "-o", PRED_OUT],
                       input=predictor_text().encode(), capture_output=True)
    if p.returncode != 0:
        sys.stderr.write(p.stderr.decode()[:8000])
        return None
    return PRED_OUT


def build_all(force: bool = False) -> dict:
    """Every fragment; returns {name: path or None}."""
    return {"divsufsort": build(force), "lzbuffer": build_lzbuffer(force), "coder": build_coder(force), "zpaql": build_zpaql(force),
            "predictor": build_predictor(force)}


if __name__ == "__main__":
    r5 = build_predictor(force=True)
    print(r5 or "reference predictor did not build")
    r4 = build_zpaql(force=True)
    print(r4 or "reference ZPAQL interpreter did not build")
    r3 = build_coder(force=True)
    print(r3 or "reference arithmetic coder did not build")
    r = build(force=True)
    print(r or "reference suffix sorter did not build")
    r2 = build_lzbuffer(force=True)
    print(r2 or "reference LZBuffer did not build")
    sys.exit(0 if r and r2 and r3 and r4 and r5 else 1)
