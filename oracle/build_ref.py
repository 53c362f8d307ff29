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
  * the block framing on top of that: Compressor.writeTag / startBlock / startSegment / postProcess / compress / endSegment /
    endBlock, ZPAQL.read / write, Encoder.init / compress -> libcompressor_ref.so (build_compressor()); and the way back:
    Decompresser.findBlock / findFilename / readComment / decompress / readSegmentEnd, Decoder.decompress / skip / init,
    PostProcessor.init / write -> libdecompresser_ref.so (build_decompresser())
  * makeConfig, itos and the numeric-method expansion inside compressBlock (LibZPAQ.cs:125-290, 360-367, 388-1044), still
    C++ text -> libfrontend_ref.so (build_frontend())

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


COMP_SRC = "/root/reference/ZPAQSharp/Compressor.cs"
COMP_OUT = os.path.join(OUT_DIR, "libcompressor_ref.so")


def compressor_text() -> str:
    """The block FRAMING of the reference: Compressor.writeTag / startBlock(level) / startBlock(hcomp) / startSegment /
    postProcess / compress / endSegment / endBlock (Compressor.cs:27-99, 133-249, 294-299), ZPAQL.read / write
    (ZPAQL.cs:112-179) and Encoder.init / compress / encode (Encoder.cs:26-103) as they lie, on top of the predictor
    fragment.  Textual repairs: C# `char[] x` -> `const char* x`, `= null` -> `= 0`, `Reader` / `Writer` parameters ->
    the harness's handle types, one comment whose second slash the formatter ate, and postProcess's C#-ified signature
    (`string pcomp, string comment`; its body still uses the C++ one: `const char* pcomp, int len`)."""
    base = predictor_text()
    fix = lambda t: re.sub(r"Array\.Resize\(ref\s+([\w\.]+),\s*", r"\1.resize(", t).replace(".Length", ".size()").replace("@", "")
    cs = lambda t: re.sub(r"\bchar\[\]\s+(\w+)", r"const char* \1", fix(t)).replace("= null", "= 0")
    rd = cs(_method_text(ZPAQL_SRC, "int read(Reader in2) // Read header")).replace("int read(Reader in2)", "int read(ReaderRef in2)")
    rd = re.sub(r"(?m)^(\s*)/ Get header", r"\1// Get header", rd)
    wr = cs(_method_text(ZPAQL_SRC, "bool write(Writer out2, bool pp)")).replace("bool write(Writer out2, bool pp)", "bool write(WriterRef out2, bool pp)")
    e_init = cs(_method_text(ENC_SRC, "void init()"))
    e_comp = cs(_method_text(ENC_SRC, "void compress(int c) // c is 0..255 or EOF"))
    e_enc = _method_text(ENC_SRC, "void encode(int y, int p)")
    c_tag = cs(_method_text(COMP_SRC, "void writeTag()"))
    c_lvl = cs(_method_text(COMP_SRC, "void startBlock(int level)"))
    c_hc = cs(_method_text(COMP_SRC, "void startBlock(char[] hcomp)"))
    c_seg = cs(_method_text(COMP_SRC, "void startSegment(char[] filename = null, char[] comment = null)"))
    c_pp = cs(_method_text(COMP_SRC, "void postProcess(string pcomp = null, string comment = null)"))
    assert "postProcess(string pcomp = 0, string comment = 0)" in c_pp
    c_pp = c_pp.replace("postProcess(string pcomp = 0, string comment = 0)", "postProcess(const char* pcomp = 0, int len = 0)")
    c_cmp = cs(_method_text(COMP_SRC, "bool compress(int n = -1)"))
    c_end = cs(_method_text(COMP_SRC, "void endSegment(char[] sha1string = null)"))
    c_eb = cs(_method_text(COMP_SRC, "void endBlock()"))
    # ZPAQL.read uses compsize: it has to be declared in front of the interpreter struct
    csz = "static const int compsize[256] = {0, 2, 3, 2, 3, 4, 6, 6, 3, 5};          // Component.cs:27-43\n"
    assert base.count(csz) == 1 and base.count("struct RefVM {") == 1
    base = base.replace(csz, "").replace("struct RefVM {", csz + r"""
struct RefSink { unsigned char* p; unsigned long long cap, n; void put(int c) { if (n < cap) p[n] = (unsigned char)c; ++n; } };
struct Reader { virtual int get() = 0;                                       // libzpaq Reader (LICENSE / libzpaq.h): get(), read()
                virtual int read(char* buf, int n) { int i = 0, c; while (i < n && (c = get()) >= 0) buf[i++] = (char)c; return i; } };
struct ReaderRef { Reader* r; ReaderRef(Reader* r_ = 0) : r(r_) {} int get() { return r->get(); } int read(char* b, int n) { return r->read(b, n); } };
struct WriterRef { RefSink* s; WriterRef(RefSink* s_ = 0) : s(s_) {} void put(int c) { s->put(c & 255); }
                   void write(const char* b, int n) { for (int i = 0; i < n; ++i) put(b[i]); } };
struct RefVM {""", 1)
    tail_vm = "  uint H(int i) { return hv(i); }\n};"
    assert base.count(tail_vm) == 1
    base = base.replace(tail_vm, "  uint H(int i) { return hv(i); }\n  void* rcode; int rcode_size;\n" + rd + "\n" + wr + "\n};", 1)
    dup = "struct RefSink { unsigned char* p; unsigned long long cap, n; void put(int c) { if (n < cap) p[n] = (unsigned char)c; ++n; } };\nstruct RefBlockEncoder {"
    assert base.count(dup) == 1
    base = base.replace(dup, "struct RefBlockEncoder {", 1)
    return base + r"""
static int toushort(const char* p) { return (p[0] & 255) + 256 * (p[1] & 255); }      // libzpaq toU16
struct RefPredictorN : RefPredictor { int predict() { return predict0(); } void update(int y) { update0(y); } };   // Predictor.cs:176-180, 200-204 (NOJIT)
struct RefEncoder {                                    // Encoder.cs:15-103: init, compress, encode are reference text
  WriterRef out; uint low, high; RefPredictorN pr; Arr<char> buf;
""" + e_init + "\n" + e_comp + "\n" + e_enc + r"""
};
struct MemoryReader : Reader {                         // Compressor.cs:324-340
  const char* p; MemoryReader(const char* p_, int off) : p(p_ + off) {} MemoryReader(MemoryReader* o) : p(o->p) {}
  int get() { return *p++ & 255; }
};
struct BoundedReader : Reader { const unsigned char* p; unsigned long long n, i; int get() { return i < n ? p[i++] : -1; } };
struct RefSHA1 { void put(int) {} unsigned long long usize() { return 0; } const char* result() { return 0; } };
struct RefPZ { Arr<byte> header; int hbegin, hend; RefSHA1* sha1; void initp() {} void run(int) {} void flush() {} };
enum State { INIT, BLOCK1, SEG1, BLOCK2, SEG2 };       // Compressor.cs:314-322
struct RefCompressor {
  RefEncoder enc; RefVM& z; RefPZ pz; ReaderRef in; RefSHA1 sha1; State state; bool verify;
  RefCompressor() : z(enc.pr.z), state(INIT), verify(false) { pz.hbegin = pz.hend = 0; pz.sha1 = 0; }
""" + "\n".join([c_tag, c_lvl, c_hc, c_seg, c_pp, c_cmp, c_end, c_eb]) + r"""
};
// One archive block exactly as LibZPAQ.compressBlock drives the Compressor (LibZPAQ.cs:296-325): [writeTag,] startBlock,
// startSegment(filename, comment), postProcess(pcomp, len), compress() to EOF, endSegment(sha1), endBlock.
// level > 0: startBlock(int level) and the reference's own model table; else startBlock(hcomp) with `hdr`.
extern "C" long long ref_compress_block(int level, const unsigned char* hdr, const unsigned char* pcomp, int pcomp_len,
                                        const char* filename, const char* comment, const unsigned char* input, unsigned long long n,
                                        const unsigned char* sha1string, int with_tag, unsigned char* out, unsigned long long cap) {
  try {
    RefCompressor* co = new RefCompressor();
    RefPredictor& P = co->enc.pr;
    P.initTables = false; P.pcode = 0; P.pcode_size = 0; P.c8 = 1; P.hmap4 = 1;
    RefSink sink; sink.p = out; sink.cap = cap; sink.n = 0;
    co->enc.out = WriterRef(&sink); co->enc.low = 1; co->enc.high = 0xFFFFFFFFu;
    if (with_tag) co->writeTag();
    if (level > 0) co->startBlock(level); else co->startBlock((const char*)hdr);
    co->startSegment(filename, comment);
    co->postProcess(pcomp_len > 0 ? (const char*)pcomp : 0, pcomp_len);
    BoundedReader br; br.p = input; br.n = n; br.i = 0;
    co->in = ReaderRef(&br);
    while (co->compress(1 << 20)) ;
    co->endSegment((const char*)sha1string);
    co->endBlock();
    delete co;
    return (long long)sink.n;
  } catch (const std::exception&) { return -1; }
}
"""


def build_compressor(force: bool = False) -> str | None:
    if not (os.path.exists(COMP_SRC) and os.path.exists(PRED_SRC) and os.path.exists(ZPAQL_SRC)):
        return COMP_OUT if os.path.exists(COMP_OUT) else None
    if os.path.exists(COMP_OUT) and not force and os.path.getmtime(COMP_OUT) >= os.path.getmtime(__file__):
        return COMP_OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    p = subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-w", "-fpermissive", "-std=c++14", "-x", "c++", "-", "-o", COMP_OUT],
                       input=compressor_text().encode(), capture_output=True)
    if p.returncode != 0:
        sys.stderr.write(p.stderr.decode()[:8000])
        return None
    return COMP_OUT


DECOMP_SRC = "/root/reference/ZPAQSharp/Decompresser.cs"
PP_SRC = "/root/reference/ZPAQSharp/PostProcessor.cs"
DECOMP_OUT = os.path.join(OUT_DIR, "libdecompresser_ref.so")


def decompresser_text() -> str:
    """The DECODE side of the framing: Decompresser.findBlock / findFilename / readComment / decompress / readSegmentEnd
    (Decompresser.cs:29-194), Decoder.decompress / skip / init / decode (Decoder.cs:32-111, 136-158), PostProcessor.init /
    write / getState (PostProcessor.cs:27-91) and ZPAQL.flush (ZPAQL.cs:194-199) as they lie, on top of the compressor
    fragment, driven like LibZPAQ.decompress (LibZPAQ.cs:65-79).  Textual repairs as before plus: `nt c` -> `int c`
    (Decoder.cs:70), and the C#-ified signatures of findFilename / readComment / readSegmentEnd whose bodies still use the
    C++ parameter names (filename, comment, sha1string as a writable buffer).  Supplied by the harness because their C#
    text is not C++ any more: Decoder.get (unbuffered here), ZPAQL.outc (Decoder.cs:113-123, ZPAQL.cs:201-207: the
    comma of libzpaq's `(outbuf[bufptr]=ch, ++bufptr==size)` became `&&`), ZPAQL.clear / initp (H and M are `hv` / `mv` here).
    Also ZPAQL.memory / pow2 (ZPAQL.cs:58-81, 1306-1311) as they lie -> ref_block_memory."""
    base = compressor_text()
    fix = lambda t: re.sub(r"Array\.Resize\(ref\s+([\w\.]+),\s*", r"\1.resize(", t).replace(".Length", ".size()").replace("@", "")
    cs = lambda t: fix(t).replace("= null", "= 0")
    flush = cs(_method_text(ZPAQL_SRC, "void flush() // write outbuf"))
    mem = cs(_method_text(ZPAQL_SRC, "double memory() // Return memory requirement in bytes"))
    pw = _method_text(ZPAQL_SRC, "static double pow2(int x)")
    d_dec = cs(_method_text(DEC_SRC, "int decompress() // return a byte or EOF"))
    d_skip = cs(_method_text(DEC_SRC, "int skip() // skip to the end of the segment"))
    assert "\tnt c = -1;" in d_skip.replace("    ", "\t") or "nt c = -1;" in d_skip
    d_skip = re.sub(r"(?m)^(\s*)nt c = -1;", r"\1int c = -1;", d_skip)
    d_init = cs(_method_text(DEC_SRC, "void init() // initialize at start of block"))
    d_bit = _method_text(DEC_SRC, "int decode(int p)")
    p_init = cs(_method_text(PP_SRC, "void init(int h, int m)"))
    p_write = cs(_method_text(PP_SRC, "int write(int c) // Input a byte, return state"))
    p_state = cs(_method_text(PP_SRC, "int getState()"))
    f_blk = cs(_method_text(DECOMP_SRC, "bool findBlock(double* memptr = null)"))
    f_name = cs(_method_text(DECOMP_SRC, "bool findFilename(Writer out2 = null)")).replace("findFilename(Writer out2 = 0)", "findFilename(WriterRef filename = 0)")
    f_cmt = cs(_method_text(DECOMP_SRC, "void readComment(Writer out2 = null)")).replace("readComment(Writer out2 = 0)", "readComment(WriterRef comment = 0)")
    f_dec = cs(_method_text(DECOMP_SRC, "bool decompress(int n = -1)"))
    f_end = cs(_method_text(DECOMP_SRC, "void readSegmentEnd(char[] sha1string = null)")).replace("readSegmentEnd(char[] sha1string = 0)", "readSegmentEnd(char* sha1string = 0)")
    assert "WriterRef filename" in f_name and "WriterRef comment" in f_cmt and "char* sha1string" in f_end
    stub = "  void outc(int) {}\n"
    assert base.count(stub) == 1
    base = base.replace(stub, r"""  WriterRef output; ShaRef sha1; Arr<char> outbuf; int bufptr;
  void outc(int ch) { if (ch < 0 || (outbuf[bufptr] = ch, ++bufptr == (int)outbuf.size())) flush(); }   // libzpaq ZPAQL::outc
""" + flush + r"""
  void clear() { cend = hbegin = hend = 0; a = b = c = d = 0; f = pc = 0; header.resize(0); hv.resize(0); mv.resize(0); memset(r, 0, sizeof r); }  // ZPAQL.cs:33-42
  void initp() { hv.resize(1, header[4]); mv.resize(1, header[5]); memset(r, 0, sizeof r); a = b = c = d = 0; f = 0; pc = 0; }   // ZPAQL.cs:52-56, 1010-1026
""" + mem + r"""
""", 1)
    wref = "struct WriterRef { RefSink* s; WriterRef(RefSink* s_ = 0) : s(s_) {} void put(int c) { s->put(c & 255); }"
    assert base.count(wref) == 1
    base = base.replace(wref, wref + " explicit operator bool() const { return s != 0; }", 1)
    base = base.replace("struct RefVM {", "struct ShaRef { explicit operator bool() const { return false; } void write(const char*, int) {} };\nstruct RefVM {", 1)
    en = "enum { NONE, CONS, CM, ICM, MATCH, AVG, MIX2, MIX, ISSE, SSE };\n"
    assert base.count(en) == 1
    base = base.replace(en, "", 1).replace("struct ShaRef {", en + pw + "\nstruct ShaRef {", 1)
    return base + r"""
struct RefDecoder : Reader {                           // Decoder.cs:16-159: decompress, skip, init, decode are reference text
  ReaderRef in; uint low, high, curr; RefPredictorN pr;
  int get() { return in.get(); }                       // Decoder.cs:113-123 reads `in` through a 64 KB buffer
""" + "\n".join([d_dec, d_skip, d_init, d_bit]) + r"""
};
struct RefPostProcessor {                              // PostProcessor.cs:11-101
  int state, hsize, ph, pm; RefVM z;
""" + "\n".join([p_init, p_write, p_state]) + r"""
};
enum DState { BLOCK, FILENAME, COMMENT, DATA, SEGEND };   // Decompresser.cs:206-219
enum DecodeState { FIRSTSEG, SEG, SKIP };
struct RefDecompresser {
  RefDecoder dec; RefVM& z; RefPostProcessor pp; DState state; DecodeState decode_state;
  RefDecompresser() : z(dec.pr.z), state(BLOCK), decode_state(FIRSTSEG) {
    RefPredictor& P = dec.pr; P.initTables = false; P.pcode = 0; P.pcode_size = 0; P.c8 = 1; P.hmap4 = 1;
    dec.low = 1; dec.high = 0xFFFFFFFFu; dec.curr = 0;
    z.outbuf.resize(1 << 14); z.bufptr = 0; pp.z.outbuf.resize(1 << 14); pp.z.bufptr = 0;
    pp.state = pp.hsize = pp.ph = pp.pm = 0; pp.z.cend = pp.z.hbegin = pp.z.hend = 0; pp.z.rcode = 0; pp.z.rcode_size = 0; z.rcode = 0; z.rcode_size = 0;
  }
""" + "\n".join([f_blk, f_name, f_cmt, f_dec, f_end]) + r"""
};
// LibZPAQ.decompress (LibZPAQ.cs:65-79): every segment of every block of `arc` into `out`; per segment 21 bytes into
// `marks` (readSegmentEnd: 0 = no checksum, else 1 + the stored SHA-1).  Returns bytes written, -1 on a reference error.
// ZPAQL.memory() (ZPAQL.cs:58-81) of a block header as stored in an archive (read by ZPAQL.read).
extern "C" double ref_block_memory(const unsigned char* hdr) {
  try { RefVM* v = new RefVM(); v->rcode = 0; v->rcode_size = 0; MemoryReader m((const char*)hdr, 0); v->read(&m); double r = v->memory(); delete v; return r; }
  catch (const std::exception&) { return -1; }
}
extern "C" long long ref_decompress(const unsigned char* arc, unsigned long long n, unsigned char* out, unsigned long long cap,
                                    unsigned char* marks, int max_segs, int* nsegs) {
  *nsegs = 0;
  try {
    RefDecompresser* d = new RefDecompresser();
    BoundedReader br; br.p = arc; br.n = n; br.i = 0;
    d->dec.in = ReaderRef(&br);
    RefSink sink; sink.p = out; sink.cap = cap; sink.n = 0;
    d->pp.z.output = WriterRef(&sink);
    while (d->findBlock()) {
      while (d->findFilename()) {
        d->readComment();
        d->decompress();
        char s[21]; memset(s, 0, sizeof s);
        d->readSegmentEnd(s);
        if (*nsegs < max_segs) memcpy(marks + 21 * *nsegs, s, 21);
        ++*nsegs;
      }
    }
    delete d;
    return (long long)sink.n;
  } catch (const std::exception&) { return -1; }
}
"""


def build_decompresser(force: bool = False) -> str | None:
    if not all(os.path.exists(f) for f in (DECOMP_SRC, PP_SRC, DEC_SRC, COMP_SRC, PRED_SRC, ZPAQL_SRC)):
        return DECOMP_OUT if os.path.exists(DECOMP_OUT) else None
    if os.path.exists(DECOMP_OUT) and not force and os.path.getmtime(DECOMP_OUT) >= os.path.getmtime(__file__):
        return DECOMP_OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    p = subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-w", "-fpermissive", "-std=c++14", "-x", "c++", "-", "-o", DECOMP_OUT],
                       input=decompresser_text().encode(), capture_output=True)
    if p.returncode != 0:
        sys.stderr.write(p.stderr.decode()[:8000])
        return None
    return DECOMP_OUT


LIBZ_SRC = "/root/reference/ZPAQSharp/LibZPAQ.cs"
FRONT_OUT = os.path.join(OUT_DIR, "libfrontend_ref.so")


def frontend_text() -> str:
    """The method-string front end of the reference, still pristine C++ inside LibZPAQ.cs: makeConfig (LibZPAQ.cs:388-1044,
    method string -> config text + args[9]), itos (:360-367), and the statement range of compressBlock that derives the
    block-size argument and expands a numeric method "LB,R,t" into its x-method (:125-141, 158-290; the SHA-1 statements
    in between are left out).  lg / nbits come from LZBuffer.cs:118-135.  Repairs: `char[] method` -> `const char*`,
    `.Length` -> `.size()`, C# `in` / `@in` -> the harness's buffer, `MAX`."""
    mk = _method_text(LIBZ_SRC, "string makeConfig(char[] method, int args[])").replace(
        "string makeConfig(char[] method, int args[])", "string makeConfig(const char* method, int args[])").replace(".Length", ".size()")
    itos = _method_text(LIBZ_SRC, "string itos(long x, int n = 1)")
    lg = _method_text(LZ_SRC, "int lg(unsigned x)")
    nb = _method_text(LZ_SRC, "int nbits(unsigned x)")
    lines = open(LIBZ_SRC, encoding="utf-8-sig").read().split("\n")
    i0 = next(i for i, l in enumerate(lines) if "const unsigned n =in.Length;" in l)
    i1 = next(i for i, l in enumerate(lines) if "// Get hash of input" in l)
    i2 = next(i for i, l in enumerate(lines) if "// Expand default methods" in l)
    i3 = next(i for i, l in enumerate(lines) if l.strip() == "// Compress" and i > i2)
    assert i0 < i1 < i2 < i3
    expand = "\n".join(lines[i0:i1] + lines[i2:i3])
    expand = expand.replace("=in.Length", "= in.size()").replace("in.data()", "in.data()").replace(".Length", ".size()").replace("@", "")
    return r"""
#include <string>
#include <vector>
#include <ctype.h>
#include <stdlib.h>
#include <string.h>
#include <stdexcept>
using std::string;
#define assert(x) ((void)0)
#define MAX(a, b) ((a) > (b) ? (a) : (b))
static void error(const char* msg) { throw std::runtime_error(msg); }
""" + "\n".join([lg, nb, itos, mk]) + r"""
struct InBuf { const unsigned char* p; unsigned n; unsigned size() const { return n; } const unsigned char* data() const { return p; } };
static string expandMethod(InBuf in, string method) {
""" + expand + r"""
  return method;
}
extern "C" int ref_make_config(const char* method, int* args, char* out, int cap) {
  try { string s = makeConfig(method, args); if ((int)s.size() >= cap) return -2; memcpy(out, s.c_str(), s.size() + 1); return (int)s.size(); }
  catch (const std::exception&) { return -1; }
}
extern "C" int ref_expand_method(const char* method, const unsigned char* data, unsigned n, char* out, int cap) {
  try { InBuf in; in.p = data; in.n = n; string s = expandMethod(in, method); if ((int)s.size() >= cap) return -2;
        memcpy(out, s.c_str(), s.size() + 1); return (int)s.size(); }
  catch (const std::exception&) { return -1; }
}
"""


def build_frontend(force: bool = False) -> str | None:
    if not (os.path.exists(LIBZ_SRC) and os.path.exists(LZ_SRC)):
        return FRONT_OUT if os.path.exists(FRONT_OUT) else None
    if os.path.exists(FRONT_OUT) and not force and os.path.getmtime(FRONT_OUT) >= os.path.getmtime(__file__):
        return FRONT_OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    p = subprocess.run(["g++", "-O1", "-fPIC", "-shared", "-w", "-fpermissive", "-std=c++14", "-x", "c++", "-", "-o", FRONT_OUT],
                       input=frontend_text().encode(), capture_output=True)
    if p.returncode != 0:
        sys.stderr.write(p.stderr.decode()[:8000])
        return None
    return FRONT_OUT


def build_all(force: bool = False) -> dict:
    """Every fragment; returns {name: path or None}."""
    return {"divsufsort": build(force), "lzbuffer": build_lzbuffer(force), "coder": build_coder(force), "zpaql": build_zpaql(force),
            "predictor": build_predictor(force), "compressor": build_compressor(force),
            "decompresser": build_decompresser(force), "frontend": build_frontend(force)}


if __name__ == "__main__":
    r6 = build_compressor(force=True)
    print(r6 or "reference Compressor framing did not build")
    r7 = build_decompresser(force=True)
    print(r7 or "reference Decompresser did not build")
    r8 = build_frontend(force=True)
    print(r8 or "reference makeConfig / method expansion did not build")
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
    sys.exit(0 if r and r2 and r3 and r4 and r5 and r6 and r7 and r8 else 1)
