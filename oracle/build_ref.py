"""Builds the fragments of the reference that are compilable here:

  * the body of /root/reference/ZPAQSharp/divsufsort.cs (libdivsufsort-lite, still C text inside a C# class)
    -> oracle/_ref/libdivsufsort_ref.so                                   (build())
  * the body of /root/reference/ZPAQSharp/LZBuffer.cs (the LZ77 / BWT pre-processor, still C++ text) together with
    `e8e9` from LibZPAQ.cs:371-384, around a harness of the three library classes the reference does not contain
    (Array<T> as documented in LICENSE:618-634, Reader, StringBuffer) -> oracle/_ref/liblzbuffer_ref.so  (build_lzbuffer())

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


if __name__ == "__main__":
    r = build(force=True)
    print(r or "reference suffix sorter did not build")
    r2 = build_lzbuffer(force=True)
    print(r2 or "reference LZBuffer did not build")
    sys.exit(0 if r and r2 else 1)
