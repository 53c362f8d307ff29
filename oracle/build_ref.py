"""Builds the one fragment of the reference that is compilable here: the body of
/root/reference/ZPAQSharp/divsufsort.cs (libdivsufsort-lite, still C text inside a C# class).

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


if __name__ == "__main__":
    r = build(force=True)
    print(r or "reference fragment did not build")
    sys.exit(0 if r else 1)
