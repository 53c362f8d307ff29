"""oracle/frontend.py -- CPU ORACLE front end (TEST INFRASTRUCTURE ONLY; never imported by
zpaqsharp_b200/).

Restates the string-processing half of the reference's compress path:

  expand_method()   LibZPAQ.cs:117-283   level digit "LB,R,t" -> "x..." method string
  make_config()     LibZPAQ.cs:388-1044  method string -> ZPAQL config text + args[9]
  compile_config()  Compiler.cs:13-478   config text -> block header bytes + PCOMP bytes
  builtin_model()   Compressor.cs:45-83  min/mid/max models (kept here as config source and
                                         checked against the reference bytecode by the tests)

The byte/bit half (codec, LZ77, BWT, framing) is oracle/zpq_oracle.cpp.

Parity status: the only material in the reference that pins this code is (a) the min.cfg source
<-> bytecode pair (LICENSE:391-400 <-> Compressor.cs:49-50) and (b) the mid/max bytecodes
(Compressor.cs:53-72), both in tests/golden/reference_kat.json.  The Compiler helpers
Next/MatchToken/RToken are corrupt in the C# text and are restored from SURVEY.md appendix B.
"""
from __future__ import annotations

import numpy as np

COMPSIZE = [0, 2, 3, 2, 3, 4, 6, 6, 3, 5]           # Component.cs:27-43
COMPNAME = ["", "const", "cm", "icm", "match", "avg", "mix2", "mix", "isse", "sse"]  # Compiler.cs:516


def lg(x: int) -> int:        # LZBuffer.cs:118-127
    return int(x).bit_length()


def nbits(x: int) -> int:     # LZBuffer.cs:130-135
    return bin(int(x)).count("1")


# ------------------------------------------------------------------------------------------
# ZPAQL mnemonics, Compiler.cs:535-569.  Generated from the ISA's field structure
# (ZPAQL.cs:256-321) instead of being listed.
# ------------------------------------------------------------------------------------------
def _opcode_names():
    reg = ["a", "b", "c", "d", "*b", "*c", "*d"]
    names = [""] * 256
    for d, r in enumerate(reg):
        names[d * 8 + 0] = r + "<>a"
        names[d * 8 + 1] = r + "++"
        names[d * 8 + 2] = r + "--"
        names[d * 8 + 3] = r + "!"
        names[d * 8 + 4] = r + "=0"
    names[0] = "error"
    for d in range(4):
        names[d * 8 + 7] = reg[d] + "=r"
    names[39], names[47], names[55] = "jt", "jf", "r=a"
    names[56], names[57], names[59], names[60], names[63] = "halt", "out", "hash", "hashd", "jmp"
    for d, r in enumerate(reg):
        for s, q in enumerate(reg + [""]):
            names[64 + d * 8 + s] = r + "=" + q
    ops = ["+=", "-=", "*=", "/=", "%=", "&=", "&~", "|=", "^=", "<<=", ">>=", "==", "<", ">"]
    for x, o in enumerate(ops):
        for s, q in enumerate(reg + [""]):
            names[128 + x * 8 + s] = "a" + o + q
    names[255] = "lj"
    names += ["post", "pcomp", "end", "if", "ifnot", "else", "endif", "do",
              "while", "until", "forever", "ifl", "ifnotl", "elsel", ";"]
    return names


OPCODES = _opcode_names()
(POST, PCOMP, END, IF, IFNOT, ELSE, ENDIF, DO, WHILE, UNTIL, FOREVER, IFL, IFNOTL, ELSEL, SEMI) = range(256, 271)
JT, JF, JMP, LJ = 39, 47, 63, 255


class ConfigError(Exception):
    pass


class _Compiler:
    """Compiler.cs:13-478."""

    def __init__(self, text: str, args):
        self.s = text + "\0"
        self.i = 0
        self.args = list(args) if args is not None else [0] * 9
        self.state = 0
        self.line = 1
        self.pcomp_cmd = ""

    # restored: SURVEY appendix B, Compiler.cs:191-226
    def next(self):
        s = self.s
        while s[self.i] != "\0":
            ch = s[self.i]
            if ch == "\n":
                self.line += 1
            if ch == "(":
                self.state += 1 + (1 if self.state < 0 else 0)
            elif self.state > 0 and ch == ")":
                self.state -= 1
            elif self.state < 0 and ch <= " ":
                self.state = 0
            elif self.state == 0 and ch > " ":
                self.state = -1
                break
            self.i += 1
        if s[self.i] == "\0":
            raise ConfigError("unexpected end of config")

    # restored: Compiler.cs:229-240
    def match(self, word: str) -> bool:
        a, k = self.i, 0
        s = self.s
        while s[a] > " " and s[a] != "(" and k < len(word):
            if s[a].lower() != word[k].lower():
                return False
            a += 1
            k += 1
        return k == len(word) and (s[a] <= " " or s[a] == "(")

    def _atoi(self, pos: int) -> int:
        s = self.s
        while s[pos] in " \t\n\r\v\f":
            pos += 1
        sign = 1
        if s[pos] in "+-":
            sign = -1 if s[pos] == "-" else 1
            pos += 1
        v = 0
        while s[pos].isdigit():
            v = v * 10 + ord(s[pos]) - 48
            pos += 1
        return sign * v

    # restored: Compiler.cs:244-276
    def number(self, low: int, high: int) -> int:
        self.next()
        s, i = self.s, self.i
        r = 0
        if s[i] == "$" and "1" <= s[i + 1] <= "9":
            if s[i + 2] == "+":
                r = self._atoi(i + 3)
            r += self.args[ord(s[i + 1]) - ord("1")]
        elif s[i] == "-" or s[i].isdigit():
            r = self._atoi(i)
        else:
            raise ConfigError("line %d: expected a number" % self.line)
        if r < low:
            raise ConfigError("line %d: number too low" % self.line)
        if r > high:
            raise ConfigError("line %d: number too high" % self.line)
        return r

    def expect(self, word: str):
        self.next()
        if not self.match(word):
            raise ConfigError("line %d: expected %s" % (self.line, word))

    def one_of(self, names) -> int:
        self.next()
        for k, w in enumerate(names):
            if w and self.match(w):
                return k
        raise ConfigError("line %d: unexpected token %r" % (self.line, self.s[self.i:self.i + 12]))

    # Compiler.cs:319-478.  Returns (program bytes incl. trailing 0, terminating token).
    def program(self):
        code = bytearray()
        if_stack, do_stack = [], []

        def pop(st):
            if not st:
                raise ConfigError("unmatched IF or DO")
            return st.pop()

        while True:
            op = self.one_of(OPCODES)
            if op in (POST, PCOMP, END):
                break
            operand = operand2 = -1
            if op == IF:
                op, operand = JF, 0
                if_stack.append(len(code) + 1)
            elif op == IFNOT:
                op, operand = JT, 0
                if_stack.append(len(code) + 1)
            elif op in (IFL, IFNOTL):
                code.append(JT if op == IFL else JF)
                code.append(3)
                op, operand, operand2 = LJ, 0, 0
                if_stack.append(len(code) + 1)
            elif op in (ELSE, ELSEL):
                if op == ELSE:
                    op, operand = JMP, 0
                else:
                    op, operand, operand2 = LJ, 0, 0
                a = pop(if_stack)
                if code[a - 1] != LJ:
                    j = len(code) - a + 1 + (1 if op == LJ else 0)
                    if j > 127:
                        raise ConfigError("IF too big, try IFL, IFNOTL")
                    code[a] = j
                else:
                    j = len(code) + 2 + (1 if op == LJ else 0)
                    code[a] = j & 255
                    code[a + 1] = (j >> 8) & 255
                if_stack.append(len(code) + 1)
            elif op == ENDIF:
                a = pop(if_stack)
                j = len(code) - a - 1
                if code[a - 1] != LJ:
                    if j > 127:
                        raise ConfigError("IF too big, try IFL, IFNOTL, ELSEL")
                    code[a] = j
                else:
                    j = len(code)
                    code[a] = j & 255
                    code[a + 1] = (j >> 8) & 255
            elif op == DO:
                do_stack.append(len(code))
            elif op in (WHILE, UNTIL, FOREVER):
                a = pop(do_stack)
                j = a - len(code) - 2
                if j >= -127:
                    op = {WHILE: JT, UNTIL: JF, FOREVER: JMP}[op]
                    operand = j & 255
                else:
                    j = a
                    if op == WHILE:
                        code += bytes([JF, 3])
                    if op == UNTIL:
                        code += bytes([JT, 3])
                    op, operand, operand2 = LJ, j & 255, j >> 8
            elif (op & 7) == 7:
                if op == LJ:
                    operand = self.number(0, 65535)
                    operand2 = operand >> 8
                    operand &= 255
                elif op in (JT, JF, JMP):
                    operand = self.number(-128, 127) & 255
                else:
                    operand = self.number(0, 255)
            if 0 <= op <= 255:
                code.append(op)
            if operand >= 0:
                code.append(operand)
            if operand2 >= 0:
                code.append(operand2)
            if len(code) > 65535 - 8:
                raise ConfigError("program too big")
        code.append(0)
        return bytes(code), op


def compile_config(text: str, args=None):
    """Compiler.cs:13-111.  Returns (hdr, pcomp, pcomp_cmd): `hdr` is the block header as written
    to the archive by ZPAQL.write(out,false) (ZPAQL.cs:158-179); `pcomp` is the PCOMP program
    including its END byte (empty when the config has none)."""
    c = _Compiler(text, args)
    c.expect("comp")
    head = bytearray(7)
    for k in range(2, 7):
        head[k] = c.number(0, 255)
    n = head[6]
    comp = bytearray()
    for i in range(n):
        c.number(i, i)
        t = c.one_of(COMPNAME)
        comp.append(t)
        clen = COMPSIZE[t]
        if clen < 1:
            raise ConfigError("invalid component")
        for _ in range(1, clen):
            comp.append(c.number(0, 255))
    comp.append(0)
    c.expect("hcomp")
    hprog, op = c.program()
    hsize = 7 + len(comp) - 2 + len(hprog)
    head[0], head[1] = hsize & 255, hsize >> 8
    hdr = bytes(head) + bytes(comp) + hprog
    pcomp = b""
    if op == POST:
        c.number(0, 0)
        c.expect("end")
    elif op == PCOMP:
        c.next()
        j = c.i
        while c.s[j] not in "\0;":
            j += 1
        c.pcomp_cmd = c.s[c.i:j]
        c.i = j + (1 if c.s[j] == ";" else 0)
        pcomp, op = c.program()
        if op != END:
            raise ConfigError("expected END")
    elif op != END:
        raise ConfigError("expected END or POST 0 END or PCOMP cmd ; ... END")
    return hdr, pcomp, c.pcomp_cmd


# ------------------------------------------------------------------------------------------
# Built-in models, Compressor.cs:45-83 (bytecode there; config source here, own formatting).
# ------------------------------------------------------------------------------------------
MIN_CFG = """comp 1 2 0 0 2
  0 icm 16
  1 isse 19 0
hcomp
  *b=a a=0 d=0 hash b-- hash *d=a
  d++ b-- hash b-- hash *d=a
  halt
end
"""

MID_CFG = """comp 3 3 0 0 8
  0 icm 5
  1 isse 13 0
  2 isse 17 1
  3 isse 18 2
  4 isse 18 3
  5 isse 19 4
  6 match 22 24
  7 mix 16 0 7 24 255
hcomp
  c++ *c=a b=c a=0
  d= 1 hash *d=a
  b-- d++ hash *d=a
  b-- d++ hash *d=a
  b-- d++ hash *d=a
  b-- d++ hash *d=a
  b-- d++ hash b-- hash *d=a
  d++ a=*c a<<= 8 *d=a
  halt
end
"""

MAX_CFG = """comp 5 9 0 0 22
  0 const 160
  1 icm 5
  2 isse 13 1
  3 isse 16 2
  4 isse 18 3
  5 isse 19 4
  6 isse 19 5
  7 isse 20 6
  8 match 22 24
  9 icm 17
  10 isse 19 9
  11 icm 13
  12 icm 13
  13 icm 13
  14 icm 14
  15 mix 16 0 15 24 255
  16 mix 8 0 16 10 255
  17 mix2 0 15 16 24 0
  18 sse 8 17 32 255
  19 mix2 8 17 18 16 255
  20 sse 16 19 32 255
  21 mix2 0 19 20 16 0
hcomp
  c++ *c=a b=c a=0
  d= 2 hash *d=a b--
  d++ hash *d=a b--
  d++ hash *d=a b--
  d++ hash *d=a b--
  d++ hash *d=a b--
  d++ hash b-- hash *d=a b--
  d++ hash *d=a b--
  d++ a=*c a&~ 32
  a> 64 if
    a< 91 if
      d++ hashd d--
      *d<>a a+=*d a*= 20 *d=a
      jmp 9
    endif
  endif
  a=*d a== 0 ifnot
    d++ *d=a d--
  endif
  *d=0
  d++ d++ b=c b-- a=0 hash *d=a
  d++ b-- a=0 hash *d=a
  d++ b-- a=0 hash *d=a
  d++ a=b a-= 212 b=a a=0 hash
  *d=a b<>a a-= 216 b<>a a=*b a&= 60 hashd
  d++ a=*c a<<= 9 *d=a
  d++ d++ d++ d++ d++ *d=a
  halt
end
"""


def builtin_model(level: int):
    """startBlock(int level), Compressor.cs:45-83 -> (hdr, pcomp)."""
    src = {1: MIN_CFG, 2: MID_CFG, 3: MAX_CFG}.get(level)
    if src is None:
        raise ConfigError("compression level must be 1..3")
    hdr, pcomp, _ = compile_config(src, None)
    return hdr, pcomp


# ------------------------------------------------------------------------------------------
# Method expansion, LibZPAQ.cs:117-283
# ------------------------------------------------------------------------------------------
def block_arg0(n: int) -> int:                      # LibZPAQ.cs:125
    return max(lg(n + 4095) - 20, 0)


def expand_method(method: str, data: bytes) -> str:
    n = len(data)
    arg0 = block_arg0(n)
    if not method[0].isdigit():
        return method
    # type from "LB,R,t", LibZPAQ.cs:128-141
    commas, arg = 0, [0, 0, 0, 0]
    for ch in method[1:]:
        if commas >= 4:
            break
        if ch in ",.":
            commas += 1
        elif ch.isdigit() and commas < 4:
            arg[commas] = arg[commas] * 10 + ord(ch) - 48
    typ = 512 if commas == 0 else arg[1] * 4 + arg[2]
    level = ord(method[0]) - 48
    doe8 = (typ & 2) * 2
    m = "x%d" % arg0
    htsz = ",%d" % (19 + arg0 + (1 if arg0 <= 6 else 0))
    sasz = ",%d" % (21 + arg0)
    if level == 0:
        m = "0%d,0" % arg0
    elif level == 1:
        if typ < 40:
            m += ",0"
        else:
            m += ",%d," % (1 + doe8)
            if typ < 80:
                m += "4,0,1,15"
            elif typ < 128:
                m += "4,0,2,16"
            elif typ < 256:
                m += "4,0,2" + htsz
            elif typ < 960:
                m += "5,0,3" + htsz
            else:
                m += "6,0,3" + htsz
    elif level == 2:
        if typ < 32:
            m += ",0"
        else:
            m += ",%d," % (1 + doe8)
            if typ < 64:
                m += "4,0,3" + htsz
            else:
                m += "4,0,7" + sasz + ",1"
    elif level == 3:
        if typ < 20:
            m += ",0"
        elif typ < 48:
            m += ",%d,4,0,3%s" % (1 + doe8, htsz)
        elif typ >= 640 or (typ & 1):
            m += ",%dci1" % (3 + doe8)
        else:
            m += ",%d,12,0,7%s,1c0,0,511i2" % (2 + doe8, sasz)
    elif level == 4:
        if typ < 12:
            m += ",0"
        elif typ < 24:
            m += ",%d,4,0,3%s" % (1 + doe8, htsz)
        elif typ < 48:
            m += ",%d,5,0,7%s1c0,0,511" % (2 + doe8, sasz)
        elif typ < 900:
            m += ",%dci1,1,1,1,2a" % doe8
            if typ & 1:
                m += "w"
            m += "m"
        else:
            m += ",%dci1" % (3 + doe8)
    else:  # 5..9, LibZPAQ.cs:233-282
        m += ",%d" % doe8
        m += "w2c0,1010,255i1" if (typ & 1) else "w1i1"
        m += "c256ci1,1,1,1,1,1,2a"
        NR = 1 << 12
        r = [0] * NR
        if n:
            p = np.frombuffer(data, dtype=np.uint8)
            # gap to the previous occurrence of the same byte value (first occurrence: i - 0)
            order = np.argsort(p, kind="stable")
            sp = p[order]
            prev = np.zeros(n, dtype=np.int64)
            same = np.concatenate(([False], sp[1:] == sp[:-1]))
            prev_sorted = np.where(same, np.concatenate(([0], order[:-1])), 0)
            prev[order] = prev_sorted
            k = np.arange(n, dtype=np.int64) - prev
            k = k[(k > 0) & (k < NR)]
            r = np.bincount(k, minlength=NR).tolist()
        n1 = n - r[1] - r[2] - r[3]
        for _ in range(2):
            period, score, t = 0, 0.0, 0
            j = 5
            while j < NR and t < n1:
                s = r[j] / (256.0 + n1 - t)
                if s > score:
                    score, period = s, j
                t += r[j]
                j += 1
            if period > 4 and score > 0.1:
                m += "c0,0,%d,255i1" % (999 + period)
                if period <= 255:
                    m += "c0,%di1" % period
                n1 -= r[period]
                r[period] = 0
            else:
                break
        m += "c0,2,0,255i1c0,3,0,0,255i1c0,4,0,0,0,255i1mm16ts19t0"
    return m


# ------------------------------------------------------------------------------------------
# PCOMP programs emitted by makeConfig, LibZPAQ.cs:427-830.  ZPAQL source, comments dropped.
# ------------------------------------------------------------------------------------------
_E8E9_AT_EOF = """
      a=b a==d ifnot
        a+= 4 a<d if
          a=*b a&= 254 a== 232 if
            c=b b++ b++ b++ b++ a=*b a++ a&= 254 a== 0 if
              b-- a=*b
              b-- a<<= 8 a+=*b
              b-- a<<= 8 a+=*b
              a-=b a++
              *b=a a>>= 8 b++
              *b=a a>>= 8 b++
              *b=a b++
            endif
            b=c
          endif
        endif
        a=*b out b++
      forever
    endif
"""


def _pcomp_lazy2(args, doe8):   # LibZPAQ.cs:427-571
    rb = args[0] - 4 if args[0] > 4 else 0
    out = "" if doe8 else " out\n"
    p = "pcomp lazy2 3 ;\n  a> 255 if\n"
    if doe8:
        p += "    b=0 d=r 4 do" + _E8E9_AT_EOF
    p += """    a=0 b=0 c=0 d=0 r=a 1 r=a 2 r=a 3 r=a 4
    halt
  endif
  a<<=d a+=c c=a
  a= 8 a+=d d=a
  a=r 1 a== 0 if
    a= 1 r=a 2
    a=c a&= 3 a> 0 if
      a-- a<<= 3 r=a 3
      a=c a>>= 2 c=a
      b=r 3 a&= 7 a+=b r=a 3
      a=c a>>= 3 c=a
      a=d a-= 5 d=a
      a= 1 r=a 1
    else
      a=c a>>= 2 c=a
      d-- d--
      a= 3 r=a 1
    endif
  endif
  do a=r 1 a== 1 if a=d a> 2 if
    a=c a&= 1 a== 1 if
      a=c a>>= 1 c=a
      b=r 2 a=c a&= 1 a+=b a+=b r=a 2
      a=c a>>= 1 c=a
      d-- d--
    else
      a=c a>>= 1 c=a
      a=r 2 a<<= 2 b=a
      a=c a&= 3 a+=b r=a 2
      a=c a>>= 2 c=a
      d-- d-- d--
"""
    p += "      a= 5 r=a 1\n" if rb else "      a= 2 r=a 1\n"
    p += "    endif\n  forever endif endif\n"
    if rb:
        p += ("  a=r 1 a== 5 if a=d a> %d if\n    a=c a&= %d r=a 5\n    a=c a>>= %d c=a\n"
              "    a=d a-= %d d=a\n    a= 2 r=a 1\n  endif endif\n") % (rb - 1, (1 << rb) - 1, rb, rb)
    p += """  a=r 1 a== 2 if a=r 3 a>d ifnot
    a=c r=a 6 a=d r=a 7
    b=r 3 a= 1 a<<=b d=a
    a-- a&=c a+=d
"""
    if rb:
        p += "    a<<= %d d=r 5 a+=d a-= %d\n" % (rb, (1 << rb) - 1)
    p += """    d=a b=r 4 a=b a-=d c=a
    d=r 2 do a=d a> 0 if d--
      a=*c *b=a c++ b++
""" + out + """    forever endif
    a=b r=a 4
    a=r 6 b=r 3 a>>=b c=a
    a=r 7 a-=b d=a
    a=0 r=a 1
  endif endif
  do a=r 1 a== 3 if a=d a> 1 if
    a=c a&= 1 a== 1 if
      a=c a>>= 1 c=a
      b=r 2 a&= 1 a+=b a+=b r=a 2
      a=c a>>= 1 c=a
      d-- d--
    else
      a=c a>>= 1 c=a
      d--
      a= 4 r=a 1
    endif
  forever endif endif
  a=r 1 a== 4 if a=d a> 7 if
    b=r 4 a=c *b=a
""" + out + """    b++ a=b r=a 4
    a=c a>>= 8 c=a
    a=d a-= 8 d=a
    a=r 2 a-- r=a 2 a== 0 if
      a=0 r=a 1
    endif
  endif endif
  halt
end
"""
    return p


def _pcomp_lzpre(args, doe8):   # LibZPAQ.cs:574-638
    out = "" if doe8 else " out\n"
    p = "pcomp lzpre c ;\n  a> 255 if\n"
    if doe8:
        p += "    d=b b=0 do" + _E8E9_AT_EOF
    p += """    b=0 c=0 d=0 a=0 r=a 1 r=a 2
  halt
  endif
  c=a a=d a== 0 if
    a=c a>>= 6 a++ d=a
    a== 1 if
      a+=c r=a 1 a=0 r=a 2
    else
      d++ a=c a&= 63 a+= $3 r=a 1 a=0 r=a 2
    endif
  else
    a== 1 if
      a=c *b=a b++
""" + out + """      a=r 1 a-- a== 0 if d=0 endif r=a 1
    else
      a> 2 if
        a=r 2 a<<= 8 a|=c r=a 2 d--
      else
        a=r 2 a<<= 8 a|=c c=a a=b a-=c a-- c=a
        d=r 1
        do
          a=*c *b=a c++ b++
""" + out + """        d-- a=d a> 0 while
      endif
    endif
  endif
  halt
end
"""
    return p


def _pcomp_bwtrle(args, doe8):   # LibZPAQ.cs:641-795
    p = """pcomp bwtrle c ;
  a> 255 ifnot
    *b=a b++
  elsel
    b-- a=*b
    b-- a<<= 8 a+=*b
    b-- a<<= 8 a+=*b
    b-- a<<= 8 a+=*b c=a r=a 1
    a=b r=a 2
    do
      a=b a> 0 if
        b-- a=*b a++ a&= 255 d=a d! *d++
      forever
    endif
    d=0 d! *d= 1 a=0
    do
      a+=*d *d=a d--
    d<>a a! a> 255 a! d<>a until
    b=0 do
      a=c a>b if
        d=*b d! *d++ d=*d d-- *d=b
      b++ forever
    endif
    b=c b++ c=r 2 do
      a=c a>b if
        d=*b d! *d++ d=*d d-- *d=b
      b++ forever
    endif
"""
    if args[0] <= 4:
        p += """    b=0 do
      a=c a>b if
        d=b a=*d a<<= 8 a+=*b *d=a
      b++ forever
    endif
    d=r 1 b=0 do
      a=d a== 0 ifnot
        a=*d a>>= 8 d=a
"""
        p += " *b=*d b++\n" if doe8 else " a=*d out\n"
        p += "      forever\n    endif\n"
        if doe8:
            p += "    d=b b=0 do" + _E8E9_AT_EOF
        p += "  endif\n  halt\nend\n"
    elif doe8:
        p += """    a=r 2 a-- r=a 2
    c=0 d=r 1 do
      a=d a== 0 ifnot
        d=*d
        b=d a=*b a<<= 24 b=a
        a=r 4 r=a 5 a>>= 8 a|=b r=a 4
        a=c a> 3 if
          a=r 5 a&= 254 a== 232 if
            a=r 4 a>>= 24 b=a a++ a&= 254 a< 2 if
              a=r 4 a-=c a+= 4 a<<= 8 a>>= 8
              b<>a a<<= 24 a+=b r=a 4
            endif
          endif
        endif
        a=c a> 3 if a=r 5 out endif c++
      forever
    endif
    b=r 4
    a=c a> 3 a=b if out endif a>>= 8 b=a
    a=c a> 2 a=b if out endif a>>= 8 b=a
    a=c a> 1 a=b if out endif a>>= 8 b=a
    a=c a> 0 a=b if out endif
  endif
  halt
end
"""
    else:
        p += """    d=r 1 do
      a=d a== 0 ifnot
        d=*d
        b=d a=*b out
      forever
    endif
  endif
  halt
end
"""
    return p


_PCOMP_E8E9 = """pcomp e8e9 d ;
  a> 255 if
    a=c a> 4 if
      c= 4
    else
      a! a+= 5 a<<= 3 d=a a=b a>>=d b=a
    endif
    do a=c a> 0 if
      a=b out a>>= 8 b=a c--
    forever endif
  else
    *b=b a<<= 24 d=a a=b a>>= 8 a+=d b=a c++
    a=c a> 4 if
      a=*b out
      a&= 254 a== 232 if
        a=b a>>= 24 a++ a&= 254 a== 0 if
          a=b a>>= 24 a<<= 24 d=a
          a=b a-=c a+= 5
          a<<= 8 a>>= 8 a|=d b=a
        endif
      endif
    endif
  endif
  halt
end
"""   # LibZPAQ.cs:798-826


def make_config(method: str):
    """LibZPAQ.cs:388-1044.  Returns (config_text, args[9])."""
    typ = method[0]
    if typ not in "xs0i":
        raise ConfigError("method must begin with x, s, 0 or i")
    args = [0] * 9
    pos = 1
    s = method + "\0"
    i = 0
    while i < 9 and (s[pos].isdigit() or s[pos] in ",."):      # LibZPAQ.cs:405-415
        if s[pos].isdigit():
            args[i] = args[i] * 10 + ord(s[pos]) - 48
        else:
            i += 1
            if i < 9:
                args[i] = 0
        pos += 1
    if typ == "0":
        return "comp 0 0 0 0 0 hcomp end\n", args

    level = args[1] & 3
    doe8 = 4 <= args[1] <= 7
    if level == 1:
        hdr, pcomp = "comp 9 16 0 $1+20 ", _pcomp_lazy2(args, doe8)
    elif level == 2:
        hdr, pcomp = "comp 9 16 0 $1+20 ", _pcomp_lzpre(args, doe8)
    elif level == 3:
        hdr, pcomp = "comp 9 16 $1+20 $1+20 ", _pcomp_bwtrle(args, doe8)
    else:
        hdr, pcomp = "comp 9 16 0 0 ", (_PCOMP_E8E9 if doe8 else "end\n")

    ncomp = 0
    membits = args[0] + 20
    sb = 5
    comp = ""
    hcomp = "hcomp\nc-- *c=a a+= 255 d=a *d=c\n"           # LibZPAQ.cs:842-847
    if level == 2:                                          # LibZPAQ.cs:848-864
        hcomp += ("  a=r 1 a== 0 if\n    a= %d\n  else a== 1 if\n    a=*c r=a 2\n"
                  "    a> 63 if a>>= 6 a++ a++\n    else a++ a++ endif\n  else\n    a--\n"
                  "  endif endif\n  r=a 1\n") % (111 + 57 * (1 if doe8 else 0))

    while s[pos] != "\0" and ncomp < 254:                   # LibZPAQ.cs:867-1042
        v = [ord(s[pos])]
        pos += 1
        if s[pos].isdigit():
            v.append(ord(s[pos]) - 48)
            pos += 1
            while s[pos].isdigit() or s[pos] in ",.":
                if s[pos].isdigit():
                    v[-1] = v[-1] * 10 + ord(s[pos]) - 48
                else:
                    v.append(0)
                pos += 1
        cmd = chr(v[0])

        if cmd == "c":
            while len(v) < 3:
                v.append(0)
            comp += "%d " % ncomp
            sb = 11
            sb += lg(v[2]) if v[2] < 256 else 6
            for k in range(3, len(v)):
                if v[k] < 512:
                    sb += nbits(v[k]) * 3 // 4
            if sb > membits:
                sb = membits
            if v[1] % 1000 == 0:
                comp += "icm %d\n" % (sb - 6 - v[1] // 1000)
            else:
                comp += "cm %d %d\n" % (sb - 2 - v[1] // 1000, v[1] % 1000 - 1)
            hcomp += "d= %d *d=0\n" % ncomp
            if 1 < v[2] <= 255:
                if lg(v[2]) != lg(v[2] - 1):
                    hcomp += "a=c a&= %d hashd\n" % (v[2] - 1)
                else:
                    hcomp += "a=c a%%= %d hashd\n" % v[2]
            elif 1000 <= v[2] <= 1255:
                hcomp += ("a= 255 a+= %d d=a a=*d a-=c a> 255 if a= 255 endif d= %d hashd\n"
                          % (v[2] - 1000, ncomp))
            for k in range(3, len(v)):
                if k == 3:
                    hcomp += "b=c "
                if v[k] == 255:
                    hcomp += "a=*b hashd\n"
                elif 0 < v[k] < 255:
                    hcomp += "a=*b a&= %d hashd\n" % v[k]
                elif 256 <= v[k] < 512:
                    hcomp += "a=r 1 a> 1 if\n  a=r 2 a< 64 if\n    a=*b "
                    if v[k] < 511:
                        hcomp += "a&= %d" % (v[k] - 256)
                    hcomp += (" hashd\n  else\n    a>>= 6 hashd a=r 1 hashd\n  endif\nelse\n"
                              "  a= 255 hashd a=r 2 hashd\nendif\n")
                elif v[k] >= 1256:
                    hcomp += "a= %d a<<= 8 a+= %d a+=b b=a\n" % (((v[k] - 1000) >> 8) & 255, (v[k] - 1000) & 255)
                elif v[k] > 1000:
                    hcomp += "a= %d a+=b b=a\n" % (v[k] - 1000)
                if v[k] < 512 and k < len(v) - 1:
                    hcomp += "b++ "
            ncomp += 1

        if cmd in "mts" and ncomp > (1 if cmd == "t" else 0):
            if len(v) <= 1:
                v.append(8)
            if len(v) <= 2:
                v.append(24 + 8 * (1 if cmd == "s" else 0))
            if cmd == "s" and len(v) <= 3:
                v.append(255)
            comp += "%d" % ncomp
            sb = 5 + v[1] * 3 // 4
            if cmd == "m":
                comp += " mix %d 0 %d %d 255\n" % (v[1], ncomp, v[2])
            elif cmd == "t":
                comp += " mix2 %d %d %d %d 255\n" % (v[1], ncomp - 1, ncomp - 2, v[2])
            else:
                comp += " sse %d %d %d %d\n" % (v[1], ncomp - 1, v[2], v[3])
            if v[1] > 8:
                hcomp += "d= %d *d=0 b=c a=0\n" % ncomp
                while v[1] >= 16:
                    hcomp += "a<<= 8 a+=*b"
                    if v[1] > 16:
                        hcomp += " b++"
                    hcomp += "\n"
                    v[1] -= 8
                if v[1] > 8:
                    hcomp += "a<<= 8 a+=*b a>>= %d\n" % (16 - v[1])
                hcomp += "a<<= 8 *d=a\n"
            ncomp += 1

        if cmd == "i" and ncomp > 0:
            hcomp += "d= %d b=c a=*d d++\n" % (ncomp - 1)
            k = 1
            while k < len(v) and ncomp < 254:
                for j in range(v[k] % 10):
                    hcomp += "hash "
                    if k < len(v) - 1 or j < v[k] % 10 - 1:
                        hcomp += "b++ "
                    sb += 6
                hcomp += "*d=a"
                if k < len(v) - 1:
                    hcomp += " d++"
                hcomp += "\n"
                if sb > membits:
                    sb = membits
                comp += "%d isse %d %d\n" % (ncomp, sb - 6 - v[k] // 10, ncomp - 1)
                ncomp += 1
                k += 1

        if cmd == "a":
            if len(v) <= 1:
                v.append(24)
            while len(v) < 4:
                v.append(0)
            comp += "%d match %d %d\n" % (ncomp, membits - v[3] - 2, membits - v[2])
            hcomp += "d= %d a=*d a*= %d a+=*c a++ *d=a\n" % (ncomp, v[1])
            sb = 5 + (membits - v[2]) * 3 // 4
            ncomp += 1

        if cmd == "w":
            defaults = [None, 1, 65, 26, 223, 20, 0]
            for k in range(1, 7):
                if len(v) <= k:
                    v.append(defaults[k])
            comp += "%d icm %d\n" % (ncomp, membits - 6 - v[6])
            for k in range(1, v[1]):
                comp += "%d isse %d %d\n" % (ncomp + k, membits - 6 - v[6], ncomp + k - 1)
            hcomp += "a=*c a&= %d a-= %d a&= 255 a< %d if\n" % (v[4], v[2], v[3])
            for k in range(v[1]):
                hcomp += ("  d= %d" % ncomp) if k == 0 else "  d++"
                hcomp += " a=*d a*= %d a+=*c a++ *d=a\n" % v[5]
            hcomp += "else\n"
            for k in range(v[1] - 1, 0, -1):
                hcomp += "  d= %d a=*d d++ *d=a\n" % (ncomp + k - 1)
            hcomp += "  d= %d *d=0\nendif\n" % ncomp
            ncomp += v[1] - 1
            sb = membits - v[6]
            ncomp += 1

    return hdr + "%d\n" % ncomp + comp + hcomp + "halt\n" + pcomp, args


def plan_block(method: str, data: bytes):
    """What compressBlock (LibZPAQ.cs:117-300) derives before touching the codec:
    returns dict(method=expanded, hdr=..., pcomp=..., args=[9], comment=str)."""
    m = expand_method(method, data)
    text, args = make_config(m)
    hdr, pcomp, cmd = compile_config(text, args)
    return {"method": m, "config": text, "hdr": hdr, "pcomp": pcomp, "args": args,
            "pcomp_cmd": cmd, "comment": str(len(data))}
