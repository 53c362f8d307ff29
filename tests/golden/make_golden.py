#!/usr/bin/env python3
"""Extract the known-answer material the reference holds for the codec path and
store it as small JSON fixtures (run in the build container only; the GPU box
has no /root/reference).

Sources (all under /root/reference/ZPAQSharp):
  StateTable.cs:21-149    sns[1024]     next-state table literal
  Predictor.cs:1359-1392  sdt2k[256]    = 2048/i
  Predictor.cs:1395-1524  sdt[1024]     = (1<<17)/(i*2+3)*2
  Predictor.cs:1527-1697  ssquasht[1344]
  Predictor.cs:1700-1790  stdt[712]     stretch run lengths
  Predictor.cs:76-77      stsum / sqsum table checksums
  Compressor.cs:48-74     min/mid/max model bytecode
  Component.cs:27-43      compsize[]
  Compressor.cs:27-43     13-byte locator tag; Decompresser.cs:34,43 rolling-hash constants
"""
import json, re, sys, os

REF = "/root/reference/ZPAQSharp"

def literal(path, start_pat, cast_char=False):
    src = open(os.path.join(REF, path), encoding="utf-8-sig").read()
    i = src.index(start_pat)
    i = src.index("{", i)
    j = src.index("}", i)
    body = src[i + 1:j]
    body = re.sub(r"//[^\n]*", "", body)
    out = []
    for tok in body.split(","):
        tok = tok.strip()
        if not tok:
            continue
        m = re.fullmatch(r"\(char\)\s*(-?\d+)", tok)
        if m:
            out.append(int(m.group(1)) & 255)
        else:
            out.append(int(tok, 0) & (255 if cast_char else 0xFFFFFFFF))
    return out

def main():
    g = {}
    g["sns"] = literal("StateTable.cs", "sns = new byte[1024]")
    g["sdt2k"] = literal("Predictor.cs", "sdt2k = new int[256]")
    g["sdt"] = literal("Predictor.cs", "sdt = new int[1024]")
    g["ssquasht"] = literal("Predictor.cs", "ssquasht = new ushort[1344]")
    g["stdt"] = literal("Predictor.cs", "stdt = new byte[712]")
    g["stsum"] = 3887533746
    g["sqsum"] = 2278286169
    models = literal("Compressor.cs", "models[]", cast_char=True)
    ms, p = [], 0
    while models[p] + 256 * models[p + 1] > 0:
        n = models[p] + 256 * models[p + 1] + 2
        ms.append(models[p:p + n]); p += n
    g["models"] = ms
    g["compsize"] = literal("Component.cs", "compsize = new int[256]")[:16]
    g["tag"] = [0x37,0x6b,0x53,0x74,0xa0,0x31,0x83,0xd3,0x8c,0xb2,0x28,0xb0,0xd3]
    g["findblock_seed"] = [0x3D49B113, 0x29EB7F93, 0x2614BE13, 0x3828EB13]
    g["findblock_hit"] = [0xB16B88F1, 0xFF5376F1, 0x72AC5BF1, 0x2F909AF1]
    for k in ("sns","sdt2k","sdt","ssquasht","stdt"):
        print(k, len(g[k]), file=sys.stderr)
    print("models", [len(m) for m in ms], file=sys.stderr)
    json.dump(g, open(os.path.join(os.path.dirname(__file__), "reference_kat.json"), "w"),
              separators=(",", ":"))

if __name__ == "__main__":
    main()
