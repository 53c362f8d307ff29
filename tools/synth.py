"""Deterministic synthetic inputs -- thin ctypes wrapper over tools/synth.c.
Shared by tests/ and bench.py; not part of the product and not part of the oracle."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libzpqsynth.so")
_lib = None

BLOCK_1MB = (0x100000 << 0) - 4096   # 1,044,480  (LibZPAQ.cs:94, arg0 = 0)
BLOCK_4MB = (0x100000 << 2) - 4096   # 4,190,208
BLOCK_16MB = (0x100000 << 4) - 4096  # 16,773,120


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "synth.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src])
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.synth_fill.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
    return _lib


def fill(out: np.ndarray, kind: str, first_block: int, count: int, block_bytes: int, threads: int | None = None):
    """Write blocks [first_block, first_block+count) of `kind` ("text" | "mixed") into `out`."""
    L = _load()
    mixed = 1 if kind == "mixed" else 0
    assert out.dtype == np.uint8 and out.size >= count * block_bytes
    base = out.ctypes.data
    threads = threads or min(32, os.cpu_count() or 1)
    step = max(1, (count + threads * 4 - 1) // (threads * 4))

    def work(lo):
        n = min(step, count - lo)
        L.synth_fill(mixed, first_block + lo, n, block_bytes, C.c_void_p(base + lo * block_bytes))

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(0, count, step)))
    return out


def blocks(kind: str, first_block: int, count: int, block_bytes: int) -> np.ndarray:
    out = np.empty(count * block_bytes, dtype=np.uint8)
    return fill(out, kind, first_block, count, block_bytes)
