#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native ZPAQ block codec.

Metric (BASELINE.json): compress & decompress MB/s (MB = 10^6 input bytes), method-2 (mid.cfg =
Compressor.startBlock(2)), 1 MB blocks (1,044,480 B), byte-exact, 1/2/4/8 B200.

A "step" is one pass of the compress hot path over one batch of synthetic mixed text/binary blocks
(the batch is as many blocks as fit resident on one GPU: one warp per block, ~106 MiB of model
state per block).  Blocks are independent, so N GPUs = N ranks each coding its own batch (weak
scaling, no collective on the data path; NCCL is used for the barrier and the max-over-ranks only).

  value     compress MB/s, whole job, inputs and outputs resident in HBM (kernel path only)
  e2e       the same through the public C ABI with HOST buffers: pinned H2D of every block and
            D2H of every archive inside the timed region
  roofline  HBM roofline of the coding kernel from the algorithmic bytes of SURVEY.md 8d
  cpu_baseline / --impl reference
            the CPU oracle (a C++ restatement of the reference; ZPAQSharp itself cannot be built,
            SURVEY.md 8c) timed on the box's host cores on a bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOCK = (0x100000 << 0) - 4096          # 1,044,480 B  (LibZPAQ.cs:94)
TOTAL_BLOCKS = 8192                     # the 8 GB stream of BASELINE.json configs[1]
LEVEL = 2                               # mid.cfg
ALGO_BYTES_PER_INPUT_BYTE = 950.0       # A(C2a), state read+write at the reference's granularity
METRIC = "compress MB/s, method-2 (mid.cfg) 1 MB blocks, byte-exact"


def shard_blocks(total: int, rank: int, world: int):
    """Contiguous block range of `rank` (blocks are independent: SURVEY.md 8e)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_max_time(t: float, device: str) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return t
    x = torch.tensor([t], dtype=torch.float64, device=device)
    dist.all_reduce(x, op=dist.ReduceOp.MAX)
    return float(x.item())


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _reference_block_coder():
    """The reference's own text for the hot loop, when oracle/_ref holds it: Predictor.init / predict0 / update0 / find,
    ZPAQL.execute and Encoder.encode compiled by oracle/build_ref.py (prebuilt; nothing is read from /root/reference at
    run time).  Returns code(block) -> coded bytes of the block, or None."""
    import ctypes as C
    import hashlib
    try:
        from oracle import build_ref, frontend
        path = build_ref.build_predictor()
        if not path or not os.path.exists(path):
            return None
        L = C.CDLL(path)
        kat = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kat.json")))
        tabs = [np.asarray(kat["sdt2k"], dtype=np.int32), np.asarray(kat["sdt"], dtype=np.int32), np.asarray(kat["ssquasht"], dtype=np.uint16),
                np.asarray(kat["stdt"], dtype=np.int32), np.asarray(kat["sns"], dtype=np.uint8)]
        L.ref_predictor_tables.argtypes = [C.c_void_p] * 5
        L.ref_predictor_tables(*[t.ctypes.data for t in tabs])
        L.ref_code_block.argtypes = [C.c_char_p, C.c_ulonglong, C.c_char_p, C.c_ulonglong, C.c_char_p, C.c_ulonglong, C.c_void_p, C.c_ulonglong]
        L.ref_code_block.restype = C.c_longlong
        hdr = bytes(frontend.builtin_model(LEVEL)[0])
    except Exception:
        return None

    def code(b):
        hashlib.sha1(b).digest()                    # compressBlock hashes the block first (LibZPAQ.cs:143-155; the SHA1 class itself is missing)
        cap = len(b) + len(b) // 4 + 4096
        out = C.create_string_buffer(cap)
        n = L.ref_code_block(hdr, len(hdr), b"\x00", 1, b, len(b), out, cap)      # ctypes releases the GIL: one block per thread
        if n < 0:
            raise RuntimeError("reference coder failed")
        return out.raw[:n]
    return code


def _reference_block_codec():
    """The reference's own text for the WHOLE block path, when oracle/_ref holds it: Compressor.writeTag / startBlock(level) /
    startSegment / postProcess / compress / endSegment / endBlock over Encoder, Predictor and ZPAQL text, and the way back
    through Decompresser / Decoder / PostProcessor (oracle/build_ref.py; prebuilt, nothing is read from /root/reference at
    run time).  Returns (compress(block) -> archive block, decompress(archive, n) -> bytes) or None."""
    import ctypes as C
    import hashlib
    try:
        from oracle import build_ref
        pc, pd = build_ref.build_compressor(), build_ref.build_decompresser()
        if not pc or not pd or not os.path.exists(pc) or not os.path.exists(pd):
            return None
        kat = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kat.json")))
        tabs = [np.asarray(kat["sdt2k"], dtype=np.int32), np.asarray(kat["sdt"], dtype=np.int32), np.asarray(kat["ssquasht"], dtype=np.uint16),
                np.asarray(kat["stdt"], dtype=np.int32), np.asarray(kat["sns"], dtype=np.uint8)]
        Lc, Ld = C.CDLL(pc), C.CDLL(pd)
        for L in (Lc, Ld):
            L.ref_predictor_tables.argtypes = [C.c_void_p] * 5
            L.ref_predictor_tables(*[t.ctypes.data for t in tabs])
        Lc.ref_compress_block.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_ulonglong,
                                          C.c_char_p, C.c_int, C.c_void_p, C.c_ulonglong]
        Lc.ref_compress_block.restype = C.c_longlong
        Ld.ref_decompress.argtypes = [C.c_char_p, C.c_ulonglong, C.c_void_p, C.c_ulonglong, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        Ld.ref_decompress.restype = C.c_longlong
    except Exception:
        return None

    def comp(b):
        sha = hashlib.sha1(b).digest()              # compressBlock hashes the block first (LibZPAQ.cs:143-155; the SHA1 class itself is missing)
        cap = len(b) + len(b) // 4 + 4096
        out = C.create_string_buffer(cap)
        n = Lc.ref_compress_block(LEVEL, None, None, 0, None, str(len(b)).encode(), b, len(b), sha, 1, out, cap)   # ctypes releases the GIL
        if n < 0 or n > cap:
            raise RuntimeError("reference Compressor text failed")
        return out.raw[:n]

    def decomp(a, n):
        out = C.create_string_buffer(n + 64)
        marks = C.create_string_buffer(21 * 4)
        nseg = C.c_int(0)
        m = Ld.ref_decompress(a, len(a), out, n + 64, marks, 4, C.byref(nseg))
        if m < 0:
            raise RuntimeError("reference Decompresser text failed")
        return out.raw[:m]
    return comp, decomp


def cpu_baseline(seconds_budget: float = 20.0, decompress: bool = True):
    """Time the CPU path (one block per host thread) on a bounded sample of the workload: the reference's own text when
    oracle/_ref holds it (kind "reference": the whole Compressor / Decompresser path, else only the per-bit hot loop), else
    the oracle port (kind "port")."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle as po
    from tools import synth
    po.build()
    cores = os.cpu_count() or 1
    data = synth.blocks("mixed", 0, cores, BLOCK)
    blocks = [data[i * BLOCK:(i + 1) * BLOCK].tobytes() for i in range(cores)]
    codec = _reference_block_codec()
    ref_code = None if codec else _reference_block_coder()

    def comp(b):
        return po.compress_block_level(b, LEVEL)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        arcs = list(ex.map(comp, blocks))
    t_port = time.perf_counter() - t0
    out = {"value": cores * BLOCK / 1e6 / t_port, "unit": "MB/s", "cores": cores, "kind": "port",
           "sample": "%d blocks of %d B (one per host thread), mid.cfg, oracle C++ -O2" % (cores, BLOCK),
           "seconds": t_port}
    if codec is not None:
        t0 = time.perf_counter()
        with ThreadPoolExecutor(cores) as ex:
            rarcs = list(ex.map(codec[0], blocks))
        t_ref = time.perf_counter() - t0
        assert rarcs == arcs, "reference text and oracle disagree"          # whole archive blocks, byte for byte
        out.update({"value": cores * BLOCK / 1e6 / t_ref, "kind": "reference", "seconds": t_ref, "port_value": cores * BLOCK / 1e6 / t_port,
                    "sample": "%d blocks of %d B (one per host thread), mid.cfg; the reference's own Compressor.startBlock(2) .. endBlock "
                              "path (Compressor, Encoder, Predictor.init/predict0/update0/find, ZPAQL text compiled -O2 from "
                              "/root/reference by oracle/build_ref.py, + SHA-1 of the block); archive blocks checked byte for byte "
                              "against the oracle's" % (cores, BLOCK)})
    elif ref_code is not None:
        t0 = time.perf_counter()
        with ThreadPoolExecutor(cores) as ex:
            coded = list(ex.map(ref_code, blocks))
        t_ref = time.perf_counter() - t0
        # the reference text must have produced the coded payload of the oracle's archive block (13-byte tag, zPQ header,
        # segment header in front; 00 00 00 00 FD sha1[20] FF behind)
        assert all(a[-26 - len(c):-26] == c for a, c in zip(arcs, coded)), "reference text and oracle disagree"
        out.update({"value": cores * BLOCK / 1e6 / t_ref, "kind": "reference", "seconds": t_ref, "port_value": cores * BLOCK / 1e6 / t_port,
                    "sample": "%d blocks of %d B (one per host thread), mid.cfg; the reference's own Predictor.init/predict0/update0/find, "
                              "ZPAQL.execute and Encoder.encode text compiled -O2 from /root/reference by oracle/build_ref.py "
                              "(+ SHA-1 of the block); coded bytes checked against the oracle's archives" % (cores, BLOCK)})
    if decompress and out["seconds"] < seconds_budget:
        dec = (lambda a: codec[1](a, BLOCK)) if codec else (lambda a: po.decompress(a, cap=BLOCK + 64)[0])
        t0 = time.perf_counter()
        with ThreadPoolExecutor(cores) as ex:
            back = list(ex.map(dec, arcs))
        t_d = time.perf_counter() - t0
        assert all(b == s for b, s in zip(back, blocks))
        out["decompress_value"] = cores * BLOCK / 1e6 / t_d        # the reference's Decompresser text when present, else the oracle port
        out["decompress_kind"] = "reference" if codec else "port"
    t0 = time.perf_counter()
    (codec[0] if codec else (ref_code or comp))(blocks[0])
    out["single_core_value"] = BLOCK / 1e6 / (time.perf_counter() - t0)
    return out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores, rank 0 only: the reference's own
    text for the per-bit hot loop compiled into oracle/_ref (kind "reference"), or the oracle port when that is missing."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workload_config(None, args.gpus)
    vals = []
    for _ in range(args.warmup):
        cpu_baseline(decompress=False)
    t_steps = []
    base = None
    for _ in range(max(1, args.steps)):
        t0 = time.perf_counter()
        base = cpu_baseline(decompress=False)
        t_steps.append(time.perf_counter() - t0)
        vals.append(base["value"])
    v = sum(vals) / len(vals)
    base["value"] = v
    _emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "MB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(t_steps) / len(t_steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": cfg, "cpu_baseline": base,
        "e2e": {"value": v, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "kind reference: Compressor/Encoder/Predictor/ZPAQL text of the reference compiled from /root/reference (oracle/build_ref.py); "
                "kind port: the C++ oracle restating it (ZPAQSharp as a whole is not buildable)",
    })


def workload_config(batch_blocks, n_gpus):
    return {"workload": "BASELINE configs[1]: mid.cfg (Compressor.startBlock(2)), 1,044,480-byte blocks of the "
                        "8192-block (8 GB) synthetic mixed text/binary stream",
            "block_bytes": BLOCK, "stream_blocks": TOTAL_BLOCKS, "batch_blocks_per_gpu": batch_blocks,
            "parallelism": "blocks x%d GPUs, no collective" % n_gpus,
            "l2_policy": "inputs larger than L2 (batch >= 1 GB per GPU, per-block state ~106 MiB)",
            "sha1": True}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-blocks", type=int, default=0, help="blocks per GPU per step (0 = one resident wave)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decompress", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from tools import synth
    from zpaqsharp_b200 import libzpaq as z

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    ctx = z.Context([local])
    stream = torch.cuda.Stream(device=dev)       # the library launches on this stream; events are recorded on it
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    hdr = z.builtin_model(LEVEL)
    state = z.device_state_bytes(hdr)

    # batch = one resident wave: what fits next to the I/O buffers
    free, total_mem = torch.cuda.mem_get_info()
    if args.batch_blocks:
        B = args.batch_blocks
    else:
        io_per_block = BLOCK * 5                      # input + slots + out (+ slack), device resident path
        B = int((free - (3 << 30)) // (state + io_per_block))
        B = max(1, min(B, 148 * 16, TOTAL_BLOCKS // max(world, 1)))
    first = (rank * B) % TOTAL_BLOCKS

    host_in = torch.empty(B * BLOCK, dtype=torch.uint8).pin_memory()
    synth.fill(host_in.numpy(), "mixed", first, B, BLOCK)
    offs = np.arange(0, (B + 1) * BLOCK, BLOCK, dtype=np.uint64)
    out_cap = B * (BLOCK + BLOCK // 4 + 8192)
    host_out = torch.empty(out_cap, dtype=torch.uint8).pin_memory()
    d_in = host_in.to(dev, non_blocking=False)
    d_out = torch.empty(out_cap, dtype=torch.uint8, device=dev)

    def step_device():
        return ctx.compress_blocks_model_dev(d_in.data_ptr(), offs, hdr, d_out.data_ptr(), out_cap)

    def step_host():
        return ctx.compress_blocks_level(host_in.numpy(), offs, LEVEL, out=host_out.numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for _ in range(args.warmup):
        ooff = step_device()
    st = ctx.stats()
    resident = st.resident_blocks

    # ---- timed: device-resident path ----
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    codec_ms = []
    launches = 0
    e0.record(stream)
    for _ in range(args.steps):
        ooff = step_device()
        s = ctx.stats()
        codec_ms.append(s.codec_kernel_ms)
        launches += s.launches
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    t_dev = reduce_max_time(e0.elapsed_time(e1) / 1e3, dev)
    in_bytes = B * BLOCK
    value = world * args.steps * in_bytes / 1e6 / t_dev
    ratio = float(ooff[B]) / in_bytes

    # ---- timed: end to end through the host-buffer ABI ----
    archive, hoff = step_host()            # warm the host path (pinned buffers, staging)
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(args.steps):
        archive, hoff = step_host()
        s = ctx.stats()
        h2d += s.h2d_bytes
        d2h += s.d2h_bytes
        launches += s.launches
    torch.cuda.synchronize()
    t_e2e = reduce_max_time(time.perf_counter() - t0, dev)
    e2e_value = world * args.steps * in_bytes / 1e6 / t_e2e

    # ---- decompression of the same batch (device decode + host framing parse) ----
    dec = None
    if not args.no_decompress:
        arc_np = archive
        back = np.empty(in_bytes + 64, dtype=np.uint8)
        out, o2, sha, bst = ctx.decompress_blocks(arc_np, hoff, out=back)        # warm-up + check
        ok = bool(np.array_equal(out, host_in.numpy())) and set(sha.tolist()) == {1}
        barrier()
        t0 = time.perf_counter()
        out, o2, sha, bst = ctx.decompress_blocks(arc_np, hoff, out=back)
        torch.cuda.synchronize()
        t_d = reduce_max_time(time.perf_counter() - t0, dev)
        s = ctx.stats()
        launches += s.launches
        dec = {"e2e_value": world * in_bytes / 1e6 / t_d, "unit": "MB/s", "codec_kernel_ms": s.codec_kernel_ms,
               "round_trip_identical": ok, "sha1_verified_blocks": int((sha == 1).sum())}

    # ---- roofline of the coding kernel ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    k_ms = sum(codec_ms) / len(codec_ms)
    achieved = ALGO_BYTES_PER_INPUT_BYTE * in_bytes / (k_ms / 1e3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("dram_bytes_per_input_byte")
        if traffic is not None:
            traffic = traffic * in_bytes
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": st.kernel.decode(errors="replace"), "kernel_ms": k_ms,
                "algorithmic_bytes_per_input_byte": ALGO_BYTES_PER_INPUT_BYTE,
                "peak_source": "MEASURED_PEAKS.json (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "traffic_source": "profiles/traffic.json: DRAM bytes per input byte from an ncu --set full capture of the same kernel, scaled to this launch" if traffic is not None else None,
                "note": "latency-bound: a dependent chain per role warp and block, 11 blocks per SM is all HBM holds; see DESIGN.md 2.3"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "MB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(B, world),
            "e2e": {"value": e2e_value, "unit": "MB/s", "h2d_bytes_per_step": h2d // max(args.steps, 1),
                    "d2h_bytes_per_step": d2h // max(args.steps, 1)},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "compression_ratio": ratio, "resident_blocks_per_gpu": int(resident),
            "state_bytes_per_block": int(st.state_bytes_per_block), "decompress": dec,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _emit(obj):
    """Write the one JSON line to the REAL stdout (libraries such as NCCL print banners to fd 1: it is parked on
    stderr for the whole run, see the bottom of this file)."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


if __name__ == "__main__":
    # stdout carries exactly one JSON line: everything else that is written to fd 1 (e.g. "NCCL version ...") goes to stderr
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    main()
