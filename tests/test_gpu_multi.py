"""GPU, two or more devices: ONE zpq_ctx over several GPUs (the north-star's host call).  The library partitions the blocks of
a batch over the devices -- compress by input bytes, decompress by archive bytes --, runs one host thread and one stream per
device and reassembles archives / restored bytes in block order (replaces the loops of LibZPAQ.cs:100-107 and :65-79).  Skipped
on a single-GPU box; `scripts/gpu_r02_multi.sh` runs it under `gpurun --gpus 2`."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ndev():
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for line in out.splitlines() if line.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.skipif(_ndev() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("what", [("level", 2), ("level", 1), ("method", "x0,0c256,0,255,255"), ("method", "x0,2,12,0,7,21,1c0,0,511i2m"),
                                  ("method", "x0,3ci1"), ("method", "x0,1,4,0,7,21,1")])
def test_context_on_all_devices_matches_one_device(zlib_, oracle, what):
    from tools import synth
    nd = _ndev()
    nblk, size = 37, 30000                                   # ragged: does not divide by the device count
    data = synth.blocks("mixed", 1200, nblk, size).tobytes()
    cuts = [0] + sorted(set(int(x) for x in np.linspace(1, len(data) - 1, nblk - 1))) + [len(data)]
    cuts[5] = cuts[4]                                        # an empty block
    offs = np.asarray(cuts, dtype=np.uint64)
    with zlib_.Context([0]) as one, zlib_.Context(list(range(nd))) as many:
        f = (lambda c: c.compress_blocks_level(data, offs, what[1])) if what[0] == "level" else (lambda c: c.compress_blocks(data, offs, what[1]))
        a1, o1 = f(one)
        am, om = f(many)
        assert am.tobytes() == a1.tobytes() and om.tolist() == o1.tolist()            # ordered reassembly, byte for byte
        ref0 = (oracle.compress_block_level(data[cuts[0]:cuts[1]], what[1]) if what[0] == "level"
                else oracle.compress_block(data[cuts[0]:cuts[1]], what[1]))
        assert a1[:int(o1[1])].tobytes() == ref0
        out, oo, sha, bst = many.decompress_blocks(am, om)
        assert out.tobytes() == data and oo.tolist() == cuts and set(sha.tolist()) == {1} and not bst.any()
        # a damaged block in the second device's share: isolated, the rest restored
        bad = bytearray(am.tobytes())
        k = nblk - 3
        bad[(int(om[k]) + int(om[k + 1])) // 2] ^= 0x5A
        try:
            out2, oo2, sha2, bst2 = many.decompress_blocks(bytes(bad), om)
        except zlib_.ZpaqError as e:
            assert e.code == zlib_.E_CORRUPT and e.block_status[k] != 0
            assert sum(1 for s in e.block_status if s) == 1
        else:
            # a stored (n = 0) block has no coder to notice the damage: its bytes change, the SHA-1 check on the device says so
            assert sha2[k] == 2 and [int(x) for i, x in enumerate(sha2) if i != k] == [1] * (nblk - 1)
            assert oo2.tolist()[:k + 1] == cuts[:k + 1] and out2[:cuts[k]].tobytes() == data[:cuts[k]]
