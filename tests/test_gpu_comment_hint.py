"""GPU: the decimal number a segment comment starts with is only a hint for the output slot (the reference ignores comments on
decode, LibZPAQ.cs:65-79): archives whose comments say something else still decode."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_comments_that_are_not_the_size(gpu_ctx, oracle):
    from tools import synth
    data = synth.blocks("mixed", 2100, 1, 50000).tobytes()
    arcs = [oracle.compress_block_level(data[:30000], 1, comment="2024 backup"),
            oracle.compress_block_level(data[30000:], 1, comment="1"),
            oracle.compress_block_level(data[:20000], 2, comment="99999999999 huge"),
            oracle.compress_block_level(data[:100], 1, comment="")]
    offs = np.concatenate([[0], np.cumsum([len(a) for a in arcs])]).astype(np.uint64)
    out, ooff, sha, bst = gpu_ctx.decompress_blocks(b"".join(arcs), offs, out=np.empty(1 << 20, dtype=np.uint8))
    assert out.tobytes() == data + data[:20000] + data[:100] and set(sha.tolist()) == {1} and not bst.any()
