"""A line-by-line Python model of the native post-processors of zpaqsharp_b200/csrc/zpq_post.cu (lz_bytes_segment, lz_bits_segment,
un_e8e9, the list walk of k_post_bwt) -- test infrastructure: tests/test_native_post_model.py holds it to the stored PCOMP programs
run by the oracle's ZPAQL machine on thousands of valid, truncated and damaged streams, far more than a GPU test can afford, so
that a case where the native form would quietly differ from the program (instead of handing the block back: ODD) shows up on the
CPU.  The kernels themselves are held to the same programs on the device by tests/test_gpu_postproc.py."""
OK, OVERFLOW, ODD = 0, 1, 2
M32 = 0xFFFFFFFF


def un_e8e9(x: bytearray):
    """zpq_post.cu: un_e8e9 (the warp looks at 32 positions and applies the first hit: sequentially that is every hit in order)."""
    n = len(x)
    i = 0
    while i + 4 < n:
        if (x[i] & 254) == 0xE8 and ((x[i + 4] + 1) & 254) == 0:
            a = (x[i + 1] | x[i + 2] << 8 | x[i + 3] << 16) - i
            x[i + 1], x[i + 2], x[i + 3] = a & 255, (a >> 8) & 255, (a >> 16) & 255
        i += 1
    return x


def lz_copy(M: bytearray, b: int, src: int, length: int):
    for i in range(length):                      # ascending byte loop: overlap repeats the period
        M[b + i] = M[src + i]


def lz_bytes_segment(data: bytes, mlimit: int, min_match: int):
    n, pos = len(data), 0
    M = bytearray()
    while pos < n:
        c = data[pos]
        if c < 64:
            length = c + 1
            avail = min(length, n - pos - 1)
            if len(M) + avail > mlimit:
                return ODD, None
            M += data[pos + 1:pos + 1 + avail]
            pos += 1 + length
        else:
            nb = (c >> 6) + 1
            if pos + 1 + nb > n:
                break
            off = int.from_bytes(data[pos + 1:pos + 1 + nb], "big")
            length = (c & 63) + min_match
            b = len(M)
            if length == 0 or off + 1 > b or b + length > mlimit:
                return ODD, None
            M += bytes(length)
            lz_copy(M, b, b - off - 1, length)
            pos += 1 + nb
    return OK, M


def lz_bits_segment(data: bytes, mlimit: int, rb: int):
    bits = nbit = state = length = m = r5 = 0
    M = bytearray()
    for byte in data:
        bits = (bits + (byte << (nbit & 31))) & M32
        nbit = (nbit + 8) & M32
        if state == 0:
            length = 1
            if bits & 3:
                m = ((bits & 3) - 1) * 8
                bits >>= 2
                m += bits & 7
                bits >>= 3
                nbit = (nbit - 5) & M32
                state = 1
            else:
                bits >>= 2
                nbit = (nbit - 2) & M32
                state = 3
        while state == 1 and nbit > 2:
            if bits & 1:
                bits >>= 1
                length = (length + length + (bits & 1)) & M32
                bits >>= 1
                nbit = (nbit - 2) & M32
            else:
                bits >>= 1
                length = ((length << 2) + (bits & 3)) & M32
                bits >>= 2
                nbit = (nbit - 3) & M32
                state = 5 if rb else 2
        if rb and state == 5 and nbit > rb - 1:
            r5 = bits & ((1 << rb) - 1)
            bits >>= rb
            nbit = (nbit - rb) & M32
            state = 2
        if state == 2 and not (m > nbit):
            off = ((bits & ((1 << m) - 1)) + (1 << m)) & M32
            if rb:
                off = ((off << rb) + r5 - ((1 << rb) - 1)) & M32
            ptr = len(M)
            if off == 0 or off > ptr or ptr + length > mlimit:
                return ODD, None
            M += bytes(length)
            lz_copy(M, ptr, ptr - off, length)
            bits >>= m
            nbit = (nbit - m) & M32
            state = 0
        while state == 3 and nbit > 1:
            if bits & 1:
                bits >>= 1
                length = (length + length + (bits & 1)) & M32
                bits >>= 1
                nbit = (nbit - 2) & M32
            else:
                bits >>= 1
                nbit = (nbit - 1) & M32
                state = 4
        if state == 4 and nbit > 7:
            if len(M) + 1 > mlimit:
                return ODD, None
            M.append(bits & 255)
            bits >>= 8
            nbit = (nbit - 8) & M32
            length = (length - 1) & M32
            if length == 0:
                state = 0
    return OK, M


def bwt_segment(data: bytes, ph: int):
    """k_post_bwt: size, start index, C[v] = 1 + bytes below v (all `size` bytes counted), stable filing of every position but idx,
    then the walk from idx to position 0 (the sub-list split of the kernel does not change the order of what is emitted)."""
    nin = len(data)
    if nin < 5 or nin - 4 + 256 > (1 << ph):
        return ODD, None
    size = nin - 4
    idx = int.from_bytes(data[size:size + 4], "little")
    if idx >= size:
        return ODD, None
    cnt = [0] * 256
    for b in range(size):
        cnt[data[b]] += 1
    C, run = [0] * 256, 1
    for v in range(256):
        C[v] = run
        run += cnt[v]
    T = [0] * (size + 1)
    for b in range(size):
        if b == idx:
            continue
        v = data[b]
        T[C[v]] = b
        C[v] += 1
    out = bytearray()
    p = idx
    while p != 0:
        p = T[p]
        out.append(data[p])
        if len(out) > size:
            return ODD, None
    return OK, out


def restore(kind: int, e8: int, param: int, ph: int, pm: int, data: bytes):
    """One segment through the native path: (status, restored bytes or None)."""
    mlimit = 1 << pm
    if kind == 5:
        return OK, bytes(un_e8e9(bytearray(data)))
    if kind == 3:
        rc, M = lz_bytes_segment(data, mlimit, param)
    elif kind == 2:
        rc, M = lz_bits_segment(data, mlimit, param)
    elif kind == 4:
        rc, M = bwt_segment(data, ph)
    else:
        return ODD, None
    if rc != OK:
        return rc, None
    if e8:
        un_e8e9(M)
    return OK, bytes(M)


def bwt_walk_split(T, idx: int, size: int, data: bytes, max_split: int = 2048):
    """The list traversal of k_post_bwt as the kernel does it: sub-lists start at idx and at every other multiple of `stride`, each
    ends at the next multiple (or at position 0); one pass measures them, the heads are ranked along the chain from idx, a second
    pass emits.  Returns the emitted bytes."""
    stride = 64
    while (size + stride - 1) // stride > max_split:
        stride <<= 1
    nsplit = (size + stride - 1) // stride
    slen, send, soff = [0] * nsplit, [0] * nsplit, [None] * nsplit
    for j in range(nsplit):
        p, n = (j * stride if j else idx), 0
        if p != 0:
            while True:
                p = T[p]
                n += 1
                if (p & (stride - 1)) == 0 or n > size:
                    break
        slen[j] = n
        send[j] = p // stride if (p & (stride - 1)) == 0 else 0
    cur = off = 0
    for _ in range(nsplit + 1):
        if soff[cur] is not None:
            break
        soff[cur] = off
        off += slen[cur]
        e = send[cur]
        if e == 0 or e >= nsplit:
            break
        cur = e
    out = bytearray(off)
    for j in range(nsplit):
        if soff[j] is None:
            continue
        p = j * stride if j else idx
        for k in range(slen[j]):
            p = T[p]
            out[soff[j] + k] = data[p]
    return bytes(out)


def bwt_list(data: bytes, size: int, idx: int):
    cnt = [0] * 256
    for b in range(size):
        cnt[data[b]] += 1
    C, run = [0] * 256, 1
    for v in range(256):
        C[v] = run
        run += cnt[v]
    T = [0] * (size + 1)
    for b in range(size):
        if b != idx:
            T[C[data[b]]] = b
            C[data[b]] += 1
    return T


def gap_hist_chunked(data: bytes, bins: int = 4096):
    """k_gap_hist: per 4096-position chunk, every position walks back through the chunk and the 4095 bytes in front of it."""
    n = len(data)
    gap = [0] * bins
    for c0 in range(0, n, bins):
        lo = c0 - bins if c0 >= bins else 0
        hi = min(n, c0 + bins)
        buf = data[lo:hi]
        for i in range(c0, hi):
            at, v = i - lo, data[i]
            reach = min(bins - 1, i)
            k = 1
            while k <= reach and buf[at - k] != v:
                k += 1
            if k <= reach:
                gap[k] += 1
            elif 0 < i < bins:
                gap[i] += 1
    return gap
