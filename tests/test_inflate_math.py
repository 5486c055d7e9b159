"""CPU checks of the arithmetic the on-device inflaters rely on (csrc/fq_inflate.cuh, csrc/fq_gzip.cu, fqgpu_api.cu),
restated in Python against zlib and RFC 1951: the kernels themselves are covered by tests/test_gpu_gzip.py and
tests/test_gpu_bgzf.py on the GPU."""
import os
import zlib

POLY = 0xEDB88320


def gf_mul(a: int, b: int) -> int:
    """a * b modulo the CRC-32 polynomial, bit 31 = x^0 (crc_mul in fqgpu_api.cu, gf_mul in fq_gzip.cu)."""
    p = 0
    for i in range(32):
        if (a >> (31 - i)) & 1:
            p ^= b
        b = (b >> 1) ^ (POLY if b & 1 else 0)
    return p


def xpow8(nbytes: int) -> int:
    r, sq = 0x80000000, 0x00800000  # 1, x^8
    while nbytes:
        if nbytes & 1:
            r = gf_mul(sq, r)
        sq = gf_mul(sq, sq)
        nbytes >>= 1
    return r


_TAB = []
for _t in range(256):
    _c = _t
    for _ in range(8):
        _c = (_c >> 1) ^ POLY if _c & 1 else _c >> 1
    _TAB.append(_c)


def raw_register(data: bytes) -> int:
    """The CRC register after `data` when it starts at zero (what one thread of gz_crc_slices_kernel computes)."""
    v = 0
    for b in data:
        v = _TAB[(v ^ b) & 0xFF] ^ (v >> 8)
    return v


def test_crc_of_slices_folded_in_order_equals_zlib():
    """raw(A ++ B) = raw(A) * x^(8|B|) + raw(B); the member's CRC = ~(ones * x^(8 n) + raw(all)): slices of a fixed size
    folded Horner-wise, a tree over runs of slices, a tail, batches joined on the host."""
    data = os.urandom(70_001)
    S = 4096
    nfull, tail = len(data) // S, len(data) % S
    xs = xpow8(S)
    horner = 0
    for k in range(nfull):
        horner = gf_mul(xs, horner) ^ raw_register(data[k * S:(k + 1) * S])
    # the same as the fold kernel does it: 1024 runs of q slices (zero slices in front), then a pairwise tree
    q = (nfull + 1023) // 1024
    pad = 1024 * q - nfull
    part = []
    for t in range(1024):
        v = 0
        for j in range(q):
            virt = t * q + j
            if virt >= pad:
                v = gf_mul(xs, v) ^ raw_register(data[(virt - pad) * S:(virt - pad + 1) * S])
        part.append(v)
    op, d = xpow8(S * q), 1
    while d < 1024:
        for t in range(0, 1024, 2 * d):
            part[t] = gf_mul(op, part[t]) ^ part[t + d]
        op = gf_mul(op, op)
        d *= 2
    assert part[0] == horner
    batch = gf_mul(xpow8(tail), horner) ^ raw_register(data[nfull * S:])
    reg = gf_mul(xpow8(len(data)), 0xFFFFFFFF) ^ batch
    assert reg ^ 0xFFFFFFFF == zlib.crc32(data)
    more = os.urandom(5_000)  # a second batch behind it
    reg2 = gf_mul(xpow8(len(more)), reg) ^ raw_register(more)
    assert reg2 ^ 0xFFFFFFFF == zlib.crc32(data + more)


def test_length_and_distance_codes_computed_instead_of_looked_up():
    len_base = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
    len_extra = [0] * 8 + [1] * 4 + [2] * 4 + [3] * 4 + [4] * 4 + [5] * 4 + [0]
    dist_base = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
                 8193, 12289, 16385, 24577]
    dist_extra = [0, 0, 0, 0] + [e for e in range(1, 14) for _ in (0, 1)]
    for s in range(29):  # len_code() of fq_inflate.cuh
        extra = 0 if s < 8 else (s - 4) >> 2
        base = 3 + s if s < 8 else ((4 + (s & 3)) << extra) + 3
        if s == 28:
            base, extra = 258, 0
        assert (base, extra) == (len_base[s], len_extra[s])
    for d in range(30):  # dist_code()
        extra = 0 if d < 4 else (d - 2) >> 1
        base = 1 + d if d < 4 else ((2 + (d & 1)) << extra) + 1
        assert (base, extra) == (dist_base[d], dist_extra[d])


def test_kraft_terms_three_lengths_per_lookup():
    """The block-start search sums 2^(7 - len) over the code-length code three 3-bit lengths at a time (kraft3[512]);
    a complete code sums to 128."""
    kraft3 = [((128 >> (x & 7)) & 127) + ((128 >> ((x >> 3) & 7)) & 127) + ((128 >> (x >> 6)) & 127) for x in range(512)]
    lens = [3, 3, 3, 3, 3, 3, 4, 4, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 3]  # 7 codes of 3 bits and 2 of 4: complete
    assert sum(2 ** (7 - n) for n in lens if n) == 128
    pre = 0
    for k, n in enumerate(lens):
        pre |= n << (3 * k)
    total = 0
    for _ in range(7):
        total += kraft3[pre & 511]
        pre >>= 9
    assert total == 128
