"""CPU model (numpy) of the GPU formulation -- per-byte run tables + per-column rank walk -- checked against the
oracle's restated pbc_dec.  It guards the arithmetic the kernels rely on (pbwt_kernels.cu: parse_runs / lookup_runs):
   start[i]  = ranks before byte i's run,  delta[i] = bit ? (m-n1) - zeros_before : -ones_before,
   j = last byte with start[j] <= rank (uniform binary search),  rank' = rank + delta[j],  bit = rank' >= m-n1.
"""
import numpy as np

from cohorts import edge_rows, haplo_matrix, random_matrix


def rle_len(c):
    v = c >> 1
    return (v & 15) << (4 * (v >> 4))


def tables(rle, m):
    c = np.frombuffer(rle, dtype=np.uint8).astype(np.int64)
    L = rle_len(c)
    b = c & 1
    start = np.cumsum(L) - L
    ones_before = np.cumsum(L * b) - L * b
    n1 = int((L * b).sum())
    zeros_before = start - ones_before
    delta = np.where(b == 1, (m - n1) - zeros_before, -ones_before)
    return start, delta, n1


def uniform_search(start, r):
    lo = np.zeros_like(r)
    n = len(start)
    while n > 1:
        half = n >> 1
        lo = lo + np.where(start[lo + half] <= r, half, 0)
        n -= half
    return lo


def walk_file(pbf_bytes, orc):
    p = orc.Pbf(pbf_bytes)
    m, n, shift = p.m, p.n, p.shift
    buf = np.frombuffer(pbf_bytes, dtype=np.uint8)
    pos = 16
    rank = [None, None]
    out = np.zeros((n, m), np.uint8)
    for k in range(n):
        if buf[pos] == ord('S'):
            for g in range(2):
                S = np.frombuffer(pbf_bytes, dtype=np.int32, count=m, offset=pos + 1 + 4 * m * g)
                rank[g] = np.empty(m, np.int64)
                rank[g][S] = np.arange(m)
            pos += 1 + 8 * m
        assert buf[pos] == ord('B')
        pos += 1
        for g in range(2):
            l = int(np.frombuffer(pbf_bytes, dtype=np.int32, count=1, offset=pos)[0])
            rle = pbf_bytes[pos + 4:pos + 4 + l]
            pos += 4 + l
            start, delta, n1 = tables(rle, m)
            if n1 == 0 or n1 == m:
                bit = np.full(m, int(n1 == m and n1 > 0), np.uint8)
            else:
                j = uniform_search(start, rank[g])
                rank[g] = rank[g] + delta[j]
                bit = (rank[g] >= m - n1).astype(np.uint8)
            out[k] |= bit << g
    p.close()
    return out


def test_model_matches_oracle(oracle):
    for mat, shift in ((haplo_matrix(300, 97, 3), 5), (random_matrix(150, 40, 4), 3), (edge_rows(5000), 4)):
        pbf = oracle.encode_pbf(mat, shift=shift)
        assert (walk_file(pbf, oracle) == mat).all()


def test_model_zero_length_and_split_runs(oracle):
    # non-canonical streams the decoder must accept: zero-length bytes, same-bit neighbours
    m = 40
    # 5 zeros, zero-length 1-run (byte 1), 3 zeros, 12 ones, zero-length 0-run (digit 0 at position 1 = byte 32),
    # then 20 zeros as two bytes (16 + 4); plane 1: 32 zeros as one digit byte, 8 ones
    rows = [(bytes([5 << 1, 1, 3 << 1, 12 << 1 | 1, 32, (16 + 1) << 1, 4 << 1]), bytes([(16 + 2) << 1, 8 << 1 | 1]))]
    # same-bit neighbours that the canonical encoder would have merged; zero-length 1-run via byte 33
    rows.append((bytes([10 << 1 | 1, 10 << 1 | 1, 33, 10 << 1, 10 << 1]), bytes([(16 + 2) << 1 | 1, 33, 8 << 1])))
    rows.append((bytes([1, 7 << 1, 1, 33, 13 << 1 | 1, (16 + 1) << 1, 4 << 1 | 1]), bytes([(16 + 2) << 1, 8 << 1])))
    pbf = oracle.encode_pbf_rle(m, rows, shift=2)
    want = oracle.decode_all(pbf)
    assert (walk_file(pbf, oracle) == want).all()
