"""CPU model (numpy) of the GPU formulation -- per-byte run tables + per-column rank walk -- checked against the
oracle's restated pbc_dec.  It guards the arithmetic the kernels rely on (pbwt_kernels.cu: parse_runs / lookup_runs):
   start[i]  = ranks before byte i's run,  delta[i] = bit ? (m-n1) - zeros_before : -ones_before,
   j = last byte with start[j] <= rank (uniform binary search),  rank' = rank + delta[j],  bit = rank' >= m-n1.
"""
import numpy as np

from cohorts import edge_rows, fake_tag_matrix, haplo_matrix, random_matrix


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


# ---------------------------------------------------------------------------------------------------------------------
# Split scan: a row's effect on a rank is a piecewise translation -- one piece per run -- and so is the
# composition of 32 rows (compose.cu); the per-group marginals push ONE BIT per rank through the same partition
# (marginal.cu), and cross whole row groups by scattering the bit vector through the group's composite map.

def run_map(rle, m):
    """Forward map of one row, one piece per run (bytes of one symbol merged, empty bytes skipped): (starts, deltas, n1)."""
    start, delta, n1 = tables(rle, m)
    c = np.frombuffer(rle, dtype=np.uint8).astype(np.int64)
    L, b = rle_len(c), c & 1
    if n1 == 0 or n1 == m:
        return np.zeros(1, np.int64), np.zeros(1, np.int64), n1          # a constant row moves nothing
    keep, prev = [], 2
    for i in range(len(c)):
        if L[i] > 0 and b[i] != prev:
            keep.append(i)
        if L[i] > 0:
            prev = b[i]
    return start[keep], delta[keep], n1


def apply_map(S, D, r):
    return r + D[np.searchsorted(S, r, side="right") - 1]


def compose(SA, DA, SB, DB, m):
    """(B after A) as pieces in A's input coordinates: every piece of A is cut at the starts of B inside its image."""
    S, D = [], []
    for p in range(len(SA)):
        s0, e0 = SA[p], (SA[p + 1] if p + 1 < len(SA) else m)
        cur, end = s0 + DA[p], e0 + DA[p]                                 # image of the piece
        k = np.searchsorted(SB, cur, side="right") - 1
        S.append(s0); D.append(DA[p] + DB[k])
        for j in range(k + 1, len(SB)):
            if SB[j] >= end:
                break
            S.append(s0 + (SB[j] - cur)); D.append(DA[p] + DB[j])
    return np.array(S, np.int64), np.array(D, np.int64)


def composite(maps, m):
    """Pairwise tree composition, like the levels of pbwt_compose_kernel."""
    level = list(maps)
    while len(level) > 1:
        level = [compose(*level[i][:2], *level[i + 1][:2], m) if i + 1 < len(level) else level[i] for i in range(0, len(level), 2)]
    return level[0]


def plane0_rows(pbf_bytes, orc):
    """(m, [plane-0 RLE of every row], [S0 of every block]) of a .pbf image."""
    p = orc.Pbf(pbf_bytes)
    m, n = p.m, p.n
    p.close()
    buf = np.frombuffer(pbf_bytes, dtype=np.uint8)
    pos, rows, snaps = 16, [], []
    for _ in range(n):
        if buf[pos] == ord('S'):
            snaps.append(np.frombuffer(pbf_bytes, dtype=np.int32, count=m, offset=pos + 1).astype(np.int64))
            pos += 1 + 8 * m
        pos += 1
        l0 = int(np.frombuffer(pbf_bytes, dtype=np.int32, count=1, offset=pos)[0])
        rows.append(pbf_bytes[pos + 4:pos + 4 + l0])
        pos += 4 + l0
        l1 = int(np.frombuffer(pbf_bytes, dtype=np.int32, count=1, offset=pos)[0])
        pos += 4 + l1
    return m, rows, snaps


def test_composite_of_a_row_group_is_the_rows_applied_in_turn(oracle):
    for mat, K in ((haplo_matrix(64, 301, 21), 32), (random_matrix(48, 97, 22), 16), (edge_rows(257)[:32], 32)):
        pbf = oracle.encode_pbf(mat, shift=13)
        m, rows, _ = plane0_rows(pbf, oracle)
        r_all = np.arange(m, dtype=np.int64)
        for g0 in range(0, len(rows) - K + 1, K):
            maps = [run_map(rows[r], m) for r in range(g0, g0 + K)]
            S, D = composite([(a, b) for a, b, _ in maps], m)
            want = r_all.copy()
            for a, b, _ in maps:
                want = apply_map(a, b, want)
            assert (apply_map(S, D, r_all) == want).all()
            assert S[0] == 0 and (np.diff(S) > 0).all()
            assert len(S) <= 1 + sum(len(a) for a, _, _ in maps)          # every run boundary has exactly one pre-image
            assert sorted(want.tolist()) == list(range(m))                # still a permutation of the ranks


def gather_partition(V, rle, m):
    """Next group vector through the row's INVERSE run table (entries sorted by where they land), and the ones that land
    behind the zeros -- marginal.cu's gather, bit by bit."""
    S, D, n1 = run_map(rle, m)
    if n1 == 0:
        return V, 0
    if n1 == m:
        return V, int(V.sum())
    dst = S + D
    order = np.argsort(dst, kind="stable")
    ts, back = dst[order], -D[order]                                      # landing start, distance back to the source
    pos = np.arange(m, dtype=np.int64)
    k = np.searchsorted(ts, pos, side="right") - 1
    Vn = V[pos + back[k]]
    return Vn, int(Vn[m - n1:].sum())


def test_group_marginals_by_bit_vector_partition(oracle):
    rng = np.random.default_rng(5)
    for mat, shift, K in ((haplo_matrix(150, 203, 23), 6, 32), (random_matrix(70, 64, 24), 5, 32)):
        pbf = oracle.encode_pbf(mat, shift=shift)
        m, rows, snaps = plane0_rows(pbf, oracle)
        in_group = rng.random(m) < 0.4
        BS = 1 << shift
        for r, rle in enumerate(rows):
            if r % BS == 0:
                V = in_group[snaps[r // BS]]                              # bit i = the column at rank i is in the group
                V_seg = V.copy()                                          # the vector in front of the current row group
            # crossing a whole row group through its composite map (the seed kernel's scatter) lands on the same vector
            if r % BS and (r % BS) % K == 0:
                g0 = r - K
                S, D = composite([run_map(rows[q], m)[:2] for q in range(g0, r)], m)
                Vs = np.zeros(m, bool)
                Vs[apply_map(S, D, np.arange(m, dtype=np.int64))] = V_seg
                assert (Vs == V).all()
                V_seg = V.copy()
            V, cnt = gather_partition(V, rle, m)
            assert cnt == int(((mat[r] & 1) == 1)[in_group].sum())        # ones of the plane-0 row inside the group


def run_table(rle, m):
    """margpiece.cu's run table of a plane-0 row: per run (RLE bytes of the same symbol merged, empty bytes skipped) its start,
    its landing place (0-runs in front in order, 1-runs behind the m - n1 zeros in order, pbwt.c:79-88) and its index in landing
    order; runs alternate, so with b0 = bit of the first run and nzr = number of 0-runs:  zi(j) = (bit_j ? nzr : 0) + (j >> 1)."""
    c = np.frombuffer(rle, dtype=np.uint8).astype(np.int64)
    L, b = rle_len(c), c & 1
    keep = L > 0
    L, b = L[keep], b[keep]
    first = np.ones(len(L), bool)
    first[1:] = b[1:] != b[:-1]
    run_id = np.cumsum(first) - 1
    R = int(run_id[-1]) + 1
    length = np.bincount(run_id, weights=L, minlength=R).astype(np.int64)
    bit = b[first]
    start = np.cumsum(length) - length
    ones_before = np.cumsum(length * bit) - length * bit
    n1 = int((length * bit).sum())
    zt = m - n1
    land = np.where(bit == 1, zt + ones_before, start - ones_before)
    nzr = int((bit == 0).sum())
    j = np.arange(R)
    zi = np.where(bit == 1, nzr, 0) + (j >> 1)
    assert sorted(zi.tolist()) == list(range(R)) and (np.diff(land[np.argsort(zi)]) > 0).all()   # landing order is a permutation, ascending
    return start, land, zi, zt, R


def compose_list(d, dl, m, U, V, S):
    """One more row on a piece list (d, dl): the runs tile the list's axis at U (ascending), run t goes to V[t] on the other side of
    the row and is the S[t]-th run there; every run takes the pieces it overlaps, clipped -- margpiece.cu's segmented copy."""
    R = len(U)
    UU = np.append(U, m)
    lo = np.searchsorted(d, U, side="right") - 1
    hi = np.searchsorted(d, UU[1:] - 1, side="right") - 1
    C = np.zeros(R, dtype=np.int64)
    C[S] = hi - lo + 1
    off = np.concatenate([[0], np.cumsum(C)])
    t_of = np.zeros(R, dtype=np.int64)
    t_of[S] = np.arange(R)
    o = np.arange(off[-1])
    s = np.searchsorted(off, o, side="right") - 1
    t = t_of[s]
    i = lo[t] + (o - off[s])
    return V[t] + np.maximum(d[i], U[t]) - U[t], dl[i] + U[t] - V[t]


def tail_count(d, dl, m, zt, P):
    """members of the group on ranks [zt, m) of the list's axis: per piece two look-ups in the reference row's prefix counts"""
    e = np.append(d[1:], m)
    a = np.maximum(d, zt)
    ok = e > a
    return int((P[e[ok] + dl[ok]] - P[a[ok] + dl[ok]]).sum())


def test_group_marginals_by_piece_lists(oracle):
    """The chains of margpiece.cu on encoded files: rows 0..15 of every 32-row group forward from the group vector in front of
    it, rows 31..16 backward from the one behind it; ones of the group in row k = members on ranks [m - n1_k, m) behind row k."""
    rng = np.random.default_rng(7)
    for mat, shift in ((haplo_matrix(200, 203, 29), 6), (random_matrix(96, 64, 30), 5), (edge_rows(300), 5)):
        pbf = oracle.encode_pbf(mat, shift=shift)
        m, rows, snaps = plane0_rows(pbf, oracle)
        in_group = rng.random(m) < 0.4
        BS, K, H = 1 << shift, 32, 16
        want = [int(((mat[r] & 1) == 1)[in_group].sum()) for r in range(len(rows))]
        # the group vector in front of every row (what the seed kernel stores every 32 rows), by the bit-vector partition
        vec = []
        for r, rle in enumerate(rows):
            if r % BS == 0:
                V = in_group[snaps[r // BS]]
            vec.append(V)
            V, _ = gather_partition(V, rle, m)
        vec.append(V)
        got = [None] * len(rows)
        for g0 in range(0, len(rows), K):
            full = g0 + K <= len(rows) and (g0 % BS) + K <= BS
            # forward
            P = np.concatenate([[0], np.cumsum(vec[g0])])
            d, dl = np.array([0]), np.array([0])
            for r in range(g0, min(g0 + (H if full else K), len(rows))):
                start, land, zi, zt, R = run_table(rows[r], m)
                d, dl = compose_list(d, dl, m, start, land, zi)
                assert d[0] == 0 and (np.diff(d) > 0).all() and len(d) <= 1 + sum(run_table(rows[q], m)[4] - 1 for q in range(g0, r + 1))
                got[r] = tail_count(d, dl, m, zt, P)
            if not full:
                continue
            # backward from the vector behind the group: in landing order the runs are U = land, V = start, and go back to run order
            P = np.concatenate([[0], np.cumsum(vec[g0 + K])])
            d, dl = np.array([0]), np.array([0])
            for r in range(g0 + K - 1, g0 + H - 1, -1):
                start, land, zi, zt, R = run_table(rows[r], m)
                got[r] = tail_count(d, dl, m, zt, P)               # (the list still belongs to row r + 1)
                U, V, S = np.zeros(R, np.int64), np.zeros(R, np.int64), np.zeros(R, np.int64)
                U[zi], V[zi], S[zi] = land, start, np.arange(R)
                d, dl = compose_list(d, dl, m, U, V, S)
        assert got == want


def both_planes(pbf_bytes, orc):
    """(m, shift, [(plane-0 RLE, plane-1 RLE) per row], [(S0, S1) per block])."""
    p = orc.Pbf(pbf_bytes)
    m, n, shift = p.m, p.n, p.shift
    p.close()
    buf = np.frombuffer(pbf_bytes, dtype=np.uint8)
    pos, rows, snaps = 16, [], []
    for _ in range(n):
        if buf[pos] == ord('S'):
            snaps.append(tuple(np.frombuffer(pbf_bytes, dtype=np.int32, count=m, offset=pos + 1 + 4 * m * g).astype(np.int64) for g in range(2)))
            pos += 1 + 8 * m
        pos += 1
        rec = []
        for _g in range(2):
            l = int(np.frombuffer(pbf_bytes, dtype=np.int32, count=1, offset=pos)[0])
            rec.append(pbf_bytes[pos + 4:pos + 4 + l])
            pos += 4 + l
        rows.append(tuple(rec))
    return m, shift, rows, snaps


def test_plane1_pairs_backwards_then_plane0_forwards(oracle):
    """The split scan's joint codes: every 1 bit of a plane-1 row is traced BACK through the non-empty plane-1 rows in front
    of it to the block's snapshot (which names the column: plane1.cu), and that column is walked FORWARD through plane 0 to
    the same row (WALK_QUERY), where it counts as other-ALT (plane-0 bit set, code 3) or missing (code 2)."""
    for mat, shift in ((haplo_matrix(200, 151, 25, p_missing_row=0.3, p_multi_row=0.3), 6), (random_matrix(40, 33, 26), 4)):
        pbf = oracle.encode_pbf(mat, shift=shift)
        m, shift, rows, snaps = both_planes(pbf, oracle)
        BS = 1 << shift
        for v, (rle0, rle1) in enumerate(rows):
            b0 = v - v % BS
            start, _, n1 = tables(rle1, m)
            c = np.frombuffer(rle1, dtype=np.uint8).astype(np.int64)
            ones = np.concatenate([np.arange(s, s + l) for s, l, b in zip(start, rle_len(c), c & 1) if b and l] or [np.zeros(0, np.int64)])
            assert len(ones) == n1
            # backwards: undo the partitions of the plane-1 rows in front (constant rows move nothing)
            x = ones.copy()
            for u in range(v - 1, b0 - 1, -1):
                S, D, k1 = run_map(rows[u][1], m)
                if k1 == 0 or k1 == m:
                    continue
                dst = S + D
                order = np.argsort(dst, kind="stable")
                x = x - D[order][np.searchsorted(dst[order], x, side="right") - 1]
            cols = snaps[v // BS][1][x]
            # forwards through plane 0 up to (and including) row v
            inv0 = np.empty(m, np.int64)
            inv0[snaps[v // BS][0]] = np.arange(m)
            r = inv0[cols]
            for u in range(b0, v):
                S, D, _ = run_map(rows[u][0], m)
                r = apply_map(S, D, r)
            S, D, k0 = run_map(rle0, m)
            bit0 = np.full(len(r), k0 == m) if k0 in (0, m) else apply_map(S, D, r) >= m - k0
            assert sorted(cols.tolist()) == np.flatnonzero(mat[v] & 2).tolist()
            assert int(bit0.sum()) == int((mat[v] == 3).sum()) and int((~bit0).sum()) == int((mat[v] == 2).sum())


# ---------------------------------------------------------------------------------------------------------------------
# Row-offset chase (index.cu): the records of a block are a chain of dependent length reads; K chasers share it, chaser l
# starting at the first position of its stretch of the byte range that VERIFIES as a record start.  The pieces count only
# if every chaser lands exactly on its successor's start and the rows add up -- otherwise the block is walked alone.

def verify_start(img, p, end, m, need=4):
    def rec(q):
        if q + 9 > end or img[q] != ord('B'):
            return None
        l0 = int.from_bytes(img[q + 1:q + 5], "little", signed=True)
        if l0 < 0 or q + 9 + l0 > end:
            return None
        l1 = int.from_bytes(img[q + 5 + l0:q + 9 + l0], "little", signed=True)
        if l1 < 0 or q + 9 + l0 + l1 > end:
            return None
        return l0, l1
    first = rec(p)
    if first is None:
        return False
    l0, l1 = first
    for off, l in ((p + 5, l0), (p + 9 + l0, l1)):           # run lengths of both planes of the first record sum to m
        c = np.frombuffer(img[off:off + l], dtype=np.uint8).astype(np.int64)
        if int(rle_len(c).sum()) != m:
            return False
    for v in range(need):                                      # ... and `need` well-formed records chain from here
        if p == end:
            return v > 0
        r = rec(p)
        if r is None:
            return False
        p += 9 + r[0] + r[1]
    return True


def team_chase(img, first, end, m, rows, K):
    span = end - first
    starts = [first]
    for l in range(1, K):
        s = None
        for p in range(first + span * l // K, first + span * (l + 1) // K):
            if img[p] == ord('B') and verify_start(img, p, end, m):
                s = p
                break
        starts.append(s)
    pieces = []
    for l in range(K):
        if starts[l] is None:
            pieces.append([])
            continue
        stop = next((s for s in starts[l + 1:] if s is not None), end)
        o, mine = starts[l], []
        while o < stop:
            l0 = int.from_bytes(img[o + 1:o + 5], "little")
            l1 = int.from_bytes(img[o + 5 + l0:o + 9 + l0], "little")
            mine.append(o)
            o += 9 + l0 + l1
        if o != stop:
            return None                                        # ran past the successor: fall back to one chaser
        pieces.append(mine)
    chain = [o for mine in pieces for o in mine]
    return chain if len(chain) == rows else None


def test_team_chase_stitches_the_true_chain(oracle):
    m = 2048
    mat = fake_tag_matrix(300, m, 9)            # runs of 256..511 zeros are coded with the byte 0x42 = 'B'
    pbf = oracle.encode_pbf(mat, shift=13)
    assert pbf.count(b"B") > 4 * 300                            # far more 'B' bytes than records
    first = 16 + 1 + 8 * m                                      # behind the 'S' record
    true_chain, o = [], first
    for _ in range(300):
        l0 = int.from_bytes(pbf[o + 1:o + 5], "little")
        l1 = int.from_bytes(pbf[o + 5 + l0:o + 9 + l0], "little")
        true_chain.append(o)
        o += 9 + l0 + l1
    end = o
    assert pbf[end:end + 1] == b"I"                             # the index record follows the last row
    fakes = sum(1 for q in range(first, end) if pbf[q] == ord("B") and q not in set(true_chain))
    assert fakes > 300 and not any(verify_start(pbf, q, end, m) for q in range(first, end) if pbf[q] == ord("B") and q not in set(true_chain))
    for K in (2, 6, 12):
        assert team_chase(pbf, first, end, m, 300, K) == true_chain
