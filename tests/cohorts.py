"""Seeded synthetic genotype matrices used by the parity tests (test infrastructure)."""
import numpy as np


def haplo_matrix(n_rows, m, seed, p_missing_row=1 / 16, p_multi_row=1 / 16, founders=12, switch=0.02):
    """A [n_rows, m] uint8 matrix of 2-bit codes with haplotype-block structure so that the PBWT compresses:
    every column copies one of `founders` founder haplotypes and switches founder with probability `switch`
    per row; rows sprinkle missing (2) / other-ALT (3) codes."""
    rng = np.random.default_rng(seed)
    f = rng.integers(0, founders, size=m)
    freq = np.clip(rng.beta(0.3, 2.0, size=n_rows), 0, 1)
    out = np.zeros((n_rows, m), dtype=np.uint8)
    for k in range(n_rows):
        alleles = (rng.random(founders) < freq[k]).astype(np.uint8)
        row = alleles[f]
        noise = rng.random(m) < 0.002
        row = row ^ noise.astype(np.uint8)
        r = rng.random()
        if r < p_missing_row:
            idx = rng.integers(0, m, size=max(1, m // 50))
            row[idx] = 2
        elif r < p_missing_row + p_multi_row:
            idx = rng.integers(0, m, size=max(1, m // 30))
            row[idx] = 3
        out[k] = row
        sw = rng.random(m) < switch
        f = np.where(sw, rng.integers(0, founders, size=m), f)
    return out


def random_matrix(n_rows, m, seed, probs=(0.55, 0.3, 0.08, 0.07)):
    rng = np.random.default_rng(seed)
    return rng.choice(4, size=(n_rows, m), p=probs).astype(np.uint8)


def edge_rows(m):
    """Rows that exercise the RLE alphabet: all-0, all-1, alternating, long runs with every nibble position."""
    rows = []
    z = np.zeros(m, np.uint8)
    rows.append(z.copy())
    rows.append(np.ones(m, np.uint8))
    rows.append((np.arange(m) & 1).astype(np.uint8))
    rows.append(((np.arange(m) >> 1) & 1).astype(np.uint8) * 3)
    for run in (15, 16, 17, 255, 256, 257, 4095, 4096, 65535, 65536, 1048576):
        if run < m:
            r = z.copy(); r[run:] = 1; rows.append(r)
            r = z.copy(); r[:run] = 2; rows.append(r)
            r = ((np.arange(m) // run) & 1).astype(np.uint8); rows.append(r)
    r = z.copy(); r[m - 1] = 1; rows.append(r)
    r = z.copy(); r[0] = 3; rows.append(r)
    rows.append(np.full(m, 2, np.uint8))
    rows.append(np.full(m, 3, np.uint8))
    return np.array(rows, dtype=np.uint8)


def fake_tag_matrix(n_rows, m, seed):
    """Rows whose plane-0 runs of 256..511 zeros are coded with the byte 0x42 = 'B', the record tag: the record bytes are
    full of fake tags (the device-side row index starts its chasers in the middle of a block, index.cu)."""
    rng = np.random.default_rng(seed)
    mat = np.zeros((n_rows, m), np.uint8)
    for r in range(n_rows):
        pos = 0
        while pos < m:
            pos += int(rng.integers(256, 512))
            if pos < m:
                mat[r, pos:pos + int(rng.integers(1, 4))] = 1 + 2 * int(rng.integers(0, 2))
                pos += 3
    return mat
