"""GPU parity of the device-side PBWT encoder (b200_enc_*: pbf_open_w / pbf_write / pbf_close, pbwt.c:199-219, 288-311,
264-286) against the oracle's restatement of the reference encoder -- the .pbf images must be byte-identical -- and
against the golden files the reference's own pbfview wrote."""
import os

import numpy as np
import pytest

from cohorts import edge_rows, haplo_matrix, random_matrix

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
EX1 = np.array([[0, 1, 2, 0], [2, 0, 1, 1], [1, 0, 1, 1], [0, 1, 0, 1], [1, 2, 0, 0], [1, 0, 1, 2], [0, 1, 1, 1]], np.uint8)


@pytest.fixture(scope="module")
def b200():
    import bgt_b200
    return bgt_b200


@pytest.fixture(scope="module")
def ctx(b200):
    c = b200.Context(0)
    yield c
    c.close()


def gpu_encode(b200, ctx, mat, shift, chunks=None, as_bits=False):
    m = mat.shape[1]
    enc = b200.Encoder(ctx, m, shift)
    bounds = [0] + list(chunks or []) + [mat.shape[0]]
    for a, b in zip(bounds[:-1], bounds[1:]):
        part = mat[a:b]
        if as_bits:
            words = (m + 31) // 32
            bits = np.zeros((part.shape[0], 2, words * 32), np.uint8)
            bits[:, 0, :m] = part & 1
            bits[:, 1, :m] = part >> 1
            packed = np.packbits(bits.reshape(part.shape[0], 2, words, 32), axis=-1, bitorder="little").view(np.uint32).reshape(part.shape[0], 2, words)
            enc.write_bits(packed)
        else:
            enc.write((part & 1).astype(np.uint8), (part >> 1).astype(np.uint8))
    img = enc.finish()
    enc.close()
    return img


def test_ex1_matches_reference_file(b200, ctx):
    """ex1.pim through the reference's `pbfview -Sb` (golden ex1.pbf, shift 13)."""
    with open(os.path.join(GOLD, "ex1.pbf"), "rb") as f:
        want = f.read()
    assert gpu_encode(b200, ctx, EX1, 13) == want


@pytest.mark.parametrize("name,shift", [("hap_200x96", 5), ("rnd_300x37", 5)])
def test_golden_files(b200, ctx, name, shift):
    mat = np.load(os.path.join(GOLD, name + ".npy"))
    with open(os.path.join(GOLD, "%s.s%d.pbf" % (name, shift)), "rb") as f:
        want = f.read()
    assert gpu_encode(b200, ctx, mat, shift) == want


@pytest.mark.parametrize("rows,m,shift,seed", [(300, 200, 7, 5), (1000, 1031, 8, 6), (257, 33, 4, 7), (64, 1, 3, 8), (50, 2, 13, 9),
                                               (130, 70001, 6, 10), (8300, 500, 13, 11)])
def test_encoder_equals_oracle(b200, ctx, rows, m, shift, seed):
    from oracle import oracle as orc
    mat = haplo_matrix(rows, m, seed) if seed % 2 else random_matrix(rows, m, seed)
    assert gpu_encode(b200, ctx, mat, shift) == orc.encode_pbf(mat, shift=shift)


def test_rle_alphabet_edges_and_batches(b200, ctx):
    """all-0 / all-1 / alternating rows, runs at every hex-digit boundary (pbwt.c:24-36), written in uneven batches, as
    bytes and as bit planes; the decoder reads the image back to the same matrix."""
    from oracle import oracle as orc
    m = 70001
    e = edge_rows(m)
    mat = np.concatenate([e, e[::-1]])
    want = orc.encode_pbf(mat, shift=4)
    assert gpu_encode(b200, ctx, mat, 4, chunks=[1, 2, 17, 40]) == want
    assert gpu_encode(b200, ctx, mat, 4, chunks=[33], as_bits=True) == want
    pb = b200.Pbf.from_bytes(ctx, want)
    q = b200.Query.columns(ctx, pb)
    got = b200.scan(ctx, pb, q, 0, mat.shape[0], hap_bytes=True)
    assert ((got["hap_bytes"][0] | got["hap_bytes"][1] << 1) == mat).all()
    q.close(); pb.close()


def test_wide_cohort_round_trip(b200, ctx):
    """m = 1,000,001 columns (config-5 width, not a multiple of 32): encode on the device, decode on the device, and the
    first rows against the oracle's encoder."""
    from oracle import oracle as orc
    m, rows = 1000001, 40
    mat = haplo_matrix(rows, m, 21, founders=40)
    img = gpu_encode(b200, ctx, mat, 5)
    assert img == orc.encode_pbf(mat, shift=5)
    pb = b200.Pbf.from_bytes(ctx, img)
    q = b200.Query.columns(ctx, pb)
    got = b200.scan(ctx, pb, q, 0, rows, hap_bytes=True)
    assert ((got["hap_bytes"][0] | got["hap_bytes"][1] << 1) == mat).all()
    q.close(); pb.close()
