"""GPU parity of the BGZF inflate kernel (inflate.cu; bgzf.c:225-249 inflate_block, :318-351 bgzf_read_block) against
zlib: every DEFLATE block type (stored, fixed, dynamic), overlapping matches, multi-block files, the empty EOF block,
and the site-only .bcf / .csi written by the reference library."""
import gzip
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

from cohorts import haplo_matrix

pytestmark = pytest.mark.gpu
EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")   # bgzf.c:51-57


def bgzf_block(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, extra=b""):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    raw = c.compress(data) + c.flush()
    xlen = 6 + len(extra)
    total = 12 + xlen + len(raw) + 8
    sub = extra + b"BC" + struct.pack("<HH", 2, total - 1)
    return b"\x1f\x8b\x08\x04\0\0\0\0\x00\xff" + struct.pack("<H", xlen) + sub + raw + struct.pack("<II", zlib.crc32(data), len(data))


def bgzf_file(data, chunk=65280, **kw):
    out = b"".join(bgzf_block(data[i:i + chunk], **kw) for i in range(0, len(data), chunk))
    return out + EOF_BLOCK


@pytest.fixture(scope="module")
def b200():
    import bgt_b200
    return bgt_b200


@pytest.fixture(scope="module")
def ctx(b200):
    c = b200.Context(0)
    yield c
    c.close()


def payloads():
    rng = np.random.default_rng(3)
    text = b"".join(b"11\t%d\t.\tA\tC\t0\t.\tAN=%d;AC=%d\n" % (1000 + 10 * i, 200000 - i % 7, i % 311) for i in range(20000))
    return {
        "random": rng.integers(0, 256, 200000, dtype=np.uint8).tobytes(),
        "text": text,
        "zeros": bytes(150000),
        "period3": b"abc" * 50000,
        "skewed": rng.choice(np.arange(8, dtype=np.uint8), 300000, p=[.5, .2, .1, .08, .05, .04, .02, .01]).tobytes(),
        "tiny": b"x",
        "empty": b"",
    }


@pytest.mark.parametrize("level,strategy", [(0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY),
                                            (9, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)])
def test_inflate_equals_zlib(b200, ctx, level, strategy):
    for name, data in payloads().items():
        f = bgzf_file(data, level=level, strategy=strategy)
        assert gzip.decompress(f) == data                      # the test's own writer is sane
        assert b200.bgzf_inflate(ctx, f) == data, (name, level, strategy)


def test_block_geometry(b200, ctx):
    """full 64 KiB blocks, ragged block sizes, extra gzip subfields in front of 'BC', no EOF block."""
    rng = np.random.default_rng(4)
    data = (b"ACGT" * 7 + rng.integers(0, 4, 1000, dtype=np.uint8).tobytes()) * 300
    assert b200.bgzf_inflate(ctx, bgzf_file(data, chunk=65536)) == data
    assert b200.bgzf_inflate(ctx, bgzf_file(data, chunk=777)[:-len(EOF_BLOCK)]) == data
    assert b200.bgzf_inflate(ctx, bgzf_file(data, chunk=30000, extra=b"XY\x03\x00abc")) == data


def test_corrupt_file_is_reported(b200, ctx):
    f = bytearray(bgzf_file(b"hello world " * 5000))
    with pytest.raises(b200.B200Error):
        b200.bgzf_inflate(ctx, bytes(f[:100]))                 # truncated
    f[40] ^= 0xff; f[41] ^= 0x55; f[60] ^= 0xAA                # damaged DEFLATE stream
    with pytest.raises(b200.B200Error):
        b200.bgzf_inflate(ctx, bytes(f))
    with pytest.raises(b200.B200Error):
        b200.bgzf_inflate(ctx, b"not a bgzf file at all........")


def test_reference_written_bcf_and_csi(b200, ctx, ref, tmp_path):
    """the site side of a BGT database as the reference library writes it (mksites: vcf_write1 + bcf_index_build)."""
    mat = haplo_matrix(5000, 64, 12)
    prefix = str(tmp_path / "s.bgt")
    with open(prefix + ".pbf", "wb") as f:
        f.write(ref.encode_pbf(mat, shift=13))
    subprocess.run([ref.MKSITES, prefix], check=True, stderr=subprocess.DEVNULL)
    for ext in (".bcf", ".bcf.csi"):
        raw = open(prefix + ext, "rb").read()
        assert b200.bgzf_inflate(ctx, raw) == gzip.decompress(raw), ext
