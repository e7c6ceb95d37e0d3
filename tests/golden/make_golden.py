"""Regenerate tests/golden/* with the UNMODIFIED reference (oracle/_ref/pbfview, built from /root/reference).

Run from the repo root in the build container:  python tests/golden/make_golden.py
The .pbf files are the reference encoder's bytes for the seeded matrices in the .npy files.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from cohorts import haplo_matrix, random_matrix  # noqa: E402
from oracle import oracle as orc  # noqa: E402

EX1 = np.array([[0, 1, 2, 0], [2, 0, 1, 1], [1, 0, 1, 1], [0, 1, 0, 1], [1, 2, 0, 0], [1, 0, 1, 2], [0, 1, 1, 1]], np.uint8)

with open(os.path.join(HERE, "ex1.pbf"), "wb") as f:
    f.write(orc.ref_run(["pbfview", "-Sb", "-"], stdin=orc.pim_text(EX1), seekable_stdout=True))
for name, mat in (("hap_200x96", haplo_matrix(200, 96, 21)), ("rnd_300x37", random_matrix(300, 37, 22))):
    np.save(os.path.join(HERE, name + ".npy"), mat)
    with open(os.path.join(HERE, name + ".s5.pbf"), "wb") as f:
        f.write(orc.ref_run(["pbfview", "-Sb", "-s", "5", "-"], stdin=orc.pim_text(mat), seekable_stdout=True))
print("golden written")
