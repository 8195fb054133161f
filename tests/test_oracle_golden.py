"""The plain-C oracle against the committed golden fixtures (generated from the compiled reference by
tests/golden/gen_golden.py).  No GPU and no /root/reference needed."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_bit_equal, canon_triangles, sha
from onepiece_b200 import scenes
from oracle import oracleapi

FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "integrate_*.npz")))


def load(path):
    g = np.load(path)
    c = g["cam"]
    cam = scenes.Camera(float(c[0]), float(c[1]), float(c[2]), float(c[3]), int(c[4]), int(c[5]), float(c[6]))
    return g, cam


def test_fixtures_exist():
    assert len(FIXTURES) >= 3


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_oracle_reproduces_golden(path):
    g, cam = load(path)
    ov = oracleapi.OracleVolume(cam, float(g["res"]), float(g["trunc"]))
    for d, c, T in zip(g["depth"], g["bgr"], g["poses"]):
        ov.integrate(d, c, T)
    ids, vox = ov.download()
    assert np.array_equal(ids, g["ids"])
    assert_bit_equal(vox[:16], g["voxels_head"], "first cubes")
    assert sha(vox) == str(g["voxels_sha"])
    pts, col = ov.extract_mesh()
    assert len(pts) == int(g["n_vertices"]) and len(pts) == 3 * int(g["n_triangles"])
    canon = canon_triangles(pts, col)
    assert_bit_equal(canon[:64], g["mesh_head"], "first triangles")
    assert sha(canon) == str(g["mesh_sha"])
