"""N>1 host logic of the partitioned fusion path (onepiece_b200/fusion.py) on CPU: world_size-2/3 gloo process groups
drive exchange_halo / gather_mesh with an oracle-backed volume, and the union of the per-rank Marching-Cubes meshes must
be the unsharded mesh, triangle for triangle."""
import os
import socket
import tempfile

import numpy as np
import pytest

from conftest import canon_triangles
from fusion_common import gloo_worker, small_scene


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_owner_and_shard_range():
    from onepiece_b200.fusion import halo_peers, owner_of, shard_range
    ids = np.array([[-9, 0, 0], [-8, 0, 0], [-1, 5, 0], [0, 0, 0], [3, 0, 0], [4, 0, 0], [8, 0, 0]], np.int32)
    assert owner_of(ids, 0, 4, 2).tolist() == [1, 0, 1, 0, 0, 1, 0]   # floor division: -9//4 = -3 -> 1, -8//4 = -2 -> 0
    assert halo_peers(0, 4) == (3, 1) and halo_peers(3, 4) == (2, 0)
    for n, w in [(10, 3), (307200, 8), (5, 8), (0, 2)]:
        parts = [shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1


@pytest.mark.parametrize("world,axis,slab", [(2, 0, 2), (2, 1, 1), (3, 2, 1)])
def test_gloo_halo_exchange_reproduces_the_unsharded_mesh(world, axis, slab):
    import torch.multiprocessing as mp
    _, _, _, ids, _, (pts, col) = small_scene()
    assert len(pts) > 1000
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "mesh.npz")
        mp.spawn(gloo_worker, args=(world, _free_port(), axis, slab, out), nprocs=world, join=True)
        z = np.load(out)
    stats = z["stats"]
    assert stats[:, 0].sum() == len(ids)                      # ownership is a partition
    assert (stats[:, 0] > 0).all()                            # every rank holds part of the surface band
    assert stats[:, 2].sum() == stats[:, 1].sum() > 0         # every exported cube arrived somewhere
    assert np.array_equal(canon_triangles(z["points"], z["colors"]), canon_triangles(pts, col))
    # without the exchange the seams are missing: the per-rank vertex counts only add up thanks to the ghosts
    assert stats[:, 3].sum() == len(pts)
