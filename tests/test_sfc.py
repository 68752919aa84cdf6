"""HOMME's space-filling-curve decomposition (spacecurve_mod.F90:39-1040, cube_mod.F90:1457-1587) as restated in
hommexx_b200/driver/homme_driver.cpp: nested Hilbert / meandering-Peano / Cinco refinements for ne = 2^a 3^b 5^c,
laid over the six faces in the order 1, 2, 6, 4, 5, 3 so that the curve never jumps."""
import numpy as np
import pytest

from hommexx_b200 import homme
from oracle import oraclelib


def sfc(ne, npart=1, part=0):
    cfg = homme.preset("ne4", ne=ne, npart=npart)
    cfg.part_id = part
    h = homme.Homme(cfg, oraclelib.ORACLE_LIB, init="none")     # driver only: no dycore call is made
    g, conn = h.local_gids(), h.connections()
    h.close()
    return g, conn


def face_mesh(ne):
    """Mesh(i, j) of CubeTopology, read back from face 5 (fifth on the curve, laid down without reflection)."""
    g, _ = sfc(ne)
    seg = g[4 * ne * ne:5 * ne * ne]
    assert (seg // (ne * ne) == 4).all()           # face 5
    mesh = np.empty((ne, ne), dtype=int)           # [i][j]
    for k, gid in enumerate(seg):
        mesh[gid % ne, (gid // ne) % ne] = k
    return mesh


def test_hilbert_peano_cinco_visit_orders():
    """The sub-cell visiting orders the reference documents position by position (spacecurve_mod.F90:683-769
    Hilbert, :506-681 PeanoM, :39-503 Cinco)."""
    hil = face_mesh(2)
    assert [tuple(np.argwhere(hil == k)[0]) for k in range(4)] == [(0, 0), (0, 1), (1, 1), (1, 0)]
    pea = face_mesh(3)
    assert [tuple(np.argwhere(pea == k)[0]) for k in range(9)] == [(0, 0), (0, 1), (0, 2), (1, 2), (2, 2), (2, 1),
                                                                   (1, 1), (1, 0), (2, 0)]
    cin = face_mesh(5)
    path = [tuple(np.argwhere(cin == k)[0]) for k in range(25)]
    assert path[0] == (0, 0) and path[-1] == (4, 0)                      # enters and leaves along the first axis
    assert all(abs(a[0] - b[0]) + abs(a[1] - b[1]) == 1 for a, b in zip(path[:-1], path[1:]))


@pytest.mark.parametrize("ne", [2, 3, 4, 5, 6, 8, 9, 10, 12, 15, 16, 20, 30])
def test_curve_is_continuous_over_the_whole_cubed_sphere(ne):
    """For ne = 2^a 3^b 5^c consecutive elements of the curve share an EDGE everywhere, including where the curve
    passes from one cube face to the next."""
    g, conn = sfc(ne)
    assert sorted(g.tolist()) == list(range(6 * ne * ne))
    edge_nbr = {}
    for t in conn:                                 # (lid1,gid1,pos1,pid1, lid2,gid2,pos2,pid2), 1-based; pos <= 4 = edge
        if t[2] <= 4:
            edge_nbr.setdefault(t[1] - 1, set()).add(t[5] - 1)
    assert all(b in edge_nbr[a] for a, b in zip(g[:-1], g[1:]))
    faces = g // (ne * ne) + 1                     # face order along the curve
    assert [int(faces[k * ne * ne]) for k in range(6)] == [1, 2, 6, 4, 5, 3]


@pytest.mark.parametrize("ne", [7, 11, 13])
def test_non_factorable_ne_uses_the_sampled_power_of_two_curve(ne):
    """cube_mod.F90:1466-1523: every element still appears exactly once; the order follows the 2^k curve."""
    g, _ = sfc(ne)
    assert sorted(g.tolist()) == list(range(6 * ne * ne))
    assert sorted(face_mesh(ne).ravel().tolist()) == list(range(ne * ne))


@pytest.mark.parametrize("ne,npart", [(30, 8), (8, 3), (6, 4), (20, 7)])
def test_every_part_is_one_connected_patch(ne, npart):
    """genspacepart (spacecurve_mod.F90:1218-1273) cuts the continuous curve into contiguous runs: each rank owns ONE
    edge-connected patch, sizes differ by at most one."""
    sizes = []
    for part in range(npart):
        g, conn = sfc(ne, npart, part)
        sizes.append(len(g))
        mine = set(g.tolist())
        nbr = {}
        for t in conn:
            if t[2] <= 4 and (t[5] - 1) in mine:
                nbr.setdefault(t[1] - 1, set()).add(t[5] - 1)
        seen, stack = {int(g[0])}, [int(g[0])]
        while stack:
            for b in nbr.get(stack.pop(), ()):
                if b not in seen:
                    seen.add(b); stack.append(b)
        assert seen == mine, (ne, npart, part, len(seen), len(mine))
    assert sum(sizes) == 6 * ne * ne and max(sizes) - min(sizes) <= 1
