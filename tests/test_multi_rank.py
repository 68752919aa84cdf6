"""N > 1 path. CPU part (gloo, world_size 2): the host-side partition and connectivity each
rank hands to init_connectivity/add_connection are mutually consistent. GPU part: 2-rank NCCL
run of the CUDA dycore, bit-identical to the single-GPU run (skipped with fewer than 2 GPUs)."""
import os
import pathlib
import subprocess
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]


def _rank_main(rank, world, port, ne, q):
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    from hommexx_b200 import homme
    from oracle import oraclelib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = homme.preset("ne4", ne=ne, npart=world)
        cfg.part_id = rank
        h = homme.Homme(cfg, oraclelib.ORACLE_LIB)  # driver only: no dycore call is made
        gids = h.local_gids()
        conn = h.connections()                  # (lid1,gid1,pos1,pid1, lid2,gid2,pos2,pid2), all 1-based
        sph = h.array("spheremp").copy()
        every = [None] * world
        dist.all_gather_object(every, (gids, conn, float(sph.sum())))
        # 1. the partition covers every element exactly once, sizes follow genspacepart
        allg = np.concatenate([g for g, _, _ in every])
        assert sorted(allg.tolist()) == list(range(h.nelem))
        sizes = [len(g) for g, _, _ in every]
        base, extra = divmod(h.nelem, world)
        assert sizes == [base + (1 if r < extra else 0) for r in range(world)]
        # 2. every connection to another rank has its mirror image on that rank
        mirror = {}
        for r, (_, c, _) in enumerate(every):
            for t in c:
                assert t[3] == r + 1
                mirror[(r + 1, t[1], t[2])] = (t[7], t[5], t[6], t[4])  # (my pid, gid, pos) -> remote (pid, gid, pos, lid)
        for (pid, gid, pos), (rpid, rgid, rpos, rlid) in mirror.items():
            back = mirror.get((rpid, rgid, rpos))
            assert back is not None, ("no mirror for", pid, gid, pos)
            assert back[0] == pid and back[1] == gid and back[2] == pos
        # 3. remote lids are the owner's local ids
        for r, (_, c, _) in enumerate(every):
            for t in c:
                owner_gids = every[t[7] - 1][0]
                assert owner_gids[t[4] - 1] == t[5] - 1
        # 4. the sphere's area is the sum over ranks
        assert abs(sum(s for _, _, s in every) - 4 * np.pi) < 1e-9
        h.close()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ne", [4, 5])
def test_partition_and_connectivity_two_ranks_gloo(ne):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + ne
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, ne, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


@pytest.mark.gpu
def test_two_gpu_run_is_bit_identical_to_single_gpu():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", str(ROOT / "tests" / "multi_gpu_parity.py"), "--preset", "ne8"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "bit-identical" in r.stdout


@pytest.mark.parametrize("ne,npart", [(6, 8), (10, 4), (7, 3)])
def test_partition_properties_many_parts(ne, npart):
    """The SFC partition the 4- and 8-GPU runs use, checked in one process: every element owned once, sizes as
    genspacepart (spacecurve_mod.F90:1232-1264), parts spatially compact (at most two edge-connected patches and a
    perimeter-sized halo), and every off-rank connection mirrored by the owner."""
    from hommexx_b200 import homme
    from oracle import oraclelib
    parts = []
    for r in range(npart):
        cfg = homme.preset("ne4", ne=ne, npart=npart)
        cfg.part_id = r
        h = homme.Homme(cfg, oraclelib.ORACLE_LIB)
        parts.append((h.local_gids(), h.connections(), float(h.array("spheremp").sum())))
        nelem = h.nelem
        h.close()
    allg = np.concatenate([g for g, _, _ in parts])
    assert sorted(allg.tolist()) == list(range(nelem))
    base, extra = divmod(nelem, npart)
    assert [len(g) for g, _, _ in parts] == [base + (1 if r < extra else 0) for r in range(npart)]
    assert abs(sum(s for _, _, s in parts) - 4 * np.pi) < 1e-9
    mirror, halo = {}, []
    for r, (gids, conn, _) in enumerate(parts):
        mine = set((gids + 1).tolist())
        # edge-connected patch: flood fill over on-rank EDGE connections (pos 1..4) reaches every element
        adj = {g: set() for g in mine}
        off = 0
        for t in conn:
            assert t[3] == r + 1
            if t[7] == r + 1:
                if t[2] <= 4:
                    adj[t[1]].add(t[5])
            else:
                off += 1
                mirror[(r + 1, t[1], t[2])] = (t[7], t[5], t[6])
        left, patches = set(mine), 0
        while left:
            patches += 1
            todo = [next(iter(left))]
            while todo:
                g = todo.pop()
                if g not in left:
                    continue
                left.discard(g)
                todo.extend(adj[g] & left)
        # the per-face curves are joined end to start, which is not always an element adjacency: a part
        # that spans a face change may be two patches, never more
        assert patches <= 2, f"part {r} is scattered over {patches} patches"
        halo.append(off)
    for (pid, gid, pos), (rpid, rgid, rpos) in mirror.items():
        assert mirror.get((rpid, rgid, rpos)) == (pid, gid, pos)
    # compactness: the halo of a part is a perimeter, far below its 8 * size connections
    assert max(halo) < 0.75 * 8 * (base + 1)
