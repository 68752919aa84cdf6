"""Multi-GPU parity of the SFC-partitioned dycore, run under torch.distributed.run:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_parity.py [--preset ne8] [--calls 2]

Every rank owns one contiguous run of the space-filling curve, exchanges DSS / min-max halos
with its neighbour ranks over NCCL, and runs `calls` prim_run_subcycle_c calls. Rank 0 then
repeats the run on the WHOLE mesh on its own GPU and the gathered per-element results must be
bit-identical (the node-centric DSS keeps the reference's unpack order for off-rank sharers
too, so the partition must not change a single bit)."""
import argparse
import ctypes as C
import os
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def new_comm(lib, dist, torch, rank, world, local_rank):
    """A fresh ncclUniqueId from rank 0 to every rank, then hommexx_b200_set_comm (before the session starts)."""
    idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = (C.c_ubyte * 128)()
        assert lib.hommexx_b200_nccl_unique_id(raw) == 0
        idbuf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
    dist.broadcast(idbuf, 0)
    raw = (C.c_ubyte * 128)(*idbuf.cpu().tolist())
    lib.hommexx_b200_set_comm(rank, world, local_rank, raw)


def one_case(preset, over, calls, dist, torch, rank, world, local_rank):
    """Multi-rank run of `preset`, then the whole mesh on rank 0's GPU alone: every bit must agree."""
    from hommexx_b200 import homme
    import distinct_tracers

    cfg = homme.preset(preset, npart=world, **over)
    cfg.part_id = rank
    libpath = homme.cuda_lib_path(cfg.nlev, cfg.qsize_d)
    lib = homme.load_dycore(libpath)
    new_comm(lib, dist, torch, rank, world, local_rank)
    h = homme.Homme(cfg, libpath)
    if cfg.qsize > 4:
        distinct_tracers.install(h)
    h.init_dycore()
    for _ in range(calls):
        h.run_subcycle()
    h.push_results()
    mine = {k: v.copy() for k, v in h.state().items()}  # the views die with the driver
    gids = h.local_gids()
    tl = h.time_levels()
    h.close()

    gathered = [None] * world
    dist.all_gather_object(gathered, (gids, mine, tl))
    ok = True
    if rank == 0:
        lib.hommexx_b200_set_comm(0, 1, local_rank, None)
        cfg1 = homme.preset(preset, npart=1, **over)
        h1 = homme.Homme(cfg1, libpath)
        if cfg1.qsize > 4:
            distinct_tracers.install(h1)
        h1.init_dycore()
        for _ in range(calls):
            h1.run_subcycle()
        h1.push_results()
        ref = {k: v.copy() for k, v in h1.state().items()}
        ref_gids = h1.local_gids()
        pos = {int(g): i for i, g in enumerate(ref_gids)}
        assert h1.time_levels() == tl
        nel = sum(len(g) for g, _, _ in gathered)
        assert nel == h1.nelem, (nel, h1.nelem)
        for g_r, st_r, tl_r in gathered:
            assert tl_r == tl
            idx = np.array([pos[int(g)] for g in g_r])
            for k, v in st_r.items():
                if not np.array_equal(v, ref[k][idx]):
                    d = np.abs(v - ref[k][idx]).max()
                    print(f"MISMATCH field {k}: max |diff| = {d:.3e}", flush=True)
                    ok = False
        h1.close()
        print(f"multi_gpu_parity: {world} ranks, preset {preset} {over}, {nel} elements, {calls} calls: "
              f"{'bit-identical to the single-GPU run' if ok else 'FAILED'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    return bool(int(flag.item()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="ne8")
    ap.add_argument("--calls", type=int, default=2)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT / "tests"))
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    ok = True
    # Three multi-rank sessions in ONE process (a session's halo tables must not leak into the next one): the
    # 4-tracer build, then the benchmarked (72, 40) build with 40 distinct tracers, then the first one again on a
    # different mesh.
    for preset, over in ((args.preset, {}), ("ne8", dict(qsize=40, qsize_d=40)), ("ne4", {})):
        ok = one_case(preset, over, args.calls, dist, torch, rank, world, local_rank) and ok
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
