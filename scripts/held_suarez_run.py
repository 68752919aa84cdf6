#!/usr/bin/env python
"""BASELINE configs[4]: Held-Suarez forced run at ne256 (393 216 elements), nlev 72, qsize 40 on the 8 GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 \
        scripts/held_suarez_run.py --ne 256 --calls 30

Every call is hxx_held_suarez_forcing (the Held-Suarez physics of physics/heldsuarez/held_suarez_mod.F90 evaluated on
the device, FM / FT never leave HBM) followed by prim_run_subcycle_c with ftype = 0. Reports whole-job element-steps/s,
SYPD, device memory per GPU and the clocks under load as ONE JSON line on stdout (rank 0). Namelist: the reference's
ne120 benchmark file scaled the way HOMME scales with resolution (tstep ~ 1/ne, nu ~ dx^3.2)."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import pathlib
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402  (ClockSampler, emit; points fd 1 at stderr so the JSON line stays alone on stdout)


def host_gb_available():
    for ln in open("/proc/meminfo"):
        if ln.startswith("MemAvailable"):
            return int(ln.split()[1]) / 1e6
    return 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ne", type=int, default=256)
    ap.add_argument("--calls", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--qsize", type=int, default=40)
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from hommexx_b200 import homme

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    # the driver keeps the Fortran-side arrays (v, T, dp3d, Qdp, Q, FM, FT ~ 136 field tiles per element) in host
    # memory on every rank, plus the global vorticity field of the JW tracer initialisation
    ne = args.ne
    while ne > 30:
        per_rank = 6 * ne * ne / world * 136 * 16 * 72 * 8 / 1e9 + 6 * ne * ne * 16 * 72 * 8 / 1e9
        if per_rank * world < 0.8 * host_gb_available():
            break
        ne -= 32
    if ne != args.ne and rank == 0:
        bench.log(f"host memory: {host_gb_available():.0f} GB available, running ne{ne} instead of ne{args.ne}")
    scale = 120.0 / ne
    nu = 1e13 * scale ** 3.2
    cfg = homme.preset("ne120", ne=ne, npart=world, qsize=args.qsize, tstep=75.0 * scale, nu=nu, nu_p=nu, nu_q=nu,
                       nu_s=nu, ftype=0)
    cfg.part_id = rank
    libpath = homme.cuda_lib_path(cfg.nlev, cfg.qsize_d, os.environ.get("HXX_FLAVOUR", "fma"))
    if not libpath.exists():
        libpath = homme.cuda_lib_path(cfg.nlev, cfg.qsize_d)
    lib = homme.load_dycore(libpath)
    lib.hommexx_b200_event_elapsed_ms.restype = C.c_double
    if world > 1:
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = (C.c_ubyte * 128)()
            assert lib.hommexx_b200_nccl_unique_id(raw) == 0
            idbuf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        raw = (C.c_ubyte * 128)(*idbuf.cpu().tolist())
        lib.hommexx_b200_set_comm(rank, world, local_rank, raw)
    else:
        lib.hommexx_b200_set_comm(0, 1, local_rank, None)
    t0 = time.perf_counter()
    h = homme.Homme(cfg, libpath)
    h.init_dycore()
    bench.log(f"[rank {rank}] setup {time.perf_counter() - t0:.1f}s: ne={cfg.ne} nelem={h.nelem} local={h.nelemd}")
    dyn = cfg.rsplit * cfg.qsplit

    def barrier():
        lib.hommexx_b200_sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def call():
        h.held_suarez_forcing()
        return h.run_subcycle()

    for _ in range(args.warmup):
        call()
    barrier()
    sampler = bench.ClockSampler(local_rank)
    sampler.start()
    l0 = lib.hommexx_b200_launch_count()
    barrier()
    lib.hommexx_b200_event_record(0)
    for _ in range(args.calls):
        nstep = call()
    lib.hommexx_b200_event_record(1)
    barrier()
    ms = lib.hommexx_b200_event_elapsed_ms(0, 1)
    launches = lib.hommexx_b200_launch_count() - l0
    clocks = sampler.stop()
    free_b, total_b = torch.cuda.mem_get_info()
    used_gb = (total_b - free_b) / 1e9
    stats = torch.tensor([ms, used_gb], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    ms, used_gb = float(stats[0]), float(stats[1])
    # sanity of the forced state on this rank: finite, surface pressure in a physical range
    ps = h.get_field("ps_v")
    tt = h.get_field("t")
    ok = bool(np.isfinite(ps).all() and np.isfinite(tt).all() and ps.min() > 4e4 and ps.max() < 1.2e5 and tt.min() > 150.0)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        steps_per_s = dyn * args.calls / (ms * 1e-3)
        out = {"metric": "element_steps_per_s", "value": h.nelem * steps_per_s, "unit": "element-steps/s",
               "sypd": steps_per_s * cfg.tstep / 365.0, "n_gpus": world, "calls": args.calls, "warmup": args.warmup,
               "ms_per_call": ms / args.calls, "nstep": int(nstep), "simulated_days": nstep * cfg.tstep / 86400.0,
               "device_memory_gb_per_gpu_max": used_gb, "state_finite_and_physical": bool(int(flag.item())),
               "gpu_launches": int(launches), "clocks": clocks, "library": libpath.name, "dtype": "f64",
               "config": {"workload": f"Held-Suarez forced preqx ne{cfg.ne} ({h.nelem} elements) nlev{cfg.nlev} "
                                      f"qsize{cfg.qsize}, JW initial state, forcing evaluated on the device every call",
                          "namelist": f"homme-ne120-v1.nl scaled: tstep {cfg.tstep:g} nu {nu:.3g} rsplit {cfg.rsplit} "
                                      f"hypervis_subcycle {cfg.hypervis_subcycle} ftype 0",
                          "partition": f"HOMME space-filling curve, {world} parts, {h.nelemd} elements on rank 0"}}
        bench.emit(out)
    h.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
