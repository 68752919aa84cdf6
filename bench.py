#!/usr/bin/env python
"""Benchmark of the preqx dycore timestep (BASELINE.json: element-steps/s and SYPD at ne30,
nlev 72, qsize 40 on B200, with the fraction of the HBM roofline).

A "step" of this bench is ONE prim_run_subcycle_c call = rsplit*qsplit dynamics steps (3 for the
ne30 benchmark namelist test/reg_test/benchmarks/v1/homme-ne30-v1.nl) on synthetic
Jablonowski-Williamson baroclinic-wave initial data. `value` counts DYNAMICS steps:
element-steps/s = elements * dynamics steps / second, whole job.

  python bench.py [--gpus N --steps K --warmup W]         the CUDA dycore (one process per GPU;
                                                          for N > 1 launch with torch.distributed.run)
  python bench.py --impl reference ...                    the CPU arm: the reference's algorithm on
                                                          the host cores (oracle port; the reference's
                                                          own Kokkos/Fortran build needs toolchains
                                                          this image lacks)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import pathlib
import statistics
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

F_BYTES = 16 * 72 * 8  # one field tile of one element at nlev = 72


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# Native libraries (NCCL's version banner, NCCL_DEBUG output) write to file descriptor 1. The contract is
# ONE JSON line on stdout, so fd 1 is pointed at stderr for the whole run and the line goes to the
# saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload(args, n_gpus):
    """N=1: BASELINE configs[1] (ne30, homme-ne30-v1.nl). N>1: BASELINE configs[3], Jablonowski-Williamson at
    ne120 (86 400 elements, homme-ne120-v1.nl) SFC-partitioned over the N GPUs: STRONG scaling over 2/4/8.
    --weak restores the round-1 weak-scaling meshes (~5400 elements per GPU, ne = round(sqrt(900 N)))."""
    from hommexx_b200 import homme
    from oracle import oraclelib
    if args.ne:
        ne = args.ne
    elif n_gpus == 1:
        ne = 30
    elif args.weak:
        ne = int(round((900.0 * n_gpus) ** 0.5))
    else:
        ne = 120
    over = dict(ne=ne, npart=n_gpus)
    if ne == 120:  # BASELINE configs[3]: the reference's own ne120 namelist (homme-ne120-v1.nl)
        if args.qsize:
            over.update(qsize=args.qsize)
        return homme.preset("ne120", **over)
    if ne != 30:
        # the namelist's time step and hyperviscosity are tuned to ne30; other meshes follow HOMME's
        # usual scaling (tstep ~ 1/ne, nu ~ dx^3.2: homme-ne120-v1.nl has tstep 75, nu 1e13)
        nu = 1e15 * (30.0 / ne) ** 3.2
        over.update(tstep=300.0 * 30.0 / ne, nu=nu, nu_p=nu, nu_q=nu, nu_s=nu)
    if args.qsize:
        over.update(qsize=args.qsize)
    return homme.preset("ne30", **over)


def euler_advect_stage_bytes(nelem, qsize):
    """Algorithmic HBM bytes of one euler_advect launch of each of the 3 stages of a tracer step (DESIGN.md section
    4): per tracer read qdp + write qdp (+ read qtens_biharmonic on the hyperviscosity stage) + 2 qlim rows; per
    element read derived_dp, divdp_proj, divdp, vn0 (2) (+ dpdiss_biharmonic on stage 3) and read+write the DSS
    variable. Stage 2 also does that stage's min/max pass but reads the tracers once: the same bytes as stage 1."""
    return [nelem * ((2 + hv) * qsize * F_BYTES + qsize * 2 * 72 * 8 + (5 + hv + 2) * F_BYTES) for hv in (0, 0, 1)]


def euler_advect_bytes(nelem, qsize):
    return sum(euler_advect_stage_bytes(nelem, qsize)) / 3.0


def step_bytes_per_elem_step(cfg):
    """algorithmic tiles per element per dynamics step."""
    q, hv, rs, qs = cfg.qsize, cfg.hypervis_subcycle, cfg.rsplit, cfg.qsplit
    caar = 14 + 3 * 12 + 18 + 5 * 11 + 12
    hvt = hv * 54
    euler = (45 + 31 * q) / qs
    remap = (2 * (q + 3) + 6) / (rs * qs)
    updq = 2 * q / (rs * qs)
    misc = 10
    return (caar + hvt + euler + remap + updq + misc) * F_BYTES


def cpu_arm_library():
    """What bench.py's CPU arm times, best first:
    1. kind "reference": oracle/_ref/libref_hommexx_72_40_omp.so — the reference's own src/share/cxx sources in its
       production CPU configuration (AVX2 vector packs, OpenMP over elements), built by oracle/Makefile (ref_timing)
       against the Kokkos stand-in of oracle/ref_shim; prebuilt where /root/reference exists, it travels with the snapshot.
    2. kind "port": the oracle compiled ON THIS HOST with -O3 -march=native (make -C oracle native), else the strict
       parity build of the oracle."""
    from oracle import oraclelib
    ref = ROOT / "oracle" / "_ref" / "libref_hommexx_72_40_omp.so"
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        flags = ""
    if ref.exists() and " avx2" in flags and " fma" in flags and not os.environ.get("HXX_CPU_ARM_PORT"):
        return ref, "reference", ("the reference's own src/share/cxx sources (CaarFunctorImpl, EulerStepFunctorImpl, "
                                  "HyperviscosityFunctorImpl, RemapFunctor, BoundaryExchange, prim_driver ...) compiled "
                                  "with g++ -O3 -mavx2 -mfma -fopenmp, HOMMEXX_AVX_VERSION 2 (vector packs of 4 levels), "
                                  "against oracle/ref_shim's OpenMP Kokkos stand-in")
    out = ROOT / "oracle" / "_native" / "liboracle_native.so"
    try:
        subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), "native"], check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL, timeout=300)
        if out.exists():
            return out, "port", "oracle port, gcc -O3 -march=native -ffp-contract=fast -fopenmp (built on this host)"
    except (OSError, subprocess.SubprocessError):
        pass
    return oraclelib.ORACLE_LIB, "port", "oracle port, gcc -O3 -mavx2 -ffp-contract=off -fopenmp (strict parity build)"


def cpu_sample_mesh(args, cores):
    """The CPU arm runs the GPU arm's own configuration (ne30, 5400 elements) when the host has the cores to
    finish in minutes, else a smaller mesh of the same namelist (element-steps/s is per-element throughput)."""
    if args.ref_ne:
        return args.ref_ne
    return 30 if cores >= 12 else 15


def run_reference(args):
    """CPU arm: the oracle port of the reference's algorithm, all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)   # torchrun exports OMP_NUM_THREADS=1 to its workers
    os.environ.setdefault("OMP_PROC_BIND", "false")
    from hommexx_b200 import homme
    lib, kind, flags = cpu_arm_library()
    cfg = workload(args, 1)
    ne_s = cpu_sample_mesh(args, cores)
    cfg = workload(args, 1)
    scfg = homme.preset("ne30", ne=ne_s, qsize=cfg.qsize)
    h = homme.Homme(scfg, lib)
    h.init_dycore()
    dyn = scfg.rsplit * scfg.qsplit
    for _ in range(args.warmup):
        h.run_subcycle()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        h.run_subcycle()
    dt = time.perf_counter() - t0
    val = h.nelem * dyn * args.steps / dt
    sample = (f"{flags}; OpenMP over elements on "
              f"{cores} host threads, ne={ne_s} ({h.nelem} elements) nlev {scfg.nlev} qsize {scfg.qsize}, "
              f"homme-ne30-v1.nl namelist, {args.steps} prim_run_subcycle_c calls after {args.warmup} warm-up")
    h.close()
    out = {"impl": "reference", "metric": "element_steps_per_s", "value": val, "unit": "element-steps/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "sypd": None,
           "config": {"workload": f"preqx ne{cfg.ne} nlev{cfg.nlev} qsize{cfg.qsize}", "sample": sample,
                      "same_config": ne_s == cfg.ne},
           "cpu_baseline": {"value": val, "unit": "element-steps/s", "cores": cores, "kind": kind, "sample": sample,
                            "same_config": ne_s == cfg.ne},
           "e2e": {"value": val, "unit": "element-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


PINNED = []   # host pointers currently page-locked with cudaHostRegister


def pin_driver_arrays(h, torch):
    """Page-lock the driver's Fortran-layout arrays so the e2e copies run from pinned memory."""
    rt = torch.cuda.cudart()
    pinned = PINNED
    names = ("v", "T", "dp3d", "Qdp", "Q", "ps_v", "omega_p")
    if sum(h.array(n).nbytes for n in names) > (32 << 30):
        return pinned   # ne120 shards: tens of GB per rank; page-locking that much is slow and can fail
    for name in names:
        a = h.array(name)
        r = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
        if int(r) == 0:
            pinned.append(a.ctypes.data)
    return pinned


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--ne", type=int, default=0, help="override the mesh (default: 30 at N=1, weak-scaled for N>1)")
    ap.add_argument("--qsize", type=int, default=0)
    ap.add_argument("--ref-ne", type=int, default=0, help="mesh of the bounded CPU sample (default: ne30 itself)")
    ap.add_argument("--weak", action="store_true", help="N>1: weak-scaling meshes instead of ne120 strong scaling")
    ap.add_argument("--flavour", default="fma", choices=["fma", "strict"],
                    help="build that is benchmarked: fma = multiply-adds contracted (parity <= 1e-11, the default), "
                         "strict = --fmad=false (bit-identical to the oracle)")
    ap.add_argument("--no-other-build", "--no-fma", dest="no_other", action="store_true",
                    help="skip the timing of the other build beside the benchmarked one")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    from hommexx_b200 import homme
    from oracle import oraclelib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    n_gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the dycore has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    cfg = workload(args, n_gpus)
    cfg.part_id = rank
    flavour = "" if args.flavour == "strict" else "fma"
    if flavour and not homme.cuda_lib_path(cfg.nlev, cfg.qsize_d, flavour).exists():
        flavour = ""   # only the flagship (nlev, qsize_d) has an FMA build
    libpath = homme.cuda_lib_path(cfg.nlev, cfg.qsize_d, flavour)
    lib = homme.load_dycore(libpath)
    lib.hommexx_b200_event_elapsed_ms.restype = C.c_double
    lib.hommexx_b200_profile.argtypes = [C.c_ulonglong]
    lib.hommexx_b200_profile_read.restype = C.c_double
    lib.hommexx_b200_profile_read.argtypes = [C.c_int, C.POINTER(C.c_int64)]
    lib.hommexx_b200_kernel_id.argtypes = [C.c_char_p]
    lib.hommexx_b200_kernel_name.restype = C.c_char_p
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

    t_setup = time.perf_counter()
    h = homme.Homme(cfg, libpath)
    h.init_dycore()
    log(f"[rank {rank}] setup {time.perf_counter() - t_setup:.1f}s: ne={cfg.ne} nelem={h.nelem} local={h.nelemd} "
        f"qsize={cfg.qsize} nlev={cfg.nlev}")
    dyn = cfg.rsplit * cfg.qsplit

    def barrier():
        lib.hommexx_b200_sync()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    kid = lib.hommexx_b200_kernel_id(b"euler_advect")
    kids = [kid, lib.hommexx_b200_kernel_id(b"euler_advect_mm"), lib.hommexx_b200_kernel_id(b"euler_advect_hv")]
    for _ in range(args.warmup):
        h.run_subcycle()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    lib.hommexx_b200_profile(sum(1 << k for k in kids))   # CUDA-event pair around every launch of the dominant kernel
    l0 = lib.hommexx_b200_launch_count()
    barrier()
    lib.hommexx_b200_event_record(0)
    for _ in range(args.steps):
        h.run_subcycle()
    lib.hommexx_b200_event_record(1)
    barrier()
    ms = lib.hommexx_b200_event_elapsed_ms(0, 1)
    launches = lib.hommexx_b200_launch_count() - l0
    clocks = sampler.stop()
    nl = C.c_int64()
    stage_ms, stage_n = [], []
    for k in kids:   # the three stages of a tracer step are three variants of the kernel
        n_ = C.c_int64()
        stage_ms.append(lib.hommexx_b200_profile_read(k, C.byref(n_)))
        stage_n.append(n_.value)
    k_ms = sum(stage_ms)
    nl.value = sum(stage_n)
    lib.hommexx_b200_profile(0)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    nelem_global, nelem_local = h.nelem, h.nelemd
    value = h.nelem * dyn * args.steps / (ms * 1e-3)
    sypd = (dyn * args.steps / (ms * 1e-3)) * cfg.tstep / 365.0  # simulated years per wall day

    peak, peak_src = peaks()
    adv_bytes = euler_advect_bytes(h.nelemd, cfg.qsize)
    k_avg_ms = k_ms / max(1, nl.value)
    achieved = adv_bytes / (k_avg_ms * 1e-3) / 1e9 if k_avg_ms > 0 else 0.0
    traffic, fp64_pct = None, None
    tf = ROOT / "profiles" / "roofline_traffic.json"
    if tf.exists():
        prof = json.loads(tf.read_text())
        traffic = prof.get("euler_advect_dram_bytes_per_launch")
        fp64_pct = prof.get("euler_advect_fp64_pipe_pct")
    sb = euler_advect_stage_bytes(h.nelemd, cfg.qsize)
    per_stage = {}
    for name, b_, ms_, n_ in zip(("stage1 <HV=0,TAVG=0,MM=0>", "stage2 <0,0,1> (+ the stage's min/max pass)",
                                  "stage3 <1,1,0> (hyperviscosity + time average)"), sb, stage_ms, stage_n):
        if n_:
            gbs = b_ / (ms_ / n_ * 1e-3) / 1e9
            per_stage[name] = {"avg_launch_ms": ms_ / n_, "achieved": gbs, "frac": gbs / peak, "launches": n_}
    roofline = {"bound": "hbm", "kernel": "euler_advect_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "per_stage": per_stage,
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": adv_bytes, "avg_launch_ms": k_avg_ms, "launches_timed": nl.value,
                "kernel_share_of_step": k_ms / ms if ms > 0 else None,
                # the kernel's second ceiling (FP64 is uncontracted, --fmad=false): pipe utilisation of its three
                # stages from the committed ncu capture (profiles/r1f_ncu_summary.txt)
                "fp64_pipe_pct_ncu": fp64_pct}
    step_bytes = step_bytes_per_elem_step(cfg)
    step_gbs = value / n_gpus * step_bytes / 1e9

    # ---- per-kernel breakdown (untimed extra pass; CUDA events around every launch) ----------
    breakdown = None
    if not args.no_breakdown:
        nk = 0
        while lib.hommexx_b200_kernel_name(nk):
            nk += 1
        lib.hommexx_b200_profile((1 << nk) - 1)
        lib.hommexx_b200_event_record(2)
        h.run_subcycle()
        lib.hommexx_b200_event_record(3)
        tot = lib.hommexx_b200_event_elapsed_ms(2, 3)
        breakdown = {}
        for i in range(nk):
            n = C.c_int64()
            t = lib.hommexx_b200_profile_read(i, C.byref(n))
            if n.value:
                breakdown[lib.hommexx_b200_kernel_name(i).decode()] = {"launches": n.value, "ms": round(t, 4),
                                                                       "share": round(t / tot, 4)}
        breakdown["_subcycle_ms"] = round(tot, 4)
        lib.hommexx_b200_profile(0)
        if rank == 0:
            log("per-kernel breakdown of one prim_run_subcycle_c:")
            for k_, v_ in sorted(breakdown.items(), key=lambda kv: -kv[1]["ms"] if isinstance(kv[1], dict) else 0):
                log(f"  {k_:34s} {v_}")

    # ---- e2e: the reference-facing C ABI with HOST buffers, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        pin_driver_arrays(h, torch)
        st = h.state()
        h2d = sum(st[k].nbytes for k in ("v", "T", "dp3d", "Qdp", "ps_v"))
        d2h = sum(st[k].nbytes for k in ("v", "T", "dp3d", "Qdp", "Q", "ps_v", "omega_p"))
        h.push_results()            # warm the staging buffers
        barrier()
        t0 = time.perf_counter()
        h.upload_state()            # init_elements_states_c: host (Fortran layout) -> device, transposed
        for _ in range(args.steps):
            h.run_subcycle()        # prim_run_subcycle_c (host scalars in/out, device flag read back)
        h.push_results()            # cxx_push_results_to_f90: device -> host, all time levels
        barrier()
        e_s = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_s = float(t.item())
        e2e = {"value": h.nelem * dyn * args.steps / e_s, "unit": "element-steps/s",
               "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
               "what": "init_elements_states_c + steps x prim_run_subcycle_c + cxx_push_results_to_f90 from pinned "
                       "host arrays in the Fortran layout (the reference benchmark pushes once per run: statefreq=9999)"}

    # ---- the CAM-coupled pattern: forcing pushed and results pulled on EVERY call ----------------
    e2e_coupled = None
    if not args.no_e2e and n_gpus == 1:
        for name in ("FM", "FT", "FQ"):
            a = h.array(name)
            if int(torch.cuda.cudart().cudaHostRegister(a.ctypes.data, a.nbytes, 0)) == 0:
                PINNED.append(a.ctypes.data)
        fo = h.forcing()
        nrep = max(2, min(5, args.steps))
        h2d_c = sum(fo[k].nbytes for k in ("FM", "FT", "FQ"))
        d2h_c = d2h + st["Qdp"].nbytes
        h.push_forcing()            # first use allocates the device forcing arrays
        barrier()
        t0 = time.perf_counter()
        for _ in range(nrep):
            h.push_forcing()        # f90_push_forcing_to_cxx: FM, FT, FQ host -> device, Qdp device -> host
            h.run_subcycle()
            h.push_results()        # cxx_push_results_to_f90
        barrier()
        c_s = time.perf_counter() - t0
        e2e_coupled = {"value": h.nelem * dyn * nrep / c_s, "unit": "element-steps/s", "calls": nrep,
                       "h2d_bytes_per_step": h2d_c, "d2h_bytes_per_step": d2h_c,
                       "what": "per call: f90_push_forcing_to_cxx + prim_run_subcycle_c + cxx_push_results_to_f90 "
                               "(prim_driver_mod.F90:1322-1404, the pattern of a CAM-coupled run); PCIe-bound"}

    def unpin_all():
        # the driver frees its arrays on close: a registration left behind would poison the next allocation that
        # lands on the same virtual addresses (cudaMemcpyAsync: invalid argument)
        rt = torch.cuda.cudart()
        while PINNED:
            rt.cudaHostUnregister(PINNED.pop())

    # ---- the other build of the same sources, timed beside the benchmarked one (N = 1) --------------
    BUILD_NOTE = {"fma": "--fmad=true: multiply-adds contracted; matches the oracle to <= 1e-11 (measured ~1e-14) on v, T, "
                         "dp3d, ps, Qdp, Q after 10 steps (tests/test_cuda_q40.py::test_fma_build_*)",
                  "": "--fmad=false: no contraction; bit-identical to the oracle per phase and after 10-12 steps "
                      "(tests/test_cuda_q40.py, tests/test_cuda_timestep.py)"}
    other = None
    other_flav = "" if flavour else "fma"
    other_path = homme.cuda_lib_path(cfg.nlev, cfg.qsize_d, other_flav)
    if not args.no_other and n_gpus == 1 and other_path.exists() and other_path != libpath:
        unpin_all()
        h.close()
        h = None
        hf = homme.Homme(cfg, other_path)
        lf = hf.lib
        lf.hommexx_b200_event_elapsed_ms.restype = C.c_double
        lf.hommexx_b200_set_comm(0, 1, local_rank, None)
        hf.init_dycore()
        for _ in range(args.warmup):
            hf.run_subcycle()
        lf.hommexx_b200_sync()
        lf.hommexx_b200_event_record(0)
        for _ in range(args.steps):
            hf.run_subcycle()
        lf.hommexx_b200_event_record(1)
        lf.hommexx_b200_sync()
        fms = lf.hommexx_b200_event_elapsed_ms(0, 1)
        other = {"value": hf.nelem * dyn * args.steps / (fms * 1e-3), "unit": "element-steps/s",
                 "ms_per_step": fms / args.steps, "library": other_path.name, "what": BUILD_NOTE[other_flav]}
        hf.close()

    # ---- CPU baseline beside it (rank 0, N = 1) -------------------------------------------------
    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        os.environ["OMP_NUM_THREADS"] = str(cores)
        olib, kind, flags = cpu_arm_library()
        ne_s = cpu_sample_mesh(args, cores)
        scfg = homme.preset("ne30", ne=ne_s, qsize=cfg.qsize)
        ho = homme.Homme(scfg, olib)
        ho.init_dycore()
        ho.run_subcycle()
        t0 = time.perf_counter()
        nrep = 2
        for _ in range(nrep):
            ho.run_subcycle()
        dt = time.perf_counter() - t0
        cpu = {"value": ho.nelem * dyn * nrep / dt, "unit": "element-steps/s", "cores": cores, "kind": kind,
               "same_config": ne_s == cfg.ne,
               "sample": f"{flags}; {nrep} prim_run_subcycle_c calls at ne={ne_s} ({ho.nelem} elements), "
                         f"nlev {scfg.nlev}, qsize {scfg.qsize}, OpenMP over elements on {cores} host threads"}
        ho.close()

    if rank == 0:
        out = {"metric": "element_steps_per_s", "value": value, "unit": "element-steps/s", "n_gpus": n_gpus,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "weak" if (args.weak or n_gpus == 1) else "strong", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "sypd": sypd, "dynamics_steps_per_bench_step": dyn,
               "config": {"workload": f"preqx ne{cfg.ne} ({nelem_global} elements) nlev{cfg.nlev} qsize{cfg.qsize}, JW baroclinic "
                                      f"wave, homme-ne{120 if cfg.ne == 120 else 30}-v1.nl namelist (tstep {cfg.tstep:g} rsplit {cfg.rsplit} qsplit "
                                      f"{cfg.qsplit} hypervis_subcycle {cfg.hypervis_subcycle} limiter {cfg.limiter_option})",
                          "partition": f"SFC, {n_gpus} part(s), {nelem_local} elements on rank 0",
                          "l2": "working set (>9 GB at ne30) far exceeds the 126 MB L2; no flush needed",
                          "step": "one prim_run_subcycle_c call"},
               "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline,
               "step_roofline": {"algorithmic_bytes_per_element_step": step_bytes, "achieved_gbs_per_gpu": step_gbs,
                                 "frac": step_gbs / peak, "peak": peak},
               "e2e": e2e, "e2e_coupled": e2e_coupled, "build": {"library": libpath.name, "what": BUILD_NOTE[flavour]},
               ("strict_build" if flavour else "fma_build"): other, "cpu_baseline": cpu, "breakdown": breakdown}
        emit(out)
    if h is not None:
        unpin_all()
        h.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
