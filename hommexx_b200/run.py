"""Namelist-driven run of the preqx dycore: the stand-in for `prim_main` on this side of the C ABI.

    python -m hommexx_b200.run hommexx_b200/namelists/prtcA-r3-dry.nl [--nmax N] [--held-suarez]

Reads a HOMME `ctl_nl` / `vert_nl` namelist (the reference's own .nl files work: unknown keys are ignored),
builds the mesh and the Jablonowski-Williamson state with the C++ driver, steps `prim_run_subcycle_c` until
`nmax` dynamics steps are done (prim_main.F90:285-331) and, every `statefreq` steps, pushes the results and
prints a `prim_printstate`-style block: min / max of the prognostic fields at the current time level and
the global integrals of the `elem%accum` energies and tracer masses the dycore's diagnostics wrote
(Diagnostics.cpp:37-185; global_integral = sum(spheremp * f) / (4 pi), prim_state_mod.F90).
Restarts follow prim_main.F90:181-262,355: `restartfreq` > 0 (steps) writes `restartdir/R<nstep>.npz` (elem%state and
the time levels, WriteRestart), `runtype = 1` with `restartfile` resumes from one; the resumed run is bit-identical
to the uninterrupted one. Host orchestration only; the dycore is the CUDA library (no CPU fallback).
"""
from __future__ import annotations

import argparse
import re
import sys

import numpy as np

from . import homme


def parse_namelist(text: str) -> dict:
    """Fortran namelist groups -> {group: {key: value}}; comments (!), .true./.false., quoted strings, numbers."""
    groups: dict = {}
    cur = None
    for raw in text.splitlines():
        line = raw.split("!", 1)[0].strip()
        if not line:
            continue
        if line.startswith("&"):
            cur = groups.setdefault(line[1:].strip().lower(), {})
            continue
        if line.startswith("/"):
            cur = None
            continue
        if cur is None or "=" not in line:
            continue
        for item in re.split(r",(?![^'\"]*['\"])", line):
            if "=" not in item:
                continue
            k, v = (x.strip() for x in item.split("=", 1))
            cur[k.lower()] = _value(v)
    return groups


def _value(v: str):
    v = v.strip().rstrip(",")
    if v.lower() in (".true.", "t", ".t."):
        return True
    if v.lower() in (".false.", "f", ".f."):
        return False
    if v[:1] in "'\"":
        return v.strip("'\"")
    try:
        return int(v)
    except ValueError:
        pass
    try:
        return float(v.lower().replace("d", "e"))
    except ValueError:
        return v


def config_from_namelist(nl: dict, **over) -> homme.Config:
    """ctl_nl -> Config, with namelist_mod.F90's names (se_ftype, vert_remap_q_alg, statefreq, moisture='dry'|...)."""
    c = nl.get("ctl_nl", {})
    v = nl.get("vert_nl", {})
    vfile = str(v.get("vfile_int", ""))
    nlev = 26 if "26" in vfile else 72
    qsize = int(c.get("qsize", 4))
    kw = dict(ne=int(c.get("ne", 4)), nlev=nlev, qsize=qsize, qsize_d=40 if (nlev == 72 and qsize > 4) else 4)
    names = dict(vert_remap_q_alg="remap_alg", limiter_option="limiter_option", rsplit="rsplit", qsplit="qsplit",
                 tstep_type="time_step_type", energy_fixer="energy_fixer", statefreq="state_frequency", nu="nu",
                 nu_p="nu_p", nu_q="nu_q", nu_s="nu_s", nu_div="nu_div", nu_top="nu_top",
                 hypervis_order="hypervis_order", hypervis_subcycle="hypervis_subcycle",
                 hypervis_scaling="hypervis_scaling", se_ftype="ftype", tstep="tstep", u_perturb="u_perturb")
    for k, dst in names.items():
        if k in c:
            kw[dst] = type(getattr(homme.Config(), dst))(c[k])
    if "moisture" in c:
        kw["moisture"] = 0 if str(c["moisture"]).lower() == "dry" else 1
    if "disable_diagnostics" in c:
        kw["disable_diagnostics"] = int(bool(c["disable_diagnostics"]))
    if str(c.get("test_case", "jw_baroclinic")).lower() not in ("jw_baroclinic", "baroclinic"):
        raise SystemExit(f"test_case {c['test_case']!r}: only the Jablonowski-Williamson case is built into the driver")
    kw.update(over)
    return homme.Config(**kw)


def global_integral(h, f):
    """sum(spheremp f) / (4 pi) over this rank's elements; f [n, ..., 4, 4]."""
    sph = h.array("spheremp").reshape(h.nelemd, *([1] * (f.ndim - 3)), 4, 4)
    return (f * sph).sum(axis=(0, f.ndim - 2, f.ndim - 1)) / (4.0 * np.pi)


def printstate(h, out=sys.stdout):
    """min / max at the current time level and the integrals of the accumulated diagnostics."""
    nstep, nm1, n0, np1 = h.time_levels()
    st = h.state()
    g = 9.80616
    print(f" nstep= {nstep}  time= {nstep * h.cfg.tstep / 86400.0:.6f} [day]", file=out)
    for nm, a in (("u", st["v"][:, n0 - 1, :, 0]), ("v", st["v"][:, n0 - 1, :, 1]), ("T", st["T"][:, n0 - 1]),
                  ("dp3d", st["dp3d"][:, n0 - 1]), ("ps", st["ps_v"][:, n0 - 1])):
        print(f" {nm:5s}= {a.min():24.15e} {a.max():24.15e} {a.sum():24.15e}", file=out)
    res = {"nstep": nstep}
    if not h.cfg.disable_diagnostics:
        acc = h.accum()
        for nm in ("KEner", "IEner", "PEner"):
            res[nm] = float(global_integral(h, acc[nm][:, 1]) / g)          # ivar 1 = after the last advance
            print(f" {nm:5s}= {res[nm]:24.15e}  [J/m^2]", file=out)
        res["TOTE"] = res["KEner"] + res["IEner"] + res["PEner"]
        print(f" TOTE = {res['TOTE']:24.15e}  [J/m^2]", file=out)
        qm = global_integral(h, acc["Qmass"][:, 1, :h.cfg.qsize]) / g
        for q, m in enumerate(np.atleast_1d(qm)[:4]):
            print(f" Q{q + 1:<2d} mass = {m:22.15e}  [kg/m^2]", file=out)
        res["Qmass"] = [float(x) for x in np.atleast_1d(qm)]
    return res


def run(cfg: homme.Config, libpath, nmax: int, held_suarez: bool = False, out=sys.stdout, restart_in=None,
        restartfreq: int = 0, restartdir: str = "."):
    h = homme.Homme(cfg, libpath)
    h.set_last_step(nmax)  # nEndStep: switches the last step's diagnostics on (prim_driver.cpp:55-57)
    if restart_in:         # runtype = 1: elem%state and tl come from the file, then the usual initialisation
        h.read_restart(restart_in)
    h.init_dycore()
    history = []
    nstep = h.time_levels()[0]
    while nstep < nmax:
        if held_suarez:
            from . import held_suarez as hs
            nstep = hs.forced_step(h)
        else:
            nstep = h.run_subcycle()
        if nstep % cfg.state_frequency == 0 or nstep >= nmax:  # prim_main.F90:312-321
            h.push_results()
            history.append(printstate(h, out))
        if restartfreq > 0 and nstep % restartfreq == 0:                       # prim_main.F90:355
            import os
            os.makedirs(restartdir, exist_ok=True)
            h.write_restart(os.path.join(restartdir, f"R{nstep:09d}.npz"))
    h.close()
    return history


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("namelist")
    ap.add_argument("--nmax", type=int, default=0, help="dynamics steps (default: the namelist's nmax, else 12)")
    ap.add_argument("--held-suarez", action="store_true", help="Held-Suarez forcing through f90_push_forcing_to_cxx")
    args = ap.parse_args(argv)
    nl = parse_namelist(open(args.namelist).read())
    cfg = config_from_namelist(nl)
    ctl = nl.get("ctl_nl", {})
    nmax = args.nmax or int(ctl.get("nmax", 12))
    restart_in = str(ctl["restartfile"]) if int(ctl.get("runtype", 0)) == 1 else None
    # the product has one backend: the CUDA library (load_dycore raises if it is not built; no CPU fallback)
    run(cfg, homme.cuda_lib_path(cfg.nlev, cfg.qsize_d), nmax, args.held_suarez, restart_in=restart_in,
        restartfreq=max(0, int(float(ctl.get("restartfreq", 0)) * 86400.0 / cfg.tstep)),  # days -> steps, namelist_mod.F90:482
        restartdir=str(ctl.get("restartdir", "./restart")))


if __name__ == "__main__":
    main()
