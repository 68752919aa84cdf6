#!/usr/bin/env python
"""Run a few prim_run_subcycle_c calls of the CUDA dycore and nothing else (the target of ncu captures).
  python scripts/prof_step.py [--ne 30] [--qsize 40] [--steps 1]"""
import argparse
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from hommexx_b200 import homme  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ne", type=int, default=30)
ap.add_argument("--qsize", type=int, default=40)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--flavour", default="", help='"" = strict build, "fma" = the FMA-contracted build')
a = ap.parse_args()
cfg = homme.preset("ne30", ne=a.ne, qsize=a.qsize)
h = homme.Homme(cfg, homme.cuda_lib_path(cfg.nlev, cfg.qsize_d, a.flavour))
h.init_dycore()
for _ in range(a.steps):
    h.run_subcycle()
h.lib.hommexx_b200_sync()
h.close()
