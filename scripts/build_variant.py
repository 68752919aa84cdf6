#!/usr/bin/env python
"""Build an experimental variant of the CUDA dycore (same sources, extra -D flags) into
hommexx_b200/csrc/variants/<name>/, selected at run time with HXX_VARIANT=<name>.
  python scripts/build_variant.py <name> [-DFOO=1 ...] [--nlev 72 --qd 40]
Used for A/B timing of kernel parameters in one gpurun call."""
import pathlib
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402

name = sys.argv[1]
flags = [a for a in sys.argv[2:] if a.startswith("-D")]
if "--fmad=true" in sys.argv:
    g.NVCC_FLAGS.remove("--fmad=false")
nlev, qd = 72, 40
csrc = ROOT / "hommexx_b200" / "csrc"
out = csrc / "variants" / name
obj = csrc / "build" / f"variant_{name}"
out.mkdir(parents=True, exist_ok=True)
obj.mkdir(parents=True, exist_ok=True)
inc, lib = g.nccl_paths()


def cc(src):
    o = obj / (src.stem + ".o")
    cmd = ["nvcc", *g.NVCC_FLAGS, f"-DHXX_NLEV={nlev}", f"-DHXX_QSIZE_D={qd}", *flags, "-I", str(ROOT / "include"),
           "-DHXX_WITH_NCCL", "-I", str(inc), "-c", str(src), "-o", str(o)]
    r = subprocess.run(cmd, stdout=open(obj / (src.stem + ".log"), "w"), stderr=subprocess.STDOUT)
    assert r.returncode == 0, open(obj / (src.stem + ".log")).read()[-2000:]
    return o


with ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(cc, sorted(csrc.glob("*.cu"))))
subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o",
                str(out / f"libhommexx_b200_nlev{nlev}_q{qd}.so"), *map(str, objs), "-L", str(lib), "-l:libnccl.so.2",
                "-Xlinker", f"-rpath={lib}"], check=True)
print("built", out)
