"""Convert the reference's vertical-coordinate tables (input DATA, not source) into
the flat format the harness reads: one file per grid holding hyai, hybi, hyam, hybm.

Source tables: /root/reference/test/vcoord/{acme-72i,acme-72m,cami-26,camm-26}.ascii
(format per src/share/hybvcoord_mod.F90:82-111: "<n> ! name" header then n values, twice).
Run once in the build container; the outputs under hommexx_b200/data/ are committed.
"""
import re, sys, pathlib

def read_blocks(path):
    toks = pathlib.Path(path).read_text().split("\n")
    vals, blocks, n = [], [], None
    for line in toks:
        if "!" in line:
            if n is not None:
                assert len(vals) == n, (path, n, len(vals))
                blocks.append(vals)
            n = int(line.split("!")[0].split()[0]); vals = []
        else:
            vals += [float(x) for x in line.split()]
    assert len(vals) == n
    blocks.append(vals)
    return blocks

ref = pathlib.Path("/root/reference/test/vcoord")
out = pathlib.Path(__file__).resolve().parents[1] / "hommexx_b200" / "data"
for name, fi, fm in (("acme-72", "acme-72i.ascii", "acme-72m.ascii"), ("cam-26", "cami-26.ascii", "camm-26.ascii")):
    hyai, hybi = read_blocks(ref / fi)
    hyam, hybm = read_blocks(ref / fm)
    assert len(hyai) == len(hybi) == len(hyam) + 1 == len(hybm) + 1
    with open(out / f"vcoord-{name}.txt", "w") as f:
        f.write(f"# nlev={len(hyam)}; rows: hyai hybi (nlev+1 lines) then hyam hybm (nlev lines)\n")
        for a, b in zip(hyai, hybi): f.write(f"{a!r} {b!r}\n")
        for a, b in zip(hyam, hybm): f.write(f"{a!r} {b!r}\n")
    print(name, len(hyam))
