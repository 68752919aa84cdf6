"""Import the reference's known-answer vectors for the sphere operators into tests/golden/.

Source: /root/reference/test/unit_tests/inputs/{gradient,divergence,vorticity}_sphere_np4.in
(text blocks "name\n values\n"; Fortran column-major arrays, e.g. elem_Dinv(np,np,2,2)).
They are produced by the Fortran operators of src/share/derivative_mod_base.F90 and are not
referenced by any reference test, so they are free golden vectors.
Run once in the build container; tests/golden/sphere_kats.json is committed.
"""
import json, pathlib

ref = pathlib.Path("/root/reference/test/unit_tests/inputs")
out = pathlib.Path(__file__).resolve().parents[1] / "tests" / "golden" / "sphere_kats.json"
kats = {}
for op in ("gradient", "divergence", "vorticity"):
    lines = (ref / f"{op}_sphere_np4.in").read_text().split("\n")
    blocks, name = {}, None
    for ln in lines[1:]:
        s = ln.strip()
        if not s:
            continue
        try:
            vals = [float(x.replace("D", "E")) for x in s.split()]
            blocks.setdefault(name, []).extend(vals)
        except ValueError:
            name = s
    blocks = {k.replace(" ", "_"): v for k, v in blocks.items()}
    assert blocks["np"] == [4.0]
    kats[op] = blocks
    print(op, {k: len(v) for k, v in blocks.items()})
out.write_text(json.dumps(kats))
