"""(reference build, library under test) sessions side by side for the per-functor parity tests: both driven through
the phase-level hooks of include/hommexx_b200.h section C from one common state. TEST INFRASTRUCTURE."""
import numpy as np

import distinct_tracers
from hommexx_b200 import homme
from reference_lib import reference_lib

# every array both libraries name (the oracle's dp_star has no counterpart View in the reference)
FIELDS = ["v", "t", "dp3d", "ps_v", "phi", "omega_p", "eta_dot_dpdn", "derived_vn0", "derived_dp", "divdp", "divdp_proj",
          "dpdiss_ave", "dpdiss_biharmonic", "qdp", "qtens_biharmonic", "qlim", "Q", "vtens", "ttens", "dptens"]


class Pair:
    """(reference, other library) sessions of one configuration, resettable to a common non-trivial state."""

    def __init__(self, cfg, libpath, backend, fields=None, warm_calls=1):
        self.cfg = cfg
        self.fields = list(fields or FIELDS)
        self.hr = homme.Homme(cfg, reference_lib(cfg.nlev, cfg.qsize_d))
        self.ho = homme.Homme(cfg, libpath)
        for h in (self.hr, self.ho):
            if cfg.qsize > 4:
                distinct_tracers.install(h)
            h.init_dycore()
        assert self.hr.lib.hommexx_b200_backend() == b"reference-serial"
        assert self.ho.lib.hommexx_b200_backend() == backend
        for _ in range(warm_calls):
            self.hr.run_subcycle()
        self.snap = {n: self.hr.get_field(n) for n in self.fields}
        for n, a in self.snap.items():
            assert a.size == self.ho.field_size(n), n
            assert np.isfinite(a).all(), n

    def reset(self):
        for n, a in self.snap.items():
            self.hr.set_field(n, a)
            self.ho.set_field(n, a)

    def call(self, name, *args):
        for h in (self.hr, self.ho):
            getattr(h.lib, name)(*args)

    def same(self, what, names=None, changed=(), skip=()):
        for n in names or self.fields:
            if n in skip:
                continue
            a, b = self.hr.get_field(n), self.ho.get_field(n)
            assert not np.isnan(a).any(), (what, n)
            assert np.array_equal(a, b), (what, n, float(np.abs(a - b).max()))
        for n in changed:  # the phase really wrote what it is supposed to write
            assert not np.array_equal(self.hr.get_field(n), self.snap[n]), (what, n, "unchanged")

    def close(self):
        self.hr.close()
        self.ho.close()
