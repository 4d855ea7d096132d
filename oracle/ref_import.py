"""
Import the UNMODIFIED reference (/root/reference/pyfem.py + utils.py) in the build
container -- TEST INFRASTRUCTURE.  Only oracle/make_golden.py and the optional
`-m "not gpu"` cross-check use it; /root/reference does not exist on the GPU box,
so nothing that runs there may call this.

The reference imports matplotlib and pyamg at module level (pyfem.py:6-8); neither
is on the assembly path and neither is installed here, so they are stubbed.  Its
profiler decorator samples the on/off flag at decoration time and writes
profiler.log into cwd (utils.py:61,119-127), so timer_off() runs before
`import pyfem`.
"""
import os
import sys
import types

REF_DIR = os.environ.get("PYFEM_REFERENCE_DIR", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_DIR, "pyfem.py"))


class _SpsolveAMG:
    """Stand-in for pyamg.ruge_stuben_solver(K): Helmholtz.__init__ builds one
    (pyfem.py:2098) although the assembly path never uses it."""

    def __init__(self, K):
        self.K = K

    def solve(self, b, tol=1e-8):
        from scipy.sparse.linalg import spsolve
        return spsolve(self.K.tocsc(), b)


def load():
    """Return the reference `pyfem` module (cached)."""
    if "pyfem" in sys.modules and getattr(sys.modules["pyfem"], "__file__", "").startswith(REF_DIR):
        return sys.modules["pyfem"]
    if not available():
        raise RuntimeError(f"reference not found at {REF_DIR}")
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.tri", "matplotlib.pylab"):
        sys.modules.setdefault(name, types.ModuleType(name))
    amg = types.ModuleType("pyamg")
    amg.ruge_stuben_solver = _SpsolveAMG
    amg.smoothed_aggregation_solver = _SpsolveAMG
    sys.modules.setdefault("pyamg", amg)
    sys.path.insert(0, REF_DIR)
    try:
        import utils as ref_utils  # the reference's utils.py
        ref_utils.timer_off()
        import pyfem as ref_pyfem
    finally:
        sys.path.remove(REF_DIR)
    return ref_pyfem
