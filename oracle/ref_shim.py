"""Import the *unmodified* reference (catniplab/vlgp at /root/reference) so golden vectors can be made from it.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` and by CPU tests that run in the build container;
/root/reference does not exist on the GPU box, so nothing that runs there may import this module.

The reference does not run on SciPy >= 1.14 because ``scipy.linalg.solve`` lost its ``sym_pos`` keyword
(call sites: /root/reference/vlgp/core.py:89,110,193,211,226,230,465).  The shim rebinds the *name* ``solve`` inside
``vlgp.core`` to a wrapper that maps ``sym_pos=True`` to ``assume_a="pos"`` (same LAPACK ``posv`` driver).  The reference
source tree is not touched.
"""
import os
import sys
import tempfile

REFERENCE_ROOT = os.environ.get("VLGP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "vlgp"))


def load():
    """Return the reference ``vlgp`` package with the sym_pos shim installed."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # ``import vlgp`` opens ./vlgp.log (vlgp/__init__.py:7-12): import from a scratch directory.
    cwd = os.getcwd()
    scratch = tempfile.mkdtemp(prefix="vlgp_ref_")
    os.chdir(scratch)
    try:
        import vlgp  # noqa: F401
        import vlgp.core
        import scipy.linalg

        if not getattr(vlgp.core.solve, "_vlgp_shim", False):
            def solve(a, b, sym_pos=False, **kw):
                return scipy.linalg.solve(a, b, assume_a="pos" if sym_pos else "gen", **kw)

            solve._vlgp_shim = True
            vlgp.core.solve = solve
    finally:
        os.chdir(cwd)
    return vlgp
