"""Lane-level (32 lanes, NumPy) emulation of the FP64 tensor-core block sweep of the H-step kernel: the products that the
HALF_LAST instantiation leaves out multiply exact zeros, so its result is the full sweep's bit for bit and equals
-B^-1 (scripts/dmma_half_last_emulation.py; hstep_dmma.cu)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_half_last_sweep_is_identical_to_the_full_sweep(capsys):
    spec = importlib.util.spec_from_file_location("half_last", os.path.join(ROOT, "scripts", "dmma_half_last_emulation.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()                                  # asserts inside: zero operands, identical tiles, inverse to 1e-12
    out = capsys.readouterr().out
    assert "W=50 NB=7: DMMA 378 -> 351" in out
