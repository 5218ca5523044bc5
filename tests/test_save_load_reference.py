"""save / load against the REFERENCE's own functions (vlgp/util.py:181-208), in the build container only: the file this
package writes is read by the reference's loader and the other way round.  The reference
tree does not exist on the GPU box, so the test skips there; nothing here touches the device."""
import os
import sys

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "vlgp")), reason="reference tree not present")


def _result():
    rng = np.random.default_rng(0)
    trials = [dict(y=rng.poisson(0.3, (40, 6)).astype(float), mu=rng.standard_normal((40, 2)),
                   v=rng.random((40, 2)), w=rng.random((40, 2)), x=np.ones((40, 1, 6))) for _ in range(3)]
    params = dict(a=rng.standard_normal((2, 6)), b=rng.standard_normal((1, 6)), noise=rng.random(6), omega=np.array([1e-3, 5e-3]),
                  sigma=np.ones(2), zdim=2, ydim=6, xdim=1, rank=50, gp_noise=1e-4, dt=1,
                  likelihood=np.array(["poisson"] * 6), transform=np.mean)
    return dict(trials=trials, params=params, config=dict(max_iter=3, window=20, callbacks=[]))


def _same(a, b):
    assert a["params"].keys() - {"transform"} == b["params"].keys() - {"transform"}
    for k in ("a", "b", "noise", "omega", "sigma"):
        assert np.array_equal(a["params"][k], b["params"][k])
    assert len(a["trials"]) == len(b["trials"])
    for ta, tb in zip(a["trials"], b["trials"]):
        for k in ("y", "mu", "v", "w"):
            assert np.array_equal(ta[k], tb[k])
    assert a["config"]["max_iter"] == b["config"]["max_iter"]


def test_files_cross_load_between_this_package_and_the_reference(tmp_path, monkeypatch):
    sys.path.insert(0, REF)
    try:
        import importlib

        ref_util = importlib.import_module("vlgp.util")
    except Exception as e:  # pragma: no cover - the reference's imports are not under test
        pytest.skip("reference util not importable here: %r" % (e,))
    finally:
        sys.path.remove(REF)
    from vlgp_b200 import util

    res = _result()
    # ours -> reference.  The reference's loader calls np.load without allow_pickle, which current NumPy refuses for any
    # pickled dict -- its own files included -- so the loader is run with that one default restored (the reference
    # source is untouched), exactly what a user of the reference on this NumPy has to do.
    util.save(res, tmp_path / "ours", ext="npy")
    real_load = np.load
    monkeypatch.setattr(ref_util.np, "load", lambda p, *a, **k: real_load(p, *a, **{**k, "allow_pickle": True}))
    back = ref_util.load(tmp_path / "ours.npy")
    _same(res, back)
    # reference -> ours (the reference pickles whatever it is given, the sklearn bound method included when it can)
    ref_res = {k: v for k, v in res.items()}
    ref_res["params"] = {k: v for k, v in res["params"].items() if k != "transform"}
    ref_util.save(ref_res, tmp_path / "theirs", ext="npy")
    monkeypatch.undo()
    _same(res, util.load(tmp_path / "theirs.npy"))
    with pytest.raises(FileNotFoundError):
        util.load(tmp_path / "missing.npy")
