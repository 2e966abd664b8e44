"""Run the reference's own ``model/point_net2.py`` and ``model/project_to_2d.py`` VERBATIM on top of
the restated third-party ops (oracle/thirdparty_ops.py).  TEST INFRASTRUCTURE ONLY.

The files are loaded from (first hit): $SN2_REFERENCE_ROOT, /root/reference, oracle/_ref (git-ignored
staging written by oracle/stage_ref.py; it travels to the GPU box, /root/reference does not).
Nothing from the reference is copied into tracked files.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

from . import thirdparty_ops as tp

_HERE = os.path.dirname(os.path.abspath(__file__))
_CACHE = {}


def reference_root() -> str | None:
    for cand in (os.environ.get("SN2_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "model", "point_net2.py")) and os.path.isfile(
            os.path.join(cand, "model", "project_to_2d.py")
        ):
            return cand
    return None


def _stub_modules() -> dict:
    tg = types.ModuleType("torch_geometric")
    tgnn = types.ModuleType("torch_geometric.nn")
    for name in ("knn_interpolate", "PointConv", "fps", "radius", "global_max_pool", "knn"):
        setattr(tgnn, name, getattr(tp, name))
    tg.nn = tgnn
    ts = types.ModuleType("torch_scatter")
    for name in ("scatter_max", "scatter_mean", "scatter_add"):
        setattr(ts, name, getattr(tp, name))
    ut = types.ModuleType("utils")
    utu = types.ModuleType("utils.utils")

    def get_trained_model_path_from_experiment(path, experiment_id):  # unused import at point_net2.py:6
        raise NotImplementedError("oracle stub")

    utu.get_trained_model_path_from_experiment = get_trained_model_path_from_experiment
    ut.utils = utu
    return {
        "torch_geometric": tg,
        "torch_geometric.nn": tgnn,
        "torch_scatter": ts,
        "utils": ut,
        "utils.utils": utu,
    }


def load_reference():
    """-> (point_net2 module, project_to_2d module) or raises FileNotFoundError."""
    root = reference_root()
    if root is None:
        raise FileNotFoundError("reference model files not found (no /root/reference, no oracle/_ref)")
    if root in _CACHE:
        return _CACHE[root]
    stubs = _stub_modules()
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        mods = []
        for fname, modname in (("point_net2.py", "sn2_reference_point_net2"), ("project_to_2d.py", "sn2_reference_project_to_2d")):
            spec = importlib.util.spec_from_file_location(modname, os.path.join(root, "model", fname))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mods.append(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _CACHE[root] = tuple(mods)
    return _CACHE[root]
