"""TEST INFRASTRUCTURE ONLY -- import the reference's OWN box-op source files under the numpy `megengine` shim.

`load()` returns a namespace with the reference modules
    structures.{boxes, boxcoder, box_convert, container, op_patch}
    layers.common.{anchor_generator, matcher, post_processing, roi_pool, function}
executed from `/root/reference/basedet/...` (read-only, never copied).  Only available in the build container;
`available()` is False on the GPU box, where the committed golden vectors (tests/golden/*.npz) are used instead.
"""
import functools
import importlib
import os
import sys
import types

REFERENCE = os.environ.get("BASEDET_REFERENCE", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mge_shim")
_loaded = None


def available():
    return os.path.isdir(os.path.join(REFERENCE, "basedet", "structures"))


def _pkg(name, path=None):
    m = types.ModuleType(name)
    m.__path__ = [path] if path else []
    sys.modules[name] = m
    return m


def load():
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE)
    if "megengine" in sys.modules and not getattr(sys.modules["megengine"], "__file__", "").startswith(_SHIM):
        raise RuntimeError("a real megengine is already imported")
    sys.path.insert(0, _SHIM)
    try:
        import megengine  # noqa: F401  (the shim)

        root = os.path.join(REFERENCE, "basedet")
        # Stub packages: only the box-op files are executed, never the heavy package __init__s
        # (basedet/layers/__init__.py pulls in basecore, backbones, losses ...).
        _pkg("basedet", root)
        utils = _pkg("basedet.utils")
        utils.cached_property = functools.cached_property
        _pkg("basedet.layers", os.path.join(root, "layers"))
        _pkg("basedet.layers.common", os.path.join(root, "layers", "common"))
        blocks = _pkg("basedet.layers.blocks")
        blocks.SinkhornDistance = type("SinkhornDistance", (), {})
        losses = _pkg("basedet.layers.losses")
        losses.iou_loss = None
        ns = types.SimpleNamespace()
        ns.structures = importlib.import_module("basedet.structures")  # real __init__: numpy + megengine only
        for mod in ("boxes", "boxcoder", "box_convert", "container", "op_patch"):
            setattr(ns, mod, importlib.import_module("basedet.structures." + mod))
        for mod in ("function", "anchor_generator", "matcher", "post_processing", "roi_pool"):
            setattr(ns, mod, importlib.import_module("basedet.layers.common." + mod))
        ns.Tensor = megengine.Tensor
        ns.F = megengine.functional
        _loaded = ns
        return ns
    finally:
        sys.path.remove(_SHIM)


def load_method(rel_path, class_name, method_name, extra_globals=None):
    """AST-extract ONE method of a reference class (e.g. models/det/fcos.py: FCOS.get_ground_truth) and compile it
    against the shim -- the model modules themselves cannot be imported (basecore, backbones ...).  The returned
    function takes ``self`` explicitly; tests pass a SimpleNamespace carrying the few attributes the method reads."""
    import ast

    import numpy as np

    ref = load()
    path = os.path.join(REFERENCE, "basedet", rel_path)
    tree = ast.parse(open(path).read(), path)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == class_name:
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name == method_name:
                    mod = ast.Module(body=[fn], type_ignores=[])
                    sys.path.insert(0, _SHIM)
                    try:
                        import megengine as mge
                    finally:
                        sys.path.remove(_SHIM)
                    glb = {"F": ref.F, "mge": mge, "np": np, "Boxes": ref.boxes.Boxes, "layers": None}
                    glb.update(extra_globals or {})
                    exec(compile(mod, path, "exec"), glb)
                    return glb[method_name]
    raise KeyError("%s.%s not found in %s" % (class_name, method_name, path))
