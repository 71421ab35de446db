"""CPU-only: the C-ABI library loads and exports every symbol include/bdet.h declares; the ctypes table mirrors
the header; the package refuses to work without the library (no fallback).  No compute calls (no GPU here)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bdet.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = re.findall(r"\b(?:int|size_t|const char\s*\*)\s+(bdet_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S)
    return {name: [a.strip() for a in args.split(",") if a.strip() and a.strip() != "void"] for name, args in decls}


def test_header_declares_the_hot_path():
    fns = header_functions()
    for required in ("bdet_anchors_grid", "bdet_points_grid", "bdet_pairwise", "bdet_pairwise_batched", "bdet_match",
                     "bdet_match_rows", "bdet_box_encode", "bdet_box_decode", "bdet_point_encode", "bdet_point_decode",
                     "bdet_assign_targets", "bdet_topk", "bdet_score_filter_topk", "bdet_nms", "bdet_roi_assign_levels",
                     "bdet_roi_align_fwd", "bdet_roi_align_bwd", "bdet_boxes_scale_clip", "bdet_cond_take"):
        assert required in fns, required


def test_library_exports_every_declared_symbol():
    from basedet_b200 import _lib, build

    path = build.build()
    lib = _lib.load()
    fns = header_functions()
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (bdet_\w+)", out))
    assert set(fns) <= exported, sorted(set(fns) - exported)
    assert exported <= set(fns), "exported but undeclared: %s" % sorted(exported - set(fns))
    for name in fns:
        assert hasattr(lib, name)
    assert lib.bdet_abi_version() == 1


def test_ctypes_table_mirrors_header():
    from basedet_b200 import _lib

    fns = header_functions()
    assert set(_lib.SIGNATURES) == set(fns), set(_lib.SIGNATURES) ^ set(fns)
    for name, args in fns.items():
        assert len(_lib.SIGNATURES[name][1]) == len(args), (name, len(_lib.SIGNATURES[name][1]), args)


def test_signatures_are_plain_c():
    src = open(HEADER).read()
    assert "torch" not in src and "at::" not in src and "std::" not in src
    assert 'extern "C"' in src


def test_host_argument_errors_without_gpu():
    """Argument validation mirrors the reference's Python asserts and happens before any launch."""

    from basedet_b200 import _lib

    lib = _lib.load()
    rc = lib.bdet_pairwise(None, 3, 1, None, 4, 1, None, 1, 0, None)  # boxes with < 4 columns
    assert rc == -1 and b"4 columns" in lib.bdet_last_error()
    thr = _lib.farr([0.5, 0.4])
    labs = _lib.iarr([0, -1, 1])
    rc = lib.bdet_match(None, 8, 0, None, 1, 8, 1, thr, labs, 3, 0, None, None, None, 0, None)
    assert rc == -1 and b"sorted" in lib.bdet_last_error()  # matcher.py:23
    assert lib.bdet_match_workspace(100, 120087, 16) > 0
    assert lib.bdet_nms_workspace(5000, 1) > 5000 * 16 * 8  # sorted boxes + kept list + chunk-local mask words


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from basedet_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under basedet_b200/ may import it."""
    pkg = os.path.join(ROOT, "basedet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, f)


def test_segment_descriptors_from_views_of_one_allocation():
    """ops._segments (host logic of the batched top-k / filter calls): (image, level) segments that live in several
    tensors are described as element offsets from the lowest base pointer, image-major."""
    import torch

    from basedet_b200 import ops

    store = torch.zeros(2 * (6 + 10), dtype=torch.float32)
    lvl0 = store[:12].view(2, 6)           # level 0: 2 images x 6
    lvl1 = store[12:].view(2, 10)          # level 1: 2 images x 10
    base, starts, lens = ops._segments([lvl1, lvl0])   # order of the list = level order, not address order
    assert base.data_ptr() == store.data_ptr()
    assert lens == [10, 6, 10, 6]
    assert starts == [12, 0, 22, 6]
