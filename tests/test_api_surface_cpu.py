"""Drop-in check of the Python boundary: every public method of the reference's ``Engine``, the
``UNetModel`` constructor, ``get_model`` and the ``nn.py`` operator seam exist here with the same parameter names in
the same order (extra OPTIONAL parameters are allowed).  The reference is parsed with ``ast`` (no import, so its
Lightning / wandb dependencies are not needed); skipped where /root/reference is absent (the GPU box)."""
import ast
import inspect
import os

import pytest

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present")


def _functions(path, cls=None):
    tree = ast.parse(open(path).read())
    body = tree.body
    if cls is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    out = {}
    for n in body:
        if isinstance(n, ast.FunctionDef):
            a = n.args
            names = [x.arg for x in a.posonlyargs + a.args]
            n_required = len(names) - len(a.defaults)
            out[n.name] = (names, n_required)
    return out


def _check(ref_fns, obj, skip=()):
    missing, mismatched = [], []
    for name, (ref_args, n_req) in ref_fns.items():
        if name in skip or (name.startswith("_") and name != "__init__"):
            continue
        fn = getattr(obj, name, None)
        if fn is None:
            missing.append(name)
            continue
        sig = inspect.signature(fn)
        mine = [p.name for p in sig.parameters.values()
                if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
        if inspect.isclass(obj) and mine and mine[0] != "self" and ref_args and ref_args[0] == "self":
            ref_cmp = ref_args[1:]
        else:
            ref_cmp = ref_args
        if mine[: len(ref_cmp)] != ref_cmp:
            mismatched.append((name, ref_cmp, mine))
            continue
        # parameters beyond the reference's must be optional
        extra = list(sig.parameters.values())[len(ref_cmp):]
        if any(p.default is p.empty and p.kind == p.POSITIONAL_OR_KEYWORD for p in extra):
            mismatched.append((name, "extra required parameter", mine))
    return missing, mismatched


def test_engine_methods_match_reference():
    from probabilisticdeepdiffusionmodels_b200.engine import Engine
    ref = _functions(os.path.join(REF, "engine.py"), "Engine")
    missing, mismatched = _check(ref, Engine)
    assert not missing, missing
    assert not mismatched, mismatched


def test_unet_and_factories_match_reference():
    from probabilisticdeepdiffusionmodels_b200 import modules, nn as seam, unet
    ref_unet = _functions(os.path.join(REF, "modules", "unet.py"), "UNetModel")
    missing, mismatched = _check({"__init__": ref_unet["__init__"]}, unet.UNetModel)
    assert not missing and not mismatched, (missing, mismatched)
    for cls in ("ResBlock", "AttentionBlock", "Upsample", "Downsample"):
        ref_init = _functions(os.path.join(REF, "modules", "unet.py"), cls)["__init__"]
        missing, mismatched = _check({"__init__": ref_init}, getattr(unet, cls))
        assert not missing and not mismatched, (cls, missing, mismatched)
    ref_nn = _functions(os.path.join(REF, "modules", "nn.py"))
    missing, mismatched = _check(ref_nn, seam)
    assert not missing, missing
    assert not mismatched, mismatched
    ref_mod = _functions(os.path.join(REF, "modules", "__init__.py"))
    missing, mismatched = _check(ref_mod, modules)
    assert not missing and not mismatched, (missing, mismatched)
