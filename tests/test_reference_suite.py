"""The reference's OWN tests of the path, unmodified, run against the drop-in.

oracle/Makefile copies tests/test_{pcrv,mindex,mi,lreg,pce}.py of the reference to oracle/_ref/tests (git-ignored; it
travels to the GPU box with the snapshot).  Here every ``import pytuq...`` of those files is answered with the
corresponding ``pytuq_b200`` module, so their assertions exercise the CUDA path through the unchanged API."""
import ast
import importlib
import importlib.util
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = os.path.join(ROOT, "oracle", "_ref", "tests")
FILES = ("test_pcrv.py", "test_mindex.py", "test_mi.py", "test_lreg.py", "test_pce.py")


def _discover():
    found = []
    for fn in FILES:
        path = os.path.join(REF_TESTS, fn)
        if not os.path.isfile(path):
            continue
        tree = ast.parse(open(path).read())
        found += [(fn, node.name) for node in tree.body if isinstance(node, ast.FunctionDef) and node.name.startswith("test")]
    return found


def _alias_modules():
    """{'pytuq[.x.y]': module pytuq_b200[.x.y]} for every module of the package (the module objects themselves
    are not touched)."""
    import pkgutil
    import pytuq_b200
    out = {"pytuq": pytuq_b200}
    for info in pkgutil.walk_packages(pytuq_b200.__path__, "pytuq_b200."):
        if info.name.split(".")[-1].startswith(("_", "lib")):      # private modules, the shared libraries next to them
            continue
        out["pytuq" + info.name[len("pytuq_b200"):]] = importlib.import_module(info.name)
    return out


@pytest.fixture(scope="module")
def reference_test_modules():
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "pytuq" or k.startswith("pytuq.")}
    sys.modules.update(_alias_modules())
    mods = {}
    try:
        for fn in FILES:
            path = os.path.join(REF_TESTS, fn)
            if os.path.isfile(path):
                spec = importlib.util.spec_from_file_location("pcq_reftest_" + fn[:-3], path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                mods[fn] = mod
        yield mods
    finally:
        for k in [k for k in sys.modules if k == "pytuq" or k.startswith("pytuq.")]:
            del sys.modules[k]
        sys.modules.update(saved)


CASES = _discover()


@pytest.mark.skipif(not CASES, reason="oracle/_ref/tests is absent (made by `make -C oracle` where /root/reference exists)")
@pytest.mark.parametrize("fn,name", CASES or [("none", "none")], ids=[f"{f[:-3]}::{n}" for f, n in CASES] or ["absent"])
def test_reference_test_passes_on_the_drop_in(reference_test_modules, fn, name):
    import pytuq_b200
    mod = reference_test_modules[fn]
    for attr in ("PCRV", "PC1d", "PCRV_mvn", "lsq", "PCE", "get_mi", "get_npc", "encode_mindex", "mi_addfront"):
        if hasattr(mod, attr):      # the names the reference tests import resolve to this package
            assert getattr(mod, attr).__module__.startswith("pytuq_b200"), (attr, getattr(mod, attr).__module__)
    getattr(mod, name)()


def test_the_whole_reference_suite_of_the_path_was_found():
    if not CASES:
        pytest.skip("oracle/_ref/tests is absent")
    assert len(CASES) == 26, CASES
