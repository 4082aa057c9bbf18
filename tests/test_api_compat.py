"""Drop-in surface check against the reference's own modules (runs only where /root/reference is
mounted, i.e. in the build container; skipped on the GPU box): every public class / function of the
reference's pose-path modules exists under the same name in the mirror with the same parameter names,
so `from binDeltaModels import OneBinDeltaModel` etc. keep working in the learn*/evaluate* scripts."""
import importlib.util
import inspect
import os
import sys
import types

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multi-modal-regression_b200")

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")

# module -> names the mirror deliberately leaves to the reference (datasets / disk I/O)
LEFT_TO_REFERENCE = {
    "objectnetHelperFunctions": set(),
    "binDeltaGenerators": set(),
    "binDeltaModels": set(), "binDeltaLosses": set(), "poseModels": set(),
    "axisAngle": set(), "quaternion": set(), "featureModels": set(), "helperFunctions": set(),
}


def _load_reference(name):
    """Import /root/reference/<name>.py in isolation with stubs for what the container lacks."""
    for stub in ("tensorboardX", "progressbar"):
        sys.modules.setdefault(stub, types.ModuleType(stub))
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in list(LEFT_TO_REFERENCE) +
                  ["featureModels", "helperFunctions", "dataGenerators"]}
    try:
        sys.path.insert(0, REF)
        for k in saved_mods:
            sys.modules.pop(k, None)
        spec = importlib.util.spec_from_file_location("_ref_" + name, os.path.join(REF, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        sys.path[:] = saved_path
        for k, v in saved_mods.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v


def _public(mod, modname):
    out = {}
    for n, obj in vars(mod).items():
        if n.startswith("_"):
            continue
        if (inspect.isclass(obj) or inspect.isfunction(obj)) and getattr(obj, "__module__", "") == mod.__name__:
            out[n] = obj
    return out


def _params(obj):
    fn = obj.__init__ if inspect.isclass(obj) else obj
    try:
        return [p for p in inspect.signature(fn).parameters if p != "self"]
    except (TypeError, ValueError):
        return None


@pytest.mark.parametrize("modname", sorted(LEFT_TO_REFERENCE))
def test_mirror_exposes_reference_names(modname):
    ref = _load_reference(modname)
    # the deployment layout of INTEGRATION.md: the mirror first, the reference right behind it (the
    # generators take the image side of the datasets from the reference's dataGenerators)
    for p in (REF, PKG):
        if p in sys.path:
            sys.path.remove(p)
    sys.path[:0] = [PKG, REF]
    try:
        sys.modules.pop(modname, None)
        ours = importlib.import_module(modname)
        _check(modname, ref, ours)
    finally:
        sys.path.remove(REF)
        for k in ("dataGenerators", "helperFunctions", "featureModels"):
            m = sys.modules.get(k)
            if m is not None and os.path.dirname(os.path.abspath(getattr(m, "__file__", ""))) == REF:
                sys.modules.pop(k)


def _check(modname, ref, ours):
    assert os.path.dirname(os.path.abspath(ours.__file__)) == PKG, "mirror module is shadowed"
    missing, mismatched = [], []
    for n, robj in _public(ref, modname).items():
        if n in LEFT_TO_REFERENCE[modname]:
            continue
        oobj = getattr(ours, n, None)
        if oobj is None:
            missing.append(n)
            continue
        rp, op = _params(robj), _params(oobj)
        if rp is not None and op is not None and rp != op[:len(rp)]:
            mismatched.append((n, rp, op))
    assert not missing, "%s: names of the reference missing in the mirror: %s" % (modname, missing)
    assert not mismatched, "%s: parameter lists differ: %s" % (modname, mismatched)
