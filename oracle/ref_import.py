"""Import the LIVE reference (/root/reference) for golden-vector generation.

TEST INFRASTRUCTURE ONLY, and only usable in the build container: the reference
tree does not exist on the GPU box, so nothing in tests/, smoke() or bench.py
calls this at run time -- only ``oracle/make_golden.py`` does, and its outputs
are committed under tests/golden/.  Recipe: SURVEY.md App. B (stub two absent
off-path modules, alias a Pillow constant that Pillow 12 dropped).
"""
import os
import sys
import types

REF_ROOT = os.environ.get("MDIR_REF_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "mdir"))


def import_reference():
    """Returns the imported ``mdir`` package (cirtorch is importable afterwards)."""
    if not available():
        raise RuntimeError("reference tree %s not present (GPU box?)" % REF_ROOT)
    sys.dont_write_bytecode = True
    if "h5py" not in sys.modules:
        sys.modules["h5py"] = types.ModuleType("h5py")          # daan file readers only
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        mpl.rcParams = {"font.size": 10.0}
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    from PIL import Image
    if not hasattr(Image, "ANTIALIAS"):
        Image.ANTIALIAS = Image.LANCZOS
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import mdir  # noqa: F401  (side effects: torch threads=3, cv2 threads=1)
    return mdir
