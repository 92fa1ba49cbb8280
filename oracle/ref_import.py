"""Import the LIVE reference: /root/reference in the build container, else the verbatim copy that
``baseline/stage_ref.py`` placed under baseline/_ref/ (git-ignored; it travels to the GPU box with the gpurun
snapshot, where /root/reference does not exist).

TEST INFRASTRUCTURE ONLY: used by ``oracle/make_golden.py`` (golden vectors committed under tests/golden/),
by the integration tests (validate() un-patched vs after mdir_b200.install()) and by bench.py's CPU-baseline
legs.  Never imported by the product package.  Recipe: SURVEY.md App. B (stub two absent off-path modules,
alias a Pillow constant that Pillow 12 dropped).
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    for cand in (os.environ.get("MDIR_REF_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "mdir")):
            return cand
    return "/root/reference"


REF_ROOT = _find_root()


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
