#!/usr/bin/env python
"""Stage the UNMODIFIED reference next to the repo so that it travels to the GPU box.

    python baseline/stage_ref.py          # /root/reference -> baseline/_ref/   (git-ignored, NOT gpurun-ignored)

The reference (jenicek/mdir) is a plain source tree: it has no setup.py / pyproject.toml, so
`pip install --target baseline/_ref /root/reference` cannot work ("neither 'setup.py' nor 'pyproject.toml'
found"); it is meant to be used from a checkout on sys.path (README.md:24-27).  This script therefore
copies the checkout verbatim.  baseline/_ref/ is listed in .gitignore (reference sources never enter this
repository's history) but not in .gpurunignore, so the copy ships with the gpurun snapshot exactly like
the built .so files do.  Consumers: oracle/ref_import.py (tests + bench CPU baselines only -- never the
product path; tests/test_cpu_boundary.py::test_product_never_imports_oracle guards that).
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("MDIR_REF_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def stage(force=False):
    """Returns the staged path, or None when the source checkout is absent (GPU box: the copy is already there)."""
    if not os.path.isdir(os.path.join(SRC, "mdir")):
        return DST if os.path.isdir(os.path.join(DST, "mdir")) else None
    if os.path.isdir(DST):
        if not force and os.path.isdir(os.path.join(DST, "mdir")):
            return DST
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", ".git"))
    return DST


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
