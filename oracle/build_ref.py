#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- stage the UNMODIFIED reference package for the GPU box.

The reference (rongmon/rbvfit 2.4.0, pure Python) lives read-only under /root/reference in the build container and
does not exist on the GPU box.  This recipe installs it with pip -- from a scratch copy under /tmp, because the
build writes egg-info into the source tree -- into ``oracle/_ref/`` (git-ignored, NOT gpurun-ignored, so it travels
with the snapshot exactly like the built ``.so``).  No reference source enters the repository's history.

    python -m oracle.build_ref            # (re)install;  __graft_entry__.build() calls it when /root/reference exists

``--no-deps``: astropy / emcee / corner / matplotlib are not installable here (no network); ``oracle/refshim.py``
provides the minimal stand-ins the reference's module-level imports need.  ``--ignore-requires-python``: the
reference pins python < 3.11 in its metadata, its hot-path code runs unchanged on 3.12.
Users: ``bench.py --impl reference`` (the reference's own ``vfit.lnprob`` under a fork pool on the box's host cores)
and the ``-m gpu`` test that drives the reference's ``vfit`` with a ``GpuVoigtModel`` as the instrument model.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"
TARGET = os.path.join(HERE, "_ref")


def installed() -> bool:
    return os.path.isfile(os.path.join(TARGET, "rbvfit", "vfit_mcmc.py"))


def build(force: bool = False) -> str:
    """Returns the install directory; raises if the reference is neither staged nor available to stage."""
    if installed() and not force:
        return TARGET
    if not os.path.isdir(os.path.join(REFERENCE, "src", "rbvfit")):
        raise RuntimeError(f"{REFERENCE} is not present and oracle/_ref has not been staged")
    with tempfile.TemporaryDirectory(prefix="rbvfit_ref_") as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REFERENCE, src, ignore=shutil.ignore_patterns(".git"))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--ignore-requires-python", "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0 or not installed():
            raise RuntimeError("pip install of the reference failed:\n" + res.stdout + res.stderr)
    return TARGET


if __name__ == "__main__":
    print(build(force=True))
