"""TEST INFRASTRUCTURE — dev container only.  Stages the UNMODIFIED reference for the GPU box.

`/root/reference` does not exist on the GPU box; `baseline/_ref/` is git-ignored but travels with the gpurun
snapshot.  This script puts the reference there so that (i) the `-m gpu` test in which the reference's own
`KronLaplace` drives `B200GGN` on the device can run, and (ii) `bench.py --impl reference --workload cora|pubmed`
can time the reference's own classes (`kind: "reference"`).

Recipe: the install the base contract prescribes —
    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref <copy of /root/reference under /tmp>
— succeeds but installs only `laplace_torch-0.2.1.dist-info`: the reference's pyproject declares
`py-modules = ["laplace"]` although `laplace` is a package directory, and never lists the vendored `curvlinops`
nor `gnn` (pyproject.toml:43-44).  The three pure-Python packages the wheel leaves out are therefore copied next
to the dist-info, byte for byte (*.py only).  Nothing under baseline/_ref is tracked by git; no reference source
enters the repository's history.  `oracle/ref_loader.py` finds the staged tree when /root/reference is absent.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("LGNN_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
PACKAGES = ("laplace", "curvlinops", "gnn")


def stage(verbose: bool = True) -> bool:
    if not os.path.isdir(os.path.join(SRC, "laplace")):
        if verbose:
            print(f"[stage_reference] {SRC} not present: nothing staged (GPU box / CI)")
        return False
    os.makedirs(DST, exist_ok=True)
    if not any(n.endswith(".dist-info") for n in os.listdir(DST)):
        tmp = "/tmp/lgnn_refcopy"
        shutil.rmtree(tmp, ignore_errors=True)
        shutil.copytree(SRC, tmp, ignore=shutil.ignore_patterns(".git", "docs", "examples", "logo", "tests"))
        r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                            "--find-links", "/opt/wheelhouse", "--target", DST, tmp], capture_output=True, text=True)
        if verbose:
            print("[stage_reference] pip install rc", r.returncode, (r.stdout.strip().splitlines() or [""])[-1])
        shutil.rmtree(tmp, ignore_errors=True)
    for pkg in PACKAGES:
        dst = os.path.join(DST, pkg)
        shutil.rmtree(dst, ignore_errors=True)
        shutil.copytree(os.path.join(SRC, pkg), dst,
                        ignore=lambda d, names: [n for n in names
                                                 if not (n.endswith(".py") or os.path.isdir(os.path.join(d, n)))
                                                 or n == "__pycache__"])
    if verbose:
        n = sum(len(fs) for _, _, fs in os.walk(DST))
        print(f"[stage_reference] staged {', '.join(PACKAGES)} under {DST} ({n} files)")
    return True


if __name__ == "__main__":
    stage()
