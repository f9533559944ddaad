#!/usr/bin/env python
"""Put the UNMODIFIED reference next to the repo, under the git-ignored baseline/_ref/ (it travels to the GPU box
with the gpurun snapshot; nothing of it enters the history):

  1. `pip install --no-index --no-build-isolation --no-deps --target baseline/_ref /root/reference`
     (the reference's own setup.py; --no-deps because the wheelhouse has no numpy wheel and numpy is installed anyway);
  2. the reference's `examples/` tree (not part of its package) copied beside it, so that its example models
     (`examples/variational_autoencoder/{iwae,vae_mnist}.py`, `examples/bayesian_neural_nets/{bnn_vi,bnn_sgmcmc}.py`)
     can be imported unmodified -- by tests/test_reference_examples.py against THIS package, and by
     `bench.py --impl reference` against the reference itself.

Usage: python tools/vendor_reference.py [--reference /root/reference] [--force]
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")


def vendor(reference="/root/reference", force=False):
    """Returns DEST, or None when the reference tree is not available (e.g. on the GPU box: use what travelled)."""
    done = os.path.join(DEST, ".vendored")
    if os.path.exists(done) and not force:
        return DEST
    if not os.path.isdir(reference):
        return DEST if os.path.isdir(os.path.join(DEST, "zhusuan")) else None
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    src = reference
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
           "--find-links", "/opt/wheelhouse", "--target", DEST]
    r = subprocess.run(cmd + [src], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        # a read-only source tree: the build wants to write egg-info next to setup.py -> install from a copy
        tmp = "/tmp/zs_reference_copy"
        shutil.rmtree(tmp, ignore_errors=True)
        shutil.copytree(reference, tmp, ignore=shutil.ignore_patterns(".git"))
        r = subprocess.run(cmd + [tmp], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n" + r.stdout)
    shutil.copytree(os.path.join(reference, "examples"), os.path.join(DEST, "examples"),
                    ignore=shutil.ignore_patterns("__pycache__", "data", "result"))
    with open(done, "w") as f:
        f.write("reference: %s\n" % reference)
    return DEST


if __name__ == "__main__":
    ref = "/root/reference"
    if "--reference" in sys.argv:
        ref = sys.argv[sys.argv.index("--reference") + 1]
    print(vendor(ref, force="--force" in sys.argv))
