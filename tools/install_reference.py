#!/usr/bin/env python
"""Put the UNMODIFIED reference where bench.py's reference legs can import it on the GPU box: baseline/_ref/.

    python tools/install_reference.py            (dev container; /root/reference must exist)

baseline/_ref/ is git-ignored (the reference's sources never enter this repo's history) but NOT gpurun-ignored, so it
travels to the GPU box with the working tree like the built .so files.  The contract's
`pip install --no-index --target baseline/_ref /root/reference` cannot work -- the reference is a research repo
without setup.py / pyproject.toml (pip: "neither 'setup.py' nor 'pyproject.toml' found") -- so this recipe copies the
Python files the hot path imports (model.py, models/, utils/, settings.py, train_and_test.py, push.py) verbatim and
records their sha256 in baseline/_ref/MANIFEST.json.  Nothing is edited; the two environment shims the reference needs
(`Tensor.cuda` no-op on the CPU arm because model.py:391 hard-codes .cuda(); a stub backbone whose repr starts with
"RES", model.py:107-115) live in bench.py, outside the reference's files.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("MGPROTO_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["model.py", "settings.py", "train_and_test.py", "push.py"]
DIRS = ["models", "utils"]


def install(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print("reference not present at %s: nothing to install" % SRC)
        return False
    os.makedirs(DST, exist_ok=True)
    pip = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                          "--target", os.path.join(DST, "_pip"), SRC], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True)
    outcome = "pip rc=%d: %s" % (pip.returncode, pip.stdout.strip().splitlines()[-1][:200] if pip.stdout.strip() else "")
    shutil.rmtree(os.path.join(DST, "_pip"), ignore_errors=True)
    man = {"source": SRC, "pip_install": outcome, "files": {}}
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    for d in DIRS:
        os.makedirs(os.path.join(DST, d), exist_ok=True)
        for f in sorted(os.listdir(os.path.join(SRC, d))):
            if f.endswith(".py"):
                shutil.copyfile(os.path.join(SRC, d, f), os.path.join(DST, d, f))
    for base, _, files in os.walk(DST):
        for f in sorted(files):
            if f.endswith(".py"):
                p = os.path.join(base, f)
                man["files"][os.path.relpath(p, DST)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    json.dump(man, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    if verbose:
        print("installed %d reference files into %s (%s)" % (len(man["files"]), DST, outcome))
    return True


if __name__ == "__main__":
    install()
