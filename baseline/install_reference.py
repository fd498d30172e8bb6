"""
Installs the UNMODIFIED reference (aubin-tchoi/shot-fpfh, mounted read-only at /root/reference in the build container)
into baseline/_ref, where `bench.py --impl reference` and the `-m gpu` drop-in test import it from. baseline/_ref is
git-ignored (no reference source enters the history) but travels to the GPU box with the rest of the tree.

    python baseline/install_reference.py [--force]

Recipe, in this order (the outcome is recorded in baseline/_ref/INSTALL_RECORD.json and in DESIGN.md):
  1. `pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target baseline/_ref` from a
     copy of the tree (the mount is read-only). The reference's build backend is poetry-core, which is neither installed
     nor in the offline wheelhouse: this step fails here with `ModuleNotFoundError: No module named 'poetry'`.
  2. The package is pure Python (26 .py files, no build step: pyproject.toml only lists the package directory), so what
     pip would have installed is the directory itself: `shot_fpfh/` (+ `scripts/`, `config/` for the CLI) are copied
     as they are.
"""

from __future__ import annotations

import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
SOURCE = "/root/reference"


def install(force: bool = False) -> dict | None:
    record_path = os.path.join(TARGET, "INSTALL_RECORD.json")
    if os.path.exists(record_path) and not force:
        with open(record_path) as f:
            return json.load(f)
    if not os.path.isdir(os.path.join(SOURCE, "shot_fpfh")):
        return None  # not the build container: the prebuilt baseline/_ref (if any) is what there is
    if os.path.isdir(TARGET):
        shutil.rmtree(TARGET)
    record = {"source": SOURCE}
    with tempfile.TemporaryDirectory() as tmp:
        copy = os.path.join(tmp, "reference")
        shutil.copytree(SOURCE, copy)
        proc = subprocess.run(
            [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links", "/opt/wheelhouse",
             "--no-deps", "--target", TARGET, copy],
            capture_output=True, text=True,
        )
    if proc.returncode == 0 and os.path.isdir(os.path.join(TARGET, "shot_fpfh")):
        record["method"] = "pip install --no-index --no-build-isolation --no-deps --target"
    else:
        record["pip_error"] = (proc.stderr or proc.stdout).strip().splitlines()[-1:]
        record["method"] = "copy of the pure-Python package directories (no build step exists)"
        os.makedirs(TARGET, exist_ok=True)
        for name in ("shot_fpfh", "scripts", "config"):
            shutil.copytree(os.path.join(SOURCE, name), os.path.join(TARGET, name))
    n_files = sum(len([f for f in files if f.endswith(".py")]) for _, _, files in os.walk(TARGET))
    record["python_files"] = n_files
    with open(record_path, "w") as f:
        json.dump(record, f, indent=1)
    return record


if __name__ == "__main__":
    print(json.dumps(install(force="--force" in sys.argv), indent=1))
