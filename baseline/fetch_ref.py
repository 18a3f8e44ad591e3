"""Vendor the UNMODIFIED reference next to the GPU: /root/reference -> baseline/_ref/ (git-ignored, NOT
gpurun-ignored, so it travels to the GPU box with the snapshot; nothing of it enters the repo's history).

What travels: the env path (`gym_narde/`), the two wrappers its tests need (`web/narde_patched.py`,
`my_game/narde_game_manager.py`), the reference's own runnable tests, and the two caller scripts of
SURVEY.md 8(f)-2 (`evaluate_model.py`, `train_deepq_pytorch.py`) with the checkpoint `evaluate` loads.
Used by: tests/test_gpu_reference_suite.py (the reference's tests and scripts run against the facade),
bench.py's cpu_baseline leg (the Python reference timed on the GPU box's host cores).

The base contract's `pip install --target baseline/_ref /root/reference` is tried first (it installs the
`gym_narde` package); the loose scripts and tests are not part of that package and are copied as files.

    python baseline/fetch_ref.py          # run by __graft_entry__.build() when /root/reference exists
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("NARDE_REFERENCE_ROOT", "/root/reference")

FILES = [
    "gym_narde/__init__.py", "gym_narde/envs/__init__.py", "gym_narde/envs/narde.py", "gym_narde/envs/narde_env.py",
    "gym_narde/envs/rendering.py",
    "web/narde_patched.py", "my_game/__init__.py", "my_game/narde_game_manager.py",
    "tests/test_move_validation.py", "tests/test_doubles_sequence.py", "tests/test_narde_game_manager.py",
    "evaluate_model.py", "train_deepq_pytorch.py", "saved_models/narde_model_final.pt", "LICENSE",
]


def available() -> bool:
    return os.path.isfile(os.path.join(DEST, "gym_narde", "envs", "narde.py"))


def pip_install() -> str:
    """The contract's offline install of the reference package; returns a one-line outcome."""
    tmp = "/tmp/narde_ref_src"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(SRC, tmp, ignore=shutil.ignore_patterns("saved_models", "web", "*.png", "*.log", ".git"))
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
           "--find-links", "/opt/wheelhouse", "--target", DEST, "--upgrade", tmp]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        return "pip install --no-deps ok" if r.returncode == 0 else "pip install failed: " + (r.stderr.strip().splitlines() or ["?"])[-1]
    except Exception as e:  # noqa: BLE001
        return "pip install not run: %s" % e
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def fetch(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print("fetch_ref: %s not present (GPU box): using the prebuilt baseline/_ref" % SRC)
        return available()
    os.makedirs(DEST, exist_ok=True)
    outcome = pip_install()
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DEST, rel)
        if not os.path.isfile(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
    with open(os.path.join(DEST, "FETCHED.txt"), "w") as f:
        f.write("unmodified files of %s; %s\n" % (SRC, outcome))
    if verbose:
        print("fetch_ref: %d files -> %s (%s)" % (len(FILES), DEST, outcome))
    return available()


if __name__ == "__main__":
    sys.exit(0 if fetch() else 1)
