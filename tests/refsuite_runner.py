"""Runs the UNMODIFIED reference callers (vendored by baseline/fetch_ref.py into baseline/_ref) in a fresh
process, either on the real reference env (`--impl reference`) or on this repo's GPU facade aliased as
`gym_narde` (`--impl facade`), and prints one JSON line with the outcome -- TEST INFRASTRUCTURE.

  --what tests     tests/test_move_validation.py, test_doubles_sequence.py, test_narde_game_manager.py: per-test outcome
  --what evaluate  evaluate_model.evaluate(num_games=..) with the shipped checkpoint: its printed report
  --what train     train_deepq_pytorch.main(episodes=..): its printed episode lines

Every RNG the callers use (numpy global, `random`, torch) is seeded identically in both modes, and the facade
consumes the global numpy stream exactly like the reference env (narde_env.py:29,107-115), so the two modes must
print the same thing.
"""
import argparse
import contextlib
import importlib.util
import io
import json
import logging
import os
import random
import shutil
import sys
import tempfile
import unittest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def setup(impl):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, ROOT)
    import gymnasium_stub
    gymnasium_stub.install()
    if impl == "facade":
        import gym_narde_b200
        import gym_narde_b200.envs
        import gym_narde_b200.envs.narde
        import gym_narde_b200.envs.narde_env
        gym_narde_b200._register_with_gymnasium()       # the stub arrived after the package was imported
        sys.modules["gym_narde"] = gym_narde_b200
        sys.modules["gym_narde.envs"] = gym_narde_b200.envs
        sys.modules["gym_narde.envs.narde"] = gym_narde_b200.envs.narde
        sys.modules["gym_narde.envs.narde_env"] = gym_narde_b200.envs.narde_env
        sys.path.append(REF)                            # web/, my_game/, the scripts (gym_narde is already aliased)
    else:
        sys.path.insert(0, REF)
    logging.disable(logging.CRITICAL)                   # the reference logs every move at INFO level


def seed_all(seed):
    import numpy as np
    import torch
    np.random.seed(seed)
    random.seed(seed)
    torch.manual_seed(seed)


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def run_tests():
    out = {}
    for k, f in enumerate(("test_move_validation.py", "test_doubles_sequence.py", "test_narde_game_manager.py")):
        seed_all(100 + k)
        mod = load(os.path.join(REF, "tests", f), "ref_" + f[:-3])
        suite = unittest.TestLoader().loadTestsFromModule(mod)
        res = unittest.TestResult()
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            suite.run(res)
        bad = {t.id().split(".", 1)[1]: "fail" for t, _ in res.failures}
        bad.update({t.id().split(".", 1)[1]: "error" for t, _ in res.errors})
        bad.update({t.id().split(".", 1)[1]: "skip" for t, _ in res.skipped})

        def names(s):
            for t in s:
                if isinstance(t, unittest.TestSuite):
                    yield from names(t)
                else:
                    yield t.id().split(".", 1)[1]
        out[f] = {n: bad.get(n, "ok") for n in names(unittest.TestLoader().loadTestsFromModule(mod))}
    return out


def run_script(what, n):
    tmp = tempfile.mkdtemp(prefix="narde_refsuite_")
    os.makedirs(os.path.join(tmp, "saved_models"))
    shutil.copy(os.path.join(REF, "saved_models", "narde_model_final.pt"), os.path.join(tmp, "saved_models"))
    os.chdir(tmp)
    buf = io.StringIO()
    try:
        seed_all(7)
        with contextlib.redirect_stdout(buf):
            if what == "evaluate":
                import evaluate_model
                seed_all(7)
                evaluate_model.evaluate(model_name="narde_model_final.pt", num_games=n, render=False)
            else:
                import train_deepq_pytorch
                seed_all(7)
                train_deepq_pytorch.main(episodes=n, max_steps=1000, epsilon=0.9, epsilon_decay=0.9)
    finally:
        os.chdir(ROOT)
        shutil.rmtree(tmp, ignore_errors=True)
    lines = [l for l in buf.getvalue().splitlines() if l.strip() and "saved to" not in l and "using device" not in l]
    return lines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["facade", "reference"], required=True)
    ap.add_argument("--what", choices=["tests", "evaluate", "train"], required=True)
    ap.add_argument("-n", type=int, default=3)
    a = ap.parse_args()
    setup(a.impl)
    res = run_tests() if a.what == "tests" else run_script(a.what, a.n)
    import gym_narde
    print("REFSUITE " + json.dumps({"impl": a.impl, "what": a.what, "env_module": gym_narde.__name__ + " @ " + os.path.dirname(gym_narde.__file__),
                                    "result": res}))


if __name__ == "__main__":
    main()
