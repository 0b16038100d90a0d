"""Test-side access to the real reference rec.io built into oracle/_ref (see build_ref.py).  Test infrastructure only."""
import json
import os
import subprocess
import sys

from . import build_ref

HERE = os.path.dirname(os.path.abspath(__file__))


def available() -> bool:
    return build_ref.available()


def call(requests):
    """runs a batch of requests in ONE reference interpreter; returns the list of replies"""
    if not available():
        raise RuntimeError("oracle/_ref is not built (python oracle/build_ref.py, needs /root/reference)")
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    p = subprocess.run([sys.executable, os.path.join(HERE, "ref_runner.py")], input=json.dumps({"requests": requests}),
                       capture_output=True, text=True, env=env, check=False)
    if p.returncode != 0:
        raise RuntimeError("reference runner failed:\n" + p.stderr[-4000:])
    return json.loads(p.stdout)["replies"]
