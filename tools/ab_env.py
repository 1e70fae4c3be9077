"""Run tools/ab_variants.py's child once per value of an environment variable (e.g. RL_LEAF_MAX)."""
import os
import subprocess
import sys

var, values = sys.argv[1], sys.argv[2].split(",")
here = os.path.dirname(os.path.abspath(__file__))
for v in values:
    env = dict(os.environ, **{var: v})
    out = subprocess.run([sys.executable, os.path.join(here, "ab_variants.py")] + sys.argv[3:], env=env, capture_output=True, text=True)
    for line in out.stdout.strip().splitlines():
        print(f"{var}={v:4s} {line}", flush=True)
