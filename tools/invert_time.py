"""Time sgw_invert_epsilon at the size of a full Si64 q-point for a few panel plans (one JSON line each)."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from sternheimergw_b200 import Context  # noqa: E402

ctx = Context(0)
plans = [{}, {"SGW_GJ_NB": "32"}, {"SGW_GJ_NB": "128"}, {"SGW_GJ_NB": "96"}, {"SGW_GJ_SB": "2"}, {"SGW_GJ_PANEL": "global"}]
for env in plans:
    for k in ("SGW_GJ_NB", "SGW_GJ_SB", "SGW_GJ_PANEL"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for rep in range(2):
        r = bench.invert_epsilon_full(ctx)
    print(json.dumps({"plan": env, **r}), flush=True)
