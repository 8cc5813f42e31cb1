#!/usr/bin/env python
"""A/B of kernel tuning knobs on config 2 (each setting in a fresh process): prints phase times.
usage: python profiles/prof_knobs.py"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sets = [{}, {"FEGPU_GATHER_BATCH": "4"}, {"FEGPU_COMPACT": "0"}, {"FEGPU_COMPACT": "0", "FEGPU_GATHER_BATCH": "4"}]
if len(sys.argv) > 1:
    sets = [dict(kv.split("=") for kv in a.split(",") if kv) for a in sys.argv[1:]]
for s in sets:
    env = dict(os.environ); env.update(s)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "prof_elastic.py"), "128"], env=env, capture_output=True, text=True)
    print(json.dumps(s), r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:], flush=True)
