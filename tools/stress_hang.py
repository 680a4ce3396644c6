#!/usr/bin/env python
"""stress_hang.py [launches] [moves per launch] - many short launches of the headline kernel ((H2O)20 x 4096 chains) with a watchdog:
prints how many completed; exits 3 if one launch does not come back within 30 s (a deadlock in the warp protocol)."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from transrot_b200 import Config, Engine, read_dbase, system_from_config, workload_config

n_launch = int(sys.argv[1]) if len(sys.argv) > 1 else 200
moves = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
golden = os.path.join(ROOT, "tests", "golden")
dbase = read_dbase(os.path.join(golden, "shipped_species.dbase.txt"))
base = Config.parseConfig(os.path.join(golden, "sample8.config.txt"))
cfg = workload_config(base, os.environ.get("WL", "h2o20"), movePerPt=500)
state = {"i": -1, "t": time.time()}

def watchdog():
    while True:
        time.sleep(2)
        if time.time() - state["t"] > 30:
            print(f"HANG: launch {state['i']} has not returned for 30 s", flush=True)
            os._exit(3)

threading.Thread(target=watchdog, daemon=True).start()
eng = Engine(0)
eng.set_system(system_from_config(dbase, cfg))
eng.set_schedules(cfg)
eng.set_options(trace_moves=0)
seeds = [1000 + i for i in range(4096)]
eng.init_chains_seeded(seeds, cfg.spaceLen, cfg.maxPropFailures)
print(eng.kernel_name(), flush=True)
t0 = time.time()
for i in range(n_launch):
    state["i"] = i; state["t"] = time.time()
    eng.run(moves)
    eng.synchronize()
    if bool(eng.summary()["done"].all()):
        eng.init_chains_seeded([s + 7919 * (i + 1) for s in seeds], cfg.spaceLen, cfg.maxPropFailures)
state["t"] = time.time() + 1e9
print(f"{n_launch} launches of {moves} moves completed in {time.time() - t0:.1f} s", flush=True)
