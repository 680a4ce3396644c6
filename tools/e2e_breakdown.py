#!/usr/bin/env python
"""Where the end-to-end time of one job goes (host clock around each public call), for one workload."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from transrot_b200 import Config, Engine, read_dbase, system_from_config, workload_config

name = sys.argv[1] if len(sys.argv) > 1 else "h2o256"
n = int(sys.argv[2]) if len(sys.argv) > 2 else (888 if name == "h2o256" else 4096)
mpp = int(sys.argv[3]) if len(sys.argv) > 3 else 500
golden = os.path.join(ROOT, "tests", "golden")
dbase = read_dbase(os.path.join(golden, "shipped_species.dbase.txt"))
cfg = workload_config(Config.parseConfig(os.path.join(golden, "sample8.config.txt")), name, movePerPt=mpp, writeConfHeatCapacitiesEnabled=False)
system = system_from_config(dbase, cfg)
seeds = np.arange(1000, 1000 + n, dtype=np.int64)
with Engine(0) as e:
    e.set_system(system); e.set_schedules(cfg)
    for rep in range(3):
        t = [time.perf_counter()]
        e.set_options(trace_moves=0); e.init_chains_seeded(seeds, cfg.spaceLen, cfg.maxPropFailures); t.append(time.perf_counter())
        e.run(0); e.synchronize(); t.append(time.perf_counter())
        m = e.pack_minima(); t.append(time.perf_counter())
        s = e.summary(); t.append(time.perf_counter())
        p = e.points(); t.append(time.perf_counter())
        se, st, sx = e.snapshots(with_sites=True); t.append(time.perf_counter())
        d = np.diff(t) * 1e3
        print(f"{name} x{n} mpp={mpp} rep {rep}: init {d[0]:.1f} ms, run {d[1]:.1f}, minima {d[2]:.2f}, summary {d[3]:.2f}, points {d[4]:.2f}, snapshots+sites {d[5]:.1f} ({sx.nbytes/1e6:.0f} MB); kernel {e.last_run_ms():.1f}")
