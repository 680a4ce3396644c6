"""How long do device and oracle take identical decisions?  (H2O)20, full shipped schedule shortened to N moves."""
import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from helpers import compare_traces, oracle_start, workload_config
from transrot_b200 import Config, Engine, read_dbase, system_from_config

g = os.path.join("tests", "golden")
dbase = read_dbase(os.path.join(g, "shipped_species.dbase.txt")); base = Config.parseConfig(os.path.join(g, "sample8.config.txt"))
name = sys.argv[1] if len(sys.argv) > 1 else "h2o20"
mpp = int(sys.argv[2]) if len(sys.argv) > 2 else 25000
cfg = workload_config(base, name, movePerPt=mpp)          # 4 teeth x 10 points x mpp
n = cfg.total_moves()
seeds = [1000, 1001, 1002]
with Engine(0) as e:
    e.set_system(system_from_config(dbase, cfg)); e.set_schedules(cfg)
    e.set_options(trace_moves=n)
    e.init_chains_seeded(seeds, cfg.spaceLen, cfg.maxPropFailures)
    t0 = time.time(); e.run(0); tr = e.trace(); print("device", time.time() - t0, "s")
    mins = e.pack_minima()
for i, s in enumerate(seeds):
    sp = oracle_start(dbase, cfg, s)
    t0 = time.time(); e0 = sp.write_snapshot(0); res = sp.sawtooth_anneal(e0, trace_cap=n, snap_cap=8)
    first, err_rel, err_cond = compare_traces(tr[i], res.trace, res.trace_abs)
    emin, tmin = sp.min_energy()
    print(f"{name} seed {s}: {n} moves, first divergence {first}, err_cond {err_cond:.2e}, oracle {time.time() - t0:.1f}s, "
          f"accepted {int(res.trace['accepted'].sum())}, min energy dev {mins['energy'][i]:.12f} oracle {emin:.12f} tooth {mins['tooth'][i]}/{tmin}")
