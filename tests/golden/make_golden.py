"""Generates tests/golden/kat.json.

The reference cannot run in this image (no JVM), so these known answers come from (i) the canonical mt19937ar
vector, (ii) numpy.random.RandomState as an independent MT19937, (iii) the reference's own six combination-rule
goldens (src/interaction_params.txt:6,7,12), (iv) SURVEY.md Appendix C (surveyor's scratch Python), and (v) the C
oracle cross-checked against oracle/pyref.py.  Run from the repo root: python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import oracle_start, workload_config  # noqa: E402
from oracle.oracle import JavaMT, OracleSpace  # noqa: E402
from transrot_b200 import Config, read_dbase, read_input  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
dbase = read_dbase(os.path.join(G, "shipped_species.dbase.txt"))
cfg = Config.parseConfig(os.path.join(G, "sample8.config.txt"))
out = {}

g = JavaMT()
g.set_seed_array([0x123, 0x234, 0x345, 0x456])
out["mt19937ar_init_by_array_123_234_345_456"] = [g.next(32) for _ in range(5)]           # canonical mt19937ar.out
g = JavaMT(42)
out["setSeed42_next32"] = [g.next(32) for _ in range(8)]
assert out["setSeed42_next32"] == [int(v) for v in np.random.RandomState(np.array([0, 42], dtype=np.uint32)).randint(0, 2**32, size=8, dtype=np.uint64)]
g = JavaMT(42)
out["setSeed42_nextDouble"] = [g.next_double() for _ in range(3)]
sp = OracleSpace(dbase)
sp.set_seed(42)
out["seed42_fanout_M_S_R"] = sp.stream_seeds()
pt = sp.pair_table()
out["combination_rule"] = {"N-Cl": [pt[0, 5, 2], pt[0, 5, 3]], "N-O": [pt[0, 6, 2], pt[0, 6, 3]], "Cl-O": [pt[5, 6, 2], pt[5, 6, 3]]}
system, xyz = read_input(os.path.join(G, "sample8.input.xyz"), dbase)
sp.set_config(cfg)
sp.read_input([("NH4+", 4, False), ("Cl-", 4, False)], xyz)
out["input_xyz_total_energy"] = sp.calc_energy()
out["input_xyz_first_com"] = [float(v) for v in sp.coords()[1][0]]
# shipped schedule: temperatures passed to writeAcceptance in tooth 1 and delT of every tooth (IEEE-exact)
t, dels, temps = cfg.maxTemp, [], []
save = t
for tooth in range(cfg.numTeeth):
    d = t / (cfg.ptsPerTooth - 1)
    dels.append(d)
    for y in range(cfg.ptsPerTooth):
        if tooth == 0:
            temps.append(t)
        if not (abs(t) < 0.005):      # heat-capacity `continue` (Main.java:228)
            t -= d
            if t < 0:
                t = 0
    save *= cfg.tempDecreasePerTooth
    t = save
out["shipped_tooth1_temperatures"] = temps
out["shipped_delT_per_tooth"] = dels
# oracle run of the shipped sample for seed 42, first 2000 moves of a shortened schedule (decision fingerprint)
short = workload_config(cfg, "sample8", movePerPt=500, ptsPerTooth=5, numTeeth=2)
o = oracle_start(dbase, short, 42)
e0 = o.calc_energy()
res = o.sawtooth_anneal(e0, trace_cap=2000)
out["sample8_seed42"] = {
    "space_len": o.space_len, "start_energy": e0,
    "accepted_first2000": int(res.trace["accepted"].sum()), "rotations_first2000": int(res.trace["kind"].sum()),
    "bounds_first2000": int(res.trace["bounds"].sum()), "mol_sum_first2000": int(res.trace["mol"].sum()),
    "running_after_2000": float(res.trace["running"][-1]),
}
with open(os.path.join(G, "kat.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps(out, indent=1)[:1500])
