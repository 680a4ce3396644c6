#!/usr/bin/env python
"""Small runs of every hot kernel family for compute-sanitizer (tools/sanitize.sh): the crew kernel with 16- and 32-lane
teams, the team kernels (half-warp, warp, block per chain), a resumed (`max_moves`) run and the probes.  No oracle, no
torch: just the C ABI, so the sanitizer only sees this library's kernels.  Prints one line per case."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from transrot_b200 import (Config, Engine, TR_KERNEL_AUTO, TR_KERNEL_CREW, TR_KERNEL_TEAM, read_dbase, system_from_config,  # noqa: E402
                           workload_config)

golden = os.path.join(ROOT, "tests", "golden")
dbase = read_dbase(os.path.join(golden, "shipped_species.dbase.txt"))
base = Config.parseConfig(os.path.join(golden, "sample8.config.txt"))

# (label, workload, kernel, threads_per_chain, chains, moves per point, resumed chunks)
CASES = {
    "crew16": ("h2o20", TR_KERNEL_AUTO, 0, 31, 60, None),
    "crew16_resume": ("h2o20", TR_KERNEL_AUTO, 0, 9, 50, (1, 37, 120)),
    "crew32": ("mixed64", TR_KERNEL_CREW, 0, 5, 30, None),
    "crew_small": ("sample8", TR_KERNEL_AUTO, 0, 17, 60, None),
    "team16": ("h2o20", TR_KERNEL_TEAM, 16, 9, 40, None),
    "team32": ("mixed64", TR_KERNEL_AUTO, 0, 5, 30, None),
    "team128": ("h2o256", TR_KERNEL_AUTO, 0, 2, 12, None),
    "team256": ("mixed64", TR_KERNEL_AUTO, 256, 2, 20, None),
}


def run(label):
    name, kernel, tpc, n, mpp, chunks = CASES[label]
    cfg = workload_config(base, name, movePerPt=mpp, ptsPerTooth=3, numTeeth=2, finale=True)
    system = system_from_config(dbase, cfg)
    seeds = [1000 + i for i in range(n)]
    with Engine(0) as e:
        e.set_system(system)
        e.set_schedules(cfg)
        e.set_kernel(kernel)
        e.set_options(trace_moves=0, threads_per_chain=tpc)
        e.init_chains_seeded(seeds, cfg.spaceLen, cfg.maxPropFailures)
        for c in chunks or ():
            e.run(c)
        e.run(0)
        s = e.summary()
        assert s["done"].all() and (s["moves"] == cfg.total_moves()).all(), label
        x, _ = e.sites()
        en = e.total_energy(x[:2])
        m = e.pack_minima()
        e.min_structures()
        e.rng_state()
        print(f"{label}: {n} chains x {cfg.total_moves()} moves on {name}, min {m['energy'].min():.6f}, E[0] {en[0]:.6f}")


if __name__ == "__main__":
    for lab in (sys.argv[1:] or list(CASES)):
        run(lab)
