"""BASELINE.json's five configurations: particles and initial box of each named workload."""
from __future__ import annotations

from .host import Config

WORKLOADS = {
    "sample8": dict(particles=[("NH4+", 4), ("Cl-", 4)], spaceLen=5.0),
    "h2o20": dict(particles=[("H2O", 20)], spaceLen=10.0),
    "nh4cl32": dict(particles=[("NH4+", 32), ("Cl-", 32)], spaceLen=14.0),
    "mixed64": dict(particles=[("NH4+", 16), ("Cl-", 16), ("H2O", 32)], spaceLen=14.0),
    "h2o256": dict(particles=[("H2O", 256)], spaceLen=24.0),
}


def workload_config(base: Config, name: str, **kw) -> Config:
    """`base` (an annealing block, e.g. the shipped config.txt) with the particles and box of workload `name`."""
    w = WORKLOADS[name]
    return base.copy(particles=list(w["particles"]), spaceLen=w["spaceLen"], **kw)


def sweep_schedules(base: Config, max_temps=(2000.0, 4000.0, 6000.0, 8000.0, 10000.0, 12000.0, 14000.0, 16000.0),
                    magwalk_probs=(0.0, 0.05, 0.10, 0.15, 0.20, 0.25, 0.30, 0.35)):
    """BASELINE config #3: the annealing-parameter sweep, sawtooth peaks x magwalk probabilities
    (translation = rotation).  Returns the list of parameter sets; chain c uses set (c // seeds_per_set)."""
    return [base.copy(maxTemp=t, magwalkProbTrans=p, magwalkProbRot=p) for t in max_temps for p in magwalk_probs]
