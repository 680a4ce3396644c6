"""CPU tests of the oracle (the checker itself): pinned against the reference's only in-tree goldens, the canonical
MT19937 vector, numpy's independent MT19937, SURVEY.md Appendix C, and an independent Python restatement."""
import json
import math
import os

import numpy as np
import pytest

from helpers import oracle_start, workload_config
from oracle.oracle import JavaMT, OracleSpace
from oracle.pyref import PyMT, PySpace
from transrot_b200 import read_input


@pytest.fixture(scope="module")
def kat(golden_dir):
    with open(os.path.join(golden_dir, "kat.json")) as fh:
        return json.load(fh)


def test_mt19937ar_canonical_vector(kat):
    g = JavaMT()
    g.set_seed_array([0x123, 0x234, 0x345, 0x456])
    assert [g.next(32) for _ in range(5)] == [1067595299, 955945823, 477289528, 4107218783, 4228976476]
    assert kat["mt19937ar_init_by_array_123_234_345_456"] == [1067595299, 955945823, 477289528, 4107218783, 4228976476]


@pytest.mark.parametrize("seed", [0, 42, -1, 2**40 + 17, -(2**62) + 5])
def test_set_seed_long_matches_numpy_mt19937(seed):
    """MersenneTwister.setSeed(long) = init_by_array({hi32, lo32}); numpy's RandomState is an independent MT19937."""
    g = JavaMT(seed)
    u = seed & 0xFFFFFFFFFFFFFFFF
    rs = np.random.RandomState(np.array([u >> 32, u & 0xFFFFFFFF], dtype=np.uint32))
    want = rs.randint(0, 2**32, size=1500, dtype=np.uint64)
    got = np.array([g.next(32) for _ in range(1500)], dtype=np.uint64)
    assert np.array_equal(got, want)


def test_bits_stream_generator_derived_draws(kat):
    g = JavaMT(42)
    assert [g.next_double() for _ in range(3)] == kat["setSeed42_nextDouble"] == [0.433919173241921, 0.6391116549738762, 0.5736409594260623]
    # nextInt: power-of-two path and the rejection path agree with the Python restatement
    for n in (3, 8, 20, 64, 100):
        a, b = JavaMT(7), PyMT(7)
        assert [a.next_int(n) for _ in range(200)] == [b.next_int(n) for _ in range(200)]
    a, b = JavaMT(-5), PyMT(-5)
    assert [a.next_long() for _ in range(20)] == [b.next_long() for _ in range(20)]


def test_seed_fanout(dbase, kat):
    sp = OracleSpace(dbase)
    sp.set_seed(42)
    assert sp.stream_seeds() == [8004395815503774571, -6657214946198437198, -7864936058809080539] == kat["seed42_fanout_M_S_R"]


def test_combination_rule_matches_reference_goldens(dbase):
    """The six full-precision values of the reference's src/interaction_params.txt:6,7,12 (bit-exact)."""
    pt = OracleSpace(dbase).pair_table()
    N, Cl, O = 0, 5, 6
    assert (pt[N, Cl, 2], pt[N, Cl, 3]) == (3099.1447746519393, 6911464.963027542)
    assert (pt[N, O, 2], pt[N, O, 3]) == (690.5308569597016, 741339.4434172494)
    assert (pt[Cl, O, 2], pt[Cl, O, 3]) == (2670.6432816052015, 5425977.317740446)
    assert np.array_equal(pt, np.transpose(pt, (1, 0, 2)))


def test_input_xyz_energy_and_com(dbase, sample_cfg, golden_dir, kat):
    system, xyz = read_input(os.path.join(golden_dir, "sample8.input.xyz"), dbase)
    sp = OracleSpace(dbase)
    sp.set_config(sample_cfg)
    sp.read_input([("NH4+", 4, False), ("Cl-", 4, False)], xyz)
    assert sp.calc_energy() == pytest.approx(-460.8203874165545, rel=1e-13)
    assert sp.calc_energy() == kat["input_xyz_total_energy"]
    assert np.allclose(sp.coords()[1][0], [1.909373313536205, 0.9268264482318975, 0.5437754582402793], rtol=0, atol=1e-15)


def test_shipped_schedule_temperatures(dbase, sample_cfg, kat):
    cfg = workload_config(sample_cfg, "sample8", movePerPt=3)
    sp = oracle_start(dbase, cfg, 42)
    res = sp.sawtooth_anneal(sp.calc_energy(), points_cap=cfg.num_points(), snap_cap=cfg.numTeeth)
    temps = res.points["temp"]
    assert list(temps[:10]) == kat["shipped_tooth1_temperatures"]
    assert temps[9] == 0.0 and temps[19] == 1.1368683772161603e-12 and temps[39] == 9.094947017729282e-13   # Appendix A Q11
    # the `continue` at Main.java:228 skips the movie frame on the last point of every tooth (round(t,2) == 0)
    assert list(res.points["movie_written"][[8, 9, 19, 29, 39]]) == [1, 0, 0, 0, 0]
    assert list(res.snap_tooth) == [1, 2, 3, 4]
    assert res.moves == cfg.total_moves() == 120


@pytest.mark.parametrize("name,seed", [("sample8", 42), ("h2o20", 7), ("mixed64", 3)])
def test_c_oracle_matches_python_restatement(dbase, sample_cfg, name, seed):
    """Two independently written restatements (C + glibc, Python + numpy MT) take identical decisions and agree on
    energies to rounding for the first few hundred moves, incl. propagate, box growth and the initial rotations."""
    n = 150 if name == "mixed64" else 400
    cfg = workload_config(sample_cfg, name, movePerPt=100, ptsPerTooth=3, numTeeth=2)
    c = oracle_start(dbase, cfg, seed)
    p = PySpace(dbase, cfg, seed)
    p.propagate()
    assert p.spaceLen == c.space_len
    sites, com = c.coords()
    assert np.max(np.abs(np.array([a for m in p.xyz for a in m]) - sites)) < 1e-13
    e0c, e0p = c.calc_energy(), p.calc_energy()
    assert e0c == pytest.approx(e0p, rel=1e-12)
    res = c.sawtooth_anneal(e0c, trace_cap=n)
    ref = p.sawtooth_anneal(e0p, n)
    assert [int(v) for v in res.trace["mol"]] == [r[0] for r in ref]
    assert [int(v) for v in res.trace["kind"]] == [r[1] for r in ref]
    assert [int(v) for v in res.trace["accepted"]] == [r[2] for r in ref]
    assert np.allclose(res.trace["running"], [r[3] for r in ref], rtol=1e-11, atol=1e-9)
    assert list(res.trace["temp"]) == [r[4] for r in ref]


def test_finale_and_static_modes(dbase, sample_cfg):
    cfg = workload_config(sample_cfg, "sample8", movePerPt=50, ptsPerTooth=3, numTeeth=2, finale=True, writeConfHeatCapacitiesEnabled=False)
    sp = oracle_start(dbase, cfg, 5)
    res = sp.sawtooth_anneal(sp.calc_energy(), points_cap=99, snap_cap=9, trace_cap=10**4)
    assert len(res.points) == cfg.num_points() == 9
    assert list(res.points["finale"]) == [0] * 6 + [1] * 3 and list(res.points["tooth"]) == [0, 0, 0, 1, 1, 1, 1, 1, 1]
    assert list(res.snap_tooth) == [1, 2]                      # the pre-finale tooth writes no Output file (Main.java:245-257)
    assert np.all(res.points["temp"][6:] == 0.0) and res.trace["magwalk"][300:].sum() == 0
    static = workload_config(sample_cfg, "sample8", movePerPt=200, staticTemp=True, eqConfigs=50, maxTemp=300.0)
    sp = oracle_start(dbase, static, 5)
    res = sp.sawtooth_anneal(sp.calc_energy(), points_cap=9, snap_cap=9, trace_cap=200)
    assert len(res.points) == 1 and res.moves == 200 and np.all(res.trace["temp"] == 300.0)
    # equilibration skip: z > eqConfigs (Main.java:216)
    assert res.points["totalEnergy"][0] == pytest.approx(float(np.sum(res.trace["running"][51:])), rel=1e-12)
