"""More parity of the CUDA path (through the C ABI) against the CPU oracle: schedule modes, the Born-Mayer
term, explicit starts with frozen particles, parameter sweeps, resumability, ragged chain counts, every kernel
variant, and 2-GPU sharding invariance.  All tests need a B200.  Bars as in test_gpu_parity.py."""
import os

import numpy as np
import pytest

from helpers import compare_traces, oracle_start, rel_err, workload_config
from oracle.oracle import JavaMT, OracleSpace
from transrot_b200 import Config, Engine, TR_KERNEL_AUTO, TR_KERNEL_CREW, read_dbase, read_input, system_from_config

pytestmark = pytest.mark.gpu

TOL = 1e-12


def _check_records(engine, dbase, cfg, seeds, pair_table=None, trace_moves=0):
    """Whole-anneal comparison of every per-point / per-tooth record with the oracle."""
    summ, pts = engine.summary(), engine.points()
    se, st, sx = engine.snapshots()
    sites, com = engine.sites()
    mins = engine.pack_minima()
    tr = engine.trace() if trace_moves else None
    for i, seed in enumerate(seeds):
        sp = oracle_start(dbase, cfg, seed, pair_table)
        e0 = sp.write_snapshot(0)
        res = sp.sawtooth_anneal(e0, points_cap=cfg.num_points() + 4, snap_cap=cfg.numTeeth + 2, trace_cap=trace_moves)
        assert summ["done"][i] == 1 and summ["moves"][i] == res.moves
        assert summ["evaluated"][i] == res.evaluated
        # executed pair evaluations: eStart + eEnd like the reference, or eEnd only (crew2 takes eStart from its pair cache)
        assert summ["pair_evals"][i] in (res.pair_evals, res.pair_evals // 2)
        assert summ["points"][i] == len(res.points)
        p = pts[i, : len(res.points)]
        for key in ("tooth", "point", "accepted", "total", "movie_written", "finale"):
            assert np.array_equal(p[key], res.points[key]), (key, p[key], res.points[key])
        assert np.array_equal(p["temp"], res.points["temp"])
        assert np.max(rel_err(p["total_energy"], res.points["totalEnergy"])) <= 1e-10
        assert np.max(rel_err(p["total_squared_energy"], res.points["totalSquaredEnergy"])) <= 1e-10
        mw = res.points["movie_written"] == 1
        if mw.any():
            assert np.max(rel_err(p["movie_energy"][mw], res.points["movieEnergy"][mw])) <= 1e-10
        n_snap = len(res.snap_tooth)
        assert summ["snapshots"][i] == n_snap + 1
        assert list(st[i, : n_snap + 1]) == [0] + list(res.snap_tooth)
        assert np.max(rel_err(se[i, 1 : n_snap + 1], res.snap_energy)) <= 1e-10
        assert np.max(np.abs(sx[i, 1 : n_snap + 1] - res.snap_coords)) <= 1e-9
        osites, ocom = sp.coords()
        assert np.max(np.abs(sites[i] - osites)) <= 1e-9 and np.max(np.abs(com[i] - ocom)) <= 1e-9
        emin, tmin = sp.min_energy()
        assert mins["tooth"][i] == tmin and abs(mins["energy"][i] - emin) <= 1e-10 * abs(emin)
        if tr is not None:
            first, _, err_cond = compare_traces(tr[i], res.trace, res.trace_abs)
            assert first == -1, f"chain {i}: decisions diverge at step {first}"
            assert err_cond <= TOL


@pytest.mark.parametrize("name", ["sample8", "h2o20"])
def test_finale_heat_capacity_and_growing_teeth(engine, dbase, sample_cfg, name):
    """0 K finale (steps / 100, magwalk off, repeated tooth index, no Output file before it), the `continue` of
    Main.java:228 with heat capacities on, and Points Added per Tooth."""
    cfg = workload_config(sample_cfg, name, movePerPt=300, ptsPerTooth=4, ptsAddedPerTooth=1, numTeeth=3, finale=True,
                          writeConfHeatCapacitiesEnabled=True, maxTemp=800.0)
    seeds = [7, 8, 9, 10, 11]
    engine.set_system(system_from_config(dbase, cfg))
    engine.set_schedules(cfg)
    engine.set_options(trace_moves=cfg.total_moves())
    engine.init_chains_seeded(seeds, cfg.spaceLen, cfg.maxPropFailures)
    engine.run(0)
    _check_records(engine, dbase, cfg, seeds, trace_moves=cfg.total_moves())


def test_static_temperature_is_the_energies_txt_stream(engine, dbase, sample_cfg):
    """Static Temperature (Main.java:74-80): one tooth, one point, T constant; the trace's `running` column is what
    writeEnergy logs per move, and the point sums skip the first Equilibration Configurations moves (z > eqConfigs)."""
    cfg = workload_config(sample_cfg, "h2o20", movePerPt=3000, staticTemp=True, eqConfigs=500, maxTemp=250.0)
    seeds = [21, 22, 23]
    engine.set_system(system_from_config(dbase, cfg))
    engine.set_schedules(cfg)
    engine.set_options(trace_moves=3000)
    engine.init_chains_seeded(seeds, cfg.spaceLen, cfg.maxPropFailures)
    engine.run(0)
    tr, pts, summ = engine.trace(), engine.points(), engine.summary()
    for i, seed in enumerate(seeds):
        sp = oracle_start(dbase, cfg, seed)
        res = sp.sawtooth_anneal(sp.calc_energy(), points_cap=4, snap_cap=4, trace_cap=3000)
        first, _, err_cond = compare_traces(tr[i], res.trace, res.trace_abs)
        assert first == -1 and err_cond <= TOL
        assert summ["points"][i] == 1 and np.all(tr[i]["temp"] == 250.0)
        assert np.max(rel_err(tr[i]["running"], res.trace["running"])) <= 1e-11
        assert abs(pts[i, 0]["total_energy"] - float(np.sum(tr[i]["running"][501:]))) <= 1e-9 * abs(pts[i, 0]["total_energy"])
        assert abs(pts[i, 0]["total_energy"] - res.points["totalEnergy"][0]) <= 1e-10 * abs(res.points["totalEnergy"][0])


def test_born_mayer_and_ghost_sites(engine, golden_dir, sample_cfg):
    """The A*exp(-B*r) term (Atom.java:103) and a ghost site (in the energy and the box test, out of the centre of
    mass): total / delta energies and a decision trace on a synthetic two-species cluster."""
    dbase = read_dbase(os.path.join(golden_dir, "born_mayer.dbase.txt"))
    cfg = sample_cfg.copy(particles=[("XY", 9), ("ZG", 7)], spaceLen=9.0, movePerPt=1500, ptsPerTooth=4, numTeeth=2, maxTemp=900.0)
    system = system_from_config(dbase, cfg)
    seeds = [3, 4, 5]
    engine.set_system(system)
    engine.set_schedules(cfg)
    engine.set_options(trace_moves=cfg.total_moves())
    engine.init_chains_seeded(seeds, cfg.spaceLen, cfg.maxPropFailures)
    engine.run(0)
    _check_records(engine, dbase, cfg, seeds, trace_moves=cfg.total_moves())
    sp = oracle_start(dbase, cfg, 3)
    x, _ = sp.coords()
    e, e_abs = sp.calc_energy_abs()
    assert abs(engine.total_energy(x[None])[0] - e) <= TOL * e_abs


def test_parameter_sweep_over_schedules(engine, dbase, sample_cfg):
    """BASELINE config #3 in miniature: chains = sawtooth peaks x magwalk probabilities x seeds, one launch, each chain
    against the oracle run with its own parameter set."""
    base = workload_config(sample_cfg, "nh4cl32", movePerPt=400, ptsPerTooth=3, numTeeth=2)
    scheds, idx, seeds = [], [], []
    for ti, tmax in enumerate((2000.0, 9000.0)):
        for pi, prob in enumerate((0.0, 0.25)):
            scheds.append(base.copy(maxTemp=tmax, magwalkProbTrans=prob, magwalkProbRot=prob))
            for s in range(2):
                idx.append(ti * 2 + pi)
                seeds.append(1000 + len(seeds))
    n_moves = base.total_moves()
    engine.set_system(system_from_config(dbase, base))
    engine.set_schedules(scheds)
    engine.set_options(trace_moves=n_moves)
    engine.init_chains_seeded(seeds, base.spaceLen, base.maxPropFailures, sched_idx=idx)
    engine.run(0)
    tr, mins = engine.trace(), engine.pack_minima()
    for i, seed in enumerate(seeds):
        cfg = scheds[idx[i]]
        sp = oracle_start(dbase, cfg, seed)
        e0 = sp.write_snapshot(0)
        res = sp.sawtooth_anneal(e0, trace_cap=n_moves, snap_cap=4)
        first, _, err_cond = compare_traces(tr[i], res.trace, res.trace_abs)
        assert first == -1 and err_cond <= TOL, (i, first, err_cond)
        assert np.array_equal(tr[i]["temp"], res.trace["temp"])
        emin, tmin = sp.min_energy()
        assert mins["tooth"][i] == tmin and abs(mins["energy"][i] - emin) <= 1e-10 * abs(emin)


def test_explicit_start_with_frozen_particle_and_imported_rng(engine, dbase, sample_cfg, golden_dir, tmp_path):
    """"Use Input.xyz" (Space.java:178-233) with a frozen particle (`1F NH4+`): never picked, still in the energy; the
    three MersenneTwister states are imported mid-stream (624 words + mti) instead of a seed."""
    text = open(os.path.join(golden_dir, "sample8.input.xyz")).read().replace("4 NH4+ | 4 Cl-", "3 NH4+ | 1F NH4+ | 4 Cl-")
    f = tmp_path / "in.xyz"
    f.write_text(text)
    system, xyz = read_input(str(f), dbase)
    cfg = sample_cfg.copy(movePerPt=2500, ptsPerTooth=3, numTeeth=2, spaceLen=12.0)
    n_moves = cfg.total_moves()
    states = []
    oracles = []
    for seed in (91, 92):
        sp = OracleSpace(dbase)
        sp.set_seed(seed)
        sp.set_config(cfg)
        sp.read_input([("NH4+", 3, False), ("NH4+", 1, True), ("Cl-", 4, False)], xyz)
        st = []
        for k in range(3):                              # three streams left mid-block (mti neither 0 nor 624)
            g = JavaMT(seed * 10 + k)
            for _ in range(37 + 100 * k + seed):
                g.next(32)
            st.append(g.state())
        sp.set_rng_state(np.stack(st))
        states.append(sp.rng_state())
        oracles.append(sp)
    engine.set_system(system)
    engine.set_schedules(cfg)
    engine.set_options(trace_moves=n_moves)
    engine.init_chains_explicit(2, xyz, cfg.spaceLen, mt_state=np.stack(states))
    assert np.array_equal(engine.rng_state(), np.stack(states))
    engine.run(0)
    tr = engine.trace()
    sites, com = engine.sites()
    for i, sp in enumerate(oracles):
        res = sp.sawtooth_anneal(sp.calc_energy(), trace_cap=n_moves)
        first, _, err_cond = compare_traces(tr[i], res.trace, res.trace_abs)
        assert first == -1 and err_cond <= TOL
        assert 3 not in set(tr[i]["mol"].tolist())      # the frozen particle is never drawn (Space.java:735-738)
        osites, ocom = sp.coords()
        assert np.max(np.abs(sites[i] - osites)) <= 1e-9
        assert np.array_equal(engine.rng_state()[i], sp.rng_state())


def test_resumed_launches_equal_one_launch(engine, dbase, sample_cfg):
    """tr_run(max_moves) stops after any move and the next launch resumes bit-identically (chain state is a POD)."""
    cfg = workload_config(sample_cfg, "h2o20", movePerPt=700, ptsPerTooth=4, numTeeth=2, finale=True)
    seeds = [1000 + i for i in range(9)]
    engine.set_system(system_from_config(dbase, cfg))
    engine.set_schedules(cfg)
    engine.set_options(trace_moves=0)

    def collect():
        s = engine.summary()
        return s, engine.points(), engine.snapshots(), engine.sites(), engine.rng_state()

    engine.init_chains_seeded(seeds, cfg.spaceLen, cfg.maxPropFailures)
    engine.run(0)
    one = collect()
    engine.init_chains_seeded(seeds, cfg.spaceLen, cfg.maxPropFailures)
    for chunk in (1, 699, 1, 1234, 17):
        engine.run(chunk)
        assert not engine.summary()["done"].any()
    engine.run(0)
    many = collect()
    assert one[0].tobytes() == many[0].tobytes() and one[1].tobytes() == many[1].tobytes()
    for a, b in zip(one[2], many[2]):
        assert np.array_equal(a, b)
    assert np.array_equal(one[3][0], many[3][0]) and np.array_equal(one[3][1], many[3][1])
    assert np.array_equal(one[4], many[4])


@pytest.mark.parametrize("n_chains", [1, 5, 37, 300])
def test_ragged_chain_counts(engine, dbase, sample_cfg, n_chains):
    """Chain counts that do not fill the last block / the last energy warp: every chain still runs and matches."""
    cfg = workload_config(sample_cfg, "h2o20", movePerPt=200, ptsPerTooth=3, numTeeth=1, maxTemp=500.0)
    seeds = [5000 + i for i in range(n_chains)]
    engine.set_system(system_from_config(dbase, cfg))
    engine.set_schedules(cfg)
    engine.set_options(trace_moves=cfg.total_moves())
    engine.init_chains_seeded(seeds, cfg.spaceLen, cfg.maxPropFailures)
    engine.run(0)
    summ, tr = engine.summary(), engine.trace()
    assert summ["done"].all() and (summ["moves"] == cfg.total_moves()).all()
    for i in sorted({0, n_chains // 2, n_chains - 1}):
        sp = oracle_start(dbase, cfg, seeds[i])
        res = sp.sawtooth_anneal(sp.calc_energy(), trace_cap=cfg.total_moves())
        first, _, err_cond = compare_traces(tr[i], res.trace, res.trace_abs)
        assert first == -1 and err_cond <= TOL


@pytest.mark.parametrize("name,tpcs", [("h2o20", (16, 32, 128)), ("mixed64", (32, 256)), ("sample8", (16, 32))])
def test_every_kernel_variant_takes_the_same_decisions(engine, dbase, sample_cfg, name, tpcs):
    """threads_per_chain = 0 (automatic: crew kernel up to 64 sites, else a team kernel), the crew kernel forced, 16 / 32
    (team kernels) and 128 / 256 (block per chain) are the same Markov chain: identical decision traces, energies
    within the bar."""
    cfg = workload_config(sample_cfg, name, movePerPt=500, ptsPerTooth=3, numTeeth=2, maxTemp=700.0)
    seeds = [300, 301, 302]
    n = cfg.total_moves()
    system = system_from_config(dbase, cfg)
    ref = None
    for tpc in (0, "crew") + tuple(tpcs):
        # "crew": the warp-specialised kernel also where the automatic choice prefers a team kernel
        engine.set_kernel(TR_KERNEL_CREW if tpc == "crew" else TR_KERNEL_AUTO)
        engine.set_system(system)
        engine.set_schedules(cfg)
        engine.set_options(trace_moves=n, threads_per_chain=0 if tpc == "crew" else tpc)
        engine.init_chains_seeded(seeds, cfg.spaceLen, cfg.maxPropFailures)
        engine.set_kernel(TR_KERNEL_AUTO)
        engine.run(0)
        tr, mins = engine.trace(), engine.pack_minima()
        if ref is None:
            ref = (tr, mins)
            sp = oracle_start(dbase, cfg, seeds[0])
            res = sp.sawtooth_anneal(sp.calc_energy(), trace_cap=n)
            assert compare_traces(tr[0], res.trace, res.trace_abs)[0] == -1
            continue
        for key in ("mol", "kind", "magwalk", "bounds", "accepted", "axis", "drew_rand"):
            assert np.array_equal(tr[key], ref[0][key]), (tpc, key)
        scale = np.maximum(np.abs(ref[0]["e_start"]), np.abs(ref[0]["e_end"])) + 1e-300
        assert np.max(np.abs(tr["dE"] - ref[0]["dE"]) / np.maximum(scale, 1.0)) <= 1e-10
        assert np.array_equal(mins["tooth"], ref[1]["tooth"])
        assert np.max(rel_err(mins["energy"], ref[1]["energy"])) <= 1e-10
    engine.set_options(trace_moves=0, threads_per_chain=0)


def test_two_gpu_sharding_is_invisible_per_chain(dbase, sample_cfg):
    """Chains are sharded by global index (seed = 1000 + global id): the union of two half-size contexts on two GPUs
    equals one context bit for bit.  Skipped on a single-GPU box."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cfg = workload_config(sample_cfg, "h2o20", movePerPt=300, ptsPerTooth=3, numTeeth=2)
    system = system_from_config(dbase, cfg)
    seeds = np.arange(1000, 1000 + 64, dtype=np.int64)

    def run(dev, lo, hi):
        with Engine(dev) as e:
            e.set_system(system)
            e.set_schedules(cfg)
            e.init_chains_seeded(seeds[lo:hi], cfg.spaceLen, cfg.maxPropFailures, first_chain=lo)
            e.run(0)
            return e.summary(), e.sites()[0], e.pack_minima()

    whole = run(0, 0, 64)
    a, b = run(0, 0, 32), run(1, 32, 64)
    assert whole[0].tobytes() == np.concatenate([a[0], b[0]]).tobytes()
    assert np.array_equal(whole[1], np.concatenate([a[1], b[1]]))
    assert whole[2].tobytes() == np.concatenate([a[2], b[2]]).tobytes()


def test_error_convention_through_the_abi(dbase, sample_cfg):
    """Non-zero status + tr_last_error text instead of exceptions across the ABI; the host turns them into the
    RuntimeException("Error: ...") the reference would have thrown (Main.java:85-87)."""
    from transrot_b200 import TransRotError

    with Engine(0) as e:
        with pytest.raises(TransRotError, match="Error: .*tr_init_chains"):
            e.run(0)                                             # no chains yet
        cfg = workload_config(sample_cfg, "h2o20")
        e.set_system(system_from_config(dbase, cfg))
        with pytest.raises(TransRotError, match="Error: .*tr_set_schedules"):
            e.init_chains_seeded([1], cfg.spaceLen, cfg.maxPropFailures)
        e.set_schedules(cfg)
        with pytest.raises(TransRotError, match="threads_per_chain"):
            e.set_options(threads_per_chain=48)
        big = sample_cfg.copy(particles=[("H2O", 600)], spaceLen=40.0)          # 1800 sites > TR_MAX_SITES
        with pytest.raises(TransRotError, match="exceed the limit"):
            e.set_system(system_from_config(dbase, big))
    with pytest.raises(TransRotError, match="out of range"):
        Engine(99)


def _untemper(y: int) -> int:
    """Inverse of the MT19937 tempering (MersenneTwister.next, commons-math3)."""
    y ^= y >> 18
    y ^= (y << 15) & 0xEFC60000
    t = y
    for _ in range(5):
        t = y ^ ((t << 7) & 0x9D2C5680)
    y = t & 0xFFFFFFFF
    t = y
    for _ in range(3):
        t = y ^ (t >> 11)
    return t & 0xFFFFFFFF


def test_abandoned_chain_is_an_error_of_the_run(engine, dbase, sample_cfg):
    """CHAIN_ABORT_RNG surfaces as TR_ERNG.  The warp-specialised kernels look a fixed number of words ahead for
    Molecule.java:68's nextInt(3) rejection loop; an R stream whose next 624 outputs are all 0xffffffff (every try rejected:
    bits = 2^31 - 1) cannot be served from that window, the chain is abandoned (done = 2) and tr_get_summary reports it instead
    of returning a result that silently differs from the reference (which would keep drawing until the block regenerates)."""
    from transrot_b200 import TransRotError

    assert _untemper(0) == 0
    cfg = workload_config(sample_cfg, "h2o20", movePerPt=300, ptsPerTooth=3, numTeeth=1)
    engine.set_system(system_from_config(dbase, cfg))
    engine.set_schedules(cfg)
    engine.set_options(trace_moves=0)
    engine.init_chains_seeded([1000, 1001], cfg.spaceLen, cfg.maxPropFailures)
    assert "crew2" in engine.kernel_name()
    sites, _ = engine.sites()
    states = engine.rng_state().copy()
    space_len = float(engine.summary()["space_len"][0])
    assert float(engine.summary()["space_len"][1]) == space_len
    stuck = states.copy()
    stuck[1, 2, :624] = _untemper(0xFFFFFFFF)          # stream R of chain 1
    stuck[1, 2, 624] = 0
    engine.init_chains_explicit(2, sites, space_len, mt_state=stuck)
    engine.run(0)
    with pytest.raises(TransRotError, match="abandoned"):
        engine.summary()
    # the same start with the untouched streams runs to the end
    engine.init_chains_explicit(2, sites, space_len, mt_state=states)
    engine.run(0)
    assert (engine.summary()["done"] == 1).all()


def test_asymmetric_pair_table_keeps_off_the_pair_cache(engine, dbase, sample_cfg):
    """The C ABI accepts a pair table with [t1][t2] != [t2][t1] (the reference never builds one).  The kernels that keep one cached
    energy per particle pair would then return whichever orientation was evaluated last: the automatic choice falls back to the
    kernels that evaluate eStart and eEnd afresh (decisions still equal to the oracle, which walks the same table), and forcing a
    pair-cache kernel is an error."""
    import dataclasses

    from transrot_b200 import TransRotError
    from transrot_b200.engine import TR_KERNEL_CREW2

    cfg = workload_config(sample_cfg, "h2o20", movePerPt=400, ptsPerTooth=3, numTeeth=1)
    system = system_from_config(dbase, cfg)
    table = system.pair_table.copy()
    assert table.shape[0] >= 2
    table[0, 1, 2] = 3.0 * table[1, 0, 2] + 1.0                  # C of (type 0 moved, type 1 other) only
    skewed = dataclasses.replace(system, pair_table=table)
    n_moves = 1200
    engine.set_system(skewed)
    engine.set_schedules(cfg)
    engine.set_options(trace_moves=n_moves)
    try:
        engine.init_chains_seeded([1000, 1001], cfg.spaceLen, cfg.maxPropFailures)
        assert "crew2" not in engine.kernel_name() and "pair cache" not in engine.kernel_name()
        engine.run(n_moves)
        engine.synchronize()
        tr = engine.trace()
        for i, seed in enumerate([1000, 1001]):
            sp = oracle_start(dbase, cfg, seed, pair_table=table)
            res = sp.sawtooth_anneal(sp.calc_energy(), trace_cap=n_moves)
            first, err_rel, err_cond = compare_traces(tr[i], res.trace, res.trace_abs)
            assert first == -1 and err_cond <= 1e-12
        engine.set_kernel(TR_KERNEL_CREW2)
        with pytest.raises(TransRotError, match="not symmetric"):
            engine.init_chains_seeded([1000], cfg.spaceLen, cfg.maxPropFailures)
    finally:
        engine.set_kernel(TR_KERNEL_AUTO)
        engine.set_options(trace_moves=0)


@pytest.mark.parametrize("tpc", [0, 16])
def test_degenerate_schedules(engine, dbase, sample_cfg, tpc):
    """Points per Tooth = 1 makes delT = t / 0 = Infinity (and 0 / 0 = NaN in the finale, where every uphill move is
    then accepted: SURVEY Q11); Moves per Point = 1 makes every move the last of its point."""
    cases = [
        workload_config(sample_cfg, "sample8", movePerPt=40, ptsPerTooth=1, numTeeth=3, finale=True, maxTemp=500.0),
        workload_config(sample_cfg, "sample8", movePerPt=1, ptsPerTooth=7, ptsAddedPerTooth=3, numTeeth=4, maxTemp=900.0),
        workload_config(sample_cfg, "h2o20", movePerPt=2, ptsPerTooth=2, numTeeth=2, finale=True, writeConfHeatCapacitiesEnabled=True),
    ]
    for cfg in cases:
        seeds = [61, 62, 63]
        n = cfg.total_moves()
        engine.set_system(system_from_config(dbase, cfg))
        engine.set_schedules(cfg)
        engine.set_options(trace_moves=n, threads_per_chain=tpc)
        engine.init_chains_seeded(seeds, cfg.spaceLen, cfg.maxPropFailures)
        engine.run(0)
        _check_records(engine, dbase, cfg, seeds, trace_moves=n)
    engine.set_options(trace_moves=0, threads_per_chain=0)


def test_long_run_decisions_identical(engine, dbase, sample_cfg):
    """The stated number of initial steps: 200 000 moves of the shipped sawtooth shape on (H2O)20 with every decision
    identical to the oracle (tools/long_parity.py: 1 000 000 moves, three seeds, also identical).  Trial geometries
    use CUDA's sin/cos instead of glibc's, so coordinates - and with them the energies - drift apart at the 1e-11
    level over such a run; the final minima still agree to 1e-12."""
    cfg = workload_config(sample_cfg, "h2o20", movePerPt=5000)
    n = cfg.total_moves()
    seeds = [1000, 1001]
    engine.set_system(system_from_config(dbase, cfg))
    engine.set_schedules(cfg)
    engine.set_options(trace_moves=n)
    engine.init_chains_seeded(seeds, cfg.spaceLen, cfg.maxPropFailures)
    engine.run(0)
    tr, mins = engine.trace(), engine.pack_minima()
    for i, seed in enumerate(seeds):
        sp = oracle_start(dbase, cfg, seed)
        e0 = sp.write_snapshot(0)
        res = sp.sawtooth_anneal(e0, trace_cap=n, snap_cap=8)
        first, _, err_cond = compare_traces(tr[i], res.trace, res.trace_abs)
        assert first == -1, f"seed {seed}: decisions diverge at move {first} of {n}"
        assert err_cond <= 1e-10
        emin, tmin = sp.min_energy()
        assert mins["tooth"][i] == tmin and abs(mins["energy"][i] - emin) <= 1e-12 * abs(emin)
    engine.set_options(trace_moves=0)


@pytest.mark.parametrize("tpc", [0, 32])
def test_movie_frames(engine, dbase, sample_cfg, tpc):
    """The structure Space.writeMovie appends after every point (Space.java:740-790): its energy is the point record's
    movie_energy, the last frame of a tooth is the write(x+1) structure, skipped frames (Main.java:228) stay empty."""
    from transrot_b200 import format_movie_frame

    cfg = workload_config(sample_cfg, "sample8", movePerPt=400, ptsPerTooth=4, numTeeth=2, finale=True,
                          writeConfHeatCapacitiesEnabled=True, maxTemp=700.0)
    system = system_from_config(dbase, cfg)
    seeds = [11, 12, 13]
    engine.set_system(system)
    engine.set_schedules(cfg)
    engine.set_options(trace_moves=0, threads_per_chain=tpc)
    engine.set_record_movie(True)
    engine.init_chains_seeded(seeds, cfg.spaceLen, cfg.maxPropFailures)
    engine.run(0)
    pts, frames = engine.points(), engine.movie()
    se, st, sx = engine.snapshots()
    summ = engine.summary()
    for i, seed in enumerate(seeds):
        n = int(summ["points"][i])
        p = pts[i, :n]
        written = p["movie_written"] == 1
        assert written.any() and (~written).any()
        e = engine.total_energy(frames[i, :n][written])
        assert np.max(rel_err(e, p["movie_energy"][written])) <= 1e-12
        assert not frames[i, :n][~written].any()
        # last point of tooth 0 (not the finale): the structure write(1) saw
        last0 = int(np.nonzero((p["tooth"] == 0) & (p["finale"] == 0))[0][-1])
        if written[last0]:
            assert np.array_equal(frames[i, last0], sx[i, 1])
        sp = oracle_start(dbase, cfg, seed)
        res = sp.sawtooth_anneal(sp.write_snapshot(0), points_cap=n + 2, snap_cap=4)
        assert np.max(rel_err(p["movie_energy"][written], res.points["movieEnergy"][written])) <= 1e-10
    text = format_movie_frame(system, frames[0, 0], float(pts[0, 0]["movie_energy"]))
    assert text.split("\n")[0].strip() == "24" and text.count("\n") == 2 + 24
    engine.set_record_movie(False)
    engine.set_options(trace_moves=0, threads_per_chain=0)


def test_replayed_output_directory(engine, dbase, sample_cfg, tmp_path):
    """Device buffers -> the reference's output directory (Output0..N.xyz, movies, acceptance ratios, minimum copy);
    every OutputN.xyz restarts through read_input and reproduces its header energy on the device."""
    from transrot_b200 import replay_outputs

    cfg = workload_config(sample_cfg, "sample8", movePerPt=300, ptsPerTooth=3, numTeeth=3, writeAcceptanceRatios=True)
    system = system_from_config(dbase, cfg)
    engine.set_system(system)
    engine.set_schedules(cfg)
    engine.set_options(trace_moves=0)
    engine.set_record_movie(True)
    engine.init_chains_seeded([42], cfg.spaceLen, cfg.maxPropFailures)
    engine.run(0)
    n = int(engine.summary()["points"][0])
    se, st, sx = engine.snapshots()
    mins = engine.pack_minima()
    out = tmp_path / "run"
    replay_outputs(str(out), system, cfg, engine.points()[0, :n], se[0], st[0], sx[0], engine.movie()[0], int(mins["tooth"][0]))
    names = sorted(os.listdir(out))
    assert [f"Output{k}.xyz" for k in range(4)] == [x for x in names if x.startswith("Output") and "_" not in x]
    assert {"Output0_1_Movie.xyz", "Output1_2_Movie.xyz", "Output2_3_Movie.xyz", "acceptance_ratios.txt"} <= set(names)
    assert f"Min_Energy_Structure_{int(mins['tooth'][0])}.xyz" in names
    assert len(open(out / "acceptance_ratios.txt").read().splitlines()) == 9
    # the shipped annealing block has heat capacities on: the t = 0 point of each tooth hits the `continue` of
    # Main.java:228 (no heat-capacity line, no movie frame)
    assert cfg.writeConfHeatCapacitiesEnabled
    assert open(out / "Output0_1_Movie.xyz").read().count("Energy:") == 2
    assert len(open(out / "configurational_heat_capacities.txt").read().splitlines()) == 6
    for k in range(4):
        sys_k, xyz = read_input(str(out / f"Output{k}.xyz"), dbase)
        header = float(open(out / f"Output{k}.xyz").read().split("\n")[1].split()[1])
        # reference quirk kept: Double.toString switches to "4.37E-4" below 1e-3 and Space.write splits it on "."
        # (Space.java:582-591), so such a coordinate is written as "4.371995E-40" and restarts as ~0
        tiny = np.abs(sx[0, k]) < 1e-3
        assert header == se[0, k] and np.max(np.abs(xyz - sx[0, k])[~tiny]) <= 5.1e-11 and np.all(np.abs(xyz[tiny]) < 1e-30)
        # the file rounds coordinates to 10 decimals: the restart energy moves by ~1e-8 relative at most
        if not tiny.any():
            assert abs(engine.total_energy(xyz[None])[0] - header) <= 1e-7 * abs(header)
    engine.set_system(system)
    engine.set_record_movie(False)
