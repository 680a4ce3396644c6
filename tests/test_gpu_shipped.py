"""Parity of the kernels that SHIP AND ARE TIMED: `trace_moves = 0`, automatic kernel choice, i.e. exactly the template
instantiations bench.py launches (the decision-trace tests of test_gpu_parity.py run the generic tracing instantiations).
Every BASELINE configuration anneals a whole (shortened) sawtooth schedule on the device and every record a host can
read back - per-point counters and energy sums, movie energies, every write(n) snapshot with its structure, the tracked
minimum, the final structure and centres of mass, the three MersenneTwister states - is compared with the oracle.

Bars: integer records and RNG states bit-exact (an accept/reject decision that differed anywhere in the run would change
the acceptance counters and the final RNG state); energies 1e-10 relative over a whole run (the per-evaluation bar is
1e-12 of the summed |pair terms|, test_gpu_parity.py; over tens of thousands of committed moves the trial geometries'
sin/cos - CUDA libm vs glibc - let coordinates drift apart at the 1e-11 level, profiles/parity_r2.md).
"""
import numpy as np
import pytest

from helpers import oracle_start, rel_err, workload_config
from transrot_b200 import sweep_schedules, system_from_config

pytestmark = pytest.mark.gpu


def _compare_whole_run(engine, dbase, cfg, seeds, sched_cfgs=None, sched_idx=None):
    summ, pts = engine.summary(), engine.points()
    se, st, sx = engine.snapshots()
    sites, com = engine.sites()
    mins = engine.pack_minima()
    rng = engine.rng_state()
    ms = engine.min_structures()
    worst = 0.0
    for i, seed in enumerate(seeds):
        c = cfg if sched_cfgs is None else sched_cfgs[sched_idx[i]]
        sp = oracle_start(dbase, c, seed)
        e0 = sp.write_snapshot(0)
        res = sp.sawtooth_anneal(e0, points_cap=c.num_points() + 4, snap_cap=c.numTeeth + 2)
        # ---- integers: bit-exact
        assert summ["done"][i] == 1 and summ["moves"][i] == res.moves == c.total_moves()
        assert summ["evaluated"][i] == res.evaluated
        assert summ["pair_evals"][i] in (res.pair_evals, res.pair_evals // 2), "executed pair evaluations: all of eStart+eEnd, or eEnd only"
        assert summ["points"][i] == len(res.points)
        p = pts[i, : len(res.points)]
        for key in ("tooth", "point", "accepted", "total", "movie_written", "finale"):
            assert np.array_equal(p[key], res.points[key]), (i, key, p[key], res.points[key])
        assert np.array_equal(p["temp"], res.points["temp"])
        assert np.array_equal(rng[i], sp.rng_state()), f"chain {i}: RNG states differ - some draw or decision differed"
        n_snap = len(res.snap_tooth)
        assert summ["snapshots"][i] == n_snap + 1 and list(st[i, : n_snap + 1]) == [0] + list(res.snap_tooth)
        emin, tmin = sp.min_energy()
        assert mins["tooth"][i] == tmin
        # ---- energies and structures
        errs = [
            np.max(rel_err(p["total_energy"], res.points["totalEnergy"])),
            np.max(rel_err(p["total_squared_energy"], res.points["totalSquaredEnergy"])),
            abs(se[i, 0] - e0) / abs(e0),
            np.max(rel_err(se[i, 1 : n_snap + 1], res.snap_energy)),
            abs(mins["energy"][i] - emin) / abs(emin),
            abs(summ["running_energy"][i] - res.final_energy) / abs(res.final_energy),
        ]
        mw = res.points["movie_written"] == 1
        if mw.any():
            errs.append(np.max(rel_err(p["movie_energy"][mw], res.points["movieEnergy"][mw])))
        worst = max(worst, float(max(errs)))
        assert max(errs) <= 1e-10, (i, errs)
        assert np.max(np.abs(sx[i, 1 : n_snap + 1] - res.snap_coords)) <= 1e-9
        osites, ocom = sp.coords()
        assert np.max(np.abs(sites[i] - osites)) <= 1e-9 and np.max(np.abs(com[i] - ocom)) <= 1e-9
        # the packed minimum structure is the snapshot of the minimum tooth
        k = [0] + list(res.snap_tooth)
        assert np.array_equal(ms[i], sx[i, k.index(tmin)])
    return worst


@pytest.mark.parametrize(
    "name,n_chains,mpp,teeth",
    [("sample8", 6, 3000, 3), ("h2o20", 6, 2500, 3), ("nh4cl32", 3, 2000, 2), ("mixed64", 3, 2000, 2), ("h2o256", 2, 300, 2)],
)
def test_shipped_kernel_whole_run(engine, dbase, sample_cfg, name, n_chains, mpp, teeth):
    cfg = workload_config(sample_cfg, name, movePerPt=mpp, ptsPerTooth=5, numTeeth=teeth, finale=(name == "h2o20"))
    seeds = [1000 + i for i in range(n_chains)]
    engine.set_system(system_from_config(dbase, cfg))
    engine.set_schedules(cfg)
    engine.set_options(trace_moves=0)                       # the non-tracing instantiations: what bench.py times
    engine.init_chains_seeded(seeds, cfg.spaceLen, cfg.maxPropFailures)
    engine.run(0)
    _compare_whole_run(engine, dbase, cfg, seeds)


KERNEL_PREFIX = {1: "anneal_kernel<L=32", 3: "crew2_kernel<L=32", 4: "anneal_kernel<L=32"}


@pytest.mark.parametrize("kernel", [1, 3, 4], ids=["team", "crew2", "team_pc"])
@pytest.mark.parametrize("name,n_chains,mpp,per_block", [("nh4cl32", 3, 2000, 0), ("mixed64", 30, 600, 13), ("h2o256", 2, 300, 0)])
def test_global_pair_cache_kernels_whole_run(engine, dbase, sample_cfg, name, n_chains, mpp, per_block, kernel):
    """Every kernel family that can run 33 and more particles, forced with tr_set_kernel whatever the automatic choice is: the
    plain team kernel (TR_KERNEL_TEAM), and the two that keep the pair-energy cache in global memory - the crew2 variants with
    whole-warp teams (TR_KERNEL_CREW2) and the team kernel on particle-major positions (TR_KERNEL_TEAM_PC).  Same whole-run
    comparison; mixed64 also with many chains per block."""
    from transrot_b200.engine import TR_KERNEL_AUTO
    TR_KERNEL_CREW2 = kernel

    cfg = workload_config(sample_cfg, name, movePerPt=mpp, ptsPerTooth=5, numTeeth=2)
    seeds = [1000 + i for i in range(n_chains)]
    engine.set_system(system_from_config(dbase, cfg))
    engine.set_schedules(cfg)
    engine.set_options(trace_moves=0)
    engine.set_kernel(TR_KERNEL_CREW2, per_block)
    try:
        engine.init_chains_seeded(seeds, cfg.spaceLen, cfg.maxPropFailures)
        assert engine.kernel_name().startswith(KERNEL_PREFIX[kernel]) and (kernel == 3 or ("pair cache" in engine.kernel_name()) == (kernel == 4))
        engine.run(0)
        if n_chains <= 4:
            _compare_whole_run(engine, dbase, cfg, seeds)
        else:
            summ, pts, rng = engine.summary(), engine.points(), engine.rng_state()
            assert summ["done"].all() and (summ["moves"] == cfg.total_moves()).all()
            for i in (0, 12, 13, 29):
                sp = oracle_start(dbase, cfg, seeds[i])
                res = sp.sawtooth_anneal(sp.write_snapshot(0), points_cap=cfg.num_points() + 2, snap_cap=4)
                p = pts[i, : len(res.points)]
                assert np.array_equal(p["accepted"], res.points["accepted"]) and np.array_equal(p["total"], res.points["total"])
                assert np.array_equal(rng[i], sp.rng_state())
    finally:
        engine.set_kernel(TR_KERNEL_AUTO)


@pytest.mark.parametrize("kernel", [1, 3, 4], ids=["team", "crew2", "team_pc"])
@pytest.mark.parametrize("name,n_chains,n_moves", [("nh4cl32", 2, 6000), ("mixed64", 2, 6000), ("h2o256", 2, 1500)])
def test_global_pair_cache_kernels_decision_trace(engine, dbase, sample_cfg, name, n_chains, n_moves, kernel):
    """Per-move decision trace of the same kernels (eStart from the global-memory pair cache) against the oracle."""
    from helpers import compare_traces
    from transrot_b200.engine import TR_KERNEL_AUTO
    TR_KERNEL_CREW2 = kernel

    cfg = workload_config(sample_cfg, name, movePerPt=max(n_moves // 4, 1), ptsPerTooth=5, numTeeth=2, maxTemp=600.0)
    seeds = [1000 + i for i in range(n_chains)]
    engine.set_system(system_from_config(dbase, cfg))
    engine.set_schedules(cfg)
    engine.set_options(trace_moves=n_moves)
    engine.set_kernel(TR_KERNEL_CREW2)
    try:
        engine.init_chains_seeded(seeds, cfg.spaceLen, cfg.maxPropFailures)
        engine.run(n_moves)
        engine.synchronize()
        tr = engine.trace()
    finally:
        engine.set_kernel(TR_KERNEL_AUTO)
        engine.set_options(trace_moves=0)
    for i, seed in enumerate(seeds):
        sp = oracle_start(dbase, cfg, seed)
        res = sp.sawtooth_anneal(sp.calc_energy(), trace_cap=n_moves)
        first, err_rel, err_cond = compare_traces(tr[i], res.trace, res.trace_abs)
        assert first == -1, f"chain {i}: decisions diverge at step {first}: gpu={tr[i][first]} oracle={res.trace[first]}"
        assert err_cond <= 1e-12, err_cond


def test_shipped_kernel_full_block_of_chains(engine, dbase, sample_cfg):
    """The launch geometry of the headline run (many chains per block, every energy team busy, several blocks): 300
    (H2O)20 chains, a sample of them against the oracle, all of them against the acceptance statistics' sanity."""
    cfg = workload_config(sample_cfg, "h2o20", movePerPt=600, ptsPerTooth=4, numTeeth=2)
    seeds = [1000 + i for i in range(300)]
    engine.set_system(system_from_config(dbase, cfg))
    engine.set_schedules(cfg)
    engine.set_options(trace_moves=0)
    engine.set_kernel(0, 28)          # full blocks, as 4096 chains have them (300 chains alone would spread as 3 per block)
    try:
        engine.init_chains_seeded(seeds, cfg.spaceLen, cfg.maxPropFailures)
        engine.run(0)
    finally:
        engine.set_kernel(0)
    summ, pts, rng, mins = engine.summary(), engine.points(), engine.rng_state(), engine.pack_minima()
    assert summ["done"].all() and (summ["moves"] == cfg.total_moves()).all()
    for i in (0, 1, 13, 14, 27, 28, 149, 298, 299):
        sp = oracle_start(dbase, cfg, seeds[i])
        res = sp.sawtooth_anneal(sp.write_snapshot(0), points_cap=cfg.num_points() + 2, snap_cap=4)
        p = pts[i, : len(res.points)]
        assert np.array_equal(p["accepted"], res.points["accepted"]) and np.array_equal(p["total"], res.points["total"])
        assert np.array_equal(rng[i], sp.rng_state())
        emin, tmin = sp.min_energy()
        assert mins["tooth"][i] == tmin and abs(mins["energy"][i] - emin) <= 1e-10 * abs(emin)


def test_shipped_kernel_parameter_sweep(engine, dbase, sample_cfg):
    """BASELINE config #3 as stated: (NH4Cl)32, chains = 8 sawtooth peaks x 8 magwalk probabilities x seeds (here 1 seed
    per set, shortened schedule), non-tracing kernel, a sample of parameter sets against the oracle."""
    base = workload_config(sample_cfg, "nh4cl32", movePerPt=300, ptsPerTooth=3, numTeeth=2)
    scheds = sweep_schedules(base)
    assert len(scheds) == 64
    idx = list(range(64))
    seeds = [1000 + i for i in range(64)]
    engine.set_system(system_from_config(dbase, base))
    engine.set_schedules(scheds)
    engine.set_options(trace_moves=0)
    engine.init_chains_seeded(seeds, base.spaceLen, base.maxPropFailures, sched_idx=idx)
    engine.run(0)
    summ, pts, rng = engine.summary(), engine.points(), engine.rng_state()
    assert summ["done"].all()
    for i in (0, 7, 9, 36, 63):
        sp = oracle_start(dbase, scheds[i], seeds[i])
        res = sp.sawtooth_anneal(sp.write_snapshot(0), points_cap=base.num_points() + 2, snap_cap=4)
        p = pts[i, : len(res.points)]
        assert np.array_equal(p["temp"], res.points["temp"])
        assert np.array_equal(p["accepted"], res.points["accepted"])
        assert np.array_equal(rng[i], sp.rng_state())
        assert np.max(rel_err(p["total_energy"], res.points["totalEnergy"])) <= 1e-10


def test_multi_engine_on_available_gpus(dbase, sample_cfg):
    """tr_create_multi / tr_gather_minima / tr_gather_min_structures (csrc/multi.cu): all visible GPUs driven from this
    one thread; per-chain results equal a single context's bit for bit, the NCCL gathers return every record in global
    chain order and the winner with its structure.  With one GPU the same calls run without NCCL."""
    import torch

    from transrot_b200 import Engine, MultiEngine

    n_dev = min(torch.cuda.device_count(), 8)
    cfg = workload_config(sample_cfg, "h2o20", movePerPt=300, ptsPerTooth=3, numTeeth=2)
    system = system_from_config(dbase, cfg)
    n = 67                                                   # ragged over 2, 4 and 8 devices
    seeds = np.arange(1000, 1000 + n, dtype=np.int64)
    with Engine(0) as e:
        e.set_system(system)
        e.set_schedules(cfg)
        e.init_chains_seeded(seeds, cfg.spaceLen, cfg.maxPropFailures)
        e.run(0)
        ref_min, ref_struct, ref_summ = e.pack_minima(), e.min_structures(), e.summary()
    with MultiEngine(list(range(n_dev))) as m:
        m.set_system(system)
        m.set_schedules(cfg)
        m.set_options()
        m.init_chains_seeded(seeds, cfg.spaceLen, cfg.maxPropFailures)
        m.run(0)
        m.synchronize()
        allm, best, x = m.gather_minima()
        structs = m.gather_min_structures()
        parts = [m.ctx(i).summary() for i in range(n_dev)]
    assert allm.tobytes() == ref_min.tobytes()
    assert np.array_equal(structs, ref_struct)
    assert np.concatenate(parts).tobytes() == ref_summ.tobytes()
    w = int(np.argmin(ref_min["energy"]))
    assert best["chain"] == w and best["energy"] == ref_min["energy"][w] and best["tooth"] == ref_min["tooth"][w]
    assert np.array_equal(x, ref_struct[w])
