#!/usr/bin/env python
"""bench.py — MC move attempts/sec of the annealing hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference algorithm on the host cores

Default workload (BASELINE.json configs[1]): (H2O)_20 TIP3P, 4096 independent seeded chains per
GPU, the shipped annealing block (10000 K, 4 teeth x 10 points x 20000 moves = 800 000 move attempts per
chain), seed = 1000 + global chain index, initial structures from the device-side Space.propagate.
`--workload` selects the other BASELINE configurations: sample8 (#1), nh4cl32 (#3: the 8 x 8 annealing-parameter sweep
x 64 seeds), mixed64 (#4: with the gather of every chain's minimum structure), h2o256 (#5).
A step = one whole annealing job over all chains: `value` times the anneal kernel with the chain state
already resident in HBM (CUDA events on the launch stream); `e2e` times the public call sequence a host
makes — seeds up from host memory, propagate, anneal, per-chain minima / summaries / point records / write(n)
structures and the winning structure back to host memory.  Under torchrun (N > 1) chains shard over ranks with no
data-path collective; the only exchange is the final NCCL all-gather of per-chain minima (and, for mixed64, of the
minimum structures), inside the e2e region.  After the timed loops a few of the chains that were just timed are re-run
through the CPU oracle (untimed) and compared: "parity_check".
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "MC move attempts/sec"
UNIT = "moves/s"
FLOPS_PER_PAIR = 19.0          # A = 0 for every shipped species
# Chains per GPU.  BASELINE.json fixes 4096 for (H2O)20 (and the shipped sample is run the same way); for the other
# configurations the count is a whole number of waves of the kernel the library picks on a 148-SM B200 - a partial last wave
# only measures idle SMs ((NH4Cl)32: 4096 chains = 1.73 waves of 16 chains per SM: 594 M moves/s; 4736 = 2 waves: 645):
#   nh4cl32: 2 waves x 148 SMs x 16 chains = 4736 = 64 parameter sets x 74 seeds;  mixed64: 2 x 148 x 12 = 3552;
#   h2o256: 2 x 148 x 3 = 888.
CHAINS_PER_GPU = {"sample8": 4096, "h2o20": 4096, "nh4cl32": 4736, "mixed64": 3552, "h2o256": 888}
SWEEP_SEEDS = 74               # config #3: 8 sawtooth peaks x 8 magwalk probabilities x 74 seeds = 4736 chains


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="h2o20", choices=["sample8", "h2o20", "nh4cl32", "mixed64", "h2o256"])
    ap.add_argument("--chains-per-gpu", type=int, default=0)
    ap.add_argument("--moves-per-point", type=int, default=0, help="override Moves per Point (default: shipped 20000)")
    ap.add_argument("--threads-per-chain", type=int, default=0)
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 team kernel, 2 crew kernel, 3 crew2 kernel (tr_set_kernel)")
    ap.add_argument("--chains-per-block", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--parity-chains", type=int, default=4, help="chains re-run through the oracle after the timed loops (0 = skip)")
    return ap.parse_args()


def load_problem(args):
    """(dbase, base config of the workload, list of parameter sets, system).  One parameter set except for nh4cl32."""
    from transrot_b200 import Config, read_dbase, sweep_schedules, system_from_config, workload_config

    golden = os.path.join(ROOT, "tests", "golden")
    dbase = read_dbase(os.path.join(golden, "shipped_species.dbase.txt"))
    base = Config.parseConfig(os.path.join(golden, "sample8.config.txt"))
    # timing runs: every Write* option / Static Temperature / Finale off
    kw = dict(writeEnergiesEnabled=False, writeConfHeatCapacitiesEnabled=False, writeAcceptanceRatios=False, finale=False,
              staticTemp=False)
    if args.moves_per_point > 0:
        kw["movePerPt"] = args.moves_per_point
    cfg = workload_config(base, args.workload, **kw)
    scheds = sweep_schedules(cfg) if args.workload == "nh4cl32" else [cfg]
    return dbase, cfg, scheds, system_from_config(dbase, cfg)


def chains_per_gpu(args):
    return args.chains_per_gpu or CHAINS_PER_GPU[args.workload]


def sched_index(args, global_chain: np.ndarray, n_sets: int) -> np.ndarray:
    """Parameter set of a chain by its GLOBAL id: consecutive blocks of SWEEP_SEEDS seeds share a set."""
    return ((global_chain // SWEEP_SEEDS) % n_sets).astype(np.int32)


def jdk_probe():
    """What the reference arm would need to run the REAL reference: a JDK and the reference sources."""
    java, javac = shutil.which("java"), shutil.which("javac")
    ref = os.environ.get("TRANSROT_REF", "/root/reference")
    have_src = os.path.exists(os.path.join(ref, "src", "main", "Main.java"))
    version = None
    if java:
        try:
            version = subprocess.run([java, "-version"], capture_output=True, text=True, timeout=20).stderr.splitlines()[0]
        except Exception:
            version = "unknown"
    return {"java": java, "javac": javac, "reference_sources": ref if have_src else None, "jdk_version": version,
            "looked_for": "java and javac on PATH; src/main/Main.java under $TRANSROT_REF or /root/reference"}


# ------------------------------------------------------------------------------------------ CPU arm


def cpu_port_rate(dbase, cfg, seconds: float, first_seed: int = 1000):
    """The oracle (C restatement of the reference algorithm) on every host core: one chain per thread
    (ctypes releases the GIL), each a bounded sample of the same workload.  Returns (moves/s, cores, sample)."""
    from oracle.oracle import oracle_from_config

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # calibrate one core on a tiny slice, then size the sample to ~`seconds`
    probe = cfg.copy(movePerPt=200, ptsPerTooth=2, numTeeth=1)
    sp = oracle_from_config(dbase, probe, first_seed)
    t0 = time.perf_counter()
    sp.sawtooth_anneal(sp.calc_energy())
    per_move = (time.perf_counter() - t0) / probe.total_moves()
    pts = cfg.num_points()
    mpp = max(1, min(cfg.movePerPt, int(seconds / per_move / pts)))
    sample_cfg = cfg.copy(movePerPt=mpp)
    spaces = []
    for i in range(cores):
        s = oracle_from_config(dbase, sample_cfg, first_seed + i)
        spaces.append((s, s.calc_energy()))
    done = [0] * cores

    def work(i):
        s, e0 = spaces[i]
        done[i] = s.sawtooth_anneal(e0).moves

    th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    sample = (f"{cores} chains (one per core), full sawtooth schedule with Moves per Point = {mpp} "
              f"({sample_cfg.total_moves()} moves per chain), {dt:.1f} s wall")
    return sum(done) / dt, cores, sample, dt


def jvm_rate(probe, cfg, seconds: float):
    """The REAL reference: P concurrent JVM instances (README.md:12,84-90), one per host core, each a bounded sample of
    the workload.  Only reachable when jdk_probe() found java + javac + the reference sources; never the case in this
    pool's images (kept so that the first box with a JDK reports kind = "jvm" without a code change)."""
    import tempfile

    ref = probe["reference_sources"]
    jar = os.path.join(ref, "src", "resources", "commons-math3-3.6.1.jar")
    out = os.path.join(ROOT, "oracle", "_ref")
    classes = os.path.join(out, "classes")
    os.makedirs(classes, exist_ok=True)
    if not os.path.exists(os.path.join(classes, "main", "Main.class")):
        subprocess.check_call([probe["javac"], "-cp", jar, "-d", classes] + sorted(
            os.path.join(ref, "src", "main", f) for f in os.listdir(os.path.join(ref, "src", "main")) if f.endswith(".java")))
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    mpp = max(50, min(cfg.movePerPt, int(seconds * 25000 / cfg.num_points())))      # ~25 k moves/s/core is SURVEY §6's estimate
    sample_cfg = cfg.copy(movePerPt=mpp)
    work = tempfile.mkdtemp(prefix="jvm_arm_")
    open(os.path.join(work, "config.txt"), "w").write(sample_cfg.to_text())
    shutil.copyfile(os.path.join(ref, "src", "dbase.txt"), os.path.join(work, "dbase.txt"))
    procs = []
    t0 = time.perf_counter()
    for i in range(cores):
        o = os.path.join(work, f"out{i}")
        os.makedirs(o)
        procs.append(subprocess.Popen([probe["java"], "-cp", f"{classes}:{jar}", "main.Main", "-c", os.path.join(work, "config.txt"), "-d",
                                       os.path.join(work, "dbase.txt"), "-o", o, "-s", str(1000 + i)], stdout=subprocess.DEVNULL,
                                      stderr=subprocess.DEVNULL))
    for p in procs:
        p.wait()
    dt = time.perf_counter() - t0
    shutil.rmtree(work, ignore_errors=True)
    sample = (f"{cores} JVM instances (one per core, {probe['jdk_version']}), full sawtooth schedule with Moves per Point = {mpp} "
              f"({sample_cfg.total_moves()} moves per chain), {dt:.1f} s wall incl. JVM start")
    return cores * sample_cfg.total_moves() / dt, cores, sample, dt


def run_reference(args):
    """`--impl reference`: the reference's CPU algorithm for the path, timed on the host cores: the real JVM program when
    a JDK and the reference sources are present (kind "jvm"), else the oracle port (kind "port": a C restatement, faster
    than the JVM original - a conservative baseline).  The probe result is part of the line."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dbase, cfg, scheds, system = load_problem(args)
    probe = jdk_probe()
    use_jvm = bool(probe["java"] and probe["javac"] and probe["reference_sources"])
    rates, secs = [], []
    sample = ""
    cores = 1
    per_step = max(2.0, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        if use_jvm:
            r, cores, sample, dt = jvm_rate(probe, cfg, per_step)
        else:
            r, cores, sample, dt = cpu_port_rate(dbase, cfg, per_step, first_seed=1000)
        if i >= args.warmup:
            rates.append(r)
            secs.append(dt)
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_descr(args, cfg, scheds, system, chains_per_gpu(args)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "jvm" if use_jvm else "port", "sample": sample},
        "jdk_probe": probe,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for ln in out.splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # under load = the upper half of the samples (the sampler also sees the untimed gaps)
        s = sorted(sm)
        return {"sm_mhz": float(np.median(s[len(s) // 2:])), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def workload_descr(args, cfg, scheds, system, cpg):
    sweep = ""
    if len(scheds) > 1:
        sweep = (f"; annealing-parameter sweep: {len(scheds)} sets (Max Temperature 2000..16000 K x magwalk probability 0..0.35) "
                 f"x {SWEEP_SEEDS} seeds, set = (global chain // {SWEEP_SEEDS}) % {len(scheds)}")
    return {
        "workload": f"{args.workload}: {system.n_mol} particles / {system.n_sites} sites, {cpg} independent seeded "
                    f"chains per GPU, sawtooth anneal {cfg.maxTemp:g} K, {cfg.numTeeth} teeth x {cfg.ptsPerTooth} points x "
                    f"{cfg.movePerPt} moves ({cfg.total_moves()} move attempts per chain){sweep}",
        "chains_per_gpu": cpg,
        "moves_per_chain": cfg.total_moves(),
        "pair_evals_per_move_ref": system.pair_evals_per_move(),
        "seed_rule": "1000 + global chain index",
        "l2": "flushed between timed steps (256 MiB write); chain state (RNG 30 MB + records) is rewritten by the untimed init",
        "parallelism": f"chains sharded over {args.gpus} GPU(s), no data-path collective; final NCCL all-gather of minima"
                       + (" and of every chain's minimum structure" if args.workload == "mixed64" else ""),
    }


def parity_check(dbase, scheds, sidx, seeds, local_ids, pts, summ, snap_e, mins, rng):
    """Re-run chains the bench has just timed through the CPU oracle (full length, one host thread each, untimed) and compare:
    per-point acceptance counters and the final MersenneTwister states (equal only if every draw and every accept/reject
    decision of the run was equal), snapshot / minimum energies."""
    from oracle.oracle import oracle_from_config

    out = [None] * len(local_ids)

    def work(k):
        i = local_ids[k]
        cfg = scheds[int(sidx[i])] if sidx is not None else scheds[0]
        sp = oracle_from_config(dbase, cfg, int(seeds[i]))
        e0 = sp.write_snapshot(0)
        res = sp.sawtooth_anneal(e0, points_cap=cfg.num_points() + 4, snap_cap=cfg.numTeeth + 2)
        p = pts[i, : len(res.points)]
        same = (int(summ["moves"][i]) == res.moves and int(summ["evaluated"][i]) == res.evaluated
                and np.array_equal(p["accepted"], res.points["accepted"]) and np.array_equal(p["total"], res.points["total"])
                and np.array_equal(p["temp"], res.points["temp"]) and np.array_equal(rng[k], sp.rng_state()))
        emin, tmin = sp.min_energy()
        n = len(res.snap_tooth)
        ref = np.concatenate([[e0], res.snap_energy, [emin], res.points["totalEnergy"]])
        got = np.concatenate([snap_e[i, : n + 1], [mins["energy"][i]], p["total_energy"]])
        err = float(np.max(np.abs(ref - got) / np.abs(ref)))
        out[k] = (bool(same and int(mins["tooth"][i]) == tmin), err)

    th = [threading.Thread(target=work, args=(k,)) for k in range(len(local_ids))]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    return {
        "chains": [int(seeds[i]) - 1000 for i in local_ids],
        "moves_per_chain": int(summ["moves"][local_ids[0]]),
        "decisions_equal": all(o[0] for o in out),
        "max_rel_err": max(o[1] for o in out),
        "what": "per-point accepted/total counters, temperatures, evaluated-move count, final state of the three MersenneTwister "
                "streams and minimum tooth equal to the oracle's; max_rel_err over snapshot energies, minimum energy and per-point "
                "energy sums",
        "oracle_seconds": round(time.perf_counter() - t0, 1),
    }


def run_ours(args):
    import torch
    import torch.distributed as dist

    from transrot_b200 import Engine, MINIMUM_DTYPE, gather_minima, global_minimum

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this backend has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dbase, cfg, scheds, system = load_problem(args)
    cpg = chains_per_gpu(args)
    first = rank * cpg
    gids = np.arange(first, first + cpg, dtype=np.int64)
    seeds = 1000 + gids
    sidx = sched_index(args, gids, len(scheds)) if len(scheds) > 1 else None
    pinned_seeds = torch.from_numpy(seeds).pin_memory()
    gather_structs = args.workload == "mixed64"
    row = system.n_sites * 3

    eng = Engine(local)
    # a non-default torch stream: the library launches on it, so torch.cuda.Event brackets see the kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    eng.set_system(system)
    eng.set_schedules(scheds)
    eng.set_kernel(args.kernel, args.chains_per_block, args.blocks_per_sm)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    minima_dev = torch.empty(cpg * MINIMUM_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    structs_dev = torch.empty(cpg * row, dtype=torch.float64, device="cuda") if gather_structs else None
    structs_all = torch.empty(world * cpg * row, dtype=torch.float64, device="cuda") if gather_structs and world > 1 else None
    structs_host = torch.empty(world * cpg * row, dtype=torch.float64).pin_memory() if gather_structs else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def init(record):
        eng.set_options(trace_moves=0, record_points=record, record_snapshots=record, threads_per_chain=args.threads_per_chain)
        eng.init_chains_seeded(pinned_seeds.numpy(), cfg.spaceLen, cfg.maxPropFailures, sched_idx=sidx, first_chain=first)

    # FP64 peak of this device, measured live (not in MEASURED_PEAKS.json): the roofline denominator
    peak_tflops, peak_clock = eng.measure_fp64_peak()

    # ---- value: the anneal kernel, chain state resident in HBM
    def device_step():
        init(False)
        flush.zero_()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count()
        e0.record(stream)
        eng.run(0)
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), eng.last_run_ms(), eng.launch_count() - l0

    for _ in range(args.warmup):
        device_step()
    sampler = ClockSampler(local)
    sampler.start()
    ms_dev, ms_kernel, launches = 0.0, 0.0, 0
    for _ in range(args.steps):
        a, b, c = device_step()
        ms_dev += a
        ms_kernel += b
        launches += c
    summ = eng.summary()
    assert int(summ["done"].sum()) == cpg, "not every chain finished"
    moves = int(summ["moves"].sum())
    pair_evals = int(summ["pair_evals"].sum())
    evaluated = int(summ["evaluated"].sum())

    # ---- e2e: the host-facing call sequence, host buffers in and out
    h2d = seeds.nbytes + (sidx.nbytes if sidx is not None else 0)
    keep = {}
    pinned = {}

    def pinned_like(name, shape, dtype):
        """Caller-owned page-locked result buffers (what a host that cares about transfer time passes to tr_get_*)."""
        if name not in pinned:
            nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
            t = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            pinned[name] = (t, t.numpy().view(dtype).reshape(shape))
        return pinned[name][1]

    def e2e_step():
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        init(True)                                   # H2D seeds (+ propagate on the device)
        eng.run(0)
        eng.pack_minima_device(minima_dev.data_ptr())
        allmin = gather_minima(minima_dev, world)    # NCCL all-gather of 24 B per chain
        s = eng.summary()
        n_pts, n_snap = eng.max_points()
        from transrot_b200 import POINT_DTYPE
        pts = eng.points(out=pinned_like("pts", (cpg, n_pts), POINT_DTYPE))
        # every write(n) structure: what the host turns into OutputN.xyz
        se, st, sx = eng.snapshots(out=(pinned_like("se", (cpg, n_snap), np.float64), pinned_like("st", (cpg, n_snap), np.int32),
                                        pinned_like("sx", (cpg, n_snap, system.n_sites, 3), np.float64)))
        best = global_minimum(allmin)
        nbytes = allmin.nbytes // world + s.nbytes + pts.nbytes + se.nbytes + st.nbytes + sx.nbytes
        if gather_structs:                           # config #4: every chain's minimum structure on every rank, then to the host
            eng.pack_min_structures_device(structs_dev.data_ptr())
            src = structs_dev
            if world > 1:
                dist.all_gather_into_tensor(structs_all, structs_dev)
                src = structs_all
            structs_host.copy_(src, non_blocking=True)
            nbytes += structs_host.numel() * 8
        if first <= best[1] < first + cpg:           # the owner fetches the winning structure
            x, e, t = eng.min_structure(best[1] - first)
            nbytes += x.nbytes
        torch.cuda.synchronize()
        keep.update(n=nbytes, best=best, summ=s, pts=pts, se=se, allmin=allmin)
        return (time.perf_counter() - t0) * 1e3

    e2e_step()   # one warm-up of the host path (allocations, pinned staging)
    ms_e2e = 0.0
    for _ in range(args.steps):
        ms_e2e += e2e_step()
    clocks = sampler.stop()

    # ---- parity of the chains just timed (rank 0, untimed)
    parity = None
    if rank == 0 and args.parity_chains > 0:
        ids = sorted({int(v) for v in np.linspace(0, cpg - 1, args.parity_chains)})
        rng_all = eng.rng_state()
        mins_local = keep["allmin"][:cpg] if world == 1 else keep["allmin"][first:first + cpg]
        parity = parity_check(dbase, scheds, sidx, seeds, ids, keep["pts"], keep["summ"], keep["se"], mins_local, rng_all[ids])

    # max over ranks
    t = torch.tensor([ms_dev, ms_kernel, ms_e2e], dtype=torch.float64, device="cuda")
    tot = torch.tensor([moves, pair_evals, launches, evaluated], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_dev, ms_kernel, ms_e2e = [float(v) for v in t.cpu()]
    moves_all, pairs_all, launches_all, evaluated_all = [float(v) for v in tot.cpu()]
    steps = args.steps
    value = moves_all * steps / (ms_dev * 1e-3)
    e2e_value = moves_all * steps / (ms_e2e * 1e-3)
    # roofline of the dominant kernel: executed pair evaluations x 19 flops over its device time, per GPU
    achieved_tflops = (pairs_all / world) * steps * FLOPS_PER_PAIR / (ms_kernel * 1e-3) / 1e12
    # the same time against the work the REFERENCE does for these moves (eStart and eEnd recomputed, Space.java:711-718)
    ref_pairs = evaluated_all / world * system.pair_evals_per_move()
    ref_work_tflops = ref_pairs * steps * FLOPS_PER_PAIR / (ms_kernel * 1e-3) / 1e12

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_descr(args, cfg, scheds, system, cpg),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(keep["n"]),
                    "ms_per_step": ms_e2e / steps},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": {
                "bound": "fp64", "bound_note": "FP64 pipe: the path is neither HBM- nor tensor-bound (chain state stays on chip, "
                                               "no contraction); achieved = EXECUTED site-pair evaluations x 19 flop / kernel time",
                "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
                "frac": achieved_tflops / peak_tflops, "traffic": None,
                "kernel": eng.kernel_name(),
                "flops_per_pair_eval": FLOPS_PER_PAIR,
                "pair_evals_per_launch": pairs_all / world,
                "pair_evals_per_evaluated_move": pairs_all / max(evaluated_all, 1.0),
                "reference_work_equivalent": {"tflops": ref_work_tflops, "frac": ref_work_tflops / peak_tflops,
                                              "note": "the reference's own pair evaluations for the same moves (eStart and eEnd) over the "
                                                      "same time; equals achieved unless the kernel caches eStart"},
                "peak_source": f"DFMA micro-benchmark run live on this device (tr_measure_fp64_peak), implied SM clock {peak_clock:.0f} MHz; "
                               "MEASURED_PEAKS.json has no FP64 figure",
                "ms_per_launch": ms_kernel / steps,
            },
            "best_minimum": {"energy": keep["best"][0], "chain": keep["best"][1], "tooth": keep["best"][2]},
            "jdk_probe": jdk_probe(),
        }
        if parity is not None:
            line["parity_check"] = parity
        if not args.no_cpu_baseline:
            r, cores, sample, _ = cpu_port_rate(dbase, cfg, args.cpu_seconds)
            line["cpu_baseline"] = {"value": r, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
        if parity is not None and not (parity["decisions_equal"] and parity["max_rel_err"] <= 1e-9):
            sys.stderr.write(f"bench.py: PARITY CHECK FAILED: {parity}\n")
            if world > 1:
                dist.barrier()
                dist.destroy_process_group()
            eng.close()
            sys.exit(1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
