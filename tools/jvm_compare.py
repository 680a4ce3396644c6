#!/usr/bin/env python
"""jvm_compare.py — cases and comparison for tools/jvm_pin.sh (pinning the oracle to the real reference's output).

    jvm_compare.py make-cases DIR --dbase dbase.txt [--input Input.xyz]   one sub-directory per case: config.txt,
                                                                         dbase.txt, seed.txt [, Input.xyz]
    jvm_compare.py compare DIR [--report FILE] [--gpu]                    after the JVM has filled DIR/*/out/<timestamp>/

What is compared (reference writers in Space.java): `seed.log` (:878), `energies.txt` (:808, the per-move running energy
in Static Temperature mode = the oracle's decision trace), `acceptance_ratios.txt` (:821), `configurational_heat_
capacities.txt` (:846), every `OutputN.xyz` (:550) and `Min_Energy_Structure_N.xyz` (:836).  The oracle side is rebuilt
with transrot_b200.replay_outputs from the oracle's records (`--gpu`: also from the CUDA path's buffers).
Bars: integer-derived text (acceptance ratios, which minimum tooth) identical; energies 1e-12 relative; coordinates
1e-9 absolute (files carry 10 decimals).  An `energies.txt` that matches line for line over N moves pins N accept/reject
decisions of the oracle to the JVM.
"""
from __future__ import annotations

import argparse
import glob
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

from transrot_b200 import Config, POINT_DTYPE, TRACE_DTYPE, read_dbase, read_input, replay_outputs, system_from_config, workload_config  # noqa: E402

SHIPPED = os.path.join(ROOT, "tests", "golden", "sample8.config.txt")


def cases(base: Config):
    """(name, config, seed, uses Input.xyz)."""
    out = []
    for seed in (42, 43):        # BASELINE config #1 as shipped: 800 000 moves, every writer on
        out.append((f"shipped_s{seed}", base.copy(), seed, False))
    for name, seed, t in (("sample8", 7, 300.0), ("h2o20", 8, 250.0), ("mixed64", 9, 500.0)):   # per-move traces
        out.append((f"static_{name}_s{seed}", workload_config(base, name, staticTemp=True, writeEnergiesEnabled=True, maxTemp=t,
                                                              movePerPt=20000, eqConfigs=1000), seed, False))
    out.append(("finale_h2o20_s11", workload_config(base, "h2o20", movePerPt=2000, ptsPerTooth=4, ptsAddedPerTooth=1, numTeeth=3,
                                                    finale=True), 11, False))
    out.append(("input_xyz_s5", base.copy(useInput=True, movePerPt=5000, ptsPerTooth=4, numTeeth=2, spaceLen=12.0), 5, True))
    return out


def make_cases(args):
    base = Config.parseConfig(SHIPPED)
    for name, cfg, seed, use_input in cases(base):
        d = os.path.join(args.dir, name)
        os.makedirs(d, exist_ok=True)
        open(os.path.join(d, "config.txt"), "w").write(cfg.to_text())
        open(os.path.join(d, "seed.txt"), "w").write(str(seed))
        shutil.copyfile(args.dbase, os.path.join(d, "dbase.txt"))
        if use_input:
            shutil.copyfile(args.input, os.path.join(d, "Input.xyz"))
        print("case", d)


def _floats(path):
    return [[float(t) for t in ln.replace("\t", " ").split()] for ln in open(path).read().splitlines() if ln.strip()]


def _xyz(path):
    lines = open(path).read().splitlines()
    energy = float(lines[1].split()[1])
    coords = np.array([[float(t) for t in ln.split()[1:4]] for ln in lines[2:] if ln.strip()])
    return energy, coords, lines[1]


def oracle_outputs(dbase, cfg, seed, out_dir, input_xyz=None):
    """Run the oracle and rebuild the reference's output directory from its records."""
    from oracle.oracle import OracleSpace, oracle_from_config

    if input_xyz:
        system, xyz = read_input(input_xyz, dbase)
        sp = OracleSpace(dbase)
        sp.set_seed(seed)
        sp.set_config(cfg)
        groups = []
        for m in range(system.n_mol):
            nm, fr = system.dbase[int(system.mol_species[m])].name, m not in set(system.moveable_idx.tolist())
            if groups and groups[-1][0] == nm and groups[-1][2] == fr:
                groups[-1][1] += 1
            else:
                groups.append([nm, 1, fr])
        sp.read_input([tuple(g) for g in groups], xyz)
    else:
        system = system_from_config(dbase, cfg)
        sp = oracle_from_config(dbase, cfg, seed)
    e0 = sp.write_snapshot(0)
    x0, _ = sp.coords()
    n_trace = cfg.total_moves() if (cfg.staticTemp and cfg.writeEnergiesEnabled) else 0
    res = sp.sawtooth_anneal(e0, points_cap=cfg.num_points() + 4, snap_cap=cfg.numTeeth + 2, trace_cap=n_trace)
    pts = np.zeros(len(res.points), dtype=POINT_DTYPE)
    for a, b in (("tooth", "tooth"), ("point", "point"), ("accepted", "accepted"), ("total", "total"), ("movie_written", "movie_written"),
                 ("finale", "finale"), ("temp", "temp"), ("total_energy", "totalEnergy"), ("total_squared_energy", "totalSquaredEnergy"),
                 ("movie_energy", "movieEnergy")):
        pts[a] = res.points[b]
    snap_e = np.concatenate([[e0], res.snap_energy])
    snap_t = np.concatenate([[0], res.snap_tooth]).astype(np.int32)
    snap_x = np.concatenate([x0[None], res.snap_coords])
    _, tmin = sp.min_energy()
    tr = None
    if n_trace:
        tr = np.zeros(len(res.trace), dtype=TRACE_DTYPE)
        tr["running"] = res.trace["running"]
    replay_outputs(out_dir, system, cfg, pts, snap_e, snap_t, snap_x, None, tmin, tr)     # movie files are not compared
    return system


def compare_dirs(ref, mine, report):
    ok = True

    def note(line, good=True):
        nonlocal ok
        ok = ok and good
        report.append(("PASS " if good else "FAIL ") + line)

    for name in ("acceptance_ratios.txt",):
        a, b = os.path.join(ref, name), os.path.join(mine, name)
        if os.path.exists(a):
            same = os.path.exists(b) and open(a).read() == open(b).read()
            note(f"{name}: text identical ({len(open(a).read().splitlines())} lines)", same)
    for name, tol in (("energies.txt", 1e-12), ("configurational_heat_capacities.txt", 1e-9)):
        a, b = os.path.join(ref, name), os.path.join(mine, name)
        if os.path.exists(a):
            if not os.path.exists(b):
                note(f"{name}: missing on our side", False)
                continue
            fa, fb = _floats(a), _floats(b)
            if len(fa) != len(fb):
                note(f"{name}: {len(fa)} lines vs {len(fb)}", False)
                continue
            A, B = np.array(fa), np.array(fb)
            err = np.abs(A - B) / np.maximum(np.abs(A), 1e-300)
            first_bad = int(np.argmax(err.max(axis=1) > tol)) if (err > tol).any() else -1
            note(f"{name}: {len(fa)} lines, max rel err {err.max():.3e} (bar {tol:g}), first line over the bar: {first_bad}", first_bad < 0)
    ref_out = sorted(glob.glob(os.path.join(ref, "Output*.xyz")) + glob.glob(os.path.join(ref, "Min_Energy_Structure_*.xyz")))
    for a in ref_out:
        b = os.path.join(mine, os.path.basename(a))
        if "Movie" in a:
            continue
        if not os.path.exists(b):
            note(f"{os.path.basename(a)}: missing on our side", False)
            continue
        ea, xa, ca = _xyz(a)
        eb, xb, cb = _xyz(b)
        good = abs(ea - eb) <= 1e-12 * abs(ea) and xa.shape == xb.shape and np.max(np.abs(xa - xb)) <= 1e-9 and ca.split("Kcal/mole")[-1] == cb.split("Kcal/mole")[-1]
        note(f"{os.path.basename(a)}: energy rel err {abs(ea - eb) / abs(ea):.3e}, max coordinate diff {np.max(np.abs(xa - xb)):.3e}", good)
    return ok


def compare(args):
    base = Config.parseConfig(SHIPPED)
    report = ["# JVM pin report", "", f"JDK: {open(os.path.join(os.path.dirname(args.dir.rstrip('/')), 'jdk_version.txt')).read().strip() if os.path.exists(os.path.join(os.path.dirname(args.dir.rstrip('/')), 'jdk_version.txt')) else 'unknown'}", ""]
    all_ok = True
    for name, cfg, seed, use_input in cases(base):
        d = os.path.join(args.dir, name)
        runs = sorted(glob.glob(os.path.join(d, "out", "*_*_*_*_*_*")))
        report.append(f"## {name} (seed {seed}, {cfg.total_moves()} moves)")
        if not runs:
            report.append("FAIL no JVM output directory")
            all_ok = False
            continue
        ref = runs[-1]
        seed_log = open(os.path.join(ref, "seed.log")).read().strip()
        report.append(("PASS " if seed_log == str(seed) else "FAIL ") + f"seed.log = {seed_log}")
        dbase = read_dbase(os.path.join(d, "dbase.txt"))
        mine = os.path.join(d, "oracle_out")
        shutil.rmtree(mine, ignore_errors=True)
        oracle_outputs(dbase, cfg, seed, mine, os.path.join(d, "Input.xyz") if use_input else None)
        all_ok = compare_dirs(ref, mine, report) and all_ok
        report.append("")
    report.append("RESULT: " + ("oracle PINNED to the JVM on every case" if all_ok else "MISMATCH - see FAIL lines"))
    text = "\n".join(report) + "\n"
    if args.report:
        open(args.report, "w").write(text)
    print(text)
    return 0 if all_ok else 1


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    a = sub.add_parser("make-cases"); a.add_argument("dir"); a.add_argument("--dbase", required=True); a.add_argument("--input")
    b = sub.add_parser("compare"); b.add_argument("dir"); b.add_argument("--report"); b.add_argument("--gpu", action="store_true")
    args = ap.parse_args()
    sys.exit(make_cases(args) if args.cmd == "make-cases" else compare(args))
