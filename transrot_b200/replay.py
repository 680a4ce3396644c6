"""Rebuilds the reference's output directory for one chain from the buffers the device returns - the host-side
half of the seam (the Java twin is sketched in INTEGRATION.md as `replayOutputs`).

Mirrors the file writes of Main.sawtoothAnneal's y-loop (Main.java:191-258) and of Main.main (Main.java:64):
`Output0.xyz` / `OutputN.xyz` (Space.write), `OutputX_Y_Movie.xyz` (Space.writeMovie, one frame per point),
`acceptance_ratios.txt`, `configurational_heat_capacities.txt`, `energies.txt` (Static Temperature only) and
`Min_Energy_Structure_N.xyz` (Space.writeMinEnergy, only when Number of Teeth > 1)."""
from __future__ import annotations

import os
import shutil
from typing import Optional

import numpy as np

from .host import Config, System
from .writers import (
    format_acceptance_line, format_energies, format_heat_capacity_line, format_movie_frame, format_output_xyz,
)


def replay_outputs(out_dir: str, system: System, cfg: Config, points: np.ndarray, snap_energy: np.ndarray,
                   snap_tooth: np.ndarray, snap_sites: np.ndarray, movie: Optional[np.ndarray] = None,
                   min_tooth: Optional[int] = None, trace: Optional[np.ndarray] = None) -> None:
    """points: POINT_DTYPE records of ONE chain (only the emitted ones); snap_*: its write(n) snapshots;
    movie: [points, S, 3] frames (Engine.movie()[chain]) or None; trace: TRACE_DTYPE records for energies.txt."""
    os.makedirs(out_dir, exist_ok=True)

    def append(name: str, text: str):
        with open(os.path.join(out_dir, name), "a") as fh:
            fh.write(text)

    for k in range(len(snap_tooth)):
        if snap_tooth[k] < 0:
            continue
        with open(os.path.join(out_dir, f"Output{int(snap_tooth[k])}.xyz"), "w") as fh:      # Space.write(n)
            fh.write(format_output_xyz(system, snap_sites[k], float(snap_energy[k])))
    for i, p in enumerate(points):
        if cfg.writeAcceptanceRatios:
            append("acceptance_ratios.txt", format_acceptance_line(float(p["temp"]), int(p["accepted"]), int(p["total"])))
        if cfg.writeConfHeatCapacitiesEnabled:
            line = format_heat_capacity_line(float(p["temp"]), float(p["total_energy"]), float(p["total_squared_energy"]), cfg.movePerPt)
            if line is not None:
                append("configurational_heat_capacities.txt", line)
        if p["movie_written"] and movie is not None:
            x = int(p["tooth"])
            append(f"Output{x}_{x + 1}_Movie.xyz", format_movie_frame(system, movie[i], float(p["movie_energy"])))
    if cfg.staticTemp and cfg.writeEnergiesEnabled and trace is not None:
        append("energies.txt", format_energies(trace["running"]))
    if cfg.numTeeth > 1 and not cfg.staticTemp and min_tooth is not None:
        shutil.copyfile(os.path.join(out_dir, f"Output{int(min_tooth)}.xyz"), os.path.join(out_dir, f"Min_Energy_Structure_{int(min_tooth)}.xyz"))
