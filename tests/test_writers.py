"""Output-file formats (reference Space.write / writeMovie / writeAcceptance / heat capacities): JVM number formatting
and the restart round trip OutputN.xyz -> Input.xyz (Space.java:187, 550-612)."""
import os

import numpy as np
import pytest

from transrot_b200 import (
    format_acceptance_line, format_heat_capacity_line, format_movie_frame, format_output_xyz, java_double_to_string,
    precision_round, read_input,
)


@pytest.mark.parametrize("value,text", [
    (1.0, "1.0"), (100.0, "100.0"), (1234567.0, "1234567.0"), (1.0e7, "1.0E7"), (123456789.0, "1.23456789E8"),
    (0.001, "0.001"), (9.999e-4, "9.999E-4"), (1.5e-10, "1.5E-10"), (-0.5, "-0.5"), (0.05, "0.05"), (0.0, "0.0"),
    (-460.8203874165545, "-460.8203874165545"), (1.1368683772161603e-12, "1.1368683772161603E-12"),
    (float("nan"), "NaN"), (float("inf"), "Infinity"),
])
def test_java_double_to_string(value, text):
    assert java_double_to_string(value) == text


def test_precision_round_is_half_up_on_the_decimal_string():
    assert precision_round(1.00000000005, 10) == 1.0000000001       # HALF_UP on "1.00000000005", not banker's / binary
    assert precision_round(-1.00000000005, 10) == -1.0000000001
    assert precision_round(2.5, 0) == 3.0 and precision_round(0.12345678904, 10) == 0.123456789


def test_output_xyz_format_and_restart_round_trip(dbase, golden_dir, tmp_path):
    system, xyz = read_input(os.path.join(golden_dir, "sample8.input.xyz"), dbase)
    text = format_output_xyz(system, xyz, -460.8203874165545)
    lines = text.split("\n")
    assert lines[0] == "          24" and lines[1] == "Energy: -460.8203874165545 Kcal/mole  4 NH4+ | 4 Cl-"
    assert len(lines) == 2 + 24 + 1 and lines[-1] == ""
    # symbol padded to two columns, integer part right-aligned in four, ten decimals
    sym, *cols = lines[2].split("  ")
    assert lines[2].startswith(" N ") and all(len(c.strip().split(".")[1]) == 10 for c in lines[2].split()[1:])
    f = tmp_path / "Output4.xyz"
    f.write_text(text)
    system2, xyz2 = read_input(str(f), dbase)                       # the reference strips "Energy: ... Kcal/mole"
    assert system2.n_mol == system.n_mol and np.max(np.abs(xyz2 - xyz)) <= 5.1e-11
    assert format_movie_frame(system, xyz, 1.5).split("\n")[1] == "Energy: 1.5 Kcal/mole"


def test_coordinate_columns():
    from transrot_b200.writers import _coord

    assert _coord(1.909373313536205) == "     1.9093733135"
    assert _coord(-12.5) == "   -12.5000000000"
    assert _coord(123.25) == "   123.2500000000"
    assert _coord(2.0e-4) == "     2.0E-4000000"        # the reference splits Double.toString on "." (Space.java:582): kept as is


def test_acceptance_and_heat_capacity_lines():
    assert format_acceptance_line(8888.888888888889, 12345, 20000) == "8888.89 0.61725\n"
    assert format_acceptance_line(1.1368683772161603e-12, 1, 3) == "0.00 0.33333\n"
    assert format_heat_capacity_line(1.1368683772161603e-12, -10.0, 5.0, 100) is None      # the `continue` of Main.java:228
    line = format_heat_capacity_line(300.0, -2000.0, 40010.0, 100)
    t, avg, cv = line.strip().split("\t")
    assert t == "300.0" and avg == "-20.0" and float(cv) == pytest.approx((400.1 - 400.0) / (0.0019872 * 300.0 ** 2))


def test_replay_outputs_from_synthetic_buffers(dbase, sample_cfg, golden_dir, tmp_path):
    """The output directory is rebuilt from plain arrays: no device needed to exercise the host side of the seam."""
    from transrot_b200 import POINT_DTYPE, replay_outputs

    system, xyz = read_input(os.path.join(golden_dir, "sample8.input.xyz"), dbase)
    cfg = sample_cfg.copy(movePerPt=10, ptsPerTooth=2, numTeeth=2, writeAcceptanceRatios=True, writeConfHeatCapacitiesEnabled=True)
    pts = np.zeros(4, dtype=POINT_DTYPE)
    pts["tooth"] = [0, 0, 1, 1]
    pts["point"] = [0, 1, 0, 1]
    pts["temp"] = [100.0, 0.0, 80.0, 0.0]
    pts["accepted"] = [4, 9, 5, 8]
    pts["total"] = [10, 20, 10, 20]
    pts["movie_written"] = [1, 0, 1, 0]                # the t = 0 points hit the `continue` of Main.java:228
    pts["total_energy"] = [-4000.0, -4100.0, -4200.0, -4300.0]
    pts["total_squared_energy"] = [1600400.0, 1681000.0, 1764400.0, 1849000.0]
    pts["movie_energy"] = [-400.5, np.nan, -420.25, np.nan]
    movie = np.stack([xyz + 0.01 * i for i in range(4)])
    out = tmp_path / "job"
    replay_outputs(str(out), system, cfg, pts, np.array([-390.0, -410.0, -430.0]), np.array([0, 1, 2]), np.stack([xyz, xyz + 0.1, xyz + 0.2]),
                   movie, min_tooth=2)
    assert sorted(os.listdir(out)) == ["Min_Energy_Structure_2.xyz", "Output0.xyz", "Output0_1_Movie.xyz", "Output1.xyz", "Output1_2_Movie.xyz",
                                      "Output2.xyz", "acceptance_ratios.txt", "configurational_heat_capacities.txt"]
    assert open(out / "acceptance_ratios.txt").read() == "100.00 0.40000\n0.00 0.45000\n80.00 0.50000\n0.00 0.40000\n"
    assert len(open(out / "configurational_heat_capacities.txt").read().splitlines()) == 2
    assert open(out / "Min_Energy_Structure_2.xyz").read() == open(out / "Output2.xyz").read()
    assert open(out / "Output0_1_Movie.xyz").read().split("\n")[1] == "Energy: -400.5 Kcal/mole"


def test_output_xyz_reproduces_the_reference_sample_byte_for_byte(dbase, golden_dir):
    """The reference's src/Input.xyz (fixture sample8.input.xyz, byte-identical copy) is a sample of the very format
    Space.write emits (Space.java:571-612): re-formatting its 24 atoms must give back its atom lines 3-26 exactly, and
    the atom-count line."""
    path = os.path.join(golden_dir, "sample8.input.xyz")
    system, xyz = read_input(path, dbase)
    want = open(path).read().split("\n")
    got = format_output_xyz(system, xyz, -460.8203874165545).split("\n")
    assert got[0] == want[0]
    assert got[2:26] == want[2:26]
    assert len([ln for ln in want[2:] if ln.strip()]) == 24
    # the comment line: Space.write puts "Energy: ... Kcal/mole  " in front of the particle list readInput wants
    assert got[1].endswith(want[1]) and got[1].startswith("Energy: ")


def test_config_to_text_round_trips_through_the_parser(sample_cfg, tmp_path):
    """Config.to_text writes a config.txt that Config.parseConfig (the reference's grammar, Config.java:36-48) reads back
    to the same values: what tools/jvm_pin.sh feeds the JVM."""
    from transrot_b200 import Config, workload_config

    for cfg in (sample_cfg, workload_config(sample_cfg, "mixed64", staticTemp=True, maxTemp=0.00001, finale=True, eqConfigs=7)):
        f = tmp_path / "config.txt"
        f.write_text(cfg.to_text())
        assert Config.parseConfig(str(f)) == cfg
