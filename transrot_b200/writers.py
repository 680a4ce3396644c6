"""Output-file formats of the reference, rebuilt from the buffers the device returns.

The Java host keeps its own `write*` methods; these mirrors exist so that a host without a JVM (this repo's
Python harness) produces byte-compatible files - `OutputN.xyz` (Space.java:550-612), movie frames (Space.java:740-
790), `acceptance_ratios.txt` (Space.java:821-834), `configurational_heat_capacities.txt` (Main.java:226-233,
Space.java:846-859), `energies.txt` (Space.java:808-819) - and so that an `OutputN.xyz` can be fed back through
`read_input` exactly as the reference allows (restart via Input.xyz, Space.java:187).

Number formatting follows the JVM: `Double.toString` (shortest repr; plain decimal for 1e-3 <= |d| < 1e7, else
"d.dddE-n"), `Precision.round(x, 10, ROUND_HALF_UP)` = `new BigDecimal(Double.toString(x)).setScale(10, HALF_UP)`
(commons-math3 3.6.1), `new BigDecimal(double).setScale(n, HALF_UP)` on the exact binary value.
"""
from __future__ import annotations

from decimal import ROUND_HALF_UP, Decimal
from typing import List, Sequence

import numpy as np

from .host import System

BOLTZMANN_CONSTANT = 0.0019872   # Space.java:32


def java_double_to_string(d: float) -> str:
    """java.lang.Double.toString for finite and non-finite doubles (JDK 19+ shortest-digits behaviour)."""
    d = float(d)
    if d != d:
        return "NaN"
    if d in (float("inf"), float("-inf")):
        return "Infinity" if d > 0 else "-Infinity"
    if d == 0.0:
        return "-0.0" if str(d).startswith("-") else "0.0"
    sign = "-" if d < 0 else ""
    a = abs(d)
    r = repr(a)
    if "e" in r or "E" in r:
        mant, e = r.lower().split("e")
        exp10 = int(e)
    else:
        mant, exp10 = r, 0
    ip, _, fp = mant.partition(".")
    all_digits = (ip + fp).lstrip("0")
    # decimal exponent of the first significant digit
    if ip.strip("0"):
        first = len(ip.lstrip("0")) - 1 + exp10
    else:
        first = -(len(fp) - len(fp.lstrip("0")) + 1) + exp10
    all_digits = all_digits.rstrip("0") or "0"
    if 1e-3 <= a < 1e7:
        if first >= 0:
            int_part = all_digits[: first + 1].ljust(first + 1, "0")
            frac = all_digits[first + 1:] or "0"
        else:
            int_part = "0"
            frac = "0" * (-first - 1) + all_digits
        return f"{sign}{int_part}.{frac}"
    frac = all_digits[1:] or "0"
    return f"{sign}{all_digits[0]}.{frac}E{first}"


def precision_round(x: float, scale: int = 10) -> float:
    """commons-math3 Precision.round(x, scale, BigDecimal.ROUND_HALF_UP)."""
    x = float(x)
    if x != x or x in (float("inf"), float("-inf")):
        return x
    q = Decimal(java_double_to_string(x).replace("E", "e")).quantize(Decimal(1).scaleb(-scale), rounding=ROUND_HALF_UP)
    r = float(q)
    return -0.0 if r == 0.0 and x < 0 else r


def _coord(v: float) -> str:
    """One coordinate as Space.write / writeMovie print it: rounded to 10 decimals, integer part right-aligned in four
    columns, fraction right-padded with zeros to ten (Space.java:582-591)."""
    parts = java_double_to_string(precision_round(v, 10)).split(".")
    return "  " + " " * max(0, 4 - len(parts[0])) + parts[0] + "." + parts[1] + "0" * max(0, 10 - len(parts[1]))


def _atom_lines(system: System, sites: np.ndarray) -> List[str]:
    out = []
    for sym, (x, y, z) in zip(system.symbols(), np.asarray(sites, dtype=np.float64).reshape(-1, 3)):
        if "*" in sym:                     # ghosts are not written (Space.java:574,748)
            continue
        out.append("\n " + sym + (" " if len(sym) == 1 else "") + _coord(x) + _coord(y) + _coord(z))
    return out


def num_atoms(system: System) -> int:
    """Space.numAtoms (Space.java:614-624): ghosts excluded."""
    return sum(1 for s in system.symbols() if "*" not in s)


def molecule_string(system: System) -> str:
    """The `3 NH4+ | 4 Cl-` summary of Space.write (Space.java:556-568): runs of equal names in `space` order."""
    names = [system.dbase[int(sp)].name for sp in system.mol_species]
    runs, last, count = [], "", 0
    for n in names:
        if n != last:
            if last:
                runs.append(f"{count} {last}")
            last, count = n, 0
        count += 1
    runs.append(f"{count} {last}")
    return " | ".join(runs)


def format_output_xyz(system: System, sites: np.ndarray, energy: float) -> str:
    """Contents of OutputN.xyz (Space.java:550-612)."""
    head = "          " + str(num_atoms(system)) + "\nEnergy: " + java_double_to_string(energy) + " Kcal/mole  " + molecule_string(system)
    return head + "".join(_atom_lines(system, sites)) + "\n"


def format_movie_frame(system: System, sites: np.ndarray, energy: float) -> str:
    """One frame appended to OutputX_Y_Movie.xyz (Space.java:740-790)."""
    head = "          " + str(num_atoms(system)) + "\nEnergy: " + java_double_to_string(energy) + " Kcal/mole"
    return head + "".join(_atom_lines(system, sites)) + "\n"


def _bigdecimal_scale(x: float, scale: int) -> str:
    """new BigDecimal(double).setScale(scale, RoundingMode.HALF_UP).toString() - on the exact binary value."""
    q = Decimal(float(x)).quantize(Decimal(1).scaleb(-scale), rounding=ROUND_HALF_UP)
    return f"{q:.{scale}f}"


def format_acceptance_line(temperature: float, accepted: int, total: int) -> str:
    """Space.writeAcceptance (Space.java:821-834): cumulative ratio within the tooth."""
    return _bigdecimal_scale(temperature, 2) + " " + _bigdecimal_scale(accepted / total, 5) + "\n"


def format_heat_capacity_line(temperature: float, total_energy: float, total_squared_energy: float, moves_per_point: int):
    """Main.java:226-233 + Space.writeConfigurationalHeatCapacity; None when round(t, 2) == 0 (the `continue`)."""
    rounded = float(Decimal(float(temperature)).quantize(Decimal("0.01"), rounding=ROUND_HALF_UP))
    if rounded == 0:
        return None
    avg = total_energy / moves_per_point
    avg2 = total_squared_energy / moves_per_point
    cv = (avg2 - avg ** 2) / (BOLTZMANN_CONSTANT * temperature ** 2)
    return "\t".join(java_double_to_string(v) for v in (rounded, avg, cv)) + "\n"


def format_energies(running: Sequence[float]) -> str:
    """energies.txt (Space.writeEnergy, Space.java:808-819): startingEnergy after every move."""
    return "".join(java_double_to_string(v) + "\n" for v in running)
