"""Host-side mirror of the reference's data model for the annealing hot path.

The reference's Java host keeps parsing (`Space.readDB`, `Config.parseConfig`,
`Space.readInput`, `Space.readParams`/`setupPairVals`) and flattens the result into
the arrays the C ABI takes (include/transrot_b200.h).  No JVM exists in this image,
so this module restates exactly that flattening in Python for the tests, the bench
and `smoke()`.  It computes nothing on the hot path: energies, moves, RNG and the
schedule all live in the CUDA library.

Citations are into /root/reference/src/main/.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field, replace
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

K_VALUE = 332.0637133  # Atom.java:10


class TransRotError(RuntimeError):
    """Stands for the RuntimeException("Error: ...") the reference throws (Main.java:85-87)."""


# --------------------------------------------------------------------------- dbase


@dataclass
class AtomDef:
    """One dbase atom slot (Atom.java:53-70).  `type_id` plays the role of the UUID (Atom.java:35,54)."""

    symbol: str
    x: float
    y: float
    z: float
    a: float
    b: float
    c: float
    d: float
    q: float
    mass: float
    type_id: int = -1

    @property
    def ghost(self) -> bool:
        return "*" in self.symbol


@dataclass
class ParticleDef:
    """One dbase particle (Molecule.java:25-31)."""

    name: str
    radius: float
    atoms: List[AtomDef]


_DB_COUNT = re.compile(r"^\d+$")
_DB_NAME = re.compile(r"^(?P<pname>.+?) {2,}(?P<radius>\d+(.\d+)?)$")
_DB_ATOM = re.compile(r"^(?P<aname>[A-Za-z0-9*_-]+)(?P<vars>( {2,}-?\d+(.\d+)?){8} {2,}\d+(.\d+)?)+$")


def _significant_lines(text: str) -> List[str]:
    return [ln.strip() for ln in text.splitlines() if ln.strip() and not ln.strip().startswith("//")]


def read_dbase(path: str) -> List[ParticleDef]:
    """Space.readDB (Space.java:119-176): same grammar, same error texts."""
    try:
        with open(path, "r") as fh:
            lines = _significant_lines(fh.read())
    except OSError:
        raise TransRotError(f"Error: File {path} not found.")
    out: List[ParticleDef] = []
    names: List[str] = []
    i = 0
    curr = 1
    next_type = 0
    while i < len(lines):
        if not _DB_COUNT.search(lines[i]):
            raise TransRotError(f"Error in line {curr} of database file: Expected atom count.")
        count = int(lines[i])
        i += 1
        curr += 1
        if i >= len(lines):
            raise TransRotError(f"Error in line {curr} of database file: Unexpected end of file.")
        m = _DB_NAME.search(lines[i])
        if not m:
            raise TransRotError(f"Error in line {curr} of database file: Unexpected format for particle definition.")
        pname = m.group("pname")
        if pname in names:
            raise TransRotError(f"Error in line {curr} of database file: Particle identifiers must be unique.")
        names.append(pname)
        radius = float(m.group("radius"))
        i += 1
        curr += 1
        atoms: List[AtomDef] = []
        nghost = 0
        for _ in range(count):
            if i >= len(lines):
                raise TransRotError(f"Error in line {curr} of database file: Unexpected end of file.")
            am = _DB_ATOM.search(lines[i])
            if not am:
                raise TransRotError(f"Error in line {curr} of database file: Unexpected format for atom definition.")
            vals = [float(v) for v in re.split(r" {2,}", am.group("vars").strip())]
            sym = am.group("aname")
            if "*" in sym:
                nghost += 1
            if nghost == count:
                raise TransRotError(
                    f"Error in line {curr} of database file: A particle may not be comprised of only ghost atoms."
                )
            atoms.append(AtomDef(sym, *vals[:9], type_id=next_type))
            next_type += 1
            i += 1
            curr += 1
        out.append(ParticleDef(pname, radius, atoms))
    return out


# -------------------------------------------------------------------------- config

_VAR_TYPES = {  # Config.java:50-95
    "Max Temperature": "DOUBLE",
    "Moves per Point": "INT",
    "Points per Tooth": "INT",
    "Points Added per Tooth": "INT",
    "Number of Teeth": "INT",
    "Temperature Decrease per Tooth": "DOUBLE",
    "Max Translation Distance": "DOUBLE",
    "Magwalk Translation Multiplication Factor": "DOUBLE",
    "Magwalk Translation Probability": "DOUBLE",
    "Max Rotation": "DOUBLE",
    "Magwalk Rotation Probability": "DOUBLE",
    "Length of Cubic Space": "DOUBLE",
    "Max Failures During Propagation": "INT",
    "Use Input.xyz": "BOOLEAN",
    "Choose All Interaction Parameters": "BOOLEAN",
    "0K Finale": "BOOLEAN",
    "Static Temperature": "BOOLEAN",
    "Write Energies When Static Temperature": "BOOLEAN",
    "Write Configurational Heat Capacities": "BOOLEAN",
    "Number of Equilibration Configurations": "INT",
    "Write Acceptance Ratios": "BOOLEAN",
}
_VAR_RE = {"INT": re.compile(r"^\d+$"), "DOUBLE": re.compile(r"^\d+(\.\d+)?$"), "BOOLEAN": re.compile(r"^(true|false)$")}
_CFG_SETTING = re.compile(r"^(?P<setting>.+): +(?P<value>.+)$")
_CFG_PARTICLE = re.compile(r"^(?P<particle>(\S+\s?)+) {2,}(?P<count>\d+)$")
_CFG_UNIT = re.compile(r"\([a-zA-Z0-9/_-]+\)")


@dataclass
class Config:
    """Config.java:12-33, same field names."""

    maxTemp: float = 10000.0
    movePerPt: int = 20000
    ptsPerTooth: int = 10
    ptsAddedPerTooth: int = 0
    numTeeth: int = 4
    tempDecreasePerTooth: float = 0.8
    maxTrans: float = 0.25
    magwalkFactorTrans: float = 10.0
    magwalkProbTrans: float = 0.1
    maxRot: float = 0.523599
    magwalkProbRot: float = 0.1
    spaceLen: float = 5.0
    maxPropFailures: int = 20
    useInput: bool = False
    chooseParams: bool = False
    finale: bool = False
    staticTemp: bool = False
    writeEnergiesEnabled: bool = False
    writeConfHeatCapacitiesEnabled: bool = False
    eqConfigs: int = 0
    writeAcceptanceRatios: bool = False
    particles: List[Tuple[str, int]] = field(default_factory=list)

    def copy(self, **kw) -> "Config":
        kw.setdefault("particles", list(self.particles))
        return replace(self, **kw)

    def num_points(self) -> int:
        """Iterations of the y-loop at Main.java:191 over the whole run (finale included)."""
        if self.staticTemp:
            return 1
        pts, n = self.ptsPerTooth, 0
        for _ in range(self.numTeeth):
            n += pts
            pts += self.ptsAddedPerTooth
        if self.finale:
            n += pts
        return n

    def total_moves(self) -> int:
        return self.num_points() * self.movePerPt

    def to_text(self) -> str:
        """A config.txt the reference's own parser accepts (Config.java:36-48,102-169): what tools/jvm_pin.sh feeds
        the JVM.  Doubles are written in plain decimal notation (the reference's DOUBLE regex has no exponent)."""
        def d(x: float) -> str:
            t = repr(float(x))
            if "e" in t or "E" in t:
                t = f"{float(x):.17f}".rstrip("0")
                t += "0" if t.endswith(".") else ""
            return t
        b = lambda v: "true" if v else "false"
        rows = [
            ("Max Temperature (K)", d(self.maxTemp)), ("Moves per Point", str(self.movePerPt)),
            ("Points per Tooth", str(self.ptsPerTooth)), ("Points Added per Tooth", str(self.ptsAddedPerTooth)),
            ("Number of Teeth", str(self.numTeeth)), ("Temperature Decrease per Tooth", d(self.tempDecreasePerTooth)),
            ("Max Translation Distance (Angstroms)", d(self.maxTrans)),
            ("Magwalk Translation Multiplication Factor", d(self.magwalkFactorTrans)),
            ("Magwalk Translation Probability", d(self.magwalkProbTrans)), ("Max Rotation (Radians)", d(self.maxRot)),
            ("Magwalk Rotation Probability", d(self.magwalkProbRot)), ("Length of Cubic Space", d(self.spaceLen)),
            ("Max Failures During Propagation", str(self.maxPropFailures)), ("Use Input.xyz (true/false)", b(self.useInput)),
            ("Choose All Interaction Parameters (true/false)", b(self.chooseParams)), ("0K Finale (true/false)", b(self.finale)),
            ("Static Temperature (true/false)", b(self.staticTemp)),
            ("Write Energies When Static Temperature (true/false)", b(self.writeEnergiesEnabled)),
            ("Write Configurational Heat Capacities (true/false)", b(self.writeConfHeatCapacitiesEnabled)),
            ("Number of Equilibration Configurations", str(self.eqConfigs)),
            ("Write Acceptance Ratios (true/false)", b(self.writeAcceptanceRatios)),
        ]
        out = ["// Sawtooth Annealing Configuration Options"] + [f"{k + ':':<54}{v}" for k, v in rows]
        out += ["", "// List of Molecules to be Used"] + [f"{n}  {c}" for n, c in self.particles]
        return "\n".join(out) + "\n"

    @staticmethod
    def parseConfig(filename: str) -> "Config":
        """Config.parseConfig (Config.java:102-169).  Particle lines are returned in file order; the caller
        feeds them to `SystemBuilder.add` like the reference feeds `Space.add` (Config.java:149)."""
        try:
            with open(filename, "r") as fh:
                raw = fh.read().splitlines()
        except OSError:
            raise TransRotError(f"Error: File {filename} not found.")
        values: Dict[str, str] = {}
        particles: List[Tuple[str, int]] = []
        finished = False
        for curr, ln in enumerate(raw, start=1):
            line = ln.strip()
            if line.startswith("//") or not line:
                continue
            # matched lazily, in the reference's order: the particle pattern backtracks badly on setting lines
            sm = _CFG_SETTING.search(line)
            pm = None if sm else _CFG_PARTICLE.search(line)
            if sm:
                if finished:
                    raise TransRotError(
                        f"Error on line {curr} in config file: Particle counts must be defined after all "
                        "configuration settings in configuration file."
                    )
                name = _CFG_UNIT.split(sm.group("setting"))[0].strip()
                value = sm.group("value")
                if name not in _VAR_TYPES:
                    raise TransRotError(f"Error on line {curr} in config file: Configuration setting '{name}' not expected.")
                if not _VAR_RE[_VAR_TYPES[name]].match(value):
                    raise TransRotError(f"Error on line {curr} in config file: Unexpected type for '{name}' configuration")
                values[name] = value
            elif pm:
                finished = True
                pname = pm.group("particle")
                if any(p == pname for p, _ in particles):
                    raise TransRotError(
                        f"Error on line {curr} in config file: Particle counts can only be specified once per particle."
                    )
                particles.append((pname, int(pm.group("count"))))
            else:
                raise TransRotError(f"Error on line {curr} in config file: Unexpected format.")
        if len(values) != len(_VAR_TYPES):
            missing = next(n for n in _VAR_TYPES if n not in values)
            raise TransRotError(f"Error in config file: Missing setting '{missing}'.")
        b = lambda k: values[k] == "true"
        return Config(
            maxTemp=float(values["Max Temperature"]),
            movePerPt=int(values["Moves per Point"]),
            ptsPerTooth=int(values["Points per Tooth"]),
            ptsAddedPerTooth=int(values["Points Added per Tooth"]),
            numTeeth=int(values["Number of Teeth"]),
            tempDecreasePerTooth=float(values["Temperature Decrease per Tooth"]),
            maxTrans=float(values["Max Translation Distance"]),
            magwalkFactorTrans=float(values["Magwalk Translation Multiplication Factor"]),
            magwalkProbTrans=float(values["Magwalk Translation Probability"]),
            maxRot=float(values["Max Rotation"]),
            magwalkProbRot=float(values["Magwalk Rotation Probability"]),
            spaceLen=float(values["Length of Cubic Space"]),
            maxPropFailures=int(values["Max Failures During Propagation"]),
            useInput=b("Use Input.xyz"),
            chooseParams=b("Choose All Interaction Parameters"),
            finale=b("0K Finale"),
            staticTemp=b("Static Temperature"),
            writeEnergiesEnabled=b("Write Energies When Static Temperature"),
            writeConfHeatCapacitiesEnabled=b("Write Configurational Heat Capacities"),
            eqConfigs=int(values["Number of Equilibration Configurations"]),
            writeAcceptanceRatios=b("Write Acceptance Ratios"),
            particles=particles,
        )


# -------------------------------------------------------------------- pair table


def setup_pair_vals(dbase: Sequence[ParticleDef]) -> np.ndarray:
    """Space.setupPairVals (Space.java:346-366): [type1][type2][A,B,C,D]."""
    atoms = [a for p in dbase for a in p.atoms]
    n = len(atoms)
    tab = np.zeros((n, n, 4), dtype=np.float64)
    for a1 in atoms:
        for a2 in atoms:
            v = (math.sqrt(a1.a * a2.a), (a1.b + a2.b) / 2, math.sqrt(a1.c * a2.c), math.sqrt(a1.d * a2.d))
            tab[a1.type_id, a2.type_id] = v
            tab[a2.type_id, a1.type_id] = v
    return tab


_P_ALIAS = re.compile(r"^(?P<particle>.+) *= *(?P<aliases>([A-Za-z0-9-_]+, *)*[A-Za-z0-9-_]+)$")
_P_PARAM = re.compile(r"^(?P<firstID>[A-Za-z0-9-_]+) *- *(?P<secondID>[A-Za-z0-9-_]+)(?P<values>( {2,}\d+(.\d+)?){4})$")


def read_params(path: str, dbase: Sequence[ParticleDef]) -> np.ndarray:
    """Space.readParams (Space.java:234-343): alias lines + explicit unlike-alias rows; like-alias pairs
    come from the combination rules (Space.java:326-340).  Pairs never mentioned stay absent in the
    reference's HashMap (a later lookup would NPE); here they are NaN so misuse is loud."""
    try:
        with open(path, "r") as fh:
            raw = fh.read().splitlines()
    except OSError:
        raise TransRotError(f"Error: File {path} not found.")
    aliased: List[str] = []
    aliases: Dict[str, List[AtomDef]] = {}
    params: List[Tuple[str, str, List[float]]] = []
    for curr, ln in enumerate(raw, start=1):
        line = ln.strip()
        if not line:
            continue
        am = _P_ALIAS.search(line)
        pm = None if am else _P_PARAM.search(line)
        if am:
            pname = am.group("particle")
            if pname in aliased:
                raise TransRotError(
                    f"Error on line {curr} in interaction parameters file: Cannot alias the same particle type twice."
                )
            names = [s.strip() for s in am.group("aliases").split(",")]
            part = next((p for p in dbase if p.name == pname), None)
            if part is None:
                raise TransRotError(f"Error on line {curr} in interaction parameters file: Aliased particle does not exist.")
            if len(part.atoms) != len(names):
                raise TransRotError(
                    f"Error on line {curr} in interaction parameters file: Number of atom aliases does not match "
                    "number of atoms defined in the input database."
                )
            for al, at in zip(names, part.atoms):
                aliases.setdefault(al, []).append(at)
            aliased.append(pname)
        elif pm:
            vals = [float(v) for v in re.split(r" {2,}", pm.group("values").strip())]
            params.append((pm.group("firstID"), pm.group("secondID"), vals))
        else:
            raise TransRotError(f"Error on line {curr} in interaction parameters file: File incorrectly formatted.")
    if len(aliased) != len(dbase):
        raise TransRotError(
            "Error in interaction parameters file: Must provide aliases for all defined particles in the database."
        )
    if len(params) < (len(aliases) * (len(aliases) - 1)) // 2:
        raise TransRotError("Error in interaction parameters file: Not all interactions are defined.")
    n = sum(len(p.atoms) for p in dbase)
    tab = np.full((n, n, 4), np.nan, dtype=np.float64)
    seen = set()
    for first, second, vals in params:
        for al in (first, second):
            if al not in aliases:
                raise TransRotError(
                    f"Error in interaction parameters file: Atom alias '{al}' not defined in particle alias lists."
                )
        for a1 in aliases[first]:
            for a2 in aliases[second]:
                if (a1.type_id, a2.type_id) in seen or (a2.type_id, a1.type_id) in seen:
                    raise TransRotError(
                        f"Error in interaction parameters file: Interaction between '{first}' and '{second}' "
                        "cannot be defined twice."
                    )
                seen.add((a1.type_id, a2.type_id))
                seen.add((a2.type_id, a1.type_id))
                tab[a1.type_id, a2.type_id] = vals
                tab[a2.type_id, a1.type_id] = vals
    for al, ats in aliases.items():
        for a1 in ats:
            for a2 in ats:
                v = (math.sqrt(a1.a * a2.a), (a1.b + a2.b) / 2, math.sqrt(a1.c * a2.c), math.sqrt(a1.d * a2.d))
                tab[a1.type_id, a2.type_id] = v
                tab[a2.type_id, a1.type_id] = v
    return tab


# ------------------------------------------------------------------------- system


@dataclass
class System:
    """Flat, chain-independent description of `space`: what `tr_set_system` takes."""

    dbase: List[ParticleDef]
    mol_species: np.ndarray   # int32 [N]   index into dbase
    mol_site_off: np.ndarray  # int32 [N+1]
    mol_radius: np.ndarray    # f64   [N]
    site_type: np.ndarray     # int32 [S]   dbase atom slot
    site_q: np.ndarray        # f64   [S]
    site_def: np.ndarray      # f64   [S,3] template coordinates (Atom.defX/Y/Z)
    site_mass: np.ndarray     # f64   [S]
    site_ghost: np.ndarray    # int32 [S]
    moveable_idx: np.ndarray  # int32 [n_moveable] indices into space, in moveableMolecules order
    pair_table: np.ndarray    # f64   [T,T,4]

    @property
    def n_mol(self) -> int:
        return int(self.mol_species.shape[0])

    @property
    def n_sites(self) -> int:
        return int(self.site_type.shape[0])

    @property
    def n_types(self) -> int:
        return int(self.pair_table.shape[0])

    def pair_evals_per_move(self) -> float:
        """P_ref averaged over a uniform pick from the movable particles."""
        ns = np.diff(self.mol_site_off).astype(np.int64)
        S = int(ns.sum())
        per = 2 * ns * (S - ns)
        return float(per[self.moveable_idx].mean())

    def symbols(self) -> List[str]:
        out = []
        for sp in self.mol_species:
            out.extend(a.symbol for a in self.dbase[int(sp)].atoms)
        return out


class SystemBuilder:
    """`Space.add` (Space.java:46-69): radius-sorted insertion (strictly-greater goes first, ties keep order)."""

    def __init__(self, dbase: Sequence[ParticleDef], pair_table: Optional[np.ndarray] = None):
        self.dbase = list(dbase)
        self.pair_table = setup_pair_vals(self.dbase) if pair_table is None else np.asarray(pair_table, dtype=np.float64)
        self._list: List[int] = []
        self._frozen: List[bool] = []

    def _species(self, name: str) -> int:
        for i, p in enumerate(self.dbase):
            if p.name == name:
                return i
        return -1

    def add(self, name: str, count: int = 1) -> bool:
        sp = self._species(name)
        if sp < 0:
            return False
        for _ in range(count):
            r = self.dbase[sp].radius
            for n, other in enumerate(self._list):
                if r > self.dbase[other].radius:
                    self._list.insert(n, sp)
                    self._frozen.insert(n, False)
                    break
            else:
                self._list.append(sp)
                self._frozen.append(False)
        return True

    def add_in_order(self, name: str, count: int, frozen: bool = False) -> bool:
        """`readInput` order (Space.java:200-220): no sorting, optional frozen flag."""
        sp = self._species(name)
        if sp < 0:
            return False
        self._list.extend([sp] * count)
        self._frozen.extend([frozen] * count)
        return True

    def build(self) -> System:
        species = np.asarray(self._list, dtype=np.int32)
        offs = [0]
        st, sq, sd, sm, sg, rad = [], [], [], [], [], []
        for sp in self._list:
            p = self.dbase[sp]
            rad.append(p.radius)
            for a in p.atoms:
                st.append(a.type_id)
                sq.append(a.q)
                sd.append((a.x, a.y, a.z))
                sm.append(a.mass)
                sg.append(1 if a.ghost else 0)
            offs.append(offs[-1] + len(p.atoms))
        mov = [i for i, f in enumerate(self._frozen) if not f]
        return System(
            dbase=self.dbase,
            mol_species=species,
            mol_site_off=np.asarray(offs, dtype=np.int32),
            mol_radius=np.asarray(rad, dtype=np.float64),
            site_type=np.asarray(st, dtype=np.int32),
            site_q=np.asarray(sq, dtype=np.float64),
            site_def=np.asarray(sd, dtype=np.float64).reshape(-1, 3),
            site_mass=np.asarray(sm, dtype=np.float64),
            site_ghost=np.asarray(sg, dtype=np.int32),
            moveable_idx=np.asarray(mov, dtype=np.int32),
            pair_table=self.pair_table,
        )


def system_from_config(dbase: Sequence[ParticleDef], cfg: Config, pair_table: Optional[np.ndarray] = None) -> System:
    b = SystemBuilder(dbase, pair_table)
    for name, count in cfg.particles:
        if not b.add(name, count):
            raise TransRotError(f"Error in config file: Unable to find particle '{name}' in dbase.")
    return b.build()


_XYZ_ENERGY = re.compile(r"Energy: -?\d*.?\d* Kcal/mole")


def read_input(path: str, dbase: Sequence[ParticleDef], pair_table: Optional[np.ndarray] = None) -> Tuple[System, np.ndarray]:
    """Space.readInput (Space.java:178-233).  Returns the system (Input.xyz order, frozen groups excluded from
    `moveable_idx`) and the site coordinates [S,3]."""
    try:
        with open(path, "r") as fh:
            lines = fh.read().split("\n")
    except OSError:
        raise TransRotError(f"Error: File {path} not found.")
    curr = 1
    try:
        num_atoms = int(lines[0].strip())
        info = _XYZ_ENERGY.sub("", lines[1], count=1).split("|")
        curr = 2
        b = SystemBuilder(dbase, pair_table)
        coords: List[Tuple[float, float, float]] = []
        li = 2
        for item in info:
            parts = item.strip().split(" ")
            if len(parts) < 2:
                raise TransRotError("Error: Incorrect format on line 2 of input file")
            num = parts[0]
            frozen = num.upper().endswith("F")
            count = int(num[:-1] if frozen else num)
            name = " ".join(parts[1:])
            if not any(p.name == name for p in dbase):
                raise TransRotError("Error: Unable to find one or more molecules from input file in dbase.txt.")
            part = next(p for p in dbase if p.name == name)
            for _ in range(count):
                for a in part.atoms:
                    vals = re.split(r" {2,}", lines[li].strip())
                    li += 1
                    curr += 1
                    if vals[0] != a.symbol:
                        raise TransRotError(f"Error: Unexpected atom on line {curr} of input file.")
                    coords.append((float(vals[1]), float(vals[2]), float(vals[3])))
            b.add_in_order(name, count, frozen)
        if li < len(lines) and lines[li] != "" :
            raise TransRotError("Error: Molecule counts must match # of lines in input file.")
        if num_atoms != len(coords):
            raise TransRotError("Error: Atom count at top of input file does not match number of molecules read.")
    except ValueError:
        raise TransRotError(f"Error on line {curr} in input file: File incorrectly formatted.")
    return b.build(), np.asarray(coords, dtype=np.float64).reshape(-1, 3)


def center_of_mass(system: System, sites: np.ndarray) -> np.ndarray:
    """Molecule.setCenterOfMass (Molecule.java:137-156), same accumulation order.  sites: [..., S, 3]."""
    sites = np.asarray(sites, dtype=np.float64)
    lead = sites.shape[:-2]
    out = np.zeros(lead + (system.n_mol, 3), dtype=np.float64)
    for m in range(system.n_mol):
        tot = 0.0
        acc = np.zeros(lead + (3,), dtype=np.float64)
        for s in range(int(system.mol_site_off[m]), int(system.mol_site_off[m + 1])):
            if system.site_ghost[s]:
                continue
            tot += float(system.site_mass[s])
            acc = acc + sites[..., s, :] * float(system.site_mass[s])
        out[..., m, :] = acc / tot
    return out
