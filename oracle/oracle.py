"""ctypes binding of the C oracle (oracle/transrot_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs.  Never imported by transrot_b200/.
PARITY UNPINNED — see transrot_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "transrot_oracle.c")
    hdr = os.path.join(_HERE, "transrot_oracle.h")
    if (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class MT(C.Structure):
    _fields_ = [("mt", C.c_uint32 * 624), ("mti", C.c_int32)]


class OrConfig(C.Structure):
    _fields_ = [
        ("maxTemp", C.c_double),
        ("tempDecreasePerTooth", C.c_double),
        ("maxTrans", C.c_double),
        ("magwalkFactorTrans", C.c_double),
        ("magwalkProbTrans", C.c_double),
        ("maxRot", C.c_double),
        ("magwalkProbRot", C.c_double),
        ("spaceLen", C.c_double),
        ("movePerPt", C.c_int32),
        ("ptsPerTooth", C.c_int32),
        ("ptsAddedPerTooth", C.c_int32),
        ("numTeeth", C.c_int32),
        ("maxPropFailures", C.c_int32),
        ("eqConfigs", C.c_int32),
        ("finale", C.c_int32),
        ("staticTemp", C.c_int32),
        ("writeConfHeatCapacitiesEnabled", C.c_int32),
        ("pad_", C.c_int32),
    ]


TRACE_DTYPE = np.dtype(
    [
        ("mol", "<i4"),
        ("kind", "i1"),
        ("magwalk", "i1"),
        ("bounds", "i1"),
        ("accepted", "i1"),
        ("axis", "<i4"),
        ("drew_rand", "<i4"),
        ("e_start", "<f8"),
        ("e_end", "<f8"),
        ("dE", "<f8"),
        ("rand", "<f8"),
        ("temp", "<f8"),
        ("running", "<f8"),
    ],
    align=True,
)
POINT_DTYPE = np.dtype(
    [
        ("tooth", "<i4"),
        ("point", "<i4"),
        ("accepted", "<i4"),
        ("total", "<i4"),
        ("movie_written", "<i4"),
        ("finale", "<i4"),
        ("temp", "<f8"),
        ("totalEnergy", "<f8"),
        ("totalSquaredEnergy", "<f8"),
        ("movieEnergy", "<f8"),
    ],
    align=True,
)
assert TRACE_DTYPE.itemsize == 64 and POINT_DTYPE.itemsize == 56


class AnnealOut(C.Structure):
    _fields_ = [
        ("trace", C.c_void_p), ("trace_cap", C.c_int64), ("trace_len", C.c_int64),
        ("points", C.c_void_p), ("points_cap", C.c_int64), ("points_len", C.c_int64),
        ("snap_tooth", C.c_void_p), ("snap_energy", C.c_void_p), ("snap_coords", C.c_void_p),
        ("snap_cap", C.c_int64), ("snap_len", C.c_int64),
        ("trace_abs", C.c_void_p),
        ("moves", C.c_int64), ("evaluated", C.c_int64), ("pair_evals", C.c_int64),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int32)
    vp = C.c_void_p
    sig = {
        "or_mt_set_seed_int": (None, [C.POINTER(MT), C.c_uint32]),
        "or_mt_set_seed_array": (None, [C.POINTER(MT), C.POINTER(C.c_uint32), C.c_int]),
        "or_mt_set_seed_long": (None, [C.POINTER(MT), C.c_int64]),
        "or_mt_next": (C.c_uint32, [C.POINTER(MT), C.c_int]),
        "or_mt_next_double": (C.c_double, [C.POINTER(MT)]),
        "or_mt_next_int": (C.c_int32, [C.POINTER(MT), C.c_int32]),
        "or_mt_next_long": (C.c_int64, [C.POINTER(MT)]),
        "or_new": (vp, []),
        "or_free": (None, [vp]),
        "or_dbase_add": (C.c_int, [vp, C.c_double, C.c_int, dp, ip]),
        "or_add": (C.c_int, [vp, C.c_int, C.c_int]),
        "or_setup_pair_vals": (None, [vp]),
        "or_set_pair": (None, [vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]),
        "or_num_types": (C.c_int, [vp]),
        "or_get_pair_table": (None, [vp, dp]),
        "or_set_config": (None, [vp, C.POINTER(OrConfig)]),
        "or_get_config": (None, [vp, C.POINTER(OrConfig)]),
        "or_set_seed": (None, [vp, C.c_int64]),
        "or_get_stream_seeds": (None, [vp, C.POINTER(C.c_int64)]),
        "or_get_rng": (None, [vp, C.POINTER(MT)]),
        "or_set_rng": (None, [vp, C.POINTER(MT)]),
        "or_propagate": (C.c_int, [vp]),
        "or_propagate_until_placed": (C.c_int, [vp]),
        "or_read_input": (C.c_int, [vp, C.c_int, ip, ip, ip, dp]),
        "or_num_molecules": (C.c_int, [vp]),
        "or_num_moveable": (C.c_int, [vp]),
        "or_num_sites": (C.c_int, [vp]),
        "or_get_layout": (None, [vp, ip, ip, ip, ip]),
        "or_get_coords": (None, [vp, dp, dp]),
        "or_set_coords": (None, [vp, dp, dp]),
        "or_calc_energy": (C.c_double, [vp]),
        "or_calc_energy_abs": (C.c_double, [vp, dp]),
        "or_mol_energy": (C.c_double, [vp, C.c_int, C.c_int]),
        "or_move": (C.c_int, [vp, C.c_int, C.c_double, C.c_double, dp, vp]),
        "or_rotate": (C.c_int, [vp, C.c_int, C.c_double, C.c_double, dp, vp]),
        "or_sawtooth_anneal": (C.c_double, [vp, C.c_double, C.POINTER(AnnealOut)]),
        "or_write_snapshot": (C.c_double, [vp, C.c_int]),
        "or_get_min": (None, [vp, dp, ip]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class JavaMT:
    """commons-math3 MersenneTwister restated in C; KAT target."""

    def __init__(self, seed: Optional[int] = None):
        self.s = MT()
        if seed is not None:
            self.set_seed(seed)

    def set_seed(self, seed: int):
        lib().or_mt_set_seed_long(C.byref(self.s), C.c_int64(seed))

    def set_seed_array(self, key: Sequence[int]):
        k = (C.c_uint32 * len(key))(*key)
        lib().or_mt_set_seed_array(C.byref(self.s), k, len(key))

    def next(self, bits: int = 32) -> int:
        return int(lib().or_mt_next(C.byref(self.s), bits))

    def next_double(self) -> float:
        return float(lib().or_mt_next_double(C.byref(self.s)))

    def next_int(self, n: int) -> int:
        return int(lib().or_mt_next_int(C.byref(self.s), n))

    def next_long(self) -> int:
        return int(lib().or_mt_next_long(C.byref(self.s)))

    def state(self) -> np.ndarray:
        out = np.empty(625, dtype=np.uint32)
        out[:624] = np.ctypeslib.as_array(self.s.mt)
        out[624] = self.s.mti
        return out


class AnnealResult:
    def __init__(self):
        self.final_energy = 0.0
        self.trace = np.zeros(0, dtype=TRACE_DTYPE)
        self.points = np.zeros(0, dtype=POINT_DTYPE)
        self.snap_tooth = np.zeros(0, dtype=np.int32)
        self.snap_energy = np.zeros(0, dtype=np.float64)
        self.snap_coords = np.zeros((0, 0, 3), dtype=np.float64)
        self.trace_abs = np.zeros(0, dtype=np.float64)
        self.moves = 0
        self.evaluated = 0
        self.pair_evals = 0


class OracleSpace:
    """One reference process: a `Space`, its static `Config` and the three static RNGs."""

    def __init__(self, dbase, pair_table: Optional[np.ndarray] = None):
        """dbase: sequence of objects with .radius and .atoms[i].(x y z a b c d q mass ghost)."""
        self.L = lib()
        self.h = self.L.or_new()
        for p in dbase:
            params = np.asarray(
                [[a.x, a.y, a.z, a.a, a.b, a.c, a.d, a.q, a.mass] for a in p.atoms], dtype=np.float64
            ).reshape(-1)
            ghost = np.asarray([1 if a.ghost else 0 for a in p.atoms], dtype=np.int32)
            self.L.or_dbase_add(self.h, float(p.radius), len(p.atoms), _dp(params), _ip(ghost))
        self.names = [p.name for p in dbase]
        if pair_table is None:
            self.L.or_setup_pair_vals(self.h)
        else:
            T = self.L.or_num_types(self.h)
            pt = np.asarray(pair_table, dtype=np.float64).reshape(T, T, 4)
            for i in range(T):
                for j in range(T):
                    self.L.or_set_pair(self.h, i, j, *[float(v) for v in pt[i, j]])

    def __del__(self):
        try:
            if self.h:
                self.L.or_free(self.h)
                self.h = None
        except Exception:
            pass

    # ---- configuration
    def set_config(self, cfg):
        c = OrConfig(
            maxTemp=cfg.maxTemp, tempDecreasePerTooth=cfg.tempDecreasePerTooth, maxTrans=cfg.maxTrans,
            magwalkFactorTrans=cfg.magwalkFactorTrans, magwalkProbTrans=cfg.magwalkProbTrans, maxRot=cfg.maxRot,
            magwalkProbRot=cfg.magwalkProbRot, spaceLen=cfg.spaceLen, movePerPt=cfg.movePerPt,
            ptsPerTooth=cfg.ptsPerTooth, ptsAddedPerTooth=cfg.ptsAddedPerTooth, numTeeth=cfg.numTeeth,
            maxPropFailures=cfg.maxPropFailures, eqConfigs=cfg.eqConfigs, finale=int(cfg.finale),
            staticTemp=int(cfg.staticTemp), writeConfHeatCapacitiesEnabled=int(cfg.writeConfHeatCapacitiesEnabled),
        )
        self.L.or_set_config(self.h, C.byref(c))

    @property
    def space_len(self) -> float:
        c = OrConfig()
        self.L.or_get_config(self.h, C.byref(c))
        return float(c.spaceLen)

    def add(self, name: str, count: int) -> bool:
        if name not in self.names:
            return False
        return bool(self.L.or_add(self.h, self.names.index(name), count))

    def set_seed(self, seed: int):
        self.L.or_set_seed(self.h, C.c_int64(seed))

    def stream_seeds(self):
        out = (C.c_int64 * 3)()
        self.L.or_get_stream_seeds(self.h, out)
        return [int(v) for v in out]

    def rng_state(self) -> np.ndarray:
        """[3, 625] uint32: streams M, S, R; 624 state words + mti."""
        arr = (MT * 3)()
        self.L.or_get_rng(self.h, arr)
        out = np.empty((3, 625), dtype=np.uint32)
        for i in range(3):
            out[i, :624] = np.ctypeslib.as_array(arr[i].mt)
            out[i, 624] = arr[i].mti
        return out

    def set_rng_state(self, st: np.ndarray):
        arr = (MT * 3)()
        for i in range(3):
            for k in range(624):
                arr[i].mt[k] = int(st[i, k])
            arr[i].mti = int(st[i, 624])
        self.L.or_set_rng(self.h, arr)

    # ---- initial state
    def propagate_until_placed(self) -> int:
        return int(self.L.or_propagate_until_placed(self.h))

    def read_input(self, groups, coords: np.ndarray):
        """groups: list of (name, count, frozen)."""
        sp = np.asarray([self.names.index(g[0]) for g in groups], dtype=np.int32)
        ct = np.asarray([g[1] for g in groups], dtype=np.int32)
        fr = np.asarray([1 if g[2] else 0 for g in groups], dtype=np.int32)
        xyz = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1)
        if not self.L.or_read_input(self.h, len(groups), _ip(sp), _ip(ct), _ip(fr), _dp(xyz)):
            raise RuntimeError("or_read_input failed")

    # ---- geometry
    @property
    def n_mol(self) -> int:
        return int(self.L.or_num_molecules(self.h))

    @property
    def n_sites(self) -> int:
        return int(self.L.or_num_sites(self.h))

    @property
    def n_moveable(self) -> int:
        return int(self.L.or_num_moveable(self.h))

    def layout(self):
        N, S, M = self.n_mol, self.n_sites, self.n_moveable
        sp = np.zeros(N, dtype=np.int32)
        off = np.zeros(N + 1, dtype=np.int32)
        st = np.zeros(S, dtype=np.int32)
        mv = np.zeros(max(M, 1), dtype=np.int32)
        self.L.or_get_layout(self.h, _ip(sp), _ip(off), _ip(st), _ip(mv))
        return sp, off, st, mv[:M]

    def coords(self):
        sites = np.zeros((self.n_sites, 3), dtype=np.float64)
        com = np.zeros((self.n_mol, 3), dtype=np.float64)
        self.L.or_get_coords(self.h, _dp(sites), _dp(com))
        return sites, com

    def set_coords(self, sites: np.ndarray, com: Optional[np.ndarray] = None):
        s = np.ascontiguousarray(sites, dtype=np.float64)
        if com is None:
            self.L.or_set_coords(self.h, _dp(s), None)
        else:
            c = np.ascontiguousarray(com, dtype=np.float64)
            self.L.or_set_coords(self.h, _dp(s), _dp(c))

    def pair_table(self) -> np.ndarray:
        T = int(self.L.or_num_types(self.h))
        out = np.zeros((T, T, 4), dtype=np.float64)
        self.L.or_get_pair_table(self.h, _dp(out))
        return out

    # ---- energies and moves
    def calc_energy(self) -> float:
        return float(self.L.or_calc_energy(self.h))

    def calc_energy_abs(self):
        """(calcEnergy(), sum of |pair term|): the second value is the conditioning scale of the first."""
        a = C.c_double()
        e = float(self.L.or_calc_energy_abs(self.h, C.byref(a)))
        return e, float(a.value)

    def mol_energy(self, m: int, n: int) -> float:
        return float(self.L.or_mol_energy(self.h, m, n))

    def _step(self, fn, m: int, max_step: float, temp: float):
        rec = np.zeros(1, dtype=TRACE_DTYPE)
        dE = C.c_double(0.0)
        acc = fn(self.h, m, float(max_step), float(temp), C.byref(dE), rec.ctypes.data)
        return int(acc), float(dE.value), rec[0]

    def move(self, m: int, maxD: float, temp: float):
        return self._step(self.L.or_move, m, maxD, temp)

    def rotate(self, m: int, maxRot: float, temp: float):
        return self._step(self.L.or_rotate, m, maxRot, temp)

    def write_snapshot(self, n: int) -> float:
        return float(self.L.or_write_snapshot(self.h, n))

    def min_energy(self):
        e = C.c_double()
        t = C.c_int32()
        self.L.or_get_min(self.h, C.byref(e), C.byref(t))
        return float(e.value), int(t.value)

    def sawtooth_anneal(self, starting_energy: float, trace_cap: int = 0, points_cap: int = 0, snap_cap: int = 0) -> AnnealResult:
        S = self.n_sites
        res = AnnealResult()
        trace = np.zeros(max(trace_cap, 1), dtype=TRACE_DTYPE)
        trace_abs = np.zeros(max(trace_cap, 1), dtype=np.float64)
        points = np.zeros(max(points_cap, 1), dtype=POINT_DTYPE)
        st = np.zeros(max(snap_cap, 1), dtype=np.int32)
        se = np.zeros(max(snap_cap, 1), dtype=np.float64)
        sc = np.zeros((max(snap_cap, 1), S, 3), dtype=np.float64)
        out = AnnealOut(
            trace=trace.ctypes.data if trace_cap else None, trace_cap=trace_cap,
            points=points.ctypes.data if points_cap else None, points_cap=points_cap,
            snap_tooth=st.ctypes.data, snap_energy=se.ctypes.data, snap_coords=sc.ctypes.data, snap_cap=snap_cap,
            trace_abs=trace_abs.ctypes.data,
        )
        res.final_energy = float(self.L.or_sawtooth_anneal(self.h, float(starting_energy), C.byref(out)))
        res.trace = trace[: out.trace_len].copy()
        res.trace_abs = trace_abs[: out.trace_len].copy()
        res.points = points[: out.points_len].copy()
        res.snap_tooth = st[: out.snap_len].copy()
        res.snap_energy = se[: out.snap_len].copy()
        res.snap_coords = sc[: out.snap_len].copy()
        res.moves, res.evaluated, res.pair_evals = int(out.moves), int(out.evaluated), int(out.pair_evals)
        return res


def oracle_from_config(dbase, cfg, seed: int, pair_table=None) -> OracleSpace:
    """Main.main init sequence up to (not including) write(0) (Main.java:34-61) for a propagated start."""
    sp = OracleSpace(dbase, pair_table)
    sp.set_seed(seed)
    sp.set_config(cfg)
    for name, count in cfg.particles:
        if not sp.add(name, count):
            raise RuntimeError(f"unknown particle {name}")
    sp.propagate_until_placed()
    return sp
