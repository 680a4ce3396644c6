"""ctypes binding of include/transrot_b200.h — the same calls the Java host would bind with Panama
(INTEGRATION.md).  Thin by design: every number on the hot path is produced by the CUDA library.
There is no CPU fallback; a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from .host import Config, System, TransRotError

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TR_LIB", os.path.join(_HERE, "libtransrot_b200.so"))

TR_MAX_MOL_SITES = 16


class TrSystem(C.Structure):
    _fields_ = [
        ("n_mol", C.c_int32), ("n_sites", C.c_int32), ("n_types", C.c_int32), ("n_moveable", C.c_int32),
        ("mol_site_off", C.c_void_p), ("mol_radius", C.c_void_p), ("site_type", C.c_void_p), ("site_q", C.c_void_p),
        ("site_def", C.c_void_p), ("site_mass", C.c_void_p), ("site_ghost", C.c_void_p), ("moveable_idx", C.c_void_p),
        ("pair_abcd", C.c_void_p),
    ]


SCHEDULE_DTYPE = np.dtype(
    [
        ("max_temp", "<f8"), ("temp_decrease_per_tooth", "<f8"), ("max_trans", "<f8"), ("magwalk_factor_trans", "<f8"),
        ("magwalk_prob_trans", "<f8"), ("max_rot", "<f8"), ("magwalk_prob_rot", "<f8"),
        ("moves_per_point", "<i4"), ("points_per_tooth", "<i4"), ("points_added_per_tooth", "<i4"), ("num_teeth", "<i4"),
        ("eq_configs", "<i4"), ("finale", "<i4"), ("static_temp", "<i4"), ("write_conf_heat_capacities", "<i4"),
    ],
    align=True,
)
POINT_DTYPE = np.dtype(
    [
        ("tooth", "<i4"), ("point", "<i4"), ("accepted", "<i4"), ("total", "<i4"), ("movie_written", "<i4"), ("finale", "<i4"),
        ("temp", "<f8"), ("total_energy", "<f8"), ("total_squared_energy", "<f8"), ("movie_energy", "<f8"),
    ],
    align=True,
)
TRACE_DTYPE = np.dtype(
    [
        ("mol", "<i4"), ("kind", "i1"), ("magwalk", "i1"), ("bounds", "i1"), ("accepted", "i1"), ("axis", "<i4"),
        ("drew_rand", "<i4"), ("e_start", "<f8"), ("e_end", "<f8"), ("dE", "<f8"), ("rand", "<f8"), ("temp", "<f8"),
        ("running", "<f8"),
    ],
    align=True,
)
SUMMARY_DTYPE = np.dtype(
    [
        ("start_energy", "<f8"), ("running_energy", "<f8"), ("min_energy", "<f8"), ("space_len", "<f8"),
        ("moves", "<i8"), ("evaluated", "<i8"), ("accepted", "<i8"), ("pair_evals", "<i8"),
        ("min_tooth", "<i4"), ("snapshots", "<i4"), ("points", "<i4"), ("done", "<i4"), ("prop_retries", "<i4"), ("pad_", "<i4"),
    ],
    align=True,
)
MINIMUM_DTYPE = np.dtype([("energy", "<f8"), ("chain", "<i8"), ("tooth", "<i4"), ("pad_", "<i4")], align=True)
assert SCHEDULE_DTYPE.itemsize == 88 and POINT_DTYPE.itemsize == 56 and TRACE_DTYPE.itemsize == 64
assert SUMMARY_DTYPE.itemsize == 88 and MINIMUM_DTYPE.itemsize == 24

_SIGNATURES = {
    "tr_create": [C.c_int, C.POINTER(C.c_void_p)],
    "tr_destroy": [C.c_void_p],
    "tr_last_error": [C.c_void_p],
    "tr_set_stream": [C.c_void_p, C.c_void_p],
    "tr_synchronize": [C.c_void_p],
    "tr_set_system": [C.c_void_p, C.POINTER(TrSystem)],
    "tr_set_schedules": [C.c_void_p, C.c_void_p, C.c_int32],
    "tr_set_options": [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32],
    "tr_init_chains_seeded": [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_int32],
    "tr_init_chains_explicit": [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double],
    "tr_run": [C.c_void_p, C.c_int64],
    "tr_last_run_ms": [C.c_void_p, C.POINTER(C.c_float)],
    "tr_launch_count": [C.c_void_p, C.POINTER(C.c_int64)],
    "tr_get_summary": [C.c_void_p, C.c_void_p],
    "tr_get_sites": [C.c_void_p, C.c_void_p, C.c_void_p],
    "tr_get_points": [C.c_void_p, C.c_void_p, C.c_int64],
    "tr_get_snapshots": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64],
    "tr_set_record_movie": [C.c_void_p, C.c_int32],
    "tr_get_movie": [C.c_void_p, C.c_void_p, C.c_int64],
    "tr_get_trace": [C.c_void_p, C.c_void_p],
    "tr_get_rng_state": [C.c_void_p, C.c_void_p],
    "tr_pack_minima": [C.c_void_p, C.c_void_p, C.c_void_p],
    "tr_get_min_structure": [C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)],
    "tr_max_points": [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)],
    "tr_total_energy": [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p],
    "tr_delta_energy": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p],
    "tr_rng_words": [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p],
    "tr_measure_fp64_peak": [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)],
    "tr_set_kernel": [C.c_void_p, C.c_int32, C.c_int32, C.c_int32],
    "tr_abi_sizes": [C.c_void_p, C.c_int32],
    "tr_kernel_name": [C.c_void_p],
    "tr_pack_min_structures": [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p],
    # multi-GPU, one host thread (csrc/multi.cu)
    "tr_create_multi": [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)],
    "tr_destroy_multi": [C.c_void_p],
    "tr_multi_last_error": [C.c_void_p],
    "tr_multi_size": [C.c_void_p],
    "tr_multi_ctx": [C.c_void_p, C.c_int32],
    "tr_multi_chain_range": [C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)],
    "tr_multi_set_system": [C.c_void_p, C.POINTER(TrSystem)],
    "tr_multi_set_schedules": [C.c_void_p, C.c_void_p, C.c_int32],
    "tr_multi_set_options": [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32],
    "tr_multi_init_chains_seeded": [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_int32],
    "tr_multi_run": [C.c_void_p, C.c_int64],
    "tr_multi_synchronize": [C.c_void_p],
    "tr_multi_last_run_ms": [C.c_void_p, C.POINTER(C.c_float)],
    "tr_gather_minima": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "tr_gather_min_structures": [C.c_void_p, C.c_void_p],
}
_RESTYPES = {"tr_destroy": None, "tr_destroy_multi": None, "tr_multi_chain_range": None, "tr_last_error": C.c_char_p,
             "tr_multi_last_error": C.c_char_p, "tr_kernel_name": C.c_char_p, "tr_multi_ctx": C.c_void_p, "tr_multi_size": C.c_int32}
TR_KERNEL_AUTO, TR_KERNEL_TEAM, TR_KERNEL_CREW, TR_KERNEL_CREW2, TR_KERNEL_TEAM_PC = 0, 1, 2, 3, 4
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load_library() -> C.CDLL:
    """Load libtransrot_b200.so.  Raises if it was not built: nothing here can run without it."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TransRotError(
            f"Error: {LIB_PATH} is missing; build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    for name, args in _SIGNATURES.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, C.c_int)
    _check_abi(L)
    _lib = L
    return L


def _check_abi(L: C.CDLL):
    """The struct layouts of this module against the library's own sizeof / offsetof (tr_abi_sizes)."""
    n = L.tr_abi_sizes(None, 0)
    got = (C.c_int32 * n)()
    L.tr_abi_sizes(got, n)
    off = lambda dt, f: dt.fields[f][1]
    want = [
        C.sizeof(TrSystem), TrSystem.mol_site_off.offset, TrSystem.pair_abcd.offset,
        SCHEDULE_DTYPE.itemsize, off(SCHEDULE_DTYPE, "moves_per_point"), off(SCHEDULE_DTYPE, "write_conf_heat_capacities"),
        POINT_DTYPE.itemsize, off(POINT_DTYPE, "temp"), off(POINT_DTYPE, "movie_energy"),
        TRACE_DTYPE.itemsize, off(TRACE_DTYPE, "e_start"), off(TRACE_DTYPE, "running"),
        SUMMARY_DTYPE.itemsize, off(SUMMARY_DTYPE, "moves"), off(SUMMARY_DTYPE, "min_tooth"),
        MINIMUM_DTYPE.itemsize, off(MINIMUM_DTYPE, "chain"), off(MINIMUM_DTYPE, "tooth"),
    ]
    if list(got) != want:
        raise TransRotError(f"Error: struct layouts of {LIB_PATH} differ from this binding: library {list(got)}, binding {want}")


def schedule_from_config(cfg: Config) -> np.ndarray:
    """The Config fields Main.sawtoothAnneal reads (Config.java:12-33) as one tr_schedule."""
    s = np.zeros(1, dtype=SCHEDULE_DTYPE)
    s["max_temp"] = cfg.maxTemp
    s["temp_decrease_per_tooth"] = cfg.tempDecreasePerTooth
    s["max_trans"] = cfg.maxTrans
    s["magwalk_factor_trans"] = cfg.magwalkFactorTrans
    s["magwalk_prob_trans"] = cfg.magwalkProbTrans
    s["max_rot"] = cfg.maxRot
    s["magwalk_prob_rot"] = cfg.magwalkProbRot
    s["moves_per_point"] = cfg.movePerPt
    s["points_per_tooth"] = cfg.ptsPerTooth
    s["points_added_per_tooth"] = cfg.ptsAddedPerTooth
    s["num_teeth"] = cfg.numTeeth
    s["eq_configs"] = cfg.eqConfigs
    s["finale"] = int(cfg.finale)
    s["static_temp"] = int(cfg.staticTemp)
    s["write_conf_heat_capacities"] = int(cfg.writeConfHeatCapacitiesEnabled)
    return s


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


class Engine:
    """One tr_ctx: one CUDA device, one batch of chains."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.tr_create(int(device), C.byref(h))
        if rc != 0:
            raise TransRotError((self.L.tr_last_error(None) or b"").decode() or f"tr_create failed ({rc})")
        self.h = h
        self.system: Optional[System] = None
        self.n_chains = 0
        self.trace_moves = 0          # pending option
        self.live_trace_moves = 0     # what the live chain batch was allocated with
        self._owned = True

    @classmethod
    def _borrow(cls, handle, system=None):
        """An Engine view of a context owned by a MultiEngine."""
        e = cls.__new__(cls)
        e.L = load_library()
        e.h = C.c_void_p(handle)
        e.system = system
        e.n_chains = 0
        e.trace_moves = e.live_trace_moves = 0
        e._owned = False
        return e

    def close(self):
        if getattr(self, "h", None):
            if self._owned:
                self.L.tr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc: int):
        if rc != 0:
            raise TransRotError((self.L.tr_last_error(self.h) or b"").decode() or f"error {rc}")

    # ---- problem definition
    def set_stream(self, cuda_stream: int):
        self._ck(self.L.tr_set_stream(self.h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._ck(self.L.tr_synchronize(self.h))

    def set_system(self, system: System):
        c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
        arrs = dict(
            mol_site_off=c(system.mol_site_off, np.int32), mol_radius=c(system.mol_radius, np.float64),
            site_type=c(system.site_type, np.int32), site_q=c(system.site_q, np.float64),
            site_def=c(system.site_def, np.float64), site_mass=c(system.site_mass, np.float64),
            site_ghost=c(system.site_ghost, np.int32), moveable_idx=c(system.moveable_idx, np.int32),
            pair_abcd=c(system.pair_table, np.float64),
        )
        ts = TrSystem(n_mol=system.n_mol, n_sites=system.n_sites, n_types=system.n_types, n_moveable=len(system.moveable_idx),
                      **{k: v.ctypes.data for k, v in arrs.items()})
        self._ck(self.L.tr_set_system(self.h, C.byref(ts)))
        self.system = system
        self.n_chains = 0

    def set_schedules(self, schedules):
        """schedules: a Config, a sequence of Configs, or a SCHEDULE_DTYPE array."""
        if isinstance(schedules, Config):
            arr = schedule_from_config(schedules)
        elif isinstance(schedules, np.ndarray):
            arr = np.ascontiguousarray(schedules, dtype=SCHEDULE_DTYPE)
        else:
            arr = np.concatenate([schedule_from_config(c) for c in schedules])
        self.schedules = arr
        self._ck(self.L.tr_set_schedules(self.h, arr.ctypes.data, len(arr)))
        self.n_chains = 0

    def set_options(self, trace_moves: int = 0, record_points: bool = True, record_snapshots: bool = True, threads_per_chain: int = 0):
        self._ck(self.L.tr_set_options(self.h, int(trace_moves), int(record_points), int(record_snapshots), int(threads_per_chain)))
        self.trace_moves = int(trace_moves)

    def set_kernel(self, kernel: int = TR_KERNEL_AUTO, chains_per_block: int = 0, blocks_per_sm: int = 0):
        """Kernel family / launch geometry for the next init_chains_* (experiments and tests; 0 = automatic)."""
        self._ck(self.L.tr_set_kernel(self.h, int(kernel), int(chains_per_block), int(blocks_per_sm)))

    def kernel_name(self) -> str:
        return (self.L.tr_kernel_name(self.h) or b"").decode()

    def init_chains_seeded(self, seeds: Sequence[int], space_len: float, max_prop_failures: int,
                           sched_idx: Optional[Sequence[int]] = None, first_chain: int = 0):
        s = np.ascontiguousarray(seeds, dtype=np.int64)
        idx = None if sched_idx is None else np.ascontiguousarray(sched_idx, dtype=np.int32)
        if idx is not None and len(idx) != len(s):
            raise TransRotError(f"Error: {len(s)} seeds but {len(idx)} schedule indices")
        self.n_chains = 0
        self._ck(self.L.tr_init_chains_seeded(self.h, len(s), int(first_chain), s.ctypes.data, _ptr(idx), float(space_len), int(max_prop_failures)))
        self.n_chains = len(s)
        self.live_trace_moves = self.trace_moves

    def init_chains_explicit(self, n_chains: int, sites: np.ndarray, space_len: float, seeds: Optional[Sequence[int]] = None,
                             mt_state: Optional[np.ndarray] = None, sched_idx: Optional[Sequence[int]] = None, first_chain: int = 0):
        S = self.system.n_sites
        x = np.ascontiguousarray(sites, dtype=np.float64).reshape(-1, S, 3)
        sd = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.int64)
        st = None if mt_state is None else np.ascontiguousarray(mt_state, dtype=np.uint32).reshape(n_chains, 3, 625)
        idx = None if sched_idx is None else np.ascontiguousarray(sched_idx, dtype=np.int32)
        n_chains = int(n_chains)
        if x.shape[0] not in (1, n_chains):
            raise TransRotError(f"Error: {x.shape[0]} structures for {n_chains} chains (give 1 or one per chain)")
        for name, a in (("seeds", sd), ("schedule indices", idx)):
            if a is not None and len(a) != n_chains:
                raise TransRotError(f"Error: {len(a)} {name} for {n_chains} chains")
        self.n_chains = 0
        self._ck(self.L.tr_init_chains_explicit(self.h, int(n_chains), int(first_chain), x.ctypes.data, x.shape[0], _ptr(sd), _ptr(st), _ptr(idx), float(space_len)))
        self.n_chains = int(n_chains)
        self.live_trace_moves = self.trace_moves

    # ---- hot path
    def run(self, max_moves: int = 0):
        self._ck(self.L.tr_run(self.h, int(max_moves)))

    def last_run_ms(self) -> float:
        ms = C.c_float()
        self._ck(self.L.tr_last_run_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def launch_count(self) -> int:
        n = C.c_int64()
        self._ck(self.L.tr_launch_count(self.h, C.byref(n)))
        return int(n.value)

    # ---- results
    def summary(self) -> np.ndarray:
        out = np.zeros(self.n_chains, dtype=SUMMARY_DTYPE)
        self._ck(self.L.tr_get_summary(self.h, out.ctypes.data))
        return out

    def sites(self, with_com: bool = True):
        S, N = self.system.n_sites, self.system.n_mol
        x = np.zeros((self.n_chains, S, 3), dtype=np.float64)
        com = np.zeros((self.n_chains, N, 3), dtype=np.float64) if with_com else None
        self._ck(self.L.tr_get_sites(self.h, x.ctypes.data, _ptr(com)))
        return x, com

    def max_points(self):
        p, s = C.c_int64(), C.c_int64()
        self._ck(self.L.tr_max_points(self.h, C.byref(p), C.byref(s)))
        return int(p.value), int(s.value)

    def points(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Per-point records [n_chains, max_points].  `out`: a caller-owned buffer of that shape (e.g. page-locked memory, so
        that the copy runs at full PCIe speed), else a fresh array."""
        cap, _ = self.max_points()
        if out is None:
            out = np.zeros((self.n_chains, cap), dtype=POINT_DTYPE)
        assert out.shape == (self.n_chains, cap) and out.dtype == POINT_DTYPE and out.flags.c_contiguous
        self._ck(self.L.tr_get_points(self.h, out.ctypes.data, cap))
        return out

    def snapshots(self, with_sites: bool = True, out=None):
        """write(n) snapshots: (energy [n_chains, cap], tooth [n_chains, cap], sites [n_chains, cap, S, 3] or None).  `out`: a
        caller-owned (energy, tooth, sites) triple of those shapes to copy into (e.g. page-locked memory)."""
        _, cap = self.max_points()
        if out is not None:
            e, t, x = out
            assert e.shape == (self.n_chains, cap) and t.shape == (self.n_chains, cap) and e.dtype == np.float64 and t.dtype == np.int32
            assert x is None or (x.shape == (self.n_chains, cap, self.system.n_sites, 3) and x.dtype == np.float64 and x.flags.c_contiguous)
        else:
            e = np.zeros((self.n_chains, cap), dtype=np.float64)
            t = np.zeros((self.n_chains, cap), dtype=np.int32)
            x = np.zeros((self.n_chains, cap, self.system.n_sites, 3), dtype=np.float64) if with_sites else None
        self._ck(self.L.tr_get_snapshots(self.h, e.ctypes.data, t.ctypes.data, _ptr(x), cap))
        return e, t, x

    def set_record_movie(self, on: bool = True):
        """Keep the structure Space.writeMovie appends after every point (before init_chains_*)."""
        self._ck(self.L.tr_set_record_movie(self.h, int(bool(on))))

    def movie(self) -> np.ndarray:
        """[n_chains, points, S, 3]: frame p belongs to point record p."""
        cap, _ = self.max_points()
        out = np.zeros((self.n_chains, cap, self.system.n_sites, 3), dtype=np.float64)
        self._ck(self.L.tr_get_movie(self.h, out.ctypes.data, cap))
        return out

    def trace(self) -> np.ndarray:
        out = np.zeros((self.n_chains, self.live_trace_moves), dtype=TRACE_DTYPE)
        self._ck(self.L.tr_get_trace(self.h, out.ctypes.data))
        return out

    def rng_state(self) -> np.ndarray:
        out = np.zeros((self.n_chains, 3, 625), dtype=np.uint32)
        self._ck(self.L.tr_get_rng_state(self.h, out.ctypes.data))
        return out

    def pack_minima(self, device_ptr: int = 0) -> np.ndarray:
        out = np.zeros(self.n_chains, dtype=MINIMUM_DTYPE)
        self._ck(self.L.tr_pack_minima(self.h, out.ctypes.data, C.c_void_p(device_ptr) if device_ptr else None))
        return out

    def pack_minima_device(self, device_ptr: int):
        self._ck(self.L.tr_pack_minima(self.h, None, C.c_void_p(device_ptr)))

    def min_structure(self, chain: int):
        x = np.zeros((self.system.n_sites, 3), dtype=np.float64)
        e, t = C.c_double(), C.c_int32()
        self._ck(self.L.tr_get_min_structure(self.h, int(chain), x.ctypes.data, C.byref(e), C.byref(t)))
        return x, float(e.value), int(t.value)

    def min_structures(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        """[count, S, 3]: the write(min_tooth) structure of chains first .. first+count-1, packed on the device."""
        count = self.n_chains - first if count is None else count
        out = np.zeros((count, self.system.n_sites, 3), dtype=np.float64)
        self._ck(self.L.tr_pack_min_structures(self.h, int(first), int(count), out.ctypes.data, None))
        return out

    def pack_min_structures_device(self, device_ptr: int, first: int = 0, count: Optional[int] = None):
        count = self.n_chains - first if count is None else count
        self._ck(self.L.tr_pack_min_structures(self.h, int(first), int(count), None, C.c_void_p(device_ptr)))

    # ---- probes
    def total_energy(self, sites: np.ndarray) -> np.ndarray:
        S = self.system.n_sites
        x = np.ascontiguousarray(sites, dtype=np.float64).reshape(-1, S, 3)
        out = np.zeros(x.shape[0], dtype=np.float64)
        self._ck(self.L.tr_total_energy(self.h, x.ctypes.data, x.shape[0], out.ctypes.data))
        return out

    def delta_energy(self, sites: np.ndarray, mol: Sequence[int], trial: np.ndarray):
        """trial: [n, n_sites_of_mol_max<=16, 3] trial coordinates of particle mol[i] (rows beyond atoms.size() ignored)."""
        S = self.system.n_sites
        x = np.ascontiguousarray(sites, dtype=np.float64).reshape(-1, S, 3)
        n = x.shape[0]
        m = np.ascontiguousarray(mol, dtype=np.int32)
        t = np.zeros((n, TR_MAX_MOL_SITES, 3), dtype=np.float64)
        tr_in = np.asarray(trial, dtype=np.float64)
        t[:, : tr_in.shape[1], :] = tr_in
        es = np.zeros(n, dtype=np.float64)
        ee = np.zeros(n, dtype=np.float64)
        de = np.zeros(n, dtype=np.float64)
        self._ck(self.L.tr_delta_energy(self.h, x.ctypes.data, m.ctypes.data, t.ctypes.data, n, es.ctypes.data, ee.ctypes.data, de.ctypes.data))
        return es, ee, de

    def rng_words(self, seeds: Sequence[int], k: int) -> np.ndarray:
        s = np.ascontiguousarray(seeds, dtype=np.int64)
        out = np.zeros((len(s), 3, k), dtype=np.uint32)
        self._ck(self.L.tr_rng_words(self.h, s.ctypes.data, len(s), int(k), out.ctypes.data))
        return out

    def measure_fp64_peak(self):
        t, mhz = C.c_double(), C.c_double()
        self._ck(self.L.tr_measure_fp64_peak(self.h, C.byref(t), C.byref(mhz)))
        return float(t.value), float(mhz.value)


class MultiEngine:
    """One tr_multi: every GPU of a box driven from one host thread (csrc/multi.cu), the way a single-JVM host
    would.  Chains are sharded in contiguous global ranges; the final gather runs over NCCL."""

    def __init__(self, devices: Sequence[int]):
        self.L = load_library()
        d = (C.c_int32 * len(devices))(*[int(x) for x in devices])
        h = C.c_void_p()
        rc = self.L.tr_create_multi(d, len(devices), C.byref(h))
        if rc != 0:
            raise TransRotError((self.L.tr_multi_last_error(None) or b"").decode() or f"tr_create_multi failed ({rc})")
        self.h = h
        self.devices = list(devices)
        self.system: Optional[System] = None
        self.n_chains = 0
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            self.L.tr_destroy_multi(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc != 0:
            raise TransRotError((self.L.tr_multi_last_error(self.h) or b"").decode() or f"error {rc}")

    def chain_range(self, i: int, n_total: Optional[int] = None):
        lo, hi = C.c_int64(), C.c_int64()
        self.L.tr_multi_chain_range(int(i), len(self.devices), int(self.n_chains if n_total is None else n_total), C.byref(lo), C.byref(hi))
        return int(lo.value), int(hi.value)

    def ctx(self, i: int) -> Engine:
        """Per-device view for the tr_get_* calls (local chain j of device i is global chain lo_i + j)."""
        e = Engine._borrow(self.L.tr_multi_ctx(self.h, int(i)), self.system)
        lo, hi = self.chain_range(i)
        e.n_chains = hi - lo
        return e

    def set_system(self, system: System):
        c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
        arrs = dict(
            mol_site_off=c(system.mol_site_off, np.int32), mol_radius=c(system.mol_radius, np.float64),
            site_type=c(system.site_type, np.int32), site_q=c(system.site_q, np.float64),
            site_def=c(system.site_def, np.float64), site_mass=c(system.site_mass, np.float64),
            site_ghost=c(system.site_ghost, np.int32), moveable_idx=c(system.moveable_idx, np.int32),
            pair_abcd=c(system.pair_table, np.float64),
        )
        ts = TrSystem(n_mol=system.n_mol, n_sites=system.n_sites, n_types=system.n_types, n_moveable=len(system.moveable_idx),
                      **{k: v.ctypes.data for k, v in arrs.items()})
        self._ck(self.L.tr_multi_set_system(self.h, C.byref(ts)))
        self.system = system
        self.n_chains = 0

    def set_schedules(self, schedules):
        if isinstance(schedules, Config):
            arr = schedule_from_config(schedules)
        elif isinstance(schedules, np.ndarray):
            arr = np.ascontiguousarray(schedules, dtype=SCHEDULE_DTYPE)
        else:
            arr = np.concatenate([schedule_from_config(c) for c in schedules])
        self._ck(self.L.tr_multi_set_schedules(self.h, arr.ctypes.data, len(arr)))
        self.n_chains = 0

    def set_options(self, trace_moves: int = 0, record_points: bool = True, record_snapshots: bool = True, threads_per_chain: int = 0):
        self._ck(self.L.tr_multi_set_options(self.h, int(trace_moves), int(record_points), int(record_snapshots), int(threads_per_chain)))

    def init_chains_seeded(self, seeds: Sequence[int], space_len: float, max_prop_failures: int,
                           sched_idx: Optional[Sequence[int]] = None, first_chain: int = 0):
        s = np.ascontiguousarray(seeds, dtype=np.int64)
        idx = None if sched_idx is None else np.ascontiguousarray(sched_idx, dtype=np.int32)
        if idx is not None and len(idx) != len(s):
            raise TransRotError(f"Error: {len(s)} seeds but {len(idx)} schedule indices")
        self.n_chains = 0
        self._ck(self.L.tr_multi_init_chains_seeded(self.h, len(s), int(first_chain), s.ctypes.data, _ptr(idx), float(space_len), int(max_prop_failures)))
        self.n_chains = len(s)

    def run(self, max_moves: int = 0):
        self._ck(self.L.tr_multi_run(self.h, int(max_moves)))

    def synchronize(self):
        self._ck(self.L.tr_multi_synchronize(self.h))

    def last_run_ms(self) -> float:
        ms = C.c_float()
        self._ck(self.L.tr_multi_last_run_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def gather_minima(self, with_structure: bool = True):
        """(all records ordered by global chain id, winner record, winner structure [S, 3] or None)."""
        allm = np.zeros(self.n_chains, dtype=MINIMUM_DTYPE)
        best = np.zeros(1, dtype=MINIMUM_DTYPE)
        x = np.zeros((self.system.n_sites, 3), dtype=np.float64) if with_structure else None
        self._ck(self.L.tr_gather_minima(self.h, allm.ctypes.data, best.ctypes.data, _ptr(x)))
        return allm, best[0], x

    def gather_min_structures(self) -> np.ndarray:
        out = np.zeros((self.n_chains, self.system.n_sites, 3), dtype=np.float64)
        self._ck(self.L.tr_gather_min_structures(self.h, out.ctypes.data))
        return out
