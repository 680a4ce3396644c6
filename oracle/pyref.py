"""Independent pure-Python restatement of the hot path, used ONLY to cross-check the C oracle.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (no JVM in this image; see transrot_oracle.h).
It shares no code with oracle/transrot_oracle.c: the Mersenne Twister is numpy's
(numpy.random.RandomState seeded with init_by_array([hi32, lo32]) reproduces
MersenneTwister.setSeed(long)), and the control flow is re-derived from
the reference sources.  Slow by construction (Python loops): use for a few hundred moves.

Citations are into /root/reference/src/main/.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np

K_VALUE = 332.0637133
KB = 0.0019872


class PyMT:
    """MersenneTwister + BitsStreamGenerator on top of numpy's MT19937."""

    def __init__(self, seed: int):
        s = seed & 0xFFFFFFFFFFFFFFFF
        self.rs = np.random.RandomState(np.array([s >> 32, s & 0xFFFFFFFF], dtype=np.uint32))

    def next32(self) -> int:
        return int(self.rs.randint(0, 2**32, dtype=np.uint64))

    def next(self, bits: int) -> int:
        return self.next32() >> (32 - bits)

    def next_double(self) -> float:
        return float(((self.next(26) << 26) | self.next(26)) * 2.0**-52)

    def next_int(self, n: int) -> int:
        if n & -n == n:
            return (n * self.next(31)) >> 31
        while True:
            bits = self.next(31)
            val = bits % n
            # Java int overflow test `bits - val + (n-1) < 0`
            if ((bits - val + (n - 1)) & 0xFFFFFFFF) < 0x80000000:
                return val

    def next_long(self) -> int:
        v = (self.next32() << 32) | self.next32()
        return v - (1 << 64) if v >= (1 << 63) else v


class PySpace:
    """Space + Config + the three static generators, for one run."""

    def __init__(self, dbase, cfg, seed: int, pair_table: Optional[np.ndarray] = None):
        self.dbase = dbase
        self.cfg = cfg
        g = PyMT(seed)                        # Main.java:24,34
        self.rM = PyMT(g.next_long())         # Main.r
        self.rS = PyMT(g.next_long())         # Space.r
        self.rR = PyMT(g.next_long())         # Molecule.r
        self.spaceLen = cfg.spaceLen
        self.species: List[int] = []
        for name, count in cfg.particles:     # Space.add: radius-sorted, ties keep order
            sp = [p.name for p in dbase].index(name)
            for _ in range(count):
                for n, o in enumerate(self.species):
                    if dbase[sp].radius > dbase[o].radius:
                        self.species.insert(n, sp)
                        break
                else:
                    self.species.append(sp)
        self.T = sum(len(p.atoms) for p in dbase)
        if pair_table is None:
            atoms = [a for p in dbase for a in p.atoms]
            pair_table = np.zeros((self.T, self.T, 4))
            for a1 in atoms:
                for a2 in atoms:
                    pair_table[a1.type_id, a2.type_id] = (
                        math.sqrt(a1.a * a2.a), (a1.b + a2.b) / 2, math.sqrt(a1.c * a2.c), math.sqrt(a1.d * a2.d))
        self.pair = pair_table
        self.xyz: List[List[List[float]]] = []   # [mol][atom][3]
        self.com: List[List[float]] = []
        self.moveable: List[int] = []

    # ---- initial state (Space.java:72-117, Molecule.java:51-62,137-156)
    def _com(self, sp: int, atoms):
        tot = px = py = pz = 0.0
        for a, (x, y, z) in zip(self.dbase[sp].atoms, atoms):
            if a.ghost:
                continue
            tot += a.mass
            px += x * a.mass
            py += y * a.mass
            pz += z * a.mass
        return [px / tot, py / tot, pz / tot]

    def _propagate_once(self) -> bool:
        self.xyz, self.com, radii = [], [], []
        L = self.spaceLen
        for sp in self.species:
            c = 0
            while True:
                x = self.rS.next_double() * L - L / 2
                y = self.rS.next_double() * L - L / 2
                z = self.rS.next_double() * L - L / 2
                ok = True
                for (cx, cy, cz), r in zip(self.com, radii):
                    d = math.sqrt((cx - x) * (cx - x) + (cy - y) * (cy - y) + (cz - z) * (cz - z))
                    if d < self.dbase[sp].radius + r:
                        c += 1
                        ok = False
                        break
                if ok:
                    atoms = [[a.x + x, a.y + y, a.z + z] for a in self.dbase[sp].atoms]
                    self.xyz.append(atoms)
                    self.com.append(self._com(sp, atoms))
                    radii.append(self.dbase[sp].radius)
                    break
                if c >= self.cfg.maxPropFailures:
                    return False
        for m in range(len(self.species)):
            trial = self._rotate_temp(m, 2 * math.pi)
            if trial is not None:
                self.xyz[m] = trial[0]
        self.moveable = list(range(len(self.species)))
        return True

    def propagate(self):
        while not self._propagate_once():
            self.spaceLen *= 1.1

    # ---- moves
    def _rotate_temp(self, m: int, maxRot: float):
        """Molecule.rotateTemp: returns (trial coords, axis) and applies the (a-x)+x round trip."""
        atoms = self.xyz[m]
        if len(atoms) <= 1:
            return None
        rot = (self.rR.next_double() - 0.5) * 2 * maxRot
        c = self.rR.next_int(3)
        x, y, z = self.com[m]
        cs, sn = math.cos(rot), math.sin(rot)
        trial = []
        for a in atoms:
            ax, ay, az = a[0] - x, a[1] - y, a[2] - z
            if c == 0:
                t = [ax, cs * ay + sn * az, -1 * sn * ay + cs * az]
            elif c == 1:
                t = [cs * ax - sn * az, ay, sn * ax + cs * az]
            else:
                t = [cs * ax + sn * ay, -1 * sn * ax + cs * ay, az]
            a[0], a[1], a[2] = ax + x, ay + y, az + z
            trial.append([t[0] + x, t[1] + y, t[2] + z])
        return trial, c

    def _pair(self, ta, qa, pa, tb, qb, pb) -> float:
        A, B, C, D = self.pair[ta, tb]
        dx, dy, dz = pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]
        r = math.sqrt(dx * dx + dy * dy + dz * dz)
        return (K_VALUE * qa * qb / r) + A * math.exp(-1 * B * r) - (C / math.pow(r, 6)) + (D / math.pow(r, 12))

    def _mol_energy(self, m: int, coords_m, n: int) -> float:
        e = 0.0
        for a, pa in zip(self.dbase[self.species[m]].atoms, coords_m):
            for b, pb in zip(self.dbase[self.species[n]].atoms, self.xyz[n]):
                e += self._pair(a.type_id, a.q, pa, b.type_id, b.q, pb)
        return e

    def calc_energy(self) -> float:
        tot = 0.0
        N = len(self.species)
        for x in range(N):
            for y in range(N):
                if y > x:
                    tot += self._mol_energy(x, self.xyz[x], y)
        return tot

    def _decide(self, m: int, trial, temp: float, new_com):
        lim = self.spaceLen * 0.75
        for t in trial:
            if any(v > lim or v < self.spaceLen * -0.75 for v in t):
                return 0.0, 0
        eS = eE = 0.0
        for n in range(len(self.species)):
            if n != m:
                eS += self._mol_energy(m, self.xyz[m], n)
                eE += self._mol_energy(m, trial, n)
        if eE - eS > 0:
            rnd = self.rS.next_double()
            if temp == 0:
                return 0.0, 0
            if rnd > math.pow(math.e, -1 * (eE - eS) / (KB * temp)):
                return 0.0, 0
        self.xyz[m] = [list(t) for t in trial]
        if new_com is not None:
            self.com[m] = new_com
        return eE - eS, 1

    def move(self, m: int, maxD: float, temp: float):
        dx = (self.rS.next_double() * 2 * maxD) - maxD
        dy = (self.rS.next_double() * 2 * maxD) - maxD
        dz = (self.rS.next_double() * 2 * maxD) - maxD
        trial = [[a[0] + dx, a[1] + dy, a[2] + dz] for a in self.xyz[m]]
        c = self.com[m]
        return self._decide(m, trial, temp, [c[0] + dx, c[1] + dy, c[2] + dz])

    def rotate(self, m: int, maxRot: float, temp: float):
        trial, _ = self._rotate_temp(m, maxRot)
        return self._decide(m, trial, temp, None)

    # ---- Main.sawtoothAnneal
    def sawtooth_anneal(self, startingEnergy: float, max_moves: int):
        """Returns the list of (mol, kind, accepted, running energy, t) for the first max_moves moves."""
        C = self.cfg
        numTeeth, pts = (1, 1) if C.staticTemp else (C.numTeeth, C.ptsPerTooth)
        t = saveT = C.maxTemp
        double = False
        pR, pT, maxD, maxRot = C.magwalkProbRot, C.magwalkProbTrans, C.maxTrans, C.maxRot
        out = []
        x = 0
        while x < numTeeth:
            delT = t / (pts - 1) if pts != 1 else (math.inf if t > 0 else math.nan)
            for _y in range(pts):
                for _z in range(C.movePerPt):
                    m = self.moveable[self.rS.next_int(len(self.moveable))]
                    if self.rM.next_double() >= 0.5 and len(self.xyz[m]) > 1:
                        kind = 1
                        if self.rM.next_double() >= pR:
                            dE, acc = self.rotate(m, maxRot, t)
                        else:
                            dE, acc = self.rotate(m, 2 * math.pi, t)
                    else:
                        kind = 0
                        if self.rM.next_double() >= pT:
                            dE, acc = self.move(m, maxD, t)
                        else:
                            dE, acc = self.move(m, maxD * C.magwalkFactorTrans, t)
                    startingEnergy += dE
                    out.append((m, kind, acc, startingEnergy, t))
                    if len(out) >= max_moves:
                        return out
                if C.writeConfHeatCapacitiesEnabled and abs(t) < 0.005:
                    continue
                if not C.staticTemp:
                    t -= delT
                if t < 0:
                    t = 0
            saveT *= C.tempDecreasePerTooth
            t = saveT
            pts += C.ptsAddedPerTooth
            if x == numTeeth - 1 and not double and C.finale:
                t = 0
                double = True
                x -= 1
                maxD /= 100
                maxRot /= 100
                pR = pT = 0
            x += 1
        return out
