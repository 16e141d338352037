"""ctypes wrapper of the CPU oracle (oracle/mmm_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module; nothing under ``multimm_b200/`` does.  PARITY UNPINNED — see mmm_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")

NUM_TERMS = 10
TERM_NAMES = ("EV", "COB", "SCB", "CHB", "SC", "LAM", "CF", "BOND", "LOOP", "ANGLE")


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc only; a few seconds)."""
    src = os.path.join(HERE, "mmm_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return LIB_PATH


class _Params(C.Structure):
    _fields_ = [
        ("n", C.c_int64),
        ("ev_form", C.c_int32), ("ev", C.c_double * 4),
        ("cob_form", C.c_int32), ("cob", C.c_double * 3),
        ("scb_form", C.c_int32), ("scb", C.c_double * 5),
        ("chb_form", C.c_int32), ("chb", C.c_double * 2),
        ("sc_form", C.c_int32), ("sc", C.c_double * 6),
        ("lam_form", C.c_int32), ("lam", C.c_double * 6),
        ("cf_form", C.c_int32), ("cf", C.c_double * 5),
        ("loop_form", C.c_int32),
        ("cutoff", C.c_double),
        ("s", C.c_void_p), ("chrom", C.c_void_p), ("cstr", C.c_void_p),
        ("nb", C.c_int64), ("bi", C.c_void_p), ("bj", C.c_void_p), ("br0", C.c_void_p), ("bk", C.c_void_p),
        ("nl", C.c_int64), ("li", C.c_void_p), ("lj", C.c_void_p), ("lr0", C.c_void_p), ("lk", C.c_void_p),
        ("na", C.c_int64), ("ai", C.c_void_p), ("aj", C.c_void_p), ("ak", C.c_void_p),
        ("at0", C.c_void_p), ("akt", C.c_void_p),
    ]


class _Report(C.Structure):
    _fields_ = [
        ("iterations", C.c_int64), ("evaluations", C.c_int64),
        ("e_initial", C.c_double), ("e_final", C.c_double), ("rms_force", C.c_double),
        ("converged", C.c_int32), ("ls_status", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_energy_forces.restype = C.c_int
        _lib.orc_energy_forces.argtypes = [C.POINTER(_Params), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib.orc_threads_used.restype = C.c_int
        _lib.orc_threads_used.argtypes = [C.c_int]
        _lib.orc_count_pairs.restype = C.c_int64
        _lib.orc_count_pairs.argtypes = [C.POINTER(_Params), C.c_void_p]
        _lib.orc_minimize.restype = C.c_int
        _lib.orc_minimize.argtypes = [C.POINTER(_Params), C.c_void_p, C.c_double, C.c_int64, C.c_int,
                                      C.POINTER(_Report), C.c_int]
        _lib.orc_hilbert_points.restype = None
        _lib.orc_hilbert_points.argtypes = [C.c_int64, C.c_int, C.c_void_p]
        _lib.orc_backbone_bonds.restype = C.c_int64
        _lib.orc_backbone_bonds.argtypes = [C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
        _lib.orc_backbone_angles.restype = C.c_int64
        _lib.orc_backbone_angles.argtypes = [C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
        _lib.orc_cell_list.restype = None
        _lib.orc_cell_list.argtypes = [C.c_int64, C.c_void_p, C.c_float, C.c_int32, C.c_float,
                                       C.c_void_p, C.c_void_p]
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class System:
    """Plain description of one MultiMM system: what the add_* methods of model.py build."""

    n: int
    ev: tuple | None = None        # (form, [epsilon, r_small, sigma, power])
    cob: tuple | None = None       # (form, [rc, Ea, Eb])
    scb: tuple | None = None       # (form, [rsc, Ea1, Ea2, Eb1, Eb2])
    chb: tuple | None = None       # (form, [kC, dE])
    sc: tuple | None = None        # (form, [C, R1, R2, x0, y0, z0])
    lam: tuple | None = None       # (form, [B, R1, R2, x0, y0, z0])
    cf: tuple | None = None        # (form, [G, R1, x0, y0, z0])
    loop_form: int = 0
    cutoff: float = 0.0
    s: np.ndarray | None = None
    chrom: np.ndarray | None = None
    cstr: np.ndarray | None = None
    bonds: tuple | None = None     # (i, j, r0, k)
    loops: tuple | None = None     # (i, j, r0, k)
    angles: tuple | None = None    # (i, j, k, theta0, ktheta)
    _keep: list = field(default_factory=list, repr=False)

    def _c(self) -> _Params:
        p = _Params()
        p.n = int(self.n)
        self._keep = []

        def setg(name, val, width):
            if val is None:
                setattr(p, name + "_form", -1)
            else:
                form, g = val
                setattr(p, name + "_form", int(form))
                arr = getattr(p, name)
                g = list(g) + [0.0] * (width - len(g))
                for q in range(width):
                    arr[q] = float(g[q])

        setg("ev", self.ev, 4)
        setg("cob", self.cob, 3)
        setg("scb", self.scb, 5)
        setg("chb", self.chb, 2)
        setg("sc", self.sc, 6)
        setg("lam", self.lam, 6)
        setg("cf", self.cf, 5)
        p.loop_form = int(self.loop_form)
        p.cutoff = float(self.cutoff)

        def keep(a, dt):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dt)
            self._keep.append(a)
            return _ptr(a)

        p.s = keep(self.s, np.int8)
        p.chrom = keep(self.chrom, np.int32)
        p.cstr = keep(self.cstr, np.float64)
        if self.bonds is not None and len(self.bonds[0]):
            p.nb = len(self.bonds[0])
            p.bi, p.bj = keep(self.bonds[0], np.int32), keep(self.bonds[1], np.int32)
            p.br0, p.bk = keep(self.bonds[2], np.float64), keep(self.bonds[3], np.float64)
        if self.loops is not None and len(self.loops[0]):
            p.nl = len(self.loops[0])
            p.li, p.lj = keep(self.loops[0], np.int32), keep(self.loops[1], np.int32)
            p.lr0, p.lk = keep(self.loops[2], np.float64), keep(self.loops[3], np.float64)
        if self.angles is not None and len(self.angles[0]):
            p.na = len(self.angles[0])
            p.ai, p.aj, p.ak = (keep(self.angles[q], np.int32) for q in range(3))
            p.at0, p.akt = keep(self.angles[3], np.float64), keep(self.angles[4], np.float64)
        return p


def energy_forces(sysd: System, x: np.ndarray, want_forces: bool = True, nthreads: int = 0):
    """Per-term energies (10,) and forces (N,3) at positions x (N,3) nm."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    assert x.shape == (sysd.n, 3)
    p = sysd._c()
    e = np.zeros(NUM_TERMS)
    f = np.zeros_like(x) if want_forces else None
    lib().orc_energy_forces(C.byref(p), _ptr(x), _ptr(e), _ptr(f), int(nthreads))
    return e, f


def host_threads() -> int:
    """Hardware threads this process may use (affinity mask, not OMP_NUM_THREADS)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def threads_used(nthreads: int = 0) -> int:
    """How many OpenMP threads a call with this `nthreads` argument really runs on."""
    return int(lib().orc_threads_used(int(nthreads)))


def count_pairs(sysd: System, x: np.ndarray) -> int:
    x = np.ascontiguousarray(x, dtype=np.float64)
    p = sysd._c()
    return int(lib().orc_count_pairs(C.byref(p), _ptr(x)))


def minimize(sysd: System, x: np.ndarray, tol: float = 10.0, max_iter: int = 0, nthreads: int = 0):
    """liblbfgs restatement; returns (x_min, report dict)."""
    x = np.array(x, dtype=np.float64, order="C", copy=True)
    p = sysd._c()
    rep = _Report()
    lib().orc_minimize(C.byref(p), _ptr(x), float(tol), int(max_iter), 1, C.byref(rep), int(nthreads))
    return x, {k: getattr(rep, k) for k, _ in _Report._fields_}


def hilbert_points(n: int, p: int = 8) -> np.ndarray:
    out = np.zeros((n, 3), dtype=np.int32)
    lib().orc_hilbert_points(int(n), int(p), _ptr(out))
    return out


def backbone_bonds(n: int, chr_ends) -> np.ndarray:
    ce = np.ascontiguousarray(chr_ends, dtype=np.int64)
    out = np.zeros(max(n, 1), dtype=np.int32)
    c = lib().orc_backbone_bonds(int(n), _ptr(ce), len(ce), _ptr(out))
    return out[:c].copy()


def backbone_angles(n: int, chr_ends) -> np.ndarray:
    ce = np.ascontiguousarray(chr_ends, dtype=np.int64)
    out = np.zeros(max(n, 1), dtype=np.int32)
    c = lib().orc_backbone_angles(int(n), _ptr(ce), len(ce), _ptr(out))
    return out[:c].copy()


def cell_list(xyzc: np.ndarray, cell: float, dim: int, origin: float):
    xyzc = np.ascontiguousarray(xyzc, dtype=np.float32)
    n = xyzc.shape[0]
    keys = np.zeros(n, dtype=np.uint32)
    order = np.zeros(n, dtype=np.int32)
    lib().orc_cell_list(n, _ptr(xyzc), np.float32(cell), int(dim), np.float32(origin), _ptr(keys), _ptr(order))
    return keys, order
