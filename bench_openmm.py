"""bench_openmm.py — the OpenMM arm of bench.py.

Runs the UNMODIFIED reference: `multimm.config.SimulationConfig` -> `multimm.model.MultiMM` ->
its own `set_radiuses / initialize_simulation / add_forcefield` (model.py:722-857) build the
`openmm.System`; then exactly what `min_energy` does (model.py:862-886) — Platform lookup,
`Simulation(...)`, `context.setPositions`, `minimizeEnergy()` — with timers around it.  None of this
repo's engine is on that path.

It needs `import openmm` AND `import multimm` (the reference package, e.g. pip-installed with
`--target baseline/_ref`, together with its import-time dependencies).  Neither exists in the build
image (no wheel, no network): `available()` says which import failed and bench.py falls back to the
CPU oracle port.  The control flow below is exercised on the CPU with stand-in modules
(tests/test_bench_contract.py); it lights up unchanged the day a box has OpenMM.
"""
from __future__ import annotations

import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def available() -> tuple[bool, str]:
    """(usable, one-line reason)."""
    if os.path.isdir(REF_DIR) and REF_DIR not in sys.path:
        sys.path.append(REF_DIR)
    for mod in ("openmm", "multimm.config", "multimm.model"):
        try:
            importlib.import_module(mod)
        except Exception as e:  # ModuleNotFoundError here; anything else is reported as it is
            return False, f"import {mod} failed ({type(e).__name__}: {e})"
    return True, "openmm and the reference package import"


def platforms() -> list[str]:
    import openmm as mm

    return [mm.Platform.getPlatform(i).getName() for i in range(mm.Platform.getNumPlatforms())]


def _build(config_kw: dict, platform_name: str, threads: int):
    """The reference's own objects, up to the point where min_energy would call minimizeEnergy()."""
    import openmm as mm
    from multimm.config import SimulationConfig as RefConfig
    from multimm.model import MultiMM as RefMultiMM
    from openmm.app import Simulation

    kw = dict(config_kw)
    kw["PLATFORM"] = platform_name
    if platform_name == "CPU":
        kw["CPU_THREADS"] = threads
    args = RefConfig(**kw)
    m = RefMultiMM(args)
    m.set_radiuses()
    m.initialize_simulation()
    m.add_forcefield()
    # model.py:862-878, verbatim in behaviour (no fallback here: a missing platform is reported)
    platform = mm.Platform.getPlatformByName(platform_name)
    if platform_name == "CPU":
        platform.setPropertyDefaultValue("Threads", str(threads))
    sim = Simulation(m.pdb.topology, m.system, m.integrator, platform)
    sim.context.setPositions(m.pdb.positions)
    return m, sim


def time_arm(config_kw: dict, platform_name: str, steps: int, warmup: int, threads: int,
             minimize_cap_s: float = 120.0) -> dict:
    """Force evaluations/s (setPositions + getState(energy, forces): what LocalEnergyMinimizer does
    once per evaluation) and minimizeEnergy() wall time with OpenMM's defaults (tolerance
    10 kJ/mol/nm, unlimited iterations), stopped by a MinimizationReporter after `minimize_cap_s`."""
    import openmm as mm

    t0 = time.perf_counter()
    m, sim = _build(config_kw, platform_name, threads)
    build_s = time.perf_counter() - t0
    pos = m.pdb.positions
    ctx = sim.context

    def one():
        ctx.setPositions(pos)
        st = ctx.getState(getEnergy=True, getForces=True)
        return st.getPotentialEnergy()

    for _ in range(max(warmup, 1)):
        e0 = one()
    t0 = time.perf_counter()
    for _ in range(steps):
        e0 = one()
    dt = time.perf_counter() - t0
    out = dict(platform=ctx.getPlatform().getName(), threads=threads if platform_name == "CPU" else None,
               force_evals_per_s=steps / dt, ms_per_eval=1e3 * dt / steps, build_seconds=build_s,
               energy_start_kj_mol=e0.value_in_unit(mm.unit.kilojoule_per_mole))

    capped = {"hit": False, "iterations": 0}
    reporter = None
    if hasattr(mm, "MinimizationReporter"):  # OpenMM >= 8.1
        t_start = time.perf_counter()

        class Cap(mm.MinimizationReporter):
            def report(self, iteration, x, grad, args):
                capped["iterations"] = iteration
                if time.perf_counter() - t_start > minimize_cap_s:
                    capped["hit"] = True
                    return True
                return False

        reporter = Cap()
    ctx.setPositions(pos)
    t0 = time.perf_counter()
    if reporter is not None:
        sim.minimizeEnergy(reporter=reporter)
    else:
        sim.minimizeEnergy()
    wall = time.perf_counter() - t0
    e1 = ctx.getState(getEnergy=True).getPotentialEnergy().value_in_unit(mm.unit.kilojoule_per_mole)
    out["minimize"] = dict(wall_seconds=wall, capped=capped["hit"], cap_seconds=minimize_cap_s,
                           iterations_seen=capped["iterations"], energy_final_kj_mol=e1)
    return out
