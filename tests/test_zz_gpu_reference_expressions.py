"""GPU energies against the REFERENCE's own expressions: the golden of
tests/golden/make_golden_forcefield.py (the reference's unmodified add_* methods, their Lepton
strings evaluated in FP64) compared with the engine driven by this repo's host mirror.  (The file
sorts last on purpose: it re-checks, against an independent source, what test_gpu_parity.py checks
against the oracle.)"""
import json
import os

import numpy as np
import pytest

from common import O
from multimm_b200 import model
from multimm_b200.config import SimulationConfig

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
G = np.load(os.path.join(GOLD, "forcefield_golden.npz"))
AUDIT = json.load(open(os.path.join(GOLD, "forcefield_golden_expressions.json")))
ALL_ON = dict(EV_USE_EXCLUDED_VOLUME=True, COB_USE_COMPARTMENT_BLOCKS=True, SCB_USE_SUBCOMPARTMENT_BLOCKS=True,
              CHB_USE_CHROMOSOMAL_BLOCKS=True, SC_USE_SPHERICAL_CONTAINER=True, IBL_USE_B_LAMINA_INTERACTION=True,
              CF_USE_CENTRAL_FORCE=True, POL_USE_HARMONIC_BOND=True, LE_USE_HARMONIC_BOND=True, POL_USE_HARMONIC_ANGLE=True)
# default forms run on the Newton-3 kernel, the alternate forms on the generic gather path (FP64
# body): the north star's 1e-5 for every one of them
TOL = {"default_forms": 1e-5, "fixed_loop_distances": 1e-5, "alt1": 1e-5, "alt2": 1e-5, "alt3": 1e-5}


@pytest.mark.parametrize("case", list(AUDIT))
def test_engine_matches_the_reference_expressions(built_lib, case):
    from multimm_b200.engine import Engine

    n = len(G["x"])
    args = SimulationConfig(PLATFORM="B200", LOOPS_PATH="unused.bedpe", OUT_PATH="/tmp/unused", N_BEADS=n, **ALL_ON,
                            **AUDIT[case]["overrides"])
    m = model.MultiMM.__new__(model.MultiMM)
    m.args, m.engine, m.timings = args, Engine(n), {}
    m.chr_ends, m.Cs, m.ms, m.ns, m.ds = G["chr_ends"], G["Cs"], G["ms"], G["ns"], G["ds"]
    m.chrom_spin, m.chrom_strength = G["chrom_spin"], G["chrom_strength"]
    m.set_radiuses()
    m.mass_center = np.average(G["x"], axis=0)
    m.add_forcefield()
    m.engine.set_positions(G["x"])
    e, f = m.engine.energy_forces()
    m.engine.close()
    want = G[f"{case}.energies"]
    for t, name in enumerate(O.TERM_NAMES):
        assert abs(e[t] - want[t]) <= TOL[case] * max(abs(want[t]), 1e-12) + 1e-9, (case, name, e[t], want[t])
    assert np.isfinite(f).all()
