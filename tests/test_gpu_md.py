"""MD relaxation on the GPU against numpy restatements of the integrators
([OpenMM] LangevinIntegrator / VerletIntegrator / BrownianIntegrator) driven by the oracle's forces
and by the same Philox4x32-10 noise stream."""
import os

import numpy as np
import pytest

from common import O, make_case, to_engine, to_oracle

pytestmark = pytest.mark.gpu

KB = 0.008314462618
MASS = 16427.889


from md_ref import normals3  # noqa: E402  (numpy restatement of the device noise stream)


def test_initial_velocities_are_maxwell_boltzmann_and_reproducible(built_lib):
    case = make_case(20000, terms=("EV",), seed=1)
    eng = to_engine(case)
    eng.set_velocities_to_temperature(310.0, seed=7)
    v = eng.get_velocities()
    assert np.allclose(v, np.sqrt(KB * 310.0 / MASS) * normals3(7, 20000, 0, 0), rtol=1e-12, atol=1e-15)
    temp = MASS * (v ** 2).sum() / (3 * 20000 * KB)
    assert abs(temp - 310.0) < 0.02 * 310.0  # 1/sqrt(3N/2) = 0.6 %
    eng.set_velocities_to_temperature(310.0, seed=8)
    assert not np.array_equal(v, eng.get_velocities())
    eng.close()


@pytest.mark.parametrize("integrator", ["verlet", "langevin", "brownian"])
def test_steps_match_numpy_restatement(built_lib, integrator):
    """Three steps of each integrator: positions and velocities against numpy with the ORACLE's forces."""
    case = make_case(400, n_chrom=2, seed=5, noise=0.02)
    sysd = to_oracle(case)
    eng = to_engine(case)
    dt, temp, gamma, seed = 0.002, 310.0, 0.5, 11
    eng.md_configure(integrator, dt, temp, gamma, MASS, seed)
    eng.set_velocities_to_temperature(temp, seed)
    x, v = case["x"].copy(), eng.get_velocities()
    for step in range(3):
        f = O.energy_forces(sysd, x)[1]
        nrm = normals3(seed, 400, step, 1)
        if integrator == "verlet":
            v = v + dt * f / MASS
            x = x + dt * v
        elif integrator == "langevin":
            a = np.exp(-gamma * dt)
            v = a * v + (1 - a) / gamma * f / MASS + np.sqrt(KB * temp * (1 - a * a) / MASS) * nrm
            x = x + dt * v
        else:
            dx = dt / (gamma * MASS) * f + np.sqrt(2 * KB * temp * dt / (gamma * MASS)) * nrm
            x, v = x + dx, dx / dt
    rep = eng.md_run(3)
    assert rep["step"] == 3
    # displacements are ~1e-4 nm per step; GPU forces carry the usual 1e-4 relative error
    assert np.abs(eng.get_positions() - x).max() < 1e-7
    assert np.abs(eng.get_velocities() - v).max() < 1e-4 * np.abs(v).max()
    e_ref = O.energy_forces(sysd, eng.get_positions(), want_forces=False)[0].sum()
    assert abs(rep["potential"] - e_ref) <= 1e-5 * abs(e_ref)
    assert rep["kinetic"] == pytest.approx(0.5 * MASS * (eng.get_velocities() ** 2).sum(), rel=1e-12)
    eng.close()


@pytest.mark.parametrize("e_boost", ["above", "below"])
def test_amd_steps_match_numpy_restatement(built_lib, e_boost):
    """[OpenMM] amd.AMDIntegrator (model.py:794-800): leapfrog with the force scaled by
    (alpha / (alpha + E - V))^2 while the total potential energy V is at or below E, unscaled above it.
    Three steps against numpy with the ORACLE's forces and energies, both branches of step(E - V)."""
    case = make_case(400, n_chrom=2, seed=6, noise=0.02)
    sysd = to_oracle(case)
    eng = to_engine(case)
    e0 = O.energy_forces(sysd, case["x"])[0].sum()
    alpha = 0.05 * abs(e0)
    e_thr = e0 + 0.3 * abs(e0) if e_boost == "above" else e0 - 0.3 * abs(e0)
    dt, temp, seed = 0.002, 310.0, 12
    eng.md_configure("amd", dt, temp, 0.0, MASS, seed, amd_alpha=alpha, amd_e=e_thr)
    eng.set_velocities_to_temperature(temp, seed)
    x, v = case["x"].copy(), eng.get_velocities()
    boosts = []
    for step in range(3):
        e_terms, f = O.energy_forces(sysd, x)
        pot = e_terms.sum()
        boost = (alpha / (alpha + e_thr - pot)) ** 2 if e_thr - pot >= 0 else 1.0
        boosts.append(boost)
        v = v + dt * f * boost / MASS
        x = x + dt * v
    assert all(b < 0.05 for b in boosts) if e_boost == "above" else boosts == [1.0] * 3
    rep = eng.md_run(3)
    assert rep["step"] == 3
    assert np.abs(eng.get_positions() - x).max() < 1e-7
    assert np.abs(eng.get_velocities() - v).max() < 1e-4 * np.abs(v).max()
    eng.close()


def test_amd_defaults_are_the_reference_config_values_and_bad_alpha_is_refused(built_lib):
    case = make_case(300, n_chrom=1, seed=2, terms=("EV", "BOND"))
    eng = to_engine(case)
    eng.md_configure("amd", 0.001, 310.0, 0.0, MASS, 1)  # SIM_AMD_ALPHA = 100, SIM_AMD_E = 1000 (config.py:255-256)
    assert eng.md_run(2)["step"] == 2
    with pytest.raises(Exception, match="alpha"):
        eng.md_configure("amd", 0.001, 310.0, 0.0, MASS, 1, amd_alpha=0.0, amd_e=10.0)
    eng.close()


def test_verlet_conserves_energy(built_lib):
    case = make_case(1500, n_chrom=2, seed=9, terms=("EV", "SCB", "SC", "BOND", "LOOP", "ANGLE"))
    eng = to_engine(case)
    eng.minimize(10.0, 300)
    eng.md_configure("verlet", 0.001, 310.0, 0.0, MASS, 3)
    eng.set_velocities_to_temperature(310.0, 3)
    tot = []
    for _ in range(10):
        rep = eng.md_run(100)
        tot.append(rep["potential"] + rep["kinetic"])
    kin = rep["kinetic"]
    assert max(tot) - min(tot) < 0.05 * kin, (tot, kin)  # leapfrog: bounded fluctuation, no drift
    eng.close()


def test_langevin_thermostat_holds_the_temperature(built_lib):
    case = make_case(3000, n_chrom=2, seed=10, terms=("EV", "SC", "BOND", "ANGLE"))
    eng = to_engine(case)
    eng.minimize(10.0, 200)
    eng.md_configure("langevin", 0.001, 310.0, 20.0, MASS, 4)  # strong friction: equilibrates in ~0.1 ps
    temps = [eng.md_run(200)["temperature"] for _ in range(10)]
    assert abs(np.mean(temps[3:]) - 310.0) < 0.05 * 310.0, temps
    eng.close()


def test_driver_runs_md_and_writes_the_reference_outputs(built_lib, tmp_path):
    from multimm_b200 import cif, run

    gold = os.path.join(os.path.dirname(__file__), "golden")
    ini = tmp_path / "c.ini"
    out = tmp_path / "out"
    ini.write_text(f"[Main]\nPLATFORM = B200\nN_BEADS = 3000\nLOOPS_PATH = {gold}/synthetic_loops.bedpe\n"
                   f"OUT_PATH = {out}\nSAVE_PLOTS = False\nMIN_MAX_ITERATIONS = 200\nSIM_RUN_MD = True\n"
                   "SIM_N_STEPS = 300\nSIM_SAMPLING_STEP = 100\nTRJ_FRAMES = 6\n")
    assert run.main(["-c", str(ini)]) == 0
    for rel in ("model/MultiMM_afterMD.cif", "md_frames/frame_1.cif", "md_frames/frame_3.cif",
                "metadata/MultiMM_annealing.dcd", "metadata/md_thermodynamics.tsv"):
        assert (out / rel).exists(), rel
    frames = cif.read_dcd(str(out / "metadata/MultiMM_annealing.dcd"))
    assert frames.shape == (6, 3000, 3)
    last = cif.read_cif_coordinates(str(out / "model/MultiMM_afterMD.cif"), include_hetatm=True)
    assert np.allclose(frames[-1], last, atol=2e-3)
    rows = (out / "metadata/md_thermodynamics.tsv").read_text().strip().split("\n")
    assert len(rows) == 4 and rows[0].startswith("step")


def test_md_steps_without_report_give_the_same_trajectory(built_lib):
    """run_md asks for energies only where a sample is recorded (every chunk paid one more
    force evaluation and a host read).  Chunks without a report must not change the trajectory."""
    from common import make_case, to_engine

    case = make_case(1500, n_chrom=2, seed=8)
    out = []
    for silent in (False, True):
        eng = to_engine(case)
        eng.md_configure("langevin", 0.001, 310.0, 0.5, 16427.889, seed=4)
        eng.set_velocities_to_temperature(310.0, 4)
        n0 = eng.launch_count
        for _ in range(4):
            eng.md_run(5, want_report=not silent)
        rep = eng.md_run(5)
        out.append((rep, eng.get_positions(), eng.launch_count - n0))
        eng.close()
    assert out[0][0] == out[1][0] and np.array_equal(out[0][1], out[1][1])
    assert out[1][2] < out[0][2]  # four evaluations (and their reductions) fewer
