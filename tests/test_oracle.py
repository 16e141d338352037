"""Pins the CPU oracle (oracle/mmm_oracle.c).  The reference's own tests hold no numeric vectors
for this path (tests/test_simulations.py checks file existence only) and OpenMM / hilbertcurve
are not installed, so the pins are: closed-form known answers derived from the energy strings
in src/multimm/model.py, finite differences in FP64, and geometric invariants of the Hilbert
curve."""
import numpy as np
import pytest

from common import DEF, O, backbone, make_case, radii, to_oracle


# ---------------------------------------------------------------------------------------------
# closed-form known answers
# ---------------------------------------------------------------------------------------------
def two(x=0.1):
    return np.array([[0.0, 0.0, 0.0], [x, 0.0, 0.0]])


def test_kat_ev_powerlaw():
    # model.py:199  epsilon*(sigma/(r+r_small))^EV_POWER at r = sigma = 0.1: 100*(0.1/0.15)^6
    e, f = O.energy_forces(O.System(n=2, ev=(0, [100.0, 0.05, 0.1, 6.0])), two(0.1))
    assert e[0] == pytest.approx(100.0 * (0.1 / 0.15) ** 6, rel=1e-14)
    assert e[0] == pytest.approx(8.779149519890261, rel=1e-13)
    # dE/dr = -6 E / (r + rs): repulsive, along -x on bead 0
    assert f[0, 0] == pytest.approx(-6.0 * e[0] / 0.15, rel=1e-13)
    assert np.allclose(f[0], -f[1])


def test_kat_ev_gaussian_core():
    e, _ = O.energy_forces(O.System(n=2, ev=(1, [100.0, 0.05, 0.1, 6.0])), two(0.1))
    assert e[0] == pytest.approx(100.0 * np.exp(-0.5), rel=1e-14)  # model.py:209


def test_kat_bond_and_loop():
    # [OpenMM] HarmonicBondForce 1/2 k (r-r0)^2 at r = 0.11: 1/2 * 3e5 * 1e-4 = 15
    e, f = O.energy_forces(O.System(n=2, bonds=([0], [1], [0.1], [3.0e5])), two(0.11))
    assert e[7] == pytest.approx(15.0, rel=1e-12)
    assert f[0, 0] == pytest.approx(3.0e5 * 0.01, rel=1e-12)  # pulled towards bead 1
    e, _ = O.energy_forces(O.System(n=2, loops=([0], [1], [0.2], [3.0e4])), two(0.3))
    assert e[8] == pytest.approx(0.5 * 3.0e4 * 0.01, rel=1e-12)
    # fene_soft (model.py:665, no 1/2): k d^2/(1 + d^2/r0^2)
    e, _ = O.energy_forces(O.System(n=2, loops=([0], [1], [0.2], [3.0e4]), loop_form=1), two(0.3))
    assert e[8] == pytest.approx(3.0e4 * 0.01 / (1 + 0.01 / 0.04), rel=1e-12)
    # gaussian_tether (model.py:686): k (1 - exp(-d^2/(r0/2)^2))
    e, _ = O.energy_forces(O.System(n=2, loops=([0], [1], [0.2], [3.0e4]), loop_form=2), two(0.3))
    assert e[8] == pytest.approx(3.0e4 * (1 - np.exp(-0.01 / 0.01)), rel=1e-12)


def test_kat_angle():
    # 90 degree corner of the Hilbert lattice: 1/2 * 100 * (pi/2)^2
    x = np.array([[0.1, 0, 0], [0, 0, 0], [0, 0.1, 0.0]])
    e, f = O.energy_forces(O.System(n=3, angles=([0], [1], [2], [np.pi], [100.0])), x)
    assert e[9] == pytest.approx(0.5 * 100.0 * (np.pi / 2) ** 2, rel=1e-14)
    assert e[9] == pytest.approx(123.37005501361698, rel=1e-13)
    # torque-free, force-free in total
    assert np.allclose(f.sum(axis=0), 0.0, atol=1e-9)
    assert np.allclose(np.cross(x, f).sum(axis=0), 0.0, atol=1e-9)
    # straight segment: theta = pi, zero energy and zero force
    x = np.array([[-0.1, 0, 0], [0, 0, 0], [0.1, 0, 0.0]])
    e, f = O.energy_forces(O.System(n=3, angles=([0], [1], [2], [np.pi], [100.0])), x)
    assert e[9] == 0.0 and np.all(f == 0.0)


def test_kat_compartment_blocks():
    r, rc = 0.2, 0.15
    g = np.exp(-r * r / (2 * rc * rc))
    def ecob(s1, s2, form=0):
        sysd = O.System(n=2, cob=(form, [rc, 1.0, 2.0]), s=np.array([s1, s2], dtype=np.int8))
        return O.energy_forces(sysd, two(r))[0][1]
    # model.py:246-250: A-A -> Ea, B-B -> Eb, mixed / unlabelled -> 0
    assert ecob(1, 2) == pytest.approx(-1.0 * g)
    assert ecob(2, 2) == pytest.approx(-1.0 * g)
    assert ecob(-1, -2) == pytest.approx(-2.0 * g)
    assert ecob(1, -1) == 0.0 and ecob(0, 1) == 0.0
    # yukawa depends on particle 1 (lower index) only, model.py:262-266
    assert ecob(1, -2, form=1) == pytest.approx(-1.0 * np.exp(-r / rc) / r)
    assert ecob(-2, 1, form=1) == pytest.approx(-2.0 * np.exp(-r / rc) / r)
    assert ecob(0, 1, form=1) == 0.0
    # theta: -E step(rc - r)
    assert ecob(1, 1, form=2) == 0.0
    sysd = O.System(n=2, cob=(2, [rc, 1.0, 2.0]), s=np.array([-1, -1], dtype=np.int8))
    assert O.energy_forces(sysd, two(0.1))[0][1] == -2.0
    assert O.energy_forces(sysd, two(rc))[0][1] == -2.0  # step(0) = 1 [OpenMM]

    def escb(s1, s2):
        sysd = O.System(n=2, scb=(0, [rc, 1.0, 1.33, 1.66, 2.0]), s=np.array([s1, s2], dtype=np.int8))
        return O.energy_forces(sysd, two(r))[0][2]
    # model.py:322-328: Ea1 <-> s=2, Ea2 <-> s=1, Eb1 <-> s=-1, Eb2 <-> s=-2, equal labels only
    assert escb(2, 2) == pytest.approx(-1.0 * g)
    assert escb(1, 1) == pytest.approx(-1.33 * g)
    assert escb(-1, -1) == pytest.approx(-1.66 * g)
    assert escb(-2, -2) == pytest.approx(-2.0 * g)
    assert escb(1, 2) == 0.0 and escb(0, 0) == 0.0


def test_kat_chromosomal_blocks():
    r = 0.7
    def echb(c1, c2, form=0):
        sysd = O.System(n=2, chb=(form, [0.3, 1e-4]), chrom=np.array([c1, c2], dtype=np.int32))
        return O.energy_forces(sysd, two(r))[0][3]
    assert echb(3, 3) == pytest.approx(1e-4 * (0.3 * r ** 4 - r ** 3 + r ** 2))  # model.py:416-419
    assert echb(3, 4) == 0.0
    assert echb(1, 1, 1) == pytest.approx(-1e-4 * np.exp(-0.3 * r * r))        # model.py:428-431
    assert echb(1, 1, 2) == pytest.approx(-1e-4 / (1 + 0.3 * r * r))            # model.py:440-443


def test_kat_external_terms():
    r1, r2 = 1.0, 2.0
    x = np.array([[2.5, 0, 0], [0.4, 0, 0.0]])
    sysd = O.System(n=2, sc=(0, [1000.0, r1, r2, 0, 0, 0]))
    e, f = O.energy_forces(sysd, x)
    # model.py:454-456: outer wall on bead 0, inner wall on bead 1
    assert e[4] == pytest.approx(1000.0 * (0.5 ** 2 + 0.6 ** 2))
    assert f[0, 0] == pytest.approx(-2 * 1000.0 * 0.5) and f[1, 0] == pytest.approx(2 * 1000.0 * 0.6)
    # lamina sin^8, only B beads (s in {-1,-2}), model.py:503-505
    rho = 1.25
    xs = np.array([[rho, 0, 0], [rho, 0, 0.0]]) + np.array([[0, 0, 0], [0, 3.0, 0]])
    sysd = O.System(n=2, lam=(0, [400.0, r1, r2, 0, 0, 0]), s=np.array([-1, 1], dtype=np.int8))
    e, _ = O.energy_forces(sysd, xs)
    assert e[5] == pytest.approx(400.0 * (np.sin(np.pi * 0.25) ** 8 - 1.0))
    # central force, model.py:584-586
    sysd = O.System(n=2, cf=(0, [20.0, r1, 0, 0, 0]), cstr=np.array([0.5, 0.0]))
    e, _ = O.energy_forces(sysd, x)
    assert e[6] == pytest.approx(20.0 * 0.5 * 1.5 ** 2)


def test_radii():
    # set_radiuses, model.py:1016-1067; SURVEY row R quotes S1 and S3
    r1, r2, rc = radii(10 ** 4)
    assert (round(r2, 3), round(r1, 3), rc) == (2.154, 1.260, pytest.approx(0.15))
    r1, r2, _ = radii(2 * 10 ** 5)
    assert (round(r2, 3), round(r1, 3)) == (5.848, 3.420)


# ---------------------------------------------------------------------------------------------
# finite differences: every term, every functional form
# ---------------------------------------------------------------------------------------------
ALL_FORMS = [
    {}, {"EV": 1}, {"COB": 1, "SCB": 1}, {"COB": 2, "SCB": 2}, {"CHB": 1}, {"CHB": 2},
    {"LAM": 1, "CF": 1, "LOOP": 1}, {"LAM": 2, "CF": 2, "LOOP": 2}, {"LAM": 3},
]


@pytest.mark.parametrize("forms", ALL_FORMS)
def test_forces_are_minus_gradient(forms):
    case = make_case(48, n_chrom=3, seed=5, forms=forms, chb_de=0.5, noise=0.02,
                     terms=("EV", "COB", "SCB", "CHB", "SC", "LAM", "CF", "BOND", "LOOP", "ANGLE"))
    sysd = to_oracle(case)
    x = case["x"]
    terms_with_force = [t for t in range(10) if not (forms.get("COB") == 2 and t in (1, 2))]
    _, f = O.energy_forces(sysd, x)
    h = 1e-6
    rng = np.random.default_rng(0)
    # directional derivatives along random directions, per term subset (theta has zero force)
    for _ in range(6):
        d = rng.normal(size=x.shape)
        d /= np.linalg.norm(d)
        ep = O.energy_forces(sysd, x + h * d, want_forces=False)[0][terms_with_force].sum()
        em = O.energy_forces(sysd, x - h * d, want_forces=False)[0][terms_with_force].sum()
        fd = -(ep - em) / (2 * h)
        an = float((f * d).sum())
        assert fd == pytest.approx(an, rel=2e-6, abs=1e-6 * np.abs(f).max())


def test_cutoff_truncates_but_keeps_chb():
    case = make_case(300, n_chrom=2, seed=2, terms=("EV", "SCB", "CHB"))
    x = case["x"]
    e_all, _ = O.energy_forces(to_oracle(case), x)
    e_cut, _ = O.energy_forces(to_oracle(case, cutoff=0.35), x)
    assert e_cut[0] < e_all[0] and e_cut[3] == pytest.approx(e_all[3], rel=1e-14)
    npairs = O.count_pairs(to_oracle(case, cutoff=0.35), x)
    d = np.linalg.norm(x[:, None] - x[None], axis=-1)
    assert abs(npairs - int((np.triu(d < 0.35, 1)).sum())) <= 2  # FP32 vs FP64 boundary pairs


def test_threads_agree():
    case = make_case(500, n_chrom=3, seed=8)
    e1, f1 = O.energy_forces(to_oracle(case), case["x"], nthreads=1)
    e4, f4 = O.energy_forces(to_oracle(case), case["x"], nthreads=4)
    assert np.allclose(e1, e4, rtol=1e-12) and np.allclose(f1, f4, rtol=1e-9, atol=1e-9)


# ---------------------------------------------------------------------------------------------
# topology quirks (model.py:628-635, 711-719)
# ---------------------------------------------------------------------------------------------
def test_backbone_single_chromosome():
    n = 50
    bi = O.backbone_bonds(n, [0, n])
    ai = O.backbone_angles(n, [0, n])
    assert bi[0] == 1 and len(bi) == n - 2          # bead 0 is unbonded: i = 0 is in chr_ends
    assert ai[0] == 1 and len(ai) == n - 3          # i = 0 skipped (in chr_ends); N-1 is not < N-2
    b2, a2 = backbone(n, [0, n])
    assert np.array_equal(bi, b2) and np.array_equal(ai, a2)


def test_backbone_genome_wide_counts():
    n, ends = 1000, np.array([0, 100, 250, 600, 1000])
    bi = O.backbone_bonds(n, ends)
    ai = O.backbone_angles(n, ends)
    c = len(ends) - 1
    assert len(bi) == n - 1 - c                       # N - 1 - C
    assert len(ai) == (n - 2) - c - (c - 1)
    assert 99 in bi and 100 not in bi                 # (e_k - 1, e_k) joins consecutive chromosomes
    assert 98 in ai and 99 not in ai and 100 not in ai
    b2, a2 = backbone(n, ends)
    assert np.array_equal(bi, b2) and np.array_equal(ai, a2)


# ---------------------------------------------------------------------------------------------
# Hilbert curve (hilbertcurve 2.0.5 restatement): invariants + first points
# ---------------------------------------------------------------------------------------------
def test_hilbert_first_points_and_unit_steps():
    pts = O.hilbert_points(2 ** 15, 8)
    assert pts[:4].tolist() == [[0, 0, 0], [0, 1, 0], [1, 1, 0], [1, 0, 0]]
    steps = np.abs(np.diff(pts, axis=0)).sum(axis=1)
    assert np.all(steps == 1)                          # consecutive points are lattice neighbours


@pytest.mark.parametrize("k", [1, 2, 3, 4, 6])
def test_hilbert_fills_cubes(k):
    # the first 2^(3k) points of the order-8 curve are a bijection onto a (2^k)^3 cube at the origin
    m = 2 ** (3 * k)
    pts = O.hilbert_points(m, 8)
    assert pts.min() == 0 and pts.max() == 2 ** k - 1
    assert len({tuple(p) for p in pts.tolist()}) == m


def test_hilbert_order_consistency():
    # the "undo excess work" loop runs p-1 times and permutes the axes cyclically each time,
    # so curves whose orders differ by a multiple of 3 share their first points
    a = O.hilbert_points(512, 8)
    assert np.array_equal(a, O.hilbert_points(512, 5))
    assert np.array_equal(a[:, [1, 2, 0]], O.hilbert_points(512, 6))


# ---------------------------------------------------------------------------------------------
# L-BFGS restatement
# ---------------------------------------------------------------------------------------------
def test_lbfgs_converges_and_meets_openmm_rule():
    case = make_case(120, n_chrom=2, seed=3, noise=0.0)
    sysd = to_oracle(case)
    x1, rep = O.minimize(sysd, case["x"], tol=10.0)
    assert rep["converged"] == 1 and rep["e_final"] < rep["e_initial"]
    e, f = O.energy_forces(sysd, x1)
    assert e.sum() == pytest.approx(rep["e_final"], rel=1e-12)
    # per-particle RMS force below the tolerance (the meaning of OpenMM's epsilon scaling)
    assert np.sqrt((f ** 2).sum() / case["n"]) <= 10.0
    assert rep["rms_force"] == pytest.approx(np.sqrt((f ** 2).sum() / case["n"]), rel=1e-9)


def test_lbfgs_max_iterations():
    case = make_case(120, n_chrom=1, seed=4, noise=0.0)
    _, rep = O.minimize(to_oracle(case), case["x"], tol=1e-6, max_iter=5)
    assert rep["iterations"] == 5 and rep["converged"] == 0


def test_lbfgs_quadratic_exact():
    # two beads on a spring: minimum at r = r0, energy 0
    sysd = O.System(n=2, bonds=([0], [1], [0.1], [3.0e5]))
    x1, rep = O.minimize(sysd, two(0.13), tol=1e-3)
    assert rep["converged"] == 1
    assert np.linalg.norm(x1[0] - x1[1]) == pytest.approx(0.1, abs=1e-6)


# ---------------------------------------------------------------------------------------------
# cell list
# ---------------------------------------------------------------------------------------------
def test_cell_list_sorted_and_stable():
    rng = np.random.default_rng(1)
    x = rng.uniform(-2, 2, size=(2000, 3)).astype(np.float32)
    keys, order = O.cell_list(x, 0.5, 8, -2.0)
    assert np.all(np.diff(keys.astype(np.int64)) >= 0)
    assert sorted(order.tolist()) == list(range(2000))
    same = keys[1:] == keys[:-1]
    assert np.all(order[1:][same] > order[:-1][same])  # stable within a cell
    # the key of each bead is the Morton code of its cell
    c = np.clip(np.floor((x[order] + 2.0) / 0.5).astype(int), 0, 7)
    def spread(v):
        out = np.zeros_like(v)
        for b in range(10):
            out |= ((v >> b) & 1) << (3 * b)
        return out
    assert np.array_equal(keys, (spread(c[:, 0]) | (spread(c[:, 1]) << 1) | (spread(c[:, 2]) << 2)).astype(np.uint32))


# ---------------------------------------------------------------------------------------------
# Hilbert curve: the same algorithm in n dimensions (tests only)
# ---------------------------------------------------------------------------------------------
def _skilling_points(n_points, p, n):
    """Skilling's transpose -> axes decode as hilbertcurve 2.x applies it, for any dimension n: the
    distance is written as an n*p-bit string, axis a takes bits a, a+n, ...; Gray decode; undo the
    excess work."""
    out = []
    for h in range(n_points):
        bits = format(h, f"0{n * p}b")
        x = [int(bits[a::n], 2) for a in range(n)]
        z = 2 << (p - 1)
        t = x[n - 1] >> 1
        for i in range(n - 1, 0, -1):
            x[i] ^= x[i - 1]
        x[0] ^= t
        q = 2
        while q != z:
            pm = q - 1
            for i in range(n - 1, -1, -1):
                if x[i] & q:
                    x[0] ^= pm
                else:
                    t = (x[0] ^ x[i]) & pm
                    x[0] ^= t
                    x[i] ^= t
            q <<= 1
        out.append(x)
    return out


def test_hilbert_algorithm_reproduces_the_packages_documented_example():
    """hilbertcurve's README example: HilbertCurve(p=1, n=2).points_from_distances(range(4)) is
    [0,0], [0,1], [1,1], [1,0].  The package is not installable here; the n-dimensional form of the
    restated algorithm reproduces that example, and its n = 3 case IS the oracle's generator."""
    assert _skilling_points(4, 1, 2) == [[0, 0], [0, 1], [1, 1], [1, 0]]
    for p in (1, 2, 3, 8):
        m = min(512, 8 ** p)
        assert np.array_equal(np.array(_skilling_points(m, p, 3)), O.hilbert_points(m, p))
    # every order-p curve in 2-D is a Hamiltonian path of unit steps on the 2^p x 2^p grid
    pts = np.array(_skilling_points(64, 3, 2))
    assert len({tuple(q) for q in pts}) == 64 and (np.abs(np.diff(pts, axis=0)).sum(axis=1) == 1).all()
