/*
 * mmm_oracle.h — CPU oracle for the MultiMM hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this; the product (multimm_b200/) never does.
 *
 * PARITY — what pins this oracle, and what does not.
 * Pinned to the reference's own code, executed in the build container:
 *   - the force field: tests/golden/make_golden_forcefield.py runs the UNMODIFIED add_* methods of
 *     src/multimm/model.py against a recording stand-in for `openmm` and evaluates the Lepton
 *     strings, parameters and bond / angle / loop lists they produce; tests/test_forcefield_golden.py
 *     holds every per-term energy of this oracle to those values (1e-10), for every functional form;
 *   - the input loaders, the start curves, the .cif / .psf writers and the structure report
 *     (tests/golden/make_golden*.py).
 * NOT pinned (restated from public documentation, marked [OpenMM] in the sources): OpenMM 8.5.1's
 * own evaluation of those expressions and its LocalEnergyMinimizer / liblbfgs (uv.lock:2462-2463),
 * and hilbertcurve 2.0.5 (uv.lock:1129-1130) — neither is installed or installable here, and the
 * reference's tests hold no numeric vectors for the path (tests/test_simulations.py asserts file
 * existence only).  For those parts: parity unpinned; closed-form known answers, finite
 * differences and curve invariants in tests/test_oracle.py stand in.
 */
#ifndef MMM_ORACLE_H
#define MMM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NUM_TERMS 10

typedef struct {
  int64_t n;
  /* pair terms: form (-1 = off) and globals as in include/multimm_b200.h */
  int32_t ev_form;  double ev[4];   /* epsilon, r_small, sigma, power */
  int32_t cob_form; double cob[3];  /* rc, Ea, Eb */
  int32_t scb_form; double scb[5];  /* rsc, Ea1, Ea2, Eb1, Eb2 */
  int32_t chb_form; double chb[2];  /* kC, dE */
  int32_t sc_form;  double sc[6];   /* C, R1, R2, x0, y0, z0 */
  int32_t lam_form; double lam[6];  /* B, R1, R2, x0, y0, z0 */
  int32_t cf_form;  double cf[5];   /* G, R1, x0, y0, z0 */
  int32_t loop_form;
  double cutoff;                    /* 0 = NoCutoff */
  const int8_t *s;                  /* [n] compartment label, may be NULL */
  const int32_t *chrom;             /* [n] chromosome id, may be NULL */
  const double *cstr;               /* [n] chrom_strength, may be NULL */
  int64_t nb; const int32_t *bi, *bj; const double *br0, *bk;
  int64_t nl; const int32_t *li, *lj; const double *lr0, *lk;
  int64_t na; const int32_t *ai, *aj, *ak; const double *at0, *akt;
} orc_params;

typedef struct {
  int64_t iterations, evaluations;
  double e_initial, e_final, rms_force;
  int32_t converged, ls_status;
} orc_min_report;

/* Per-term energies e[ORC_NUM_TERMS] and forces f[n*3] (may be NULL) at x[n*3]. */
int orc_energy_forces(const orc_params *p, const double *x, double *e, double *f, int nthreads);
/* Number of threads a parallel region that asks for nthreads (<= 0: the OpenMP default) gets. */
int orc_threads_used(int nthreads);
/* Number of unordered pairs within the cutoff (or n(n-1)/2 when cutoff == 0). */
int64_t orc_count_pairs(const orc_params *p, const double *x);
/* liblbfgs restatement with OpenMM's LocalEnergyMinimizer settings; x is updated in place. */
int orc_minimize(const orc_params *p, double *x, double tol, int64_t max_iter, int single_precision,
                 orc_min_report *rep, int nthreads);
/* First n points of the order-p 3-D Hilbert curve (hilbertcurve 2.0.5 algorithm). */
void orc_hilbert_points(int64_t n, int p, int32_t *ijk);
/* Backbone topology rules of model.py:628-635 and 711-719. Return the count written. */
int64_t orc_backbone_bonds(int64_t n, const int64_t *chr_ends, int64_t n_ends, int32_t *bi);
int64_t orc_backbone_angles(int64_t n, const int64_t *chr_ends, int64_t n_ends, int32_t *ai);
/* Cell keys and stable sorted order for cutoff mode (FP32 arithmetic identical to the GPU). */
void orc_cell_list(int64_t n, const float *xyzc /* centred, n*3 */, float cell, int32_t dim,
                   float origin, uint32_t *key_sorted, int32_t *order);

#ifdef __cplusplus
}
#endif
#endif
