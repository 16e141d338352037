/*
 * mmm_oracle.c — CPU oracle for the MultiMM hot path.  TEST INFRASTRUCTURE ONLY; see
 * mmm_oracle.h for who may load it and for what pins it: the energy terms are held to the
 * reference's own force-field builder (tests/test_forcefield_golden.py); OpenMM's evaluation
 * engine, its minimiser and the hilbertcurve package are restated from documentation (parity
 * unpinned there), with known answers and finite differences in tests/test_oracle.py.
 *
 * FP64 throughout, plain loops.  Each function cites the reference lines it restates
 * (paths relative to /root/reference/).  OpenMM conventions restated from its public
 * documentation are marked [OpenMM]:
 *   HarmonicBondForce   E = 1/2 k (r - r0)^2
 *   HarmonicAngleForce  E = 1/2 k (theta - theta0)^2
 *   CustomNonbondedForce, NoCutoff: sum over i < j, particle "1" is the lower index
 *   delta(x) = 1 if x == 0 else 0;  step(x) = 0 if x < 0 else 1;  max; ^
 */
#include "mmm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define T_EV 0
#define T_COB 1
#define T_SCB 2
#define T_CHB 3
#define T_SC 4
#define T_LAM 5
#define T_CF 6
#define T_BOND 7
#define T_LOOP 8
#define T_ANGLE 9

/* ------------------------------------------------------------------------------------ */
/* pair terms                                                                            */
/* ------------------------------------------------------------------------------------ */

/* E(s1,s2) of add_compartment_blocks, src/multimm/model.py:246-250 (gaussian / theta) */
static double cob_strength(const double *g, int s1, int s2) {
  double a1 = (s1 == 1) + (s1 == 2), a2 = (s2 == 1) + (s2 == 2);
  double b1 = (s1 == -1) + (s1 == -2), b2 = (s2 == -1) + (s2 == -2);
  return g[1] * a1 * a2 + g[2] * b1 * b2;
}
/* yukawa variant uses s1 on both factors, model.py:262-266 (particle 1 = lower index) */
static double cob_strength_yukawa(const double *g, int s1) {
  double a1 = (s1 == 1) + (s1 == 2), b1 = (s1 == -1) + (s1 == -2);
  return g[1] * a1 * a1 + g[2] * b1 * b1;
}
/* E(s1,s2) of add_subcompartment_blocks, model.py:322-328 */
static double scb_strength(const double *g, int s1, int s2) {
  return g[1] * ((s1 == 2) && (s2 == 2)) + g[2] * ((s1 == 1) && (s2 == 1)) +
         g[3] * ((s1 == -1) && (s2 == -1)) + g[4] * ((s1 == -2) && (s2 == -2));
}

/* Energy of one pair per term (e4[0..3]) and dE/dr per term summed (returned through *dedr).
 * in_cut: whether the truncatable terms (EV, COB, SCB) are evaluated for this pair. */
static inline void pair_eval(const orc_params *p, int s1, int s2, int c1, int c2, double r,
                             int in_cut, double e4[4], double *dedr) {
  double de = 0.0;
  e4[0] = e4[1] = e4[2] = e4[3] = 0.0;
  if (in_cut) {
    if (p->ev_form == 0) { /* model.py:199  epsilon*(sigma/(r+r_small))^EV_POWER */
      double eps = p->ev[0], rs = p->ev[1], sig = p->ev[2], pw = p->ev[3];
      double e = eps * pow(sig / (r + rs), pw);
      e4[0] = e;
      de += -pw * e / (r + rs);
    } else if (p->ev_form == 1) { /* model.py:209  epsilon*exp(-r^2/(2 sigma^2)) */
      double eps = p->ev[0], sig = p->ev[2];
      double e = eps * exp(-r * r / (2.0 * sig * sig));
      e4[0] = e;
      de += -e * r / (sig * sig);
    }
    if (p->cob_form >= 0) {
      double rc = p->cob[0];
      if (p->cob_form == 0) { /* model.py:246-250 */
        double E = cob_strength(p->cob, s1, s2);
        double g = exp(-r * r / (2.0 * rc * rc));
        e4[1] = -E * g;
        de += E * g * r / (rc * rc);
      } else if (p->cob_form == 1) { /* model.py:262-266, lambda = r_comp (model.py:268) */
        double E = cob_strength_yukawa(p->cob, s1);
        double g = exp(-r / rc);
        e4[1] = -E * g / r;
        de += E * g * (1.0 / (rc * r) + 1.0 / (r * r));
      } else { /* model.py:279-283: -E*step(rc-r), zero force */
        double E = cob_strength(p->cob, s1, s2);
        e4[1] = -E * ((rc - r) >= 0.0 ? 1.0 : 0.0);
      }
    }
    if (p->scb_form >= 0) {
      double rc = p->scb[0];
      double E = scb_strength(p->scb, s1, s2);
      if (p->scb_form == 0) { /* model.py:322-328 */
        double g = exp(-r * r / (2.0 * rc * rc));
        e4[2] = -E * g;
        de += E * g * r / (rc * rc);
      } else if (p->scb_form == 1) { /* model.py:342-348 */
        double g = exp(-r / rc);
        e4[2] = -E * g / r;
        de += E * g * (1.0 / (rc * r) + 1.0 / (r * r));
      } else { /* model.py:363-369 */
        e4[2] = -E * ((rc - r) >= 0.0 ? 1.0 : 0.0);
      }
    }
  }
  /* CHB is never truncated: the default polynomial grows with r. */
  if (p->chb_form >= 0 && c1 == c2) {
    double kc = p->chb[0], dE = p->chb[1];
    if (p->chb_form == 0) { /* model.py:416-419 */
      double r2 = r * r;
      e4[3] = dE * (kc * r2 * r2 - r2 * r + r2);
      de += dE * (4.0 * kc * r2 * r - 3.0 * r2 + 2.0 * r);
    } else if (p->chb_form == 1) { /* model.py:428-431 */
      double g = exp(-kc * r * r);
      e4[3] = -dE * g;
      de += 2.0 * kc * r * dE * g;
    } else { /* model.py:440-443 */
      double q = 1.0 / (1.0 + kc * r * r);
      e4[3] = -dE * q;
      de += dE * q * q * 2.0 * kc * r;
    }
  }
  *dedr = de;
}

/* FP32 inclusion test shared bit-for-bit with the GPU cell-list path: centred float
 * coordinates, r2 = fma(dz,dz,fma(dy,dy,dx*dx)), pair kept iff r2 < rc^2 (float). */
static inline int pair_in_cut(const float *xc, int64_t i, int64_t j, float rc2) {
  float dx = xc[3 * i] - xc[3 * j], dy = xc[3 * i + 1] - xc[3 * j + 1],
        dz = xc[3 * i + 2] - xc[3 * j + 2];
  float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
  return r2 < rc2;
}

static float *centred_floats(const orc_params *p, const double *x) {
  /* centre = arithmetic mean of the positions (the engine uses the same definition) */
  int64_t n = p->n;
  double c[3] = {0, 0, 0};
  for (int64_t i = 0; i < n; i++)
    for (int d = 0; d < 3; d++) c[d] += x[3 * i + d];
  for (int d = 0; d < 3; d++) c[d] /= (double)n;
  float *xc = (float *)malloc(sizeof(float) * 3 * (size_t)n);
  for (int64_t i = 0; i < n; i++)
    for (int d = 0; d < 3; d++) xc[3 * i + d] = (float)(x[3 * i + d] - c[d]);
  return xc;
}

int64_t orc_count_pairs(const orc_params *p, const double *x) {
  int64_t n = p->n;
  if (p->cutoff <= 0.0) return n * (n - 1) / 2;
  float *xc = centred_floats(p, x);
  float rc2 = (float)p->cutoff * (float)p->cutoff;
  int64_t cnt = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : cnt)
  for (int64_t i = 0; i < n; i++)
    for (int64_t j = i + 1; j < n; j++) cnt += pair_in_cut(xc, i, j, rc2);
  free(xc);
  return cnt;
}

/* ------------------------------------------------------------------------------------ */
/* external terms                                                                        */
/* ------------------------------------------------------------------------------------ */

/* returns energy; *dedrho = dE/d rho */
static double sc_eval(const double *g, double rho, double *dedrho) {
  /* model.py:454-456  C*(max(0,r-R2)^2 + max(0,R1-r)^2) */
  double C = g[0], R1 = g[1], R2 = g[2];
  double a = fmax(0.0, rho - R2), b = fmax(0.0, R1 - rho);
  *dedrho = 2.0 * C * (a - b);
  return C * (a * a + b * b);
}

static double lam_eval(int form, const double *g, int s, double rho, double *dedrho) {
  double B = g[0], R1 = g[1], R2 = g[2];
  double on = (s == -1) + (s == -2); /* (delta(s+1)+delta(s+2)) */
  *dedrho = 0.0;
  if (on == 0.0) return 0.0;
  if (form == 0) { /* model.py:503-505  B*(sin(pi*(r-R1)/(R2-R1))^8 - 1) */
    double w = M_PI / (R2 - R1), a = w * (rho - R1);
    double sn = sin(a), cs = cos(a);
    double s2 = sn * sn, s4 = s2 * s2, s7 = s4 * s2 * sn;
    *dedrho = B * 8.0 * s7 * cs * w;
    return B * (s4 * s4 - 1.0);
  } else if (form == 1) { /* model.py:513-517 */
    double sg = 0.1 * (R2 - R1), q = 2.0 * sg * sg;
    double g1 = exp(-(rho - R1) * (rho - R1) / q), g2 = exp(-(rho - R2) * (rho - R2) / q);
    *dedrho = B * (g1 * 2.0 * (rho - R1) / q + g2 * 2.0 * (rho - R2) / q);
    return -B * (g1 + g2);
  } else if (form == 2) { /* model.py:524-527 */
    double r0 = 0.5 * (R1 + R2);
    *dedrho = 2.0 * B * (rho - r0);
    return B * (rho - r0) * (rho - r0);
  } else { /* model.py:534-538 */
    double lm = 0.05 * (R2 - R1);
    double ea = exp((rho - R2) / lm), eb = exp(-(rho - R1) / lm);
    double fa = 1.0 / (1.0 + ea), fb = 1.0 / (1.0 + eb);
    /* d/drho 1/(1+ea) = -ea/(lm (1+ea)^2); d/drho 1/(1+eb) = +eb/(lm (1+eb)^2) */
    *dedrho = -B * (-ea * fa * fa / lm + eb * fb * fb / lm);
    return -B * (fa + fb);
  }
}

static double cf_eval(int form, const double *g, double c, double rho, double *dedrho) {
  double G = g[0], R1 = g[1];
  if (form == 0) { /* model.py:584-586  G*chrom_s*(r-R1)^2 */
    *dedrho = 2.0 * G * c * (rho - R1);
    return G * c * (rho - R1) * (rho - R1);
  } else if (form == 1) { /* model.py:594-599  -G*chrom_s*exp(-r^2/(2 sigma^2)), sigma=0.5 R1 */
    double sg = 0.5 * R1, e = exp(-rho * rho / (2.0 * sg * sg));
    *dedrho = G * c * e * rho / (sg * sg);
    return -G * c * e;
  } else { /* model.py:607-612  -G*chrom_s/(1+exp((r-R1)/lambda)), lambda=0.2 R1 */
    double lm = 0.2 * R1, ea = exp((rho - R1) / lm), q = 1.0 / (1.0 + ea);
    *dedrho = G * c * ea * q * q / lm;
    return -G * c * q;
  }
}

/* ------------------------------------------------------------------------------------ */
/* bonded terms                                                                          */
/* ------------------------------------------------------------------------------------ */

static double bond_like(int form, double r, double r0, double k, double *dedr) {
  double d = r - r0;
  if (form == 0) { /* [OpenMM] HarmonicBondForce; model.py:630-635, 653-659 */
    *dedr = k * d;
    return 0.5 * k * d * d;
  } else if (form == 1) { /* model.py:664-680: k*(r-r0)^2/(1+alpha*(r-r0)^2), alpha=1/r0^2 */
    double al = 1.0 / (r0 * r0), q = 1.0 + al * d * d;
    *dedr = 2.0 * k * d / (q * q);
    return k * d * d / q;
  } else { /* model.py:685-701: k*(1-exp(-(r-r0)^2/sigma^2)), sigma=r0/2 */
    double sg = 0.5 * r0, g = exp(-d * d / (sg * sg));
    *dedr = k * g * 2.0 * d / (sg * sg);
    return k * (1.0 - g);
  }
}

static double eval_bonds(int form, int64_t nb, const int32_t *bi, const int32_t *bj,
                         const double *r0, const double *k, const double *x, double *f) {
  double e = 0.0;
  for (int64_t b = 0; b < nb; b++) {
    int64_t i = bi[b], j = bj[b];
    double dx = x[3 * i] - x[3 * j], dy = x[3 * i + 1] - x[3 * j + 1],
           dz = x[3 * i + 2] - x[3 * j + 2];
    double r = sqrt(dx * dx + dy * dy + dz * dz), dedr;
    e += bond_like(form, r, r0[b], k[b], &dedr);
    if (f && r > 0.0) {
      double s = -dedr / r;
      f[3 * i] += s * dx; f[3 * i + 1] += s * dy; f[3 * i + 2] += s * dz;
      f[3 * j] -= s * dx; f[3 * j + 1] -= s * dy; f[3 * j + 2] -= s * dz;
    }
  }
  return e;
}

/* [OpenMM] HarmonicAngleForce on the Reference platform: theta = acos(clamped cosine),
 * E = 1/2 k (theta-theta0)^2; forces through the cross-product form with |u x v| floored at
 * 1e-6.  model.py:711-719 adds (i, i+1, i+2) with the middle bead as the vertex. */
static double eval_angles(int64_t na, const int32_t *ai, const int32_t *aj, const int32_t *ak,
                          const double *t0, const double *kt, const double *x, double *f) {
  double e = 0.0;
  for (int64_t a = 0; a < na; a++) {
    int64_t i = ai[a], j = aj[a], k = ak[a];
    double u[3], v[3];
    for (int d = 0; d < 3; d++) {
      u[d] = x[3 * i + d] - x[3 * j + d];
      v[d] = x[3 * k + d] - x[3 * j + d];
    }
    double uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
    double vv = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    double uv = u[0] * v[0] + u[1] * v[1] + u[2] * v[2];
    double pv[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2],
                    u[0] * v[1] - u[1] * v[0]};
    double rp = sqrt(pv[0] * pv[0] + pv[1] * pv[1] + pv[2] * pv[2]);
    if (rp < 1e-6) rp = 1e-6;
    double cs = uv / sqrt(uu * vv), th;
    if (cs >= 1.0) th = 0.0;
    else if (cs <= -1.0) th = M_PI;
    else th = acos(cs);
    double dth = th - t0[a];
    e += 0.5 * kt[a] * dth * dth;
    if (f) {
      double dedth = kt[a] * dth;
      /* dtheta/dx_i = (u x p)/(|u|^2 |p|),  dtheta/dx_k = -(v x p)/(|v|^2 |p|) */
      double ta = -dedth / (uu * rp), tc = dedth / (vv * rp);
      double fa[3] = {ta * (u[1] * pv[2] - u[2] * pv[1]), ta * (u[2] * pv[0] - u[0] * pv[2]),
                      ta * (u[0] * pv[1] - u[1] * pv[0])};
      double fc[3] = {tc * (v[1] * pv[2] - v[2] * pv[1]), tc * (v[2] * pv[0] - v[0] * pv[2]),
                      tc * (v[0] * pv[1] - v[1] * pv[0])};
      for (int d = 0; d < 3; d++) {
        f[3 * i + d] += fa[d];
        f[3 * k + d] += fc[d];
        f[3 * j + d] -= fa[d] + fc[d];
      }
    }
  }
  return e;
}

/* ------------------------------------------------------------------------------------ */
/* full evaluation                                                                       */
/* ------------------------------------------------------------------------------------ */

/* Threads a parallel region asking for `nthreads` really gets (what bench.py reports as "cores":
 * launchers such as torchrun export OMP_NUM_THREADS=1, which silently serialises nthreads <= 0). */
int orc_threads_used(int nthreads) {
  int got = 1;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
  {
#pragma omp single
    got = omp_get_num_threads();
  }
#else
  (void)nthreads;
#endif
  return got;
}

int orc_energy_forces(const orc_params *p, const double *x, double *e, double *f, int nthreads) {
  const int64_t n = p->n;
  for (int t = 0; t < ORC_NUM_TERMS; t++) e[t] = 0.0;
  if (f) memset(f, 0, sizeof(double) * 3 * (size_t)n);
  int any_pair = p->ev_form >= 0 || p->cob_form >= 0 || p->scb_form >= 0 || p->chb_form >= 0;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif

  if (any_pair && n > 1) {
    float *xc = NULL;
    float rc2 = 0.f;
    if (p->cutoff > 0.0) {
      xc = centred_floats(p, x);
      rc2 = (float)p->cutoff * (float)p->cutoff;
    }
    double *fl_all = f ? (double *)calloc((size_t)nthreads * 3 * (size_t)n, sizeof(double)) : NULL;
    double *et_all = (double *)calloc((size_t)nthreads * 4, sizeof(double));
    const int chb_on = p->chb_form >= 0;
#pragma omp parallel num_threads(nthreads)
    {
#ifdef _OPENMP
      int tid = omp_get_thread_num();
#else
      int tid = 0;
#endif
      double *fl = fl_all ? fl_all + (size_t)tid * 3 * (size_t)n : NULL;
      double et[4] = {0, 0, 0, 0};
      /* [OpenMM] NoCutoff: every i < j, no exclusions (no addExclusion anywhere in model.py) */
#pragma omp for schedule(dynamic, 16)
      for (int64_t i = 0; i < n; i++) {
        const double xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
        const int si = p->s ? p->s[i] : 0, ci = p->chrom ? p->chrom[i] : 0;
        double fxi = 0, fyi = 0, fzi = 0;
        for (int64_t j = i + 1; j < n; j++) {
          int in_cut = xc ? pair_in_cut(xc, i, j, rc2) : 1;
          const int cj = p->chrom ? p->chrom[j] : 0;
          if (!in_cut && !(chb_on && ci == cj)) continue;
          double dx = xi - x[3 * j], dy = yi - x[3 * j + 1], dz = zi - x[3 * j + 2];
          double r = sqrt(dx * dx + dy * dy + dz * dz);
          double e4[4], dedr;
          pair_eval(p, si, p->s ? p->s[j] : 0, ci, cj, r, in_cut, e4, &dedr);
          et[0] += e4[0]; et[1] += e4[1]; et[2] += e4[2]; et[3] += e4[3];
          if (fl) {
            double sc = -dedr / r;
            fxi += sc * dx; fyi += sc * dy; fzi += sc * dz;
            fl[3 * j] -= sc * dx; fl[3 * j + 1] -= sc * dy; fl[3 * j + 2] -= sc * dz;
          }
        }
        if (fl) { fl[3 * i] += fxi; fl[3 * i + 1] += fyi; fl[3 * i + 2] += fzi; }
      }
      for (int t = 0; t < 4; t++) et_all[tid * 4 + t] = et[t];
    }
    for (int t = 0; t < nthreads; t++)
      for (int q = 0; q < 4; q++) e[q] += et_all[t * 4 + q];
    if (f) {
#pragma omp parallel for num_threads(nthreads)
      for (int64_t q = 0; q < 3 * n; q++) {
        double a = 0;
        for (int t = 0; t < nthreads; t++) a += fl_all[(size_t)t * 3 * (size_t)n + q];
        f[q] += a;
      }
    }
    free(fl_all);
    free(et_all);
    free(xc);
  }

  /* external terms, one bead at a time (CustomExternalForce) */
  for (int64_t i = 0; i < n; i++) {
    const int si = p->s ? p->s[i] : 0;
    if (p->sc_form >= 0) {
      double dx = x[3 * i] - p->sc[3], dy = x[3 * i + 1] - p->sc[4], dz = x[3 * i + 2] - p->sc[5];
      double rho = sqrt(dx * dx + dy * dy + dz * dz), d;
      e[T_SC] += sc_eval(p->sc, rho, &d);
      if (f && rho > 0) { f[3 * i] -= d * dx / rho; f[3 * i + 1] -= d * dy / rho; f[3 * i + 2] -= d * dz / rho; }
    }
    if (p->lam_form >= 0) {
      double dx = x[3 * i] - p->lam[3], dy = x[3 * i + 1] - p->lam[4], dz = x[3 * i + 2] - p->lam[5];
      double rho = sqrt(dx * dx + dy * dy + dz * dz), d;
      e[T_LAM] += lam_eval(p->lam_form, p->lam, si, rho, &d);
      if (f && rho > 0) { f[3 * i] -= d * dx / rho; f[3 * i + 1] -= d * dy / rho; f[3 * i + 2] -= d * dz / rho; }
    }
    if (p->cf_form >= 0) {
      double dx = x[3 * i] - p->cf[2], dy = x[3 * i + 1] - p->cf[3], dz = x[3 * i + 2] - p->cf[4];
      double rho = sqrt(dx * dx + dy * dy + dz * dz), d;
      e[T_CF] += cf_eval(p->cf_form, p->cf, p->cstr ? p->cstr[i] : 0.0, rho, &d);
      if (f && rho > 0) { f[3 * i] -= d * dx / rho; f[3 * i + 1] -= d * dy / rho; f[3 * i + 2] -= d * dz / rho; }
    }
  }
  if (p->nb > 0) e[T_BOND] = eval_bonds(0, p->nb, p->bi, p->bj, p->br0, p->bk, x, f);
  if (p->nl > 0) e[T_LOOP] = eval_bonds(p->loop_form, p->nl, p->li, p->lj, p->lr0, p->lk, x, f);
  if (p->na > 0) e[T_ANGLE] = eval_angles(p->na, p->ai, p->aj, p->ak, p->at0, p->akt, x, f);
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* L-BFGS: restatement of liblbfgs (bundled with OpenMM) as configured by                */
/* [OpenMM] LocalEnergyMinimizer::minimize, reached from model.py:886                    */
/*   m = 6, linesearch = BACKTRACKING_STRONG_WOLFE, ftol 1e-4, wolfe 0.9,                */
/*   max_linesearch 40, min_step 1e-20, max_step 1e20,                                   */
/*   epsilon = tol / max(1, sqrt(sum|x_i|^2 / N)), stop when |g|/max(1,|x|) <= epsilon   */
/* ------------------------------------------------------------------------------------ */

static double vdot(const double *a, const double *b, int64_t n) {
  double s = 0;
  for (int64_t i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}

static double total_energy_grad(const orc_params *p, const double *x, double *g, double *f,
                                int nthreads) {
  double e[ORC_NUM_TERMS], tot = 0;
  orc_energy_forces(p, x, e, f, nthreads);
  for (int t = 0; t < ORC_NUM_TERMS; t++) tot += e[t];
  for (int64_t i = 0; i < 3 * p->n; i++) g[i] = -f[i];
  return tot;
}

int orc_minimize(const orc_params *p, double *x, double tol, int64_t max_iter, int single_precision,
                 orc_min_report *rep, int nthreads) {
  (void)single_precision; /* xtol only matters for the More-Thuente search, unused here */
  const int m = 6;
  const int64_t n3 = 3 * p->n;
  const double ftol = 1e-4, wolfe = 0.9, min_step = 1e-20, max_step = 1e20;
  const int max_ls = 40;
  double *g = malloc(sizeof(double) * n3), *gp = malloc(sizeof(double) * n3);
  double *xp = malloc(sizeof(double) * n3), *d = malloc(sizeof(double) * n3);
  double *f = malloc(sizeof(double) * n3);
  double *S = malloc(sizeof(double) * n3 * m), *Y = malloc(sizeof(double) * n3 * m);
  double ys_[6], alpha[6];
  memset(rep, 0, sizeof(*rep));

  double norm = vdot(x, x, n3) / (double)p->n;
  norm = norm < 1.0 ? 1.0 : sqrt(norm);
  const double epsilon = tol / norm;

  double fx = total_energy_grad(p, x, g, f, nthreads);
  rep->evaluations = 1;
  rep->e_initial = fx;
  for (int64_t i = 0; i < n3; i++) d[i] = -g[i];
  double xnorm = sqrt(vdot(x, x, n3)), gnorm = sqrt(vdot(g, g, n3));
  if (xnorm < 1.0) xnorm = 1.0;
  int64_t k = 1;
  int end = 0, status = 0, converged = 0;
  if (gnorm / xnorm <= epsilon) {
    converged = 1;
  } else {
    double step = 1.0 / sqrt(vdot(d, d, n3));
    for (;;) {
      memcpy(xp, x, sizeof(double) * n3);
      memcpy(gp, g, sizeof(double) * n3);
      /* backtracking line search, strong Wolfe */
      {
        const double dec = 0.5, inc = 2.1;
        double dginit = vdot(g, d, n3);
        if (dginit > 0) { status = -1; break; }
        const double finit = fx, dgtest = ftol * dginit;
        int count = 0;
        for (;;) {
          double width;
          for (int64_t i = 0; i < n3; i++) x[i] = xp[i] + step * d[i];
          fx = total_energy_grad(p, x, g, f, nthreads);
          rep->evaluations++;
          count++;
          if (fx > finit + step * dgtest) {
            width = dec;
          } else {
            double dg = vdot(g, d, n3);
            if (dg < wolfe * dginit) width = inc;
            else if (dg > -wolfe * dginit) width = dec;
            else break;
          }
          if (step < min_step) { status = -2; break; }
          if (step > max_step) { status = -3; break; }
          if (count >= max_ls) { status = -4; break; }
          step *= width;
        }
        if (status < 0) { /* liblbfgs reverts to the previous point and gives up */
          memcpy(x, xp, sizeof(double) * n3);
          memcpy(g, gp, sizeof(double) * n3);
          fx = finit;
          break;
        }
      }
      rep->iterations++;
      xnorm = sqrt(vdot(x, x, n3));
      gnorm = sqrt(vdot(g, g, n3));
      if (xnorm < 1.0) xnorm = 1.0;
      if (gnorm / xnorm <= epsilon) { converged = 1; break; }
      if (max_iter != 0 && max_iter < k + 1) break;
      double *s = S + (size_t)end * n3, *y = Y + (size_t)end * n3;
      for (int64_t i = 0; i < n3; i++) { s[i] = x[i] - xp[i]; y[i] = g[i] - gp[i]; }
      double ys = vdot(y, s, n3), yy = vdot(y, y, n3);
      ys_[end] = ys;
      int bound = (m <= k) ? m : (int)k;
      k++;
      end = (end + 1) % m;
      for (int64_t i = 0; i < n3; i++) d[i] = -g[i];
      int j = end;
      for (int q = 0; q < bound; q++) {
        j = (j + m - 1) % m;
        alpha[j] = vdot(S + (size_t)j * n3, d, n3) / ys_[j];
        const double *yj = Y + (size_t)j * n3;
        for (int64_t i = 0; i < n3; i++) d[i] -= alpha[j] * yj[i];
      }
      for (int64_t i = 0; i < n3; i++) d[i] *= ys / yy;
      for (int q = 0; q < bound; q++) {
        double beta = vdot(Y + (size_t)j * n3, d, n3) / ys_[j];
        const double *sj = S + (size_t)j * n3;
        for (int64_t i = 0; i < n3; i++) d[i] += (alpha[j] - beta) * sj[i];
        j = (j + 1) % m;
      }
      step = 1.0;
    }
  }
  rep->e_final = fx;
  rep->rms_force = sqrt(vdot(g, g, n3) / (double)p->n);
  rep->converged = converged;
  rep->ls_status = status;
  free(g); free(gp); free(xp); free(d); free(f); free(S); free(Y);
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Hilbert curve: hilbertcurve 2.0.5 `point_from_distance` (Skilling's transpose -> axes)  */
/* as called by initial_structure_tools.py:157-161 with p = 8, n = 3.                      */
/* ------------------------------------------------------------------------------------ */
void orc_hilbert_points(int64_t n, int p, int32_t *ijk) {
  for (int64_t h = 0; h < n; h++) {
    uint32_t X[3] = {0, 0, 0};
    /* transpose: bit string of length 3p, MSB first; x[a] takes characters a, a+3, a+6, ... */
    for (int b = 0; b < 3 * p; b++) {
      int bit = (int)((h >> (3 * p - 1 - b)) & 1);
      X[b % 3] = (X[b % 3] << 1) | (uint32_t)bit;
    }
    uint32_t z = 2u << (p - 1);
    uint32_t t = X[2] >> 1;
    for (int i = 2; i > 0; i--) X[i] ^= X[i - 1];
    X[0] ^= t;
    for (uint32_t q = 2; q != z; q <<= 1) {
      uint32_t pm = q - 1;
      for (int i = 2; i >= 0; i--) {
        if (X[i] & q) {
          X[0] ^= pm;
        } else {
          t = (X[0] ^ X[i]) & pm;
          X[0] ^= t;
          X[i] ^= t;
        }
      }
    }
    ijk[3 * h] = (int32_t)X[0];
    ijk[3 * h + 1] = (int32_t)X[1];
    ijk[3 * h + 2] = (int32_t)X[2];
  }
}

/* ------------------------------------------------------------------------------------ */
/* backbone topology                                                                     */
/* ------------------------------------------------------------------------------------ */
static int in_list(int64_t v, const int64_t *a, int64_t na) {
  for (int64_t q = 0; q < na; q++)
    if (a[q] == v) return 1;
  return 0;
}

/* model.py:628-635: bond (i, i+1) for i in [0, N-2] unless i is in chr_ends */
int64_t orc_backbone_bonds(int64_t n, const int64_t *chr_ends, int64_t n_ends, int32_t *bi) {
  int64_t c = 0;
  for (int64_t i = 0; i < n - 1; i++)
    if (!in_list(i, chr_ends, n_ends)) bi[c++] = (int32_t)i;
  return c;
}

/* model.py:711-719: angle (i, i+1, i+2) for i in [0, N-3] unless i in chr_ends or chr_ends-1 */
int64_t orc_backbone_angles(int64_t n, const int64_t *chr_ends, int64_t n_ends, int32_t *ai) {
  int64_t c = 0;
  for (int64_t i = 0; i < n - 2; i++)
    if (!in_list(i, chr_ends, n_ends) && !in_list(i + 1, chr_ends, n_ends)) ai[c++] = (int32_t)i;
  return c;
}

/* ------------------------------------------------------------------------------------ */
/* cell list for cutoff mode: FP32 arithmetic shared bit-for-bit with the GPU             */
/*   cell coordinate c_a = min(dim-1, max(0, (int)floorf((x_a - origin) / cell)))         */
/*   key = 30-bit Morton code of (cx,cy,cz); order = stable sort by key                   */
/* ------------------------------------------------------------------------------------ */
static uint32_t spread3(uint32_t v) {
  v &= 0x3ff;
  v = (v | (v << 16)) & 0x030000FF;
  v = (v | (v << 8)) & 0x0300F00F;
  v = (v | (v << 4)) & 0x030C30C3;
  v = (v | (v << 2)) & 0x09249249;
  return v;
}

typedef struct { uint32_t key; int32_t idx; } kv_t;
static int kv_cmp(const void *a, const void *b) {
  const kv_t *p = (const kv_t *)a, *q = (const kv_t *)b;
  if (p->key != q->key) return p->key < q->key ? -1 : 1;
  return p->idx < q->idx ? -1 : (p->idx > q->idx);
}

void orc_cell_list(int64_t n, const float *xyzc, float cell, int32_t dim, float origin,
                   uint32_t *key_sorted, int32_t *order) {
  kv_t *kv = (kv_t *)malloc(sizeof(kv_t) * (size_t)n);
  for (int64_t i = 0; i < n; i++) {
    uint32_t c[3];
    for (int d = 0; d < 3; d++) {
      int v = (int)floorf((xyzc[3 * i + d] - origin) / cell);
      if (v < 0) v = 0;
      if (v > dim - 1) v = dim - 1;
      c[d] = (uint32_t)v;
    }
    kv[i].key = spread3(c[0]) | (spread3(c[1]) << 1) | (spread3(c[2]) << 2);
    kv[i].idx = (int32_t)i;
  }
  qsort(kv, (size_t)n, sizeof(kv_t), kv_cmp);
  for (int64_t i = 0; i < n; i++) { key_sorted[i] = kv[i].key; order[i] = kv[i].idx; }
  free(kv);
}
