// mmm_pairmath.cuh — pair-term arithmetic shared by the gather kernel (mmm_pair.cu) and the
// cut-off cell-list kernel (mmm_cells.cu): MUFU wrappers and the generic any-form pair evaluation.
#pragma once
#include "mmm_internal.cuh"

namespace pairmath {

__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rsqrt(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// per-thread constants of the i-bead
struct IBead {
  float x, y, z;
  int w;
  float a_scb, a_cob;  // eps_i / rc^2 for the SCB / COB Gaussian (0 when the bead's label has none)
};

struct Acc {
  float fx, fy, fz;     // force with all prefactors applied (everything except far EV)
  float ux, uy, uz;     // EV force in units of p * eps * sigma^p (far tiles)
  float eev;            // sum w^p            (x eps sigma^p)
  float gscb, gcob;     // sum of matching Gaussians (x eps_i)
  float echb;           // sum r^2 (kC r^2 - r + 1) over same-chromosome pairs (x dE)
};

// Generic pair: any functional form, runtime switches, no tile skipping.  Used for the
// non-default forms (model.py:205-211, 258-288, 338-378, 424-445) and non-integer EV powers.
// The deltas are taken in FP32 from the centred FP32 coordinates (exactly what the specialised
// kernels see, and the oracle's inclusion test in cut-off mode), everything after that is FP64
// with the accurate libdevice functions: this is the slow path, it buys the north star's
// 1e-5 / 1e-4 bars for every form instead of the 3e-5 / 2e-4 the lg2/ex2 formulation reached.
// cut2 > 0: plain truncation, the pair counts iff r2 < cut2 with r2 = fma(dz,dz,fma(dy,dy,dx*dx))
// in FP32 (the oracle's pair_in_cut, bit for bit).  Returns whether the pair contributed.
__device__ __forceinline__ bool pair_generic(const float4 pj, const IBead& b, const int si,
                                             const bool i_lower, const PairParams& c, bool live,
                                             double& fx, double& fy, double& fz, double e4[4],
                                             const float cut2 = 0.0f) {
  const float dxf = b.x - pj.x, dyf = b.y - pj.y, dzf = b.z - pj.z;
  const float r2f = fmaf(dzf, dzf, fmaf(dyf, dyf, dxf * dxf));
  if (cut2 > 0.0f && !(r2f < cut2)) live = false;
  if (!live) return false;
  const double dx = (double)dxf, dy = (double)dyf, dz = (double)dzf;
  const double r2 = dx * dx + dy * dy + dz * dz;
  const double r = sqrt(r2), inv_r = 1.0 / r;
  const int wj = __float_as_int(pj.w);
  const int sj = (wj & 7) - 2;
  double dedr = 0.0;  // dE/dr summed over terms
  if (c.ev_form == MMM_EV_POWERLAW) {  // model.py:199
    const double w = 1.0 / (r + c.d_ev[1]);
    const double en = c.d_ev[0] * pow(c.d_ev[2] * w, c.d_ev[3]);
    e4[0] += en;
    dedr -= c.d_ev[3] * en * w;
  } else if (c.ev_form == MMM_EV_GAUSSIAN_CORE) {  // model.py:209
    const double is2 = 1.0 / (c.d_ev[2] * c.d_ev[2]);
    const double en = c.d_ev[0] * exp(-0.5 * r2 * is2);
    e4[0] += en;
    dedr -= en * r * is2;
  }
  if (c.cob_form >= 0) {
    const bool ai = si > 0, bi = si < 0, aj = sj > 0, bj = sj < 0;
    double E;
    if (c.cob_form == MMM_BLOCK_YUKAWA) {
      // model.py:262-266 uses s1 on both factors; particle 1 is the lower index [OpenMM]
      const int s1 = i_lower ? si : sj;
      E = s1 > 0 ? c.d_cob[1] : (s1 < 0 ? c.d_cob[2] : 0.0);
    } else {
      E = (ai && aj) ? c.d_cob[1] : ((bi && bj) ? c.d_cob[2] : 0.0);
    }
    if (c.cob_form == MMM_BLOCK_GAUSSIAN) {  // model.py:246-250
      const double irc2 = 1.0 / (c.d_cob[0] * c.d_cob[0]);
      const double g = exp(-0.5 * r2 * irc2);
      e4[1] -= E * g;
      dedr += E * g * r * irc2;
    } else if (c.cob_form == MMM_BLOCK_YUKAWA) {
      const double il = 1.0 / c.d_cob[0];
      const double g = exp(-r * il);
      e4[1] -= E * g * inv_r;
      dedr += E * g * (il * inv_r + inv_r * inv_r);
    } else {  // model.py:279-283: -E step(rc - r), no force
      e4[1] += (c.d_cob[0] - r >= 0.0) ? -E : 0.0;
    }
  }
  if (c.scb_form >= 0) {
    double E = 0.0;
    if (si == sj && si != 0) E = c.d_scb[si == 2 ? 1 : (si == 1 ? 2 : (si == -1 ? 3 : 4))];
    if (c.scb_form == MMM_BLOCK_GAUSSIAN) {  // model.py:322-328
      const double irc2 = 1.0 / (c.d_scb[0] * c.d_scb[0]);
      const double g = exp(-0.5 * r2 * irc2);
      e4[2] -= E * g;
      dedr += E * g * r * irc2;
    } else if (c.scb_form == MMM_BLOCK_YUKAWA) {  // model.py:342-348
      const double il = 1.0 / c.d_scb[0];
      const double g = exp(-r * il);
      e4[2] -= E * g * inv_r;
      dedr += E * g * (il * inv_r + inv_r * inv_r);
    } else {  // model.py:363-369
      e4[2] += (c.d_scb[0] - r >= 0.0) ? -E : 0.0;
    }
  }
  if (c.chb_form >= 0 && ((b.w ^ wj) & 0xFFFF00) == 0) {
    const double kc = c.d_chb[0], de = c.d_chb[1];
    if (c.chb_form == MMM_CHB_POLYNOMIAL) {  // model.py:416-419
      e4[3] += de * r2 * (kc * r2 + 1.0 - r);
      dedr += de * r * (4.0 * kc * r2 - 3.0 * r + 2.0);
    } else if (c.chb_form == MMM_CHB_GAUSSIAN) {  // model.py:428-431
      const double g = exp(-kc * r2);
      e4[3] -= de * g;
      dedr += 2.0 * kc * r * de * g;
    } else {  // model.py:440-443
      const double q = 1.0 / (kc * r2 + 1.0);
      e4[3] -= de * q;
      dedr += de * q * q * 2.0 * kc * r;
    }
  }
  const double fs = -dedr * inv_r;
  fx += fs * dx;
  fy += fs * dy;
  fz += fs * dz;
  return true;
}

}  // namespace pairmath
