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
// cut2 > 0: plain truncation, the pair counts iff r2 < cut2 with r2 = fma(dz,dz,fma(dy,dy,dx*dx))
// on the centred FP32 coordinates (the oracle's pair_in_cut, bit for bit).  Returns whether the
// pair contributed.
__device__ __forceinline__ bool pair_generic(const float4 pj, const IBead& b, const int si,
                                             const bool i_lower, const PairParams& c, bool live,
                                             float& fx, float& fy, float& fz, float e4[4],
                                             const float cut2 = 0.0f) {
  const float dx = b.x - pj.x, dy = b.y - pj.y, dz = b.z - pj.z;
  float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
  if (cut2 > 0.0f && !(r2 < cut2)) live = false;
  if (!live) r2 = 1.0f;
  const float inv_r = fast_rsqrt(r2);
  const float r = r2 * inv_r;
  const int wj = __float_as_int(pj.w);
  const int sj = (wj & 7) - 2;
  float dedr = 0.0f;  // dE/dr summed over terms
  float e[4] = {0.f, 0.f, 0.f, 0.f};
  if (c.ev_form == MMM_EV_POWERLAW) {
    const float w = fast_rcp(r + c.ev_rs);
    const float en = c.ev_eps * fast_ex2(c.ev_power * fast_lg2(c.ev_sigma * w));
    e[0] = en;
    dedr -= c.ev_power * en * w;
  } else if (c.ev_form == MMM_EV_GAUSSIAN_CORE) {
    const float is2 = 1.0f / (c.ev_sigma * c.ev_sigma);
    const float en = c.ev_eps * fast_ex2(-0.72134752f * r2 * is2);
    e[0] = en;
    dedr -= en * r * is2;
  }
  if (c.cob_form >= 0) {
    const bool ai = si > 0, bi = si < 0, aj = sj > 0, bj = sj < 0;
    float E;
    if (c.cob_form == MMM_BLOCK_YUKAWA) {
      // model.py:262-266 uses s1 on both factors; particle 1 is the lower index [OpenMM]
      const int s1 = i_lower ? si : sj;
      E = s1 > 0 ? c.cob_ea : (s1 < 0 ? c.cob_eb : 0.0f);
    } else {
      E = (ai && aj) ? c.cob_ea : ((bi && bj) ? c.cob_eb : 0.0f);
    }
    if (c.cob_form == MMM_BLOCK_GAUSSIAN) {
      const float irc2 = 1.0f / (c.cob_rc * c.cob_rc);
      const float g = fast_ex2(-0.72134752f * r2 * irc2);
      e[1] = -E * g;
      dedr += E * g * r * irc2;
    } else if (c.cob_form == MMM_BLOCK_YUKAWA) {
      const float il = 1.0f / c.cob_rc;
      const float g = fast_ex2(-1.44269504f * r * il);
      e[1] = -E * g * inv_r;
      dedr += E * g * (il * inv_r + inv_r * inv_r);
    } else {
      e[1] = (c.cob_rc - r >= 0.0f) ? -E : 0.0f;
    }
  }
  if (c.scb_form >= 0) {
    float E = 0.0f;
    if (si == sj && si != 0) E = c.scb_e[si == 2 ? 0 : (si == 1 ? 1 : (si == -1 ? 2 : 3))];
    if (c.scb_form == MMM_BLOCK_GAUSSIAN) {
      const float irc2 = 1.0f / (c.scb_rc * c.scb_rc);
      const float g = fast_ex2(-0.72134752f * r2 * irc2);
      e[2] = -E * g;
      dedr += E * g * r * irc2;
    } else if (c.scb_form == MMM_BLOCK_YUKAWA) {
      const float il = 1.0f / c.scb_rc;
      const float g = fast_ex2(-1.44269504f * r * il);
      e[2] = -E * g * inv_r;
      dedr += E * g * (il * inv_r + inv_r * inv_r);
    } else {
      e[2] = (c.scb_rc - r >= 0.0f) ? -E : 0.0f;
    }
  }
  if (c.chb_form >= 0 && ((b.w ^ wj) & 0xFFFF00) == 0) {
    if (c.chb_form == MMM_CHB_POLYNOMIAL) {
      e[3] = c.chb_de * r2 * fmaf(c.chb_kc, r2, 1.0f - r);
      dedr += c.chb_de * r * fmaf(-3.0f, r, fmaf(4.0f * c.chb_kc, r2, 2.0f));
    } else if (c.chb_form == MMM_CHB_GAUSSIAN) {
      const float g = fast_ex2(-1.44269504f * c.chb_kc * r2);
      e[3] = -c.chb_de * g;
      dedr += 2.0f * c.chb_kc * r * c.chb_de * g;
    } else {
      const float q = fast_rcp(fmaf(c.chb_kc, r2, 1.0f));
      e[3] = -c.chb_de * q;
      dedr += c.chb_de * q * q * 2.0f * c.chb_kc * r;
    }
  }
  if (live) {
    const float fs = -dedr * inv_r;
    fx = fmaf(fs, dx, fx);
    fy = fmaf(fs, dy, fy);
    fz = fmaf(fs, dz, fz);
    e4[0] += e[0]; e4[1] += e[1]; e4[2] += e[2]; e4[3] += e[3];
  }
  return live;
}

}  // namespace pairmath
