// mmm_pair.cu — exact all-pairs ("NoCutoff") pair kernel for sm_100a.
//
// Replaces the four CustomNonbondedForce objects of model.py:164-451 (EV, COB, SCB, CHB), which
// the reference leaves at OpenMM's default NoCutoff method with no exclusions: every i != j
// pair contributes.  One fused pass evaluates all active pair terms.
//
// Formulation (deterministic gather): a work item is (block of 256 consecutive i-beads) x
// (chunk of j-tiles).  Each thread owns one i-bead in registers; j-beads are staged 256 at a
// time in shared memory as float4 {x,y,z,typebits} and read by broadcast LDS.128.  Forces are
// accumulated in FP32 inside a stage and folded into FP64 per thread after each stage, so the
// summation order is fixed and no atomics touch the result.  Every unordered pair is visited
// twice (once from each side); energies are halved at the end.
//
// Tile classification (warp-uniform, from the 32-bead bounding boxes written by k_prepare):
//   far   : box distance^2 >= rg2  -> Gaussian block terms are < 2^-26 of their prefactor and
//           are skipped (their whole tail is < 1e-7 of the term); EV (+CHB) only.
//   chrom : CHB needs work only where the chromosome ranges of the two tiles overlap; if both
//           tiles are single-chromosome the per-pair comparison is dropped as well.
//   diag  : the tile that contains the i-bead itself masks the self pair.
// The roofline that bounds this kernel is the FP32-FMA / MUFU issue rate, not HBM: a stage of
// 4 KB is reused by 256 threads x 256 pairs.
#include "mmm_internal.cuh"
#include "mmm_pairmath.cuh"

namespace {

using namespace pairmath;

enum : int { PM_GAUSS = 1, PM_CHB = 2, PM_CHBMASK = 4, PM_SELF = 8 };

template <int P>
__device__ __forceinline__ float powi(float w) {
  if constexpr (P == 1) {
    return w;
  } else if constexpr (P % 2 == 0) {
    const float hf = powi<P / 2>(w);
    return hf * hf;
  } else {
    return w * powi<P - 1>(w);
  }
}

// One tile of 32 j-beads against this thread's i-bead.  EVP: integer EV power; GK: 0 none,
// 1 SCB, 2 COB, 3 both Gaussian block terms.
template <int MODE, int EVP, int GK>
__device__ __forceinline__ void pair_tile(const float4* __restrict__ sj, const IBead& b,
                                          const PairParams& c, const int self_j, Acc& a) {
  constexpr bool kGauss = (MODE & PM_GAUSS) != 0 && GK != 0;
  constexpr bool kChb = (MODE & PM_CHB) != 0;
  constexpr bool kMask = (MODE & PM_CHBMASK) != 0;
  constexpr bool kSelf = (MODE & PM_SELF) != 0;
  constexpr bool kScaled = kGauss || kChb || kSelf;  // accumulate into fx (scaled) or ux (EV units)
  const float evf = (float)EVP * c.ev_pref;
  const float kc4 = 4.0f * c.chb_kc;
#pragma unroll 8
  for (int jj = 0; jj < MMM_TILE; ++jj) {
    const float4 pj = sj[jj];
    const float dx = b.x - pj.x, dy = b.y - pj.y, dz = b.z - pj.z;
    float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    float live = 1.0f;
    if (kSelf) {
      if (jj == self_j) { r2 = 1.0f; live = 0.0f; }
    }
    const float inv_r = fast_rsqrt(r2);
    const float r = r2 * inv_r;
    const float w = fast_rcp(r + c.ev_rs);
    const float wp = powi<EVP>(w);
    float fs;
    if (kSelf) {
      a.eev = fmaf(live, wp, a.eev);
      fs = evf * live * wp * w * inv_r;
    } else {
      a.eev += wp;
      fs = wp * w * inv_r;
      if (kScaled) fs *= evf;
    }
    const int xr = b.w ^ __float_as_int(pj.w);
    if (kGauss) {
      float g = fast_ex2(r2 * c.g_c);
      if (kSelf) g *= live;
      if (GK & 1) {
        const float g1 = ((xr & 0x7) == 0) ? g : 0.0f;
        a.gscb += g1;
        fs = fmaf(-b.a_scb, g1, fs);
      }
      if (GK & 2) {
        const float g2 = ((xr & 0x18) == 0) ? g : 0.0f;
        a.gcob += g2;
        fs = fmaf(-b.a_cob, g2, fs);
      }
    }
    if (kChb) {
      // E = dE (kC r^4 - r^3 + r^2);  -(dE/dr)/r = -dE (4 kC r^2 - 3 r + 2)
      const float q = fmaf(c.chb_kc, r2, 1.0f - r);
      float e = r2 * q;
      float bb = fmaf(-3.0f, r, fmaf(kc4, r2, 2.0f));
      if (kMask || kSelf) {
        bool same = (xr & 0xFFFF00) == 0;
        if (kSelf) same = same && (jj != self_j);
        e = same ? e : 0.0f;
        bb = same ? bb : 0.0f;
      }
      a.echb += e;
      fs = fmaf(-c.chb_de, bb, fs);
    }
    if (kScaled) {
      a.fx = fmaf(fs, dx, a.fx);
      a.fy = fmaf(fs, dy, a.fy);
      a.fz = fmaf(fs, dz, a.fz);
    } else {
      a.ux = fmaf(fs, dx, a.ux);
      a.uy = fmaf(fs, dy, a.uy);
      a.uz = fmaf(fs, dz, a.uz);
    }
  }
}

struct PairArgs {
  const float4* pos4;
  const TileInfo* tiles;
  double* fpair;   // [nchunk][3][npad]
  double* epair;   // [items][4]
  int* counter;
  const int* skip;
  int64_t n, npad;
  int ntiles, chunk_tiles, nchunk, n_items;
  PairParams pp;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// EVP = integer EV power (0 selects the generic any-form path); GK Gaussian kinds; CHB polynomial on/off.
template <int EVP, int GK, bool CHB>
__global__ void __launch_bounds__(MMM_IBLOCK) k_pair_exact(const PairArgs A) {
  __shared__ __align__(16) float4 s_j[MMM_STAGE];
  __shared__ __align__(16) TileInfo s_tiles[MMM_STAGE / MMM_TILE];
  __shared__ double s_red[4][MMM_IBLOCK / 32];
  __shared__ int s_item;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const PairParams& c = A.pp;
  constexpr int kTilesPerStage = MMM_STAGE / MMM_TILE;
  if (A.skip && *A.skip) return;

  for (;;) {
    __syncthreads();  // protects s_item, s_j and s_red across items
    if (tid == 0) s_item = atomicAdd(A.counter, 1);
    __syncthreads();
    const int item = s_item;
    if (item >= A.n_items) break;
    const int iblk = item / A.nchunk, chunk = item - iblk * A.nchunk;
    const int64_t i = (int64_t)iblk * MMM_IBLOCK + tid;
    const int itile = (int)(i / MMM_TILE);

    const float4 pi = A.pos4[i];
    IBead b;
    b.x = pi.x; b.y = pi.y; b.z = pi.z; b.w = __float_as_int(pi.w);
    const int si = (b.w & 7) - 2;
    b.a_scb = 0.0f; b.a_cob = 0.0f;
    float e_scb_i = 0.0f, e_cob_i = 0.0f;
    if (c.scb_form >= 0 && si != 0) e_scb_i = c.scb_e[si == 2 ? 0 : (si == 1 ? 1 : (si == -1 ? 2 : 3))];
    if (c.cob_form >= 0 && si != 0) e_cob_i = si > 0 ? c.cob_ea : c.cob_eb;
    b.a_scb = e_scb_i * c.g_inv_rc2;
    b.a_cob = e_cob_i * c.g_inv_rc2;
    const TileInfo ib = A.tiles[itile];

    double dfx = 0.0, dfy = 0.0, dfz = 0.0, de0 = 0.0, de1 = 0.0, de2 = 0.0, de3 = 0.0;

    const int jt0 = chunk * A.chunk_tiles;
    const int jt1 = min(jt0 + A.chunk_tiles, A.ntiles);
    for (int jt = jt0; jt < jt1; jt += kTilesPerStage) {
      __syncthreads();
      s_j[tid] = A.pos4[(int64_t)jt * MMM_TILE + tid];  // npad is a multiple of MMM_STAGE
      if (tid < 2 * kTilesPerStage)
        reinterpret_cast<float4*>(s_tiles)[tid] =
            reinterpret_cast<const float4*>(A.tiles + jt)[tid];
      __syncthreads();
      const int nsub = min(kTilesPerStage, jt1 - jt);

      if constexpr (EVP > 0) {
        Acc a;
        a.fx = a.fy = a.fz = a.ux = a.uy = a.uz = 0.0f;
        a.eev = a.gscb = a.gcob = a.echb = 0.0f;
        for (int sub = 0; sub < nsub; ++sub) {
          const TileInfo jb = s_tiles[sub];
          const float4* sj = s_j + sub * MMM_TILE;
          const int jtile = jt + sub;
          // chromosome ranges
          int chb_mode = 0;
          if (CHB) {
            const bool overlap = !(ib.cmax < jb.cmin || jb.cmax < ib.cmin);
            const bool uniform = (ib.cmin == ib.cmax) && (jb.cmin == jb.cmax);
            chb_mode = overlap ? (uniform ? PM_CHB : (PM_CHB | PM_CHBMASK)) : 0;
          }
          if (jtile == itile) {
            pair_tile<PM_GAUSS | (CHB ? (PM_CHB | PM_CHBMASK) : 0) | PM_SELF, EVP, GK>(sj, b, c, lane, a);
            continue;
          }
          bool near = false;
          if (GK != 0) {
            const float ddx = fmaxf(0.0f, fmaxf(ib.lox - jb.hix, jb.lox - ib.hix));
            const float ddy = fmaxf(0.0f, fmaxf(ib.loy - jb.hiy, jb.loy - ib.hiy));
            const float ddz = fmaxf(0.0f, fmaxf(ib.loz - jb.hiz, jb.loz - ib.hiz));
            near = fmaf(ddz, ddz, fmaf(ddy, ddy, ddx * ddx)) < c.rg2;
          }
          if (!near) {
            if (!CHB || chb_mode == 0) pair_tile<0, EVP, GK>(sj, b, c, -1, a);
            else if (chb_mode == PM_CHB) pair_tile<PM_CHB, EVP, GK>(sj, b, c, -1, a);
            else pair_tile<PM_CHB | PM_CHBMASK, EVP, GK>(sj, b, c, -1, a);
          } else {
            if (!CHB || chb_mode == 0) pair_tile<PM_GAUSS, EVP, GK>(sj, b, c, -1, a);
            else pair_tile<PM_GAUSS | PM_CHB | PM_CHBMASK, EVP, GK>(sj, b, c, -1, a);
          }
        }
        // fold the stage into FP64
        const float evf = (float)EVP * c.ev_pref;
        dfx += (double)fmaf(evf, a.ux, a.fx);
        dfy += (double)fmaf(evf, a.uy, a.fy);
        dfz += (double)fmaf(evf, a.uz, a.fz);
        de0 += (double)a.eev;
        de1 += (double)a.gcob;
        de2 += (double)a.gscb;
        de3 += (double)a.echb;
      } else {
        double fx = 0.0, fy = 0.0, fz = 0.0, e4[4] = {0.0, 0.0, 0.0, 0.0};
        for (int sub = 0; sub < nsub; ++sub) {
          const int jtile = jt + sub;
          const float4* sj = s_j + sub * MMM_TILE;
#pragma unroll 2
          for (int jj = 0; jj < MMM_TILE; ++jj) {
            const int64_t j = (int64_t)jtile * MMM_TILE + jj;
            pair_generic(sj[jj], b, si, i < j, c, j != i, fx, fy, fz, e4);
          }
        }
        dfx += fx; dfy += fy; dfz += fz;
        de0 += e4[0]; de1 += e4[1]; de2 += e4[2]; de3 += e4[3];
      }
    }

    // per-(chunk, bead) partial force, SoA planes
    double* fp = A.fpair + (size_t)chunk * 3 * (size_t)A.npad;
    fp[i] = dfx;
    fp[(size_t)A.npad + i] = dfy;
    fp[2 * (size_t)A.npad + i] = dfz;

    // energies: apply per-bead prefactors, halve (each unordered pair was seen twice), reduce
    if constexpr (EVP > 0) {
      de0 *= 0.5 * (double)c.ev_pref;
      de1 *= -0.5 * (double)e_cob_i;
      de2 *= -0.5 * (double)e_scb_i;
      de3 *= 0.5 * (double)c.chb_de;
    } else {
      de0 *= 0.5; de1 *= 0.5; de2 *= 0.5; de3 *= 0.5;
    }
    if (i >= A.n) { de0 = de1 = de2 = de3 = 0.0; }
    de0 = warp_sum(de0); de1 = warp_sum(de1); de2 = warp_sum(de2); de3 = warp_sum(de3);
    if (lane == 0) { s_red[0][warp] = de0; s_red[1][warp] = de1; s_red[2][warp] = de2; s_red[3][warp] = de3; }
    __syncthreads();
    if (tid < 4) {
      double s = 0.0;
#pragma unroll
      for (int w8 = 0; w8 < MMM_IBLOCK / 32; ++w8) s += s_red[tid][w8];
      A.epair[(size_t)item * 4 + tid] = s;
    }
  }
}

template <int EVP, int GK, bool CHB>
int launch_variant(mmm_system* h, const PairArgs& A) {
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pair_exact<EVP, GK, CHB>, MMM_IBLOCK, 0);
  if (occ < 1) occ = 1;
  int grid = h->sm_count * occ;
  if (grid > A.n_items) grid = A.n_items;
  k_pair_exact<EVP, GK, CHB><<<grid, MMM_IBLOCK, 0, h->stream>>>(A);
  return 0;
}

template <int EVP>
int launch_evp(mmm_system* h, const PairArgs& A, int gk, bool chb) {
  switch (gk * 2 + (chb ? 1 : 0)) {
    case 0: return launch_variant<EVP, 0, false>(h, A);
    case 1: return launch_variant<EVP, 0, true>(h, A);
    case 2: return launch_variant<EVP, 1, false>(h, A);
    case 3: return launch_variant<EVP, 1, true>(h, A);
    case 4: return launch_variant<EVP, 2, false>(h, A);
    case 5: return launch_variant<EVP, 2, true>(h, A);
    case 6: return launch_variant<EVP, 3, false>(h, A);
    default: return launch_variant<EVP, 3, true>(h, A);
  }
}

}  // namespace

// The specialised path covers the reference's default forms: power-law EV with an integer
// power in {3, 6}, Gaussian COB/SCB sharing one range, polynomial CHB.
bool mmm_pair_fast_path(const mmm_system* h) { return mmm_pair_fast_path_pp(h->pp); }

bool mmm_pair_fast_path_pp(const PairParams& p) {
  if (p.ev_form != MMM_EV_POWERLAW) return false;
  if (!(p.ev_power == 6.0f || p.ev_power == 3.0f)) return false;
  if (p.cob_form != MMM_FORM_OFF && p.cob_form != MMM_BLOCK_GAUSSIAN) return false;
  if (p.scb_form != MMM_FORM_OFF && p.scb_form != MMM_BLOCK_GAUSSIAN) return false;
  if (p.cob_form >= 0 && p.scb_form >= 0 && p.cob_rc != p.scb_rc) return false;
  if (p.chb_form != MMM_FORM_OFF && p.chb_form != MMM_CHB_POLYNOMIAL) return false;
  return true;
}

// pp_override (cut-off mode): evaluate this parameter set instead of the handle's (the exact
// CHB-only pass that runs beside the cell-list pass); no event timing in that case.
int mmm_launch_pair_exact(mmm_system* h, const int* d_skip, const PairParams* pp_override) {
  PairArgs A;
  A.pos4 = h->d_pos4;
  A.tiles = h->d_tiles;
  A.fpair = h->d_fpair;
  A.epair = h->d_epair;
  A.counter = h->d_counter;
  A.skip = d_skip;
  A.n = h->n;
  A.npad = h->npad;
  A.ntiles = (int)h->ntiles;
  A.nchunk = h->nchunk;
  A.chunk_tiles = h->chunk_tiles;
  A.n_items = (int)((h->npad / MMM_IBLOCK) * h->nchunk);
  A.pp = pp_override ? *pp_override : h->pp;
  MMM_CUDA(h, cudaMemsetAsync(h->d_counter, 0, sizeof(int), h->stream));
  const bool timed = pp_override == nullptr;
  const bool collect = timed && h->ev_cursor >= 0 && (size_t)(2 * h->ev_cursor + 1) < h->ev_pool.size();
  cudaEvent_t ea = collect ? h->ev_pool[2 * h->ev_cursor] : h->ev_a;
  cudaEvent_t eb = collect ? h->ev_pool[2 * h->ev_cursor + 1] : h->ev_b;
  if (collect) h->ev_cursor++;
  if (timed && !h->capturing) MMM_CUDA(h, cudaEventRecord(ea, h->stream));
  if (mmm_pair_fast_path_pp(A.pp)) {
    const int gk = (A.pp.scb_form >= 0 ? 1 : 0) | (A.pp.cob_form >= 0 ? 2 : 0);
    const bool chb = A.pp.chb_form >= 0;
    if (A.pp.ev_power == 6.0f) launch_evp<6>(h, A, gk, chb);
    else launch_evp<3>(h, A, gk, chb);
  } else {
    launch_variant<0, 0, false>(h, A);
  }
  h->launches++;
  MMM_CUDA(h, cudaGetLastError());
  if (timed && !h->capturing) MMM_CUDA(h, cudaEventRecord(eb, h->stream));
  return MMM_OK;
}
