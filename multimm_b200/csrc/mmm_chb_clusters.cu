// mmm_chb_clusters.cu — CHB on cluster centroids: the surrogate the COARSE stage of the opt-in
// two-stage minimisation uses for the chromosomal-block term (model.py:386-451, polynomial form).
//
// CHB's polynomial dE (kC r^4 - r^3 + r^2) grows with r and cannot be truncated, so a cut-off
// evaluation has to pay an exact all-pairs pass over every same-chromosome pair (1.08e9 pairs at
// N = 2e5: as much as the whole truncated EV/SCB pass).  The term is weak (dE = 1e-4) and smooth,
// and the coarse stage only has to come close to the minimum — the exact stage that follows meets
// the stopping rule on the reference's potential.  So, when mmm_set_chb_surrogate(h, 1) is set in
// cut-off mode, CHB is evaluated on CLUSTERS: runs of at most 32 consecutive beads of one chromosome
// (the 32-bead tiles of the chain order, split where the chromosome changes),
//     E_s = dE * sum_{T < T', same chromosome} n_T n_T' f(|c_T - c_T'|),   f(r) = kC r^4 - r^3 + r^2,
// with c_T the centroid.  It is a proper potential (a function of the positions through the
// centroids) whose gradient is exact: every bead of T feels -dE sum_T' n_T' f'(r)/r (c_T - c_T'), so
// L-BFGS's line search stays consistent.  O(N + clusters^2 / chromosomes): 9e5 interactions at
// N = 2e5 instead of 1.08e9.  Never used in exact mode, never by default.
//   k_cl_centroid   one warp per cluster: FP64 centroid of its beads             24 B read per bead
//   k_cl_forces     one thread per cluster, loop over the clusters of its chromosome (FP64)
// k_assemble adds the cluster's force to each of its beads.
#include <algorithm>
#include <numeric>

#include "mmm_internal.cuh"

namespace {

__global__ void __launch_bounds__(256) k_cl_centroid(const double* __restrict__ x, const int* __restrict__ cl_start,
                                                     int ncl, double* __restrict__ cen, const int* __restrict__ skip) {
  if (skip && *skip) return;
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= ncl) return;
  const int b0 = cl_start[c], b1 = cl_start[c + 1];
  double sx = 0.0, sy = 0.0, sz = 0.0;
  const int i = b0 + lane;
  if (i < b1) { sx = x[3 * (size_t)i]; sy = x[3 * (size_t)i + 1]; sz = x[3 * (size_t)i + 2]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
    sz += __shfl_xor_sync(0xffffffffu, sz, o);
  }
  if (lane == 0) {
    const double inv = 1.0 / (double)(b1 - b0);
    cen[4 * (size_t)c] = sx * inv;
    cen[4 * (size_t)c + 1] = sy * inv;
    cen[4 * (size_t)c + 2] = sz * inv;
    cen[4 * (size_t)c + 3] = (double)(b1 - b0);
  }
}

// Thread q handles the q-th cluster in chromosome-sorted order; its chromosome's clusters are the
// slots [range.x, range.y) of that order.
__global__ void __launch_bounds__(128) k_cl_forces(const double* __restrict__ cen, const int* __restrict__ by_chrom,
                                                   const int2* __restrict__ range, int ncl, double kc, double de,
                                                   double* __restrict__ force, double* __restrict__ epair_slots,
                                                   const int* __restrict__ skip) {
  if (skip && *skip) return;
  __shared__ double s_red[4];
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0.0;
  if (q < ncl) {
    const int T = by_chrom[q];
    const int2 rg = range[q];
    const double cx = cen[4 * (size_t)T], cy = cen[4 * (size_t)T + 1], cz = cen[4 * (size_t)T + 2], nT = cen[4 * (size_t)T + 3];
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int p = rg.x; p < rg.y; ++p) {
      if (p == q) continue;
      const int U = by_chrom[p];
      const double dx = cx - cen[4 * (size_t)U], dy = cy - cen[4 * (size_t)U + 1], dz = cz - cen[4 * (size_t)U + 2];
      const double nU = cen[4 * (size_t)U + 3];
      const double r2 = dx * dx + dy * dy + dz * dz, r = sqrt(r2);
      e += nU * r2 * (kc * r2 - r + 1.0);                  // f(r)
      const double g = -nU * (4.0 * kc * r2 - 3.0 * r + 2.0);  // -f'(r) / r
      fx += g * dx; fy += g * dy; fz += g * dz;
    }
    force[3 * (size_t)T] = de * fx;
    force[3 * (size_t)T + 1] = de * fy;
    force[3 * (size_t)T + 2] = de * fz;
    e *= 0.5 * de * nT;  // every cluster pair is seen from both sides
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = e;
  __syncthreads();
  if (threadIdx.x == 0) {
    double* slot = epair_slots + 4 * (size_t)blockIdx.x;
    slot[0] = slot[1] = slot[2] = 0.0;
    slot[3] = s_red[0] + s_red[1] + s_red[2] + s_red[3];  // the CHB energy slot
  }
}

template <typename T>
int upload_vec(mmm_system* h, T** dptr, const std::vector<T>& v) {
  if (*dptr) { cudaFree(*dptr); *dptr = nullptr; }
  if (v.empty()) return MMM_OK;
  MMM_CUDA(h, cudaMalloc((void**)dptr, v.size() * sizeof(T)));
  MMM_CUDA(h, cudaMemcpyAsync(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
  MMM_CUDA(h, cudaStreamSynchronize(h->stream));
  return MMM_OK;
}

}  // namespace

int mmm_chb_clusters_blocks(const mmm_system* h) { return (h->n_clusters + 127) / 128; }

// Clusters are static (chromosome ids and the chain order do not change): built once per scratch.
int mmm_chb_clusters_build(mmm_system* h) {
  const int n = (int)h->n;
  std::vector<int> start, chrom, of_bead((size_t)n);
  for (int i = 0; i < n; ++i) {
    const int c = h->h_chrom.empty() ? 0 : h->h_chrom[(size_t)i];
    if (i % MMM_TILE == 0 || c != chrom.back()) { start.push_back(i); chrom.push_back(c); }
    of_bead[(size_t)i] = (int)start.size() - 1;
  }
  const int ncl = (int)start.size();
  start.push_back(n);
  std::vector<int> by_chrom((size_t)ncl);
  std::iota(by_chrom.begin(), by_chrom.end(), 0);
  std::stable_sort(by_chrom.begin(), by_chrom.end(), [&](int a, int b) { return chrom[(size_t)a] < chrom[(size_t)b]; });
  std::vector<int2> range((size_t)ncl);
  for (int q = 0; q < ncl;) {
    int e = q;
    while (e < ncl && chrom[(size_t)by_chrom[(size_t)e]] == chrom[(size_t)by_chrom[(size_t)q]]) ++e;
    for (int p = q; p < e; ++p) range[(size_t)p] = make_int2(q, e);
    q = e;
  }
  h->n_clusters = ncl;
  int rc;
  if ((rc = upload_vec(h, &h->d_cl_start, start))) return rc;
  if ((rc = upload_vec(h, &h->d_cl_of_bead, of_bead))) return rc;
  if ((rc = upload_vec(h, &h->d_cl_by_chrom, by_chrom))) return rc;
  if ((rc = upload_vec(h, &h->d_cl_range, range))) return rc;
  if (h->d_cl_cen) { cudaFree(h->d_cl_cen); h->d_cl_cen = nullptr; }
  if (h->d_cl_force) { cudaFree(h->d_cl_force); h->d_cl_force = nullptr; }
  MMM_CUDA(h, cudaMalloc((void**)&h->d_cl_cen, sizeof(double) * 4 * (size_t)ncl));
  MMM_CUDA(h, cudaMalloc((void**)&h->d_cl_force, sizeof(double) * 3 * (size_t)ncl));
  return MMM_OK;
}

int mmm_launch_chb_clusters(mmm_system* h, const int* d_skip) {
  const int ncl = h->n_clusters;
  k_cl_centroid<<<(ncl + 7) / 8, 256, 0, h->stream>>>(h->d_x, h->d_cl_start, ncl, h->d_cl_cen, d_skip);
  k_cl_forces<<<mmm_chb_clusters_blocks(h), 128, 0, h->stream>>>(h->d_cl_cen, h->d_cl_by_chrom, h->d_cl_range, ncl,
                                                                 h->pp.d_chb[0], h->pp.d_chb[1], h->d_cl_force,
                                                                 h->d_epair + 4 * (size_t)h->cl_item0, d_skip);
  h->launches += 2;
  MMM_CUDA(h, cudaGetLastError());
  return MMM_OK;
}
