// mmm_analysis.cu — the O(N^2) part of the structure report:
// plots.py:663 `dmat = distance.cdist(V, V); mean_dist = np.mean(dmat)` materialises an N x N FP64
// matrix (320 GB at N = 2e5; the reference's own heat-map is switched off at N >= 50 000,
// model.py:1095).  Here the sum of all pair distances is a tiled gather pass over the centred FP32
// copy: j-beads staged 256 at a time in shared memory, FP32 within a stage, FP64 across, fixed
// summation order.  Issue-bound like the gather pair kernel (7 instructions per ordered pair).
#include "mmm_internal.cuh"

namespace {

__global__ void __launch_bounds__(256) k_sum_distances(const float4* __restrict__ pos4, int64_t n, int64_t npad,
                                                       double* __restrict__ part) {
  __shared__ float4 s_j[256];
  __shared__ double s_red[8];
  const int tid = threadIdx.x;
  const int64_t i = (int64_t)blockIdx.x * 256 + tid;
  const float4 pi = pos4[i < npad ? i : 0];
  double acc = 0.0;
  for (int64_t j0 = 0; j0 < n; j0 += 256) {
    __syncthreads();
    s_j[tid] = pos4[j0 + tid];  // npad is a multiple of 256
    __syncthreads();
    const int cnt = (int)(n - j0 < 256 ? n - j0 : 256);
    float a = 0.0f;
#pragma unroll 8
    for (int jj = 0; jj < cnt; ++jj) {
      const float dx = pi.x - s_j[jj].x, dy = pi.y - s_j[jj].y, dz = pi.z - s_j[jj].z;
      a += sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
    }
    acc += (double)a;
  }
  if (i >= n) acc = 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((tid & 31) == 0) s_red[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += s_red[w];
    part[blockIdx.x] = s;
  }
}

}  // namespace

extern "C" int mmm_mean_pair_distance(mmm_handle h, double* mean_out) {
  if (!h || !mean_out) return MMM_ERR_ARG;
  if (!h->positions_set) return mmm_fail(h, MMM_ERR_STATE, "positions were never set");
  cudaSetDevice(h->device);
  int rc = mmm_launch_prepare(h, nullptr);  // d_x -> centred FP32 copy
  if (rc) return rc;
  const int blocks = (int)((h->n + 255) / 256);
  double* d_part = nullptr;
  MMM_CUDA(h, cudaMalloc((void**)&d_part, sizeof(double) * blocks));
  k_sum_distances<<<blocks, 256, 0, h->stream>>>(h->d_pos4, h->n, h->npad, d_part);
  h->launches++;
  std::vector<double> part((size_t)blocks);
  cudaError_t e = cudaMemcpyAsync(part.data(), d_part, sizeof(double) * blocks, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_part);
  if (e != cudaSuccess) return mmm_fail(h, MMM_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e));
  double s = 0.0;
  for (double p : part) s += p;
  // np.mean over the full N x N matrix, zero diagonal included (plots.py:664)
  *mean_out = s / ((double)h->n * (double)h->n);
  return MMM_OK;
}
