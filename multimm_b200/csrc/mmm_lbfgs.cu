// mmm_lbfgs.cu — on-device L-BFGS; replaces Simulation.minimizeEnergy() (model.py:886), i.e.
// [OpenMM] LocalEnergyMinimizer driving its bundled liblbfgs with
//   m = 6, LBFGS_LINESEARCH_BACKTRACKING_STRONG_WOLFE, ftol 1e-4, wolfe 0.9, max_linesearch 40,
//   epsilon = tol / max(1, rms |x_i|), stop when |g| / max(1,|x|) <= epsilon.
//
// No host round trip per iteration.  Every evaluation is the same kernel sequence
//     k_apply -> k_prepare -> k_pair -> k_assemble -> k_dots (whose last block decides)
// captured once as a CUDA graph and replayed
// and all control flow (Armijo / Wolfe tests, step scaling, acceptance, convergence, history
// rotation) lives in LbfgsState on the device.  The host enqueues batches of evaluations and
// looks at one pinned int per batch; once `done` is set every kernel returns at its first
// instruction.
//
// Vector-free two-loop recursion: k_dots produces, in ONE pass over the vectors, every inner
// product the update needs (MMM_NDOT = 6m + 7); k_decide keeps the Gram matrix of the history,
// runs the two-loop recursion on 2m+1 coefficients in a single thread and k_apply forms
// d = sum_b delta_b b in one more pass.  Algebraically identical to liblbfgs' loop; only the
// floating-point association differs.  HBM-bound: k_dots reads (2m+4) x 24 B per bead,
// k_apply reads/writes (2m+8) x 24 B per bead.
#include <math.h>
#include <string.h>

#include <chrono>

#include "mmm_internal.cuh"

namespace {

constexpr int M = MMM_LBFGS_M;
// layout of the dot-product vector
constexpr int D_DS = 0, D_DY = M, D_YS = 2 * M, D_YY = 3 * M, D_GS = 4 * M, D_GY = 5 * M;
constexpr int D_DD = 6 * M, D_DYN = 6 * M + 1, D_YYN = 6 * M + 2, D_GD = 6 * M + 3, D_GYN = 6 * M + 4,
              D_GG = 6 * M + 5, D_XX = 6 * M + 6;
constexpr int kDotBlock = 128;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One pass over x, g, gp, d and the valid history slots; y = g - gp is formed on the fly.
struct DecideArgs {
  LbfgsState* st;
  const double* epair; int n_items;
  const double* epart; int n_eblocks;
  const double* dpart; int n_dblocks;
  double* eterms;
  unsigned* ticket;  // blocks of k_dots that have finished (last-block-done)
};

__device__ void decide(const DecideArgs& A);
__device__ void decide_serial(LbfgsState* st, const double* s_val, double* eterms, double (*G)[MMM_LBFGS_M][MMM_LBFGS_M]);

// decide != 0: the block that finishes last goes on to play liblbfgs on the sums (one launch less
// per evaluation, which is what a small system's iteration time is made of).
__global__ void __launch_bounds__(kDotBlock) k_dots(const LbfgsState* __restrict__ st, int64_t n3,
                                                    const double* __restrict__ x, const double* __restrict__ g,
                                                    const double* __restrict__ gp, const double* __restrict__ d,
                                                    const double* __restrict__ S, const double* __restrict__ Y,
                                                    double* __restrict__ dpart, const DecideArgs DA, const int decide_too) {
  if (st->done) return;
  __shared__ double s_red[kDotBlock / 32][MMM_NDOT];
  const int bound = st->bound;
  double acc[MMM_NDOT];
#pragma unroll
  for (int q = 0; q < MMM_NDOT; ++q) acc[q] = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * kDotBlock + threadIdx.x; e < n3; e += (int64_t)gridDim.x * kDotBlock) {
    const double xe = x[e], ge = g[e], de = d[e], ye = ge - gp[e];
#pragma unroll
    for (int l = 0; l < M; ++l) {
      if (l < bound) {
        const double sl = S[(size_t)l * n3 + e], yl = Y[(size_t)l * n3 + e];
        acc[D_DS + l] += de * sl;
        acc[D_DY + l] += de * yl;
        acc[D_YS + l] += ye * sl;
        acc[D_YY + l] += ye * yl;
        acc[D_GS + l] += ge * sl;
        acc[D_GY + l] += ge * yl;
      }
    }
    acc[D_DD] += de * de;
    acc[D_DYN] += de * ye;
    acc[D_YYN] += ye * ye;
    acc[D_GD] += ge * de;
    acc[D_GYN] += ge * ye;
    acc[D_GG] += ge * ge;
    acc[D_XX] += xe * xe;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < MMM_NDOT; ++q) {
    const double v = warp_sum_d(acc[q]);
    if (lane == 0) s_red[warp][q] = v;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < MMM_NDOT; q += kDotBlock) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kDotBlock / 32; ++w) s += s_red[w][q];
    dpart[(size_t)blockIdx.x * MMM_NDOT + q] = s;
  }
  if (!decide_too) return;
  __shared__ bool s_last;
  __threadfence();  // this block's partials are visible before its ticket is
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(DA.ticket, 1u);
    s_last = t == gridDim.x - 1;
    if (s_last) *DA.ticket = 0u;  // ready for the next evaluation
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  decide(DA);
}

// One block (the last of k_dots).  Warps reduce the partial sums in a fixed order — independent of
// which block happened to be last — then thread 0 plays liblbfgs on coefficients held in shared memory.
__device__ void decide(const DecideArgs& A) {
  LbfgsState* st = A.st;
  __shared__ double s_val[MMM_NUM_TERMS + MMM_NDOT];
  __shared__ double s_G[3][M][M];  // Gss, Gsy, Gyy
  constexpr int kWarps = kDotBlock / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q = threadIdx.x; q < 3 * M * M; q += kDotBlock) {
    const int m = q / (M * M), r = (q / M) % M, c = q % M;
    s_G[m][r][c] = m == 0 ? st->Gss[r][c] : (m == 1 ? st->Gsy[r][c] : st->Gyy[r][c]);
  }
  for (int q = warp; q < MMM_NUM_TERMS + MMM_NDOT; q += kWarps) {
    double acc = 0.0;
    if (q < 4) {
      for (int b = lane; b < A.n_items; b += 32) acc += A.epair[(size_t)b * 4 + q];
    } else if (q < MMM_NUM_TERMS) {
      for (int b = lane; b < A.n_eblocks; b += 32) acc += A.epart[(size_t)b * 6 + (q - 4)];
    } else {
      for (int b = lane; b < A.n_dblocks; b += 32) acc += A.dpart[(size_t)b * MMM_NDOT + (q - MMM_NUM_TERMS)];
    }
    acc = warp_sum_d(acc);
    if (lane == 0) s_val[q] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) decide_serial(st, s_val, A.eterms, s_G);
  __syncthreads();
  for (int q = threadIdx.x; q < 3 * M * M; q += kDotBlock) {
    const int m = q / (M * M), r = (q / M) % M, c = q % M;
    (m == 0 ? st->Gss[r][c] : (m == 1 ? st->Gsy[r][c] : st->Gyy[r][c])) = s_G[m][r][c];
  }
}

__device__ void decide_serial(LbfgsState* st, const double* s_val, double* eterms, double (*G)[M][M]) {
  double (*Gss)[M] = G[0];
  double (*Gsy)[M] = G[1];
  double (*Gyy)[M] = G[2];
  const double* D = s_val + MMM_NUM_TERMS;
  double f = 0.0;
  for (int t = 0; t < MMM_NUM_TERMS; ++t) {
    f += s_val[t];
    st->e_terms[t] = s_val[t];
    eterms[t] = s_val[t];
  }
  st->evaluations++;
  const double ftol = 1e-4, wolfe = 0.9, min_step = 1e-20, max_step = 1e20;
  const int max_ls = 40;
  const bool finite = isfinite(f) && isfinite(D[D_GG]);

  if (st->phase == 1) {  // first evaluation
    st->e_initial = f;
    st->fx = f;
    const double gnorm = sqrt(D[D_GG]);
    double xnorm = sqrt(D[D_XX]);
    if (xnorm < 1.0) xnorm = 1.0;
    st->gnorm = gnorm;
    st->xnorm = xnorm;
    if (!finite) { st->done = 4; st->flag = APPLY_NONE; return; }
    if (gnorm / xnorm <= st->epsilon) { st->done = 1; st->flag = APPLY_NONE; return; }
    st->step = 1.0 / gnorm;  // d = -g
    st->finit = f;
    st->dginit = -D[D_GG];
    st->ls_count = 0;
    st->phase = 2;
    st->flag = APPLY_INIT;
    return;
  }

  // line search (backtracking, strong Wolfe)
  st->ls_count++;
  const double step = st->step, dg = D[D_GD];
  bool accept = false;
  double width = 0.5;
  if (!finite || !(f <= st->finit + step * ftol * st->dginit)) width = 0.5;
  else if (dg < wolfe * st->dginit) width = 2.1;
  else if (dg > -wolfe * st->dginit) width = 0.5;
  else accept = true;

  if (!accept) {
    int code = 0;
    if (step < min_step) code = -2;
    else if (step > max_step) code = -3;
    else if (st->ls_count >= max_ls) code = -4;
    if (code) {  // liblbfgs reverts to the previous point and returns
      st->done = 3;
      st->ls_status = code;
      st->flag = APPLY_RESTORE;
    } else {
      st->step = step * width;
      st->flag = APPLY_RETRY;
    }
    return;
  }

  st->iterations++;
  st->fx = f;
  const double gnorm = sqrt(D[D_GG]);
  double xnorm = sqrt(D[D_XX]);
  if (xnorm < 1.0) xnorm = 1.0;
  st->gnorm = gnorm;
  st->xnorm = xnorm;
  if (gnorm / xnorm <= st->epsilon) { st->done = 1; st->flag = APPLY_NONE; return; }
  if (st->max_iter != 0 && st->max_iter < st->k + 1) { st->done = 2; st->flag = APPLY_NONE; return; }

  // history update: s_e = step * d, y_e = g - gp go to slot e
  const int e = st->end, bound_old = st->bound;
  const double ys = step * D[D_DYN], yy = D[D_YYN];
  double gs[M], gy[M];
  for (int l = 0; l < M; ++l) {
    gs[l] = D[D_GS + l];
    gy[l] = D[D_GY + l];
    if (l == e || l >= bound_old) continue;
    Gss[e][l] = Gss[l][e] = step * D[D_DS + l];
    Gsy[e][l] = step * D[D_DY + l];  // s_e . y_l
    Gsy[l][e] = D[D_YS + l];         // s_l . y_e
    Gyy[e][l] = Gyy[l][e] = D[D_YY + l];
  }
  Gss[e][e] = step * step * D[D_DD];
  Gsy[e][e] = ys;
  Gyy[e][e] = yy;
  st->ys[e] = ys;
  gs[e] = step * D[D_GD];
  gy[e] = D[D_GYN];
  const double gg = D[D_GG];

  const int bound = (M <= st->k) ? M : (int)st->k;
  st->k++;
  const int end = (e + 1) % M;
  st->end = end;
  st->bound = bound;

  // two-loop recursion on coefficients over {s_l}, {y_l}, g
  double cs[M], cy[M], alpha[M], cg = -1.0;
  for (int l = 0; l < M; ++l) { cs[l] = 0.0; cy[l] = 0.0; alpha[l] = 0.0; }
  int j = end;
  for (int q = 0; q < bound; ++q) {
    j = (j + M - 1) % M;
    double sd = cg * gs[j];
    for (int l = 0; l < bound; ++l) sd += cs[l] * Gss[j][l] + cy[l] * Gsy[j][l];
    alpha[j] = sd / st->ys[j];
    cy[j] -= alpha[j];
  }
  const double scale = ys / yy;
  cg *= scale;
  for (int l = 0; l < M; ++l) { cs[l] *= scale; cy[l] *= scale; }
  for (int q = 0; q < bound; ++q) {
    double yd = cg * gy[j];
    for (int l = 0; l < bound; ++l) yd += cs[l] * Gsy[l][j] + cy[l] * Gyy[j][l];
    const double beta = yd / st->ys[j];
    cs[j] += alpha[j] - beta;
    j = (j + 1) % M;
  }
  double dginit = cg * gg;
  for (int l = 0; l < bound; ++l) dginit += cs[l] * gs[l] + cy[l] * gy[l];
  for (int l = 0; l < M; ++l) { st->delta[l] = cs[l]; st->delta[M + l] = cy[l]; }
  st->delta[2 * M] = cg;
  st->slot = e;
  st->finit = f;
  st->dginit = dginit;
  st->step = 1.0;
  st->ls_count = 0;
  st->flag = APPLY_ACCEPT;
  if (!(dginit < 0.0)) {  // LBFGSERR_INCREASEGRADIENT: keep the accepted point and stop
    st->done = 3;
    st->ls_status = -1;
    st->flag = APPLY_NONE;
  }
}

// Element-wise update of the L-BFGS vectors and the trial point.
__global__ void __launch_bounds__(256) k_apply(const LbfgsState* __restrict__ st, int64_t n3,
                                               double* __restrict__ x, double* __restrict__ g,
                                               double* __restrict__ xp, double* __restrict__ gp,
                                               double* __restrict__ d, double* __restrict__ S,
                                               double* __restrict__ Y) {
  const int flag = st->flag;
  if (flag == APPLY_NONE) return;
  const double step = st->step;
  const int slot = st->slot, bound = st->bound;
  double cs[M], cy[M];
#pragma unroll
  for (int l = 0; l < M; ++l) { cs[l] = st->delta[l]; cy[l] = st->delta[M + l]; }
  const double cg = st->delta[2 * M];
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n3; e += (int64_t)gridDim.x * blockDim.x) {
    if (flag == APPLY_RETRY) {
      x[e] = xp[e] + step * d[e];
    } else if (flag == APPLY_INIT) {
      const double xe = x[e], ge = g[e];
      xp[e] = xe;
      gp[e] = ge;
      d[e] = -ge;
      x[e] = xe - step * ge;
    } else if (flag == APPLY_ACCEPT) {
      const double xe = x[e], ge = g[e];
      const double sn = xe - xp[e], yn = ge - gp[e];
      S[(size_t)slot * n3 + e] = sn;
      Y[(size_t)slot * n3 + e] = yn;
      xp[e] = xe;
      gp[e] = ge;
      double dn = cg * ge;
#pragma unroll
      for (int l = 0; l < M; ++l) {
        if (l < bound) {
          const double sl = (l == slot) ? sn : S[(size_t)l * n3 + e];
          const double yl = (l == slot) ? yn : Y[(size_t)l * n3 + e];
          dn += cs[l] * sl + cy[l] * yl;
        }
      }
      d[e] = dn;
      x[e] = xe + step * dn;
    } else {  // APPLY_RESTORE
      x[e] = xp[e];
      g[e] = gp[e];
    }
  }
}

__global__ void k_clear_flag(LbfgsState* st) { st->flag = APPLY_NONE; }

}  // namespace

int mmm_launch_dots_decide(mmm_system* h) {
  const int64_t n3 = 3 * h->n;
  DecideArgs A;
  A.st = h->d_lb;
  A.epair = h->d_epair; A.n_items = (int)h->n_items;
  A.epart = h->d_epart; A.n_eblocks = h->n_red_blocks;
  A.dpart = h->d_dpart; A.n_dblocks = h->n_dot_blocks;
  A.eterms = h->d_eterms;
  A.ticket = reinterpret_cast<unsigned*>(h->d_counter + 2);
  k_dots<<<h->n_dot_blocks, kDotBlock, 0, h->stream>>>(h->d_lb, n3, h->d_x, h->d_g, h->d_gp, h->d_d, h->d_S,
                                                       h->d_Y, h->d_dpart, A, 1);
  h->launches += 1;
  MMM_CUDA(h, cudaGetLastError());
  return MMM_OK;
}

// One L-BFGS evaluation as it is enqueued: apply -> prepare -> pair -> assemble -> dots (+ decide).
static int enqueue_evaluation(mmm_system* h, int grid_apply) {
  const int64_t n3 = 3 * h->n;
  const int* d_done = &h->d_lb->done;
  k_apply<<<grid_apply, 256, 0, h->stream>>>(h->d_lb, n3, h->d_x, h->d_g, h->d_xp, h->d_gp, h->d_d, h->d_S, h->d_Y);
  h->launches++;
  int rc;
  if ((rc = mmm_evaluate(h, d_done))) return rc;
  return mmm_launch_dots_decide(h);
}

int mmm_run_minimize(mmm_system* h, double tol, int64_t max_iter, mmm_min_report* out) {
  const int64_t n3 = 3 * h->n;
  const auto t0 = std::chrono::steady_clock::now();
  // vectors start at zero so the first k_dots pass reads defined values
  MMM_CUDA(h, cudaMemsetAsync(h->d_xp, 0, sizeof(double) * n3, h->stream));
  MMM_CUDA(h, cudaMemsetAsync(h->d_gp, 0, sizeof(double) * n3, h->stream));
  MMM_CUDA(h, cudaMemsetAsync(h->d_d, 0, sizeof(double) * n3, h->stream));
  MMM_CUDA(h, cudaMemsetAsync(h->d_counter + 2, 0, sizeof(int), h->stream));

  // epsilon = tol / max(1, sqrt(sum |x_i|^2 / N)), [OpenMM] LocalEnergyMinimizer::minimize
  // (one host read of |x|^2 before the loop starts; nothing is read back per iteration)
  LbfgsState init;
  memset(&init, 0, sizeof(init));
  init.phase = 1;
  init.k = 1;
  init.max_iter = max_iter;
  init.epsilon = -1.0;  // filled below
  {
    // |x|^2 through the dot kernel: run it once with a dummy state
    MMM_CUDA(h, cudaMemcpyAsync(h->d_lb, &init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
    MMM_CUDA(h, cudaMemsetAsync(h->d_g, 0, sizeof(double) * n3, h->stream));
    DecideArgs none{};
    k_dots<<<h->n_dot_blocks, kDotBlock, 0, h->stream>>>(h->d_lb, n3, h->d_x, h->d_g, h->d_gp, h->d_d,
                                                         h->d_S, h->d_Y, h->d_dpart, none, 0);
    h->launches++;
    std::vector<double> part((size_t)h->n_dot_blocks * MMM_NDOT);
    MMM_CUDA(h, cudaMemcpyAsync(part.data(), h->d_dpart, part.size() * sizeof(double), cudaMemcpyDeviceToHost,
                                h->stream));
    MMM_CUDA(h, cudaStreamSynchronize(h->stream));
    double xx = 0.0;
    for (int b = 0; b < h->n_dot_blocks; ++b) xx += part[(size_t)b * MMM_NDOT + D_XX];
    double norm = xx / (double)h->n;
    norm = norm < 1.0 ? 1.0 : sqrt(norm);
    init.epsilon = tol / norm;
  }
  MMM_CUDA(h, cudaMemcpyAsync(h->d_lb, &init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
  const int* d_done = &h->d_lb->done;

  // first evaluation at x0 (also sizes every scratch buffer, so nothing allocates after this point)
  int rc;
  if ((rc = mmm_evaluate(h, d_done))) return rc;
  if ((rc = mmm_launch_dots_decide(h))) return rc;

  const int grid_apply = std::min<int64_t>((n3 + 255) / 256, (int64_t)h->sm_count * 8);
  h->sort_age = 0;  // cut-off mode: the loop starts a fresh Morton-order period, with or without the graph

  // One period of evaluations captured as a CUDA graph and replayed: the 6-7 small launches of an
  // evaluation cost more host and launch latency than GPU time on a small system (configs[0]).  In
  // cut-off mode a period is one life of the Morton order (kResortEvery evaluations, the first of
  // which re-sorts).  Not with several GPUs (the collective stays a plain stream operation).
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  int period = 1;
  const bool use_graph = !h->nccl_comm && !h->no_graph;
  if (use_graph) {
    period = (h->pair_mode == 3 && h->cut_n3) ? mmm_cutoff_resort_period() : 1;
    MMM_CUDA(h, cudaStreamSynchronize(h->stream));
    h->capturing = true;
    h->sort_age = 0;
    const int64_t launches_before = h->launches;
    cudaError_t ce = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
    rc = MMM_OK;
    if (ce == cudaSuccess) {
      for (int q = 0; q < period && rc == MMM_OK; ++q) rc = enqueue_evaluation(h, grid_apply);
      ce = cudaStreamEndCapture(h->stream, &graph);
    }
    h->capturing = false;
    h->launches_per_period = h->launches - launches_before;
    h->launches = launches_before;
    h->sort_age = 0;  // the replayed period starts with a re-sort, exactly as captured
    if (rc != MMM_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess || !graph || cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) {
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      graph = nullptr;
      gexec = nullptr;
      period = 1;  // fall back to plain launches
    }
  }

  int batch = 4;
  LbfgsState fin;
  for (;;) {
    const auto tb0 = std::chrono::steady_clock::now();
    for (int b = 0; b < batch; ++b) {
      if (gexec) {
        MMM_CUDA(h, cudaGraphLaunch(gexec, h->stream));
        h->launches += h->launches_per_period;
      } else if ((rc = enqueue_evaluation(h, grid_apply))) {
        return rc;
      }
    }
    MMM_CUDA(h, cudaMemcpyAsync(h->h_done, d_done, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    MMM_CUDA(h, cudaStreamSynchronize(h->stream));
    if (*h->h_done) break;
    // aim for ~50 ms of queued work per host look
    const double tb = std::chrono::duration<double>(std::chrono::steady_clock::now() - tb0).count();
    const double per_eval = tb / batch;
    int want = (int)(0.05 / (per_eval > 1e-6 ? per_eval : 1e-6));
    batch = want < 4 ? 4 : (want > 512 ? 512 : want);
    // several GPUs: every rank must enqueue the same number of evaluations (each contains a
    // collective), so the batch size may not depend on this rank's clock
    if (h->nccl_comm) batch = 8;
  }
  if (gexec) cudaGraphExecDestroy(gexec);
  if (graph) cudaGraphDestroy(graph);
  h->sort_age = 0;  // whatever comes next re-sorts
  // a failed line search leaves a pending RESTORE (x <- xp)
  k_apply<<<grid_apply, 256, 0, h->stream>>>(h->d_lb, n3, h->d_x, h->d_g, h->d_xp, h->d_gp, h->d_d, h->d_S, h->d_Y);
  k_clear_flag<<<1, 1, 0, h->stream>>>(h->d_lb);
  h->launches += 2;
  MMM_CUDA(h, cudaMemcpyAsync(&fin, h->d_lb, sizeof(fin), cudaMemcpyDeviceToHost, h->stream));
  MMM_CUDA(h, cudaStreamSynchronize(h->stream));
  MMM_CUDA(h, cudaGetLastError());
  if (out) {
    out->iterations = fin.iterations;
    out->evaluations = fin.evaluations;
    out->e_initial = fin.e_initial;
    out->e_final = fin.fx;
    out->rms_force = fin.gnorm / sqrt((double)h->n);
    out->converged = fin.done == 1;
    out->ls_status = fin.done == 3 ? fin.ls_status : 0;
    out->wall_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  if (fin.done == 4) return mmm_fail(h, MMM_ERR_NUMERIC, "non-finite energy or gradient at the start of minimisation");
  return MMM_OK;
}
