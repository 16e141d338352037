// microbench_mma.cu — could the tensor cores take FMA-pipe work off the exact pair kernel's hot loop?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/microbench_mma scripts/microbench_mma.cu
// The hot body spends 19 FMA-pipe lane-operations per pair; 3 of them form r^2 and 6 accumulate the two
// forces.  r^2 = |xi|^2 + |xj|^2 - 2 xi.xj and F_i = sum_j fs_ij (xi - xj) are contractions, and the legacy
// warp-level MMA (mma.sync m16n8k8, TF32 operands split hi + lo, FP32 accumulate) could form them:
//   r^2 of a 16 x 8 block of pairs: 2 MMAs (K = 8 slots x 2: hi.hi, lo.hi, hi.lo cross terms and the norms
//     in three TF32 parts each);
//   i-side force: [sum fs xj, sum fs] = FS[16 x 8] . Xj[8 x 4]: 3 MMAs per block (fs hi/lo x xj hi/lo), the
//     r^2 accumulator layout IS the A-operand layout after a permutation of the contraction index;
//   j side stays on the FMA pipe (it needs the transposed ownership).
// That leaves 14 FMA-pipe operations per pair and adds 80 MMAs per 2 048 pairs.  This file measures
//   mma      : issue rate of mma.sync.m16n8k8.tf32 alone (cycles per MMA per SM sub-partition)
//   hot      : the shipped body (19 FMA-pipe ops + 2 MUFU per pair), as in microbench_issue.cu
//   hot_mma  : the body with r^2 and the i-side force on the tensor cores — same dependencies and
//              instruction mix as the real thing would have, values meaningless
// and prints cycles per warp-pair per SMSP.
#include <cuda_runtime.h>
#include <stdio.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 x, u64 y, u64 z) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(y), "l"(z)); return d; }
__device__ __forceinline__ u64 mul2(u64 x, u64 y) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(x), "l"(y)); return d; }
__device__ __forceinline__ u64 add2(u64 x, u64 y) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(x), "l"(y)); return d; }
__device__ __forceinline__ float fsqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float frcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// D = A (16 x 8, tf32) . B (8 x 8, tf32) + C, FP32 accumulate
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

constexpr int ITERS = 4096;

__global__ void k_mma(float* out) {
  float acc[8][4];
  unsigned a[4], b[2];
  for (int q = 0; q < 4; ++q) a[q] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + q);
  b[0] = __float_as_uint(0.5f); b[1] = __float_as_uint(0.25f);
  for (int q = 0; q < 8; ++q) for (int r = 0; r < 4; ++r) acc[q][r] = 0.f;
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int q = 0; q < 8; ++q) mma_tf32(acc[q], a, b);
  }
  float s = 0;
  for (int q = 0; q < 8; ++q) for (int r = 0; r < 4; ++r) s += acc[q][r];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shipped body (reference point): lane = (a, b), 8 i-beads as 4 packed pairs, j duplicated in shared memory
__global__ void k_hot(float* out, float rs, int trips) {
  __shared__ float4 s_xy[32];
  __shared__ float2 s_z[32];
  if (threadIdx.x < 32) {
    const float t = 0.1f * threadIdx.x;
    s_xy[threadIdx.x] = make_float4(-t, -t, -2 * t, -2 * t);
    s_z[threadIdx.x] = make_float2(-3 * t, -3 * t);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, a = lane >> 2, b = lane & 3;
  u64 x2[4], y2[4], z2[4], fx[4], fy[4], fz[4];
  for (int m = 0; m < 4; ++m) {
    x2[m] = pk2(1.0f + lane + m, 1.5f + lane + m); y2[m] = pk2(2.0f + m, 2.5f + m); z2[m] = pk2(3.0f + m, 3.5f + m);
    fx[m] = fy[m] = fz[m] = pk2(0.f, 0.f);
  }
  u64 ev2 = pk2(0.f, 0.f);
  const u64 rs2 = pk2(rs, rs);
  float jacc = 0.f;
#pragma unroll 1
  for (int t = 0; t < trips; ++t) {
#pragma unroll 2
    for (int g = 0; g < 4; ++g) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int jl = (((2 * g + k) ^ a) << 2) | b;
        const float4 nxy = s_xy[jl];
        const float2 nz = s_z[jl];
        const u64 njx = pk2(nxy.x, nxy.y), njy = pk2(nxy.z, nxy.w), njz = pk2(nz.x, nz.y);
        u64 ax = pk2(0.f, 0.f), ay = ax, az = ax;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const u64 dx = add2(x2[m], njx), dy = add2(y2[m], njy), dz = add2(z2[m], njz);
          const u64 r2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
          float r2a, r2b;
          unpk2(r2, r2a, r2b);
          const u64 r = pk2(fsqrt(r2a), fsqrt(r2b));
          const u64 q = fma2(rs2, r, r2);
          float qa, qb;
          unpk2(q, qa, qb);
          const u64 wr = pk2(frcp(qa), frcp(qb));
          const u64 w = mul2(r, wr);
          const u64 w2 = mul2(w, w);
          const u64 w3 = mul2(w2, w);
          const u64 wp = mul2(w3, w3);
          const u64 fs = mul2(wp, wr);
          ev2 = add2(ev2, wp);
          fx[m] = fma2(fs, dx, fx[m]); fy[m] = fma2(fs, dy, fy[m]); fz[m] = fma2(fs, dz, fz[m]);
          ax = fma2(fs, dx, ax); ay = fma2(fs, dy, ay); az = fma2(fs, dz, az);
        }
        float lo, hi;
        unpk2(ax, lo, hi); jacc += lo + hi;
        unpk2(ay, lo, hi); jacc += lo + hi;
        unpk2(az, lo, hi); jacc += lo + hi;
      }
    }
  }
  float s = jacc, lo, hi;
  for (int m = 0; m < 4; ++m) {
    unpk2(fx[m], lo, hi); s += lo + hi; unpk2(fy[m], lo, hi); s += lo + hi; unpk2(fz[m], lo, hi); s += lo + hi;
  }
  unpk2(ev2, lo, hi);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + lo + hi;
}

// Tensor-core variant.  Lane (g = lane >> 2, t = lane & 3).  A step = 64 i-beads x 32 j-beads per warp as
// 4 (mt) x 4 (nt) blocks of 16 x 8 pairs; a lane owns in block (mt, nt) the pairs
// i in {g + 16 mt, g + 8 + 16 mt} x j in {2t + 8 nt, 2t + 1 + 8 nt}: accumulator registers c0..c3.
//   JSIDE_MMA false: j side on the FMA pipe (dx from registers), i side on the tensor cores
template <bool I_SIDE_MMA>
__global__ void k_hot_mma(float* out, float rs, int trips) {
  __shared__ unsigned s_a[4][2][4][32];   // r^2 A fragments of the 4 m-tiles (2 MMAs x 4 registers), per lane
  __shared__ unsigned s_b[4][4][32];      // r^2 B fragments of the 4 n-tiles (2 MMAs x 2 registers), per lane
  __shared__ unsigned s_bx[4][4][32];     // force B fragments (xj hi / lo, 2 registers each), per lane
  __shared__ float2 s_xj[3][16];          // j coordinates as natural pairs (2t, 2t + 1) per n-tile
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  if (threadIdx.x < 32) {
    for (int m = 0; m < 4; ++m) for (int h = 0; h < 2; ++h) for (int q = 0; q < 4; ++q)
      s_a[m][h][q][lane] = __float_as_uint(0.01f * (lane + m + h + q) + 1.0f);
    for (int n = 0; n < 4; ++n) for (int q = 0; q < 4; ++q) {
      s_b[n][q][lane] = __float_as_uint(0.02f * (lane + n + q) + 0.5f);
      s_bx[n][q][lane] = __float_as_uint(0.03f * (lane + n + q) + 0.25f);
    }
    if (lane < 16) for (int d = 0; d < 3; ++d) s_xj[d][lane] = make_float2(0.1f * lane + d, 0.1f * lane + d + 0.05f);
  }
  __syncthreads();
  float xi[8], yi[8], zi[8];  // i-beads g + 8 k
  for (int k = 0; k < 8; ++k) { xi[k] = 1.0f + g + 8 * k; yi[k] = 2.0f + k; zi[k] = 3.0f + k; }
  float fi[4][4];             // i-side accumulators of the 4 m-tiles (tensor-core variant)
  u64 fxi[8], fyi[8], fzi[8]; // i-side accumulators on the FMA pipe (packed over the two j of a register pair)
  for (int m = 0; m < 4; ++m) for (int q = 0; q < 4; ++q) fi[m][q] = 0.f;
  for (int k = 0; k < 8; ++k) fxi[k] = fyi[k] = fzi[k] = pk2(0.f, 0.f);
  u64 ev2 = pk2(0.f, 0.f);
  const u64 rs2 = pk2(rs, rs);
  float jacc = 0.f;
#pragma unroll 1
  for (int tr = 0; tr < trips; ++tr) {
#pragma unroll 1
    for (int nt = 0; nt < 4; ++nt) {
      unsigned b1[2] = {s_b[nt][0][lane], s_b[nt][1][lane]}, b2[2] = {s_b[nt][2][lane], s_b[nt][3][lane]};
      unsigned bxh[2] = {s_bx[nt][0][lane], s_bx[nt][1][lane]}, bxl[2] = {s_bx[nt][2][lane], s_bx[nt][3][lane]};
      const float2 xj = s_xj[0][4 * nt + t], yj = s_xj[1][4 * nt + t], zj = s_xj[2][4 * nt + t];
      u64 ax = pk2(0.f, 0.f), ay = ax, az = ax;  // j-side partial sums of the lane's two j-beads
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        unsigned a1[4], a2[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { a1[q] = s_a[mt][0][q][lane]; a2[q] = s_a[mt][1][q][lane]; }
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        mma_tf32(c, a1, b1);
        mma_tf32(c, a2, b2);  // c = r^2 of the lane's four pairs of this block
        unsigned fh[4], fl[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {  // rows g (h = 0) and g + 8 (h = 1): one packed pair over the two j each
          const int k = 2 * mt + h;
          const u64 r2 = pk2(fabsf(c[2 * h]) + 1.0f, fabsf(c[2 * h + 1]) + 1.0f);
          float r2a, r2b;
          unpk2(r2, r2a, r2b);
          const u64 r = pk2(fsqrt(r2a), fsqrt(r2b));
          const u64 q = fma2(rs2, r, r2);
          float qa, qb;
          unpk2(q, qa, qb);
          const u64 wr = pk2(frcp(qa), frcp(qb));
          const u64 w = mul2(r, wr);
          const u64 w2 = mul2(w, w);
          const u64 w3 = mul2(w2, w);
          const u64 wp = mul2(w3, w3);
          const u64 fs = mul2(wp, wr);
          ev2 = add2(ev2, wp);
          // deltas for the j side (scalar subtractions: xi is not a register pair)
          const u64 dx = pk2(xi[k] - xj.x, xi[k] - xj.y), dy = pk2(yi[k] - yj.x, yi[k] - yj.y), dz = pk2(zi[k] - zj.x, zi[k] - zj.y);
          ax = fma2(fs, dx, ax); ay = fma2(fs, dy, ay); az = fma2(fs, dz, az);
          if (I_SIDE_MMA) {
            // fs = hi + lo in TF32 parts: hi by masking (ALU pipe), lo = fs - hi (one packed subtraction)
            float fa, fb;
            unpk2(fs, fa, fb);
            const unsigned ha = __float_as_uint(fa) & 0xffffe000u, hb = __float_as_uint(fb) & 0xffffe000u;
            const u64 lo2 = add2(fs, pk2(-__uint_as_float(ha), -__uint_as_float(hb)));
            float la, lb;
            unpk2(lo2, la, lb);
            // accumulator layout -> A-operand layout: (row h, column 2t) -> a[h], (row h, column 2t + 1) -> a[h + 2]
            fh[h] = ha; fh[h + 2] = hb;
            fl[h] = __float_as_uint(la); fl[h + 2] = __float_as_uint(lb);
          } else {
            fxi[k] = fma2(fs, dx, fxi[k]); fyi[k] = fma2(fs, dy, fyi[k]); fzi[k] = fma2(fs, dz, fzi[k]);
          }
        }
        if (I_SIDE_MMA) {
          mma_tf32(fi[mt], fh, bxh);
          mma_tf32(fi[mt], fl, bxh);
          mma_tf32(fi[mt], fh, bxl);
        }
      }
      float lo, hi;
      unpk2(ax, lo, hi); jacc += lo - hi;
      unpk2(ay, lo, hi); jacc += lo - hi;
      unpk2(az, lo, hi); jacc += lo - hi;
    }
  }
  float s = jacc, lo, hi;
  for (int m = 0; m < 4; ++m) for (int q = 0; q < 4; ++q) s += fi[m][q];
  for (int k = 0; k < 8; ++k) {
    unpk2(fxi[k], lo, hi); s += lo + hi; unpk2(fyi[k], lo, hi); s += lo + hi; unpk2(fzi[k], lo, hi); s += lo + hi;
  }
  unpk2(ev2, lo, hi);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + lo + hi;
}

template <typename F>
float time_ms(F launch) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  return best;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double clk = khz * 1e3;
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 2048 * 4);
  printf("device %s, %d SMs, max clock %.0f MHz (cycles below assume it)\n", p.name, sms, clk / 1e6);
  for (int wps : {4, 8, 16, 32}) {
    const int threads = wps * 32 > 1024 ? 1024 : wps * 32, blocks = sms * (wps * 32 / threads);
    const float ms = time_ms([&] { k_mma<<<blocks, threads>>>(out); });
    const double mma_per_smsp = 8.0 * ITERS * wps / 4.0;
    printf("%-10s %3d warps/SM %8.3f cycles per mma.sync.m16n8k8.tf32 per SMSP (%.1f TFLOP/s)\n", "mma", wps,
           ms * 1e-3 * clk / mma_per_smsp, 8.0 * ITERS * wps * sms * 2048.0 / (ms * 1e-3) / 1e12);
  }
  for (int per_sm : {1, 2, 3}) {
    const int threads = 256, blocks = sms * per_sm, trips = 2048;
    for (int variant = 0; variant < 3; ++variant) {
      const float ms = time_ms([&] {
        if (variant == 0) k_hot<<<blocks, threads>>>(out, 0.05f, trips);
        else if (variant == 1) k_hot_mma<false><<<blocks, threads>>>(out, 0.05f, trips);
        else k_hot_mma<true><<<blocks, threads>>>(out, 0.05f, trips);
      });
      const double cyc_pair = ms * 1e-3 * clk / (64.0 * trips * (per_sm * 8 / 4.0));
      printf("%-22s %2d warps/SM %8.3f cycles per warp-pair per SMSP; %.3e pairs/s\n",
             variant == 0 ? "hot (shipped body)" : variant == 1 ? "hot, r^2 by MMA" : "hot, r^2 + i side MMA",
             per_sm * 8, cyc_pair, 64.0 * trips * (double)blocks * threads / (ms * 1e-3));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
