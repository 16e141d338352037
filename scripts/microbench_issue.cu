// microbench_issue.cu — what bounds the Newton-3 pair kernel's hot loop on sm_100a?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/microbench_issue scripts/microbench_issue.cu
// Measures, per SM sub-partition and in cycles per warp-instruction, at several occupancies:
//   ffma2        : packed FFMA2 only (independent chains)
//   ffma2+alu    : FFMA2 interleaved 1:1 with independent integer adds — does a packed instruction hold
//                  the issue port for both of its pipe cycles (then 3 cycles per pair of instructions)
//                  or only the FMA pipe (then 2)?
//   ffma2+mufu   : 38 FFMA2 + 8 MUFU per trip (the hot loop's ratio 9.5 : 2 per pair, x4), independent chains:
//                  do the FMA and the MUFU pipe overlap when nothing depends on anything?
//   hot          : the hot body itself (EV power 6, far tile, no CHB), register resident, j from shared
//   hot rsq+rcp  : the same with rsqrt + rcp(r + rs) instead of sqrt + rcp(r^2 + rs r)
//   hot sqrt+rcp/2 : one MUFU.RCP per TWO pairs (t = 1 / (qa qb); 1/qa = t qb; 1/qb = t qa): 1.5 MUFU per pair
//   hot rsq+ser4/2 : far tiles only need 1 MUFU per pair: w = y / (1 + rs y), y = rsqrt(r^2), as a series in rs y
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 x, u64 y, u64 z) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(y), "l"(z)); return d; }
__device__ __forceinline__ u64 mul2(u64 x, u64 y) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(x), "l"(y)); return d; }
__device__ __forceinline__ u64 add2(u64 x, u64 y) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(x), "l"(y)); return d; }
__device__ __forceinline__ float fsqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float frcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float frsq(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr int ITERS = 2048;

__global__ void k_ffma2(float* out, float a, float b) {
  u64 v[8];
  for (int q = 0; q < 8; ++q) v[q] = pk2(threadIdx.x + q, threadIdx.x - q);
  const u64 A = pk2(a, a), B = pk2(b, b);
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = fma2(v[q], A, B);
  }
  float s = 0, lo, hi;
  for (int q = 0; q < 8; ++q) { unpk2(v[q], lo, hi); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// three distinct 64-bit register operands per FFMA2 (6 register reads)
__global__ void k_ffma2_3reg(float* out, float a, float b) {
  u64 v[8], w[8];
  for (int q = 0; q < 8; ++q) { v[q] = pk2(threadIdx.x + q, threadIdx.x - q); w[q] = pk2(a + q, b - q); }
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = fma2(v[q], w[q], w[(q + 3) & 7]);
  }
  float s = 0, lo, hi;
  for (int q = 0; q < 8; ++q) { unpk2(v[q], lo, hi); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2_alu(float* out, float a, float b, int c) {
  u64 v[8];
  int w[8];
  for (int q = 0; q < 8; ++q) { v[q] = pk2(threadIdx.x + q, threadIdx.x - q); w[q] = threadIdx.x + q; }
  const u64 A = pk2(a, a), B = pk2(b, b);
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        v[q] = fma2(v[q], A, B);
        asm volatile("xor.b32 %0, %0, %1;" : "+r"(w[q]) : "r"(c));  // LOP3 on the ALU pipe, independent
      }
  }
  float s = 0, lo, hi;
  for (int q = 0; q < 8; ++q) { unpk2(v[q], lo, hi); s += lo + hi + w[q]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2_mufu(float* out, float a, float b) {
  u64 v[8];
  float m[8];
  for (int q = 0; q < 8; ++q) { v[q] = pk2(threadIdx.x + q, threadIdx.x - q); m[q] = 1.0f + threadIdx.x + q; }
  const u64 A = pk2(a, a), B = pk2(b, b);
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    // 38 FFMA2 + 8 MUFU: the hot loop's mix for 4 pairs (9.5 packed + 2 MUFU each), all independent
#pragma unroll
    for (int q = 0; q < 8; ++q) { v[q] = fma2(v[q], A, B); asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(m[q])); }
#pragma unroll
    for (int u = 0; u < 3; ++u)
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = fma2(v[q], A, B);
#pragma unroll
    for (int q = 0; q < 6; ++q) v[q] = fma2(v[q], A, B);
  }
  float s = 0, lo, hi;
  for (int q = 0; q < 8; ++q) { unpk2(v[q], lo, hi); s += lo + hi + m[q]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the hot body: 8 i-beads (4 packed pairs) against the j-beads of a 32-bead tile in shared memory, 64 pairs per trip
template <int VARIANT>
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
          u64 fs, wp;
          if (VARIANT == 0) {
            const u64 r = pk2(fsqrt(r2a), fsqrt(r2b));
            const u64 q = fma2(rs2, r, r2);
            float qa, qb;
            unpk2(q, qa, qb);
            const u64 wr = pk2(frcp(qa), frcp(qb));
            const u64 w = mul2(r, wr);
            const u64 w2 = mul2(w, w);
            const u64 w3 = mul2(w2, w);
            wp = mul2(w3, w3);
            fs = mul2(wp, wr);
          } else if (VARIANT == 2) {
            // one MUFU.RCP for the two halves: t = 1 / (qa qb); 1/qa = t qb, 1/qb = t qa
            const u64 r = pk2(fsqrt(r2a), fsqrt(r2b));
            const u64 q = fma2(rs2, r, r2);
            float qa, qb;
            unpk2(q, qa, qb);
            const float t = frcp(qa * qb);
            const u64 wr = pk2(t * qb, t * qa);
            const u64 w = mul2(r, wr);
            const u64 w2 = mul2(w, w);
            const u64 w3 = mul2(w2, w);
            wp = mul2(w3, w3);
            fs = mul2(wp, wr);
          } else if (VARIANT == 3 || VARIANT == 4) {
            // far tiles (r >= 0.9 nm, rs / r <= 0.056): w = y / (1 + rs y) as a truncated series in t = rs y
            const u64 y = pk2(frsq(r2a), frsq(r2b));
            const u64 t = mul2(rs2, y);
            const u64 one = pk2(1.0f, 1.0f), mone = pk2(-1.0f, -1.0f);
            u64 P;
            if (VARIANT == 3) P = fma2(t, fma2(t, fma2(t, fma2(t, one, mone), one), mone), one);  // 1 - t + t^2 - t^3 + t^4
            else P = fma2(t, fma2(t, one, mone), one);                                            // 1 - t + t^2
            const u64 w = mul2(y, P);
            const u64 w2 = mul2(w, w);
            const u64 w3 = mul2(w2, w);
            wp = mul2(w3, w3);
            fs = mul2(mul2(wp, w), y);
          } else {
            const u64 y = pk2(frsq(r2a), frsq(r2b));
            const u64 r = fma2(r2, y, rs2);  // r + rs
            float qa, qb;
            unpk2(r, qa, qb);
            const u64 w = pk2(frcp(qa), frcp(qb));
            const u64 w2 = mul2(w, w);
            const u64 w3 = mul2(w2, w);
            wp = mul2(w3, w3);
            fs = mul2(mul2(wp, w), y);
          }
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

// The hot body with the four register pairs of a j-bead processed PHASE BY PHASE (all r^2, then all
// square roots, then all reciprocals, ...) instead of pair by pair: does the source order change what
// ptxas's schedule achieves?
__global__ void k_hot_phased(float* out, float rs, int trips) {
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
        u64 dx[4], dy[4], dz[4], r2[4], r[4], wr[4], fs[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          dx[m] = add2(x2[m], njx); dy[m] = add2(y2[m], njy); dz[m] = add2(z2[m], njz);
          r2[m] = fma2(dz[m], dz[m], fma2(dy[m], dy[m], mul2(dx[m], dx[m])));
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          float p, q;
          unpk2(r2[m], p, q);
          r[m] = pk2(fsqrt(p), fsqrt(q));
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const u64 qq = fma2(rs2, r[m], r2[m]);
          float p, q;
          unpk2(qq, p, q);
          wr[m] = pk2(frcp(p), frcp(q));
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const u64 w = mul2(r[m], wr[m]);
          const u64 w2 = mul2(w, w);
          const u64 w3 = mul2(w2, w);
          const u64 wp = mul2(w3, w3);
          ev2 = add2(ev2, wp);
          fs[m] = mul2(wp, wr[m]);
        }
        u64 ax = pk2(0.f, 0.f), ay = ax, az = ax;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          fx[m] = fma2(fs[m], dx[m], fx[m]); fy[m] = fma2(fs[m], dy[m], fy[m]); fz[m] = fma2(fs[m], dz[m], fz[m]);
          ax = fma2(fs[m], dx[m], ax); ay = fma2(fs[m], dy[m], ay); az = fma2(fs[m], dz[m], az);
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
  printf("%-14s %9s %14s %s\n", "kernel", "warps/SM", "cyc/instr/SMSP", "note");
  for (int wps : {4, 8, 16, 32}) {  // warps per SM -> per SMSP = wps / 4
    const int threads = wps * 32 > 1024 ? 1024 : wps * 32, blocks = sms * (wps * 32 / threads);
    auto cyc = [&](float ms, double instr_per_thread) {
      const double warp_instr_per_smsp = instr_per_thread * (double)wps / 4.0;
      return ms * 1e-3 * clk / warp_instr_per_smsp;
    };
    float ms = time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 1.0000001f, 1e-9f); });
    printf("%-14s %9d %14.3f FFMA2, 2 uniform operands\n", "ffma2", wps, cyc(ms, 32.0 * ITERS));
    ms = time_ms([&] { k_ffma2_3reg<<<blocks, threads>>>(out, 1.0000001f, 1e-9f); });
    printf("%-14s %9d %14.3f FFMA2, 3 register-pair operands\n", "ffma2_3reg", wps, cyc(ms, 32.0 * ITERS));
    ms = time_ms([&] { k_ffma2_alu<<<blocks, threads>>>(out, 1.0000001f, 1e-9f, 3); });
    printf("%-14s %9d %14.3f per (FFMA2 + LOP3) pair: 2 = FFMA2 leaves the issue port free, 3 = it holds it\n",
           "ffma2+alu", wps, cyc(ms, 32.0 * ITERS));
    ms = time_ms([&] { k_ffma2_mufu<<<blocks, threads>>>(out, 1.0000001f, 1e-9f); });
    printf("%-14s %9d %14.3f per 4 pairs' worth (38 FFMA2 + 8 MUFU, independent): 76 = FMA-pipe bound, 64 = MUFU bound, 140 = no overlap\n",
           "ffma2+mufu", wps, cyc(ms, 1.0 * ITERS));
  }
  for (int cfg = 0; cfg < 3; ++cfg) {
    const int threads = 256, per_sm = cfg == 0 ? 1 : (cfg == 1 ? 2 : 3), blocks = sms * per_sm, trips = 4096;
    for (int variant = 0; variant < 6; ++variant) {
      float ms = time_ms([&] {
        if (variant == 0) k_hot<0><<<blocks, threads>>>(out, 0.05f, trips);
        else if (variant == 1) k_hot<1><<<blocks, threads>>>(out, 0.05f, trips);
        else if (variant == 2) k_hot<2><<<blocks, threads>>>(out, 0.05f, trips);
        else if (variant == 3) k_hot<3><<<blocks, threads>>>(out, 0.05f, trips);
        else if (variant == 4) k_hot<4><<<blocks, threads>>>(out, 0.05f, trips);
        else k_hot_phased<<<blocks, threads>>>(out, 0.05f, trips);
      });
      const double pairs = 64.0 * trips * (double)blocks * threads;
      const double cyc_pair = ms * 1e-3 * clk / (64.0 * trips * (per_sm * 8 / 4.0));
      printf("%-14s %9d %14.3f cycles per warp-pair per SMSP (floor 19); %.3e pairs/s\n",
             variant == 0 ? "hot sqrt+rcp" : variant == 1 ? "hot rsq+rcp" : variant == 2 ? "hot sqrt+rcp/2" :
             variant == 3 ? "hot rsq+ser4" : variant == 4 ? "hot rsq+ser2" : "hot phased", per_sm * 8, cyc_pair, pairs / (ms * 1e-3));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
