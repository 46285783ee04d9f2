// Microbenchmark (diagnostics): issue rate of the butterfly inner loop as FFMA2 (fma.rn.f32x2) vs scalar FFMA,
// per SM sub-partition, with 1..3 resident warps per sub-partition.  Prints cycles per butterfly (= 4 FMA lanes-ops).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long A64;
__device__ __forceinline__ A64 pk(float x, float y) { A64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ float2 upk(A64 r) { float2 o; asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r)); return o; }
__device__ __forceinline__ A64 fma2(A64 a, A64 b, A64 c) { A64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int VAR>
__global__ void __launch_bounds__(128, 3) k(float *out, long long *cyc, float t, int iters) {
    __shared__ float2 tp[40][2];
    if (threadIdx.x < 40) { tp[threadIdx.x][0] = make_float2(t, -t); tp[threadIdx.x][1] = make_float2(-t, t); }
    __syncthreads();
    const int tpar = __popc(threadIdx.x) & 1;
    A64 v[64];
    float re[64], im[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) { v[j] = pk(threadIdx.x + j, 1.f); re[j] = threadIdx.x + j; im[j] = 1.f; }
    const A64 TA = pk(t, -t), TB = pk(-t, t);
    long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int b = 0; b < 6; ++b) {
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                if (j & (1 << b)) continue;
                const int p = j | (1 << b);
                if (VAR >= 3) continue;
                if (VAR == 0) {  // packed butterfly as in the sweep kernels
                    const A64 u = v[j], w = v[p];
                    v[j] = fma2(TA, w, u);
                    v[p] = fma2(TB, u, w);
                } else if (VAR == 1) {  // scalar butterfly: a' = a - i t b, b' = b - i t a
                    const float ur = re[j], ui = im[j], wr = re[p], wi = im[p];
                    re[j] = fmaf(t, wi, ur);
                    im[j] = fmaf(-t, wr, ui);
                    re[p] = fmaf(t, ui, wr);
                    im[p] = fmaf(-t, ur, wi);
                } else if (VAR == 2) {  // packed, only the first half of each butterfly (pure FFMA2 stream, in place)
                    v[j] = fma2(TA, v[p], v[j]);
                }
            }
        }
        if (VAR == 3 || VAR == 4) {  // as in the sweep kernels: multipliers from shared memory per stage, parity-swapped
#pragma unroll
            for (int rep = 0; rep < (VAR == 4 ? 4 : 1); ++rep)
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                const A64 *pp = reinterpret_cast<const A64 *>(&tp[rep * 6 + b][0]);
                const A64 SA = pp[tpar], SB = pp[tpar ^ 1];
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                    if (j & (1 << b)) continue;
                    const bool odd = (__builtin_popcount(j) & 1) != 0;
                    const A64 u = v[j], w = v[j | (1 << b)];
                    v[j] = fma2(odd ? SB : SA, w, u);
                    v[j | (1 << b)] = fma2(odd ? SA : SB, u, w);
                }
            }
        }
    }
    long long c1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) { const float2 f = upk(v[j]); acc += f.x + f.y + re[j] + im[j]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

__global__ void spin(long long n, long long *o) {
    const long long c0 = clock64();
    while (clock64() - c0 < n) {}
    if (threadIdx.x == 0) *o = clock64() - c0;
}
static double sm_mhz() {  // SM clock right now: spin for a known number of SM cycles, time it with events
    long long *o; cudaMalloc(&o, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); spin<<<1, 32>>>(4000000, o); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h; cudaMemcpy(&h, o, 8, cudaMemcpyDeviceToHost); cudaFree(o);
    return h / (ms * 1e3);
}

template <int VAR>
void run(const char *name, int want, int iters) {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<VAR>, 128, 0);
    const int per_sm = want < occ ? want : occ;
    float *out; long long *cyc;
    const int grid = 148 * per_sm;
    cudaMalloc(&out, sizeof(float) * grid * 128);
    cudaMalloc(&cyc, sizeof(long long) * grid);
    k<VAR><<<grid, 128>>>(out, cyc, 0.3f, iters);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<VAR><<<grid, 128>>>(out, cyc, 0.3f, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mhz = sm_mhz();
    long long h[148 * 4];
    cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < grid; ++i) mean += h[i]; mean /= grid;
    const double bf = (double)iters * 6 * 32 * per_sm * (VAR == 4 ? 4 : 1);  // butterflies per sub-partition (one warp of each CTA)
    printf("%-34s warps/SMSP=%d (occ %d)  cycles/butterfly/SMSP: %.2f by clock64, %.2f by events @ %.0f MHz  (%.3f ms)\n", name,
           per_sm, occ, mean / bf, ms * 1e3 * mhz / bf, mhz, ms);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w = 1; w <= 3; ++w) {
        run<0>("FFMA2 butterfly (2 FFMA2)", w, 4000);
        run<1>("scalar butterfly (4 FFMA)", w, 4000);
        run<2>("FFMA2 half butterfly (1 FFMA2)", w, 4000);
        run<3>("FFMA2 bfly, T from smem, parity", w, 4000);
        run<4>("same, 24 straight-line stages", w, 1000);
    }
    return 0;
}
