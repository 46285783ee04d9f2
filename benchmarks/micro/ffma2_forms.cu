// Microbenchmark (diagnostics): which source forms of the tan-form butterfly ptxas turns into FFMA2 with operand
// modifiers (half swap / per-half negate / scalar broadcast), and what each costs per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long A64;
__device__ __forceinline__ A64 pk(float x, float y) { A64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ float2 upk(A64 r) { float2 o; asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r)); return o; }
__device__ __forceinline__ A64 fma2(A64 a, A64 b, A64 c) { A64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// VAR 0: parity-swapped storage, t from kernel parameter        v[j] += (t,-t) * v[p] ; v[p] += (-t,t) * v[j]
// VAR 1: natural storage, half swap in source, t kernel param   a' = a + t * (b.y, -b.x) ; b' = b + t * (a.y, -a.x)
// VAR 2: VAR 0 with t read from shared memory per stage (value uniform, register not provably so)
// VAR 3: VAR 1 with t read from shared memory per stage
// VAR 4: VAR 0 with a per-thread sign on t (what the parity-swapped layout needs in the sweep kernels)
// VAR 5: odd-parity amplitudes stored as -i*x: skew update v[j] += s v[p]; v[p] -= s v[j], s per thread (scalar broadcast)
// VAR 6: VAR 5 with s = kernel parameter
// VAR 7: VAR 5 layout, one multiplier per CTA read from __constant__ memory at a block-uniform index
__constant__ float c_tab[4096];

template <int VAR>
__global__ void __launch_bounds__(128, 3) k(float *out, long long *cyc, float t, int iters) {
    __shared__ float ts[64];
    if (threadIdx.x < 64) ts[threadIdx.x] = t + 1e-3f * threadIdx.x;
    __syncthreads();
    const float sgn = (__popc(threadIdx.x) & 1) ? -1.f : 1.f;
    A64 v[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = pk(threadIdx.x + j, 1.f);
    long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            float tt = t;
            if (VAR == 2 || VAR == 3) tt = ts[(it + b) & 63];
            if (VAR == 4 || VAR == 5) tt = ts[(it + b) & 63] * sgn;
            if (VAR == 7) tt = c_tab[(blockIdx.x + iters + b) & 4095];
            const A64 TA = pk(tt, -tt), TB = pk(-tt, tt), TT = pk(tt, tt), TN = pk(-tt, -tt);
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                if (j & (1 << b)) continue;
                const int p = j | (1 << b);
                const A64 u = v[j], w = v[p];
                if (VAR == 0 || VAR == 2 || VAR == 4) {
                    v[j] = fma2(TA, w, u);
                    v[p] = fma2(TB, u, w);
                } else if (VAR == 5 || VAR == 6 || VAR == 7) {
                    const bool odd = (__builtin_popcount(j) & 1) != 0;
                    v[j] = fma2(odd ? TN : TT, w, u);
                    v[p] = fma2(odd ? TT : TN, u, w);
                } else {
                    const float2 uf = upk(u), wf = upk(w);
                    v[j] = fma2(TT, pk(wf.y, -wf.x), u);
                    v[p] = fma2(TT, pk(uf.y, -uf.x), w);
                }
            }
        }
    }
    long long c1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) { const float2 f = upk(v[j]); acc += f.x + f.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}
__global__ void spin(long long n, long long *o) {
    const long long c0 = clock64();
    while (clock64() - c0 < n) {}
    if (threadIdx.x == 0) *o = clock64() - c0;
}
static double sm_mhz() {
    long long *o; cudaMalloc(&o, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); spin<<<1, 32>>>(4000000, o); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h; cudaMemcpy(&h, o, 8, cudaMemcpyDeviceToHost); cudaFree(o);
    return h / (ms * 1e3);
}
template <int VAR>
void run(const char *name, int per_sm, int iters) {
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
    const double bf = (double)iters * 6 * 32 * per_sm;
    printf("%-52s warps/SMSP=%d  %.2f cycles/butterfly/SMSP  (%.3f ms @ %.0f MHz)\n", name, per_sm, ms * 1e3 * mhz / bf, ms, mhz);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    { float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 0.3f + 1e-5f * i; cudaMemcpyToSymbol(c_tab, h, sizeof(h)); }
    for (int w = 1; w <= 3; w += 2) {
        run<0>("swapped storage, t = kernel param", w, 4000);
        run<1>("natural storage + half swap, t = kernel param", w, 4000);
        run<2>("swapped storage, t from smem", w, 4000);
        run<3>("natural storage + half swap, t from smem", w, 4000);
        run<4>("swapped storage, t from smem * per-thread sign", w, 4000);
        run<5>("-i storage (skew form), s from smem * per-thread sign", w, 4000);
        run<6>("-i storage (skew form), s = kernel param", w, 4000);
        run<7>("skew form, s = __constant__[block-uniform index]", w, 4000);
    }
    return 0;
}
