// Microbenchmark (diagnostics): per-sub-partition cost of the phase-layer instruction sequences (XU pipe: MUFU, I2FP).
#include <cstdio>
#include <cuda_runtime.h>
// VAR 0: sin.approx only   1: sin+cos   2: int->float only   3: int->float + sin + cos   4: magic-number conversion + sin + cos
// VAR 5: full phase step (IADD, convert, sincos, complex multiply)   6: same with the magic-number conversion
template <int VAR>
__global__ void __launch_bounds__(128, 3) k(float *out, int seed, int iters) {
    float x[32], y[32];
    int t[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) { x[j] = threadIdx.x + j; y[j] = 1.f; t[j] = seed * (threadIdx.x * 32 + j + 1); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (VAR == 0) { x[j] = __sinf(x[j]); }
            else if (VAR == 1) { float s, c; __sincosf(x[j], &s, &c); x[j] = s; y[j] += c; }
            else if (VAR == 2) { t[j] += seed; x[j] += (float)t[j]; }
            else if (VAR == 3) { t[j] += seed; float s, c; __sincosf((float)t[j] * 1.4629180792671596e-09f, &s, &c); x[j] += s; y[j] += c; }
            else if (VAR == 4) {
                t[j] += seed;
                const float f = __int_as_float(((unsigned)t[j] >> 9) + 0x4B000000u);   // 2^23 + top 23 bits
                float s, c; __sincosf(fmaf(f, 7.4901405658478574e-07f, -6.2831853071795862f * 1.0f * 8388608.0f * 1.1920928955078125e-07f * 8388608.0f), &s, &c);
                x[j] += s; y[j] += c;
            }
            else if (VAR == 5 || VAR == 6) {
                t[j] += seed;
                float ang;
                if (VAR == 5) ang = (float)t[j] * 1.4629180792671596e-09f;
                else ang = fmaf(__int_as_float(((unsigned)t[j] >> 9) + 0x4B000000u), 7.4901405658478574e-07f, -6.2831855f);
                float s, c; __sincosf(ang, &s, &c);
                const float nx = x[j] * c - y[j] * s, ny = y[j] * c + x[j] * s;
                x[j] = nx; y[j] = ny;
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) acc += x[j] + y[j] + (float)t[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
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
    float *out;
    const int grid = 148 * per_sm;
    cudaMalloc(&out, sizeof(float) * grid * 128);
    k<VAR><<<grid, 128>>>(out, 12345, iters);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<VAR><<<grid, 128>>>(out, 12345, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mhz = sm_mhz();
    const double el = (double)iters * 32 * per_sm;  // elements per sub-partition
    printf("%-46s warps/SMSP=%d  %.2f cycles/element/SMSP  (%.3f ms @ %.0f MHz)\n", name, per_sm, ms * 1e3 * mhz / el, ms, mhz);
    cudaFree(out);
}
int main() {
    for (int w = 1; w <= 3; w += 2) {
        run<0>("sin.approx", w, 4000);
        run<1>("sin + cos", w, 4000);
        run<2>("int -> float (I2FP)", w, 4000);
        run<3>("I2FP + sin + cos", w, 4000);
        run<4>("magic-number convert + sin + cos", w, 4000);
        run<5>("phase step: I2FP, sincos, complex multiply", w, 4000);
        run<6>("phase step: magic convert, sincos, cmul", w, 4000);
    }
    return 0;
}
