// probe_sweep.cu -- kernel-level attribution harness for the LO MID sweep (diagnostics; not part of the library).
// Build:  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr \
//              -DQMC_PROBES -I qumcmc_b200/csrc -o benchmarks/probe_sweep benchmarks/probe_sweep.cu qumcmc_b200/csrc/model.o ...
// It includes sweep.cu directly to reach the kernels in its anonymous namespace.
#include "../qumcmc_b200/csrc/sweep.cu"

#include <cstdio>
#include <vector>

template <int MODE, int MINB, int TPC = 1>
static float run_probe(const SweepArgs &a, const int *order, int tiles, int chains, int smem_bytes, int reps) {
    cudaFuncSetAttribute(lo_probe_kernel<MODE, MINB, TPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    lo_probe_kernel<MODE, MINB, TPC><<<dim3(tiles / TPC, chains), NT, smem_bytes>>>(a, order, 0);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) lo_probe_kernel<MODE, MINB, TPC><<<dim3(tiles / TPC, chains), NT, smem_bytes>>>(a, order, 0);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(err));
    return ms / reps;
}

int main(int argc, char **argv) {
    const int n = 20, chains = argc > 1 ? atoi(argv[1]) : 512, tiles = 1 << (n - 13), reps = 3;
    SweepArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n;
    a.n_total = n;
    a.pf_dist = argc > 2 ? atoi(argv[2]) : 222;
    cudaMalloc(&a.state, sizeof(float2) * ((size_t)chains << n));
    cudaMemset(a.state, 0, sizeof(float2) * ((size_t)chains << n));
    std::vector<ChainParams> hp(chains);
    for (int c = 0; c < chains; ++c) {
        ChainParams &p = hp[c];
        memset(&p, 0, sizeof(p));
        p.r = 5; p.cfx = 123456789; p.shift = 20; p.scale = 0.5f; p.active = 1;
        for (int q = 0; q < n; ++q) {
            p.tq[q] = 0.3f + 0.01f * q;
        }
    }
    cudaMalloc(&a.params, sizeof(ChainParams) * chains);
    cudaMemcpy(a.params, hp.data(), sizeof(ChainParams) * chains, cudaMemcpyHostToDevice);
    int4 *af;
    cudaMalloc(&af, sizeof(int4) * 2 * ((size_t)1 << (n - 6)));
    cudaMemset(af, 1, sizeof(int4) * 2 * ((size_t)1 << (n - 6)));
    a.af_lo = af;
    for (int i = 0; i < 64; ++i) a.dr_lo[i] = 1000 + i;
    std::vector<int> ho(chains);
    for (int c = 0; c < chains; ++c) ho[c] = c;
    int *order;
    cudaMalloc(&order, sizeof(int) * chains);
    cudaMemcpy(order, ho.data(), sizeof(int) * chains, cudaMemcpyHostToDevice);
    const double gb = 2.0 * 8.0 * (double)((size_t)chains << n) / 1e9;  // read + write of every amplitude
    const double tiles_per_sm = (double)tiles * chains / 148.0;
    auto report = [&](const char *name, float ms) {
        printf("%-44s %8.3f ms  %7.1f GB/s-equivalent  %7.0f cycles/tile/SM @1.965GHz\n", name, ms, gb / (ms * 1e-3),
               ms * 1e-3 * 1.965e9 / tiles_per_sm);
    };
    const int S = TILE * 8;
    report("full (loads+bfly+phase+transposes), 3 CTA/SM", run_probe<15, 3>(a, order, tiles, chains, S, reps));
    report("full, 4 tiles per CTA", run_probe<15, 3, 4>(a, order, tiles, chains, S, reps));
    report("movement only (1|8), 3 CTA/SM", run_probe<9, 3>(a, order, tiles, chains, S, reps));
    report("compute only (2|4|8), 3 CTA/SM", run_probe<14, 3>(a, order, tiles, chains, S, reps));
    report("compute only, 4 tiles per CTA", run_probe<14, 3, 4>(a, order, tiles, chains, S, reps));
    report("compute only, 16 tiles per CTA", run_probe<14, 3, 16>(a, order, tiles, chains, S, reps));
    report("bfly only (2), 3 CTA/SM", run_probe<2, 3>(a, order, tiles, chains, S, reps));
    report("bfly only, 4 tiles per CTA", run_probe<2, 3, 4>(a, order, tiles, chains, S, reps));
    report("bfly only, 16 tiles per CTA", run_probe<2, 3, 16>(a, order, tiles, chains, S, reps));
    report("bfly + transposes (2|8), 3 CTA/SM", run_probe<10, 3>(a, order, tiles, chains, S, reps));
    report("bfly + phase (2|4), 3 CTA/SM", run_probe<6, 3>(a, order, tiles, chains, S, reps));
    report("phase only (4), 3 CTA/SM", run_probe<4, 3>(a, order, tiles, chains, S, reps));
    report("phase only, 16 tiles per CTA", run_probe<4, 3, 16>(a, order, tiles, chains, S, reps));
    report("transposes only (8), 3 CTA/SM", run_probe<8, 3>(a, order, tiles, chains, S, reps));
    return 0;
}
