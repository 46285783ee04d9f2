// sweep.cu -- HBM-resident statevector path (n above the shared-memory limits). Placeholder.
#include "common.cuh"
struct SweepWorkspace { int dummy; };
void qmc_sweep_free(qmc_model *m) { delete m->sweep; m->sweep = nullptr; }
int qmc_sweep_run(qmc_model *m, int, int64_t, int64_t, const uint64_t *, uint64_t, uint32_t, double, double, double,
                  double, uint64_t, int, const double *, const int *, const double *, const int *, uint64_t *,
                  uint8_t *, uint64_t *, int64_t *, int64_t *, void *, void *, cudaStream_t) {
    qmc_set_error("HBM sweep path not built yet (n=%d)", m->n);
    return QMC_ERR_UNSUPPORTED;
}
