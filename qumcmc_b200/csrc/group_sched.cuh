// L2-resident group schedule of the two-group sweep paths (sweep.cu: complex64, sweep128.cu: complex128), n <= 20.
//
// A proposal is r sweeps over one statevector, alternately on the LO and the HI qubit group.  The step-major schedule
// launches sweep k of EVERY chain, then sweep k + 1: each launch streams all statevectors through HBM (r - 1 read +
// write passes per proposal).  Here `group` chains whose statevectors fit the L2 together run ALL their sweeps back to
// back, then the next group starts: a sweep finds the lines its predecessor left dirty in the L2 and rewrites them in
// place, so HBM sees about one write-back per chain and proposal.  order[] is sorted by r inside a parity class, so a
// group holds one or two r values.  Groups alternate over a few side streams so that the tail of one group's launch is
// filled by CTAs of another's; the streams fork from and join the caller's stream through events.
// The kernels, their arguments and the per-chain arithmetic are those of the step-major schedule: results are bit-identical.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

enum { QMC_GS_FIRST = 0, QMC_GS_MID = 1, QMC_GS_FINAL = 2, QMC_GS_FINAL1 = 3, QMC_GS_HI_FIRST = 4, QMC_GS_HI_MID = 5 };

struct QmcGroupLaunch {
    int kind;    // QMC_GS_*: role of the sweep
    int off;     // first position in order[]
    int count;   // chains
    int stream;  // side stream index
};

struct QmcGroupSched {
    std::vector<cudaStream_t> aux;
    std::vector<cudaEvent_t> ev_join;
    cudaEvent_t ev_fork = nullptr;
    std::vector<int> rpos;  // r of the chain at position i of order[]
    std::vector<QmcGroupLaunch> plan;  // launch list of the current hop

    int ensure(int count) {
        if (!ev_fork) QMC_CUDA_TRY(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        while ((int)aux.size() < count) {
            cudaStream_t q = nullptr;
            cudaEvent_t e = nullptr;
            QMC_CUDA_TRY(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
            aux.push_back(q);
            QMC_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ev_join.push_back(e);
        }
        return QMC_OK;
    }
    void destroy() {
        for (auto e : ev_join) cudaEventDestroy(e);
        if (ev_fork) cudaEventDestroy(ev_fork);
        for (auto q : aux) cudaStreamDestroy(q);
        ev_join.clear();
        aux.clear();
        ev_fork = nullptr;
    }
};

// Knobs (read per call: the parity tests switch schedules inside one process).  QUMCMC_L2_GROUP_MIB = statevector bytes
// per group (0 = step-major schedule), QUMCMC_L2_STREAMS = side streams.  Defaults measured on B200
// (profiles/r2_l2_group_sweep.txt): about 64 MiB of statevectors in flight is what stays resident (C3, complex64: 16 MiB x 4
// streams 46.7 k proposals/s, 24 MiB x 3 streams 46.4 k, 16 MiB x 5 = 80 MiB in flight 42.4 k).  The schedule costs one
// kernel launch per group and sweep -- 10.6 k launches per hop of 4096 chains at 24 MiB x 3 (15.9 k at 16 MiB x 4), about
// 3 us of one host core each -- so the default takes the larger group: 24 MiB x 3 streams for complex64, 16 MiB x 3 for
// complex128 (one 2^20 statevector per group: 18.2 k proposals/s vs 15.0 k step-major).
struct QmcGroupKnobs {
    int group = 0;    // chains per group (0: off)
    int streams = 1;
    bool prefetch = false;
    bool debug = false;
};
inline QmcGroupKnobs qmc_group_knobs(int64_t state_bytes, int default_mib, int default_streams) {
    QmcGroupKnobs k;
    const int mib = getenv("QUMCMC_L2_GROUP_MIB") ? atoi(getenv("QUMCMC_L2_GROUP_MIB")) : default_mib;
    const int str = getenv("QUMCMC_L2_STREAMS") ? atoi(getenv("QUMCMC_L2_STREAMS")) : default_streams;
    k.streams = std::min(8, std::max(1, str));
    if (mib > 0) k.group = (int)std::max<int64_t>(1, std::min<int64_t>(4096, ((int64_t)mib << 20) / state_bytes));
    k.prefetch = getenv("QUMCMC_L2_PF") ? atoi(getenv("QUMCMC_L2_PF")) != 0 : false;
    k.debug = getenv("QUMCMC_L2_DEBUG") != nullptr;
    return k;
}

// The launch list of one hop (pure host logic: qmc_debug_group_plan exposes it to the CPU tests).
// rhist[r] = chains with r layers; order[] = odd-r chains (r descending), then even-r chains (r descending).  A chain with
// r layers takes r sweeps, alternately on the LO and the HI group, the last one always LO: odd r starts with LO.
// Launches of one stream execute in list order; a group lives on one stream, so every chain sees its sweeps in order.
inline void qmc_group_plan(int nc, int r_max, const std::vector<int> &rhist, int group, int streams, std::vector<int> &rpos,
                           std::vector<QmcGroupLaunch> &out) {
    out.clear();
    rpos.resize((size_t)nc);
    int base[2] = {0, 0};
    {
        int k = 0;
        for (int parity = 1; parity >= 0; --parity) {
            if (parity == 0) base[1] = k;
            for (int r = r_max; r >= 1; --r)
                if ((r & 1) == parity)
                    for (int i = 0; i < rhist[r]; ++i) rpos[k++] = r;
        }
    }
    int gi = 0;
    for (int cls = 0; cls < 2; ++cls) {  // cls 0: r odd (LO sweeps at odd steps), cls 1: r even (LO sweeps at even steps)
        const int end = cls == 0 ? base[1] : nc;
        for (int g0 = base[cls]; g0 < end; g0 += group, ++gi) {
            const int cnt = std::min(group, end - g0);
            const int sg = gi % streams;
            const int rg = rpos[g0];
            int gt = cnt;  // chains of the group with r > step: a prefix, r descends
            for (int step = 1; step <= rg; ++step) {
                int eq = 0;
                while (gt > 0 && rpos[g0 + gt - 1] <= step) { --gt; ++eq; }
                const bool lo_step = ((step & 1) == 1) == (cls == 0);
                if (lo_step) {
                    if (gt > 0) out.push_back({step == 1 ? QMC_GS_FIRST : QMC_GS_MID, g0, gt, sg});
                    if (eq > 0) out.push_back({step == 1 ? QMC_GS_FINAL1 : QMC_GS_FINAL, g0 + gt, eq, sg});
                } else if (gt > 0) {
                    out.push_back({step == 1 ? QMC_GS_HI_FIRST : QMC_GS_HI_MID, g0, gt, sg});
                }
            }
        }
    }
}

// lo(kind, off, count, stream) / hi(first, off, count, stream) launch one sweep over order[off .. off + count).
template <typename LaunchLo, typename LaunchHi>
int qmc_group_schedule(QmcGroupSched &gs, const QmcGroupKnobs &kn, cudaStream_t st, int nc, int r_max,
                       const std::vector<int> &rhist, LaunchLo lo, LaunchHi hi, int &launched) {
    int rc = gs.ensure(kn.streams);
    if (rc != QMC_OK) return rc;
    QMC_CUDA_TRY(cudaEventRecord(gs.ev_fork, st));
    for (int i = 0; i < kn.streams; ++i) QMC_CUDA_TRY(cudaStreamWaitEvent(gs.aux[i], gs.ev_fork, 0));
    qmc_group_plan(nc, r_max, rhist, kn.group, kn.streams, gs.rpos, gs.plan);
    for (const QmcGroupLaunch &L : gs.plan) {
        rc = L.kind >= QMC_GS_HI_FIRST ? hi(L.kind == QMC_GS_HI_FIRST, L.off, L.count, gs.aux[L.stream])
                                       : lo(L.kind, L.off, L.count, gs.aux[L.stream]);
        if (rc != QMC_OK) return rc;
        ++launched;
    }
    for (int i = 0; i < kn.streams; ++i) {
        QMC_CUDA_TRY(cudaEventRecord(gs.ev_join[i], gs.aux[i]));
        QMC_CUDA_TRY(cudaStreamWaitEvent(st, gs.ev_join[i], 0));
    }
    return QMC_OK;
}
