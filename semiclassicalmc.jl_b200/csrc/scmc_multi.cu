// scmc_multi.cu -- several GPUs of one box behind ONE C-ABI object (include/scmc.h, scmc_multi_*): what a non-Python caller (the Julia
// shim, a C program) uses instead of the reference's one-Slurm-job-step-per-temperature farm (src/Launcher.jl:239-345, merged from files
// by src/IO.jl:64-143).
//
// Replicas / temperatures / restarts never interact, so a multi object is one scmc_t per device, each owning a contiguous range of GLOBAL
// replica ids (the ids enter the counter-based stream: a chain does not depend on how replicas are split).  Calls fan out to the devices
// from one host thread per device; nothing crosses between GPUs while sweeping.  The only exchange of the path -- the per-replica measure!
// rows at the end of a measurement interval -- is ONE ncclAllGather over NVLink (single process, ncclCommInitAll), after which every
// device holds the rows of all replicas and device 0 hands them to the caller.  NCCL is loaded at first use (dlopen), so the library keeps
// loading on machines without it.
#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "handle.h"

namespace {

struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& err) {
        if (lib) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (lib) break;
        }
        if (!lib) { err = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : ""); return false; }
        auto sym = [&](const char* n) { return dlsym(lib, n); };
        CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        AllGather = (decltype(AllGather))sym("ncclAllGather");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        if (!CommInitAll || !CommDestroy || !AllGather || !GroupStart || !GroupEnd) { err = "NCCL library lacks a required symbol"; return false; }
        return true;
    }
};
Nccl g_nccl;

}  // namespace

struct scmc_multi {
    int n_dev = 0;
    int64_t R_total = 0, per = 0;                     // per: padded shard size (ncclAllGather needs equal counts)
    std::vector<int> dev;
    std::vector<scmc_t*> h;
    std::vector<int64_t> first, count;                // replica range of every device (local numbering of this object)
    std::vector<ncclComm_t> comm;
    std::vector<double*> d_send, d_recv;              // [per][NOBS], [n_dev][per][NOBS] on every device
    bool nccl_ready = false;
};

using namespace scmc;

namespace {

// f(i) on one host thread per device; the first failure (status, message) is reported in the caller's thread
template <class F>
int32_t fan_out(scmc_multi* m, F&& f) {
    std::vector<int32_t> st(m->n_dev, SCMC_OK);
    std::vector<std::string> msg(m->n_dev);
    std::vector<std::thread> th;
    for (int i = 0; i < m->n_dev; i++)
        th.emplace_back([&, i] {
            st[i] = f(i);
            if (st[i] != SCMC_OK) msg[i] = scmc_last_error();
        });
    for (auto& t : th) t.join();
    for (int i = 0; i < m->n_dev; i++)
        if (st[i] != SCMC_OK) { set_error("device " + std::to_string(m->dev[i]) + ": " + msg[i]); return st[i]; }
    return SCMC_OK;
}

int32_t nccl_fail(ncclResult_t r, const char* what) {
    set_error(std::string("NCCL error in ") + what + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    return SCMC_ERR_CUDA;
}

int32_t ensure_nccl(scmc_multi* m) {
    if (m->nccl_ready) return SCMC_OK;
    std::string err;
    if (!g_nccl.load(err)) { set_error(err); return SCMC_ERR_UNSUPPORTED; }
    m->comm.assign(m->n_dev, nullptr);
    ncclResult_t r = g_nccl.CommInitAll(m->comm.data(), m->n_dev, m->dev.data());
    if (r != ncclSuccess) return nccl_fail(r, "ncclCommInitAll");
    m->d_send.assign(m->n_dev, nullptr); m->d_recv.assign(m->n_dev, nullptr);
    for (int i = 0; i < m->n_dev; i++) {
        SCMC_CK(cudaSetDevice(m->dev[i]));
        SCMC_CK(cudaMalloc(&m->d_send[i], (size_t)m->per * SCMC_NOBS * 8));
        SCMC_CK(cudaMemset(m->d_send[i], 0, (size_t)m->per * SCMC_NOBS * 8));
        SCMC_CK(cudaMalloc(&m->d_recv[i], (size_t)m->n_dev * m->per * SCMC_NOBS * 8));
    }
    m->nccl_ready = true;
    return SCMC_OK;
}

}  // namespace

extern "C" {

int32_t scmc_multi_create(scmc_multi_t** out, int32_t n_dev, const int32_t* devices, int32_t precision, int64_t n_sites, int32_t d,
                          int64_t n_replicas_total, int64_t replica_offset, int32_t z_max, const int32_t* nbr_site, const int32_t* nbr_label,
                          int32_t n_labels, const double* J, const double* K, const double* gen_re, const double* gen_im,
                          const double* gensq_re, const double* gensq_im, const int32_t* colour) {
    SCMC_ARG(out != nullptr, "scmc_multi_create: null pointer");
    *out = nullptr;
    SCMC_ARG(n_dev >= 1 && n_dev <= 64, "scmc_multi_create: 1..64 devices");
    SCMC_ARG(n_replicas_total >= n_dev, "scmc_multi_create: fewer replicas than devices");
    scmc_multi* m = new scmc_multi();
    m->n_dev = n_dev; m->R_total = n_replicas_total; m->per = (n_replicas_total + n_dev - 1) / n_dev;
    m->dev.resize(n_dev); m->h.assign(n_dev, nullptr); m->first.resize(n_dev); m->count.resize(n_dev);
    const int64_t base = n_replicas_total / n_dev, rem = n_replicas_total % n_dev;
    for (int i = 0; i < n_dev; i++) {
        m->dev[i] = devices ? devices[i] : i;
        m->count[i] = base + (i < rem ? 1 : 0);
        m->first[i] = i * base + std::min<int64_t>(i, rem);
    }
    const int32_t st = fan_out(m, [&](int i) {
        return scmc_create(&m->h[i], m->dev[i], precision, n_sites, d, m->count[i], replica_offset + m->first[i], z_max, nbr_site, nbr_label,
                           n_labels, J, K, gen_re, gen_im, gensq_re, gensq_im, colour);
    });
    if (st != SCMC_OK) { const std::string keep = scmc_last_error(); scmc_multi_destroy(m); set_error(keep); return st; }
    *out = m;
    return SCMC_OK;
}

int32_t scmc_multi_destroy(scmc_multi_t* m) {
    if (!m) return SCMC_OK;
    for (int i = 0; i < m->n_dev; i++) {
        if (m->nccl_ready) {
            cudaSetDevice(m->dev[i]);
            if (m->d_send[i]) cudaFree(m->d_send[i]);
            if (m->d_recv[i]) cudaFree(m->d_recv[i]);
            if (m->comm[i]) g_nccl.CommDestroy(m->comm[i]);
        }
        if (m->h[i]) scmc_destroy(m->h[i]);
    }
    delete m;
    return SCMC_OK;
}

int32_t scmc_multi_count(scmc_multi_t* m, int32_t* n_dev, int64_t* n_replicas_total) {
    SCMC_ARG(m, "null argument");
    if (n_dev) *n_dev = m->n_dev;
    if (n_replicas_total) *n_replicas_total = m->R_total;
    return SCMC_OK;
}

int32_t scmc_multi_handle(scmc_multi_t* m, int32_t i, scmc_t** h, int64_t* first_replica, int64_t* n_replicas) {
    SCMC_ARG(m && i >= 0 && i < m->n_dev, "scmc_multi_handle: bad device index");
    if (h) *h = m->h[i];
    if (first_replica) *first_replica = m->first[i];
    if (n_replicas) *n_replicas = m->count[i];
    return SCMC_OK;
}

int32_t scmc_multi_randomize(scmc_multi_t* m, uint64_t seed) {
    SCMC_ARG(m, "null argument");
    return fan_out(m, [&](int i) { return scmc_randomize(m->h[i], seed); });
}

int32_t scmc_multi_set_stream_version(scmc_multi_t* m, int32_t version) {
    SCMC_ARG(m, "null argument");
    for (int i = 0; i < m->n_dev; i++) SCMC_TRY(scmc_set_stream_version(m->h[i], version));
    return SCMC_OK;
}

int32_t scmc_multi_set_sweep_mode(scmc_multi_t* m, int32_t mode, int64_t batch) {
    SCMC_ARG(m, "null argument");
    for (int i = 0; i < m->n_dev; i++) SCMC_TRY(scmc_set_sweep_mode(m->h[i], mode, batch));
    return SCMC_OK;
}

int32_t scmc_multi_sweep(scmc_multi_t* m, int64_t n_sweeps, int32_t hits, const double* beta, const double* sigma, uint64_t seed,
                         uint64_t sweep_counter, int64_t* accepted, int64_t* attempted) {
    SCMC_ARG(m && beta && sigma, "null argument");
    return fan_out(m, [&](int i) {
        const int64_t f = m->first[i];
        return scmc_sweep(m->h[i], n_sweeps, hits, beta + f, sigma + f, seed, sweep_counter, accepted ? accepted + f : nullptr,
                          attempted ? attempted + f : nullptr);
    });
}

int32_t scmc_multi_energy(scmc_multi_t* m, double* E) {
    SCMC_ARG(m && E, "null argument");
    return fan_out(m, [&](int i) { return scmc_energy(m->h[i], E + m->first[i]); });
}

int32_t scmc_multi_optimize(scmc_multi_t* m, int64_t n_sweeps, double* E) {
    SCMC_ARG(m && E, "null argument");
    return fan_out(m, [&](int i) { return scmc_optimize(m->h[i], n_sweeps, E + m->first[i]); });
}

int32_t scmc_multi_get_state(scmc_multi_t* m, int64_t r0, int64_t count, double* re, double* im) {
    SCMC_ARG(m && re && im && r0 >= 0 && count >= 0 && r0 + count <= m->R_total, "scmc_multi_get_state: bad replica range");
    return fan_out(m, [&](int i) -> int32_t {
        const int64_t lo = std::max(r0, m->first[i]), hi = std::min(r0 + count, m->first[i] + m->count[i]);
        if (lo >= hi) return SCMC_OK;
        int64_t info[8];
        SCMC_TRY(scmc_info(m->h[i], info));
        const size_t per_rep = (size_t)info[6] * (size_t)m->h[i]->d;
        return scmc_get_state(m->h[i], lo - m->first[i], hi - lo, re + (size_t)(lo - r0) * per_rep, im + (size_t)(lo - r0) * per_rep);
    });
}

// measure! of every replica: each device reduces its replicas' rows; ONE ncclAllGather; device 0 -> host
int32_t scmc_multi_measure(scmc_multi_t* m, double* obs) {
    SCMC_ARG(m && obs, "null argument");
    if (m->n_dev == 1) return scmc_measure(m->h[0], obs);
    SCMC_TRY(ensure_nccl(m));
    SCMC_TRY(fan_out(m, [&](int i) -> int32_t {
        SCMC_CK(cudaSetDevice(m->dev[i]));
        SCMC_TRY(launch_measure(m->h[i], m->d_send[i]));
        return SCMC_OK;
    }));
    ncclResult_t r = g_nccl.GroupStart();
    if (r != ncclSuccess) return nccl_fail(r, "ncclGroupStart");
    for (int i = 0; i < m->n_dev; i++) {
        r = g_nccl.AllGather(m->d_send[i], m->d_recv[i], (size_t)m->per * SCMC_NOBS, ncclDouble, m->comm[i], m->h[i]->stream);
        if (r != ncclSuccess) { g_nccl.GroupEnd(); return nccl_fail(r, "ncclAllGather"); }
    }
    r = g_nccl.GroupEnd();
    if (r != ncclSuccess) return nccl_fail(r, "ncclGroupEnd");
    for (int i = 0; i < m->n_dev; i++) {
        SCMC_CK(cudaSetDevice(m->dev[i]));
        SCMC_CK(cudaStreamSynchronize(m->h[i]->stream));
    }
    SCMC_CK(cudaSetDevice(m->dev[0]));
    for (int i = 0; i < m->n_dev; i++)
        SCMC_CK(cudaMemcpy(obs + (size_t)m->first[i] * SCMC_NOBS, m->d_recv[0] + (size_t)i * m->per * SCMC_NOBS, (size_t)m->count[i] * SCMC_NOBS * 8,
                           cudaMemcpyDeviceToHost));
    return SCMC_OK;
}

// run! on every device at once.  p->T, p->T_i, p->sigma: [R_total]; out->series [rows][R_total][NOBS], mean [R_total][NOBS], acc_rate [R_total],
// therm_energy [n_log][R_total] (any of them NULL to skip)
int32_t scmc_multi_run_metropolis(scmc_multi_t* m, const scmc_run_params* p, scmc_run_results* out) {
    SCMC_ARG(m && p && out && p->T && p->T_i && p->sigma, "null argument");
    const int64_t rate = std::max<int64_t>(1, p->measurement_rate);
    const int64_t total_all = p->n_therm + p->n_measure;
    const int64_t total = p->stop_sweep > 0 ? std::min<int64_t>(total_all, p->stop_sweep) : total_all;
    const int64_t meas_start = std::max<int64_t>(p->n_therm + 1, (int64_t)p->sweep0 + 1);
    const int64_t rows_max = total >= meas_start ? total / rate - (meas_start - 1) / rate : 0;
    const int64_t therm_end = std::min<int64_t>(p->n_therm, total);
    const int64_t n_log = std::max<int64_t>(0, therm_end / rate - (int64_t)p->sweep0 / rate);
    std::vector<std::vector<double>> ser(m->n_dev), thm(m->n_dev);
    std::vector<int64_t> rows(m->n_dev, 0);
    SCMC_TRY(fan_out(m, [&](int i) -> int32_t {
        const int64_t f = m->first[i], c = m->count[i];
        scmc_run_params q = *p;
        q.T = p->T + f; q.T_i = p->T_i + f; q.sigma = p->sigma + f;
        scmc_run_results o{};
        if (out->series) { ser[i].assign((size_t)std::max<int64_t>(rows_max, 1) * c * SCMC_NOBS, 0.0); o.series = ser[i].data(); }
        if (out->therm_energy) { thm[i].assign((size_t)std::max<int64_t>(n_log, 1) * c, 0.0); o.therm_energy = thm[i].data(); }
        o.mean = out->mean ? out->mean + (size_t)f * SCMC_NOBS : nullptr;
        o.acc_rate = out->acc_rate ? out->acc_rate + f : nullptr;
        const int32_t st = scmc_run_metropolis(m->h[i], &q, &o);
        rows[i] = o.n_rows;
        return st;
    }));
    out->n_rows = rows[0];
    for (int i = 0; i < m->n_dev; i++) {
        const int64_t f = m->first[i], c = m->count[i];
        if (out->series)
            for (int64_t r = 0; r < rows[i]; r++)
                memcpy(out->series + ((size_t)r * m->R_total + f) * SCMC_NOBS, ser[i].data() + (size_t)r * c * SCMC_NOBS, (size_t)c * SCMC_NOBS * 8);
        if (out->therm_energy)
            for (int64_t r = 0; r < n_log; r++) memcpy(out->therm_energy + (size_t)r * m->R_total + f, thm[i].data() + (size_t)r * c, (size_t)c * 8);
    }
    return SCMC_OK;
}

// cooling loop of runAnnealing! on every device at once; result arrays [R_total]
int32_t scmc_multi_anneal(scmc_multi_t* m, const scmc_anneal_params* p, scmc_anneal_results* out) {
    SCMC_ARG(m && p && out, "null argument");
    const bool logs = p->log_rate > 0 && out->energy_log && out->beta_log && out->n_log_cap > 0;
    std::vector<std::vector<double>> el(m->n_dev), bl(m->n_dev);
    std::vector<int64_t> nl(m->n_dev, 0);
    SCMC_TRY(fan_out(m, [&](int i) {
        const int64_t f = m->first[i], c = m->count[i];
        scmc_anneal_results o{};
        o.E = out->E ? out->E + f : nullptr; o.beta = out->beta ? out->beta + f : nullptr;
        o.sigma = out->sigma ? out->sigma + f : nullptr; o.sweeps = out->sweeps ? out->sweeps + f : nullptr;
        if (logs) {
            el[i].assign((size_t)out->n_log_cap * c, 0.0); bl[i].assign((size_t)out->n_log_cap * c, 0.0);
            o.energy_log = el[i].data(); o.beta_log = bl[i].data(); o.n_log_cap = out->n_log_cap;
        }
        const int32_t st = scmc_anneal(m->h[i], p, &o);
        nl[i] = o.n_log;
        return st;
    }));
    out->n_log = 0;
    if (logs) {
        // devices whose replicas all stopped early wrote fewer rows: pad with their last row so that every replica has the same grid
        for (int i = 0; i < m->n_dev; i++) out->n_log = std::max(out->n_log, nl[i]);
        for (int i = 0; i < m->n_dev; i++) {
            const int64_t f = m->first[i], c = m->count[i];
            for (int64_t r = 0; r < out->n_log; r++) {
                const int64_t src = nl[i] ? std::min(r, nl[i] - 1) : -1;
                for (int64_t q = 0; q < c; q++) {
                    out->energy_log[r * m->R_total + f + q] = src >= 0 ? el[i][(size_t)src * c + q] : 0.0;
                    out->beta_log[r * m->R_total + f + q] = src >= 0 ? bl[i][(size_t)src * c + q] : 0.0;
                }
            }
        }
    }
    return SCMC_OK;
}

}  // extern "C"
