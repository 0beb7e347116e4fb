// emb_multi.cpp -- one process, every GPU of the box: the C-ABI form of SURVEY 8e.
//
// The hot path shards by global sample index with no exchange (device d owns the contiguous range emb_shard_range(n, d, D);
// the Philox stream is keyed by the global index, so the union equals the single-device result for any D).  One host thread per
// device drives the existing single-device entry point; the ONE collective of the job sums the verification histograms over the
// devices with ncclAllReduce (ncclCommInitAll inside this process).  NCCL is bound at run time (dlopen of libnccl.so.2 -- the
// copy torch ships or the system one) so that libemb200.so keeps libcudart as its only link-time dependency; without NCCL,
// or with a single device, the histograms are summed through the host (a few KB).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/emb200.h"

namespace {

// the four NCCL calls used, with the ABI of nccl.h 2.x (ncclUint64 = 5, ncclSum = 0)
struct Nccl {
    void* lib = nullptr;
    int (*CommInitAll)(void** comms, int ndev, const int* devlist) = nullptr;
    int (*AllReduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t st) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*CommDestroy)(void* comm) = nullptr;
    bool ok() const { return CommInitAll && AllReduce && GroupStart && GroupEnd && CommDestroy; }
};
Nccl& nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            n.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (n.lib) break;
        }
        if (!n.lib) return;
        n.CommInitAll = (decltype(n.CommInitAll))dlsym(n.lib, "ncclCommInitAll");
        n.AllReduce = (decltype(n.AllReduce))dlsym(n.lib, "ncclAllReduce");
        n.GroupStart = (decltype(n.GroupStart))dlsym(n.lib, "ncclGroupStart");
        n.GroupEnd = (decltype(n.GroupEnd))dlsym(n.lib, "ncclGroupEnd");
        n.CommDestroy = (decltype(n.CommDestroy))dlsym(n.lib, "ncclCommDestroy");
    });
    return n;
}

thread_local std::string g_multi_err;

}  // namespace

extern "C" {

void emb_shard_range(int64_t n, int32_t shard, int32_t n_shards, int64_t* first, int64_t* count) {
    const int64_t per = n_shards > 0 ? (n + n_shards - 1) / n_shards : n;
    const int64_t f = std::min<int64_t>(n, (int64_t)shard * per);
    if (first) *first = f;
    if (count) *count = std::max<int64_t>(0, std::min<int64_t>(n, f + per) - f);
}

int emb_nccl_available(void) { return nccl().ok() ? 1 : 0; }

// counts[d]: device pointers on device devices[d], `len` uint64 each -> every one holds the sum (one collective)
int emb_allreduce_histograms(int32_t n_devices, const int32_t* devices, unsigned long long* const* counts, int64_t len) {
    if (n_devices <= 0 || !counts || len < 0) return EMB_E_ARG;
    if (n_devices == 1 || len == 0) return 0;
    Nccl& N = nccl();
    if (N.ok()) {
        std::vector<void*> comms((size_t)n_devices, nullptr);
        std::vector<int> devs((size_t)n_devices);
        for (int d = 0; d < n_devices; ++d) devs[(size_t)d] = devices ? devices[d] : d;
        if (N.CommInitAll(comms.data(), n_devices, devs.data()) == 0) {
            int rc = N.GroupStart();
            for (int d = 0; d < n_devices && rc == 0; ++d) {
                cudaSetDevice(devs[(size_t)d]);
                rc = N.AllReduce(counts[d], counts[d], (size_t)len, /*ncclUint64*/ 5, /*ncclSum*/ 0, comms[(size_t)d], 0);
            }
            if (rc == 0) rc = N.GroupEnd();
            for (int d = 0; d < n_devices; ++d) {
                cudaSetDevice(devs[(size_t)d]);
                cudaStreamSynchronize(0);
            }
            for (auto c : comms)
                if (c) N.CommDestroy(c);
            if (rc == 0) return 0;
        }
    }
    // no NCCL (or it failed): the few KB go through the host
    std::vector<unsigned long long> sum((size_t)len, 0ull), tmp((size_t)len);
    for (int d = 0; d < n_devices; ++d) {
        cudaSetDevice(devices ? devices[d] : d);
        if (cudaMemcpy(tmp.data(), counts[d], (size_t)len * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return EMB_E_CUDA;
        for (int64_t i = 0; i < len; ++i) sum[(size_t)i] += tmp[(size_t)i];
    }
    for (int d = 0; d < n_devices; ++d) {
        cudaSetDevice(devices ? devices[d] : d);
        if (cudaMemcpy(counts[d], sum.data(), (size_t)len * 8, cudaMemcpyHostToDevice) != cudaSuccess) return EMB_E_CUDA;
    }
    return 0;
}

int emb_sample_tracks_multi(const emb_model* m, const emb_rng* rng, int64_t n, int32_t T, const emb_sample_opts* opts,
                            int32_t n_devices, const emb_track_out* outs, unsigned long long* hist_initial,
                            unsigned long long* hist_transition) {
    if (!m || !rng || !opts || !outs || n < 0 || T < 1) return EMB_E_ARG;
    int avail = emb_device_count();
    if (avail <= 0) return EMB_E_CUDA;
    const int D = n_devices > 0 ? n_devices : avail;
    if (D > avail) return EMB_E_ARG;
    emb_model_info info;
    int rc0 = emb_model_get_info(m, &info);
    if (rc0) return rc0;
    const bool want_hist = hist_initial || hist_transition;
    const int64_t len_i = (int64_t)info.n_initial * 64, len_t = (int64_t)info.n_dyn * 64;
    // per-device histogram accumulators (device memory, initial followed by transition: one buffer = one collective)
    std::vector<unsigned long long*> dh((size_t)D, nullptr);
    std::vector<int> rcs((size_t)D, 0);
    std::vector<std::string> errs((size_t)D);
    std::vector<std::thread> th;
    for (int d = 0; d < D; ++d) {
        th.emplace_back([&, d] {
            int64_t first = 0, cnt = 0;
            emb_shard_range(n, d, D, &first, &cnt);
            if (cudaSetDevice(d) != cudaSuccess) { rcs[(size_t)d] = EMB_E_CUDA; return; }
            if (want_hist) {
                if (cudaMalloc((void**)&dh[(size_t)d], (size_t)(len_i + len_t) * 8) != cudaSuccess ||
                    cudaMemset(dh[(size_t)d], 0, (size_t)(len_i + len_t) * 8) != cudaSuccess) { rcs[(size_t)d] = EMB_E_CUDA; return; }
            }
            if (cnt == 0) return;
            emb_sample_opts o = *opts;
            o.device = d;
            o.stream = nullptr;
            if (o.start_per_sample) o.start_per_sample = nullptr;   // per-sample presets are a single-device feature
            emb_rng r = *rng;
            r.first_sample = rng->first_sample + (uint64_t)first;
            emb_track_out out = outs[d];
            const bool host = (o.mem & 0xFF) == EMB_MEM_HOST;
            std::vector<unsigned long long> hh;             // host-memory call: the pass accumulates into host counters ...
            if (want_hist && host) hh.assign((size_t)(len_i + len_t), 0ull);
            unsigned long long* hbase = host ? hh.data() : dh[(size_t)d];
            out.hist_initial = hist_initial ? hbase : nullptr;
            out.hist_transition = hist_transition ? hbase + len_i : nullptr;
            rcs[(size_t)d] = emb_sample_tracks(m, &r, cnt, T, &o, &out);
            if (rcs[(size_t)d] == 0 && want_hist && host &&     // ... which go to the device for the collective
                cudaMemcpy(dh[(size_t)d], hh.data(), hh.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess)
                rcs[(size_t)d] = EMB_E_CUDA;
            if (rcs[(size_t)d]) errs[(size_t)d] = emb_last_error();
        });
    }
    for (auto& t : th) t.join();
    int rc = 0;
    for (int d = 0; d < D; ++d)
        if (rcs[(size_t)d] && !rc) {
            rc = rcs[(size_t)d];
            g_multi_err = errs[(size_t)d];
        }
    if (rc == 0 && want_hist) {
        rc = emb_allreduce_histograms(D, nullptr, dh.data(), len_i + len_t);
        if (rc == 0) {
            cudaSetDevice(0);
            std::vector<unsigned long long> tmp((size_t)(len_i + len_t));
            if (cudaMemcpy(tmp.data(), dh[0], tmp.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess) rc = EMB_E_CUDA;
            if (rc == 0 && hist_initial) for (int64_t i = 0; i < len_i; ++i) hist_initial[i] += tmp[(size_t)i];
            if (rc == 0 && hist_transition) for (int64_t i = 0; i < len_t; ++i) hist_transition[i] += tmp[(size_t)(len_i + i)];
        }
    }
    for (int d = 0; d < D; ++d)
        if (dh[(size_t)d]) {
            cudaSetDevice(d);
            cudaFree(dh[(size_t)d]);
        }
    return rc;
}

const char* emb_multi_last_error(void) { return g_multi_err.c_str(); }

}  // extern "C"
