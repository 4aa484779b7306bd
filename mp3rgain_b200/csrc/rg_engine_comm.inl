// rg_engine_comm.inl -- album mode across GPUs: the engine owns the one exchange step of the path.
// Included at the end of rg_engine.cu.
//
// analyze_album (src/replaygain.rs:1033-1074) accumulates every track's histogram into one and takes the maximum of the
// peaks.  Sharded over ranks (one process per GPU, tracks split by bytes) that loop becomes: every rank accumulates its own
// tracks on its device (K5's u64 atomics), then ONE exchange -- all-reduce(sum) of the 12000 u64 counters and all-reduce(max)
// of two u64 words {peak bit pattern, error key}, issued as one NCCL group on the engine's own stream right behind the last
// K5 -- and every rank runs the percentile kernel on the merged histogram.  The error key carries the reference's "first
// failing file aborts the album" (:1055) across ranks: (2^54-1 - global index of the rank's first failing file) << 8 |
// -status, so the maximum names the failing file with the smallest global index and its status on every rank.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the library the host process already uses, if any): the engine has no
// link-time dependency on it and single-GPU callers never load it.  A host that has another transport (MPI, gloo -- the
// CPU tests) attaches an exchange function instead; the two buffers then travel through pinned host memory.
#include <dlfcn.h>
#include <mutex>

// the few NCCL declarations the engine needs (nccl.h, NCCL 2.x ABI)
typedef struct ncclComm* rg_ncclComm_t;
typedef struct { char internal[128]; } rg_ncclUniqueId;
enum { RG_NCCL_SUCCESS = 0, RG_NCCL_UINT64 = 5, RG_NCCL_SUM = 0, RG_NCCL_MAX = 2 };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(rg_ncclUniqueId*) = nullptr;
    int (*CommInitRank)(rg_ncclComm_t*, int, rg_ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(rg_ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, rg_ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) return;
        api.GetUniqueId = (int (*)(rg_ncclUniqueId*))dlsym(api.lib, "ncclGetUniqueId");
        api.CommInitRank = (int (*)(rg_ncclComm_t*, int, rg_ncclUniqueId, int))dlsym(api.lib, "ncclCommInitRank");
        api.CommDestroy = (int (*)(rg_ncclComm_t))dlsym(api.lib, "ncclCommDestroy");
        api.AllReduce = (int (*)(const void*, void*, size_t, int, int, rg_ncclComm_t, cudaStream_t))dlsym(api.lib, "ncclAllReduce");
        api.GroupStart = (int (*)())dlsym(api.lib, "ncclGroupStart");
        api.GroupEnd = (int (*)())dlsym(api.lib, "ncclGroupEnd");
        api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.GroupStart || !api.GroupEnd) {
            dlclose(api.lib);
            api.lib = nullptr;
        }
    });
    if (!api.lib) { g_last_error = "libnccl.so.2 could not be loaded: multi-GPU album mode needs NCCL (or an attached exchange function)"; return nullptr; }
    return &api;
}
static int fail_nccl(NcclApi* api, int r, const char* what) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s failed: %s", what, api->GetErrorString ? api->GetErrorString(r) : "NCCL error");
    g_last_error = buf;
    return RG_ERR_CUDA;
}

extern "C" int rg_comm_unique_id(uint8_t id[RG_COMM_ID_BYTES]) {
    if (!id) return RG_ERR_INVALID_ARG;
    NcclApi* api = nccl_api();
    if (!api) return RG_ERR_NO_DEVICE;
    rg_ncclUniqueId u;
    int r = api->GetUniqueId(&u);
    if (r != RG_NCCL_SUCCESS) return fail_nccl(api, r, "ncclGetUniqueId");
    static_assert(sizeof u == RG_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    memcpy(id, &u, sizeof u);
    return RG_OK;
}

static void comm_drop(rg_engine* e) {
    if (e->comm && e->comm_owned) {
        NcclApi* api = nccl_api();
        if (api) api->CommDestroy((rg_ncclComm_t)e->comm);
    }
    e->comm = nullptr; e->comm_owned = false; e->xchg_fn = nullptr; e->xchg_ctx = nullptr; e->rank = 0; e->nranks = 1;
}

extern "C" int rg_engine_comm_init(rg_engine* e, int rank, int nranks, const uint8_t id[RG_COMM_ID_BYTES]) {
    if (!e || !id || nranks < 1 || rank < 0 || rank >= nranks) return RG_ERR_INVALID_ARG;
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    NcclApi* api = nccl_api();
    if (!api) return RG_ERR_NO_DEVICE;
    CK(cudaSetDevice(e->device));
    comm_drop(e);
    rg_ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    rg_ncclComm_t comm = nullptr;
    int r = api->CommInitRank(&comm, nranks, u, rank);
    if (r != RG_NCCL_SUCCESS) return fail_nccl(api, r, "ncclCommInitRank");
    e->comm = comm; e->comm_owned = true; e->rank = rank; e->nranks = nranks;
    return RG_OK;
}

extern "C" int rg_engine_comm_attach(rg_engine* e, void* nccl_comm, int rank, int nranks) {
    if (!e || !nccl_comm || nranks < 1 || rank < 0 || rank >= nranks) return RG_ERR_INVALID_ARG;
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    if (!nccl_api()) return RG_ERR_NO_DEVICE;
    comm_drop(e);
    e->comm = nccl_comm; e->comm_owned = false; e->rank = rank; e->nranks = nranks;
    return RG_OK;
}

extern "C" int rg_engine_comm_attach_host(rg_engine* e, int rank, int nranks, rg_exchange_fn fn, void* ctx) {
    if (!e || !fn || nranks < 1 || rank < 0 || rank >= nranks) return RG_ERR_INVALID_ARG;
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    comm_drop(e);
    e->xchg_fn = fn; e->xchg_ctx = ctx; e->rank = rank; e->nranks = nranks;
    return RG_OK;
}

extern "C" int rg_engine_comm_release(rg_engine* e) {
    if (!e) return RG_ERR_INVALID_ARG;
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    CK(cudaSetDevice(e->device));
    comm_drop(e);
    return RG_OK;
}

#define RG_ERRKEY_TOP ((1ull << 54) - 1)   // keys stay below 2^62: they order the same as u64 (NCCL) and as i64 (a host transport)

static int album_sharded_impl(rg_engine* e, rg_corpus* c, uint64_t shard_offset, rg_track_result* out, rg_album_result* album,
                              uint64_t* album_hist_out, int64_t* first_bad_global) {
    if (first_bad_global) *first_bad_global = -1;
    int rc = album_buffers(e);
    if (rc != RG_OK) return rc;
    unsigned long long* d_hist = e->d_album.as<unsigned long long>();
    CK(e->d_album_peak.ensure(64));
    unsigned long long* d_x = e->d_album_peak.as<unsigned long long>();   // [0] peak bits (max), [1] error key (max)
    CK(cudaMemsetAsync(d_hist, 0, RG_HIST_BINS * sizeof(unsigned long long), e->stream));
    CK(cudaMemsetAsync(d_x, 0, 2 * sizeof(unsigned long long), e->stream));
    int bad = -1;
    rc = analyze_corpus_impl(e, c, out, nullptr, d_hist, d_x, true, &bad);
    if (rc != RG_OK && bad < 0) return rc;   // not a file's failure (CUDA, memory): this rank cannot take part any more
    CK(e->h_xchg.ensure((RG_HIST_BINS + 8) * sizeof(uint64_t)));
    uint64_t* hx = e->h_xchg.as<uint64_t>();
    if (bad >= 0) {
        hx[0] = ((RG_ERRKEY_TOP - (shard_offset + (uint64_t)bad)) << 8) | (uint64_t)(uint8_t)(-rc);
        CK(cudaMemcpyAsync(d_x + 1, hx, sizeof(uint64_t), cudaMemcpyHostToDevice, e->stream));
    }
    CK(cudaEventRecord(e->ev[5], e->stream));
    if (e->nranks > 1 && e->comm) {
        NcclApi* api = nccl_api();
        if (!api) return RG_ERR_NO_DEVICE;
        int r = api->GroupStart();
        if (r == RG_NCCL_SUCCESS) r = api->AllReduce(d_hist, d_hist, RG_HIST_BINS, RG_NCCL_UINT64, RG_NCCL_SUM, (rg_ncclComm_t)e->comm, e->stream);
        if (r == RG_NCCL_SUCCESS) r = api->AllReduce(d_x, d_x, 2, RG_NCCL_UINT64, RG_NCCL_MAX, (rg_ncclComm_t)e->comm, e->stream);
        const int r2 = api->GroupEnd();
        if (r != RG_NCCL_SUCCESS || r2 != RG_NCCL_SUCCESS) return fail_nccl(api, r != RG_NCCL_SUCCESS ? r : r2, "ncclAllReduce");
        e->stats.kernel_launches += 2;
    } else if (e->nranks > 1 && e->xchg_fn) {
        // another transport: the 96 KB travel through pinned host memory
        CK(cudaMemcpyAsync(hx + 8, d_hist, RG_HIST_BINS * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaMemcpyAsync(hx + 2, d_x, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        if (e->xchg_fn(e->xchg_ctx, hx + 8, RG_HIST_BINS, hx + 2, 2) != 0) { g_last_error = "the attached exchange function failed"; return RG_ERR_CUDA; }
        CK(cudaMemcpyAsync(d_hist, hx + 8, RG_HIST_BINS * sizeof(uint64_t), cudaMemcpyHostToDevice, e->stream));
        CK(cudaMemcpyAsync(d_x, hx + 2, 2 * sizeof(uint64_t), cudaMemcpyHostToDevice, e->stream));
        e->stats.d2h_bytes += (RG_HIST_BINS + 2) * sizeof(uint64_t);
        e->stats.h2d_bytes += (RG_HIST_BINS + 2) * sizeof(uint64_t);
    } else if (e->nranks > 1) {
        g_last_error = "rg_analyze_album_sharded: the engine is one rank of several but has no communicator";
        return RG_ERR_INVALID_ARG;
    }
    CK(cudaEventRecord(e->ev[6], e->stream));
    CK(e->d_misc.ensure(64));
    k_album_percentile<<<1, K5_THREADS, 0, e->stream>>>(d_hist, d_x, e->d_misc.as<double>());
    e->stats.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(hx, e->d_misc.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(hx + 3, d_x + 1, sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
    if (album_hist_out) CK(cudaMemcpyAsync(album_hist_out, d_hist, RG_HIST_BINS * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->stats.d2h_bytes += 4 * sizeof(uint64_t) + (album_hist_out ? RG_HIST_BINS * sizeof(uint64_t) : 0);
    float ms = 0;
    if (cudaEventElapsedTime(&ms, e->ev[5], e->ev[6]) == cudaSuccess) e->stats.last_ms_exchange = ms;
    const uint64_t key = hx[3];
    if (key) {   // some rank's file failed: the album fails everywhere, with that file's status (src/replaygain.rs:1055)
        if (first_bad_global) *first_bad_global = (int64_t)(RG_ERRKEY_TOP - (key >> 8));
        return -(int)(key & 0xFF);
    }
    double h[3];
    memcpy(h, hx, sizeof h);
    album->album_loudness_db = h[0];
    album->album_gain_db = h[1];
    album->album_peak = h[2];
    return RG_OK;
}

extern "C" int rg_analyze_album_sharded_corpus(rg_engine* e, rg_corpus* c, uint64_t shard_offset, rg_track_result* out,
                                               rg_album_result* album, uint64_t* album_hist_out, int64_t* first_bad_global) {
    if (!e || !c || !album || (!out && !c->len.empty())) return RG_ERR_INVALID_ARG;
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    return album_sharded_impl(e, c, shard_offset, out, album, album_hist_out, first_bad_global);
}

extern "C" int rg_analyze_album_sharded(rg_engine* e, const rg_input* in, size_t n, uint64_t shard_offset, rg_track_result* out,
                                        rg_album_result* album, uint64_t* album_hist_out, int64_t* first_bad_global) {
    if (!e || !album || (n && (!in || !out))) return RG_ERR_INVALID_ARG;
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    rg_corpus c;
    int rc = host_corpus(e, in, n, c);
    if (rc != RG_OK) return rc;
    return album_sharded_impl(e, &c, shard_offset, out, album, album_hist_out, first_bad_global);
}
