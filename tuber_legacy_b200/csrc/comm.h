// The small exchanges of a multi-GPU frame: two all-gathers of a few words / one histogram per
// rank, one broadcast of buffer handles, one integer sum (which is also the barrier). Two transports behind one interface:
//
//   NcclComm   one process per GPU (or any mix): NCCL over NVLink / NVSwitch. libnccl.so.2 is opened at run time
//              (dlopen) rather than linked, so a host that already carries an NCCL (e.g. the one bundled with
//              PyTorch) shares that copy instead of loading a second one.
//   LocalComm  several contexts inside ONE process (one host thread per context, one or several GPUs): a
//              rendezvous in host memory plus peer copies. This is also what lets a single-GPU box run the
//              sharded frame end to end in the tests (NCCL refuses two ranks on one GPU).
//
// The instance payload never goes through either of them: the emit kernels store straight into the presenting
// rank's buffers (peer mapping: CUDA IPC between processes, plain pointers inside one process).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>

#include <condition_variable>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace tb {

constexpr size_t kCommIdBytes = 128;
static_assert(sizeof(ncclUniqueId) == kCommIdBytes, "tb_shard id is an ncclUniqueId");
constexpr char kLocalIdMagic[8] = {'T', 'B', 'L', 'O', 'C', 'A', 'L', 0};

struct Comm {
    int rank = 0, nranks = 1;
    std::string err;
    virtual ~Comm() {}
    // every rank contributes `bytes` from d_send; d_recv (nranks * bytes) holds them rank-major afterwards
    virtual bool all_gather(const void* d_send, void* d_recv, size_t bytes, cudaStream_t s) = 0;
    virtual bool broadcast(void* d_buf, size_t bytes, int root, cudaStream_t s) = 0;
    // *total (if given) = sum of every rank's `value`. Also THE barrier: everything every rank enqueued on its stream before
    // the call is complete (and visible) after it
    virtual bool sum(cudaStream_t s, int value, int* total) = 0;
    bool barrier(cudaStream_t s) { return sum(s, 0, nullptr); }
    virtual bool same_process() const = 0;
    // true if another rank of the group computes on this rank's GPU (only the process-local transport allows that): kernels
    // of different ranks may then queue behind each other, so no kernel of one rank may WAIT for a kernel of another
    virtual bool shares_device() const { return false; }
};

// one word in and out of the reduction without a copy engine (see peek_async in tuber_b200.cu)
__global__ void comm_put_kernel(int* d, int v) { *d = v; }
__global__ void comm_get_kernel(int* pinned, const int* d) { *pinned = *d; }

// ---- NCCL, loaded at run time --------------------------------------------------------------------------------------

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;

    static NcclApi& get() {
        static NcclApi api;
        static std::once_flag once;
        std::call_once(once, [] { api.load(); });
        return api;
    }
    bool ok() const { return lib != nullptr; }

  private:
    void load() {
        const char* names[] = {getenv("TB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        // an NCCL the host process already loaded wins (RTLD_NOLOAD matches by soname)
        for (const char* n : names)
            if (n && !lib) lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_LOCAL);
        for (const char* n : names)
            if (n && !lib) lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (!lib) {
            err = std::string("libnccl.so.2 not found (") + (dlerror() ? dlerror() : "?") + ")";
            return;
        }
        bool all = true;
        auto sym = [&](const char* name) {
            void* p = dlsym(lib, name);
            if (!p) all = false;
            return p;
        };
        GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
        CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
        AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
        Broadcast = reinterpret_cast<decltype(Broadcast)>(sym("ncclBroadcast"));
        AllReduce = reinterpret_cast<decltype(AllReduce)>(sym("ncclAllReduce"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
        if (!all) {
            err = "libnccl.so.2 lacks a required symbol";
            lib = nullptr;
        }
    }
};

struct NcclComm : Comm {
    ncclComm_t comm = nullptr;
    int* d_token = nullptr;
    int* h_token = nullptr;  // pinned
    bool init(int rank_, int nranks_, const uint8_t id[kCommIdBytes]) {
        NcclApi& api = NcclApi::get();
        if (!api.ok()) {
            err = api.err;
            return false;
        }
        rank = rank_;
        nranks = nranks_;
        ncclUniqueId uid;
        memcpy(&uid, id, sizeof(uid));
        ncclResult_t r = api.CommInitRank(&comm, nranks, uid, rank);
        if (r != ncclSuccess) {
            err = std::string("ncclCommInitRank: ") + api.GetErrorString(r);
            comm = nullptr;
            return false;
        }
        if (cudaMalloc((void**)&d_token, 2 * sizeof(int)) != cudaSuccess) {
            err = "cudaMalloc(barrier token)";
            return false;
        }
        cudaMemset(d_token, 0, 2 * sizeof(int));
        if (cudaMallocHost((void**)&h_token, sizeof(int)) != cudaSuccess) {
            err = "cudaMallocHost(barrier token)";
            return false;
        }
        return true;
    }
    ~NcclComm() override {
        if (comm) NcclApi::get().CommDestroy(comm);
        cudaFree(d_token);
        if (h_token) cudaFreeHost(h_token);
    }
    bool check(ncclResult_t r, const char* what) {
        if (r == ncclSuccess) return true;
        err = std::string(what) + ": " + NcclApi::get().GetErrorString(r);
        return false;
    }
    bool all_gather(const void* d_send, void* d_recv, size_t bytes, cudaStream_t s) override {
        return check(NcclApi::get().AllGather(d_send, d_recv, bytes, ncclUint8, comm, s), "ncclAllGather");
    }
    bool broadcast(void* d_buf, size_t bytes, int root, cudaStream_t s) override {
        return check(NcclApi::get().Broadcast(d_buf, d_buf, bytes, ncclUint8, root, comm, s), "ncclBroadcast");
    }
    bool sum(cudaStream_t s, int value, int* total) override {
        // stream order: every kernel enqueued before has finished (its peer stores are visible system-wide at kernel
        // end) before this rank enters the all-reduce, and the all-reduce finishes only when every rank has entered it
        comm_put_kernel<<<1, 1, 0, s>>>(d_token, value);
        if (!check(NcclApi::get().AllReduce(d_token, d_token + 1, 1, ncclInt32, ncclSum, comm, s), "ncclAllReduce")) return false;
        comm_get_kernel<<<1, 1, 0, s>>>(h_token, d_token + 1);
        if (cudaStreamSynchronize(s) != cudaSuccess) {
            err = "barrier: stream synchronisation failed";
            return false;
        }
        if (total) *total = *h_token;
        return true;
    }
    bool same_process() const override { return false; }
};

// ---- several contexts inside one process -----------------------------------------------------------------------------

struct LocalGroup {
    std::mutex m;
    std::condition_variable cv;
    int nranks = 0, arrived = 0;
    uint64_t generation = 0;
    std::vector<const void*> send;
    std::vector<int> device, value;
    void wait_all() {  // host barrier of the group's threads
        std::unique_lock<std::mutex> lk(m);
        const uint64_t gen = generation;
        if (++arrived == nranks) {
            arrived = 0;
            ++generation;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return generation != gen; });
        }
    }
};

struct LocalComm : Comm {
    std::shared_ptr<LocalGroup> g;
    int device = 0;
    static std::shared_ptr<LocalGroup> group_of(const uint8_t id[kCommIdBytes], int nranks) {
        static std::mutex reg_m;
        static std::map<std::string, std::weak_ptr<LocalGroup>> reg;
        std::lock_guard<std::mutex> lk(reg_m);
        const std::string key(reinterpret_cast<const char*>(id), kCommIdBytes);
        std::shared_ptr<LocalGroup> p = reg[key].lock();
        if (!p) {
            p = std::make_shared<LocalGroup>();
            p->nranks = nranks;
            p->send.assign(nranks, nullptr);
            p->device.assign(nranks, 0);
            p->value.assign(nranks, 0);
            reg[key] = p;
        }
        return p;
    }
    bool init(int rank_, int nranks_, const uint8_t id[kCommIdBytes], int device_) {
        rank = rank_;
        nranks = nranks_;
        device = device_;
        g = group_of(id, nranks_);
        if (g->nranks != nranks_) {
            err = "ranks of one local group disagree about its size";
            return false;
        }
        g->device[rank] = device_;
        g->wait_all();
        for (int r = 0; r < nranks; ++r) {  // peer access for the direct stores and the copies below
            if (g->device[r] == device) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, device, g->device[r]);
            if (can && cudaDeviceEnablePeerAccess(g->device[r], 0) != cudaSuccess) cudaGetLastError();  // already enabled is fine
        }
        return true;
    }
    bool sync(cudaStream_t s) {
        if (cudaStreamSynchronize(s) == cudaSuccess) return true;
        err = "local transport: stream synchronisation failed";
        return false;
    }
    bool all_gather(const void* d_send, void* d_recv, size_t bytes, cudaStream_t s) override {
        bool ok = sync(s);  // my contribution is complete
        g->send[rank] = d_send;
        g->wait_all();
        for (int r = 0; ok && r < nranks; ++r)
            ok = cudaMemcpyPeerAsync(static_cast<char*>(d_recv) + (size_t)r * bytes, device, g->send[r], g->device[r], bytes, s) == cudaSuccess;
        ok = ok && sync(s);
        g->wait_all();  // every rank has read every send buffer: they may be reused
        if (!ok && err.empty()) err = "local transport: all_gather failed";
        return ok;
    }
    bool broadcast(void* d_buf, size_t bytes, int root, cudaStream_t s) override {
        bool ok = sync(s);
        g->send[rank] = d_buf;
        g->wait_all();
        if (rank != root) ok = ok && cudaMemcpyPeerAsync(d_buf, device, g->send[root], g->device[root], bytes, s) == cudaSuccess && sync(s);
        g->wait_all();
        if (!ok && err.empty()) err = "local transport: broadcast failed";
        return ok;
    }
    bool sum(cudaStream_t s, int v, int* total) override {
        const bool ok = sync(s);
        g->value[rank] = v;
        g->wait_all();
        int t = 0;
        for (int r = 0; r < nranks; ++r) t += g->value[r];
        g->wait_all();  // every rank has read every value: they may be overwritten
        if (total) *total = t;
        return ok;
    }
    bool same_process() const override { return true; }
    bool shares_device() const override {
        for (int r = 0; r < nranks; ++r)
            if (r != rank && g->device[r] == device) return true;
        return false;
    }
};

}  // namespace tb
