// Multi-GPU inside the product: one process per GPU, every rank holds a full scene + BVH replica and renders
// its share of the sample indices into a buffer of plain sums; this file combines those buffers over NVLink with NCCL and
// turns them into the frame.  The reference is single-GPU (example/path_tracer/pt_pass.cpp:39-57), so there is no
// interface to mirror: the entry points are the ones SURVEY.md 8b proposes (pb2_comm_* in include/pb2.h).
//
//   * NCCL is bound at run time (dlopen of libnccl.so.2): libpb2.so stays loadable on a single-GPU box without NCCL, and a
//     process that already holds a copy (torch.distributed in bench.py) shares it.
//   * The collective runs on the communicator's own stream, ordered after the scene's stream by an event, and reads the sum
//     buffer OUT OF PLACE (results land in a staging buffer / the frame buffer).  The scene's next k_accumulate — the only
//     kernel that writes the sum buffer — waits for it (Scene::accumulate_gate); everything before it in the next render
//     call (generate, eight rounds of extend / shade / shadow) overlaps the reduction of the previous one.
//   * mode PB2_REDUCE_ROOT: ncclReduce to `root`, finalize there.  mode PB2_REDUCE_ALL: ncclReduceScatter, every rank
//     finalizes its 1/N of the pixels, ncclAllGather in place in the frame buffer — no rank is the sink, and every rank ends
//     up with the frame (a display or a writer may sit on any of them).
#include "scene.cuh"
#include <cstring>
#include <dlfcn.h>
#include <nccl.h>

namespace pb2 {
void finalize_sum_on(cudaStream_t st, const float4 *sum, float4 *frame, uint64_t n, uint32_t spp); // wavefront.cu

namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi &nccl() {
    static NcclApi api;
    if (api.handle) return api;
    const char *names[] = { "libnccl.so.2", "libnccl.so" };
    void *h = nullptr;
    for (const char *n : names)
        if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) throw std::runtime_error(std::string("pb2_comm: cannot load NCCL (") + dlerror() + "); multi-GPU rendering needs libnccl.so.2");
    auto sym = [&](const char *name) {
        void *p = dlsym(h, name);
        if (!p) throw std::runtime_error(std::string("pb2_comm: libnccl lacks ") + name);
        return p;
    };
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.Reduce = reinterpret_cast<decltype(api.Reduce)>(sym("ncclReduce"));
    api.ReduceScatter = reinterpret_cast<decltype(api.ReduceScatter)>(sym("ncclReduceScatter"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.handle = h;
    return api;
}
void nccl_check(ncclResult_t r, const char *what) {
    if (r != ncclSuccess) throw std::runtime_error(std::string(what) + " failed: " + nccl().GetErrorString(r));
}
}// namespace

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, size = 1, device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t scene_ready = nullptr, done = nullptr;
    cudaEvent_t t_begin = nullptr, t_end = nullptr; // device time of the last reduction (on the communicator's stream)
    uint64_t last_bytes = 0;
    DevBuf<float4> staging; // reduced sums: the whole image on the root (ROOT mode) or this rank's slice (ALL mode)
    uint64_t reductions = 0;
    ~Comm() {
        if (stream) cudaStreamSynchronize(stream);
        if (comm) nccl().CommDestroy(comm);
        if (scene_ready) cudaEventDestroy(scene_ready);
        if (done) cudaEventDestroy(done);
        if (t_begin) cudaEventDestroy(t_begin);
        if (t_end) cudaEventDestroy(t_end);
        if (stream) cudaStreamDestroy(stream);
    }
};

void comm_unique_id(uint8_t id[PB2_COMM_ID_BYTES]) {
    static_assert(PB2_COMM_ID_BYTES == sizeof(ncclUniqueId), "pb2.h and nccl.h disagree on the id size");
    ncclUniqueId u;
    nccl_check(nccl().GetUniqueId(&u), "ncclGetUniqueId");
    memcpy(id, &u, sizeof u);
}
Comm *comm_create(int n_ranks, int rank, const uint8_t *id) {
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) throw std::runtime_error("pb2_comm_create: bad rank / size");
    auto c = std::make_unique<Comm>();
    c->rank = rank, c->size = n_ranks;
    PB2_CUDA(cudaGetDevice(&c->device));
    // Highest priority: the render kernels are persistent and fill every SM, so a collective queued at normal priority gets no
    // CTA slot until a whole trace launch has drained — on BOTH ranks at once, or it spins waiting for its peer.  Measured on
    // 2 x B200 (Cornell, 64 spp per GPU per step): reduce-scatter + all-gather at normal priority cost 15 ms per 88 ms step.
    int prio_lo = 0, prio_hi = 0;
    PB2_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    PB2_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi));
    PB2_CUDA(cudaEventCreateWithFlags(&c->scene_ready, cudaEventDisableTiming));
    PB2_CUDA(cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming));
    PB2_CUDA(cudaEventCreate(&c->t_begin));
    PB2_CUDA(cudaEventCreate(&c->t_end));
    if (n_ranks > 1) { // a single rank needs no NCCL at all
        if (!id) throw std::runtime_error("pb2_comm_create: the id of pb2_comm_unique_id (from rank 0) is required for more than one rank");
        ncclUniqueId u;
        memcpy(&u, id, sizeof u);
        nccl_check(nccl().CommInitRank(&c->comm, n_ranks, u, rank), "ncclCommInitRank");
    }
    return c.release();
}
void comm_destroy(Comm *c) { delete c; }

// sum buffers of all ranks -> frame = sum / total_spp.  Asynchronous: ordered after the scene's stream, runs on the
// communicator's stream; the scene's next accumulate kernel and pb2_comm_synchronize wait for it.
void comm_reduce_frames(Comm &c, Scene &s, const float4 *sum, float4 *frame, uint64_t n_pixels, uint32_t total_spp, int mode, int root) {
    if (!sum || !n_pixels || !total_spp) throw std::runtime_error("pb2_comm_reduce_frames: bad argument");
    if (root < 0 || root >= c.size) throw std::runtime_error("pb2_comm_reduce_frames: bad root");
    PB2_CUDA(cudaEventRecord(c.scene_ready, s.stream));
    PB2_CUDA(cudaStreamWaitEvent(c.stream, c.scene_ready, 0));
    PB2_CUDA(cudaEventRecord(c.t_begin, c.stream));
    c.last_bytes = n_pixels * sizeof(float4);
    if (c.size == 1) {
        if (!frame) throw std::runtime_error("pb2_comm_reduce_frames: frame buffer missing");
        finalize_sum_on(c.stream, sum, frame, n_pixels, total_spp);
    } else if (mode == PB2_REDUCE_ALL && n_pixels % (uint64_t)c.size == 0) {
        if (!frame) throw std::runtime_error("pb2_comm_reduce_frames: PB2_REDUCE_ALL needs the frame buffer on every rank");
        const uint64_t slice = n_pixels / c.size; // pixels per rank
        c.staging.ensure(slice);
        nccl_check(nccl().ReduceScatter(sum, c.staging.ptr, slice * 4, ncclFloat, ncclSum, c.comm, c.stream), "ncclReduceScatter");
        finalize_sum_on(c.stream, c.staging.ptr, frame + (uint64_t)c.rank * slice, slice, total_spp);
        nccl_check(nccl().AllGather(frame + (uint64_t)c.rank * slice, frame, slice * 4, ncclFloat, c.comm, c.stream), "ncclAllGather"); // in place
    } else { // PB2_REDUCE_ROOT, and the fallback for pixel counts the ranks do not divide
        if (c.rank == root) {
            if (!frame) throw std::runtime_error("pb2_comm_reduce_frames: frame buffer missing on the root");
            c.staging.ensure(n_pixels);
        }
        nccl_check(nccl().Reduce(sum, c.rank == root ? c.staging.ptr : nullptr, n_pixels * 4, ncclFloat, ncclSum, root, c.comm, c.stream), "ncclReduce");
        if (c.rank == root) finalize_sum_on(c.stream, c.staging.ptr, frame, n_pixels, total_spp);
    }
    PB2_CUDA(cudaEventRecord(c.t_end, c.stream));
    PB2_CUDA(cudaEventRecord(c.done, c.stream));
    // the sum buffer is being read: the scene's next k_accumulate waits for this reduction (wavefront.cu).  The event belongs
    // to the scene, so neither object's lifetime depends on the other's.
    if (!s.accumulate_gate) PB2_CUDA(cudaEventCreateWithFlags(&s.accumulate_gate, cudaEventDisableTiming));
    PB2_CUDA(cudaEventRecord(s.accumulate_gate, c.stream));
    s.gate_pending = true;
    ++c.reductions;
}
void comm_synchronize(Comm &c) { PB2_CUDA(cudaStreamSynchronize(c.stream)); }
// device time of the last reduction (collectives + finalize, from the moment the communicator's stream could start on it) and the
// bytes of one rank's sum buffer; synchronises the communicator's stream
void comm_last_reduction(Comm &c, float *ms, uint64_t *bytes) {
    PB2_CUDA(cudaStreamSynchronize(c.stream));
    float t = 0.f;
    if (c.reductions) PB2_CUDA(cudaEventElapsedTime(&t, c.t_begin, c.t_end));
    if (ms) *ms = t;
    if (bytes) *bytes = c.last_bytes;
}
int comm_nccl_version() {
    int v = 0;
    nccl_check(nccl().GetVersion(&v), "ncclGetVersion");
    return v;
}
}// namespace pb2
