// dist.cu — multi-GPU path inside one box: one process per GPU, amplitudes sharded on the top
// log2(world) qubits (rank r owns the amplitudes whose high index bits equal r), NCCL over NVLink.
//
// Replaces the reference's qubit-count "sharding" and root-gather sync
// (src/quantum_geometric/distributed/quantum_distributed_operations.c:260-316, documented as defective in
// BASELINE.md §4 #13) and the serial ncclCommInitRank loop of core/multi_gpu_operations.c:46-116 (#14).
//
//   * gates acting diagonally on rank qubits (RZ/CZ/phases/controls/cost layer) cost no traffic: the sweep
//     kernel ORs the rank bits into the global index its masks are evaluated on;
//   * a gate acting non-diagonally on a rank qubit first swaps that qubit with the top local one: each rank
//     exchanges the contiguous half of its shard it does not keep with one peer (ncclSend/ncclRecv pair);
//     plan.cpp's mapper chooses which local qubit to evict and inserts these EXCHANGE pseudo-runs;
//   * the Gram rows are sharded like the amplitudes: per-rank partial (P+1)x(P+1) matrices are summed with a
//     single ncclAllReduce before Q = C - v v^H is formed.
#include <nccl.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "ctx.hpp"

#include <dlfcn.h>

namespace qgt {

// NCCL is bound at run time, not at link time: a host process that already carries its own libnccl.so.2
// (PyTorch bundles a newer one than the system's) keeps using it, and merely loading this library never
// drags a second NCCL into the process.
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.handle) return QGT_B200_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // the copy the host process already uses
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return fail(QGT_B200_ERR_HARDWARE, std::string("cannot load libnccl.so.2: ") + dlerror());
    NcclApi a;
    a.handle = h;
    bool ok = true;
    auto sym = [&](const char* name) { void* p = dlsym(h, name); if (!p) ok = false; return p; };
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
    a.Send = (decltype(a.Send))sym("ncclSend");
    a.Recv = (decltype(a.Recv))sym("ncclRecv");
    a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
    a.Broadcast = (decltype(a.Broadcast))sym("ncclBroadcast");
    a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
    a.ReduceScatter = (decltype(a.ReduceScatter))sym("ncclReduceScatter");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    if (!ok) return fail(QGT_B200_ERR_HARDWARE, "libnccl.so.2 lacks a required symbol");
    g_nccl = a;
    return QGT_B200_OK;
}

struct DistState {
    ncclComm_t comm = nullptr;
    DevBuf bounce;       // receive buffer of an exchange: half a column
    DevBuf red;          // small reduction buffer
    DevBuf seg_edges, seg_vw;
};

static int nccl_fail(ncclResult_t r, const char* what) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s", what, (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    return fail(QGT_B200_ERR_HARDWARE, buf);
}

void dist_shutdown(qgt_b200_ctx* c) {
    if (!c->dist) return;
    if (c->dist->comm) g_nccl.CommDestroy(c->dist->comm);
    c->dist->bounce.release(); c->dist->red.release(); c->dist->seg_edges.release(); c->dist->seg_vw.release();
    delete c->dist;
    c->dist = nullptr;
}

int dist_allreduce_device(qgt_b200_ctx* c, double* d_buf, size_t count) {
    if (c->world == 1) return QGT_B200_OK;
    ncclResult_t r = g_nccl.AllReduce(d_buf, d_buf, count, ncclDouble, ncclSum, c->dist->comm, c->stream);
    if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce");
    return QGT_B200_OK;
}

int dist_allreduce_host(qgt_b200_ctx* c, double* v, int n) {
    if (c->world == 1) return QGT_B200_OK;
    int rc = c->dist->red.reserve((size_t)n * sizeof(double));
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync(c->dist->red.ptr, v, n * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "allreduce staging");
    if ((rc = dist_allreduce_device(c, (double*)c->dist->red.ptr, (size_t)n))) return rc;
    e = cudaMemcpyAsync(v, c->dist->red.ptr, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "allreduce readback");
    return QGT_B200_OK;
}

// One grouped exchange: rank bits b_0 < b_1 < ... of `mask` trade places with the top k local qubits (b_i <-> nloc-k+i).
// The shard splits into 2^k blocks by its top k local bits; block j goes to the rank whose mask bits spell j and lands
// there at block index (my mask bits); the block whose index equals my own mask bits stays.  Out of place (src != dst):
// every block is received straight into its final position - no bounce buffer, no second pass over HBM.
int dist_exchange_multi(qgt_b200_ctx* c, const cplx* src, cplx* dst, uint64_t D, unsigned mask) {
    if (c->world == 1 || !c->dist) return fail(QGT_B200_ERR_INTERNAL, "exchange on an unsharded state");
    if (!mask || src == dst) return fail(QGT_B200_ERR_INTERNAL, "exchange needs a mask and distinct buffers");
    int bits[16], k = 0;
    for (int b = 0; b < 16; b++) if (mask >> b & 1u) bits[k++] = b;
    const uint64_t blk = D >> k;
    int mine = 0;
    for (int i = 0; i < k; i++) mine |= ((c->rank >> bits[i]) & 1) << i;
    ncclResult_t r = g_nccl.GroupStart();
    for (int j = 0; j < (1 << k) && r == ncclSuccess; j++) {
        if (j == mine) continue;
        int peer = c->rank;
        for (int i = 0; i < k; i++) peer = (peer & ~(1 << bits[i])) | (((j >> i) & 1) << bits[i]);
        r = g_nccl.Send(src + (uint64_t)j * blk, blk * 2, ncclDouble, peer, c->dist->comm, c->stream);
        if (r == ncclSuccess) r = g_nccl.Recv(dst + (uint64_t)j * blk, blk * 2, ncclDouble, peer, c->dist->comm, c->stream);
    }
    if (r == ncclSuccess) r = g_nccl.GroupEnd();
    if (r != ncclSuccess) return nccl_fail(r, "exchange send/recv");
    cudaError_t e = cudaMemcpyAsync(dst + (uint64_t)mine * blk, src + (uint64_t)mine * blk, blk * sizeof(cplx), cudaMemcpyDeviceToDevice, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "exchange local block");
    c->stats.exchange_bytes += (double)(((uint64_t)(1 << k) - 1) * blk * sizeof(cplx));
    return QGT_B200_OK;
}

// in-place form for a lone state (qgt_b200_apply_circuit): through a scratch column and back
int dist_exchange(qgt_b200_ctx* c, cplx* col, uint64_t D, unsigned mask) {
    if (c->world == 1 || !c->dist) return fail(QGT_B200_ERR_INTERNAL, "exchange on an unsharded state");
    int rc = c->dist->bounce.reserve(D * sizeof(cplx));
    if (rc) return rc;
    if ((rc = dist_exchange_multi(c, col, (cplx*)c->dist->bounce.ptr, D, mask))) return rc;
    cudaError_t e = cudaMemcpyAsync(col, c->dist->bounce.ptr, D * sizeof(cplx), cudaMemcpyDeviceToDevice, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "exchange copy");
    return QGT_B200_OK;
}

// cost tables with qubits renamed by each segment's logical -> physical map
int upload_segment_costs(qgt_b200_ctx* c, const qgt_b200_circuit& circ, const std::vector<MappedSegment>& segs) {
    c->seg_cost.clear();
    if (!circ.num_edges && !circ.vertex_weights) return QGT_B200_OK;
    const int n = circ.num_qubits;
    std::vector<QgtDevEdge> ed;
    std::vector<double> vw;
    for (const MappedSegment& sg : segs) {
        for (size_t k = 0; k < circ.num_edges; k++) {
            QgtDevEdge e;
            e.i = sg.phys_of_logical[circ.edges[k].i]; e.j = sg.phys_of_logical[circ.edges[k].j]; e.w = circ.edges[k].weight;
            ed.push_back(e);
        }
        if (circ.vertex_weights) {
            std::vector<double> w(n, 0.0);
            for (int q = 0; q < n; q++) w[sg.phys_of_logical[q]] = circ.vertex_weights[q];
            vw.insert(vw.end(), w.begin(), w.end());
        }
    }
    int rc;
    if ((rc = c->dist->seg_edges.reserve(std::max<size_t>(16, ed.size() * sizeof(QgtDevEdge))))) return rc;
    if ((rc = c->dist->seg_vw.reserve(std::max<size_t>(16, vw.size() * sizeof(double))))) return rc;
    cudaError_t e = cudaSuccess;
    if (!ed.empty()) e = cudaMemcpyAsync(c->dist->seg_edges.ptr, ed.data(), ed.size() * sizeof(QgtDevEdge), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && !vw.empty()) e = cudaMemcpyAsync(c->dist->seg_vw.ptr, vw.data(), vw.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "segment cost upload");
    for (size_t s = 0; s < segs.size(); s++) {
        QgtCostTable t;
        t.edges = circ.num_edges ? (const QgtDevEdge*)c->dist->seg_edges.ptr + s * circ.num_edges : nullptr;
        t.num_edges = (int)circ.num_edges;
        t.vertex_weights = circ.vertex_weights ? (const double*)c->dist->seg_vw.ptr + s * n : nullptr;
        t.n = n;
        c->seg_cost.push_back(t);
    }
    return QGT_B200_OK;
}

int dist_apply_circuit(qgt_b200_state* s, const qgt_b200_circuit* circ, const double* theta) {
    qgt_b200_ctx* c = s->ctx;
    CircuitPlan plan;
    std::vector<MappedSegment> segs;
    std::string err;
    int rc = build_plan_sharded(*circ, theta, c->opt, s->nloc, true, plan, segs, err);
    if (rc) return fail(rc, err);
    PlanImage img;
    stats_begin(c);
    if ((rc = upload_plan(c, *circ, plan, img))) return rc;
    if ((rc = upload_segment_costs(c, *circ, segs))) return rc;
    if ((rc = apply_plan_inplace(c, plan, s->d, s->D))) return rc;
    if ((rc = stats_end(c))) return rc;
    c->stats.num_runs = (int)plan.runs.size();
    c->stats.tile_qubits = plan.K;
    return QGT_B200_OK;
}

int dist_qgt(qgt_b200_ctx* c, const qgt_b200_circuit* circ, const double* theta,
             double* metric, double* berry, double* q_full, qgt_b200_state* psi_out) {
    const int n = circ->num_qubits, P = circ->num_params;
    int gbits = 0;
    while ((1 << gbits) < c->world) gbits++;
    const int nloc = n - gbits;
    if (nloc < 4) return fail(QGT_B200_ERR_INVALID_ARG, "state too small for this many ranks");
    const uint64_t D = (uint64_t)1 << nloc;
    CircuitPlan plan;
    std::vector<MappedSegment> segs;
    std::string err;
    int rc = build_plan_sharded(*circ, theta, c->opt, nloc, psi_out != nullptr, plan, segs, err);
    if (rc) return fail(rc, err);
    Program prog;
    // every rank must build the same program: agree on the smallest slot count
    size_t slots = workspace_slots(c, D, ((size_t)256 << 20) + D * sizeof(cplx));
    {
        int rc2 = c->dist->red.reserve(sizeof(double));
        if (rc2) return rc2;
        double v = (double)slots;
        cudaError_t e0 = cudaMemcpyAsync(c->dist->red.ptr, &v, sizeof v, cudaMemcpyHostToDevice, c->stream);
        if (e0 != cudaSuccess) return cuda_fail(e0, "slot staging");
        ncclResult_t r = g_nccl.AllReduce(c->dist->red.ptr, c->dist->red.ptr, 1, ncclDouble, ncclMin, c->dist->comm, c->stream);
        if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce(min)");
        e0 = cudaMemcpyAsync(&v, c->dist->red.ptr, sizeof v, cudaMemcpyDeviceToHost, c->stream);
        if (e0 == cudaSuccess) e0 = cudaStreamSynchronize(c->stream);
        if (e0 != cudaSuccess) return cuda_fail(e0, "slot readback");
        slots = (size_t)v;
        if (slots > 5) slots -= 1;               // one column is the spare the grouped exchanges receive into
    }
    const bool use_fused = choose_fused(c, plan, slots);
    if ((rc = use_fused ? build_fused_program(plan, slots, psi_out != nullptr, prog, err, c->fused_traj)
                        : build_qgt_program(plan, slots, psi_out != nullptr, prog, err))) return fail(rc, err);
    if ((rc = c->arena.reserve(((size_t)prog.num_slots + 1) * D * sizeof(cplx)))) return rc;     // + the spare column of the exchanges
    const size_t cm = (size_t)(P + 1) * (P + 1);
    if ((rc = c->cmat.reserve(std::max<size_t>(16, cm * sizeof(cplx))))) return rc;
    if ((rc = c->outbuf.reserve(std::max<size_t>(16, (size_t)P * P * 4 * sizeof(double))))) return rc;
    PlanImage img;
    stats_begin(c);
    if ((rc = upload_plan(c, *circ, plan, img))) return rc;
    if ((rc = upload_segment_costs(c, *circ, segs))) return rc;
    cudaError_t e = cudaMemsetAsync(c->cmat.ptr, 0, cm * sizeof(cplx), c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "memset");
    if ((rc = run_program(c, *circ, plan, prog, (cplx*)c->arena.ptr, D, (cplx*)c->cmat.ptr))) return rc;
    double* d_metric = (double*)c->outbuf.ptr;
    double* d_berry = d_metric + (size_t)P * P;
    cplx* d_q = (cplx*)(d_berry + (size_t)P * P);
    if (prog.fused) {
        // partial A and self transition matrices of the shards: two allreduces, then the host assembly (identical on every rank)
        if ((rc = fused_finish(c, plan, metric, berry, q_full))) return rc;
    } else {
        // partial Gram matrices of the shards -> one allreduce of (P+1)^2 complex numbers
        c->timer.begin(c->stream, 2, "gram allreduce");
        rc = dist_allreduce_device(c, (double*)c->cmat.ptr, cm * 2);
        c->timer.end(c->stream);
        if (rc) return rc;
        e = launch_finalize((const cplx*)c->cmat.ptr, P, d_metric, d_berry, d_q, c->stream);
        if (e != cudaSuccess) return cuda_fail(e, "finalize launch");
    }
    const size_t pp = (size_t)P * P;
    if (P > 0 && !prog.fused) {
        if (metric && e == cudaSuccess) e = cudaMemcpyAsync(metric, d_metric, pp * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (berry && e == cudaSuccess) e = cudaMemcpyAsync(berry, d_berry, pp * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (q_full && e == cudaSuccess) e = cudaMemcpyAsync(q_full, d_q, pp * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream);
    }
    if (psi_out && e == cudaSuccess)
        e = cudaMemcpyAsync(psi_out->d, (cplx*)c->arena.ptr + (size_t)c->psi_phys_slot * D, D * sizeof(cplx), cudaMemcpyDeviceToDevice, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "result copy");
    if ((rc = stats_end(c))) return rc;
    c->stats.num_runs = (int)plan.runs.size();
    c->stats.resident_columns = prog.resident;
    c->stats.blocks = prog.blocks;
    c->stats.tile_qubits = plan.K;
    return QGT_B200_OK;
}

}  // namespace qgt

extern "C" {

int qgt_b200_dist_unique_id(uint8_t id[QGT_B200_NCCL_ID_BYTES]) {
    if (!id) return qgt::fail(QGT_B200_ERR_INVALID_ARG, "id is NULL");
    static_assert(sizeof(ncclUniqueId) <= QGT_B200_NCCL_ID_BYTES, "unique id does not fit");
    int lrc = qgt::nccl_load();
    if (lrc) return lrc;
    ncclUniqueId uid;
    ncclResult_t r = qgt::g_nccl.GetUniqueId(&uid);
    if (r != ncclSuccess) return qgt::nccl_fail(r, "ncclGetUniqueId");
    std::memset(id, 0, QGT_B200_NCCL_ID_BYTES);
    std::memcpy(id, &uid, sizeof uid);
    return QGT_B200_OK;
}

int qgt_b200_dist_init(qgt_b200_ctx* c, int rank, int world, const uint8_t id[QGT_B200_NCCL_ID_BYTES]) {
    if (!c || !id) return qgt::fail(QGT_B200_ERR_INVALID_ARG, "ctx/id is NULL");
    if (world < 1 || (world & (world - 1)) || rank < 0 || rank >= world)
        return qgt::fail(QGT_B200_ERR_INVALID_ARG, "world must be a power of two and 0 <= rank < world");
    if (c->dist) return qgt::fail(-17 /* QGT_ERROR_ALREADY_INITIALIZED */, "communicator already initialised");
    cudaSetDevice(c->device);
    if (world == 1) { c->rank = 0; c->world = 1; return QGT_B200_OK; }
    int lrc = qgt::nccl_load();
    if (lrc) return lrc;
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof uid);
    qgt::DistState* d = new qgt::DistState();
    ncclResult_t r = qgt::g_nccl.CommInitRank(&d->comm, world, uid, rank);
    if (r != ncclSuccess) { delete d; return qgt::nccl_fail(r, "ncclCommInitRank"); }
    c->dist = d;
    c->rank = rank;
    c->world = world;
    return QGT_B200_OK;
}

int qgt_b200_dist_world(const qgt_b200_ctx* c, int* rank, int* world) {
    if (!c) return qgt::fail(QGT_B200_ERR_INVALID_ARG, "ctx is NULL");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    return QGT_B200_OK;
}

int qgt_b200_dist_barrier(qgt_b200_ctx* c) {
    if (!c) return qgt::fail(QGT_B200_ERR_INVALID_ARG, "ctx is NULL");
    double v = 0.0;
    return qgt::dist_allreduce_host(c, &v, 1);
}


// Generic collectives on caller buffers (host or device pointers): what the collective slots of the reference's
// ComputeBackendOps bind (compute_backend.h:215-319 of the reference; its CUDA backend, supercomputer/backends/
// compute_cuda.cu:1043-1330, stages host buffers the same way).  dtype / op use the reference's ComputeDataType /
// ComputeReduceOp numbering.  `count` is per rank for scatter / gather / allgather / reduce_scatter.  With one rank
// every call is the copy the reference's single-node path makes.
int qgt_b200_dist_collective(qgt_b200_ctx* c, int kind, const void* send, void* recv, size_t count, int dtype, int op, int root) {
    using namespace qgt;
    if (!c) return fail(QGT_B200_ERR_INVALID_ARG, "ctx is NULL");
    static const size_t esize[] = {4, 8, 8, 16, 4, 8, 1};
    if (dtype < 0 || dtype > 6) return fail(QGT_B200_ERR_INVALID_ARG, "unknown data type");
    if (kind < QGT_B200_COLL_BROADCAST || kind > QGT_B200_COLL_REDUCE_SCATTER) return fail(QGT_B200_ERR_INVALID_ARG, "unknown collective");
    if (kind == QGT_B200_COLL_BROADCAST) { if (!recv) recv = const_cast<void*>(send); send = recv; }
    if ((!send || !recv) && count) return fail(QGT_B200_ERR_INVALID_ARG, "send/recv is NULL");
    if (root < 0 || root >= c->world) return fail(QGT_B200_ERR_INVALID_ARG, "root out of range");
    const int W = c->world;
    size_t send_n = count, recv_n = count;                 // elements
    if (kind == QGT_B200_COLL_SCATTER || kind == QGT_B200_COLL_REDUCE_SCATTER) send_n = count * W;
    if (kind == QGT_B200_COLL_GATHER || kind == QGT_B200_COLL_ALLGATHER) recv_n = count * W;
    if (W == 1) {
        if (send != recv && count) {
            cudaSetDevice(c->device);
            cudaError_t e = cudaMemcpy(recv, send, count * esize[dtype], cudaMemcpyDefault);
            if (e != cudaSuccess) return cuda_fail(e, "single-rank collective copy");
        }
        return QGT_B200_OK;
    }
    if (!c->dist) return fail(QGT_B200_ERR_NOT_INIT, "communicator not initialised");
    cudaSetDevice(c->device);
    // complex types travel as pairs of their real type (sum / avg act component-wise; prod/min/max are refused)
    ncclDataType_t nt; size_t mul = 1;
    switch (dtype) {
        case 0: nt = ncclFloat; break;
        case 1: nt = ncclDouble; break;
        case 2: nt = ncclFloat; mul = 2; break;
        case 3: nt = ncclDouble; mul = 2; break;
        case 4: nt = ncclInt32; break;
        case 5: nt = ncclInt64; break;
        default: nt = ncclUint8; break;
    }
    ncclRedOp_t ro = ncclSum;
    if (kind == QGT_B200_COLL_ALLREDUCE || kind == QGT_B200_COLL_REDUCE_SCATTER) {
        switch (op) {
            case 0: ro = ncclSum; break;
            case 1: ro = ncclProd; break;
            case 2: ro = ncclMin; break;
            case 3: ro = ncclMax; break;
            case 4: ro = ncclAvg; break;
            default: return fail(QGT_B200_ERR_INVALID_ARG, "unknown reduction");
        }
        if (mul == 2 && (op == 1 || op == 2 || op == 3)) return fail(QGT_B200_ERR_UNSUPPORTED, "prod/min/max are not defined component-wise for complex data");
    }
    auto is_dev = [](const void* p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
    };
    const size_t sb = send_n * esize[dtype], rb = recv_n * esize[dtype];
    const bool sdev = is_dev(send), rdev = is_dev(recv);
    const void* ds = send; void* dr = recv;
    const size_t r_pad = (kind == QGT_B200_COLL_BROADCAST || rdev) ? 0 : ((rb + 255) & ~(size_t)255);
    const size_t need = r_pad + (sdev ? 0 : sb);
    if (need) {
        int rc = c->dist->red.reserve(need + 256);
        if (rc) return rc;
        char* base = (char*)c->dist->red.ptr;
        if (!rdev) dr = base;
        if (!sdev) {
            void* stage = base + r_pad;
            const bool sender = !(kind == QGT_B200_COLL_BROADCAST || kind == QGT_B200_COLL_SCATTER) || c->rank == root;
            if (sender) {
                cudaError_t e = cudaMemcpyAsync(stage, send, sb, cudaMemcpyHostToDevice, c->stream);
                if (e != cudaSuccess) return cuda_fail(e, "collective staging");
            }
            ds = stage;
            if (kind == QGT_B200_COLL_BROADCAST) dr = stage;       // in place
        }
    }
    ncclResult_t r = ncclSuccess;
    ncclComm_t comm = c->dist->comm;
    switch (kind) {
        case QGT_B200_COLL_BROADCAST: r = g_nccl.Broadcast(ds, dr, send_n * mul, nt, root, comm, c->stream); break;
        case QGT_B200_COLL_ALLREDUCE: r = g_nccl.AllReduce(ds, dr, send_n * mul, nt, ro, comm, c->stream); break;
        case QGT_B200_COLL_ALLGATHER: r = g_nccl.AllGather(ds, dr, count * mul, nt, comm, c->stream); break;
        case QGT_B200_COLL_REDUCE_SCATTER: r = g_nccl.ReduceScatter(ds, dr, count * mul, nt, ro, comm, c->stream); break;
        case QGT_B200_COLL_SCATTER:
            g_nccl.GroupStart();
            if (c->rank == root)
                for (int p = 0; p < W && r == ncclSuccess; p++)
                    r = g_nccl.Send((const char*)ds + (size_t)p * count * esize[dtype], count * mul, nt, p, comm, c->stream);
            if (r == ncclSuccess) r = g_nccl.Recv(dr, count * mul, nt, root, comm, c->stream);
            g_nccl.GroupEnd();
            break;
        default:   // gather
            g_nccl.GroupStart();
            r = g_nccl.Send(ds, count * mul, nt, root, comm, c->stream);
            if (c->rank == root)
                for (int p = 0; p < W && r == ncclSuccess; p++)
                    r = g_nccl.Recv((char*)dr + (size_t)p * count * esize[dtype], count * mul, nt, p, comm, c->stream);
            g_nccl.GroupEnd();
            break;
    }
    if (r != ncclSuccess) return nccl_fail(r, "collective");
    cudaError_t e = cudaSuccess;
    const bool receiver = kind != QGT_B200_COLL_GATHER || c->rank == root;
    if (!rdev && receiver) e = cudaMemcpyAsync(recv, dr, rb, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    return e == cudaSuccess ? QGT_B200_OK : cuda_fail(e, "collective");
}

}  // extern "C"
