// dist.cu — multi-GPU path: one process per GPU, amplitudes sharded on the top log2(world) qubits.
// (placeholder until the sharded executor lands: world == 1 everywhere)
#include "ctx.hpp"

namespace qgt {

struct DistState { int dummy; };

void dist_shutdown(qgt_b200_ctx* c) { delete c->dist; c->dist = nullptr; }

int dist_allreduce_host(qgt_b200_ctx* c, double* v, int n) {
    (void)v; (void)n;
    if (c->world == 1) return QGT_B200_OK;
    return fail(QGT_B200_ERR_UNSUPPORTED, "multi-GPU reductions not built");
}

int dist_apply_circuit(qgt_b200_state*, const qgt_b200_circuit*, const double*) {
    return fail(QGT_B200_ERR_UNSUPPORTED, "multi-GPU circuit application not built");
}

int dist_qgt(qgt_b200_ctx*, const qgt_b200_circuit*, const double*, double*, double*, double*, qgt_b200_state*) {
    return fail(QGT_B200_ERR_UNSUPPORTED, "multi-GPU QGT not built");
}

}  // namespace qgt

extern "C" {
int qgt_b200_dist_unique_id(uint8_t*) { return qgt::fail(QGT_B200_ERR_UNSUPPORTED, "multi-GPU not built"); }
int qgt_b200_dist_init(qgt_b200_ctx*, int, int, const uint8_t*) { return qgt::fail(QGT_B200_ERR_UNSUPPORTED, "multi-GPU not built"); }
int qgt_b200_dist_world(const qgt_b200_ctx* c, int* rank, int* world) {
    if (!c) return qgt::fail(QGT_B200_ERR_INVALID_ARG, "ctx is NULL");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    return QGT_B200_OK;
}
int qgt_b200_dist_export_ipc(qgt_b200_ctx*, size_t, uint8_t*) { return qgt::fail(QGT_B200_ERR_UNSUPPORTED, "multi-GPU not built"); }
int qgt_b200_dist_import_ipc(qgt_b200_ctx*, const uint8_t*) { return qgt::fail(QGT_B200_ERR_UNSUPPORTED, "multi-GPU not built"); }
int qgt_b200_dist_barrier(qgt_b200_ctx*) { return QGT_B200_OK; }
}
