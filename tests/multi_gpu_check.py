"""Multi-GPU parity check of the sharded path; one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

Every rank compares its shard / the all-reduced QGT with the CPU oracle (sizes the oracle finishes in seconds)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle.oracle import Oracle  # noqa: E402
from quantum_geometric_tensor_b200 import api, circuits as K  # noqa: E402

ALL_KINDS = [K.X, K.Y, K.Z, K.H, K.S, K.T, K.SDG, K.TDG, K.SX, K.RX, K.RY, K.RZ, K.PHASE, K.CNOT, K.CY, K.CZ,
             K.CH, K.SWAP, K.CRX, K.CRY, K.CRZ, K.ZZ]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = api.Context(local)
    uid = [api.Context.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.dist_init(rank, world, uid[0])
    orc = Oracle()
    failures = []

    def check(name, err, tol):
        ok = err < tol
        if rank == 0 or not ok:
            print(f"[rank {rank}] {name}: err={err:.2e} {'ok' if ok else 'FAIL'}", flush=True)
        if not ok:
            failures.append(name)

    # 1. sharded gate application incl. exchanges, identity layout restored
    for n, circ in ((14, K.hea_layers(14, 2)), (13, K.random_circuit(13, 80, 31, kinds=ALL_KINDS, share_params=True)),
                    (12, K.qaoa_maxcut(12, 2))):
        th = K.default_angles(max(1, circ.num_params), n)
        st = ctx.state(n).init(circ.initial_state)
        st.apply(circ, th)
        shard = st.download()
        nrm = st.norm2()
        st.close()
        ref = orc.apply(circ, th)
        dloc = ref.size // world
        check(f"state {circ.name}", float(np.abs(shard - ref[rank * dloc:(rank + 1) * dloc]).max()), 1e-13)
        check(f"norm {circ.name}", abs(nrm - 1.0), 1e-12)

    # 2. QGT: partial Grams + allreduce
    for circ in (K.config("c1"), K.random_circuit(11, 60, 77, kinds=ALL_KINDS, share_params=True), K.qaoa_maxcut(10, 3)):
        th = K.default_angles(max(1, circ.num_params), 5)
        q = ctx.qgt(circ, th)
        qo = orc.qgt(circ, th)
        check(f"qgt {circ.name}", float(np.abs(q - qo).max() / np.abs(qo).max()), 1e-10)
    # blocked schedule on shards
    ctx.set_option("max_slots", 7)
    circ = K.hea_layers(11, 2)
    th = K.default_angles(circ.num_params)
    q = ctx.qgt(circ, th)
    ctx.set_option("max_slots", 0)
    qo = orc.qgt(circ, th)
    check("qgt blocked hea n=11", float(np.abs(q - qo).max() / np.abs(qo).max()), 1e-10)

    # 3. 22 qubits, 3 full layers (P = 132): every rank qubit is exchanged in and out several times and the columns do
    #    not all fit per rank on 2 GPUs' worth of slots when capped: a 3 x 3 sub-block against the reference-generated
    #    golden (tests/golden_big, oracle/_ref), with both schedules
    z = np.load(os.path.join(ROOT, "tests", "golden_big", "hea_n22_l3_cols.npz"))
    circ = K.hea_layers(22, 3)
    th = z["theta"]
    cols = [int(x) for x in z["cols"]]
    for label, opts in (("gram schedule", {"fused": 0}), ("fused schedule, capped slots", {"fused": 1, "max_slots": 40}),
                        ("fused schedule, 9 slots (phi's images share one column, launches walked range by range)", {"fused": 1, "max_slots": 9})):
        for k_, v_ in opts.items():
            ctx.set_option(k_, v_)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        q = ctx.qgt(circ, th)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = ctx.stats()
        ctx.set_option("fused", -1)
        ctx.set_option("max_slots", 0)
        sub = q[np.ix_(cols, cols)]
        check(f"qgt n=22 L=3 sub-block metric ({label})", float(np.abs(sub.real - z["metric"]).max() / np.abs(z["metric"]).max()), 1e-10)
        check(f"qgt n=22 L=3 sub-block curvature ({label})", float(np.abs(-2 * sub.imag - z["curvature"]).max() / np.abs(z["curvature"]).max()), 1e-10)
        check(f"qgt n=22 hermitian ({label})", float(np.abs(q - q.conj().T).max()), 1e-12)
        if st["exchange_bytes"] <= 0:
            failures.append("no exchange happened at n=22")
        if rank == 0:
            print(f"[rank 0] n=22 P=132 on {world} GPUs ({label}): {dt * 1e3:.1f} ms, exchange bytes/rank {st['exchange_bytes']:.3e} "
                  f"({st['exchange_bytes'] / max(st['ms_exchange'], 1e-9) * 1e-6:.0f} GB/s), sweep {st['ms_sweep']:.1f} ms gram {st['ms_gram']:.1f} ms "
                  f"exchange {st['ms_exchange']:.1f} ms other {st['ms_other']:.1f} ms blocks {st['blocks']}", flush=True)

    # 4. adjoint energy gradient on sharded states: fused pair path (ansatz, exchanges inside one program) and the per-run
    #    path (QAOA cost layers, exchanges between programs), against the oracle's per-parameter derivative columns
    def ring(c, seed=3):
        rng = np.random.default_rng(seed)
        n = c.num_qubits
        c.edges = [(q, (q + 1) % n, float(rng.uniform(0.5, 1.5))) for q in range(n)]
        c.vertex_weights = [float(v) for v in rng.uniform(-1, 1, n)]
        return c
    for circ, want_fused in ((ring(K.hea_layers(14, 2)), 1), (K.qaoa_maxcut(12, 2), 0), (ring(K.hea_layers(13, 1)), 1)):
        th = K.default_angles(circ.num_params, 17)
        e, g = ctx.expectation_gradient(circ, th)
        st = ctx.stats()
        eo, go = orc.expectation_gradient(circ, th)
        check(f"gradient energy {circ.name}", abs(e - eo) / max(1.0, abs(eo)), 1e-10)
        check(f"gradient {circ.name} (fused={st['fused']})", float(np.abs(g - go).max() / max(1.0, np.abs(go).max())), 1e-10)
        if world <= 4 and st["fused"] != want_fused:
            failures.append(f"gradient path of {circ.name}: fused={st['fused']}")

    # 5. the generic collectives behind the ComputeBackendOps table (host buffers staged, NCCL underneath)
    x = np.arange(6, dtype=np.float64) + 10.0 * rank
    out = np.zeros(6)
    ctx.collective(1, x, out, 6, 1, 0)                                       # allreduce sum, f64
    check("collective allreduce", float(np.abs(out - (world * np.arange(6) + 10.0 * sum(range(world)))).max()), 1e-12)
    z = (np.arange(4) + 1j * rank).astype(np.complex64)
    zo = np.zeros(4, dtype=np.complex64)
    ctx.collective(1, z, zo, 4, 2, 0)                                        # allreduce sum, complex64 (pairs of f32)
    check("collective allreduce c64", float(np.abs(zo - (world * np.arange(4) + 1j * sum(range(world)))).max()), 1e-6)
    b = np.full(5, float(rank + 1))
    ctx.collective(0, None, b, 5, 1, 0, root=world - 1)                      # broadcast in place from the last rank
    check("collective broadcast", float(np.abs(b - world).max()), 0.5)
    ag = np.zeros(3 * world)
    ctx.collective(4, np.full(3, float(rank)), ag, 3, 1)                     # allgather
    check("collective allgather", float(np.abs(ag - np.repeat(np.arange(world), 3)).max()), 0.5)
    rs = np.zeros(2)
    ctx.collective(5, np.arange(2 * world, dtype=np.float64), rs, 2, 1, 0)   # reduce_scatter sum
    check("collective reduce_scatter", float(np.abs(rs - world * (np.arange(2) + 2 * rank)).max()), 1e-12)
    sc = np.zeros(2)
    ctx.collective(2, np.arange(2 * world, dtype=np.float64) if rank == 0 else np.zeros(2 * world), sc, 2, 1, 0, root=0)   # scatter
    check("collective scatter", float(np.abs(sc - (np.arange(2) + 2 * rank)).max()), 0.5)
    ga = np.zeros(2 * world)
    ctx.collective(3, np.full(2, float(rank)), ga, 2, 1, 0, root=0)          # gather
    if rank == 0:
        check("collective gather", float(np.abs(ga - np.repeat(np.arange(world), 2)).max()), 0.5)

    # 6. sampling and argmax on a sharded state: every rank uploads its slice of one seeded vector and must receive the
    #    inverse-CDF indices / the global winner of the whole vector
    for n in (16, 19):
        rng = np.random.default_rng(100 + n)
        full = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
        full[(5 << (n - 3)) + 77] *= 40.0                                   # a clear maximum, in a high shard
        full /= np.linalg.norm(full)
        L = (1 << n) // world
        s = ctx.state(n).upload(full[rank * L:(rank + 1) * L])
        u = np.concatenate([rng.uniform(0, 1, 300), [0.0, 1e-12, 0.5, 1 - 1e-12]])
        got = s.sample(u).astype(np.int64)
        cdf = np.cumsum(np.abs(full) ** 2)
        ref = np.searchsorted(cdf, u, side="right")
        gap = np.minimum(np.abs(cdf[np.minimum(ref, cdf.size - 1)] - u), np.abs(u - np.where(ref > 0, cdf[np.maximum(ref, 1) - 1], 0)))
        far = gap > 1e-10
        bad = int(np.count_nonzero(got[far] != np.minimum(ref[far], cdf.size - 1)))
        lo = np.where(got > 0, cdf[np.maximum(got, 1) - 1], 0.0)
        bad += int(np.count_nonzero(~((lo - 1e-11 <= u) & (u < cdf[got] + 1e-11))))
        check(f"sharded sampling n={n}", float(bad), 0.5)
        idx, p = s.argmax()
        check(f"sharded argmax n={n}", float(abs(idx - int(np.argmax(np.abs(full) ** 2))) + abs(p - np.max(np.abs(full) ** 2))), 1e-12)
        s.close()

    ctx.close()
    fl = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(fl)
    dist.destroy_process_group()
    if int(fl.item()):
        print(f"[rank {rank}] FAILURES: {failures}")
        sys.exit(1)
    if rank == 0:
        print("multi-GPU check passed")


if __name__ == "__main__":
    main()
