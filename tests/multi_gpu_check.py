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

    # 3. a size no single check needs the oracle for: norm preservation + Hermiticity at 2^26 amplitudes per... n = 26
    circ = K.hea(26, 24)
    th = K.default_angles(24)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    q = ctx.qgt(circ, th)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = ctx.stats()
    check("qgt n=26 hermitian", float(np.abs(q - q.conj().T).max()), 1e-12)
    check("qgt n=26 first-layer diag", float(np.abs(np.diag(q.real)[:24] - 0.25).max()), 1e-12)
    if rank == 0:
        print(f"[rank 0] n=26 P=24 on {world} GPUs: {dt * 1e3:.1f} ms, exchange bytes/rank {st['exchange_bytes']:.3e}, "
              f"sweep {st['ms_sweep']:.1f} ms gram {st['ms_gram']:.1f} ms other {st['ms_other']:.1f} ms", flush=True)

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
