"""CPU tests of the multi-GPU host logic: logical->physical qubit mapping, EXCHANGE pseudo-runs, per-segment
cost tables, partial-Gram reduction — first with all ranks simulated in one process, then with two real
processes over torch.distributed/gloo exchanging half shards (what NCCL send/recv does on the GPUs)."""
import os
import socket
import sys

import numpy as np
import pytest

from quantum_geometric_tensor_b200 import api, circuits as K

import plan_interp as pi

ALL_KINDS = [K.X, K.Y, K.Z, K.H, K.S, K.T, K.SDG, K.TDG, K.SX, K.RX, K.RY, K.RZ, K.PHASE, K.CNOT, K.CY, K.CZ,
             K.CH, K.SWAP, K.CRX, K.CRY, K.CRZ, K.ZZ]


@pytest.mark.parametrize("world,slots", [(2, 64), (4, 64), (8, 64), (2, 6), (4, 9)])
def test_sharded_hea_qgt_and_state(oracle, world, slots):
    c = K.hea_layers(7, 2)
    th = K.default_angles(c.num_params)
    plan = api.plan_dump_sharded(c, th, world, True, tile_qubits=4, reg_qubits=2, column_slots=slots)
    assert any(r["exchange"] >= 0 for r in plan["runs"]), "a hardware-efficient ansatz needs qubit exchanges"
    Q, psi = pi.run_program_sharded(plan, c, world)
    assert np.abs(Q - oracle.qgt(c, th)).max() < 1e-12
    assert np.abs(psi - oracle.apply(c, th)).max() < 1e-13      # identity layout restored at the end


@pytest.mark.parametrize("seed,world", [(0, 2), (1, 4), (2, 2), (3, 8)])
def test_sharded_random_circuits(oracle, seed, world):
    c = K.random_circuit(7, 40, 900 + seed, kinds=ALL_KINDS, share_params=True)
    th = K.default_angles(max(1, c.num_params), seed)
    plan = api.plan_dump_sharded(c, th, world, True, tile_qubits=4, reg_qubits=2, column_slots=200)
    Q, psi = pi.run_program_sharded(plan, c, world)
    assert np.abs(psi - oracle.apply(c, th)).max() < 1e-13
    assert np.abs(Q - oracle.qgt(c, th)).max() < 1e-12


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_qaoa_cost_tables_follow_the_qubit_map(oracle, world):
    c = K.qaoa_maxcut(8, 2)
    c.vertex_weights = list(np.linspace(-0.4, 0.4, 8))
    th = K.default_angles(c.num_params, 3)
    plan = api.plan_dump_sharded(c, th, world, True, tile_qubits=4, reg_qubits=2, column_slots=64)
    Q, psi = pi.run_program_sharded(plan, c, world)
    assert np.abs(psi - oracle.apply(c, th)).max() < 1e-13
    assert np.abs(Q - oracle.qgt(c, th)).max() < 1e-12


def test_exchanges_are_batched_over_rank_bits(monkeypatch):
    """A hardware-efficient ansatz on 8 ranks: the three rank qubits come in with ONE grouped exchange per visit
    instead of three pairwise ones (fewer exchanges and 7/8 instead of 3/2 of the shard moved)."""
    c = K.hea_layers(9, 3)
    th = K.default_angles(c.num_params)
    def exchanged(plan):
        ex = [r["exchange_mask"] for r in plan["runs"] if r["exchange"] >= 0]
        vol = sum(1.0 - 0.5 ** bin(m).count("1") for m in ex)
        return len(ex), vol
    monkeypatch.setenv("QGT_B200_BATCH_EXCHANGES", "0")
    n1, v1 = exchanged(api.plan_dump_sharded(c, th, 8, True, tile_qubits=4, reg_qubits=2))
    monkeypatch.setenv("QGT_B200_BATCH_EXCHANGES", "1")
    plan = api.plan_dump_sharded(c, th, 8, True, tile_qubits=4, reg_qubits=2)
    n2, v2 = exchanged(plan)
    assert any(bin(r["exchange_mask"]).count("1") > 1 for r in plan["runs"])
    assert n2 < n1 and v2 < v1


def test_diagonal_gates_on_rank_qubits_need_no_exchange():
    c = K.Circuit(6)
    for q in range(6):
        c.add(K.H, q) if q < 4 else None
    c.add(K.RZ, 5, -1, 0)
    c.add(K.CZ, 4, 5)
    c.add(K.CNOT, 1, 5)          # control on a rank qubit, target local
    c.add(K.ZZ, 5, 2, 1)
    plan = api.plan_dump_sharded(c, np.array([0.3, 0.4]), 4, False)
    assert all(r["exchange"] < 0 for r in plan["runs"])


def _gloo_worker(rank, world, port, result_path):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import numpy as np
    import torch
    from oracle.oracle import Oracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    c = K.hea_layers(7, 2)
    th = K.default_angles(c.num_params)
    plan = api.plan_dump_sharded(c, th, world, True, tile_qubits=4, reg_qubits=2, column_slots=10)
    prog, P, nloc = plan["program"], plan["P"], plan["nloc"]
    tabs = pi.segment_cost_tables(plan, c)
    slots = [np.zeros(1 << nloc, dtype=np.complex128) for _ in range(prog["slots"])]
    Cm = np.zeros((P + 1, P + 1), dtype=np.complex128)
    for ins in prog["instrs"]:
        k = ins["k"]
        if k == "init":
            slots[ins["dst"]] = pi.initial_shard(c, rank, nloc)
        elif k == "copy":
            slots[ins["dst"]] = slots[ins["src"]].copy()
        elif k == "sweep":
            run = plan["runs"][ins["run"]]
            if run["exchange"] >= 0:
                # grouped exchange exactly as dist.cu's dist_exchange_multi: block j of my shard goes to the rank whose
                # mask bits spell j, the block I receive from it lands at index j; my own block stays
                bits = [b for b in range(8) if (run["exchange_mask"] >> b) & 1]
                k = len(bits)
                blk = (1 << nloc) >> k
                mine = sum(((rank >> b) & 1) << i for i, b in enumerate(bits))
                for (src, dst, ovr, acc, extra) in ins["cols"]:
                    col = slots[dst]
                    new = col.copy()
                    reqs, recvs = [], []
                    for j in range(1 << k):
                        if j == mine:
                            continue
                        peer = rank
                        for i, b in enumerate(bits):
                            peer = (peer & ~(1 << b)) | (((j >> i) & 1) << b)
                        send = torch.from_numpy(np.ascontiguousarray(col[j * blk:(j + 1) * blk]).view(np.float64).copy())
                        recv = torch.empty_like(send)
                        reqs += [dist.isend(send, peer), dist.irecv(recv, peer)]
                        recvs.append((j, recv))
                    for q in reqs:
                        q.wait()
                    for j, recv in recvs:
                        new[j * blk:(j + 1) * blk] = recv.numpy().view(np.complex128)
                    slots[dst] = new
                continue
            res = [(dst, acc, sum(pi.sweep_shard(plan, c, ins["run"], slots[src], rank, o, tabs) for o in [ovr] + list(extra)))
                   for (src, dst, ovr, acc, extra) in ins["cols"]]
            for dst, acc, v in res:
                slots[dst] = slots[dst] + v if acc else v
        elif k == "gram":
            for sa, ia in zip(ins["a"], ins["aid"]):
                for sb, ib in zip(ins["b"], ins["bid"]):
                    Cm[ia, ib] = np.vdot(slots[sa], slots[sb])
                    Cm[ib, ia] = np.conj(Cm[ia, ib])
    t = torch.from_numpy(Cm.view(np.float64).copy())
    dist.all_reduce(t)                                           # the P x P Gram allreduce
    Cm = t.numpy().view(np.complex128).reshape(P + 1, P + 1)
    v = Cm[:P, P]
    Q = Cm[:P, :P] - np.outer(v, v.conj())
    orc = Oracle()
    err_q = float(np.abs(Q - orc.qgt(c, th)).max())
    psi_ref = orc.apply(c, th)
    lo = rank << nloc
    err_psi = float(np.abs(slots[prog["psi"]] - psi_ref[lo:lo + (1 << nloc)]).max())
    with open(f"{result_path}.{rank}", "w") as f:
        f.write(f"{err_q} {err_psi}")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_multi_process_gloo_exchange_and_allreduce(tmp_path, world):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    res = str(tmp_path / "res")
    mp.spawn(_gloo_worker, args=(world, port, res), nprocs=world, join=True)
    for r in range(world):
        err_q, err_psi = [float(x) for x in open(f"{res}.{r}").read().split()]
        assert err_q < 1e-12 and err_psi < 1e-13
