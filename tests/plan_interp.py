"""numpy interpreter of the planner's JSON dump (qgt_b200_plan_dump).   TEST INFRASTRUCTURE ONLY.

Executes the fused-run plan and the column schedule exactly as the CUDA executor would (same op
order, same derivative overrides, same Gram calls) so the host-side planning logic can be checked
against the oracle without a GPU.
"""
from __future__ import annotations

import numpy as np


def _cost_energy(circ, idx):
    e = np.zeros(idx.shape, dtype=np.float64)
    for (i, j, w) in circ.edges:
        e += w * (((idx >> i) ^ (idx >> j)) & 1)
    if circ.vertex_weights is not None:
        for q in range(circ.num_qubits):
            e += circ.vertex_weights[q] * (1 - 2 * ((idx >> q) & 1))
    return e


def apply_op(state: np.ndarray, op: dict, circ) -> np.ndarray:
    n = circ.num_qubits
    idx = np.arange(1 << n, dtype=np.uint64)
    cm = np.uint64(op["cmask"])
    ctrl_ok = (idx & cm) == cm
    m = np.array(op["m"], dtype=np.float64)
    t = op["type"]
    out = state.copy()
    zero_fail = bool(op["flags"] & 1)
    if t in (0, 1, 2, 3):
        tb = np.uint64(1) << np.uint64(op["target"])
        lo = (idx & tb) == 0
        i0 = idx[lo & ctrl_ok]
        i1 = i0 | tb
        a0, a1 = state[i0], state[i1]
        if t == 3:
            out[i0], out[i1] = a1, a0
        else:
            M = (m[0::2] + 1j * m[1::2]).reshape(2, 2)
            out[i0] = M[0, 0] * a0 + M[0, 1] * a1
            out[i1] = M[1, 0] * a0 + M[1, 1] * a1
        if zero_fail:
            out[~ctrl_ok] = 0
    elif t == 4:
        pm = np.uint64(op["pmask"])
        par = np.zeros(idx.shape, dtype=np.uint64)
        v = idx & pm
        for b in range(n):
            par ^= (v >> np.uint64(b)) & np.uint64(1)
        d = np.where(par == 1, m[2] + 1j * m[3], m[0] + 1j * m[1])
        out = np.where(ctrl_ok, state * d, 0 if zero_fail else state)
    elif t == 5:
        e = _cost_energy(circ, idx.astype(np.int64))
        out = state * np.exp(-1j * m[0] * e)
        if op["flags"] & 2:
            out = out * (-1j * m[1] * e)
    else:
        raise ValueError(t)
    return out


def run_sweep(plan: dict, run_idx: int, src: np.ndarray, ovr: int, circ) -> np.ndarray:
    run = plan["runs"][run_idx]
    v = src
    for i, op in enumerate(run["ops"]):
        v = apply_op(v, run["dops"][str(i)] if i == ovr else op, circ)
    return v


def initial_state(circ) -> np.ndarray:
    dim = 1 << circ.num_qubits
    if circ.initial_state == 1:
        return np.full(dim, 1.0 / np.sqrt(dim), dtype=np.complex128)
    v = np.zeros(dim, dtype=np.complex128)
    v[0] = 1
    return v


def apply_plan(plan: dict, circ, state: np.ndarray) -> np.ndarray:
    for r in range(len(plan["runs"])):
        state = run_sweep(plan, r, state, -1, circ)
    return state


def run_program(plan: dict, circ):
    """Returns (Q, psi or None, counters)."""
    prog = plan["program"]
    P = plan["P"]
    dim = 1 << circ.num_qubits
    slots = [np.zeros(dim, dtype=np.complex128) for _ in range(prog["slots"])]
    Cm = np.zeros((P + 1, P + 1), dtype=np.complex128)
    seen = np.zeros((P + 1, P + 1), dtype=np.int32)
    counters = {"sweep_cols": 0, "gram_pairs": 0, "launches": 0}
    for ins in prog["instrs"]:
        k = ins["k"]
        if k == "init":
            slots[ins["dst"]] = initial_state(circ)
        elif k == "copy":
            slots[ins["dst"]] = slots[ins["src"]].copy()
        elif k == "sweep":
            counters["launches"] += 1
            dsts = [c[1] for c in ins["cols"]]
            assert len(set(dsts)) == len(dsts), "two items of one launch share a destination"
            srcs_oop = {c[0] for c in ins["cols"] if c[0] != c[1]}
            assert not (srcs_oop & set(dsts)), "a launch reads a column another item of it writes"
            results = []
            for (src, dst, ovr, acc, extra) in ins["cols"]:
                v = run_sweep(plan, ins["run"], slots[src], ovr, circ)
                for e in extra:                      # product rule inside one dense stage: the item sums the variants
                    v = v + run_sweep(plan, ins["run"], slots[src], e, circ)
                results.append((dst, acc, v))
                counters["sweep_cols"] += 1
            for dst, acc, v in results:
                slots[dst] = slots[dst] + v if acc else v
        elif k == "gram":
            for sa, ia in zip(ins["a"], ins["aid"]):
                for sb, ib in zip(ins["b"], ins["bid"]):
                    val = np.vdot(slots[sa], slots[sb])
                    Cm[ia, ib] = val
                    Cm[ib, ia] = np.conj(val)
                    seen[ia, ib] += 1
                    seen[ib, ia] += 1
                    counters["gram_pairs"] += 1
        else:
            raise ValueError(k)
    v = Cm[:P, P]
    Q = Cm[:P, :P] - np.outer(v, v.conj())
    psi = slots[prog["psi"]] if prog["psi_final"] else None
    return Q, psi, counters, seen


# ---- sharded states (one shard per rank) -----------------------------------------------------------------
def _apply_op_shard(shard, op, circ, gidx, nloc, edges, vweights):
    """apply_op on one shard: gidx = global amplitude indices of the shard's entries (rank bits included)."""
    cm = np.uint64(op["cmask"])
    ctrl_ok = (gidx & cm) == cm
    m = np.array(op["m"], dtype=np.float64)
    t = op["type"]
    out = shard.copy()
    zero_fail = bool(op["flags"] & 1)
    lidx = np.arange(shard.size, dtype=np.uint64)
    if t in (0, 1, 2, 3):
        assert op["target"] < nloc, "non-diagonal op on a rank qubit"
        tb = np.uint64(1) << np.uint64(op["target"])
        lo = (lidx & tb) == 0
        i0 = lidx[lo & ctrl_ok]
        i1 = i0 | tb
        a0, a1 = shard[i0], shard[i1]
        if t == 3:
            out[i0], out[i1] = a1, a0
        else:
            M = (m[0::2] + 1j * m[1::2]).reshape(2, 2)
            out[i0] = M[0, 0] * a0 + M[0, 1] * a1
            out[i1] = M[1, 0] * a0 + M[1, 1] * a1
        if zero_fail:
            out[~ctrl_ok] = 0
    elif t == 4:
        pm = np.uint64(op["pmask"])
        par = np.zeros(gidx.shape, dtype=np.uint64)
        v = gidx & pm
        for b in range(circ.num_qubits):
            par ^= (v >> np.uint64(b)) & np.uint64(1)
        d = np.where(par == 1, m[2] + 1j * m[3], m[0] + 1j * m[1])
        out = np.where(ctrl_ok, shard * d, 0 if zero_fail else shard)
    elif t == 5:
        g = gidx.astype(np.int64)
        e = np.zeros(g.shape)
        for (i, j, w) in edges:
            e += w * (((g >> i) ^ (g >> j)) & 1)
        if vweights is not None:
            for q in range(circ.num_qubits):
                e += vweights[q] * (1 - 2 * ((g >> q) & 1))
        out = shard * np.exp(-1j * m[0] * e)
        if op["flags"] & 2:
            out = out * (-1j * m[1] * e)
    return out


def segment_cost_tables(plan, circ):
    """edge lists / vertex weights renamed by each segment's logical -> physical qubit map"""
    tabs = []
    for seg in plan["segments"]:
        phys = seg["phys"]
        edges = [(phys[i], phys[j], w) for (i, j, w) in circ.edges]
        vw = None
        if circ.vertex_weights is not None:
            vw = [0.0] * circ.num_qubits
            for q in range(circ.num_qubits):
                vw[phys[q]] = circ.vertex_weights[q]
        tabs.append((edges, vw))
    return tabs


def sweep_shard(plan, circ, run_idx, shard, rank, ovr, tabs):
    run = plan["runs"][run_idx]
    nloc = plan["nloc"]
    gidx = (np.uint64(rank) << np.uint64(nloc)) | np.arange(shard.size, dtype=np.uint64)
    edges, vw = tabs[run["segment"]]
    v = shard
    for i, op in enumerate(run["ops"]):
        v = _apply_op_shard(v, run["dops"][str(i)] if i == ovr else op, circ, gidx, nloc, edges, vw)
    return v


def exchange_all_ranks(cols, mask):
    """cols[rank] = shard (modified in place).  Rank bits b_0 < b_1 < ... of `mask` trade places with the top k local
    qubits (b_i <-> nloc-k+i): one grouped all-to-all (dist.cu: dist_exchange_multi)."""
    world = len(cols)
    dloc = cols[0].size
    nloc = dloc.bit_length() - 1
    bits = [b for b in range(16) if (mask >> b) & 1]
    k = len(bits)
    full = np.concatenate(cols)
    idx = np.arange(full.size, dtype=np.int64)
    src = idx.copy()
    for i, b in enumerate(bits):                      # new[idx] = old[idx with bit (nloc+b) and bit (nloc-k+i) swapped]
        hi, lo = nloc + b, nloc - k + i
        bh, bl = (src >> hi) & 1, (src >> lo) & 1
        src = src ^ ((bh ^ bl) << hi) ^ ((bh ^ bl) << lo)
    out = full[src]
    for r in range(world):
        cols[r][:] = out[r * dloc:(r + 1) * dloc]


def initial_shard(circ, rank, nloc):
    dim = 1 << nloc
    if circ.initial_state == 1:
        return np.full(dim, 2.0 ** (-0.5 * circ.num_qubits), dtype=np.complex128)
    v = np.zeros(dim, dtype=np.complex128)
    if rank == 0:
        v[0] = 1
    return v


def run_program_sharded(plan, circ, world):
    """all ranks simulated in one process; returns (Q, full psi or None)"""
    prog = plan["program"]
    P, nloc = plan["P"], plan["nloc"]
    tabs = segment_cost_tables(plan, circ)
    slots = [[np.zeros(1 << nloc, dtype=np.complex128) for _ in range(prog["slots"])] for _ in range(world)]
    Cm = np.zeros((P + 1, P + 1), dtype=np.complex128)
    for ins in prog["instrs"]:
        k = ins["k"]
        if k == "init":
            for r in range(world):
                slots[r][ins["dst"]] = initial_shard(circ, r, nloc)
        elif k == "copy":
            for r in range(world):
                slots[r][ins["dst"]] = slots[r][ins["src"]].copy()
        elif k == "sweep":
            run = plan["runs"][ins["run"]]
            if run["exchange"] >= 0:
                for (src, dst, ovr, acc, extra) in ins["cols"]:
                    assert src == dst and ovr < 0 and not acc
                    cols = [slots[r][dst] for r in range(world)]
                    exchange_all_ranks(cols, run["exchange_mask"])
                continue
            for r in range(world):
                res = [(dst, acc, sum(sweep_shard(plan, circ, ins["run"], slots[r][src], r, o, tabs) for o in [ovr] + list(extra)))
                       for (src, dst, ovr, acc, extra) in ins["cols"]]
                for dst, acc, v in res:
                    slots[r][dst] = slots[r][dst] + v if acc else v
        elif k == "gram":
            for sa, ia in zip(ins["a"], ins["aid"]):
                for sb, ib in zip(ins["b"], ins["bid"]):
                    val = sum(np.vdot(slots[r][sa], slots[r][sb]) for r in range(world))   # the allreduce
                    Cm[ia, ib] = val
                    Cm[ib, ia] = np.conj(val)
    v = Cm[:P, P]
    Q = Cm[:P, :P] - np.outer(v, v.conj())
    psi = np.concatenate([slots[r][prog["psi"]] for r in range(world)]) if prog["psi_final"] else None
    return Q, psi


# ---- fused schedule (plan.hpp: build_fused_program) --------------------------------------------------------------
def _stage_of_op(run):
    """lowered-op index -> run-relative index of the dense stage holding it"""
    m, sidx = {}, 0
    for sp in run["subs"]:
        for st in sp.get("stages", []):
            for o in st["ops"]:
                m[o] = sidx
            sidx += 1
    return m


def _fused_launch(plan, circ, ins, slots_by_rank, apply_fn, A, D, v):
    """One INSTR_FUSED.  slots_by_rank[r] = list of this rank's columns; apply_fn(rank, state, op) applies a lowered op.
    Works at op granularity (no stage matrices): T[mu][b] = <lambda after op b | d(op b) phi before op b>, taken only
    for occurrences in stages >= rho_from, exactly the terms the device contracts from its transition matrices."""
    run = plan["runs"][ins["run"]]
    ops, dops = run["ops"], run["dops"]
    stage_of = _stage_of_op(run)
    world = len(slots_by_rank)
    results = []
    phi_out = []
    for (src, dst, ovr, acc, extra, cid, rho_from, self_, phi_dst) in ins["cols"]:
        ovrs = [ovr] + list(extra) if ovr >= 0 else [None]
        if phi_dst >= 0:                   # pair mode: this item also writes the advanced phi
            adv = []
            for r in range(world):
                st_ = slots_by_rank[r][ins["phi"]]
                for op in ops:
                    st_ = apply_fn(r, st_, op)
                adv.append(st_)
            phi_out.append((phi_dst, adv))
        total = [0 for _ in range(world)]
        for o in ovrs:
            for r in range(world):
                lam = slots_by_rank[r][src].copy()
                phi = slots_by_rank[r][ins["phi"]].copy()
                for i, op in enumerate(ops):
                    phi_prev = phi
                    phi = apply_fn(r, phi, op)
                    lam = apply_fn(r, lam, dops[str(i)] if i == o else op)
                    if op["param"] >= 0 and not self_:
                        assert i in stage_of, "a parameterised op sits outside every dense stage"
                        if stage_of[i] >= rho_from:
                            A[cid, op["param"]] += np.vdot(lam, apply_fn(r, phi_prev, dops[str(i)]))
                total[r] = total[r] + lam
        results.append((dst, acc, total))
        if self_ and rho_from == 0:
            # pairs inside one stage, the diagonal and the projections: from phi alone
            for r in range(world):
                traj = [slots_by_rank[r][ins["phi"]]]
                for op in ops:
                    traj.append(apply_fn(r, traj[-1], op))
                sidx = 0
                for sp in run["subs"]:
                    for st in sp.get("stages", []):
                        if st["params"]:
                            e = max(st["ops"])
                            cols = {}
                            for a in st["ops"]:
                                p = ops[a]["param"]
                                if p < 0:
                                    continue
                                x = apply_fn(r, traj[a], dops[str(a)])
                                for j in range(a + 1, e + 1):
                                    x = apply_fn(r, x, ops[j])
                                cols[p] = cols.get(p, 0) + x
                            for mu, cm in cols.items():
                                v[mu] += np.vdot(cm, traj[e + 1])
                                for nu, cn in cols.items():
                                    D[mu, nu] += np.vdot(cm, cn)
                        sidx += 1
    dsts = [d for d, _, _ in results]
    assert len(set(dsts)) == len(dsts), "two items of one launch share a destination"
    assert ins["phi"] not in dsts, "a fused launch overwrites the phi it reads"
    assert len(phi_out) <= 1 and (not phi_out or len(ins["cols"]) == 1), "pair mode: one item per launch"
    for dst, acc, total in results:
        for r in range(world):
            slots_by_rank[r][dst] = slots_by_rank[r][dst] + total[r] if acc else total[r]
    for dst, adv in phi_out:
        for r in range(world):
            slots_by_rank[r][dst] = adv[r]


def run_program_fused(plan: dict, circ, world: int = 1):
    """Returns (Q, psi or None, counters) for a fused program, all ranks simulated in one process."""
    prog = plan["program"]
    assert prog["fused"] == 1
    P, nloc = plan["P"], plan["nloc"]
    dim = 1 << nloc
    tabs = segment_cost_tables(plan, circ) if world > 1 else None
    slots = [[np.zeros(dim, dtype=np.complex128) for _ in range(prog["slots"])] for _ in range(world)]
    A = np.zeros((P, P), dtype=np.complex128)
    D = np.zeros((P, P), dtype=np.complex128)
    v = np.zeros(P, dtype=np.complex128)
    counters = {"column_passes": 0, "launches": 0}

    def apply_fn_for(run):
        if world == 1:
            return lambda r, st, op: apply_op(st, op, circ)
        edges, vw = tabs[run["segment"]]
        def f(r, st, op):
            gidx = (np.uint64(r) << np.uint64(nloc)) | np.arange(st.size, dtype=np.uint64)
            return _apply_op_shard(st, op, circ, gidx, nloc, edges, vw)
        return f

    for ins in prog["instrs"]:
        k = ins["k"]
        if k == "init":
            for r in range(world):
                slots[r][ins["dst"]] = initial_shard(circ, r, nloc) if world > 1 else initial_state(circ)
        elif k == "copy":
            for r in range(world):
                slots[r][ins["dst"]] = slots[r][ins["src"]].copy()
        elif k == "sweep":
            run = plan["runs"][ins["run"]]
            counters["launches"] += 1
            if run["exchange"] >= 0:
                for (src, dst, ovr, acc, extra) in ins["cols"]:
                    assert src == dst and ovr < 0 and not acc
                    exchange_all_ranks([slots[r][dst] for r in range(world)], run["exchange_mask"])
                continue
            f = apply_fn_for(run)
            for (src, dst, ovr, acc, extra) in ins["cols"]:
                assert ovr < 0 and not acc
                counters["column_passes"] += 1
                for r in range(world):
                    st = slots[r][src]
                    for op in run["ops"]:
                        st = f(r, st, op)
                    slots[r][dst] = st
        elif k == "fused":
            counters["launches"] += 1
            counters["column_passes"] += len(ins["cols"])
            _fused_launch(plan, circ, ins, slots, apply_fn_for(plan["runs"][ins["run"]]), A, D, v)
        else:
            raise ValueError(k)
    Cm = A + A.conj().T + D
    Q = Cm - np.outer(v, v.conj())
    psi = np.concatenate([slots[r][prog["psi"]] for r in range(world)]) if prog["psi_final"] else None
    return Q, psi, counters


def run_gradient_program(plan: dict, circ, psi: np.ndarray):
    """Interprets an adjoint-gradient program (qgt_b200_plan_dump_gradient) on the CPU: slot 0 = psi, slot 2 = H psi.
    Returns dE/dtheta.  Fused launches contribute A[0, nu] = <Lambda|G_nu|chi>, Gram instructions <Lambda|column>."""
    prog = plan["program"]
    P = plan["P"]
    dim = psi.size
    slots = [[np.zeros(dim, dtype=np.complex128) for _ in range(prog["slots"])]]
    idx = np.arange(dim, dtype=np.uint64)
    slots[0][0] = psi.copy()
    slots[0][2] = _cost_energy(circ, idx.astype(np.int64)) * psi
    A = np.zeros((max(P, 1), max(P, 1)), dtype=np.complex128)
    Dm = np.zeros_like(A)
    v = np.zeros(max(P, 1), dtype=np.complex128)
    grad = np.zeros(P)
    f = lambda r, st, op: apply_op(st, op, circ)
    for ins in prog["instrs"]:
        k = ins["k"]
        if k == "sweep":
            run = plan["runs"][ins["run"]]
            ops, dops = run["ops"], run["dops"]
            results = []
            for (src, dst, ovr, acc, extra) in ins["cols"]:
                total = 0
                for o in ([ovr] + list(extra) if ovr >= 0 else [None]):
                    st = slots[0][src]
                    for i, op in enumerate(ops):
                        st = apply_op(st, dops[str(i)] if i == o else op, circ)
                    total = total + st
                results.append((dst, acc, total))
            dsts = [d for d, _, _ in results]
            assert len(set(dsts)) == len(dsts)
            for dst, acc, total in results:
                slots[0][dst] = slots[0][dst] + total if acc else total
        elif k == "fused":
            _fused_launch(plan, circ, ins, slots, f, A, Dm, v)
        elif k == "gram":
            assert ins["aid"] == [P] and len(ins["a"]) == 1
            for b, bid in zip(ins["b"], ins["bid"]):
                grad[bid] += -2.0 * np.vdot(slots[0][ins["a"][0]], slots[0][b]).real
        else:
            raise ValueError(k)
    if prog["fused"] == 1:
        grad += -2.0 * A[0, :P].real
    return grad


def run_gradient_program_sharded(plan: dict, circ, psi: np.ndarray, world: int):
    """The same for a state sharded over `world` ranks (all simulated in this process): psi arrives in the identity qubit
    layout (what the forward plan restores), Lambda = H psi is formed shard by shard, exchanges move both states, the
    partial results of the ranks are summed (the allreduce)."""
    prog = plan["program"]
    P, nloc = plan["P"], plan["nloc"]
    dim = 1 << nloc
    tabs = segment_cost_tables(plan, circ)
    energy = _cost_energy(circ, np.arange(psi.size, dtype=np.int64))
    slots = [[np.zeros(dim, dtype=np.complex128) for _ in range(prog["slots"])] for _ in range(world)]
    for r in range(world):
        slots[r][0] = psi[r * dim:(r + 1) * dim].copy()
        slots[r][2] = (energy * psi)[r * dim:(r + 1) * dim].copy()
    A = np.zeros((max(P, 1), max(P, 1)), dtype=np.complex128)
    Dm = np.zeros_like(A)
    v = np.zeros(max(P, 1), dtype=np.complex128)
    grad = np.zeros(P)

    def apply_fn_for(run):
        edges, vw = tabs[run["segment"]]
        def f(r, st, op):
            gidx = (np.uint64(r) << np.uint64(nloc)) | np.arange(st.size, dtype=np.uint64)
            return _apply_op_shard(st, op, circ, gidx, nloc, edges, vw)
        return f

    for ins in prog["instrs"]:
        k = ins["k"]
        run = plan["runs"][ins["run"]] if "run" in ins else None
        if k == "sweep":
            if run["exchange"] >= 0:
                for col in ins["cols"]:
                    src, dst, ovr, acc = col[0], col[1], col[2], col[3]
                    assert src == dst and ovr < 0 and not acc
                    exchange_all_ranks([slots[r][dst] for r in range(world)], run["exchange_mask"])
                continue
            f = apply_fn_for(run)
            ops, dops = run["ops"], run["dops"]
            for r in range(world):
                results = []
                for col in ins["cols"]:
                    src, dst, ovr, acc, extra = col[0], col[1], col[2], col[3], col[4]
                    total = 0
                    for o in ([ovr] + list(extra) if ovr >= 0 else [None]):
                        st = slots[r][src]
                        for i, op in enumerate(ops):
                            st = f(r, st, dops[str(i)] if i == o else op)
                        total = total + st
                    results.append((dst, acc, total))
                for dst, acc, total in results:
                    slots[r][dst] = slots[r][dst] + total if acc else total
        elif k == "fused":
            _fused_launch(plan, circ, ins, slots, apply_fn_for(run), A, Dm, v)
        elif k == "gram":
            assert ins["aid"] == [P] and len(ins["a"]) == 1
            for b, bid in zip(ins["b"], ins["bid"]):
                grad[bid] += -2.0 * sum(np.vdot(slots[r][ins["a"][0]], slots[r][b]) for r in range(world)).real
        else:
            raise ValueError(k)
    if prog["fused"] == 1:
        grad += -2.0 * A[0, :P].real
    return grad
