"""Host-side domain decomposition and halo plan, checked on CPU: single-process consistency
of the plans of all ranks, and a world_size-2 `gloo` run that performs the halo exchange the plan
describes with torch.distributed send/recv and checks every ghost against its owner's value."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tria(dim, reps):
    import openifem_b200 as ifem

    t = ifem.Triangulation(dim)
    lo, hi = (0,) * dim, (1,) * dim
    ifem.GridGenerator.subdivided_hyper_rectangle(t, reps, lo, hi, True)
    return t


@pytest.mark.parametrize("dim,reps,size", [(2, (6, 8), 2), (2, (5, 9), 3), (3, (4, 4, 8), 4), (3, (3, 3, 5), 2)])
def test_partition_is_a_partition_and_plans_match(dim, reps, size):
    import openifem_b200 as ifem

    t = _tria(dim, reps)
    parts = [ifem.Partition(t, 2, 1, r, size) for r in range(size)]
    n_nodes = {0: int(np.prod([2 * k + 1 for k in reps])), 1: int(np.prod([k + 1 for k in reps]))}
    for which in (0, 1):
        owned = []
        for r, p in enumerate(parts):
            c = p.counts(which)
            l2g = p.local_to_global(which)
            assert len(np.unique(l2g)) == len(l2g)
            o = l2g[: c["n_owned"]]
            assert np.all(np.diff(o) > 0)  # owned nodes ascending in global id
            owned.append(o)
        allo = np.concatenate(owned)
        assert len(allo) == n_nodes[which] and len(np.unique(allo)) == n_nodes[which]  # disjoint cover
        # the k-th message r sends to s is exactly what s expects as its k-th message from r (one message per
        # ghost layer), in the same order
        for r, p in enumerate(parts):
            l2g_r = p.local_to_global(which)
            for s, q in enumerate(parts):
                if s == r:
                    continue
                out = [x for x in p.neighbours(which) if x["rank"] == s]
                back = [x for x in q.neighbours(which) if x["rank"] == r]
                assert len(out) == len(back) <= 2
                l2g_s = q.local_to_global(which)
                for a, b in zip(out, back):
                    sent = l2g_r[a["send_local"]]
                    expect = l2g_s[b["recv_offset"]: b["recv_offset"] + b["recv_count"]]
                    assert np.array_equal(sent, expect)
                    assert np.all(a["send_local"] < p.counts(which)["n_owned"])
            c = p.counts(which)
            assert c["n_owned"] <= c["n_layer1"] <= c["n_local"]
            # every ghost is received exactly once
            got = np.zeros(c["n_local"], dtype=int)
            for nb in p.neighbours(which):
                got[nb["recv_offset"]: nb["recv_offset"] + nb["recv_count"]] += 1
            assert np.all(got[c["n_owned"]:] == 1) and np.all(got[: c["n_owned"]] == 0)
    # every cell is local somewhere; the ghost layer is one cell deep for slabs
    assert sum(p.counts(0)["n_local_cells"] for p in parts) >= int(np.prod(reps))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _halo_worker(rank, size, port, dim, reps, q):
    try:
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist

        import openifem_b200 as ifem

        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=size)
        t = ifem.Triangulation(dim)
        ifem.GridGenerator.subdivided_hyper_rectangle(t, reps, (0,) * dim, (1,) * dim, True)
        part = ifem.Partition(t, 2, 1, rank, size)
        ok = True
        for which, bs in ((0, dim), (1, 1)):
            c = part.counts(which)
            l2g = part.local_to_global(which).astype(np.int64)
            f = lambda g, k: np.sin(0.37 * g + 1.3 * k)  # value of component k at global node g
            v = np.full((c["n_local"], bs), np.nan)
            for k in range(bs):
                v[: c["n_owned"], k] = f(l2g[: c["n_owned"]], k)
            vt = torch.from_numpy(v)
            reqs, bufs = [], []
            for nb in part.neighbours(which):
                if len(nb["send_local"]):
                    sb = vt[torch.from_numpy(nb["send_local"].astype(np.int64))].contiguous()
                    bufs.append(sb)
                    reqs.append(dist.isend(sb, nb["rank"]))
                if nb["recv_count"]:
                    rb = vt[nb["recv_offset"]: nb["recv_offset"] + nb["recv_count"]]
                    reqs.append(dist.irecv(rb, nb["rank"]))
            for r in reqs:
                r.wait()
            for k in range(bs):
                ok = ok and np.allclose(v[:, k], f(l2g, k), atol=0, rtol=0)
            # a dot product over owned entries summed over ranks equals the global dot product
            loc = torch.tensor([float((v[: c["n_owned"]] ** 2).sum())], dtype=torch.float64)
            dist.all_reduce(loc)
            n_glob = int(np.prod([(2 if which == 0 else 1) * r_ + 1 for r_ in reps]))
            ref = sum(float((f(np.arange(n_glob), k) ** 2).sum()) for k in range(bs))
            ok = ok and abs(loc.item() - ref) < 1e-9 * ref
        dist.destroy_process_group()
        q.put((rank, bool(ok), ""))
    except Exception as e:  # pragma: no cover
        q.put((rank, False, repr(e)))


@pytest.mark.parametrize("dim,reps", [(2, (4, 6)), (3, (3, 3, 4))])
def test_halo_exchange_world2_gloo(dim, reps):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_halo_worker, args=(r, 2, port, dim, reps, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, msg in res:
        assert ok, f"rank {rank}: {msg}"


@pytest.mark.parametrize("dim,reps,axis,size", [(3, (3, 3, 12), 2, 2), (3, (3, 3, 12), 2, 4), (2, (12, 4), 0, 3), (2, (4, 12), 1, 2)])
def test_slabs_of_a_band_refined_mesh_keep_hanging_nodes_with_their_masters(dim, reps, axis, size):
    """locally refined meshes (csrc/partition.cpp plane_slabs): every rank's owned nodes tile the mesh, and a hanging node has
    the owner of all its masters - the condition under which the condensation of the hanging rows is local to a rank
    (csrc/hanging.cu); the slab axis is found automatically (last axis with enough admissible planes)"""
    import openifem_b200 as ifem

    t = _tria(dim, reps)
    v, c, _ = t.get_mesh()
    x = v[c].mean(axis=1)[:, axis]
    t.execute_refinement(((x > 0.34) & (x < 0.67)).astype(np.uint8))
    v, c, _ = t.get_mesh()
    hv, hk, hm = t.hanging()
    assert hv.size > 0
    parts = [ifem.Partition(t, 1, 1, r, size) for r in range(size)]
    owner = np.full(v.shape[0], -1)
    for r, p in enumerate(parts):
        n_owned = p.counts(0)["n_owned"]
        l2g = p.local_to_global(0)
        assert n_owned > 0
        assert np.all(owner[l2g[:n_owned]] == -1)
        owner[l2g[:n_owned]] = r
    assert np.all(owner >= 0)
    # the FE_Q(1) node of a vertex: nodes are numbered in lexicographic (z, y, x) order of their quantised position (csrc/mesh.cpp
    # spatial_renumber)
    lo, hi = v.min(axis=0), v.max(axis=0)
    q = np.rint((v - lo) / np.where(hi > lo, hi - lo, 1.0) * float((1 << 21) - 1)).astype(np.int64)
    order = np.lexsort(tuple(q[:, d] for d in range(dim)))
    vert_node = np.empty(v.shape[0], dtype=int)
    vert_node[order] = np.arange(v.shape[0])
    for h, k, m in zip(hv, hk, hm):
        owners = {owner[vert_node[h]]} | {owner[vert_node[j]] for j in m[:k]}
        assert len(owners) == 1, (h, m[:k], owners)
