"""N > 1 host-side logic on CPU: two gloo ranks shard the pair rows the way the sharded GPU context
does (mesh.shard_range), exchange their slices with an all-gather, and reproduce the unsharded result;
the canonical kinetic-energy reduction (per-warp partials, all-gathered, summed in a fixed pattern) is
bit-identical for any rank count. The arithmetic here is the oracle's -- this checks decomposition and
collective plumbing, not the CUDA kernels (those are covered on the GPU box by tools/check_sharded.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def canonical_ke(warp_partials: np.ndarray, threads: int = 128) -> float:
    """Host restatement of canonical_sum (csrc/mesh_kernels.cuh): thread t adds entries t, t+128, ...;
    warps reduce with a shuffle-down tree; warp 0 adds the per-warp results in warp order."""
    acc = np.zeros(threads)
    for b in range(len(warp_partials)):
        acc[b % threads] += warp_partials[b]
    lanes = acc.reshape(threads // 32, 32).copy()
    o = 16
    while o > 0:
        lanes[:, :o] = lanes[:, :o] + lanes[:, o:2 * o]
        o >>= 1
    total = 0.0
    for w in range(threads // 32):
        total += lanes[w, 0]
    return float(total)


def warp_partials(vel: np.ndarray, lo: int, hi: int) -> np.ndarray:
    """Per-warp kinetic energy of vertices [lo,hi) (lo multiple of 32), shuffle-down tree order."""
    n_w = (hi - lo + 31) // 32
    out = np.zeros(n_w)
    for w in range(n_w):
        lanes = np.zeros(32)
        a, b = lo + 32 * w, min(lo + 32 * w + 32, hi)
        v = vel[a:b]
        lanes[:b - a] = 0.5 * (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2])
        o = 16
        while o > 0:
            lanes[:o] = lanes[:o] + lanes[o:2 * o]
            o >>= 1
        out[w] = lanes[0]
    return out


def worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import sys
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from np_shape_lab_b200 import mesh as M, workloads as W
    from oracle import oracle as O
    w = W.build("S8_disc")
    n = w.n_vertices
    rng = np.random.default_rng(3)
    pos = w.mesh.pos + 0.01 * rng.standard_normal(w.mesh.pos.shape)
    vel = 0.05 * rng.standard_normal(pos.shape)
    lo, hi = M.shard_range(n, rank, world)
    span = M.shard_range(n, 0, world)[1]
    o = O.Oracle(w.mesh.edges, w.mesh.faces, w.params["l0"], w.params)
    o.set_state(pos, vel, w.q)
    o.dressup()
    o.force(lo, hi - 1)                          # owned rows x all columns
    mine = torch.zeros(span, 3, dtype=torch.float64)
    mine[:hi - lo] = torch.from_numpy(o.vec("pforce")[lo:hi])
    parts = [torch.zeros(span, 3, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, mine)                  # the reference gathers forces, src/newforces.cpp:268-271
    gathered = torch.cat(parts)[:n].numpy()
    # kinetic energy: warp partials of the owned range, all-gathered, canonical sum
    wp = torch.zeros(span // 32, dtype=torch.float64)
    mine_wp = warp_partials(vel, lo, hi)
    wp[:len(mine_wp)] = torch.from_numpy(mine_wp)
    allwp = [torch.zeros(span // 32, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allwp, wp)
    ke = canonical_ke(torch.cat(allwp).numpy())
    # the NCCL unique id travels as an opaque object from rank 0 (bench.py does the same over NCCL)
    box = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    if rank == 0:
        o.force()
        full = o.vec("pforce")
        q.put((np.array_equal(gathered, full), ke, canonical_ke(warp_partials(vel, 0, n)),
               box[0] == bytes(range(128)), (lo, hi)))
    else:
        q.put((box[0] == bytes(range(128)), (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_sharding_reproduces_unsharded_result():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0 = [o for o in outs if len(o) == 5][0]
    r1 = [o for o in outs if len(o) == 2][0]
    same_force, ke_sharded, ke_single, id_ok, (lo0, hi0) = r0
    assert same_force and id_ok and r1[0]
    assert ke_sharded == ke_single                      # bit-identical, not just close
    assert lo0 == 0 and hi0 == r1[1][0] and hi0 % 32 == 0 and r1[1][1] == 642


@pytest.mark.parametrize("n,world", [(372, 2), (972, 4), (40962, 8), (163842, 8), (40962, 1)])
def test_shard_ranges_tile_the_rows(n, world):
    from np_shape_lab_b200 import mesh as M
    cur = 0
    for r in range(world):
        lo, hi = M.shard_range(n, r, world)
        assert lo == cur and lo % 32 == 0 and hi >= lo
        cur = hi
    assert cur == n
    # never more than one warp of rows away from the reference's ceil(N/P) slices
    ref = M.row_partition(n, 0, world)
    assert 0 <= M.shard_range(n, 0, world)[1] - ref[1] < 32


# ---- the symmetric pair kernel's work decomposition (host-side planner, no device needed) -----------------------

def _pair_count(units, n, rb):
    """Number of pairs (i, j), j > i, that the units cover, and a coverage check on a coarse grid."""
    total = 0
    for I, c, j0, j1 in units:
        i0, i1 = I * rb, min((I + 1) * rb, n)
        for i in (range(i0, i1) if j0 < i1 else ()):  # rows that overlap the column range (diagonal units)
            total += max(0, j1 - max(j0, i + 1))
        if j0 >= i1:
            total += (i1 - i0) * (j1 - j0)
    return total


@pytest.mark.parametrize("n", [4096, 10242, 40962])
def test_symmetric_units_cover_every_pair_exactly_once_on_any_rank_count(n):
    """Units are a function of the vertex count only; the ranks of a 1/2/4/8-GPU job share them out by virtual rank.
    Together they cover each unordered pair once, and the shares are balanced."""
    from np_shape_lab_b200 import capi
    all_units, geo1 = capi.plan_pair_units(n, 0, 1)
    rb = geo1["rows_per_block"]
    assert rb == 6 * 64 and geo1["column_granularity"] == 32  # default shape of k_pair_sym: 6 rows x 64 threads
    assert _pair_count(all_units.tolist(), n, rb) == n * (n - 1) // 2
    keys1 = sorted((int(u[0]), int(u[1])) for u in all_units)
    assert len(set(keys1)) == len(keys1)
    work = np.array([(u[3] - u[2]) for u in all_units])
    assert np.all(np.diff(-((work + 127) // 128)) >= 0)  # launch order: heaviest (most column tiles) first
    for world in (2, 4, 8):
        per_rank, keys = [], []
        for r in range(world):
            u, geo = capi.plan_pair_units(n, r, world)
            assert geo == geo1  # the decomposition does not depend on the rank count
            keys += [(int(a[0]), int(a[1])) for a in u]
            per_rank.append(_pair_count(u.tolist(), n, rb))
        assert sorted(keys) == keys1
        assert sum(per_rank) == n * (n - 1) // 2
        if n >= 40962:
            assert max(per_rank) <= 1.01 * min(per_rank), per_rank  # balanced shares
    with pytest.raises(capi.NpslError):
        capi.plan_pair_units(n, 0, 3)
