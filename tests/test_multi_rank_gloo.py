"""N > 1 path on the CPU: world_size 2 and 3 with the gloo backend. Each rank assembles its strip of element rows
(the kernels' phase functions, emulated on the host), exchanges the interface rows/entries with its neighbours
through membranealefem.jl_b200/host/partition.py -- the same code bench.py runs over NCCL -- and must end up with
the single-rank result on every row/entry it touches; the all-reduced residual norm must equal the global one."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, motion_code, n1, n2, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import mafb200 as maf
        from emu_driver import Emu
        part = maf.pkg.host.partition
        p = maf.Params(motion=maf.Motion(motion_code), scenario=maf.F_PULL, num1el=n1, num2el=n2, length=8.0,
                       output=False)
        mesh = maf.Mesh(p, pull_speed=0.5)
        xms, cps = maf.synthetic_state(mesh, p)           # counter-based: identical on every rank
        emu = Emu(mesh, p)
        colptr, rowval = emu.pattern()
        e_first, e_last = part.strip_elements(n1, n2, world, rank)
        r, K = emu.assemble(xms, cps, 0.5, 0.5, mode=1, el_first=e_first, el_last=e_last)
        rng = part.touched_ranges(mesh, colptr, e_first, e_last)
        # only the touched ranges may be non-zero (that is all a strip writes)
        nz = K.data
        assert np.all(r[:rng[0] - 1] == 0) and np.all(r[rng[1]:] == 0)
        assert np.all(nz[:rng[2] - 1] == 0) and np.all(nz[rng[3]:] == 0)
        mine = torch.tensor(list(rng), dtype=torch.int64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        ranges = [t.tolist() for t in allr]
        t_r, t_k, t_n = torch.from_numpy(r.copy()), torch.from_numpy(nz.copy()), torch.zeros(1, dtype=torch.float64)
        ex = part.InterfaceExchange(dist, ranges, rank, t_r)
        ex(t_r, t_k, t_n)
        # reference: the whole mesh on one rank
        r_all, K_all = emu.assemble(xms, cps, 0.5, 0.5, mode=1)
        rows = slice(rng[0] - 1, rng[1])
        slots = slice(rng[2] - 1, rng[3])
        er = np.abs(t_r.numpy()[rows] - r_all[rows]).max() / np.abs(r_all).max()
        ek = np.abs(t_k.numpy()[slots] - K_all.data[slots]).max() / np.abs(K_all.data).max()
        en = abs(float(t_n) - float(r_all @ r_all)) / float(r_all @ r_all)
        q.put((rank, er, ek, en, ex.bytes_per_step()))
        dist.destroy_process_group()
    except Exception as e:   # surface the failure in the parent
        import traceback
        q.put((rank, "error", traceback.format_exc(), str(e), 0))


@pytest.mark.parametrize("world,motion_code,n1,n2", [(2, 5, 6, 8), (3, 3, 5, 9), (2, 2, 4, 6)])
def test_strips_plus_interface_exchange_equal_single_rank(world, motion_code, n1, n2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, motion_code, n1, n2, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    for (rank, er, ek, en, nbytes) in res:
        assert er != "error", ek
        assert er < 1e-13 and ek < 1e-13 and en < 1e-13, (rank, er, ek, en)
        assert nbytes > 0


def test_partition_arithmetic():
    sys.path.insert(0, ROOT)
    import mafb200 as maf
    part = maf.pkg.host.partition
    # strips tile the element range exactly
    for n2, world in [(1001, 8), (17, 2), (9, 3), (1001, 1)]:
        nxt = 1
        for r in range(world):
            a, b = part.strip_elements(7, n2, world, r)
            assert a == nxt and b >= a
            nxt = b + 1
        assert nxt == 7 * n2 + 1
    # strips thinner than two element rows would couple rank k with rank k + 2: refused, not mis-summed
    for n2, world in [(4, 4), (3, 2), (5, 3), (2, 3)]:
        with pytest.raises(ValueError, match="at least 2 rows"):
            part.strip_rows(n2, world, 0)
    assert part.strip_rows(4, 2, 1) == (2, 4) and part.strip_rows(1, 1, 0) == (0, 1)
    ranges = [(1, 100, 1, 1000), (81, 200, 801, 2000), (181, 300, 1801, 3000)]
    ov = part.overlaps(ranges, 1)
    assert [(o.peer, o.rows, o.slots) for o in ov] == [(0, slice(80, 100), slice(800, 1000)),
                                                       (2, slice(180, 200), slice(1800, 2000))]
    assert part.owned_rows(ranges, 0) == slice(0, 100) and part.owned_rows(ranges, 1) == slice(100, 200)
    assert part.owned_rows(ranges, 2) == slice(200, 300)
    assert part.owned_slots(ranges, 0) == slice(0, 1000) and part.owned_slots(ranges, 1) == slice(1000, 2000)
    assert part.owned_slots(ranges, 2) == slice(2000, 3000)


def test_library_strips_match_the_host_partition():
    """The strip arithmetic behind the C ABI (maf_create_strip -> maf_host.h::strip_range, here through tests/emu) cuts
    the same strips and touches the same ranges as host/partition.py; thin strips are refused by both."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import mafb200 as maf
    from emu_driver import Emu
    part = maf.pkg.host.partition
    p = maf.Params(motion=maf.ALEVB, scenario=maf.F_PULL, num1el=5, num2el=9, length=8.0, output=False)
    mesh = maf.Mesh(p, pull_speed=0.5)
    emu = Emu(mesh, p)
    colptr, _ = emu.pattern()
    for world in (1, 2, 3, 4):
        prev_rows = prev_slots = None
        for rank in range(world):
            els, rows, slots = emu.strip_range(rank, world)
            assert els == part.strip_elements(5, 9, world, rank)
            assert (rows[0], rows[1], slots[0], slots[1]) == part.touched_ranges(mesh, colptr, *els)
            if prev_rows is not None:       # neighbouring strips overlap (the two shared node rows), nothing else
                assert prev_rows[0] < rows[0] <= prev_rows[1] < rows[1] and prev_slots[0] < slots[0] <= prev_slots[1] < slots[1]
            prev_rows, prev_slots = rows, slots
    with pytest.raises(RuntimeError, match="too few element rows"):
        emu.strip_range(0, 5)
