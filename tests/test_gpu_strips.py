"""Strips of element rows behind the C ABI (include/maf.h: maf_create_strip, maf_peer_attach_local, maf_peer_export /
maf_peer_attach, maf_assemble_strip, maf_assemble_strip_host): the reference's per-task chunks and their sum
(FiniteElement.jl:88-89, 144-147) spread over several GPUs. One handle per strip; the interface rows / entries are
added by the upper strip of each pair from the lower strip's memory (peer loads over NVLink).

In one process, all strips on the devices that exist (one GPU: every strip on device 0 -- the flag protocol and the
slice arithmetic are the same; two or more GPUs: one strip per GPU with peer access). A second test runs one PROCESS
per strip with CUDA IPC exports exchanged over torch.distributed (gloo)."""
import os
import socket
import sys

import numpy as np
import pytest

import mafb200 as maf
from cases import check_strict, make_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ndev():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("name,nranks", [("alevb_pull_17x17", 2), ("alevb_pull_17x17", 3), ("lag_pull_17x17", 4),
                                         ("eul_pull_5x4", 2), ("alevb_pull_fine_19x18", 3)])
@pytest.mark.parametrize("mode", [maf.SCATTER_ATOMIC, maf.SCATTER_DETERMINISTIC])
def test_strips_in_one_process_equal_the_whole(name, nranks, mode):
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    whole = maf.Assembler(hm, p)
    r_w, nz_w, rn_w = whole.assemble(xms, cps, time, dt, scatter_mode=mode)
    colptr, rowval = whole.pattern()
    ndev = _ndev()
    strips = [maf.Assembler(hm, p, device=k % ndev, strip=(k, nranks)) for k in range(nranks)]
    infos = [s.strip_info() for s in strips]
    # the owned ranges tile 1..nmdf and 1..nnz; memory per strip is a fraction of the whole
    assert infos[0]["own_rows"][0] == 1 and infos[-1]["own_rows"][1] == hm.nmdf
    assert infos[0]["own_slots"][0] == 1 and infos[-1]["own_slots"][1] == whole.nnz
    for a, b in zip(infos[:-1], infos[1:]):
        assert a["own_rows"][1] + 1 == b["own_rows"][0] and a["own_slots"][1] + 1 == b["own_slots"][0]
        assert a["rows"][1] >= b["rows"][0] and a["slots"][1] >= b["slots"][0]          # the interface overlaps
    for k, s in enumerate(strips):
        s.peer_attach_local(strips[k - 1] if k > 0 else None, strips[k + 1] if k + 1 < nranks else None)
    for rep in range(3):          # several steps: the flags carry the step number
        outs = [None] * nranks
        # device-pointer entry point: every strip is launched asynchronously, then collected
        for s in strips:
            s.state_set(xms, cps)
        for s in strips:
            s.assemble_strip(None, None, time, dt, scatter_mode=mode)
        for k, s in enumerate(strips):
            i = s.strip_info()
            s.sync()
            outs[k] = s.download(i["own_rows"][0], i["own_rows"][1] - i["own_rows"][0] + 1,
                                 i["own_slots"][0], i["own_slots"][1] - i["own_slots"][0] + 1)
        r = np.concatenate([o[0] for o in outs])
        nz = np.concatenate([o[1] for o in outs])
        assert np.abs(nz - nz_w).max() <= 1e-13 * np.abs(nz_w).max()
        assert np.abs(r - r_w).max() <= 1e-13 * max(np.abs(r_w).max(), 1e-300) + 1e-14
    check_strict(name, r, nz, colptr, rowval, xms, cps, time, dt, f"{nranks} strips, mode {mode}")
    if mode == maf.SCATTER_DETERMINISTIC and nranks == 2:
        # host-buffer entry point (blocking: the lower strip first would wait for nobody, the upper one pulls)
        import threading
        res = [None] * nranks
        th = [threading.Thread(target=lambda k=k: res.__setitem__(k, strips[k].assemble_strip_host(
            xms, cps, time, dt, scatter_mode=mode))) for k in range(nranks)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        r2 = np.concatenate([o[0] for o in res])
        nz2 = np.concatenate([o[1] for o in res])
        assert np.array_equal(r2, r) and np.array_equal(nz2, nz)        # deterministic path: bitwise
        assert abs(sum(o[2] for o in res) - rn_w) <= 1e-12 * rn_w
    for s in strips:
        assert s.strip_timings()["result_bytes"] < 0.8 * 8 * (whole.nnz + hm.nmdf) or nranks == 1
        s.close()
    whole.close()


def test_strip_handles_refuse_what_they_cannot_do():
    p, hm, om, xms, cps, time, dt, args = make_case("alevb_pull_17x17")
    with pytest.raises(maf.MafError, match="too few element rows"):
        maf.Assembler(hm, p, strip=(0, 9))
    s = maf.Assembler(hm, p, strip=(0, 2))
    with pytest.raises(maf.MafError, match="strip"):
        s.assemble(xms, cps, time, dt)
    with pytest.raises(maf.MafError, match="fixed element range"):
        s.set_element_range(1, 10)
    with pytest.raises(maf.MafError, match="attach"):
        s.state_set(xms, cps)
        s.assemble_strip(None, None, time, dt)
    s.close()


def _free_port():
    sk = socket.socket()
    sk.bind(("127.0.0.1", 0))
    port = sk.getsockname()[1]
    sk.close()
    return port


def _ipc_worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import mafb200 as m
        from cases import make_case as mk
        p, hm, om, xms, cps, time, dt, args = mk("alevb_pull_17x17")
        dev = rank % torch.cuda.device_count()
        s = m.Assembler(hm, p, device=dev, strip=(rank, world))
        mine = torch.tensor(list(s.peer_export()), dtype=torch.uint8)
        allh = [torch.zeros(64, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(allh, mine)
        s.peer_attach(bytes(allh[rank - 1].tolist()) if rank > 0 else None,
                      bytes(allh[rank + 1].tolist()) if rank + 1 < world else None)
        outs = []
        for step in range(3):
            r_own, nz_own, rn = s.assemble_strip_host(xms, cps, time, dt, scatter_mode=m.SCATTER_DETERMINISTIC)
            t = torch.tensor([rn], dtype=torch.float64)
            dist.all_reduce(t)
            outs.append((r_own.copy(), nz_own.copy(), float(t.item())))
        dist.barrier()
        q.put((rank, outs[-1][0], outs[-1][1], outs[-1][2],
               all(np.array_equal(o[1], outs[0][1]) for o in outs)))
        s.close()
        dist.destroy_process_group()
    except Exception:
        import traceback
        q.put((rank, "error", traceback.format_exc(), 0.0, False))


@pytest.mark.parametrize("world", [2, 3])
def test_one_process_per_strip_over_cuda_ipc(world):
    import torch.multiprocessing as mp
    p, hm, om, xms, cps, time, dt, args = make_case("alevb_pull_17x17")
    whole = maf.Assembler(hm, p)
    r_w, nz_w, rn_w = whole.assemble(xms, cps, time, dt, scatter_mode=maf.SCATTER_DETERMINISTIC)
    whole.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ipc_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted((q.get(timeout=600) for _ in range(world)), key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
    for t in res:
        assert not isinstance(t[1], str), t[2]
    r = np.concatenate([t[1] for t in res])
    nz = np.concatenate([t[2] for t in res])
    # every strip sums its own elements in ascending element id and the interface adds (lower strip) + (upper strip):
    # the same numbers as the whole mesh up to the association of that last addition; bitwise equal from step to step
    assert np.abs(nz - nz_w).max() <= 1e-13 * np.abs(nz_w).max() and np.abs(r - r_w).max() <= 1e-13 * np.abs(r_w).max()
    for t in res:
        assert abs(t[3] - rn_w) <= 1e-12 * rn_w and t[4]
