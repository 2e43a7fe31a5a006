#!/usr/bin/env python
"""bench.py -- residual + tangent assembly (calc_r_K) throughput on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (one calc_r_K: area elements + Neumann boundary elements + scatter) over the
synthetic flat F_PULL patch of BASELINE.json configs[4] (1001 x 1001 = 1 002 001 elements, ALEVB, perturbed state of
SURVEY.md 8(d)5). `value` = elements/s with the inputs resident in HBM; `e2e` = the same call through the C ABI's
host-buffer entry point (maf_assemble: H2D of xms/cps, D2H of r and nzval inside the timed region).
N > 1 (torchrun): the patch is cut into N strips of element rows (strong scaling), one strip handle per rank
(maf_create_strip: sliced tables and result buffers); after its kernels the upper strip of every pair adds the lower
strip's interface sums straight from its memory over NVLink (CUDA IPC peer loads, flag-ordered, inside the library)
and the residual norm is all-reduced over NCCL.
--impl reference times the CPU restatement of the reference algorithm (the reference itself is Julia, which this
image does not have) on a bounded sample of the same workload with all host threads.

Process hygiene: the timed repo arm maps libmembrane_b200.so only; everything that needs the oracle -- the
cpu_baseline leg, the CPU side of `newton_iteration`, the `parity_spot` check -- runs in child processes of this
script (`--impl reference`, `--parity-spot`), and the reference arm never loads the product library.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "FP64 residual+tangent assembly throughput (calc_r_K)"
UNIT = "Melem/s"
# algorithmic work per element, SURVEY.md 8(d): bytes B_el = 8*25*ndf^2 + 8*ndf + 8*(3+ndf) + 36, FLOPs F_el
B_EL = {"LAG": 3324, "EUL": 9972, "ALEV": 12988, "ALEVB": 12988}
F_EL = {"LAG": 0.20e6, "EUL": 0.24e6, "ALEV": 0.41e6, "ALEVB": 0.41e6}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--motion", default="ALEVB", choices=["LAG", "EUL", "ALEV", "ALEVB"])
    ap.add_argument("--n", "--patch-n", type=int, default=int(os.environ.get("MAF_BENCH_N", "1001")),
                    help="elements per direction of the synthetic patch (under torchrun use --patch-n or MAF_BENCH_N: "
                         "torchrun's own parser rejects the abbreviation-like --n)")
    ap.add_argument("--scatter", default="atomic", choices=["atomic", "deterministic"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-newton", action="store_true", help="skip the newton_iteration object (configs 1-4)")
    ap.add_argument("--no-spot", action="store_true", help="skip the parity_spot child process")
    ap.add_argument("--parity-spot", action="store_true", help="child mode: single-element parity check on the bench "
                    "patch against the oracle and the extended-precision truth; prints one JSON object")
    ap.add_argument("--newton-cpu", action="store_true", help="with --impl reference: also time the CPU port on the "
                    "17x17 meshes of configs 1-4 (for the newton_iteration object)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def mark(self):
        """Number of samples received so far (to cut the window of the timed region out of a longer recording)."""
        return len(self.rows)

    def stop(self, first=0):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows[first:]:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------- CPU arm
def lm_pattern(LM, nmdf):
    """0-based CSC pattern = union over elements of (active LM rows x active LM cols)."""
    keys = []
    for e in range(LM.shape[1]):
        act = LM[:, e][LM[:, e] != 0].astype(np.int64) - 1
        keys.append((act[None, :] * nmdf + act[:, None]).ravel())       # col * nmdf + row
    keys = np.unique(np.concatenate(keys))
    cols, rows = keys // nmdf, keys % nmdf
    colptr = np.zeros(nmdf + 1, dtype=np.int64)
    np.add.at(colptr, cols + 1, 1)
    return np.cumsum(colptr), rows


CPU_FLAGS = "g++ -O3 -march=native -ffp-contract=off -fcx-limited-range (oracle/Makefile: libmaf_oracle_native.so, " \
            "compiled on this machine)"


def cpu_reference_rate(motion, steps, warmup, target_s=1.5):
    """Times the oracle's calc_r_K (complex-step tangent, the reference's chunked threading scheme, private
    accumulators, sum, serial Neumann loop -- FiniteElement.jl:75-200) on a bounded sample of the workload.
    Inputs come from the pure-Python host mirror (mafb200 is imported, libmembrane_b200.so is NOT loaded)."""
    import mafb200 as maf
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    mcode = getattr(orc, motion)

    def setup(n):
        om = orc.Mesh(motion=mcode, scenario=orc.F_PULL, num1el=n, num2el=n, pull_speed=0.5, kind="native")
        p = maf.Params(motion=getattr(maf, motion), scenario=maf.F_PULL, num1el=n, num2el=n, output=False)
        hm = maf.Mesh(p, pull_speed=0.5)
        xms, cps = maf.synthetic_state(hm, p)
        colptr, rows = lm_pattern(om.LM, om.nmdf)
        return om, xms, cps, colptr, rows

    om, xms, cps, colptr, rows = setup(17)
    t0 = time.perf_counter()
    om.calc_r_K_fast(xms, cps, 0.5, 0.5, colptr, rows, nthreads=cores, want_out=False)
    rate = om.numel / (time.perf_counter() - t0)
    n = int(min(160, max(17, np.sqrt(rate * target_s))))
    if n >= 18:
        n = max(n, 19)
    om, xms, cps, colptr, rows = setup(n)
    for _ in range(max(1, min(warmup, 2))):
        om.calc_r_K_fast(xms, cps, 0.5, 0.5, colptr, rows, nthreads=cores, want_out=False)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        om.calc_r_K_fast(xms, cps, 0.5, 0.5, colptr, rows, nthreads=cores, want_out=False)
        ts.append(time.perf_counter() - t0)
    tot = sum(ts)
    return {"value": om.numel * steps / tot / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n}x{n}-element F_PULL {motion} patch ({om.numel} elements/step, {steps} steps, same "
                      f"perturbed state generator), C++ restatement of the reference algorithm (complex-step tangent, "
                      f"chunked std::thread tasks with private accumulators as FiniteElement.jl:88-147), "
                      f"{CPU_FLAGS}, {cores} threads; Julia itself is not installed"}, tot / steps * 1e3


NEWTON_CONFIGS = [("1 lag-pull", "LAG"), ("2 eul-pull", "EUL"), ("3 ale-pull", "ALEVB"), ("4 translate-ale", "ALEVB")]


def newton_cpu_times():
    """CPU port on the 17 x 17 meshes of BASELINE.json configs 1-4: ms per calc_r_K (all host threads, reference
    threading scheme) on the deformed state after the predictor of step 1 -- the CPU side of `newton_iteration`."""
    import mafb200 as maf
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    out = {}
    for name, motion in NEWTON_CONFIGS:
        p = maf.Params(motion=getattr(maf, motion), scenario=maf.F_PULL, num1el=17, num2el=17, output=False)
        mesh, xms, cps = maf.prepare_input(p, pull_speed=0.5, dts=[0.5], t0=0.0, t0_id=0)
        maf.update_xms(p.motion, xms, cps, 0.5, mesh.dofs)
        om = orc.Mesh(motion=getattr(orc, motion), scenario=orc.F_PULL, num1el=17, num2el=17, pull_speed=0.5,
                      kind="native")
        ts = []
        for _ in range(6):
            t0 = time.perf_counter()
            om.calc_r_K(xms, cps, 0.5, 0.5, nthreads=cores)
            ts.append(time.perf_counter() - t0)
        out[name] = {"cpu_port_assemble_ms": float(np.median(ts[1:]) * 1e3), "threads": cores}
    return out


def newton_gpu_times(maf, device):
    """GPU side of `newton_iteration` (BASELINE.json metric, second half; loop of FiniteElement.jl:29-55): the four
    small configs on the reference's default 17 x 17 mesh, two time steps each (pull_speed 0.5, dt 0.5). Per Newton
    iteration: wall time of the assembly call through the C ABI with host buffers (maf_assemble: H2D state, kernels,
    D2H r + nzval), of the device-resident variant (maf_assemble_resident: no state upload), the device time of the
    kernels alone, and the host solve, timed separately (SciPy SuperLU stands in for Julia's UMFPACK `\\`): from
    scratch in every iteration like the reference (:38), and through host/solver.py (value buffer owned by the solver,
    column ordering chosen once per pattern)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    out = {}
    for name, motion in NEWTON_CONFIGS:
        p = maf.Params(motion=getattr(maf, motion), scenario=maf.F_PULL, num1el=17, num2el=17, output=False)
        args = dict(pull_speed=0.5, dts=[0.5, 0.5], t0=0.0, t0_id=0, device=device)
        mesh, xms, cps = maf.prepare_input(p, **args)
        if name.startswith("4"):   # translate-ale emulation: converged tether state, then an in-plane
            maf.run_analysis(mesh, xms, cps, p, **args)          # Dirichlet velocity on the pulled nodes
            U = maf.Dof.Unknown
            for (unk, node, val) in mesh.inh_dir_bcs:
                cps[node - 1, mesh.dofs[U.vx] - 1] = 0.2
                cps[node - 1, mesh.dofs[U.vmx] - 1] = 0.2
        res = {}
        maf.calc_r_K(mesh, xms, cps, 0.5, 0.5, p, device=device)     # warm-up: handle creation, first launches
        for label, kw in (("host_buffers+scratch_lu", {}), ("resident+pattern_solver", {"resident": True, "solver": "pattern"})):
            x, c = xms.copy(), cps.copy()
            timers = {}
            t0 = time.perf_counter()
            hist = maf.run_analysis(mesh, x, c, p, timers=timers, **kw, **args)
            wall = time.perf_counter() - t0
            it = timers["iterations"]
            res[label] = {"iterations": it, "assemble_call_ms": timers["assembly_s"] / it * 1e3,
                          "host_solve_ms": timers["solve_s"] / it * 1e3, "newton_iteration_ms": wall / it * 1e3,
                          "eps_first_step": hist[0]}
            if kw.get("solver") == "pattern":     # the one-time analysis of the pattern is not a per-iteration cost
                ps = maf.pkg.host.analysis._pattern_solver(mesh, p, args)
                once = ps.timers["order_s"]
                steady = (timers["solve_s"] - once) / max(it - 1, 1) * 1e3
                res[label].update({"host_solve_ms": steady,
                                   "newton_iteration_ms": res[label]["assemble_call_ms"] + steady,
                                   "pattern_analysis_once_ms": once * 1e3, "ordering": ps.ordering,
                                   "fill_nnz_L_plus_U": int(ps.fill)})
        asm = maf.pkg.host.analysis._assembler(mesh, p, args)
        # kernels alone, device-resident state: CUDA events around the launches of one assembly (maf_timings)
        asm.state_set(xms, cps)
        dev_ms, call_ms = [], []
        r_buf, k_buf = np.empty(mesh.nmdf), np.empty(asm.nnz)
        for _ in range(12):
            t0 = time.perf_counter()
            asm.assemble_resident(0.5, 0.5, r=r_buf, nzval=k_buf)
            call_ms.append((time.perf_counter() - t0) * 1e3)
            tm = asm.timings()
            dev_ms.append(tm["zero_ms"] + tm["area_ms"] + tm["bdry_ms"])
        res["kernels_device_ms"] = float(np.median(dev_ms[2:]))
        res["kernels_device_split_ms"] = {k: tm[k] for k in ("zero_ms", "area_ms", "bdry_ms")}
        res["graph_replays"] = asm.kernel_info()["graph_replays"]
        res["assemble_resident_call_ms"] = float(np.median(call_ms[2:]))
        res["nmdf"], res["nnz"], res["numel"] = int(mesh.nmdf), int(asm.nnz), int(mesh.numel)
        maf.pkg.host.analysis.close_assemblers(mesh)
        out[name] = res
    return out


def run_child(extra, timeout=900):
    """Runs this script in a child process (keeps the oracle out of the timed process) and returns its JSON line."""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR",
                                                              "MASTER_PORT", "LOCAL_WORLD_SIZE", "GROUP_RANK")}
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__)] + extra, capture_output=True, text=True,
                           timeout=timeout, env=env, cwd=ROOT)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        return json.loads(lines[-1]) if lines else {"error": (r.stderr or "no output")[-400:]}
    except Exception as e:   # the bench line must still be printed
        return {"error": repr(e)[:400]}


def parity_spot(a):
    """Child mode (--parity-spot): a few dozen single elements of the bench patch through the library against the
    oracle's element routine and the extended-precision truth (tests/spot_parity.py; the GPU test suite runs > 300)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import mafb200 as maf
    from oracle import oracle as orc
    from spot_parity import check_elements, select_elements
    p = maf.Params(motion=getattr(maf, a.motion), scenario=maf.F_PULL, num1el=a.n, num2el=a.n, output=False)
    mesh = maf.Mesh(p, pull_speed=0.5)
    xms, cps = maf.synthetic_state(mesh, p)
    kw = dict(motion=getattr(orc, a.motion), scenario=orc.F_PULL, num1el=a.n, num2el=a.n, length=p.length,
              pull_speed=0.5)
    om, ot = orc.Mesh(**kw), orc.Mesh(kind="truth", **kw)
    asm = maf.Assembler(mesh, p, device=0)
    dx = torch.from_numpy(np.ascontiguousarray(xms.T)).cuda()
    dc = torch.from_numpy(np.ascontiguousarray(cps.T)).cuda()
    els = select_elements(mesh, n_random=12)
    els = els[::max(1, len(els) // 40)]
    w = check_elements(asm, mesh, om, ot, dx.data_ptr(), dc.data_ptr(), xms, cps, 0.5, 0.5, els)
    return {"elements_checked": w["n"], "what": "single elements via maf_set_element_range(el, el) vs the oracle's "
            "elem_r_K scattered through LM and vs the extended-precision truth (tests/spot_parity.py)",
            "dof_numbering_equal": bool(np.array_equal(om.ID, mesh.ID)),
            "max_rel_diff_K_vs_oracle": w["K_rel_oracle"], "max_abs_diff_r_vs_oracle": w["r_abs_oracle"],
            "strict_rule_worst_ratio": max(w["K_strict"], w["r_strict"]),
            "gpu_error_in_eps_E": w["K_in_epsE"], "oracle_error_in_eps_E": w["oracle_in_epsE"],
            "rule": "|x - truth| <= 1e-11 |truth| + eps E per entry (tests/cases.py), ratio <= 1 passes"}


# ----------------------------------------------------------------------------------------------- main
def main():
    a = parse()
    # exactly ONE line may reach stdout (the JSON): libraries that print there (e.g. the NCCL version banner) are
    # sent to stderr for the whole run, the JSON line is written to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rc = run(a, real_stdout)
    real_stdout.flush()
    return rc


def run(a, out_stream):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.parity_spot:
        out_stream.write(json.dumps(parity_spot(a)) + "\n")
        return 0

    if a.impl == "reference":
        if rank != 0:
            return 0
        cb, ms = cpu_reference_rate(a.motion, a.steps, a.warmup)
        out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus,
               "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": f"synthetic flat F_PULL patch, {a.motion}, bounded CPU sample (see cpu_baseline)"},
               "cpu_baseline": cb,
               "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0}
        if a.newton_cpu:
            out["newton_cpu"] = newton_cpu_times()
        out_stream.write(json.dumps(out) + "\n")
        return 0

    import torch
    import torch.distributed as dist
    import mafb200 as maf
    if rank == 0 and not os.path.exists(maf.pkg.capi.LIB_PATH):   # normally prebuilt (__graft_entry__.build)
        subprocess.check_call([sys.executable, "-c", "import __graft_entry__ as g; g.build_cuda()"], cwd=ROOT)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    motion = getattr(maf, a.motion)
    p = maf.Params(motion=motion, scenario=maf.F_PULL, num1el=a.n, num2el=a.n, output=False)
    t_setup = time.perf_counter()
    mesh = maf.Mesh(p, pull_speed=0.5)
    xms, cps = maf.synthetic_state(mesh, p)
    free0 = torch.cuda.mem_get_info(dev)[0]
    asm = maf.Assembler(mesh, p, device=local_rank, strip=(rank, world) if world > 1 else None)
    t_setup = time.perf_counter() - t_setup
    mode = maf.SCATTER_ATOMIC if a.scatter == "atomic" else maf.SCATTER_DETERMINISTIC
    dt = 0.5

    # strips of element rows: one strip handle per rank, neighbours attached through CUDA IPC exports
    sinfo = None
    if world > 1:
        sinfo = asm.strip_info()
        mine = torch.tensor(list(asm.peer_export()), dtype=torch.uint8, device=dev)
        allh = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.all_gather(allh, mine)
        asm.peer_attach(bytes(allh[rank - 1].cpu().tolist()) if rank > 0 else None,
                        bytes(allh[rank + 1].cpu().tolist()) if rank + 1 < world else None)
    info = asm.range_info()
    my_elems = info["elements"][1] - info["elements"][0] + 1

    d_x = torch.from_numpy(np.ascontiguousarray(xms.T)).to(dev)
    d_c = torch.from_numpy(np.ascontiguousarray(cps.T)).to(dev)
    d_r = d_k = None
    if world == 1:
        d_r = torch.zeros(mesh.nmdf, dtype=torch.float64, device=dev)
        d_k = torch.zeros(asm.nnz, dtype=torch.float64, device=dev)
    d_n = torch.zeros(1, dtype=torch.float64, device=dev)
    d_n2 = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(2)]
    nstep, pending = [0], [None, None]
    torch.cuda.synchronize()
    mem_used = free0 - torch.cuda.mem_get_info(dev)[0]
    # everything (kernels, interface copies, NCCL ops, timing events) is ordered on the handle's own stream
    torch.cuda.synchronize()
    ext = torch.cuda.ExternalStream(asm.stream(), device=dev)
    torch.cuda.set_stream(ext)

    def step():
        if world == 1:
            asm.assemble_device(d_x.data_ptr(), d_c.data_ptr(), dt, dt, scatter_mode=mode, d_r=d_r.data_ptr(),
                                d_nzval=d_k.data_ptr(), d_rnorm2=d_n.data_ptr(), stream=None)
        else:
            # kernels of the strip, interface sums pulled from the lower neighbour's memory over NVLink, partial |r|^2
            # The one NCCL collective of a step, the residual norm, runs on NCCL's stream beside the next step's
            # zero-fill and kernels (two norm buffers, used alternately); the last one is awaited inside the timed
            # region (drain()).
            slot = nstep[0] % 2
            nstep[0] += 1
            if pending[slot] is not None:
                pending[slot].wait()
            asm.assemble_strip(d_x.data_ptr(), d_c.data_ptr(), dt, dt, scatter_mode=mode,
                               d_rnorm2_partial=d_n2[slot].data_ptr())
            pending[slot] = dist.all_reduce(d_n2[slot], async_op=True)

    def drain():
        for k in range(2):
            if pending[k] is not None:
                pending[k].wait()
                pending[k] = None
        if world > 1 and nstep[0] > 0:
            d_n.copy_(d_n2[(nstep[0] - 1) % 2])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # the sampler needs ~0.1 s to deliver its first line: it is started before the warm-up and only the samples
    # that arrive from the start of the timed region on are used
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(a.warmup):
        step()
    drain()
    barrier()
    s_first = sampler.mark()
    l0 = asm.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(a.steps):          # no host synchronisation inside the timed region
        step()
    drain()
    ev1.record()
    barrier()
    # device time of the dominant kernel in every launch of the timed region (event ring inside the library), and the
    # other parts of the LAST step
    n_ring = min(a.steps, 64)
    area_ms = asm.area_kernel_times(n_ring)
    last = asm.timings()
    xch_ms = [asm.strip_timings()["exchange_ms"]] if world > 1 else []
    ms = ev0.elapsed_time(ev1)
    launches = asm.launch_count() - l0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    rnorm2 = float(d_n.item())
    # a timed region shorter than the sampling period (many GPUs, few steps) may have seen no sample: keep the same
    # load running, untimed, until two samples have arrived (every rank runs the same number of extra steps)
    extra = 0
    if world > 1:
        need = torch.tensor([1 if (rank == 0 and sampler.mark() - s_first < 2) else 0], dtype=torch.int64, device=dev)
        dist.broadcast(need, 0)
        short = bool(need.item())
    else:
        short = sampler.mark() - s_first < 2
    if short:
        for _ in range(int(max(1.0, 400.0 / max(ms / a.steps, 1e-3)))):   # ~0.4 s of the same steps
            step()
            extra += 1
        drain()
        barrier()
    clocks = sampler.stop(s_first) if rank == 0 else None
    if clocks is not None and extra:
        clocks["note"] = f"timed region shorter than the sampling period: {extra} more untimed steps of the same load were sampled"
    value = mesh.numel * a.steps / (ms * 1e-3) / 1e6

    # ---- the deterministic scatter path on the same workload (N = 1): bitwise reproducible, ascending element id ----
    det = None
    if world == 1 and a.scatter == "atomic" and not a.no_e2e:
        dmode = maf.SCATTER_DETERMINISTIC
        try:
            for _ in range(2):
                asm.assemble_device(d_x.data_ptr(), d_c.data_ptr(), dt, dt, scatter_mode=dmode, d_r=d_r.data_ptr(),
                                    d_nzval=d_k.data_ptr(), d_rnorm2=d_n.data_ptr(), stream=None)
            torch.cuda.synchronize()
            nd = max(3, min(a.steps, 5))
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            d0.record()
            for _ in range(nd):
                asm.assemble_device(d_x.data_ptr(), d_c.data_ptr(), dt, dt, scatter_mode=dmode, d_r=d_r.data_ptr(),
                                    d_nzval=d_k.data_ptr(), d_rnorm2=d_n.data_ptr(), stream=None)
            d1.record()
            torch.cuda.synchronize()
            dms = d0.elapsed_time(d1) / nd
            ki = asm.kernel_info()
            det = {"value": mesh.numel / (dms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": dms,
                   "fraction_of_atomics_rate": (mesh.numel / (dms * 1e-3) / 1e6) / value,
                   "staging_bytes": ki["staging_bytes"], "staging_fraction_of_nzval": ki["staging_bytes"] / (asm.nnz * 8.0),
                   "band_rows": ki["band_rows"], "rnorm2": float(d_n.item())}
        except maf.MafError as e:
            det = {"error": str(e)[:200]}
        # leave the atomics result in the buffers for the checks below
        asm.assemble_device(d_x.data_ptr(), d_c.data_ptr(), dt, dt, scatter_mode=mode, d_r=d_r.data_ptr(),
                            d_nzval=d_k.data_ptr(), d_rnorm2=d_n.data_ptr(), stream=None)
        torch.cuda.synchronize()

    # ---- e2e through the host-buffer entry point (pinned host memory) ------------------------------------------
    e2e = None
    if not a.no_e2e and world == 1:
        hx = torch.from_numpy(np.ascontiguousarray(xms.T)).pin_memory()
        hc = torch.from_numpy(np.ascontiguousarray(cps.T)).pin_memory()
        hr = torch.empty(mesh.nmdf, dtype=torch.float64).pin_memory()
        hk = torch.empty(asm.nnz, dtype=torch.float64).pin_memory()
        xs, cs = hx.numpy().T, hc.numpy().T          # column-major views for the ctypes call
        for _ in range(max(1, a.warmup)):
            asm.assemble(xs, cs, dt, dt, scatter_mode=mode, r=hr.numpy(), nzval=hk.numpy())
        tot = 0.0
        for _ in range(a.steps):
            asm.assemble(xs, cs, dt, dt, scatter_mode=mode, r=hr.numpy(), nzval=hk.numpy())
            tot += asm.timings()["total_ms"]
        tm = asm.timings()
        e2e = {"value": mesh.numel * a.steps / (tot * 1e-3) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int((3 + mesh.ndf) * mesh.numnp * 8),
               "d2h_bytes_per_step": int((mesh.nmdf + asm.nnz + 1) * 8),
               "ms_per_step": tot / a.steps, "last_step_ms": tm,
               "note": "maf_assemble with pinned host buffers: H2D xms+cps, kernels, D2H r + nzval (the K values "
                       "the host solver needs) inside the timed region (CUDA events on the handle's stream)"}
        err = float(np.abs(hr.numpy() - d_r.cpu().numpy()).max())
        e2e["max_abs_diff_r_vs_device_path"] = err
        # host-link roofline of this box for this traffic: one plain pinned device-to-host copy of nzval, nothing else
        # running (what the e2e call cannot beat: its copies hide the kernels, not the other way round)
        link_s = float("inf")
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            asm.download(1, 0, 1, asm.nnz, nz_out=hk.numpy())
            torch.cuda.synchronize()
            if rep > 0:
                link_s = min(link_s, time.perf_counter() - t0)
        e2e["host_link_roofline"] = {"d2h_gbs": asm.nnz * 8 / link_s / 1e9, "ms_for_nzval": link_s * 1e3,
                                     "e2e_fraction_of_link_roofline": link_s / (tot / a.steps * 1e-3),
                                     "how": "one pinned device-to-host copy of nzval, best of 2 after a warm-up"}
    elif not a.no_e2e:
        # N > 1, through the strip handle's host-buffer entry point (maf_assemble_strip_host): every rank uploads the
        # node rows its strip reads from pinned host memory, assembles, takes part in the interface exchange and copies
        # the rows of r / entries of nzval it OWNS (every entry leaves exactly one GPU) into pinned host memory
        own_r = sinfo["own_rows"][1] - sinfo["own_rows"][0] + 1
        own_k = sinfo["own_slots"][1] - sinfo["own_slots"][0] + 1
        hx = torch.from_numpy(np.ascontiguousarray(xms.T)).pin_memory()
        hc = torch.from_numpy(np.ascontiguousarray(cps.T)).pin_memory()
        hr = torch.empty(own_r, dtype=torch.float64).pin_memory()
        hk = torch.empty(own_k, dtype=torch.float64).pin_memory()
        xs, cs = hx.numpy().T, hc.numpy().T
        rn_host = 0.0
        for _ in range(max(1, a.warmup)):
            asm.assemble_strip_host(xs, cs, dt, dt, scatter_mode=mode, r_own=hr.numpy(), nzval_own=hk.numpy())
        barrier()
        tot = 0.0
        for _ in range(a.steps):
            _, _, rn_part = asm.assemble_strip_host(xs, cs, dt, dt, scatter_mode=mode, r_own=hr.numpy(),
                                                    nzval_own=hk.numpy())
            tot += asm.timings()["total_ms"]        # CUDA events on the handle's stream around the whole call
            rn_host = rn_part
        barrier()
        t = torch.tensor([tot], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot = float(t.item())
        rn_t = torch.tensor([rn_host], dtype=torch.float64, device=dev)
        dist.all_reduce(rn_t)
        nodes = info["nodes"][1] - info["nodes"][0] + 1
        nb = torch.tensor([own_r + own_k + 1, (3 + mesh.ndf) * nodes], dtype=torch.int64, device=dev)
        dist.all_reduce(nb)
        # host-link roofline of this box for this traffic: every rank copies its owned nzval slice to pinned host
        # memory at the same time, nothing else running (what the e2e step cannot beat)
        # (one untimed repetition first; then the best of three, each the max over ranks of a copy started together)
        link_s = float("inf")
        for rep in range(4):
            barrier()
            t0 = time.perf_counter()
            asm.download(1, 0, sinfo["own_slots"][0], own_k, nz_out=hk.numpy())
            torch.cuda.synchronize()
            tl = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            dist.all_reduce(tl, op=dist.ReduceOp.MAX)
            if rep > 0:
                link_s = min(link_s, float(tl.item()))
        kb = torch.tensor([own_k * 8], dtype=torch.int64, device=dev)
        dist.all_reduce(kb)
        link_gbs = float(kb.item()) / link_s / 1e9
        e2e_s = tot / a.steps * 1e-3
        e2e = {"value": mesh.numel * a.steps / (tot * 1e-3) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(nb[1].item()) * 8, "d2h_bytes_per_step": int(nb[0].item()) * 8,
               "ms_per_step": tot / a.steps,
               "note": "per rank, maf_assemble_strip_host: H2D of the node rows the strip reads (pinned host memory), "
                       "strip assembly, interface sums over NVLink, D2H of the owned rows of r / entries of nzval "
                       "into pinned host memory (every entry leaves exactly one GPU); CUDA events on the handle's "
                       "stream around the call, max over ranks",
               "rnorm2_host": float(rn_t.item()),
               "host_link_roofline": {"aggregate_d2h_gbs": link_gbs, "ms_for_nzval": link_s * 1e3,
                                      "e2e_fraction_of_link_roofline": link_s / e2e_s,
                                      "how": f"{world} ranks copy their owned nzval slices to pinned host memory "
                                             "concurrently, nothing else running; max over ranks, best of 3 "
                                             "(a reference point measured with one plain copy per rank: the strip "
                                             "pipeline of the e2e call can come out slightly ahead of it)"}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (area_kernel) ---------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    area = float(np.mean(area_ms))
    other = {k: float(last[k]) for k in ("zero_ms", "bdry_ms", "gather_ms")}
    other["note"] = ("last step; atomics path: the Neumann boundary kernels run beside the area kernel on a second "
                     "stream (bdry_ms = what is left to wait for)")
    fp64_peak = maf.fp64_peak_tflops(local_rank)
    ach_gbs = my_elems * B_EL[a.motion] / (area * 1e-3) / 1e9
    ach_tf = my_elems * F_EL[a.motion] / (area * 1e-3) / 1e12
    # DRAM bytes of one launch of this kernel on this workload, from the committed `ncu --set full` capture
    traffic, traffic_src = None, None
    tfile = os.path.join(ROOT, "profiles", "r2_area_kernel_alevb_1001_traffic.json")
    if a.motion == "ALEVB" and a.n == 1001 and a.scatter == "atomic" and world == 1 and os.path.exists(tfile):
        tj = json.load(open(tfile))
        traffic, traffic_src = tj["dram_read"] + tj["dram_write"], "profiles/r2_area_kernel_alevb_1001_full.txt"
    roofline = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": f"area_kernel<{a.motion}>", "kernel_ms": area,
                "peak_source": "MEASURED_PEAKS.json (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_element": B_EL[a.motion],
                "fp64": {"achieved_tflops": ach_tf, "peak_tflops": fp64_peak, "frac": ach_tf / fp64_peak,
                         "algorithmic_flop_per_element": F_EL[a.motion],
                         "peak_source": "maf_fp64_peak DFMA microbenchmark, this run",
                         "note": "the kernel runs on the FP64 CUDA cores and is on the compute side of the roofline "
                                 "(arithmetic intensity ~32 FLOP/B vs machine balance ~5): the HBM fraction above is "
                                 "the schema's headline, this one is the relevant ceiling. What limits it below the "
                                 "DFMA peak (ncu, DESIGN.md section 4): the shared-memory data pipe (69 %) and the "
                                 "latency chain of the Gauss-point items at 3 resident CTAs per SM"},
                "other_kernels_ms": other}

    # ---- everything that needs the oracle runs in child processes (this process maps libmembrane_b200.so only) ---
    cpu, newton, spot = None, None, None
    if not a.no_cpu and world == 1:
        child = run_child(["--impl", "reference", "--steps", "4", "--warmup", "1", "--motion", a.motion] +
                          ([] if a.no_newton else ["--newton-cpu"]))
        cpu = child.get("cpu_baseline", child)
        if not a.no_newton:
            newton = newton_gpu_times(maf, local_rank)
            for name, v in (child.get("newton_cpu") or {}).items():
                newton.setdefault(name, {}).update(v)
            newton["note"] = ("per Newton iteration on the reference's default 17x17 F_PULL mesh (configs 1-4; 4 = "
                              "translate-ale emulation), 2 time steps: assembly call through the C ABI, host solve "
                              "timed separately (SciPy SuperLU), CPU port of the reference assembly beside it")
    if not a.no_spot and world == 1 and a.scatter == "atomic":
        spot = run_child(["--parity-spot", "--motion", a.motion, "--patch-n", str(a.n)])

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"synthetic flat F_PULL patch {a.n}x{a.n} = {mesh.numel} elements, {a.motion} "
                                  f"(ndf {mesh.ndf}), perturbed state (SURVEY 8(d)5)",
                      "knots": "reference rule (Mesh.jl:176-181): centre-refined knots for >= 18 elements/direction",
                      "numel": mesh.numel, "numnp": mesh.numnp, "nmdf": mesh.nmdf, "nnz": asm.nnz,
                      "pattern": "P_blk", "scatter": a.scatter, "elements_per_rank": my_elems,
                      "l2": "no flush: each step streams r + nzval (%.1f GB) >> 126 MB L2" % (asm.nnz * 8 / 1e9),
                      "parallelism": (f"{world} strips of element rows, one strip handle per rank (maf_create_strip); "
                                      f"interface sums read from the lower neighbour's memory over NVLink (CUDA IPC "
                                      f"peer loads inside the library, {float(np.mean(xch_ms)):.3f} ms per step on "
                                      f"rank 0 incl. waiting for the neighbour); NCCL all-reduce of the residual norm")
                      if world > 1 else "single GPU",
                      "device_mem_used_gb_rank0": mem_used / 1e9,
                      "setup_s": t_setup, "kernel": asm.kernel_info()},
           "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
           "newton_iteration": newton, "parity_spot": spot, "value_deterministic": det, "rnorm2": rnorm2}
    out_stream.write(json.dumps(out) + "\n")
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
