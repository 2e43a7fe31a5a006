"""`solve` and `restart` of the reference (src/MembraneAleFem.jl:44-107)."""
import os
import pickle

from .analysis import run_analysis
from .input import prepare_input
from .params import check_params


def _display_params(p, **args):
    """display_params (Params.jl:134-156): text dump + serialised params/args for `restart`."""
    if not p.output:
        return
    with open(os.path.join(args["out_path"], "params.txt"), "a") as f:
        f.write(f"{p}\n\n{args}\n\n")
    with open(os.path.join(args["out_path"], "params.dat"), "wb") as f:
        pickle.dump(p, f)
    with open(os.path.join(args["out_path"], "args.dat"), "wb") as f:
        pickle.dump(args, f)


def solve(p, **args):
    """MembraneAleFem.solve(p; args...) (MembraneAleFem.jl:44-60). Returns (mesh, xms, cps, newton_histories)."""
    check_params(p, **args)
    _display_params(p, **args)
    mesh, xms, cps = prepare_input(p, **args)
    hist = run_analysis(mesh, xms, cps, p, **args)
    return mesh, xms, cps, hist


def restart(p_file, a_file, **r_args):
    """MembraneAleFem.restart(p_file, a_file; r_args...) (MembraneAleFem.jl:75-107): reload the serialised
    parameters and arguments, overlay the new keyword arguments and re-enter `solve`.

    params.dat / args.dat are pickles (the reference uses Julia's Serialization the same way): loading them executes
    whatever they contain, so only files this package wrote itself (`_display_params`) may be passed in."""
    with open(p_file, "rb") as f:
        p = pickle.load(f)
    with open(a_file, "rb") as f:
        args = pickle.load(f)
    for k in ("in_path", "Δts", "dts", "t0", "t0_id"):
        assert k in r_args or k in ("Δts", "dts"), f"restart needs '{k}'"
    args = {**args, **r_args}
    args.setdefault("in_xms", f"t{args['t0_id']}-xms.txt")
    args.setdefault("in_cps", f"t{args['t0_id']}-cps.txt")
    return solve(p, **args)
