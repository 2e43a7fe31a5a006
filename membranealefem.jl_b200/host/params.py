"""`Params` and `check_params` of the reference (src/input/Params.jl:36-125)."""
import dataclasses
import os

from .enums import (ALEVB, F_BEND, F_CAVI, F_COUE, F_POIS, F_PULL, NDERS, NEN, POLY, STATIC, ZDIM, Motion, Scenario)


@dataclasses.dataclass(frozen=True)
class Params:
    """Params.jl:36-55 (same field names; `zv`, `adb`, `am`, `ek`, `enr` stand for ζv, αdb, αm, εk, εnr)."""
    motion: Motion = ALEVB
    scenario: Scenario = F_PULL
    num1el: int = 17
    num2el: int = 17
    output: bool = True
    length: float = 64.0
    kb: float = 1.0
    kg: float = -0.5
    zv: float = 1.0
    pn: float = 0.0
    poly: int = 2
    gp1d: int = 3
    nders: int = 2
    nen: int = None
    adb: float = None
    am: float = 1.0
    ek: float = 1.0e-15
    enr: float = 1.0e-12

    def __post_init__(self):
        if self.nen is None:
            object.__setattr__(self, "nen", (self.poly + 1) ** ZDIM)
        if self.adb is None:
            object.__setattr__(self, "adb", self.length ** 2)
        object.__setattr__(self, "motion", Motion(self.motion))
        object.__setattr__(self, "scenario", Scenario(self.scenario))


def check_params(p: Params, **args):
    """Params.jl:70-125. AssertionError on the same conditions, with the same messages."""
    ks = args.keys()
    assert "Δts" in ks or "dts" in ks, "\nneed list of time steps 'Δts'"
    assert "t0" in ks, "\nneed initial time 't0'"
    assert "t0_id" in ks, "\nneed initial time ID 't0_id'"
    if p.output:
        assert "out_file" in ks, "\nneed output file 'out_file'"
        assert "out_path" in ks, "\nneed output path 'out_path'"
        os.makedirs(args["out_path"], exist_ok=True)
    assert p.nen == (p.poly + 1) ** ZDIM, "\nincorrect number of element nodes"
    assert p.nen == NEN
    assert p.poly == POLY
    assert p.nders == NDERS
    if p.scenario == F_BEND:
        assert "bend_tm" in ks, "\nneed final ramp-up time 'bend_tm'"
        assert "bend_mf" in ks, "\nneed final applied bending moment 'bend_mf'"
        assert args["bend_mf"] == p.kb / 2 / p.length, "\nα = 1/2 is the half-angle"
        assert p.motion != STATIC, f"\nmesh cannot be static for {F_BEND.name} scenario"
        assert p.motion != ALEVB, f"\n{F_BEND.name} not implemented for {ALEVB.name} motion"
        assert p.pn == 0.0, f"{F_BEND.name} with normal pressure not implemented"
    elif p.scenario == F_PULL:
        assert "pull_speed" in ks, "\nneed pull speed 'pull_speed'"
    elif p.scenario == F_CAVI:
        assert p.motion == STATIC, "\nneed static motion for cavity scenario"
    elif p.scenario == F_COUE:
        assert p.motion == STATIC, "\nneed static motion for Couette scenario"
    elif p.scenario == F_POIS:
        assert p.motion == STATIC, "\nneed static motion for Poiseuille scenario"
    assert p.scenario in (F_CAVI, F_COUE, F_POIS, F_PULL, F_BEND), f"\n{p.scenario} scenario not implemented!"


def get_dts(args):
    return list(args["Δts"] if "Δts" in args else args["dts"])
