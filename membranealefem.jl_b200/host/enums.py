"""Enumerations of the reference (src/input/Enums.jl:24-156, src/input/Dof.jl:28-51), same names and codes."""
import enum


class Scenario(enum.IntEnum):   # Enums.jl:24-30
    F_CAVI = 1
    F_COUE = 2
    F_POIS = 3
    F_PULL = 4
    F_BEND = 5


class Topology(enum.IntEnum):   # Enums.jl:43-46
    FLAT = 1
    CYLINDER = 2


class Motion(enum.IntEnum):     # Enums.jl:67-73
    STATIC = 1
    EUL = 2
    LAG = 3
    ALEV = 4
    ALEVB = 5


class Boundary(enum.IntEnum):   # Enums.jl:89-94
    BOTTOM = 1
    RIGHT = 2
    TOP = 3
    LEFT = 4


class Corner(enum.IntEnum):     # Enums.jl:109-114
    BOTTOM_LEFT = 1
    BOTTOM_RIGHT = 2
    TOP_LEFT = 3
    TOP_RIGHT = 4


class Neumann(enum.IntEnum):    # Enums.jl:131-135
    SHEAR = 1
    STRETCH = 2
    MOMENT = 3


class Curve(enum.IntEnum):      # Enums.jl:151-154
    CLAMPED = 1
    CLOSED = 2


class Dof:
    """module Dof (src/input/Dof.jl)."""

    class Unknown(enum.IntEnum):    # Dof.jl:28-33
        vx = 1
        vy = 2
        vz = 3
        vmx = 4
        vmy = 5
        vmz = 6
        lam = 7    # Dof.λ
        pm = 8

    class Position(enum.IntEnum):   # Dof.jl:49-51
        xm = 1
        ym = 2
        zm = 3


# compile-time constants of the reference (src/input/Params.jl:161-179)
POLY, GP1D, NDERS, ZDIM, XDIM, NEN, VOIGT, NEDBDF = 2, 3, 2, 2, 3, 9, 3, 3

STATIC, EUL, LAG, ALEV, ALEVB = Motion.STATIC, Motion.EUL, Motion.LAG, Motion.ALEV, Motion.ALEVB
F_CAVI, F_COUE, F_POIS, F_PULL, F_BEND = (Scenario.F_CAVI, Scenario.F_COUE, Scenario.F_POIS, Scenario.F_PULL,
                                          Scenario.F_BEND)
BOTTOM, RIGHT, TOP, LEFT = Boundary.BOTTOM, Boundary.RIGHT, Boundary.TOP, Boundary.LEFT
SHEAR, STRETCH, MOMENT = Neumann.SHEAR, Neumann.STRETCH, Neumann.MOMENT
CLAMPED, CLOSED = Curve.CLAMPED, Curve.CLOSED
FLAT, CYLINDER = Topology.FLAT, Topology.CYLINDER
