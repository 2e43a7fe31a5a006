"""Host logic of the contraction-phase plan (maf_config.h::build_config): the built-in tuned plans and MAF_PLAN
overrides are valid plans (every chunk exactly once), invalid texts are refused. Runs tools/print_plan.cpp on the CPU."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOTIONS = {"STATIC": 1, "EUL": 2, "LAG": 3, "ALEV": 4, "ALEVB": 5}


@pytest.fixture(scope="module")
def print_plan(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("plan") / "print_plan")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "membranealefem.jl_b200", "csrc"),
                           os.path.join(ROOT, "tools", "print_plan.cpp"), "-o", exe])
    return exe


def _run(exe, motion, plan=None):
    env = dict(os.environ)
    env.pop("MAF_PLAN", None)
    if plan is not None:
        env["MAF_PLAN"] = plan
    return subprocess.run([exe, str(MOTIONS[motion])], env=env, capture_output=True, text=True)


def _assignment(out):
    """chunk id -> (warp, round) from the printer's lines"""
    got = {}
    for m in re.finditer(r"chunk\s+(\d+) .*-> warp (-?\d+) round (-?\d+)", out):
        got[int(m.group(1))] = (int(m.group(2)), int(m.group(3)))
    return got


@pytest.mark.parametrize("motion", sorted(MOTIONS))
def test_built_in_plan_places_every_chunk_once(print_plan, motion):
    r = _run(print_plan, motion)
    assert r.returncode == 0, r.stderr
    n = int(re.search(r"nchunks (\d+)", r.stdout).group(1))
    got = _assignment(r.stdout)
    assert sorted(got) == list(range(n))
    assert all(0 <= w < 4 and rd >= 0 for w, rd in got.values())
    assert len(set(got.values())) == n          # no two chunks in the same (warp, round) slot


def test_override_is_taken_literally_and_invalid_plans_are_refused(print_plan):
    r = _run(print_plan, "LAG")
    n = int(re.search(r"nchunks (\d+)", r.stdout).group(1))
    ids = list(range(n))
    plan = ",".join(map(str, ids[::-1][: n // 2])) + "/" + ",".join(map(str, ids[::-1][n // 2:]))
    r = _run(print_plan, "LAG", plan)
    assert r.returncode == 0, r.stderr
    got = _assignment(r.stdout)
    for w, part in enumerate(plan.split("/")):
        for rd, c in enumerate(int(x) for x in part.split(",")):
            assert got[c] == (w, rd)
    for bad in ("0,0,1", "0,1", ",".join(map(str, ids)) + f",{n}", "a,b", "0/1/2/3/4"):
        assert _run(print_plan, "LAG", bad).returncode != 0, bad
