"""CPU model of the producer / consumer protocol of ``conv3d_wgrad_f32_ws_*`` (predict_pv_yield_b200/csrc/conv3d_wgrad_f32.cu).

The kernel's correctness rests on one host-checkable invariant: with ``kWsD`` step packages in flight and a ring of
``kWsR`` input-plane slots addressed by a running position (a run of consecutive output times starts with three new
planes, every further step adds one), the producer never overwrites a slot that a consumer may still read -- for ANY
sequence of run lengths, including runs of a single step at the boundaries of a CTA's range.  The model below replays
the kernel's bookkeeping (``pos`` / ``next`` / ``n % kWsD``) with the constants parsed from the source and lets the
producer run as far ahead as the ``empty`` barriers allow (the worst case for slot reuse).  It also checks the
work-splitting arithmetic both roles repeat independently (they must agree on every step without communicating).
"""
import os
import random
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "predict_pv_yield_b200", "csrc", "conv3d_wgrad_f32.cu")).read()


def _const(name):
    m = re.search(rf"constexpr int {name} = ([^;]+);", SRC)
    assert m, name
    expr = m.group(1).strip()
    return int(eval(expr, {"kWsD": _const("kWsD")} if "kWsD" in expr and name != "kWsD" else {}))


D = _const("kWsD")
R = _const("kWsR")


def runs_of_range(g_begin, g_end, To):
    """(t0, nstep) of the runs both roles derive from a CTA's step range (conv3d_wgrad_f32_ws body)."""
    g = g_begin
    out = []
    while g < g_end:
        col = g // To
        t0 = g - col * To
        nstep = min(g_end - g, To - t0)
        out.append((col, t0, nstep))
        g += nstep
    return out


def simulate(run_lengths, lead):
    """Replay the ring bookkeeping.  ``lead`` = how many steps the producer is ahead of the slowest consumer when it
    loads a package (1 .. D; D is the most the empty barriers allow).  Returns the number of steps checked."""
    # the schedule of steps: for each global step n -> (slots it READS, slots its package WRITES, package slot d)
    steps = []
    pos, nxt = 0, 0
    n = 0
    for L in run_lengths:
        for s in range(L):
            if s == 0:
                pos = nxt
                writes = [pos, (pos + 1) % R, (pos + 2) % R]
            else:
                pos = (pos + 1) % R
                writes = [(pos + 2) % R]
            nxt = (pos + 3) % R
            reads = [pos, (pos + 1) % R, (pos + 2) % R]
            steps.append((reads, writes, n % D))
            n += 1
    # the producer loads package m when steps <= m - lead are complete: steps m - lead + 1 .. m - 1 may still be reading
    for m, (_, writes, d) in enumerate(steps):
        for k in range(max(0, m - lead + 1), m):
            reads_k, _, d_k = steps[k]
            assert not (set(writes) & set(reads_k)), (m, k, writes, reads_k)
            assert d != d_k, (m, k)  # the gz buffer of the package is not the one an in-flight step reads
        # and the planes a step reads are exactly the three most recently allocated positions of its run
        reads_m = steps[m][0]
        assert len(set(reads_m)) == 3
    return len(steps)


def test_ring_constants():
    assert D >= 2, "the producer must be able to run ahead of the consumers"
    assert R >= 3 * D, "ring too small for back-to-back run starts (three new planes per step)"


@pytest.mark.parametrize("pattern", [
    [15] * 8,                 # the BASELINE shape: runs of To = 15 steps
    [1] * 40,                 # every step starts a run: three new planes per step, the worst case for the ring
    [1, 15, 1, 1, 2, 15, 3],  # short runs at the boundaries of a CTA's range
    [2] * 30,
    [3, 1] * 20,
])
def test_no_slot_is_overwritten_while_in_use(pattern):
    for lead in range(1, D + 1):
        assert simulate(pattern, lead) == sum(pattern)


def test_random_run_patterns():
    rng = random.Random(518)
    for _ in range(200):
        pattern = [rng.choice([1, 1, 2, 3, 5, 11, 15, 17]) for _ in range(rng.randint(1, 30))]
        simulate(pattern, D)


def test_a_smaller_ring_would_fail():
    """The model is able to detect the hazard: the same bookkeeping on a ring one slot short collides."""
    global R
    keep = R
    try:
        R = 3 * D - 1
        with pytest.raises(AssertionError):
            simulate([1] * 20, D)
    finally:
        R = keep


@pytest.mark.parametrize("total,ctas,To", [(28800, 148, 15), (28800, 296, 15), (7, 148, 3), (1000, 7, 11), (59, 59, 1)])
def test_both_roles_split_the_steps_identically(total, ctas, To):
    """Every step of the (sample, tile, time) space is processed exactly once, runs never cross a column."""
    ctas = min(ctas, total)
    seen = []
    for cta in range(ctas):
        g_begin, g_end = total * cta // ctas, total * (cta + 1) // ctas
        assert g_end > g_begin, "a CTA without work would never complete its barriers' first phase"
        for col, t0, nstep in runs_of_range(g_begin, g_end, To):
            assert 0 < nstep <= To - t0
            seen.extend(col * To + t0 + s for s in range(nstep))
    assert seen == list(range(total))
