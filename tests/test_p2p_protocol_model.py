"""Model check of the arrival-counter protocol of the peer-memory result exchange
(csrc/p2p_exchange.cu, api.cu::mcl_concept_scan_sharded_p2p).

Per step e = 1, 2, ... every rank executes, in stream order,
    push 1   its pieces into every peer's receive area [e & 1] + a bump of that peer's counter
             (one delivery per destination, each arbitrarily late: stores, fence, then the atomic),
    wait x   for its own counter to reach a target,
    merge    of receive area [e & 1] -- which must hold every peer's pieces of step e,
and, unless the step keeps only the rows the rank merged itself (MCL_SHARDED_LOCAL_ROWS),
    push 2   the merged rows into every peer's result area + a bump of the peer's second counter,
    wait f   for the second counter (monotone, counts the full steps),
    copy     of the result area -- which must hold every peer's merged rows of step e.
All ranks run the same sequence of full / local steps.

What is checked under random adversarial schedules (deliveries as late as the waits allow):
  * a merge / copy never reads pieces of another step (stale or newer),
  * no delivery overwrites pieces that their reader has not consumed yet,
  * nothing deadlocks.
The library's scheme for push 1 -- counters alternating with the step's parity, target = steps of
that parity so far * (world - 1) -- must pass; a single monotone counter (the first
implementation) must NOT: with local steps a fast peer's arrival of step e+1 can stand in for a
slow peer's of step e.  That the model finds that hole within a few hundred schedules is what shows
it is sharp enough to mean something when it passes.  No GPU needed."""
import random

import pytest


def run_schedule(world, modes, rng, parity_counters):
    """One random schedule of len(modes) steps (modes[e-1] = True: full step).  Returns the first
    violation as a string, or None."""
    steps = len(modes)
    n_ctr = 2 if parity_counters else 1
    xcount = [[0] * n_ctr for _ in range(world)]
    fcount = [0] * world
    # xin[r][parity][src] / fin[r][src] = step whose pieces of `src` lie in rank r's areas
    xin = [[[0] * world for _ in range(2)] for _ in range(world)]
    fin = [[0] * world for _ in range(world)]
    merged = [0] * world                 # last step whose receive area the rank has read
    copied = [0] * world                 # last FULL step whose result area the rank has read
    fulls_upto = [0] * (steps + 1)       # number of full steps among 1..e
    for e in range(1, steps + 1):
        fulls_upto[e] = fulls_upto[e - 1] + (1 if modes[e - 1] else 0)
    prev_full = [0] * (steps + 1)        # last full step before e
    for e in range(1, steps + 1):
        prev_full[e] = e - 1 if (e > 1 and modes[e - 2]) else (prev_full[e - 1] if e > 1 else 0)
    in_flight = []                       # (phase, dst, src, step): stores + bump of one destination, not landed
    pc = [("push1", 1)] * world          # next action of every rank
    done = [False] * world

    def ready(r):
        kind, e = pc[r]
        if kind in ("push1", "push2"):
            return True
        if kind == "merge":
            c = (e & 1) if parity_counters else 0
            target = ((e + 1) >> 1) * (world - 1) if parity_counters else e * (world - 1)
            return xcount[r][c] >= target
        return fcount[r] >= fulls_upto[e] * (world - 1)          # "copy"

    while not all(done) or in_flight:
        ranks = [r for r in range(world) if not done[r] and ready(r)]
        if not ranks and not in_flight:
            return "deadlock"
        # adversary: deliveries are lazy, so that ranks run ahead of their peers' stores
        if ranks and (not in_flight or rng.random() < 0.8):
            r = rng.choice(ranks)
            kind, e = pc[r]
            if kind == "push1":
                in_flight += [("x", dst, r, e) for dst in range(world) if dst != r]
                pc[r] = ("merge", e)
            elif kind == "merge":
                for src in range(world):
                    if src != r and xin[r][e & 1][src] != e:
                        return f"rank {r} merged step {e} with pieces of step {xin[r][e & 1][src]} from rank {src}"
                merged[r] = e
                pc[r] = ("push2", e) if modes[e - 1] else ("push1", e + 1)
                if not modes[e - 1] and e == steps:
                    done[r] = True
            elif kind == "push2":
                in_flight += [("f", dst, r, e) for dst in range(world) if dst != r]
                pc[r] = ("copy", e)
            else:
                for src in range(world):
                    if src != r and fin[r][src] != e:
                        return f"rank {r} copied step {e} with merged rows of step {fin[r][src]} from rank {src}"
                copied[r] = e
                pc[r] = ("push1", e + 1)
                if e == steps:
                    done[r] = True
            continue
        phase, dst, src, e = in_flight.pop(rng.randrange(len(in_flight)))
        if phase == "x":
            if e - 2 > merged[dst]:
                return f"rank {dst}: pieces of step {e} from {src} overwrote step {e - 2} before it was merged"
            xin[dst][e & 1][src] = e             # the data lands, THEN the counter moves (fence + atomic)
            xcount[dst][(e & 1) if parity_counters else 0] += 1
        else:
            if prev_full[e] > copied[dst]:
                return f"rank {dst}: merged rows of step {e} from {src} overwrote step {prev_full[e]} before the copy"
            fin[dst][src] = e
            fcount[dst] += 1
    return None


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("kind", ["local", "full", "mixed"])
def test_parity_counters_are_safe_under_random_schedules(world, kind):
    rng = random.Random(1000 + world)
    for i in range(250):
        modes = {"local": [False] * 6, "full": [True] * 6,
                 "mixed": [rng.random() < 0.5 for _ in range(7)]}[kind]
        assert run_schedule(world, modes, rng, parity_counters=True) is None, (i, modes)


def test_single_counter_is_not():
    """The model finds the race of a single monotone counter in local-rows mode (three or more
    ranks: with two the only peer's arrivals come in order)."""
    rng = random.Random(7)
    found = None
    for _ in range(3000):
        found = run_schedule(4, [False] * 6, rng, parity_counters=False)
        if found:
            break
    assert found and "merged step" in found


def test_single_counter_suffices_when_every_step_is_full():
    """... and only there: the second wait of a full step holds every rank back until all merged
    rows have arrived, so no peer runs ahead -- which is why the SECOND counter may stay monotone."""
    rng = random.Random(9)
    for _ in range(300):
        assert run_schedule(4, [True] * 5, rng, parity_counters=False) is None


def test_targets_match_the_library_formula():
    # api.cu: xcount_target = ((epoch + 1) >> 1) * (world - 1) on counter [epoch & 1]
    for world in (2, 4, 8):
        seen = [0, 0]
        for e in range(1, 50):
            seen[e & 1] += world - 1
            assert ((e + 1) >> 1) * (world - 1) == seen[e & 1]
