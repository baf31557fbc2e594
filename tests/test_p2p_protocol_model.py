"""Model check of the arrival-counter protocol of the peer-memory result exchange
(csrc/p2p_exchange.cu, api.cu::mcl_concept_scan_sharded_p2p) in its hardest mode: every rank
takes only the rows it merged (MCL_SHARDED_LOCAL_ROWS), so nothing after the merge holds a fast
rank back.

Per step e = 1, 2, ... a rank (stream order)  pushes its pieces into every peer's receive area
[e & 1] and bumps that peer's counter -- one delivery per destination, each arbitrarily late --,
waits for its own counter to reach a target, then merges area [e & 1] (which must hold every
peer's pieces of step e, nothing older, nothing newer).

The library's scheme -- counters alternating with the step's parity, target = steps of that parity
so far * (world - 1) -- must be safe under EVERY schedule; a single monotone counter (the first
implementation) is not: a fast peer's arrival of step e+1 can stand in for a slow peer's of step e.
Random adversarial schedules find that hole within a few hundred runs, which is what shows the
model is sharp enough to mean something when it passes.  No GPU needed."""
import random

import pytest


def run_schedule(world, steps, rng, parity_counters):
    """One random schedule.  Returns the first violation as a string, or None."""
    n_ctr = 2 if parity_counters else 1
    counter = [[0] * n_ctr for _ in range(world)]
    # area[r][parity][src] = step whose pieces of `src` lie in rank r's receive area
    area = [[[0] * world for _ in range(2)] for _ in range(world)]
    in_flight = []                       # (dst, src, step): stores + bump of one destination, not yet landed
    pc = [("push", 1)] * world           # next action of every rank
    done = [False] * world
    while not all(done) or in_flight:
        moves = [("deliver", i) for i in range(len(in_flight))]
        for r in range(world):
            if done[r]:
                continue
            kind, e = pc[r]
            if kind == "push":
                moves.append(("rank", r))
            else:
                c = (e & 1) if parity_counters else 0
                target = ((e + 1) >> 1) * (world - 1) if parity_counters else e * (world - 1)
                if counter[r][c] >= target:
                    moves.append(("rank", r))
        if not moves:
            return "deadlock"
        # adversary: deliveries are lazy, so that ranks run ahead of their peers' stores
        ranks = [m for m in moves if m[0] == "rank"]
        kind, x = rng.choice(ranks) if ranks and rng.random() < 0.8 else rng.choice(moves)
        if kind == "deliver":
            dst, src, e = in_flight.pop(x)
            if area[dst][e & 1][src] > e:
                return f"rank {dst}: pieces of step {e} from {src} overwrote newer ones"
            area[dst][e & 1][src] = e            # the data lands, THEN the counter moves (fence + atomic)
            counter[dst][(e & 1) if parity_counters else 0] += 1
            continue
        r = x
        kind, e = pc[r]
        if kind == "push":
            for dst in range(world):
                if dst != r:
                    in_flight.append((dst, r, e))
            pc[r] = ("merge", e)
        else:                                    # the wait was satisfied: merge area [e & 1]
            for src in range(world):
                if src != r and area[r][e & 1][src] != e:
                    return (f"rank {r} merged step {e} with pieces of step {area[r][e & 1][src]} "
                            f"from rank {src}")
            if e == steps:
                done[r] = True
            else:
                pc[r] = ("push", e + 1)
    return None


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_parity_counters_are_safe_under_random_schedules(world):
    rng = random.Random(1000 + world)
    for _ in range(400):
        assert run_schedule(world, steps=6, rng=rng, parity_counters=True) is None


def test_single_counter_is_not():
    """The model finds the race of a single monotone counter (three or more ranks: with two the
    only peer's arrivals come in order)."""
    rng = random.Random(7)
    found = None
    for _ in range(3000):
        found = run_schedule(4, steps=6, rng=rng, parity_counters=False)
        if found:
            break
    assert found and "merged step" in found


def test_targets_match_the_library_formula():
    # api.cu: xcount_target = ((epoch + 1) >> 1) * (world - 1) on counter [epoch & 1]
    for world in (2, 4, 8):
        seen = [0, 0]
        for e in range(1, 50):
            seen[e & 1] += world - 1
            assert ((e + 1) >> 1) * (world - 1) == seen[e & 1]
