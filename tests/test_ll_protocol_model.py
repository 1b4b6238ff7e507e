"""A model of the row-sum exchange between GPUs (k_ice_exchange, LL form: kernels.cu st_ll / ld_ll)
run under random schedules on the CPU.  Not the CUDA code -- the PROTOCOL it implements, with the
memory system at its most adversarial: every 8-byte word a rank stores into a peer is delivered on
its own, at any later time, in any order (also across rows, passes and the two halves of one value).

Each rank, per pass p: (1) stores the sums of its own rows into buffer p & 1 of EVERY rank as two
words (tag(p) | low half, tag(p) | high half); (2) reads all rows of its own buffer p & 1, accepting
a row once BOTH words carry tag(p); (3) only then goes on to pass p + 1.  Checked on every accepted
row: the value is the one the owner computed for THIS pass (no torn or stale value), and -- the
hazard two alternating buffers must exclude -- no word of pass p + 2 is ever delivered into a buffer
whose rank has not finished reading pass p.  Also checked: every schedule terminates (no deadlock)."""
import random

import pytest


def value_of(rank, row, epoch):
    return (rank * 1000003 + row * 7919 + epoch * 104729) & 0xFFFFFFFFFFFFFFFF


def run(seed, world, rows_per_rank, passes, epoch_base, nbuf=2, tagged=True):
    rng = random.Random(seed)
    n = world * rows_per_rank
    owner = [r // rows_per_rank for r in range(n)]
    # buf[rank][parity][row] = [word0, word1]; a word = (tag, half); leftovers of an earlier call
    buf = [[[[(epoch_base - 1 - par, 0xDEAD), (epoch_base - 1 - par, 0xBEEF)] for _ in range(n)]
            for par in range(2)] for _ in range(world)]
    in_flight = []                                # (dest rank, parity, row, word index, tag, half)
    sent = [0] * world                            # passes whose stores have been issued
    done_reading = [0] * world                    # passes whose statistics are complete
    pending = [set() for _ in range(world)]       # rows still to accept in the current pass
    steps = 0
    while min(done_reading) < passes:
        steps += 1
        assert steps < 2000000, "schedule does not terminate"
        actions = []
        for a in range(world):
            if sent[a] == done_reading[a] and sent[a] < passes:
                actions.append(("send", a))
            if sent[a] > done_reading[a]:
                actions.append(("read", a))
        if in_flight:
            actions.extend([("deliver", None)] * 3)
        kind, a = rng.choice(actions)
        if kind == "send":
            p = sent[a]
            tag = epoch_base + p + 1
            for row in range(a * rows_per_rank, (a + 1) * rows_per_rank):
                v = value_of(a, row, tag)
                for dest in range(world):
                    in_flight.append((dest, p % nbuf, row, 0, tag, v & 0xFFFFFFFF))
                    in_flight.append((dest, p % nbuf, row, 1, tag, v >> 32))
            sent[a] += 1
            pending[a] = set(range(n))
        elif kind == "deliver":
            k = rng.randrange(len(in_flight))
            dest, par, row, w, tag, half = in_flight.pop(k)
            # the pass this word belongs to is tag - epoch_base - 1; its buffer must be free:
            # the destination has finished reading the pass that used this buffer before
            assert done_reading[dest] >= tag - epoch_base - 1 - 1, "buffer overwritten while still read"
            buf[dest][par][row][w] = (tag, half)
        else:
            p = done_reading[a]
            tag = epoch_base + p + 1
            row = rng.choice(sorted(pending[a]))
            w0, w1 = buf[a][p % nbuf][row]
            if (w0[0] == tag and w1[0] == tag) if tagged else (w0[0] == tag):   # ld_ll: both tags current
                got = (w1[1] << 32) | w0[1]
                assert got == value_of(owner[row], row, tag), "torn or stale value accepted"
                pending[a].discard(row)
                if not pending[a]:
                    done_reading[a] += 1
    assert not in_flight or all(t <= epoch_base + passes for (_, _, _, _, t, _) in in_flight)
    return steps


@pytest.mark.parametrize("world", [2, 3, 4])
def test_ll_exchange_under_random_schedules(world):
    for seed in range(60):
        run(seed, world, rows_per_rank=3, passes=5, epoch_base=10 + seed)


def test_ll_exchange_a_second_call_never_matches_leftover_words():
    """Two normalisations in a row on the same buffers: the second starts from the epoch where the
    first ended (tb_store keeps px.epoch growing), so words left from the first never carry a
    current tag."""
    for seed in range(40):
        run(1000 + seed, 2, rows_per_rank=4, passes=4, epoch_base=100)
        run(2000 + seed, 2, rows_per_rank=4, passes=4, epoch_base=100 + 4 + 1)


def test_the_model_has_teeth():
    """The same schedules break a protocol with ONE buffer (a fast rank overwrites what a slow one
    still reads) and one that trusts the tag of the first word only (torn values)."""
    def breaks(**kw):
        for seed in range(200):
            try:
                run(seed, 3, rows_per_rank=3, passes=5, epoch_base=10, **kw)
            except AssertionError:
                return True
        return False
    assert breaks(nbuf=1)
    assert breaks(tagged=False)
    assert not breaks()
