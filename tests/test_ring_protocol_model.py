"""A model of the per-warp stream ring of the K1 kernels (tiles.cu / rowsum.cu: `issue` / `step`)
under random completion times of the asynchronous copies.  Not the CUDA code -- the PROTOCOL:

* a ring of two stages, each with an mbarrier of count 1; the warp's program order is
  issue(0), issue(1), then per piece: wait on the stage's phase parity, read the stage, __syncwarp,
  re-issue the stage with the piece two ahead;
* `issue` arms the barrier with the byte count and starts a copy whose bytes land one chunk at a
  time, at any later moment (the scheduler decides); the phase completes when all bytes are in;
* an empty piece is neither armed nor waited for and does not flip the parity bit.

Checked: a read always finds the complete bytes of the piece it expects, no chunk ever lands in a
stage the warp has not finished reading, and the schedule terminates.  Two mutated protocols -- a
copy issued three ahead into a ring of two, and a parity bit flipped for empty pieces -- are broken
by the same schedules."""
import random

import pytest


def run(seed, ahead=2, flip_on_empty=False):
    rng = random.Random(seed)
    nstage = 2
    npieces = rng.randint(1, 14)
    sizes = [0 if rng.random() < 0.25 else rng.randint(1, 4) for _ in range(npieces)]   # chunks per piece
    stage_data = [[None] * 4 for _ in range(nstage)]      # what each chunk slot holds: (piece, chunk)
    phase = [0] * nstage                                  # index of the current (incomplete) phase
    pending = [0] * nstage                                # chunks the armed phase still waits for
    in_flight = []                                        # (stage, piece, chunk)
    reading = [None]                                      # stage the warp reads right now
    parity = [0] * nstage

    def issue(stage, piece):
        if sizes[piece] == 0:
            return
        assert pending[stage] == 0, "stage armed twice"
        pending[stage] = sizes[piece]
        for c in range(sizes[piece]):
            in_flight.append((stage, piece, c))

    def land_one():
        k = rng.randrange(len(in_flight))
        stage, piece, c = in_flight.pop(k)
        assert reading[0] != stage, "a copy lands in a stage that is being read"
        stage_data[stage][c] = (piece, c)
        pending[stage] -= 1
        if pending[stage] == 0:
            phase[stage] += 1                             # complete_tx: the phase is over

    def wait(stage, par):
        spins = 0
        while phase[stage] % 2 == par:                    # the phase of this parity is not complete yet
            assert in_flight, "wait can never be satisfied"
            land_one()
            spins += 1
            assert spins < 10000
        # between any two instructions of the warp copies may also make progress
        while in_flight and rng.random() < 0.5:
            land_one()

    for p in range(min(ahead, npieces)):
        issue(p % nstage, p)
    for ci in range(npieces):
        stage = ci % nstage
        while in_flight and rng.random() < 0.3:
            land_one()
        if sizes[ci]:
            wait(stage, parity[stage])
            parity[stage] ^= 1
            reading[0] = stage
            for c in range(sizes[ci]):
                assert stage_data[stage][c] == (ci, c), "read of bytes that are not this piece's"
                while in_flight and rng.random() < 0.3:   # copies of OTHER stages land meanwhile
                    land_one()
            reading[0] = None                             # __syncwarp: every lane has read the stage
        elif flip_on_empty:
            parity[stage] ^= 1
        if ci + ahead < npieces:
            issue((ci + ahead) % nstage, ci + ahead)
    while in_flight:
        land_one()


def test_ring_under_random_completion_times():
    for seed in range(3000):
        run(seed)


def test_the_ring_model_has_teeth():
    def breaks(**kw):
        for seed in range(400):
            try:
                run(seed, **kw)
            except AssertionError:
                return True
        return False
    assert breaks(ahead=3)                # a third copy in flight with two stages
    assert breaks(flip_on_empty=True)     # parity out of step with the barrier's phase
    assert not breaks()
