"""Executable models of two pieces of kernel control logic: the loop control of K2's accumulate kernel (dbgsom_b200/csrc/accumulate.cu, fourth version) and, at the end of the file, the tensor-memory hand-over of the streamed BMU search.

The kernel's control state is small enough to restate exactly: a prefetch cursor that cuts the sorted sample sequence
into batches (never across a segment boundary), a 64-slot ring of row offsets with slots 0-7 mirrored behind the end,
and a byte FIFO that carries "rows of the batch | 0x80 if it starts a segment" from the prefetch side to the consuming
side.  This model steps one warp through a team's range and checks the invariants the CUDA code relies on: every
position is consumed exactly once and in order, with the right segment; a ring slot is never overwritten while a
position that still has to be read lives in it; reads through the mirror see the right position; the FIFO never holds
more than STAGES - 1 batches.  (The GPU parity tests check the kernel itself; this pins the design on the CPU.)
"""
import numpy as np
import pytest


def run_team(offsets, p_begin, p_end, U, STAGES):
    """Returns the (position, segment) pairs in consumption order."""
    M = len(offsets) - 1
    perm = np.arange(offsets[M])  # identity permutation: slot contents == positions
    # segment containing p_begin: largest j with offsets[j] <= p_begin
    lo = int(np.searchsorted(offsets, p_begin, side="right") - 1)
    pp, pseg, pseg_end, pnew = p_begin, lo, int(offsets[lo + 1]), True
    filled = p_begin
    ring = np.full(72, -1, dtype=np.int64)

    def perm_at(p):
        return int(perm[p if p < p_end else p_end - 1])

    nxt = [perm_at(filled + lane) for lane in range(32)]
    fifo = 0
    staged = {}  # stage -> list of positions copied
    consumed = []

    def issue(stage, fifo_pos):
        nonlocal pp, pseg, pseg_end, pnew, filled, nxt, fifo
        if pp >= p_end:
            return
        while pp >= pseg_end:
            pseg += 1
            pseg_end = int(offsets[pseg + 1])
            pnew = True
        lim = min(p_end, pseg_end)
        nb = min(lim - pp, U)
        assert nb >= 1
        if pp + U > filled:
            busy = {q & 63 for q in range(pp, filled)}  # positions fetched into the ring but not yet read
            for lane in range(32):
                slot = (filled + lane) & 63
                assert slot not in busy, "refill overwrites a slot that is still to be read"
                ring[slot] = nxt[lane]
                if slot < 8:
                    ring[slot + 64] = nxt[lane]
            filled += 32
            nxt = [perm_at(filled + lane) for lane in range(32)]
        assert filled >= pp + U
        base = pp & 63
        rows = []
        for u in range(U):  # all U slots are read, nb rows are copied
            assert base + u < 72
            v = ring[base + u]
            if u < nb:
                assert v == perm_at(pp + u) == pp + u, "ring / mirror returned the wrong position"
                rows.append((int(v), pseg))
        staged[stage] = rows
        assert (fifo >> (8 * fifo_pos)) & 0xFF == 0, "FIFO slot still occupied"
        fifo |= (nb | (0x80 if pnew else 0)) << (8 * fifo_pos)
        pnew = False
        pp += nb

    for s in range(STAGES - 1):
        issue(s, s)
    stage, seg = 0, ~lo
    while fifo & 0x7F:
        nb = fifo & 0x7F
        if fifo & 0x80:
            nseg = ~seg
            if seg >= 0:
                nseg = seg + 1
                while offsets[nseg + 1] == offsets[nseg]:
                    nseg += 1
            seg = nseg
        fifo >>= 8
        assert fifo >> (8 * (STAGES - 2)) == 0, "more than STAGES - 2 batches left after the pop"
        issue(STAGES - 1 if stage == 0 else stage - 1, STAGES - 2)
        rows = staged.pop(stage)
        assert len(rows) == nb
        for pos, pseg_at_issue in rows:
            assert pseg_at_issue == seg, "consuming side disagrees with the prefetch side about the segment"
            assert offsets[seg] <= pos < offsets[seg + 1]
            consumed.append((pos, seg))
        stage = 0 if stage + 1 == STAGES else stage + 1
    return consumed


@pytest.mark.parametrize("U,STAGES", [(4, 4), (8, 4), (2, 3), (2, 4)])
def test_every_position_once_in_order_with_its_segment(U, STAGES):
    rng = np.random.default_rng(U * 10 + STAGES)
    for trial in range(60):
        M = int(rng.integers(1, 40))
        counts = rng.integers(0, 30, size=M)
        counts[rng.random(M) < 0.4] = 0           # many empty segments (dead neurons)
        if trial % 5 == 0:
            counts[rng.integers(0, M)] += 700     # one long segment: the ring wraps many times
        if counts.sum() == 0:
            counts[rng.integers(0, M)] = 1
        offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        total = int(offsets[-1])
        n_teams = int(rng.integers(1, 7))
        seen = []
        for team in range(n_teams):
            p_begin, p_end = total * team // n_teams, total * (team + 1) // n_teams
            if p_begin >= p_end:
                continue
            got = run_team(offsets, p_begin, p_end, U, STAGES)
            assert [p for p, _ in got] == list(range(p_begin, p_end))
            seen.extend(got)
        assert len(seen) == total
        seg_of = np.repeat(np.arange(M), counts)
        assert [s for _, s in seen] == seg_of.tolist()


# ------------------------------------------------------------------------------------------------------------------
# The tensor-memory hand-over of the streamed BMU search (dbgsom_b200/csrc/bmu_tc.cu, SGT > 0): two 256-column regions
# alternate as sum and partial accumulator, guarded by one full / empty mbarrier pair per region and per-region phase
# bits on both sides.  Model: the MMA thread and the epilogue (all its warps move in lock step here) as coroutines over
# mbarriers with the hardware's parity semantics; checked: no deadlock, the MMA stream never overwrites a region whose
# contents are still needed, every tile's final sum holds each of its chains exactly once.
class _Mbar:
    def __init__(self):
        self.phase = 0  # parity of the phase in progress

    def done(self, parity):  # try_wait.parity: has the phase with this parity completed?
        return self.phase != parity

    def complete(self):
        self.phase ^= 1


def _run_hand_over(n_tiles, nseg):
    full, empty = [_Mbar(), _Mbar()], [_Mbar(), _Mbar()]
    region = [None, None]  # contents: None | ("chain", tile, seg) | ("sum", tile, {segs})
    sums = {}

    def mma():
        acc, rph = 0, 0
        for t in range(n_tiles):
            for sg in range(nseg):
                reg = acc if sg == 0 else acc ^ 1
                while not empty[reg].done(((rph >> reg) & 1) ^ 1):
                    yield
                assert region[reg] is None, f"tile {t} chain {sg} overwrites {region[reg]}"
                region[reg] = ("chain", t, sg)
                yield  # the chain's MMAs run
                full[reg].complete()  # tcgen05.commit
                rph ^= 1 << reg
            acc ^= 1

    def epilogue():
        acc, eph = 0, 0
        for t in range(n_tiles):
            R, Q = acc, acc ^ 1
            while not full[R].done((eph >> R) & 1):
                yield
            eph ^= 1 << R
            assert region[R] == ("chain", t, 0)
            region[R] = ("sum", t, {0})
            for sg in range(1, nseg):
                while not full[Q].done((eph >> Q) & 1):
                    yield
                eph ^= 1 << Q
                assert region[Q] == ("chain", t, sg)
                assert sg not in region[R][2]
                region[R][2].add(sg)
                region[Q] = None
                empty[Q].complete()  # all epilogue warps of both CTAs have arrived
                yield
            sums[t] = set(region[R][2])
            region[R] = None
            empty[R].complete()
            acc ^= 1
            yield

    actors = [mma(), epilogue()]
    alive = [True, True]
    idle_rounds = 0
    while any(alive):
        before = (full[0].phase, full[1].phase, empty[0].phase, empty[1].phase, repr(region))
        for i, a in enumerate(actors):
            if alive[i]:
                try:
                    next(a)
                except StopIteration:
                    alive[i] = False
        after = (full[0].phase, full[1].phase, empty[0].phase, empty[1].phase, repr(region))
        idle_rounds = idle_rounds + 1 if before == after else 0
        assert idle_rounds < 8, "deadlock: neither side makes progress"
    return sums


@pytest.mark.parametrize("nseg", [2, 3, 4, 8, 16])
def test_tensor_memory_regions_alternate_without_deadlock(nseg):
    for n_tiles in (1, 2, 3, 7):
        sums = _run_hand_over(n_tiles, nseg)
        assert sorted(sums) == list(range(n_tiles))
        assert all(s == set(range(nseg)) for s in sums.values())
