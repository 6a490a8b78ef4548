"""Model check of the synchronisation protocol of the experimental 2-CTA Ozaki kernel (csrc/ozaki_gemm2.cuh).

The kernel has not run on hardware yet, and its risk is not the arithmetic (shared with the validated 1-CTA kernel) but
the plumbing between the two CTAs of a pair: per-CTA full barriers, the peer's relay warp (or, on the tensor-map load path,
both CTAs' copies completing the leader's barrier), multicast commits, the tempty barrier that lives in the leader.
This test restates that plumbing — the loops of the five warp roles, line for line, with mbarrier semantics (arrival counts, transaction bytes, phase parity), asynchronous bulk copies that land in any
order, MMAs that execute in issue order some time after they were issued, and commits that arrive once all earlier
MMAs are done — and runs it under many random interleavings.  It asserts
  * no deadlock: every role of both CTAs finishes;
  * every MMA finds, in BOTH CTAs' shared memory, exactly the digit tiles of its (tile, pass, k step) — nothing is
    overwritten before the tensor core has read it, nothing is consumed before it has landed;
  * the epilogue of each CTA reads complete accumulators of the right (tile, pass), and no MMA overwrites tensor memory
    that an epilogue warp of either CTA has not finished reading.
It is a model (Python, no GPU): it proves the protocol, not the PTX.  tools/ozaki_variants.py is the hardware test."""
import random

import pytest

GROUPS_PER_PASS, MAX_S, STAGES = 4, 8, 4


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.phase ^= 1
            self.pending = self.count

    def arrive(self):
        assert self.pending > 0, "more arrivals than the barrier expects in one phase"
        self.pending -= 1
        self._check()

    def expect_tx(self, nbytes):  # mbarrier.arrive.expect_tx
        self.tx += nbytes
        self.arrive()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._check()

    def done(self, parity):  # try_wait.parity: has the phase with this parity completed?
        return self.phase != parity


class Cta:
    def __init__(self):
        self.full = [MBar(1) for _ in range(STAGES)]
        self.empty = [MBar(1) for _ in range(STAGES)]
        self.pfull = [MBar(1) for _ in range(STAGES)]
        self.tfull, self.tempty = MBar(1), MBar(8)
        self.smem = [dict() for _ in range(STAGES)]  # stage -> {("A"|"B", sub-slot h): (tile, pass, k step)}
        self.tmem = [None] * GROUPS_PER_PASS          # slot -> [unit, MMAs accumulated]


class Sim:
    def __init__(self, S, ksteps, tiles, seed, tma=False):
        self.S, self.ksteps, self.tiles, self.tma = S, ksteps, tiles, tma
        self.npass = (S + GROUPS_PER_PASS - 1) // GROUPS_PER_PASS
        self.cta = [Cta(), Cta()]
        self.rng = random.Random(seed)
        self.loads = []       # bulk copies in flight (complete in any order)
        self.tensor = []      # tensor-core queue of the pair: MMAs and commits, executed in issue order
        self.reads_done = {}  # (rank, unit) -> epilogue warps that finished reading
        self.finished = set()

    def pass_shape(self, ps):
        g_hi = self.S + 1 - GROUPS_PER_PASS * ps
        g_lo = max(2, g_hi - GROUPS_PER_PASS + 1)
        d_hi = min(self.S, g_hi - 1)
        sub = 2 if 2 * d_hi <= MAX_S else 1
        return g_hi, g_lo, d_hi, sub

    # ---- the warp roles: generators that yield ("wait", barrier, parity) or ("step",) ----
    def producer(self, rank):
        me = self.cta[rank]
        stage, phase = 0, 0
        for tile in range(self.tiles):
            for ps in range(self.npass):
                _, _, d_hi, sub = self.pass_shape(ps)
                for ks in range(0, self.ksteps, sub):
                    nsub = min(sub, self.ksteps - ks)
                    yield ("wait", me.empty[stage], phase ^ 1)
                    a_bytes, b_bytes = d_hi * 4096, d_hi * 2048
                    if self.tma:  # cp.async.bulk.tensor.cta_group::2: both CTAs' bytes complete the LEADER's barrier
                        if rank == 0:
                            me.full[stage].expect_tx(2 * (a_bytes + b_bytes) * nsub)
                        signal = 0
                    else:
                        me.full[stage].expect_tx((a_bytes + b_bytes) * nsub)
                        signal = rank
                    for h in range(nsub):
                        self.loads.append((rank, stage, ("A", h), (tile, ps, ks + h), a_bytes, signal))
                        self.loads.append((rank, stage, ("B", h), (tile, ps, ks + h), b_bytes, signal))
                    yield ("step",)
                    stage += 1
                    if stage == STAGES:
                        stage, phase = 0, phase ^ 1
        self.finished.add(("producer", rank))

    def relay(self):  # warp 1 of the peer CTA
        me, leader = self.cta[1], self.cta[0]
        stage, phase = 0, 0
        for tile in range(0 if self.tma else self.tiles):  # the tensor-map path needs no relay
            for ps in range(self.npass):
                _, _, _, sub = self.pass_shape(ps)
                for ks in range(0, self.ksteps, sub):
                    yield ("wait", me.full[stage], phase)
                    leader.pfull[stage].arrive()  # remote mbarrier.arrive.release.cluster
                    yield ("step",)
                    stage += 1
                    if stage == STAGES:
                        stage, phase = 0, phase ^ 1
        self.finished.add(("relay", 1))

    def mma(self):  # warp 1 of the leader CTA
        me = self.cta[0]
        stage, phase, unit = 0, 0, 0
        for tile in range(self.tiles):
            for ps in range(self.npass):
                g_hi, g_lo, d_hi, sub = self.pass_shape(ps)
                yield ("wait", me.tempty, (unit & 1) ^ 1)
                for ks in range(0, self.ksteps, sub):
                    nsub = min(sub, self.ksteps - ks)
                    yield ("wait", me.full[stage], phase)
                    if not self.tma:
                        yield ("wait", me.pfull[stage], phase)
                    for h in range(nsub):
                        first = (ks + h) > 0
                        for gi in range(GROUPS_PER_PASS):
                            for t in range(1, MAX_S + 1):
                                g = g_hi - gi
                                u = g - t
                                if g >= g_lo and t <= self.S and 1 <= u <= self.S:
                                    accumulate = True if t > max(1, g - self.S) else first
                                    self.tensor.append(("mma", stage, h, (tile, ps, ks + h), g - g_lo, unit, accumulate, t <= d_hi and u <= d_hi))
                    self.tensor.append(("commit", "empty", stage))
                    yield ("step",)
                    stage += 1
                    if stage == STAGES:
                        stage, phase = 0, phase ^ 1
                self.tensor.append(("commit", "tfull", None))
                yield ("step",)
                unit += 1
        self.finished.add(("mma", 0))

    def epilogue(self, rank, warp):
        me, leader = self.cta[rank], self.cta[0]
        unit = 0
        for tile in range(self.tiles):
            for ps in range(self.npass):
                g_hi, g_lo, _, _ = self.pass_shape(ps)
                yield ("wait", me.tfull, unit & 1)
                for g in range(g_lo, g_hi + 1):  # tcgen05.ld of every group accumulator of the pass
                    pairs = sum(1 for t in range(1, self.S + 1) if 1 <= g - t <= self.S)
                    assert me.tmem[g - g_lo] == [unit, pairs * self.ksteps], f"CTA {rank} warp {warp}: accumulator of unit {unit} group {g} is {me.tmem[g - g_lo]}"
                    yield ("step",)  # other roles may run between two loads: a premature MMA would be seen
                self.reads_done[(rank, unit)] = self.reads_done.get((rank, unit), 0) + 1
                leader.tempty.arrive()  # local for the leader, remote for the peer
                yield ("step",)
                unit += 1
        self.finished.add(("epilogue", rank, warp))

    # ---- asynchronous hardware ----
    def land_a_load(self):
        rank, stage, part, tag, nbytes, signal = self.loads.pop(self.rng.randrange(len(self.loads)))
        self.cta[rank].smem[stage][part] = tag
        self.cta[signal].full[stage].complete_tx(nbytes)

    def run_tensor_op(self):
        op = self.tensor.pop(0)
        if op[0] == "commit":
            _, which, stage = op
            for c in self.cta:  # multicast, mask 0b11
                (c.empty[stage] if which == "empty" else c.tfull).arrive()
            return
        _, stage, h, tag, slot, unit, accumulate, digits_staged = op
        assert digits_staged, "an MMA uses a digit that its pass does not stage"
        for rank, c in enumerate(self.cta):  # the MMA reads A and its half of B from BOTH CTAs' shared memory
            assert c.smem[stage].get(("A", h)) == tag and c.smem[stage].get(("B", h)) == tag, \
                f"MMA of {tag} found {c.smem[stage]} in stage {stage} of CTA {rank}"
            if not accumulate:
                if unit > 0:
                    for r in (0, 1):
                        assert self.reads_done.get((r, unit - 1), 0) == 4, f"MMA of unit {unit} overwrites accumulators CTA {r} is still reading"
                c.tmem[slot] = [unit, 1]
            else:
                assert c.tmem[slot] is not None and c.tmem[slot][0] == unit, f"accumulating into {c.tmem[slot]} during unit {unit}"
                c.tmem[slot][1] += 1

    def run(self):
        actors = {("producer", 0): self.producer(0), ("producer", 1): self.producer(1), ("relay", 1): self.relay(), ("mma", 0): self.mma()}
        for rank in (0, 1):
            for w in range(4):
                actors[("epilogue", rank, w)] = self.epilogue(rank, w)
        blocked = {}
        expected = set(actors)
        while True:
            choices = []
            for name, gen in actors.items():
                if name in blocked:
                    bar, parity = blocked[name]
                    if not bar.done(parity):
                        continue
                choices.append(("actor", name))
            if self.loads:
                choices.append(("load", None))
            if self.tensor:
                choices.append(("tensor", None))
            if not choices:
                break
            kind, name = self.rng.choice(choices)
            if kind == "load":
                self.land_a_load()
            elif kind == "tensor":
                self.run_tensor_op()
            else:
                blocked.pop(name, None)
                try:
                    ev = next(actors[name])
                    if ev[0] == "wait":
                        blocked[name] = (ev[1], ev[2])
                except StopIteration:
                    del actors[name]
        assert self.finished == expected, f"deadlock: {sorted(expected - self.finished)} never finished (blocked: {sorted(blocked)})"
        assert not self.loads and not self.tensor


@pytest.mark.parametrize("tma", [False, True])
@pytest.mark.parametrize("S", [8, 7])
@pytest.mark.parametrize("ksteps,tiles", [(4, 1), (4, 3), (8, 2), (12, 2)])
def test_protocol_under_random_interleavings(S, ksteps, tiles, tma):
    for seed in range(25):
        Sim(S, ksteps, tiles, seed, tma).run()


def test_the_model_catches_a_missing_relay():
    """Sanity of the model itself: if the leader did not wait for the peer's operands, some interleaving must be caught."""

    class NoPeerWait(Sim):
        def mma(self):
            for ev in super().mma():
                if ev[0] == "wait" and any(ev[1] is b for b in self.cta[0].pfull):
                    continue  # skip the wait on peer_full
                yield ev

        def relay(self):  # keep the barrier phases moving so the only difference is the missing wait
            yield from super().relay()

    caught = 0
    for seed in range(40):
        try:
            NoPeerWait(8, 8, 2, seed).run()
        except AssertionError:
            caught += 1
    assert caught > 0
