"""Protocol model of the CTA-pair MLP kernels (spin-nerf_b200/csrc/mlp_tc.cu: mlp_fwd_kernel, mlp_tc_bwd.cu: mlp_dgrad_kernel).

The kernels synchronise a weight producer (6 lanes), a peer-CTA relay, one MMA-issuing thread and two epilogue groups per
CTA through mbarriers that are waited on by PARITY.  A parity wait is only correct if the waiter is never more than one
phase behind the barrier — the bug class that does not show up in a parity test of the outputs unless the timing happens
to hit it (one such bug — a producer lane that skipped a ring revolution — was found on the GPU while these kernels were
written).  This test restates the loops of every role with the kernels' own index / phase formulas on top of an exact
mbarrier model (pending count, tx count, phase) and runs them as a randomised discrete-event simulation:

  * every parity wait that passes must have been satisfied by exactly the phase the role meant to wait for,
  * every MMA must find, in BOTH CTAs, the weight chunk of its own layer in the ring slot it reads,
  * no ring slot is overwritten while a tile slot still has to read it, and nothing deadlocks,

for many rounds and random latencies of copies, commits, remote arrives and epilogues.  CPU only.
"""
import heapq
import random

import pytest

K_SLOTS = 3          # ring slots (one group = two half-chunks each)
FWD_CHUNKS = [1, 4, 4, 4, 4, 4, 1, 4, 4, 4, 4, 1]       # c_step_chunks (mlp_tc.cu)
DG_CHUNKS = [2, 4, 4, 4, 4, 4, 4, 4, 4]                 # c_dg_chunks (mlp_tc_bwd.cu)


class Barrier:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _maybe_complete(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self, expect_tx=0):
        assert self.pending > 0, "more arrivals than the barrier was initialised for"
        self.tx += expect_tx
        self.pending -= 1
        self._maybe_complete()

    def complete_tx(self, n):
        self.tx -= n
        self._maybe_complete()

    def parity_ready(self, parity):
        return (self.phase & 1) != parity


class Sim:
    """Cooperative threads (generators).  A thread yields ("wait", barrier, parity, meant_phase) or ("sleep", dt)."""

    def __init__(self, seed):
        self.rng = random.Random(seed)
        self.now, self.seq = 0.0, 0
        self.events = []          # (time, seq, callable)
        self.waiting = []         # (thread, barrier, parity, meant_phase)
        self.threads = {}
        self.done = set()

    def at(self, dt, fn):
        self.seq += 1
        heapq.heappush(self.events, (self.now + dt, self.seq, fn))

    def spawn(self, name, gen):
        self.threads[name] = gen
        self.at(0.0, lambda: self._step(name))

    def _step(self, name):
        gen = self.threads[name]
        try:
            req = next(gen)
        except StopIteration:
            self.done.add(name)
            return
        if req[0] == "sleep":
            self.at(req[1], lambda: self._step(name))
        else:
            _, bar, parity, meant = req
            self.waiting.append((name, bar, parity, meant))
            self._poll()

    def _poll(self):
        still = []
        for name, bar, parity, meant in self.waiting:
            if bar.parity_ready(parity):
                # the wait passes: it must be because phase `meant` (0-based) has completed and the barrier is no further
                assert bar.phase == meant + 1, f"{name}: parity wait satisfied by phase {bar.phase - 1}, meant {meant}"
                self.at(self.rng.uniform(20, 200), lambda n=name: self._step(n))      # wake-up latency
            else:
                still.append((name, bar, parity, meant))
        self.waiting = still

    def run(self):
        while self.events:
            t, _, fn = heapq.heappop(self.events)
            self.now = t
            fn()
            self._poll()
        assert not self.waiting, "deadlock: " + ", ".join(w[0] for w in self.waiting)
        assert self.done == set(self.threads), "threads did not finish: " + str(set(self.threads) - self.done)


def simulate(chunks, rounds, seed, fwd_phase_formula, lane_skips_wait_bug=False):
    """chunks: per-step chunk counts; fwd_phase_formula: True -> group barrier index s % 4 with phase it*(n/4) + s//4
    (mlp_fwd_kernel), False -> running step counter gs: index gs % 4, phase gs // 4 (mlp_dgrad_kernel)."""
    sim = Sim(seed)
    rng = sim.rng
    n_steps = len(chunks)
    if fwd_phase_formula:
        assert n_steps % 4 == 0

    def gidx(it, s):                       # (barrier index, phase number meant) of step s in round it
        if fwd_phase_formula:
            return s & 3, it * (n_steps // 4) + (s >> 2)
        gs = it * n_steps + s
        return gs & 3, gs >> 2

    cta = []
    for rank in (0, 1):
        cta.append(dict(full=[[Barrier(2 if rank == 0 else 1) for _ in range(4)] for _ in range(2)],
                        empty=[Barrier(1) for _ in range(K_SLOTS)], acc=[Barrier(1), Barrier(1)],
                        ring=[[None, None] for _ in range(K_SLOTS)], readers_left=[[0, 0] for _ in range(K_SLOTS)]))
    act = [Barrier(4), Barrier(4)]          # leader's act_ready[t]: (2 column halves folded into) 2 warps x 2 CTAs

    def producer(rank, lane):
        me = cta[rank]
        slot_of_lane, sub = lane % K_SLOTS, lane // K_SLOTS
        grp = 0
        for it in range(rounds):
            for s in range(n_steps):
                nch = chunks[s]
                for g in range((nch + 1) // 2):
                    stage, rev = grp % K_SLOTS, grp // K_SLOTS
                    if stage == slot_of_lane:
                        in_group = min(2, nch - 2 * g)
                        bi, ph = gidx(it, s)
                        # both lanes of the slot wait for EVERY release (phase ^ 1 parity, release number rev - 1)
                        if not (lane_skips_wait_bug and sub >= in_group):
                            yield ("wait", me["empty"][stage], (rev & 1) ^ 1, rev - 1)
                        if sub == 0:
                            me["full"][g][bi].arrive(expect_tx=in_group)
                            if g == 0 and nch <= 2:
                                me["full"][1][bi].arrive()
                        if sub < in_group:
                            assert me["readers_left"][stage][sub] == 0, "slot overwritten while a tile slot still reads it"
                            tag = (it, s, 2 * g + sub)

                            def land(me=me, stage=stage, sub=sub, tag=tag, bar=me["full"][g][bi]):
                                me["ring"][stage][sub] = tag
                                me["readers_left"][stage][sub] = 2          # tile slots 0 and 1 of this CTA pair
                                bar.complete_tx(1)
                            sim.at(rng.uniform(300, 2500), land)
                            yield ("sleep", rng.uniform(5, 700))             # bulk-copy issue is slow per thread
                    grp += 1

    def relay():
        peer, leader = cta[1], cta[0]
        for it in range(rounds):
            for s in range(n_steps):
                bi, ph = gidx(it, s)
                for g in range(2):
                    yield ("wait", peer["full"][g][bi], ph & 1, ph)
                    sim.at(rng.uniform(100, 900), lambda b=leader["full"][g][bi]: b.arrive())

    def issuer():
        leader = cta[0]
        grp = 0
        act_phase = [0, 0]
        for it in range(rounds):
            for s in range(n_steps):
                nch = chunks[s]
                bi, ph = gidx(it, s)
                yield ("wait", leader["full"][0][bi], ph & 1, ph)
                for t in range(2):
                    yield ("wait", act[t], act_phase[t] & 1, act_phase[t])
                    act_phase[t] += 1
                    for c in range(nch):
                        stage = (grp + (c >> 1)) % K_SLOTS
                        if t == 0 and c == 2:
                            yield ("wait", leader["full"][1][bi], ph & 1, ph)
                        for me in cta:                                       # the pair's MMA reads both CTAs' half chunks
                            assert me["ring"][stage][c & 1] == (it, s, c), \
                                f"MMA of step {(it, s, c)} tile slot {t} found {me['ring'][stage][c & 1]} in the ring"
                            me["readers_left"][stage][c & 1] -= 1
                        yield ("sleep", rng.uniform(400, 700))               # four blocking MMA issues
                        if t == 1 and ((c & 1) or c == nch - 1):
                            def release(stage=stage):
                                for me in cta:
                                    me["empty"][stage].arrive()
                            sim.at(rng.uniform(50, 600), release)            # commit: after the MMAs have completed
                    sim.at(rng.uniform(300, 900), lambda t=t: [me["acc"][t].arrive() for me in cta])
                grp += (nch + 1) >> 1

    def epilogue(rank, t):
        me = cta[rank]
        lat = (lambda: rng.uniform(50, 800)) if rank == 1 else (lambda: rng.uniform(5, 50))   # remote vs local arrive
        for it in range(rounds):
            yield ("sleep", rng.uniform(500, 4000))                          # prologue: encode the tile
            for w in range(2):
                sim.at(lat(), lambda: act[t].arrive())
            for s in range(n_steps):
                yield ("wait", me["acc"][t], (it * n_steps + s) & 1, it * n_steps + s)
                yield ("sleep", rng.uniform(300, 4000))                      # drain TMEM, write the next A tile
                if s != n_steps - 1:                                         # the last step is followed by the next prologue
                    for w in range(2):
                        sim.at(lat(), lambda: act[t].arrive())

    for rank in (0, 1):
        for lane in range(2 * K_SLOTS):
            sim.spawn(f"producer{rank}.{lane}", producer(rank, lane))
        for t in (0, 1):
            sim.spawn(f"epilogue{rank}.{t}", epilogue(rank, t))
    sim.spawn("relay", relay())
    sim.spawn("issuer", issuer())
    sim.run()
    return sim.now


@pytest.mark.parametrize("seed", range(8))
def test_forward_pair_protocol(seed):
    assert simulate(FWD_CHUNKS, rounds=7, seed=seed, fwd_phase_formula=True) > 0


@pytest.mark.parametrize("seed", range(8))
def test_dgrad_pair_protocol(seed):
    assert simulate(DG_CHUNKS, rounds=9, seed=100 + seed, fwd_phase_formula=False) > 0


def test_model_flags_the_revolution_skipping_producer():
    """With the original producer (second lane of a slot only waits when its group has a second chunk) the simulation must
    fail: that lane's next parity wait is satisfied by an older release and it overwrites a slot that is still being read."""
    failures = 0
    for seed in range(6):
        try:
            simulate(FWD_CHUNKS, rounds=5, seed=seed, fwd_phase_formula=True, lane_skips_wait_bug=True)
        except AssertionError:
            failures += 1
    assert failures == 6


def test_model_catches_a_lane_that_skips_a_ring_revolution():
    """The bug found on the GPU: the second lane of a slot did not wait for the slot's release when its group had a single
    chunk, so its next parity wait was satisfied by the release before last.  The barrier model must flag exactly that."""
    sim = Sim(0)
    b = Barrier(1)
    b.arrive()                         # one release has happened (phase 1) ...

    def lane():
        yield ("wait", b, 0, 2)        # ... the lane waits for the THIRD one with parity 0: passes at once, wrongly
    sim.spawn("lane", lane())
    with pytest.raises(AssertionError, match="meant 2"):
        sim.run()
