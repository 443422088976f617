"""Protocol models of the CTA-pair MLP kernels (spin-nerf_b200/csrc/mlp_tc.cu: mlp_fwd_ts_kernel, mlp_tc_bwd.cu: mlp_dgrad_kernel).

The kernels synchronise weight-producer lanes, a peer-CTA relay, one MMA-issuing warp, the epilogue warps and (training) a
stash store lane through mbarriers that are waited on by PARITY.  A parity wait is only correct if the waiter is never more
than one phase behind the barrier — the bug class that does not show up in a parity test of the outputs unless the timing
happens to hit it.  Two such bugs were found on the GPU: a producer lane that skipped a ring revolution (round 1), and the
round-2 forward kernel's weight ring, where the second lane of a slot did not arrive on the slot's full barrier and the
issuer released the ring's empty pad groups without looking at them — one train step in ~1000 deadlocked
(profiles/r02_summary.md).  This file restates the loops of every role with the kernels' own index / phase formulas on top of
an exact mbarrier model (pending count, tx count, phase) and runs them as a randomised discrete-event simulation with
occasional long stalls of single lanes:

  * every parity wait that passes must have been satisfied by exactly the phase the role meant to wait for,
  * every MMA must find, in BOTH CTAs, the weight chunk of its own layer in the ring slot it reads,
  * no ring slot is overwritten while a batch still has to read it, and nothing deadlocks,

for many rounds and random latencies of copies, commits, remote arrives and epilogues; the shipped protocols must pass, the
two buggy ones must be flagged.  CPU only.
"""
import heapq
import random

import pytest

K_SLOTS = 3          # ring slots (one group = two half-chunks each)
R1_FWD_CHUNKS = [1, 4, 4, 4, 4, 4, 1, 4, 4, 4, 4, 1]    # round 1's forward schedule (3-slot ring; the kernel was replaced by mlp_fwd_ts_kernel)
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


def simulate(chunks, rounds, seed, fwd_phase_formula, lane_skips_wait_bug=False, stall_prob=0.0):
    """chunks: per-step chunk counts; fwd_phase_formula: True -> group barrier index s % 4 with phase it*(n/4) + s//4
    (round 1's forward kernel), False -> running step counter gs: index gs % 4, phase gs // 4 (mlp_dgrad_kernel)."""
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
                        if rng.random() < stall_prob:
                            yield ("sleep", rng.uniform(3000, 12000))        # the lane is descheduled for a few microseconds
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
                    if rng.random() < stall_prob:
                        yield ("sleep", rng.uniform(3000, 12000))
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


@pytest.mark.parametrize("seed", range(4))
def test_three_slot_ring_with_odd_chunk_counts(seed):
    """Round 1's forward schedule on the 3-slot ring (one-chunk groups exercise the second producer lane's idle phases)."""
    assert simulate(R1_FWD_CHUNKS, rounds=7, seed=seed, fwd_phase_formula=True) > 0


@pytest.mark.parametrize("seed", range(8))
def test_dgrad_pair_protocol(seed):
    assert simulate(DG_CHUNKS, rounds=9, seed=100 + seed, fwd_phase_formula=False) > 0


def test_dgrad_pair_protocol_under_heavy_stalls():
    for seed in range(4):
        assert simulate(DG_CHUNKS, rounds=7, seed=200 + seed, fwd_phase_formula=False, stall_prob=0.25) > 0
        assert simulate(R1_FWD_CHUNKS, rounds=5, seed=300 + seed, fwd_phase_formula=True, stall_prob=0.25) > 0


def test_model_flags_the_revolution_skipping_producer():
    """With the original producer (second lane of a slot only waits when its group has a second chunk) the simulation must
    fail: that lane's next parity wait is satisfied by an older release and it overwrites a slot that is still being read."""
    failures = 0
    for seed in range(6):
        try:
            simulate(R1_FWD_CHUNKS, rounds=5, seed=seed, fwd_phase_formula=True, lane_skips_wait_bug=True)
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


# ---------------------------------------------------------------------------------------------------------------------
# mlp_fwd_ts_kernel (round 2): 4-slot weight ring with static slots, N-split accumulators, activations in tensor memory
# ---------------------------------------------------------------------------------------------------------------------
TS_SLOTS = 4
TS_G_NCH = [1, 0, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 0, 2, 2, 2, 2, 2, 2, 2, 2, 1, 0]      # c_g_nch: chunks per group, 3 empty pad groups
TS_L_GROUP0 = [0, 2, 4, 6, 8, 10, 14, 16, 18, 20]                                       # c_l_group0
TS_L_KIND = ["L0", "WIDE", "WIDE", "WIDE", "WIDE", "L5", "WIDE", "WIDE", "WIDE", "VIEWS"]  # c_l_kind
TS_REVS = len(TS_G_NCH) // TS_SLOTS                                                     # ring revolutions per round


def ts_reads_of_group():
    """group -> number of MMA batches that read it (4 per 256-wide layer: 2 tile slots x 2 accumulator halves; 2 for the views layer)."""
    reads = [0] * len(TS_G_NCH)
    for L, g0 in enumerate(TS_L_GROUP0):
        nb = 2 if TS_L_KIND[L] == "VIEWS" else 4
        for g in {"L0": [g0], "WIDE": [g0, g0 + 1], "L5": [g0, g0 + 1, g0 + 2], "VIEWS": [g0, g0 + 1, g0 + 2]}[TS_L_KIND[L]]:
            reads[g] = nb
    return reads


def test_ts_tables_are_consistent():
    reads = ts_reads_of_group()
    assert len(TS_G_NCH) % TS_SLOTS == 0
    # every group with chunks is read by some layer, the empty ones by none, and a layer's groups sit in consecutive ring slots
    assert [r > 0 for r in reads] == [n > 0 for n in TS_G_NCH]
    assert all((g0 & 3) in (0, 2) for g0 in TS_L_GROUP0)
    assert sum(TS_G_NCH) == 39                                                          # kFwdChunks (mlp_tc.cuh)


def simulate_ts(rounds, seed, train=True, protocol="shipped", stall_prob=0.02):
    """protocol "shipped": both producer lanes of a slot wait for every release and arrive on every full phase (count 3 on the
    leader, 2 on the peer), the issuer waits for the full barrier of the empty pad groups before it releases their slots.
    protocol "r2_first": only lane (slot, 0) arrives (counts 2 / 1) and empty groups are released unseen — what round 2 first
    shipped.  stall_prob: chance that a producer lane / the relay is descheduled for a few microseconds before a wait."""
    assert protocol in ("shipped", "r2_first")
    sim = Sim(seed)
    rng = sim.rng
    shipped = protocol == "shipped"
    reads = ts_reads_of_group()

    cta = []
    for rank in (0, 1):
        nfull = (3 if rank == 0 else 2) if shipped else (2 if rank == 0 else 1)
        cta.append(dict(full=[Barrier(nfull) for _ in range(TS_SLOTS)], empty=[Barrier(1) for _ in range(TS_SLOTS)],
                        accfull=[Barrier(1), Barrier(1)], written=[Barrier(1) for _ in range(4)], free=[Barrier(1) for _ in range(4)],
                        ring=[[None, None] for _ in range(TS_SLOTS)], readers_left=[[0, 0] for _ in range(TS_SLOTS)]))
    accfree = [Barrier(2), Barrier(2)]      # leader only; the 16 warps of a CTA are folded into one arrival
    aready = [Barrier(2), Barrier(2)]
    remote = lambda: rng.uniform(50, 800)

    def maybe_stall():
        return ("sleep", rng.uniform(3000, 12000) if rng.random() < stall_prob else rng.uniform(0, 30))

    def producer(rank, lane):
        me = cta[rank]
        slot, sub = lane & 3, lane >> 2
        for it in range(rounds):
            for g, nch in enumerate(TS_G_NCH):
                if (g & 3) != slot:
                    continue
                rev = it * TS_REVS + (g >> 2)
                yield maybe_stall()
                yield ("wait", me["empty"][slot], (rev & 1) ^ 1, rev - 1)
                if sub == 0 and nch:
                    me["full"][slot].arrive(expect_tx=nch)
                elif sub == 0 or shipped:
                    me["full"][slot].arrive()
                if sub < nch:
                    assert me["readers_left"][slot][sub] == 0, "ring slot overwritten while a batch still has to read it"
                    tag = (it, g)

                    def land(me=me, slot=slot, sub=sub, tag=tag, g=g):
                        me["ring"][slot][sub] = tag
                        me["readers_left"][slot][sub] = reads[g]
                        me["full"][slot].complete_tx(1)
                    sim.at(rng.uniform(300, 2500), land)
                    yield ("sleep", rng.uniform(5, 700))

    def relay():
        peer, leader = cta[1], cta[0]
        for it in range(rounds):
            for g in range(len(TS_G_NCH)):
                rev = it * TS_REVS + (g >> 2)
                yield maybe_stall()
                yield ("wait", peer["full"][g & 3], rev & 1, rev)
                sim.at(remote(), lambda b=leader["full"][g & 3]: b.arrive())

    pipe_t = [0.0]

    def mma(dur):                           # the tensor pipe executes the issued blocks in order; returns when this one completes
        pipe_t[0] = max(pipe_t[0], sim.now) + dur
        return pipe_t[0] - sim.now

    def issuer():
        leader = cta[0]
        batch = 0
        ar_n, af_n = [0, 0], [0, 0]

        def full_wait(slot, rev):
            return ("wait", leader["full"][slot], rev & 1, rev)

        def read(it, g, last):
            slot = g & 3
            for me in cta:
                for sub in range(TS_G_NCH[g]):
                    assert me["ring"][slot][sub] == (it, g), f"MMA of group {(it, g)} found {me['ring'][slot][sub]} in ring slot {slot}"
                    me["readers_left"][slot][sub] -= 1
                    assert me["readers_left"][slot][sub] >= 0
            done = mma(256.0 * TS_G_NCH[g])
            if last:
                empty_commit(slot, done)

        def empty_commit(slot, done=None):
            done = mma(0.0) if done is None else done            # tcgen05.commit: after everything issued so far
            for me in cta:
                assert me["readers_left"][slot] == [0, 0]
                sim.at(done + rng.uniform(50, 600), lambda b=me["empty"][slot]: b.arrive())

        for it in range(rounds):
            rev0 = it * TS_REVS
            for L, g0 in enumerate(TS_L_GROUP0):
                kind = TS_L_KIND[L]
                s0, rev = g0 & 3, rev0 + (g0 >> 2)
                nb = 2 if kind == "VIEWS" else 4
                yield full_wait(s0, rev)
                for b in range(nb):
                    t, h = (b, 0) if kind == "VIEWS" else (b >> 1, b & 1)
                    first, last = b == 0, b == nb - 1
                    acc = batch & 1
                    batch += 1
                    if h == 0:
                        yield ("wait", aready[t], ar_n[t] & 1, ar_n[t])
                        ar_n[t] += 1
                    yield ("wait", accfree[acc], (af_n[acc] & 1) ^ 1, af_n[acc] - 1)
                    af_n[acc] += 1
                    yield ("sleep", rng.uniform(20, 120))
                    if kind == "L0":
                        read(it, g0, last)
                        if last:
                            if shipped:
                                yield full_wait(1, rev)
                            empty_commit(1)
                    else:
                        read(it, g0, last)
                        if first:
                            yield full_wait(s0 + 1, rev)
                        read(it, g0 + 1, last)
                        if kind == "L5":
                            if first:
                                yield full_wait(0, rev + 1)
                            read(it, g0 + 2, last)
                            if last:
                                if shipped:
                                    yield full_wait(1, rev + 1)
                                empty_commit(1)
                        elif kind == "VIEWS":
                            if first:
                                yield full_wait(2, rev)
                            read(it, g0 + 2, last)
                            if last:
                                if shipped:
                                    yield full_wait(3, rev)
                                empty_commit(3)
                    done = mma(0.0)
                    for me in cta:                               # commit of the batch -> acc_full of both CTAs
                        sim.at(done + rng.uniform(50, 400), lambda bar=me["accfull"][acc]: bar.arrive())

    def epilogue(rank):
        me = cta[rank]
        lat = remote if rank == 1 else (lambda: rng.uniform(5, 50))
        batch = 0
        full_n, free_n, pend = [0, 0], [0, 0, 0, 0], [False] * 4
        for it in range(rounds):
            for t in range(2):
                yield ("sleep", rng.uniform(300, 1500))          # prologue: encode slot t
                sim.at(lat(), lambda t=t: aready[t].arrive())
            for L, kind in enumerate(TS_L_KIND):
                for t in range(2):
                    for h in ((0,) if kind == "VIEWS" else (0, 1)):
                        acc = batch & 1
                        batch += 1
                        yield ("wait", me["accfull"][acc], full_n[acc] & 1, full_n[acc])
                        full_n[acc] += 1
                        sb = 2 * t + h
                        if train and kind != "VIEWS" and pend[sb]:
                            yield ("wait", me["free"][sb], free_n[sb] & 1, free_n[sb])
                            free_n[sb] += 1
                            pend[sb] = False
                        yield ("sleep", rng.uniform(100, 300))   # tcgen05.ld
                        sim.at(lat(), lambda acc=acc: accfree[acc].arrive())
                        yield ("sleep", rng.uniform(300, 1200))  # bias, ReLU, packs, tcgen05.st
                        if kind != "VIEWS":
                            pend[sb] = train
                            if h == 1:
                                sim.at(lat(), lambda t=t: aready[t].arrive())
                                if train:
                                    me["written"][sb - 1].arrive()
                                    me["written"][sb].arrive()

    def store_lane(rank):
        me = cta[rank]
        k = 0
        for it in range(rounds):
            for L in range(len(TS_L_KIND) - 1):
                for b in range(4):
                    yield ("wait", me["written"][b], (k >> 2) & 1, k >> 2)
                    yield ("sleep", rng.uniform(100, 900))       # the bulk copy of buffer b - 1 has read its source
                    if k > 0:
                        me["free"][(b + 3) & 3].arrive()
                    k += 1
        me["free"][3].arrive()

    for rank in (0, 1):
        for lane in range(2 * TS_SLOTS):
            sim.spawn(f"producer{rank}.{lane}", producer(rank, lane))
        sim.spawn(f"epilogue{rank}", epilogue(rank))
        if train:
            sim.spawn(f"store{rank}", store_lane(rank))
    sim.spawn("relay", relay())
    sim.spawn("issuer", issuer())
    sim.run()
    return sim.now


@pytest.mark.parametrize("train", [False, True])
@pytest.mark.parametrize("seed", range(6))
def test_forward_ts_protocol(seed, train):
    assert simulate_ts(rounds=8, seed=seed, train=train) > 0


def test_forward_ts_protocol_under_heavy_stalls():
    for seed in range(4):
        assert simulate_ts(rounds=6, seed=1000 + seed, stall_prob=0.25) > 0


def test_model_flags_round_twos_first_ring_protocol():
    """What deadlocked one train step in ~1000 on the GPU: with one arriving lane per slot and empty groups released unseen, a
    lane (or the relay) that is descheduled between two consecutive releases of its slot waits for a parity that has already
    come round again.  The model must flag that protocol (deadlock or a wait satisfied by the wrong phase) under the same
    stalls that the shipped protocol survives."""
    failures = 0
    for seed in range(8):
        try:
            simulate_ts(rounds=6, seed=1000 + seed, protocol="r2_first", stall_prob=0.25)
        except AssertionError:
            failures += 1
    assert failures >= 6, failures
