"""Randomised interleaving check of the barrier protocol of tcconv_pair_kernel (ovc_tcconv_pair.cuh).

Every role of both CTAs is a generator that yields before each barrier operation; a scheduler picks runnable roles at
random.  mbarrier model: arrival count, pending count, phase bit; wait(parity) passes when the phase with that parity has
completed (phase bit != parity), as mbarrier.try_wait.parity does.  Detects: deadlock, an arrival that would complete a
phase a waiter has not yet consumed (phase overrun), a weight slot / activation buffer overwritten before the MMAs that
read it were committed, and MMAs issued on data that has not landed in BOTH CTAs.
    python tools/pair_protocol_sim.py [trials]
"""
import random
import sys

NABUF, SLOTS, NISS = 2, 3, 2          # small ring so that wrap-around happens often
NK8, K = 5, 2                         # 5 channel chunks x 2 taps = 10 weight slots per tile


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects"
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase ^ 1

    def passed(self, parity):
        return self.phase != parity


class Cta:
    def __init__(self):
        self.a_full = [Bar(1) for _ in range(NABUF)]        # 128 producer arrivals modelled as one
        self.a_empty = [Bar(NISS) for _ in range(NABUF)]
        self.pa_full = [Bar(1) for _ in range(NABUF)]
        self.b_full = [Bar(1) for _ in range(SLOTS)]
        self.b_empty = [Bar(NISS) for _ in range(SLOTS)]
        self.pb_full = [Bar(1) for _ in range(SLOTS)]
        self.acc_full = Bar(NISS)
        self.slot_data = [None] * SLOTS                     # which weight step the slot holds
        self.buf_data = [None] * NABUF                      # which channel chunk the A buffer holds
        self.slot_readers = [0] * SLOTS                     # issued-but-uncommitted MMA groups reading the slot
        self.buf_readers = [0] * NABUF


def wait(bar, parity):
    while not bar.passed(parity):
        yield "blocked"
    yield "ok"


def tma(c):
    slot, phase = 0, 1
    for it in range(NK8 * K):
        yield from wait(c.b_empty[slot], phase)
        assert c.slot_readers[slot] == 0, "weight slot overwritten while MMAs still read it"
        c.slot_data[slot] = it
        c.b_full[slot].arrive()
        yield "ok"
        slot += 1
        if slot == SLOTS:
            slot, phase = 0, phase ^ 1


def producer(c):
    for q in range(NK8):
        buf = q % NABUF
        yield from wait(c.a_empty[buf], ((q // NABUF) & 1) ^ 1)
        assert c.buf_readers[buf] == 0, "A buffer overwritten while MMAs still read it"
        c.buf_data[buf] = q
        c.a_full[buf].arrive()
        yield "ok"
    yield from wait(c.acc_full, 0)                          # epilogue


def forward_b(peer, leader):
    slot, phase = 0, 0
    for it in range(NK8 * K):
        yield from wait(peer.b_full[slot], phase)
        leader.pb_full[slot].arrive()
        yield "ok"
        slot += 1
        if slot == SLOTS:
            slot, phase = 0, phase ^ 1


def forward_a(peer, leader):
    for q in range(NK8):
        buf = q % NABUF
        yield from wait(peer.a_full[buf], (q // NABUF) & 1)
        leader.pa_full[buf].arrive()
        yield "ok"


def issuer(leader, peer):
    slot, bphase = 0, 0
    it = 0
    for q in range(NK8):
        buf = q % NABUF
        aph = (q // NABUF) & 1
        yield from wait(leader.a_full[buf], aph)
        yield from wait(leader.pa_full[buf], aph)
        assert leader.buf_data[buf] == q and peer.buf_data[buf] == q, "MMA on a stale / missing A chunk"
        for tap in range(K):
            yield from wait(leader.b_full[slot], bphase)
            yield from wait(leader.pb_full[slot], bphase)
            assert leader.slot_data[slot] == it and peer.slot_data[slot] == it, "MMA on a stale / missing weight slot"
            for c in (leader, peer):
                c.slot_readers[slot] += 1
                c.buf_readers[buf] += 1
            yield "ok"                                      # the MMAs run asynchronously; the commit retires them
            for c in (leader, peer):
                c.slot_readers[slot] -= 1
                c.b_empty[slot].arrive()                    # tcgen05.commit ... multicast: both CTAs
            yield "ok"
            it += 1
            slot += 1
            if slot == SLOTS:
                slot, bphase = 0, bphase ^ 1
        for c in (leader, peer):
            c.buf_readers[buf] -= K
            c.a_empty[buf].arrive()
        yield "ok"
    for c in (leader, peer):
        c.acc_full.arrive()
    yield "ok"


def run(seed):
    rng = random.Random(seed)
    L, P = Cta(), Cta()
    roles = {
        "L.tma": tma(L), "P.tma": tma(P), "L.prod": producer(L), "P.prod": producer(P),
        "L.iss0": issuer(L, P), "L.iss1": issuer(L, P), "P.fwd_b": forward_b(P, L), "P.fwd_a": forward_a(P, L),
    }
    blocked = set()
    while roles:
        runnable = [n for n in roles if n not in blocked] or None
        if runnable is None:
            raise RuntimeError(f"deadlock: {sorted(roles)} all blocked")
        name = rng.choice(runnable)
        try:
            state = next(roles[name])
        except StopIteration:
            del roles[name]
            blocked.clear()
            continue
        if state == "blocked":
            blocked.add(name)
        else:
            blocked.clear()


if __name__ == "__main__":
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    for s in range(trials):
        run(s)
    print(f"pair protocol: {trials} random interleavings, no deadlock / overrun / stale operand")
