"""Protocol model of the attention kernels' warp-role synchronisation (attention.cu), run on the CPU.

The kernels coordinate a TMA producer, a single-thread MMA issuer (whose tcgen05 work retires IN ORDER and signals
mbarriers through tcgen05.commit) and 8..24 softmax warps through 1-bit-parity mbarriers and named barriers.  A wrong
parity, a missing wait or a buffer reused one tile too early is a hang or silent corruption on the GPU.  This test replays
the exact sequence of waits / arrives / commits of every kernel variant under thousands of random interleavings and random
latencies and checks, at every access, that the buffer holds the tile the accessor expects:

  * K / V stage j is not overwritten before every product that reads it has retired;
  * S_j is complete when a softmax warp loads it, and is not overwritten before all warps of the group have loaded it;
  * P_j is fully written (all warps) before PV_j starts, and not overwritten while PV_{j-PB} may still read it;
  * O is only rescaled while no PV product is in flight, and PV_j starts after every rescale of tile j;
  * nobody passes an mbarrier wait early through parity aliasing; the run terminates (no deadlock).

The model mirrors attention_kernel<.., SB, PB, ..> (incl. the BKV == 64 "lazy pv_done wait" with two p_full barriers) and the
two-tile attention_fa_kernel (two query tiles per CTA, early S release, MMA issue order PV0, S1, PV1, S0, MUFU token).  It knows
nothing about arithmetic — only about who may touch what, when.  (The round-1 ping-pong kernel this file also modelled was
measured slower and removed from the product in round 2; its model went with it.)
"""
import random

import pytest


class MBar(object):
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects in one phase"
        if self.pending == 0:
            self.phase ^= 1
            self.pending = self.count

    def done(self, parity):          # mbarrier.try_wait.parity: has the phase with this parity completed?
        return self.phase != parity


class NamedBar(object):
    def __init__(self, warps):
        self.warps, self.n, self.gen = warps, 0, 0

    def arrive(self):
        self.n += 1
        assert self.n <= self.warps, "named barrier over-subscribed"
        if self.n == self.warps:
            self.n, self.gen = 0, self.gen + 1


class Sim(object):
    """Agents are generators; they yield a predicate (block until true) or None (just a scheduling point)."""

    def __init__(self, seed):
        self.rng = random.Random(seed)
        self.agents = []
        self.pipe = []            # the tensor pipe: FIFO of ("mma", start_fn, end_fn) / ("commit", mbar), retires in order

    def spawn(self, gen):
        self.agents.append([gen, None])

    def tensor_pipe(self):
        while True:
            if not self.pipe:
                if self.issuer_done:
                    return
                yield None
                continue
            kind, a, b = self.pipe[0]
            if kind == "commit":
                self.pipe.pop(0)
                a.arrive()
                continue
            a()                                       # product starts reading its operands
            for _ in range(self.rng.randint(0, 3)):
                yield None
            b()                                       # product retires
            self.pipe.pop(0)

    def run(self, limit=300_000):
        self.issuer_done = False
        steps = 0
        while self.agents:
            ready = [ag for ag in self.agents if ag[1] is None or ag[1]()]
            assert ready, "deadlock: every live role is blocked"
            ag = self.rng.choice(ready)
            try:
                ag[1] = next(ag[0])
            except StopIteration:
                self.agents.remove(ag)
            steps += 1
            assert steps < limit, "livelock (roles spin without progress)"


def delay(sim, hi=3):
    for _ in range(sim.rng.randint(0, hi)):
        yield None


def wait(bar, parity):
    return lambda: bar.done(parity)


# ---------------------------------------------------------------------------------------------------------------------
# attention_kernel<DK, DVP, BKV, KV_STAGES, SB, PB, SW = 2>
# ---------------------------------------------------------------------------------------------------------------------
def simulate_attention_kernel(seed, ntiles, ST, SB, PB, lazy, rescale_prob=0.3, mutate=None):
    """mutate (self-checks of the checker): 'single_p_full' = lazy waits with ONE p_full barrier, 'no_last_wait' = lazy waits
    without the forced pv_done wait on the last tile — the two hazards found while writing the BKV == 64 variant."""
    sim = Sim(seed)
    PF = 2 if (lazy and mutate != "single_p_full") else 1
    NW = 8
    k_full, k_empty = [MBar(1) for _ in range(ST)], [MBar(1) for _ in range(ST)]
    v_full, v_empty = [MBar(1) for _ in range(ST)], [MBar(1) for _ in range(ST)]
    s_full, p_full, pv_done = [MBar(1) for _ in range(2)], [MBar(NW) for _ in range(PF)], MBar(1)
    pair = [NamedBar(2) for _ in range(4)]
    kst, vst = [None] * ST, [None] * ST            # tile held by each K / V stage
    kbusy, vbusy = [0] * ST, [0] * ST              # products currently reading the stage
    S = [dict(tile=None, readers=set()) for _ in range(SB)]
    P = [dict(tile=None, writers=set(), busy=0) for _ in range(PB)]
    O = dict(pv_inflight=0, pv_retired=-1, rescaled=[set() for _ in range(ntiles)])

    def tma():
        for j in range(ntiles):
            st, ph = j % ST, (j // ST) & 1
            yield wait(k_empty[st], ph ^ 1)
            assert kbusy[st] == 0, "K stage overwritten while a product reads it"
            yield from delay(sim)
            kst[st] = j
            k_full[st].arrive()
            yield wait(v_empty[st], ph ^ 1)
            assert vbusy[st] == 0, "V stage overwritten while a product reads it"
            yield from delay(sim)
            vst[st] = j
            v_full[st].arrive()

    def issue_S(j):
        st = j % ST
        yield wait(k_full[st], (j // ST) & 1)
        buf = S[j % SB]

        def start():
            assert kst[st] == j, f"S_{j} reads K stage holding tile {kst[st]}"
            assert buf["tile"] is None or len(buf["readers"]) == NW, f"S buffer overwritten before all warps loaded S_{buf['tile']}"
            kbusy[st] += 1
            buf["tile"], buf["readers"] = None, set()

        def end():
            kbusy[st] -= 1
            buf["tile"] = j
        sim.pipe.append(("mma", start, end))
        sim.pipe.append(("commit", k_empty[st], None))
        sim.pipe.append(("commit", s_full[j % SB], None))

    def mma():
        yield from issue_S(0)
        if SB == 2 and ntiles > 1:
            yield from issue_S(1)
        for j in range(ntiles):
            st = j % ST
            yield wait(p_full[j % PF], (j // PF) & 1)
            if SB == 1 and j + 1 < ntiles:
                yield from issue_S(j + 1)
            yield wait(v_full[st], (j // ST) & 1)
            pb = P[j % PB]

            def start(j=j, st=st, pb=pb):
                assert vst[st] == j, f"PV_{j} reads V stage holding tile {vst[st]}"
                assert pb["tile"] == j and len(pb["writers"]) == NW, f"PV_{j} started before P_{j} was complete"
                assert len(O["rescaled"][j]) == NW, f"PV_{j} started before every warp settled O for tile {j}"
                vbusy[st] += 1
                pb["busy"] += 1
                O["pv_inflight"] += 1

            def end(j=j, st=st, pb=pb):
                vbusy[st] -= 1
                pb["busy"] -= 1
                O["pv_inflight"] -= 1
                O["pv_retired"] = j
            sim.pipe.append(("mma", start, end))
            sim.pipe.append(("commit", v_empty[st], None))
            sim.pipe.append(("commit", pv_done, None))
            if SB == 2 and j + 2 < ntiles:
                yield from issue_S(j + 2)
        sim.issuer_done = True

    def softmax(w):
        quarter = w & 3
        # the rescale decision is identical in the two warps of a quarter (same rows): draw it per (quarter, tile)
        rs = [random.Random(seed * 1000 + quarter * 100 + j).random() < rescale_prob for j in range(ntiles)]
        for j in range(ntiles):
            yield wait(s_full[j % SB], (j // SB) & 1)
            buf = S[j % SB]
            assert buf["tile"] == j, f"warp {w} loaded S buffer holding tile {buf['tile']} instead of {j}"
            buf["readers"].add(w)
            yield from delay(sim)
            g0 = pair[quarter].gen                      # pair_sync(): row-max exchange with the quarter's other warp
            pair[quarter].arrive()
            yield lambda g0=g0: pair[quarter].gen != g0
            rescale = j > 0 and rs[j]
            if PB == 1 and j > 0:
                yield wait(pv_done, (j - 1) & 1)
                assert O["pv_retired"] >= j - 1, f"parity aliasing: warp {w} passed pv_done({j - 1}) early"
            pb = P[j % PB]
            assert pb["busy"] == 0, f"warp {w} writes P_{j} while a PV product still reads the buffer"
            if pb["tile"] != j:
                pb["tile"], pb["writers"] = j, set()
            yield from delay(sim)
            pb["writers"].add(w)
            if j > 0:
                if PB == 2 and ((not lazy) or rescale or (j == ntiles - 1 and mutate != "no_last_wait")):
                    yield wait(pv_done, (j - 1) & 1)
                    assert O["pv_retired"] >= j - 1, f"parity aliasing: warp {w} passed pv_done({j - 1}) early"
                if rescale:
                    assert O["pv_inflight"] == 0 and O["pv_retired"] == j - 1, "O rescaled while a PV product is in flight"
                    yield from delay(sim)
                    assert O["pv_inflight"] == 0, "a PV product started during the rescale"
            O["rescaled"][j].add(w)
            p_full[j % PF].arrive()
        yield wait(pv_done, (ntiles - 1) & 1)
        assert O["pv_retired"] == ntiles - 1, f"epilogue of warp {w} read O before PV_{ntiles - 1} retired"

    sim.spawn(tma())
    sim.spawn(mma())
    sim.spawn(sim.tensor_pipe())
    for w in range(NW):
        sim.spawn(softmax(w))
    sim.run()


VARIANTS = {   # name: (KV_STAGES, SB, PB, lazy pv_done wait + two p_full barriers)
    "bkv128_2cta (default d<=64)": (2, 1, 1, False),
    "bkv128_d80/d160": (2, 2, 2, False),
    "bkv128_d160_1stage": (1, 2, 2, False),
    "bkv64_sb2 (VDB_ATT_BKV=64)": (4, 2, 2, True),
    "bkv64_3cta (short contexts)": (2, 1, 1, False),
}


@pytest.mark.parametrize("name", sorted(VARIANTS))
@pytest.mark.parametrize("ntiles", [1, 2, 3, 5, 8])
def test_attention_kernel_protocol(name, ntiles):
    ST, SB, PB, lazy = VARIANTS[name]
    for seed in range(60):
        simulate_attention_kernel(seed, ntiles, ST, SB, PB, lazy)


@pytest.mark.parametrize("mutation", ["single_p_full", "no_last_wait"])
def test_the_model_catches_the_hazards_it_was_written_for(mutation):
    """Self-check of the checker: each protocol bug found (and fixed) while writing the lazy-wait variant must be detected."""
    caught = 0
    for seed in range(100):
        try:
            simulate_attention_kernel(seed, 6, 4, 2, 2, True, rescale_prob=0.0, mutate=mutation)
        except AssertionError:
            caught += 1
    assert caught >= 20, f"only {caught}/100 interleavings expose the '{mutation}' bug: the model lost its teeth"


# ---------------------------------------------------------------------------------------------------------------------
# attention_fa_kernel<DVP, KV_STAGES, POLY, TOKEN, ONES>: two query tiles (softmax groups) per CTA, one S / P / O buffer each
# ---------------------------------------------------------------------------------------------------------------------
def simulate_fa_kernel(seed, ntiles, ST=3, token=True, rescale_prob=0.3, mutate=None):
    """mutate (self-checks): 'no_s_free' = the next S product is queued without waiting for the group to have loaded the
    previous scores; 'no_pv_wait' = a group writes P_j without waiting for PV_{j-1}; 'k_release_g0' = the K stage is released
    after group 0's product instead of group 1's; 'surplus_pass' = the last token hand-over is not skipped."""
    sim = Sim(seed)
    NW = 4                                                        # warps per softmax group (one per TMEM lane quarter)
    q_full = MBar(1)
    k_full, k_empty = [MBar(1) for _ in range(ST)], [MBar(1) for _ in range(ST)]
    v_full, v_empty = [MBar(1) for _ in range(ST)], [MBar(1) for _ in range(ST)]
    s_full, s_free = [MBar(1) for _ in range(2)], [MBar(NW) for _ in range(2)]
    p_full, pv_done = [MBar(NW) for _ in range(2)], [MBar(1) for _ in range(2)]
    tok = [NamedBar(2 * NW) for _ in range(2)]                    # bar 1 + g: group g syncs (4 warps), the other group arrives (4 warps)
    q_loaded = [False]
    kst, vst = [None] * ST, [None] * ST
    kbusy, vbusy = [0] * ST, [0] * ST
    S = [dict(tile=None, readers=set(), busy=0) for _ in range(2)]
    P = [dict(tile=None, writers=set(), busy=0) for _ in range(2)]
    O = [dict(pv_inflight=0, pv_retired=-1, settled=[set() for _ in range(ntiles)]) for _ in range(2)]
    in_exp = [0, 0]                                               # warps of each group inside their exp2 phase

    def tma():
        yield from delay(sim)
        q_loaded[0] = True
        q_full.arrive()
        for j in range(ntiles):
            st, ph = j % ST, (j // ST) & 1
            yield wait(k_empty[st], ph ^ 1)
            assert kbusy[st] == 0, "K stage overwritten while a product reads it"
            yield from delay(sim)
            kst[st] = j
            k_full[st].arrive()
            yield wait(v_empty[st], ph ^ 1)
            assert vbusy[st] == 0, "V stage overwritten while a product reads it"
            yield from delay(sim)
            vst[st] = j
            v_full[st].arrive()

    def issue_S(g, j):
        st = j % ST
        yield wait(k_full[st], (j // ST) & 1)
        buf = S[g]

        def start():
            assert q_loaded[0], "S product before Q arrived"
            assert kst[st] == j, f"S_{g}({j}) reads K stage holding tile {kst[st]}"
            assert buf["tile"] is None or len(buf["readers"]) == NW, \
                f"S buffer of group {g} overwritten before all its warps loaded S({buf['tile']})"
            kbusy[st] += 1
            buf["tile"], buf["readers"] = None, set()

        def end():
            kbusy[st] -= 1
            buf["tile"] = j
        sim.pipe.append(("mma", start, end))
        if g == (0 if mutate == "k_release_g0" else 1):
            sim.pipe.append(("commit", k_empty[st], None))
        sim.pipe.append(("commit", s_full[g], None))

    def issue_PV(g, j):
        st = j % ST
        yield wait(p_full[g], j & 1)
        yield wait(v_full[st], (j // ST) & 1)
        pb, og = P[g], O[g]

        def start():
            assert vst[st] == j, f"PV_{g}({j}) reads V stage holding tile {vst[st]}"
            assert pb["tile"] == j and len(pb["writers"]) == NW, f"PV_{g}({j}) started before P was complete"
            assert len(og["settled"][j]) == NW, f"PV_{g}({j}) started before every warp settled O"
            vbusy[st] += 1
            pb["busy"] += 1
            og["pv_inflight"] += 1

        def end():
            vbusy[st] -= 1
            pb["busy"] -= 1
            og["pv_inflight"] -= 1
            og["pv_retired"] = j
        sim.pipe.append(("mma", start, end))
        if g == 1:
            sim.pipe.append(("commit", v_empty[st], None))
        sim.pipe.append(("commit", pv_done[g], None))

    def next_S(g, j):
        if mutate != "no_s_free":
            yield wait(s_free[g], (j - 1) & 1)
        yield from issue_S(g, j)

    def mma():
        yield wait(q_full, 0)
        yield from issue_S(0, 0)
        yield from issue_S(1, 0)
        if ntiles > 1:
            yield from next_S(0, 1)
        for j in range(ntiles):
            yield from issue_PV(0, j)
            if j + 1 < ntiles:
                yield from next_S(1, j + 1)
            yield from issue_PV(1, j)
            if j + 2 < ntiles:
                yield from next_S(0, j + 2)
        sim.issuer_done = True

    def softmax(g, w):
        rs = [random.Random(seed * 1000 + g * 500 + w * 50 + j).random() < rescale_prob for j in range(ntiles)]
        if token and g == 1 and True:
            tok[0].arrive()                                         # prime the ring: group 0 goes first
        for j in range(ntiles):
            yield wait(s_full[g], j & 1)
            buf = S[g]
            assert buf["tile"] == j, f"group {g} warp {w} loaded S buffer holding tile {buf['tile']} instead of {j}"
            buf["readers"].add(w)
            yield from delay(sim)
            s_free[g].arrive()                                      # the scores live in registers from here on
            yield from delay(sim)                                   # row max, rescale decision
            if j > 0:
                if mutate != "no_pv_wait":
                    yield wait(pv_done[g], (j - 1) & 1)
                    assert O[g]["pv_retired"] >= j - 1, f"parity aliasing: group {g} warp {w} passed pv_done({j - 1}) early"
                if rs[j]:
                    assert O[g]["pv_inflight"] == 0, "O rescaled while a PV product of the group is in flight"
                    yield from delay(sim)
            O[g]["settled"][j].add(w)
            if token:
                g0 = tok[g].gen
                tok[g].arrive()                                     # bar.sync 1 + g, 256
                yield lambda g0=g0: tok[g].gen != g0
                in_exp[g] += 1
                assert in_exp[g ^ 1] == 0, "both groups inside their exp2 phase (the token is not exclusive)"
            pb = P[g]
            assert pb["busy"] == 0, f"group {g} warp {w} writes P({j}) while PV({pb['tile']}) still reads the buffer"
            if pb["tile"] != j:
                pb["tile"], pb["writers"] = j, set()
            yield from delay(sim)
            pb["writers"].add(w)
            if token:
                in_exp[g] -= 1
                if not (j == ntiles - 1 and g == 1) or mutate == "surplus_pass":
                    tok[g ^ 1].arrive()                             # bar.arrive 1 + (g ^ 1), 256
            p_full[g].arrive()
        yield wait(pv_done[g], (ntiles - 1) & 1)
        assert O[g]["pv_retired"] == ntiles - 1, f"epilogue of group {g} warp {w} read O before the last PV retired"

    sim.spawn(tma())
    sim.spawn(mma())
    sim.spawn(sim.tensor_pipe())
    for g in range(2):
        for w in range(NW):
            sim.spawn(softmax(g, w))
    sim.run()
    if token:
        assert all(t.n == 0 for t in tok), "a token hand-over was left dangling at kernel exit"


@pytest.mark.parametrize("token", [True, False])
@pytest.mark.parametrize("ntiles", [1, 2, 3, 4, 7, 32])
def test_two_tile_kernel_protocol(ntiles, token):
    for seed in range(40 if ntiles < 32 else 6):
        simulate_fa_kernel(seed, ntiles, ST=3, token=token)
        simulate_fa_kernel(seed + 1000, ntiles, ST=2, token=token)


@pytest.mark.parametrize("mutation", ["no_s_free", "no_pv_wait", "k_release_g0", "surplus_pass"])
def test_two_tile_model_self_check(mutation):
    """each mutation removes one wait / moves one release of the real kernel: the model must notice"""
    caught = 0
    for seed in range(100):
        try:
            simulate_fa_kernel(seed, 6, ST=2, rescale_prob=0.5, mutate=mutation)
        except AssertionError:
            caught += 1
    assert caught >= 10, f"only {caught}/100 interleavings expose the '{mutation}' bug: the model lost its teeth"
