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

The model mirrors attention_kernel<.., SB, PB, ..> (incl. the BKV == 64 "lazy pv_done wait" with two p_full barriers) and
attention_pp_kernel<.., G, ..> (exp2 token ring).  It knows nothing about arithmetic — only about who may touch what, when.
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
# attention_pp_kernel<DVP, G, KV_STAGES>: G softmax groups, single S / P buffer each, exp2 token ring
# ---------------------------------------------------------------------------------------------------------------------
def simulate_pp_kernel(seed, ntiles, G, ST, rescale_prob=0.3, mutate=None):
    """mutate: 'no_prime' = the token ring is never primed (must deadlock), 'surplus_pass' = the last group hands the token on
    after its last tile (must leave a dangling arrival)."""
    sim = Sim(seed)
    NW = 8
    k_full, k_empty = [MBar(1) for _ in range(ST)], [MBar(1) for _ in range(ST)]
    v_full, v_empty = [MBar(1) for _ in range(ST)], [MBar(1) for _ in range(ST)]
    s_full, p_full, pv_done = [MBar(1) for _ in range(G)], [MBar(NW) for _ in range(G)], [MBar(1) for _ in range(G)]
    pair = [[NamedBar(2) for _ in range(4)] for _ in range(G)]
    token = [NamedBar(2 * NW) for _ in range(G)]      # 8 waiting warps + 8 arriving warps
    kst, vst, kbusy, vbusy = [None] * ST, [None] * ST, [0] * ST, [0] * ST
    S = [dict(tile=None, readers=set()) for _ in range(G)]
    P = [dict(tile=None, writers=set(), busy=0) for _ in range(G)]
    O = [dict(pv_inflight=0, pv_retired=-1, settled=[set() for _ in range(ntiles)]) for _ in range(G)]
    in_exp = set()                                     # groups currently holding the exp2 token (must never be two)

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

    def issue_S(g, j):
        st = j % ST
        yield wait(k_full[st], (j // ST) & 1)
        buf = S[g]

        def start():
            assert kst[st] == j, f"S_{g}({j}) reads K stage holding tile {kst[st]}"
            assert buf["tile"] is None or len(buf["readers"]) == NW, "S buffer overwritten before all warps of the group loaded it"
            kbusy[st] += 1
            buf["tile"], buf["readers"] = None, set()

        def end():
            kbusy[st] -= 1
            buf["tile"] = j
        sim.pipe.append(("mma", start, end))
        if g == G - 1:
            sim.pipe.append(("commit", k_empty[st], None))
        sim.pipe.append(("commit", s_full[g], None))

    def mma():
        for g in range(G):
            yield from issue_S(g, 0)
        for j in range(ntiles):
            st = j % ST
            for g in range(G):
                yield wait(p_full[g], j & 1)
                if j + 1 < ntiles:
                    yield from issue_S(g, j + 1)
                yield wait(v_full[st], (j // ST) & 1)

                def start(j=j, st=st, g=g):
                    assert vst[st] == j, f"PV_{g}({j}) reads V stage holding tile {vst[st]}"
                    assert P[g]["tile"] == j and len(P[g]["writers"]) == NW, f"PV_{g}({j}) started before P was complete"
                    assert len(O[g]["settled"][j]) == NW
                    vbusy[st] += 1
                    P[g]["busy"] += 1
                    O[g]["pv_inflight"] += 1

                def end(j=j, st=st, g=g):
                    vbusy[st] -= 1
                    P[g]["busy"] -= 1
                    O[g]["pv_inflight"] -= 1
                    O[g]["pv_retired"] = j
                sim.pipe.append(("mma", start, end))
                if g == G - 1:
                    sim.pipe.append(("commit", v_empty[st], None))
                sim.pipe.append(("commit", pv_done[g], None))
        sim.issuer_done = True

    def softmax(g, w):
        quarter = w & 3
        rs = [random.Random(seed * 1000 + g * 7919 + quarter * 100 + j).random() < rescale_prob for j in range(ntiles)]
        if g == G - 1 and mutate != "no_prime":
            token[0].arrive()                           # prime the ring
        for j in range(ntiles):
            yield wait(s_full[g], j & 1)
            assert S[g]["tile"] == j, f"group {g} warp {w} loaded S holding tile {S[g]['tile']} instead of {j}"
            S[g]["readers"].add(w)
            yield from delay(sim)
            g0 = pair[g][quarter].gen
            pair[g][quarter].arrive()
            yield lambda g0=g0: pair[g][quarter].gen != g0
            rescale = j > 0 and rs[j]
            if j > 0:
                yield wait(pv_done[g], (j - 1) & 1)
                assert O[g]["pv_retired"] >= j - 1, "parity aliasing on pv_done"
            t0 = token[g].gen                           # token_wait(): bar.sync 13 + g, 512
            token[g].arrive()
            yield lambda t0=t0: token[g].gen != t0
            in_exp.add((g, w))
            assert all(x[0] == g for x in in_exp), f"two groups in their exp2 phase at once: {sorted(in_exp)}"
            assert P[g]["busy"] == 0, "P written while PV still reads it"
            if P[g]["tile"] != j:
                P[g]["tile"], P[g]["writers"] = j, set()
            yield from delay(sim)
            P[g]["writers"].add(w)
            in_exp.discard((g, w))
            if mutate == "surplus_pass" or not (j == ntiles - 1 and g == G - 1):
                token[(g + 1) % G].arrive()             # token_pass(): bar.arrive
            if rescale:
                assert O[g]["pv_inflight"] == 0 and O[g]["pv_retired"] == j - 1
                yield from delay(sim)
                assert O[g]["pv_inflight"] == 0
            O[g]["settled"][j].add(w)
            p_full[g].arrive()
        yield wait(pv_done[g], (ntiles - 1) & 1)
        assert O[g]["pv_retired"] == ntiles - 1

    sim.spawn(tma())
    sim.spawn(mma())
    sim.spawn(sim.tensor_pipe())
    for g in range(G):
        for w in range(NW):
            sim.spawn(softmax(g, w))
    sim.run()
    assert all(t.n == 0 for t in token), "a token hand-over was left dangling at kernel exit"


@pytest.mark.parametrize("G", [2, 3])
@pytest.mark.parametrize("ntiles", [1, 2, 3, 6])
def test_pingpong_kernel_protocol(G, ntiles):
    for seed in range(40):
        simulate_pp_kernel(seed, ntiles, G, ST=4)
        simulate_pp_kernel(seed + 1000, ntiles, G, ST=2)


@pytest.mark.parametrize("mutation", ["no_prime", "surplus_pass"])
def test_pingpong_model_self_check(mutation):
    with pytest.raises(AssertionError):
        simulate_pp_kernel(0, 3, 3, ST=4, mutate=mutation)
