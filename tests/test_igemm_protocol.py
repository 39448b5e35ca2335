"""Protocol model of igemm_kernel's three roles (igemm.cu): TMA producer -> smem ring (full / empty mbarriers) -> MMA issuer
-> two TMEM accumulator stages (tmem_full / tmem_empty) -> eight epilogue warps, for a persistent CTA walking several
output tiles.  Same simulator as tests/test_attention_protocol.py (random interleavings, random latencies, in-order tensor
pipe); checks that a shared-memory stage is never refilled while an MMA still reads it, that the MMA of tile i+2 never
overwrites the accumulator stage before all eight epilogue warps have drained tile i, and that every epilogue warp reads
exactly the tile it expects (1-bit parity aliasing would show up here)."""
import pytest

from test_attention_protocol import MBar, Sim, delay, wait


def simulate_igemm(seed, ntiles, kblocks, STAGES, EW=8, mutate=None):
    sim = Sim(seed)
    full, empty = [MBar(1) for _ in range(STAGES)], [MBar(1) for _ in range(STAGES)]
    tmem_full, tmem_empty = [MBar(1) for _ in range(2)], [MBar(EW) for _ in range(2)]
    stage_data, stage_busy = [None] * STAGES, [0] * STAGES          # (tile, kb) held by a stage; MMAs reading it
    acc = [dict(tile=None, kb_done=0, readers=set()) for _ in range(2)]

    def producer():
        stage, phase = 0, 0
        for t in range(ntiles):
            for kb in range(kblocks):
                yield wait(empty[stage], phase ^ 1)
                assert stage_busy[stage] == 0, "smem stage refilled while an MMA still reads it"
                yield from delay(sim)
                stage_data[stage] = (t, kb)
                full[stage].arrive()                                   # TMA complete_tx
                stage += 1
                if stage == STAGES:
                    stage, phase = 0, phase ^ 1

    def mma():
        stage, phase = 0, 0
        for it in range(ntiles):
            a, aphase = it & 1, (it >> 1) & 1
            if mutate != "no_tmem_empty_wait":
                yield wait(tmem_empty[a], aphase ^ 1)
            for kb in range(kblocks):
                yield wait(full[stage], phase)

                def start(it=it, kb=kb, stage=stage, a=a):
                    assert stage_data[stage] == (it, kb), f"MMA of tile {it} k-block {kb} reads a stage holding {stage_data[stage]}"
                    if kb == 0:
                        prev = acc[a]
                        assert prev["tile"] is None or len(prev["readers"]) == EW, \
                            f"accumulator stage {a} overwritten before the epilogue drained tile {prev['tile']}"
                        acc[a] = dict(tile=it, kb_done=0, readers=set())
                    stage_busy[stage] += 1

                def end(stage=stage, a=a):
                    stage_busy[stage] -= 1
                    acc[a]["kb_done"] += 1
                sim.pipe.append(("mma", start, end))
                sim.pipe.append(("commit", empty[stage], None))
                stage += 1
                if stage == STAGES:
                    stage, phase = 0, phase ^ 1
            sim.pipe.append(("commit", tmem_full[a], None))
        sim.issuer_done = True

    def epilogue(w):
        for it in range(ntiles):
            a, aphase = it & 1, (it >> 1) & 1
            yield wait(tmem_full[a], aphase)
            assert acc[a]["tile"] == it and acc[a]["kb_done"] == kblocks, \
                f"epilogue warp {w} read accumulator stage {a} holding tile {acc[a]['tile']} ({acc[a]['kb_done']} k-blocks) instead of {it}"
            yield from delay(sim, 4)
            acc[a]["readers"].add(w)
            tmem_empty[a].arrive()

    sim.spawn(producer())
    sim.spawn(mma())
    sim.spawn(sim.tensor_pipe())
    for w in range(EW):
        sim.spawn(epilogue(w))
    sim.run()


@pytest.mark.parametrize("STAGES", [4, 5, 6, 8])
@pytest.mark.parametrize("ntiles,kblocks", [(1, 1), (1, 9), (3, 5), (4, 2), (7, 45), (5, 1)])
def test_igemm_roles_protocol(STAGES, ntiles, kblocks):
    for seed in range(25):
        simulate_igemm(seed, ntiles, kblocks, STAGES)


def test_model_catches_a_missing_accumulator_handshake():
    caught = 0
    for seed in range(60):
        try:
            simulate_igemm(seed, 6, 2, 4, mutate="no_tmem_empty_wait")
        except AssertionError:
            caught += 1
    assert caught >= 20, f"only {caught}/60 interleavings expose the missing tmem_empty wait"
