"""CPU replay of the CIPS kernel's mbarrier protocol (cips-3d_b200/csrc/cips_tc.cu) on the tile order the
library really uses (c3d_debug_cips_tile_order, host only).

Agents, each an in-order program like its warp(s): the weight producer, the two MMA issuers, the epilogue.
Barriers are phase counters; a wait on phase n must find the barrier at exactly n completed phases - one more
and the kernel's 1-bit parity wait would alias (the race fixed in r01g) - so the replay asserts
  * no parity aliasing on full[] / acc_ready[] / epi_done[],
  * an MMA reads A-operand chunk kc/2 only in the version written for its layer,
  * the first MMA into accumulator block nc (kc == 0, overwrite) comes after the epilogue drained that block,
  * the epilogue overwrites A-operand chunk j in place only after every MMA of the layer that reads it,
  * no deadlock, under many random interleavings.
MMA completion is modelled as immediate; tcgen05.commit only ever delays an arrive, which the random
scheduler covers by delaying the committing agent.
"""
import ctypes as C
import random

import pytest

L_LAYERS, ITERS = 18, 3


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g
    g.build()
    import cips3d_b200
    return cips3d_b200


def _order(pkg):
    lib = pkg._lib.load()
    full = (C.c_uint16 * 32)()
    inn = (C.c_uint16 * 4)()
    lib.c3d_debug_cips_tile_order.argtypes = [C.POINTER(C.c_uint16), C.POINTER(C.c_uint16)]
    lib.c3d_debug_cips_tile_order.restype = C.c_int
    stages = lib.c3d_debug_cips_tile_order(full, inn)
    assert stages > 0
    dec = lambda e: (e & 15, (e >> 4) & 15, (e >> 8) & 15, e >> 12)
    return [dec(e) for e in full], [dec(e) for e in inn], stages


class Bar:
    def __init__(self, count):
        self.count, self.arr, self.phase = count, 0, 0

    def arrive(self):
        self.arr += 1
        if self.arr == self.count:
            self.arr, self.phase = 0, self.phase + 1


def _replay(full, inn, stages, seed):
    rnd = random.Random(seed)
    layers = [inn] + [full] * (L_LAYERS - 1)
    fullb = [Bar(1) for _ in range(stages)]
    emptyb = [Bar(2) for _ in range(stages)]
    epi_done = [Bar(1) for _ in range(4)]       # the 16 epilogue warps move together here
    ready = [Bar(2) for _ in range(4)]
    xver, drained = [-1] * 4, [0] * 4
    reads = [[0] * 4 for _ in range(ITERS * L_LAYERS + 1)]

    def producer():
        g = 0
        for _ in range(ITERS):
            for l in range(L_LAYERS):
                for _t in layers[l]:
                    st, n = g % stages, g // stages
                    yield lambda st=st, n=n: emptyb[st].phase >= n
                    assert emptyb[st].phase == n, "empty[] parity alias"
                    fullb[st].arrive()
                    g += 1

    def issuer(me):
        g = 0
        for it in range(ITERS):
            for l in range(L_LAYERS):
                ph = it * L_LAYERS + l
                waited = -1
                for kc, nc, need, rdy in layers[l]:
                    st, n = g % stages, g // stages
                    mine = (nc >> 1) == me
                    if mine:
                        for j in range(waited + 1, need + 1):
                            yield lambda j=j, ph=ph: epi_done[j].phase >= ph + 1
                            assert epi_done[j].phase == ph + 1, "epi_done[] parity alias"
                        waited = max(waited, need)
                    yield lambda st=st, n=n: fullb[st].phase >= n + 1
                    assert fullb[st].phase == n + 1, "full[] parity alias"
                    if mine:
                        assert xver[kc // 2] == ph, "MMA reads a stale / future A-operand chunk"
                        if kc == 0:
                            assert drained[nc] == ph, "accumulator block overwritten before it was drained"
                        reads[ph][kc // 2] += 1
                        for j in range(4):
                            if rdy >> j & 1:
                                ready[j].arrive()
                    emptyb[st].arrive()
                    g += 1

    def epilogue():
        for it in range(ITERS):
            for j in range(4):
                xver[j] = drained[j] = it * L_LAYERS        # staging of the input tile
                epi_done[j].arrive()
            for l in range(L_LAYERS):
                ph = it * L_LAYERS + l
                for j in range(4):
                    yield lambda j=j, ph=ph: ready[j].phase >= ph + 1
                    assert ready[j].phase == ph + 1, "acc_ready[] parity alias"
                    assert reads[ph][j] == sum(1 for kc, *_ in layers[l] if kc // 2 == j), \
                        "epilogue overwrites an A-operand chunk that this layer's MMAs still read"
                    xver[j] = drained[j] = ph + 1
                    if l != L_LAYERS - 1:
                        epi_done[j].arrive()

    agents = [producer(), issuer(0), issuer(1), epilogue()]
    cond = [None] * 4
    alive = [True] * 4
    while any(alive):
        progressed = False
        idx = list(range(4))
        rnd.shuffle(idx)
        for i in idx:
            if not alive[i]:
                continue
            for _ in range(rnd.randint(1, 40)):
                if cond[i] is not None and not cond[i]():
                    break
                progressed = True
                try:
                    cond[i] = next(agents[i])
                except StopIteration:
                    alive[i] = False
                    break
        assert progressed, f"deadlock (seed {seed})"


def test_tile_order_is_a_permutation_with_valid_dependencies(pkg):
    full, inn, stages = _order(pkg)
    assert stages >= 2
    assert sorted((kc, nc) for kc, nc, _, _ in full) == [(kc, nc) for kc in range(8) for nc in range(4)]
    assert [(kc, nc) for kc, nc, _, _ in inn] == [(0, nc) for nc in range(4)]
    for order in (full, inn):
        seen = set()
        for kc, nc, need, _ in order:
            if nc not in seen:
                assert kc == 0, "first MMA into an accumulator block must be the overwriting one (kc == 0)"
                seen.add(nc)
            assert need >= nc and need >= kc // 2
        for me in range(2):                      # each issuer commits every acc_ready[j] exactly once per layer
            for j in range(4):
                assert sum(1 for _, nc, _, rdy in order if (nc >> 1) == me and rdy >> j & 1) == 1


@pytest.mark.parametrize("seed", range(12))
def test_protocol_replay(pkg, seed):
    full, inn, stages = _order(pkg)
    _replay(full, inn, stages, seed)


def test_replay_catches_the_r01g_ring_race(pkg):
    """Negative control: with the pre-fix protocol (an issuer neither observes nor releases the tiles it does
    not own) and the pre-fix tile order, some interleaving must trip the full[] alias assertion - otherwise
    the replay above proves nothing."""
    _, inn, stages = _order(pkg)
    old, done = [], set()
    for j in range(4):                              # the r01f staircase
        for kc in range(min(2 * (j + 1), 8)):
            for nc in range(j + 1):
                if (kc, nc) not in done:
                    done.add((kc, nc))
                    old.append((kc, nc, j, 0))
    layers = [inn] + [old] * (L_LAYERS - 1)
    tripped = 0
    for seed in range(40):
        rnd = random.Random(seed)
        fullb = [Bar(1) for _ in range(stages)]
        emptyb = [Bar(1) for _ in range(stages)]
        state = {"alias": False}

        def producer():
            g = 0
            for _ in range(2):
                for l in range(L_LAYERS):
                    for _t in layers[l]:
                        st, n = g % stages, g // stages
                        yield lambda st=st, n=n: emptyb[st].phase >= n
                        fullb[st].arrive()
                        g += 1

        def issuer(me):
            g = 0
            for _ in range(2):
                for l in range(L_LAYERS):
                    for kc, nc, need, _r in layers[l]:
                        st, n = g % stages, g // stages
                        if (nc >> 1) == me:
                            # 1-bit parity wait: passes when an odd number of phases separates us from the barrier
                            yield lambda st=st, n=n: (fullb[st].phase - (n + 1)) % 2 == 0 and fullb[st].phase >= n - 1
                            if fullb[st].phase != n + 1:
                                state["alias"] = True
                                return
                            emptyb[st].arrive()
                        g += 1

        agents = [producer(), issuer(0), issuer(1)]
        cond, alive = [None] * 3, [True] * 3
        for _round in range(20000):
            if state["alias"] or not any(alive):
                break
            progressed = False
            for i in rnd.sample(range(3), 3):
                if not alive[i]:
                    continue
                for _ in range(rnd.randint(1, 60)):
                    if cond[i] is not None and not cond[i]():
                        break
                    progressed = True
                    try:
                        cond[i] = next(agents[i])
                    except StopIteration:
                        alive[i] = False
                        break
            if not progressed:
                break
        tripped += state["alias"]
    assert tripped > 0
