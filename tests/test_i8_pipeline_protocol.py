"""CPU suite: randomised discrete-event simulation of the mbarrier protocol of the INT8 digit-plane GLM kernel
(cmdstan_b200/csrc/glm_i8.cu), which itself needs tcgen05 / tensor memory and runs on hardware only.  Roles as in the kernel: a
producer thread (q planes + scales of a chain tile, X planes through a ring of NB slots), one MMA-issuing thread (first product of
tile it + 1 before the second product of tile it; completions committed to mbarriers in issue order by an in-order tensor pipe),
NEW epilogue warps (read Eta, release it early or late, write the dm planes, flush G at the end of a chain tile).  Every wait is
the kernel's parity wait.  Checked under random speeds and latencies: no deadlock; no ring slot, q buffer, scale buffer, dm buffer
or accumulator is overwritten while something still reads it; every read sees the data of the iteration it belongs to.
The r02u bug -- epilogue warps waiting on q_full, which the MMA thread can advance twice while a warp is still in its flush -- is
kept as a mutant that the model must catch."""
import random

import pytest

NEW = 4     # epilogue warps (16 in the kernel: the protocol only counts arrivals)
NS = 3      # digit planes (6 in the kernel)


class MBar:
    def __init__(self, count=1):
        self.count, self.arrived, self.phase = count, 0, 0

    def arrive(self):
        self.arrived += 1
        assert self.arrived <= self.count, "more arrivals than the barrier expects in one phase"
        if self.arrived == self.count:
            self.arrived = 0
            self.phase += 1

    def test(self, parity):   # mbarrier.try_wait.parity: true once the phase with this parity has completed
        return (self.phase & 1) != (parity & 1)


def simulate(n_rt, n_my, nb, seed, early_release=True, wait_q_full_in_epilogue=False, skip_dm_empty_wait=False, limit=400_000):
    rng = random.Random(seed)
    n_it = n_rt * n_my
    x_full, x_empty = [MBar() for _ in range(nb)], [MBar() for _ in range(nb)]
    q_full, q_empty, a_full, dm_empty, g_full = MBar(2), MBar(), MBar(), MBar(), MBar()   # q_full: digits + scales, two copies
    a_empty, dm_full, g_empty = MBar(NEW), MBar(NEW), MBar(NEW)
    now = [0]
    # ---- state the hazards are checked on
    slot_tag = [None] * nb            # (it, plane) held by a ring slot
    slot_readers = [0] * nb           # MMAs issued but not yet executed that read the slot
    slot_inflight = [False] * nb
    q_tag, q_readers, q_inflight = [None], [0], [False]
    sc_tag = [None, None]             # scale buffers (j & 1)
    eta_tag, eta_read = [None], [set()]          # iteration whose Eta sits in the accumulator; warps that have read it
    dm_tag, dm_written, dm_readers = [None], [set()], [0]
    g_tag, g_done, g_read = [None], [0], [set(range(NEW))]   # chain tile in the G accumulator, second products executed, warps that read
    pending = []                      # in-flight TMA copies: (ready time, fn)
    tensor = []                       # in-order tensor pipe: (duration, fn at execution end)
    tensor_busy = [0]

    def tma(latency, fn):
        pending.append((now[0] + latency, fn))

    def enqueue(duration, fn):
        tensor.append([duration, fn])

    # ---- producer
    def producer():
        slot, npass = 0, 0
        for j in range(n_my):
            if j > 0:
                while not q_empty.test(j - 1):
                    yield
            assert q_readers[0] == 0 and not q_inflight[0], "q planes overwritten while a first product still reads them"
            q_inflight[0] = True

            def q_done(j=j):
                q_tag[0] = j
                q_inflight[0] = False
                q_full.arrive()

            def sc_done(j=j):
                sc_tag[j & 1] = j
                q_full.arrive()
            tma(rng.randint(1, 30), q_done)
            tma(rng.randint(1, 30), sc_done)
            for rt in range(n_rt):
                for i in range(NS):
                    if npass > 0:
                        while not x_empty[slot].test(npass - 1):
                            yield
                    assert slot_readers[slot] == 0 and not slot_inflight[slot], f"ring slot {slot} overwritten while an MMA still reads it"
                    slot_inflight[slot] = True

                    def x_done(slot=slot, tag=(j * n_rt + rt, i)):
                        slot_tag[slot] = tag
                        slot_inflight[slot] = False
                        x_full[slot].arrive()
                    tma(rng.randint(1, 60), x_done)
                    slot += 1
                    if slot == nb:
                        slot, npass = 0, npass + 1
                    if rng.random() < 0.3:
                        yield

    # ---- MMA issuer
    def mma():
        st = dict(a_slot=0, a_par=0, a_it=0, a_rt=0, a_j=0, b_slot=0, b_it=0, b_rt=0, b_j=0)

        def issue_a():
            it, j = st["a_it"], st["a_j"]
            if it > 0:
                while not a_empty.test(it - 1):
                    yield
            if st["a_rt"] == 0:
                while not q_full.test(j):
                    yield
            for i in range(NS):
                s = st["a_slot"]
                while not x_full[s].test(st["a_par"]):
                    yield
                slot_readers[s] += 1
                q_readers[0] += 1

                def run(s=s, it=it, i=i, j=j):
                    assert slot_tag[s] == (it, i) and not slot_inflight[s], f"first product of tile {it} plane {i} reads slot {s} holding {slot_tag[s]}"
                    assert q_tag[0] == j and not q_inflight[0], f"first product of chain tile {j} reads the q planes of {q_tag[0]}"
                    if i == 0:   # accumulate = 0: the Eta accumulator is overwritten
                        assert it == 0 or eta_read[0] == set(range(NEW)), f"Eta of tile {it - 1} overwritten before warps {set(range(NEW)) - eta_read[0]} read it"
                        eta_tag[0], eta_read[0] = None, set()
                    if i == NS - 1:
                        eta_tag[0] = it
                    slot_readers[s] -= 1
                    q_readers[0] -= 1
                enqueue(rng.randint(1, 8), run)
                st["a_slot"] += 1
                if st["a_slot"] == nb:
                    st["a_slot"], st["a_par"] = 0, st["a_par"] ^ 1
                if rng.random() < 0.3:
                    yield
            enqueue(0, a_full.arrive)
            st["a_it"] += 1
            st["a_rt"] += 1
            if st["a_rt"] == n_rt:
                enqueue(0, q_empty.arrive)
                st["a_rt"], st["a_j"] = 0, st["a_j"] + 1

        def issue_b():
            it, j = st["b_it"], st["b_j"]
            while not dm_full.test(it):
                yield
            if st["b_rt"] == 0 and j > 0:
                while not g_empty.test(j - 1):
                    yield
            first = st["b_rt"] == 0
            for i in range(NS):
                s = st["b_slot"]
                slot_readers[s] += 1
                dm_readers[0] += 1

                def run(s=s, it=it, i=i, j=j, first=first):
                    assert slot_tag[s] == (it, i) and not slot_inflight[s], f"second product of tile {it} plane {i} reads slot {s} holding {slot_tag[s]}"
                    assert dm_tag[0] == it and dm_written[0] == set(range(NEW)), f"second product of tile {it} reads dm planes of {dm_tag[0]} written by {dm_written[0]}"
                    if first and i == 0:   # accumulate = 0: the G accumulator is overwritten
                        assert g_read[0] == set(range(NEW)), f"G of chain tile {g_tag[0]} overwritten before warps {set(range(NEW)) - g_read[0]} read it"
                        g_tag[0], g_done[0], g_read[0] = j, 0, set()
                    if i == NS - 1:
                        g_done[0] += 1
                    slot_readers[s] -= 1
                    dm_readers[0] -= 1
                enqueue(rng.randint(1, 8), run)
                enqueue(0, x_empty[s].arrive)
                st["b_slot"] = (s + 1) % nb
            enqueue(0, dm_empty.arrive)
            st["b_it"] += 1
            st["b_rt"] += 1
            if st["b_rt"] == n_rt:
                enqueue(0, g_full.arrive)
                st["b_rt"], st["b_j"] = 0, st["b_j"] + 1

        yield from issue_a()
        for it in range(n_it):
            if it + 1 < n_it:
                yield from issue_a()
            yield from issue_b()

    # ---- epilogue warps
    bar_sync = [0, 0]   # arrivals, generation

    def epilogue(w):
        it = 0
        for j in range(n_my):
            if wait_q_full_in_epilogue:
                while not q_full.test(j):
                    yield
            for rt in range(n_rt):
                while not a_full.test(it):
                    yield
                assert eta_tag[0] == it, f"warp {w} reads Eta of tile {eta_tag[0]} for tile {it}"
                assert sc_tag[j & 1] == j, f"warp {w} reads the scales of chain tile {sc_tag[j & 1]} for chain tile {j}"
                eta_read[0].add(w)
                if early_release:
                    a_empty.arrive()
                for _ in range(rng.randint(0, 10)):   # recombination + link
                    yield
                assert sc_tag[j & 1] == j
                if not early_release:
                    a_empty.arrive()
                if it > 0 and not skip_dm_empty_wait:
                    while not dm_empty.test(it - 1):
                        yield
                assert dm_readers[0] == 0, f"warp {w} overwrites the dm planes while a second product still reads them"
                if dm_tag[0] != it:
                    dm_tag[0], dm_written[0] = it, set()
                dm_written[0].add(w)
                dm_full.arrive()
                it += 1
                if rng.random() < 0.15:               # a warp that falls behind
                    for _ in range(rng.randint(0, 40)):
                        yield
            while not g_full.test(j):
                yield
            assert g_tag[0] == j and g_done[0] == n_rt, f"warp {w} reads G of chain tile {g_tag[0]} after {g_done[0]} of {n_rt} products"
            g_read[0].add(w)
            g_empty.arrive()
            for _ in range(rng.randint(0, 80 if rng.random() < 0.3 else 8)):   # flush: logs, shuffles, stores -- sometimes slow
                yield
            gen = bar_sync[1]                         # bar.sync 1, NEW (twice in the kernel; once suffices for the model)
            bar_sync[0] += 1
            if bar_sync[0] == NEW:
                bar_sync[0], bar_sync[1] = 0, gen + 1
            while bar_sync[1] == gen:
                yield

    roles = {"producer": producer(), "mma": mma(), **{f"epi{w}": epilogue(w) for w in range(NEW)}}
    alive = set(roles)
    while alive:
        now[0] += 1
        if now[0] > limit:
            raise TimeoutError(f"deadlock: {sorted(alive)} still waiting")
        for p in [p for p in pending if p[0] <= now[0]]:
            pending.remove(p)
            p[1]()
        if tensor:                                    # in-order tensor pipe
            if tensor_busy[0] >= tensor[0][0]:
                tensor.pop(0)[1]()
                tensor_busy[0] = 0
            else:
                tensor_busy[0] += 1
        for name in rng.sample(sorted(alive), len(alive)):
            if rng.random() < 0.75:
                try:
                    next(roles[name])
                except StopIteration:
                    alive.discard(name)
    while tensor:                                     # drain
        tensor.pop(0)[1]()
    assert g_done[0] == n_rt and g_tag[0] == n_my - 1


@pytest.mark.parametrize("early_release", [True, False])
def test_protocol_random_schedules(early_release):
    for seed in range(60):
        rng = random.Random(7000 + seed)
        simulate(n_rt=rng.choice([1, 1, 2, 3, 7]), n_my=rng.choice([1, 2, 3, 5]), nb=2 * NS + rng.choice([0, 0, 1, NS]), seed=seed, early_release=early_release)


def test_the_model_catches_the_q_full_aliasing_deadlock():
    """r02u: with one row tile per chain tile the MMA thread runs two chain tiles ahead of a warp that is still in its flush; a
    parity wait on q_full then sees phase j + 2 as phase j and never returns."""
    caught = 0
    for seed in range(40):
        try:
            simulate(n_rt=1, n_my=5, nb=2 * NS, seed=seed, wait_q_full_in_epilogue=True, limit=60_000)
        except TimeoutError:
            caught += 1
    assert caught > 0


def test_the_model_catches_a_missing_dm_empty_wait():
    """without the wait for the previous tile's second product the dm planes are overwritten under it"""
    caught = 0
    for seed in range(40):
        try:
            simulate(n_rt=5, n_my=2, nb=2 * NS, seed=seed, skip_dm_empty_wait=True, limit=60_000)
        except (AssertionError, TimeoutError):
            caught += 1
    assert caught > 0
