"""CPU suite: randomised discrete-event simulation of the producer/consumer protocol of the warp-asynchronous GLM kernels
(cmdstan_b200/csrc/glm_dmma.cu: glm_dmma3_kernel with 2 X stages -- validated on the GPU -- and the experimental
glm_dmma4_kernel with 3): eight warps, "full" mbarriers with phase parity completed by the TMA, "done" counters whose LAST
arriver issues the refill of a buffer (X tile NST iterations ahead; the q tile of the next chain tile).  Checked under random
warp speeds and TMA latencies: no deadlock, a buffer is never refilled while a warp still uses it, every wait sees the data of
the iteration it belongs to (no parity aliasing), every warp consumes every (chain tile, row tile) exactly once."""
import random

NW = 8


class MBar:
    def __init__(self):
        self.phase = 0          # completed phases
        self.pending = None     # (ready_time, payload) of the in-flight TMA

    def expect(self, now, latency, payload):
        assert self.pending is None, "expect_tx on a barrier whose previous phase is still in flight"
        self.pending = (now + latency, payload)

    def tick(self, now, data_slot):
        if self.pending and now >= self.pending[0]:
            data_slot[0] = self.pending[1]
            self.pending = None
            self.phase += 1

    def test(self, parity):     # mbarrier.try_wait.parity: true once the phase with this parity has completed
        return (self.phase & 1) != parity


def simulate(nst, n_rt, n_ct, seed):
    rng = random.Random(seed)
    n_it = n_rt * n_ct
    fullX = [MBar() for _ in range(nst)]
    fullB = MBar()
    xdata = [[None] for _ in range(nst)]     # which iteration's tile sits in the buffer
    bdata = [None]
    doneX = [0] * nst
    doneB = [0]
    users = [set() for _ in range(nst)]      # warps currently reading X buffer b
    busers = set()
    consumed = [[] for _ in range(NW)]
    now = [0]

    def load_x(it2):
        b = it2 % nst
        assert not users[b], f"refill of X buffer {b} for iteration {it2} while warps {users[b]} still read it"
        fullX[b].expect(now[0], rng.randint(1, 40), it2)

    def load_b(ct2):
        assert not busers, "refill of the q tile while warps still read it"
        fullB.expect(now[0], rng.randint(1, 40), ct2)

    def warp(w):
        ct = rt = 0
        buf = 0
        xphase = 0
        for it in range(n_it):
            if rt == 0:
                while not fullB.test(ct & 1):
                    yield
                assert bdata[0] == ct, f"warp {w} reads q tile {bdata[0]} for chain tile {ct}"
            while not fullX[buf].test((xphase >> buf) & 1):
                yield
            xphase ^= 1 << buf
            assert xdata[buf][0] == it, f"warp {w} at iteration {it} sees the tile of iteration {xdata[buf][0]}"
            users[buf].add(w)
            busers.add(w)
            for _ in range(rng.randint(0, 6)):      # phase A
                yield
            busers.discard(w)
            if rt == n_rt - 1 and ct + 1 < n_ct:
                doneB[0] += 1
                if doneB[0] == NW:
                    doneB[0] = 0
                    load_b(ct + 1)
            for _ in range(rng.randint(0, 12)):     # link + phase B
                yield
            users[buf].discard(w)
            consumed[w].append((ct, rt))
            if it + nst < n_it:
                doneX[buf] += 1
                if doneX[buf] == NW:
                    doneX[buf] = 0
                    load_x(it + nst)
            if rt == n_rt - 1:
                rt = 0
                ct += 1
            else:
                rt += 1
            buf = (buf + 1) % nst
            if rng.random() < 0.1:                  # a warp that falls behind
                for _ in range(rng.randint(0, 60)):
                    yield

    load_b(0)
    for b in range(min(nst, n_it)):
        load_x(b)
    gens = [warp(w) for w in range(NW)]
    alive = set(range(NW))
    while alive:
        now[0] += 1
        assert now[0] < 2_000_000, "deadlock"
        for b in range(nst):
            fullX[b].tick(now[0], xdata[b])
        fullB.tick(now[0], bdata)
        for w in rng.sample(sorted(alive), len(alive)):
            if rng.random() < 0.7:
                try:
                    next(gens[w])
                except StopIteration:
                    alive.discard(w)
    want = [(c, r) for c in range(n_ct) for r in range(n_rt)]
    assert all(cons == want for cons in consumed)


def test_protocol_two_and_three_stages():
    for seed in range(40):
        rng = random.Random(1000 + seed)
        simulate(nst=rng.choice([2, 3]), n_rt=rng.choice([1, 2, 3, 10, 11]), n_ct=rng.choice([1, 2, 5, 16]), seed=seed)
