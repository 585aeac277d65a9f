"""Discrete simulation of the mbarrier protocol of the tcgen05 GEMM with an A transform (gg_gemm_tc.cuh, round 2):
per stage s the barriers full_raw(s) [1 arrival: TMA], full_ab(s) and gt_full(s) [one arrival per converter thread],
empty(s) [2 arrivals: MMA commit + the g_t bulk-store thread], plus the two accumulators' tmem_full / tmem_empty.
Roles run under random interleavings; every wait uses the parity the kernel uses (phase bit flips when the stage index
wraps).  Checked: no deadlock; no arrival lands while the barrier's previous phase is still unconsumed by a waiter that
needs it (parity aliasing: a waiter one phase behind would sail through a wait two phases later); the producer never
overwrites a stage whose MMAs or whose bulk store have not finished; the store thread never reads a stage before every
converter has written its g_t chunk.
Run: python tools/gemm_protocol_sim.py"""
import random


class Bar:
    """mbarrier with a fixed arrival count; phase = number of completed phases (parity = phase & 1)"""

    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier's count in one phase"
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def passed(self, parity):
        """try_wait.parity(parity): true when the phase with that parity has completed, i.e. current parity differs"""
        return (self.phase & 1) != parity


def sim(seed, stages=2, conv=4, tiles=7, kblocks=4):
    rng = random.Random(seed)
    full_raw = [Bar(1) for _ in range(stages)]
    full_ab = [Bar(conv) for _ in range(stages)]
    gt_full = [Bar(conv) for _ in range(stages)]
    empty = [Bar(2) for _ in range(stages)]
    tmem_full = [Bar(1) for _ in range(2)]
    tmem_empty = [Bar(1) for _ in range(2)]          # (the kernel: 256 epilogue threads; one arrival stands for them)
    stage_item = [None] * stages                      # work item (tile, kb) whose data the stage holds
    stage_gt = [0] * stages                           # converters that have written g_t for the current item
    mma_done = [True] * stages
    store_done = [True] * stages

    def ring(role_items):
        s, ph = 0, 0
        for it in role_items:
            yield it, s, ph
            s += 1
            if s == stages:
                s, ph = 0, ph ^ 1

    items = [(t, k) for t in range(tiles) for k in range(kblocks)]

    def producer():
        for it, s, ph in ring(items):
            while not empty[s].passed(ph ^ 1):
                yield
            assert mma_done[s] and store_done[s], ("stage overwritten while in use", it, s)
            stage_item[s], stage_gt[s], mma_done[s], store_done[s] = it, 0, False, False
            full_raw[s].arrive()
            yield

    def converter(_c):
        for it, s, ph in ring(items):
            while not full_raw[s].passed(ph):
                yield
            assert stage_item[s] == it, ("converter read a stale stage", it, stage_item[s])
            stage_gt[s] += 1
            yield
            full_ab[s].arrive()
            gt_full[s].arrive()
            yield

    def mma():
        acc, acc_ph = 0, 0
        it_ring = ring(items)
        for t in range(tiles):
            while not tmem_empty[acc].passed(acc_ph ^ 1):
                yield
            for k in range(kblocks):
                it, s, ph = next(it_ring)
                while not full_ab[s].passed(ph):
                    yield
                assert stage_item[s] == it
                yield
                mma_done[s] = True
                empty[s].arrive()                     # tcgen05.commit -> empty(s)
            tmem_full[acc].arrive()
            acc += 1
            if acc == 2:
                acc, acc_ph = 0, acc_ph ^ 1
            yield

    def store():
        for it, s, ph in ring(items):
            while not gt_full[s].passed(ph):
                yield
            assert stage_item[s] == it and stage_gt[s] == conv, ("bulk store before every converter wrote g_t", it)
            yield                                      # cp.async.bulk ... wait_group.read
            store_done[s] = True
            empty[s].arrive()
            yield

    def epilogue():
        acc, acc_ph = 0, 0
        for t in range(tiles):
            while not tmem_full[acc].passed(acc_ph):
                yield
            yield
            tmem_empty[acc].arrive()
            acc += 1
            if acc == 2:
                acc, acc_ph = 0, acc_ph ^ 1
            yield

    roles = [producer(), mma(), store(), epilogue()] + [converter(c) for c in range(conv)]
    alive = list(range(len(roles)))
    idle_rounds = 0
    while alive:
        before = (tuple(b.phase for bs in (full_raw, full_ab, gt_full, empty, tmem_full, tmem_empty) for b in bs),
                  tuple(b.pending for bs in (full_raw, full_ab, gt_full, empty, tmem_full, tmem_empty) for b in bs),
                  tuple(stage_gt), len(alive))
        rng.shuffle(alive)
        for r in list(alive):
            for _ in range(rng.randint(1, 3)):
                try:
                    next(roles[r])
                except StopIteration:
                    alive.remove(r)
                    break
        after = (tuple(b.phase for bs in (full_raw, full_ab, gt_full, empty, tmem_full, tmem_empty) for b in bs),
                 tuple(b.pending for bs in (full_raw, full_ab, gt_full, empty, tmem_full, tmem_empty) for b in bs),
                 tuple(stage_gt), len(alive))
        idle_rounds = idle_rounds + 1 if before == after else 0
        if idle_rounds > 50:
            raise RuntimeError(("deadlock", seed))
    return True


if __name__ == "__main__":
    for seed in range(2000):
        sim(seed, stages=random.Random(seed).choice([2, 3]), tiles=random.Random(seed + 1).randint(1, 9),
            kblocks=random.Random(seed + 2).choice([1, 4, 8]))
    print("ok")
