"""An executable model of the synchronisation protocol of the persistent tcgen05 GEMM (vit.rs_b200/csrc/gemm_tc.cu).

The kernel is a set of roles that never share a program counter — per CTA one TMA producer thread, one MMA issuer thread
(leader CTA only), eight epilogue warps — coupled only by mbarriers, the in-order tensor pipe, asynchronous TMA / st.async
completions and one global atomic counter (the tile scheduler).  This file restates each role statement by statement as a
Python generator and runs them under a random scheduler with arbitrary delays of every asynchronous completion.  The model
checks what a GPU run cannot show except by hanging or by silently wrong numbers:

  * no deadlock: some role or completion can always make progress until every role has left its loop;
  * no phase aliasing: a parity wait (mbarrier.try_wait.parity) never passes for another phase than the one the waiter means,
    and no barrier runs two phases ahead of a waiter (where the parity test would block for ever);
  * no data hazard: a shared-memory stage is not refilled while an MMA that reads it is in flight, an MMA reads the k-block it
    expects, an accumulator is not overwritten before all 8 (16 for a CTA pair) epilogue warps have read it, a tile-queue slot
    is not overwritten before every consumer has read it;
  * every work unit is processed exactly once, in the same order by every role of a cluster, and the global counter is re-armed
    for the next launch by the last cluster to drain it.

It is a model of the protocol as written in the source, not of the hardware: tests/test_pipeline_model.py runs it over the
shapes of the training step (few and many k-blocks per tile, more and fewer units than clusters, slow epilogues, slow loads).
"""
import random

STAGE_BYTES = 1  # one abstract byte count per CTA and stage


class ProtocolError(AssertionError):
    pass


class MBar:
    """mbarrier: pending arrivals + transaction bytes of the current phase; the phase completes when both reach zero."""

    def __init__(self, name, count):
        self.name, self.count, self.pending, self.tx, self.phase = name, count, count, 0, 0

    def _settle(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self, intended, n=1):
        if self.phase != intended:
            raise ProtocolError(f"{self.name}: arrival meant for phase {intended} lands in phase {self.phase}")
        if self.pending < n:
            raise ProtocolError(f"{self.name}: more arrivals than the barrier expects")
        self.pending -= n
        self._settle()

    def arrive_expect_tx(self, nbytes, intended):
        if self.phase != intended:
            raise ProtocolError(f"{self.name}: expect_tx meant for phase {intended} lands in phase {self.phase}")
        self.tx += nbytes
        self.pending -= 1
        self._settle()

    def complete_tx(self, nbytes, intended):
        if self.phase != intended:
            raise ProtocolError(f"{self.name}: bytes meant for phase {intended} land in phase {self.phase}")
        self.tx -= nbytes
        self._settle()

    def parity_passes(self, parity):  # mbarrier.try_wait.parity: true once the phase with this parity has completed
        return (self.phase & 1) != parity


class Sim:
    """Random scheduler over role generators, asynchronous completions and in-order tensor pipes (one FIFO per key of `pipes`)."""

    def __init__(self, rng, slow=None):
        self.rng = rng
        self.slow = slow or {}  # role kind -> probability of being skipped when picked (makes that role slow)
        self.threads, self.events, self.pipes = [], [], {}

    # ---- scheduling ------------------------------------------------------------------------------------------------
    def spawn(self, kind, name, gen):
        self.threads.append({"kind": kind, "name": name, "gen": gen, "wait": None, "done": False})

    def post(self, fn):  # an asynchronous completion that may be delayed arbitrarily (TMA, st.async, remote arrive)
        self.events.append(fn)

    def run(self, max_steps=10_000_000):
        for _ in range(max_steps):
            ready = [t for t in self.threads if not t["done"] and (t["wait"] is None or t["wait"][0].parity_passes(t["wait"][1]))]
            heads = [p for p in self.pipes.values() if p]
            n = len(ready) + len(self.events) + len(heads)
            if n == 0:
                if all(t["done"] for t in self.threads):
                    return
                stuck = [(t["name"], t["wait"][0].name, t["wait"][1], t["wait"][0].phase) for t in self.threads if not t["done"]]
                raise ProtocolError(f"deadlock: {stuck[:8]} ...")
            i = self.rng.randrange(n)
            if i < len(ready):
                t = ready[i]
                if self.rng.random() < self.slow.get(t["kind"], 0.0):
                    continue
                if t["wait"] is not None:
                    bar, parity, intended = t["wait"]
                    if bar.phase != intended + 1:  # the parity test passed for a phase the waiter does not mean
                        raise ProtocolError(f"{t['name']}: waits for phase {intended} of {bar.name}, parity passed in phase {bar.phase}")
                    t["wait"] = None
                try:
                    req = next(t["gen"])
                except StopIteration:
                    t["done"] = True
                    continue
                if req is not None:
                    bar, parity, intended = req
                    if bar.phase > intended + 1:
                        raise ProtocolError(f"{t['name']}: {bar.name} is in phase {bar.phase}, the waiter still means phase {intended}")
                    t["wait"] = req
            elif i < len(ready) + len(self.events):
                if self.rng.random() < self.slow.get("event", 0.0):
                    continue
                self.events.pop(i - len(ready))()
            else:
                if self.rng.random() < self.slow.get("pipe", 0.0):
                    continue
                heads[i - len(ready) - len(self.events)].pop(0)()  # the tensor pipe retires in order
        raise ProtocolError("step limit reached")


class Launch(Sim):
    """One kernel launch: `clusters` clusters of CG CTAs working through `total_units` units of `kb` k-blocks each."""

    def __init__(self, rng, sched, total_units, kb, clusters, CG=2, STAGES=6, SLOTS=4, EPI=8, dynamic=True, epi_cost=3, slow=None,
                 fault=None):
        super().__init__(rng, slow)
        self.sched, self.total, self.kb, self.clusters = sched, total_units, kb, clusters
        self.fault = fault  # a deliberately broken protocol (tests: the checker must notice)
        self.CG, self.STAGES, self.SLOTS, self.EPI, self.dynamic, self.epi_cost = CG, STAGES, SLOTS, EPI, dynamic, epi_cost
        self.processed = {}  # unit -> how many epilogue warps finished it
        self.order = {}      # (cluster, role) -> [units]
        for cl in range(clusters):
            self._build_cluster(cl)

    # ---- one cluster -----------------------------------------------------------------------------------------------
    def _build_cluster(self, cl):
        CG, STAGES, SLOTS, EPI = self.CG, self.STAGES, self.SLOTS, self.EPI
        B = lambda name, count: MBar(f"cl{cl}.{name}", count)
        full = [[B(f"full[{c}][{s}]", 1) for s in range(STAGES)] for c in range(CG)]     # only the leader's are used
        empty = [[B(f"empty[{c}][{s}]", 1) for s in range(STAGES)] for c in range(CG)]
        tfull = [[B(f"tfull[{c}][{a}]", 1) for a in range(2)] for c in range(CG)]
        tempty = [[B(f"tempty[{c}][{a}]", CG * EPI - (1 if self.fault == "tempty_short" else 0)) for a in range(2)]
                  for c in range(CG)]  # only the leader's are used
        sfull = [[B(f"sfull[{c}][{q}]", 1) for q in range(SLOTS)] for c in range(CG)]
        sempty = [[B(f"sempty[{c}][{q}]", 2 * EPI + 2 if CG == 2 else EPI + 1) for q in range(SLOTS)] for c in range(CG)]
        slot = [[None] * SLOTS for _ in range(CG)]          # (it, unit) as delivered by st.async
        slot_readers = [[0] * SLOTS for _ in range(CG)]     # consumers that still have to read the current content
        content = [[None] * STAGES for _ in range(CG)]      # (unit, kb) in the shared-memory stage, "loading" while TMA writes
        readers = [[0] * STAGES for _ in range(CG)]         # MMAs in flight that read the stage
        acc_unit = [[None, None] for _ in range(CG)]        # unit whose complete accumulator sits in TMEM buffer a
        acc_readers = [[0, 0] for _ in range(CG)]           # epilogue warps that still have to read it
        pipe = self.pipes.setdefault(cl, [])
        total, kbn, unit_step = self.total, self.kb, self.clusters
        consumers_per_cta = [EPI + (1 if c == 0 else 1) for c in range(CG)] if CG == 2 else [EPI + 1]
        # (leader CTA: 8 epilogue warps + the MMA issuer read its slot; peer CTA: 8 epilogue warps + its TMA producer)

        def next_unit(c, it, who):
            q = it % SLOTS
            yield (sfull[c][q], (it // SLOTS) & 1, it // SLOTS)
            got = slot[c][q]
            if got is None or got[0] != it:
                raise ProtocolError(f"cl{cl} {who}: slot {q} of CTA {c} holds {got}, expected iteration {it}")
            slot_readers[c][q] -= 1
            k = it // SLOTS
            self.post(lambda: sempty[0][q].arrive(k))  # (remote) arrive on the leader's empty barrier
            self.order.setdefault((cl, who), []).append(got[1])
            return got[1]

        def fetch(it):
            if self.dynamic:
                v = self.sched[0]
                self.sched[0] += 1
                return v
            return cl + it * unit_step

        def producer(r):
            stage, phase, use = 0, 0, [0] * STAGES
            lookahead = total >= 4 * unit_step
            prefetched = fetch(0) if (r == 0 and lookahead) else 0
            it = 0
            while True:
                if r == 0:
                    q = it % SLOTS
                    if it >= SLOTS and self.fault != "no_sempty_wait":
                        yield (sempty[0][q], ((it // SLOTS) - 1) & 1, it // SLOTS - 1)
                    unit = prefetched if lookahead else fetch(it)
                    if lookahead and unit < total:
                        prefetched = fetch(it + 1)
                    for rr in range(CG):  # sched_publish: arrive.expect_tx on the CTA's full barrier, st.async of the unit number
                        if slot_readers[rr][q] != 0:
                            raise ProtocolError(f"cl{cl}: slot {q} of CTA {rr} republished with {slot_readers[rr][q]} readers outstanding")
                        k = it // SLOTS

                        def expect(rr=rr, q=q, k=k):
                            sfull[rr][q].arrive_expect_tx(4, k)

                        def deliver(rr=rr, q=q, k=k, it=it, unit=unit):
                            slot[rr][q] = (it, unit)
                            slot_readers[rr][q] = consumers_per_cta[rr]
                            sfull[rr][q].complete_tx(4, k)
                        if rr == 0:
                            expect()
                        else:
                            self.post(expect)
                        self.post(deliver)
                    self.order.setdefault((cl, "producer0"), []).append(unit)
                    if unit >= total:
                        if self.dynamic:
                            self.sched[1] += 1
                            if self.sched[1] == unit_step:  # the last cluster to drain the queue re-arms it
                                self.sched[1] = 0
                                self.sched[0] = 0
                        return
                else:
                    unit = yield from next_unit(r, it, f"producer{r}")
                    if unit >= total:
                        return
                for kb in range(kbn):
                    yield (empty[r][stage], phase ^ (0 if self.fault == "empty_parity" else 1), use[stage] - 1)
                    if readers[r][stage] != 0:
                        raise ProtocolError(f"cl{cl} producer{r}: stage {stage} refilled with {readers[r][stage]} MMAs in flight on it")
                    k = use[stage]
                    if r == 0:
                        full[0][stage].arrive_expect_tx(CG * STAGE_BYTES, k)
                    content[r][stage] = "loading"

                    def landed(r=r, stage=stage, unit=unit, kb=kb, k=k):
                        content[r][stage] = (unit, kb)
                        full[0][stage].complete_tx(STAGE_BYTES, k)
                    self.post(landed)
                    use[stage] += 1
                    stage += 1
                    if stage == STAGES:
                        stage, phase = 0, phase ^ 1
                    yield None
                it += 1

        def issuer():
            stage, acc, phase, acc_phase, use, acc_use = 0, 0, 0, 0, [0] * STAGES, [0, 0]
            it = 0
            while True:
                unit = yield from next_unit(0, it, "issuer")
                if unit >= total:
                    return
                yield (tempty[0][acc], acc_phase ^ 1, acc_use[acc] - 1)
                for c in range(CG):
                    if acc_readers[c][acc] != 0:
                        raise ProtocolError(f"cl{cl} issuer: accumulator {acc} of CTA {c} overwritten, {acc_readers[c][acc]} warps have not read it")
                    acc_unit[c][acc] = None
                for kb in range(kbn):
                    yield (full[0][stage], phase, use[stage])
                    for c in range(CG):
                        readers[c][stage] += 1

                    def mma(stage=stage, unit=unit, kb=kb):
                        for c in range(CG):
                            if content[c][stage] != (unit, kb):
                                raise ProtocolError(f"cl{cl}: MMA of unit {unit} k-block {kb} reads stage {stage} of CTA {c} holding {content[c][stage]}")
                            readers[c][stage] -= 1
                    pipe.append(mma)
                    k = use[stage]

                    def freed(stage=stage, k=k):  # tcgen05.commit -> empty barrier of both CTAs (multicast)
                        for c in range(CG):
                            empty[c][stage].arrive(k)
                    pipe.append(freed)
                    use[stage] += 1
                    stage += 1
                    if stage == STAGES:
                        stage, phase = 0, phase ^ 1
                    yield None
                ka = acc_use[acc]

                def published(acc=acc, unit=unit, ka=ka):  # tcgen05.commit -> tfull of both CTAs
                    for c in range(CG):
                        acc_unit[c][acc] = unit
                        acc_readers[c][acc] = EPI
                        tfull[c][acc].arrive(ka)
                pipe.append(published)
                acc_use[acc] += 1
                acc += 1
                if acc == 2:
                    acc, acc_phase = 0, acc_phase ^ 1
                it += 1

        def epilogue(c, w):
            acc, acc_phase, acc_use = 0, 0, [0, 0]
            it = 0
            while True:
                unit = yield from next_unit(c, it, f"epi{c}.{w}")
                if unit >= total:
                    return
                yield (tfull[c][acc], acc_phase, acc_use[acc])
                for _ in range(self.epi_cost):  # tcgen05.ld, arithmetic, TMA stores
                    if acc_unit[c][acc] != unit:
                        raise ProtocolError(f"cl{cl} epi{c}.{w}: reads accumulator {acc} for unit {unit}, it holds {acc_unit[c][acc]}")
                    yield None
                acc_readers[c][acc] -= 1
                self.processed[unit] = self.processed.get(unit, 0) + 1
                ka = acc_use[acc]
                self.post(lambda acc=acc, ka=ka: tempty[0][acc].arrive(ka))
                acc_use[acc] += 1
                acc += 1
                if acc == 2:
                    acc, acc_phase = 0, acc_phase ^ 1
                it += 1

        for r in range(CG):
            self.spawn("producer", f"cl{cl}.producer{r}", producer(r))
        self.spawn("issuer", f"cl{cl}.issuer", issuer())
        for c in range(CG):
            for w in range(EPI):
                self.spawn("epilogue", f"cl{cl}.epi{c}.{w}", epilogue(c, w))

    # ---- what must hold when the launch has drained -----------------------------------------------------------------------
    def check_complete(self):
        want = self.CG * self.EPI
        for u in range(self.total):
            if self.processed.get(u, 0) != want:
                raise ProtocolError(f"unit {u} finished by {self.processed.get(u, 0)} epilogue warps, expected {want}")
        if any(u >= self.total for u in self.processed):
            raise ProtocolError("a unit beyond the problem was processed")
        for cl in range(self.clusters):
            seqs = {who: tuple(s) for (c, who), s in self.order.items() if c == cl}
            if len(set(seqs.values())) != 1:
                raise ProtocolError(f"cluster {cl}: roles disagree on the unit sequence: { {k: v[:6] for k, v in seqs.items()} }")
        if self.dynamic and self.sched != [0, 0]:
            raise ProtocolError(f"tile counter not re-armed: {self.sched}")


def simulate(seed, total_units, kb, clusters, launches=2, **kw):
    """Run `launches` consecutive launches sharing the global tile counter; raises ProtocolError on any violation."""
    rng = random.Random(seed)
    sched = [0, 0]
    for _ in range(launches):
        launch = Launch(rng, sched, total_units, kb, clusters, **kw)
        launch.run()
        launch.check_complete()
    return True
