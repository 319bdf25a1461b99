"""An executable model of the synchronisation protocol of the persistent attention backward kernel
(attn_bwd_persist_kernel, vit.rs_b200/csrc/attention_tc.cu; replaces attention_backward, train_vit.rs:559-601).

One CTA walks over its (batch, head) pairs with six roles that share nothing but mbarriers, the in-order tensor pipe and
TMA completions: a loader thread, a statistics warp, the MMA issuer warp, two SIMT groups that alternate over the
(key tile j) x (query sub-tile s) iterations, and a read-out group.  Every barrier's parity is DERIVED from running indices
(head G, iteration n, key tile j) instead of being carried in a flipped variable, so a wrong formula for some sequence length
shows up only as a hang or as silently wrong gradients at that length.  This model restates each role with the kernel's own
parity expressions and checks, for every T in [1, 256] and several heads per CTA, under a random scheduler:

  * no deadlock and no parity aliasing (a wait passes for the phase it means; no barrier runs two phases ahead of a waiter);
  * every arrival lands in the phase it is meant for (arrival counts per phase: 1, 128 or 256);
  * operand lifetimes: a shared-memory tile (Q_i / dO_i rows, K_j, V_j, the four dS tiles) is not refilled or overwritten while an
    MMA that reads it is in flight or before the MMAs that must read it were issued; a score buffer / the dV, dK, dQ accumulators
    are not overwritten before their consumer has read them; the row statistics of a head are not replaced while a group reads
    them.
"""
from tests.pipeline_model import MBar, ProtocolError, Sim

TILE, SUB = 128, 64


def sub_at(j, t, nsub):  # attention_tc.cu: sub_at
    return 2 * nsub - 3 - t if ((j & 1) and not (nsub & 1) and t >= nsub - 2) else t


class AttnBwd(Sim):
    def __init__(self, rng, T, nheads, slow=None, work=2, fault=None):
        super().__init__(rng, slow)
        self.T, self.nheads, self.fault = T, nheads, fault
        NT, NSUB = (T + TILE - 1) // TILE, (T + SUB - 1) // SUB
        N = NT * NSUB
        self.NT, self.NSUB, self.N = NT, NSUB, N
        B = MBar
        load0, load1, load2 = B("load0", 1), B("load1", 1), B("load2", 1)
        s_full = [B(f"s_full[{i}]", 1) for i in range(2)]
        p_full = [B(f"p_full[{i}]", 128) for i in range(2)]
        free_kv = [B(f"free_kv[{i}]", 1) for i in range(2)]
        stat_full = [B(f"stat_full[{i}]", 1) for i in range(2)]
        stat_free = [B(f"stat_free[{i}]", 256) for i in range(2)]
        ds_free = [B(f"ds_free[{i}]", 1) for i in range(4)]
        acc_full, acc_free = B("acc_full", 1), B("acc_free", 128 - (1 if fault == "acc_free_short" else 0))
        dq_full, dq_free, q0_free = B("dq_full", 1), B("dq_free", 128), B("q0_free", 1)
        pipe = self.pipes.setdefault(0, [])
        # ---- instrumentation: what each buffer holds and who still reads it ----
        tile = {}        # ("Q", i) / ("dO", i) / ("K", j) / ("V", j) -> head whose data it holds, or "loading" / "staging"
        tile_readers = {}  # same keys -> MMAs in flight that read it
        must_read = {}   # same keys -> MMAs of the current head that are still to be ISSUED against the tile
        ds_tile = [None] * 4      # (G, j, s) whose dS^T the shared tile holds
        ds_readers = [0] * 4
        score_buf = [None, None]  # (G, n, state): "scores" (complete S^T / dP^T), "packed" (P^T / dS^T in place)
        acc_kv = {"holds": None, "read": True}   # dV_j / dK_j accumulators: (G, j) complete, read out?
        acc_q = {"holds": None, "read": True}    # dQ accumulators of head G
        stats = [None, None]       # head whose row statistics the buffer holds
        stats_readers = [0, 0]
        self.stored = []           # (G, what) in the order the read-out group stored them
        self.done_heads = 0

        def uses(G):  # MMAs of head G still to be issued against each operand tile (lifetime check of the loader's refills)
            for j in range(NT):
                # scores (2 MMA groups: S^T reads K_j and Q_s, dP^T reads V_j and dO_s), dV (dO_s), dK (Q_s), dQ (K_j)
                must_read[(G, ("K", j))] = NSUB + (NSUB + 1) // 2  # scores per sub-tile + one dQ_i per query tile
                must_read[(G, ("V", j))] = NSUB
            for i in range(NT):
                subs = [s for s in range(NSUB) if s // 2 == i]
                must_read[(G, ("Q", i))] = 2 * len(subs) * NT   # per key tile: a score MMA group and a dK MMA group per sub-tile
                must_read[(G, ("dO", i))] = 2 * len(subs) * NT

        def read_op(G, keys):  # an MMA group is issued: it reads these tiles until it retires
            for k in keys:
                if tile.get(k) != G:
                    raise ProtocolError(f"T={T} head {G}: an MMA reads {k}, which holds {tile.get(k)}")
                tile_readers[k] = tile_readers.get(k, 0) + 1
                must_read[(G, k)] -= 1

            def retire():
                for k in keys:
                    tile_readers[k] -= 1
            return retire

        def refill(G, keys, bar, phase):  # loader: TMA into tiles for head G
            for k in keys:
                if tile_readers.get(k, 0) != 0:
                    raise ProtocolError(f"T={T}: {k} refilled for head {G} with {tile_readers[k]} MMAs in flight on it")
                if G > 0 and must_read.get((G - 1, k), 1) != 0:
                    raise ProtocolError(f"T={T}: {k} refilled for head {G}; head {G - 1} still has MMAs to issue on it "
                                        f"({must_read.get((G - 1, k), 'not even started')})")
                if tile.get(k) == "staging":
                    raise ProtocolError(f"T={T}: {k} refilled for head {G} while the read-out group's store still reads it")
                tile[k] = "loading"

                def landed(k=k, G=G, bar=bar, phase=phase):
                    tile[k] = G
                    bar.complete_tx(1, phase)
                self.post(landed)

        def loader():
            for G in range(nheads):
                par = (G - 1) & 1
                if G > 0:
                    yield (free_kv[0], par, G - 1)
                load0.arrive_expect_tx(4, G)
                refill(G, [("K", 0), ("V", 0)], load0, G)
                if G > 0 and fault != "no_q0_wait":
                    yield (q0_free, par, G - 1)
                refill(G, [("Q", 0), ("dO", 0)], load0, G)
                if NT > 1:
                    if G > 0:
                        yield (dq_full, par, G - 1)
                    load1.arrive_expect_tx(2, G)
                    refill(G, [("Q", 1), ("dO", 1)], load1, G)
                    if G > 0:
                        yield (free_kv[1], par, G - 1)
                    load2.arrive_expect_tx(2, G)
                    refill(G, [("K", 1), ("V", 1)], load2, G)
                yield None

        def statistics():
            for G in range(nheads):
                sb = G & 1
                if G >= 2:
                    yield (stat_free[sb], ((G >> 1) - 1) & 1, (G >> 1) - 1)
                if stats_readers[sb] != 0:
                    raise ProtocolError(f"T={T}: statistics buffer {sb} rewritten for head {G} while a group reads it")
                stats[sb] = G
                yield None
                stat_full[sb].arrive(G >> 1)

        def issuer():
            for G in range(nheads):
                uses(G)
                gpar = G & 1
                state = {"q1": False, "kv1": False}
                dq_ready = G == 0
                done_mask = 0

                def wait_tile1(n):
                    if n >= N:
                        return
                    j = n // NSUB
                    s_ = sub_at(j, n - j * NSUB, NSUB)
                    if not state["q1"] and s_ >= TILE // SUB:
                        yield (load1, gpar, G)
                        state["q1"] = True
                    if not state["kv1"] and j > 0:
                        yield (load2, gpar, G)
                        state["kv1"] = True

                def issue_scores(n):
                    j = n // NSUB
                    s_ = sub_at(j, n - j * NSUB, NSUB)
                    bx = n & 1
                    i = s_ >> 1
                    ret = read_op(G, [("K", j), ("Q", i), ("V", j), ("dO", i)])
                    # (K_j and V_j are read once each by the pair of MMA groups: count the pair as one use of each)

                    def scores(ret=ret, bx=bx, n=n, G=G):
                        prev = score_buf[bx]
                        if prev is not None and prev[2] != "consumed":
                            raise ProtocolError(f"T={T} head {G}: scores of iteration {n} overwrite buffer {bx} holding {prev}")
                        score_buf[bx] = (G, n, "scores")
                        ret()
                    pipe.append(scores)
                    k = G * ((N + 1 - bx) >> 1) + (n >> 1)
                    pipe.append(lambda bx=bx, k=k: s_full[bx].arrive(k))

                yield (load0, gpar, G)
                yield from wait_tile1(0)
                yield from wait_tile1(1)
                issue_scores(0)
                if N > 1:
                    issue_scores(1)
                yield None
                for m in range(N):
                    j = m // NSUB
                    t_ = m - j * NSUB
                    s_ = sub_at(j, t_, NSUB)
                    bx = m & 1
                    if t_ == 0:
                        done_mask = 0
                    per_buf = (N + 1 - bx) >> 1
                    done_mask |= 1 << s_
                    partner = s_ ^ 1
                    pair_done = partner >= NSUB or bool((done_mask >> partner) & 1)
                    yield (p_full[bx], (G * per_buf + (m >> 1)) & 1, G * per_buf + (m >> 1))
                    if t_ == 0 and G * NT + j > 0:
                        yield (acc_free, (G * NT + j - 1) & 1, G * NT + j - 1)
                    yield from wait_tile1(m + 2)
                    if pair_done and not dq_ready:
                        yield (dq_free, (G - 1) & 1, G - 1)
                        dq_ready = True
                    # ---- one elected region: dV / dK of iteration m, scores of m + 2, dQ_i when the query tile is complete ----
                    i = s_ >> 1
                    if score_buf[bx] != (G, m, "packed"):
                        raise ProtocolError(f"T={T} head {G}: dV / dK of iteration {m} read buffer {bx} holding {score_buf[bx]}")
                    if t_ == 0:
                        if not acc_kv["read"]:
                            raise ProtocolError(f"T={T} head {G}: dV / dK of key tile {j} restarted before {acc_kv['holds']} was read out")
                        acc_kv["holds"] = None
                    ret = read_op(G, [("dO", i), ("Q", i)])

                    def dvdk(ret=ret, bx=bx, m=m, G=G):
                        score_buf[bx] = (G, m, "consumed")
                        ret()
                    pipe.append(dvdk)
                    if m + 2 < N:
                        issue_scores(m + 2)
                    if j == NT - 1 and s_ < 2 and pair_done:
                        kq = G
                        pipe.append(lambda kq=kq: q0_free.arrive(kq))
                    if pair_done:
                        subs = [s for s in (2 * i, 2 * i + 1) if s < NSUB]
                        for s2 in subs:
                            if ds_tile[s2 & 3] != (G, j, s2):
                                raise ProtocolError(f"T={T} head {G}: dQ_{i} of key tile {j} reads dS tile {s2 & 3} holding {ds_tile[s2 & 3]}")
                            ds_readers[s2 & 3] += 1
                        if j == 0:
                            if not acc_q["read"]:
                                raise ProtocolError(f"T={T} head {G}: dQ restarted before head {acc_q['holds']} was read out")
                        ret = read_op(G, [("K", j)])

                        def dq(ret=ret, subs=subs):
                            for s2 in subs:
                                ds_readers[s2 & 3] -= 1
                            ret()
                        pipe.append(dq)
                        kd = G * NT + j
                        for s2 in subs:
                            pipe.append(lambda b=s2 & 3, kd=kd: ds_free[b].arrive(kd))
                    if t_ == NSUB - 1:
                        ka = G * NT + j

                        def tile_done(ka=ka, j=j, G=G):
                            acc_kv["holds"], acc_kv["read"] = (G, j), False
                            acc_full.arrive(ka)
                        pipe.append(tile_done)
                    yield None

                def head_done(G=G):
                    acc_q["holds"], acc_q["read"] = G, False
                    dq_full.arrive(G)
                pipe.append(head_done)
                yield None

        def group(g):
            per_buf = (N + 1 - g) >> 1
            for G in range(nheads):
                yield (stat_full[G & 1], (G >> 1) & 1, G >> 1)
                if stats[G & 1] != G:
                    raise ProtocolError(f"T={T}: group {g} reads statistics of head {stats[G & 1]} for head {G}")
                stats_readers[G & 1] += 1
                for j in range(NT):
                    for t_ in range(NSUB):
                        n = j * NSUB + t_
                        if (n & 1) != g:
                            continue
                        s_ = sub_at(j, t_, NSUB)
                        yield (s_full[g], (G * per_buf + (n >> 1)) & 1, G * per_buf + (n >> 1))
                        if score_buf[g] != (G, n, "scores"):
                            raise ProtocolError(f"T={T} head {G}: group {g} reads iteration {n} from a buffer holding {score_buf[g]}")
                        bs = s_ & 3
                        if G * NT + j > 0:
                            yield (ds_free[bs], (G * NT + j - 1) & 1, G * NT + j - 1)
                        for _ in range(work):
                            yield None
                        if ds_readers[bs] != 0:
                            raise ProtocolError(f"T={T} head {G}: dS tile {bs} overwritten with {ds_readers[bs]} dQ MMAs in flight on it")
                        ds_tile[bs] = (G, j, s_)
                        score_buf[g] = (G, n, "packed")
                        p_full[g].arrive(G * per_buf + (n >> 1), 128)
                stats_readers[G & 1] -= 1
                stat_free[G & 1].arrive(G >> 1, 128)

        def readout():
            for G in range(nheads):
                for j in range(NT):
                    yield (acc_full, (G * NT + j) & 1, G * NT + j)
                    if acc_kv["holds"] != (G, j):
                        raise ProtocolError(f"T={T}: read-out of key tile {(G, j)} finds {acc_kv['holds']}")
                    yield None
                    acc_kv["read"] = True
                    acc_free.arrive(G * NT + j, 128)
                    for k in (("V", j), ("K", j)):  # dV_j / dK_j leave through the dead V_j / K_j tiles
                        if tile_readers.get(k, 0) != 0:
                            raise ProtocolError(f"T={T}: {k} used as a staging tile with {tile_readers[k]} MMAs in flight on it")
                        if must_read.get((G, k), 1) != 0:
                            raise ProtocolError(f"T={T}: {k} used as a staging tile; head {G} still has {must_read.get((G, k))} MMAs to issue on it")
                        tile[k] = "staging"
                    yield None
                    for k in (("V", j), ("K", j)):  # cp.async.bulk.wait_group.read 0: the stores have read the tiles
                        tile[k] = "dead"
                    self.stored.append((G, f"dkv{j}"))
                    free_kv[j].arrive(G)
                yield (dq_full, G & 1, G)
                if acc_q["holds"] != G:
                    raise ProtocolError(f"T={T}: read-out of dQ of head {G} finds head {acc_q['holds']}")
                yield None
                acc_q["read"] = True
                dq_free.arrive(G, 128)
                self.stored.append((G, "dq"))
                self.done_heads += 1
                yield None

        self.spawn("loader", "loader", loader())
        self.spawn("statistics", "statistics", statistics())
        self.spawn("issuer", "issuer", issuer())
        self.spawn("group", "groupA", group(0))
        self.spawn("group", "groupB", group(1))
        self.spawn("readout", "readout", readout())

    def check_complete(self):
        want = [(G, w) for G in range(self.nheads) for w in [f"dkv{j}" for j in range(self.NT)] + ["dq"]]
        if self.stored != want:
            raise ProtocolError(f"T={self.T}: stores {self.stored[:8]} ..., expected {want[:8]} ...")


def simulate(seed, T, nheads, **kw):
    import random
    sim = AttnBwd(random.Random(seed), T, nheads, **kw)
    sim.run()
    sim.check_complete()
    return True
