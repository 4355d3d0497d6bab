"""The static program of the two-tile field kernel (csrc/field_tc2.cu), checked without a GPU.

The launcher builds, on the host, the MMA warp's slot list and the epilogue warps' event list; the kernel's roles walk them
and meet only at mbarriers.  `onerf_debug_two_tile_program` returns exactly those tables.  This test replays them in a
small discrete-event model of the barrier protocol - tcgen05 completion is asynchronous (FIFO, arbitrarily late), roles are
scheduled at random - and checks, for every branch configuration and several tile pairs in a row:
  * no deadlock, every role finishes;
  * no mbarrier is completed twice before its (parity-tracking) waiter has consumed the previous phase;
  * XS (the single shared-memory buffer of the encoded input) is only rewritten when no issued MMA can still read it,
    and every X-fed slot reads the X of its own tile and layer;
  * a tile's activations are only overwritten when no issued MMA can still read them, and every hidden slot reads the
    output of the layer in front of it;
  * a tile's raw features are only replaced after the tile's last X production of the pair.
"""
import ctypes
import os
import random

import pytest

from object_nerf_b200 import _lib

EV_XGEN, EPI_DIR = 5, 4
SE_TWO, SE_H1, SE_TILE = 8, 16, 32
XG_TILE, XG_FULL, XG_WAIT_F, XG_RELEASE_F = 1, 2, 4, 8
SLOT_WAIT_H, SLOT_WAIT_XS = 1, 2


def _program(want_scene, want_object, train):
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("library not built")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    buf = (ctypes.c_uint32 * 16384)()
    n = lib.onerf_debug_two_tile_program(int(want_scene), int(want_object), int(train), buf, 16384)
    assert n > 0, n
    w = list(buf[:n])
    n_slots, n_events, n_layers, per = w[:4]
    o = 4
    slots = []
    for _ in range(n_slots):
        tile, layer, half, flags = w[o:o + 4]
        rec = w[o + 4:o + 20]
        o += per
        hdr, gmeta = rec[1], rec[2]
        ngroups = (hdr >> 20) & 0xff
        groups = [(gmeta >> (4 * g)) & 15 for g in range(ngroups)]
        slots.append(dict(rec=rec, tile=tile, layer=layer, half=half, acc=flags >> 4, wait_h=bool(flags & SLOT_WAIT_H),
                          wait_xs=bool(flags & SLOT_WAIT_XS), reads_x=any(not (g >> 3) for g in groups),
                          reads_h=any(g >> 3 for g in groups), hdr=hdr, ngroups=ngroups,
                          rec_acc=(hdr >> 16) & 1, rec_tile=(hdr >> 17) & 1))
    events = []
    for _ in range(n_events):
        x, y, z, ww, ed = w[o:o + 5]
        o += 5
        events.append(dict(x=x, w=ww, ed=ed))
    layers = []
    for _ in range(n_layers):
        N, nhalf, nx, nh, ng = w[o:o + 5]
        groups = w[o + 5:o + 5 + ng]
        o += 5 + 8
        layers.append(dict(N=N, nhalf=nhalf, nslab_x=nx, nslab_h=nh,
                           groups=[dict(first=g & 31, cnt=(g >> 5) & 7, from_h=(g >> 8) & 1) for g in groups]))
    assert o == n
    for s_ in slots:
        s_["layer_info"] = layers[s_["layer"]]
    return slots, events, n_layers


class Bar:
    """An mbarrier with ONE waiter that tracks the phase parity itself (as every waiter in the kernel does)."""

    def __init__(self, name, count=1, pre=0):
        self.name, self.count, self.pending, self.completed, self.consumed = name, count, 0, pre, 0

    def arrive(self):
        self.pending += 1
        if self.pending == self.count:
            self.pending = 0
            self.completed += 1
            assert self.completed - self.consumed <= 1, f"{self.name}: phase completed twice before the waiter consumed one"

    def ready(self):
        return self.completed > self.consumed

    def consume(self):
        assert self.ready()
        self.consumed += 1


def _simulate(slots, events, n_layers, n_pairs, seed):
    rng = random.Random(seed)
    # (layers with no hidden input start a branch; the layer in front of a hidden slot is layer - 1)
    acc_free = [Bar("acc_free0", pre=1), Bar("acc_free1", pre=1)]
    acc_ready = [Bar("acc_ready0"), Bar("acc_ready1")]
    h_ready = [Bar("h_ready0"), Bar("h_ready1")]
    xs_ready = Bar("xs_ready")
    f_ready = [Bar("f_ready0"), Bar("f_ready1")]
    f_free = [Bar("f_free0"), Bar("f_free1")]
    state = dict(xs=None, h=[None, None], f=[None, None])
    inflight = []          # issued, not yet complete: (pair, slot index)
    # which X production does a slot read?  count X-fed layers per tile in program order
    xuse_of_layer = {}
    for s in slots:
        if s["reads_x"] and s["layer"] not in xuse_of_layer:
            xuse_of_layer[s["layer"]] = len(xuse_of_layer)
    n_xuse = len(xuse_of_layer)
    epi_slot_of_event, k = {}, 0
    for i, e in enumerate(events):
        if (e["w"] & 0xff) != EV_XGEN:
            epi_slot_of_event[i] = k
            k += 1
    assert k == len(slots)
    xgen_seen = {}         # (pair, tile) -> number of X productions done

    def mma():
        for pair in range(n_pairs):
            for si, s in enumerate(slots):
                assert s["rec_acc"] == s["acc"] and s["rec_tile"] == s["tile"]
                yield lambda s=s: acc_free[s["acc"]].ready()
                acc_free[s["acc"]].consume()
                if s["wait_h"]:
                    yield lambda s=s: h_ready[s["tile"]].ready()
                    h_ready[s["tile"]].consume()
                if s["wait_xs"]:
                    yield lambda: xs_ready.ready()
                    xs_ready.consume()
                if s["reads_x"]:
                    assert state["xs"] == (pair, s["tile"], xuse_of_layer[s["layer"]]), (state["xs"], pair, s)
                if s["reads_h"]:
                    assert state["h"][s["tile"]] == (pair, s["layer"] - 1), (state["h"], pair, s)
                inflight.append((pair, si))
                yield None

    def completer():
        while True:
            yield lambda: bool(inflight)
            pair, si = inflight.pop(0)
            acc_ready[slots[si]["acc"]].arrive()

    def epilogue():
        for pair in range(n_pairs):
            for ei, e in enumerate(events):
                w = e["w"]
                if (w & 0xff) == EV_XGEN:
                    xf = w >> 8
                    t = xf & XG_TILE
                    if e["x"] != 0xff:
                        yield lambda e=e: acc_ready[e["x"]].ready()          # peek: not consumed
                    if xf & XG_WAIT_F:
                        yield lambda t=t: f_ready[t].ready()
                        f_ready[t].consume()
                    assert state["f"][t] == pair, (state["f"], pair, t)
                    assert not any(slots[si]["reads_x"] for (_, si) in inflight), "XS rewritten under an MMA that reads it"
                    u = xgen_seen.get((pair, t), 0)
                    xgen_seen[(pair, t)] = u + 1
                    state["xs"] = (pair, t, u)
                    xs_ready.arrive()
                    if xf & XG_RELEASE_F:
                        assert u == n_xuse - 1
                        f_free[t].arrive()
                    yield None
                    continue
                s = slots[epi_slot_of_event[ei]]
                kind, acc, t = w & 7, (w >> 16) & 15, 1 if (w & SE_TILE) else 0
                assert acc == s["acc"] and t == s["tile"] and bool(w & SE_H1) == bool(s["half"])
                yield lambda acc=acc: acc_ready[acc].ready()
                acc_ready[acc].consume()
                acc_free[acc].arrive()
                if kind == EPI_DIR or ((w & SE_TWO) and not (w & SE_H1)):
                    yield None
                    continue
                assert not any(slots[si]["tile"] == t and slots[si]["reads_h"] for (_, si) in inflight), \
                    "activations overwritten under an MMA that reads them"
                state["h"][t] = (pair, s["layer"])
                h_ready[t].arrive()
                yield None

    def gather():
        for pair in range(n_pairs):
            for t in range(2):
                if pair > 0:
                    yield lambda t=t: f_free[t].ready()
                    f_free[t].consume()
                    assert xgen_seen.get((pair - 1, t), 0) == n_xuse
                state["f"][t] = pair
                f_ready[t].arrive()
                yield None

    actors = {"mma": mma(), "epi": epilogue(), "gather": gather(), "done": completer()}
    waiting = {k: None for k in actors}
    finished = set()
    steps = 0
    while len(finished) < 3:
        runnable = [k for k in actors if k not in finished and (waiting[k] is None or waiting[k]())]
        if not ({"mma", "epi", "gather"} - finished):
            break
        runnable = [k for k in runnable if not (k == "done" and not inflight and waiting[k] is not None and not waiting[k]())]
        assert runnable, f"deadlock after {steps} steps: xs={state['xs']} inflight={inflight[:3]} finished={finished}"
        k = rng.choice(runnable)
        try:
            waiting[k] = next(actors[k])
        except StopIteration:
            finished.add(k)
        steps += 1
        assert steps < 200000
    return steps


@pytest.mark.parametrize("cfg", [(1, 1, 0), (1, 0, 0), (0, 1, 0), (1, 1, 1)], ids=["both", "scene", "object", "both-train"])
def test_two_tile_program_is_deadlock_and_hazard_free(cfg):
    slots, events, n_layers = _program(*cfg)
    assert len(slots) > 0 and len(events) >= len(slots)
    for seed in range(12):
        _simulate(slots, events, n_layers, n_pairs=3, seed=seed)


def test_two_tile_program_tables_are_consistent():
    slots, events, n_layers = _program(1, 1, 1)
    # training: exactly one slot per tile carries the dump of X, and it is a slot that waits for X
    dump_x = [s for s in slots if s["hdr"] & (1 << 28)]
    assert sorted(s["tile"] for s in dump_x) == [0, 1] and all(s["wait_xs"] and s["layer"] == 0 for s in dump_x)
    assert all(s["hdr"] & (1 << 29) for s in dump_x)        # one-half layer: drained in the same slot
    # every activation slot 1..16 is written by some event of each tile, masks only where a LeakyReLU exists
    seen = {}
    for e in events:
        if (e["w"] & 0xff) != EV_XGEN and e["ed"] & 0xff:
            seen.setdefault(1 if e["w"] & SE_TILE else 0, set()).add((e["ed"] & 0xff) - 1)
    assert seen[0] == set(range(1, 17)) and seen[1] == set(range(1, 17))
    slots_inf, events_inf, _ = _program(1, 1, 0)
    assert not any(s["hdr"] & (3 << 28) for s in slots_inf)


def test_the_model_detects_broken_programs():
    """The checks above have teeth: three protocol bugs, each caught under every schedule tried."""
    slots, events, n_layers = _program(1, 1, 0)

    def caught(sl, ev):
        n = 0
        for seed in range(6):
            try:
                _simulate(sl, ev, n_layers, 3, seed)
            except AssertionError:
                n += 1
        return n

    ev = [dict(e) for e in events]
    for e in ev:                                    # X of tile B produced without waiting for A's MMAs of the layer
        if (e["w"] & 0xff) == EV_XGEN:
            e["x"] = 0xff
    assert caught(slots, ev) == 6
    sl = [dict(s) for s in slots]
    for s in sl:                                    # a hidden layer's first slot does not wait for the activations
        if s["wait_h"] and s["layer"] == 8:
            s["wait_h"] = False
    assert caught(sl, events) == 6
    ev = [dict(e) for e in events]
    done = 0
    for e in ev:                                    # raw features released after the FIRST X production of a tile
        if (e["w"] & 0xff) == EV_XGEN:
            xf = (e["w"] >> 8) & ~XG_RELEASE_F
            if (xf & XG_WAIT_F) and done < 2:
                xf |= XG_RELEASE_F
                done += 1
            e["w"] = EV_XGEN | (xf << 8)
    assert caught(slots, ev) == 6


def test_mma_records_and_the_weight_producer_walk_the_same_ring_stages():
    """The MMA warp reads 64-byte slot records, the weight producer walks layers[].groups[]: two tables, one ring.  They must
    describe the same stages in the same order, and the operand offsets in the records must be the ones the layouts imply."""
    for cfg in [(1, 1, 0), (1, 0, 0), (0, 1, 0), (1, 1, 1)]:
        slots, events, n_layers = _program(*cfg)
        for s in slots:
            L, rec = s["layer_info"], s["rec"]
            half_n = L["N"] >> (L["nhalf"] - 1)
            assert rec[0] == (1 << 4) | (1 << 7) | (1 << 10) | ((half_n >> 3) << 17) | ((128 >> 4) << 24)     # instruction descriptor
            assert (rec[1] & 0xffff) == (half_n * 64) >> 4                                                    # bytes of one slab of this half / 16
            assert s["ngroups"] == len(L["groups"]) <= 5
            x_slabs, h_slabs = [], []
            for g, G in enumerate(L["groups"]):
                meta = (rec[2] >> (4 * g)) & 15
                assert (meta & 7) == G["cnt"] and (meta >> 3) == G["from_h"] and 1 <= G["cnt"] <= 4
                for i in range(G["cnt"]):
                    slab = G["first"] + i
                    rel = (rec[4 + 2 * g + (i >> 1)] >> (16 * (i & 1))) & 0xffff
                    if G["from_h"]:
                        assert rel == 16 * slab                         # TMEM columns: 32 bf16 of K = 16 packed columns
                        h_slabs.append(slab)
                    else:
                        assert rel == (slab >> 1) * 1024 + (slab & 1) * 4   # SWIZZLE_128B atoms of 64 K: 16-byte units
                        x_slabs.append(slab)
            assert x_slabs == list(range(L["nslab_x"])) and h_slabs == list(range(L["nslab_h"]))
            assert s["reads_x"] == (L["nslab_x"] > 0) and s["reads_h"] == (L["nslab_h"] > 0)
            assert s["wait_h"] == (L["nslab_h"] > 0 and s["half"] == 0)
            assert s["wait_xs"] == (L["nslab_x"] > 0 and s["half"] == 0)
