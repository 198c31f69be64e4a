"""Static proof, on the launch plan itself, that Fdtd2D.run_streamed is race-free and complete for a given configuration.

run_streamed issues one kernel pass per (row block, pass level) on several streams plus uploads of the medium and
downloads of Ez; the only ordering between them is stream order and the events each item waits for.  Fdtd2D.streamed_plan
returns exactly that as data.  For many configurations (grid heights, block plans, both schedules, windows, stream
counts, slabs) this test builds the happens-before relation and checks that
  * every row an item READS was written by items (or uploads) that happen before it,
  * any two operations NOT ordered by happens-before touch disjoint rows of every array (no write/read, read/write or
    write/write overlap on naz, on either ping-pong state set, on either Ez array),
  * every pass level covers the stored rows exactly once and the downloads cover them exactly once.
A pass of depth d producing rows [lo, hi) reads rows [lo-d, hi+d) of the input set and of naz (the temporally blocked
kernel recomputes d halo rows per side; rows it merely prefetches beyond that are never used).  No GPU needed."""
import itertools
import types

import numpy as np
import pytest

from simulation_b200 import fd2d


def _split(nsteps, tb):
    """pass depths as the library cuts them (fdtd2d_plan): greedy, instantiated depths 1, 2, 3, 4, 6, 8, 12"""
    out, left = [], int(nsteps)
    while left > 0:
        d = min(tb, left)
        d = 12 if d >= 12 else (8 if d > 8 else (d - 1 if d in (5, 7) else d))
        out.append(d)
        left -= d
    return out


def _planner(rows_alloc, row_base=0, tblock=6):
    stub = types.SimpleNamespace(row_base=row_base, rows_alloc=rows_alloc, tblock=tblock, np_dtype=np.dtype(np.float32))
    stub._depths = lambda n, tb=None: _split(n, tb or tblock)      # (the product asks the library: Fdtd2D.pass_depths)
    stub._default_block_rows = lambda schedule, least: fd2d.Fdtd2D._default_block_rows(stub, schedule, least)
    return lambda *a, **k: fd2d.Fdtd2D.streamed_plan(stub, *a, **k)


def _overlap(a, b):
    return max(a[0], b[0]) < min(a[1], b[1])


def check_plan(plan, lossy=False):
    depths, P, edges = plan["depths"], plan["levels"], plan["edges"]
    lo_all, hi_all = plan["stored_rows"]
    B = len(edges) - 1
    clip = lambda lo, hi: (max(lo, lo_all), min(hi, hi_all))
    # ---- operations in issue order: (name, stream, waits, reads {array: rows}, writes {array: rows})
    ops = []
    for k in range(B):
        ops.append((("upload", k), "up", [], {}, {"naz": (edges[k], edges[k + 1])}))
    for it in plan["items"]:
        b, p, (lo, hi) = it["block"], it["level"], it["rows"]
        d = depths[p]
        assert hi - lo >= min(4 * max(depths), hi_all - lo_all), f"block {b} level {p}: {hi - lo} rows"
        reads = {"naz": clip(lo - d, hi + d), f"set{p % 2}": clip(lo - d, hi + d)}
        writes = {f"set{(p + 1) % 2}": (lo, hi)}
        if p == P - 1 or lossy:
            writes[f"ez{(p + 1) % 2}"] = (lo, hi)
        ops.append((("item", (b, p)), f"lane{it['lane']}", list(it["waits"]), reads, writes))
        if p == P - 1:
            ops.append((("download", b), "down", [("item", (b, p))], {f"ez{P % 2}": (lo, hi)}, {}))
    index = {op[0]: i for i, op in enumerate(ops)}
    # ---- happens-before as bitsets (ops are listed in issue order, so every edge points backwards)
    before = [0] * len(ops)
    last_on = {}
    for i, (name, stream, waits, _, _) in enumerate(ops):
        m = 0
        preds = [index[w] for w in waits] + ([last_on[stream]] if stream in last_on else [])
        for j in preds:
            assert j < i, f"{name} waits for something issued later"
            m |= before[j] | (1 << j)
        before[i] = m
        last_on[stream] = i
    hb = lambda i, j: bool((before[j] >> i) & 1)            # op i happens before op j
    # ---- every level tiles the stored rows; downloads too
    for p in range(P):
        rows = sorted(it["rows"] for it in plan["items"] if it["level"] == p)
        assert rows[0][0] == lo_all and rows[-1][1] == hi_all and all(a[1] == b[0] for a, b in zip(rows, rows[1:])), (p, rows)
    # ---- reads are satisfied by writers that happen before
    producers = {}
    for i, (name, _, _, _, writes) in enumerate(ops):
        for arr, rows in writes.items():
            producers.setdefault((arr, name[0], name[1][1] if name[0] == "item" else None), []).append((rows, i))
    for j, (name, _, _, reads, _) in enumerate(ops):
        if name[0] != "item":
            continue
        b, p = name[1]
        for rows, i in producers.get(("naz", "upload", None), []):
            if _overlap(rows, reads["naz"]):
                assert hb(i, j), f"{name} reads naz rows {reads['naz']} before upload {ops[i][0]} is known to be done"
        if p > 0:
            for rows, i in producers.get((f"set{p % 2}", "item", p - 1), []):
                if _overlap(rows, reads[f"set{p % 2}"]):
                    assert hb(i, j), f"{name} reads rows {reads[f'set{p % 2}']} that {ops[i][0]} produces, unordered"
    # ---- unordered operations touch disjoint rows
    for i, j in itertools.combinations(range(len(ops)), 2):
        if hb(i, j) or hb(j, i):
            continue
        (ni, _, _, ri, wi), (nj, _, _, rj, wj) = ops[i], ops[j]
        for arr, rows in wi.items():
            for other in (rj, wj):
                assert not (arr in other and _overlap(rows, other[arr])), f"race on {arr}: {ni} writes {rows}, {nj} touches {other[arr]}"
        for arr, rows in wj.items():
            assert not (arr in ri and _overlap(rows, ri[arr])), f"race on {arr}: {nj} writes {rows}, {ni} reads {ri[arr]}"
    return len(ops)


CONFIGS = [
    # rows_alloc, row_base, nsteps, kwargs
    (32768, 0, 96, {}),                                                     # the bench's e2e call
    (32768, 0, 96, {"schedule": "wavefront"}),
    (32768, 0, 96, {"window": None}),
    (32768, 0, 96, {"window": 3, "streams": 5}),
    (32768, 0, 100, {"streams": 4}),                                        # last pass shallower (depths 6 x 16 + 4)
    (32768 + 192, 32768 - 96, 96, {}),                                      # a slab of the 8-GPU e2e leg (96 ghost rows per side)
    (16384, 0, 24, {"block_rows": 512}),
    (1500, 0, 52, {"blocks": 5, "streams": 4}),
    (1500, 0, 52, {"blocks": 9, "streams": 3, "schedule": "wavefront"}),
    (1500, 0, 40, {"block_rows": [64, 128, 256, 512, 256, 128, 64], "streams": 5}),
    (1500, 0, 40, {"block_rows": [24, 30, 24] * 30, "streams": 5}),
    (1500, 0, 40, {"block_rows": [24, 30, 24] * 30, "streams": 5, "schedule": "wavefront"}),
    (1030, 0, 12, {"block_rows": 1024}),                                    # remainder shorter than a pass allows
    (496, 252, 48, {"blocks": 4, "streams": 3}),                            # slab test geometry
    (700, 0, 7, {"blocks": 2, "tblock": 4}),
    (47, 0, 10, {"block_rows": [26, 19, 12, 102]}),                         # too short to cut (found by tools/fuzz_emulated_2d_modes.py)
    (45, 0, 13, {"block_rows": [60, 111], "schedule": "wavefront"}),
    (4096, 0, 600, {}),                                                     # long run on a short grid: falls back to the wavefront
    (9000, 0, 61, {"tblock": 1, "streams": 16, "block_rows": 700}),
    (32768, 0, 96, {"tblock": 12}),                                         # deep passes: 8 levels of 12 steps
    (32768, 0, 20, {"tblock": 12}),                                         # the driver's bench call: depths 12 + 8
    (32768 + 40, 32768 - 20, 20, {"tblock": 12}),                           # ... on a slab with a 20-row ghost band
    (16384, 0, 50, {"tblock": 12, "block_rows": 1024, "streams": 3}),       # depths 12 x 4 + 2
]


@pytest.mark.parametrize("lossy", [False, True])
@pytest.mark.parametrize("rows_alloc,row_base,nsteps,kw", CONFIGS)
def test_streamed_plan_is_complete_and_race_free(rows_alloc, row_base, nsteps, kw, lossy):
    plan = _planner(rows_alloc, row_base)(nsteps, **kw)
    assert check_plan(plan, lossy) > 0


def test_skewed_plan_lets_a_block_run_as_soon_as_it_has_arrived():
    plan = _planner(32768)(96)
    assert plan["schedule"] == "skewed"
    for it in plan["items"]:
        for kind, key in it["waits"]:
            if kind == "upload":
                assert key == it["block"]                       # never a later block's upload
            else:
                assert key[0] <= it["block"]                    # never a later block's pass
    wave = _planner(32768)(96, schedule="wavefront")
    assert any(kind == "item" and key[0] > it["block"] for it in wave["items"] for kind, key in it["waits"])


def test_the_checker_catches_a_broken_plan():
    plan = _planner(1500)(52, blocks=5, streams=4)
    for it in plan["items"]:                                    # drop the dependency on the previous level
        it["waits"] = [w for w in it["waits"] if w[0] != "item"] or [("upload", it["block"])]
    with pytest.raises(AssertionError):
        check_plan(plan)
    plan = _planner(1500)(52, blocks=5, streams=4, schedule="wavefront")
    for it in plan["items"]:                                    # wavefront blocks with the skewed schedule's (weaker) waits
        if it["level"] > 0:
            it["waits"] = [("item", (it["block"], it["level"] - 1))]
    with pytest.raises(AssertionError):
        check_plan(plan)
