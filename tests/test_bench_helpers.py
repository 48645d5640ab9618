"""Host-side logic of bench.py that runs without a GPU: the band split and its feedback balancing."""
import importlib.util
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_band_rows_partition_the_frame():
    b = _bench()
    for h in (2160, 4320, 600, 389):
        for n in (1, 2, 4, 8):
            rows = [b.band_rows(h, r, n) for r in range(n)]
            assert rows[0][0] == 0 and rows[-1][1] == h
            for (a0, a1), (b0, b1) in zip(rows, rows[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(y0 % 64 == 0 for y0, _ in rows)


def test_rebalanced_moves_rows_towards_the_slow_ranks_and_keeps_a_partition():
    b = _bench()
    bounds = [0, 256, 512, 832, 1088, 1344, 1664, 1920, 2160]
    times = [0.15, 0.19, 0.19, 0.21, 0.19, 0.19, 0.19, 0.15]
    new = b.rebalanced(bounds, times, 2160)
    assert new[0] == 0 and new[-1] == 2160 and len(new) == len(bounds)
    assert all(y1 - y0 >= 16 for y0, y1 in zip(new, new[1:]))
    assert all(y % 16 == 0 for y in new[:-1])
    assert new[1] > bounds[1] and new[-2] < bounds[-2]          # the fast edge ranks take rows from their neighbours
    # equal times: nothing moves
    same = b.rebalanced([0, 544, 1088, 1632, 2160], [0.544, 0.544, 0.544, 0.528], 2160)      # equal cost per row
    assert same == [0, 544, 1088, 1632, 2160]
    # a rank that is far too slow cannot squeeze its neighbours below the quantum
    tight = b.rebalanced([0, 16, 32, 2160], [1.0, 1.0, 100.0], 2160)
    assert all(y1 - y0 >= 16 for y0, y1 in zip(tight, tight[1:])) and tight[-1] == 2160


def test_upload_chunking_rule():
    b = _bench()
    nbytes = 3089856 * 32                      # C4's vertex buffer
    for world in (2, 4, 8):
        assert b.nbytes_ok(nbytes, 4 * world)  # four chunks of one slice per rank, 16-byte aligned
    assert not b.nbytes_ok(0, 8)
    assert not b.nbytes_ok(1000, 8)            # would leave unaligned slices: the caller falls back to one piece


def test_timed_e2e_reports_the_median_run():
    b = _bench()
    calls = {"steps": 0, "finishes": 0}
    clock = {"t": 0.0}
    cost = iter([0.0] * 4 + [1.0] * 5 + [9.0] * 5 + [2.0] * 5)       # warm-up, then three runs of five steps: one hiccup run

    def step(k):
        calls["steps"] += 1
        clock["t"] += next(cost)

    def finish():
        calls["finishes"] += 1

    real = b.time.perf_counter
    b.time.perf_counter = lambda: clock["t"]
    try:
        median, runs = b.timed_e2e(step, finish, 5, None, None, 0)
    finally:
        b.time.perf_counter = real
    assert calls == {"steps": 4 + 15, "finishes": 1 + 3}
    assert runs == [5.0, 45.0, 10.0] and median == 10.0
