"""CPU model of the streaming kernels' shared-memory ring plan (kernels_feather_tma.cu: k_fts_ring_plan) and of the
producer warps' wait rule (sb_stream.cuh: stream_producer).  The device code cannot run here; this restates the two
small algorithms and checks, over random tile-size sequences (incl. empty tiles and tiles that fill half the ring), the
invariants the kernels rely on:
  * a tile's ring region lies inside the ring and is disjoint from the regions of every tile that may still be in flight
    when its copies land (tiles need[seq] .. seq-1 of the CTA's sequence);
  * need[seq] also covers the reuse of the tile's stage entry (seq - STAGES), is non-decreasing and < = seq;
  * every mbarrier parity wait a producer warp issues is unambiguous: when it waits for tile t it already knows tile
    t - STAGES consumed (a parity wait can only tell "this phase" from "the next")."""
import numpy as np
import pytest

RING_UNITS, STAGES, PRODUCERS = 800, 16, 4          # SB_FTS_RING_BYTES / 128, SB_FTT_STAGES, SB_FTS_PRODUCER_WARPS


def ring_plan(units):
    """k_fts_ring_plan for one CTA's tile sequence -> (start[], need[])"""
    start, need, hist = [], [], {}
    head = tail = oldest = 0
    for seq, u in enumerate(units):
        while True:
            if seq - oldest < STAGES:
                if seq == oldest:
                    head = tail = 0
                if head >= tail:
                    if head + u <= RING_UNITS:
                        s = head
                        break
                    if u < tail:
                        s = 0
                        break
                elif head + u < tail:
                    s = head
                    break
            oldest += 1
            tail = hist[oldest % STAGES] if oldest < seq else head
        head = s + u
        hist[seq % STAGES] = s
        start.append(s)
        need.append(oldest)
    return start, need


@pytest.mark.parametrize("seed", range(12))
def test_ring_plan_and_producer_waits(seed):
    rng = np.random.default_rng(seed)
    kinds = [rng.integers(65, 110, 400),                                  # one-camera tiles (8 KB table + a small box)
             rng.choice([0, 70, 100, 200, 300, 384], 400),                # mixed, with empty tiles and 3-camera tiles
             np.r_[rng.integers(65, 100, 200), np.zeros(60, int), rng.integers(300, 385, 100)],
             rng.choice([0, 0, 0, 384], 400)]
    for units in kinds:
        units = [int(u) for u in units]
        start, need = ring_plan(units)
        prev_need = 0
        for seq, (s, u, nd) in enumerate(zip(start, units, need)):
            assert 0 <= s and s + u <= RING_UNITS
            assert prev_need <= nd <= seq and seq - nd < STAGES          # stage entry of tile seq - STAGES is free, too
            prev_need = nd
            for j in range(nd, seq):                                      # tiles that may still be in flight
                if u and units[j]:
                    assert s + u <= start[j] or start[j] + units[j] <= s, (seq, j)
        # the producer warps (warp w handles tiles w, w + P, ...): what each one knows consumed before every wait
        for pw in range(PRODUCERS):
            known = -1
            for seq in range(pw, len(units), PRODUCERS):
                nd = max(need[seq], seq - STAGES + 1)
                if nd <= 0:
                    continue
                t, w = nd - 1, seq - PRODUCERS
                if w >= 0 and t > w:
                    assert known >= w - STAGES                            # stepping wait on the warp's previous tile
                    known = max(known, w)
                assert known >= t - STAGES, (pw, seq, t, known)           # the wait for tile t is unambiguous
                known = max(known, t)
