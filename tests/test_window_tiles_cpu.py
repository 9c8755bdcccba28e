"""CPU: the tile-skipping arithmetic of the sliding-window kernels, restated in Python line by line
(`decode_window` in csrc/paged_decode.cu; `win_tg / win_skip / real_tile` in csrc/attention_fwd_sm100.cu; `tile_live`
in csrc/attention_fwd.cu) and checked by brute force against the oracle's mask (``oracle.golden.window_mask``, pinned
to the reference's ``_generate_window_mask``): a kernel may skip a KV tile only if NO query row it serves sees ANY key
of it, and it must visit every tile that holds a visible key exactly once, in order."""

import itertools
import random

import torch

from oracle import golden

K_DECODE_TILE = 64    # kTile in paged_decode.cu / attention_fwd.cu
K_TC_TILE = 128       # kBN in attention_fwd_sm100.cu
K_TC_ROWS = 256       # query rows of one CTA (two 128-row tiles)


def decode_window(seq_len, win_local, win_global, tile=K_DECODE_TILE):
    """csrc/paged_decode.cu: decode_window().  Returns (lo, g, visible real tiles in visiting order)."""
    tiles = (seq_len + tile - 1) // tile if seq_len > 0 else 0
    if win_local < 0 and win_global < 0:
        return 0, 0, list(range(tiles))
    lo = max(0, seq_len - 1 - win_local) if win_local >= 0 else seq_len
    g = min(win_global, seq_len) if win_global >= 0 else 0
    t_g = (g + tile - 1) // tile
    t_lo = tiles if lo >= seq_len else lo // tile
    if t_g >= t_lo:
        return lo, g, list(range(tiles))
    n_vis = t_g + (tiles - t_lo)
    return lo, g, [v if v < t_g else t_lo + (v - t_g) for v in range(n_vis)]


def test_decode_visible_tiles_match_the_mask():
    rng = random.Random(0)
    cases = [(1, 0, -1), (64, 0, 0), (65, 63, 64), (4096, 1000, -1), (4096, -1, 100), (300, 5000, 5000)]
    cases += [(rng.randint(1, 3000), rng.choice([-1, 0, 1, 63, 64, 65, 500, 4000]), rng.choice([-1, 0, 1, 64, 100, 2000]))
              for _ in range(300)]
    for seq_len, wl, wg in cases:
        lo, g, tiles = decode_window(seq_len, wl, wg)
        mask = golden.window_mask(1, seq_len, None if wl < 0 else wl, None if wg < 0 else wg)[0]  # [seq_len]
        assert tiles == sorted(set(tiles)), "every tile at most once, in order"
        per_key = torch.tensor([(k < g or k >= lo) for k in range(seq_len)])
        assert torch.equal(per_key, mask), (seq_len, wl, wg)              # the in-tile mask of the kernel
        need = sorted({k // K_DECODE_TILE for k in range(seq_len) if mask[k]})
        assert set(need) <= set(tiles), (seq_len, wl, wg)                 # nothing visible is skipped
        # and nothing is read for nothing, except that the two ranges merge when they touch
        if len(tiles) != (seq_len + K_DECODE_TILE - 1) // K_DECODE_TILE:
            assert tiles == need, (seq_len, wl, wg)


def tcgen05_visible_tiles(q_len, kv_len, m0, win_local, win_global):
    """csrc/attention_fwd_sm100.cu: n_t[] / win_tg / win_skip / real_tile for the CTA whose first query row is m0.
    Returns, per 128-row query tile, the real KV tiles it visits (in order)."""
    off = kv_len - q_len
    has_win = win_local >= 0 or win_global >= 0
    win_g = win_global if (has_win and win_global >= 0) else 0
    n_t = []
    for t in range(2):
        first = m0 + t * 128
        n = 0
        if first < q_len:
            last = min(first + 128, q_len) - 1
            n_end = min(kv_len, off + last + 1)
            n = (n_end + K_TC_TILE - 1) // K_TC_TILE if n_end > 0 else 0
        n_t.append(n)
    win_tg = win_skip = 0
    if has_win:
        win_tg = (win_g + K_TC_TILE - 1) // K_TC_TILE
        if win_local >= 0:
            t_lo = max(0, off + m0 - win_local) // K_TC_TILE
            if t_lo > win_tg:
                win_skip = t_lo - win_tg
            n_t = [n - win_skip if n > win_tg else n for n in n_t]
        else:
            n_t = [min(n, win_tg) for n in n_t]
    return [[j if j < win_tg else j + win_skip for j in range(n)] for n in n_t]


def test_prefill_visible_tiles_match_the_mask():
    rng = random.Random(1)
    for _ in range(120):
        q_len = rng.randint(1, 1500)
        kv_len = q_len + rng.choice([0, 0, 17, 300, 1000])
        wl = rng.choice([-1, 0, 1, 100, 127, 128, 129, 700, 4000])
        wg = rng.choice([-1, 0, 1, 128, 130, 1000])
        mask = golden.window_mask(q_len, kv_len, None if wl < 0 else wl, None if wg < 0 else wg)  # [q_len, kv_len]
        for m0 in range(0, q_len, K_TC_ROWS):
            visited = tcgen05_visible_tiles(q_len, kv_len, m0, wl, wg)
            for t in range(2):
                rows = range(m0 + t * 128, min(m0 + (t + 1) * 128, q_len))
                tiles = visited[t]
                assert tiles == sorted(set(tiles))
                if not len(rows):
                    assert tiles == []
                    continue
                seen = mask[rows.start:rows.stop].any(dim=0)  # keys any row of the tile sees
                need = sorted({k // K_TC_TILE for k in range(kv_len) if seen[k]})
                assert set(need) <= set(tiles), (q_len, kv_len, wl, wg, m0, t)
                # the skip decision is taken per CTA (256 rows), so a query tile may visit a KV tile only its sibling
                # sees; but a tile NO row of the CTA sees must lie inside the visited span (ranges that touch merge)
                cta_seen = mask[m0:min(m0 + K_TC_ROWS, q_len)].any(dim=0)
                cta_need = {k // K_TC_TILE for k in range(kv_len) if cta_seen[k]}
                for e in set(tiles) - cta_need:
                    assert min(cta_need) < e < max(cta_need), (q_len, kv_len, wl, wg, m0, t, e)

def test_general_kernel_tile_live_matches_the_mask():
    """csrc/attention_fwd.cu: tile_live() of a 64-row query block (one CTA), 64-key tiles."""
    rng = random.Random(2)
    for _ in range(150):
        q_len = rng.randint(1, 700)
        kv_len = q_len + rng.choice([0, 9, 400])
        wl = rng.choice([-1, 0, 5, 63, 64, 200, 3000])
        wg = rng.choice([-1, 0, 3, 64, 300])
        if wl < 0 and wg < 0:
            continue
        off = kv_len - q_len
        win_g = wg if wg >= 0 else 0
        mask = golden.window_mask(q_len, kv_len, None if wl < 0 else wl, None if wg < 0 else wg)
        for m0 in range(0, q_len, 64):
            last_row = min(m0 + 64, q_len) - 1
            n_end = min(kv_len, off + last_row + 1)
            n_tiles = (n_end + 63) // 64 if n_end > 0 else 0
            cta_lo = (off + m0 - wl) if wl >= 0 else 2 ** 31 - 1
            live = [it for it in range(n_tiles) if it * 64 < win_g or it * 64 + 63 >= cta_lo]
            seen = mask[m0:last_row + 1].any(dim=0)
            need = {k // 64 for k in range(kv_len) if seen[k]}
            assert need <= set(live), (q_len, kv_len, wl, wg, m0)
