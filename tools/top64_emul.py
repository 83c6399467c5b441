"""Lane-by-lane emulation of top64_of_256_half (swem_b200/csrc/topl.cuh): the 60-comparator 16-input network checked with the
0-1 principle, the run merges, the two prune-merges with their re-spreading over the lanes, and the final rank layout --
checked against numpy on random distinct words.  Prints the min/max/select and shuffle counts per (pixel, side).

    python tools/top64_emul.py
"""
import numpy as np
N16 = [
[(0,13),(1,12),(2,15),(3,14),(4,8),(5,6),(7,11),(9,10)],
[(0,5),(1,7),(2,9),(3,4),(6,13),(8,14),(10,15),(11,12)],
[(0,1),(2,3),(4,5),(6,8),(7,9),(10,11),(12,13),(14,15)],
[(0,2),(1,3),(4,10),(5,11),(6,7),(8,9),(12,14),(13,15)],
[(1,2),(3,12),(4,6),(5,7),(8,10),(9,11),(13,14)],
[(1,4),(2,6),(5,8),(7,10),(9,13),(11,14)],
[(2,4),(3,6),(9,12),(11,13)],
[(3,5),(6,8),(7,9),(10,12)],
[(3,4),(5,6),(7,8),(9,10),(11,12)],
[(6,7),(8,9)],
]

def check_net16():
    x = ((np.arange(65536)[:, None] >> np.arange(16)[None, :]) & 1).astype(np.int8)
    for layer in N16:
        for (i, j) in layer:
            hi = np.maximum(x[:, i], x[:, j]); lo = np.minimum(x[:, i], x[:, j])
            x[:, i] = hi; x[:, j] = lo
    assert np.all(x[:, :-1] >= x[:, 1:])
    return sum(len(l) for l in N16)
rng = np.random.default_rng(0)

def shfl_xor(a_col, mask):           # a_col: [16 lanes] -> value from lane ^ mask
    return a_col[np.arange(16) ^ mask]

def top64(vals):
    """vals: [256] distinct. returns [64] sorted desc via the emulated network. a[l, k]."""
    a = vals.reshape(16, 16).copy()          # any assignment
    L = np.arange(16)
    nalu = nshfl = 0
    # phase 1a: in-lane 16-sort (descending: lower reg = larger)
    for layer in N16:
        for (i, j) in layer:
            hi = np.maximum(a[:, i], a[:, j]); lo = np.minimum(a[:, i], a[:, j])
            a[:, i] = hi; a[:, j] = lo; nalu += 2
    # phase 1b/1c: merges to 32 and 64 (mirrored + half cleaners), as sort_desc_half for sz = 32, 64
    R = 16
    for sz in (32, 64):
        lmask = sz // R - 1
        take_max = (L & (sz // R // 2)) == 0
        y = np.stack([shfl_xor(a[:, R - 1 - k], lmask) for k in range(R)], 1); nshfl += R
        a = np.where(take_max[:, None], np.maximum(a, y), np.minimum(a, y)); nalu += 2 * R
        d = sz >> 2
        while d > 0:
            if d < R:
                for k in range(R):
                    if (k & d) == 0:
                        hi = np.maximum(a[:, k], a[:, k | d]); lo = np.minimum(a[:, k], a[:, k | d])
                        a[:, k] = hi; a[:, k | d] = lo; nalu += 2
            else:
                ld = d // R
                tm = (L & ld) == 0
                y = np.stack([shfl_xor(a[:, k], ld) for k in range(R)], 1); nshfl += R
                a = np.where(tm[:, None], np.maximum(a, y), np.minimum(a, y)); nalu += 2 * R
            d >>= 1
    # check: runs of 64 sorted
    for r in range(4):
        run = a[4*r:4*r+4].reshape(-1)
        assert np.all(run[:-1] > run[1:])
    # level 1: prune-merge runs (0,1) and (2,3); 16 regs -> 8 regs
    hi1 = (L >> 2) & 1
    t = np.stack([shfl_xor(a[:, 8 + m], 7) for m in range(8)], 1); nshfl += 8
    c = np.stack([np.maximum(a[:, x], t[:, 7 - x]) for x in range(8)], 1); nalu += 8
    b = np.stack([np.where(hi1 == 1, c[:, 7 - m], c[:, m]) for m in range(8)], 1); nalu += 8
    b45 = (L & 3) ^ np.where(hi1 == 1, 3, 0)
    # d = 32: lane ^ 2, max if bit5 == 0
    for (lm, bit) in ((2, (b45 >> 1) & 1), (1, b45 & 1), (7, hi1)):
        y = np.stack([shfl_xor(b[:, m], lm) for m in range(8)], 1); nshfl += 8
        b = np.where((bit == 0)[:, None], np.maximum(b, y), np.minimum(b, y)); nalu += 16
    for d in (4, 2, 1):
        for k in range(8):
            if (k & d) == 0:
                hi = np.maximum(b[:, k], b[:, k | d]); lo = np.minimum(b[:, k], b[:, k | d])
                b[:, k] = hi; b[:, k | d] = lo; nalu += 2
    # check: S0 / S1 sorted 64 with i = b45*16 + hi1*8 + m
    for g in range(2):
        out = np.zeros(64, dtype=vals.dtype)
        for l in range(8 * g, 8 * g + 8):
            for m in range(8):
                out[b45[l] * 16 + hi1[l] * 8 + m] = b[l, m]
        ref = np.sort(vals.reshape(16, 16)[8*g:8*g+8].reshape(-1))[::-1][:64]
        assert np.array_equal(out, ref), (g,)
    # level 2: prune-merge S0 (lanes 0-7) and S1 (lanes 8-15); 8 regs -> 4 regs
    hi2 = (L >> 3) & 1
    t = np.stack([shfl_xor(b[:, 4 + x], 12) for x in range(4)], 1); nshfl += 4
    c = np.stack([np.maximum(b[:, x], t[:, 3 - x]) for x in range(4)], 1); nalu += 4
    e = np.stack([np.where(hi2 == 1, c[:, 3 - n], c[:, n]) for n in range(4)], 1); nalu += 4
    lb = np.where(hi2 == 1, L ^ 12, L)
    hi = (lb >> 2) & 1
    b45 = (lb & 3) ^ np.where(hi == 1, 3, 0)
    for (lm, bit) in ((2, (b45 >> 1) & 1), (1, b45 & 1), (7, hi), (12, hi2)):
        y = np.stack([shfl_xor(e[:, n], lm) for n in range(4)], 1); nshfl += 4
        e = np.where((bit == 0)[:, None], np.maximum(e, y), np.minimum(e, y)); nalu += 8
    for d in (2, 1):
        for k in range(4):
            if (k & d) == 0:
                h_ = np.maximum(e[:, k], e[:, k | d]); l_ = np.minimum(e[:, k], e[:, k | d])
                e[:, k] = h_; e[:, k | d] = l_; nalu += 2
    out = np.zeros(64, dtype=vals.dtype)
    rank = b45 * 16 + hi * 8 + hi2 * 4
    for l in range(16):
        for n in range(4):
            out[rank[l] + n] = e[l, n]
    return out, nalu, nshfl, rank

for trial in range(300):
    vals = rng.permutation(100000)[:256].astype(np.int64)
    out, nalu, nshfl, rank = top64(vals)
    ref = np.sort(vals)[::-1][:64]
    assert np.array_equal(out, ref), trial
print('16-input network: %d compare-exchanges, sorts every 0-1 input' % check_net16())
print('top-64 of 256: ok on 300 random inputs;', nalu, 'min/max/select +', nshfl, 'shuffles; rank base per lane', rank)
