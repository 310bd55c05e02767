"""CPU model of the tile sort of bilateral_driving_b200/csrc/binning.cu (block_merge_sort + fix_depth_ties): the
position rule of the binary-search merges (left run counts strictly smaller partners, right run smaller-or-equal
ones) must give a permutation for any segment length, also when equal sentinel words are present, and the tie
repair must restore Gaussian-id order among equal depths.  The CUDA kernels follow exactly these rules; their
parity on the GPU is covered by tests/test_render_gpu.py."""
import numpy as np

SENT = np.uint64(0xFFFFFFFFFFFFFFFF)


def merge_sort_model(keys):
    n = len(keys)
    a = keys.copy()
    for c0 in range(0, n, 32):                       # warp_sort32 on 32-key chunks
        a[c0:c0 + 32] = np.sort(a[c0:c0 + 32])
    m = 32
    while m < n:
        b = np.full(n, np.uint64(0x1234))              # poison: every slot must be written exactly once
        written = np.zeros(n, dtype=np.int32)
        for i in range(n):
            run = i // m
            right = run & 1
            pair0 = (run & ~1) * m
            other0 = pair0 if right else pair0 + m
            olen = max(0, min(m, n - other0))
            partner = a[other0:other0 + olen]
            x = a[i]
            lo = int(np.searchsorted(partner, x, side="right" if right else "left"))
            pos = i - (m if right else 0) + lo
            b[pos] = x
            written[pos] += 1
        assert (written == 1).all()
        a = b
        m <<= 1
    return a


def test_merge_rule_sorts_any_length_with_unique_keys():
    rng = np.random.default_rng(0)
    for n in list(range(1, 70)) + [127, 128, 129, 315, 1000, 1354, 2047, 2048]:
        depth = rng.integers(0, 1 << 20, n, dtype=np.uint64)
        slot = rng.permutation(n).astype(np.uint64)
        keys = (depth << np.uint64(32)) | slot
        assert (merge_sort_model(keys) == np.sort(keys)).all()


def test_merge_rule_keeps_equal_sentinels_apart():
    rng = np.random.default_rng(1)
    for n in (40, 100, 333, 1025):
        depth = rng.integers(0, 1 << 20, n, dtype=np.uint64)
        keys = (depth << np.uint64(32)) | np.arange(n, dtype=np.uint64)
        keys[rng.random(n) < 0.2] = SENT              # positions the emission never filled
        assert (merge_sort_model(keys) == np.sort(keys)).all()


def test_tie_repair_restores_gaussian_id_order():
    rng = np.random.default_rng(2)
    n = 500
    depth = rng.integers(0, 40, n, dtype=np.uint64)      # many exact ties
    slot = rng.permutation(n).astype(np.uint64)           # slots are handed out in arbitrary order
    gid = rng.permutation(100000)[:n]                     # Gaussian id of each slot
    keys = np.sort((depth << np.uint64(32)) | slot)       # what the merge sort leaves: (depth, slot) order
    out = keys.copy()
    i = 0
    while i + 1 < n:                                       # fix_depth_ties: insertion sort inside equal-depth runs
        e = i + 1
        while e < n and (out[e] >> np.uint64(32)) == (out[i] >> np.uint64(32)):
            e += 1
        for a in range(i + 1, e):
            ka = out[a]
            ida = gid[int(ka & np.uint64(0xFFFFFFFF))]
            b = a - 1
            while b >= i and gid[int(out[b] & np.uint64(0xFFFFFFFF))] > ida:
                out[b + 1] = out[b]
                b -= 1
            out[b + 1] = ka
        i = e
    d = (out >> np.uint64(32)).astype(np.int64)
    g = gid[(out & np.uint64(0xFFFFFFFF)).astype(np.int64)]
    order = np.lexsort((g, d))                            # gsplat: ascending depth, ties by Gaussian id
    assert (order == np.arange(n)).all()
