"""A lane-by-lane Python model of `merge_insert` (hannoy_b200/csrc/sorted.cuh) — the in-place batched merge both heaps of the walk
use — fuzzed against `sorted(union)[:keep]`.  It pins the algorithm the kernel implements (final positions of the new keys as a
bit mask over the new array, old entries filling the free cells in order, blocks of MERGE_BLOCK tiles moved from the top down with
all loads of a block before its stores), independently of a GPU: a change of the device code that breaks the argument shows up
here first."""
import random

MERGE_BLOCK = 2


def merge_insert_model(a, length, has, key, keep, desc):
    """Mirrors the device code statement by statement; `a` is the array (list), lane L contributes key[L] iff has[L]."""
    lanes = [l for l in range(32) if has[l]]
    new_len = min(length + len(lanes), keep)
    if not lanes or new_len <= 0:
        return max(new_len, 0)
    before = (lambda x, k: x > k) if desc else (lambda x, k: x < k)
    fp = [0xFFFFFFFF] * 32
    for l in lanes:                                   # each lane binary-searches its own key ...
        lo, hi = 0, length
        while lo < hi:
            mid = (lo + hi) >> 1
            if before(a[mid], key[l]):
                lo = mid + 1
            else:
                hi = mid
        fp[l] = lo
    for l in lanes:                                   # ... and ranks it among the new keys: old position + rank = final position
        fp[l] += sum(1 for s in lanes if before(key[s], key[l]))
    t0 = min(fp) >> 5
    t_top = (new_len - 1) >> 5
    d = t_top - t0
    tb = t_top - (d % MERGE_BLOCK if d >= 0 else -((-d) % MERGE_BLOCK))   # C++ remainder
    while tb >= t0:
        below = sum(1 for l in range(32) if fp[l] < (tb << 5))
        held = {}
        for i in range(MERGE_BLOCK):                  # all loads of the block ...
            if tb + i > t_top:
                continue
            f = 0
            for l in range(32):
                if (fp[l] >> 5) == tb + i:
                    f |= 1 << (fp[l] & 31)
            for lane in range(32):
                j = ((tb + i) << 5) + lane
                src = j - below - bin(f & ((1 << lane) - 1)).count("1")
                if j < new_len and not (f >> lane) & 1 and src != j:
                    assert 0 <= src < length
                    held[j] = a[src]
            below += bin(f).count("1")
        for j, v in held.items():                     # ... before its stores
            a[j] = v
        tb -= MERGE_BLOCK
    for l in lanes:
        if fp[l] < new_len:
            a[fp[l]] = key[l]
    return new_len


def test_merge_model_equals_sorted_union():
    rng = random.Random(7)
    for _ in range(20000):
        desc = rng.random() < 0.5
        length = rng.choice([0, 1, 5, 31, 32, 33, 64, 100, 200, 255, 256, 257, 300, 511, 800])
        univ = rng.sample(range(1, 5000), length + 32)
        old = sorted(univ[:length], reverse=desc)
        cap = length + 40
        a = old + [None] * (cap - length)
        m = rng.choice([0, 1, 2, 5, 14, 32])
        lanes = rng.sample(range(32), m)
        has, key = [False] * 32, [0] * 32
        for i, l in enumerate(lanes):
            has[l], key[l] = True, univ[length + i]
        keep = min(cap, max(0, rng.choice([length + m, length - 3, length, 10, 0, length + m - 1, 200])))
        n = merge_insert_model(a, length, has, key, keep, desc)
        want = sorted(old + [key[l] for l in lanes], reverse=desc)[:keep] if m else old[:min(length, keep)]
        assert n == len(want) and a[:n] == want, (length, m, keep, desc)


def test_queue_window_trim_is_a_prefix():
    """Dead queue entries (distance bits above the result set's maximum) form a prefix of the descending array: the window start
    found by the kernel's binary search equals the count of dead entries."""
    rng = random.Random(3)
    for _ in range(2000):
        n = rng.randint(1, 300)
        q = sorted((rng.randrange(1, 1000) for _ in range(n)), reverse=True)
        mb = rng.randrange(0, 1100)
        if not q[0] > mb:
            continue
        lo, hi = 1, n
        while lo < hi:
            mid = (lo + hi) >> 1
            if q[mid] > mb:
                lo = mid + 1
            else:
                hi = mid
        assert lo == sum(1 for x in q if x > mb)
