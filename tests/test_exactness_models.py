"""CPU models of the three restructurings whose exactness the CUDA kernels rely on, checked against the
literal sequential algorithms in pure Python (no GPU, no oracle):

 * k_segment: the sequential FP64 running sum of Segment_by_Entropy (data_segmentation.adb:75-90) replayed
   as a composition of integer maps "k -> k + a[k mod 4]" over two binades (zip-ada_b200/csrc/b2_segment.cu);
 * k_mtf_seq / k_mtf_seq8: move-to-front kept as the PLACE of every symbol instead of the list
   (zip-ada_b200/csrc/b2_mtf.cu; bzip2-encoding.adb:384-396);
 * k_ent_pm: the forward package-merge may stop as soon as a list repeats the one below
   (zip-ada_b200/csrc/b2_pm.cuh; huffman-encoding-length_limited_coding.adb:131-163);
 * k_keys0: round 0 of the rotation sort by seven characters in radix B plus a bucketed eighth, doubling from 7
   (zip-ada_b200/csrc/b2_bwt.cu; the order of bzip2-encoding.adb:229-255).
"""
import math
import random
from fractions import Fraction


# ---------------------------------------------------------------------------------------------------
# 1. FP64 replay
# ---------------------------------------------------------------------------------------------------
def term_map(d, B, coarse):
    """The map of fl (k u' + d) on k, u' = 2^(B-52), when the result lies on the grid u' (coarse False)
    or 2 u' (coarse True): four increments indexed by k mod 4 (mirror of pf_term in b2_segment.cu)."""
    x = Fraction(d) * Fraction(2) ** (52 - B)          # exact: d is a double, the scale a power of two
    X = math.floor(x)
    fr = x - X
    a = [0, 0, 0, 0]
    if not coarse:
        m = X + (1 if fr > Fraction(1, 2) else 0)
        tie = fr == Fraction(1, 2)
        for r in range(4):
            a[r] = m + (1 if (tie and ((r + m) & 1)) else 0)
    else:
        for b in range(2):
            Y = X + b
            odd = Y & 1
            m = (Y >> 1) + (1 if (odd and fr > 0) else 0)
            tie = bool(odd) and fr == 0
            for qp in range(2):
                a[2 * qp + b] = 2 * (m + (1 if (tie and ((qp + m) & 1)) else 0)) - b
    return a


def compose(f, g):                                      # first f, then g (pf_compose)
    return [f[i] + g[(i + f[i]) & 3] for i in range(4)]


def t_table():
    inv = 1.0 / 16000.0
    return [0.0] + [-(c * inv) * math.log(c * inv) for c in range(1, 16002)]


def _replay_case(rng, E0, n_terms, T):
    """Serial chain of IEEE adds against the composed maps; returns the number of adds checked."""
    eb = math.frexp(E0)[1] - 1                          # E0 in [2^eb, 2^(eb+1))
    B = eb - 1 if E0 < 1.5 * 2.0 ** eb else eb          # two binades [2^B, 2^(B+2)) as in the kernel
    u = 2.0 ** (B - 52)
    split = 2.0 ** (B + 1)
    k = int(Fraction(E0) / Fraction(u))
    assert float(k) * u == E0
    e = E0
    total = [0, 0, 0, 0]
    k0 = k
    checked = 0
    for _ in range(n_terms):
        c = rng.randrange(1, 16000)
        for d in (-T[c], T[c + 1]):                     # "remove the old element, add the new one" (:75-83)
            e2 = e + d                                  # IEEE double add, round to nearest even
            if not (2.0 ** B <= e2 < 2.0 ** (B + 2)):
                return checked                          # the kernel restarts here; nothing to compare
            f = term_map(d, B, e2 >= split)
            k = k + f[k & 3]
            assert Fraction(k) * Fraction(u) == Fraction(e2), (E0, d, e, e2, k)
            total = compose(total, f)
            assert k0 + total[k0 & 3] == k              # the composed map gives the same sum
            e = e2
            checked += 1
    return checked


def test_fp64_sum_replayed_by_integer_maps():
    rng = random.Random(20261017)
    T = t_table()
    checked = 0
    # sums inside a binade, just above / just below a power of two (intermediate sums cross it), ties
    for E0 in (2.9, 3.000000000000001, 1.02, 2.0000000001, 3.9999, 4.3, 5.545, 1.999, 0.51, 0.26, 7.6):
        for _ in range(6):
            checked += _replay_case(rng, E0 * (1 + rng.random() * 1e-9), 400, T)
    assert checked > 20_000


def test_ties_go_to_the_even_neighbour_in_both_grids():
    # d = half a grid step: the result depends on the parity of k (fine grid) or of k / 2 (coarse grid)
    B = 0
    u = 2.0 ** (B - 52)
    for k in (2 ** 52 + 4, 2 ** 52 + 5, 2 ** 52 + 6, 2 ** 52 + 7):
        e = float(k) * u
        f = term_map(u / 2, B, False)
        assert Fraction(k + f[k & 3]) * Fraction(u) == Fraction(e + u / 2)
    for k in (2 ** 53 + 4, 2 ** 53 + 6, 2 ** 53 + 8, 2 ** 53 + 10):
        e = float(k) * u
        for d in (u, 3 * u, -u, 0.5 * u, 1.5 * u):
            f = term_map(d, B, True)
            assert Fraction(k + f[k & 3]) * Fraction(u) == Fraction(e + d), (k, d)


def test_map_composition_is_associative():
    rng = random.Random(7)
    T = t_table()
    maps = [term_map(rng.choice((-1, 1)) * T[rng.randrange(1, 16000)], 1, rng.random() < 0.5) for _ in range(64)]
    left = [0, 0, 0, 0]
    for m in maps:
        left = compose(left, m)
    halves = [[0, 0, 0, 0], [0, 0, 0, 0]]
    for i, m in enumerate(maps):
        halves[i >= 32] = compose(halves[i >= 32], m)
    assert compose(halves[0], halves[1]) == left


# ---------------------------------------------------------------------------------------------------
# 2. move-to-front by places
# ---------------------------------------------------------------------------------------------------
def test_mtf_places_equal_the_shifted_list():
    rng = random.Random(11)
    for n_used in (1, 2, 7, 64, 65, 256):
        symbols = rng.sample(range(256), n_used)
        lst = sorted(symbols)                            # the list starts in increasing order (:374-376)
        code = {b: i for i, b in enumerate(sorted(symbols))}
        place = [255] * 256
        for pos, b in enumerate(lst):
            place[code[b]] = pos
        data = [rng.choice(symbols[:max(1, n_used // 3)]) if rng.random() < 0.7 else rng.choice(symbols) for _ in range(5000)]
        for b in data:
            r_list = lst.index(b)                        # the reference: linear search, shift (:384-396)
            lst.insert(0, lst.pop(r_list))
            c = code[b]
            r = place[c]
            for j in range(256):                         # the kernels: per-byte compare-and-add on packed places
                if place[j] < r:
                    place[j] += 1
            place[c] = 0
            assert r == r_list
        assert [place[code[b]] for b in lst] == list(range(n_used))


# ---------------------------------------------------------------------------------------------------
# 3. forward package-merge with early stop
# ---------------------------------------------------------------------------------------------------
def package_merge_lengths(weights, max_bits, early_stop):
    leaves = sorted(weights)
    n = len(leaves)
    need = 2 * n - 2
    lists = [(list(leaves), [False] * n)]               # (weights, is_package)
    for lev in range(1, max_bits):
        prev = lists[-1][0]
        pk = [prev[2 * b] + prev[2 * b + 1] for b in range(len(prev) // 2)]
        cur, flag, i, j = [], [], 0, 0
        while (i < n or j < len(pk)) and len(cur) < need:
            if j < len(pk) and (i >= n or pk[j] <= leaves[i]):      # a package goes before a leaf of equal weight
                cur.append(pk[j]); flag.append(True); j += 1
            else:
                cur.append(leaves[i]); flag.append(False); i += 1
        lists.append((cur, flag))
        if early_stop and cur == prev:
            while len(lists) < max_bits:
                lists.append((cur, flag))                # every list above repeats this one
            break
    lens = [0] * n
    k = need
    for lev in range(max_bits - 1, -1, -1):
        flag = lists[lev][1]
        p = sum(flag[:k])
        for a in range(k - p):
            lens[a] += 1
        k = 2 * p
    return lens


def test_package_merge_may_stop_when_a_list_repeats():
    rng = random.Random(5)
    for _ in range(300):
        n = rng.choice((2, 3, 5, 17, 40, 90, 258))
        kind = rng.random()
        if kind < 0.3:
            w = [1] * n                                              # Avoid_Zeros on an empty cluster
        elif kind < 0.6:
            w = [max(1, int(1000 * rng.random() ** 6)) for _ in range(n)]
        else:
            w = [rng.randrange(1, 50) for _ in range(n)]
        for max_bits in (15, 16, 17):
            if (1 << max_bits) < n:
                continue
            full = package_merge_lengths(w, max_bits, False)
            assert package_merge_lengths(w, max_bits, True) == full
            assert sum(Fraction(1, 2 ** l) for l in full) <= 1 and max(full) <= max_bits


# ---------------------------------------------------------------------------------------------------
# 4. selector sweep: cost excesses clipped at 7
# ---------------------------------------------------------------------------------------------------
def test_clipped_excess_picks_the_same_coder():
    """k_ent_cost keeps, per group, the bits of every coder above the cheapest one clipped at 7 (4-bit fields);
    the sweep adds the coder's place 1..6 in the selector list and takes the strict minimum, lowest coder on ties
    (bzip2-encoding.adb:683-695).  A coder 7 or more bits above the cheapest can never win."""
    rng = random.Random(3)
    for _ in range(20000):
        ec = rng.randrange(2, 7)
        bits = [rng.randrange(0, 400) if rng.random() < 0.5 else rng.randrange(100, 120) for _ in range(ec)]
        places = rng.sample(range(1, ec + 1), ec)
        exact = min(range(ec), key=lambda c: (bits[c] + places[c], c))
        mn = min(bits)
        clipped = min(range(ec), key=lambda c: (min(bits[c] - mn, 7) + places[c], c))
        assert exact == clipped


# ---------------------------------------------------------------------------------------------------
# 5. chunk chain: 32 probes per trip
# ---------------------------------------------------------------------------------------------------
def search32(tincl, lo, hi, target):
    """The warp-wide search of k_cut_chain (b2_chunks.cu): first index in [lo, hi) whose inclusive prefix
    reaches `target`, or hi; lane l probes the end of the l-th of 32 sub-ranges."""
    while lo < hi:
        span = hi - lo
        step = (span + 31) // 32
        probes = [lo + min((l + 1) * step, span) - 1 for l in range(32)]
        hits = [tincl[p] >= target for p in probes]
        if not any(hits):
            lo = hi
            break
        k = hits.index(True)
        hi = probes[k]
        lo = lo + k * step
    return lo


def test_search32_is_a_lower_bound():
    import bisect
    rng = random.Random(9)
    for _ in range(400):
        n = rng.randrange(1, 5000)
        inc, run = [], 0
        for _ in range(n):
            run += rng.choice((0, 0, 1, 5, 2048, 2560))
            inc.append(run)
        lo = rng.randrange(0, n)
        hi = rng.randrange(lo, n + 1)
        for target in (0, 1, inc[lo], inc[min(n - 1, (lo + hi) // 2)] + rng.choice((0, 1)), inc[-1] + 7):
            want = lo + bisect.bisect_left(inc[lo:hi], target)
            assert search32(inc, lo, hi, target) == want


# ---------------------------------------------------------------------------------------------------
# 6. radix scatter: dispatch order of the tiles and the look-back chain
# ---------------------------------------------------------------------------------------------------
def tiles_rr(ntiles_per_block, group):
    """Mirror of build_tiles_rr (b2_bwt.cu): blocks in groups; inside a group tile k of every block, then
    tile k + 1 of every block; every tile remembers the place of the tile before it in its block."""
    rr, last = [], [None] * len(ntiles_per_block)
    for g0 in range(0, len(ntiles_per_block), group):
        blocks = range(g0, min(len(ntiles_per_block), g0 + group))
        for k in range(max(ntiles_per_block[j] for j in blocks)):
            for j in blocks:
                if ntiles_per_block[j] > k:
                    rr.append((j, k, last[j]))
                    last[j] = len(rr) - 1
    return rr


def test_scatter_dispatch_order_and_lookback_offsets():
    rng = random.Random(13)
    for group in (1, 4, 128):
        nt = [rng.randrange(1, 12) for _ in range(rng.randrange(1, 300))]
        rr = tiles_rr(nt, group)
        assert len(rr) == sum(nt)
        seen = {}
        for pos, (j, k, prev) in enumerate(rr):
            assert k == seen.get(j, 0)                      # the tiles of a block come in order
            seen[j] = k + 1
            assert (prev is None) == (k == 0)
            if prev is not None:
                assert prev < pos and rr[prev][:2] == (j, k - 1)   # the predecessor was dispatched earlier: no deadlock
        # a tile's offset for a digit = digit base of the block + counts of the digit in the block's earlier tiles,
        # obtained by walking the chain until an inclusive state is met
        counts = [[rng.randrange(0, 9) for _ in range(4)] for _ in rr]       # 4 digits are enough here
        inclusive = [None] * len(rr)
        for pos, (j, k, prev) in enumerate(rr):                              # dispatch order = completion order here
            excl = [0] * 4 if prev is None else inclusive[prev]
            inclusive[pos] = [e + c for e, c in zip(excl, counts[pos])]
            want = [sum(counts[p][d] for p, (jj, kk, _) in enumerate(rr) if jj == j and kk < k) for d in range(4)]
            assert excl == want


# ---------------------------------------------------------------------------------------------------
# Round 0 of the rotation sort in seven passes (k_keys0, b2_bwt.cu): seven characters in radix B and the
# eighth cut down to q order-preserving buckets make a key below 2^56; equal keys share seven characters,
# the doubling goes on from 7, and the sorted rotations - hence the BWT string and the origin pointer
# (bzip2-encoding.adb:229-255, :273-280) - are those of the plain sort.
# ---------------------------------------------------------------------------------------------------
def _doubling_sort(text, key0, h0):
    """Cyclic prefix doubling: ranks from the round-0 key, then (rank[i], rank[i + h]) with h = h0, 2 h0, ..."""
    n = len(text)
    order = sorted(range(n), key=lambda i: key0[i])
    rank = [0] * n
    for pos, i in enumerate(order):
        rank[i] = pos if pos == 0 or key0[i] != key0[order[pos - 1]] else rank[order[pos - 1]]
    h = h0
    while h < n and len(set(rank)) < n:
        pair = [(rank[i], rank[(i + h) % n]) for i in range(n)]
        order = sorted(range(n), key=lambda i: pair[i])
        new = [0] * n
        for pos, i in enumerate(order):
            new[i] = pos if pos == 0 or pair[i] != pair[order[pos - 1]] else new[order[pos - 1]]
        rank = new
        h *= 2
    return rank


def test_seven_pass_round0_key_orders_like_the_rotations():
    rng = random.Random(20260117)
    for trial in range(60):
        B = rng.choice([129, 137, 200, 256])
        q = (1 << 56) // B ** 7
        assert q >= 1 and B ** 7 * q <= 1 << 56
        n = rng.choice([1, 2, 5, 7, 8, 9, 40, 300])
        kind = trial % 3
        if kind == 0:
            text = [rng.randrange(B) for _ in range(n)]
        elif kind == 1:
            unit = [rng.randrange(B) for _ in range(rng.choice([1, 2, 3]))]           # periodic: equal rotations
            text = (unit * n)[:n]
        else:
            text = [rng.choice([0, 1, B - 2, B - 1]) for _ in range(n)]                 # long common prefixes, extreme codes
        key0 = []
        for i in range(n):
            k = 0
            for j in range(7):
                k = k * B + text[(i + j) % n]
            c8 = text[(i + 7) % n]
            k = k * q + (c8 * q) // B
            assert k < 1 << 56
            key0.append(k)
        rank = _doubling_sort(text, key0, 7)
        rot = lambda i: text[i:] + text[:i]
        plain = sorted(range(n), key=lambda i: (rot(i), i))
        # rotations in rank order are in lexicographic order, equal ranks = equal rotations
        by_rank = sorted(range(n), key=lambda i: (rank[i], i))
        assert [rot(i) for i in by_rank] == [rot(i) for i in plain], (trial, B, n)
        for a, b in zip(by_rank, by_rank[1:]):
            assert (rank[a] == rank[b]) == (rot(a) == rot(b)), (trial, a, b)
