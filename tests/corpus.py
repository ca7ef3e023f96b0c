"""Seeded synthetic corpora of SURVEY.md §8(d), defined with integer arithmetic only so that numpy (tests,
golden scripts, CPU baselines) and torch on the GPU (bench.py) produce the SAME bytes.

  markov_text   order-3 byte Markov chain trained on the reference's doc/*.txt + zip_lib/*.ad? (the model is
                the committed fixture tests/golden/markov3_model.npz, made by tools/make_markov_model.py;
                nothing of /root/reference is read at run time).  The stream is a concatenation of independent
                chains of CHAIN bytes, so any part of it can be generated on its own and in parallel.
  random_bytes  uniform bytes (a counter hash)
  sparse_binary zero runs (about geometric, mean ~200) between short bursts of non-zero bytes, one burst in
                16 is a dense 64-byte record
  mixed         16 MiB stripes cycling {markov text, random, sparse} (BASELINE.json configs[2])
  entries       archive entries of 1-64 KiB, sizes about log-uniform, 3/4 text and 1/4 random (configs[4])

`xp` is the array module: numpy (default) or torch; with torch pass device=.  All index arithmetic is int64;
products wrap identically in both (two's complement) and are masked to 32 bits.
"""
import os
import numpy as np

CHAIN = 16384                      # bytes per Markov chain
M32 = 0xFFFFFFFF
_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class _B:
    """The few array operations the generators need, for numpy or torch."""

    def __init__(self, xp=np, device=None):
        self.xp = xp
        self.torch = xp is not np
        self.device = device

    def arange(self, a, b):
        if self.torch:
            return self.xp.arange(a, b, dtype=self.xp.int64, device=self.device)
        return np.arange(a, b, dtype=np.int64)

    def asarray(self, a, dtype=None):
        if self.torch:
            t = self.xp.as_tensor(np.ascontiguousarray(a), device=self.device)
            return t
        return np.asarray(a)

    def zeros_u8(self, shape):
        if self.torch:
            return self.xp.zeros(shape, dtype=self.xp.uint8, device=self.device)
        return np.zeros(shape, np.uint8)

    def searchsorted_right(self, sorted_arr, v):
        if self.torch:
            return self.xp.searchsorted(sorted_arr, v, right=True)
        return np.searchsorted(sorted_arr, v, side="right").astype(np.int64)

    def searchsorted_left(self, sorted_arr, v):
        if self.torch:
            return self.xp.searchsorted(sorted_arr, v, right=False)
        return np.searchsorted(sorted_arr, v, side="left").astype(np.int64)

    def where(self, c, a, b):
        return self.xp.where(c, a, b)

    def cummax(self, a):
        if self.torch:
            return self.xp.cummax(a, dim=0).values
        return np.maximum.accumulate(a)

    def clip(self, a, lo, hi):
        if self.torch:
            return a.clamp(lo, hi)
        return np.clip(a, lo, hi)

    def u8(self, a):
        if self.torch:
            return a.to(self.xp.uint8)
        return a.astype(np.uint8)

    def cat(self, parts):
        if self.torch:
            return self.xp.cat(parts)
        return np.concatenate(parts)


def h32(idx, seed):
    """32-bit hash of (int64 array idx >= 0, python int seed): murmur3's finaliser, twice."""
    lo = idx & M32
    hi = idx >> 32
    x = lo ^ ((hi * 0x9E3779B1 + (seed & M32) * 0x85EBCA77 + 0x165667B1) & M32)
    for _ in range(2):
        x = x ^ (x >> 16)
        x = (x * 0x85EBCA6B) & M32
        x = x ^ (x >> 13)
        x = (x * 0xC2B2AE35) & M32
        x = x ^ (x >> 16)
        x = (x + 0x9E3779B9) & M32
    return x


_model_cache = {}


def _model(B):
    key = ("t" if B.torch else "n", str(B.device))
    if key in _model_cache:
        return _model_cache[key]
    z = np.load(os.path.join(_GOLDEN, "markov3_model.npz"))
    pairs = z["pairs"].astype(np.int64)
    counts = z["counts"].astype(np.int64)
    ctx = pairs >> 8
    nxt = pairs & 255
    uctx, first = np.unique(ctx, return_index=True)
    cum = np.cumsum(counts)                                # inclusive, global
    before = np.concatenate([[0], cum])[first]             # cumulative count before each context
    last = np.concatenate([first[1:], [pairs.size]]) - 1
    total = cum[last] - before
    m = dict(uctx=B.asarray(uctx), before=B.asarray(before), total=B.asarray(total), cum=B.asarray(cum), nxt=B.asarray(nxt), n_ctx=int(uctx.size))
    _model_cache[key] = m
    return m


def markov_chains(chain_ids, seed, B):
    """Bytes of the chains `chain_ids` (int64 array): array of shape (len(chain_ids), CHAIN), uint8."""
    m = _model(B)
    K = int(chain_ids.shape[0])
    out = B.zeros_u8((K, CHAIN))
    cidx = h32(chain_ids, seed ^ 0x00C0FFEE) % m["n_ctx"]
    ctx = m["uctx"][cidx]
    base = chain_ids * CHAIN
    n_ctx = m["n_ctx"]
    for t in range(CHAIN):
        r = h32(base + t, seed)
        # contexts without a successor in the training text restart from a hashed context
        ci = B.clip(B.searchsorted_left(m["uctx"], ctx), 0, n_ctx - 1)
        ok = m["uctx"][ci] == ctx
        ci = B.where(ok, ci, r % n_ctx)
        target = m["before"][ci] + (r >> 3) % m["total"][ci]
        j = B.searchsorted_right(m["cum"], target)
        b = m["nxt"][j]
        ctx = ((B.where(ok, ctx, m["uctx"][ci]) << 8) | b) & 0xFFFFFF
        out[:, t] = B.u8(b)
    return out


def markov_text(nbytes, seed=0x5EED0001, xp=np, device=None, chain0=0):
    """`nbytes` of Markov text: chains chain0, chain0 + 1, ... back to back."""
    B = _B(xp, device)
    nch = (nbytes + CHAIN - 1) // CHAIN
    parts = []
    for c0 in range(0, nch, 65536):
        c1 = min(nch, c0 + 65536)
        parts.append(markov_chains(B.arange(chain0 + c0, chain0 + c1), seed, B).reshape(-1))
    out = parts[0] if len(parts) == 1 else B.cat(parts)
    return out[:nbytes]


def random_bytes(nbytes, seed=0x5EED0002, xp=np, device=None, pos0=0, lo=0, hi=255):
    B = _B(xp, device)
    parts = []
    for a in range(0, nbytes, 1 << 26):
        b = min(nbytes, a + (1 << 26))
        h = h32(B.arange(pos0 + a, pos0 + b), seed)
        parts.append(B.u8(lo + (h >> 8) % (hi - lo + 1)))
    if not parts:
        return B.zeros_u8((0,))
    return parts[0] if len(parts) == 1 else B.cat(parts)


def sparse_binary(nbytes, seed=0x5EED0003, xp=np, device=None, pos0=0):
    """Position p starts a burst with probability 1/200; a burst is 1..8 non-zero bytes, or a dense 64-byte
    record one time in 16; everything not covered by a burst is zero."""
    B = _B(xp, device)
    parts = []
    for a in range(0, nbytes, 1 << 25):
        b = min(nbytes, a + (1 << 25))
        lead = min(64, pos0 + a)                       # bursts that started before the piece still cover its head
        p = B.arange(pos0 + a - lead, pos0 + b)
        h = h32(p, seed)
        start = (h % 200) == 0
        h2 = h32(p, seed ^ 0x5A5A5A5A)
        ln = B.where((h2 & 15) == 0, 64 + 0 * p, 1 + ((h2 >> 4) & 7))
        cover = B.cummax(B.where(start, p + ln, 0 * p))
        val = 1 + (h32(p, seed ^ 0x0F0F0F0F) >> 8) % 255
        byte = B.where(p < cover, val, 0 * p)
        parts.append(B.u8(byte[lead:]))
    if not parts:
        return B.zeros_u8((0,))
    return parts[0] if len(parts) == 1 else B.cat(parts)


def mixed(nbytes, seed=0x5EED0004, stripe=16 << 20, xp=np, device=None, lo=0, hi=None):
    """Stripes of `stripe` bytes (a multiple of CHAIN) cycling {markov text, random, sparse binary}.
    lo / hi: only the bytes [lo, hi) of the stream (a rank's slice)."""
    assert stripe % CHAIN == 0
    B = _B(xp, device)
    hi = nbytes if hi is None else min(hi, nbytes)
    out = B.zeros_u8((max(0, hi - lo),))
    if hi <= lo:
        return out
    per = stripe // CHAIN
    s0, s1 = lo // stripe, (hi - 1) // stripe + 1
    # all text stripes in one go (the chains are independent: one loop of CHAIN steps for all of them)
    ids = []
    for s in range(s0, s1):
        if s % 3 == 0:
            a, e = max(lo, s * stripe), min(hi, (s + 1) * stripe)
            ids.append(B.arange(a // CHAIN, (e - 1) // CHAIN + 1))
    if ids:
        ids = B.cat(ids)
        for g0 in range(0, int(ids.shape[0]), 131072):
            g = ids[g0:g0 + 131072]
            ch = markov_chains(g, seed, B)
            gl = g.cpu().tolist() if B.torch else g.tolist()
            k = 0
            while k < len(gl):                         # runs of consecutive chains are contiguous in the stream
                k1 = k
                while k1 + 1 < len(gl) and gl[k1 + 1] == gl[k1] + 1:
                    k1 += 1
                a, e = gl[k] * CHAIN, (gl[k1] + 1) * CHAIN
                ca, ce = max(a, lo), min(e, hi)
                out[ca - lo:ce - lo] = ch[k:k1 + 1].reshape(-1)[ca - a:ce - a]
                k = k1 + 1
    for s in range(s0, s1):
        a, e = max(lo, s * stripe), min(hi, (s + 1) * stripe)
        if s % 3 == 1:
            out[a - lo:e - lo] = random_bytes(e - a, seed + 1, xp, device, pos0=a)
        elif s % 3 == 2:
            out[a - lo:e - lo] = sparse_binary(e - a, seed + 2, xp, device, pos0=a)
    return out


def entries(n_entries, seed=0x5EED0005):
    """Config-5 shaped archive entries: sizes about log-uniform in 1..64 KiB (uniform inside a hashed octave),
    three quarters Markov text, one quarter random bytes.  Returns (flat uint8 array with every entry 16-byte
    aligned, offsets, sizes, kinds) as numpy arrays."""
    i = np.arange(n_entries, dtype=np.int64)
    h = h32(i, seed)
    octave = (h % 6).astype(np.int64)
    base = np.int64(1024) << octave
    sizes = base + (h32(i, seed ^ 0x11111111) % base)
    kinds = ((h >> 8) % 4 == 3).astype(np.int64)            # 1 = random
    padded = (sizes + 15) & ~np.int64(15)
    offs = np.concatenate([[0], np.cumsum(padded)[:-1]]).astype(np.int64)
    total = int(padded.sum())
    flat = np.zeros(total, np.uint8)
    # text entries are consecutive slices of one Markov stream, random entries of one hashed stream
    tsz = int(sizes[kinds == 0].sum())
    text = markov_text(tsz, seed + 1) if tsz else np.zeros(0, np.uint8)
    rsz = int(sizes[kinds == 1].sum())
    rnd = random_bytes(rsz, seed + 2) if rsz else np.zeros(0, np.uint8)
    tp = rp = 0
    for k in range(n_entries):
        n = int(sizes[k]); o = int(offs[k])
        if kinds[k] == 0:
            flat[o:o + n] = text[tp:tp + n]; tp += n
        else:
            flat[o:o + n] = rnd[rp:rp + n]; rp += n
    return flat, offs.astype(np.uint64), sizes.astype(np.uint64), kinds


def workload(name, nbytes, seed, xp=np, device=None, lo=0, hi=None):
    """Bytes [lo, hi) of the `nbytes` stream of the named corpus (the whole stream by default)."""
    hi = nbytes if hi is None else min(hi, nbytes)
    if name == "markov":
        c0 = lo // CHAIN
        t = markov_text(hi - c0 * CHAIN, seed, xp, device, chain0=c0)
        return t[lo - c0 * CHAIN:]
    if name == "mixed":
        return mixed(nbytes, seed, 16 << 20, xp, device, lo, hi)
    if name == "random":
        return random_bytes(hi - lo, seed, xp, device, pos0=lo)
    if name == "sparse":
        return sparse_binary(hi - lo, seed, xp, device, pos0=lo)
    raise ValueError(name)
