"""Seeded synthetic inputs.  No file of /root/reference is read at run time."""
import numpy as np

_LETTERS = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
_LW = np.array([12.7, 9.1, 8.2, 7.5, 7.0, 6.7, 6.3, 6.1, 6.0, 4.3, 4.0, 2.8, 2.8, 2.4, 2.4, 2.2, 2.0, 2.0, 1.9,
                1.5, 1.0, 0.8, 0.15, 0.15, 0.1, 0.07])


def _vocab(rng, nwords=30000):
    lens = np.clip(rng.poisson(4.2, nwords) + 1, 1, 14)
    width = 16
    tab = np.zeros((nwords, width), np.uint8)
    p = _LW / _LW.sum()
    for i in range(nwords):
        tab[i, :lens[i]] = rng.choice(_LETTERS, size=lens[i], p=p)
    return tab, lens


def text(nbytes, seed=0x5EED0001):
    """English-like text: Zipf-distributed pseudo-words, punctuation, line breaks."""
    rng = np.random.default_rng(seed)
    tab, lens = _vocab(np.random.default_rng(12345))
    nwords = tab.shape[0]
    ranks = np.arange(1, nwords + 1, dtype=np.float64)
    p = 1.0 / ranks ** 1.07
    p /= p.sum()
    cdf = np.cumsum(p)
    out = []
    total = 0
    while total < nbytes:
        k = min(4_000_000, (nbytes - total) // 5 + 1024)
        ids = np.searchsorted(cdf, rng.random(k)).clip(0, nwords - 1)
        l = lens[ids]
        sep = np.full(k, 32, np.uint8)
        r = rng.random(k)
        sep[r < 0.08] = ord(",")
        sep[r < 0.045] = ord(".")
        sep[r < 0.012] = 10
        # word bytes followed by a separator (", " / ". " get a space via a second separator column)
        rows = np.zeros((k, 18), np.uint8)
        rows[:, :16] = tab[ids]
        rows[np.arange(k), l] = sep
        extra = (sep == ord(",")) | (sep == ord("."))
        rows[np.arange(k)[extra], l[extra] + 1] = 32
        flat = rows.reshape(-1)
        flat = flat[flat != 0]
        out.append(flat)
        total += flat.size
    return np.concatenate(out)[:nbytes].copy()


def random_bytes(nbytes, seed=0x5EED0002, lo=0, hi=255):
    return np.random.default_rng(seed).integers(lo, hi + 1, nbytes, dtype=np.uint8)


def sparse_binary(nbytes, seed=0x5EED0003):
    """Zero runs (geometric, mean 200) separated by 1..64 non-zero bytes."""
    rng = np.random.default_rng(seed)
    out = np.zeros(nbytes, np.uint8)
    pos = 0
    gaps = rng.geometric(1 / 200.0, nbytes // 100 + 16)
    lens = rng.integers(1, 65, gaps.size)
    for g, l in zip(gaps, lens):
        pos += int(g)
        if pos >= nbytes:
            break
        e = min(nbytes, pos + int(l))
        out[pos:e] = rng.integers(1, 256, e - pos, dtype=np.uint8)
        pos = e
    return out


def mixed(nbytes, stripe=1 << 20, seed=0x5EED0004):
    """Stripes cycling {text, random, sparse binary} (config 3 shape, smaller stripes)."""
    parts = []
    k = 0
    total = 0
    while total < nbytes:
        n = min(stripe, nbytes - total)
        kind = k % 3
        parts.append(text(n, seed + k) if kind == 0 else random_bytes(n, seed + k) if kind == 1 else sparse_binary(n, seed + k))
        total += n
        k += 1
    return np.concatenate(parts)
