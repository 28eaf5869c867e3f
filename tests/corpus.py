"""Deterministic test inputs (SURVEY.md section 8d classes, small sizes)."""
import random
import zlib

WORDS = None


def _words():
    global WORDS
    if WORDS is None:
        rng = random.Random(1)
        WORDS = ["".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(2, 12))) for _ in range(4096)]
    return WORDS


def text(n, seed=0):
    """class T: Zipf-ish words."""
    rng = random.Random(seed)
    w = _words()
    out = []
    size = 0
    while size < n:
        k = min(int(rng.paretovariate(1.1)) - 1, len(w) - 1)
        s = w[k] + (" " if rng.random() < 0.9 else "\n")
        out.append(s)
        size += len(s)
    return "".join(out).encode()[:n]


def pattern(n):
    """class P: programs/test_trailing_bytes.c:74-75"""
    return bytes(((i % 123) + (i % 1023)) & 0xff for i in range(n))


def stride(n, s=7):
    """class S: programs/test_litrunlen_overflow.c:36-41 style"""
    return bytes((s * k) % 251 for k in range(n))


def rand(n, seed=0):
    return random.Random(seed).randbytes(n)


def zeros(n):
    return bytes(n)


def mixed(n, seed=0):
    q = n // 4
    return text(q, seed) + pattern(q) + rand(q, seed) + zeros(n - 3 * q)


def all_classes(n, seed=0):
    return {"T": text(n, seed), "P": pattern(n), "S": stride(n), "R": rand(n, seed), "Z": zeros(n), "M": mixed(n, seed)}


def zlib_raw(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=-15):
    c = zlib.compressobj(level, zlib.DEFLATED, wbits, 9, strategy)
    return c.compress(data) + c.flush()
