"""A tiny DEFLATE *assembler* for tests: builds dynamic-Huffman blocks from explicit code
lengths and token lists, so that odd-but-valid and invalid streams can be produced at will
(the same technique as the reference's programs/test_util.c:210-237 put_bits(), restated)."""
import random


class BitWriter:
    def __init__(self):
        self.acc = 0
        self.n = 0
        self.out = bytearray()

    def put(self, value, nbits):
        self.acc |= (value & ((1 << nbits) - 1)) << self.n
        self.n += nbits
        while self.n >= 8:
            self.out.append(self.acc & 0xff)
            self.acc >>= 8
            self.n -= 8

    def put_code(self, code, length):
        for i in reversed(range(length)):       # Huffman codewords go MSB-first
            self.put((code >> i) & 1, 1)

    def align(self):
        if self.n:
            self.put(0, 8 - self.n)

    def bytes(self):
        self.align()
        return bytes(self.out)


def canonical(lens):
    """symbol -> canonical codeword for the given lengths (0 = unused)."""
    maxl = max(lens) if lens else 0
    cnt = [0] * (maxl + 2)
    for l in lens:
        if l:
            cnt[l] += 1
    code = 0
    nxt = [0] * (maxl + 2)
    for l in range(1, maxl + 1):
        nxt[l] = code
        code = (code + cnt[l]) << 1
    codes = [0] * len(lens)
    for s, l in enumerate(lens):
        if l:
            codes[s] = nxt[l]
            nxt[l] += 1
    return codes


LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
LEN_EXTRA = [0] * 8 + [1] * 4 + [2] * 4 + [3] * 4 + [4] * 4 + [5] * 4 + [0]
OFF_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577]
OFF_EXTRA = [0, 0, 0, 0] + [i // 2 for i in range(2, 28)]
PERM = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]
# a fixed complete precode using all 19 symbols: 13 of length 4, 6 of length 5
PRECODE_LENS = [4] * 13 + [5] * 6


def len_slot(length):
    for s in range(28, -1, -1):
        if length >= LEN_BASE[s]:
            return s
    raise ValueError(length)


def off_slot(offset):
    for s in range(29, -1, -1):
        if offset >= OFF_BASE[s]:
            return s
    raise ValueError(offset)


def write_code_lengths(bw, seq, pre_lens=PRECODE_LENS, rle=True):
    pre_codes = canonical(pre_lens)
    i = 0
    while i < len(seq):
        v = seq[i]
        run = 1
        while i + run < len(seq) and seq[i + run] == v:
            run += 1
        if rle and v == 0 and run >= 11:
            r = min(run, 138)
            bw.put_code(pre_codes[18], pre_lens[18])
            bw.put(r - 11, 7)
            i += r
        elif rle and v == 0 and run >= 3:
            r = min(run, 10)
            bw.put_code(pre_codes[17], pre_lens[17])
            bw.put(r - 3, 3)
            i += r
        elif rle and i > 0 and seq[i - 1] == v and run >= 3:
            r = min(run, 6)
            bw.put_code(pre_codes[16], pre_lens[16])
            bw.put(r - 3, 2)
            i += r
        else:
            bw.put_code(pre_codes[v], pre_lens[v])
            i += 1


def dynamic_block(bw, litlen_lens, offset_lens, tokens, bfinal=1, hlit=None, hdist=None, emit_eob=True):
    """tokens: ints (literal byte) or (length, offset) tuples.  EOB appended."""
    ll = list(litlen_lens) + [0] * (288 - len(litlen_lens))
    ol = list(offset_lens) + [0] * (32 - len(offset_lens))
    if hlit is None:
        hlit = max(257, max([i + 1 for i, l in enumerate(ll) if l] or [0]))
    if hdist is None:
        hdist = max(1, max([i + 1 for i, l in enumerate(ol) if l] or [0]))
    bw.put(bfinal, 1)
    bw.put(2, 2)
    bw.put(hlit - 257, 5)
    bw.put(hdist - 1, 5)
    bw.put(19 - 4, 4)
    for s in PERM:
        bw.put(PRECODE_LENS[s], 3)
    write_code_lengths(bw, ll[:hlit] + ol[:hdist])
    lc, oc = canonical(ll), canonical(ol)
    for t in tokens:
        if isinstance(t, int):
            bw.put_code(lc[t], ll[t])
        else:
            length, offset = t
            s = len_slot(length)
            bw.put_code(lc[257 + s], ll[257 + s])
            bw.put(length - LEN_BASE[s], LEN_EXTRA[s])
            o = off_slot(offset)
            bw.put_code(oc[o], ol[o])
            bw.put(offset - OFF_BASE[o], OFF_EXTRA[o])
    if emit_eob:
        bw.put_code(lc[256], ll[256])


def random_complete_lens(rng, nsyms_used, maxlen, deep_bias=0.7):
    """Code lengths of a complete prefix code with `nsyms_used` leaves and depth <= maxlen."""
    leaves = [0]
    while len(leaves) < nsyms_used:
        cands = [i for i, d in enumerate(leaves) if d < maxlen]
        if not cands:
            break
        if rng.random() < deep_bias:
            i = max(cands, key=lambda k: (leaves[k], rng.random()))
        else:
            i = rng.choice(cands)
        d = leaves.pop(i)
        leaves += [d + 1, d + 1]
    return leaves


def odd_code_stream(rng, n_tokens=400, maxlen=15, n_lit_syms=None, n_off_syms=None):
    """A valid stream whose Huffman codes are arbitrary complete codes (long codewords,
    many subtables), plus the bytes it must decode to."""
    n_lit_syms = n_lit_syms or rng.randint(20, 284)
    n_off_syms = n_off_syms or rng.randint(2, 30)
    lit_syms = set(rng.sample(range(256), min(n_lit_syms - 1, 250)))
    len_syms = set(rng.sample(range(257, 286), rng.randint(1, 25)))
    syms = sorted(lit_syms | len_syms | {256})
    lens = random_complete_lens(rng, len(syms), maxlen, deep_bias=rng.choice([0.3, 0.7, 0.95]))
    rng.shuffle(lens)
    ll = [0] * 288
    for s, l in zip(syms, lens):
        ll[s] = l
    osyms = sorted(rng.sample(range(30), n_off_syms))
    olens = random_complete_lens(rng, len(osyms), maxlen, deep_bias=rng.choice([0.3, 0.9]))
    rng.shuffle(olens)
    ol = [0] * 32
    for s, l in zip(osyms, olens):
        ol[s] = l
    out = bytearray()
    tokens = []
    lits = sorted(lit_syms)
    for _ in range(n_tokens):
        if out and rng.random() < 0.4:
            s = rng.choice(sorted(len_syms)) - 257
            length = LEN_BASE[s] + rng.randrange(1 << LEN_EXTRA[s]) if s < 28 else 258
            if s == 27 and length == 258:
                length = 257      # 258 belongs to symbol 285
            cands = [o for o in osyms if OFF_BASE[o] <= len(out)]
            if not cands:
                continue
            o = rng.choice(cands)
            offset = min(len(out), OFF_BASE[o] + rng.randrange(1 << OFF_EXTRA[o]))
            if off_slot(offset) != o:
                offset = OFF_BASE[o]
            tokens.append((length, offset))
            for _k in range(length):
                out.append(out[-offset])
        else:
            b = rng.choice(lits)
            tokens.append(b)
            out.append(b)
    bw = BitWriter()
    dynamic_block(bw, ll, ol, tokens)
    return bw.bytes(), bytes(out)
