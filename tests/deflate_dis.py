"""A slow, readable DEFLATE *disassembler* for tests and tuning: returns per-block token
statistics of a raw stream (independent of both the oracle and the kernels)."""
import deflate_asm as da


class BitReader:
    def __init__(self, data):
        self.d, self.p = data, 0

    def bit(self):
        byte = self.p >> 3
        v = (self.d[byte] >> (self.p & 7)) & 1 if byte < len(self.d) else 0
        self.p += 1
        return v

    def bits(self, n):
        v = 0
        for i in range(n):
            v |= self.bit() << i
        return v


def _decoder(lens):
    codes = da.canonical(lens)
    table = {}
    for s, l in enumerate(lens):
        if l:
            table[(l, codes[s])] = s
    def dec(br):
        code = 0
        for l in range(1, 16):
            code = (code << 1) | br.bit()
            if (l, code) in table:
                return table[(l, code)]
        raise ValueError("bad code")
    return dec


def disassemble(data):
    br = BitReader(data)
    blocks = []
    out_len = 0
    while True:
        start = br.p
        bfinal = br.bit()
        btype = br.bits(2)
        info = {"type": btype, "lits": 0, "matches": 0, "match_bytes": 0, "hist_len": {}, "off_bits": 0}
        if btype == 0:
            br.p = (br.p + 7) & ~7
            ln = br.bits(16); br.bits(16)
            br.p += 8 * ln
            info["lits"] = ln
            out_len += ln
        else:
            if btype == 2:
                hlit, hdist, hclen = 257 + br.bits(5), 1 + br.bits(5), 4 + br.bits(4)
                pl = [0] * 19
                for i in range(hclen):
                    pl[da.PERM[i]] = br.bits(3)
                pd = _decoder(pl)
                lens = []
                while len(lens) < hlit + hdist:
                    s = pd(br)
                    if s < 16: lens.append(s)
                    elif s == 16: lens += [lens[-1]] * (3 + br.bits(2))
                    elif s == 17: lens += [0] * (3 + br.bits(3))
                    else: lens += [0] * (11 + br.bits(7))
                ll, ol = lens[:hlit], lens[hlit:hlit + hdist]
            else:
                ll = [8] * 144 + [9] * 112 + [7] * 24 + [8] * 8
                ol = [5] * 32
            info["header_bits"] = br.p - start
            ld, od = _decoder(ll), _decoder(ol)
            while True:
                s = ld(br)
                if s < 256:
                    info["lits"] += 1; out_len += 1
                elif s == 256:
                    break
                else:
                    k = min(s - 257, 28)
                    ln = da.LEN_BASE[k] + br.bits(da.LEN_EXTRA[k])
                    o = od(br)
                    off = da.OFF_BASE[min(o, 29)] + br.bits(da.OFF_EXTRA[min(o, 29)])
                    info["matches"] += 1; info["match_bytes"] += ln; out_len += ln
                    info["hist_len"][ln] = info["hist_len"].get(ln, 0) + 1
        info["bits"] = br.p - start
        blocks.append(info)
        if bfinal:
            break
    return blocks, out_len
