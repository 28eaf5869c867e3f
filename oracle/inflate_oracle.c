/*
 * inflate_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * A deliberately simple restatement of the reference decompressor: canonical
 * Huffman decoding one bit at a time, no tables, no fast paths.  What is kept
 * exactly is the CONTRACT: decoded bytes, the verdict for every malformed input,
 * actual_in and actual_out.
 *
 * The reference's bit reader pulls bytes ahead of consumption and counts virtual
 * zero bytes past the end ("overread_count", lib/deflate_decompress.c:214-254).
 * Restated in terms of P = number of bits consumed so far and n = in_nbytes:
 *   R1  bits beyond the input read as zero;
 *   R2  REFILL_BITS() at the top of the generic loop (decompress_template.h:686)
 *       fails when more than 8 bytes would have been over-read, i.e. when a
 *       litlen symbol starts at P >= 8n + 9  -> BAD_DATA;
 *   R3  at the end of the final block all over-read bytes must still be
 *       unconsumed (decompress_template.h:754): P > 8n -> BAD_DATA;
 *   R4  a stored block header requires the byte-aligned position to be real
 *       (decompress_template.h:264): ceil8(P) > 8n -> BAD_DATA.
 * (Inside a block header no output is produced, so an over-read there ends in
 * BAD_DATA at R2/R3/R4 or at a failed SAFETY_CHECK -- the same verdict.)
 */
#include "oracle.h"
#include <string.h>

typedef struct {
	const uint8_t *in;
	size_t n;
	uint64_t P;		/* bits consumed */
} bits_t;

static unsigned getbit(bits_t *b)
{
	uint64_t byte = b->P >> 3;
	unsigned v = byte < b->n ? (b->in[byte] >> (b->P & 7)) & 1 : 0;	/* R1 */
	b->P++;
	return v;
}

static unsigned getbits(bits_t *b, unsigned cnt)
{
	unsigned v = 0;
	for (unsigned i = 0; i < cnt; i++)
		v |= getbit(b) << i;
	return v;
}

/* Canonical code built by the rules of build_decode_table (lib/deflate_decompress.c:721-853). */
typedef struct {
	int special;			/* 0 normal, 1 "every pattern decodes to sym special_sym with 1 bit" */
	unsigned special_sym;
	unsigned count[16];
	unsigned first[16];		/* first canonical code of each length */
	unsigned offs[16];		/* index into sorted[] of the first symbol of each length */
	uint16_t sorted[288];
	unsigned maxlen;
} code_t;

static int build_code(code_t *c, const uint8_t *lens, unsigned nsyms, unsigned max_codeword_len)
{
	memset(c, 0, sizeof(*c));
	for (unsigned s = 0; s < nsyms; s++)
		c->count[lens[s]]++;
	unsigned maxlen = max_codeword_len;
	while (maxlen > 1 && c->count[maxlen] == 0)		/* :753-758 */
		maxlen--;
	c->maxlen = maxlen;
	uint32_t used = 0;
	for (unsigned l = 1; l <= maxlen; l++)			/* :781-786 */
		used = (used << 1) + c->count[l];
	if (used > (1u << maxlen))				/* overfull :800 */
		return 0;
	if (used < (1u << maxlen)) {				/* incomplete :804-853 */
		c->special = 1;
		if (used == 0) {
			c->special_sym = 0;
		} else {
			if (used != (1u << (maxlen - 1)) || c->count[1] != 1)
				return 0;
			for (unsigned s = 0; s < nsyms; s++)
				if (lens[s] == 1) { c->special_sym = s; break; }
		}
		return 1;
	}
	unsigned code = 0, idx = 0;
	for (unsigned l = 1; l <= 15; l++) {
		c->first[l] = code;
		c->offs[l] = idx;
		code = (code + c->count[l]) << 1;
		idx += c->count[l];
	}
	unsigned next[16];
	memcpy(next, c->offs, sizeof(next));
	for (unsigned s = 0; s < nsyms; s++)
		if (lens[s])
			c->sorted[next[lens[s]]++] = (uint16_t)s;
	return 1;
}

static unsigned decode_sym(const code_t *c, bits_t *b)
{
	if (c->special) {
		getbit(b);
		return c->special_sym;
	}
	unsigned code = 0;
	for (unsigned l = 1; l <= c->maxlen; l++) {
		code = (code << 1) | getbit(b);		/* codewords are stored MSB-first in LSB-first bit order */
		if (code - c->first[l] < c->count[l])
			return c->sorted[c->offs[l] + (code - c->first[l])];
	}
	return 0;	/* unreachable for a complete code */
}

/* Appendix-A tables (ref: lib/deflate_decompress.c:576-587, 616-627) */
static const uint16_t len_base[31] = {3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258,258,258};
static const uint8_t len_extra[31] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0,0,0};
static const uint16_t off_base[32] = {1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577,24577,24577};
static const uint8_t off_extra[32] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13,13,13};

static int inflate_raw(const uint8_t *in, size_t n, uint8_t *out, size_t avail, int exact,
		       size_t *actual_in, size_t *actual_out)
{
	static const uint8_t perm[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};
	bits_t b = {in, n, 0};
	size_t op = 0;
	unsigned bfinal;
	do {
		bfinal = getbit(&b);
		unsigned btype = getbits(&b, 2);
		code_t lit, off;
		uint8_t lens[288 + 32 + 140];
		if (btype == 2) {					/* dynamic: decompress_template.h:85-245 */
			unsigned hlit = 257 + getbits(&b, 5), hdist = 1 + getbits(&b, 5), hclen = 4 + getbits(&b, 4);
			uint8_t plens[19];
			memset(plens, 0, sizeof(plens));
			for (unsigned i = 0; i < hclen; i++)
				plens[perm[i]] = (uint8_t)getbits(&b, 3);
			code_t pre;
			if (!build_code(&pre, plens, 19, 7))
				return ORC_BAD_DATA;
			unsigned i = 0, total = hlit + hdist;
			while (i < total) {
				unsigned ps = decode_sym(&pre, &b);
				if (ps < 16) { lens[i++] = (uint8_t)ps; continue; }
				unsigned rep; uint8_t val;
				if (ps == 16) {
					if (i == 0) return ORC_BAD_DATA;	/* :202 */
					val = lens[i - 1];
					rep = 3 + getbits(&b, 2);
				} else if (ps == 17) {
					val = 0; rep = 3 + getbits(&b, 3);
				} else {
					val = 0; rep = 11 + getbits(&b, 7);
				}
				for (unsigned k = 0; k < rep; k++) lens[i + k] = val;	/* array has the 138-entry slack */
				i += rep;
			}
			if (i != total) return ORC_BAD_DATA;			/* :245 */
			if (!build_code(&off, lens + hlit, hdist, 15)) return ORC_BAD_DATA;	/* :331 */
			if (!build_code(&lit, lens, hlit, 15)) return ORC_BAD_DATA;		/* :332 */
		} else if (btype == 0) {				/* stored: decompress_template.h:247-285 */
			uint64_t Pa = (b.P + 7) & ~(uint64_t)7;
			if (Pa > (uint64_t)n * 8) return ORC_BAD_DATA;		/* R4 */
			size_t B = (size_t)(Pa >> 3);
			if (n - B < 4) return ORC_BAD_DATA;
			unsigned len = in[B] | (in[B + 1] << 8), nlen = in[B + 2] | (in[B + 3] << 8);
			if (len != (nlen ^ 0xffffu)) return ORC_BAD_DATA;
			if (len > avail - op) return ORC_INSUFFICIENT_SPACE;
			if (len > n - (B + 4)) return ORC_BAD_DATA;
			memcpy(out + op, in + B + 4, len);
			op += len;
			b.P = (uint64_t)(B + 4 + len) * 8;
			continue;
		} else if (btype == 1) {				/* static: decompress_template.h:287-327 */
			unsigned i;
			for (i = 0; i < 144; i++) lens[i] = 8;
			for (; i < 256; i++) lens[i] = 9;
			for (; i < 280; i++) lens[i] = 7;
			for (; i < 288; i++) lens[i] = 8;
			for (; i < 320; i++) lens[i] = 5;
			build_code(&off, lens + 288, 32, 15);
			build_code(&lit, lens, 288, 15);
		} else {
			return ORC_BAD_DATA;				/* :290 */
		}
		for (;;) {						/* generic loop: decompress_template.h:680-738 */
			if (b.P >= (uint64_t)n * 8 + 9) return ORC_BAD_DATA;	/* R2 */
			unsigned sym = decode_sym(&lit, &b);
			if (sym < 256) {
				if (op == avail) return ORC_INSUFFICIENT_SPACE;	/* :700-701 */
				out[op++] = (uint8_t)sym;
				continue;
			}
			if (sym == 256) break;
			unsigned length = len_base[sym - 257] + getbits(&b, len_extra[sym - 257]);
			if (length > avail - op) return ORC_INSUFFICIENT_SPACE;	/* :708-709 */
			unsigned osym = decode_sym(&off, &b);
			unsigned offset = off_base[osym] + getbits(&b, off_extra[osym]);
			if (offset > op) return ORC_BAD_DATA;			/* :727 */
			for (unsigned k = 0; k < length; k++, op++) out[op] = out[op - offset];
		}
	} while (!bfinal);
	if (b.P > (uint64_t)n * 8) return ORC_BAD_DATA;			/* R3 */
	if (actual_in) *actual_in = (size_t)((b.P + 7) >> 3);		/* :757-762 */
	if (!exact) { if (actual_out) *actual_out = op; }
	else if (op != avail) return ORC_SHORT_OUTPUT;				/* :765-770 */
	return ORC_SUCCESS;
}

int oracle_decompress(int format, const void *in_, size_t n, void *out, size_t avail, int exact,
		      size_t *actual_in, size_t *actual_out)
{
	const uint8_t *in = (const uint8_t *)in_;
	size_t pos = 0, ain = 0, aout = 0;
	int r;
	if (format == ORC_RAW)
		return inflate_raw(in, n, (uint8_t *)out, avail, exact, actual_in, actual_out);
	if (format == ORC_ZLIB) {					/* lib/zlib_decompress.c:45-94 */
		if (n < 6) return ORC_BAD_DATA;
		unsigned hdr = (in[0] << 8) | in[1];
		if (hdr % 31) return ORC_BAD_DATA;
		if (((hdr >> 8) & 0xf) != 8) return ORC_BAD_DATA;
		if ((hdr >> 12) > 7) return ORC_BAD_DATA;
		if ((hdr >> 5) & 1) return ORC_BAD_DATA;
		r = inflate_raw(in + 2, n - 6, (uint8_t *)out, avail, exact, &ain, &aout);
		if (r != ORC_SUCCESS) return r;
		if (exact) aout = avail;
		const uint8_t *t = in + 2 + ain;
		uint32_t want = ((uint32_t)t[0] << 24) | (t[1] << 16) | (t[2] << 8) | t[3];
		if (oracle_adler32(1, out, aout) != want) return ORC_BAD_DATA;
		if (actual_in) *actual_in = 2 + ain + 4;
		if (actual_out && !exact) *actual_out = aout;
		return ORC_SUCCESS;
	}
	/* gzip: lib/gzip_decompress.c:45-134 */
	if (n < 18) return ORC_BAD_DATA;
	if (in[0] != 0x1f || in[1] != 0x8b || in[2] != 8) return ORC_BAD_DATA;
	unsigned flg = in[3];
	pos = 10;
	if (flg & 0xE0) return ORC_BAD_DATA;
	if (flg & 0x04) {
		unsigned xlen = in[pos] | (in[pos + 1] << 8);
		pos += 2;
		if (n - pos < (size_t)xlen + 8) return ORC_BAD_DATA;
		pos += xlen;
	}
	if (flg & 0x08) { while (in[pos++] != 0 && pos != n) {} if (n - pos < 8) return ORC_BAD_DATA; }
	if (flg & 0x10) { while (in[pos++] != 0 && pos != n) {} if (n - pos < 8) return ORC_BAD_DATA; }
	if (flg & 0x02) { pos += 2; if (pos > n || n - pos < 8) return ORC_BAD_DATA; }
	r = inflate_raw(in + pos, n - 8 - pos, (uint8_t *)out, avail, exact, &ain, &aout);
	if (r != ORC_SUCCESS) return r;
	if (exact) aout = avail;
	const uint8_t *t = in + pos + ain;
	uint32_t crc = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24);
	uint32_t isz = t[4] | (t[5] << 8) | (t[6] << 16) | ((uint32_t)t[7] << 24);
	if (oracle_crc32(0, out, aout) != crc) return ORC_BAD_DATA;
	if ((uint32_t)aout != isz) return ORC_BAD_DATA;
	if (actual_in) *actual_in = pos + ain + 8;
	if (actual_out && !exact) *actual_out = aout;
	return ORC_SUCCESS;
}

size_t oracle_compress_bound(int format, size_t n)
{
	size_t blocks = (n + 4999) / 5000;
	if (blocks < 1) blocks = 1;
	return 5 * blocks + n + (format == ORC_GZIP ? 18 : format == ORC_ZLIB ? 6 : 0);
}

size_t oracle_compress_stored(int format, int level, const void *in_, size_t n, void *out_, size_t avail)
{
	const uint8_t *in = (const uint8_t *)in_;
	uint8_t *out = (uint8_t *)out_, *p = out;
	size_t overhead = format == ORC_GZIP ? 18 : format == ORC_ZLIB ? 6 : 0;
	if (overhead && avail <= overhead) return 0;
	size_t nblocks = n ? (n + 65534) / 65535 : 1;
	if (n + 5 * nblocks > avail - overhead) return 0;
	if (format == ORC_GZIP) {
		uint8_t h[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, (uint8_t)(level < 2 ? 4 : level >= 8 ? 2 : 0), 255};
		memcpy(p, h, 10); p += 10;
	} else if (format == ORC_ZLIB) {
		unsigned hint = level < 2 ? 0 : level < 6 ? 1 : level < 8 ? 2 : 3;
		unsigned hdr = (8u << 8) | (7u << 12) | (hint << 6);
		hdr |= 31 - (hdr % 31);
		*p++ = (uint8_t)(hdr >> 8); *p++ = (uint8_t)hdr;
	}
	for (size_t bl = 0; bl < nblocks; bl++) {
		size_t off = bl * 65535, len = n - off > 65535 ? 65535 : n - off;
		*p++ = bl + 1 == nblocks;
		*p++ = (uint8_t)len; *p++ = (uint8_t)(len >> 8); *p++ = (uint8_t)~len; *p++ = (uint8_t)(~len >> 8);
		if (len) memcpy(p, in + off, len);
		p += len;
	}
	if (format == ORC_GZIP) {
		uint32_t c = oracle_crc32(0, in, n), s = (uint32_t)n;
		for (int i = 0; i < 4; i++) *p++ = (uint8_t)(c >> (8 * i));
		for (int i = 0; i < 4; i++) *p++ = (uint8_t)(s >> (8 * i));
	} else if (format == ORC_ZLIB) {
		uint32_t a = oracle_adler32(1, in, n);
		for (int i = 3; i >= 0; i--) *p++ = (uint8_t)(a >> (8 * i));
	}
	return (size_t)(p - out);
}
