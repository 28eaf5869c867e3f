/* checksum_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h). */
#include "oracle.h"

/* ref: lib/crc32.c:211-219 crc32_slice1 with the table entries of lib/crc32_tables.h
 * regenerated bit-by-bit from the generator 0xEDB88320 (lib/crc32.c:51-57). */
uint32_t oracle_crc32(uint32_t crc, const void *buf, size_t len)
{
	const uint8_t *p = (const uint8_t *)buf;
	if (!p) return 0;			/* lib/crc32.c:259-260 */
	crc = ~crc;				/* lib/crc32.c:261 */
	for (size_t i = 0; i < len; i++) {
		crc ^= p[i];
		for (int k = 0; k < 8; k++)
			crc = (crc >> 1) ^ ((crc & 1) ? 0xEDB88320u : 0);
	}
	return ~crc;
}

/* ref: lib/adler32.c:31 (DIVISOR 65521), :54 (MAX_CHUNK_LEN 5552), :75-103 ADLER32_CHUNK */
uint32_t oracle_adler32(uint32_t adler, const void *buf, size_t len)
{
	const uint8_t *p = (const uint8_t *)buf;
	if (!p) return 1;			/* lib/adler32.c:159-160 */
	uint32_t s1 = adler & 0xffff, s2 = adler >> 16;
	while (len) {
		size_t n = len < 5552 ? len : 5552;
		len -= n;
		while (n--) {
			s1 += *p++;
			s2 += s1;
		}
		s1 %= 65521;
		s2 %= 65521;
	}
	return (s2 << 16) | s1;
}
