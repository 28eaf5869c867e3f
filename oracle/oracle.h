/*
 * oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's algorithms for the hot path, used only as
 * the checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 * The product (libdeflate_b200/) never links, loads or executes anything here.
 *
 * Parity pin: every function below is checked against the UNMODIFIED reference
 * compiled into oracle/_ref/libdeflate_ref.so (tests/test_oracle_pin.py) and
 * against the golden fixtures under tests/golden/.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stddef.h>
#include <stdint.h>

enum { ORC_SUCCESS = 0, ORC_BAD_DATA = 1, ORC_SHORT_OUTPUT = 2, ORC_INSUFFICIENT_SPACE = 3 };
enum { ORC_RAW = 0, ORC_ZLIB = 1, ORC_GZIP = 2 };

/* ref: lib/crc32.c:256-262 (slice-by-1 formulation of crc32_slice1, lib/crc32.c:211-219) */
uint32_t oracle_crc32(uint32_t crc, const void *buf, size_t len);
/* ref: lib/adler32.c:75-119,156-162 */
uint32_t oracle_adler32(uint32_t adler, const void *buf, size_t len);
/* ref: lib/decompress_template.h:44-772 + lib/{gzip,zlib}_decompress.c; exact == (actual_out_ret == NULL) */
int oracle_decompress(int format, const void *in, size_t in_nbytes, void *out, size_t out_avail,
		      int exact, size_t *actual_in, size_t *actual_out);
/* ref: lib/deflate_compress.c:4088-4135 (+6 zlib, +18 gzip) */
size_t oracle_compress_bound(int format, size_t in_nbytes);
/* ref: lib/deflate_compress.c:2393-2443 (level 0 stored blocks) with the wrappers */
size_t oracle_compress_stored(int format, int level, const void *in, size_t in_nbytes, void *out, size_t out_avail);
#endif
