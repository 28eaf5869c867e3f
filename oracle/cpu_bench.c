/*
 * cpu_bench.c -- TEST/BENCH INFRASTRUCTURE ONLY.
 *
 * pthread harness around the UNMODIFIED reference (oracle/_ref/libdeflate_ref.so):
 * one compressor + one decompressor per thread (legal per libdeflate.h:56-57,
 * 178-179) over disjoint contiguous chunk ranges, mirroring the chunk loop of
 * programs/benchmark.c:443-509.  Used by bench.py for (1) producing the reference's
 * own L6 streams that the decompress benchmark inflates, (2) the cpu_baseline /
 * `--impl reference` timings.  Built as a shared library for ctypes.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "libdeflate.h"

typedef struct {
	int op;			/* 0 compress, 1 decompress, 2 crc32, 3 adler32 */
	int fmt, level;
	const uint8_t *in;	/* compress/checksum: n chunks of chunk_size; decompress: stream slab */
	size_t chunk_size;
	const size_t *in_off;	/* decompress: offsets/sizes of streams */
	const size_t *in_sz;
	uint8_t *out;		/* compress: slab with out_stride per chunk; decompress: n * chunk_size */
	size_t out_stride;
	size_t *out_sz;		/* compress: produced sizes */
	uint32_t *sums;
	size_t lo, hi;
	int failures;
} job_t;

static double now(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void *worker(void *arg)
{
	job_t *j = (job_t *)arg;
	struct libdeflate_compressor *c = NULL;
	struct libdeflate_decompressor *d = NULL;
	if (j->op == 0) c = libdeflate_alloc_compressor(j->level);
	if (j->op == 1) d = libdeflate_alloc_decompressor();
	for (size_t i = j->lo; i < j->hi; i++) {
		if (j->op == 0) {
			const void *src = j->in + i * j->chunk_size;
			void *dst = j->out + i * j->out_stride;
			size_t r;
			if (j->fmt == 2) r = libdeflate_gzip_compress(c, src, j->chunk_size, dst, j->out_stride);
			else if (j->fmt == 1) r = libdeflate_zlib_compress(c, src, j->chunk_size, dst, j->out_stride);
			else r = libdeflate_deflate_compress(c, src, j->chunk_size, dst, j->out_stride);
			j->out_sz[i] = r;
			if (!r) j->failures++;
		} else if (j->op == 1) {
			const void *src = j->in + j->in_off[i];
			void *dst = j->out + i * j->chunk_size;
			size_t aout = 0;
			enum libdeflate_result r;
			if (j->fmt == 2) r = libdeflate_gzip_decompress(d, src, j->in_sz[i], dst, j->chunk_size, &aout);
			else if (j->fmt == 1) r = libdeflate_zlib_decompress(d, src, j->in_sz[i], dst, j->chunk_size, &aout);
			else r = libdeflate_deflate_decompress(d, src, j->in_sz[i], dst, j->chunk_size, &aout);
			if (r != LIBDEFLATE_SUCCESS || aout != j->chunk_size) j->failures++;
		} else if (j->op == 2) {
			j->sums[i] = libdeflate_crc32(0, j->in + i * j->chunk_size, j->chunk_size);
		} else {
			j->sums[i] = libdeflate_adler32(1, j->in + i * j->chunk_size, j->chunk_size);
		}
	}
	if (c) libdeflate_free_compressor(c);
	if (d) libdeflate_free_decompressor(d);
	return NULL;
}

static double run(job_t proto, size_t n, int nthreads, int *failures)
{
	if (nthreads < 1) nthreads = 1;
	if ((size_t)nthreads > n) nthreads = (int)(n ? n : 1);
	pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
	job_t *jobs = (job_t *)malloc(sizeof(job_t) * nthreads);
	double t0 = now();
	for (int t = 0; t < nthreads; t++) {
		jobs[t] = proto;
		jobs[t].lo = n * t / nthreads;
		jobs[t].hi = n * (t + 1) / nthreads;
		jobs[t].failures = 0;
		pthread_create(&th[t], NULL, worker, &jobs[t]);
	}
	int f = 0;
	for (int t = 0; t < nthreads; t++) {
		pthread_join(th[t], NULL);
		f += jobs[t].failures;
	}
	double dt = now() - t0;
	if (failures) *failures = f;
	free(th);
	free(jobs);
	return dt;
}

/* returns seconds; out_sz[i] = compressed size of chunk i (0 = failure) */
double cpub_compress(int fmt, int level, const uint8_t *in, size_t chunk_size, size_t n,
		     uint8_t *out, size_t out_stride, size_t *out_sz, int nthreads, int *failures)
{
	job_t j;
	memset(&j, 0, sizeof(j));
	j.op = 0; j.fmt = fmt; j.level = level; j.in = in; j.chunk_size = chunk_size;
	j.out = out; j.out_stride = out_stride; j.out_sz = out_sz;
	return run(j, n, nthreads, failures);
}

double cpub_decompress(int fmt, const uint8_t *streams, const size_t *off, const size_t *sz, size_t n,
		       uint8_t *out, size_t chunk_size, int nthreads, int *failures)
{
	job_t j;
	memset(&j, 0, sizeof(j));
	j.op = 1; j.fmt = fmt; j.in = streams; j.in_off = off; j.in_sz = sz; j.out = out; j.chunk_size = chunk_size;
	return run(j, n, nthreads, failures);
}

double cpub_checksum(int is_adler, const uint8_t *in, size_t chunk_size, size_t n, uint32_t *sums, int nthreads)
{
	job_t j;
	memset(&j, 0, sizeof(j));
	j.op = is_adler ? 3 : 2; j.in = in; j.chunk_size = chunk_size; j.sums = sums;
	return run(j, n, nthreads, NULL);
}
