/*
 * synth.c -- synthetic chunk generator for bench.py and the tests.
 *
 * Counter-based: chunk i depends only on (SEED, i, class), so any rank can produce
 * any shard.  state = splitmix64(SEED ^ chunk_index), SEED = 0x5EEDDEF1A7E.
 * Classes: 0 T text-like (Zipf(1.1) words from a fixed 4096-word dictionary, the
 * primary throughput corpus, every chunk unique), 1 P (i%123)+(i%1023)
 * (programs/test_trailing_bytes.c:74-75), 2 S stride pattern
 * (programs/test_litrunlen_overflow.c:36-41), 3 R uniform random, 4 Z zeros,
 * 5 M quarter each of T/P/R/Z, 6 = robustness mix (class = chunk_index % 6).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SEED 0x5EEDDEF1A7EULL
#define NWORDS 4096

static uint64_t splitmix64(uint64_t *s)
{
	uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

static char g_words[NWORDS][12];
static uint8_t g_wlen[NWORDS];
static uint32_t g_cdf[NWORDS];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void init_tables(void)
{
	uint64_t s = 1;
	for (int w = 0; w < NWORDS; w++) {
		uint64_t r = splitmix64(&s);
		g_wlen[w] = 2 + (uint8_t)(r % 11);
		for (int k = 0; k < g_wlen[w]; k++) {
			r = splitmix64(&s);
			g_words[w][k] = 'a' + (char)(r % 26);
		}
	}
	double total = 0, acc = 0;
	for (int w = 0; w < NWORDS; w++) total += pow((double)(w + 1), -1.1);
	for (int w = 0; w < NWORDS; w++) {
		acc += pow((double)(w + 1), -1.1);
		double v = acc / total * 4294967296.0;
		g_cdf[w] = v >= 4294967295.0 ? 0xffffffffu : (uint32_t)v;
	}
	g_cdf[NWORDS - 1] = 0xffffffffu;
}

static void fill_text(uint8_t *out, size_t n, uint64_t *s)
{
	size_t pos = 0;
	while (pos < n) {
		uint64_t r = splitmix64(s);
		uint32_t u = (uint32_t)(r >> 32);
		int lo = 0, hi = NWORDS - 1;
		while (lo < hi) {
			int mid = (lo + hi) >> 1;
			if (g_cdf[mid] < u) lo = mid + 1; else hi = mid;
		}
		size_t l = g_wlen[lo];
		if (l > n - pos) l = n - pos;
		memcpy(out + pos, g_words[lo], l);
		pos += l;
		if (pos < n) out[pos++] = (r & 15) ? ' ' : '\n';
	}
}

static void fill_chunk(uint8_t *out, size_t n, uint64_t index, int cls)
{
	uint64_t s = SEED ^ index;
	splitmix64(&s);
	if (cls == 6) cls = (int)(index % 6);
	switch (cls) {
	case 0: fill_text(out, n, &s); break;
	case 1: for (size_t i = 0; i < n; i++) out[i] = (uint8_t)((i % 123) + (i % 1023)); break;
	case 2: { unsigned stride = 1 + (unsigned)(index % 13); for (size_t i = 0; i < n; i++) out[i] = (uint8_t)((stride * i) % 251); break; }
	case 3: for (size_t i = 0; i < n; i += 8) { uint64_t r = splitmix64(&s); memcpy(out + i, &r, n - i < 8 ? n - i : 8); } break;
	case 4: memset(out, 0, n); break;
	default: {
		size_t q = n / 4;
		fill_text(out, q, &s);
		for (size_t i = 0; i < q; i++) out[q + i] = (uint8_t)((i % 123) + (i % 1023));
		for (size_t i = 0; i < q; i += 8) { uint64_t r = splitmix64(&s); memcpy(out + 2 * q + i, &r, q - i < 8 ? q - i : 8); }
		memset(out + 3 * q, 0, n - 3 * q);
	} break;
	}
}

typedef struct { uint8_t *out; size_t chunk, first, lo, hi; int cls; } job_t;

static void *worker(void *a)
{
	job_t *j = (job_t *)a;
	for (size_t i = j->lo; i < j->hi; i++) fill_chunk(j->out + i * j->chunk, j->chunk, j->first + i, j->cls);
	return NULL;
}

void synth_fill(uint8_t *out, size_t chunk_size, size_t first_chunk, size_t n_chunks, int cls, int nthreads)
{
	pthread_once(&g_once, init_tables);
	if (nthreads < 1) nthreads = 1;
	if ((size_t)nthreads > n_chunks) nthreads = n_chunks ? (int)n_chunks : 1;
	pthread_t th[256];
	job_t jobs[256];
	if (nthreads > 256) nthreads = 256;
	for (int t = 0; t < nthreads; t++) {
		jobs[t] = (job_t){out, chunk_size, first_chunk, n_chunks * t / nthreads, n_chunks * (t + 1) / nthreads, cls};
		pthread_create(&th[t], NULL, worker, &jobs[t]);
	}
	for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}
