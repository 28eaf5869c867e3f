// inflate_resolve.cu -- second half of batched DEFLATE decompression for sm_100a: the LZ77
// side.  (First half: inflate_kernel.cu, which turns every stream into tokens.)
//
// What it computes: for every chunk whose stream decoded, the bytes the reference's decode
// loop writes (ref: lib/decompress_template.h:381-430 literals, :590-670 match copies with
// forward byte-by-byte overlap semantics) -- from the chunk's token stream: packed literal
// bytes + 4-byte {literal run, length, offset} records (format: ldb_common.cuh).
//
// B200 mapping -- one 256-thread CTA per chunk, three CTAs per SM:
//   * the output lives in a 64 KiB shared-memory RING (index = position mod 65536, shifted so
//     that 16-byte rows of the ring are 16-byte rows of the destination): the 32 KiB LZ77
//     window never leaves the SM and never touches L2/HBM; finished bytes leave as coalesced
//     16-byte rows, the only global stores of the kernel;
//   * records are taken 256 at a time (one per thread); a CTA-wide prefix sum of
//     {literals, literals + length} gives every thread its literal source and its destination;
//   * literals are placed at once; a match whose source lies wholly before the block is copied
//     at once (all of those are independent of each other); the others set "pending" bits over
//     their destination range in a bitmap and are resolved in rounds: a match is ready when no
//     byte of its source is pending -- exact dependency tracking, so the number of rounds is the
//     depth of the dependency chains inside one block, not the number of matches;
//   * a copy moves up to 16 bytes per step (aligned word loads, funnel shifts, word stores);
//     an overlapping match (offset < length) doubles its effective offset after every period, so
//     a run of 258 equal bytes takes log2 steps, not 258.
//
// Algorithmic HBM bytes per chunk: actual_out written once (+ the token stream read once, which
// is the price of the two-kernel split; see inflate_kernel.cu).
#include "ldb_common.cuh"

#define RES_THREADS 256
#define RES_WARPS   (RES_THREADS / 32)
#define RES_RING    65536u
#define RES_MASK    (RES_RING - 1)
#define RES_SPAN    16384u	// most output bytes one block of records may cover
#define RES_FLUSH   8192u	// finished bytes that trigger a write-out
#define RES_LIT_FAST 8u		// literal runs up to this are placed by the owning thread

// shared memory layout
#define RES_SM_RING  0
#define RES_SM_PEND  (RES_SM_RING + RES_RING)			// u32[RES_SPAN / 32]
#define RES_SM_LIST  (RES_SM_PEND + RES_SPAN / 8)		// u32[3 * RES_THREADS]: {dst, src, n}
#define RES_SM_WSUM  (RES_SM_LIST + 12 * RES_THREADS)		// u32[2 * RES_WARPS]
#define RES_SM_MISC  (RES_SM_WSUM + 8 * RES_WARPS)		// u32[8]
#define RES_SM_BYTES (RES_SM_MISC + 32)

// the scratch slot of a chunk: what the decoder can emit is bounded both by the output room
// (a record stands for >= 3 bytes or for up to 2^31 literals) and by the input (a literal takes
// >= 1 bit, a match >= 2 bits).  Same formula on host and device.
__host__ __device__ static inline size_t res_tok_cap(size_t in_nbytes, size_t out_avail)
{
	if (out_avail > 0xfffffff0u) out_avail = 0xfffffff0u;
	if (in_nbytes > 0xfffffff0u) in_nbytes = 0xfffffff0u;
	size_t a = out_avail + out_avail / 3 + 64;
	size_t b = 25 * (in_nbytes + 16) + 64;
	size_t c = a < b ? a : b;
	return (c + 15) & ~(size_t)15;
}
size_t ldb_inflate_tok_cap(size_t in_nbytes, size_t out_avail) { return res_tok_cap(in_nbytes, out_avail); }

// ---- slot sizes -> exclusive prefix sums (one CTA; n + 1 outputs) ---------------------------
__global__ void __launch_bounds__(1024)
ldb_inflate_caps_kernel(const size_t *in_nbytes, const size_t *out_avail, u64 *tok_off, size_t n)
{
	__shared__ u64 wsum[32];
	__shared__ u64 carry_s;
	const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) carry_s = 0;
	__syncthreads();
	for (size_t base = 0; base < n; base += 1024) {
		size_t i = base + tid;
		u64 v = i < n ? (u64)res_tok_cap(in_nbytes[i], out_avail[i]) : 0;
		u64 incl = v;
		for (int o = 1; o < 32; o <<= 1) {
			u64 t = __shfl_up_sync(LDB_FULL_MASK, incl, o);
			if (lane >= (u32)o) incl += t;
		}
		if (lane == 31) wsum[warp] = incl;
		__syncthreads();
		u64 before = carry_s;
		for (u32 w = 0; w < warp; w++) before += wsum[w];
		if (i < n) tok_off[i] = before + incl - v;
		__syncthreads();
		if (tid == 1023) carry_s = before + incl;
		__syncthreads();
	}
	if (tid == 0) tok_off[n] = carry_s;
}

int ldb_launch_inflate_caps(const size_t *d_in_nbytes, const size_t *d_out_avail, u64 *d_tok_off, size_t n, void *stream)
{
	LDB_LAUNCH(ldb_inflate_caps_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, d_in_nbytes, d_out_avail, d_tok_off, n);
	LDB_CUDA_CHECK_RET(cudaGetLastError());
	return 0;
}

// ---- pending bitmap (bit i = output byte P + i is not final yet) --------------------------------
__device__ __forceinline__ void res_bits_set(u32 *bm, u32 lo, u32 hi)	// [lo, hi), hi > lo
{
	u32 wl = lo >> 5, wh = (hi - 1) >> 5;
	u32 ml = 0xffffffffu << (lo & 31), mh = 0xffffffffu >> (31 - ((hi - 1) & 31));
	if (wl == wh) {
		atomicOr(&bm[wl], ml & mh);
	} else {
		atomicOr(&bm[wl], ml);
		for (u32 w = wl + 1; w < wh; w++) atomicOr(&bm[w], 0xffffffffu);
		atomicOr(&bm[wh], mh);
	}
}
__device__ __forceinline__ void res_bits_clear(u32 *bm, u32 lo, u32 hi)
{
	u32 wl = lo >> 5, wh = (hi - 1) >> 5;
	u32 ml = 0xffffffffu << (lo & 31), mh = 0xffffffffu >> (31 - ((hi - 1) & 31));
	if (wl == wh) {
		atomicAnd(&bm[wl], ~(ml & mh));
	} else {
		atomicAnd(&bm[wl], ~ml);
		for (u32 w = wl + 1; w < wh; w++) atomicAnd(&bm[w], 0u);
		atomicAnd(&bm[wh], ~mh);
	}
}
__device__ __forceinline__ bool res_bits_any(const u32 *bm, u32 lo, u32 hi)
{
	const volatile u32 *b = bm;
	u32 wl = lo >> 5, wh = (hi - 1) >> 5;
	u32 ml = 0xffffffffu << (lo & 31), mh = 0xffffffffu >> (31 - ((hi - 1) & 31));
	if (wl == wh) return (b[wl] & ml & mh) != 0;
	u32 acc = b[wl] & ml;
	for (u32 w = wl + 1; w < wh; w++) acc |= b[w];
	acc |= b[wh] & mh;
	return acc != 0;
}

// ---- copies inside the ring --------------------------------------------------------------------
// m <= 16 bytes from ring position qs to qd; the two ranges do not overlap.
__device__ __forceinline__ void res_copy_piece(u8 *ring, u32 qd, u32 qs, u32 m)
{
	const u32 *r32 = (const u32 *)ring;
	const u32 sa = qs & ~3u, ssh = 8 * (qs & 3);
	const u32 need = (qs & 3) + m;		// source bytes counted from the aligned start
	u32 w0 = r32[(sa & RES_MASK) >> 2], w1 = 0, w2 = 0, w3 = 0, w4 = 0;
	if (need > 4) w1 = r32[((sa + 4) & RES_MASK) >> 2];
	if (need > 8) w2 = r32[((sa + 8) & RES_MASK) >> 2];
	if (need > 12) w3 = r32[((sa + 12) & RES_MASK) >> 2];
	if (need > 16) w4 = r32[((sa + 16) & RES_MASK) >> 2];
	u32 v0 = __funnelshift_r(w0, w1, ssh), v1 = __funnelshift_r(w1, w2, ssh);
	u32 v2 = __funnelshift_r(w2, w3, ssh), v3 = __funnelshift_r(w3, w4, ssh);
	// head: single bytes up to the next word boundary of the destination
	u32 h = (4 - (qd & 3)) & 3;
	if (h > m) h = m;
	if (h > 0) ring[qd & RES_MASK] = (u8)v0;
	if (h > 1) ring[(qd + 1) & RES_MASK] = (u8)(v0 >> 8);
	if (h > 2) ring[(qd + 2) & RES_MASK] = (u8)(v0 >> 16);
	// body: whole words, re-aligned to the destination
	const u32 hsh = 8 * h;
	u32 u0 = __funnelshift_r(v0, v1, hsh), u1 = __funnelshift_r(v1, v2, hsh);
	u32 u2 = __funnelshift_r(v2, v3, hsh), u3 = v3 >> hsh;
	u32 *d32 = (u32 *)ring;
	const u32 qa = qd + h, rem = m - h, nw = rem >> 2;
	if (nw > 0) d32[(qa & RES_MASK) >> 2] = u0;
	if (nw > 1) d32[((qa + 4) & RES_MASK) >> 2] = u1;
	if (nw > 2) d32[((qa + 8) & RES_MASK) >> 2] = u2;
	if (nw > 3) d32[((qa + 12) & RES_MASK) >> 2] = u3;
	// tail: the last rem & 3 bytes
	const u32 t = rem & 3;
	if (t) {
		u32 tw = nw == 0 ? u0 : (nw == 1 ? u1 : (nw == 2 ? u2 : u3));
		const u32 qt = qa + 4 * nw;
		ring[qt & RES_MASK] = (u8)tw;
		if (t > 1) ring[(qt + 1) & RES_MASK] = (u8)(tw >> 8);
		if (t > 2) ring[(qt + 2) & RES_MASK] = (u8)(tw >> 16);
	}
}

// n bytes to qd from 'off' bytes back, with the byte-by-byte forward semantics of a DEFLATE
// match (the source may run into the destination).  After each full period the effective
// offset doubles: the bytes just written repeat the pattern.
__device__ __forceinline__ void res_copy_match(u8 *ring, u32 qd, u32 off, u32 n)
{
	u32 eff = off;
	while (n) {
		u32 m = n < 16 ? n : 16;
		if (m > eff) m = eff;
		res_copy_piece(ring, qd, qd - eff, m);
		qd += m;
		n -= m;
		if (eff < 16) eff += eff;
	}
}

// ---- the kernel ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(RES_THREADS, 3)
ldb_inflate_resolve_kernel(ldb_inflate_args a, u32 *work_counter)
{
	LDB_DYN_SMEM(sm);
	u8 *ring = sm + RES_SM_RING;
	u32 *pend = (u32 *)(sm + RES_SM_PEND);
	u32 *list = (u32 *)(sm + RES_SM_LIST);
	u32 *wsum = (u32 *)(sm + RES_SM_WSUM);
	volatile u32 *misc = (volatile u32 *)(sm + RES_SM_MISC);	// 0 chunk, 1 list count, 2 first lits, 3 P', 4 L'
	const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	for (u32 i = tid; i < RES_SPAN / 32; i += RES_THREADS) pend[i] = 0;

	for (;;) {
		__syncthreads();	// the previous chunk's write-out has left the ring
		if (tid == 0) misc[0] = atomicAdd(work_counter, 1u);
		__syncthreads();
		const size_t idx = misc[0];
		if (idx >= a.count) break;
		const size_t c = a.first + idx;
		const u32 n_rec = a.tok_counts[2 * c];
		if (n_rec == 0) continue;
		const u8 *lit = a.tok_base + (a.tok_off[c] - a.tok_origin);
		const u32 *rec_end = (const u32 *)(a.tok_base + (a.tok_off[c + 1] - a.tok_origin));
		u8 *const out = (u8 *)a.out_ptrs[c];
		// shifted coordinates: q = output position + a0, so that q % 16 is the alignment phase of
		// the destination and the ring index q % 65536 keeps 16-byte rows together
		const u32 a0 = (u32)(uintptr_t)out & 15;
		u8 *const gbase = out - a0;
		u32 P = a0;		// where the next block of records starts writing
		u32 L = 0;		// literal bytes consumed
		u32 flushed = a0;	// everything below has been written to 'out'
		u32 r0 = 0;		// first record of the next block
		u32 big_rem = 0;	// what is left of a literal run longer than RES_SPAN

		// write-out of ring bytes [flushed, upto): 16-byte rows, single bytes at ragged ends
		auto write_out = [&](u32 upto) {
			if (upto <= flushed) return;
			u32 lo16 = (flushed + 15) & ~15u, hi16 = upto & ~15u;
			if (lo16 >= hi16) {
				for (u32 q = flushed + tid; q < upto; q += RES_THREADS) gbase[q] = ring[q & RES_MASK];
			} else {
				for (u32 q = flushed + tid; q < lo16; q += RES_THREADS) gbase[q] = ring[q & RES_MASK];
				for (u32 q = lo16 + 16 * tid; q < hi16; q += 16 * RES_THREADS)
					*(uint4 *)(gbase + q) = *(const uint4 *)(ring + (q & RES_MASK));
				for (u32 q = hi16 + tid; q < upto; q += RES_THREADS) gbase[q] = ring[q & RES_MASK];
			}
			flushed = upto;
		};

		while (r0 < n_rec) {
			// ---- records of this block, one per thread ------------------------------------
			const u32 i = r0 + tid;
			const bool valid = i < n_rec;
			u32 r = valid ? rec_end[-1 - (s32)i] : LDB_TOK_PURE_FLAG;
			u32 lits, mlen = 0, off = 0;
			if (r & LDB_TOK_PURE_FLAG) {
				lits = r & 0x7fffffffu;
				if (tid == 0 && big_rem) lits = big_rem;
			} else {
				lits = (r >> 23) & 255;
				mlen = ((r >> 15) & 255) + 3;
				off = (r & 32767) + 1;
			}
			if (tid == 0) { misc[1] = 0; misc[2] = lits; }
			// CTA-wide exclusive prefix sums of {lits, lits + mlen}
			u32 li = lits, ti = lits + mlen;
			for (int o = 1; o < 32; o <<= 1) {
				u32 x = __shfl_up_sync(LDB_FULL_MASK, li, o), y = __shfl_up_sync(LDB_FULL_MASK, ti, o);
				if (lane >= (u32)o) { li += x; ti += y; }
			}
			if (lane == 31) { wsum[2 * warp] = li; wsum[2 * warp + 1] = ti; }
			__syncthreads();
			u32 lbase = 0, tbase = 0;
			for (u32 w = 0; w < warp; w++) { lbase += wsum[2 * w]; tbase += wsum[2 * w + 1]; }
			const u32 lit_src = L + lbase + li - lits;	// first literal of this record
			const u32 q_lit = P + tbase + ti - lits - mlen;	// where its literals go
			const u32 q_m = q_lit + lits;			// where its match goes
			const u32 q_end = q_m + mlen;
			// a block never covers more than RES_SPAN bytes: cut it at the first record that would
			const bool inc = valid && (q_end - P <= RES_SPAN);
			const u32 n_inc = (u32)__syncthreads_count(inc);
			if (n_inc == 0) {
				// the first record is a literal run longer than the span: move one span of it
				const u32 first_lits = misc[2];
				for (u32 k = tid; k < RES_SPAN; k += RES_THREADS) ring[(P + k) & RES_MASK] = __ldg(lit + L + k);
				P += RES_SPAN;
				L += RES_SPAN;
				big_rem = first_lits - RES_SPAN;
				__syncthreads();
				write_out(P & ~15u);
				continue;
			}
			if (inc && tid == n_inc - 1) { misc[3] = q_end; misc[4] = lit_src + lits; }

			// ---- literals, independent matches, pending bits --------------------------------
			bool pending = false;
			if (inc) {
				if (lits <= RES_LIT_FAST) {
#pragma unroll
					for (u32 k = 0; k < RES_LIT_FAST; k++)
						if (k < lits) ring[(q_lit + k) & RES_MASK] = __ldg(lit + lit_src + k);
				} else {
					u32 e = atomicAdd((u32 *)&misc[1], 1u);
					list[3 * e] = q_lit;
					list[3 * e + 1] = lit_src;
					list[3 * e + 2] = lits;
				}
				if (mlen) {
					if (q_m - off + mlen <= P) res_copy_match(ring, q_m, off, mlen);
					else {
						pending = true;
						res_bits_set(pend, q_m - P, q_end - P);
					}
				}
			}
			__syncthreads();
			// long literal runs: one warp per run, 32 bytes per step
			const u32 n_list = misc[1];
			if (n_list) {
				for (u32 e = warp; e < n_list; e += RES_WARPS) {
					const u32 qd = list[3 * e], src = list[3 * e + 1], n = list[3 * e + 2];
					for (u32 k = lane; k < n; k += 32) ring[(qd + k) & RES_MASK] = __ldg(lit + src + k);
				}
			}
			// ---- dependent matches, in rounds ----------------------------------------------
			// needed source bytes: [q_m - off, min(q_m - off + mlen, q_m)); only those at or after P
			// can be pending
			u32 need_lo = 0, need_hi = 0;
			if (pending) {
				u32 s0 = q_m - off, s1 = s0 + mlen;
				if (s1 > q_m) s1 = q_m;
				need_lo = s0 > P ? s0 - P : 0;
				need_hi = s1 > P ? s1 - P : 0;
			}
			while (__syncthreads_or(pending)) {
				const bool ready = pending && !(need_hi > need_lo && res_bits_any(pend, need_lo, need_hi));
				__syncthreads();
				if (ready) {
					res_copy_match(ring, q_m, off, mlen);
					res_bits_clear(pend, q_m - P, q_end - P);
					pending = false;
				}
			}
			P = misc[3];
			L = misc[4];
			r0 += n_inc;
			big_rem = 0;
			if (r0 >= n_rec) write_out(P);
			else if (P - flushed >= RES_FLUSH) write_out(P & ~15u);
			__syncthreads();	// misc[] and the list are rewritten by the next block
		}
	}
}

int ldb_launch_inflate_resolve(const ldb_inflate_args &a, const ldb_launch_cfg &cfg, void *stream)
{
	if (a.count == 0) return 0;
	u32 *d_counter = ldb_inflate_resolve_counter(a, cfg);
	LDB_CUDA_CHECK_RET(cudaFuncSetAttribute(ldb_inflate_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RES_SM_BYTES));
	int per_sm = (cfg.max_smem_optin + 1024) / (RES_SM_BYTES + 1024);
	if (per_sm < 1) per_sm = 1;
	if (per_sm > 3) per_sm = 3;
	size_t blocks = (size_t)cfg.num_sms * per_sm;
	if (blocks > a.count) blocks = a.count;
	LDB_LAUNCH(ldb_inflate_resolve_kernel, dim3((unsigned)blocks), dim3(RES_THREADS), RES_SM_BYTES, (cudaStream_t)stream, a, d_counter);
	LDB_CUDA_CHECK_RET(cudaGetLastError());
	return 0;
}
