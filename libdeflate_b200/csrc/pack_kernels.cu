// pack_kernels.cu -- gather the chunks of a batch into one dense buffer, on the device.
//
// Why: a compressed batch sits in compress_bound()-sized slots (the only size a caller can
// allocate up front, ref: libdeflate.h:99-116), ~3.4x larger than what was produced.  Before the
// bytes cross PCIe (host forms) or NVLink (the multi-GPU gather), they are packed back to back:
// chunk i goes to dense + offsets[i], offsets = exclusive prefix sums of the sizes rounded up to
// 16 bytes (so every chunk start stays 16-byte aligned for vector loads; <= 15 B of padding each).
// Algorithmic bytes: sum(sizes) read + written once.
#include "ldb_common.cuh"

// offsets[0..n] = exclusive prefix sums of align16(sizes[i]); one CTA, tiles of 1024
__global__ void __launch_bounds__(1024)
ldb_pack_offsets_kernel(const size_t *sizes, u64 *offsets, size_t n)
{
	__shared__ u64 wsum[32];
	__shared__ u64 carry_s;
	const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) carry_s = 0;
	__syncthreads();
	for (size_t base = 0; base < n; base += 1024) {
		size_t i = base + tid;
		u64 v = i < n ? ((u64)sizes[i] + 15) & ~(u64)15 : 0;
		u64 incl = v;
		for (int o = 1; o < 32; o <<= 1) {
			u64 t = __shfl_up_sync(LDB_FULL_MASK, incl, o);
			if (lane >= (u32)o) incl += t;
		}
		if (lane == 31) wsum[warp] = incl;
		__syncthreads();
		u64 before = carry_s;
		for (u32 w = 0; w < warp; w++) before += wsum[w];
		if (i < n) offsets[i] = before + incl - v;
		__syncthreads();
		if (tid == 1023) carry_s = before + incl;
		__syncthreads();
	}
	if (tid == 0) offsets[n] = carry_s;
}

// one CTA per chunk (grid-stride): 16-byte rows when the source is 16-byte aligned, else bytes
__global__ void __launch_bounds__(256)
ldb_pack_copy_kernel(const void *const *ptrs, const size_t *sizes, const u64 *offsets, u8 *dense, size_t dense_avail, size_t n)
{
	for (size_t c = blockIdx.x; c < n; c += gridDim.x) {
		const u8 *src = (const u8 *)ptrs[c];
		const size_t len = sizes[c];
		const u64 off = offsets[c];
		if (!src || off + len > dense_avail) continue;	// the caller sees offsets[n] > dense_avail
		u8 *dst = dense + off;
		if ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
			const size_t rows = len >> 4;
			for (size_t r = threadIdx.x; r < rows; r += blockDim.x) ((uint4 *)dst)[r] = ((const uint4 *)src)[r];
			for (size_t i = (rows << 4) + threadIdx.x; i < len; i += blockDim.x) dst[i] = src[i];
		} else {
			for (size_t i = threadIdx.x; i < len; i += blockDim.x) dst[i] = src[i];
		}
	}
}

int ldb_launch_pack(const void *const *d_ptrs, const size_t *d_sizes, size_t n, void *d_dense, size_t dense_avail,
		    u64 *d_offsets, const ldb_launch_cfg &cfg, void *stream)
{
	if (n == 0) return 0;
	LDB_LAUNCH(ldb_pack_offsets_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, d_sizes, d_offsets, n);
	LDB_CUDA_CHECK_RET(cudaGetLastError());
	size_t blocks = (size_t)cfg.num_sms * 8;
	if (blocks > n) blocks = n;
	LDB_LAUNCH(ldb_pack_copy_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, d_ptrs, d_sizes, d_offsets,
		   (u8 *)d_dense, dense_avail, n);
	LDB_CUDA_CHECK_RET(cudaGetLastError());
	return 0;
}
