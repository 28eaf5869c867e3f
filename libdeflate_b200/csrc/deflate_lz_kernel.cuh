// deflate_lz_kernel.cuh -- PLACEHOLDER until the LZ77 + Huffman kernel lands:
// every level currently emits stored blocks (valid, ratio 1.0).
#pragma once
size_t ldb_deflate_scratch_bytes(const ldb_launch_cfg &cfg) { (void)cfg; return 4096; }
static int ldb_launch_deflate_lz(const ldb_deflate_args &a, const ldb_launch_cfg &cfg, void *stream)
{
	size_t blocks = a.n < (size_t)cfg.num_sms * 8 ? a.n : (size_t)cfg.num_sms * 8;
	LDB_LAUNCH(ldb_deflate_stored_kernel, dim3((unsigned)blocks), dim3(DEF_THREADS), 0, (cudaStream_t)stream, a);
	LDB_CUDA_CHECK_RET(cudaGetLastError());
	return 0;
}
