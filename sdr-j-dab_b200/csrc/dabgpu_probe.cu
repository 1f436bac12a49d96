// dabgpu_probe.cu -- integer-ALU peak probe.  MEASURED_PEAKS.json holds only HBM and bf16-GEMM peaks; the
// Viterbi group is bound by the integer pipes, so bench.py measures their peak on the same
// GPU in the same run with this dependent-chain-free micro-benchmark: every thread keeps 16 independent
// accumulators and issues add / min (ALU pipe) and mad.lo (FMA pipe) instructions back to back.
#include "dabgpu_internal.h"

template <int KIND>
__global__ void __launch_bounds__ (256) int_peak_kernel (unsigned *sink, unsigned seed, int iters) {
	unsigned r [16];
#pragma unroll
	for (int i = 0; i < 16; i ++) r [i] = seed + threadIdx. x * 16 + i;
	const unsigned b = seed | 1u, c = seed * 3u + 7u;
	for (int it = 0; it < iters; it ++) {
#pragma unroll
		for (int i = 0; i < 16; i ++) {
			if (KIND == 0)      asm volatile ("add.u32 %0, %0, %1;" : "+r" (r [i]) : "r" (b));
			else if (KIND == 1) asm volatile ("min.u32 %0, %0, %1;" : "+r" (r [i]) : "r" (c));
			else if (KIND == 3) r [i] = __vadd2 (r [i], b);
			else if (KIND == 4) r [i] = __vminu2 (r [i], c + i);
			else if (KIND == 5) { bool ph, pl; r [i] = __vibmin_u16x2 (r [i], c + i, &ph, &pl); if (!pl) r [(i + 1) & 15] |= 4u; if (!ph) r [(i + 2) & 15] |= 8u; }
			else if ((i & 1) == 0) asm volatile ("add.u32 %0, %0, %1;" : "+r" (r [i]) : "r" (b));
			else                asm volatile ("mad.lo.u32 %0, %0, %1, %2;" : "+r" (r [i]) : "r" (b), "r" (c));
		}
	}
	unsigned x = 0;
#pragma unroll
	for (int i = 0; i < 16; i ++) x ^= r [i];
	if (x == 0x12345678u) sink [0] = x;        // keeps the chain alive, practically never taken
}

// ops[0] = add-only, ops[1] = min-only, ops[2] = add + mad.lo mix, ops[3] = packed 16x2 add, ops[4] = packed 16x2 min,
// ops[5] = packed min with predicates + 2 predicated ORs (counted as one), each in instructions x lanes per second
extern "C" int dabgpu_int_peak (dabgpu_t *h, double *ops) {
	if (!h || !ops) return DABGPU_ERR_ARG;
	CUDA_TRY (h, cudaSetDevice (h -> device));
	int sms = 0;
	CUDA_TRY (h, cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, h -> device));
	CUDA_TRY (h, h -> d_aux. ensure (256));
	const int iters = 4096, grid = sms * 8;
	cudaEvent_t e0, e1;
	CUDA_TRY (h, cudaEventCreate (&e0)); CUDA_TRY (h, cudaEventCreate (&e1));
	for (int kind = 0; kind < 6; kind ++) {
		float best = 1e30f;
		for (int rep = 0; rep < 4; rep ++) {
			CUDA_TRY (h, cudaEventRecord (e0, h -> stream));
			if (kind == 0) int_peak_kernel<0><<<grid, 256, 0, h -> stream>>> ((unsigned *) h -> d_aux. p, 12345u + rep, iters);
			if (kind == 1) int_peak_kernel<1><<<grid, 256, 0, h -> stream>>> ((unsigned *) h -> d_aux. p, 12345u + rep, iters);
			if (kind == 2) int_peak_kernel<2><<<grid, 256, 0, h -> stream>>> ((unsigned *) h -> d_aux. p, 12345u + rep, iters);
			if (kind == 3) int_peak_kernel<3><<<grid, 256, 0, h -> stream>>> ((unsigned *) h -> d_aux. p, 12345u + rep, iters);
			if (kind == 4) int_peak_kernel<4><<<grid, 256, 0, h -> stream>>> ((unsigned *) h -> d_aux. p, 12345u + rep, iters);
			if (kind == 5) int_peak_kernel<5><<<grid, 256, 0, h -> stream>>> ((unsigned *) h -> d_aux. p, 12345u + rep, iters);
			CUDA_TRY (h, cudaEventRecord (e1, h -> stream));
			CUDA_TRY (h, cudaEventSynchronize (e1));
			float ms = 0;
			CUDA_TRY (h, cudaEventElapsedTime (&ms, e0, e1));
			if (rep > 0 && ms < best) best = ms;
		}
		ops [kind] = (double) grid * 256.0 * 16.0 * iters / (best * 1e-3);
	}
	cudaEventDestroy (e0); cudaEventDestroy (e1);
	return DABGPU_OK;
}
