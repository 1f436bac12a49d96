// f32x2_bench.cu -- issue-rate micro-benchmark for the packed FP32 instructions of sm_100a (FFMA2 / FADD2 / FMUL2) against
// their scalar forms, all operands in per-thread registers: 8 independent chains per thread, 8 CTAs of 256 threads per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/f32x2_bench tools/f32x2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
template <int KIND> __global__ void __launch_bounds__ (256) k (float2 *out, const float2 *in) {
	__shared__ float2 sm [2048];
	float2 x [8];
#pragma unroll
	for (int i = 0; i < 8; i ++) x [i] = in [threadIdx. x + 256 * i];
	float2 A = in [threadIdx. x + 2048], B = in [threadIdx. x + 2304];
	for (int i = threadIdx. x; i < 2048; i += 256) sm [i] = in [i];
	__syncthreads ();
	int acc [8];
#pragma unroll
	for (int i = 0; i < 8; i ++) acc [i] = threadIdx. x + i;
	const int m1 = __float_as_int (A. x) | 0xff00, m2 = __float_as_int (B. y) | 5;
#pragma unroll 1
	for (int it = 0; it < ITER; it ++) {
#pragma unroll
		for (int i = 0; i < 8; i ++) {
			if (KIND == 0) { x [i]. x = fmaf (x [i]. x, A. x, B. x); x [i]. y = fmaf (x [i]. y, A. y, B. y); }        // 2 FFMA
			if (KIND == 1) x [i] = __ffma2_rn (x [i], A, B);                                                         // 1 FFMA2
			if (KIND == 2) { x [i]. x = x [i]. x + B. x; x [i]. y = x [i]. y + B. y; }                               // 2 FADD
			if (KIND == 3) x [i] = __fadd2_rn (x [i], B);                                                            // 1 FADD2
			if (KIND == 4) { x [i]. x = x [i]. x * A. x; x [i]. y = x [i]. y * A. y; }                               // 2 FMUL
			if (KIND == 5) x [i] = __fmul2_rn (x [i], A);                                                            // 1 FMUL2
			if (KIND == 6) x [i] = __ffma2_rn (make_float2 (x [i]. y, x [i]. x), A, B);                              // FFMA2, swapped data operand
			if (KIND == 7) x [i] = __ffma2_rn (x [i], make_float2 (A. x, A. x), B);                                  // FFMA2, broadcast scalar
			if (KIND == 8) { x [i] = __ffma2_rn (x [i], A, B); acc [i] = (acc [i] & m1) ^ m2; }                     // FFMA2 + independent LOP3
			if (KIND == 9) { x [i]. x = fmaf (x [i]. x, A. x, B. x); x [i]. y = fmaf (x [i]. y, A. y, B. y); acc [i] = (acc [i] & m1) ^ m2; }
			if (KIND == 10) { x [i] = __ffma2_rn (x [i], A, sm [(threadIdx. x + 32 * i + it) & 2047]); }             // FFMA2 + LDS.64
			if (KIND == 11) { const float2 s = sm [(threadIdx. x + 32 * i + it) & 2047]; x [i]. x = fmaf (x [i]. x, A. x, s. x); x [i]. y = fmaf (x [i]. y, A. y, s. y); }
			if (KIND == 12) { x [i] = __fadd2_rn (x [i], make_float2 (-x [(i + 1) & 7]. x, -x [(i + 1) & 7]. y)); }  // FADD2 reg-reg with negation
			if (KIND == 13) { x [i]. x -= x [(i + 1) & 7]. x; x [i]. y -= x [(i + 1) & 7]. y; }
		}
	}
	int at = 0;
#pragma unroll
	for (int i = 0; i < 8; i ++) at += acc [i];
	float2 s = make_float2 ((float) at, 0.f);
#pragma unroll
	for (int i = 0; i < 8; i ++) { s. x += x [i]. x; s. y += x [i]. y; }
	out [blockIdx. x * 256 + threadIdx. x] = s;
}
template <int KIND> void run (const char *name, int sms, float2 *out, const float2 *in, int instr_per_elem) {
	cudaEvent_t e0, e1; cudaEventCreate (&e0); cudaEventCreate (&e1);
	const int grid = sms * 8;
	k<KIND><<<grid, 256>>> (out, in);
	cudaEventRecord (e0);
	for (int r = 0; r < 5; r ++) k<KIND><<<grid, 256>>> (out, in);
	cudaEventRecord (e1); cudaEventSynchronize (e1);
	float ms; cudaEventElapsedTime (&ms, e0, e1);
	const double elems = 5.0 * grid * 8.0 /* warps */ * ITER * 8;      // warp-level chain updates (2 fp32 lanes-ops x 32 each)
	const double per_clk = elems / (ms * 1e-3) / sms / 1.965e9;
	printf ("%-34s %8.3f ms  %6.3f updates/SM/clk = %6.1f fp32 lane-ops/SM/clk, %5.2f warp-instr/SM/clk\n", name, ms / 5, per_clk, per_clk * 64, per_clk * instr_per_elem);
}
int main () {
	cudaDeviceProp p; cudaGetDeviceProperties (&p, 0);
	float2 *out, *in; cudaMalloc (&out, (size_t) p. multiProcessorCount * 8 * 256 * sizeof (float2));
	cudaMalloc (&in, 4096 * sizeof (float2)); cudaMemset (in, 0, 4096 * sizeof (float2));
	const int n = p. multiProcessorCount;
	printf ("%s, %d SMs\n", p. name, n);
	run<0> ("2 x FFMA", n, out, in, 2);
	run<1> ("FFMA2", n, out, in, 1);
	run<2> ("2 x FADD", n, out, in, 2);
	run<3> ("FADD2", n, out, in, 1);
	run<4> ("2 x FMUL", n, out, in, 2);
	run<5> ("FMUL2", n, out, in, 1);
	run<6> ("FFMA2 swapped data operand", n, out, in, 1);
	run<7> ("FFMA2 broadcast scalar operand", n, out, in, 1);
	run<8> ("FFMA2 + LOP3", n, out, in, 2);
	run<9> ("2 x FFMA + LOP3", n, out, in, 3);
	run<10> ("FFMA2 + LDS.64 (+addr)", n, out, in, 2);
	run<11> ("2 x FFMA + LDS.64 (+addr)", n, out, in, 3);
	run<12> ("FADD2 reg - reg", n, out, in, 1);
	run<13> ("2 x FADD reg - reg", n, out, in, 2);
	return 0;
}
