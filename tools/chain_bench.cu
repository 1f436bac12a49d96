// Dependent-chain micro-benchmark for the acquisition kernel's level / window-sum walks (one warp, as there):
// cycles per step of  a = fma (a, c, t_k)  (a) from registers only, (b) with a lane-0 128-bit record every four steps,
// (c) with the inputs of the next 32 steps fetched by eight 128-bit broadcast loads as well, (d) = (c) with two FFMAs a step.
// Measured on B200: 4.41 / 4.53 / 6.57 / 5.47 cycles a step (the kernel's own loops: ~7.5).  Requesting the loads two blocks ahead
// (two register sets) was tried in the kernel and made it slower (219 registers, 6.1 against 5.7 ns per sample).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/chain_bench tools/chain_bench.cu && tools/_build/chain_bench
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k (float *out, long long *cyc, int variant, int nblk) {
	__shared__ __align__ (16) float s_t [1024 + 64], s_r [1024 + 64];
	for (int i = threadIdx. x; i < 1088; i += 32) { s_t [i] = 1e-6f * (float) (i & 15); s_r [i] = 0.f; }
	__syncwarp ();
	const int lane = threadIdx. x;
	float a = 0.3f, ap = a, app = a;
	const float c_hi = 0.99999f, c_lo = 1.3580322644202169613e-08f;
	const long long t0 = clock64 ();
	for (int rep = 0; rep < nblk; rep ++) {
		float4 tn [8];
#pragma unroll
		for (int q = 0; q < 8; q ++) tn [q] = reinterpret_cast<const float4 *> (s_t) [q];
		for (int i0 = 0; i0 < 1024; i0 += 32) {
			float tc [32];
#pragma unroll
			for (int q = 0; q < 8; q ++) { tc [4 * q] = tn [q]. x; tc [4 * q + 1] = tn [q]. y; tc [4 * q + 2] = tn [q]. z; tc [4 * q + 3] = tn [q]. w; if (variant >= 2) tn [q] = reinterpret_cast<const float4 *> (s_t + i0 + 32) [q]; }
#pragma unroll
			for (int q = 0; q < 32; q += 4) {
				float r [4];
#pragma unroll
				for (int j = 0; j < 4; j ++) {
					if (variant == 3) { const float u = __fmaf_rn (app, c_lo, tc [q + j]); app = ap; ap = a; a = __fmaf_rn (a, c_hi, u); }
					else a = __fmaf_rn (a, c_hi, tc [q + j]);
					r [j] = a;
				}
				if (variant >= 1 && lane == 0) *reinterpret_cast<float4 *> (&s_r [i0 + q]) = make_float4 (r [0], r [1], r [2], r [3]);
			}
		}
	}
	const long long t1 = clock64 ();
	if (lane == 0) { out [0] = a + s_r [5]; cyc [0] = t1 - t0; }
}
int main () {
	float *o; long long *c, h;
	cudaMalloc (&o, 4); cudaMalloc (&c, 8);
	const char *name [4] = { "registers only", "+ lane-0 STS.128 per 4 steps", "+ LDS.128 prefetch of the next block", "+ second FFMA (lagged term)" };
	for (int v = 0; v < 4; v ++) {
		k<<<1, 32>>> (o, c, v, 64);
		k<<<1, 32>>> (o, c, v, 64);
		cudaMemcpy (&h, c, 8, cudaMemcpyDeviceToHost);
		printf ("variant %d (%s): %.2f cycles per step\n", v, name [v], (double) h / (64.0 * 1024.0));
	}
	return 0;
}
