// dabgpu_resample.cu -- airspyHandler's sample-rate conversion (airspy-handler.cpp:138-148 tables, 342-370
// data_available): int16 I,Q at `in_rate` samples/s -> complex floats at 2 048 000 samples/s, linear interpolation in
// blocks of 1 ms.  Block b reads input samples [b R, b R + R], R = in_rate / 1000 (the last one is shared with block
// b + 1, exactly the reference's convBuffer[0] = convBuffer[convBufferSize]) and writes 2048 output samples; blocks
// are independent, so the whole recording converts in one launch and the result can go straight into
// dabgpu_decode_cf32_dev.  The products and the sum are rounded separately (no FMA contraction): bit-exact with the
// reference's float arithmetic.
#include <math.h>
#include "dabgpu_internal.h"

struct ResampleMap { int16_t base [2048]; float ratio [2048]; };

__global__ void __launch_bounds__ (256) resample_i16_kernel (const short2 *in, int R, long long nblocks, const ResampleMap *M, float2 *out) {
	const long long b = blockIdx. x;
	if (b >= nblocks) return;
	const short2 *src = in + b * R;
	float2 *dst = out + b * 2048;
	for (int j = threadIdx. x; j < 2048; j += blockDim. x) {
		const int k = M -> base [j];
		const float r = M -> ratio [j], q = __fsub_rn (1.0f, r);
		const short2 a = src [k + 1], c = src [k];
		const float ax = (float) a. x / 2048.0f, ay = (float) a. y / 2048.0f, cx = (float) c. x / 2048.0f, cy = (float) c. y / 2048.0f;
		dst [j] = make_float2 (__fadd_rn (__fmul_rn (ax, r), __fmul_rn (cx, q)), __fadd_rn (__fmul_rn (ay, r), __fmul_rn (cy, q)));
	}
}

static int resample_map (dabgpu *h, int32_t in_rate, const ResampleMap **d_map) {
	ResampleMap m;
	const float inVal = (float) (in_rate / 1000);
	for (int i = 0; i < 2048; i ++) {                        // airspy-handler.cpp:141-146
		m. base [i] = (int16_t) (int) floor (i * (inVal / 2048.0));
		m. ratio [i] = (float) (i * (inVal / 2048.0) - m. base [i]);
	}
	void *d = nullptr;
	int rc = dab_device_table (h, (9ll << 44) | (long long) in_rate, &m, sizeof (m), &d);
	if (rc) return rc;
	*d_map = (const ResampleMap *) d;
	return DABGPU_OK;
}

extern "C" int dabgpu_resample_i16_dev (dabgpu_t *h, const int16_t *d_iq, size_t n_in, int32_t in_rate, float *d_out, size_t *n_out, size_t *consumed) {
	if (!h || !n_out || !consumed || in_rate < 1000 || in_rate % 1000 != 0 || in_rate / 1000 > 32000 || (n_in > 0 && (!d_iq || !d_out)))
		return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_resample_i16: bad argument (in_rate must be a multiple of 1000)");
	CUDA_TRY (h, cudaSetDevice (h -> device));
	const int R = in_rate / 1000;
	const long long nblocks = n_in > 0 ? (long long) ((n_in - 1) / R) : 0;
	*n_out = (size_t) nblocks * 2048; *consumed = (size_t) nblocks * R;
	if (nblocks == 0) return DABGPU_OK;
	const ResampleMap *M;
	int rc = resample_map (h, in_rate, &M);
	if (rc) return rc;
	resample_i16_kernel<<<(unsigned) nblocks, 256, 0, h -> stream>>> ((const short2 *) d_iq, R, nblocks, M, (float2 *) d_out);
	h -> launches ++;
	CUDA_TRY (h, cudaGetLastError ());
	return DABGPU_OK;
}

extern "C" int dabgpu_resample_i16 (dabgpu_t *h, const int16_t *iq, size_t n_in, int32_t in_rate, float *out, size_t *n_out, size_t *consumed) {
	if (!h || !n_out || !consumed || (n_in > 0 && (!iq || !out))) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_resample_i16: bad argument");
	CUDA_TRY (h, cudaSetDevice (h -> device));
	const size_t ib = n_in * 2 * sizeof (int16_t);
	const int R = in_rate >= 1000 ? in_rate / 1000 : 1;
	const size_t ob = (n_in > 0 ? (n_in - 1) / R : 0) * 2048 * sizeof (float2);
	CUDA_TRY (h, h -> d_in. ensure (ib + 16)); CUDA_TRY (h, h -> d_out. ensure (ob + 16));
	CUDA_TRY (h, h -> h_in. ensure (ib + 16));
	if (ib) { memcpy (h -> h_in. p, iq, ib); CUDA_TRY (h, cudaMemcpyAsync (h -> d_in. p, h -> h_in. p, ib, cudaMemcpyHostToDevice, h -> stream)); }
	int rc = dabgpu_resample_i16_dev (h, (const int16_t *) h -> d_in. p, n_in, in_rate, (float *) h -> d_out. p, n_out, consumed);
	if (rc) return rc;
	if (*n_out) {
		CUDA_TRY (h, cudaMemcpyAsync (out, h -> d_out. p, *n_out * sizeof (float2), cudaMemcpyDeviceToHost, h -> stream));
	}
	CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
	return DABGPU_OK;
}
