// probe: which hardware warp slots share a scheduler (SMSP)?  Two warps running a dependent-free FFMA loop take
// twice as long when they sit on the same scheduler.  One CTA of 16 warps on one SM; a mask picks the active pair.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned mask, long long *out, float a, float b) {
	unsigned wid;
	asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
	const int w = threadIdx.x >> 5;
	__shared__ long long t0s, t1s;
	__syncthreads();
	long long t0 = clock64();
	if ((mask >> w) & 1u) {
		float r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
		for (int i = 0; i < 20000; ++i) {
			r0 = fmaf(r0, a, b); r1 = fmaf(r1, a, b); r2 = fmaf(r2, a, b); r3 = fmaf(r3, a, b);
			r4 = fmaf(r4, a, b); r5 = fmaf(r5, a, b); r6 = fmaf(r6, a, b); r7 = fmaf(r7, a, b);
		}
		if (r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7 == 1.2345f) out[63] = 1;
	}
	__syncthreads();
	long long t1 = clock64();
	if (threadIdx.x == 0) out[0] = t1 - t0;
	if ((threadIdx.x & 31) == 0) out[1 + w] = wid;
}
int main() {
	long long *d, h[64];
	cudaMalloc(&d, 64 * sizeof(long long));
	for (int other = 0; other < 16; ++other) {
		unsigned mask = 1u | (1u << other);
		k<<<1, 512>>>(mask, d, 0.999f, 0.001f);
		cudaMemcpy(h, d, 64 * sizeof(long long), cudaMemcpyDeviceToHost);
		printf("warps {0,%2d} slots {%lld,%lld}: %lld cycles\n", other, h[1], h[1 + other], h[0]);
	}
	return 0;
}
