// probe: which hardware warp slots do the warps of small CTAs get? (B200)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int *out, int spin) {
	unsigned smid, wid;
	asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
	asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
	if ((threadIdx.x & 31) == 0) {
		int i = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
		out[2 * i] = smid;
		out[2 * i + 1] = wid;
	}
	long long t0 = clock64();
	while (clock64() - t0 < spin) {}
}
int main(int argc, char **argv) {
	int threads = argc > 1 ? atoi(argv[1]) : 64, blocks = 148 * 4, regs_dummy = 0;
	int n = blocks * threads / 32;
	int *d, *h = (int *)malloc(n * 2 * sizeof(int));
	cudaMalloc(&d, n * 2 * sizeof(int));
	k<<<blocks, threads>>>(d, 2000000);
	cudaMemcpy(h, d, n * 2 * sizeof(int), cudaMemcpyDeviceToHost);
	// print the CTAs that landed on SM 0 and SM 1
	for (int sm = 0; sm < 2; ++sm) {
		printf("SM %d:", sm);
		for (int i = 0; i < n; ++i)
			if (h[2 * i] == sm) printf(" [cta %d w%d slot %d]", i / (threads / 32), i % (threads / 32), h[2 * i + 1]);
		printf("\n");
	}
	return 0;
}
