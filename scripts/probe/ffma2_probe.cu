// ffma2_probe.cu -- FFMA2 (fma.rn.f32x2) issue rate on sm_100a by operand form and by warps per scheduler.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 bc(float a) { u64 r; asm volatile("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(a)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE, int NACC>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b) {
	u64 acc[NACC];
	float t[16];
	for (int i = 0; i < 16; ++i) t[i] = a + i * 1e-3f;
	for (int i = 0; i < NACC; ++i) acc[i] = pk(threadIdx.x + i, threadIdx.x - i);
	const u64 xa = pk(a, b), xb = pk(b, a);
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int u = 0; u < 16; ++u) {
#pragma unroll
			for (int i = 0; i < NACC; ++i) {
				if (MODE == 0) acc[i] = fma2(acc[i], xa, xb);                 // packed operands
				if (MODE == 1) acc[i] = fma2(bc(t[(u + i) & 15]), xa, acc[i]); // broadcast multiplier, like the sweep kernel
				if (MODE == 2) acc[i] = fma2(bc(t[(u + i) & 15]), acc[i], bc(t[u])); // two broadcast operands
			}
		}
	}
	float s = 0;
	for (int i = 0; i < NACC; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i])); s += lo + hi; }
	if (s == 123.456f) out[0] = s;
}

template <int MODE, int NACC> void run(const char *name, int threads, int blocks_per_sm) {
	float *d; cudaMalloc(&d, 4);
	int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int iters = 4096, blocks = sms * blocks_per_sm;
	double best = 0;
	for (int rep = 0; rep < 6; ++rep) {
		cudaEventRecord(e0);
		k<MODE, NACC><<<blocks, threads>>>(d, iters, 0.999f, 0.001f);
		cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		double tf = 2.0 * 2 * NACC * 16 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
		if (rep && tf > best) best = tf;
	}
	printf("%-44s threads %3d blocks/SM %d: %.2f TFLOP/s\n", name, threads, blocks_per_sm, best);
}
int main() {
	run<0, 8>("packed operands, 8 chains", 256, 8);
	run<1, 8>("broadcast multiplier, 8 chains", 256, 8);
	run<2, 8>("two broadcast operands, 8 chains", 256, 8);
	run<1, 16>("broadcast multiplier, 16 chains, 1 warp/SMSP", 128, 1);
	run<1, 16>("broadcast multiplier, 16 chains, 2 warps/SMSP", 256, 1);
	run<0, 16>("packed operands, 16 chains, 1 warp/SMSP", 128, 1);
	run<1, 4>("broadcast multiplier, 4 chains, 1 warp/SMSP", 128, 1);
	run<1, 2>("broadcast multiplier, 2 chains, 1 warp/SMSP", 128, 1);
	run<1, 1>("broadcast multiplier, 1 chain, 1 warp/SMSP", 128, 1);
	return 0;
}
