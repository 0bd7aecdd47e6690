// npb_alg2.cuh -- argument block and race noise shared by the two conjugate Algorithm 2 sweep kernels (npb_alg2.cu: one step at
// a time; npb_alg2_tile.cu: a tile of steps evaluated ahead, corrected after every move)
#pragma once
#include "npb_alg8_tile4.cuh"

struct A2Args {
	const float *X;        // [N, D]
	const double *X64;     // [N, D] the same rows in double: what the sufficient statistics add and subtract
	const int32_t *order;  // [n_sweeps, N]
	npb_z_t *z;            // [N, C]
	int *counts;           // [C, 32]
	int *kocc, *overflow;
	unsigned long long *st; // [C, 4]
	double *sx, *sxx;      // [C, 32, D], [C, 32, D, D]
	float *mu, *P, *ld;    // [C, 32, D], [C, 32, D, D], [C, 32]
	const float *G;        // [N + 2] the count-dependent constant of the predictive
	const float *lp0;      // [N] prior-predictive log-density of every item (a new cluster's candidate)
	const float *P0;       // [D, D] Lambda_0^-1
	float mu0[NPB_MAX_D];
	float ld0, kappa0, nu0, log2_alpha;
	int N, C, n_sweeps;
	uint32_t sweep0;
	uint64_t seed;
	int tile;              // k_a2_tile: steps evaluated ahead (1 = strictly one step at a time, same results)
};

__device__ __forceinline__ float a2_noise(uint32_t a, uint32_t b, uint32_t c) {
	return neg_lg2_exp1(npb_mix32(npb_mix32(a ^ (b * 0x9E3779B1u)) ^ (c * 0x85EBCA77u)));
}

