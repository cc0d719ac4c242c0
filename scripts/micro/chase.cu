// Microbenchmark: how many divergent, dependent 32-byte loads per second can a B200 sustain from an
// L2-resident array?  This is the access pattern of the stackless BVH walk (one 256-bit node fetch per lane
// per step, the next address depends on the loaded data).  Build: nvcc -arch=sm_100a -O3 -o chase chase.cu
//   ./chase [MB=36] [blocksPerSM=10] [activeLanes=32] [steps=512] [alu=0]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128) chase(const float4* __restrict__ nodes, int n, int steps, int activeLanes, int alu, unsigned* out) {
	const int lane = threadIdx.x & 31;
	unsigned idx = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u % (unsigned) n;
	float acc = 0.0f;
	if (lane < activeLanes) {
		for (int s = 0; s < steps; s++) {
			float4 lo, hi;
			asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
				: "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
				: "l"(nodes + 2 * (size_t) idx));
			idx = __float_as_uint(hi.w);
			float t = lo.x;
			for (int a = 0; a < alu; a++) t = t * hi.x + lo.y;      // dependent filler work, like the box test
			acc += t;
		}
	}
	if (acc == 12345.678f) out[0] = idx;
	if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = idx;
}

int main(int argc, char** argv) {
	const int mb = argc > 1 ? atoi(argv[1]) : 36;
	const int bps = argc > 2 ? atoi(argv[2]) : 10;
	const int lanes = argc > 3 ? atoi(argv[3]) : 32;
	const int steps = argc > 4 ? atoi(argv[4]) : 512;
	const int alu = argc > 5 ? atoi(argv[5]) : 0;
	const int n = mb * 1024 * 1024 / 32;
	std::vector<unsigned> perm(n);
	for (int i = 0; i < n; i++) perm[i] = i;
	std::mt19937 rng(1);
	std::shuffle(perm.begin(), perm.end(), rng);
	std::vector<float> h((size_t) n * 8, 0.5f);
	for (int i = 0; i < n; i++) { unsigned nx = perm[(i + 1) % n]; memcpy(&h[(size_t) perm[i] * 8 + 7], &nx, 4); }
	float4* d; unsigned* out;
	cudaMalloc(&d, (size_t) n * 32); cudaMalloc(&out, 8);
	cudaMemcpy(d, h.data(), (size_t) n * 32, cudaMemcpyHostToDevice);
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	const int grid = p.multiProcessorCount * bps;
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	float best = 1e9f;
	for (int rep = 0; rep < 4; rep++) {
		cudaEventRecord(a);
		chase<<<grid, 128>>>(d, n, steps, lanes, alu, out);
		cudaEventRecord(b); cudaEventSynchronize(b);
		float ms; cudaEventElapsedTime(&ms, a, b);
		if (rep > 0 && ms < best) best = ms;
	}
	const double loads = (double) grid * 4 * lanes * steps;
	int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
	printf("array %3d MB  blocks/SM %2d  active lanes %2d  alu %3d: %7.3f ms  %7.1f G lane-loads/s  %.3f per SM-cycle (at %.2f GHz)\n",
		mb, bps, lanes, alu, best, loads / best / 1e6, loads / (best * 1e-3) / p.multiProcessorCount / (clk * 1e3), clk * 1e-6);
	return 0;
}
