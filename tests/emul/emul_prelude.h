// Stepping harness: the product's CUDA sources compiled for the host, kernels run thread by thread.
// TEST INFRASTRUCTURE ONLY (see README.md).  Force-included (-include) in front of every translation unit of
// gcm_b200/csrc by build_emul.py.
#pragma once
#define GCMB_EMUL 1
#include <cuda_runtime.h>  // types only; every runtime call is redirected below

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <functional>

namespace gcmb_emul {
inline thread_local uint3 t_blockIdx, t_threadIdx;
inline thread_local dim3 t_blockDim, t_gridDim;
inline thread_local double acc_v;
inline thread_local long long acc_c;

template<typename F>
void launch(dim3 grid, dim3 block, F body) {
	t_gridDim = grid;
	t_blockDim = block;
	for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
		t_blockIdx = {bx, by, bz};
		acc_v = 0;
		acc_c = 0;
		// threads run from the last to the first so that thread 0 sees complete block sums
		for (long long t = (long long) block.x * block.y * block.z - 1; t >= 0; t--) {
			t_threadIdx = {(unsigned) (t % block.x), (unsigned) ((t / block.x) % block.y), (unsigned) (t / ((long long) block.x * block.y))};
			body();
		}
	}
}

// kernels written as barrier-separated phases: one call per block, the kernel loops over its threads
template<typename F>
void launch_blocks(dim3 grid, dim3 block, F body) {
	t_gridDim = grid;
	t_blockDim = block;
	t_threadIdx = {0, 0, 0};
	for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
		t_blockIdx = {bx, by, bz};
		body();
	}
}

inline cudaError_t e_malloc(void** p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t e_free(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t e_memcpy(void* d, const void* s, size_t n) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t e_memset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t ok() { return cudaSuccess; }
struct Ev { std::chrono::steady_clock::time_point t; };
inline cudaError_t ev_create(cudaEvent_t* e) { *e = reinterpret_cast<cudaEvent_t>(new Ev); return cudaSuccess; }
inline cudaError_t ev_record(cudaEvent_t e) { reinterpret_cast<Ev*>(e)->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t ev_elapsed(float* ms, cudaEvent_t a, cudaEvent_t b) {
	*ms = std::chrono::duration<float, std::milli>(reinterpret_cast<Ev*>(b)->t - reinterpret_cast<Ev*>(a)->t).count();
	return cudaSuccess;
}
inline cudaError_t ev_destroy(cudaEvent_t e) { delete reinterpret_cast<Ev*>(e); return cudaSuccess; }
}  // namespace gcmb_emul

#define blockIdx gcmb_emul::t_blockIdx
#define threadIdx gcmb_emul::t_threadIdx
#define blockDim gcmb_emul::t_blockDim
#define gridDim gcmb_emul::t_gridDim
#define __syncthreads() ((void) 0)
#undef __shared__
#define __shared__ static thread_local
using std::min;
using std::max;

#define cudaGetDeviceCount(p) (*(p) = 1, cudaSuccess)
#define cudaGetLastError() gcmb_emul::ok()
#define cudaSetDevice(d) gcmb_emul::ok()
#define cudaStreamCreateWithFlags(p, f) (*(p) = nullptr, cudaSuccess)
#define cudaStreamSynchronize(s) gcmb_emul::ok()
#define cudaStreamCreateWithPriority(p, f, prio) (*(p) = nullptr, cudaSuccess)
#define cudaDeviceGetStreamPriorityRange(lo, hi) (*(lo) = 0, *(hi) = 0, cudaSuccess)
#define cudaStreamWaitEvent(s, e, f) gcmb_emul::ok()
#define cudaEventCreateWithFlags(p, f) gcmb_emul::ev_create(p)
#define cudaStreamDestroy(s) gcmb_emul::ok()
#define cudaEventCreate(p) gcmb_emul::ev_create(p)
#define cudaEventRecord(e, s) gcmb_emul::ev_record(e)
#define cudaEventSynchronize(e) gcmb_emul::ok()
#define cudaEventElapsedTime(ms, a, b) gcmb_emul::ev_elapsed(ms, a, b)
#define cudaEventDestroy(e) gcmb_emul::ev_destroy(e)
#define cudaMalloc(p, n) gcmb_emul::e_malloc((void**) (p), n)
#define cudaFree(p) gcmb_emul::e_free(p)
#define cudaHostAlloc(p, n, f) gcmb_emul::e_malloc((void**) (p), n)
#define cudaFreeHost(p) gcmb_emul::e_free(p)
#define cudaMemsetAsync(d, v, n, s) gcmb_emul::e_memset(d, v, n)
#define cudaMemcpyAsync(d, s, n, k, st) gcmb_emul::e_memcpy(d, s, n)
#define cudaMemcpy(d, s, n, k) gcmb_emul::e_memcpy(d, s, n)
#define cudaMemset(d, v, n) gcmb_emul::e_memset(d, v, n)
#define cudaGetErrorString(e) "emulated CUDA error"

#define GCMB_GLOBAL static
#define GCMB_DEV static inline
#define GCMB_BOUNDS(n)
#define GCMB_BOUNDS2(n, b)
#define GCMB_LAUNCH(kernel, grid, block, stream, ...) \
	gcmb_emul::launch(dim3(grid), dim3(block), [&]() { kernel(__VA_ARGS__); })
#define GCMB_LAUNCH_COOP(kernel, grid, block, smem, stream, ...) \
	gcmb_emul::launch_blocks(dim3(grid), dim3(block), [&]() { kernel(__VA_ARGS__); })
#define GCMB_BLOCK_THREADS(tid) for (int tid = 0; tid < (int) gcmb_emul::t_blockDim.x; tid++)
#define GCMB_DYN_SMEM_RAW(name)                                                       \
	alignas(128) static thread_local unsigned char name##_storage_[256 * 1024];        \
	unsigned char* name = name##_storage_
#define cudaFuncSetAttribute(k, a, v) cudaSuccess
#define cudaGetDevice(p) (*(p) = 0, cudaSuccess)
#define GCMB_EMUL_BLOCK_SUM(v, c) \
	do { gcmb_emul::acc_v += (v); gcmb_emul::acc_c += (c); (v) = gcmb_emul::acc_v; (c) = gcmb_emul::acc_c; } while (0)

