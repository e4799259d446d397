// cuda_shim.h -- TEST HARNESS ONLY.  A minimal stand-in for the CUDA runtime and the device
// intrinsics so that psiclass_b200/csrc/*.cu{,h} can be compiled by g++ into a serial, one-lane
// (Coop width 1) program.  It exists so the kernel logic can be debugged in a container that has
// no GPU; it is built only by tests/shim/build_shim.sh into tests/shim/_build/, is never loaded by
// the product (psiclass_b200/ loads libpsc.so built by nvcc and fails loudly without a CUDA
// device) and no benchmark number ever comes from it.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <math.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)

struct int4 { int x, y, z, w; };
static inline int4 make_int4(int x, int y, int z, int w) { int4 r = {x, y, z, w}; return r; }
struct int2 { int x, y; };
static inline int2 make_int2(int x, int y) { int2 r = {x, y}; return r; }
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
static inline const char *cudaGetErrorString(cudaError_t) { return "shim"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 8; return 0; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = 0; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = 0; return 0; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = 0; return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return 0; }
struct cudaDeviceProp { int multiProcessorCount; size_t totalGlobalMem; char name[64]; };
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { p->multiProcessorCount = 1; p->totalGlobalMem = (size_t)8 << 30; strcpy(p->name, "cpu-shim"); return 0; }
static inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = (size_t)8 << 30; *t = (size_t)8 << 30; return 0; }

enum cudaLimit { cudaLimitStackSize };
static inline cudaError_t cudaDeviceSetLimit(cudaLimit, size_t) { return 0; }
static inline long long clock64() { return 0; }

// serial kernel launch: one "block" of one "thread"
#define PSC_LAUNCH(kernel, grid, block, stream, ...)                                              \
	do {                                                                                            \
		gridDim = dim3(1); blockDim = dim3(1); blockIdx = dim3(0); threadIdx = dim3(0);              \
		kernel(__VA_ARGS__);                                                                         \
	} while (0)
