/* TEST INFRASTRUCTURE ONLY (oracle/).  Host stand-in for the CUDA runtime so that the reference's
 * own solver classes (/root/reference/src, compiled where they lie, never copied) build with g++ as
 * plain single-threaded host C++.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs may use anything built from this.  Recipe: oracle/ref/build_ref.sh. */
#ifndef GPSAT_ORACLE_CUDA_STUB_H
#define GPSAT_ORACLE_CUDA_STUB_H
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <cstdint>
#include <cstddef>
#include <cmath>
#include <climits>
#include <ctime>
#include <cassert>
#include <new>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __shared__
#define __constant__

struct gpsat_stub_dim3 { unsigned x, y, z; };
static const gpsat_stub_dim3 threadIdx = {0, 0, 0};
static const gpsat_stub_dim3 blockIdx  = {0, 0, 0};
static const gpsat_stub_dim3 blockDim  = {1, 1, 1};
static const gpsat_stub_dim3 gridDim   = {1, 1, 1};

typedef int cudaError;
typedef int cudaError_t;
typedef void *cudaEvent_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost,
                      cudaMemcpyDeviceToDevice, cudaMemcpyDefault };

template <class T> static inline cudaError cudaMalloc(T **p, size_t n)
{ *p = (T *)std::malloc(n ? n : 1); return *p ? 0 : 2; }
template <class T> static inline cudaError cudaMallocPitch(T **p, size_t *pitch, size_t w, size_t h)
{ *pitch = w; *p = (T *)std::malloc(w * h ? w * h : 1); return *p ? 0 : 2; }
static inline cudaError cudaFree(void *p) { std::free(p); return 0; }
static inline cudaError cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) std::memcpy(d, s, n); return 0; }
static inline cudaError cudaMemset(void *d, int v, size_t n) { std::memset(d, v, n); return 0; }
static inline cudaError cudaDeviceReset() { return 0; }
static inline cudaError cudaDeviceSynchronize() { return 0; }
static inline const char *cudaGetErrorString(cudaError) { return "stub"; }
static inline cudaError cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return 0; }
static inline cudaError cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return 0; }
static inline cudaError cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
static inline cudaError cudaEventDestroy(cudaEvent_t) { return 0; }

/* single host thread: the atomics degenerate to their sequential meaning */
static inline unsigned atomicInc(unsigned *a, unsigned lim) { unsigned o = *a; *a = (o >= lim) ? 0 : o + 1; return o; }
static inline int atomicExch(int *a, int v) { int o = *a; *a = v; return o; }
static inline unsigned atomicAdd(unsigned *a, unsigned v) { unsigned o = *a; *a += v; return o; }
static inline int atomicAdd(int *a, int v) { int o = *a; *a += v; return o; }
static inline long long clock64() { return (long long)std::clock(); }
#endif
