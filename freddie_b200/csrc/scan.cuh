// scan.cuh -- exclusive scans and fills shared by the cluster-prep and split-stage kernels (templates only: the
// header is included by several translation units).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef long long i64;
typedef unsigned int u32;
typedef unsigned char u8;
typedef unsigned long long u64;

// ---- exclusive scans (three launches: block sums, scan of the sums by one block, apply) ----
#define CP_SCAN_THREADS 256
#define CP_SCAN_ITEMS 8
template <class TIn, class TOut>
__global__ void k_cp_scan_sums(const TIn* __restrict__ in, i64 n, TOut* __restrict__ sums) {
  __shared__ TOut sh[CP_SCAN_THREADS / 32];
  const i64 base = (i64)blockIdx.x * CP_SCAN_THREADS * CP_SCAN_ITEMS;
  TOut v = 0;
  for (int k = 0; k < CP_SCAN_ITEMS; ++k) {
    const i64 i = base + (i64)k * CP_SCAN_THREADS + threadIdx.x;
    if (i < n) v += (TOut)in[i];
  }
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    TOut s = 0;
    for (int k = 0; k < CP_SCAN_THREADS / 32; ++k) s += sh[k];
    sums[blockIdx.x] = s;
  }
}
template <class TOut>
__global__ void k_cp_scan_top(TOut* sums, int nb, TOut* total) {  // one block; nb is small (n / 2048)
  __shared__ TOut carry;
  __shared__ TOut sh[CP_SCAN_THREADS];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < nb; b0 += CP_SCAN_THREADS) {
    const int i = b0 + threadIdx.x;
    const TOut v = i < nb ? sums[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < CP_SCAN_THREADS; o <<= 1) {
      const TOut x = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += x;
      __syncthreads();
    }
    if (i < nb) sums[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += sh[CP_SCAN_THREADS - 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
template <class TIn, class TOut>
__global__ void k_cp_scan_apply(const TIn* __restrict__ in, i64 n, const TOut* __restrict__ sums, TOut* __restrict__ out) {
  // the items of a block in their order: thread-strided chunks of CP_SCAN_THREADS
  __shared__ TOut sh[CP_SCAN_THREADS];
  __shared__ TOut carry;
  const i64 base = (i64)blockIdx.x * CP_SCAN_THREADS * CP_SCAN_ITEMS;
  if (threadIdx.x == 0) carry = sums[blockIdx.x];
  __syncthreads();
  for (int k = 0; k < CP_SCAN_ITEMS; ++k) {
    const i64 i = base + (i64)k * CP_SCAN_THREADS + threadIdx.x;
    const TOut v = i < n ? (TOut)in[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < CP_SCAN_THREADS; o <<= 1) {
      const TOut x = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += x;
      __syncthreads();
    }
    if (i < n) out[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += sh[CP_SCAN_THREADS - 1];
    __syncthreads();
  }
}

template <class T>
__global__ void k_cp_fill(T* p, i64 n, T v) {
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) p[i] = v;
}
