// scan.cuh -- exclusive prefix sums (per-track segment counts -> segment offsets; track lengths ->
// shard split points).  Three small kernels: per-tile reduce, scan of the tile sums by one block, and
// tile-local warp-shuffle scan + offset (+ a caller-supplied carry).  out has n+1 entries (out[n] = carry + total).
#pragma once
#include <cuda_runtime.h>

namespace rt {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

// exclusive scan of one value per thread across the block; returns the exclusive prefix, total in *total
template <typename T>
__device__ __forceinline__ T block_excl_scan(T v, T *total) {
    __shared__ T warp_sums[kScanThreads / 32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    T inc = warp_incl_scan(v, lane);
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        T s = lane < kScanThreads / 32 ? warp_sums[lane] : T(0);
        T si = warp_incl_scan(s, lane);
        if (lane < kScanThreads / 32) warp_sums[lane] = si - s;
        if (lane == kScanThreads / 32 - 1) *total = si;
    }
    __syncthreads();
    T r = warp_sums[w] + inc - v;
    __syncthreads();
    return r;
}

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(kScanThreads) k_scan_tile_sums(const Tin *in, Tout *tile_sums, long long n) {
    __shared__ Tout total;
    long long base = blockIdx.x * (long long)kScanTile + threadIdx.x * kScanItems;
    Tout s = 0;
#pragma unroll
    for (int q = 0; q < kScanItems; ++q)
        if (base + q < n) s += (Tout)in[base + q];
    block_excl_scan<Tout>(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: in-place exclusive scan of the tile sums
template <typename Tout>
__global__ void __launch_bounds__(kScanThreads) k_scan_tile_offsets(Tout *tile_sums, long long n_tiles) {
    __shared__ Tout total;
    Tout carry = 0;
    for (long long b = 0; b < n_tiles; b += kScanThreads) {
        long long i = b + threadIdx.x;
        Tout v = i < n_tiles ? tile_sums[i] : Tout(0);
        Tout ex = block_excl_scan<Tout>(v, &total);
        if (i < n_tiles) tile_sums[i] = carry + ex;
        carry += total;
        __syncthreads();
    }
}

// optimistic evaluation: does the batch fit the Segment columns and did the record pool hold every record?  Checked by the
// thread that writes the scan's total (the flag is cleared with the control block at the start of the call).
struct ScanGuard {
    int *cancel;  // nullptr: no guard
    long long base, cap;
    const int *pool_cursor;
    int pool_blocks;
};

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const Tin *in, Tout *out, const Tout *tile_offsets,
                                                             long long n, Tout carry, const ScanGuard guard) {
    __shared__ Tout total;
    long long base = blockIdx.x * (long long)kScanTile + threadIdx.x * kScanItems;
    Tout v[kScanItems];
    Tout s = 0;
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) {
        v[q] = (base + q < n) ? (Tout)in[base + q] : Tout(0);
        s += v[q];
    }
    Tout ex = block_excl_scan<Tout>(s, &total) + tile_offsets[blockIdx.x] + carry;
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) {
        if (base + q < n) out[base + q] = ex;
        ex += v[q];
        if (base + q == n - 1) {
            out[n] = ex;
            if (guard.cancel && ((long long)ex - guard.base > guard.cap || *guard.pool_cursor > guard.pool_blocks)) *guard.cancel = 1;
        }
    }
}

// grid-stride sum (total track length of a shard; only used to size chunks)
__global__ void k_sum_double(const double *x, long long n, double *out) {
    double s = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += x[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

}  // namespace rt
