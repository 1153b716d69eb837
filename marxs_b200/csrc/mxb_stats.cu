// mxb_stats.cu — reductions of the analysis callers as kernels, sm_100a.
//
// sigma_clipped_stats (reference marxs/analysis/analysis.py:9-25 -> astropy.stats.sigma_clipped_stats with its
// defaults: cenfunc = median, stdfunc = std, non-finite values masked, clip to [median - sigma std, median + sigma std],
// stop when a round clips nothing or after maxiters rounds, then mean / median / std of what is left).
//
// Nothing is sorted, compacted or copied: the surviving set is always "finite values inside [lo, hi]" of the ORIGINAL
// column, so a round is a handful of streaming passes over the column (8 B per value each):
//   count + sum -> mean;  sum (x - mean)^2 -> std;  8 radix passes (8 bits of the order-preserving 64-bit key each)
//   that select the exact middle element(s) -> median.
// The whole iteration runs on the device: every kernel reads the state block (window, ranks, `done`) from global
// memory, so all (maxiters + 1) rounds are enqueued at once and the host synchronises only to read the four results.
// Sums are reduced per block and then in block order by one block: bit-reproducible from run to run.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mxb.h"

namespace {

constexpr int kBlocks = 1184;     // 148 SMs x 8 CTAs of 256 threads: full occupancy
constexpr int kUnroll = 4;        // independent 8-byte loads in flight per thread (the passes are latency bound otherwise)
constexpr int kThreads = 256;

struct ClipState {
    double lo, hi;                     // surviving window (inclusive)
    double mean, sigma;
    unsigned long long cnt, prev_cnt;
    unsigned long long key[2];         // radix-select prefixes of the two middle elements
    unsigned long long rank[2];        // their ranks among the values that share the prefix
    int done, round, maxiters, pad;
    unsigned hist[2][256];
    double out[4];                     // mean, median, std, count
    double part_sum[kBlocks];
    unsigned long long part_cnt[kBlocks];
};

__device__ __forceinline__ unsigned long long order_key(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double key_value(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ bool alive(double x, double lo, double hi) {
    return (x >= lo) && (x <= hi) && (fabs(x) <= 1.7976931348623157e308);      // NaN fails every comparison
}

__device__ __forceinline__ double block_sum(double v, double* sm) {
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kThreads / 32; ++w) r += sm[w];
    __syncthreads();
    return r;     // valid in thread 0
}

__global__ void clip_init(ClipState* st, double sigma, int maxiters) {
    if (threadIdx.x == 0) {
        st->lo = -1.7976931348623157e308;
        st->hi = 1.7976931348623157e308;
        st->sigma = sigma;
        st->cnt = 0;
        st->prev_cnt = ~0ULL;
        st->done = 0;
        st->round = 0;
        st->maxiters = maxiters;
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        st->out[0] = st->out[1] = st->out[2] = nan;
        st->out[3] = 0.0;
    }
}

// pass 1: count and sum of the surviving values (per-block partials)
__global__ void __launch_bounds__(kThreads) clip_count_sum(const double* __restrict__ x, long long n, ClipState* st) {
    __shared__ double sm[kThreads / 32];
    if (st->done) return;
    const double lo = st->lo, hi = st->hi;
    double s = 0.0;
    unsigned long long c = 0;
    const long long T = (long long)gridDim.x * kThreads;
    for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n; i += kUnroll * T) {
        double v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) v[u] = (i + u * T < n) ? __ldg(x + i + u * T) : __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (alive(v[u], lo, hi)) {
                s += v[u];
                ++c;
            }
    }
    const double bs = block_sum(s, sm);
    const double bc = block_sum((double)c, sm);      // < 2^53 values per block: exact
    if (threadIdx.x == 0) {
        st->part_sum[blockIdx.x] = bs;
        st->part_cnt[blockIdx.x] = (unsigned long long)bc;
    }
}

// one block: mean, convergence test, ranks of the middle element(s)
// (one warp: lane l adds the partials l, l + 32, ... in order, then a fixed shuffle tree: deterministic)
__device__ __forceinline__ double warp_reduce_partials(const double* p) {
    double s = 0.0;
    for (int b = threadIdx.x; b < kBlocks; b += 32) s += p[b];
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    return __shfl_sync(0xffffffffu, s, 0);
}
__global__ void clip_after_sum(ClipState* st) {
    if (st->done) return;
    const double s = warp_reduce_partials(st->part_sum);
    unsigned long long c = 0;
    for (int b = threadIdx.x; b < kBlocks; b += 32) c += st->part_cnt[b];
    for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if (threadIdx.x) return;
    st->cnt = c;
    if (c == 0 || c == st->prev_cnt) {      // everything clipped, or the last round clipped nothing: out[] is final
        if (c == 0) {
            const double nan = __longlong_as_double(0x7ff8000000000000LL);
            st->out[0] = st->out[1] = st->out[2] = nan;
            st->out[3] = 0.0;
        }
        st->done = 1;
        return;
    }
    st->mean = s / (double)c;
    st->key[0] = st->key[1] = 0ULL;
    st->rank[0] = (c - 1) / 2;
    st->rank[1] = c / 2;
    for (int k = 0; k < 256; ++k) st->hist[0][k] = st->hist[1][k] = 0u;      // (also zeroed by every clip_pick)
}

// pass 2: sum of squared deviations from the mean
__global__ void __launch_bounds__(kThreads) clip_m2(const double* __restrict__ x, long long n, ClipState* st) {
    __shared__ double sm[kThreads / 32];
    if (st->done) return;
    const double lo = st->lo, hi = st->hi, mean = st->mean;
    double s = 0.0;
    const long long T = (long long)gridDim.x * kThreads;
    for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n; i += kUnroll * T) {
        double v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) v[u] = (i + u * T < n) ? __ldg(x + i + u * T) : __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (alive(v[u], lo, hi)) {
                const double d = v[u] - mean;
                s = fma(d, d, s);
            }
    }
    const double bs = block_sum(s, sm);
    if (threadIdx.x == 0) st->part_sum[blockIdx.x] = bs;
}

// radix pass: histogram of the next 8 key bits of the survivors that carry the prefix of middle element 0 / 1
__global__ void __launch_bounds__(kThreads) clip_hist(const double* __restrict__ x, long long n, ClipState* st, int shift) {
    __shared__ unsigned h[2][256];
    if (st->done) return;
    h[0][threadIdx.x] = 0u;
    h[1][threadIdx.x] = 0u;
    __syncthreads();
    const double lo = st->lo, hi = st->hi;
    const unsigned long long ka = st->key[0], kb = st->key[1];
    const bool same = (ka == kb);
    // (shift == 56: no prefix yet; a 64-bit shift by 64 is undefined, so the mask form is used)
    const unsigned long long mask = (shift == 56) ? 0ULL : (~0ULL << (shift + 8));
    const long long n_up = (n + 31) & ~31LL;      // whole warps stay in the loop (match_any needs the full mask)
    const long long T = (long long)gridDim.x * kThreads;
    for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n_up; i += kUnroll * T) {
        double v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) v[u] = (i + u * T < n) ? __ldg(x + i + u * T) : __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const bool ok = alive(v[u], lo, hi);
            const unsigned long long key = order_key(v[u]);
            const unsigned d = (unsigned)(key >> shift) & 255u;
            const bool in_a = ok && ((key & mask) == ka);
            const bool in_b = ok && !same && ((key & mask) == kb);
            // lanes of a warp that count the same bin are combined: a narrow distribution puts the whole column into
            // one or two bins of the leading bytes
            const unsigned tag = in_a ? d : (in_b ? 256u + d : 512u);
            const unsigned peers = __match_any_sync(0xffffffffu, tag);
            if (tag < 512u && (int)(threadIdx.x & 31) == __ffs(peers) - 1)
                atomicAdd(&h[tag >> 8][tag & 255u], (unsigned)__popc(peers));
        }
    }
    __syncthreads();
    if (h[0][threadIdx.x]) atomicAdd(&st->hist[0][threadIdx.x], h[0][threadIdx.x]);
    if (h[1][threadIdx.x]) atomicAdd(&st->hist[1][threadIdx.x], h[1][threadIdx.x]);
}

// one warp: descend one byte for both middle elements (lane l owns bins 8l .. 8l+7; a warp scan finds the owner)
__global__ void clip_pick(ClipState* st, int shift) {
    if (st->done) return;
    const int lane = threadIdx.x;
    const bool same = (st->key[0] == st->key[1]);
    for (int m = 0; m < 2; ++m) {
        const unsigned* h = st->hist[(m == 1 && !same) ? 1 : 0];
        unsigned v[8];
        unsigned long long tot = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            v[k] = h[lane * 8 + k];
            tot += v[k];
        }
        unsigned long long inc = tot;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        const unsigned long long exc = inc - tot, r = st->rank[m];
        const unsigned ball = __ballot_sync(0xffffffffu, (r >= exc) && (r < inc));
        const int owner = ball ? __ffs(ball) - 1 : 31;
        __syncwarp();
        if (lane == owner) {
            unsigned long long cum = exc;
            int d = lane * 8;
            for (int k = 0; k < 8; ++k) {
                d = lane * 8 + k;
                if (r < cum + v[k] || d == 255) break;
                cum += v[k];
            }
            st->rank[m] = r - cum;
            st->key[m] |= (unsigned long long)d << shift;
        }
        __syncwarp();
    }
    for (int k = lane; k < 256; k += 32) st->hist[0][k] = st->hist[1][k] = 0u;
}

// one block: results of the round, next window
__global__ void clip_after_round(ClipState* st) {
    if (st->done) return;
    const double m2 = warp_reduce_partials(st->part_sum);
    if (threadIdx.x) return;
    const double std = sqrt(m2 / (double)st->cnt);
    const double med = 0.5 * (key_value(st->key[0]) + key_value(st->key[1]));
    st->out[0] = st->mean;
    st->out[1] = med;
    st->out[2] = std;
    st->out[3] = (double)st->cnt;
    st->prev_cnt = st->cnt;
    if (st->round >= st->maxiters) {
        st->done = 1;
        return;
    }
    st->round += 1;
    st->lo = fmax(st->lo, med - st->sigma * std);
    st->hi = fmin(st->hi, med + st->sigma * std);
}

__global__ void clip_copy_out(const ClipState* st, double* out) {
    if (threadIdx.x < 4) out[threadIdx.x] = st->out[threadIdx.x];
}

}  // namespace

extern "C" {

size_t mxb_sigma_clip_workspace(void) { return sizeof(ClipState); }

int mxb_sigma_clip_stats(const double* x, int64_t n, double sigma, int maxiters, double* out4, void* workspace, void* stream) {
    if (!out4 || !workspace || n < 0 || (n > 0 && !x) || maxiters < 0 || !(sigma >= 0.0)) return MXB_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    ClipState* st = (ClipState*)workspace;
    clip_init<<<1, 32, 0, s>>>(st, sigma, maxiters);
    for (int r = 0; r <= maxiters && n > 0; ++r) {
        clip_count_sum<<<kBlocks, kThreads, 0, s>>>(x, (long long)n, st);
        clip_after_sum<<<1, 32, 0, s>>>(st);
        clip_m2<<<kBlocks, kThreads, 0, s>>>(x, (long long)n, st);
        for (int shift = 56; shift >= 0; shift -= 8) {
            clip_hist<<<kBlocks, kThreads, 0, s>>>(x, (long long)n, st, shift);
            clip_pick<<<1, 32, 0, s>>>(st, shift);
        }
        clip_after_round<<<1, 32, 0, s>>>(st);
    }
    clip_copy_out<<<1, 32, 0, s>>>(st, out4);
    return cudaGetLastError() == cudaSuccess ? MXB_OK : MXB_ECUDA;
}

}  // extern "C"
