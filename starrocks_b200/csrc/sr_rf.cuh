// sr_rf.cuh -- runtime filters: what HashJoinBuildOperator::set_finishing publishes to the scans of
// the probe side (be/src/exec/pipeline/hashjoin/hash_join_build_operator.cpp:100-215):
//   * MinMaxRuntimeFilter  (be/src/runtime/runtime_filter.h:584-): [min, max] of the build keys + has_null;
//   * SimdBlockFilter      (runtime_filter.h:79-240, runtime_filter.cpp:26-35,114-123): split-block bloom filter, one
//     32-byte bucket of 8 words per key: bucket = hash & mask, word k gets bit (uint32(hash >> log_buckets) * SALT[k]) >> 27;
//     hash = phmap_mix<8>(std::hash<T>(value)) (runtime_filter.h:1270-1276, base/phmap/phmap_utils.h:86-95).
// The directory this file builds is byte-identical to the reference's `_directory` (tests compare it with the CPU
// restatement, which is pinned by the reference's own SimdBlockFilter tests), so filters can cross between GPU and CPU BEs.
// A bucket is exactly one 32-byte DRAM sector: a membership test costs one sector read.
#pragma once

#include "sr_host.cuh"

namespace srd {

struct RfDev {
    const uint32_t* dir; // nullptr: no bloom part
    unsigned long long dir_mask;
    long long min_value, max_value;
    int32_t log_buckets;
    int32_t has_null;
    int32_t value_id; // column the scan tests (scan integration only)
    int32_t in_count; // > 0: runtime IN filter -- the sorted distinct build keys; membership is EXACT and replaces the bloom test
    const long long* in_values;
};

__device__ __forceinline__ unsigned long long rf_value_hash(long long v) {
    const unsigned long long a = (unsigned long long)v, k = 0xde5fb9d2630458e9ull;
    return __umul64hi(a, k) + a * k;
}
__device__ __forceinline__ uint32_t rf_salt(int i) {
    constexpr uint32_t SALT[8] = {0x47b6137bu, 0x44974d91u, 0x8824ad5bu, 0xa2b7289du, 0x705495c7u, 0x2df1424bu, 0x9efc4947u, 0x5c6bfb31u};
    return SALT[i];
}
// membership test of a non-NULL value: range first (free), then the 8 bits of its bucket
// runtime IN filter (HashJoiner::_create_runtime_in_filters, hash_joiner.cpp:563-609: at most
// max_pushdown_conditions_per_column = 1024 build rows): binary search in the sorted key list (<= 8 KB, L1 / L2 resident)
__device__ __forceinline__ bool rf_in_list(const RfDev& r, long long v) {
    int lo = 0, hi = r.in_count;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(r.in_values + mid) < v) lo = mid + 1; else hi = mid;
    }
    return lo < r.in_count && __ldg(r.in_values + lo) == v;
}
__device__ __forceinline__ bool rf_test(const RfDev& r, long long v) {
    if (v < r.min_value || v > r.max_value) return false;
    if (r.in_count > 0) return rf_in_list(r, v);
    if (r.dir == nullptr) return true;
    const unsigned long long h = rf_value_hash(v);
    const uint32_t key = (uint32_t)(h >> r.log_buckets);
    uint32_t w[8];
    ldg_nc_u32x8(r.dir + 8 * (h & r.dir_mask), w); // the bucket is one 32-byte sector: one 256-bit load
    uint32_t miss = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) miss |= ~w[i] & (1u << ((key * rf_salt(i)) >> 27));
    return miss == 0;
}

// R rows at once: all bucket reads are issued before any is tested (2 x LDG.128 per row in flight), which is what the
// scan needs -- the directory usually sits in L2 and the test is latency-, not bandwidth-bound.
// active: bit r = row r is to be tested; returns the bits of the rows that pass.
template <int R>
__device__ __forceinline__ uint32_t rf_test_rows(const RfDev& rf, const long long (&v)[R], uint32_t active) {
    uint32_t in = 0;
#pragma unroll
    for (int r = 0; r < R; r++) in |= (((active >> r) & 1u) && v[r] >= rf.min_value && v[r] <= rf.max_value ? 1u : 0u) << r;
    if (rf.in_count > 0) {
        uint32_t pass = 0;
#pragma unroll
        for (int r = 0; r < R; r++)
            if ((in >> r) & 1u) pass |= (rf_in_list(rf, v[r]) ? 1u : 0u) << r;
        return pass;
    }
    if (rf.dir == nullptr || in == 0) return in;
    uint32_t w[R][8];
    uint32_t key[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const unsigned long long h = rf_value_hash(v[r]);
        key[r] = (uint32_t)(h >> rf.log_buckets);
#pragma unroll
        for (int i = 0; i < 8; i++) w[r][i] = 0;
        if ((in >> r) & 1u) ldg_nc_u32x8(rf.dir + 8 * (h & rf.dir_mask), w[r]); // one 256-bit load per bucket (L1 tag rate bound)
    }
    uint32_t pass = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
        uint32_t miss = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) miss |= ~w[r][i] & (1u << ((key[r] * rf_salt(i)) >> 27));
        pass |= (miss == 0 ? 1u : 0u) << r;
    }
    return pass & in;
}

// stats: [0] min, [1] max, [2] inserted, [3] has_null
__device__ __forceinline__ void rf_insert_one(uint32_t* dir, unsigned long long dir_mask, int log_buckets, long long v) {
    const unsigned long long h = rf_value_hash(v);
    const uint32_t key = (uint32_t)(h >> log_buckets);
    uint32_t* b = dir + 8 * (h & dir_mask);
#pragma unroll
    for (int i = 0; i < 8; i++) atomicOr(b + i, 1u << ((key * rf_salt(i)) >> 27));
}

// min/max/count/has_null + bloom insert over one column of a chunk (any integer-class type)
__global__ void __launch_bounds__(256) k_rf_insert_col(DCol col, int64_t n, uint32_t* __restrict__ dir, unsigned long long dir_mask, int log_buckets,
                                                        int insert_nulls, long long* __restrict__ stats) {
    long long mn = 0x7fffffffffffffffll, mx = (long long)0x8000000000000000ll, cnt = 0;
    int nul = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (col.nulls && col.nulls[i]) {
            nul |= insert_nulls;
            continue;
        }
        const long long v = load_int(col.data, col.type, i);
        mn = min(mn, v);
        mx = max(mx, v);
        cnt++;
        if (dir) rf_insert_one(dir, dir_mask, log_buckets, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(SR_FULL_MASK, mn, o));
        mx = max(mx, __shfl_xor_sync(SR_FULL_MASK, mx, o));
        nul |= __shfl_xor_sync(SR_FULL_MASK, nul, o);
    }
    cnt = warp_sum(cnt);
    if (lane_id() == 0) {
        if (cnt) {
            atomicMin(&stats[0], mn);
            atomicMax(&stats[1], mx);
            atomicAdd((unsigned long long*)&stats[2], (unsigned long long)cnt);
        }
        if (nul) atomicOr((int*)&stats[3], 1);
    }
}

// RuntimeFilter::evaluate over one column: selection[i] = has_null for NULL rows, else rf_test(value)
__global__ void __launch_bounds__(256) k_rf_evaluate(const __grid_constant__ RfDev r, DCol col, int64_t n, uint8_t* __restrict__ sel, int merge_and) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (merge_and && sel[i] == 0) continue;
        bool pass;
        if (col.nulls && col.nulls[i])
            pass = r.has_null != 0;
        else
            pass = rf_test(r, load_int(col.data, col.type, i));
        sel[i] = pass ? 1 : 0;
    }
}

__global__ void k_rf_merge_stats(long long* __restrict__ stats, long long mn, long long mx, long long cnt, int has_null) {
    stats[0] = min(stats[0], mn);
    stats[1] = max(stats[1], mx);
    stats[2] += cnt;
    stats[3] |= has_null;
}

__global__ void __launch_bounds__(256) k_or_u32(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] |= src[i];
}

} // namespace srd

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct sr_rf {
    sr_ctx* ctx = nullptr;
    int32_t key_type = 0;
    int32_t log_num_buckets = 0;
    uint64_t dir_mask = 0;
    DevBuf dir, stats;
    Staged staged;
    // runtime IN filter: the distinct non-NULL keys while at most SR_RF_IN_FILTER_ROW_LIMIT rows were inserted
    bool in_enabled = false, in_dirty = false;
    int64_t in_rows = 0;
    std::vector<long long> in_host; // sorted, distinct
    DevBuf in_dev;
    // host copy of the stats, refreshed lazily
    bool stats_valid = false;
    long long hstats[4] = {0, 0, 0, 0};
    size_t dir_bytes() const { return log_num_buckets ? ((size_t)32 << log_num_buckets) : 0; }
};

static int32_t rf_init(sr_rf* rf, sr_ctx* ctx, int32_t key_type, int64_t expected_rows, int32_t with_bloom) {
    if (srd::type_width(key_type) == 0 || srd::type_width(key_type) > 8 || srd::is_float_class(key_type))
        return sr_fail(ctx, SR_ERR_NOT_SUPPORTED, "runtime filter key type %d (integer-class keys of at most 8 bytes)", key_type);
    rf->ctx = ctx;
    rf->key_type = key_type;
    SR_TRY(rf->stats.reserve(ctx, 32));
    const long long init[4] = {0x7fffffffffffffffll, (long long)0x8000000000000000ll, 0, 0};
    SR_CUDA(ctx, cudaMemcpyAsync(rf->stats.p, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    SR_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // `init` is a stack buffer
    // runtime_in_filter_row_limit (hash_joiner.h:270); a pure MinMaxRuntimeFilter (with_bloom = 0) stays a range test
    rf->in_enabled = with_bloom && expected_rows <= SR_RF_IN_FILTER_ROW_LIMIT;
    if (with_bloom) { // SimdBlockFilter::init
        const uint64_t nums = (uint64_t)std::max<int64_t>(1, expected_rows);
        int log_heap_space = 0;
        while ((1ull << log_heap_space) < nums) log_heap_space++; // ceil(log2(nums))
        rf->log_num_buckets = std::max(1, log_heap_space - 5);
        rf->dir_mask = (1ull << std::min(63, rf->log_num_buckets)) - 1;
        SR_TRY(rf->dir.reserve(ctx, rf->dir_bytes()));
        SR_CUDA(ctx, cudaMemsetAsync(rf->dir.p, 0, rf->dir_bytes(), ctx->stream));
    }
    return SR_OK;
}

static int32_t rf_read_stats(sr_rf* rf) {
    if (rf->stats_valid) return SR_OK;
    sr_ctx* ctx = rf->ctx;
    SR_CUDA(ctx, cudaMemcpyAsync(rf->hstats, rf->stats.p, sizeof(rf->hstats), cudaMemcpyDeviceToHost, ctx->stream));
    SR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    rf->stats_valid = true;
    return SR_OK;
}

static int32_t rf_device_desc(sr_rf* rf, srd::RfDev* d) {
    SR_TRY(rf_read_stats(rf));
    d->dir = rf->log_num_buckets ? rf->dir.as<uint32_t>() : nullptr;
    d->dir_mask = rf->dir_mask;
    d->min_value = rf->hstats[0];
    d->max_value = rf->hstats[1];
    d->log_buckets = rf->log_num_buckets;
    d->has_null = rf->hstats[3] != 0;
    d->value_id = -1;
    d->in_count = 0;
    d->in_values = nullptr;
    if (rf->in_enabled && !rf->in_host.empty()) {
        sr_ctx* ctx = rf->ctx;
        if (rf->in_dirty) {
            SR_TRY(rf->in_dev.reserve(ctx, sizeof(long long) * rf->in_host.size()));
            SR_CUDA(ctx, cudaMemcpyAsync(rf->in_dev.p, rf->in_host.data(), sizeof(long long) * rf->in_host.size(), cudaMemcpyHostToDevice, ctx->stream));
            SR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            rf->in_dirty = false;
        }
        d->in_count = (int32_t)rf->in_host.size();
        d->in_values = rf->in_dev.as<long long>();
    }
    return SR_OK;
}
