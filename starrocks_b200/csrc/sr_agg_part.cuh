// sr_agg_part.cuh -- radix-partitioned push of the hash aggregate (device side; included by sr_agg.cuh).
//
// Replaces, for large batches, the row-at-a-time lazy_emplace + update of
//   AggHashMap*::compute_agg_states            be/src/exec/aggregate/agg_hash_map.h:303-361
//   Aggregator::compute_batch_agg_states       be/src/exec/aggregator.cpp:1616-1640
// whose GPU transcription (k_agg_push) pays one L2 atomic transaction per row and state word (measured round 1:
// ~60-90 G atomics/s whatever the table size).  Here the batch is first moved next to the table slice it will update:
//
//   k_aggp_hist      one pass over the key columns: rows per BUCKET (bucket = top bits of the row's home slot)
//   k_aggp_prepare   exclusive scan -> bucket bases, write cursors, tile list of the second scatter level
//   k_aggp_scatter   packed RECORDS (key, evaluated function inputs, null mask; W 8-byte words) are counting-sorted by
//                    bucket inside a shared-memory tile, then every bucket's run is copied out contiguously (space
//                    reserved with ONE atomicAdd per (tile, bucket) on the bucket cursor).  Up to 2^9 buckets take one
//                    level; up to 2^15 take two (fan-out <= 2^8 each, so runs stay >= 16 records long)
//   k_aggp_apply_smem one CTA per bucket = per table SLICE: the slice's key / state arrays are loaded into shared
//                    memory (or initialised there when the table is still empty), the bucket's records are applied with
//                    shared-memory atomics (64-bit CAS claim, 32-bit adds with carry), the slice is written back with
//                    coalesced stores.  HBM traffic per row: W*8 bytes x (1 write + 1 read) per level + the table once.
//   k_aggp_apply_l2  fallback when a bucket spans more than one slice (tables beyond 2^15 slices) and for the records of
//                    slices that overflowed: global atomics on the L2-prefetched table range, as in round 1.
#pragma once

namespace srd {

constexpr int AGGP_BLOCK = 512;
constexpr int AGGP_MAX_BITS = 15;       // buckets of one push: the histogram of 2^15 counters lives in shared memory
constexpr int AGGP_ONE_LEVEL_BITS = 9;  // up to 2^9 buckets are scattered in one level
constexpr int AGGP_THREAD_WORDS = 16;   // record words a thread keeps in registers per tile
constexpr int AGGP_MAX_WORDS = 2 + SR_MAX_AGG_FNS + 1;
constexpr int AGGP_MAX_FAN = 512;    // fan-out of one scatter level

enum PartWordKind : int32_t { WK_KEY_LO = 0, WK_KEY_HI = 1, WK_VALUE = 2, WK_NULLS = 3 };

struct PartPlan {
    int32_t bits;         // log2(buckets)
    int32_t bits2;        // bits of the second scatter level (0: one level)
    int32_t bucket_shift; // bucket = (hash & mask) >> bucket_shift
    int32_t words;        // W: 8-byte words per record
    int32_t null_word;    // word holding the null mask (bit f: input of function f is NULL), -1: no input is nullable
    int32_t pad;
    int32_t word_kind[AGGP_MAX_WORDS];
    int32_t word_fn[AGGP_MAX_WORDS];
    int32_t val_word[SR_MAX_AGG_FNS]; // word of function f's input, -1 for COUNT(*)
    // SIMPLE plans (one non-nullable group-by column of <= 8 bytes, every function input a plain non-nullable column):
    // word 0 = the key column's value & simple_key_mask, word w = value id word_vid[w]; no expression interpreter
    int32_t simple;
    int32_t word_vid[AGGP_MAX_WORDS];
    unsigned long long simple_key_mask;
};

__host__ __device__ constexpr int aggp_row_group(int W) { return W <= 2 ? 4 : (W <= 4 ? 2 : 1); } // rows produced together
__host__ __device__ constexpr int aggp_rows_per_thread(int W) {
    return AGGP_THREAD_WORDS / W / aggp_row_group(W) > 0 ? AGGP_THREAD_WORDS / W / aggp_row_group(W) * aggp_row_group(W) : aggp_row_group(W);
}

__device__ __forceinline__ uint32_t aggp_bucket(const AggDev& a, const PartPlan& pl, const HKey& key) {
    return (uint32_t)((hkey_hash(a, key) & a.mask) >> pl.bucket_shift);
}

// ---- histogram --------------------------------------------------------------------------------------------------
constexpr int AGGP_HIST_ROWS = 8; // key loads in flight per thread
__global__ void __launch_bounds__(AGGP_BLOCK, 1) k_aggp_hist(const AggDev* __restrict__ ad, const __grid_constant__ VTab vt, const __grid_constant__ PartPlan pl,
                                                              int64_t row_base, int64_t n, uint32_t* __restrict__ hist) {
    extern __shared__ uint32_t s_h[];
    const AggDev& a = *ad;
    const int P = 1 << pl.bits;
    for (int i = threadIdx.x; i < P; i += AGGP_BLOCK) s_h[i] = 0;
    __syncthreads();
    const int64_t tile = (int64_t)AGGP_BLOCK * AGGP_HIST_ROWS;
    const int64_t ntiles = (n + tile - 1) / tile;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t r0 = t * tile + threadIdx.x;
        HKey key[AGGP_HIST_ROWS];
#pragma unroll
        for (int k = 0; k < AGGP_HIST_ROWS; k++) {
            const int64_t r = r0 + (int64_t)k * AGGP_BLOCK;
            key[k].lo = key[k].hi = SR_AGG_EMPTY;
            if (r < n) {
                ChunkLoader ld{vt, row_base + r};
                agg_pack_key(a, ld, key[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < AGGP_HIST_ROWS; k++) {
            const int64_t r = r0 + (int64_t)k * AGGP_BLOCK;
            // rows whose packed key equals the empty marker live in the table's special slot: the scatter pass applies
            // them directly, they are not staged
            if (r < n && !hkey_is_empty(a, key[k])) atomicAdd(&s_h[aggp_bucket(a, pl, key[k])], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P; i += AGGP_BLOCK) {
        const uint32_t c = s_h[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// ---- bucket bases, cursors, tile list of the second level (one CTA) -------------------------------------------------
__global__ void __launch_bounds__(1024) k_aggp_prepare(const uint32_t* __restrict__ hist, int bits, int bits2, int tile_rows, uint64_t* __restrict__ base,
                                                        unsigned long long* __restrict__ cursor, unsigned long long* __restrict__ cursor1,
                                                        uint32_t* __restrict__ tile_start) {
    __shared__ uint32_t s_scan[1024 / 32 + 1];
    __shared__ uint64_t s_running;
    const int P = 1 << bits;
    if (threadIdx.x == 0) s_running = 0;
    __syncthreads();
    for (int b0 = 0; b0 < P; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const uint32_t v = i < P ? hist[i] : 0;
        uint32_t tot;
        const uint32_t ex = block_excl_scan<1024>(v, s_scan, &tot);
        const uint64_t run = s_running;
        if (i < P) {
            base[i] = run + ex;
            cursor[i] = run + ex;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_running = run + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) base[P] = s_running;
    if (bits2 == 0) return;
    __syncthreads(); // base[] written by this block is visible to it after the barrier
    const int F1 = 1 << (bits - bits2);
    if (threadIdx.x == 0) s_running = 0;
    __syncthreads();
    for (int b0 = 0; b0 < F1; b0 += 1024) { // F1 <= 1024: one iteration
        const int i = b0 + threadIdx.x;
        uint32_t tiles = 0;
        if (i < F1) {
            const uint64_t lo = base[(size_t)i << bits2], hi = base[(size_t)(i + 1) << bits2];
            cursor1[i] = lo;
            tiles = (uint32_t)((hi - lo + (uint64_t)tile_rows - 1) / (uint64_t)tile_rows);
        }
        uint32_t tot;
        const uint32_t ex = block_excl_scan<1024>(tiles, s_scan, &tot);
        const uint64_t run = s_running;
        if (i < F1) tile_start[i] = (uint32_t)run + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_running = run + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_start[F1] = (uint32_t)s_running;
}

// ---- scatter ----------------------------------------------------------------------------------------------------
struct ScatterArgs {
    int64_t row_base, n;                 // FROM_CHUNK: rows [row_base, row_base + n) of the bound chunk
    const unsigned long long* src;       // !FROM_CHUNK: records sorted by first-level bucket
    const uint64_t* base;                // bucket bases (2^bits + 1)
    const uint32_t* tile_start;          // !FROM_CHUNK: first tile of every first-level bucket (+ total)
    unsigned long long* cursor;          // write cursors of the destination buckets
    unsigned long long* dst;             // destination records
    int32_t fan_bits;                    // log2 of this level's fan-out
    int32_t local_shift;                 // local bucket = (bucket >> local_shift) & (fan - 1)
};

template <int W>
__device__ __forceinline__ void aggp_make_record(const AggDev& a, const PartPlan& pl, const VTab& vt, int64_t row, unsigned long long (&rec)[W], HKey& key) {
    ChunkLoader ld{vt, row};
    agg_pack_key(a, ld, key);
    unsigned long long nm = 0;
#pragma unroll
    for (int w = 0; w < W; w++) {
        const int kind = pl.word_kind[w];
        if (kind == WK_KEY_LO) {
            rec[w] = key.lo;
        } else if (kind == WK_KEY_HI) {
            rec[w] = key.hi;
        } else if (kind == WK_NULLS) {
            rec[w] = nm; // the null word is the last one: every function has been evaluated
        } else {
            const int f = pl.word_fn[w];
            int64_t bits;
            const bool nul = eval_expr(a.fns[f].input, ld, bits);
            rec[w] = nul ? 0ull : (unsigned long long)bits;
            nm |= (nul ? 1ull : 0ull) << f;
        }
    }
}

template <int W>
__device__ __forceinline__ void aggp_make_record_simple(const PartPlan& pl, const VTab& vt, int64_t row, unsigned long long (&rec)[W], HKey& key) {
    ChunkLoader ld{vt, row};
#pragma unroll
    for (int w = 0; w < W; w++) {
        int64_t bits;
        ld.load(pl.word_vid[w], bits);
        rec[w] = (unsigned long long)bits;
    }
    rec[0] &= pl.simple_key_mask;
    key.lo = rec[0];
    key.hi = 0;
}

template <int W>
__device__ __forceinline__ void aggp_load_record(const unsigned long long* __restrict__ p, unsigned long long (&rec)[W]) {
    if (W == 2) {
        const ulonglong2 v = __ldg((const ulonglong2*)p);
        rec[0] = v.x;
        rec[1] = v.y;
    } else {
#pragma unroll
        for (int w = 0; w < W; w++) rec[w] = __ldg(p + w);
    }
}
template <int W>
__device__ __forceinline__ void aggp_store_record(unsigned long long* p, const unsigned long long* s) {
    if (W == 2) {
        *(ulonglong2*)p = *(const ulonglong2*)s;
    } else {
#pragma unroll
        for (int w = 0; w < W; w++) p[w] = s[w];
    }
}

template <int W>
constexpr size_t aggp_scatter_smem() {
    return (size_t)AGGP_BLOCK * aggp_rows_per_thread(W) * (W * 8 + 6) + (size_t)AGGP_MAX_FAN * 20;
}

// One tile = T consecutive rows (first level: of the input chunk; second level: of one first-level bucket).  The tile's
// records go to shared memory in ROW order as they are produced (a few rows per thread at a time, so that their loads are
// in flight together and nothing has to stay in registers across the barriers), the tile histogram is scanned, every
// row claims a position inside its bucket's run, and the copy-out walks the tile in BUCKET order through the inverse
// permutation: consecutive threads write consecutive records of a bucket's run.
enum ScatterMode { SCATTER_RECORDS = 0, SCATTER_CHUNK = 1, SCATTER_CHUNK_SIMPLE = 2 };
template <int W, int MODE>
__global__ void __launch_bounds__(AGGP_BLOCK, 2) k_aggp_scatter(const AggDev* __restrict__ ad, const __grid_constant__ VTab vt, const __grid_constant__ PartPlan pl,
                                                                 const __grid_constant__ ScatterArgs sa) {
    constexpr int R = aggp_rows_per_thread(W);
    constexpr int G = aggp_row_group(W);
    constexpr int T = AGGP_BLOCK * R;
    constexpr bool FROM_CHUNK = MODE != SCATTER_RECORDS;
    static_assert(R % G == 0, "row groups");
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ uint32_t s_scan[AGGP_BLOCK / 32 + 1];
    unsigned long long* s_rec = (unsigned long long*)s_raw;              // T records in row order
    unsigned long long* s_gbase = s_rec + (size_t)T * W;                  // destination of the bucket's run
    uint32_t* s_hist = (uint32_t*)(s_gbase + AGGP_MAX_FAN);               // records of the tile per bucket
    uint32_t* s_start = s_hist + AGGP_MAX_FAN;                            // first sorted position of the bucket
    uint32_t* s_cur = s_start + AGGP_MAX_FAN;                             // next free sorted position of the bucket
    uint16_t* s_lb = (uint16_t*)(s_cur + AGGP_MAX_FAN);                   // row -> bucket (0xFFFF: no record)
    uint16_t* s_perm = s_lb + T;                                          // sorted position -> row of the tile
    uint16_t* s_bkt = s_perm + T;                                         // sorted position -> bucket
    const AggDev& a = *ad;
    const int tid = threadIdx.x;
    const int F = 1 << sa.fan_bits;
    const uint32_t fmask = (uint32_t)F - 1;
    const int per = (F + AGGP_BLOCK - 1) / AGGP_BLOCK;
    const int64_t ntiles = FROM_CHUNK ? (sa.n + T - 1) / T : (int64_t)sa.tile_start[1 << (pl.bits - pl.bits2)];
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int64_t row0;        // first row / record of the tile
        int tile_n;
        uint32_t cbase = 0;  // cursor index of local bucket 0
        if (FROM_CHUNK) {
            row0 = tile * T;
            tile_n = (int)(sa.n - row0 < T ? sa.n - row0 : T);
        } else {
            // first-level bucket b1 owning this tile: last entry of tile_start <= tile
            const int F1 = 1 << (pl.bits - pl.bits2);
            int lo = 0, hi = F1;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(sa.tile_start + mid) <= (uint32_t)tile) lo = mid; else hi = mid;
            }
            const uint64_t pb = sa.base[(size_t)lo << pl.bits2], pe = sa.base[(size_t)(lo + 1) << pl.bits2];
            row0 = (int64_t)pb + (tile - (int64_t)__ldg(sa.tile_start + lo)) * T;
            tile_n = (int)((int64_t)pe - row0 < T ? (int64_t)pe - row0 : T);
            cbase = (uint32_t)lo << pl.bits2;
        }
        for (int i = tid; i < F; i += AGGP_BLOCK) s_hist[i] = 0;
        __syncthreads();
#pragma unroll 1
        for (int k0 = 0; k0 < R; k0 += G) {
            unsigned long long rec[G][W];
            HKey key[G];
#pragma unroll
            for (int g = 0; g < G; g++) {
                const int q = (k0 + g) * AGGP_BLOCK + tid;
                if (q < tile_n) {
                    if (MODE == SCATTER_CHUNK_SIMPLE) {
                        aggp_make_record_simple<W>(pl, vt, sa.row_base + row0 + q, rec[g], key[g]);
                    } else if (MODE == SCATTER_CHUNK) {
                        aggp_make_record<W>(a, pl, vt, sa.row_base + row0 + q, rec[g], key[g]);
                    } else {
                        aggp_load_record<W>(sa.src + (size_t)(row0 + q) * W, rec[g]);
                        key[g].lo = rec[g][0];
                        key[g].hi = a.wide ? rec[g][W > 1 ? 1 : 0] : 0ull;
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < G; g++) {
                const int q = (k0 + g) * AGGP_BLOCK + tid;
                if (q < tile_n) {
                    uint32_t l = 0xFFFFu;
                    if (FROM_CHUNK && hkey_is_empty(a, key[g])) { // the table's special slot: applied here, exactly once, never staged
                        AccPtrs gp;
                        acc_ptrs_global(a, gp);
                        ChunkLoader ld{vt, sa.row_base + row0 + q};
                        agg_apply_row<false>(a, gp, (long long)a.cap, ld);
                    } else {
                        aggp_store_record<W>(s_rec + (size_t)q * W, rec[g]);
                        l = (aggp_bucket(a, pl, key[g]) >> sa.local_shift) & fmask;
                        atomicAdd(&s_hist[l], 1u);
                    }
                    s_lb[q] = (uint16_t)l;
                }
            }
        }
        __syncthreads();
        {   // exclusive scan of the tile histogram; reserve the destination runs
            uint32_t local = 0;
            for (int i = 0; i < per; i++) {
                const int idx = tid * per + i;
                if (idx < F) local += s_hist[idx];
            }
            uint32_t tot;
            uint32_t run = block_excl_scan<AGGP_BLOCK>(local, s_scan, &tot);
            for (int i = 0; i < per; i++) {
                const int idx = tid * per + i;
                if (idx < F) {
                    const uint32_t c = s_hist[idx];
                    s_start[idx] = run;
                    s_cur[idx] = run;
                    run += c;
                    if (c) s_gbase[idx] = atomicAdd(sa.cursor + cbase + idx, (unsigned long long)c);
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < R; k++) {
            const int q = k * AGGP_BLOCK + tid;
            if (q < tile_n) {
                const uint32_t l = s_lb[q];
                if (l != 0xFFFFu) {
                    const uint32_t pos = atomicAdd(&s_cur[l], 1u);
                    s_perm[pos] = (uint16_t)q;
                    s_bkt[pos] = (uint16_t)l;
                }
            }
        }
        __syncthreads();
        const int staged = (int)s_cur[F - 1]; // records of the tile (rows of the special slot dropped out)
        for (int pos = tid; pos < staged; pos += AGGP_BLOCK) {
            const uint32_t l = s_bkt[pos];
            const unsigned long long d = s_gbase[l] + (unsigned long long)(pos - s_start[l]);
            aggp_store_record<W>(sa.dst + d * W, s_rec + (size_t)s_perm[pos] * W);
        }
        __syncthreads();
    }
}

// ---- apply: shared-memory slices ------------------------------------------------------------------------------------
// multi-word add into shared memory through native 32-bit atomics: every carry out of a word is seen by exactly one
// adder (atom returns the old value) and forwarded as a +1 to the next word.  Readers look at the words after a barrier.
__device__ __forceinline__ void smem_add_i128(long long* lo, long long* hi, long long v) {
    const uint32_t base_lo = (uint32_t)__cvta_generic_to_shared(lo), base_hi = (uint32_t)__cvta_generic_to_shared(hi);
    const uint32_t sign = v < 0 ? 0xFFFFFFFFu : 0u;
    const uint32_t w[4] = {(uint32_t)v, (uint32_t)((unsigned long long)v >> 32), sign, sign};
    uint32_t carry = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t addr = (i < 2 ? base_lo : base_hi) + 4u * (uint32_t)(i & 1);
        if (w[i] == 0xFFFFFFFFu && carry) continue; // + 2^32: the word is unchanged, the carry moves on
        const uint32_t add = w[i] + carry;
        carry = 0;
        if (add) {
            uint32_t old;
            asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(add) : "memory");
            carry = (old + add) < old ? 1u : 0u;
        }
    }
}
__device__ __forceinline__ void acc_apply_slice(int32_t mode, long long* a0, long long* a1, long long slot, long long bits) {
    if (mode == M_SUM_I128)
        smem_add_i128(a0 + slot, a1 + slot, bits);
    else
        acc_apply_shared(mode, a0, a1, slot, bits);
}

// shared-memory image of one slice: [keys | cnt | per fn: acc0, acc1?, accn?], S slots each
__device__ __forceinline__ void aggp_slice_ptrs(const AggDev& a, unsigned char* smem, int S, unsigned long long*& keys, AccPtrs& p) {
    keys = (unsigned long long*)smem;
    long long* q = (long long*)(keys + (size_t)S * (a.wide ? 2 : 1));
    p.cnt = q;
    q += S;
    for (int f = 0; f < SR_MAX_AGG_FNS; f++) {
        p.acc0[f] = p.acc1[f] = p.accn[f] = nullptr;
        if (f < a.num_fns && a.fns[f].mode != M_COUNT_STAR) {
            p.acc0[f] = q;
            q += S;
            if (a.fns[f].mode == M_SUM_I128) {
                p.acc1[f] = q;
                q += S;
            }
            if (a.fns[f].track_n) {
                p.accn[f] = q;
                q += S;
            }
        }
    }
}

struct ApplyArgs {
    const unsigned long long* rec;
    const uint64_t* base;
    uint32_t num_buckets;
    int32_t fresh; // the table holds no group yet: slices are initialised in shared memory instead of loaded
    uint32_t* fail_list;            // buckets whose slice overflowed (their records are re-applied by k_aggp_apply_l2)
    unsigned long long* fail_count;
};

__global__ void __launch_bounds__(AGGP_BLOCK, 2) k_aggp_apply_smem(const AggDev* __restrict__ ad, const __grid_constant__ PartPlan pl, const __grid_constant__ ApplyArgs aa) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ int s_fail;
    __shared__ uint32_t s_new[AGGP_BLOCK / 32];
    const AggDev& a = *ad;
    const int S = 1 << a.slice_log2;
    const uint32_t smask = (uint32_t)S - 1;
    const int W = pl.words;
    const int tid = threadIdx.x;
    // the per-function array pointers are indexed by a run-time function number: kept in shared memory (a local array
    // would live in local memory)
    __shared__ AccPtrs sp, gp;
    __shared__ unsigned long long* s_keys_p;
    if (tid == 0) {
        unsigned long long* kp;
        AccPtrs t;
        aggp_slice_ptrs(a, s_raw, S, kp, t);
        sp = t;
        s_keys_p = kp;
        acc_ptrs_global(a, t);
        gp = t;
    }
    __syncthreads();
    unsigned long long* const s_keys = s_keys_p;
    const bool fast2 = W == 2 && !a.wide; // (key, one value)
    const int kw = a.wide ? 2 : 1;
    for (uint32_t b = blockIdx.x; b < aa.num_buckets; b += gridDim.x) {
        const int64_t r0 = (int64_t)aa.base[b], r1 = (int64_t)aa.base[b + 1];
        if (r1 <= r0) continue;
        const size_t g0 = (size_t)b << a.slice_log2; // first slot of the slice
        if (tid == 0) s_fail = 0;
        if (aa.fresh) {
            for (int i = tid; i < S * kw; i += AGGP_BLOCK) s_keys[i] = SR_AGG_EMPTY;
            for (int i = tid; i < S; i += AGGP_BLOCK) {
                sp.cnt[i] = 0;
                for (int f = 0; f < a.num_fns; f++) {
                    if (sp.acc0[f]) sp.acc0[f][i] = acc_init_value(a.fns[f].mode);
                    if (sp.acc1[f]) sp.acc1[f][i] = 0;
                    if (sp.accn[f]) sp.accn[f][i] = 0;
                }
            }
        } else {
            for (int i = tid; i < S * kw; i += AGGP_BLOCK) s_keys[i] = a.hkeys[g0 * kw + i];
            for (int i = tid; i < S; i += AGGP_BLOCK) sp.cnt[i] = gp.cnt[g0 + i];
            for (int f = 0; f < a.num_fns; f++) {
                if (sp.acc0[f])
                    for (int i = tid; i < S; i += AGGP_BLOCK) sp.acc0[f][i] = gp.acc0[f][g0 + i];
                if (sp.acc1[f])
                    for (int i = tid; i < S; i += AGGP_BLOCK) sp.acc1[f][i] = gp.acc1[f][g0 + i];
                if (sp.accn[f])
                    for (int i = tid; i < S; i += AGGP_BLOCK) sp.accn[f][i] = gp.accn[f][g0 + i];
            }
        }
        __syncthreads();
        uint32_t my_new = 0;
        for (int64_t q = r0 + tid; q < r1; q += AGGP_BLOCK) {
            const unsigned long long* rp = aa.rec + (size_t)q * W;
            HKey key;
            unsigned long long w1 = 0;
            if (fast2) { // one 16-byte load
                const ulonglong2 v = __ldg((const ulonglong2*)rp);
                key.lo = v.x;
                w1 = v.y;
                key.hi = 0;
            } else {
                key.lo = __ldg(rp);
                key.hi = a.wide ? __ldg(rp + 1) : 0ull;
            }
            uint32_t s = (uint32_t)hkey_hash(a, key) & smask; // slice_log2 <= log2(cap): the low bits of the home slot
            long long slot = -1;
            for (int tries = 0; tries < S; tries++) {
                if (!a.wide) {
                    unsigned long long cur = *(volatile unsigned long long*)(s_keys + s);
                    if (cur == SR_AGG_EMPTY) {
                        cur = atomicCAS(s_keys + s, SR_AGG_EMPTY, key.lo);
                        if (cur == SR_AGG_EMPTY) {
                            my_new++;
                            cur = key.lo;
                        }
                    }
                    if (cur == key.lo) {
                        slot = s;
                        break;
                    }
                } else {
                    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(s_keys + 2 * s);
                    HKey cur;
                    asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(cur.lo), "=l"(cur.hi) : "r"(addr) : "memory");
                    if (cur.lo == SR_AGG_EMPTY && cur.hi == SR_AGG_EMPTY) {
                        asm volatile(
                                "{ .reg .b128 c, n, o;\n"
                                "  mov.b128 c, {%2, %3};\n"
                                "  mov.b128 n, {%4, %5};\n"
                                "  atom.shared.cas.b128 o, [%6], c, n;\n"
                                "  mov.b128 {%0, %1}, o; }"
                                : "=l"(cur.lo), "=l"(cur.hi)
                                : "l"(SR_AGG_EMPTY), "l"(SR_AGG_EMPTY), "l"(key.lo), "l"(key.hi), "r"(addr)
                                : "memory");
                        if (cur.lo == SR_AGG_EMPTY && cur.hi == SR_AGG_EMPTY) {
                            my_new++;
                            cur = key;
                        }
                    }
                    if (cur.lo == key.lo && cur.hi == key.hi) {
                        slot = s;
                        break;
                    }
                }
                s = (s + 1) & smask;
            }
            if (slot < 0) {
                s_fail = 1;
                continue;
            }
            smem_add_u64(sp.cnt + slot, 1ull);
            const unsigned long long nm = pl.null_word >= 0 ? __ldg(rp + pl.null_word) : 0ull;
#pragma unroll 1
            for (int f = 0; f < a.num_fns; f++) {
                const AggFnDev& fn = a.fns[f];
                if (fn.mode == M_COUNT_STAR || ((nm >> f) & 1ull)) continue;
                const long long bits = (long long)(fast2 ? w1 : __ldg(rp + pl.val_word[f]));
                acc_apply_slice(fn.mode, sp.acc0[f], sp.acc1[f], slot, bits);
                if (fn.track_n) smem_add_u64(sp.accn[f] + slot, 1ull);
            }
        }
        __syncthreads();
        if (s_fail) {
            // the slice cannot take every group of its bucket: leave the table untouched, hand the bucket back
            if (tid == 0) aa.fail_list[atomicAdd(aa.fail_count, 1ull)] = b;
        } else {
            for (int i = tid; i < S * kw; i += AGGP_BLOCK) a.hkeys[g0 * kw + i] = s_keys[i];
            for (int i = tid; i < S; i += AGGP_BLOCK) gp.cnt[g0 + i] = sp.cnt[i];
            for (int f = 0; f < a.num_fns; f++) {
                if (sp.acc0[f])
                    for (int i = tid; i < S; i += AGGP_BLOCK) gp.acc0[f][g0 + i] = sp.acc0[f][i];
                if (sp.acc1[f])
                    for (int i = tid; i < S; i += AGGP_BLOCK) gp.acc1[f][g0 + i] = sp.acc1[f][i];
                if (sp.accn[f])
                    for (int i = tid; i < S; i += AGGP_BLOCK) gp.accn[f][g0 + i] = sp.accn[f][i];
            }
            const uint32_t wn = warp_sum(my_new);
            if (lane_id() == 0) s_new[tid >> 5] = wn;
            __syncthreads();
            if (tid == 0) {
                unsigned long long t = 0;
                for (int w = 0; w < AGGP_BLOCK / 32; w++) t += s_new[w];
                if (t) atomicAdd(a.ngroups, t);
            }
        }
        __syncthreads();
    }
}

// ---- apply: global atomics on an L2-prefetched table range ----------------------------------------------------------
// records [r0, r1), or the records named by `list` (refused by the admission limit before a growth).  [s_lo, s_hi): the
// slots the records map to, requested into L2 with full-line prefetches up front.
__global__ void __launch_bounds__(AGG_BLOCK) k_aggp_apply_l2(const AggDev* __restrict__ ad, const __grid_constant__ PartPlan pl, const unsigned long long* __restrict__ rec,
                                                              int64_t r0, int64_t r1, const uint64_t* __restrict__ list, unsigned long long s_lo, unsigned long long s_hi,
                                                              uint64_t* __restrict__ fail_list, unsigned long long* __restrict__ fail_count) {
    const AggDev& a = *ad;
    AccPtrs p;
    acc_ptrs_global(a, p);
    const int W = pl.words;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    if (s_hi > s_lo) {
        const unsigned long long lines = ((s_hi - s_lo) * 8 + 127) / 128; // 8-byte state words, 128-byte lines
        for (unsigned long long l = (unsigned long long)tid; l < lines; l += (unsigned long long)nth) {
            const unsigned long long w = s_lo + l * 16;
            if (a.wide) {
                prefetch_l2(a.hkeys + 2 * w);
                prefetch_l2(a.hkeys + 2 * w + 16);
            } else {
                prefetch_l2(a.hkeys + w);
            }
            prefetch_l2(p.cnt + w);
            for (int f = 0; f < a.num_fns; f++) {
                if (p.acc0[f]) prefetch_l2(p.acc0[f] + w);
                if (p.acc1[f]) prefetch_l2(p.acc1[f] + w);
                if (p.accn[f]) prefetch_l2(p.accn[f] + w);
            }
        }
    }
    unsigned long long known_groups = *(volatile unsigned long long*)a.ngroups;
    for (int64_t q = r0 + tid; q < r1; q += nth) {
        const int64_t i = list ? (int64_t)list[q] : q;
        const unsigned long long* rp = rec + (size_t)i * W;
        const HKey key{rp[0], a.wide ? rp[1] : 0ull};
        bool inserted;
        const long long slot = agg_find_slot_key(a, key, inserted, known_groups);
        agg_count_new_groups(a, inserted, known_groups);
        if (slot < 0) { // table / slice full: remember the record, the host grows the table and re-applies the list
            fail_list[atomicAdd(fail_count, 1ull)] = (uint64_t)i;
            continue;
        }
        atomicAdd((unsigned long long*)p.cnt + slot, 1ull);
        const unsigned long long nm = pl.null_word >= 0 ? rp[pl.null_word] : 0ull;
#pragma unroll 1
        for (int f = 0; f < a.num_fns; f++) {
            const AggFnDev& fn = a.fns[f];
            if (fn.mode == M_COUNT_STAR || ((nm >> f) & 1ull)) continue;
            acc_apply(fn.mode, p.acc0[f], p.acc1[f], slot, (long long)rp[pl.val_word[f]]);
            if (fn.track_n) atomicAdd((unsigned long long*)p.accn[f] + slot, 1ull);
        }
    }
}

} // namespace srd
