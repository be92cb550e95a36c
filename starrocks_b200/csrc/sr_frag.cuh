// sr_frag.cuh -- the fused pipeline fragment: scan -> filter -> [hash-join probe]* -> aggregate
// in ONE pass over the fact columns.  This is the hot kernel of the SSB Q4.1 / Q1.1 configs.
//
// What it fuses (reference call stack):
//   OlapChunkSource::_read_chunk_from_storage filter step   scan/olap_chunk_source.cpp:676-722
//   HashJoinProbeOperator::push_chunk/pull_chunk x N        hashjoin/hash_join_probe_operator.cpp:55-96
//     -> JoinHashMap::probe (lookup_init + _probe_from_ht + _probe_output/_build_output)
//   AggregateBlockingSinkOperator::push_chunk               aggregate/aggregate_blocking_sink_operator.cpp:101-138
//
// B200 design (HBM-bound integer/gather work, no tensor cores):
//  * persistent CTAs (grid = SMs x resident CTAs), each thread owns 4 consecutive rows so the
//    first-touched column of a tile is one 128-bit ld.global.nc.L1::no_allocate per thread;
//  * late materialisation: a column is only loaded for rows still alive, so DRAM sectors of
//    later columns are skipped when all 8 rows of a 32-byte sector are already filtered out;
//  * every range-mapped join is first tested against its 1-bit-per-key bitmap; the bitmaps of
//    the earliest (most selective) joins are copied into shared memory once per CTA, later
//    ones are read through L1/L2 (they stay resident: SSB dimensions are a few hundred KB);
//  * the build row index (first[]) and payload columns are only fetched for rows that survive
//    all joins; the group-by / SUM update goes to shared-memory accumulators when the table is
//    small, flushed once per CTA;
//  * join order: measured pass rates on a sample of the first batch (adaptive), most selective
//    first -- inner joins commute and the result is order independent.
#pragma once

#include "sr_agg.cuh"

namespace srd {

struct FragJoinDev {
    JoinDev j;
    int32_t key_value_id;
    int32_t smem_off; // word offset of the bitmap copy in dynamic shared memory, -1 = global
    int32_t bitmap_words;
    int32_t use_bitmap; // range-mapped table: test the bitmap; otherwise probe the hash table
    int32_t need_head;  // a payload column of this join is read downstream
    int32_t pad;
};

struct FragDev {
    int32_t num_preds, num_exprs, num_joins, pad;
    CPred preds[8];
    CExpr exprs[4];
    FragJoinDev joins[SR_MAX_FRAG_JOINS];
    unsigned long long* rows_passed;
};

constexpr int FRAG_BLOCK = 512;
constexpr int FRAG_ROWS = 4;
constexpr int FRAG_TILE = FRAG_BLOCK * FRAG_ROWS;

struct FragLoader {
    const VTab& vt;
    int64_t row;
    uint32_t bidx[SR_MAX_FRAG_JOINS];
    __device__ __forceinline__ bool load(int id, int64_t& bits) const {
        const VDesc& d = vt.v[id];
        if (d.src < 0) {
            const bool nul = d.nulls != nullptr && d.nulls[row] != 0;
            bits = is_float_class(d.type) ? __double_as_longlong(load_double(d.data, d.type, row)) : load_int(d.data, d.type, row);
            return nul;
        }
        const int64_t r = bidx[d.src];
        const bool nul = d.nulls != nullptr && d.nulls[r] != 0;
        if (is_float_class(d.type))
            bits = __double_as_longlong(d.type == SR_TYPE_FLOAT ? (double)__ldg((const float*)d.data + r) : __ldg((const double*)d.data + r));
        else
            bits = load_int_cached(d.data, d.type, r);
        return nul;
    }
};

// load one fact value for each alive row of the thread's 4-row group
__device__ __forceinline__ void load_rows4(const VDesc& d, int64_t row0, uint32_t alive, int64_t vals[FRAG_ROWS], uint32_t& nullmask) {
    nullmask = 0;
    const int w = type_width(d.type);
    if (w == 4 && alive == 0xF && !is_float_class(d.type) && (((uintptr_t)d.data) & 15) == 0) {
        const int4 v = ldg_stream_v4((const int32_t*)d.data + row0); // row0 % 4 == 0 and cudaMalloc alignment -> 16B aligned
        vals[0] = v.x;
        vals[1] = v.y;
        vals[2] = v.z;
        vals[3] = v.w;
    } else {
#pragma unroll
        for (int r = 0; r < FRAG_ROWS; r++) {
            if (alive & (1u << r)) {
                vals[r] = is_float_class(d.type) ? __double_as_longlong(load_double(d.data, d.type, row0 + r)) : load_int(d.data, d.type, row0 + r);
            }
        }
    }
    if (d.nulls) {
#pragma unroll
        for (int r = 0; r < FRAG_ROWS; r++)
            if ((alive & (1u << r)) && d.nulls[row0 + r]) nullmask |= 1u << r;
    }
}

__device__ __forceinline__ bool frag_join_hit(const FragJoinDev& fj, const uint32_t* smem, int64_t key) {
    if (fj.use_bitmap) {
        if (key < fj.j.min_value || key > fj.j.max_value) return false;
        const uint64_t idx = (uint64_t)(key - fj.j.min_value);
        const uint32_t word = fj.smem_off >= 0 ? smem[fj.smem_off + (idx >> 5)] : __ldg(fj.j.bitmap + (idx >> 5));
        return (word >> (idx & 31)) & 1u;
    }
    return join_lookup(fj.j, key) != 0;
}

template <bool SMEM_AGG>
__global__ void __launch_bounds__(FRAG_BLOCK) k_fragment(const FragDev* __restrict__ fdp, const AggDev* __restrict__ adp, VTab vt, int64_t n,
                                                          int32_t agg_smem_word_off) {
    extern __shared__ __align__(16) uint32_t smem[];
    const FragDev& fd = *fdp;
    const AggDev& ad = *adp;
    // stage the bitmaps of the leading joins in shared memory
    for (int j = 0; j < fd.num_joins; j++) {
        const FragJoinDev& fj = fd.joins[j];
        if (fj.smem_off >= 0)
            for (int w = threadIdx.x; w < fj.bitmap_words; w += blockDim.x) smem[fj.smem_off + w] = fj.j.bitmap[w];
    }
    AccPtrs acc;
    if (SMEM_AGG) {
        acc_ptrs_smem(ad, (long long*)(smem + agg_smem_word_off), acc);
        acc_smem_init(ad, acc);
    } else {
        acc_ptrs_global(ad, acc);
    }
    __syncthreads();

    unsigned long long passed = 0;
    const int64_t num_tiles = (n + FRAG_TILE - 1) / FRAG_TILE;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * FRAG_TILE + (int64_t)threadIdx.x * FRAG_ROWS;
        uint32_t alive = 0;
#pragma unroll
        for (int r = 0; r < FRAG_ROWS; r++)
            if (row0 + r < n) alive |= 1u << r;
        int64_t vals[FRAG_ROWS];
        uint32_t nullmask;
        // ---- scan conjuncts (ColumnPredicate form) ----
#pragma unroll 1
        for (int p = 0; p < fd.num_preds && alive; p++) {
            const CPred& pr = fd.preds[p];
            load_rows4(vt.v[pr.value_id], row0, alive, vals, nullmask);
#pragma unroll
            for (int r = 0; r < FRAG_ROWS; r++)
                if ((alive & (1u << r)) && !eval_pred(pr, vals[r], (nullmask >> r) & 1u)) alive &= ~(1u << r);
        }
        // ---- generic boolean conjuncts ----
#pragma unroll 1
        for (int e = 0; e < fd.num_exprs && alive; e++) {
#pragma unroll 1
            for (int r = 0; r < FRAG_ROWS; r++) {
                if (alive & (1u << r)) {
                    ChunkLoader ld{vt, row0 + r};
                    int64_t bits;
                    const bool nul = eval_expr(fd.exprs[e], ld, bits);
                    if (nul || bits == 0) alive &= ~(1u << r);
                }
            }
        }
        // ---- join probes, most selective first ----
#pragma unroll 1
        for (int j = 0; j < fd.num_joins && alive; j++) {
            const FragJoinDev& fj = fd.joins[j];
            load_rows4(vt.v[fj.key_value_id], row0, alive, vals, nullmask);
            alive &= ~nullmask; // NULL keys never match (join_hash_table.cpp:166-170)
#pragma unroll
            for (int r = 0; r < FRAG_ROWS; r++)
                if ((alive & (1u << r)) && !frag_join_hit(fj, smem, vals[r])) alive &= ~(1u << r);
        }
        // ---- survivors: build row indexes, group slot, aggregate update ----
        if (alive) {
#pragma unroll 1
            for (int r = 0; r < FRAG_ROWS; r++) {
                if (!(alive & (1u << r))) continue;
                FragLoader ld{vt, row0 + r, {0, 0, 0, 0, 0, 0}};
                for (int j = 0; j < fd.num_joins; j++) {
                    const FragJoinDev& fj = fd.joins[j];
                    if (fj.need_head) {
                        int64_t key;
                        ld.load(fj.key_value_id, key);
                        ld.bidx[j] = join_lookup(fj.j, key);
                    }
                }
                const long long slot = agg_find_slot(ad, ld);
                if (slot >= 0) agg_apply_row(ad, acc, slot, ld);
                passed++;
            }
        }
    }
    if (SMEM_AGG) {
        __syncthreads();
        acc_smem_flush(ad, acc);
    }
    passed = warp_sum(passed);
    if (lane_id() == 0 && passed) atomicAdd(fd.rows_passed, passed);
}

// adaptive join ordering: independent pass counts of each join on a sample of rows
__global__ void __launch_bounds__(256) k_frag_sample(const FragDev* __restrict__ fdp, VTab vt, int64_t n, unsigned long long* __restrict__ counts) {
    const FragDev& fd = *fdp;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += (int64_t)gridDim.x * blockDim.x) {
        ChunkLoader ld{vt, row};
        for (int j = 0; j < fd.num_joins; j++) {
            const FragJoinDev& fj = fd.joins[j];
            int64_t key;
            const bool nul = ld.load(fj.key_value_id, key);
            const bool hit = !nul && join_lookup(fj.j, key) != 0;
            const uint32_t m = __ballot_sync(__activemask(), hit);
            if (hit && (m & lanemask_lt()) == 0) atomicAdd(&counts[j], (unsigned long long)__popc(m));
        }
    }
}

} // namespace srd

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct sr_fragment {
    sr_ctx* ctx = nullptr;
    std::vector<sr_pred> preds;
    std::vector<sr_expr> exprs;
    int32_t num_joins = 0;
    sr_frag_join joins[SR_MAX_FRAG_JOINS];
    sr_agg* agg = nullptr;
    bool compiled = false;
    VReg reg;          // fact + payload values
    int num_fact_values = 0;
    std::vector<int32_t> value_src;       // per value: -1 fact, j = payload of (ordered) join j
    std::vector<const BuildCol*> value_col; // payload column for src >= 0
    srd::FragDev host;
    DevBuf dev, counters;
    Staged staged;
    size_t smem_bytes = 0;
    int32_t agg_smem_word_off = 0;
    bool smem_agg = false;
    int grid = 0;
    int order[SR_MAX_FRAG_JOINS];
    double pass_rate[SR_MAX_FRAG_JOINS];
};

struct FragTypeCtx {
    sr_fragment* f;
};

// slot type: fact chunk first, then payload columns of the joins
static int32_t frag_slot_type(void* user, int32_t slot) {
    sr_fragment* f = ((FragTypeCtx*)user)->f;
    const int k = f->staged.find(slot);
    if (k >= 0) return f->staged.cols[k].type;
    for (int j = 0; j < f->num_joins; j++)
        for (int p = 0; p < f->joins[j].num_payload; p++)
            if (f->joins[j].payload_build_slots[p] == slot) {
                const BuildCol* bc = f->joins[j].join->find_col(slot);
                return bc ? bc->type : 0;
            }
    return 0;
}
static bool frag_slot_nullable(void* user, int32_t slot) {
    sr_fragment* f = ((FragTypeCtx*)user)->f;
    const int k = f->staged.find(slot);
    if (k >= 0) return f->staged.cols[k].nulls != nullptr;
    for (int j = 0; j < f->num_joins; j++)
        for (int p = 0; p < f->joins[j].num_payload; p++)
            if (f->joins[j].payload_build_slots[p] == slot) {
                const BuildCol* bc = f->joins[j].join->find_col(slot);
                return bc && bc->nullable;
            }
    return false;
}

static int32_t frag_bind_vtab(sr_fragment* f, VTab* vt) {
    sr_ctx* ctx = f->ctx;
    vt->n = (int32_t)f->reg.slots.size();
    for (size_t k = 0; k < f->reg.slots.size(); k++) {
        if (f->value_src[k] < 0) {
            const int c = f->staged.find(f->reg.slots[k]);
            if (c < 0) return sr_fail(ctx, SR_ERR_INVALID_ARGUMENT, "fact chunk misses slot %d", f->reg.slots[k]);
            if (f->staged.cols[c].type != f->reg.types[k]) return sr_fail(ctx, SR_ERR_INVALID_ARGUMENT, "fact slot %d changed type", f->reg.slots[k]);
            vt->v[k].data = f->staged.cols[c].data;
            vt->v[k].nulls = f->staged.cols[c].nulls;
            vt->v[k].type = f->staged.cols[c].type;
            vt->v[k].src = -1;
        } else {
            const BuildCol* bc = f->value_col[k];
            vt->v[k].data = bc->data.p;
            vt->v[k].nulls = bc->nullable ? (const uint8_t*)bc->nulls.p : nullptr;
            vt->v[k].type = bc->type;
            vt->v[k].src = f->value_src[k];
        }
    }
    return SR_OK;
}

static int32_t frag_compile(sr_fragment* f) {
    sr_ctx* ctx = f->ctx;
    FragTypeCtx tc{f};
    srd::FragDev& h = f->host;
    memset(&h, 0, sizeof(h));
    f->reg = VReg();
    if (f->preds.size() > 8 || f->exprs.size() > 4) return sr_fail(ctx, SR_ERR_NOT_SUPPORTED, "too many scan conjuncts for the fused fragment");
    for (size_t k = 0; k < f->preds.size(); k++) SR_TRY(compile_pred(ctx, &f->preds[k], &f->reg, frag_slot_type, &tc, &h.preds[k]));
    for (size_t k = 0; k < f->exprs.size(); k++) {
        SR_TRY(compile_expr(ctx, &f->exprs[k], &f->reg, frag_slot_type, &tc, &h.exprs[k]));
        if (h.exprs[k].result_is_double) return sr_fail(ctx, SR_ERR_INVALID_ARGUMENT, "filter expression %zu is not boolean", k);
    }
    h.num_preds = (int32_t)f->preds.size();
    h.num_exprs = (int32_t)f->exprs.size();
    h.num_joins = f->num_joins;
    // join keys (identity order for now; reordered after sampling)
    for (int j = 0; j < f->num_joins; j++) {
        sr_join* jn = f->joins[j].join;
        const int c = f->staged.find(f->joins[j].probe_key_slot);
        if (c < 0) return sr_fail(ctx, SR_ERR_INVALID_ARGUMENT, "fact chunk misses probe key slot %d", f->joins[j].probe_key_slot);
        if (srd::is_float_class(f->staged.cols[c].type) || f->staged.cols[c].width > 8)
            return sr_fail(ctx, SR_ERR_INVALID_ARGUMENT, "probe key slot %d has a non-integer type", f->joins[j].probe_key_slot);
        if (f->staged.cols[c].width != srd::type_width(jn->desc.key_types[0]))
            return sr_fail(ctx, SR_ERR_INVALID_ARGUMENT, "probe key slot %d width differs from the join's key type", f->joins[j].probe_key_slot);
        const int id = f->reg.add(f->joins[j].probe_key_slot, f->staged.cols[c].type);
        if (id >= SR_MAX_VALUES) return sr_fail(ctx, SR_ERR_NOT_SUPPORTED, "too many distinct columns");
        h.joins[j].key_value_id = id;
        h.joins[j].j = jn->dev();
        h.joins[j].smem_off = -1;
        h.joins[j].use_bitmap = jn->method != SR_JOIN_METHOD_LINEAR_CHAINED ? 1 : 0;
        h.joins[j].bitmap_words = (int32_t)((std::max<int64_t>(jn->bucket_size, 1) + 31) / 32);
        h.joins[j].need_head = 0;
        f->order[j] = j;
    }
    // aggregate (its expressions may reference fact slots and payload slots)
    SR_TRY(agg_compile(f->agg, frag_slot_type, frag_slot_nullable, &tc));
    // the fragment and its aggregate must share one value table: rebuild the aggregate's
    // value ids on top of the fragment registry
    {
        VReg merged = f->reg;
        std::vector<int> remap(f->agg->reg.slots.size());
        for (size_t k = 0; k < f->agg->reg.slots.size(); k++) {
            remap[k] = merged.add(f->agg->reg.slots[k], f->agg->reg.types[k]);
            if (remap[k] >= SR_MAX_VALUES) return sr_fail(ctx, SR_ERR_NOT_SUPPORTED, "too many distinct columns");
        }
        srd::AggDev& ah = f->agg->host;
        for (int k = 0; k < ah.num_keys; k++) ah.key_value_id[k] = remap[ah.key_value_id[k]];
        for (int q = 0; q < ah.num_fns; q++)
            for (int k = 0; k < ah.fns[q].input.num_nodes; k++)
                if (ah.fns[q].input.nodes[k].op == srd::C_LOAD_I || ah.fns[q].input.nodes[k].op == srd::C_LOAD_D)
                    ah.fns[q].input.nodes[k].arg = remap[ah.fns[q].input.nodes[k].arg];
        f->agg->reg = merged;
        f->reg = merged;
        SR_TRY(agg_upload(f->agg));
    }
    // classify values: fact column or payload of join j
    f->value_src.assign(f->reg.slots.size(), -1);
    f->value_col.assign(f->reg.slots.size(), nullptr);
    for (size_t k = 0; k < f->reg.slots.size(); k++) {
        if (f->staged.find(f->reg.slots[k]) >= 0) continue;
        bool found = false;
        for (int j = 0; j < f->num_joins && !found; j++)
            for (int p = 0; p < f->joins[j].num_payload && !found; p++)
                if (f->joins[j].payload_build_slots[p] == f->reg.slots[k]) {
                    f->value_src[k] = j;
                    f->value_col[k] = f->joins[j].join->find_col(f->reg.slots[k]);
                    h.joins[j].need_head = 1;
                    found = true;
                }
        if (!found) return sr_fail(ctx, SR_ERR_INVALID_ARGUMENT, "slot %d is neither a fact column nor a join payload", f->reg.slots[k]);
    }
    SR_TRY(f->counters.reserve(ctx, 128));
    SR_CUDA(ctx, cudaMemsetAsync(f->counters.p, 0, 128, ctx->stream));
    h.rows_passed = f->counters.as<unsigned long long>();
    return SR_OK;
}

// sample pass rates, order the joins, assign shared-memory bitmaps, upload, pick the grid
static int32_t frag_plan(sr_fragment* f, const VTab& vt, int64_t n) {
    sr_ctx* ctx = f->ctx;
    srd::FragDev& h = f->host;
    SR_TRY(f->dev.reserve(ctx, sizeof(srd::FragDev)));
    SR_CUDA(ctx, cudaMemcpyAsync(f->dev.p, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
    const int64_t sample = std::min<int64_t>(n, 1 << 16);
    unsigned long long counts[SR_MAX_FRAG_JOINS] = {0};
    if (f->num_joins > 1 && sample > 0) {
        unsigned long long* dcounts = f->counters.as<unsigned long long>() + 8;
        SR_CUDA(ctx, cudaMemsetAsync(dcounts, 0, sizeof(counts), ctx->stream));
        srd::k_frag_sample<<<grid_for(sample, 256), 256, 0, ctx->stream>>>((const srd::FragDev*)f->dev.p, vt, sample, dcounts);
        SR_LAUNCH_CHECK(ctx);
        SR_CUDA(ctx, cudaMemcpyAsync(counts, dcounts, sizeof(counts), cudaMemcpyDeviceToHost, ctx->stream));
        SR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    for (int j = 0; j < f->num_joins; j++) f->pass_rate[j] = sample > 0 ? (double)counts[j] / (double)sample : 1.0;
    std::vector<int> ord(f->num_joins);
    for (int j = 0; j < f->num_joins; j++) ord[j] = j;
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) {
        if (f->pass_rate[a] != f->pass_rate[b]) return f->pass_rate[a] < f->pass_rate[b];
        return h.joins[a].bitmap_words < h.joins[b].bitmap_words;
    });
    srd::FragJoinDev reordered[SR_MAX_FRAG_JOINS];
    std::vector<int> inv(f->num_joins);
    for (int q = 0; q < f->num_joins; q++) {
        reordered[q] = h.joins[ord[q]];
        inv[ord[q]] = q;
        f->order[q] = ord[q];
    }
    for (int q = 0; q < f->num_joins; q++) h.joins[q] = reordered[q];
    for (size_t k = 0; k < f->value_src.size(); k++)
        if (f->value_src[k] >= 0) f->value_src[k] = inv[f->value_src[k]];
    // shared memory: aggregate accumulators first (8-byte aligned), then bitmaps greedily in probe order
    size_t words = 0;
    f->smem_agg = f->agg->smem_bytes > 0;
    f->agg_smem_word_off = 0;
    if (f->smem_agg) words += f->agg->smem_bytes / 4;
    const size_t budget_words = (96 * 1024) / 4; // keeps >= 2 CTAs of 512 threads resident per SM
    for (int q = 0; q < f->num_joins; q++) {
        srd::FragJoinDev& fj = h.joins[q];
        fj.smem_off = -1;
        if (fj.use_bitmap && words + (size_t)fj.bitmap_words <= budget_words) {
            fj.smem_off = (int32_t)words;
            words += (size_t)fj.bitmap_words;
            words = (words + 3) & ~(size_t)3;
        } else {
            break; // later joins see few rows; their bitmaps stay in L1/L2
        }
    }
    f->smem_bytes = words * 4;
    SR_CUDA(ctx, cudaMemcpyAsync(f->dev.p, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
    SR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // occupancy-sized persistent grid
    int per_sm = 0;
    SR_CUDA(ctx, cudaFuncSetAttribute(srd::k_fragment<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(f->smem_bytes, 1)));
    SR_CUDA(ctx, cudaFuncSetAttribute(srd::k_fragment<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(f->smem_bytes, 1)));
    if (f->smem_agg)
        SR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, srd::k_fragment<true>, srd::FRAG_BLOCK, f->smem_bytes));
    else
        SR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, srd::k_fragment<false>, srd::FRAG_BLOCK, f->smem_bytes));
    if (per_sm < 1) return sr_fail(ctx, SR_ERR_CUDA, "fragment kernel does not fit on an SM (smem %zu)", f->smem_bytes);
    f->grid = per_sm * ctx->num_sms;
    f->compiled = true;
    return SR_OK;
}

static int32_t frag_push(sr_fragment* f, const sr_chunk_view* fact) {
    sr_ctx* ctx = f->ctx;
    if (f->agg->finished) return sr_fail(ctx, SR_ERR_STATE, "fragment push after sink_finish");
    SR_TRY(f->staged.stage(ctx, fact));
    const int64_t n = fact->num_rows;
    bool first = !f->compiled;
    if (first) SR_TRY(frag_compile(f));
    VTab vt;
    SR_TRY(frag_bind_vtab(f, &vt));
    if (first) SR_TRY(frag_plan(f, vt, n));
    if (n == 0) return SR_OK;
    SR_TRY(agg_check_nullability(f->agg, vt));
    if (f->agg->smem_bytes == 0 && f->smem_agg) {
        // nullability change forced the aggregate to global accumulation
        f->smem_agg = false;
    }
    sr_agg* a = f->agg;
    const srd::AggDev& ah = a->host;
    const bool hash = !ah.dense && ah.num_keys > 0;
    if (hash) {
        // survivors are unknown before the pass: make the table large enough for the worst case
        while ((uint64_t)a->ngroups_host + (uint64_t)n > a->host.limit) {
            if (a->host.cap >= (1ull << 33)) return sr_fail(ctx, SR_ERR_OUT_OF_MEMORY, "aggregate table would exceed 2^33 slots; push smaller batches");
            SR_TRY(agg_grow(a, a->host.cap * 4));
        }
    }
    const int grid = (int)std::min<int64_t>(f->grid, (n + srd::FRAG_TILE - 1) / srd::FRAG_TILE);
    if (f->smem_agg)
        srd::k_fragment<true><<<grid, srd::FRAG_BLOCK, f->smem_bytes, ctx->stream>>>((const srd::FragDev*)f->dev.p, (const srd::AggDev*)a->dev.p, vt, n,
                                                                                    f->agg_smem_word_off);
    else
        srd::k_fragment<false><<<grid, srd::FRAG_BLOCK, f->smem_bytes, ctx->stream>>>((const srd::FragDev*)f->dev.p, (const srd::AggDev*)a->dev.p, vt, n,
                                                                                     f->agg_smem_word_off);
    SR_LAUNCH_CHECK(ctx);
    if (hash) {
        uint64_t ng;
        int32_t ovf, bad;
        SR_TRY(agg_read_counters(a, &ng, &ovf, &bad));
        a->ngroups_host = (int64_t)ng;
        if (ovf) return sr_fail(ctx, SR_ERR_STATE, "aggregate hash table overflow (internal)");
    }
    return SR_OK;
}
