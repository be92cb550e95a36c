// sr_scan.cuh -- scan predicates -> uint8 selection vector -> order-preserving compaction.
// Replaces K1-K4 of SURVEY.md section 2b:
//   ColumnPredicate::evaluate/evaluate_and   be/src/storage/column_operator_predicate.h:41-111
//   ChunkPredicateEvaluator::eval_conjuncts  be/src/exprs/chunk_predicate_evaluator.cpp:84-149
//   Chunk::filter / t_filter_range<T>        be/src/column/chunk.cpp:362, column_filter_range.cpp:39-148
// HBM-bound streaming kernels: coalesced 4/8-byte loads per lane, selection kept as bytes (the
// reference's Filter), block counts + one scan, then a scatter that keeps input order.
#pragma once

#include "sr_host.cuh"

namespace srd {

struct ChunkLoader {
    const VTab& vt;
    int64_t row;
    __device__ __forceinline__ bool load(int id, int64_t& bits) const {
        const VDesc& d = vt.v[id];
        const bool nul = d.nulls != nullptr && d.nulls[row] != 0;
        bits = is_float_class(d.type) ? __double_as_longlong(load_double(d.data, d.type, row)) : load_int(d.data, d.type, row);
        return nul;
    }
};

struct ScanProg {
    CPred preds[16];
    int32_t num_preds;
    int32_t num_exprs;
    CExpr exprs[8];
};

constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

// selection[i] = all conjuncts true-and-not-null.  block_counts[b] = survivors of tile b.
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_select(const ScanProg* __restrict__ prog, const __grid_constant__ VTab vt, int64_t n,
                                                             uint8_t* __restrict__ sel, uint32_t* __restrict__ block_counts) {
    __shared__ uint32_t s_cnt[SCAN_BLOCK / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const int64_t row = base + k * SCAN_BLOCK + threadIdx.x;
        if (row < n) {
            ChunkLoader ld{vt, row};
            bool pass = true;
            for (int p = 0; p < prog->num_preds && pass; p++) {
                int64_t bits;
                const bool nul = ld.load(prog->preds[p].value_id, bits);
                pass = eval_pred(prog->preds[p], bits, nul);
            }
            for (int e = 0; e < prog->num_exprs && pass; e++) {
                int64_t bits;
                const bool nul = eval_expr(prog->exprs[e], ld, bits);
                pass = !nul && bits != 0;
            }
            sel[row] = pass ? 1 : 0;
            mine += pass ? 1u : 0u;
        }
    }
    if (block_counts) {
        mine = warp_sum(mine);
        if (lane_id() == 0) s_cnt[threadIdx.x >> 5] = mine;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int w = 0; w < SCAN_BLOCK / 32; w++) t += s_cnt[w];
            block_counts[blockIdx.x] = t;
        }
    }
}

// single-block exclusive scan of `n` uint32 counts into uint64 offsets; total -> *total
__global__ void __launch_bounds__(1024) k_scan_counts(const uint32_t* __restrict__ counts, int64_t n, uint64_t* __restrict__ offsets,
                                                       uint64_t* __restrict__ total) {
    __shared__ uint32_t s_scan[1024 / 32 + 1];
    __shared__ uint64_t s_running;
    if (threadIdx.x == 0) s_running = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const uint32_t v = i < n ? counts[i] : 0;
        uint32_t tot;
        const uint32_t ex = block_excl_scan<1024>(v, s_scan, &tot);
        const uint64_t run = s_running;
        if (i < n) offsets[i] = run + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_running = run + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_running;
}

struct CompactCol {
    const void* src;
    void* dst;
    int32_t width;
    int32_t pad;
};
struct CompactArgs {
    CompactCol c[2 * SR_MAX_OUT_COLS];
    int32_t n;
};

template <typename T>
__device__ __forceinline__ void copy_elem(const void* src, void* dst, int64_t s, int64_t d) {
    ((T*)dst)[d] = ((const T*)src)[s];
}

// grid = (tiles, columns).  Keeps input order: out position = tile offset + rank inside tile.
__global__ void __launch_bounds__(SCAN_BLOCK) k_compact(const uint8_t* __restrict__ sel, const uint64_t* __restrict__ tile_offsets,
                                                         CompactArgs args, int64_t n) {
    __shared__ uint32_t s_scan[SCAN_BLOCK / 32 + 1];
    const CompactCol col = args.c[blockIdx.y];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    uint64_t out = tile_offsets[blockIdx.x];
#pragma unroll 1
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const int64_t row = base + k * SCAN_BLOCK + threadIdx.x;
        const uint32_t keep = (row < n && sel[row]) ? 1u : 0u;
        uint32_t tot;
        const uint32_t rank = block_excl_scan<SCAN_BLOCK>(keep, s_scan, &tot);
        if (keep) {
            const int64_t d = (int64_t)(out + rank);
            switch (col.width) {
            case 1:
                copy_elem<uint8_t>(col.src, col.dst, row, d);
                break;
            case 2:
                copy_elem<uint16_t>(col.src, col.dst, row, d);
                break;
            case 4:
                copy_elem<uint32_t>(col.src, col.dst, row, d);
                break;
            case 8:
                copy_elem<uint64_t>(col.src, col.dst, row, d);
                break;
            default:
                copy_elem<int4>(col.src, col.dst, row, d);
                break;
            }
        }
        out += tot;
    }
}

// K11: dst[j] = src[index[j]] (Column::append_selective); index 0 of a build column is the
// sentinel row.  null_out (optional): 1 when index == 0 (outer join miss) or src null.
struct GatherCol {
    const void* src;
    const uint8_t* src_nulls;
    void* dst;
    uint8_t* dst_nulls;
    int32_t width;
    int32_t zero_is_null; // build side of an outer join
};
struct GatherArgs {
    GatherCol c[SR_MAX_OUT_COLS];
    int32_t n;
};

__global__ void __launch_bounds__(256) k_gather(const uint32_t* __restrict__ index, int64_t n, GatherArgs args) {
    const GatherCol col = args.c[blockIdx.y];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = index[i];
        switch (col.width) {
        case 1:
            copy_elem<uint8_t>(col.src, col.dst, s, i);
            break;
        case 2:
            copy_elem<uint16_t>(col.src, col.dst, s, i);
            break;
        case 4:
            copy_elem<uint32_t>(col.src, col.dst, s, i);
            break;
        case 8:
            copy_elem<uint64_t>(col.src, col.dst, s, i);
            break;
        default:
            copy_elem<int4>(col.src, col.dst, s, i);
            break;
        }
        if (col.dst_nulls) {
            uint8_t nu = col.src_nulls ? col.src_nulls[s] : 0;
            if (col.zero_is_null && s == 0) nu = 1;
            col.dst_nulls[i] = nu;
        }
    }
}

} // namespace srd

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct sr_scan {
    sr_ctx* ctx = nullptr;
    std::vector<sr_pred> preds;
    std::vector<sr_expr> exprs;
    std::vector<int32_t> out_slots;
    // compiled against the first chunk's slot types
    bool compiled = false;
    VReg reg;
    DevBuf prog;
    Staged staged;
    DevBuf sel, block_counts, tile_offsets;
    std::vector<DevBuf> out_bufs; // 2 per out slot
};

static int32_t staged_slot_type(void* user, int32_t slot) {
    const Staged* st = (const Staged*)user;
    const int k = st->find(slot);
    return k < 0 ? 0 : st->cols[k].type;
}

static int32_t scan_compile(sr_scan* s) {
    sr_ctx* ctx = s->ctx;
    if (s->preds.size() > 16 || s->exprs.size() > 8) return sr_fail(ctx, SR_ERR_NOT_SUPPORTED, "too many scan conjuncts");
    std::vector<uint8_t> hostbuf(sizeof(srd::ScanProg), 0);
    srd::ScanProg* hp = (srd::ScanProg*)hostbuf.data();
    s->reg = VReg();
    for (size_t k = 0; k < s->preds.size(); k++) SR_TRY(compile_pred(ctx, &s->preds[k], &s->reg, staged_slot_type, &s->staged, &hp->preds[k]));
    for (size_t k = 0; k < s->exprs.size(); k++) {
        SR_TRY(compile_expr(ctx, &s->exprs[k], &s->reg, staged_slot_type, &s->staged, &hp->exprs[k]));
        if (hp->exprs[k].result_is_double) return sr_fail(ctx, SR_ERR_INVALID_ARGUMENT, "filter expression %zu is not boolean", k);
    }
    hp->num_preds = (int32_t)s->preds.size();
    hp->num_exprs = (int32_t)s->exprs.size();
    SR_TRY(s->prog.reserve(ctx, sizeof(srd::ScanProg)));
    SR_CUDA(ctx, cudaMemcpyAsync(s->prog.p, hp, sizeof(srd::ScanProg), cudaMemcpyHostToDevice, ctx->stream));
    SR_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // hostbuf goes out of scope
    s->compiled = true;
    return SR_OK;
}

// runs predicate evaluation; leaves selection bytes in s->sel (device) and, when
// want_counts, tile counts/offsets + the total in ctx->pinned[0] (after a sync)
static int32_t scan_select(sr_scan* s, const sr_chunk_view* in, bool want_counts, int64_t* total_out) {
    sr_ctx* ctx = s->ctx;
    SR_TRY(s->staged.stage(ctx, in));
    if (!s->compiled) SR_TRY(scan_compile(s));
    VTab vt;
    SR_TRY(bind_vtab(ctx, s->reg, s->staged, &vt));
    const int64_t n = in->num_rows;
    const int tiles = grid_for(n, srd::SCAN_TILE);
    SR_TRY(s->sel.reserve(ctx, (size_t)std::max<int64_t>(n, 1)));
    if (want_counts) {
        SR_TRY(s->block_counts.reserve(ctx, sizeof(uint32_t) * (size_t)tiles));
        SR_TRY(s->tile_offsets.reserve(ctx, sizeof(uint64_t) * (size_t)tiles));
    }
    if (n > 0) {
        srd::k_scan_select<<<tiles, srd::SCAN_BLOCK, 0, ctx->stream>>>((const srd::ScanProg*)s->prog.p, vt, n, s->sel.as<uint8_t>(),
                                                                      want_counts ? s->block_counts.as<uint32_t>() : nullptr);
        SR_LAUNCH_CHECK(ctx);
    }
    if (want_counts) {
        if (n > 0) {
            srd::k_scan_counts<<<1, 1024, 0, ctx->stream>>>(s->block_counts.as<uint32_t>(), tiles, s->tile_offsets.as<uint64_t>(),
                                                           ctx->dscratch);
            SR_LAUNCH_CHECK(ctx);
            SR_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, ctx->dscratch, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
            SR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            *total_out = (int64_t)ctx->pinned[0];
        } else {
            *total_out = 0;
        }
    }
    return SR_OK;
}
