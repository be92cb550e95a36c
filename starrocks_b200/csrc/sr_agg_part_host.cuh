// sr_agg_part_host.cuh -- host side of the radix-partitioned aggregate push (kernels: sr_agg_part.cuh).
#pragma once

static const int64_t kPartitionedMinRows = 1 << 20;   // smaller batches take the direct (global atomics) push
static const int64_t kPartitionedMaxRows = 1ll << 30; // staged records per round
static const uint64_t kPartitionSliceBytes = 32ull << 20; // L2 mode: table bytes one bucket maps to

template <int W, int MODE>
static int32_t aggp_launch_scatter(sr_ctx* ctx, int grid, const srd::AggDev* dev, const VTab& vt, const srd::PartPlan& pl, const srd::ScatterArgs& sa) {
    const size_t smem = srd::aggp_scatter_smem<W>();
    SR_CUDA(ctx, cudaFuncSetAttribute(srd::k_aggp_scatter<W, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    srd::k_aggp_scatter<W, MODE><<<grid, srd::AGGP_BLOCK, smem, ctx->stream>>>(dev, vt, pl, sa);
    SR_LAUNCH_CHECK(ctx);
    return SR_OK;
}

static int32_t aggp_scatter(sr_ctx* ctx, bool from_chunk, int grid, const srd::AggDev* dev, const VTab& vt, const srd::PartPlan& pl, const srd::ScatterArgs& sa) {
    if (from_chunk && pl.simple) {
        switch (pl.words) { // (plans of more than 4 words take the general kernel)
        case 1:
            return aggp_launch_scatter<1, srd::SCATTER_CHUNK_SIMPLE>(ctx, grid, dev, vt, pl, sa);
        case 2:
            return aggp_launch_scatter<2, srd::SCATTER_CHUNK_SIMPLE>(ctx, grid, dev, vt, pl, sa);
        case 3:
            return aggp_launch_scatter<3, srd::SCATTER_CHUNK_SIMPLE>(ctx, grid, dev, vt, pl, sa);
        case 4:
            return aggp_launch_scatter<4, srd::SCATTER_CHUNK_SIMPLE>(ctx, grid, dev, vt, pl, sa);
        default:
            break;
        }
    }
    switch (pl.words) {
#define SR_AGGP_CASE(W)                                                                                   \
    case W:                                                                                               \
        return from_chunk ? aggp_launch_scatter<W, srd::SCATTER_CHUNK>(ctx, grid, dev, vt, pl, sa)        \
                          : aggp_launch_scatter<W, srd::SCATTER_RECORDS>(ctx, grid, dev, vt, pl, sa);
        SR_AGGP_CASE(1)
        SR_AGGP_CASE(2)
        SR_AGGP_CASE(3)
        SR_AGGP_CASE(4)
        SR_AGGP_CASE(5)
        SR_AGGP_CASE(6)
        SR_AGGP_CASE(7)
        SR_AGGP_CASE(8)
        SR_AGGP_CASE(9)
        SR_AGGP_CASE(10)
        SR_AGGP_CASE(11)
#undef SR_AGGP_CASE
    default:
        return sr_fail(ctx, SR_ERR_NOT_SUPPORTED, "record of %d words", pl.words);
    }
}

// shared memory one slice of the table needs in k_aggp_apply_smem
static size_t aggp_slice_smem(const srd::AggDev& h) { return ((size_t)1 << h.slice_log2) * agg_slot_bytes(h); }

// Rows [0, n) of the bound chunk: scatter by the top bits of the home slot, then apply bucket by bucket.
static int32_t agg_push_partitioned(sr_agg* a, const VTab& vt, int64_t n, bool fresh) {
    sr_ctx* ctx = a->ctx;
    int64_t done = 0;
    const bool trace = getenv("SR_AGG_TRACE") != nullptr; // phase times (CUDA events) on stderr
    cudaEvent_t tev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (trace)
        for (auto& e : tev) SR_CUDA(ctx, cudaEventCreate(&e));
    while (done < n) {
        const srd::AggDev& h = a->host;
        const int64_t m = std::min<int64_t>(n - done, kPartitionedMaxRows); // bounded by the staging buffers
        int log2cap = 0;
        while ((1ull << log2cap) < h.cap) log2cap++;
        srd::PartPlan pl;
        memset(&pl, 0, sizeof(pl));
        // record layout
        int w = 0;
        pl.word_kind[w++] = srd::WK_KEY_LO;
        if (h.wide) pl.word_kind[w++] = srd::WK_KEY_HI;
        bool any_nullable = false;
        for (int f = 0; f < h.num_fns; f++) {
            pl.val_word[f] = -1;
            const srd::AggFnDev& fn = h.fns[f];
            if (fn.mode == srd::M_COUNT_STAR) continue;
            pl.word_kind[w] = srd::WK_VALUE;
            pl.word_fn[w] = f;
            pl.val_word[f] = w++;
            for (int k = 0; k < fn.input.num_nodes; k++)
                if ((fn.input.nodes[k].op == srd::C_LOAD_I || fn.input.nodes[k].op == srd::C_LOAD_D) && vt.v[fn.input.nodes[k].arg].nulls) any_nullable = true;
        }
        pl.null_word = -1;
        if (any_nullable) {
            pl.word_kind[w] = srd::WK_NULLS;
            pl.null_word = w++;
        }
        pl.words = w;
        // SIMPLE plan?
        pl.simple = (h.num_keys == 1 && !h.wide && !h.key_nullable[0] && !any_nullable && w <= 4) ? 1 : 0;
        if (pl.simple) {
            pl.word_vid[0] = h.key_value_id[0];
            pl.simple_key_mask = h.key_width[0] == 8 ? ~0ull : ((1ull << (8 * h.key_width[0])) - 1);
            if (vt.v[h.key_value_id[0]].nulls) pl.simple = 0;
            for (int f = 0; f < h.num_fns; f++) {
                if (pl.val_word[f] < 0) continue;
                if (h.fns[f].input.form != srd::F_COL) pl.simple = 0;
                pl.word_vid[pl.val_word[f]] = h.fns[f].input.nodes[0].arg;
            }
        }
        // buckets: one per slice when the slices of the whole table can be counted in one shared-memory histogram
        const size_t slice_smem = aggp_slice_smem(h);
        const bool smem_mode = log2cap - h.slice_log2 <= srd::AGGP_MAX_BITS && slice_smem <= 100 * 1024 && !getenv("SR_AGG_PARTITION_FORCE_L2");
        if (smem_mode) {
            pl.bits = log2cap - h.slice_log2;
            pl.bucket_shift = h.slice_log2;
        } else {
            const uint64_t table_bytes = (h.cap + 1) * agg_slot_bytes(h);
            int log2p = 1;
            while (log2p < srd::AGGP_ONE_LEVEL_BITS && (table_bytes >> log2p) > kPartitionSliceBytes) log2p++;
            if (log2p > log2cap) log2p = log2cap;
            pl.bits = log2p;
            pl.bucket_shift = log2cap - log2p;
        }
        pl.bits2 = pl.bits > srd::AGGP_ONE_LEVEL_BITS ? pl.bits / 2 : 0;
        const int P = 1 << pl.bits;
        const int F1 = 1 << (pl.bits - pl.bits2);
        const int tile_rows = srd::AGGP_BLOCK * srd::aggp_rows_per_thread(pl.words);
        const size_t rec_bytes = (size_t)m * pl.words * 8;
        SR_TRY(a->part_hist.reserve(ctx, sizeof(uint32_t) * (size_t)P));
        SR_TRY(a->part_base.reserve(ctx, sizeof(uint64_t) * ((size_t)P + 1)));
        SR_TRY(a->part_cursor.reserve(ctx, sizeof(uint64_t) * ((size_t)P + F1)));
        SR_TRY(a->part_tiles.reserve(ctx, sizeof(uint32_t) * ((size_t)F1 + 1)));
        SR_TRY(a->part_rec[0].reserve(ctx, rec_bytes + 16));
        if (pl.bits2) SR_TRY(a->part_rec[1].reserve(ctx, rec_bytes + 16));
        unsigned long long* cursor = a->part_cursor.as<unsigned long long>();
        unsigned long long* cursor1 = cursor + P;
        unsigned long long* rec_final = a->part_rec[0].as<unsigned long long>();
        unsigned long long* rec_mid = pl.bits2 ? a->part_rec[1].as<unsigned long long>() : nullptr;
        const srd::AggDev* dev = (const srd::AggDev*)a->dev.p;
        const int grid2 = ctx->num_sms * 2;
        if (trace) SR_CUDA(ctx, cudaEventRecord(tev[0], ctx->stream));
        SR_CUDA(ctx, cudaMemsetAsync(a->part_hist.p, 0, sizeof(uint32_t) * (size_t)P, ctx->stream));
        {
            const size_t hsm = sizeof(uint32_t) * (size_t)P;
            SR_CUDA(ctx, cudaFuncSetAttribute(srd::k_aggp_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(uint32_t) << srd::AGGP_MAX_BITS)));
            const int hgrid = ctx->num_sms * (hsm > 64 * 1024 ? 1 : 2);
            srd::k_aggp_hist<<<hgrid, srd::AGGP_BLOCK, hsm, ctx->stream>>>(dev, vt, pl, done, m, a->part_hist.as<uint32_t>());
            SR_LAUNCH_CHECK(ctx);
        }
        if (trace) SR_CUDA(ctx, cudaEventRecord(tev[1], ctx->stream));
        srd::k_aggp_prepare<<<1, 1024, 0, ctx->stream>>>(a->part_hist.as<uint32_t>(), pl.bits, pl.bits2, tile_rows, a->part_base.as<uint64_t>(), cursor, cursor1,
                                                          a->part_tiles.as<uint32_t>());
        SR_LAUNCH_CHECK(ctx);
        srd::ScatterArgs sa;
        memset(&sa, 0, sizeof(sa));
        sa.row_base = done;
        sa.n = m;
        sa.base = a->part_base.as<uint64_t>();
        sa.tile_start = a->part_tiles.as<uint32_t>();
        if (pl.bits2 == 0) {
            sa.cursor = cursor;
            sa.dst = rec_final;
            sa.fan_bits = pl.bits;
            sa.local_shift = 0;
            SR_TRY(aggp_scatter(ctx, true, grid2, dev, vt, pl, sa));
            if (trace) SR_CUDA(ctx, cudaEventRecord(tev[2], ctx->stream));
        } else {
            sa.cursor = cursor1;
            sa.dst = rec_mid;
            sa.fan_bits = pl.bits - pl.bits2;
            sa.local_shift = pl.bits2;
            SR_TRY(aggp_scatter(ctx, true, grid2, dev, vt, pl, sa));
            if (trace) SR_CUDA(ctx, cudaEventRecord(tev[2], ctx->stream));
            sa.src = rec_mid;
            sa.cursor = cursor;
            sa.dst = rec_final;
            sa.fan_bits = pl.bits2;
            sa.local_shift = 0;
            SR_TRY(aggp_scatter(ctx, false, grid2, dev, vt, pl, sa));
        }
        if (trace) SR_CUDA(ctx, cudaEventRecord(tev[3], ctx->stream));
        unsigned long long* fail_count = a->counters.as<unsigned long long>() + 4;
        SR_CUDA(ctx, cudaMemsetAsync(fail_count, 0, 8, ctx->stream));
        std::vector<uint64_t> bounds;
        uint64_t failed_buckets = 0;
        if (smem_mode) {
            SR_TRY(a->part_fail.reserve(ctx, sizeof(uint32_t) * (size_t)P));
            srd::ApplyArgs aa;
            aa.rec = rec_final;
            aa.base = a->part_base.as<uint64_t>();
            aa.num_buckets = (uint32_t)P;
            aa.fresh = fresh ? 1 : 0;
            fresh = false;
            aa.fail_list = a->part_fail.as<uint32_t>();
            aa.fail_count = fail_count;
            SR_CUDA(ctx, cudaFuncSetAttribute(srd::k_aggp_apply_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)slice_smem));
            srd::k_aggp_apply_smem<<<std::min(P, grid2), srd::AGGP_BLOCK, slice_smem, ctx->stream>>>(dev, pl, aa);
            SR_LAUNCH_CHECK(ctx);
            if (trace) SR_CUDA(ctx, cudaEventRecord(tev[4], ctx->stream));
            SR_CUDA(ctx, cudaMemcpyAsync(ctx->pinned + 8, a->counters.p, 48, cudaMemcpyDeviceToHost, ctx->stream));
            SR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            a->ngroups_host = (int64_t)ctx->pinned[8];
            failed_buckets = ctx->pinned[8 + 4];
        }
        if (trace && smem_mode) {
            float t[4];
            for (int k = 0; k < 4; k++) cudaEventElapsedTime(&t[k], tev[k], tev[k + 1]);
            fprintf(stderr, "[sr_agg partitioned push] rows %lld, %d-word records, 2^%d buckets (%d + %d bits), slices of %d slots: histogram %.3f ms, scatter %.3f ms, "
                            "scatter-2 %.3f ms, apply %.3f ms, %llu slices overflowed\n",
                    (long long)m, pl.words, pl.bits, pl.bits - pl.bits2, pl.bits2, 1 << h.slice_log2, t[0], t[1], t[2], t[3], (unsigned long long)failed_buckets);
        }
        if (!smem_mode || failed_buckets > 0) {
            // global-atomics apply: every bucket (L2 mode) or the buckets whose slice overflowed, after a growth
            bounds.resize((size_t)P + 1);
            SR_CUDA(ctx, cudaMemcpyAsync(bounds.data(), a->part_base.p, sizeof(uint64_t) * ((size_t)P + 1), cudaMemcpyDeviceToHost, ctx->stream));
            std::vector<uint32_t> todo;
            if (smem_mode) {
                todo.resize((size_t)failed_buckets);
                SR_CUDA(ctx, cudaMemcpyAsync(todo.data(), a->part_fail.p, sizeof(uint32_t) * (size_t)failed_buckets, cudaMemcpyDeviceToHost, ctx->stream));
            }
            SR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (smem_mode) {
                if (a->host.cap >= (1ull << 33)) return sr_fail(ctx, SR_ERR_OUT_OF_MEMORY, "aggregate table would exceed 2^33 slots");
                SR_TRY(agg_grow(a, a->host.cap * 2));
            } else {
                todo.resize((size_t)P);
                for (int b = 0; b < P; b++) todo[(size_t)b] = (uint32_t)b;
            }
            SR_TRY(a->part_fail64.reserve(ctx, sizeof(uint64_t) * (size_t)m));
            SR_CUDA(ctx, cudaMemsetAsync(fail_count, 0, 8, ctx->stream));
            for (uint32_t b : todo) {
                const int64_t r0 = (int64_t)bounds[b], r1 = (int64_t)bounds[(size_t)b + 1];
                if (r1 <= r0) continue;
                const int agrid = std::min(grid_for(r1 - r0, srd::AGG_BLOCK), ctx->num_sms * 8);
                // the slot range to prefetch is only known in L2 mode (after a growth a bucket maps to several ranges)
                const unsigned long long s_lo = smem_mode ? 0ull : (unsigned long long)b << pl.bucket_shift;
                const unsigned long long s_hi = smem_mode ? 0ull : (unsigned long long)(b + 1) << pl.bucket_shift;
                srd::k_aggp_apply_l2<<<agrid, srd::AGG_BLOCK, 0, ctx->stream>>>((const srd::AggDev*)a->dev.p, pl, rec_final, r0, r1, nullptr, s_lo, s_hi,
                                                                                 a->part_fail64.as<uint64_t>(), fail_count);
                SR_LAUNCH_CHECK(ctx);
            }
            fresh = false;
            // records refused by the admission limit / a full slice: grow, re-apply them
            while (true) {
                SR_CUDA(ctx, cudaMemcpyAsync(ctx->pinned + 8, a->counters.p, 48, cudaMemcpyDeviceToHost, ctx->stream));
                SR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                a->ngroups_host = (int64_t)ctx->pinned[8];
                const uint64_t failed = ctx->pinned[8 + 4];
                if (failed == 0) break;
                if (a->host.cap >= (1ull << 33)) return sr_fail(ctx, SR_ERR_OUT_OF_MEMORY, "aggregate table would exceed 2^33 slots");
                SR_CUDA(ctx, cudaMemsetAsync((uint8_t*)a->counters.p + 8, 0, 8, ctx->stream)); // overflow / range flags
                SR_CUDA(ctx, cudaMemsetAsync(fail_count, 0, 8, ctx->stream));
                SR_TRY(agg_grow(a, a->host.cap * 2));
                SR_TRY(a->part_fail64b.reserve(ctx, sizeof(uint64_t) * (size_t)failed));
                const int agrid = std::min(grid_for((int64_t)failed, srd::AGG_BLOCK), ctx->num_sms * 8);
                srd::k_aggp_apply_l2<<<agrid, srd::AGG_BLOCK, 0, ctx->stream>>>((const srd::AggDev*)a->dev.p, pl, rec_final, 0, (int64_t)failed, a->part_fail64.as<uint64_t>(), 0, 0,
                                                                                 a->part_fail64b.as<uint64_t>(), fail_count);
                SR_LAUNCH_CHECK(ctx);
                std::swap(a->part_fail64, a->part_fail64b);
            }
        }
        // keep the load below the admission limit for the pushes that follow
        while ((uint64_t)a->ngroups_host > a->host.limit) {
            if (a->host.cap >= (1ull << 33)) return sr_fail(ctx, SR_ERR_OUT_OF_MEMORY, "aggregate table would exceed 2^33 slots");
            SR_TRY(agg_grow(a, a->host.cap * 2));
        }
        a->partitioned_pushes++;
        done += m;
    }
    if (trace)
        for (auto& e : tev) cudaEventDestroy(e);
    return SR_OK;
}
