// Drives the GPU operators through the reference's pipeline protocol (PipelineDriver::process pull/push/finishing,
// dependency of the probe on the build) on an SSB Q4.1-shaped plan, the way be/test/exec/pipeline/pipeline_test_base
// assembles OpFactories without an FE:
//
//   build pipelines (x4):  GpuScanOperator(dimension, predicate) -> GpuHashJoinBuildOperator
//   probe pipeline  (A):   GpuScanOperator(lineorder) -> GpuHashJoinProbeOperator x4 -> GpuAggregateBlockingSinkOperator
//   probe pipeline  (B):   GpuScanOperator(lineorder) -> GpuFragmentSinkOperator            (fused form)
//   probe pipeline  (C):   A, with every build operator publishing a runtime filter (min/max + bloom) through the
//                          RuntimeFilterHub and the lineorder scan consuming them (local_rf_block until they are ready)
//   result pipeline:       GpuAggregateBlockingSourceOperator -> ResultSink (collects rows)
//
// A, B and C must give the groups of the CPU ORACLE (oracle/sr_oracle.cpp: the restatement of the reference's operators)
// run on the same chunks -- dimension scans + join builds + the chunk-at-a-time probe / aggregate pipeline
// (orc_fragment_run); a row-at-a-time evaluation of the query in this file cross-checks the oracle.  Exit code 0 = pass.  Needs a CUDA device (run by tests/test_host_pipeline.py -m gpu).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <map>
#include <thread>
#include <tuple>

#include "../../oracle/sr_oracle.h"
#include "../../starrocks_b200/host/gpu/gpu_operators.h"

using namespace starrocks;
using namespace starrocks::pipeline;

namespace {

enum Slots { LO_ORDERDATE = 0, LO_CUSTKEY, LO_SUPPKEY, LO_PARTKEY, LO_REVENUE, LO_SUPPLYCOST, C_CUSTKEY = 10, C_REGION, C_NATION, S_SUPPKEY = 20, S_REGION,
             P_PARTKEY = 30, P_MFGR, D_DATEKEY = 40, D_YEAR, OUT_REV = 50, OUT_COST };

struct Rng {
    uint64_t s;
    uint32_t next(uint32_t lo, uint32_t hi) { // [lo, hi)
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        return lo + (uint32_t)((s >> 33) % (hi - lo));
    }
};

ChunkPtr make_chunk(const std::vector<std::pair<SlotId, std::vector<int32_t>>>& cols) {
    auto c = std::make_shared<Chunk>();
    for (auto& [slot, v] : cols) c->append_column(std::make_shared<Int32Column>(SR_TYPE_INT, v), slot);
    return c;
}

std::vector<ChunkPtr> split(const ChunkPtr& whole, size_t chunk_size) {
    std::vector<ChunkPtr> out;
    for (size_t off = 0; off < whole->num_rows(); off += chunk_size) out.push_back(whole->slice(off, std::min(chunk_size, whole->num_rows() - off)));
    return out;
}

class ResultSink final : public Operator {
public:
    ResultSink() : Operator(nullptr, 99, "result_sink", 99, false, 0) {}
    bool has_output() const override { return false; }
    bool need_input() const override { return !_finished; }
    bool is_finished() const override { return _finished; }
    StatusOr<ChunkPtr> pull_chunk(RuntimeState*) override { return Status::InternalError("sink"); }
    Status push_chunk(RuntimeState*, const ChunkPtr& c) override {
        chunks.push_back(c);
        return Status::OK();
    }
    Status set_finishing(RuntimeState*) override {
        _finished = true;
        return Status::OK();
    }
    std::vector<ChunkPtr> chunks;

private:
    bool _finished = false;
};

#define CHECK_OK(expr)                                                              \
    do {                                                                            \
        Status _st = (expr);                                                        \
        if (!_st.ok()) {                                                            \
            fprintf(stderr, "FAILED %s: %s\n", #expr, _st.to_string().c_str());     \
            exit(2);                                                                \
        }                                                                           \
    } while (0)

void run_to_finish(PipelineDriver& d, RuntimeState* state, const char* name) {
    CHECK_OK(d.prepare(state));
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(300);
    while (std::chrono::steady_clock::now() < deadline) {
        auto st = d.process(state);
        if (!st.ok()) {
            fprintf(stderr, "driver %s failed: %s\n", name, st.status().to_string().c_str());
            exit(2);
        }
        if (st.value() == PipelineDriver::FINISH) return;
        if (st.value() == PipelineDriver::PRECONDITION_BLOCK) {
            fprintf(stderr, "driver %s is still blocked on its dependency\n", name);
            exit(2);
        }
        // READY without progress (the asynchronous scan's IO task is still running) or PENDING_FINISH: the real driver parks
        // in the poller (pipeline_driver_poller.cpp); here the thread naps
        std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
    fprintf(stderr, "driver %s made no progress\n", name);
    exit(2);
}

using Groups = std::map<std::pair<int32_t, int32_t>, std::pair<int64_t, int64_t>>;

Groups collect(const ResultSink& sink) {
    Groups g;
    for (auto& c : sink.chunks) {
        auto* y = (const int32_t*)c->get_column_by_slot_id(D_YEAR)->raw_data();
        auto* n = (const int32_t*)c->get_column_by_slot_id(C_NATION)->raw_data();
        auto* r = (const int64_t*)c->get_column_by_slot_id(OUT_REV)->raw_data();
        auto* s = (const int64_t*)c->get_column_by_slot_id(OUT_COST)->raw_data();
        for (size_t i = 0; i < c->num_rows(); i++) g[{y[i], n[i]}] = {r[i], s[i]};
    }
    return g;
}

sr_expr col_expr(int32_t slot) {
    sr_expr e{};
    e.nodes[0].op = SR_EX_COL;
    e.nodes[0].slot_id = slot;
    e.num_nodes = 1;
    return e;
}

} // namespace

int main(int argc, char** argv) {
    const size_t n_fact = argc > 1 ? (size_t)atoll(argv[1]) : 1000000;
    const int n_cust = 30000, n_supp = 2000, n_part = 20000, n_dates = 2556;
    sr_ctx* ctx = sr_ctx_create(0, nullptr);
    if (!ctx) {
        fprintf(stderr, "sr_ctx_create failed: %s\n", sr_last_error(nullptr));
        return 3;
    }
    RuntimeState state(4096);
    Rng rng{20240921};

    // ---- synthetic SSB-shaped tables ----
    std::vector<int32_t> c_key(n_cust), c_region(n_cust), c_nation(n_cust), s_key(n_supp), s_region(n_supp), p_key(n_part), p_mfgr(n_part), d_key(n_dates),
            d_year(n_dates);
    for (int i = 0; i < n_cust; i++) {
        c_key[i] = i + 1;
        c_region[i] = rng.next(0, 5);
        c_nation[i] = c_region[i] * 5 + rng.next(0, 5);
    }
    for (int i = 0; i < n_supp; i++) {
        s_key[i] = i + 1;
        s_region[i] = rng.next(0, 5);
    }
    for (int i = 0; i < n_part; i++) {
        p_key[i] = i + 1;
        p_mfgr[i] = rng.next(0, 5);
    }
    for (int i = 0; i < n_dates; i++) {
        d_year[i] = 1992 + i / 366;
        d_key[i] = d_year[i] * 10000 + (i % 366) + 101; // unique, sparse like yyyymmdd
    }
    std::vector<int32_t> lo_date(n_fact), lo_cust(n_fact), lo_supp(n_fact), lo_part(n_fact), lo_rev(n_fact), lo_cost(n_fact);
    for (size_t i = 0; i < n_fact; i++) {
        lo_date[i] = d_key[rng.next(0, n_dates)];
        lo_cust[i] = rng.next(1, n_cust + 1);
        lo_supp[i] = rng.next(1, n_supp + 1);
        lo_part[i] = rng.next(1, n_part + 1);
        lo_rev[i] = rng.next(81000, 10400001);
        lo_cost[i] = rng.next(54000, 125001);
    }
    // ---- the checker: row-at-a-time evaluation of the query ----
    Groups expect;
    for (size_t i = 0; i < n_fact; i++) {
        if (s_region[lo_supp[i] - 1] != 1 || c_region[lo_cust[i] - 1] != 1 || p_mfgr[lo_part[i] - 1] > 1) continue;
        const int32_t year = lo_date[i] / 10000;
        auto& g = expect[{year, c_nation[lo_cust[i] - 1]}];
        g.first += lo_rev[i];
        g.second += lo_cost[i];
    }

    // ---- descriptors ----
    auto eq_pred = [](int32_t slot, int64_t v) {
        sr_pred p{};
        p.slot_id = slot;
        p.op = SR_PRED_EQ;
        p.ilo = v;
        return p;
    };
    sr_pred supp_pred = eq_pred(S_REGION, 1), cust_pred = eq_pred(C_REGION, 1), part_pred{};
    part_pred.slot_id = P_MFGR;
    part_pred.op = SR_PRED_IN;
    part_pred.in_list[0] = 0;
    part_pred.in_list[1] = 1;
    part_pred.in_count = 2;
    struct Dim {
        const char* name;
        ChunkPtr table;
        sr_pred* pred;
        int32_t key_slot, probe_slot;
        std::vector<int32_t> payload;
    };
    std::vector<Dim> dims = {
            {"supplier", make_chunk({{S_SUPPKEY, s_key}, {S_REGION, s_region}}), &supp_pred, S_SUPPKEY, LO_SUPPKEY, {}},
            {"customer", make_chunk({{C_CUSTKEY, c_key}, {C_REGION, c_region}, {C_NATION, c_nation}}), &cust_pred, C_CUSTKEY, LO_CUSTKEY, {C_NATION}},
            {"part", make_chunk({{P_PARTKEY, p_key}, {P_MFGR, p_mfgr}}), &part_pred, P_PARTKEY, LO_PARTKEY, {}},
            {"dates", make_chunk({{D_DATEKEY, d_key}, {D_YEAR, d_year}}), nullptr, D_DATEKEY, LO_ORDERDATE, {D_YEAR}},
    };
    ChunkPtr lineorder = make_chunk({{LO_ORDERDATE, lo_date}, {LO_CUSTKEY, lo_cust}, {LO_SUPPKEY, lo_supp}, {LO_PARTKEY, lo_part}, {LO_REVENUE, lo_rev},
                                     {LO_SUPPLYCOST, lo_cost}});
    sr_agg_desc agg_desc{};
    agg_desc.num_group_keys = 2;
    agg_desc.group_slots[0] = D_YEAR;
    agg_desc.group_slots[1] = C_NATION;
    agg_desc.group_types[0] = agg_desc.group_types[1] = SR_TYPE_INT;
    agg_desc.has_ranges = 1;
    agg_desc.group_min[0] = 1992;
    agg_desc.group_max[0] = 1998;
    agg_desc.group_min[1] = 0;
    agg_desc.group_max[1] = 24;
    agg_desc.num_fns = 2;
    agg_desc.fns[0] = sr_agg_fn{SR_AGG_SUM, SR_TYPE_INT, OUT_REV, 0, col_expr(LO_REVENUE)};
    agg_desc.fns[1] = sr_agg_fn{SR_AGG_SUM, SR_TYPE_INT, OUT_COST, 0, col_expr(LO_SUPPLYCOST)};

    // ---- the oracle: the same plan through the CPU restatement of the reference's operators ----
    Groups oracle_groups;
    {
        std::vector<orc_join*> ojoins;
        orc_fragment_desc ofd{};
        ofd.num_joins = (int32_t)dims.size();
        ofd.agg = agg_desc;
        for (size_t k = 0; k < dims.size(); k++) {
            Dim& dm = dims[k];
            std::vector<int32_t> outs = {dm.key_slot};
            for (int32_t p : dm.payload) outs.push_back(p);
            sr_scan_desc sd{};
            sd.preds = dm.pred;
            sd.num_preds = dm.pred ? 1 : 0;
            sd.out_slots = outs.data();
            sd.num_out_slots = (int32_t)outs.size();
            std::vector<sr_col_view> views;
            sr_chunk_view tv = make_chunk_view(*dm.table, &views);
            std::vector<std::vector<int32_t>> kept(outs.size(), std::vector<int32_t>(dm.table->num_rows()));
            std::vector<void*> od(outs.size());
            std::vector<uint8_t*> on(outs.size(), nullptr);
            for (size_t c = 0; c < outs.size(); c++) od[c] = kept[c].data();
            const int64_t rows = orc_scan_filter(&sd, &tv, od.data(), on.data());
            if (rows < 0) {
                fprintf(stderr, "oracle scan failed: %s\n", orc_last_error());
                return 2;
            }
            sr_join_desc jd{};
            jd.join_type = SR_JOIN_INNER;
            jd.num_keys = 1;
            jd.build_key_slots[0] = dm.key_slot;
            jd.probe_key_slots[0] = dm.probe_slot;
            jd.key_types[0] = SR_TYPE_INT;
            jd.enable_range_direct_mapping = 1;
            jd.num_build_out = (int32_t)dm.payload.size();
            for (size_t p = 0; p < dm.payload.size(); p++) jd.build_out_slots[p] = dm.payload[p];
            orc_join_options opt{1, 1, 1 << 20, 32 << 20, 0, 0};
            orc_join* oj = orc_join_create(&jd, &opt);
            std::vector<sr_col_view> bcols;
            for (size_t c = 0; c < outs.size(); c++) bcols.push_back(sr_col_view{kept[c].data(), nullptr, SR_TYPE_INT, outs[c]});
            sr_chunk_view bv{bcols.data(), (int32_t)bcols.size(), SR_MEM_HOST, rows};
            if (!oj || orc_join_append_build(oj, &bv) != SR_OK || orc_join_build(oj) != SR_OK) {
                fprintf(stderr, "oracle join build failed: %s\n", orc_last_error());
                return 2;
            }
            ojoins.push_back(oj);
            ofd.joins[k].join = oj;
            ofd.joins[k].probe_key_slot = dm.probe_slot;
            ofd.joins[k].num_payload = (int32_t)dm.payload.size();
            for (size_t p = 0; p < dm.payload.size(); p++) ofd.joins[k].payload_build_slots[p] = dm.payload[p];
        }
        std::vector<sr_col_view> fviews;
        sr_chunk_view fv = make_chunk_view(*lineorder, &fviews);
        orc_agg* oresult = orc_agg_create(&agg_desc);
        int64_t passed = 0;
        if (!oresult || orc_fragment_run(&ofd, &fv, 4, oresult, &passed) != SR_OK) {
            fprintf(stderr, "oracle fragment run failed: %s\n", orc_last_error());
            return 2;
        }
        const int64_t g = orc_agg_num_groups(oresult);
        std::vector<int32_t> oy((size_t)g), on_((size_t)g);
        std::vector<int64_t> orv((size_t)g), oc((size_t)g);
        std::vector<uint8_t> nul0((size_t)g), nul1((size_t)g), nul2((size_t)g), nul3((size_t)g);
        void* od[4] = {oy.data(), on_.data(), orv.data(), oc.data()};
        uint8_t* onl[4] = {nul0.data(), nul1.data(), nul2.data(), nul3.data()};
        if (orc_agg_output(oresult, od, onl) != SR_OK) {
            fprintf(stderr, "oracle output failed: %s\n", orc_last_error());
            return 2;
        }
        for (int64_t i = 0; i < g; i++) oracle_groups[{oy[(size_t)i], on_[(size_t)i]}] = {orv[(size_t)i], oc[(size_t)i]};
        orc_agg_destroy(oresult);
        for (auto* oj : ojoins) orc_join_destroy(oj);
        if (oracle_groups != expect) {
            fprintf(stderr, "the oracle (%zu groups) and the row-at-a-time evaluation (%zu groups) disagree\n", oracle_groups.size(), expect.size());
            return 2;
        }
        printf("oracle: %zu groups from %lld joined rows (equal to the row-at-a-time evaluation)\n", oracle_groups.size(), (long long)passed);
    }

    Groups results[3];
    size_t rows_after_scan[3] = {0, 0, 0};
    for (int variant = 0; variant < 3; variant++) {
        const bool fused = variant == 1, with_rf = variant == 2;
        RuntimeFilterHub hub;
        std::vector<std::pair<OperatorPtr, OperatorPtr>> build_ops;
        // ---- build pipelines: scan(dim) -> hash join build; the joiner is shared with the probe side ----
        std::vector<GpuHashJoinerFactoryPtr> joiner_factories;
        std::vector<std::vector<int32_t>> out_slot_store(dims.size());
        // slots still needed downstream of join k on the probe side
        std::vector<std::vector<int32_t>> probe_out = {{LO_ORDERDATE, LO_CUSTKEY, LO_PARTKEY, LO_REVENUE, LO_SUPPLYCOST},
                                                       {LO_ORDERDATE, LO_PARTKEY, LO_REVENUE, LO_SUPPLYCOST},
                                                       {LO_ORDERDATE, LO_REVENUE, LO_SUPPLYCOST, C_NATION},
                                                       {LO_REVENUE, LO_SUPPLYCOST, C_NATION}};
        for (size_t k = 0; k < dims.size(); k++) {
            Dim& dm = dims[k];
            out_slot_store[k] = {dm.key_slot};
            for (int32_t p : dm.payload) out_slot_store[k].push_back(p);
            sr_scan_desc sd{};
            sd.preds = dm.pred;
            sd.num_preds = dm.pred ? 1 : 0;
            sd.out_slots = out_slot_store[k].data();
            sd.num_out_slots = (int32_t)out_slot_store[k].size();
            sr_join_desc jd{};
            jd.join_type = SR_JOIN_INNER;
            jd.num_keys = 1;
            jd.build_key_slots[0] = dm.key_slot;
            jd.probe_key_slots[0] = dm.probe_slot;
            jd.key_types[0] = SR_TYPE_INT;
            jd.enable_range_direct_mapping = 1;
            jd.num_build_out = (int32_t)dm.payload.size();
            for (size_t p = 0; p < dm.payload.size(); p++) jd.build_out_slots[p] = dm.payload[p];
            jd.num_probe_out = (int32_t)probe_out[k].size();
            for (size_t p = 0; p < probe_out[k].size(); p++) jd.probe_out_slots[p] = probe_out[k][p];
            auto jf = std::make_shared<GpuHashJoinerFactory>(ctx, jd);
            joiner_factories.push_back(jf);
            GpuScanOperatorFactory scan_f(1, 1, ctx, sd, {split(dm.table, 4096)});
            GpuHashJoinBuildOperatorFactory build_f(2, 20 + (int)k, jf);
            build_ops.emplace_back(scan_f.create(1, 0), build_f.create(1, 0));
            if (with_rf) {
                hub.add_holder(20 + (int)k);
                static_cast<GpuHashJoinBuildOperator*>(build_ops.back().second.get())
                        ->set_runtime_filters(&hub, {GpuRuntimeFilterBuildDesc{(int32_t)k, 0, true, false}});
            }
        }
        auto run_builds = [&]() {
            for (size_t k = 0; k < dims.size(); k++) {
                PipelineDriver build_driver({build_ops[k].first, build_ops[k].second});
                run_to_finish(build_driver, &state, dims[k].name);
                build_driver.close(&state);
                printf("    %s | %s\n", build_ops[k].second->runtime_profile()->name().c_str(), build_ops[k].second->unique_metrics()->to_string().c_str());
            }
        };
        if (!with_rf) run_builds();
        // ---- probe pipeline ----
        sr_scan_desc fact_scan{};
        std::vector<int32_t> fact_out = {LO_ORDERDATE, LO_CUSTKEY, LO_SUPPKEY, LO_PARTKEY, LO_REVENUE, LO_SUPPLYCOST};
        fact_scan.out_slots = fact_out.data();
        fact_scan.num_out_slots = (int32_t)fact_out.size();
        GpuScanOperatorFactory fact_f(3, 3, ctx, fact_scan, {split(lineorder, 4096)});
        GpuAggregatorPtr aggregator;
        GpuFragmentPtr fragment;
        Operators ops = {fact_f.create(1, 0)};
        auto* fact_scan_op = static_cast<GpuScanOperator*>(ops[0].get());
        if (with_rf) {
            std::vector<GpuRuntimeFilterProbeDesc> probes;
            for (size_t k = 0; k < dims.size(); k++) probes.push_back(GpuRuntimeFilterProbeDesc{(int32_t)k, 20 + (int32_t)k, dims[k].probe_slot});
            fact_scan_op->set_runtime_filters(&hub, probes);
        }
        GpuAggregatorFactoryPtr agg_f;
        if (!fused) {
            for (size_t k = 0; k < dims.size(); k++) {
                GpuHashJoinProbeOperatorFactory pf(4 + (int)k, 4 + (int)k, joiner_factories[k]);
                ops.push_back(pf.create(1, 0));
            }
            agg_f = std::make_shared<GpuAggregatorFactory>(ctx, agg_desc);
            GpuAggregateBlockingSinkOperatorFactory sink_f(8, 8, agg_f);
            ops.push_back(sink_f.create(1, 0));
            aggregator = agg_f->get_or_create(0);
        } else {
            sr_fragment_desc fd{};
            fd.num_joins = (int32_t)dims.size();
            std::vector<GpuHashJoinerPtr> joiners;
            for (size_t k = 0; k < dims.size(); k++) {
                fd.joins[k].probe_key_slot = dims[k].probe_slot;
                fd.joins[k].num_payload = (int32_t)dims[k].payload.size();
                for (size_t p = 0; p < dims[k].payload.size(); p++) fd.joins[k].payload_build_slots[p] = dims[k].payload[p];
                joiners.push_back(joiner_factories[k]->get());
            }
            fd.agg = agg_desc;
            fragment = std::make_shared<GpuFragment>(ctx, fd, joiners);
            ops.push_back(std::make_shared<GpuFragmentSinkOperator>(nullptr, 8, 8, 0, fragment, (size_t)1 << 18));
            aggregator = fragment->aggregator();
        }
        PipelineDriver probe_driver(ops);
        if (with_rf) {
            // the probe-side driver must park until every build operator has published its collector
            probe_driver.set_local_rf_holders(fact_scan_op->rf_holders());
            auto st = probe_driver.process(&state);
            if (!st.ok() || st.value() != PipelineDriver::PRECONDITION_BLOCK || !probe_driver.local_rf_block()) {
                fprintf(stderr, "probe driver was not blocked on its local runtime filters\n");
                return 2;
            }
            run_builds();
            if (probe_driver.local_rf_block()) {
                fprintf(stderr, "runtime filters were not published by the build operators\n");
                return 2;
            }
        }
        run_to_finish(probe_driver, &state, fused ? "probe(fused)" : with_rf ? "probe(per-operator + runtime filters)" : "probe(per-operator)");
        rows_after_scan[variant] = (size_t)fact_scan_op->rows_after_scan();
        // ---- result pipeline ----
        auto sink = std::make_shared<ResultSink>();
        auto source = std::make_shared<GpuAggregateBlockingSourceOperator>(nullptr, 9, 9, 0, aggregator);
        PipelineDriver result_driver({source, sink});
        run_to_finish(result_driver, &state, "result");
        results[variant] = collect(*sink);
        result_driver.close(&state);
        probe_driver.close(&state);
        printf("%s path: %zu groups, %zu rows left the scan, %zu rows moved between operators\n",
               fused ? "fused fragment" : with_rf ? "per-operator + runtime filters" : "per-operator", results[variant].size(), rows_after_scan[variant],
               probe_driver.rows_moved());
        for (auto& op : ops) printf("    %s | %s | %s\n", op->runtime_profile()->name().c_str(), op->common_metrics()->to_string().c_str(), op->unique_metrics()->to_string().c_str());
        ops.clear(); // scans go before the hub that owns their filters
    }
    int rc = 0;
    {
        // ---- HASH_PARTITIONED exchange sink: 8 channels on lo_custkey; every row must arrive once, on the channel
        // ReduceOp(fnv_hash(key), 8) names (hash_util.hpp:127-134,242-244), in input order per channel ----
        const int nch = 8;
        std::vector<int32_t> rowid(n_fact);
        for (size_t i = 0; i < n_fact; i++) rowid[i] = (int32_t)i;
        auto whole = make_chunk({{LO_CUSTKEY, lo_cust}, {LO_REVENUE, rowid}});
        sr_part_desc pd{};
        pd.hash_fn = SR_HASH_FNV;
        pd.reduce_op = SR_REDUCE_MULHI;
        pd.num_channels = nch;
        pd.num_part_slots = 1;
        pd.part_slots[0] = LO_CUSTKEY;
        std::vector<std::vector<ChunkPtr>> received(nch);
        auto sink = std::make_shared<GpuExchangeSinkOperator>(nullptr, 11, 11, 0, ctx, pd, [&](int32_t ch, const ChunkPtr& c) {
            received[ch].push_back(c);
            return Status::OK();
        }, (size_t)1 << 18);
        CHECK_OK(sink->prepare(&state));
        for (auto& c : split(whole, 4096)) CHECK_OK(sink->push_chunk(&state, c));
        CHECK_OK(sink->set_finishing(&state));
        size_t total = 0;
        bool ok = true;
        for (int ch = 0; ch < nch; ch++) {
            int32_t last = -1;
            for (auto& c : received[ch]) {
                auto* k = (const int32_t*)c->get_column_by_slot_id(LO_CUSTKEY)->raw_data();
                auto* id = (const int32_t*)c->get_column_by_slot_id(LO_REVENUE)->raw_data();
                for (size_t i = 0; i < c->num_rows(); i++) {
                    uint32_t h = 0x811C9DC5u; // FNV seed (HashUtil::FNV_SEED)
                    for (int b = 0; b < 4; b++) h = (((uint32_t)k[i] >> (8 * b) & 0xFFu) ^ h) * 0x01000193u;
                    const int32_t want = (int32_t)(((uint64_t)h * (uint64_t)nch) >> 32);
                    ok = ok && want == ch && id[i] > last && lo_cust[id[i]] == k[i];
                    last = id[i];
                    total++;
                }
            }
        }
        if (!ok || total != n_fact) {
            fprintf(stderr, "exchange sink: %zu of %zu rows arrived, placement/order %s\n", total, n_fact, ok ? "ok" : "WRONG");
            rc = 1;
        }
        printf("exchange sink: %zu rows over %d channels, placement and per-channel order verified\n", total, nch);
    }
    {
        // ---- two-phase aggregate: GROUP BY lo_custkey, SUM(lo_revenue), AVG(lo_supplycost), COUNT(*) ----
        //   first phase:  GpuScanOperator(lineorder) -> GpuAggregateStreamingSinkOperator | ...StreamingSourceOperator ->
        //   merge phase:  GpuAggregateBlockingSinkOperator(merge desc)                    | ...BlockingSourceOperator -> sink
        // in the three TStreamingPreaggregationMode settings; AUTO runs with a tiny table budget so that it flushes and
        // passes through.  lo_custkey has 30 000 values over 1 M rows sorted by nothing: reduction ~33 per full table.
        struct G {
            int64_t rev = 0, cost = 0, cnt = 0;
        };
        const char* names[4] = {"AUTO", "FORCE_STREAMING", "FORCE_PREAGGREGATION", "LIMITED_MEM"};
        // run 5's second key: lo_partkey folded to 13 values -> 390 000 groups over 1 M rows, a middling reduction (~1.4 on the
        // first batch) that sends AUTO into SELECTIVE_PREAGG
        std::vector<int32_t> lo_part13(n_fact);
        for (size_t i = 0; i < n_fact; i++) lo_part13[i] = lo_part[i] % 13;
        ChunkPtr lineorder13 = make_chunk({{LO_CUSTKEY, lo_cust}, {LO_PARTKEY, lo_part13}, {LO_REVENUE, lo_rev}, {LO_SUPPLYCOST, lo_cost}});
        // (a) GROUP BY lo_custkey: 30 000 groups over 1 M rows, reduction ~33 -> AUTO keeps pre-aggregating;
        // (b) GROUP BY lo_custkey, lo_partkey: ~1 M groups, reduction ~1 -> AUTO must flush, pass through and probe again
        // (c) LIMITED_MEM on the low-reduction input: streams as soon as the table reaches its byte limit;
        // (d) AUTO on the middling-reduction input: SELECTIVE_PREAGG (known groups aggregated, the rest streamed)
        for (int run = 0; run < 6; run++) {
            const int mode = run < 3 ? run : run == 4 ? 3 : 0;
            const bool two_keys = run >= 3;
            const bool folded = run == 5;
            const std::vector<int32_t>& key2 = folded ? lo_part13 : lo_part;
            sr_agg_desc q{};
            q.num_group_keys = two_keys ? 2 : 1;
            q.group_slots[0] = LO_CUSTKEY;
            q.group_slots[1] = LO_PARTKEY;
            q.group_types[0] = q.group_types[1] = SR_TYPE_INT;
            q.num_fns = 3;
            q.fns[0] = sr_agg_fn{SR_AGG_SUM, SR_TYPE_INT, OUT_REV, 0, col_expr(LO_REVENUE)};
            q.fns[1] = sr_agg_fn{SR_AGG_AVG, SR_TYPE_INT, OUT_COST, 0, col_expr(LO_SUPPLYCOST)};
            q.fns[2] = sr_agg_fn{SR_AGG_COUNT_STAR, SR_TYPE_INT, OUT_COST + 1, 0, sr_expr{}};
            sr_agg_desc p1{}, p2{};
            if (sr_agg_two_phase_descs(&q, &p1, &p2) != SR_OK) {
                fprintf(stderr, "sr_agg_two_phase_descs failed\n");
                return 2;
            }
            std::map<std::pair<int32_t, int32_t>, G> want;
            for (size_t i = 0; i < n_fact; i++) {
                G& g = want[{lo_cust[i], two_keys ? key2[i] : 0}];
                g.rev += lo_rev[i];
                g.cost += lo_cost[i];
                g.cnt++;
            }
            sr_scan_desc sd{};
            std::vector<int32_t> outs = {LO_CUSTKEY, LO_PARTKEY, LO_REVENUE, LO_SUPPLYCOST};
            sd.out_slots = outs.data();
            sd.num_out_slots = (int32_t)outs.size();
            GpuScanOperatorFactory scan_f(30, 30, ctx, sd, {split(folded ? lineorder13 : lineorder, 4096)});
            auto streaming_f = std::make_shared<GpuStreamingAggregatorFactory>(ctx, p1, (GpuStreamingPreaggMode)mode, /*max_ht_bytes=*/(size_t)256 << 10,
                                                                               /*pass_through_batches=*/1);
            auto streaming = streaming_f->get_or_create(0);
            auto final_f = std::make_shared<GpuAggregatorFactory>(ctx, p2);
            GpuAggregateStreamingSinkOperatorFactory stream_sink_f(31, 31, streaming_f);
            GpuAggregateStreamingSourceOperatorFactory stream_source_f(32, 32, streaming_f);
            PipelineDriver first({scan_f.create(1, 0), stream_sink_f.create(1, 0)});
            GpuAggregateBlockingSinkOperatorFactory merge_sink_f(33, 33, final_f);
            PipelineDriver second({stream_source_f.create(1, 0), merge_sink_f.create(1, 0)});
            run_to_finish(first, &state, "first phase");
            run_to_finish(second, &state, "merge phase");
            auto sink = std::make_shared<ResultSink>();
            PipelineDriver result_driver({std::make_shared<GpuAggregateBlockingSourceOperator>(nullptr, 34, 34, 0, final_f->get_or_create(0)), sink});
            run_to_finish(result_driver, &state, "two-phase result");
            size_t groups = 0;
            bool ok = true;
            for (auto& c : sink->chunks) {
                auto* k = (const int32_t*)c->get_column_by_slot_id(LO_CUSTKEY)->raw_data();
                auto* k2 = two_keys ? (const int32_t*)c->get_column_by_slot_id(LO_PARTKEY)->raw_data() : nullptr;
                auto* r = (const int64_t*)c->get_column_by_slot_id(OUT_REV)->raw_data();
                auto* a = (const double*)c->get_column_by_slot_id(OUT_COST)->raw_data();
                auto* n = (const int64_t*)c->get_column_by_slot_id(OUT_COST + 1)->raw_data();
                for (size_t i = 0; i < c->num_rows(); i++, groups++) {
                    auto it = want.find({k[i], k2 ? k2[i] : 0});
                    const double avg = it == want.end() ? 0 : (double)it->second.cost / (double)it->second.cnt;
                    ok = ok && it != want.end() && r[i] == it->second.rev && n[i] == it->second.cnt && std::abs(a[i] - avg) <= 1e-9 * avg;
                }
            }
            if (!ok || groups != want.size()) {
                fprintf(stderr, "two-phase aggregate (%s): %zu groups (want %zu), values %s\n", names[mode], groups, want.size(), ok ? "ok" : "WRONG");
                rc = 1;
            }
            if (run == 3 && (streaming->rows_streamed() == 0 || streaming->num_flushes() == 0)) {
                fprintf(stderr, "AUTO never left pre-aggregation on a low-reduction input\n");
                rc = 1;
            }
            if (run == 4 && (!streaming->memory_limited() || streaming->rows_streamed() == 0)) {
                fprintf(stderr, "LIMITED_MEM never hit its limit / never streamed\n");
                rc = 1;
            }
            if (run == 5 && (streaming->num_selective_phases() == 0 || streaming->rows_selected_into_table() == 0 || streaming->rows_streamed() == 0)) {
                fprintf(stderr, "AUTO never took the SELECTIVE_PREAGG leg on a middling-reduction input (%d phases, %zu rows into the table, %zu streamed)\n",
                        streaming->num_selective_phases(), streaming->rows_selected_into_table(), streaming->rows_streamed());
                rc = 1;
            }
            printf("two-phase aggregate %-20s %s: %zu groups; first phase passed %zu rows through, emitted %zu pre-aggregated rows, %d table flushes, "
                   "%d selective phases (%zu rows aggregated into known groups)\n",
                   names[mode], folded ? "(2 keys, reduction ~2.5)" : two_keys ? "(2 keys, reduction ~1)" : "(1 key, reduction ~33)", groups,
                   streaming->rows_streamed(), streaming->rows_from_table(), streaming->num_flushes(), streaming->num_selective_phases(),
                   streaming->rows_selected_into_table());
            result_driver.close(&state);
            second.close(&state);
            first.close(&state);
        }
    }
    if (results[0] != oracle_groups) {
        fprintf(stderr, "per-operator pipeline differs from the oracle (%zu vs %zu groups)\n", results[0].size(), expect.size());
        rc = 1;
    }
    if (results[2] != oracle_groups) {
        fprintf(stderr, "per-operator pipeline with runtime filters differs from the oracle (%zu vs %zu groups)\n", results[2].size(), expect.size());
        rc = 1;
    }
    {
        // the filters are exact enough to drop (almost) every row the joins would drop: the scan must emit at least the
        // joined rows and far fewer than the table (bloom false positives only)
        size_t joined = 0;
        for (size_t i = 0; i < n_fact; i++)
            joined += s_region[lo_supp[i] - 1] == 1 && c_region[lo_cust[i] - 1] == 1 && p_mfgr[lo_part[i] - 1] <= 1;
        if (rows_after_scan[2] < joined || rows_after_scan[2] > joined + n_fact / 50 || rows_after_scan[0] != n_fact) {
            fprintf(stderr, "runtime filters: %zu rows left the scan, %zu join (of %zu)\n", rows_after_scan[2], joined, n_fact);
            rc = 1;
        }
        printf("runtime filters: scan output %zu -> %zu rows (%zu rows join)\n", rows_after_scan[0], rows_after_scan[2], joined);
    }
    if (results[1] != oracle_groups) {
        fprintf(stderr, "fused pipeline differs from the oracle (%zu vs %zu groups)\n", results[1].size(), expect.size());
        rc = 1;
    }
    printf("kernel launches: %lld\n", (long long)sr_ctx_kernel_launches(ctx));
    sr_ctx_destroy(ctx);
    printf(rc == 0 ? "PIPELINE_Q41_OK\n" : "PIPELINE_Q41_MISMATCH\n");
    return rc;
}
