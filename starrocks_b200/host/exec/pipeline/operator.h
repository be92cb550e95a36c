// API-identical shim of the pipeline operator interface so the GPU operators can be exercised the way the BE's
// PipelineDriver exercises them.  Names, signatures and contracts follow
//   be/src/exec/pipeline/operator.h:44-352   (Operator), :354-456 (OperatorFactory)
//   be/src/exec/pipeline/source_operator.h:37-189 (SourceOperator / SourceOperatorFactory)
//   be/src/exec/pipeline/operator_with_dependency.h:41-49
//   be/src/base/status.h, statusor.h (Status / StatusOr, RETURN_IF_ERROR)
// In a real BE build these headers are the BE's own; only this directory's gpu/*.h would be added.
#pragma once

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../column/chunk.h"

namespace starrocks {

class Status {
public:
    Status() = default;
    static Status OK() { return Status(); }
    static Status InternalError(std::string m) { return Status(1, std::move(m)); }
    static Status InvalidArgument(std::string m) { return Status(2, std::move(m)); }
    static Status NotSupported(std::string m) { return Status(3, std::move(m)); }
    static Status MemoryLimitExceeded(std::string m) { return Status(4, std::move(m)); }
    static Status EndOfFile(std::string m) { return Status(5, std::move(m)); }
    bool ok() const { return _code == 0; }
    bool is_end_of_file() const { return _code == 5; }
    int code() const { return _code; }
    const std::string& message() const { return _msg; }
    std::string to_string() const { return ok() ? "OK" : _msg; }

private:
    Status(int c, std::string m) : _code(c), _msg(std::move(m)) {}
    int _code = 0;
    std::string _msg;
};

template <typename T>
class StatusOr {
public:
    StatusOr(Status s) : _status(std::move(s)) {}
    StatusOr(T v) : _value(std::move(v)) {}
    bool ok() const { return _status.ok(); }
    const Status& status() const { return _status; }
    T& value() { return _value; }
    T& operator*() { return _value; }

private:
    Status _status;
    T _value{};
};

#define RETURN_IF_ERROR(stmt)          \
    do {                               \
        Status _s = (stmt);            \
        if (!_s.ok()) return _s;       \
    } while (0)

// RuntimeState: only what the operators read (be/src/runtime/runtime_state.h:130 chunk_size = batch_size,
// default vector_chunk_size 4096, be/src/common/config.h:915)
class RuntimeState {
public:
    explicit RuntimeState(int chunk_size = 4096) : _chunk_size(chunk_size) {}
    int chunk_size() const { return _chunk_size; }
    bool is_cancelled() const { return _cancelled; }
    void set_cancelled() { _cancelled = true; }

private:
    int _chunk_size;
    bool _cancelled = false;
};

// RuntimeProfile, the part the operators use (be/src/common/runtime_profile.h): named counters with a unit, created
// once (add_counter) and bumped with COUNTER_UPDATE / SCOPED_TIMER.  Operator owns two of them, like the reference
// (be/src/exec/pipeline/operator.h:296-335): _common_metrics (filled by the driver: rows pushed / pulled, operator
// time) and _unique_metrics (the operator's own: BuildHashTableTime, SearchHashTableTime, AggComputeTime, ...).
enum class TUnit { UNIT, TIME_NS, BYTES };
class RuntimeProfile {
public:
    struct Counter {
        TUnit unit;
        std::atomic<int64_t> v{0};
        void update(int64_t d) { v.fetch_add(d, std::memory_order_relaxed); }
        void set(int64_t x) { v.store(x, std::memory_order_relaxed); }
        int64_t value() const { return v.load(std::memory_order_relaxed); }
    };
    explicit RuntimeProfile(std::string name) : _name(std::move(name)) {}
    Counter* add_counter(const std::string& name, TUnit unit) {
        auto& c = _counters[name];
        if (!c) {
            c = std::make_unique<Counter>();
            c->unit = unit;
        }
        return c.get();
    }
    Counter* get_counter(const std::string& name) const {
        auto it = _counters.find(name);
        return it == _counters.end() ? nullptr : it->second.get();
    }
    const std::string& name() const { return _name; }
    std::string to_string() const {
        std::string out = _name + ":";
        for (auto& kv : _counters) {
            const int64_t v = kv.second->value();
            out += " " + kv.first + "=";
            if (kv.second->unit == TUnit::TIME_NS)
                out += std::to_string(v / 1000) + "us";
            else
                out += std::to_string(v);
        }
        return out;
    }

private:
    std::string _name;
    std::map<std::string, std::unique_ptr<Counter>> _counters;
};
#define ADD_COUNTER(profile, name, unit) (profile)->add_counter(name, unit)
#define ADD_TIMER(profile, name) (profile)->add_counter(name, TUnit::TIME_NS)
#define COUNTER_UPDATE(c, v) (c)->update(v)
#define COUNTER_SET(c, v) (c)->set(v)
class ScopedTimer {
public:
    explicit ScopedTimer(RuntimeProfile::Counter* c) : _c(c), _t0(std::chrono::steady_clock::now()) {}
    ~ScopedTimer() { _c->update(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - _t0).count()); }

private:
    RuntimeProfile::Counter* _c;
    std::chrono::steady_clock::time_point _t0;
};
#define SR_CONCAT_(a, b) a##b
#define SR_CONCAT(a, b) SR_CONCAT_(a, b)
#define SCOPED_TIMER(c) ScopedTimer SR_CONCAT(_scoped_timer_, __LINE__)(c)

// MemTracker, the part an operator touches (be/src/runtime/mem_tracker.h: consume / release / consumption)
class MemTracker {
public:
    void consume(int64_t b) { _bytes.fetch_add(b, std::memory_order_relaxed); }
    void release(int64_t b) { _bytes.fetch_sub(b, std::memory_order_relaxed); }
    void set(int64_t b) { _bytes.store(b, std::memory_order_relaxed); }
    int64_t consumption() const { return _bytes.load(std::memory_order_relaxed); }

private:
    std::atomic<int64_t> _bytes{0};
};

namespace pipeline {

class OperatorFactory;

class Operator {
public:
    Operator(OperatorFactory* factory, int32_t id, std::string name, int32_t plan_node_id, bool is_subordinate, int32_t driver_sequence)
            : _factory(factory), _id(id), _name(std::move(name)), _plan_node_id(plan_node_id), _driver_sequence(driver_sequence) {
        (void)is_subordinate;
        const std::string label = _name + " (plan_node_id=" + std::to_string(_plan_node_id) + ")";
        _runtime_profile = std::make_shared<RuntimeProfile>(label);
        _common_metrics = std::make_shared<RuntimeProfile>("CommonMetrics");
        _unique_metrics = std::make_shared<RuntimeProfile>("UniqueMetrics");
        _mem_tracker = std::make_shared<MemTracker>();
        _push_chunk_num_counter = ADD_COUNTER(_common_metrics.get(), "PushChunkNum", TUnit::UNIT);
        _push_row_num_counter = ADD_COUNTER(_common_metrics.get(), "PushRowNum", TUnit::UNIT);
        _pull_chunk_num_counter = ADD_COUNTER(_common_metrics.get(), "PullChunkNum", TUnit::UNIT);
        _pull_row_num_counter = ADD_COUNTER(_common_metrics.get(), "PullRowNum", TUnit::UNIT);
        _total_timer = ADD_TIMER(_common_metrics.get(), "OperatorTotalTime");
    }
    virtual ~Operator() = default;

    virtual Status prepare(RuntimeState* state) { return Status::OK(); }
    // be/src/exec/pipeline/operator.h:60-66: per-driver state that must be created on the driver's own thread
    virtual Status prepare_local_state(RuntimeState* state) { return Status::OK(); }
    // be/src/exec/pipeline/operator.h:143: make the operator reusable (multi-cast / cache operators call it)
    virtual Status reset_state(RuntimeState* state, const std::vector<ChunkPtr>& refill_chunks) { return Status::OK(); }
    virtual Status set_finishing(RuntimeState* state) { return Status::OK(); }
    virtual Status set_finished(RuntimeState* state) { return Status::OK(); }
    virtual Status set_cancelled(RuntimeState* state) { return Status::OK(); }
    virtual void close(RuntimeState* state) {}

    virtual bool has_output() const = 0;
    virtual bool need_input() const = 0;
    virtual bool is_finished() const = 0;
    virtual bool pending_finish() const { return false; }

    virtual StatusOr<ChunkPtr> pull_chunk(RuntimeState* state) = 0;
    virtual Status push_chunk(RuntimeState* state, const ChunkPtr& chunk) = 0;

    int32_t get_id() const { return _id; }
    int32_t get_plan_node_id() const { return _plan_node_id; }
    RuntimeProfile* runtime_profile() { return _runtime_profile.get(); }
    RuntimeProfile* common_metrics() { return _common_metrics.get(); }
    RuntimeProfile* unique_metrics() { return _unique_metrics.get(); }
    MemTracker* mem_tracker() const { return _mem_tracker.get(); }
    // PipelineDriver is the one filling the common metrics (pipeline_driver.cpp:340-420)
    void update_push_metrics(int64_t rows) {
        COUNTER_UPDATE(_push_chunk_num_counter, 1);
        COUNTER_UPDATE(_push_row_num_counter, rows);
    }
    void update_pull_metrics(int64_t rows) {
        COUNTER_UPDATE(_pull_chunk_num_counter, 1);
        COUNTER_UPDATE(_pull_row_num_counter, rows);
    }
    RuntimeProfile::Counter* total_timer() { return _total_timer; }
    std::string get_raw_name() const { return _name; }
    std::string get_name() const { return _name + "_" + std::to_string(_plan_node_id) + (is_finished() ? "(X)" : "(O)"); }

protected:
    OperatorFactory* _factory;
    const int32_t _id;
    const std::string _name;
    const int32_t _plan_node_id;
    const int32_t _driver_sequence;
    std::shared_ptr<RuntimeProfile> _runtime_profile, _common_metrics, _unique_metrics;
    std::shared_ptr<MemTracker> _mem_tracker;
    RuntimeProfile::Counter *_push_chunk_num_counter, *_push_row_num_counter, *_pull_chunk_num_counter, *_pull_row_num_counter, *_total_timer;
};
using OperatorPtr = std::shared_ptr<Operator>;
using Operators = std::vector<OperatorPtr>;

class SourceOperator : public Operator {
public:
    using Operator::Operator;
    bool need_input() const override { return false; }
    Status push_chunk(RuntimeState* state, const ChunkPtr& chunk) override { return Status::InternalError("Shouldn't push chunk to source operator"); }
};

// be/src/exec/pipeline/operator_with_dependency.h:41-49
class OperatorWithDependency : public Operator {
public:
    using Operator::Operator;
    virtual bool is_ready() const = 0;
};

class OperatorFactory {
public:
    OperatorFactory(int32_t id, std::string name, int32_t plan_node_id) : _id(id), _name(std::move(name)), _plan_node_id(plan_node_id) {}
    virtual ~OperatorFactory() = default;
    virtual OperatorPtr create(int32_t degree_of_parallelism, int32_t driver_sequence) = 0;
    virtual bool is_source() const { return false; }
    int32_t id() const { return _id; }
    int32_t plan_node_id() const { return _plan_node_id; }
    virtual Status prepare(RuntimeState* state) { return Status::OK(); }
    virtual void close(RuntimeState* state) {}
    std::string get_name() const { return _name + "_(" + std::to_string(_plan_node_id) + ")"; }
    virtual bool support_event_scheduler() const { return false; }

protected:
    const int32_t _id;
    const std::string _name;
    const int32_t _plan_node_id;
};
using OpFactoryPtr = std::shared_ptr<OperatorFactory>;
using OpFactories = std::vector<OpFactoryPtr>;

class SourceOperatorFactory : public OperatorFactory {
public:
    using OperatorFactory::OperatorFactory;
    bool is_source() const override { return true; }
    void set_degree_of_parallelism(size_t dop) { _degree_of_parallelism = dop; }
    size_t degree_of_parallelism() const { return _degree_of_parallelism; }

protected:
    size_t _degree_of_parallelism = 1;
};

// be/src/exec/pipeline/runtime_filter_types.h:140-245.  The build operator of a join publishes its filters once
// (set_collector, first caller wins); drivers whose source consumes them stay PRECONDITION_BLOCKed until every holder
// is ready (pipeline_driver.h:357-364 local_rf_block).  The collector type is the execution layer's business.
class RuntimeFilterHolder {
public:
    void set_collector(std::shared_ptr<void> collector) {
        void* expected = nullptr;
        if (_collector.compare_exchange_strong(expected, collector.get(), std::memory_order_release, std::memory_order_acquire))
            _collector_ownership = std::move(collector);
    }
    void* get_collector() const { return _collector.load(std::memory_order_acquire); }
    bool is_ready() const { return get_collector() != nullptr; }

private:
    std::shared_ptr<void> _collector_ownership;
    std::atomic<void*> _collector{nullptr};
};

class RuntimeFilterHub { // keyed by the plan node id of the join that produces the filters (factory level: sequence -1)
public:
    void add_holder(int32_t plan_node_id) { _holders.emplace(plan_node_id, std::make_unique<RuntimeFilterHolder>()); }
    void set_collector(int32_t plan_node_id, std::shared_ptr<void> collector) { get_holder(plan_node_id)->set_collector(std::move(collector)); }
    RuntimeFilterHolder* get_holder(int32_t plan_node_id) {
        auto it = _holders.find(plan_node_id);
        return it == _holders.end() ? nullptr : it->second.get();
    }
    std::vector<RuntimeFilterHolder*> gather_holders(const std::set<int32_t>& ids) {
        std::vector<RuntimeFilterHolder*> out;
        for (int32_t id : ids)
            if (auto* h = get_holder(id)) out.push_back(h);
        return out;
    }

private:
    std::unordered_map<int32_t, std::unique_ptr<RuntimeFilterHolder>> _holders;
};

// The pull/push loop of PipelineDriver::process (be/src/exec/pipeline/pipeline_driver.cpp:270-500) for one driver:
// for each adjacent pair, pull when the upstream has output and the downstream needs input, propagate finishing,
// enforce the chunk_size limit (:372-378).  Returns when the sink is finished or no operator can make progress
// (the real driver would park in the poller; tests call process() again after the dependency is released).
class PipelineDriver {
public:
    explicit PipelineDriver(Operators ops) : _operators(std::move(ops)), _finishing_sent(_operators.size(), false) {}
    enum State { READY, PRECONDITION_BLOCK, PENDING_FINISH, FINISH };

    Status prepare(RuntimeState* state) {
        for (auto& op : _operators) RETURN_IF_ERROR(op->prepare(state));
        for (auto& op : _operators) RETURN_IF_ERROR(op->prepare_local_state(state));
        return Status::OK();
    }

    void set_local_rf_holders(std::vector<RuntimeFilterHolder*> holders) { _local_rf_holders = std::move(holders); }
    bool local_rf_block() { // pipeline_driver.h:357-364
        if (_all_local_rf_ready) return false;
        _all_local_rf_ready = std::all_of(_local_rf_holders.begin(), _local_rf_holders.end(), [](auto* h) { return h->is_ready(); });
        return !_all_local_rf_ready;
    }

    StatusOr<State> process(RuntimeState* state) {
        const size_t n = _operators.size();
        if (local_rf_block()) return PRECONDITION_BLOCK;
        for (auto& op : _operators) {
            auto* dep = dynamic_cast<OperatorWithDependency*>(op.get());
            if (dep != nullptr && !dep->is_ready()) return PRECONDITION_BLOCK;
        }
        while (true) {
            bool progressed = false;
            for (size_t i = _first_unfinished; i + 1 < n; i++) {
                auto& curr = _operators[i];
                auto& next = _operators[i + 1];
                if (curr->has_output() && next->need_input() && !next->is_finished()) {
                    StatusOr<ChunkPtr> maybe = [&]() {
                        SCOPED_TIMER(curr->total_timer());
                        return curr->pull_chunk(state);
                    }();
                    if (!maybe.ok() && !maybe.status().is_end_of_file()) return maybe.status();
                    if (maybe.ok() && maybe.value() != nullptr && maybe.value()->num_rows() > 0) {
                        if ((int)maybe.value()->num_rows() > state->chunk_size())
                            return Status::InternalError("Intermediate chunk size must not be greater than " + std::to_string(state->chunk_size()) +
                                                         ", actually " + std::to_string(maybe.value()->num_rows()) + " after " + curr->get_name());
                        curr->update_pull_metrics((int64_t)maybe.value()->num_rows());
                        next->update_push_metrics((int64_t)maybe.value()->num_rows());
                        {
                            SCOPED_TIMER(next->total_timer());
                            RETURN_IF_ERROR(next->push_chunk(state, maybe.value()));
                        }
                        _rows_moved += maybe.value()->num_rows();
                    }
                    progressed = true;
                }
                if (curr->is_finished() && !_finishing_sent[i + 1]) { // :424-437
                    RETURN_IF_ERROR(next->set_finishing(state));
                    _finishing_sent[i + 1] = true;
                    progressed = true;
                }
            }
            while (_first_unfinished + 1 < n && _operators[_first_unfinished]->is_finished() && _finishing_sent[_first_unfinished + 1])
                _first_unfinished++;
            if (_operators.back()->is_finished()) {
                // PENDING_FINISH (pipeline_driver.cpp:880-905): asynchronous work queued by an operator must drain first
                for (auto& op : _operators)
                    if (op->pending_finish()) return PENDING_FINISH;
                for (auto& op : _operators) RETURN_IF_ERROR(op->set_finished(state));
                return FINISH;
            }
            if (!progressed) return READY;
        }
    }

    void close(RuntimeState* state) {
        for (auto& op : _operators) op->close(state);
    }
    size_t rows_moved() const { return _rows_moved; }

private:
    Operators _operators;
    std::vector<bool> _finishing_sent;
    size_t _first_unfinished = 0;
    size_t _rows_moved = 0;
    std::vector<RuntimeFilterHolder*> _local_rf_holders;
    bool _all_local_rf_ready = false;
};

} // namespace pipeline
} // namespace starrocks
